#!/usr/bin/env python
"""bench.py — SBS frames/s of the B200 hot path (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one batch of `--batch` synthetic 1080p frames
(config "1080p 300-frame synthetic clip, default config.json stereo params", BASELINE.json
configs[1]); frames are independent, so with N GPUs every rank processes its own batch (frame-range
sharding, no collective; "scaling": "weak") and the reported value is total frames / max-over-ranks
device time.  The K timed steps are K consecutive batches of ONE continuous clip: they overlap in the
slot pipeline the way consecutive seconds of a video do, and the timed region is bracketed by a
barrier + full drain + synchronize on both sides (pipeline fill and drain are inside it).

  value  frames/s with the inputs already resident in HBM (vsc_submit_device), timed with CUDA
         events across all slot streams (vsc_timer_begin/end)
  e2e    frames/s through the public StereoGenerator submit/collect API with pinned HOST buffers:
         H2D of every frame and D2H of every SBS result are inside the timed region
  roofline      dominant kernel: algorithmic bytes per frame / its mean device
                time (CUDA events on its stream), against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the oracle port of the reference's CPU path timed on this box's host cores
--impl reference times that CPU port alone (the reference itself is Python + OpenCV + torch and
cannot travel to the GPU box; oracle/ is its restatement).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')   # one hardware queue per frame slot (before CUDA init)
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(ROOT, 'video-stereo-converter_b200')]

import numpy as np  # noqa: E402

H, W = 1080, 1920
DEPTH_DTYPE = np.uint8
WORKLOAD = '1080p (1920x1080) synthetic clip, uint8 depth, default config.json stereo params'
N_DISTINCT = 48          # distinct synthetic frames cycled through (inputs 48 * 8.3 MB > 126 MB L2)
METRIC = 'SBS frames/sec (1080p, default stereo params)'
if os.environ.get('VSC_BENCH_WORKLOAD') == '4k':    # side measurement (BASELINE.json configs[2]), never the headline line
    H, W, DEPTH_DTYPE, N_DISTINCT = 2160, 3840, np.uint16, 16
    WORKLOAD = '4K (3840x2160) synthetic clip, uint16 depth, default config.json stereo params'
    METRIC = 'SBS frames/sec (4K, 16-bit depth, default stereo params)'


def algorithmic_bytes(h, w, depth_itemsize):
    """rgb in + depth in + SBS out, per frame."""
    return h * w * 3 + h * w * depth_itemsize + h * 2 * w * 3


def hbm_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index: int):
        self.index, self.samples, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        if os.environ.get('VSC_BENCH_NO_CLOCKS'):
            return
        while not self._stop.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s and s[0].replace('.', '').isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for s in self.samples for i in range(4) if len(s) > 2 + i and s[2 + i].lower() == 'active'})
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(self.samples)}


def make_frames(n, h, w, dtype, seed0=0):
    from vsc_b200.synthetic import make_pair
    return [make_pair(h, w, seed0 + i, dtype) for i in range(n)]


# ------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    from vsc_b200 import StereoGenerator, StereoParams
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    params = StereoParams()
    slots, batch, group = args.slots, args.batch, args.group
    gen = StereoGenerator(f'cuda:{local_rank}', n_slots=slots, group_size=group)
    # distinct frames per rank (frame-range shard of the synthetic clip)
    n_distinct = min(N_DISTINCT, max(batch, slots))
    frames = make_frames(n_distinct, H, W, DEPTH_DTYPE, seed0=rank * 1000)
    d_rgb = [torch.from_numpy(r).cuda() for r, _ in frames]
    d_dep = [torch.from_numpy(d).cuda() for _, d in frames]
    d_out = [[torch.empty((H, 2 * W, 3), dtype=torch.uint8, device='cuda') for _ in range(group)] for _ in range(slots)]
    torch.cuda.synchronize()

    # The clip is one continuous stream of frames: step k is frames [k*batch, (k+1)*batch) and consecutive steps
    # overlap in the slot pipeline exactly as consecutive seconds of a video do.  The timed region is bracketed by
    # a full drain + synchronize on both sides (timed()), never between steps.
    class Pipe:
        def __init__(self):
            self.free, self.busy, self.last = list(range(slots)), [], None

    def device_step(step, pipe):
        """one batch, inputs resident in HBM; a finished slot is reused at once (no head-of-line blocking);
        every submission carries `group` frames that share the slot's stream and one hole-filling launch"""
        for i0 in range(0, batch, group):
            if not pipe.free:
                s = gen.wait_any(pipe.busy)
                gen.wait(s)
                pipe.busy.remove(s)
                pipe.free.append(s)
            s = pipe.free.pop(0)
            tri = []
            for k in range(min(group, batch - i0)):
                f = (step * batch + i0 + k) % n_distinct
                tri.append((d_rgb[f].data_ptr(), d_dep[f].data_ptr(), d_out[s][k].data_ptr()))
            gen.submit_device_group(s, tri, DEPTH_DTYPE, H, W, params)
            pipe.busy.append(s)

    def device_drain(pipe):
        for s in pipe.busy:
            gen.wait(s)
        pipe.free += pipe.busy
        pipe.busy = []

    from concurrent.futures import ThreadPoolExecutor
    # loader threads of this rank: never more than its share of the host cores (N ranks share the box)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    loaders = ThreadPoolExecutor(max_workers=max(1, min(group, cores // max(1, world) - 1)))
    for s_ in range(slots):          # allocate the pinned staging buffers outside the timed region
        for k_ in range(group):
            gen.pinned_inputs(s_, H, W, DEPTH_DTYPE, k_)
    e2e_reads = []

    def e2e_step(step, pipe):
        for i0 in range(0, batch, group):
            if not pipe.free:
                s = gen.wait_any(pipe.busy)
                pipe.last = gen.collect(s, copy=False)
                pipe.busy.remove(s)
                pipe.free.append(s)
            s = pipe.free.pop(0)
            n = min(group, batch - i0)

            def load(k, s=s, i0=i0):
                # the loader pool's job: decode straight into the slot's pinned buffers (here: memcpy of a prepared frame)
                f = (step * batch + i0 + k) % n_distinct
                prgb, pdep = gen.pinned_inputs(s, H, W, DEPTH_DTYPE, k)
                np.copyto(prgb, frames[f][0])
                np.copyto(pdep, frames[f][1])
            list(loaders.map(load, range(n)))
            gen.submit_pinned(s, params, n)
            pipe.busy.append(s)
        if pipe.last is not None:        # host read of the newest finished SBS frame (pinned, already transferred)
            last = pipe.last[-1] if isinstance(pipe.last, list) else pipe.last
            e2e_reads.append(int(last[0, 0, 0]))

    def e2e_drain(pipe):
        for s in pipe.busy:
            pipe.last = gen.collect(s, copy=False)
        pipe.free += pipe.busy
        pipe.busy = []
        last = pipe.last[-1] if isinstance(pipe.last, list) else pipe.last
        e2e_reads.append(int(last[0, 0, 0]))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, drain, steps):
        pipe = Pipe()
        barrier()
        gen.timer_begin()
        t0 = time.perf_counter()
        for k in range(steps):
            fn(k, pipe)
        drain(pipe)
        ms = gen.timer_end()
        wall = (time.perf_counter() - t0) * 1e3
        barrier()
        if dist is not None:
            t = torch.tensor([ms, wall], device='cuda', dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1])
        return ms, wall

    wp = Pipe()
    for k in range(args.warmup):
        device_step(k, wp)
    device_drain(wp)
    launches_per_frame = gen.last_frame_launches(0) / group
    with ClockSampler(local_rank) as clk:
        ms_dev, wall_dev = timed(device_step, device_drain, args.steps)
    wp = Pipe()
    for k in range(max(1, args.warmup // 2)):
        e2e_step(k, wp)
    e2e_drain(wp)
    ms_e2e, wall_e2e = timed(e2e_step, e2e_drain, args.steps)

    # per-kernel times for the roofline (separate, untimed pass with event pairs around every launch)
    kt = {}          # kernel name -> list over profiled frames of its summed device ms in that frame
    if rank == 0:
        gen.set_profiling(True)
        # one slot submission (`group` frames sharing one hole-filling launch sequence) in flight at a time, exactly
        # the launch structure of the timed legs; times are divided by the frames per submission
        nprof = 6
        for i in range(nprof):
            tri = [(d_rgb[(i * group + k) % n_distinct].data_ptr(), d_dep[(i * group + k) % n_distinct].data_ptr(),
                    d_out[0][k].data_ptr()) for k in range(group)]
            gen.submit_device_group(0, tri, DEPTH_DTYPE, H, W, params)
            gen.wait(0)
            if i >= 2:
                per = {}
                for name, t in gen.kernel_times(0):
                    per[name] = per.get(name, 0.0) + t / group
                for name, t in per.items():
                    kt.setdefault(name, []).append(t)
        gen.set_profiling(False)
    frames_total = batch * args.steps * world
    out = None
    if rank == 0:
        peak, peak_src = hbm_peak()
        per_kernel = {k: float(np.mean(v)) for k, v in kt.items()}   # device ms per frame, one slot submission in flight
        # launches of a kernel per slot submission: the hole-filling kernels run once for all `group` frames
        per_launch = {k: per_kernel[k] * (group if k.startswith('telea_') else 1) for k in per_kernel}
        frame_serial = float(sum(per_kernel.values()))
        dom = max(per_kernel, key=per_kernel.get)
        bytes_frame = algorithmic_bytes(H, W, np.dtype(DEPTH_DTYPE).itemsize)
        achieved = bytes_frame / (per_kernel[dom] * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, 'profiles', 'dominant_kernel_traffic.json')) as f:
                tj = json.load(f)
                if tj.get('kernel') == dom:
                    traffic = tj.get('dram_bytes_per_launch')
        except Exception:
            pass
        out = {
            'metric': METRIC, 'value': frames_total / (ms_dev * 1e-3), 'unit': 'frames/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_dev / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8/f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'frames_per_step_per_gpu': batch, 'slots_in_flight': slots, 'frames_per_slot': group,
                       'distinct_frames': n_distinct, 'l2_policy': 'inputs larger than L2 (%d distinct frames cycled)' % n_distinct,
                       'sharding': 'frame-range, no collective'},
            'e2e': {'value': frames_total / (ms_e2e * 1e-3), 'unit': 'frames/s',
                    'h2d_bytes_per_step': batch * (H * W * 3 + H * W * np.dtype(DEPTH_DTYPE).itemsize),
                    'd2h_bytes_per_step': batch * H * 2 * W * 3, 'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': int(round(launches_per_frame * batch * args.steps)),
            'launches_per_frame': launches_per_frame,
            'clocks': clk.summary(),
            'roofline': {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
                         'algorithmic_bytes_per_launch': bytes_frame * (group if dom.startswith('telea_') else 1),
                         'frames_per_launch': group if dom.startswith('telea_') else 1,
                         'kernel_ms_per_launch': per_launch[dom], 'kernel_ms_per_frame': per_kernel[dom],
                         'share_of_serial_frame': per_kernel[dom] / frame_serial if frame_serial else None,
                         'whole_path': {'achieved': frames_total / world / (ms_dev * 1e-3) * bytes_frame / 1e9,
                                        'frac': frames_total / world / (ms_dev * 1e-3) * bytes_frame / 1e9 / peak}},
            'kernel_ms_per_frame': {k: round(v, 4) for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1])},
            'wall_ms': {'device_leg': wall_dev, 'e2e_leg': wall_e2e},
        }
        if world == 1 and not args.no_cpu_baseline:
            out['cpu_baseline'] = cpu_baseline(sample_frames=args.cpu_frames)
    gen.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return out


def _use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU baseline is meant to use every host core."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    os.environ['OMP_NUM_THREADS'] = str(n)


def cpu_baseline(sample_frames=1, h=H, w=W, warm=False):
    """Oracle port of the reference CPU path on this box's host cores (bounded sample)."""
    _use_all_host_threads()
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import oracle as O
    O.build()
    frames = make_frames(sample_frames, h, w, DEPTH_DTYPE, seed0=0)
    if warm:
        O.process_frame(frames[0][0][:64], frames[0][1][:64], O.Params())
    t0 = time.perf_counter()
    for r, d in frames:
        O.process_frame(r, d, O.Params())
    dt = time.perf_counter() - t0
    return {'value': sample_frames * (h / H) / dt, 'unit': 'frames/s', 'cores': O.num_threads(), 'kind': 'port',
            'sample': f'{sample_frames} synthetic {w}x{h} frame(s), default params, oracle/ C+numpy port of '
                      f'helper/stereo_core.py with OpenMP ({O.num_threads()} threads); {dt:.1f} s',
            'host_cpus': os.cpu_count()}


def run_reference(args, rank, world):
    """--impl reference: the CPU port of the reference's own implementation, rank 0 only."""
    if rank != 0:
        return None
    _use_all_host_threads()
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import oracle as O
    O.build()
    total = args.steps + args.warmup
    rows = H if total <= 16 else max(136, (H * 16 // total) // 8 * 8)
    if os.environ.get('VSC_BENCH_TINY'):     # unit tests only: keep the CPU suite fast
        rows = 64
    frames = make_frames(2, rows, W, DEPTH_DTYPE, seed0=0)
    for k in range(args.warmup):
        O.process_frame(*frames[k % 2], O.Params())
    t0 = time.perf_counter()
    for k in range(args.steps):
        O.process_frame(*frames[k % 2], O.Params())
    dt = time.perf_counter() - t0
    value = args.steps * (rows / H) / dt
    base = {'value': value, 'unit': 'frames/s', 'cores': O.num_threads(), 'kind': 'port',
            'sample': f'each step = one synthetic {W}x{rows} frame ({rows}/{H} of a 1080p frame), default params'}
    return {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'frames/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8/f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'note': 'CPU port (oracle/) of helper/stereo_core.py; the Python reference cannot travel to the GPU box'},
            'cpu_baseline': base, 'e2e': {'value': value, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=300, help='frames per step per GPU (BASELINE.json configs[1]: a 300-frame clip)')
    ap.add_argument('--slots', type=int, default=30, help='slots (CUDA streams) per GPU')
    ap.add_argument('--group', type=int, default=4, help='frames per slot submission (share a stream and one hole-filling launch)')
    ap.add_argument('--cpu-frames', type=int, default=4, help='frames of the bounded CPU-baseline sample (about 3 s each on 16 threads)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        out = run_reference(args, rank, world)
    else:
        if world == 1 and args.gpus > 1:
            # launched without torchrun: spawn one rank per GPU ourselves
            cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={args.gpus}',
                   '--master-addr', '127.0.0.1', '--master-port', '29531', os.path.abspath(__file__)] + sys.argv[1:]
            sys.exit(subprocess.call(cmd))
        out = run_ours(args, rank, world, local_rank)
    if out is not None:
        print(json.dumps(out), flush=True)


if __name__ == '__main__':
    main()
