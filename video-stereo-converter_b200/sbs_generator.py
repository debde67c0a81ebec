#!/usr/bin/env python
"""Stereo 3D Generator on B200 — drop-in for the reference's sbs_generator.py.

Same command line, files and exit codes as /root/reference/sbs_generator.py
(`python sbs_generator.py <workflow_dir> [--cpu] [--no-interactive]`, :131-133; resume by skipping
existing sbs_<n>.png :178-185; exit code 100 on a GPU failure :41,317; free-space deletion :279-290;
a carriage-return progress line on stdout that the orchestrator scrapes), but the frame loop
(:304-328) is a batched, pinned-host, multi-stream pipeline:

    loader pool --(decode straight into the slot's pinned buffers)--> submit (H2D | kernels | D2H on
    the slot's CUDA stream) --> collect --> saver pool (PNG encode) --> in-order publish

and a frame-range sharder spreads one clip over the GPUs of the box (`--gpus N`, or launch under
torchrun with one process per GPU).  Frames are independent, so there is no collective.
There is no CPU path: `--cpu` is refused.
"""
from __future__ import annotations

import os
import sys
import threading
import time
from argparse import ArgumentParser, RawDescriptionHelpFormatter
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GPU_ERROR_EXIT_CODE = 100      # reference: sbs_generator.py:41


def _load_config_api():
    """Prefer the reference's own config_manager when this file sits in the reference tree."""
    try:
        from helper.config_manager import ConfigError, get_path, load_config  # type: ignore
        return ConfigError, get_path, load_config
    except Exception:
        from vsc_b200.workflow import ConfigError, get_path, load_config
        return ConfigError, get_path, load_config


def build_parser() -> ArgumentParser:
    p = ArgumentParser(formatter_class=RawDescriptionHelpFormatter,
                       description='Stereo 3D Generator - Create side-by-side stereo images (B200 CUDA path)',
                       epilog='Example:\n  python sbs_generator.py "D:/Video-Processing/workflow"\n')
    p.add_argument('workflow_path', type=Path, help='Path to workflow directory containing config.json')
    p.add_argument('--cpu', action='store_true', help='(reference flag) refused: this implementation has no CPU path')
    p.add_argument('--no-interactive', action='store_true', help='Exit on error instead of waiting for user input (for orchestrator)')
    p.add_argument('--gpus', type=int, default=0, help='number of GPUs to shard the clip over (0 = all visible, or the torchrun world)')
    p.add_argument('--slots', type=int, default=6, help='frames in flight per GPU')
    p.add_argument('--io-threads', type=int, default=8, help='loader and saver threads per GPU')
    p.add_argument('--raw-sink', default=None, metavar='PATH',
                   help="hand the SBS frames to an encoder instead of writing sbs_*.png: raw rgb24, in clip order, into PATH "
                        "(a file or FIFO; '-' = stdout); geometry in PATH.json.  Processes every frame pair (no resume)")
    return p


def process_shard(pairs, output_dir: Path, params, device_index: int, slots: int, io_threads: int, free_space_mode: str,
                  no_interactive: bool, progress=None, sink=None, sink_index=None) -> int:
    """Run `pairs` [(frame_path, depth_path, frame_num)] through one GPU.  Returns frames written.
    With `sink` (a RawFrameSink shared by all GPUs of the process) frames go to it at position
    `sink_index[frame_num]` instead of into sbs_<n>.png."""
    import cv2
    import numpy as np
    from vsc_b200 import StereoGenerator, load_image_pair
    from vsc_b200.sharder import InOrderPublisher

    gen = StereoGenerator(f'cuda:{device_index}', n_slots=slots)
    loaders = ThreadPoolExecutor(max_workers=io_threads, thread_name_prefix='load')
    savers = ThreadPoolExecutor(max_workers=io_threads, thread_name_prefix='save')
    save_failed = threading.Event()
    written = [0]
    lock = threading.Lock()

    def load(item):
        try:
            return load_image_pair(item[0], item[1])
        except Exception as e:          # reference: the loader logs and skips (sbs_generator.py:229-230)
            print(f'  Error loading {item[2]}: {e}')
            return None

    def save(sbs, item):
        if sink is not None:
            try:
                sink.put(sink_index[item[2]], sbs)
            except Exception as e:
                print(f'\nRaw sink failed at SBS frame #{item[2]}: {e}')
                save_failed.set()
                return
            with lock:
                written[0] += 1
            if progress:
                progress(item)
            return
        final = str(output_dir / f'sbs_{item[2]}.png')
        staged = InOrderPublisher.staged_path(final)
        for attempt in range(3):        # reference: 3 retries, 60 s apart (sbs_generator.py:241-262)
            try:
                ok, buf = cv2.imencode('.png', cv2.cvtColor(sbs, cv2.COLOR_RGB2BGR))
                if not ok:
                    raise IOError(f'PNG encode failed for {final}')
                with open(staged, 'wb') as f:
                    f.write(buf.tobytes())
                InOrderPublisher.mark_ready(final)
                break
            except Exception as e:
                print(f'\nSave failed for SBS frame #{item[2]} ({attempt + 1}/3): {e}')
                if attempt == 2:
                    save_failed.set()
                    return
                time.sleep(60 if not os.environ.get('VSC_FAST_RETRY') else 0.01)
        if free_space_mode in ('frame', 'all'):
            Path(item[0]).unlink(missing_ok=True)
        if free_space_mode in ('depth', 'all'):
            Path(item[1]).unlink(missing_ok=True)
        with lock:
            written[0] += 1
        if progress:
            progress(item)

    prefetch = [loaders.submit(load, it) for it in pairs[:slots * 2]]
    nxt_load = len(prefetch)
    inflight = []           # (slot, item)
    free_slots = list(range(slots))
    save_futs = []
    try:
        for i, item in enumerate(pairs):
            loaded = prefetch[i].result()
            if nxt_load < len(pairs):
                prefetch.append(loaders.submit(load, pairs[nxt_load]))
                nxt_load += 1
            if loaded is None:
                if sink is not None:
                    sink.skip(sink_index[item[2]])      # keep the stream gap-free
                continue
            if save_failed.is_set():
                break
            if not free_slots:
                # reuse whichever slot finishes first (a slow frame must not stall the others)
                s = gen.wait_any([sl for sl, _ in inflight])
                it = next(t for sl, t in inflight if sl == s)
                inflight.remove((s, it))
                save_futs.append(savers.submit(save, gen.collect(s), it))
                free_slots.append(s)
            s = free_slots.pop(0)
            gen.submit(s, loaded[0], loaded[1], params)
            inflight.append((s, item))
            prefetch[i] = None
        while inflight:
            s, it = inflight.pop(0)
            save_futs.append(savers.submit(save, gen.collect(s), it))
        for f in save_futs:
            f.result()
    finally:
        loaders.shutdown(wait=False, cancel_futures=True)
        savers.shutdown(wait=True)
        gen.close()
    if save_failed.is_set():
        print('\nERROR: Failed to write output file.' + (' Exiting (non-interactive mode).' if no_interactive else ''))
    return written[0]


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    if args.raw_sink == '-':
        sys.stdout = sys.stderr          # stdout carries the frames
    if not args.workflow_path.is_dir():
        print(f'ERROR: Workflow directory not found: {args.workflow_path}')
        return 0
    ConfigError, get_path, load_config = _load_config_api()
    try:
        config = load_config(args.workflow_path)
    except ConfigError as e:
        print(f'ERROR: {e}')
        return 0
    frames_dir = get_path(args.workflow_path, config, 'frames')
    depth_dir = get_path(args.workflow_path, config, 'depth_maps')
    output_dir = get_path(args.workflow_path, config, 'sbs')
    if not frames_dir.exists():
        print(f'ERROR: Frames directory not found: {frames_dir}')
        return 0
    if not depth_dir.exists():
        print(f'ERROR: Depth directory not found: {depth_dir}')
        return 0
    output_dir.mkdir(parents=True, exist_ok=True)
    if args.cpu:
        print('ERROR: --cpu is not supported: the B200 SBS path has no CPU fallback (use the reference implementation).')
        return 2

    from vsc_b200 import StereoParams
    from vsc_b200._lib import VscCudaError
    from vsc_b200.sharder import InOrderPublisher, barrier, dist_env, shard_items
    from vsc_b200.workflow import find_frame_pairs
    sc = config['stereo']
    params = StereoParams(**{k: sc[k] for k in ('max_disparity', 'convergence', 'super_sampling', 'edge_softness',
                                                'artifact_smoothing', 'depth_gamma', 'sharpen')})
    rank, world, local_rank = dist_env()
    if rank == 0:
        print('Scanning for frame pairs...')
    all_pairs, missing, first, last = find_frame_pairs(frames_dir, depth_dir)
    if missing and rank == 0:
        print(f'Missing depth maps: {missing} of {missing + len(all_pairs)} frames in range of frame_{first} to frame_{last}')
    pairs = [p for p in all_pairs if args.raw_sink or not (output_dir / f'sbs_{p[2]}.png').exists()]
    skipped = len(all_pairs) - len(pairs)
    if rank == 0:
        print(f'Found: {len(all_pairs)} frame pairs, {skipped} already processed, {len(pairs)} to process')
    if not pairs:
        if rank == 0:
            print('All frames already processed.')
        return 0

    import torch
    if not torch.cuda.is_available():
        print('ERROR: no CUDA device available and this implementation has no CPU path')
        return GPU_ERROR_EXIT_CODE
    ngpu = torch.cuda.device_count()
    free_space_mode = config.get('free_space', {}).get('sbs_generator', 'none')
    if rank == 0:
        print(f'Using GPU: {torch.cuda.get_device_name(0)} x{world if world > 1 else (args.gpus or ngpu)}')
        print(f'Parameters: disparity={params.max_disparity}, convergence={params.convergence}, '
              f'super_sampling={params.super_sampling}, edge_softness={params.edge_softness}, '
              f'smoothing={params.artifact_smoothing}, gamma={params.depth_gamma}, sharpen={params.sharpen}')

    finals = [str(output_dir / f'sbs_{p[2]}.png') for p in pairs]
    sink = sink_index = None
    if args.raw_sink:
        if world > 1:
            print('ERROR: --raw-sink needs all GPUs in one process (use --gpus N, not torchrun)')
            return 2
        import cv2
        from vsc_b200.sharder import RawFrameSink
        probe = cv2.imread(str(pairs[0][0]), cv2.IMREAD_COLOR)
        if probe is None:
            print(f'ERROR: cannot read {pairs[0][0]}')
            return 0
        sink = RawFrameSink(args.raw_sink, len(pairs), probe.shape[0], 2 * probe.shape[1], frame_numbers=[p[2] for p in pairs])
        sink_index = {p[2]: i for i, p in enumerate(pairs)}
    t0 = time.time()
    done = [0]

    def progress(_item):
        done[0] += 1
        if rank == 0:
            el = time.time() - t0
            sys.stdout.write(f'\r{skipped + done[0] * max(1, world)}/{len(all_pairs)} [{el:.0f}s, {done[0] * max(1, world) / max(el, 1e-6):.2f}img/s]')
            sys.stdout.flush()

    try:
        if world > 1:
            # torchrun: this process is one rank = one GPU
            import torch.distributed as dist
            dist.init_process_group('nccl' if torch.cuda.is_available() else 'gloo')
            mine = shard_items(pairs, world, rank)
            pub = InOrderPublisher(finals)
            t = threading.Thread(target=pub.run, daemon=True) if rank == 0 else None
            if t:
                t.start()
            n = process_shard(mine, output_dir, params, local_rank, args.slots, args.io_threads, free_space_mode, args.no_interactive, progress)
            barrier()
            if t:
                pub.run(timeout_s=600)
            dist.destroy_process_group()
        else:
            g = args.gpus or ngpu
            g = max(1, min(g, ngpu, len(pairs)))
            pub = InOrderPublisher([] if sink is not None else finals)
            t = threading.Thread(target=pub.run, daemon=True)
            t.start()
            if g == 1:
                n = process_shard(pairs, output_dir, params, 0, args.slots, args.io_threads, free_space_mode, args.no_interactive, progress,
                                  sink, sink_index)
            else:
                # one worker thread per GPU; each owns a StereoGenerator (its own CUDA context state, streams, pinned ring)
                results = [0] * g
                errors = []

                def work(k):
                    try:
                        results[k] = process_shard(shard_items(pairs, g, k), output_dir, params, k, args.slots, args.io_threads,
                                                   free_space_mode, args.no_interactive, progress, sink, sink_index)
                    except BaseException as e:   # noqa: BLE001
                        errors.append(e)
                ths = [threading.Thread(target=work, args=(k,)) for k in range(g)]
                for th in ths:
                    th.start()
                for th in ths:
                    th.join()
                if errors:
                    raise errors[0]
                n = sum(results)
            pub.run(timeout_s=600)
            if sink is not None and not sink.close():
                print(f'\nWARNING: raw sink holds {sink.written} of {len(pairs)} frames (stream ends at the first missing frame)')
    except VscCudaError as e:
        print(f'\nERROR: GPU failure - {e}')
        return GPU_ERROR_EXIT_CODE
    if rank == 0:
        print(f'\nDone! Processed {n if world == 1 else len(pairs)} of {len(pairs)} frames.')
    return 0


if __name__ == '__main__':
    sys.exit(main())
