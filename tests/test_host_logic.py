"""Host-side logic that needs no GPU: frame-range sharding (incl. a world-size-2 gloo run), in-order
publishing, workflow/config handling and the CLI's early exits."""
import json
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from conftest import ROOT
from vsc_b200 import sharder
from vsc_b200.workflow import ConfigError, find_frame_pairs, get_path, load_config

PKG = os.path.join(ROOT, 'video-stereo-converter_b200')


@pytest.mark.parametrize('n,world', [(0, 4), (1, 8), (10, 3), (300, 8), (10000, 4), (17, 17)])
def test_contiguous_ranges_partition(n, world):
    r = sharder.contiguous_ranges(n, world)
    assert len(r) == world and r[0][0] == 0 and r[-1][1] == n
    assert all(a <= b for a, b in r) and all(r[i][1] == r[i + 1][0] for i in range(world - 1))
    per = -(-n // world) if n else 0
    assert all(b - a <= per for a, b in r)


@pytest.mark.parametrize('n,world,block', [(0, 2, 16), (5, 8, 16), (300, 8, 16), (10000, 4, 64), (33, 2, 16)])
def test_block_cyclic_is_a_partition_and_stays_in_step(n, world, block):
    parts = [sharder.block_cyclic(n, world, r, block) for r in range(world)]
    allidx = sorted(i for p in parts for i in p)
    assert allidx == list(range(n))
    assert all(p == sorted(p) for p in parts)
    # ranks advance together: after k items each, the done set covers a prefix up to a window of world*block
    for k in range(0, max(len(p) for p in parts) + 1, 7):
        done = sorted(i for p in parts for i in p[:k])
        if done:
            first_gap = next((j for j, v in enumerate(done) if v != j), len(done))
            assert done[-1] - first_gap < world * block + block


def test_in_order_publisher(tmp_path):
    finals = [str(tmp_path / f'sbs_{i:06d}.png') for i in range(6)]
    pub = sharder.InOrderPublisher(finals)

    def write(i):
        with open(pub.staged_path(finals[i]), 'wb') as f:
            f.write(b'x')
        pub.mark_ready(finals[i])

    visible = lambda: sorted(p.name for p in tmp_path.glob('sbs_*.png'))   # noqa: E731 - what the orchestrator globs
    write(2); write(1)
    assert pub.publish_available() == 0 and visible() == []          # frame 0 missing: nothing may appear
    write(0)
    assert pub.publish_available() == 3 and visible() == ['sbs_000000.png', 'sbs_000001.png', 'sbs_000002.png']
    write(5)
    assert pub.publish_available() == 0
    write(3); write(4)
    assert pub.run(timeout_s=2) and pub.done() and len(visible()) == 6
    assert not list(tmp_path.glob('*.ready')) and not list(tmp_path.glob('.*.part'))


def test_in_order_publisher_steps_over_skipped_frames_and_reports_publication(tmp_path):
    """A frame that cannot be produced is reported with mark_skipped: the frames after it still become visible (the
    reference logs an unreadable frame and carries on, sbs_generator.py:229-230), and on_published fires only after
    the final rename (inputs may be deleted then, not before)."""
    finals = [str(tmp_path / f'sbs_{i:06d}.png') for i in range(5)]
    seen = []
    pub = sharder.InOrderPublisher(finals, on_published=lambda i: seen.append((i, os.path.exists(finals[i]))))

    def write(i):
        with open(pub.staged_path(finals[i]), 'wb') as f:
            f.write(b'x')
        pub.mark_ready(finals[i])

    write(0); write(2); write(3)
    assert pub.publish_available() == 1                                # frame 1 unknown: 2 and 3 wait
    pub.mark_skipped(finals[1])
    assert pub.publish_available() == 3 and pub.skipped == [finals[1]]
    pub.mark_skipped(finals[4])
    assert pub.run(timeout_s=2) and pub.done()
    assert sorted(p.name for p in tmp_path.glob('sbs_*.png')) == ['sbs_000000.png', 'sbs_000002.png', 'sbs_000003.png']
    assert seen == [(0, True), (2, True), (3, True)]
    assert not list(tmp_path.glob('.*.skip')) and not list(tmp_path.glob('*.ready'))
    # two threads publishing the same list must neither skip nor crash
    finals2 = [str(tmp_path / f'b_{i:06d}.png') for i in range(200)]
    pub2 = sharder.InOrderPublisher(finals2)
    import threading
    ths = [threading.Thread(target=pub2.run, kwargs=dict(poll_s=0.001, timeout_s=10)) for _ in range(2)]
    for t in ths:
        t.start()
    for f in finals2:
        with open(pub2.staged_path(f), 'wb') as fh:
            fh.write(b'y')
        pub2.mark_ready(f)
    for t in ths:
        t.join()
    assert pub2.done() and len(list(tmp_path.glob('b_*.png'))) == 200


def test_sweep_parameter_grid():
    """sbs_sweep.py: slider names / ranges of sbs_tester.py:356-362, cartesian product."""
    sys.path.insert(0, PKG)
    import sbs_sweep
    assert sbs_sweep.parse_param('max_disparity=20:60:10') == ('max_disparity', [20.0, 30.0, 40.0, 50.0, 60.0])
    assert sbs_sweep.parse_param('depth_gamma=0.2, 0.35,0.5') == ('depth_gamma', [0.2, 0.35, 0.5])
    assert sbs_sweep.parse_param('super_sampling=1:4:1.5')[1] == [1.0, 2.5, 4.0]
    for bad in ('gamma=1', 'max_disparity=1', 'convergence=60', 'sharpen=0:16:0', 'sharpen', 'edge_softness='):
        with pytest.raises(ValueError):
            sbs_sweep.parse_param(bad)
    base = {k: 1.0 for k in sbs_sweep.SLIDERS}
    g = sbs_sweep.grid(base, ['max_disparity=10,20', 'sharpen=0:16:8'])
    assert len(g) == 6 and g[0]['max_disparity'] == 10.0 and g[0]['sharpen'] == 0.0 and g[5] == {**base, 'max_disparity': 20.0, 'sharpen': 16.0}
    assert sbs_sweep.grid(base, []) == [base]
    with pytest.raises(ValueError):
        sbs_sweep.grid(base, ['sharpen=1', 'sharpen=2'])


def test_raw_frame_sink_writes_in_clip_order(tmp_path):
    """frames arrive out of order from several workers, the stream is in clip order."""
    import threading
    n, h, w = 40, 3, 8
    path = str(tmp_path / 'sbs.rgb')
    sink = sharder.RawFrameSink(path, n, h, w, max_ahead=6, frame_numbers=range(100, 100 + n))

    def worker(ids):
        for i in ids:
            sink.put(i, np.full((h, w, 3), i, np.uint8))
    ths = [threading.Thread(target=worker, args=(list(range(k, n, 3)),)) for k in range(3)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert sink.close() and sink.written == n
    data = np.fromfile(path, np.uint8).reshape(n, h, w, 3)
    assert (data == np.arange(n, dtype=np.uint8)[:, None, None, None]).all()
    meta = json.loads(open(path + '.json').read())
    assert meta['pix_fmt'] == 'rgb24' and (meta['width'], meta['height'], meta['frames']) == (w, h, n)
    assert meta['frame_numbers'][0] == 100 and meta['bytes_per_frame'] == h * w * 3
    with pytest.raises(ValueError):
        sharder.RawFrameSink(str(tmp_path / 'x.rgb'), 1, h, w).put(0, np.zeros((h, w), np.uint8))


def test_raw_frame_sink_stops_at_a_gap(tmp_path):
    path = str(tmp_path / 'gap.rgb')
    sink = sharder.RawFrameSink(path, 4, 2, 2)
    for i in (0, 1, 3):
        sink.put(i, np.full((2, 2, 3), i, np.uint8))
    assert sink.close(timeout_s=5) is False and sink.written == 2        # frame 2 never came: the stream ends before it
    assert os.path.getsize(path) == 2 * 12


def _workflow(tmp_path, stereo=None, n=3):
    wf = tmp_path / 'wf'
    for d in ('frames', 'depth_maps', 'sbs'):
        (wf / d).mkdir(parents=True)
    cfg = {'input_video': 'in.mkv', 'output_video': 'out.mkv',
           'directories': {'frames': 'frames', 'depth_maps': 'depth_maps', 'sbs': 'sbs', 'chunks': 'chunks'},
           'stereo': stereo or {'max_disparity': 50.0, 'convergence': -10, 'super_sampling': 3.0, 'edge_softness': 20.0,
                                'artifact_smoothing': 1.0, 'depth_gamma': 0.2, 'sharpen': 14.0},
           'depth': {'save_16bit': False}, 'encoding': {'crf': 19, 'preset': 'slow'},
           'free_space': {'sbs_generator': 'none', 'chunk_generator': 'none'}}
    (wf / 'config.json').write_text(json.dumps(cfg))
    for i in range(n):
        (wf / 'frames' / f'frame_{i:06d}.png').write_bytes(b'')
    return wf


def test_workflow_config_rules(tmp_path):
    wf = _workflow(tmp_path)
    cfg = load_config(wf)
    assert cfg['stereo']['convergence'] == -10                      # ints are accepted for floats
    assert get_path(wf, cfg, 'sbs') == wf / 'sbs'
    with pytest.raises(KeyError):
        get_path(wf, cfg, 'nope')
    bad = json.loads((wf / 'config.json').read_text())
    del bad['stereo']['sharpen']
    (wf / 'config.json').write_text(json.dumps(bad))
    with pytest.raises(ConfigError):
        load_config(wf)
    bad['stereo']['sharpen'] = 'x'
    (wf / 'config.json').write_text(json.dumps(bad))
    with pytest.raises(ConfigError):
        load_config(wf)
    (wf / 'config.json').write_text('{not json')
    with pytest.raises(ConfigError):
        load_config(wf)
    with pytest.raises(ConfigError):
        load_config(tmp_path / 'missing')


def test_frame_pair_discovery_prefers_tif(tmp_path):
    wf = _workflow(tmp_path, n=4)
    (wf / 'depth_maps' / 'depth_frame_000000.png').write_bytes(b'')
    (wf / 'depth_maps' / 'depth_frame_000001.png').write_bytes(b'')
    (wf / 'depth_maps' / 'depth_frame_000001.tif').write_bytes(b'')
    (wf / 'depth_maps' / 'depth_frame_000003.tif').write_bytes(b'')
    pairs, missing, first, last = find_frame_pairs(wf / 'frames', wf / 'depth_maps')
    assert [p[2] for p in pairs] == ['000000', '000001', '000003']
    assert pairs[1][1].suffix == '.tif' and pairs[0][1].suffix == '.png'
    assert (missing, first, last) == (1, '000002', '000002')


def _run_cli(*args):
    return subprocess.run([sys.executable, os.path.join(PKG, 'sbs_generator.py'), *map(str, args)], capture_output=True, text=True, timeout=120)


def test_cli_early_exits_match_the_reference(tmp_path):
    r = _run_cli(tmp_path / 'nope')
    assert r.returncode == 0 and 'Workflow directory not found' in r.stdout
    wf = _workflow(tmp_path, n=0)
    r = _run_cli(wf, '--no-interactive')
    assert r.returncode == 0 and 'All frames already processed.' in r.stdout
    r = _run_cli(wf, '--cpu')
    assert r.returncode != 0 and 'no CPU' in r.stdout
    (wf / 'config.json').write_text('{}')
    r = _run_cli(wf)
    assert r.returncode == 0 and r.stdout.startswith('ERROR')


GLOO_WORKER = textwrap.dedent('''
    import os, sys, json
    sys.path.insert(0, sys.argv[1])
    import torch, torch.distributed as dist
    from vsc_b200 import sharder
    rank, world, local = sharder.dist_env()
    dist.init_process_group('gloo')
    n = int(sys.argv[2])
    mine = sharder.shard_items(list(range(n)), world, rank, block=4)
    # "process" the shard: every rank reports what it did; no data-path collective, only bookkeeping
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    t = sharder.all_reduce_max(float(rank + 1))
    sharder.barrier()
    if rank == 0:
        print(json.dumps({'parts': gathered, 'tmax': t, 'world': world}))
    dist.destroy_process_group()
''')


def test_sharding_under_torchrun_gloo_world2(tmp_path):
    script = tmp_path / 'w.py'
    script.write_text(GLOO_WORKER)
    env = dict(os.environ, OMP_NUM_THREADS='1')
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr', '127.0.0.1',
                        '--master-port', '29617', str(script), PKG, '37'], capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith('{')][-1]
    out = json.loads(line)
    assert out['world'] == 2 and out['tmax'] == 2.0
    assert sorted(out['parts'][0] + out['parts'][1]) == list(range(37))
    assert not set(out['parts'][0]) & set(out['parts'][1])


@pytest.mark.parametrize('force_port', [False, True])
def test_bench_reference_arm_prints_contract_line(force_port):
    """--impl reference: the reference's CPU path on the host cores (the unmodified module when it is available or
    staged, else the oracle port), same metric/unit, h2d/d2h 0 (tiny sample via env override)."""
    env = dict(os.environ, VSC_BENCH_TINY='1')
    if force_port:
        env['VSC_BENCH_FORCE_PORT'] = '1'
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d['impl'] == 'reference' and d['unit'] == 'frames/s' and d['higher_is_better'] is True
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert d['cpu_baseline']['kind'] in (('port',) if force_port else ('reference', 'port'))
    assert d['cpu_baseline']['cores'] >= 1 and d['value'] > 0


def test_div3_identity(tmp_path):
    """The back end divides the 3x3 pooling sums by 3 with a multiply and two FMAs (csrc/vsc_kernels.cuh div3_exact);
    the identity is checked against the IEEE division for every float of the range on hosts with FMA hardware
    (~20 s), on a 1/64 sample otherwise (software fmaf)."""
    import shutil
    import subprocess
    if shutil.which('gcc') is None:
        pytest.skip('no gcc')
    here = os.path.dirname(os.path.abspath(__file__))
    hw_fma = False
    try:
        hw_fma = ' fma ' in open('/proc/cpuinfo').read()
    except OSError:
        pass
    exe = str(tmp_path / 'div3_check')
    flags = ['-O2', '-ffp-contract=off'] + (['-mfma'] if hw_fma else [])
    subprocess.run(['gcc'] + flags + ['-o', exe, os.path.join(here, 'div3_check.c'), '-lm'], check=True)
    out = subprocess.run([exe, '1' if hw_fma else '64'], check=True, capture_output=True, text=True, timeout=600).stdout.split()
    assert int(out[0]) == 0 and int(out[1]) > 1_000_000
