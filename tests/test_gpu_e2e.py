"""End-to-end parity of StereoGenerator.process_frame (CUDA, through the C ABI) against the oracle,
plus size-independent properties at the benchmark sizes."""
import numpy as np
import pytest

import oracle as O
from vsc_b200 import StereoGenerator, StereoParams
from vsc_b200 import _lib
from vsc_b200.synthetic import make_pair

pytestmark = pytest.mark.gpu

CASES = [
    ((120, 160), np.uint8, {}),
    ((135, 240), np.uint16, {}),
    ((100, 180), np.uint8, dict(super_sampling=1.0, edge_softness=0.0, depth_gamma=1.0, max_disparity=30.0, convergence=5.0,
                                artifact_smoothing=5.0)),
    ((90, 150), np.uint16, dict(super_sampling=2.5, edge_softness=3.0, depth_gamma=0.5, max_disparity=20.0, convergence=-7.0,
                                artifact_smoothing=0.0, sharpen=0.0)),
    ((96, 200), np.float32, dict(super_sampling=2.0, edge_softness=0.0, depth_gamma=1.0, max_disparity=100.0, convergence=-50.0,
                                 artifact_smoothing=5.0)),
    ((64, 256), np.uint8, dict(super_sampling=4.0, edge_softness=30.0, depth_gamma=2.0, max_disparity=100.0, convergence=50.0,
                               artifact_smoothing=2.5, sharpen=16.0)),
    ((80, 120), np.uint16, dict(super_sampling=1.3, edge_softness=0.5, depth_gamma=0.1, max_disparity=5.0, convergence=0.0,
                                artifact_smoothing=0.1, sharpen=0.5)),
]


@pytest.fixture(scope='module')
def gen():
    g = StereoGenerator('cuda', n_slots=3)
    yield g
    g.close()


@pytest.mark.parametrize('shape,dt,kw', CASES)
def test_process_frame_matches_oracle(gen, shape, dt, kw):
    rgb, depth = make_pair(shape[0], shape[1], seed=7, depth_dtype=dt)
    out = gen.process_frame(rgb, depth, StereoParams(**kw))
    ref = O.process_frame(rgb, depth, O.Params(**kw))
    assert out.shape == ref.shape == (shape[0], 2 * shape[1], 3) and out.dtype == np.uint8
    diff = np.abs(out.astype(int) - ref.astype(int))
    assert np.array_equal(out, ref), f'{(diff > 0).sum()} values differ, max {diff.max()}'


def test_default_params_and_none(gen):
    rgb, depth = make_pair(72, 128, seed=1)
    assert np.array_equal(gen.process_frame(rgb, depth), gen.process_frame(rgb, depth, StereoParams()))


def test_inputs_not_mutated_and_output_owned(gen):
    rgb, depth = make_pair(72, 128, seed=2)
    r0, d0 = rgb.copy(), depth.copy()
    a = gen.process_frame(rgb, depth)
    b = gen.process_frame(rgb, depth)
    assert np.array_equal(rgb, r0) and np.array_equal(depth, d0)
    assert a is not b and np.array_equal(a, b) and a.flags['C_CONTIGUOUS'] and a.flags['OWNDATA']


def test_invalid_crop_raises(gen):
    """md=5, conv=50: a crop offset is negative; the reference raises RuntimeError."""
    rgb, depth = make_pair(72, 128, seed=2)
    with pytest.raises(RuntimeError):
        gen.process_frame(rgb, depth, StereoParams(max_disparity=5.0, convergence=50.0))
    with pytest.raises(RuntimeError):
        O.process_frame(rgb, depth, O.Params(max_disparity=5.0, convergence=50.0))
    # md=25, conv=-50 (offset 0) is the first valid point
    out = gen.process_frame(rgb, depth, StereoParams(max_disparity=25.0, convergence=-50.0))
    assert np.array_equal(out, O.process_frame(rgb, depth, O.Params(max_disparity=25.0, convergence=-50.0)))


def test_near_black_quirk(gen):
    """An all-ones frame with artifact_smoothing > 0 comes back all-255 (stereo_core.py:404-407)."""
    rgb = np.ones((64, 96, 3), np.uint8)
    _, depth = make_pair(64, 96, seed=3)
    out = gen.process_frame(rgb, depth)
    assert np.array_equal(out, O.process_frame(rgb, depth))
    assert (out >= 254).all()      # the x255 branch was taken (torch 2.11 gives 254 after the unsharp round trip)
    out0 = gen.process_frame(rgb, depth, StereoParams(artifact_smoothing=0.0))
    assert np.array_equal(out0, O.process_frame(rgb, depth, O.Params(artifact_smoothing=0.0)))


def test_near_black_quirk_many_rows(gen):
    """Same quirk on a frame with more (row, segment) work items than the conditional re-run's grid holds, so its
    CTAs stride over several items (and one eye only: the other eye's maximum is above 1.0)."""
    h, w = 420, 700
    rgb = np.ones((h, w, 3), np.uint8)
    _, depth = make_pair(h, w, seed=5)
    out = gen.process_frame(rgb, depth)
    assert np.array_equal(out, O.process_frame(rgb, depth)) and (out >= 254).all()
    rgb[:, :8] = 7                       # reaches only the left part of the views
    out = gen.process_frame(rgb, depth)
    assert np.array_equal(out, O.process_frame(rgb, depth))


def test_flat_depth(gen):
    rgb, _ = make_pair(64, 96, seed=4)
    depth = np.full((64, 96), 77, np.uint8)
    assert np.array_equal(gen.process_frame(rgb, depth), O.process_frame(rgb, depth))


def test_cpu_device_is_refused():
    with pytest.raises(RuntimeError):
        StereoGenerator('cpu')


def test_async_slots_match_sync(gen):
    frames = [make_pair(90, 160, seed=s, depth_dtype=np.uint16) for s in range(7)]
    sync = [gen.process_frame(r, d) for r, d in frames]
    batch = gen.process_batch(frames)
    assert len(batch) == len(sync)
    for a, b in zip(sync, batch):
        assert np.array_equal(a, b)


def test_queue_scratch_overflow_is_detected_and_rerun():
    """The march's queue scratch is sized optimistically; a frame that needs more must be re-run with a larger
    scratch and still be exact (vsc_wait), also through the asynchronous slots."""
    g = StereoGenerator('cuda', n_slots=2)
    try:
        lib = _lib.load()
        kw = dict(edge_softness=0.0, depth_gamma=1.0, super_sampling=2.0, max_disparity=60.0)
        rgb, depth = make_pair(120, 200, seed=21)
        ref = O.process_frame(rgb, depth, O.Params(**kw))
        _lib.check(lib.vsc_debug_set_telea_capacity(g._ctx.handle, 64))
        assert np.array_equal(g.process_frame(rgb, depth, StereoParams(**kw)), ref)
        _lib.check(lib.vsc_debug_set_telea_capacity(g._ctx.handle, 64))
        outs = g.process_batch([(rgb, depth)] * 3, StereoParams(**kw))
        assert all(np.array_equal(o, ref) for o in outs)
    finally:
        g.close()
    g = StereoGenerator('cuda', n_slots=1, group_size=3)         # the whole group is re-run when one frame overflows
    try:
        _lib.check(lib.vsc_debug_set_telea_capacity(g._ctx.handle, 64))
        rgb2, depth2 = make_pair(120, 200, seed=22)
        outs = g.process_batch([(rgb, depth), (rgb2, depth2), (rgb, depth)], StereoParams(**kw))
        assert np.array_equal(outs[0], ref) and np.array_equal(outs[2], ref)
        assert np.array_equal(outs[1], O.process_frame(rgb2, depth2, O.Params(**kw)))
    finally:
        g.close()


def test_grouped_slots_match_single_frames():
    """Slots that take several frames per submission (shared stream, one hole-filling launch for all of them)
    must give exactly the single-frame results, for full and partial groups and for mixed sizes."""
    frames = [make_pair(90, 160, seed=s, depth_dtype=np.uint16) for s in range(8)] + [make_pair(72, 128, seed=40)]
    single = StereoGenerator('cuda', n_slots=1)
    ref = [single.process_frame(r, d) for r, d in frames]
    single.close()
    g = StereoGenerator('cuda', n_slots=2, group_size=3)
    try:
        outs = g.process_batch(frames)
        assert len(outs) == len(ref)
        for a, b in zip(outs, ref):
            assert np.array_equal(a, b)
        g.submit_frames(0, frames[:2])
        g.submit_frames(1, frames[2:5])
        r1, r0 = g.collect(1), g.collect(0)
        assert all(np.array_equal(a, b) for a, b in zip(r0 + r1, ref[:5]))
        assert np.array_equal(g.process_frame(*frames[5]), ref[5])
        with pytest.raises(ValueError):
            g.submit_frames(0, frames[:4])
    finally:
        g.close()


def _sweep_cases(n=28, seed=2024):
    """Random points of the sbs_tester sliders (reference sbs_tester.py:356-362: ranges and step sizes), mixed sizes / dtypes."""
    rng = np.random.default_rng(seed)
    step = lambda lo, hi, st: float(lo + st * rng.integers(0, int(round((hi - lo) / st)) + 1))   # noqa: E731
    cases = []
    for i in range(n):
        kw = dict(max_disparity=step(5, 100, 0.5), convergence=step(-50, 50, 1.0), super_sampling=round(step(1.0, 4.0, 0.1), 1),
                  edge_softness=step(0, 30, 0.5), artifact_smoothing=round(step(0, 5, 0.1), 1), depth_gamma=round(step(0.1, 2.0, 0.05), 2),
                  sharpen=step(0, 16, 0.5))
        h, w = int(rng.integers(40, 110)), int(rng.integers(160, 330))
        cases.append((h, w, [np.uint8, np.uint16, np.float32][i % 3], kw))
    return cases


@pytest.mark.parametrize('h,w,dt,kw', _sweep_cases())
def test_tester_slider_sweep_matches_oracle(gen, h, w, dt, kw):
    """BASELINE.json configs[4] (the sbs_tester parameter sweep), at sizes the oracle finishes quickly: bit-exact, or
    the same refusal when the convergence crop leaves no valid window."""
    rgb, depth = make_pair(h, w, seed=h * 1000 + w, depth_dtype=dt)
    p = StereoParams(**kw)
    try:
        want = O.process_frame(rgb, depth, O.Params(**kw))
    except (RuntimeError, ValueError):
        with pytest.raises(RuntimeError):
            gen.process_frame(rgb, depth, p)
        return
    got = gen.process_frame(rgb, depth, p)
    assert got.shape == want.shape and np.array_equal(got, want), (kw, int((got != want).sum()))


def test_results_do_not_depend_on_load():
    """The hole-filling dataflow must not depend on timing: frames processed while many other frames are in flight
    (8 slots x 2 frames, several rounds) equal, bit for bit, the same frames processed alone.  The solo context also
    runs the 16-warp (latency) variant of the march, the loaded one the 8-warp (throughput) variant."""
    frames = [make_pair(270, 480, seed=s, depth_dtype=np.uint16 if s % 2 else np.uint8) for s in range(10)]
    single = StereoGenerator('cuda', n_slots=1)
    ref = [single.process_frame(r, d) for r, d in frames]
    single.close()
    g = StereoGenerator('cuda', n_slots=8, group_size=2)
    try:
        for _ in range(3):
            order = [(i * 3) % len(frames) for i in range(3 * len(frames))]
            outs = g.process_batch([frames[i] for i in order if frames[i][1].dtype == np.uint8])
            exp = [ref[i] for i in order if frames[i][1].dtype == np.uint8]
            assert len(outs) == len(exp) and all(np.array_equal(a, b) for a, b in zip(outs, exp))
            outs = g.process_batch([frames[i] for i in order if frames[i][1].dtype == np.uint16])
            exp = [ref[i] for i in order if frames[i][1].dtype == np.uint16]
            assert all(np.array_equal(a, b) for a, b in zip(outs, exp))
    finally:
        g.close()


def test_ready_and_wait_any(gen):
    frames = [make_pair(90, 160, seed=s) for s in range(5)]
    ref = [gen.process_frame(r, d) for r, d in frames]
    for s in range(3):
        gen.submit(s, *frames[s])
    done = []
    pending = [0, 1, 2]
    while pending:
        s = gen.wait_any(pending)
        assert gen.ready(s)
        done.append((s, gen.collect(s)))
        pending.remove(s)
    for s, out in done:
        assert np.array_equal(out, ref[s])
    with pytest.raises(RuntimeError):
        gen.ready(0)            # nothing in flight any more


def test_module_level_helpers():
    import torch
    from vsc_b200 import apply_depth_gamma, forward_warp_stereo, normalize_depth
    rng = np.random.default_rng(0)
    d = torch.from_numpy(rng.random((1, 1, 40, 90), dtype=np.float32) * 7 + 3)
    n = normalize_depth(d)
    assert n.shape == d.shape and np.array_equal(n.numpy().ravel(), O.normalize_depth(d.numpy().ravel()))
    g = apply_depth_gamma(n, 0.2)
    assert np.array_equal(g.numpy().ravel(), O.apply_gamma(n.numpy().ravel(), 0.2))
    img = torch.from_numpy(rng.integers(0, 256, (1, 3, 40, 90)).astype(np.float32))
    lw, lm, rw, rm = forward_warp_stereo(img, g, 33.0)
    ol, olm = O.warp(img[0].numpy(), g[0, 0].numpy(), 33.0, +1)
    orr, orm = O.warp(img[0].numpy(), g[0, 0].numpy(), 33.0, -1)
    assert lw.shape == (1, 3, 40, 90) and lm.shape == (1, 1, 40, 90)
    assert np.array_equal(lw[0].numpy(), ol) and np.array_equal(rw[0].numpy(), orr)
    assert np.array_equal(lm[0, 0].numpy().astype(np.uint8), olm) and np.array_equal(rm[0, 0].numpy().astype(np.uint8), orm)


# ---- benchmark sizes ---------------------------------------------------------------------------
def _compare_full(gen, h, w, dt, kw):
    rgb, depth = make_pair(h, w, seed=0, depth_dtype=dt)
    out = gen.process_frame(rgb, depth, StereoParams(**kw))
    assert np.array_equal(out, gen.process_frame(rgb, depth, StereoParams(**kw))), 'not deterministic'
    ref = O.process_frame(rgb, depth, O.Params(**kw))
    diff = np.abs(out.astype(int) - ref.astype(int))
    assert np.array_equal(out, ref), f'{(diff > 0).sum()} values differ, max {diff.max()}'


def test_1080p_default_full_frame(gen):
    """BASELINE.json configs[0]/[1]: one 1920x1080 frame, uint8 depth, default parameters."""
    _compare_full(gen, 1080, 1920, np.uint8, {})


def test_4k_u16_full_frame(gen):
    """BASELINE.json configs[2]: 3840x2160, 16-bit depth, full-width SBS output."""
    _compare_full(gen, 2160, 3840, np.uint16, {})


def test_8k_band_aggressive(gen):
    """BASELINE.json configs[4] geometry (7680 wide, md=100, conv=-50, smoothing 5) on a 256-row band so
    that the oracle stays cheap; stretched_w must be 7929."""
    kw = dict(max_disparity=100.0, convergence=-50.0, super_sampling=1.0, edge_softness=0.0, depth_gamma=1.0,
              artifact_smoothing=5.0)
    assert _lib.geometry(3840, 7680, StereoParams(**kw)).stretched_w == 7929
    _compare_full(gen, 256, 7680, np.uint16, kw)
    kw['super_sampling'] = 2.0
    _compare_full(gen, 128, 7680, np.uint16, kw)


def test_8k_full_frame_sweep_point(gen):
    """One FULL 7680x3840 frame of BASELINE.json configs[4]: max disparity, convergence -50, aggressive hole filling
    (no edge softening, smoothing 5), 16-bit depth, super_sampling 1 (the largest the CPU side can check: the
    reference itself needs > 60 GB of host memory beyond SS 2).  Bit-exact against the oracle."""
    kw = dict(max_disparity=100.0, convergence=-50.0, super_sampling=1.0, edge_softness=0.0, depth_gamma=1.0,
              artifact_smoothing=5.0)
    _compare_full(gen, 3840, 7680, np.uint16, kw)
