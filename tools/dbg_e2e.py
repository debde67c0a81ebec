"""debug helper (not a test): bisect an end-to-end mismatch stage by stage"""
import ctypes as C, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'video-stereo-converter_b200'), os.path.join(ROOT, 'oracle')]
import numpy as np
import oracle as O
from vsc_b200 import _lib, StereoGenerator, StereoParams
from vsc_b200.synthetic import make_pair
gen = StereoGenerator('cuda', 1)
lib = _lib.load()
lib.vsc_debug_fetch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
def fetch(which, shape, dt):
    a = np.empty(shape, dt)
    _lib.check(lib.vsc_debug_fetch(gen._ctx.handle, which, _lib.ptr(a), a.nbytes))
    return a
cases = [((120, 160), np.uint8, {}), ((96, 200), np.float32, dict(super_sampling=2.0, edge_softness=0.0, depth_gamma=1.0, max_disparity=100.0, convergence=-50.0, artifact_smoothing=5.0))]
for shape, dt, kw in cases:
    rgb, depth = make_pair(shape[0], shape[1], seed=7, depth_dtype=dt)
    taps = {}
    p = O.Params(**kw)
    ref = O.process_frame(rgb, depth, p, taps)
    out = gen.process_frame(rgb, depth, StereoParams(**kw))
    g = O.geometry(shape[0], shape[1], p)
    h, w, sw, hs, ws = shape[0], shape[1], g['stretched_w'], g['hs'], g['ws']
    print(shape, kw, 'SBS bad', int((out != ref).any(axis=2).sum()))
    print('  rgb_st', int((fetch(0, (h, sw, 3), np.uint8) != taps['rgb_stretched']).sum()))
    print('  depth_norm', int((fetch(1, (h, sw), np.float32) != taps['depth_norm']).sum()))
    print('  depth_ss', int((fetch(2, (hs, ws), np.float32) != taps['depth_ss']).sum()))
    for i, side in enumerate(('left', 'right')):
        va = fetch(3 + i, (hs, ws, 4), np.uint8)
        wref = taps['warp_' + side].transpose(1, 2, 0).astype(np.uint8)
        print('  warp', side, 'colour', int((va[:, :, :3] != wref).any(axis=2).sum()), 'mask', int((va[:, :, 3] != taps['mask_' + side]).sum()),
              'vmask', int((fetch(7 + i, (hs, ws), np.uint8) != taps['mask_' + side]).sum()))
        if p.artifact_smoothing > 0:
            vb = fetch(5 + i, (hs, ws, 4), np.uint8)
        else:
            vb = va
        crop = g['left_crop'] if i == 0 else g['right_crop']
        bad = (vb[:, :, :3] != taps['inpaint_' + side]).any(axis=2)
        known = taps['mask_' + side] > 0
        print('  final view', side, 'bad in window', int(bad[:, crop:crop + g['crop_w']].sum()), 'bad on originally-valid px in window',
              int((bad & known)[:, crop:crop + g['crop_w']].sum()), 'bad anywhere', int(bad.sum()))
        sm = taps['smooth_' + side]
        M = O.dilate3(((1 - taps['mask_' + side].astype(np.float32)) * 255).astype(np.uint8)) > 0
        print('     bilateral (outside M) bad', int(((vb[:, :, :3] != sm).any(axis=2) & ~M).sum()))
