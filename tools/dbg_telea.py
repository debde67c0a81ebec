"""debug helper (not a test): run small inpaint cases on the GPU and report where they deviate"""
import ctypes as C, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'video-stereo-converter_b200'), os.path.join(ROOT, 'oracle')]
import numpy as np
import oracle as O
from vsc_b200 import _lib
from vsc_b200.synthetic import make_rgb
lib = _lib.load()
ctx = _lib.Context(0, 1)
def run(img, hole, name):
    h, w = hole.shape
    valid = (~hole).astype(np.uint8)
    img = img.copy(); img[hole] = 0
    outs = []
    for rep in range(2):
        out = img.copy()
        _lib.check(lib.vsc_stage_inpaint(ctx.handle, _lib.ptr(out), _lib.ptr(valid), h, w, 0, w))
        outs.append(out)
    tt = np.empty((h, w), np.float32); st = np.empty((h, w), np.uint8)
    _lib.check(lib.vsc_debug_telea_state(ctx.handle, 0, _lib.ptr(tt), _lib.ptr(st), None, h * w))
    mask = ((1 - valid.astype(np.float32)) * 255).astype(np.uint8)
    M = O.dilate3(mask)
    ref, t = O.telea(img, M, 3, return_t=True)
    t = t[1:-1, 1:-1]
    bad = (outs[0] != ref).any(axis=2)
    print(f'{name}: holes {int((M>0).sum())} mismatch {int(bad.sum())} nondet {int((outs[0]!=outs[1]).any(axis=2).sum())}')
    # compare T: hole pixels
    hm = M > 0
    tdiff = (tt != t) & hm
    print('   T mismatches on holes:', int(tdiff.sum()), ' st f-bits hist on holes:', np.bincount(st[hm] & 3, minlength=4))
    ring = ((st >> 2) & 3) == 3
    tr = np.where(ring, -tt, tt)
    near = ring & ~hm
    print('   ring px', int(near.sum()), 'ring T mismatches', int((tr[near] != t[near]).sum()))
    if bad.any():
        ys, xs = np.nonzero(bad)
        k = np.argmin(t[ys, xs])  # earliest (smallest T) wrong pixel
        y, x = ys[k], xs[k]
        print('   first bad (by T) at', (y, x), 'T ref', t[y, x], 'T gpu', tt[y, x], 'ref', ref[y, x], 'gpu', outs[0][y, x])
        print('   bad T range', t[ys, xs].min(), t[ys, xs].max(), ' good hole count', int((hm & ~bad).sum()))
h, w = 40, 60
img = make_rgb(h, w, seed=2)
hole = np.zeros((h, w), bool); hole[20, 30] = True; run(img, hole, 'single')
hole = np.zeros((h, w), bool); hole[10:30, 30] = True; run(img, hole, 'vline')
hole = np.zeros((h, w), bool); hole[15:25, 20:40] = True; run(img, hole, 'block')
hole = np.zeros((h, w), bool); hole[:, :12] = True; run(img, hole, 'band')
rng = np.random.default_rng(1)
hole = rng.random((h, w)) < 0.01; run(img, hole, 'sparse')
h, w = 120, 170
img = make_rgb(h, w, seed=2)
hole = rng.random((h, w)) < 0.004; run(img, hole, '1px-big')
