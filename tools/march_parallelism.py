"""analysis helper (CPU only): how long is the dependency chain of the colour pass of the hole filling?
For the hole mask of one synthetic 1080p default frame (one eye, after the 3x3 dilation), take the computation order
of the oracle's two-pass Telea and give every hole pixel the level 1 + max(level of the earlier hole pixels within
Chebyshev distance 4) - the earliest step at which an ordered dataflow could run it.  tasks / levels = the average
parallelism available to a design without generation barriers.
usage: python tools/march_parallelism.py [rows]      (rows: use only the top part of the frame, default 360)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'video-stereo-converter_b200'), os.path.join(ROOT, 'oracle')]
import numpy as np, cv2
import oracle as O
from vsc_b200.synthetic import make_pair
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 360
rgb, depth = make_pair(rows, 1920, seed=0)
taps = {}
O.process_frame(rgb, depth, O.Params(), taps)
for side in ('left', 'right'):
    valid = taps['mask_' + side] > 0
    mask = cv2.dilate((~valid).astype(np.uint8) * 255, np.ones((3, 3), np.uint8))
    img = np.zeros(mask.shape + (3,), np.uint8)
    _, order = O.telea_two_pass(img, mask, 3, return_order=True)
    ys, xs = np.nonzero((order >= 0) & (order < 2**31 - 1))
    idx = np.argsort(order[ys, xs])
    ys, xs = ys[idx], xs[idx]
    H, W = mask.shape
    level = np.zeros((H + 8, W + 8), np.int32)          # padded by 4
    ordp = np.full((H + 8, W + 8), -1, np.int64); ordp[4:-4, 4:-4] = order
    t0 = time.time()
    for k, (y, x) in enumerate(zip(ys, xs)):
        win_o = ordp[y:y + 9, x:x + 9]; win_l = level[y:y + 9, x:x + 9]
        m = (win_o >= 0) & (win_o < k)
        level[y + 4, x + 4] = 1 + (win_l[m].max() if m.any() else 0)
    # per cluster (holes linked within 15 px, the GPU's tile clustering to a good approximation): size, depth
    from scipy import ndimage
    lab, ncl = ndimage.label(cv2.dilate(mask, np.ones((15, 15), np.uint8)) > 0, structure=np.ones((3, 3)))
    cl = lab[ys, xs]
    lv = level[ys + 4, xs + 4]
    sizes = np.bincount(cl, minlength=ncl + 1); depths = np.zeros(ncl + 1, np.int64)
    np.maximum.at(depths, cl, lv)
    top = np.argsort(-sizes)[:5]
    print(f'{side}: {ncl} clusters; largest: ' + ', '.join(f'{sizes[c]} px / depth {depths[c]} (x{sizes[c] / max(depths[c], 1):.0f})' for c in top if sizes[c]))
    n, depth_ = len(ys), int(level.max())
    print(f'{side}: {n} hole pixels, dependency depth {depth_}, average parallelism {n / max(depth_, 1):.1f} '
          f'(analysis {time.time() - t0:.1f} s, {rows} rows)')
