"""debug helper: cycle seeds over slots and report slow frames"""
import os, sys, time
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'video-stereo-converter_b200')]
import numpy as np, torch
from vsc_b200 import StereoGenerator, StereoParams
from vsc_b200.synthetic import make_pair
h, w, slots = 1080, 1920, int(os.environ.get('SLOTS', '30'))
seeds = [int(a) for a in sys.argv[1:]]
g = StereoGenerator('cuda:0', slots)
d_out = [torch.empty((h, 2 * w, 3), dtype=torch.uint8, device='cuda') for _ in range(slots)]
fr = [make_pair(h, w, s) for s in seeds]
dr = [torch.from_numpy(r).cuda() for r, _ in fr]; dd = [torch.from_numpy(d).cuda() for _, d in fr]
ND = len(seeds)
def run(n, log=False):
    infl = []; lat = []
    for i in range(n):
        s = i % slots
        if len(infl) == slots:
            s0, i0, t0 = infl.pop(0); g.wait(s0); lat.append((time.perf_counter() - t0, seeds[i0 % ND], g.last_frame_ms(s0)))
        g.submit_device(s, dr[i % ND].data_ptr(), dd[i % ND].data_ptr(), np.uint8, h, w, d_out[s].data_ptr(), StereoParams()); infl.append((s, i, time.perf_counter()))
    while infl:
        s0, i0, t0 = infl.pop(0); g.wait(s0); lat.append((time.perf_counter() - t0, seeds[i0 % ND], g.last_frame_ms(s0)))
    return lat
run(2 * slots)
g.timer_begin(); lat = run(120); ms = g.timer_end()
print(f'seeds {seeds} fps {120/(ms*1e-3):.1f}')
by = {}
for l, s, dev in lat: by.setdefault(s, []).append((l, dev))
for s, v in by.items(): print('  seed', s, 'mean submit->done ms %.1f' % (1e3 * np.mean([a for a, _ in v])), 'mean device ms %.1f max %.1f' % (np.mean([b for _, b in v]), max(b for _, b in v)))
