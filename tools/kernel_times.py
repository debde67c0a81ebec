"""profiling helper (not a test): per-kernel device times of one frame, one slot in flight
usage: python tools/kernel_times.py [H W [u16]]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'video-stereo-converter_b200')]
import numpy as np
from vsc_b200 import StereoGenerator
from vsc_b200.synthetic import make_pair
h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1080, 1920)
dt = np.uint16 if len(sys.argv) > 3 and sys.argv[3] == 'u16' else np.uint8
gen = StereoGenerator('cuda', int(os.environ.get('SLOTS', '1')))
rgb, depth = make_pair(h, w, seed=1, depth_dtype=dt)
for i in range(3):
    gen.process_frame(rgb, depth)
gen.set_profiling(True)
tot = {}
for i in range(3):
    gen.process_frame(rgb, depth)
    for n, t in gen.kernel_times(0):
        tot[n] = tot.get(n, 0.0) + t / 3
for n, t in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f'  {n:28s} {t:8.3f} ms')
print(f'  {"sum":28s} {sum(tot.values()):8.3f} ms   ({h}x{w}, {np.dtype(dt).name} depth)')
