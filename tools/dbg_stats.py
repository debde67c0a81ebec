"""profiling helper (not a test): Telea phase counters for one 1080p default frame (stats build)"""
import ctypes as C, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ['VSC_B200_LIB'] = os.path.join(ROOT, 'video-stereo-converter_b200', 'lib', 'libvsc_b200_stats.so')
sys.path[:0] = [os.path.join(ROOT, 'video-stereo-converter_b200')]
import numpy as np
from vsc_b200 import _lib, StereoGenerator, StereoParams
from vsc_b200.synthetic import make_pair
h, w, dt = (int(sys.argv[1]), int(sys.argv[2]), np.uint16 if len(sys.argv) > 3 and sys.argv[3] == 'u16' else np.uint8) if len(sys.argv) > 2 else (1080, 1920, np.uint8)
gen = StereoGenerator('cuda', 1)
lib = _lib.load()
rgb, depth = make_pair(h, w, seed=0, depth_dtype=dt)
for i in range(3):
    gen.process_frame(rgb, depth)
gen.set_profiling(True)
gen.process_frame(rgb, depth)
for n, t in gen.kernel_times(0):
    if t > 0.05: print(f'  {n:28s} {t:8.3f} ms')
st = (C.c_ulonglong * 64)()
_lib.check(lib.vsc_debug_telea_stats(gen._ctx.handle, st))
names = ['c_wait', 'c_pop', 'c_sort', 'c_part', 'c_total', 'n_pops', 'n_pix', 'n_gen', 'n_polls', 'n_clusters', 'max_total', 'max_pops', 'max_c_load', 'max_c_min4', 'max_c_inp', 'max_c_rel']
for v in range(2):
    for p in range(2):
        d = {names[i]: st[(v * 2 + p) * 16 + i] for i in range(16)}
        print('view', v, 'outer' if p == 0 else 'main ', {k: (f'{x/1e6:.2f}Mcyc' if k.startswith('c_') or k.startswith('max_c') or k == 'max_total' else x) for k, x in d.items()})
