"""profiling helper (not a test): Telea phase counters for one 1080p default frame (stats build)"""
import ctypes as C, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault('VSC_B200_LIB', os.path.join(ROOT, 'video-stereo-converter_b200', 'lib', 'libvsc_b200_stats.so'))
sys.path[:0] = [os.path.join(ROOT, 'video-stereo-converter_b200')]
import numpy as np
import json
from vsc_b200 import _lib, StereoGenerator, StereoParams
from vsc_b200.synthetic import make_pair
h, w, dt = (int(sys.argv[1]), int(sys.argv[2]), np.uint16 if len(sys.argv) > 3 and sys.argv[3] == 'u16' else np.uint8) if len(sys.argv) > 2 else (1080, 1920, np.uint8)
gen = StereoGenerator('cuda', 1)
P = StereoParams(**json.loads(os.environ.get('VSC_PARAMS', '{}')))      # e.g. VSC_PARAMS='{"edge_softness": 0, "super_sampling": 1}'
lib = _lib.load()
rgb, depth = make_pair(h, w, seed=0, depth_dtype=dt)
for i in range(3):
    gen.process_frame(rgb, depth, P)
gen.set_profiling(True)
gen.process_frame(rgb, depth, P)
for n, t in gen.kernel_times(0):
    if t > 0.05: print(f'  {n:28s} {t:8.3f} ms')
st = (C.c_ulonglong * 64)()
_lib.check(lib.vsc_debug_telea_stats(gen._ctx.handle, st))
names = ['sum_total', 'n_clusters', 'sum_band', 'sum_outer', 'sum_order', 'sum_colour', 'tasks_holes', 'tasks_ring',
         'max_total', 'max_band', 'max_outer', 'max_order', 'max_colour', 'max_ntask', 'max_nband', 'max_gens_holes',
         'max_streams', 'max_sweeps', 'max_c_sort', 'max_c_claim', 'max_c_sweep', 'max_c_push',
         'b_load', 'b_terms', 'b_sums', 'b_final', 'b_release', 'b_handover', 'b_idle', 'b_tasks', 'b_continued']
cyc = {'sum_total', 'sum_band', 'sum_outer', 'sum_order', 'sum_colour', 'max_total', 'max_band', 'max_outer', 'max_order',
       'max_colour', 'max_c_sort', 'max_c_claim', 'max_c_sweep', 'max_c_push', 'b_load', 'b_terms', 'b_sums', 'b_final', 'b_release', 'b_handover', 'b_idle'}
for v in range(2):
    d = {names[i]: st[v * 32 + i] for i in range(len(names))}
    print('view', v, {k: (f'{x / 1e6:.3f}Mcyc' if k in cyc else x) for k, x in d.items()})
