"""debug helper (not a test): per-seed frame latency and Telea stats (stats build)"""
import ctypes as C, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ['VSC_B200_LIB'] = os.path.join(ROOT, 'video-stereo-converter_b200', 'lib', 'libvsc_b200_stats.so')
sys.path[:0] = [os.path.join(ROOT, 'video-stereo-converter_b200')]
import numpy as np
from vsc_b200 import _lib, StereoGenerator
from vsc_b200.synthetic import make_pair
gen = StereoGenerator('cuda', 1); lib = _lib.load()
names = ['c_wait', 'c_pop', 'c_sort', 'c_part', 'c_total', 'n_pops', 'n_pix', 'n_gen', 'n_polls', 'n_clusters', 'max_total', 'max_pops']
for seed in [int(a) for a in sys.argv[1:]]:
    rgb, depth = make_pair(1080, 1920, seed=seed)
    gen.process_frame(rgb, depth)
    gen.set_profiling(True); gen.process_frame(rgb, depth)
    kt = dict(gen.kernel_times(0))
    st = (C.c_ulonglong * 64)(); _lib.check(lib.vsc_debug_telea_stats(gen._ctx.handle, st))
    print('seed', seed, 'frame ms %.1f' % gen.last_frame_ms(0), 'cluster ms %.1f' % kt['telea_cluster_kernel'])
    for v in range(2):
        for p in range(2):
            d = {names[i]: st[(v * 2 + p) * 16 + i] for i in range(12)}
            print('   v%d %s pix %d pops %d gens %d clusters %d total %.1fM max_total %.1fM max_pops %d sort %.1fM part %.1fM' % (
                v, 'outer' if p == 0 else 'main ', d['n_pix'], d['n_pops'], d['n_gen'], d['n_clusters'], d['c_total'] / 1e6, d['max_total'] / 1e6, d['max_pops'], d['c_sort'] / 1e6, d['c_part'] / 1e6))
