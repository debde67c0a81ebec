"""debug helper: single-frame latency for a few parameter sets"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'video-stereo-converter_b200')]
import numpy as np
from vsc_b200 import StereoGenerator, StereoParams
from vsc_b200.synthetic import make_pair
g = StereoGenerator('cuda', 1)
cases = [((1080, 1920), np.uint8, {}),
         ((1080, 1920), np.uint8, dict(edge_softness=0.0)),
         ((1080, 1920), np.uint8, dict(max_disparity=100.0, convergence=-50.0, super_sampling=1.0, edge_softness=0.0, depth_gamma=1.0, artifact_smoothing=5.0)),
         ((1080, 1920), np.uint8, dict(max_disparity=100.0, convergence=-50.0, super_sampling=2.0, edge_softness=0.0, depth_gamma=1.0, artifact_smoothing=5.0)),
         ((3840, 7680), np.uint16, dict(max_disparity=100.0, convergence=-50.0, super_sampling=1.0, edge_softness=0.0, depth_gamma=1.0, artifact_smoothing=5.0)),
         ((3840, 7680), np.uint16, dict(max_disparity=100.0, convergence=0.0, super_sampling=2.0, edge_softness=20.0, depth_gamma=0.2, artifact_smoothing=5.0))]
for shape, dt, kw in cases:
    rgb, d = make_pair(shape[0], shape[1], 0, dt)
    g.process_frame(rgb, d, StereoParams(**kw))
    g.set_profiling(True); g.process_frame(rgb, d, StereoParams(**kw)); g.set_profiling(False)
    kt = {}
    for n, t in g.kernel_times(0): kt[n] = kt.get(n, 0) + t
    print(shape, kw, 'frame ms %.1f' % g.last_frame_ms(0), {k: round(v, 2) for k, v in kt.items() if v > 0.3}, flush=True)
