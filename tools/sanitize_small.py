"""small end-to-end run for compute-sanitizer (a few parameter sets, small frames)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'video-stereo-converter_b200')]
import numpy as np
from vsc_b200 import StereoGenerator, StereoParams
from vsc_b200.synthetic import make_pair
g = StereoGenerator('cuda', 2)
cases = [((72, 128), np.uint8, {}), ((65, 131), np.uint16, dict(super_sampling=2.5, edge_softness=3.0)),
         ((64, 100), np.float32, dict(super_sampling=1.0, edge_softness=0.0, depth_gamma=1.0, max_disparity=30.0, convergence=-14.0, artifact_smoothing=5.0))]
for shape, dt, kw in cases:
    rgb, d = make_pair(shape[0], shape[1], 3, dt)
    out = g.process_frame(rgb, d, StereoParams(**kw))
    print(shape, kw, out.shape, int(out.sum()))
outs = g.process_batch([make_pair(72, 128, s) for s in range(4)])
print('batch ok', len(outs))
g.close()
