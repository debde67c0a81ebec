"""profiling driver (not a test): N slot submissions of G frames each through the device-resident path, one in flight
usage: python tools/prof_one_frame.py [H W [N [G]]]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'video-stereo-converter_b200')]
import numpy as np, torch
from vsc_b200 import StereoGenerator, StereoParams
from vsc_b200.synthetic import make_pair
h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1080, 1920)
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4
G = int(sys.argv[4]) if len(sys.argv) > 4 else 1
dt = np.uint16 if h > 1080 else np.uint8
gen = StereoGenerator('cuda', int(os.environ.get('SLOTS', '2')), G)    # >= 2 slots: the throughput configuration of the march
frames = [make_pair(h, w, seed=i, depth_dtype=dt) for i in range(max(2, G))]
d_rgb = [torch.from_numpy(r).cuda() for r, _ in frames]
d_dep = [torch.from_numpy(d).cuda() for _, d in frames]
d_out = [torch.empty((h, 2 * w, 3), dtype=torch.uint8, device='cuda') for _ in range(G)]
for i in range(n):
    tri = [(d_rgb[(i + k) % len(frames)].data_ptr(), d_dep[(i + k) % len(frames)].data_ptr(), d_out[k].data_ptr()) for k in range(G)]
    gen.submit_device_group(0, tri, dt, h, w, StereoParams())
    gen.wait(0)
print('ok', gen.last_frame_ms(0), 'ms/submission', gen.last_frame_launches(0), 'launches', G, 'frames')
