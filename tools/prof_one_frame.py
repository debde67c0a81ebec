"""profiling driver (not a test): N identical-size frames through the device-resident path, one in flight"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'video-stereo-converter_b200')]
import numpy as np, torch
from vsc_b200 import StereoGenerator, StereoParams
from vsc_b200.synthetic import make_pair
h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1080, 1920)
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4
dt = np.uint16 if h > 1080 else np.uint8
gen = StereoGenerator('cuda', 1)
frames = [make_pair(h, w, seed=i, depth_dtype=dt) for i in range(2)]
d_rgb = [torch.from_numpy(r).cuda() for r, _ in frames]
d_dep = [torch.from_numpy(d).cuda() for _, d in frames]
d_out = torch.empty((h, 2 * w, 3), dtype=torch.uint8, device='cuda')
for i in range(n):
    gen.submit_device(0, d_rgb[i % 2].data_ptr(), d_dep[i % 2].data_ptr(), dt, h, w, d_out.data_ptr(), StereoParams())
    gen.wait(0)
print('ok', gen.last_frame_ms(0), 'ms/frame', gen.last_frame_launches(0), 'launches')
