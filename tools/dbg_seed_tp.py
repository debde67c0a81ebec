"""debug helper: 30-slot throughput when every frame is the same seed"""
import os, sys
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'video-stereo-converter_b200')]
import numpy as np, torch
from vsc_b200 import StereoGenerator, StereoParams
from vsc_b200.synthetic import make_pair
h, w, slots, n = 1080, 1920, int(os.environ.get('SLOTS', '30')), 60
g = StereoGenerator('cuda:0', slots)
d_out = [torch.empty((h, 2 * w, 3), dtype=torch.uint8, device='cuda') for _ in range(slots)]
for seed in [int(a) for a in sys.argv[1:]]:
    r, d = make_pair(h, w, seed)
    dr, dd = torch.from_numpy(r).cuda(), torch.from_numpy(d).cuda()
    def run(n):
        infl = []
        for i in range(n):
            s = i % slots
            if len(infl) == slots: g.wait(infl.pop(0))
            g.submit_device(s, dr.data_ptr(), dd.data_ptr(), np.uint8, h, w, d_out[s].data_ptr(), StereoParams()); infl.append(s)
        while infl: g.wait(infl.pop(0))
    run(30)
    g.timer_begin(); run(n); ms = g.timer_end()
    print(f'seed {seed} fps {n/(ms*1e-3):.1f}', flush=True)
