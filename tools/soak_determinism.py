"""soak check (GPU box): every frame processed under full multi-slot load must equal, bit for bit, the same frame
processed alone (the march's dataflow must not depend on timing).  usage: python tools/soak_determinism.py [rounds]"""
import os, sys
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'video-stereo-converter_b200')]
import numpy as np, torch
from vsc_b200 import StereoGenerator, StereoParams
from vsc_b200.synthetic import make_pair
h, w, slots, G, ND = 1080, 1920, 30, 4, 24
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 6
fr = [make_pair(h, w, s) for s in range(ND)]
dr = [torch.from_numpy(r).cuda() for r, _ in fr]; dd = [torch.from_numpy(d).cuda() for _, d in fr]
solo = StereoGenerator('cuda:0', 1, 1)
ref = []
for i in range(ND):
    o = torch.empty((h, 2 * w, 3), dtype=torch.uint8, device='cuda')
    solo.submit_device(0, dr[i].data_ptr(), dd[i].data_ptr(), np.uint8, h, w, o.data_ptr(), StereoParams()); solo.wait(0)
    ref.append(o)
solo.close()
gen = StereoGenerator('cuda:0', slots, G)
outs = [[torch.empty((h, 2 * w, 3), dtype=torch.uint8, device='cuda') for _ in range(G)] for _ in range(slots)]
what = [[None] * G for _ in range(slots)]
bad = checked = 0
def check(s):
    global bad, checked
    gen.wait(s)
    for k in range(G):
        if what[s][k] is not None:
            checked += 1
            bad += int(not torch.equal(outs[s][k], ref[what[s][k]]))
n = 0
for rnd in range(rounds):
    for s in range(slots):
        if rnd: check(s)
        ids = [(n + k * 7) % ND for k in range(G)]; n += 1
        gen.submit_device_group(s, [(dr[i].data_ptr(), dd[i].data_ptr(), outs[s][k].data_ptr()) for k, i in enumerate(ids)], np.uint8, h, w, StereoParams())
        what[s] = ids
for s in range(slots): check(s)
print(f'soak: {checked} frames under load compared with their solo result, {bad} differ')
sys.exit(1 if bad else 0)
