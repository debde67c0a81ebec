"""debug helper: per-kernel elapsed times under full multi-slot load (profiling events on every slot)"""
import os, sys, collections
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'video-stereo-converter_b200')]
import numpy as np, torch
from vsc_b200 import StereoGenerator, StereoParams
from vsc_b200.synthetic import make_pair
h, w, slots = 1080, 1920, int(os.environ.get('SLOTS', '30'))
g = StereoGenerator('cuda:0', slots)
fr = [make_pair(h, w, s) for s in range(8)]
dr = [torch.from_numpy(r).cuda() for r, _ in fr]; dd = [torch.from_numpy(d).cuda() for _, d in fr]
d_out = [torch.empty((h, 2 * w, 3), dtype=torch.uint8, device='cuda') for _ in range(slots)]
acc = collections.defaultdict(list); frame_ms = []
def run(n, collect):
    free, busy = list(range(slots)), []
    for i in range(n):
        if not free:
            s = g.wait_any(busy); g.wait(s); busy.remove(s); free.append(s)
            if collect:
                per = collections.defaultdict(float)
                for name, t in g.kernel_times(s): per[name] += t
                for k, v in per.items(): acc[k].append(v)
                frame_ms.append(g.last_frame_ms(s))
        s = free.pop(0)
        g.submit_device(s, dr[i % 8].data_ptr(), dd[i % 8].data_ptr(), np.uint8, h, w, d_out[s].data_ptr(), StereoParams()); busy.append(s)
    for s in busy: g.wait(s)
run(60, False)
g.set_profiling(True)
g.timer_begin(); run(150, True); ms = g.timer_end()
print('slots', slots, 'fps %.1f' % (150 / (ms * 1e-3)), 'mean frame device ms %.1f' % np.mean(frame_ms))
for k, v in sorted(acc.items(), key=lambda kv: -np.mean(kv[1])):
    if np.mean(v) > 0.05: print(f'  {k:28s} mean {np.mean(v):8.3f} ms  max {np.max(v):8.3f}')
