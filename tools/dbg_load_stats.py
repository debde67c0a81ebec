"""profiling helper (not a test): Telea phase counters of ONE group launch, solo vs under full multi-slot load (stats build)"""
import ctypes as C, sys, os
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault('VSC_B200_LIB', os.path.join(ROOT, 'video-stereo-converter_b200', 'lib', 'libvsc_b200_stats.so'))
sys.path[:0] = [os.path.join(ROOT, 'video-stereo-converter_b200')]
import numpy as np, torch
from vsc_b200 import _lib, StereoGenerator, StereoParams
from vsc_b200.synthetic import make_pair
h, w, slots, G = 1080, 1920, int(os.environ.get('SLOTS', '30')), 4
gen = StereoGenerator('cuda:0', slots, G)
lib = _lib.load()
fr = [make_pair(h, w, s) for s in range(8)]
dr = [torch.from_numpy(r).cuda() for r, _ in fr]; dd = [torch.from_numpy(d).cuda() for _, d in fr]
outs = [[torch.empty((h, 2 * w, 3), dtype=torch.uint8, device='cuda') for _ in range(G)] for _ in range(slots)]
def submit(s, k):
    gen.submit_device_group(s, [(dr[(k + i) % 8].data_ptr(), dd[(k + i) % 8].data_ptr(), outs[s][i].data_ptr()) for i in range(G)], np.uint8, h, w, StereoParams())
names = ['c_wait', 'c_pop', 'c_sort', 'c_part', 'c_total', 'n_pops', 'n_pix', 'n_gen', 'n_polls', 'n_clusters', 'max_total']
def show(tag):
    st = (C.c_ulonglong * 64)()
    _lib.check(lib.vsc_debug_telea_stats(gen._ctx.handle, st))
    tot = {n: 0 for n in names}
    for b in range(4):
        for i, n in enumerate(names):
            tot[n] = max(tot[n], st[b * 16 + i]) if n == 'max_total' else tot[n] + st[b * 16 + i]
    print(tag, {k: (f'{v/1e6:.1f}M' if k.startswith('c_') or k == 'max_total' else v) for k, v in tot.items()})
# solo: one group on slot 0
for _ in range(2):
    submit(0, 0); gen.wait(0)
show('solo  ')
# loaded: keep every slot busy, slot 0's launch in the middle of the stream of work
gen.timer_begin()
n = 0
for rnd in range(3):
    for s in range(slots):
        if rnd: gen.wait(s)
        submit(s, n); n += G
gen.wait(0)
show('loaded')
for s in range(slots): gen.wait(s)
ms = gen.timer_end()
print('fps %.1f' % (n / (ms * 1e-3)))
