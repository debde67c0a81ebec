"""debug helper (not a test): device-resident throughput of one process on one GPU"""
import os, sys, time
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'video-stereo-converter_b200')]
import numpy as np, torch
dev = int(sys.argv[1]) if len(sys.argv) > 1 else 0
torch.cuda.set_device(dev)
if os.environ.get('WORLD_SIZE'):
    dev = int(os.environ['LOCAL_RANK']); torch.cuda.set_device(dev)
    import torch.distributed as dist
    dist.init_process_group('nccl', device_id=torch.device('cuda', dev)); dist.barrier()
    print('max conn', os.environ.get('CUDA_DEVICE_MAX_CONNECTIONS'), flush=True)
from vsc_b200 import StereoGenerator, StereoParams
from vsc_b200.synthetic import make_pair
h, w, slots, n = 1080, 1920, 30, 96
g = StereoGenerator(f'cuda:{dev}', slots)
ND = int(os.environ.get('ND', '8')); S0 = int(os.environ.get('S0', '0'))
frames = [make_pair(h, w, S0 + i) for i in range(ND)]
d_rgb = [torch.from_numpy(r).cuda() for r, _ in frames]; d_dep = [torch.from_numpy(d).cuda() for _, d in frames]
d_out = [torch.empty((h, 2 * w, 3), dtype=torch.uint8, device='cuda') for _ in range(slots)]
def run(n):
    infl = []; tsub = 0.0
    for i in range(n):
        s = i % slots
        if len(infl) == slots: g.wait(infl.pop(0))
        t = time.perf_counter()
        g.submit_device(s, d_rgb[i % ND].data_ptr(), d_dep[i % ND].data_ptr(), np.uint8, h, w, d_out[s].data_ptr(), StereoParams()); infl.append(s)
        tsub += time.perf_counter() - t
    while infl: g.wait(infl.pop(0))
    return tsub
run(60)
for rep in range(2):
    g.timer_begin(); t0 = time.perf_counter(); tsub = run(n); ms = g.timer_end()
    print(f'dev {dev} pid {os.getpid()} fps {n/(ms*1e-3):.1f} submit-cpu-ms/frame {tsub/n*1e3:.3f} OMP={os.environ.get("OMP_NUM_THREADS")} cpus={os.cpu_count()} aff={len(os.sched_getaffinity(0))}', flush=True)
