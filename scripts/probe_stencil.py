"""Device timing of the stencil stages at 12 MP (24 B/px algorithmic: read x, write y) -> gpurun_out/probe_stencil.json."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reconfigisp_b200 import ops

PEAK = 6532.2
N, H, W = 1, 3000, 4000


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


x = torch.rand(N, 3, H, W, device='cuda')
x255 = x * 255
res = {}
px = N * H * W


def rec(name, ms, bpp=24):
    res[name] = dict(ms=round(ms, 3), MPps=round(px / ms / 1e3, 1), frac=round(bpp * px / ms / 1e6 / PEAK, 3))
    print(name, res[name], flush=True)


win = torch.tensor([3] * N, dtype=torch.int32, device='cuda')
rec('bilateral_w3', timeit(lambda: ops.bilateral(x255, win, torch.full((N,), 30., device='cuda'), torch.full((N,), 3., device='cuda'))))
for k in (3, 5, 9, 15):
    rec('median_k%d' % k, timeit(lambda: ops.median(x255, k), iters=3, warm=1))
rec('fastnlm_b3_s3', timeit(lambda: ops.fastnlm(x255, win, win, torch.full((N,), 20., device='cuda')), iters=3, warm=1))
rec('guided_r4', timeit(lambda: ops.guided_filter(x, 4, 1e-3)))
rec('sharpen', timeit(lambda: ops.sharpen(x, torch.full((N,), 0.5, device='cuda'))))
os.makedirs('gpurun_out', exist_ok=True)
json.dump(res, open('gpurun_out/probe_stencil.json', 'w'), indent=1)
