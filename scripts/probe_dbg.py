import os, sys, torch
sys.path.insert(0, os.getcwd())
from reconfigisp_b200 import ops
N,H,W=4,3000,4000
raw=torch.rand(N,1,H,W,device='cuda')
def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for k in ('nearest','bilinear'):
    print(os.environ.get('RISP_FUSED_DBG','0'), k, '%.4f ms' % timeit(lambda: ops.demosaic(raw,k)))
