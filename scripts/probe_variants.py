"""Time the fused step / fwd kernels for each library variant given on the command line."""
import os, subprocess, sys
code = r'''
import sys, os, torch
sys.path.insert(0, os.getcwd())
from reconfigisp_b200 import ops
N, H, W = 4, 3000, 4000
raw = torch.rand(N, 1, H, W, device='cuda'); gt = torch.rand(N, 3, H, W, device='cuda')
ident = [0.0] * 30
ident[6] = ident[17] = ident[28] = 1.0
params = torch.tensor([[1.05, 1.0, 0.95] + ident + [0.5] + [0.25, 0.5, 0.75]], device='cuda')
chain = ops.Chain(['gain', 'poly10', 'gamma', ('gtm', 4)])
def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
out = []
for kind in ('nearest', 'bilinear', 'malvar'):
    step = ops.PipelineStep(N, H, W, kind, chain, 'cuda')
    ms = timeit(lambda: step(raw, gt, params))
    msf = timeit(lambda: ops.pipeline_fwd(raw, kind, chain, params))
    out.append('%s step %.3f ms (%.0f MP/s, %.3f of peak)  fwd %.3f ms' % (kind, ms, N*H*W/ms/1e3, 16*N*H*W/ms/1e6/6532.2, msf))
print(' | '.join(out))
'''
for lib in sys.argv[1:]:
    env = dict(os.environ)
    if lib != 'default':
        env['RISP_LIB_PATH'] = os.path.abspath(lib)
    r = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True)
    print(lib, '::', r.stdout.strip() or r.stderr[-400:], flush=True)
