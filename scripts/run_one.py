"""Run a few launches of one fused pipeline kernel (for ncu):  python scripts/run_one.py step|fwd A|B|C|D bilinear [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reconfigisp_b200 import ops

mode, sig, kind = sys.argv[1], sys.argv[2], sys.argv[3]
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
N, H, W = int(os.environ.get('PROBE_N', 4)), 3000, 4000
g = torch.Generator(device='cpu').manual_seed(3)
raw = torch.rand(N, 1, H, W, generator=g).cuda()
gt = torch.rand(N, 3, H, W, generator=g).cuda()
ident = [0.0] * 30
ident[6] = ident[17] = ident[28] = 1.0
P = {'gain': [1.05, 1.0, 0.95], 'poly10': ident, 'gamma': [0.5], 'gtm': [0.25, 0.5, 0.75]}
st = {'A': ['gain', 'poly10', 'gamma', ('gtm', 4)], 'B': ['gamma', 'poly10', 'gain'], 'C': ['gamma', 'poly10'], 'D': ['gamma', ('gtm', 4)],
      'none': []}[sig]
chain = ops.Chain(st)
vals = []
for s in st:
    vals += P[s if isinstance(s, str) else s[0]]
params = torch.tensor([vals], device='cuda') if vals else None
if mode == 'step':
    step = ops.PipelineStep(N, H, W, kind, chain, 'cuda')
    for _ in range(iters):
        step(raw, gt, params)
elif mode == 'fwd':
    for _ in range(iters):
        ops.pipeline_fwd(raw, kind, chain, params) if st else ops.demosaic(raw, kind)
torch.cuda.synchronize()
print('done')
