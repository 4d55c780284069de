"""Tiny driver for ncu captures: a few launches of the fused pipeline step on 4 x 12MP frames."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reconfigisp_b200 import ops
N, H, W = int(os.environ.get('PROBE_N', 4)), 3000, 4000
kind = os.environ.get('PROBE_DM', 'bilinear')
raw = torch.rand(N, 1, H, W, device='cuda'); gt = torch.rand(N, 3, H, W, device='cuda')
ident = [0.0] * 30
ident[6] = ident[17] = ident[28] = 1.0
params = torch.tensor([[1.05, 1.0, 0.95] + ident + [0.5] + [0.25, 0.5, 0.75]], device='cuda')
chain = ops.Chain(['gain', 'poly10', 'gamma', ('gtm', 4)])
step = ops.PipelineStep(N, H, W, kind, chain, 'cuda')
for _ in range(6):
    step(raw, gt, params)
y = ops.pipeline_fwd(raw, kind, chain, params)
torch.cuda.synchronize()
print('done', float(step.loss))
