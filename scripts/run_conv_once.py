import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reconfigisp_b200 import ops
B, S = 64, 256
x = torch.randn(B, 64, S, S, device='cuda'); w = torch.randn(64, 64, 3, 3, device='cuda') / 24.0; b = torch.zeros(64, device='cuda')
xb = ops.to_blocked(x)
with torch.no_grad():
    for _ in range(3): ops.conv2d_tc(xb, w, b)
torch.cuda.synchronize(); print('done')
