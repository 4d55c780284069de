import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reconfigisp_b200 import ops
Cin, Cout, K = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (64, 64, 3)))
x = torch.randn(4, Cin, 256, 256, device='cuda'); w = torch.randn(Cout, Cin, K, K, device='cuda') * 0.05; b = torch.randn(Cout, device='cuda')
xb = ops.to_blocked(x)
for _ in range(3):
    y = ops.conv2d_tc(xb, w, b)
torch.cuda.synchronize()
