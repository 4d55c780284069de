"""3x3 64->64 tensor-core convolution: alternative tile configurations (risp_debug_tc_variant) at B = 4 and B = 64."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reconfigisp_b200 import ops, _lib as L
for B in (4, 64):
    x = torch.randn(B, 64, 256, 256, device='cuda'); w = torch.randn(64, 64, 3, 3, device='cuda') / 24.0; b = torch.zeros(64, device='cuda')
    xb = ops.to_blocked(x)
    ref = None
    for v in (0, 1, 2, 3):
        L.call('risp_debug_tc_variant', v)
        with torch.no_grad():
            for _ in range(3):
                y = ops.conv2d_tc(xb, w, b)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                y = ops.conv2d_tc(xb, w, b)
            e1.record(); torch.cuda.synchronize()
        ref = y if ref is None else ref
        us = e0.elapsed_time(e1) * 100
        print('B=%d variant %d: %.1f us  %.1f TF/s useful  maxdiff vs variant 0 %.1e' % (B, v, us, 2.0 * B * 65536 * 64 * 64 * 9 / us / 1e6, float((y - ref).abs().max())), flush=True)
L.call('risp_debug_tc_variant', 0)
