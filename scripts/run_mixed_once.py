"""A few launches of the fused mixed-op of one sRGB step at 12 MP (for ncu):  ncu -k regex:mixed_ ... python scripts/run_mixed_once.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reconfigisp_b200.modules.super_prune_fifteen_demos_four_bayer_two import mixed_op_probe


def timed(fn, iters, warm):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    return 1.0


mixed_op_probe(timed, 1, 3000, 4000, 6532.2, K_ext=9)
torch.cuda.synchronize()
print('done')
