"""HBM fraction of the fused mixed-op kernels of one sRGB step in isolation (6 classical + 9 materialised candidates, 12 MP)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reconfigisp_b200.modules.super_prune_fifteen_demos_four_bayer_two import mixed_op_probe


def timed(fn, steps, warm):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


print(json.dumps(mixed_op_probe(timed, 1, 3000, 4000, 6532.2)))
