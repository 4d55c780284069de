import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reconfigisp_b200.modules.super_prune_fifteen_demos_four_bayer_two import mixed_op_probe
def timed(fn, steps, warm):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    return 1.0
print(mixed_op_probe(timed, 1, 3000, 4000, 6532.2))
