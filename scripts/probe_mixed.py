"""Fused mixed-op of one sRGB step in isolation at 12 MP (fwd 132 B/px, bwd 252 B/px): fraction of the measured HBM peak."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reconfigisp_b200.modules.super_prune_fifteen_demos_four_bayer_two import mixed_op_probe


def timed(fn, iters, warm):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)


peak = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs'] if os.path.exists('MEASURED_PEAKS.json') else 6532.2
for K in (9, 3):
    print(K, json.dumps(mixed_op_probe(timed, 1, 3000, 4000, peak, K_ext=K)), flush=True)
