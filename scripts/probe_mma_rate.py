import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reconfigisp_b200 import _lib as L
out = torch.zeros(2, dtype=torch.int64, device='cuda')
for NP in (16, 32, 64, 128, 256):
    for n_acc in (1, 2, 4):
        if n_acc * NP > 512: continue
        for split3 in (0, 1):
            for a_pw in (136, 137, 144):
                iters = 512
                L.call('risp_debug_mma_rate', L.ptr(out), NP, iters, n_acc, split3, a_pw, L.stream())
                torch.cuda.synchronize()
                n = iters * (3 if split3 else 1)
                iss, tot = out.tolist()
                print('N=%3d acc=%d split3=%d a_pw=%d : issue %.1f cyc/MMA  total %.1f cyc/MMA  (ideal %.0f)' % (NP, n_acc, split3, a_pw, iss / n, tot / n, 128 * NP * 8 / 2048.), flush=True)
