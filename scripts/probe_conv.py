"""Accuracy (vs fp64 CPU) and speed of the two convolution kernels (FFMA direct vs tcgen05 3xTF32)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from reconfigisp_b200 import ops

def timeit(fn, iters=10, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

g = torch.Generator().manual_seed(1)
for (Cin, Cout, K) in ((64, 64, 3), (15, 64, 9), (64, 32, 5), (32, 3, 5), (64, 32, 1), (4, 64, 3), (64, 4, 3)):
    xs = torch.randn(2, Cin, 40, 136, generator=g); w = torch.randn(Cout, Cin, K, K, generator=g) / (K * Cin ** 0.5); b = torch.randn(Cout, generator=g) * 0.1
    ref = F.conv2d(xs.double(), w.double(), b.double(), padding=K // 2)
    yd = ops.conv2d(xs.cuda(), w.cuda(), b.cuda())
    yt = ops.from_blocked(ops.conv2d_tc(ops.to_blocked(xs.cuda()), w.cuda(), b.cuda()), Cout)
    e_d = float((yd.cpu().double() - ref).abs().max()); e_t = float((yt.cpu().double() - ref).abs().max())
    N, H, W = 4, 256, 256
    x = torch.randn(N, Cin, H, W, device='cuda'); wc, bc = w.cuda(), b.cuda()
    xb = ops.to_blocked(x)
    t_d = timeit(lambda: ops.conv2d(x, wc, bc))
    t_t = timeit(lambda: ops.conv2d_tc(xb, wc, bc))
    fl = 2.0 * N * H * W * Cin * Cout * K * K
    print('%2d->%2d k%d  err direct %.2e  tc %.2e | direct %.3f ms (%.1f TF/s)  tc %.3f ms (%.1f TF/s fp32-equivalent)' %
          (Cin, Cout, K, e_d, e_t, t_d, fl / t_d / 1e9, t_t, fl / t_t / 1e9), flush=True)
