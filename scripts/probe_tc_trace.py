"""Timeline of one CTA of the tcgen05 convolution (needs a -DRISP_TC_TRACE build: RISP_LIB_PATH=build/trace/lib.so)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reconfigisp_b200 import ops, _lib as L
L.call('risp_debug_tc_variant', int(os.environ.get('TC_VARIANT', '0')))
for Cin, Cout, K in ((64, 64, 3), (3, 64, 9), (64, 3, 9), (64, 32, 5), (32, 64, 5))[:int(os.environ.get('TC_CASES', '5'))]:
    x = torch.randn(4, Cin, 256, 256, device='cuda'); w = torch.randn(Cout, Cin, K, K, device='cuda') * 0.05; b = torch.randn(Cout, device='cuda')
    xb = ops.to_blocked(x)
    for _ in range(3):
        y = ops.conv2d_tc(xb, w, b)
    torch.cuda.synchronize()
    out = (ctypes.c_longlong * 16)()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        y = ops.conv2d_tc(xb, w, b)
    e1.record(); torch.cuda.synchronize()
    print('%.1f us per launch' % (e0.elapsed_time(e1) * 100))
    L.call('risp_debug_tc_trace', out)
    t = list(out)
    print('%2d->%2d k%d: prologue %d | stage loop %d | drain %d | epilogue %d | total %d ;  boundary-wait %d  weight-wait %d  issue %d  refill %d' % (
        Cin, Cout, K, t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[4] - t[0], t[8], t[9], t[10], t[11]), flush=True)
    print('   epilogue detail (thread 0): own work %d of %d, waiting for the other warps %d; TMEM loads %d' % (t[15] - t[3], t[4] - t[3], t[4] - t[15], t[7]))
