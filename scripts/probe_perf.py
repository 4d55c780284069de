"""Quick device-side timing probe of the hot kernels (not the bench contract; used while tuning)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reconfigisp_b200 import ops

PEAK = 6532.2


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    H, W = 3000, 4000
    px = N * H * W
    dev = 'cuda'
    raw = torch.rand(N, 1, H, W, device=dev)
    gt = torch.rand(N, 3, H, W, device=dev)
    x = torch.rand(N, 3, H, W, device=dev)
    dy = torch.randn(N, 3, H, W, device=dev)
    ident = [0.0] * 30
    ident[6] = ident[17] = ident[28] = 1.0
    params = torch.tensor([[1.05, 1.0, 0.95] + ident + [0.5] + [0.25, 0.5, 0.75]], device=dev)
    chain = ops.Chain(['gain', 'poly10', 'gamma', ('gtm', 4)])
    res = {}

    def rec(name, ms, bytes_per_px):
        gbs = bytes_per_px * px / ms / 1e6
        res[name] = dict(ms=round(ms, 4), MPps=round(px / ms / 1e3, 1), GBps=round(gbs, 1), frac=round(gbs / PEAK, 3))
        print(name, res[name], flush=True)

    for kind in ('nearest', 'bilinear', 'malvar'):
        step = ops.PipelineStep(N, H, W, kind, chain, dev)
        rec('step_' + kind, timeit(lambda: step(raw, gt, params)), 16)
        rec('fwd_' + kind, timeit(lambda: ops.pipeline_fwd(raw, kind, chain, params)), 16)
        rec('demosaic_' + kind, timeit(lambda: ops.demosaic(raw, kind)), 16)
    for name, ch, p in (('gamma', ops.Chain(['gamma']), params[:, 33:34]), ('gain', ops.Chain(['gain']), params[:, 0:3]),
                        ('poly10', ops.Chain(['poly10']), params[:, 3:33]), ('gtm', ops.Chain([('gtm', 4)]), params[:, 34:37]),
                        ('chain4', chain, params)):
        p = p.contiguous()
        rec('fwd_' + name, timeit(lambda: ops.chain_apply(x, ch, p)), 24)
        xg = x.clone().requires_grad_()
        pg = p.clone().requires_grad_()
        y = ops.chain_apply(xg, ch, pg)
        rec('bwd_' + name, timeit(lambda: torch.autograd.grad(y, (xg, pg), dy, retain_graph=True)), 36)
    rec('mse_fwd', timeit(lambda: ops.mse_loss(x, gt)), 24)
    rec('copy_torch', timeit(lambda: x.copy_(gt)), 24)
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(res, open('gpurun_out/probe_N%d.json' % N, 'w'), indent=1)


if __name__ == '__main__':
    main()
