"""Device-side timing of the fused pipeline kernels (step / fwd) per signature and demosaic; RISP_FUSED=0 gives the old kernels."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reconfigisp_b200 import ops

PEAK = 6532.2
N, H, W = int(os.environ.get('PROBE_N', 4)), 3000, 4000


def timeit(fn, iters=20, warm=3):
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
    dev = 'cuda'
    g = torch.Generator(device='cpu').manual_seed(3)
    raw = torch.rand(N, 1, H, W, generator=g).to(dev)
    gt = torch.rand(N, 3, H, W, generator=g).to(dev)
    ident = [0.0] * 30
    ident[6] = ident[17] = ident[28] = 1.0
    P = {'gain': [1.05, 1.0, 0.95], 'poly10': ident, 'gamma': [0.5], 'gtm': [0.25, 0.5, 0.75]}
    sigs = {'A_gain_poly_gamma_gtm': ['gain', 'poly10', 'gamma', ('gtm', 4)], 'B_gamma_poly_gain': ['gamma', 'poly10', 'gain'],
            'C_gamma_poly': ['gamma', 'poly10'], 'D_gamma_gtm': ['gamma', ('gtm', 4)]}
    res = {}
    px = N * H * W
    kinds = sys.argv[1:] or ['bilinear', 'nearest', 'malvar']
    for name, st in sigs.items():
        chain = ops.Chain(st)
        vals = []
        for s in st:
            vals += P[s if isinstance(s, str) else s[0]]
        params = torch.tensor([vals], device=dev)
        for kind in kinds:
            step = ops.PipelineStep(N, H, W, kind, chain, dev)
            ms = timeit(lambda: step(raw, gt, params))
            msf = timeit(lambda: ops.pipeline_fwd(raw, kind, chain, params))
            res['%s/%s' % (name, kind)] = dict(step_ms=round(ms, 4), step_frac=round(16 * px / ms / 1e6 / PEAK, 3),
                                               fwd_ms=round(msf, 4), fwd_frac=round(16 * px / msf / 1e6 / PEAK, 3))
            print(name, kind, res['%s/%s' % (name, kind)], flush=True)
    for kind in kinds:
        ms = timeit(lambda: ops.demosaic(raw, kind))
        res['demosaic/' + kind] = dict(ms=round(ms, 4), frac=round(16 * px / ms / 1e6 / PEAK, 3))
        print('demosaic', kind, res['demosaic/' + kind], flush=True)
    os.makedirs('gpurun_out', exist_ok=True)
    tag = os.environ.get('PROBE_TAG', 'fused' + os.environ.get('RISP_FUSED', '1'))
    json.dump(res, open('gpurun_out/probe_step_%s.json' % tag, 'w'), indent=1)


if __name__ == '__main__':
    main()
