"""Measures BASELINE.json's secondary configs on one B200 and writes gpurun_out/configs.json (copied to profiles/):
  cfg2  fixed pipeline, 12 MP frames, N in {1,4,8}: fused inference and the fused tuning step
  cfg2e the SAME pipeline written the way the reference writes it (one eager torch op after another + autograd)
        on the same B200 -- SURVEY.md §8d(iv) "the GPU baseline to beat".  Plain torch, independent of oracle/.
  cfg3  supernet search iteration (n_step=3, nothing pruned), 256^2, N=4 and N=64
  cfg4  12 MP frame -> 63 tiles of 512 (stride 480) -> pipeline on all tiles in one batch -> ramp blend
Not the bench contract (bench.py is)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from reconfigisp_b200 import ops, patch
from reconfigisp_b200.synthetic import synthetic_frames

PEAK = json.load(open('MEASURED_PEAKS.json')).get('hbm_gbs', 6532.2) if os.path.exists('MEASURED_PEAKS.json') else 6532.2


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


# ---- eager torch version of Bayer_02_Demosaic_02_sRGB_11_13_01_14 (reference style: tools_origin.py ops) --------
def eager_pipeline(raw, p_gain, p_poly, p_gamma, p_gtm):
    N, _, H, W = raw.shape
    yy, xx = torch.meshgrid(torch.arange(H, device=raw.device), torch.arange(W, device=raw.device), indexing='ij')
    r_m = ((yy % 2 == 0) & (xx % 2 == 0)).float(); b_m = ((yy % 2 == 1) & (xx % 2 == 1)).float(); g_m = 1 - r_m - b_m
    kg = torch.tensor([[0, 1, 0], [1, 4, 1], [0, 1, 0]], device=raw.device).float().view(1, 1, 3, 3) / 4
    krb = torch.tensor([[1, 2, 1], [2, 4, 2], [1, 2, 1]], device=raw.device).float().view(1, 1, 3, 3) / 4
    pad = lambda t: F.pad(t, (1, 1, 1, 1), mode='reflect')
    R = F.conv2d(pad(raw * r_m), krb); G = F.conv2d(pad(raw * g_m), kg); B = F.conv2d(pad(raw * b_m), krb)
    x = torch.cat([B, G, R], dim=1)
    x = x * (p_gain * 5).view(1, 3, 1, 1)                                                  # WbManual
    b, g, r = x[:, 0:1], x[:, 1:2], x[:, 2:3]                                               # WbQuadratic (:313-359)
    phi = torch.cat([b * b, g * g, r * r, b * g, b * r, g * r, b, g, r, torch.ones_like(b)], dim=1)
    P = (p_poly * 10 - 5).view(3, 10)
    x = torch.clamp(torch.einsum('ck,nkhw->nchw', P, phi), 0, 1)
    x = torch.clamp(x, 1e-6, 1) ** p_gamma.view(1, 1, 1, 1)                                 # Gamma
    xs = torch.linspace(0, 1, 5, device=raw.device)                                         # GtmManual (:409-440)
    ys = torch.cat([torch.zeros(1, device=raw.device), p_gtm, torch.ones(1, device=raw.device)])
    out = x
    for k in range(4):
        seg = (x >= xs[k]) & (x < xs[k + 1])
        out = torch.where(seg, ys[k] + (x - xs[k]) * (ys[k + 1] - ys[k]) / (xs[k + 1] - xs[k]), out)
    return torch.clamp(out, 0, 1)


def main():
    res = {'peak_GBps': PEAK}
    dev = 'cuda'
    H, W = 3000, 4000
    ident = [0.0] * 30
    ident[6] = ident[17] = ident[28] = 1.0
    params = torch.tensor([[1.05, 1.0, 0.95] + ident + [0.5] + [0.25, 0.5, 0.75]], device=dev)
    chain = ops.Chain(['gain', 'poly10', 'gamma', ('gtm', 4)])
    for N in (1, 4, 8):
        raw, gt = torch.rand(N, 1, H, W, device=dev), torch.rand(N, 3, H, W, device=dev)
        px = N * H * W
        step = ops.PipelineStep(N, H, W, 'bilinear', chain, dev)
        for name, fn in (('step', lambda: step(raw, gt, params)), ('fwd', lambda: ops.pipeline_fwd(raw, 'bilinear', chain, params))):
            ms = timeit(fn, iters=20)
            res['cfg2_%s_N%d' % (name, N)] = dict(ms=round(ms, 4), MPps=round(px / ms / 1e3, 1), frac_hbm=round(16 * px / ms / 1e6 / PEAK, 3))
            print('cfg2', name, N, res['cfg2_%s_N%d' % (name, N)], flush=True)
        del raw, gt, step
    # eager torch baseline on the same GPU (N=1: the unfused graph keeps ~25 full-size tensors alive)
    raw, gt = torch.rand(1, 1, H, W, device=dev), torch.rand(1, 3, H, W, device=dev)
    lg = [torch.full((3,), 0.2, device=dev, requires_grad=True), torch.tensor([(v + 5) / 10 for v in ident], device=dev, requires_grad=True),
          torch.tensor([0.5], device=dev, requires_grad=True), torch.tensor([0.25, 0.5, 0.75], device=dev, requires_grad=True)]

    def eager_step():
        loss = F.mse_loss(eager_pipeline(raw, *lg), gt)
        return torch.autograd.grad(loss, lg)

    def eager_fwd():
        with torch.no_grad():
            return eager_pipeline(raw, *lg)
    px = H * W
    for name, fn in (('step', eager_step), ('fwd', eager_fwd)):
        ms = timeit(fn, iters=5, warm=2)
        res['cfg2e_eager_torch_%s_N1' % name] = dict(ms=round(ms, 3), MPps=round(px / ms / 1e3, 1))
        print('cfg2e', name, res['cfg2e_eager_torch_%s_N1' % name], flush=True)
    del raw, gt
    torch.cuda.empty_cache()
    # cfg4: split inference
    frame = torch.rand(1, 1, H, W, device=dev)
    fn4 = lambda: patch.split_inference(lambda t: ops.pipeline_fwd(t, 'bilinear', chain, params), frame, 512, 480)
    ms = timeit(fn4, iters=10)
    res['cfg4_split_63tiles'] = dict(ms=round(ms, 3), MPps=round(H * W / ms / 1e3, 1))
    print('cfg4', res['cfg4_split_63tiles'], flush=True)
    # cfg3: search iteration
    from reconfigisp_b200.search import DartsModel
    for N, iters in ((4, 4), (64, 2)):
        opt = {'model': 'darts', 'network_G': {'which_model_G': 'SuperPruneFifteenDemosFourBayerTwo', 'n_step': 3, 'n_modules': 15,
                                               'prune_threshold': 0.2, 'weight_seed': 10},
               'train': {'lr_G': 1e-3, 'momentum_G': 0.9, 'lr_meta': 1e-3, 'beta1': 0.9, 'beta2': 0.999, 'pixel_criterion': 'l2'}}
        m = DartsModel(opt)
        r, g = synthetic_frames(2 * N, 256, 256, seed=10, pin=True)
        m.feed_data((r[:N], g[:N], r[N:], g[N:]))

        def it():
            m.optimize_alphas(); m.optimize_parameters()
        ms = timeit(it, iters=iters, warm=3)      # the side streams' allocator pools settle in the first iterations
        # 9.05 MFLOP/px fwd for the CNN candidates; fwd + data-gradient ~ 2x; 5 passes
        tf = 5 * 2 * 9.05e6 * N * 256 * 256 / (ms / 1e3) / 1e12
        res['cfg3_search_iter_N%d' % N] = dict(ms=round(ms, 1), iters_per_s=round(1e3 / ms, 3), MPps_per_pass=round(5 * N * 65536 / ms / 1e3, 2),
                                               cnn_TFLOPs_useful=round(tf, 1), pruned=m.netG.pruned_paths)
        print('cfg3', N, res['cfg3_search_iter_N%d' % N], flush=True)
        del m
        torch.cuda.empty_cache()
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(res, open('gpurun_out/configs.json', 'w'), indent=1)


if __name__ == '__main__':
    main()
