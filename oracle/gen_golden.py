"""Generate `tests/golden/*.npz` by RUNNING THE REFERENCE's own in-repo code on CPU.

    python -m oracle.gen_golden            (only in the build container: needs /root/reference)

Everything recorded here comes out of reference code objects (`models.modules.*`,
`utils.util_path_restore`) loaded by `oracle/ref_loader.py`; the five un-shipped kernel
modules are bound to `oracle/isp_oracle.py` (those ops are "parity unpinned", SPEC.md).
The vectors are small; the script is committed so they can be regenerated.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader as RL           # noqa: E402
from oracle import isp_oracle as O            # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def npy(t):
    return t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)


def save(name, **kw):
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **{k: npy(v) for k, v in kw.items()})
    print('wrote', name, {k: npy(v).shape for k, v in kw.items()})


def synth_bgr(N, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(N, 3, H, W, generator=g)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = RL.load_reference(weight_seed=10)
    T = ref.tools_origin
    g = torch.Generator().manual_seed(10)

    # ---- A5 WbQuadratic (tools_origin.py:313-359) -------------------------------------
    x = (torch.rand(2, 3, 6, 8, generator=g) * 1.2 - 0.1).requires_grad_()
    p = torch.rand(2, 30, generator=g)
    p = (0.5 + 0.08 * (p - 0.5)).requires_grad_()          # near identity so the clamp is partly active
    p.data[:, 6] = 0.6; p.data[:, 17] = 0.6; p.data[:, 28] = 0.6
    dy = torch.randn(2, 3, 6, 8, generator=g)
    y = T.WbQuadratic()(x, p)
    dx, dp = torch.autograd.grad(y, (x, p), dy)
    save('wb_quadratic', x=x, p=p, dy=dy, y=y, dx=dx, dp=dp)

    # ---- A6 GtmManual (tools_origin.py:409-440) + the __main__ smoke (:807-820) --------
    x = (torch.rand(3, 3, 5, 7, generator=g) * 1.4 - 0.2)
    x.view(-1)[:6] = torch.tensor([0., 0.25, 0.5, 0.75, 1.0, 0.9])   # exact knot hits / borders
    x.requires_grad_()
    p = torch.tensor([[0.3, 0.5, 0.7], [0.1, 0.2, 0.3], [0.9, 0.9, 0.9]], requires_grad=True)
    dy = torch.randn(3, 3, 5, 7, generator=g)
    y = T.GtmManual(4)(x, p)
    dx, dp = torch.autograd.grad(y, (x, p), dy)
    smoke = T.GtmManual(4)(torch.full((1, 3, 4, 4), 0.9), torch.tensor([[0.3, 0.5, 0.7]]))
    save('gtm_manual', x=x, p=p, dy=dy, y=y, dx=dx, dp=dp, smoke=smoke)

    # ---- A19 conditional-module FC (tools_origin.py:109-163) ---------------------------
    with RL.cpu_only():
        cm = T.ConditionalWbManual(in_channels=(12, 5))
        img = synth_bgr(2, 9, 10, 3) * 1.1 - 0.05
        flat = (torch.randn(cm.total_params, generator=g) * 0.1).requires_grad_()
        out = cm._fc_forward(img, flat)
        dflat, = torch.autograd.grad(out, flat, torch.ones_like(out))
        hist = O.histc_planes(img, 4)
    save('cond_fc', img=img, flat=flat, out=out, dflat=dflat, hist=hist, in_channels=np.array([12, 5]),
         out_channel=np.array(3))

    # ---- A15-A18 CNN candidates with seeded weights (arch files) ------------------------
    P = ref.tools_proxy
    with RL.cpu_only():
        ref.counter['n'] = 0
        nets = dict(srcnn_res3=P.ProxyNet(3, 'x'), srcnn_res1=P.ProxyNet(1, 'x'),
                    srcnn_demosaic=P.ProxyDemosaicNet(0, 'x'),
                    path14l_bayer=P.PathRestore14lBayer(0, 'x'), path14l_bgr=P.PathRestore14lBgr(0, 'x'))
        x3 = synth_bgr(2, 20, 24, 5).requires_grad_()
        raw = torch.rand(2, 1, 20, 24, generator=g).requires_grad_()
        outs = {}
        for name, net in nets.items():
            if name.startswith('srcnn_res'):
                par = torch.rand(2, int(name[-1]), generator=g)
                y = net(x3, par)
                dx, = torch.autograd.grad(y.square().sum(), x3)
                outs[name + '_par'] = par
            elif name == 'path14l_bgr':
                y = net(x3, None)
                dx, = torch.autograd.grad(y.square().sum(), x3)
            else:
                y = net(raw, None)
                dx, = torch.autograd.grad(y.square().sum(), raw)
            outs[name + '_y'] = y
            outs[name + '_dx'] = dx
    save('cnn_candidates', x3=x3, raw=raw, seeds=np.arange(10, 15), **outs)

    # ---- A22 supernet forward / alpha+param grads (super_prune…:175-214) ----------------
    with RL.cpu_only():
        ref.counter['n'] = 0
        net = ref.super_prune.SuperPruneFifteenDemosFourBayerTwo(n_step=2, threshold=0.2, module_path='/x/')
        ga = torch.Generator().manual_seed(11)
        with torch.no_grad():
            for a in net.alphas:
                a.copy_(torch.randn(a.shape, generator=ga) * 1.2)        # some paths get pruned
            for q in net.trainable_parameters:
                if q.numel():
                    q.add_(torch.randn(q.shape, generator=ga) * 0.3)
        raw = (torch.rand(2, 1, 16, 16, generator=ga) * 0.8 + 0.1)
        gt = synth_bgr(2, 16, 16, 12)
        y = net(raw)
        loss = ((y - gt) ** 2).mean()
        nz = [q for q in net.trainable_parameters if q.numel()]
        grads = torch.autograd.grad(loss, list(net.alphas) + nz, allow_unused=True)
        kw = {'alpha%d' % i: a for i, a in enumerate(net.alphas)}
        kw.update({'dalpha%d' % i: grads[i] for i in range(len(net.alphas))})
        names = [n for n, q in net.named_parameters() if n.startswith('param_')]
        for n, q, gr in zip(names, nz, grads[len(net.alphas):]):
            kw['logit_' + n] = q
            kw['dlogit_' + n] = gr if gr is not None else torch.zeros_like(q)
        for i, m in enumerate(net.intermediate_results):
            kw['inter%d' % i] = m
        save('supernet', raw=raw, gt=gt, y=y, loss=loss, pruned=np.array(net.pruned_paths),
             threshold=np.array(0.2), n_step=np.array(2), weight_seed=np.array(10), **kw)

    # ---- A20/A21 fixed pipelines (origin_universal.py / isp_universal.py) ---------------
    with RL.cpu_only():
        raw = (torch.rand(2, 1, 24, 32, generator=g) * 0.9)
        kw = dict(raw=raw)
        for tag, arch in (('classical', 'Bayer_02_Demosaic_02_sRGB_11_13_01'),
                          ('sid', 'Bayer_01_Demosaic_03_sRGB_01_13_11'),
                          ('s7isp', 'Bayer_01_Demosaic_01_sRGB_04_01_13'),
                          ('all_origin', 'Bayer_02_Demosaic_01_sRGB_05_02_03_04_06_07_08_10_12_15'),
                          ('nlm', 'Bayer_02_Demosaic_01_sRGB_09_01')):
            ref.counter['n'] = 0
            net = ref.origin_universal.OriginUniversal('/x/', arch)
            y = net(raw)
            kw[tag + '_y'] = y
            kw[tag + '_keys'] = np.array(list(net.state_dict().keys()))
            for i, m in enumerate(net.intermediate_results):
                kw['%s_inter%d' % (tag, i)] = m
        ref.counter['n'] = 0
        arch = 'Bayer_02_Demosaic_01_sRGB_11_13_01_14_05'
        net = ref.isp_universal.IspUniversal('/x/', (None,) * 7, arch)
        gt = synth_bgr(2, 24, 32, 13)
        y = net(raw)
        loss = ((y - gt) ** 2).mean()
        nz = [q for q in net.trainable_parameters if q.numel()]
        grads = torch.autograd.grad(loss, nz)
        kw.update(isp_y=y, isp_loss=loss, isp_gt=gt, isp_keys=np.array(list(net.state_dict().keys())))
        for q, gr, key in zip(nz, grads, net.state_dict().keys()):
            kw['isp_dlogit_' + key] = gr
        save('fixed_pipelines', **kw)

    # ---- A25 patch split / merge (util_path_restore.py:47-134) --------------------------
    U = ref.util_path_restore
    rng = np.random.RandomState(10)
    img = rng.rand(37, 45, 3).astype(np.float32)
    patches, pos, cnt = U.whole2patch(img, (16, 16), (12, 12))
    merged = U.patch2whole(patches * 0.5 + 0.1, pos, cnt, (12, 12))
    mask = U.create_patch_mask((16, 16), (2, 2))
    # the shipped test geometry: 3000x4000 frame, 512/480 (SID_test.yml:18-19): origins only
    H, W, h, s = 3000, 4000, 512, 480
    ys = list(range(0, H - h, s)) + [H - h]
    xs = list(range(0, W - h, s)) + [W - h]
    save('patch', img=img, patches=patches, pos=pos, cnt=cnt, merged=merged, mask=mask,
         mask512=U.create_patch_mask((512, 512), (16, 16))[:40, :40], ys=np.array(ys), xs=np.array(xs))


if __name__ == '__main__':
    main()
