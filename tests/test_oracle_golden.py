"""The CPU oracle against the golden vectors recorded from the reference's own code
(`oracle/gen_golden.py`).  This is what "pins" the oracle."""
import numpy as np
import torch

from oracle import isp_oracle as O
from oracle import pipeline_oracle as PO

T = torch.from_numpy


def close(a, b, tol=1e-6, rel=False):
    a = a.detach().numpy() if torch.is_tensor(a) else np.asarray(a)
    assert a.shape == b.shape, (a.shape, b.shape)
    scale = max(1.0, float(np.abs(b).max())) if rel else 1.0
    assert np.abs(a - b).max() <= tol * scale, np.abs(a - b).max()


def test_wb_quadratic(golden):
    g = golden('wb_quadratic')
    x, p = T(g['x']).requires_grad_(), T(g['p']).requires_grad_()
    y = O.wb_quadratic(x, p)
    dx, dp = torch.autograd.grad(y, (x, p), T(g['dy']))
    close(y, g['y']); close(dx, g['dx'], 1e-5); close(dp, g['dp'], 1e-4)


def test_gtm_manual(golden):
    g = golden('gtm_manual')
    x, p = T(g['x']).requires_grad_(), T(g['p']).requires_grad_()
    y = O.gtm_manual(x, p)
    dx, dp = torch.autograd.grad(y, (x, p), T(g['dy']))
    close(y, g['y'], 0); close(dx, g['dx'], 0); close(dp, g['dp'], 1e-6)
    # the reference's __main__ smoke: 0.9 with knots [.3,.5,.7] -> 0.88 (tools_origin.py:807-820)
    assert abs(float(g['smoke'].max()) - 0.88) < 1e-6 and abs(float(g['smoke'].min()) - 0.88) < 1e-6
    close(O.gtm_manual(torch.full((1, 3, 4, 4), 0.9), torch.tensor([[.3, .5, .7]])), g['smoke'], 0)


def test_conditional_fc(golden):
    g = golden('cond_fc')
    img, flat = T(g['img']), T(g['flat']).requires_grad_()
    hist = O.histc_planes(img, 4)
    close(hist, g['hist'], 0)
    out = O.fc_params(hist, flat, [12, 5, 3])
    dflat, = torch.autograd.grad(out, flat, torch.ones_like(out))
    close(out, g['out']); close(dflat, g['dflat'], 1e-5)


def test_cnn_candidates(golden):
    g = golden('cnn_candidates')
    x3, raw = T(g['x3']).requires_grad_(), T(g['raw']).requires_grad_()
    nets = [('srcnn_res3', PO.Net('srcnn_res', 3, 10)), ('srcnn_res1', PO.Net('srcnn_res', 1, 11)),
            ('srcnn_demosaic', PO.Net('srcnn_demosaic', 0, 12)), ('path14l_bayer', PO.Net('path14l_bayer', 0, 13)),
            ('path14l_bgr', PO.Net('path14l_bgr', 0, 14))]
    for name, net in nets:
        if name.startswith('srcnn_res'):
            y = net(x3, T(g[name + '_par'])); inp = x3
        elif name == 'path14l_bgr':
            y = net(x3); inp = x3
        else:
            y = net(raw); inp = raw
        dx, = torch.autograd.grad(y.square().sum(), inp)
        close(y, g[name + '_y'], 2e-6); close(dx, g[name + '_dx'], 1e-5, rel=True)


def _load_supernet(g):
    net = PO.Supernet(int(g['n_step']), float(g['threshold']), int(g['weight_seed']))
    with torch.no_grad():
        for i, a in enumerate(net.alphas):
            a.copy_(T(g['alpha%d' % i]))
        for k in range(int(g['n_step'])):
            for (name, _), lg in zip(PO.SRGB_STEP, net.logits[2 + k]):
                if lg.numel():
                    lg.copy_(T(g['logit_param_step%d_%s' % (k + 1, name)]))
    return net


def test_supernet(golden):
    g = golden('supernet')
    net = _load_supernet(g)
    y, inter = net.forward(T(g['raw']))
    assert list(g['pruned']) == net.pruned_paths and sum(net.pruned_paths) > 0
    close(y, g['y'], 1e-6)
    for i, m in enumerate(inter):
        close(m, g['inter%d' % i], 1e-6)
    loss = ((y - T(g['gt'])) ** 2).mean()
    nz = [l for l in net.trainable if l.numel()]
    grads = torch.autograd.grad(loss, net.alphas + nz, allow_unused=True)
    for i in range(len(net.alphas)):
        close(grads[i], g['dalpha%d' % i], 1e-6)
    names = ['param_step%d_%s' % (k + 1, n) for k in range(int(g['n_step'])) for n, _ in PO.SRGB_STEP
             if n in PO.DEFAULT_LOGITS]
    for n, gr in zip(names, grads[len(net.alphas):]):
        ref = g['dlogit_' + n]
        close(torch.zeros_like(T(ref)) if gr is None else gr, ref, 1e-6)


def test_fixed_pipelines(golden):
    g = golden('fixed_pipelines')
    raw = T(g['raw'])
    for tag, arch in (('classical', 'Bayer_02_Demosaic_02_sRGB_11_13_01'), ('sid', 'Bayer_01_Demosaic_03_sRGB_01_13_11'),
                      ('s7isp', 'Bayer_01_Demosaic_01_sRGB_04_01_13'),
                      ('all_origin', 'Bayer_02_Demosaic_01_sRGB_05_02_03_04_06_07_08_10_12_15'),
         ('nlm', 'Bayer_02_Demosaic_01_sRGB_09_01')):
        pipe = PO.FixedPipeline(arch, 'origin', 10)
        y, inter = pipe.forward(raw)
        close(y, g[tag + '_y'], 2e-6)
        for i, m in enumerate(inter):
            close(m, g['%s_inter%d' % (tag, i)], 2e-6)
        assert [k for k, l in zip(pipe.keys, pipe.logits) if l.numel()] == list(g[tag + '_keys'])
    pipe = PO.FixedPipeline('Bayer_02_Demosaic_01_sRGB_11_13_01_14_05', 'isp', 10)
    y, _ = pipe.forward(raw)
    close(y, g['isp_y'], 1e-6)
    loss = ((y - T(g['isp_gt'])) ** 2).mean()
    nz = [(k, l) for k, l in zip(pipe.keys, pipe.logits) if l.numel()]
    grads = torch.autograd.grad(loss, [l for _, l in nz])
    assert [k for k, _ in nz] == list(g['isp_keys'])
    for (k, _), gr in zip(nz, grads):
        close(gr, g['isp_dlogit_' + k], 1e-6)


def test_patch_split_merge(golden):
    g = golden('patch')
    patches, pos, cnt = O.whole2patch(g['img'], (16, 16), (12, 12))
    assert np.array_equal(pos, g['pos']) and np.array_equal(patches, g['patches']) and np.array_equal(cnt, g['cnt'])
    merged = O.patch2whole(patches * 0.5 + 0.1, pos, cnt, (12, 12))
    assert np.array_equal(merged, g['merged'])
    assert np.array_equal(O.create_patch_mask((16, 16), (2, 2)), g['mask'])
    assert np.array_equal(O.create_patch_mask((512, 512), (16, 16))[:40, :40], g['mask512'])
    assert O.patch_origins(3000, 512, 480) == list(g['ys']) and O.patch_origins(4000, 512, 480) == list(g['xs'])
    assert len(g['ys']) * len(g['xs']) == 63


def test_bayer_masks_bit_exact():
    H, W = 6, 10
    mR, mG1, mG2, mB = O.bayer_masks(H, W)
    raw = torch.arange(H * W, dtype=torch.float32).view(1, 1, H, W)
    p = O.pack_rggb(raw)
    assert torch.equal(p[0, 0], raw[0, 0][mR].view(H // 2, W // 2))
    assert torch.equal(p[0, 1], raw[0, 0][mG1].view(H // 2, W // 2))
    assert torch.equal(p[0, 2], raw[0, 0][mG2].view(H // 2, W // 2))
    assert torch.equal(p[0, 3], raw[0, 0][mB].view(H // 2, W // 2))
    assert torch.equal(O.unpack_rggb(p), raw)
    assert (mR.int() + mG1.int() + mG2.int() + mB.int()).eq(1).all()


def test_anchors_identity_at_default_logits():
    """In-repo anchors: the default logits make wbmanual / wbquadratic / gtm ~identity."""
    x = torch.rand(1, 3, 8, 8)
    sg = lambda v: torch.sigmoid(torch.tensor([v]))
    assert (PO.m_wbmanual(x, sg([-1.38] * 3)) - x).abs().max() < 6e-3
    assert (PO.m_wbquadratic(x, sg(PO.WBQ_INIT)) - x).abs().max() < 2e-3
    assert (PO.m_gtmmanual(x, sg([-1.099, 0, 1.099])) - x).abs().max() < 1e-3


def test_darts_step_against_reference_model(golden):
    """`PO.darts_step` (the restatement the GPU test of round 1 compared with) held to a run of the reference's own
    `DartsModel.optimize_alphas / optimize_parameters` (oracle/gen_golden_darts.py), n_step = 3, pruning active."""
    g = golden('darts_model')
    n_al = 5
    oG, oV = PO.Supernet(3, 0.2, 10), PO.Supernet(3, 0.2, 10)
    with torch.no_grad():
        for i, a in enumerate(oG.alphas):
            a.copy_(T(g['alpha0_%d' % i]))
    r = PO.darts_step(oG, oV, T(g['img']), T(g['gt']), T(g['vimg']), T(g['vgt']), lr_G=0.01, momentum_G=0.9, lr_meta=0.01)
    assert abs(float(r['val_loss'].detach()) - float(g['it0_val_loss'])) <= 1e-6
    for i in range(n_al):
        close(r['alpha_grad'][i], g['it0_alpha_grad_%d' % i], 2e-6)
    # the reference applies the Adam step on the alphas before optimize_parameters: repeat the plain pass there
    with torch.no_grad():
        for i, a in enumerate(oG.alphas):
            a.copy_(T(g['it0_alpha_%d' % i]))
    y, _ = oG.forward(T(g['img']))
    loss = ((y - T(g['gt'])) ** 2).mean()
    assert abs(float(loss.detach()) - float(g['it0_loss'])) <= 1e-6
    nz = [p for p in oG.trainable if p.numel() > 0]
    grads = torch.autograd.grad(loss, nz, allow_unused=True)
    assert len(nz) == int(g['n_params'])
    for i, gr in enumerate(grads):
        close(torch.zeros_like(nz[i]) if gr is None else gr, g['it0_param_grad_%d' % i], 2e-6)


def _summary(t):
    f = t.detach().reshape(-1).double()
    head = f[:24].float() if f.numel() >= 24 else torch.nn.functional.pad(f.float(), (0, 24 - f.numel()))
    return torch.cat([head, torch.tensor([float(f.sum()), float(f.abs().sum())])])


def test_finetune_step_against_reference_model(golden):
    """The oracle's proxy / target / loss pieces replayed with the reference's host RNG draws against a run of the
    reference's own `DartsFtModel.finetune_proxies` (oracle/gen_golden_darts_ft.py): first-step loss and weight gradients of
    every flagged proxy.  The FIFO holds the sRGB intermediates of the training pass (darts_ft_model.py:194-201)."""
    import random
    g = golden('darts_ft')
    names = [str(n) for n in g['names']]
    assert names == ['crysisengine', 'whiteworld', 'bilateral', 'median', 'fastnlm']
    oG = PO.Supernet(3, 0.2, 10)
    y, mids = oG.forward(T(g['img']))
    assert abs(float(((y - T(g['gt'])) ** 2).mean().detach()) - float(g['loss_G'])) <= 1e-6
    mem = [t.detach() for t in mids if t.shape[1] == 3]
    assert len(mem) == int(g['n_ft_data'])
    step_names = [n for n, _ in PO.SRGB_STEP]
    random.seed(int(g['ft_seed'])); torch.manual_seed(int(g['ft_seed']))
    for k, name in enumerate(names):
        net = oG.steps[-1][step_names.index(name)]
        data = mem[int(random.random() * len(mem))]
        param = torch.rand(1, net.P).repeat(data.shape[0], 1)
        sd = {kk: v.clone().requires_grad_() for kk, v in net.sd.items()}
        loss = O.mse(O.srcnn_res(data, param, sd), PO.ORIGIN[name](data, param))
        assert abs(float(loss.detach()) - float(g['losses'][k, 0])) <= 2e-6 * max(1.0, float(g['losses'][k, 0])), name
        grads = torch.autograd.grad(loss, list(sd.values()))
        ref = T(g['grads_' + name])[0]
        for i, gr in enumerate(grads):
            sm, rf = _summary(gr), ref[i].numpy()
            close(sm[:24], rf[:24], 2e-5 * max(1e-3, float(np.abs(rf[:24]).max())))       # first values of the tensor
            close(sm[24:], rf[24:], 2e-5 * max(1e-3, float(rf[25])))                        # its sum and abs-sum
        # consume the second step's draws so that the next proxy sees the generator state the reference saw
        int(random.random() * len(mem)); torch.rand(1, net.P)
