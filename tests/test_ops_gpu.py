"""GPU parity: every C-ABI op (through reconfigisp_b200.ops) against the CPU oracle on the same
seeded inputs, plus the golden vectors recorded from the reference.  Tolerances: outputs in [0,1]
max-abs <= 1e-4 (north_star); reduced gradients rel 1e-3 / abs 1e-5; index work bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import isp_oracle as O   # noqa: E402  (the checker)

TOL = 1e-4


def dev(t):
    return t.cuda()


def maxabs(a, b):
    return float((a.detach().cpu().double() - b.detach().cpu().double()).abs().max())


def relclose(a, b, rtol=1e-3, atol=1e-5):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a - b).abs()
    lim = atol + rtol * b.abs().max()
    assert float(err.max()) <= float(lim), (float(err.max()), float(lim))


@pytest.fixture(scope='module')
def ops():
    import reconfigisp_b200.ops as ops
    return ops


def rand_img(N, H, W, seed, lo=-0.05, hi=1.1):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(N, 3, H, W, generator=g) * (hi - lo) + lo
    # exact 0 / 1 values exercise the inclusive clamp masks
    x.view(-1)[:4] = torch.tensor([0., 1., 0.5, 0.25])
    return x


SHAPES = [(2, 8, 12), (1, 17, 13), (3, 64, 96)]


@pytest.mark.parametrize('shape', SHAPES)
def test_pointwise_ops_fwd_bwd(ops, shape):
    N, H, W = shape
    g = torch.Generator().manual_seed(1)
    cases = {
        'gamma': (lambda x, p: O.gamma_manual(x, p), ops.gamma, torch.rand(N, 1, generator=g) * 0.9 + 0.05),
        'gain': (lambda x, p: O.wb_manual(x, p), ops.gain, torch.rand(N, 3, generator=g) * 2),
        'gain_clip': (lambda x, p: torch.clamp(O.wb_manual(x, p), 0, 1), ops.gain_clip, torch.rand(N, 3, generator=g) * 2),
        'poly10': (lambda x, p: O.wb_quadratic(x, (p + 5) / 10), ops.poly10, torch.randn(N, 30, generator=g) * 0.3),
        'ccm': (lambda x, p: O.ccm(x, p), ops.ccm, torch.eye(3).view(1, 9).repeat(N, 1) + torch.randn(N, 9, generator=g) * 0.2),
    }
    for name, (ofn, gfn, p) in cases.items():
        x = rand_img(N, H, W, 3)
        dy = torch.randn(N, 3, H, W, generator=g)
        xo, po = x.clone().requires_grad_(), p.clone().requires_grad_()
        yo = ofn(xo, po)
        dxo, dpo = torch.autograd.grad(yo, (xo, po), dy)
        xg, pg = dev(x).requires_grad_(), dev(p).requires_grad_()
        yg = gfn(xg, pg)
        dxg, dpg = torch.autograd.grad(yg, (xg, pg), dev(dy))
        assert maxabs(yg, yo) <= TOL, (name, maxabs(yg, yo))
        if name == 'gamma':
            # d/dx of x^g is unbounded near 0: compare where the oracle's own gradient is moderate
            m = (dxo.abs() < 50)
            assert float(((dxg.cpu() - dxo) * m).abs().max()) <= 1e-3 * 50, name
        else:
            relclose(dxg, dxo)
        relclose(dpg, dpo)


def test_gtm_and_goldens(ops, golden):
    g = golden('gtm_manual')
    x, p = torch.from_numpy(g['x']), torch.from_numpy(g['p'])
    xg = dev(x).requires_grad_()
    pg = dev(p).requires_grad_()
    # the module semantics (knots of batch element 0) are applied by the host layer: expand row 0
    y = ops.gtm(xg, pg[0:1].expand(x.shape[0], 3), 4)
    dx, dp = torch.autograd.grad(y, (xg, pg), dev(torch.from_numpy(g['dy'])))
    assert maxabs(y, torch.from_numpy(g['y'])) <= 1e-6
    assert maxabs(dx, torch.from_numpy(g['dx'])) <= 1e-5
    relclose(dp, torch.from_numpy(g['dp']))
    sm = ops.gtm(dev(torch.full((1, 3, 4, 4), 0.9)), dev(torch.tensor([[0.3, 0.5, 0.7]])), 4)
    assert abs(float(sm.max()) - 0.88) < 1e-6 and abs(float(sm.min()) - 0.88) < 1e-6

    g = golden('wb_quadratic')
    x, p = torch.from_numpy(g['x']), torch.from_numpy(g['p'])
    xg, pg = dev(x).requires_grad_(), dev(p).requires_grad_()
    y = ops.poly10(xg, pg * 10 - 5)
    dx, dp = torch.autograd.grad(y, (xg, pg), dev(torch.from_numpy(g['dy'])))
    assert maxabs(y, torch.from_numpy(g['y'])) <= 1e-5
    relclose(dx, torch.from_numpy(g['dx']))
    relclose(dp, torch.from_numpy(g['dp']))


def test_fused_chain_matches_sequence(ops):
    N, H, W = 2, 32, 48
    g = torch.Generator().manual_seed(5)
    x = rand_img(N, H, W, 7, 0.0, 1.0)
    gains = torch.rand(1, 3, generator=g) * 0.5 + 0.8
    coef = (torch.tensor(O_WBQ_IDENTITY()) + torch.randn(30, generator=g) * 0.05).view(1, 30)
    gm = torch.tensor([[0.6]])
    knots = torch.tensor([[0.2, 0.55, 0.8]])
    chain = ops.Chain(['gain', 'poly10', 'gamma', ('gtm', 4)])
    params = torch.cat([gains, coef, gm, knots], dim=1)

    def oracle(x, params):
        a = O.wb_manual(x, params[:, 0:3].expand(N, 3))
        b = O.wb_quadratic(a, ((params[:, 3:33] + 5) / 10).expand(N, 30))
        c = O.gamma_manual(b, params[:, 33:34].expand(N, 1))
        return O.gtm_manual(c, params[:, 34:37].expand(N, 3), 4)
    xo, po = x.clone().requires_grad_(), params.clone().requires_grad_()
    yo = oracle(xo, po)
    dy = torch.randn(N, 3, H, W, generator=g)
    dxo, dpo = torch.autograd.grad(yo, (xo, po), dy)
    xg, pg = dev(x).requires_grad_(), dev(params).requires_grad_()
    yg = ops.chain_apply(xg, chain, pg)
    dxg, dpg = torch.autograd.grad(yg, (xg, pg), dev(dy))
    assert maxabs(yg, yo) <= TOL
    m = dxo.abs() < 50
    assert float(((dxg.cpu() - dxo) * m).abs().max()) <= 5e-2
    relclose(dpg, dpo, rtol=2e-3)


def O_WBQ_IDENTITY():
    v = [0.0] * 30
    v[6] = v[17] = v[28] = 1.0
    return v


@pytest.mark.parametrize('shape', [(2, 8, 12), (1, 64, 256), (2, 34, 132)])
def test_bayer_index_ops_bit_exact(ops, shape):
    N, H, W = shape
    raw = torch.arange(N * H * W, dtype=torch.float32).view(N, 1, H, W)
    p = ops.pack_rggb(dev(raw))
    assert torch.equal(p.cpu(), O.pack_rggb(raw))
    assert torch.equal(ops.unpack_rggb(p).cpu(), raw)
    x12 = torch.arange(N * 12 * (H // 2) * (W // 2), dtype=torch.float32).view(N, 12, H // 2, W // 2)
    assert torch.equal(ops.pixel_shuffle2(dev(x12)).cpu(), torch.nn.functional.pixel_shuffle(x12, 2))
    assert torch.equal(ops.pixel_unshuffle2(ops.pixel_shuffle2(dev(x12))).cpu(), x12)
    # nearest demosaic only moves samples: bit-exact
    g = torch.Generator().manual_seed(2)
    r = torch.rand(N, 1, H, W, generator=g)
    assert torch.equal(ops.demosaic(dev(r), 'nearest').cpu(), O.demosaic_nearest(r))


@pytest.mark.parametrize('shape', [(2, 8, 12), (1, 64, 256), (2, 34, 132), (1, 70, 520)])
def test_demosaic(ops, shape):
    N, H, W = shape
    g = torch.Generator().manual_seed(4)
    raw = torch.rand(N, 1, H, W, generator=g)
    assert maxabs(ops.demosaic(dev(raw), 'bilinear'), O.demosaic_bilinear(raw)) <= 1e-6
    assert maxabs(ops.demosaic(dev(raw), 'malvar', 1.0), O.demosaic_laplacian(raw, 1.0)) <= 2e-6
    assert maxabs(ops.demosaic(dev(raw * 255), 'malvar', 255.0), O.demosaic_laplacian(raw * 255, 255.0)) <= 1e-3
    # adjoint of nearest
    ro = raw.clone().requires_grad_()
    d = torch.randn(N, 3, H, W, generator=g)
    dro, = torch.autograd.grad(O.demosaic_nearest(ro), ro, d)
    rg = dev(raw).requires_grad_()
    drg, = torch.autograd.grad(ops.demosaic(rg, 'nearest'), rg, dev(d))
    assert maxabs(drg, dro) <= 1e-6


def test_blc_wb(ops):
    N, H, W = 2, 16, 24
    g = torch.Generator().manual_seed(6)
    raw = torch.rand(N, 1, H, W, generator=g)
    p = torch.cat([torch.rand(N, 1, generator=g) * 0.1, torch.rand(N, 4, generator=g) + 0.8], dim=1)
    ro, po = raw.clone().requires_grad_(), p.clone().requires_grad_()
    yo = O.bayer_wb(O.black_level(ro, po[:, 0:1]), po[:, 1:5])
    d = torch.randn(N, 1, H, W, generator=g)
    dro, dpo = torch.autograd.grad(yo, (ro, po), d)
    rg, pg = dev(raw).requires_grad_(), dev(p).requires_grad_()
    yg = ops.bayer_blc_wb(rg, pg)
    drg, dpg = torch.autograd.grad(yg, (rg, pg), dev(d))
    assert maxabs(yg, yo) <= 1e-6
    relclose(drg, dro); relclose(dpg, dpo)


@pytest.mark.parametrize('kind', ['nearest', 'bilinear', 'malvar'])
def test_fused_pipeline_fwd_and_step(ops, kind):
    N, H, W = 2, 40, 136
    g = torch.Generator().manual_seed(8)
    raw = torch.rand(N, 1, H, W, generator=g) * 0.9
    gt = torch.rand(N, 3, H, W, generator=g)
    chain = ops.Chain(['gain', 'poly10', 'gamma', ('gtm', 4)])
    params = torch.cat([torch.tensor([[1.1, 0.9, 1.2]]), (torch.tensor(O_WBQ_IDENTITY()) + torch.randn(30, generator=g) * 0.03).view(1, 30),
                        torch.tensor([[0.55]]), torch.tensor([[0.2, 0.5, 0.8]])], dim=1)
    dm = {'nearest': O.demosaic_nearest, 'bilinear': O.demosaic_bilinear, 'malvar': lambda r: O.demosaic_laplacian(r, 1.0)}[kind]

    def oracle(params):
        a = O.wb_manual(dm(raw), params[:, 0:3].expand(N, 3))
        b = O.wb_quadratic(a, ((params[:, 3:33] + 5) / 10).expand(N, 30))
        c = O.gamma_manual(b, params[:, 33:34].expand(N, 1))
        return O.gtm_manual(c, params[:, 34:37].expand(N, 3), 4)
    po = params.clone().requires_grad_()
    yo = oracle(po)
    lo = O.mse(yo, gt)
    dpo, = torch.autograd.grad(lo, po)
    yg = ops.pipeline_fwd(dev(raw), kind, chain, dev(params))
    assert maxabs(yg, yo) <= TOL
    pg = dev(params).requires_grad_()
    lg = ops.pipeline_mse(pg, dev(raw), dev(gt), kind, chain)
    dpg, = torch.autograd.grad(lg, pg)
    assert abs(float(lg.detach()) - float(lo.detach())) <= 1e-5 * max(1.0, float(lo.detach()))
    relclose(dpg, dpo, rtol=2e-3, atol=1e-6)
    # per-image parameter rows
    pN = params.repeat(N, 1) * (1 + 0.01 * torch.arange(N).view(N, 1))
    pgN = dev(pN).requires_grad_()
    lgN = ops.pipeline_mse(pgN, dev(raw), dev(gt), kind, chain)
    dgN, = torch.autograd.grad(lgN, pgN)

    def oracleN(pp):
        a = O.wb_manual(dm(raw), pp[:, 0:3])
        b = O.wb_quadratic(a, (pp[:, 3:33] + 5) / 10)
        c = O.gamma_manual(b, pp[:, 33:34])
        outs = [O.gtm_manual(c[i:i + 1], pp[i:i + 1, 34:37], 4) for i in range(N)]
        return torch.cat(outs)
    poN = pN.clone().requires_grad_()
    loN = O.mse(oracleN(poN), gt)
    doN, = torch.autograd.grad(loN, poN)
    relclose(dgN, doN, rtol=2e-3, atol=1e-6)


def test_stats_hist_kth(ops):
    N, H, W = 2, 33, 47
    x = rand_img(N, H, W, 9)
    st = ops.plane_stats(dev(x)).cpu()
    assert torch.equal(st[..., 0], x.amin(dim=(2, 3))) and torch.equal(st[..., 2], x.amax(dim=(2, 3)))
    assert maxabs(st[..., 1], x.mean(dim=(2, 3))) <= 1e-6
    for bins in (4, 16, 256):
        assert torch.equal(ops.histc01(dev(x), bins).cpu(), O.histc_planes(x, bins))
    k = torch.tensor([1, 200])
    assert torch.equal(ops.kth_largest(dev(x), k).cpu(), O.kth_largest_per_plane(x, k))
    xl = torch.rand(1, 3, 512, 768, generator=torch.Generator().manual_seed(1)) * 255
    k = torch.tensor([int(0.37 * 512 * 768)])
    assert torch.equal(ops.kth_largest(dev(xl), k).cpu(), O.kth_largest_per_plane(xl, k))
    lm = ops.loglum_mean(dev(x.clamp(0, 1) * 255), 1 / 255.).cpu()
    ref = torch.log(O._lum(x.clamp(0, 1)) + 1e-6).mean(dim=(1, 2))
    assert maxabs(lm, ref) <= 1e-4


def test_grayworld_and_tone_ops(ops):
    N, H, W = 2, 24, 40
    x = rand_img(N, H, W, 11, 0.0, 0.9)
    xo = x.clone().requires_grad_()
    yo = O.wb_grayworld(xo)
    d = torch.randn(N, 3, H, W, generator=torch.Generator().manual_seed(3))
    dxo, = torch.autograd.grad(yo, xo, d)
    xg = dev(x).requires_grad_()
    yg = ops.grayworld(xg)
    dxg, = torch.autograd.grad(yg, xg, dev(d))
    assert maxabs(yg, yo) <= TOL
    relclose(dxg, dxo)
    x255 = x * 255
    wp, mg = torch.tensor([0.5, 0.8]), torch.tensor([0.5, 0.3])
    assert maxabs(ops.tone_reinhard(dev(x255), dev(wp), dev(mg), 255.) / 255, O.tone_reinhard(x255, wp, mg) / 255) <= TOL
    assert maxabs(ops.tone_crysis(dev(x255), dev(mg), 255.) / 255, O.tone_crysis(x255, mg) / 255) <= TOL
    ex = torch.tensor([5.5, 2.0])
    assert maxabs(ops.tone_filmic(dev(x255), dev(wp), dev(ex), 255.) / 255, O.tone_filmic(x255, wp, ex) / 255) <= TOL
    ratio = torch.tensor([0.02, 0.5])
    assert maxabs(ops.whiteworld(dev(x255), dev(ratio), 255.) / 255, O.wb_whiteworld(x255, ratio.numpy()) / 255) <= TOL


def test_stencils(ops):
    N, H, W = 2, 37, 45
    x = rand_img(N, H, W, 13, 0.0, 1.0)
    x255 = x * 255
    for k in (3, 5, 9):
        assert torch.equal(ops.median(dev(x255), k).cpu(), O.denoise_median(x255, k)), k
    win = torch.tensor([3, 5], dtype=torch.int32)
    sc, ss = torch.tensor([30., 12.]), torch.tensor([3., 50.])
    yb = ops.bilateral(dev(x255), dev(win), dev(sc), dev(ss))
    assert maxabs(yb / 255, O.denoise_bilateral(x255, win, sc, ss) / 255) <= TOL
    assert maxabs(ops.guided_filter(dev(x), 2, 1e-2), O.guided_filter(x, 2, 1e-2)) <= TOL
    # non-local means: per-image sizes, small and large decay, and the p == 1 corner of the wrapper (17 / 17)
    for blk, srch, h in (([3, 3], [3, 3], [12., 60.]), ([5, 3], [7, 9], [25., 8.]), ([17, 1], [17, 5], [40., 3.])):
        blk, srch, h = torch.tensor(blk, dtype=torch.int32), torch.tensor(srch, dtype=torch.int32), torch.tensor(h)
        yn = ops.fastnlm(dev(x255), dev(blk), dev(srch), dev(h))
        assert maxabs(yn / 255, O.denoise_fastnlm(x255, blk, srch, h) / 255) <= TOL, (blk, srch)
    a = torch.tensor([[0.7], [1.5]])
    xo, ao = x.clone().requires_grad_(), a.clone().requires_grad_()
    yo = O.sharpen(xo, ao)
    d = torch.randn(N, 3, H, W, generator=torch.Generator().manual_seed(5))
    dxo, dao = torch.autograd.grad(yo, (xo, ao), d)
    xg, ag = dev(x).requires_grad_(), dev(a).requires_grad_()
    yg = ops.sharpen(xg, ag)
    dxg, dag = torch.autograd.grad(yg, (xg, ag), dev(d))
    assert maxabs(yg, yo) <= TOL
    relclose(dxg, dxo); relclose(dag, dao)


def test_alpha_prune_and_mixed_op(ops):
    g = torch.Generator().manual_seed(17)
    for K, thr in ((2, 0.2), (4, 0.5), (15, 0.2)):
        alpha = torch.randn(K, generator=g) * 1.5
        ao = alpha.clone().requires_grad_()
        post_o, npr = O.prune_probs(ao, thr)
        up = torch.randn(K, generator=g)
        dao, = torch.autograd.grad(post_o, ao, up)
        ag = dev(alpha).requires_grad_()
        cnt = torch.zeros(1, dtype=torch.int32, device='cuda')
        post_g = ops.alpha_prune(ag, thr, cnt)
        dag, = torch.autograd.grad(post_g, ag, dev(up))
        assert maxabs(post_g, post_o) <= 1e-6 and int(cnt.item()) == npr
        assert maxabs(dag, dao) <= 1e-6
    # sRGB-style stage: 6 classical candidates in registers + 3 materialised ones
    N, H, W = 2, 16, 24
    x = rand_img(N, H, W, 19, 0.0, 1.0)
    ext = [rand_img(N, H, W, 20 + i, 0.0, 1.0) for i in range(3)]
    chain = ops.Chain(['gamma', 'gain_clip', 'skip', 'gain', 'poly10', ('gtm', 4)])
    gm, gw = torch.tensor([[0.5]]), torch.tensor([[1.1, 0.95, 1.05]])
    gn = torch.tensor([[1.0, 1.2, 0.9]])
    coef = (torch.tensor(O_WBQ_IDENTITY()) + torch.randn(30, generator=g) * 0.05).view(1, 30)
    kn = torch.tensor([[0.3, 0.5, 0.7]])
    params = torch.cat([gm, gw, gn, coef, kn], dim=1)
    w = torch.softmax(torch.randn(9, generator=g), dim=0)
    w[2] = 0.0   # a pruned branch
    w = w / w.sum()

    def oracle(x, params, w, ext):
        p = params.expand(N, -1)
        outs = [O.gamma_manual(x, p[:, 0:1]), torch.clamp(O.wb_manual(x, p[:, 1:4]), 0, 1), x, O.wb_manual(x, p[:, 4:7]),
                O.wb_quadratic(x, (p[:, 7:37] + 5) / 10), O.gtm_manual(x, p[:, 37:40], 4)] + list(ext)
        y = 0
        for o, wk in zip(outs, w):
            if wk < 1e-9:
                continue
            y = y + o * wk
        return y
    xo, po, wo = x.clone().requires_grad_(), params.clone().requires_grad_(), w.clone().requires_grad_()
    eo = [e.clone().requires_grad_() for e in ext]
    yo = oracle(xo, po, wo, eo)
    d = torch.randn(N, 3, H, W, generator=g)
    go = torch.autograd.grad(yo, [xo, po, wo] + eo, d)
    xg, pg, wg = dev(x).requires_grad_(), dev(params).requires_grad_(), dev(w).requires_grad_()
    eg = [dev(e).requires_grad_() for e in ext]
    yg = ops.mixed_op(xg, chain, pg, wg, eg)
    gg = torch.autograd.grad(yg, [xg, pg, wg] + eg, dev(d))
    assert maxabs(yg, yo) <= TOL
    m = go[0].abs() < 50
    assert float(((gg[0].cpu() - go[0]) * m).abs().max()) <= 5e-2
    relclose(gg[1], go[1], rtol=2e-3)
    relclose(gg[2], go[2], rtol=2e-3)
    for a, b in zip(gg[3:], go[3:]):
        relclose(a, b)
    # Bayer-domain stage: skip + one materialised candidate
    xb = torch.rand(N, 1, H, W, generator=g)
    e1 = torch.rand(N, 1, H, W, generator=g)
    wb = torch.tensor([0.3, 0.7])
    xo, wo, e1o = xb.clone().requires_grad_(), wb.clone().requires_grad_(), e1.clone().requires_grad_()
    yo = e1o * wo[0] + xo * wo[1]
    d1 = torch.randn(N, 1, H, W, generator=g)
    go = torch.autograd.grad(yo, (xo, wo, e1o), d1)
    # candidate order of the Bayer step: [path_bayer (ext), skip] -> kernel order is [skip..., ext...]
    xg, e1g = dev(xb).requires_grad_(), dev(e1).requires_grad_()
    wg = dev(torch.tensor([0.7, 0.3])).requires_grad_()
    yg = ops.mixed_op(xg, ops.Chain(['skip']), None, wg, [e1g])
    gg = torch.autograd.grad(yg, (xg, wg, e1g), dev(d1))
    assert maxabs(yg, yo) <= 1e-6
    relclose(gg[0], go[0]); relclose(gg[1].flip(0), go[1]); relclose(gg[2], go[2])


def test_loss(ops):
    g = torch.Generator().manual_seed(23)
    y, gt = torch.rand(2, 3, 15, 21, generator=g), torch.rand(2, 3, 15, 21, generator=g)
    for fn, ofn in ((ops.mse_loss, lambda a, b: ((a - b) ** 2).mean()), (ops.l1_loss, lambda a, b: (a - b).abs().mean())):
        yo = y.clone().requires_grad_()
        lo = ofn(yo, gt)
        do, = torch.autograd.grad(lo * 3.0, yo)
        yg = dev(y).requires_grad_()
        lg = fn(yg, dev(gt))
        dg, = torch.autograd.grad(lg * 3.0, yg)
        assert abs(float(lg.detach()) - float(lo.detach())) <= 1e-6
        assert maxabs(dg, do) <= 1e-7


def test_patch_split_merge_bit_exact(ops, golden):
    g = golden('patch')
    img = torch.from_numpy(g['img']).permute(2, 0, 1).contiguous()          # (C,H,W)
    tiles, pos = ops.whole2patch(dev(img), (16, 16), (12, 12))
    assert np.array_equal(np.array(pos), g['pos'])
    assert np.array_equal(tiles.permute(0, 2, 3, 1).cpu().numpy(), g['patches'])
    merged = ops.patch2whole(tiles * 0.5 + 0.1, (37, 45), (12, 12))
    assert np.array_equal(merged.permute(1, 2, 0).cpu().numpy(), g['merged'])
    # the shipped geometry (SID_test.yml:18-19): 3000x4000, 512/480 -> 63 tiles; identity round trip
    frame = torch.rand(3, 3000, 4000, generator=torch.Generator().manual_seed(1))
    tiles, pos = ops.whole2patch(dev(frame), (512, 512), (480, 480))
    assert tiles.shape[0] == 63
    back = ops.patch2whole(tiles, (3000, 4000), (480, 480))
    assert maxabs(back, frame) <= 1e-6


def test_errors_map_to_reference_exceptions(ops):
    x = torch.rand(1, 3, 8, 8).cuda()
    with pytest.raises(ValueError):
        ops.Chain(['poly10', 'ccm'])
    with pytest.raises(RuntimeError):
        ops.gamma(torch.rand(1, 3, 8, 8), torch.rand(1, 1))          # CPU tensor: no fallback
    with pytest.raises(NotImplementedError):
        r = torch.rand(1, 1, 8, 8).cuda().requires_grad_()
        ops.demosaic(r, 'bilinear').sum().backward()
    with pytest.raises(ValueError):
        ops.demosaic(torch.rand(1, 1, 8, 6).cuda(), 'bilinear')      # W % 4 != 0


@pytest.mark.parametrize('cfg', [(3, 64, 64, 3), (2, 15, 64, 9), (2, 64, 32, 5), (2, 32, 3, 5), (2, 64, 32, 1), (1, 4, 64, 3), (2, 64, 4, 3), (1, 32, 12, 5)])
def test_conv2d_fwd_and_data_gradient(ops, cfg):
    """The dense convolution of the CNN candidates against torch's fp64 convolution on the CPU."""
    import torch.nn.functional as F
    N, Cin, Cout, K = cfg
    H, W = 23, 40
    g = torch.Generator().manual_seed(31 + K + Cin)
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, K, K, generator=g) * (1.0 / (K * Cin ** 0.5))
    b = torch.randn(Cout, generator=g) * 0.1
    res = torch.randn(N, Cout, H, W, generator=g)
    d = torch.randn(N, Cout, H, W, generator=g)
    for relu_in, relu_out, use_res, res_relu in ((False, False, False, False), (True, True, False, False), (False, False, True, True),
                                                 (True, False, True, False)):
        xo, ro = x.double().requires_grad_(), res.double().requires_grad_()
        yo = F.conv2d(torch.relu(xo) if relu_in else xo, w.double(), b.double(), padding=K // 2)
        if relu_out:
            yo = torch.relu(yo)
        if use_res:
            yo = yo + (torch.relu(ro) if res_relu else ro)
        go = torch.autograd.grad(yo, (xo, ro) if use_res else (xo,), d.double())
        xg, rg = dev(x).requires_grad_(), dev(res).requires_grad_()
        yg = ops.conv2d(xg, dev(w), dev(b), relu_in, relu_out, rg if use_res else None, res_relu)
        gg = torch.autograd.grad(yg, (xg, rg) if use_res else (xg,), dev(d))
        assert maxabs(yg, yo) <= 2e-5, (cfg, relu_in, relu_out, use_res)
        for a, bb in zip(gg, go):
            assert maxabs(a, bb) <= 5e-5, (cfg, relu_in, relu_out, use_res)


def test_decode_codes_matches_loader_normalisation(ops):
    g = torch.Generator().manual_seed(2)
    raw = torch.randint(0, 1024, (2, 1, 17, 23), generator=g, dtype=torch.int32)
    gt = torch.randint(0, 256, (2, 3, 17, 23), generator=g, dtype=torch.int32)
    assert torch.equal(ops.decode_codes(raw.to(torch.int16).cuda(), 1023.).cpu(), raw.float() / 1023.)
    assert torch.equal(ops.decode_codes(gt.to(torch.uint8).cuda(), 255.).cpu(), gt.float() / 255.)
    big = torch.randint(0, 16384, (1, 1, 64, 128), generator=g, dtype=torch.int32)
    assert torch.equal(ops.decode_codes(big.to(torch.int16).cuda(), 16383.).cpu(), big.float() / 16383.)


@pytest.mark.parametrize('cfg', [(3, 64, 64, 3), (2, 15, 64, 9), (2, 64, 32, 5), (2, 32, 3, 5), (2, 64, 32, 1), (1, 4, 64, 3), (2, 64, 4, 3), (1, 32, 12, 5)])
def test_conv2d_tensor_core_path(ops, cfg):
    """tcgen05 implicit-GEMM convolution (3-term TF32 split) against torch's fp64 convolution on the CPU."""
    import torch.nn.functional as F
    N, Cin, Cout, K = cfg
    for H, W in ((23, 40), (37, 300)):
        g = torch.Generator().manual_seed(77 + K + Cin)
        x = torch.randn(N, Cin, H, W, generator=g)
        w = torch.randn(Cout, Cin, K, K, generator=g) * (1.0 / (K * Cin ** 0.5))
        b = torch.randn(Cout, generator=g) * 0.1
        res = torch.randn(N, Cout, H, W, generator=g)
        d = torch.randn(N, Cout, H, W, generator=g)
        for relu_in, relu_out, use_res, res_relu in ((False, False, False, False), (True, True, False, False), (True, False, True, True)):
            xo, ro = x.double().requires_grad_(), res.double().requires_grad_()
            yo = F.conv2d(torch.relu(xo) if relu_in else xo, w.double(), b.double(), padding=K // 2)
            if relu_out:
                yo = torch.relu(yo)
            if use_res:
                yo = yo + (torch.relu(ro) if res_relu else ro)
            go = torch.autograd.grad(yo, (xo, ro) if use_res else (xo,), d.double())
            xg, rg = dev(x).requires_grad_(), dev(res).requires_grad_()
            yb = ops.conv2d_tc(ops.to_blocked(xg), dev(w), dev(b), relu_in, relu_out, ops.to_blocked(rg) if use_res else None, res_relu)
            yg = ops.from_blocked(yb, Cout)
            gg = torch.autograd.grad(yg, (xg, rg) if use_res else (xg,), dev(d))
            # the tensor core rounds its accumulator toward zero once per MMA: the error grows with the number of
            # chained MMAs (taps x channels / 8) and is RELATIVE to the accumulated magnitude
            assert maxabs(yg, yo) <= 4e-5 * max(1.0, float(yo.detach().abs().max())), (cfg, H, W, relu_in, relu_out, use_res, maxabs(yg, yo))
            for a, bb in zip(gg, go):
                err = (a.detach().cpu().double() - bb).abs()
                tol = 1.5e-4 * max(1.0, float(bb.abs().max()))
                if relu_out:
                    # an output within rounding distance of 0 can fall on the other side of the ReLU than in the fp64
                    # reference; that flips one mask bit and perturbs the gradient inside that unit's receptive field
                    # (each flipped unit touches K*K*Cin gradient elements: ~1e-5 flips per unit x 64 units x 81 taps)
                    assert float((err > tol).double().mean()) <= 1e-2, (cfg, H, W, float((err > tol).double().mean()))
                    assert float(err.median()) <= 0.05 * tol
                else:
                    assert float(err.max()) <= tol, (cfg, H, W, relu_in, relu_out, use_res, float(err.max()))


@pytest.mark.parametrize('cfg', [(2, 15, 64, 9), (2, 64, 32, 5), (2, 32, 3, 5), (3, 5, 20, 3), (2, 7, 9, 1)])
@pytest.mark.parametrize('path', ['direct', 'tc'])
def test_conv2d_weight_and_bias_gradient(ops, cfg, path):
    """Weight / bias gradients (proxy fine-tuning, darts_ft_model.py:206-246) against torch's fp64 autograd."""
    import torch.nn.functional as F
    N, Cin, Cout, K = cfg
    H, W = 37, 70
    g = torch.Generator().manual_seed(5 + K + Cin)
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, K, K, generator=g) * (1.0 / (K * Cin ** 0.5))
    b = torch.randn(Cout, generator=g) * 0.1
    d = torch.randn(N, Cout, H, W, generator=g) / (H * W) ** 0.5
    for relu_in, relu_out in ((False, False), (True, True)):
        wo, bo = w.double().requires_grad_(), b.double().requires_grad_()
        xo = x.double()
        yo = F.conv2d(torch.relu(xo) if relu_in else xo, wo, bo, padding=K // 2)
        if relu_out:
            yo = torch.relu(yo)
        gwo, gbo = torch.autograd.grad(yo, (wo, bo), d.double())
        wg, bg = dev(w).requires_grad_(), dev(b).requires_grad_()
        if path == 'direct':
            yg = ops.conv2d(dev(x), wg, bg, relu_in, relu_out)
        else:
            yg = ops.from_blocked(ops.conv2d_tc(ops.to_blocked(dev(x)), wg, bg, relu_in, relu_out), Cout)
        gw, gb = torch.autograd.grad(yg, (wg, bg), dev(d))
        tw = 2e-5 * max(1.0, float(gwo.abs().max()))
        # an output ReLU mask bit can flip for outputs within rounding distance of 0 (tensor-core path mostly)
        errw = (gw.cpu().double() - gwo).abs()
        assert float(errw.max()) <= (20 if relu_out else 1) * tw, (cfg, path, relu_in, float(errw.max()), tw)
        assert maxabs(gb, gbo) <= (20 if relu_out else 1) * 2e-5 * max(1.0, float(gbo.abs().max())), (cfg, path, relu_in)


@pytest.mark.parametrize('shape', [(2, 38, 136), (1, 5, 4), (1, 64, 520), (3, 9, 128)])
def test_stencil_marching_fast_paths(ops, shape):
    """W % 4 == 0 takes the register-marching kernels (risp_march.cuh): median 3x3 (replicate border) bit-exact, bilateral with
    per-image window 3 / 1 (reflect-101, circular support), unsharp mask 5x5 -- strips with a partial last warp, one-lane
    frames, chunk seams."""
    N, H, W = shape
    x = rand_img(N, H, W, 31, 0.0, 1.0)
    x255 = x * 255
    assert torch.equal(ops.median(dev(x255), 3).cpu(), O.denoise_median(x255, 3))
    for k in (5, 7):
        if H > k and W > k:
            assert torch.equal(ops.median(dev(x255), k).cpu(), O.denoise_median(x255, k)), k
    win = torch.tensor([3, 1, 3][:N], dtype=torch.int32)
    sc, ss = torch.tensor([30., 12., 80.][:N]), torch.tensor([3., 50., 1.5][:N])
    yb = ops.bilateral(dev(x255), dev(win), dev(sc), dev(ss), max_window=3)
    assert maxabs(yb / 255, O.denoise_bilateral(x255, win, sc, ss) / 255) <= TOL
    a = torch.tensor([[0.7], [1.5], [0.2]][:N])
    if H > 2:
        assert maxabs(ops.sharpen(dev(x), dev(a)), O.sharpen(x, a)) <= TOL


@pytest.mark.parametrize('shape', [(2, 37, 44), (1, 9, 8), (1, 130, 260), (2, 61, 45)])
@pytest.mark.parametrize('radius', [1, 3, 4, 8])
def test_guided_filter_marching_and_tile_kernels_match_oracle(ops, shape, radius):
    """W % 4 == 0 takes the register-marching kernels (running column sums + shuffle windows, strips with an 8-column halo,
    several strips / row chunks at the larger shape), other widths the shared-memory tile kernels; both against the oracle,
    including a flat region where var -> 0 (the cancellation-sensitive case of s2/n - mean^2)."""
    N, H, W = shape
    if min(H, W) <= radius:
        pytest.skip('window larger than the frame')
    x = rand_img(N, H, W, 31 + radius, 0.0, 1.0)
    x[:, :, : H // 3, : W // 2] = 0.37
    for eps in (1e-2, 1e-3):
        assert maxabs(ops.guided_filter(dev(x), radius, eps), O.guided_filter(x, radius, eps)) <= TOL, (shape, radius, eps)
