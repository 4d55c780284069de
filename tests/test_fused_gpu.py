"""GPU parity of the packed-fp32 fused pipeline kernels (csrc/risp_fused.cu): every pre-instantiated chain signature x
every demosaic, forward / MSE step / backward-with-upstream-dy, against (a) the CPU oracle and (b) this library's own
op-by-op path (demosaic kernel + chain kernels + loss kernel under autograd), on ragged widths (partial last strip, a
single strip, several strips), frame borders, per-image parameter rows, exact 0 / 1 pixels (inclusive clamp masks), gains
of exactly zero (the gain is folded into the polynomial: no division anywhere) and tone-curve knots outside [0,1].
Tolerances: outputs max-abs <= 1e-4; reduced gradients rel 2e-3 of the largest entry."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import isp_oracle as O   # noqa: E402  (the checker)

TOL = 1e-4
IDENT = [0.0] * 30
IDENT[6] = IDENT[17] = IDENT[28] = 1.0

SIGS = {
    'A': ['gain', 'poly10', 'gamma', ('gtm', 4)],
    'B': ['gamma', 'poly10', 'gain'],
    'C': ['gamma', 'poly10'],
    'D': ['gamma', ('gtm', 4)],
}
DM = {'nearest': O.demosaic_nearest, 'bilinear': O.demosaic_bilinear, 'malvar': lambda r: O.demosaic_laplacian(r, 1.0)}


def stage_params(st, g, jitter=0.03):
    vals = []
    for s in st:
        name = s if isinstance(s, str) else s[0]
        if name == 'gain':
            vals += [1.1, 0.9, 1.2]
        elif name == 'poly10':
            vals += (torch.tensor(IDENT) + torch.randn(30, generator=g) * jitter).tolist()
        elif name == 'gamma':
            vals += [0.55]
        elif name == 'gtm':
            vals += [0.2, 0.5, 0.8]
    return torch.tensor([vals])


def oracle_chain(x, st, p):
    """p: (N, P) kernel-level parameter rows."""
    N, o = x.shape[0], 0
    for s in st:
        name = s if isinstance(s, str) else s[0]
        if name == 'gain':
            x = O.wb_manual(x, p[:, o:o + 3]); o += 3
        elif name == 'poly10':
            x = O.wb_quadratic(x, (p[:, o:o + 30] + 5) / 10); o += 30
        elif name == 'gamma':
            x = O.gamma_manual(x, p[:, o:o + 1]); o += 1
        elif name == 'gtm':
            x = torch.cat([O.gtm_manual(x[i:i + 1], p[i:i + 1, o:o + 3], 4) for i in range(N)]); o += 3
    return x


def relclose(a, b, rtol=2e-3, atol=1e-6):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    assert a.shape == b.shape
    lim = atol + rtol * float(b.abs().max())
    assert float((a - b).abs().max()) <= lim, (float((a - b).abs().max()), lim)


@pytest.fixture(scope='module')
def ops():
    import reconfigisp_b200.ops as ops
    return ops


@pytest.mark.parametrize('kind', ['nearest', 'bilinear', 'malvar'])
@pytest.mark.parametrize('sig', sorted(SIGS))
@pytest.mark.parametrize('shape', [(2, 12, 132), (1, 6, 8), (3, 50, 520)])
def test_fused_signatures_vs_oracle(ops, kind, sig, shape):
    N, H, W = shape
    g = torch.Generator().manual_seed(sum(map(ord, kind + sig)) + shape[1] * shape[2])
    raw = torch.rand(N, 1, H, W, generator=g) * 1.05
    raw.view(-1)[:3] = torch.tensor([0., 1., 0.5])
    gt = torch.rand(N, 3, H, W, generator=g)
    st = SIGS[sig]
    chain = ops.Chain(st)
    params = stage_params(st, g)
    # the oracle runs in float64 here (fp64 as the tie-breaker): near-black pixels go through x^gamma with
    # a slope of ~100, which turns the fp32 ORACLE's own rounding (3e-5 against fp64 on these inputs) into > 1e-4
    po = params.double().requires_grad_()
    yo = oracle_chain(DM[kind](raw.double()), st, po.expand(N, -1))
    lo = O.mse(yo, gt.double())
    dpo, = torch.autograd.grad(lo, po)
    # forward
    yg = ops.pipeline_fwd(raw.cuda(), kind, chain, params.cuda())
    assert float((yg.cpu().double() - yo.detach()).abs().max()) <= TOL
    # fused step, shared row
    pg = params.cuda().requires_grad_()
    lg = ops.pipeline_mse(pg, raw.cuda(), gt.cuda(), kind, chain)
    dpg, = torch.autograd.grad(lg, pg)
    assert abs(float(lg.detach()) - float(lo.detach())) <= 1e-5 * max(1.0, float(lo.detach()))
    relclose(dpg, dpo)
    # per-image rows
    pN = params.repeat(N, 1) * (1 + 0.01 * torch.arange(N).view(N, 1))
    poN = pN.double().requires_grad_()
    loN = O.mse(oracle_chain(DM[kind](raw.double()), st, poN), gt.double())
    doN, = torch.autograd.grad(loN, poN)
    pgN = pN.cuda().requires_grad_()
    lgN = ops.pipeline_mse(pgN, raw.cuda(), gt.cuda(), kind, chain)
    dgN, = torch.autograd.grad(lgN, pgN)
    relclose(dgN, doN)


@pytest.mark.parametrize('kind', ['nearest', 'bilinear', 'malvar'])
def test_fused_demosaic_only_is_the_demosaic(ops, kind):
    """The demosaic entry point runs through the fused kernel (chain = skip): bit-exact index work for nearest, oracle
    tolerance for the interpolating kinds, on a width with a partial last strip and the smallest legal frame."""
    for (N, H, W) in [(2, 14, 260), (1, 4, 4), (1, 40, 1028)]:
        g = torch.Generator().manual_seed(H * W)
        raw = torch.rand(N, 1, H, W, generator=g)
        y = ops.demosaic(raw.cuda(), kind).cpu()
        ref = DM[kind](raw)
        if kind == 'nearest':
            assert torch.equal(y, ref)
        else:
            assert float((y - ref).abs().max()) <= 1e-6


def test_fused_step_equals_op_by_op_path(ops):
    """The single-pass step against this library's unfused kernels under autograd, at a size with many items per CTA."""
    N, H, W = 2, 96, 1000
    g = torch.Generator().manual_seed(21)
    raw, gt = torch.rand(N, 1, H, W, generator=g).cuda(), torch.rand(N, 3, H, W, generator=g).cuda()
    for sig, st in SIGS.items():
        chain = ops.Chain(st)
        p1 = stage_params(st, g).cuda().requires_grad_()
        l1 = ops.pipeline_mse(p1, raw, gt, 'bilinear', chain)
        d1, = torch.autograd.grad(l1, p1)
        p2 = p1.detach().clone().requires_grad_()
        y2 = ops.chain_apply(ops.demosaic(raw, 'bilinear'), chain, p2)
        l2 = ops.mse_loss(y2, gt)
        d2, = torch.autograd.grad(l2, p2)
        assert abs(float(l1.detach()) - float(l2.detach())) <= 1e-6
        relclose(d1, d2, rtol=1e-3)
        # bit-reproducible: fixed summation order
        l3 = ops.pipeline_mse(p1, raw, gt, 'bilinear', chain)
        d3, = torch.autograd.grad(l3, p1)
        assert torch.equal(d1, d3) and torch.equal(l1, l3)


def test_fused_zero_gain_and_out_of_range_knots(ops):
    """gain = 0 exactly (folded gain: gradients still exact, no division) and tone-curve knots outside [0,1] (the final
    clamp of GtmManual becomes active: the kernel's block-uniform slow path)."""
    N, H, W = 1, 20, 136
    g = torch.Generator().manual_seed(5)
    raw = torch.rand(N, 1, H, W, generator=g)
    gt = torch.rand(N, 3, H, W, generator=g)
    st = SIGS['A']
    chain = ops.Chain(st)
    for gains, knots in (([0.0, 0.9, 1.2], [0.2, 0.5, 0.8]), ([1.1, 0.9, 1.2], [-0.2, 0.5, 1.3]), ([1.0, 1.0, 1.0], [0.6, 0.3, 0.9])):
        params = torch.tensor([gains + (torch.tensor(IDENT) + torch.randn(30, generator=g) * 0.03).tolist() + [0.55] + knots])
        po = params.clone().requires_grad_()
        lo = O.mse(oracle_chain(O.demosaic_bilinear(raw), st, po), gt)
        dpo, = torch.autograd.grad(lo, po)
        pg = params.cuda().requires_grad_()
        lg = ops.pipeline_mse(pg, raw.cuda(), gt.cuda(), 'bilinear', chain)
        dpg, = torch.autograd.grad(lg, pg)
        assert abs(float(lg.detach()) - float(lo.detach())) <= 1e-5
        relclose(dpg, dpo)


def test_fused_saturated_pixels_clamp_masks(ops):
    """Frames full of exact 0 / 1 raw values: polynomial outputs land exactly on the clamp bounds (inclusive masks) and the
    tone curve sees x == 1 (pass-through pixel)."""
    N, H, W = 1, 16, 128
    g = torch.Generator().manual_seed(9)
    raw = (torch.rand(N, 1, H, W, generator=g) > 0.5).float()
    gt = torch.rand(N, 3, H, W, generator=g)
    st = SIGS['A']
    chain = ops.Chain(st)
    params = torch.tensor([[1.0, 1.0, 1.0] + IDENT + [0.5] + [0.25, 0.5, 0.75]])
    po = params.clone().requires_grad_()
    yo = oracle_chain(O.demosaic_nearest(raw), st, po)
    lo = O.mse(yo, gt)
    dpo, = torch.autograd.grad(lo, po)
    yg = ops.pipeline_fwd(raw.cuda(), 'nearest', chain, params.cuda())
    assert float((yg.cpu() - yo.detach()).abs().max()) <= TOL
    pg = params.cuda().requires_grad_()
    lg = ops.pipeline_mse(pg, raw.cuda(), gt.cuda(), 'nearest', chain)
    dpg, = torch.autograd.grad(lg, pg)
    relclose(dpg, dpo)


@pytest.mark.parametrize('sig', sorted(SIGS))
def test_fused_l1_step(ops, sig):
    """pixel_criterion 'l1' (isp_model.py:44-49) through the same single pass: loss and gradients vs the fp64 oracle."""
    N, H, W = 2, 24, 264
    g = torch.Generator().manual_seed(77)
    raw, gt = torch.rand(N, 1, H, W, generator=g), torch.rand(N, 3, H, W, generator=g)
    st = SIGS[sig]
    chain = ops.Chain(st)
    assert ops.fused_chain_supported(chain)
    params = stage_params(st, g)
    po = params.double().requires_grad_()
    lo = (oracle_chain(O.demosaic_bilinear(raw.double()), st, po.expand(N, -1)) - gt.double()).abs().mean()
    dpo, = torch.autograd.grad(lo, po)
    pg = params.cuda().requires_grad_()
    lg = ops.pipeline_l1(pg, raw.cuda(), gt.cuda(), 'bilinear', chain)
    dpg, = torch.autograd.grad(lg, pg)
    assert abs(float(lg.detach()) - float(lo.detach())) <= 1e-5
    relclose(dpg, dpo)


def test_uint8_pack_and_crop_decode(ops):
    """tensor2bgr on the device (utils/util.py:118-135: clip(x*255, 0, 255) truncated, HWC), the u8 output of the blend
    (test_split.py:107) and the loader's even-aligned crop + /16383 (sid_sony_ratio_rggb2bgr_dataset.py:109-136): bit-exact."""
    import numpy as np
    g = torch.Generator().manual_seed(3)
    x = torch.rand(3, 37, 52, generator=g) * 1.2 - 0.1
    ref = np.clip(np.transpose(x.numpy(), [1, 2, 0]) * 255, 0, 255).astype(np.uint8)
    assert np.array_equal(ops.to_u8_hwc(x.cuda()).cpu().numpy(), ref)
    H, W, ps, st = 120, 136, 48, 40
    frame = torch.rand(3, H, W, generator=g) * 1.1 - 0.05
    tiles, _ = ops.whole2patch(frame.cuda(), (ps, ps), (st, st))
    merged = ops.patch2whole(tiles, (H, W), (st, st), clip01=True).cpu()
    u8, f = ops.patch2whole_u8(tiles, (H, W), (st, st), want_float=True)
    assert torch.equal(f.cpu(), merged)
    assert np.array_equal(u8.cpu().numpy(), (np.clip(np.transpose(merged.numpy(), [1, 2, 0]), 0, 1) * 255.).astype(np.uint8))
    codes = torch.randint(0, 16384, (2, 1, 60, 80), generator=g, dtype=torch.int32).to(torch.int16)
    crop = ops.crop_decode(codes.cuda(), 12, 34, 32, 40, 16383.)
    assert torch.equal(crop.cpu(), codes[:, :, 12:44, 34:74].float() / 16383.)
    with pytest.raises(ValueError):
        ops.crop_decode(codes.cuda(), 13, 34, 32, 40, 16383.)        # odd origin breaks the CFA phase


@pytest.mark.parametrize('kind', ['nearest', 'bilinear'])
def test_fused_tone_curve_pixels_exactly_on_knots(ops, kind):
    """The backward-carrying kernels look the tone curve's segment up from floor(4x) (FFMA.RZ / LOP3 / LDS.64 table).  Frames
    whose pixels sit EXACTLY on the knots 0, 1/4, 1/2, 1 (gamma = 1 passes powers of two through lg2 / ex2 unchanged) must pick
    the upper segment like the reference's half-open `(x >= start) & (x < end)` (tools_origin.py:435) and treat x == 1 as the
    pass-through pixel; the four segment slopes are made very different so that a wrong pick moves the gamma gradient."""
    N, H, W = 2, 12, 264
    g = torch.Generator().manual_seed(31)
    vals = torch.tensor([0.0, 0.25, 0.5, 1.0, 0.125, 0.0625])
    raw = vals[torch.randint(0, len(vals), (N, 1, H, W), generator=g)]
    gt = torch.rand(N, 3, H, W, generator=g)
    st = SIGS['D']
    chain = ops.Chain(st)
    params = torch.tensor([[1.0, 0.1, 0.7, 0.8]])
    po = params.double().requires_grad_()
    yo = oracle_chain(DM[kind](raw.double()), st, po.expand(N, -1))
    lo = O.mse(yo, gt.double())
    dpo, = torch.autograd.grad(lo, po)
    yg = ops.pipeline_fwd(raw.cuda(), kind, chain, params.cuda())
    assert float((yg.cpu().double() - yo.detach()).abs().max()) <= TOL
    pg = params.cuda().requires_grad_()
    lg = ops.pipeline_mse(pg, raw.cuda(), gt.cuda(), kind, chain)
    dpg, = torch.autograd.grad(lg, pg)
    assert abs(float(lg.detach()) - float(lo.detach())) <= 1e-5
    relclose(dpg, dpo, rtol=1e-3)
    # the same frames through the 37-parameter chain (identity polynomial, unit gains): the tone curve sits behind gamma there
    stA = SIGS['A']
    chainA = ops.Chain(stA)
    pA = torch.tensor([[1.0, 1.0, 1.0] + IDENT + [1.0, 0.1, 0.7, 0.8]])
    poA = pA.double().requires_grad_()
    loA = O.mse(oracle_chain(DM[kind](raw.double()), stA, poA.expand(N, -1)), gt.double())
    dpoA, = torch.autograd.grad(loA, poA)
    pgA = pA.cuda().requires_grad_()
    lgA = ops.pipeline_mse(pgA, raw.cuda(), gt.cuda(), kind, chainA)
    dpgA, = torch.autograd.grad(lgA, pgA)
    assert abs(float(lgA.detach()) - float(loA.detach())) <= 1e-5
    relclose(dpgA, dpoA, rtol=1e-3)
