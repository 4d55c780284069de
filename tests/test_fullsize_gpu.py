"""Parity at BASELINE.json's full frame size (3000 x 4000, 12 MP), where the CPU oracle is too slow to run the
whole frame: size-independent properties plus oracle checks on crops cut out of the full-size result.
  * the fused kernels against the oracle on interior and border crops (the oracle runs on the crop + halo)
  * fused single-pass step == the unfused op-by-op path (same loss, same gradients)
  * bit-reproducibility (no float atomics), batch-mean consistency, adjointness of the nearest demosaic,
    bit-exact code normalisation and the 63-tile split / blend round trip at the shipped geometry."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import isp_oracle as O   # noqa: E402  (the checker)

H, W = 3000, 4000


@pytest.fixture(scope='module')
def ops():
    import reconfigisp_b200.ops as ops
    return ops


@pytest.fixture(scope='module')
def frame():
    from reconfigisp_b200.synthetic import synthetic_frames
    raw, gt = synthetic_frames(1, H, W, seed=10)
    return raw, gt


def _params():
    ident = [0.0] * 30
    ident[6] = ident[17] = ident[28] = 1.0
    g = torch.Generator().manual_seed(3)
    poly = torch.tensor(ident) + torch.randn(30, generator=g) * 0.02
    return torch.cat([torch.tensor([1.08, 0.97, 1.12]), poly, torch.tensor([0.55]), torch.tensor([0.22, 0.5, 0.81])]).view(1, 37)


def _oracle_pipeline(raw, params, kind):
    dm = {'nearest': O.demosaic_nearest, 'bilinear': O.demosaic_bilinear, 'malvar': lambda r: O.demosaic_laplacian(r, 1.0)}[kind]
    N = raw.shape[0]
    a = O.wb_manual(dm(raw), params[:, 0:3].expand(N, 3))
    b = O.wb_quadratic(a, ((params[:, 3:33] + 5) / 10).expand(N, 30))
    c = O.gamma_manual(b, params[:, 33:34].expand(N, 1))
    return O.gtm_manual(c, params[:, 34:37].expand(N, 3), 4)


def maxabs(a, b):
    return float((a.detach().cpu().double() - b.detach().cpu().double()).abs().max())


@pytest.mark.parametrize('kind', ['nearest', 'bilinear', 'malvar'])
def test_full_frame_inference_matches_oracle_on_crops(ops, frame, kind):
    raw, _ = frame
    chain = ops.Chain(['gain', 'poly10', 'gamma', ('gtm', 4)])
    params = _params()
    y = ops.pipeline_fwd(raw.cuda(), kind, chain, params.cuda()).cpu()
    assert y.shape == (1, 3, H, W)
    hl = 4                                                      # even halo >= the widest stencil, keeps the CFA phase
    for (y0, x0, h, w) in ((0, 0, 48, 64), (H - 48, W - 64, 48, 64), (1502, 1998, 40, 72), (0, W - 132, 36, 132), (H - 34, 2, 34, 60)):
        ya, xa = max(0, y0 - hl), max(0, x0 - hl)
        yb, xb = min(H, y0 + h + hl), min(W, x0 + w + hl)
        ref = _oracle_pipeline(raw[:, :, ya:yb, xa:xb], params, kind)
        # rows / columns of the crop whose stencil support is inside the crop or on the true frame border
        ref = ref[:, :, y0 - ya: y0 - ya + h, x0 - xa: x0 - xa + w]
        assert maxabs(y[:, :, y0:y0 + h, x0:x0 + w], ref) <= 1e-4, (kind, y0, x0)


def test_full_frame_step_equals_unfused_path_and_is_reproducible(ops, frame):
    raw, gt = frame
    raw, gt = raw.cuda(), gt.cuda()
    chain = ops.Chain(['gain', 'poly10', 'gamma', ('gtm', 4)])
    p1 = _params().cuda().requires_grad_()
    l1 = ops.pipeline_mse(p1, raw, gt, 'bilinear', chain)
    g1, = torch.autograd.grad(l1, p1)
    # op-by-op: demosaic kernel, one chain kernel per stage, loss kernel, autograd in between
    p2 = _params().cuda().requires_grad_()
    x = ops.demosaic(raw, 'bilinear')
    x = ops.gain(x, p2[:, 0:3])
    x = ops.poly10(x, p2[:, 3:33])
    x = ops.gamma(x, p2[:, 33:34])
    x = ops.gtm(x, p2[:, 34:37], 4)
    l2 = ops.mse_loss(x, gt)
    g2, = torch.autograd.grad(l2, p2)
    assert abs(float(l1.detach()) - float(l2.detach())) <= 2e-6 * float(l2.detach())
    err = (g1 - g2).abs().max()
    assert float(err) <= 1e-3 * float(g2.abs().max()), (float(err), float(g2.abs().max()))
    # fixed-order reductions: a second run is bit-identical
    p3 = _params().cuda().requires_grad_()
    l3 = ops.pipeline_mse(p3, raw, gt, 'bilinear', chain)
    g3, = torch.autograd.grad(l3, p3)
    assert torch.equal(l1.detach(), l3.detach()) and torch.equal(g1, g3)
    # two copies of the frame: the loss is a mean, the gradient of shared parameters too
    p4 = _params().cuda().requires_grad_()
    l4 = ops.pipeline_mse(p4, raw.repeat(2, 1, 1, 1), gt.repeat(2, 1, 1, 1), 'bilinear', chain)
    g4, = torch.autograd.grad(l4, p4)
    assert abs(float(l4.detach()) - float(l1.detach())) <= 1e-6 * float(l1.detach())
    assert float((g4 - g1).abs().max()) <= 1e-5 * float(g1.abs().max())


def test_full_frame_linear_and_index_properties(ops, frame):
    raw, gt = frame
    raw, gt = raw.cuda(), gt.cuda()
    g = torch.Generator(device='cuda').manual_seed(5)
    # nearest demosaic and its adjoint: <D r, y> == <r, D^T y>
    r = raw.clone().requires_grad_()
    yv = torch.randn(1, 3, H, W, device='cuda', generator=g)
    d = ops.demosaic(r, 'nearest')
    rt, = torch.autograd.grad(d, r, yv)
    lhs, rhs = float((d.detach().double() * yv.double()).sum()), float((raw.double() * rt.double()).sum())
    assert abs(lhs - rhs) <= 1e-6 * max(1.0, abs(lhs))          # fp32 sums inside the adjoint, fp64 dot products outside
    # bilinear demosaic is linear: D(a r1 + b r2) == a D r1 + b D r2 (fp32 rounding only)
    r2 = torch.rand(1, 1, H, W, device='cuda', generator=g)
    lin = ops.demosaic(0.25 * raw + 0.5 * r2, 'bilinear') - (0.25 * ops.demosaic(raw, 'bilinear') + 0.5 * ops.demosaic(r2, 'bilinear'))
    assert float(lin.abs().max()) <= 2e-6, float(lin.abs().max())
    # pack / unpack and pixel shuffles are permutations: exact round trips
    assert torch.equal(ops.unpack_rggb(ops.pack_rggb(raw)), raw), 'pack/unpack'
    assert torch.equal(ops.pixel_shuffle2(ops.pixel_unshuffle2(gt)), gt), 'pixel (un)shuffle'
    # code normalisation is bit-exact with the loader's division at full size
    codes = torch.round(raw * 1023).to(torch.int16)
    # (against the CPU division the reference's loader performs; torch's CUDA `/ scalar` multiplies by the reciprocal)
    assert torch.equal(ops.decode_codes(codes, 1023.).cpu(), codes.cpu().float() / 1023.), 'decode_codes'
    # identity-parameter chain is the identity on [0,1] data up to rounding; clip stages are idempotent
    ident = [0.0] * 30
    ident[6] = ident[17] = ident[28] = 1.0
    pid = torch.tensor([[1., 1., 1.] + ident + [1.0] + [0.25, 0.5, 0.75]], device='cuda')
    chain = ops.Chain(['gain', 'poly10', 'gamma', ('gtm', 4)])
    y1 = ops.chain_apply(gt, chain, pid)
    # (gamma = 1 still goes through lg2 / ex2: a few 1e-7 relative)
    assert float((y1 - gt).abs().max()) <= 2e-5, float((y1 - gt).abs().max())
    # split into the 63 shipped tiles and blend back: identity up to the blend's fp32 rounding
    tiles, pos = ops.whole2patch(gt[0], (512, 512), (480, 480))
    assert tiles.shape[0] == 63
    back = ops.patch2whole(tiles, (H, W), (480, 480))
    assert float((back - gt[0]).abs().max()) <= 1e-6, float((back - gt[0]).abs().max())


def test_full_frame_stencils_match_oracle_on_crops(ops, frame):
    """Register-marching stencils at 12 MP (many strips / row chunks / running sums over long chunks): median 3x3 bit-exact and
    the guided filter within tolerance against the oracle on interior and border crops (the oracle sees the crop + halo; on
    the frame border the crop includes the border itself, so the border rules are compared too)."""
    raw, gt = frame
    x = gt.cuda()                                           # (1,3,H,W) in [0,1]
    x255 = x * 255
    med = ops.median(x255, 3)
    gf = ops.guided_filter(x, 4, 1e-3)
    halo = 8
    for (y0, x0) in ((0, 0), (0, W - 160), (H - 96, 0), (H - 96, W - 160), (1480, 2000), (777, 3880)):
        h, w = 96, 160
        ya, yb, xa, xb = max(0, y0 - halo), min(H, y0 + h + halo), max(0, x0 - halo), min(W, x0 + w + halo)
        # crops that touch the frame border keep it (same border rule); inner edges carry a halo that is cut off again
        cy, cx = y0 - ya, x0 - xa
        crop = x[:, :, ya:yb, xa:xb].cpu()
        m_ref = O.denoise_median(crop * 255, 3)[:, :, cy:cy + h, cx:cx + w]
        assert torch.equal(med[:, :, y0:y0 + h, x0:x0 + w].cpu(), m_ref), (y0, x0)
        g_ref = O.guided_filter(crop, 4, 1e-3)[:, :, cy:cy + h, cx:cx + w]
        assert float((gf[:, :, y0:y0 + h, x0:x0 + w].cpu() - g_ref).abs().max()) <= 1e-4, (y0, x0)


def test_mixed_op_specialised_kernel_matches_sum_of_candidates(ops):
    """The supernet's sRGB mixed-op at 2 MP (compile-time candidate signature, batched candidate loads, many blocks) against
    the same weighted sum built from the single-candidate chain kernels + torch, forward and every gradient."""
    from reconfigisp_b200.modules.super_prune_fifteen_demos_four_bayer_two import SRGB_CLASSICAL
    g = torch.Generator().manual_seed(21)
    N, Hh, Ww, K_ext = 2, 1000, 1000, 9
    dev = 'cuda'
    x = torch.rand(N, 3, Hh, Ww, generator=g).to(dev).requires_grad_()
    ext = [torch.rand(N, 3, Hh, Ww, generator=g).to(dev).requires_grad_() for _ in range(K_ext)]
    ident = [0.0] * 30
    ident[6] = ident[17] = ident[28] = 1.0
    row = torch.tensor([0.6, 1.1, 0.9, 1.05, 1.05, 1.0, 0.95] + ident + [0.3, 0.5, 0.8]) + torch.randn(40, generator=g) * 0.01
    table = row.view(1, -1).repeat(N, 1).to(dev).requires_grad_()
    w = torch.softmax(torch.randn(len(SRGB_CLASSICAL) + K_ext, generator=g), 0).to(dev).requires_grad_()
    dy = torch.randn(N, 3, Hh, Ww, generator=g).to(dev)
    chain = ops.Chain([op for _, op in SRGB_CLASSICAL])
    y = ops.mixed_op(x, chain, table, w, ext)
    got = torch.autograd.grad(y, [x, table, w] + ext, dy)
    # reference: one chain kernel per classical candidate on its slice of the table
    offs, o = [], 0
    for _, op in SRGB_CLASSICAL:
        c1 = ops.Chain([op])
        offs.append((o, c1.P, c1))
        o += c1.P
    yr = 0
    for j, (o0, P, c1) in enumerate(offs):
        t = x if P == 0 and c1.names == ['skip'] else ops.chain_apply(x, c1, table[:, o0:o0 + P] if P else None)
        yr = yr + w[j] * t
    for e in range(K_ext):
        yr = yr + w[len(offs) + e] * ext[e]
    ref = torch.autograd.grad(yr, [x, table, w] + ext, dy)
    assert float((y - yr).detach().abs().max()) <= 2e-6
    for a, b in zip(got, ref):
        scale = max(1e-6, float(b.abs().max()))
        assert float((a - b).abs().max()) <= 2e-4 * scale, (a.shape, float((a - b).abs().max()), scale)
