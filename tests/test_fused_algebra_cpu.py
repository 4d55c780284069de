"""The algebra the fused step kernel (csrc/risp_fused.cu / risp_fused.cuh) relies on, restated in fp64 torch and held against
autograd through the ORACLE's stage functions (oracle/isp_oracle.py, the checker) -- CPU only, no kernel involved.  What is
checked is the derivation, so that a parity failure on the GPU can be pinned on code rather than on mathematics:

  1. wbmanual -> wbquadratic folded into one polynomial q' = q * gmon(g): both parameter gradients follow from the single
     accumulator S[c][i] = sum e_c phi_i(x0) (Euler's identity for the monomials), including g = 0 exactly;
  2. the tone curve as a table (a_k, s_k): out = s_k x + a_k equals the reference's segment formula (tools_origin.py:429-435),
     and the knot gradients are the second differences of the hinge sums A_0 = sum d x, A_k = sum d max(x - x_k, 0);
  3. gamma's mask [y >= eps] and the polynomial's clamp mask [0 <= u <= 1] are the single compare max(sat(u), eps) == u.
"""
import torch

from oracle import isp_oracle as O

torch.manual_seed(0)
EPS = O.GAMMA_EPS if hasattr(O, 'GAMMA_EPS') else 1e-8


def monomials(x):                      # x: (N,3,H,W) BGR -> (N,10,H,W): b2 g2 r2 bg br gr b g r 1  (WbQuadratic's order)
    b, g, r = x[:, 0], x[:, 1], x[:, 2]
    return torch.stack([b * b, g * g, r * r, b * g, b * r, g * r, b, g, r, torch.ones_like(b)], dim=1)


def test_folded_gain_polynomial_gradients_from_one_accumulator():
    N, H, W = 2, 9, 11
    x0 = torch.rand(N, 3, H, W, dtype=torch.float64)
    gt = torch.rand(N, 3, H, W, dtype=torch.float64)
    ident = torch.zeros(30, dtype=torch.float64)
    ident[6] = ident[17] = ident[28] = 1.0
    for gains in ([1.1, 0.9, 1.2], [0.0, 0.9, 1.2], [1.0, 0.0, 0.0]):
        g = torch.tensor([gains], dtype=torch.float64, requires_grad=True)
        q = (ident + 0.2 * torch.randn(30, dtype=torch.float64)).view(1, 30).requires_grad_()
        # oracle: the two stages one after the other (kernel-level q -> the reference's [0,1] parameter is (q + 5) / 10)
        y = O.wb_quadratic(O.wb_manual(x0, g.expand(N, -1)), ((q + 5) / 10).expand(N, -1))
        loss = ((y - gt) ** 2).mean()
        dg_ref, dq_ref = torch.autograd.grad(loss, [g, q])
        # kernel algebra: one polynomial in x0 with q' = q * gmon(g); e = dL/du masked by the inclusive clamp
        gb, gg, gr = g.detach()[0]
        gmon = torch.stack([gb * gb, gg * gg, gr * gr, gb * gg, gb * gr, gg * gr, gb, gg, gr, torch.tensor(1.0, dtype=torch.float64)])
        qd = q.detach().view(3, 10)
        phi = monomials(x0)                                                     # (N,10,H,W)
        u = torch.einsum('ci,nihw->nchw', qd * gmon, phi)
        assert torch.allclose(torch.clamp(u, 0, 1), y.detach(), atol=1e-12)
        e = 2 * (torch.clamp(u, 0, 1) - gt) / gt.numel() * ((u >= 0) & (u <= 1))
        S = torch.einsum('nchw,nihw->ci', e, phi)                               # the ONLY per-pixel accumulation
        dq = gmon * S
        one, zero = torch.tensor(1.0, dtype=torch.float64), torch.tensor(0.0, dtype=torch.float64)
        dgb = torch.stack([2 * gb, zero, zero, gg, gr, zero, one, zero, zero, zero])
        dgg = torch.stack([zero, 2 * gg, zero, gb, zero, gr, zero, one, zero, zero])
        dgr = torch.stack([zero, zero, 2 * gr, zero, gb, gg, zero, zero, one, zero])
        dg = torch.stack([(qd * dgb * S).sum(), (qd * dgg * S).sum(), (qd * dgr * S).sum()])
        assert torch.allclose(dq.view(1, 30), dq_ref, atol=1e-12), gains
        assert torch.allclose(dg.view(1, 3), dg_ref, atol=1e-12), gains


def tone_table(knots):
    y = [0.0] + list(knots) + [1.0]
    s = [(y[k + 1] - y[k]) * 4.0 for k in range(4)]
    a = [y[k] - 0.25 * k * s[k] for k in range(4)]
    return torch.tensor(a + [0.0], dtype=torch.float64), torch.tensor(s + [1.0], dtype=torch.float64)     # entry 4: x == 1


def test_tone_curve_table_and_hinge_sum_gradients():
    x = torch.cat([torch.rand(1, 3, 16, 16, dtype=torch.float64).view(-1),
                   torch.tensor([0.0, 0.25, 0.5, 0.75, 1.0, 0.25 - 1e-12, 0.75 + 1e-12], dtype=torch.float64)]).view(1, 1, 1, -1)
    d = torch.randn_like(x)
    for knots in ([0.1, 0.7, 0.8], [0.25, 0.5, 0.75], [0.6, 0.3, 0.9]):            # monotone or not: all in [0,1] (fast path)
        p = torch.tensor([knots], dtype=torch.float64, requires_grad=True)
        xr = x.clone().requires_grad_()
        out = O.gtm_manual(xr, p, 4)
        dx_ref, dp_ref = torch.autograd.grad(out, [xr, p], d)
        a, s = tone_table(knots)
        k = torch.clamp(torch.floor(4 * x), max=4).long()                           # the table entry (test_tone_curve_index_cpu.py)
        assert torch.allclose(s[k] * x + a[k], out.detach(), atol=1e-12)
        assert torch.allclose(d * s[k], dx_ref, atol=1e-12)
        A = [(d * x).sum()] + [(d * torch.clamp(x - 0.25 * j, min=0)).sum() for j in (1, 2, 3)] + [torch.tensor(0.0, dtype=torch.float64)]
        dp = torch.stack([4 * (A[j - 1] - 2 * A[j] + A[j + 1]) for j in (1, 2, 3)])
        assert torch.allclose(dp.view(1, 3), dp_ref, atol=1e-10), knots
        # the hinge as one saturating add: x - x_k < 1 on [0,1], so clamp(., 0, 1) == max(., 0)
        for j in (1, 2, 3):
            assert torch.equal(torch.clamp(x - 0.25 * j, 0, 1), torch.clamp(x - 0.25 * j, min=0))


def test_merged_gamma_and_clamp_mask():
    u = torch.cat([torch.randn(4096, dtype=torch.float32) * 0.7 + 0.5,
                   torch.tensor([0.0, 1.0, EPS, EPS / 2, 1.0 + 1e-7, -0.0, 1e-8, 9.9e-9, float('nan')], dtype=torch.float32)])
    y = torch.clamp(u, 0, 1)
    y = torch.where(torch.isnan(u), torch.zeros_like(u), y)                          # __saturatef(NaN) = 0
    xc = torch.clamp(y, min=float(EPS))                                              # what gamma computes anyway
    separate = (u == y) & (y >= float(EPS))
    merged = xc == u
    assert torch.equal(separate, merged)
