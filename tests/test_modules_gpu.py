"""GPU parity of the module / container layer against the golden vectors recorded from the
reference's own containers (oracle/gen_golden.py) and against the CPU oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import isp_oracle as O            # noqa: E402
from oracle import pipeline_oracle as PO      # noqa: E402

T = torch.from_numpy


def maxabs(a, b):
    b = T(b) if isinstance(b, np.ndarray) else b
    return float((a.detach().cpu().double() - b.detach().cpu().double()).abs().max())


def relclose(a, b, rtol=2e-3, atol=1e-6):
    b = T(b) if isinstance(b, np.ndarray) else b
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    assert a.shape == b.shape, (a.shape, b.shape)
    assert float((a - b).abs().max()) <= atol + rtol * float(b.abs().max()), (float((a - b).abs().max()), float(b.abs().max()))


def relclose_relu_net(a, b, rtol=1e-3, flip_frac=5e-2):
    """Gradients through ReLU networks: a pre-activation within rounding distance of 0 can land on the other
    side of the ReLU than in the reference, which changes the gradient inside that unit's receptive field.
    So: all but a small fraction of the elements must agree to rtol (one flipped unit of SRCNNRes touches a 13x13
    patch, i.e. ~6% of the 20x24 golden image), and the outliers must stay small."""
    b = T(b) if isinstance(b, np.ndarray) else b
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    err, scale = (a - b).abs(), float(b.abs().max())
    assert float((err > rtol * scale).double().mean()) <= flip_frac, float((err > rtol * scale).double().mean())
    assert float(err.max()) <= 2e-2 * scale, (float(err.max()), scale)


def test_cnn_candidates_match_reference(golden):
    from reconfigisp_b200.modules import tools_proxy as P
    g = golden('cnn_candidates')
    x3, raw = T(g['x3']).cuda().requires_grad_(), T(g['raw']).cuda().requires_grad_()
    specs = [('srcnn_res3', P.ProxyNet(3, None), 10), ('srcnn_res1', P.ProxyNet(1, None), 11),
             ('srcnn_demosaic', P.ProxyDemosaicNet(0, None), 12), ('path14l_bayer', P.PathRestore14lBayer(0, None), 13),
             ('path14l_bgr', P.PathRestore14lBgr(0, None), 14)]
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    for name, net, seed in specs:
        net.load_state_dict(P.seeded_state_dict(net, seed))
        net = net.cuda().requires_grad_(False)        # candidate nets are frozen constants on this path
        if name.startswith('srcnn_res'):
            y = net(x3, T(g[name + '_par']).cuda()); inp = x3
        elif name == 'path14l_bgr':
            y = net(x3, None); inp = x3
        else:
            y = net(raw, None); inp = raw
        dx, = torch.autograd.grad(y.square().sum(), inp)
        assert maxabs(y, g[name + '_y']) <= 1e-4, name
        relclose_relu_net(dx, g[name + '_dx'])
        # state-dict keys are part of the API (tools_proxy.py load())
        assert list(net.state_dict().keys()) == list(PO.seeded_weights(PO.ARCH_SHAPES[
            {'srcnn_res3': 'srcnn_res', 'srcnn_res1': 'srcnn_res'}.get(name, name)](int(name[-1]) if name[-1].isdigit() else 0), 0).keys())


ARCHS = (('classical', 'Bayer_02_Demosaic_02_sRGB_11_13_01'), ('sid', 'Bayer_01_Demosaic_03_sRGB_01_13_11'),
         ('s7isp', 'Bayer_01_Demosaic_01_sRGB_04_01_13'), ('all_origin', 'Bayer_02_Demosaic_01_sRGB_05_02_03_04_06_07_08_10_12_15'),
         ('nlm', 'Bayer_02_Demosaic_01_sRGB_09_01'))


@pytest.mark.parametrize('fuse', [False, True])
def test_origin_universal_matches_reference(golden, fuse):
    from reconfigisp_b200.modules.origin_universal import OriginUniversal
    g = golden('fixed_pipelines')
    raw = T(g['raw']).cuda()
    for tag, arch in ARCHS:
        net = OriginUniversal('/x/', arch, weight_seed=10, fuse=fuse).cuda()
        y = net(raw)
        assert list(net.state_dict().keys()) == list(g[tag + '_keys']), tag
        assert maxabs(y, g[tag + '_y']) <= 1e-4, (tag, maxabs(y, g[tag + '_y']))
        inter = net.intermediate_results
        assert len(inter) == len(net.all_modules)
        for i in range(len(inter)):
            assert maxabs(inter[i], g['%s_inter%d' % (tag, i)]) <= 1e-4, (tag, i)


@pytest.mark.parametrize('fuse', [False, True])
def test_isp_universal_grads_match_reference(golden, fuse):
    from reconfigisp_b200.modules.isp_universal import IspUniversal
    from reconfigisp_b200 import ops
    g = golden('fixed_pipelines')
    raw, gt = T(g['raw']).cuda(), T(g['isp_gt']).cuda()
    net = IspUniversal('/x/', (None,) * 7, 'Bayer_02_Demosaic_01_sRGB_11_13_01_14_05', weight_seed=10, fuse=fuse).cuda()
    assert list(net.state_dict().keys()) == list(g['isp_keys'])
    y = net(raw)
    assert maxabs(y, g['isp_y']) <= 1e-4
    loss = ops.mse_loss(y, gt)
    assert abs(float(loss.detach()) - float(g['isp_loss'])) <= 1e-6
    nz = [(k, q) for k, q in net.named_parameters()]
    grads = torch.autograd.grad(loss, [q for _, q in nz])
    for (k, _), gr in zip(nz, grads):
        relclose(gr, g['isp_dlogit_' + k])


def test_isp_universal_state_dict_loads_into_origin(golden):
    """S7ISP_test.yml:29-30: a checkpoint tuned with IspUniversal loads strict into OriginUniversal."""
    from reconfigisp_b200.modules.isp_universal import IspUniversal
    from reconfigisp_b200.modules.origin_universal import OriginUniversal
    arch = 'Bayer_01_Demosaic_03_sRGB_01_13_11'
    a = IspUniversal('/x/', (None,) * 5, arch, weight_seed=10)
    b = OriginUniversal('/x/', arch, weight_seed=10)
    b.load_state_dict(a.state_dict(), strict=True)
    with pytest.raises(ValueError):
        OriginUniversal('/x/', '01_02', weight_seed=1)
    with pytest.raises(AssertionError):
        OriginUniversal('/x/', 'sRGB_16', weight_seed=1)


def test_supernet_matches_reference(golden):
    from reconfigisp_b200.modules.super_prune_fifteen_demos_four_bayer_two import SuperPruneFifteenDemosFourBayerTwo
    from reconfigisp_b200 import ops
    g = golden('supernet')
    net = SuperPruneFifteenDemosFourBayerTwo(int(g['n_step']), float(g['threshold']), '/x/', weight_seed=int(g['weight_seed'])).cuda()
    names = [n for n, _ in net.named_parameters()]
    assert names[:2] == ['alpha_bayer', 'alpha_demosaic'] and len(names) == 2 + 13 * int(g['n_step'])
    with torch.no_grad():
        for i, a in enumerate(net.alphas):
            a.copy_(T(g['alpha%d' % i]))
        for n, q in net.named_parameters():
            if n.startswith('param_'):
                q.copy_(T(g['logit_' + n]))
    y = net(T(g['raw']).cuda())
    assert net.pruned_paths == list(g['pruned'])
    assert maxabs(y, g['y']) <= 1e-4
    for i, m in enumerate(net.intermediate_results):
        assert maxabs(m, g['inter%d' % i]) <= 1e-4, i
    loss = ops.mse_loss(y, T(g['gt']).cuda())
    assert abs(float(loss.detach()) - float(g['loss'])) <= 2e-5 * float(g['loss'])
    nz = [q for q in net.trainable_parameters if q.nelement()]
    grads = torch.autograd.grad(loss, list(net.alphas) + nz, allow_unused=True)
    for i in range(len(net.alphas)):
        relclose(grads[i], g['dalpha%d' % i], rtol=2e-3, atol=1e-7)
    pnames = [n for n in names if n.startswith('param_')]
    for n, gr in zip(pnames, grads[len(net.alphas):]):
        assert gr is not None, n                               # dummy-gradient quirk: zeros, not None
        relclose(gr, g['dlogit_' + n], rtol=2e-3, atol=1e-7)


def test_conditional_module_matches_reference(golden):
    from reconfigisp_b200.modules import tools_origin as TO
    g = golden('cond_fc')
    cm = TO.ConditionalWbManual(in_channels=(12, 5))
    flat = T(g['flat']).cuda().requires_grad_()
    out = cm._fc_forward(T(g['img']).cuda(), flat)
    dflat, = torch.autograd.grad(out, flat, torch.ones_like(out))
    assert maxabs(out, g['out']) <= 1e-6
    relclose(dflat, g['dflat'], rtol=1e-4)
    # and the full module against the oracle
    img = T(g['img']).cuda().clamp(0, 1)
    y = cm(img, flat)
    yo = O.wb_manual(img.cpu(), T(g['out']) * 5) if False else None
    p = O.fc_params(O.histc_planes(img.cpu(), 4), T(g['flat']), [12, 5, 3])
    assert maxabs(y, O.wb_manual(img.cpu(), p * 5)) <= 1e-4


def test_plugin_boundary_run_signatures():
    """The five kernel modules resolve by bare name and honour the layouts of tools_origin.py."""
    from reconfigisp_b200 import isp_kernels
    isp_kernels.install()
    import whitebalance, gamma, demosaic, globaltonemapping, spatialnoisereduction   # noqa: E401
    g = torch.Generator().manual_seed(3)
    x = torch.rand(2, 3, 16, 24, generator=g)
    xd = x.cuda()
    nhwc = xd.permute(0, 2, 3, 1)                                    # what the reference wrappers pass
    gm = torch.tensor([[0.4], [0.7]])
    out = gamma.Gamma().run(nhwc, 'manual', {'gamma': gm.cuda()})
    assert out.shape == (2, 16, 24, 3) and maxabs(out.permute(0, 3, 1, 2), O.gamma_manual(x, gm)) <= 1e-4
    gain = torch.tensor([[1., 2., 0.5], [0.3, 1.2, 5.0]])
    out = whitebalance.WhiteBalance().run(nhwc, 'manual', {'gain': gain.cuda()})
    assert maxabs(out.permute(0, 3, 1, 2), O.wb_manual(x, gain)) <= 1e-5
    out = whitebalance.WhiteBalance().run(nhwc, 'grayworld', {'input': {}, 'output': {}})
    assert maxabs(out.permute(0, 3, 1, 2), O.wb_grayworld(x)) <= 1e-4
    ratio = np.array([0.1, 0.4], dtype=np.float32)
    out = whitebalance.WhiteBalance().run(nhwc * 255., 'whiteworld', {'white_point_ratio': ratio})
    assert maxabs(out.permute(0, 3, 1, 2) / 255, O.wb_whiteworld(x * 255, ratio) / 255) <= 1e-4
    wp, mg = np.array([0.6, 0.9], np.float32), np.array([0.5, 0.2], np.float32)
    out = globaltonemapping.GlobalToneMapping().run(nhwc * 255., 'reinhard', {'white_point': wp, 'middle_grey': mg})
    assert maxabs(out.permute(0, 3, 1, 2) / 255, O.tone_reinhard(x * 255, T(wp), T(mg)) / 255) <= 1e-4
    out = globaltonemapping.GlobalToneMapping().run(nhwc * 255., 'filmic', {'white_point': wp, 'exposure_bias': mg * 9 + 1})
    assert maxabs(out.permute(0, 3, 1, 2) / 255, O.tone_filmic(x * 255, T(wp), T(mg * 9 + 1)) / 255) <= 1e-4
    out = globaltonemapping.GlobalToneMapping().run(nhwc * 255., 'crysisengine', {'lum_adapted': mg})
    assert maxabs(out.permute(0, 3, 1, 2) / 255, O.tone_crysis(x * 255, T(mg)) / 255) <= 1e-4
    raw = torch.rand(2, 1, 16, 24, generator=g)
    desc = {'input': {'format': 'RGGB', 'bitdepth': 10}, 'output': {'format': 'BGR', 'bitdepth': 8}}
    out = demosaic.Demosaic().run(raw.cuda(), 'nearestneighbor', desc)
    assert out.shape == (2, 3, 16, 24) and torch.equal(out.cpu(), O.demosaic_nearest(raw))
    out = demosaic.Demosaic().run((raw.cuda() * 255).permute(0, 2, 3, 1), 'laplacian', desc)
    assert out.shape == (2, 16, 24, 3) and maxabs(out.permute(0, 3, 1, 2) / 255, O.demosaic_laplacian(raw * 255, 255.) / 255) <= 1e-5
    out = spatialnoisereduction.SpatialNoiseReduction().run(nhwc * 255., 'median', {'size': 5})
    assert torch.equal(out.permute(0, 3, 1, 2).cpu(), O.denoise_median(x * 255, 5))
    win = torch.tensor([3, 3], dtype=torch.int32)
    out = spatialnoisereduction.SpatialNoiseReduction().run(nhwc * 255., 'bilateral', {
        'window_length': win.cuda(), 'sigma_color': torch.tensor([20., 50.]).cuda(), 'sigma_space': torch.tensor([5., 2.]).cuda()})
    assert maxabs(out.permute(0, 3, 1, 2) / 255, O.denoise_bilateral(x * 255, win, [20., 50.], [5., 2.]) / 255) <= 1e-4
    out = spatialnoisereduction.SpatialNoiseReduction().run(nhwc * 255., 'fastnlm', {
        'block_size': win.cuda(), 'search_block': torch.tensor([5, 3], dtype=torch.int32).cuda(), 'decay_factor': torch.tensor([20., 50.]).cuda()})
    assert maxabs(out.permute(0, 3, 1, 2) / 255, O.denoise_fastnlm(x * 255, win, [5, 3], [20., 50.]) / 255) <= 1e-4
    with pytest.raises(ValueError):
        gamma.Gamma().run(nhwc, 'auto', {})


def test_tuning_step_cuda_graph_matches_eager():
    """IspModel replays the fused tuning step from a CUDA graph; parameters after several steps must match the
    eagerly launched step (same kernels, same order), also when new data is fed into the persistent buffers."""
    from reconfigisp_b200.tuning import IspModel
    from reconfigisp_b200.synthetic import synthetic_frames

    def opt(graph):
        return {'model': 'isp', 'is_train': True, 'cuda_graph': graph,
                'network_G': {'which_model_G': 'OriginUniversal', 'architecture': 'Bayer_02_Demosaic_02_sRGB_11_13_01_14', 'weight_seed': 10},
                'train': {'lr_G': 1e-2, 'beta1': 0.9, 'beta2': 0.99, 'pixel_criterion': 'l2', 'lr_scheme': 'MultiStepLR',
                          'lr_steps': [1000], 'lr_gamma': 0.5}, 'path': {'pretrain_model_G': None}}
    raw, gt = synthetic_frames(4, 64, 96, seed=3)
    raw_c, gt_c = torch.round(raw * 1023).to(torch.int16), torch.round(gt * 255).to(torch.uint8)
    models = [IspModel(opt(True)), IspModel(opt(True))]
    models[0].use_graph = False          # same (capturable) Adam arithmetic, launched eagerly
    losses = [[], []]
    for k, m in enumerate(models):
        for step in range(8):
            sl = slice(0, 2) if step % 2 == 0 else slice(2, 4)
            m.feed_data((raw_c[sl], gt_c[sl]) if step >= 6 else (raw[sl], gt[sl]))
            m.optimize_parameters()
            losses[k].append(float(m.log_dict['loss']))
    assert models[1]._graph is not None, 'the graph path was not taken'
    assert max(abs(a - b) for a, b in zip(*losses)) <= 1e-6 * max(losses[0]), losses
    for pa, pb in zip(models[0].netG.trainable_parameters, models[1].netG.trainable_parameters):
        if pa.numel():
            assert maxabs(pa, pb.detach()) <= 1e-5


def test_relu_gradient_outliers_are_mask_flips(golden):
    """Why `relclose_relu_net` tolerates a few per cent of outliers: they must be ReLU mask flips and nothing else.  For
    SRCNNRes (conv9-ReLU-conv5-ReLU-conv5 + x) every input-gradient element that disagrees with the fp64 oracle by more
    than rtol has to lie inside the receptive field (radius 4 through conv9, 2+4 through conv5 o conv9) of a unit whose
    fp64 pre-activation is within rounding distance of zero -- and away from such units the tolerance is the tight one."""
    import torch.nn.functional as F
    from reconfigisp_b200.modules import tools_proxy as P
    g = golden('cnn_candidates')
    net = P.ProxyNet(3, None)
    sd = P.seeded_state_dict(net, 10)
    net.load_state_dict(sd)
    net = net.cuda().requires_grad_(False)
    x = T(g['x3'])
    par = T(g['srcnn_res3_par'])
    xg = x.cuda().requires_grad_()
    y = net(xg, par.cuda())
    dx, = torch.autograd.grad(y.square().sum(), xg)
    # fp64 oracle with its pre-activations
    xd = x.double().requires_grad_()
    sdd = {k: v.double() for k, v in sd.items()}
    N, _, H, W = x.shape
    feat = torch.cat([xd.amin(dim=(2, 3)), xd.mean(dim=3).mean(dim=2), xd.amax(dim=(2, 3)), par.double()], dim=1).view(N, -1, 1, 1).expand(-1, -1, H, W)
    a1 = F.conv2d(torch.cat([xd, feat], dim=1), sdd['srcnn.0.weight'], sdd['srcnn.0.bias'], padding=4)
    a2 = F.conv2d(torch.relu(a1), sdd['srcnn.2.weight'], sdd['srcnn.2.bias'], padding=2)
    yd = xd + F.conv2d(torch.relu(a2), sdd['srcnn.4.weight'], sdd['srcnn.4.bias'], padding=2)
    dref, = torch.autograd.grad(yd.square().sum(), xd)
    err = (dx.cpu().double() - dref).abs().amax(dim=1, keepdim=True)                  # (N,1,H,W)
    scale = float(dref.abs().max())
    eps1, eps2 = 3e-5 * float(a1.detach().abs().max()), 3e-5 * float(a2.detach().abs().max())
    near1 = F.max_pool2d((a1.abs() < eps1).any(dim=1, keepdim=True).double(), 9, 1, 4)    # receptive field of a layer-1 unit
    near2 = F.max_pool2d((a2.abs() < eps2).any(dim=1, keepdim=True).double(), 13, 1, 6)   # ... of a layer-2 unit
    allowed = (near1 + near2) > 0
    outlier = err > 1e-3 * scale
    assert not bool((outlier & ~allowed).any()), 'gradient outliers away from any near-zero pre-activation'
    assert float(err[~allowed].max() if bool((~allowed).any()) else 0.0) <= 2e-4 * scale
    assert float(err.max()) <= 2e-2 * scale


@pytest.mark.parametrize('crit', ['l2', 'l1'])
def test_tuning_fast_step_matches_autograd_step(crit):
    """The graph-free tuning step ([logits -> table] kernel, single-pass fused step, [d table -> d logits] kernel, flat
    gradient buffer) against the autograd formulation of the same step (`fast_step: False`): losses and parameters after
    several Adam updates, MSE and L1 (isp_model.py:44-49, :128-142)."""
    from reconfigisp_b200.tuning import IspModel
    from reconfigisp_b200.synthetic import synthetic_frames

    def opt(fast):
        return {'model': 'isp', 'is_train': True, 'cuda_graph': False, 'fast_step': fast,
                'network_G': {'which_model_G': 'OriginUniversal', 'architecture': 'Bayer_02_Demosaic_02_sRGB_11_13_01_14', 'weight_seed': 10},
                'train': {'lr_G': 1e-2, 'beta1': 0.9, 'beta2': 0.99, 'pixel_criterion': crit, 'lr_scheme': 'MultiStepLR',
                          'lr_steps': [1000], 'lr_gamma': 0.5}, 'path': {'pretrain_model_G': None}}
    raw, gt = synthetic_frames(2, 64, 96, seed=5)
    models = [IspModel(opt(True)), IspModel(opt(False))]
    losses, grads0 = [[], []], [None, None]
    for k, m in enumerate(models):
        m.feed_data((raw.cuda(), gt.cuda()))
        for step in range(5):
            m.optimize_parameters()
            losses[k].append(float(m.log_dict['loss'].detach()))
            if step == 0:
                grads0[k] = [p.grad.detach().clone() for p in m.netG.trainable_parameters if p.numel()]
    assert models[0]._fast is not None and getattr(models[1], '_fast', None) is None
    # the first step is the same function evaluated two ways: loss identical, gradients to rounding
    assert abs(losses[0][0] - losses[1][0]) <= 1e-7
    for ga, gb in zip(*grads0):
        relclose(ga, gb, rtol=2e-5, atol=1e-9)
    # afterwards Adam (lr 1e-2, update ~ lr * g / |g| in the first steps) amplifies those roundings: trajectories stay close
    assert max(abs(a - b) for a, b in zip(*losses)) <= 2e-3 * max(losses[1]), losses
    for pa, pb in zip(models[0].netG.trainable_parameters, models[1].netG.trainable_parameters):
        if pa.numel():
            assert maxabs(pa.detach(), pb.detach()) <= 2e-3
    # state-dict keys and shapes are untouched by the flat re-layout
    sa, sb = models[0].netG.state_dict(), models[1].netG.state_dict()
    assert list(sa.keys()) == list(sb.keys()) and all(sa[k].shape == sb[k].shape for k in sa)


# ---- SRCNNRes with its constant channels folded into a border-class bias table (search path) ---------------------------
def _ref_stats3(x):
    """srcnn_res_arch.py:36-40, verbatim reductions (CPU)."""
    mn, _ = torch.min(x, dim=3)
    mn, _ = torch.min(mn, dim=2)
    me = torch.mean(torch.mean(x, dim=3), dim=2)
    mx, _ = torch.max(x, dim=3)
    mx, _ = torch.max(mx, dim=2)
    return torch.cat([mn, me, mx], dim=1)


def test_plane_stats3_routes_min_max_gradient_like_the_reference():
    from reconfigisp_b200 import ops
    g = torch.Generator().manual_seed(3)
    x = torch.rand(3, 3, 18, 22, generator=g).clamp(0.25, 0.75)        # many ties at both extremes
    x[1, 2] = 0.5                                                       # a constant plane: min == max everywhere
    w = torch.randn(3, 9, generator=g)
    xc = x.clone().requires_grad_()
    sc = _ref_stats3(xc)
    gc, = torch.autograd.grad((sc * w).sum(), xc)
    xg = x.cuda().requires_grad_()
    sg = ops.plane_stats3(xg)
    gg, = torch.autograd.grad((sg * w.cuda()).sum(), xg)
    assert torch.equal(sg[:, :3].cpu(), sc[:, :3]) and torch.equal(sg[:, 6:].cpu(), sc[:, 6:])     # min / max bit-exact
    assert maxabs(sg[:, 3:6], sc[:, 3:6].detach()) <= 1e-6
    assert maxabs(gg, gc) <= 1e-6          # same single pixel receives the min / max gradient


@pytest.mark.parametrize('K,C,shape', [(9, 64, (2, 20, 24)), (9, 64, (1, 8, 8)), (5, 32, (2, 11, 150)), (3, 3, (1, 6, 5))])
def test_blocked_class_sums_match_torch(K, C, shape):
    from reconfigisp_b200 import ops
    N, H, W = shape
    g = torch.Generator().manual_seed(5)
    d = torch.randn(N, C, H, W, generator=g).cuda()
    m = torch.randn(N, C, H, W, generator=g).cuda()
    PAD = K // 2
    cy = torch.tensor([y if y < PAD else (K - (H - y) if y >= H - PAD else PAD) for y in range(H)])
    cx = torch.tensor([x if x < PAD else (K - (W - x) if x >= W - PAD else PAD) for x in range(W)])
    oh = torch.zeros(H, W, K * K, dtype=torch.float64)
    oh[torch.arange(H).view(H, 1), torch.arange(W).view(1, W), cy.view(H, 1) * K + cx.view(1, W)] = 1
    for mask in (None, m):
        dm = d if mask is None else d * (mask > 0)
        ref = torch.einsum('nchw,hwk->nkc', dm.double().cpu(), oh)
        out = ops.blocked_class_sums(ops.to_blocked(d), None if mask is None else ops.to_blocked(mask), C, K)
        assert out.shape == (N, K * K, (C + 15) // 16 * 16)
        assert maxabs(out[..., :C], ref) <= 1e-4 * max(1.0, float(ref.abs().max()))
        assert float(out[..., C:].abs().max()) == 0.0 if out.shape[-1] > C else True


@pytest.mark.parametrize('P_,shape', [(2, (2, 20, 24)), (1, (1, 8, 8)), (5, (2, 37, 150)), (3, (1, 64, 260))])
def test_srcnn_res_folded_constant_channels_match_materialised(P_, shape):
    from reconfigisp_b200.modules import tools_proxy as P
    N, H, W = shape
    g = torch.Generator().manual_seed(11)
    net = P.ProxyNet(P_, None)
    net.load_state_dict(P.seeded_state_dict(net, 20 + P_))
    net = net.cuda()
    x0 = torch.rand(N, 3, H, W, generator=g).clamp(0.1, 0.9).cuda()
    p0 = torch.rand(N, P_, generator=g).cuda()
    dy = torch.randn(N, 3, H, W, generator=g).cuda()
    res = {}
    for mode in ('folded', 'materialised'):
        net.requires_grad_(mode == 'materialised')       # frozen -> folded table; trainable -> the reference's concat
        x, p = x0.clone().requires_grad_(), p0.clone().requires_grad_()
        y = net(x, p)
        res[mode] = (y,) + torch.autograd.grad(y, [x, p], dy)
    (yf, dxf, dpf), (ym, dxm, dpm) = res['folded'], res['materialised']
    assert maxabs(yf, ym) <= 2e-5 * max(1.0, float(ym.detach().abs().max())), (maxabs(yf, ym), float(ym.detach().abs().max()))
    relclose_relu_net(dxf, dxm, rtol=1e-4, flip_frac=2e-2)
    relclose(dpf, dpm, rtol=2e-3, atol=1e-5)


@pytest.mark.parametrize('active', [list(range(8)), [0, 3, 7], [5, 2]])
def test_srcnn_res_bank_matches_individual_proxies(active):
    """The grouped evaluation of a step's SRCNNRes proxies (one launch per layer for all members, shared statistics and input
    gradient paths) against the members evaluated one by one: outputs, d x and every parameter gradient."""
    from reconfigisp_b200.modules import tools_proxy as P
    g = torch.Generator().manual_seed(5)
    Ps = [2, 1, 2, 1, 3, 1, 3, 5]                      # reinhard, crysis, filmic, whiteworld, bilateral, median, fastnlm, bm3d
    nets = []
    for k, p_ in enumerate(Ps):
        net = P.ProxyNet(p_, None)
        net.load_state_dict(P.seeded_state_dict(net, 40 + k))
        nets.append(net.cuda().requires_grad_(False))
    N, H, W = 2, 37, 150
    x0 = torch.rand(N, 3, H, W, generator=g).clamp(0.1, 0.9).cuda()
    p0 = [torch.rand(1, Ps[s], generator=g).cuda() for s in active]
    dys = [torch.randn(N, 3, H, W, generator=g).cuda() for _ in active]
    bank = P.SRCNNResBank(nets)
    assert bank.usable(x0)
    x = x0.clone().requires_grad_()
    ps = [p.clone().requires_grad_() for p in p0]
    ys = bank(x, ps, active)
    got = torch.autograd.grad(ys, [x] + ps, dys)
    xr = x0.clone().requires_grad_()
    pr = [p.clone().requires_grad_() for p in p0]
    yr = [nets[s](xr, q.expand(N, -1)) for s, q in zip(active, pr)]
    ref = torch.autograd.grad(yr, [xr] + pr, dys)
    for a, b in zip(ys, yr):
        assert maxabs(a, b.detach()) <= 2e-5 * max(1.0, float(b.detach().abs().max()))
    relclose_relu_net(got[0], ref[0], rtol=1e-4, flip_frac=2e-2)
    for a, b in zip(got[1:], ref[1:]):
        relclose(a, b, rtol=2e-3, atol=1e-5)
