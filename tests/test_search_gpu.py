"""GPU parity of the DARTS second-order step (search.DartsModel) against the CPU oracle's restatement of
darts_model.py:159-324, and of the tiled-inference path against the reference's patch2whole semantics."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import isp_oracle as O            # noqa: E402
from oracle import pipeline_oracle as PO      # noqa: E402


def relclose(a, b, rtol=3e-3, atol=1e-7):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    assert a.shape == b.shape
    assert float((a - b).abs().max()) <= atol + rtol * float(b.abs().max()), (float((a - b).abs().max()), float(b.abs().max()))


def _opt(n_step):
    return {'model': 'darts', 'network_G': {'which_model_G': 'SuperPruneFifteenDemosFourBayerTwo', 'n_step': n_step, 'n_modules': 15,
                                           'prune_threshold': 0.2, 'weight_seed': 10},
            'train': {'lr_G': 0.01, 'momentum_G': 0.9, 'lr_meta': 0.01, 'beta1': 0.9, 'beta2': 0.999, 'pixel_criterion': 'l2'}}


def test_darts_step_matches_oracle():
    from reconfigisp_b200.search import DartsModel
    g = torch.Generator().manual_seed(10)
    N, H, W, n_step = 2, 16, 16, 1
    img, vimg = torch.rand(N, 1, H, W, generator=g) * 0.8 + 0.1, torch.rand(N, 1, H, W, generator=g) * 0.8 + 0.1
    gt, vgt = torch.rand(N, 3, H, W, generator=g), torch.rand(N, 3, H, W, generator=g)
    alphas0 = [torch.randn(k, generator=g) * 0.5 for k in (2, 4, 15)]
    # oracle
    oG, oV = PO.Supernet(n_step, 0.2, 10), PO.Supernet(n_step, 0.2, 10)
    with torch.no_grad():
        for a, v in zip(oG.alphas, alphas0):
            a.copy_(v)
    ref = PO.darts_step(oG, oV, img, gt, vimg, vgt, lr_G=0.01, momentum_G=0.9, lr_meta=0.01)
    # B200
    m = DartsModel(_opt(n_step))
    with torch.no_grad():
        for a, v in zip(m.netG.alphas, alphas0):
            a.copy_(v.cuda())
    m.feed_data((img, gt, vimg, vgt))
    m.optimize_alphas()
    ref_val = ref['val_loss']
    ref_val = float(ref_val.detach()) if torch.is_tensor(ref_val) else float(ref_val)
    assert abs(float(m.val_loss) - ref_val) <= 1e-5
    for a, r in zip(m.netG.alphas, ref['alpha_grad']):
        relclose(a.grad, r)
    # parameter step with the alphas restored (the oracle did not apply the Adam update)
    with torch.no_grad():
        for a, v in zip(m.netG.alphas, alphas0):
            a.copy_(v.cuda())
    m.optimize_parameters()
    nz = [p for p in m.netG.trainable_parameters if p.nelement() > 0]
    assert abs(float(m.log_dict['loss'].detach()) - float(ref['loss2'].detach())) <= 1e-5
    for p, r in zip(nz, ref['param_grad']):
        if r is None:
            continue
        relclose(p.grad, r)


def test_split_inference_matches_reference_blend():
    from reconfigisp_b200 import ops
    from reconfigisp_b200.patch import split_inference
    from reconfigisp_b200.modules.origin_universal import OriginUniversal
    g = torch.Generator().manual_seed(4)
    H, W, ps, st = 120, 136, 48, 40
    raw = torch.rand(1, 1, H, W, generator=g)
    net = OriginUniversal('/x/', 'Bayer_02_Demosaic_03_sRGB_01_13_11', weight_seed=10).cuda()
    out = split_inference(lambda t: net(t), raw.cuda(), ps, st, chunk=4)
    # oracle: the reference's numpy split / per-tile model / blend (test_split.py:82-108)
    pipe = PO.FixedPipeline('Bayer_02_Demosaic_03_sRGB_01_13_11', 'origin', 10)
    img = raw[0].permute(1, 2, 0).numpy()
    patches, pos, cnt = O.whole2patch(img, (ps, ps), (st, st))
    outs = [pipe.forward(torch.from_numpy(p).permute(2, 0, 1).unsqueeze(0))[0][0].permute(1, 2, 0).detach().numpy() for p in patches]
    merged = np.clip(O.patch2whole(np.array(outs), pos, cnt, (st, st)), 0, 1)
    assert float(np.abs(out.permute(1, 2, 0).cpu().numpy() - merged).max()) <= 1e-4


def _opt_ft(n_step, ft_steps=1):
    o = _opt(n_step)
    o['model'] = 'darts_ft'
    o['network_G']['which_model_G'] = 'SuperPruneFifteenDemosFourBayerTwoFt'
    o['proxy_ft_params'] = {'memory_size': 4, 'ft_steps': ft_steps}
    return o


def test_proxy_finetune_step_matches_oracle():
    """darts_ft_model.py:185-231: one fine-tune step per flagged proxy -- loss and weight/bias gradients against
    the oracle's SRCNNRes + classical targets under autograd, with the reference's host RNG draws."""
    import random
    from reconfigisp_b200.networks import create_model
    n_step = 2
    m = create_model(_opt_ft(n_step))
    assert [n for n, *_ in m.ft_nets] == ['crysisengine', 'whiteworld', 'bilateral', 'median', 'fastnlm']
    g = torch.Generator().manual_seed(4)
    mem = [torch.rand(2, 3, 24, 32, generator=g) * 0.9 + 0.05 for _ in range(3)]
    m.ft_data = [t.cuda() for t in mem]
    before = {n: [p.detach().clone() for p in proxy.parameters()] for n, proxy, *_ in m.ft_nets}
    random.seed(3); torch.manual_seed(3)
    m.finetune_proxies()
    # oracle replay
    random.seed(3); torch.manual_seed(3)
    oS = PO.Supernet(n_step, 0.2, 10)
    names = [n for n, _ in PO.SRGB_STEP]
    for name, proxy, _, _ in m.ft_nets:
        data = mem[int(random.random() * len(mem))]
        param = torch.rand(1, oS.steps[-1][names.index(name)].P).repeat(data.shape[0], 1)
        sd = {k: v.clone().requires_grad_() for k, v in oS.steps[-1][names.index(name)].sd.items()}
        out = O.srcnn_res(data, param, sd)
        gt = PO.ORIGIN[name](data, param)
        loss = O.mse(out, gt)
        keys = [k for k, _ in proxy.named_parameters()]
        grads = torch.autograd.grad(loss, [sd[k] for k in keys])
        assert abs(float(m.log_dict['ft_' + name]) - float(loss)) <= 2e-5 * max(1.0, float(loss)), name
        for k, p, r in zip(keys, proxy.parameters(), grads):
            relclose(p.grad, r, rtol=3e-3)
        # Adam moved the weights, and load_proxy_nets copied them into every sRGB step
        idx = names.index(name)
        for p0, p1, p2 in zip(before[name], m.netG.all_modules[-1][idx].parameters(), m.netG.all_modules[-2][idx].parameters()):
            assert float((p1 - p0).abs().max()) > 0
            assert torch.equal(p1, p2)


def test_proxy_finetune_reduces_proxy_error_and_search_continues():
    """The online loop of train_ft.py: search iterations fill the FIFO, fine-tuning lowers the proxy-vs-original
    error, and the tuned weights are the ones the next supernet forward uses."""
    import random
    from reconfigisp_b200.networks import create_model
    from reconfigisp_b200 import ops
    opt = _opt_ft(1, ft_steps=12)
    opt['train']['lr_G'] = 1e-3
    m = create_model(opt)
    g = torch.Generator().manual_seed(6)
    batch = lambda: (torch.rand(2, 1, 32, 32, generator=g) * 0.8 + 0.1, torch.rand(2, 3, 32, 32, generator=g),
                     torch.rand(2, 1, 32, 32, generator=g) * 0.8 + 0.1, torch.rand(2, 3, 32, 32, generator=g))
    for _ in range(3):
        m.feed_data(batch()); m.optimize_alphas(); m.optimize_parameters()
    assert 0 < len(m.ft_data) <= 4 and all(t.shape[1] == 3 for t in m.ft_data)
    name, proxy, target, _ = m.ft_nets[0]
    data, param = m.ft_data[-1], torch.full((2, 1), 0.5).cuda()
    with torch.no_grad():
        e0 = float(ops.mse_loss(proxy(data, param), target(data, param)))
    random.seed(0); torch.manual_seed(0)
    for _ in range(4):
        m.finetune_proxies()
    with torch.no_grad():
        e1 = float(ops.mse_loss(proxy(data, param), target(data, param)))
    assert e1 < e0, (e0, e1)
    m.feed_data(batch()); m.optimize_alphas(); m.optimize_parameters()
    assert torch.isfinite(m.log_dict['loss'])


def test_darts_model_matches_reference_run(golden):
    """Two full search iterations (optimize_alphas + optimize_parameters) of `search.DartsModel`, n_step = 3 with two
    candidates pruned, against the golden recorded from the reference's own `DartsModel` (oracle/gen_golden_darts.py):
    validation loss, alpha gradients, alphas after Adam, training loss, parameter gradients and parameters after
    SGD-momentum -- after BOTH iterations (the second one exercises the momentum buffer in the virtual step, :212-218)."""
    from reconfigisp_b200.search import DartsModel
    g = golden('darts_model')
    T = torch.from_numpy
    o = _opt(3)
    m = DartsModel(o)
    with torch.no_grad():
        for i, a in enumerate(m.netG.alphas):
            a.copy_(T(g['alpha0_%d' % i]).cuda())
    m.feed_data((T(g['img']), T(g['gt']), T(g['vimg']), T(g['vgt'])))
    for it in range(2):
        m.optimize_alphas()
        assert abs(float(m.val_loss) - float(g['it%d_val_loss' % it])) <= 2e-5, it
        for i, a in enumerate(m.netG.alphas):
            relclose(a.grad, T(g['it%d_alpha_grad_%d' % (it, i)]), rtol=5e-3, atol=2e-7)
            relclose(a, T(g['it%d_alpha_%d' % (it, i)]), rtol=1e-3, atol=1e-5)
        m.optimize_parameters()
        assert m.netG.pruned_paths == list(g['it%d_pruned' % it]), it
        assert abs(float(m.log_dict['loss'].detach()) - float(g['it%d_loss' % it])) <= 2e-5, it
        nz = [p for p in m.netG.trainable_parameters if p.nelement() > 0]
        assert len(nz) == int(g['n_params'])
        for i, p in enumerate(nz):
            relclose(p.grad, T(g['it%d_param_grad_%d' % (it, i)]), rtol=5e-3, atol=2e-7)
            relclose(p, T(g['it%d_param_%d' % (it, i)]), rtol=1e-4, atol=1e-6)


def test_finetune_proxies_matches_reference_run(golden):
    """`search_ft.DartsFtModel`: one training pass (fills the FIFO) + `finetune_proxies()` with ft_steps = 2 against the golden
    recorded from the reference's own `DartsFtModel` (oracle/gen_golden_darts_ft.py): training loss, the loss of BOTH
    fine-tune steps of every flagged proxy (the second one sees the weights Adam produced from the first) and the first-step
    weight / bias gradients (first 24 values, sum, abs-sum of every tensor) -- same host RNG draws as the reference."""
    import random
    from reconfigisp_b200.networks import create_model
    g = golden('darts_ft')
    T = torch.from_numpy
    o = _opt_ft(3, ft_steps=2)
    o['proxy_ft_params']['memory_size'] = 8
    o['train']['lr_G'] = float(g['lr'])
    m = create_model(o)
    names = [str(n) for n in g['names']]
    assert [n for n, *_ in m.ft_nets] == names
    m.feed_data((T(g['img']), T(g['gt']), T(g['vimg']), T(g['vgt'])))
    m.optimize_parameters()
    assert abs(float(m.log_dict['loss'].detach()) - float(g['loss_G'])) <= 2e-5
    assert len(m.ft_data) == int(g['n_ft_data'])
    losses, grads = [], {n: [] for n in names}
    loss_fn = m._loss
    m._loss = lambda a, b: (losses.append(loss_fn(a, b)) or losses[-1])

    def summary(t):
        f = t.detach().reshape(-1).double().cpu()
        head = f[:24].float() if f.numel() >= 24 else torch.nn.functional.pad(f.float(), (0, 24 - f.numel()))
        return torch.cat([head, torch.tensor([float(f.sum()), float(f.abs().sum())])])
    for name, proxy, _, optim in m.ft_nets:
        orig = optim.step

        def step(orig=orig, proxy=proxy, name=name):
            grads[name].append(torch.stack([summary(p.grad) for p in proxy.parameters()]))
            return orig()
        optim.step = step
    random.seed(int(g['ft_seed'])); torch.manual_seed(int(g['ft_seed']))
    m.finetune_proxies()
    got = torch.stack([l.detach().cpu() for l in losses]).view(len(names), 2)
    ref = T(g['losses'])
    for k, name in enumerate(names):
        assert abs(float(got[k, 0]) - float(ref[k, 0])) <= 2e-5 * max(1.0, float(ref[k, 0])), (name, float(got[k, 0]), float(ref[k, 0]))
        # step 2 runs on Adam-updated weights: every weight moved by ~lr with the SIGN of its gradient, which is fragile for
        # near-zero gradients, so the bar is relative
        assert abs(float(got[k, 1]) - float(ref[k, 1])) <= 2e-2 * float(ref[k, 1]), (name, float(got[k, 1]), float(ref[k, 1]))
        rg = T(g['grads_' + name])[0]
        for i in range(rg.shape[0]):
            a, b = grads[name][0][i], rg[i]
            assert float((a[:24] - b[:24]).abs().max()) <= 3e-3 * max(1e-6, float(b[:24].abs().max())), (name, i)
            assert abs(float(a[25]) - float(b[25])) <= 3e-3 * float(b[25]) + 1e-7, (name, i)
