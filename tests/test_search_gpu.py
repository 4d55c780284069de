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
    assert abs(float(m.val_loss) - float(ref['val_loss'])) <= 1e-5
    for a, r in zip(m.netG.alphas, ref['alpha_grad']):
        relclose(a.grad, r)
    # parameter step with the alphas restored (the oracle did not apply the Adam update)
    with torch.no_grad():
        for a, v in zip(m.netG.alphas, alphas0):
            a.copy_(v.cuda())
    m.optimize_parameters()
    nz = [p for p in m.netG.trainable_parameters if p.nelement() > 0]
    assert abs(float(m.log_dict['loss']) - float(ref['loss2'])) <= 1e-5
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
