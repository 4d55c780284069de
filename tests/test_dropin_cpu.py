"""Boundary 1 drop-in: the reference's UNMODIFIED `codes/models/modules/tools_origin.py`
imported over this repository's plugin directory `reconfigisp_b200/isp_kernels/` (the five module names it resolves through
sys.path, tools_origin.py:8-17).  Runs in the build container only (needs /root/reference, skipped elsewhere) and without a
GPU: `reconfigisp_b200.ops` -- what the plugins call -- is patched to the CPU oracle, so what is exercised is the contract
between the reference wrappers and the plugins: option names, NHWC / NCHW layouts, 0-255 scaling, numpy vs tensor
parameters, integer window tensors, dict descriptors.  Every wrapper class of the file is called once and its result is
compared with the oracle applied to the wrapper's documented parameter mapping."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from oracle import isp_oracle as O
from oracle import ref_loader as RL

pytestmark = pytest.mark.skipif(not RL.reference_available(), reason='reference tree not present (build container only)')


@pytest.fixture(scope='module')
def tools_origin(request):
    import reconfigisp_b200.ops as ops
    from reconfigisp_b200 import isp_kernels
    saved = {k: getattr(ops, k) for k in ('gamma', 'gain', 'grayworld', 'whiteworld', 'demosaic', 'tone_reinhard', 'tone_crysis',
                                          'tone_filmic', 'bilateral', 'median', 'fastnlm')}
    calls = []

    def log(name, fn):
        def w(*a, **k):
            calls.append(name)
            return fn(*a, **k)
        return w
    ops.gamma = log('gamma', lambda x, g: O.gamma_manual(x, g))
    ops.gain = log('gain', lambda x, g: O.wb_manual(x, g))
    ops.grayworld = log('grayworld', lambda x: O.wb_grayworld(x))
    ops.whiteworld = log('whiteworld', lambda x, r, s=1.0: O.wb_whiteworld(x, r))
    ops.demosaic = log('demosaic', lambda raw, kind, clip_hi=1.0: {'nearest': O.demosaic_nearest, 'bilinear': O.demosaic_bilinear,
                                                                  'malvar': lambda r: O.demosaic_laplacian(r, clip_hi)}[kind](raw))
    ops.tone_reinhard = log('reinhard', lambda x, wp, mg, s=1.0: O.tone_reinhard(x, wp, mg))
    ops.tone_crysis = log('crysis', lambda x, la, s=1.0: O.tone_crysis(x, la))
    ops.tone_filmic = log('filmic', lambda x, wp, eb, s=1.0: O.tone_filmic(x, wp, eb))
    ops.bilateral = log('bilateral', lambda x, win, sc, ss, max_window=None: O.denoise_bilateral(x, win, sc, ss))
    ops.median = log('median', lambda x, k: O.denoise_median(x, k))
    ops.fastnlm = log('fastnlm', lambda x, b, s, h, max_halo=None: O.denoise_fastnlm(x, b, s, h))
    here = isp_kernels.install()                      # what a deployment does: put the plugin directory on sys.path
    for name in ('whitebalance', 'gamma', 'demosaic', 'globaltonemapping', 'spatialnoisereduction'):
        sys.modules.pop(name, None)
    codes = os.path.join(RL.REFERENCE_ROOT, 'codes')
    sys.path.insert(0, codes)
    sys.modules.pop('models.modules.tools_origin', None)
    with RL.cpu_only():
        T = importlib.import_module('models.modules.tools_origin')
    assert os.path.dirname(sys.modules['whitebalance'].__file__) == here      # OUR plugins were resolved
    assert T.__file__.startswith(RL.REFERENCE_ROOT)                           # the reference's own, unmodified file

    def restore():
        for k, v in saved.items():
            setattr(ops, k, v)
        for name in ('whitebalance', 'gamma', 'demosaic', 'globaltonemapping', 'spatialnoisereduction', 'models.modules.tools_origin'):
            sys.modules.pop(name, None)
        sys.path.remove(codes)
    request.addfinalizer(restore)
    T._calls = calls
    return T


def close(a, b, tol=1e-5):
    assert a.shape == b.shape, (a.shape, b.shape)
    assert float((a.detach() - b.detach()).abs().max()) <= tol, float((a.detach() - b.detach()).abs().max())


def test_reference_wrappers_run_on_our_plugins(tools_origin):
    T = tools_origin
    g = torch.Generator().manual_seed(2)
    N, H, W = 2, 12, 16
    x = torch.rand(N, 3, H, W, generator=g)
    raw = torch.rand(N, 1, H, W, generator=g)
    with RL.cpu_only():
        # differentiable wrappers, image in [0,1]
        p1 = torch.rand(N, 1, generator=g)
        close(T.Gamma()(x, p1), O.gamma_manual(x, p1))
        close(T.Grayworld()(x, None), O.wb_grayworld(x))
        p3 = torch.rand(N, 3, generator=g)
        close(T.WbManual()(x, p3), O.wb_manual(x, p3 * 5))                               # gain = p*5 (:214)
        close(T.DemosaicNearest()(raw, None), O.demosaic_nearest(raw), 0)
        # Origin* wrappers: x255, numpy parameters, /255 (:462-473, :535-546, :655-662, :696-710, :742-751, :785-797)
        close(T.OriginDemosBilinear()(raw, None), O.demosaic_bilinear(raw * 255) / 255)
        close(T.OriginDemosLaplacian()(raw, None), O.demosaic_laplacian(raw * 255, 255.) / 255)
        p2 = torch.rand(N, 2, generator=g)
        close(T.OriginToneReinhard()(x, p2), O.tone_reinhard(x * 255, p2[:, 0], p2[:, 1]) / 255)
        close(T.OriginToneCrysis()(x, p1), O.tone_crysis(x * 255, p1[:, 0]) / 255)
        close(T.OriginToneFilmic()(x, p2), O.tone_filmic(x * 255, p2[:, 0], p2[:, 1] * 9 + 1) / 255)   # exposure = p*9+1 (:613)
        close(T.OriginWbWhiteworld()(x, p1), O.wb_whiteworld(x * 255, p1[:, 0]) / 255)
        pb = torch.rand(N, 3, generator=g)
        win = O.bilateral_window_from_param(pb[:, 0])                                                    # .int()*7 quirk (:698)
        close(T.OriginNoiseBilateral()(x, pb), O.denoise_bilateral(x * 255, win, pb[:, 1] * 99 + 1, pb[:, 2] * 99 + 1) / 255)
        pm = torch.tensor([[0.45], [0.9]])
        close(T.OriginNoiseMedian()(x, pm), O.denoise_median(x * 255, O.median_size_from_param(0.45)) / 255, 0)   # batch row 0 (:744)
        pn = torch.rand(N, 3, generator=g)
        blk = O.bilateral_window_from_param(pn[:, 0])
        srch = O.bilateral_window_from_param(pn[:, 1])
        close(T.OriginNoiseFastnlm()(x, pn), O.denoise_fastnlm(x * 255, blk, srch, pn[:, 2] * 99 + 1) / 255)
    # every plugin family was reached through the reference's own call sites
    assert set(T._calls) >= {'gamma', 'gain', 'grayworld', 'whiteworld', 'demosaic', 'reinhard', 'crysis', 'filmic', 'bilateral',
                             'median', 'fastnlm'}
