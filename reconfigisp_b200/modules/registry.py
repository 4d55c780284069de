"""Stage registry and architecture-string grammar (isp_universal.py:62-208, origin_universal.py:36-141,
super_prune_fifteen_demos_four_bayer_two.py:35-171), as data.

  architecture := token ('_' token)*          token := 'Bayer' | 'Demosaic' | 'sRGB' | <1-based index>
  A domain token switches the pool; every index token appends one stage; `step` counts stages from 1
  across domains and names the state-dict key  param_step{step}_{name}.
"""
from collections import namedtuple

import numpy as np

from . import tools_origin as T
from . import tools_proxy as P

DOMAINS = ('Bayer', 'Demosaic', 'sRGB')

# proxy checkpoints below module_path (isp_universal.py:34-52); value = (param channels, relative path)
_EXP = 'proxy_nets/experiments/'
PROXY_CKPT = {
    'reinhard': (2, _EXP + '006_reinhard_residual_multistepLR2/models/400000_G.pth'),
    'crysisengine': (1, _EXP + '007_crysis_residual_multistepLR/models/400000_G.pth'),
    'filmic': (2, _EXP + '009_filmic_residual_multistepLR/models/400000_G.pth'),
    'whiteworld': (1, _EXP + '008_whiteworld_residual_multistepLR/models/400000_G.pth'),
    'bilateral': (3, _EXP + '013_bilateral_residual_multistepLR2/models/400000_G.pth'),
    'median': (1, _EXP + '010_median_residual_multistepLR/models/400000_G.pth'),
    'fastnlm': (3, _EXP + '014_fastnlm_residual_multistepLR2/models/400000_G.pth'),
    'bilinear': (0, _EXP + '015_demosaic_bilinear_multistepLR/models/400000_G.pth'),
    'laplacian': (0, _EXP + '016_demosaic_laplacian_multistepLR/models/400000_G.pth'),
    'path_bayer': (0, _EXP + '020_denoise_path_restore_14l_bayer_aug_multistepLR/models/800000_G.pth'),
    'path_bgr': (0, _EXP + '019_path_restore_14l_rgb/models/path_restore_14l_rgb.pth'),
    'bm3d': (5, _EXP + '022_bm3d_residual_multistepLR_mc/models/400000_G.pth'),
}

_WBQ = [0.0] * 30
_WBQ[6] = _WBQ[17] = _WBQ[28] = 0.406
# default logits (isp_universal.py:102-127): sigmoid -> identity-ish module parameters
DEFAULT_LOGITS = {
    'gamma': [0.], 'reinhard': [0., 0.], 'crysisengine': [0.], 'filmic': [0., 0.], 'grayworld': [], 'whiteworld': [0.],
    'bilateral': [0., 0., 0.], 'median': [0.], 'fastnlm': [0., 0., 0.], 'skip': [], 'wbmanual': [-1.38, -1.38, -1.38],
    'path_bgr': [], 'wbquadratic': _WBQ, 'gtmmanual': [-1.099, 0, 1.099], 'bm3d': [-1.946, 1.099, -1.099, -1.099, 2.708],
    'conditional_gamma': [0.], 'conditional_wb_manual': [-1.38, -1.38, -1.38], 'conditional_wb_quadratic': _WBQ,
    'path_bayer': [], 'nearest': [], 'bilinear': [], 'laplacian': [], 'demosaicnet': [],
}

NAMES = {
    'Bayer': ['path_bayer', 'skip'],
    'Demosaic': ['nearest', 'bilinear', 'laplacian', 'demosaicnet'],
    'sRGB': ['gamma', 'reinhard', 'crysisengine', 'filmic', 'grayworld', 'whiteworld', 'bilateral', 'median', 'fastnlm',
             'skip', 'wbmanual', 'path_bgr', 'wbquadratic', 'gtmmanual', 'bm3d',
             'conditional_gamma', 'conditional_wb_manual', 'conditional_wb_quadratic',     # IspUniversal only, 16-18
             'ten_layer_net', 'two_layer_net', 'toy_net'],                                 # 19-21: undefined upstream
}
N_SRGB_ORIGIN = 15        # OriginUniversal / the supernets stop at bm3d
CONDITIONAL = {'conditional_gamma': (T.ConditionalGamma, 'gamma_in_channels'),
               'conditional_wb_manual': (T.ConditionalWbManual, 'wb_manual_in_channels'),
               'conditional_wb_quadratic': (T.ConditionalWbQuadratic, 'wb_quadratic_in_channels')}

CLASSICAL = {'gamma': T.Gamma, 'grayworld': T.Grayworld, 'skip': T.Skip, 'wbmanual': T.WbManual,
             'wbquadratic': T.WbQuadratic, 'nearest': T.DemosaicNearest, 'demosaicnet': T.DemosaicNet}
ORIGIN = {'reinhard': T.OriginToneReinhard, 'crysisengine': T.OriginToneCrysis, 'filmic': T.OriginToneFilmic,
          'whiteworld': T.OriginWbWhiteworld, 'bilateral': T.OriginNoiseBilateral, 'median': T.OriginNoiseMedian,
          'fastnlm': T.OriginNoiseFastnlm, 'bilinear': T.OriginDemosBilinear, 'laplacian': T.OriginDemosLaplacian}
NETS = {'path_bayer': P.PathRestore14lBayer, 'path_bgr': P.PathRestore14lBgr, 'bilinear': P.ProxyDemosaicNet,
        'laplacian': P.ProxyDemosaicNet}

Stage = namedtuple('Stage', 'step domain name')


def parse_architecture(architecture, n_srgb=len(NAMES['sRGB'])):
    """'Bayer_01_Demosaic_03_sRGB_01_13_11' -> [Stage(1,'Bayer','path_bayer'), ...]  (same errors as the
    reference: ValueError without a leading domain, AssertionError for an index out of range)."""
    domain, step, stages = None, 0, []
    for tok in architecture.split('_'):
        if tok in DOMAINS:
            domain = tok
            continue
        if domain is None:
            raise ValueError('Domain (Bayer, Demosaic, sRGB) is not specified in ISP architecture!')
        step += 1
        limit = n_srgb if domain == 'sRGB' else len(NAMES[domain])
        idx = int(tok)
        assert 1 <= idx <= limit
        stages.append(Stage(step, domain, NAMES[domain][idx - 1]))
    return stages


def build_net(name, ckpt_path, weight_seed=None):
    """A CNN candidate.  `ckpt_path` None (or weight_seed given) -> seeded stand-in weights."""
    n_par = PROXY_CKPT[name][0]
    cls = NETS.get(name, P.ProxyNet)
    if weight_seed is not None:
        net = cls(n_par, None)
        net.load_state_dict(P.seeded_state_dict(net, weight_seed))
    else:
        net = cls(n_par, ckpt_path)
    for q in net.parameters():          # candidate-net weights are frozen constants
        q.requires_grad_(False)
    return net


def conditional_init(total, global_logits, rng=np.random):
    """isp_universal.py:184-190: N(0, 0.01) FC weights followed by the global logits."""
    return list(rng.randn(total - len(global_logits)) * 0.01) + list(global_logits)
