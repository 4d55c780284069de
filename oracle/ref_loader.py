"""Loader that makes the reference's own Python importable on CPU  --  TEST INFRASTRUCTURE.

Used ONLY by `oracle/gen_golden.py` (in the build container, where `/root/reference`
exists) to run the reference's in-repo code and record golden vectors.  Nothing in the
product, the `-m gpu` tests, `smoke()` or `bench.py` touches this file at run time.

Recipe:
  1. put `<reference>/codes` on sys.path (the code uses absolute package names `models.*`);
  2. register the five un-shipped kernel modules (`tools_origin.py:13-17`) in `sys.modules`,
     each `run(img, option, params)` dispatching to `oracle/isp_oracle.py`;
  3. neutralise the hard-coded CUDA placement (`origin_universal.py:22`, `super_prune…:33`);
  4. replace the `load()` of the four proxy-net classes by deterministic seeded weights
     (the trained checkpoints under /DATA/module are not shipped).
"""
import contextlib
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

from . import isp_oracle as O

REFERENCE_ROOT = os.environ.get('RISP_REFERENCE_ROOT', '/root/reference')


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'codes', 'models'))


# -- oracle-backed stand-ins for /DATA/ISP_Kernels --------------------------------------
def _nhwc(fn):
    """Wrap an NCHW oracle function for the NHWC calling convention of the wrappers."""
    return lambda img, *a: fn(img.permute(0, 3, 1, 2), *a).permute(0, 2, 3, 1)


class _WhiteBalance:
    def run(self, img, option, params):
        if option == 'grayworld':
            return _nhwc(O.wb_grayworld)(img)
        if option == 'manual':
            return _nhwc(O.wb_manual)(img, params['gain'])
        if option == 'whiteworld':
            return _nhwc(O.wb_whiteworld)(img, params['white_point_ratio'])
        raise ValueError(option)


class _Gamma:
    def run(self, img, option, params):
        assert option == 'manual'
        return _nhwc(O.gamma_manual)(img, params['gamma'])


class _Demosaic:
    def run(self, img, option, params):
        if option == 'nearestneighbor':
            return O.demosaic_nearest(img)                       # NCHW in, NCHW out
        if option == 'bilinear':
            return _nhwc(O.demosaic_bilinear)(img)
        if option == 'laplacian':
            return _nhwc(lambda r: O.demosaic_laplacian(r, 255.0))(img)
        if option == 'demosaicnet':
            # external network, weights not shipped: SRCNNDemosaic-architecture stand-in with
            # the fixed seeded weights of `isp_oracle.demosaicnet_standin_state` (oracle/SPEC.md)
            return O.srcnn_demosaic(img, O.demosaicnet_standin_state())
        raise ValueError(option)


class _GlobalToneMapping:
    def run(self, img, option, params):
        if option == 'reinhard':
            return _nhwc(O.tone_reinhard)(img, params['white_point'], params['middle_grey'])
        if option == 'crysisengine':
            return _nhwc(O.tone_crysis)(img, params['lum_adapted'])
        if option == 'filmic':
            return _nhwc(O.tone_filmic)(img, params['white_point'], params['exposure_bias'])
        raise ValueError(option)


class _SpatialNoiseReduction:
    def run(self, img, option, params):
        if option == 'bilateral':
            return _nhwc(O.denoise_bilateral)(img, params['window_length'], params['sigma_color'],
                                              params['sigma_space'])
        if option == 'median':
            return _nhwc(O.denoise_median)(img, params['size'])
        if option == 'fastnlm':
            return _nhwc(O.denoise_fastnlm)(img, params['block_size'], params['search_block'], params['decay_factor'])
        raise ValueError(option)


def _install_kernel_stubs():
    for name, cls_name, cls in (('whitebalance', 'WhiteBalance', _WhiteBalance), ('gamma', 'Gamma', _Gamma),
                                ('demosaic', 'Demosaic', _Demosaic),
                                ('globaltonemapping', 'GlobalToneMapping', _GlobalToneMapping),
                                ('spatialnoisereduction', 'SpatialNoiseReduction', _SpatialNoiseReduction)):
        m = types.ModuleType(name)
        setattr(m, cls_name, cls)
        sys.modules[name] = m


def seeded_state_dict(module, seed):
    """Deterministic weights independent of nn.init: N(0,1)*0.05 drawn key by key."""
    g = torch.Generator().manual_seed(seed)
    return {k: torch.randn(v.shape, generator=g) * 0.05 for k, v in module.state_dict().items()}


@contextlib.contextmanager
def cpu_only():
    """Make `.to(cuda)`, `.cuda()` no-ops while reference containers are built / run."""
    orig_mod_to, orig_t_cuda, orig_t_to = nn.Module.to, torch.Tensor.cuda, torch.Tensor.to

    def t_to(self, *a, **k):
        if a and isinstance(a[0], (torch.device, str)) and 'cuda' in str(a[0]):
            return self
        return orig_t_to(self, *a, **k)

    nn.Module.to = lambda self, *a, **k: self
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.Tensor.to = t_to
    try:
        yield
    finally:
        nn.Module.to, torch.Tensor.cuda, torch.Tensor.to = orig_mod_to, orig_t_cuda, orig_t_to


def load_reference(weight_seed=10):
    """Returns a namespace with the reference modules importable on CPU."""
    assert reference_available(), 'reference tree not present'
    codes = os.path.join(REFERENCE_ROOT, 'codes')
    if codes not in sys.path:
        sys.path.insert(0, codes)
    _install_kernel_stubs()
    import models.modules.tools_origin as tools_origin
    import models.modules.tools_proxy as tools_proxy

    counter = {'n': 0}

    def fake_load(self, load_path, strict_load=True):
        # one deterministic weight set per network instance, in construction order
        self.load_state_dict(seeded_state_dict(self, weight_seed + counter['n']))
        counter['n'] += 1

    for cls in (tools_proxy.ProxyNet, tools_proxy.ProxyDemosaicNet,
                tools_proxy.PathRestore14lBayer, tools_proxy.PathRestore14lBgr):
        cls.load = fake_load

    import models.modules.super_prune_fifteen_demos_four_bayer_two as sp
    import models.modules.super_prune_fifteen_demos_four_bayer_two_ft as sp_ft
    import models.modules.origin_universal as ou
    # isp_universal.py:92-94 names three classes that no file defines
    import builtins
    for nm in ('TenLayerNet', 'TwoLayerNet', 'ToyNet'):
        if not hasattr(builtins, nm):
            setattr(builtins, nm, type(nm, (nn.Module,), {}))
    import models.modules.isp_universal as iu
    # util_path_restore.py:3 imports skimage (absent here) for an unrelated SSIM helper
    if 'skimage' not in sys.modules:
        sk, skm = types.ModuleType('skimage'), types.ModuleType('skimage.measure')
        skm.compare_ssim = None
        sk.measure = skm
        sys.modules['skimage'], sys.modules['skimage.measure'] = sk, skm
    import utils.util_path_restore as upr
    return types.SimpleNamespace(tools_origin=tools_origin, tools_proxy=tools_proxy, counter=counter,
                                 super_prune=sp, super_prune_ft=sp_ft, origin_universal=ou,
                                 isp_universal=iu, util_path_restore=upr)
