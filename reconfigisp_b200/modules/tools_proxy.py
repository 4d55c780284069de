"""CNN candidates: proxy networks and Path-Restore denoisers -- same class names, constructor
signatures, state-dict keys and `load()` behaviour as the reference's `tools_proxy.py` and the four
architecture files (`srcnn_res_arch.py`, `srcnn_demosaic_arch.py`, `path_14l_bayer_arch.py`,
`path_14l_bgr_arch.py`).

The parameter containers are `nn.Conv2d` modules laid out exactly like the reference's
`nn.Sequential`s (so checkpoints load with `strict=True`); the forward passes are written against
`conv.conv2d`, the dense-convolution entry of this package.
"""
import logging
from collections import OrderedDict

import torch
import torch.nn as nn

from .. import ops
from . import conv


def _strip_module_prefix(state):
    """tools_proxy.py:28-40: checkpoints saved from DataParallel carry a 'module.' prefix."""
    clean = OrderedDict()
    for k, v in state.items():
        clean[k[7:] if k.startswith('module.') else k] = v
    return clean


class _Loadable(nn.Module):
    def _maybe_load(self, load_path, strict_load):
        self.logger = logging.getLogger('base')
        if load_path is not None:
            self.load(load_path, strict_load)

    def load(self, load_path, strict_load=True):
        self.logger.info('Loading model for ProxyNet [{:s}] ...'.format(load_path))
        self.load_state_dict(_strip_module_prefix(torch.load(load_path, map_location='cpu')), strict=strict_load)


def seeded_state_dict(module, seed):
    """Deterministic stand-in weights (the trained checkpoints are not shipped): N(0,1)*0.05 drawn
    key by key from a CPU generator -- the recipe the oracle and the golden vectors use."""
    g = torch.Generator().manual_seed(seed)
    return OrderedDict((k, torch.randn(v.shape, generator=g) * 0.05) for k, v in module.state_dict().items())


class SRCNNRes(nn.Module):
    """srcnn_res_arch.py:6-53.  Input = x ++ [per-image min, mean, max of every channel] ++ params,
    broadcast over the frame; conv9-ReLU-conv5-ReLU-conv5; global residual."""

    def __init__(self, param_channel):
        super().__init__()
        self.srcnn = nn.Sequential(nn.Conv2d(3 + 9 + param_channel, 64, 9, 1, 4), nn.ReLU(),
                                   nn.Conv2d(64, 32, 5, 1, 2), nn.ReLU(), nn.Conv2d(32, 3, 5, 1, 2))

    def forward(self, x, param_vec):
        x = x.float()
        N, _, H, W = x.shape
        feat_min = x.amin(dim=(2, 3))
        feat_mean = torch.mean(torch.mean(x, dim=3), dim=2)
        feat_max = x.amax(dim=(2, 3))
        try:
            feat = torch.cat([feat_min, feat_mean, feat_max, param_vec], dim=1)
        except Exception:
            raise ValueError(feat_min.size(), None if param_vec is None else param_vec.size())
        feat_in = torch.cat([x, feat.view(N, -1, 1, 1).expand(N, feat.shape[1], H, W)], dim=1)
        c1, c2, c3 = self.srcnn[0], self.srcnn[2], self.srcnn[4]
        h = conv.conv2d_blocked(conv.to_blocked(feat_in), c1.weight, c1.bias, relu_out=True)
        h = conv.conv2d_blocked(h, c2.weight, c2.bias, relu_out=True)
        h = conv.conv2d_blocked(h, c3.weight, c3.bias, residual=conv.to_blocked(x))
        return conv.from_blocked(h, 3)


class SRCNNDemosaic(nn.Module):
    """srcnn_demosaic_arch.py:6-55: RGGB pack -> conv9-ReLU-conv1-ReLU-conv5 -> PixelShuffle(2)."""

    def __init__(self, param_channel):
        super().__init__()
        self.srcnn = nn.Sequential(nn.Conv2d(4 + param_channel, 64, 9, 1, 4), nn.ReLU(),
                                   nn.Conv2d(64, 32, 1, 1, 0), nn.ReLU(), nn.Conv2d(32, 12, 5, 1, 2),
                                   nn.PixelShuffle(2))

    def forward(self, x, param_vec):
        h = ops.pack_rggb(x.float())
        if param_vec is not None:
            N, _, h2, w2 = h.shape
            h = torch.cat([h, param_vec.view(N, -1, 1, 1).expand(N, param_vec.shape[1], h2, w2)], dim=1)
        c1, c2, c3 = self.srcnn[0], self.srcnn[2], self.srcnn[4]
        h = conv.conv2d_blocked(conv.to_blocked(h), c1.weight, c1.bias, relu_out=True)
        h = conv.conv2d_blocked(h, c2.weight, c2.bias, relu_out=True)
        h = conv.conv2d_blocked(h, c3.weight, c3.bias)
        return ops.pixel_shuffle2(conv.from_blocked(h, 12))


class ResidualBlock(nn.Module):
    """path_14l_bayer_arch.py:6-21.  Quirk kept: the block starts with an IN-PLACE ReLU, so the skip
    connection carries relu(x), not x."""

    def __init__(self, inchannel, outchannel, shortcut=None):
        super().__init__()
        self.basic = nn.Sequential(nn.ReLU(inplace=True), nn.Conv2d(inchannel, outchannel, 3, 1, 1),
                                   nn.ReLU(inplace=True), nn.Conv2d(outchannel, outchannel, 3, 1, 1))
        self.shortcut = shortcut

    def forward(self, x):
        """x and the result are channel-blocked tensors (see modules/conv.py)."""
        c1, c2 = self.basic[1], self.basic[3]
        t = conv.conv2d_blocked(x, c1.weight, c1.bias, relu_in=True, relu_out=True)
        return conv.conv2d_blocked(t, c2.weight, c2.bias, residual=x, residual_relu=True)


def _trunk(seq, h):
    """planar in, planar out; the 14 layers in between stay in the channel-blocked tensor-core layout"""
    first, blocks, last = seq[0], seq[1], seq[3]
    hb = conv.conv2d_blocked(conv.to_blocked(h), first.weight, first.bias)
    for blk in blocks:
        hb = blk(hb)
    hb = conv.conv2d_blocked(hb, last.weight, last.bias, relu_in=True)
    return conv.from_blocked(hb, last.weight.shape[0])


class Path14lBayer(nn.Module):
    """path_14l_bayer_arch.py:24-88; packed half-resolution, no global residual."""

    def __init__(self, param_channel):
        super().__init__()
        self.path_restore_14l = nn.Sequential(nn.Conv2d(4 + param_channel, 64, 3, 1, 1),
                                              nn.Sequential(*[ResidualBlock(64, 64) for _ in range(6)]),
                                              nn.ReLU(inplace=True), nn.Conv2d(64, 4, 3, 1, 1), nn.PixelShuffle(2))

    def forward(self, x, param_vec):
        h = ops.pack_rggb(x.float())
        if param_vec is not None:
            N, _, h2, w2 = h.shape
            h = torch.cat([h, param_vec.view(N, -1, 1, 1).expand(N, param_vec.shape[1], h2, w2)], dim=1)
        return ops.unpack_rggb(_trunk(self.path_restore_14l, h))


class Path14lBgr(nn.Module):
    """path_14l_bgr_arch.py:25-86; BGR -> RGB, trunk, RGB -> BGR."""

    def __init__(self, param_channel):
        super().__init__()
        self.path_restore_14l = nn.Sequential(nn.Conv2d(3 + param_channel, 64, 3, 1, 1),
                                              nn.Sequential(*[ResidualBlock(64, 64) for _ in range(6)]),
                                              nn.ReLU(inplace=True), nn.Conv2d(64, 3, 3, 1, 1))

    def forward(self, x, param_vec):
        h = x.float().flip(1)
        if param_vec is not None:
            N, _, H, W = h.shape
            h = torch.cat([h, param_vec.view(N, -1, 1, 1).expand(N, param_vec.shape[1], H, W)], dim=1)
        return _trunk(self.path_restore_14l, h).flip(1)


class ProxyNet(SRCNNRes, _Loadable):
    """tools_proxy.py:17-40."""

    def __init__(self, param_channel, load_path, strict_load=True):
        super().__init__(param_channel)
        self._maybe_load(load_path, strict_load)


class ProxyDemosaicNet(SRCNNDemosaic, _Loadable):
    """tools_proxy.py:43-66."""

    def __init__(self, param_channel, load_path, strict_load=True):
        super().__init__(param_channel)
        self._maybe_load(load_path, strict_load)


class PathRestore14lBayer(Path14lBayer, _Loadable):
    """tools_proxy.py:69-92."""

    def __init__(self, param_channel, load_path, strict_load=True):
        super().__init__(param_channel)
        self._maybe_load(load_path, strict_load)


class PathRestore14lBgr(Path14lBgr, _Loadable):
    """tools_proxy.py:95-118."""

    def __init__(self, param_channel, load_path, strict_load=True):
        super().__init__(param_channel)
        self._maybe_load(load_path, strict_load)


_STANDIN = {}


def demosaicnet_standin(device):
    """DemosaicNet's weights are external to the reference; the default stand-in is a seeded
    SRCNNDemosaic (oracle/SPEC.md).  Replace with `isp_kernels.demosaic.register_demosaicnet`."""
    key = str(device)
    if key not in _STANDIN:
        net = ProxyDemosaicNet(0, None)
        net.load_state_dict(seeded_state_dict(net, 4))
        for q in net.parameters():
            q.requires_grad_(False)
        _STANDIN[key] = net.to(device)
    return _STANDIN[key]
