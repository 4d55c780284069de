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


def shared_blocked(x):
    """Channel-blocked copy of a stage input: the one `prepare_shared` made for all candidates of a supernet step (their
    gradients then meet in ONE from_blocked on the way back), else a fresh one."""
    d = getattr(x, '_risp_shared', None)
    return d['blk'] if d is not None and d['version'] == x._version else conv.to_blocked(x)


def shared_stats3(x):
    """(N, 9) per-image [min | mean | max] of a BGR stage input, shared by the eight SRCNNRes proxies of a supernet step."""
    d = getattr(x, '_risp_shared', None)
    return d['st3'] if d is not None and d['version'] == x._version else ops.plane_stats3(x)


def prepare_shared(x):
    """Called by the supernet on the main stream before the candidates of a step fan out over side streams; the cache is
    only valid inside that step's autograd graph, so `release_shared` drops it once the candidates have run."""
    x._risp_shared = {'version': x._version, 'blk': conv.to_blocked(x), 'st3': ops.plane_stats3(x)}


def release_shared(x):
    if hasattr(x, '_risp_shared'):
        del x._risp_shared


def _frozen(*convs):
    for c in convs:
        if c.weight.requires_grad or (c.bias is not None and c.bias.requires_grad):
            return False
    return True


def _derived(module, name, convs, make):
    """Weight-derived constants (folded tables, flipped copies), cached on the module and rebuilt when a weight changes."""
    key = tuple((c.weight._version, c.weight.data_ptr(), None if c.bias is None else c.bias._version) for c in convs)
    cache = module.__dict__.setdefault('_risp_derived', {})
    if name not in cache or cache[name][0] != key:
        with torch.no_grad():
            cache[name] = (key, make())
        if torch.cuda.is_available():
            torch.cuda.current_stream().synchronize()     # built once, then read from whichever stream the candidate runs on
    return cache[name][1]


class SRCNNRes(nn.Module):
    """srcnn_res_arch.py:6-53.  Input = x ++ [per-image min, mean, max of every channel] ++ params,
    broadcast over the frame; conv9-ReLU-conv5-ReLU-conv5; global residual.

    The 9+P broadcast channels are never materialised when the network is frozen (the search path): a channel that is
    constant over the frame contributes  value * (sum of the taps that fall inside the frame)  to the first convolution,
    i.e. a per-image bias that depends only on which of the 9 x 9 border classes the output pixel is in.  That table is
    (N, 81, 64) = features (N, 9+P) x S (9+P, 81*64) with S precomputed from the weights; the convolution itself runs on
    the 3 image channels (1 chunk of 8 instead of 2-3), and the gradient of the features is the class-sums of the masked
    upstream gradient times S^T (`ops.blocked_class_sums`).  The statistics and the blocked copy of x are computed once per
    stage input and shared by the eight proxies."""

    def __init__(self, param_channel):
        super().__init__()
        self.srcnn = nn.Sequential(nn.Conv2d(3 + 9 + param_channel, 64, 9, 1, 4), nn.ReLU(),
                                   nn.Conv2d(64, 32, 5, 1, 2), nn.ReLU(), nn.Conv2d(32, 3, 5, 1, 2))

    def _folded(self):
        c1 = self.srcnn[0]

        def make():
            W = c1.weight.detach()
            K, PAD = W.shape[-1], W.shape[-1] // 2
            cls = torch.arange(K, device=W.device).view(K, 1)
            tap = torch.arange(K, device=W.device).view(1, K)
            # taps of an output pixel in border class a that land inside the frame (class PAD = interior: all of them)
            M = torch.where(cls < PAD, tap >= PAD - cls, torch.where(cls == PAD, torch.ones_like(tap, dtype=torch.bool), tap <= K - cls - 1 + PAD))
            M = M.to(torch.float64)
            S = torch.einsum('ocyx,ay,bx->cabo', W[:, 3:].double(), M, M).reshape(W.shape[1] - 3, K * K * W.shape[0])
            return (W[:, :3].contiguous(), S.float().contiguous(), c1.bias.detach().float().repeat(K * K).contiguous())
        return _derived(self, 'folded', [c1], make)

    def forward(self, x, param_vec):
        x = x.float()
        N, _, H, W = x.shape
        c1, c2, c3 = self.srcnn[0], self.srcnn[2], self.srcnn[4]
        if not _frozen(c1) or min(H, W) < 8:
            return self._forward_materialised(x, param_vec)     # fine-tuning needs the weight gradients of all 12+P channels
        st = shared_stats3(x)
        try:
            feat = torch.cat([st, param_vec], dim=1)
        except Exception:
            raise ValueError(st.size(), None if param_vec is None else param_vec.size())
        w_img, S, b_rep = self._folded()
        tab = ops.bias_table(feat, S, b_rep)                     # (N, 81*64) bias of every border class
        xb = shared_blocked(x)
        h = conv.conv2d_blocked(xb, w_img, None, relu_out=True, bias_tab=tab)
        h = conv.conv2d_blocked(h, c2.weight, c2.bias, relu_out=True)
        h = conv.conv2d_blocked(h, c3.weight, c3.bias, residual=xb)
        return conv.from_blocked(h, 3)

    def _forward_materialised(self, x, param_vec):
        N, _, H, W = x.shape
        st = ops.plane_stats3(x)                                # [min | mean | max], gradient of min / max to the first extremum
        try:
            feat = torch.cat([st, param_vec], dim=1)
        except Exception:
            raise ValueError(st.size(), None if param_vec is None else param_vec.size())
        feat_in = torch.cat([x, feat.view(N, -1, 1, 1).expand(N, feat.shape[1], H, W)], dim=1)
        c1, c2, c3 = self.srcnn[0], self.srcnn[2], self.srcnn[4]
        h = conv.conv2d_blocked(conv.to_blocked(feat_in), c1.weight, c1.bias, relu_out=True)
        h = conv.conv2d_blocked(h, c2.weight, c2.bias, relu_out=True)
        h = conv.conv2d_blocked(h, c3.weight, c3.bias, residual=conv.to_blocked(x))
        return conv.from_blocked(h, 3)


class SRCNNResBank:
    """The SRCNNRes proxies of ONE supernet step evaluated together (`ops.srcnn_res_bank`): their prepared weights, folded
    tables and biases stacked slot by slot, rebuilt when any weight changes.  Only for frozen members (the search path);
    `usable()` says whether the grouped path applies to an input."""

    def __init__(self, nets):
        self.nets = list(nets)                       # slot order
        self._cache = None
        self._pidx = {}

    def usable(self, x):
        return (all(_frozen(n.srcnn[0], n.srcnn[2], n.srcnn[4]) for n in self.nets) and min(x.shape[2], x.shape[3]) >= 8
                and len(self.nets) <= 8 and all(n.srcnn[0].weight.shape[0] == 64 for n in self.nets))

    def bank(self):
        key = tuple((c.weight._version, c.weight.data_ptr(), c.bias._version) for n in self.nets for c in (n.srcnn[0], n.srcnn[2], n.srcnn[4]))
        if self._cache is None or self._cache[0] != key:
            with torch.no_grad():
                Ps = [n.srcnn[0].weight.shape[1] - 12 for n in self.nets]
                Pmax = max(Ps)
                F = 9 + Pmax
                W1, W1T, W2, W2T, W3, W3T, S, BR, B2, B3 = ([] for _ in range(10))
                for n in self.nets:
                    c1, c2, c3 = n.srcnn[0], n.srcnn[2], n.srcnn[4]
                    w_img, S_n, b_rep = n._folded()
                    W1.append(ops._tc_weights(w_img, False)); W1T.append(ops._tc_weights(w_img, True))
                    W2.append(ops._tc_weights(c2.weight, False)); W2T.append(ops._tc_weights(c2.weight, True))
                    W3.append(ops._tc_weights(c3.weight, False)); W3T.append(ops._tc_weights(c3.weight, True))
                    Sp = torch.zeros((F, S_n.shape[1]), device=S_n.device, dtype=torch.float32)
                    Sp[:S_n.shape[0]] = S_n
                    S.append(Sp); BR.append(b_rep)
                    B2.append(c2.bias.detach().float()); B3.append(c3.bias.detach().float())
                st = lambda l: torch.stack(l).contiguous()
                d = dict(W1=st(W1), W1T=st(W1T), W2=st(W2), W2T=st(W2T), W3=st(W3), W3T=st(W3T), S=st(S), BR=st(BR), B2=st(B2), B3=st(B3),
                         Ps=Ps, Pmax=Pmax, F=F, J=int(S[0].shape[1]), par_index=self._par_index)
            if torch.cuda.is_available():
                torch.cuda.current_stream().synchronize()
            self._cache = (key, d)
            self._pidx = {}
        return self._cache[1]

    def _par_index(self, active, device):
        k = (active, str(device))
        if k not in self._pidx:
            d = self._cache[1]
            idx = [g * d['Pmax'] + j for g, slot in enumerate(active) for j in range(d['Ps'][slot])]
            self._pidx[k] = torch.tensor(idx, dtype=torch.int64, device=device)
        return self._pidx[k]

    def __call__(self, x, pars, active):
        """x (N,3,H,W); pars: list of (1, P_slot) tensors in [0,1] for the active slots -> list of G outputs (N,3,H,W)."""
        y = ops.srcnn_res_bank(x, torch.cat(pars, dim=1), self.bank(), list(active))
        N = x.shape[0]
        return list(y.view(len(active), N, *y.shape[1:]).unbind(0))


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
        if _frozen(c1, c2) and c1.weight.shape[0] == c1.weight.shape[1] == c2.weight.shape[0]:
            return ops.resblock_tc(x, c1.weight, c1.bias, c2.weight, c2.bias)     # both convs + a two-launch backward
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
        seq = self.path_restore_14l
        x = x.float()
        if param_vec is None and _frozen(seq[0], seq[3]):
            # frozen network (search path): the BGR<->RGB flips (:65, :84) are folded into the first layer's input channels
            # and the last layer's output channels, and the blocked copy of x is the one the other candidates share
            def make():
                return (seq[0].weight.detach().flip(1).contiguous(), seq[3].weight.detach().flip(0).contiguous(),
                        seq[3].bias.detach().flip(0).contiguous())
            w_first, w_last, b_last = _derived(self, 'flipped', [seq[0], seq[3]], make)
            hb = conv.conv2d_blocked(shared_blocked(x), w_first, seq[0].bias)
            for blk in seq[1]:
                hb = blk(hb)
            return conv.from_blocked(conv.conv2d_blocked(hb, w_last, b_last, relu_in=True), 3)
        h = x.flip(1)
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
