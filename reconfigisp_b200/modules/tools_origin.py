"""Candidate modules of the ISP search space -- same class names and `forward(img, params)`
signatures as the reference's `codes/models/modules/tools_origin.py`, re-implemented on the
sm_100a kernels (reconfigisp_b200.ops).

Conventions: `img` fp32 NCHW in [0,1], BGR plane order; `params` (N,P) in [0,1]
(sigmoid of the learnt logits, repeated over the batch) or None.  The reference permutes to NHWC,
calls an external kernel and permutes back (e.g. tools_origin.py:59-71); here every module works on
the NCHW planes directly, and the wrappers' [0,1]->range mappings are tiny torch ops on (N,P).
"""
import torch
import torch.nn as nn

from .. import ops


class Skip(nn.Module):
    """tools_origin.py:256-262 -- returns the SAME tensor object."""
    chain_op = ('skip', 0)

    def forward(self, img, params=None):
        return img

    @staticmethod
    def kernel_params(params):
        return None


class Gamma(nn.Module):
    """tools_origin.py:48-73; y = clamp(x, 1e-8, 1) ** gamma, one gamma in [0,1] per image."""
    chain_op = ('gamma', 0)

    def forward(self, img, params):
        return ops.gamma(img, params)

    @staticmethod
    def kernel_params(params):
        return params


class Grayworld(nn.Module):
    """tools_origin.py:22-45; gray-world white balance, output clipped to [0,1], no parameters."""

    def forward(self, img, params=None):
        return ops.grayworld(img)


class WbManual(nn.Module):
    """tools_origin.py:200-225; per-channel gain, [0,1] -> [0,5] (:214)."""
    chain_op = ('gain', 0)

    def forward(self, img, params=None):
        return ops.gain(img, params * 5)

    @staticmethod
    def kernel_params(params):
        return params * 5


class WbQuadratic(nn.Module):
    """tools_origin.py:313-359; 3x10 quadratic colour transform, [0,1] -> [-5,5] (:326), clip [0,1]."""
    chain_op = ('poly10', 0)

    def forward(self, img, params):
        return ops.poly10(img, params * 10 - 5)

    @staticmethod
    def kernel_params(params):
        return params * 10 - 5


class GtmManual(nn.Module):
    """tools_origin.py:409-440; piecewise-linear tone curve with n_seg segments.  Quirk kept: the
    knots of batch element 0 are used for the whole batch (:422)."""

    def __init__(self, n_seg):
        super().__init__()
        self.n_seg = n_seg
        self.chain_op = ('gtm', n_seg)

    def forward(self, imgs, params):
        return ops.gtm(imgs, params[0:1], self.n_seg)

    @staticmethod
    def kernel_params(params):
        return params[0:1]


class DemosaicNearest(nn.Module):
    """tools_origin.py:265-286; (N,1,H,W) RGGB -> (N,3,H,W) BGR, differentiable."""
    demosaic_kind = 'nearest'

    def forward(self, img, params=None):
        return ops.demosaic(img, 'nearest')


class DemosaicNet(nn.Module):
    """tools_origin.py:289-310; external network (weights not shipped): served by the callable
    registered in reconfigisp_b200.isp_kernels.demosaic (default: seeded SRCNNDemosaic stand-in)."""

    def forward(self, img, params=None):
        from ..isp_kernels import demosaic as dm
        return dm.Demosaic().run(img, 'demosaicnet', {'input': {'format': 'RGGB'}})


# ---- conditional modules (sRGB 16-18) ------------------------------------------------------------------
class ConditionalModuleBGR(nn.Module):
    """tools_origin.py:77-163.  Per-image 3 x hist_bin histogram -> FC stack whose weights are slices
    of the flat parameter vector -> + global parameters -> sigmoid.  The reference computes the
    histogram on the CPU with a device->host round trip per channel per image (:124); here it is one
    kernel launch on the device (bit-identical bin rule)."""

    def __init__(self, in_channels, out_channel):
        super().__init__()
        assert in_channels[0] % 3 == 0
        self.hist_bin = in_channels[0] // 3
        self.in_out_channels = [ch for ch in in_channels] + [out_channel]
        self.total_params = sum(i * o + o for i, o in zip(self.in_out_channels[:-1], self.in_out_channels[1:]))
        self.total_params += out_channel
        self.module_params = out_channel

    def _fc_forward(self, img, params):
        assert params.size(0) == self.total_params
        assert img.size(1) == 3
        feat = ops.histc01(img.detach(), self.hist_bin)            # (N, 3*hist_bin), no gradient (:129)
        idx, layers = 0, len(self.in_out_channels) - 1
        for k in range(layers):
            ci, co = self.in_out_channels[k], self.in_out_channels[k + 1]
            weight = params[idx: idx + ci * co].view(ci, co)
            idx += ci * co
            bias = params[idx: idx + co]
            idx += co
            feat = torch.matmul(feat, weight) + bias
            if k != layers - 1:
                feat = torch.relu(feat)
        assert idx == self.total_params - self.module_params
        # quirk kept (:158): a single scalar `params[idx]` is read as the "global" parameter
        return torch.sigmoid(params[idx] + feat)


class ConditionalGamma(ConditionalModuleBGR):
    """tools_origin.py:167-197."""

    def __init__(self, in_channels):
        super().__init__(in_channels, 1)

    def forward(self, img, params):
        return ops.gamma(img, self._fc_forward(img, params))


class ConditionalWbManual(ConditionalModuleBGR):
    """tools_origin.py:229-253."""

    def __init__(self, in_channels):
        super().__init__(in_channels, 3)

    def forward(self, img, params=None):
        return ops.gain(img, self._fc_forward(img, params) * 5)


class ConditionalWbQuadratic(ConditionalModuleBGR):
    """tools_origin.py:363-406."""

    def __init__(self, in_channels):
        super().__init__(in_channels, 30)

    def forward(self, img, params):
        return ops.poly10(img, self._fc_forward(img, params) * 10 - 5)


# ---- original (non-differentiable) algorithms that have proxy nets -----------------------------------------
# The reference scales to [0,255], detaches the parameters through numpy (a host sync) and scales back
# (e.g. tools_origin.py:528-546).  The operators below are scale-free or take the data scale as an
# argument, so they run on the [0,1] planes with parameters that never leave the device.
class OriginDemosBilinear(nn.Module):
    """tools_origin.py:445-475."""
    demosaic_kind = 'bilinear'

    def forward(self, img, params=None):
        return ops.demosaic(img.detach(), 'bilinear')


class OriginDemosLaplacian(nn.Module):
    """tools_origin.py:479-509; Malvar-He-Cutler, clipped to the 8-bit output range."""
    demosaic_kind = 'malvar'

    def forward(self, img, params=None):
        return ops.demosaic(img.detach(), 'malvar', 1.0)


class OriginToneReinhard(nn.Module):
    """tools_origin.py:513-550; params (N,2): white_point, middle_grey."""

    def forward(self, img, params):
        p = params.detach()
        return ops.tone_reinhard(img, p[:, 0], p[:, 1])


class OriginToneCrysis(nn.Module):
    """tools_origin.py:554-588; params (N,1): lum_adapted."""
    chain_op = ('crysis', 0)

    def forward(self, img, params):
        return ops.tone_crysis(img, params.detach()[:, 0])

    @staticmethod
    def kernel_params(params):
        return (1.0 / torch.clamp(params.detach()[:, 0:1], min=1e-3))


class OriginToneFilmic(nn.Module):
    """tools_origin.py:592-630; params (N,2): white_point, exposure_bias [0,1] -> [1,10] (:613)."""
    chain_op = ('filmic', 0)

    def forward(self, img, params):
        p = params.detach()
        return ops.tone_filmic(img, p[:, 0], p[:, 1] * 9. + 1.)

    @staticmethod
    def kernel_params(params):
        p = params.detach()
        w = torch.clamp(p[:, 0], min=1e-3)
        return torch.stack([p[:, 1] * 9. + 1., 1.0 / ops._hable(w)], dim=1)


class OriginWbWhiteworld(nn.Module):
    """tools_origin.py:634-669; params (N,1): white point ratio."""

    def forward(self, img, params):
        return ops.whiteworld(img, params.detach()[:, 0])


class OriginNoiseBilateral(nn.Module):
    """tools_origin.py:673-717.  Quirk kept (:698): `.int()` truncates p in [0,1) to 0, so the window
    is 3 unless p == 1.  sigma_color is given on the 0-255 scale (:699) and applied to [0,1] data."""

    def forward(self, img, params):
        p = params.detach()
        window = (p[:, 0].int() * 7) * 2 + 3
        sigma_color = (p[:, 1] * 99 + 1) / 255.0
        sigma_space = p[:, 2] * 99 + 1
        return ops.bilateral(img, window, sigma_color, sigma_space, max_window=17)


class OriginNoiseMedian(nn.Module):
    """tools_origin.py:721-758; k = 2*int(p[0]*7)+3 for the whole batch (:744).  The window size is a
    launch parameter, so this module reads one scalar back from the device (as the reference does)."""

    def forward(self, img, params):
        k = 2 * int(params.detach()[0, 0].float().mul(7).item()) + 3
        return ops.median(img, k)


class OriginNoiseFastnlm(nn.Module):
    """tools_origin.py:762-804.  Same `.int()*7` quirk as the bilateral window (:786-787): block and search sizes
    are 3 unless p == 1 (then 17); decay h = p*99+1 on the 0-255 scale (:788), applied to [0,1] data here."""

    def forward(self, img, params):
        p = params.detach()
        block = (p[:, 0].int() * 7) * 2 + 3
        search = (p[:, 1].int() * 7) * 2 + 3
        h = (p[:, 2] * 99 + 1) / 255.0
        return ops.fastnlm(img, block, search, h, max_halo=16)
