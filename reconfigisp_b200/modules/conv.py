"""Dense convolution entry used by the CNN candidates (NCHW fp32, stride 1, 'same' zero padding)
with the fusions the candidate architectures need: bias, ReLU on the input (the leading in-place
ReLU of `ResidualBlock`, path_14l_bayer_arch.py:9-21), ReLU on the output, residual add (optionally
of relu(residual)).

ROUND-1 STATUS: this routes to the library convolution of torch (im2col + cuBLAS SGEMM; library
code, like a cuBLAS call) -- the hand-written sm_100a implicit-GEMM kernel is the next item of the
build plan (DESIGN.md "CNN candidates").  cuDNN is switched OFF here on purpose: on B200 its fp32
5x5 engines return ~3e-3 max-abs error against an fp32 CPU convolution even with allow_tf32=False
(measured, scripts/diag_conv2.py), which breaks the 1e-4 parity bar; the non-cuDNN path is exact to
~2e-6.  Everything around the convolution (pack / pixel-shuffle, statistics, mixed-op, loss) already
runs on this package's own kernels.
"""
import torch
import torch.nn.functional as F


def conv2d(x, weight, bias=None, relu_in=False, relu_out=False, residual=None, residual_relu=False):
    if not x.is_cuda:
        raise RuntimeError('reconfigisp_b200 CNN candidates run on CUDA tensors only (no CPU fallback)')
    k = weight.shape[-1]
    h = torch.relu(x) if relu_in else x
    with torch.backends.cudnn.flags(enabled=False):
        y = F.conv2d(h, weight, bias, stride=1, padding=k // 2)
    if relu_out:
        y = torch.relu(y)
    if residual is not None:
        y = y + (torch.relu(residual) if residual_relu else residual)
    return y
