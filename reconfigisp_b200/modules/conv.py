"""Dense convolution entry used by the CNN candidates (NCHW fp32, stride 1, 'same' zero padding) with the
fusions the candidate architectures need: bias, ReLU on the input (the leading in-place ReLU of
`ResidualBlock`, path_14l_bayer_arch.py:9-21), ReLU on the output, residual add (optionally of
relu(residual)).  Runs on this package's own sm_100a kernel (`csrc/risp_conv.cu`, exact fp32 accumulation).

Why not cuDNN: on B200 its fp32 5x5 engines return ~3e-3 max-abs error against an fp32 CPU convolution even
with allow_tf32=False (measured, scripts/diag_conv2.py), which breaks the 1e-4 parity bar of this path.
"""
from .. import ops


def conv2d(x, weight, bias=None, relu_in=False, relu_out=False, residual=None, residual_relu=False):
    if not x.is_cuda:
        raise RuntimeError('reconfigisp_b200 CNN candidates run on CUDA tensors only (no CPU fallback)')
    return ops.conv2d(x, weight, bias, relu_in, relu_out, residual, residual_relu)


# ---- tensor-core path on channel-blocked activations (csrc/risp_conv_tc.cu) -------------------------------
def to_blocked(x):
    return ops.to_blocked(x)


def from_blocked(xb, channels):
    return ops.from_blocked(xb, channels)


def conv2d_blocked(xb, weight, bias=None, relu_in=False, relu_out=False, residual=None, residual_relu=False, bias_tab=None):
    """Same contract as `conv2d`, but input / residual / output are channel-blocked (N,H,C16/4,W,4) tensors and
    the arithmetic runs on the tcgen05 tensor cores (3-term TF32 split, fp32-accurate)."""
    return ops.conv2d_tc(xb, weight, bias, relu_in, relu_out, residual, residual_relu, bias_tab)
