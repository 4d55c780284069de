"""Shared helpers of the five plugin modules (`Kernel().run(img, option, params)`)."""
import numpy as np
import torch

from reconfigisp_b200 import ops


def nhwc_to_nchw(img):
    """The reference wrappers hand over `img.permute(0, 2, 3, 1)` of an NCHW-contiguous tensor
    (tools_origin.py:33,59,211): undoing the permute is free.  A genuinely NHWC-contiguous tensor
    is re-laid out once (device copy, plumbing)."""
    t = img.permute(0, 3, 1, 2)
    return t if t.is_contiguous() else t.contiguous()


def nchw_to_nhwc(img):
    return img.permute(0, 2, 3, 1)


def dev_vec(v, like):
    """numpy / tensor per-image parameter -> 1-D float tensor on the image's device."""
    if torch.is_tensor(v):
        return v.detach().to(device=like.device, dtype=torch.float32).reshape(-1)
    return torch.as_tensor(np.asarray(v, dtype=np.float32), device=like.device).reshape(-1)
