"""Synthetic RGGB raws + BGR ground truth for tests and the benchmark (no datasets are available).

Definition: a smooth RGB scene (low-frequency sinusoids + a few rectangles) in [0,1] is
mosaiced to RGGB with the reference's phase (R=(0,0) G1=(0,1) G2=(1,0) B=(1,1), srcnn_demosaic_arch.py:39-42),
shot+read noise is added, the result is clamped and QUANTISED TO 10 BIT and divided by 1023 like the
reference's loaders (`/1023.`, s7isp_rggb2bgr_dataset.py:123); the ground truth is the clean scene in BGR
order quantised to 8 bit (`/255.`, :134).  Seed 10 = the reference's manual_seed / test_seed.
"""
import math

import torch


def _scene(H, W, g):
    yy = torch.linspace(0, 1, H).view(H, 1)
    xx = torch.linspace(0, 1, W).view(1, W)
    chans = []
    for c in range(3):
        f = torch.rand(4, generator=g) * 6 + 1
        ph = torch.rand(4, generator=g) * 2 * math.pi
        s = 0.45 + 0.2 * torch.sin(f[0] * yy + ph[0]) * torch.cos(f[1] * xx + ph[1]) \
            + 0.15 * torch.sin(f[2] * (yy + xx) + ph[2]) + 0.1 * torch.cos(f[3] * (yy - xx) + ph[3])
        chans.append(s)
    img = torch.stack(chans)                                        # (3,H,W) B,G,R
    for _ in range(6):
        y0, x0 = int(torch.randint(0, H - 8, (1,), generator=g)), int(torch.randint(0, W - 8, (1,), generator=g))
        h, w = int(torch.randint(8, max(9, H // 3), (1,), generator=g)), int(torch.randint(8, max(9, W // 3), (1,), generator=g))
        img[:, y0:y0 + h, x0:x0 + w] = torch.rand(3, 1, 1, generator=g) * 0.9 + 0.05
    return img.clamp_(0, 1)


def synthetic_frames(N, H, W, seed=10, pin=False):
    """-> raw (N,1,H,W), gt (N,3,H,W) float32 CPU tensors (optionally pinned)."""
    assert H % 2 == 0 and W % 2 == 0
    g = torch.Generator().manual_seed(seed)
    base = _scene(H, W, g)
    raw = torch.empty((N, 1, H, W), dtype=torch.float32, pin_memory=pin)
    gt = torch.empty((N, 3, H, W), dtype=torch.float32, pin_memory=pin)
    for n in range(N):
        # frames differ by an even (phase-preserving) cyclic shift and a gain, so each is distinct data
        sy, sx = 2 * int(torch.randint(0, H // 2, (1,), generator=g)), 2 * int(torch.randint(0, W // 2, (1,), generator=g))
        scene = torch.roll(base, (sy, sx), dims=(1, 2)) * (0.8 + 0.2 * float(torch.rand(1, generator=g)))
        gt[n] = torch.round(scene * 255) / 255
        m = torch.empty((H, W), dtype=torch.float32)
        m[0::2, 0::2] = scene[2, 0::2, 0::2]
        m[0::2, 1::2] = scene[1, 0::2, 1::2]
        m[1::2, 0::2] = scene[1, 1::2, 0::2]
        m[1::2, 1::2] = scene[0, 1::2, 1::2]
        noise = torch.randn((H, W), generator=g)
        m = m + noise * torch.sqrt(m * 1e-3 + 1e-5)                # shot + read noise
        raw[n, 0] = torch.round(m.clamp_(0, 1) * 1023) / 1023
    return raw, gt
