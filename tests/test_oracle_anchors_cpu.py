"""Secondary anchors for the oracle ops whose arithmetic lives in the un-shipped /DATA/ISP_Kernels ("parity unpinned",
oracle/SPEC.md): OpenCV implements the same published algorithms, so the oracle is held to it where it applies.
CPU only; skipped when cv2 is not importable."""
import numpy as np
import pytest
import torch

cv2 = pytest.importorskip('cv2')

from oracle import isp_oracle as O   # noqa: E402


def _hwc(t):
    return np.ascontiguousarray(t.permute(1, 2, 0).numpy())


def test_median_is_cv2_median_blur():
    x = torch.rand(2, 3, 41, 53, generator=torch.Generator().manual_seed(1)) * 255
    for k in (3, 5):                                           # cv2 supports float32 for k = 3, 5
        ref = torch.stack([torch.stack([torch.from_numpy(cv2.medianBlur(x[n, c].numpy().copy(), k)) for c in range(3)]) for n in range(2)])
        assert torch.equal(O.denoise_median(x, k), ref), k


def test_bilateral_is_cv2_bilateral_filter():
    x = torch.rand(2, 3, 41, 53, generator=torch.Generator().manual_seed(1)) * 255
    win, sc, ss = torch.tensor([5, 9]), [25., 60.], [3., 10.]
    mine = O.denoise_bilateral(x, win, sc, ss)
    for n in range(2):
        ref = cv2.bilateralFilter(_hwc(x[n]), int(win[n]), sc[n], ss[n])       # BORDER_REFLECT_101, L1 colour distance
        # OpenCV's float path interpolates exp() from a table: a few 1e-4 on 0-255 data
        assert float(np.abs(_hwc(mine[n]) - ref).max()) <= 5e-4, n


def test_bilinear_demosaic_phase_and_plane_order_match_cv2():
    """R at (0,0) / BGR planes / bilinear kernels: OpenCV's COLOR_BayerBG2BGR is this CFA (its names refer to the second
    row); it rounds to integers, so interior pixels agree within half a code.  The other three phases are far off."""
    img = (torch.rand(1, 3, 40, 52, generator=torch.Generator().manual_seed(2)) * 255).round()
    raw = O.mosaic_bgr(img)
    mine = _hwc(O.demosaic_bilinear(raw)[0])
    r8 = raw[0, 0].numpy().astype(np.uint8)
    err = {}
    for name in ('COLOR_BayerBG2BGR', 'COLOR_BayerGB2BGR', 'COLOR_BayerRG2BGR', 'COLOR_BayerGR2BGR'):
        ref = cv2.cvtColor(r8, getattr(cv2, name)).astype(np.float32)
        err[name] = float(np.abs(mine - ref)[2:-2, 2:-2].max())
    assert err['COLOR_BayerBG2BGR'] <= 0.5, err
    assert min(v for k, v in err.items() if k != 'COLOR_BayerBG2BGR') > 50, err
    # nearest-neighbour keeps the sampled values of every plane untouched (pure data movement)
    near = O.demosaic_nearest(raw)
    assert torch.equal(near[0, 2, 0::2, 0::2], raw[0, 0, 0::2, 0::2]) and torch.equal(near[0, 0, 1::2, 1::2], raw[0, 0, 1::2, 1::2])


def test_malvar_reproduces_linear_ramps_and_constants():
    """Malvar-He-Cutler is exact on images whose planes are affine in (y, x) with a common gradient: the Laplacian
    correction terms vanish.  (No OpenCV counterpart exists; this pins the published kernels' normalisation.)"""
    yy, xx = torch.meshgrid(torch.arange(24.), torch.arange(32.), indexing='ij')
    ramp = 0.2 + 0.01 * yy + 0.005 * xx
    img = torch.stack([ramp + 0.1, ramp, ramp - 0.05]).unsqueeze(0)       # B, G, R: same gradient, different offsets
    out = O.demosaic_laplacian(O.mosaic_bgr(img), 1.0)
    assert float((out - img)[:, :, 2:-2, 2:-2].abs().max()) <= 1e-6
    const = torch.full((1, 3, 16, 16), 0.4)
    assert float((O.demosaic_laplacian(O.mosaic_bgr(const), 1.0) - const).abs().max()) <= 1e-6


def test_fastnlm_is_opencv_fast_nl_means():
    """The wrapper's option name and parameters (`fastnlm`: block_size, search_block, decay_factor, tools_origin.py:762-804)
    are OpenCV's.  The oracle's definition reproduces `cv2.fastNlMeansDenoising` on a 3-channel 8-bit image (patch distance
    averaged jointly over the channels, weights exp(-d2/h^2)) to within OpenCV's own 8-bit output rounding and fixed-point
    weight table: <= 1 code max, <= 0.3 code mean, for several (block, search, h).  The Lab-space
    `fastNlMeansDenoisingColored` is a different operator (it denoises L and ab separately): measured 1-3.5 codes mean
    away on the same inputs, recorded here so the distance is on file (oracle/SPEC.md)."""
    cv2 = pytest.importorskip('cv2')
    rng = np.random.RandomState(3)
    H, W = 48, 56
    yy, xx = np.mgrid[0:H, 0:W]
    clean = np.stack([120 + 60 * np.sin(xx / 9.0), 100 + 50 * np.cos(yy / 7.0), 140 + 40 * np.sin((xx + yy) / 11.0)], -1)
    u8 = np.clip(clean + rng.randn(H, W, 3) * 8, 0, 255).round().astype(np.uint8)
    x = torch.from_numpy(u8.astype(np.float32)).permute(2, 0, 1)[None]
    for b, s, h in ((3, 3, 10.0), (3, 7, 10.0), (5, 11, 20.0)):
        ours = O.denoise_fastnlm(x, [b], [s], [h])[0].permute(1, 2, 0).numpy()
        cvg = cv2.fastNlMeansDenoising(u8, None, h, b, s).astype(np.float32)
        d = np.abs(ours - cvg)
        assert d.max() <= 1.0 and d.mean() <= 0.3, (b, s, h, d.max(), d.mean())
        lab = cv2.fastNlMeansDenoisingColored(u8, None, h, h, b, s).astype(np.float32)
        assert 0.3 < np.abs(ours - lab).mean() < 6.0          # a different operator, by a known margin
