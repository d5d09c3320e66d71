"""Drop-in for the evaluation helpers of the reference's utils/util_image.py that sit on the callers' hot loop
(scripts/denoising_virnet_*.py, scripts/sisr_virnet_syn.py, the trainers' validation phase), on the device:

* `img_as_ubyte` / `calculate_psnr` / `calculate_ssim` / `batch_PSNR` / `batch_SSIM` (utils/util_image.py:16-116),
* the 8-fold flip / rotate self-ensemble built on `data_aug_np` / `inverse_data_aug_np` (:391-466) as the SIDD / DND
  scripts use it (scripts/denoising_virnet_real_sidd.py:120-136, dnd_submission_py/pytorch_wrapper.py:17-32): the eight
  augmented copies go through the network as ONE batch (two when the image is not square).

PSNR is bit-exact with the reference on uint8 inputs (integer sum of squared differences; the Y channel follows
MATLAB's rgb2ycbcr with round-half-even in fp64); SSIM agrees to fp64 rounding (the reference's cv2.filter2D sums the
121 window taps in another order)."""
from __future__ import annotations

import ctypes as C
import math

import torch

from .. import lib as _l
from ..ops import _ptr, _stream

_WIN = None


def _ssim_window():
    """np.outer(cv2.getGaussianKernel(11, 1.5), ...) (utils/util_image.py:22-23): exp(-(i-5)^2 / (2 * 1.5^2)), normalised."""
    global _WIN
    if _WIN is None:
        ax = torch.arange(11, dtype=torch.float64) - 5
        k1 = torch.exp(-(ax ** 2) / (2 * 1.5 ** 2))
        k1 = k1 / k1.sum()
        _WIN = (C.c_double * 121)(*torch.outer(k1, k1).flatten().tolist())
    return _WIN


def _need_cuda(t):
    if not t.is_cuda:
        raise RuntimeError("virnet_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")


def img_as_ubyte(x: torch.Tensor) -> torch.Tensor:
    """NCHW (or CHW) float image in [0, 1] -> [N, H, W, C] (or [H, W, C]) uint8, like skimage.img_as_ubyte applied to
    `x.clamp(0, 1)` transposed to HWC (scripts/sisr_virnet_syn.py:147, denoising_virnet_syn.py:137)."""
    _need_cuda(x)
    squeeze = x.dim() == 3
    x4 = (x.unsqueeze(0) if squeeze else x).contiguous().float()
    n, c, h, w = x4.shape
    out = torch.empty(n, h, w, c, device=x.device, dtype=torch.uint8)
    _l.check(_l.load().vk_to_u8(_ptr(x4), _ptr(out), n, c, h, w, _stream()), "vk_to_u8")
    return out[0] if squeeze else out


def _check_pair(im1, im2):
    if im1.shape != im2.shape:
        raise ValueError("Input images must have the same dimensions.")
    _need_cuda(im1)
    if im1.dtype != torch.uint8 or im2.dtype != torch.uint8:
        raise TypeError("uint8 images expected (use img_as_ubyte)")
    a, b = im1.contiguous(), im2.contiguous()
    if a.dim() == 2:
        a, b = a.unsqueeze(-1), b.unsqueeze(-1)
    return a, b


def calculate_psnr(im1, im2, border=0, ycbcr=False) -> float:
    """utils/util_image.py:68-89 on device uint8 [H, W, C] (or [H, W]) images; returns a Python float."""
    a, b = _check_pair(im1, im2)
    h, w, c = a.shape
    ssd = torch.zeros(1, device=a.device, dtype=torch.int64)
    _l.check(_l.load().vk_psnr_u8(_ptr(a), _ptr(b), h, w, c, int(border), int(bool(ycbcr)), _ptr(ssd), _stream()),
             "vk_psnr_u8")
    count = (h - 2 * border) * (w - 2 * border) * (1 if ycbcr else c)
    mse = float(ssd.item()) / count
    if mse == 0:
        return float("inf")
    return 20 * math.log10(255.0 / math.sqrt(mse))


def calculate_ssim(im1, im2, border=0, ycbcr=False) -> float:
    """utils/util_image.py:39-66 on device uint8 images: mean of the per-channel SSIM maps."""
    a, b = _check_pair(im1, im2)
    h, w, c = a.shape
    nch = 1 if ycbcr else c
    sums = torch.zeros(nch, device=a.device, dtype=torch.float64)
    _l.check(_l.load().vk_ssim_u8(_ptr(a), _ptr(b), h, w, c, int(border), int(bool(ycbcr)), _ssim_window(), _ptr(sums),
                                  _stream()), "vk_ssim_u8")
    npix = (h - 2 * border - 10) * (w - 2 * border - 10)
    per_channel = (sums / npix).tolist()
    return sum(per_channel) / nch


def batch_PSNR(img, imclean, border=0, ycbcr=False) -> float:
    """utils/util_image.py:91-103: mean PSNR over a batch of NCHW float images in [0, 1]."""
    a, b = img_as_ubyte(img), img_as_ubyte(imclean)
    return sum(calculate_psnr(b[i], a[i], border, ycbcr) for i in range(a.shape[0])) / a.shape[0]


def batch_SSIM(img, imclean, border=0, ycbcr=False) -> float:
    """utils/util_image.py:105-116."""
    a, b = img_as_ubyte(img), img_as_ubyte(imclean)
    return sum(calculate_ssim(b[i], a[i], border, ycbcr) for i in range(a.shape[0])) / a.shape[0]


def self_ensemble(net, x: torch.Tensor, clip: bool = False) -> torch.Tensor:
    """mean over the 8 flips / rotations m of inverse_data_aug(net(data_aug(x, m))[0], m) for NCHW fp32 `x`:
    one gather kernel, ONE batched forward of 8N images (4N + 4N when H != W), one averaging kernel."""
    _need_cuda(x)
    x = x.contiguous().float()
    n, c, h, w = x.shape
    lib = _l.load()
    buf = torch.empty(8 * n * c * h * w, device=x.device, dtype=torch.float32)
    half = 4 * n * c * h * w
    _l.check(lib.vk_aug8(_ptr(x), _ptr(buf), buf.data_ptr() + 4 * half, n * c, h, w, _stream()), "vk_aug8")
    with torch.no_grad():
        if h == w:
            mu = net(buf.view(8 * n, c, h, w))[0].contiguous()
            mu_a, mu_b = mu[:4 * n], mu[4 * n:]
        else:
            mu_a = net(buf[:half].view(4 * n, c, h, w))[0].contiguous()
            mu_b = net(buf[half:].view(4 * n, c, w, h))[0].contiguous()
    out = torch.empty_like(x)
    _l.check(lib.vk_aug8_merge(_ptr(mu_a), _ptr(mu_b), _ptr(out), n * c, h, w, int(clip), _stream()), "vk_aug8_merge")
    return out
