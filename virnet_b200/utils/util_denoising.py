"""Drop-in for the hot-path part of the reference's utils/util_denoising.py: `noise_estimate_fun`
(utils/util_denoising.py:54-63), the variance-map prior of the real-noise trainer
(train_denoising_real.py:164), as one CUDA kernel (vk_noise_estimate).

The Gaussian window is the reference's `inverse_gamma_kernel` (:24-35): OpenCV's getGaussianKernel formula
exp(-(i - (k-1)/2)^2 / (2 s^2)) normalised to sum 1, s = 0.3 ((k-1)/2 - 1) + 0.8, outer product, renormalised,
computed in float64 on the host and stored as fp32 — no OpenCV dependency."""
from __future__ import annotations

import math

import torch

from .. import ops

_WINDOWS = {}


def gaussian_window(k_size: int, device) -> torch.Tensor:
    key = (int(k_size), str(device))
    win = _WINDOWS.get(key)
    if win is None:
        scale = 0.3 * ((k_size - 1) * 0.5 - 1) + 0.8
        ax = torch.arange(k_size, dtype=torch.float64) - (k_size - 1) / 2
        k1 = torch.exp(-(ax ** 2) / (2 * scale ** 2))
        k1 = k1 / k1.sum()
        k2 = torch.outer(k1, k1)
        win = (k2 / k2.sum()).to(torch.float32).to(device).contiguous()
        _WINDOWS[key] = win
    return win


def noise_estimate_fun(im_noisy, im_gt, k_size):
    """Estimate the variance map: N x c x h x w -> N x c x h x w, clamped at 1e-10."""
    if not im_noisy.is_cuda:
        raise RuntimeError("virnet_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
    return ops.noise_estimate(im_noisy.contiguous().float(), im_gt.contiguous().float(),
                              gaussian_window(int(k_size), im_noisy.device))
