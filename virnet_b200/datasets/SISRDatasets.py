"""Device-side replacement for the arithmetic of the reference's SISR training dataset
(datasets/SISRDatasets.py:66-122, GeneralTrainFloder.__getitem__, Gaussian-noise branch): the CPU loader keeps
decoding, cropping and flipping HR patches; the anisotropic Gaussian blur, the clip, the Direct / bicubic
down-sampling and the noise are one C-ABI call over the whole batch (vk_sisr_degrade).  JPEG noise
(util_image.jpeg_compress, an OpenCV codec round trip) stays on the CPU path of the reference.

Per-sample scalars (lam1, lam2, theta, noise level) are drawn with Python's `random` in the reference's order."""
from __future__ import annotations

import math
import random

import numpy as np
import torch

from .. import ops
from ..loss.resize_right import downsample_matrix


def shifted_anisotropic_Gaussian(k_size=21, sf=4, lambda_1=1.2, lambda_2=5.0, theta=0.0, shift=True):
    """utils/util_sisr.py:60-93: k x k softmax-normalised Gaussian with covariance U diag(l1, l2) U^T and the
    (s1, s2, rho) description of its marginals.  Host-side numpy (441 values per sample)."""
    c, s = np.cos(theta), np.sin(theta)
    rot = np.array([[c, -s], [s, c]])
    cov = rot @ np.diag([lambda_1, lambda_2]) @ rot.T
    inv = np.linalg.inv(cov)
    center = k_size // 2 + 0.5 * (sf - k_size % 2) if shift else k_size // 2
    xs, ys = np.meshgrid(range(k_size), range(k_size))
    z = np.stack([xs, ys], 2).astype(np.float32) - center                    # k x k x 2, float32 like the reference
    q = -0.5 * np.einsum("ija,ab,ijb->ij", z, inv, z).reshape(-1)
    e = np.exp(q - q.max())
    kernel = (e / e.sum()).reshape(k_size, k_size)
    s1, s2 = cov[0, 0], cov[1, 1]
    rho = cov[0, 1] / (math.sqrt(s1) * math.sqrt(s2))
    return kernel, np.array([s1, s2, rho])


class GeneralTrainGPU:
    def __init__(self, sf, k_size=21, kernel_shift=False, downsampler="Bicubic", noise_level=(0.1, 15)):
        self.sf, self.k_size, self.kernel_shift = int(sf), int(k_size), bool(kernel_shift)
        self.downsampler, self.noise_level = downsampler, noise_level

    def draw(self):
        """One sample's (kernel, kernel_infos, std), consuming `random` like SISRDatasets.py:78-99 (Gaussian noise)."""
        lam1 = random.uniform(0.2, self.sf)
        lam2 = random.uniform(lam1, self.sf) if random.random() < 0.7 else lam1
        theta = random.uniform(0, np.pi)
        kernel, infos = shifted_anisotropic_Gaussian(self.k_size, self.sf, lam1 ** 2, lam2 ** 2, theta, self.kernel_shift)
        std = random.uniform(self.noise_level[0], self.noise_level[1]) / 255.0
        return kernel, infos, std

    def degrade(self, im_hr, kernels=None, std=None, noise=None):
        """im_hr [N, C, H, W] fp32 on the device (cropped, augmented).  Returns (im_hr, im_lr, im_blur, kernel_infos
        [N, 3], nlevel [N, 1, 1, 1]) like the reference's batch; kernels / std / noise override the internal draws."""
        if not im_hr.is_cuda:
            raise RuntimeError("virnet_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        n, c, H, W = im_hr.shape
        dev = im_hr.device
        infos = None
        if kernels is None or std is None:
            ks, inf, sd = zip(*[self.draw() for _ in range(n)])
            kernels = torch.from_numpy(np.stack(ks)).float() if kernels is None else kernels
            std = torch.tensor(sd, dtype=torch.float32) if std is None else std
            infos = torch.from_numpy(np.stack(inf)).float()
        rh = downsample_matrix(H, self.sf, self.downsampler, dev)
        rw = downsample_matrix(W, self.sf, self.downsampler, dev)
        if noise is None:
            noise = torch.randn(n, c, rh.shape[0], rw.shape[0], device=dev, dtype=torch.float32)
        std = std.to(dev, torch.float32).reshape(n).contiguous()
        im_blur, im_lr = ops.sisr_degrade(im_hr.contiguous().float(), kernels.to(dev, torch.float32).contiguous(), rh, rw,
                                          noise.to(dev).contiguous(), std)
        return im_hr, im_lr, im_blur, (None if infos is None else infos.to(dev)), std.view(n, 1, 1, 1)
