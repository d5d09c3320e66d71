"""Device-side replacement for the arithmetic of the reference's synthetic denoising dataset
(datasets/DenoisingDatasets.py:180-253, SimulateTrain): the CPU loader keeps decoding and cropping uint8 patches
(cv2.imread + crop_patch, :220-226 and datasets/__init__.py:29-39); everything after the crop — float conversion,
sigma map, noise, clipping, flip/rotate augmentation, variance map — is one CUDA kernel over the whole batch
(vk_synth_denoise), so a GPU that consumes >3000 patches/s is not starved by one CPU worker per GPU.

The per-sample random scalars are drawn with Python's `random` in the reference's order (centre h, centre w, scale,
up, down for 'niid' / one level for 'iid'; then the augmentation flag), so a seeded run produces the same maps."""
from __future__ import annotations

import random

import torch

from .. import ops


class SimulateTrainGPU:
    def __init__(self, pch_size=128, chn=3, mode="niid", clip=False):
        self.pch_size, self.chn, self.mode, self.clip = pch_size, chn, mode.lower(), clip
        self.sigma_min, self.sigma_max = 0, 75
        if self.mode not in ("niid", "iid"):
            raise ValueError("Plsase Input corrected noise type: iid or niid")

    def draw_sigma_params(self):
        """One sample's sigma-map scalars, consuming `random` exactly like generate_sigma_niid / generate_sigma_iid."""
        p = self.pch_size
        if self.mode == "niid":
            center = [random.uniform(0, p), random.uniform(0, p)]
            scale = random.uniform(p / 4, p / 4 * 3)
            up = random.uniform(self.sigma_min / 255.0, self.sigma_max / 255.0)
            down = random.uniform(self.sigma_min / 255.0, self.sigma_max / 255.0)
            if up < down:
                up, down = down, up
            up += 5 / 255.0
            return [center[0], center[1], scale, up, down, 0.0]
        level = random.uniform(self.sigma_min / 255.0, self.sigma_max / 255.0)
        return [0.0, 0.0, -1.0, 0.0, 0.0, level]

    def synthesize(self, patches_u8, params=None, aug=None, noise=None):
        """patches_u8: uint8 [N, P, P, C] RGB crops on the device.  Returns (im_noisy, im_gt, sigma_map_gt) NCHW fp32.
        params / aug / noise override the internal draws (parity tests; per-sample draw interleaving is the caller's)."""
        if not patches_u8.is_cuda:
            raise RuntimeError("virnet_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        n = patches_u8.shape[0]
        dev = patches_u8.device
        if params is None or aug is None:
            rows, flags = [], []
            for _ in range(n):
                rows.append(self.draw_sigma_params())
                flags.append(random.randint(0, 7))
            params = torch.tensor(rows, dtype=torch.float64) if params is None else params
            aug = torch.tensor(flags, dtype=torch.int32) if aug is None else aug
        if noise is None:
            noise = torch.randn(patches_u8.shape, device=dev, dtype=torch.float32)
        return ops.synth_denoise(patches_u8.contiguous(), params.to(dev, torch.float64).contiguous(),
                                 aug.to(dev, torch.int32).contiguous(), noise.to(dev).contiguous(), clip=self.clip)
