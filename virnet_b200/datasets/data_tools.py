"""Drop-in for the hot-path part of the reference's datasets/data_tools.py: `MixUp_AUG`
(datasets/data_tools.py:12-30), applied to every batch by train_denoising_real.py:163.

The random draws are made exactly like the reference's (CPU `torch.randperm`, then a CPU Beta(0.6, 0.6)
`rsample((bs, 1))`), so a seeded run mixes the same pairs with the same coefficients; the blend of both
tensors is one CUDA kernel (vk_mixup)."""
from __future__ import annotations

import torch

from .. import ops


class MixUp_AUG:
    def __init__(self):
        self.dist = torch.distributions.beta.Beta(torch.tensor([0.6]), torch.tensor([0.6]))

    def draw(self, bs, device):
        indices = torch.randperm(bs)
        lam = self.dist.rsample((bs, 1)).view(-1)
        return indices.to(device, non_blocking=True), lam.to(device, non_blocking=True)

    def aug(self, rgb_gt, rgb_noisy, draws=None):
        if not rgb_gt.is_cuda:
            raise RuntimeError("virnet_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        indices, lam = draws if draws is not None else self.draw(rgb_gt.size(0), rgb_gt.device)
        return ops.mixup(rgb_gt.contiguous().float(), rgb_noisy.contiguous().float(), indices, lam.contiguous().float())
