"""Fused replacement for the reference's loss/ELBO_simple.py:elbo_denoising_simple.

Same call signature and return tuple `(loss, lh, kl_gauss, kl_Igamma)` (reference
loss/ELBO_simple.py:23-53).  One CUDA kernel computes the three means and the gradients
w.r.t. `mu` and `sigma_est` in a single pass over the data; autograd receives them through
a custom Function, so `loss.backward()` works as with the reference.
"""
from __future__ import annotations

import math

import torch

from .. import ops

_DIGAMMA_CACHE = {}


def _digamma(v: float) -> float:
    if v not in _DIGAMMA_CACHE:
        _DIGAMMA_CACHE[v] = float(torch.digamma(torch.tensor(v, dtype=torch.float64)))
    return _DIGAMMA_CACHE[v]


class _ElboDenoiseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, sigma_est, im_noisy, im_gt, beta0, eps2, alpha0):
        mu_c, sg_c = mu.contiguous(), sigma_est.contiguous()
        d_mu = torch.empty_like(mu_c)
        d_sigma = torch.empty_like(sg_c)
        out4 = ops.elbo_denoise(mu_c, sg_c, im_noisy.contiguous(), im_gt.contiguous(), beta0.contiguous(),
                                eps2=eps2, alpha0=alpha0, digamma_am1=_digamma(alpha0 - 1.0), d_mu=d_mu,
                                d_sigma=d_sigma)
        ctx.save_for_backward(d_mu, d_sigma)
        return out4[0], out4[1], out4[2], out4[3]

    @staticmethod
    def backward(ctx, g_loss, g_lh, g_kg, g_ig):
        # only `loss` is differentiated by the trainers; lh / kl terms are logged values
        d_mu, d_sigma = ctx.saved_tensors
        return d_mu * g_loss, d_sigma * g_loss, None, None, None, None, None


def elbo_denoising_simple(mu, sigma_est, im_noisy, im_gt, eps2, alpha0, beta0):
    if isinstance(mu, (list, tuple)):
        if len(mu) != 1:
            raise NotImplementedError("deep-supervision list form is unused by the shipped networks")
        mu = mu[0]
    alpha0_f = float(alpha0.item() if torch.is_tensor(alpha0) else alpha0)
    eps2_f = float(eps2)
    if not torch.is_tensor(beta0):
        raise TypeError("beta0 must be a tensor shaped like sigma_est")
    beta0 = beta0.expand_as(sigma_est)
    return _ElboDenoiseFn.apply(mu, sigma_est, im_noisy, im_gt, beta0, eps2_f, alpha0_f)
