"""Fused replacements for the reference's loss/ELBO_simple.py: elbo_denoising_simple and elbo_sisr.

Same call signature and return tuple `(loss, lh, kl_gauss, kl_Igamma)` (reference
loss/ELBO_simple.py:23-53).  One CUDA kernel computes the three means and the gradients
w.r.t. `mu` and `sigma_est` in a single pass over the data; autograd receives them through
a custom Function, so `loss.backward()` works as with the reference.
"""
from __future__ import annotations

import math

import torch

from .. import ops
from .resize_right import downsample_matrix

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
        # deep-supervision form (loss/ELBO_simple.py:30-36,44-49): lh and kl_gauss are averaged over the list, the
        # inverse-Gamma KL is shared -> the mean of the per-element fused losses, term by term
        outs = [elbo_denoising_simple(m, sigma_est, im_noisy, im_gt, eps2, alpha0, beta0) for m in mu]
        return tuple(sum(o[i] for o in outs) / len(outs) for i in range(4))
    alpha0_f = float(alpha0.item() if torch.is_tensor(alpha0) else alpha0)
    eps2_f = float(eps2)
    if not torch.is_tensor(beta0):
        raise TypeError("beta0 must be a tensor shaped like sigma_est")
    beta0 = beta0.expand_as(sigma_est)
    return _ElboDenoiseFn.apply(mu, sigma_est, im_noisy, im_gt, beta0, eps2_f, alpha0_f)


# ---------------------------------------------------------------------------
# super-resolution (loss/ELBO_simple.py:82-138)
# ---------------------------------------------------------------------------
def _f(v) -> float:
    return float(v.item() if torch.is_tensor(v) else v)


def sisr_draws(kinfo_est, mu, kappa0):
    """The random draws elbo_sisr consumes, made with torch's generator in the reference's order
    (Gamma(kappa0-1, beta).rsample() == standard_gamma / beta, ELBO_simple.py:61-64; randn_like(rho) :76;
    randn_like(mu) :56) so that the same seed gives the same loss as the reference on the same device."""
    conc = torch.full((kinfo_est.shape[0], 2), kappa0 - 1.0, device=kinfo_est.device, dtype=torch.float32)
    gamma_draw = torch._standard_gamma(conc)
    rho_draw = torch.randn(kinfo_est.shape[0], 1, device=kinfo_est.device, dtype=torch.float32)
    z_draw = torch.randn_like(mu, dtype=torch.float32)
    return gamma_draw, rho_draw, z_draw


class _ElboSisrFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, sigma_est, kinfo_est, im_hr, im_lr, prior_mean, prior_logmean, kinfo_gt, draws, rh, rw, hyper):
        gamma_draw, rho_draw, z_draw = draws
        n = mu.shape[0]
        terms, kernel, d_mu, d_sigma, d_kinfo = ops.elbo_sisr(
            mu.contiguous().float(), im_hr.contiguous().float(), im_lr.contiguous().float(),
            sigma_est.reshape(n).contiguous().float(), kinfo_est.contiguous().float(), kinfo_gt.contiguous().float(),
            prior_mean, prior_logmean, gamma_draw.contiguous(), rho_draw.reshape(n).contiguous(), z_draw.contiguous(),
            rh, rw, **hyper)
        ctx.save_for_backward(d_mu, d_sigma, d_kinfo)
        ctx.sigma_shape = sigma_est.shape
        ctx.mark_non_differentiable(kernel)
        return (terms[0], terms[1], terms[2], terms[3], terms[4], terms[5], terms[6], terms[7], kernel)

    @staticmethod
    def backward(ctx, g_loss, *unused):
        # the trainers differentiate `loss` only (train_SISR.py:224); the other entries are logged values
        d_mu, d_sigma, d_kinfo = ctx.saved_tensors
        return (d_mu * g_loss, (d_sigma * g_loss).view(ctx.sigma_shape), d_kinfo * g_loss) + (None,) * 9


def elbo_sisr(mu, sigma_est, kinfo_est, im_hr, im_lr, sigma_prior, alpha0, kinfo_gt, kappa0, r2, eps2, sf, k_size,
              penalty_K, shift, downsampler, draws=None):
    """Same arguments and return value as the reference: (loss, [lh, kl_rnet, kl_snet, kl_knet, kl_knet0,
    kl_knet1, kl_knet2, kernel]).  `draws` (extra, optional): (gamma_draw [N,2], rho_draw [N,1], z_draw like mu)
    to replace the internal random draws — used by the parity tests."""
    if isinstance(mu, (list, tuple)):
        # deep-supervision form (loss/ELBO_simple.py:105-110,128-133): ONE kernel re-parameterisation, one z draw per
        # list element (the reference's generator order), lh and kl_rnet averaged, the other terms shared
        if draws is None:
            g_draw, r_draw, z0 = sisr_draws(kinfo_est, mu[0], _f(kappa0))
            zs = [z0] + [torch.randn_like(m, dtype=torch.float32) for m in mu[1:]]
        else:
            g_draw, r_draw, zs = draws[0], draws[1], list(draws[2])
        outs = [elbo_sisr(m, sigma_est, kinfo_est, im_hr, im_lr, sigma_prior, alpha0, kinfo_gt, kappa0, r2, eps2, sf, k_size,
                          penalty_K, shift, downsampler, draws=(g_draw, r_draw, z)) for m, z in zip(mu, zs)]
        loss = sum(o[0] for o in outs) / len(outs)
        detail = [sum(o[1][i] for o in outs) / len(outs) for i in range(7)] + [outs[0][1][7]]
        return loss, detail
    if sigma_est.numel() != mu.shape[0]:
        raise NotImplementedError("elbo_sisr expects the per-image noise variance of noise_avg=True (N x 1 x 1 x 1)")
    n, _, H, W = mu.shape
    h, w = im_lr.shape[2], im_lr.shape[3]
    alpha0_f, kappa0_f, sf, k_size = _f(alpha0), _f(kappa0), int(sf), int(k_size)
    if draws is None:
        draws = sisr_draws(kinfo_est, mu, kappa0_f)
    rh = downsample_matrix(H, sf, downsampler, mu.device)
    rw = downsample_matrix(W, sf, downsampler, mu.device)
    if rh.shape[0] != h or rw.shape[0] != w:
        raise ValueError(f"im_lr is {h}x{w} but down-sampling {H}x{W} by {sf} gives {rh.shape[0]}x{rw.shape[0]}")
    # per-image statistics of the prior variance (N x 1 x 1 x 1 for Gaussian noise, a map with add_jpeg)
    sp = sigma_prior.float().expand(n, *sigma_prior.shape[1:]).reshape(n, -1)
    prior_mean, prior_logmean = sp.mean(1).contiguous(), sp.log().mean(1).contiguous()
    center = k_size // 2 + 0.5 * (sf - k_size % 2) if shift else float(k_size // 2)
    hyper = dict(k_size=k_size, center=float(center), alpha0=alpha0_f, digamma_am1=_digamma(alpha0_f - 1.0),
                 kappa0=kappa0_f, r2=float(r2), eps2=float(eps2), pk0=float(penalty_K[0]), pk1=float(penalty_K[1]))
    out = _ElboSisrFn.apply(mu, sigma_est, kinfo_est, im_hr, im_lr, prior_mean, prior_logmean, kinfo_gt, tuple(draws),
                            rh, rw, hyper)
    return out[0], list(out[1:])
