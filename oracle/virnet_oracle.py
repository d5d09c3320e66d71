"""CPU oracle for the VIRNet hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product path (virnet_b200/) never does.

What it is: a functional restatement, in plain fp32 torch CPU ops, of the reference's
forward path and losses.  The reference (zsyOAOA/VIRNet) is pure PyTorch: its arithmetic
lives in the third-party dependency `torch` (README.md:23 pins pytorch==1.13.0;
environment.yml:52 pins 1.12.0; this image has torch 2.11.0) through the call sites
cited on each function below.  Everything here works on a flat `state_dict` with the
reference's parameter names, so the same weights can be fed to the reference, to this
oracle and to the CUDA path.

Parity pinning: the reference ships no tests, golden vectors or checkpoints
(SURVEY.md §8c), so the oracle is pinned against outputs of the reference itself, run
in the build container by tools/gen_golden.py (which imports /root/reference) and
committed under tests/golden/.  tests/test_oracle_golden.py replays them.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

# networks/VIRNet.py:15-16 and networks/KNet.py:5-6
SNET_LOG_MAX, SNET_LOG_MIN = math.log(1e2), math.log(1e-10)
KNET_LOG_MAX, KNET_LOG_MIN = math.log(1e2), math.log(1e-4)


@dataclass
class NetCfg:
    """Constructor arguments of VIRAttResUNet / VIRAttResUNetSR (networks/VIRNet.py:22-29, 52-62)."""
    im_chn: int = 3
    sigma_chn: int = 1
    kernel_chn: int = 3
    n_feat: Sequence[int] = (96, 192, 288)
    dep_S: int = 5
    dep_K: int = 8
    n_resblocks: int = 3
    noise_cond: bool = True
    kernel_cond: bool = True
    extra_mode: str = "Input"
    noise_avg: bool = False
    sisr: bool = False

    @property
    def extra_chn(self) -> int:
        c = self.sigma_chn if self.noise_cond else 0
        if self.sisr and self.kernel_cond:
            c += self.kernel_chn
        return c


# --------------------------------------------------------------------------------------
# parameter construction in the reference's creation order (drives the RNG stream)
# --------------------------------------------------------------------------------------
def _conv_params(sd: SD, name: str, cin: int, cout: int, k: int, bias: bool = True, transposed: bool = False):
    """Default nn.Conv2d / nn.ConvTranspose2d init (kaiming_uniform(a=sqrt(5)) + uniform bias)."""
    m = (torch.nn.ConvTranspose2d(cin, cout, k, stride=k) if transposed
         else torch.nn.Conv2d(cin, cout, k, bias=bias))
    sd[name + ".weight"] = m.weight.detach()
    if m.bias is not None:
        sd[name + ".bias"] = m.bias.detach()


def _att_layer_params(sd: SD, name: str, out_chn: int, extra_chn: int):
    nf1, nf2 = out_chn // 8, out_chn // 4                       # networks/AttResUNet.py:15-16
    _conv_params(sd, name + ".conv1", extra_chn, nf1, 1)
    _conv_params(sd, name + ".conv2", nf1, nf2, 1)
    _conv_params(sd, name + ".mul_conv", nf2, out_chn, 1)
    _conv_params(sd, name + ".add_conv", nf2, out_chn, 1)


def build_state_dict(cfg: NetCfg) -> SD:
    """Create parameters exactly as the reference constructors do, consuming the global torch
    RNG in the same order: SNet -> (KNet) -> RNet (networks/VIRNet.py:31-40, 66-78)."""
    sd: SD = {}
    # --- SNet: DnCNN (networks/DnCNN.py:22-29), orthogonal init + zero bias (:46-52)
    names = []
    _conv_params(sd, "SNet.conv1", cfg.im_chn, 64, 3); names.append("SNet.conv1")
    for ii in range(1, cfg.dep_S - 1):
        nm = f"SNet.mid_layer.{2 * (ii - 1)}"
        _conv_params(sd, nm, 64, 64, 3); names.append(nm)
    _conv_params(sd, "SNet.conv_last", 64, cfg.sigma_chn, 3); names.append("SNet.conv_last")
    gain = torch.nn.init.calculate_gain("leaky_relu", 0.25)
    for nm in names:                                             # module order == creation order
        torch.nn.init.orthogonal_(sd[nm + ".weight"], gain=gain)
        sd[nm + ".bias"].zero_()
    # --- KNet (networks/KNet.py:41-50)
    if cfg.sisr:
        _conv_params(sd, "KNet.head", cfg.im_chn, 64, 9, bias=False)
        for b in range(cfg.dep_K):
            _conv_params(sd, f"KNet.body.{b}.body.0", 64, 64, 3)
            _conv_params(sd, f"KNet.body.{b}.body.2", 64, 64, 3)
            _conv_params(sd, f"KNet.body.{b}.body.3.body.0", 64, 64 // 16, 1)
            _conv_params(sd, f"KNet.body.{b}.body.3.body.2", 64 // 16, 64, 1)
        _conv_params(sd, "KNet.tail.0", 64, cfg.kernel_chn, 3)
    # --- RNet: AttResUNet (networks/AttResUNet.py:109-139)
    mode = cfg.extra_mode.lower()
    assert mode in ("null", "input", "down", "both")
    nf = list(cfg.n_feat)
    depth = len(nf)
    extra = cfg.extra_chn
    head_in = cfg.im_chn if mode in ("down", "null") else cfg.im_chn + extra
    _conv_params(sd, "RNet.head", head_in, nf[0], 3)
    extra_down = extra if mode in ("down", "both") else 0
    for ii in range(depth):
        for b in range(cfg.n_resblocks):
            p = f"RNet.down_path.{ii}.body.{b}"
            if extra_down > 0:
                _att_layer_params(sd, p + ".sft1", nf[ii], extra_down)
                _att_layer_params(sd, p + ".sft2", nf[ii], extra_down)
            _conv_params(sd, p + ".conv1", nf[ii], nf[ii], 3)
            _conv_params(sd, p + ".conv2", nf[ii], nf[ii], 3)
        if ii + 1 < depth:
            _conv_params(sd, f"RNet.down_path.{ii}.downsampler", nf[ii], nf[ii + 1], 3)
    for k, jj in enumerate(reversed(range(depth - 1))):
        _conv_params(sd, f"RNet.up_path.{k}.upsampler", nf[jj + 1], nf[jj], 2, transposed=True)
        for b in range(cfg.n_resblocks):
            p = f"RNet.up_path.{k}.body.{b}"
            _conv_params(sd, p + ".conv1", nf[jj], nf[jj], 3)
            _conv_params(sd, p + ".conv2", nf[jj], nf[jj], 3)
    _conv_params(sd, "RNet.tail", nf[0], cfg.im_chn, 3)
    return sd


# --------------------------------------------------------------------------------------
# forward restatement
# --------------------------------------------------------------------------------------
def _conv(sd: SD, name: str, x: Tensor, stride: int = 1, padding: int = 1) -> Tensor:
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride=stride, padding=padding)


def pad_input(x: Tensor, mod: int) -> Tensor:
    """utils/util_net.py:20-25 — reflect pad bottom/right up to a multiple of `mod`."""
    h, w = x.shape[-2:]
    bottom = int(math.ceil(h / mod) * mod - h)
    right = int(math.ceil(w / mod) * mod - w)
    return F.pad(x, pad=(0, right, 0, bottom), mode="reflect")


def dncnn(sd: SD, x: Tensor, cfg: NetCfg) -> Tensor:
    """networks/DnCNN.py:37-44 — conv+LReLU(0.25) x (dep-1), conv; optional global average."""
    h = F.leaky_relu(_conv(sd, "SNet.conv1", x), 0.25)
    for ii in range(1, cfg.dep_S - 1):
        h = F.leaky_relu(_conv(sd, f"SNet.mid_layer.{2 * (ii - 1)}", h), 0.25)
    out = _conv(sd, "SNet.conv_last", h)
    if cfg.noise_avg:
        out = out.mean(dim=(2, 3), keepdim=True)                # nn.AdaptiveAvgPool2d((1,1)) DnCNN.py:30-33
    return out


def att_layer(sd: SD, p: str, extra: Tensor) -> Tuple[Tensor, Tensor]:
    """networks/AttResUNet.py:27-32 — SFT-style (mul, add) from the conditioning maps."""
    f1 = F.leaky_relu(_conv(sd, p + ".conv1", extra, padding=0), 0.2)
    f2 = F.leaky_relu(_conv(sd, p + ".conv2", f1, padding=0), 0.2)
    mul = torch.sigmoid(_conv(sd, p + ".mul_conv", f2, padding=0))
    add = _conv(sd, p + ".add_conv", f2, padding=0)
    return mul, add


def att_res_block(sd: SD, p: str, x: Tensor, extra: Optional[Tensor]) -> Tensor:
    """networks/AttResUNet.py:48-60 — pre-activation residual block with optional modulation."""
    if extra is not None:
        mul1, add1 = att_layer(sd, p + ".sft1", extra)
        f1 = _conv(sd, p + ".conv1", F.leaky_relu(x * mul1 + add1, 0.2))
        mul2, add2 = att_layer(sd, p + ".sft2", extra)
        f2 = _conv(sd, p + ".conv2", F.leaky_relu(f1 * mul2 + add2, 0.2))
    else:
        f1 = _conv(sd, p + ".conv1", F.leaky_relu(x, 0.2))
        f2 = _conv(sd, p + ".conv2", F.leaky_relu(f1, 0.2))
    return x + f2


def att_res_unet(sd: SD, x_in: Tensor, extra_in: Optional[Tensor], cfg: NetCfg) -> Tensor:
    """networks/AttResUNet.py:141-175."""
    mode = cfg.extra_mode.lower()
    depth = len(cfg.n_feat)
    h, w = x_in.shape[-2:]
    x = pad_input(x_in, 2 ** (depth - 1))
    extra = pad_input(extra_in, 2 ** (depth - 1)) if mode != "null" else None
    if mode in ("input", "both"):
        x = _conv(sd, "RNet.head", torch.cat([x, extra], 1))
    else:
        x = _conv(sd, "RNet.head", x)
    blocks: List[Tensor] = []
    extra_down = [extra] if mode in ("down", "both") else None
    for ii in range(depth):
        e = extra_down[ii] if extra_down is not None else None
        for b in range(cfg.n_resblocks):
            x = att_res_block(sd, f"RNet.down_path.{ii}.body.{b}", x, e)
        if ii != depth - 1:
            blocks.append(x)
            x = _conv(sd, f"RNet.down_path.{ii}.downsampler", x, stride=2)
            if extra_down is not None:
                extra_down.append(F.interpolate(extra, x.shape[-2:], mode="nearest"))
    for k in range(depth - 1):
        up = f"RNet.up_path.{k}"
        x = F.conv_transpose2d(x, sd[up + ".upsampler.weight"], sd[up + ".upsampler.bias"], stride=2)
        x = x + blocks[-k - 1]
        for b in range(cfg.n_resblocks):
            x = att_res_block(sd, f"{up}.body.{b}", x, None)
    return _conv(sd, "RNet.tail", x)[..., :h, :w] + x_in


def kernel_net(sd: SD, x: Tensor, cfg: NetCfg) -> Tensor:
    """networks/KNet.py:52-59 (+ RB_Layer :37-39, CALayer :23-26)."""
    h = F.conv2d(x, sd["KNet.head.weight"], None, stride=4, padding=4)
    for b in range(cfg.dep_K):
        p = f"KNet.body.{b}.body"
        f = F.leaky_relu(_conv(sd, p + ".0", h), 0.2)
        f = _conv(sd, p + ".2", f)
        y = f.mean(dim=(2, 3), keepdim=True)
        y = F.leaky_relu(_conv(sd, p + ".3.body.0", y, padding=0), 0.2)
        y = torch.sigmoid(_conv(sd, p + ".3.body.2", y, padding=0))
        h = f * y + h
    out = _conv(sd, "KNet.tail.0", h).mean(dim=(2, 3), keepdim=True)
    lam12 = torch.exp(torch.clamp(out[:, :2], min=KNET_LOG_MIN, max=KNET_LOG_MAX))
    rho = torch.tanh(out[:, -1]).unsqueeze(1)
    return torch.cat((lam12, rho), dim=1)


def vir_denoise_forward(sd: SD, x: Tensor, cfg: NetCfg) -> Tuple[Tensor, Tensor]:
    """networks/VIRNet.py:42-46."""
    sigma = torch.exp(torch.clamp(dncnn(sd, x, cfg), min=SNET_LOG_MIN, max=SNET_LOG_MAX))
    extra = sigma.sqrt() if cfg.noise_cond else None
    mu = att_res_unet(sd, x, extra, cfg)
    return mu, sigma


def vir_sisr_forward(sd: SD, x: Tensor, sf: int, cfg: NetCfg) -> Tuple[Tensor, Tensor, Tensor]:
    """networks/VIRNet.py:80-97."""
    sigma = torch.exp(torch.clamp(dncnn(sd, x, cfg), min=SNET_LOG_MIN, max=SNET_LOG_MAX))
    kinfo = kernel_net(sd, x, cfg)
    x_up = F.interpolate(x, scale_factor=sf, mode="nearest")
    h_up, w_up = x_up.shape[-2:]
    extra = None
    if cfg.noise_cond or cfg.kernel_cond:
        parts = []
        if cfg.kernel_cond:
            parts.append(kinfo.repeat(1, 1, h_up, w_up))
        if cfg.noise_cond:
            if cfg.noise_avg:
                parts.append(sigma.sqrt().repeat(1, 1, h_up, w_up))
            else:
                parts.append(F.interpolate(sigma.sqrt(), scale_factor=sf, mode="nearest"))
        extra = torch.cat(parts, 1)
    mu = att_res_unet(sd, x_up, extra, cfg)
    return mu, kinfo.squeeze(-1).squeeze(-1), sigma


# --------------------------------------------------------------------------------------
# loss restatement (loss/ELBO_simple.py)
# --------------------------------------------------------------------------------------
def kl_inverse_gamma_simple(beta_q: Tensor, alpha_p, beta_p: Tensor) -> Tensor:
    """loss/ELBO_simple.py:12-14."""
    return (alpha_p * (beta_p / beta_q - 1) + alpha_p * (beta_q.log() - beta_p.log())).mean()


def kl_gauss_simple(mu_q: Tensor, mu_p: Tensor, var_p) -> Tensor:
    """loss/ELBO_simple.py:16."""
    return 0.5 * ((mu_q - mu_p) ** 2 / var_p).mean()


def likelihood(x: Tensor, mu_q: Tensor, var_q, alpha_q, beta_q: Tensor) -> Tensor:
    """loss/ELBO_simple.py:18-21."""
    alpha_q = torch.as_tensor(alpha_q, dtype=torch.float32)
    t = 0.5 * (beta_q.log() - torch.digamma(alpha_q) + alpha_q / beta_q * ((x - mu_q) ** 2 + var_q))
    return (t + 0.5 * math.log(2 * math.pi)).mean()


def elbo_denoising_simple(mu, sigma_est, im_noisy, im_gt, eps2, alpha0, beta0):
    """loss/ELBO_simple.py:23-53 (single-output form; the list form is unused by the shipped nets)."""
    kl_gauss = kl_gauss_simple(mu, im_gt, eps2)
    beta = sigma_est * alpha0
    kl_ig = kl_inverse_gamma_simple(beta, alpha0 - 1, beta0)
    lh = likelihood(im_noisy, mu, eps2, alpha0 - 1, beta)
    return lh + kl_gauss + kl_ig, lh, kl_gauss, kl_ig


# --------------------------------------------------------------------------------------
# SISR loss restatement (loss/ELBO_simple.py:55-138, utils/util_sisr.py:26-58,127-144,
# ResizeRight/resize_right.py:29-76,146-320, ResizeRight/interp_methods.py:34-42)
# --------------------------------------------------------------------------------------
def _cubic(x):
    """ResizeRight/interp_methods.py:34-42 (Keys cubic, a = -0.5), numpy float64."""
    import numpy as np
    ax = np.abs(x)
    ax2, ax3 = ax ** 2, ax ** 3
    return ((1.5 * ax3 - 2.5 * ax2 + 1.0) * (ax <= 1.0) +
            (-0.5 * ax3 + 2.5 * ax2 - 4.0 * ax + 2.0) * ((1.0 < ax) & (ax <= 2.0)))


def resize_matrix(in_sz: int, sf: int, downsampler: str = "bicubic"):
    """The 1-D operator applied along one axis by `conv_multi_kernel_tensor`'s down-sampler, as a dense
    [out_sz, in_sz] float32 matrix.  'direct': rows select every sf-th sample (util_sisr.py:137-138).
    'bicubic': ResizeRight.resize(scale_factors=1/sf), antialiased cubic — projected grid
    (resize_right.py:251-262), field of view with mirror folding (:265-285), stretched kernel
    (:304-315), weights normalised per output (:288-299); torch float32 arithmetic as the reference runs it."""
    import numpy as np
    if downsampler.lower() == "direct":
        out_sz = (in_sz + sf - 1) // sf
        m = np.zeros((out_sz, in_sz), dtype=np.float32)
        m[np.arange(out_sz), np.arange(out_sz) * sf] = 1.0
        return torch.from_numpy(m)
    if downsampler.lower() != "bicubic":
        raise ValueError("downsampler must be Direct or Bicubic")
    scale = 1.0 / sf
    out_sz = int(math.ceil(scale * in_sz))
    eps = float(np.finfo(np.float32).eps)
    support = 4.0 / scale
    # the reference evaluates these in torch: arange is int64, divisions promote to float32
    grid = (torch.arange(out_sz) / scale + (in_sz - 1) / 2 - (out_sz - 1) / (2 * scale))
    left = (grid - support / 2 - eps).ceil().long()
    fov = left[:, None] + torch.arange(int(math.ceil(support - eps)))
    mirror = torch.cat((torch.arange(in_sz), torch.arange(in_sz - 1, -1, -1)))
    fov = mirror[torch.remainder(fov, mirror.shape[0])]
    d = (grid[:, None] - fov).to(torch.float32)
    w = scale * torch.from_numpy(_cubic((scale * d).numpy().astype(np.float32)).astype(np.float32))
    sw = w.sum(1, keepdim=True)
    sw[sw == 0] = 1
    w = w / sw
    m = torch.zeros(out_sz, in_sz, dtype=torch.float32)
    m.scatter_add_(1, fov, w)
    return m


def sigma2kernel(k_cov: Tensor, k_size: int, sf: int, shift: bool) -> Tensor:
    """utils/util_sisr.py:26-58: softmax over the k x k grid of -0.5 z^T Sigma^-1 z; N x 1 x k x k."""
    inv = torch.inverse(k_cov)
    center = k_size // 2 + 0.5 * (sf - k_size % 2) if shift else k_size // 2
    ax = torch.arange(k_size, dtype=k_cov.dtype)
    X, Y = torch.meshgrid(ax, ax, indexing="ij")
    Z = torch.stack((X, Y), dim=2).view(1, -1, 2, 1) - center
    q = -0.5 * Z.transpose(2, 3).matmul(inv).matmul(Z).squeeze(-1).squeeze(-1)
    return F.softmax(q, dim=1).view(-1, 1, k_size, k_size)


def blur_downsample(im_hr: Tensor, kernel: Tensor, sf: int, downsampler: str) -> Tensor:
    """utils/util_sisr.py:127-144: reflect pad k//2, per-sample cross-correlation (grouped conv3d), then
    ::sf sampling or the antialiased cubic resize along H and W."""
    n, c, h, w = im_hr.shape
    k = kernel.shape[-1]
    pad = F.pad(im_hr, (k // 2,) * 4, mode="reflect")
    blur = F.conv2d(pad.reshape(1, n * c, h + k - 1, w + k - 1),
                    kernel.expand(n, c, k, k).reshape(n * c, 1, k, k), groups=n * c).view(n, c, h, w)
    mh, mw = resize_matrix(h, sf, downsampler), resize_matrix(w, sf, downsampler)
    return torch.einsum("yh,nchw,xw->ncyx", mh.to(blur), blur, mw.to(blur))


def elbo_sisr(mu, sigma_est, kinfo_est, im_hr, im_lr, sigma_prior, alpha0, kinfo_gt, kappa0, r2, eps2, sf, k_size,
              penalty_K, shift, downsampler, *, gamma_draw, rho_draw, z_draw):
    """loss/ELBO_simple.py:82-138 with the three random draws made explicit:
      gamma_draw [N,2] ~ Gamma(kappa0-1, 1)  (Gamma(alpha, beta).rsample() == gamma_draw / beta, :61-64)
      rho_draw   [N,1] ~ N(0,1)              (:76)
      z_draw     like mu ~ N(0,1)            (:56)
    Returns (loss, [lh, kl_rnet, kl_snet, kl_knet, kl_knet0, kl_knet1, kl_knet2, kernel])."""
    kl_rnet = kl_gauss_simple(mu, im_hr, eps2)
    beta0 = sigma_prior * alpha0
    beta = sigma_est * alpha0
    kl_snet = kl_inverse_gamma_simple(beta, alpha0 - 1, beta0)
    kl_k0 = kl_inverse_gamma_simple(kappa0 * kinfo_est[:, 0], kappa0 - 1, kappa0 * kinfo_gt[:, 0])
    kl_k1 = kl_inverse_gamma_simple(kappa0 * kinfo_est[:, 1], kappa0 - 1, kappa0 * kinfo_gt[:, 1])
    kl_k2 = kl_gauss_simple(kinfo_est[:, 2], kinfo_gt[:, 2], r2) * penalty_K[0]
    kl_knet = (kl_k0 + kl_k1 + kl_k2) / 3 * penalty_K[1]
    # reparameterised covariance (:66-80)
    k_var = 1.0 / (gamma_draw / (kinfo_est[:, :2] * kappa0))
    v1, v2 = torch.chunk(k_var, 2, dim=1)
    rho = kinfo_est[:, 2].unsqueeze(1) + math.sqrt(r2) * rho_draw
    direction = v1.detach().sqrt() * v2.detach().sqrt() * torch.clamp(rho, min=-1, max=1)
    k_cov = torch.cat([v1, direction, direction, v2], dim=1).view(-1, 1, 2, 2)
    kernel = sigma2kernel(k_cov, k_size, sf, shift)
    # likelihood (:55-59)
    zz = mu + z_draw * math.sqrt(eps2)
    zz_blur = blur_downsample(zz, kernel, sf, downsampler)
    alpha_q = torch.as_tensor(alpha0 - 1, dtype=torch.float32)
    lh = (0.5 * math.log(2 * math.pi) + 0.5 * (beta.log() - torch.digamma(alpha_q)) +
          0.5 * alpha_q / beta * (im_lr - zz_blur) ** 2).mean()
    loss = lh + kl_rnet + kl_snet + kl_knet
    return loss, [lh, kl_rnet, kl_snet, kl_knet, kl_k0, kl_k1, kl_k2, kernel]


def reference_draws(n: int, mu_shape, kappa0: float, generator=None, device="cpu"):
    """The reference's draw ORDER (Gamma rsample -> randn_like(rho) -> randn_like(mu)) on torch's global
    generator: calling this right after the same `torch.manual_seed` reproduces what elbo_sisr consumes."""
    conc = torch.full((n, 2), float(kappa0 - 1), device=device)
    gamma_draw = torch._standard_gamma(conc)
    rho_draw = torch.randn(n, 1, device=device)
    z_draw = torch.randn(*mu_shape, device=device)
    return gamma_draw, rho_draw, z_draw


# --------------------------------------------------------------------------------------
# real-noise trainer's data-side operators (utils/util_denoising.py:24-63, datasets/data_tools.py:12-30)
# --------------------------------------------------------------------------------------
def inverse_gamma_window(k_size: int) -> Tensor:
    """utils/util_denoising.py:24-41: cv2.getGaussianKernel(k, s) (s > 0: exp(-(i-(k-1)/2)^2 / 2s^2), sum 1)
    outer product, renormalised; float64 -> float32."""
    scale = 0.3 * ((k_size - 1) * 0.5 - 1) + 0.8
    i = torch.arange(k_size, dtype=torch.float64)
    k1 = torch.exp(-((i - (k_size - 1) / 2) ** 2) / (2 * scale ** 2))
    k1 = k1 / k1.sum()
    k2 = k1[:, None] * k1[None, :]
    return (k2 / k2.sum()).to(torch.float32)


def noise_estimate_fun(im_noisy: Tensor, im_gt: Tensor, k_size: int) -> Tensor:
    """utils/util_denoising.py:43-63: depthwise Gaussian-window mean of the squared error, reflect pad, clamp 1e-10."""
    c = im_noisy.shape[1]
    kernel = inverse_gamma_window(k_size).expand(c, 1, k_size, k_size)
    err2 = (im_noisy - im_gt) ** 2
    out = F.conv2d(F.pad(err2, [k_size // 2] * 4, mode="reflect"), kernel, groups=c)
    return out.clamp_(min=1e-10)


def mixup(rgb_gt: Tensor, rgb_noisy: Tensor, indices: Tensor, lam: Tensor):
    """datasets/data_tools.py:21-30 with the draws (randperm, Beta(0.6, 0.6) rsample) made explicit."""
    lam = lam.view(-1, 1, 1, 1)
    return lam * rgb_gt + (1 - lam) * rgb_gt[indices], lam * rgb_noisy + (1 - lam) * rgb_noisy[indices]


# --------------------------------------------------------------------------------------
# synthetic denoising batches (datasets/DenoisingDatasets.py:180-253, utils/util_denoising.py:12-22,
# utils/util_image.py:391-436) — numpy, one sample at a time like the reference
# --------------------------------------------------------------------------------------
def data_aug_np(image, mode: int):
    """utils/util_image.py:391-436."""
    import numpy as np
    if mode == 0:
        out = image
    elif mode == 1:
        out = np.flipud(image)
    elif mode in (2, 3):
        out = np.rot90(image)
    elif mode in (4, 5):
        out = np.rot90(image, k=2)
    elif mode in (6, 7):
        out = np.rot90(image, k=3)
    else:
        raise ValueError("Invalid choice of image transformation")
    if mode in (3, 5, 7):
        out = np.flipud(out)
    return out.copy()


def synth_denoise_sample(patch_u8, params, aug_flag: int, noise, clip: bool = False):
    """SimulateTrain.__getitem__ (:216-253) after crop_patch, with its random draws as arguments:
    params = [center_h, center_w, scale, up (after swap and + 5/255), down, iid_level]; scale <= 0 -> 'iid'.
    patch_u8 HWC uint8, noise HWC float32 ~ N(0,1).  Returns (im_noisy, im_gt, sigma_map_gt) CHW float32 tensors."""
    import numpy as np
    P = patch_u8.shape[0]
    im_gt = np.multiply(np.asarray(patch_u8), 1.0 / 255, dtype=np.float32)      # skimage.img_as_float32 on uint8
    ch, cw, scale, up, down, level = [float(v) for v in params]
    if scale > 0:
        ii, jj = [x.astype(np.float64) for x in np.meshgrid(np.arange(P), np.arange(P), indexing="ij")]
        kk = np.exp((-(ii - ch) ** 2 - (jj - cw) ** 2) / (2 * scale ** 2))
        kk /= kk.sum()
        sigma_map = (down + (kk - kk.min()) / (kk.max() - kk.min()) * (up - down)).astype(np.float32)
    else:
        sigma_map = (np.ones([P, P]) * level).astype(np.float32)
    sigma_map = sigma_map[:, :, np.newaxis]
    nz = np.asarray(noise, dtype=np.float32) * sigma_map
    im_noisy = im_gt + nz.astype(np.float32)
    if clip:
        im_noisy = np.clip(im_noisy, 0.0, 1.0)
    im_gt, im_noisy, sigma_map = [data_aug_np(x, aug_flag) for x in (im_gt, im_noisy, sigma_map)]
    sg = np.square(sigma_map)
    sg = np.where(sg < 1e-10, 1e-10, sg)
    tt = lambda x: torch.from_numpy(x.transpose((2, 0, 1)).copy())
    return tt(im_noisy), tt(im_gt), tt(sg.astype(np.float32))


# --------------------------------------------------------------------------------------
# SISR training-pair synthesis (datasets/SISRDatasets.py:86-104, utils/util_sisr.py:60-93,108-125) — numpy
# --------------------------------------------------------------------------------------
def sisr_degrade_sample(im_hr, kernel, sf: int, downsampler: str, noise, std: float):
    """One sample, HWC float32: scipy.ndimage.convolve(mode='reflect') restated as a correlation with the flipped
    kernel on np.pad(mode='symmetric'), clip, Direct / ResizeRight-bicubic down-sampling (the numpy path of ResizeRight
    computes in float64), Gaussian noise, clip.  Returns (im_blur, im_lr) HWC float32."""
    import numpy as np
    k = kernel.shape[0]
    r = k // 2
    H, W, C = im_hr.shape
    pad = np.pad(im_hr.astype(np.float64), ((r, r), (r, r), (0, 0)), mode="symmetric")
    kf = kernel[::-1, ::-1].astype(np.float64)
    blur = np.zeros((H, W, C), dtype=np.float64)
    for i in range(k):
        for j in range(k):
            blur += kf[i, j] * pad[i:i + H, j:j + W]
    blur = np.clip(blur.astype(np.float32), 0.0, 1.0)
    if downsampler.lower() == "direct":
        im_blur = blur[::sf, ::sf]
    else:
        mh = resize_matrix(H, sf, "bicubic").double().numpy()
        mw = resize_matrix(W, sf, "bicubic").double().numpy()
        im_blur = np.einsum("yh,hwc,xw->yxc", mh, blur.astype(np.float64), mw).astype(np.float32)
    im_lr = np.clip(im_blur + np.asarray(noise, dtype=np.float32) * np.float32(std), 0.0, 1.0).astype(np.float32)
    return im_blur.astype(np.float32), im_lr


# --------------------------------------------------------------------------------------
# one reference training step (train_denoising_syn.py:175-184), used as the CPU baseline
# --------------------------------------------------------------------------------------
def clip_grad_norm_(params: List[Tensor], max_norm: float) -> Tensor:
    """torch.nn.utils.clip_grad_norm_ semantics (L2, eps 1e-6, coefficient clamped to 1)."""
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(p.grad) for p in params]))
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for p in params:
        p.grad.mul_(coef)
    return total


def denoise_loss_and_grads(sd: SD, cfg: NetCfg, im_noisy, im_gt, sigma_gt, alpha0=24.5, eps2=1e-6):
    """Forward + ELBO + backward through autograd on the restated graph; returns
    (loss tuple, mu, sigma, grads dict)."""
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    mu, sigma = vir_denoise_forward(leaf, im_noisy, cfg)
    beta0 = alpha0 * sigma_gt
    loss, lh, klg, klig = elbo_denoising_simple(mu, sigma, im_noisy, im_gt, eps2, alpha0, beta0)
    loss.backward()
    grads = {k: v.grad for k, v in leaf.items()}
    return (loss.detach(), lh.detach(), klg.detach(), klig.detach()), mu.detach(), sigma.detach(), grads
