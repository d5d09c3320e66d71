"""Drop-in replacements for the reference's networks/VIRNet.py model classes.

Same constructor arguments, `forward` signatures, sub-module attribute names and
`state_dict` layout as the reference (networks/VIRNet.py:18-46 and :48-97), so
train_denoising_syn.py / train_SISR.py / scripts/testing_demo.py can import these instead.
The forward and backward passes run as one CUDA program over libvirnet_sm100.so
(virnet_b200/engine.py); autograd sees the whole network as a single node, so DDP hooks,
`loss.backward()`, `clip_grad_norm_` and torch optimizers keep working unchanged.

Extra (non-reference) knob: `precision` — "tf32" (fp32 storage, TF32 tensor-core MMA; matches the
reference's default GPU numerics and the 1e-3 parity bar) or "bf16" (training throughput mode).
It can also be set with the VIRNET_B200_PRECISION environment variable.
"""
from __future__ import annotations

import os

import torch
from torch import nn

from .containers import AttResUNet, DnCNN, KernelNet


def _default_precision():
    return os.environ.get("VIRNET_B200_PRECISION", "tf32").lower()


def _param_grads(eng, params):
    """Per-parameter gradients for autograd: views into ONE private copy of the engine's flat gradient buffer (a single
    copy kernel; the engine's own buffer is overwritten by the next backward, and autograd may keep what it is given)."""
    flat = eng.flat_grads.clone()
    out = []
    for p in params:
        if not p.requires_grad:
            out.append(None)
            continue
        o = eng.flat_offsets[eng.param_index[id(p)]]
        out.append(flat[o:o + p.numel()].view_as(p))
    return tuple(out)


class _EngineMixin:
    def mark_params_dirty(self):
        """Call after writing parameters through `.data` (EMA, manual re-initialisation, `p.data.copy_`): such writes do
        not bump the tensors' version counters, so the packed tensor-core operands would otherwise stay stale.
        Optimizer steps, `load_state_dict` and in-place ops on the parameters themselves are detected automatically."""
        if self._engine is not None:
            self._engine.mark_params_dirty()

    def release_buffers(self):
        """Free every cached activation buffer of the engine (e.g. after a validation sweep over many image sizes)."""
        if self._engine is not None:
            self._engine.release_buffers()


class _WholeNetFn(torch.autograd.Function):
    """One autograd node for SNet + RNet: forward saves activations inside the engine,
    backward returns the parameter gradients computed by the dgrad / wgrad kernels."""

    @staticmethod
    def forward(ctx, x, engine, need_grad, *params):
        # grad mode is always off inside Function.forward and ctx.needs_input_grad ignores torch.no_grad(): the module
        # decides before apply() whether this forward will be differentiated (no_grad / eval forwards save nothing)
        mu, sigma = engine.forward(x, save=need_grad)
        ctx.engine = engine
        ctx.gen = engine.saved["gen"] if need_grad else None
        ctx.params = params
        return mu, sigma

    @staticmethod
    def backward(ctx, g_mu, g_sigma):
        eng = ctx.engine
        eng.backward(g_mu, g_sigma, gen=ctx.gen)
        return (None, None, None) + _param_grads(eng, ctx.params)


class _SRNetFn(torch.autograd.Function):
    """One autograd node for SNet + KNet + the SFT-modulated RNet (parameter gradients only: the reference never
    differentiates w.r.t. the LR image, train_SISR.py)."""

    @staticmethod
    def forward(ctx, x, sf, engine, need_grad, *params):
        mu, kinfo, sigma = engine.forward_sr(x, sf, save=need_grad)
        ctx.engine, ctx.params = engine, params
        ctx.gen = engine.saved["gen"] if need_grad else None
        return mu, kinfo, sigma

    @staticmethod
    def backward(ctx, g_mu, g_kinfo, g_sigma):
        eng = ctx.engine
        eng.backward_sr(g_mu, g_kinfo, g_sigma, gen=ctx.gen)
        return (None, None, None, None) + _param_grads(eng, ctx.params)


class VIRAttResUNet(_EngineMixin, nn.Module):
    """Denoising: sigma = exp(clamp(SNet(x))), mu = RNet(x, sqrt(sigma)); returns (mu, sigma)."""

    def __init__(self, im_chn, sigma_chn=3, n_feat=[64, 128, 192], dep_S=5, n_resblocks=2, noise_cond=True,
                 extra_mode="Input", noise_avg=False, precision=None):
        super().__init__()
        self.SNet = DnCNN(im_chn, sigma_chn, dep=dep_S, noise_avg=noise_avg)
        self.noise_cond = noise_cond
        extra_chn = sigma_chn if noise_cond else 0
        self.RNet = AttResUNet(im_chn, extra_chn=extra_chn, out_chn=im_chn, n_feat=n_feat,
                               n_resblocks=n_resblocks, extra_mode=extra_mode)
        self.precision = (precision or _default_precision()).lower()
        self._engine = None

    def engine(self):
        from ..engine import DenoiseEngine
        if self._engine is None or self._engine.precision != self.precision:
            object.__setattr__(self, "_engine", DenoiseEngine(self, self.precision))
        return self._engine

    def forward(self, x):
        eng = self.engine()
        params = tuple(self.parameters())
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        mu, sigma = _WholeNetFn.apply(x, eng, need_grad, *params)
        return mu, sigma


class VIRAttResUNetSR(_EngineMixin, nn.Module):
    """Super-resolution variant (SNet + KNet + SFT-modulated RNet); returns (mu, kinfo, sigma)."""

    def __init__(self, im_chn, sigma_chn=1, kernel_chn=3, n_feat=[64, 128, 192], dep_S=5, dep_K=8,
                 noise_cond=True, kernel_cond=True, n_resblocks=1, extra_mode="Down", noise_avg=True,
                 precision=None):
        super().__init__()
        self.noise_cond, self.noise_avg, self.kernel_cond = noise_cond, noise_avg, kernel_cond
        extra_chn = (kernel_chn if kernel_cond else 0) + (sigma_chn if noise_cond else 0)
        self.SNet = DnCNN(im_chn, sigma_chn, dep=dep_S, noise_avg=noise_avg)
        self.KNet = KernelNet(im_chn, kernel_chn, num_blocks=dep_K)
        self.RNet = AttResUNet(im_chn, extra_chn=extra_chn, out_chn=im_chn, n_feat=n_feat,
                               n_resblocks=n_resblocks, extra_mode=extra_mode)
        self.precision = (precision or _default_precision()).lower()
        self._engine = None

    def engine(self):
        from ..engine import DenoiseEngine
        if self._engine is None or self._engine.precision != self.precision:
            object.__setattr__(self, "_engine", DenoiseEngine(self, self.precision, sr=True))
        return self._engine

    def forward(self, x, sf):
        """(mu, kinfo_est [N,3], sigma [N,1,1,1]) as networks/VIRNet.py:80-97; differentiable w.r.t. the parameters."""
        eng = self.engine()
        params = tuple(self.parameters())
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        return _SRNetFn.apply(x, int(sf), eng, need_grad, *params)
