"""Tensor-level wrappers over the C ABI (device pointers come from torch tensors).

PyTorch is plumbing here: it owns device memory and the CUDA stream.  All
arithmetic on the hot path happens inside libvirnet_sm100.so.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import lib as _l
from .lib import (VK_BF16, VK_TF32, VK_CONV3X3_S1, VK_CONV3X3_S2, VK_CONVT2X2_S2, VK_CONV1X1,
                  VK_CONV2X2_S2, VK_CONV3X3_S2_DGRAD,
                  VK_EPI_STD, VK_EPI_NCHW_F32)

TORCH_DTYPE = {VK_BF16: torch.bfloat16, VK_TF32: torch.float32}


def chan_align(dtype: int) -> int:
    """Channel pitch granularity: one UMMA K step is 32 bytes."""
    return 16 if dtype == VK_BF16 else 8


def chan_pad(c: int, dtype: int) -> int:
    a = chan_align(dtype)
    return (c + a - 1) // a * a


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# ---------------------------------------------------------------------------
# optional per-launch timing (CUDA events on the launching stream); used by bench.py for the
# roofline of the dominant kernel.  Off by default: zero overhead on the product path.
# ---------------------------------------------------------------------------
_PROFILE = None


def start_profile():
    global _PROFILE
    _PROFILE = []
    return _PROFILE


def stop_profile():
    """Returns [{family, flops, bytes, ms}] for every library call since start_profile()."""
    global _PROFILE
    recs, _PROFILE = _PROFILE or [], None
    torch.cuda.synchronize()
    return [dict(family=f, flops=fl, bytes=by, ms=s.elapsed_time(e)) for f, fl, by, s, e in recs]


class _Prof:
    __slots__ = ("family", "flops", "bytes", "s")

    def __init__(self, family, flops=0.0, nbytes=0.0):
        self.family, self.flops, self.bytes = family, flops, nbytes

    def __enter__(self):
        if _PROFILE is not None:
            self.s = torch.cuda.Event(enable_timing=True)
            self.s.record()
        return self

    def __exit__(self, *exc):
        if _PROFILE is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            _PROFILE.append((self.family, self.flops, self.bytes, self.s, e))
        return False


def _conv_flops(kind, n, ih, iw, k, cout, wrows):
    """Algorithmic FLOPs (2*MAC) of one vk_conv_igemm call."""
    if kind == VK_CONV3X3_S1:
        return 2.0 * n * ih * iw * 9 * k * cout
    if kind in (VK_CONV3X3_S2, VK_CONV3X3_S2_DGRAD):
        # S2: ih,iw = fine grid -> coarse output; S2_DGRAD: ih,iw = coarse grid; 9 taps per coarse pixel either way
        pix = ((ih + 1) // 2) * ((iw + 1) // 2) if kind == VK_CONV3X3_S2 else ih * iw
        return 2.0 * n * pix * 9 * k * cout
    if kind == VK_CONVT2X2_S2:
        return 2.0 * n * ih * iw * 4 * k * cout
    if kind == VK_CONV2X2_S2:
        return 2.0 * n * (ih // 2) * (iw // 2) * 4 * k * cout
    return 2.0 * n * ih * iw * k * cout


# ---------------------------------------------------------------------------
# layout helpers (torch ops; used for parameter packing and in tests)
# ---------------------------------------------------------------------------
def to_nhwc(x_nchw: torch.Tensor, dtype: int, ld: int | None = None) -> torch.Tensor:
    n, c, h, w = x_nchw.shape
    ld = ld or chan_pad(c, dtype)
    out = torch.zeros(n, h, w, ld, device=x_nchw.device, dtype=TORCH_DTYPE[dtype])
    out[..., :c] = x_nchw.permute(0, 2, 3, 1)
    return out


def from_nhwc(x_nhwc: torch.Tensor, c: int) -> torch.Tensor:
    return x_nhwc[..., :c].permute(0, 3, 1, 2).float().contiguous()


def pack_conv_weight(w: torch.Tensor, dtype: int, ldx: int | None = None, wrows: int | None = None):
    """OIHW [Cout,Cin,kh,kw] -> K-major [taps][wrows][ldx] (tap = r*kw + s)."""
    co, ci, kh, kw = w.shape
    ldx = ldx or chan_pad(ci, dtype)
    wrows = wrows or (co + 15) // 16 * 16
    out = torch.zeros(kh * kw, wrows, ldx, device=w.device, dtype=TORCH_DTYPE[dtype])
    out[:, :co, :ci] = w.permute(2, 3, 0, 1).reshape(kh * kw, co, ci)
    return out


def pack_conv_weight_dgrad(w: torch.Tensor, dtype: int, ldx: int | None = None, wrows: int | None = None):
    """Weights of the input-gradient conv of a 3x3 s1 p1 conv: rotate 180 deg, swap Cin/Cout."""
    wt = w.flip(2, 3).permute(1, 0, 2, 3).contiguous()
    return pack_conv_weight(wt, dtype, ldx, wrows)


def pack_convT_weight(w: torch.Tensor, dtype: int, ldx: int | None = None):
    """ConvTranspose2d weight [Cin,Cout,2,2] -> [1][4*Cout][ldx], row = (dy*2+dx)*Cout + co."""
    ci, co, kh, kw = w.shape
    assert kh == 2 and kw == 2 and co % 16 == 0
    ldx = ldx or chan_pad(ci, dtype)
    out = torch.zeros(1, 4 * co, ldx, device=w.device, dtype=TORCH_DTYPE[dtype])
    out[0, :, :ci] = w.permute(2, 3, 1, 0).reshape(4 * co, ci)
    return out


# ---------------------------------------------------------------------------
# vk_conv_igemm
# ---------------------------------------------------------------------------
def conv_igemm(x, w_packed, *, dtype, kind, cout, bias=None, ldo=0, epi=VK_EPI_STD, resid=None, mask=None,
               out1=None, out2=None, alpha=0.2, round_out2=False, act_expclamp=False, clamp=(0.0, 0.0),
               crop=(0, 0), out_hw=(0, 0), tune=None, cin=None, sft=None):
    """x: NHWC [n,ih,iw,ldx] tensor of the storage dtype; w_packed: [taps][wrows][ldx]."""
    assert x.is_cuda and x.is_contiguous() and w_packed.is_contiguous()
    assert x.dtype == TORCH_DTYPE[dtype] and w_packed.dtype == TORCH_DTYPE[dtype]
    n, ih, iw, ldx = x.shape
    assert w_packed.shape[2] == ldx, (w_packed.shape, ldx)
    a = _l.vk_conv_args()
    a.dtype, a.kind = dtype, kind
    a.x, a.n, a.ih, a.iw, a.ldx = _ptr(x), n, ih, iw, ldx
    a.w, a.wrows = _ptr(w_packed), w_packed.shape[1]
    a.bias = _ptr(bias)
    a.cout, a.ldo, a.epi = cout, ldo, epi
    a.resid, a.mask, a.out1, a.out2 = _ptr(resid), _ptr(mask), _ptr(out1), _ptr(out2)
    a.alpha = alpha
    a.round_out2 = int(round_out2)
    a.act_expclamp = int(act_expclamp)
    a.clamp_lo, a.clamp_hi = clamp
    a.crop_h, a.crop_w = crop
    a.out_h, a.out_w = out_hw
    if sft is not None:
        mul, add = sft                                   # fp32 [n, C] per-sample modulation of out2
        assert mul.dtype == torch.float32 and add.dtype == torch.float32 and mul.shape == add.shape
        assert mul.is_contiguous() and add.is_contiguous() and mul.shape[0] == n and mul.shape[1] >= cout
        a.sft_mul, a.sft_add, a.sft_ld = _ptr(mul), _ptr(add), mul.shape[1]
    if tune:
        a.force_tiles_per_cta = tune.get("p", 0)
        a.force_chunk_bytes = tune.get("chunk", 0)
        a.force_stages = tune.get("stages", 0)
        a.force_tw = tune.get("tw", 0)
        a.force_impl = tune.get("impl", 0)
        a.force_nt = tune.get("nt", 0)
        a.cta_timing = _ptr(tune.get("cta_timing"))
    with _Prof("conv_igemm", _conv_flops(kind, n, ih, iw, cin or ldx, cout, a.wrows) if _PROFILE is not None else 0.0):
        _l.check(_l.load().vk_conv_igemm(C.byref(a), _stream()), "vk_conv_igemm")


# ---------------------------------------------------------------------------
# vk_conv_wgrad
# ---------------------------------------------------------------------------
def _wgrad_args(a, b, dw, dtype, kind, m_valid, n_valid, dbias, tune, swapped=False):
    g = _l.vk_wgrad_args()
    g.dtype, g.kind = dtype, kind
    g.swapped = int(bool(swapped))
    assert not (swapped and dbias is not None), "swapped operands: the kernel cannot produce the bias gradient"
    g.a, g.n, g.gh, g.gw, g.lda, g.m_valid = _ptr(a), a.shape[0], a.shape[1], a.shape[2], a.shape[3], m_valid
    g.b, g.bh, g.bw, g.ldb, g.n_valid = _ptr(b), b.shape[1], b.shape[2], b.shape[3], n_valid
    g.dw, g.dbias = _ptr(dw), _ptr(dbias)
    if tune:
        g.force_ksplit = tune.get("ksplit", 0)
        g.force_k_rows = tune.get("k_rows", 0)
        g.force_stages = tune.get("stages", 0)
    return g


WGRAD_MAX_SLICES = 148      # a split-K wave never has more K slices than SMs


def conv_wgrad_plan(a, b, *, dtype, kind, m_valid, n_valid, dbias=None, tune=None, max_slices=WGRAD_MAX_SLICES,
                    swapped=False):
    """(K slices, bias slots) vk_conv_wgrad will write in the deterministic layout for these operands."""
    g = _wgrad_args(a, b, None, dtype, kind, m_valid, n_valid, dbias, tune, swapped)
    g.max_slices = max_slices
    sl, bs = C.c_int32(0), C.c_int32(0)
    _l.check(_l.load().vk_conv_wgrad_plan(C.byref(g), C.byref(sl), C.byref(bs)), "vk_conv_wgrad_plan")
    return sl.value, bs.value


def conv_wgrad(a, b, dw, *, dtype, kind, m_valid, n_valid, dbias=None, tune=None, partials=None, dbias_partials=None,
               swapped=False):
    """a: M operand NHWC (conv: dY; convT: X); b: N operand NHWC (conv: X; convT: dY_up).
    dw: fp32 [taps, m_valid, n_valid] workspace, accumulated into with red.add — or, deterministic form, `partials`
    fp32 [slices, taps, m_valid, n_valid] (+ `dbias_partials` [bias slots, m_valid]) written with plain stores, one slab
    per K slice (sizes from conv_wgrad_plan); wgrad_unpack_batched sums the slabs in order."""
    assert a.is_contiguous() and b.is_contiguous() and dw.is_contiguous() and dw.dtype == torch.float32
    assert a.dtype == TORCH_DTYPE[dtype] and b.dtype == TORCH_DTYPE[dtype]
    g = _wgrad_args(a, b, dw, dtype, kind, m_valid, n_valid, dbias, tune, swapped)
    if partials is not None:
        assert partials.dtype == torch.float32 and partials.is_contiguous() and partials.shape[1:] == dw.shape
        assert dbias is None or (dbias_partials is not None and dbias_partials.dtype == torch.float32)
        g.max_slices, g.partials, g.dbias_partials = partials.shape[0], _ptr(partials), _ptr(dbias_partials)
    taps = dw.shape[0]
    with _Prof("conv_wgrad", 2.0 * g.n * g.gh * g.gw * taps * m_valid * n_valid):
        _l.check(_l.load().vk_conv_wgrad(C.byref(g), _stream()), "vk_conv_wgrad")


def wgrad_unpack(ws, out, accumulate=False):
    """ws [taps, M, N] fp32 -> out [M, N, kh, kw] (parameter layout)."""
    taps, m, n = ws.shape
    assert out.numel() == ws.numel() and out.is_contiguous() and out.dtype == torch.float32
    with _Prof("wgrad_unpack", 0.0, 8.0 * ws.numel()):
        _l.check(_l.load().vk_wgrad_unpack(_ptr(ws), _ptr(out), taps, m, n, int(accumulate), _stream()), "vk_wgrad_unpack")


def wgrad_unpack_batched(descs_dev, ndesc, max_mn, accumulate=False):
    """All layers' [taps, M, N] workspaces -> parameter-layout gradients in one launch."""
    with _Prof("wgrad_unpack"):
        _l.check(_l.load().vk_wgrad_unpack_batched(_ptr(descs_dev), ndesc, max_mn, int(accumulate), _stream()),
                 "vk_wgrad_unpack_batched")


# ---------------------------------------------------------------------------
# HBM-bound kernels
# ---------------------------------------------------------------------------
def pack_input(img, out, *, dtype, sf=1, extra=None, extra_is_map=True, extra_sqrt_mask=0, esf=1):
    """img NCHW fp32 -> out NHWC [n, hp, wp, ld] (reflect pad, nearest upsample, optional extra channels)."""
    n, c, h, w = img.shape
    _, hp, wp, ld = out.shape
    assert img.dtype == torch.float32 and img.is_contiguous() and out.is_contiguous()
    e = eh = ew = 0
    if extra is not None:
        assert extra.dtype == torch.float32 and extra.is_contiguous()
        e = extra.shape[1]
        if extra_is_map:
            eh, ew = extra.shape[2], extra.shape[3]
    with _Prof("pack_input"):
        _l.check(_l.load().vk_pack_input(dtype, _ptr(img), n, c, h, w, sf, _ptr(extra), e, int(extra_is_map),
                                         extra_sqrt_mask, eh, ew, esf, _ptr(out), hp, wp, ld, _stream()), "vk_pack_input")


def pack_grad(g, out, *, dtype):
    n, c, h, w = g.shape
    _, hp, wp, ld = out.shape
    assert g.dtype == torch.float32 and g.is_contiguous() and out.is_contiguous()
    with _Prof("pack_grad"):
        _l.check(_l.load().vk_pack_grad(dtype, _ptr(g), n, c, h, w, _ptr(out), hp, wp, ld, _stream()), "vk_pack_grad")


def sigma_head_bwd(sigma, g_sigma, g_in, chan, out, *, dtype, log_lo, log_hi):
    n, sc, h, w = sigma.shape
    ld = out.shape[-1]
    hp = wp = ld_in = 0
    if g_in is not None:
        _, hp, wp, ld_in = g_in.shape
    assert g_sigma is None or (g_sigma.is_contiguous() and g_sigma.dtype == torch.float32)
    with _Prof("sigma_head_bwd"):
        _l.check(_l.load().vk_sigma_head_bwd(dtype, _ptr(sigma), _ptr(g_sigma), _ptr(g_in), ld_in, chan, hp, wp, _ptr(out),
                                             ld, n, sc, h, w, log_lo, log_hi, _stream()), "vk_sigma_head_bwd")


REDUCE_MAX_BLOCKS = 592     # VK_REDUCE_MAX_BLOCKS (include/virnet_b200.h): slots of the atomic-free two-pass reductions


def elbo_ws(device):
    """Scratch of vk_elbo_denoise: one (lh, kl_gauss, kl_Igamma) slot per thread block."""
    return torch.empty(3 * REDUCE_MAX_BLOCKS, device=device, dtype=torch.float64)


def adam_ws(ngroups, device):
    """Scratch of vk_adam_clip_step(_dev): one squared-norm slot per (parameter group, thread block)."""
    return torch.empty(ngroups * REDUCE_MAX_BLOCKS, device=device, dtype=torch.float64)


def channel_sum_ws(c, device):
    """Scratch of the atomic-free vk_channel_sum for up to `c` channels."""
    return torch.empty(REDUCE_MAX_BLOCKS * c, device=device, dtype=torch.float32)


def elbo_denoise(mu, sigma, noisy, gt, beta0, *, eps2, alpha0, digamma_am1, beta0_scale=1.0, grad_scale=1.0, d_mu=None, d_sigma=None,
                 acc3=None, out4=None):
    n, c, h, w = mu.shape
    sc = sigma.shape[1]
    for t in (mu, sigma, noisy, gt, beta0):
        assert t.dtype == torch.float32 and t.is_contiguous()
    # the kernel indexes the prior with sigma's channel layout (the reference broadcasts beta0 against beta,
    # loss/ELBO_simple.py:38-40; callers expand a 1-channel prior before coming here)
    assert beta0.shape == sigma.shape, f"beta0 {tuple(beta0.shape)} must have sigma's shape {tuple(sigma.shape)}"
    assert noisy.shape == mu.shape and gt.shape == mu.shape and sigma.shape[0] == n and sigma.shape[2:] == (h, w)
    if acc3 is None:
        acc3 = elbo_ws(mu.device)
    assert acc3.dtype == torch.float64 and acc3.is_contiguous()
    if out4 is None:
        out4 = torch.empty(4, device=mu.device, dtype=torch.float32)
    with _Prof("elbo_denoise"):
        _l.check(_l.load().vk_elbo_denoise(_ptr(mu), _ptr(sigma), _ptr(noisy), _ptr(gt), _ptr(beta0), beta0_scale, n, c, sc, h, w,
                                           eps2, alpha0, digamma_am1, grad_scale, _ptr(d_mu), _ptr(d_sigma), _ptr(acc3),
                                           acc3.numel(), _ptr(out4), _stream()), "vk_elbo_denoise")
    return out4


def pack_weights(descs_dev, ndesc, max_elems, *, dtype, round_tf32=True):
    with _Prof("pack_weights"):
        _l.check(_l.load().vk_pack_weights(dtype, _ptr(descs_dev), ndesc, max_elems, int(round_tf32), _stream()),
                 "vk_pack_weights")


def channel_sum(x, c, out, *, dtype, ws=None):
    """out[c] += per-channel sums of x.  ws (channel_sum_ws): per-block partials summed in block order, bit-reproducible;
    without it one atomicAdd per (block, channel)."""
    npix = x.numel() // x.shape[-1]
    assert ws is None or (ws.dtype == torch.float32 and ws.is_contiguous())
    with _Prof("channel_sum"):
        _l.check(_l.load().vk_channel_sum(dtype, _ptr(x), npix, x.shape[-1], c, _ptr(out), _ptr(ws),
                                          0 if ws is None else ws.numel(), _stream()), "vk_channel_sum")


def channel_sum_batched(x, c, out, *, dtype):
    """x [n, ..., ld] -> out [n, c] += per-sample sums over the pixels."""
    n = x.shape[0]
    npix = x.numel() // (x.shape[-1] * n)
    assert x.is_contiguous() and out.is_contiguous() and out.shape == (n, c) and out.dtype == torch.float32
    with _Prof("channel_sum"):
        _l.check(_l.load().vk_channel_sum_batched(dtype, _ptr(x), n, npix, x.shape[-1], c, _ptr(out), _stream()),
                 "vk_channel_sum_batched")


def adam_clip_step(params, grads, exp_avg, exp_avg_sq, groups_dev, ngroups, max_group_elems, sq_ws, *, grad_scale, lr,
                   beta1, beta2, eps, step, norms_out=None):
    with _Prof("adam_clip"):
        _l.check(_l.load().vk_adam_clip_step(_ptr(params), _ptr(grads), _ptr(exp_avg), _ptr(exp_avg_sq), _ptr(groups_dev),
                                             ngroups, max_group_elems, _ptr(sq_ws), sq_ws.numel(), grad_scale, lr, beta1,
                                             beta2, eps, step, _ptr(norms_out), _stream()), "vk_adam_clip_step")


def adam_clip_step_dev(params, grads, exp_avg, exp_avg_sq, groups_dev, ngroups, max_group_elems, sq_ws, hyper_dev, *,
                       grad_scale, beta1, beta2, eps, norms_out=None):
    """hyper_dev: fp32 [3] on the device = (lr, 1 - beta1^step, sqrt(1 - beta2^step))."""
    with _Prof("adam_clip"):
        _l.check(_l.load().vk_adam_clip_step_dev(_ptr(params), _ptr(grads), _ptr(exp_avg), _ptr(exp_avg_sq),
                                                 _ptr(groups_dev), ngroups, max_group_elems, _ptr(sq_ws), sq_ws.numel(),
                                                 grad_scale, beta1, beta2, eps, _ptr(hyper_dev), _ptr(norms_out), _stream()),
                 "vk_adam_clip_step_dev")


# ---------------------------------------------------------------------------
# super-resolution forward path: small per-sample kernels
# ---------------------------------------------------------------------------
def knet_head(x, w, out, *, dtype):
    """x NCHW fp32, w [cout, c, 9, 9] fp32 -> out NHWC [n, oh, ow, ld] (Conv2d k9 s4 p4, no bias)."""
    n, c, h, wd = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous() and w.is_contiguous() and out.is_contiguous()
    assert out.shape[1] == (h - 1) // 4 + 1 and out.shape[2] == (wd - 1) // 4 + 1
    with _Prof("knet_head"):
        _l.check(_l.load().vk_knet_head(dtype, _ptr(x), _ptr(w), _ptr(out), n, c, h, wd, w.shape[0], out.shape[-1],
                                        _stream()), "vk_knet_head")


def ca_layer(f, skip, w1, b1, w2, b2, out, *, dtype, c, alpha=0.2):
    """out = f * sigmoid(W2 lrelu(W1 mean(f) + b1) + b2) + skip; f/skip/out NHWC [n, h, w, ld]."""
    n, ld = f.shape[0], f.shape[-1]
    npix = f.shape[1] * f.shape[2]
    for t in (f, skip, out):
        assert t.is_contiguous() and t.shape == f.shape
    with _Prof("ca_layer"):
        _l.check(_l.load().vk_ca_layer(dtype, _ptr(f), _ptr(skip), _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), _ptr(out), n,
                                       npix, c, w1.shape[0], ld, alpha, _stream()), "vk_ca_layer")


def gap_head(x, out, *, exp_mask=0, tanh_mask=0, lo=0.0, hi=0.0):
    """x NCHW fp32 -> out [n, c] = head(mean over pixels)."""
    n, c = x.shape[0], x.shape[1]
    assert x.dtype == torch.float32 and x.is_contiguous() and out.is_contiguous() and out.numel() == n * c
    with _Prof("gap_head"):
        _l.check(_l.load().vk_gap_head(_ptr(x), n, c, x.shape[2] * x.shape[3], exp_mask, tanh_mask, lo, hi, _ptr(out),
                                       _stream()), "vk_gap_head")


def sft_mlp(extra, att, mul, add, *, sqrt_mask=0, alpha=0.2):
    """att: an AttLayer container (conv1, conv2, mul_conv, add_conv 1x1 convs); extra fp32 [n, e]."""
    n, e = extra.shape
    c1, c2, c = att.conv1.out_channels, att.conv2.out_channels, att.mul_conv.out_channels
    assert extra.dtype == torch.float32 and extra.is_contiguous() and mul.shape == (n, c) and add.shape == (n, c)
    with _Prof("sft_mlp"):
        _l.check(_l.load().vk_sft_mlp(_ptr(extra), n, e, sqrt_mask, _ptr(att.conv1.weight), _ptr(att.conv1.bias), c1,
                                      _ptr(att.conv2.weight), _ptr(att.conv2.bias), c2, _ptr(att.mul_conv.weight),
                                      _ptr(att.mul_conv.bias), _ptr(att.add_conv.weight), _ptr(att.add_conv.bias), c,
                                      alpha, _ptr(mul), _ptr(add), _stream()), "vk_sft_mlp")


def upsample_nearest_nchw(x, out, sf):
    n, c, h, w = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous() and out.is_contiguous() and out.shape == (n, c, h * sf, w * sf)
    with _Prof("upsample_nearest"):
        _l.check(_l.load().vk_upsample_nearest(_ptr(x), _ptr(out), n, c, h, w, sf, _stream()), "vk_upsample_nearest")


# ---------------------------------------------------------------------------
# backward of the super-resolution network's small layers
# ---------------------------------------------------------------------------
def sft_bwd_det_ws_floats(n, npix, c):
    """Floats of scratch the deterministic form of sft_bwd needs for [n, npix, *] tensors with c channels."""
    need = _l.load().vk_sft_bwd_det_ws_floats(n, npix, c)
    if need < 0:
        raise _l.VkError("vk_sft_bwd_det_ws_floats: bad shape")
    return need


def sft_bwd(g, x, mul, gx, dmul, dadd, *, dtype, c, resid=None, det_ws=None):
    """gx = g * mul (+ resid); dmul/dadd [n, c] += sums over pixels of g * x and g.  NHWC tensors [n, h, w, ld].
    det_ws (fp32 scratch of at least sft_bwd_det_ws_floats(...) elements): the bit-reproducible form without atomics."""
    n, ld = g.shape[0], g.shape[-1]
    npix = g.shape[1] * g.shape[2]
    for t in (g, x, gx) + ((resid,) if resid is not None else ()):
        assert t.is_contiguous() and t.shape == g.shape
    assert mul.shape == (n, c) and dmul.shape == (n, c) and dadd.shape == (n, c)
    if det_ws is not None:
        assert det_ws.dtype == torch.float32 and det_ws.is_cuda and det_ws.is_contiguous()
        with _Prof("sft_bwd"):
            _l.check(_l.load().vk_sft_bwd_det(dtype, _ptr(g), _ptr(x), _ptr(mul), _ptr(resid), _ptr(gx), _ptr(dmul),
                                              _ptr(dadd), n, npix, c, ld, _ptr(det_ws), det_ws.numel(), _stream()),
                     "vk_sft_bwd_det")
        return
    with _Prof("sft_bwd"):
        _l.check(_l.load().vk_sft_bwd(dtype, _ptr(g), _ptr(x), _ptr(mul), _ptr(resid), _ptr(gx), _ptr(dmul), _ptr(dadd), n,
                                      npix, c, ld, _stream()), "vk_sft_bwd")


def sft_mlp_bwd(extra, att, dmul, dadd, grads, d_extra, *, sqrt_mask=0, alpha=0.2):
    """att: AttLayer container; grads: dict parameter -> gradient view (accumulated); d_extra [n, e] accumulated."""
    n, e = extra.shape
    c1, c2, c = att.conv1.out_channels, att.conv2.out_channels, att.mul_conv.out_channels
    gp = [grads(p) for p in (att.conv1.weight, att.conv1.bias, att.conv2.weight, att.conv2.bias, att.mul_conv.weight,
                             att.mul_conv.bias, att.add_conv.weight, att.add_conv.bias)]
    with _Prof("sft_mlp_bwd"):
        _l.check(_l.load().vk_sft_mlp_bwd(
            _ptr(extra), n, e, sqrt_mask, _ptr(att.conv1.weight), _ptr(att.conv1.bias), c1, _ptr(att.conv2.weight),
            _ptr(att.conv2.bias), c2, _ptr(att.mul_conv.weight), _ptr(att.mul_conv.bias), _ptr(att.add_conv.weight),
            _ptr(att.add_conv.bias), c, alpha, _ptr(dmul), _ptr(dadd), *[_ptr(t) for t in gp], _ptr(d_extra), _stream()),
            "vk_sft_mlp_bwd")


def ca_layer_bwd(g, f, ca, df, grads, *, dtype, c, alpha=0.2, det_ws=None):
    """ca: CALayer container (body[0], body[2] 1x1 convs); df = g * s + dy / npix; parameter grads accumulated.
    det_ws (fp32 scratch, at least n * (2 r c + r + c) elements): the bit-reproducible form without atomics."""
    n, ld = g.shape[0], g.shape[-1]
    npix = g.shape[1] * g.shape[2]
    w1, b1, w2, b2 = ca.body[0].weight, ca.body[0].bias, ca.body[2].weight, ca.body[2].bias
    if det_ws is not None:
        assert det_ws.dtype == torch.float32 and det_ws.is_cuda and det_ws.is_contiguous()
        with _Prof("ca_layer_bwd"):
            _l.check(_l.load().vk_ca_layer_bwd_det(dtype, _ptr(g), _ptr(f), _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), _ptr(df),
                                                   _ptr(grads(w1)), _ptr(grads(b1)), _ptr(grads(w2)), _ptr(grads(b2)), n,
                                                   npix, c, w1.shape[0], ld, alpha, _ptr(det_ws), det_ws.numel(),
                                                   _stream()), "vk_ca_layer_bwd_det")
        return
    with _Prof("ca_layer_bwd"):
        _l.check(_l.load().vk_ca_layer_bwd(dtype, _ptr(g), _ptr(f), _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), _ptr(df),
                                           _ptr(grads(w1)), _ptr(grads(b1)), _ptr(grads(w2)), _ptr(grads(b2)), n, npix, c,
                                           w1.shape[0], ld, alpha, _stream()), "vk_ca_layer_bwd")


def gap_head_bwd(gout, outv, gx, *, dtype, c, exp_mask=0, tanh_mask=0, lo=0.0, hi=0.0):
    """gx NHWC [n, h, w, ld] = gout[n, c] * head'(outv[n, c]) / (h * w)."""
    n, ld = gx.shape[0], gx.shape[-1]
    hw = gx.shape[1] * gx.shape[2]
    assert gout.is_contiguous() and outv.is_contiguous() and gout.numel() == n * c and outv.numel() == n * c
    with _Prof("gap_head_bwd"):
        _l.check(_l.load().vk_gap_head_bwd(dtype, _ptr(gout), _ptr(outv), n, c, hw, exp_mask, tanh_mask, lo, hi, _ptr(gx),
                                           ld, _stream()), "vk_gap_head_bwd")


def knet_head_wgrad(x, g, gw, *, dtype, det_ws=None):
    """gw [cout, c, 9, 9] += weight gradient of the KNet head.  det_ws (fp32 scratch, at least n * gw.numel() elements):
    per-sample slots added in sample order instead of atomics."""
    n, c, h, wd = x.shape
    with _Prof("knet_head_wgrad"):
        if det_ws is not None:
            assert det_ws.dtype == torch.float32 and det_ws.is_cuda and det_ws.is_contiguous()
            _l.check(_l.load().vk_knet_head_wgrad_det(dtype, _ptr(x), _ptr(g), _ptr(gw), n, c, h, wd, gw.shape[0],
                                                      g.shape[-1], _ptr(det_ws), det_ws.numel(), _stream()),
                     "vk_knet_head_wgrad_det")
        else:
            _l.check(_l.load().vk_knet_head_wgrad(dtype, _ptr(x), _ptr(g), _ptr(gw), n, c, h, wd, gw.shape[0], g.shape[-1],
                                                  _stream()), "vk_knet_head_wgrad")


# ---------------------------------------------------------------------------
# super-resolution negative ELBO (forward value + gradients in one call)
# ---------------------------------------------------------------------------
def elbo_sisr(mu, im_hr, im_lr, sigma_est, kinfo_est, kinfo_gt, prior_mean, prior_logmean, gamma_draw, rho_draw,
              z_draw, rh, rw, *, k_size, center, alpha0, digamma_am1, kappa0, r2, eps2, pk0, pk1, ws_cache=None):
    """All tensors fp32 contiguous CUDA.  Returns (terms[8], kernel [n,1,k,k], d_mu, d_sigma [n], d_kinfo [n,3]).

    Scratch memory belongs to the CALLER: `ws_cache` is a dict the caller keeps alive (one per trainer); buffers put in
    it are never freed here, so a CUDA graph that captured their addresses stays valid whatever other shapes are
    evaluated later.  Without a cache every call allocates its own scratch buffer from torch's caching allocator
    (stream-ordered, and from the graph's private pool during a capture)."""
    n, c, H, W = mu.shape
    h, w = im_lr.shape[2], im_lr.shape[3]
    for t in (mu, im_hr, im_lr, sigma_est, kinfo_est, kinfo_gt, prior_mean, prior_logmean, gamma_draw, rho_draw,
              z_draw, rh, rw):
        assert t.dtype == torch.float32 and t.is_cuda and t.is_contiguous()
    assert im_hr.shape == mu.shape and z_draw.shape == mu.shape and im_lr.shape[:2] == (n, c)
    assert rh.shape == (h, H) and rw.shape == (w, W), "down-sampling operators do not match the LR / HR sizes"
    assert sigma_est.numel() == n and kinfo_est.shape == (n, 3) and kinfo_gt.shape == (n, 3)
    assert gamma_draw.shape == (n, 2) and rho_draw.numel() == n and prior_mean.numel() == n
    lib = _l.load()
    need = lib.vk_elbo_sisr_ws_bytes(n, c, H, W, h, w, k_size)
    if need < 0:
        raise _l.VkError("vk_elbo_sisr_ws_bytes: bad shape")
    if ws_cache is None:
        ws = torch.empty(need, dtype=torch.uint8, device=mu.device)
    else:
        key = (mu.device, need, torch.cuda.current_stream(mu.device).cuda_stream)
        ws = ws_cache.get(key)
        if ws is None:
            ws = ws_cache[key] = torch.empty(need, dtype=torch.uint8, device=mu.device)
    dev = mu.device
    d_mu = torch.empty_like(mu)
    d_sigma = torch.empty(n, device=dev, dtype=torch.float32)
    d_kinfo = torch.empty(n, 3, device=dev, dtype=torch.float32)
    kernel = torch.empty(n, 1, k_size, k_size, device=dev, dtype=torch.float32)
    terms = torch.empty(8, device=dev, dtype=torch.float32)
    a = _l.vk_elbo_sisr_args()
    a.mu, a.im_hr, a.im_lr, a.sigma_est = _ptr(mu), _ptr(im_hr), _ptr(im_lr), _ptr(sigma_est)
    a.kinfo_est, a.kinfo_gt, a.prior_mean, a.prior_logmean = _ptr(kinfo_est), _ptr(kinfo_gt), _ptr(prior_mean), _ptr(prior_logmean)
    a.gamma_draw, a.rho_draw, a.z_draw, a.rh, a.rw = _ptr(gamma_draw), _ptr(rho_draw), _ptr(z_draw), _ptr(rh), _ptr(rw)
    a.d_mu, a.d_sigma, a.d_kinfo, a.kernel, a.terms = _ptr(d_mu), _ptr(d_sigma), _ptr(d_kinfo), _ptr(kernel), _ptr(terms)
    a.ws, a.ws_bytes = _ptr(ws), need
    a.n, a.c, a.H, a.W, a.h, a.w, a.k_size = n, c, H, W, h, w, k_size
    a.center, a.alpha0, a.digamma_am1 = center, alpha0, digamma_am1
    a.kappa0, a.r2, a.eps2, a.pk0, a.pk1 = kappa0, r2, eps2, pk0, pk1
    with _Prof("elbo_sisr"):
        _l.check(lib.vk_elbo_sisr(C.byref(a), _stream()), "vk_elbo_sisr")
    return terms, kernel, d_mu, d_sigma, d_kinfo


# ---------------------------------------------------------------------------
# real-noise denoising trainer: variance-map prior and MixUp
# ---------------------------------------------------------------------------
def noise_estimate(im_noisy, im_gt, window, out=None, floor=1e-10):
    """out = clamp_min(window (*) (im_noisy - im_gt)^2, floor), reflect padding; NCHW fp32, window [k, k] fp32."""
    n, c, h, w = im_noisy.shape
    for t in (im_noisy, im_gt, window):
        assert t.dtype == torch.float32 and t.is_cuda and t.is_contiguous()
    assert im_gt.shape == im_noisy.shape and window.shape[0] == window.shape[1]
    if out is None:
        out = torch.empty_like(im_noisy)
    with _Prof("noise_estimate"):
        _l.check(_l.load().vk_noise_estimate(_ptr(im_noisy), _ptr(im_gt), _ptr(window), window.shape[0], _ptr(out), n * c,
                                             h, w, floor, _stream()), "vk_noise_estimate")
    return out


def mixup(a, b, perm, lam):
    """(lam * a + (1 - lam) * a[perm], same for b); a, b [n, ...] fp32, perm int64 [n], lam fp32 [n] (all CUDA)."""
    n = a.shape[0]
    for t in (a, b, lam):
        assert t.dtype == torch.float32 and t.is_cuda and t.is_contiguous()
    assert a.shape == b.shape and perm.dtype == torch.int64 and perm.is_cuda and perm.numel() == n and lam.numel() == n
    out_a, out_b = torch.empty_like(a), torch.empty_like(b)
    with _Prof("mixup"):
        _l.check(_l.load().vk_mixup(_ptr(a), _ptr(b), _ptr(perm), _ptr(lam), _ptr(out_a), _ptr(out_b), n, a.numel() // n,
                                    _stream()), "vk_mixup")
    return out_a, out_b


def synth_denoise(patches_u8, params, aug, noise, *, clip=False):
    """patches uint8 [n, p, p, c]; params fp64 [n, 6]; aug int32 [n]; noise fp32 [n, p, p, c] (all CUDA, contiguous).
    Returns (im_noisy, im_gt, sigma_gt) NCHW fp32."""
    n, p, p2, c = patches_u8.shape
    assert p == p2 and patches_u8.dtype == torch.uint8 and params.dtype == torch.float64 and params.shape == (n, 6)
    assert aug.dtype == torch.int32 and aug.numel() == n and noise.dtype == torch.float32 and noise.shape == patches_u8.shape
    for t in (patches_u8, params, aug, noise):
        assert t.is_cuda and t.is_contiguous()
    dev = patches_u8.device
    im_noisy = torch.empty(n, c, p, p, device=dev, dtype=torch.float32)
    im_gt = torch.empty_like(im_noisy)
    sigma_gt = torch.empty(n, 1, p, p, device=dev, dtype=torch.float32)
    with _Prof("synth_denoise"):
        _l.check(_l.load().vk_synth_denoise(_ptr(patches_u8), _ptr(params), _ptr(aug), _ptr(noise), n, p, c, int(clip),
                                            _ptr(im_noisy), _ptr(im_gt), _ptr(sigma_gt), _stream()), "vk_synth_denoise")
    return im_noisy, im_gt, sigma_gt


def sft_mlp_batched(descs_dev, n_layers, max_c, extra, *, sqrt_mask=0, alpha=0.2):
    n, e = extra.shape
    with _Prof("sft_mlp"):
        _l.check(_l.load().vk_sft_mlp_batched(_ptr(descs_dev), n_layers, max_c, _ptr(extra), n, e, sqrt_mask, alpha,
                                              _stream()), "vk_sft_mlp_batched")


def sft_mlp_bwd_batched(descs_dev, n_layers, max_c, extra, d_extra, *, sqrt_mask=0, alpha=0.2, det_ws=None,
                        params_per_sample=0):
    """det_ws (fp32 scratch, at least n * params_per_sample + n_layers * n * e elements; params_per_sample = parameters of
    all the AttLayers): per-(sample, layer) slots added in a fixed order instead of racing atomics."""
    n, e = extra.shape
    with _Prof("sft_mlp_bwd"):
        if det_ws is not None:
            assert det_ws.dtype == torch.float32 and det_ws.is_cuda and det_ws.is_contiguous() and params_per_sample > 0
            _l.check(_l.load().vk_sft_mlp_bwd_batched_det(_ptr(descs_dev), n_layers, max_c, _ptr(extra), n, e, sqrt_mask,
                                                          alpha, _ptr(d_extra), params_per_sample, _ptr(det_ws),
                                                          det_ws.numel(), _stream()), "vk_sft_mlp_bwd_batched_det")
        else:
            _l.check(_l.load().vk_sft_mlp_bwd_batched(_ptr(descs_dev), n_layers, max_c, _ptr(extra), n, e, sqrt_mask, alpha,
                                                      _ptr(d_extra), _stream()), "vk_sft_mlp_bwd_batched")


def sisr_degrade(im_hr, kernels, rh, rw, noise, std):
    """im_hr [n, c, H, W], kernels [n, k, k], rh [h, H], rw [w, W], noise [n, c, h, w], std [n] (fp32 CUDA contiguous).
    Returns (im_blur, im_lr) [n, c, h, w]."""
    n, c, H, W = im_hr.shape
    h, w, k = rh.shape[0], rw.shape[0], kernels.shape[-1]
    for t in (im_hr, kernels, rh, rw, noise, std):
        assert t.dtype == torch.float32 and t.is_cuda and t.is_contiguous()
    assert kernels.shape == (n, k, k) and rh.shape == (h, H) and rw.shape == (w, W) and noise.shape == (n, c, h, w)
    assert std.numel() == n
    lib = _l.load()
    need = lib.vk_sisr_degrade_ws_bytes(n, c, H, W, w)
    ws = torch.empty(need, dtype=torch.uint8, device=im_hr.device)
    im_blur = torch.empty(n, c, h, w, device=im_hr.device, dtype=torch.float32)
    im_lr = torch.empty_like(im_blur)
    with _Prof("sisr_degrade"):
        _l.check(lib.vk_sisr_degrade(_ptr(im_hr), _ptr(kernels), k, _ptr(rh), _ptr(rw), _ptr(noise), _ptr(std),
                                     _ptr(im_blur), _ptr(im_lr), _ptr(ws), need, n, c, H, W, h, w, _stream()),
                 "vk_sisr_degrade")
    return im_blur, im_lr


# ---------------------------------------------------------------------------
# spatially varying conditioning maps (csrc/vk_sft_spatial.cu)
# ---------------------------------------------------------------------------
class ExtraSource:
    """Where the conditioning channels of a pixel come from: per-sample constants `cst` [n, ec] and / or maps `map`
    [n, em, eh, ew] (nearest-upsampled by esf to hh x ww, reflect padded to hp x wp).  Keeps the tensors alive."""

    def __init__(self, cst, map_, esf, sqrt_mask, hh, ww, hp, wp):
        for t in (cst, map_):
            assert t is None or (t.dtype == torch.float32 and t.is_cuda and t.is_contiguous())
        self.cst, self.map = cst, map_
        s = _l.vk_extra_src()
        s.cst, s.map = _ptr(cst), _ptr(map_)
        s.ec = 0 if cst is None else cst.shape[1]
        s.em, s.eh, s.ew = (0, 0, 0) if map_ is None else (map_.shape[1], map_.shape[2], map_.shape[3])
        s.esf, s.sqrt_mask, s.hh, s.ww, s.hp, s.wp = int(esf), int(sqrt_mask), hh, ww, hp, wp
        self.c = s
        self.e = s.ec + s.em


def pack_input_mixed(img, out, src: ExtraSource, *, dtype, sf=1):
    n, c, h, w = img.shape
    assert img.dtype == torch.float32 and img.is_contiguous() and out.is_contiguous()
    assert out.shape[:3] == (n, src.c.hp, src.c.wp)
    with _Prof("pack_input"):
        _l.check(_l.load().vk_pack_input_mixed(dtype, _ptr(img), n, c, h, w, sf, C.byref(src.c), _ptr(out), out.shape[-1],
                                               _stream()), "vk_pack_input_mixed")


def sft_apply(x, out, att, src: ExtraSource, *, dtype, c, alpha=0.2, round_tf32=False):
    """out = lrelu(x * mul + add), (mul, add) = AttLayer `att` evaluated per pixel on the conditioning source."""
    n, h, w, ld = x.shape
    assert x.is_contiguous() and out.is_contiguous() and out.shape == x.shape
    a = _l.vk_sft_apply_args()
    a.dtype, a.n, a.h, a.w, a.c, a.ld = dtype, n, h, w, c, ld
    a.c1, a.c2 = att.conv1.out_channels, att.conv2.out_channels
    a.x, a.out = _ptr(x), _ptr(out)
    for nm, p_ in (("w1", att.conv1.weight), ("b1", att.conv1.bias), ("w2", att.conv2.weight), ("b2", att.conv2.bias),
                   ("wm", att.mul_conv.weight), ("bm", att.mul_conv.bias), ("wa", att.add_conv.weight),
                   ("ba", att.add_conv.bias)):
        setattr(a, nm, p_.data_ptr())
    a.extra = src.c
    a.alpha, a.round_tf32 = alpha, int(round_tf32)
    with _Prof("sft_apply"):
        _l.check(_l.load().vk_sft_apply(C.byref(a), _stream()), "vk_sft_apply")


def sft_apply_bwd(g, x, gx, att, src: ExtraSource, grad_view, *, dtype, c, resid=None, d_cst=None, d_map=None, alpha=0.2):
    """Backward of sft_apply.  g = dL/d(x*mul+add) (lrelu' already applied by the producing dgrad), x = the modulated
    features; writes gx = g * mul (+ resid), accumulates the AttLayer's parameter gradients into grad_view(param) (four
    pixel-K tensor-core GEMMs over per-pixel operands) and the conditioning gradients into d_cst / d_map."""
    n, h, w, ld = x.shape
    assert g.shape == x.shape and gx.shape == x.shape and g.is_contiguous() and x.is_contiguous() and gx.is_contiguous()
    tdt = TORCH_DTYPE[dtype]
    c1, c2 = att.conv1.out_channels, att.conv2.out_channels
    ld1, ld2 = (c1 + 15) // 16 * 16, (c2 + 15) // 16 * 16
    dev = x.device
    dm = torch.empty_like(x)
    f2, dq2 = torch.empty(n, h, w, ld2, device=dev, dtype=tdt), torch.empty(n, h, w, ld2, device=dev, dtype=tdt)
    f1, dq1 = torch.empty(n, h, w, ld1, device=dev, dtype=tdt), torch.empty(n, h, w, ld1, device=dev, dtype=tdt)
    ev = torch.empty(n, h, w, 16, device=dev, dtype=tdt)
    a = _l.vk_sft_apply_bwd_args()
    a.dtype, a.n, a.h, a.w, a.c, a.ld = dtype, n, h, w, c, ld
    a.c1, a.c2, a.ld1, a.ld2 = c1, c2, ld1, ld2
    a.g, a.x, a.resid, a.gx = _ptr(g), _ptr(x), _ptr(resid), _ptr(gx)
    a.dm, a.f2, a.dq2, a.f1, a.dq1, a.ev = _ptr(dm), _ptr(f2), _ptr(dq2), _ptr(f1), _ptr(dq1), _ptr(ev)
    for nm, p_ in (("w1", att.conv1.weight), ("b1", att.conv1.bias), ("w2", att.conv2.weight), ("b2", att.conv2.bias),
                   ("wm", att.mul_conv.weight), ("bm", att.mul_conv.bias), ("wa", att.add_conv.weight),
                   ("ba", att.add_conv.bias)):
        setattr(a, nm, p_.data_ptr())
    a.d_cst, a.d_map = _ptr(d_cst), _ptr(d_map)
    a.extra = src.c
    a.alpha = alpha
    with _Prof("sft_apply_bwd"):
        _l.check(_l.load().vk_sft_apply_bwd(C.byref(a), _stream()), "vk_sft_apply_bwd")
    e = src.e

    def gemm(m_op, n_op, conv, m_valid, n_valid):
        conv_wgrad(m_op, n_op, grad_view(conv.weight).view(1, m_valid, n_valid), dtype=dtype, kind=VK_CONV1X1,
                   m_valid=m_valid, n_valid=n_valid, dbias=grad_view(conv.bias))

    gemm(dm, f2, att.mul_conv, c, c2)
    gemm(g, f2, att.add_conv, c, c2)
    gemm(dq2, f1, att.conv2, c2, c1)
    gemm(dq1, ev, att.conv1, c1, e)


def extra_head_grad(g_r0, c, src: ExtraSource, *, dtype, d_cst=None, d_map=None):
    """Conditioning-channel gradient of the head conv's packed input g_r0 [n, hp, wp, ld] -> d_cst / d_map (accumulated)."""
    n, hp, wp, ld = g_r0.shape
    with _Prof("extra_head_grad"):
        _l.check(_l.load().vk_extra_head_grad(dtype, _ptr(g_r0), n, ld, c, C.byref(src.c), _ptr(d_cst), _ptr(d_map),
                                              _stream()), "vk_extra_head_grad")
