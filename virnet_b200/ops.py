"""Tensor-level wrappers over the C ABI (device pointers come from torch tensors).

PyTorch is plumbing here: it owns device memory and the CUDA stream.  All
arithmetic on the hot path happens inside libvirnet_sm100.so.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import lib as _l
from .lib import (VK_BF16, VK_TF32, VK_CONV3X3_S1, VK_CONV3X3_S2, VK_CONVT2X2_S2, VK_CONV1X1,
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
               crop=(0, 0), tune=None):
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
    if tune:
        a.force_tiles_per_cta = tune.get("p", 0)
        a.force_chunk_bytes = tune.get("chunk", 0)
        a.force_stages = tune.get("stages", 0)
        a.force_tw = tune.get("tw", 0)
    _l.check(_l.load().vk_conv_igemm(C.byref(a), _stream()), "vk_conv_igemm")
