"""Host-side construction of the down-sampling operators used by the SISR likelihood.

`conv_multi_kernel_tensor` (reference utils/util_sisr.py:127-144) down-samples the blurred image either by
taking every sf-th pixel ('Direct') or with ResizeRight's antialiased cubic resize ('Bicubic',
ResizeRight/resize_right.py:29-76).  For a given (size, scale) both are FIXED separable linear maps, so the
CUDA loss (csrc/vk_sisr_loss.cu) applies them — and their transposes in the backward pass — as two small dense
matrices.  This module builds those matrices; they are cached per (size, sf, mode, device).

The cubic weights follow ResizeRight step by step: output centres projected onto the input axis
(resize_right.py:251-262), a window of ceil(4 sf) taps starting at ceil(centre - 2 sf - eps) (:265-276),
mirror ("reflect-with-edge") folding of out-of-range taps (:278-284), the cubic kernel stretched by the scale
for antialiasing (:304-315, interp_methods.py:34-42) and per-output normalisation (:288-299).
"""
from __future__ import annotations

import math
from functools import lru_cache

import torch


def _keys_cubic(t: torch.Tensor) -> torch.Tensor:
    a = t.abs()
    a2, a3 = a * a, a * a * a
    near = (1.5 * a3 - 2.5 * a2 + 1.0) * (a <= 1.0).to(t.dtype)
    far = (-0.5 * a3 + 2.5 * a2 - 4.0 * a + 2.0) * ((a > 1.0) & (a <= 2.0)).to(t.dtype)
    return near + far


@lru_cache(maxsize=64)
def _matrix_cpu(in_sz: int, sf: int, mode: str) -> torch.Tensor:
    if mode == "direct":
        out_sz = -(-in_sz // sf)
        m = torch.zeros(out_sz, in_sz)
        m[torch.arange(out_sz), torch.arange(out_sz) * sf] = 1.0
        return m
    scale = 1.0 / sf
    out_sz = int(math.ceil(scale * in_sz))
    eps = torch.finfo(torch.float32).eps
    width = 4.0 / scale
    centres = torch.arange(out_sz) / scale + (in_sz - 1) / 2 - (out_sz - 1) / (2 * scale)
    first = (centres - width / 2 - eps).ceil().long()
    taps = first[:, None] + torch.arange(int(math.ceil(width - eps)))
    period = 2 * in_sz
    folded = torch.remainder(taps, period)
    folded = torch.where(folded < in_sz, folded, period - 1 - folded)
    wts = scale * _keys_cubic(scale * (centres[:, None] - folded))
    norm = wts.sum(1, keepdim=True)
    norm[norm == 0] = 1
    m = torch.zeros(out_sz, in_sz)
    m.scatter_add_(1, folded, wts / norm)
    return m


_DEVICE_CACHE = {}


def downsample_matrix(in_sz: int, sf: int, downsampler: str, device) -> torch.Tensor:
    """Dense [out_sz, in_sz] fp32 operator for one axis; `downsampler` is 'Direct' or 'Bicubic'."""
    mode = downsampler.lower()
    if mode not in ("direct", "bicubic"):
        raise ValueError("Please input the corrected downsampler: Direct or Bicubic!")
    key = (in_sz, int(sf), mode, str(device))
    m = _DEVICE_CACHE.get(key)
    if m is None:
        m = _matrix_cpu(in_sz, int(sf), mode).to(device).contiguous()
        _DEVICE_CACHE[key] = m
    return m
