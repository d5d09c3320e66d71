"""Execution engine of the denoising network (VIRAttResUNet): turns the parameter
containers into one CUDA program over libvirnet_sm100.so.

Forward dataflow (reference: networks/VIRNet.py:42-46, DnCNN.py:37-44, AttResUNet.py:141-175),
all activations NHWC in the compute dtype (bf16, or fp32 storage with TF32 MMA):

  pack_input(x) -> SNet convs (bias+LReLU(.25) fused) -> last conv with exp(clamp) epilogue -> sigma
  pack_input(x, sqrt(sigma)) [reflect pad + concat fused] -> head conv (dual write: X, LReLU(X))
  resblock: conv1 epilogue writes LReLU(conv1+b);  conv2 epilogue writes X' = X + conv2 + b and LReLU(X')
  stride-2 conv via TMA element strides; ConvT as 1x1 GEMM + depth-to-space epilogue fused with `+ bridge`
  tail conv epilogue: + bias, crop, + x_in, NCHW fp32 store

Backward mirrors it: dgrad = the same implicit-GEMM kernel over rotated/transposed packed weights
with the LeakyReLU' mask (taken from the sign of the saved activated tensor) and the residual
gradient fused in the epilogue; wgrad = pixel-K GEMM with split-K fp32 reduction.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from collections import OrderedDict
from typing import Dict, List, Optional

import torch
from torch import nn

from . import lib as _l
from . import ops
from .ops import (VK_BF16, VK_TF32, VK_CONV3X3_S1, VK_CONV3X3_S2, VK_CONVT2X2_S2, VK_CONV2X2_S2,
                  VK_CONV3X3_S2_DGRAD, VK_EPI_NCHW_F32)

SNET_LOG_MAX, SNET_LOG_MIN = math.log(1e2), math.log(1e-10)     # networks/VIRNet.py:15-16
KNET_LOG_MAX, KNET_LOG_MIN = math.log(1e2), math.log(1e-4)      # networks/KNet.py:5-6

PRECISIONS = {"bf16": VK_BF16, "tf32": VK_TF32}


def _pad16(c: int) -> int:
    return (c + 15) // 16 * 16


def _wrows(c: int) -> int:
    """Rows of a packed weight operand (GEMM N).  Layers with at most 16 output rows (tail, SNet end, head dgrad) are
    padded to 32: a CTA pair then issues M=256 MMAs with 16 rows from each CTA — half as many MMAs per tile, and an N=16
    MMA costs the same ~44 clocks of issue time as an N=32 one."""
    return 32 if c <= 16 else _pad16(c)


class _Layer:
    """One Conv2d / ConvTranspose2d: parameters plus their packed GEMM operands."""

    def __init__(self, name: str, mod: nn.Module, kind: str, need_dgrad: bool):
        self.name, self.mod, self.kind, self.need_dgrad = name, mod, kind, need_dgrad
        w = mod.weight
        if kind == "convT":
            self.cin, self.cout, self.taps = w.shape[0], w.shape[1], 4
        else:
            self.cout, self.cin, self.taps = w.shape[0], w.shape[1], w.shape[2] * w.shape[3]
        self.wf = self.wd = None        # packed forward / dgrad operands
        self.ws = None                  # fp32 wgrad workspace view [taps][M][N]

    @property
    def weight(self):
        return self.mod.weight

    @property
    def bias(self):
        return self.mod.bias


class _NoSave(dict):
    """Activation table of a forward that nobody will differentiate: drops what is put in, so every tensor dies with
    its last Python reference and torch's caching allocator recycles the block for a later layer."""

    def __setitem__(self, k, v):
        pass


class DenoiseEngine:
    """Buffer ownership.  A forward that saves for backward (`save=True`) writes its activations into a named buffer
    SET keyed by the input shape; sets live in a small LRU (`max_cached_shapes`, default 2) so that a validation loop
    over many image sizes cannot pin memory without bound.  A set in use by a forward whose backward has not run yet is
    never handed to another forward (a second differentiable forward of the same shape gets its own set), and each
    saved forward carries a generation number that `backward` checks.  Forwards with `save=False` (eval / no_grad)
    use no named buffers at all: every activation is a fresh allocation that is released as soon as the next layer
    has consumed it, so inference needs the live set only (a few tensors), and never touches saved activations."""

    def __init__(self, net: nn.Module, precision: str = "tf32", sr: bool = False):
        self.net = net
        self.sr = sr
        self.precision = precision
        self.dtype = PRECISIONS[precision]
        self.tdt = ops.TORCH_DTYPE[self.dtype]
        snet, rnet = net.SNet, net.RNet
        self.im_chn = rnet.in_chn
        self.sigma_chn = snet.conv_last.out_channels
        self.noise_cond = bool(net.noise_cond)
        self.extra_mode = rnet.extra_mode
        # conditioning channels of RNet (networks/VIRNet.py:36-40, 66-75): [kernel code (SISR, kernel_cond)] + [sigma]
        self.kernel_cond = bool(getattr(net, "kernel_cond", False)) if sr else False
        self.kc = net.KNet.tail[0].out_channels if self.kernel_cond else 0
        self.sc_extra = self.sigma_chn if self.noise_cond else 0
        self.n_extra = self.kc + self.sc_extra
        self.noise_avg = bool(snet.noise_avg)
        if rnet.extra_chn != self.n_extra:
            raise ValueError("RNet.extra_chn does not match the conditioning channels of the wrapper module")
        if not sr and self.noise_avg:
            # the reference itself cannot run this: RNet reflect-pads the 1x1 sigma "map" (utils/util_net.py:20-25)
            raise NotImplementedError("VIRAttResUNet(noise_avg=True): the reference's RNet cannot pad a 1x1 sigma map either")
        # per-pixel sigma map (noise_avg=False) -> the SFT layers see spatially varying conditioning
        self.spatial_extra = self.sc_extra > 0 and not self.noise_avg
        self.head_extra = self.n_extra if self.extra_mode in ("input", "both") else 0
        self.use_sft = self.extra_mode in ("down", "both") and self.n_extra > 0
        self.depth, self.n_feat, self.n_res = rnet.depth, rnet.n_feat, rnet.n_resblocks

        # ---- layer table (forward order) ----
        L: List[_Layer] = []
        s_convs = snet.conv_layers()
        self.s_layers = []
        for i, m in enumerate(s_convs):
            nm = "SNet.conv1" if i == 0 else ("SNet.conv_last" if i == len(s_convs) - 1 else f"SNet.mid{i}")
            self.s_layers.append(_Layer(nm, m, "conv", need_dgrad=i > 0))
        L += self.s_layers
        self.head = _Layer("RNet.head", rnet.head, "conv", need_dgrad=self.head_extra > 0)
        L.append(self.head)
        self.down = []
        for ii, blk in enumerate(rnet.down_path):
            res = [(_Layer(f"RNet.down{ii}.b{b}.conv1", rb.conv1, "conv", True),
                    _Layer(f"RNet.down{ii}.b{b}.conv2", rb.conv2, "conv", True)) for b, rb in enumerate(blk.body)]
            ds = _Layer(f"RNet.down{ii}.ds", blk.downsampler, "conv_s2", True) if ii + 1 < self.depth else None
            self.down.append((res, ds))
            for a, b in res:
                L += [a, b]
            if ds is not None:
                L.append(ds)
        self.up = []
        for k, blk in enumerate(rnet.up_path):
            us = _Layer(f"RNet.up{k}.us", blk.upsampler, "convT", True)
            res = [(_Layer(f"RNet.up{k}.b{b}.conv1", rb.conv1, "conv", True),
                    _Layer(f"RNet.up{k}.b{b}.conv2", rb.conv2, "conv", True)) for b, rb in enumerate(blk.body)]
            self.up.append((us, res))
            L.append(us)
            for a, b in res:
                L += [a, b]
        self.tail = _Layer("RNet.tail", rnet.tail, "conv", True)
        L.append(self.tail)
        # KNet (super-resolution only): the 3x3 convs run on the tensor-core kernels, the 9x9 head, the
        # channel-attention MLPs and the SFT MLPs read their fp32 parameters directly (vk_sisr.cu)
        self.k_blocks, self.k_tail = [], None
        if sr:
            knet = net.KNet
            for b, rb in enumerate(knet.body):
                self.k_blocks.append((_Layer(f"KNet.body{b}.conv1", rb.body[0], "conv", True),
                                      _Layer(f"KNet.body{b}.conv2", rb.body[2], "conv", True), rb.body[3]))
                L += [self.k_blocks[-1][0], self.k_blocks[-1][1]]
            self.k_tail = _Layer("KNet.tail", knet.tail[0], "conv", True)
            L.append(self.k_tail)
        self.layers = L

        self.wgrad_side_stream = os.environ.get("VIRNET_B200_WGRAD_STREAM", "1") != "0"
        # deterministic weight gradients: every split-K slice of vk_conv_wgrad stores its partial sums in its own slab
        # and the unpack kernel adds the slabs in slice order (no fp32 atomics).  Together with the atomic-free loss /
        # norm / bias reductions the denoising training step is then run-to-run bit-identical; costs one extra pass
        # over the partial slabs (about 0.6 GB at batch 32).  Also set by the trainers' `deterministic=True`.
        self.deterministic = os.environ.get("VIRNET_B200_DETERMINISTIC", "0") == "1"
        self._det_ran: Dict[int, _Layer] = {}          # layers whose wgrad ran in the current backward (det mode)
        self._det_tables: Dict = {}
        self.grad_sync = None          # dp.BucketedGradSync when the trainer overlaps the gradient all-reduce (world > 1)
        self._wg_stream, self._wg_events, self._wg_i = None, [], 0
        self._flat_key = None
        self._packed_version = None
        self._sets: "OrderedDict" = OrderedDict()      # (signature, slot) -> {"bufs": {...}, "owner": generation or None}
        self._cur: Optional[Dict] = None               # buffer set of the running forward / backward
        self._transient = False                        # True while a save=False forward runs
        self._tables: Dict = {}                        # per-batch-size SFT tables (small)
        self.max_cached_shapes = int(os.environ.get("VIRNET_B200_MAX_CACHED_SHAPES", "2"))
        self._gen = 0
        self._pending: Dict[int, Dict] = {}            # generation -> saved activations awaiting their backward
        self.saved = None                              # most recent saved forward (what the fused trainers consume)

    # ------------------------------------------------------------------
    # parameters: one flat fp32 buffer (params) + one for grads + wgrad workspace
    # ------------------------------------------------------------------
    def _params(self):
        return list(self.net.parameters())

    def _ensure_flat(self):
        params = self._params()
        dev = params[0].device
        if dev.type != "cuda":
            raise _l.VkError("virnet_b200 runs on CUDA (sm_100a) only: move the module to the GPU first; "
                             "there is no CPU fallback")
        key = (dev, tuple(p.data_ptr() for p in params))
        if key == self._flat_key:
            return
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4            # keep every tensor 16-byte aligned
        flat = torch.zeros(total, device=dev, dtype=torch.float32)
        for p, o in zip(params, offs):
            flat[o:o + p.numel()].view_as(p).copy_(p.data)
            p.data = flat[o:o + p.numel()].view_as(p)
        self.flat_params, self.flat_offsets, self.flat_total = flat, offs, total
        self.flat_grads = torch.zeros(total, device=dev, dtype=torch.float32)
        self.param_index = {id(p): i for i, p in enumerate(params)}
        self._flat_key = (dev, tuple(p.data_ptr() for p in params))
        self._build_packing(dev)
        self._packed_version = None
        self.release_buffers()

    def grad_view(self, p: torch.Tensor) -> torch.Tensor:
        i = self.param_index[id(p)]
        o = self.flat_offsets[i]
        return self.flat_grads[o:o + p.numel()].view_as(p)

    def _build_packing(self, dev):
        dt, tdt = self.dtype, self.tdt
        cp = lambda c: ops.chan_pad(c, dt)
        descs = []
        max_elems = 0
        ws_total = 0

        def add(src, dst, dim0, dim1, taps, rows, ld, dst_taps, mode):
            nonlocal max_elems
            d = _l.vk_pack_desc()
            d.src, d.dst = src.data_ptr(), dst.data_ptr()
            d.dim0, d.dim1, d.taps, d.rows, d.ld, d.dst_taps, d.mode = dim0, dim1, taps, rows, ld, dst_taps, mode
            descs.append(d)
            max_elems = max(max_elems, dst.numel())

        for ly in self.layers:
            w = ly.weight
            if ly.kind == "convT":
                ly.wf = torch.empty(1, 4 * ly.cout, cp(ly.cin), device=dev, dtype=tdt)
                add(w, ly.wf, ly.cin, ly.cout, 4, 4 * ly.cout, cp(ly.cin), 1, 2)
                ly.wd = torch.empty(4, _pad16(ly.cin), cp(ly.cout), device=dev, dtype=tdt)
                add(w, ly.wd, ly.cin, ly.cout, 4, _pad16(ly.cin), cp(ly.cout), 4, 3)
            else:
                ly.wf = torch.empty(ly.taps, _wrows(ly.cout), cp(ly.cin), device=dev, dtype=tdt)
                add(w, ly.wf, ly.cout, ly.cin, ly.taps, _wrows(ly.cout), cp(ly.cin), ly.taps, 0)
                if ly.need_dgrad:
                    ly.wd = torch.empty(ly.taps, _wrows(ly.cin), cp(ly.cout), device=dev, dtype=tdt)
                    add(w, ly.wd, ly.cout, ly.cin, ly.taps, _wrows(ly.cin), cp(ly.cout), ly.taps,
                        4 if ly.kind == "conv_s2" else 1)
            ws_total += w.numel()
        arr = (_l.vk_pack_desc * len(descs))(*descs)
        raw = bytes(memoryview(arr))
        self._pack_descs = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
        self._pack_n, self._pack_max = len(descs), max_elems
        # wgrad workspace: [taps][M][N] per layer, one flat buffer
        self.flat_ws = torch.zeros(ws_total, device=dev, dtype=torch.float32)
        o = 0
        udescs = []
        for ly in self.layers:
            n = ly.weight.numel()
            m_, n_ = (ly.cin, ly.cout) if ly.kind == "convT" else (ly.cout, ly.cin)
            ly.ws = self.flat_ws[o:o + n].view(ly.taps, m_, n_)
            o += n
            d = _l.vk_unpack_desc()
            d.ws, d.out, d.taps, d.mn = ly.ws.data_ptr(), self.grad_view(ly.weight).data_ptr(), ly.taps, m_ * n_
            udescs.append(d)
        arr = (_l.vk_unpack_desc * len(udescs))(*udescs)
        self._unpack_descs = torch.frombuffer(bytearray(bytes(memoryview(arr))), dtype=torch.uint8).to(dev)
        self._unpack_n, self._unpack_max = len(udescs), max(d.mn for d in udescs)
        self._det_tables = {}
        for ly in self.layers:
            ly.det = None
        self._csum_ws = ops.channel_sum_ws(max(max(ly.cout, ly.cin) for ly in self.layers), dev)
        self._csum_ws2 = ops.channel_sum_ws(16, dev)       # the side stream's own scratch (narrow-output bias sums)

    def mark_params_dirty(self):
        self._packed_version = None

    def _ensure_packed(self):
        ver = tuple(p._version for p in self._params())
        if ver == self._packed_version:
            return
        ops.pack_weights(self._pack_descs, self._pack_n, self._pack_max, dtype=self.dtype, round_tf32=True)
        self._packed_version = ver

    # ------------------------------------------------------------------
    # buffers
    # ------------------------------------------------------------------
    def release_buffers(self):
        """Drop every cached activation / gradient buffer (pending backwards are invalidated)."""
        self._sets.clear()
        self._tables = {}
        self._pending.clear()
        self._cur, self.saved = None, None

    def _begin(self, sig, save: bool):
        """Select the buffer set of a forward with input signature `sig`; returns the table the forward records its
        activations in."""
        self._transient = not save
        if not save:
            self._cur = None
            return _NoSave()
        slot = 0
        while True:
            key = (sig, slot)
            st = self._sets.get(key)
            if st is None:
                st = self._sets[key] = {"bufs": {}, "owner": None}
                break
            if st["owner"] is None or st["owner"] not in self._pending:
                break
            slot += 1                                   # its backward is still pending: leave it alone
        self._sets.move_to_end(key)
        for k in [k for k, v in self._sets.items() if v["owner"] is None or v["owner"] not in self._pending]:
            if len(self._sets) <= self.max_cached_shapes:
                break
            if k != key:
                del self._sets[k]
        self._gen += 1
        st["owner"] = self._gen
        self._cur = st
        A: Dict = {"gen": self._gen, "set": st}
        return A

    def _commit(self, A):
        """Register the saved activations of a differentiable forward; returns its generation number."""
        self._pending[A["gen"]] = A
        # forwards whose outputs were dropped without a backward would otherwise pin their sets forever
        while len(self._pending) > 8:
            self._pending.pop(next(iter(self._pending)))
        self.saved = A
        return A["gen"]

    def _take_saved(self, gen):
        if gen is None:
            A = self.saved
            if A is None:
                raise _l.VkError("backward called without a saved forward")
            gen = A["gen"]
        A = self._pending.pop(gen, None)
        if A is None:
            raise _l.VkError("backward of a forward whose saved activations are gone: it was already back-propagated, or "
                             "more than 8 differentiable forwards were pending (retain_graph / double backward are not supported)")
        if A["set"]["owner"] != gen:
            raise _l.VkError("stale saved activations: their buffers were reused by a later forward")
        if self.saved is A:
            self.saved = None
        self._cur, self._transient = A["set"], False
        return A

    def _buf(self, name, shape, dtype=None):
        dtype = dtype or self.tdt
        if self._transient:
            return torch.empty(shape, device=self.flat_params.device, dtype=dtype)
        bufs = self._cur["bufs"]
        key = (name, tuple(shape), dtype)
        b = bufs.get(key)
        if b is None:
            b = bufs[key] = torch.empty(shape, device=self.flat_params.device, dtype=dtype)
        return b

    # ------------------------------------------------------------------
    # conv helpers
    # ------------------------------------------------------------------
    def _conv(self, x, ly: _Layer, kind, *, w=None, cout=None, bias=True, **kw):
        ops.conv_igemm(x, ly.wf if w is None else w, dtype=self.dtype, kind=kind, cout=ly.cout if cout is None else cout,
                       bias=ly.bias if bias else None, round_out2=self.dtype == VK_TF32, **kw)

    # ------------------------------------------------------------------
    # forward
    # ------------------------------------------------------------------
    def forward(self, x: torch.Tensor, save: bool):
        self._ensure_flat()
        self._ensure_packed()
        if x.dtype != torch.float32 or not x.is_cuda:
            raise _l.VkError("input must be a CUDA fp32 NCHW tensor")
        x = x.contiguous()
        N, C, H, W = x.shape
        assert C == self.im_chn
        dt = self.dtype
        cp = lambda c: ops.chan_pad(c, dt)
        mod = 2 ** (self.depth - 1)
        Hp, Wp = (H + mod - 1) // mod * mod, (W + mod - 1) // mod * mod
        if Hp > 2 * H - 1 or Wp > 2 * W - 1:
            raise _l.VkError("image too small for reflect padding")
        A = self._begin(("den", N, H, W), save)

        # ---- SNet ----
        xs = self._buf("xs", (N, H, W, cp(C)))
        ops.pack_input(x, xs, dtype=dt)
        A["xs"] = xs
        cur = xs
        for i, ly in enumerate(self.s_layers[:-1]):
            o = self._buf(f"s{i}", (N, H, W, cp(ly.cout)))
            self._conv(cur, ly, VK_CONV3X3_S1, ldo=cp(ly.cout), out2=o, alpha=0.25)
            A[f"s{i}"] = o
            cur = o
        sigma = torch.empty(N, self.sigma_chn, H, W, device=x.device, dtype=torch.float32)
        self._conv(cur, self.s_layers[-1], VK_CONV3X3_S1, epi=VK_EPI_NCHW_F32, out1=sigma, act_expclamp=True,
                   clamp=(SNET_LOG_MIN, SNET_LOG_MAX))

        # ---- RNet ----
        if self.extra_mode != "null" and self.n_extra == 0:
            raise _l.VkError("extra_mode != 'Null' needs conditioning maps (noise_cond=True): the reference fails on "
                             "pad_input(None) here as well (networks/AttResUNet.py:147-149)")
        if self.use_sft:
            # extra_mode 'Down' / 'Both': the per-pixel sigma map modulates the down path through the AttLayers
            src = ops.ExtraSource(None, sigma, 1, (1 << self.sigma_chn) - 1, H, W, Hp, Wp)
            mu = self._rnet_forward(A, x, 1, None, src, save)
            if save:
                A["sigma"], A["src"], A["x"] = sigma, src, x
                A["shape"] = (N, C, H, W, Hp, Wp, None)
                A["rnet_general"] = True
                self._commit(A)
            return mu, sigma
        cin0 = C + self.head_extra
        r0 = self._buf("r0", (N, Hp, Wp, cp(cin0)))
        if self.head_extra:
            ops.pack_input(x, r0, dtype=dt, extra=sigma, extra_is_map=True, extra_sqrt_mask=(1 << self.sigma_chn) - 1)
        else:
            ops.pack_input(x, r0, dtype=dt)
        A["r0"] = r0
        nf = self.n_feat
        h, w = Hp, Wp
        X = self._buf("X.head", (N, h, w, nf[0]))
        Act = self._buf("A.head", (N, h, w, nf[0]))
        self._conv(r0, self.head, VK_CONV3X3_S1, ldo=nf[0], out1=X, out2=Act, alpha=0.2)
        bridges = []
        dims = [(h, w)]
        for ii, (res, ds) in enumerate(self.down):
            c = nf[ii]
            for b, (c1, c2) in enumerate(res):
                A[f"d{ii}.{b}.a"] = Act
                Bt = self._buf(f"d{ii}.{b}.B", (N, h, w, c))
                self._conv(Act, c1, VK_CONV3X3_S1, ldo=c, out2=Bt, alpha=0.2)
                A[f"d{ii}.{b}.b"] = Bt
                Xn = self._buf(f"d{ii}.{b}.X", (N, h, w, c))
                last = b == len(res) - 1
                An = None if last else self._buf(f"d{ii}.{b}.A", (N, h, w, c))
                self._conv(Bt, c2, VK_CONV3X3_S1, ldo=c, resid=X, out1=Xn, out2=An, alpha=0.2)
                X, Act = Xn, An
            if ds is not None:
                bridges.append(X)
                A[f"d{ii}.x"] = X
                h2, w2 = (h + 1) // 2, (w + 1) // 2
                Xd = self._buf(f"d{ii}.ds.X", (N, h2, w2, nf[ii + 1]))
                Ad = self._buf(f"d{ii}.ds.A", (N, h2, w2, nf[ii + 1]))
                self._conv(X, ds, VK_CONV3X3_S2, ldo=nf[ii + 1], out1=Xd, out2=Ad, alpha=0.2)
                X, Act, h, w = Xd, Ad, h2, w2
                dims.append((h, w))
        for k, (us, res) in enumerate(self.up):
            lvl = self.depth - 2 - k
            c = nf[lvl]
            A[f"u{k}.x"] = X
            h, w = dims[lvl]
            Xu = self._buf(f"u{k}.us.X", (N, h, w, c))
            Au = self._buf(f"u{k}.us.A", (N, h, w, c))
            self._conv(X, us, VK_CONVT2X2_S2, ldo=c, resid=bridges[lvl], out1=Xu, out2=Au, alpha=0.2)
            X, Act = Xu, Au
            for b, (c1, c2) in enumerate(res):
                A[f"u{k}.{b}.a"] = Act
                Bt = self._buf(f"u{k}.{b}.B", (N, h, w, c))
                self._conv(Act, c1, VK_CONV3X3_S1, ldo=c, out2=Bt, alpha=0.2)
                A[f"u{k}.{b}.b"] = Bt
                Xn = self._buf(f"u{k}.{b}.X", (N, h, w, c))
                last = b == len(res) - 1
                An = None if last else self._buf(f"u{k}.{b}.A", (N, h, w, c))
                self._conv(Bt, c2, VK_CONV3X3_S1, ldo=c, resid=X, out1=Xn, out2=An, alpha=0.2)
                X, Act = Xn, An
        A["tail.x"] = X
        mu = torch.empty(N, C, H, W, device=x.device, dtype=torch.float32)
        self._conv(X, self.tail, VK_CONV3X3_S1, epi=VK_EPI_NCHW_F32, resid=x, out1=mu, crop=(H, W))
        if save:
            A["sigma"] = sigma
            A["shape"] = (N, C, H, W, Hp, Wp, dims)
            self._commit(A)
        return mu, sigma

    # ------------------------------------------------------------------
    # backward
    # ------------------------------------------------------------------
    def _wgrad(self, ly: _Layer, a, b, kind):
        """a: M operand, b: N operand (see include/virnet_b200.h vk_wgrad_args)."""
        m_valid, n_valid = (ly.cin, ly.cout) if ly.kind == "convT" else (ly.cout, ly.cin)
        dbias = None
        if ly.bias is not None and ly.kind != "convT":
            dbias = self.grad_view(ly.bias)
        kw = {}
        # narrow-OUTPUT 3x3 layers (tail, last SNet layer): pass the operands the other way round so that the narrow dY is
        # the N operand (vk_wgrad_args.swapped); the bias gradient is then a plain channel sum of dY
        swap_bias = None
        if (kind == VK_CONV3X3_S1 and self.dtype == VK_BF16 and ly.cout <= 16 < ly.cin and a.shape[-1] == 16
                and os.environ.get("VIRNET_B200_NO_WGRAD_SWAP") is None):
            swap_bias, dbias = dbias, None
            a, b, m_valid, n_valid = b, a, n_valid, m_valid
            kw["swapped"] = True
        if self.deterministic:
            key = (tuple(a.shape), tuple(b.shape), kind)
            ent = ly.det
            if ent is None or ent["key"] != key:
                slices, bias_slots = ops.conv_wgrad_plan(a, b, dtype=self.dtype, kind=kind, m_valid=m_valid,
                                                         n_valid=n_valid, dbias=dbias, swapped=kw.get("swapped", False))
                dev = self.flat_params.device
                ent = ly.det = {"key": key, "slices": slices, "bias_slots": bias_slots,
                                "partials": torch.empty((slices,) + tuple(ly.ws.shape), device=dev, dtype=torch.float32),
                                "dbias": None if dbias is None else torch.empty(bias_slots, m_valid, device=dev,
                                                                                dtype=torch.float32)}
                self._det_tables = {}
            self._det_ran[id(ly)] = ly
            kw.update(partials=ent["partials"], dbias_partials=ent["dbias"])
        ws = self._wg_stream
        if ws is None:
            ops.conv_wgrad(a, b, ly.ws, dtype=self.dtype, kind=kind, m_valid=m_valid, n_valid=n_valid, dbias=dbias, **kw)
            if swap_bias is not None:
                ops.channel_sum(b, n_valid, swap_bias, dtype=self.dtype, ws=self._csum_ws)
            return
        # weight gradients are leaves of the backward graph: run them on a side stream so their CTAs fill the SMs
        # the persistent dgrad kernels leave idle in their last (partial) round
        ev = self._wg_events[self._wg_i % len(self._wg_events)]
        self._wg_i += 1
        ev.record()
        ws.wait_event(ev)
        with torch.cuda.stream(ws):
            ops.conv_wgrad(a, b, ly.ws, dtype=self.dtype, kind=kind, m_valid=m_valid, n_valid=n_valid, dbias=dbias, **kw)
            if swap_bias is not None:
                ops.channel_sum(b, n_valid, swap_bias, dtype=self.dtype, ws=self._csum_ws2)

    def _bucket_done(self, first_layer: _Layer):
        """Every layer from `first_layer` (forward order) to the end of the previous bucket has had its weight-gradient
        kernels issued: hand that range of the flat gradient buffer to the overlapped all-reduce (dp.BucketedGradSync)."""
        if self.grad_sync is not None:
            self.grad_sync.bucket_ready(self.layers.index(first_layer), self._wg_stream)

    def _det_table(self):
        """Descriptor table of the deterministic unpack: per layer the split-K slabs written in this backward (summed in
        slice order) plus, where the kernel produced them, the bias slots; layers whose wgrad did not run keep their
        zero workspace.  Cached per set of layers / shapes."""
        key = tuple((id(ly), ly.det["key"]) for ly in self.layers if id(ly) in self._det_ran)
        hit = self._det_tables.get(key)
        if hit is not None:
            return hit
        descs, offs = [], []
        for ly in self.layers:
            offs.append(len(descs))
            m_, n_ = (ly.cin, ly.cout) if ly.kind == "convT" else (ly.cout, ly.cin)
            d = _l.vk_unpack_desc()
            d.out, d.taps, d.mn = self.grad_view(ly.weight).data_ptr(), ly.taps, m_ * n_
            if id(ly) in self._det_ran:
                ent = ly.det
                d.ws, d.nslices, d.slice_stride = ent["partials"].data_ptr(), ent["slices"], ly.taps * m_ * n_
                descs.append(d)
                if ent["dbias"] is not None:
                    b = _l.vk_unpack_desc()
                    b.ws, b.out, b.taps, b.mn = ent["dbias"].data_ptr(), self.grad_view(ly.bias).data_ptr(), 1, m_
                    b.nslices, b.slice_stride = ent["bias_slots"], m_
                    descs.append(b)
            else:
                d.ws, d.nslices, d.slice_stride = ly.ws.data_ptr(), 1, 0
                descs.append(d)
        offs.append(len(descs))
        arr = (_l.vk_unpack_desc * len(descs))(*descs)
        table = torch.frombuffer(bytearray(bytes(memoryview(arr))), dtype=torch.uint8).to(self.flat_params.device)
        hit = self._det_tables[key] = (table, offs, max(d.mn for d in descs))
        return hit

    def _begin_wgrads(self):
        """Start of a backward pass: clear the gradient buffers the weight-gradient kernels accumulate into."""
        self.flat_grads.zero_()
        if self.deterministic:
            self._det_ran = {}          # flat_ws is never written in this mode and stays zero
        else:
            self.flat_ws.zero_()

    def unpack_range(self, i0: int, i1: int):
        """Workspace -> parameter layout for layers [i0, i1) (a sub-range of the batched descriptor table)."""
        sz = C.sizeof(_l.vk_unpack_desc)
        if self.deterministic:
            table, offs, max_mn = self._det_table()
            d0, d1 = offs[i0], offs[i1]
            ops.wgrad_unpack_batched(table[d0 * sz:d1 * sz], d1 - d0, max_mn, accumulate=False)
            return
        descs = self._unpack_descs[i0 * sz:i1 * sz]
        ops.wgrad_unpack_batched(descs, i1 - i0, self._unpack_max, accumulate=False)

    def _unpack_all(self):
        self.unpack_range(0, len(self.layers))

    def layer_flat_range(self, i0: int, i1: int):
        """[begin, end) of the flat parameter / gradient buffer covered by layers [i0, i1) (weights and biases)."""
        def first_off(ly):
            return self.flat_offsets[self.param_index[id(ly.weight)]]
        begin = first_off(self.layers[i0])
        end = first_off(self.layers[i1]) if i1 < len(self.layers) else self.flat_total
        return begin, end

    def _dgrad(self, g, ly: _Layer, kind, cout, **kw):
        ops.conv_igemm(g, ly.wd, dtype=self.dtype, kind=kind, cout=cout, bias=None, **kw)

    def _resblock_bwd(self, tag, c1, c2, gX, shape):
        A = self._saved_A
        N, h, w, c = shape
        a, bt = A[tag + ".a"], A[tag + ".b"]
        self._wgrad(c2, gX, bt, VK_CONV3X3_S1)
        gF = self._buf("g." + tag + ".F", (N, h, w, c))
        self._dgrad(gX, c2, VK_CONV3X3_S1, c, ldo=c, mask=bt, out1=gF, alpha=0.2)
        self._wgrad(c1, gF, a, VK_CONV3X3_S1)
        gXp = self._buf("g." + tag + ".X", (N, h, w, c))
        self._dgrad(gF, c1, VK_CONV3X3_S1, c, ldo=c, mask=a, resid=gX, out1=gXp, alpha=0.2)
        return gXp

    def backward(self, g_mu: Optional[torch.Tensor], g_sigma: Optional[torch.Tensor], gen: Optional[int] = None):
        """Accumulates parameter gradients into self.flat_grads (zeroed here first).  `gen`: generation of the forward to
        differentiate (default: the most recent saved one)."""
        A = self._saved_A = self._take_saved(gen)
        if A.get("rnet_general"):
            return self._backward_general_denoise(A, g_mu, g_sigma)
        N, C, H, W, Hp, Wp, dims = A["shape"]
        dt = self.dtype
        cp = lambda c: ops.chan_pad(c, dt)
        nf = self.n_feat
        self._begin_wgrads()
        if self.grad_sync is not None:
            self.grad_sync.begin()
        if self.wgrad_side_stream and self._wg_stream is None:
            self._wg_stream = torch.cuda.Stream(device=self.flat_params.device)
            self._wg_events = [torch.cuda.Event() for _ in range(8)]
        if not self.wgrad_side_stream:
            self._wg_stream = None
        gR0 = None
        if g_mu is not None:
            g_mu = g_mu.contiguous().float()
            G = self._buf("g.mu", (N, Hp, Wp, cp(C)))
            ops.pack_grad(g_mu, G, dtype=dt)
            # tail
            self._wgrad(self.tail, G, A["tail.x"], VK_CONV3X3_S1)
            h, w = dims[0]
            gX = self._buf("g.tail.X", (N, h, w, nf[0]))
            self._dgrad(G, self.tail, VK_CONV3X3_S1, nf[0], ldo=nf[0], out1=gX)
            # up path, reversed
            g_bridge = {}
            for k in reversed(range(len(self.up))):
                us, res = self.up[k]
                lvl = self.depth - 2 - k
                c = nf[lvl]
                h, w = dims[lvl]
                for b in reversed(range(len(res))):
                    gX = self._resblock_bwd(f"u{k}.{b}", res[b][0], res[b][1], gX, (N, h, w, c))
                g_bridge[lvl] = gX
                xlow = A[f"u{k}.x"]
                self._wgrad(us, xlow, gX, VK_CONVT2X2_S2)
                ops.channel_sum(gX, c, self.grad_view(us.bias), dtype=dt, ws=self._csum_ws)
                self._bucket_done(us)            # everything from this up block to the tail has its gradient
                hl, wl = dims[lvl + 1]
                gXl = self._buf(f"g.u{k}.low", (N, hl, wl, nf[lvl + 1]))
                self._dgrad(gX, us, VK_CONV2X2_S2, nf[lvl + 1], ldo=nf[lvl + 1], out1=gXl)
                gX = gXl
            # down path, reversed
            for ii in reversed(range(self.depth)):
                res, ds = self.down[ii]
                c = nf[ii]
                h, w = dims[ii]
                if ds is not None:
                    # gX is the gradient w.r.t. the stride-2 conv output (coarse grid)
                    self._wgrad(ds, gX, A[f"d{ii}.x"], VK_CONV3X3_S2)
                    gXf = self._buf(f"g.d{ii}.ds", (N, h, w, c))
                    self._dgrad(gX, ds, VK_CONV3X3_S2_DGRAD, c, ldo=c, resid=g_bridge[ii], out1=gXf, out_hw=(h, w))
                    gX = gXf
                for b in reversed(range(len(res))):
                    gX = self._resblock_bwd(f"d{ii}.{b}", res[b][0], res[b][1], gX, (N, h, w, c))
                if ii > 0:
                    self._bucket_done(res[0][0])     # this level of the down path (and its down-sampler) is complete
            # head
            self._wgrad(self.head, gX, A["r0"], VK_CONV3X3_S1)
            if self.head_extra:
                cin0 = C + self.head_extra
                gR0 = self._buf("g.r0", (N, Hp, Wp, cp(cin0)))
                self._dgrad(gX, self.head, VK_CONV3X3_S1, cin0, ldo=cp(cin0), out1=gR0)
        # ---- SNet ----
        if g_sigma is not None or gR0 is not None:
            sc = self.sigma_chn
            if g_sigma is not None:
                g_sigma = g_sigma.contiguous().float()
            GS = self._buf("g.sig", (N, H, W, cp(sc)))
            ops.sigma_head_bwd(A["sigma"], g_sigma, gR0, C, GS, dtype=dt, log_lo=SNET_LOG_MIN, log_hi=SNET_LOG_MAX)
            g = GS
            nS = len(self.s_layers)
            for i in reversed(range(nS)):
                ly = self.s_layers[i]
                inp = A["xs"] if i == 0 else A[f"s{i - 1}"]
                self._wgrad(ly, g, inp, VK_CONV3X3_S1)
                if i > 0:
                    gn = self._buf(f"g.s{i - 1}", (N, H, W, cp(ly.cin)))
                    self._dgrad(g, ly, VK_CONV3X3_S1, ly.cin, ldo=cp(ly.cin), mask=inp, out1=gn, alpha=0.25)
                    g = gn
        # ---- workspace -> parameter-layout gradients ----
        if self.grad_sync is not None:
            self._bucket_done(self.layers[0])    # the rest: down level 0, head, SNet
            self.grad_sync.finish()
        else:
            if self._wg_stream is not None:
                torch.cuda.current_stream().wait_stream(self._wg_stream)
            self._unpack_all()
        self._saved_A = None
        A["set"]["owner"] = None

    # ------------------------------------------------------------------
    # super-resolution network (networks/VIRNet.py:80-97): forward, and backward w.r.t. every parameter
    # ------------------------------------------------------------------
    def _sft_tables(self, N):
        """(mul, add) buffers of every AttLayer, their gradient accumulators (one flat buffer, zeroed per backward) and
        the device descriptor table of the batched SFT-MLP kernels; cached per batch size."""
        key = ("sft_tables", N, self._flat_key)
        hit = self._tables.get(key)
        if hit is not None:
            return hit[:4]
        rnet, nfeat, dev, f32 = self.net.RNet, self.n_feat, self.flat_params.device, torch.float32
        layers = [(ii, b, which, getattr(rb, which)) for ii, blk in enumerate(rnet.down_path)
                  for b, rb in enumerate(blk.body) for which in ("sft1", "sft2")]
        total = sum(N * nfeat[ii] for ii, _, _, _ in layers)
        vals = torch.empty(2 * total, device=dev, dtype=f32)       # mul | add
        grads = torch.zeros(2 * total, device=dev, dtype=f32)      # dmul | dadd
        sft, dmd, descs, o = {}, {}, [], 0
        for ii, b, which, att in layers:
            c = nfeat[ii]
            mul, add = vals[o:o + N * c].view(N, c), vals[total + o:total + o + N * c].view(N, c)
            dm, dd = grads[o:o + N * c].view(N, c), grads[total + o:total + o + N * c].view(N, c)
            o += N * c
            sft[(ii, b, which)], dmd[(ii, b, which)] = (mul, add), (dm, dd)
            d = _l.vk_sft_desc()
            ps = (att.conv1.weight, att.conv1.bias, att.conv2.weight, att.conv2.bias, att.mul_conv.weight,
                  att.mul_conv.bias, att.add_conv.weight, att.add_conv.bias)
            for nm, p_ in zip(("w1", "b1", "w2", "b2", "wm", "bm", "wa", "ba"), ps):
                setattr(d, nm, p_.data_ptr())
                setattr(d, "g" + nm, self.grad_view(p_).data_ptr())
            d.mul, d.add, d.dmul, d.dadd = mul.data_ptr(), add.data_ptr(), dm.data_ptr(), dd.data_ptr()
            d.c1, d.c2, d.c = att.conv1.out_channels, att.conv2.out_channels, c
            assert d.c1 + d.c2 <= c
            descs.append(d)
        arr = (_l.vk_sft_desc * len(descs))(*descs)
        table = torch.frombuffer(bytearray(bytes(memoryview(arr))), dtype=torch.uint8).to(dev)
        n_params = sum(p_.numel() for _, _, _, att in layers for p_ in att.parameters())
        out = (sft, table, len(descs), max(nfeat), dmd, grads, vals, n_params)
        self._tables[key] = out
        return out[:4]

    def _rnet_forward(self, S, x, sf, extra_cst, src, save, sqrt_mask=0):
        """AttResUNet.forward (networks/AttResUNet.py:141-175) for every extra_mode.  x: NCHW fp32 (the LR image when
        sf > 1: the nearest up-sampling is fused into the packing); extra_cst [N, E] fp32: per-sample constant
        conditioning values (SFT by conv epilogue); src: ops.ExtraSource when a conditioning channel varies per pixel
        (SFT by vk_sft_apply).  Records what the backward needs in S; returns mu (NCHW fp32, caller-owned)."""
        dt, dev, f32 = self.dtype, x.device, torch.float32
        cp = lambda c: ops.chan_pad(c, dt)
        N, C, h, w = x.shape
        Hh, Ww = h * sf, w * sf
        mod = 2 ** (self.depth - 1)
        Hp, Wp = (Hh + mod - 1) // mod * mod, (Ww + mod - 1) // mod * mod
        if Hp > 2 * Hh - 1 or Wp > 2 * Ww - 1:
            raise _l.VkError("image too small for reflect padding")
        nfeat, rnet = self.n_feat, self.net.RNet
        sft_const = self.use_sft and src is None
        sft_spatial = self.use_sft and src is not None
        rnd = dt == VK_TF32
        sft = None
        if sft_const:
            sft, sft_descs, sft_n, sft_maxc = self._sft_tables(N)
            ops.sft_mlp_batched(sft_descs, sft_n, sft_maxc, extra_cst, sqrt_mask=sqrt_mask)
        cin0 = C + self.head_extra
        r0 = self._buf("sr.r0", (N, Hp, Wp, cp(cin0)))
        if self.head_extra and src is not None:
            ops.pack_input_mixed(x, r0, src, dtype=dt, sf=sf)
        elif self.head_extra:
            ops.pack_input(x, r0, dtype=dt, sf=sf, extra=extra_cst, extra_is_map=False, extra_sqrt_mask=sqrt_mask)
        else:
            ops.pack_input(x, r0, dtype=dt, sf=sf)
        S["r0"] = r0
        hh, ww = Hp, Wp

        def att(ii, b, which):
            return getattr(rnet.down_path[ii].body[b], which)

        def modulated(name, X, ii, b, which, c):
            """lrelu(X * mul + add) with the AttLayer evaluated per pixel (spatially varying conditioning)."""
            out = self._buf(name, tuple(X.shape))
            ops.sft_apply(X, out, att(ii, b, which), src, dtype=dt, c=c, alpha=0.2, round_tf32=rnd)
            return out

        X = self._buf("sr.X.head", (N, hh, ww, nfeat[0]))
        if sft_spatial:
            self._conv(r0, self.head, VK_CONV3X3_S1, ldo=nfeat[0], out1=X)
            Act = modulated("sr.A.head", X, 0, 0, "sft1", nfeat[0])
        else:
            Act = self._buf("sr.A.head", (N, hh, ww, nfeat[0]))
            self._conv(r0, self.head, VK_CONV3X3_S1, ldo=nfeat[0], out1=X, out2=Act, alpha=0.2,
                       sft=sft[(0, 0, "sft1")] if sft_const else None)
        bridges, dims = [], [(hh, ww)]
        for ii, (res, ds) in enumerate(self.down):
            c = nfeat[ii]
            for b, (c1, c2) in enumerate(res):
                tag = f"d{ii}.{b}"
                last = b == len(res) - 1
                Xn = self._buf(f"sr.{tag}.X", (N, hh, ww, c))
                if sft_spatial:
                    F1 = self._buf(f"sr.{tag}.F1", (N, hh, ww, c))
                    self._conv(Act, c1, VK_CONV3X3_S1, ldo=c, out1=F1)
                    Bt = modulated(f"sr.{tag}.B", F1, ii, b, "sft2", c)
                    S[tag + ".x"], S[tag + ".a"], S[tag + ".f1"], S[tag + ".b"] = X, Act, F1, Bt
                    self._conv(Bt, c2, VK_CONV3X3_S1, ldo=c, resid=X, out1=Xn)
                    X, Act = Xn, (None if last else modulated(f"sr.{tag}.A", Xn, ii, b + 1, "sft1", c))
                    continue
                Bt = self._buf(f"sr.{tag}.B", (N, hh, ww, c))
                if sft_const:
                    F1 = self._buf(f"sr.{tag}.F1", (N, hh, ww, c)) if save else None
                    # conv1 emits a2 = lrelu(fea1 * mul2 + add2); training also keeps fea1 (needed for d mul2)
                    self._conv(Act, c1, VK_CONV3X3_S1, ldo=c, out1=F1, out2=Bt, alpha=0.2, sft=sft[(ii, b, "sft2")])
                    S[tag + ".x"], S[tag + ".f1"] = X, F1
                else:
                    self._conv(Act, c1, VK_CONV3X3_S1, ldo=c, out2=Bt, alpha=0.2)
                S[tag + ".a"], S[tag + ".b"] = Act, Bt
                if last:
                    self._conv(Bt, c2, VK_CONV3X3_S1, ldo=c, resid=X, out1=Xn)
                    X, Act = Xn, None
                else:
                    An = self._buf(f"sr.{tag}.A", (N, hh, ww, c))
                    self._conv(Bt, c2, VK_CONV3X3_S1, ldo=c, resid=X, out1=Xn, out2=An, alpha=0.2,
                               sft=sft[(ii, b + 1, "sft1")] if sft_const else None)
                    X, Act = Xn, An
            if ds is not None:
                bridges.append(X)
                S[f"d{ii}.xlast"] = X
                h2, w2 = (hh + 1) // 2, (ww + 1) // 2
                Xd = self._buf(f"sr.d{ii}.ds.X", (N, h2, w2, nfeat[ii + 1]))
                if sft_spatial:
                    self._conv(X, ds, VK_CONV3X3_S2, ldo=nfeat[ii + 1], out1=Xd)
                    hh, ww = h2, w2
                    Ad = modulated(f"sr.d{ii}.ds.A", Xd, ii + 1, 0, "sft1", nfeat[ii + 1])
                else:
                    Ad = self._buf(f"sr.d{ii}.ds.A", (N, h2, w2, nfeat[ii + 1]))
                    self._conv(X, ds, VK_CONV3X3_S2, ldo=nfeat[ii + 1], out1=Xd, out2=Ad, alpha=0.2,
                               sft=sft[(ii + 1, 0, "sft1")] if sft_const else None)
                    hh, ww = h2, w2
                X, Act = Xd, Ad
                dims.append((hh, ww))
        for k, (us, res) in enumerate(self.up):
            lvl = self.depth - 2 - k
            c = nfeat[lvl]
            S[f"u{k}.x"] = X
            hh, ww = dims[lvl]
            Xu = self._buf(f"sr.u{k}.us.X", (N, hh, ww, c))
            Au = self._buf(f"sr.u{k}.us.A", (N, hh, ww, c))
            self._conv(X, us, VK_CONVT2X2_S2, ldo=c, resid=bridges[lvl], out1=Xu, out2=Au, alpha=0.2)
            X, Act = Xu, Au
            for b, (c1, c2) in enumerate(res):
                S[f"u{k}.{b}.a"] = Act
                Bt = self._buf(f"sr.u{k}.{b}.B", (N, hh, ww, c))
                self._conv(Act, c1, VK_CONV3X3_S1, ldo=c, out2=Bt, alpha=0.2)
                S[f"u{k}.{b}.b"] = Bt
                Xn = self._buf(f"sr.u{k}.{b}.X", (N, hh, ww, c))
                last = b == len(res) - 1
                An = None if last else self._buf(f"sr.u{k}.{b}.A", (N, hh, ww, c))
                self._conv(Bt, c2, VK_CONV3X3_S1, ldo=c, resid=X, out1=Xn, out2=An, alpha=0.2)
                X, Act = Xn, An
        S["tail.x"] = X
        # tail: + bias, crop, + x_in (for SISR the nearest-upsampled LR image, AttResUNet.py:173) -> NCHW fp32
        if sf > 1:
            x_res = self._buf("sr.xup", (N, C, Hh, Ww), f32)
            ops.upsample_nearest_nchw(x, x_res, sf)
        else:
            x_res = x
        mu = torch.empty(N, C, Hh, Ww, device=dev, dtype=f32)
        self._conv(X, self.tail, VK_CONV3X3_S1, epi=VK_EPI_NCHW_F32, resid=x_res, out1=mu, crop=(Hh, Ww))
        S["rnet_dims"] = (Hh, Ww, Hp, Wp, dims)
        S["sft"] = sft
        return mu

    def forward_sr(self, x: torch.Tensor, sf: int, save: bool = False):
        """x: LR image NCHW fp32 -> (mu [N,C,H*sf,W*sf], kinfo [N,3], sigma [N,1,1,1] or, with noise_avg=False, the
        per-pixel variance map [N,sigma_chn,H,W]) — networks/VIRNet.py:80-97 for every constructor configuration."""
        net = self.net
        if self.extra_mode != "null" and self.n_extra == 0:
            raise _l.VkError("extra_mode != 'Null' needs conditioning maps (noise_cond or kernel_cond): the reference "
                             "fails on pad_input(None) here as well (networks/AttResUNet.py:147-149)")
        self._ensure_flat()
        self._ensure_packed()
        if x.dtype != torch.float32 or not x.is_cuda:
            raise _l.VkError("input must be a CUDA fp32 NCHW tensor")
        x = x.contiguous()
        N, C, h, w = x.shape
        sf = int(sf)
        dt, dev = self.dtype, x.device
        cp = lambda c: ops.chan_pad(c, dt)
        f32 = torch.float32
        S = self._begin(("sr", N, h, w, sf), save)

        # ---- SNet (DnCNN.py:37-44): global average of the log-variance (noise_avg) or a per-pixel map ----
        xs = self._buf("sr.xs", (N, h, w, cp(C)))
        ops.pack_input(x, xs, dtype=dt)
        S["xs"] = xs
        cur = xs
        for i, ly in enumerate(self.s_layers[:-1]):
            o = self._buf(f"sr.s{i}", (N, h, w, cp(ly.cout)))
            self._conv(cur, ly, VK_CONV3X3_S1, ldo=cp(ly.cout), out2=o, alpha=0.25)
            S[f"s{i}"] = o
            cur = o
        sc = self.sigma_chn
        if self.noise_avg:
            logvar = self._buf("sr.logvar", (N, sc, h, w), f32)
            self._conv(cur, self.s_layers[-1], VK_CONV3X3_S1, epi=VK_EPI_NCHW_F32, out1=logvar)
            sigma = torch.empty(N, sc, 1, 1, device=dev, dtype=f32)
            ops.gap_head(logvar, sigma, exp_mask=(1 << sc) - 1, lo=SNET_LOG_MIN, hi=SNET_LOG_MAX)
        else:
            sigma = torch.empty(N, sc, h, w, device=dev, dtype=f32)
            self._conv(cur, self.s_layers[-1], VK_CONV3X3_S1, epi=VK_EPI_NCHW_F32, out1=sigma, act_expclamp=True,
                       clamp=(SNET_LOG_MIN, SNET_LOG_MAX))

        # ---- KNet (KNet.py:52-59) ----
        knet = net.KNet
        nfk = knet.head.out_channels
        kh, kw = (h - 1) // 4 + 1, (w - 1) // 4 + 1
        H = self._buf("sr.k.h0", (N, kh, kw, cp(nfk)))
        ops.knet_head(x, knet.head.weight, H, dtype=dt)
        for b, (c1, c2, ca) in enumerate(self.k_blocks):
            A = self._buf(f"sr.k{b}.a", (N, kh, kw, cp(nfk)))
            self._conv(H, c1, VK_CONV3X3_S1, ldo=cp(nfk), out2=A, alpha=0.2)
            F_ = self._buf(f"sr.k{b}.f", (N, kh, kw, cp(nfk)))
            self._conv(A, c2, VK_CONV3X3_S1, ldo=cp(nfk), out1=F_)
            Hn = self._buf(f"sr.k.h{b + 1}", (N, kh, kw, cp(nfk)))
            ops.ca_layer(F_, H, ca.body[0].weight, ca.body[0].bias, ca.body[2].weight, ca.body[2].bias, Hn, dtype=dt,
                         c=nfk, alpha=0.2)
            S[f"k{b}.h"], S[f"k{b}.a"], S[f"k{b}.f"] = H, A, F_
            H = Hn
        S["k.hlast"] = H
        kcn = self.k_tail.cout
        kraw = self._buf("sr.k.raw", (N, kcn, kh, kw), f32)
        self._conv(H, self.k_tail, VK_CONV3X3_S1, epi=VK_EPI_NCHW_F32, out1=kraw)
        kinfo = torch.empty(N, kcn, device=dev, dtype=f32)
        ops.gap_head(kraw, kinfo, exp_mask=0b011, tanh_mask=1 << (kcn - 1), lo=KNET_LOG_MIN, hi=KNET_LOG_MAX)

        # ---- conditioning (VIRNet.py:84-95): [kinfo if kernel_cond] + [sqrt(sigma) if noise_cond], constants per
        # sample except the sigma map of noise_avg=False ----
        kc = self.kc
        n_cst = kc + (self.sc_extra if self.noise_avg else 0)
        extra = None
        if n_cst:
            extra = self._buf("sr.extra", (N, n_cst), f32)
            if kc:
                extra[:, :kc].copy_(kinfo)
            if n_cst > kc:
                extra[:, kc:].copy_(sigma.view(N, sc))
        sqrt_mask = (((1 << self.sc_extra) - 1) << kc) if self.sc_extra else 0
        src = None
        if self.spatial_extra and self.extra_mode != "null":
            Hh, Ww = h * sf, w * sf
            mod = 2 ** (self.depth - 1)
            src = ops.ExtraSource(extra, sigma, sf, sqrt_mask, Hh, Ww, (Hh + mod - 1) // mod * mod,
                                  (Ww + mod - 1) // mod * mod)
        mu = self._rnet_forward(S, x, sf, extra, src, save, sqrt_mask)
        if save:
            Hh, Ww, Hp, Wp, dims = S["rnet_dims"]
            S["x"], S["sigma"], S["kinfo"], S["extra"] = x, sigma, kinfo, extra
            S["src"] = src
            S["shape"] = (N, C, h, w, sf, Hh, Ww, Hp, Wp, dims, kh, kw, sqrt_mask)
            self._commit(S)
        return mu, kinfo, sigma

    def _sft_block_bwd(self, ii, b, c1, c2, gX, shape, dm, dd):
        """Backward of one SFT-modulated AttResBlock (AttResUNet.py:48-60); returns the gradient w.r.t. its input."""
        S = self._saved_A
        N, hh, ww, c = shape
        tag = f"d{ii}.{b}"
        m1 = S["sft"][(ii, b, "sft1")][0]
        m2 = S["sft"][(ii, b, "sft2")][0]
        self._wgrad(c2, gX, S[tag + ".b"], VK_CONV3X3_S1)
        G2 = self._buf(f"g.sr.{tag}.G2", (N, hh, ww, c))
        self._dgrad(gX, c2, VK_CONV3X3_S1, c, ldo=c, mask=S[tag + ".b"], out1=G2, alpha=0.2)
        gF1 = self._buf(f"g.sr.{tag}.F1", (N, hh, ww, c))
        det_ws = self._det_scratch(ops.sft_bwd_det_ws_floats(N, hh * ww, c)) if self.deterministic else None
        ops.sft_bwd(G2, S[tag + ".f1"], m2, gF1, dm[(ii, b, "sft2")], dd[(ii, b, "sft2")], dtype=self.dtype, c=c,
                    det_ws=det_ws)
        self._wgrad(c1, gF1, S[tag + ".a"], VK_CONV3X3_S1)
        G1 = self._buf(f"g.sr.{tag}.G1", (N, hh, ww, c))
        self._dgrad(gF1, c1, VK_CONV3X3_S1, c, ldo=c, mask=S[tag + ".a"], out1=G1, alpha=0.2)
        gXp = self._buf(f"g.sr.{tag}.X", (N, hh, ww, c))
        ops.sft_bwd(G1, S[tag + ".x"], m1, gXp, dm[(ii, b, "sft1")], dd[(ii, b, "sft1")], dtype=self.dtype, c=c, resid=gX,
                    det_ws=det_ws)
        return gXp

    def _det_scratch(self, floats: int):
        """fp32 scratch of the deterministic forms of the small per-sample kernels (slots of partial sums that a second
        launch adds in a fixed order); grown on demand, kept for the life of the engine (CUDA graphs capture its address),
        used on the main stream only."""
        buf = getattr(self, "_det_ws_buf", None)
        if buf is None or buf.numel() < floats:
            if buf is not None:
                self._det_ws_old = getattr(self, "_det_ws_old", []) + [buf]      # a captured graph may still point at it
            buf = self._det_ws_buf = torch.empty(max(floats, 1 << 16), device=self.flat_params.device, dtype=torch.float32)
        return buf

    def _sft_block_bwd_spatial(self, ii, b, c1, c2, gX, shape, src, d_cst, d_map):
        """Backward of one AttResBlock whose AttLayers run per pixel (vk_sft_apply_bwd + four 1x1 weight-gradient GEMMs
        per AttLayer); returns the gradient w.r.t. the block input."""
        S = self._saved_A
        N, hh, ww, c = shape
        tag = f"d{ii}.{b}"
        blk = self.net.RNet.down_path[ii].body[b]
        self._wgrad(c2, gX, S[tag + ".b"], VK_CONV3X3_S1)
        G2 = self._buf(f"g.sr.{tag}.G2", (N, hh, ww, c))
        self._dgrad(gX, c2, VK_CONV3X3_S1, c, ldo=c, mask=S[tag + ".b"], out1=G2, alpha=0.2)
        gF1 = self._buf(f"g.sr.{tag}.F1", (N, hh, ww, c))
        ops.sft_apply_bwd(G2, S[tag + ".f1"], gF1, blk.sft2, src, self.grad_view, dtype=self.dtype, c=c, d_cst=d_cst,
                          d_map=d_map)
        self._wgrad(c1, gF1, S[tag + ".a"], VK_CONV3X3_S1)
        G1 = self._buf(f"g.sr.{tag}.G1", (N, hh, ww, c))
        self._dgrad(gF1, c1, VK_CONV3X3_S1, c, ldo=c, mask=S[tag + ".a"], out1=G1, alpha=0.2)
        gXp = self._buf(f"g.sr.{tag}.X", (N, hh, ww, c))
        ops.sft_apply_bwd(G1, S[tag + ".x"], gXp, blk.sft1, src, self.grad_view, dtype=self.dtype, c=c, resid=gX,
                          d_cst=d_cst, d_map=d_map)
        return gXp

    def _rnet_backward(self, S, g_mu, N, C, sqrt_mask):
        """Backward of _rnet_forward for every extra_mode: parameter gradients of RNet (convs through the dgrad / wgrad
        kernels, AttLayers per sample or per pixel) and the gradient w.r.t. the conditioning values.
        Returns (d_cst [N, n_cst] fp32 or None, d_map like src.map or None)."""
        dt, dev, f32 = self.dtype, g_mu.device, torch.float32
        cp = lambda c: ops.chan_pad(c, dt)
        nf = self.n_feat
        Hh, Ww, Hp, Wp, dims = S["rnet_dims"]
        src = S.get("src")
        extra = S.get("extra")
        n_cst = 0 if extra is None else extra.shape[1]
        d_cst = torch.zeros(N, n_cst, device=dev, dtype=f32) if n_cst else None
        d_map = torch.zeros_like(src.map) if (src is not None and src.map is not None) else None
        sft_const = self.use_sft and src is None
        sft_spatial = self.use_sft and src is not None
        G = self._buf("g.sr.mu", (N, Hp, Wp, cp(C)))
        ops.pack_grad(g_mu.contiguous().float(), G, dtype=dt)
        self._wgrad(self.tail, G, S["tail.x"], VK_CONV3X3_S1)
        hh, ww = dims[0]
        gX = self._buf("g.sr.tail.X", (N, hh, ww, nf[0]))
        self._dgrad(G, self.tail, VK_CONV3X3_S1, nf[0], ldo=nf[0], out1=gX)
        g_bridge = {}
        for k in reversed(range(len(self.up))):
            us, res = self.up[k]
            lvl = self.depth - 2 - k
            c = nf[lvl]
            hh, ww = dims[lvl]
            for b in reversed(range(len(res))):
                gX = self._resblock_bwd(f"u{k}.{b}", res[b][0], res[b][1], gX, (N, hh, ww, c))
            g_bridge[lvl] = gX
            self._wgrad(us, S[f"u{k}.x"], gX, VK_CONVT2X2_S2)
            ops.channel_sum(gX, c, self.grad_view(us.bias), dtype=dt, ws=self._csum_ws)
            hl, wl = dims[lvl + 1]
            gXl = self._buf(f"g.sr.u{k}.low", (N, hl, wl, nf[lvl + 1]))
            self._dgrad(gX, us, VK_CONV2X2_S2, nf[lvl + 1], ldo=nf[lvl + 1], out1=gXl)
            gX = gXl
        dm = dd = None
        if sft_const:
            tables = self._tables[("sft_tables", N, self._flat_key)]
            sft_descs, sft_n, sft_maxc, dmd, sft_grads = tables[1], tables[2], tables[3], tables[4], tables[5]
            sft_grads.zero_()
            dm = {k: v[0] for k, v in dmd.items()}
            dd = {k: v[1] for k, v in dmd.items()}
        for ii in reversed(range(self.depth)):
            res, ds = self.down[ii]
            c = nf[ii]
            hh, ww = dims[ii]
            if ds is not None:
                self._wgrad(ds, gX, S[f"d{ii}.xlast"], VK_CONV3X3_S2)
                gXf = self._buf(f"g.sr.d{ii}.ds", (N, hh, ww, c))
                self._dgrad(gX, ds, VK_CONV3X3_S2_DGRAD, c, ldo=c, resid=g_bridge[ii], out1=gXf, out_hw=(hh, ww))
                gX = gXf
            for b in reversed(range(len(res))):
                if sft_const:
                    gX = self._sft_block_bwd(ii, b, res[b][0], res[b][1], gX, (N, hh, ww, c), dm, dd)
                elif sft_spatial:
                    gX = self._sft_block_bwd_spatial(ii, b, res[b][0], res[b][1], gX, (N, hh, ww, c), src, d_cst, d_map)
                else:
                    gX = self._resblock_bwd(f"d{ii}.{b}", res[b][0], res[b][1], gX, (N, hh, ww, c))
        # head conv: weight gradient, and the gradient w.r.t. its conditioning channels
        self._wgrad(self.head, gX, S["r0"], VK_CONV3X3_S1)
        if sft_const:
            # SFT MLPs: (dmul, dadd) -> their 1x1 convs and the conditioning values, every AttLayer in one launch
            if self.deterministic:
                n_params = tables[7]
                ops.sft_mlp_bwd_batched(sft_descs, sft_n, sft_maxc, extra, d_cst, sqrt_mask=sqrt_mask,
                                        det_ws=self._det_scratch(N * n_params + sft_n * N * extra.shape[1]),
                                        params_per_sample=n_params)
            else:
                ops.sft_mlp_bwd_batched(sft_descs, sft_n, sft_maxc, extra, d_cst, sqrt_mask=sqrt_mask)
        if self.head_extra:
            cin0 = C + self.head_extra
            gR0 = self._buf("g.sr.r0", (N, Hp, Wp, cp(cin0)))
            self._dgrad(gX, self.head, VK_CONV3X3_S1, cin0, ldo=cp(cin0), out1=gR0)
            if src is not None:
                ops.extra_head_grad(gR0, C, src, dtype=dt, d_cst=d_cst, d_map=d_map)
            else:
                E = n_cst
                kc = self.kc
                hsum = torch.zeros(N, cin0, device=dev, dtype=f32)
                if self.deterministic:                    # per sample through the atomic-free, block-ordered form
                    for n in range(N):
                        ops.channel_sum(gR0[n], cin0, hsum[n], dtype=dt, ws=self._csum_ws)
                else:
                    ops.channel_sum_batched(gR0, cin0, hsum, dtype=dt)
                # the head saw [kinfo, sqrt(sigma)] as constant planes: chain the sqrt for the variance channels
                hext = hsum[:, C:C + E].clone()
                if E > kc:
                    hext[:, kc:] = hext[:, kc:] * 0.5 / extra[:, kc:].sqrt().clamp_min(1e-20)
                d_cst += hext
        return d_cst, d_map

    def _snet_backward_map(self, S, g_total, N, h, w, prefix):
        """SNet backward for a per-pixel variance output sigma = exp(clamp(SNet(x))) (VIRNet.py:42-43,81): g_total is
        dL/dsigma [N, sc, h, w] fp32 (loss gradient + conditioning gradient)."""
        dt = self.dtype
        cp = lambda c: ops.chan_pad(c, dt)
        sc = self.sigma_chn
        GS = self._buf(prefix + "g.sig", (N, h, w, cp(sc)))
        ops.sigma_head_bwd(S["sigma"], g_total.contiguous(), None, 0, GS, dtype=dt, log_lo=SNET_LOG_MIN, log_hi=SNET_LOG_MAX)
        g = GS
        for i in reversed(range(len(self.s_layers))):
            ly = self.s_layers[i]
            inp = S["xs"] if i == 0 else S[f"s{i - 1}"]
            self._wgrad(ly, g, inp, VK_CONV3X3_S1)
            if i > 0:
                gn = self._buf(f"{prefix}g.s{i - 1}", (N, h, w, cp(ly.cin)))
                self._dgrad(g, ly, VK_CONV3X3_S1, ly.cin, ldo=cp(ly.cin), mask=inp, out1=gn, alpha=0.25)
                g = gn

    def _begin_backward(self, dev):
        self._begin_wgrads()
        if self.wgrad_side_stream and self._wg_stream is None:
            self._wg_stream = torch.cuda.Stream(device=dev)
            self._wg_events = [torch.cuda.Event() for _ in range(8)]
        if not self.wgrad_side_stream:
            self._wg_stream = None

    def _end_backward(self, A):
        if self._wg_stream is not None:
            torch.cuda.current_stream().wait_stream(self._wg_stream)
        self._unpack_all()
        self._saved_A = None
        A["set"]["owner"] = None

    def _backward_general_denoise(self, A, g_mu, g_sigma):
        """VIRAttResUNet with extra_mode 'Down' / 'Both': RNet modulated by the per-pixel sigma map."""
        N, C, H, W, Hp, Wp, _ = A["shape"]
        dev = A["x"].device
        self._begin_backward(dev)
        g_total = torch.zeros_like(A["sigma"])
        if g_sigma is not None:
            g_total += g_sigma.float()
        if g_mu is not None:
            _, d_map = self._rnet_backward(A, g_mu, N, C, (1 << self.sigma_chn) - 1)
            g_total += d_map
        self._snet_backward_map(A, g_total, N, H, W, "")
        self._end_backward(A)

    def backward_sr(self, g_mu, g_kinfo, g_sigma, gen: Optional[int] = None):
        """Accumulates the gradients of every parameter (SNet, KNet, RNet incl. the SFT MLPs) into flat_grads."""
        S = self._saved_A = self._take_saved(gen)
        if "kinfo" not in S:
            raise _l.VkError("backward_sr called without a saved super-resolution forward")
        N, C, h, w, sf, Hh, Ww, Hp, Wp, dims, kh, kw, sqrt_mask = S["shape"]
        dt, dev, f32 = self.dtype, S["x"].device, torch.float32
        cp = lambda c: ops.chan_pad(c, dt)
        net = self.net
        kcn, sc = self.k_tail.cout, self.sigma_chn         # outputs of KNet / SNet (always produced, VIRNet.py:81-82)
        kc = self.kc                                       # conditioning channels taken from kinfo (0 without kernel_cond)
        spatial = S.get("src") is not None                 # per-pixel sigma map (noise_avg=False)
        self._begin_backward(dev)
        d_cst = d_map = None
        if g_mu is not None:
            d_cst, d_map = self._rnet_backward(S, g_mu, N, C, sqrt_mask)
        gk = torch.zeros(N, kcn, device=dev, dtype=f32)
        if kc and d_cst is not None:
            gk += d_cst[:, :kc]
        if g_kinfo is not None:
            gk += g_kinfo.reshape(N, kcn).float()
        if spatial:
            gs_map = torch.zeros_like(S["sigma"])
            if d_map is not None:
                gs_map += d_map
            if g_sigma is not None:
                gs_map += g_sigma.float()
        else:
            gs = torch.zeros(N, sc, device=dev, dtype=f32)
            if d_cst is not None and d_cst.shape[1] > kc:
                gs += d_cst[:, kc:]
            if g_sigma is not None:
                gs += g_sigma.reshape(N, sc).float()
        # ---- KNet ----
        knet = net.KNet
        nfk = knet.head.out_channels
        Gt = self._buf("g.sr.k.tail", (N, kh, kw, cp(kcn)))
        ops.gap_head_bwd(gk.contiguous(), S["kinfo"], Gt, dtype=dt, c=kcn, exp_mask=0b011, tanh_mask=1 << (kcn - 1),
                         lo=KNET_LOG_MIN, hi=KNET_LOG_MAX)
        self._wgrad(self.k_tail, Gt, S["k.hlast"], VK_CONV3X3_S1)
        gH = self._buf("g.sr.k.h", (N, kh, kw, cp(nfk)))
        self._dgrad(Gt, self.k_tail, VK_CONV3X3_S1, nfk, ldo=cp(nfk), out1=gH)
        for b in reversed(range(len(self.k_blocks))):
            c1, c2, ca = self.k_blocks[b]
            dF = self._buf(f"g.sr.k{b}.df", (N, kh, kw, cp(nfk)))
            ca_ws = None
            if self.deterministic:
                r = ca.body[0].weight.shape[0]
                ca_ws = self._det_scratch(N * (2 * r * nfk + r + nfk))
            ops.ca_layer_bwd(gH, S[f"k{b}.f"], ca, dF, self.grad_view, dtype=dt, c=nfk, det_ws=ca_ws)
            self._wgrad(c2, dF, S[f"k{b}.a"], VK_CONV3X3_S1)
            gA = self._buf(f"g.sr.k{b}.ga", (N, kh, kw, cp(nfk)))
            self._dgrad(dF, c2, VK_CONV3X3_S1, nfk, ldo=cp(nfk), mask=S[f"k{b}.a"], out1=gA, alpha=0.2)
            self._wgrad(c1, gA, S[f"k{b}.h"], VK_CONV3X3_S1)
            gHn = self._buf(f"g.sr.k{b}.gh", (N, kh, kw, cp(nfk)))
            self._dgrad(gA, c1, VK_CONV3X3_S1, nfk, ldo=cp(nfk), resid=gH, out1=gHn)
            gH = gHn
        ops.knet_head_wgrad(S["x"], gH, self.grad_view(knet.head.weight), dtype=dt,
                            det_ws=self._det_scratch(N * knet.head.weight.numel()) if self.deterministic else None)
        # ---- SNet: per-pixel map head, or the global-average head ----
        if spatial:
            self._snet_backward_map(S, gs_map, N, h, w, "sr.")
        else:
            GS = self._buf("g.sr.sig", (N, h, w, cp(sc)))
            ops.gap_head_bwd(gs.contiguous(), S["sigma"].reshape(N, sc), GS, dtype=dt, c=sc, exp_mask=(1 << sc) - 1,
                             lo=SNET_LOG_MIN, hi=SNET_LOG_MAX)
            g = GS
            nS = len(self.s_layers)
            for i in reversed(range(nS)):
                ly = self.s_layers[i]
                inp = S["xs"] if i == 0 else S[f"s{i - 1}"]
                self._wgrad(ly, g, inp, VK_CONV3X3_S1)
                if i > 0:
                    gn = self._buf(f"g.sr.s{i - 1}", (N, h, w, cp(ly.cin)))
                    self._dgrad(g, ly, VK_CONV3X3_S1, ly.cin, ldo=cp(ly.cin), mask=inp, out1=gn, alpha=0.25)
                    g = gn
        if self._wg_stream is not None:
            torch.cuda.current_stream().wait_stream(self._wg_stream)
        self._unpack_all()
        self._saved_A = None
        S["set"]["owner"] = None
