"""ctypes binding of libvirnet_sm100.so (the C ABI declared in include/virnet_b200.h).

The shared library is built in-tree by ``__graft_entry__.build()`` (or ``make -C
virnet_b200/csrc``).  There is no fallback: if the library is missing, or a call
returns an error, we raise — the product path never drops to a CPU / eager
implementation.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libvirnet_sm100.so"

VK_BF16, VK_TF32 = 0, 1
VK_CONV3X3_S1, VK_CONV3X3_S2, VK_CONVT2X2_S2, VK_CONV1X1, VK_CONV2X2_S2, VK_CONV3X3_S2_DGRAD = 0, 1, 2, 3, 4, 5
VK_EPI_STD, VK_EPI_NCHW_F32 = 0, 1


class VkError(RuntimeError):
    pass


class vk_conv_args(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("kind", C.c_int32),
        ("x", C.c_void_p),
        ("n", C.c_int32), ("ih", C.c_int32), ("iw", C.c_int32), ("ldx", C.c_int32),
        ("w", C.c_void_p),
        ("wrows", C.c_int32),
        ("bias", C.c_void_p),
        ("cout", C.c_int32), ("ldo", C.c_int32), ("epi", C.c_int32),
        ("resid", C.c_void_p), ("mask", C.c_void_p), ("out1", C.c_void_p), ("out2", C.c_void_p),
        ("alpha", C.c_float),
        ("round_out2", C.c_int32), ("act_expclamp", C.c_int32),
        ("clamp_lo", C.c_float), ("clamp_hi", C.c_float),
        ("crop_h", C.c_int32), ("crop_w", C.c_int32),
        ("out_h", C.c_int32), ("out_w", C.c_int32),
        ("sft_mul", C.c_void_p), ("sft_add", C.c_void_p), ("sft_ld", C.c_int32), ("pad_", C.c_int32),
        ("force_tiles_per_cta", C.c_int32), ("force_chunk_bytes", C.c_int32),
        ("force_stages", C.c_int32), ("force_tw", C.c_int32),
        ("force_impl", C.c_int32), ("force_nt", C.c_int32),
        ("cta_timing", C.c_void_p),
    ]


class vk_wgrad_args(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("kind", C.c_int32),
        ("a", C.c_void_p),
        ("n", C.c_int32), ("gh", C.c_int32), ("gw", C.c_int32), ("lda", C.c_int32), ("m_valid", C.c_int32),
        ("b", C.c_void_p),
        ("bh", C.c_int32), ("bw", C.c_int32), ("ldb", C.c_int32), ("n_valid", C.c_int32),
        ("dw", C.c_void_p), ("dbias", C.c_void_p),
        ("force_ksplit", C.c_int32), ("force_k_rows", C.c_int32), ("force_stages", C.c_int32),
        ("max_slices", C.c_int32), ("partials", C.c_void_p), ("dbias_partials", C.c_void_p),
        ("swapped", C.c_int32), ("pad_", C.c_int32),
    ]


class vk_elbo_sisr_args(C.Structure):
    _fields_ = ([(k, C.c_void_p) for k in ("mu", "im_hr", "im_lr", "sigma_est", "kinfo_est", "kinfo_gt", "prior_mean",
                                           "prior_logmean", "gamma_draw", "rho_draw", "z_draw", "rh", "rw", "d_mu",
                                           "d_sigma", "d_kinfo", "kernel", "terms", "ws")] +
                [("ws_bytes", C.c_int64)] +
                [(k, C.c_int32) for k in ("n", "c", "H", "W", "h", "w", "k_size")] +
                [(k, C.c_float) for k in ("center", "alpha0", "digamma_am1", "kappa0", "r2", "eps2", "pk0", "pk1")])


class vk_sft_desc(C.Structure):
    _fields_ = ([(k, C.c_void_p) for k in ("w1", "b1", "w2", "b2", "wm", "bm", "wa", "ba", "gw1", "gb1", "gw2", "gb2",
                                           "gwm", "gbm", "gwa", "gba", "mul", "add", "dmul", "dadd")] +
                [(k, C.c_int32) for k in ("c1", "c2", "c", "pad_")])


class vk_extra_src(C.Structure):
    _fields_ = [("cst", C.c_void_p), ("map", C.c_void_p),
                ("ec", C.c_int32), ("em", C.c_int32), ("eh", C.c_int32), ("ew", C.c_int32), ("esf", C.c_int32),
                ("sqrt_mask", C.c_uint32),
                ("hh", C.c_int32), ("ww", C.c_int32), ("hp", C.c_int32), ("wp", C.c_int32)]


class vk_sft_apply_args(C.Structure):
    _fields_ = ([(k, C.c_int32) for k in ("dtype", "n", "h", "w", "c", "ld", "c1", "c2")] +
                [(k, C.c_void_p) for k in ("x", "out", "w1", "b1", "w2", "b2", "wm", "bm", "wa", "ba")] +
                [("extra", vk_extra_src), ("alpha", C.c_float), ("round_tf32", C.c_int32)])


class vk_sft_apply_bwd_args(C.Structure):
    _fields_ = ([(k, C.c_int32) for k in ("dtype", "n", "h", "w", "c", "ld", "c1", "c2", "ld1", "ld2")] +
                [(k, C.c_void_p) for k in ("g", "x", "resid", "gx", "dm", "f2", "dq2", "f1", "dq1", "ev", "w1", "b1", "w2",
                                           "b2", "wm", "bm", "wa", "ba", "d_cst", "d_map")] +
                [("extra", vk_extra_src), ("alpha", C.c_float), ("pad_", C.c_int32)])


class vk_pack_desc(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p),
                ("dim0", C.c_int32), ("dim1", C.c_int32), ("taps", C.c_int32),
                ("rows", C.c_int32), ("ld", C.c_int32), ("dst_taps", C.c_int32),
                ("mode", C.c_int32), ("pad_", C.c_int32)]


class vk_unpack_desc(C.Structure):
    _fields_ = [("ws", C.c_void_p), ("out", C.c_void_p), ("taps", C.c_int32), ("mn", C.c_int32),
                ("nslices", C.c_int32), ("pad_", C.c_int32), ("slice_stride", C.c_int64)]


class vk_adam_group(C.Structure):
    _fields_ = [("begin", C.c_int64), ("end", C.c_int64), ("max_norm", C.c_float), ("pad_", C.c_int32)]


_lib = None

# every symbol include/virnet_b200.h declares: (name, restype, argtypes)
_SIGNATURES = {
    "vk_conv_igemm": (C.c_int, [C.POINTER(vk_conv_args), C.c_void_p]),
    "vk_conv_wgrad": (C.c_int, [C.POINTER(vk_wgrad_args), C.c_void_p]),
    "vk_conv_wgrad_plan": (C.c_int, [C.POINTER(vk_wgrad_args), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "vk_wgrad_unpack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "vk_wgrad_unpack_batched": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_void_p]),
    "vk_sizeof_wgrad_args": (C.c_uint32, []),
    "vk_pack_input": (C.c_int, [C.c_int32, C.c_void_p] + [C.c_int32] * 5 + [C.c_void_p] + [C.c_int32] * 6
                      + [C.c_void_p] + [C.c_int32] * 3 + [C.c_void_p]),
    "vk_pack_grad": (C.c_int, [C.c_int32, C.c_void_p] + [C.c_int32] * 4 + [C.c_void_p] + [C.c_int32] * 3 + [C.c_void_p]),
    "vk_sigma_head_bwd": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int32] * 4 + [C.c_void_p]
                          + [C.c_int32] * 5 + [C.c_float, C.c_float, C.c_void_p]),
    "vk_elbo_denoise": (C.c_int, [C.c_void_p] * 5 + [C.c_float] + [C.c_int32] * 5 + [C.c_float] * 4 + [C.c_void_p] * 3
                        + [C.c_int32, C.c_void_p, C.c_void_p]),
    "vk_pack_weights": (C.c_int, [C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_void_p]),
    "vk_channel_sum": (C.c_int, [C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64,
                                 C.c_void_p]),
    "vk_channel_sum_batched": (C.c_int, [C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                         C.c_void_p]),
    "vk_adam_clip_step": (C.c_int, [C.c_void_p] * 5 + [C.c_int32, C.c_int64, C.c_void_p, C.c_int32] + [C.c_float] * 5
                          + [C.c_int32, C.c_void_p, C.c_void_p]),
    "vk_adam_clip_step_dev": (C.c_int, [C.c_void_p] * 5 + [C.c_int32, C.c_int64, C.c_void_p, C.c_int32] + [C.c_float] * 4
                              + [C.c_void_p, C.c_void_p, C.c_void_p]),
    "vk_knet_head": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int32] * 6 + [C.c_void_p]),
    "vk_ca_layer": (C.c_int, [C.c_int32] + [C.c_void_p] * 7 + [C.c_int32] * 5 + [C.c_float, C.c_void_p]),
    "vk_gap_head": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_uint32, C.c_uint32, C.c_float, C.c_float,
                              C.c_void_p, C.c_void_p]),
    "vk_sft_mlp": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                             C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_float,
                             C.c_void_p, C.c_void_p, C.c_void_p]),
    "vk_sft_bwd": (C.c_int, [C.c_int32] + [C.c_void_p] * 7 + [C.c_int32] * 4 + [C.c_void_p]),
    "vk_sft_bwd_det_ws_floats": (C.c_int64, [C.c_int32] * 3),
    "vk_sft_bwd_det": (C.c_int, [C.c_int32] + [C.c_void_p] * 7 + [C.c_int32] * 4 + [C.c_void_p, C.c_int64, C.c_void_p]),
    "vk_sft_mlp_bwd": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int32,
                                 C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_int32, C.c_float] + [C.c_void_p] * 12),
    "vk_ca_layer_bwd": (C.c_int, [C.c_int32] + [C.c_void_p] * 11 + [C.c_int32] * 5 + [C.c_float, C.c_void_p]),
    "vk_ca_layer_bwd_det": (C.c_int, [C.c_int32] + [C.c_void_p] * 11 + [C.c_int32] * 5 + [C.c_float, C.c_void_p, C.c_int64,
                                                                                              C.c_void_p]),
    "vk_gap_head_bwd": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_uint32,
                                  C.c_uint32, C.c_float, C.c_float, C.c_void_p, C.c_int32, C.c_void_p]),
    "vk_knet_head_wgrad": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int32] * 6 + [C.c_void_p]),
    "vk_knet_head_wgrad_det": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int32] * 6 +
                               [C.c_void_p, C.c_int64, C.c_void_p]),
    "vk_upsample_nearest": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 5 + [C.c_void_p]),
    "vk_sft_mlp_batched": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_uint32,
                                     C.c_float, C.c_void_p]),
    "vk_sft_mlp_bwd_batched": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_uint32,
                                         C.c_float, C.c_void_p, C.c_void_p]),
    "vk_sft_mlp_bwd_batched_det": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                             C.c_uint32, C.c_float, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                             C.c_void_p]),
    "vk_sizeof_sft_desc": (C.c_uint32, []),
    "vk_sft_apply": (C.c_int, [C.POINTER(vk_sft_apply_args), C.c_void_p]),
    "vk_sizeof_sft_apply_args": (C.c_uint32, []),
    "vk_sft_apply_bwd": (C.c_int, [C.POINTER(vk_sft_apply_bwd_args), C.c_void_p]),
    "vk_sizeof_sft_apply_bwd_args": (C.c_uint32, []),
    "vk_extra_head_grad": (C.c_int, [C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(vk_extra_src),
                                     C.c_void_p, C.c_void_p, C.c_void_p]),
    "vk_pack_input_mixed": (C.c_int, [C.c_int32, C.c_void_p] + [C.c_int32] * 5 + [C.POINTER(vk_extra_src), C.c_void_p,
                                                                                  C.c_int32, C.c_void_p]),
    "vk_synth_denoise": (C.c_int, [C.c_void_p] * 4 + [C.c_int32] * 4 + [C.c_void_p] * 4),
    "vk_noise_estimate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                    C.c_int32, C.c_float, C.c_void_p]),
    "vk_aug8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "vk_aug8_merge": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "vk_to_u8": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 4 + [C.c_void_p]),
    "vk_psnr_u8": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 5 + [C.c_void_p, C.c_void_p]),
    "vk_ssim_u8": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 5 + [C.c_void_p, C.c_void_p, C.c_void_p]),
    "vk_mixup": (C.c_int, [C.c_void_p] * 6 + [C.c_int32, C.c_int64, C.c_void_p]),
    "vk_sisr_degrade_ws_bytes": (C.c_int64, [C.c_int32] * 5),
    "vk_sisr_degrade": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32] + [C.c_void_p] * 7 + [C.c_int64] + [C.c_int32] * 6 +
                        [C.c_void_p]),
    "vk_elbo_sisr_ws_bytes": (C.c_int64, [C.c_int32] * 7),
    "vk_sizeof_elbo_sisr_args": (C.c_uint32, []),
    "vk_elbo_sisr": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vk_sizeof_conv_args": (C.c_uint32, []),
    "vk_version": (C.c_char_p, []),
    "vk_launch_count": (C.c_uint64, []),
}


def exported_symbols():
    return sorted(_SIGNATURES)


def load():
    """Load the shared library (once).  Raises VkError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise VkError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C virnet_b200/csrc` (there is no CPU fallback)")
    lib = C.CDLL(os.fspath(LIB_PATH))
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if (lib.vk_sizeof_conv_args() != C.sizeof(vk_conv_args)
            or lib.vk_sizeof_wgrad_args() != C.sizeof(vk_wgrad_args)
            or lib.vk_sizeof_elbo_sisr_args() != C.sizeof(vk_elbo_sisr_args)
            or lib.vk_sizeof_sft_desc() != C.sizeof(vk_sft_desc)
            or lib.vk_sizeof_sft_apply_args() != C.sizeof(vk_sft_apply_args)
            or lib.vk_sizeof_sft_apply_bwd_args() != C.sizeof(vk_sft_apply_bwd_args)):
        raise VkError("argument struct layout mismatch between lib.py and the built library; rebuild")
    _lib = lib
    return lib


def check(code: int, what: str):
    if code == 0:
        return
    if code < 0:
        names = {-1: "VK_E_BADARG", -2: "VK_E_UNSUPPORTED", -3: "VK_E_NODRIVER"}
        raise VkError(f"{what}: {names.get(code, code)}")
    raise VkError(f"{what}: cudaError {code}")


def launch_count() -> int:
    return int(load().vk_launch_count())
