"""ctypes binding of libmsp_b200.so (the C ABI declared in include/msp_b200.h).

There is deliberately no fallback: if the shared object is missing the import of any product
module raises, and every call that returns a negative status raises RuntimeError carrying
msp_last_error().  (The reference swallows per-batch exceptions, train_model.py:122-130, so the
text is also printed to stderr to stay visible.)
"""
from __future__ import annotations

import ctypes as C
import os
import sys
from pathlib import Path

import torch  # noqa: F401  (first: libmsp_b200.so links the CUDA runtime dynamically and must share torch's instance)

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libmsp_b200.so"

P = C.c_void_p
I = C.c_int
LL = C.c_longlong
F = C.c_float
D = C.c_double


class ConvDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "N", "H", "W", "C", "x_cs", "Ho", "Wo", "K", "y_cs", "KH", "KW", "stride", "pad_t", "pad_l",
        "relu", "win_px", "Wp", "stat_rows")]


class UnpackItem(C.Structure):
    _fields_ = [("partials", C.c_void_p), ("dst", C.c_void_p), ("split_stride", C.c_longlong)] + \
               [(n, C.c_int32) for n in ("splits", "K", "C_true", "taps", "Cpad", "rowwin_KH", "rowwin_KW", "rowwin_cpp",
                                         "accumulate", "reserved")]


class BnActDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "N", "H", "W", "C", "x_cs", "y_cs", "act", "r_C", "r_cs", "r_stride")]


# name -> argtypes (all functions return int status unless listed in _RESTYPES)
SIGNATURES = {
    "msp_last_error": [],
    "msp_version": [],
    "msp_launch_count": [],
    "msp_conv_set_policy": [I, I],
    "msp_conv_last_kernel": [],
    "msp_pack_weights": [P, I, I, I, I, I, I, P, P, P],
    "msp_conv_fprop": [C.POINTER(ConvDesc), P, P, P, P, P, P, P],
    "msp_conv_dgrad": [C.POINTER(ConvDesc), P, P, P, I, P],
    "msp_conv_transpose_fprop": [C.POINTER(ConvDesc), P, P, P, I, P, P],
    "msp_conv_wgrad_splits": [C.POINTER(ConvDesc)],
    "msp_conv_wgrad": [C.POINTER(ConvDesc), P, P, P, P],
    "msp_unpack_wgrad": [C.POINTER(ConvDesc), P, I, P, P],
    "msp_unpack_wgrad_batched": [I, C.POINTER(UnpackItem), P],
    "msp_fold_upconv_weights": [P, I, I, P, P],
    "msp_unfold_upconv_wgrad": [P, P, P, P, I, I, P, I, P],
    "msp_upconv2x_fprop": [C.POINTER(ConvDesc), P, P, P, P, P],
    "msp_upconv2x_dgrad": [C.POINTER(ConvDesc), P, P, P, I, P],
    "msp_upconv2x_wgrad_splits": [C.POINTER(ConvDesc)],
    "msp_upconv2x_wgrad_class": [C.POINTER(ConvDesc), P, P, I, I, P, P],
    "msp_pack_weights_rowwin": [P, I, I, I, I, I, P, P],
    "msp_pack_weights_batched": [P, I, I, P],
    "msp_nchw_f32_to_rowwin_bf16": [P, I, I, I, I, I, I, I, P, P],
    "msp_nchw_bf16_to_rowwin_bf16": [P, I, I, I, I, I, I, I, P, P],
    "msp_nchw_f32_to_nhwc_bf16": [P, I, I, I, I, I, P, P],
    "msp_nhwc_bf16_to_nchw_f32": [P, I, I, I, I, I, P, P],
    "msp_nchw_f32_grad_to_nhwc_bf16": [P, I, I, I, I, I, P, P],
    "msp_bn_finalize": [P, P, I, D, F, F, P, P, P, P, I, I, P],
    "msp_reduce_rows": [P, I, I, P, I, P],
    "msp_bn_act_fwd": [C.POINTER(BnActDesc), P, P, P, P, P, P, P, P, P],
    "msp_bn_act_bwd_reduce": [C.POINTER(BnActDesc), P, P, P, P, P, P, P, P, P, P, P, I, P],
    "msp_bn_act_bwd_reduce_rows": [C.POINTER(BnActDesc), P, P, P, P, P, P, P, P, P, I, C.POINTER(C.c_int), P],
    "msp_bn_act_bwd_apply": [C.POINTER(BnActDesc), P, P, P, P, P, P, P, P, P, P, D, P, P, I, P],
    "msp_bn_eval_prepare": [P, I, F, P, P],
    "msp_maxpool_fwd": [P, I, I, I, I, I, I, I, I, P, P, I, I, I, P],
    "msp_maxpool_bwd": [P, P, I, I, I, I, I, I, I, I, I, I, P, I, I, P],
    "msp_upsample2x_fwd": [P, I, I, I, I, I, P, I, P],
    "msp_upsample2x_bwd": [P, I, I, I, I, I, P, I, P],
    "msp_upsample_bilinear2x_fwd": [P, I, I, I, I, I, P, I, P],
    "msp_upsample_bilinear2x_bwd": [P, I, I, I, I, I, P, I, P],
    "msp_avgpool_fwd": [P, I, I, I, I, P, P],
    "msp_avgpool_bwd": [P, I, I, I, P, I, P],
    "msp_copy_channels": [P, LL, I, I, P, I, P],
    "msp_channel_sum": [P, LL, I, I, P, P, I, P],
    "msp_add_relu_fwd": [P, P, LL, I, I, I, P, I, P],
    "msp_relu_bwd": [P, P, LL, I, I, I, P, I, P],
    "msp_add": [P, P, LL, I, I, I, P, I, P],
    "msp_gate_mul_fwd": [P, P, I, I, I, I, I, I, P, I, P],
    "msp_gate_mul_bwd": [P, P, P, I, I, I, I, I, I, I, P, I, I, P, I, P],
    "msp_final_conv_act_fwd": [P, I, I, I, I, I, P, P, I, I, P, P, P],
    "msp_final_conv_act_bwd": [P, I, I, I, I, I, P, I, I, P, P, P, I, P, P, P, I, P],
    "msp_dice_sums": [P, P, I, I, LL, I, I, I, P, P],
    "msp_dice_finalize": [P, I, I, I, F, P, P, P],
    "msp_dice_bwd": [P, P, I, I, LL, I, I, I, P, F, P, P, P],
    "msp_scale_to_float": [P, D, P, P],
    "msp_ce_prob_fwd_bwd": [P, P, I, I, LL, F, F, P, P, P, P],
    "msp_bce_fwd_bwd": [P, P, LL, I, F, P, P, P, P],
    "msp_softmax_ce_fwd_bwd": [P, P, I, I, F, F, P, P, P, P],
    "msp_softmax_ce_soft_fwd_bwd": [P, P, I, I, F, F, P, P, P, P],
    "msp_softmax_ce_spatial_fwd_bwd": [P, P, I, I, LL, F, F, P, P, P, P],
    "msp_confusion_binary": [P, P, I, I, I, LL, F, I, P, P],
    "msp_confusion_multiclass": [P, P, I, I, I, LL, P, P],
    "msp_topk_hits": [P, P, I, I, LL, I, P, P],
    "msp_rowpair_distances": [P, P, I, LL, I, P, P],
    "msp_triplet_hinge": [P, I, P, I, P, P],
    "msp_optim_sqnorm": [I, C.POINTER(C.c_void_p), C.POINTER(C.c_longlong), P, I, P],
    "msp_optim_norm": [P, P, P],
    "msp_optim_add_scalar": [I, C.POINTER(C.c_void_p), F, P],
    "msp_optim_clip": [I, C.POINTER(C.c_void_p), C.POINTER(C.c_longlong), P, F, P],
    "msp_optim_sgd": [I, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_longlong),
                      F, F, F, F, I, I, P],
    "msp_optim_adamw": [I, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                        C.POINTER(C.c_longlong), F, D, D, F, F, P, P],
    "msp_u8_to_f32_nchw": [P, I, I, LL, I, D, P, P],
    "msp_color_jitter": [P, I, I, LL, C.POINTER(C.c_int), F, F, F, F, C.POINTER(C.c_float), P, P, P],
    "msp_p2p_buffer_bytes": [I, I],
    "msp_p2p_alloc": [LL, C.POINTER(C.c_void_p), P],
    "msp_p2p_open": [P, C.POINTER(C.c_void_p)],
    "msp_p2p_close": [P],
    "msp_p2p_free": [P],
    "msp_p2p_allreduce_sum_f32": [P, I, I, I, I, C.POINTER(C.c_void_p), P, P],
    "msp_p2p_stats_exchange": [P, I, I, P, P, I, P, I, D, F, F, P, P, P, P, I, I, I, C.POINTER(C.c_void_p), P, P, P],
}
_RESTYPES = {"msp_last_error": C.c_char_p, "msp_conv_last_kernel": C.c_char_p, "msp_launch_count": C.c_longlong, "msp_p2p_buffer_bytes": C.c_longlong}


class MspError(RuntimeError):
    pass


def _load() -> C.CDLL:
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m medsegpretrainimagenet_b200.build` "
            "(there is no CPU / PyTorch fallback for the B200 hot path)")
    lib = C.CDLL(str(LIB_PATH))
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    return lib


lib = _load()


def last_error() -> str:
    return lib.msp_last_error().decode("utf-8", "replace")


def launch_count() -> int:
    return int(lib.msp_launch_count())


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = f"{what} failed (status {rc}): {last_error()}"
        if os.environ.get("MSP_QUIET_ERRORS") != "1":
            print("[msp_b200] " + msg, file=sys.stderr, flush=True)
        raise MspError(msg)


def call(name: str, *args) -> None:
    check(getattr(lib, name)(*args), name)
