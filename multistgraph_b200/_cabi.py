"""ctypes binding of include/matgcn.h.

The product path has no fallback: if ``csrc/libmatgcn.so`` is missing or a call fails,
the caller gets an exception, never a silent PyTorch/CPU path.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_int, c_longlong, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libmatgcn.so")
ABI_VERSION = 4
FLAG_EXACT = 0
FLAG_TF32 = 1
FLAG_BF16 = 2

_lib = None

_F = c_void_p  # device float*

_SIGNATURES = {
    "matgcn_abi_version": (c_int, []),
    "matgcn_last_error": (c_char_p, []),
    "matgcn_launch_count": (ctypes.c_ulonglong, []),
    "matgcn_tc_launch_count": (ctypes.c_ulonglong, []),
    "matgcn_propagate_fwd": (c_int, [_F, c_int, c_int, c_int, _F, c_int, _F, c_int, c_void_p]),
    "matgcn_debug_set_timeline": (c_int, [c_void_p]),
    "matgcn_debug_set_mode": (c_int, [c_int]),
    "matgcn_debug_set_timeline_skip": (c_int, [c_int]),
    "matgcn_gemm_debug": (c_int, [c_int, c_int, c_int, c_int, c_int, _F, c_int, _F, c_int, _F, c_int, c_int, c_int,
                                  c_void_p]),
    "matgcn_set_fused_tail": (c_int, [c_int]),
    "matgcn_set_recurrent_kernel": (c_int, [c_int]),
    "matgcn_set_dr_pass": (c_int, [c_int]),
    "matgcn_rec_timing": (c_int, [c_int]),
    "matgcn_rec_timing_read": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "matgcn_propagate_fwd_bf16": (c_int, [_F, c_int, c_int, c_int, _F, c_int, _F, c_void_p]),
    "matgcn_propagate_fwd_bf16_twin": (c_int, [_F, c_int, c_int, c_int, _F, c_int, _F, c_void_p]),
    "matgcn_gemm_debug_bf16": (c_int, [c_int, c_int, c_int, c_int, c_int, _F, c_int, _F, c_int, _F, c_int, c_int, c_void_p]),
    "matgcn_adaptive_adj_fwd": (c_int, [_F, _F, c_int, c_int, _F, c_int, c_void_p]),
    "matgcn_adaptive_adj_bwd": (c_int, [_F, _F, _F, _F, c_int, c_int, c_int, _F, _F, _F, c_void_p]),
    "matgcn_nodeweights_fwd": (c_int, [_F, _F, _F, _F, c_int, c_int, c_int, c_int, c_int, _F, _F, c_void_p]),
    "matgcn_nodeweights_fwd_ex": (c_int, [_F, _F, _F, _F, c_int, c_int, c_int, c_int, c_int, _F, _F, c_int, c_void_p]),
    "matgcn_nodeweights_bwd_ex": (c_int, [_F, _F, _F, _F, _F, _F, c_int, c_int, c_int, c_int, c_int, _F, _F, _F, _F, c_int,
                                          c_void_p]),
    "matgcn_nodeweights_bwd": (c_int, [_F, _F, _F, _F, _F, _F, c_int, c_int, c_int, c_int, c_int,
                                       _F, _F, _F, _F, c_void_p]),
    "matgcn_encoder_layer_fwd_ws_bytes": (c_size_t, [c_int] * 6),
    "matgcn_encoder_layer_bwd_ws_bytes": (c_size_t, [c_int] * 7),
    "matgcn_encoder_layer_y_offset": (c_size_t, [c_int] * 6),
    "matgcn_encoder_layer_y_tstride": (c_size_t, [c_int] * 6),
    "matgcn_encoder_layer_slot_offset": (c_size_t, [c_char_p] + [c_int] * 6),
    "matgcn_encoder_layer_fwd": (c_int, [c_int] * 7 + [_F, c_longlong, _F, _F, _F, _F, _F, _F, _F, _F, _F, _F,
                                                       _F, _F, c_int, c_void_p]),
    "matgcn_encoder_layer_bwd": (c_int, [c_int] * 8 + [_F, c_longlong, _F, _F, _F, _F, _F, _F, _F, _F,
                                                       _F, _F, _F, _F, _F, _F, _F, _F, _F, _F, _F, _F, c_int, c_void_p]),
    "matgcn_encoder_layer_chain_ok": (c_int, [c_int] * 8),
    "matgcn_encoder_layer_fwd_chained": (c_int, [c_int] * 7 + [_F, c_longlong, c_void_p, _F, _F, _F, _F, _F, _F, _F, _F, _F, _F,
                                                               _F, _F, c_int, c_void_p]),
    "matgcn_encoder_layer_bwd_chained": (c_int, [c_int] * 8 + [_F, c_longlong, _F, _F, _F, _F, _F, _F, _F, _F,
                                                               _F, _F, _F, _F, _F, _F, _F, _F, _F, _F, _F, _F, c_int, _F, c_void_p,
                                                               c_void_p]),
    "matgcn_dense_gru_layer_fwd_ws_bytes": (c_size_t, [c_int] * 5),
    "matgcn_dense_gru_layer_bwd_ws_bytes": (c_size_t, [c_int] * 5),
    "matgcn_dense_gru_layer_y_offset": (c_size_t, [c_int] * 5),
    "matgcn_dense_gru_layer_fwd": (c_int, [c_int] * 5 + [_F, c_longlong, _F, _F, _F, _F, _F, _F, c_int, c_void_p]),
    "matgcn_dense_gru_layer_bwd": (c_int, [c_int] * 5 + [_F, c_longlong, _F, c_longlong, _F, _F, _F, _F, _F, _F, _F, _F,
                                                         _F, _F, c_int, c_void_p]),
    "matgcn_grad_sumsq": (c_int, [_F, c_longlong, c_void_p, c_void_p]),
    "matgcn_adam_clip_step": (c_int, [_F, _F, _F, _F, c_longlong, c_void_p] + [ctypes.c_float] * 7
                              + [c_longlong, c_int, _F, c_void_p]),
    "matgcn_adam_clip_step_dev": (c_int, [_F, _F, _F, _F, c_longlong, c_void_p, ctypes.c_float, ctypes.c_float, c_void_p]
                                  + [ctypes.c_float] * 4 + [c_void_p, c_int, _F, c_void_p]),
    "matgcn_step_tick": (c_int, [c_void_p, c_void_p, c_void_p]),
    "matgcn_assemble_windows": (c_int, [_F, c_longlong, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_int, _F, _F,
                                        c_void_p, c_void_p]),
    "matgcn_head_dropout_scale": (ctypes.c_float, [ctypes.c_float]),
    "matgcn_head_fwd": (c_int, [_F, c_longlong, c_int, c_longlong, c_int, _F, _F, c_int, ctypes.c_float, ctypes.c_ulonglong, _F,
                                c_void_p]),
    "matgcn_head_bwd": (c_int, [_F, c_longlong, c_int, c_longlong, c_int, _F, c_int, ctypes.c_float, ctypes.c_ulonglong, _F, _F, _F,
                                _F, c_void_p]),
    "matgcn_head_fwd_dev": (c_int, [_F, c_longlong, c_int, c_longlong, c_int, _F, _F, c_int, ctypes.c_float, ctypes.c_ulonglong,
                                    c_void_p, _F, c_void_p]),
    "matgcn_head_bwd_dev": (c_int, [_F, c_longlong, c_int, c_longlong, c_int, _F, c_int, ctypes.c_float, ctypes.c_ulonglong,
                                    c_void_p, _F, _F, _F, _F, c_void_p]),
    "matgcn_head_dropout_mask": (c_int, [c_longlong, ctypes.c_float, ctypes.c_ulonglong, _F, c_void_p]),
    "matgcn_masked_mae_fwd": (c_int, [_F, _F, c_void_p, c_void_p, c_void_p, ctypes.c_float, ctypes.c_float, ctypes.c_float, c_void_p,
                                      _F, c_void_p]),
    "matgcn_masked_mae_bwd": (c_int, [_F, _F, c_void_p, c_void_p, c_void_p, ctypes.c_float, ctypes.c_float, ctypes.c_float, c_void_p,
                                      _F, _F, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


class MatgcnError(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes handle; raises if the library was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MatgcnError(
                "CUDA library %s is not built (run `python -m multistgraph_b200.build` or "
                "__graft_entry__.build()); there is no CPU fallback for the Multi-ATGCN encoder" % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        got = handle.matgcn_abi_version()
        if got != ABI_VERSION:
            raise MatgcnError("libmatgcn.so ABI %d != expected %d (stale build?)" % (got, ABI_VERSION))
        _lib = handle
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().matgcn_last_error()
        raise MatgcnError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))
