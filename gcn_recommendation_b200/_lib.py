"""ctypes binding of liblgcn_b200.so (include/lgcn.h).

There is NO CPU fallback: if the CUDA library is missing or a tensor is not a contiguous
CUDA tensor of the expected dtype, the call raises.
"""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LGCN_B200_LIB", os.path.join(_HERE, "liblgcn_b200.so"))

c_i32, c_i64, c_f32, c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p

SPMM_PLAIN, SPMM_ADD, SPMM_MEAN, SPMM_ADAM = 0, 1, 2, 3
SPMM_F_STREAM_HINTS = 1
SPMM_F_NO_RING = 2
SPMM_F_BIG_PATH = 4
SPMM_F_COLD_FIRST = 8
SPMM_F_FORCE_RING = 16
SPMM_F_NO_PREFETCH = 32
SPMM_F_ALT_X = 64
SPMM_F_ALT_LAYER0 = 128
SPMM_F_LONG_DONE = 256
BPR_GP_INCLUDES_GF, BPR_NO_GRAD = 1, 2
ABI_VERSION = 7


class SpmmArgs(ctypes.Structure):
    """struct lgcn_spmm_args (include/lgcn.h)."""
    _fields_ = [
        ("rowptr", c_vp), ("colval", c_vp), ("x", c_vp),
        ("n_rows", c_i64), ("d", c_i32), ("mode", c_i32),
        ("y", c_vp), ("addend", c_vp), ("layers", c_vp * 8), ("n_layers", c_i32),
        ("n_long", c_i32), ("long_row_ids", c_vp), ("long_rowptr", c_vp), ("long_colval", c_vp),
        ("long_seg_ptr", c_vp), ("seg_len", c_i32), ("n_seg", c_i32), ("seg_ws", c_vp),
        ("addend2", c_vp), ("p", c_vp), ("m", c_vp), ("v", c_vp), ("adam_scalars", c_vp),
        ("beta1", c_f32), ("beta2", c_f32), ("eps", c_f32), ("g_out", c_vp),
        ("flags", c_i32), ("x_rowflag", c_vp), ("addend_rowflag", c_vp), ("zero_row", c_vp),
        ("y_rowflag", c_vp),
        ("x_alt", c_vp), ("alt_begin", c_i64), ("alt_rows", c_i64),
        ("g_skip", c_vp), ("skip_begin", c_i64), ("skip_rows", c_i64),
        ("chunk_order", c_vp), ("long_done", c_vp),
    ]


_SIGNATURES = {
    "lgcn_abi_version": (ctypes.c_int, []),
    "lgcn_error_string": (ctypes.c_char_p, [ctypes.c_int]),
    "lgcn_csr_from_sorted_coo": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "lgcn_check_indices": (ctypes.c_int, [c_vp, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "lgcn_edge_weights": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "lgcn_spmm": (ctypes.c_int, [ctypes.POINTER(SpmmArgs), c_vp]),
    "lgcn_sizeof_spmm_args": (ctypes.c_size_t, []),
    "lgcn_spmm_chunk_rows": (ctypes.c_int, [c_i64, c_i32, c_i32]),
    "lgcn_spmm_launches": (ctypes.c_int, [c_i64, c_i32, c_i32, c_i32, ctypes.POINTER(ctypes.c_int32)]),
    "lgcn_spmm_kernel_name": (ctypes.c_int, [c_i64, c_i32, c_i32, c_i32, c_i32, ctypes.c_char_p,
                                             ctypes.c_size_t]),
    "lgcn_bpr_fused": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i64, c_f32,
                                      c_f32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "lgcn_bpr_partial": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i64, c_vp, c_vp]),
    "lgcn_bpr_apply": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i64, c_f32,
                                      c_f32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "lgcn_zero_rows": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i64, c_vp]),
    "lgcn_sample_bpr": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64, ctypes.c_uint64, c_vp, c_i64, c_vp,
                                       c_vp, c_vp, c_i64, c_vp]),
    "lgcn_adam_tick": (ctypes.c_int, [c_vp, c_vp, c_f32, c_f32, c_f32, c_vp]),
    "lgcn_adam": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_f32, c_f32, c_f32,
                                 c_vp]),
    "lgcn_fusion_proj_fwd": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_vp, c_vp]),
    "lgcn_fusion_force_simt": (None, [ctypes.c_int]),
    "lgcn_fusion_proj_bwd": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_vp,
                                            c_vp, c_vp, c_vp]),
    "lgcn_score_topk_workspace": (ctypes.c_size_t, [c_i64, c_i64, c_i32, c_i32]),
    "lgcn_score_topk": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i32, c_vp, c_vp, c_i32,
                                       c_vp, c_vp, c_vp, ctypes.c_size_t, c_vp]),
    "lgcn_score_tc_workspace": (ctypes.c_size_t, [c_i64, c_i64, c_i32]),
    "lgcn_score_tc_launches": (ctypes.c_int, [c_i64, c_i64]),
    "lgcn_score_tc_prepare": (ctypes.c_int, [c_vp, c_i64, c_i32, c_vp, ctypes.c_size_t, c_vp]),
    "lgcn_score_tc_topk": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i32, c_vp, c_vp, c_i32,
                                          c_vp, c_vp, c_vp, c_vp, ctypes.c_size_t, c_vp]),
    "lgcn_eval_metrics": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i32, c_vp, c_vp]),
}

_lib = None


class LgcnError(RuntimeError):
    pass


def exported_symbols():
    """Names include/lgcn.h declares (used by the CPU-side ABI test)."""
    return sorted(_SIGNATURES)


def load():
    """Load the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LgcnError(
                f"{LIB_PATH} is missing: run `python -m gcn_recommendation_b200.build` "
                "(nvcc, sm_100a). There is no CPU fallback for the LightGCN hot path.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if lib.lgcn_abi_version() != ABI_VERSION:
            raise LgcnError("liblgcn_b200.so ABI version mismatch; rebuild the library")
        if lib.lgcn_sizeof_spmm_args() != ctypes.sizeof(SpmmArgs):
            raise LgcnError("lgcn_spmm_args layout mismatch between lgcn.h and _lib.py")
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        msg = load().lgcn_error_string(int(rc)).decode()
        raise LgcnError(f"lgcn call failed ({rc}): {msg}")


_DT = {"f32": torch.float32, "i32": torch.int32, "i64": torch.int64, "f64": torch.float64,
       "u8": torch.uint8}


def ptr(t, dt="f32", allow_none=False):
    """Device pointer of a contiguous CUDA tensor (borrowed; the tensor must outlive the call)."""
    if t is None:
        if allow_none:
            return None
        raise LgcnError("required tensor is None")
    if not t.is_cuda:
        raise LgcnError("the LightGCN kernels need CUDA tensors (no CPU fallback)")
    if t.dtype != _DT[dt]:
        raise LgcnError(f"expected dtype {_DT[dt]}, got {t.dtype}")
    if not t.is_contiguous():
        raise LgcnError("tensor must be contiguous")
    return t.data_ptr()


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream
