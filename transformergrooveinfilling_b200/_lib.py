"""ctypes binding of libgroove_b200.so (C ABI declared in include/groove_b200.h).

The library is the ONLY compute path: there is no CPU / eager fallback.  If it cannot be built or
loaded, importing the compute entry points raises."""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import build as _build

PREC_FP32, PREC_BF16, PREC_FP32_TC = 0, 1, 2
PATH_FP32_SIMT, PATH_FUSED_D32, PATH_FUSED_D256, PATH_GEMM_TC, PATH_GEMM_TC_SPLIT = 0, 1, 2, 3, 4
T_STEPS = 32


class GtConfig(C.Structure):
    _fields_ = [
        ("d_model", C.c_int32), ("nhead", C.c_int32), ("dim_ff", C.c_int32), ("n_enc", C.c_int32),
        ("n_dec", C.c_int32), ("e_src", C.c_int32), ("e_tgt", C.c_int32), ("precision", C.c_int32),
        ("dropout", C.c_float), ("reserved", C.c_int32),
    ]


_p = C.c_void_p
_i64 = C.c_int64
_u64 = C.c_uint64
_f = C.c_float
_cfgp = C.POINTER(GtConfig)

# name -> (restype, argtypes); must list every symbol include/groove_b200.h declares
SIGNATURES = {
    "gt_version": (C.c_int, []),
    "gt_last_error": (C.c_char_p, []),
    "gt_path_kind": (C.c_int, [_cfgp]),
    "gt_param_count": (_i64, [_cfgp]),
    "gt_param_layout": (C.c_int, [_cfgp, C.POINTER(_i64), C.POINTER(_i64), C.c_int]),
    "gt_workspace_bytes": (_i64, [_cfgp, _i64, C.c_int]),
    "gt_forward": (C.c_int, [_cfgp, _p, _p, _p, _p, _i64, _p, _p, _i64, C.c_int, _u64, _u64, _i64, _p]),
    "gt_backward": (C.c_int, [_cfgp, _p, _p, _p, _p, _i64, _p, _p, _p, _p, _i64, _u64, _u64, _i64, _p]),
    "gt_loss_scratch_floats": (_i64, [_i64]),
    "gt_loss": (C.c_int, [_p, _p, _i64, _f, _p, _p, _f, _p, _p]),
    "gt_loss_voices": (C.c_int, [_p, _p, _i64, C.c_int, _f, _p, _p, _f, _p, _p]),
    "gt_eval_scratch_floats": (_i64, [_i64, C.c_int]),
    "gt_eval_metrics": (C.c_int, [_p, _p, _i64, C.c_int, _p, _p, _p]),
    "gt_train_step": (C.c_int, [_cfgp, _p, _p, _p, _p, _i64, _f, _p, _p, _p, _p, _i64, _u64, _u64, _i64, _p]),
    "gt_predict": (C.c_int, [_cfgp, _p, _p, _p, _i64, _f, _p, _p, _i64, _p]),
    "gt_predict_variant": (C.c_int, [_cfgp, _p, _p, _p, _i64, _f, _p, _p, _i64, C.c_int, _p]),
    "gt_sgd_step": (C.c_int, [_p, _p, _i64, _f, _f, _p]),
    "gt_adam_step": (C.c_int, [_p, _p, _p, _p, _i64, _f, _f, _f, _f, _i64, _f, _p]),
    "gt_peer_alloc": (C.c_int, [_i64, C.POINTER(C.c_void_p), C.c_char_p]),
    "gt_peer_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "gt_peer_close": (C.c_int, [_p]),
    "gt_peer_free": (C.c_int, [_p]),
    "gt_peer_publish": (C.c_int, [_p, _i64, _p, _i64, _p]),
    "gt_sgd_step_peers": (C.c_int, [_p, C.POINTER(C.c_void_p), C.c_int, _i64, _p, _i64, _f, _f, _p]),
    "gt_adam_step_peers": (C.c_int, [_p, C.POINTER(C.c_void_p), C.c_int, _i64, _p, _p, _p, _i64, _f, _f, _f, _f, _i64, _f, _p]),
    "gt_graph_train_create": (C.c_int, [_cfgp, _p, _p, _p, _p, _i64, _f, _p, _p, _p, _p, _i64, C.c_int, _f, _p, _p, _u64, _p,
                                        _p, _p, _p, _p, _i64, _p, C.POINTER(C.c_void_p)]),
    "gt_graph_launch": (C.c_int, [_p, C.c_int, _p]),
    "gt_graph_destroy": (C.c_int, [_p]),
    "gt_train_steps": (C.c_int, [_cfgp, _p, _p, _p, _p, _p, _i64, _i64, C.c_int, _f, _p, _p, _p, _p, _p, _p, _i64, C.c_int, _f, _p, _p,
                                 _i64, _u64, _u64, _p]),
    "gt_grad_buckets": (C.c_int, [_cfgp, C.POINTER(_i64), C.POINTER(_i64), C.c_int]),
    "gt_grad_events_enable": (C.c_int, [C.c_int]),
    "gt_grad_bucket_wait": (C.c_int, [C.c_int, _p]),
    "gt_launch_count": (_i64, [C.c_int]),
    "gt_profile_enable": (C.c_int, [C.c_int, C.c_int]),
    "gt_profile_collect": (C.c_int, [C.POINTER(C.c_double), C.POINTER(_i64)]),
    "gt_profile_collect_class": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(_i64)]),
    "gt_debug_dropout_mask": (C.c_int, [_u64, _u64, C.c_int32, _f, _i64, _i64, _p, _p]),
    "gt_debug_gemm": (C.c_int, [C.c_int, _p, _i64, _i64, _p, _i64, _i64, _p, _i64, _i64, _i64, _i64, C.c_int, _p, _p, _i64, _p, _i64,
                                _f, _f, _u64, _u64, C.c_int32, _i64, _i64, _p]),
    "gt_debug_gemm_scratch": (C.c_int, [_p, _i64]),
    "gt_debug_umma_rate": (C.c_int, [C.c_int, C.c_int, C.c_int, _p, _p]),
    "gt_debug_tc_gemm": (C.c_int, [_p, _p, _p, C.c_int, C.c_int, C.c_int, C.c_int, _p]),
}

_lock = threading.Lock()
_lib = None


def lib_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True):
    """Load (building first if needed) the CUDA library.  Raises RuntimeError on any failure."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = _build.LIB
        if build_if_missing:
            path = _build.build()
        if not os.path.exists(path):
            raise RuntimeError(f"groove_b200: CUDA library {path} is missing and no CPU fallback exists")
        lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        if lib.gt_version() != 1:
            raise RuntimeError("groove_b200: ABI version mismatch")
        _lib = lib
        return lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().gt_last_error()
        raise RuntimeError(f"groove_b200 {what} failed: {msg.decode() if msg else 'unknown error'}")


def ptr(t) -> int:
    """Device pointer of a torch tensor (None -> NULL)."""
    return 0 if t is None else t.data_ptr()


def stream_ptr(device) -> int:
    import torch
    return torch.cuda.current_stream(device).cuda_stream
