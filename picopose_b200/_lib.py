"""ctypes binding of libpicopose_b200.so (the C ABI declared in include/picopose_b200.h).

There is deliberately no fallback: if the shared library is missing or a call
fails, a RuntimeError is raised -- the product path never routes through
PyTorch ops or the CPU oracle.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
# PICOPOSE_B200_LIB names another build of the same library (A/B runs of two builds in one process tree)
LIB_PATH = os.environ.get("PICOPOSE_B200_LIB") or os.path.join(HERE, "lib", "libpicopose_b200.so")

MODE_BF16, MODE_FP32, MODE_BF16X3 = 0, 1, 2
MODES = {"bf16": MODE_BF16, "fp32": MODE_FP32, "bf16x3": MODE_BF16X3}
MATCH_FAST_KEYS = 0x100   # PP_MATCH_FAST_KEYS

# name -> (restype, argtypes); must list every symbol include/picopose_b200.h declares
_vp, _i, _i64, _sz, _f = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_float
SIGNATURES = {
    "pp_version": (_i, []),
    "pp_last_error": (C.c_char_p, []),
    "pp_check_device_faults": (_i, []),
    "pp_launch_count": (C.c_longlong, []),
    "pp_profile_gemm_events": (None, [_vp, _vp]),
    "pp_corr_lookup": (_i, [C.POINTER(_vp), C.POINTER(_i), C.POINTER(_i), _i, _vp, _i, _i, _i, _i, _vp, _vp]),
    "pp_corr_lookup_tiled": (_i, [C.POINTER(_vp), C.POINTER(_i), C.POINTER(_i), _i, _vp, _i, _i, _i, _i, _vp, _vp]),
    "pp_volume_retile": (_i, [_vp, _vp, _i64, _i, _i, _i, _vp]),
    "pp_bilinear_sample": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "pp_grid_sample": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "pp_match_kp": (_i, [_i, _i]),
    "pp_match_prepare": (_i, [_vp, _i64, _i, _i, _i, _i, _vp, _vp, _vp]),
    "pp_match_scores_workspace": (_sz, [_i, _i, _i]),
    "pp_match_query_meta_bytes": (_sz, [_i, _i]),
    "pp_match_prepare_query": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "pp_match_scores": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz,
                             _i, _vp]),
    "pp_topk": (_i, [_vp, _i, _i, _i, _i64, _vp, _vp, _vp]),
    "pp_match_templates_workspace": (_sz, [_i, _i, _i, _i, _i, _i]),
    "pp_match_templates": (_i, [_vp, _vp, _vp, _vp, _i64, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _sz,
                                _i, _vp]),
    "pp_match_templates_dense": (_i, [_vp, _i64, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp,
                                      _vp, _sz, _i, _vp]),
    "pp_topk_pairs": (_i, [_vp, _i, _i, _i, _i64, _vp, _vp]),
    "pp_topk_merge": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "pp_xchg_bytes": (_sz, [_i, _i, _i]),
    "pp_xchg_create": (_i, [_sz, C.POINTER(_vp), _vp]),
    "pp_xchg_open": (_i, [_vp, C.POINTER(_vp)]),
    "pp_xchg_close": (_i, [_vp]),
    "pp_xchg_destroy": (_i, [_vp]),
    "pp_xchg_push": (_i, [_vp, _sz, _vp, _i, _sz, _vp]),
    "pp_xchg_push_signal": (_i, [_vp, _sz, _sz, _vp, _sz, _sz, _vp, _vp, _sz, _i, _i, C.c_uint32, _vp]),
    "pp_xchg_signal": (_i, [_vp, _sz, _i, _i, C.c_uint32, _vp]),
    "pp_xchg_wait": (_i, [_vp, _sz, _i, C.c_uint32, _vp]),
    "pp_topk_exchange": (_i, [_vp, _i, _i, _i, _i64, _vp, _i, _i, _i, _i, C.c_uint32, _vp, _vp, _vp]),
    "pp_match_similarity_workspace": (_sz, [_i, _i]),
    "pp_match_similarity": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _sz, _i, _vp]),
    "pp_match_similarity_dense_workspace": (_sz, [_i, _i, _i, _i, _i]),
    "pp_match_similarity_dense": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _sz, _i, _vp]),
    "pp_correlation_pyramid": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _i, C.POINTER(_vp), _i, _vp]),
    "pp_correlation_pyramid_tiled": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _i, C.POINTER(_vp), _i, _vp]),
    "pp_windowed_correlation_prepare": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "pp_windowed_correlation_prepare_all": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, C.POINTER(_vp), _vp]),
    "pp_windowed_correlation": (_i, [_vp, C.POINTER(_vp), _i, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "pp_windowed_correlation_conv1x1": (_i, [_vp, C.POINTER(_vp), _i, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp, _vp]),
    "pp_select_templates": (_i, [C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i64), _i, _i, _i, _vp, _i, _i, _vp]),
    "pp_init_correspondences": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "pp_stage3_correspondences": (_i, [_vp, _vp, _i, _i, _i, _f, _vp, _vp, _vp]),
}

_lock = threading.Lock()
_lib = None


def load() -> C.CDLL:
    """Loads the shared library once; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            # not a fallback: the same CUDA sources are compiled in-tree when the toolchain is at hand
            try:
                from . import build as _build
                _build.build()
            except Exception as exc:  # noqa: BLE001
                raise RuntimeError(
                    f"{LIB_PATH} is missing and could not be built ({exc}); run `python -m picopose_b200.build` "
                    "(nvcc, sm_100a). picopose_b200 has no PyTorch/CPU fallback path.") from exc
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export the symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    return load().pp_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise RuntimeError(f"picopose_b200 {what} failed ({rc}): {last_error()}")


def check_device_faults() -> None:
    check(load().pp_check_device_faults(), "device fault check")


# ---- torch helpers (torch is only the allocator / stream provider) -----------------------------

def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "picopose_b200 runs on a CUDA (sm_100) device only: got a tensor on "
                f"'{t.device}'; there is no CPU fallback")


def require_inference(what: str, *tensors) -> None:
    """The kernels have no backward: fail loudly instead of silently cutting the autograd graph (the reference calls
    these functions under autograd in forward_train, model/picopose.py:114-137 -- training keeps the reference modules)."""
    import torch
    if torch.is_grad_enabled():
        for t in tensors:
            if isinstance(t, torch.Tensor) and t.requires_grad:
                raise RuntimeError(
                    f"picopose_b200.{what} is inference-only (no autograd): an input requires grad and grad mode is on. "
                    "Call it under torch.no_grad() (as run_test.py:165 does) or use the reference module for training.")


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def stream_of(t) -> int:
    import torch
    return torch.cuda.current_stream(t.device).cuda_stream
