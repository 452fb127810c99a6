"""ctypes binding of libtopo_b200.so (the C ABI declared in include/topo_b200.h).

There is no CPU fallback: if the library is missing the import fails loudly, and every compute
entry point refuses non-CUDA tensors.  PyTorch only supplies device memory and the stream.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtopo_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED = 0, 1, 2, 3


class TopoError(RuntimeError):
    pass


class ComplexView(C.Structure):
    """topo_complex_view"""
    _fields_ = [("probs", C.c_void_p), ("pos", C.c_void_p), ("act_idx", C.c_void_p),
                ("counts", C.c_void_p), ("row_off", C.c_void_p), ("batch", C.c_int64)]


class CombineParams(C.Structure):
    """topo_combine_params"""
    _fields_ = [("channels", C.c_int), ("n_msgs", C.c_int),
                ("agg", C.c_void_p * 3), ("w", C.c_void_p * 3), ("scale", C.c_void_p * 3),
                ("x", C.c_void_p), ("att_w1", C.c_void_p), ("att_b1", C.c_void_p),
                ("att_w2", C.c_void_p), ("att_b2", C.c_void_p),
                ("ln_gamma", C.c_void_p), ("ln_beta", C.c_void_p),
                ("ln_eps", C.c_float), ("apply_ln", C.c_int),
                ("saved_m", C.c_void_p * 3), ("saved_pre", C.c_void_p * 3), ("saved_score", C.c_void_p), ("saved_layout", C.c_int), ("weight_images", C.c_void_p), ("max_ctas", C.c_int)]


class ImageJob(C.Structure):
    """topo_image_job"""
    _fields_ = [("w", C.c_void_p), ("scale", C.c_void_p), ("att_w1", C.c_void_p), ("dst", C.c_void_p)]


WEIGHT_IMAGE_BYTES = 24576


class WgradJob(C.Structure):
    """topo_wgrad_job"""
    _fields_ = [("wprod", C.c_void_p), ("w", C.c_void_p), ("scale", C.c_void_p), ("g_w", C.c_void_p), ("g_scale", C.c_void_p),
                ("partials", C.c_void_p), ("n_rows_dev", C.c_void_p), ("rows", C.c_int64), ("partial_offset", C.c_int32),
                ("count", C.c_int32), ("max_ctas", C.c_int32), ("reserved_", C.c_int32)]


class CombineGrads(C.Structure):
    """topo_combine_grads"""
    _fields_ = [("g_agg", C.c_void_p * 3), ("g_x", C.c_void_p), ("g_wprod", C.c_void_p * 3),
                ("g_att_w1", C.c_void_p), ("g_att_b1", C.c_void_p), ("g_att_w2", C.c_void_p),
                ("g_att_b2", C.c_void_p), ("g_ln_gamma", C.c_void_p), ("g_ln_beta", C.c_void_p), ("cta_partials", C.c_void_p)]


CTA_PARTIAL_FLOATS = 16704          # TOPO_CTA_PARTIAL_FLOATS


_P, _I64, _I32, _F = C.c_void_p, C.c_int64, C.c_int, C.c_float
_PP = C.POINTER(C.c_void_p)

# name -> argtypes; every function returns int except the ones listed in _NON_STATUS
SIGNATURES = {
    "topo_version": [],
    "topo_last_error": [],
    "topo_tables_create": [_I32, _PP],
    "topo_tables_create_ex": [_I32, _I32, _PP],
    "topo_tables_destroy": [_P],
    "topo_tables_sizes": [_P, C.POINTER(_I64), C.POINTER(_I64)],
    "topo_tables_simplex_vertices": [_P, _I32, _P],
    "topo_tables_faces": [_P, _I32, _P],
    "topo_tables_cofaces": [_P, _I32, _P],
    "topo_tables_face_matrix": [_P, _I32, _P, _P],
    "topo_hard_concrete_fwd": [_P, _P, _P, C.POINTER(_I64), _I64, _I32, _I32, _P, _P],
    "topo_hard_concrete_bwd": [_P, _P, _P, C.POINTER(_I64), _I64, _I32, _P, _P, _P, _P, _P],
    "topo_hard_concrete_bwd_workspace_floats": [C.POINTER(_I64), _I64],
    "topo_binary_gumbel_fwd": [_P, _P, _F, _I64, _P, _P],
    "topo_binary_gumbel_bwd": [_P, _P, _F, _I64, _P, _P, _P],
    "topo_rectify_fwd": [_P, _P, _F, _I64, _P, _P],
    "topo_rectify_bwd": [_P, _P, _P, _P, _F, _I64, _P, _P, _P],
    "topo_active_sets": [_P, _P, _I64, _P, _P, _P, _P, _P],
    "topo_penalties_fwd": [_P, _P, _I64, _F, _F, _P, _P, _P],
    "topo_penalties_bwd": [_P, _P, _I64, _F, _F, _P, _P, _P, _P],
    "topo_embed_fwd": [_P, C.POINTER(ComplexView), _I32, _I32, _P, _P, _P],
    "topo_embed_bwd": [_P, C.POINTER(ComplexView), _I32, _I32, _P, _P, _P, _P, _P],
    "topo_layernorm_fwd": [_I64, _I32, _P, _P, _P, _F, _P, _P],
    "topo_layernorm_bwd": [_I64, _I32, _P, _P, _F, _P, _P, _P, _P, _P, _P],
    "topo_layernorm_bwd_workspace_floats": [_I64, _I32],
    "topo_operators_count": [_P, _P, _P, _P, _P, _I64, _P, _P],
    "topo_operators_fill": [_P, _P, _P, _P, _P, _I64, _P, _PP, _PP, _PP, _P],
    "topo_operators_bwd": [_P, _P, _P, _P, _P, _I64, _P, _PP, _P, _P],
    "topo_sccn_aggregate_fwd": [_P, C.POINTER(ComplexView), _I32, _PP, _PP, _PP, _PP, _P],
    "topo_sccn_aggregate_bwd": [_P, C.POINTER(ComplexView), _I32, _PP, _PP, _PP, _PP, _PP, _PP, _PP, _P, _P],
    "topo_spmm_csr": [_I64, _P, _P, _P, _P, _I32, _P, _P],
    "topo_sddmm_csr": [_I64, _P, _P, _P, _P, _I32, _P, _P],
    "topo_sccn_prepare_images": [C.POINTER(ImageJob), _I32, _I32, _P],
    "topo_sccn_finish_weight_grads": [C.POINTER(WgradJob), _I32, _I32, _P],
    "topo_sccn_combine_grid": [_I64, _I32],
    "topo_sccn_combine_fwd": [C.POINTER(CombineParams), _I64, _P, _P, _P],
    "topo_sccn_combine_fwd_tc": [C.POINTER(CombineParams), _I64, _P, _P, _P],
    "topo_sccn_combine_fwd_tc2": [C.POINTER(CombineParams), _I64, _P, _P, _P],
    "topo_sccn_combine_bwd": [C.POINTER(CombineParams), _I64, _P, _P, C.POINTER(CombineGrads), _P, _P],
    "topo_sccn_combine_bwd_attention": [C.POINTER(CombineParams), _I64, _P, _P, C.POINTER(CombineGrads), _P, _P],
    "topo_sccn_combine_bwd_conv": [C.POINTER(CombineParams), _I64, _P, C.POINTER(CombineGrads), _P, _P],
    "topo_sccn_combine_bwd_conv_tc": [C.POINTER(CombineParams), _I64, _P, C.POINTER(CombineGrads), _P, _P],
    "topo_sccn_combine_bwd_tc": [C.POINTER(CombineParams), _I64, _P, _P, C.POINTER(CombineGrads), _P],
    "topo_distance_padded_size": [C.POINTER(_I64), _I32],
    "topo_distance_image_bytes": [_I64, C.POINTER(_I64), _I32],
    "topo_distance_logq_words": [_I64, C.POINTER(_I64), _I32],
    "topo_distance_workspace_floats": [_I64, _I64, _I32],
    "topo_distance_prepare": [_P, _I64, _I64, C.POINTER(_I64), _I32, _F, _P, _P, _P, _P],
    "topo_distance_rows": [_P, _P, _P, _I64, C.POINTER(_I64), _I32, _I64, _I64, _I64, _I64, _P, _P, _P],
    "topo_distance_block": [_P, _P, _P, _I64, _I64, _P, _P, _P, _I64, _I64, C.POINTER(_I64), _I32, _P, _P, _I64, _P],
    "topo_cross_attention_fwd": [_P, _P, _P, _P, _I64, _I64, _I32, _P, _P, _P],
    "topo_cross_attention_bwd": [_P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I32, _I64, _P, _P, _P, _P, _P],
}
_NON_STATUS = {"topo_version": C.c_int, "topo_last_error": C.c_char_p, "topo_tables_destroy": None,
               "topo_debug_fwd16_mask": None, "topo_debug_fwd16_stamps": None, "topo_debug_bwd_stamps": None,
               "topo_distance_padded_size": C.c_int64, "topo_distance_image_bytes": C.c_int64,
               "topo_distance_workspace_floats": C.c_int64, "topo_distance_logq_words": C.c_int64,
               "topo_sccn_combine_grid": C.c_int, "topo_hard_concrete_bwd_workspace_floats": C.c_int64,
               "topo_layernorm_bwd_workspace_floats": C.c_int64}

# unit-test / measurement entry points: only in libtopo_b200_debug.so (csrc/build.py build_debug()), never in the product library
DEBUG_LIB_PATH = os.path.join(_HERE, "libtopo_b200_debug.so")
DEBUG_SIGNATURES = {
    "topo_debug_gemm_tf32x3": [_P, _P, _I64, _I32, _P, _P],
    "topo_debug_gemm_bf16x3": [_P, _P, _I64, _I32, _I32, _I32, _I32, _P, _P],
    "topo_debug_fwd16_mask": [_I32],
    "topo_debug_fwd16_stamps": [_P],
    "topo_debug_bwd_stamps": [_P],
}


def load_debug() -> C.CDLL:
    """The debug twin of the library (all product entry points plus DEBUG_SIGNATURES) for tests and scripts."""
    if not os.path.exists(DEBUG_LIB_PATH):
        raise ImportError(f"{DEBUG_LIB_PATH} is missing: build it with `python topo_audio_autoencoder_b200/csrc/build.py`")
    dbg = C.CDLL(DEBUG_LIB_PATH)
    for name, argtypes in {**SIGNATURES, **DEBUG_SIGNATURES}.items():
        fn = getattr(dbg, name)
        fn.argtypes = argtypes
        fn.restype = _NON_STATUS.get(name, C.c_int)
    return dbg


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python topo_audio_autoencoder_b200/csrc/build.py` "
            "(or __graft_entry__.build()).  There is no CPU or PyTorch fallback for this package.")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header / library mismatch
        fn.argtypes = argtypes
        fn.restype = _NON_STATUS.get(name, C.c_int)
    return lib


# kernels launched by one call of each compute entry point (csrc/*.cu), for bench.py's gpu_launches
KERNELS_PER_CALL = {
    "topo_tables_face_matrix": 1, "topo_hard_concrete_fwd": 1, "topo_hard_concrete_bwd": 2,
    "topo_binary_gumbel_fwd": 1, "topo_binary_gumbel_bwd": 1, "topo_rectify_fwd": 4, "topo_rectify_bwd": 4,
    "topo_active_sets": 2, "topo_penalties_fwd": 1, "topo_penalties_bwd": 1, "topo_embed_fwd": 1,
    "topo_embed_bwd": 2, "topo_layernorm_fwd": 1, "topo_layernorm_bwd": 2, "topo_operators_count": 2,
    "topo_operators_fill": 1, "topo_operators_bwd": 1, "topo_sccn_aggregate_fwd": 2, "topo_sccn_aggregate_bwd": 3,
    "topo_spmm_csr": 1, "topo_sddmm_csr": 1, "topo_sccn_prepare_images": 1, "topo_sccn_finish_weight_grads": 1, "topo_sccn_combine_fwd": 1, "topo_sccn_combine_fwd_tc": 1, "topo_sccn_combine_fwd_tc2": 1, "topo_sccn_combine_bwd": 2,
    "topo_sccn_combine_bwd_attention": 1, "topo_sccn_combine_bwd_conv": 1, "topo_sccn_combine_bwd_conv_tc": 1, "topo_sccn_combine_bwd_tc": 1, "topo_distance_prepare": 2,
    "topo_distance_rows": 3, "topo_distance_block": 3, "topo_cross_attention_fwd": 1, "topo_cross_attention_bwd": 2,
}


class _Instrumented:
    """The loaded library with per-entry-point call counting and, on request, CUDA-event timing of
    every compute call on torch's current stream (the stream the kernels are enqueued on)."""

    def __init__(self, cdll):
        self._cdll = cdll
        self.calls = {}
        self.timing = None        # None, or dict name -> list of (start_event, end_event)
        for name in SIGNATURES:
            fn = getattr(cdll, name)
            if name in KERNELS_PER_CALL:
                setattr(self, name, self._wrap(name, fn))
            else:
                setattr(self, name, fn)

    def _wrap(self, name, fn):
        def call(*args):
            self.calls[name] = self.calls.get(name, 0) + 1
            if self.timing is None:
                return fn(*args)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*args)
            e1.record()
            self.timing.setdefault(name, []).append((e0, e1))
            return rc
        return call

    def reset_counts(self):
        self.calls = {}

    def kernel_launches(self) -> int:
        return sum(KERNELS_PER_CALL[k] * v for k, v in self.calls.items())

    def start_timing(self):
        self.timing = {}

    def stop_timing(self):
        """-> dict name -> (calls, total_ms); synchronises."""
        torch.cuda.synchronize()
        out = {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in (self.timing or {}).items()}
        self.timing = None
        return out


lib = _Instrumented(_load())


def check(status: int) -> None:
    if status != OK:
        kind = {ERR_INVALID: "invalid argument", ERR_CUDA: "CUDA error", ERR_UNSUPPORTED: "unsupported"}.get(status, "error")
        raise TopoError(f"libtopo_b200 {kind}: {lib.topo_last_error().decode()}")


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t: Optional[torch.Tensor], dtype=torch.float32) -> Optional[int]:
    """Device pointer of a contiguous CUDA tensor (None passes through as NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise TopoError("libtopo_b200 kernels need CUDA tensors; there is no CPU fallback")
    if t.dtype != dtype:
        raise TopoError(f"expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise TopoError("expected a contiguous tensor")
    return t.data_ptr()


def ptr_array(tensors: Sequence[Optional[torch.Tensor]], n: int, dtype=torch.float32):
    arr = (C.c_void_p * n)()
    for i in range(n):
        t = tensors[i] if i < len(tensors) else None
        arr[i] = ptr(t, dtype)
    return arr


def i64_array(values: Sequence[int]):
    return (C.c_int64 * len(values))(*[int(v) for v in values])
