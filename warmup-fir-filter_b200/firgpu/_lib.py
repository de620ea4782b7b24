"""ctypes binding of ``libfirgpu.so`` (the C ABI in ``include/firgpu.h``).

The library is built in-tree by ``__graft_entry__.build()`` (``make -C
warmup-fir-filter_b200/csrc``).  There is no CPU implementation behind this
module: if the shared object is missing, importing it raises, and every compute
call needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("FIRGPU_LIB") or HERE / "libfirgpu.so")      # FIRGPU_LIB: another build of the same library (kernel experiments)

# enum firgpu_dtype
U8, I16, I32, F32, F64 = 0, 1, 2, 3, 4
# enum firgpu_path
PATH_NONE, PATH_GENERIC, PATH_STAGED, PATH_EMPTY = 0, 1, 2, 3

_i64, _vp, _ci = C.c_int64, C.c_void_p, C.c_int

# name -> (restype, argtypes); one entry per symbol declared in include/firgpu.h
SIGNATURES = {
    "firgpu_version": (_ci, []),
    "firgpu_last_error": (C.c_char_p, []),
    "firgpu_last_path": (_ci, []),
    "firgpu_launch_count": (_i64, []),
    "firgpu_device_count": (_ci, [C.POINTER(_ci)]),
    "firgpu_fir1d_fixed": (_ci, [_vp, _vp, _ci, _ci, _i64, _i64, _i64, _i64, _vp, _ci, _ci, _ci, _ci, _i64, _i64, _vp]),
    "firgpu_fir1d_fixed_multi": (_ci, [_vp, _vp, _ci, _ci, _i64, _i64, _i64, _i64, _i64, _vp, _ci, _ci, _ci, _ci, _ci, _vp]),
    "firgpu_fir1d_ideal": (_ci, [_vp, _vp, _ci, _ci, _i64, _i64, _i64, _i64, _vp, _ci, _ci, _i64, _i64, _vp]),
    "firgpu_fir1d_fixed_ex": (_ci, [_vp, _vp, _ci, _ci, _i64, _i64, _i64, _i64, _vp, _ci, _ci, _ci, _ci, _i64, _i64, _ci, _ci, _ci, _ci, _ci, _vp]),
    "firgpu_fir1d_ideal_ex": (_ci, [_vp, _vp, _ci, _ci, _i64, _i64, _i64, _i64, _vp, _ci, _ci, _i64, _i64, _ci, _ci, _ci, _ci, _vp]),
    "firgpu_fir1d_fixed_ex_host": (_ci, [_vp, _vp, _ci, _ci, _i64, _i64, _i64, _i64, _vp, _ci, _ci, _ci, _ci, _ci, _ci, _ci, _ci, _ci, _ci]),
    "firgpu_fir1d_ideal_ex_host": (_ci, [_vp, _vp, _ci, _ci, _i64, _i64, _i64, _i64, _vp, _ci, _ci, _ci, _ci, _ci, _ci, _ci]),
    "firgpu_fir1d_compare": (_ci, [_vp, _ci, _i64, _i64, _i64, _vp, _vp, _ci, _ci, _ci, _vp, _vp]),
    "firgpu_fir1d_compare_host": (_ci, [_vp, _ci, _i64, _i64, _i64, _vp, _vp, _ci, _ci, _ci, _vp, _ci]),
    "firgpu_fir2d_fixed": (_ci, [_vp, _vp, _ci, _ci, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _ci, _ci, _ci, _ci, _i64, _i64, _vp]),
    "firgpu_fir2d_ideal": (_ci, [_vp, _vp, _ci, _ci, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _ci, _ci, _i64, _i64, _vp]),
    "firgpu_fir1d_fixed_host": (_ci, [_vp, _vp, _ci, _ci, _i64, _i64, _i64, _i64, _vp, _ci, _ci, _ci, _ci, _ci]),
    "firgpu_fir1d_fixed_multi_host": (_ci, [_vp, _vp, _ci, _ci, _i64, _i64, _i64, _i64, _i64, _vp, _ci, _ci, _ci, _ci, _ci, _ci]),
    "firgpu_fir1d_ideal_host": (_ci, [_vp, _vp, _ci, _ci, _i64, _i64, _i64, _i64, _vp, _ci, _ci, _ci]),
    "firgpu_fir2d_fixed_host": (_ci, [_vp, _vp, _ci, _ci, _i64, _i64, _i64, _vp, _ci, _ci, _ci, _ci, _ci]),
    "firgpu_fir2d_ideal_host": (_ci, [_vp, _vp, _ci, _ci, _i64, _i64, _i64, _vp, _ci, _ci, _ci]),
    "firgpu_stream_create_fixed": (_ci, [C.POINTER(_vp), _ci, _ci, _vp, _ci, _ci, _ci, _ci, _ci, _ci, _ci, _ci, _ci]),
    "firgpu_stream_create_ideal": (_ci, [C.POINTER(_vp), _ci, _ci, _vp, _ci, _ci, _ci, _ci, _ci, _ci]),
    "firgpu_stream_process": (_ci, [_vp, _vp, _vp, _i64, _vp]),
    "firgpu_stream_process_host": (_ci, [_vp, _vp, _vp, _i64]),
    "firgpu_stream_reset": (_ci, [_vp, _vp]),
    "firgpu_stream_history": (_ci, [_vp, C.POINTER(_i64)]),
    "firgpu_stream_destroy": (_ci, [_vp]),
    "firgpu_record_begin": (_ci, [_vp]),
    "firgpu_record_end": (_ci, [_vp, C.POINTER(_vp)]),
    "firgpu_record_abort": (_ci, [_vp]),
    "firgpu_graph_launch": (_ci, [_vp, _vp]),
    "firgpu_graph_nodes": (_ci, [_vp, C.POINTER(_i64)]),
    "firgpu_graph_destroy": (_ci, [_vp]),
    "firgpu_fir1d_fixed_complex": (_ci, [_vp, _vp, _ci, _ci, _i64, _i64, _i64, _i64, _vp, _vp, _ci, _ci, _ci, _i64, _i64, _ci, _ci, _ci, _ci, _vp]),
    "firgpu_fir1d_fixed_complex_host": (_ci, [_vp, _vp, _ci, _ci, _i64, _i64, _i64, _i64, _vp, _vp, _ci, _ci, _ci, _ci, _ci, _ci, _ci, _ci]),
    "firgpu_fir1d_fixed_plan": (_ci, [_vp, _ci, _ci, _ci, C.POINTER(_ci), C.POINTER(_ci), C.POINTER(_ci)]),
    "firgpu_get_device": (_ci, [C.POINTER(_ci)]),
    "firgpu_set_device": (_ci, [_ci]),
    "firgpu_fir1d_fixed_host_multi": (_ci, [_vp, _vp, _ci, _ci, _i64, _i64, _i64, _i64, _vp, _ci, _ci, _ci, _ci, C.POINTER(_ci), _ci]),
    "firgpu_pinned_alloc": (_ci, [C.POINTER(_vp), C.c_size_t]),
    "firgpu_pinned_free": (_ci, [_vp]),
    "firgpu_release_cache": (_ci, []),
}


class FirGpuError(RuntimeError):
    """A libfirgpu call returned non-zero (negative: rejected argument,
    positive: cudaError_t)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"libfirgpu error {code}: {message}")
        self.code = code


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
                "(nvcc, sm_100a).  firgpu has no CPU fallback."
            )
        handle = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise FirGpuError(rc, lib().firgpu_last_error().decode("utf-8", "replace"))


def last_path() -> int:
    return int(lib().firgpu_last_path())


def launch_count() -> int:
    return int(lib().firgpu_launch_count())


def device_count() -> int:
    n = _ci(0)
    check(lib().firgpu_device_count(C.byref(n)))
    return n.value
