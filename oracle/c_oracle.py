"""ctypes loader for ``oracle/fir_oracle.c`` (and, when it was built in this
container, ``oracle/_ref/libfir1d_ref.so`` = the reference's own C++ sources).

TEST INFRASTRUCTURE ONLY — see ``oracle/fir_oracle.c``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "_build" / "libfiroracle.so"
REF_LIB_PATH = HERE / "_ref" / "libfir1d_ref.so"
REFERENCE_ROOT = Path(os.environ.get("FIR_REFERENCE_ROOT", "/root/reference"))

_DT_IN = {np.dtype(np.uint8): 0, np.dtype(np.int16): 1, np.dtype(np.int32): 2,
          np.dtype(np.float32): 3, np.dtype(np.float64): 4}
_DT_OUT = {np.dtype(np.uint8): 0, np.dtype(np.int16): 1}


def build(ref: bool = True) -> None:
    """Compile the C restatement; also the reference C++ when its sources are
    present (this container only — the GPU box uses the prebuilt file)."""
    subprocess.run(["make", "-s", "-C", str(HERE)], check=True)
    if ref and (REFERENCE_ROOT / "fir_1d" / "model" / "cpp" / "fir_1d_ref.cpp").exists():
        subprocess.run(["make", "-s", "-C", str(HERE), "ref", f"REFERENCE={REFERENCE_ROOT}"], check=True)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            build(ref=False)
        _lib = C.CDLL(str(LIB_PATH))
        i64, vp, ci = C.c_int64, C.c_void_p, C.c_int
        _lib.firo_fir1d_fixed.argtypes = [vp, vp, ci, ci, i64, i64, i64, i64, vp, ci, ci, ci, ci, ci]
        _lib.firo_fir1d_ideal.argtypes = [vp, vp, ci, i64, i64, i64, i64, vp, ci, ci, ci]
        _lib.firo_fir2d_fixed.argtypes = [vp, vp, ci, ci, i64, i64, i64, i64, i64, i64, i64, vp, ci, ci, ci, ci, ci]
        _lib.firo_fir2d_ideal.argtypes = [vp, vp, ci, i64, i64, i64, i64, i64, i64, i64, vp, ci, ci, ci]
        for f in (_lib.firo_fir1d_fixed, _lib.firo_fir1d_ideal, _lib.firo_fir2d_fixed,
                  _lib.firo_fir2d_ideal, _lib.firo_max_threads):
            f.restype = ci
    return _lib


def max_threads() -> int:
    return int(lib().firo_max_threads())


def _as2d(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x)
    return x.reshape(1, -1) if x.ndim == 1 else x.reshape(-1, x.shape[-1])


def fir1d_fixed(x, hq, frac_bits, acc_bits, *, dilation=1, out_dtype=np.uint8, threads=1):
    """x: [..., cols] of u8/i16/i32/f32/f64; hq: quantised integer taps."""
    x = np.asarray(x)
    x2 = _as2d(x)
    hq = np.ascontiguousarray(hq, dtype=np.int32)
    y = np.empty(x2.shape, dtype=out_dtype)
    rc = lib().firo_fir1d_fixed(x2.ctypes.data, y.ctypes.data, _DT_IN[x2.dtype], _DT_OUT[y.dtype],
                                x2.shape[0], x2.shape[1], x2.shape[1], x2.shape[1],
                                hq.ctypes.data, hq.size, dilation, frac_bits, acc_bits, threads)
    if rc != 0:
        raise ValueError(f"firo_fir1d_fixed rc={rc}")
    return y.reshape(x.shape)


def fir1d_ideal(x, h, *, dilation=1, threads=1):
    x = np.asarray(x)
    x2 = _as2d(x)
    h = np.ascontiguousarray(h, dtype=np.float64)
    y = np.empty(x2.shape, dtype=np.float64)
    rc = lib().firo_fir1d_ideal(x2.ctypes.data, y.ctypes.data, _DT_IN[x2.dtype],
                                x2.shape[0], x2.shape[1], x2.shape[1], x2.shape[1],
                                h.ctypes.data, h.size, dilation, threads)
    if rc != 0:
        raise ValueError(f"firo_fir1d_ideal rc={rc}")
    return y.reshape(x.shape)


def _as3d(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x)
    return x.reshape((-1,) + x.shape[-2:])


def fir2d_fixed(x, hq2d, frac_bits, acc_bits, *, out_dtype=np.uint8, threads=1):
    x = np.asarray(x)
    x3 = _as3d(x)
    hq = np.ascontiguousarray(hq2d, dtype=np.int32)
    b, hgt, wid = x3.shape
    y = np.empty(x3.shape, dtype=out_dtype)
    rc = lib().firo_fir2d_fixed(x3.ctypes.data, y.ctypes.data, _DT_IN[x3.dtype], _DT_OUT[y.dtype],
                                b, hgt, wid, wid, wid, hgt * wid, hgt * wid,
                                hq.ctypes.data, hq.shape[0], hq.shape[1], frac_bits, acc_bits, threads)
    if rc != 0:
        raise ValueError(f"firo_fir2d_fixed rc={rc}")
    return y.reshape(x.shape)


def fir2d_ideal(x, h2d, *, threads=1):
    x = np.asarray(x)
    x3 = _as3d(x)
    h = np.ascontiguousarray(h2d, dtype=np.float64)
    b, hgt, wid = x3.shape
    y = np.empty(x3.shape, dtype=np.float64)
    rc = lib().firo_fir2d_ideal(x3.ctypes.data, y.ctypes.data, _DT_IN[x3.dtype],
                                b, hgt, wid, wid, wid, hgt * wid, hgt * wid,
                                h.ctypes.data, h.shape[0], h.shape[1], threads)
    if rc != 0:
        raise ValueError(f"firo_fir2d_ideal rc={rc}")
    return y.reshape(x.shape)


# ---- the reference's own C++ (only where oracle/_ref was built) -------------
_ref = None


def ref_lib():
    """None when oracle/_ref/libfir1d_ref.so is absent."""
    global _ref
    if _ref is None and REF_LIB_PATH.exists():
        _ref = C.CDLL(str(REF_LIB_PATH))
        _ref.ref_fir1d_idel_full.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_void_p]
        _ref.ref_fir1d_idel_full.restype = C.c_long
        _ref.ref_fir1d_stream.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_void_p]
        _ref.ref_complex_real_tap_mac.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p]
    return _ref


def ref_fir1d_full(x, h) -> np.ndarray:
    """Fir1D::fir_1d_idel of the reference (full convolution, N+L-1)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    h = np.ascontiguousarray(h, dtype=np.float64)
    y = np.empty(x.size + h.size - 1, dtype=np.float64)
    n = ref_lib().ref_fir1d_idel_full(x.ctypes.data, x.size, h.ctypes.data, h.size, y.ctypes.data)
    return y[:n]


def ref_fir1d_stream(x, h) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float64)
    h = np.ascontiguousarray(h, dtype=np.float64)
    y = np.empty(x.size, dtype=np.float64)
    ref_lib().ref_fir1d_stream(x.ctypes.data, x.size, h.ctypes.data, h.size, y.ctypes.data)
    return y


def ref_complex_real_tap_mac(z_interleaved, h):
    z = np.ascontiguousarray(z_interleaved, dtype=np.float64)
    h = np.ascontiguousarray(h, dtype=np.float64)
    re, im = C.c_double(), C.c_double()
    ref_lib().ref_complex_real_tap_mac(z.ctypes.data, h.ctypes.data, h.size, C.byref(re), C.byref(im))
    return re.value, im.value
