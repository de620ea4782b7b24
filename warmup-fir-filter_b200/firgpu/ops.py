"""Array-level entry points of the B200 FIR path.

Every function validates and quantises on the host exactly like the reference
(``firgpu.quant``), then makes ONE call into ``libfirgpu.so``:

* NumPy arrays / Python sequences / CPU torch tensors -> the ``*_host`` C entry
  points (chunked H2D -> kernel -> D2H pipeline inside the library);
* CUDA torch tensors (or anything else exposing a device pointer through
  ``DeviceArray``) -> the stream-ordered device entry points, result stays on
  the device.

Shapes: 1-D functions filter along the last axis, every leading index is an
independent, separately zero-padded row (reference:
fir_1d/sim/vector/gen_fixed_output.py:44-59).  2-D functions filter the last
two axes.  There is no CPU arithmetic path in this module.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib, quant

_NP_CODE = {np.dtype(np.uint8): _lib.U8, np.dtype(np.int16): _lib.I16, np.dtype(np.int32): _lib.I32,
            np.dtype(np.float32): _lib.F32, np.dtype(np.float64): _lib.F64}
_OUT_FIXED = {np.dtype(np.uint8): _lib.U8, np.dtype(np.int16): _lib.I16}
_OUT_IDEAL = {np.dtype(np.float64): _lib.F64, np.dtype(np.float32): _lib.F32}


@dataclass
class DeviceArray:
    """A raw device buffer for callers that do not use torch: ``ptr`` is a
    CUDA device pointer to ``shape`` elements of ``dtype`` (C order)."""
    ptr: int
    shape: tuple
    dtype: np.dtype
    device: int = 0
    stream: int = 0


_TAP_CACHE: dict = {}


def _tap_key(h, frac_bits: int, acc_bits: int, coeff_bits: int):
    """Cache key of an already validated + quantised tap set, or None when the
    taps are not plain Python numbers (arrays, exotic scalars: never cached)."""
    try:
        key = (tuple(h), frac_bits, acc_bits, coeff_bits)
        hash(key)
    except TypeError:
        return None
    return key if all(type(v) in (int, float) for v in key[0]) else None


def _remember_taps(key, hq: np.ndarray) -> None:
    if key is not None:
        if len(_TAP_CACHE) > 256:
            _TAP_CACHE.clear()
        hq.setflags(write=False)
        _TAP_CACHE[key] = hq


class _on_device:
    """`with torch.cuda.device(d)` only when d is not already current (the
    context manager costs more than a small launch)."""

    def __init__(self, torch, device):
        self.cm = None if torch.cuda.current_device() == device.index else torch.cuda.device(device)

    def __enter__(self):
        if self.cm is not None:
            self.cm.__enter__()

    def __exit__(self, *exc):
        if self.cm is not None:
            self.cm.__exit__(*exc)


def _is_torch(x) -> bool:
    return type(x).__module__.split(".")[0] == "torch"


def _raw_stream(torch, device) -> int:
    """cudaStream_t of torch's current stream on ``device`` as an integer.  torch.cuda.current_stream() builds a
    Stream object per call (about 7 microseconds); the raw accessor is what launch-latency-sensitive callers use."""
    try:
        return torch._C._cuda_getCurrentRawStream(device.index if device.index is not None else torch.cuda.current_device())
    except AttributeError:                      # an older / newer torch without the private accessor
        return torch.cuda.current_stream(device).cuda_stream


def _np_container(x, lo: int = 0, hi: int = 255, raw: bool = False) -> np.ndarray:
    """Any host input -> C-contiguous array in one of the five containers the
    kernels read.  Integer types the kernels do not read are clamped here to
    the sample range [lo, hi] (rounding is the identity on integers,
    fir_1d_ref.py:35-41); with ``raw`` (samples used as they are) they become
    float64 instead."""
    a = np.asarray(x)
    if a.dtype == object:
        a = np.asarray(x, dtype=np.float64)
    if a.dtype in _NP_CODE:
        pass
    elif raw:
        a = a.astype(np.float64)
    elif a.dtype == np.bool_:
        a = a.astype(np.uint8)
    elif a.dtype == np.int8:
        a = a.astype(np.int16)
    elif a.dtype.kind in "iu":
        a = np.clip(a, lo, hi).astype(np.uint8 if (lo >= 0 and hi <= 255) else np.int16)
    elif a.dtype.kind == "f":
        if a.dtype == np.float16 and not raw:
            # the reference adds 0.5 in the array's own precision (fir_1d_ref.py:37 on np.float16 scalars); the only
            # value where that differs from exact arithmetic is the largest float16 below 0.5, which ties up to 1.0
            a = np.where(a == np.float16(0.49975586), np.float32(0.5), a.astype(np.float32))
        else:
            a = a.astype(np.float32 if a.dtype.itemsize < 4 else np.float64)
    else:
        raise TypeError(f"unsupported sample dtype {a.dtype}")
    return np.ascontiguousarray(a)


def _torch_container(t, lo: int = 0, hi: int = 255):
    import torch
    if t.dtype in (torch.uint8, torch.int16, torch.int32, torch.float32, torch.float64):
        pass
    elif t.dtype == torch.bool:
        t = t.to(torch.uint8)
    elif t.dtype == torch.int8:
        t = t.to(torch.int16)
    elif t.dtype in (torch.int64,):
        t = t.clamp(lo, hi).to(torch.uint8 if (lo >= 0 and hi <= 255) else torch.int16)
    elif t.dtype == torch.float16:
        wide = t.to(torch.float32)                       # (the float16 tie below 0.5, see _np_container)
        t = torch.where(t == 0.49975586, torch.full_like(wide, 0.5), wide)
    elif t.dtype == torch.bfloat16:
        t = t.to(torch.float32)
    else:
        raise TypeError(f"unsupported sample dtype {t.dtype}")
    return t.contiguous()


def _torch_code(t) -> int:
    import torch
    return {torch.uint8: _lib.U8, torch.int16: _lib.I16, torch.int32: _lib.I32,
            torch.float32: _lib.F32, torch.float64: _lib.F64}[t.dtype]


def _torch_dtype(np_dtype):
    import torch
    return {np.dtype(np.uint8): torch.uint8, np.dtype(np.int16): torch.int16,
            np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}[np.dtype(np_dtype)]


def _check_finite_torch(t) -> None:
    import torch
    if t.is_floating_point() and t.numel():
        ok = torch.isfinite(t)
        if not bool(ok.all()):
            index = int(torch.argmin(ok.reshape(-1).to(torch.uint8)))
            raise ValueError(f"Invalid x[{index}]={float(t.reshape(-1)[index])}: x must be finite.")


def _cuda_rows(x, allow_copy: bool, lo: int = 0, hi: int = 255):
    """CUDA tensor -> (tensor to read, rows, cols, row pitch in elements).
    1-D tensors and 2-D tensors whose last axis is dense are read in place,
    pitch included (that is how a shard view of a wider buffer arrives);
    anything else is made contiguous, which is only legal without halos."""
    import torch
    native = x.dtype in (torch.uint8, torch.int16, torch.int32, torch.float32, torch.float64)
    if native and x.is_contiguous():                       # the common case first: no stride arithmetic
        shape = x.shape
        if len(shape) == 1:
            return x, 1, int(shape[0]), int(shape[0])
        if len(shape) == 2:
            return x, int(shape[0]), int(shape[1]), int(shape[1])
        rows, cols = _rows_cols(shape)
        return x, rows, cols, cols
    if native and x.dim() == 1 and (x.stride(0) == 1 or x.shape[0] <= 1):
        return x, 1, int(x.shape[0]), int(x.shape[0])
    if native and x.dim() == 2 and (x.stride(1) == 1 or x.shape[1] <= 1) and (x.stride(0) >= x.shape[1] or x.shape[0] <= 1):
        return x, int(x.shape[0]), int(x.shape[1]), int(max(x.stride(0), x.shape[1]))
    if native and x.is_contiguous():
        rows, cols = _rows_cols(x.shape)
        return x, rows, cols, cols
    if not allow_copy:
        raise ValueError("this tensor would need a dtype/layout copy, which is not possible for in-place shard views")
    t = _torch_container(x, lo, hi)
    rows, cols = _rows_cols(t.shape)
    return t, rows, cols, cols


def _cuda_imgs(t, what: str):
    """CUDA tensor [..., H, W] -> (n_img, H, W, row pitch, image stride) in
    elements when it can be addressed in place: last axis dense, at most three
    axes (or contiguous leading axes), non-overlapping rows.  Column / row
    views of a larger image and row bands qualify."""
    n, hgt, wid = _imgs(t.shape)
    if t.is_contiguous():
        return n, hgt, wid, wid, hgt * wid
    ok = t.dim() in (2, 3) and (t.stride(-1) == 1 or wid <= 1) and (t.stride(-2) >= wid or hgt <= 1)
    if ok and t.dim() == 3:
        ok = t.stride(0) >= (hgt - 1) * t.stride(-2) + wid or n <= 1
    if not ok:
        raise ValueError(f"{what} must be contiguous or a dense-last-axis view of at most three axes")
    ld = int(t.stride(-2)) if hgt > 1 else wid
    img_stride = int(t.stride(0)) if (t.dim() == 3 and n > 1) else hgt * ld
    return n, hgt, wid, ld, img_stride


def _cuda_out(out, x, out_dtype):
    """The caller's ``out`` (or a fresh tensor) for a device call.  The kernels write through the raw pointer with
    the element size of ``out_dtype``, so a tensor of another dtype, on the host, or on another GPU must never get
    that far (it would be overrun, or fault)."""
    import torch
    want = _torch_dtype(out_dtype)
    if out is None:
        return torch.empty(x.shape, dtype=want, device=x.device)
    if not (_is_torch(out) and out.is_cuda and out.device == x.device and out.dtype == want):
        raise ValueError(f"out must be a {want} tensor on {x.device} (got "
                         f"{getattr(out, 'dtype', type(out).__name__)} on {getattr(out, 'device', 'host')})")
    return out


class _device_scope:
    """Make ``device`` current for a raw-pointer call (no torch involved)."""

    def __init__(self, device: int):
        self.device, self.prev = int(device), None

    def __enter__(self):
        L = _lib.lib()
        cur = C.c_int(0)
        _lib.check(L.firgpu_get_device(C.byref(cur)))
        if cur.value != self.device:
            _lib.check(L.firgpu_set_device(self.device))
            self.prev = cur.value

    def __exit__(self, *exc):
        if self.prev is not None:
            _lib.lib().firgpu_set_device(self.prev)


def _host_out(out, shape, dtype):
    """The caller's ``out`` (or a fresh array) for a host-buffer call: must be a
    C-contiguous array / CPU tensor of exactly the expected shape and dtype,
    because the library writes through its raw pointer."""
    if out is None:
        return np.empty(shape, dtype=dtype)
    arr = out.numpy() if _is_torch(out) else out
    if not isinstance(arr, np.ndarray) or arr.dtype != np.dtype(dtype) or tuple(arr.shape) != tuple(shape) or not arr.flags["C_CONTIGUOUS"] \
            or not arr.flags["WRITEABLE"]:
        raise ValueError(f"out must be a writable C-contiguous {np.dtype(dtype).name} array of shape {tuple(shape)}")
    return arr


def _rows_cols(shape) -> tuple[int, int]:
    if len(shape) == 0:
        raise ValueError("x must have at least one dimension")
    cols = int(shape[-1])
    rows = 1
    for s in shape[:-1]:
        rows *= int(s)
    return rows, cols


def _imgs(shape) -> tuple[int, int, int]:
    if len(shape) < 2:
        raise ValueError("2-D FIR needs an input with at least two dimensions")
    n = 1
    for s in shape[:-2]:
        n *= int(s)
    return n, int(shape[-2]), int(shape[-1])


def _taps_ptr(arr: np.ndarray) -> int:
    """Address of a (contiguous, caller-kept-alive) tap array; ctypes converts the integer for a c_void_p argument."""
    return arr.__array_interface__["data"][0]


# ---------------------------------------------------------------------------
# 1-D
# ---------------------------------------------------------------------------
def fir1d_fixed(x, h, frac_bits: int = 12, acc_bits: int = 32, coeff_bits: int = 16, *,
                dilation: int = 1, out_dtype=None, out=None, device: int = -1,
                halo_left_valid: int = 0, halo_right_valid: int = 0, validate: bool = True,
                centre=None, in_range=(0, 255), out_range=(0, 255)):
    """Bit-true fixed-point FIR (fir_1d/model/python/fir_1d_fixed_ref.py:12-130)
    along the last axis of ``x``.

    ``dilation=2`` filters interleaved I/Q.  ``halo_*_valid`` say how many
    scalars before / after every row are real neighbour samples the kernel may
    read instead of assuming zeros (device inputs only: a shard of a longer
    signal).  Returns uint8 (or ``out_dtype=np.int16``: same values in an int16
    container) of the same shape, on the same side (host / device) as ``x``.

    ``centre`` (taps ahead of the output; default ``len(h) // 2``), ``in_range``
    and ``out_range`` turn the reference's fixed choices into parameters (the
    extended C entry points): ``centre=0`` is the causal alignment,
    ``in_range = out_range = (-32768, 32767)`` the signed int16 path.  The
    defaults ARE the reference model.
    """
    # A cache hit means this exact tap set already passed every coefficient-side
    # check; the reference's order (taps, samples, bit widths, Q range) is kept.
    key = _tap_key(h, frac_bits, acc_bits, coeff_bits)
    hq = _TAP_CACHE.get(key) if key is not None else None
    if hq is None:
        quant.validate_taps(h)
    in_range, out_range = (int(in_range[0]), int(in_range[1])), (int(out_range[0]), int(out_range[1]))
    if out_dtype is None:
        out_dtype = np.uint8 if (out_range[0] >= 0 and out_range[1] <= 255) else np.int16
    if _is_torch(x):
        if validate:
            _check_finite_torch(x)
    elif not isinstance(x, DeviceArray):
        x = _np_container(x, *in_range)
        if validate:
            quant.check_samples_finite(x)
    if hq is None:
        quant.check_bits(frac_bits, acc_bits, coeff_bits)
        hq = quant.quantize_taps(h, frac_bits, coeff_bits)
        _remember_taps(key, hq)
    ex = None
    if (centre is not None and centre != len(hq) // 2) or in_range != (0, 255) or out_range != (0, 255):
        ex = (len(hq) // 2 if centre is None else int(centre),) + in_range + out_range
    return fir1d_fixed_q(x, hq, frac_bits, acc_bits, dilation=dilation, out_dtype=out_dtype, out=out,
                         device=device, halo_left_valid=halo_left_valid, halo_right_valid=halo_right_valid, ex=ex)


def fir1d_fixed_plan(h, frac_bits: int = 12, acc_bits: int = 32, coeff_bits: int = 16) -> dict:
    """Which datapath would serve these taps on uint8 / int16 samples — answered on the host from the quantised taps
    alone (``firgpu_fir1d_fixed_plan``): ``{"path": "staged" | "generic", "post": "shift" | "hi16" | "wrap", "mac":
    "dp2a" | "imad"}``.  Between 66 and 129 taps the answer is the one for int16 samples (uint8 samples: generic)."""
    quant.validate_taps(h)
    quant.check_bits(frac_bits, acc_bits, coeff_bits)
    hq = np.ascontiguousarray(quant.quantize_taps(h, frac_bits, coeff_bits), dtype=np.int32)
    path, post, mac = C.c_int(0), C.c_int(0), C.c_int(0)
    _lib.check(_lib.lib().firgpu_fir1d_fixed_plan(_taps_ptr(hq), hq.size, frac_bits, acc_bits, C.byref(path), C.byref(post), C.byref(mac)))
    staged = path.value == _lib.PATH_STAGED
    return {"path": "staged" if staged else "generic", "post": ("shift", "hi16", "wrap")[post.value] if staged else None,
            "mac": ("dp2a", "imad")[mac.value] if staged else None}


def fir1d_fixed_q(x, hq, frac_bits: int, acc_bits: int, *, dilation: int = 1, out_dtype=np.uint8, out=None,
                  device: int = -1, halo_left_valid: int = 0, halo_right_valid: int = 0, ex=None):
    """Same as :func:`fir1d_fixed` with already-quantised integer taps and no
    sample validation: the thin layer over ``firgpu_fir1d_fixed[_host]``
    (``ex = (centre, in_lo, in_hi, out_lo, out_hi)`` selects the ``_ex`` forms)."""
    L = _lib.lib()
    hq = np.ascontiguousarray(hq, dtype=np.int32).reshape(-1)
    out_code = _OUT_FIXED[np.dtype(out_dtype)]
    if isinstance(x, DeviceArray):
        rows, cols = _rows_cols(x.shape)
        if not isinstance(out, DeviceArray):
            raise ValueError("DeviceArray inputs need an explicit out=DeviceArray")
        if np.dtype(out.dtype) not in _OUT_FIXED or tuple(out.shape) != tuple(x.shape) or out.device != x.device:
            raise ValueError(f"out must be a uint8 / int16 DeviceArray of shape {tuple(x.shape)} on device {x.device}")
        args = (x.ptr, out.ptr, _NP_CODE[np.dtype(x.dtype)], _OUT_FIXED[np.dtype(out.dtype)], rows, cols, cols, cols,
                _taps_ptr(hq), hq.size, dilation, frac_bits, acc_bits, halo_left_valid, halo_right_valid)
        with _device_scope(x.device):
            if ex is None:
                _lib.check(L.firgpu_fir1d_fixed(*args, x.stream))
            else:
                _lib.check(L.firgpu_fir1d_fixed_ex(*args, *ex, x.stream))
        return out
    if _is_torch(x) and x.is_cuda:
        import torch
        t, rows, cols, ld = _cuda_rows(x, not (halo_left_valid or halo_right_valid), *((ex[1], ex[2]) if ex else (0, 255)))
        y = _cuda_out(out, x, out_dtype)
        _, y_rows, y_cols, ld_y = _cuda_rows(y, allow_copy=False)
        if (y_rows, y_cols) != (rows, cols):
            raise ValueError(f"out has shape {tuple(y.shape)}, expected {tuple(x.shape)}")
        with _on_device(torch, x.device):
            stream = _raw_stream(torch, x.device)
            if ex is None:
                _lib.check(L.firgpu_fir1d_fixed(t.data_ptr(), y.data_ptr(), _torch_code(t), _OUT_FIXED[np.dtype(out_dtype)],
                                                rows, cols, ld, ld_y, _taps_ptr(hq), hq.size, dilation, frac_bits,
                                                acc_bits, halo_left_valid, halo_right_valid, stream))
            else:
                _lib.check(L.firgpu_fir1d_fixed_ex(t.data_ptr(), y.data_ptr(), _torch_code(t), _OUT_FIXED[np.dtype(out_dtype)],
                                                   rows, cols, ld, ld_y, _taps_ptr(hq), hq.size, dilation, frac_bits,
                                                   acc_bits, halo_left_valid, halo_right_valid, *ex, stream))
        return y
    if halo_left_valid or halo_right_valid:
        raise ValueError("halo_*_valid apply to device-resident shards only")
    was_torch = _is_torch(x)
    a = _np_container(x.numpy() if was_torch else x, *((ex[1], ex[2]) if ex else (0, 255)))
    rows, cols = _rows_cols(a.shape)
    y_np = _host_out(out, a.shape, out_dtype)
    y = out if out is not None else y_np
    if ex is None:
        _lib.check(L.firgpu_fir1d_fixed_host(a.ctypes.data, y_np.ctypes.data, _NP_CODE[a.dtype], out_code, rows, cols,
                                             cols, cols, _taps_ptr(hq), hq.size, dilation, frac_bits, acc_bits, device))
    else:
        _lib.check(L.firgpu_fir1d_fixed_ex_host(a.ctypes.data, y_np.ctypes.data, _NP_CODE[a.dtype], out_code, rows, cols,
                                                cols, cols, _taps_ptr(hq), hq.size, dilation, frac_bits, acc_bits, *ex,
                                                device))
    if was_torch and out is None:
        import torch
        return torch.from_numpy(y)
    return y


def fir1d_ideal(x, h, *, dilation: int = 1, out_dtype=np.float64, out=None, device: int = -1,
                halo_left_valid: int = 0, halo_right_valid: int = 0, validate: bool = True,
                centre=None, raw: bool = False, in_range=(0, 255)):
    """Ideal floating-point FIR (fir_1d/model/python/fir_1d_ref.py:43-65) along
    the last axis.  ``out_dtype=np.float64`` is bit-identical to the reference
    (unfused multiply and add, taps in ascending order); ``np.float32`` runs a
    float32 datapath (|err| <= 1e-6 * 255 * sum|h|).  ``centre=0`` is the
    causal alignment and ``raw=True`` filters the samples as given (no round /
    clamp): together the C++ ``Fir1D`` model (fir_1d/model/cpp/fir_1d_ref.cpp)."""
    quant.validate_taps(h)
    L = _lib.lib()
    hh = np.ascontiguousarray([float(v) for v in h], dtype=np.float64)
    out_code = _OUT_IDEAL[np.dtype(out_dtype)]
    in_range = (int(in_range[0]), int(in_range[1]))
    ex = None
    if (centre is not None and centre != hh.size // 2) or raw or in_range != (0, 255):
        ex = (hh.size // 2 if centre is None else int(centre),) + in_range + (1 if raw else 0,)
    if _is_torch(x) and x.is_cuda:
        import torch
        if validate:
            _check_finite_torch(x)
        if raw and x.dtype not in (torch.uint8, torch.int16, torch.int32, torch.float32, torch.float64):
            x = x.to(torch.float64)              # raw samples of an exotic dtype: exact in float64, never clamped
        t, rows, cols, ld = _cuda_rows(x, not (halo_left_valid or halo_right_valid), *in_range)
        y = _cuda_out(out, x, out_dtype)
        _, y_rows, y_cols, ld_y = _cuda_rows(y, allow_copy=False)
        if (y_rows, y_cols) != (rows, cols):
            raise ValueError(f"out has shape {tuple(y.shape)}, expected {tuple(x.shape)}")
        with _on_device(torch, x.device):
            stream = _raw_stream(torch, x.device)
            if ex is None:
                _lib.check(L.firgpu_fir1d_ideal(t.data_ptr(), y.data_ptr(), _torch_code(t), out_code, rows, cols, ld, ld_y,
                                                _taps_ptr(hh), hh.size, dilation, halo_left_valid, halo_right_valid, stream))
            else:
                _lib.check(L.firgpu_fir1d_ideal_ex(t.data_ptr(), y.data_ptr(), _torch_code(t), out_code, rows, cols, ld, ld_y,
                                                   _taps_ptr(hh), hh.size, dilation, halo_left_valid, halo_right_valid,
                                                   *ex, stream))
        return y
    if halo_left_valid or halo_right_valid:
        raise ValueError("halo_*_valid apply to device-resident shards only")
    was_torch = _is_torch(x)
    a = _np_container(x.numpy() if was_torch else x, *in_range, raw=raw)
    if validate:
        quant.check_samples_finite(a)
    rows, cols = _rows_cols(a.shape)
    y = _host_out(out, a.shape, out_dtype)
    if ex is None:
        _lib.check(L.firgpu_fir1d_ideal_host(a.ctypes.data, y.ctypes.data, _NP_CODE[a.dtype], out_code, rows, cols, cols,
                                             cols, _taps_ptr(hh), hh.size, dilation, device))
    else:
        _lib.check(L.firgpu_fir1d_ideal_ex_host(a.ctypes.data, y.ctypes.data, _NP_CODE[a.dtype], out_code, rows, cols, cols,
                                                cols, _taps_ptr(hh), hh.size, dilation, *ex, device))
    if was_torch and out is None:
        import torch
        return torch.from_numpy(y)
    return y


def fir1d_fixed_iq(z, h, frac_bits: int = 12, acc_bits: int = 32, coeff_bits: int = 16, **kw):
    """Interleaved complex baseband [I0,Q0,I1,Q1,...] with real taps: Re and
    Im are filtered independently (the only product lib/mycomplex.cpp:12-22
    defines is complex x real).  ``z`` may also have a trailing axis of 2."""
    shape = tuple(z.shape)
    if len(shape) >= 2 and shape[-1] == 2:
        flat = z.reshape(shape[:-2] + (shape[-2] * 2,))
        return fir1d_fixed(flat, h, frac_bits, acc_bits, coeff_bits, dilation=2, **kw).reshape(shape)
    return fir1d_fixed(z, h, frac_bits, acc_bits, coeff_bits, dilation=2, **kw)


def fir1d_fixed_complex(z, h, frac_bits: int = 12, acc_bits: int = 32, coeff_bits: int = 16, *, in_range=(0, 255),
                        out_range=(0, 255), out_dtype=None, out=None, device: int = -1, validate: bool = True):
    """Interleaved complex baseband ``[I0,Q0,I1,Q1,...]`` (or a trailing axis of 2) filtered with COMPLEX taps ``h``:
    ``y[n] = sum_k h[k] * z[n - (k - L//2)]`` in the fixed-point model — Re and Im of the taps quantised like the 1-D
    model's taps, one accumulator per output component, the 1-D post stage on each (the reference defines
    no complex product, so this specification is ours: parity unpinned except for purely real taps, where it equals
    :func:`fir1d_fixed_iq`).  ``in_range = out_range = (-32768, 32767)`` is signed baseband on int16 containers."""
    hc = [complex(v) for v in h]
    quant.validate_taps([v.real for v in hc])
    quant.validate_taps([v.imag for v in hc])
    quant.check_bits(frac_bits, acc_bits, coeff_bits)
    hq_re = np.ascontiguousarray(quant.quantize_taps([v.real for v in hc], frac_bits, coeff_bits), dtype=np.int32)
    hq_im = np.ascontiguousarray(quant.quantize_taps([v.imag for v in hc], frac_bits, coeff_bits), dtype=np.int32)
    in_range, out_range = (int(in_range[0]), int(in_range[1])), (int(out_range[0]), int(out_range[1]))
    if out_dtype is None:
        out_dtype = np.uint8 if (out_range[0] >= 0 and out_range[1] <= 255) else np.int16
    out_code = _OUT_FIXED[np.dtype(out_dtype)]
    shape = tuple(z.shape)
    pairs = len(shape) >= 2 and shape[-1] == 2
    L = _lib.lib()
    if _is_torch(z) and z.is_cuda:
        import torch
        if validate:
            _check_finite_torch(z)
        flat = z.reshape(shape[:-2] + (shape[-2] * 2,)) if pairs else z
        t, rows, cols, ld = _cuda_rows(flat, True, *in_range)
        if cols % 2:
            raise ValueError("interleaved I/Q needs an even number of scalars per row")
        y = _cuda_out(out.reshape(flat.shape) if (out is not None and pairs) else out, flat, out_dtype)
        _, y_rows, y_cols, ld_y = _cuda_rows(y, allow_copy=False)
        if (y_rows, y_cols) != (rows, cols):
            raise ValueError(f"out has shape {tuple(y.shape)}, expected {tuple(flat.shape)}")
        with _on_device(torch, z.device):
            _lib.check(L.firgpu_fir1d_fixed_complex(t.data_ptr(), y.data_ptr(), _torch_code(t), out_code, rows, cols, ld, ld_y,
                                                    _taps_ptr(hq_re), _taps_ptr(hq_im), hq_re.size, frac_bits, acc_bits, 0, 0,
                                                    *in_range, *out_range, _raw_stream(torch, z.device)))
        return y.reshape(shape) if pairs else y
    was_torch = _is_torch(z)
    a = _np_container(z.numpy() if was_torch else z, *in_range)
    if validate:
        quant.check_samples_finite(a)
    flat = a.reshape(a.shape[:-2] + (a.shape[-2] * 2,)) if pairs else a
    rows, cols = _rows_cols(flat.shape)
    if cols % 2:
        raise ValueError("interleaved I/Q needs an even number of scalars per row")
    y_np = _host_out(out, a.shape, out_dtype)
    _lib.check(L.firgpu_fir1d_fixed_complex_host(flat.ctypes.data, y_np.ctypes.data, _NP_CODE[flat.dtype], out_code, rows, cols, cols,
                                                 cols, _taps_ptr(hq_re), _taps_ptr(hq_im), hq_re.size, frac_bits, acc_bits,
                                                 *in_range, *out_range, device))
    if was_torch and out is None:
        import torch
        return torch.from_numpy(y_np)
    return out if out is not None else y_np


def compare_metrics(x, h, frac_bits: int = 12, acc_bits: int = 32, coeff_bits: int = 16, *, device: int = -1,
                    validate: bool = True) -> dict:
    """Fixed-vs-ideal report metrics of one case, fused: both models are
    evaluated per sample in one pass over ``x`` ([rows, cols] or 1-D) and only
    seven reductions leave the GPU.  Same keys and definitions as the
    reference's ``_compute_metrics`` (fir_1d/sim/vector/
    gen_3tap_compare_report.py:67-112); float64 sums agree with NumPy's to about
    1e-12 relative (different summation order), counts are exact."""
    quant.validate_taps(h)
    is_t = _is_torch(x)
    if is_t:
        if validate:
            _check_finite_torch(x)
    else:
        x = _np_container(x)
        if validate:
            quant.check_samples_finite(x)
    quant.check_bits(frac_bits, acc_bits, coeff_bits)
    hq = np.ascontiguousarray(quant.quantize_taps(h, frac_bits, coeff_bits), dtype=np.int32)
    hh = np.ascontiguousarray([float(v) for v in h], dtype=np.float64)
    out = np.zeros(7, dtype=np.float64)
    L = _lib.lib()
    if is_t and x.is_cuda:
        import torch
        t, rows, cols, ld = _cuda_rows(x, allow_copy=True)
        with _on_device(torch, x.device):
            stream = _raw_stream(torch, x.device)
            _lib.check(L.firgpu_fir1d_compare(t.data_ptr(), _torch_code(t), rows, cols, ld, _taps_ptr(hq), _taps_ptr(hh),
                                              hq.size, frac_bits, acc_bits, out.ctypes.data, stream))
    else:
        a = _np_container(x.numpy() if is_t else x)
        rows, cols = _rows_cols(a.shape)
        _lib.check(L.firgpu_fir1d_compare_host(a.ctypes.data, _NP_CODE[a.dtype], rows, cols, cols, _taps_ptr(hq),
                                               _taps_ptr(hh), hq.size, frac_bits, acc_bits, out.ctypes.data, device))
    n = rows * cols
    if n == 0:
        return {"num_samples": 0, "max_abs_err": 0.0, "mae": 0.0, "rmse": 0.0, "mean_err": 0.0, "sat_low_ratio": 0.0,
                "sat_high_ratio": 0.0, "sat_ratio": 0.0, "clip_needed_ratio": 0.0}
    low, high = float(out[4]) / n, float(out[5]) / n
    return {"num_samples": int(n), "max_abs_err": float(out[3]), "mae": float(out[0]) / n,
            "rmse": float(np.sqrt(out[1] / n)), "mean_err": float(out[2]) / n, "sat_low_ratio": low,
            "sat_high_ratio": high, "sat_ratio": low + high, "clip_needed_ratio": float(out[6]) / n}


# ---------------------------------------------------------------------------
# 2-D
# ---------------------------------------------------------------------------
def fir2d_fixed(x, h2d, frac_bits: int = 12, acc_bits: int = 32, coeff_bits: int = 16, *, out_dtype=np.uint8,
                out=None, device: int = -1, halo_top_valid: int = 0, halo_bottom_valid: int = 0,
                validate: bool = True):
    """2-D fixed-point FIR over the last two axes (specification: DESIGN.md,
    "fir_2d"; the reference ships no fir_2d).  One accumulator per pixel over
    the whole window, then the 1-D model's wrap / round / shift / saturate."""
    hh = quant.validate_taps_2d(h2d)
    if _is_torch(x):
        if validate:
            _check_finite_torch(x)
    else:
        x = _np_container(x)
        if validate:
            quant.check_samples_finite(x)
    quant.check_bits(frac_bits, acc_bits, coeff_bits)
    hq = quant.quantize_taps(hh, frac_bits, coeff_bits)
    return fir2d_fixed_q(x, hq, frac_bits, acc_bits, out_dtype=out_dtype, out=out, device=device,
                         halo_top_valid=halo_top_valid, halo_bottom_valid=halo_bottom_valid)


def fir2d_fixed_q(x, hq2d, frac_bits: int, acc_bits: int, *, out_dtype=np.uint8, out=None, device: int = -1,
                  halo_top_valid: int = 0, halo_bottom_valid: int = 0):
    """:func:`fir2d_fixed` with already-quantised integer taps ``[kh, kw]`` and
    no validation: the thin layer over ``firgpu_fir2d_fixed[_host]``."""
    hq = np.ascontiguousarray(hq2d, dtype=np.int32)
    kh, kw = hq.shape
    L = _lib.lib()
    out_code = _OUT_FIXED[np.dtype(out_dtype)]
    if _is_torch(x) and x.is_cuda:
        import torch
        t = _torch_container(x) if not (halo_top_valid or halo_bottom_valid) else x
        n, hgt, wid = _imgs(t.shape)
        ld = t.stride(-2) if t.dim() >= 2 else wid
        img_stride = t.stride(-3) if t.dim() >= 3 else hgt * ld
        y = _cuda_out(out, t, out_dtype)
        if not y.is_contiguous() or tuple(y.shape) != tuple(t.shape):
            raise ValueError("out must be a contiguous tensor of the input's shape")
        with _on_device(torch, x.device):
            stream = _raw_stream(torch, x.device)
            _lib.check(L.firgpu_fir2d_fixed(t.data_ptr(), y.data_ptr(), _torch_code(t), out_code, n, hgt, wid, ld, wid,
                                            img_stride, hgt * wid, _taps_ptr(hq), kh, kw, frac_bits, acc_bits,
                                            halo_top_valid, halo_bottom_valid, stream))
        return y
    if halo_top_valid or halo_bottom_valid:
        raise ValueError("halo_*_valid apply to device-resident shards only")
    was_torch = _is_torch(x)
    a = _np_container(x.numpy() if was_torch else x)
    n, hgt, wid = _imgs(a.shape)
    y = _host_out(out, a.shape, out_dtype)
    _lib.check(L.firgpu_fir2d_fixed_host(a.ctypes.data, y.ctypes.data, _NP_CODE[a.dtype], out_code, n, hgt, wid,
                                         _taps_ptr(hq), kh, kw, frac_bits, acc_bits, device))
    if was_torch and out is None:
        import torch
        return torch.from_numpy(y)
    return y


def fir1d_fixed_multi(x, h_sets, frac_bits: int = 12, acc_bits: int = 32, coeff_bits: int = 16, *, out_dtype=np.uint8,
                      device: int = -1, validate: bool = True):
    """Several tap sets (all of one length) over ONE pass of the input — the
    coefficient loop of the output-vector generators
    (fir_1d/sim/vector/gen_fixed_output.py:91-104) as a single call.  Returns an
    array of shape ``[len(h_sets)] + x.shape``; groups of four tap sets are
    served by one kernel launch that reads ``x`` from HBM once."""
    h_sets = [list(h) for h in h_sets]
    for h in h_sets:
        quant.validate_taps(h)
    if len({len(h) for h in h_sets}) != 1:
        raise ValueError("all tap sets of one multi-filter call must have the same length")
    is_t = _is_torch(x)
    if is_t:
        if validate:
            _check_finite_torch(x)
    else:
        x = _np_container(x)
        if validate:
            quant.check_samples_finite(x)
    quant.check_bits(frac_bits, acc_bits, coeff_bits)
    hq = np.ascontiguousarray(np.stack([quant.quantize_taps(h, frac_bits, coeff_bits) for h in h_sets]), dtype=np.int32)
    nf, n_taps = hq.shape
    L = _lib.lib()
    out_code = _OUT_FIXED[np.dtype(out_dtype)]
    if is_t and x.is_cuda:
        import torch
        t, rows, cols, ld = _cuda_rows(x, allow_copy=True)
        # every tap set's output starts 256-byte aligned (the kernels' vector stores and the flattened pass for odd
        # image widths need 16): the result is a view [nf, ...] of a buffer with a padded stride between tap sets
        n = rows * cols
        stride = -(-n // 256) * 256
        flat = torch.empty(nf * stride, dtype=_torch_dtype(out_dtype), device=x.device)
        with _on_device(torch, x.device):
            stream = _raw_stream(torch, x.device)
            _lib.check(L.firgpu_fir1d_fixed_multi(t.data_ptr(), flat.data_ptr(), _torch_code(t), out_code, rows, cols, ld,
                                                  cols, stride, _taps_ptr(hq), nf, n_taps, 1, frac_bits, acc_bits,
                                                  stream))
        return flat.view(nf, stride)[:, :n].unflatten(1, tuple(x.shape)) if x.dim() > 0 and n else flat.view(nf, stride)[:, :n].reshape((nf,) + tuple(x.shape))
    a = _np_container(x.numpy() if is_t else x)
    rows, cols = _rows_cols(a.shape)
    y = np.empty((nf,) + a.shape, dtype=out_dtype)
    _lib.check(L.firgpu_fir1d_fixed_multi_host(a.ctypes.data, y.ctypes.data, _NP_CODE[a.dtype], out_code, rows, cols,
                                               cols, cols, rows * cols, _taps_ptr(hq), nf, n_taps, 1, frac_bits,
                                               acc_bits, device))
    if is_t:
        import torch
        return torch.from_numpy(y)
    return y


def fir2d_ideal(x, h2d, *, out_dtype=np.float64, out=None, device: int = -1, validate: bool = True,
                halo_top_valid: int = 0, halo_bottom_valid: int = 0):
    """float64 (bit-identical to the scalar specification) or float32 2-D FIR
    over the last two axes; taps visited row-major, unfused multiply + add.
    CUDA tensors may be column / row views of a larger image (``x`` and ``out``
    alike) and row bands with ``halo_*_valid`` real rows above / below."""
    hh = np.ascontiguousarray(quant.validate_taps_2d(h2d))
    kh, kw = hh.shape
    L = _lib.lib()
    out_code = _OUT_IDEAL[np.dtype(out_dtype)]
    if _is_torch(x) and x.is_cuda:
        import torch
        if validate:
            _check_finite_torch(x)
        native = x.dtype in (torch.uint8, torch.int16, torch.int32, torch.float32, torch.float64)
        in_place = native and not x.is_contiguous() and x.dim() in (2, 3) and x.stride(-1) == 1
        if (halo_top_valid or halo_bottom_valid) and not (native and (x.is_contiguous() or in_place)):
            raise ValueError("halo rows need a tensor that can be read in place")
        t = x if (native and (x.is_contiguous() or in_place)) else _torch_container(x)
        try:
            n, hgt, wid, ld, img_stride = _cuda_imgs(t, "x")
        except ValueError:
            if halo_top_valid or halo_bottom_valid:
                raise
            t = _torch_container(x)                      # odd view without halos: filter a dense copy
            n, hgt, wid, ld, img_stride = _cuda_imgs(t, "x")
        y = out if out is not None else torch.empty(t.shape, dtype=_torch_dtype(out_dtype), device=x.device)
        if tuple(y.shape) != tuple(t.shape) or y.dtype != _torch_dtype(out_dtype) or not y.is_cuda:
            raise ValueError(f"out must be a CUDA {np.dtype(out_dtype)} tensor of the input's shape")
        _, _, _, ld_y, img_stride_y = _cuda_imgs(y, "out")
        with _on_device(torch, x.device):
            stream = _raw_stream(torch, x.device)
            _lib.check(L.firgpu_fir2d_ideal(t.data_ptr(), y.data_ptr(), _torch_code(t), out_code, n, hgt, wid, ld, ld_y,
                                            img_stride, img_stride_y, _taps_ptr(hh), kh, kw, halo_top_valid,
                                            halo_bottom_valid, stream))
        return y
    if halo_top_valid or halo_bottom_valid:
        raise ValueError("halo_*_valid apply to device-resident shards only")
    was_torch = _is_torch(x)
    a = _np_container(x.numpy() if was_torch else x)
    if validate:
        quant.check_samples_finite(a)
    n, hgt, wid = _imgs(a.shape)
    y = _host_out(out, a.shape, out_dtype)
    _lib.check(L.firgpu_fir2d_ideal_host(a.ctypes.data, y.ctypes.data, _NP_CODE[a.dtype], out_code, n, hgt, wid,
                                         _taps_ptr(hh), kh, kw, device))
    if was_torch and out is None:
        import torch
        return torch.from_numpy(y)
    return y


# ---------------------------------------------------------------------------
# several devices, one host array
# ---------------------------------------------------------------------------
def fir1d_fixed_multi_gpu(x, h, frac_bits: int = 12, acc_bits: int = 32, coeff_bits: int = 16, *, devices, dilation: int = 1,
                          out_dtype=np.uint8, out=None, validate: bool = True):
    """:func:`fir1d_fixed` on a HOST array, spread over ``devices`` (CUDA ordinals) by one call of
    ``firgpu_fir1d_fixed_host_multi``: rows split evenly, a single long row cut into segments that read the taps'
    reach from the neighbouring host samples.  No collective, no device-to-device traffic."""
    quant.validate_taps(h)
    a = _np_container(x.numpy() if _is_torch(x) else x)
    if validate:
        quant.check_samples_finite(a)
    quant.check_bits(frac_bits, acc_bits, coeff_bits)
    hq = np.ascontiguousarray(quant.quantize_taps(h, frac_bits, coeff_bits), dtype=np.int32)
    rows, cols = _rows_cols(a.shape)
    y = _host_out(out, a.shape, out_dtype)
    devs = (C.c_int * len(devices))(*[int(d) for d in devices])
    _lib.check(_lib.lib().firgpu_fir1d_fixed_host_multi(a.ctypes.data, y.ctypes.data, _NP_CODE[a.dtype], _OUT_FIXED[np.dtype(out_dtype)],
                                                       rows, cols, cols, cols, _taps_ptr(hq), hq.size, dilation, frac_bits,
                                                       acc_bits, devs, len(devices)))
    return y


# ---------------------------------------------------------------------------
# streaming with the history on the device
# ---------------------------------------------------------------------------
class FirStream:
    """The causal filter ``y[n] = sum_k h[k] x[n - k*dilation]`` over a signal delivered in blocks; the
    ``(len(h) - 1) * dilation`` most recent samples stay in HBM between calls (``firgpu_stream_*``; the reference's
    C++ ``Fir1D`` keeps them in ``reg_``, fir_1d/model/cpp/fir_1d_ref.cpp:16-33).

    ``model="fixed"``: the fixed-point model with integer taps ``hq`` (quantise with :mod:`firgpu.quant`);
    ``model="ideal"``: the floating-point model with real taps.  ``process`` takes a CUDA tensor (device in, device
    out, stream-ordered on torch's current stream) or a host array (pipelined copies)."""

    def __init__(self, taps, *, model: str = "fixed", in_dtype=np.int16, out_dtype=None, dilation: int = 1, frac_bits: int = 12,
                 acc_bits: int = 32, in_range=(0, 255), out_range=(0, 255), raw: bool = False, device: int = -1):
        L = _lib.lib()
        self._h = C.c_void_p()
        self.model, self.in_dtype, self.dilation = model, np.dtype(in_dtype), int(dilation)
        in_code = _NP_CODE[self.in_dtype]
        if model == "fixed":
            self.out_dtype = np.dtype(out_dtype or (np.uint8 if (out_range[0] >= 0 and out_range[1] <= 255) else np.int16))
            self.taps = np.ascontiguousarray(taps, dtype=np.int32).reshape(-1)
            _lib.check(L.firgpu_stream_create_fixed(C.byref(self._h), in_code, _OUT_FIXED[self.out_dtype], _taps_ptr(self.taps),
                                                    self.taps.size, self.dilation, frac_bits, acc_bits, int(in_range[0]),
                                                    int(in_range[1]), int(out_range[0]), int(out_range[1]), device))
        elif model == "ideal":
            self.out_dtype = np.dtype(out_dtype or np.float64)
            self.taps = np.ascontiguousarray([float(v) for v in taps], dtype=np.float64)
            _lib.check(L.firgpu_stream_create_ideal(C.byref(self._h), in_code, _OUT_IDEAL[self.out_dtype], _taps_ptr(self.taps),
                                                    self.taps.size, self.dilation, int(in_range[0]), int(in_range[1]),
                                                    1 if raw else 0, device))
        else:
            raise ValueError("model must be 'fixed' or 'ideal'")

    @property
    def history(self) -> int:
        n = C.c_int64(0)
        _lib.check(_lib.lib().firgpu_stream_history(self._h, C.byref(n)))
        return n.value

    def process(self, x, out=None):
        """Filter the next block.  CUDA tensor in -> CUDA tensor out (asynchronous); host array in -> host array out."""
        L = _lib.lib()
        if _is_torch(x) and x.is_cuda:
            import torch
            if x.dim() != 1 or not x.is_contiguous() or x.dtype != _torch_dtype_any(self.in_dtype):
                raise ValueError(f"a block must be a contiguous 1-D {self.in_dtype} tensor")
            y = _cuda_out(out, x, self.out_dtype)
            if tuple(y.shape) != tuple(x.shape) or not y.is_contiguous():
                raise ValueError("out must be a contiguous tensor of the block's shape")
            with _on_device(torch, x.device):
                _lib.check(L.firgpu_stream_process(self._h, x.data_ptr(), y.data_ptr(), x.numel(),
                                                   _raw_stream(torch, x.device)))
            return y
        a = np.ascontiguousarray(x.numpy() if _is_torch(x) else x, dtype=self.in_dtype).reshape(-1)
        y = _host_out(out, a.shape, self.out_dtype)
        _lib.check(L.firgpu_stream_process_host(self._h, a.ctypes.data, y.ctypes.data, a.size))
        return y

    def reset(self, stream: int = 0) -> None:
        _lib.check(_lib.lib().firgpu_stream_reset(self._h, stream))

    def close(self) -> None:
        if self._h is not None and self._h.value:
            _lib.lib().firgpu_stream_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _torch_dtype_any(np_dtype):
    import torch
    return {np.dtype(np.uint8): torch.uint8, np.dtype(np.int16): torch.int16, np.dtype(np.int32): torch.int32,
            np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}[np.dtype(np_dtype)]


# ---------------------------------------------------------------------------
# recorded launch sequences
# ---------------------------------------------------------------------------
class record:
    """Context manager over ``firgpu_record_begin`` / ``firgpu_record_end``: the firgpu device calls issued inside the
    block on ``stream`` (a raw ``cudaStream_t`` handle, not the NULL stream) by this thread are recorded, not executed;
    ``replay()`` then submits the whole sequence at once (a CUDA graph).  For callers that manage device memory
    themselves (``DeviceArray``, or CUDA tensors with every ``out=`` preallocated): nothing inside the block may
    allocate, and each distinct call must have run once eagerly before.

        with firgpu.record(stream) as rec:
            firgpu.fir1d_fixed_q(x, hq, 15, 32, out_dtype=np.int16, out=y)
        rec.replay()
    """

    def __init__(self, stream: int):
        self.stream = int(stream)
        self._g = C.c_void_p()
        self.kernels = 0

    def __enter__(self):
        _lib.check(_lib.lib().firgpu_record_begin(self.stream))
        return self

    def __exit__(self, exc_type, exc, tb):
        L = _lib.lib()
        if exc_type is not None:
            L.firgpu_record_abort(self.stream)
            return False
        _lib.check(L.firgpu_record_end(self.stream, C.byref(self._g)))
        n = C.c_int64(0)
        _lib.check(L.firgpu_graph_nodes(self._g, C.byref(n)))
        self.kernels = n.value
        return False

    def replay(self, stream: int | None = None) -> None:
        _lib.check(_lib.lib().firgpu_graph_launch(self._g, self.stream if stream is None else int(stream)))

    def close(self) -> None:
        if self._g is not None and self._g.value:
            _lib.lib().firgpu_graph_destroy(self._g)
            self._g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Recording:
    """The same for torch users, with torch's graph-safe allocator: ``fn`` (a callable issuing firgpu device calls
    on CUDA tensors; it may allocate its outputs) runs once eagerly, once under ``torch.cuda.graph``, and never
    again — ``replay()`` resubmits what it launched and returns the tensors the recorded pass produced, refreshed
    from the inputs' current contents.

        rec = firgpu.Recording(lambda: [firgpu.fir1d_fixed_multi(img, taps) for img in images])
        outs = rec.replay()
    """

    def __init__(self, fn):
        import torch
        warm = torch.cuda.Stream()
        warm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(warm):
            fn()                                             # first use configures kernels; not recorded
        torch.cuda.current_stream().wait_stream(warm)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        before = _lib.launch_count()
        with torch.cuda.graph(self.graph):
            self.result = fn()
        self.kernels = _lib.launch_count() - before

    def replay(self):
        self.graph.replay()
        return self.result


# ---------------------------------------------------------------------------
# pinned host buffers without torch
# ---------------------------------------------------------------------------
class PinnedArray:
    """A NumPy array over page-locked memory from ``firgpu_pinned_alloc``; the
    ``*_host`` entry points DMA such buffers directly (no staging copy)."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        self._ptr = C.c_void_p()
        _lib.check(_lib.lib().firgpu_pinned_alloc(C.byref(self._ptr), nbytes))
        buf = (C.c_char * max(nbytes, 1)).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape, dtype=np.int64))).reshape(self.shape)

    def close(self) -> None:
        if self._ptr is not None and self._ptr.value:
            self.array = None
            _lib.lib().firgpu_pinned_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
