"""Vectorised NumPy (int64 / float64) restatement of the reference FIR models.

TEST INFRASTRUCTURE ONLY — see the header of ``oracle/fir_oracle_py.py`` for
the rules (never on the product path; only tests / smoke() / bench.py's CPU
baseline legs import it) and for the parity status (1-D pinned against the
reference's KATs and reference-made golden vectors; I/Q pinned by composition;
fir_2d PARITY UNPINNED — the reference has no fir_2d).

Same arithmetic as the scalar oracle, expressed as shift-and-add over whole
arrays so that 2**28-sample and 16384x16384 inputs finish in seconds.  int64
is exact here as long as |sum_k hq[k] * x| < 2**62, i.e. for every tap count
up to MAX_TAPS with 32-bit coefficients (255 * 2**31 * 1024 < 2**50).
"""
from __future__ import annotations

import numpy as np

MAX_TAPS = 1024
PIXEL_LO, PIXEL_HI = 0, 255          # fir_1d_fixed_ref.py:50-51


def prep_samples(x, lo: int = PIXEL_LO, hi: int = PIXEL_HI) -> np.ndarray:
    """round-half-up (floor(x + 0.5) in the precision of the container: float64 for lists and float64 arrays) then clamp to [lo, hi]
    (fir_1d/model/python/fir_1d_ref.py:35-41; the reference's range is
    [0, 255]).  Integer inputs skip the rounding, which is the identity on
    them.  Returns uint8 for the reference range, int64 otherwise."""
    a = np.asarray(x)
    out_t = np.uint8 if (lo, hi) == (PIXEL_LO, PIXEL_HI) else np.int64
    if a.dtype == np.uint8 and out_t == np.uint8:
        return a
    if a.dtype.kind in "iub":
        return np.clip(a.astype(np.int64), lo, hi).astype(out_t)
    if a.dtype.kind == "f" and a.dtype.itemsize < 8:
        # the reference adds 0.5 to the array's own scalars: `np.float32 + 0.5` is a float32 add under NumPy 2 promotion
        # (fir_1d_ref.py:37, `math.floor(sample + 0.5)`), which differs from float64 at one value per format, the largest
        # one below 0.5 (it ties up to 1.0)
        r = np.floor(a + a.dtype.type(0.5)).astype(np.float64)
    else:
        r = np.floor(a.astype(np.float64) + 0.5)
    return np.clip(r, lo, hi).astype(out_t)


def quantize_taps(h, frac_bits: int, coeff_bits: int) -> np.ndarray:
    """rint(h * 2**frac) -> clip to signed coeff_bits -> integer
    (fir_1d_fixed_ref.py:52-55,75-81)."""
    lo = -(1 << (coeff_bits - 1))
    hi = (1 << (coeff_bits - 1)) - 1
    scaled = np.rint(np.asarray(h, dtype=np.float64) * float(1 << frac_bits))
    return np.clip(scaled, lo, hi).astype(np.int64)


def post_accumulate(acc: np.ndarray, frac_bits: int, acc_bits: int,
                    out_lo: int = PIXEL_LO, out_hi: int = PIXEL_HI) -> np.ndarray:
    """wrap to acc_bits two's complement -> + 2**(frac-1), not re-wrapped ->
    arithmetic shift -> saturate (fir_1d_fixed_ref.py:110-126).  ``acc`` is
    int64 and is consumed."""
    if acc_bits < 63:
        acc &= np.int64((1 << acc_bits) - 1)
        acc -= (acc >> np.int64(acc_bits - 1)) << np.int64(acc_bits)
    # acc_bits >= 63: |acc| < 2**62 so the mask + sign restore is the identity.
    if frac_bits > 62:
        # acc + 2**(frac-1) lies in [0, 2**frac) for every |acc| < 2**61.
        acc[...] = 0
    else:
        acc += np.int64(1 << (frac_bits - 1))
        acc >>= np.int64(frac_bits)
    np.clip(acc, out_lo, out_hi, out=acc)
    return acc


def _shifted_views(xs: np.ndarray, n_taps: int, dilation: int, centre: int | None = None):
    """Yield (k, view) where view[..., n] == xs[..., n - (k - centre)*dilation]
    with zeros outside; centre = L//2 is the reference's alignment
    (fir_1d_fixed_ref.py:86,99-104), centre = 0 the causal form of the C++
    Fir1D (fir_1d/model/cpp/fir_1d_ref.cpp:16-33)."""
    centre = n_taps // 2 if centre is None else centre
    n = xs.shape[-1]
    left = (n_taps - 1 - centre) * dilation      # reach towards lower indices
    right = centre * dilation                    # reach towards higher indices
    pad = [(0, 0)] * (xs.ndim - 1) + [(left, right)]
    xp = np.pad(xs, pad)
    for k in range(n_taps):
        start = left - (k - centre) * dilation
        yield k, xp[..., start:start + n]


def fir1d_fixed(x, h, frac_bits: int = 12, acc_bits: int = 32, coeff_bits: int = 16,
                *, dilation: int = 1, taps_are_quantized: bool = False,
                out_dtype=np.uint8, centre: int | None = None,
                in_range=(PIXEL_LO, PIXEL_HI), out_range=(PIXEL_LO, PIXEL_HI)) -> np.ndarray:
    """Fixed-point FIR along the last axis; every leading index is an
    independent, separately zero-padded row (gen_fixed_output.py:44-59).
    ``centre`` / ``in_range`` / ``out_range`` turn the reference's fixed
    choices (L//2, [0,255], [0,255]) into parameters (extended entry points;
    any non-default value is OUR specification, not reference-pinned)."""
    xs = prep_samples(x, *in_range)
    hq = np.asarray(h, dtype=np.int64) if taps_are_quantized else quantize_taps(h, frac_bits, coeff_bits)
    assert 1 <= hq.size <= MAX_TAPS
    acc = np.zeros(xs.shape, dtype=np.int64)
    xi = xs.astype(np.int64)
    for k, view in _shifted_views(xi, hq.size, dilation, centre):
        if hq[k] != 0:
            acc += view * hq[k]
    return post_accumulate(acc, frac_bits, acc_bits, *out_range).astype(out_dtype)


def fir1d_ideal(x, h, *, dilation: int = 1, dtype=np.float64, centre: int | None = None, raw: bool = False,
                in_range=(PIXEL_LO, PIXEL_HI)) -> np.ndarray:
    """float64 model along the last axis: acc = acc + h[k] * xs[...] in
    ascending k, separate multiply and add (fir_1d_ref.py:55-63).  The padded
    zeros contribute +-0.0, which never changes a non-negative-zero sum, so
    including them is bit-identical to the reference's skip."""
    xs = np.asarray(x, dtype=dtype) if raw else prep_samples(x, *in_range).astype(dtype)
    hh = np.asarray(h, dtype=dtype)
    acc = np.zeros(xs.shape, dtype=dtype)
    for k, view in _shifted_views(xs, hh.size, dilation, centre):
        acc = acc + hh[k] * view
    return acc


def fir1d_full_causal(x, h) -> np.ndarray:
    """Fir1D::fir_1d_idel (fir_1d/model/cpp/fir_1d_ref.cpp:37-63): raw float64
    samples, causal alignment, all N + L - 1 outputs = the causal model on the
    signal extended by L - 1 zeros."""
    x = np.asarray(x, dtype=np.float64)
    if x.size == 0 or len(h) == 0:
        return np.zeros(0)
    ext = np.concatenate([x, np.zeros(len(h) - 1)])
    return fir1d_ideal(ext, h, centre=0, raw=True)


def fir1d_fixed_iq(z, h, frac_bits: int = 12, acc_bits: int = 32, coeff_bits: int = 16,
                   out_dtype=np.uint8) -> np.ndarray:
    """Interleaved I/Q with real taps = the 1-D model on z[0::2] and z[1::2]
    (lib/mycomplex.cpp:12-22: complex*double and complex+complex only)."""
    z = np.asarray(z)
    out = np.empty(z.shape, dtype=out_dtype)
    out[..., 0::2] = fir1d_fixed(z[..., 0::2], h, frac_bits, acc_bits, coeff_bits, out_dtype=out_dtype)
    out[..., 1::2] = fir1d_fixed(z[..., 1::2], h, frac_bits, acc_bits, coeff_bits, out_dtype=out_dtype)
    return out


# --------------------------------------------------------------------------
# 2-D (parity unpinned — our spec; see fir_oracle_py.fir2d_fixed)
# --------------------------------------------------------------------------
def fir1d_fixed_complex(z, h, frac_bits: int = 12, acc_bits: int = 32, coeff_bits: int = 16, *, out_dtype=np.int16,
                        in_range=(PIXEL_LO, PIXEL_HI), out_range=(PIXEL_LO, PIXEL_HI)) -> np.ndarray:
    """Interleaved I/Q [I0,Q0,I1,Q1,...] with COMPLEX taps — OUR specification (parity unpinned: lib/mycomplex.h:10-41
    defines complex + complex and complex * double only, so the reference fixes how a real-scalar product and a sum
    act per component, not a complex product).  Written directly from the definition, not as the two real filters the
    product composes it from:

        hq_re[k], hq_im[k] = the 1-D model's quantisation of Re h[k], Im h[k]          (fir_1d_fixed_ref.py:79-81)
        xs = the 1-D model's sample preparation of every scalar                         (fir_1d_ref.py:35-41)
        acc_re[n] = sum_k hq_re[k] * xs_re[m] - hq_im[k] * xs_im[m],   m = n - (k - L//2), zero outside [0, N)
        acc_im[n] = sum_k hq_re[k] * xs_im[m] + hq_im[k] * xs_re[m]
        y_re, y_im = the 1-D post stage on each accumulator (wrap once, + bias, >> frac, saturate)   (:110-126)

    Imaginary taps of zero reduce it to ``fir1d_fixed_iq`` (real taps), which IS pinned by composition."""
    z = np.asarray(z)
    hc = np.asarray(h, dtype=np.complex128).reshape(-1)
    hq_re = quantize_taps(hc.real, frac_bits, coeff_bits)
    hq_im = quantize_taps(hc.imag, frac_bits, coeff_bits)
    xs = prep_samples(z, *in_range).astype(np.int64)
    xr, xi = xs[..., 0::2], xs[..., 1::2]
    acc_re = np.zeros(xr.shape, dtype=np.int64)
    acc_im = np.zeros(xr.shape, dtype=np.int64)
    views_r = dict(_shifted_views(xr, hc.size, 1, None))
    views_i = dict(_shifted_views(xi, hc.size, 1, None))
    for k in range(hc.size):
        acc_re += hq_re[k] * views_r[k] - hq_im[k] * views_i[k]
        acc_im += hq_re[k] * views_i[k] + hq_im[k] * views_r[k]
    out = np.empty(z.shape, dtype=out_dtype)
    out[..., 0::2] = post_accumulate(acc_re, frac_bits, acc_bits, *out_range).astype(out_dtype)
    out[..., 1::2] = post_accumulate(acc_im, frac_bits, acc_bits, *out_range).astype(out_dtype)
    return out


def _shifted_views_2d(xs: np.ndarray, kh: int, kw: int):
    ch, cw = kh // 2, kw // 2
    hgt, wid = xs.shape[-2:]
    top, bottom = kh - 1 - ch, ch
    left, right = kw - 1 - cw, cw
    pad = [(0, 0)] * (xs.ndim - 2) + [(top, bottom), (left, right)]
    xp = np.pad(xs, pad)
    for m in range(kh):
        r0 = top - (m - ch)
        for n in range(kw):
            c0 = left - (n - cw)
            yield m, n, xp[..., r0:r0 + hgt, c0:c0 + wid]


def fir2d_fixed(x, h2d, frac_bits: int = 12, acc_bits: int = 32, coeff_bits: int = 16,
                *, taps_are_quantized: bool = False, out_dtype=np.uint8) -> np.ndarray:
    xs = prep_samples(x).astype(np.int64)
    h2d = np.asarray(h2d)
    hq = h2d.astype(np.int64) if taps_are_quantized else quantize_taps(h2d, frac_bits, coeff_bits)
    kh, kw = hq.shape
    acc = np.zeros(xs.shape, dtype=np.int64)
    for m, n, view in _shifted_views_2d(xs, kh, kw):
        if hq[m, n] != 0:
            acc += view * hq[m, n]
    return post_accumulate(acc, frac_bits, acc_bits).astype(out_dtype)


def fir2d_ideal(x, h2d, dtype=np.float64) -> np.ndarray:
    xs = prep_samples(x).astype(dtype)
    hh = np.asarray(h2d, dtype=dtype)
    kh, kw = hh.shape
    acc = np.zeros(xs.shape, dtype=dtype)
    for m, n, view in _shifted_views_2d(xs, kh, kw):
        acc = acc + hh[m, n] * view
    return acc


def fir2d_separable_fixed(x, h_row, h_col, frac_bits: int = 12, acc_bits: int = 32,
                          coeff_bits: int = 16) -> np.ndarray:
    """Row pass then column pass, each a full 1-D reference model (two
    roundings).  This one IS reference-pinned, by composition."""
    rows_done = fir1d_fixed(x, h_row, frac_bits, acc_bits, coeff_bits)
    cols_done = fir1d_fixed(np.swapaxes(rows_done, -1, -2), h_col, frac_bits, acc_bits, coeff_bits)
    return np.ascontiguousarray(np.swapaxes(cols_done, -1, -2))
