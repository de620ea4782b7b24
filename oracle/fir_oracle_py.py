"""Scalar (pure-Python, unbounded-int) restatement of the reference FIR models.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product
path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the
checker.  The shipped path is the CUDA library behind ``include/firgpu.h``.

Each function cites the reference lines (relative to /root/reference) whose
arithmetic it restates.  The code is organised differently from the reference
(one shared "post" stage, explicit index helpers, parametrised saturation
bounds, a 2-D and an interleaved-I/Q form) but must produce the same numbers;
``tests/test_oracle_vs_reference.py`` pins it against the unmodified reference
whenever /root/reference is importable, and ``tests/golden/*.npz`` pin it
everywhere else.

Parity status: fir_1d fixed / ideal — PINNED (reference KATs + reference-made
golden vectors).  I/Q — pinned by composition (reference run on z[0::2] and
z[1::2]).  fir_2d — PARITY UNPINNED: the reference ships no fir_2d code, tests
or vectors (fir_2d/model/cpp/CMakeLists.txt is 0 bytes); the 2-D functions
below are *our* specification, anchored to the reference only through the
degenerate single-row / single-column kernels and the separable composition.
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np

PIXEL_LO = 0      # fir_1d/model/python/fir_1d_fixed_ref.py:51
PIXEL_HI = 255    # fir_1d/model/python/fir_1d_fixed_ref.py:50


# --------------------------------------------------------------------------
# shared stages
# --------------------------------------------------------------------------
def prep_sample(sample, lo: int = PIXEL_LO, hi: int = PIXEL_HI) -> int:
    """One input sample -> integer pixel.

    Round half up as floor(s + 0.5) evaluated in float64
    (fir_1d/model/python/fir_1d_ref.py:35-38), then clamp to [0, 255]
    (fir_1d_ref.py:40-41).
    """
    rounded = math.floor(sample + 0.5)
    if rounded < lo:
        return lo
    if rounded > hi:
        return hi
    return rounded


def quantize_taps(h: Sequence[float], frac_bits: int, coeff_bits: int) -> list[int]:
    """Real taps -> signed Q-format integers.

    scale by 2**frac_bits, numpy rint (ties-to-even), clip to the signed
    coeff_bits range, cast (fir_1d_fixed_ref.py:52-55,75-81).  Kept on numpy
    on purpose: the reference does exactly these three numpy calls and the
    result of ``h * SCALE`` in float64 is part of the contract.
    """
    lo = -(1 << (coeff_bits - 1))
    hi = (1 << (coeff_bits - 1)) - 1
    scaled = np.rint(np.asarray(h, dtype=np.float64) * float(1 << frac_bits))
    return [int(v) for v in np.clip(scaled, lo, hi).astype(np.int64)]


def wrap_signed(acc: int, acc_bits: int) -> int:
    """Keep the low acc_bits of acc and read them as two's complement
    (fir_1d_fixed_ref.py:93,110-115).  Applied once per output, after the
    whole sum."""
    acc &= (1 << acc_bits) - 1
    if acc >> (acc_bits - 1):
        acc -= 1 << acc_bits
    return acc


def post_accumulate(acc: int, frac_bits: int, acc_bits: int,
                    out_lo: int = PIXEL_LO, out_hi: int = PIXEL_HI) -> int:
    """wrap -> + half LSB (not re-wrapped) -> arithmetic shift -> saturate
    (fir_1d_fixed_ref.py:110-126)."""
    acc = wrap_signed(acc, acc_bits)
    acc += 1 << (frac_bits - 1)
    val = acc >> frac_bits
    return out_lo if val < out_lo else out_hi if val > out_hi else val


# --------------------------------------------------------------------------
# 1-D
# --------------------------------------------------------------------------
def fir1d_fixed(x: Sequence, h: Sequence[float], frac_bits: int = 12,
                acc_bits: int = 32, coeff_bits: int = 16, *,
                dilation: int = 1, centre: int | None = None,
                in_range=(PIXEL_LO, PIXEL_HI), out_range=(PIXEL_LO, PIXEL_HI)) -> np.ndarray:
    """Same-mode, centre-aligned, zero-padded fixed-point FIR of one signal.

    y[n] = post( sum_k hq[k] * xs[n - (k - L//2) * dilation] ), xs = 0 outside
    [0, N)  (fir_1d_fixed_ref.py:83-107; dilation == 1 is the reference, other
    values are the interleaved-channel generalisation used for I/Q).
    """
    xs = [prep_sample(s, *in_range) for s in x]
    hq = quantize_taps(h, frac_bits, coeff_bits)
    n_samples, n_taps = len(xs), len(hq)
    centre = n_taps // 2 if centre is None else centre
    out = np.empty(n_samples, dtype=np.uint8 if tuple(out_range) == (PIXEL_LO, PIXEL_HI) else np.int64)
    for n in range(n_samples):
        acc = 0
        for k in range(n_taps):
            src = n - (k - centre) * dilation
            if 0 <= src < n_samples:
                acc += xs[src] * hq[k]
        out[n] = post_accumulate(acc, frac_bits, acc_bits, *out_range)
    return out


def fir1d_ideal(x: Sequence, h: Sequence[float], *, dilation: int = 1, centre: int | None = None,
                raw: bool = False) -> list[float]:
    """float64 model: same indexing, taps visited in ascending k, one multiply
    then one add per visited tap, out-of-range taps skipped, no output clamp
    (fir_1d/model/python/fir_1d_ref.py:43-65)."""
    xs = [float(s) for s in x] if raw else [prep_sample(s) for s in x]
    n_samples, n_taps = len(xs), len(h)
    centre = n_taps // 2 if centre is None else centre
    out = [0.0] * n_samples
    for n in range(n_samples):
        acc = 0.0
        for k in range(n_taps):
            src = n - (k - centre) * dilation
            if 0 <= src < n_samples:
                acc += h[k] * xs[src]
        out[n] = acc
    return out


def fir1d_fixed_iq(z_interleaved: Sequence, h: Sequence[float], frac_bits: int = 12,
                   acc_bits: int = 32, coeff_bits: int = 16) -> np.ndarray:
    """Complex samples [I0,Q0,I1,Q1,...] with real taps: MyComplex has only
    complex*double and complex+complex (lib/mycomplex.cpp:12-22), so Re and Im
    are filtered independently.  Built by composition of the 1-D model."""
    z = list(z_interleaved)
    out = np.empty(len(z), dtype=np.uint8)
    out[0::2] = fir1d_fixed(z[0::2], h, frac_bits, acc_bits, coeff_bits)
    out[1::2] = fir1d_fixed(z[1::2], h, frac_bits, acc_bits, coeff_bits)
    return out


# --------------------------------------------------------------------------
# 2-D  (parity unpinned: our spec, see module docstring)
# --------------------------------------------------------------------------
def fir2d_fixed(x, h2d, frac_bits: int = 12, acc_bits: int = 32,
                coeff_bits: int = 16) -> np.ndarray:
    """y[i,j] = post( sum_m sum_n hq[m,n] * xs[i-m+Kh//2, j-n+Kw//2] ), zero
    outside the image, ONE accumulator / wrap / round / saturate per pixel.
    Pre- and post-processing are those of the 1-D model
    (fir_1d_fixed_ref.py:79-81,110-126); the double sum is the course formula
    (fir_1d/docs/DSP_2025_2_2nd_assignment_v4.pdf p.14)."""
    rows = [[prep_sample(s) for s in row] for row in x]
    height = len(rows)
    width = len(rows[0]) if height else 0
    kh = len(h2d)
    kw = len(h2d[0])
    hq = [quantize_taps(r, frac_bits, coeff_bits) for r in h2d]
    out = np.empty((height, width), dtype=np.uint8)
    for i in range(height):
        for j in range(width):
            acc = 0
            for m in range(kh):
                si = i - m + kh // 2
                if not 0 <= si < height:
                    continue
                for n in range(kw):
                    sj = j - n + kw // 2
                    if 0 <= sj < width:
                        acc += rows[si][sj] * hq[m][n]
            out[i, j] = post_accumulate(acc, frac_bits, acc_bits)
    return out


def fir2d_ideal(x, h2d) -> np.ndarray:
    """float64 2-D model; taps visited m-major, n-minor, ascending."""
    rows = [[prep_sample(s) for s in row] for row in x]
    height = len(rows)
    width = len(rows[0]) if height else 0
    kh = len(h2d)
    kw = len(h2d[0])
    out = np.zeros((height, width), dtype=np.float64)
    for i in range(height):
        for j in range(width):
            acc = 0.0
            for m in range(kh):
                si = i - m + kh // 2
                if not 0 <= si < height:
                    continue
                for n in range(kw):
                    sj = j - n + kw // 2
                    if 0 <= sj < width:
                        acc += float(h2d[m][n]) * rows[si][sj]
            out[i, j] = acc
    return out
