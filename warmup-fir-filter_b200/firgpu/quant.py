"""Host-side contract of the fixed-point FIR path: argument validation with the
reference's error texts, and the float64 coefficient quantisation.

These steps stay on the host on purpose.  They touch a handful of taps (or do
one vectorised pass over float samples) and their float64 behaviour —
``np.rint`` ties-to-even on ``h * 2**frac_bits`` — is part of the bit-exact
contract (reference: fir_1d/model/python/fir_1d_fixed_ref.py:39-81 and
fir_1d/model/python/fir_1d_ref.py:9-33).  The C ABI takes integer taps.
"""
from __future__ import annotations

import math

import numpy as np

MAX_ABS_H_COEFF = 8.0            # fir_1d_ref.py:6
VALID_COEFF_BITS = (8, 16, 32)   # fir_1d_fixed_ref.py:43


def validate_taps(h) -> None:
    """Empty -> finite -> magnitude, tap by tap in index order
    (fir_1d_ref.py:9-24)."""
    if len(h) == 0:
        raise ValueError("Invalid h: h coefficients must not be empty.")
    for index, coeff in enumerate(h):
        if not math.isfinite(coeff):
            raise ValueError(f"Invalid h[{index}]={coeff}: h coefficients must be finite.")
        if abs(coeff) > MAX_ABS_H_COEFF:
            raise ValueError(f"Invalid h[{index}]={coeff}: |h| must be <= {MAX_ABS_H_COEFF}.")


def validate_taps_2d(h2d) -> np.ndarray:
    """2-D kernels: same three checks on every entry, row-major order; the
    index in the message is the (row, col) pair."""
    rows = [list(r) for r in h2d]
    if len(rows) == 0 or any(len(r) == 0 for r in rows):
        raise ValueError("Invalid h: h coefficients must not be empty.")
    if any(len(r) != len(rows[0]) for r in rows):
        raise ValueError("Invalid h: 2-D h must be rectangular.")
    for i, r in enumerate(rows):
        for j, coeff in enumerate(r):
            if not math.isfinite(coeff):
                raise ValueError(f"Invalid h[{i}][{j}]={coeff}: h coefficients must be finite.")
            if abs(coeff) > MAX_ABS_H_COEFF:
                raise ValueError(f"Invalid h[{i}][{j}]={coeff}: |h| must be <= {MAX_ABS_H_COEFF}.")
    return np.array(rows, dtype=np.float64)


def check_samples_finite(a: np.ndarray) -> None:
    """fir_1d_ref.py:27-33, vectorised: integer containers are always finite;
    for float containers report the first offending flat index."""
    if a.dtype.kind != "f" or a.size == 0:
        return
    ok = np.isfinite(a)
    if not ok.all():
        index = int(np.argmin(ok.reshape(-1)))
        sample = a.reshape(-1)[index]
        raise ValueError(f"Invalid x[{index}]={float(sample)}: x must be finite.")


def check_bits(frac_bits: int, acc_bits: int, coeff_bits: int) -> None:
    """fir_1d_fixed_ref.py:39-47 (same order: frac, acc, coeff)."""
    if frac_bits <= 0:
        raise ValueError(f"Invalid frac_bits={frac_bits}. frac_bits must be > 0.")
    if acc_bits <= 0:
        raise ValueError(f"Invalid acc_bits={acc_bits}. acc_bits must be > 0.")
    if coeff_bits not in VALID_COEFF_BITS:
        raise ValueError(f"Invalid coeff_bits={coeff_bits}. coeff_bits must be one of {VALID_COEFF_BITS}.")


def quantize_taps(h, frac_bits: int, coeff_bits: int) -> np.ndarray:
    """Real taps -> int32 Q-format taps: range check against the representable
    real interval, then rint / clip / integer cast in float64
    (fir_1d_fixed_ref.py:52-58,67-81).  Works for 1-D and 2-D tap arrays; the
    message indexes a flat 1-D ``h`` the way the reference does."""
    lo_int = -(1 << (coeff_bits - 1))
    hi_int = (1 << (coeff_bits - 1)) - 1
    scale = 1 << frac_bits
    lo_real, hi_real = lo_int / scale, hi_int / scale
    arr = np.asarray(h, dtype=np.float64)
    flat = arr.reshape(-1)
    for index in range(flat.size):
        coeff = flat[index]
        if coeff < lo_real or coeff > hi_real:
            where = f"h[{index}]" if arr.ndim == 1 else f"h[{index // arr.shape[1]}][{index % arr.shape[1]}]"
            shown = h[index] if arr.ndim == 1 else float(coeff)
            raise ValueError(f"Invalid {where}={shown}: out of Q-format real range [{lo_real}, {hi_real}].")
    q = np.clip(np.rint(arr * scale), lo_int, hi_int)
    return q.astype(np.int64).astype(np.int32)
