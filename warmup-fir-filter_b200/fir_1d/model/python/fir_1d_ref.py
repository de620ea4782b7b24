"""Drop-in for the reference's ``fir_1d.model.python.fir_1d_ref``.

``fir_1d_ideal(x, h)`` keeps the reference signature and return type
(fir_1d/model/python/fir_1d_ref.py:43-65) but evaluates the float64 model on
the GPU (``firgpu_fir1d_ideal``: unfused multiply and add in ascending tap
order, so results are bit-identical).  The small validators keep their names
because the fixed-point module and user code import them
(fir_1d/model/python/fir_1d_fixed_ref.py:6).
"""
from __future__ import annotations

import math
from collections.abc import Sequence

from firgpu import ops as _ops
from firgpu import quant as _quant

MAX_ABS_H_COEFF = _quant.MAX_ABS_H_COEFF


def _validate_h_coefficients(h: Sequence[float]) -> None:
    _quant.validate_taps(h)


def _validate_x(x: Sequence[int | float]) -> list[int | float]:
    """Host-list form of the sample check (fir_1d_ref.py:27-33)."""
    out = list(x)
    for index, sample in enumerate(out):
        if not math.isfinite(sample):
            raise ValueError(f"Invalid x[{index}]={sample}: x must be finite.")
    return out


def _round_half_up_x(x: Sequence[int | float]) -> list[int]:
    return [math.floor(sample + 0.5) for sample in x]


def _clamp_x(x: Sequence[int]) -> list[int]:
    return [0 if s < 0 else 255 if s > 255 else s for s in x]


def fir_1d_ideal(x, h: Sequence[float]):
    """Same-mode, centre-aligned, zero-padded float64 FIR; output not clamped.
    Sequences and NumPy arrays return ``list[float]`` like the reference; a
    CUDA tensor returns a CUDA float64 tensor (no host round trip)."""
    y = _ops.fir1d_ideal(x, h)
    if _ops._is_torch(y):
        return y if y.is_cuda else y.tolist()
    return y.tolist()
