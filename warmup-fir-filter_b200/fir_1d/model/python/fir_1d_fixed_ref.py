"""Drop-in for the reference's ``fir_1d.model.python.fir_1d_fixed_ref``.

``fir_1d_fixed_golden`` keeps the reference signature, defaults, error texts
and return type (fir_1d/model/python/fir_1d_fixed_ref.py:12-130).  Validation
and the float64 tap quantisation run on the host (``firgpu.quant``); the
clamp, multiply-accumulate, acc_bits wrap, rounding shift and saturation run
in libfirgpu on the GPU.  There is no CPU fallback.
"""
from __future__ import annotations

import numpy as np
import numpy.typing as npt

from firgpu import ops as _ops

from .fir_1d_ref import _clamp_x, _round_half_up_x, _validate_h_coefficients, _validate_x  # noqa: F401

MAX_ABS_H_COEFF = 8.0


def fir_1d_fixed_golden(x, h, frac_bits: int = 12, acc_bits: int = 32, coeff_bits: int = 16) -> npt.NDArray[np.uint8]:
    """Bit-true fixed-point FIR of one signal: uint8 pixels (any finite input is
    rounded half up and clamped to [0,255]) x signed Q-format taps, wrapping
    ``acc_bits`` accumulator, round, shift, saturate to uint8.

    Sequences / NumPy arrays return ``np.ndarray[uint8]`` of ``len(x)``; a CUDA
    tensor returns a CUDA uint8 tensor."""
    return _ops.fir1d_fixed(x, h, frac_bits, acc_bits, coeff_bits)
