"""``fir_2d_ideal(x, h2d)``: float64 2-D FIR over an image (or a batch of
images in the leading axes), centre-aligned, zero outside the image, output
not clamped — the 2-D analogue of fir_1d/model/python/fir_1d_ref.py:43-65."""
from __future__ import annotations

from firgpu import ops as _ops
from firgpu.quant import MAX_ABS_H_COEFF  # noqa: F401


def fir_2d_ideal(x, h2d):
    return _ops.fir2d_ideal(x, h2d)
