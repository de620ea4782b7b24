"""``fir_2d_fixed_golden(x, h2d, frac_bits=12, acc_bits=32, coeff_bits=16)``:
bit-true 2-D fixed-point FIR with the 1-D model's pre- and post-processing
(fir_1d/model/python/fir_1d_fixed_ref.py:79-81,110-126) and one accumulator
per pixel over the whole window.  ``fir_2d_separable_fixed`` is the row pass
followed by the column pass (two roundings), i.e. two reference 1-D models."""
from __future__ import annotations

import numpy as np

from firgpu import ops as _ops

MAX_ABS_H_COEFF = 8.0


def fir_2d_fixed_golden(x, h2d, frac_bits: int = 12, acc_bits: int = 32, coeff_bits: int = 16):
    return _ops.fir2d_fixed(x, h2d, frac_bits, acc_bits, coeff_bits)


def fir_2d_separable_fixed(x, h_row, h_col, frac_bits: int = 12, acc_bits: int = 32, coeff_bits: int = 16):
    rows_done = _ops.fir1d_fixed(x, h_row, frac_bits, acc_bits, coeff_bits)
    col_kernel = [[float(v)] for v in h_col]
    return _ops.fir2d_fixed(rows_done, col_kernel, frac_bits, acc_bits, coeff_bits)
