"""Coefficient bank of the verification runs: four 3-tap and four 5-tap
kernels, keyed as in the reference (fir_1d/sim/vector/h_coeff.py:3-16; the
generators iterate these dicts in insertion order, which fixes the order of
the output files)."""
from fractions import Fraction as _F


def _taps(*ratios):
    return [float(_F(r)) if isinstance(r, str) else float(r) for r in ratios]


h_coeff_3tap_map = {
    "moving_avg": [1 / 3] * 3,
    "simple_lp": _taps("1/4", "1/2", "1/4"),
    "edge": [-1.0, 0, 1.0],
    "sharpen": _taps("-1/8", "5/4", "-1/8"),
}

h_coeff_5tap_map = {
    "moving_avg": [1 / 5] * 5,
    "simple_lp": _taps("1/16", "4/16", "6/16", "4/16", "1/16"),
    "edge": [-1 / 8, -2 / 8, 0, 2 / 8, 1 / 8],
    "sharpen": _taps("-1/16", "-4/16", "26/16", "-4/16", "-1/16"),
}
