"""GPU-backed counterpart of the reference's C++ ``Fir1D`` class
(fir_1d/model/cpp/fir_1d_ref.h:8-27, fir_1d_ref.cpp:7-63): the causal float64
FIR on raw samples, as a stateful streaming object and as the block-wise
"full" convolution.

* ``process_sample(x_n)`` / ``process_block(x)``: ``y[n] = sum_k h[k] * x[n-k]``
  with an (L-1)-sample history carried between calls (``reg_`` in the C++,
  fir_1d_ref.cpp:16-33); ``reset()`` zeroes it.
* ``Fir1D.fir_1d_idel(x, h)``: all ``N + L - 1`` outputs of the causal model
  (fir_1d_ref.cpp:37-63).

Arithmetic runs in ``firgpu_fir1d_ideal_ex`` (centre = 0, raw samples, float64,
unfused multiply + add in ascending tap order — the order of the C++ loops), so
results are bit-identical to the C++ compiled without FMA contraction.
"""
from __future__ import annotations

import numpy as np

from firgpu import ops as _ops


class Fir1D:
    def __init__(self, h=None):
        self._h: list[float] = []
        self._state = np.zeros(0, dtype=np.float64)
        if h is not None:
            self.set_taps(h)

    def set_taps(self, h) -> None:
        self._h = [float(v) for v in h]
        self.reset()

    def taps(self) -> list[float]:
        return list(self._h)

    def reset(self) -> None:
        self._state = np.zeros(max(len(self._h) - 1, 0), dtype=np.float64)

    def process_block(self, x) -> np.ndarray:
        """Filter one chunk; outputs line up with the chunk's samples and the
        chunk's tail becomes the history for the next call."""
        if not self._h:
            return np.zeros(len(x), dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        if x.size == 0:
            return np.zeros(0, dtype=np.float64)
        hist = self._state.size
        window = np.concatenate([self._state, x])
        y = _ops.fir1d_ideal(window, self._h, centre=0, raw=True)[hist:]
        self._state = window[window.size - hist:].copy() if hist else self._state
        return y

    def process_sample(self, x_n: float) -> float:
        return float(self.process_block([x_n])[0]) if self._h else 0.0

    @staticmethod
    def fir_1d_idel(x, h) -> np.ndarray:
        """Block model: full convolution, length ``len(x) + len(h) - 1``."""
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        h = [float(v) for v in h]
        if x.size == 0 or not h:
            return np.zeros(0, dtype=np.float64)
        return _ops.fir1d_ideal(np.concatenate([x, np.zeros(len(h) - 1)]), h, centre=0, raw=True)
