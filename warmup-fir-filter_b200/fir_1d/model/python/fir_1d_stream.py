"""GPU-backed counterpart of the reference's C++ ``Fir1D`` class
(fir_1d/model/cpp/fir_1d_ref.h:8-27, fir_1d_ref.cpp:7-63): the causal float64
FIR on raw samples, as a stateful streaming object and as the block-wise
"full" convolution.

* ``process_sample(x_n)`` / ``process_block(x)``: ``y[n] = sum_k h[k] * x[n-k]``
  with the (L-1)-sample history carried between calls (``reg_`` in the C++,
  fir_1d_ref.cpp:16-33); ``reset()`` zeroes it.  The history lives in HBM inside a
  ``firgpu_stream_*`` object: a block that is a CUDA tensor is filtered where it
  lies and nothing travels to the host.
* ``Fir1D.fir_1d_idel(x, h)``: all ``N + L - 1`` outputs of the causal model
  (fir_1d_ref.cpp:37-63).

Arithmetic: float64, unfused multiply + add in ascending tap order — the order
of the C++ loops — so results are bit-identical to the C++ compiled without FMA
contraction.  Like the C++ class (and unlike the Python model) nothing is
validated: any tap magnitude is accepted, and NaN / inf samples propagate.

``process_sample`` is here for interface parity; one sample per call is a
host -> device -> host round trip per sample (tens of microseconds) and no way
to use a GPU.  Feed blocks.
"""
from __future__ import annotations

import numpy as np

from firgpu import ops as _ops


class Fir1D:
    def __init__(self, h=None):
        self._h: list[float] = []
        self._stream = None
        if h is not None:
            self.set_taps(h)

    def set_taps(self, h) -> None:
        self._h = [float(v) for v in h]
        if self._stream is not None:
            self._stream.close()
        self._stream = None              # created on first use (needs a device)

    def taps(self) -> list[float]:
        return list(self._h)

    def reset(self) -> None:
        if self._stream is not None:
            self._stream.reset()

    def _state(self):
        if self._stream is None:
            self._stream = _ops.FirStream(self._h, model="ideal", in_dtype=np.float64, out_dtype=np.float64, raw=True)
        return self._stream

    def process_block(self, x):
        """Filter one chunk; outputs line up with the chunk's samples and the
        chunk's tail becomes the history for the next call.  A float64 CUDA
        tensor stays on the device; anything else is treated as a host array."""
        if _ops._is_torch(x) and x.is_cuda:
            if not self._h:
                return x.new_zeros(x.shape)
            return self._state().process(x)
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        if not self._h or x.size == 0:
            return np.zeros(x.size, dtype=np.float64)
        return self._state().process(x)

    def process_sample(self, x_n: float) -> float:
        return float(self.process_block([x_n])[0]) if self._h else 0.0

    @staticmethod
    def fir_1d_idel(x, h) -> np.ndarray:
        """Block model: full convolution, length ``len(x) + len(h) - 1``."""
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        h = [float(v) for v in h]
        if x.size == 0 or not h:
            return np.zeros(0, dtype=np.float64)
        s = _ops.FirStream(h, model="ideal", in_dtype=np.float64, out_dtype=np.float64, raw=True)
        try:
            return s.process(np.concatenate([x, np.zeros(len(h) - 1)]))
        finally:
            s.close()
