"""``fir_1d.sim``: the GPU-backed vector generators live in ``vector``.  When
``FIR_REFERENCE_ROOT`` points at a checkout of the reference, its ``fir_1d/sim``
directory is appended to this package's search path, so ``fir_1d.sim.tests``
(the reference's own pytest suite) imports next to — and runs against — the
GPU-backed modules."""
import os as _os
from pathlib import Path as _Path

_root = _os.environ.get("FIR_REFERENCE_ROOT")
if _root:
    _extra = _Path(_root) / "fir_1d" / "sim"
    if _extra.is_dir() and str(_extra) not in __path__:
        __path__.append(str(_extra))
