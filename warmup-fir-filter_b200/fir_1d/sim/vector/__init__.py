"""Vector generators.  Only the two model-calling generators (and the
coefficient bank) are re-implemented here; the reference's image decoding,
compare reports and PNG restore are file plumbing outside the hot path.  When
``FIR_REFERENCE_ROOT`` points at a checkout of the reference, its
``fir_1d/sim/vector`` directory is appended to this package's search path so
that those untouched modules (``gen_input_vectors``, ``gen_*_compare_report``,
``restore_images``) import under their usual names next to the GPU-backed ones.
"""
import os as _os
from pathlib import Path as _Path

_root = _os.environ.get("FIR_REFERENCE_ROOT")
if _root:
    _extra = _Path(_root) / "fir_1d" / "sim" / "vector"
    if _extra.is_dir() and str(_extra) not in __path__:
        __path__.append(str(_extra))
