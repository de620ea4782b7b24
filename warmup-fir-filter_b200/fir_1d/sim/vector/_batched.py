"""Shared driver of the two output-vector generators.

The reference filters an image one row at a time through a Python call per
row (fir_1d/sim/vector/gen_fixed_output.py:44-59, gen_ideal_output.py:40-49).
Rows are independent and zero-padded separately, so here a whole image is ONE
batched call into libfirgpu ([height, width] = rows x cols).  File discovery,
naming and the skip-unless-overwrite rule are those of
gen_fixed_output.py:19-31,62-107.
"""
from __future__ import annotations

from pathlib import Path
from typing import Callable

import numpy as np

INPUT_SUFFIX = "_x_u8.npy"


def vector_root() -> Path:
    """The directory that holds ``input/`` and ``output/`` for every pipeline stage.  The reference resolves them
    next to its own ``fir_1d/sim/vector`` modules (gen_input_vectors.py:15, gen_fixed_output.py:16-17,
    gen_3tap_compare_report.py:18, restore_images.py:27); in an overlay (``FIR_REFERENCE_ROOT``) the reference's
    untouched stages still do, so the GPU-backed generators must use the same place or the stages would read and
    write different trees.  ``FIR_VECTOR_ROOT`` overrides it for all stages (a read-only reference checkout)."""
    import os
    if os.environ.get("FIR_VECTOR_ROOT"):
        return Path(os.environ["FIR_VECTOR_ROOT"]).resolve()
    ref = os.environ.get("FIR_REFERENCE_ROOT")
    if ref and (Path(ref) / "fir_1d" / "sim" / "vector").is_dir():
        return (Path(ref) / "fir_1d" / "sim" / "vector").resolve()
    return Path(__file__).resolve().parent


def list_input_files(input_dir: Path) -> list[Path]:
    found = [p for p in Path(input_dir).glob("*.npy") if p.name.endswith(INPUT_SUFFIX)]
    found.sort(key=lambda p: p.name.lower())
    return found


def load_image_u8(path: Path) -> np.ndarray:
    x = np.load(path)
    if x.ndim != 2:
        raise ValueError(f"{path.name}: expected 2D array, got shape={x.shape}")
    return x if x.dtype == np.uint8 else x.astype(np.uint8)


def case_stem(path: Path) -> str:
    return path.name[: -len(INPUT_SUFFIX)] if path.name.endswith(INPUT_SUFFIX) else path.stem


def generate_outputs(*, input_dir: Path, out_dir: Path, coeff_map: dict[str, list[float]], name_format: str,
                     run_image: Callable[[np.ndarray, list[float]], np.ndarray], overwrite: bool,
                     run_image_multi: Callable[[np.ndarray, list[list[float]]], np.ndarray] | None = None) -> int:
    """For every input image x every coefficient set that is missing (or all,
    with ``overwrite``): filter, save, count.  ``name_format`` receives
    ``stem`` and ``coeff``.  With ``run_image_multi`` all missing coefficient
    sets of an image are computed by one call (the image is read once)."""
    inputs = list_input_files(input_dir)
    if not inputs:
        raise FileNotFoundError(f"No input .npy files found in {input_dir}")
    out_dir.mkdir(parents=True, exist_ok=True)
    written = 0
    for path in inputs:
        todo = []
        for coeff_name, taps in coeff_map.items():
            target = out_dir / name_format.format(stem=case_stem(path), coeff=coeff_name)
            if overwrite or not target.exists():
                todo.append((target, taps))
        if not todo:
            continue
        image = load_image_u8(path)
        if run_image_multi is not None and len(todo) > 1:
            outputs = run_image_multi(image, [taps for _, taps in todo])
            for (target, _), y in zip(todo, outputs):
                np.save(target, y)
                written += 1
            continue
        for target, taps in todo:
            np.save(target, run_image(image, taps))
            written += 1
    return written
