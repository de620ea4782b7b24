"""Drop-in for the reference's ``fir_1d.sim.vector.gen_ideal_output``
(fir_1d/sim/vector/gen_ideal_output.py:37-118): float64 outputs named
``{case}__{coeff}_ideal_{3|5}tap_y_f64.npy`` under ``ideal_{3|5}tap/``, one
batched GPU call per image."""
from __future__ import annotations

import argparse
from pathlib import Path
from time import perf_counter

import numpy as np

from firgpu import ops as _ops
from fir_1d.model.python.fir_1d_ref import fir_1d_ideal  # noqa: F401  (re-exported like the reference)
from fir_1d.sim.vector import _batched
from fir_1d.sim.vector.h_coeff import h_coeff_3tap_map, h_coeff_5tap_map

THIS_FILE = Path(__file__).resolve()
DEFAULT_INPUT_DIR = _batched.vector_root() / "input"      # the same tree as the reference's other stages (see vector_root)
DEFAULT_OUTPUT_DIR = _batched.vector_root() / "output"

_iter_input_npy_files = _batched.list_input_files
_load_input_image_u8 = _batched.load_image_u8
_case_stem_from_input = _batched.case_stem


def _run_ideal_rowwise(x_u8: np.ndarray, h: list[float]) -> np.ndarray:
    """[height, width] uint8 -> [height, width] float64, every row an
    independent ``fir_1d_ideal`` (gen_ideal_output.py:37-50) — one launch."""
    x_u8 = np.asarray(x_u8)
    if x_u8.ndim != 2:
        raise ValueError(f"expected 2D array, got shape={x_u8.shape}")
    return _ops.fir1d_ideal(x_u8, h)


def _generate(tap_label: str, coeff_map, input_dir: Path, output_dir: Path, overwrite: bool) -> int:
    return _batched.generate_outputs(
        input_dir=Path(input_dir).resolve(),
        out_dir=Path(output_dir).resolve() / f"ideal_{tap_label}",
        coeff_map=coeff_map,
        name_format="{stem}__{coeff}_ideal_" + tap_label + "_y_f64.npy",
        run_image=_run_ideal_rowwise,
        overwrite=overwrite,
    )


def generate_ideal_3tap_output_vector(input_dir: Path = DEFAULT_INPUT_DIR, output_dir: Path = DEFAULT_OUTPUT_DIR, *,
                                      overwrite: bool = False) -> int:
    return _generate("3tap", h_coeff_3tap_map, input_dir, output_dir, overwrite)


def generate_ideal_5tap_output_vector(input_dir: Path = DEFAULT_INPUT_DIR, output_dir: Path = DEFAULT_OUTPUT_DIR, *,
                                      overwrite: bool = False) -> int:
    return _generate("5tap", h_coeff_5tap_map, input_dir, output_dir, overwrite)


def _build_argparser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Generate FIR 1D ideal output vectors for 3tap/5tap filters (GPU).")
    p.add_argument("--input-dir", type=Path, default=DEFAULT_INPUT_DIR)
    p.add_argument("--output-dir", type=Path, default=DEFAULT_OUTPUT_DIR)
    p.add_argument("--tap", choices=("all", "3", "5"), default="all")
    p.add_argument("--overwrite", action="store_true")
    return p


def main(argv=None) -> int:
    args = _build_argparser().parse_args(argv)
    t0 = perf_counter()
    counts = {}
    kw = dict(input_dir=args.input_dir, output_dir=args.output_dir, overwrite=args.overwrite)
    try:
        if args.tap in ("all", "3"):
            counts["ideal_3tap"] = generate_ideal_3tap_output_vector(**kw)
        if args.tap in ("all", "5"):
            counts["ideal_5tap"] = generate_ideal_5tap_output_vector(**kw)
    except Exception as exc:
        print(f'[FAIL] gen_ideal_output file=gen_ideal_output.py generated=0 failed=1 '
              f'elapsed={perf_counter() - t0:.2f}s error="{exc}"')
        raise
    detail = " ".join(f"{k}={v}" for k, v in counts.items())
    print(f"[OK] gen_ideal_output file=gen_ideal_output.py generated={sum(counts.values())} failed=0 "
          f"elapsed={perf_counter() - t0:.2f}s out={Path(args.output_dir).resolve()} {detail}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
