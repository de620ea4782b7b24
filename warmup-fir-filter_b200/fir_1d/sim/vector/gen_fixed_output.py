"""Drop-in for the reference's ``fir_1d.sim.vector.gen_fixed_output``: same
public functions, keyword defaults, output naming
(``{case}__{coeff}_fixed_{3|5}tap_y_u8.npy`` under ``fixed_{3|5}tap/``) and CLI
(fir_1d/sim/vector/gen_fixed_output.py:110-197), with each image filtered by
one batched GPU call instead of ``height`` Python calls."""
from __future__ import annotations

import argparse
from pathlib import Path
from time import perf_counter

import numpy as np

from firgpu import ops as _ops
from fir_1d.model.python.fir_1d_fixed_ref import fir_1d_fixed_golden  # noqa: F401  (re-exported like the reference)
from fir_1d.sim.vector import _batched
from fir_1d.sim.vector.h_coeff import h_coeff_3tap_map, h_coeff_5tap_map

THIS_FILE = Path(__file__).resolve()
DEFAULT_INPUT_DIR = _batched.vector_root() / "input"      # the same tree as the reference's other stages (see vector_root)
DEFAULT_OUTPUT_DIR = _batched.vector_root() / "output"

_iter_input_npy_files = _batched.list_input_files
_load_input_image_u8 = _batched.load_image_u8
_case_stem_from_input = _batched.case_stem


def _run_fixed_rowwise(x_u8: np.ndarray, h: list[float], *, frac_bits: int, acc_bits: int, coeff_bits: int) -> np.ndarray:
    """[height, width] uint8 -> [height, width] uint8, every row an independent
    ``fir_1d_fixed_golden`` (gen_fixed_output.py:34-60) — one launch."""
    x_u8 = np.asarray(x_u8)
    if x_u8.ndim != 2:
        raise ValueError(f"expected 2D array, got shape={x_u8.shape}")
    return _ops.fir1d_fixed(x_u8, h, frac_bits, acc_bits, coeff_bits)


def _generate(tap_label: str, coeff_map, input_dir: Path, output_dir: Path, frac_bits: int, acc_bits: int,
              coeff_bits: int, overwrite: bool) -> int:
    return _batched.generate_outputs(
        input_dir=Path(input_dir).resolve(),
        out_dir=Path(output_dir).resolve() / f"fixed_{tap_label}",
        coeff_map=coeff_map,
        name_format="{stem}__{coeff}_fixed_" + tap_label + "_y_u8.npy",
        run_image=lambda img, taps: _run_fixed_rowwise(img, taps, frac_bits=frac_bits, acc_bits=acc_bits,
                                                       coeff_bits=coeff_bits),
        overwrite=overwrite,
        run_image_multi=lambda img, tap_sets: _ops.fir1d_fixed_multi(img, tap_sets, frac_bits, acc_bits, coeff_bits),
    )


def generate_fixed_3tap_output_vector(input_dir: Path = DEFAULT_INPUT_DIR, output_dir: Path = DEFAULT_OUTPUT_DIR, *,
                                      frac_bits: int = 12, acc_bits: int = 32, coeff_bits: int = 16,
                                      overwrite: bool = False) -> int:
    return _generate("3tap", h_coeff_3tap_map, input_dir, output_dir, frac_bits, acc_bits, coeff_bits, overwrite)


def generate_fixed_5tap_output_vector(input_dir: Path = DEFAULT_INPUT_DIR, output_dir: Path = DEFAULT_OUTPUT_DIR, *,
                                      frac_bits: int = 12, acc_bits: int = 32, coeff_bits: int = 16,
                                      overwrite: bool = False) -> int:
    return _generate("5tap", h_coeff_5tap_map, input_dir, output_dir, frac_bits, acc_bits, coeff_bits, overwrite)


def _build_argparser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Generate FIR 1D fixed output vectors for 3tap/5tap filters (GPU).")
    p.add_argument("--input-dir", type=Path, default=DEFAULT_INPUT_DIR)
    p.add_argument("--output-dir", type=Path, default=DEFAULT_OUTPUT_DIR)
    p.add_argument("--tap", choices=("all", "3", "5"), default="all")
    p.add_argument("--frac-bits", type=int, default=12)
    p.add_argument("--acc-bits", type=int, default=32)
    p.add_argument("--coeff-bits", type=int, default=16)
    p.add_argument("--overwrite", action="store_true")
    return p


def main(argv=None) -> int:
    args = _build_argparser().parse_args(argv)
    t0 = perf_counter()
    counts = {}
    kw = dict(input_dir=args.input_dir, output_dir=args.output_dir, frac_bits=args.frac_bits, acc_bits=args.acc_bits,
              coeff_bits=args.coeff_bits, overwrite=args.overwrite)
    try:
        if args.tap in ("all", "3"):
            counts["fixed_3tap"] = generate_fixed_3tap_output_vector(**kw)
        if args.tap in ("all", "5"):
            counts["fixed_5tap"] = generate_fixed_5tap_output_vector(**kw)
    except Exception as exc:
        print(f'[FAIL] gen_fixed_output file=gen_fixed_output.py generated=0 failed=1 '
              f'elapsed={perf_counter() - t0:.2f}s error="{exc}"')
        raise
    detail = " ".join(f"{k}={v}" for k, v in counts.items())
    print(f"[OK] gen_fixed_output file=gen_fixed_output.py generated={sum(counts.values())} failed=0 "
          f"elapsed={perf_counter() - t0:.2f}s out={Path(args.output_dir).resolve()} {detail}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
