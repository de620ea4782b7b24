"""Fused fixed-vs-ideal compare rows straight from the input vectors.

The reference's report stage (fir_1d/sim/vector/gen_{3,5}tap_compare_report.py)
loads a saved float64 ideal file and a saved uint8 fixed file per case and
reduces them with NumPy (`_compute_metrics`, :67-112).  Here each case is one
GPU pass over the *input* image that evaluates both models per pixel and
returns only the reductions (``firgpu_fir1d_compare``), so the 8-byte-per-pixel
ideal array is never written to HBM or disk.  Rows carry the same keys as the
reference's report rows (:316-335), and ``summarize`` mirrors `_summarize_rows`
(:115-155), so the published accuracy tables can be regenerated from
``*_x_u8.npy`` inputs alone.
"""
from __future__ import annotations

from pathlib import Path
from typing import Any

import numpy as np

from firgpu import ops as _ops
from fir_1d.sim.vector import _batched
from fir_1d.sim.vector.h_coeff import h_coeff_3tap_map, h_coeff_5tap_map

METRIC_KEYS = ("max_abs_err", "mae", "rmse", "mean_err", "sat_low_ratio", "sat_high_ratio", "sat_ratio", "clip_needed_ratio")


def fused_compare_rows(input_dir: Path, tap: str, *, frac_bits: int = 12, acc_bits: int = 32, coeff_bits: int = 16) -> list[dict[str, Any]]:
    """One row per (input image, coefficient set) of the ``tap`` ("3" or "5")
    bank, sorted by (case_stem, coeff_name) like the reference's shared keys."""
    coeff_map = {"3": h_coeff_3tap_map, "5": h_coeff_5tap_map}[str(tap)]
    files = _batched.list_input_files(Path(input_dir))
    if not files:
        raise FileNotFoundError(f"No input .npy files found in {input_dir}")
    rows = []
    for path in files:
        image = _batched.load_image_u8(path)
        stem = _batched.case_stem(path)
        for coeff_name, taps in coeff_map.items():
            m = _ops.compare_metrics(image, taps, frac_bits, acc_bits, coeff_bits)
            rows.append({"key": f"{stem}__{coeff_name}", "case_stem": stem, "coeff_name": coeff_name,
                         "height": int(image.shape[0]), "width": int(image.shape[1]), **m})
    rows.sort(key=lambda r: (r["case_stem"], r["coeff_name"]))
    return rows


def summarize(rows: list[dict[str, Any]]) -> dict[str, Any]:
    if not rows:
        return {"num_cases": 0, "num_samples_total": 0}
    out: dict[str, Any] = {"num_cases": len(rows), "num_samples_total": int(sum(r["num_samples"] for r in rows))}
    for key in METRIC_KEYS:
        out[f"avg_{key}"] = float(np.mean([r[key] for r in rows]))
    for key in ("max_abs_err", "mae", "rmse", "sat_ratio"):
        out[f"max_{key}"] = float(np.max([r[key] for r in rows]))
    return out
