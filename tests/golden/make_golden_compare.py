#!/usr/bin/env python
"""Report-level golden numbers FROM THE UNMODIFIED REFERENCE.

Runs only where /root/reference and the reference pipeline's full outputs
(--full-dir, made by make_golden.py) exist.  Writes

  tests/golden/compare_rows.json      per-case metrics of all 28 + 28 ideal/fixed
                                      file pairs, computed by the reference's own
                                      `_compute_metrics`, and the reference's own
                                      `_summarize_rows` over them
  tests/golden/image_large_input.npz  the uint8 input of the 4499x2999 case (the
                                      six smaller inputs are in images_inputs.npz)
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import numpy as np

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
from fir_1d.sim.vector import gen_3tap_compare_report as rep3  # noqa: E402
from fir_1d.sim.vector import gen_5tap_compare_report as rep5  # noqa: E402
from fir_1d.sim.vector.h_coeff import h_coeff_3tap_map, h_coeff_5tap_map  # noqa: E402

HERE = Path(__file__).resolve().parent


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--full-dir", type=Path, default=Path("/tmp/firgolden"))
    args = ap.parse_args()
    inp, out = args.full_dir / "input", args.full_dir / "output"
    stems = sorted(p.name[: -len("_x_u8.npy")] for p in inp.glob("*_x_u8.npy"))
    result = {}
    for tap, cmap, mod in (("3tap", h_coeff_3tap_map, rep3), ("5tap", h_coeff_5tap_map, rep5)):
        rows = []
        for stem in stems:
            for cname in sorted(cmap):
                yi = np.load(out / f"ideal_{tap}" / f"{stem}__{cname}_ideal_{tap}_y_f64.npy")
                yf = np.load(out / f"fixed_{tap}" / f"{stem}__{cname}_fixed_{tap}_y_u8.npy")
                m = mod._compute_metrics(yi, yf)
                rows.append({"key": f"{stem}__{cname}", "case_stem": stem, "coeff_name": cname,
                             "height": int(yi.shape[0]), "width": int(yi.shape[1]), **m})
        rows.sort(key=lambda r: (r["case_stem"], r["coeff_name"]))
        result[tap] = {"rows": rows, "summary": mod._summarize_rows(rows)}
    (HERE / "compare_rows.json").write_text(json.dumps(result, indent=1) + "\n")
    big = [s for s in stems if "4499x2999" in s][0]
    np.savez_compressed(HERE / "image_large_input.npz", **{big: np.load(inp / f"{big}_x_u8.npy")})
    for tap in result:
        s = result[tap]["summary"]
        print(tap, {k: round(s[k], 6) for k in ("avg_mae", "avg_rmse", "avg_mean_err", "max_max_abs_err")}, s["num_samples_total"])


if __name__ == "__main__":
    main()
