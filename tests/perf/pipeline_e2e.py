#!/usr/bin/env python
"""End-to-end run of the drop-in pipeline driver's model stages (ideal + fixed
output generation, 3 and 5 taps) on the reference's 7 input images, timed, with
every one of the 112 output files checked against the SHA-256 of the reference
pipeline's own output (tests/golden/manifest.json).  The reference needs 448 s
for the same work (SURVEY.md section 6, 8-core Xeon, single thread).

    python tests/perf/pipeline_e2e.py [--workdir /tmp/firgpu_pipeline]
"""
from __future__ import annotations

import argparse
import hashlib
import json
import shutil
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path[:0] = [str(ROOT), str(ROOT / "warmup-fir-filter_b200")]
from fir_1d.sim.vector import gen_fixed_output as gf, gen_ideal_output as gi  # noqa: E402
from firgpu import _lib  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--workdir", type=Path, default=Path("/tmp/firgpu_pipeline"))
    args = ap.parse_args()
    golden = ROOT / "tests" / "golden"
    manifest = json.loads((golden / "manifest.json").read_text())
    inp, out = args.workdir / "input", args.workdir / "output"
    shutil.rmtree(args.workdir, ignore_errors=True)
    inp.mkdir(parents=True)
    n_px = 0
    for blob in (np.load(golden / "images_inputs.npz"), np.load(golden / "image_large_input.npz")):
        for stem in blob.keys():
            np.save(inp / f"{stem}_x_u8.npy", blob[stem])
            n_px += blob[stem].size
    launches0 = _lib.launch_count()
    t0 = time.perf_counter()
    counts = {
        "ideal_3tap": gi.generate_ideal_3tap_output_vector(input_dir=inp, output_dir=out, overwrite=True),
        "ideal_5tap": gi.generate_ideal_5tap_output_vector(input_dir=inp, output_dir=out, overwrite=True),
        "fixed_3tap": gf.generate_fixed_3tap_output_vector(input_dir=inp, output_dir=out, overwrite=True),
        "fixed_5tap": gf.generate_fixed_5tap_output_vector(input_dir=inp, output_dir=out, overwrite=True),
    }
    elapsed = time.perf_counter() - t0
    launches = _lib.launch_count() - launches0
    bad = 0
    for kind, suffix in (("fixed", "u8"), ("ideal", "f64")):
        for key, want in manifest[kind].items():
            stem_coeff, tap = key.rsplit("_", 1)
            path = out / f"{kind}_{tap}" / f"{stem_coeff}_{kind}_{tap}_y_{suffix}.npy"
            got = hashlib.sha256(np.ascontiguousarray(np.load(path)).tobytes()).hexdigest()
            bad += got != want
    print(json.dumps({"stage": "ideal + fixed output generation, 3 and 5 taps, 7 images", "files": sum(counts.values()),
                      "pixels_x_filters": n_px * 16, "seconds": round(elapsed, 2), "firgpu_kernel_launches": int(launches),
                      "files_not_matching_reference_sha256": bad, "reference_seconds_same_stages": 448.3}))
    assert bad == 0 and sum(counts.values()) == 112


if __name__ == "__main__":
    main()
