#!/usr/bin/env python
"""Generate the committed golden fixtures FROM THE UNMODIFIED REFERENCE.

Runs only where /root/reference exists (the build container).  It imports the
reference's own ``fir_1d_fixed_golden`` / ``fir_1d_ideal`` and its row-wise
generators — nothing from this repo's product or oracle code — and writes:

  tests/golden/fir1d_cases.npz     KATs + seeded random cases: inputs, params,
                                   reference outputs (fixed u8, ideal f64)
  tests/golden/images_small.npz    the two 64x64 BMP cases: input + all 16
                                   fixed and 16 ideal reference outputs
  tests/golden/images_inputs.npz   the u8 inputs of the 6 cases <= 1.1 Mpx
                                   (PIL-decoded here; JPEG decode cannot be
                                   repeated on the GPU box)
  tests/golden/image_crops.npz     16 rows (top 2, middle 12, bottom 2) of the 4499x2999 case + their
                                   16 fixed reference outputs
  tests/golden/manifest.json       SHA-256 of every full-size reference output
                                   (56 fixed + 56 ideal) and of the 7 inputs

Usage:  python tests/golden/make_golden.py [--full-dir /tmp/firgolden]
``--full-dir`` points at a scratch directory holding the reference pipeline's
full outputs (input/, output/{fixed,ideal}_{3,5}tap/); when it is missing the
script regenerates it with the reference's own generators (about 4 minutes on
4 processes; the reference needs ``input/`` to pre-exist).
"""
from __future__ import annotations

import argparse
import hashlib
import json
import sys
from multiprocessing import Process
from pathlib import Path

import numpy as np

sys.dont_write_bytecode = True
REFERENCE = Path("/root/reference")
sys.path.insert(0, str(REFERENCE))

from fir_1d.model.python.fir_1d_fixed_ref import fir_1d_fixed_golden  # noqa: E402
from fir_1d.model.python.fir_1d_ref import fir_1d_ideal  # noqa: E402
from fir_1d.sim.vector import gen_fixed_output as ref_gen_fixed  # noqa: E402
from fir_1d.sim.vector import gen_ideal_output as ref_gen_ideal  # noqa: E402
from fir_1d.sim.vector.gen_input_vectors import generate_input_vector_jsons  # noqa: E402
from fir_1d.sim.vector.h_coeff import h_coeff_3tap_map, h_coeff_5tap_map  # noqa: E402

HERE = Path(__file__).resolve().parent


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def regenerate_full(full: Path) -> None:
    inp, out = full / "input", full / "output"
    inp.mkdir(parents=True, exist_ok=True)
    out.mkdir(parents=True, exist_ok=True)
    generate_input_vector_jsons(output_dir=inp, overwrite=True)
    kw = dict(input_dir=inp, output_dir=out, overwrite=True)
    jobs = [ref_gen_fixed.generate_fixed_3tap_output_vector, ref_gen_fixed.generate_fixed_5tap_output_vector,
            ref_gen_ideal.generate_ideal_3tap_output_vector, ref_gen_ideal.generate_ideal_5tap_output_vector]
    procs = [Process(target=f, kwargs=kw) for f in jobs]
    [p.start() for p in procs]
    [p.join() for p in procs]
    assert all(p.exitcode == 0 for p in procs)


# ---- 1. KATs + seeded random cases through the reference functions ----------
def case_list():
    """(name, x, h, frac_bits, acc_bits, coeff_bits) — x as ndarray."""
    f8 = np.float64
    cases = [
        # the reference's own unit-test vectors (fir_1d/sim/tests/test_1d_fixed.py:9-40,95-104)
        ("kat_q412_center", np.array([10, 20, 30, 40], f8), [0.25, 0.5, 0.25], 12, 32, 16),
        ("kat_round_clamp_in", np.array([-1.2, 0.5, 1.5, 254.6, 300.2], f8), [1.0], 12, 32, 16),
        ("kat_sat_high", np.array([255, 255], f8), [7.999755859375], 12, 32, 16),
        ("kat_sat_low", np.array([255, 255], f8), [-8.0], 12, 32, 16),
        ("kat_even_taps", np.array([255, 255, 255, 255], f8), [0.5, 0.25], 12, 32, 16),
        # test_1d_ideal.py:27-42
        ("kat_ideal_prep", np.array([-1.2, 0.49, 0.5, 1.5, 254.6, 300.2], f8), [1.0], 12, 32, 16),
        ("kat_ideal_noclamp", np.array([255, 255], f8), [5.0], 12, 32, 16),
        ("kat_ideal_len4", np.array([3, 7, 11, 15, 19], f8), [0.1, 0.5, 0.3, 0.1], 12, 32, 16),
        # SURVEY.md §8(c)5 extras
        ("impulse_3", np.array([0, 0, 100, 0, 0], f8), [0.25, 0.5, 1.0], 12, 32, 16),
        ("impulse_5", np.array([0, 0, 0, 100, 0, 0, 0], f8), [0.1, 0.2, 0.3, 0.4, 0.5], 12, 32, 16),
        ("half_up", np.array([1, 3, 5], f8), [0.5], 12, 32, 16),
        ("wrap_acc16", np.array([255, 100, 7], f8), [1.0], 12, 16, 16),
        ("wrap_acc20", np.array([255, 100, 7], f8), [1.0], 12, 20, 16),
        ("wrap_acc21", np.array([255, 100, 7], f8), [1.0], 12, 21, 16),
        ("c32_acc32", np.array([255, 1, 2], f8), [7.5, -7.5, 0.001], 28, 32, 32),
        ("c32_acc64", np.array([255, 1, 2], f8), [7.5, -7.5, 0.001], 28, 64, 32),
        ("q115_simple_lp5", np.arange(10, 90, 10).astype(f8), h_coeff_5tap_map["simple_lp"], 15, 32, 16),
        ("n_lt_l_1", np.array([200], f8), [0.2] * 5, 12, 32, 16),
        ("n_lt_l_2", np.array([200, 100], f8), [0.2] * 5, 12, 32, 16),
        ("i16_container", np.array([-300, -1, 0, 128, 255, 256, 32767, -32768], np.int16), [1.0], 12, 32, 16),
        ("empty", np.array([], f8), [0.25, 0.5, 0.25], 12, 32, 16),
        ("tie_even_coeff", np.array([1, 2, 3, 255], f8), [0.5 + 2.0 ** -13], 12, 32, 16),
        ("orientation", np.array([1, 2, 4, 8, 16, 32], f8), [-1.0, 0.0, 1.0], 12, 32, 16),
        ("c8_q17", np.array([0, 255, 128, 3, 77], f8), [0.9921875, -1.0, 0.5], 7, 32, 8),
        ("neg_round", np.array([1, 1, 1, 1], f8), [-0.5, 0.0, 0.0], 12, 32, 16),
    ]
    rng = np.random.default_rng(20261018)
    for t in range(260):
        n_taps = int(rng.integers(1, 10))
        cb = int(rng.choice([8, 16, 32]))
        fb = int(rng.integers(1, min(cb + 6, 40)))
        ab = int(rng.choice([8, 12, 16, 20, 24, 31, 32, 33, 40, 48, 62, 63, 64, 80]))
        n = int(rng.integers(0, 96))
        hi = min(8.0, ((1 << (cb - 1)) - 1) / (1 << fb))
        lo = max(-8.0, -(1 << (cb - 1)) / (1 << fb))
        h = [float(v) for v in rng.uniform(lo, hi, size=n_taps)]
        kind = int(rng.integers(0, 4))
        if kind == 0:
            x = rng.integers(0, 256, size=n).astype(np.uint8)
        elif kind == 1:
            x = rng.integers(-400, 600, size=n).astype(np.int16)
        elif kind == 2:
            x = rng.uniform(-50, 320, size=n)
        else:
            x = (rng.integers(-2, 515, size=n) * 0.5).astype(np.float64)   # exact .5 ties
        cases.append((f"rand_{t:03d}", x, h, fb, ab, cb))
    # the 8 course coefficient sets on random rows, both Q formats that accept them
    for name, h in list(h_coeff_3tap_map.items()) + list(h_coeff_5tap_map.items()):
        tag = f"{len(h)}tap_{name}"
        x = rng.integers(0, 256, size=257).astype(np.uint8)
        cases.append((f"coef_q412_{tag}", x, list(h), 12, 32, 16))
        if max(h) <= 32767 / 32768 and min(h) >= -1.0:
            cases.append((f"coef_q115_{tag}", x, list(h), 15, 32, 16))
    return cases


def write_case_file() -> int:
    blob, index = {}, []
    for name, x, h, fb, ab, cb in case_list():
        y_fixed = fir_1d_fixed_golden(x.tolist(), h, frac_bits=fb, acc_bits=ab, coeff_bits=cb)
        y_ideal = np.array(fir_1d_ideal(x.tolist(), h), dtype=np.float64)
        blob[f"{name}/x"] = x
        blob[f"{name}/h"] = np.array(h, dtype=np.float64)
        blob[f"{name}/y_fixed"] = np.asarray(y_fixed, dtype=np.uint8)
        blob[f"{name}/y_ideal"] = y_ideal
        index.append({"name": name, "frac_bits": fb, "acc_bits": ab, "coeff_bits": cb})
    blob["index_json"] = np.frombuffer(json.dumps(index).encode(), dtype=np.uint8)
    np.savez_compressed(HERE / "fir1d_cases.npz", **blob)
    return len(index)


# ---- 2. image-derived fixtures from the reference pipeline's full outputs ---
def coeff_sets():
    for tap, cmap in (("3tap", h_coeff_3tap_map), ("5tap", h_coeff_5tap_map)):
        for cname in cmap:
            yield tap, cname


def write_image_files(full: Path) -> dict:
    inp, out = full / "input", full / "output"
    in_files = sorted(inp.glob("*_x_u8.npy"), key=lambda p: p.name.lower())
    manifest = {"inputs": {}, "fixed": {}, "ideal": {}}
    small, inputs, crops = {}, {}, {}
    for p in in_files:
        stem = p.name[: -len("_x_u8.npy")]
        x = np.load(p)
        manifest["inputs"][stem] = {"shape": list(x.shape), "sha256": sha(x)}
        is_small = x.size <= 64 * 64
        if x.size <= 1_200_000:
            inputs[stem] = x
        else:
            crop_rows = np.r_[0:2, x.shape[0] // 2 - 6:x.shape[0] // 2 + 6, x.shape[0] - 2:x.shape[0]]
            crops[f"{stem}/rows"] = crop_rows
            crops[f"{stem}/x"] = x[crop_rows]
        for tap, cname in coeff_sets():
            yf = np.load(out / f"fixed_{tap}" / f"{stem}__{cname}_fixed_{tap}_y_u8.npy")
            yi = np.load(out / f"ideal_{tap}" / f"{stem}__{cname}_ideal_{tap}_y_f64.npy")
            assert yf.dtype == np.uint8 and yi.dtype == np.float64 and yf.shape == x.shape == yi.shape
            manifest["fixed"][f"{stem}__{cname}_{tap}"] = sha(yf)
            manifest["ideal"][f"{stem}__{cname}_{tap}"] = sha(yi)
            if is_small:
                small[f"{stem}/{tap}/{cname}/fixed"] = yf
                small[f"{stem}/{tap}/{cname}/ideal"] = yi
            if x.size > 1_200_000:
                crops[f"{stem}/{tap}/{cname}/fixed"] = yf[crop_rows]
        if is_small:
            small[f"{stem}/x"] = x
    np.savez_compressed(HERE / "images_small.npz", **small)
    np.savez_compressed(HERE / "images_inputs.npz", **inputs)
    np.savez_compressed(HERE / "image_crops.npz", **crops)
    return manifest


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--full-dir", type=Path, default=Path("/tmp/firgolden"))
    args = ap.parse_args()
    if not (args.full_dir / "output" / "ideal_5tap").exists():
        regenerate_full(args.full_dir)
    n_cases = write_case_file()
    manifest = write_image_files(args.full_dir)
    manifest["note"] = ("SHA-256 of raw array bytes of the reference pipeline's outputs "
                        "(python pipeline_fir_1d.py equivalent, unmodified reference, Q4.12/acc32/coeff16)")
    manifest["coeff_3tap"] = {k: list(v) for k, v in h_coeff_3tap_map.items()}
    manifest["coeff_5tap"] = {k: list(v) for k, v in h_coeff_5tap_map.items()}
    (HERE / "manifest.json").write_text(json.dumps(manifest, indent=1, sort_keys=True) + "\n")
    print(f"wrote {n_cases} function-level cases, {len(manifest['fixed'])} fixed + "
          f"{len(manifest['ideal'])} ideal file hashes")


if __name__ == "__main__":
    main()
