#!/usr/bin/env python
"""Per-call latency of the drop-in model functions at the reference's own call
granularity (one image row per call, fir_1d/sim/vector/gen_fixed_output.py:44-59),
next to the scalar Python port of the reference's loop on the same host.

    python tests/perf/dropin_latency.py [--profile]
"""
import argparse
import cProfile
import json
import pstats
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path[:0] = [str(ROOT), str(ROOT / "warmup-fir-filter_b200")]
from fir_1d.model.python.fir_1d_fixed_ref import fir_1d_fixed_golden  # noqa: E402
from fir_1d.model.python.fir_1d_ref import fir_1d_ideal  # noqa: E402
from oracle import fir_oracle_py as opy  # noqa: E402  (checker / CPU comparison only)

H5 = [1 / 16, 4 / 16, 6 / 16, 4 / 16, 1 / 16]


def per_call(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--profile", action="store_true")
    args = ap.parse_args()
    rng = np.random.default_rng(0)
    for width in (64, 1280, 4499):
        row_u8 = rng.integers(0, 256, size=width).astype(np.uint8)
        row_list = row_u8.astype(float).tolist()
        want = fir_1d_fixed_golden(row_u8, H5)
        assert np.array_equal(want, np.asarray(opy.fir1d_fixed(row_list, H5), dtype=np.uint8))
        out = {"width": width,
               "fixed_uint8_array_us": per_call(lambda: fir_1d_fixed_golden(row_u8, H5), 200) * 1e6,
               "fixed_float_list_us": per_call(lambda: fir_1d_fixed_golden(row_list, H5), 100) * 1e6,
               "ideal_uint8_array_us": per_call(lambda: fir_1d_ideal(row_u8, H5), 100) * 1e6,
               "python_port_fixed_us": per_call(lambda: opy.fir1d_fixed(row_list, H5), 3) * 1e6}
        print(json.dumps({k: (round(v, 1) if isinstance(v, float) else v) for k, v in out.items()}), flush=True)
    if args.profile:
        row_u8 = rng.integers(0, 256, size=4499).astype(np.uint8)
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(300):
            fir_1d_fixed_golden(row_u8, H5)
        pr.disable()
        pstats.Stats(pr).sort_stats("cumulative").print_stats(18)


if __name__ == "__main__":
    main()
