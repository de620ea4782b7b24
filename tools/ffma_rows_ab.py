"""A/B of the general 5x5 tensor-map kernel with 0 / 1 / 2 kernel rows accumulated by FFMA (FIRGPU_T2_FFMA_ROWS,
read by the library on every call): 16384^2 int16 -> int16, int16 -> uint8 and uint8 -> uint8, CUDA events around
back-to-back launches, every output hashed (the three settings must be bit-equal).  One JSON line per setting.
The FFMA variants are compiled only into a library built with `make EXTRA=-DFIRGPU_T2_FFMA=1` (they measured slower,
profiles/README.md R2.5c); the shipped library ignores the setting and every line then times the all-dp2a kernel.

    python tools/ffma_rows_ab.py [--reps 10] [--size 16384]
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "warmup-fir-filter_b200"))


def main():
    import torch
    import firgpu
    from firgpu import quant

    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--size", type=int, default=16384)
    args = ap.parse_args()
    peak = json.loads((Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json").read_text()).get("hbm_gbs", 6545.9)
    n = args.size
    g = torch.Generator(device="cuda").manual_seed(1)
    base = torch.randint(-64, 320, (1, n, n), generator=g, device="cuda", dtype=torch.int16)
    kr = np.random.default_rng(0).uniform(-0.2, 0.2, size=(5, 5))
    lp = np.array([1, 4, 6, 4, 1]) / 16.0
    k_norm = np.outer(lp, lp) + np.random.default_rng(1).uniform(-0.002, 0.002, size=(5, 5))   # gain about 1, not rank-1
    cases = [("i16->i16 general q4.12", base, np.int16, kr, 12, 32, 4),
             ("i16->i16 gain-1 q1.15", base, np.int16, k_norm, 15, 32, 4),
             ("i16->i16 general q4.12 acc_bits 16 (live wrap)", base, np.int16, kr, 12, 16, 4),
             ("i16->u8 general q4.12", base, np.uint8, kr, 12, 32, 3),
             ("u8->u8 general q4.12", base.clamp(0, 255).to(torch.uint8), np.uint8, kr, 12, 32, 2)]
    for tag, img, out_dtype, k, frac, acc, bytes_per_px in cases:
        hq = quant.quantize_taps(np.asarray(k), frac, 16)
        out = torch.empty(img.shape, dtype=torch.uint8 if out_dtype is np.uint8 else torch.int16, device="cuda")
        ref_hash = None
        for fr in (0, 1, 2, 1, 0):
            os.environ["FIRGPU_T2_FFMA_ROWS"] = str(fr)
            out.zero_()
            for _ in range(2):
                firgpu.fir2d_fixed_q(img, hq, frac, acc, out_dtype=out_dtype, out=out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                firgpu.fir2d_fixed_q(img, hq, frac, acc, out_dtype=out_dtype, out=out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            h = hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest()[:16]
            if ref_hash is None:
                ref_hash = h
            gpx = n * n / ms / 1e6
            print(json.dumps({"case": tag, "ffma_rows": fr, "ms": round(ms, 4), "Gpixel_per_s": round(gpx, 1),
                              "frac_of_hbm_peak": round(gpx * bytes_per_px / peak, 4), "sha": h,
                              "equal_to_all_dp2a": h == ref_hash}), flush=True)


if __name__ == "__main__":
    main()
