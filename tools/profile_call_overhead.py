#!/usr/bin/env python
"""Where the per-call host time of the device entry points goes (small input, so the kernel is negligible).
python tools/profile_call_overhead.py"""
import cProfile, pstats, sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "warmup-fir-filter_b200")]
import firgpu
from firgpu import quant

LP5 = [1 / 16, 4 / 16, 6 / 16, 4 / 16, 1 / 16]
x = torch.randint(0, 256, (64, 64), device="cuda", dtype=torch.uint8)
y = torch.empty_like(x)
hq = quant.quantize_taps(LP5, 12, 16)
cases = {
    "fir1d_fixed(x, h, out=y, validate=False)": lambda: firgpu.fir1d_fixed(x, LP5, 12, 32, 16, out=y, validate=False),
    "fir1d_fixed(x, h) (allocates, validates)": lambda: firgpu.fir1d_fixed(x, LP5, 12, 32, 16),
    "fir1d_fixed_q(x, hq, out=y)": lambda: firgpu.fir1d_fixed_q(x, hq, 12, 32, out=y),
}
for name, fn in cases.items():
    for _ in range(200):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5000):
        fn()
    dt = time.perf_counter() - t0
    torch.cuda.synchronize()
    print(f"{name}: {dt / 5000 * 1e6:.2f} us per call (host submit)")
pr = cProfile.Profile()
pr.enable()
for _ in range(5000):
    cases["fir1d_fixed(x, h, out=y, validate=False)"]()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
