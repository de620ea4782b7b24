#!/usr/bin/env python
"""End-to-end throughput of ONE process driving 1, 2, 4, 8 GPUs through firgpu_fir1d_fixed_host_multi: a single pinned
host array in, a single pinned host array out (the bench workload, 5-tap Q1.15 int16).

    python tools/bench_e2e_multi.py [log2_samples]
"""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "warmup-fir-filter_b200")]
import firgpu  # noqa: E402

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 30)
h = [1 / 16, 4 / 16, 6 / 16, 4 / 16, 1 / 16]
x = torch.randint(0, 256, (n,), dtype=torch.int16).pin_memory()
y = torch.empty(n, dtype=torch.int16).pin_memory()
xn, yn = x.numpy(), y.numpy()
ref = None
for g in (1, 2, 4, 8):
    if g > torch.cuda.device_count():
        break
    devs = list(range(g))
    for _ in range(2):
        firgpu.fir1d_fixed_multi_gpu(xn, h, 15, 32, 16, devices=devs, out_dtype=np.int16, out=yn, validate=False)
    reps = 4
    t0 = time.perf_counter()
    for _ in range(reps):
        firgpu.fir1d_fixed_multi_gpu(xn, h, 15, 32, 16, devices=devs, out_dtype=np.int16, out=yn, validate=False)
    dt = (time.perf_counter() - t0) / reps
    if ref is None:
        ref = yn.copy()
    print(json.dumps({"gpus": g, "samples": n, "ms": round(dt * 1e3, 2), "gsamples_s": round(n / dt / 1e9, 2),
                      "gb_s_each_way": round(2 * n / dt / 1e9, 1), "equal_to_1gpu": bool(np.array_equal(ref, yn))}), flush=True)
