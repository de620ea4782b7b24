#!/usr/bin/env python
"""Host-buffer (end-to-end) throughput of firgpu_fir1d_fixed_host for the
bench workload, for a given FIRGPU_CHUNK_MB (read once at library load)."""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "warmup-fir-filter_b200")]
import firgpu  # noqa: E402

n = 1 << 28
x = torch.randint(0, 256, (n,), dtype=torch.int16).pin_memory()
y = torch.empty(n, dtype=torch.int16).pin_memory()
h = [1 / 16, 4 / 16, 6 / 16, 4 / 16, 1 / 16]
xn, yn = x.numpy(), y.numpy()
for _ in range(2):
    firgpu.fir1d_fixed(xn, h, 15, 32, 16, out_dtype=np.int16, out=yn, validate=False)
reps = 6
t0 = time.perf_counter()
for _ in range(reps):
    firgpu.fir1d_fixed(xn, h, 15, 32, 16, out_dtype=np.int16, out=yn, validate=False)
dt = (time.perf_counter() - t0) / reps
print({"chunk_mb": os.environ.get("FIRGPU_CHUNK_MB", "default"), "ms": round(dt * 1e3, 2), "gsamples_s": round(n / dt / 1e9, 2),
       "gb_s_each_way": round(2 * n / dt / 1e9, 1)})
# pageable buffers for comparison
xp, yp = np.array(xn), np.empty_like(yn)
firgpu.fir1d_fixed(xp, h, 15, 32, 16, out_dtype=np.int16, out=yp, validate=False)
t0 = time.perf_counter()
for _ in range(3):
    firgpu.fir1d_fixed(xp, h, 15, 32, 16, out_dtype=np.int16, out=yp, validate=False)
dt = (time.perf_counter() - t0) / 3
print({"pageable": True, "ms": round(dt * 1e3, 2), "gsamples_s": round(n / dt / 1e9, 2)})
assert np.array_equal(yp, yn)

# config 4 through the host entry point: one 16384 x 16384 int16 image, 3x3 and 5x5, pinned buffers
img = torch.randint(0, 256, (16384, 16384), dtype=torch.int16).pin_memory()
out = torch.empty_like(img).pin_memory()
for k, tag in (([[1 / 9] * 3] * 3, "3x3"), (np.outer(h, h).tolist(), "5x5")):
    for _ in range(2):
        firgpu.fir2d_fixed(img.numpy(), k, 12, 32, 16, out_dtype=np.int16, out=out.numpy(), validate=False)
    t0 = time.perf_counter()
    for _ in range(4):
        firgpu.fir2d_fixed(img.numpy(), k, 12, 32, 16, out_dtype=np.int16, out=out.numpy(), validate=False)
    dt = (time.perf_counter() - t0) / 4
    print({"2d": tag, "ms": round(dt * 1e3, 2), "gpixels_s": round(img.numel() / dt / 1e9, 2), "gb_s_each_way": round(2 * img.numel() / dt / 1e9, 1)})
