#!/usr/bin/env python
"""Golden vectors from the reference's own C++ (fir_1d/model/cpp/fir_1d_ref.cpp),
compiled in place into oracle/_ref/libfir1d_ref.so by `make -C oracle ref`.
Runs only where /root/reference exists; writes tests/golden/cpp_fir1d_cases.npz:
raw float64 inputs, taps, Fir1D::fir_1d_idel outputs (full convolution) and
Fir1D::process_sample streams."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import c_oracle  # noqa: E402

c_oracle.build(ref=True)
assert c_oracle.ref_lib() is not None, "oracle/_ref was not built (reference tree missing?)"
rng = np.random.default_rng(20261019)
blob = {}
names = []
cases = [("pdf_p13", np.array([1.0, 2.0, 3.0, 4.0, 5.0]), np.array([1 / 3, 1 / 3, 1 / 3]))]
for t in range(60):
    n = int(rng.integers(1, 300))
    l = int(rng.integers(1, 12))
    kind = t % 3
    x = rng.uniform(-500, 500, size=n) if kind == 0 else rng.integers(0, 256, size=n).astype(np.float64) if kind == 1 \
        else rng.normal(0, 1e3, size=n)
    h = rng.uniform(-8, 8, size=l)
    cases.append((f"rand_{t:02d}", x, h))
for name, x, h in cases:
    blob[f"{name}/x"] = x
    blob[f"{name}/h"] = h
    blob[f"{name}/full"] = c_oracle.ref_fir1d_full(x, h)
    blob[f"{name}/stream"] = c_oracle.ref_fir1d_stream(x, h)
    names.append(name)
blob["names"] = np.array(names)
np.savez_compressed(Path(__file__).resolve().parent / "cpp_fir1d_cases.npz", **blob)
print(f"wrote {len(names)} C++ reference cases")
