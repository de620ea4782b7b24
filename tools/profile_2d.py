#!/usr/bin/env python
"""One launch of each 2-D kernel family on the config-4 shape (16384^2 int16),
for ncu:  ncu --set full -k regex:fir2d_tiled ... python tools/profile_2d.py"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "warmup-fir-filter_b200")]
import firgpu  # noqa: E402

LP5 = [1 / 16, 4 / 16, 6 / 16, 4 / 16, 1 / 16]
g = torch.Generator(device="cuda").manual_seed(4321)
x = torch.randint(0, 256, (1, 16384, 16384), generator=g, device="cuda", dtype=torch.int16)
y = torch.empty_like(x)
rng = np.random.default_rng(0)
for name, k in (("3x3 box", [[1 / 9] * 3] * 3), ("5x5 rank-1", np.outer(LP5, LP5).tolist()),
                ("5x5 general", rng.uniform(-0.2, 0.2, size=(5, 5)).tolist())):
    firgpu.fir2d_fixed(x, k, 12, 32, 16, out_dtype=np.int16, out=y, validate=False)
    torch.cuda.synchronize()
    print(name, "done")
