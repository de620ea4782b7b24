#!/usr/bin/env python
"""One launch each of the secondary fast kernels, for ncu:
ideal fast (u8 -> f64), row-streaming tensor-map kernel (short rows), NF=4 multi-filter, u8 stream."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "warmup-fir-filter_b200")]
import firgpu  # noqa: E402

LP5 = [1 / 16, 4 / 16, 6 / 16, 4 / 16, 1 / 16]
g = torch.Generator(device="cuda").manual_seed(1)
x8 = torch.randint(0, 256, (1 << 26,), generator=g, device="cuda", dtype=torch.uint8)
firgpu.fir1d_ideal(x8, LP5, validate=False)
xs = torch.randint(0, 256, (65536, 4096), generator=g, device="cuda", dtype=torch.int16)
firgpu.fir1d_fixed(xs, LP5, 15, 32, 16, out_dtype=np.int16, validate=False)
img = x8[: 8192 * 8192].view(8192, 8192)
firgpu.fir1d_fixed_multi(img, [LP5, [0.2] * 5, [-1 / 8, -2 / 8, 0, 2 / 8, 1 / 8], [-1 / 16, -4 / 16, 26 / 16, -4 / 16, -1 / 16]], validate=False)
firgpu.fir1d_fixed(x8, LP5, validate=False)
torch.cuda.synchronize()
print("done")
