#!/usr/bin/env python
"""A few launches of ONE round-2 kernel case, for ncu (--set full) and launch lists:

    python tools/profile_r2.py CASE [REPS]

CASE: headline | wrap | wrap_u8 | imad | u8i16 | narrow128 | narrow4096 | image | multi | 2d5x5 | 2d5x5_u8 | 2d3x3 | 2d7x7 | 2d7x7rank1 | iq | ideal32
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "warmup-fir-filter_b200")]
import firgpu  # noqa: E402
from firgpu import quant  # noqa: E402

LP5 = [1 / 16, 4 / 16, 6 / 16, 4 / 16, 1 / 16]
HW = [0.9, -0.7, 0.8, 0.3, -0.2]
case = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
g = torch.Generator(device="cuda").manual_seed(1)


def i16(shape):
    return torch.randint(-64, 320, shape, generator=g, device="cuda", dtype=torch.int16)


if case in ("headline", "wrap", "imad", "iq"):
    x = i16((1 << 28,))
    y = torch.empty_like(x)
    if case == "headline":
        fn = lambda: firgpu.fir1d_fixed(x, LP5, 15, 32, 16, out_dtype=np.int16, out=y, validate=False)
    elif case == "wrap":
        fn = lambda: firgpu.fir1d_fixed(x, HW, 12, 16, 16, out_dtype=np.int16, out=y, validate=False)
    elif case == "imad":
        fn = lambda: firgpu.fir1d_fixed(x, HW, 28, 32, 32, out_dtype=np.int16, out=y, validate=False)
    else:
        fn = lambda: firgpu.fir1d_fixed(x, LP5, 15, 32, 16, dilation=2, out_dtype=np.int16, out=y, validate=False)
elif case == "wrap_u8":
    x = i16((1 << 28,)).clamp(0, 255).to(torch.uint8)
    y = torch.empty_like(x)
    fn = lambda: firgpu.fir1d_fixed(x, HW, 12, 16, 16, out=y, validate=False)
elif case == "u8i16":
    x = i16((1 << 28,)).clamp(0, 255).to(torch.uint8)
    y = torch.empty(1 << 28, dtype=torch.int16, device="cuda")
    fn = lambda: firgpu.fir1d_fixed(x, LP5, 12, 32, 16, out_dtype=np.int16, out=y, validate=False)
elif case.startswith("narrow"):
    cols = int(case[6:])
    x = i16(((1 << 27) // cols, cols))
    y = torch.empty_like(x)
    fn = lambda: firgpu.fir1d_fixed(x, LP5, 15, 32, 16, out_dtype=np.int16, out=y, validate=False)
elif case in ("image", "multi"):
    x = i16((2999, 4499)).clamp(0, 255).to(torch.uint8)
    if case == "image":
        y = torch.empty_like(x)
        fn = lambda: firgpu.fir1d_fixed(x, LP5, 12, 32, 16, out=y, validate=False)
    else:
        from fir_1d.sim.vector.h_coeff import h_coeff_5tap_map
        sets = list(h_coeff_5tap_map.values())
        fn = lambda: firgpu.fir1d_fixed_multi(x, sets, validate=False)
elif case.startswith("2d"):
    u8 = case.endswith("_u8")
    img = i16((1, 16384, 16384))
    if u8:
        img = img.clamp(0, 255).to(torch.uint8)
    out = torch.empty_like(img)
    g7 = np.array([1, 6, 15, 20, 15, 6, 1]) / 64.0
    k = [[1 / 9] * 3] * 3 if "3x3" in case else np.outer(g7, g7).tolist() if "7x7rank1" in case else np.random.default_rng(7).uniform(-0.1, 0.1, size=(7, 7)).tolist() if "7x7" in case else (np.outer(LP5, LP5).tolist() if "rank1" in case else np.random.default_rng(0).uniform(-0.2, 0.2, size=(5, 5)).tolist())
    hq = quant.quantize_taps(np.asarray(k), 12, 16)
    fn = lambda: firgpu.fir2d_fixed_q(img, hq, 12, 32, out_dtype=np.uint8 if u8 else np.int16, out=out)
elif case == "ideal32":
    x = i16((1 << 26,)).clamp(0, 255).to(torch.uint8)
    y = torch.empty(1 << 26, dtype=torch.float32, device="cuda")
    fn = lambda: firgpu.fir1d_ideal(x, LP5, out_dtype=np.float32, out=y, validate=False)
else:
    raise SystemExit(f"unknown case {case}")
for _ in range(reps):
    fn()
torch.cuda.synchronize()
print("done", case)
