#!/usr/bin/env python
"""A compact pass through every kernel family for compute-sanitizer
(memcheck / racecheck / initcheck / synccheck), small enough to finish under
the tool's slowdown.  Each result is also checked against the oracle.

    compute-sanitizer --tool racecheck python tests/perf/sweep_all_kernels.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path[:0] = [str(ROOT), str(ROOT / "warmup-fir-filter_b200")]
import firgpu  # noqa: E402
from firgpu import _lib, shard  # noqa: E402
from oracle import fir_oracle_np as onp  # noqa: E402

rng = np.random.default_rng(0)
LP5 = [1 / 16, 4 / 16, 6 / 16, 4 / 16, 1 / 16]
n_checked = 0


def sig(n, dtype):
    x = rng.integers(0, 256, size=n).astype(dtype)
    if dtype != np.uint8:
        x[rng.random(n) < 0.02] = -700
    return x


def check(got, want, what):
    global n_checked
    assert np.array_equal(got.cpu().numpy() if hasattr(got, "cpu") else got, want), what
    n_checked += 1


# staged 1-D: containers x taps x dilation, multi-tile, ragged tails, unaligned views, halos
for in_dt, out_dt in ((np.int16, np.int16), (np.uint8, np.uint8), (np.int16, np.uint8), (np.uint8, np.int16)):
    for L, dil, n in ((5, 1, 70_001), (3, 1, 40_000), (9, 1, 33_333), (5, 2, 66_002), (2, 1, 17)):
        h = (np.ones(L) / (2 * L)).tolist()
        x = sig(n, in_dt)
        y = firgpu.fir1d_fixed(torch.from_numpy(x).cuda(), h, 15, 32, 16, dilation=dil, out_dtype=out_dt)
        assert _lib.last_path() == _lib.PATH_STAGED
        check(y, onp.fir1d_fixed(x, h, 15, 32, 16, dilation=dil, out_dtype=out_dt), ("staged", in_dt, out_dt, L, dil, n))
base = sig(5 * 4499 + 3, np.uint8)
view = torch.from_numpy(base).cuda()[1:1 + 5 * 4499].view(5, 4499)
check(firgpu.fir1d_fixed(view, LP5), onp.fir1d_fixed(base[1:1 + 5 * 4499].reshape(5, 4499), LP5), "unaligned rows")
x = sig(300_000, np.int16)
want = onp.fir1d_fixed(x, LP5, 15, 32, 16, out_dtype=np.int16)
for p in shard.plan_segments(x.size, 3, 5, 1, align=4096):
    buf, body = shard.scatter_segment(x, p, "cuda")
    y = firgpu.fir1d_fixed(body, LP5, 15, 32, 16, out_dtype=np.int16, halo_left_valid=p.before, halo_right_valid=p.after)
    check(y, want[p.start:p.stop], "segment halos")
# multi-filter
sets = [[0.25, 0.5, 0.25], [1 / 3] * 3, [-1.0, 0, 1.0], [-0.125, 1.25, -0.125]]
xi = sig(37 * 1280, np.uint8).reshape(37, 1280)
check(firgpu.fir1d_fixed_multi(torch.from_numpy(xi).cuda(), sets), np.stack([onp.fir1d_fixed(xi, h) for h in sets]), "multi")
# generic 1-D (wrap live, long filter, float input) and ideal
xs = sig(9000, np.int16)
check(firgpu.fir1d_fixed(torch.from_numpy(xs).cuda(), [0.9, -0.7, 0.8], 12, 16, 16), onp.fir1d_fixed(xs, [0.9, -0.7, 0.8], 12, 16, 16), "generic wrap")
h33 = (np.ones(33) / 40).tolist()
check(firgpu.fir1d_fixed(torch.from_numpy(xs).cuda(), h33), onp.fir1d_fixed(xs, h33), "generic 33 taps")
xf = rng.uniform(-10, 300, size=5000)
check(firgpu.fir1d_fixed(torch.from_numpy(xf).cuda(), LP5), onp.fir1d_fixed(xf, LP5), "generic f64 in")
check(firgpu.fir1d_ideal(torch.from_numpy(xs).cuda(), LP5), onp.fir1d_ideal(xs, LP5), "ideal f64")
# 2-D: tiled general / separable, generic
for dt, shape in ((np.int16, (2, 150, 1000)), (np.uint8, (1, 90, 2016)), (np.int16, (1, 40, 333))):
    img = sig(int(np.prod(shape)), dt).reshape(shape)
    for k in ([[1 / 9] * 3] * 3, np.outer(LP5, LP5).tolist(), rng.uniform(-0.2, 0.2, size=(5, 5)).tolist(), [[v] for v in LP5]):
        check(firgpu.fir2d_fixed(torch.from_numpy(img).cuda(), k, out_dtype=dt), onp.fir2d_fixed(img, k, out_dtype=dt), ("2d", dt, shape))
img = sig(64 * 64, np.uint8).reshape(64, 64)
check(firgpu.fir2d_ideal(torch.from_numpy(img).cuda(), [[1 / 9] * 3] * 3), onp.fir2d_ideal(img, [[1 / 9] * 3] * 3), "2d ideal")
# fused compare + host pipelines
m = firgpu.compare_metrics(torch.from_numpy(xi).cuda(), LP5)
assert m["num_samples"] == xi.size
check(firgpu.fir1d_fixed(x, LP5, 15, 32, 16, out_dtype=np.int16), want, "host pipeline")
check(firgpu.fir2d_fixed(img, [[1 / 9] * 3] * 3), onp.fir2d_fixed(img, [[1 / 9] * 3] * 3), "host 2d")
torch.cuda.synchronize()
print(f"sanitize_cases ok: {n_checked} results checked, {_lib.launch_count()} firgpu launches")
