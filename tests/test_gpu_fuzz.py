"""Seeded random parity sweep over the dispatch space of the C ABI: every
container, tap count (short, long-bucket, generic), Q format, accumulator
width (incl. live wrap), dilation, alignment, column views at odd offsets, the
multi-filter pass, both 2-D models and the floating-point model — each case
small enough for the oracle to finish instantly, bit-exact unless stated.

The directed tests pin each kernel family; this sweep is there to catch the
seams BETWEEN families (conversion pass -> staged, flattened pass + row ends,
re-pitched 2-D copies, zero-padded off-centre taps, short-row routing ...)."""
from __future__ import annotations

import os

import numpy as np
import pytest

from oracle import fir_oracle_np as onp

pytestmark = pytest.mark.gpu

SEED = int(os.environ.get("FIRGPU_FUZZ_SEED", "0"))          # another seed = another sweep
SCALE = int(os.environ.get("FIRGPU_FUZZ_SCALE", "3"))        # multiplies the number of cases

DTYPES = (np.uint8, np.int16, np.int32, np.float32, np.float64)
EDGE_COLS = (1, 2, 3, 5, 8, 15, 16, 17, 31, 33, 63, 64, 65, 127, 255, 256, 257, 1023, 1025, 2047, 2049, 4095, 4097, 4499, 8191, 8193)


def make(rng, shape, dtype):
    if dtype in (np.float32, np.float64):
        x = rng.uniform(-40, 300, size=shape)
        ties = rng.random(shape) < 0.2
        x[ties] = np.floor(x[ties]) + 0.5                    # exact .5 ties
        return x.astype(dtype)
    x = rng.integers(0, 256, size=shape)
    if dtype != np.uint8:
        wild = rng.random(shape) < 0.05
        x[wild] = rng.integers(-3000, 3000, size=int(wild.sum()))
    return x.astype(dtype)


def pick_shape(rng, max_rows=40, max_cols=9000):
    rows = int(rng.choice((1, 1, 2, 3, 7, 16, 17, int(rng.integers(1, max_rows + 1)))))
    cols = int(rng.choice(EDGE_COLS)) if rng.random() < 0.6 else int(rng.integers(1, max_cols + 1))
    return rows, cols


def random_taps(rng, n_taps, bound):
    h = rng.uniform(-bound, bound, size=n_taps)
    h[rng.random(n_taps) < 0.15] = 0.0
    return h.tolist()


def test_fuzz_fir1d_fixed():
    import torch
    import firgpu
    rng = np.random.default_rng(20261018 + SEED)
    for case in range(260 * SCALE):
        rows, cols = pick_shape(rng)
        dtype = DTYPES[int(rng.integers(len(DTYPES)))]
        out_dtype = (np.uint8, np.int16)[int(rng.integers(2))]
        n_taps = int(rng.choice((1, 2, 3, 4, 5, 5, 5, 6, 7, 8, 9, 9, 11, 13, 17, 21, 33, 40, 70)))
        frac = int(rng.choice((8, 12, 12, 15)))
        acc = int(rng.choice((32, 32, 32, 48, 64, 24, 20)))
        bound = {8: 7.9, 12: 7.9, 15: 0.99}[frac] / max(1.0, n_taps / 4) if rng.random() < 0.8 else {8: 7.9, 12: 7.9, 15: 0.99}[frac]
        h = random_taps(rng, n_taps, bound)
        dilation = 2 if (rng.random() < 0.15 and n_taps <= 9) else 1
        kw = {}
        if rng.random() < 0.15 and dilation == 1:
            kw["centre"] = int(rng.integers(0, n_taps))
        shape = (cols,) if (rows == 1 and rng.random() < 0.5) else (rows, cols)
        x = make(rng, shape, dtype)
        want = onp.fir1d_fixed(x, h, frac, acc, 16, dilation=dilation, out_dtype=out_dtype, **kw)
        xt = torch.from_numpy(x).cuda()
        got = firgpu.fir1d_fixed(xt, h, frac, acc, 16, dilation=dilation, out_dtype=out_dtype, **kw)
        tag = (case, shape, np.dtype(dtype).name, np.dtype(out_dtype).name, n_taps, frac, acc, dilation, kw)
        assert np.array_equal(got.cpu().numpy(), want), tag
        if case % 4 == 0:                                     # host entry point (chunked pipeline)
            assert np.array_equal(firgpu.fir1d_fixed(x, h, frac, acc, 16, dilation=dilation, out_dtype=out_dtype, **kw), want), tag + ("host",)
        if len(shape) == 2 and cols > 40 and case % 3 == 0 and dilation == 1:      # column view: odd base, odd pitch
            c0 = int(rng.integers(1, 20))
            w = cols - c0 - int(rng.integers(0, 10))
            view = xt[:, c0:c0 + w]
            got = firgpu.fir1d_fixed(view, h, frac, acc, 16, out_dtype=out_dtype, **kw)
            assert np.array_equal(got.cpu().numpy(), onp.fir1d_fixed(x[:, c0:c0 + w], h, frac, acc, 16, out_dtype=out_dtype, **kw)), tag + ("view", c0, w)


def test_fuzz_multi_filter_and_ideal():
    import torch
    import firgpu
    rng = np.random.default_rng(77001 + SEED)
    for case in range(80 * SCALE):
        rows, cols = pick_shape(rng)
        dtype = DTYPES[int(rng.integers(len(DTYPES)))]
        n_taps = int(rng.choice((1, 2, 3, 5, 5, 7, 9, 12)))
        x = make(rng, (rows, cols), dtype)
        xt = torch.from_numpy(x).cuda()
        sets = [random_taps(rng, n_taps, 1.5) for _ in range(int(rng.choice((1, 3, 4, 4, 5, 8))))]
        want = np.stack([onp.fir1d_fixed(x, h) for h in sets])
        assert np.array_equal(firgpu.fir1d_fixed_multi(xt, sets).cpu().numpy(), want), (case, rows, cols, np.dtype(dtype).name, n_taps, len(sets))
        h = sets[0]
        want = onp.fir1d_ideal(x, h)
        assert np.array_equal(firgpu.fir1d_ideal(xt, h).cpu().numpy(), want), (case, "ideal", rows, cols, np.dtype(dtype).name, n_taps)
        got32 = firgpu.fir1d_ideal(xt, h, out_dtype=np.float32).cpu().numpy().astype(np.float64)
        assert np.abs(got32 - want).max() <= 1e-6 * 255 * max(np.abs(h).sum(), 1e-30) + 1e-30, (case, "ideal32")
        m = firgpu.compare_metrics(xt, h)
        fixed = onp.fir1d_fixed(x, h).astype(np.float64)
        d = fixed - want
        assert m["max_abs_err"] == np.abs(d).max() and abs(m["mae"] - np.abs(d).mean()) <= 1e-9 * max(1.0, np.abs(d).mean()), (case, "compare")


def test_fuzz_fir2d():
    import torch
    import firgpu
    rng = np.random.default_rng(424242 + SEED)
    for case in range(90 * SCALE):
        n_img = int(rng.choice((1, 1, 2, 3)))
        hgt = int(rng.choice((1, 2, 5, 16, 33, int(rng.integers(1, 120)))))
        wid = int(rng.choice((1, 2, 7, 16, 31, 64, 255, 256, 333, 1000, 1001, 1102, int(rng.integers(1, 1500)))))
        dtype = DTYPES[int(rng.integers(len(DTYPES)))]
        kh, kw = int(rng.choice((1, 3, 3, 5, 5, 7))), int(rng.choice((1, 3, 3, 5, 5, 7)))
        k = rng.uniform(-0.5, 0.5, size=(kh, kw))
        if rng.random() < 0.3:                                # rank-1 kernels take the separable path
            k = np.outer(rng.uniform(-0.7, 0.7, size=kh), rng.uniform(-0.7, 0.7, size=kw))
        k = k.tolist()
        acc = int(rng.choice((32, 32, 24, 64)))
        x = make(rng, (n_img, hgt, wid), dtype)
        xt = torch.from_numpy(x).cuda()
        out_dtype = np.int16 if (dtype == np.int16 and rng.random() < 0.7) else np.uint8
        tag = (case, (n_img, hgt, wid), np.dtype(dtype).name, kh, kw, acc, np.dtype(out_dtype).name)
        want = onp.fir2d_fixed(x, k, 12, acc, 16, out_dtype=out_dtype)
        assert np.array_equal(firgpu.fir2d_fixed(xt, k, 12, acc, 16, out_dtype=out_dtype).cpu().numpy(), want), tag
        if case % 5 == 0:
            assert np.array_equal(firgpu.fir2d_fixed(x, k, 12, acc, 16, out_dtype=out_dtype), want), tag + ("host",)
        assert np.array_equal(firgpu.fir2d_ideal(xt, k).cpu().numpy(), onp.fir2d_ideal(x, k)), tag + ("ideal",)


def test_fuzz_extended_modes_and_shards():
    """The extended entry points (alignment, clamp ranges, raw samples) and
    shard seams (segments with real halo samples, row bands with halo rows) at
    random: OUR specification restated by the oracle, equal to the reference
    model at the defaults."""
    import torch
    import firgpu
    from firgpu import shard
    rng = np.random.default_rng(9090 + SEED)
    ranges = ((0, 255), (-32768, 32767), (-100, 100), (0, 1023), (-7, 300))
    for case in range(60 * SCALE):
        rows, cols = pick_shape(rng, max_rows=12, max_cols=6000)
        n_taps = int(rng.choice((1, 2, 3, 5, 5, 7, 9, 12, 20)))
        h = random_taps(rng, n_taps, 1.2)
        x = rng.integers(-32768, 32768, size=(rows, cols)).astype(np.int16)
        xt = torch.from_numpy(x).cuda()
        kw = {"centre": int(rng.integers(0, n_taps)) if rng.random() < 0.6 else None,
              "in_range": ranges[int(rng.integers(len(ranges)))], "out_range": ranges[int(rng.integers(len(ranges)))]}
        frac, acc = int(rng.choice((12, 15))), int(rng.choice((32, 48, 24)))
        if frac == 15:
            h = [v * 0.8 for v in h]
        tag = (case, rows, cols, n_taps, frac, acc, kw)
        want = onp.fir1d_fixed(x, h, frac, acc, 16, out_dtype=np.int16, **kw)
        assert np.array_equal(firgpu.fir1d_fixed(xt, h, frac, acc, 16, out_dtype=np.int16, **kw).cpu().numpy(), want), tag
        raw = bool(rng.random() < 0.5)
        xf = rng.normal(0, 500, size=(rows, cols)) if raw else x
        ikw = {"centre": kw["centre"], "raw": raw} if raw else {"centre": kw["centre"], "in_range": kw["in_range"]}
        want = onp.fir1d_ideal(xf, h, **ikw)
        assert np.array_equal(firgpu.fir1d_ideal(torch.from_numpy(xf).cuda(), h, **ikw).cpu().numpy(), want), tag + ("ideal", raw)
    for case in range(12 * SCALE):                           # shard seams
        n = int(rng.integers(5_000, 120_000))
        n_taps = int(rng.choice((3, 5, 9, 17)))
        dil = int(rng.choice((1, 1, 2)))
        if dil == 2:
            n -= n % 2
            n_taps = min(n_taps, 9)
        h = random_taps(rng, n_taps, 0.9)
        dtype = (np.uint8, np.int16, np.float32)[int(rng.integers(3))]
        x = make(rng, (n,), dtype)
        want = onp.fir1d_fixed(x, h, 12, 32, 16, dilation=dil)
        parts = int(rng.integers(2, 6))
        for p in shard.plan_segments(n, parts, n_taps, dil, align=int(rng.choice((16, 64, 4096)))):
            buf, body = shard.scatter_segment(x, p, "cuda")
            y = firgpu.fir1d_fixed(body, h, 12, 32, 16, dilation=dil, halo_left_valid=p.before, halo_right_valid=p.after)
            assert np.array_equal(y.cpu().numpy(), want[p.start:p.stop]), (case, n, n_taps, dil, np.dtype(dtype).name, parts, p)
        hgt, wid = int(rng.integers(20, 200)), int(rng.choice((64, 333, 1000, 1102)))
        kh = int(rng.choice((3, 5)))
        k = rng.uniform(-0.4, 0.4, size=(kh, kh)).tolist()
        img = make(rng, (2, hgt, wid), (np.uint8, np.int16)[int(rng.integers(2))])
        want = onp.fir2d_fixed(img, k)
        got = np.empty_like(want)
        for band in shard.plan_bands(hgt, int(rng.integers(2, 5)), kh):
            if band.size == 0:
                continue
            buf, body = shard.scatter_band(img, band, "cuda")
            y = firgpu.fir2d_fixed(body, k, halo_top_valid=band.before, halo_bottom_valid=band.after)
            got[:, band.start:band.stop, :] = y.cpu().numpy()
        assert np.array_equal(got, want), (case, "bands", hgt, wid, kh)
