"""GPU parity, 1-D fixed-point path, through the C ABI (ctypes) and the
reference-shaped Python entry points.  Bit-exact against the golden vectors
made by the unmodified reference and against the pinned oracle."""
from __future__ import annotations

import hashlib
import re

import numpy as np
import pytest

from oracle import c_oracle, fir_oracle_np as onp

pytestmark = pytest.mark.gpu

SIMPLE_LP5 = [1 / 16, 4 / 16, 6 / 16, 4 / 16, 1 / 16]


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available()
    return torch


@pytest.fixture(scope="module")
def fg():
    import firgpu
    from firgpu import _lib
    assert _lib.lib().firgpu_version() >= 100
    return firgpu


def golden_fn():
    from fir_1d.model.python.fir_1d_fixed_ref import fir_1d_fixed_golden
    return fir_1d_fixed_golden


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def signal(rng, n, dtype=np.int16, wild=0.01):
    """99 % pixels in [0,255], 1 % anywhere in the container's range
    (exercises the input clamp) — SURVEY §8(d) input recipe."""
    x = rng.integers(0, 256, size=n).astype(dtype)
    if dtype != np.uint8 and wild > 0:
        info = np.iinfo(dtype)
        mask = rng.random(n) < wild
        x[mask] = rng.integers(info.min, info.max, size=int(mask.sum()), endpoint=True).astype(dtype)
    return x


# ---- the reference's own unit tests, restated (fir_1d/sim/tests/test_1d_fixed.py) ----
def test_reference_kats(fg):
    f = golden_fn()
    y = f([10, 20, 30, 40], [0.25, 0.5, 0.25])                                   # :9-16
    assert isinstance(y, np.ndarray) and y.dtype == np.uint8 and y.tolist() == [10, 20, 30, 28]
    assert f([-1.2, 0.5, 1.5, 254.6, 300.2], [1.0]).tolist() == [0, 1, 2, 255, 255]   # :19-24
    assert f([255, 255], [7.999755859375]).tolist() == [255, 255]                # :27-33
    assert f([255, 255], [-8.0]).tolist() == [0, 0]                              # :35-40
    y = f([255, 255, 255, 255], [0.5, 0.25])                                     # :95-104
    assert y.shape == (4,) and y.dtype == np.uint8 and y.min() >= 0 and y.max() <= 255


@pytest.mark.parametrize("kwargs, pattern", [
    (dict(x=[1.0, float("nan")], h=[1.0]), r"x\[1\].*finite"),
    (dict(x=[1.0], h=[float("inf")]), r"h\[0\].*finite"),
    (dict(x=[1.0], h=[]), r"must not be empty"),
    (dict(x=[1.0], h=[1.0], coeff_bits=12), r"Invalid coeff_bits=12"),
    (dict(x=[1.0], h=[1.0], frac_bits=0), r"frac_bits must be > 0"),
    (dict(x=[1.0], h=[1.0], acc_bits=0), r"acc_bits must be > 0"),
    (dict(x=[1.0], h=[8.0]), r"out of Q-format real range"),
    (dict(x=[1.0], h=[1.0], frac_bits=7, coeff_bits=8), r"out of Q-format real range"),
    (dict(x=[1.0], h=[8.0 + 1e-6]), r"\|h\| must be <="),
])
def test_reference_error_contract(fg, kwargs, pattern):
    with pytest.raises(ValueError) as err:
        golden_fn()(**kwargs)
    assert re.search(pattern, str(err.value))


def test_every_golden_case_bit_exact(fg, golden_cases):
    f = golden_fn()
    for c in golden_cases:
        y = f(c["x"].tolist(), c["h"], frac_bits=c["frac_bits"], acc_bits=c["acc_bits"], coeff_bits=c["coeff_bits"])
        assert y.dtype == np.uint8 and np.array_equal(y, c["y_fixed"]), c["name"]
        y2 = f(c["x"], c["h"], frac_bits=c["frac_bits"], acc_bits=c["acc_bits"], coeff_bits=c["coeff_bits"])
        assert np.array_equal(y2, c["y_fixed"]), c["name"] + " (ndarray input)"


def test_reference_images_by_hash(fg, images_inputs, manifest):
    """Config 1: the reference's own images, all 8 coefficient sets, compared
    with the SHA-256 of the reference pipeline's output files."""
    from fir_1d.sim.vector.gen_fixed_output import _run_fixed_rowwise
    from firgpu import _lib
    n = 0
    for stem in images_inputs.keys():
        x = images_inputs[stem]
        for tap in ("3tap", "5tap"):
            for name, h in manifest[f"coeff_{tap}"].items():
                y = _run_fixed_rowwise(x, h, frac_bits=12, acc_bits=32, coeff_bits=16)
                assert y.shape == x.shape and y.dtype == np.uint8
                assert sha(y) == manifest["fixed"][f"{stem}__{name}_{tap}"], (stem, name, tap)
                assert _lib.last_path() == _lib.PATH_STAGED
                n += 1
    assert n == 48


def test_odd_width_image_rows(fg, image_crops, manifest):
    stem = "case_005_img_006_4499x2999_gray"
    x = image_crops[f"{stem}/x"]
    for tap in ("3tap", "5tap"):
        for name, h in manifest[f"coeff_{tap}"].items():
            assert np.array_equal(fg.fir1d_fixed(x, h), image_crops[f"{stem}/{tap}/{name}/fixed"])


# ---- the staged kernel against the oracle -------------------------------------------
@pytest.mark.parametrize("in_dtype", [np.uint8, np.int16])
@pytest.mark.parametrize("out_dtype", [np.uint8, np.int16])
@pytest.mark.parametrize("n_taps", [1, 2, 3, 4, 5, 6, 7, 8, 9])
def test_staged_all_tap_counts(fg, torch, in_dtype, out_dtype, n_taps):
    from firgpu import _lib
    rng = np.random.default_rng(1000 + n_taps)
    for frac, dil in ((12, 1), (15, 1), (15, 2), (12, 2)):
        h = rng.uniform(-0.45, 0.45, size=n_taps).tolist()
        for n in (1, 2, 7, 64, 1000, 16384, 16385, 40000 + 2 * n_taps):
            if dil == 2 and n % 2:
                n += 1
            x = signal(rng, n, in_dtype)
            xt = torch.from_numpy(x).cuda()
            y = fg.fir1d_fixed(xt, h, frac, 32, 16, dilation=dil, out_dtype=out_dtype)
            assert _lib.last_path() == _lib.PATH_STAGED
            want = c_oracle.fir1d_fixed(x, onp.quantize_taps(h, frac, 16), frac, 32, dilation=dil, out_dtype=out_dtype)
            got = y.cpu().numpy()
            assert got.dtype == np.dtype(out_dtype)
            assert np.array_equal(got, want), (n_taps, frac, dil, n)


def test_generic_kernel_serves_what_staged_declines(fg, torch):
    """A live acc_bits wrap and 32-bit taps are served by the staged kernel (WRAP post mode, IMAD for taps beyond
    int16); filters beyond the largest bucket, accumulators wider than 32 bits whose sums do not fit int32, and small
    float / int32 inputs go to the generic kernels."""
    from firgpu import _lib
    rng = np.random.default_rng(9)
    x = signal(rng, 5000, np.int16)
    xt = torch.from_numpy(x).cuda()
    S, G = _lib.PATH_STAGED, _lib.PATH_GENERIC
    cases = [
        (12, 16, 16, [0.9, -0.7, 0.8, 0.3, -0.2], S),     # acc_bits wrap is live
        (28, 32, 32, [7.5, -7.5, 0.001], S),              # 32-bit taps, wrap live
        (12, 32, 16, rng.uniform(-0.1, 0.1, size=130).tolist(), G),  # longer than the largest staged bucket (129)
        (12, 20, 16, [1.0], S),                           # SURVEY 8(c)5 wrap case
        (3, 12, 8, [7.0, -6.5, 3.25, 1.0], S),            # tiny frac_bits, 8-bit taps, wrap live
        (30, 64, 32, [0.4, 0.3], G),                      # 64-bit accumulator, sums beyond int32
        (28, 40, 32, [0.001, 0.002], S),                  # wider accumulator, but the sums provably fit int32
    ]
    for frac, acc, cb, h, path in cases:
        y = fg.fir1d_fixed(xt, h, frac, acc, cb).cpu().numpy()
        assert _lib.last_path() == path, (frac, acc, cb)
        assert np.array_equal(y, onp.fir1d_fixed(x, h, frac, acc, cb)), (frac, acc, cb)
    for dt in (np.int32, np.float32, np.float64):
        xs = (rng.uniform(-20, 300, size=3001)).astype(dt)
        y = fg.fir1d_fixed(torch.from_numpy(xs).cuda(), SIMPLE_LP5).cpu().numpy()
        assert _lib.last_path() == _lib.PATH_GENERIC
        assert np.array_equal(y, onp.fir1d_fixed(xs, SIMPLE_LP5))


def test_batched_rows_pitched_and_unaligned(fg, torch):
    """Rows x cols with pitch > cols, odd pitches (no 16-byte alignment), and
    base pointers offset by one element: the patch path of the staged kernel."""
    from firgpu import _lib
    rng = np.random.default_rng(21)
    for dtype in (np.uint8, np.int16):
        for rows, cols, pitch in ((5, 4499, 4499), (3, 641, 650), (7, 33000, 33008), (2, 16384, 16400), (64, 64, 64)):
            base = signal(rng, rows * pitch + 3, dtype)
            for off in (0, 1, 3):
                buf = torch.from_numpy(base).cuda()
                view = buf[off:off + rows * pitch].view(rows, pitch)[:, :cols]
                y = fg.fir1d_fixed(view, SIMPLE_LP5, 15, 32, 16)
                assert _lib.last_path() == _lib.PATH_STAGED
                xs = base[off:off + rows * pitch].reshape(rows, pitch)[:, :cols]
                assert np.array_equal(y.cpu().numpy(), onp.fir1d_fixed(xs, SIMPLE_LP5, 15, 32, 16)), (dtype, rows, cols, pitch, off)


def test_segment_halos_make_seamless_shards(fg, torch):
    """A long signal cut into segments with duplicated halos gives exactly the
    unsharded result (config 2 / 5 sharding), for both dilations."""
    from firgpu import shard
    rng = np.random.default_rng(33)
    for dil, n in ((1, 200_000), (2, 131_072 + 4096), (1, 70_001)):
        x = signal(rng, n, np.int16)
        want = c_oracle.fir1d_fixed(x, onp.quantize_taps(SIMPLE_LP5, 15, 16), 15, 32, dilation=dil, out_dtype=np.int16, threads=4)
        for parts in (2, 3, 8):
            pieces = shard.plan_segments(n, parts, 5, dil, align=4096)
            got = np.empty(n, dtype=np.int16)
            for p in pieces:
                if p.size == 0:
                    continue
                buf, body = shard.scatter_segment(x, p, "cuda")
                y = fg.fir1d_fixed(body, SIMPLE_LP5, 15, 32, 16, dilation=dil, out_dtype=np.int16,
                                   halo_left_valid=p.before, halo_right_valid=p.after)
                got[p.start:p.stop] = y.cpu().numpy()
            assert np.array_equal(got, want), (dil, n, parts)


def test_host_entry_point_pipeline(fg):
    """firgpu_fir1d_fixed_host: pageable and pinned buffers, many rows, and a
    row long enough to be cut into column segments with halos."""
    rng = np.random.default_rng(44)
    for shape in ((1, 40_000_000), (300, 70_001), (4096, 4096)):
        x = signal(rng, shape[0] * shape[1], np.int16).reshape(shape)
        want = c_oracle.fir1d_fixed(x, onp.quantize_taps(SIMPLE_LP5, 15, 16), 15, 32, out_dtype=np.int16, threads=8)
        y = fg.fir1d_fixed(x, SIMPLE_LP5, 15, 32, 16, out_dtype=np.int16)
        assert np.array_equal(y, want), shape
        pin_x, pin_y = fg.PinnedArray(shape, np.int16), fg.PinnedArray(shape, np.int16)
        pin_x.array[...] = x
        fg.fir1d_fixed(pin_x.array, SIMPLE_LP5, 15, 32, 16, out_dtype=np.int16, out=pin_y.array)
        assert np.array_equal(pin_y.array, want), shape
        pin_x.close()
        pin_y.close()


def test_iq_is_two_independent_filters(fg, torch):
    """Config 5 semantics (lib/mycomplex.cpp:12-22): interleaved I/Q with real
    taps == the 1-D model on z[0::2] and z[1::2]."""
    rng = np.random.default_rng(55)
    z = signal(rng, 2 * 300_001, np.int16)
    zt = torch.from_numpy(z).cuda()
    y = fg.fir1d_fixed_iq(zt, SIMPLE_LP5, 15, 32, 16, out_dtype=np.int16).cpu().numpy()
    assert np.array_equal(y[0::2], onp.fir1d_fixed(z[0::2], SIMPLE_LP5, 15, 32, 16, out_dtype=np.int16))
    assert np.array_equal(y[1::2], onp.fir1d_fixed(z[1::2], SIMPLE_LP5, 15, 32, 16, out_dtype=np.int16))
    y2 = fg.fir1d_fixed_iq(zt.view(-1, 2), SIMPLE_LP5, 15, 32, 16, out_dtype=np.int16)
    assert y2.shape == (300_001, 2) and np.array_equal(y2.cpu().numpy().reshape(-1), y)


def test_full_size_stream_properties(fg, torch):
    """Config 2 at full size (2**28 int16 samples): windows against the
    oracle (start, a tile seam, middle, end) + size-independent properties."""
    n = 1 << 28
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = torch.randint(0, 256, (n,), generator=g, device="cuda", dtype=torch.int16)
    wild = torch.rand(n, generator=g, device="cuda") < 0.01
    x[wild] = torch.randint(-32768, 32768, (int(wild.sum()),), generator=g, device="cuda", dtype=torch.int32).to(torch.int16)
    del wild
    y = fg.fir1d_fixed(x, SIMPLE_LP5, 15, 32, 16, out_dtype=np.int16)
    hq = onp.quantize_taps(SIMPLE_LP5, 15, 16)
    for start in (0, 8192 - 64, (n // 2) - 5000, n - (1 << 20)):
        stop = min(n, start + (1 << 20))
        lo, hi = max(0, start - 2), min(n, stop + 2)
        xs = x[lo:hi].cpu().numpy()
        want = c_oracle.fir1d_fixed(xs, hq, 15, 32, out_dtype=np.int16, threads=8)
        # interior of the window is independent of the window's own zero padding
        a, b = (0 if lo == 0 else 2), (xs.size if hi == n else xs.size - 2)
        assert np.array_equal(y[lo + a:lo + b].cpu().numpy(), want[a:b]), start
    # ... and every one of the 2**28 outputs against the C oracle on all host cores
    import os
    full = c_oracle.fir1d_fixed(x.cpu().numpy(), hq, 15, 32, out_dtype=np.int16, threads=max(1, len(os.sched_getaffinity(0))))
    assert np.array_equal(y.cpu().numpy(), full)
    del full
    # range and DC-gain properties of a unity-gain low-pass over clamped pixels
    assert int(y.min()) >= 0 and int(y.max()) <= 255
    xc = x.clamp(0, 255)
    # round-half-up on sixteenths biases the mean by +1/32
    assert abs(float(y[2:-2].double().mean()) - float(xc.double().mean())) < 0.05
    # shift invariance: filtering a shifted copy shifts the output (away from the ends)
    y_shift = fg.fir1d_fixed(x[4096:], SIMPLE_LP5, 15, 32, 16, out_dtype=np.int16)
    assert torch.equal(y_shift[2:1 << 22], y[4096 + 2:4096 + (1 << 22)])
    # u8 container of the same pixels gives the same values
    y8 = fg.fir1d_fixed(xc.to(torch.uint8), SIMPLE_LP5, 15, 32, 16)
    assert torch.equal(y8.to(torch.int16), y)


def test_imad_datapath_matches_dp2a(fg, torch):
    """The alternative IMAD datapath (FIRGPU_MAC=1, kept for A/B evidence) is
    bit-identical."""
    import subprocess, sys, os, textwrap
    code = textwrap.dedent("""
        import sys, numpy as np, torch
        sys.path[:0] = [%r, %r]
        import firgpu
        from oracle import fir_oracle_np as onp
        rng = np.random.default_rng(3)
        x = rng.integers(-400, 700, size=100003).astype(np.int16)
        h = [1/16, 4/16, 6/16, 4/16, 1/16]
        for frac in (12, 15):
            y = firgpu.fir1d_fixed(torch.from_numpy(x).cuda(), h, frac, 32, 16).cpu().numpy()
            assert np.array_equal(y, onp.fir1d_fixed(x, h, frac, 32, 16))
        print("ok")
    """) % (str(__import__("pathlib").Path(__file__).resolve().parents[1]),
            str(__import__("pathlib").Path(__file__).resolve().parents[1] / "warmup-fir-filter_b200"))
    env = dict(os.environ, FIRGPU_MAC="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def test_multi_filter_reads_input_once(fg, torch, manifest):
    """firgpu_fir1d_fixed_multi[_host]: the generators' coefficient loop as one
    call.  Four tap sets on uint8 go through the NF=4 staged kernel; any other
    count / container falls back to single passes — same values either way."""
    from firgpu import _lib
    rng = np.random.default_rng(66)
    for tap in ("3tap", "5tap"):
        sets = list(manifest[f"coeff_{tap}"].values())
        for shape in ((37, 4499), (64, 64), (5, 70_000), (1, 300_001)):
            x = rng.integers(0, 256, size=shape).astype(np.uint8)
            want = np.stack([onp.fir1d_fixed(x, h) for h in sets])
            before = _lib.launch_count()
            got = fg.fir1d_fixed_multi(torch.from_numpy(x).cuda(), sets)
            # one pass over the input, row ends included
            assert _lib.launch_count() - before == 1 and _lib.last_path() == _lib.PATH_STAGED
            assert got.shape == (4,) + shape and np.array_equal(got.cpu().numpy(), want), (tap, shape)
            assert np.array_equal(fg.fir1d_fixed_multi(x, sets), want), (tap, shape, "host")
    x = rng.integers(-100, 400, size=(3, 10_001)).astype(np.int16)
    for sets in ([[0.25, 0.5, 0.25]] * 3, [[0.1, 0.2, 0.3, 0.2, 0.1], [0.2] * 5, [0.5, 0, 0, 0, 0.5], [1 / 16, 4 / 16, 6 / 16, 4 / 16, 1 / 16], [0, 0, 1, 0, 0]]):
        want = np.stack([onp.fir1d_fixed(x, h) for h in sets])
        assert np.array_equal(fg.fir1d_fixed_multi(torch.from_numpy(x).cuda(), sets).cpu().numpy(), want)
        assert np.array_equal(fg.fir1d_fixed_multi(x, sets), want)
    # mixed post modes inside one group (one tap set cannot be pre-scaled to a 16-bit shift)
    sets = [[0.25, 0.5, 0.25], [-0.125, 1.25, -0.125], [7.5, -7.0, 0.25], [1 / 3] * 3]
    xu = rng.integers(0, 256, size=(9, 5000)).astype(np.uint8)
    want = np.stack([onp.fir1d_fixed(xu, h) for h in sets])
    assert np.array_equal(fg.fir1d_fixed_multi(torch.from_numpy(xu).cuda(), sets).cpu().numpy(), want)
    with pytest.raises(ValueError, match="same length"):
        fg.fir1d_fixed_multi(xu, [[1.0], [0.5, 0.5]])


def test_wide_signed_int16_fast_path(fg, torch):
    """Extended entry point on int16 containers with non-reference clamp ranges:
    served by the wide datapath of the staged kernel (IMAD on full-width
    samples, saturating pack) whenever int32 provably suffices — bit-exact
    against the parametrised oracle, tile edges, ragged tails, shard halos and
    interleaved I/Q included."""
    from firgpu import _lib, shard
    rng = np.random.default_rng(77)
    full = (-32768, 32767)
    for n_taps in (1, 2, 3, 5, 8, 9):
        h = rng.uniform(-0.2, 0.2, size=n_taps).tolist()            # sum|h| < 2: int32 is provably enough in Q1.15
        for dil, n in ((1, 1), (1, 9), (1, 16384), (1, 70_003), (2, 66_002), (1, 300_001)):
            x = rng.integers(-32768, 32768, size=n).astype(np.int16)
            xt = torch.from_numpy(x).cuda()
            for in_r, out_r in ((full, full), ((-1000, 2000), (-300, 300)), ((0, 255), full)):
                got = fg.fir1d_fixed(xt, h, 15, 32, 16, dilation=dil, in_range=in_r, out_range=out_r, out_dtype=np.int16)
                assert _lib.last_path() == _lib.PATH_STAGED, (n_taps, dil, n, in_r, out_r)
                want = onp.fir1d_fixed(x, h, 15, 32, 16, dilation=dil, in_range=in_r, out_range=out_r, out_dtype=np.int16)
                assert np.array_equal(got.cpu().numpy(), want), (n_taps, dil, n, in_r, out_r)
    # rows with pitch, and segments with halos
    x = rng.integers(-32768, 32768, size=(5, 33_000)).astype(np.int16)
    h = [0.1, -0.2, 0.4, -0.2, 0.1]
    view = torch.from_numpy(x).cuda()[:, :32_999]
    got = fg.fir1d_fixed(view, h, 15, 32, 16, in_range=full, out_range=full)
    assert _lib.last_path() == _lib.PATH_STAGED
    assert np.array_equal(got.cpu().numpy(), onp.fir1d_fixed(x[:, :32_999], h, 15, 32, 16, in_range=full, out_range=full, out_dtype=np.int16))
    sig = rng.integers(-32768, 32768, size=200_000).astype(np.int16)
    want = onp.fir1d_fixed(sig, h, 15, 32, 16, in_range=full, out_range=full, out_dtype=np.int16)
    for p in shard.plan_segments(sig.size, 3, 5, 1, align=4096):
        buf, body = shard.scatter_segment(sig, p, "cuda")
        y = fg.fir1d_fixed(body, h, 15, 32, 16, in_range=full, out_range=full, halo_left_valid=p.before, halo_right_valid=p.after)
        assert np.array_equal(y.cpu().numpy(), want[p.start:p.stop])
    # saturation really happens, and the int32 proof sends big taps to the generic kernel
    loud = np.full(4096, 32767, np.int16)
    y = fg.fir1d_fixed(torch.from_numpy(loud).cuda(), [0.9, 0.9], 15, 32, 16, in_range=full, out_range=full)
    assert _lib.last_path() == _lib.PATH_STAGED and int(y[10]) == 32767
    y = fg.fir1d_fixed(torch.from_numpy(loud).cuda(), [0.999, 0.999, 0.999], 15, 32, 16, in_range=full, out_range=full)
    assert _lib.last_path() == _lib.PATH_GENERIC            # 32767 * 3 * 32735 > 2^31: wrap is live, int64 kernel
    assert np.array_equal(y.cpu().numpy(), onp.fir1d_fixed(loud, [0.999, 0.999, 0.999], 15, 32, 16, in_range=full, out_range=full, out_dtype=np.int16))


def test_long_filters_run_in_staged_buckets(fg, torch):
    """10..129 taps: served by the staged kernel through zero-padded tap-count
    buckets (13/17/25/33/49/65/97/129) that keep the centre tap in place — bit-exact,
    even and odd lengths, both containers, tile edges and shard halos."""
    from firgpu import _lib, shard
    rng = np.random.default_rng(88)
    for n_taps in (10, 11, 13, 14, 16, 17, 21, 32, 33, 40, 48, 49, 63, 64, 65, 66, 96, 97, 98, 128, 129):
        h = (rng.uniform(-1, 1, size=n_taps) / n_taps).tolist()
        for dtype in (np.uint8, np.int16):
            for frac in (15, 12):
                for n in (1, 50, 4095, 16384, 70_001):
                    x = signal(rng, n, dtype)
                    y = fg.fir1d_fixed(torch.from_numpy(x).cuda(), h, frac, 32, 16, out_dtype=dtype)
                    # (the 97 / 129-tap buckets exist for int16 samples; uint8 samples beyond 65 taps: generic kernel)
                    want_path = _lib.PATH_GENERIC if dtype is np.uint8 and n_taps > 65 else _lib.PATH_STAGED
                    assert _lib.last_path() == want_path, (n_taps, dtype, frac, n)
                    assert np.array_equal(y.cpu().numpy(), onp.fir1d_fixed(x, h, frac, 32, 16, out_dtype=dtype)), (n_taps, dtype, frac, n)
    h = (rng.uniform(-1, 1, size=31) / 31).tolist()
    x = signal(rng, 5 * 9000, np.int16).reshape(5, 9000)
    assert np.array_equal(fg.fir1d_fixed(torch.from_numpy(x).cuda(), h, 15, 32, 16).cpu().numpy(), onp.fir1d_fixed(x, h, 15, 32, 16))
    sig = signal(rng, 150_000, np.int16)
    want = onp.fir1d_fixed(sig, h, 15, 32, 16, out_dtype=np.int16)
    for p in shard.plan_segments(sig.size, 3, 31, 1, align=4096):
        buf, body = shard.scatter_segment(sig, p, "cuda")
        y = fg.fir1d_fixed(body, h, 15, 32, 16, out_dtype=np.int16, halo_left_valid=p.before, halo_right_valid=p.after)
        assert _lib.last_path() == _lib.PATH_STAGED
        assert np.array_equal(y.cpu().numpy(), want[p.start:p.stop])
    host = fg.fir1d_fixed(sig, h, 15, 32, 16, out_dtype=np.int16)
    assert np.array_equal(host, want)


def test_float_and_int32_containers_use_the_fast_kernels(fg, torch):
    """int32 / float32 / float64 samples: one streaming pass rounds (half up, in
    float64) and clamps them to uint8 pixels, then the staged / tiled kernels
    filter — bit-exact, exact .5 ties, shard halos, 1-D and 2-D."""
    from firgpu import _lib, shard
    rng = np.random.default_rng(99)
    n = 70_003
    halves = (rng.integers(-6, 520, size=n) * 0.5)                 # every other value is an exact .5 tie
    for x in (rng.uniform(-30, 300, size=n), halves, halves.astype(np.float32), rng.integers(-500, 700, size=n).astype(np.int32),
              rng.uniform(-5, 260, size=(6, 9001)).astype(np.float32)):
        for h, frac in ((SIMPLE_LP5, 15), ([0.25, 0.5, 0.25], 12), ((np.ones(21) / 42).tolist(), 15)):
            for out_dtype in (np.uint8, np.int16):
                y = fg.fir1d_fixed(torch.from_numpy(x).cuda(), h, frac, 32, 16, out_dtype=out_dtype)
                assert _lib.last_path() == _lib.PATH_STAGED, (x.dtype, len(h), out_dtype)
                assert np.array_equal(y.cpu().numpy(), onp.fir1d_fixed(x, h, frac, 32, 16, out_dtype=out_dtype)), (x.dtype, len(h), out_dtype)
    x = rng.uniform(-30, 300, size=150_000)
    want = onp.fir1d_fixed(x, SIMPLE_LP5, 15, 32, 16)
    for p in shard.plan_segments(x.size, 3, 5, 1, align=4096):
        buf, body = shard.scatter_segment(x, p, "cuda")
        y = fg.fir1d_fixed(body, SIMPLE_LP5, 15, 32, 16, halo_left_valid=p.before, halo_right_valid=p.after)
        assert np.array_equal(y.cpu().numpy(), want[p.start:p.stop])
    # wrap-live taps on a float container: uint8 copy + generic kernel
    y = fg.fir1d_fixed(torch.from_numpy(x).cuda(), [0.9, -0.7, 0.8], 12, 16, 16)
    assert np.array_equal(y.cpu().numpy(), onp.fir1d_fixed(x, [0.9, -0.7, 0.8], 12, 16, 16))
    img = rng.uniform(-20, 280, size=(2, 150, 1000))
    for k in ([[1 / 9] * 3] * 3, np.outer(SIMPLE_LP5, SIMPLE_LP5).tolist()):
        y = fg.fir2d_fixed(torch.from_numpy(img).cuda(), k)
        assert _lib.last_path() == _lib.PATH_STAGED
        assert np.array_equal(y.cpu().numpy(), onp.fir2d_fixed(img, k))
    assert np.array_equal(fg.fir2d_fixed(img, k), onp.fir2d_fixed(img, k))           # host entry point
    assert np.array_equal(fg.fir1d_fixed(x, SIMPLE_LP5), onp.fir1d_fixed(x, SIMPLE_LP5))
    # multi-filter pass (one conversion for all tap sets) and the floating-point model
    from fir_1d.sim.vector.h_coeff import h_coeff_5tap_map
    sets = list(h_coeff_5tap_map.values())
    rows = rng.uniform(-20, 280, size=(33, 1501)).astype(np.float32)
    want = np.stack([onp.fir1d_fixed(rows, h) for h in sets])
    assert np.array_equal(fg.fir1d_fixed_multi(torch.from_numpy(rows).cuda(), sets).cpu().numpy(), want)
    assert _lib.last_path() == _lib.PATH_STAGED
    assert np.array_equal(fg.fir1d_fixed_multi(rows, sets), want)
    for src in (x, rows):
        got = fg.fir1d_ideal(torch.from_numpy(src).cuda(), SIMPLE_LP5)
        assert _lib.last_path() == _lib.PATH_STAGED
        assert np.array_equal(got.cpu().numpy(), onp.fir1d_ideal(src, SIMPLE_LP5))
        assert np.array_equal(np.asarray(fg.fir1d_ideal(src, SIMPLE_LP5)), onp.fir1d_ideal(src, SIMPLE_LP5))


def test_sample_rounding_matches_float64_floor_at_every_boundary(fg, torch):
    """The device rounds float samples from their bit pattern; the reference
    computes floor(x + 0.5) with a float64 add (fir_1d_ref.py:35-41).  Identity
    filter over every x.5 tie, its float neighbours, the float64 add's one
    inexact tie (0.5 - 2^-54 -> 1), signed zeros, denormals and huge values —
    through the generic kernel (short input), the conversion pass + staged
    kernel (long input) and the floating-point model."""
    ints = np.arange(-3, 260, dtype=np.float64)
    rng = np.random.default_rng(4)
    special = np.concatenate([
        ints, ints + 0.5, np.nextafter(ints + 0.5, -np.inf), np.nextafter(ints + 0.5, np.inf), np.nextafter(ints, -np.inf),
        np.nextafter(ints, np.inf),
        [0.0, -0.0, 5e-324, -5e-324, 1e-300, 1e300, -1e300, 0.49999999999999994, 0.5, 0.25, 2.0 ** -30, 255.49999999999997, 1.7e308],
        np.exp(rng.uniform(-40, 40, size=500)), -np.exp(rng.uniform(-40, 40, size=100))])
    assert onp.fir1d_fixed(np.array([0.49999999999999994]), [1.0])[0] == 1          # the reference's own behaviour
    assert onp.fir1d_fixed(np.array([0.49999997], np.float32), [1.0])[0] == 1       # ... and on a float32 array (float32 add)
    assert onp.fir1d_fixed([0.49999997017860413], [1.0])[0] == 0                    # ... the same value in a list: float64 add
    i32 = np.arange(-3, 260, dtype=np.float32)
    special32 = np.concatenate([
        i32, i32 + np.float32(0.5), np.nextafter(i32 + np.float32(0.5), np.float32(-np.inf)), np.nextafter(i32 + np.float32(0.5), np.float32(np.inf)),
        np.nextafter(i32, np.float32(-np.inf)), np.float32([0.0, -0.0, 1e-45, -1e-45, 1e-30, 3e38, -3e38, 0.49999997, 255.49998]),
        np.exp(rng.uniform(-40, 40, size=500)).astype(np.float32)]).astype(np.float32)
    for base in (special, special32):
        for reps in (1, 8):                                   # 1: generic kernel; 8: > 4096 samples, conversion pass
            x = np.tile(base, reps)
            # the add is evaluated in the container's precision (np.float32 + 0.5 stays float32: fir_1d_ref.py:37)
            with np.errstate(over="ignore"):
                want = np.clip(np.floor(x + x.dtype.type(0.5)).astype(np.float64), 0, 255).astype(np.uint8)
            for src in (x, torch.from_numpy(x).cuda()):
                y = fg.fir1d_fixed(src, [1.0])
                y = y.cpu().numpy() if hasattr(y, "cpu") else y
                assert np.array_equal(y, want), (x.dtype, reps, type(src))
            assert np.array_equal(np.asarray(fg.fir1d_ideal(x, [1.0]), dtype=np.float64), want.astype(np.float64))
            img = x[: (x.size // 64) * 64].reshape(-1, 64)
            assert np.array_equal(fg.fir2d_fixed(img, [[0, 0, 0], [0, 1.0, 0], [0, 0, 0]]), want[: img.size].reshape(img.shape))


def test_dense_rows_of_odd_pitch_take_the_flattened_pass(fg, torch):
    """Row batches whose pitch is not a multiple of 16 bytes (odd image widths,
    the reference's 4499-wide image) are filtered as one long row by the
    staged kernel, which zero-pads every row on its own inside the kernel:
    bit-exact with per-row zero padding, for every container pair, I/Q
    dilation, long filters and the four-tap-set pass."""
    from firgpu import _lib
    from fir_1d.sim.vector.h_coeff import h_coeff_3tap_map, h_coeff_5tap_map
    rng = np.random.default_rng(77)
    for shape in ((7, 1501), (33, 4499), (100, 255), (3, 1502)):
        for in_dtype in (np.uint8, np.int16):
            x = rng.integers(0, 256, size=shape).astype(in_dtype)
            if in_dtype == np.int16:
                x[rng.random(shape) < 0.03] = -4000
                x[rng.random(shape) < 0.03] = 9000
            xt = torch.from_numpy(x).cuda()
            for out_dtype in (np.uint8, np.int16):
                for h, frac in ((SIMPLE_LP5, 15), ([-1.0, 0.0, 1.0], 12), ([0.2] * 9, 12)):
                    got = fg.fir1d_fixed(xt, h, frac, 32, 16, out_dtype=out_dtype)
                    assert _lib.last_path() == _lib.PATH_STAGED, (shape, in_dtype, out_dtype, len(h))
                    assert np.array_equal(got.cpu().numpy(), onp.fir1d_fixed(x, h, frac, 32, 16, out_dtype=out_dtype)), (shape, in_dtype, out_dtype, len(h))
            if shape[1] % 2 == 0:                                        # interleaved I/Q rows
                got = fg.fir1d_fixed(xt, SIMPLE_LP5, 15, 32, 16, dilation=2, out_dtype=in_dtype)
                assert _lib.last_path() == _lib.PATH_STAGED
                assert np.array_equal(got.cpu().numpy(), onp.fir1d_fixed(x, SIMPLE_LP5, 15, 32, 16, dilation=2, out_dtype=in_dtype))
            if in_dtype == np.uint8:
                for sets in (list(h_coeff_3tap_map.values()), list(h_coeff_5tap_map.values())):
                    got = fg.fir1d_fixed_multi(xt, sets)
                    assert _lib.last_path() == _lib.PATH_STAGED
                    assert np.array_equal(got.cpu().numpy(), np.stack([onp.fir1d_fixed(x, h) for h in sets])), shape
    x = rng.integers(0, 256, size=(5, 3001)).astype(np.uint8)            # long filter: 21 taps -> 25-tap bucket
    h21 = (np.ones(21) / 42).tolist()
    assert np.array_equal(fg.fir1d_fixed(torch.from_numpy(x).cuda(), h21, 15, 32, 16).cpu().numpy(), onp.fir1d_fixed(x, h21, 15, 32, 16))
    assert _lib.last_path() == _lib.PATH_STAGED


def test_long_filters_int16_in_uint8_out(fg, torch):
    """10..129 taps on an int16 container with the drop-in's default uint8 output:
    served by the long-filter buckets of the staged kernel, bit-exact."""
    from firgpu import _lib
    rng = np.random.default_rng(123)
    x = signal(rng, 300_001, np.int16)
    xt = torch.from_numpy(x).cuda()
    for n_taps in (10, 13, 21, 33, 50, 65, 80, 129):
        h = (rng.uniform(-1, 1, size=n_taps) / n_taps * 4).tolist()
        got = fg.fir1d_fixed(xt, h, 12, 32, 16)
        assert _lib.last_path() == _lib.PATH_STAGED, n_taps
        assert got.dtype == torch.uint8 and np.array_equal(got.cpu().numpy(), onp.fir1d_fixed(x, h, 12, 32, 16)), n_taps
    rows = signal(rng, (5, 40_000), np.int16)
    h = (np.ones(17) / 17).tolist()
    assert np.array_equal(fg.fir1d_fixed(torch.from_numpy(rows).cuda(), h).cpu().numpy(), onp.fir1d_fixed(rows, h))


def test_long_filters_uint8_in_int16_out(fg, torch):
    """10..65 taps on uint8 samples with the values delivered in an int16 container (32-byte stores per 16-byte input
    chunk): the last container pair to get long-filter buckets — bit-exact, every bucket, odd lengths and offsets,
    a live accumulator wrap, rows, and a column view whose stores are not 32-byte aligned."""
    from firgpu import _lib
    rng = np.random.default_rng(321)
    x = signal(rng, 300_001, np.uint8)
    xt = torch.from_numpy(x).cuda()
    for n_taps in (10, 13, 16, 21, 33, 50, 64, 65, 90):
        h = (rng.uniform(-1, 1, size=n_taps) / n_taps * 4).tolist()
        for frac, acc in ((12, 32), (15, 32), (12, 16)):
            got = fg.fir1d_fixed(xt, h, frac, acc, 16, out_dtype=np.int16)
            assert _lib.last_path() == (_lib.PATH_STAGED if n_taps <= 65 else _lib.PATH_GENERIC), (n_taps, frac, acc)
            assert got.dtype == torch.int16
            assert np.array_equal(got.cpu().numpy(), onp.fir1d_fixed(x, h, frac, acc, 16, out_dtype=np.int16)), (n_taps, frac, acc)
    rows = signal(rng, (6, 40_016), np.uint8)
    rt = torch.from_numpy(rows).cuda()
    h = (rng.uniform(-1, 1, size=25) / 8).tolist()
    for c0, width in ((0, 40_016), (16, 40_000), (3, 39_999)):
        view = rt[:, c0:c0 + width]
        got = fg.fir1d_fixed(view, h, 12, 32, 16, out_dtype=np.int16)
        assert np.array_equal(got.cpu().numpy(), onp.fir1d_fixed(rows[:, c0:c0 + width], h, 12, 32, 16, out_dtype=np.int16)), (c0, width)
