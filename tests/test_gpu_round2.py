"""GPU parity for what round 2 added to the hot path: the live ``acc_bits`` wrap
and 32-bit taps on the staged kernels, dense row batches in one launch (row
ends inside the kernel), BASELINE configs 3 / 4 / 5 at full size, streaming
with the history on the device, recorded launch sequences, the multi-device
host entry, and the argument checks the advisor asked for.  Everything goes
through the C ABI (ctypes); the oracle is the checker only."""
from __future__ import annotations

import os

import numpy as np
import pytest

from oracle import c_oracle, fir_oracle_np as onp

pytestmark = pytest.mark.gpu

SIMPLE_LP5 = [1 / 16, 4 / 16, 6 / 16, 4 / 16, 1 / 16]
CORES = max(1, len(os.sched_getaffinity(0)))


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available()
    return torch


@pytest.fixture(scope="module")
def fg():
    import firgpu
    from firgpu import _lib
    assert _lib.lib().firgpu_version() >= 200
    return firgpu


def signal(rng, shape, dtype=np.int16, wild=0.01):
    """SURVEY 8(d) input recipe: 99 % pixels in [0,255], 1 % anywhere in the container."""
    x = rng.integers(0, 256, size=shape).astype(dtype)
    if dtype != np.uint8 and wild > 0:
        info = np.iinfo(dtype)
        mask = rng.random(shape) < wild
        x[mask] = rng.integers(info.min, info.max, size=int(mask.sum()), endpoint=True).astype(dtype)
    return x


def device_signal(torch, shape, seed, wild=0.01):
    """The same recipe generated on the device (full-size configs)."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randint(0, 256, shape, generator=g, device="cuda", dtype=torch.int16)
    n = x.numel()
    n_wild = int(n * wild)
    where = torch.randint(0, n, (n_wild,), generator=g, device="cuda")
    x.view(-1)[where] = torch.randint(-32768, 32768, (n_wild,), generator=g, device="cuda", dtype=torch.int32).to(torch.int16)
    return x


# ---------------------------------------------------------------------------------------------------------------------
# live acc_bits wrap, taps beyond int16 (reference: fir_1d_fixed_ref.py:93,110-118)
# ---------------------------------------------------------------------------------------------------------------------
def test_reference_wrap_vectors_on_the_fast_kernels(fg, torch):
    """SURVEY 8(c)5 known answers, computed by the unmodified reference: the wrap happens once per output, the
    rounding bias is added after it and is not re-wrapped."""
    from firgpu import _lib
    x = np.array([255, 100, 7] * 2000, dtype=np.uint8)           # long enough to be a device call of useful size
    xt = torch.from_numpy(x).cuda()
    for acc_bits, want in ((16, [0, 4, 7]), (20, [0, 100, 7]), (21, [255, 100, 7])):
        y = fg.fir1d_fixed(xt, [1.0], 12, acc_bits, 16).cpu().numpy()
        assert _lib.last_path() == _lib.PATH_STAGED, acc_bits
        assert y[:3].tolist() == want and np.array_equal(y, np.tile(np.array(want, np.uint8), 2000)), acc_bits
    # coeff_bits = 32, frac_bits = 28: taps far beyond int16 -> IMAD datapath, still the staged kernel
    x3 = np.array([255, 1, 2], dtype=np.uint8)
    for acc_bits, want in ((32, [0, 8, 1]),):
        y = fg.fir1d_fixed(torch.from_numpy(x3).cuda(), [7.5, -7.5, 0.001], 28, acc_bits, 32).cpu().numpy()
        assert _lib.last_path() == _lib.PATH_STAGED
        assert y.tolist() == want
    y = fg.fir1d_fixed(torch.from_numpy(x3).cuda(), [7.5, -7.5, 0.001], 28, 64, 32).cpu().numpy()
    assert _lib.last_path() == _lib.PATH_GENERIC and y.tolist() == [0, 8, 0]          # 64-bit accumulator: int64 kernel


def test_golden_wrap_cases_now_run_on_the_staged_kernel(fg, torch, golden_cases):
    """Every reference-made golden case whose accumulator is at most 32 bits wide and whose filter has at most 9
    taps is served by a fast kernel, bit-exact (device path; the golden file also holds the wrap-live cases)."""
    from firgpu import _lib
    served = wrap_live = 0
    for c in golden_cases:
        if c["x"].size == 0:
            continue
        xt = torch.from_numpy(np.ascontiguousarray(c["x"])).cuda()
        y = fg.fir1d_fixed(xt, c["h"], c["frac_bits"], c["acc_bits"], c["coeff_bits"]).cpu().numpy()
        assert np.array_equal(y, c["y_fixed"]), c["name"]
        hq = onp.quantize_taps(c["h"], c["frac_bits"], c["coeff_bits"])
        int_container = c["x"].dtype in (np.uint8, np.int16)
        if int_container and c["acc_bits"] <= 32 and c["frac_bits"] <= 31 and len(c["h"]) <= 9:
            assert _lib.last_path() == _lib.PATH_STAGED, c["name"]
            served += 1
            if 255 * int(np.abs(hq).sum()) >= 2 ** (c["acc_bits"] - 1):
                wrap_live += 1
    assert served > 50 and wrap_live > 5, (served, wrap_live)


@pytest.mark.parametrize("in_dtype", [np.uint8, np.int16])
@pytest.mark.parametrize("out_dtype", [np.uint8, np.int16])
def test_wrap_sweep_against_the_oracle(fg, torch, in_dtype, out_dtype):
    """acc_bits 8 ... 32 x frac_bits x coeff_bits 8 / 16 / 32 x 1 ... 9 taps x both dilations, tile edges and ragged
    tails: the accumulator really wraps in most of these, and taps of 32 bits do not fit the dp2a operands."""
    from firgpu import _lib
    rng = np.random.default_rng(2024)
    n_live = 0
    for n_taps in range(1, 10):
        for acc_bits, frac_bits, coeff_bits in ((8, 3, 8), (12, 6, 8), (16, 12, 16), (20, 12, 16), (24, 15, 16), (24, 20, 32),
                                                (31, 15, 16), (32, 28, 32), (32, 31, 32), (17, 1, 16)):
            lim = min(7.9, (2 ** (coeff_bits - 1) - 1) / 2 ** frac_bits)
            h = rng.uniform(-lim, lim, size=n_taps).tolist()
            hq = onp.quantize_taps(h, frac_bits, coeff_bits)
            n_live += 255 * int(np.abs(hq).sum()) >= 2 ** (acc_bits - 1)
            for dil, n in ((1, 5), (1, 16385), (2, 40_002), (1, 70_001)):
                x = signal(rng, n, in_dtype)
                y = fg.fir1d_fixed(torch.from_numpy(x).cuda(), h, frac_bits, acc_bits, coeff_bits, dilation=dil, out_dtype=out_dtype)
                assert _lib.last_path() == _lib.PATH_STAGED, (n_taps, acc_bits, frac_bits, coeff_bits, dil, n)
                want = c_oracle.fir1d_fixed(x, hq, frac_bits, acc_bits, dilation=dil, out_dtype=out_dtype, threads=4)
                assert np.array_equal(y.cpu().numpy(), want), (n_taps, acc_bits, frac_bits, coeff_bits, dil, n)
    assert n_live > 40
    # rows with pitch, shard halos and the long-filter buckets keep working under a live wrap
    from firgpu import shard
    h = [0.9, -0.7, 0.8, 0.3, -0.2]
    rows = signal(rng, (5, 33_000), in_dtype)
    view = torch.from_numpy(rows).cuda()[:, :32_999]
    got = fg.fir1d_fixed(view, h, 12, 16, 16, out_dtype=out_dtype)
    assert _lib.last_path() == _lib.PATH_STAGED
    assert np.array_equal(got.cpu().numpy(), onp.fir1d_fixed(rows[:, :32_999], h, 12, 16, 16, out_dtype=out_dtype))
    sig = signal(rng, 150_000, in_dtype)
    want = onp.fir1d_fixed(sig, h, 12, 16, 16, out_dtype=out_dtype)
    for p in shard.plan_segments(sig.size, 3, 5, 1, align=4096):
        buf, body = shard.scatter_segment(sig, p, "cuda")
        y = fg.fir1d_fixed(body, h, 12, 16, 16, out_dtype=out_dtype, halo_left_valid=p.before, halo_right_valid=p.after)
        assert _lib.last_path() == _lib.PATH_STAGED
        assert np.array_equal(y.cpu().numpy(), want[p.start:p.stop])
    if in_dtype == out_dtype:
        h21 = rng.uniform(-0.9, 0.9, size=21).tolist()
        got = fg.fir1d_fixed(torch.from_numpy(sig).cuda(), h21, 12, 18, 16, out_dtype=out_dtype)
        assert _lib.last_path() == _lib.PATH_STAGED
        assert np.array_equal(got.cpu().numpy(), onp.fir1d_fixed(sig, h21, 12, 18, 16, out_dtype=out_dtype))


def test_wrap_host_entry_and_multi_filter(fg, torch):
    rng = np.random.default_rng(5)
    x = signal(rng, (300, 5001), np.uint8)
    sets = [[0.9, -0.7, 0.8], [0.25, 0.5, 0.25], [7.5, -7.0, 0.25], [1 / 3] * 3]
    want = np.stack([onp.fir1d_fixed(x, h, 12, 16, 16) for h in sets])
    assert np.array_equal(fg.fir1d_fixed_multi(torch.from_numpy(x).cuda(), sets, 12, 16, 16).cpu().numpy(), want)
    assert np.array_equal(fg.fir1d_fixed_multi(x, sets, 12, 16, 16), want)
    assert np.array_equal(fg.fir1d_fixed(x, sets[0], 12, 16, 16), want[0])


# ---------------------------------------------------------------------------------------------------------------------
# dense row batches: one launch, row ends inside the kernel (gen_fixed_output.py:44-59 semantics)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("in_dtype", [np.uint8, np.int16])
def test_dense_row_batches_are_one_launch(fg, torch, in_dtype):
    """Every row keeps its own zero padding although the batch is staged as one long row: all widths from the
    narrowest the kernel takes up to several tiles, odd and even, every tap count, both output containers,
    interleaved I/Q, the four-tap-set pass, live wrap."""
    from firgpu import _lib
    from fir_1d.sim.vector.h_coeff import h_coeff_3tap_map, h_coeff_5tap_map
    rng = np.random.default_rng(31)
    widths = (24, 25, 31, 32, 33, 40, 63, 64, 65, 100, 127, 128, 255, 256, 500, 641, 1280, 4499, 16385, 40_000)
    for cols in widths:
        rows = max(2, min(700, 300_000 // cols))
        x = signal(rng, (rows, cols), in_dtype, wild=0.03)
        xt = torch.from_numpy(x).cuda()
        for n_taps in (1, 2, 3, 4, 5, 8, 9):
            h = rng.uniform(-0.4, 0.4, size=n_taps).tolist()
            for out_dtype in (np.uint8, np.int16):
                before = _lib.launch_count()
                got = fg.fir1d_fixed(xt, h, 12, 32, 16, out_dtype=out_dtype)
                assert _lib.last_path() == _lib.PATH_STAGED
                if cols >= 32:
                    assert _lib.launch_count() - before == 1, (cols, n_taps)
                assert np.array_equal(got.cpu().numpy(), onp.fir1d_fixed(x, h, 12, 32, 16, out_dtype=out_dtype)), (cols, n_taps, out_dtype)
        if cols % 2 == 0:
            got = fg.fir1d_fixed(xt, SIMPLE_LP5, 15, 32, 16, dilation=2, out_dtype=in_dtype)
            assert np.array_equal(got.cpu().numpy(), onp.fir1d_fixed(x, SIMPLE_LP5, 15, 32, 16, dilation=2, out_dtype=in_dtype)), cols
        got = fg.fir1d_fixed(xt, [0.9, -0.7, 0.8, 0.3, -0.2], 12, 16, 16)               # live wrap
        assert np.array_equal(got.cpu().numpy(), onp.fir1d_fixed(x, [0.9, -0.7, 0.8, 0.3, -0.2], 12, 16, 16)), cols
        if in_dtype == np.uint8:
            for sets in (list(h_coeff_3tap_map.values()), list(h_coeff_5tap_map.values())):
                before = _lib.launch_count()
                got = fg.fir1d_fixed_multi(xt, sets)
                if cols >= 32:
                    assert _lib.launch_count() - before == 1, cols
                assert np.array_equal(got.cpu().numpy(), np.stack([onp.fir1d_fixed(x, h) for h in sets])), cols


def test_reference_largest_image_is_one_launch(fg, torch, image_large, manifest):
    """The reference's 2999 x 4499 image (odd pitch): four coefficient sets per launch, hashes of the reference
    pipeline's own output files."""
    import hashlib
    from firgpu import _lib
    stem, x = image_large
    xt = torch.from_numpy(x).cuda()
    for tap in ("3tap", "5tap"):
        names, sets = zip(*manifest[f"coeff_{tap}"].items())
        before = _lib.launch_count()
        got = fg.fir1d_fixed_multi(xt, list(sets))
        assert _lib.launch_count() - before == 1
        for name, y in zip(names, got.cpu().numpy()):
            assert hashlib.sha256(np.ascontiguousarray(y).tobytes()).hexdigest() == manifest["fixed"][f"{stem}__{name}_{tap}"], (name, tap)


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE configs 3, 4, 5 at full size
# ---------------------------------------------------------------------------------------------------------------------
def test_config3_full_size_batched_channels(fg, torch):
    """8192 channels x 2^20 int16 samples, 5-tap Q1.15 (17.2 GB in + 17.2 GB out resident): 64 random rows in full and
    both ends of EVERY row against the C oracle, plus the ragged variant cols = 1 000 000 (pitched view and dense)."""
    from firgpu import _lib
    rows, cols = 8192, 1 << 20
    free, _ = torch.cuda.mem_get_info()
    if free < 40 * 2 ** 30:
        pytest.skip(f"needs 40 GB of free HBM, have {free / 2 ** 30:.0f}")
    hq = onp.quantize_taps(SIMPLE_LP5, 15, 16)
    x = torch.empty((rows, cols), dtype=torch.int16, device="cuda")
    for b in range(0, rows, 1024):                                        # seed 1234 + row block
        x[b:b + 1024] = device_signal(torch, (1024, cols), 1234 + b // 1024)
    y = fg.fir1d_fixed(x, SIMPLE_LP5, 15, 32, 16, out_dtype=np.int16)
    assert _lib.last_path() == _lib.PATH_STAGED
    rng = np.random.default_rng(3)
    pick = np.sort(rng.choice(rows, size=64, replace=False))
    idx = torch.from_numpy(pick).cuda()
    want = c_oracle.fir1d_fixed(x[idx].cpu().numpy(), hq, 15, 32, out_dtype=np.int16, threads=CORES)
    assert np.array_equal(y[idx].cpu().numpy(), want)
    edge = 4096
    head = c_oracle.fir1d_fixed(x[:, :edge + 2].cpu().numpy(), hq, 15, 32, out_dtype=np.int16, threads=CORES)
    assert np.array_equal(y[:, :edge].cpu().numpy(), head[:, :edge])
    tail = c_oracle.fir1d_fixed(x[:, -(edge + 2):].cpu().numpy(), hq, 15, 32, out_dtype=np.int16, threads=CORES)
    assert np.array_equal(y[:, -edge:].cpu().numpy(), tail[:, 2:])
    assert int(y.min()) >= 0 and int(y.max()) <= 255
    # rows are independent: a row filtered alone equals its row of the batch
    alone = fg.fir1d_fixed(x[4097], SIMPLE_LP5, 15, 32, 16, out_dtype=np.int16)
    assert torch.equal(alone, y[4097])
    # ragged: 1 000 000 columns as a pitched view of the same buffer (in place, output pitched too)
    rc = 1_000_000
    yv = torch.zeros_like(y[:2048])
    fg.fir1d_fixed(x[:2048, :rc], SIMPLE_LP5, 15, 32, 16, out_dtype=np.int16, out=yv[:, :rc])
    assert _lib.last_path() == _lib.PATH_STAGED
    assert torch.equal(yv[:, :rc - 2], y[:2048, :rc - 2])                 # same samples in reach -> same outputs
    assert int(yv[:, rc:].abs().max()) == 0                               # nothing written beyond the ragged end
    endw = c_oracle.fir1d_fixed(x[:2048, rc - 4098:rc].cpu().numpy(), hq, 15, 32, out_dtype=np.int16, threads=CORES)
    assert np.array_equal(yv[:, rc - 4096:rc].cpu().numpy(), endw[:, 2:])
    del yv, y
    dense = x[:1024, :rc].contiguous()
    yd = fg.fir1d_fixed(dense, SIMPLE_LP5, 15, 32, 16, out_dtype=np.int16)
    pick = np.sort(rng.choice(1024, size=16, replace=False))
    want = c_oracle.fir1d_fixed(dense[torch.from_numpy(pick).cuda()].cpu().numpy(), hq, 15, 32, out_dtype=np.int16, threads=CORES)
    assert np.array_equal(yd[torch.from_numpy(pick).cuda()].cpu().numpy(), want)


def test_config5_full_size_complex_iq(fg, torch):
    """2^27 complex samples = 2^28 interleaved int16 scalars, 5-tap Q1.15, real taps (lib/mycomplex.cpp:12-22): every
    output against the C oracle; Re and Im are two independent filters; segments with halos are seamless."""
    from firgpu import _lib, shard
    n = 1 << 28
    z = device_signal(torch, (n,), 5678)
    y = fg.fir1d_fixed_iq(z, SIMPLE_LP5, 15, 32, 16, out_dtype=np.int16)
    assert _lib.last_path() == _lib.PATH_STAGED
    hq = onp.quantize_taps(SIMPLE_LP5, 15, 16)
    full = c_oracle.fir1d_fixed(z.cpu().numpy(), hq, 15, 32, dilation=2, out_dtype=np.int16, threads=CORES)
    assert np.array_equal(y.cpu().numpy(), full)
    # the definition: I and Q filtered separately by the reference-pinned 1-D model (windows)
    for start in (0, n // 2 - 4096, n - (1 << 21)):
        zs = z[start:start + (1 << 21)].cpu().numpy()
        for c in (0, 1):
            want = c_oracle.fir1d_fixed(zs[c::2], hq, 15, 32, out_dtype=np.int16, threads=4)
            got = full[start:start + (1 << 21)][c::2]
            a, b = (0 if start == 0 else 2), (want.size if start + (1 << 21) == n else want.size - 2)
            assert np.array_equal(got[a:b], want[a:b]), (start, c)
    del full
    for parts in (2, 4, 8):                                               # the 8-GPU plan of BASELINE config 5, on one GPU
        for p in shard.plan_segments(n, parts, 5, 2, align=4096):
            lead = shard.aligned_lead(p.before, 2)
            buf = torch.zeros(lead + p.size + p.after, dtype=torch.int16, device="cuda")
            buf[lead - p.before:] = z[p.lo:p.hi]
            got = fg.fir1d_fixed(buf[lead:lead + p.size], SIMPLE_LP5, 15, 32, 16, dilation=2, out_dtype=np.int16,
                                 halo_left_valid=p.before, halo_right_valid=p.after)
            assert torch.equal(got, y[p.start:p.stop]), (parts, p.start)
            del buf, got


def test_config4_full_size_image_batch_and_band_seams(fg, torch):
    """[4, 16384, 16384] int16, 3x3 and 5x5 (rank-1 and general), Q4.12: crops at all four borders, the centre and
    every row-band seam of the 2 / 4 / 8-GPU plans against the C oracle, and every band of those plans (filtered as
    a shard with its halo rows) equal to the unsharded result."""
    from firgpu import _lib, quant, shard
    b_img, hgt, wid = 4, 16384, 16384
    x = device_signal(torch, (b_img, hgt, wid), 4321)
    rng = np.random.default_rng(4)
    kernels = {"box3": [[1 / 9] * 3] * 3, "sharpen3": [[0, -1, 0], [-1, 5, -1], [0, -1, 0]],
               "lp5 (rank-1)": np.outer(SIMPLE_LP5, SIMPLE_LP5).tolist(),
               "random5 (general)": (rng.uniform(-1, 1, size=(5, 5)) * (7.9 / 25)).tolist()}
    seams = sorted({b.start for parts in (2, 4, 8) for b in shard.plan_bands(hgt, parts, 5)} - {0})
    for name, k in kernels.items():
        hq = quant.quantize_taps(np.asarray(k), 12, 16)
        kh = hq.shape[0]
        y = fg.fir2d_fixed_q(x, hq, 12, 32, out_dtype=np.int16)
        assert _lib.last_path() == _lib.PATH_STAGED, name
        r = kh // 2
        crops = [(0, 0), (0, wid - 1024), (hgt - 1024, 0), (hgt - 1024, wid - 1024), (hgt // 2 - 512, wid // 2 - 512)]
        crops += [(s - 512, int(rng.integers(0, wid - 1024))) for s in seams]
        for img in (0, b_img - 1):
            for r0, c0 in crops:
                r1, c1 = r0 + 1024, c0 + 1024
                xs = x[img, r0:r1, c0:c1].cpu().numpy()
                want = c_oracle.fir2d_fixed(xs, hq, 12, 32, out_dtype=np.int16, threads=CORES)
                # the crop's own zero padding is right only where the crop touches the image border
                a0, a1 = (0 if r0 == 0 else r), (1024 if r1 == hgt else 1024 - r)
                b0, b1 = (0 if c0 == 0 else r), (1024 if c1 == wid else 1024 - r)
                assert np.array_equal(y[img, r0 + a0:r0 + a1, c0 + b0:c0 + b1].cpu().numpy(), want[a0:a1, b0:b1]), (name, img, r0, c0)
        for parts in (2, 4, 8):
            for band in shard.plan_bands(hgt, parts, kh):
                buf = x[:, band.lo:band.hi, :]                            # the band with its halo rows, in place
                got = fg.fir2d_fixed_q(buf[:, band.before:band.before + band.size, :], hq, 12, 32, out_dtype=np.int16,
                                       halo_top_valid=band.before, halo_bottom_valid=band.after)
                assert torch.equal(got, y[:, band.start:band.stop, :]), (name, parts, band.start)
                del got
        del y


# ---------------------------------------------------------------------------------------------------------------------
# streaming with the history on the device (fir_1d/model/cpp/fir_1d_ref.cpp:16-35)
# ---------------------------------------------------------------------------------------------------------------------
def test_stream_blocks_equal_one_call(fg, torch):
    """Any blocking of a signal — blocks shorter than the history included — gives the bits of one causal call over
    the whole signal: fixed-point (every container, I/Q, live wrap, signed range) and floating-point (raw float64 =
    the C++ Fir1D, and the rounded / clamped pixel model with float32 output), device blocks and host blocks."""
    from firgpu import _lib, quant
    rng = np.random.default_rng(77)
    n = 200_003
    cuts = [0, 1, 2, 3, 7, 8, 50, 4096, 4097, 70_000, 70_001, 150_000, n]
    for in_dtype in (np.uint8, np.int16, np.int32, np.float32, np.float64):
        x = (rng.uniform(-20, 300, size=n)).astype(in_dtype) if np.dtype(in_dtype).kind == "f" else signal(rng, n, in_dtype) \
            if in_dtype != np.int32 else rng.integers(-500, 700, size=n).astype(np.int32)
        xt = torch.from_numpy(x).cuda()
        for h, frac, acc, dil in ((SIMPLE_LP5, 15, 32, 1), ([0.9, -0.7, 0.8], 12, 16, 1), (SIMPLE_LP5, 15, 32, 2),
                                  (rng.uniform(-0.1, 0.1, size=17).tolist(), 12, 32, 1)):
            hq = quant.quantize_taps(h, frac, 16)
            want = onp.fir1d_fixed(x, h, frac, acc, 16, dilation=dil, centre=0, out_dtype=np.int16)
            s = fg.FirStream(hq, model="fixed", in_dtype=in_dtype, out_dtype=np.int16, dilation=dil, frac_bits=frac, acc_bits=acc)
            assert s.history == (len(h) - 1) * dil
            got = torch.cat([s.process(xt[a:b]) for a, b in zip(cuts[:-1], cuts[1:])])
            assert np.array_equal(got.cpu().numpy(), want), (in_dtype, len(h), dil, "device")
            s.reset()
            torch.cuda.synchronize()
            host = np.concatenate([s.process(x[a:b]) for a, b in zip(cuts[:-1], cuts[1:])])
            assert np.array_equal(host, want), (in_dtype, len(h), dil, "host")
            s.close()
        hh = [0.5, -0.25, 0.125, 0.3, -0.7]
        for out_dtype, raw in ((np.float64, False), (np.float32, False), (np.float64, True)):
            if raw and in_dtype != np.float64:
                continue
            one = fg.fir1d_ideal(xt, hh, centre=0, raw=raw, out_dtype=out_dtype)
            s = fg.FirStream(hh, model="ideal", in_dtype=in_dtype, out_dtype=out_dtype, raw=raw)
            got = torch.cat([s.process(xt[a:b]) for a, b in zip(cuts[:-1], cuts[1:])])
            assert torch.equal(got, one), (in_dtype, out_dtype, raw)
            if out_dtype == np.float64:
                assert np.array_equal(one.cpu().numpy(), onp.fir1d_ideal(x, hh, centre=0, raw=raw))
            s.close()
    # signed int16 range (the reference's stated future work), device blocks
    full = (-32768, 32767)
    xs = rng.integers(-32768, 32768, size=n).astype(np.int16)
    h = [0.1, -0.2, 0.4, -0.2, 0.1]
    s = fg.FirStream(quant.quantize_taps(h, 15, 16), model="fixed", in_dtype=np.int16, frac_bits=15, in_range=full, out_range=full)
    got = torch.cat([s.process(torch.from_numpy(xs[a:b]).cuda()) for a, b in zip(cuts[:-1], cuts[1:])])
    assert np.array_equal(got.cpu().numpy(), onp.fir1d_fixed(xs, h, 15, 32, 16, centre=0, in_range=full, out_range=full, out_dtype=np.int16))
    assert _lib.launch_count() > 0


def test_stream_full_size_signal_in_megasample_blocks(fg, torch):
    """A 2^28-sample int16 signal delivered as 256 device-resident blocks of 2^20 samples: bit-identical to one call,
    and to the C oracle of the causal model on windows around block seams."""
    from firgpu import quant
    n, blk = 1 << 28, 1 << 20
    x = device_signal(torch, (n,), 99)
    hq = quant.quantize_taps(SIMPLE_LP5, 15, 16)
    one = fg.fir1d_fixed(x, SIMPLE_LP5, 15, 32, 16, centre=0, out_dtype=np.int16)
    y = torch.empty_like(one)
    s = fg.FirStream(hq, model="fixed", in_dtype=np.int16, out_dtype=np.int16, frac_bits=15)
    for b in range(0, n, blk):
        s.process(x[b:b + blk], out=y[b:b + blk])
    assert torch.equal(y, one)
    for seam in (blk, 17 * blk, n - blk):
        xs = x[seam - 5000:seam + 5000].cpu().numpy()
        want = onp.fir1d_fixed(xs, SIMPLE_LP5, 15, 32, 16, centre=0, out_dtype=np.int16)
        assert np.array_equal(y[seam - 5000 + 4:seam + 5000].cpu().numpy(), want[4:])
    # a float64 stream (the C++ Fir1D) over a long host signal: pipelined host blocks
    from fir_1d.model.python.fir_1d_stream import Fir1D
    sig = np.random.default_rng(1).normal(0, 100, size=(1 << 24) + 3)
    h = [0.5, -0.25, 0.125, 0.3, -0.7]
    f = Fir1D(h)
    parts = [f.process_block(sig[a:b]) for a, b in ((0, 9_000_001), (9_000_001, sig.size))]
    assert np.array_equal(np.concatenate(parts), onp.fir1d_ideal(sig, h, centre=0, raw=True))
    dev = f.process_block(torch.from_numpy(sig[:1000]).cuda())                # device block: stays on the device
    assert dev.is_cuda and dev.dtype == torch.float64


# ---------------------------------------------------------------------------------------------------------------------
# recorded launch sequences
# ---------------------------------------------------------------------------------------------------------------------
def test_recorded_sequences_replay_bit_exact(fg, torch, images_inputs, manifest):
    """firgpu_record_* / firgpu_graph_launch through the raw-stream context manager (preallocated outputs) and through
    the torch wrapper: a replay recomputes from the buffers' CURRENT contents; the reference's images x coefficient
    sets become one submission."""
    from firgpu import _lib, quant
    rng = np.random.default_rng(8)
    hq = quant.quantize_taps(SIMPLE_LP5, 15, 16)
    x = torch.from_numpy(signal(rng, 1 << 20, np.int16)).cuda()
    y = torch.empty_like(x)
    img = torch.from_numpy(signal(rng, (300, 4499), np.uint8)).cuda()
    yi = torch.empty_like(img)
    hq3 = quant.quantize_taps([0.25, 0.5, 0.25], 12, 16)

    def calls():
        fg.fir1d_fixed_q(x, hq, 15, 32, out_dtype=np.int16, out=y)
        fg.fir1d_fixed_q(img, hq3, 12, 32, out_dtype=np.uint8, out=yi)

    calls()                                                                  # eager first use
    side = torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        with fg.record(side.cuda_stream) as rec:
            calls()
    assert rec.kernels == 2
    for seed in (1, 2):
        x.copy_(torch.from_numpy(signal(np.random.default_rng(seed), 1 << 20, np.int16)))
        img.copy_(torch.from_numpy(signal(np.random.default_rng(seed), (300, 4499), np.uint8)))
        y.zero_()
        yi.zero_()
        torch.cuda.synchronize()
        rec.replay()
        side.synchronize()
        assert np.array_equal(y.cpu().numpy(), onp.fir1d_fixed(x.cpu().numpy(), SIMPLE_LP5, 15, 32, 16, out_dtype=np.int16))
        assert np.array_equal(yi.cpu().numpy(), onp.fir1d_fixed(img.cpu().numpy(), [0.25, 0.5, 0.25]))
    rec.close()
    # torch wrapper: the reference's six small images x 8 coefficient sets (allocating outputs inside)
    images = [torch.from_numpy(images_inputs[k]).cuda() for k in images_inputs.keys()]
    sets3, sets5 = list(manifest["coeff_3tap"].values()), list(manifest["coeff_5tap"].values())
    rec = fg.Recording(lambda: [(fg.fir1d_fixed_multi(im, sets3), fg.fir1d_fixed_multi(im, sets5)) for im in images])
    assert rec.kernels == 2 * len(images)                                    # one launch per image and tap count
    before = _lib.launch_count()
    outs = rec.replay()
    torch.cuda.synchronize()
    for im, (o3, o5) in zip(images, outs):
        xs = im.cpu().numpy()
        assert np.array_equal(o3.cpu().numpy(), np.stack([onp.fir1d_fixed(xs, h) for h in sets3]))
        assert np.array_equal(o5.cpu().numpy(), np.stack([onp.fir1d_fixed(xs, h) for h in sets5]))
    assert _lib.launch_count() == before                                     # (torch replays its own graph object)


# ---------------------------------------------------------------------------------------------------------------------
# host entry points: several devices, pipelined compare
# ---------------------------------------------------------------------------------------------------------------------
def test_multi_device_host_entry(fg, torch):
    """firgpu_fir1d_fixed_host_multi: one host array partitioned over a device list (the same device may appear
    more than once: the partitioning, halo reads from the host array and per-share pipelines are what is tested)."""
    rng = np.random.default_rng(12)
    n_dev = torch.cuda.device_count()
    lists = [[0], [0, 0], [0, 0, 0]] + ([list(range(n_dev))] if n_dev > 1 else [])
    for shape in ((1, 5_000_003), (37, 70_001), (2, 300_000)):
        x = signal(rng, shape, np.int16)
        want = c_oracle.fir1d_fixed(x, onp.quantize_taps(SIMPLE_LP5, 15, 16), 15, 32, out_dtype=np.int16, threads=CORES)
        for devs in lists:
            y = fg.fir1d_fixed_multi_gpu(x, SIMPLE_LP5, 15, 32, 16, devices=devs, out_dtype=np.int16)
            assert np.array_equal(y, want), (shape, devs)
    z = signal(rng, (1, 2_000_002), np.int16)
    want = c_oracle.fir1d_fixed(z, onp.quantize_taps(SIMPLE_LP5, 15, 16), 15, 32, dilation=2, out_dtype=np.int16, threads=CORES)
    assert np.array_equal(fg.fir1d_fixed_multi_gpu(z, SIMPLE_LP5, 15, 32, 16, devices=[0, 0, 0], dilation=2, out_dtype=np.int16), want)


def test_compare_host_is_pipelined_and_agrees_with_the_device_entry(fg, torch):
    rng = np.random.default_rng(13)
    x = signal(rng, (6000, 4499), np.uint8)
    a = fg.compare_metrics(x, [0.25, 0.5, 0.25])
    b = fg.compare_metrics(torch.from_numpy(x).cuda(), [0.25, 0.5, 0.25])
    for k in a:
        assert a[k] == pytest.approx(b[k], rel=1e-12, abs=1e-15), k
    assert a["num_samples"] == x.size


# ---------------------------------------------------------------------------------------------------------------------
# argument checks (advisor findings)
# ---------------------------------------------------------------------------------------------------------------------
def test_mismatched_out_tensors_are_rejected_before_the_kernel_writes(fg, torch):
    x = torch.zeros(4096, dtype=torch.int16, device="cuda")
    with pytest.raises(ValueError, match="out must be"):
        fg.fir1d_ideal(x, [1.0], out=torch.empty(4096, dtype=torch.float32, device="cuda"))          # default out_dtype is float64
    with pytest.raises(ValueError, match="out must be"):
        fg.fir1d_fixed(x, [1.0], out_dtype=np.int16, out=torch.empty(4096, dtype=torch.uint8, device="cuda"))
    with pytest.raises(ValueError, match="out must be"):
        fg.fir1d_fixed(x, [1.0], out=torch.empty(4096, dtype=torch.uint8))                           # host tensor
    with pytest.raises(ValueError, match="out must be"):
        fg.fir2d_fixed(x.view(64, 64), [[1.0]], out=torch.empty((64, 64), dtype=torch.int16, device="cuda"))
    if torch.cuda.device_count() > 1:
        with pytest.raises(ValueError, match="out must be"):
            fg.fir1d_fixed(x, [1.0], out=torch.empty(4096, dtype=torch.uint8, device="cuda:1"))
    # raw device pointers: extended parameters are honoured, conflicting containers rejected
    xs = torch.from_numpy(np.arange(-3000, 3000, dtype=np.int16)).cuda()
    ys = torch.empty_like(xs)
    full = (-32768, 32767)
    fg.fir1d_fixed(fg.DeviceArray(xs.data_ptr(), (xs.numel(),), np.int16, device=0), [0.5], 15, 32, 16, centre=0, in_range=full,
                   out_range=full, out=fg.DeviceArray(ys.data_ptr(), (ys.numel(),), np.int16, device=0))
    torch.cuda.synchronize()
    assert np.array_equal(ys.cpu().numpy(), onp.fir1d_fixed(xs.cpu().numpy(), [0.5], 15, 32, 16, centre=0, in_range=full,
                                                            out_range=full, out_dtype=np.int16))
    with pytest.raises(ValueError, match="out must be"):
        fg.fir1d_fixed(fg.DeviceArray(xs.data_ptr(), (xs.numel(),), np.int16), [0.5], out=fg.DeviceArray(ys.data_ptr(), (5,), np.int16))


def test_fir2d_live_wrap_runs_on_the_tensor_map_kernel(fg, torch):
    """The 2-D model inherits the 1-D post stage (wrap to acc_bits once per pixel, bias not re-wrapped): narrow
    accumulators are served by the tensor-map kernel's WRAP post mode — general, rank-1 and single-row / single-column
    windows, uint8 / int16 pixels, both outputs, band shards with halo rows — bit-exact against the oracle."""
    from firgpu import _lib, quant, shard
    rng = np.random.default_rng(71)
    kernels = {"3x3": rng.uniform(-0.9, 0.9, size=(3, 3)), "5x5": rng.uniform(-0.5, 0.5, size=(5, 5)),
               "5x5 rank-1": np.outer([0.9, -0.7, 0.8, 0.3, -0.2], [0.5, 0.9, -0.8, 0.7, 0.4]), "3x5": rng.uniform(-0.6, 0.6, size=(3, 5)),
               "5x1": rng.uniform(-0.9, 0.9, size=(5, 1)), "1x5": rng.uniform(-0.9, 0.9, size=(1, 5))}
    n_live = 0
    for dtype in (np.uint8, np.int16):
        img = signal(rng, (2, 150, 1000), dtype, wild=0.03)
        it = torch.from_numpy(img).cuda()
        for name, k in kernels.items():
            for acc_bits in (12, 16, 20):
                hq = quant.quantize_taps(k, 12, 16)
                n_live += 255 * int(np.abs(hq).sum()) >= 2 ** (acc_bits - 1)
                for out_dtype in ((np.uint8, np.int16) if dtype == np.int16 else (np.uint8,)):
                    got = fg.fir2d_fixed(it, k.tolist(), 12, acc_bits, 16, out_dtype=out_dtype)
                    assert _lib.last_path() == _lib.PATH_STAGED, (dtype, name, acc_bits, out_dtype)
                    assert np.array_equal(got.cpu().numpy(), onp.fir2d_fixed(img, k.tolist(), 12, acc_bits, 16, out_dtype=out_dtype)), (dtype, name, acc_bits)
    assert n_live >= 20
    k = kernels["5x5"]
    hq = quant.quantize_taps(k, 12, 16)
    img = signal(rng, (1, 300, 640), np.int16)
    want = onp.fir2d_fixed(img, k.tolist(), 12, 16, 16, out_dtype=np.int16)
    it = torch.from_numpy(img).cuda()
    for band in shard.plan_bands(300, 3, 5):
        got = fg.fir2d_fixed_q(it[:, band.lo:band.hi, :][:, band.before:band.before + band.size, :], hq, 12, 16, out_dtype=np.int16,
                               halo_top_valid=band.before, halo_bottom_valid=band.after)
        assert np.array_equal(got.cpu().numpy(), want[:, band.start:band.stop, :]), band.start
    # short pitched rows of a 1-D batch under a live wrap: the row-streaming form of the same kernel
    rows = signal(rng, (64, 3008), np.int16)
    view = torch.from_numpy(rows).cuda()[:, :3000]
    h = [0.9, -0.7, 0.8, 0.3, -0.2]
    got = fg.fir1d_fixed(view, h, 12, 16, 16, out_dtype=np.int16)
    assert _lib.last_path() == _lib.PATH_STAGED
    assert np.array_equal(got.cpu().numpy(), onp.fir1d_fixed(rows[:, :3000], h, 12, 16, 16, out_dtype=np.int16))


def test_complex_taps_on_interleaved_iq(fg, torch):
    """Complex taps (SURVEY 8(f)4's last step; specification ours, parity unpinned — the reference's MyComplex has no
    complex product): the product composes them from two real filters over the interleaved stream, the oracle
    evaluates the definition directly.  Anchors: purely real taps reduce to the real-tap I/Q filter (pinned by
    composition), h = [j] rotates every sample by 90 degrees, 1 ... 8 complex taps, reference and signed ranges, live
    wrap, several rows, the trailing-axis-of-2 layout, host arrays long enough to be cut into segments."""
    from firgpu import _lib
    rng = np.random.default_rng(81)
    full = (-32768, 32767)
    for n_taps in (1, 2, 3, 4, 5, 8):
        h = (rng.uniform(-0.4, 0.4, size=n_taps) + 1j * rng.uniform(-0.4, 0.4, size=n_taps)).tolist()
        for shape in ((2,), (10,), (16386,), (3, 5000), (70_002,)):
            for in_r, out_r, dt in (((0, 255), (0, 255), np.uint8), ((0, 255), (0, 255), np.int16), (full, full, np.int16)):
                z = signal(rng, shape, dt, wild=0.05) if in_r == (0, 255) else rng.integers(-3000, 3000, size=shape).astype(np.int16)
                for acc_bits in (32, 18):
                    got = fg.fir1d_fixed_complex(torch.from_numpy(z).cuda(), h, 12, acc_bits, 16, in_range=in_r, out_range=out_r, out_dtype=np.int16)
                    want = onp.fir1d_fixed_complex(z, h, 12, acc_bits, 16, in_range=in_r, out_range=out_r, out_dtype=np.int16)
                    assert np.array_equal(got.cpu().numpy(), want), (n_taps, shape, in_r, acc_bits)
    # fast kernels serve it: up to four signed complex taps on the wide datapath, reference ranges through the buckets
    z = rng.integers(-3000, 3000, size=1 << 20).astype(np.int16)
    fg.fir1d_fixed_complex(torch.from_numpy(z).cuda(), [0.2 + 0.1j, 0.3 - 0.2j, 0.1 + 0.3j], 15, 32, 16, in_range=full, out_range=full)
    assert _lib.last_path() == _lib.PATH_STAGED
    # anchors
    zr = signal(rng, 100_000, np.int16)
    h_real = [0.1, 0.2, 0.4, 0.2, 0.1]
    a = fg.fir1d_fixed_complex(torch.from_numpy(zr).cuda(), h_real, 12, 32, 16, out_dtype=np.int16).cpu().numpy()
    b = fg.fir1d_fixed_iq(torch.from_numpy(zr).cuda(), h_real, 12, 32, 16, out_dtype=np.int16).cpu().numpy()
    assert np.array_equal(a, b) and np.array_equal(a, onp.fir1d_fixed_iq(zr, h_real, 12, 32, 16, out_dtype=np.int16))
    zs = rng.integers(-1000, 1000, size=20_000).astype(np.int16)
    rot = fg.fir1d_fixed_complex(torch.from_numpy(zs).cuda(), [1j], 12, 32, 16, in_range=full, out_range=full).cpu().numpy()
    assert np.array_equal(rot[0::2], -zs[1::2]) and np.array_equal(rot[1::2], zs[0::2])
    # [N, 2] layout and the host entry point (long enough for column segments with halos)
    pairs = zs.reshape(-1, 2)
    h = [0.3 + 0.1j, -0.2 + 0.25j, 0.15 - 0.3j]
    want = onp.fir1d_fixed_complex(zs, h, 12, 32, 16, in_range=full, out_range=full, out_dtype=np.int16)
    got = fg.fir1d_fixed_complex(torch.from_numpy(pairs).cuda(), h, 12, 32, 16, in_range=full, out_range=full)
    assert got.shape == pairs.shape and np.array_equal(got.cpu().numpy().reshape(-1), want)
    big = rng.integers(-3000, 3000, size=40_000_000).astype(np.int16)
    want = onp.fir1d_fixed_complex(big, h, 12, 32, 16, in_range=full, out_range=full, out_dtype=np.int16)
    assert np.array_equal(fg.fir1d_fixed_complex(big, h, 12, 32, 16, in_range=full, out_range=full), want)


def test_ideal_wide_groups_every_alignment_and_the_fallback():
    """The floating-point model's 32-byte-group form (one 256-bit store per lane, aligned 4-sample chunk loads) is taken
    where input and output rows share their alignment, the 16-byte form elsewhere; both must give the oracle's result —
    float64 bit for bit, float32 within the stated tolerance — at every tap count, ragged widths and view offsets."""
    import torch
    import firgpu
    from firgpu import _lib
    rng = np.random.default_rng(77)
    for dtype in (np.uint8, np.int16):
        base = rng.integers(0, 256, size=(5, 9000)).astype(dtype)
        if dtype == np.int16:
            base[rng.random(base.shape) < 0.05] = -7
            base[rng.random(base.shape) < 0.05] = 300
        bt = torch.from_numpy(base).cuda()
        for n_taps in (1, 2, 3, 5, 8, 9):
            h = rng.uniform(-2, 2, size=n_taps).tolist()
            tol = 1e-6 * 255 * float(np.sum(np.abs(h)))
            for cx, cy, width in ((0, 0, 8192), (0, 0, 4107), (4, 4, 3001), (3, 3, 519), (1, 0, 2050), (0, 5, 777), (8, 0, 6), (2, 2, 31)):
                want = onp.fir1d_ideal(base[:, cx:cx + width], h)
                for odt, tdt in ((np.float64, torch.float64), (np.float32, torch.float32)):
                    out = torch.full((5, 9100), -3.0, dtype=tdt, device="cuda")
                    firgpu.fir1d_ideal(bt[:, cx:cx + width], h, out_dtype=odt, out=out[:, cy:cy + width])
                    assert _lib.last_path() == _lib.PATH_STAGED
                    got = out.cpu().numpy()
                    if odt is np.float64:
                        assert np.array_equal(got[:, cy:cy + width], want), (dtype, n_taps, cx, cy, width)
                    else:
                        assert np.abs(got[:, cy:cy + width] - want).max() <= tol, (dtype, n_taps, cx, cy, width)
                    assert (got[:, :cy] == -3.0).all() and (got[:, cy + width:] == -3.0).all()
    # one long row with shard halos on the wide form
    x = rng.integers(0, 256, size=(1 << 20) + 24).astype(np.uint8)
    h = [0.1, 0.2, 0.4, 0.2, 0.1]
    full = onp.fir1d_ideal(x, h)
    xt = torch.from_numpy(x).cuda()
    body = xt[8:8 + (1 << 20)]
    y = firgpu.fir1d_ideal(body, h, halo_left_valid=2, halo_right_valid=2)
    assert np.array_equal(y.cpu().numpy(), full[8:8 + (1 << 20)])


def test_2d_ideal_rows_kernel_bit_identical():
    """fir2d_ideal_rows_kernel (four output rows per lane, 32-byte groups; taken when pitches and offsets are multiples of
    the group): every window, heights that are not multiples of four, widths ragged inside an aligned pitch, view offsets,
    several images, row bands with halo rows — float64 bit-identical to the oracle, float32 within the tolerance, and
    nothing written outside the view."""
    import torch
    import firgpu
    from firgpu import _lib, shard
    rng = np.random.default_rng(123)
    kernels = [rng.uniform(-2, 2, size=s).tolist() for s in ((3, 3), (5, 5), (3, 5), (5, 3))]
    for dtype in (np.uint8, np.int16):
        base = rng.integers(0, 256, size=(3, 43, 1024)).astype(dtype)
        if dtype == np.int16:
            base[rng.random(base.shape) < 0.05] = -9
            base[rng.random(base.shape) < 0.05] = 999
        bt = torch.from_numpy(base).cuda()
        for k in kernels:
            tol = 1e-6 * 255 * float(np.abs(k).sum())
            for (r0, hgt), (c0, width) in (((0, 43), (0, 1024)), ((0, 1), (0, 1000)), ((2, 2), (8, 520)), ((1, 7), (16, 9)),
                                           ((0, 40), (3, 800)), ((5, 13), (8, 1016))):
                view = bt[:, r0:r0 + hgt, c0:c0 + width]
                want = onp.fir2d_ideal(base[:, r0:r0 + hgt, c0:c0 + width], k)
                for odt, tdt in ((np.float64, torch.float64), (np.float32, torch.float32)):
                    out = torch.full((3, 43, 1024), -5.0, dtype=tdt, device="cuda")
                    firgpu.fir2d_ideal(view, k, out_dtype=odt, out=out[:, r0:r0 + hgt, c0:c0 + width])
                    assert _lib.last_path() == _lib.PATH_STAGED
                    got = out.cpu().numpy()
                    inner = got[:, r0:r0 + hgt, c0:c0 + width]
                    if odt is np.float64:
                        assert np.array_equal(inner, want), (dtype, np.shape(k), r0, hgt, c0, width)
                    else:
                        assert np.abs(inner - want).max() <= tol, (dtype, np.shape(k), r0, hgt, c0, width)
                    got[:, r0:r0 + hgt, c0:c0 + width] = -5.0
                    assert (got == -5.0).all(), "wrote outside the view"
    x = rng.integers(0, 256, size=(1, 301, 2048)).astype(np.uint8)
    k = kernels[1]
    want = onp.fir2d_ideal(x, k)
    got = np.empty_like(want)
    for band in shard.plan_bands(301, 3, 5):
        buf, body = shard.scatter_band(x, band, "cuda")
        y = firgpu.fir2d_ideal(body, k, halo_top_valid=band.before, halo_bottom_valid=band.after)
        got[..., band.start:band.stop, :] = y.cpu().numpy()
    assert np.array_equal(got, want)


def test_u8_in_int16_out_store_alignments():
    """uint8 samples into an int16 container: a thread's 16 outputs are one 256-bit store where the destination is 32-byte
    aligned and two 128-bit stores elsewhere; every alignment of the output view must give the oracle's values and write
    nothing outside it."""
    import torch
    import firgpu
    rng = np.random.default_rng(5)
    x = rng.integers(0, 256, size=(3, 70_000)).astype(np.uint8)
    xt = torch.from_numpy(x).cuda()
    for h in ([0.0625, 0.25, 0.375, 0.25, 0.0625], [-0.25, 1.5, -0.25]):
        for c0, width in ((0, 70_000), (8, 65_536), (16, 40_000), (3, 33_333), (24, 1_000)):
            out = torch.full((3, 70_100), -77, dtype=torch.int16, device="cuda")
            firgpu.fir1d_fixed(xt[:, c0:c0 + width], h, out_dtype=np.int16, out=out[:, c0:c0 + width])
            got = out.cpu().numpy()
            want = onp.fir1d_fixed(x[:, c0:c0 + width], h).astype(np.int16)
            assert np.array_equal(got[:, c0:c0 + width], want), (len(h), c0, width)
            assert (got[:, :c0] == -77).all() and (got[:, c0 + width:] == -77).all(), (len(h), c0, width)
    flat = torch.from_numpy(x.reshape(-1)).cuda()
    y = firgpu.fir1d_fixed(flat, [0.25, 0.5, 0.25], out_dtype=np.int16)
    assert np.array_equal(y.cpu().numpy(), onp.fir1d_fixed(x.reshape(-1), [0.25, 0.5, 0.25]).astype(np.int16))


# ---------------------------------------------------------------------------------------------------------------------
# general 5x5: the newest kernel rows accumulated by FFMA, exactly (csrc/fir2d_tiled.cuh, template parameter FR)
# ---------------------------------------------------------------------------------------------------------------------
def test_general_5x5_ffma_rows_are_bit_exact(fg, torch, monkeypatch):
    """0 / 1 / 2 kernel rows of the general 5x5 kernel run as float multiply-adds whose partial sums the host proved to
    stay inside one 2^23-wide window of exactly representable integers.  Every setting must equal the integer oracle:
    saturated regions (all pixels 255 under taps at the admission limit 255 * sum|taps| < 2^23, one count below it and
    one above it, where the library must fall back to fewer rows), all-negative rows (the window's base), the
    shift-by-16 pre-scale (Q1.15: float sum shifted left by one), plain shift, the live acc_bits wrap, uint8 / int16
    pixels and the narrowing output, ragged strips and bands.  The FFMA variants were measured slower than dp2a alone
    (profiles/README.md R2.5c) and are compiled only with -DFIRGPU_T2_FFMA=1; a library without them ignores the
    setting and the three passes check the shipped kernel against the same edge cases."""
    from firgpu import _lib
    rng = np.random.default_rng(2025)

    def kernel(row0, row1=None, scale=200):
        hq = rng.integers(-scale, scale + 1, size=(5, 5)).astype(np.int32)
        hq[0] = row0
        if row1 is not None:
            hq[1] = row1
        return hq

    kernels = [
        ("random q4.12", rng.integers(-1200, 1201, size=(5, 5)).astype(np.int32), 12, 32),
        ("random q1.15 (pre-scale 2)", rng.integers(-3000, 3001, size=(5, 5)).astype(np.int32), 15, 32),
        ("random q4.12, 20-bit accumulator", rng.integers(-1200, 1201, size=(5, 5)).astype(np.int32), 12, 20),
        ("random q4.12, live 16-bit wrap", rng.integers(-1200, 1201, size=(5, 5)).astype(np.int32), 12, 16),
        ("row 0 all negative", kernel([-900, -1, -2000, -30, -1500]), 12, 32),
        ("row 0 at the limit", kernel([8224, -8224, 8224, -8224, 0]), 12, 32),              # 255 * 32896 = 2^23 - 128
        ("row 0 at the limit, negative", kernel([-8224, -8224, -8224, -8223, -1]), 12, 32),
        ("row 0 one above the limit", kernel([8224, -8224, 8225, -8224, 0]), 12, 32),       # falls back to dp2a
        ("rows 0 + 1 at the limit", kernel([4000, -4000, 4000, -4000, 448], [4000, 4000, -4000, -4000, 448]), 12, 32),
        ("rows 0 + 1 above, row 0 below", kernel([4000, -4000, 4000, -4000, 448], [4000, 4000, -4000, -4000, 449]), 12, 32),
        ("wide taps, plain shift q8.8", kernel([30000, -2000, 500, -300, 96], scale=3000), 8, 32),
    ]
    for dtype, out_dtypes, shape in ((np.int16, (np.int16, np.uint8), (2, 75, 2000)), (np.uint8, (np.uint8,), (1, 140, 4064))):
        x = signal(rng, shape, dtype, wild=0.02)
        x[..., 10:60, 100:700] = 255                     # saturated block: the extreme partial sums
        x[..., 20:50, 900:1200] = np.where(rng.random((30, 300)) < 0.5, 255, 0).astype(dtype)
        xt = torch.from_numpy(x).cuda()
        for name, hq, frac, acc_bits in kernels:
            for out_dtype in out_dtypes:
                want = c_oracle.fir2d_fixed(x, hq, frac, acc_bits, out_dtype=out_dtype, threads=CORES)
                for rows in ("0", "1", "2"):
                    monkeypatch.setenv("FIRGPU_T2_FFMA_ROWS", rows)
                    got = fg.fir2d_fixed_q(xt, hq, frac, acc_bits, out_dtype=out_dtype)
                    assert _lib.last_path() == _lib.PATH_STAGED, (name, dtype, rows)
                    assert np.array_equal(got.cpu().numpy(), want), (name, dtype, out_dtype, rows)
    monkeypatch.delenv("FIRGPU_T2_FFMA_ROWS")
