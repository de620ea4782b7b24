#!/usr/bin/env python
"""Kernel-level sweep over the shapes BASELINE.json names (not the contract
bench: see bench.py).  Device-resident inputs, CUDA events, one line per case:

    python tools/bench_matrix.py [--cases stream,u8,iq,batched,taps,images,ideal,2d] [--reps 20]
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "warmup-fir-filter_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

import firgpu  # noqa: E402
from firgpu import _lib, quant  # noqa: E402

LP5 = [1 / 16, 4 / 16, 6 / 16, 4 / 16, 1 / 16]
LP3 = [0.25, 0.5, 0.25]
PEAK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0


REPLAY = False      # --replay: time a recorded sequence (firgpu.Recording, a CUDA graph of the same launches) instead of Python calls


def timed(fn, reps, warm=3, recordable=True):
    if REPLAY and recordable:
        rec = firgpu.Recording(fn)
        fn = rec.replay
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3


def report(name, seconds, samples, bytes_per_sample):
    gbs = samples * bytes_per_sample / seconds / 1e9
    path = {1: "generic", 2: "staged", 3: "empty"}.get(_lib.last_path(), "?")
    print(json.dumps({"case": name, "ms": round(seconds * 1e3, 4), "gsamples_s": round(samples / seconds / 1e9, 1),
                      "gb_s": round(gbs, 1), "frac_of_measured_peak": round(gbs / PEAK, 3), "path": path}), flush=True)


def rand_i16(n, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randint(-64, 320, (n,), generator=g, device="cuda", dtype=torch.int16)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="stream,signed,u8,iq,taps,batched,images,ideal,2d")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--replay", action="store_true", help="submit each case as a recorded sequence (device time without per-call Python)")
    args = ap.parse_args()
    global REPLAY
    REPLAY = args.replay
    cases = args.cases.split(",")
    n = 1 << 28
    if "stream" in cases:
        x = rand_i16(n)
        y = torch.empty_like(x)
        for frac, tag in ((15, "q1.15 (hi16 post)"), (12, "q4.12 (shift post)")):
            s = timed(lambda: firgpu.fir1d_fixed(x, LP5, frac, 32, 16, out_dtype=np.int16, out=y, validate=False), args.reps)
            report(f"stream 2^28 i16->i16 5tap {tag}", s, n, 4)
        y8 = torch.empty(n, dtype=torch.uint8, device="cuda")
        s = timed(lambda: firgpu.fir1d_fixed(x, LP5, 15, 32, 16, out=y8, validate=False), args.reps)
        report("stream 2^28 i16->u8 5tap q1.15", s, n, 3)
        del x, y, y8
    if "signed" in cases:
        g = torch.Generator(device="cuda").manual_seed(3)
        x = torch.randint(-32768, 32768, (n,), generator=g, device="cuda", dtype=torch.int32).to(torch.int16)
        y = torch.empty_like(x)
        full = (-32768, 32767)
        s = timed(lambda: firgpu.fir1d_fixed(x, LP5, 15, 32, 16, in_range=full, out_range=full, out=y, validate=False), args.reps)
        report("signed int16 stream 2^28 i16->i16 5tap q1.15 (wide datapath)", s, n, 4)
        s = timed(lambda: firgpu.fir1d_fixed(x, LP5, 15, 32, 16, dilation=2, in_range=full, out_range=full, out=y, validate=False), args.reps)
        report("signed int16 I/Q 2^27 complex 5tap q1.15 (wide datapath)", s, n, 4)
        s = timed(lambda: firgpu.fir1d_fixed(x[: 1 << 26], LP5, 15, 48, 16, centre=0, in_range=full, out_range=full, out=y[: 1 << 26], validate=False), 5)
        report("signed int16 causal 2^26 5tap (centred 9-tap form, wide datapath)", s, 1 << 26, 4)
        del x, y
    if "wrap" in cases:
        # live acc_bits wrap (WRAP post mode) and taps beyond int16 (IMAD datapath) on the staged kernel
        x = rand_i16(n)
        y = torch.empty_like(x)
        hw = [0.9, -0.7, 0.8, 0.3, -0.2]
        for frac, acc, cb, tag in ((12, 16, 16, "acc_bits 16"), (12, 20, 16, "acc_bits 20"), (12, 24, 16, "acc_bits 24"),
                                   (28, 32, 32, "coeff_bits 32, frac 28 (IMAD)")):
            s = timed(lambda: firgpu.fir1d_fixed(x, hw, frac, acc, cb, out_dtype=np.int16, out=y, validate=False), args.reps)
            report(f"stream 2^28 i16->i16 5tap wrap-live {tag}", s, n, 4)
        xu = x.clamp(0, 255).to(torch.uint8)
        yu = torch.empty_like(xu)
        s = timed(lambda: firgpu.fir1d_fixed(xu, hw, 12, 16, 16, out=yu, validate=False), args.reps)
        report("stream 2^28 u8->u8 5tap wrap-live acc_bits 16", s, n, 2)
        del x, y, xu, yu
    if "small" in cases:
        # back-to-back small launches (a strong-scaled shard, a streaming block): 16 passes per recorded sequence, so
        # that programmatic dependent launch between them is visible (compare with FIRGPU_NO_PDL=1)
        for m in (1 << 20, 1 << 22, 1 << 24, 1 << 25):
            x = rand_i16(m)
            y = torch.empty_like(x)
            s = timed(lambda: [firgpu.fir1d_fixed(x, LP5, 15, 32, 16, out_dtype=np.int16, out=y, validate=False) for _ in range(16)], args.reps)
            report(f"16 x stream {m} i16->i16 5tap q1.15 back to back (per pass)", s / 16, m, 4)
            xz = x.view(-1)
            s = timed(lambda: [firgpu.fir1d_fixed(xz, LP5, 15, 32, 16, dilation=2, out_dtype=np.int16, out=y, validate=False) for _ in range(16)], args.reps)
            report(f"16 x iq {m} scalars i16 5tap q1.15 back to back (per pass)", s / 16, m, 4)
            del x, y
    if "streamblocks" in cases:
        # device-resident streaming (firgpu_stream_*): a 2^28-sample int16 signal delivered in blocks, against one call
        from firgpu import quant as _q
        x = rand_i16(n)
        y = torch.empty_like(x)
        hq = _q.quantize_taps(LP5, 15, 16)
        s = timed(lambda: firgpu.fir1d_fixed(x, LP5, 15, 32, 16, centre=0, out_dtype=np.int16, out=y, validate=False), args.reps)
        report("causal one call 2^28 i16->i16 5tap q1.15", s, n, 4)
        for lg in (20, 22, 24, 26, 28):
            blk = 1 << lg
            st = firgpu.FirStream(hq, model="fixed", in_dtype=np.int16, out_dtype=np.int16, frac_bits=15)
            s = timed(lambda: [st.process(x[b:b + blk], out=y[b:b + blk]) for b in range(0, n, blk)], max(3, args.reps // 4))
            report(f"stream object, 2^28 samples in device blocks of 2^{lg} (main + history kernel per block)", s, n, 4)
            st.close()
        g = torch.Generator(device="cuda").manual_seed(9)
        xf = torch.randn(1 << 26, generator=g, device="cuda", dtype=torch.float64)
        yf = torch.empty_like(xf)
        s = timed(lambda: firgpu.fir1d_ideal(xf, LP5, centre=0, raw=True, out=yf, validate=False), 5)
        report("C++ Fir1D model, one call 2^26 f64->f64 5tap", s, 1 << 26, 16)
        for lg in (20, 22, 24):
            blk = 1 << lg
            st = firgpu.FirStream(LP5, model="ideal", in_dtype=np.float64, out_dtype=np.float64, raw=True)
            s = timed(lambda: [st.process(xf[b:b + blk], out=yf[b:b + blk]) for b in range(0, 1 << 26, blk)], 5)
            report(f"C++ Fir1D model as a stream object, 2^26 samples in device blocks of 2^{lg}", s, 1 << 26, 16)
            st.close()
        del x, y, xf, yf
    if "u8" in cases:
        x = rand_i16(n).clamp(0, 255).to(torch.uint8)
        y = torch.empty_like(x)
        for h, tag in ((LP5, "5tap"), (LP3, "3tap")):
            s = timed(lambda: firgpu.fir1d_fixed(x, h, 12, 32, 16, out=y, validate=False), args.reps)
            report(f"stream 2^28 u8->u8 {tag} q4.12", s, n, 2)
        s = timed(lambda: firgpu.fir1d_fixed(x, LP5, 15, 32, 16, out=y, validate=False), args.reps)
        report("stream 2^28 u8->u8 5tap q1.15", s, n, 2)
        del x, y
    if "iq" in cases:
        z = rand_i16(n)                      # 2^27 complex samples interleaved
        y = torch.empty_like(z)
        s = timed(lambda: firgpu.fir1d_fixed(z, LP5, 15, 32, 16, dilation=2, out_dtype=np.int16, out=y, validate=False), args.reps)
        report("iq 2^27 complex i16 5tap q1.15", s, n, 4)
        del z, y
    if "taps" in cases:
        x = rand_i16(n)
        y = torch.empty_like(x)
        for L in (1, 2, 3, 4, 5, 7, 9):
            h = (np.ones(L) / (2 * L)).tolist()
            s = timed(lambda: firgpu.fir1d_fixed(x, h, 15, 32, 16, out_dtype=np.int16, out=y, validate=False), args.reps)
            report(f"stream 2^28 i16->i16 {L}tap q1.15", s, n, 4)
        for L in (13, 17, 33, 65, 97, 129):
            hl = (np.ones(L) / (2 * L)).tolist()
            s = timed(lambda: firgpu.fir1d_fixed(x, hl, 15, 32, 16, out_dtype=np.int16, out=y, validate=False), args.reps)
            report(f"stream 2^28 i16->i16 {L}tap q1.15 (long-filter bucket)", s, n, 4)
        h130 = (np.ones(130) / 260).tolist()
        s = timed(lambda: firgpu.fir1d_fixed(x[: 1 << 26], h130, 15, 32, 16, out_dtype=np.int16, out=y[: 1 << 26], validate=False), 5)
        report("stream 2^26 i16->i16 130tap (generic kernel)", s, 1 << 26, 4)
        del x, y
    if "batched" in cases:
        for rows, cols in ((8192, 1 << 20), (8192, 1_000_000), (1 << 16, 4096), (1 << 20, 256)):
            x = rand_i16(rows * cols).view(rows, cols)
            y = torch.empty_like(x)
            s = timed(lambda: firgpu.fir1d_fixed(x, LP5, 15, 32, 16, out_dtype=np.int16, out=y, validate=False), max(3, args.reps // 4))
            report(f"batched {rows}x{cols} i16->i16 5tap q1.15", s, rows * cols, 4)
            del x, y
    if "images" in cases:
        for hgt, wid in ((853, 1280), (2999, 4499), (64, 64)):
            x = rand_i16(hgt * wid).clamp(0, 255).to(torch.uint8).view(hgt, wid)
            y = torch.empty_like(x)
            s = timed(lambda: firgpu.fir1d_fixed(x, LP5, 12, 32, 16, out=y, validate=False), args.reps)
            report(f"image rows {hgt}x{wid} u8->u8 5tap q4.12", s, hgt * wid, 2)
    if "ideal" in cases:
        x = rand_i16(1 << 26).clamp(0, 255).to(torch.uint8)
        y = torch.empty(1 << 26, dtype=torch.float64, device="cuda")
        s = timed(lambda: firgpu.fir1d_ideal(x, LP5, out=y, validate=False), 5)
        report("ideal 2^26 u8->f64 5tap", s, 1 << 26, 9)
        y32 = torch.empty(1 << 26, dtype=torch.float32, device="cuda")
        s = timed(lambda: firgpu.fir1d_ideal(x, LP5, out_dtype=np.float32, out=y32, validate=False), 5)
        report("ideal 2^26 u8->f32 5tap", s, 1 << 26, 5)
        del x, y, y32
        # the reference's image shapes (rows are filtered independently: every row has two edge tiles)
        for hgt, wid in ((2999, 4499), (1080, 1920), (8192, 16384)):
            x = rand_i16(hgt * wid).clamp(0, 255).to(torch.uint8).view(hgt, wid)
            y = torch.empty(hgt, wid, dtype=torch.float64, device="cuda")
            s = timed(lambda: firgpu.fir1d_ideal(x, LP5, out=y, validate=False), 10)
            report(f"ideal rows {hgt}x{wid} u8->f64 5tap", s, hgt * wid, 9)
            del x, y
    if "2d" in cases:
        img = rand_i16(16384 * 16384).view(1, 16384, 16384)
        out = torch.empty_like(img)
        k3 = [[1 / 9] * 3] * 3
        k5 = np.outer(LP5, LP5).tolist()
        kr = np.random.default_rng(0).uniform(-0.2, 0.2, size=(5, 5)).tolist()
        for k, tag in ((k3, "3x3 box"), (k5, "5x5 lp rank-1"), (kr, "5x5 general")):
            hq = quant.quantize_taps(np.asarray(k), 12, 16)
            s = timed(lambda: firgpu.fir2d_fixed_q(img, hq, 12, 32, out_dtype=np.int16, out=out), args.reps, warm=2)
            report(f"2d 16384^2 i16->i16 {tag} q4.12", s, 16384 * 16384, 4)
        img8 = rand_i16(16384 * 16384).clamp(0, 255).to(torch.uint8).view(1, 16384, 16384)
        out8 = torch.empty_like(img8)
        for k, tag in ((k3, "3x3 box"), (k5, "5x5 lp rank-1"), (kr, "5x5 general")):
            hq = quant.quantize_taps(np.asarray(k), 12, 16)
            s = timed(lambda: firgpu.fir2d_fixed_q(img8, hq, 12, 32, out=out8), args.reps, warm=2)
            report(f"2d 16384^2 u8->u8 {tag} q4.12", s, 16384 * 16384, 2)
        del img, out, img8, out8
        small8 = rand_i16(8192 * 8192).clamp(0, 255).to(torch.uint8).view(1, 8192, 8192)
        outf = torch.empty((1, 8192, 8192), dtype=torch.float64, device="cuda")
        s = timed(lambda: firgpu.fir2d_ideal(small8, k3, out=outf, validate=False), 3, warm=1)
        report("2d ideal 8192^2 u8->f64 3x3 (generic kernel)", s, 8192 * 8192, 9)
        del small8, outf
        batch = rand_i16(4 * 8192 * 8192).view(4, 8192, 8192)
        outb = torch.empty_like(batch)
        hq = quant.quantize_taps(np.asarray(k3), 12, 16)
        s = timed(lambda: firgpu.fir2d_fixed_q(batch, hq, 12, 32, out_dtype=np.int16, out=outb), args.reps, warm=2)
        report("2d batch 4 x 8192^2 i16->i16 3x3 box q4.12", s, 4 * 8192 * 8192, 4)
    if "multi" in cases:
        from fir_1d.sim.vector.h_coeff import h_coeff_5tap_map
        sets = list(h_coeff_5tap_map.values())
        for hgt, wid in ((2999, 4499), (8192, 16384)):
            x = rand_i16(hgt * wid).clamp(0, 255).to(torch.uint8).view(hgt, wid)
            s = timed(lambda: firgpu.fir1d_fixed_multi(x, sets, validate=False), args.reps)
            report(f"multi-filter x4 {hgt}x{wid} u8 -> 4 x u8 5tap q4.12 (1 read + 4 writes)", s, hgt * wid, 5)
            s = timed(lambda: [firgpu.fir1d_fixed(x, h, validate=False) for h in sets], args.reps)
            report(f"4 single passes {hgt}x{wid} u8 -> u8 5tap q4.12 (same 5 B/px accounting)", s, hgt * wid, 5)
    if "2dodd" in cases:
        # the reference's largest image (odd width and pitch): re-pitched copy + tiled kernel vs the generic kernel
        img = rand_i16(2999 * 4499).clamp(0, 255).to(torch.uint8).view(2999, 4499)
        out = torch.empty_like(img)
        for k, tag in (([[1 / 9] * 3] * 3, "3x3 box"), (np.outer(LP5, LP5).tolist(), "5x5 lp")):
            hq = quant.quantize_taps(np.asarray(k), 12, 16)
            s = timed(lambda: firgpu.fir2d_fixed_q(img, hq, 12, 32, out=out), args.reps)
            report(f"2d 2999x4499 u8->u8 {tag} q4.12 (odd pitch)", s, 2999 * 4499, 2)
    if "longtaps" in cases:
        # the long-filter buckets alone (the last session of round 2 added 97 / 129 taps and uint8 in / int16 out)
        x = rand_i16(1 << 27)
        y = torch.empty_like(x)
        x8 = x.clamp(0, 255).to(torch.uint8)
        for L in (65, 97, 129):
            hl = (np.ones(L) / (2 * L)).tolist()
            s = timed(lambda: firgpu.fir1d_fixed(x, hl, 15, 32, 16, out_dtype=np.int16, out=y, validate=False), args.reps)
            report(f"stream 2^27 i16->i16 {L}tap q1.15 (long-filter bucket)", s, 1 << 27, 4)
        for L in (21, 65):
            hl = (np.ones(L) / (2 * L)).tolist()
            s = timed(lambda: firgpu.fir1d_fixed(x8, hl, 15, 32, 16, out_dtype=np.int16, out=y, validate=False), args.reps)
            report(f"stream 2^27 u8->i16 {L}tap q1.15 (long-filter bucket)", s, 1 << 27, 3)
        h130 = (np.ones(130) / 260).tolist()
        s = timed(lambda: firgpu.fir1d_fixed(x[: 1 << 24], h130, 15, 32, 16, out_dtype=np.int16, out=y[: 1 << 24], validate=False), 3)
        report("stream 2^24 i16->i16 130tap (generic kernel)", s, 1 << 24, 4)
        del x, y, x8
    if "2dwin" in cases:
        # windows beyond 5x5 and even sizes on the tensor-map kernel (FIRGPU_FORCE_GENERIC=1 in the environment times the
        # one-thread-per-pixel int64 kernel they ran on before)
        img = rand_i16(8192 * 8192).view(1, 8192, 8192)
        out = torch.empty_like(img)
        g7 = np.array([1, 6, 15, 20, 15, 6, 1]) / 64.0
        rng7 = np.random.default_rng(7)
        for hq, tag in ((np.outer([1, 6, 15, 20, 15, 6, 1], [1, 6, 15, 20, 15, 6, 1]), "7x7 binomial rank-1"),
                        (quant.quantize_taps(rng7.uniform(-0.1, 0.1, size=(7, 7)), 12, 16), "7x7 general"),
                        (quant.quantize_taps(rng7.uniform(-0.2, 0.2, size=(3, 7)), 12, 16), "3x7 general"),
                        (quant.quantize_taps(rng7.uniform(-0.2, 0.2, size=(7, 3)), 12, 16), "7x3 general"),
                        (quant.quantize_taps(np.outer(g7, [0.25, 0.5, 0.25]), 12, 16), "7x3 as quantised"),
                        (quant.quantize_taps(rng7.uniform(-0.2, 0.2, size=(4, 4)), 12, 16), "4x4 general (zero-padded to 5x5)"),
                        (quant.quantize_taps(rng7.uniform(-0.3, 0.3, size=(2, 2)), 12, 16), "2x2 general (zero-padded to 3x3)")):
            hq = np.ascontiguousarray(hq, dtype=np.int32)
            s = timed(lambda: firgpu.fir2d_fixed_q(img, hq, 12, 32, out_dtype=np.int16, out=out), args.reps, warm=2)
            report(f"2d 8192^2 i16->i16 {tag} q4.12", s, 8192 * 8192, 4)
        del img, out
    if "compare" in cases:
        # fused fixed-vs-ideal metrics: 1 B read per sample, seven numbers out (the call synchronises to return them)
        for hgt, wid in ((64, 64), (2999, 4499), (8192, 16384)):
            x = rand_i16(hgt * wid).clamp(0, 255).to(torch.uint8).view(hgt, wid)
            s = timed(lambda: firgpu.compare_metrics(x, LP5, validate=False), 5, recordable=False)    # (the call synchronises to return its numbers)
            report(f"compare metrics {hgt}x{wid} u8 5tap q4.12 (fixed + ideal + 7 reductions)", s, hgt * wid, 1)
    if "iqideal" in cases:
        z = rand_i16(1 << 27)                      # 2^26 complex samples interleaved
        y = torch.empty(1 << 27, dtype=torch.float64, device="cuda")
        s = timed(lambda: firgpu.fir1d_ideal(z, LP5, dilation=2, out=y, validate=False), 5)
        report("iq ideal 2^26 complex i16->f64 5tap", s, 1 << 27, 10)
        del z, y
    if "narrow" in cases:
        # batches of short aligned rows: row-streaming tensor-map kernel vs flattened pass + row-end fix-up
        for cols in (64, 128, 256, 512, 1024, 2048, 4096):
            rows = (1 << 27) // cols
            x = rand_i16(rows * cols).view(rows, cols)
            y = torch.empty_like(x)
            s = timed(lambda: firgpu.fir1d_fixed(x, LP5, 15, 32, 16, out_dtype=np.int16, out=y, validate=False), args.reps)
            report(f"batched {rows}x{cols} i16->i16 5tap q1.15", s, rows * cols, 4)
            del x, y
    if "2dmixed" in cases:
        img = rand_i16(16384 * 16384).view(16384, 16384)
        out = torch.empty(16384, 16384, dtype=torch.uint8, device="cuda")
        for k, tag in (([[1 / 9] * 3] * 3, "3x3 box"), (np.outer(LP5, LP5).tolist(), "5x5 lp")):
            hq = quant.quantize_taps(np.asarray(k), 12, 16)
            s = timed(lambda: firgpu.fir2d_fixed_q(img, hq, 12, 32, out=out), args.reps)
            report(f"2d 16384^2 i16->u8 {tag} q4.12 (int16 in, uint8 out: the drop-in default)", s, 16384 * 16384, 3)
    if "causal" in cases:
        g = torch.Generator(device="cuda").manual_seed(9)
        x = torch.randn(1 << 26, generator=g, device="cuda", dtype=torch.float64)
        y = torch.empty_like(x)
        s = timed(lambda: firgpu.fir1d_ideal(x, LP5, centre=0, raw=True, out=y, validate=False), 5)
        report("causal raw f64->f64 2^26 5tap (C++ Fir1D model)", s, 1 << 26, 16)
        del x, y
    if "ideal2d" in cases:
        img = rand_i16(8192 * 8192).clamp(0, 255).to(torch.uint8).view(8192, 8192)
        k5 = np.outer(LP5, LP5).tolist()
        for k, tag in (([[1 / 9] * 3] * 3, "3x3"), (k5, "5x5")):
            for odt, tdt, nb in ((np.float64, torch.float64, 9), (np.float32, torch.float32, 5)):
                out = torch.empty(8192, 8192, dtype=tdt, device="cuda")
                s = timed(lambda: firgpu.fir2d_ideal(img, k, out_dtype=odt, out=out, validate=False), 5)
                report(f"2d ideal 8192^2 u8->{np.dtype(odt).name} {tag}", s, 8192 * 8192, nb)
                del out
    if "containers" in cases:
        # int32 / float sample containers: rounding pass to uint8 + staged kernel.  Algorithmic bytes per
        # sample = container read + 1 output byte; the uint8 temporary (1 write + 1 read) is extra traffic.
        m = 1 << 27
        g = torch.Generator(device="cuda").manual_seed(5)
        base = torch.rand(m, generator=g, device="cuda", dtype=torch.float32) * 300.0 - 20.0
        y = torch.empty(m, dtype=torch.uint8, device="cuda")
        for dt, nb in ((torch.float32, 4), (torch.float64, 8), (torch.int32, 4)):
            x = base.to(dt)
            s = timed(lambda: firgpu.fir1d_fixed(x, LP5, 15, 32, 16, out=y, validate=False), args.reps)
            report(f"stream 2^27 {str(dt).split('.')[1]}->u8 5tap q1.15 (prep + staged)", s, m, nb + 1)
            del x
        k3 = [[1 / 9] * 3] * 3
        img = (torch.rand(8192 * 8192, generator=g, device="cuda", dtype=torch.float32) * 300.0 - 20.0).view(8192, 8192)
        out = torch.empty(8192, 8192, dtype=torch.uint8, device="cuda")
        s = timed(lambda: firgpu.fir2d_fixed(img, k3, 12, 32, 16, out=out, validate=False), args.reps)
        report("2d 8192^2 float32->u8 3x3 box q4.12 (prep + tiled)", s, 8192 * 8192, 5)


if __name__ == "__main__":
    main()
