#!/usr/bin/env python
"""BASELINE configs 3, 4, 5 on N GPUs (the contract bench, bench.py, measures
config 2).  Same timing rules: one process per GPU, barrier + synchronize on
both sides, CUDA events on the launching stream, max over ranks.  STRONG
scaling: the global problem is fixed and sharded with no data-path collective.

    python tests/perf/bench_configs.py                                   # 1 GPU
    python -m torch.distributed.run --nproc-per-node N ... tests/perf/bench_configs.py
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[2]
sys.path[:0] = [str(ROOT), str(ROOT / "warmup-fir-filter_b200")]
import firgpu  # noqa: E402
from firgpu import _lib, quant, shard  # noqa: E402
from oracle import c_oracle  # noqa: E402

LP5 = [1 / 16, 4 / 16, 6 / 16, 4 / 16, 1 / 16]
PEAK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--rows", type=int, default=8192)
    args = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    def timed(fn):
        for _ in range(3):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b) / args.reps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) * 1e-3

    def report(name, seconds, units, bytes_per_unit, verified):
        if rank == 0:
            gbs = units * bytes_per_unit / seconds / 1e9
            print(json.dumps({"config": name, "n_gpus": world, "scaling": "strong", "ms": round(seconds * 1e3, 4),
                              "gunits_s": round(units / seconds / 1e9, 1), "gb_s_total": round(gbs, 1),
                              "frac_of_measured_peak_per_gpu": round(gbs / world / PEAK, 3), "verified_vs_oracle": verified}), flush=True)

    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    hq5 = quant.quantize_taps(LP5, 15, 16)

    # ---- config 3: 8192 channels x 2^20 samples, channel-sharded, no halo --------------
    rows, cols = args.rows, 1 << 20
    mine = shard.plan_rows(rows, world)[rank]
    x = torch.randint(-64, 320, (mine.size, cols), generator=g, device=dev, dtype=torch.int16)
    y = torch.empty_like(x)
    fn = lambda: firgpu.fir1d_fixed_q(x, hq5, 15, 32, out_dtype=np.int16, out=y)
    fn()
    want = c_oracle.fir1d_fixed(x[:2, :70000].cpu().numpy(), hq5, 15, 32, out_dtype=np.int16)
    ok = bool(np.array_equal(y[:2, :69998].cpu().numpy(), want[:, :69998]))
    report(f"config 3: batched fir_1d 5-tap int16, {rows} channels x 2^20, channel-sharded", timed(fn), rows * cols, 4, ok)
    del x, y

    # ---- config 5: interleaved I/Q 2^27 complex samples, contiguous segments + halo ----
    n = 1 << 28
    pieces = shard.plan_segments(n, world, 5, 2, align=4096)
    me = pieces[rank]
    lead = shard.aligned_lead(me.before, 2)
    buf = torch.zeros(lead + me.size + me.after, dtype=torch.int16, device=dev)
    buf[lead:lead + me.size] = torch.randint(-64, 320, (me.size,), generator=g, device=dev, dtype=torch.int16)
    if world > 1:
        shard.exchange_halos(buf, lead, pieces, rank, dist)
    body = buf[lead:lead + me.size]
    y = torch.empty(me.size, dtype=torch.int16, device=dev)
    fn = lambda: firgpu.fir1d_fixed_q(body, hq5, 15, 32, dilation=2, out_dtype=np.int16, out=y,
                                      halo_left_valid=me.before, halo_right_valid=me.after)
    fn()
    win = buf[lead - me.before:lead + 40000].cpu().numpy()
    want = c_oracle.fir1d_fixed(win, hq5, 15, 32, dilation=2, out_dtype=np.int16)
    ok = bool(np.array_equal(y[:30000].cpu().numpy(), want[me.before:me.before + 30000]))
    report("config 5: complex I/Q 5-tap int16, 2^27 complex samples, segments with halo", timed(fn), n, 4, ok)
    del buf, y

    # ---- config 4: [4, 16384, 16384] int16 images, row-band sharded with halo ----------
    b_img, hgt, wid = 4, 16384, 16384
    rng = np.random.default_rng(0)
    for name, k in (("3x3 box", [[1 / 9] * 3] * 3), ("5x5 lp (rank-1)", np.outer(LP5, LP5).tolist()),
                    ("5x5 random (non-separable)", rng.uniform(-0.2, 0.2, size=(5, 5)).tolist())):
        kh = len(k)
        bands = shard.plan_bands(hgt, world, kh)
        me = bands[rank]
        buf = torch.zeros((b_img, me.before + me.size + me.after, wid), dtype=torch.int16, device=dev)
        buf[:, me.before:me.before + me.size, :] = torch.randint(-64, 320, (b_img, me.size, wid), generator=g, device=dev, dtype=torch.int16)
        if world > 1:
            shard.exchange_band_halos(buf, bands, rank, dist)
        body = buf[:, me.before:me.before + me.size, :]
        out = torch.empty((b_img, me.size, wid), dtype=torch.int16, device=dev)
        hq = quant.quantize_taps(np.asarray(k), 12, 16)
        fn = lambda: firgpu.fir2d_fixed_q(body, hq, 12, 32, out_dtype=np.int16, out=out,
                                          halo_top_valid=me.before, halo_bottom_valid=me.after)
        fn()
        assert _lib.last_path() == _lib.PATH_STAGED
        crop = buf[1, :me.before + 300, 5000:5400].cpu().numpy()
        want = c_oracle.fir2d_fixed(crop, hq, 12, 32, out_dtype=np.int16, threads=4)
        r = kh // 2
        # interior of the crop (away from the crop's own zero padding), rows measured from the band start
        top = 0 if me.before == 0 and rank == 0 else r
        ok = bool(np.array_equal(out[1, top:250, 5000 + r:5400 - r].cpu().numpy(),
                                 want[me.before + top:me.before + 250, r:400 - r]))
        report(f"config 4: fir_2d {name} int16, {b_img} x {hgt} x {wid}, row-band sharded with halo", timed(fn), b_img * hgt * wid, 4, ok)
        del buf, out, body
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
