#!/usr/bin/env python
"""Row-band shard of BASELINE config 4 at N GPUs on ONE GPU: [4, 16384 / N, 16384] int16, by stages per item
(FIRGPU_T2_SPI, read per launch).  python tools/shard2d_sweep.py N"""
import json, os, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "warmup-fir-filter_b200")]
import firgpu
from firgpu import quant
n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 8
hgt = 16384 // n_gpus
x = torch.randint(0, 256, (4, hgt, 16384), device="cuda", dtype=torch.int16)
y = torch.empty_like(x)
LP = [1 / 16, 4 / 16, 6 / 16, 4 / 16, 1 / 16]
kernels = {"3x3": np.full((3, 3), 1 / 9), "5x5 rank-1": np.outer(LP, LP), "5x5 general": np.random.default_rng(0).uniform(-0.2, 0.2, size=(5, 5))}
for name, k in kernels.items():
    hq = quant.quantize_taps(k, 12, 16)
    for spi in [0] + list(range(2, 15)):
        if spi:
            os.environ["FIRGPU_T2_SPI"] = str(spi)
        else:
            os.environ.pop("FIRGPU_T2_SPI", None)
        fn = lambda: firgpu.fir2d_fixed_q(x, hq, 12, 32, out_dtype=np.int16, out=y)
        rec = firgpu.Recording(lambda: [fn() for _ in range(10)])
        rec.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); rec.replay(); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        print(json.dumps({"shard_of": n_gpus, "kernel": name, "stages_per_item": spi or "auto", "ms": round(ms, 4),
                          "frac": round(x.numel() * 4 / (ms * 1e-3) / 1e9 / 6545.9, 3)}), flush=True)
