#!/usr/bin/env python
"""uint8 samples -> int16 container (3 B/sample): replayed device time of the staged kernel, 2^28 samples, 5 / 3 taps."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "warmup-fir-filter_b200"))
import firgpu  # noqa: E402

PEAK = json.load(open(ROOT / "MEASURED_PEAKS.json")).get("hbm_gbs", 6545.9) if (ROOT / "MEASURED_PEAKS.json").exists() else 6545.9
n = 1 << 28
x = torch.randint(0, 256, (n,), dtype=torch.uint8, device="cuda")
y = torch.empty(n, dtype=torch.int16, device="cuda")
for taps in ([0.0625, 0.25, 0.375, 0.25, 0.0625], [0.25, 0.5, 0.25]):
    fn = lambda: firgpu.fir1d_fixed(x, taps, out_dtype=np.int16, out=y, validate=False)
    fn()
    rec = firgpu.Recording(lambda: [fn() for _ in range(10)])
    rec.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rec.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 10)
    print(json.dumps({"case": f"u8->i16 2^28 {len(taps)}tap q4.12", "ms": round(best, 5), "gsamples_s": round(n / best / 1e6, 1),
                      "frac_of_measured_peak": round(n * 3 / best / 1e6 / PEAK, 3)}))
