#!/usr/bin/env python
"""Float32-output ideal kernel: wide-group variant (FIRGPU_IDEAL_F32_WIDE=2|4|8) against the default form.

One process per setting (the knob is read once); each prints replayed device time for 2^26 uint8 / int16 samples and rows of
the reference's largest image, plus a SHA-256 of every output so the settings can be compared bit for bit.

    python tools/ideal_f32_wide.py            # runs all settings
"""
import hashlib
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "warmup-fir-filter_b200"))


def child():
    import numpy as np
    import torch
    import firgpu
    PEAK = json.load(open(ROOT / "MEASURED_PEAKS.json")).get("hbm_gbs", 6545.9) if (ROOT / "MEASURED_PEAKS.json").exists() else 6545.9
    LP5 = [0.0625, 0.25, 0.375, 0.25, 0.0625]
    g = torch.Generator(device="cuda").manual_seed(7)
    out = {}
    iq = os.environ.get("IDEAL_IQ") == "1"
    cases = (("iq i16 2^26 complex", torch.int16, (1, 1 << 27), LP5), ("iq i16 4096x8192 complex 9tap", torch.int16, (4096, 16384), [0.1] * 9)) if iq else None
    for name, dt, shape, taps in cases or (("u8 2^26", torch.uint8, (1, 1 << 26), LP5), ("i16 2^26", torch.int16, (1, 1 << 26), LP5),
                                  ("u8 2999x4499", torch.uint8, (2999, 4499), LP5), ("u8 8192x16384 3tap", torch.uint8, (8192, 16384), [0.25, 0.5, 0.25]),
                                  ("i16 4096x4104 9tap", torch.int16, (4096, 4104), [0.1, 0.1, 0.1, 0.1, 0.2, 0.1, 0.1, 0.1, 0.1])):
        if dt == torch.uint8:
            x = torch.randint(0, 256, shape, dtype=torch.uint8, device="cuda", generator=g)
        else:
            x = torch.randint(-300, 600, shape, dtype=torch.int16, device="cuda", generator=g)
        odt = np.float64 if os.environ.get("IDEAL_OUT") == "f64" else np.float32
        y = torch.empty(shape, dtype=torch.float64 if odt is np.float64 else torch.float32, device="cuda")
        fn = (lambda: firgpu.fir1d_ideal(x, taps, dilation=2, out_dtype=odt, out=y, validate=False)) if iq else \
             (lambda: firgpu.fir1d_ideal(x, taps, out_dtype=odt, out=y, validate=False))
        fn()
        torch.cuda.synchronize()
        digest = hashlib.sha256(y.cpu().numpy().tobytes()).hexdigest()[:16]
        rec = firgpu.Recording(lambda: [fn() for _ in range(20)])
        for _ in range(3):
            rec.replay()
        torch.cuda.synchronize()
        big = torch.empty(1 << 28, dtype=torch.uint8, device="cuda")
        best = 1e9
        for _ in range(5):
            big.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rec.replay()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 20)
        n = shape[0] * shape[1]
        bps = x.element_size() + y.element_size()
        out[name] = {"ms": round(best, 5), "gsps": round(n / best / 1e6, 1), "frac": round(n * bps / best / 1e6 / PEAK, 3), "sha": digest}
    print(json.dumps({"out": os.environ.get("IDEAL_OUT", "f32"), "wide": os.environ.get("FIRGPU_IDEAL_IQ_WIDE", os.environ.get("FIRGPU_IDEAL_F32_WIDE", os.environ.get("FIRGPU_IDEAL_F64_WIDE", "default"))),
                      "threads": os.environ.get("FIRGPU_IDEAL_THREADS", "default"), "cases": out}))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        runs = [{"IDEAL_IQ": "1", "IDEAL_OUT": o, "FIRGPU_IDEAL_IQ_WIDE": w} for o in ("f64", "f32") for w in ("0", "1", "2")] + \
               [{"FIRGPU_IDEAL_F32_WIDE": "0"}, {}, {"IDEAL_OUT": "f64", "FIRGPU_IDEAL_F64_WIDE": "0"}, {"IDEAL_OUT": "f64"}]
        for extra in runs:
            r = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, **extra), capture_output=True, text=True)
            print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-800:], flush=True)
