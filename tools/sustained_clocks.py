#!/usr/bin/env python
"""Back-to-back launches against one launch taken alone, with SM clock and board power sampled during the sustained
part: why instruction-heavier kernels lose more between an isolated ncu duration and a replayed sequence than the
headline kernel does.  One JSON line per case."""
import json
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "warmup-fir-filter_b200"))
import firgpu  # noqa: E402
from firgpu import quant  # noqa: E402

import pynvml  # noqa: E402

pynvml.nvmlInit()
H = pynvml.nvmlDeviceGetHandleByIndex(0)
LP5 = [0.0625, 0.25, 0.375, 0.25, 0.0625]


class Sampler:
    def __init__(self):
        self.clk, self.pw, self.on, self.stop = [], [], False, False
        self.t = threading.Thread(target=self.run, daemon=True)
        self.t.start()

    def run(self):
        while not self.stop:
            if self.on:
                self.clk.append(pynvml.nvmlDeviceGetClockInfo(H, pynvml.NVML_CLOCK_SM))
                self.pw.append(pynvml.nvmlDeviceGetPowerUsage(H) / 1000.0)
            time.sleep(0.002)


def i16(shape):
    g = torch.Generator(device="cuda").manual_seed(1)
    return torch.randint(0, 256, shape, dtype=torch.int16, device="cuda", generator=g)


def measure(name, fn, nbytes):
    fn()
    one = firgpu.Recording(lambda: fn())
    many = firgpu.Recording(lambda: [fn() for _ in range(50)])
    torch.cuda.synchronize()
    alone = []
    for _ in range(5):                       # one launch after 50 ms of idle
        time.sleep(0.05)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); one.replay(); e1.record()
        torch.cuda.synchronize()
        alone.append(e0.elapsed_time(e1))
    s = Sampler()
    per = []
    t_end = time.time() + 1.5
    s.on = True
    while time.time() < t_end:               # 1.5 s back to back
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); many.replay(); e1.record()
        torch.cuda.synchronize()
        per.append(e0.elapsed_time(e1) / 50)
    s.on = False
    s.stop = True
    peak = 6545.9
    print(json.dumps({"case": name, "alone_ms_min": round(min(alone), 5), "alone_frac": round(nbytes / min(alone) / 1e6 / peak, 3),
                      "sustained_ms_first": round(per[0], 5), "sustained_ms_median": round(float(np.median(per)), 5),
                      "sustained_frac": round(nbytes / float(np.median(per)) / 1e6 / peak, 3),
                      "sm_mhz_median": int(np.median(s.clk)) if s.clk else None, "sm_mhz_min": int(min(s.clk)) if s.clk else None,
                      "power_w_median": round(float(np.median(s.pw)), 1) if s.pw else None, "power_w_max": round(max(s.pw), 1) if s.pw else None}), flush=True)


x = i16((1 << 28,)); y = torch.empty_like(x)
measure("headline 2^28 i16 5tap", lambda: firgpu.fir1d_fixed(x, LP5, 15, 32, 16, out_dtype=np.int16, out=y, validate=False), (1 << 28) * 4)
del x, y
for cols in (256, 4096):
    x = i16(((1 << 27) // cols, cols)); y = torch.empty_like(x)
    measure(f"rows {(1 << 27) // cols}x{cols} i16 5tap (row ends in-kernel)", lambda: firgpu.fir1d_fixed(x, LP5, 15, 32, 16, out_dtype=np.int16, out=y, validate=False), (1 << 27) * 4)
    del x, y
x = i16((1 << 28,)).to(torch.uint8); y = torch.empty_like(x)
measure("u8->u8 2^28 5tap q4.12", lambda: firgpu.fir1d_fixed(x, LP5, 12, 32, 16, out=y, validate=False), (1 << 28) * 2)
del x, y
img = i16((1, 16384, 16384)); out = torch.empty_like(img)
for tag, k in (("3x3", [[1 / 9] * 3] * 3), ("5x5 general", np.random.default_rng(0).uniform(-0.2, 0.2, size=(5, 5)).tolist())):
    hq = quant.quantize_taps(np.asarray(k), 12, 16)
    measure(f"2d 16384^2 i16 {tag}", lambda: firgpu.fir2d_fixed_q(img, hq, 12, 32, out_dtype=np.int16, out=out), 16384 * 16384 * 4)
