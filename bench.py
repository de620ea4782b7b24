#!/usr/bin/env python
"""Benchmark of the fixed-point FIR hot path (BASELINE.json metric: output
Gsamples/s of the 5-tap Q1.15 int16 FIR, and its fraction of the HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one rank's batch of synthetic input:
a contiguous 2**28-sample segment (BASELINE config 2) of a stream of N * 2**28
int16 samples, cut with a duplicated taps-1 halo copied at shard time (weak
scaling, no collective in the timed region).  Inputs (512 MiB per rank) are
larger than the 126 MB L2, so every step streams from HBM.

Printed JSON (one line, rank 0): the base contract plus ``roofline``,
``cpu_baseline``, ``e2e``, ``clocks``, ``gpu_launches`` and ``configs`` — BASELINE
configs 3, 4 and 5 measured in the same run, STRONG-scaled over the N ranks
(fixed global problem, sharded as SURVEY 8(e) says: channels / segments with
halo / row bands with halo, no collective in any timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
PKG = ROOT / "warmup-fir-filter_b200"
for _p in (str(ROOT), str(PKG)):
    if _p not in sys.path:
        sys.path.insert(0, _p)

SEGMENT = 1 << 28                      # samples per rank (BASELINE config 2)
TAPS = [1 / 16, 4 / 16, 6 / 16, 4 / 16, 1 / 16]      # h_coeff_5tap_map["simple_lp"]
FRAC_BITS, ACC_BITS, COEFF_BITS = 15, 32, 16         # Q1.15
BYTES_PER_SAMPLE = 4                   # int16 in + int16 out
METRIC = "output Gsamples/s, 5-tap Q15 FIR"
UNIT = "Gsamples/s"


def workload_name(world: int) -> str:
    return (f"fir_1d 5-tap Q1.15 int16 FIR, synthetic stream of {world} x 2^28 samples, "
            f"contiguous segments with taps-1 halo, 1 segment per B200")


# ----------------------------------------------------------------------------
# clocks: NVML polled from a thread during the timed regions
# ----------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int):
        self.samples, self.reason_bits, self.max_mhz = [], 0, None
        self._stop = threading.Event()
        self._active = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        except Exception as exc:          # no NVML: report that instead of inventing numbers
            self.error = repr(exc)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            if self._active.is_set():
                try:
                    self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                    self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    pass
            time.sleep(0.002)

    def region(self, on: bool):
        (self._active.set if on else self._active.clear)()

    def summary(self) -> dict:
        self._stop.set()
        if self._thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": getattr(self, "error", "nvml unavailable")}
        self._thread.join(timeout=1)
        bits = self.reason_bits & ~0x1
        reasons = [name for bit, name in self.REASONS.items() if bits & bit]
        med = int(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(self.samples)}


# ----------------------------------------------------------------------------
# CPU baselines (oracle/ is executed here only as the reported baseline)
# ----------------------------------------------------------------------------
def _py_port_chunk(args):
    from oracle import fir_oracle_py as opy
    seed, n = args
    rng = np.random.default_rng(seed)
    x = rng.integers(0, 256, size=n).astype(np.int16).tolist()
    t0 = time.perf_counter()
    opy.fir1d_fixed(x, TAPS, FRAC_BITS, ACC_BITS, COEFF_BITS)
    return time.perf_counter() - t0


def python_port_rate(cores: int, per_core: int, pool=None) -> tuple[float, float]:
    """The reference's own algorithm shape (pure-Python per-sample loop,
    fir_1d_fixed_ref.py:95-128) as restated in oracle/fir_oracle_py.py, one
    independent chunk per process.  Returns (samples/s aggregate, seconds)."""
    import multiprocessing as mp
    own = pool is None
    if own:
        pool = mp.get_context("spawn").Pool(cores)
    try:
        t0 = time.perf_counter()
        inner = pool.map(_py_port_chunk, [(s, per_core) for s in range(cores)])
        dt = time.perf_counter() - t0
        if own:                       # a fresh pool pays interpreter start-up inside the wall time
            dt = max(inner)
    finally:
        if own:
            pool.close()
            pool.join()
    return cores * per_core / dt, dt


def c_port_rate(threads: int, n: int, min_seconds: float) -> tuple[float, float, int]:
    from oracle import c_oracle, fir_oracle_np as onp
    rng = np.random.default_rng(1)
    x = rng.integers(0, 256, size=n).astype(np.int16)
    hq = onp.quantize_taps(TAPS, FRAC_BITS, COEFF_BITS)
    c_oracle.fir1d_fixed(x[: 1 << 20], hq, FRAC_BITS, ACC_BITS, out_dtype=np.int16, threads=threads)
    reps, t0 = 0, time.perf_counter()
    while True:
        c_oracle.fir1d_fixed(x, hq, FRAC_BITS, ACC_BITS, out_dtype=np.int16, threads=threads)
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds:
            return reps * n / dt, dt, reps


def numpy_port_rate(n: int, min_seconds: float) -> tuple[float, float, int]:
    """The vectorised NumPy restatement (oracle/fir_oracle_np.py) on one core: SURVEY 8(d) asks for it as a second,
    clearly labelled CPU line next to the scalar port."""
    from oracle import fir_oracle_np as onp
    rng = np.random.default_rng(2)
    x = rng.integers(0, 256, size=n).astype(np.int16)
    onp.fir1d_fixed(x[: 1 << 16], TAPS, FRAC_BITS, ACC_BITS, COEFF_BITS, out_dtype=np.int16)
    reps, t0 = 0, time.perf_counter()
    while True:
        onp.fir1d_fixed(x, TAPS, FRAC_BITS, ACC_BITS, COEFF_BITS, out_dtype=np.int16)
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds:
            return reps * n / dt, dt, reps


def host_cores() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_baseline_block(seconds_c: float = 5.0) -> dict:
    cores = host_cores()
    per_core = 1 << 21
    rate_py, dt_py = python_port_rate(cores, per_core)
    rate_c, dt_c, reps = c_port_rate(cores, 1 << 26, seconds_c)
    rate_np, dt_np, reps_np = numpy_port_rate(1 << 24, 3.0)
    return {
        "value": rate_py / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": (f"oracle/fir_oracle_py.py (pure-Python per-sample loop, the reference's algorithm shape) on "
                   f"{cores} processes x 2^21 samples of the same 5-tap Q1.15 int16 workload, {dt_py:.1f} s"),
        "compiled_port": {"value": rate_c / 1e9, "unit": UNIT, "cores": cores,
                          "sample": f"oracle/fir_oracle.c (gcc -O2 -fopenmp) {reps} x 2^26 samples, {dt_c:.1f} s"},
        "numpy_port": {"value": rate_np / 1e9, "unit": UNIT, "cores": 1,
                       "sample": f"oracle/fir_oracle_np.py (vectorised int64 shift-and-add) {reps_np} x 2^24 samples, {dt_np:.1f} s"},
    }


def reference_arm(args, rank: int) -> None:
    """--impl reference: the reference's CPU implementation of the path.  The
    reference is pure Python and cannot travel to the GPU box, so this times
    its faithful port (oracle/fir_oracle_py.py) on every host core."""
    if rank != 0:
        return
    import multiprocessing as mp
    cores = host_cores()
    per_core = 1 << 17
    pool = mp.get_context("spawn").Pool(cores)
    try:
        for _ in range(args.warmup):
            python_port_rate(cores, 1 << 12, pool)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            python_port_rate(cores, per_core, pool)
        dt = time.perf_counter() - t0
    finally:
        pool.close()
        pool.join()
    value = args.steps * cores * per_core / dt / 1e9
    sample = f"{cores} processes x 2^17 samples per step of the 5-tap Q1.15 int16 stream (bounded sample of the 2^28 workload)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32 accumulate over u8 x int16 (Python ints)", "data": "synthetic",
        "config": {"workload": workload_name(args.gpus), "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------
def make_segment(torch, dist, rank: int, world: int, device):
    """This rank's segment of the global stream in HBM: [lead pad | halo before
    | body | halo after]; the halos are the neighbours' edge samples, exchanged
    once at shard time (outside every timed region)."""
    from firgpu import shard
    pieces = shard.plan_segments(SEGMENT * world, world, len(TAPS), 1, align=4096)
    piece = pieces[rank]
    lead = shard.aligned_lead(piece.before, 2)
    buf = torch.zeros(lead + piece.size + piece.after, dtype=torch.int16, device=device)
    body = buf[lead:lead + piece.size]
    g = torch.Generator(device=device).manual_seed(1234 + rank)
    body.copy_(torch.randint(0, 256, (piece.size,), generator=g, device=device, dtype=torch.int16))
    n_wild = piece.size // 100                # 1 % anywhere in int16: exercises the input clamp
    where = torch.randint(0, piece.size, (n_wild,), generator=g, device=device)
    body[where] = torch.randint(-32768, 32768, (n_wild,), generator=g, device=device, dtype=torch.int32).to(torch.int16)
    del where
    if world > 1:
        shard.exchange_halos(buf, lead, pieces, rank, dist)
    torch.cuda.synchronize()
    return piece, lead, buf, body


# ----------------------------------------------------------------------------
# BASELINE configs 3, 4, 5: strong scaling at the N the driver passes
# ----------------------------------------------------------------------------
def device_signal(torch, shape, seed, device):
    """SURVEY 8(d) input recipe on the device: 99 % pixels in [0,255], 1 % anywhere in int16."""
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.randint(0, 256, shape, generator=g, device=device, dtype=torch.int16)
    n = x.numel()
    where = torch.randint(0, n, (n // 100,), generator=g, device=device)
    x.view(-1)[where] = torch.randint(-32768, 32768, (n // 100,), generator=g, device=device, dtype=torch.int32).to(torch.int16)
    return x


def other_configs(args, torch, dist, rank: int, world: int, device, peak: float) -> dict:
    """One step = one pass over this rank's shard.  The `reps` timed steps are recorded once (firgpu.Recording: the
    library's batch entry, a CUDA graph of exactly the launches the public calls make) and submitted together, so
    that a 20-microsecond kernel is not timed behind 25 microseconds of Python per call.  Timing: barrier +
    synchronize, CUDA events on the launching stream around the submission, max over ranks.  `consistent` is a
    self-consistency check between two different kernel paths of the library (parity against the oracle is the job
    of tests/, which cover these configs at full size)."""
    import firgpu
    from firgpu import _lib, quant, shard

    reps = max(5, min(args.steps, 20))
    out = {}

    def timed(step):
        def steps():
            for _ in range(reps):
                step()
        rec = firgpu.Recording(steps)          # runs `steps` once eagerly (warm-up), then records it
        rec.replay()                           # second warm-up pass, as a replay
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rec.replay()
        b.record()
        torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b) / reps], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), rec.kernels // reps

    def all_ok(flag: bool) -> bool:
        t = torch.tensor([1 if flag else 0], device=device, dtype=torch.int32)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    def entry(name, ms, kernels, units, unit_name, ok, how, partition):
        gbs = units * BYTES_PER_SAMPLE / (ms * 1e-3) / 1e9
        out[name] = {"ms": ms, "value": units / (ms * 1e-3) / 1e9, "unit": unit_name, "scaling": "strong", "n_gpus": world,
                     "roofline": {"bound": "hbm", "achieved_per_gpu": gbs / world, "peak": peak, "unit": "GB/s",
                                  "frac": gbs / world / peak, "bytes_per_unit": BYTES_PER_SAMPLE},
                     "kernels_per_step": kernels, "steps": reps, "partition": partition, "consistent": ok, "consistency_check": how}

    hq5 = quant.quantize_taps(TAPS, FRAC_BITS, COEFF_BITS)

    # ---- config 3: 8192 channels x 2^20 samples, channel-sharded, no halo ----------------------------
    rows, cols = 8192, 1 << 20
    mine = shard.plan_rows(rows, world)[rank]
    need = 2 * mine.size * cols * 2 + (2 << 30)
    if torch.cuda.mem_get_info(device)[0] > need:
        x = torch.empty((mine.size, cols), dtype=torch.int16, device=device)
        for b in range(0, mine.size, 512):
            e = min(b + 512, mine.size)
            x[b:e] = device_signal(torch, (e - b, cols), 1234 + (mine.start + b) // 512, device)
        y = torch.empty_like(x)
        ms, kernels = timed(lambda: firgpu.fir1d_fixed_q(x, hq5, FRAC_BITS, ACC_BITS, out_dtype=np.int16, out=y))
        assert _lib.last_path() == _lib.PATH_STAGED
        r = mine.size // 2                      # a channel filtered alone (single-row tile walk) equals its row of the batch
        ok = torch.equal(firgpu.fir1d_fixed_q(x[r], hq5, FRAC_BITS, ACC_BITS, out_dtype=np.int16), y[r])
        entry("config3", ms, kernels, rows * cols, "Gsamples/s", all_ok(ok), "one channel filtered alone == its row of the batch",
              f"channels x{world}, no halo")
        del x, y
    else:
        out["config3"] = {"skipped": "not enough free HBM for 8192 x 2^20 int16 in + out on this rank"}

    # ---- config 5: interleaved I/Q, 2^27 complex samples, contiguous segments + halo -----------------
    n = 1 << 28
    pieces = shard.plan_segments(n, world, len(TAPS), 2, align=4096)
    me = pieces[rank]
    lead = shard.aligned_lead(me.before, 2)
    buf = torch.zeros(lead + me.size + me.after, dtype=torch.int16, device=device)
    buf[lead:lead + me.size] = device_signal(torch, (me.size,), 5678 + rank, device)
    if world > 1:
        shard.exchange_halos(buf, lead, pieces, rank, dist)
    body = buf[lead:lead + me.size]
    y = torch.empty(me.size, dtype=torch.int16, device=device)
    ms, kernels = timed(lambda: firgpu.fir1d_fixed_q(body, hq5, FRAC_BITS, ACC_BITS, dilation=2, out_dtype=np.int16, out=y,
                                                    halo_left_valid=me.before, halo_right_valid=me.after))
    assert _lib.last_path() == _lib.PATH_STAGED
    w = min(me.size, 1 << 22)                   # Re of a window filtered as a plain 1-D signal (dilation 1) equals Re of the I/Q pass
    win = buf[lead - me.before:lead + w + me.after]
    re = firgpu.fir1d_fixed_q(win[0::2].contiguous(), hq5, FRAC_BITS, ACC_BITS, out_dtype=np.int16)
    skip = me.before // 2
    ok = torch.equal(re[skip:skip + w // 2 - 2], y[0:w:2][:w // 2 - 2])
    entry("config5", ms, kernels, n, "Gsamples/s", all_ok(ok), "Re of a window filtered as a 1-D signal == Re of the interleaved pass",
          f"segments x{world}, halo {me.before}+{me.after} scalars duplicated at shard time")
    del buf, y, body, win, re

    # ---- config 4: [4, 16384, 16384] int16 images, row bands + halo ------------------------------------
    b_img, hgt, wid = 4, 16384, 16384
    lp = np.asarray(TAPS)
    k_general = np.random.default_rng(0).uniform(-0.2, 0.2, size=(5, 5))
    for key, k in (("config4_3x3", np.full((3, 3), 1 / 9)), ("config4_5x5_rank1", np.outer(lp, lp)), ("config4_5x5_general", k_general)):
        kh = k.shape[0]
        bands = shard.plan_bands(hgt, world, kh)
        me = bands[rank]
        buf = torch.zeros((b_img, me.before + me.size + me.after, wid), dtype=torch.int16, device=device)
        buf[:, me.before:me.before + me.size, :] = device_signal(torch, (b_img, me.size, wid), 4321 + rank, device)
        if world > 1:
            shard.exchange_band_halos(buf, bands, rank, dist)
        body = buf[:, me.before:me.before + me.size, :]
        res = torch.empty((b_img, me.size, wid), dtype=torch.int16, device=device)
        hq = quant.quantize_taps(k, 12, 16)
        ms, kernels = timed(lambda: firgpu.fir2d_fixed_q(body, hq, 12, 32, out_dtype=np.int16, out=res,
                                                        halo_top_valid=me.before, halo_bottom_valid=me.after))
        assert _lib.last_path() == _lib.PATH_STAGED
        # a sub-band with its own halo rows, filtered as a shard of its own, equals the same rows of the band
        r0 = me.size // 2
        sub = firgpu.fir2d_fixed_q(body[1:2, r0:r0 + 256, :], hq, 12, 32, out_dtype=np.int16, halo_top_valid=min(r0, kh - 1 - kh // 2),
                                   halo_bottom_valid=min(me.size - r0 - 256, kh // 2))
        ok = torch.equal(sub[0], res[1, r0:r0 + 256, :])
        entry(key, ms, kernels, b_img * hgt * wid, "Gpixels/s", all_ok(ok), "a 256-row sub-band with halo rows == the same rows of the band",
              f"row bands x{world}, {kh - 1} halo rows duplicated at shard time")
        del buf, res, body, sub
    return out


def gpu_arm(args, rank: int, local_rank: int, world: int) -> None:
    import torch
    import torch.distributed as dist

    import firgpu
    from firgpu import _lib, quant

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL logs (its version banner included) go to stdout by default: stdout carries ONE JSON line only
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    piece, lead, buf, body = make_segment(torch, dist, rank, world, device)
    y = torch.empty(piece.size, dtype=torch.int16, device=device)
    hq = quant.quantize_taps(TAPS, FRAC_BITS, COEFF_BITS)

    def step():
        firgpu.fir1d_fixed_q(body, hq, FRAC_BITS, ACC_BITS, out_dtype=np.int16, out=y,
                             halo_left_valid=piece.before, halo_right_valid=piece.after)

    # Parity is established by tests/ and __graft_entry__.smoke(); here only the kernel family is asserted
    # (the oracle is not used in this arm: it is executed by the cpu_baseline / --impl reference legs alone).
    step()
    assert _lib.last_path() == _lib.PATH_STAGED, "the staged sm_100a kernel must serve the benchmark shape"

    clocks = ClockSampler(local_rank if "CUDA_VISIBLE_DEVICES" not in os.environ else
                          int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank]))
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = _lib.launch_count()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks.region(True)
    start.record()
    for _ in range(args.steps):
        step()
    end.record()
    torch.cuda.synchronize()
    clocks.region(False)
    launches = _lib.launch_count() - launches0
    if world > 1:
        dist.barrier()
    ms_total = torch.tensor([start.elapsed_time(end)], device=device, dtype=torch.float64)
    ms_ranks = [float(ms_total.item()) / args.steps]
    if world > 1:
        gathered = [torch.zeros_like(ms_total) for _ in range(world)]
        dist.all_gather(gathered, ms_total)
        ms_ranks = [float(t.item()) / args.steps for t in gathered]
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_total = float(ms_total.item())
    ms_per_step = ms_total / args.steps
    value = world * piece.size * args.steps / (ms_total * 1e-3) / 1e9

    # ---- end to end through the public API with host buffers -----------------
    e2e_value, e2e_ok, e2e_steps = None, None, 0
    n_host = piece.before + piece.size + piece.after
    if not args.no_e2e:
        e2e_steps = max(3, min(args.steps, 8))
        x_host = torch.empty(n_host, dtype=torch.int16).pin_memory()
        y_host = torch.empty(n_host, dtype=torch.int16).pin_memory()
        x_host.copy_(buf[lead - piece.before:lead + piece.size + piece.after])
        x_np, y_np = x_host.numpy(), y_host.numpy()
        for _ in range(2):
            firgpu.fir1d_fixed(x_np, TAPS, FRAC_BITS, ACC_BITS, COEFF_BITS, out_dtype=np.int16, out=y_np, device=local_rank)
        barrier()
        clocks.region(True)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            firgpu.fir1d_fixed(x_np, TAPS, FRAC_BITS, ACC_BITS, COEFF_BITS, out_dtype=np.int16, out=y_np, device=local_rank)
        torch.cuda.synchronize()
        e2e_s = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
        clocks.region(False)
        if world > 1:
            dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        e2e_value = world * piece.size * e2e_steps / float(e2e_s.item()) / 1e9
        e2e_ok = bool(torch.equal(y_host[piece.before:piece.before + piece.size], y.cpu())) if rank == 0 else None

    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    configs = None
    if not args.no_configs:
        del buf, body, y
        if not args.no_e2e:
            del x_host, y_host, x_np, y_np
        torch.cuda.empty_cache()
        clocks.region(True)
        configs = other_configs(args, torch, dist, rank, world, device, peak)
        clocks.region(False)

    clock_summary = clocks.summary()
    if rank == 0:
        achieved = BYTES_PER_SAMPLE * piece.size / (ms_per_step * 1e-3) / 1e9
        traffic, traffic_src = None, None
        tp = ROOT / "profiles" / "traffic.json"
        if tp.exists():
            tj = json.loads(tp.read_text())
            traffic = tj.get("fir1d_staged_i16_i16_5tap_2p28_dram_bytes_per_launch")
            traffic_src = tj.get("source")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32 accumulate (dp2a: int16 taps x u8 samples), int16 containers",
            "data": "synthetic",
            "config": {"workload": workload_name(world), "taps": "simple_lp 5-tap", "frac_bits": FRAC_BITS,
                       "acc_bits": ACC_BITS, "coeff_bits": COEFF_BITS, "samples_per_gpu": piece.size,
                       "l2": "inputs (512 MiB/GPU) larger than L2, no flush needed", "parallelism": f"segments x{world}"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "bytes_per_launch": BYTES_PER_SAMPLE * piece.size, "kernel": "fir1d_staged_kernel<int16,int16,5,1>"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * n_host * world,
                    "d2h_bytes_per_step": 2 * n_host * world, "steps": e2e_steps,
                    "api": "firgpu.fir1d_fixed(host int16 ndarray, pinned) -> firgpu_fir1d_fixed_host", "matches_device_run": e2e_ok},
            "gpu_launches": int(launches), "clocks": clock_summary,
            "ms_per_step_by_rank": [round(v, 6) for v in ms_ranks],      # the headline takes the slowest
        }
        if configs is not None:
            line["configs"] = configs
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_block()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=("ours", "reference"), default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-buffer leg")
    ap.add_argument("--no-configs", action="store_true", help="profiling runs only: skip BASELINE configs 3 / 4 / 5")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    gpu_arm(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
