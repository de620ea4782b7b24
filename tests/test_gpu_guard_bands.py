"""Out-of-bounds WRITE detection without compute-sanitizer (closed on this GPU
pool): every kernel family writes into the middle of a larger buffer whose
surroundings hold a sentinel; after the call the sentinel must be intact and
the payload must equal the oracle.  Shapes are chosen to end mid-tile /
mid-vector, with aligned and unaligned output bases and row pitches."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import fir_oracle_np as onp

pytestmark = pytest.mark.gpu

LP5 = [1 / 16, 4 / 16, 6 / 16, 4 / 16, 1 / 16]
SENTINEL = 0x5A


def guarded(torch, shape, dtype, lead, pitch_extra):
    """Rows of `shape[1]` elements at pitch shape[1]+pitch_extra, preceded by
    `lead` guard elements and followed by 4096; everything pre-filled with the
    sentinel byte pattern.  Returns (whole buffer, view to write into)."""
    rows, cols = shape
    pitch = cols + pitch_extra
    total = lead + rows * pitch + 4096
    buf = torch.full((total * torch.empty((), dtype=dtype).element_size(),), SENTINEL, dtype=torch.uint8, device="cuda").view(dtype)
    view = buf[lead:lead + rows * pitch].view(rows, pitch)[:, :cols]
    return buf, view


def check_guards(torch, buf, view, lead, shape, pitch_extra, what):
    rows, cols = shape
    pitch = cols + pitch_extra
    raw = buf.view(torch.uint8)
    es = buf.element_size()
    mask = torch.ones(raw.numel(), dtype=torch.bool, device="cuda")
    for r in range(rows):
        a = (lead + r * pitch) * es
        mask[a:a + cols * es] = False
    assert bool((raw[mask] == SENTINEL).all()), f"{what}: wrote outside its output"


@pytest.mark.parametrize("lead, pitch_extra", [(0, 0), (8, 8), (3, 5), (1, 0)])
def test_1d_kernels_stay_inside_their_output(lead, pitch_extra):
    import torch
    import firgpu
    rng = np.random.default_rng(lead * 10 + pitch_extra)
    cases = []
    for in_dt, out_dt in ((np.int16, np.int16), (np.uint8, np.uint8), (np.int16, np.uint8), (np.uint8, np.int16)):
        for shape in ((1, 1), (1, 16385), (3, 4099), (20, 300), (2, 70_001)):
            cases.append((in_dt, out_dt, shape, dict(h=LP5, frac=15)))
    cases.append((np.int16, np.int16, (2, 40_003), dict(h=(np.ones(31) / 62).tolist(), frac=15)))        # long-filter bucket
    cases.append((np.int16, np.int16, (2, 40_003), dict(h=LP5, frac=15, in_range=(-32768, 32767), out_range=(-32768, 32767))))   # wide
    cases.append((np.int16, np.int16, (1, 40_002), dict(h=LP5, frac=15, dilation=2)))                     # I/Q
    cases.append((np.int16, np.int16, (2, 5_003), dict(h=[0.9, -0.7, 0.8], frac=12, acc_bits=16)))        # generic
    for in_dt, out_dt, shape, kw in cases:
        x = rng.integers(0, 256, size=shape).astype(in_dt)
        tdt = torch.int16 if out_dt == np.int16 else torch.uint8
        buf, view = guarded(torch, shape, tdt, lead, pitch_extra)
        extra = {k: v for k, v in kw.items() if k in ("in_range", "out_range", "dilation")}
        firgpu.fir1d_fixed(torch.from_numpy(x).cuda(), kw["h"], kw["frac"], kw.get("acc_bits", 32), 16, out_dtype=out_dt, out=view, **extra)
        want = onp.fir1d_fixed(x, kw["h"], kw["frac"], kw.get("acc_bits", 32), 16, out_dtype=out_dt, **extra)
        assert np.array_equal(view.cpu().numpy(), want), (in_dt, out_dt, shape, kw)
        check_guards(torch, buf, view, lead, shape, pitch_extra, (in_dt, out_dt, shape, kw))


def test_ideal_multi_and_2d_kernels_stay_inside_their_output():
    import torch
    import firgpu
    rng = np.random.default_rng(5)
    # ideal fast + generic
    for shape, lead in (((3, 4099), 0), ((1, 2049), 1), ((5, 2048), 2)):
        x = rng.integers(0, 256, size=shape).astype(np.uint8)
        buf, view = guarded(torch, shape, torch.float64, lead, 2)
        firgpu.fir1d_ideal(torch.from_numpy(x).cuda(), LP5, out=view)
        assert np.array_equal(view.cpu().numpy(), onp.fir1d_ideal(x, LP5))
        check_guards(torch, buf, view, lead, shape, 2, ("ideal", shape))
    # 2-D tiled: outputs are contiguous images; guard = sentinel before and after
    for dtype, shape in ((np.int16, (2, 70, 1000)), (np.uint8, (1, 131, 2016)), (np.int16, (1, 67, 1016))):
        img = rng.integers(0, 256, size=shape).astype(dtype)
        tdt = torch.int16 if dtype == np.int16 else torch.uint8
        n = int(np.prod(shape))
        for lead in (0, 64):
            buf = torch.full(((lead + n + 4096) * np.dtype(dtype).itemsize,), SENTINEL, dtype=torch.uint8, device="cuda").view(tdt)
            out = buf[lead:lead + n].view(shape)
            for k in ([[1 / 9] * 3] * 3, np.outer(LP5, LP5).tolist(), rng.uniform(-0.2, 0.2, size=(5, 5)).tolist()):
                firgpu.fir2d_fixed(torch.from_numpy(img).cuda(), k, out_dtype=dtype, out=out)
                assert np.array_equal(out.cpu().numpy(), onp.fir2d_fixed(img, k, out_dtype=dtype))
                raw = buf.view(torch.uint8)
                es = np.dtype(dtype).itemsize
                assert bool((raw[:lead * es] == SENTINEL).all()) and bool((raw[(lead + n) * es:] == SENTINEL).all()), (dtype, shape, lead)
