"""GPU parity of the 2-D path.  fir_2d is PARITY UNPINNED against the reference
(it ships no fir_2d); what is pinned: (i) single-row / single-column kernels
reduce bit-exactly to the reference-pinned 1-D model, (ii) the separable form
equals two 1-D passes, and the full 2-D kernels are bit-exact against the three
agreeing restatements of our specification in oracle/."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import c_oracle, fir_oracle_np as onp

pytestmark = pytest.mark.gpu

LP5 = [1 / 16, 4 / 16, 6 / 16, 4 / 16, 1 / 16]
K3_BOX = [[1 / 9] * 3] * 3
K3_SHARPEN = [[0, -1, 0], [-1, 5, -1], [0, -1, 0]]
K3_LAPLACE = [[0, 1, 0], [1, -4, 1], [0, 1, 0]]
K5_LP = np.outer(LP5, LP5).tolist()


def cuda(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def image(rng, shape, dtype=np.int16):
    x = rng.integers(0, 256, size=shape).astype(dtype)
    if dtype != np.uint8:
        mask = rng.random(shape) < 0.01
        x[mask] = rng.integers(-32768, 32767, size=int(mask.sum())).astype(dtype)
    return x


def test_degenerate_kernels_equal_reference_pinned_1d():
    import firgpu
    from fir_2d.model.python.fir_2d_fixed_ref import fir_2d_fixed_golden, fir_2d_separable_fixed
    rng = np.random.default_rng(1)
    x = image(rng, (37, 53), np.uint8)
    for h in ([0.25, 0.5, 0.25], LP5, [-1 / 8, -2 / 8, 0, 2 / 8, 1 / 8], [0.5, 0.25]):
        rows = fir_2d_fixed_golden(x, [h])
        assert rows.dtype == np.uint8 and np.array_equal(rows, onp.fir1d_fixed(x, h))
        cols = fir_2d_fixed_golden(x, [[v] for v in h])
        assert np.array_equal(cols, onp.fir1d_fixed(x.T, h).T)
        sep = fir_2d_separable_fixed(x, h, h)
        assert np.array_equal(sep, onp.fir2d_separable_fixed(x, h, h))
    got = firgpu.fir2d_fixed(cuda(x), [[1.0]]).cpu().numpy()
    assert np.array_equal(got, x)


@pytest.mark.parametrize("dtype", [np.uint8, np.int16])
@pytest.mark.parametrize("out_dtype", [np.uint8, np.int16])
def test_full_kernels_against_oracle(dtype, out_dtype):
    import firgpu
    rng = np.random.default_rng(2)
    shapes = [(1, 1, 1), (1, 2, 3), (2, 5, 300), (1, 64, 64), (3, 97, 517), (1, 300, 2050), (2, 33, 4499)]
    kernels = [(K3_BOX, 12), (K3_SHARPEN, 12), (K3_LAPLACE, 12), (K5_LP, 12), (K5_LP, 15),
               (rng.uniform(-0.3, 0.3, size=(5, 5)).tolist(), 12), (rng.uniform(-0.4, 0.4, size=(3, 5)).tolist(), 12),
               (rng.uniform(-0.4, 0.4, size=(4, 2)).tolist(), 12), (rng.uniform(-0.1, 0.1, size=(7, 7)).tolist(), 12)]
    for shape in shapes:
        x = image(rng, shape, dtype)
        for k, frac in kernels:
            hq = onp.quantize_taps(k, frac, 16)
            want = c_oracle.fir2d_fixed(x, hq, frac, 32, out_dtype=out_dtype, threads=8)
            got = firgpu.fir2d_fixed(cuda(x), k, frac, 32, 16, out_dtype=out_dtype).cpu().numpy()
            assert got.dtype == np.dtype(out_dtype)
            assert np.array_equal(got, want), (shape, np.shape(k), frac)


def test_tiled_tma_kernel_serves_aligned_shapes():
    """The TMA tensor-map kernel (general and rank-1 separable variants) must
    be the one that runs for 16-byte-aligned uint8 / int16 images, and it must
    be bit-exact — including all four zero-padded borders, which it gets from
    the TMA unit's out-of-bounds fill, and strips / bands that end mid-tile."""
    import firgpu
    from firgpu import _lib
    rng = np.random.default_rng(7)
    lp3 = [0.25, 0.5, 0.25]
    kernels = [("box3", K3_BOX, 12), ("sharpen3", K3_SHARPEN, 12), ("lp5 rank-1", K5_LP, 12), ("lp5 q1.15", K5_LP, 15),
               ("rand5", rng.uniform(-0.3, 0.3, size=(5, 5)).tolist(), 12),
               ("3x5 rank-1", np.outer(lp3, LP5).tolist(), 12), ("5x3 rank-1", np.outer(LP5, lp3).tolist(), 12),
               ("3x5", rng.uniform(-0.4, 0.4, size=(3, 5)).tolist(), 12), ("5x3", rng.uniform(-0.4, 0.4, size=(5, 3)).tolist(), 12),
               ("5x1", [[v] for v in LP5], 12), ("3x1", [[v] for v in lp3], 12)]
    for dtype, shapes in ((np.int16, [(1, 1, 8), (2, 70, 1000), (1, 131, 2048), (3, 67, 1016), (1, 200, 2024)]),
                          (np.uint8, [(1, 3, 16), (2, 70, 2016), (1, 140, 4064), (1, 61, 2032)])):
        for shape in shapes:
            x = image(rng, shape, dtype)
            xt = cuda(x)
            for name, k, frac in kernels:
                hq = onp.quantize_taps(k, frac, 16)
                want = c_oracle.fir2d_fixed(x, hq, frac, 32, out_dtype=dtype, threads=8)
                got = firgpu.fir2d_fixed(xt, k, frac, 32, 16, out_dtype=dtype)
                assert _lib.last_path() == _lib.PATH_STAGED, (name, shape)
                assert np.array_equal(got.cpu().numpy(), want), (name, dtype, shape)


def test_2d_generic_cases_wrap_and_float_inputs():
    import firgpu
    rng = np.random.default_rng(3)
    x = rng.uniform(-30, 300, size=(2, 40, 70))
    k = rng.uniform(-2, 2, size=(3, 3)).tolist()
    for acc in (16, 24, 32, 64):
        assert np.array_equal(firgpu.fir2d_fixed(cuda(x), k, 12, acc, 16).cpu().numpy(), onp.fir2d_fixed(x, k, 12, acc, 16)), acc
    xi = image(rng, (40, 70), np.int32)
    assert np.array_equal(firgpu.fir2d_fixed(cuda(xi), k).cpu().numpy(), onp.fir2d_fixed(xi, k))


def test_2d_ideal_bit_identical_and_f32_tolerance():
    import firgpu
    from fir_2d.model.python.fir_2d_ref import fir_2d_ideal
    rng = np.random.default_rng(4)
    x = image(rng, (2, 61, 130), np.uint8)
    for k in (K3_BOX, K5_LP, rng.uniform(-1, 1, size=(3, 5)).tolist()):
        want = onp.fir2d_ideal(x, k)
        assert np.array_equal(fir_2d_ideal(cuda(x), k).cpu().numpy(), want)
        assert np.array_equal(fir_2d_ideal(x, k), want)
        got32 = firgpu.fir2d_ideal(cuda(x), k, out_dtype=np.float32).cpu().numpy().astype(np.float64)
        assert np.max(np.abs(got32 - want)) <= 1e-6 * 255 * float(np.sum(np.abs(k)))


def test_2d_ideal_fast_kernel_bit_identical():
    """fir2d_ideal_fast_kernel (uint8 / int16 pixels, 3x3 / 5x5 / 3x5 / 5x3): even
    and odd widths, column views (every store / load alignment), several
    images, row bands with halo rows — bit-identical to the oracle in float64,
    within the stated tolerance in float32, and nothing written outside."""
    import torch
    import firgpu
    from firgpu import _lib, shard
    rng = np.random.default_rng(41)
    kernels = [rng.uniform(-2, 2, size=s).tolist() for s in ((3, 3), (5, 5), (3, 5), (5, 3))]
    for dtype in (np.uint8, np.int16):
        for shape in ((1, 9, 16), (2, 33, 1031), (1, 70, 2050), (3, 5, 777)):
            x = image(rng, shape, dtype)
            for k in kernels:
                got = firgpu.fir2d_ideal(cuda(x), k)
                assert _lib.last_path() == _lib.PATH_STAGED, (dtype, shape)
                assert np.array_equal(got.cpu().numpy(), onp.fir2d_ideal(x, k)), (dtype, shape, np.shape(k))
    for dtype in (np.float32, np.float64, np.int32):          # rounded / clamped to uint8 pixels first, then the same kernel
        x = (rng.uniform(-30, 300, size=(2, 90, 777))).astype(dtype)
        for k in kernels[:2]:
            got = firgpu.fir2d_ideal(cuda(x), k)
            assert _lib.last_path() == _lib.PATH_STAGED, dtype
            assert np.array_equal(got.cpu().numpy(), onp.fir2d_ideal(x, k)), dtype
    wide = image(rng, (2, 40, 1500), np.uint8)
    wide_t = cuda(wide)
    k = kernels[1]
    for out_dtype, tdt in ((np.float64, torch.float64), (np.float32, torch.float32)):
        for c0, width in ((0, 1500), (1, 1498), (3, 1001), (2, 640)):
            out = torch.full((2, 40, 1500), -7.0, dtype=tdt, device="cuda")
            firgpu.fir2d_ideal(wide_t[:, :, c0:c0 + width], k, out_dtype=out_dtype, out=out[:, :, c0:c0 + width])
            assert _lib.last_path() == _lib.PATH_STAGED
            want = onp.fir2d_ideal(wide[:, :, c0:c0 + width], k)
            got = out.cpu().numpy()
            if out_dtype is np.float64:
                assert np.array_equal(got[:, :, c0:c0 + width], want), (c0, width)
            else:
                assert np.abs(got[:, :, c0:c0 + width] - want).max() <= 1e-6 * 255 * np.abs(k).sum()
            assert (got[:, :, :c0] == -7.0).all() and (got[:, :, c0 + width:] == -7.0).all(), (c0, width)
    x = image(rng, (1, 300, 514), np.int16)
    want = onp.fir2d_ideal(x, k)
    got = np.empty_like(want)
    for band in shard.plan_bands(300, 3, 5):
        buf, body = shard.scatter_band(x, band, "cuda")
        y = firgpu.fir2d_ideal(body, k, halo_top_valid=band.before, halo_bottom_valid=band.after)
        got[..., band.start:band.stop, :] = y.cpu().numpy()
    assert np.array_equal(got, want)


def test_row_band_shards_are_seamless():
    """Config 4 sharding: row bands with duplicated halo rows give exactly the
    whole-image result."""
    import firgpu
    from firgpu import shard
    from firgpu import _lib
    rng = np.random.default_rng(5)
    # a 2200-byte pitch is not TMA-able: such bands are copied to an aligned pitch first, then take the same kernel
    for width, path in ((1100, _lib.PATH_STAGED), (1104, _lib.PATH_STAGED)):
        x = image(rng, (2, 700, width))
        for k in (K3_SHARPEN, K5_LP):
            kh = len(k)
            want = c_oracle.fir2d_fixed(x, onp.quantize_taps(k, 12, 16), 12, 32, out_dtype=np.int16, threads=8)
            for parts in (2, 3, 8):
                got = np.empty_like(want)
                for band in shard.plan_bands(x.shape[-2], parts, kh):
                    buf, body = shard.scatter_band(x, band, "cuda")
                    y = firgpu.fir2d_fixed(body, k, 12, 32, 16, out_dtype=np.int16,
                                           halo_top_valid=band.before, halo_bottom_valid=band.after)
                    assert _lib.last_path() == path
                    got[:, band.start:band.stop, :] = y.cpu().numpy()
                assert np.array_equal(got, want), (width, kh, parts)


def test_2d_images_of_odd_width_and_pitch():
    """uint8 / int16 images the tensor-map kernel cannot read in place (odd
    widths such as the reference's 4499, pitches that are not multiples of 16
    bytes, column views, unaligned bases) are re-pitched into an aligned
    temporary and run on the tiled kernel: bit-exact, nothing written outside."""
    import torch
    import firgpu
    from firgpu import _lib
    rng = np.random.default_rng(15)
    for dtype in (np.uint8, np.int16):
        for shape in ((1, 200, 4499), (2, 97, 333), (3, 64, 1001), (1, 70, 1102)):
            x = image(rng, shape, dtype)
            for k in (K3_BOX, K3_SHARPEN, K5_LP, rng.uniform(-0.4, 0.4, size=(5, 5)).tolist(), rng.uniform(-1, 1, size=(3, 5)).tolist()):
                got = firgpu.fir2d_fixed(cuda(x), k, 12, 32, 16, out_dtype=dtype)
                assert _lib.last_path() == _lib.PATH_STAGED, (dtype, shape, np.shape(k))
                assert np.array_equal(got.cpu().numpy(), onp.fir2d_fixed(x, k, 12, 32, 16, out_dtype=dtype)), (dtype, shape, np.shape(k))
    # int16 container in, uint8 out (the drop-in's default output type): narrowed to uint8 pixels first
    from fir_2d.model.python.fir_2d_fixed_ref import fir_2d_fixed_golden
    x16 = image(rng, (2, 300, 1024), np.int16)
    for k in (K3_BOX, K5_LP):
        got = fir_2d_fixed_golden(cuda(x16), k)
        assert _lib.last_path() == _lib.PATH_STAGED and got.dtype == torch.uint8
        assert np.array_equal(got.cpu().numpy(), onp.fir2d_fixed(x16, k))
    x = image(rng, (1, 120, 2000), np.uint8)
    xt = cuda(x)
    for c0, width in ((1, 1998), (3, 1001), (16, 1024)):                  # column views: unaligned base / odd width
        got = firgpu.fir2d_fixed_q(xt[:, :, c0:c0 + width], onp.quantize_taps(K5_LP, 12, 16), 12, 32)
        assert np.array_equal(got.cpu().numpy(), onp.fir2d_fixed(x[:, :, c0:c0 + width], K5_LP)), (c0, width)


def test_2d_host_entry_point():
    import firgpu
    rng = np.random.default_rng(6)
    x = image(rng, (2, 3000, 4100))         # 49 MB of int16: several row bands per image
    k = K5_LP
    want = c_oracle.fir2d_fixed(x, onp.quantize_taps(k, 12, 16), 12, 32, out_dtype=np.int16, threads=8)
    got = firgpu.fir2d_fixed(x, k, 12, 32, 16, out_dtype=np.int16)
    assert np.array_equal(got, want)
    x8 = image(rng, (1, 200, 333), np.uint8)
    assert np.array_equal(firgpu.fir2d_fixed(x8, K3_BOX), onp.fir2d_fixed(x8, K3_BOX))


def test_full_size_image_properties():
    """Config 4 at full size (16384 x 16384 int16): crops against the oracle at
    all four borders and the centre + global properties."""
    import torch
    import firgpu
    n = 16384
    g = torch.Generator(device="cuda").manual_seed(4321)
    x = torch.randint(0, 256, (1, n, n), generator=g, device="cuda", dtype=torch.int16)
    for k in (K3_BOX, K5_LP):
        y = firgpu.fir2d_fixed(x, k, 12, 32, 16, out_dtype=np.int16)
        hq = onp.quantize_taps(k, 12, 16)
        r = len(k) // 2
        for (i0, j0) in ((0, 0), (0, n - 512), (n - 512, 0), (n - 512, n - 512), (n // 2 - 256, n // 2 - 256), (4000, 8100)):
            ia, ib = max(0, i0 - r), min(n, i0 + 512 + r)
            ja, jb = max(0, j0 - r), min(n, j0 + 512 + r)
            crop = x[0, ia:ib, ja:jb].cpu().numpy()
            want = c_oracle.fir2d_fixed(crop, hq, 12, 32, out_dtype=np.int16, threads=8)
            got = y[0, i0:i0 + 512, j0:j0 + 512].cpu().numpy()
            assert np.array_equal(got, want[i0 - ia:i0 - ia + 512, j0 - ja:j0 - ja + 512]), (len(k), i0, j0)
        assert int(y.min()) >= 0 and int(y.max()) <= 255
        # transposition symmetry of a symmetric kernel
        yt = firgpu.fir2d_fixed(x.transpose(1, 2).contiguous(), k, 12, 32, 16, out_dtype=np.int16)
        assert torch.equal(yt.transpose(1, 2), y)
        del y, yt


def test_windows_up_to_7x7_and_even_sizes_run_on_the_tensor_map_kernel():
    """Every window from 1x2 to 7x7 — odd sizes instantiated, even sizes by a zero row / column appended at the end, which
    keeps the model's centre L // 2 — is served by the tensor-map kernel and is bit-exact against the oracle: general
    and exactly rank-1 kernels, uint8 / int16 pixels, the narrowing output, Q4.12 / Q1.15, a live 16-bit accumulator
    wrap, borders, ragged strips, several images, and row-band shards of an even-height window (asymmetric halo)."""
    import firgpu
    from firgpu import _lib, quant, shard
    rng = np.random.default_rng(77)
    g7 = np.array([1, 6, 15, 20, 15, 6, 1]) / 64.0
    x16 = image(rng, (2, 90, 1000), np.int16)
    x8 = image(rng, (1, 75, 2032), np.uint8)
    t16, t8 = cuda(x16), cuda(x8)
    n = 0
    for kh in range(1, 8):
        for kw in range(1, 8):
            if kh == 1 and kw == 1:
                continue
            ks = [(rng.uniform(-0.25, 0.25, size=(kh, kw)), 12, 32)]
            if max(kh, kw) >= 6 or (kh % 2 == 0 and kw % 2 == 0):
                ks.append((np.outer(g7[:kh] / g7[:kh].sum(), g7[:kw] / g7[:kw].sum()), 15, 32))      # rank-1 before quantisation
                ks.append((rng.uniform(-0.5, 0.5, size=(kh, kw)), 12, 16))                            # live wrap
            for k, frac, acc in ks:
                hq = quant.quantize_taps(k, frac, 16)
                for x, xt, outs in ((x16, t16, (np.int16, np.uint8)), (x8, t8, (np.uint8,))):
                    for out_dtype in outs:
                        want = c_oracle.fir2d_fixed(x, hq, frac, acc, out_dtype=out_dtype, threads=8)
                        got = firgpu.fir2d_fixed_q(xt, hq, frac, acc, out_dtype=out_dtype)
                        assert _lib.last_path() == _lib.PATH_STAGED, (kh, kw, x.dtype)
                        assert np.array_equal(got.cpu().numpy(), want), (kh, kw, frac, acc, x.dtype, out_dtype)
                        n += 1
    assert n >= 200
    # an exactly rank-1 integer 7x7 kernel (the separable instantiation) and its row-band shards; a 6x4 window in bands
    a = np.array([1, 6, 15, 20, 15, 6, 1], dtype=np.int32)
    img = image(rng, (1, 400, 1024), np.int16)
    it = cuda(img)
    for hq, frac in ((np.outer(a, a), 12), (rng.integers(-300, 300, size=(6, 4)).astype(np.int32), 12)):
        kh = hq.shape[0]
        want = c_oracle.fir2d_fixed(img, hq, frac, 32, out_dtype=np.int16, threads=8)
        assert np.array_equal(firgpu.fir2d_fixed_q(it, hq, frac, 32, out_dtype=np.int16).cpu().numpy(), want)
        got = np.empty_like(want)
        for band in shard.plan_bands(400, 3, kh):
            buf, body = shard.scatter_band(img, band, "cuda")
            y = firgpu.fir2d_fixed_q(body, hq, frac, 32, out_dtype=np.int16, halo_top_valid=band.before, halo_bottom_valid=band.after)
            assert _lib.last_path() == _lib.PATH_STAGED
            got[:, band.start:band.stop, :] = y.cpu().numpy()
        assert np.array_equal(got, want), kh
    # batches of short rows through the 1-D entry: 2, 4, 6 and 7 taps now take the row-streaming form as 3 and 5 do
    rows = image(rng, (40, 1000), np.int16)
    rt = cuda(rows)
    for taps in (2, 4, 6, 7):
        h = rng.uniform(-0.3, 0.3, size=taps).tolist()
        got = firgpu.fir1d_fixed(rt, h, 12, 32, 16, out_dtype=np.int16)
        assert _lib.last_path() == _lib.PATH_STAGED
        assert np.array_equal(got.cpu().numpy(), onp.fir1d_fixed(rows, h, 12, 32, 16, out_dtype=np.int16)), taps
