"""GPU parity: the ideal (floating-point) model, and the reference-shaped
row-wise generators / pipeline driver with their file contract
(fir_1d/sim/tests/test_1d_ideal.py, test_fixed_output.py, test_ideal_ output.py
restated against the GPU-backed drop-in)."""
from __future__ import annotations

import hashlib
import re

import numpy as np
import pytest

from oracle import fir_oracle_np as onp

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ideal_fn():
    from fir_1d.model.python.fir_1d_ref import fir_1d_ideal
    return fir_1d_ideal


# ---- fir_1d/sim/tests/test_1d_ideal.py ---------------------------------------------
def test_reference_ideal_kats():
    f = ideal_fn()
    y = f([10, 20, 30, 40], [0.25, 0.5, 0.25])                                     # :8-16
    assert isinstance(y, list) and y == pytest.approx([10.0, 20.0, 30.0, 27.5], abs=1e-12)
    assert len(f([1, 2, 3, 4, 5], [0.1, 0.2, 0.3, 0.4])) == 5                       # :19-24
    assert f([-1.2, 0.49, 0.5, 1.5, 254.6, 300.2], [1.0]) == [0.0, 0.0, 1.0, 2.0, 255.0, 255.0]   # :27-34
    assert f([255, 255], [5.0]) == [1275.0, 1275.0]                                 # :37-42 not clamped
    assert f([1.0], [8.0]) == [8.0] and f([1.0], [-8.0]) == [-8.0]                  # :67-69


@pytest.mark.parametrize("kwargs, pattern", [
    (dict(x=[1.0, float("inf")], h=[1.0]), r"x\[1\].*finite"),
    (dict(x=[1.0], h=[float("nan")]), r"h\[0\].*finite"),
    (dict(x=[1.0], h=[]), r"must not be empty"),
    (dict(x=[1.0], h=[8.0 + 1e-6]), r"\|h\| must be <="),
])
def test_reference_ideal_errors(kwargs, pattern):
    with pytest.raises(ValueError) as err:
        ideal_fn()(**kwargs)
    assert re.search(pattern, str(err.value))


def test_ideal_f64_is_bit_identical_on_every_golden_case(golden_cases):
    f = ideal_fn()
    for c in golden_cases:
        y = np.array(f(c["x"].tolist(), c["h"]), dtype=np.float64)
        assert np.array_equal(y, c["y_ideal"]), c["name"]


def test_ideal_images_by_hash(images_inputs, manifest):
    from fir_1d.sim.vector.gen_ideal_output import _run_ideal_rowwise
    for stem in ("case_003_img_004_64x64_gray", "case_004_img_005_64x64_noise_gray", "case_001_img_002_640x762_gray"):
        x = images_inputs[stem]
        for tap in ("3tap", "5tap"):
            for name, h in manifest[f"coeff_{tap}"].items():
                y = _run_ideal_rowwise(x, h)
                assert y.dtype == np.float64 and y.shape == x.shape
                assert sha(y) == manifest["ideal"][f"{stem}__{name}_{tap}"], (stem, name, tap)


def test_ideal_f32_within_stated_tolerance():
    """float32 datapath: |err| <= 1e-6 * 255 * sum|h| (relative 1e-6 of the
    output's full scale), the tolerance BASELINE.json's north star names."""
    import torch
    import firgpu
    rng = np.random.default_rng(2)
    x = rng.integers(0, 256, size=(7, 50_001)).astype(np.uint8)
    for h in ([1 / 3] * 3, [1 / 16, 4 / 16, 6 / 16, 4 / 16, 1 / 16], [-1 / 16, -4 / 16, 26 / 16, -4 / 16, -1 / 16], [7.3, -2.1]):
        want = onp.fir1d_ideal(x, h)
        got = firgpu.fir1d_ideal(torch.from_numpy(x).cuda(), h, out_dtype=np.float32).cpu().numpy()
        assert got.dtype == np.float32
        tol = 1e-6 * 255 * float(np.sum(np.abs(h)))
        assert np.max(np.abs(got.astype(np.float64) - want)) <= tol
        got64 = firgpu.fir1d_ideal(torch.from_numpy(x).cuda(), h).cpu().numpy()
        assert np.array_equal(got64, want)
        # even width: the fast kernel's float32 variant
        from firgpu import _lib
        xe = np.ascontiguousarray(x[:, :50_000])
        got32 = firgpu.fir1d_ideal(torch.from_numpy(xe).cuda(), h, out_dtype=np.float32)
        assert _lib.last_path() == _lib.PATH_STAGED and got32.dtype == torch.float32
        assert np.max(np.abs(got32.cpu().numpy().astype(np.float64) - onp.fir1d_ideal(xe, h))) <= tol
        host = firgpu.fir1d_ideal(x, h)
        assert np.array_equal(host, want)


def test_fixed_tracks_ideal_within_one_lsb_when_unsaturated():
    """Cross-check of the two models on the GPU (what the reference's compare
    reports measure): a unity-gain low-pass never saturates, so the fixed
    output is the rounded ideal output up to coefficient quantisation."""
    import torch
    import firgpu
    rng = np.random.default_rng(6)
    x = torch.from_numpy(rng.integers(0, 256, size=1 << 20).astype(np.uint8)).cuda()
    h = [1 / 16, 4 / 16, 6 / 16, 4 / 16, 1 / 16]
    fixed = firgpu.fir1d_fixed(x, h).double()
    ideal = firgpu.fir1d_ideal(x, h)
    assert float((fixed - ideal).abs().max()) <= 0.5 + 1e-9


# ---- generators (fir_1d/sim/tests/test_fixed_output.py, "test_ideal_ output.py") ----
SMALL = np.array([[10, 20, 30, 40, 50, 60, 70, 80],
                  [80, 70, 60, 50, 40, 30, 20, 10],
                  [0, 10, 0, 10, 0, 10, 0, 10],
                  [255, 200, 150, 100, 50, 0, 25, 75]], dtype=np.uint8)      # output_test_common.py:12-25


def test_generators_file_contract_and_values(tmp_path):
    from fir_1d.model.python.fir_1d_fixed_ref import fir_1d_fixed_golden
    from fir_1d.model.python.fir_1d_ref import fir_1d_ideal
    from fir_1d.sim.vector import gen_fixed_output as gf, gen_ideal_output as gi
    from fir_1d.sim.vector.h_coeff import h_coeff_3tap_map, h_coeff_5tap_map
    inp, out = tmp_path / "input", tmp_path / "output"
    inp.mkdir()
    np.save(inp / "case_000_small_x_u8.npy", SMALL)
    assert gf.generate_fixed_3tap_output_vector(input_dir=inp, output_dir=out) == len(h_coeff_3tap_map)
    assert gf.generate_fixed_5tap_output_vector(input_dir=inp, output_dir=out) == len(h_coeff_5tap_map)
    assert gi.generate_ideal_3tap_output_vector(input_dir=inp, output_dir=out) == 4
    assert gi.generate_ideal_5tap_output_vector(input_dir=inp, output_dir=out) == 4
    for sub, suffix, dtype in (("fixed_3tap", "_fixed_3tap_y_u8.npy", np.uint8), ("fixed_5tap", "_fixed_5tap_y_u8.npy", np.uint8),
                               ("ideal_3tap", "_ideal_3tap_y_f64.npy", np.float64), ("ideal_5tap", "_ideal_5tap_y_f64.npy", np.float64)):
        files = sorted((out / sub).glob("*.npy"))
        assert len(files) == 4 and all("__" in f.name and f.name.endswith(suffix) for f in files)
        for f in files:
            y = np.load(f)
            assert y.shape == SMALL.shape and y.dtype == dtype
    # rows are independent calls of the 1-D model (test_fixed_output.py:81-96)
    y = np.load(out / "fixed_3tap" / "case_000_small__simple_lp_fixed_3tap_y_u8.npy")
    for r in (0, 2):
        assert np.array_equal(y[r], fir_1d_fixed_golden(SMALL[r].tolist(), h_coeff_3tap_map["simple_lp"]))
        assert np.array_equal(y[r], onp.fir1d_fixed(SMALL[r], h_coeff_3tap_map["simple_lp"]))
    yi = np.load(out / "ideal_5tap" / "case_000_small__sharpen_ideal_5tap_y_f64.npy")
    for r in (1, 3):
        assert np.allclose(yi[r], fir_1d_ideal(SMALL[r].tolist(), h_coeff_5tap_map["sharpen"]))
    # idempotent skip unless overwrite (gen_fixed_output.py:95-96)
    assert gf.generate_fixed_3tap_output_vector(input_dir=inp, output_dir=out) == 0
    assert gf.generate_fixed_3tap_output_vector(input_dir=inp, output_dir=out, overwrite=True) == 4
    # keyword plumbing of the bit widths
    out2 = tmp_path / "q115"
    assert gf.generate_fixed_5tap_output_vector(input_dir=inp, output_dir=out2, frac_bits=12, acc_bits=16, coeff_bits=16) == 4
    y16 = np.load(out2 / "fixed_5tap" / "case_000_small__sharpen_fixed_5tap_y_u8.npy")
    assert np.array_equal(y16, onp.fir1d_fixed(SMALL, h_coeff_5tap_map["sharpen"], 12, 16, 16))


def test_pipeline_driver_model_stages(tmp_path, monkeypatch):
    import pipeline_fir_1d as pipe
    from fir_1d.sim.vector import gen_fixed_output as gf, gen_ideal_output as gi
    inp, out = tmp_path / "input", tmp_path / "output"
    inp.mkdir()
    np.save(inp / "case_000_small_x_u8.npy", SMALL)
    monkeypatch.setenv("FIR_VECTOR_ROOT", str(tmp_path))          # one vector tree for every stage (_batched.vector_root)
    res = pipe.run_pipeline(tap="all", overwrite_vectors=True, skip_input=True, skip_ideal=False, skip_fixed=False,
                            skip_report=True, skip_restore=True, restore_kind="all", ideal_policy="clip",
                            overwrite_images=False, strict_report=False, strict_restore=False, top_k=5)
    assert res == {"selected_taps": ["3", "5"], "ideal_counts": {"ideal_3tap": 4, "ideal_5tap": 4},
                   "fixed_counts": {"fixed_3tap": 4, "fixed_5tap": 4}}
    with pytest.raises(ImportError, match="FIR_REFERENCE_ROOT"):
        pipe.run_pipeline(tap="3", overwrite_vectors=False, skip_input=False, skip_ideal=True, skip_fixed=True,
                          skip_report=True, skip_restore=True, restore_kind="all", ideal_policy="clip",
                          overwrite_images=False, strict_report=False, strict_restore=False, top_k=5)


# ---- fused compare metrics ------------------------------------------
def numpy_metrics(y_ideal, y_fixed):
    """The reference's `_compute_metrics` (gen_3tap_compare_report.py:67-112) restated on arrays."""
    diff = y_fixed.astype(np.float64) - y_ideal.astype(np.float64)
    n = diff.size
    return {"num_samples": n, "max_abs_err": float(np.abs(diff).max()), "mae": float(np.abs(diff).mean()),
            "rmse": float(np.sqrt(np.mean(np.square(diff)))), "mean_err": float(diff.mean()),
            "sat_low_ratio": float(np.mean(y_fixed == 0)), "sat_high_ratio": float(np.mean(y_fixed == 255)),
            "sat_ratio": float(np.mean(y_fixed == 0)) + float(np.mean(y_fixed == 255)),
            "clip_needed_ratio": float(np.mean((y_ideal < 0.0) | (y_ideal > 255.0)))}


def test_fused_compare_matches_file_based_metrics(images_small, images_inputs, manifest):
    """Against the reference pipeline's own saved ideal / fixed outputs for
    the two 64x64 cases, and against oracle outputs for a larger image."""
    import torch
    import firgpu
    stems = sorted({k.split("/")[0] for k in images_small.keys()})
    for stem in stems:
        x = images_small[f"{stem}/x"]
        for tap in ("3tap", "5tap"):
            for name, h in manifest[f"coeff_{tap}"].items():
                want = numpy_metrics(images_small[f"{stem}/{tap}/{name}/ideal"], images_small[f"{stem}/{tap}/{name}/fixed"])
                for got in (firgpu.compare_metrics(x, h), firgpu.compare_metrics(torch.from_numpy(x).cuda(), h)):
                    assert got["num_samples"] == want["num_samples"]
                    for key in ("sat_low_ratio", "sat_high_ratio", "clip_needed_ratio", "max_abs_err"):
                        assert got[key] == want[key], (stem, tap, name, key)
                    for key in ("mae", "rmse", "mean_err", "sat_ratio"):
                        assert got[key] == pytest.approx(want[key], rel=1e-11, abs=1e-13), (stem, tap, name, key)
    # the documented worst case: case_004 (64x64 noise) x edge  (fir_1d/docs/*_compare_analysis_v1.md)
    noise = images_small["case_004_img_005_64x64_noise_gray/x"]
    assert firgpu.compare_metrics(noise, manifest["coeff_3tap"]["edge"])["rmse"] == pytest.approx(36.4074, abs=5e-5)
    assert firgpu.compare_metrics(noise, manifest["coeff_5tap"]["edge"])["rmse"] == pytest.approx(13.4034, abs=5e-5)
    x = images_inputs["case_001_img_002_640x762_gray"]
    for h, fb, ab in ((manifest["coeff_5tap"]["sharpen"], 12, 32), (manifest["coeff_3tap"]["edge"], 12, 16), ([0.3, 0.3], 9, 32)):
        want = numpy_metrics(onp.fir1d_ideal(x, h), onp.fir1d_fixed(x, h, fb, ab, 16))
        got = firgpu.compare_metrics(x, h, fb, ab, 16)
        for key, v in want.items():
            assert got[key] == pytest.approx(v, rel=1e-11, abs=1e-13), key
    assert firgpu.compare_metrics(np.zeros((0, 5), np.uint8), [1.0])["num_samples"] == 0


def test_fused_compare_rows_and_summary(tmp_path, images_small, manifest):
    from fir_1d.sim.vector import gen_fused_compare as fused
    for stem in sorted({k.split("/")[0] for k in images_small.keys()}):
        np.save(tmp_path / f"{stem}_x_u8.npy", images_small[f"{stem}/x"])
    rows = fused.fused_compare_rows(tmp_path, "5")
    assert [r["key"] for r in rows] == sorted(r["key"] for r in rows) and len(rows) == 8
    assert set(rows[0]) >= {"key", "case_stem", "coeff_name", "height", "width", "num_samples", "max_abs_err", "mae", "rmse",
                            "mean_err", "sat_low_ratio", "sat_high_ratio", "sat_ratio", "clip_needed_ratio"}
    want_rows = []
    for r in rows:
        want_rows.append(numpy_metrics(images_small[f"{r['case_stem']}/5tap/{r['coeff_name']}/ideal"],
                                       images_small[f"{r['case_stem']}/5tap/{r['coeff_name']}/fixed"]))
    summary = fused.summarize(rows)
    assert summary["num_cases"] == 8 and summary["num_samples_total"] == 8 * 64 * 64
    assert summary["avg_rmse"] == pytest.approx(np.mean([w["rmse"] for w in want_rows]), rel=1e-11)
    assert summary["max_max_abs_err"] == max(w["max_abs_err"] for w in want_rows)


# ---- extended entry points: causal / streaming (C++ Fir1D) and signed int16 ---------
def test_fir1d_class_matches_reference_cpp_vectors():
    """fir_1d.model.python.fir_1d_stream.Fir1D against vectors produced by the
    reference's own C++ Fir1D (tests/golden/cpp_fir1d_cases.npz): bit-identical,
    whole-signal, sample-by-sample and in ragged chunks."""
    import torch
    import firgpu
    from conftest import GOLDEN
    from fir_1d.model.python.fir_1d_stream import Fir1D
    blob = np.load(GOLDEN / "cpp_fir1d_cases.npz")
    rng = np.random.default_rng(0)
    for i, name in enumerate(blob["names"]):
        x, h = blob[f"{name}/x"], blob[f"{name}/h"]
        assert np.array_equal(Fir1D.fir_1d_idel(x, h), blob[f"{name}/full"]), name
        f = Fir1D(h)
        assert f.taps() == h.tolist()
        cuts = np.sort(rng.integers(0, x.size + 1, size=3))
        chunks = np.split(x, cuts)
        y = np.concatenate([f.process_block(c) for c in chunks])
        assert np.array_equal(y, blob[f"{name}/stream"]), name
        if i < 3:
            f.reset()
            assert [f.process_sample(v) for v in x[:20]] == blob[f"{name}/stream"][:20].tolist()
        # device tensors, with the history exposed as real neighbour samples
        if x.size > len(h):
            xt = torch.from_numpy(x).cuda()
            k = len(h) - 1
            y_tail = firgpu.fir1d_ideal(xt[k:], h.tolist(), centre=0, raw=True, halo_left_valid=k)
            assert np.array_equal(y_tail.cpu().numpy(), blob[f"{name}/stream"][k:]), name
    assert Fir1D().process_sample(3.0) == 0.0 and Fir1D.fir_1d_idel([], [1.0]).size == 0


def test_signed_int16_and_alignment_modes_against_oracle():
    """in/out ranges and the tap alignment as parameters:
    our specification, equal to the reference model at the defaults."""
    import torch
    import firgpu
    from firgpu import _lib
    rng = np.random.default_rng(12)
    x = rng.integers(-32768, 32768, size=(3, 20_001)).astype(np.int16)
    xt = torch.from_numpy(x).cuda()
    h = [0.4, -0.3, 0.2, 0.1, -0.05]
    full = (-32768, 32767)
    for centre in (None, 0, 1, 4):
        for in_r, out_r in ((full, full), ((0, 255), full), ((-100, 100), (-7, 300)), ((0, 255), (0, 255))):
            for frac, acc in ((12, 32), (15, 48), (12, 24)):
                want = onp.fir1d_fixed(x, h, frac, acc, 16, centre=centre, in_range=in_r, out_range=out_r, out_dtype=np.int16)
                got = firgpu.fir1d_fixed(xt, h, frac, acc, 16, centre=centre, in_range=in_r, out_range=out_r, out_dtype=np.int16)
                assert np.array_equal(got.cpu().numpy(), want), (centre, in_r, out_r, frac, acc)
                host = firgpu.fir1d_fixed(x, h, frac, acc, 16, centre=centre, in_range=in_r, out_range=out_r, out_dtype=np.int16)
                assert np.array_equal(host, want), (centre, in_r, out_r, frac, acc, "host")
    # defaults through the _ex entry point are the reference model on the fast kernel
    y = firgpu.fir1d_fixed(xt, h, 12, 32, 16, centre=2, in_range=(0, 255), out_range=(0, 255))
    assert _lib.last_path() == _lib.PATH_STAGED and np.array_equal(y.cpu().numpy(), onp.fir1d_fixed(x, h))
    assert firgpu.fir1d_fixed(xt, h, in_range=full, out_range=full).dtype == torch.int16
    # off-centre alignments run on the centred fast kernels with zero taps added in front / behind
    for centre in (0, 1, 4):
        for kw in ({}, {"in_range": full, "out_range": full}):
            y = firgpu.fir1d_fixed(xt, h, 15, 32, 16, centre=centre, out_dtype=np.int16, **kw)
            assert _lib.last_path() == _lib.PATH_STAGED, (centre, kw)
            assert np.array_equal(y.cpu().numpy(), onp.fir1d_fixed(x, h, 15, 32, 16, centre=centre, out_dtype=np.int16, **kw))
    h17 = rng.uniform(-0.2, 0.2, size=17).tolist()                       # causal 17 taps -> centred 33-tap bucket
    y = firgpu.fir1d_fixed(xt, h17, 12, 32, 16, centre=0, out_dtype=np.int16)
    assert _lib.last_path() == _lib.PATH_STAGED
    assert np.array_equal(y.cpu().numpy(), onp.fir1d_fixed(x, h17, 12, 32, 16, centre=0, out_dtype=np.int16))
    # ideal model with alignment / clamp range
    for centre in (0, 3):
        want = onp.fir1d_ideal(x, h, centre=centre, in_range=(-50, 900))
        assert np.array_equal(firgpu.fir1d_ideal(xt, h, centre=centre, in_range=(-50, 900)).cpu().numpy(), want)
        assert np.array_equal(firgpu.fir1d_ideal(x, h, centre=centre, in_range=(-50, 900)), want)
    with pytest.raises(firgpu._lib.FirGpuError):
        firgpu.fir1d_fixed(xt, h, centre=5)


def test_ideal_fast_kernel_bit_identical():
    """fir1d_ideal_fast_kernel (uint8 / int16 -> float64): every tap count,
    ragged lengths, several rows, shard halos — bit-identical to the oracle
    (and therefore to the reference's float64 model)."""
    import torch
    import firgpu
    from firgpu import _lib, shard
    rng = np.random.default_rng(31)
    for dtype in (np.uint8, np.int16):
        for n_taps in range(1, 10):
            h = rng.uniform(-8, 8, size=n_taps).tolist()
            for shape in ((1, 1), (1, 7), (3, 2048), (2, 2050), (5, 4098), (1, 100_003), (7, 1501), (4, 515), (6, 1027)):
                x = rng.integers(0, 256, size=shape).astype(dtype)
                if dtype == np.int16:
                    x[rng.random(shape) < 0.05] = -999
                    x[rng.random(shape) < 0.05] = 12345
                got = firgpu.fir1d_ideal(torch.from_numpy(x).cuda(), h)
                assert _lib.last_path() == _lib.PATH_STAGED, (dtype, n_taps, shape)
                assert np.array_equal(got.cpu().numpy(), onp.fir1d_ideal(x, h)), (dtype, n_taps, shape)
    # odd widths and pitches: the kernel shifts every row's tile grid to keep its 16-byte stores aligned; column
    # views of a wider image give rows whose first sample / first output sit at every alignment
    wide = rng.integers(0, 256, size=(9, 4499)).astype(np.uint8)
    wide_t = torch.from_numpy(wide).cuda()
    h = [0.1, -0.2, 0.3, 0.4, 0.05]
    for out_dtype, exact in ((np.float64, True), (np.float32, False)):
        for c0 in (0, 1, 2, 3, 5):
            for width in (4499 - 5, 1920, 777):
                view = wide_t[:, c0:c0 + width]
                out = torch.full((9, 4499), -7.0, dtype=torch.float64 if exact else torch.float32, device="cuda")
                firgpu.fir1d_ideal(view, h, out_dtype=out_dtype, out=out[:, c0:c0 + width])
                assert _lib.last_path() == _lib.PATH_STAGED
                want = onp.fir1d_ideal(wide[:, c0:c0 + width], h)
                got = out.cpu().numpy()
                if exact:
                    assert np.array_equal(got[:, c0:c0 + width], want), (c0, width)
                else:
                    assert np.abs(got[:, c0:c0 + width] - want).max() <= 1e-6 * 255 * np.abs(h).sum()
                assert (got[:, :c0] == -7.0).all() and (got[:, c0 + width:] == -7.0).all(), (c0, width)   # nothing outside the view
    x = rng.integers(-300, 600, size=50_000).astype(np.int16)
    want = onp.fir1d_ideal(x, h)
    for p in shard.plan_segments(x.size, 3, 5, 1, align=64):
        buf, body = shard.scatter_segment(x, p, "cuda")
        y = firgpu.fir1d_ideal(body, h, halo_left_valid=p.before, halo_right_valid=p.after)
        assert np.array_equal(y.cpu().numpy(), want[p.start:p.stop])


def test_largest_reference_image_by_hash(image_large, manifest):
    """4499 x 2999 (odd width, unaligned pitch): all 8 fixed and 8 ideal files
    of the reference pipeline, by SHA-256 — with the other six images
    (test_gpu_fir1d_fixed.py, above) the GPU path reproduces the complete
    output set of `python pipeline_fir_1d.py`."""
    from fir_1d.sim.vector.gen_fixed_output import _run_fixed_rowwise
    from fir_1d.sim.vector.gen_ideal_output import _run_ideal_rowwise
    stem, x = image_large
    for tap in ("3tap", "5tap"):
        for name, h in manifest[f"coeff_{tap}"].items():
            assert sha(_run_fixed_rowwise(x, h, frac_bits=12, acc_bits=32, coeff_bits=16)) == manifest["fixed"][f"{stem}__{name}_{tap}"], (name, tap)
            assert sha(_run_ideal_rowwise(x, h)) == manifest["ideal"][f"{stem}__{name}_{tap}"], (name, tap)


def test_published_accuracy_tables_from_inputs_alone(tmp_path, images_inputs, image_large, compare_rows):
    """SURVEY 8(f) #2 end to end: the fused compare path on the 7 input images
    reproduces every row the reference's report code computes from its saved
    ideal / fixed files, and the aggregate numbers printed in the reference's
    docs (fir_1d/docs/fir_1d_3tap_compare_analysis_v1.md:44-83,
    fir_1d_5tap_compare_analysis_v1.md:44-85, fir_1d_fixed_3tap_vs_5tap_comparison_v1.md:20-21)."""
    from fir_1d.sim.vector import gen_fused_compare as fused
    for stem in images_inputs.keys():
        np.save(tmp_path / f"{stem}_x_u8.npy", images_inputs[stem])
    np.save(tmp_path / f"{image_large[0]}_x_u8.npy", image_large[1])
    out = {}
    for tap in ("3", "5"):
        rows = fused.fused_compare_rows(tmp_path, tap)
        want_rows = compare_rows[f"{tap}tap"]["rows"]
        assert [r["key"] for r in rows] == [w["key"] for w in want_rows] and len(rows) == 28
        for r, w in zip(rows, want_rows):
            assert (r["height"], r["width"], r["num_samples"]) == (w["height"], w["width"], w["num_samples"])
            for key in ("max_abs_err", "sat_low_ratio", "sat_high_ratio", "clip_needed_ratio"):
                assert r[key] == w[key], (r["key"], key)
            for key in ("mae", "rmse", "mean_err", "sat_ratio"):
                assert r[key] == pytest.approx(w[key], rel=1e-10, abs=1e-13), (r["key"], key)
        summary = fused.summarize(rows)
        for key, v in compare_rows[f"{tap}tap"]["summary"].items():
            if key in summary:
                assert summary[key] == pytest.approx(v, rel=1e-10), key
        n_total = sum(r["num_samples"] for r in rows)
        by_coeff = {c: float(np.mean([r["rmse"] for r in rows if r["coeff_name"] == c])) for c in ("edge", "moving_avg", "sharpen", "simple_lp")}
        out[tap] = dict(summary=summary, w_mae=sum(r["num_samples"] * r["mae"] for r in rows) / n_total,
                        w_rmse=sum(r["num_samples"] * r["rmse"] for r in rows) / n_total, by_coeff=by_coeff,
                        worst=max(rows, key=lambda r: r["rmse"]))
    r4 = lambda v: round(v, 4)
    s3, s5 = out["3"], out["5"]
    assert (r4(s3["summary"]["avg_mae"]), r4(s3["summary"]["avg_rmse"]), r4(s3["summary"]["avg_mean_err"]), s3["summary"]["max_max_abs_err"]) == (2.4078, 5.8063, 2.2903, 255.0)
    assert (r4(s3["w_mae"]), r4(s3["w_rmse"])) == (0.6592, 2.6526)
    assert {k: r4(v) for k, v in s3["by_coeff"].items()} == {"edge": 21.9354, "moving_avg": 0.2432, "sharpen": 0.7780, "simple_lp": 0.2686}
    assert s3["worst"]["key"] == "case_004_img_005_64x64_noise_gray__edge" and r4(s3["worst"]["rmse"]) == 36.4074
    assert (r4(s5["summary"]["avg_mae"]), r4(s5["summary"]["avg_rmse"]), r4(s5["summary"]["avg_mean_err"]), s5["summary"]["max_max_abs_err"]) == (1.1302, 2.9380, 0.8871, 118.0625)
    assert (r4(s5["w_mae"]), r4(s5["w_rmse"])) == (0.3243, 1.2739)
    assert {k: r4(v) for k, v in s5["by_coeff"].items()} == {"edge": 8.1020, "moving_avg": 0.2570, "sharpen": 3.1345, "simple_lp": 0.2586}
    assert s5["worst"]["key"] == "case_004_img_005_64x64_noise_gray__edge" and r4(s5["worst"]["rmse"]) == 13.4034
    assert round(100 * (s5["summary"]["avg_mae"] / s3["summary"]["avg_mae"] - 1), 2) == -53.06
    assert round(100 * (s5["summary"]["avg_rmse"] / s3["summary"]["avg_rmse"] - 1), 2) == -49.40


def test_causal_raw_float64_model_on_the_fast_kernel():
    """The C++ Fir1D model (causal alignment, raw float64 samples,
    fir_1d/model/cpp/fir_1d_ref.cpp:16-63) runs on the register kernel for up to
    nine taps: bit-identical to the oracle for every tap count, ragged lengths,
    several rows, views at odd offsets, and block-wise streaming with carried
    state equals one call over the whole signal."""
    import torch
    import firgpu
    from firgpu import _lib
    from fir_1d.model.python.fir_1d_stream import Fir1D
    rng = np.random.default_rng(53)
    for n_taps in range(1, 10):
        h = rng.uniform(-3, 3, size=n_taps).tolist()
        for shape in ((1, 5), (1, 70_001), (3, 2049), (4, 1501)):
            x = rng.normal(0, 1e3, size=shape)
            got = firgpu.fir1d_ideal(torch.from_numpy(x).cuda(), h, centre=0, raw=True)
            assert _lib.last_path() == _lib.PATH_STAGED, (n_taps, shape)
            assert np.array_equal(got.cpu().numpy(), onp.fir1d_ideal(x, h, centre=0, raw=True)), (n_taps, shape)
    big = rng.normal(0, 50, size=200_003)
    h = [0.5, -0.25, 0.125, 0.3, -0.7]
    want = onp.fir1d_ideal(big, h, centre=0, raw=True)
    big_t = torch.from_numpy(big).cuda()
    for c0 in (1, 2, 3):                                              # misaligned views
        got = firgpu.fir1d_ideal(big_t[c0:c0 + 100_000], h, centre=0, raw=True)
        assert np.array_equal(got.cpu().numpy(), onp.fir1d_ideal(big[c0:c0 + 100_000], h, centre=0, raw=True)), c0
    f = Fir1D(h)
    parts = [f.process_block(big[a:b]) for a, b in ((0, 7), (7, 50_000), (50_000, 50_001), (50_001, big.size))]
    assert np.array_equal(np.concatenate(parts), want)


def test_interleaved_iq_floating_point_model_on_the_fast_kernel():
    """dilation = 2 (interleaved I/Q, real taps) of the floating-point model on
    the register kernel: equal to the model applied to z[0::2] and z[1::2]
    separately, bit for bit in float64, for every tap count, even and odd
    scalar counts, several rows and float containers."""
    import torch
    import firgpu
    from firgpu import _lib
    rng = np.random.default_rng(61)
    for dtype in (np.uint8, np.int16, np.float32):
        for n_taps in range(1, 10):
            h = rng.uniform(-2, 2, size=n_taps).tolist()
            for shape in ((1, 2), (1, 10_000), (3, 4098), (2, 4501), (5, 1026)):
                z = rng.integers(0, 256, size=shape).astype(dtype)
                if dtype == np.int16:
                    z[rng.random(shape) < 0.05] = -77
                if dtype == np.float32:
                    z = z + rng.uniform(-0.5, 0.5, size=shape).astype(np.float32)
                want = onp.fir1d_ideal(z, h, dilation=2)
                if shape[1] % 2 == 0:                      # the definition: I and Q filtered separately
                    sep = np.empty_like(want)
                    sep[..., 0::2] = onp.fir1d_ideal(z[..., 0::2], h)
                    sep[..., 1::2] = onp.fir1d_ideal(z[..., 1::2], h)
                    assert np.array_equal(want, sep)
                got = firgpu.fir1d_ideal(torch.from_numpy(z).cuda(), h, dilation=2)
                if dtype != np.float32 or z.size >= 4096:
                    assert _lib.last_path() == _lib.PATH_STAGED, (dtype, n_taps, shape)
                assert np.array_equal(got.cpu().numpy(), want), (dtype, n_taps, shape)
                got32 = firgpu.fir1d_ideal(torch.from_numpy(z).cuda(), h, dilation=2, out_dtype=np.float32).cpu().numpy()
                assert np.abs(got32 - want).max() <= 1e-6 * 255 * np.abs(h).sum() + 1e-30
