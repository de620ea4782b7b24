"""The oracle (oracle/: scalar Python, vectorised NumPy, plain C) against the
golden vectors made by the UNMODIFIED reference (tests/golden/make_golden.py).
This is what pins the checker that the GPU parity tests rely on."""
from __future__ import annotations

import hashlib

import numpy as np
import pytest

from oracle import c_oracle, fir_oracle_np as onp, fir_oracle_py as opy


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_case_file_covers_reference_kats(golden_cases):
    by_name = {c["name"]: c for c in golden_cases}
    # fir_1d/sim/tests/test_1d_fixed.py:9-16 and test_1d_ideal.py:8-16
    assert by_name["kat_q412_center"]["y_fixed"].tolist() == [10, 20, 30, 28]
    assert by_name["kat_q412_center"]["y_ideal"].tolist() == [10.0, 20.0, 30.0, 27.5]
    # test_1d_fixed.py:19-24
    assert by_name["kat_round_clamp_in"]["y_fixed"].tolist() == [0, 1, 2, 255, 255]
    # SURVEY §8(c)5 extras, produced by the reference itself
    assert by_name["wrap_acc16"]["y_fixed"].tolist() == [0, 4, 7]
    assert by_name["wrap_acc21"]["y_fixed"].tolist() == [255, 100, 7]
    assert by_name["c32_acc32"]["y_fixed"].tolist() == [0, 8, 1]
    assert by_name["c32_acc64"]["y_fixed"].tolist() == [0, 8, 0]
    assert by_name["q115_simple_lp5"]["y_fixed"].tolist() == [11, 20, 30, 40, 50, 60, 64, 51]
    assert by_name["i16_container"]["y_fixed"].tolist() == [0, 0, 0, 128, 255, 255, 255, 0]
    assert len(golden_cases) >= 280


def test_scalar_oracle_matches_reference_outputs(golden_cases):
    for c in golden_cases:
        y = opy.fir1d_fixed(c["x"].tolist(), c["h"], c["frac_bits"], c["acc_bits"], c["coeff_bits"])
        assert y.dtype == np.uint8 and np.array_equal(y, c["y_fixed"]), c["name"]
        yi = np.array(opy.fir1d_ideal(c["x"].tolist(), c["h"]), dtype=np.float64)
        assert np.array_equal(yi, c["y_ideal"]), c["name"]      # bit-identical float64


def test_numpy_oracle_matches_reference_outputs(golden_cases):
    for c in golden_cases:
        y = onp.fir1d_fixed(c["x"], c["h"], c["frac_bits"], c["acc_bits"], c["coeff_bits"])
        assert np.array_equal(y, c["y_fixed"]), c["name"]
        assert np.array_equal(onp.fir1d_ideal(c["x"], c["h"]), c["y_ideal"]), c["name"]


def test_c_oracle_matches_reference_outputs(golden_cases):
    for c in golden_cases:
        if c["x"].size == 0:
            continue
        hq = onp.quantize_taps(c["h"], c["frac_bits"], c["coeff_bits"])
        x = c["x"] if c["x"].dtype != np.int64 else c["x"].astype(np.int32)
        y = c_oracle.fir1d_fixed(x, hq, c["frac_bits"], c["acc_bits"])
        assert np.array_equal(y, c["y_fixed"]), c["name"]
        assert np.array_equal(c_oracle.fir1d_ideal(x, c["h"]), c["y_ideal"]), c["name"]


def _coeffs(manifest):
    for tap in ("3tap", "5tap"):
        for name, h in manifest[f"coeff_{tap}"].items():
            yield tap, name, h


def test_small_images_every_output(images_small, manifest):
    """Both 64x64 BMP cases: all 16 fixed and 16 ideal reference files."""
    stems = sorted({k.split("/")[0] for k in images_small.keys()})
    assert len(stems) == 2
    for stem in stems:
        x = images_small[f"{stem}/x"]
        for tap, name, h in _coeffs(manifest):
            want_f = images_small[f"{stem}/{tap}/{name}/fixed"]
            want_i = images_small[f"{stem}/{tap}/{name}/ideal"]
            assert np.array_equal(onp.fir1d_fixed(x, h), want_f)
            assert np.array_equal(onp.fir1d_ideal(x, h), want_i)
            hq = onp.quantize_taps(h, 12, 16)
            assert np.array_equal(c_oracle.fir1d_fixed(x, hq, 12, 32), want_f)
            assert np.array_equal(c_oracle.fir1d_ideal(x, h), want_i)
            assert sha(want_f) == manifest["fixed"][f"{stem}__{name}_{tap}"]


def test_full_images_by_hash(images_inputs, manifest):
    """The six images up to 1.1 Mpx: oracle output must hash to the SHA-256 of
    the reference pipeline's full-size output files (96 files)."""
    checked = 0
    for stem in images_inputs.keys():
        x = images_inputs[stem]
        assert sha(x) == manifest["inputs"][stem]["sha256"]
        for tap, name, h in _coeffs(manifest):
            hq = onp.quantize_taps(h, 12, 16)
            assert sha(c_oracle.fir1d_fixed(x, hq, 12, 32, threads=4)) == manifest["fixed"][f"{stem}__{name}_{tap}"]
            assert sha(c_oracle.fir1d_ideal(x, h, threads=4)) == manifest["ideal"][f"{stem}__{name}_{tap}"]
            checked += 2
    assert checked == 96


def test_large_image_rows(image_crops, manifest):
    """16 rows of the 4499x2999 case (odd width, unaligned pitch)."""
    stem = "case_005_img_006_4499x2999_gray"
    x = image_crops[f"{stem}/x"]
    assert x.shape == (16, 4499)
    for tap, name, h in _coeffs(manifest):
        want = image_crops[f"{stem}/{tap}/{name}/fixed"]
        assert np.array_equal(onp.fir1d_fixed(x, h), want)
        assert np.array_equal(c_oracle.fir1d_fixed(x, onp.quantize_taps(h, 12, 16), 12, 32), want)


def test_iq_is_two_independent_filters():
    rng = np.random.default_rng(5)
    z = rng.integers(-300, 600, size=2 * 173).astype(np.int16)
    h = [1 / 16, 4 / 16, 6 / 16, 4 / 16, 1 / 16]
    want = opy.fir1d_fixed_iq(z.tolist(), h, 15, 32, 16)
    assert np.array_equal(onp.fir1d_fixed_iq(z, h, 15, 32, 16), want)
    hq = onp.quantize_taps(h, 15, 16)
    assert np.array_equal(c_oracle.fir1d_fixed(z, hq, 15, 32, dilation=2), want)


def test_2d_oracles_agree_and_reduce_to_1d():
    """fir_2d is our specification (parity unpinned); the three restatements
    must agree with each other and, for single-row / single-column kernels,
    with the reference-pinned 1-D model."""
    rng = np.random.default_rng(11)
    x = rng.integers(0, 256, size=(13, 17)).astype(np.uint8)
    h2 = rng.uniform(-0.4, 0.4, size=(3, 5))
    a = opy.fir2d_fixed(x.tolist(), h2.tolist())
    b = onp.fir2d_fixed(x, h2)
    c = c_oracle.fir2d_fixed(x, onp.quantize_taps(h2, 12, 16), 12, 32)
    assert np.array_equal(a, b) and np.array_equal(a, c)
    assert np.array_equal(opy.fir2d_ideal(x.tolist(), h2.tolist()), onp.fir2d_ideal(x, h2))
    assert np.array_equal(c_oracle.fir2d_ideal(x, h2), onp.fir2d_ideal(x, h2))
    h1 = [0.25, 0.5, 0.25]
    assert np.array_equal(onp.fir2d_fixed(x, [h1]), onp.fir1d_fixed(x, h1))
    assert np.array_equal(onp.fir2d_fixed(x, [[v] for v in h1]), onp.fir1d_fixed(x.T, h1).T)
    sep = onp.fir2d_separable_fixed(x, h1, h1)
    again = onp.fir1d_fixed(onp.fir1d_fixed(x, h1).T, h1).T
    assert np.array_equal(sep, again)


@pytest.mark.parametrize("acc_bits", [8, 16, 31, 32, 33, 48, 64])
def test_numpy_and_c_oracle_fuzz_against_scalar(acc_bits):
    rng = np.random.default_rng(acc_bits)
    for _ in range(20):
        n_taps = int(rng.integers(1, 12))
        cb = int(rng.choice([8, 16, 32]))
        fb = int(rng.integers(1, cb + 4))
        n = int(rng.integers(1, 80))
        hi = min(8.0, ((1 << (cb - 1)) - 1) / (1 << fb))
        lo = max(-8.0, -(1 << (cb - 1)) / (1 << fb))
        h = rng.uniform(lo, hi, size=n_taps).tolist()
        x = rng.integers(-50, 400, size=n).astype(np.int16)
        for dil in (1, 2, 3):
            want = opy.fir1d_fixed(x.tolist(), h, fb, acc_bits, cb, dilation=dil)
            assert np.array_equal(onp.fir1d_fixed(x, h, fb, acc_bits, cb, dilation=dil), want)
            hq = onp.quantize_taps(h, fb, cb)
            assert np.array_equal(c_oracle.fir1d_fixed(x, hq, fb, acc_bits, dilation=dil), want)


def test_causal_raw_model_matches_reference_cpp_vectors():
    """tests/golden/cpp_fir1d_cases.npz = Fir1D::fir_1d_idel / process_sample of
    the reference's own C++ (compiled by oracle/Makefile, make_golden_cpp.py):
    the oracle's causal raw-sample model must be bit-identical."""
    from conftest import GOLDEN
    blob = np.load(GOLDEN / "cpp_fir1d_cases.npz")
    assert len(blob["names"]) == 61
    assert np.allclose(blob["pdf_p13/full"], [1 / 3, 1, 2, 3, 4, 3, 5 / 3])      # course PDF p.13 example
    for name in blob["names"]:
        x, h = blob[f"{name}/x"], blob[f"{name}/h"].tolist()
        assert np.array_equal(onp.fir1d_full_causal(x, h), blob[f"{name}/full"]), name
        assert np.array_equal(onp.fir1d_ideal(x, h, centre=0, raw=True), blob[f"{name}/stream"]), name
        assert np.array_equal(np.array(opy.fir1d_ideal(x.tolist(), h, centre=0, raw=True)), blob[f"{name}/stream"]), name


def test_parametrised_oracle_reduces_to_reference_at_defaults(golden_cases):
    for c in golden_cases[:80]:
        y = onp.fir1d_fixed(c["x"], c["h"], c["frac_bits"], c["acc_bits"], c["coeff_bits"], centre=len(c["h"]) // 2,
                            in_range=(0, 255), out_range=(0, 255))
        assert np.array_equal(y, c["y_fixed"]), c["name"]
    rng = np.random.default_rng(1)
    x = rng.integers(-40000, 40000, size=300)
    h = [0.4, -0.3, 0.2, 0.1]
    for centre in range(4):
        for rng_in, rng_out in (((0, 255), (0, 255)), ((-32768, 32767), (-32768, 32767)), ((-100, 100), (-5, 300))):
            a = opy.fir1d_fixed(x.tolist(), h, 12, 32, 16, centre=centre, in_range=rng_in, out_range=rng_out)
            b = onp.fir1d_fixed(x, h, 12, 32, 16, centre=centre, in_range=rng_in, out_range=rng_out, out_dtype=np.int64)
            assert np.array_equal(np.asarray(a, dtype=np.int64), b), (centre, rng_in, rng_out)


def test_largest_image_by_hash(image_large, manifest):
    """The 4499x2999 case in full: with it, the oracle reproduces all 56 fixed
    and 56 ideal output files of the reference pipeline."""
    stem, x = image_large
    assert sha(x) == manifest["inputs"][stem]["sha256"]
    for tap, name, h in _coeffs(manifest):
        hq = onp.quantize_taps(h, 12, 16)
        assert sha(c_oracle.fir1d_fixed(x, hq, 12, 32, threads=8)) == manifest["fixed"][f"{stem}__{name}_{tap}"]
        assert sha(c_oracle.fir1d_ideal(x, h, threads=8)) == manifest["ideal"][f"{stem}__{name}_{tap}"]


def test_reference_report_numbers_are_the_published_ones(compare_rows):
    """compare_rows.json (reference code on reference outputs) carries the
    accuracy tables the reference's docs publish
    (fir_1d/docs/fir_1d_{3,5}tap_compare_analysis_v1.md)."""
    s3, s5 = compare_rows["3tap"]["summary"], compare_rows["5tap"]["summary"]
    assert s3["num_cases"] == s5["num_cases"] == 28 and s3["num_samples_total"] == 67_975_252
    assert (round(s3["avg_mae"], 4), round(s3["avg_rmse"], 4), round(s3["avg_mean_err"], 4), s3["max_max_abs_err"]) == (2.4078, 5.8063, 2.2903, 255.0)
    assert (round(s5["avg_mae"], 4), round(s5["avg_rmse"], 4), round(s5["avg_mean_err"], 4), s5["max_max_abs_err"]) == (1.1302, 2.9380, 0.8871, 118.0625)


def test_complex_tap_oracle_anchors():
    """oracle/fir_oracle_np.py::fir1d_fixed_complex is OUR specification (the reference has no complex product); what
    can be anchored is anchored: purely real taps = the real-tap I/Q filter (two independent runs of the reference-pinned
    1-D model), h = [j] is a rotation by 90 degrees, linearity in the taps before the rounding, and a scalar restatement."""
    rng = np.random.default_rng(5)
    z = rng.integers(-300, 600, size=2 * 257).astype(np.int16)
    h = [0.1, 0.2, 0.4, 0.2, 0.1]
    assert np.array_equal(onp.fir1d_fixed_complex(z, h, 12, 32, 16, out_dtype=np.uint8), onp.fir1d_fixed_iq(z, h, 12, 32, 16))
    full = (-32768, 32767)
    zs = rng.integers(-1000, 1000, size=200).astype(np.int16)
    rot = onp.fir1d_fixed_complex(zs, [1j], 12, 32, 16, in_range=full, out_range=full)
    assert np.array_equal(rot[0::2], -zs[1::2]) and np.array_equal(rot[1::2], zs[0::2])
    # scalar restatement of the definition with unbounded Python ints
    hc = [0.3 + 0.1j, -0.2 + 0.25j, 0.15 - 0.3j, 0.05 + 0.05j]
    hr = onp.quantize_taps([v.real for v in hc], 12, 16).tolist()
    hi = onp.quantize_taps([v.imag for v in hc], 12, 16).tolist()
    n, L, c = zs.size // 2, len(hc), len(hc) // 2
    want = np.empty(zs.size, dtype=np.int64)
    for acc_bits in (32, 14):
        for i in range(n):
            ar = ai = 0
            for k in range(L):
                m = i - (k - c)
                if 0 <= m < n:
                    xr, xi = int(zs[2 * m]), int(zs[2 * m + 1])
                    ar += hr[k] * xr - hi[k] * xi
                    ai += hr[k] * xi + hi[k] * xr
            want[2 * i] = onp.post_accumulate(np.array([ar], dtype=np.int64), 12, acc_bits, *full)[0]
            want[2 * i + 1] = onp.post_accumulate(np.array([ai], dtype=np.int64), 12, acc_bits, *full)[0]
        got = onp.fir1d_fixed_complex(zs, hc, 12, acc_bits, 16, in_range=full, out_range=full)
        assert np.array_equal(got, want.astype(np.int16)), acc_bits
