"""Host-side logic that needs no GPU: the validation / quantisation contract,
container normalisation, the shard planner, and the C-ABI library itself
(loads, exports every symbol include/firgpu.h declares, rejects bad arguments
before touching CUDA, serves empty inputs without a device)."""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from firgpu import _lib, ops, quant, shard
from oracle import fir_oracle_np as onp

ROOT = Path(__file__).resolve().parents[1]


# ---- C ABI -----------------------------------------------------------------------
def declared_functions() -> list[str]:
    text = (ROOT / "include" / "firgpu.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(firgpu_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = declared_functions()
    assert len(names) >= 17
    handle = C.CDLL(str(_lib.LIB_PATH))
    for name in names:
        assert hasattr(handle, name), f"{name} declared in include/firgpu.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes binding and header disagree"
    assert _lib.lib().firgpu_version() == 200


def test_header_has_no_torch_or_cuda_types():
    text = (ROOT / "include" / "firgpu.h").read_text()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    assert "torch" not in code and "cudaStream_t" not in code and "#include <cuda" not in code


def test_abi_rejects_bad_arguments_without_a_device():
    L = _lib.lib()
    x = np.zeros(8, np.uint8)
    y = np.zeros(8, np.uint8)
    hq = np.array([1, 2, 1], np.int32)
    px, py, ph = x.ctypes.data, y.ctypes.data, hq.ctypes.data
    call = L.firgpu_fir1d_fixed
    assert call(px, py, 9, 0, 1, 8, 8, 8, ph, 3, 1, 12, 32, 0, 0, None) == -1      # dtype
    assert call(px, py, 0, 4, 1, 8, 8, 8, ph, 3, 1, 12, 32, 0, 0, None) == -1      # out dtype
    assert call(px, py, 0, 0, -1, 8, 8, 8, ph, 3, 1, 12, 32, 0, 0, None) == -2     # shape
    assert call(px, py, 0, 0, 2, 8, 4, 8, ph, 3, 1, 12, 32, 0, 0, None) == -2      # ld < cols
    assert call(px, py, 0, 0, 1, 8, 8, 8, ph, 0, 1, 12, 32, 0, 0, None) == -3      # taps
    assert call(px, py, 0, 0, 1, 8, 8, 8, ph, 2000, 1, 12, 32, 0, 0, None) == -3
    assert call(px, py, 0, 0, 1, 8, 8, 8, ph, 3, 1, 0, 32, 0, 0, None) == -4       # bits
    assert call(px, py, 0, 0, 1, 8, 8, 8, ph, 3, 0, 12, 32, 0, 0, None) == -5      # dilation
    assert call(None, py, 0, 0, 1, 8, 8, 8, ph, 3, 1, 12, 32, 0, 0, None) == -6    # NULL
    assert call(px, py, 0, 0, 1, 8, 8, 8, ph, 3, 1, 12, 32, -1, 0, None) == -7     # halo
    assert b"halo" in L.firgpu_last_error()
    assert L.firgpu_fir2d_fixed(px, py, 0, 0, 1, 2, 4, 4, 4, 8, 8, ph, 40, 40, 12, 32, 0, 0, None) == -3
    # zero-sized problems are served without a device
    assert call(None, None, 0, 0, 0, 8, 8, 8, ph, 3, 1, 12, 32, 0, 0, None) == 0
    assert _lib.last_path() == _lib.PATH_EMPTY
    assert L.firgpu_fir1d_fixed_host(None, None, 0, 0, 1, 0, 0, 0, ph, 3, 1, 12, 32, -1) == 0


def test_no_cpu_fallback_without_a_device():
    """On a box without CUDA the product path must fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(_lib.FirGpuError) as err:
        ops.fir1d_fixed([10, 20, 30, 40], [0.25, 0.5, 0.25])
    assert err.value.code > 0       # a cudaError_t, not a silently computed result


def test_product_package_never_imports_the_oracle():
    for path in (ROOT / "warmup-fir-filter_b200").rglob("*.py"):
        text = path.read_text()
        assert "oracle" not in re.sub(r"#.*", "", text), f"{path} mentions the oracle"
    for path in (ROOT / "warmup-fir-filter_b200" / "csrc").glob("*.cu*"):
        assert "oracle" not in path.read_text()


# ---- validation / quantisation contract (fir_1d_ref.py:9-33, fir_1d_fixed_ref.py:39-81) ----
@pytest.mark.parametrize("h, pattern", [
    ([], r"must not be empty"),
    ([float("inf")], r"h\[0\].*finite"),
    ([0.5, float("nan")], r"h\[1\].*finite"),
    ([8.0 + 1e-6], r"\|h\| must be <= 8\.0"),
    ([9.0, float("nan")], r"h\[0\]=9\.0: \|h\| must be"),     # index order, not check order
])
def test_tap_validation_messages(h, pattern):
    with pytest.raises(ValueError, match=pattern):
        quant.validate_taps(h)


def test_sample_validation_reports_first_bad_index():
    quant.check_samples_finite(np.array([1, 2, 3], np.int16))
    with pytest.raises(ValueError, match=r"x\[2\]=nan: x must be finite"):
        quant.check_samples_finite(np.array([1.0, 2.0, np.nan, np.inf]))
    with pytest.raises(ValueError, match=r"x\[1\]=-inf"):
        quant.check_samples_finite(np.array([1.0, -np.inf], np.float32))


def test_bit_width_checks_and_order():
    with pytest.raises(ValueError, match=r"Invalid frac_bits=0\. frac_bits must be > 0\."):
        quant.check_bits(0, 0, 12)
    with pytest.raises(ValueError, match=r"Invalid acc_bits=-3\. acc_bits must be > 0\."):
        quant.check_bits(12, -3, 12)
    with pytest.raises(ValueError, match=r"Invalid coeff_bits=12\. coeff_bits must be one of \(8, 16, 32\)\."):
        quant.check_bits(12, 32, 12)
    quant.check_bits(12, 32, 16)


def test_quantisation_matches_reference_rules():
    assert quant.quantize_taps([0.25, 0.5, 0.25], 12, 16).tolist() == [1024, 2048, 1024]
    assert quant.quantize_taps([1 / 3] * 3, 12, 16).tolist() == [1365] * 3
    assert quant.quantize_taps([1 / 5] * 5, 15, 16).tolist() == [6554] * 5
    assert quant.quantize_taps([0.5 + 2.0 ** -13], 12, 16).tolist() == [2048]          # ties to even
    assert quant.quantize_taps([0.5 + 3 * 2.0 ** -13], 12, 16).tolist() == [2050]
    assert quant.quantize_taps([7.999755859375, -8.0], 12, 16).tolist() == [32767, -32768]
    assert quant.quantize_taps([7.5, -7.5, 0.001], 28, 32).tolist() == onp.quantize_taps([7.5, -7.5, 0.001], 28, 32).tolist()
    with pytest.raises(ValueError, match=r"h\[0\]=8\.0: out of Q-format real range \[-8\.0, 7\.999755859375\]"):
        quant.quantize_taps([8.0], 12, 16)
    with pytest.raises(ValueError, match=r"h\[2\]=1\.0: out of Q-format real range"):
        quant.quantize_taps([-1.0, 0, 1.0], 15, 16)                                    # 3-tap edge is not Q1.15
    q2 = quant.quantize_taps(np.array([[0.25, 0.5], [-0.125, 1.0]]), 12, 16)
    assert q2.shape == (2, 2) and q2.dtype == np.int32 and q2.tolist() == [[1024, 2048], [-512, 4096]]


def test_coefficient_bank_quantises_to_survey_values():
    from fir_1d.sim.vector.h_coeff import h_coeff_3tap_map, h_coeff_5tap_map
    assert list(h_coeff_3tap_map) == ["moving_avg", "simple_lp", "edge", "sharpen"]
    assert list(h_coeff_5tap_map) == ["moving_avg", "simple_lp", "edge", "sharpen"]
    q = lambda h: quant.quantize_taps(h, 12, 16).tolist()
    assert q(h_coeff_3tap_map["edge"]) == [-4096, 0, 4096]
    assert q(h_coeff_3tap_map["sharpen"]) == [-512, 5120, -512]
    assert q(h_coeff_5tap_map["moving_avg"]) == [819] * 5
    assert q(h_coeff_5tap_map["simple_lp"]) == [256, 1024, 1536, 1024, 256]
    assert q(h_coeff_5tap_map["edge"]) == [-512, -1024, 0, 1024, 512]
    assert q(h_coeff_5tap_map["sharpen"]) == [-256, -1024, 6656, -1024, -256]


def test_host_container_normalisation():
    c = ops._np_container
    assert c([1, 2, 300, -4]).tolist() == [1, 2, 255, 0] and c([1, 2]).dtype == np.uint8
    assert c([1.5, 2]).dtype == np.float64
    assert c(np.array([1, -1], np.int8)).dtype == np.int16
    assert c(np.array([70000, 3], np.uint32)).tolist() == [255, 3]
    assert c(np.array([2 ** 40, -2 ** 40], np.int64)).tolist() == [255, 0]
    assert c([10 ** 30, -10 ** 30, 7]).tolist() == [1e30, -1e30, 7.0]          # beyond int64 -> float64
    assert c(np.array([1.5], np.float16)).dtype == np.float32
    assert c(np.array([True, False])).tolist() == [1, 0]
    a = np.arange(12, dtype=np.int16).reshape(3, 4)[:, ::2]
    assert c(a).flags["C_CONTIGUOUS"]
    with pytest.raises(TypeError):
        c(np.array(["a"]))


def test_empty_input_is_served_without_a_device():
    from fir_1d.model.python.fir_1d_fixed_ref import fir_1d_fixed_golden
    from fir_1d.model.python.fir_1d_ref import fir_1d_ideal
    y = fir_1d_fixed_golden([], [0.25, 0.5, 0.25])
    assert isinstance(y, np.ndarray) and y.dtype == np.uint8 and y.shape == (0,)
    assert fir_1d_ideal([], [1.0]) == []


def test_dropin_modules_keep_reference_names():
    import fir_1d.model.python.fir_1d_fixed_ref as fx
    import fir_1d.model.python.fir_1d_ref as idl
    import fir_1d.sim.vector.gen_fixed_output as gf
    import fir_1d.sim.vector.gen_ideal_output as gi
    import pipeline_fir_1d as pipe
    import inspect
    assert fx.MAX_ABS_H_COEFF == idl.MAX_ABS_H_COEFF == 8.0
    sig = inspect.signature(fx.fir_1d_fixed_golden)
    assert list(sig.parameters) == ["x", "h", "frac_bits", "acc_bits", "coeff_bits"]
    assert [p.default for p in sig.parameters.values()][2:] == [12, 32, 16]
    assert list(inspect.signature(idl.fir_1d_ideal).parameters) == ["x", "h"]
    for name in ("_validate_h_coefficients", "_validate_x", "_round_half_up_x", "_clamp_x"):
        assert callable(getattr(idl, name)) and getattr(fx, name) is getattr(idl, name)
    assert idl._round_half_up_x([-1.2, 0.49, 0.5, 1.5, 0.49999999999999994]) == [-1, 0, 1, 2, 1]
    assert idl._clamp_x([-5, 0, 255, 256]) == [0, 0, 255, 255]
    with pytest.raises(ValueError, match=r"x\[1\].*finite"):
        idl._validate_x([1, float("inf")])
    for mod, names in ((gf, ("generate_fixed_3tap_output_vector", "generate_fixed_5tap_output_vector", "_run_fixed_rowwise")),
                       (gi, ("generate_ideal_3tap_output_vector", "generate_ideal_5tap_output_vector", "_run_ideal_rowwise"))):
        for n in names:
            assert callable(getattr(mod, n))
    p = inspect.signature(gf.generate_fixed_3tap_output_vector).parameters
    assert list(p) == ["input_dir", "output_dir", "frac_bits", "acc_bits", "coeff_bits", "overwrite"]
    assert p["frac_bits"].kind is inspect.Parameter.KEYWORD_ONLY
    want = ["tap", "overwrite_vectors", "skip_input", "skip_ideal", "skip_fixed", "skip_report", "skip_restore",
            "restore_kind", "ideal_policy", "overwrite_images", "strict_report", "strict_restore", "top_k"]
    assert list(inspect.signature(pipe.run_pipeline).parameters) == want
    ns = pipe._build_argparser().parse_args(["--tap", "5", "--skip-input", "--top-k", "3"])
    assert set(vars(ns)) == set(want) and ns.tap == "5" and ns.skip_input and ns.top_k == 3


def test_generators_file_contract_without_inputs(tmp_path):
    from fir_1d.sim.vector.gen_fixed_output import generate_fixed_3tap_output_vector
    with pytest.raises(FileNotFoundError, match="No input .npy files found"):
        generate_fixed_3tap_output_vector(input_dir=tmp_path, output_dir=tmp_path / "o")
    from fir_1d.sim.vector import _batched
    np.save(tmp_path / "case_000_a_x_u8.npy", np.zeros((2, 3), np.int32))
    np.save(tmp_path / "Case_001_b_x_u8.npy", np.zeros(4, np.uint8))
    np.save(tmp_path / "other.npy", np.zeros(4, np.uint8))
    files = _batched.list_input_files(tmp_path)
    assert [f.name for f in files] == ["case_000_a_x_u8.npy", "Case_001_b_x_u8.npy"]
    assert _batched.load_image_u8(files[0]).dtype == np.uint8
    with pytest.raises(ValueError, match="expected 2D array"):
        _batched.load_image_u8(files[1])
    assert _batched.case_stem(files[0]) == "case_000_a"


# ---- shard planner ------------------------------------------------------------------
def test_taps_reach_follows_centre_alignment():
    assert shard.taps_reach(5) == (2, 2)
    assert shard.taps_reach(3) == (1, 1)
    assert shard.taps_reach(4) == (1, 2)          # even L: centre = L//2 reaches further ahead
    assert shard.taps_reach(2) == (0, 1)
    assert shard.taps_reach(1) == (0, 0)
    assert shard.taps_reach(5, 2) == (4, 4)


@pytest.mark.parametrize("total, parts, align", [(1 << 28, 8, 4096), (1000, 3, 1), (10, 4, 4), (0, 2, 8), (5, 8, 1), (4097, 2, 4096)])
def test_split_even_covers_exactly(total, parts, align):
    ranges = shard.split_even(total, parts, align)
    assert len(ranges) == parts and ranges[0][0] == 0 and ranges[-1][1] == total
    for (a, b), (c, d) in zip(ranges, ranges[1:]):
        assert b == c and a <= b
    for a, b in ranges[:-1]:
        assert b % align == 0 or b == total


def test_segment_plan_halos():
    pieces = shard.plan_segments(1 << 20, 4, 5, 1)
    assert [p.before for p in pieces] == [0, 2, 2, 2] and [p.after for p in pieces] == [2, 2, 2, 0]
    assert all(p.size == 1 << 18 for p in pieces)
    iq = shard.plan_segments(1 << 20, 2, 5, 2)
    assert iq[1].before == 4 and iq[0].after == 4 and iq[1].start % 2 == 0
    with pytest.raises(ValueError):
        shard.plan_segments(100, 2, 5, 2, align=3)
    bands = shard.plan_bands(16384, 8, 5)
    assert bands[0].before == 0 and bands[0].after == 2 and bands[7].after == 0 and bands[3].lo == 3 * 2048 - 2
    assert shard.aligned_lead(2, 2) == 128 and shard.aligned_lead(0, 2) == 0 and shard.aligned_lead(4, 1) == 256      # 256-byte boundaries


def test_sharded_result_equals_whole_signal_on_cpu():
    """Planner + halo semantics, with the oracle standing in for the kernel."""
    rng = np.random.default_rng(8)
    x = rng.integers(-100, 400, size=10_000).astype(np.int16)
    h = [1 / 16, 4 / 16, 6 / 16, 4 / 16, 1 / 16]
    for dil in (1, 2):
        want = onp.fir1d_fixed(x, h, 15, 32, 16, dilation=dil)
        for parts in (2, 3, 7):
            got = np.empty_like(want)
            for p in shard.plan_segments(x.size, parts, 5, dil, align=64):
                local = onp.fir1d_fixed(x[p.lo:p.hi], h, 15, 32, 16, dilation=dil)
                got[p.start:p.stop] = local[p.before:p.before + p.size]
            assert np.array_equal(got, want)


def test_host_out_buffers_are_validated_before_the_library_writes():
    x = np.arange(8, dtype=np.uint8)
    for bad in (np.zeros(7, np.uint8), np.zeros(8, np.int16), np.zeros(16, np.uint8)[::2], np.zeros((2, 4), np.uint8)):
        with pytest.raises(ValueError, match="out must be a writable C-contiguous uint8 array of shape"):
            ops.fir1d_fixed(x, [1.0], out=bad)
    with pytest.raises(ValueError, match="float64 array of shape"):
        ops.fir1d_ideal(x, [1.0], out=np.zeros(8, np.float32))
    frozen = np.zeros(8, np.uint8)
    frozen.setflags(write=False)
    with pytest.raises(ValueError, match="writable"):
        ops.fir1d_fixed(x, [1.0], out=frozen)


def test_image_view_addressing():
    """ops._cuda_imgs: which [..., H, W] views can be addressed in place (pitch /
    image stride in elements) and which must be refused — pure stride logic,
    checked here on CPU tensors."""
    import torch
    from firgpu import ops
    base = torch.zeros(3, 40, 100, dtype=torch.int16)
    assert ops._cuda_imgs(base, "x") == (3, 40, 100, 100, 4000)
    assert ops._cuda_imgs(base[:, :, 5:71], "x") == (3, 40, 66, 100, 4000)          # column view
    assert ops._cuda_imgs(base[:, 7:19, :], "x") == (3, 12, 100, 100, 4000)         # row band of every image
    assert ops._cuda_imgs(base[1, 3:9, 2:50], "x") == (1, 6, 48, 100, 600)          # one image, window
    assert ops._cuda_imgs(base[:1, :, 1:2], "x")[:3] == (1, 40, 1)
    four = torch.zeros(2, 3, 8, 16)
    assert ops._cuda_imgs(four, "x") == (6, 8, 16, 16, 128)                         # contiguous leading axes fold into the batch
    for bad in (base[:, :, ::2], base.transpose(1, 2), four[:, :, :, 2:9]):
        with pytest.raises(ValueError):
            ops._cuda_imgs(bad, "x")


# ---- round 2: host-side decisions that need no device -------------------------------------------------------------
@pytest.mark.parametrize("h, frac, acc, cb, want", [
    ([1 / 16, 4 / 16, 6 / 16, 4 / 16, 1 / 16], 15, 32, 16, ("staged", "hi16", "dp2a")),      # the headline taps: pre-scaled by 2, one byte permute
    ([-0.125, 1.25, -0.125], 12, 32, 16, ("staged", "shift", "dp2a")),                        # sharpen: 5120 * 16 does not fit int16
    ([0.9, -0.7, 0.8, 0.3, -0.2], 12, 16, 16, ("staged", "wrap", "dp2a")),                    # 255 * sum|hq| >= 2^15: the wrap is live
    ([1.0], 12, 20, 16, ("staged", "wrap", "dp2a")),
    ([0.9, -0.7, 0.8, 0.3, -0.2], 12, 24, 16, ("staged", "shift", "dp2a")),                   # same taps, wrap provably idle
    ([7.5, -7.5, 0.001], 28, 32, 32, ("staged", "wrap", "imad")),                             # 32-bit taps
    ([0.001, 0.002], 28, 40, 32, ("staged", "wrap", "imad")),                                 # wider accumulator, sums fit int32
    ([0.4, 0.3], 30, 64, 32, ("generic", None, None)),                                        # sums beyond int32 in a 64-bit accumulator
    ([0.01] * 21, 12, 18, 16, ("staged", "wrap", "dp2a")),                                    # long-filter bucket under a live wrap
    ([0.01] * 70, 12, 32, 16, ("staged", "hi16", "dp2a")),                                    # 97-tap bucket
    ([0.005] * 129, 12, 32, 16, ("staged", "hi16", "dp2a")),                                  # the largest bucket
    ([0.005] * 130, 12, 32, 16, ("generic", None, None)),                                     # beyond the largest bucket
])
def test_fixed_point_datapath_is_chosen_from_the_taps_alone(h, frac, acc, cb, want):
    import firgpu
    got = firgpu.fir1d_fixed_plan(h, frac, acc, cb)
    assert (got["path"], got["post"], got["mac"]) == want


def test_vector_root_is_one_tree_for_all_pipeline_stages(monkeypatch, tmp_path):
    """The GPU-backed generators default to the SAME input / output tree as the reference's untouched stages
    (advisor finding, round 1): FIR_VECTOR_ROOT, else the reference's fir_1d/sim/vector, else next to the drop-in."""
    from pathlib import Path
    from fir_1d.sim.vector import _batched
    monkeypatch.delenv("FIR_VECTOR_ROOT", raising=False)
    monkeypatch.delenv("FIR_REFERENCE_ROOT", raising=False)
    assert _batched.vector_root() == Path(_batched.__file__).resolve().parent
    ref = tmp_path / "ref"
    (ref / "fir_1d" / "sim" / "vector").mkdir(parents=True)
    monkeypatch.setenv("FIR_REFERENCE_ROOT", str(ref))
    assert _batched.vector_root() == (ref / "fir_1d" / "sim" / "vector").resolve()
    monkeypatch.setenv("FIR_VECTOR_ROOT", str(tmp_path / "elsewhere"))
    assert _batched.vector_root() == (tmp_path / "elsewhere").resolve()


def test_stream_and_record_entry_points_reject_bad_arguments_without_a_device():
    import ctypes as C
    from firgpu import _lib
    L = _lib.lib()
    h = C.c_void_p()
    taps = (C.c_int32 * 3)(1, 2, 1)
    assert L.firgpu_stream_create_fixed(C.byref(h), 9, _lib.U8, taps, 3, 1, 12, 32, 0, 255, 0, 255, 0) == -1       # dtype
    assert L.firgpu_stream_create_fixed(C.byref(h), _lib.U8, _lib.U8, taps, 0, 1, 12, 32, 0, 255, 0, 255, 0) == -3  # taps
    assert L.firgpu_stream_create_fixed(C.byref(h), _lib.U8, _lib.F64, taps, 3, 1, 12, 32, 0, 255, 0, 255, 0) == -1
    assert L.firgpu_stream_create_fixed(C.byref(h), _lib.U8, _lib.U8, taps, 3, 0, 12, 32, 0, 255, 0, 255, 0) == -5  # dilation
    assert L.firgpu_stream_create_fixed(C.byref(h), _lib.U8, _lib.U8, taps, 3, 1, 0, 32, 0, 255, 0, 255, 0) == -4   # bits
    assert L.firgpu_stream_process(None, None, None, 5, None) == -6
    assert L.firgpu_record_begin(None) == -6                      # the NULL stream cannot record
    assert L.firgpu_graph_launch(None, None) == -6
    assert b"NULL" in L.firgpu_last_error()
