"""Oracle vs. the reference itself — runs only where /root/reference exists
(the build container).  The GPU box never sees the reference; there the
committed golden vectors (test_oracle_golden.py) carry the pin."""
from __future__ import annotations

import importlib
import sys
from pathlib import Path

import numpy as np
import pytest

from oracle import c_oracle, fir_oracle_np as onp, fir_oracle_py as opy

REFERENCE = Path("/root/reference")
pytestmark = pytest.mark.skipif(not (REFERENCE / "fir_1d").is_dir(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    """Import the reference's model modules without letting its ``fir_1d``
    package shadow (or be shadowed by) the drop-in package of the same name."""
    saved = {k: v for k, v in sys.modules.items() if k == "fir_1d" or k.startswith("fir_1d.")}
    for k in saved:
        del sys.modules[k]
    sys.dont_write_bytecode = True
    sys.path.insert(0, str(REFERENCE))
    try:
        fixed = importlib.import_module("fir_1d.model.python.fir_1d_fixed_ref")
        ideal = importlib.import_module("fir_1d.model.python.fir_1d_ref")
        assert str(REFERENCE) in fixed.__file__
        yield fixed.fir_1d_fixed_golden, ideal.fir_1d_ideal
    finally:
        sys.path.remove(str(REFERENCE))
        for k in [k for k in sys.modules if k == "fir_1d" or k.startswith("fir_1d.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_fuzz_fixed_and_ideal(ref):
    golden, ideal = ref
    rng = np.random.default_rng(77)
    for _ in range(300):
        n_taps = int(rng.integers(1, 10))
        cb = int(rng.choice([8, 16, 32]))
        fb = int(rng.integers(1, min(cb + 6, 40)))
        ab = int(rng.choice([8, 16, 24, 32, 33, 48, 64]))
        n = int(rng.integers(0, 70))
        hi = min(8.0, ((1 << (cb - 1)) - 1) / (1 << fb))
        lo = max(-8.0, -(1 << (cb - 1)) / (1 << fb))
        h = rng.uniform(lo, hi, size=n_taps).tolist()
        x = rng.uniform(-40, 300, size=n)
        want = golden(x.tolist(), h, frac_bits=fb, acc_bits=ab, coeff_bits=cb)
        assert np.array_equal(opy.fir1d_fixed(x.tolist(), h, fb, ab, cb), want)
        assert np.array_equal(onp.fir1d_fixed(x, h, fb, ab, cb), want)
        want_i = np.array(ideal(x.tolist(), h), dtype=np.float64)
        assert np.array_equal(onp.fir1d_ideal(x, h), want_i)
        if n:
            assert np.array_equal(c_oracle.fir1d_ideal(x, h), want_i)
            assert np.array_equal(c_oracle.fir1d_fixed(x, onp.quantize_taps(h, fb, cb), fb, ab), want)


def test_reference_cpp_full_convolution_slice():
    """Fir1D::fir_1d_idel (fir_1d/model/cpp/fir_1d_ref.cpp:37-63) is the full
    convolution; the Python model's same-mode output is its slice
    [L//2, L//2 + N) — same products, same ascending-k order."""
    if c_oracle.ref_lib() is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(3)
    for n_taps in (1, 2, 3, 4, 5, 8):
        x = rng.integers(0, 256, size=41).astype(np.float64)
        h = rng.uniform(-1, 1, size=n_taps)
        full = c_oracle.ref_fir1d_full(x, h)
        assert full.size == x.size + n_taps - 1
        same = onp.fir1d_ideal(x, h.tolist())
        assert np.allclose(full[n_taps // 2:n_taps // 2 + x.size], same, rtol=0, atol=1e-9)
        stream = c_oracle.ref_fir1d_stream(x, h)
        assert np.allclose(stream, full[:x.size], rtol=0, atol=1e-9)


def test_reference_complex_times_real_tap_is_componentwise():
    """lib/mycomplex.cpp:12-22: sum_k z[k] * h[k] with real h acts on Re and
    Im separately — the semantics of the dilation-2 (I/Q) kernel mode."""
    if c_oracle.ref_lib() is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(4)
    z = rng.integers(0, 256, size=10).astype(np.float64)
    h = rng.uniform(-1, 1, size=5)
    re, im = c_oracle.ref_complex_real_tap_mac(z, h)
    assert re == pytest.approx(float(np.dot(z[0::2], h)), abs=1e-9)
    assert im == pytest.approx(float(np.dot(z[1::2], h)), abs=1e-9)


def test_half_up_rounding_follows_the_container_precision(ref):
    """`math.floor(sample + 0.5)` (fir_1d_ref.py:37) runs on the container's own scalars: float64 for lists and float64
    arrays, float32 / float16 for such arrays (NumPy 2 promotion).  The three oracles agree with the reference on every
    x.5 tie, its neighbours and the one value per format where the precisions differ (the largest below 0.5)."""
    fixed, ideal = ref
    for dt in (np.float64, np.float32, np.float16):
        ints = np.arange(-2, 258).astype(dt)
        half = ints + dt(0.5)
        x = np.concatenate([ints, half, np.nextafter(half, dt(-np.inf)), np.nextafter(half, dt(np.inf)),
                            np.array([0.49999999999999994, 0.49999997, 0.49975586, 0.5, 255.49999], dtype=np.float64).astype(dt)])
        x = x[np.isfinite(x)]
        want = fixed(x, [1.0])
        assert np.array_equal(onp.fir1d_fixed(x, [1.0]), want), dt
        assert list(opy.fir1d_fixed(list(x), [1.0])) == want.tolist(), dt
        if dt != np.float16:                          # (the C oracle reads float32 / float64 containers)
            assert np.array_equal(c_oracle.fir1d_fixed(x, onp.quantize_taps([1.0], 12, 16), 12, 32), want), dt
        assert np.array_equal(np.asarray(ideal(x, [1.0])), onp.fir1d_ideal(x, [1.0])), dt
    assert fixed(np.array([0.49999997], np.float32), [1.0])[0] == 1 and fixed([0.49999997017860413], [1.0])[0] == 0
