"""The reference's OWN pytest suite (fir_1d/sim/tests, 39 tests, unmodified, run
from where it lies under /root/reference) against the GPU-backed drop-in,
through the overlay described in INTEGRATION.md.  Only where the reference tree
exists (this container).  Without a GPU the 22 validation / error-contract
tests pass and the 17 computing tests fail with the library's CUDA error —
which also proves there is no CPU fallback; with a GPU all 39 must pass."""
from __future__ import annotations

import os
import re
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
REFERENCE = Path("/root/reference")
pytestmark = pytest.mark.skipif(not (REFERENCE / "fir_1d" / "sim" / "tests").is_dir(), reason="reference tree not present")


def test_reference_suite_runs_against_the_dropin(tmp_path):
    env = dict(os.environ, PYTHONPATH=str(ROOT / "warmup-fir-filter_b200"), FIR_REFERENCE_ROOT=str(REFERENCE),
               PYTHONDONTWRITEBYTECODE="1")
    out = subprocess.run([sys.executable, "-m", "pytest", "-p", "no:cacheprovider", f"--rootdir={tmp_path}",
                          str(REFERENCE / "fir_1d" / "sim" / "tests"), "-q"],
                         cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    text = out.stdout + out.stderr
    assert "warmup-fir-filter_b200" in subprocess.run(
        [sys.executable, "-c", "import fir_1d.model.python.fir_1d_fixed_ref as m; print(m.__file__)"],
        cwd=tmp_path, env=env, capture_output=True, text=True).stdout, "the overlay did not take precedence"
    passed = int((re.search(r"(\d+) passed", text) or [0, 0])[1])
    failed = int((re.search(r"(\d+) failed", text) or [0, 0])[1])
    assert passed + failed == 39, text[-2000:]
    import torch
    if torch.cuda.is_available():
        assert failed == 0, text[-3000:]
    else:
        assert passed == 22 and failed == 17, text[-2000:]
        assert text.count("FirGpuError") >= 17 and "CUDA error" in text      # loud failure, no silent CPU result
