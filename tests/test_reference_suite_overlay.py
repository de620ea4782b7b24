"""The reference's OWN pytest suite (fir_1d/sim/tests, 39 tests, unmodified) and
its own pipeline driver stages against the GPU-backed drop-in, through the
overlay described in INTEGRATION.md.

Where the reference comes from: ``/root/reference`` in the build container; on
the GPU box (where that path does not exist) the byte-for-byte copy that
``tools/stage_reference.py`` puts under the git-ignored ``baseline/_ref/``.
Without either, the tests skip.

* no GPU (build container): the 22 validation / error-contract tests pass and
  the 17 computing tests fail with the library's CUDA error — which also proves
  there is no CPU fallback;
* GPU: all 39 must pass, and ``pipeline_fir_1d --overwrite-vectors`` with NO
  stage skipped (image decode -> ideal -> fixed -> compare reports -> PNG
  restore, all five on one vector tree) reproduces the reference pipeline's
  output files by SHA-256 and its published report numbers.
"""
from __future__ import annotations

import hashlib
import json
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
STAGED = ROOT / "baseline" / "_ref" / "reference_tree"
REFERENCE = Path("/root/reference") if (Path("/root/reference") / "fir_1d" / "sim" / "tests").is_dir() else STAGED
needs_reference = pytest.mark.skipif(not (REFERENCE / "fir_1d" / "sim" / "tests").is_dir(),
                                     reason="reference tree neither at /root/reference nor staged under baseline/_ref")


def _has_gpu() -> bool:
    import torch
    return torch.cuda.is_available()


def _run_suite(tmp_path):
    env = dict(os.environ, PYTHONPATH=str(ROOT / "warmup-fir-filter_b200"), FIR_REFERENCE_ROOT=str(REFERENCE),
               PYTHONDONTWRITEBYTECODE="1")
    out = subprocess.run([sys.executable, "-m", "pytest", "-p", "no:cacheprovider", f"--rootdir={tmp_path}",
                          str(REFERENCE / "fir_1d" / "sim" / "tests"), "-q"],
                         cwd=tmp_path, env=env, capture_output=True, text=True, timeout=900)
    text = out.stdout + out.stderr
    assert "warmup-fir-filter_b200" in subprocess.run(
        [sys.executable, "-c", "import fir_1d.model.python.fir_1d_fixed_ref as m; print(m.__file__)"],
        cwd=tmp_path, env=env, capture_output=True, text=True).stdout, "the overlay did not take precedence"
    passed = int((re.search(r"(\d+) passed", text) or [0, 0])[1])
    failed = int((re.search(r"(\d+) failed", text) or [0, 0])[1])
    return passed, failed, text


@needs_reference
def test_reference_suite_runs_against_the_dropin(tmp_path):
    passed, failed, text = _run_suite(tmp_path)
    assert passed + failed == 39, text[-2000:]
    if _has_gpu():
        assert failed == 0, text[-3000:]
    else:
        assert passed == 22 and failed == 17, text[-2000:]
        assert text.count("FirGpuError") >= 17 and "CUDA error" in text      # loud failure, no silent CPU result


@pytest.mark.gpu
@needs_reference
def test_reference_suite_39_passed_on_the_gpu(tmp_path):
    """the reference's tests, unmodified, against the GPU-backed drop-in.  The log is kept under
    gpurun_out/ when that directory exists (copied to profiles/ by hand after a run)."""
    passed, failed, text = _run_suite(tmp_path)
    log = ROOT / "gpurun_out"
    if log.is_dir():
        (log / "reference_suite_on_gpu.log").write_text(text)
    assert (passed, failed) == (39, 0), text[-3000:]


@pytest.mark.gpu
@needs_reference
def test_pipeline_overlay_end_to_end_no_stage_skipped(tmp_path, manifest, compare_rows):
    """`python -m pipeline_fir_1d --overwrite-vectors --overwrite-images` exactly as INTEGRATION.md A documents it:
    the reference's untouched image-decode, report and restore stages around the two GPU stages, all on ONE vector
    tree (FIR_VECTOR_ROOT: the reference checkout may be read-only).  Outputs = the reference pipeline's own files."""
    root = tmp_path / "vector"
    env = dict(os.environ, PYTHONPATH=str(ROOT / "warmup-fir-filter_b200"), FIR_REFERENCE_ROOT=str(REFERENCE),
               FIR_VECTOR_ROOT=str(root), PYTHONDONTWRITEBYTECODE="1")
    out = subprocess.run([sys.executable, "-m", "pipeline_fir_1d", "--overwrite-vectors", "--overwrite-images"],
                         cwd=tmp_path, env=env, capture_output=True, text=True, timeout=1800)
    text = out.stdout + out.stderr
    assert out.returncode == 0 and "[OK] pipeline_fir_1d" in text and "skipped=0" in text, text[-3000:]
    inputs = sorted((root / "input").glob("*_x_u8.npy"))
    assert len(inputs) == 7
    n = 0
    for kind, suffix in (("fixed", "y_u8"), ("ideal", "y_f64")):
        for tap in ("3tap", "5tap"):
            files = sorted((root / "output" / f"{kind}_{tap}").glob("*.npy"))
            assert len(files) == 28, (kind, tap, len(files))
            for f in files:
                key = f.name.replace(f"_{kind}_{tap}_{suffix}.npy", f"_{tap}")
                if key in manifest[kind]:              # (JPEG-decoded inputs depend on libjpeg; the manifest has what the reference made here)
                    assert hashlib.sha256(np.load(f).tobytes()).hexdigest() == manifest[kind][key], f.name
                    n += 1
    assert n >= 96
    for tap in ("3", "5"):
        rep = json.loads((root / "output" / f"report_{tap}tap" / f"compare_{tap}tap_summary.json").read_text())
        want = compare_rows[f"{tap}tap"]["summary"]                     # the reference's report code on the reference's own outputs
        assert rep["overall"]["num_cases"] == want["num_cases"] == 28
        for key in ("avg_mae", "avg_rmse", "avg_max_abs_err", "max_rmse"):
            assert rep["overall"][key] == pytest.approx(want[key], rel=1e-12), (tap, key)
    assert any((root.parent / "output_img").rglob("*.png")), "restore stage wrote no images"
