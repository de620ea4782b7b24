"""Shared test plumbing.

* ``-m "not gpu"`` : oracle vs. the reference-made golden vectors, host logic,
  C-ABI symbol checks, gloo world_size-2 sharding — no CUDA device needed.
* ``-m gpu``       : parity of the CUDA path (through the C ABI) against the
  oracle and the golden vectors.  These fail (not skip) when the extension is
  missing on a box that has a GPU.
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
PKG = ROOT / "warmup-fir-filter_b200"
GOLDEN = ROOT / "tests" / "golden"
for p in (str(ROOT), str(PKG)):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # The tests need the in-tree library; a fresh checkout builds it once (nvcc cross-compiles without a GPU).
    lib = PKG / "firgpu" / "libfirgpu.so"
    if not lib.exists():
        import subprocess
        subprocess.run(["make", "-s", "-j8", "-C", str(PKG / "csrc")], check=True)


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class GoldenCases:
    """tests/golden/fir1d_cases.npz: inputs, parameters and the outputs of the
    UNMODIFIED reference functions (made by tests/golden/make_golden.py)."""

    def __init__(self):
        self.blob = np.load(GOLDEN / "fir1d_cases.npz")
        self.index = json.loads(bytes(self.blob["index_json"]).decode())

    def __iter__(self):
        for meta in self.index:
            n = meta["name"]
            yield dict(name=n, x=self.blob[f"{n}/x"], h=self.blob[f"{n}/h"].tolist(),
                       y_fixed=self.blob[f"{n}/y_fixed"], y_ideal=self.blob[f"{n}/y_ideal"],
                       frac_bits=meta["frac_bits"], acc_bits=meta["acc_bits"], coeff_bits=meta["coeff_bits"])


@pytest.fixture(scope="session")
def golden_cases():
    return list(GoldenCases())


@pytest.fixture(scope="session")
def manifest():
    return json.loads((GOLDEN / "manifest.json").read_text())


@pytest.fixture(scope="session")
def images_small():
    return np.load(GOLDEN / "images_small.npz")


@pytest.fixture(scope="session")
def images_inputs():
    return np.load(GOLDEN / "images_inputs.npz")


@pytest.fixture(scope="session")
def image_crops():
    return np.load(GOLDEN / "image_crops.npz")


@pytest.fixture(scope="session")
def image_large():
    """(stem, uint8 [2999, 4499]) — the reference's largest test image."""
    blob = np.load(GOLDEN / "image_large_input.npz")
    stem = list(blob.keys())[0]
    return stem, blob[stem]


@pytest.fixture(scope="session")
def compare_rows():
    """Per-case metrics and summaries computed by the reference's own report
    code on its own outputs (tests/golden/make_golden_compare.py)."""
    return json.loads((GOLDEN / "compare_rows.json").read_text())
