#!/usr/bin/env python
"""Stage the UNMODIFIED reference files that the overlay tests need on the GPU box.

/root/reference exists only in the build container.  The reference's own test-suite (fir_1d/sim/tests, 39 tests) is the
parity contract for the drop-in, and its file-plumbing stages (image decode, compare reports, PNG
restore) are what pipeline_fir_1d.py runs around the two GPU stages — so a byte-for-byte copy of those files travels
to the box inside baseline/_ref/ (git-ignored: never part of this repository's history, never imported by the
product; only tests/test_reference_suite_overlay.py points FIR_REFERENCE_ROOT at it).

    python tools/stage_reference.py            # no-op when /root/reference is absent
"""
from __future__ import annotations

import shutil
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
SRC = Path("/root/reference")
DST = ROOT / "baseline" / "_ref" / "reference_tree"
KEEP = ("fir_1d/__init__.py", "fir_1d/sim/__init__.py", "fir_1d/sim/tests", "fir_1d/sim/img", "fir_1d/sim/vector", "fir_1d/model",
        "pipeline_fir_1d.py")


def stage() -> bool:
    if not SRC.is_dir():
        return False
    if DST.exists():
        shutil.rmtree(DST)
    for rel in KEEP:
        s, d = SRC / rel, DST / rel
        if s.is_dir():
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "input", "output", "output_img"))
        elif s.is_file():
            d.parent.mkdir(parents=True, exist_ok=True)
            shutil.copy2(s, d)
    (DST / "STAGED_FROM").write_text(f"{SRC} (unmodified copy, made by tools/stage_reference.py)\n")
    return True


if __name__ == "__main__":
    print("staged" if stage() else "reference tree absent: nothing staged", DST)
    sys.exit(0)
