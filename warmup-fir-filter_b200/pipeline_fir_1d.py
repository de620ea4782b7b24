"""Drop-in for the reference's ``pipeline_fir_1d.py`` driver: same
``run_pipeline`` keyword signature, result dict and CLI flags
(pipeline_fir_1d.py:34-98,102-175).

The two model stages (ideal / fixed output generation) run on the GPU through
``fir_1d.sim.vector.gen_{ideal,fixed}_output``.  Input-vector generation,
compare reports and PNG restore are the reference's own file-plumbing modules;
they are imported under their usual names when ``FIR_REFERENCE_ROOT`` exposes
them (see ``fir_1d/sim/vector/__init__.py``) and must otherwise be skipped
with the matching ``--skip-*`` flag.
"""
from __future__ import annotations

import argparse
import importlib
from pathlib import Path
from time import perf_counter
from typing import Any

from fir_1d.sim.vector import _batched, gen_fixed_output, gen_ideal_output

STAGE_FLAGS = ("skip_input", "skip_ideal", "skip_fixed", "skip_report", "skip_restore")


def _selected_taps(tap: str) -> list[str]:
    return ["3", "5"] if tap == "all" else [tap]


def _reference_stage(module: str, attr: str):
    try:
        return getattr(importlib.import_module(f"fir_1d.sim.vector.{module}"), attr)
    except ImportError as exc:
        raise ImportError(
            f"stage module fir_1d.sim.vector.{module} is the reference's own file plumbing; set "
            f"FIR_REFERENCE_ROOT to a checkout of the reference or skip the stage ({exc})") from exc


def _log_stage(message: str) -> None:
    print(f"[pipeline] {message}")


def run_pipeline(*, tap: str, overwrite_vectors: bool, skip_input: bool, skip_ideal: bool, skip_fixed: bool,
                 skip_report: bool, skip_restore: bool, restore_kind: str, ideal_policy: str, overwrite_images: bool,
                 strict_report: bool, strict_restore: bool, top_k: int) -> dict[str, Any]:
    taps = _selected_taps(tap)
    results: dict[str, Any] = {"selected_taps": taps}
    # ONE vector tree for all five stages, passed explicitly: the reference's own stages default to directories next
    # to their source files, the GPU-backed ones must not default to somewhere else (see _batched.vector_root).
    root = _batched.vector_root()
    input_dir, output_dir = root / "input", root / "output"
    if not skip_input:
        input_dir.mkdir(parents=True, exist_ok=True)       # (the reference's stage creates the leaf only)

    if not skip_input:
        _log_stage("Generate input vectors")
        make_inputs = _reference_stage("gen_input_vectors", "generate_input_vector_jsons")
        results["input_manifest"] = make_inputs(output_dir=input_dir, overwrite=overwrite_vectors)

    if not skip_ideal:
        _log_stage("Generate ideal outputs (GPU)")
        results["ideal_counts"] = {
            f"ideal_{t}tap": getattr(gen_ideal_output, f"generate_ideal_{t}tap_output_vector")(
                input_dir=input_dir, output_dir=output_dir, overwrite=overwrite_vectors)
            for t in taps}

    if not skip_fixed:
        _log_stage("Generate fixed outputs (GPU)")
        results["fixed_counts"] = {
            f"fixed_{t}tap": getattr(gen_fixed_output, f"generate_fixed_{t}tap_output_vector")(
                input_dir=input_dir, output_dir=output_dir, overwrite=overwrite_vectors)
            for t in taps}

    if not skip_report:
        _log_stage("Generate compare reports")
        results["report_results"] = {
            f"report_{t}tap": _reference_stage(f"gen_{t}tap_compare_report", f"generate_{t}tap_compare_report")(
                ideal_dir=output_dir / f"ideal_{t}tap", fixed_dir=output_dir / f"fixed_{t}tap",
                report_dir=output_dir / f"report_{t}tap", top_k=top_k, strict=strict_report)
            for t in taps}

    if not skip_restore:
        _log_stage("Restore output images")
        summary = _reference_stage("restore_images", "restore_images")(
            vector_output_dir=output_dir, output_img_dir=root.parent / "output_img",
            kind=restore_kind, tap=tap, ideal_policy=ideal_policy, overwrite=overwrite_images, strict=strict_restore)
        results["restore_summary"] = {k: summary[k] for k in ("num_converted", "num_skipped")}

    return results


def _build_argparser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="FIR 1D end-to-end pipeline with the model stages on the GPU.")
    p.add_argument("--tap", choices=("all", "3", "5"), default="all")
    p.add_argument("--overwrite-vectors", action="store_true")
    for stage in ("input", "ideal", "fixed", "report", "restore"):
        p.add_argument(f"--skip-{stage}", action="store_true", help=f"Skip the {stage} stage.")
    p.add_argument("--restore-kind", choices=("all", "ideal", "fixed"), default="all")
    p.add_argument("--ideal-policy", choices=("clip", "normalize"), default="clip")
    p.add_argument("--overwrite-images", action="store_true")
    p.add_argument("--strict-report", action="store_true")
    p.add_argument("--strict-restore", action="store_true")
    p.add_argument("--top-k", type=int, default=5)
    return p


def main(argv=None) -> None:
    args = _build_argparser().parse_args(argv)
    t0 = perf_counter()
    skipped = sum(int(getattr(args, f)) for f in STAGE_FLAGS)
    out = (_batched.vector_root() / "output").resolve()
    try:
        summary = run_pipeline(**{k: v for k, v in vars(args).items()})
    except Exception as exc:
        print(f'[FAIL] pipeline_fir_1d file=pipeline_fir_1d.py generated=0 skipped={skipped} failed=1 '
              f'elapsed={perf_counter() - t0:.2f}s out={out} error="{exc}"')
        raise
    done = len([k for k in summary if k != "selected_taps"])
    print(f"[OK] pipeline_fir_1d file=pipeline_fir_1d.py generated={done} skipped={skipped} failed=0 "
          f"elapsed={perf_counter() - t0:.2f}s out={out}")


if __name__ == "__main__":
    main()
