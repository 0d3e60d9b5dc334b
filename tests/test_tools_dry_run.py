"""Dry run of the GPU-side tool scripts bench.py launches as diagnostic legs (tools/bench_shape_sens.py, tools/bench_assembly_variants.py) and of
the tuning sweep (tools/sweep_assembly.py): their control flow, argument handling and JSON output, with the device context replaced by the CPU
test double (tests/host_standin.py) and tiny meshes.  Timings and variant switches mean nothing here; the point is that the scripts cannot fail
on host-side logic the first time they meet a GPU."""
import io
import json
import os
import runpy
import sys
from contextlib import redirect_stdout

import pytest

import wae_b200 as W
from host_standin import HostStandIn
from wae_b200 import helmholtz, nlevp, shape

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KNOBS = ("WAE_GATHER_SLOTS", "WAE_GATHER_CTAS", "WAE_GATHER_THREADS", "WAE_ASM_VARIANT", "WAE_LU_NBO", "WAE_LU_LEAF", "WAE_LU_SKIP_UPPER", "WAE_LU_GEMM", "WAE_LU_SOLVE_PF", "WAE_EIGS_PAIRED")


def _run(monkeypatch, tool, argv):
    ctx = HostStandIn()
    for m in (W, helmholtz, shape, nlevp):
        monkeypatch.setattr(m, "get_context", lambda device=None: ctx)
    monkeypatch.setattr(sys, "argv", [tool] + argv)
    for k in KNOBS:
        monkeypatch.delenv(k, raising=False)
    buf = io.StringIO()
    with redirect_stdout(buf):
        runpy.run_path(os.path.join(ROOT, "tools", tool), run_name="__main__")
    for k in KNOBS:
        os.environ.pop(k, None)  # the tools set them for the library; leave the process environment clean for the other tests
    return json.loads([l for l in buf.getvalue().splitlines() if l.startswith("{\"")][-1])


def test_shape_sensitivity_bench_tool(monkeypatch):
    out = _run(monkeypatch, "bench_shape_sens.py", ["3", "3", "24", "3"])
    assert out["surface_points"] > 100 and out["finite"] and out["householder"]["flag"] >= 0
    assert out["oracle"]["ok"] and out["oracle"]["points"] == 3


@pytest.mark.parametrize("tool", ["bench_assembly_variants.py", "sweep_assembly.py"])
def test_assembly_tools(monkeypatch, tool):
    out = _run(monkeypatch, tool, ["3", "quad", "2"])
    assert out["tets"] == 162 and out["order"] == "quad"
    if tool == "sweep_assembly.py":
        assert out["n_failed"] == 0 and len(out["best"]) == 8 and all(r["ok"] for r in out["best"])
    else:
        assert set(out["variants"]) == {"0", "1", "2", "3", "0_again"} and all(v["ok"] for v in out["variants"].values())
        assert len(out["layouts"]) == 6 and all(r["ok"] for r in out["layouts"])


def test_lu_knob_tool(monkeypatch):
    out = _run(monkeypatch, "bench_lu_knobs.py", ["3", "3", "24", "quad", "1"])
    assert out["householder"]["rel_diff"] < 1e-9 and out["householder"]["paired"]["flag"] >= 0
    assert len(out["combos"]) == 10 and all(r.get("ok") for r in out["combos"]) and "free_error" not in out, out
    assert [(r["nbo"], r["leaf"], r["skip_upper"], r["gemm"], r["solve_pf"]) for r in out["combos"]] == [
        (128, 64, 0, 1, 0), (64, 64, 0, 1, 0), (256, 64, 0, 1, 0), (128, 32, 0, 1, 0), (128, 128, 0, 1, 0), (128, 64, 1, 1, 0), (256, 64, 1, 1, 0),
        (128, 64, 0, 2, 0), (128, 64, 1, 2, 0), (128, 64, 0, 1, 1)]
