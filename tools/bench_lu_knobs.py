"""Short sweep of the numeric-LU tuning knobs (WAE_LU_NBO: outer block width of the pivot-block factorisation, read at every factorisation;
WAE_LU_LEAF: nested-dissection leaf size, read by the symbolic analysis; WAE_LU_SKIP_UPPER: opt-in skip of the pivot-square tiles above the
diagonal in the trailing updates; WAE_LU_GEMM=2: opt-in cp.async ring in the DMMA GEMM; WAE_LU_SOLVE_PF=1: opt-in L2 prefetch
hint in the triangular window kernels of the solves) on the config-2 tube: per combination one analysis, `reps`
factorisations (CUDA-event time of the numeric phase), one refined solve and its residual.  Prints ONE JSON line.  bench.py runs this in
a subprocess as a diagnostic leg; the defaults (NBO 128, LEAF 64) are the measured configuration and are not changed by anything here.

    python tools/bench_lu_knobs.py [nx=20] [ny=20] [nz=300] [order=quad] [reps=2]
"""
import json
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import wae_b200 as W  # noqa: E402

nx, ny, nz = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (20, 20, 300)))
order = sys.argv[4] if len(sys.argv) > 4 else "quad"
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
KNOBS = ("WAE_LU_NBO", "WAE_LU_LEAF", "WAE_LU_SKIP_UPPER", "WAE_LU_GEMM", "WAE_LU_SOLVE_PF")
for k in KNOBS:
    os.environ.pop(k, None)
hz = 0.5 / nz
mesh = W.kuhn_box((nx, ny, nz), (0, 0, -0.25), (0.05, 0.05, 0.25), jitter=0.1, seed=12345, flame_layer=(nz // 2, nz // 2 + 1))
c = np.where(mesh.points[2, mesh.tetrahedra].sum(axis=1) / 4 < 0, 347.2, 694.4)
gam, rho = 1.4, 1.225
q = 101325.0 * 3 * math.pi * 0.025**2 * gam / (gam - 1)
dscrp = {"Interior": ("interior", ()), "Outlet": ("admittance", ("Y", 1e15)),
         "Flame": ("flame", (gam, rho, q, [0.025, 0.025, -0.6 * hz], [0, 0, 1.0], "n", "τ", 1.0, 0.001))}
L = W.discretize(mesh, dscrp, c, order=order)
dev = L.device()
ctx = dev.ctx
z = 340 * 2 * math.pi
op = L(z)
op.materialize(0)
rng = np.random.default_rng(0)
b = rng.standard_normal(L.size()) + 1j * rng.standard_normal(L.size())
out = {"tube": [nx, ny, nz], "order": order, "dofs": int(L.size()), "reps": reps, "combos": []}
analyses = {}
for nbo, leaf, skip, gemm, spf in ((None, None, None, None, None), (64, None, None, None, None), (256, None, None, None, None), (None, 32, None, None, None),
                                   (None, 128, None, None, None), (None, None, 1, None, None), (256, None, 1, None, None), (None, None, None, 2, None),
                                   (None, None, 1, 2, None), (None, None, None, None, 1)):
    for k, v in zip(KNOBS, (nbo, leaf, skip, gemm, spf)):
        os.environ.pop(k, None)
        if v is not None:
            os.environ[k] = str(v)
    row = {"nbo": nbo or 128, "leaf": leaf or 64, "skip_upper": skip or 0, "gemm": gemm or 1, "solve_pf": spf or 0}
    try:
        if leaf not in analyses:  # only the leaf size enters the symbolic analysis: the other knobs are read by every factorisation / solve
            analyses[leaf] = ctx.lu_analyze(dev.fid)
        lid, lu_nnz, lu_flops = analyses[leaf]
        ms = []
        for _ in range(reps):
            ctx.lu_factor(lid, 0)
            ms.append(ctx.last_ms("factor"))
        symf = 0.5 if ctx.last_ms("factor_sym") > 0.5 else 1.0
        sol = []
        for _ in range(3):
            x = ctx.lu_solve(lid, b)
            sol.append(ctx.last_ms("solve"))
        sol_ms = min(sol)
        res = float(np.abs(op.matvec(x) - b).max() / np.abs(b).max())
        if not out["combos"]:  # default configuration only: what a second / eighth right-hand side costs in the same pass over the factor
            for nr in (2, 8):   # (2 = the direct and the adjoint Arnoldi recurrence of householder advancing together)
                B = rng.standard_normal((L.size(), nr)) + 1j * rng.standard_normal((L.size(), nr))
                t = []
                for _ in range(2):
                    ctx.lu_solve(lid, B)
                    t.append(ctx.last_ms("solve"))
                row[f"solve{nr}_ms"] = float(min(t))
        row.update({"factor_ms": float(min(ms)), "factor_nnz": float(lu_nnz), "factor_flops": float(lu_flops) * symf, "tflops": symf * lu_flops / max(min(ms), 1e-9) / 1e9,
                    "solve_ms": float(sol_ms), "residual": res, "ok": bool(res <= 1e-6)})
    except Exception as e:  # noqa: BLE001 -- diagnostic only
        row["error"] = repr(e)[:200]
    out["combos"].append(row)
for k in KNOBS:
    os.environ.pop(k, None)
# one eigenpair with the default host loop and with WAE_EIGS_PAIRED=1 (direct and adjoint Arnoldi recurrence as the two right-hand sides of one
# pass over the factor, wae_eigs_si_pair): wall time, solves, and the difference of the two eigenvalues
if None in analyses:
    dev._lu, dev.lu_nnz, dev.lu_flops = analyses[None]  # the family's own handle: no further analysis
    hh = {}
    for tag, flag in (("default", None), ("paired", "1")):
        os.environ.pop("WAE_EIGS_PAIRED", None)
        if flag:
            os.environ["WAE_EIGS_PAIRED"] = flag
        try:
            st, t0 = {}, time.perf_counter()
            sol, nit, fl = W.householder(L, z, maxiter=15, tol=1e-9 * abs(z), output=False, stats=st)
            hh[tag] = {"wall_s": time.perf_counter() - t0, "iterations": nit, "flag": fl, "solves": st.get("solves"), "eigs_wall_s": st.get("eigs_wall_s"),
                       "omega": [sol.params["ω"].real, sol.params["ω"].imag]}
        except Exception as e:  # noqa: BLE001 -- diagnostic only
            hh[tag] = {"error": repr(e)[:200]}
    os.environ.pop("WAE_EIGS_PAIRED", None)
    if all("omega" in v for v in hh.values()):
        a, b2 = complex(*hh["default"]["omega"]), complex(*hh["paired"]["omega"])
        hh["rel_diff"] = abs(a - b2) / abs(a)
    out["householder"] = hh
for leaf, (lid, _, _) in analyses.items():
    try:
        ctx.lu_free(lid)  # 23 GB of factors per analysis at config 2
    except Exception as e:  # noqa: BLE001 -- the sweep fits the device without it
        out["free_error"] = repr(e)[:120]
print(json.dumps(out))
