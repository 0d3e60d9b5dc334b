"""First measurement of the shape-sensitivity path on a first-order Rijke tube: GPU discretize + householder, then
discrete_adjoint_shape_sensitivity for ALL surface points through the C ABI (one kernel launch per descriptor term), timed; the CPU
oracle's literal loop (six discretize calls per point, src/shape_sensitivity.jl:43-137) on a few points as the checker and as the
per-point CPU cost.  Prints ONE JSON line.  bench.py runs this in a subprocess (diagnostic leg "shape_sensitivity").

    python tools/bench_shape_sens.py [nx=10] [ny=10] [nz=150] [oracle_points=5]
"""
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import wae_b200 as W  # noqa: E402

nx, ny, nz = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (10, 10, 150)))
n_or = int(sys.argv[4]) if len(sys.argv) > 4 else 5
GAMMA, RHO = 1.4, 1.225
Q02U0 = 101325.0 * (1200.0 / 300.0 - 1) * math.pi * 0.025**2 * GAMMA / (GAMMA - 1)
hz = 0.5 / nz
mesh = W.kuhn_box((nx, ny, nz), (0, 0, -0.25), (0.05, 0.05, 0.25), jitter=0.1, seed=12345, flame_layer=(nz // 2, nz // 2 + 1))
c = np.where(mesh.points[2, mesh.tetrahedra].sum(axis=1) / 4 < 0, 347.2, 694.4)
x_ref = [0.025, 0.025, -0.6 * hz]
ref_idx = mesh.find_tetrahedron_containing_point(x_ref)
dscrp = {"Interior": ("interior", ()), "Outlet": ("admittance", ("Y", 1e15)),
         "Flame": ("flame", (GAMMA, RHO, Q02U0, ref_idx, x_ref, [0.0, 0.0, 1.0], "n", "τ", 1.0, 0.001))}
L = W.discretize(mesh, dscrp, c)
sol, nit, flag = W.householder(L, 340 * 2 * math.pi, maxiter=20, tol=1e-9 * 340 * 2 * math.pi, output=False)
t0 = time.perf_counter()
sp_, trm, ttm = W.get_surface_points(mesh)
t1 = time.perf_counter()
ctx = W.get_context()
walls, kms = [], []
for _ in range(3):
    t = time.perf_counter()
    sens = W.discrete_adjoint_shape_sensitivity(mesh, dscrp, c, sp_, trm, ttm, L, sol)
    walls.append(time.perf_counter() - t)
    kms.append(ctx.sens_kernel_ms)
out = {"mesh": f"Rijke tube {nx}x{ny}x{nz} Kuhn cubes, P1", "tets": len(mesh.tetrahedra), "points": int(mesh.points.shape[1]),
       "surface_points": int(len(sp_)), "householder": {"iterations": nit, "flag": flag, "omega": [sol.params["ω"].real, sol.params["ω"].imag]},
       "get_surface_points_s": t1 - t0, "wall_s": float(np.median(walls)), "kernel_ms": float(np.median(kms)),
       "points_per_s_wall": len(sp_) / float(np.median(walls)), "finite": bool(np.isfinite(sens).all())}
if n_or > 0:  # checker + CPU cost per point: the oracle's literal loop on a few points
    from oracle import helmholtz as ohelm
    from oracle import nlevp as onlevp
    from oracle import shape as oshape
    from oracle.mesh import Mesh as OMesh
    raw = (mesh.points, [], [list(map(int, t)) for t in mesh.triangles], [list(map(int, t)) for t in mesh.tetrahedra],
           {k: {"dimension": v["dimension"], "simplices": list(map(int, v["simplices"]))} for k, v in mesh.domains.items()})
    mo = OMesh("m", raw=raw)
    Lo = ohelm.discretize(mo, dscrp, c)

    class S:  # the GPU eigenpair is handed to the oracle loop: the comparison isolates the sensitivity evaluation
        params, v, v_adj, eigval = dict(sol.params), sol.v, sol.v_adj, "ω"
    sub = [int(k) for k in np.linspace(0, len(sp_) - 1, n_or)]
    t = time.perf_counter()
    want = oshape.discrete_adjoint_shape_sensitivity(mo, dscrp, c, [int(sp_[k]) for k in sub], [list(map(int, trm[k])) for k in sub],
                                                     [list(map(int, ttm[k])) for k in sub], Lo, S)
    dt = time.perf_counter() - t
    pts = sp_[sub]
    scale = float(np.abs(sens).max())
    out["oracle"] = {"points": n_or, "s_per_point": dt / n_or, "max_abs_diff": float(np.abs(sens[:, pts] - want[:, pts]).max()), "scale": scale,
                     "ok": bool(np.abs(sens[:, pts] - want[:, pts]).max() <= 1e-5 * scale),
                     "speedup_wall_per_point": (dt / n_or) / (float(np.median(walls)) / len(sp_))}
print(json.dumps(out))
