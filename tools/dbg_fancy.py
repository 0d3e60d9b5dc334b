import math, os, sys
import numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import wae_b200 as W
from cases import GAMMA, N_REF, Q02U0, RHO, X_REF, load_raw_mesh, speedofsound
raw = load_raw_mesh("rijke_mm")
mg = W.Mesh("Rijke_mm.msh", scale=0.001, raw=raw)
A, B, Cm, D = np.array([[-2.0e3, 0.0], [0.0, -5.0e3]]), np.array([1.0, 1.0]), np.array([4.0e3, -1.0e3]), np.array([0.5])
cases = {
    "scalar": {"Interior": ("interior", ()), "Outlet": ("admittance", (A, B, Cm, D)),
               "Flame": ("fancyflame", (GAMMA, RHO, Q02U0, X_REF, N_REF, "n", "τ", "a", 1.0, 0.001, -1e-8))},
    "summed": {"Interior": ("interior", ()), "Outlet": ("admittance", ("Y", 1e15)),
               "Flame": ("fancyflame", (GAMMA, RHO, Q02U0, X_REF, N_REF, ("n1", "n2"), ("τ1", "τ2"), ("a1", "a2"),
                                        (0.6, 0.4), (0.001, 0.0013), (-1e-8, -2e-8)))},
}
for env in ({}, {"WAE_LU_WININV": "0"}, {"WAE_EIGS_REFINE": "1"}, {"WAE_LU_SOLVE_FUSED": "0"}, {"WAE_LU_SOLVE_UPD2": "0"}):
    for k in ("WAE_LU_WININV", "WAE_EIGS_REFINE", "WAE_LU_SOLVE_FUSED", "WAE_LU_SOLVE_UPD2"):
        os.environ.pop(k, None)
    os.environ.update(env)
    for name, dscrp in cases.items():
        Lg = W.discretize(mg, dscrp, mg.generate_field(speedofsound))
        print("=====", env, name, flush=True)
        sg, ng, fg = W.householder(Lg, 340 * 2 * math.pi, maxiter=25, tol=1e-11, output=True)
        print("flag", fg, "n", ng, flush=True)
        Lg.release()
