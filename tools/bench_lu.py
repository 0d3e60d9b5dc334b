"""LU scaling probe on a synthetic Rijke-like tube (config 2 geometry at a chosen resolution).
    python tools/bench_lu.py nx ny nz [order=quad] [nsolve=5]
"""
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import wae_b200 as W  # noqa: E402

nx, ny, nz = (int(a) for a in sys.argv[1:4])
order = sys.argv[4] if len(sys.argv) > 4 else "quad"
nsolve = int(sys.argv[5]) if len(sys.argv) > 5 else 5
t0 = time.time()
hz = 0.5 / nz
mesh = W.kuhn_box((nx, ny, nz), (0, 0, -0.25), (0.05, 0.05, 0.25), jitter=0.1, seed=12345, flame_layer=(nz // 2, nz // 2 + 1))
c = mesh.generate_field(lambda x, y, z: 347.2 if z < 0 else 694.4)
gam, rho = 1.4, 1.225
q = 101325.0 * 3 * math.pi * 0.025**2 * gam / (gam - 1)
dscrp = {"Interior": ("interior", ()), "Outlet": ("admittance", ("Y", 1e15)),
         "Flame": ("flame", (gam, rho, q, [0.025, 0.025, -0.6 * hz], [0, 0, 1.0], "n", "τ", 1.0, 0.001))}
t1 = time.time()
L = W.discretize(mesh, dscrp, c, order=order)
t2 = time.time()
dev = L.device()
ctx = dev.ctx
print(f"mesh {len(mesh.tetrahedra)} tets, dim {L.size()}, nnz {dev.nnz}; host mesh {t1-t0:.1f}s discretize {t2-t1:.1f}s (kernels {L.discretization.timing})", flush=True)
t3 = time.time()
lid = dev.lu()
t4 = time.time()
print(f"symbolic {t4-t3:.1f}s: nnz(L+U) {dev.lu_nnz:.3e} fill {dev.lu_nnz/dev.nnz:.1f} flops {dev.lu_flops:.3e}", flush=True)
z = 340 * 2 * math.pi
L(z).materialize(0)
for rep in range(int(os.environ.get('WAE_NFACTOR', '3'))):
    ctx.lu_factor(lid, 0)
    ms = ctx.last_ms("factor")
    symf = 0.5 if ctx.last_ms("factor_sym") > 0.5 else 1.0
    print(f"factor {ms:.1f} ms -> {symf*dev.lu_flops/ms/1e9:.2f} TFLOP/s (fp64, 8 flops per complex multiply-add, {'symmetric' if symf < 1 else 'general'} elimination); static pivots {ctx.last_ms('static_pivots')}", flush=True)
rng = np.random.default_rng(0)
b = rng.standard_normal(L.size()) + 1j * rng.standard_normal(L.size())
for rep in range(nsolve):
    x = ctx.lu_solve(lid, b)
    ms = ctx.last_ms("solve")
r = L(z).matvec(x) - b
print(f"solve {ms:.1f} ms (1 rhs, 1 refinement step; {2*16*dev.lu_nnz/ms/1e6:.0f} GB/s of factor traffic), residual max|Ax-b|/max|b| = {np.abs(r).max()/np.abs(b).max():.2e}")
if os.environ.get("WAE_PROBE_ONLY") == "1":
    sys.exit(0)
B = rng.standard_normal((L.size(), 8)) + 1j * rng.standard_normal((L.size(), 8))
for k in (2, 8):
    for rep in range(3):
        X = ctx.lu_solve(lid, B[:, :k])
    ms = ctx.last_ms("solve")
    r = L(z).matvec(X[:, k - 1]) - B[:, k - 1]
    print(f"solve {k} rhs {ms:.1f} ms ({2*16*dev.lu_nnz/ms/1e6:.0f} GB/s of factor traffic, {k*2*16*dev.lu_nnz/ms/1e6:.0f} GB/s x rhs), residual {np.abs(r).max()/np.abs(B[:, k - 1]).max():.2e}")
if os.environ.get("WAE_PROBE_ONLY") == "2":
    sys.exit(0)
st = {}
t5 = time.time()
sol, n, flag = W.householder(L, z, maxiter=15, tol=1e-9 * abs(z), output=True, stats=st)
t6 = time.time()
print(f"householder: {n} iterations, flag {flag}, omega {sol.params['ω']}, {t6-t5:.2f} s wall, stats {st}")
