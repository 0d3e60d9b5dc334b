"""Assembly-only timing on a synthetic Kuhn box (SURVEY 8d config 5 at a chosen size).
    python tools/bench_assembly.py [ncube=60] [order=quad] [reps=5]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import wae_b200 as W  # noqa: E402
from wae_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
order = sys.argv[2] if len(sys.argv) > 2 else "quad"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
t0 = time.time()
mesh = W.kuhn_box((n, n, n), (0, 0, 0), (1, 1, 1), jitter=0.1, seed=7)
tris, tets, dim = W.aggregate_elements(mesh, order)
t1 = time.time()
ctx = W.get_context()
ctx.mesh_set(1 if order == "lin" else 2, mesh.points.T, tets, tris, dim)
t2 = time.time()
pid, nnz = ctx.pattern_build(3, None)
t3 = time.time()
c = np.random.default_rng(7).uniform(300, 700, len(tets))
im, ik = ctx.assemble_mk(pid, c)  # includes the one-off gather-program build
t4 = time.time()
ntet = len(tets)
npts = mesh.points.shape[1]
nloc = tets.shape[1]
alg_bytes = ntet * (4 * nloc + 8) + 24 * npts + 2 * nnz * 8
print(f"mesh {ntet} tets, {dim} dofs, nnz {nnz}; host: mesh {t1-t0:.2f}s upload {t2-t1:.2f}s pattern {t3-t2:.2f}s gather-program+first {t4-t3:.2f}s")
ms = []
for r in range(reps):
    ctx.mat_free(im); ctx.mat_free(ik)
    im, ik = ctx.assemble_mk(pid, c)
    ms.append(ctx.last_ms("assemble"))
best, med = min(ms), sorted(ms)[len(ms) // 2]
print(f"M+K kernel: best {best:.3f} ms median {med:.3f} ms -> {ntet/med/1e6:.3f} Gtet/s, algorithmic {alg_bytes/1e9:.3f} GB -> {alg_bytes/med/1e6:.1f} GB/s "
      f"({alg_bytes/med/1e6/6548.5*100:.1f}% of measured HBM peak 6548.5 GB/s); {alg_bytes/ntet:.0f} B/tet")
os.environ["WAE_FORCE_ATOMIC"] = "1"
ctx.mat_free(im); ctx.mat_free(ik)
im, ik = ctx.assemble_mk(pid, c)
ms = []
for r in range(3):
    ctx.mat_free(im); ctx.mat_free(ik)
    im, ik = ctx.assemble_mk(pid, c)
    ms.append(ctx.last_ms("assemble"))
print(f"atomic generation: median {sorted(ms)[1]:.3f} ms -> {ntet/sorted(ms)[1]/1e6:.3f} Gtet/s")
