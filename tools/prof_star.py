"""Short driver of the star assembly kernel for ncu and for phase timing (WAE_STAR_DBG): n^3-cube Kuhn box, a few M+K launches.

    python tools/prof_star.py [ncube=64] [order=quad] [reps=5] [phases]
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import wae_b200 as W  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
order = sys.argv[2] if len(sys.argv) > 2 else "quad"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
phases = len(sys.argv) > 4
os.environ["WAE_ASM_GEN"] = "3"
mesh = W.kuhn_box((n, n, n), (0, 0, 0), (1, 1, 1), jitter=0.1, seed=7)
tris, tets, dim = W.aggregate_elements(mesh, order)
ctx = W.get_context()
ctx.mesh_set(1 if order == "lin" else 2, mesh.points.T, tets, tris, dim)
c = np.random.default_rng(7).uniform(300, 700, len(tets))
pid, nnz = ctx.pattern_build(3, None)
alg = len(tets) * (4 * tets.shape[1] + 8) + 24 * mesh.points.shape[1] + 2 * nnz * 8
out = {"ncube": n, "order": order, "tets": len(tets), "nnz": int(nnz), "algorithmic_bytes": alg}
for dbg in ((0, 6, 4, 2, 5, 3, 1, 7) if phases else (0,)):
    os.environ["WAE_STAR_DBG"] = str(dbg)
    im, ik = ctx.assemble_mk(pid, c)
    ms = []
    for _ in range(reps):
        ctx.assemble_mk(pid, c, reuse=(im, ik))
        ms.append(ctx.last_ms("assemble"))
    ctx.mat_free(im); ctx.mat_free(ik)
    out["dbg%d_ms" % dbg] = float(np.median(ms))
os.environ.pop("WAE_STAR_DBG", None)
out["hbm_frac_6548"] = alg / out["dbg0_ms"] / 1e6 / 6548.5
print(json.dumps(out))
