"""Assembly-only stress (BASELINE.json configs[4]): unit-cube Kuhn grid n^3 cubes (n = 203 -> 50 193 162
tetrahedra, 67 419 143 P2 DOFs, nnz ~ 1.94e9), seed-7 jitter, per-tet c ~ U(300, 700), one boundary face as admittance,
flame = 0.1 % of the tetrahedra; K, M, C, Q assembly plus 10 L(z) combines at random z.

    python tools/bench_config5.py [n=203] [reps=5] > gpurun_out/config5.json

Prints one JSON object (phases, kernel times by CUDA events, algorithmic bytes and roofline fractions)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import wae_b200 as W  # noqa: E402
from wae_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 203
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
HBM = 6548.5
try:
    HBM = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
out = {"config": f"unit-cube Kuhn grid {n}^3 cubes, P2, seed-7 jitter, per-tet c~U(300,700), Outlet admittance, flame = 0.1% of the tets",
       "hbm_peak_gbs": HBM, "host_s": {}}


def tick(name, t0):
    out["host_s"][name] = round(time.time() - t0, 2)
    print(f"[config5] {name}: {out['host_s'][name]} s", file=sys.stderr, flush=True)


t = time.time()
mesh = W.kuhn_box((n, n, n), (0, 0, 0), (1, 1, 1), jitter=0.1, seed=7)
tick("mesh generation (python)", t)
t = time.time()
tris, tets, dim = W.aggregate_elements(mesh, "quad")
tick("edge numbering / P2 DOF lists (python)", t)
ntet, npts = len(tets), mesh.points.shape[1]
ctx = W.get_context()
t = time.time()
ctx.mesh_set(2, mesh.points.T, tets, tris, dim)
tick("mesh upload", t)
t = time.time()
pid, nnz = ctx.pattern_build(3, None)
tick("sparsity pattern (host symbolic)", t)
rng = np.random.default_rng(7)
c = rng.uniform(300, 700, ntet)
t = time.time()
im, ik = ctx.assemble_mk(pid, c)  # includes the one-off build of the assembly program (P2: star program)
tick("assembly program (host symbolic) + first M,K assembly", t)
ms = []
for _ in range(reps):
    ctx.assemble_mk(pid, c, reuse=(im, ik))
    ms.append(ctx.last_ms("assemble"))
mk_ms = float(np.median(ms))
alg_mk = ntet * (4 * 10 + 8) + 24 * npts + 2 * nnz * 8
out.update({"tets": ntet, "dofs": int(dim), "points": int(npts), "nnz": int(nnz)})
out["assemble_MK"] = {"kernel_ms": mk_ms, "all_ms": ms, "Mtets_per_s": ntet / mk_ms / 1e3, "algorithmic_GB": alg_mk / 1e9,
                      "achieved_GBs": alg_mk / mk_ms / 1e6, "frac_of_hbm": alg_mk / mk_ms / 1e6 / HBM,
                      "kernel": "assemble_tet_stars<10,3>" if ctx.last_ms("star_patches") > 0 else "assemble_tet_pairs<10,3>"}
if ctx.last_ms("star_patches") > 0:
    out["assemble_MK"]["star_layout"] = {k: ctx.last_ms("star_" + k) for k in ("patches", "staged", "simplices", "sources", "smem", "threads", "ctas_per_sm")}
    out["assemble_MK"]["program_bytes_per_tet"] = ctx.last_ms("star_program_bytes") / ntet
# boundary admittance: the z = 1 face ("Outlet"), per-triangle c = 500
outlet = np.asarray(mesh.domains["Outlet"]["simplices"], dtype=np.int64)
t = time.time()
pidc = ctx.pattern_build(2, outlet)[0]
ic = ctx.assemble(pidc, _lib.OP_BOUNDARY, np.full(len(outlet), 500.0))
c_ms = ctx.last_ms("assemble")
tick("boundary pattern + C assembly", t)
# flame: 0.1 % of the tetrahedra (a contiguous block in storage order), reference tet just outside of it
nfl = max(6, ntet // 1000)
f0 = ntet // 2
flame = np.arange(f0, f0 + nfl, dtype=np.int64)
ref = f0 - 1
x_ref = mesh.points[:, mesh.tetrahedra[ref]].mean(axis=1)
t = time.time()
_, iq, nnzq = ctx.assemble_flame(flame, ref, list(x_ref), [0.0, 0.0, 1.0], 1.0)
q_ms = ctx.last_ms("assemble_flame") or ctx.last_ms("assemble")
tick("Q (flame) assembly", t)
out["assemble_C"] = {"triangles": int(len(outlet)), "kernel_ms": c_ms}
out["assemble_Q"] = {"flame_tets": int(nfl), "nnz": int(nnzq), "kernel_ms": q_ms}
t = time.time()
fid, nnz_u = ctx.family_create([im, ik, ic, iq])
tick("family (union pattern, term maps)", t)
cms = []
for k in range(10):
    z = complex(rng.uniform(500, 5000), rng.uniform(-200, 200))
    coeffs = np.array([z * z, 1.0, z * 1e15 * 0 + z, np.exp(-1j * z * 1e-3)], dtype=np.complex128)
    ctx.combine(fid, coeffs, 0)
    cms.append(ctx.last_ms("combine"))
comb_ms = float(np.median(cms[1:]))
alg_c = nnz_u * (8 + 8 + 16)  # real M, real K read once, complex L(z) written once (sub-pattern terms C, Q are negligible)
out["combine"] = {"nnz_union": int(nnz_u), "kernel_ms": comb_ms, "all_ms": cms, "algorithmic_GB": alg_c / 1e9, "achieved_GBs": alg_c / comb_ms / 1e6,
                  "frac_of_hbm": alg_c / comb_ms / 1e6 / HBM}
try:
    import torch
    out["device_mem_GB"] = torch.cuda.max_memory_allocated() / 1e9  # (library allocations are outside torch: see nvidia-smi)
except Exception:
    pass
print(json.dumps(out))
