"""Time the opt-in variants of the M/K assembly kernel (WAE_ASM_VARIANT, csrc/assembly_kernels.cu) against the default kernel on an
n^3-cube P2 (or P1) Kuhn box and check that they produce the same matrices, then a SHORT layout sweep (patch size x CTAs per SM, the
first item of DESIGN section 9: several smaller CTAs per SM whose phases drift apart; tools/sweep_assembly.py is the full sweep).
Prints ONE JSON line.  bench.py runs this in a subprocess after its own (default-kernel) assembly measurement, so that a failing
variant cannot disturb the bench.

    python tools/bench_assembly_variants.py [ncube=64] [order=quad] [reps=7]
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import wae_b200 as W  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
order = sys.argv[2] if len(sys.argv) > 2 else "quad"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 7
os.environ.pop("WAE_ASM_VARIANT", None)
mesh = W.kuhn_box((n, n, n), (0, 0, 0), (1, 1, 1), jitter=0.1, seed=7)
tris, tets, dim = W.aggregate_elements(mesh, order)
ctx = W.get_context()
ctx.mesh_set(1 if order == "lin" else 2, mesh.points.T, tets, tris, dim)
pid, nnz = ctx.pattern_build(3, None)
c = np.random.default_rng(7).uniform(300, 700, len(tets))
im, ik = ctx.assemble_mk(pid, c)
ref = (ctx.mat_get(im).copy(), ctx.mat_get(ik).copy())
scale = (np.abs(ref[0]).max(), np.abs(ref[1]).max())
out = {"ncube": n, "order": order, "tets": len(tets), "nnz": int(nnz), "reps": reps, "variants": {}}
for var in (0, 1, 2, 3, 0):
    os.environ["WAE_ASM_VARIANT"] = str(var)
    ms = []
    for _ in range(reps):
        ctx.assemble_mk(pid, c, reuse=(im, ik))
        ms.append(ctx.last_ms("assemble"))
    err = max(np.abs(ctx.mat_get(im) - ref[0]).max() / scale[0], np.abs(ctx.mat_get(ik) - ref[1]).max() / scale[1])
    key = str(var) if str(var) not in out["variants"] else str(var) + "_again"
    out["variants"][key] = {"median_ms": float(np.median(ms)), "best_ms": float(min(ms)), "Gtet_per_s": len(tets) / float(np.median(ms)) / 1e6,
                            "max_rel_diff_vs_default": float(err), "ok": bool(err <= 1e-13)}
# short layout sweep: (slots, CTAs per SM) x {variant 0, variant 3}; a fresh pattern per layout (the pair program is built on first use with
# the slot budget of the environment); every combination checked against the default kernel's matrices; failures are recorded, not fatal
KNOBS = ("WAE_GATHER_SLOTS", "WAE_GATHER_CTAS", "WAE_GATHER_THREADS", "WAE_ASM_VARIANT")
out["layouts"] = []
for slots, ctas in ((6144, 2), (4096, 3), (3072, 4)):
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ["WAE_GATHER_SLOTS"], os.environ["WAE_GATHER_CTAS"] = str(slots), str(ctas)
    try:
        pid2, _ = ctx.pattern_build(3, None)
        for var in (0, 3):
            os.environ["WAE_ASM_VARIANT"] = str(var)
            jm, jk = ctx.assemble_mk(pid2, c)
            ms = []
            for _ in range(reps):
                ctx.assemble_mk(pid2, c, reuse=(jm, jk))
                ms.append(ctx.last_ms("assemble"))
            err = max(np.abs(ctx.mat_get(jm) - ref[0]).max() / scale[0], np.abs(ctx.mat_get(jk) - ref[1]).max() / scale[1])
            ctx.mat_free(jm)
            ctx.mat_free(jk)
            out["layouts"].append({"slots": slots, "ctas_per_sm": ctas, "variant": var, "median_ms": float(np.median(ms)), "best_ms": float(min(ms)),
                                   "Gtet_per_s": len(tets) / float(np.median(ms)) / 1e6, "max_rel_diff_vs_default": float(err), "ok": bool(err <= 1e-12)})
    except Exception as e:  # noqa: BLE001 -- diagnostic only
        out["layouts"].append({"slots": slots, "ctas_per_sm": ctas, "error": repr(e)[:200]})
for k in KNOBS:
    os.environ.pop(k, None)
print(json.dumps(out))
