"""Tuning sweep of the star assembly kernel (generation 3, csrc/assembly_star.cu) on an n^3-cube Kuhn box: shared-memory budget per CTA
(WAE_STAR_SMEM = patch size and CTAs per SM) and threads per CTA (WAE_STAR_THREADS), every combination timed with CUDA events and checked
against the pair-program kernel (generation 2, WAE_ASM_GEN=2), which is timed beside it.  One line per combination + ONE JSON line at the end.

    python tools/sweep_star.py [ncube=64] [order=quad] [reps=7] [quick]
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
quick = len(sys.argv) > 4
KNOBS = ("WAE_ASM_GEN", "WAE_STAR_SMEM", "WAE_STAR_THREADS", "WAE_STAR_CTAS", "WAE_GATHER_SLOTS", "WAE_GATHER_CTAS", "WAE_GATHER_THREADS", "WAE_ASM_VARIANT")
for k in KNOBS:
    os.environ.pop(k, None)
mesh = W.kuhn_box((n, n, n), (0, 0, 0), (1, 1, 1), jitter=0.1, seed=7)
tris, tets, dim = W.aggregate_elements(mesh, order)
ctx = W.get_context()
ctx.mesh_set(1 if order == "lin" else 2, mesh.points.T, tets, tris, dim)
c = np.random.default_rng(7).uniform(300, 700, len(tets))
pid, nnz = ctx.pattern_build(3, None)
alg = len(tets) * (4 * tets.shape[1] + 8) + 24 * mesh.points.shape[1] + 2 * nnz * 8


def timed(label, extra=()):
    im, ik = ctx.assemble_mk(pid, c)
    ms = []
    for _ in range(reps):
        ctx.assemble_mk(pid, c, reuse=(im, ik))
        ms.append(ctx.last_ms("assemble"))
    vals = (ctx.mat_get(im).copy(), ctx.mat_get(ik).copy())
    ctx.mat_free(im); ctx.mat_free(ik)
    med = float(np.median(ms))
    row = {"kernel": label, "median_ms": med, "best_ms": float(min(ms)), "Gtet_per_s": len(tets) / med / 1e6,
           "hbm_frac_6548": alg / med / 1e6 / 6548.5}
    for k in extra:
        row[k] = ctx.last_ms(k)
    return row, vals


os.environ["WAE_ASM_GEN"] = "2"
ref_row, ref = timed("pairs(gen2)")
print(ref_row, flush=True)
os.environ["WAE_ASM_GEN"] = "3"
scale = (np.abs(ref[0]).max(), np.abs(ref[1]).max())
rows = [ref_row]
extra = ("star_patches", "star_staged", "star_program_bytes", "star_smem", "star_threads", "star_ctas_per_sm")
combos = [(None, None), (230400, 1024), (230400, 768), (230400, 512), (114688, 512), (114688, 384), (114688, 256), (76800, 256), (76800, 320), (57344, 256), (57344, 192), (45056, 192)]
if quick:
    combos = [(None, None), (230400, 1024), (114688, 512), (76800, 256)]
for smem, thr in combos:
    for k, v in (("WAE_STAR_SMEM", smem), ("WAE_STAR_THREADS", thr)):
        os.environ.pop(k, None)
        if v is not None:
            os.environ[k] = str(v)
    try:
        row, vals = timed("stars(gen3)", extra)
        err = max(np.abs(vals[0] - ref[0]).max() / scale[0], np.abs(vals[1] - ref[1]).max() / scale[1])
        row.update({"smem_budget": smem, "threads_env": thr, "max_rel_diff_vs_gen2": float(err), "ok": bool(err <= 1e-12),
                    "program_bytes_per_tet": row["star_program_bytes"] / len(tets), "staged_per_tet": row["star_staged"] / len(tets)})
    except Exception as e:  # noqa: BLE001 -- a combination that does not fit / launch is reported, not fatal
        row = {"kernel": "stars(gen3)", "smem_budget": smem, "threads_env": thr, "error": repr(e)[:200]}
    rows.append(row)
    print(row, flush=True)
good = sorted((r for r in rows if r.get("ok")), key=lambda r: r["median_ms"])
print(json.dumps({"ncube": n, "order": order, "tets": len(tets), "nnz": int(nnz), "algorithmic_bytes": alg, "gen2": ref_row, "best": good[:6],
                  "n_failed": len(rows) - 1 - len(good)}))
