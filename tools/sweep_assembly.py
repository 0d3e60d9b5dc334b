"""Tuning sweep of the M/K assembly kernel (csrc/assembly_kernels.cu) on an n^3-cube P2 Kuhn box: patch size (WAE_GATHER_SLOTS), CTAs per
SM / threads per CTA (WAE_GATHER_CTAS, WAE_GATHER_THREADS) and the kernel variants (WAE_ASM_VARIANT), every combination timed with CUDA
events and checked against the default configuration.  One line per combination + ONE JSON line at the end (best first).

    python tools/sweep_assembly.py [ncube=64] [order=quad] [reps=5]

The defaults (12288 slots, 1 CTA x 1024 threads, variant 0) are the measured production kernel; nothing here changes them.
"""
import itertools
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import wae_b200 as W  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
order = sys.argv[2] if len(sys.argv) > 2 else "quad"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
KNOBS = ("WAE_GATHER_SLOTS", "WAE_GATHER_CTAS", "WAE_GATHER_THREADS", "WAE_ASM_VARIANT")
for k in KNOBS:
    os.environ.pop(k, None)
mesh = W.kuhn_box((n, n, n), (0, 0, 0), (1, 1, 1), jitter=0.1, seed=7)
tris, tets, dim = W.aggregate_elements(mesh, order)
ctx = W.get_context()
ctx.mesh_set(1 if order == "lin" else 2, mesh.points.T, tets, tris, dim)
c = np.random.default_rng(7).uniform(300, 700, len(tets))
pid0, nnz = ctx.pattern_build(3, None)
im, ik = ctx.assemble_mk(pid0, c)
ref = (ctx.mat_get(im).copy(), ctx.mat_get(ik).copy())
scale = (np.abs(ref[0]).max(), np.abs(ref[1]).max())
ctx.mat_free(im); ctx.mat_free(ik)
alg = len(tets) * (4 * tets.shape[1] + 8) + 24 * mesh.points.shape[1] + 2 * nnz * 8
rows = []
# (slots, CTAs per SM): a CTA's shared memory is 16 B per slot + two program/coordinate buffers, so k CTAs need roughly slots <= 12288 / k
layouts = [(None, None), (12288, 1), (9216, 1), (6144, 1), (6144, 2), (5120, 2), (4096, 2), (4096, 3), (3072, 3), (3072, 4), (2560, 4)]
for slots, ctas in layouts:
    for k in KNOBS:
        os.environ.pop(k, None)
    if slots is not None:
        os.environ["WAE_GATHER_SLOTS"] = str(slots)
    pid, _ = ctx.pattern_build(3, None)  # a fresh pattern: the pair program is built on first use with the slot budget above
    for var, thr in itertools.product((0, 1, 2, 3), (None, 512, 768) if ctas == 1 else (None,)):
        os.environ["WAE_ASM_VARIANT"] = str(var)
        for k, v in (("WAE_GATHER_CTAS", ctas), ("WAE_GATHER_THREADS", thr)):
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = str(v)
        try:
            im, ik = ctx.assemble_mk(pid, c)
            ms = []
            for _ in range(reps):
                ctx.assemble_mk(pid, c, reuse=(im, ik))
                ms.append(ctx.last_ms("assemble"))
            err = max(np.abs(ctx.mat_get(im) - ref[0]).max() / scale[0], np.abs(ctx.mat_get(ik) - ref[1]).max() / scale[1])
            ctx.mat_free(im); ctx.mat_free(ik)
            med = float(np.median(ms))
            row = {"slots": slots, "ctas": ctas, "threads": thr, "variant": var, "median_ms": med, "best_ms": float(min(ms)),
                   "Gtet_per_s": len(tets) / med / 1e6, "hbm_frac_6548": alg / med / 1e6 / 6548.5, "max_rel_diff": float(err), "ok": bool(err <= 1e-12)}
        except Exception as e:  # noqa: BLE001 -- a combination that does not fit / launch is reported, not fatal
            row = {"slots": slots, "ctas": ctas, "threads": thr, "variant": var, "error": repr(e)[:200]}
        rows.append(row)
        print(row, flush=True)
good = sorted((r for r in rows if r.get("ok")), key=lambda r: r["median_ms"])
print(json.dumps({"ncube": n, "order": order, "tets": len(tets), "nnz": int(nnz), "algorithmic_bytes": alg, "best": good[:8], "n_failed": len(rows) - len(good)}))
