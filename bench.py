#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 hot path (contract: see the task statement / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): synthetic Rijke tube 0.05 x 0.05 x 0.5 m, Kuhn grid 20 x 20 x 300 cubes
-> 720 000 tetrahedra, 1 010 281 P2 DOFs, interior + outlet admittance (Y=1e15) + n-tau flame (n=1, tau=1 ms).
One STEP = numeric re-assembly of all operators (M, K, C, Q, aux) on the GPU followed by householder (Newton,
order 1) from 340*2*pi to |d omega| <= 1e-9 |omega|  ->  one eigenpair.  `value` = eigenpairs/s with mesh,
patterns and the symbolic LU resident on the device; `e2e` additionally uploads the vertex coordinates and the
speed-of-sound field from pinned host memory every step (the eigenvector always returns to the host).
With N > 1 every rank solves its own replica (tau differs per rank): householder does not shard
("replicas only", weak scaling).  The path that does shard -- Beyn's quadrature nodes -- is timed on every N
as well and reported under "beyn" (strong scaling: 128 nodes in total, one NCCL all-reduce of the moments).

Extra objects on the JSON line: roofline (numeric LU = DMMA ZGEMM, FP64 tensor pipe), roofline_assembly (HBM),
assembly (Mtets/s on a larger box), beyn, cpu_baseline (the scipy/SuperLU oracle on a bounded sample), clocks.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GAMMA, RHO = 1.4, 1.225
Q02U0 = 101325.0 * (1200.0 / 300.0 - 1) * math.pi * 0.025**2 * GAMMA / (GAMMA - 1)
Z0 = 340 * 2 * math.pi


def tube_case(W, ncube, order="quad"):
    nx, ny, nz = ncube
    hz = 0.5 / nz
    mesh = W.kuhn_box(ncube, (0, 0, -0.25), (0.05, 0.05, 0.25), jitter=0.1, seed=12345, flame_layer=(nz // 2, nz // 2 + 1),
                      name=f"rijke_kuhn_{nx}x{ny}x{nz}")
    c = mesh.generate_field(lambda x, y, z: 347.2 if z < 0 else 694.4) if len(mesh.tetrahedra) < 20000 else \
        np.where(mesh.points[2, mesh.tetrahedra].sum(axis=1) / 4 < 0, 347.2, 694.4)
    dscrp = {"Interior": ("interior", ()), "Outlet": ("admittance", ("Y", 1e15)),
             "Flame": ("flame", (GAMMA, RHO, Q02U0, [0.025, 0.025, -0.6 * hz], [0.0, 0.0, 1.0], "n", "τ", 1.0, 0.001))}
    return mesh, c, dscrp


def oracle_step(ncube):
    """One eigenpair with the CPU oracle (numpy/scipy restatement of the reference path) on a reduced tube.  Returns the wall time and
    its split: "asm" (discretize: Python element loops + sparse()), "factor" (the SuperLU factorisations inside the shift-invert
    ARPACK calls, timed by wrapping scipy's splu) and "other" (triangular solves, Arnoldi, L(z) sums)."""
    import importlib

    import wae_b200 as W
    from oracle.helmholtz import discretize as odisc
    from oracle.mesh import Mesh as OMesh
    from oracle.nlevp import householder as ohouse
    m, _, dscrp = tube_case(W, ncube)
    raw = (m.points, [], [list(map(int, t)) for t in m.triangles], [list(map(int, t)) for t in m.tetrahedra],
           {k: {"dimension": v["dimension"], "simplices": list(map(int, v["simplices"]))} for k, v in m.domains.items()})
    ar = importlib.import_module("scipy.sparse.linalg._eigen.arpack.arpack")
    acc = {"factor": 0.0, "n_fact": 0}
    orig = ar.splu

    def timed_splu(*a, **k):
        t = time.perf_counter()
        lu = orig(*a, **k)
        acc["factor"] += time.perf_counter() - t
        acc["n_fact"] += 1
        return lu
    ar.splu = timed_splu
    try:
        t0, c0 = time.perf_counter(), time.process_time()
        mo = OMesh("m", raw=raw)
        c = mo.generate_field(lambda x, y, z: 347.2 if z < 0 else 694.4)
        L = odisc(mo, dscrp, c, order="quad")
        t1 = time.perf_counter()
        sol, n, flag = ohouse(L, Z0, maxiter=15, tol=1e-9 * Z0)
        t2 = time.perf_counter()
    finally:
        ar.splu = orig
    return {"total": t2 - t0, "asm": t1 - t0, "factor": acc["factor"], "other": (t2 - t1) - acc["factor"], "n_fact": acc["n_fact"],
            "dim": L.size(), "ntet": len(m.tetrahedra), "nit": n, "threads": max(1, round((time.process_time() - c0) / (t2 - t0)))}


def lu_cost(W, mesh):
    """(factorisation flops, nnz(L+U), tetrahedra) of the P2 operator pattern of `mesh` under one and the same fill-reducing ordering
    (the host-only symbolic phase of libwae_b200, nested dissection; needs no GPU): the size measure the CPU sample is scaled with."""
    import ctypes as C

    import scipy.sparse as sp

    from wae_b200 import _lib
    _, tets, dim = W.aggregate_elements(mesh, "quad")
    n = tets.shape[1]
    I = np.repeat(tets, n, axis=1).ravel().astype(np.int32)
    J = np.tile(tets, (1, n)).ravel().astype(np.int32)
    A = sp.csc_matrix((np.ones(len(I), dtype=np.int8), (I, J)), shape=(dim, dim))
    A.sum_duplicates()
    A.sort_indices()
    xyz = np.ascontiguousarray(np.concatenate([mesh.points, 0.5 * (mesh.points[:, mesh.lines[:, 0]] + mesh.points[:, mesh.lines[:, 1]])], axis=1).T)
    f = _lib.lib().wae_lu_symbolic_stats
    P64, PD = C.POINTER(C.c_int64), C.POINTER(C.c_double)
    f.restype, f.argtypes = C.c_int32, [C.c_int64, P64, P64, PD, C.c_int32, PD]
    cp, rv = A.indptr.astype(np.int64), A.indices.astype(np.int64)
    out = np.zeros(8)
    if f(dim, cp.ctypes.data_as(P64), rv.ctypes.data_as(P64), xyz.ctypes.data_as(PD), 64, out.ctypes.data_as(PD)) != 0:
        raise RuntimeError("wae_lu_symbolic_stats failed")
    return float(out[2]), float(out[1]), len(tets)


def cpu_scaled(W, st, sample, tube):
    """Scale the measured CPU sample to the metric's unit (eigenpairs/s ON THE FULL WORKLOAD): every phase of the sample by the
    growth of the quantity it is proportional to -- assembly by the number of tetrahedra, the factorisations by the factorisation
    flop count, the rest (triangular solves, Arnoldi) by nnz(L+U) -- with both counts taken from the same nested-dissection symbolic
    analysis of the two patterns (SuperLU's own COLAMD fill grows faster, so this favours the CPU).  Same iteration count assumed."""
    fs, ns, ts = lu_cost(W, tube_case(W, sample)[0])
    ff, nf, tf = lu_cost(W, tube_case(W, tube)[0])
    est = st["asm"] * tf / ts + st["factor"] * ff / fs + st["other"] * nf / ns
    return 1.0 / est, {"sample_s": {k: st[k] for k in ("total", "asm", "factor", "other")}, "growth": {"tets": tf / ts, "factor_flops": ff / fs, "factor_nnz": nf / ns},
                       "estimated_full_size_s": est, "sample_eigenpairs_per_s": 1.0 / st["total"]}


T_START = time.perf_counter()
DIAG_START_BY, DIAG_END_BY = 420.0, 660.0  # seconds since process start


def diag_leg(tool, argv, timeout, env=None):
    """Run tools/<tool> in a subprocess and return the JSON object of its last JSON line; a failure, a time-out or an exhausted wall-clock
    budget is recorded in the result and never disturbs the bench itself."""
    elapsed = time.perf_counter() - T_START
    if elapsed > DIAG_START_BY:
        return {"skipped": f"diagnostic budget: {elapsed:.0f} s into the run"}
    try:
        p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool), *argv], capture_output=True, text=True,
                           timeout=max(30.0, min(timeout, DIAG_END_BY - elapsed)), env=env)
        line = [l for l in p.stdout.splitlines() if l.startswith("{")]
        return json.loads(line[-1]) if line else {"error": (p.stderr or p.stdout)[-400:]}
    except Exception as e:  # noqa: BLE001 -- diagnostic only
        return {"error": repr(e)[:400]}


def W_host():
    """The package for host-only use (mesh generators, host diagnostics); importing it needs no GPU."""
    import wae_b200 as W
    return W


class Clocks:
    """nvidia-smi sampler running during the timed region."""

    def __init__(self, dev):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(dev), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        os.unlink(self.f.name)
        return out


def ncu_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of a kernel, from the committed `ncu --set full` summaries
    (profiles/r01_traffic.json: kernel/workload -> bytes, with the capture it was read from); None if not captured."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))[key]
        return t["dram_bytes_per_launch"] / 1e9, t["source"]
    except Exception:
        return None, None


def fp64_peak(torch, dev):
    """Measured FP64 GEMM peak of this GPU (cuBLAS via torch): DGEMM 8192^3 and ZGEMM 4096^3, best of 5, TFLOP/s."""
    best = {}
    for name, n, dt, fl in (("dgemm", 8192, torch.float64, 2.0), ("zgemm", 4096, torch.complex128, 8.0)):
        a = torch.randn(n, n, dtype=dt, device=dev)
        b = torch.randn(n, n, dtype=dt, device=dev)
        torch.matmul(a, b)
        torch.cuda.synchronize()
        t = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            t.append(e0.elapsed_time(e1))
        best[name] = fl * n**3 / (min(t) * 1e-3) / 1e12
        del a, b
    torch.cuda.empty_cache()
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--tube", default="20,20,300", help="cubes of the Rijke tube grid (config 2: 20,20,300)")
    ap.add_argument("--beyn-box", default="32,32,64", help="cubes of the P1 tube used for the sharded Beyn leg")
    ap.add_argument("--beyn-edge-nodes", type=int, default=32, help="Gauss-Legendre nodes per polygon edge (4 edges -> 128 nodes)")
    ap.add_argument("--assembly-cubes", type=int, default=64, help="n for the n^3-cube P2 assembly-only leg (0 = skip)")
    ap.add_argument("--cpu-sample", default="4,4,60")
    ap.add_argument("--ref-sample", default="4,4,60")
    ap.add_argument("--skip-extras", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    tube = tuple(int(x) for x in args.tube.split(","))
    workload = (f"Rijke tube {tube[0]}x{tube[1]}x{tube[2]} Kuhn cubes, P2, interior+outlet admittance+n-tau flame; "
                "step = GPU re-assembly + householder(order 1) to |dw|<=1e-9|w|")

    # ------------------------------------------------------------------ reference arm: the CPU oracle
    if args.impl == "reference":
        if rank != 0:
            return
        sample = tuple(int(x) for x in args.ref_sample.split(","))
        for _ in range(args.warmup):
            oracle_step(sample)
        t0 = time.perf_counter()
        acc = None
        for _ in range(args.steps):
            st = oracle_step(sample)
            acc = st if acc is None else {k: (acc[k] + st[k] if k in ("total", "asm", "factor", "other") else st[k]) for k in st}
        dt = time.perf_counter() - t0
        for k in ("total", "asm", "factor", "other"):
            acc[k] /= args.steps
        val, scaling = cpu_scaled(W_host(), acc, sample, tube)
        smp = (f"reference cannot run here (pure Julia, no julia binary): CPU restatement (numpy + scipy SuperLU/ARPACK, not "
               f"Julia/UMFPACK) timed on a bounded sample of the workload: tube {sample[0]}x{sample[1]}x{sample[2]} cubes, {acc['ntet']} tets, "
               f"{acc['dim']} P2 DOFs, assembly + householder to convergence per step ({acc['total']:.2f} s, {acc['n_fact']} factorisations); "
               f"value = that time scaled phase by phase to the full workload (assembly x tets, factorisations x factorisation flops, "
               f"rest x nnz(L+U); see cpu_baseline.scaling) -- an ESTIMATE of the full-size CPU rate, the full-size CPU run itself takes hours")
        print(json.dumps({"impl": "reference", "metric": "NLEVP eigenpairs/s (householder, config 2)", "value": val, "unit": "eigenpairs/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
                          "config": {"workload": workload, "sample": smp},
                          "cpu_baseline": {"value": val, "unit": "eigenpairs/s", "cores": acc["threads"], "kind": "port", "sample": smp, "extrapolated": True,
                                           "cores_note": f"process CPU time / wall time of the sample (SuperLU is serial, OpenBLAS threads the dense kernels); host has {os.cpu_count()} cores",
                                           "scaling": scaling},
                          "e2e": {"value": val, "unit": "eigenpairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist

    import wae_b200 as W
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ctx = W.get_context(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    t_setup = time.perf_counter()
    mesh, c, dscrp = tube_case(W, tube)
    tau = 0.001 * (1.0 + 0.05 * rank)
    L = W.discretize(mesh, dscrp, c, order="quad")
    disc = L.discretization
    dv = L.device()
    lu = dv.lu()
    t_setup = time.perf_counter() - t_setup
    npts, ntet = mesh.points.shape[1], len(mesh.tetrahedra)
    pts_pinned = torch.from_numpy(np.ascontiguousarray(mesh.points.T)).pin_memory()
    c_pinned = torch.from_numpy(np.ascontiguousarray(c)).pin_memory()
    tol = 1e-9 * Z0

    def step(e2e, stats):
        if e2e:
            ctx.mesh_update_points(pts_pinned.numpy())
        ms = disc.reassemble(c_pinned.numpy())
        stats["assemble_ms"] = stats.get("assemble_ms", 0.0) + ms
        L.params["n"], L.params["τ"] = 1.0 + 0j, complex(tau)
        sol, n, flag = W.householder(L, Z0, maxiter=15, tol=tol, output=False, stats=stats)
        stats["iterations"] = stats.get("iterations", 0) + n
        if flag < 0:
            raise RuntimeError(f"householder failed with flag {flag}")
        return sol

    for _ in range(args.warmup):
        sol = step(False, {})
    omega = sol.params["ω"]

    def timed(e2e):
        stats = {}
        barrier()
        l0 = ctx.launch_count()
        clk = Clocks(local) if rank == 0 else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step(e2e, stats)
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        return ms, stats, ctx.launch_count() - l0, (clk.stop() if clk else None)

    ms, stats, launches, clocks = timed(False)
    ms_e2e, stats_e2e, _, _ = timed(True)
    value = world * args.steps / (ms * 1e-3)
    value_e2e = world * args.steps / (ms_e2e * 1e-3)
    nfac = stats["factorizations"]
    sym = ctx.last_ms("factor_sym") > 0.5   # complex-symmetric elimination (LDL^T-type): half of the LU multiply-adds
    fac_flops = dv.lu_flops * (0.5 if sym else 1.0)
    fac_tflops = fac_flops * nfac / (stats["factor_ms"] * 1e-3) / 1e12
    iters = stats["iterations"]
    h2d = 3 * npts * 8 + ntet * 8 + (iters / args.steps) * 2 * 16 * dv.dim  # points + c + Krylov start vectors per iteration
    d2h = (iters / args.steps) * 2 * 16 * dv.dim                              # eigenvector pairs back to the host

    out = {"metric": "NLEVP eigenpairs/s (householder, config 2)", "value": value, "unit": "eigenpairs/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "c128", "data": "synthetic",
           "config": {"workload": workload, "tets": ntet, "dofs": dv.dim, "nnz": dv.nnz, "factor_nnz": dv.lu_nnz,
                      "factor_flops": dv.lu_flops, "parallelism": "replicas only (householder does not shard); Beyn nodes sharded under 'beyn'",
                      "l2": "working set (LU factors, %.1f GB) is far larger than the 126 MB L2; no flush needed" % (dv.lu_nnz * 16 / 1e9),
                      "setup_s_not_timed": t_setup, "omega": [omega.real, omega.imag]},
           "e2e": {"value": value_e2e, "unit": "eigenpairs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
           "gpu_launches": int(launches), "clocks": clocks,
           "phases_ms_per_step_e2e": {"assemble": stats_e2e["assemble_ms"] / args.steps, "numeric_lu": stats_e2e["factor_ms"] / args.steps,
                                      "eigs_wall": 1e3 * stats_e2e["eigs_wall_s"] / args.steps, "iterations": stats_e2e["iterations"] / args.steps,
                                      "solves": stats_e2e["solves"] / args.steps},
           "phases_ms_per_step": {"assemble": stats["assemble_ms"] / args.steps, "numeric_lu": stats["factor_ms"] / args.steps,
                                  "eigs_wall": 1e3 * stats["eigs_wall_s"] / args.steps, "perturb_wall": 1e3 * stats["perturb_wall_s"] / args.steps,
                                  "iterations": iters / args.steps, "factorizations": nfac / args.steps, "solves": stats["solves"] / args.steps}}

    # ------------------------------------------------------------------ roofline of the dominant kernel
    peaks = fp64_peak(torch, dev)
    peak = max(peaks.values())
    out["roofline"] = {"bound": "tensor", "achieved": fac_tflops, "peak": peak, "unit": "TFLOP/s", "frac": fac_tflops / peak,
                       "traffic": ncu_traffic("lu_gemm_kernel/config2/schur_level")[0],
                       "traffic_note": "GB (dram read+write) of the largest Schur-complement launch of one factorisation, " + str(ncu_traffic("lu_gemm_kernel/config2/schur_level")[1]),
                       "kernel": "lu_gemm_kernel (complex C -= A*B^T on DMMA m8n8k4 f64) inside the numeric LU",
                       "elimination": "symmetric (LDL^T-type) + rank-1 Woodbury for the flame term" if sym else "general LU",
                       "flops_per_factorisation": fac_flops,
                       "note": ("achieved = factorisation flops from the symbolic phase (8 real flops per complex multiply-add; half of the LU "
                                "count for the symmetric elimination; includes the 2 extra solves of the Woodbury set-up in the time) / "
                                "CUDA-event time of the numeric LU (all its kernels); peak = cuBLAS FP64 GEMM measured in this run "
                                f"(dgemm 8192^3 {peaks['dgemm']:.1f}, zgemm 4096^3 {peaks['zgemm']:.1f} TFLOP/s) -- MEASURED_PEAKS.json has no FP64 figure")}
    # triangular solves vs HBM: one refined solve (2 sweeps pairs: L then U^T panel each) of a random right-hand side
    hbm = 6548.5
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        hbm_src = "MEASURED_PEAKS.json"
    except Exception:
        hbm_src = "fallback"
    rng = np.random.default_rng(1)
    bvec = rng.standard_normal(dv.dim) + 1j * rng.standard_normal(dv.dim)
    sol_ms = []
    for _ in range(4):
        ctx.lu_solve(lu, bvec)
        sol_ms.append(ctx.last_ms("solve"))
    sol_med = float(np.median(sol_ms[1:]))
    sol_bytes = 2 * (16.0 * dv.lu_nnz + 2 * 16.0 * dv.dim)  # 16 (nnz(L)+nnz(U)) + 16 d nrhs 2 per solve; one refinement step = 2 solves
    out["solve"] = {"ms": sol_med, "nrhs": 1, "refinement_steps": 1,
                    "roofline": {"bound": "hbm", "achieved": sol_bytes / sol_med / 1e6, "peak": hbm, "unit": "GB/s", "frac": sol_bytes / sol_med / 1e6 / hbm,
                                 "traffic": None, "peak_source": hbm_src,
                                 "kernel": "lu_fwd_update / lu_bwd_update / lu_fwd_tri / lu_bwd_tri (level-scheduled, windowed)",
                                 "algorithmic_bytes": sol_bytes}}
    if not args.skip_extras and rank == 0:
        out["assembly"] = assembly_leg(W, ctx, args.assembly_cubes, hbm, hbm_src) if args.assembly_cubes else None
        # re-establish the tube mesh on the context for anything that follows
    # ------------------------------------------------------------------ Beyn leg (sharded over all ranks)
    if not args.skip_extras:
        out["beyn"] = beyn_leg(W, torch, dist, ctx, args, rank, world, dev, barrier, max_over_ranks)
    # ------------------------------------------------------------------ CPU baseline (rank 0, N = 1 only)
    if rank == 0 and world == 1 and not args.skip_extras:
        sample = tuple(int(x) for x in args.cpu_sample.split(","))
        st = oracle_step(sample)
        val, scaling = cpu_scaled(W, st, sample, tube)
        out["cpu_baseline"] = {"value": val, "unit": "eigenpairs/s", "cores": st["threads"], "kind": "port", "extrapolated": True, "scaling": scaling,
                               "cores_note": f"process CPU time / wall time of the sample (SuperLU is serial, OpenBLAS threads the dense kernels); host has {os.cpu_count()} cores",
                               "sample": (f"numpy/scipy (SuperLU+ARPACK) restatement of the reference path, NOT Julia/UMFPACK; one eigenpair on "
                                          f"the same tube at {sample[0]}x{sample[1]}x{sample[2]} cubes = {st['ntet']} tets, {st['dim']} P2 DOFs "
                                          f"({dv.dim / st['dim']:.0f}x fewer DOFs than the GPU workload), {st['nit']} Newton iterations, "
                                          f"{st['total']:.1f} s measured; value = that time scaled phase by phase to the full workload (assembly x "
                                          f"tets, factorisations x factorisation flops, rest x nnz(L+U)): an estimate, the full-size CPU run "
                                          f"takes hours; host has {os.cpu_count()} cores, SuperLU is serial")}
    # ------------------------------------------------------------------ diagnostic legs (rank 0, N = 1 only), each in a subprocess and inside a
    # wall-clock budget: a leg that would start later than DIAG_START_BY seconds into the run is skipped, and none may run past DIAG_END_BY
    if rank == 0 and world == 1 and not args.skip_extras:
        out["shape_sensitivity"] = diag_leg("bench_shape_sens.py", ["10", "10", "150", "5"], 300)
        out["lu_knobs"] = diag_leg("bench_lu_knobs.py", [*(str(x) for x in tube), "quad", "2"], 240)  # "combos" + "householder" (default vs paired)
    out["wall_s"] = time.perf_counter() - T_START
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def assembly_leg(W, ctx0, n, hbm, hbm_src):
    """M+K assembly of an n^3-cube P2 box (config 5 geometry at bench size), kernel time by CUDA events."""
    from wae_b200 import _lib
    ctx = _lib.Context(ctx0.device)
    mesh = W.kuhn_box((n, n, n), (0, 0, 0), (1, 1, 1), jitter=0.1, seed=7)
    tris, tets, dim = W.aggregate_elements(mesh, "quad")
    ctx.mesh_set(2, mesh.points.T, tets, tris, dim)
    pid, nnz = ctx.pattern_build(3, None)
    c = np.random.default_rng(7).uniform(300, 700, len(tets))
    im, ik = ctx.assemble_mk(pid, c)
    ms = []
    for _ in range(5):
        ctx.assemble_mk(pid, c, reuse=(im, ik))
        ms.append(ctx.last_ms("assemble"))
    med = float(np.median(ms))
    ntet, npts = len(tets), mesh.points.shape[1]
    alg = ntet * (4 * 10 + 8) + 24 * npts + 2 * nnz * 8
    res = {"value": ntet / med / 1e3, "unit": "Mtets/s", "tets": ntet, "dofs": dim, "nnz": int(nnz), "kernel_ms": med,
           "roofline": {"bound": "hbm", "achieved": alg / med / 1e6, "peak": hbm, "unit": "GB/s", "frac": alg / med / 1e6 / hbm,
                        "traffic": ncu_traffic(f"assemble_tet_pairs/p2_box_{n}")[0], "traffic_note": "GB per launch, " + str(ncu_traffic(f"assemble_tet_pairs/p2_box_{n}")[1]),
                        "peak_source": hbm_src, "algorithmic_bytes_per_tet": alg / ntet,
                        "kernel": "assemble_tet_pairs<10,3> (P2 M+K, owner-computes pair program, persistent, TMA-staged)"}}
    ctx.close()
    # opt-in kernel variants (WAE_ASM_VARIANT; prepared from the per-phase profile): timed and checked against the
    # default kernel in a SUBPROCESS, so that nothing they do can disturb this run; the headline above is always the default kernel
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:  # the other ranks are waiting for this one: no diagnostics in the scaling runs
        return res
    diag = diag_leg("bench_assembly_variants.py", [str(n), "quad", "7"], 240,
                    env={**os.environ, "CUDA_VISIBLE_DEVICES": os.environ.get("CUDA_VISIBLE_DEVICES", str(ctx0.device))})
    res["variants"] = diag.get("variants", diag)
    if "layouts" in diag:  # short patch-size x CTAs-per-SM sweep (the default layout stays 12288 slots, one CTA per SM)
        res["layouts"] = diag["layouts"]
    return res


def beyn_leg(W, torch, dist, ctx0, args, rank, world, dev, barrier, max_over_ranks):
    """Beyn contour integration, 4 x beyn_edge_nodes quadrature nodes sharded round-robin over the ranks,
    one NCCL all-reduce of the moments (strong scaling: the total work is fixed)."""
    from wae_b200 import _lib, nlevp
    nb = tuple(int(x) for x in args.beyn_box.split(","))
    ctx = _lib.Context(ctx0.device)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    mesh = W.kuhn_box(nb, (0, 0, -0.5), (0.1, 0.1, 0.5), jitter=0.1, seed=2024, name="beyn_tube")
    c = np.full(len(mesh.tetrahedra), 347.2)
    dscrp = {"Interior": ("interior", ()), "Outlet": ("admittance", ("Y", 1e15))}
    L = W.discretize(mesh, dscrp, c, order="lin", ctx=ctx)
    dv = L.device()
    dv.lu()
    G = [z * 2 * math.pi for z in (50 + 40j, 50 - 40j, 800 - 40j, 800 + 40j)]
    l, N = 8, args.beyn_edge_nodes
    # warm-up: one node per rank
    nlevp.compute_moment_matrices(L, G[:2], l=l, K=1, N=1)
    stats = {}
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rng = np.random.default_rng(0)
    Vp = rng.random((dv.dim, l)) + 1j * rng.random((dv.dim, l))  # beyn(...; random=true), seeded: same V on every rank
    A = nlevp.compute_moment_matrices(L, G, l=l, K=1, N=N, stats=stats, V=Vp)
    Om, P = nlevp.moments2eigs(A, G, rtol=1e-8, pos_test=True)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    res = {"value": len(Om) / (ms * 1e-3), "unit": "eigenpairs/s", "scaling": "strong", "n_gpus": world, "nodes": 4 * N, "l": l,
           "dofs": dv.dim, "tets": len(mesh.tetrahedra), "eigenvalues_found": len(Om), "ms": ms,
           "node_solves_per_s": 4 * N / (ms * 1e-3), "factor_ms_rank0": stats.get("factor_ms"), "solve_ms_rank0": stats.get("solve_ms"),
           "freq_hz": sorted(float(x) for x in (Om.real / 2 / math.pi))[:8],
           "collective": "one all_reduce (NCCL) of the 2K x l x d complex moment tensor"}
    ctx.close()
    return res


if __name__ == "__main__":
    main()
