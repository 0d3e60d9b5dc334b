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
With N > 1 every rank solves its own replica of the SAME problem (equal work per GPU): householder does not shard
("replicas only", weak scaling).  The path that does shard -- Beyn's quadrature nodes -- is timed on every N
as well and reported under "beyn" (strong scaling: 128 nodes in total, one NCCL all-reduce of the moments).

The `value` and `e2e` steps ALTERNATE inside one timed region (every step bracketed by its own CUDA events), so that neither
is measured on a warmer device than the other.

Objects on the JSON line: roofline (numeric LU = DMMA ZGEMM on the FP64 tensor pipe; `roofline.legs` holds the other kernels of the
path with their own rooflines -- assembly Mtets/s + HBM fraction, triangular solves, Beyn (BASELINE.json configs[2], sharded over the
ranks) -- and `roofline.phases_ms_per_step` the split of a step), cpu_baseline (the scipy/SuperLU oracle MEASURED on a bounded sample
of the same tube, the GPU path timed on that very sample beside it and the two eigenvalues compared: `cpu_baseline.same_problem`; the
phase-wise extrapolation to the full workload is a labelled side note there, never a headline), e2e, clocks.

`--impl reference`: the CPU oracle alone, `value` = its MEASURED rate on the bounded sample (it never loads libwae_b200.so).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time


def _pin_rank_to_cores():
    """N replicas share one host: give every rank its own core set and cap the BLAS / OpenMP pools to it BEFORE numpy, scipy and torch
    start their thread pools (round 1: 8 unpinned ranks with 16-thread pools each lost 35 % at N = 8 without a single collective)."""
    try:
        local, world = int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
        cores = sorted(os.sched_getaffinity(0))
        if world > 1 and len(cores) >= world:
            per = len(cores) // world
            mine = cores[local * per:(local + 1) * per]
            os.sched_setaffinity(0, mine)
            for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
                os.environ[k] = str(max(1, len(mine)))
            return len(mine)
        return len(cores)
    except (AttributeError, OSError, ValueError):
        return None


HOST_CORES_OF_RANK = _pin_rank_to_cores()

import numpy as np  # noqa: E402

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


def workload_config(tube, sample):
    """The `config` object: identical in both arms (the driver compares them)."""
    return {"workload": (f"BASELINE.json configs[1]: Rijke tube {tube[0]}x{tube[1]}x{tube[2]} Kuhn cubes, P2, interior + outlet admittance (Y=1e15) + "
                         "n-tau flame (n=1, tau=1 ms); step = numeric re-assembly of all operators + householder(order 1) from 340*2*pi to |dw| <= 1e-9 |w| "
                         "= one eigenpair"),
            "tube_cubes": list(tube), "elements": "P2 tetrahedra", "arithmetic": "complex fp64",
            "cpu_sample": (f"the CPU arm times the same step on a bounded sample of this workload: the same tube at {sample[0]}x{sample[1]}x{sample[2]} cubes "
                           "(the full-size CPU step takes hours); the GPU arm times that sample too (cpu_baseline.same_problem)"),
            "l2": "working set (LU factors, 21 GB) is far larger than the 126 MB L2; no flush needed",
            "parallelism": "replicas only (householder does not shard); Beyn's quadrature nodes are sharded under roofline.legs.beyn"}


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
            "dim": L.size(), "ntet": len(m.tetrahedra), "nit": n, "threads": max(1, round((time.process_time() - c0) / (t2 - t0))),
            "omega": complex(sol.params["ω"]), "flag": flag}


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
    return {"what": ("SIDE NOTE, an estimate and not a measurement: the measured sample scaled phase by phase to the full workload (assembly x "
                     "tetrahedra, factorisations x factorisation flops, rest x nnz(L+U), both counts from one nested-dissection analysis of the two "
                     "patterns; SuperLU's own fill grows faster, so this favours the CPU)"),
            "growth": {"tets": tf / ts, "factor_flops": ff / fs, "factor_nnz": nf / ns}, "estimated_full_size_s_per_eigenpair": est}


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
        sm, mx, pw, reasons = [], [], [], set()
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            try:
                pw.append(float(parts[2]))
            except ValueError:
                pass
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm),
                   "sm_mhz_min": float(min(sm)), "power_w_median": float(np.median(pw)) if pw else None, "power_w_max": float(max(pw)) if pw else None}
        os.unlink(self.f.name)
        return out


def ncu_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of a kernel, from the committed `ncu --set full` summaries
    (profiles/r02_traffic.json, else r01: kernel/workload -> bytes, with the capture it was read from); None if not captured."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", name)))[key]
            return t["dram_bytes_per_launch"] / 1e9, t["source"]
        except Exception:
            continue
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
    ap.add_argument("--beyn-grid", default="32,32,236", help="cubes of the squircle cylinder of the sharded Beyn leg (config 3: 32,32,236, P2)")
    ap.add_argument("--beyn-order", default="quad")
    ap.add_argument("--beyn-edge-nodes", type=int, default=32, help="Gauss-Legendre nodes per polygon edge (4 edges -> 128 nodes)")
    ap.add_argument("--assembly-cubes", type=int, default=64, help="n for the n^3-cube P2 assembly-only leg (0 = skip)")
    ap.add_argument("--cpu-sample", default="4,4,60", help="cubes of the bounded sample both CPU legs (and the GPU same-problem leg) run")
    ap.add_argument("--skip-extras", action="store_true")
    ap.add_argument("--host-profile", action="store_true", help="one extra, untimed step under cProfile on rank 0: the host-side hot spots go to details.host_profile")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    tube = tuple(int(x) for x in args.tube.split(","))
    sample = tuple(int(x) for x in args.cpu_sample.split(","))
    config = workload_config(tube, sample)
    metric = "NLEVP eigenpairs/s (householder, config 2)"

    def cpu_sample_text(st):
        return (f"reference cannot run here (pure Julia, no julia binary): CPU restatement (numpy + scipy SuperLU/ARPACK, NOT Julia/UMFPACK) MEASURED "
                f"on a bounded sample of the workload: tube {sample[0]}x{sample[1]}x{sample[2]} cubes, {st['ntet']} tets, {st['dim']} P2 DOFs, assembly + "
                f"householder to convergence = one eigenpair per step ({st['total']:.2f} s, {st['n_fact']} factorisations, {st['nit']} Newton iterations); "
                f"value = measured eigenpairs/s ON THAT SAMPLE; host has {os.cpu_count()} cores, SuperLU is serial, OpenBLAS threads the dense kernels")

    # ------------------------------------------------------------------ reference arm: the CPU oracle, measured (never loads libwae_b200.so)
    if args.impl == "reference":
        if rank != 0:
            return
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
        val = args.steps / dt
        print(json.dumps({"impl": "reference", "metric": metric, "value": val, "unit": "eigenpairs/s", "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "c128", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": val, "unit": "eigenpairs/s", "cores": acc["threads"], "kind": "port", "sample": cpu_sample_text(acc),
                                           "extrapolated": False, "sample_dofs": acc["dim"], "sample_tets": acc["ntet"],
                                           "phases_s": {k: acc[k] for k in ("asm", "factor", "other")}, "omega": [acc["omega"].real, acc["omega"].imag],
                                           "cores_note": f"process CPU time / wall time of the sample; host has {os.cpu_count()} cores"},
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
    tau = 0.001  # the SAME replica on every rank: equal work per GPU (a rank-dependent tau changes the Newton iteration count, 6 ... 9 at N = 8)
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
        w0 = time.perf_counter()
        if e2e:
            ctx.mesh_update_points(pts_pinned.numpy())
        ms = disc.reassemble(c_pinned.numpy())
        w1 = time.perf_counter()
        stats["assemble_ms"] = stats.get("assemble_ms", 0.0) + ms
        L.params["n"], L.params["τ"] = 1.0 + 0j, complex(tau)
        sol, n, flag = W.householder(L, Z0, maxiter=15, tol=tol, output=False, stats=stats)
        stats["reassemble_wall_s"] = stats.get("reassemble_wall_s", 0.0) + w1 - w0
        stats["householder_wall_s"] = stats.get("householder_wall_s", 0.0) + time.perf_counter() - w1
        stats["iterations"] = stats.get("iterations", 0) + n
        if flag < 0:
            raise RuntimeError(f"householder failed with flag {flag}")
        return sol

    for _ in range(args.warmup):
        sol = step(False, {})
    omega = sol.params["ω"]

    host_profile = None
    if args.host_profile:
        import cProfile
        import io
        import pstats
        barrier()
        pr = cProfile.Profile()
        pr.enable()
        step(False, {})
        pr.disable()
        barrier()
        if rank == 0:
            buf = io.StringIO()
            pstats.Stats(pr, stream=buf).sort_stats("tottime").print_stats(18)
            host_profile = [l.strip() for l in buf.getvalue().splitlines() if l.strip()][-20:]
    # value and e2e steps alternate inside ONE timed region; every step has its own pair of CUDA events
    stats, stats_e2e = {}, {}
    ms_plain = ms_e2e = 0.0
    launches = 0
    barrier()
    clk = Clocks(local)  # every rank samples its own GPU (rank 0's record is the line's "clocks", all of them go to details.per_rank)
    t_region = time.perf_counter()
    for _ in range(args.steps):
        for e2e, st in ((False, stats), (True, stats_e2e)):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = ctx.launch_count()
            e0.record()
            step(e2e, st)
            e1.record()
            e1.synchronize()
            if e2e:
                ms_e2e += e0.elapsed_time(e1)
            else:
                ms_plain += e0.elapsed_time(e1)
                launches += ctx.launch_count() - l0
    barrier()
    t_region = time.perf_counter() - t_region
    clocks = clk.stop()
    per_rank = None
    if world > 1:  # every rank's own step time and phase split (rank 0 prints them: which rank, and which phase, is the slow one)
        mine = {"rank": rank, "ms_per_step": ms_plain / args.steps, "numeric_lu": stats["factor_ms"] / args.steps,
                "eigs_wall": 1e3 * stats["eigs_wall_s"] / args.steps, "perturb_wall": 1e3 * stats.get("perturb_wall_s", 0.0) / args.steps,
                "reassemble_wall": 1e3 * stats["reassemble_wall_s"] / args.steps, "householder_wall": 1e3 * stats["householder_wall_s"] / args.steps,
                "iterations": stats["iterations"] / args.steps, "factorizations": stats["factorizations"] / args.steps,
                "cores": HOST_CORES_OF_RANK, "lu_tflops": (dv.lu_flops * (0.5 if ctx.last_ms("factor_sym") > 0.5 else 1.0)) * stats["factorizations"] / (stats["factor_ms"] * 1e-3) / 1e12,
                "clocks": clocks}
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)
    ms, ms_e2e = max_over_ranks(ms_plain), max_over_ranks(ms_e2e)
    value = world * args.steps / (ms * 1e-3)
    value_e2e = world * args.steps / (ms_e2e * 1e-3)
    nfac = stats["factorizations"]
    sym = ctx.last_ms("factor_sym") > 0.5   # complex-symmetric elimination (LDL^T-type): half of the LU multiply-adds
    fac_flops = dv.lu_flops * (0.5 if sym else 1.0)
    fac_tflops = fac_flops * nfac / (stats["factor_ms"] * 1e-3) / 1e12
    iters = stats["iterations"]
    h2d = 3 * npts * 8 + ntet * 8 + (iters / args.steps) * 2 * 16 * dv.dim  # points + c + Krylov start vectors per iteration
    d2h = (iters / args.steps) * 2 * 16 * dv.dim                              # eigenvector pairs back to the host

    def phases(st):
        return {"assemble": st["assemble_ms"] / args.steps, "numeric_lu": st["factor_ms"] / args.steps,
                "combine_factor_wall": 1e3 * st["combine_factor_wall_s"] / args.steps, "eigs_wall": 1e3 * st["eigs_wall_s"] / args.steps,
                "perturb_wall": 1e3 * st.get("perturb_wall_s", 0.0) / args.steps, "reassemble_wall": 1e3 * st["reassemble_wall_s"] / args.steps,
                "householder_wall": 1e3 * st["householder_wall_s"] / args.steps, "iterations": st["iterations"] / args.steps,
                "factorizations": st["factorizations"] / args.steps, "solves": st["solves"] / args.steps}

    out = {"metric": metric, "value": value, "unit": "eigenpairs/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "c128", "data": "synthetic", "config": config,
           "details": {"tets": ntet, "dofs": dv.dim, "nnz": dv.nnz, "factor_nnz": dv.lu_nnz, "factor_flops": dv.lu_flops, "setup_s_not_timed": t_setup,
                       "omega": [omega.real, omega.imag], "timed_region_wall_s": t_region,
                       "timed_region": "value and e2e steps alternate; ms_per_step = sum of the value steps' CUDA-event times / steps (max over ranks)",
                       "host_cores_of_this_rank": HOST_CORES_OF_RANK, "per_rank": per_rank, "host_profile": host_profile},
           "e2e": {"value": value_e2e, "unit": "eigenpairs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                   "ms_per_step": ms_e2e / args.steps},
           "gpu_launches": int(launches), "clocks": clocks}

    # ------------------------------------------------------------------ roofline of the dominant kernel (+ the other kernels under "legs")
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
                                f"(dgemm 8192^3 {peaks['dgemm']:.1f}, zgemm 4096^3 {peaks['zgemm']:.1f} TFLOP/s) -- MEASURED_PEAKS.json has no FP64 figure"),
                       "phases_ms_per_step": phases(stats), "phases_ms_per_step_e2e": phases(stats_e2e), "legs": {}}
    legs = out["roofline"]["legs"]
    # triangular solves vs HBM: one refined solve (2 sweeps pairs: L then U^T panel each) of a random right-hand side
    hbm = 6548.5
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        hbm_src = "MEASURED_PEAKS.json"
    except Exception:
        hbm_src = "fallback"
    rng = np.random.default_rng(1)
    for nrhs in (1, 8):
        bvec = rng.standard_normal((dv.dim, nrhs)) + 1j * rng.standard_normal((dv.dim, nrhs))
        sol_ms = []
        for _ in range(4):
            ctx.lu_solve(lu, bvec if nrhs > 1 else bvec[:, 0])
            sol_ms.append(ctx.last_ms("solve"))
        sol_med = float(np.median(sol_ms[1:]))
        sol_bytes = 2 * (16.0 * dv.lu_nnz + 2 * 16.0 * dv.dim * nrhs)  # 16 (nnz(L)+nnz(U)) + 16 d nrhs 2 per solve; one refinement step = 2 solves
        legs["solve_nrhs%d" % nrhs] = {"ms": sol_med, "nrhs": nrhs, "refinement_steps": 1, "bound": "hbm", "achieved": sol_bytes / sol_med / 1e6, "peak": hbm,
                                       "unit": "GB/s", "frac": sol_bytes / sol_med / 1e6 / hbm, "traffic": None, "peak_source": hbm_src,
                                       "kernel": "lu_fwd_update / lu_bwd_update / lu_fwd_tri / lu_bwd_tri (level-scheduled, windowed)",
                                       "algorithmic_bytes": sol_bytes}
    # ------------------------------------------------------------------ the same problem on both sides (rank 0, N = 1): the CPU oracle's bounded
    # sample, measured, and the GPU path on that very mesh; the two eigenvalues are the tube-geometry parity check
    if rank == 0 and world == 1 and not args.skip_extras:
        st = oracle_step(sample)
        same = same_problem_leg(W, ctx, sample, st)
        out["cpu_baseline"] = {"value": 1.0 / st["total"], "unit": "eigenpairs/s", "cores": st["threads"], "kind": "port", "extrapolated": False,
                               "sample": cpu_sample_text(st), "sample_dofs": st["dim"], "sample_tets": st["ntet"],
                               "phases_s": {k: st[k] for k in ("asm", "factor", "other")},
                               "cores_note": f"process CPU time / wall time of the sample (SuperLU is serial, OpenBLAS threads the dense kernels); host has {os.cpu_count()} cores",
                               "same_problem": same}
        try:
            out["cpu_baseline"]["full_size_estimate"] = cpu_scaled(W, st, sample, tube)
        except Exception as e:  # noqa: BLE001 -- a side note only
            out["cpu_baseline"]["full_size_estimate"] = {"error": repr(e)[:200]}
    # the 23 GB of factors of the tube are not needed any more: the Beyn leg's cylinder needs 78 GB per GPU
    L.release()
    if not args.skip_extras and rank == 0 and args.assembly_cubes:
        legs["assembly"] = assembly_leg(W, ctx, args.assembly_cubes, hbm, hbm_src)
    # ------------------------------------------------------------------ Beyn leg (BASELINE.json configs[2], sharded over all ranks)
    if not args.skip_extras:
        legs["beyn"] = beyn_leg(W, torch, dist, ctx, args, rank, world, dev, barrier, max_over_ranks)
    # ------------------------------------------------------------------ diagnostic legs (rank 0, N = 1 only), each in a subprocess and inside a
    # wall-clock budget: a leg that would start later than DIAG_START_BY seconds into the run is skipped, and none may run past DIAG_END_BY
    if rank == 0 and world == 1 and not args.skip_extras:
        out["shape_sensitivity"] = diag_leg("bench_shape_sens.py", ["10", "10", "150", "5"], 300)
    out["wall_s"] = time.perf_counter() - T_START
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def same_problem_leg(W, ctx0, sample, st_cpu):
    """The GPU path on the CPU oracle's own sample (same mesh, same descriptor, same start value and tolerance): eigenpairs/s and the
    relative difference of the two eigenvalues."""
    import torch

    from wae_b200 import _lib
    ctx = _lib.Context(ctx0.device)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    mesh, c, dscrp = tube_case(W, sample)
    L = W.discretize(mesh, dscrp, c, order="quad", ctx=ctx)
    disc = L.discretization
    L.device().lu()

    def one():
        disc.reassemble(c)
        sol, n, flag = W.householder(L, Z0, maxiter=15, tol=1e-9 * Z0, output=False)
        return sol, n, flag
    for _ in range(2):
        one()
    reps = 5
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        sol, n, flag = one()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    om_g, om_c = complex(sol.params["ω"]), st_cpu["omega"]
    res = {"mesh": f"tube {sample[0]}x{sample[1]}x{sample[2]} cubes, {len(mesh.tetrahedra)} tets, {L.size()} P2 DOFs (identical on both sides)",
           "gpu_eigenpairs_per_s": 1e3 / ms, "gpu_ms_per_eigenpair": ms, "gpu_iterations": n, "gpu_flag": flag,
           "cpu_eigenpairs_per_s": 1.0 / st_cpu["total"], "cpu_s_per_eigenpair": st_cpu["total"], "cpu_iterations": st_cpu["nit"],
           "measured_ratio_gpu_over_cpu": (1e3 / ms) * st_cpu["total"],
           "omega_gpu": [om_g.real, om_g.imag], "omega_cpu": [om_c.real, om_c.imag], "omega_rel_diff": abs(om_g - om_c) / abs(om_c),
           "note": "both sides measured on the same problem in the same run; a 10^4-DOF problem is launch-latency-bound on the GPU"}
    ctx.close()
    return res


def assembly_leg(W, ctx0, n, hbm, hbm_src):
    """M+K assembly of an n^3-cube P2 box (config 5 geometry at bench size), kernel time by CUDA events: the star kernel (generation 3,
    the default) and, beside it, the pair-program kernel of round 1 (generation 2)."""
    from wae_b200 import _lib
    ctx = _lib.Context(ctx0.device)
    mesh = W.kuhn_box((n, n, n), (0, 0, 0), (1, 1, 1), jitter=0.1, seed=7)
    tris, tets, dim = W.aggregate_elements(mesh, "quad")
    ctx.mesh_set(2, mesh.points.T, tets, tris, dim)
    pid, nnz = ctx.pattern_build(3, None)
    c = np.random.default_rng(7).uniform(300, 700, len(tets))
    ntet, npts = len(tets), mesh.points.shape[1]
    alg = ntet * (4 * 10 + 8) + 24 * npts + 2 * nnz * 8

    def timed(gen):
        old = os.environ.pop("WAE_ASM_GEN", None)
        if gen is not None:
            os.environ["WAE_ASM_GEN"] = gen
        try:
            im, ik = ctx.assemble_mk(pid, c)
            ms = []
            for _ in range(7):
                ctx.assemble_mk(pid, c, reuse=(im, ik))
                ms.append(ctx.last_ms("assemble"))
            ctx.mat_free(im); ctx.mat_free(ik)
        finally:
            os.environ.pop("WAE_ASM_GEN", None)
            if old is not None:
                os.environ["WAE_ASM_GEN"] = old
        return float(np.median(ms))
    med = timed("3")
    prog = ctx.last_ms("star_program_bytes")
    layout = {k: ctx.last_ms("star_" + k) for k in ("patches", "staged", "simplices", "sources", "smem", "threads", "ctas_per_sm")}
    med2 = timed("2")
    traffic, tsrc = ncu_traffic(f"assemble_tet_stars/p2_box_{n}")
    res = {"value": ntet / med / 1e3, "unit": "Mtets/s", "tets": ntet, "dofs": dim, "nnz": int(nnz), "kernel_ms": med,
           "bound": "hbm", "achieved": alg / med / 1e6, "peak": hbm, "frac": alg / med / 1e6 / hbm, "achieved_unit": "GB/s",
           "traffic": traffic, "traffic_note": "GB per launch (dram read + write), " + str(tsrc), "peak_source": hbm_src,
           "algorithmic_bytes_per_tet": alg / ntet, "program_bytes_per_tet": prog / ntet if prog > 0 else None, "layout": layout,
           "kernel": "assemble_tet_stars<10,3> (P2 M+K, star program: sub-simplex stars summed in registers, persistent, TMA-staged program)",
           "generation_2": {"kernel": "assemble_tet_pairs<10,3> (round 1, WAE_ASM_GEN=2)", "kernel_ms": med2, "value": ntet / med2 / 1e3,
                            "frac": alg / med2 / 1e6 / hbm}}
    ctx.close()
    return res


def beyn_leg(W, torch, dist, ctx0, args, rank, world, dev, barrier, max_over_ranks):
    """BASELINE.json configs[2]: Beyn contour integration, 4 x beyn_edge_nodes quadrature nodes on the ~2 M-DOF squircle cylinder (P2 32 x 32 x
    236 cubes, 1 998 425 DOFs), l = 8, nodes sharded round-robin over the ranks, one NCCL all-reduce of the moments (strong scaling: the
    total work is fixed; reported at every N, N = 1 included)."""
    from wae_b200 import _lib, nlevp
    nb = tuple(int(x) for x in args.beyn_grid.split(","))
    ctx = _lib.Context(ctx0.device)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    t0 = time.perf_counter()
    R, Lz = 0.05, 1.0
    mesh = W.kuhn_box(nb, (-R, -R, 0.0), (R, R, Lz), jitter=0.1, seed=2024, name="cylinder")
    x, y = mesh.points[0] / R, mesh.points[1] / R
    mesh.points[0], mesh.points[1] = R * x * np.sqrt(1 - 0.5 * y * y), R * y * np.sqrt(1 - 0.5 * x * x)  # square -> disc
    c = np.full(len(mesh.tetrahedra), 347.2)
    dscrp = {"Interior": ("interior", ()), "Outlet": ("admittance", ("Y", 1e15))}
    L = W.discretize(mesh, dscrp, c, order=args.beyn_order, ctx=ctx)
    dv = L.device()
    dv.lu()
    t_setup = time.perf_counter() - t0
    G = [z * 2 * math.pi for z in (50 + 100j, 50 - 100j, 850 - 100j, 850 + 100j)]
    l, N = 8, args.beyn_edge_nodes
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
    fac = max_over_ranks(stats.get("factor_ms", 0.0))
    sol = max_over_ranks(stats.get("solve_ms", 0.0))
    sym = ctx.last_ms("factor_sym") > 0.5
    nper = len(nlevp.shard_nodes(4 * N, 0, world))
    flops = dv.lu_flops * (0.5 if sym else 1.0)
    res = {"config": f"BASELINE.json configs[2]: Beyn, {4 * N} quadrature nodes, l = {l}, K = 1, squircle cylinder {args.beyn_grid} cubes, {args.beyn_order}",
           "value": len(Om) / (ms * 1e-3), "unit": "eigenpairs/s", "scaling": "strong", "n_gpus": world, "nodes": 4 * N, "l": l,
           "dofs": dv.dim, "tets": len(mesh.tetrahedra), "factor_nnz": dv.lu_nnz, "eigenvalues_found": len(Om), "ms": ms, "seconds": ms * 1e-3,
           "node_solves_per_s": 4 * N / (ms * 1e-3), "max_rank_factor_ms": fac, "max_rank_solve_ms": sol, "nodes_per_rank": nper,
           "factor_tflops_rank0": flops * nper / (max(stats.get("factor_ms", 1.0), 1e-9) * 1e-3) / 1e12, "setup_s_not_timed": t_setup,
           "freq_hz": sorted(float(x) for x in (Om.real / 2 / math.pi))[:8],
           "collective": "one all_reduce (NCCL) of the 2K x l x d complex moment tensor",
           "speedup_note": "strong scaling: seconds at N GPUs against seconds at N = 1 of the same leg (SCALE_r02.json holds every N)"}
    ctx.close()
    return res


if __name__ == "__main__":
    main()
