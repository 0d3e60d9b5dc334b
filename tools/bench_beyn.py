"""Config 3 (BASELINE.json): Beyn contour integration, 128 quadrature nodes, ~2M-DOF synthetic cylinder, sharded over the GPUs.

    python tools/bench_beyn.py [--grid 32,32,236] [--order quad] [--edge-nodes 32] [--l 8]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_beyn.py ...

Mesh: squircle-mapped cylinder R = 0.05 m, L = 1.0 m, Kuhn grid (default P2 32 x 32 x 236 cubes -> 1 449 984 tets,
65*65*473 = 1 998 425 DOFs: the P1 100 x 100 x 196 variant of SURVEY 8d needs > 180 GB of factors), seed-2024 jitter,
passive flame (no Q term), open end (Y = 1e15) at z = L.  Contour: rectangle 50..850 Hz x +-100 Hz, 4 edges x 32 nodes.
Prints one JSON line (rank 0): wall/device time of the moment computation, eigenvalues found, per-rank factor/solve time.
"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", default="32,32,236")
    ap.add_argument("--order", default="quad")
    ap.add_argument("--edge-nodes", type=int, default=32)
    ap.add_argument("--l", type=int, default=8)
    ap.add_argument("--polish", type=int, default=1, help="householder-polish this many Beyn eigenvalues on rank 0 (not timed)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    import wae_b200 as W
    from wae_b200 import nlevp
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    t0 = time.perf_counter()
    nc = tuple(int(x) for x in args.grid.split(","))
    R, Lz = 0.05, 1.0
    mesh = W.kuhn_box(nc, (-R, -R, 0.0), (R, R, Lz), jitter=0.1, seed=2024, name="cylinder")
    x, y = mesh.points[0] / R, mesh.points[1] / R
    mesh.points[0], mesh.points[1] = R * x * np.sqrt(1 - 0.5 * y * y), R * y * np.sqrt(1 - 0.5 * x * x)  # square -> disc
    c = np.full(len(mesh.tetrahedra), 347.2)
    L = W.discretize(mesh, {"Interior": ("interior", ()), "Outlet": ("admittance", ("Y", 1e15))}, c, order=args.order)
    dv = L.device()
    ctx = dv.ctx
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    dv.lu()
    t_setup = time.perf_counter() - t0
    G = [z * 2 * math.pi for z in (50 + 100j, 50 - 100j, 850 - 100j, 850 + 100j)]
    stats = {}
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw = time.perf_counter()
    e0.record()
    rng = np.random.default_rng(0)  # beyn(...; random=true): the first l identity columns all sit in one corner of the mesh
    V = rng.random((dv.dim, args.l)) + 1j * rng.random((dv.dim, args.l))
    A = nlevp.compute_moment_matrices(L, G, l=args.l, K=1, N=args.edge_nodes, stats=stats, V=V)
    Om, P = nlevp.moments2eigs(A, G, rtol=1e-8, pos_test=True)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    fac = torch.tensor([stats.get("factor_ms", 0.0), stats.get("solve_ms", 0.0)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(fac, op=dist.ReduceOp.MAX)
    wall = time.perf_counter() - tw
    if rank == 0:
        sym = ctx.last_ms("factor_sym") > 0.5
        nodes = 4 * args.edge_nodes
        nper = len(nlevp.shard_nodes(nodes, 0, world))
        flops = dv.lu_flops * (0.5 if sym else 1.0)
        out = {"config": "Beyn, %d quadrature nodes, cylinder %s %s" % (nodes, args.grid, args.order), "n_gpus": world, "dofs": dv.dim,
               "tets": len(mesh.tetrahedra), "nnz": dv.nnz, "factor_nnz": dv.lu_nnz, "flops_per_factorisation": flops,
               "elimination": "symmetric" if sym else "general", "l": args.l, "ms": float(ms.item()), "wall_s": wall,
               "node_solves_per_s": nodes / (float(ms.item()) * 1e-3), "eigenvalues_found": int(len(Om)),
               "eigenpairs_per_s": len(Om) / (float(ms.item()) * 1e-3), "freq_hz": sorted(float(v) for v in Om.real / 2 / math.pi),
               "growth_hz": [float(v) for v in Om.imag / 2 / math.pi], "max_rank_factor_ms": float(fac[0].item()),
               "max_rank_solve_ms": float(fac[1].item()), "nodes_per_rank": nper,
               "factor_tflops_rank0": flops * nper / (stats.get("factor_ms", 1.0) * 1e-3) / 1e12, "setup_s_not_timed": t_setup}
        pol = []
        for om in sorted(Om, key=lambda z: z.real)[: args.polish]:
            sol, n, flag = W.householder(L, om, maxiter=8, tol=1e-9 * abs(om), output=False)
            pol.append({"beyn_hz": om.real / 2 / math.pi, "householder_hz": sol.params["ω"].real / 2 / math.pi, "iterations": n, "flag": flag,
                        "rel_diff": abs(sol.params["ω"] - om) / abs(om)})
        out["polish"] = pol
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
