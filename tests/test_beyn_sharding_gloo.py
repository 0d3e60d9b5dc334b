"""N>1 host logic on CPU: quadrature-node sharding + the single all-reduce of the Beyn moments
(world_size 2, gloo).  The per-rank moments come from the oracle here; on the GPU box they come from
wae_beyn_moments and the all-reduce runs over NCCL/NVLink."""
import math
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cases import load_raw_mesh, rijke_dscrp, speedofsound
    from oracle.helmholtz import discretize
    from oracle.mesh import Mesh
    from oracle.nlevp import beyn_moments, contour_nodes
    from wae_b200.nlevp import allreduce_moments, shard_nodes
    mesh = Mesh("m", scale=0.001, raw=load_raw_mesh("rijke_mm"))
    L = discretize(mesh, rijke_dscrp(0.0, 0.001), mesh.generate_field(speedofsound))
    G = [z * 2 * math.pi for z in (150 + 5j, 150 - 5j, 1000 - 5j, 1000 + 5j)]
    zs, _ = contour_nodes(G, 4)
    mine = shard_nodes(len(zs), rank, world)
    A = torch.from_numpy(beyn_moments(L, G, 3, 1, 4, nodes=mine))
    allreduce_moments(A)
    np.save(os.path.join(out_dir, f"A{rank}.npy"), A.numpy())
    # the product's own sharded routine (node shard -> per-rank moments -> one all-reduce), with the context replaced by the CPU test
    # double: every rank must end up with the serial moments
    import wae_b200 as W
    from host_standin import HostStandIn
    Lp = W.discretize(W.Mesh("m", scale=0.001, raw=load_raw_mesh("rijke_mm")), rijke_dscrp(0.0, 0.001), mesh.generate_field(speedofsound),
                      ctx=HostStandIn())
    np.save(os.path.join(out_dir, f"P{rank}.npy"), W.compute_moment_matrices(Lp, G, l=3, K=1, N=4))
    if rank == 0:
        np.save(os.path.join(out_dir, "Afull.npy"), beyn_moments(L, G, 3, 1, 4))
    dist.destroy_process_group()


def test_sharded_moments_equal_serial(tmp_path):
    from wae_b200.nlevp import shard_nodes
    parts = [set(shard_nodes(10, r, 3)) for r in range(3)]
    assert set().union(*parts) == set(range(10)) and sum(map(len, parts)) == 10
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    A0, A1, Af = (np.load(tmp_path / n) for n in ("A0.npy", "A1.npy", "Afull.npy"))
    assert np.array_equal(A0, A1)
    assert np.abs(A0 - Af).max() <= 1e-12 * np.abs(Af).max()
    P0, P1 = np.load(tmp_path / "P0.npy"), np.load(tmp_path / "P1.npy")
    assert np.array_equal(P0, P1) and P0.shape == Af.shape and np.abs(P0 - Af).max() <= 1e-10 * np.abs(Af).max()
