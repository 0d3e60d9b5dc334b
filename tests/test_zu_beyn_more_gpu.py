"""More of Beyn's contour integration on the GPU path (beyn.jl:34-110): the block-Hankel form K = 2, a random probing matrix
(random=true), the config-3 geometry (squircle cylinder, P2) at a size the oracle finishes in seconds with the reference's householder
polish, and the single-process multi-GPU entry (wae_beyn_moments_multi: needs two visible devices)."""
import math

import numpy as np
import pytest

from cases import load_raw_mesh, rijke_dscrp, speedofsound

pytestmark = pytest.mark.gpu


def _families(order="lin", n=0.0):
    import wae_b200 as W
    from oracle.helmholtz import discretize as odisc
    from oracle.mesh import Mesh as OMesh
    raw = load_raw_mesh("rijke_mm")
    mo = OMesh("m", scale=0.001, raw=raw)
    Lo = odisc(mo, rijke_dscrp(n, 0.001), mo.generate_field(speedofsound), order=order)
    mg = W.Mesh("m", scale=0.001, raw=raw)
    Lg = W.discretize(mg, rijke_dscrp(n, 0.001), mg.generate_field(speedofsound), order=order)
    return Lo, Lg


def test_beyn_block_hankel_and_random_probe():
    import wae_b200 as W
    from oracle.nlevp import beyn as obeyn
    from oracle.nlevp import beyn_moments as omoments
    Lo, Lg = _families()
    G = [z * 2 * math.pi for z in (150 + 5j, 150 - 5j, 1000 - 5j, 1000 + 5j)]
    # K = 2: four moments per node, block-Hankel matrices of size 2 d x 2 l (beyn.jl:77-83)
    Ao = omoments(Lo, G, 3, 2, 16)
    Ag = W.compute_moment_matrices(Lg, G, l=3, K=2, N=16)
    assert Ag.shape == Ao.shape == (Lo.size(), 3, 4)
    for p in range(4):
        assert np.abs(Ag[:, :, p] - Ao[:, :, p]).max() <= 1e-9 * np.abs(Ao[:, :, p]).max(), p
    Oo, _ = obeyn(Lo, G, l=3, K=2, N=16, tol=1e-8)
    Og, _ = W.beyn(Lg, G, l=3, K=2, N=16, tol=1e-8, output=False)
    assert len(Og) == len(Oo) == 2
    assert np.abs(np.sort_complex(Og) - np.sort_complex(Oo)).max() <= 1e-7 * np.abs(Oo).max()
    # random=true (beyn.jl:42-43): another probing matrix, the same eigenvalues
    Or, _ = W.beyn(Lg, G, l=4, K=1, N=16, tol=1e-8, output=False, random=True, seed=3)
    assert len(Or) == 2 and np.abs(np.sort_complex(Or) - np.sort_complex(Oo)).max() <= 1e-7 * np.abs(Oo).max()


def _cylinder(W, nc, R=0.05, Lz=1.0):
    mesh = W.kuhn_box(nc, (-R, -R, 0.0), (R, R, Lz), jitter=0.1, seed=2024, name="cylinder")
    x, y = mesh.points[0] / R, mesh.points[1] / R
    mesh.points[0], mesh.points[1] = R * x * np.sqrt(1 - 0.5 * y * y), R * y * np.sqrt(1 - 0.5 * x * x)  # square -> disc (bench.py, config 3)
    return mesh


def test_beyn_config3_geometry_reduced():
    """BASELINE.json configs[2] at 4 x 4 x 32 cubes (P2, 5 265 DOFs): Beyn on the GPU against the oracle, then householder polish."""
    import wae_b200 as W
    from oracle.helmholtz import discretize as odisc
    from oracle.mesh import Mesh as OMesh
    from oracle.nlevp import beyn_moments as omoments
    from oracle.nlevp import moments2eigs as om2e
    mesh = _cylinder(W, (4, 4, 32))
    c = np.full(len(mesh.tetrahedra), 347.2)
    dscrp = {"Interior": ("interior", ()), "Outlet": ("admittance", ("Y", 1e15))}
    Lg = W.discretize(mesh, dscrp, c, order="quad")
    raw = (mesh.points, [], [list(map(int, t)) for t in mesh.triangles], [list(map(int, t)) for t in mesh.tetrahedra],
           {k: {"dimension": v["dimension"], "simplices": list(map(int, v["simplices"]))} for k, v in mesh.domains.items()})
    mo = OMesh("m", raw=raw)
    Lo = odisc(mo, dscrp, np.full(len(mo.tetrahedra), 347.2), order="quad")
    assert Lo.size() == Lg.size() == 9 * 9 * 65
    G = [z * 2 * math.pi for z in (50 + 100j, 50 - 100j, 500 - 100j, 500 + 100j)]
    Ao = omoments(Lo, G, 8, 1, 6)
    Ag = W.compute_moment_matrices(Lg, G, l=8, K=1, N=6)
    assert np.abs(Ag - Ao).max() <= 1e-9 * np.abs(Ao).max()
    # the first l identity columns probe one corner of the mesh (beyn.jl:45-48), where the duct modes are small: the eigenvalues are
    # extracted from singular values ~1e-6 below the largest, which amplifies the 1e-10 agreement of the moments
    Oo = om2e(Ao, G, 8, 1, 1e-8, True)[0]
    Og, _ = W.moments2eigs(Ag, G, tol=1e-8, pos_test=True)
    assert len(Og) == len(Oo) == 3   # quarter-wave modes of the closed-open duct: 86.8, 260.4, 434 Hz
    assert np.abs(np.sort_complex(Og) - np.sort_complex(Oo)).max() <= 1e-5 * np.abs(Oo).max()
    for om in Og:
        sol, n, flag = W.householder(Lg, om, maxiter=8, tol=1e-9 * abs(om), output=False)
        assert flag >= 0 and abs(sol.params["ω"] - om) <= 1e-4 * abs(om)


def test_beyn_single_process_multi_gpu():
    """wae_beyn_moments_multi: one host process, one context per device, in-library ncclAllReduce -- equal to the single-device moments."""
    import torch

    import wae_b200 as W
    from wae_b200 import _lib
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible GPUs")
    raw = load_raw_mesh("rijke_mm")
    fams = []
    for dev in (0, 1):
        mg = W.Mesh("m", scale=0.001, raw=raw)
        fams.append(W.discretize(mg, rijke_dscrp(0.0, 0.001), mg.generate_field(speedofsound), order="lin", ctx=_lib.Context(dev)))
    G = [z * 2 * math.pi for z in (150 + 5j, 150 - 5j, 1000 - 5j, 1000 + 5j)]
    A1 = W.compute_moment_matrices(fams[0], G, l=5, K=1, N=8)
    A2 = W.compute_moment_matrices(fams[0], G, l=5, K=1, N=8, replicas=fams[1:])
    assert A1.shape == A2.shape and np.abs(A1 - A2).max() <= 1e-11 * np.abs(A1).max()
    O1, _ = W.beyn(fams[0], G, l=5, N=8, tol=1e-8, output=False)
    O2, _ = W.beyn(fams[0], G, l=5, N=8, tol=1e-8, output=False, replicas=fams[1:])
    assert len(O1) == len(O2) == 2 and np.abs(np.sort_complex(O1) - np.sort_complex(O2)).max() <= 1e-9 * np.abs(O1).max()
