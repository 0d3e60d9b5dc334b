"""Parity at BASELINE.json's full size (config 2: 720 000 tets, 1 010 281 P2 DOFs) through size-independent
properties -- the oracle cannot run there (SuperLU would need hours)."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big():
    import bench
    import wae_b200 as W
    mesh, c, dscrp = bench.tube_case(W, (20, 20, 300))
    L = W.discretize(mesh, dscrp, c, order="quad")
    return W, mesh, c, L


def test_assembly_properties_at_full_size(big):
    W, mesh, c, L = big
    dev = L.device()
    ctx = dev.ctx
    d = L.size()
    assert d == 1010281 and len(mesh.tetrahedra) == 720000
    ones = np.ones(d, dtype=complex)
    terms = {t.operator: i for i, t in enumerate(L.terms)}

    def apply(op, x):
        sc = [None] * len(L.terms)
        sc[terms[op]] = 1.0
        dev.combine(dev.flat(sc), 2)
        return ctx.spmm(dev.fid, 2, x)
    # partition of unity: 1^T M 1 = volume, K 1 = 0 (constants are in the kernel of the stiffness operator)
    vol = 0.05 * 0.05 * 0.5
    assert abs(np.vdot(ones, apply("M", ones)).real - vol) < 1e-12 * vol
    k1 = apply("K", ones)
    kd = apply("K", np.random.default_rng(0).standard_normal(d) + 0j)
    assert np.abs(k1).max() < 1e-9 * np.abs(kd).max()
    # aux term is exactly -M (Helmholtz.jl:572); symmetry of M and K
    assert np.abs(apply("__aux__", ones) + apply("M", ones)).max() == 0.0
    x = np.random.default_rng(1).standard_normal(d) + 0j
    y = np.random.default_rng(2).standard_normal(d) + 0j
    for op in ("M", "K"):
        assert abs(np.vdot(y, apply(op, x)) - np.vdot(apply(op, y), x)) < 1e-10 * abs(np.vdot(y, apply(op, x)))
    # scaling law of the stiffness term: K(2c) = 4 K(c) (re-assembly in place)
    kx = apply("K", x)
    L.discretization.reassemble(2.0 * c)
    assert np.abs(apply("K", x) - 4 * kx).max() <= 1e-13 * np.abs(kx).max()
    L.discretization.reassemble(c)
    assert np.array_equal(apply("K", x), kx)  # deterministic assembly: bit-identical after re-assembly


def test_solve_and_eigenpair_at_full_size(big):
    W, mesh, c, L = big
    dev = L.device()
    ctx = dev.ctx
    d = L.size()
    z = 340 * 2 * math.pi
    rng = np.random.default_rng(5)
    b = rng.standard_normal(d) + 1j * rng.standard_normal(d)
    op = L(z)
    x = op.solve(b)
    r = op.matvec(x) - b
    assert np.abs(r).max() < 1e-10 * np.abs(b).max()
    # solve -> multiply round trip for the transposed and adjoint systems
    lid = dev.lu()
    for trans in (1, 2):
        xt = ctx.lu_solve(lid, b, trans=trans)
        assert np.abs(op.matvec(xt, trans=trans) - b).max() < 1e-10 * np.abs(b).max()
    # converged eigenpair: residual of L(omega) v and of the adjoint pair
    sol, n, flag = W.householder(L, z, maxiter=15, tol=1e-9 * z, output=False)
    assert flag in (0, 1) and n <= 8
    om = sol.params["ω"]
    res = L(om).matvec(sol.v)
    ref = np.abs(L(om).to_scipy()) @ np.abs(sol.v)
    m = ref > 1e-3 * ref.max()  # ignore the penalised outlet rows (values ~1e-17 times 1e18 entries)
    assert (np.abs(res)[m] / ref[m]).max() < 1e-7
    # the same mode as the 1/8-size tube (mesh convergence of the first thermoacoustic mode): sanity of the physics
    assert abs(om - (1159.6 + 393.2j)) < 1.0
