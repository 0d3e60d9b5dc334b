"""GPU parity: CUDA assembly / combine / SpMM through the C ABI vs the CPU oracle (same inputs)."""
import math

import numpy as np
import pytest
import scipy.sparse as sp

from cases import load_raw_mesh, rijke_dscrp, speedofsound

pytestmark = pytest.mark.gpu

VAL_TOL = 1e-12  # relative to the largest entry of the operator (fp64, different summation order)


def _oracle_family(raw, order, n=0.01, tau=0.001, scale=0.001, dscrp=None, cfun=speedofsound):
    from oracle.helmholtz import discretize
    from oracle.mesh import Mesh
    mesh = Mesh("m", scale=scale, raw=raw)
    c = mesh.generate_field(cfun)
    return discretize(mesh, dscrp or rijke_dscrp(n, tau), c, order=order)


def _gpu_family(raw, order, n=0.01, tau=0.001, scale=0.001, dscrp=None, cfun=speedofsound):
    import wae_b200 as W
    mesh = W.Mesh("m", scale=scale, raw=raw)
    c = mesh.generate_field(cfun)
    return W.discretize(mesh, dscrp or rijke_dscrp(n, tau), c, order=order)


def _check_term(tg, to):
    A = sp.csc_matrix(to.coeff)
    A.sort_indices()
    colptr, rowval, nzval = tg.coeff.csc()
    # bit-exact pattern: same (colptr,rowval) as sparse() of the oracle's triplets
    assert np.array_equal(colptr, A.indptr), tg.operator
    assert np.array_equal(rowval, A.indices), tg.operator
    scale = np.abs(A.data).max()
    assert np.abs(nzval - A.data).max() <= VAL_TOL * scale, (tg.operator, np.abs(nzval - A.data).max() / scale)


@pytest.mark.parametrize("order", ["lin", "quad"])
def test_rijke_terms_match_oracle(order):
    raw = load_raw_mesh("rijke_mm")
    Lo = _oracle_family(raw, order)
    Lg = _gpu_family(raw, order)
    assert [t.operator for t in Lg.terms] == [t.operator for t in Lo.terms]
    assert Lg.size() == Lo.size() == (1006 if order == "lin" else Lo.size())
    for tg, to in zip(Lg.terms, Lo.terms):
        _check_term(tg, to)
    assert set(Lg.params) == set(Lo.params)


@pytest.mark.parametrize("order", ["lin", "quad"])
def test_rijke_combine_and_spmm(order):
    raw = load_raw_mesh("rijke_mm")
    Lo = _oracle_family(raw, order, n=1.0)
    Lg = _gpu_family(raw, order, n=1.0)
    rng = np.random.default_rng(3)
    for z in (340 * 2 * math.pi, 1075.3 + 372.1j):
        Ao = sp.csc_matrix(Lo(z))
        Ag = Lg(z).to_scipy()
        d = (Ag - Ao)
        assert abs(d).max() <= 1e-12 * abs(Ao).max()
        x = rng.standard_normal((Lo.size(), 3)) + 1j * rng.standard_normal((Lo.size(), 3))
        for trans, ref in ((0, Ao @ x), (1, Ao.T @ x), (2, Ao.conj().T @ x)):
            y = Lg(z).matvec(x, trans=trans)
            assert np.abs(y - ref).max() <= 1e-12 * np.abs(ref).max()
    # derivative evaluation L(z,1) (mode :all) and the side effect on params
    Ao1 = sp.csc_matrix(Lo(500.0 + 2j, 1))
    Ag1 = Lg(500.0 + 2j, 1).to_scipy()
    assert abs(Ag1 - Ao1).max() <= 1e-12 * abs(Ao1).max()
    assert Lg.params["ω"] == Lo.params["ω"] == 500.0 + 2j


def test_kuhn_box_quad_matches_oracle_and_atomic_path(monkeypatch):
    import wae_b200 as W
    mesh = W.kuhn_box((5, 4, 7), (0, 0, -0.1), (0.05, 0.04, 0.1), jitter=0.15, seed=5, flame_layer=(3, 4))
    raw = (mesh.points, [], [list(t) for t in mesh.triangles], [list(t) for t in mesh.tetrahedra],
           {k: {"dimension": v["dimension"], "simplices": list(map(int, v["simplices"]))} for k, v in mesh.domains.items()})
    cf = lambda x, y, z: 347.0 if z < 0 else 694.0
    dscrp = {"Interior": ("interior", ()), "Outlet": ("admittance", ("Y", 1e15)),
             "Flame": ("flame", (1.4, 1.225, 1000.0, [0.02, 0.02, -0.03], [0, 0, 1.0], "n", "τ", 1.0, 0.001))}
    Lo = _oracle_family(raw, "quad", scale=1.0, dscrp=dscrp, cfun=cf)
    Lg = _gpu_family(raw, "quad", scale=1.0, dscrp=dscrp, cfun=cf)
    for tg, to in zip(Lg.terms, Lo.terms):
        _check_term(tg, to)
    # the atomic-scatter generation must agree with the gather generation
    monkeypatch.setenv("WAE_FORCE_ATOMIC", "1")
    La = _gpu_family(raw, "quad", scale=1.0, dscrp=dscrp, cfun=cf)
    for ta, tg in zip(La.terms, Lg.terms):
        va, vg = ta.coeff.csc()[2], tg.coeff.csc()[2]
        assert np.abs(va - vg).max() <= 1e-13 * np.abs(vg).max()


def test_linear_c_stiffness_and_boundary():
    """per-vertex speed of sound: FEM.jl:2283-2424 (cc1 stiffness) and :469-525 (c1 boundary mass)."""
    import wae_b200 as W
    from oracle.helmholtz import discretize as odisc
    from oracle.mesh import Mesh as OMesh
    raw = load_raw_mesh("rijke_mm")
    dscrp = {"Interior": ("interior", ()), "Outlet": ("admittance", ("Y", 1e15))}
    for order in ("lin", "quad"):
        mo = OMesh("m", scale=0.001, raw=raw)
        mg = W.Mesh("m", scale=0.001, raw=raw)
        cfun = lambda x, y, z: 347.0 + 1000.0 * z + 300 * x
        Lo = odisc(mo, dscrp, mo.generate_field(cfun, order="lin"), order=order)
        Lg = W.discretize(mg, dscrp, mg.generate_field(cfun, order="lin"), order=order)
        for tg, to in zip(Lg.terms, Lo.terms):
            _check_term(tg, to)


def test_assembly_is_deterministic():
    """The gather generation writes every nonzero once in a fixed order: bit-identical across runs."""
    raw = load_raw_mesh("rijke_mm")
    a = _gpu_family(raw, "quad").terms[1].coeff.csc()[2]
    b = _gpu_family(raw, "quad").terms[1].coeff.csc()[2]
    assert np.array_equal(a, b)
