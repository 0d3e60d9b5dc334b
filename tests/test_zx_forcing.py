"""Forced response (docs/src/tutorial_09_forcing.md:44,74,99-100): ``L, rhs = discretize(mesh, dscrp, c, source=true)`` with a
``:speaker`` boundary, then ``sol = L(ω) \\ Array(rhs(ω))`` (Helmholtz.jl:251-258, 488-503, 524-526, 576-577; wallsrc :193-210;
element vectors FEM.jl:2557-2589, pinned to the reference's expressions in tests/golden/fem_tables.npz by test_oracle_golden.py).

The reference stores no output of this path ("parity unpinned"), so the oracle is pinned to the closed-form solution of the same
boundary-value problem in a uniform duct; the CUDA path (wae_assemble_wallsrc + combine + LU solve) is compared with the oracle.

NOTE (round 1): added after the round's GPU budget was spent -- the GPU test below has not run on a B200 yet; the file sorts late so
that `pytest -x` reaches it after the tests that have."""
import math

import numpy as np
import pytest
import scipy.sparse.linalg as spla

import wae_b200 as W
from cases import load_raw_mesh, speedofsound
from oracle import helmholtz as ohelm
from oracle import mesh as omesh


def _oracle_mesh(mesh):
    raw = (mesh.points, [], [list(map(int, t)) for t in mesh.triangles], [list(map(int, t)) for t in mesh.tetrahedra],
           {k: {"dimension": v["dimension"], "simplices": list(map(int, v["simplices"]))} for k, v in mesh.domains.items()})
    return omesh.Mesh("m", raw=raw)


def _duct(nz=40, order="quad"):
    mesh = W.kuhn_box((2, 2, nz), (0, 0, 0), (0.05, 0.05, 0.5), jitter=0.15, seed=3)
    return mesh, _oracle_mesh(mesh)


def test_speaker_family_structure_and_total_source():
    """rhs is a family of vectors with the scalar (boundary scalar..., speaker level); sum_i m_i = -i * int_Gamma c (the bases sum to one)."""
    mesh, mo = _duct(8)
    area = 0.05 * 0.05
    for order in ("lin", "quad"):
        dscrp = {"Interior": ("interior", ()), "Inlet": ("speaker", ("A", 2.0, "Y", 0.5)), "Outlet": ("admittance", ("Y", 7.0))}
        L, rhs = ohelm.discretize(mo, dscrp, np.full(len(mo.tetrahedra), 347.0), order=order, source=True)
        assert [t.operator for t in rhs.terms] == ["m"] and rhs.terms[0].params == (("ω",), ("Y",), ("A",))
        assert rhs.params == {"ω": 0j, "A": 2 + 0j, "Y": 0.5 + 0j} and L.params["Y"] == 0.5  # first definition wins (Helmholtz.jl:266-271)
        m = rhs.terms[0].coeff.toarray().ravel()
        assert m.shape == (L.size(),)
        assert abs(m.sum() - (-1j * 347.0 * area)) <= 1e-12 * 347.0 * area
        inlet = set(np.asarray(mo.triangles)[mo.domains["Inlet"]["simplices"]].ravel().tolist())
        assert all(abs(mo.points[2, i]) < 1e-12 for i in np.flatnonzero(m[: mo.points.shape[1]]))
        assert set(np.flatnonzero(m[: mo.points.shape[1]]).tolist()) <= inlet
        if order == "quad":
            assert not np.any(m[list(inlet)])  # FEM.jl:2561-2563: the P2 vertex functions integrate to zero over a triangle
        w = 2 * math.pi * 100.0
        assert np.allclose(rhs(w).toarray().ravel(), w * 0.5 * 2.0 * m, rtol=1e-14, atol=0)
        # per-point speed of sound: sum_i m_i = -i * int c with c linear per triangle
        cpt = 300.0 + 1000.0 * mo.points[0] + 500.0 * mo.points[1]
        _, rhs2 = ohelm.discretize(mo, dscrp, cpt, order=order, source=True)
        want = -1j * (300.0 + 1000.0 * 0.025 + 500.0 * 0.025) * area
        assert abs(rhs2.terms[0].coeff.toarray().sum() - want) <= 1e-12 * abs(want)


def _duct_closed_form(z, w, c, l, Y0, A, Yl):
    """p'' + k^2 p = 0 on [0, l];  z = 0 (outward normal -z): -c p' + i w Y0 (p - A) = 0;  z = l: c p' + i w Yl p = 0 -- the strong
    form of  (w^2 M + K + w Y C) p = w Y A m  with K = -c^2 int grad.grad, C = -i c int_Gamma phi phi, m = -i c int_Gamma phi."""
    k = w / c
    S = np.array([[1j * w * Y0, -c * k],
                  [-c * k * math.sin(k * l) + 1j * w * Yl * math.cos(k * l), c * k * math.cos(k * l) + 1j * w * Yl * math.sin(k * l)]])
    a, b = np.linalg.solve(S, np.array([1j * w * Y0 * A, 0.0]))
    return a * np.cos(k * z) + b * np.sin(k * z)


@pytest.mark.parametrize("f_hz, Y0, Yl, tol", [(150.0, 0.8 + 0.2j, 1.5, 2e-6), (420.0, 1e15, 0.3 - 0.1j, 5e-5)])
def test_oracle_forced_response_matches_the_duct_solution(f_hz, Y0, Yl, tol):
    """Pins the oracle's :speaker / source=true restatement: plane-wave response of a uniform duct driven by a membrane of admittance
    Y0 (Y0 = 1e15: the tutorial's prescribed-pressure limit p(0) = A) against an impedance end.  P2, 40 cells over 0.5 m."""
    c0 = 347.0
    dscrp = {"Interior": ("interior", ()), "Inlet": ("speaker", ("A", 1.0, "Y", Y0)), "Outlet": ("admittance", ("Yout", Yl))}
    w = 2 * math.pi * f_hz
    errs = []
    for nz in (20, 40):
        mesh, mo = _duct(nz)
        L, rhs = ohelm.discretize(mo, dscrp, np.full(len(mo.tetrahedra), c0), order="quad", source=True)
        rhs.params["A"] = 3.0 - 1.0j  # parameters can be reset after the discretisation (tutorial_09_forcing.md:82-86)
        sol = spla.spsolve(L(w).tocsc(), rhs(w).toarray().ravel())
        npts = mo.points.shape[1]
        want = _duct_closed_form(mo.points[2], w, c0, 0.5, Y0, 3.0 - 1.0j, Yl)
        errs.append(np.abs(sol[:npts] - want).max() / np.abs(want).max())
    assert errs[1] <= tol and errs[1] < errs[0] / 4  # discretisation error (5e-6 -> 1e-6 and 9e-5 -> 1e-5 from 20 to 40 axial cells)


@pytest.mark.gpu
@pytest.mark.parametrize("order", ["lin", "quad"])
def test_forced_response_on_the_gpu_matches_the_oracle(order):
    """tutorial_09_forcing.md on the Rijke mesh (speaker at the outlet, A = 1, Y = 1e15, active flame): the source vector, L(ω) and
    the LU solve on the device against the oracle + SuperLU; then per-point c and a second speaker with a functional admittance."""
    from cases import rijke_dscrp
    raw = load_raw_mesh("rijke_mm")
    mg, mo = W.Mesh("m", scale=0.001, raw=raw), omesh.Mesh("m", scale=0.001, raw=raw)
    c = mg.generate_field(speedofsound)
    dscrp = rijke_dscrp(0.01, 0.001)
    dscrp["Outlet"] = ("speaker", ("A", 1, "Y", 1e15))
    L, rhs = W.discretize(mg, dscrp, c, order=order, source=True)
    Lo, rhso = ohelm.discretize(mo, dscrp, c, order=order, source=True)
    assert rhs.params == rhso.params and [(t.operator, t.params) for t in rhs.terms] == [(t.operator, t.params) for t in rhso.terms]
    mo_ = rhso.terms[0].coeff.toarray().ravel()
    assert np.abs(rhs.terms[0].coeff - mo_).max() <= 1e-13 * np.abs(mo_).max()
    for f_hz in (150.0, 333.0):
        w = 2 * math.pi * f_hz
        b, bo = rhs(w), rhso(w).toarray().ravel()
        assert np.abs(b - bo).max() <= 1e-13 * np.abs(bo).max()
        sol = L(w).solve(b)
        want = spla.spsolve(Lo(w).tocsc(), bo)
        assert np.abs(sol - want).max() <= 1e-8 * np.abs(want).max()
        assert abs(sol[np.argmax(np.abs(mo_))] - 1.0) < 1e-6  # Y = 1e15: the membrane prescribes p = A

    Yf = lambda w_, k=0: (2.0 + 0.001j * w_) if k == 0 else (0.001j if k == 1 else 0.0)
    cpt = np.array([speedofsound(*mo.points[:, i]) * (1 + 0.1 * math.sin(40 * mo.points[2, i])) for i in range(mo.points.shape[1])])
    dscrp2 = {"Interior": ("interior", ()), "Outlet": ("speaker", ("A", 1.5, "Y", 0.7)), "Inlet": ("speaker", ("B", 0.5j, Yf))}
    L2, rhs2 = W.discretize(mg, dscrp2, cpt, order=order, source=True)
    Lo2, rhso2 = ohelm.discretize(mo, dscrp2, cpt, order=order, source=True)
    assert len(rhs2.terms) == len(rhso2.terms) == 2
    w = 2 * math.pi * 210.0
    b, bo = rhs2(w), rhso2(w).toarray().ravel()
    assert np.abs(b - bo).max() <= 1e-13 * np.abs(bo).max()
    sol, want = L2(w).solve(b), spla.spsolve(Lo2(w).tocsc(), bo)
    assert np.abs(sol - want).max() <= 1e-8 * np.abs(want).max()
    # in-place re-assembly refreshes the source vectors too
    L2.discretization.reassemble(1.1 * cpt)
    _, rhso3 = ohelm.discretize(mo, dscrp2, 1.1 * cpt, order=order, source=True)
    b3 = rhso3(w).toarray().ravel()
    assert np.abs(rhs2(w) - b3).max() <= 1e-13 * np.abs(b3).max()


def test_generated_kernel_tables_match_the_oracle_tables():
    """csrc/fem_gen.h (tools/gen_fem_tables.py) holds the literal tables the CUDA kernels multiply with |det|: mass, source and their
    linear-c variants for both element families -- compared here with the oracle's tables (themselves pinned to the reference's
    expressions in tests/golden/fem_tables.npz), incl. the WAE_*_TRI_SRC / _SRCC tables of the speaker kernel."""
    import os
    import re

    from oracle import fem
    src = open(os.path.join(os.path.dirname(os.path.abspath(W.__file__)), "csrc", "fem_gen.h")).read()
    tabs = {m.group(1): np.array([float(x) for x in m.group(2).split(",")])
            for m in re.finditer(r"const double (WAE_\w+)\[\d+\] = \{([^}]*)\};", src)}
    for tag, order in (("P1", 1), ("P2", 2)):
        t3, t4 = fem.tables(order, 3), fem.tables(order, 4)
        for name, want in ((f"WAE_{tag}_TRI_MASS", t3["mass"]), (f"WAE_{tag}_TRI_MASSC", t3["massc"]), (f"WAE_{tag}_TRI_SRC", t3["src"]),
                           (f"WAE_{tag}_TRI_SRCC", t3["srcc"]), (f"WAE_{tag}_TET_MASS", t4["mass"]), (f"WAE_{tag}_TET_MASSC", t4["massc"]),
                           (f"WAE_{tag}_TET_SRC", t4["src"])):
            assert name in tabs, name
            assert tabs[name].shape == (want.size,) and np.abs(tabs[name] - want.ravel()).max() <= 1e-16, name
