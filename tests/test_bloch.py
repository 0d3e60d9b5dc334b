"""Bloch-periodic reduction (BASELINE.json config 4): oracle vs analytic periodic duct on the CPU,
CUDA path vs oracle on the GPU."""
import math

import numpy as np
import pytest
import scipy.sparse as sp

DOS, LC = 6, 0.2


def _unit_cell(n=(6, 3, 3)):
    from wae_b200.meshutils import kuhn_unit_cell
    return kuhn_unit_cell(n, (0, 0, 0), (LC, 0.1, 0.1), DOS=DOS, jitter=0.1, seed=3)


def _to_oracle(m):
    from oracle.mesh import Mesh as OMesh
    raw = (m.points, [], [list(map(int, t)) for t in m.triangles], [list(map(int, t)) for t in m.tetrahedra],
           {k: {"dimension": v["dimension"], "simplices": list(map(int, v["simplices"]))} for k, v in m.domains.items()})
    mo = OMesh("u", raw=raw)
    mo.lines = [list(map(int, l)) for l in m.lines]
    mo.dos = m.dos
    assert np.array_equal(np.array(mo.tetrahedra), m.tetrahedra)
    return mo


def test_oracle_bloch_matches_analytic_periodic_duct():
    """A DOS-periodic rigid duct: the b-th Bloch mode of one cell is the plane wave with k_x = 2 pi b / (DOS L_c)."""
    from oracle.helmholtz import discretize
    from oracle.nlevp import mslp
    mo = _to_oracle(_unit_cell())
    c = np.full(len(mo.tetrahedra), 340.0)
    L = discretize(mo, {"Interior": ("interior", ())}, c, order="quad", b="b")
    assert L.size() == 588 and len(L.terms) == 7
    for bb in (1, 2):
        L.params["b"] = complex(bb)
        w_exact = 340.0 * 2 * math.pi * bb / (DOS * LC)
        sol, n, flag = mslp(L, w_exact * 1.02, maxiter=12, tol=1e-9)
        assert flag == 0 and abs(sol.params["ω"] - w_exact) < 2e-5 * w_exact


def test_bloch_dof_maps_match_blochify():
    from oracle.helmholtz import blochify
    from wae_b200.meshutils import bloch_dof_maps
    m = _unit_cell((3, 2, 2))
    d = m.dos
    npts = m.points.shape[1]
    new, image, axis, red = bloch_dof_maps(m, "quad")
    dim = npts + len(m.lines)
    ii = np.arange(dim)
    parts = blochify(ii, ii, np.ones(dim), d.naxis, d.nxbloch, d.naxis + d.nxsector, d.naxis_ln + npts, d.naxis_ln + d.nxsector_ln + npts, npts)
    assert list(parts[0][0]) == list(new) and not parts[1][0] and not parts[2][0]  # (i,i): both folded or neither -> plain
    assert red == dim - d.nxbloch - d.nxbloch_ln and new.max() == red - 1


@pytest.mark.gpu
@pytest.mark.parametrize("order", ["lin", "quad"])
def test_bloch_terms_and_mslp_sweep_match_oracle(order):
    import wae_b200 as W
    from oracle.helmholtz import discretize as odisc
    from oracle.nlevp import mslp as omslp
    m = _unit_cell()
    mo = _to_oracle(m)
    c = np.where(m.points[2, m.tetrahedra].mean(axis=1) < 0.05, 340.0, 420.0)
    dscrp = {"Interior": ("interior", ()), "Outlet": ("admittance", ("Y", 0.3))}
    Lo = odisc(mo, dscrp, c, order=order, b="b")
    Lg = W.discretize(m, dscrp, c, order=order, b="b")
    assert [t.operator for t in Lg.terms] == [t.operator for t in Lo.terms]
    for tg, to in zip(Lg.terms, Lo.terms):
        A = sp.csc_matrix(to.coeff)
        A.sort_indices()
        colptr, rowval, nz = tg.coeff.csc()
        assert np.array_equal(colptr, A.indptr) and np.array_equal(rowval, A.indices), tg.operator
        assert np.abs(nz - A.data).max() <= 1e-12 * np.abs(A.data).max()
    # config-4 style sweep: mslp over shifts and Bloch numbers (replicas; the general LU serves the unsymmetric family)
    for bb in (0, 1, 2):
        Lo.params["b"] = Lg.params["b"] = complex(bb)
        for f0 in (300.0, 900.0):
            so, no, fo = omslp(Lo, f0, maxiter=15, tol=1e-9, scale=2 * math.pi)
            sg, ng, fg = W.mslp(Lg, f0, maxiter=15, tol=1e-9, scale=2 * math.pi, output=False)
            assert fo == fg == 0
            # (b=0 from 300 Hz converges to the trivial mode omega = 0: compare on the scale of the shift there)
            assert abs(so.params["ω"] - sg.params["ω"]) <= 1e-10 * max(abs(so.params["ω"]), 2 * math.pi * f0), (bb, f0)
    v = W.bloch_expand(m, sg)
    assert len(v) == m.dos.naxis + m.dos.nxsector * DOS
