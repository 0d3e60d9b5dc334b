"""The product's HOST MIRROR of the reference interface (wavesandeigenvalues.jl_b200/{helmholtz,nlevp,shape}.py) exercised without a GPU:
the context object is replaced by the test double of tests/host_standin.py (oracle element routines + scipy behind the Context method
names), so what is tested here is the Python above the C ABI -- descriptor parsing, term and parameter bookkeeping, the Newton / mslp
loops with their flags, perturbation series, the forced-response wiring and the argument marshalling of the shape-sensitivity sequence --
against the reference's golden values and the oracle.  The CUDA kernels themselves are covered by the `-m gpu` tests only."""
import math

import numpy as np
import pytest
import scipy.sparse.linalg as spla

import wae_b200 as W
from cases import G9_APPROX_20, GAMMA, N_REF, Q02U0, RHO, X_REF, load_raw_mesh, rijke_dscrp, speedofsound
from host_standin import HostStandIn
from oracle import helmholtz as ohelm
from oracle import mesh as omesh

TOL = 1e-10


@pytest.fixture(scope="module")
def rijke():
    raw = load_raw_mesh("rijke_mm")
    mg, mo = W.Mesh("m", scale=0.001, raw=raw), omesh.Mesh("m", scale=0.001, raw=raw)
    return mg, mo, mg.generate_field(speedofsound)


def test_discretize_terms_and_householder_golden(rijke):
    """G1 (tutorial_04 notebook cell 5) through the product's discretize + LinearOperatorFamily + householder."""
    mg, mo, c = rijke
    ctx = HostStandIn()
    L = W.discretize(mg, rijke_dscrp(0.01, 0.001), c, ctx=ctx)
    Lo = ohelm.discretize(mo, rijke_dscrp(0.01, 0.001), c)
    assert [(t.operator, t.symbol, t.params) for t in L.terms] == [(t.operator, t.symbol, t.params) for t in Lo.terms]
    assert L.params == Lo.params
    for tg, to in zip(L.terms, Lo.terms):
        assert abs(tg.coeff.to_scipy() - to.coeff).max() <= 1e-12 * abs(to.coeff).max()
    sol, n, flag = W.householder(L, 340 * 2 * math.pi, maxiter=20, tol=1e-11, output=False)
    g1 = 1710.6977772393461 + 9.615018460173488j
    assert flag == 1 and n in (6, 7, 8) and abs(sol.params["ω"] - g1) / abs(g1) < TOL
    assert L.active == ["ω"] and L.mode == "all"  # restored (LinOpFam.jl:546-560 toggles them around the perturbation calls)
    # the returned vectors are normalised as at Householder.jl:189-190
    M = -L.terms[-1].coeff.to_scipy()
    assert abs(np.vdot(sol.v, M @ sol.v) - 1) < 1e-10


def test_mslp_golden_flags_and_perturb_fast(rijke):
    """G4 (mslp, flag 0) and G9 (20th-order perturb_fast! evaluated at tau + 0.5 ms) through the product's host loops."""
    mg, mo, c = rijke
    L = W.discretize(mg, rijke_dscrp(1.0, 0.001), c, ctx=HostStandIn())
    sol, n, flag = W.mslp(L, 340 * 2 * math.pi, maxiter=20, tol=1e-11, output=False)
    g4 = 1075.325211506839 + 372.1017670372039j
    assert flag == 0 and abs(sol.params["ω"] - g4) / abs(g4) < TOL
    sol2, n2, flag2 = W.mslp(L, 340 * 2 * math.pi, maxiter=3, tol=1e-11, output=False)
    assert flag2 == 1 and n2 == 3  # maxiter reached (iterative_solvers.jl:232-234)
    W.perturb_fast_bang(sol, L, "τ", 20)
    assert abs(sol("τ", 0.001 + 0.0005, 20) - G9_APPROX_20) / abs(G9_APPROX_20) < 1e-9


@pytest.mark.parametrize("order", ["lin", "quad"])
def test_forced_response_wiring(rijke, order):
    """discretize(...; source=True) / :speaker / VectorFamily / L(ω).solve(rhs(ω)) against the oracle (tutorial_09_forcing.md)."""
    mg, mo, c = rijke
    dscrp = rijke_dscrp(0.01, 0.001)
    dscrp["Outlet"] = ("speaker", ("A", 1, "Y", 1e15))
    L, rhs = W.discretize(mg, dscrp, c, order=order, source=True, ctx=HostStandIn())
    Lo, rhso = ohelm.discretize(mo, dscrp, c, order=order, source=True)
    assert rhs.params == rhso.params and [(t.operator, t.params) for t in rhs.terms] == [(t.operator, t.params) for t in rhso.terms]
    assert [(t.operator, t.params) for t in L.terms] == [(t.operator, t.params) for t in Lo.terms]
    w = 2 * math.pi * 150.0
    b, bo = rhs(w), rhso(w).toarray().ravel()
    assert np.abs(b - bo).max() <= 1e-13 * np.abs(bo).max()
    sol, want = L(w).solve(b), spla.spsolve(Lo(w).tocsc(), bo)
    assert np.abs(sol - want).max() <= 1e-9 * np.abs(want).max()
    rhs.params["A"] = 2.0  # tutorial_09_forcing.md:82-86: the excitation level can be reset afterwards
    assert np.allclose(rhs(w), 2 * b, rtol=1e-14, atol=0)
    # two speakers, per-point c, functional admittance; in-place re-assembly refreshes the source vectors
    Yf = lambda w_, k=0: (2.0 + 0.001j * w_) if k == 0 else (0.001j if k == 1 else 0.0)
    cpt = np.array([speedofsound(*mo.points[:, i]) * (1 + 0.1 * math.sin(40 * mo.points[2, i])) for i in range(mo.points.shape[1])])
    dscrp2 = {"Interior": ("interior", ()), "Outlet": ("speaker", ("A", 1.5, "Y", 0.7)), "Inlet": ("speaker", ("B", 0.5j, Yf))}
    L2, rhs2 = W.discretize(mg, dscrp2, cpt, order=order, source=True, ctx=HostStandIn())
    _, rhso2 = ohelm.discretize(mo, dscrp2, cpt, order=order, source=True)
    bo = rhso2(w).toarray().ravel()
    assert len(rhs2.terms) == 2 and np.abs(rhs2(w) - bo).max() <= 1e-13 * np.abs(bo).max()
    L2.discretization.reassemble(1.1 * cpt)
    _, rhso3 = ohelm.discretize(mo, dscrp2, 1.1 * cpt, order=order, source=True)
    b3 = rhso3(w).toarray().ravel()
    assert np.abs(rhs2(w) - b3).max() <= 1e-13 * np.abs(b3).max()
    with pytest.raises(ValueError):
        W.discretize(mg, {"Interior": ("interior", ()), "Outlet": ("speaker", ("A", 1))}, c, source=True, ctx=HostStandIn())
    # one context, two families: a context holds one mesh at a time, so re-assembling the OLDER family (other element order) after a
    # newer discretize() must make its own mesh resident again
    ctx = HostStandIn()
    other = "quad" if order == "lin" else "lin"
    La = W.discretize(mg, rijke_dscrp(0.01, 0.001), c, order=order, ctx=ctx)
    ref = [t.coeff.csc()[2].copy() for t in La.terms]
    W.discretize(mg, rijke_dscrp(0.01, 0.001), c, order=other, ctx=ctx)
    La.discretization.reassemble(c)
    assert all(np.array_equal(t.coeff.csc()[2], r) for t, r in zip(La.terms, ref))


def test_shape_sensitivity_call_sequence(rijke):
    """discrete_adjoint_shape_sensitivity of the product (normalisation, descriptor -> terms, per-point lists, begin/add/end marshalling)
    with the kernel's per-thread function replayed on the host, against the oracle's literal loop on a subset of the surface points."""
    from oracle import nlevp as onlevp
    from oracle import shape as oshape
    mg, mo, c = rijke
    ref_idx = mg.find_tetrahedron_containing_point(X_REF)
    dscrp = {"Interior": ("interior", ()), "Outlet": ("admittance", ("Y", 1e15)),
             "Flame": ("flame", (GAMMA, RHO, Q02U0, ref_idx, X_REF, N_REF, "n", "τ", 1.0, 0.001))}
    ctx = HostStandIn()
    L = W.discretize(mg, dscrp, c, ctx=ctx)
    sol, n, flag = W.householder(L, 700 * 2 * math.pi, maxiter=14, tol=1e-11, output=False)
    assert flag == 1
    sp_, trm, ttm = W.get_surface_points(mg)
    flame = set(np.asarray(mg.tetrahedra)[mg.domains["Flame"]["simplices"]].ravel().tolist())
    sub = sorted(set(range(0, len(sp_), 60)) | set([k for k, p in enumerate(sp_) if p in flame][:4]))
    pick = lambda lst: [lst[k] for k in sub]
    sens = W.discrete_adjoint_shape_sensitivity(mg, dscrp, c, sp_[sub], pick(trm), pick(ttm), L, sol, ctx=ctx)
    assert sens.shape == (3, mg.points.shape[1]) and np.count_nonzero(np.abs(sens).sum(axis=0)) == len(sub)
    Lo = ohelm.discretize(mo, dscrp, c)
    solo, _, _ = onlevp.householder(Lo, 700 * 2 * math.pi, maxiter=14, tol=1e-11)
    so, tro, tto = oshape.get_surface_points(mo)
    want = oshape.discrete_adjoint_shape_sensitivity(mo, dscrp, c, [so[k] for k in sub], [tro[k] for k in sub], [tto[k] for k in sub], Lo, solo)
    assert np.abs(sens - want).max() <= 1e-5 * np.abs(want).max()
    # the family and the solution survive the call (the context keeps one mesh at a time: same topology, patterns stay valid)
    assert abs(L(sol.params["ω"]).matvec(sol.v)).max() <= 1e-6 * abs(L(sol.params["ω"], 1).matvec(sol.v)).max() * abs(sol.params["ω"])


def test_bloch_discretize_and_unit_cell_shape_sensitivity():
    """Config-4 host path on the reference's NTNU_12 mesh: extend_mesh(unit=true) -> discretize(b) (term list, class scalars, folded
    dimension, penalty term) vs the oracle; mslp at b = 1 next to the plenum mode; then the unit-cell branch of
    discrete_adjoint_shape_sensitivity (shape_sensitivity.jl:84-118: axis points skipped, Bloch-plane points paired with their images,
    cylindrical moves, DOF folding and class phase handed to the kernel) against the independent replay of test_shape_sensitivity.py for
    ALL 846 surface points."""
    from oracle import nlevp as onlevp
    from oracle.mesh import extend_mesh as oext
    from test_bloch import NTNU_DOMS, NTNU_DSCRP, _ntnu_meshes, _ntnu_sos
    from test_shape_sensitivity import host_replay
    mg, mo = _ntnu_meshes()
    doms = NTNU_DOMS + [("CC", "half")]
    g, o = W.extend_mesh(mg, doms, unit=True), oext(mo, doms, unit=True)
    c = g.generate_field(_ntnu_sos)
    dscrp = dict(NTNU_DSCRP)
    dscrp["Outlet_high"] = ("admittance", ("Y_in", 0.2 + 0.1j))
    ctx = HostStandIn()
    L = W.discretize(g, dscrp, c, b="b", ctx=ctx)
    Lo = ohelm.discretize(o, dscrp, c, b="b")
    assert [t.operator for t in L.terms] == [t.operator for t in Lo.terms] and L.size() == Lo.size()
    for bb in (0, 1, 2):
        L.params["b"] = Lo.params["b"] = complex(bb)
        A, Ao = L(5000.0).to_scipy(), Lo(5000.0)
        assert abs(A - Ao).max() <= 1e-12 * abs(Ao).max()
    L.params["b"] = Lo.params["b"] = 1 + 0j
    sol, n, flag = W.mslp(L, 1146.0, maxiter=20, tol=1e-9, scale=2 * math.pi, output=False)
    solo, _, flo = onlevp.mslp(Lo, 1146.0, maxiter=20, tol=1e-9, scale=2 * math.pi)
    assert flag == flo == 0 and abs(sol.params["ω"] - solo.params["ω"]) <= 1e-9 * abs(solo.params["ω"])
    assert 1000 < sol.params["ω"].real / 2 / math.pi < 1250
    sp_, trm, ttm = W.get_surface_points(g)
    sens = W.discrete_adjoint_shape_sensitivity(g, dscrp, c, sp_, trm, ttm, L, sol, ctx=ctx)
    w0 = sol.params["ω"]
    v0 = sol.v / np.sqrt(np.vdot(sol.v, sol.v))
    va = sol.v_adj / np.conj(np.vdot(sol.v_adj, L(w0, 1) @ v0))
    rep = host_replay(g, dscrp, c, sp_, trm, ttm, w0, v0, va)
    scale = np.abs(rep).max()
    assert scale > 0 and np.abs(sens - rep).max() <= 1e-12 * scale
    assert np.abs(sens[:, : g.dos.naxis]).max() == 0 and np.count_nonzero(np.abs(sens).sum(axis=0)) == len(sp_) - g.dos.naxis


def test_tutorial_01_third_order_mslp_and_descriptor_variants(rijke):
    """The host loop behind tests/test_zw_tutorial_01_gpu.py (mslp(L, 245*2*pi - 82im*2*pi, order=3) at n = 1 -> G4, growth rate 59.22) and a
    sweep over descriptor variants (:fancyflame scalar and summed, state-space and functional admittance, :flameresponse, per-point c):
    the product's term list, parameters and L(z) against the oracle's."""
    mg, mo, c = rijke
    L = W.discretize(mg, rijke_dscrp(1.0, 0.001), c, ctx=HostStandIn())
    sol, n, flag = W.mslp(L, (245 - 82j) * 2 * math.pi, order=3, maxiter=15, tol=1e-10, output=False)
    g4 = 1075.325211506839 + 372.1017670372039j
    assert flag == 0 and abs(sol.params["ω"] - g4) / abs(g4) < TOL and round(abs(sol.params["ω"].imag) / 2 / math.pi, 2) == 59.22
    sol3, n3, flag3 = W.householder(L, 170 * 2 * math.pi + 350j, maxiter=15, tol=1e-10, order=3, nev=2, output=False)
    assert flag3 == 1 and abs(sol3.params["ω"] - g4) / abs(g4) < 1e-9

    Yf = lambda w_, k=0: (2.0 + 0.001j * w_) if k == 0 else (0.001j if k == 1 else 0.0)
    Ass, Bss, Css, Dss = np.array([[-50.0, 20.0], [-20.0, -80.0]]), np.array([[1.0], [0.5]]), np.array([[0.3, -0.2]]), np.array([[0.05]])
    cpt = np.array([speedofsound(*mo.points[:, i]) * (1 + 0.05 * math.cos(30 * mo.points[2, i])) for i in range(mo.points.shape[1])])
    flame = (GAMMA, RHO, Q02U0, X_REF, N_REF)
    variants = [
        ({"Interior": ("interior", ()), "Outlet": ("admittance", (Yf,)), "Flame": ("fancyflame", flame + ("n", "τ", "a", 0.8, 0.0011, 1e-9))}, c),
        ({"Interior": ("interior", ()), "Outlet": ("admittance", (Ass, Bss, Css, Dss)),
          "Flame": ("fancyflame", flame + (["n1", "n2"], ["t1", "t2"], ["a1", "a2"], [0.5, 0.3], [0.001, 0.002], [1e-9, 2e-9]))}, c),
        ({"Interior": ("interior", ()), "Inlet": ("admittance", ("Yin", 0.4 - 0.1j)), "Outlet": ("admittance", ("Y", 1e15)),
          "Flame": ("flameresponse", flame + ("ε", 0.02))}, cpt),
        # docs/src/tutorial_08_custom_FTF.md: a user closure FTF(ω, k) (value and derivatives), and the plain parameter :FTF
        ({"Interior": ("interior", ()), "Outlet": ("admittance", ("Y", 1e15)),
          "Flame": ("flame", flame + (lambda w_, k=0: (0.7 * np.exp(-0.001j * w_) * (-0.001j) ** k),))}, c),
        ({"Interior": ("interior", ()), "Outlet": ("admittance", ("Y", 1e15)), "Flame": ("flame", flame)}, c),
    ]
    for dscrp, cc in variants:
        Lg, Lo = W.discretize(mg, dscrp, cc, order="quad", ctx=HostStandIn()), ohelm.discretize(mo, dscrp, cc, order="quad")
        assert [(t.operator, t.params) for t in Lg.terms] == [(t.operator, t.params) for t in Lo.terms]
        assert Lg.params.keys() == Lo.params.keys() and all(Lg.params[k] == Lo.params[k] or (Lg.params[k] != Lg.params[k]) for k in Lo.params)
        if "FTF" in Lg.params:
            assert Lg.params["FTF"] == Lo.params["FTF"] == 0
            Lg.params["FTF"] = Lo.params["FTF"] = 0.3 - 0.2j
        for z in (900.0 + 30j, 2500.0 - 10j):
            A, Ao = Lg(z).to_scipy(), Lo(z)
            assert abs(A - Ao).max() <= 1e-12 * abs(Ao).max()
            A1, Ao1 = Lg(z, 1).to_scipy(), Lo(z, 1)
            assert abs(A1 - Ao1).max() <= 1e-12 * abs(Ao1).max()


def test_tutorial_08_through_the_host_mirror(rijke):
    """The same tutorial_08 sequence through the product's discretize / mslp / perturb_fast! / pade."""
    from cases import tutorial_08_check
    mg, mo, c = rijke
    ctx = HostStandIn()
    tutorial_08_check(W.discretize, lambda L, z, **k: W.mslp(L, z, output=False, **k), W.perturb_fast_bang, W.pade, W.polyval, mg, c, ctx=ctx)


def test_beyn_host_logic(rijke):
    """beyn / compute_moment_matrices / moments2eigs of the product (node generation, per-node scalars, K from l and d, probing matrix,
    singular-value cut, position test) against the oracle on the tutorial's contour (docs/src/tutorial_01_rijke_tube.md:202-212, N reduced)."""
    from oracle.nlevp import beyn as obeyn
    from oracle.nlevp import beyn_moments as omoments
    mg, mo, c = rijke
    L = W.discretize(mg, rijke_dscrp(0.0, 0.001), c, ctx=HostStandIn())
    Lo = ohelm.discretize(mo, rijke_dscrp(0.0, 0.001), c)
    G = [z * 2 * math.pi for z in (150 + 5j, 150 - 5j, 1000 - 5j, 1000 + 5j)]
    Ao = omoments(Lo, G, 5, 1, 8)
    Ag = W.compute_moment_matrices(L, G, l=5, K=1, N=8)
    assert Ag.shape == Ao.shape == (L.size(), 5, 2) and np.abs(Ag - Ao).max() <= 1e-10 * np.abs(Ao).max()
    Oo, Po = obeyn(Lo, G, l=5, N=8, tol=1e-8)
    Og, Pg = W.beyn(L, G, l=5, N=8, tol=1e-8, output=False)
    assert len(Og) == len(Oo) == 2 and np.abs(np.sort_complex(Og) - np.sort_complex(Oo)).max() <= 1e-8 * np.abs(Oo).max()
    f = np.sort(Og.real) / 2 / math.pi
    assert abs(f[0] - 272) < 2 and abs(f[1] - 695) < 10
    # higher moments (K = 2) and a seeded random probing matrix give the same two eigenvalues
    O2, _ = W.beyn(L, G, l=3, K=2, N=8, tol=1e-8, output=False, random=True, seed=3)
    assert len(O2) == 2 and np.abs(np.sort_complex(O2) - np.sort_complex(Og)).max() <= 1e-3 * np.abs(Og).max()


def test_release_drops_and_rebuilds_the_device_side(rijke):
    """L.release() (wae_lu_free + wae_family_free: what Julia's GC does for the reference's sparse sums and UMFPACK factors) returns the
    device side of a family; the next use rebuilds it and gives the same eigenvalue.  A family with a live LU handle cannot be freed."""
    from wae_b200 import _lib
    mg, mo, c = rijke
    ctx = HostStandIn()
    L = W.discretize(mg, rijke_dscrp(0.01, 0.001), c, ctx=ctx)
    sol, _, _ = W.householder(L, 340 * 2 * math.pi, maxiter=20, tol=1e-11, output=False)
    dev = L.device()
    fid, lid = dev.fid, dev.lu()
    with pytest.raises(_lib.WaeError):
        ctx.family_free(fid)  # its LU handle is still alive
    L.release()
    assert ctx.fams[fid] is None and ctx.lus[lid] is None and L._dev is None
    with pytest.raises(_lib.WaeError):
        ctx.lu_free(lid)  # ids are never reused
    sol2, _, _ = W.householder(L, 340 * 2 * math.pi, maxiter=20, tol=1e-11, output=False)
    assert L.device().fid != fid and abs(sol2.params["ω"] - sol.params["ω"]) <= 1e-12 * abs(sol.params["ω"])


def test_paired_eigs_switch_of_the_host_loop(rijke, monkeypatch):
    """WAE_EIGS_PAIRED=1: the host loop asks the context for both Arnoldi recurrences in one call (wae_eigs_si_pair) and falls back to two
    wae_eigs_si calls when the context refuses (general elimination); G1 either way.  (The test double has no pair routine of its own: the
    two classes below put one over its eigs_si, so that this covers the Python branch only.)"""
    from wae_b200 import _lib
    mg, mo, c = rijke
    calls = {"pair": 0, "refused": 0}

    class Pairing(HostStandIn):
        def eigs_si_pair(self, lid, fid, m_slot, nev, v0, v0_adj):
            calls["pair"] += 1
            lam, V, n1 = self.eigs_si(lid, fid, m_slot, nev, v0, trans=0)
            lam_a, Va, n2 = self.eigs_si(lid, fid, m_slot, nev, v0_adj, trans=2)
            return lam, V, lam_a, Va, n1 + n2

    class Refusing(HostStandIn):
        def eigs_si_pair(self, *a):
            calls["refused"] += 1
            raise _lib.WaeError(_lib.E_INVALID, "needs a symmetric-mode factorisation")

    monkeypatch.setenv("WAE_EIGS_PAIRED", "1")
    g1 = 1710.6977772393461 + 9.615018460173488j
    for cls in (Pairing, Refusing):
        L = W.discretize(mg, rijke_dscrp(0.01, 0.001), c, ctx=cls())
        stats = {}
        sol, n, flag = W.householder(L, 340 * 2 * math.pi, maxiter=20, tol=1e-11, output=False, stats=stats)
        assert flag == 1 and abs(sol.params["ω"] - g1) / abs(g1) < TOL and stats["solves"] > 0
    assert calls["pair"] >= 6 and calls["refused"] == 1  # the refusal is remembered for the rest of the run


def test_first_order_perturbation_without_the_normalisations(rijke):
    """householder / mslp take their Newton step from perturb(..., vectors=False): lam_1 = -(v_adj^H L01 v) / (v_adj^H L10 v) is free of the
    scaling of the vector pair, so the shortcut must reproduce the coefficient of the reference's normalised recipe (perturbation.jl:319-367)."""
    from wae_b200.nlevp import perturb
    mg, mo, c = rijke
    L = W.discretize(mg, rijke_dscrp(0.01, 0.001), c, ctx=HostStandIn())
    sol, n, flag = W.householder(L, 340 * 2 * math.pi, maxiter=20, tol=1e-11, output=False)
    rng = np.random.default_rng(5)
    v = sol.v * (3.0 - 2.0j) + 1e-3 * (rng.standard_normal(L.size()) + 1j * rng.standard_normal(L.size()))
    va = sol.v_adj * (0.5 + 4.0j) + 1e-3 * (rng.standard_normal(L.size()) + 1j * rng.standard_normal(L.size()))
    active, mode = L.active, L.mode
    L.active, L.mode = ["ω", "τ"], "householder"
    try:
        lam_ref, vecs = perturb(L, 1, v, va)
        lam_fast, none = perturb(L, 1, v, va, vectors=False)
    finally:
        L.active, L.mode = active, mode
    assert none == [None, None] and vecs[0] is not None
    assert abs(lam_fast[1] - lam_ref[1]) <= 1e-12 * abs(lam_ref[1])
