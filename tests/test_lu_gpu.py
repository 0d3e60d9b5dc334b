"""GPU parity: sparse LU (factor/solve), shift-invert Arnoldi and the NLEVP solvers vs the CPU oracle / goldens."""
import math

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from cases import load_raw_mesh, rijke_dscrp, speedofsound

pytestmark = pytest.mark.gpu

TOL = 1e-10  # BASELINE.json: eigenvalues within 1e-10 relative of the reference path


def _gpu_family(order="lin", n=0.01, tau=0.001):
    import wae_b200 as W
    mesh = W.Mesh("Rijke_mm.msh", scale=0.001, raw=load_raw_mesh("rijke_mm"))
    c = mesh.generate_field(speedofsound)
    return W.discretize(mesh, rijke_dscrp(n, tau), c, order=order)


@pytest.mark.parametrize("order", ["lin", "quad"])
def test_solve_matches_superlu(order):
    L = _gpu_family(order, n=1.0)
    d = L.size()
    rng = np.random.default_rng(11)
    B = rng.standard_normal((d, 3)) + 1j * rng.standard_normal((d, 3))
    for z in (340 * 2 * math.pi, 1075.0 + 372.0j):
        op = L(z)
        A = op.to_scipy()
        lu = spla.splu(sp.csc_matrix(A))
        dev = L.device()
        ctx = dev.ctx
        op.materialize(0)
        lid = dev.lu()
        ctx.lu_factor(lid, 0)
        for trans, Aop, code in ((0, A, "N"), (1, A.T, "T"), (2, A.conj().T, "H")):
            X = ctx.lu_solve(lid, B, trans=trans)
            Xref = lu.solve(B, trans=code)
            # backward error against the true matrix and agreement with SuperLU
            res = (np.abs(Aop @ X - B) / (abs(Aop) @ np.abs(X) + np.abs(B))).max()  # componentwise backward error
            assert res < 1e-12, (order, z, trans, res)
            assert np.abs(X - Xref).max() <= 1e-7 * np.abs(Xref).max()
        x1 = ctx.lu_solve(lid, B[:, 0])
        assert np.abs(x1 - lu.solve(B[:, 0])).max() <= 1e-7 * np.abs(x1).max()


def test_eigs_matches_arpack():
    L = _gpu_family("lin", n=0.01)
    z = 340 * 2 * math.pi
    dev = L.device()
    ctx = dev.ctx
    A = L(z).to_scipy()
    L(z).materialize(0)
    mc = [None] * len(L.terms)
    mc[-1] = -1.0
    dev.combine(dev.flat(mc), 1)
    M = -L.terms[-1].coeff.to_scipy()
    lid = dev.lu()
    ctx.lu_factor(lid, 0)
    v0 = np.ones(L.size(), dtype=complex)
    lam, V, ns = ctx.eigs_si(lid, dev.fid, 1, 1, v0, trans=0)
    lam_ref, v_ref = spla.eigs(sp.csc_matrix(A), k=1, M=sp.csc_matrix(M), sigma=0, v0=v0, tol=0)
    assert abs(lam[0] - lam_ref[0]) <= 1e-10 * abs(lam_ref[0])

    def phase(v):
        v = v / np.linalg.norm(v)
        k = np.argmax(np.abs(v))
        return v * (abs(v[k]) / v[k])
    # eigenvector parity after phase normalisation (BASELINE.json: 1e-8); the Y=1e15 penalty rows hold ~1e-17 values
    assert np.abs(phase(V[:, 0]) - phase(v_ref[:, 0])).max() <= 1e-8
    lam3, V3, _ = ctx.eigs_si(lid, dev.fid, 1, 3, v0, trans=0)
    ref3 = spla.eigs(sp.csc_matrix(A), k=3, M=sp.csc_matrix(M), sigma=0, v0=v0, tol=0)[0]
    assert np.abs(np.sort_complex(lam3) - np.sort_complex(ref3)).max() <= 1e-8 * np.abs(ref3).max()
    # adjoint problem eigs(A', M')
    lama, Va, _ = ctx.eigs_si(lid, dev.fid, 1, 1, np.conj(v0), trans=2)
    assert abs(lama[0] - np.conj(lam[0])) <= 1e-9 * abs(lam[0])


G_HOUSEHOLDER = [
    (0.001, 1710.6977772393461 + 9.615018460173488j),
    (0.001 + 0.00001, 1710.864199971756 + 9.593830019670127j),
    (0.001 + 0.0008798274754933992 + 0.001, 1707.4565281774599 - 9.11764397194075j),
]


@pytest.mark.parametrize("tau,omega", G_HOUSEHOLDER)
def test_householder_goldens_gpu(tau, omega):
    import wae_b200 as W
    L = _gpu_family("lin", n=0.01, tau=tau)
    # tol: |d omega| <= 1e-10 (6e-14 relative).  1e-11 sits at the rounding floor of the Newton step on the device (the fp64 atomics of the
    # factorisation do not fix the order of the additions), so that reaching it within maxiter was a matter of luck
    sol, n, flag = W.householder(L, 340 * 2 * math.pi, maxiter=20, tol=1e-10, output=False)
    assert flag in (0, 1)
    assert abs(sol.params["ω"] - omega) / abs(omega) < TOL


G_MSLP = [
    (0.001, 340 * 2 * math.pi, 1075.325211506839 + 372.1017670372039j),
    (0.0015, 916.7085040155473 + 494.3258317478708j, 916.7036137579256 + 494.32932528479967j),
    (0.001 + 2 * 0.0007029896606802446, 668.5373997804821 + 529.4636751544649j, 668.537399929804 + 529.4636746814361j),
]


@pytest.mark.parametrize("tau,start,omega", G_MSLP)
def test_mslp_goldens_gpu(tau, start, omega):
    import wae_b200 as W
    L = _gpu_family("lin", n=1.0, tau=tau)
    sol, n, flag = W.mslp(L, start, maxiter=20, tol=1e-10, output=False)
    assert flag == 0
    assert abs(sol.params["ω"] - omega) / abs(omega) < TOL


def test_householder_quad_and_eigenvectors_vs_oracle():
    """P2 elements, higher-order Householder update and eigenvector parity (1e-8 after phase normalisation)."""
    import wae_b200 as W
    from oracle.helmholtz import discretize as odisc
    from oracle.mesh import Mesh as OMesh
    from oracle.nlevp import householder as ohouse
    raw = load_raw_mesh("rijke_mm")
    mo = OMesh("m", scale=0.001, raw=raw)
    Lo = odisc(mo, rijke_dscrp(1.0, 0.001), mo.generate_field(speedofsound), order="quad")
    Lg = _gpu_family("quad", n=1.0)
    for order in (1, 3):
        so, no, fo = ohouse(Lo, 340 * 2 * math.pi, maxiter=25, tol=1e-10, order=order)
        sg, ng, fg = W.householder(Lg, 340 * 2 * math.pi, maxiter=25, tol=1e-10, order=order, output=False)
        wo, wg = so.params["ω"], sg.params["ω"]
        assert abs(wo - wg) / abs(wo) < TOL, (order, wo, wg)

        def phase(v):
            k = np.argmax(np.abs(v))
            return v * (abs(v[k]) / v[k])
        assert np.abs(phase(so.v) - phase(sg.v)).max() <= 1e-8 * np.abs(so.v).max()
        assert np.abs(phase(so.v_adj) - phase(sg.v_adj)).max() <= 1e-6 * np.abs(so.v_adj).max()


def test_beyn_gpu_vs_oracle():
    import wae_b200 as W
    from oracle.helmholtz import discretize as odisc
    from oracle.mesh import Mesh as OMesh
    from oracle.nlevp import beyn as obeyn
    raw = load_raw_mesh("rijke_mm")
    mo = OMesh("m", scale=0.001, raw=raw)
    Lo = odisc(mo, rijke_dscrp(0.0, 0.001), mo.generate_field(speedofsound))
    Lg = _gpu_family("lin", n=0.0)
    G = [z * 2 * math.pi for z in (150 + 5j, 150 - 5j, 1000 - 5j, 1000 + 5j)]
    # moment matrices agree to solver accuracy ...
    from oracle.nlevp import beyn_moments as omoments
    Ao = omoments(Lo, G, 5, 1, 16)
    Ag = W.compute_moment_matrices(Lg, G, l=5, K=1, N=16)
    assert np.abs(Ag - Ao).max() <= 1e-9 * np.abs(Ao).max()
    # ... and so do the eigenvalues once the noise singular values (~1e-14 vs 1e2) are cut (tol is absolute, beyn.jl:92-95)
    Oo, Po = obeyn(Lo, G, l=5, N=16, tol=1e-8)
    Og, Pg = W.beyn(Lg, G, l=5, N=16, tol=1e-8, output=False)
    assert len(Og) == len(Oo) == 2
    assert np.abs(np.sort_complex(Og) - np.sort_complex(Oo)).max() <= 1e-8 * np.abs(Oo).max()
    # the reference's own idiom (tutorial_06...jl:41-55): polish with householder
    for om in Og:
        sol, n, flag = W.householder(Lg, om, maxiter=10, tol=1e-10, output=False)
        assert abs(sol.params["ω"] - om) < 5e-2 * abs(om)


def test_fancyflame_and_state_space_admittance_match_oracle():
    """Descriptor variants of Helmholtz.discretize (Helmholtz.jl:279-285, 363-400): Gaussian-filtered n-tau flame (:fancyflame, scalar
    and summed form) and state-space outlet admittance -- the flame term is not a rank-1 update of a symmetric family only through
    its scalar, the admittance scalar is a rational function; householder must agree with the oracle to 1e-10."""
    import wae_b200 as W
    from cases import GAMMA, N_REF, Q02U0, RHO, X_REF
    from oracle.helmholtz import discretize as odisc
    from oracle.mesh import Mesh as OMesh
    from oracle.nlevp import householder as ohouse
    raw = load_raw_mesh("rijke_mm")
    mg, mo = W.Mesh("Rijke_mm.msh", scale=0.001, raw=raw), OMesh("Rijke_mm.msh", scale=0.001, raw=raw)
    A, B, Cm, D = np.array([[-2.0e3, 0.0], [0.0, -5.0e3]]), np.array([1.0, 1.0]), np.array([4.0e3, -1.0e3]), np.array([0.5])
    cases = {
        "scalar": {"Interior": ("interior", ()), "Outlet": ("admittance", (A, B, Cm, D)),
                   "Flame": ("fancyflame", (GAMMA, RHO, Q02U0, X_REF, N_REF, "n", "τ", "a", 1.0, 0.001, -1e-8))},
        "summed": {"Interior": ("interior", ()), "Outlet": ("admittance", ("Y", 1e15)),
                   "Flame": ("fancyflame", (GAMMA, RHO, Q02U0, X_REF, N_REF, ("n1", "n2"), ("τ1", "τ2"), ("a1", "a2"),
                                            (0.6, 0.4), (0.001, 0.0013), (-1e-8, -2e-8)))},
    }
    for name, dscrp in cases.items():
        Lg = W.discretize(mg, dscrp, mg.generate_field(speedofsound))
        Lo = odisc(mo, dscrp, mo.generate_field(speedofsound))
        assert [t.operator for t in Lg.terms] == [t.operator for t in Lo.terms]
        z = 1000.0 + 200.0j
        assert abs(Lg(z).to_scipy() - Lo(z)).max() <= 1e-11 * abs(Lo(z)).max(), name
        assert abs(Lg(z, 1).to_scipy() - Lo(z, 1)).max() <= 1e-11 * abs(Lo(z, 1)).max(), name
        # tol: 1e-13 relative; 1e-11 sits at the rounding floor of the step |d omega| (2e-11 ... 5e-12 from one iteration to the next, on the
        # GPU and in the oracle alike) and made the iteration count a coin toss
        sg, ng, fg = W.householder(Lg, 340 * 2 * math.pi, maxiter=25, tol=1e-10, output=False)
        so, no, fo = ohouse(Lo, 340 * 2 * math.pi, maxiter=25, tol=1e-10)
        assert fg == fo and fg >= 0
        assert abs(sg.params["ω"] - so.params["ω"]) <= TOL * abs(so.params["ω"]), (name, sg.params["ω"], so.params["ω"])


def test_perturb_fast_goldens_on_the_gpu():
    """G8/G9 (tutorial_04_perturbation_theory.md:112-155) through the CUDA path: one factorisation of L(0,0), 20 solves, one
    combine + SpMV per derivative pair; plus the eigenvector series against the oracle and perturb_norm! / perturb! consistency."""
    import wae_b200 as W
    from cases import G8_TAYLOR, G9_APPROX_1_HZ, G9_APPROX_20
    from oracle.helmholtz import discretize as odisc
    from oracle.mesh import Mesh as OMesh
    from oracle.nlevp import mslp as omslp
    from oracle.nlevp import perturb_fast_bang as ofast
    L = _gpu_family("lin", n=1.0)
    sol, n, flag = W.mslp(L, 150 * 2 * math.pi, maxiter=30, tol=1e-10, output=False)
    assert flag == 0
    W.perturb_fast_bang(sol, L, "τ", 20)
    tay = sol.eigval_pert["τ/Taylor"]
    for a, b in zip(tay, G8_TAYLOR):
        assert abs(a - b) <= 1e-5 * abs(b)
    assert abs(sol("τ", 0.0015, 20) - G9_APPROX_20) <= TOL * abs(G9_APPROX_20)
    assert abs(sol("τ", 0.0015, 1) / 2 / math.pi - G9_APPROX_1_HZ) <= TOL * abs(G9_APPROX_1_HZ)
    # Pade [10/10] from the same coefficients is closer to the exact eigenvalue at tau + 0.5 ms (G5) than the Taylor polynomial
    exact = 916.7036137579256 + 494.32932528479967j
    assert abs(sol("τ", 0.0015, 10, 10) - exact) < abs(sol("τ", 0.0015, 20) - exact)
    # oracle: same series (eigenvalue coefficients to 1e-9, eigenvector coefficients up to the phase of v0)
    mo = OMesh("Rijke_mm.msh", scale=0.001, raw=load_raw_mesh("rijke_mm"))
    Lo = odisc(mo, rijke_dscrp(1.0, 0.001), mo.generate_field(speedofsound))
    so, _, _ = omslp(Lo, 150 * 2 * math.pi, maxiter=30, tol=1e-11)
    ofast(so, Lo, "τ", 8)
    for k in range(9):
        assert abs(tay[k] - so.eigval_pert["τ/Taylor"][k]) <= 1e-8 * abs(tay[k])
    vg, vo = sol.v_pert["τ/Taylor"], so.v_pert["τ/Taylor"]
    ph = np.vdot(vo[0], vg[0])
    ph /= abs(ph)
    for k in range(4):
        assert np.abs(vg[k] - ph * vo[k]).max() <= 1e-6 * np.abs(vo[k]).max(), k
    sol2, _, _ = W.mslp(L, 150 * 2 * math.pi, maxiter=30, tol=1e-11, output=False)
    W.perturb_norm_bang(sol2, L, "τ", 6)
    sol3, _, _ = W.mslp(L, 150 * 2 * math.pi, maxiter=30, tol=1e-11, output=False)
    W.perturb_bang(sol3, L, "τ", 4)
    for k in range(5):
        assert abs(sol2.eigval_pert["τ/Taylor"][k] - tay[k]) <= 1e-8 * abs(tay[k])
        assert abs(sol3.eigval_pert["τ/Taylor"][k] - tay[k]) <= 1e-8 * abs(tay[k])
    # G10: the same at the G1 solution (householder, n = 0.01), notebook cells 5-14
    from cases import G10_APPROX_20
    L1 = _gpu_family("lin", n=0.01)
    s1, _, _ = W.householder(L1, 340 * 2 * math.pi, maxiter=20, tol=1e-11, output=False)
    W.perturb_fast_bang(s1, L1, "τ", 20)
    assert abs(s1("τ", 0.001 + 1e-5, 20) - G10_APPROX_20) <= TOL * abs(G10_APPROX_20)
    r = W.conv_radius(s1, "τ")
    assert len(r) == 20 and 1e-3 < r[-1] < 3e-3
