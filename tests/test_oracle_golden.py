"""Pin the CPU oracle to the reference's stored outputs."""
import math

import numpy as np
import pytest

from cases import load_raw_mesh, rijke_dscrp, speedofsound
from oracle import fem
from oracle.helmholtz import discretize
from oracle.mesh import Mesh
from oracle.nlevp import beyn, householder, mslp

TOL = 1e-10  # relative eigenvalue tolerance of BASELINE.json


@pytest.fixture(scope="module")
def rijke():
    mesh = Mesh("Rijke_mm.msh", scale=0.001, raw=load_raw_mesh("rijke_mm"))
    c = mesh.generate_field(speedofsound)
    return mesh, c


def test_mesh_counts(rijke):
    mesh, c = rijke
    # docs/src/tutorial_01_rijke_tube.md:62-64
    assert mesh.points.shape == (3, 1006) and len(mesh.triangles) == 1562 and len(mesh.tetrahedra) == 3380


def test_fem_tables_match_reference_expressions():
    g = np.load(__import__("os").path.join(__import__("cases").GOLDEN, "fem_tables.npz"))
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    T = fem.tables
    assert rel(T(1, 4)["mass"], g["tab_s43v1u1"]) < 1e-15 and rel(T(2, 4)["mass"], g["tab_s43v2u2"]) < 1e-15
    assert rel(T(1, 3)["mass"], g["tab_s33v1u1"]) < 1e-15 and rel(T(2, 3)["mass"], g["tab_s33v2u2"]) < 1e-15
    assert rel(T(1, 4)["src"], g["tab_s43v1"][0]) < 1e-15 and rel(T(2, 4)["src"], g["tab_s43v2"][0]) < 1e-15
    assert rel(T(1, 3)["src"], g["tab_s33v1"][0]) < 1e-15 and rel(T(2, 3)["src"], g["tab_s33v2"][0]) < 1e-15
    for k in range(g["X"].shape[0]):
        ct, c = fem.CooTrafo(g["X"][k]), g["C4"][k]
        for o, n1, n2, n3 in [(1, "s43nv1nu1", "s43nv1nu1cc1", "s43v1u1c1"), (2, "s43nv2nu2", "s43nv2nu2cc1", "s43v2u2c1")]:
            assert rel(fem.tet_stiff(ct, o), g["val_" + n1][k]) < 5e-15
            assert rel(fem.tet_stiff_cc1(ct, c, o), g["val_" + n2][k]) < 5e-15
            assert rel(fem.tet_mass_c1(ct, c, o), g["val_" + n3][k]) < 5e-15
        assert rel(fem.tet_grad_at(ct, g["nref"][k], g["xref"][k], 2), g["val_s43nv2rx"][k]) < 5e-15
        assert rel(fem.tet_grad_at(ct, g["nref"][k], g["xref"][k], 1), g["val_s43nv1rx"][k]) < 5e-15
        ct, c = fem.CooTrafo(g["XT"][k]), g["C3"][k]
        assert rel(fem.tri_mass_c1(ct, c, 1), g["val_s33v1u1c1"][k]) < 5e-15
        assert rel(fem.tri_mass_c1(ct, c, 2), g["val_s33v2u2c1"][k]) < 5e-15
        assert rel(fem.tri_src_c1(ct, c, 1), g["val_s33v1c1"][k]) < 5e-15
        assert rel(fem.tri_src_c1(ct, c, 2), g["val_s33v2c1"][k]) < 5e-15


G_HOUSEHOLDER = [  # (tau, omega) notebook cells 5, 16, 32
    (0.001, 1710.6977772393461 + 9.615018460173488j),
    (0.001 + 0.00001, 1710.864199971756 + 9.593830019670127j),
    (0.001 + 0.0008798274754933992 + 0.001, 1707.4565281774599 - 9.11764397194075j),
]
G1_TRACE_RES = [1.6243873211323859e6, 147442.983718491, 1804.6396540359215, 0.2839458865676277]
G1_TRACE_Z = [2136.2830044410593, 1753.640443755553 + 8.160610349893785j, 1711.2293969397867 + 9.58445933911067j,
              1710.6978605746078 + 9.615009671129867j, 1710.6977772393602 + 9.615018460179712j]


@pytest.mark.parametrize("tau,omega", G_HOUSEHOLDER)
def test_householder_goldens(rijke, tau, omega):
    mesh, c = rijke
    L = discretize(mesh, rijke_dscrp(0.01, tau), c)
    assert L.size() == 1006
    trace = []
    sol, n, flag = householder(L, 340 * 2 * math.pi, maxiter=20, tol=1e-11, trace=trace)
    assert abs(sol.params["ω"] - omega) / abs(omega) < TOL
    if tau == 0.001:
        for k, r in enumerate(G1_TRACE_RES):
            assert abs(trace[k + 1][1] - r) / r < 1e-6
        for k, z in enumerate(G1_TRACE_Z):
            assert abs(trace[k][3] - z) / abs(z) < 1e-9


G_MSLP = [  # (tau, start, omega) docs/src/tutorial_04_perturbation_theory.md:75-88,147-172,386-408
    (0.001, 340 * 2 * math.pi, 1075.325211506839 + 372.1017670372039j),
    (0.0015, 916.7085040155473 + 494.3258317478708j, 916.7036137579256 + 494.32932528479967j),
    (0.001 + 2 * 0.0007029896606802446, 668.5373997804821 + 529.4636751544649j, 668.537399929804 + 529.4636746814361j),
]


@pytest.mark.parametrize("tau,start,omega", G_MSLP)
def test_mslp_goldens(rijke, tau, start, omega):
    mesh, c = rijke
    L = discretize(mesh, rijke_dscrp(1, tau), c)
    sol, n, flag = mslp(L, start, maxiter=20, tol=1e-11)
    assert flag == 0
    assert abs(sol.params["ω"] - omega) / abs(omega) < TOL
    if tau == 0.001:
        assert n in (7, 8, 9)  # reference: 8; the last step is at round-off level (|dz|~1e-12)


def test_beyn_prose_values(rijke):
    """docs/src/tutorial_01_rijke_tube.md:202-212: two modes near 272 and 695 Hz (N reduced from 256 to 32)."""
    mesh, c = rijke
    L = discretize(mesh, rijke_dscrp(0.0, 0.001), c)
    G = [z * 2 * math.pi for z in (150 + 5j, 150 - 5j, 1000 - 5j, 1000 + 5j)]
    Om, P = beyn(L, G, l=5, N=32)
    f = np.sort(Om.real / 2 / math.pi)
    assert len(f) == 2 and abs(f[0] - 272) < 1.0 and abs(f[1] - 695) < 5.0
    # the reference's own idiom (tutorial_06...jl:41-55): polish with a local solver
    for om in Om:
        sol, n, flag = householder(L, om, maxiter=10, tol=1e-10)
        assert abs(sol.params["ω"] - om) < 5e-3 * abs(om)  # N=32 on a thin rectangle: quadrature-limited


def test_tutorial_01_local_solver_prose(rijke):
    """docs/src/tutorial_01_rijke_tube.md:216-269: `mslp(L, 250*2*pi)` without a stopping criterion runs its ten iterations and returns
    flag 1 ("the maximum number of iterations has been reached") at the 272 Hz mode; after `L.params[:n] = 1` the third-order call
    `mslp(L, 245*2*pi - 82im*2*pi, order=3)` lands on the active mode whose growth rate is "≈ 59.22" -- the same eigenvalue as G4."""
    mesh, c = rijke
    L = discretize(mesh, rijke_dscrp(0.0, 0.001), c)
    sol, n, flag = mslp(L, 250 * 2 * math.pi)
    assert (n, flag) == (10, 1) and abs(sol.params["ω"].real / 2 / math.pi - 272) < 0.5 and abs(sol.params["ω"].imag) < 1e-6
    L.params["n"] = 1
    sol, n, flag = mslp(L, (245 - 82j) * 2 * math.pi, order=3)
    w = sol.params["ω"]
    assert round(abs(w.imag) / 2 / math.pi, 2) == 59.22
    g4 = 1075.325211506839 + 372.1017670372039j
    assert abs(w - g4) / abs(g4) < TOL


def test_tutorial_08_custom_ftf_and_flame_response_closure(rijke):
    """docs/src/tutorial_08_custom_FTF.md: custom FTF closure == built-in n-tau model (G4); flame-response parametrisation + [8/8] Pade +
    Newton-Raphson on the scalar closure reproduces it to 1e-8."""
    from cases import tutorial_08_check
    from oracle.nlevp import pade, perturb_fast_bang, polyval
    mesh, c = rijke
    quiet = lambda L, z, **k: mslp(L, z, **k)
    tutorial_08_check(discretize, quiet, perturb_fast_bang, pade, polyval, mesh, c)


def test_perturb_fast_goldens(rijke):
    """G8/G9 (docs/src/tutorial_04_perturbation_theory.md:112-155): 20th-order power series of omega(tau) by perturb_fast!, its value
    at tau + 0.5 ms and the first-order value; perturb! (the slow algorithm) and perturb_norm! give the same eigenvalue series."""
    from cases import G8_TAYLOR, G9_APPROX_1_HZ, G9_APPROX_20
    from oracle.nlevp import perturb_bang, perturb_fast_bang, perturb_norm_bang, solution_eval
    mesh, c = rijke
    L = discretize(mesh, rijke_dscrp(1, 0.001), c)
    sol, n, flag = mslp(L, 150 * 2 * math.pi, maxiter=30, tol=1e-11)
    assert flag == 0
    mode0, active0 = L.mode, list(L.active)
    perturb_fast_bang(sol, L, "τ", 20)
    assert L.mode == mode0 and L.active == active0  # perturb_fast! restores the family (LinOpFam.jl:586-588)
    tay = sol.eigval_pert["τ/Taylor"]
    for a, b in zip(tay, G8_TAYLOR):
        assert abs(a - b) <= 1e-5 * abs(b)  # printed with 6 significant digits
    assert abs(solution_eval(sol, "τ", 0.0015, 20, 0) - G9_APPROX_20) <= 1e-11 * abs(G9_APPROX_20)
    assert abs(solution_eval(sol, "τ", 0.0015, 1, 0) / 2 / math.pi - G9_APPROX_1_HZ) <= 1e-11 * abs(G9_APPROX_1_HZ)
    sol2, _, _ = mslp(L, 150 * 2 * math.pi, maxiter=30, tol=1e-11)
    perturb_bang(sol2, L, "τ", 6)
    sol3, _, _ = mslp(L, 150 * 2 * math.pi, maxiter=30, tol=1e-11)
    perturb_norm_bang(sol3, L, "τ", 6)
    for k in range(7):
        assert abs(sol2.eigval_pert["τ/Taylor"][k] - tay[k]) <= 1e-9 * abs(tay[k])
        assert abs(sol3.eigval_pert["τ/Taylor"][k] - tay[k]) <= 1e-9 * abs(tay[k])


def test_perturb_fast_golden_at_g1(rijke):
    """G10 (notebook cells 5-14): 20th-order series at the G1 solution, evaluated 1e-5 s away; equals the exact eigenvalue G2 to 1e-11."""
    from cases import G10_APPROX_20, G10_TAYLOR_01
    from oracle.nlevp import perturb_fast_bang, solution_eval
    mesh, c = rijke
    L = discretize(mesh, rijke_dscrp(0.01, 0.001), c)
    sol, n, flag = householder(L, 340 * 2 * math.pi, maxiter=20, tol=1e-11)
    perturb_fast_bang(sol, L, "τ", 20)
    tay = sol.eigval_pert["τ/Taylor"]
    for a, b in zip(tay, G10_TAYLOR_01):
        assert abs(a - b) <= 1e-5 * abs(b)
    w = solution_eval(sol, "τ", 0.001 + 1e-5, 20, 0)
    assert abs(w - G10_APPROX_20) <= 1e-11 * abs(w)
    assert abs(w - G_HOUSEHOLDER[1][1]) <= 1e-10 * abs(w)  # G2: the exact eigenvalue at tau = 1.01 ms


def _qep1():
    """Quadratic eigenvalue problem 1 of the NLEVP collection as in docs/src/tutorial_00_NLEVP.md:28-43."""
    import scipy.sparse as sp
    from oracle.nlevp import LinearOperatorFamily, Term, pow1, pow2
    A2 = np.array([[0, 6, 0], [0, 6, 0], [0, 0, 1]], dtype=complex)
    A1 = np.array([[1, -6, 0], [2, -7, 0], [0, 0, 0]], dtype=complex)
    T = LinearOperatorFamily()
    T.push(Term(sp.csc_matrix(A2), (pow2,), (("λ",),), "λ^2", "A2"))
    T.push(Term(sp.csc_matrix(A1), (pow1,), (("λ",),), "λ", "A1"))
    T.push(Term(sp.identity(3, dtype=complex, format="csc"), (), (), "", "A0"))
    return T


def test_generic_nlevp_tutorial_00():
    """tutorial_00_NLEVP.md:144 (mslp(T, 0): eigenvalue 1/3, ten iterations, warning flag 1 = maxiter) and :282-286 (Beyn on the
    square +-2 +-2i with l = 6 finds the 5 eigenvalues inside: 1/3, 1/2, 1, +-i).  A generic family gets the auxiliary term
    -I * __aux__ (iterative_solvers.jl:119-123)."""
    T = _qep1()
    sol, n, flag = mslp(T, 0, maxiter=10)
    assert (n, flag) == (10, 1) and abs(sol.params["λ"] - 1 / 3) < 1e-12
    assert T.terms[-1].operator == "__aux__" and T.auxval == "__aux__"
    Om, P = beyn(_qep1(), [2 + 2j, -2 + 2j, -2 - 2j, 2 - 2j], l=6, N=64)
    exact = np.array([1 / 3, 0.5, 1.0, 1j, -1j])
    assert len(Om) == 5
    for z in exact:
        assert np.abs(Om - z).min() < 1e-9
    for i, lam in enumerate(Om):  # residual check of the tutorial
        v = P[:, i] / np.linalg.norm(P[:, i])
        assert np.linalg.norm(_qep1()(lam) @ v) < 1e-8


def test_gallery_rijke_tube_against_the_analytic_duct():
    """Gallery.rijke_tube (src/NLEVP/gallery.jl:171-260), the assembly-free fixture of SURVEY 8(c).  Passive (n = 0) it is a closed-open
    duct with a jump of the speed of sound, whose eigenfrequencies solve  c1 tan(w x_m / c1) = c2 cot(w (l - x_m) / c2);  first-order
    elements converge to them like h^2.  Active (n = 1, tau = 2) householder and mslp must agree, and Beyn must find the same mode."""
    import scipy.optimize as so
    from oracle.gallery import rijke_tube
    from oracle.nlevp import beyn, householder, mslp
    errs = []
    for res in (129, 257):
        L, grid = rijke_tube(res)
        assert L.size() == res and [t.operator for t in L.terms] == ["M", "K", "C", "Q", "__aux__"]
        L.params["n"] = 0j
        mid = res // 2 + 1
        xm, l, c1, c2 = grid[mid - 1], grid[-1], 1.0, 2.0  # elements 1..mid-1 (1-based) carry c_min: the jump sits at node mid-1 (0-based)
        f = lambda w: c1 * math.tan(w * xm / c1) - c2 / math.tan(w * (l - xm) / c2)
        exact = so.brentq(f, 1.0, 2.5)
        sol, n, flag = householder(L, exact * 1.05, maxiter=20, tol=1e-10)
        assert flag == 1 and abs(sol.params["ω"].imag) < 1e-9
        errs.append(abs(sol.params["ω"].real - exact) / exact)
    assert errs[0] < 2e-4 and 3.5 < errs[0] / errs[1] < 4.5  # second-order convergence to the analytic eigenfrequency
    L, _ = rijke_tube(129)
    sh, nh, fh = householder(L, 1.7, maxiter=30, tol=1e-10)
    L2, _ = rijke_tube(129)
    sm, nm, fm = mslp(L2, 1.7, maxiter=30, tol=1e-10)
    assert fh == 1 and fm == 0 and abs(sh.params["ω"] - sm.params["ω"]) < 1e-9
    w = sh.params["ω"]
    assert abs(w.imag) > 1e-3  # the flame makes the mode grow or decay
    L3, _ = rijke_tube(129)
    Om, P = beyn(L3, [w + 0.3 + 0.3j, w - 0.3 + 0.3j, w - 0.3 - 0.3j, w + 0.3 - 0.3j], l=4, N=32)
    assert len(Om) >= 1 and np.abs(Om - w).min() < 1e-8
