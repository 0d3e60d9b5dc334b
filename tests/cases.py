"""Shared test cases: the reference's tutorial set-ups (docs/src/tutorial_01_rijke_tube.md:60-175)."""
import math
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_raw_mesh(name):
    """Raw gmsh parse stored by tests/golden/make_mesh_fixtures.py -> (points, lines, tris, tets, domains)."""
    z = np.load(os.path.join(GOLDEN, name + "_mesh.npz"))
    domains = {}
    for i, (n, d) in enumerate(zip(z["dom_names"], z["dom_dims"])):
        domains[str(n)] = {"dimension": int(d), "simplices": [int(x) for x in z[f"dom_{i}"]]}
    return (z["points"], [list(map(int, r)) for r in z["lines"]], [list(map(int, r)) for r in z["tris"]],
            [list(map(int, r)) for r in z["tets"]], domains)


# gas data of the Rijke tutorial
GAMMA, RHO, TU, TB, P0, RGAS = 1.4, 1.225, 300.0, 1200.0, 101325.0, 287.05
AREA = math.pi * 0.025**2
Q02U0 = P0 * (TB / TU - 1) * AREA * GAMMA / (GAMMA - 1)
X_REF = [0.0, 0.0, -0.00101]
N_REF = [0.0, 0.0, 1.0]


def speedofsound(x, y, z):
    return math.sqrt(GAMMA * RGAS * TU) if z < 0.0 else math.sqrt(GAMMA * RGAS * TB)


def rijke_dscrp(n, tau):
    return {
        "Interior": ("interior", ()),
        "Outlet": ("admittance", ("Y", 1e15)),
        "Flame": ("flame", (GAMMA, RHO, Q02U0, X_REF, N_REF, "n", "τ", n, tau)),
    }


# High-order perturbation theory, docs/src/tutorial_04_perturbation_theory.md:112-155: Rijke P1, n = 1, tau = 1 ms, mslp to 1e-11, then
# perturb_fast!(sol, L, :tau, 20).  G8: the 21 Taylor coefficients of omega(tau) as printed (6 significant digits);
# G9: sol(:tau, tau + 0.0005, 20) and the first-order value (printed in Hz) at full precision.
G8_TAYLOR = [1075.33 + 372.102j, -2.62868e5 + 3.40796e5j, -1.79944e8 - 1.475e8j, 9.4741e10 - 1.57309e11j, 1.66943e14 + 8.14274e13j,
             -8.3483e16 + 1.86246e17j, -2.15622e20 - 9.17653e19j, 1.05357e23 - 2.58704e23j, 3.19354e26 + 1.25588e26j,
             -1.54315e29 + 4.02817e29j, -5.16822e32 - 1.94111e32j, 2.48748e35 - 6.72431e35j, 8.85193e38 + 3.23647e38j,
             -4.26474e41 + 1.17688e42j, -1.57803e45 - 5.68031e44j, 7.63541e47 - 2.13155e48j, 2.89779e51 + 1.0345e51j,
             -1.41133e54 + 3.9619e54j, -5.44414e57 - 1.93716e57j, 2.67322e60 - 7.51468e60j, 1.04148e64 + 3.70668e63j]
G9_APPROX_20 = 916.7085040155473 + 494.3258317478708j
G9_APPROX_1_HZ = 150.22496667319837 + 86.34150633981955j
# G10 (examples/tutorials/tutorial_04_perturbation_theory.ipynb cells 5-14): the same at the G1 set-up (n = 0.01, householder to 1e-11):
# perturb_fast!(sol, L, :tau, 20), then sol(:tau, tau + 1e-5, 20); first two coefficients as printed in cell 11.  (The later cells of that
# notebook -- conv_radius, the [10/10] Pade value -- were produced after L.params had been modified in place and are not reproducible.)
G10_APPROX_20 = 1710.8641999717368 + 9.593830019669932j
G10_TAYLOR_01 = (1710.7 + 9.61502j, 16655.8 - 1972.54j)


def tutorial_08_check(discretize, mslp, perturb_fast_bang, pade, polyval, mesh, c, **kw):
    """docs/src/tutorial_08_custom_FTF.md:44-203 with either implementation: (i) the n-tau flame written as a user closure FTF(ω, k) gives
    the eigenvalue of the built-in model (G4); (ii) the plain-FTF family H, solved passive, expanded to 16th order in the flame response
    and closed with the FTF by Newton-Raphson on the [8/8] Pade approximant, lands on the same eigenvalue ("works like a charm")."""
    import math

    import numpy as np
    n, tau = 1.0, 0.001
    FTF = lambda w, k=0: n * np.exp(-1j * w * tau) * (-1j * tau) ** k
    flame = (GAMMA, RHO, Q02U0, X_REF, N_REF)
    base = {"Interior": ("interior", ()), "Outlet": ("admittance", ("Y", 1e15))}
    g4 = 1075.325211506839 + 372.1017670372039j
    L = discretize(mesh, {**base, "Flame": ("flame", flame + (FTF,))}, c, **kw)
    sol, nn, flag = mslp(L, 340, maxiter=20, tol=1e-9, scale=2 * math.pi)
    assert flag == 0 and abs(sol.params["ω"] - g4) / abs(g4) < 1e-10
    H = discretize(mesh, {**base, "Flame": ("flame", flame)}, c, **kw)
    H.params["FTF"] = 0j
    sb, nn, flag = mslp(H, 340 * 2 * math.pi, maxiter=20, tol=1e-11)
    assert flag == 0 and abs(sb.params["ω"].real / 2 / math.pi - 272.064) < 1e-3
    perturb_fast_bang(sb, H, "FTF", 16)
    a, b = pade(sb.eigval_pert["FTF/Taylor"], 8, 8)
    d = lambda p: np.array([k * p[k] for k in range(1, len(p))])

    def w(z, k=0):
        P, Q = polyval(a, z), polyval(b, z)
        return P / Q if k == 0 else (polyval(d(a), z) * Q - P * polyval(d(b), z)) / Q**2
    eta = 0j
    for _ in range(10):
        om = w(eta)
        eta -= (FTF(om) - eta) / (FTF(om, 1) * w(eta, 1) - 1)
    assert abs(FTF(w(eta)) - eta) < 1e-12 and abs(w(eta) - g4) / abs(g4) < 1e-7
    return w(eta)
