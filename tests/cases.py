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
