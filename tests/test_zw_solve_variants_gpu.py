"""GPU check of the triangular-solve kernel generations against each other and against the matrix: the window-inverse kernels
(round 2, the default: csrc/lu_numeric.cu lu_wininv_kernel, lu_fwd_win / lu_bwd_win, the fused deep-level kernels, the 16-part forward
and 4-columns-per-warp backward updates) against the block-by-block substitution of round 1 (WAE_LU_WININV=0, WAE_LU_SOLVE_UPD2=0), on a
box whose top separators span several 256-column windows and whose deep levels are forced through the fused kernels
(WAE_LU_SOLVE_FUSED_MIN=1), for 1, 2, 3, 5 and 8 right-hand sides (kernel instantiations for 1, 2, 4 and 8; the staged backward kernels
of the 4- and 8-wide ones) and the three transposition modes."""
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KNOBS = ("WAE_LU_WININV", "WAE_LU_SOLVE_FUSED", "WAE_LU_SOLVE_FUSED_MIN", "WAE_LU_SOLVE_UPD2")


def _family(order):
    import wae_b200 as W
    mesh = W.kuhn_box((14, 14, 24), (0, 0, -0.25), (0.05, 0.05, 0.25), jitter=0.1, seed=7, flame_layer=(12, 13))
    c = mesh.generate_field(lambda x, y, z: 347.2 if z < 0 else 694.4)
    gam, rho = 1.4, 1.225
    q = 101325.0 * 3 * math.pi * 0.025**2 * gam / (gam - 1)
    dscrp = {"Interior": ("interior", ()), "Outlet": ("admittance", ("Y", 1e15)),
             "Flame": ("flame", (gam, rho, q, [0.025, 0.025, -0.6 * 0.5 / 24], [0, 0, 1.0], "n", "τ", 1.0, 0.001))}
    return W.discretize(mesh, dscrp, c, order=order)


@pytest.mark.parametrize("order", ["lin", "quad"])
def test_solve_kernel_generations_agree(order):
    L = _family(order)
    d = L.size()
    rng = np.random.default_rng(3)
    B = rng.standard_normal((d, 8)) + 1j * rng.standard_normal((d, 8))
    z = 340 * 2 * math.pi
    op = L(z)
    A = op.to_scipy().tocsr()
    dev = L.device()
    ctx = dev.ctx
    op.materialize(0)
    lid = dev.lu()
    old = {k: os.environ.pop(k, None) for k in KNOBS}
    settings = {
        "round1": {"WAE_LU_WININV": "0", "WAE_LU_SOLVE_UPD2": "0"},
        "default": {},
        "fused_everywhere": {"WAE_LU_SOLVE_FUSED_MIN": "1"},
        "no_fused": {"WAE_LU_SOLVE_FUSED": "0"},
        "inverse_only": {"WAE_LU_SOLVE_FUSED": "0", "WAE_LU_SOLVE_UPD2": "0"},
    }
    mats = {0: A, 1: A.T.tocsr(), 2: A.conj().T.tocsr()}
    try:
        results = {}
        for name, env in settings.items():
            for k in KNOBS:
                os.environ.pop(k, None)
            os.environ.update(env)
            ctx.lu_factor(lid, 0)  # the window inverses are written by the factorisation
            for nrhs in (1, 2, 3, 5, 8):
                for trans in (0, 1, 2):
                    X = ctx.lu_solve(lid, B[:, :nrhs] if nrhs > 1 else B[:, 0], trans=trans)
                    X = X.reshape(d, -1)
                    Aop = mats[trans]
                    Bn = B[:, :nrhs]
                    res = (np.abs(Aop @ X - Bn) / (abs(Aop) @ np.abs(X) + np.abs(Bn))).max()  # componentwise backward error
                    assert res < 1e-12, (name, order, nrhs, trans, res)
                    results[(name, nrhs, trans)] = X
        for (name, nrhs, trans), X in results.items():
            ref = results[("round1", nrhs, trans)]
            assert np.abs(X - ref).max() <= 1e-9 * np.abs(ref).max(), (name, order, nrhs, trans)
    finally:
        for k in KNOBS:
            os.environ.pop(k, None)
            if old[k] is not None:
                os.environ[k] = old[k]
        ctx.lu_factor(lid, 0)


def test_householder_with_and_without_window_inverses():
    """The Newton iteration drives L(omega) to singularity: the inverse of the last window holds entries ~ 1 / (last pivot).  The eigenvalue
    must not depend on the kernel generation (1e-10, the parity tolerance of the path)."""
    import wae_b200 as W
    old = {k: os.environ.pop(k, None) for k in KNOBS}
    try:
        om = {}
        for name, env in (("round1", {"WAE_LU_WININV": "0", "WAE_LU_SOLVE_UPD2": "0"}), ("default", {}), ("fused", {"WAE_LU_SOLVE_FUSED_MIN": "1"})):
            for k in KNOBS:
                os.environ.pop(k, None)
            os.environ.update(env)
            L = _family("quad")
            sol, n, flag = W.householder(L, 340 * 2 * math.pi, maxiter=15, tol=1e-9 * 340 * 2 * math.pi, output=False)
            assert flag >= 0, (name, flag)
            om[name] = sol.params["ω"]
            L.release()
        for name, w in om.items():
            assert abs(w - om["round1"]) <= 1e-10 * abs(w), (name, w, om["round1"])
    finally:
        for k in KNOBS:
            os.environ.pop(k, None)
            if old[k] is not None:
                os.environ[k] = old[k]
