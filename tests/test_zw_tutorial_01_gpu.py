"""GPU pin of tutorial_01's third-order local solve (docs/src/tutorial_01_rijke_tube.md:258-269).  Added after the round's GPU budget
was spent: the file sorts late so that `pytest -x` reaches it after the tests that have already run on a B200."""
import math

import pytest

from test_lu_gpu import TOL, _gpu_family

pytestmark = pytest.mark.gpu


def test_tutorial_01_third_order_mslp_gpu():
    """docs/src/tutorial_01_rijke_tube.md:258-269: mslp(L, 245*2*pi - 82im*2*pi, order=3) at n = 1 -> growth rate "≈ 59.22", the G4
    eigenvalue (a stopping criterion is given here: without one the iteration keeps factorising the converged, singular L(ω))."""
    import wae_b200 as W
    L = _gpu_family("lin", n=1.0, tau=0.001)
    sol, n, flag = W.mslp(L, (245 - 82j) * 2 * math.pi, order=3, maxiter=15, tol=1e-10, output=False)
    g4 = 1075.325211506839 + 372.1017670372039j
    assert flag == 0 and abs(sol.params["ω"] - g4) / abs(g4) < TOL
    assert round(abs(sol.params["ω"].imag) / 2 / math.pi, 2) == 59.22


def test_tutorial_08_custom_ftf_and_flame_response_closure_gpu():
    """docs/src/tutorial_08_custom_FTF.md on the GPU: custom FTF closure == G4; passive solution of the plain-FTF family, 16th-order
    perturb_fast! in the flame response, [8/8] Pade + Newton-Raphson closure -> G4 to 1e-7."""
    import wae_b200 as W
    from cases import load_raw_mesh, speedofsound, tutorial_08_check
    mesh = W.Mesh("m", scale=0.001, raw=load_raw_mesh("rijke_mm"))
    tutorial_08_check(W.discretize, lambda L, z, **k: W.mslp(L, z, output=False, **k), W.perturb_fast_bang, W.pade, W.polyval, mesh,
                      mesh.generate_field(speedofsound))
