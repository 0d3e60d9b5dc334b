"""GPU check of what the factorisation reports when its static pivoting cannot stand in for UMFPACK's row exchanges (the reference's
lu / factorize raise SingularException: perturbation.jl:329 passes check = false on purpose, every other call site relies on the
exception): an exactly zero pivot raises WAE_E_SINGULAR, a perturbed pivot that matters is caught by the probe solve after the
factorisation, lu_factor(..., check=False) accepts both, and a regular matrix of the same shape goes through untouched."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def _factor(A, check=True, pivot_eps=None):
    from wae_b200 import _lib
    A = sp.csc_matrix(A.astype(complex))
    A.sort_indices()
    old = os.environ.pop("WAE_LU_PIVOT_EPS", None)
    if pivot_eps is not None:
        os.environ["WAE_LU_PIVOT_EPS"] = repr(pivot_eps)
    ctx = _lib.Context(0)
    try:
        n = A.shape[0]
        pid, mid = ctx.mat_set(n, A.indptr, A.indices, A.data)
        fid, _ = ctx.family_create([mid])
        ctx.combine(fid, np.array([1.0], dtype=complex), 0)
        lid, _, _ = ctx.lu_analyze(fid)  # reads WAE_LU_PIVOT_EPS
        ctx.lu_factor(lid, 0, check=check)
        b = np.arange(1, n + 1, dtype=complex)
        return ctx.lu_solve(lid, b), ctx.last_ms("static_pivots"), ctx.last_ms("zero_pivots")
    finally:
        ctx.close()
        os.environ.pop("WAE_LU_PIVOT_EPS", None)
        if old is not None:
            os.environ["WAE_LU_PIVOT_EPS"] = old


def _matrix(a22):
    n = 40
    A = sp.lil_matrix((n, n))
    A.setdiag(np.linspace(2.0, 3.0, n))
    for i in range(n - 1):  # tridiagonal coupling keeps the pattern connected
        A[i, i + 1] = A[i + 1, i] = 0.1
    A[0, 0], A[0, 1], A[1, 0], A[1, 1] = 1.0, 2.0, 2.0, a22
    A[1, 2] = A[2, 1] = 0.0
    return A.tocsc()


def test_regular_matrix_goes_through():
    A = _matrix(7.0)
    x, perturbed, zero = _factor(A)
    b = np.arange(1, A.shape[0] + 1, dtype=complex)
    assert perturbed == 0 and zero == 0
    assert np.abs(A @ x - b).max() <= 1e-12 * np.abs(b).max()


def test_exactly_zero_pivot_raises_unless_check_is_off():
    from wae_b200 import _lib
    A = _matrix(4.0)  # second pivot: 4 - 2 * 2 = 0 exactly
    with pytest.raises(_lib.SingularException):
        _factor(A)
    x, perturbed, zero = _factor(A, check=False)  # lu(A, check = false): factorised with a perturbed pivot, no exception
    assert zero == 1 and perturbed >= 1


def test_perturbed_pivot_that_matters_is_caught_by_the_probe():
    from wae_b200 import _lib
    A = _matrix(4.0 + 1e-12)  # second pivot 1e-12 (2.5e-13 after equilibration): below a threshold of 1e-8 it is replaced
    with pytest.raises(_lib.SingularException):
        _factor(A, pivot_eps=1e-8)
    x, perturbed, zero = _factor(A, check=False, pivot_eps=1e-8)
    assert perturbed >= 1 and zero == 0
    # with the default threshold (1e-30) the pivot is kept and the solve is as accurate as the conditioning allows
    x, perturbed, zero = _factor(A)
    b = np.arange(1, A.shape[0] + 1, dtype=complex)
    assert perturbed == 0
    assert np.abs(A @ x - b).max() <= 1e-6 * np.abs(x).max() * abs(A).max()
