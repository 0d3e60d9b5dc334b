"""Oracle (test infrastructure): NLEVP operator family and solvers on scipy sparse.

Restates, for the hot path only:
  scalar algebra         src/NLEVP/algebra.jl:4-43 (pow0/1/2), :129-150 (exp_az, exp_delay)
  Term / family / push!  src/NLEVP/LinOpFam.jl:16-35, :131-186, :305-346
  L(z...) evaluation     src/NLEVP/LinOpFam.jl:466-529
  perturb! / perturb     src/NLEVP/LinOpFam.jl:546-560, src/NLEVP/perturbation.jl:2-121,319-367
  pade / polyval         src/NLEVP/LinOpFam.jl:622-642, :715-730
  householder            src/NLEVP/Householder.jl:21-35, :70-203
  mslp                   src/NLEVP/iterative_solvers.jl:93-252
  beyn / gauss / wn      src/NLEVP/beyn.jl:34-138, :146-209

Third-party arithmetic of the reference (not vendored there): Arpack.jl 0.4.0
``eigs(A,M;nev,sigma=0,v0)`` -> scipy.sparse.linalg.eigs (same ARPACK mode 3);
UMFPACK ``lu``/``\\`` -> scipy SuperLU; FastGaussQuadrature.gausslegendre ->
numpy.polynomial.legendre.leggauss.
"""
import math

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

# ----------------------------------------------------------------------------- algebra


def pow0(z, k=0):
    return complex(1) if k == 0 else (complex(0) if k > 0 else complex("nan"))


def pow1(z, k=0):
    if k == 0:
        return complex(z)
    if k == 1:
        return complex(1)
    return complex(0) if k > 1 else complex("nan")


def pow2(z, k=0):
    if k == 0:
        return complex(z) ** 2
    if k == 1:
        return 2 * complex(z)
    if k == 2:
        return complex(2)
    return complex(0) if k > 2 else complex("nan")


def _pow(z, k, a):
    """algebra.jl:46-61: k-th derivative of z^a for integer a."""
    if k > a > 0:
        return 0.0
    f = 1.0
    i = a
    for _ in range(k):
        f *= i
        i -= 1
    return f * complex(z) ** (a - k)


def exp_delay(w, tau, m, n):
    """algebra.jl:138-147: d^m/dw^m d^n/dtau^n exp(-i w tau)."""
    a = -1j
    f = 0.0
    for i in range(n + 1):
        f += math.comb(n, i) * _pow(tau, i, m) * (a * w) ** (n - i)
    return f * a**m * np.exp(a * w * tau)


def exp_az(z, a, k):
    return a**k * np.exp(a * z)


# ----------------------------------------------------------------------------- family


def generate_z_g_z(g):  # algebra.jl:169-179
    def z_g_z(z, n):
        return z * g(z, 0) if n == 0 else z * g(z, n) + n * g(z, n - 1)
    return z_g_z


def generate_stsp_z(A, B, C, D):  # algebra.jl:158-167
    A = np.atleast_2d(np.asarray(A, dtype=complex))
    B = np.asarray(B, dtype=complex).reshape(A.shape[0], 1)
    C = np.asarray(C, dtype=complex).reshape(1, A.shape[0])
    D = complex(np.asarray(D).ravel()[0])

    def stsp_z(z, n):
        inv = np.linalg.inv(1j * z * np.eye(A.shape[0]) - A)
        P = np.eye(A.shape[0], dtype=complex)
        for _ in range(n + 1):
            P = P @ inv
        f = (-1j) ** n * math.factorial(n) * (C @ P @ B)[0, 0]
        return f + D if n == 0 else f
    return stsp_z


def exp_ax2(z, a, n):  # algebra.jl:229-253, with the running A /= a, Z /= z^2 of the reference
    if a == 0:
        return complex(1) if n == 0 else complex(0)
    f = 0j
    A, Z = np.complex128(a) ** n, np.complex128(z) ** n
    cnst = 2**n * math.factorial(n)
    with np.errstate(all="ignore"):  # IEEE semantics as in Julia: at z = 0 the running quotient becomes Inf/NaN instead of raising
        for k in range(n // 2 + 1):
            f += cnst * 4.0 ** (-k) / math.factorial(k) / math.factorial(n - 2 * k) * A * Z
            A = A / np.complex128(a)
            Z = Z / np.complex128(z) ** 2
    return complex(f * np.exp(a * z**2))


def exp_az2mzit(z, tau, a, m, n, k):  # algebra.jl:255-274
    coeff = 0j
    for ii in range(m + 1):
        for jj in range(m - ii + 1):
            kk = m - jj - ii
            multi = math.factorial(m) / math.factorial(ii) / math.factorial(jj) / math.factorial(kk)
            coeff += multi * _pow(z, kk, n + 2 * k) * exp_ax2(z, a, jj) * exp_delay(z, tau, ii, 0)
    return coeff * (-1j) ** n


def sigma_nexp_az2mzit(*args):  # algebra.jl:313-325 (0-based: z, then J triples (n, tau, a), then m, then J triples (l, n, k))
    J = (len(args) - 2) // 6
    z, m = args[0], args[3 * J + 1]
    f = 0j
    for j in range(J):
        nn, tau, a = args[1 + 3 * j], args[2 + 3 * j], args[3 + 3 * j]
        l, n, k = args[3 * J + 2 + 3 * j], args[3 * J + 3 + 3 * j], args[3 * J + 4 + 3 * j]
        f += pow1(nn, l) * exp_az2mzit(z, tau, a, m, n, k)
    return f


class Term:
    def __init__(self, coeff, func, params, symbol, operator):
        self.coeff, self.func, self.params, self.symbol, self.operator = coeff, tuple(func), tuple(params), symbol, operator
        self.varlist = []
        for par in params:
            for v in par:
                if v not in self.varlist:
                    self.varlist.append(v)

    def scalar(self, d):
        """LinOpFam.jl:466-479 without the matrix: product of func(args..., derivs...)."""
        c = complex(1)
        for f, pars in zip(self.func, self.params):
            c *= f(*[d[p][0] for p in pars], *[d[p][1] for p in pars])
        return c


class Solution:
    def __init__(self, params, v, v_adj, eigval, auxval=""):
        self.params = dict(params)
        self.v, self.v_adj, self.eigval, self.auxval = v, v_adj, eigval, auxval
        self.eigval_pert, self.v_pert = {}, {}


class LinearOperatorFamily:
    def __init__(self, params=("λ",), values=None):
        if values is None:
            values = [complex("nan")] * len(params)
        self.terms = []
        self.eigval = params[0]
        self.auxval = params[-1] if len(params) > 1 else ""
        self.active = [self.eigval]
        self.params = {p: complex(v) for p, v in zip(params, values)}
        self.mode = "all"

    def push(self, T):
        """LinOpFam.jl:305-346."""
        for idx, t in enumerate(self.terms):
            if (t.func, t.params) == (T.func, T.params):
                coeff = t.coeff + T.coeff
                if abs(coeff).sum() == 0:
                    del self.terms[idx]
                else:
                    self.terms[idx] = Term(coeff, t.func, t.params, t.symbol, t.operator)
                return
        for pars in T.params:
            for p in pars:
                self.params.setdefault(p, complex("nan"))
        self.terms.append(T)

    def size(self):
        return self.terms[0].coeff.shape[0] if self.terms else 0

    def scalars(self, derivs):
        """Per-term scalar (or None if skipped) for derivative orders `derivs` of self.active."""
        dd = dict(zip(self.active, derivs))
        out = []
        for t in self.terms:
            if self.mode != "householder" and t.operator == "__aux__":
                out.append(None)
                continue
            if any(d > 0 and v not in t.varlist for v, d in zip(self.active, derivs)):
                out.append(None)
                continue
            d = {v: (self.params[v], dd.get(v, 0)) for v in t.varlist}
            out.append(t.scalar(d))
        return out

    def __call__(self, *args):
        """LinOpFam.jl:482-529 (oplist/in_or_ex not restated: unused on the hot path)."""
        na = len(self.active)
        if self.mode == "all":
            for v, val in zip(self.active, args):
                self.params[v] = complex(val)
        if self.mode == "all" and len(args) == na:
            derivs = [0] * na
        else:
            derivs = [int(a) for a in args[len(args) - na :]]
        coeff = sp.csc_matrix(self.terms[0].coeff.shape, dtype=complex)  # spzeros(size(L.terms[1].coeff)...): matrices or vectors
        for t, s in zip(self.terms, self.scalars(derivs)):
            if s is not None:
                coeff = coeff + s * t.coeff
        if self.mode in ("compact", "householder"):
            coeff = coeff / math.prod(math.factorial(int(a)) for a in args[len(args) - na :])
        return coeff


# ----------------------------------------------------------------------------- perturbation


def partitions(n):
    """All partitions of n as ascending lists (Kelleher's accelerated ascending rule, perturbation.jl:2-80)."""
    a = [0] * (n + 1)
    k = 1
    y = n - 1
    while k != 0:
        x = a[k - 1] + 1
        k -= 1
        while 2 * x <= y:
            a[k] = x
            y -= x
            k += 1
        l = k + 1
        while x <= y:
            a[k] = x
            a[l] = y
            yield a[: k + 2]
            x += 1
            y -= 1
        a[k] = x + y
        y = x + y - 1
        yield a[: k + 1]


def part2mult(p):
    mu = [0] * sum(p)
    for i in p:
        mu[i - 1] += 1
    return mu


def multinomcoeff(mu):
    return math.factorial(sum(mu)) / math.prod(math.factorial(m) for m in mu)


class _LstsqSolver:
    def __init__(self, A):
        self.A = sp.csc_matrix(A).toarray()

    def solve(self, b):
        return np.linalg.lstsq(self.A, b, rcond=None)[0]


def _factor_singular(A):
    """lu(L(0,0), check=false) with the QR fallback of perturbation.jl:329-332: L(0,0) is singular by construction; SuperLU raises on
    an exactly singular pivot (tiny dense examples) where UMFPACK reports a status -- then a least-squares solve stands in for SPQR
    (the null-space component of the solution is projected out right after, perturbation.jl:361)."""
    try:
        return spla.splu(sp.csc_matrix(A))
    except RuntimeError:
        return _LstsqSolver(A)


def perturb(L, N, v0, v0Adj):
    """perturbation.jl:319-367: Taylor coefficients of the auxiliary eigenvalue."""
    v0 = v0 / np.sqrt(np.vdot(v0, v0))
    L10 = L(1, 0)
    v0Adj = v0Adj / (np.vdot(v0Adj, L10 @ v0))
    lam = np.zeros(N + 1, dtype=complex)
    v = [None] * (N + 1)
    v[0] = v0
    L00 = _factor_singular(L(0, 0))
    for k in range(1, N + 1):
        r = np.zeros(len(v0), dtype=complex)
        for n in range(1, k + 1):
            r += L(0, n) @ v[k - n]
        for m in range(1, k + 1):
            for p in partitions(m):
                if p == [k]:
                    continue
                mu = part2mult(p)
                for n in range(0, k - m + 1):
                    coeff = 1
                    for g, mg in enumerate(mu):
                        coeff *= lam[g + 1] ** mg
                    r += (L(sum(mu), n) @ v[k - n - m]) * multinomcoeff(mu) * coeff
        lam[k] = -np.vdot(v0Adj, r) / np.vdot(v0Adj, L10 @ v0)
        v[k] = L00.solve(-(r + lam[k] * (L10 @ v0)))
        v[k] = v[k] - np.vdot(v0, v[k]) * v0
    return lam, v


def perturb_disk(L, N, v0, v0Adj):
    """perturbation.jl:373-450: the same power series as perturb(), summed per derivative pair (m, n) -- r = sum L(m,n) w_mn with
    w_mn = sum over the multi-indices mu with |mu| = m of v[k-n-weight(mu)] * multinomial(mu) * prod lambda_g^mu_g -- and with the
    eigenvector normalisation c = -1/2 sum_l v_l' v_(k-l).  The reference reads the multi-indices from files shipped with the package
    (compressed_perturbation_data/k/m_n); here they are generated (all partitions of every weight W <= k - n except {k})."""
    v0 = v0 / np.sqrt(np.vdot(v0, v0))
    L10v0 = L(1, 0) @ v0
    v0Adj = v0Adj / np.vdot(v0Adj, L10v0)
    den = np.vdot(v0Adj, L10v0)
    lam = np.zeros(N + 1, dtype=complex)
    v = [None] * (N + 1)
    v[0] = v0
    L00 = spla.splu(sp.csc_matrix(L(0, 0)))
    for k in range(1, N + 1):
        w = {}
        for n in range(1, k + 1):
            w[(0, n)] = v[k - n].copy()
        for W in range(1, k + 1):
            for p in partitions(W):
                if p == [k]:
                    continue
                mu = part2mult(p)
                coeff = multinomcoeff(mu)
                for g, mg in enumerate(mu):
                    coeff = coeff * lam[g + 1] ** mg
                for n in range(0, k - W + 1):
                    key = (len(p), n)
                    if key in w:
                        w[key] = w[key] + v[k - n - W] * coeff
                    else:
                        w[key] = v[k - n - W] * coeff
        r = np.zeros(len(v0), dtype=complex)
        for (m, n), wv in sorted(w.items()):
            r += L(m, n) @ wv
        lam[k] = -np.vdot(v0Adj, r) / den
        v[k] = L00.solve(-(r + lam[k] * L10v0))
        v[k] = v[k] - np.vdot(v0, v[k]) * v0
        c = 0j
        for l in range(1, k):
            c -= 0.5 * np.vdot(v[l], v[k - l])
        v[k] = v[k] + c * v[0]
    return lam, v


def perturb_norm(L, N, v0, v0Adj):
    """perturbation.jl:470-545: perturb_disk with the mass matrix Y = -coeff(__aux__) in every inner product (literal restatement,
    including the lu(Y) solve of v0Adj)."""
    Y = sp.csc_matrix(-L.terms[-1].coeff).astype(complex)
    v0 = v0 / np.sqrt(np.vdot(v0, Y @ v0))
    v0Adj = spla.splu(Y).solve(v0Adj)
    L10v0 = L(1, 0) @ v0
    v0Adj = v0Adj / np.vdot(v0Adj, Y @ L10v0)
    lam = np.zeros(N + 1, dtype=complex)
    v = [None] * (N + 1)
    v[0] = v0
    L00 = spla.splu(sp.csc_matrix(L(0, 0)))
    for k in range(1, N + 1):
        w = {}
        for n in range(1, k + 1):
            w[(0, n)] = v[k - n].copy()
        for W in range(1, k + 1):
            for p in partitions(W):
                if p == [k]:
                    continue
                mu = part2mult(p)
                coeff = multinomcoeff(mu)
                for g, mg in enumerate(mu):
                    coeff = coeff * lam[g + 1] ** mg
                for n in range(0, k - W + 1):
                    key = (len(p), n)
                    w[key] = w[key] + v[k - n - W] * coeff if key in w else v[k - n - W] * coeff
        r = np.zeros(len(v0), dtype=complex)
        for (m, n), wv in sorted(w.items()):
            r += L(m, n) @ wv
        lam[k] = -np.vdot(v0Adj, Y @ r) / np.vdot(v0Adj, Y @ L10v0)
        v[k] = L00.solve(-(r + lam[k] * L10v0))
        v[k] = v[k] - np.vdot(v0, Y @ v[k]) * v0
        c = 0j
        for l in range(1, k):
            c -= 0.5 * np.vdot(v[l], Y @ v[k - l])
        v[k] = v[k] + c * v[0]
    return lam, v


def perturb_norm_bang(sol, L, param, N, mode="compact"):
    """perturb_norm! (LinOpFam.jl:606-620)."""
    active, params, cur = L.active, L.params, L.mode
    L.params = sol.params
    L.active = [sol.eigval, param]
    L.mode = mode
    key = f"{param}/Taylor"
    try:
        sol.eigval_pert[key], sol.v_pert[key] = perturb_norm(L, N, sol.v, sol.v_adj)
        sol.eigval_pert[key][0] = sol.params[sol.eigval]
    finally:
        L.active, L.mode, L.params = active, cur, params


def perturb_fast_bang(sol, L, param, N, mode="compact"):
    """perturb_fast! (LinOpFam.jl:576-590)."""
    active, params, cur = L.active, L.params, L.mode
    L.params = sol.params
    L.active = [sol.eigval, param]
    L.mode = mode
    key = f"{param}/Taylor"
    try:
        sol.eigval_pert[key], sol.v_pert[key] = perturb_disk(L, N, sol.v, sol.v_adj)
        sol.eigval_pert[key][0] = sol.params[sol.eigval]
    finally:
        L.active, L.mode, L.params = active, cur, params


def solution_eval(sol, param, eps, Lo=0, M=0):
    """(sol::Solution)(param, eps, L, M) (LinOpFam.jl:680-696): [L/M] Pade approximant of the eigenvalue at param = eps."""
    key = f"{param}/[{Lo}/{M}]"
    if key not in sol.eigval_pert:
        sol.eigval_pert[key] = pade(sol.eigval_pert[f"{param}/Taylor"], Lo, M)
    a, b = sol.eigval_pert[key]
    d = eps - sol.params[param]
    return polyval(a, d) / polyval(b, d)


def perturb_bang(sol, L, param, N, mode="compact"):
    """LinOpFam.jl:546-560."""
    active, params, cur = L.active, L.params, L.mode
    L.params = sol.params
    L.active = [sol.eigval, param]
    L.mode = mode
    key = f"{param}/Taylor"
    try:
        sol.eigval_pert[key], sol.v_pert[key] = perturb(L, N, sol.v, sol.v_adj)
        sol.eigval_pert[key][0] = sol.params[sol.eigval]
    finally:
        L.active, L.mode, L.params = active, cur, params


def pade(w, Lo, M):
    """LinOpFam.jl:622-642."""
    A = np.zeros((M, M), dtype=complex)
    for i in range(1, M + 1):
        for j in range(1, M + 1):
            if Lo + i - j >= 0:
                A[i - 1, j - 1] = w[Lo + i - j]
    b = np.linalg.solve(A, -np.asarray(w[Lo + 1 : Lo + M + 1])) if M > 0 else np.zeros(0, dtype=complex)
    b = np.concatenate([[1.0], b])
    a = np.zeros(Lo + 1, dtype=complex)
    for l in range(Lo + 1):
        for m in range(l + 1):
            if m <= M:
                a[l] += w[l - m] * b[m]
    return a, b


def polyval(p, z):
    f = p[-1]
    for c in p[-2::-1]:
        f = f * z + c
    return f


def poly_roots(p):
    """Householder.jl:195-203: companion-matrix eigenvalues."""
    N = len(p) - 1
    C = np.zeros((N, N), dtype=complex)
    for i in range(1, N):
        C[i, i - 1] = 1
    C[:, N - 1] = -np.asarray(p[:N]) / p[N]
    return np.linalg.eigvals(C)


def householder_update(f):
    """Householder.jl:21-35."""
    o = len(f) - 1
    if o == 1:
        return -f[0] / f[1]
    if o == 2:
        return -f[0] * f[1] / (f[1] ** 2 - 0.5 * f[0] * f[2])
    if o == 3:
        return -(6 * f[0] * f[1] ** 2 - 3 * f[0] ** 2 * f[2]) / (6 * f[1] ** 3 - 6 * f[0] * f[1] * f[2] + f[0] ** 2 * f[3])
    if o == 4:
        return -(4 * f[0] * (6 * f[1] ** 3 - 6 * f[0] * f[1] * f[2] + f[0] ** 2 * f[3])) / (
            24 * f[1] ** 4 - 36 * f[0] * f[1] ** 2 * f[2] + 6 * f[0] ** 2 * f[2] ** 2 + 8 * f[0] ** 2 * f[1] * f[3] - f[0] ** 3 * f[4])
    return (5 * f[0] * (24 * f[1] ** 4 - 36 * f[0] * f[1] ** 2 * f[2] + 6 * f[0] ** 2 * f[2] ** 2 + 8 * f[0] ** 2 * f[1] * f[3] - f[0] ** 3 * f[4])) / (
        -120 * f[1] ** 5 + 240 * f[0] * f[1] ** 3 * f[2] - 60 * f[0] ** 2 * f[1] ** 2 * f[3]
        + 10 * f[0] ** 2 * f[1] * (-9 * f[2] ** 2 + f[0] * f[4]) + f[0] ** 3 * (20 * f[2] * f[3] - f[0] * f[5]))


# ----------------------------------------------------------------------------- local solvers


def _eigs(A, M, nev, v0):
    """Arpack.eigs(A, M; nev, sigma=0, v0) (Householder.jl:100-101, iterative_solvers.jl:132-133).  For the tiny dense examples of
    tutorial_00 (dimension 3) scipy's ARPACK wrapper cannot build a Krylov space of the default ncv = 20 and fails where Arpack.jl
    succeeds; there the generalised eigenproblem is solved densely and the nev eigenvalues of smallest modulus are returned."""
    if A.shape[0] <= 20:
        import scipy.linalg as sla
        lam, v = sla.eig(sp.csc_matrix(A).toarray(), sp.csc_matrix(M).toarray())
        lam = np.where(np.isfinite(lam), lam, np.inf)
        idx = np.argsort(np.abs(lam), kind="stable")[:nev]
        lam_s, v_s = lam[idx], v[:, idx].astype(complex)
        # a degenerate eigenvalue: the Krylov method returns the component of the start vector in the eigenspace, not an arbitrary
        # basis vector of it (matters for T(0) = I of tutorial_00, where all three auxiliary eigenvalues coincide)
        for i in range(len(idx)):
            cluster = np.flatnonzero(np.abs(lam - lam_s[i]) <= 1e-8 * max(1.0, abs(lam_s[i])))
            if len(cluster) > 1 and v0 is not None:
                Vc = v[:, cluster]
                proj = Vc @ np.linalg.lstsq(Vc, np.asarray(v0, dtype=complex), rcond=None)[0]
                if np.linalg.norm(proj) > 1e-12 * np.linalg.norm(v0):
                    v_s[:, i] = proj / np.linalg.norm(proj)
        return lam_s, v_s
    lam, v = spla.eigs(sp.csc_matrix(A), k=nev, M=sp.csc_matrix(M), sigma=0, v0=v0, tol=0)
    idx = np.argsort(np.abs(lam), kind="stable")
    return lam[idx], v[:, idx]


def _iterate(L, z, maxiter, tol, relax, order, nev, v0, v0_adj, kind, num_order=1, trace=None):
    """Shared skeleton of householder (Householder.jl:70-192) and mslp (iterative_solvers.jl:93-252)."""
    z = complex(z)
    z0 = complex("inf")
    lam = float("inf")
    lam0 = float("inf")
    n = 0
    active, mode = L.active, L.mode
    d = L.size()
    if v0 is None:
        L(0)
        v0 = np.ones(d, dtype=complex)
    if v0_adj is None:
        v0_adj = np.conj(v0)
    M = -L.terms[-1].coeff
    while abs(z - z0) > tol and n < maxiter:
        if trace is not None:
            trace.append((n, abs(lam), abs(z - z0), z))
        if kind == "householder":
            z0 = z
        L.params[L.eigval] = z
        L.params[L.auxval] = 0
        A = L(z)
        lams, v = _eigs(A, M, nev, v0)
        lams_adj, v_adj = _eigs(A.conj().T, M.conj().T, nev, v0_adj)
        dzs, back = [], []
        L.active = [L.auxval, L.eigval]
        for i in range(nev):
            L.params[L.auxval] = lams[i]
            sol = Solution(L.params, v[:, i], v_adj[:, i], L.auxval)
            perturb_bang(sol, L, L.eigval, order, mode="householder")
            coeffs = sol.eigval_pert[f"{L.eigval}/Taylor"]
            if kind == "householder":
                f = [math.factorial(k) * c for k, c in enumerate(coeffs)]
                dzs.append(householder_update(f))
            else:
                num, den = pade(coeffs, num_order, order - num_order)
                roots = poly_roots(num)
                dzs.append(roots[np.argsort(np.abs(roots), kind="stable")[0]])
                if z0 != complex("inf"):
                    back.append(lam0 - polyval(num, z0 - z) / polyval(den, z0 - z))
        L.active = [L.eigval]
        sel = np.argsort(np.abs(back if back else dzs), kind="stable")[0]
        lam = lams[sel]
        L.params[L.auxval] = lam
        if kind == "mslp":
            z0 = z
            lam0 = lam
        z = z + relax * dzs[sel]
        v0 = (1 - relax) * v0 + relax * v[:, sel]
        v0_adj = (1 - relax) * v0_adj + relax * v_adj[:, sel]
        n += 1
    L.params[L.eigval] = z
    if trace is not None:
        trace.append((n, abs(lam), abs(z - z0), z))
    L.active, L.mode = active, mode
    v0 = v0 / np.sqrt(np.vdot(v0, M @ v0))
    v0_adj = v0_adj / np.conj(np.vdot(v0_adj, L(L.params[L.eigval], 1) @ v0))
    return Solution(L.params, v0, v0_adj, L.eigval), n, z, z0, lam


def householder(L, z, maxiter=10, tol=0.0, relax=1.0, lam_tol=float("inf"), order=1, nev=1, v0=None, v0_adj=None, trace=None):
    sol, n, z, z0, lam = _iterate(L, z, maxiter, tol, relax, order, nev, v0, v0_adj, "householder", trace=trace)
    if n >= maxiter:
        flag = -1
    elif abs(lam) <= lam_tol:
        flag = 1
    elif abs(z - z0) <= tol:
        flag = 0
    else:
        flag = -3
    return sol, n, flag


def mslp(L, z, maxiter=10, tol=0.0, relax=1.0, lam_tol=float("inf"), order=1, nev=1, v0=None, v0_adj=None, num_order=1, scale=1, trace=None):
    if L.terms[-1].operator != "__aux__":  # iterative_solvers.jl:119-123: a generic family gets the auxiliary term -I * __aux__
        L.push(Term(-sp.identity(L.size(), dtype=complex, format="csc"), (pow1,), (("__aux__",),), "__aux__", "__aux__"))
        L.auxval = "__aux__"
    sol, n, z, z0, lam = _iterate(L, z * scale, maxiter, tol * scale, relax, order, nev, v0, v0_adj, "mslp", num_order, trace=trace)
    if n >= maxiter:
        flag = 1
    elif abs(lam) <= lam_tol:
        flag = 0
    elif abs(z - z0) <= tol * scale:
        flag = 2
    else:
        flag = -1
    return sol, n, flag


# ----------------------------------------------------------------------------- Beyn


def wn(z, G):
    """beyn.jl:185-209 winding number."""
    w = 0
    for i in range(len(G)):
        a, b = G[i], G[(i + 1) % len(G)]
        isleft = (b.real - a.real) * (z.imag - a.imag) - (z.real - a.real) * (b.imag - a.imag)
        if a.imag <= z.imag:
            if b.imag > z.imag and isleft > 0:
                w += 1
        elif b.imag <= z.imag and isleft < 0:
            w -= 1
    return w


def contour_nodes(G, N):
    """beyn.jl:112-138: Gauss-Legendre nodes/weights on each polygon edge (weights include (b-a)/2)."""
    X, W = np.polynomial.legendre.leggauss(N)
    zs, ws = [], []
    for i in range(len(G)):
        a, b = complex(G[i]), complex(G[(i + 1) % len(G)])
        zs.extend(X * (b - a) / 2 + (a + b) / 2)
        ws.extend(W * (b - a) / 2)
    return np.array(zs), np.array(ws)


def beyn_moments(L, G, l, K, N, nodes=None):
    d = L.size()
    L(0)
    V = np.zeros((d, l), dtype=complex)
    for i in range(min(l, d)):
        V[i, i] = 1
    zs, ws = contour_nodes(G, N)
    A = np.zeros((d, l, 2 * K), dtype=complex)
    sel = range(len(zs)) if nodes is None else nodes
    for j in sel:
        X = spla.splu(sp.csc_matrix(L(zs[j]))).solve(V) * ws[j]
        for p in range(2 * K):
            A[:, :, p] += zs[j] ** p * X
    return A


def moments2eigs(A, G, l, K, tol=0.0, pos_test=True):
    d = A.shape[0]
    B0 = np.zeros((d * K, l * K), dtype=complex)
    B1 = np.zeros((d * K, l * K), dtype=complex)
    for i in range(K):
        for j in range(K):
            B0[d * i : d * (i + 1), l * j : l * (j + 1)] = A[:, :, i + j]
            B1[d * i : d * (i + 1), l * j : l * (j + 1)] = A[:, :, i + j + 1]
    V, S, Wh = np.linalg.svd(B0, full_matrices=False)
    W = Wh.conj().T
    if tol > 0:
        m = S > tol
        V, S, W = V[:, m], S[m], W[:, m]
    Om, P = np.linalg.eig(V.conj().T @ B1 @ W @ np.diag(1 / S))
    P = V[:d, :] @ P
    if pos_test:
        m = np.array([wn(complex(z), G) != 0 for z in Om], dtype=bool)
        Om, P = Om[m], P[:, m]
    return Om, P, S


def beyn(L, G, l=5, K=1, N=16, tol=0.0, pos_test=True):
    """beyn.jl:34-110."""
    d = L.size()
    K = max(K, l // d + int(l % d != 0))
    A = beyn_moments(L, G, l, K, N)
    Om, P, _ = moments2eigs(A, G, l, K, tol, pos_test)
    return Om, P
