"""Host mirror of the reference's NLEVP module for the accelerated path.

Same names, argument meaning, side effects and flags as the reference (Julia symbols become
Python strings, ``push!`` is ``L.push``), with every matrix living on the GPU behind the C ABI
(include/wae_b200.h):

  pow0/pow1/pow2, exp_delay, exp_az ...   src/NLEVP/algebra.jl
  Term, Solution, LinearOperatorFamily    src/NLEVP/LinOpFam.jl:16-35, 95-112, 131-186, 305-346
  L(z), L(z,k), L(m,n)                    src/NLEVP/LinOpFam.jl:482-529  -> one wae_combine pass
  perturb!, perturb                       src/NLEVP/LinOpFam.jl:546-560, src/NLEVP/perturbation.jl:319-367
  householder                             src/NLEVP/Householder.jl:70-192
  mslp                                    src/NLEVP/iterative_solvers.jl:93-252
  beyn, inpoly, wn                        src/NLEVP/beyn.jl:34-138, 178-209

Scalar coefficient functions stay on the host (they are arbitrary closures in the reference);
only their values cross the ABI.
"""
import math
import os
import time

import numpy as np

from . import _lib

# --------------------------------------------------------------------------------------------- algebra
_NAN = complex("nan")


def pow0(z, k=0):
    """algebra.jl:4-12"""
    return complex(1) if k == 0 else (complex(0) if k > 0 else _NAN)


def pow1(z, k=0):
    """algebra.jl:16-26"""
    if k == 0:
        return complex(z)
    if k == 1:
        return complex(1)
    return complex(0) if k > 1 else _NAN


def pow2(z, k=0):
    """algebra.jl:30-42"""
    if k == 0:
        return complex(z) ** 2
    if k == 1:
        return 2 * complex(z)
    if k == 2:
        return complex(2)
    return complex(0) if k > 2 else _NAN


def pow(z, k, a):
    """k-th derivative of z^a (algebra.jl:46-75)."""
    if isinstance(a, int) and k > a > 0:
        return complex(0)
    if k < 0:
        return _NAN
    f, i = 1.0, a
    for _ in range(k):
        f *= i
        i -= 1
    return f * complex(z) ** (a - k)


def pow_a(a):
    """algebra.jl:77-107"""
    return lambda z, k=0: pow(z, k, a)


def exp_az(z, a, k):
    """algebra.jl:129-135"""
    return a**k * np.exp(a * z) if k >= 0 else _NAN


def exp_delay(w, tau, m, n):
    """d^m/dw^m d^n/dtau^n exp(-i w tau)  (algebra.jl:138-147)."""
    a = -1j
    f = 0.0
    for i in range(n + 1):
        f += math.comb(n, i) * pow(tau, i, m) * (a * w) ** (n - i)
    return f * a**m * np.exp(a * w * tau)


tau_delay = exp_delay


def generate_z_g_z(g):
    """algebra.jl:169-179: derivative rule for z*g(z)."""
    def z_g_z(z, n):
        return z * g(z, 0) if n == 0 else z * g(z, n) + n * g(z, n - 1)
    return z_g_z


def generate_stsp_z(A, B, C, D):
    """algebra.jl:158-167: n-th derivative of the state-space transfer function C (i z I - A)^-1 B + D."""
    A, B, C = np.atleast_2d(np.asarray(A, dtype=complex)), np.asarray(B, dtype=complex).reshape(-1, 1), np.asarray(C, dtype=complex).reshape(1, -1)

    def stsp_z(z, n):
        R = np.linalg.matrix_power(np.linalg.inv(1j * z * np.eye(A.shape[0]) - A), n + 1)
        f = (-1j) ** n * math.factorial(n) * (C @ R @ B)[0, 0]
        return f + (np.asarray(D, dtype=complex).ravel()[0] if n == 0 else 0)
    return stsp_z


def exp_ax2(z, a, n):
    """algebra.jl:229-253: n-th derivative of exp(a z^2) with respect to z."""
    if a == 0:
        return complex(1) if n == 0 else complex(0)
    f = 0j
    for k in range(n // 2 + 1):
        f += 2**n * math.factorial(n) * 4.0 ** (-k) / math.factorial(k) / math.factorial(n - 2 * k) * complex(a) ** (n - k) * complex(z) ** (n - 2 * k)
    return f * np.exp(a * z * z)


def exp_az2mzit(z, tau, a, m, n, k):
    """algebra.jl:255-274: d^m/dz^m d^n/dtau^n d^k/da^k exp(a z^2 - i z tau) (Gaussian-filtered time delay of :fancyflame)."""
    f = pow_a(n + 2 * k)
    coeff = 0j
    for ii in range(m + 1):
        h_ii = exp_delay(z, tau, ii, 0)
        for jj in range(m - ii + 1):
            kk = m - jj - ii
            multi = math.factorial(m) / math.factorial(ii) / math.factorial(jj) / math.factorial(kk)
            coeff += multi * f(z, kk) * exp_ax2(z, a, jj) * h_ii
    return coeff * (-1j) ** n


def Sigma_nexp_az2mzit(*args):
    """algebra.jl:313-325 (the reference's Σnexp_az2mzit): sum_j n_j exp(a_j z^2 - i z tau_j); args = (z, n_1, tau_1, a_1, ..., m, l_1, n_1, k_1, ...)."""
    J = (len(args) - 2) // 6
    z, m = args[0], args[3 * J + 1]
    f = 0j
    for j in range(J):
        nn, tau, a = args[1 + 3 * j: 4 + 3 * j]
        l, n, k = args[2 + 3 * J + 3 * j: 5 + 3 * J + 3 * j]
        f += pow1(nn, l) * exp_az2mzit(z, tau, a, m, n, k)
    return f


def generate_Sigma_y_exp_ikx(y):
    """algebra.jl:276-288"""
    N = len(y)

    def f(z, n):
        s = 0j
        for k, yk in enumerate(y):
            s += (k**n if n else 1) * yk * np.exp(2j * math.pi * k / N * z)
        return s * (2j * math.pi / N) ** n
    return f


def generate_gz_hz(g, h):
    """algebra.jl:290-299"""
    return lambda z, k: sum(math.comb(k, i) * h(z, k - i) * g(z, i) for i in range(k + 1))


def generate_1_gz(g):
    """algebra.jl:301-310"""
    return lambda z, k: (1 - g(z, k)) if k == 0 else -g(z, k)


# --------------------------------------------------------------------------------------------- device objects
_CTX = None


def get_context(device=None):
    """Process-wide libwae_b200 context (one GPU per process; LOCAL_RANK picks the device)."""
    global _CTX
    if _CTX is None:
        import os
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        _CTX = _lib.Context(device)
    return _CTX


def reset_context():
    global _CTX
    if _CTX is not None:
        _CTX.close()
    _CTX = None


class DeviceMatrix:
    """A sparse matrix resident on the GPU (what ``Term.coeff`` is in this package).

    ``parts`` is a list of (mat_id, scale) whose sum is the matrix -- ``push`` merges same-signature terms
    without a device-side sparse add (LinOpFam.jl:311-335 does ``coeff + T.coeff``)."""

    def __init__(self, ctx, dim, parts):
        self.ctx, self.dim, self.parts = ctx, dim, list(parts)

    @property
    def shape(self):
        return (self.dim, self.dim)

    def __neg__(self):
        return DeviceMatrix(self.ctx, self.dim, [(m, -s) for m, s in self.parts])

    def __add__(self, other):
        return DeviceMatrix(self.ctx, self.dim, self.parts + other.parts)

    def csc(self):
        """(colptr, rowval, nzval) of the (first) part, 0-based: the SparseMatrixCSC fields."""
        mid, s = self.parts[0]
        pid, _, nnz = self.ctx.mat_info(mid)
        colptr, rowval = self.ctx.pattern_get(pid, self.dim, nnz)
        return colptr, rowval, self.ctx.mat_get(mid) * s

    def to_scipy(self):
        import scipy.sparse as sp
        out = None
        for mid, s in self.parts:
            pid, _, nnz = self.ctx.mat_info(mid)
            colptr, rowval = self.ctx.pattern_get(pid, self.dim, nnz)
            A = sp.csc_matrix((self.ctx.mat_get(mid) * s, rowval, colptr), shape=self.shape)
            out = A if out is None else out + A
        return out

    @staticmethod
    def from_scipy(A, ctx=None):
        import scipy.sparse as sp
        ctx = ctx or get_context()
        A = sp.csc_matrix(A, dtype=complex)
        A.sort_indices()
        A.sum_duplicates()
        _, mid = ctx.mat_set(A.shape[0], A.indptr, A.indices, A.data)
        return DeviceMatrix(ctx, A.shape[0], [(mid, 1.0)])


class Term:
    """LinOpFam.jl:16-35"""

    def __init__(self, coeff, func, params, symbol, operator):
        self.coeff, self.func, self.params = coeff, tuple(func), tuple(tuple(p) for p in params)
        self.symbol, self.operator = symbol, operator
        self.varlist = []
        for par in self.params:
            for v in par:
                if v not in self.varlist:
                    self.varlist.append(v)

    def scalar(self, d):
        """LinOpFam.jl:466-479 without the matrix: prod_k func_k(args..., derivs...)."""
        c = complex(1)
        for f, pars in zip(self.func, self.params):
            c *= f(*[d[p][0] for p in pars], *[d[p][1] for p in pars])
        return c

    def __str__(self):
        return (self.symbol + "*" if self.symbol else "") + self.operator


class Solution:
    """LinOpFam.jl:95-112"""

    def __init__(self, params, v, v_adj, eigval, auxval=""):
        self.params = dict(params)
        self.v, self.v_adj, self.eigval, self.auxval = v, v_adj, eigval, auxval
        self.eigval_pert, self.v_pert = {}, {}

    def __call__(self, param, eps, Lo=0, M=0, vector=False):
        """(sol::Solution)(param, eps, L, M; vector) (LinOpFam.jl:680-696): evaluate the [Lo/M] Pade approximant (M = 0: Taylor
        polynomial of order Lo) of the eigenvalue (and eigenvector) at param = eps from the stored perturbation coefficients."""
        key = f"{param}/[{Lo}/{M}]"
        if key not in self.eigval_pert or (vector and key not in self.v_pert):
            pade_bang(self, param, Lo, M, vector=vector)
        a, b = self.eigval_pert[key]
        d = eps - self.params[param]
        val = polyval(a, d) / polyval(b, d)
        if not vector:
            return val
        A, B = self.v_pert[key]
        return val, polyval(A, d) / polyval(B, d)

    def __str__(self):
        txt = f"####Solution####\neigval:\n{self.eigval} = {self.params[self.eigval]}\n\nParameters:\n"
        for k, v in self.params.items():
            if k not in (self.eigval, self.auxval):
                txt += f"{k} = {v}\n"
        if self.auxval in self.params:
            txt += f"\nResidual:\nabs({self.auxval}) = {abs(self.params[self.auxval])}\n"
        return txt


class DeviceOperator:
    """Value of ``L(z...)``: the family evaluated with fixed term scalars.  Materialised into one of the
    family's device value slots on demand (a single pass over the shared pattern)."""

    def __init__(self, fam, coeffs):
        self.fam, self.coeffs = fam, np.asarray(coeffs, dtype=np.complex128)

    @property
    def shape(self):
        return (self.fam.dim, self.fam.dim)

    def materialize(self, slot=0):
        self.fam.combine(self.coeffs, slot)
        return slot

    def to_scipy(self, slot=0):
        import scipy.sparse as sp
        self.materialize(slot)
        colptr, rowval = self.fam.pattern()
        return sp.csc_matrix((self.fam.ctx.family_get(self.fam.fid, slot, self.fam.nnz), rowval, colptr), shape=self.shape)

    def matvec(self, x, trans=0, slot=0):
        self.materialize(slot)
        return self.fam.ctx.spmm(self.fam.fid, slot, x, trans)

    def __matmul__(self, x):
        return self.matvec(x)

    def solve(self, b, trans=0, slot=0):
        """``L(z) \\ b`` (UMFPACK in the reference)."""
        self.materialize(slot)
        lu = self.fam.lu()
        self.fam.ctx.lu_factor(lu, slot)
        return self.fam.ctx.lu_solve(lu, b, trans)


class _DeviceFamily:
    """Device side of a LinearOperatorFamily: union pattern + term maps (wae_family_create)."""

    def __init__(self, ctx, terms):
        self.ctx = ctx
        self.dim = terms[0].coeff.dim
        self.term_parts = []  # per term: list of (position in the flattened matrix list, scale)
        mats = []
        for t in terms:
            idx = []
            for mid, s in t.coeff.parts:
                idx.append((len(mats), s))
                mats.append(mid)
            self.term_parts.append(idx)
        self.n_flat = len(mats)
        self.fid, self.nnz = ctx.family_create(mats)
        self._pattern = None
        self._lu = None
        self.lu_nnz = self.lu_flops = None

    def flat(self, term_scalars):
        c = np.zeros(self.n_flat, dtype=np.complex128)
        for parts, s in zip(self.term_parts, term_scalars):
            if s is None:
                continue
            for pos, sc in parts:
                c[pos] = s * sc
        return c

    def combine(self, flat_coeffs, slot):
        self.ctx.combine(self.fid, flat_coeffs, slot)

    def pattern(self):
        if self._pattern is None:
            self._pattern = self.ctx.family_pattern_get(self.fid, self.dim, self.nnz)
        return self._pattern

    def lu(self):
        if self._lu is None:
            self._lu, self.lu_nnz, self.lu_flops = self.ctx.lu_analyze(self.fid)
        return self._lu

    def free(self):
        """Return the LU factor storage, value slots and term maps to the device (wae_lu_free, wae_family_free); the term matrices stay."""
        if self._lu is not None:
            self.ctx.lu_free(self._lu)
            self._lu = None
        if self.fid is not None:
            self.ctx.family_free(self.fid)
            self.fid = None

    def __del__(self):  # a dropped family returns its device storage (the reference relies on Julia's GC for the same objects)
        try:
            if getattr(self.ctx, "h", None):
                self.free()
        except Exception:  # noqa: BLE001 -- interpreter shutdown, context already destroyed
            pass


class LinearOperatorFamily:
    """LinOpFam.jl:131-186.  First parameter = eigenvalue, last = auxiliary eigenvalue."""

    def __init__(self, params=("λ",), values=None):
        if values is None:
            values = [_NAN] * len(params)
        self.terms = []
        self.eigval = params[0]
        self.auxval = params[-1] if len(params) > 1 else ""
        self.active = [self.eigval]
        self.params = {p: complex(v) for p, v in zip(params, values)}
        self.mode = "all"
        self._dev = None

    # -- push! (LinOpFam.jl:305-346) -------------------------------------------------------------
    def push(self, T):
        self.release()  # the device side (value slots, term maps, LU factors: 23 GB at config 2) belongs to the old term list
        for idx, t in enumerate(self.terms):
            if (t.func, t.params) == (T.func, T.params):
                coeff = t.coeff + T.coeff
                if coeff.dim <= 200000 and abs(coeff.to_scipy()).sum() == 0:
                    del self.terms[idx]
                else:
                    self.terms[idx] = Term(coeff, t.func, t.params, t.symbol, t.operator)
                return self
        for pars in T.params:
            for p in pars:
                self.params.setdefault(p, _NAN)
        self.terms.append(T)
        return self

    def __iadd__(self, T):
        return self.push(T)

    def size(self):
        return self.terms[0].coeff.dim if self.terms else 0

    def __str__(self):
        d = self.size()
        eq = "+".join(str(t) for t in self.terms if not t.operator.startswith("_"))
        pars = "".join(f"{k}\t{v}\n" for k, v in self.params.items())
        return f"{d}×{d}-dimensional operator family: \n\n{eq}\n\nParameters\n----------\n{pars}"

    # -- device side -----------------------------------------------------------------------------
    def device(self):
        if self._dev is None:
            if not self.terms:
                raise ValueError("empty operator family")
            self._dev = _DeviceFamily(self.terms[0].coeff.ctx, self.terms)
        return self._dev

    def release(self):
        """Explicitly drop the device side of the family (union pattern values, LU factors: what Julia's GC does for the reference's
        SparseMatrixCSC sums and UMFPACK factors); it is rebuilt on the next use.  The term matrices are untouched."""
        if self._dev is not None:
            self._dev.free()
            self._dev = None

    def scalars(self, derivs):
        """Per-term scalar (None if the term is skipped) -- LinOpFam.jl:501-522 without the matrices."""
        dd = dict(zip(self.active, derivs))
        out = []
        for t in self.terms:
            if self.mode != "householder" and t.operator == "__aux__":
                out.append(None)
                continue
            if any(d > 0 and v not in t.varlist for v, d in zip(self.active, derivs)):
                out.append(None)
                continue
            out.append(t.scalar({v: (self.params[v], dd.get(v, 0)) for v in t.varlist}))
        return out

    def __call__(self, *args):
        """LinOpFam.jl:482-529 (incl. the side effect on params in mode :all)."""
        na = len(self.active)
        if self.mode == "all":
            for v, val in zip(self.active, args):
                self.params[v] = complex(val)
        if self.mode == "all" and len(args) == na:
            derivs = [0] * na
        else:
            derivs = [int(a) for a in args[len(args) - na:]]
        sc = self.scalars(derivs)
        if self.mode in ("compact", "householder"):
            fac = math.prod(math.factorial(int(a)) for a in args[len(args) - na:])
            sc = [None if s is None else s / fac for s in sc]
        dev = self.device()
        return DeviceOperator(dev, dev.flat(sc))


class VectorFamily(LinearOperatorFamily):
    """The ``rhs`` of ``discretize(...; source=true)`` (Helmholtz.jl:79, 524-526): a LinearOperatorFamily whose coefficients are
    vectors.  The coefficients are boundary-sized and live on the host as dense complex arrays (assembled on the device by
    wae_assemble_wallsrc); calling the family returns ``Array(rhs(ω))``, ready for ``L(ω).solve(...)``."""

    def push(self, T):
        for idx, t in enumerate(self.terms):
            if (t.func, t.params) == (T.func, T.params):
                coeff = t.coeff + T.coeff
                if not np.any(coeff):
                    del self.terms[idx]
                else:
                    self.terms[idx] = Term(coeff, t.func, t.params, t.symbol, t.operator)
                return self
        for pars in T.params:
            for p in pars:
                self.params.setdefault(p, _NAN)
        self.terms.append(T)
        return self

    def size(self):
        return len(self.terms[0].coeff) if self.terms else 0

    def device(self):
        raise TypeError("a vector family has no device operator")

    def __call__(self, *args):
        na = len(self.active)
        if self.mode == "all":
            for v, val in zip(self.active, args):
                self.params[v] = complex(val)
        derivs = [0] * na if (self.mode == "all" and len(args) == na) else [int(a) for a in args[len(args) - na:]]
        out = np.zeros(self.size(), dtype=np.complex128)
        for t, s in zip(self.terms, self.scalars(derivs)):
            if s is not None:
                out += s * t.coeff
        if self.mode in ("compact", "householder"):
            out /= math.prod(math.factorial(int(a)) for a in args[len(args) - na:])
        return out


# --------------------------------------------------------------------------------------------- perturbation
def _partitions(n):
    a = [0] * (n + 1)
    k, y = 1, n - 1
    while k != 0:
        x = a[k - 1] + 1
        k -= 1
        while 2 * x <= y:
            a[k] = x
            y -= x
            k += 1
        l = k + 1
        while x <= y:
            a[k], a[l] = x, y
            yield a[: k + 2]
            x += 1
            y -= 1
        a[k] = x + y
        y = x + y - 1
        yield a[: k + 1]


def _part2mult(p):
    mu = [0] * sum(p)
    for i in p:
        mu[i - 1] += 1
    return mu


def perturb(L, N, v0, v0Adj, vectors=True):
    """perturbation.jl:319-367.  L must be in mode :householder/:compact with two active variables.

    Device work: L(m,n)*v products (combine + SpMV).  The factorisation of the (deliberately near-singular)
    L(0,0) and its N solves are only needed for N >= 2 -- v_N is never used for the eigenvalue coefficients --
    so Newton (order 1) runs without any extra LU (the reference factorises and solves regardless)."""
    if N == 1 and not vectors:
        # first order, coefficients only (the Newton step of householder): lam_1 = -(v0Adj^H L01 v0) / (v0Adj^H L10 v0) does not depend on
        # the scaling of v0 and v0Adj, so the normalisations of the reference (five passes over the vectors) are skipped
        L10v0 = L(1, 0).matvec(v0, slot=2)
        den = np.vdot(v0Adj, L10v0)
        lam1 = -np.vdot(v0Adj, L(0, 1).matvec(v0, slot=2)) / den
        return np.array([0.0, lam1], dtype=complex), [None, None]
    v0 = v0 / np.sqrt(np.vdot(v0, v0))
    L10v0 = L(1, 0).matvec(v0, slot=2)
    v0Adj = v0Adj / np.vdot(v0Adj, L10v0)
    den = np.vdot(v0Adj, L10v0)
    lam = np.zeros(N + 1, dtype=complex)
    v = [None] * (N + 1)
    v[0] = v0
    ctx = L.device().ctx
    lu = None
    for k in range(1, N + 1):
        r = np.zeros(len(v0), dtype=complex)
        for n in range(1, k + 1):
            r += L(0, n).matvec(v[k - n], slot=2)
        for m in range(1, k + 1):
            for p in _partitions(m):
                if p == [k]:
                    continue
                mu = _part2mult(p)
                for n in range(0, k - m + 1):
                    coeff = 1
                    for g, mg in enumerate(mu):
                        coeff *= lam[g + 1] ** mg
                    mult = math.factorial(sum(mu)) / math.prod(math.factorial(x) for x in mu)
                    r += L(sum(mu), n).matvec(v[k - n - m], slot=2) * mult * coeff
        lam[k] = -np.vdot(v0Adj, r) / den
        if k < N:
            if lu is None:
                L(0, 0).materialize(3)
                lu = L.device().lu()
                ctx.lu_factor(lu, 3, check=False)  # lu(L(0,0), check=false): the operator is singular by construction
            vk = ctx.lu_solve(lu, -(r + lam[k] * L10v0))
            v[k] = vk - np.vdot(v0, vk) * v0
    return lam, v


def perturb_disk(L, N, v0, v0Adj, weighted=False):
    """perturbation.jl:373-450 (perturb_disk) and :470-545 (perturb_norm, weighted=True): power-series coefficients of the eigenvalue
    and the eigenvector to order N with ONE factorisation of L(0,0) and N solves on the device.  The sum is organised per derivative
    pair: r = sum_(m,n) L(m,n) w_mn, w_mn = sum over the multi-indices mu with |mu| = m of v[k-n-weight(mu)] multinomial(mu)
    prod_g lambda_g^mu_g -- one combine + SpMV per pair instead of one per multi-index.  The reference reads the multi-indices from
    files shipped with the package; here all partitions of every weight W <= k-n (except {k}) are generated on the fly.
    weighted: inner products in the mass matrix Y = -coeff(__aux__) (v0'Y v0 = 1).  With Y real symmetric the reference's
    lu(Y)\v0Adj followed by v0Adj'Y(.) is v0Adj'(.), which is what is evaluated here (no factorisation of Y)."""
    dv = L.device()
    ctx = dv.ctx
    if weighted:
        Y = (-L.terms[-1].coeff).to_scipy().real
        ip = lambda a, b: np.vdot(a, Y @ b)
    else:
        ip = np.vdot
    v0 = v0 / np.sqrt(ip(v0, v0))
    L10v0 = L(1, 0).matvec(v0, slot=2)
    v0Adj = v0Adj / np.vdot(v0Adj, L10v0)
    den = np.vdot(v0Adj, L10v0)
    lam = np.zeros(N + 1, dtype=complex)
    v = [None] * (N + 1)
    v[0] = v0
    L(0, 0).materialize(3)
    lu = dv.lu()
    ctx.lu_factor(lu, 3, check=False)  # lu(L(0,0), check=false), perturbation.jl:385,493: singular by construction
    for k in range(1, N + 1):
        w = {(0, n): v[k - n].copy() for n in range(1, k + 1)}
        for W in range(1, k + 1):
            for p in _partitions(W):
                if p == [k]:
                    continue
                mu = _part2mult(p)
                coeff = math.factorial(sum(mu)) / math.prod(math.factorial(x) for x in mu)
                for g, mg in enumerate(mu):
                    coeff = coeff * lam[g + 1] ** mg
                for n in range(0, k - W + 1):
                    key = (len(p), n)
                    w[key] = w[key] + v[k - n - W] * coeff if key in w else v[k - n - W] * coeff
        r = np.zeros(len(v0), dtype=complex)
        for (m, n), wv in sorted(w.items()):
            r += L(m, n).matvec(wv, slot=2)
        lam[k] = -np.vdot(v0Adj, r) / den
        vk = ctx.lu_solve(lu, -(r + lam[k] * L10v0))
        vk = vk - ip(v0, vk) * v0
        c = 0j
        for l in range(1, k):
            c -= 0.5 * ip(v[l], v[k - l])
        v[k] = vk + c * v[0]
    return lam, v


def _perturb_driver(fn, sol, L, param, N, mode):
    active, params, cur = L.active, L.params, L.mode
    L.params = sol.params
    L.active = [sol.eigval, param]
    L.mode = mode
    key = f"{param}/Taylor"
    try:
        sol.eigval_pert[key], sol.v_pert[key] = fn(L, N, sol.v, sol.v_adj)
        sol.eigval_pert[key][0] = sol.params[sol.eigval]
    finally:
        L.active, L.mode, L.params = active, cur, params


def perturb_fast_bang(sol, L, param, N, mode="compact"):
    """perturb_fast! (LinOpFam.jl:576-590)"""
    _perturb_driver(perturb_disk, sol, L, param, N, mode)


def perturb_norm_bang(sol, L, param, N, mode="compact"):
    """perturb_norm! (LinOpFam.jl:606-620): as perturb_fast! with mass-weighted normalisation of the eigenvector series."""
    _perturb_driver(lambda L_, N_, v_, va_: perturb_disk(L_, N_, v_, va_, weighted=True), sol, L, param, N, mode)


def pade_bang(sol, param, Lo, M, vector=False):
    """pade! (LinOpFam.jl:644-676): [Lo/M] Pade approximant of the eigenvalue (and, component-wise, of the eigenvector) series."""
    key, tkey = f"{param}/[{Lo}/{M}]", f"{param}/Taylor"
    sol.eigval_pert[key] = pade(sol.eigval_pert[tkey], Lo, M)
    if vector:
        V = np.array(sol.v_pert[tkey][: Lo + M + 1])  # (order, dof)
        A = np.empty((Lo + 1, V.shape[1]), dtype=complex)
        B = np.empty((M + 1, V.shape[1]), dtype=complex)
        for i in range(V.shape[1]):
            A[:, i], B[:, i] = pade(V[:, i], Lo, M)
        sol.v_pert[key] = (A, B)


def conv_radius(sol_or_coeffs, param=None):
    """conv_radius (LinOpFam.jl:754-766): ratio estimates |a_n / a_(n+1)| of the convergence radius of the power series."""
    a = sol_or_coeffs.eigval_pert[f"{param}/Taylor"] if param is not None else sol_or_coeffs
    a = np.asarray(a)
    return np.abs(a[:-1] / a[1:])


def perturb_bang(sol, L, param, N, mode="compact", vectors=True):
    """perturb! (LinOpFam.jl:546-560).  vectors=False (internal: the Newton step of householder / mslp uses the eigenvalue coefficients
    only) lets a first-order call skip the normalisation of the eigenvector pair."""
    active, params, cur = L.active, L.params, L.mode
    L.params = sol.params
    L.active = [sol.eigval, param]
    L.mode = mode
    key = f"{param}/Taylor"
    try:
        sol.eigval_pert[key], sol.v_pert[key] = perturb(L, N, sol.v, sol.v_adj, vectors=vectors)
        sol.eigval_pert[key][0] = sol.params[sol.eigval]
    finally:
        L.active, L.mode, L.params = active, cur, params


def pade(w, Lo, M):
    """LinOpFam.jl:622-642"""
    A = np.zeros((M, M), dtype=complex)
    for i in range(1, M + 1):
        for j in range(1, M + 1):
            if Lo + i - j >= 0:
                A[i - 1, j - 1] = w[Lo + i - j]
    b = np.linalg.solve(A, -np.asarray(w[Lo + 1: Lo + M + 1])) if M > 0 else np.zeros(0, dtype=complex)
    b = np.concatenate([[1.0], b])
    a = np.zeros(Lo + 1, dtype=complex)
    for l in range(Lo + 1):
        for m in range(l + 1):
            if m <= M:
                a[l] += w[l - m] * b[m]
    return a, b


def polyval(p, z):
    """LinOpFam.jl:715-730 (Horner)"""
    f = p[-1]
    for c in p[-2::-1]:
        f = f * z + c
    return f


def poly_roots(p):
    """Householder.jl:195-203"""
    N = len(p) - 1
    Cm = np.zeros((N, N), dtype=complex)
    for i in range(1, N):
        Cm[i, i - 1] = 1
    Cm[:, N - 1] = -np.asarray(p[:N]) / p[N]
    return np.linalg.eigvals(Cm)


def householder_update(f):
    """Householder.jl:21-35"""
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


# --------------------------------------------------------------------------------------------- local solvers
# mslp status flags (iterative_solvers.jl:4-14)
itsol_converged, itsol_maxiter, itsol_slow_convergence = 0, 1, 2
itsol_impossible, itsol_singular_exception, itsol_arpack_exception, itsol_isnan, itsol_unknown = -1, -2, -3, -4, -5


def _iterate(L, z, maxiter, tol, relax, order, nev, v0, v0_adj, kind, num_order, output, scale, stats):
    z = complex(z)
    z0 = complex("inf")
    lam = float("inf")
    lam0 = float("inf")
    n = 0
    active, mode = L.active, L.mode
    if v0 is None or len(v0) == 0:
        L(0)  # reference: size(L(0)) -- zeroes the eigenvalue parameter as a side effect
        v0 = np.ones(L.size(), dtype=complex)
    if v0_adj is None or len(v0_adj) == 0:
        v0_adj = np.conj(v0)
    err = None
    dev = L.device()
    ctx = dev.ctx
    lu = dev.lu()
    # M = -L.terms[end].coeff, kept in family slot 1
    mcoef = [None] * len(L.terms)
    mcoef[-1] = -1.0
    dev.combine(dev.flat(mcoef), 1)
    paired = os.environ.get("WAE_EIGS_PAIRED", "1") == "1" and hasattr(ctx, "eigs_si_pair")
    try:
        while abs(z - z0) > tol and n < maxiter:
            if output:
                print(n, "\t\t", abs(lam), "\t", abs(z - z0) / scale, "\t", z / scale, flush=True)
            if kind == "householder":
                z0 = z
            L.params[L.eigval] = z
            L.params[L.auxval] = 0
            _t0 = time.perf_counter()
            L(z).materialize(0)
            ctx.lu_factor(lu, 0)
            _t1 = time.perf_counter()
            if paired:  # opt-in: both Arnoldi recurrences as the two right-hand sides of one pass over the factor
                try:
                    lams, v, lams_adj, v_adj, ns1 = ctx.eigs_si_pair(lu, dev.fid, 1, nev, v0, v0_adj)
                    ns2 = 0
                except _lib.WaeError as e:
                    if e.code != _lib.E_INVALID:
                        raise
                    paired = False  # general (unsymmetric) elimination: two separate calls
            if not paired:
                lams, v, ns1 = ctx.eigs_si(lu, dev.fid, 1, nev, v0, trans=0)
                lams_adj, v_adj, ns2 = ctx.eigs_si(lu, dev.fid, 1, nev, v0_adj, trans=2)
            _t2 = time.perf_counter()
            if stats is not None:
                stats["factorizations"] = stats.get("factorizations", 0) + 1
                stats["solves"] = stats.get("solves", 0) + ns1 + ns2
                stats["factor_ms"] = stats.get("factor_ms", 0.0) + ctx.last_ms("factor")
                stats["combine_factor_wall_s"] = stats.get("combine_factor_wall_s", 0.0) + _t1 - _t0
                stats["eigs_wall_s"] = stats.get("eigs_wall_s", 0.0) + _t2 - _t1
            if nev > 1:  # (one pair: nothing to sort, and v[:, idx] would copy the vectors)
                idx = np.argsort(np.abs(lams), kind="stable")
                lams, v = lams[idx], v[:, idx]
                idx = np.argsort(np.abs(lams_adj), kind="stable")
                lams_adj, v_adj = lams_adj[idx], v_adj[:, idx]
            dzs, back = [], []
            L.active = [L.auxval, L.eigval]
            for i in range(nev):
                L.params[L.auxval] = lams[i]
                sol = Solution(L.params, v[:, i], v_adj[:, i], L.auxval)
                perturb_bang(sol, L, L.eigval, order, mode="householder", vectors=False)
                coeffs = sol.eigval_pert[f"{L.eigval}/Taylor"]
                if kind == "householder":
                    dzs.append(householder_update([math.factorial(k) * c for k, c in enumerate(coeffs)]))
                else:
                    num, den = pade(coeffs, num_order, order - num_order)
                    roots = poly_roots(num)
                    dzs.append(roots[np.argsort(np.abs(roots), kind="stable")[0]])
                    if z0 != complex("inf"):
                        back.append(lam0 - polyval(num, z0 - z) / polyval(den, z0 - z))
            L.active = [L.eigval]
            if stats is not None:
                stats["perturb_wall_s"] = stats.get("perturb_wall_s", 0.0) + time.perf_counter() - _t2
            sel = np.argsort(np.abs(back if back else dzs), kind="stable")[0]
            lam = lams[sel]
            L.params[L.auxval] = lam
            if kind == "mslp":
                z0, lam0 = z, lam
            z = z + relax * dzs[sel]
            if relax == 1.0:  # the reference's formula with relax = 1, without three passes over each vector
                v0, v0_adj = v[:, sel], v_adj[:, sel]
            else:
                v0 = (1 - relax) * v0 + relax * v[:, sel]
                v0_adj = (1 - relax) * v0_adj + relax * v_adj[:, sel]
            n += 1
    except _lib.ArpackException as e:
        err = "arpack"
        if output:
            print("Error occured:", e)
    except _lib.SingularException as e:
        err = "singular"
        L.params[L.eigval] = z
        if output:
            print("Error occured:", e)
    except _lib.WaeError as e:
        # any other failure of the device path (CUDA error, out of memory, invalid argument): as the reference's catch-all
        # (Householder.jl:131-149, iterative_solvers.jl:192-210) the solver state is restored and the flag is "unknown"
        err = "unknown"
        L.params[L.eigval] = z
        L.active, L.mode = active, mode
        if output:
            print("Error occured:", e)
        return Solution(L.params, v0, v0_adj, L.eigval), n, z, z0, lam, err
    if err is None:
        L.params[L.eigval] = z
        if output:
            print(n, "\t\t", abs(lam), "\t", abs(z - z0) / scale, "\t", z / scale)
    L.active, L.mode = active, mode
    # normalisation (Householder.jl:189-190)
    Mv = ctx.spmm(dev.fid, 1, v0)
    v0 = v0 / np.sqrt(np.vdot(v0, Mv))
    dLv = L(L.params[L.eigval], 1).matvec(v0, slot=2)
    v0_adj = v0_adj / np.conj(np.vdot(v0_adj, dLv))
    return Solution(L.params, v0, v0_adj, L.eigval), n, z, z0, lam, err


def householder(L, z, maxiter=10, tol=0.0, relax=1.0, lam_tol=float("inf"), order=1, nev=1, v0=None, v0_adj=None,
                output=True, stats=None):
    """Householder.jl:70-192.  Returns (Solution, n, flag): 1 converged, 0 slow convergence, -1 maxiter,
    -4 Arnoldi failure, -6 singular factorisation."""
    if output:
        print("Launching Householder...\nIter    Res:     dz:     z:\n----------------------------------")
    sol, n, z, z0, lam, err = _iterate(L, z, maxiter, tol, relax, order, nev, v0, v0_adj, "householder", 1, output, 1, stats)
    if err == "arpack":
        flag = -4
    elif err == "singular":
        flag = -6
    elif err == "unknown":
        flag = -2
    elif n >= maxiter:
        flag = -1
    elif abs(lam) <= lam_tol:
        flag = 1
    elif abs(z - z0) <= tol:
        flag = 0
    elif z != z:
        flag = -5
    else:
        flag = -3
    if output and err is None:
        print("...finished Householder!\nNumber of steps: ", n, "\nLast step parameter variation:", abs(z0 - z),
              "\nAuxiliary eigenvalue λ residual (rhs):", abs(lam), "\nEigenvalue:", z, "\nEigenvalue/(2*pi):", z / 2 / math.pi)
    return sol, n, flag


def mslp(L, z, maxiter=10, tol=0.0, relax=1.0, lam_tol=float("inf"), order=1, nev=1, v0=None, v0_adj=None, num_order=1,
         scale=1, output=True, stats=None):
    """iterative_solvers.jl:93-252.  Flags: 0 converged, 1 maxiter, 2 slow convergence, -2 singular, -3 Arnoldi."""
    if output:
        print("Launching MSLP solver...\nIter   dz:     z:\n----------------------------------")
    if L.terms[-1].operator != "__aux__":
        # reference appends -I with a new parameter :__aux__ (iterative_solvers.jl:119-123)
        import scipy.sparse as sp
        ctx = L.terms[0].coeff.ctx
        eye = DeviceMatrix.from_scipy(-sp.identity(L.size(), dtype=complex, format="csc"), ctx)
        L.push(Term(eye, (pow1,), (("__aux__",),), "__aux__", "__aux__"))
        L.auxval = "__aux__"
    sol, n, z, z0, lam, err = _iterate(L, z * scale, maxiter, tol * scale, relax, order, nev, v0, v0_adj, "mslp", num_order,
                                       output, scale, stats)
    if err == "arpack":
        flag = itsol_arpack_exception
    elif err == "singular":
        flag = itsol_singular_exception
    elif err == "unknown":
        flag = itsol_unknown
    elif n >= maxiter:
        flag = itsol_maxiter
    elif abs(lam) <= lam_tol:
        flag = itsol_converged
    elif abs(z - z0) <= tol * scale:
        flag = itsol_slow_convergence
    elif z != z:
        flag = itsol_isnan
    else:
        flag = itsol_impossible
    return sol, n, flag


# --------------------------------------------------------------------------------------------- Beyn
def wn(z, G):
    """Winding number (beyn.jl:185-209)."""
    w = 0
    for i in range(len(G)):
        a, b = complex(G[i]), complex(G[(i + 1) % len(G)])
        isleft = (b.real - a.real) * (z.imag - a.imag) - (z.real - a.real) * (b.imag - a.imag)
        if a.imag <= z.imag:
            if b.imag > z.imag and isleft > 0:
                w += 1
        elif b.imag <= z.imag and isleft < 0:
            w -= 1
    return w


def inpoly(z, G):
    return wn(complex(z), G) != 0


def contour_nodes(G, N):
    """Gauss-Legendre nodes and weights (incl. the (b-a)/2 factor) on every polygon edge (beyn.jl:112-138)."""
    X, W = np.polynomial.legendre.leggauss(N)
    zs, ws = [], []
    for i in range(len(G)):
        a, b = complex(G[i]), complex(G[(i + 1) % len(G)])
        zs.extend(X * (b - a) / 2 + (a + b) / 2)
        ws.extend(W * (b - a) / 2)
    return np.array(zs), np.array(ws)


def shard_nodes(n_nodes, rank, world):
    """Quadrature nodes of rank `rank`: round-robin, so that every rank gets nodes from every polygon edge."""
    return np.arange(rank, n_nodes, world)


def allreduce_moments(A, group=None):
    """Sum the per-rank partial moment tensors (torch complex tensor, in place) over the process group.
    The only collective of the path: NCCL over NVLink on GPUs (gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(torch.view_as_real(A), group=group)
    return A


def compute_moment_matrices(L, G, l=5, K=1, N=16, group=None, stats=None, V=None, replicas=None):
    """Moments A_p = sum_j w_j z_j^p L(z_j)^{-1} V, p < 2K, V = first l identity columns (beyn.jl:62-74).

    The quadrature nodes are sharded round-robin over the ranks of ``group`` (torch.distributed, NCCL):
    each GPU factorises its own nodes, the moments are summed with one all-reduce.  Returns a (d, l, 2K)
    complex numpy array (on every rank).

    ``replicas``: the same family discretised on OTHER devices of this process (``discretize(..., ctx=Context(dev))``): all nodes
    are then run from this one process through wae_beyn_moments_multi -- node j on device j mod (1 + len(replicas)), one host
    thread per device and one in-library ncclAllReduce -- which is how a single-process caller (the reference's Julia beyn) scales."""
    dev = L.device()
    ctx = dev.ctx
    d = dev.dim
    L(0)
    zs, ws = contour_nodes(G, N)
    if replicas:
        fams = [L] + list(replicas)
        coeffs = np.zeros((len(zs), dev.n_flat), dtype=np.complex128)
        for j in range(len(zs)):
            L.params[L.eigval] = complex(zs[j])
            coeffs[j] = dev.flat(L.scalars([0] * len(L.active)))
        devs = [f.device() for f in fams]
        A = _lib.Context.beyn_moments_multi([x.ctx for x in devs], [x.fid for x in devs], [x.lu() for x in devs], zs, ws, coeffs, l, 2 * K, d, V=V)
        if stats is not None:
            stats["factorizations"] = stats.get("factorizations", 0) + len(zs)
            stats["factor_ms"] = stats.get("factor_ms", 0.0) + max(max(0.0, x.ctx.last_ms("beyn_factor_total")) for x in devs)
            stats["solve_ms"] = stats.get("solve_ms", 0.0) + max(max(0.0, x.ctx.last_ms("beyn_solve_total")) for x in devs)
        return np.ascontiguousarray(A)
    import torch
    import torch.distributed as dist
    rank, world = 0, 1
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    mine = shard_nodes(len(zs), rank, world)
    coeffs = np.zeros((len(mine), dev.n_flat), dtype=np.complex128)
    for r, j in enumerate(mine):
        L.params[L.eigval] = complex(zs[j])
        coeffs[r] = dev.flat(L.scalars([0] * len(L.active)))
    A = ctx.moment_buffer(2 * K, l, d)  # device tensor, == (d,l,2K) column-major
    if len(mine):
        ctx.beyn_moments(dev.fid, dev.lu(), zs[mine], ws[mine], coeffs, l, 2 * K, A.data_ptr(), V=V)
    if stats is not None:
        stats["factorizations"] = stats.get("factorizations", 0) + len(mine)
        if len(mine):  # (a rank without nodes has no timings: wae_last_ms answers -1)
            stats["factor_ms"] = stats.get("factor_ms", 0.0) + max(0.0, ctx.last_ms("beyn_factor_total"))
            stats["solve_ms"] = stats.get("solve_ms", 0.0) + max(0.0, ctx.last_ms("beyn_solve_total"))
    allreduce_moments(A, group)
    return A.permute(2, 1, 0).cpu().numpy()


def moments2eigs(A, G=None, tol=0.0, pos_test=True, output=False, rtol=0.0):
    """Block-Hankel SVD + small eigenproblem (beyn.jl:77-107, 289-323).  ``tol`` is the reference's absolute
    singular-value threshold; ``rtol`` (extension) is relative to the largest singular value."""
    d, l, K2 = A.shape
    K = K2 // 2
    B0 = np.zeros((d * K, l * K), dtype=complex)
    B1 = np.zeros((d * K, l * K), dtype=complex)
    for i in range(K):
        for j in range(K):
            B0[d * i: d * (i + 1), l * j: l * (j + 1)] = A[:, :, i + j]
            B1[d * i: d * (i + 1), l * j: l * (j + 1)] = A[:, :, i + j + 1]
    V, S, Wh = np.linalg.svd(B0, full_matrices=False)
    W = Wh.conj().T
    if output:
        print("############\nsingular values:\n", S)
    if tol > 0 or rtol > 0:
        m = S > max(tol, rtol * S[0])
        V, S, W = V[:, m], S[m], W[:, m]
    Om, P = np.linalg.eig(V.conj().T @ B1 @ W @ np.diag(1 / S))
    P = V[:d, :] @ P
    if pos_test and G is not None:
        m = np.array([inpoly(z, G) for z in Om], dtype=bool)
        Om, P = Om[m], P[:, m]
    return Om, P


def beyn(L, G, l=5, K=1, N=16, tol=0.0, pos_test=True, output=True, random=False, group=None, stats=None, seed=0, replicas=None):
    """beyn.jl:34-110.  N is the number of quadrature nodes PER polygon edge.  ``group`` / ``replicas``: see compute_moment_matrices."""
    d = L.size()
    K = max(K, l // d + int(l % d != 0))
    V = None
    if random and l < d:
        # beyn.jl:42-43 uses rand(ComplexF64,d,l) (unseeded); here the probing matrix is seeded so that every rank of a
        # sharded run draws the same V (and runs are reproducible)
        rng = np.random.default_rng(seed)
        V = rng.random((d, l)) + 1j * rng.random((d, l))
    A = compute_moment_matrices(L, G, l=min(l, d), K=K, N=N, group=group, stats=stats, V=V, replicas=replicas)
    return moments2eigs(A, G, tol=tol, pos_test=pos_test, output=output)


def bloch_expand(mesh, sol_or_vec, b="b"):
    """Bloch.jl:120-143: expand a unit-cell vector to the full annulus, sector s multiplied by exp(2 pi i b s / DOS)."""
    d = mesh.dos
    if isinstance(sol_or_vec, Solution):
        vec, B = sol_or_vec.v, sol_or_vec.params[b]
    else:
        vec, B = np.asarray(sol_or_vec), (0 if isinstance(b, str) else b)
    v = np.zeros(d.naxis + d.nxsector * d.DOS, dtype=complex)
    v[: d.naxis] = vec[: d.naxis]
    for s_ in range(d.DOS):
        v[d.naxis + s_ * d.nxsector: d.naxis + (s_ + 1) * d.nxsector] = vec[d.naxis: d.naxis + d.nxsector] * np.exp(2j * math.pi / d.DOS * B * s_)
    return v
