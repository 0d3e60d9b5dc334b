"""Oracle (test infrastructure): P1/P2 element matrices of the reference's FEM library.

Restates ``src/FEM/FEM.jl`` of the reference:
  CooTrafo            FEM.jl:2-21      -> coo_trafo
  create_indices      FEM.jl:22-32     -> create_indices
  s33v1u1/s33v2u2     FEM.jl:435-450   -> tri_mass
  s33v1u1c1/s33v2u2c1 FEM.jl:469-525   -> tri_mass_c1
  s43v1u1/s43v2u2     FEM.jl:704-738   -> tet_mass
  s43v1u1c1/s43v2u2c1 FEM.jl:764-890   -> tet_mass_c1
  s43nv1nu1/s43nv2nu2 FEM.jl:1745-1874 -> tet_stiff
  s43nv?nu?cc1        FEM.jl:2283-2424 -> tet_stiff_cc1
  s43v1/s43v2         FEM.jl:2429-2435 -> tet_src
  s43nv1rx/s43nv2rx   FEM.jl:2442-2484 -> tet_grad_at
  s33v1/s33v2(/c1)    FEM.jl:2557-2589 -> tri_src, tri_src_c1
  f1/f2               FEM.jl:2611-2633 -> shape

The reference stores each of these as a literal table / generated polynomial in
the entries of A = inv*inv'.  They are the exact integrals of the standard
Lagrange bases (P2 local order [v1,v2,v3,v4,e12,e13,e14,e23,e24,e34], triangles
[v1,v2,v3,e12,e13,e23]); here the same numbers are obtained by exact rational
integration of barycentric monomials,
    int_T l1^a l2^b l3^c l4^d = |det J| a! b! c! d! / (a+b+c+d+3)!   (tets)
    int_F l1^a l2^b l3^c      = |det|   a! b! c!    / (a+b+c+2)!     (triangles, |det| = 2 area)
and pinned against the reference's own expressions in tests/golden/fem_tables.npz.
"""
from fractions import Fraction
from functools import lru_cache
from math import factorial

import numpy as np

# ----------------------------------------------------------------------------
# tiny exact polynomial algebra in barycentric coordinates
# ----------------------------------------------------------------------------


def _pmul(p, q):
    r = {}
    for ea, ca in p.items():
        for eb, cb in q.items():
            e = tuple(x + y for x, y in zip(ea, eb))
            r[e] = r.get(e, 0) + ca * cb
    return r


def _padd(p, q, s=1):
    r = dict(p)
    for e, c in q.items():
        r[e] = r.get(e, 0) + s * c
    return r


def _pdiff(p, a):
    r = {}
    for e, c in p.items():
        if e[a] > 0:
            e2 = list(e)
            e2[a] -= 1
            r[tuple(e2)] = r.get(tuple(e2), 0) + c * e[a]
    return r


def _pint(p, nb):
    """Integral over the reference simplex with nb barycentric coords, per unit |det|."""
    tot = Fraction(0)
    for e, c in p.items():
        num = 1
        for x in e:
            num *= factorial(x)
        tot += Fraction(c) * Fraction(num, factorial(sum(e) + nb - 1))
    return tot


def _lam(k, nb):
    e = [0] * nb
    e[k] = 1
    return {tuple(e): Fraction(1)}


def _basis(order, nb):
    """Lagrange basis polynomials; nb=4 tets, nb=3 triangles."""
    lam = [_lam(k, nb) for k in range(nb)]
    if order == 1:
        return lam
    one = {tuple([0] * nb): Fraction(1)}
    phis = []
    for k in range(nb):  # vertex functions (2 l_k - 1) l_k
        phis.append(_pmul(_padd({e: 2 * c for e, c in lam[k].items()}, one, -1), lam[k]))
    for a in range(nb):  # edge functions 4 l_a l_b in the order 12,13,14,23,24,34
        for b in range(a + 1, nb):
            phis.append({e: 4 * c for e, c in _pmul(lam[a], lam[b]).items()})
    return phis


@lru_cache(maxsize=None)
def tables(order, nb):
    """Exact tables (float64) for one element family."""
    phi = _basis(order, nb)
    n = len(phi)
    lam = [_lam(k, nb) for k in range(nb)]
    mass = np.zeros((n, n))
    massc = np.zeros((n, n, nb))
    src = np.zeros(n)
    srcc = np.zeros((n, nb))
    for i in range(n):
        src[i] = float(_pint(phi[i], nb))
        for k in range(nb):
            srcc[i, k] = float(_pint(_pmul(phi[i], lam[k]), nb))
        for j in range(n):
            pij = _pmul(phi[i], phi[j])
            mass[i, j] = float(_pint(pij, nb))
            for k in range(nb):
                massc[i, j, k] = float(_pint(_pmul(pij, lam[k]), nb))
    out = dict(n=n, mass=mass, massc=massc, src=src, srcc=srcc)
    if nb == 4:
        dphi = [[_pdiff(p, a) for a in range(4)] for p in phi]
        stiff = np.zeros((n, n, 4, 4))
        stiffcc = np.zeros((n, n, 4, 4, 4, 4))
        ll = [[_pmul(lam[k], lam[l]) for l in range(4)] for k in range(4)]
        for i in range(n):
            for j in range(n):
                for a in range(4):
                    for b in range(4):
                        g = _pmul(dphi[i][a], dphi[j][b])
                        if not g:
                            continue
                        stiff[i, j, a, b] = float(_pint(g, 4))
                        for k in range(4):
                            for l in range(4):
                                stiffcc[i, j, a, b, k, l] = float(_pint(_pmul(g, ll[k][l]), 4))
        out.update(stiff=stiff, stiffcc=stiffcc, dphi=dphi)
    return out


# ----------------------------------------------------------------------------
# reference-named element routines
# ----------------------------------------------------------------------------


class CooTrafo:
    """FEM.jl:2-21.  X is 3x4 (tet) or 3x3 (triangle), last column is the origin."""

    def __init__(self, X):
        X = np.asarray(X, dtype=float)
        d, m = X.shape
        J = np.empty((d, d))
        self.orig = X[:, -1].copy()
        J[:, : m - 1] = X[:, :-1] - X[:, -1:]
        if m == 3:
            nrm = np.cross(J[:, 0], J[:, 1])
            J[:, -1] = nrm / np.linalg.norm(nrm)
        self.trafo = J
        self.inv = np.linalg.inv(J)
        self.det = np.linalg.det(J)


def create_indices(smplx):
    """FEM.jl:22-32: ii[i,j]=smplx[i], jj=ii'."""
    s = np.asarray(smplx)
    ii = np.repeat(s[:, None], len(s), axis=1)
    return ii, ii.T


def _grads(ct):
    """Rows: gradients of the four barycentric coordinates (FEM.jl:2442-2448)."""
    return np.vstack([ct.inv, -ct.inv.sum(axis=0)])


def tet_mass(ct, order):
    return tables(order, 4)["mass"] * abs(ct.det)


def tet_mass_c1(ct, c, order):
    return tables(order, 4)["massc"] @ np.asarray(c) * abs(ct.det)


def tet_stiff(ct, order):
    G = _grads(ct)
    return np.einsum("ijab,ab->ij", tables(order, 4)["stiff"], G @ G.T) * abs(ct.det)


def tet_stiff_cc1(ct, c, order):
    G = _grads(ct)
    c = np.asarray(c)
    return np.einsum("ijabkl,ab,k,l->ij", tables(order, 4)["stiffcc"], G @ G.T, c, c) * abs(ct.det)


def tet_src(ct, order):
    return tables(order, 4)["src"] * abs(ct.det)


def shape(ct, p, order):
    """f1/f2 (FEM.jl:2611-2633): basis values at physical point p."""
    x = ct.inv @ (np.asarray(p, dtype=float) - ct.orig)
    lam = np.append(x, 1.0 - x.sum())
    phi = _basis(order, 4)
    return np.array([sum(float(c) * np.prod(lam ** np.array(e)) for e, c in pk.items()) for pk in phi])


def tet_grad_at(ct, n_ref, x_ref, order):
    """grad(phi_j)(x_ref) . n_ref  (FEM.jl:2442-2484)."""
    G = _grads(ct)
    gn = G @ np.asarray(n_ref, dtype=float)
    x = ct.inv @ (np.asarray(x_ref, dtype=float) - ct.orig)
    lam = np.append(x, 1.0 - x.sum())
    dphi = tables(order, 4)["dphi"]
    out = np.zeros(len(dphi))
    for j, dj in enumerate(dphi):
        for a in range(4):
            v = sum(float(c) * np.prod(lam ** np.array(e)) for e, c in dj[a].items())
            out[j] += v * gn[a]
    return out


def tri_mass(ct, order):
    return tables(order, 3)["mass"] * abs(ct.det)


def tri_mass_c1(ct, c, order):
    return tables(order, 3)["massc"] @ np.asarray(c) * abs(ct.det)


def tri_src(ct, order):
    return tables(order, 3)["src"] * abs(ct.det)


def tri_src_c1(ct, c, order):
    return tables(order, 3)["srcc"] @ np.asarray(c) * abs(ct.det)
