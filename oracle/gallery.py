"""Oracle (test infrastructure): the assembly-free fixture of the reference's NLEVP gallery.

  rijke_tube(resolution; l, c_max, mid)      src/NLEVP/gallery.jl:171-260

One-dimensional thermoacoustic Rijke tube, first-order elements on a uniform grid:
    d/dx c^2 dp/dx + w^2 p - n exp(-i w tau) dp/dx(x_ref) = 0 on ]0,l[,   dp/dx(0) = 0,  p(l) = 0 (Y = 1e15 penalty)
as the operator family  w^2 M + K + w Y C + n exp(-i w tau) Q - lambda M.  SURVEY section 8(c) lists it as a fixture that exercises the
NLEVP solvers without the FEM path.  0-based indices.
"""
import numpy as np
import scipy.sparse as sp

from .nlevp import LinearOperatorFamily, Term, exp_delay, pow1, pow2


def rijke_tube(resolution=127, l=1.0, c_max=2.0, mid=0):
    n, tau, c_min = 1.0, 2.0, 1.0
    outlet = resolution - 1
    grid = np.linspace(0.0, l, resolution)
    e2p = [(i, i + 1) for i in range(resolution - 1)]
    if mid == 0:
        mid = resolution // 2 + 1  # 1-based number of the element that holds the flame (gallery.jl:184-187)
    ref = mid - 1                  # 1-based number of the reference element
    e2v = np.diff(grid)
    V = e2v[mid - 1]
    e2c = [c_min if i + 1 < mid else c_max for i in range(resolution)]
    m_unit = np.array([[2.0, 1.0], [1.0, 2.0]]) / 6
    k_unit = -np.array([[1.0, -1.0], [-1.0, 1.0]])
    I, J, MM, KK = [], [], [], []
    for idx, el in enumerate(e2p):
        for a in range(2):
            for b in range(2):
                I.append(el[a]); J.append(el[b])
                MM.append(m_unit[a, b] * e2v[idx])
                KK.append(k_unit[a, b] / e2v[idx] * e2c[idx] ** 2)
    d = resolution
    M = sp.csc_matrix(sp.coo_matrix((np.asarray(MM, dtype=complex), (I, J)), shape=(d, d)))
    K = sp.csc_matrix(sp.coo_matrix((np.asarray(KK, dtype=complex), (I, J)), shape=(d, d)))
    B = sp.csc_matrix(sp.coo_matrix(([-c_max * 1j], ([outlet], [outlet])), shape=(d, d)))
    grad_p_ref = np.array([-1.0, 1.0]) / e2v[ref - 1]
    el, rf = e2p[mid - 1], e2p[ref - 1]
    qi, qj, qq = [], [], []
    for i in range(2):       # rows: the two nodes of the flame element, each with int phi_i = e2v/2
        for j in range(2):   # columns: the two nodes of the reference element
            qi.append(el[i]); qj.append(rf[j]); qq.append(-grad_p_ref[j] * e2v[mid - 1] / 2)
    Q = sp.csc_matrix(sp.coo_matrix((np.asarray(qq, dtype=complex), (qi, qj)), shape=(d, d))) / V
    L = LinearOperatorFamily(["ω", "n", "τ", "Y", "λ"], [0.0, n, tau, 1e15, float("inf")])
    L.push(Term(M, (pow2,), (("ω",),), "ω^2", "M"))
    L.push(Term(K, (), (), "", "K"))
    L.push(Term(B, (pow1, pow1), (("ω",), ("Y",)), "ω*Y", "C"))
    L.push(Term(Q, (pow1, exp_delay), (("n",), ("ω", "τ")), "n*exp(-i ω τ)", "Q"))
    L.push(Term(-M, (pow1,), (("λ",),), "-λ", "__aux__"))
    return L, grid
