"""Emit wavesandeigenvalues.jl_b200/csrc/fem_gen.h: element tables for the CUDA kernels.

The numbers are exact integrals of the standard P1/P2 Lagrange bases on the
reference simplex (same derivation as the test oracle, but this script is a
build tool of the product and does not import oracle/).  They reproduce the
literal tables / generated polynomials of the reference's src/FEM/FEM.jl
(:435-450, :704-738, :1745-1874, :2429-2435, :2557-2563) to round-off, which
tests/test_oracle_golden.py pins through tests/golden/fem_tables.npz.

    python tools/gen_fem_tables.py
"""
import os
from fractions import Fraction
from math import factorial


def pmul(p, q):
    r = {}
    for ea, ca in p.items():
        for eb, cb in q.items():
            e = tuple(x + y for x, y in zip(ea, eb))
            r[e] = r.get(e, 0) + ca * cb
    return r


def pdiff(p, a):
    r = {}
    for e, c in p.items():
        if e[a]:
            e2 = list(e); e2[a] -= 1
            r[tuple(e2)] = r.get(tuple(e2), 0) + c * e[a]
    return r


def pint(p, nb):
    t = Fraction(0)
    for e, c in p.items():
        num = 1
        for x in e:
            num *= factorial(x)
        t += Fraction(c) * Fraction(num, factorial(sum(e) + nb - 1))
    return t


def basis(order, nb):
    lam = []
    for k in range(nb):
        e = [0] * nb; e[k] = 1
        lam.append({tuple(e): Fraction(1)})
    if order == 1:
        return lam
    phi = []
    for k in range(nb):
        e2 = [0] * nb; e2[k] = 2
        e1 = [0] * nb; e1[k] = 1
        phi.append({tuple(e2): Fraction(2), tuple(e1): Fraction(-1)})
    for a in range(nb):
        for b in range(a + 1, nb):
            phi.append({e: 4 * c for e, c in pmul(lam[a], lam[b]).items()})
    return phi


def lit(fr):
    return repr(float(fr))


def sym_table(name, n, fn):
    vals = [lit(fn(i, j)) for i in range(n) for j in range(n)]
    return f"static __device__ const double {name}[{n * n}] = {{{', '.join(vals)}}};\n"


# ---- star program (assembly generation 3) ----------------------------------------------------------------------------
# Every nonzero (p, q) of M and K belongs to the sub-simplex spanned by the vertices of its two DOFs (a vertex, an edge, a face or
# the tetrahedron); its sources are the elements of that simplex's star.  Because the Lagrange bases are invariant under vertex
# permutations, K_e[p, q] is the SAME linear combination of the element's gram entries g_xy (x, y vertices of the simplex) in every
# element of the star, and M_e[p, q] is one coefficient times |det|.  The kernel therefore sums g_xy and |det| over the star once
# per simplex and forms all nonzeros of the simplex from those sums ("roles").  This function derives the combinations from the
# exact integrals, checks the permutation invariance it relies on, and emits them.
def star_section(order, tag, phi, dphi):
    from itertools import permutations
    n = len(phi)
    edges = [(a, b) for a in range(4) for b in range(a + 1, 4)]

    def dof_verts(i):
        return (i,) if i < 4 else edges[i - 4]

    def dof_of(verts):
        verts = tuple(sorted(verts))
        return verts[0] if len(verts) == 1 else 4 + edges.index(verts)

    def kcoef(i, j):  # {(x, y) x<=y: coefficient of g_xy in K[i][j]}
        r = {}
        for a in range(4):
            for b in range(4):
                c = pint(pmul(dphi[i][a], dphi[j][b]), 4)
                if c != 0:
                    key = (min(a, b), max(a, b))
                    r[key] = r.get(key, 0) + c
        return {k: v for k, v in r.items() if v != 0}

    if order == 1:
        ents = [("vert", (0,), [(0, 0)], [(0, 0)]), ("edge", (0, 1), [(0, 1)], [(0, 1)])]
    else:
        ents = [
            ("vert", (0,), [(0, 0)], [(0, 0)]),
            ("edge", (0, 1), [(0, 0), (1, 1), (0, 1)], [(0, 1), (0, 4), (1, 4), (4, 4)]),
            ("face", (0, 1, 2), [(0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2)], [(2, 4), (1, 5), (0, 7), (4, 5), (4, 7), (5, 7)]),
            ("tet", (0, 1, 2, 3), [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)], [(4, 9), (5, 8), (6, 7)]),
        ]
    out = [f"// ---- star program, {tag}: roles of the nonzeros of a vertex / edge / face / tetrahedron star (see tools/gen_fem_tables.py)\n"]
    mass = []
    nrole = 0
    for name, verts, sums, roles in ents:
        body = []
        for r, (p, q) in enumerate(roles):
            kc = kcoef(p, q)
            assert set(kc) <= set(sums), (name, p, q, kc)
            # permutation invariance: relabelling the vertices maps the role onto an entry with the same coefficients
            for perm in permutations(range(4)):
                pp = dof_of([perm[v] for v in dof_verts(p)])
                qq = dof_of([perm[v] for v in dof_verts(q)])
                kp = kcoef(pp, qq)
                mapped = {(min(perm[x], perm[y]), max(perm[x], perm[y])): c for (x, y), c in kc.items()}
                assert kp == mapped, (name, p, q, perm)
                assert pint(pmul(phi[pp], phi[qq]), 4) == pint(pmul(phi[p], phi[q]), 4)
            # group equal coefficients: c1 * (S[i] + S[j]) + ...
            bycoef = {}
            for key, c in kc.items():
                bycoef.setdefault(c, []).append(sums.index(key))
            terms = []
            for c, idx in sorted(bycoef.items(), key=lambda t: min(t[1])):
                inner = " + ".join(f"S[{i}]" for i in sorted(idx))
                terms.append(f"{lit(c)} * ({inner})" if len(idx) > 1 else f"{lit(c)} * {inner}")
            body.append(f"    K[{r}] = {' + '.join(terms)};\n")
            mass.append(lit(pint(pmul(phi[p], phi[q]), 4)))
            nrole += 1
        sums_txt = ", ".join(f"g{x}{y}" for x, y in sums)
        roles_txt = ", ".join(f"({p},{q})" for p, q in roles)
        out.append(f"// {name}: S = sums over the star of [{sums_txt}] (local vertices of the simplex in canonical order), K = entries {roles_txt} of the canonical frame\n")
        out.append(f"static WAE_HD inline void wae_{tag.lower()}_star_{name}(const double* S, double* K) {{\n{''.join(body)}}}\n")
    out.append(f"// mass coefficient of every role (times the star's sum of |det|), roles in the order vert, edge, face, tet\n")
    fill = "".join(f"    m[{i}] = {v};\n" for i, v in enumerate(mass))
    out.append(f"static WAE_HD inline void wae_{tag.lower()}_star_mass(double* m) {{\n{fill}}}\n")
    out.append(f"#define WAE_{tag}_STAR_ROLES {nrole}\n")
    return "".join(out)


def main():
    out = ["// GENERATED by tools/gen_fem_tables.py -- do not edit.\n#pragma once\n#ifdef __CUDACC__\n#define WAE_HD __host__ __device__\n#else\n#define WAE_HD\n#endif\ntemplate <int S>\nstruct WaeIdx {\n  static constexpr int value = S;\n};\n"]
    for order, tag in ((1, "P1"), (2, "P2")):
        phi = basis(order, 4)
        n = len(phi)
        out.append(f"// ---- tetrahedron, {tag}: n_loc = {n}\n")
        out.append(sym_table(f"WAE_{tag}_TET_MASS", n, lambda i, j: pint(pmul(phi[i], phi[j]), 4)))
        out.append(f"static __device__ const double WAE_{tag}_TET_SRC[{n}] = {{{', '.join(lit(pint(p, 4)) for p in phi)}}};\n")
        dphi = [[pdiff(p, a) for a in range(4)] for p in phi]
        pairs = [(a, b) for a in range(4) for b in range(a, 4)]
        # K[i][j] = sum_ab S_ijab g_ab, g symmetric: fold (a,b)+(b,a) on the 10 unique g's
        body = []
        for i in range(n):
            for j in range(i, n):
                terms = []
                for q, (a, b) in enumerate(pairs):
                    c = pint(pmul(dphi[i][a], dphi[j][b]), 4)
                    if a != b:
                        c += pint(pmul(dphi[i][b], dphi[j][a]), 4)
                    if c != 0:
                        terms.append(f"{lit(c)} * g[{q}]")
                body.append(f"    K[{i * n + j}] = {' + '.join(terms) if terms else '0.0'};\n")
                if i != j:
                    body.append(f"    K[{j * n + i}] = K[{i * n + j}];\n")
        out.append(f"// full {n}x{n} stiffness (row-major, symmetric) from g[q] = grad(l_a).grad(l_b) * scale, q over a<=b\n")
        out.append(f"static __device__ __forceinline__ void wae_{tag.lower()}_tet_stiff(const double* __restrict__ g, double* __restrict__ K) {{\n{''.join(body)}}}\n")
        # visitor over the packed upper triangle: f(integral_constant<int, s>, K_s, mass coefficient of s); every expression sits
        # at its use, so a kernel that consumes the entries one by one keeps only g[] live (pair-program assembly kernel)
        vbody = []
        idx = 0
        for i in range(n):
            for j in range(i, n):
                terms = []
                for q, (a, b) in enumerate(pairs):
                    c = pint(pmul(dphi[i][a], dphi[j][b]), 4)
                    if a != b:
                        c += pint(pmul(dphi[i][b], dphi[j][a]), 4)
                    if c != 0:
                        terms.append(f"{lit(c)} * g[{q}]")
                vbody.append(f"    f(WaeIdx<{idx}>{{}}, {' + '.join(terms) if terms else '0.0'}, {lit(pint(pmul(phi[i], phi[j]), 4))});\n")
                idx += 1
        out.append(f"template <class F>\nstatic __device__ __forceinline__ void wae_{tag.lower()}_tet_sym_visit(const double* __restrict__ g, F&& f) {{\n{''.join(vbody)}}}\n")
        # linear-c mass: M_ij = |det| sum_k c_k m_ijk
        lam = basis(1, 4)
        mc = [lit(pint(pmul(pmul(phi[i], phi[j]), lam[k]), 4)) for i in range(n) for j in range(n) for k in range(4)]
        out.append(f"static __device__ const double WAE_{tag}_TET_MASSC[{n * n * 4}] = {{{', '.join(mc)}}};\n")
        # linear-c stiffness: K_ij = sum_ab g_ab sum_{k<=l} c_k c_l T_ijabkl (folded on unique ab and kl)
        scc = []
        for i in range(n):
            for j in range(n):
                for (a, b) in pairs:
                    gab = pmul(dphi[i][a], dphi[j][b])
                    if a != b:
                        g2 = pmul(dphi[i][b], dphi[j][a])
                        gab = {e: gab.get(e, 0) + g2.get(e, 0) for e in set(gab) | set(g2)}
                    for (k, l) in pairs:
                        c = pint(pmul(gab, pmul(lam[k], lam[l])), 4) * (1 if k == l else 2)
                        scc.append(lit(c))
        out.append(f"static __device__ const double WAE_{tag}_TET_STIFFCC[{n * n * 100}] = {{{', '.join(scc)}}};\n")
        out.append(star_section(order, tag, phi, dphi))
        # triangles
        phit = basis(order, 3)
        nt = len(phit)
        out.append(f"// ---- triangle, {tag}: n_loc = {nt}\n")
        out.append(sym_table(f"WAE_{tag}_TRI_MASS", nt, lambda i, j: pint(pmul(phit[i], phit[j]), 3)))
        lam3 = basis(1, 3)
        mc = [lit(pint(pmul(pmul(phit[i], phit[j]), lam3[k]), 3)) for i in range(nt) for j in range(nt) for k in range(3)]
        out.append(f"static __device__ const double WAE_{tag}_TRI_MASSC[{nt * nt * 3}] = {{{', '.join(mc)}}};\n")
        out.append(f"static __device__ const double WAE_{tag}_TRI_SRC[{nt}] = {{{', '.join(lit(pint(p, 3)) for p in phit)}}};\n")
        # linear-c source (speaker): s_i = |det| sum_k c_k int(phi_i lam_k)
        sc = [lit(pint(pmul(phit[i], lam3[k]), 3)) for i in range(nt) for k in range(3)]
        out.append(f"static __device__ const double WAE_{tag}_TRI_SRCC[{nt * 3}] = {{{', '.join(sc)}}};\n")
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "wavesandeigenvalues.jl_b200", "csrc", "fem_gen.h")
    open(dst, "w").write("".join(out))
    print("wrote", os.path.normpath(dst))


if __name__ == "__main__":
    main()
