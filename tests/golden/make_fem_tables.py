"""Generate tests/golden/fem_tables.npz from the reference's own element expressions.

Runs ONLY in the build container (reads /root/reference/src/FEM/FEM.jl, which
does not exist on the GPU box).  Julia is not installed, so the reference's
element routines -- literal tables and machine-generated polynomials in
A=inv*inv', cc=c*c', c1..c4, x,y,z -- are evaluated by translating their
assignment lines to Python with 1-based index shims.  No reference source is
copied into the repo: only the numeric inputs/outputs are stored.

    python tests/golden/make_fem_tables.py
"""
import os
import re
import sys

import numpy as np

REF = "/root/reference/src/FEM/FEM.jl"


class OneBased:
    def __init__(self, shape, dtype=complex):
        self.a = np.zeros(shape, dtype=dtype)

    def _ix(self, k):
        if isinstance(k, tuple):
            return tuple(i - 1 for i in k)
        return k - 1

    def __getitem__(self, k):
        return self.a[self._ix(k)]

    def __setitem__(self, k, v):
        self.a[self._ix(k)] = v


def func_body(src, name):
    m = re.search(r"^function " + re.escape(name) + r"\(.*?\)\s*$", src, re.M)
    assert m, name
    start = m.end()
    # body ends at the first line that is exactly 'end' at column 0
    e = re.search(r"^end\s*$", src[start:], re.M)
    return src[start : start + e.start()]


def eval_literal(body):
    """First [...] literal of the body -> 2-D float array (rows split by ';')."""
    m = re.search(r"\[(.*?)\]", body, re.S)
    txt = m.group(1).replace("\n", " ")
    rows = [r.split() for r in txt.split(";") if r.strip()]
    return np.array([[eval(x) for x in r] for r in rows], dtype=float)


def statements(body):
    """Assignment statements 'M[...]=expr'; the reference wraps a few of them over two lines."""
    cur = None
    for line in body.splitlines():
        line = line.strip()
        if re.match(r"^M\[[0-9, ]+\]\s*=", line):
            if cur:
                yield cur
            cur = line
        elif cur is not None and line and not re.match(r"^(return|end|for|cc|M=|#)", line):
            cur += line
        elif cur is not None:
            yield cur
            cur = None
    if cur:
        yield cur


def eval_assign(body, ns):
    for st in statements(body):
        exec(st, {}, ns)


def main():
    src = open(REF).read()
    rng = np.random.default_rng(20261018)
    out = {}
    ncase = 4
    X = rng.standard_normal((ncase, 3, 4))
    X[1] *= 1e-3
    C4 = 300.0 + 400.0 * rng.random((ncase, 4))
    XT = rng.standard_normal((ncase, 3, 3))
    C3 = 300.0 + 400.0 * rng.random((ncase, 3))
    nref = rng.standard_normal((ncase, 3))
    nref /= np.linalg.norm(nref, axis=1, keepdims=True)
    lam = rng.random((ncase, 4))
    lam /= lam.sum(axis=1, keepdims=True)
    out.update(X=X, C4=C4, XT=XT, C3=C3, nref=nref)

    def trafo(Xc):
        d, m = Xc.shape
        J = np.empty((3, 3))
        J[:, : m - 1] = Xc[:, :-1] - Xc[:, -1:]
        if m == 3:
            n = np.cross(J[:, 0], J[:, 1])
            J[:, 2] = n / np.linalg.norm(n)
        return J, np.linalg.inv(J), np.linalg.det(J), Xc[:, -1]

    # constant tables ---------------------------------------------------------
    for name in ["s33v1u1", "s33v2u2", "s43v1u1", "s43v2u2", "s43v1", "s43v2", "s33v1", "s33v2"]:
        out["tab_" + name] = eval_literal(func_body(src, name))

    # tet polynomials -----------------------------------------------------------
    xref = np.zeros((ncase, 3))
    for name, n in [("s43nv1nu1", 4), ("s43nv2nu2", 10), ("s43nv1nu1cc1", 4), ("s43nv2nu2cc1", 10),
                    ("s43v1u1c1", 4), ("s43v2u2c1", 10), ("s43nv2rx", 10)]:
        body = func_body(src, name)
        res = []
        for k in range(ncase):
            J, Ji, det, orig = trafo(X[k])
            A = OneBased((3, 3), float)
            A.a[:] = Ji @ Ji.T
            cc = OneBased((4, 4), float)
            cc.a[:] = np.outer(C4[k], C4[k])
            c1, c2, c3, c4 = C4[k]
            if name == "s43nv2rx":
                M = OneBased((10, 3), float)
                xr = orig + J @ lam[k, :3]
                xref[k] = xr
                x, y, z = Ji @ (xr - orig)
                eval_assign(body, dict(M=M, x=x, y=y, z=z))
                res.append(M.a @ Ji @ nref[k])
                continue
            M = OneBased((n, n), float)
            eval_assign(body, dict(M=M, A=A, cc=cc, c1=c1, c2=c2, c3=c3, c4=c4))
            res.append(M.a * abs(det))
        out["val_" + name] = np.array(res)
    out["xref"] = xref
    # s43nv1rx is a fixed 4x3 selector times inv times n_ref (FEM.jl:2442-2448)
    res = []
    for k in range(ncase):
        J, Ji, det, orig = trafo(X[k])
        Msel = np.vstack([np.eye(3), -np.ones((1, 3))])
        res.append(Msel @ Ji @ nref[k])
    out["val_s43nv1rx"] = np.array(res)

    # triangle polynomials --------------------------------------------------------
    for name, n in [("s33v1u1c1", 3), ("s33v2u2c1", 6)]:
        body = func_body(src, name)
        res = []
        for k in range(ncase):
            J, Ji, det, orig = trafo(XT[k])
            c1, c2, c4 = C3[k]
            M = OneBased((n, n), float)
            eval_assign(body, dict(M=M, c1=c1, c2=c2, c4=c4))
            res.append(M.a * abs(det))
        out["val_" + name] = np.array(res)
    for name, n in [("s33v1c1", 3), ("s33v2c1", 6)]:
        body = func_body(src, name)
        res = []
        for k in range(ncase):
            J, Ji, det, orig = trafo(XT[k])
            c1, c2, c4 = C3[k]
            M = OneBased((n,), float)
            eval_assign(body, dict(M=M, c1=c1, c2=c2, c4=c4))
            res.append(M.a * abs(det))
        out["val_" + name] = np.array(res)

    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fem_tables.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    sys.exit(main())
