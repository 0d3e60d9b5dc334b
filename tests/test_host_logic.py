"""CPU tests of the product's host side (no GPU): mesh numbering rules, scalar algebra, family evaluation
rules and the symbolic LU phase, each against the oracle restatement of the reference."""
import ctypes as C
import os
import math

import numpy as np
import pytest

import wae_b200 as W
from cases import load_raw_mesh, speedofsound
from oracle import mesh as omesh
from oracle import nlevp as onlevp
from wae_b200 import nlevp


@pytest.fixture(scope="module")
def meshes():
    raw = load_raw_mesh("rijke_mm")
    return W.Mesh("m", scale=0.001, raw=raw), omesh.Mesh("m", scale=0.001, raw=raw)


def test_simplex_ordering_and_domains(meshes):
    mg, mo = meshes
    assert np.array_equal(mg.tetrahedra, np.array(mo.tetrahedra))
    assert np.array_equal(mg.triangles, np.array(mo.triangles))
    for dom in mo.domains:
        assert list(mg.domains[dom]["simplices"]) == list(mo.domains[dom]["simplices"]), dom
        assert mg.domains[dom]["dimension"] == mo.domains[dom]["dimension"]


def test_quadratic_dof_numbering(meshes):
    mg, mo = meshes
    tg, tt, dg = W.aggregate_elements(mg, "quad")
    to, tto, do = omesh.aggregate_elements(mo, "quad")
    assert dg == do
    assert np.array_equal(tt, np.array(tto)) and np.array_equal(tg, np.array(to))
    assert np.array_equal(mg.lines, np.array(mo.lines))


def test_mesh_queries(meshes):
    mg, mo = meshes
    assert np.array_equal(mg.link_triangles_to_tetrahedra(), mo.link_triangles_to_tetrahedra())
    for dom in ("Flame", "Outlet", "Interior"):
        assert abs(mg.compute_size(dom) - mo.compute_size(dom)) <= 1e-13 * mo.compute_size(dom)
    for p in ([0.0, 0.0, -0.00101], [0.01, 0.005, 0.1], [1.0, 1.0, 1.0]):
        assert mg.find_tetrahedron_containing_point(p) == mo.find_tetrahedron_containing_point(p)
    assert np.array_equal(mg.generate_field(speedofsound), mo.generate_field(speedofsound))


def test_ragged_and_duplicate_input():
    """duplicated simplices and an empty domain (Mesh constructor, Meshutils.jl:118-147)."""
    pts = np.array([[0, 1, 0, 0, 1.0], [0, 0, 1, 0, 1.0], [0, 0, 0, 1, 1.0]])
    tets = [[0, 1, 2, 3], [3, 2, 1, 0], [1, 2, 3, 4]]
    tris = [[0, 1, 2], [2, 1, 0]]
    dom = {"A": {"dimension": 3, "simplices": [0, 1, 2]}, "E": {"dimension": 2, "simplices": []}, "T": {"dimension": 2, "simplices": [1, 0]}}
    mg = W.Mesh("x", raw=(pts, [], tris, tets, dom))
    mo = omesh.Mesh("x", raw=(pts, [], tris, tets, dom))
    assert np.array_equal(mg.tetrahedra, np.array(mo.tetrahedra)) and len(mg.tetrahedra) == 2
    assert list(mg.domains["A"]["simplices"]) == mo.domains["A"]["simplices"]
    assert list(mg.domains["T"]["simplices"]) == mo.domains["T"]["simplices"] == [0]
    assert len(mg.domains["E"]["simplices"]) == 0


def test_scalar_algebra():
    rng = np.random.default_rng(0)
    for _ in range(20):
        z = complex(rng.standard_normal(), rng.standard_normal()) * 100
        tau = complex(rng.random() * 1e-3)
        for k in range(4):
            assert nlevp.pow0(z, k) == onlevp.pow0(z, k)
            assert nlevp.pow1(z, k) == onlevp.pow1(z, k)
            assert nlevp.pow2(z, k) == onlevp.pow2(z, k)
        for m in range(4):
            for n in range(4):
                a, b = nlevp.exp_delay(z, tau, m, n), onlevp.exp_delay(z, tau, m, n)
                assert abs(a - b) <= 1e-15 * abs(b)
    # derivative consistency of exp_delay by finite differences
    z, tau, h = 500.0 + 3j, 1e-3 + 0j, 1e-4
    fd = (nlevp.exp_delay(z + h, tau, 0, 0) - nlevp.exp_delay(z - h, tau, 0, 0)) / (2 * h)
    assert abs(fd - nlevp.exp_delay(z, tau, 1, 0)) < 1e-8


def test_update_formulas_and_pade():
    rng = np.random.default_rng(1)
    for o in range(1, 6):
        f = list(rng.standard_normal(o + 1) + 1j * rng.standard_normal(o + 1))
        assert abs(nlevp.householder_update(f) - onlevp.householder_update(f)) < 1e-14
    w = list(rng.standard_normal(6) + 1j * rng.standard_normal(6))
    for Lo, M in ((1, 0), (1, 2), (2, 3)):
        a, b = nlevp.pade(w, Lo, M)
        ao, bo = onlevp.pade(w, Lo, M)
        assert np.allclose(a, ao) and np.allclose(b, bo)
    assert np.allclose(np.sort_complex(nlevp.poly_roots([2.0, -3.0, 1.0])), [1.0, 2.0])
    G = [0, 2, 2 + 2j, 2j]
    assert nlevp.wn(1 + 1j, G) == onlevp.wn(1 + 1j, G) == 1 and nlevp.wn(3 + 1j, G) == 0
    z1, w1 = nlevp.contour_nodes(G, 8)
    z2, w2 = onlevp.contour_nodes(G, 8)
    assert np.array_equal(z1, z2) and np.array_equal(w1, w2)
    assert abs(w1.sum()) < 1e-14  # closed contour


class _FakeMat:
    dim = 5
    parts = [(0, 1.0)]
    ctx = None


def _families():
    fams = []
    for mod in (nlevp, onlevp):
        L = mod.LinearOperatorFamily(["ω", "λ"], [0.0, float("inf")])
        coeff = _FakeMat() if mod is nlevp else __import__("scipy.sparse").sparse.identity(5, dtype=complex, format="csc")
        for func, params, op in (((mod.pow2,), (("ω",),), "M"), ((), (), "K"), ((mod.pow1, mod.pow1), (("ω",), ("Y",)), "C"),
                                 ((mod.pow1, mod.exp_delay), (("n",), ("ω", "τ")), "Q"), ((mod.pow1,), (("λ",),), "__aux__")):
            L.terms.append(mod.Term(coeff, func, params, "", op))
        L.params.update({"Y": 1e15 + 0j, "n": 0.5 + 0j, "τ": 1e-3 + 0j})
        fams.append(L)
    return fams


def test_family_scalar_rules_match_oracle():
    """Which terms enter L(z,...) and with which scalar (LinOpFam.jl:501-526), all three modes."""
    Lg, Lo = _families()
    for L in (Lg, Lo):
        L.params["ω"] = 300.0 + 5j
        L.params["λ"] = 0.25 + 0j
    for mode, active, derivs_list in (("all", ["ω"], [[0], [1], [2], [3]]),
                                      ("householder", ["λ", "ω"], [[0, 0], [1, 0], [0, 1], [0, 2], [1, 1], [2, 0]]),
                                      ("compact", ["ω", "τ"], [[0, 0], [1, 1], [0, 3]])):
        for L in (Lg, Lo):
            L.mode, L.active = mode, list(active)
        for derivs in derivs_list:
            sg, so = Lg.scalars(derivs), Lo.scalars(derivs)
            assert [x is None for x in sg] == [x is None for x in so], (mode, derivs)
            for a, b in zip(sg, so):
                if a is not None:
                    assert abs(a - b) <= 1e-15 * max(abs(b), 1e-300), (mode, derivs)


def test_partitions_match_oracle():
    for n in range(1, 8):
        assert [list(p) for p in nlevp._partitions(n)] == [list(p) for p in onlevp.partitions(n)]
        assert len(list(nlevp._partitions(n))) == [1, 2, 3, 5, 7, 11, 15][n - 1]


def test_symbolic_lu_phase_on_host(meshes):
    """wae_lu_symbolic_stats runs without a GPU: valid permutation, bounded fill on the P2 Rijke pattern."""
    import scipy.sparse as sp
    from wae_b200 import _lib
    mg, _ = meshes
    _, tets, dim = W.aggregate_elements(mg, "quad")
    n = tets.shape[1]
    I = np.repeat(tets, n, axis=1).ravel()
    J = np.tile(tets, (1, n)).ravel()
    A = sp.csc_matrix((np.ones(len(I)), (I, J)), shape=(dim, dim))
    A.sum_duplicates()
    A.sort_indices()
    lib = _lib.lib()
    lib.wae_lu_symbolic_stats.restype = C.c_int32
    lib.wae_lu_symbolic_stats.argtypes = [C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_double), C.c_int32,
                                          C.POINTER(C.c_double)]
    cp, rv = A.indptr.astype(np.int64), A.indices.astype(np.int64)
    out = np.zeros(8)
    for coords in (None, np.ascontiguousarray(np.random.default_rng(0).random((dim, 3)))):
        rc = lib.wae_lu_symbolic_stats(dim, cp.ctypes.data_as(C.POINTER(C.c_int64)), rv.ctypes.data_as(C.POINTER(C.c_int64)),
                                       None if coords is None else coords.ctypes.data_as(C.POINTER(C.c_double)), 64,
                                       out.ctypes.data_as(C.POINTER(C.c_double)))
        assert rc == 0
        assert out[1] >= A.nnz and out[0] >= 1
    # with the BFS key (no coordinates) the fill of this 6172-DOF tube must stay moderate
    rc = lib.wae_lu_symbolic_stats(dim, cp.ctypes.data_as(C.POINTER(C.c_int64)), rv.ctypes.data_as(C.POINTER(C.c_int64)), None, 64,
                                   out.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == 0 and out[1] < 30 * A.nnz


def _pair_program_check(mesh, order, slot_cap):
    from wae_b200 import _lib
    _, tets, dim = W.aggregate_elements(mesh, order)
    lib = _lib.lib()
    lib.wae_pair_program_check.restype = C.c_int32
    lib.wae_pair_program_check.argtypes = [C.c_int32, C.c_int64, C.POINTER(C.c_double), C.c_int64, C.POINTER(C.c_uint32), C.c_int32,
                                           C.POINTER(C.c_double)]
    xyz = np.ascontiguousarray(mesh.points.T, dtype=np.float64)
    t32 = np.ascontiguousarray(tets, dtype=np.uint32)
    out = np.zeros(8)
    rc = lib.wae_pair_program_check(1 if order == "lin" else 2, xyz.shape[0], xyz.ctypes.data_as(C.POINTER(C.c_double)), len(t32),
                                    t32.ctypes.data_as(C.POINTER(C.c_uint32)), slot_cap, out.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == 0
    nsym = 10 if order == "lin" else 55
    return dict(nnz=int(out[0]), patches=int(out[1]), staged=int(out[2]), sources=int(out[3]), units=int(out[4]), max_slots=int(out[5]),
                err=out[6], bad=int(out[7]), ntet=len(t32), nsym=nsym)


def test_assembly_pair_program_on_host(meshes):
    """The owner-computes pair program of the M/K assembly kernel, replayed on the host (no GPU): every nonzero is written exactly
    once, every slot is used once, the fixed-order sums equal the plain triplet sums; on the unstructured Rijke mesh (P1 and P2) and
    on a structured Kuhn box cut into many patches and into a single patch."""
    mg, _ = meshes
    box = W.kuhn_box((6, 5, 7), (0, 0, 0), (1, 1, 1), jitter=0.1, seed=5)
    for mesh, order, cap in ((mg, "lin", 6000), (mg, "quad", 7168), (mg, "quad", 14000), (box, "quad", 5000), (box, "quad", 4096), (box, "quad", 3072), (box, "lin", 14000)):  # incl. the slot budgets of the multi-CTA layouts bench.py times
        r = _pair_program_check(mesh, order, cap)
        assert r["bad"] == 0 and r["err"] < 1e-12, r
        assert r["max_slots"] <= cap
        # a source is owned at least once and at most twice (once per owning column patch)
        assert r["nsym"] * r["ntet"] <= r["sources"] <= 2 * r["nsym"] * r["ntet"]
        assert r["staged"] >= r["ntet"] and r["units"] <= r["nnz"]
    one = _pair_program_check(box, "lin", 14000)
    assert one["patches"] == 1 and one["staged"] == one["ntet"] and one["sources"] == 10 * one["ntet"]


def _star_program_check(mesh, order, budget, c):
    from wae_b200 import _lib
    _, tets, dim = W.aggregate_elements(mesh, order)
    lib = _lib.lib()
    pd, pi64 = C.POINTER(C.c_double), C.POINTER(C.c_int64)
    lib.wae_star_program_check.restype = C.c_int32
    lib.wae_star_program_check.argtypes = [C.c_int32, C.c_int64, pd, C.c_int64, C.POINTER(C.c_uint32), pd, C.c_int64, C.c_int64, pi64,
                                           C.POINTER(C.c_int32), pd, pd, pd]
    xyz = np.ascontiguousarray(mesh.points.T, dtype=np.float64)
    t32 = np.ascontiguousarray(tets, dtype=np.uint32)
    nloc = t32.shape[1]
    cap = len(t32) * nloc * nloc
    colptr, rowval = np.zeros(dim + 1, dtype=np.int64), np.zeros(cap, dtype=np.int32)
    vm, vk, st = np.zeros(cap), np.zeros(cap), np.zeros(16)
    cc = np.ascontiguousarray(c, dtype=np.float64)
    rc = lib.wae_star_program_check(1 if order == "lin" else 2, xyz.shape[0], xyz.ctypes.data_as(pd), len(t32), t32.ctypes.data_as(C.POINTER(C.c_uint32)),
                                    cc.ctypes.data_as(pd), budget, cap, colptr.ctypes.data_as(pi64), rowval.ctypes.data_as(C.POINTER(C.c_int32)),
                                    vm.ctypes.data_as(pd), vk.ctypes.data_as(pd), st.ctypes.data_as(pd))
    assert rc == 0
    nnz = int(st[0])
    return dict(nnz=nnz, patches=int(st[1]), staged=int(st[2]), simplices=int(st[3]), sources=int(st[4]), smem=int(st[5]), program=int(st[6]),
                bad=int(st[7]), wavefronts=st[8:12].copy(), ntet=len(t32), colptr=colptr, rowval=rowval[:nnz], vm=vm[:nnz], vk=vk[:nnz], tets=t32, xyz=xyz, dim=dim)


def _oracle_mk(xyz, tets, c, order, dim):
    """M and K = -c^2 int grad.grad as sparse() of the element triplets, from the oracle's exact tables (vectorised over the elements)."""
    import scipy.sparse as sp
    from oracle import fem
    T = fem.tables(1 if order == "lin" else 2, 4)
    X = xyz[tets[:, :4].astype(np.int64)]                       # n x 4 x 3
    J = np.transpose(X[:, :3, :] - X[:, 3:4, :], (0, 2, 1))       # columns = edges
    inv = np.linalg.inv(J)
    det = np.abs(np.linalg.det(J))
    G = np.concatenate([inv, -inv.sum(axis=1, keepdims=True)], axis=1)
    GG = np.einsum("nad,nbd->nab", G, G)
    Ke = np.einsum("ijab,nab->nij", np.asarray(T["stiff"], dtype=float), GG) * (-(c ** 2) * det)[:, None, None]
    Me = np.asarray(T["mass"], dtype=float)[None] * det[:, None, None]
    nloc = tets.shape[1]
    I = np.repeat(tets.astype(np.int64), nloc, axis=1).ravel()
    Jc = np.tile(tets.astype(np.int64), (1, nloc)).ravel()
    M = sp.coo_matrix((Me.ravel(), (I, Jc)), shape=(dim, dim)).tocsc()
    K = sp.coo_matrix((Ke.ravel(), (I, Jc)), shape=(dim, dim)).tocsc()
    for A in (M, K):
        A.sum_duplicates()
        A.sort_indices()
    return M, K


def test_assembly_star_program_on_host(meshes):
    """The star program of the third-generation M/K assembly kernel, replayed on the host with the kernel's own arithmetic (no GPU):
    pattern identical to sparse() of the element triplets, every nonzero written exactly once, values equal to the oracle's element
    tables summed per nonzero; unstructured Rijke mesh (P1, P2) and a jittered Kuhn box cut into many patches, few patches, one patch."""
    mg, _ = meshes
    box = W.kuhn_box((6, 5, 7), (0, 0, 0), (1, 1, 1), jitter=0.1, seed=5)
    rng = np.random.default_rng(3)
    for mesh, order, budget in ((mg, "lin", 40 * 1024), (mg, "quad", 112 * 1024), (mg, "quad", 60 * 1024), (box, "quad", 224 * 1024),
                                (box, "quad", 112 * 1024), (box, "quad", 48 * 1024), (box, "lin", 224 * 1024), (box, "lin", 24 * 1024)):
        ntet = len(mesh.tetrahedra)
        c = 300.0 + 400.0 * rng.random(ntet)
        r = _star_program_check(mesh, order, budget, c)
        assert r["bad"] == 0, {k: v for k, v in r.items() if np.isscalar(v)}
        assert r["smem"] <= budget
        M, K = _oracle_mk(r["xyz"], r["tets"], c, order, r["dim"])
        assert np.array_equal(r["colptr"], M.indptr) and np.array_equal(r["rowval"], M.indices)
        assert np.abs(r["vm"] - M.data).max() <= 1e-12 * np.abs(M.data).max()
        assert np.abs(r["vk"] - K.data).max() <= 1e-12 * np.abs(K.data).max()
        # every sub-simplex of every element is a source at least once; patches only repeat stars on their rim
        per = 10 if order == "lin" else 15
        assert per * ntet <= r["sources"] <= 4 * per * ntet and r["staged"] >= ntet
    small = W.kuhn_box((3, 3, 4), (0, 0, 0), (1, 1, 1), jitter=0.1, seed=6)
    one = _star_program_check(small, "lin", 224 * 1024, np.full(len(small.tetrahedra), 340.0))
    assert one["patches"] == 1 and one["staged"] == one["ntet"] and one["sources"] == 10 * one["ntet"]


def test_kuhn_box_is_conforming():
    m = W.kuhn_box((3, 2, 4), (0, 0, 0), (3, 2, 4), jitter=0.1, seed=3, flame_layer=(1, 2))
    X = m.points[:, m.tetrahedra]
    vol = np.abs(np.linalg.det(np.moveaxis(X[:, :, :3] - X[:, :, 3:4], 1, 0))).sum() / 6
    assert abs(vol - 24.0) < 1e-9  # jitter moves interior points only: the volume is preserved
    assert len(m.tetrahedra) == 3 * 2 * 4 * 6
    assert len(m.triangles) == 2 * 2 * (3 * 2 + 3 * 4 + 2 * 4)
    assert len(m.domains["Flame"]["simplices"]) == 3 * 2 * 6
    assert abs(m.compute_size("Outlet") - 6.0) < 1e-12


def _num_deriv(f, x, order, h):
    """central finite-difference derivative of the given order (complex step sizes are fine: f is analytic)."""
    if order == 0:
        return f(x)
    return (_num_deriv(f, x + h, order - 1, h) - _num_deriv(f, x - h, order - 1, h)) / (2 * h)


def test_fancyflame_and_state_space_scalars():
    """exp_az2mzit (algebra.jl:255-274), Σnexp_az2mzit (:313-325), exp_ax2 (:229-253) and generate_stsp_z (:158-167): the product's
    functions agree with the oracle's restatement and with finite differences of the closed form exp(a z^2 - i z tau)."""
    z, tau, a = 900.0 + 40.0j, 1.3e-3 + 0j, -2.0e-7 + 0j
    F = lambda z_, t_, a_: np.exp(a_ * z_ * z_ - 1j * z_ * t_)
    for m, n, k in ((0, 0, 0), (1, 0, 0), (2, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (1, 0, 1), (2, 1, 1), (3, 2, 0)):
        g = nlevp.exp_az2mzit(z, tau, a, m, n, k)
        o = onlevp.exp_az2mzit(z, tau, a, m, n, k)
        assert abs(g - o) <= 1e-12 * abs(o)
        # d^n/dtau^n d^k/da^k of the closed form is (-i z)^n z^(2k) F; the m z-derivatives by central differences
        fd = _num_deriv(lambda zz: (-1j * zz) ** n * zz ** (2 * k) * F(zz, tau, a), z, m, 0.5)
        assert abs(g - fd) <= 2e-5 * abs(fd) + 1e-30, (m, n, k)
    for nn in range(5):
        assert abs(nlevp.exp_ax2(z, a, nn) - onlevp.exp_ax2(z, a, nn)) <= 1e-12 * abs(onlevp.exp_ax2(z, a, nn))
    args = (z, 1.5 + 0j, tau, a, 0.5 + 0j, 2 * tau, 3 * a, 1, 0, 0, 0, 0, 0, 0)
    s = nlevp.Sigma_nexp_az2mzit(*args)
    assert abs(s - onlevp.sigma_nexp_az2mzit(*args)) <= 1e-12 * abs(s)
    assert abs(s - (1.5 * nlevp.exp_az2mzit(z, tau, a, 1, 0, 0) + 0.5 * nlevp.exp_az2mzit(z, 2 * tau, 3 * a, 1, 0, 0))) <= 1e-12 * abs(s)
    A, B, Cm, D = np.array([[-50.0, 20.0], [-20.0, -80.0]]), np.array([1.0, 0.5]), np.array([2.0, -1.0]), np.array([0.3])
    fg, fo = nlevp.generate_stsp_z(A, B, Cm, D), onlevp.generate_stsp_z(A, B, Cm, D)
    H = lambda w: (Cm @ np.linalg.solve(1j * w * np.eye(2) - A, B)) + 0.3
    for n in range(4):
        assert abs(fg(z, n) - fo(z, n)) <= 1e-12 * abs(fo(z, n))
        if n <= 2:  # (finite differences of higher order drown in round-off: the values are ~1e-11)
            assert abs(fg(z, n) - _num_deriv(H, z, n, 2.0)) <= 1e-4 * abs(fg(z, n)) + 1e-30


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver times beside the GPU arm) prints one JSON line with the contract's keys; its
    value is MEASURED on the bounded sample (no extrapolation), its `config` is the GPU arm's, and it never loads libwae_b200.so."""
    import json
    import subprocess
    import sys
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0', '--cpu-sample', '2,2,8']; "
            "runpy.run_path('bench.py', run_name='__main__'); "
            "print('NATIVE', [l.split()[-1] for l in open('/proc/self/maps') if 'libwae_b200' in l])")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["unit"] == "eigenpairs/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["extrapolated"] is False
    # measured: one step of the sample took ms_per_step, and that is the value
    assert abs(d["value"] * d["ms_per_step"] * 1e-3 - 1) < 1e-9 and cb["value"] == d["value"]
    assert "scaling" not in cb and cb["sample_dofs"] == 5 * 5 * 17
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    # the same config object as the GPU arm (the driver compares them)
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert d["config"] == bench.workload_config((20, 20, 300), (2, 2, 8))
    assert [l for l in out.stdout.splitlines() if l.startswith("NATIVE")][-1] == "NATIVE []"


@pytest.mark.parametrize("name,scale", [("rijke_mm", 0.001), ("ntnu_12", 1.0)])
def test_octosplit_matches_oracle(name, scale):
    """octosplit (Meshutils.jl:589-747): the vectorised product against the loop restatement -- points, sorted child lists,
    domain relabelling -- plus the size-independent properties: 8x / 4x counts, volume and surface area conserved, children
    of a boundary triangle are faces of exactly one child tetrahedron, and the refined mesh is conforming."""
    raw = load_raw_mesh(name)
    mg, mo = W.Mesh("m", scale=scale, raw=raw), omesh.Mesh("m", scale=scale, raw=raw)
    sg, so = W.octosplit(mg), omesh.octosplit(mo)
    assert np.array_equal(sg.points, so.points)
    assert np.array_equal(sg.tetrahedra, np.array(so.tetrahedra)) and np.array_equal(sg.triangles, np.array(so.triangles))
    assert len(sg.tetrahedra) == 8 * len(mg.tetrahedra) and len(sg.triangles) == 4 * len(mg.triangles)
    assert sg.points.shape[1] == mg.points.shape[1] + len(mg.lines)
    assert np.array_equal(sg.points[:, : mg.points.shape[1]], mg.points)
    for dom, d in so.domains.items():
        assert list(sg.domains[dom]["simplices"]) == list(d["simplices"]), dom
    for dom, d in mg.domains.items():
        if d["dimension"] in (2, 3) and len(d["simplices"]):
            a, b = mg.compute_size(dom), sg.compute_size(dom)
            assert abs(a - b) <= 1e-12 * a, dom
    # conforming: every face belongs to one (boundary) or two (interior) tetrahedra, boundary faces = the triangles of the surface
    t = sg.tetrahedra
    faces = np.sort(np.concatenate([t[:, [0, 1, 2]], t[:, [0, 1, 3]], t[:, [0, 2, 3]], t[:, [1, 2, 3]]]), axis=1)
    uf, cnt = np.unique(faces, axis=0, return_counts=True)
    assert cnt.max() == 2
    t2t = sg.link_triangles_to_tetrahedra()
    assert (t2t >= 0).all()
    # second level: the numbering rules hold on the refined mesh as well (P2 DOFs of the split mesh)
    tg, tt, dim = W.aggregate_elements(sg, "quad")
    assert dim == sg.points.shape[1] + len(sg.lines) and tt.max() == dim - 1


def test_octosplit_tie_break_and_empty_surface():
    """Regular tetrahedron: all three inner diagonals have equal length, the reference takes AB-CD (first branch, '<=')."""
    pts = np.array([[1.0, 1, -1, -1], [1, -1, 1, -1], [1, -1, -1, 1]])
    dom = {"A": {"dimension": 3, "simplices": [0]}}
    mg, mo = W.Mesh("x", raw=(pts, [], [], [[0, 1, 2, 3]], dom)), omesh.Mesh("x", raw=(pts, [], [], [[0, 1, 2, 3]], dom))
    sg, so = W.octosplit(mg), omesh.octosplit(mo)
    assert np.array_equal(sg.tetrahedra, np.array(so.tetrahedra)) and len(sg.triangles) == 0
    ab, cd = 4 + mg.edge_index(np.array([0]), np.array([1]))[0], 4 + mg.edge_index(np.array([2]), np.array([3]))[0]
    assert sum(1 for t in sg.tetrahedra if ab in t and cd in t) == 4
    assert list(sg.domains["A"]["simplices"]) == list(range(8))


def test_native_simplex_numbering_equals_the_numpy_path():
    """wae_sorted_unique_simplices (csrc/mesh_symbolic.cpp, thread-parallel sort) against the vectorised numpy rule it replaces for large
    inputs: unique simplices, representatives (first occurrence, input vertex order) and the raw -> unique map, for edges, triangles and
    tetrahedra, with many repeated vertex sets, ids up to 2^31 and sizes on both sides of the threading threshold."""
    import numpy as np
    from wae_b200 import meshutils as mu
    rng = np.random.default_rng(11)
    for k in (2, 3, 4):
        for n, hi in ((1, 5), (7, 3), (3000, 40), (70000, 900), (70000, 2**31)):
            s = rng.integers(0, hi, size=(n, k))
            a, ia = mu._sorted_unique(s, native=False)
            b, ib = mu._sorted_unique(s, native=True)
            assert np.array_equal(a, b) and np.array_equal(ia, ib), (k, n, hi)
    # the six edges of every tetrahedron of a Kuhn box (collect_lines): what the P2 numbering of the big configurations is built from
    mesh = mu.kuhn_box((12, 9, 11), (0, 0, 0), (1, 1, 1))
    t = mesh.tetrahedra
    pairs = np.stack([t[:, [0, 1]], t[:, [0, 2]], t[:, [0, 3]], t[:, [1, 2]], t[:, [1, 3]], t[:, [2, 3]]], axis=1).reshape(-1, 2)
    a, ia = mu._sorted_unique(pairs, native=False)
    b, ib = mu._sorted_unique(pairs, native=True)
    nx, ny, nz = 12, 9, 11  # Kuhn triangulation: axis edges + one diagonal per cube face + one body diagonal per cube
    n_edges = (nx * (ny + 1) * (nz + 1) + (nx + 1) * ny * (nz + 1) + (nx + 1) * (ny + 1) * nz
               + nx * ny * (nz + 1) + nx * (ny + 1) * nz + (nx + 1) * ny * nz + nx * ny * nz)
    assert np.array_equal(a, b) and np.array_equal(ia, ib) and len(a) == n_edges
    # ids outside [0, 2^32) are rejected by the library and fall back to the numpy rule
    big = np.array([[2**33, 1], [1, 2**33], [0, 1]])
    c, ic = mu._sorted_unique(big, native=True)
    assert np.array_equal(c, np.array([[0, 1], [2**33, 1]])) and list(ic) == [1, 1, 0]
