"""Discrete-adjoint shape sensitivity (src/shape_sensitivity.jl:16-141), CPU part: the product's host logic (surface bookkeeping,
descriptor -> terms, per-point simplex lists, post-processing) and the per-thread function of the CUDA kernel replayed on the
host (wae_shape_sens_check) against the oracle's literal restatement (six discretize calls per point on reduced domains), and
the oracle itself against true finite differences of the eigenvalue.  The GPU launch of the same function is tested in
test_shape_sensitivity_gpu.py."""
import ctypes as C
import math
import os

import numpy as np
import pytest

import wae_b200 as W
from cases import GAMMA, N_REF, Q02U0, RHO, X_REF, load_raw_mesh, speedofsound
from oracle import helmholtz as ohelm
from oracle import mesh as omesh
from oracle import nlevp as onlevp
from oracle import shape as oshape
from wae_b200 import _lib, shape

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_pd, _pi64, _pu32 = C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_uint32)


def host_replay(mesh, dscrp, c, surface_points, tri_mask, tet_mask, w0, v0, va, h=1e-9):
    """The product's host logic + the kernel's per-thread function on the host (test-only entry of the library)."""
    lib = C.CDLL(os.path.join(ROOT, "wavesandeigenvalues.jl_b200", "libwae_b200.so"))
    f = lib.wae_shape_sens_check
    f.restype = C.c_int32
    f.argtypes = [C.c_int64, _pd, C.c_int64, _pu32, C.c_int64, _pu32, C.c_int64, _pi64, C.c_double, _pd, _pd, C.c_int32, _pi64, _pi64, _pd,
                  C.c_int32, _pd, C.c_int64, _pd, C.c_double, _pd]
    xyz = np.ascontiguousarray(mesh.points.T, dtype=np.float64)
    tets = np.ascontiguousarray(mesh.tetrahedra, dtype=np.uint32)
    tris = np.ascontiguousarray(mesh.triangles, dtype=np.uint32)
    pts = np.ascontiguousarray(surface_points, dtype=np.int64)
    v0 = np.ascontiguousarray(v0, dtype=np.complex128)
    va = np.ascontiguousarray(va, dtype=np.complex128)
    out = np.zeros((len(pts), 3), dtype=np.complex128)
    P = lambda a, t: None if a is None else a.ctypes.data_as(t)
    for term in shape.sensitivity_terms(mesh, dscrp, c, w0):
        n_elem = len(tris) if term["dim"] == 2 else len(tets)
        ptr, elems, cc = shape.sensitivity_lists(term, n_elem, tri_mask, tet_mask)
        cpe = 1
        if cc is not None:
            cc = np.ascontiguousarray(cc, dtype=np.float64)
            cpe = 1 if cc.ndim == 1 else cc.shape[1]
        coef = np.array([complex(term["coef"])])
        nref = None if "n_ref" not in term else np.ascontiguousarray(term["n_ref"], dtype=np.float64)
        rc = f(xyz.shape[0], P(xyz, _pd), len(tets), P(tets, _pu32), len(tris), P(tris, _pu32), len(pts), P(pts, _pi64), h, P(v0, _pd), P(va, _pd),
               term["kind"], P(ptr, _pi64), P(elems, _pi64), P(cc, _pd), cpe, P(coef, _pd), term.get("ref_tet", 0), P(nref, _pd),
               float(term.get("nl", 0.0)), P(out, _pd))
        assert rc == 0, (rc, term["kind"])
    sens = np.zeros((3, mesh.points.shape[1]), dtype=np.complex128)
    sens[:, pts] = out.T
    return sens


@pytest.fixture(scope="module")
def rijke():
    """tutorial set-up of examples/shape/tutorial_09_shape_sensitivity.jl:5-56 (hot case, n = 1, explicit ref_idx), without octosplit."""
    raw = load_raw_mesh("rijke_mm")
    mg, mo = W.Mesh("m", scale=0.001, raw=raw), omesh.Mesh("m", scale=0.001, raw=raw)
    c = mo.generate_field(speedofsound)
    ref_idx = mo.find_tetrahedron_containing_point(X_REF)
    dscrp = {"Interior": ("interior", ()), "Outlet": ("admittance", ("Y", 1e15)),
             "Flame": ("flame", (GAMMA, RHO, Q02U0, ref_idx, X_REF, N_REF, "n", "τ", 1.0, 0.001))}
    L = ohelm.discretize(mo, dscrp, c)
    sol, n, flag = onlevp.householder(L, 700 * 2 * math.pi, maxiter=14, tol=1e-11)
    assert flag == 1
    return mg, mo, c, dscrp, L, sol, ref_idx


def _normalised(L, sol):
    w0 = sol.params[sol.eigval]
    v0 = sol.v / np.sqrt(np.vdot(sol.v, sol.v))
    va = sol.v_adj / np.conj(np.vdot(sol.v_adj, L(w0, 1) @ v0))
    return w0, v0, va


def test_surface_bookkeeping_matches_oracle(rijke):
    mg, mo = rijke[0], rijke[1]
    sg, trg, ttg = W.get_surface_points(mg)
    so, tro, tto = oshape.get_surface_points(mo)
    assert list(sg) == so and len(sg) == 783
    assert all(list(a) == b for a, b in zip(trg, tro)) and all(list(a) == b for a, b in zip(ttg, tto))
    ng, no = W.get_normal_vectors(mg), oshape.get_normal_vectors(mo)
    assert np.allclose(ng, no, rtol=1e-13, atol=0)
    # closed surface: the area vectors sum to zero; outward: the divergence theorem gives 3 * volume
    assert np.abs(ng.sum(axis=1)).max() <= 1e-12 * np.abs(ng).sum()
    cen = mg.points[:, mg.triangles].mean(axis=2)
    vol = mg.compute_size("Interior")
    assert abs(np.einsum("ij,ij->", cen, ng) / 2 / 3 - vol) <= 1e-10 * vol


def test_kernel_function_matches_oracle_loop(rijke):
    """all 783 surface points x 3 coordinates; the two sides difference the same element matrices at h = 1e-9, so they agree to the
    round-off of that difference (the oracle differences the assembled sums, the kernel the element matrices)."""
    mg, mo, c, dscrp, L, sol, ref_idx = rijke
    sp_, trm, ttm = W.get_surface_points(mg)
    w0, v0, va = _normalised(L, sol)
    got = host_replay(mg, dscrp, c, sp_, trm, ttm, w0, v0, va)
    so, tro, tto = oshape.get_surface_points(mo)
    want = oshape.discrete_adjoint_shape_sensitivity(mo, dscrp, c, so, tro, tto, L, sol)
    scale = np.abs(want).max()
    assert scale > 1
    assert np.abs(got - want).max() <= 2e-6 * scale
    mag = np.abs(want[:, sp_]).max(axis=0)  # per point: values span 6 .. 5400, the finite-difference noise of the oracle is ~2e-3
    assert (np.abs(got - want)[:, sp_].max(axis=0) <= 2e-4 * np.maximum(mag, 1e-3 * scale)).all()
    assert np.count_nonzero(np.abs(want).sum(axis=0)) == len(sp_)
    # points touching the flame / the reference tetrahedron / the outlet are all present in the comparison
    flame = set(np.asarray(mg.tetrahedra)[mg.domains["Flame"]["simplices"]].ravel().tolist())
    outlet = set(np.asarray(mg.triangles)[mg.domains["Outlet"]["simplices"]].ravel().tolist())
    assert flame & set(sp_.tolist()) and outlet & set(sp_.tolist())


def test_linear_speed_of_sound_and_descriptor_variants(rijke):
    """per-point c (linear-c stiffness and boundary mass, FEM.jl:2283-2311, 469-483), a 9-tuple flame (reference tetrahedron by
    point location) and a functional admittance, on a subset of the points."""
    mg, mo, _, _, _, _, _ = rijke
    cpt = np.array([speedofsound(*mo.points[:, i]) * (1 + 0.1 * math.sin(40 * mo.points[2, i])) for i in range(mo.points.shape[1])])
    Yf = lambda w, k=0: (2.0 + 0.001j * w) if k == 0 else (0.001j if k == 1 else 0.0)
    dscrp = {"Interior": ("interior", ()), "Outlet": ("admittance", (Yf,)), "Inlet": ("admittance", ("Yin", 0.5)),
             "Flame": ("flame", (GAMMA, RHO, Q02U0, X_REF, N_REF, "n", "τ", 0.7, 0.0012))}
    L = ohelm.discretize(mo, dscrp, cpt)
    sol, n, flag = onlevp.householder(L, 300 * 2 * math.pi, maxiter=20, tol=1e-10)
    w0, v0, va = _normalised(L, sol)
    sp_, trm, ttm = W.get_surface_points(mg)
    sub = np.r_[0:len(sp_):9]
    flame_pts = [k for k, p in enumerate(sp_) if p in set(np.asarray(mg.tetrahedra)[mg.domains["Flame"]["simplices"]].ravel().tolist())][:12]
    sub = np.unique(np.r_[sub, flame_pts])
    pick = lambda lst: [lst[k] for k in sub]
    got = host_replay(mg, dscrp, cpt, sp_[sub], pick(trm), pick(ttm), w0, v0, va)
    so, tro, tto = oshape.get_surface_points(mo)
    want = oshape.discrete_adjoint_shape_sensitivity(mo, dscrp, cpt, [so[k] for k in sub], [tro[k] for k in sub], [tto[k] for k in sub], L, sol)
    assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max()


def test_adjoint_sensitivity_is_the_eigenvalue_derivative(rijke):
    """Pins the oracle (the reference stores no output of this path): for points that touch neither the flame nor the reference
    tetrahedron the discrete-adjoint value equals the derivative of the eigenvalue w.r.t. the point position, here by central
    differences of re-converged householder solutions (step 1e-6 of a 1e-3..1e-2-sized mesh)."""
    mg, mo, c, dscrp, L, sol, ref_idx = rijke
    so, tro, tto = oshape.get_surface_points(mo)
    flame = set(np.asarray(mo.tetrahedra)[mo.domains["Flame"]["simplices"]].ravel().tolist()) | set(mo.tetrahedra[ref_idx])
    picks = [k for k in range(5, len(so), 97) if so[k] not in flame][:6]
    sens = oshape.discrete_adjoint_shape_sensitivity(mo, dscrp, c, [so[k] for k in picks], [tro[k] for k in picks], [tto[k] for k in picks], L, sol)
    w0 = sol.params[sol.eigval]
    step = 1e-6
    for k in picks:
        p = so[k]
        for crd in range(3):
            ws = []
            for sgn in (1, -1):
                m2 = omesh.Mesh("m", scale=0.001, raw=load_raw_mesh("rijke_mm"))
                m2.points[crd, p] += sgn * step
                L2 = ohelm.discretize(m2, dscrp, c)
                s2, _, fl = onlevp.householder(L2, w0, maxiter=10, tol=1e-12)
                ws.append(s2.params[s2.eigval])
            fd = (ws[0] - ws[1]) / (2 * step)
            assert abs(fd - sens[crd, p]) <= 2e-4 * max(abs(sens[:, p]).max(), 1.0), (p, crd, fd, sens[crd, p])


def test_post_processing_matches_oracle(rijke):
    mg, mo = rijke[0], rijke[1]
    sp_, trm, ttm = W.get_surface_points(mg)
    so, tro, tto = oshape.get_surface_points(mo)
    nv = W.get_normal_vectors(mg)
    rng = np.random.default_rng(3)
    sens = np.zeros((3, mg.points.shape[1]), dtype=complex)
    sens[:, sp_] = rng.standard_normal((3, len(sp_))) + 1j * rng.standard_normal((3, len(sp_)))
    a, b = W.normalize_sensitivity(sp_, nv, trm, sens), oshape.normalize_sensitivity(so, nv, tro, sens)
    assert np.abs(a - b).max() <= 1e-12 * np.abs(b).max()
    assert np.abs(W.normal_sensitivity(nv, a) - oshape.normal_sensitivity(nv, b)).max() <= 1e-12 * np.abs(b).max()
    a2, b2 = W.bound_mass_normalize(sp_, nv, trm, mg, sens), oshape.bound_mass_normalize(so, nv, tro, mo, sens)
    assert np.abs(a2 - b2).max() <= 1e-10 * np.abs(b2).max()


def test_unit_cell_meshes_are_rejected(rijke):
    mg = rijke[0]
    cell = W.kuhn_unit_cell((2, 2, 2), (0, 0, 0), (1, 1, 1), DOS=4)
    with pytest.raises(NotImplementedError):
        W.get_surface_points(cell)
