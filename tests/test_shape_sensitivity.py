"""Discrete-adjoint shape sensitivity (src/shape_sensitivity.jl:16-141), CPU part: the product's host logic (surface bookkeeping,
descriptor -> terms, per-point simplex lists, post-processing) and the per-thread function of the CUDA kernel replayed on the
host (wae_shape_sens_check) against the oracle's literal restatement (six discretize calls per point on reduced domains), and
the oracle itself against true finite differences of the eigenvalue.  The GPU launch of the same function is tested in
test_zy_shape_sensitivity_gpu.py."""
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
    _pi32, _pu8 = C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
    f.restype = C.c_int32
    f.argtypes = [C.c_int64, _pd, C.c_int64, _pu32, C.c_int64, _pu32, C.c_int64, _pi64, _pi64, C.c_double, C.c_int32, _pd, _pd, _pi32, _pu8, _pd,
                  C.c_int32, _pi64, _pi64, _pd, C.c_int32, _pd, C.c_int64, _pd, C.c_double, _pd]
    xyz = np.ascontiguousarray(mesh.points.T, dtype=np.float64)
    tets = np.ascontiguousarray(mesh.tetrahedra, dtype=np.uint32)
    tris = np.ascontiguousarray(mesh.triangles, dtype=np.uint32)
    npts = xyz.shape[0]
    pts = np.ascontiguousarray(surface_points, dtype=np.int64)
    unit = shape._is_unit(mesh)
    partner = dof_new = dof_flag = phase = None
    if unit:  # the same preparation as discrete_adjoint_shape_sensitivity
        from wae_b200.meshutils import bloch_dof_maps
        d = mesh.dos
        keep = np.flatnonzero(pts >= d.naxis)
        pts = np.ascontiguousarray(pts[keep])
        tri_mask, tet_mask = [tri_mask[k] for k in keep], [tet_mask[k] for k in keep]
        partner = np.ascontiguousarray(np.where(pts < d.naxis + d.nxbloch, npts - d.nxbloch + (pts - d.naxis), -1), dtype=np.int64)
        new, image, axis, red = bloch_dof_maps(mesh, "lin")
        dof_new = np.ascontiguousarray(new, dtype=np.int32)
        dof_flag = np.ascontiguousarray(image.astype(np.uint8) | (axis.astype(np.uint8) << 1))
        phase = np.array([np.exp(2j * math.pi / d.DOS)])
        assert len(v0) == red
    v0 = np.ascontiguousarray(v0, dtype=np.complex128)
    va = np.ascontiguousarray(va, dtype=np.complex128)
    out = np.zeros((len(pts), 3), dtype=np.complex128)
    P = lambda a, t: None if a is None else a.ctypes.data_as(t)
    for term in shape.sensitivity_terms(mesh, dscrp, c, w0, bloch=unit):
        n_elem = len(tris) if term["dim"] == 2 else len(tets)
        ptr, elems, cc = shape.sensitivity_lists(term, n_elem, tri_mask, tet_mask)
        cpe = 1
        if cc is not None:
            cc = np.ascontiguousarray(cc, dtype=np.float64)
            cpe = 1 if cc.ndim == 1 else cc.shape[1]
        coef = np.array([complex(term["coef"])])
        nref = None if "n_ref" not in term else np.ascontiguousarray(term["n_ref"], dtype=np.float64)
        rc = f(npts, P(xyz, _pd), len(tets), P(tets, _pu32), len(tris), P(tris, _pu32), len(pts), P(pts, _pi64), P(partner, _pi64), h, int(unit),
               P(v0, _pd), P(va, _pd), P(dof_new, _pi32), P(dof_flag, _pu8), P(phase, _pd),
               term["kind"], P(ptr, _pi64), P(elems, _pi64), P(cc, _pd), cpe, P(coef, _pd), term.get("ref_tet", 0), P(nref, _pd),
               float(term.get("nl", 0.0)), P(out, _pd))
        assert rc == 0, (rc, term["kind"])
    sens = np.zeros((3, npts), dtype=np.complex128)
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
    sub = np.r_[0:len(sp_):20]
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
    picks = [k for k in range(5, len(so), 190) if so[k] not in flame][:3]
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


def test_unit_cell_variant_on_the_ntnu_combustor():
    """shape_sensitivity.jl:84-118 on the reference's own annular-combustor mesh (docs/src/NTNU_12.msh -> extend_mesh(unit=true), 19 axis
    points, 396 Bloch-plane points): cylindrical moves, Bloch-plane points moving with their images, blochified operators at b = 1.
    The oracle re-discretises the unit cell six times per point (incl. the weighting matrix over all 8446 tetrahedra), so only a
    handful of points of every kind is compared; the surface bookkeeping is compared for all 846 points."""
    from oracle.mesh import extend_mesh as oext
    from test_bloch import NTNU_DOMS, NTNU_DSCRP, _ntnu_meshes, _ntnu_sos
    mg, mo = _ntnu_meshes()
    doms = NTNU_DOMS + [("CC", "half")]
    g, o = W.extend_mesh(mg, doms, unit=True), oext(mo, doms, unit=True)
    sg, trg, ttg = W.get_surface_points(g)
    so, tro, tto = oshape.get_surface_points(o)
    assert list(sg) == so and all(list(a) == b for a, b in zip(trg, tro)) and all(list(a) == b for a, b in zip(ttg, tto))
    c = o.generate_field(_ntnu_sos)
    dscrp = dict(NTNU_DSCRP)
    dscrp["Outlet_high"] = ("admittance", ("Y_in", 0.2 + 0.1j))  # a non-zero admittance so that the boundary term takes part
    L = ohelm.discretize(o, dscrp, c, b="b")
    L.params["b"] = 1 + 0j
    sol, n, flag = onlevp.mslp(L, 1000.0, maxiter=20, tol=1e-10, scale=2 * math.pi)
    assert flag == 0 and 500 < sol.params["ω"].real / 2 / math.pi < 1500  # one of the two b = 1 modes near 863 / 1124 Hz
    w0, v0, va = _normalised(L, sol)
    d = g.dos
    npts = g.points.shape[1]
    outlet = set(np.asarray(g.triangles)[g.domains["Outlet_high"]["simplices"]].ravel().tolist())
    kinds = {"axis": [k for k, p in enumerate(sg) if p < d.naxis][:1],
             "plane": [k for k, p in enumerate(sg) if d.naxis <= p < d.naxis + d.nxbloch][5:7],
             "image": [k for k, p in enumerate(sg) if p >= npts - d.nxbloch][5:7],
             "outlet": [k for k, p in enumerate(sg) if p in outlet and d.naxis + d.nxbloch <= p < npts - d.nxbloch][:2],
             "body": [k for k, p in enumerate(sg) if d.naxis + d.nxbloch <= p < npts - d.nxbloch and p not in outlet][10:12]}
    sub = sorted(k for ks in kinds.values() for k in ks)
    assert all(len(v) > 0 for v in kinds.values())
    pick = lambda lst: [lst[k] for k in sub]
    got = host_replay(g, dscrp, c, sg[sub], pick(trg), pick(ttg), w0, v0, va)
    want = oshape.discrete_adjoint_shape_sensitivity(o, dscrp, c, [so[k] for k in sub], pick(tro), pick(tto), L, sol)
    scale = np.abs(want).max()
    assert scale > 0 and np.abs(want[:, sg[kinds["axis"]]]).max() == 0  # axis points are skipped
    assert np.abs(got - want).max() <= 1e-5 * scale, np.abs(got - want).max() / scale
    if os.environ.get("WAE_TEST_VERBOSE"):
        for name, ks in kinds.items():
            print(name, [(float(np.abs(want[:, sg[k]]).max()), float(np.abs(got - want)[:, sg[k]].max())) for k in ks])
    for k in kinds["plane"] + kinds["image"] + kinds["outlet"] + kinds["body"]:
        assert np.abs(want[:, sg[k]]).max() > 0
