"""Host mirror of the reference's shape-sensitivity interface over the batched CUDA kernel (csrc/shape_sens.cu).

  get_surface_points(mesh)                                   src/Meshutils.jl:884-966
  get_normal_vectors(mesh)                                   src/Meshutils.jl:1030-1069
  discrete_adjoint_shape_sensitivity(mesh, dscrp, C, surface_points, tri_mask, tet_mask, L, sol; h)
                                                             src/shape_sensitivity.jl:16-141
  normalize_sensitivity, bound_mass_normalize, normal_sensitivity   src/shape_sensitivity.jl:149-246

The reference re-runs ``discretize`` six times per surface point on a mesh whose domains are cut down to the simplices
touching the point (``mass_weighting=false`` and no ``order``: first-order elements).  Here the host only turns the model
descriptor into a list of terms (operator kind, scalar at omega_0, simplex lists per point) and one kernel launch per term
evaluates  -v_adj^H (E(x+h) - E(x-h)) v / (2h)  for all points and coordinates at once; no operator is assembled.
Unit-cell (Bloch) meshes (:84-118): points move along their local cylindrical basis, a point of the Bloch reference plane together
with its image, the operators are the blochified ones at b = 1 -- the kernel folds the DOFs and applies the class phases itself
(flames on Bloch meshes are not on the accelerated path, as in ``discretize``).  Indices are 0-based.
"""
import numpy as np

from . import _lib
from .helmholtz import _split_c
from .nlevp import (Sigma_nexp_az2mzit, Term, exp_az2mzit, exp_delay, generate_stsp_z, generate_z_g_z, get_context, pow1, pow2)


# --------------------------------------------------------------------------------------------- surface bookkeeping
def _group(keys, vals, n):
    """vals grouped by key (stable): list of n int64 arrays."""
    order = np.argsort(keys, kind="stable")
    cuts = np.searchsorted(keys[order], np.arange(n + 1))
    v = vals[order]
    return [v[cuts[i]:cuts[i + 1]] for i in range(n)]


def _is_unit(mesh):
    return getattr(mesh, "dos", 1) != 1 and bool(getattr(mesh.dos, "unit", False))


def get_surface_points(mesh, output=False):
    """surface_points (sorted point indices of all triangles), tri_mask[k] / tet_mask[k] = indices of the triangles / tetrahedra
    that contain surface point k, in ascending order (Meshutils.jl:884-944)."""
    tri, tet = mesh.triangles, mesh.tetrahedra
    surface_points = np.unique(tri)
    npts = mesh.points.shape[1]
    slot = np.full(npts, -1, dtype=np.int64)
    slot[surface_points] = np.arange(len(surface_points))
    n = len(surface_points)
    tri_mask = _group(slot[tri].ravel(), np.repeat(np.arange(len(tri)), 3), n)
    ts = slot[tet].ravel()
    keep = ts >= 0
    tet_mask = _group(ts[keep], np.repeat(np.arange(len(tet)), 4)[keep], n)
    if _is_unit(mesh):
        # Meshutils.jl:946-964: a point of the Bloch reference plane and its image share their simplex lists.  The reference indexes
        # the lists by point number, i.e. it relies on the first naxis + nxbloch points (and the last nxbloch ones) all being surface
        # points -- true for extend_mesh(unit=true), which keeps the triangles of the two periodic planes
        d = mesh.dos
        n0 = d.naxis + d.nxbloch
        if not (np.array_equal(surface_points[:n0], np.arange(n0)) and np.array_equal(surface_points[n - d.nxbloch:], np.arange(npts - d.nxbloch, npts))):
            raise ValueError("unit-cell mesh whose Bloch planes carry no triangles: the reference's get_surface_points is undefined here")

        def uniq(a, b):
            x = np.concatenate([a, b])
            _, first = np.unique(x, return_index=True)
            return x[np.sort(first)]
        for k in range(d.naxis, n0):
            img = n - d.nxbloch + (k - d.naxis)
            for mask in (tri_mask, tet_mask):
                mask[k] = uniq(mask[k], mask[img])
                mask[img] = uniq(mask[img], mask[k])
    return surface_points, tri_mask, tet_mask


def get_normal_vectors(mesh, output=False):
    """3 x n_tri outward normals of the surface triangles, length = twice the area (Meshutils.jl:1030-1069)."""
    if mesh.tri2tet is None:
        mesh.link_triangles_to_tetrahedra()
    tri = mesh.triangles
    tet = mesh.tetrahedra[mesh.tri2tet]
    inside = (tet[:, :, None] == tri[:, None, :]).any(axis=2)  # vertex of the adjacent tetrahedron that belongs to the triangle
    D = tet[np.arange(len(tet)), np.argmin(inside, axis=1)]    # first vertex that does not
    P = mesh.points
    A, B, Cc = P[:, tri[:, 0]], P[:, tri[:, 1]], P[:, tri[:, 2]]
    N = np.cross(A - Cc, B - Cc, axis=0)
    return N * np.sign(np.einsum("ij,ij->j", N, Cc - P[:, D]))


# --------------------------------------------------------------------------------------------- descriptor -> terms
def sensitivity_terms(mesh, dscrp, C, w0, bloch=False):
    """What ``discretize(mesh_h, dscrp, C, mass_weighting=false)(w0)`` is made of (Helmholtz.jl:232-403): a list of dicts
    {kind, dim, simplices, coef[, c][, ref_tet, n_ref, nl]}; ``coef`` is the term's scalar at w0 with the parameter values the
    descriptor itself provides (the reference evaluates the freshly discretised families, not ``L``)."""
    npts = mesh.points.shape[1]
    C_tet, C_tri = _split_c(mesh, C, npts)
    params = {"ω": complex(w0), "λ": complex("inf")}
    terms = []

    def scalar(func, arg):
        t = Term(None, func, arg, "", "")
        return t.scalar({v: (params[v], 0) for v in t.varlist})

    for domain, (typ, data) in dscrp.items():
        dim = mesh.domains[domain]["dimension"]
        simplices = np.asarray(mesh.domains[domain]["simplices"], dtype=np.int64)
        if typ in ("interior", "mass"):
            terms.append({"kind": _lib.SENS_MASS, "dim": 3, "simplices": simplices, "coef": scalar((pow2,), (("ω",),))})
            if typ == "interior":
                terms.append({"kind": _lib.SENS_STIFF, "dim": 3, "simplices": simplices, "coef": 1.0 + 0j, "c": C_tet})
        elif typ == "stiff":
            funcs, args, _ = data
            for a in args:
                for p in a:
                    if p != "ω":
                        params[p] = 0.0
            terms.append({"kind": _lib.SENS_STIFF, "dim": 3, "simplices": simplices, "coef": scalar(tuple(funcs), tuple(args)), "c": C_tet})
        elif typ == "admittance":
            if len(data) == 2:
                adm_sym, adm_val = data
                params.setdefault(adm_sym, complex(adm_val))
                coef = scalar((pow1, pow1), (("ω",), (adm_sym,)))
            elif len(data) == 1:
                coef = scalar((generate_z_g_z(data[0]),), (("ω",),))
            elif len(data) == 4:
                coef = scalar((generate_z_g_z(generate_stsp_z(*data)),), (("ω",),))
            else:
                raise ValueError("Data length does not match :admittance option!")
            terms.append({"kind": _lib.SENS_BOUNDARY, "dim": 2, "simplices": simplices, "coef": coef, "c": C_tri})
        elif typ in ("flame", "flameresponse", "fancyflame"):
            if bloch:
                raise NotImplementedError(f"descriptor type {typ!r} with Bloch periodicity is not on the accelerated path")
            ref_idx = -1
            if typ == "flame" and len(data) == 9:
                gamma, rho, nglobal, x_ref, n_ref, n_sym, tau_sym, n_val, tau_val = data
            elif typ == "flame" and len(data) == 10:
                gamma, rho, nglobal, ref_idx, x_ref, n_ref, n_sym, tau_sym, n_val, tau_val = data
            if typ == "flame" and len(data) in (9, 10):
                params.setdefault(n_sym, complex(n_val))
                params.setdefault(tau_sym, complex(tau_val))
                coef = scalar((pow1, exp_delay), ((n_sym,), ("ω", tau_sym)))
            elif typ == "flame" and len(data) == 6:
                gamma, rho, nglobal, x_ref, n_ref, FTF = data
                coef = scalar((FTF,), (("ω",),))
            elif typ == "flame" and len(data) == 5:
                gamma, rho, nglobal, x_ref, n_ref = data
                params["FTF"] = 0.0
                coef = scalar((pow1,), (("FTF",),))
            elif typ == "flameresponse":
                gamma, rho, nglobal, x_ref, n_ref, eps_sym, eps_val = data
                params.setdefault(eps_sym, complex(eps_val))
                coef = scalar((pow1,), ((eps_sym,),))
            elif typ == "fancyflame":
                gamma, rho, nglobal, x_ref, n_ref, n_sym, tau_sym, a_sym, n_val, tau_val, a_val = data
                if isinstance(n_val, (int, float, complex)):
                    for s_, v_ in ((n_sym, n_val), (tau_sym, tau_val), (a_sym, a_val)):
                        params.setdefault(s_, complex(v_))
                    coef = scalar((pow1, exp_az2mzit), ((n_sym,), ("ω", tau_sym, a_sym)))
                else:
                    arg = ["ω"]
                    for ns, ts, as_, nv, tv, av in zip(n_sym, tau_sym, a_sym, n_val, tau_val, a_val):
                        params[ns], params[ts], params[as_] = complex(nv), complex(tv), complex(av)
                        arg += [ns, ts, as_]
                    coef = scalar((Sigma_nexp_az2mzit,), (tuple(arg),))
            elif typ == "flame":
                raise ValueError("Data length does not match :flame option!")
            if ref_idx < 0:
                ref_idx = mesh.find_tetrahedron_containing_point(x_ref)
                if ref_idx < 0:
                    raise ValueError("reference point x_ref is not inside the mesh")
            terms.append({"kind": _lib.SENS_FLAME, "dim": 3, "simplices": simplices, "coef": coef, "ref_tet": int(ref_idx),
                          "n_ref": np.asarray(n_ref, dtype=float), "nl": (gamma - 1) / rho * nglobal})
        else:
            raise NotImplementedError(f"descriptor type {typ!r} is not on the accelerated path")
    return terms


def sensitivity_lists(term, n_elem, tri_mask, tet_mask):
    """Per moved point the simplices of the term's domain that touch it, in the order of the domain list
    (shape_sensitivity.jl:50-69) -> CSR (ptr, elems) and the speed of sound per list entry (or None)."""
    mask = tri_mask if term["dim"] == 2 else tet_mask
    n = len(mask)
    lens = np.array([len(m) for m in mask], dtype=np.int64)
    flat = np.concatenate([np.asarray(m, dtype=np.int64) for m in mask]) if lens.sum() else np.zeros(0, dtype=np.int64)
    slot = np.repeat(np.arange(n), lens)
    pos = np.full(n_elem, -1, dtype=np.int64)
    pos[term["simplices"]] = np.arange(len(term["simplices"]))
    keep = pos[flat] >= 0
    flat, slot = flat[keep], slot[keep]
    order = np.lexsort([pos[flat], slot])
    flat, slot = flat[order], slot[order]
    if len(flat):  # a simplex listed twice for the same point counts once (the reference tests membership)
        first = np.ones(len(flat), dtype=bool)
        first[1:] = (flat[1:] != flat[:-1]) | (slot[1:] != slot[:-1])
        flat, slot = flat[first], slot[first]
    ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(slot, minlength=n), out=ptr[1:])
    c = term["c"][flat] if "c" in term else None
    return ptr, flat, c


def discrete_adjoint_shape_sensitivity(mesh, dscrp, C, surface_points, tri_mask, tet_mask, L, sol, h=1e-9, output=False, ctx=None):
    """sens (3 x N_points complex): d omega / d x_p for every listed surface point, zero elsewhere; on unit-cell meshes the three
    rows are the radial, azimuthal and axial directions of the point (and axis points are skipped, :100-105)."""
    npts = mesh.points.shape[1]
    unit = _is_unit(mesh)
    fold = {}
    if unit:
        import math

        from .meshutils import bloch_dof_maps
        d = mesh.dos
        new, image, axis, red = bloch_dof_maps(mesh, "lin")
        if L.size() != red:
            raise ValueError("shape sensitivity is a first-order path: L must be discretize(mesh, dscrp, C, b='b') with order='lin'")
        fold = {"dof_new": new, "dof_flag": image.astype(np.uint8) | (axis.astype(np.uint8) << 1), "phase": np.exp(2j * math.pi / d.DOS),
                "cylindrical": True}
    elif L.size() != npts:
        raise ValueError("shape sensitivity is a first-order path: L must be discretize(mesh, dscrp, C) with order='lin'")
    w0 = sol.params[sol.eigval]
    v0 = np.asarray(sol.v, dtype=np.complex128)
    v0 = v0 / np.sqrt(np.vdot(v0, v0))
    va = np.asarray(sol.v_adj, dtype=np.complex128)
    va = va / np.conj(np.vdot(va, L(w0, 1) @ v0))  # :24-25 (one combine + one SpMV on the device)
    ctx = ctx or get_context()
    # the context holds one mesh at a time: make this one (first-order connectivity) resident; same topology as the family's own
    # mesh_set, so its patterns stay valid
    ctx.mesh_set(1, mesh.points.T, mesh.tetrahedra, mesh.triangles, npts)
    surface_points = np.asarray(surface_points, dtype=np.int64)
    keep = np.arange(len(surface_points))
    if unit:
        keep = keep[surface_points >= d.naxis]  # axis points are skipped (:100-105)
        pts = surface_points[keep]
        on_plane = (pts >= d.naxis) & (pts < d.naxis + d.nxbloch)
        fold["partner"] = np.where(on_plane, npts - d.nxbloch + (pts - d.naxis), -1)  # :88-91
        tri_mask, tet_mask = [tri_mask[k] for k in keep], [tet_mask[k] for k in keep]
    pts = surface_points[keep]
    ctx.shape_sens_begin(pts, h, v0, va, **fold)
    for term in sensitivity_terms(mesh, dscrp, C, w0, bloch=unit):
        n_elem = len(mesh.triangles) if term["dim"] == 2 else len(mesh.tetrahedra)
        ptr, elems, c = sensitivity_lists(term, n_elem, tri_mask, tet_mask)
        ctx.shape_sens_add(term["kind"], ptr, elems, term["coef"], c=c, ref_tet=term.get("ref_tet", 0), n_ref=term.get("n_ref"),
                           nl=term.get("nl", 0.0))
    sens = np.zeros((3, npts), dtype=np.complex128)
    sens[:, pts] = ctx.shape_sens_end()
    return sens


# --------------------------------------------------------------------------------------------- post-processing (host, surface-sized)
def normalize_sensitivity(surface_points, normal_vectors, tri_mask, sens):
    """Point sensitivities spread over the adjacent triangles, weighted by their projected areas (shape_sensitivity.jl:149-184)."""
    ntri = normal_vectors.shape[1]
    A = np.linalg.norm(normal_vectors, axis=0) / 2
    lens = np.array([len(m) for m in tri_mask], dtype=np.int64)
    tri = np.concatenate([np.asarray(m, dtype=np.int64) for m in tri_mask]) if lens.sum() else np.zeros(0, dtype=np.int64)
    slot = np.repeat(np.arange(len(tri_mask)), lens)
    pnt = np.asarray(surface_points, dtype=np.int64)[slot]
    out = np.zeros((3, ntri), dtype=np.complex128)
    for crd in range(3):
        V = np.abs(normal_vectors[crd]) / 6
        vol = np.bincount(slot, weights=V[tri], minlength=len(tri_mask))
        ok = (vol[slot] != 0) & (A[tri] > 0)
        w = np.zeros(len(tri))
        w[ok] = V[tri[ok]] / vol[slot[ok]] / A[tri[ok]]
        val = sens[crd, pnt] * w
        out[crd] = np.bincount(tri, weights=val.real, minlength=ntri) + 1j * np.bincount(tri, weights=val.imag, minlength=ntri)
    return out


def bound_mass_normalize(surface_points, normal_vectors, tri_mask, mesh, sens):
    """sens on the surface points multiplied by the inverse of the first-order boundary mass matrix (shape_sensitivity.jl:191-229;
    the reference factorises this surface-sized matrix with SparseArrays.lu on the host as well)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    M = np.array([[1 / 12, 1 / 24, 1 / 24], [1 / 24, 1 / 12, 1 / 24], [1 / 24, 1 / 24, 1 / 12]])
    surface_points = np.asarray(surface_points, dtype=np.int64)
    slot = np.full(mesh.points.shape[1], -1, dtype=np.int64)
    slot[surface_points] = np.arange(len(surface_points))
    t = slot[mesh.triangles]
    nrm = np.linalg.norm(normal_vectors, axis=0)
    I = np.repeat(t, 3, axis=1).ravel()
    J = np.tile(t, (1, 3)).ravel()
    V = (M.ravel()[None, :] * nrm[:, None]).ravel()
    n = len(surface_points)
    B = spla.splu(sp.csc_matrix(sp.coo_matrix((V, (I, J)), shape=(n, n))))
    out = np.zeros_like(sens)
    for i in range(3):
        out[i, surface_points] = B.solve(sens[i, surface_points].real.copy()) + 1j * B.solve(sens[i, surface_points].imag.copy())
    return out


def normal_sensitivity(normal_vectors, normed_sens):
    """Component of the triangle sensitivities along the unit normals (shape_sensitivity.jl:237-246)."""
    n = normal_vectors / np.linalg.norm(normal_vectors, axis=0)
    return np.einsum("ij,ij->j", n, normed_sens)
