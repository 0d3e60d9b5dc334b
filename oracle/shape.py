"""Oracle (test infrastructure): discrete-adjoint shape sensitivity, restated literally.

  get_surface_points                    src/Meshutils.jl:884-966
  get_normal_vectors                    src/Meshutils.jl:1030-1069
  discrete_adjoint_shape_sensitivity    src/shape_sensitivity.jl:16-141, get_cylindrics :361-370
  normalize_sensitivity                 src/shape_sensitivity.jl:149-184
  bound_mass_normalize                  src/shape_sensitivity.jl:191-229
  normal_sensitivity                    src/shape_sensitivity.jl:237-246

Every surface point costs six calls of the oracle's ``discretize`` on a copy of the mesh whose domains are cut down to the
simplices touching the point -- exactly the reference's loop; this is the thing the product replaces by one batched kernel.
All indices 0-based.  Parity unpinned: the reference stores no output of this path (examples/shape/*.jl write VTK files only).
"""
import copy

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .helmholtz import discretize
from .mesh import Mesh


def get_surface_points(mesh):
    surface_points = sorted({int(p) for tri in mesh.triangles for p in tri})
    pos = {p: i for i, p in enumerate(surface_points)}
    tri_mask = [[] for _ in surface_points]
    tet_mask = [[] for _ in surface_points]
    for it, tri in enumerate(mesh.triangles):
        for p in tri:
            tri_mask[pos[int(p)]].append(it)
    for it, tet in enumerate(mesh.tetrahedra):
        for p in tet:
            if int(p) in pos:
                tet_mask[pos[int(p)]].append(it)
    dos = getattr(mesh, "dos", 1)
    if dos != 1 and dos.unit:  # Meshutils.jl:946-964 (the lists are indexed by POINT number there; 0-based here)
        def unique_inplace(lst):
            seen, out = set(), []
            for x in lst:
                if x not in seen:
                    seen.add(x)
                    out.append(x)
            lst[:] = out
        n = len(surface_points)
        for idx in surface_points:
            bidx = idx - dos.naxis  # 0-based: 0 <= bidx < nxbloch
            if 0 <= bidx < dos.nxbloch:
                img = n - dos.nxbloch + bidx
                for mask in (tri_mask, tet_mask):
                    mask[idx].extend(mask[img])
                    unique_inplace(mask[idx])
                    mask[img].extend(mask[idx])
                    unique_inplace(mask[img])
    return surface_points, tri_mask, tet_mask


def get_normal_vectors(mesh):
    if mesh.tri2tet is None:
        mesh.link_triangles_to_tetrahedra()
    nv = np.empty((3, len(mesh.triangles)))
    for idx, tri in enumerate(mesh.triangles):
        tet = mesh.tetrahedra[mesh.tri2tet[idx]]
        D = next(p for p in tet if p not in tri)
        A, B, C = (mesh.points[:, p] for p in tri)
        N = np.cross(A - C, B - C)
        nv[:, idx] = N * np.sign(np.dot(N, C - mesh.points[:, D]))
    return nv


def _reduced_mesh(mesh, dscrp, tris, tets):
    m = Mesh.__new__(Mesh)
    m.name = "mesh_h"
    m.points = mesh.points.copy()
    m.lines, m.triangles, m.tetrahedra, m.tri2tet = mesh.lines, mesh.triangles, mesh.tetrahedra, mesh.tri2tet
    m.dos = getattr(mesh, "dos", 1)
    if hasattr(mesh, "_lmap"):
        m._lmap = mesh._lmap
    m.domains = {}
    for dom in dscrp:
        d = copy.deepcopy(mesh.domains[dom])
        keep = set(tris) if d["dimension"] == 2 else set(tets) if d["dimension"] == 3 else set()
        d["simplices"] = [s for s in d["simplices"] if s in keep]
        m.domains[dom] = d
    return m


def get_cylindrics(pnt):
    """shape_sensitivity.jl:361-370: columns = local radial, azimuthal, axial unit vectors."""
    X = np.zeros((3, 3))
    X[:, 2] = [0, 0, 1]
    X[:, 0] = pnt
    X[2, 0] = 0.0
    X[:, 0] /= np.linalg.norm(X[:, 0])
    X[:, 1] = np.cross(X[:, 2], X[:, 0])
    return X


def discrete_adjoint_shape_sensitivity(mesh, dscrp, C, surface_points, tri_mask, tet_mask, L, sol, h=1e-9):
    w0 = sol.params[sol.eigval]
    v0 = np.asarray(sol.v, dtype=complex)
    v0 = v0 / np.sqrt(np.vdot(v0, v0))
    va = np.asarray(sol.v_adj, dtype=complex)
    va = va / np.conj(np.vdot(va, L(w0, 1) @ v0))
    if mesh.tri2tet is None:
        mesh.link_triangles_to_tetrahedra()
    dos = getattr(mesh, "dos", 1)
    unit = dos != 1 and dos.unit
    b = "b" if unit else None
    npts = mesh.points.shape[1]
    sens = np.zeros((3, npts), dtype=complex)
    for idx, pnt_idx in enumerate(surface_points):
        mh = _reduced_mesh(mesh, dscrp, tri_mask[idx], tet_mask[idx])
        pnt = mh.points[:, pnt_idx].copy()
        bloch = False
        if unit:
            if 0 <= pnt_idx - dos.naxis < dos.nxbloch:  # :86-91 identify the Bloch image point
                pnt_bloch_idx = npts - dos.nxbloch + (pnt_idx - dos.naxis)
                pnt_bloch = mesh.points[:, pnt_bloch_idx].copy()
                bloch = True
            elif pnt_idx < dos.naxis:  # :100-105 axis points are skipped
                continue
        for crd in range(3):
            mh.points[:, pnt_idx] = pnt
            if unit:
                X = get_cylindrics(pnt)
                mh.points[:, pnt_idx] += h * X[:, crd]
                if bloch:
                    mh.points[:, pnt_bloch_idx] = pnt_bloch
                    Xb = get_cylindrics(pnt_bloch)
                    mh.points[:, pnt_bloch_idx] += h * Xb[:, crd]
            else:
                mh.points[crd, pnt_idx] += h
            D_right = discretize(mh, dscrp, C, mass_weighting=False, b=b)
            if unit:
                mh.points[:, pnt_idx] -= 2 * h * X[:, crd]
                if bloch:
                    mh.points[:, pnt_bloch_idx] -= 2 * h * Xb[:, crd]
            else:
                mh.points[crd, pnt_idx] -= 2 * h
            D_left = discretize(mh, dscrp, C, mass_weighting=False, b=b)
            if unit:
                D_right.params[b] = D_left.params[b] = 1 + 0j
            D = (D_right(w0) - D_left(w0)) / (2 * h)
            sens[crd, pnt_idx] = -np.vdot(va, D @ v0)
    return sens


def normalize_sensitivity(surface_points, normal_vectors, tri_mask, sens):
    ntri = normal_vectors.shape[1]
    out = np.zeros((3, ntri), dtype=complex)
    A = np.linalg.norm(normal_vectors, axis=0) / 2
    for crd in range(3):
        V = np.abs(normal_vectors[crd]) / 6
        for idx, pnt in enumerate(surface_points):
            tris = tri_mask[idx]
            vol = sum(abs(V[t]) for t in tris)
            if vol == 0:
                continue
            for t in tris:
                if A[t] > 0:
                    out[crd, t] += sens[crd, pnt] / A[t] * (abs(V[t]) / vol)
    return out


def bound_mass_normalize(surface_points, normal_vectors, tri_mask, mesh, sens):
    M = np.array([[1 / 12, 1 / 24, 1 / 24], [1 / 24, 1 / 12, 1 / 24], [1 / 24, 1 / 24, 1 / 12]])
    pos = {p: i for i, p in enumerate(surface_points)}
    I, J, V = [], [], []
    for idx, tri in enumerate(mesh.triangles):
        mm = M * np.linalg.norm(normal_vectors[:, idx])
        for a in range(3):
            for b in range(3):
                I.append(pos[int(tri[b])]); J.append(pos[int(tri[a])]); V.append(mm[b, a])
    n = len(surface_points)
    B = spla.splu(sp.csc_matrix(sp.coo_matrix((V, (I, J)), shape=(n, n))))
    out = np.zeros_like(sens)
    sp_ = np.asarray(surface_points)
    for i in range(3):
        out[i, sp_] = B.solve(sens[i, sp_].real.copy()) + 1j * B.solve(sens[i, sp_].imag.copy())
    return out


def normal_sensitivity(normal_vectors, normed_sens):
    n = normal_vectors / np.linalg.norm(normal_vectors, axis=0)
    return np.einsum("ij,ij->j", n, normed_sens)  # LinearAlgebra.dot(v, s) with a real first argument
