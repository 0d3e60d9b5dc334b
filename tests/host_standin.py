"""TEST DOUBLE of ``wae_b200._lib.Context`` -- test infrastructure only, never imported by the product.

The product's host mirror of the reference interface (``discretize``, ``LinearOperatorFamily``, ``householder``/``mslp``, ``perturb_fast!``,
``discrete_adjoint_shape_sensitivity``, the forced response) is Python above the C ABI and cannot run without a GPU: ``wae_create`` fails
and there is no CPU fallback.  To exercise that HOST LOGIC in the ``-m "not gpu"`` suite -- term bookkeeping, parameter handling, flags,
the argument marshalling of the begin/add/end sequences -- this class stands in for the context object with the same method names and
conventions (0-based ids, CSC patterns with sorted rows and explicit zeros, trans 0/1/2 = N/T/C), doing the numeric work with the oracle's
element routines and scipy (SuperLU, ARPACK).  It says nothing about the CUDA kernels: those are checked by the ``-m gpu`` tests through
the real library.  ``shape_sens_*`` goes through the library's host-only replay of the kernel's per-thread function (wae_shape_sens_check)."""
import ctypes as C
import os

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import fem
from wae_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class HostStandIn:
    def __init__(self):
        self.pats, self.mats, self.fams, self.lus = [], [], [], []
        self._ms, self.launches = {}, 0
        self.device = -1

    # -- plumbing ------------------------------------------------------------------------------------------------------------
    def set_stream(self, ptr):
        pass

    def sync(self):
        pass

    def close(self):
        pass

    def launch_count(self):
        return self.launches

    def last_ms(self, phase):
        return self._ms.get(phase, -1.0)

    # -- mesh / patterns -------------------------------------------------------------------------------------------------------
    def mesh_set(self, order, xyz, tets, tris, dim):
        self.order, self.P = order, np.ascontiguousarray(np.asarray(xyz, dtype=float).T)
        self.tets, self.tris, self.dim = np.asarray(tets, dtype=np.int64), np.asarray(tris, dtype=np.int64), dim
        self.mesh_serial = getattr(self, "mesh_serial", 0) + 1
        return self.mesh_serial

    def mesh_update_points(self, xyz):
        self.P = np.ascontiguousarray(np.asarray(xyz, dtype=float).T)

    def _new_pattern(self, dim, I, J, kind=0, elems=None):
        A = sp.csc_matrix(sp.coo_matrix((np.ones(len(I)), (I, J)), shape=(dim, dim)))
        A.sort_indices()
        self.pats.append({"dim": dim, "colptr": A.indptr.astype(np.int64), "rowval": A.indices.astype(np.int64), "kind": kind, "elems": elems,
                          "keys": np.repeat(np.arange(dim, dtype=np.int64), np.diff(A.indptr)) * dim + A.indices})
        return len(self.pats) - 1, A.nnz

    def pattern_build(self, elem_kind, elem_ids=None):
        conn = self.tets if elem_kind == 3 else self.tris
        ids = np.arange(len(conn)) if elem_ids is None else np.asarray(elem_ids, dtype=np.int64)
        c = conn[ids]
        n = c.shape[1]
        return self._new_pattern(self.dim, np.repeat(c, n, axis=1).ravel(), np.tile(c, (1, n)).ravel(), elem_kind, ids)

    def pattern_get(self, pid, dim, nnz):
        return self.pats[pid]["colptr"], self.pats[pid]["rowval"]

    def _align(self, pid, I, J, V):
        """values of sparse(I, J, V) on the stored pattern (explicit zeros kept)"""
        p = self.pats[pid]
        pos = np.searchsorted(p["keys"], np.asarray(J, dtype=np.int64) * p["dim"] + np.asarray(I, dtype=np.int64))
        out = np.zeros(len(p["keys"]), dtype=complex)
        np.add.at(out, pos, V)
        return out

    def _store(self, pid, val, reuse=-1):
        if reuse >= 0:
            self.mats[reuse]["val"] = val
            return reuse
        self.mats.append({"pid": pid, "val": val})
        return len(self.mats) - 1

    # -- assembly (element routines of the oracle, Helmholtz.jl:405-503) -----------------------------------------------------------
    def _triplets(self, elem_kind, ids, kind, c, scale):
        conn = self.tets if elem_kind == 3 else self.tris
        I, J, V = [], [], []
        for k, e in enumerate(ids):
            s = conn[e]
            ct = fem.CooTrafo(self.P[:, s[:4 if elem_kind == 3 else 3]])
            if kind == _lib.OP_MASS:
                vv = fem.tet_mass(ct, self.order) * scale
            elif kind == _lib.OP_STIFF:
                vv = -c[k] ** 2 * fem.tet_stiff(ct, self.order) if np.ndim(c[k]) == 0 else -fem.tet_stiff_cc1(ct, c[k], self.order)
            else:
                vv = (c[k] * fem.tri_mass(ct, self.order) if np.ndim(c[k]) == 0 else fem.tri_mass_c1(ct, c[k], self.order)) * (-1j * scale)
            ii, jj = fem.create_indices(s)
            I.extend(ii.T.ravel()); J.extend(jj.T.ravel()); V.extend(np.asarray(vv).T.ravel())
        return np.asarray(I, dtype=np.int64), np.asarray(J, dtype=np.int64), np.asarray(V, dtype=complex)

    def _elements(self, pid, kind, c, scale):
        p = self.pats[pid]
        return self._align(pid, *self._triplets(p["kind"], p["elems"], kind, c, scale))

    def assemble(self, pid, kind, c=None, scale=1.0, reuse=-1):
        self._ms["assemble"] = 1e-3
        self.launches += 1
        return self._store(pid, self._elements(pid, kind, None if c is None else np.asarray(c, dtype=float), scale), reuse)

    def assemble_mk(self, pid, c, reuse=(-1, -1)):
        c = np.asarray(c, dtype=float)
        return self.assemble(pid, _lib.OP_MASS, None, 1.0, reuse[0]), self.assemble(pid, _lib.OP_STIFF, c, 1.0, reuse[1])

    def assemble_flame(self, flame_tets, ref_tet, x_ref, n_ref, nlocal, reuse=-1):
        rows = np.unique(self.tets[np.asarray(flame_tets, dtype=np.int64)])
        S = np.zeros(self.dim)
        for e in flame_tets:
            s = self.tets[e]
            np.add.at(S, s, fem.tet_src(fem.CooTrafo(self.P[:, s[:4]]), self.order))
        s = self.tets[ref_tet]
        G = np.zeros(self.dim)
        np.add.at(G, s, -nlocal * fem.tet_grad_at(fem.CooTrafo(self.P[:, s[:4]]), n_ref, x_ref, self.order))
        cols = np.unique(s)
        I, J = np.repeat(rows, len(cols)), np.tile(cols, len(rows))
        pid, nnz = (self.mats[reuse]["pid"], None) if reuse >= 0 else self._new_pattern(self.dim, I, J)
        mid = self._store(pid, self._align(pid, I, J, (S[I] * G[J]).astype(complex)), reuse)
        self._ms["assemble"] = 1e-3
        self.launches += 3
        return pid, mid, len(self.pats[pid]["keys"])

    def assemble_wallsrc(self, tri_ids, c, dim):
        out = np.zeros(dim, dtype=complex)
        c = np.asarray(c, dtype=float)
        for k, e in enumerate(np.asarray(tri_ids, dtype=np.int64)):
            s = self.tris[e]
            ct = fem.CooTrafo(self.P[:, s[:3]])
            out[s] += (c[k] * fem.tri_src(ct, self.order) if c.ndim == 1 else fem.tri_src_c1(ct, c[k], self.order)) / 1j
        self._ms["assemble"] = 1e-3
        self.launches += 1
        return out

    def assemble_bloch(self, elem_kind, elem_ids, kind, c, scale, dim_red, dof_new, dof_flag, n_class):
        """Bloch.jl:4-112 as the library does it: DOFs folded through dof_new, every element entry sorted into the classes plain / +
        (column DOF is an image) / - (row DOF is an image), each once more for entries touching an axis DOF when n_class == 6;
        n_class == 1 sums everything (the weighting matrix, Helmholtz.jl:541-549)."""
        conn = self.tets if elem_kind == 3 else self.tris
        ids = np.arange(len(conn)) if elem_ids is None else np.asarray(elem_ids, dtype=np.int64)
        I, J, V = self._triplets(elem_kind, ids, kind, None if c is None else np.asarray(c, dtype=float), scale)
        dof_new, dof_flag = np.asarray(dof_new, dtype=np.int64), np.asarray(dof_flag, dtype=np.uint8)
        fi, fj = dof_flag[I], dof_flag[J]
        cls = np.zeros(len(I), dtype=np.int64)
        if n_class > 1:
            ic, jc = (fi & 1).astype(bool), (fj & 1).astype(bool)
            cls = np.where(ic == jc, 0, np.where(jc, 1, 2))
            if n_class == 6:
                cls = cls + 3 * (((fi | fj) & 2) != 0)
        pids, mids = [], []
        for k in range(n_class):
            m = cls == k
            pid, _ = self._new_pattern(dim_red, dof_new[I[m]], dof_new[J[m]])
            pids.append(pid)
            mids.append(self._store(pid, self._align(pid, dof_new[I[m]], dof_new[J[m]], V[m])))
        self._ms["assemble"] = 1e-3
        self.launches += 1
        return pids, mids

    # -- matrices / families -----------------------------------------------------------------------------------------------------
    def mat_info(self, mid):
        m = self.mats[mid]
        return m["pid"], True, len(m["val"])

    def mat_get(self, mid):
        return self.mats[mid]["val"].copy()

    def mat_set(self, dim, colptr, rowval, nzval):
        J = np.repeat(np.arange(dim), np.diff(colptr))
        pid, _ = self._new_pattern(dim, np.asarray(rowval), J)
        return pid, self._store(pid, self._align(pid, rowval, J, np.asarray(nzval, dtype=complex)))

    def mat_free(self, mid):
        self.mats[mid] = None

    def _csc(self, pid, val):
        p = self.pats[pid]
        return sp.csc_matrix((val, p["rowval"], p["colptr"]), shape=(p["dim"], p["dim"]))

    def family_create(self, mat_ids):
        dim = self.pats[self.mats[mat_ids[0]]["pid"]]["dim"]
        I = np.concatenate([self.pats[self.mats[m]["pid"]]["rowval"] for m in mat_ids])
        J = np.concatenate([np.repeat(np.arange(dim), np.diff(self.pats[self.mats[m]["pid"]]["colptr"])) for m in mat_ids])
        pid, nnz = self._new_pattern(dim, I, J)
        self.fams.append({"pid": pid, "mats": list(map(int, mat_ids)), "slots": {}})
        return len(self.fams) - 1, nnz

    def family_pattern_get(self, fid, dim, nnz):
        p = self.pats[self.fams[fid]["pid"]]
        return p["colptr"], p["rowval"]

    def combine(self, fid, coeffs, slot):
        f = self.fams[fid]
        out = np.zeros(len(self.pats[f["pid"]]["keys"]), dtype=complex)
        for m, cf in zip(f["mats"], np.asarray(coeffs, dtype=complex)):
            if cf != 0:
                p = self.pats[self.mats[m]["pid"]]
                out[np.searchsorted(self.pats[f["pid"]]["keys"], p["keys"])] += cf * self.mats[m]["val"]
        f["slots"][slot] = out
        self.launches += 1

    def family_get(self, fid, slot, nnz):
        return self.fams[fid]["slots"][slot].copy()

    def _slot(self, fid, slot):
        return self._csc(self.fams[fid]["pid"], self.fams[fid]["slots"][slot])

    def spmm(self, fid, slot, X, trans=0):
        A = self._slot(fid, slot)
        self.launches += 1
        return (A, A.T, A.conj().T)[trans] @ np.asarray(X, dtype=complex)

    # -- LU / eigs ---------------------------------------------------------------------------------------------------------------
    def lu_analyze(self, fid):
        self.lus.append({"fid": fid, "A": None, "lu": None})
        return len(self.lus) - 1, 0, 0.0

    def lu_free(self, lid):
        if self.lus[lid] is None:
            raise _lib.WaeError(_lib.E_INVALID, f"unknown LU id {lid}")
        self.lus[lid] = None

    def family_free(self, fid):
        if self.fams[fid] is None:
            raise _lib.WaeError(_lib.E_INVALID, f"unknown family id {fid}")
        if any(S is not None and S["fid"] == fid for S in self.lus):
            raise _lib.WaeError(_lib.E_INVALID, f"family {fid} still has an LU handle (wae_lu_free first)")
        self.fams[fid] = None

    def lu_factor(self, lid, slot, check=True):
        S = self.lus[lid]
        S["A"] = self._slot(S["fid"], slot)
        try:
            S["lu"] = spla.splu(S["A"])
        except RuntimeError as e:  # exactly singular
            raise _lib.SingularException(_lib.E_SINGULAR, str(e))
        self._ms["factor"] = 0.0
        self.launches += 1

    def lu_solve(self, lid, B, trans=0):
        self._ms["solve"] = 0.0
        self.launches += 1
        return self.lus[lid]["lu"].solve(np.asarray(B, dtype=complex), trans=("N", "T", "H")[trans])

    def eigs_si(self, lid, fid, m_slot, nev, v0, trans=0):
        A, M = self.lus[lid]["A"], self._slot(fid, m_slot)
        if trans == 2:
            A, M = A.conj().T, M.conj().T
        lam, V = spla.eigs(sp.csc_matrix(A), k=nev, M=sp.csc_matrix(M), sigma=0, v0=np.asarray(v0, dtype=complex), tol=0)
        return lam, V, 20

    def moment_buffer(self, n_mom, l, d):
        import torch
        return torch.zeros((n_mom, l, d), dtype=torch.complex128)

    def beyn_moments(self, fid, lid, z, w, coeffs, l, n_mom, out_ptr, V=None):
        """beyn.jl:62-74 for the nodes handed in: A_p += w_j z_j^p L(z_j)^-1 V, V = first l identity columns unless given."""
        f = self.fams[fid]
        d = self.pats[f["pid"]]["dim"]
        A = np.ctypeslib.as_array((C.c_double * (2 * n_mom * l * d)).from_address(out_ptr)).view(np.complex128).reshape(n_mom, l, d)
        Vm = np.eye(d, l, dtype=complex) if V is None else np.asarray(V, dtype=complex)
        for zj, wj, cf in zip(np.asarray(z), np.asarray(w), np.asarray(coeffs)):
            self.combine(fid, cf, 0)
            X = spla.splu(self._slot(fid, 0)).solve(Vm)
            for p in range(n_mom):
                A[p] += (wj * zj**p) * X.T
        self._ms["beyn_factor_total"] = self._ms["beyn_solve_total"] = 0.0

    # -- shape sensitivity: the library's host replay of the kernel's per-thread function -----------------------------------------
    def shape_sens_begin(self, points, step, v, v_adj, partner=None, cylindrical=False, dof_new=None, dof_flag=None, phase=None):
        if self.order != 1:
            raise _lib.WaeError(_lib.E_INVALID, "shape sensitivity needs a first-order mesh")
        self._sens = {"pts": np.ascontiguousarray(points, dtype=np.int64), "h": float(step), "v": np.ascontiguousarray(v, dtype=complex),
                      "va": np.ascontiguousarray(v_adj, dtype=complex), "partner": None if partner is None else np.ascontiguousarray(partner, dtype=np.int64),
                      "cyl": int(bool(cylindrical)), "dn": None if dof_new is None else np.ascontiguousarray(dof_new, dtype=np.int32),
                      "df": None if dof_flag is None else np.ascontiguousarray(dof_flag, dtype=np.uint8),
                      "ph": None if phase is None else np.array([complex(phase)])}
        self._sens["out"] = np.zeros((len(self._sens["pts"]), 3), dtype=complex)
        self.sens_kernel_ms = 0.0

    def shape_sens_add(self, kind, ptr, elems, coef, c=None, ref_tet=0, n_ref=None, nl=0.0):
        pd, pi64, pu32, pi32, pu8 = (C.POINTER(t) for t in (C.c_double, C.c_int64, C.c_uint32, C.c_int32, C.c_uint8))
        f = C.CDLL(os.path.join(ROOT, "wavesandeigenvalues.jl_b200", "libwae_b200.so")).wae_shape_sens_check
        f.restype = C.c_int32
        f.argtypes = [C.c_int64, pd, C.c_int64, pu32, C.c_int64, pu32, C.c_int64, pi64, pi64, C.c_double, C.c_int32, pd, pd, pi32, pu8, pd,
                      C.c_int32, pi64, pi64, pd, C.c_int32, pd, C.c_int64, pd, C.c_double, pd]
        S = self._sens
        xyz = np.ascontiguousarray(self.P.T)
        tets, tris = np.ascontiguousarray(self.tets, dtype=np.uint32), np.ascontiguousarray(self.tris, dtype=np.uint32)
        ptr, elems = np.ascontiguousarray(ptr, dtype=np.int64), np.ascontiguousarray(elems, dtype=np.int64)
        cpe = 1
        if c is not None:
            c = np.ascontiguousarray(c, dtype=np.float64)
            cpe = 1 if c.ndim == 1 else c.shape[1]
        cf = np.array([complex(coef)])
        nr = None if n_ref is None else np.ascontiguousarray(n_ref, dtype=np.float64)
        P = lambda a, t: None if a is None else a.ctypes.data_as(t)
        rc = f(xyz.shape[0], P(xyz, pd), len(tets), P(tets, pu32), len(tris), P(tris, pu32), len(S["pts"]), P(S["pts"], pi64), P(S["partner"], pi64),
               S["h"], S["cyl"], P(S["v"], pd), P(S["va"], pd), P(S["dn"], pi32), P(S["df"], pu8), P(S["ph"], pd), int(kind), P(ptr, pi64), P(elems, pi64),
               P(c, pd), cpe, P(cf, pd), int(ref_tet), P(nr, pd), float(nl), P(S["out"], pd))
        if rc != 0:
            raise _lib.WaeError(rc, "wae_shape_sens_check")
        self.launches += 1

    def shape_sens_end(self):
        return self._sens["out"].T.copy()
