"""ctypes binding of libwae_b200.so (include/wae_b200.h).  Fails loudly when the CUDA library is missing:
there is no CPU fallback in this package."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwae_b200.so")

OK, E_INVALID, E_CUDA, E_SINGULAR, E_NOCONV, E_NOMEM = 0, -1, -2, -3, -4, -5
OP_MASS, OP_STIFF, OP_BOUNDARY = 1, 2, 3
FAMILY_SLOTS = 4


class WaeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libwae_b200 error {code}: {msg}")
        self.code = code


class SingularException(WaeError):
    """Raised on WAE_E_SINGULAR; the reference's solvers map it to flag -6 / itsol_singular_exception."""


class ArpackException(WaeError):
    """Raised on WAE_E_NOCONV; the reference's solvers map it to flag -4 / itsol_arpack_exception."""


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  This package has no CPU fallback.")
        _lib = C.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


_i32, _i64, _dbl, _vp = C.c_int32, C.c_int64, C.c_double, C.c_void_p
_pi32, _pi64, _pd = C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_double)
_pu32 = C.POINTER(C.c_uint32)

SIGNATURES = {
    "wae_create": (_i32, [C.POINTER(_vp), _i32, _i32]),
    "wae_destroy": (_i32, [_vp]),
    "wae_last_error": (C.c_char_p, [_vp]),
    "wae_set_stream": (_i32, [_vp, _vp]),
    "wae_sync": (_i32, [_vp]),
    "wae_launch_count": (_i64, [_vp]),
    "wae_last_ms": (_dbl, [_vp, C.c_char_p]),
    "wae_mesh_set": (_i32, [_vp, _i32, _i64, _pd, _i64, _pu32, _i64, _pu32, _i64]),
    "wae_mesh_update_points": (_i32, [_vp, _i64, _pd]),
    "wae_pattern_build": (_i32, [_vp, _i32, _i64, _pi64, _pi32, _pi64]),
    "wae_pattern_get": (_i32, [_vp, _i32, _pi64, _pi64]),
    "wae_assemble": (_i32, [_vp, _i32, _i32, _pd, _i32, _dbl, _pi32]),
    "wae_assemble_mk": (_i32, [_vp, _i32, _pd, _i32, _pi32, _pi32]),
    "wae_assemble_flame": (_i32, [_vp, _i64, _pi64, _i64, _pd, _pd, _dbl, _pi32, _pi32, _pi64]),
    "wae_assemble_bloch": (_i32, [_vp, _i32, _i64, _pi64, _i32, _pd, _i32, _dbl, _i64, _pi64, C.POINTER(C.c_uint8), _i32, _pi32, _pi32]),
    "wae_mat_info": (_i32, [_vp, _i32, _pi32, _pi32, _pi64]),
    "wae_mat_get": (_i32, [_vp, _i32, _pd]),
    "wae_mat_set": (_i32, [_vp, _i64, _pi64, _pi64, _pd, _pi32, _pi32]),
    "wae_mat_free": (_i32, [_vp, _i32]),
    "wae_family_create": (_i32, [_vp, _i32, _pi32, _pi32, _pi64]),
    "wae_family_pattern_get": (_i32, [_vp, _i32, _pi64, _pi64]),
    "wae_combine": (_i32, [_vp, _i32, _pd, _i32]),
    "wae_family_get": (_i32, [_vp, _i32, _i32, _pd]),
    "wae_family_spmm": (_i32, [_vp, _i32, _i32, _i32, _i32, _pd, _pd]),
    "wae_lu_analyze": (_i32, [_vp, _i32, _pi32, _pi64, _pd]),
    "wae_lu_factor": (_i32, [_vp, _i32, _i32]),
    "wae_lu_factor_ex": (_i32, [_vp, _i32, _i32, _i32]),
    "wae_lu_solve": (_i32, [_vp, _i32, _i32, _i32, _pd]),
    "wae_lu_free": (_i32, [_vp, _i32]),
    "wae_family_free": (_i32, [_vp, _i32]),
    "wae_pattern_free": (_i32, [_vp, _i32]),
    "wae_eigs_si": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _pd, _pd, _pd, _pi32]),
    "wae_eigs_si_pair": (_i32, [_vp, _i32, _i32, _i32, _i32, _pd, _pd, _pd, _pd, _pd, _pd, _pi32]),
    "wae_beyn_moments": (_i32, [_vp, _i32, _i32, _i32, _pd, _pd, _pd, _i32, _i32, _pd, _vp]),
    "wae_beyn_moments_multi": (_i32, [_i32, C.POINTER(_vp), _pi32, _pi32, _i32, _pd, _pd, _pd, _i32, _i32, _pd, _pd]),
    "wae_assemble_wallsrc": (_i32, [_vp, _i64, _pi64, _pd, _i32, _pd]),
    "wae_shape_sens_begin": (_i32, [_vp, _i64, _pi64, _pi64, _dbl, _i32, _i64, _pd, _pd, _pi64, C.POINTER(C.c_uint8), _pd]),
    "wae_shape_sens_add": (_i32, [_vp, _i32, _pi64, _pi64, _pd, _i32, _pd, _i64, _pd, _dbl]),
    "wae_shape_sens_end": (_i32, [_vp, _pd]),
    "wae_sorted_unique_simplices": (_i32, [_i64, _i32, _pi64, _pi64, _pi64, _pi64]),  # host-only: no context argument
}
# host-only diagnostics (no context, no GPU): used by the CPU tests, never by the product path
HOST_DIAGNOSTICS = ("wae_lu_symbolic_stats", "wae_pair_program_check", "wae_star_program_check", "wae_shape_sens_check")
SENS_MASS, SENS_STIFF, SENS_BOUNDARY, SENS_FLAME = 1, 2, 3, 4


def _declare(l):
    for name, (res, args) in SIGNATURES.items():
        f = getattr(l, name)
        f.restype = res
        f.argtypes = args


def _p(a, typ):
    return None if a is None else a.ctypes.data_as(typ)


def sorted_unique_simplices(simp):
    """(first, inv) of the reference's simplex numbering rule (host-side, thread-parallel sort; see include/wae_b200.h): the unique
    simplices in order are ``simp[first]``, ``inv[i]`` is the unique index of input row i."""
    simp = np.ascontiguousarray(simp, dtype=np.int64)
    n, k = simp.shape
    first = np.empty(n, dtype=np.int64)
    inv = np.empty(n, dtype=np.int64)
    nu = _i64()
    rc = lib().wae_sorted_unique_simplices(n, k, _p(simp, _pi64), _p(first, _pi64), _p(inv, _pi64), C.byref(nu))
    if rc != OK:
        raise WaeError(rc, "wae_sorted_unique_simplices: vertex ids must lie in [0, 2^32) and simplices have 2, 3 or 4 vertices")
    return first[:nu.value], inv


class Context:
    """One wae_ctx (one GPU, one host thread).  Index base 0 on the Python side; base=1 is what the Julia shim of INTEGRATION.md creates
    (element / DOF / nonzero indices cross the ABI as Julia holds them) and what tests/test_zv_index_base_gpu.py exercises."""

    def __init__(self, device=0, base=0):
        self._l = lib()
        h = _vp()
        self.base = base
        rc = self._l.wae_create(C.byref(h), device, base)
        if rc != OK:
            raise WaeError(rc, "wae_create failed (no CUDA device? this package has no CPU fallback)")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self._l.wae_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc == OK:
            return
        msg = (self._l.wae_last_error(self.h) or b"").decode(errors="replace")
        if rc == E_SINGULAR:
            raise SingularException(rc, msg)
        if rc == E_NOCONV:
            raise ArpackException(rc, msg)
        raise WaeError(rc, msg)

    # -- plumbing ------------------------------------------------------------------------------
    def set_stream(self, ptr):
        self._chk(self._l.wae_set_stream(self.h, _vp(ptr)))

    def sync(self):
        self._chk(self._l.wae_sync(self.h))

    def launch_count(self):
        return int(self._l.wae_launch_count(self.h))

    def last_ms(self, phase):
        return float(self._l.wae_last_ms(self.h, phase.encode()))

    # -- mesh / patterns / assembly --------------------------------------------------------------
    def mesh_set(self, order, xyz, tets, tris, dim):
        xyz = np.ascontiguousarray(xyz, dtype=np.float64)      # (n_pts, 3) == 3 x n_pts column-major
        tets = np.ascontiguousarray(tets, dtype=np.uint32)     # (n_tet, nloc)
        tris = np.ascontiguousarray(tris, dtype=np.uint32)     # (n_tri, nloc3)
        self._keep = (xyz, tets, tris)
        self._chk(self._l.wae_mesh_set(self.h, order, xyz.shape[0], _p(xyz, _pd), tets.shape[0], _p(tets, _pu32),
                                       tris.shape[0], _p(tris, _pu32), dim))
        self.mesh_serial = getattr(self, "mesh_serial", 0) + 1  # the context holds ONE mesh: see Discretization.reassemble
        return self.mesh_serial

    def pattern_build(self, elem_kind, elem_ids=None):
        pid, nnz = _i32(), _i64()
        if elem_ids is None:
            self._chk(self._l.wae_pattern_build(self.h, elem_kind, 0, None, C.byref(pid), C.byref(nnz)))
        else:
            ids = np.ascontiguousarray(elem_ids, dtype=np.int64)
            self._chk(self._l.wae_pattern_build(self.h, elem_kind, len(ids), _p(ids, _pi64), C.byref(pid), C.byref(nnz)))
        return pid.value, nnz.value

    def pattern_get(self, pid, dim, nnz):
        colptr = np.empty(dim + 1, dtype=np.int64)
        rowval = np.empty(nnz, dtype=np.int64)
        self._chk(self._l.wae_pattern_get(self.h, pid, _p(colptr, _pi64), _p(rowval, _pi64)))
        return colptr, rowval

    def mesh_update_points(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        self._chk(self._l.wae_mesh_update_points(self.h, xyz.shape[0], _p(xyz, _pd)))

    def assemble(self, pid, kind, c=None, scale=1.0, reuse=-1):
        mid = _i32(reuse)
        cpe = 1
        if c is not None:
            c = np.ascontiguousarray(c, dtype=np.float64)
            cpe = 1 if c.ndim == 1 else c.shape[1]
        self._chk(self._l.wae_assemble(self.h, pid, kind, _p(c, _pd), cpe, scale, C.byref(mid)))
        return mid.value

    def assemble_mk(self, pid, c, reuse=(-1, -1)):
        c = np.ascontiguousarray(c, dtype=np.float64)
        cpe = 1 if c.ndim == 1 else c.shape[1]
        im, ik = _i32(reuse[0]), _i32(reuse[1])
        self._chk(self._l.wae_assemble_mk(self.h, pid, _p(c, _pd), cpe, C.byref(im), C.byref(ik)))
        return im.value, ik.value

    def assemble_flame(self, flame_tets, ref_tet, x_ref, n_ref, nlocal, reuse=-1):
        ft = np.ascontiguousarray(flame_tets, dtype=np.int64)
        xr = np.ascontiguousarray(x_ref, dtype=np.float64)
        nr = np.ascontiguousarray(n_ref, dtype=np.float64)
        pid, mid, nnz = _i32(), _i32(reuse), _i64()
        self._chk(self._l.wae_assemble_flame(self.h, len(ft), _p(ft, _pi64), int(ref_tet), _p(xr, _pd), _p(nr, _pd), float(nlocal),
                                             C.byref(pid), C.byref(mid), C.byref(nnz)))
        return pid.value, mid.value, nnz.value

    def assemble_wallsrc(self, tri_ids, c, dim):
        """Speaker source vector over the listed triangles -> dense complex vector of length dim."""
        ids = np.ascontiguousarray(tri_ids, dtype=np.int64)
        c = np.ascontiguousarray(c, dtype=np.float64)
        cpe = 1 if c.ndim == 1 else c.shape[1]
        out = np.empty(dim, dtype=np.complex128)
        self._chk(self._l.wae_assemble_wallsrc(self.h, len(ids), _p(ids, _pi64), _p(c, _pd), cpe, _p(out, _pd)))
        return out

    def assemble_bloch(self, elem_kind, elem_ids, kind, c, scale, dim_red, dof_new, dof_flag, n_class):
        ids = None if elem_ids is None else np.ascontiguousarray(elem_ids, dtype=np.int64)
        cpe = 1
        if c is not None:
            c = np.ascontiguousarray(c, dtype=np.float64)
            cpe = 1 if c.ndim == 1 else c.shape[1]
        dn = np.ascontiguousarray(dof_new, dtype=np.int64)
        df = np.ascontiguousarray(dof_flag, dtype=np.uint8)
        pids = np.zeros(n_class, dtype=np.int32)
        mids = np.zeros(n_class, dtype=np.int32)
        self._chk(self._l.wae_assemble_bloch(self.h, elem_kind, 0 if ids is None else len(ids), _p(ids, _pi64), kind, _p(c, _pd), cpe, scale,
                                             dim_red, _p(dn, _pi64), df.ctypes.data_as(C.POINTER(C.c_uint8)), n_class, _p(pids, _pi32),
                                             _p(mids, _pi32)))
        return list(map(int, pids)), list(map(int, mids))

    def mat_info(self, mid):
        pid, cx, nnz = _i32(), _i32(), _i64()
        self._chk(self._l.wae_mat_info(self.h, mid, C.byref(pid), C.byref(cx), C.byref(nnz)))
        return pid.value, bool(cx.value), nnz.value

    def mat_get(self, mid):
        _, _, nnz = self.mat_info(mid)
        out = np.empty(nnz, dtype=np.complex128)
        self._chk(self._l.wae_mat_get(self.h, mid, _p(out, _pd)))
        return out

    def mat_set(self, dim, colptr, rowval, nzval):
        colptr = np.ascontiguousarray(colptr, dtype=np.int64)
        rowval = np.ascontiguousarray(rowval, dtype=np.int64)
        nzval = np.ascontiguousarray(nzval, dtype=np.complex128)
        pid, mid = _i32(), _i32()
        self._chk(self._l.wae_mat_set(self.h, dim, _p(colptr, _pi64), _p(rowval, _pi64), _p(nzval, _pd), C.byref(pid), C.byref(mid)))
        return pid.value, mid.value

    def mat_free(self, mid):
        self._chk(self._l.wae_mat_free(self.h, mid))

    # -- family ----------------------------------------------------------------------------------
    def family_create(self, mat_ids):
        ids = np.ascontiguousarray(mat_ids, dtype=np.int32)
        fid, nnz = _i32(), _i64()
        self._chk(self._l.wae_family_create(self.h, len(ids), _p(ids, _pi32), C.byref(fid), C.byref(nnz)))
        return fid.value, nnz.value

    def family_pattern_get(self, fid, dim, nnz):
        colptr = np.empty(dim + 1, dtype=np.int64)
        rowval = np.empty(nnz, dtype=np.int64)
        self._chk(self._l.wae_family_pattern_get(self.h, fid, _p(colptr, _pi64), _p(rowval, _pi64)))
        return colptr, rowval

    def combine(self, fid, coeffs, slot):
        cf = np.ascontiguousarray(coeffs, dtype=np.complex128)
        self._chk(self._l.wae_combine(self.h, fid, _p(cf, _pd), slot))

    def family_get(self, fid, slot, nnz):
        out = np.empty(nnz, dtype=np.complex128)
        self._chk(self._l.wae_family_get(self.h, fid, slot, _p(out, _pd)))
        return out

    def spmm(self, fid, slot, X, trans=0):
        X = np.asarray(X, dtype=np.complex128)
        one = X.ndim == 1
        Xf = np.asfortranarray(X.reshape(X.shape[0], -1))
        Y = np.empty_like(Xf, order="F")
        self._chk(self._l.wae_family_spmm(self.h, fid, slot, trans, Xf.shape[1], _p(Xf, _pd), _p(Y, _pd)))
        return Y[:, 0] if one else Y

    # -- LU / eigs / beyn ------------------------------------------------------------------------
    def lu_analyze(self, fid):
        lid, nnz, fl = _i32(), _i64(), _dbl()
        self._chk(self._l.wae_lu_analyze(self.h, fid, C.byref(lid), C.byref(nnz), C.byref(fl)))
        return lid.value, nnz.value, fl.value

    def lu_factor(self, lid, slot, check=True):
        """check=False: lu(A, check=false) of perturbation.jl:329 (a singular matrix is factorised with static pivoting instead of raising)."""
        self._chk(self._l.wae_lu_factor_ex(self.h, lid, slot, 1 if check else 0))

    def lu_solve(self, lid, B, trans=0):
        B = np.asarray(B, dtype=np.complex128)
        one = B.ndim == 1
        X = np.array(B.reshape(B.shape[0], -1), dtype=np.complex128, order="F", copy=True)
        self._chk(self._l.wae_lu_solve(self.h, lid, trans, X.shape[1], _p(X, _pd)))
        return X[:, 0] if one else X

    def lu_free(self, lid):
        self._chk(self._l.wae_lu_free(self.h, lid))

    def family_free(self, fid):
        self._chk(self._l.wae_family_free(self.h, fid))

    def pattern_free(self, pid):
        self._chk(self._l.wae_pattern_free(self.h, pid))

    def eigs_si(self, lid, fid, m_slot, nev, v0, trans=0):
        v0 = np.ascontiguousarray(v0, dtype=np.complex128)
        d = v0.shape[0]
        lam = np.empty(nev, dtype=np.complex128)
        V = np.empty((d, nev), dtype=np.complex128, order="F")
        ns = _i32()
        self._chk(self._l.wae_eigs_si(self.h, lid, fid, m_slot, trans, nev, _p(v0, _pd), _p(lam, _pd), _p(V, _pd), C.byref(ns)))
        return lam, V, ns.value

    def eigs_si_pair(self, lid, fid, m_slot, nev, v0, v0_adj):
        """eigs(A, M) and eigs(A', M') advanced together (two right-hand sides per pass over the factor) -> (lam, V, lam_adj, V_adj, solves)."""
        v0 = np.ascontiguousarray(v0, dtype=np.complex128)
        v0a = np.ascontiguousarray(v0_adj, dtype=np.complex128)
        d = v0.shape[0]
        lam, lam_a = np.empty(nev, dtype=np.complex128), np.empty(nev, dtype=np.complex128)
        V, Va = np.empty((d, nev), dtype=np.complex128, order="F"), np.empty((d, nev), dtype=np.complex128, order="F")
        ns = _i32()
        self._chk(self._l.wae_eigs_si_pair(self.h, lid, fid, m_slot, nev, _p(v0, _pd), _p(v0a, _pd), _p(lam, _pd), _p(V, _pd), _p(lam_a, _pd),
                                           _p(Va, _pd), C.byref(ns)))
        return lam, V, lam_a, Va, ns.value

    def moment_buffer(self, n_mom, l, d):
        """Zeroed (n_mom, l, d) complex tensor on this context's GPU for wae_beyn_moments (== d x l x n_mom column-major); the context is
        bound to torch's current stream so that the all-reduce that follows is ordered after the node loop."""
        import torch
        A = torch.zeros((n_mom, l, d), dtype=torch.complex128, device=f"cuda:{self.device}")
        self.set_stream(torch.cuda.current_stream().cuda_stream)
        return A

    def beyn_moments(self, fid, lid, z, w, coeffs, l, n_mom, out_ptr, V=None):
        z = np.ascontiguousarray(z, dtype=np.complex128)
        w = np.ascontiguousarray(w, dtype=np.complex128)
        cf = np.ascontiguousarray(coeffs, dtype=np.complex128)
        if V is not None:
            V = np.asfortranarray(V, dtype=np.complex128)
        self._chk(self._l.wae_beyn_moments(self.h, fid, lid, len(z), _p(z, _pd), _p(w, _pd), _p(cf, _pd), l, n_mom, _p(V, _pd), _vp(out_ptr)))

    @staticmethod
    def beyn_moments_multi(ctxs, fids, lids, z, w, coeffs, l, n_mom, dim, V=None):
        """wae_beyn_moments_multi: all nodes from ONE host process, node j on ctxs[j mod len(ctxs)] (one context per device), the partial
        moments all-reduced with NCCL inside the library.  Returns the (dim, l, n_mom) complex moments as a Fortran-ordered host array."""
        z = np.ascontiguousarray(z, dtype=np.complex128)
        w = np.ascontiguousarray(w, dtype=np.complex128)
        cf = np.ascontiguousarray(coeffs, dtype=np.complex128)
        if V is not None:
            V = np.asfortranarray(V, dtype=np.complex128)
        hs = (_vp * len(ctxs))(*[c.h for c in ctxs])
        fa = np.ascontiguousarray(fids, dtype=np.int32)
        la = np.ascontiguousarray(lids, dtype=np.int32)
        out = np.zeros((dim, l, n_mom), dtype=np.complex128, order="F")
        ctxs[0]._chk(ctxs[0]._l.wae_beyn_moments_multi(len(ctxs), hs, _p(fa, _pi32), _p(la, _pi32), len(z), _p(z, _pd), _p(w, _pd), _p(cf, _pd),
                                                        l, n_mom, _p(V, _pd), _p(out, _pd)))
        return out

    # -- shape sensitivity -----------------------------------------------------------------------
    def shape_sens_begin(self, points, step, v, v_adj, partner=None, cylindrical=False, dof_new=None, dof_flag=None, phase=None):
        pts = np.ascontiguousarray(points, dtype=np.int64)
        v = np.ascontiguousarray(v, dtype=np.complex128)
        va = np.ascontiguousarray(v_adj, dtype=np.complex128)
        par = None if partner is None else np.ascontiguousarray(partner, dtype=np.int64)
        dn = None if dof_new is None else np.ascontiguousarray(dof_new, dtype=np.int64)
        df = None if dof_flag is None else np.ascontiguousarray(dof_flag, dtype=np.uint8)
        ph = None if phase is None else np.array([complex(phase)], dtype=np.complex128)
        self._chk(self._l.wae_shape_sens_begin(self.h, len(pts), _p(pts, _pi64), _p(par, _pi64), float(step), int(bool(cylindrical)), len(v),
                                               _p(v, _pd), _p(va, _pd), _p(dn, _pi64), _p(df, C.POINTER(C.c_uint8)), _p(ph, _pd)))
        self._sens_n = len(pts)
        self.sens_kernel_ms = 0.0  # summed over the wae_shape_sens_add launches of this sequence

    def shape_sens_add(self, kind, ptr, elems, coef, c=None, ref_tet=0, n_ref=None, nl=0.0):
        ptr = np.ascontiguousarray(ptr, dtype=np.int64)
        elems = np.ascontiguousarray(elems, dtype=np.int64)
        cpe = 1
        if c is not None:
            c = np.ascontiguousarray(c, dtype=np.float64)
            cpe = 1 if c.ndim == 1 else c.shape[1]
        cf = np.array([complex(coef)], dtype=np.complex128)
        nr = None if n_ref is None else np.ascontiguousarray(n_ref, dtype=np.float64)
        self._chk(self._l.wae_shape_sens_add(self.h, kind, _p(ptr, _pi64), _p(elems, _pi64), _p(c, _pd), cpe, _p(cf, _pd), int(ref_tet),
                                             _p(nr, _pd), float(nl)))
        if len(elems):
            self.sens_kernel_ms += max(self.last_ms("shape_sens"), 0.0)

    def shape_sens_end(self):
        out = np.empty((self._sens_n, 3), dtype=np.complex128)  # == 3 x n_sp column-major
        self._chk(self._l.wae_shape_sens_end(self.h, _p(out, _pd)))
        return out.T
