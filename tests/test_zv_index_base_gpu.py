"""The ABI in 1-based (Julia) index mode -- wae_create(&h, device, 1), the mode INTEGRATION.md binds: every index array that crosses the
boundary (element DOF lists, element ids, the reference tetrahedron, CSC colptr / rowval in both directions) is handed over as Julia holds
it, and the results equal those of a 0-based context on the same mesh; the pattern that comes back is Julia's SparseMatrixCSC (colptr,
rowval) of sparse(I, J, V) of the oracle's triplets, 1-based."""
import numpy as np
import pytest

from cases import GAMMA, N_REF, RHO, Q02U0, X_REF, load_raw_mesh, speedofsound

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("order", ["lin", "quad"])
def test_one_based_context(order):
    import scipy.sparse as sp

    import wae_b200 as W
    from oracle import helmholtz as ohelm
    from oracle import mesh as omesh
    from wae_b200 import _lib
    raw = load_raw_mesh("rijke_mm")
    mesh = W.Mesh("m", scale=0.001, raw=raw)
    tris, tets, dim = W.aggregate_elements(mesh, order)
    c = mesh.generate_field(speedofsound)
    outlet = np.asarray(mesh.domains["Outlet"]["simplices"], dtype=np.int64)
    flame = np.asarray(mesh.domains["Flame"]["simplices"], dtype=np.int64)
    ref = mesh.find_tetrahedron_containing_point(X_REF)
    c_tri = c[mesh.link_triangles_to_tetrahedra()]
    res = {}
    for base in (0, 1):
        ctx = _lib.Context(0, base=base)
        try:
            ctx.mesh_set(1 if order == "lin" else 2, mesh.points.T, tets + base, tris + base, dim)
            pid, nnz = ctx.pattern_build(3, None)
            im, ik = ctx.assemble_mk(pid, c)
            pb, nb = ctx.pattern_build(2, outlet + base)
            ic = ctx.assemble(pb, _lib.OP_BOUNDARY, c_tri[outlet])
            pq, iq, nq = ctx.assemble_flame(flame + base, ref + base, X_REF, N_REF, (GAMMA - 1) / RHO * Q02U0 / mesh.compute_size("Flame"))
            r = {"mk": ctx.pattern_get(pid, dim, nnz), "M": ctx.mat_get(im), "K": ctx.mat_get(ik), "b": ctx.pattern_get(pb, dim, nb),
                 "C": ctx.mat_get(ic), "q": ctx.pattern_get(pq, dim, nq), "Q": ctx.mat_get(iq)}
            # a user matrix in CSC with the context's base, a family over (K, M, user) and one solve of (K + 1e5 M + user) x = b
            cp, rv = r["mk"]
            pu, iu = ctx.mat_set(dim, cp, rv, (1.0 + 0.5j) * r["M"])
            r["u"] = ctx.pattern_get(pu, dim, nnz)
            fid, _ = ctx.family_create([ik, im, iu])
            ctx.combine(fid, np.array([1.0, 1e5, 2.0], dtype=complex), 0)
            lid, _, _ = ctx.lu_analyze(fid)
            ctx.lu_factor(lid, 0)
            rhs = np.random.default_rng(4).standard_normal(dim) + 0j
            r["x"] = ctx.lu_solve(lid, rhs)
            res[base] = r
        finally:
            ctx.close()
    for key in ("mk", "b", "q", "u"):
        for a0, a1 in zip(res[0][key], res[1][key]):
            assert np.array_equal(a0 + 1, a1), key
    for key in ("M", "K"):
        assert np.array_equal(res[0][key], res[1][key]), key
    # (the boundary mass goes through the atomic kernel: the order of the additions is not fixed, equal to rounding)
    assert np.abs(res[0]["C"] - res[1]["C"]).max() <= 1e-14 * np.abs(res[0]["C"]).max()
    # (the flame source and the solve sweeps sum with atomics: equal to rounding -- times the condition number for x --, not bitwise)
    assert np.abs(res[0]["Q"] - res[1]["Q"]).max() <= 1e-12 * np.abs(res[0]["Q"]).max()
    assert np.abs(res[0]["x"] - res[1]["x"]).max() <= 1e-9 * np.abs(res[0]["x"]).max()
    # Julia's SparseMatrixCSC of the oracle's triplets: colptr and rowval as the reference holds them (1-based)
    mo = omesh.Mesh("m", scale=0.001, raw=raw)
    trip = {}
    ohelm.discretize(mo, {"Interior": ("interior", ())}, mo.generate_field(speedofsound), order=order, triplets=trip)
    I, J, V = trip["M"][0]
    A = sp.csc_matrix((np.ones(len(I)), (np.asarray(I), np.asarray(J))), shape=(dim, dim))
    A.sum_duplicates()
    A.sort_indices()
    cp1, rv1 = res[1]["mk"]
    assert np.array_equal(cp1, A.indptr.astype(np.int64) + 1) and np.array_equal(rv1, A.indices.astype(np.int64) + 1)
