/* wae_b200.h -- C ABI of libwae_b200.so, the B200 (sm_100a) implementation of the
 * data-parallel hot path of WavesAndEigenvalues.jl:
 *     Helmholtz.discretize -> LinearOperatorFamily L(z) -> householder / mslp / beyn.
 *
 * The reference is pure Julia and has no FFI for this path; these entry points are
 * what its host code would bind with `ccall` (INTEGRATION.md shows the Julia stubs).
 * Each group cites the reference code it replaces (paths relative to the reference
 * repository root).
 *
 * Conventions
 *  - Every function returns an int32 status: 0 ok, <0 error (WAE_E_*); the message
 *    is available from wae_last_error().  No exceptions cross the ABI.
 *  - Host arrays are caller-owned and never retained after the call returns.
 *    Device memory is owned by the context and released by wae_destroy().
 *  - Complex numbers are interleaved (re,im) doubles == Julia ComplexF64 / C99
 *    double _Complex / numpy complex128.
 *  - Index arrays handed in or out use the context's index base (wae_create:
 *    1 for Julia, 0 for C/Python).  Integer widths are explicit.
 *  - Matrices are CSC with sorted row indices inside columns, exactly the layout
 *    of Julia's SparseMatrixCSC produced by `sparse(I,J,V,dim,dim)`.
 *  - One host thread per context.  All work is issued on the context's stream
 *    (wae_set_stream) and every function that returns data to the host
 *    synchronises that stream before returning.
 */
#ifndef WAE_B200_H
#define WAE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wae_ctx wae_ctx;

enum {
  WAE_OK = 0,
  WAE_E_INVALID = -1,   /* bad argument / unknown id / wrong state           */
  WAE_E_CUDA = -2,      /* CUDA runtime failure                              */
  WAE_E_SINGULAR = -3,  /* zero / non-finite pivot: Julia shim throws SingularException(0)
                           (Householder.jl:145-148 -> flag -6; iterative_solvers.jl:206-209) */
  WAE_E_NOCONV = -4,    /* Arnoldi did not converge: shim throws ARPACKException
                           (Householder.jl:140-144 -> flag -4)                */
  WAE_E_NOMEM = -5
};

/* ---- context ------------------------------------------------------------------- */
int32_t wae_create(wae_ctx** out, int32_t device, int32_t index_base);
int32_t wae_destroy(wae_ctx* h);
const char* wae_last_error(wae_ctx* h);
/* Issue all later work on `cuda_stream` (a cudaStream_t; NULL = the legacy default stream). */
int32_t wae_set_stream(wae_ctx* h, void* cuda_stream);
int32_t wae_sync(wae_ctx* h);
/* Number of kernels this library has launched since creation (bench.py: gpu_launches). */
int64_t wae_launch_count(wae_ctx* h);
/* Device-side duration of the most recent call of the named phase, in ms (CUDA events on the
 * context stream).  Phases: "assemble", "combine", "factor", "solve", "spmv", "eigs".
 * Further keys (plain numbers, not times): "factor_sym" (1 if the last factorisation used the symmetric
 * elimination), "static_pivots" / "zero_pivots" (perturbed / exactly zero pivots of the last factorisation),
 * "eigs_residual" (worst relative Ritz residual of the pairs the last wae_eigs_si / wae_eigs_si_pair
 * returned), "star_*" (layout of the assembly program), "lu_trace_*" (WAE_LU_TRACE=1).  -1 if never set. */
double wae_last_ms(wae_ctx* h, const char* phase);

/* ---- mesh + element DOF lists ---------------------------------------------------
 * Replaces the per-element reads of mesh.points / tetrahedra / triangles inside the loops of
 * src/Helmholtz.jl:411-463,468-476,532-539.  `order` is 1 (:lin) or 2 (:quad); `tets` is
 * n_loc x n_tet (n_loc = 4 | 10, column-major == Julia vector-of-vectors flattened) and `tris`
 * n_loc3 x n_tri (3 | 6), as produced by aggregate_elements (src/FEM/FEM.jl:84-116);
 * xyz is 3 x n_pts column-major (== mesh.points); dim = number of DOFs.                     */
int32_t wae_mesh_set(wae_ctx* h, int32_t order, int64_t n_pts, const double* xyz,
                     int64_t n_tet, const uint32_t* tets, int64_t n_tri, const uint32_t* tris,
                     int64_t dim);

/* New vertex coordinates on the same topology (patterns, gather programs and the symbolic LU stay valid):
 * what shape_sensitivity.jl:111,120 needs when it re-runs discretize on a perturbed mesh.                */
int32_t wae_mesh_update_points(wae_ctx* h, int64_t n_pts, const double* xyz);

/* ---- sparsity pattern of one operator (symbolic, once per domain) -----------------
 * Replaces the I/J triplet growth + SparseArrays.sparse pattern merge (Helmholtz.jl:406-417,515;
 * FEM.jl:22-32 create_indices).  elem_kind: 3 = tetrahedra, 2 = triangles.  elem_ids index the
 * element list given to wae_mesh_set (context index base); NULL = all elements.             */
int32_t wae_pattern_build(wae_ctx* h, int32_t elem_kind, int64_t n_elem, const int64_t* elem_ids,
                          int32_t* pattern_id, int64_t* nnz);
int32_t wae_pattern_get(wae_ctx* h, int32_t pattern_id, int64_t* colptr /* dim+1 */, int64_t* rowval /* nnz */);

/* ---- numeric assembly ---------------------------------------------------------------
 * kind: element operator; values are produced directly on the pattern (duplicates summed).
 *   WAE_OP_MASS      int phi_i phi_j                      FEM.jl:704-738      (Helmholtz.jl:409-425)
 *   WAE_OP_STIFF     -c^2 int grad phi_i . grad phi_j     FEM.jl:1745-1874    (Helmholtz.jl:120-140,426-445)
 *   WAE_OP_BOUNDARY  -i c int_tri phi_i phi_j             FEM.jl:435-450      (Helmholtz.jl:151-171,446-463)
 * c: speed of sound, c_per_elem values per element of the pattern's element list, in list order
 *    (1 = constant per element; 4 / 3 = linear, vertex values: FEM.jl:764-890, 2283-2424, 469-525);
 *    ignored (may be NULL) for WAE_OP_MASS.
 * scale: real factor applied to all values (the __aux__ term is -M: Helmholtz.jl:572).
 * mat_id is in/out: a valid id (>= 0) on entry means "overwrite this matrix in place" (re-assembly with new
 * c or new coordinates, no new allocation); pass -1 to create a matrix.                       */
enum { WAE_OP_MASS = 1, WAE_OP_STIFF = 2, WAE_OP_BOUNDARY = 3 };
int32_t wae_assemble(wae_ctx* h, int32_t pattern_id, int32_t kind, const double* c,
                     int32_t c_per_elem, double scale, int32_t* mat_id);
/* Mass and stiffness of the same tetrahedral pattern in one pass over the elements. */
int32_t wae_assemble_mk(wae_ctx* h, int32_t pattern_id, const double* c, int32_t c_per_elem,
                        int32_t* mass_id, int32_t* stiff_id);
/* Speaker source vector of discretize(...; source=true) (Helmholtz.jl:488-503 opr == :m, wallsrc :193-210; FEM.jl:2557-2589):
 * out[i] = -i * sum over the listed triangles of c * |det| * int phi_i (c_per_elem 1) or of |det| * sum_k c_k int phi_i lambda_k
 * (c_per_elem 3, vertex values).  out: dim complex values (interleaved), written in full -- the dense form of the reference's
 * sparsevec(I, V, dim); the scalar (boundary_func..., pow1) stays on the host.                  */
int32_t wae_assemble_wallsrc(wae_ctx* h, int64_t n_tri, const int64_t* tri_ids, const double* c, int32_t c_per_elem, double* out);
/* Flame-response operator Q = S (x) G (Helmholtz.jl:464-487, 19-33; FEM.jl:2429-2484):
 * S_i = sum over flame tets of int phi_i, G_j = -nlocal * grad phi_j(x_ref) . n_ref on ref_tet.
 * Builds its own (dense block) pattern.                                                      */
int32_t wae_assemble_flame(wae_ctx* h, int64_t n_flame, const int64_t* flame_tets, int64_t ref_tet,
                           const double* x_ref, const double* n_ref, double nlocal,
                           int32_t* pattern_id, int32_t* mat_id, int64_t* nnz);
/* Bloch-periodic variant (src/Bloch.jl:4-112, Helmholtz.jl:509-513,541-549): the element operator `kind` over the given
 * elements, assembled on the FOLDED unit-cell DOFs and split by the reference's rule into n_class matrices:
 * 3 = (plain, +, -), 6 = (plain, +, -, axis, +axis, -axis) when the mesh has axis DOFs, 1 = everything summed (the
 * weighting matrix).  dof_new[d] = folded index of DOF d (context index base), dof_flag[d] bit 0 = image of the Bloch
 * plane, bit 1 = axis DOF (the host derives both from mesh.dos exactly as blochify does); dim_red = folded dimension.
 * pattern_ids / mat_ids receive n_class ids.  A family built from such terms is not symmetric: the general LU is used. */
int32_t wae_assemble_bloch(wae_ctx* h, int32_t elem_kind, int64_t n_elem, const int64_t* elem_ids, int32_t kind,
                           const double* c, int32_t c_per_elem, double scale, int64_t dim_red, const int64_t* dof_new,
                           const uint8_t* dof_flag, int32_t n_class, int32_t* pattern_ids, int32_t* mat_ids);

/* Values of a matrix as complex numbers in pattern order (SparseMatrixCSC.nzval). */
int32_t wae_mat_info(wae_ctx* h, int32_t mat_id, int32_t* pattern_id, int32_t* is_complex, int64_t* nnz);
int32_t wae_mat_get(wae_ctx* h, int32_t mat_id, double* nzval_complex /* 2*nnz doubles */);
/* Upload a user matrix (any CSC, e.g. a hand-built Term): builds pattern + values. */
int32_t wae_mat_set(wae_ctx* h, int64_t dim, const int64_t* colptr, const int64_t* rowval,
                    const double* nzval_complex, int32_t* pattern_id, int32_t* mat_id);
int32_t wae_mat_free(wae_ctx* h, int32_t mat_id);

/* ---- operator family:  L(z) = sum_i f_i(z) A_i  on one shared pattern ---------------
 * Replaces LinOpFam.jl:482-529 (the n_terms allocating sparse adds).  The scalar functions
 * stay on the host (arbitrary closures); only their values cross the boundary.              */
int32_t wae_family_create(wae_ctx* h, int32_t n_terms, const int32_t* mat_ids, int32_t* fam_id,
                          int64_t* nnz_union);
int32_t wae_family_pattern_get(wae_ctx* h, int32_t fam_id, int64_t* colptr, int64_t* rowval);
/* Evaluate into one of the family's value slots (slot 0..WAE_FAMILY_SLOTS-1). coeffs: n_terms
 * interleaved complex; a term whose coefficient is exactly 0 is skipped.                    */
#define WAE_FAMILY_SLOTS 4
int32_t wae_combine(wae_ctx* h, int32_t fam_id, const double* coeffs, int32_t slot);
int32_t wae_family_get(wae_ctx* h, int32_t fam_id, int32_t slot, double* nzval_complex);
/* Y = op(slot) * X, X and Y dim x nrhs column-major complex host arrays.  trans: 0 N, 1 T, 2 C. */
int32_t wae_family_spmm(wae_ctx* h, int32_t fam_id, int32_t slot, int32_t trans, int32_t nrhs,
                        const double* X, double* Y);

/* ---- sparse LU of a family slot -------------------------------------------------------
 * Replaces SparseArrays.lu / `\` (UMFPACK) at perturbation.jl:329,359, beyn.jl:65 and inside
 * Arpack.eigs (Householder.jl:100-101).  Symbolic analysis (nested dissection, supernodes,
 * frontal structure) runs once per family on the host; numeric factorisation and the
 * triangular solves run on the device.                                                      */
int32_t wae_lu_analyze(wae_ctx* h, int32_t fam_id, int32_t* lu_id, int64_t* factor_nnz, double* factor_flops);
int32_t wae_lu_factor(wae_ctx* h, int32_t lu_id, int32_t slot);
/* check = 1: as wae_lu_factor -- an exactly zero pivot (the matrix is singular for an LU without row exchanges; UMFPACK would raise
 * SingularException or pivot) returns WAE_E_SINGULAR instead of factors that solve to garbage; and when pivots were perturbed (tiny but
 * not zero), a probe solve A x = (1, ..., 1) with one refinement step follows: a relative residual above 1e-6 returns WAE_E_SINGULAR as
 * well ("factor_probe_residual" in wae_last_ms).  check = 0: the reference's
 * lu(L(0,0), check=false) of the deliberately singular operator in perturb (perturbation.jl:329): tiny and zero pivots are replaced
 * (static pivoting) and the factors are kept.                                                                                   */
int32_t wae_lu_factor_ex(wae_ctx* h, int32_t lu_id, int32_t slot, int32_t check);
/* In-place solve op(A) X = B; trans as above; X dim x nrhs column-major complex host array. */
int32_t wae_lu_solve(wae_ctx* h, int32_t lu_id, int32_t trans, int32_t nrhs, double* X);

/* ---- release (the Julia side would call these from finalizers: SparseMatrixCSC / UMFPACK factors are garbage-collected in the
 * reference, e.g. the `lu` objects of beyn.jl:65 and perturbation.jl:329 die at the end of their scope).  Everything a handle owns is
 * released by wae_destroy at the latest.  Order: LU handles of a family, then the family, then its matrices (wae_mat_free), then
 * their patterns; each call fails with WAE_E_INVALID (and releases nothing) while something still refers to the object.  Ids are
 * never reused.  wae_lu_free returns the factor storage (the largest allocation of the path) to the device.                  */
int32_t wae_lu_free(wae_ctx* h, int32_t lu_id);
int32_t wae_family_free(wae_ctx* h, int32_t fam_id);
int32_t wae_pattern_free(wae_ctx* h, int32_t pattern_id);

/* Host-only (needs no GPU and no context): the reference's simplex numbering rule for n simplices of k = 2, 3 or 4 vertices
 * (simp: n x k, row-major, ids in [0, 2^32)): unique vertex sets in ascending order of their DESCENDING-sorted vertex tuple, the
 * first occurrence of a set is its representative (src/Mesh/sorter.jl:9-31; Meshutils.jl:92-165; collect_lines! :831-840).
 * first[u] (u < *n_unique, array of n) = input row of unique simplex u, inv[i] (n) = unique index of input row i.
 * Replaces the reference's O(n^2) list insertion with one thread-parallel sort.                                 */
int32_t wae_sorted_unique_simplices(int64_t n, int32_t k, const int64_t* simp, int64_t* first, int64_t* inv, int64_t* n_unique);

/* Host-only diagnostic (needs no GPU and no context): run the symbolic phase on a CSC pattern (0-based) and
 * report out[0..7] = supernodes, nnz(L+U), factorisation flops, largest pivot block, largest row structure,
 * tree depth, factor array entries, largest per-depth update buffer.  coords (n x 3) may be NULL.          */
int32_t wae_lu_symbolic_stats(int64_t n, const int64_t* colptr, const int64_t* rowval, const double* coords,
                              int32_t leaf_size, double* out);

/* Host-only diagnostic (needs no GPU and no context): build the sparsity pattern and the owner-computes pair program of the
 * M/K assembly kernel for a tetrahedral mesh (order 1: 4, order 2: 10 DOFs per element, 0-based, n_loc x n_tet) and replay the
 * kernel's three passes (slot scatter, fixed-order summation, chunked stores) on the host with synthetic element entries.
 * out[0..7] = nnz, patches, staged elements, sources, summation units, largest slot count, max |error| against the plain
 * triplet sum, number of violated invariants (every nonzero written exactly once, slots used once, alignment).          */
int32_t wae_pair_program_check(int32_t order, int64_t n_pts, const double* xyz, int64_t n_tet, const uint32_t* tets,
                               int32_t slot_cap, double* out);

/* Host-only diagnostic (needs no GPU and no context): build the sparsity pattern and the STAR program of the third-generation
 * M/K assembly kernel (csrc/assembly_star*.c*) for a tetrahedral mesh and replay the kernel's three passes (geometry, star sums
 * and roles, chunked stores) on the host with the kernel's own arithmetic, c constant per element (n_tet values).  Returns the
 * CSC pattern (colptr: dim + 1, rowval / val_m / val_k: up to nnz_cap entries, 0-based) and stats[0..7] = nnz, patches, staged
 * elements, sub-simplices, star sources, shared memory of one CTA (bytes), program size (bytes), number of violated invariants
 * (every nonzero written exactly once, alignment of the staged pieces, bounds); stats[8..11] = simulated shared-memory wavefronts of
 * the star pass's gram gathers and their conflict-free count, the same for the store pass's record gathers (stats holds 16 doubles). */
int32_t wae_star_program_check(int32_t order, int64_t n_pts, const double* xyz, int64_t n_tet, const uint32_t* tets, const double* c,
                               int64_t smem_budget, int64_t nnz_cap, int64_t* colptr, int32_t* rowval, double* val_m, double* val_k,
                               double* stats);

/* ---- shift-invert Arnoldi: nev eigenpairs of A v = lambda M v nearest 0 ----------------
 * Replaces Arpack.eigs(A,M;nev,sigma=0,v0) (Householder.jl:100-101, iterative_solvers.jl:132-133).
 * A is the matrix factorised in lu_id, M the family slot m_slot.  trans=2 gives the adjoint
 * problem eigs(A',M').  lam: nev complex; V: dim x nev complex.                               */
int32_t wae_eigs_si(wae_ctx* h, int32_t lu_id, int32_t fam_id, int32_t m_slot, int32_t trans,
                    int32_t nev, const double* v0, double* lam, double* V, int32_t* n_solves);

/* eigs(A, M) and eigs(A', M') of one householder / mslp iteration (Householder.jl:100-101; iterative_solvers.jl:132-133) advanced together:
 * every Arnoldi step applies A^{-1} M to the direct basis vector and A^{-H} M^H to the adjoint one as the two right-hand sides of ONE pass
 * over the factor.  Same arguments and results as two wae_eigs_si calls (trans 0 with v0 -> lam, V; trans 2 with v0_adj -> lam_adj, V_adj).
 * Needs a factorisation in symmetric mode (complex-symmetric part + rank-k flame correction); returns WAE_E_INVALID otherwise and the
 * caller uses wae_eigs_si twice.                                                                                                  */
int32_t wae_eigs_si_pair(wae_ctx* h, int32_t lu_id, int32_t fam_id, int32_t m_slot, int32_t nev, const double* v0, const double* v0_adj,
                         double* lam, double* V, double* lam_adj, double* V_adj, int32_t* n_solves);


/* ---- Beyn moments --------------------------------------------------------------------
 * Replaces the integrand/gauss loop of beyn.jl:62-74,112-138 for the nodes handed in:
 *   A_p += w_j z_j^p L(z_j)^{-1} V,  p = 0..n_mom-1, V = first l identity columns (beyn.jl:45-48) or the
 *   caller's (random, beyn.jl:42-43) probing matrix.
 * coeffs: n_nodes x n_terms complex (host-evaluated term scalars at each node).
 * A_out: device pointer (dim x l x n_mom complex, column-major) accumulated in place, so that
 * the caller can all-reduce it over ranks with NCCL (torch.distributed).                    */
int32_t wae_beyn_moments(wae_ctx* h, int32_t fam_id, int32_t lu_id, int32_t n_nodes,
                         const double* z, const double* w, const double* coeffs,
                         int32_t l, int32_t n_mom, const double* V /* dim x l complex host array, NULL = identity columns */,
                         void* A_out_device);

/* The same for ALL GPUs of the box from one host process -- the reference's beyn is one process with one loop over all quadrature
 * nodes (beyn.jl:34-74, 112-138), so a single ccall must scale.  hs[0..n_ctx) are contexts on n_ctx different devices, each holding
 * the same family (fam_ids[r]) and an analysed LU of it (lu_ids[r]) -- the caller builds them with one loop over the devices
 * (INTEGRATION.md).  Node j goes to context j mod n_ctx; every context runs its nodes on its own host thread; the partial moments
 * are summed with ONE ncclAllReduce over NVLink inside the library (NCCL is loaded at run time: libnccl.so.2 or WAE_NCCL_LIB; not
 * needed for n_ctx = 1) and returned in the HOST array A_out (dim x l x n_mom complex, column-major).  The moment update
 * A_p += w z^p X is fused into the last kernel of every node's solve.  Errors of a device are reported on hs[0] (wae_last_error). */
int32_t wae_beyn_moments_multi(int32_t n_ctx, wae_ctx* const* hs, const int32_t* fam_ids, const int32_t* lu_ids, int32_t n_nodes,
                               const double* z, const double* w, const double* coeffs, int32_t l, int32_t n_mom,
                               const double* V /* dim x l complex host array, NULL = identity columns */, double* A_out);

/* ---- discrete-adjoint shape sensitivity ------------------------------------------------
 * Replaces the loop of discrete_adjoint_shape_sensitivity (src/shape_sensitivity.jl:16-141; non-unit meshes): per surface
 * point and coordinate the reference calls discretize() twice on a mesh whose domains are cut down to the simplices touching
 * the point (:50-69,:111,:120) and evaluates  sens = -v_adj' * (D_right(w0) - D_left(w0)) / (2h) * v  (:125,:129).  Here one
 * kernel thread per (point, coordinate) evaluates the element matrices of those simplices at the two positions and contracts
 * the difference with the eigenvectors; no matrix is formed.  First-order meshes only (the reference's call has no `order`).
 *   begin: n_sp moved points (context index base), step h, v and v_adj (vdim complex, normalised by the caller, :21-25).
 *          Unit-cell meshes (:84-118; b = :b evaluated at params[b] = 1): cylindrical != 0 moves every point along its local
 *          cylindrical basis (get_cylindrics, :361-370), partner[s] (or < base: none) is the image of a Bloch-plane point that
 *          moves with it, dof_new / dof_flag (per mesh point: folded DOF, bit 0 image, bit 1 axis -- the arrays given to
 *          wae_assemble_bloch) fold the DOFs as blochify does, vdim is the folded dimension and phase = exp(i 2 pi / DOS);
 *          entries touching an axis DOF are dropped (their class scalar delta(b) is 0 at b = 1) and the axis penalty term D is
 *          not evaluated (both eigenvectors vanish there).  NULL for partner / dof_new / dof_flag / phase = plain mesh.
 *   add:   one descriptor term.  ptr (n_sp+1, 0-based offsets) / elems (context index base): for every moved point the
 *          simplices of the term's domain that touch it (tet_mask / tri_mask cut with the domain, :50-69), tetrahedra for
 *          MASS / STIFF / FLAME, triangles for BOUNDARY; c: speed of sound per list entry (c_per_elem 1, or 4 / 3 vertex
 *          values) for STIFF / BOUNDARY; coef: the term's scalar at w0 (one complex number: w0^2, 1, w0*Y, n*exp(-i w0 tau));
 *          FLAME: ref_tet, n_ref and nl = (gamma-1)/rho*nglobal -- the kernel divides by the volume of the listed flame
 *          tetrahedra, which is what compute_size! returns on the reduced domain (Helmholtz.jl:325).
 *   end:   sens (3 x n_sp complex, column-major) = sum over the added terms.                                              */
enum { WAE_SENS_MASS = 1, WAE_SENS_STIFF = 2, WAE_SENS_BOUNDARY = 3, WAE_SENS_FLAME = 4 };
int32_t wae_shape_sens_begin(wae_ctx* h, int64_t n_sp, const int64_t* points, const int64_t* partner, double step, int32_t cylindrical,
                             int64_t vdim, const double* v, const double* v_adj, const int64_t* dof_new, const uint8_t* dof_flag,
                             const double* phase);
int32_t wae_shape_sens_add(wae_ctx* h, int32_t kind, const int64_t* ptr, const int64_t* elems, const double* c, int32_t c_per_elem,
                           const double* coef, int64_t ref_tet, const double* n_ref, double nl);
int32_t wae_shape_sens_end(wae_ctx* h, double* sens);
/* Host-only diagnostic (needs no GPU and no context; used by the CPU tests, never by the product path): the per-thread
 * function of the kernel above evaluated in a plain host loop.  0-based indices, first-order connectivity (4 x n_tet,
 * 3 x n_tri); ACCUMULATES one term into sens (3 x n_sp complex).                                                          */
int32_t wae_shape_sens_check(int64_t n_pts, const double* xyz, int64_t n_tet, const uint32_t* tets, int64_t n_tri, const uint32_t* tris,
                             int64_t n_sp, const int64_t* points, const int64_t* partner, double step, int32_t cylindrical, const double* v,
                             const double* v_adj, const int32_t* dof_new, const uint8_t* dof_flag, const double* phase, int32_t kind,
                             const int64_t* ptr, const int64_t* elems, const double* c, int32_t c_per_elem, const double* coef,
                             int64_t ref_tet, const double* n_ref, double nl, double* sens);

#ifdef __cplusplus
}
#endif
#endif /* WAE_B200_H */
