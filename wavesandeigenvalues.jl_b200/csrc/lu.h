// Sparse LU of a family slot: symbolic analysis on the host (once per pattern), numeric
// multifrontal factorisation and triangular solves on the device.  Replaces the UMFPACK
// calls of the reference (perturbation.jl:329,359; beyn.jl:65; inside Arpack.eigs at
// Householder.jl:100-101).  Internal interface.
#pragma once
#include <array>
#include <map>

#include "wae_internal.h"

#define WAE_LU_NB 32  // pivot block width of the blocked partial factorisation

// Result of the symbolic phase.  "Positions" are indices in the elimination order.
struct LuSymbolic {
  int64_t n = 0;
  std::vector<int32_t> perm;   // perm[pos] = original index
  std::vector<int32_t> iperm;  // iperm[orig] = pos
  int nsn = 0;                 // supernodes, numbered in elimination (post) order
  std::vector<int32_t> sn_first;   // nsn+1: pivot positions of supernode k are [sn_first[k], sn_first[k+1])
  std::vector<int32_t> sn_parent;  // assembly-tree parent (-1 root)
  std::vector<int32_t> sn_depth;   // root = 0
  std::vector<int64_t> struct_ptr; // nsn+1
  std::vector<int32_t> struct_idx; // row structure below the pivot block (positions, ascending)
  std::vector<int64_t> rel_ptr;    // nsn+1 (== struct_ptr): position of struct rows in the parent's front
  std::vector<int32_t> rel_idx;
  std::vector<int64_t> lp_off, up_off;  // offsets (complex units) of the L panel and U^T panel in the factor array
  std::vector<int64_t> upd_off;         // offset of the r x r update matrix inside its depth-level buffer
  std::vector<int64_t> dinv_off;        // offset of the inverted diagonal blocks (per block: L_kk^-1, U_kk^-1, NB x NB each)
  int64_t dinv_size = 0;
  std::vector<int64_t> level_upd_size;  // per depth: complex entries of all update matrices of that depth
  std::vector<std::vector<int32_t>> levels;  // supernodes by depth
  int64_t fac_size = 0;        // complex entries of the factor array (L panels + U^T panels)
  int64_t factor_nnz = 0;      // structural nonzeros of L+U (supernodal, incl. dense pivot blocks)
  double flops = 0;            // real flops of the numeric factorisation (8 per complex multiply-add)
  std::vector<int64_t> amap;   // per nonzero of A: destination offset in the factor array
  std::vector<int32_t> diagpos;  // nz index of A(i,i) in the pattern, -1 if structurally absent
  int max_s = 0, max_r = 0;
};

// coords: optional n x 3 node coordinates (nullptr -> BFS level structure is used for the bisection)
void wae_lu_symbolic(int64_t n, const int64_t* colptr, const int32_t* rowval, const double* coords, int leaf_size,
                     LuSymbolic& S);

struct LuSolver {
  int fam = -1;
  LuSymbolic sym;
  // device copies of the symbolic data
  DevBuf<int32_t> d_perm, d_iperm, d_sn_first, d_sn_parent, d_struct_idx, d_rel_idx, d_diagpos;
  DevBuf<int64_t> d_struct_ptr, d_lp_off, d_up_off, d_upd_off, d_dinv_off, d_amap;
  DevBuf<int32_t> d_colidx_nz;         // column index of every nonzero of A (for the scaling in the scatter)
  std::vector<DevBuf<int32_t>> d_level;  // supernode ids per depth
  std::vector<int32_t> level_half[2];      // per depth, concatenated: the even / odd entries of the size-sorted list (two front groups)
  std::vector<int64_t> level_half_ptr[2];  // depth d: [ptr[d], ptr[d + 1])
  DevBuf<int32_t> d_level_half[2];
  std::vector<DevBuf<int32_t>> d_xa_tile_ptr;  // per depth: extend-add tile prefix over the supernodes of that depth
  std::vector<int32_t> xa_tiles;       // per depth: number of extend-add tiles
  // extend-add in rounds: round k holds the k-th child of every parent, so no two supernodes of a round write the same front and the
  // additions are plain read-modify-writes (no atomics).  Per depth: the children in round order, per round its offset, count and tiles
  struct XaRound {
    int32_t off = 0, count = 0, tiles = 0;
  };
  std::vector<std::vector<XaRound>> xa_rounds;          // [depth of the children][round]
  std::vector<DevBuf<int32_t>> d_xa_round_children;     // per depth: children in round order
  std::vector<DevBuf<int32_t>> d_xa_round_tile_ptr;     // per depth: per round a tile prefix (count + 1 entries each), concatenated
  DevBuf<cplx> d_Aval;                 // copy of the factorised matrix (iterative refinement)
  DevBuf<cplx> d_Aval_csr;             // the same values in CSR order, built by the first refinement product y = A x after a factorisation
  bool aval_csr_valid = false;
  // numeric
  DevBuf<cplx> d_fac;
  DevBuf<cplx> d_dinv;                 // explicit inverses of the NB x NB diagonal blocks (triangular solves without a serial chain)
  DevBuf<cplx> d_upd[2];
  DevBuf<double> d_scale;              // equilibration D (A_s = D A D)
  DevBuf<cplx> d_work;                 // solve workspace (n x nrhs) x 3
  DevBuf<cplx> d_io;                   // cached staging buffer of wae_lu_solve / wae_beyn_moments
  DevBuf<cplx> d_arn_V, d_arn_w, d_arn_t, d_arn_c;  // cached Arnoldi workspace of wae_eigs_si
  DevBuf<cplx> d_arn2_V, d_arn2_c, d_arn_pair;      // second Krylov basis and the n x 2 operand of wae_eigs_si_pair
  DevBuf<double> d_arn2_dots;
  DevBuf<double> d_arn_dots;
  DevBuf<int32_t> d_flag;              // device-side status (bad pivot)
  DevBuf<int32_t> d_wininv_items;      // (supernode, block) work items of the window inverses, grouped by the number of blocks below
  int wininv_ptr[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // items with exactly c blocks below: [wininv_ptr[c], wininv_ptr[c + 1]), c = 1 .. 7
  bool wininv = false;                 // the last factorisation stored the window inverses (the solves use the product kernels)
  // CUDA graphs of the sweep pairs: the launch sequence of a sweep pair depends only on the symbolic structure, the number of right-hand
  // sides, the transposition and the kernel switches, so it is captured once per such key and replayed (lu_numeric.cu: lu_sweeps)
  struct SweepGraph {
    void* exec = nullptr;  // cudaGraphExec_t
    int64_t launches = 0;
  };
  std::map<std::array<int64_t, 8>, SweepGraph> sweep_graphs;
  ~LuSolver();
  int work_nrhs = 0;
  bool factored = false;
  bool check_singular = true;          // raise WAE_E_SINGULAR on an exactly zero pivot (wae_lu_factor_ex with check = 0 clears it)
  bool sym_mode = false;               // last factorisation used the symmetric elimination
  // rank-k correction A = S + sum_j f_j s_j g_j^T (flame terms) on top of the symmetric factorisation of S
  int r1_k = 0;
  DevBuf<cplx> d_Sval;                 // values of the symmetric part
  DevBuf<cplx> d_r1_Sm, d_r1_Gm;       // n x k dense copies of the s_j / g_j vectors
  DevBuf<cplx> d_r1_Z, d_r1_Zt;        // S^-1 Sm, S^-1 Gm
  DevBuf<cplx> d_r1_ZZ;                // both, side by side: the right-hand sides of the one solve that produces them
  DevBuf<cplx> d_r1_Kinv, d_r1_KinvT;  // (F^-1 + Gm^T Z)^-1 and its transpose, k x k
  DevBuf<cplx> d_r1_t;                 // k x nrhs scratch
  int refine_steps = 1;   // iterative refinement steps of wae_lu_solve / wae_beyn_moments
  int eigs_refine = 0;    // ... inside the Arnoldi operator (goldens G1-G6 hold to 1e-10 without; WAE_EIGS_REFINE overrides)
  double pivot_eps = 1e-30;  // only exact zeros are replaced: near-singular L(omega) is the normal case close to an eigenvalue
};

// numeric phase (lu_numeric.cu) -- all on the context stream, device pointers
void wae_lu_setup_device(wae_ctx* h, LuSolver& S);
// d_Aval: values that are factorised; d_full (may be NULL = d_Aval): the matrix the solves refer to (iterative refinement);
// sym: d_Aval is complex symmetric -> LDL^T-type elimination (half of the GEMM work)
void wae_lu_factor_device(wae_ctx* h, LuSolver& S, const cplx* d_Aval, const cplx* d_full, int sym);
// X <- (LU)^-1 X (tt = 0) or (LU)^-T X (tt = 1): factors only, no rank-k correction, no refinement
void wae_lu_base_solve(wae_ctx* h, LuSolver& S, int tt, int nrhs, cplx* d_X);
// X (n x nrhs, column-major, device) <- op(A)^{-1} X; trans: 0 N, 1 T, 2 C
// Beyn: A_p += w z^p x for p < n_mom, fused into the last kernel of the solve (A: dim x nrhs x n_mom complex on the device)
struct LuMomentEpilogue {
  int n_mom;
  cplx w, z;
  cplx* A;
};
void wae_lu_solve_device(wae_ctx* h, LuSolver& S, int trans, int nrhs, cplx* d_X, int refine, const LuMomentEpilogue* ep = nullptr);
// X2 (n x 2): column 0 <- A^{-1} x0, column 1 <- A^{-H} x1 in one pass over a symmetric-mode factor (no refinement)
void wae_lu_solve_pair_device(wae_ctx* h, LuSolver& S, cplx* d_X2);
