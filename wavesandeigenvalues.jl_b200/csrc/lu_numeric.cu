// Numeric phase of the sparse complex LU on the device (multifrontal, supernodal).
//
// Storage per supernode k with s pivots and r structure rows (ld = s + r), all column-major complex:
//   Lp  (ld x s)  rows 0..s-1: pivot block (L strictly below the diagonal of each NBxNB diagonal block and
//                 in all blocks below it, U on/above the diagonal inside the diagonal blocks), rows s..: L21
//   Up  (ld x s)  the U factor stored TRANSPOSED: diagonal blocks hold U_kk^T (lower incl. diagonal), blocks
//                 below hold U[k, J]^T, rows s.. hold U12^T.  L and U^T panels therefore have the same shape and
//                 the same kernels serve A x = b (forward with Lp, backward with Up) and A^T x = b (forward
//                 with Up, backward with Lp).
//   A22 (r x r)   Schur complement / update matrix, lives in a per-depth scratch buffer until the parent
//                 has absorbed it (extend-add).
// The assembly tree is processed by depth, deepest level first; all supernodes of one depth are independent
// and are handled by batched kernels (grid.z = supernode).  The frontal updates are complex GEMMs
// C -= A B^T on the FP64 tensor cores (mma.sync m8n8k4 f64 -> SASS DMMA.8x8x4).
// No pivoting across blocks (the symbolic structure is static); tiny pivots are replaced (static
// pivoting) and the solve applies iterative refinement against the original matrix.
#include <cuda_profiler_api.h>
#include <cuda_runtime.h>

#include <array>

#include <algorithm>

#include "lu.h"

#define NB WAE_LU_NB

struct SnView {
  int s, r, ld, first;
  cplx *lp, *up;
};

struct LuDev {  // plain device pointers handed to the kernels
  const int32_t* sn_first;
  const int64_t* struct_ptr;
  const int32_t* struct_idx;
  const int32_t* rel_idx;
  const int32_t* sn_parent;
  const int64_t *lp_off, *up_off, *upd_off, *dinv_off;
  cplx* fac;
  cplx* dinv;
};

__device__ __forceinline__ SnView sn_view(const LuDev& D, int k) {
  SnView v;
  v.first = D.sn_first[k];
  v.s = D.sn_first[k + 1] - v.first;
  v.r = (int)(D.struct_ptr[k + 1] - D.struct_ptr[k]);
  v.ld = v.s + v.r;
  v.lp = D.fac + D.lp_off[k];
  v.up = D.fac + D.up_off[k];
  return v;
}

__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cinv(cplx a) {
  double d = 1.0 / (a.x * a.x + a.y * a.y);
  return make_double2(a.x * d, -a.y * d);
}
__device__ __forceinline__ void catomic_sub(cplx* p, cplx v) {
  atomicAdd(&p->x, -v.x);
  atomicAdd(&p->y, -v.y);
}

// ---- equilibration + scatter of A into the fronts ----------------------------------------------
__global__ void lu_scale_kernel(const cplx* __restrict__ A, const int32_t* __restrict__ diagpos, int64_t n, double* __restrict__ d) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 1.0;
  int32_t p = diagpos[i];
  if (p >= 0) {
    cplx a = A[p];
    double m = sqrt(a.x * a.x + a.y * a.y);
    if (m > 0.0 && isfinite(m)) s = rsqrt(m);
  }
  d[i] = s;
}

__global__ void lu_scatter_kernel(const cplx* __restrict__ A, const int64_t* __restrict__ amap, const int32_t* __restrict__ rowidx,
                                  const int32_t* __restrict__ colidx, const double* __restrict__ d, int64_t nnz, cplx* __restrict__ fac) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  cplx a = A[k];
  double s = d[rowidx[k]] * d[colidx[k]];
  fac[amap[k]] = make_double2(a.x * s, a.y * s);
}

// ---- extend-add: child update matrices into the parent fronts ------------------------------------
// grid.x = tiles over all children of the level (tile_ptr prefix), block 32x8
// sym != 0: the update matrices hold their lower triangle only (x >= y); upper entries are read from the mirror position
// and only the lower part of the parent front is written (its U^T panel is regenerated from the L panel).
__global__ void __launch_bounds__(256) lu_extend_add_kernel(LuDev D, const int32_t* __restrict__ children, const int32_t* __restrict__ tile_ptr,
                                                            int nchild, const cplx* __restrict__ upd_child, cplx* __restrict__ upd_parent, int sym, int part, int atomic) {
  // part 0: every target; 1: targets in the parent's pivot panels only; 2: targets in the parent's update matrix only
  // locate the child of this tile
  int t = blockIdx.x;
  int lo = 0, hi = nchild;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (tile_ptr[mid] <= t) lo = mid; else hi = mid;
  }
  const int c = children[lo];
  const int p = D.sn_parent[c];
  if (p < 0) return;
  const int rc = (int)(D.struct_ptr[c + 1] - D.struct_ptr[c]);
  const int nt = (rc + 31) >> 5;
  const int tl = t - tile_ptr[lo];
  const int ti = tl % nt, tj = tl / nt;
  const int32_t* rel = D.rel_idx + D.struct_ptr[c];
  const cplx* U = upd_child + D.upd_off[c];
  SnView P = sn_view(D, p);
  cplx* A22 = upd_parent + D.upd_off[p];
  const int x = ti * 32 + threadIdx.x;
  if (x >= rc) return;
  const int ia = rel[x];
  for (int yy = threadIdx.y; yy < 32; yy += 8) {
    int y = tj * 32 + yy;
    if (y >= rc) break;
    int jb = rel[y];
    cplx* dst;
    if (jb < P.s) {
      if (part == 2) continue;
      if (ia >= P.s || ia / NB >= jb / NB) dst = P.lp + ia + (size_t)jb * P.ld;
      else if (sym) continue;
      else dst = P.up + jb + (size_t)ia * P.ld;
    } else {
      if (ia < P.s) {
        if (sym || part == 2) continue;
        dst = P.up + jb + (size_t)ia * P.ld;
      } else {
        if (part == 1 || (sym && ia < jb)) continue;
        dst = A22 + (ia - P.s) + (size_t)(jb - P.s) * P.r;
      }
    }
    cplx v = (sym && x < y) ? U[y + (size_t)x * rc] : U[x + (size_t)y * rc];
    if (atomic) {
      atomicAdd(&dst->x, v.x);
      atomicAdd(&dst->y, v.y);
    } else {  // a round of the extend-add: this supernode is the only writer of its parent's front
      cplx o = *dst;
      o.x += v.x;
      o.y += v.y;
      *dst = o;
    }
  }
}

// ---- step k, part 1: LU of the NB x NB diagonal block (no pivoting, static perturbation) ------------
// One thread per entry, ONE barrier per pivot: every thread forms its multiplier l_i = T[i][p] / pivot itself from the unscaled column
// (the owner of L[i][p] stores it one barrier later, when nobody reads column p any more), and the same row operations are applied to
// an identity block in the same step, which yields L^-1 for free (Gauss-Jordan); U^-1 comes out of the same sweep through the
// transposed problem (see below).  Both inverses are kept: the panel and the triangular solves multiply by them.
__global__ void __launch_bounds__(NB * NB) lu_diag_kernel(LuDev D, const int32_t* __restrict__ list, int k, double eps, int* __restrict__ flag) {
  SnView S = sn_view(D, list[blockIdx.x]);
  const int c0 = k * NB;
  if (c0 >= S.s) return;
  const int nb = min(NB, S.s - c0);
  __shared__ cplx T[NB][NB + 1];  // in place: strict lower = L, upper incl. diagonal = U
  __shared__ cplx X[NB][NB + 1];  // strict lower = L^-1 (unit diagonal implied); upper incl. diagonal = U^-1, built as the transpose
                                  // of Z = (U^T)^-1: Z[a][b] (a >= b) lives at X[b][a]
  __shared__ cplx piv[NB];        // pivots after the static perturbation
  const int i = threadIdx.x, j = threadIdx.y;  // element (row i, col j)
  const cplx zero = make_double2(0.0, 0.0), one = make_double2(1.0, 0.0);
  cplx* blk = S.lp + c0 + (size_t)c0 * S.ld;
  T[i][j] = (i < nb && j < nb) ? blk[i + (size_t)j * S.ld] : (i == j ? one : zero);  // identity padding of a short last block
  X[i][j] = i == j ? one : zero;
  __syncthreads();
  // One sweep, one barrier per pivot.  Row p of U is final when step p starts, and so is column p of W = U^T (W[i][p] = U[p][i]): the
  // forward Gauss-Jordan step of the LOWER triangular W on an identity block -- row p scaled by 1 / pivot, rows below minus W[i][p] times
  // it -- runs in the same step as the elimination of column p of T and the Gauss-Jordan step for L^-1 (round 2a: a second sweep of NB
  // barriers for U^-1).  Every thread forms its multiplier / scaled entry itself from the unscaled values; the owners store them one
  // barrier later, when nobody reads them any more.  The pivot of step p + 1 (static perturbation included) and its reciprocal are
  // formed by the ONE thread that owns T[p+1][p+1], inside step p, and read by everybody after the barrier -- instead of 1024 threads
  // each doing the double-precision division.
  __shared__ cplx pinv[NB];  // reciprocals of the pivots
  auto make_pivot = [&](int q, cplx d) {
    const double m = d.x * d.x + d.y * d.y;
    if (!(m >= eps * eps)) {  // tiny, zero or NaN pivot -> static pivoting
      if (!isfinite(m)) atomicOr(flag, 2);
      else atomicAdd(flag + 1, 1);
      if (m == 0.0) atomicOr(flag, 4);  // an exactly zero pivot: structurally / exactly singular for an LU without row exchanges
      d = make_double2(m > 0.0 && isfinite(m) ? d.x * eps / sqrt(m) : eps, m > 0.0 && isfinite(m) ? d.y * eps / sqrt(m) : 0.0);
    }
    piv[q] = d;
    pinv[q] = cinv(d);
  };
  if (i == 0 && j == 0) make_pivot(0, T[0][0]);
  __syncthreads();
  cplx lown = zero, zown = zero;
  for (int p = 0; p < nb; p++) {
    const cplx di = pinv[p];
    cplx tnew = zero, xnew = zero, znew = zero;
    bool tw = false, xw = false, zw = false;
    if (i > p) {
      const cplx l = cmul(T[i][p], di);
      if (j > p) {
        tnew = csub(T[i][j], cmul(l, T[p][j]));
        tw = true;
        if (i == p + 1 && j == p + 1 && p + 1 < nb) make_pivot(p + 1, tnew);
      } else if (j == p) {
        lown = l;
        xnew = make_double2(-l.x, -l.y);  // row i of the identity block minus l times row p (unit diagonal)
        xw = true;
      } else {
        xnew = csub(X[i][j], cmul(l, X[p][j]));
        xw = true;
      }
      if (j <= p) {  // Z[i][j] -= W[i][p] * (Z[p][j] / pivot),  W[i][p] = U[p][i] = T[p][i]
        znew = csub(X[j][i], cmul(T[p][i], cmul(X[j][p], di)));
        zw = true;
      }
    } else if (i == p && j <= p) {
      zown = cmul(X[j][p], di);  // row p of Z, scaled
    }
    // writes of this step that nobody else reads in this step (own element only): rows i > p are read by their owners alone
    if (tw) T[i][j] = tnew;
    if (xw) X[i][j] = xnew;
    if (zw) X[j][i] = znew;
    __syncthreads();
    if (j == p && i > p) T[i][p] = lown;   // column p of T is not read any more
    if (i == p && j <= p) X[j][p] = zown;  // row p of Z is not read any more
  }
  if (i == j && i < nb) T[i][i] = piv[i];
  __syncthreads();
  if (i < nb && j < nb) {
    blk[i + (size_t)j * S.ld] = T[i][j];
    // U_kk^T into the U^T panel (lower triangle incl. diagonal)
    if (i >= j) S.up[(c0 + i) + (size_t)(c0 + j) * S.ld] = T[j][i];
  }
  // stored NB x NB column-major, zero padded (both triangles share one array: strict lower = L^-1 (unit diagonal implied), upper incl.
  // diagonal = U^-1)
  cplx* inv = D.dinv + D.dinv_off[list[blockIdx.x]] + (size_t)k * 2 * NB * NB;
  const bool in = i < nb && j < nb;
  inv[i + j * NB] = i > j ? (in ? X[i][j] : zero) : (i == j && i < nb ? one : zero);
  inv[NB * NB + i + j * NB] = (i <= j && in) ? X[i][j] : zero;  // U^-1[i][j] = Z[j][i]
}

// ---- step k, part 2: panel solves below the diagonal block -------------------------------------------
//   Lp rows:  X <- X * U_kk^{-1}        Up rows:  X <- X * L_kk^{-T}  (unit diagonal)
// one thread per row, the NB row entries live in registers, the triangular factor in shared memory
// sym != 0 (complex-symmetric front, F[k,J]^T == F[J,k]): launched with grid.y == 2 as well, but the U^T-panel rows are
// computed from the (not yet solved) L-panel row copy that the symmetric copy kernel placed there.
__global__ void __launch_bounds__(128) lu_sym_copy_kernel(LuDev D, const int32_t* __restrict__ list, int k) {
  SnView S = sn_view(D, list[blockIdx.z]);
  const int c0 = k * NB;
  if (c0 >= S.s) return;
  const int nb = min(NB, S.s - c0);
  const int r0 = c0 + nb;
  const int row = blockIdx.x * 128 + threadIdx.x;
  if (row >= S.ld - r0) return;
  const cplx* src = S.lp + (r0 + row) + (size_t)c0 * S.ld;
  cplx* dst = S.up + (r0 + row) + (size_t)c0 * S.ld;
  for (int j = 0; j < nb; j++) dst[(size_t)j * S.ld] = src[(size_t)j * S.ld];
}

// sym_scale != 0 (symmetric elimination, grid.y == 1): U = D L^T, so the U^T-panel rows are the solved L-panel rows with column j scaled
// by the pivot U_jj -- no copy of the unsolved rows and no second triangular solve.
__global__ void __launch_bounds__(128) lu_panel_kernel(LuDev D, const int32_t* __restrict__ list, int k, int sym_scale) {
  SnView S = sn_view(D, list[blockIdx.z]);
  const int c0 = k * NB;
  if (c0 >= S.s) return;
  const int nb = min(NB, S.s - c0);
  const int r0 = c0 + nb;
  const int nrows = S.ld - r0;
  if ((int)(blockIdx.x * 128) >= nrows) return;
  const bool upper = blockIdx.y == 1;  // 0: Lp with U_kk, 1: Up with L_kk^T
  __shared__ cplx T[NB][NB + 1];       // T[m][j] = coefficient multiplying x_m in the equation of x_j
  __shared__ cplx idiag[NB], diag[NB];
  const cplx* blk = S.lp + c0 + (size_t)c0 * S.ld;
  for (int e = threadIdx.x; e < NB * NB; e += 128) {
    int m = e % NB, j = e / NB;
    if (m < nb && j < nb) T[m][j] = upper ? blk[j + (size_t)m * S.ld] /* L[j][m] */ : blk[m + (size_t)j * S.ld] /* U[m][j] */;
  }
  if (threadIdx.x < nb) {
    diag[threadIdx.x] = blk[threadIdx.x + (size_t)threadIdx.x * S.ld];
    idiag[threadIdx.x] = upper ? make_double2(1.0, 0.0) : cinv(diag[threadIdx.x]);
  }
  __syncthreads();
  const int row = blockIdx.x * 128 + threadIdx.x;
  if (row >= nrows) return;
  cplx* base = (upper ? S.up : S.lp) + (r0 + row) + (size_t)c0 * S.ld;
  cplx x[NB];
#pragma unroll
  for (int j = 0; j < NB; j++) x[j] = j < nb ? base[(size_t)j * S.ld] : make_double2(0.0, 0.0);
#pragma unroll
  for (int j = 0; j < NB; j++) {
    if (j < nb) {
      cplx acc = x[j];
#pragma unroll
      for (int m = 0; m < j; m++) acc = csub(acc, cmul(x[m], T[m][j]));
      x[j] = cmul(acc, idiag[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < NB; j++)
    if (j < nb) base[(size_t)j * S.ld] = x[j];
  if (sym_scale) {
    cplx* ub = S.up + (r0 + row) + (size_t)c0 * S.ld;
#pragma unroll
    for (int j = 0; j < NB; j++)
      if (j < nb) ub[(size_t)j * S.ld] = cmul(x[j], diag[j]);
  }
}

// ---- round 2: diagonal block on ONE WARP ---------------------------------------------------------------------------------------
// Phase 1 (LU): lane i holds row i of the NB x NB block in registers, the pivot row travels by shuffles, every register index is static
// (fully unrolled).  Phase 2 / 3 (L^-1, U^-1): the factors go to shared memory and lane c computes COLUMN c of the inverse by
// substitution -- every lane reads the same factor entry at the same time (broadcast), its own column stays in registers, nothing is
// written to shared memory.  No block barrier at all (round 2a: 64 barriers of a 1024-thread CTA, 34 us per block at the top of the
// tree, 1.7 ms per launch on the levels with 8000 fronts); two blocks share a 64-thread CTA.  Same LU arithmetic as lu_diag_kernel
// (multiplier l = T[i][p] / pivot, static perturbation of tiny pivots); the inverses come from substitution instead of Gauss-Jordan.
__device__ __forceinline__ cplx shfl_c(cplx v, int src) {
  return make_double2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}

#define LU_DIAG_WARPS 2
__global__ void __launch_bounds__(32 * LU_DIAG_WARPS) lu_diag_warp_kernel(LuDev D, const int32_t* __restrict__ list, int nlist, int k, double eps,
                                                                       int* __restrict__ flag) {
  __shared__ cplx Fs[LU_DIAG_WARPS][NB][NB + 1];  // the factors, row-major: strict lower = L, upper incl. diagonal = U
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int item = blockIdx.x * LU_DIAG_WARPS + wp;
  if (item >= nlist) return;
  const int sn = list[item];
  SnView S = sn_view(D, sn);
  const int c0 = k * NB;
  if (c0 >= S.s) return;
  const int nb = min(NB, S.s - c0);
  const cplx zero = make_double2(0.0, 0.0), one = make_double2(1.0, 0.0);
  cplx* blk = S.lp + c0 + (size_t)c0 * S.ld;
  cplx (*F)[NB + 1] = Fs[wp];
  cplx* inv = D.dinv + D.dinv_off[sn] + (size_t)k * 2 * NB * NB;
  {
    cplx T[NB];
#pragma unroll
    for (int j = 0; j < NB; j++)
      T[j] = (lane < nb && j < nb) ? blk[lane + (size_t)j * S.ld] : (lane == j ? one : zero);  // identity padding of a short last block
#pragma unroll
    for (int p = 0; p < NB; p++) {
      cplx d = shfl_c(T[p], p);
      const double m = d.x * d.x + d.y * d.y;
      if (!(m >= eps * eps)) {  // tiny, zero or NaN pivot -> static pivoting (every lane derives the same replacement)
        if (lane == p) {
          if (!isfinite(m)) atomicOr(flag, 2);
          else atomicAdd(flag + 1, 1);
          if (m == 0.0) atomicOr(flag, 4);
        }
        d = make_double2(m > 0.0 && isfinite(m) ? d.x * eps / sqrt(m) : eps, m > 0.0 && isfinite(m) ? d.y * eps / sqrt(m) : 0.0);
        if (lane == p) T[p] = d;
      }
      const bool below = lane > p;
      const cplx l = cmul(T[p], cinv(d));
#pragma unroll
      for (int j = p + 1; j < NB; j++) {
        const cplx t = shfl_c(T[j], p);
        if (below) T[j] = csub(T[j], cmul(l, t));
      }
      if (below) T[p] = l;
    }
    // L \ U back into the panel, U_kk^T into the U^T panel (lower triangle incl. diagonal: lane i owns row i of U = column i there)
#pragma unroll
    for (int j = 0; j < NB; j++) {
      F[lane][j] = T[j];
      if (lane < nb && j < nb) {
        blk[lane + (size_t)j * S.ld] = T[j];
        if (j >= lane) S.up[(c0 + j) + (size_t)(c0 + lane) * S.ld] = T[j];
      }
    }
  }
  __syncwarp();
  {
    // column c = lane of L^-1 (unit lower): x_c = 1, x_i = -sum_{c <= kk < i} L[i][kk] x_kk
    cplx x[NB];
#pragma unroll
    for (int i = 0; i < NB; i++) {
      cplx acc = zero;
#pragma unroll
      for (int kk = 0; kk < i; kk++) {
        const cplx f = F[i][kk];  // broadcast
        if (kk >= lane) {
          acc.x -= f.x * x[kk].x - f.y * x[kk].y;
          acc.y -= f.x * x[kk].y + f.y * x[kk].x;
        }
      }
      x[i] = i == lane ? one : acc;  // rows above the diagonal: acc == 0
    }
    // inv[i + c NB], zero padded (unit diagonal stored for i < nb)
#pragma unroll
    for (int i = 0; i < NB; i++) inv[i + lane * NB] = (i < nb && lane < nb && i >= lane) ? x[i] : zero;
  }
  {
    // column c = lane of U^-1 (upper): x_i = (delta_ic - sum_{i < kk <= c} U[i][kk] x_kk) / U[i][i], from i = c upwards
    cplx x[NB];
#pragma unroll
    for (int i = NB - 1; i >= 0; i--) {
      cplx acc = i == lane ? one : zero;
#pragma unroll
      for (int kk = i + 1; kk < NB; kk++) {
        const cplx f = F[i][kk];
        if (kk <= lane) {
          acc.x -= f.x * x[kk].x - f.y * x[kk].y;
          acc.y -= f.x * x[kk].y + f.y * x[kk].x;
        }
      }
      x[i] = i <= lane ? cmul(acc, cinv(F[i][i])) : zero;
    }
#pragma unroll
    for (int i = 0; i < NB; i++) inv[NB * NB + i + lane * NB] = (i < nb && lane < nb && i <= lane) ? x[i] : zero;
  }
}

// ---- round 2: panel rows times the explicit inverse of the diagonal block ---------------------------------------------------------
//   Lp rows:  X <- X * U_kk^-1        Up rows (general elimination only):  X <- X * L_kk^-T
// One thread per row as in lu_panel_kernel, but a PRODUCT with the inverse the diagonal-block kernel keeps instead of a substitution: the
// 528 multiply-adds of a row are independent accumulations (the substitution is a chain of 496 dependent ones: 19 us per launch on the
// top levels, where the step is pure latency).
__global__ void __launch_bounds__(128) lu_panel_inv_kernel(LuDev D, const int32_t* __restrict__ list, int k, int sym_scale) {
  const int sn = list[blockIdx.z];
  SnView S = sn_view(D, sn);
  const int c0 = k * NB;
  if (c0 >= S.s) return;
  const int nb = min(NB, S.s - c0);
  const int r0 = c0 + nb;
  const int nrows = S.ld - r0;
  if ((int)(blockIdx.x * 128) >= nrows) return;
  const bool upper = blockIdx.y == 1;
  __shared__ cplx M[NB][NB];  // M[m][j]: x_new[j] = sum_{m <= j} x[m] M[m][j]
  __shared__ cplx diag[NB];
  const int row = blockIdx.x * 128 + threadIdx.x;
  const bool live = row < nrows;
  cplx* base = (upper ? S.up : S.lp) + (r0 + (live ? row : 0)) + (size_t)c0 * S.ld;
  cplx x[NB];
#pragma unroll
  for (int m = 0; m < NB; m++) x[m] = (live && m < nb) ? base[(size_t)m * S.ld] : make_double2(0.0, 0.0);
  const cplx* inv = D.dinv + D.dinv_off[sn] + (size_t)k * 2 * NB * NB;
  for (int e = threadIdx.x; e < NB * NB; e += 128) {
    const int m = e % NB, j = e / NB;
    // Lp: U^-1[m][j] at NB^2 + m + j NB;  Up: L^-T[m][j] = L^-1[j][m] at j + m NB
    M[m][j] = upper ? inv[j + m * NB] : inv[NB * NB + m + j * NB];
  }
  if (threadIdx.x < NB) diag[threadIdx.x] = threadIdx.x < nb ? S.lp[(c0 + threadIdx.x) + (size_t)(c0 + threadIdx.x) * S.ld] : make_double2(0.0, 0.0);
  __syncthreads();
  if (!live) return;
  cplx* ub = S.up + (r0 + row) + (size_t)c0 * S.ld;
#pragma unroll
  for (int j = 0; j < NB; j++) {
    cplx a0 = make_double2(0.0, 0.0), a1 = make_double2(0.0, 0.0);  // two chains per column
#pragma unroll
    for (int m = 0; m <= j; m++) {
      const cplx c = M[m][j];
      if (m & 1) {
        a1.x += x[m].x * c.x - x[m].y * c.y;
        a1.y += x[m].x * c.y + x[m].y * c.x;
      } else {
        a0.x += x[m].x * c.x - x[m].y * c.y;
        a0.y += x[m].x * c.y + x[m].y * c.x;
      }
    }
    if (j < nb) {
      const cplx acc = make_double2(a0.x + a1.x, a0.y + a1.y);
      base[(size_t)j * S.ld] = acc;
      if (sym_scale) ub[(size_t)j * S.ld] = cmul(acc, diag[j]);
    }
  }
}

// ---- complex GEMM  C -= A * B^T  on the FP64 tensor cores -----------------------------------------------
// A: m x K (lda), B: n x K (ldb), C: m x n (ldc), all column-major complex.  CTA tile 64x64, 8 warps (4 along m,
// 2 along n), warp tile 16x32 = 2x4 DMMA m8n8k4 tiles, K tile 16, operands split into re/im planes in smem.
#define GT 64
#define GK 16
#define GLD 72  // padded leading dimension: (kk*GLD + mm) mod 16 hits every value twice -> conflict-free LDS.64

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

struct GemmProblem {
  int m, n, K, lda, ldb, ldc;
  bool beta0;  // C = -A B^T (C is not read: the Schur complement of a front whose update matrix receives its children afterwards)
  const cplx *A, *B;
  cplx* C;
};

__device__ __forceinline__ void zgemm_nt_tile(const GemmProblem& P, int tile_m, int tile_n) {
  __shared__ double As[2][GK][GLD], Bs[2][GK][GLD];  // [re/im][k][row]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp & 3) * 16, wn = (warp >> 2) * 32;
  const int m0 = tile_m * GT, n0 = tile_n * GT;
  double acc_r[2][4][2], acc_i[2][4][2];
#pragma unroll
  for (int a = 0; a < 2; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) acc_r[a][b][0] = acc_r[a][b][1] = acc_i[a][b][0] = acc_i[a][b][1] = 0.0;
  const int lr = tid & 63, lk = tid >> 6;  // loader: row lr, k-columns lk, lk+4, lk+8, lk+12
  for (int k0 = 0; k0 < P.K; k0 += GK) {
#pragma unroll
    for (int q = 0; q < 4; q++) {
      int kk = lk + 4 * q;
      cplx a = make_double2(0.0, 0.0), b = make_double2(0.0, 0.0);
      if (k0 + kk < P.K) {
        if (m0 + lr < P.m) a = P.A[(m0 + lr) + (size_t)(k0 + kk) * P.lda];
        if (n0 + lr < P.n) b = P.B[(n0 + lr) + (size_t)(k0 + kk) * P.ldb];
      }
      As[0][kk][lr] = a.x; As[1][kk][lr] = a.y;
      Bs[0][kk][lr] = b.x; Bs[1][kk][lr] = b.y;
    }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < GK; ks += 4) {
      const int kk = ks + (lane & 3), rr = lane >> 2;
      double ar[2], ai[2], br[4], bi[4];
#pragma unroll
      for (int a = 0; a < 2; a++) {
        ar[a] = As[0][kk][wm + a * 8 + rr];
        ai[a] = As[1][kk][wm + a * 8 + rr];
      }
#pragma unroll
      for (int b = 0; b < 4; b++) {
        br[b] = Bs[0][kk][wn + b * 8 + rr];
        bi[b] = Bs[1][kk][wn + b * 8 + rr];
      }
#pragma unroll
      for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
          dmma(acc_r[a][b][0], acc_r[a][b][1], ar[a], br[b]);
          dmma(acc_r[a][b][0], acc_r[a][b][1], -ai[a], bi[b]);
          dmma(acc_i[a][b][0], acc_i[a][b][1], ar[a], bi[b]);
          dmma(acc_i[a][b][0], acc_i[a][b][1], ai[a], br[b]);
        }
    }
    __syncthreads();
  }
  // C -= acc ; fragment: row = lane/4, cols = (lane%4)*2 + {0,1}
#pragma unroll
  for (int a = 0; a < 2; a++)
#pragma unroll
    for (int b = 0; b < 4; b++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        int row = m0 + wm + a * 8 + (lane >> 2), col = n0 + wn + b * 8 + (lane & 3) * 2 + e;
        if (row < P.m && col < P.n) {
          cplx* c = P.C + row + (size_t)col * P.ldc;
          cplx v = P.beta0 ? make_double2(0.0, 0.0) : *c;
          v.x -= acc_r[a][b][e];
          v.y -= acc_i[a][b][e];
          *c = v;
        }
      }
}

// Opt-in variant of the tile (WAE_LU_GEMM=2; written without GPU access, timed and residual-checked by tools/bench_lu_knobs.py): the
// operands stay interleaved complex in shared memory and arrive by cp.async (16 bytes per request, zero fill outside the matrices) through a
// ring of GP_STAGES K tiles, so that the global loads of the next tiles are in flight during the DMMAs of the current one -- the
// single-buffered tile above exposes the full load latency on the K = 32 / K = 128 pivot-block updates (one to eight K tiles per launch).
// Fragments come out of shared memory as one LDS.128 per complex number; leading dimension 66: (kk * 66 + row) mod 8 runs through all
// eight 16-byte bank groups over the eight lanes of a quarter warp (kk = lane & 3, row = lane >> 2) -> conflict-free.
// Same products in the same order per accumulator as the tile above: the results are bitwise equal.
#define GP_STAGES 3
#define GP_LD 66
#define GP_SMEM (GP_STAGES * 2 * GK * GP_LD * (int)sizeof(cplx))

__device__ __forceinline__ void cp_async16_zfill(void* smem, const void* gmem, bool valid) {
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem);
  const int bytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem), "r"(bytes) : "memory");
}

__device__ __forceinline__ void zgemm_nt_tile_async(const GemmProblem& P, int tile_m, int tile_n) {
  extern __shared__ __align__(16) unsigned char gp_smem[];
  typedef cplx (*Stage)[GK][GP_LD];
  Stage As = reinterpret_cast<Stage>(gp_smem);                                                 // [stage][k][row]
  Stage Bs = reinterpret_cast<Stage>(gp_smem + (size_t)GP_STAGES * GK * GP_LD * sizeof(cplx));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp & 3) * 16, wn = (warp >> 2) * 32;
  const int m0 = tile_m * GT, n0 = tile_n * GT;
  double acc_r[2][4][2], acc_i[2][4][2];
#pragma unroll
  for (int a = 0; a < 2; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) acc_r[a][b][0] = acc_r[a][b][1] = acc_i[a][b][0] = acc_i[a][b][1] = 0.0;
  const int lr = tid & 63, lk = tid >> 6;  // loader: row lr, k-columns lk, lk+4, lk+8, lk+12
  const bool a_row = m0 + lr < P.m, b_row = n0 + lr < P.n;
  const cplx* a_src = P.A + (a_row ? m0 + lr : 0);
  const cplx* b_src = P.B + (b_row ? n0 + lr : 0);
  auto load_stage = [&](int kt, int buf) {
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int kk = lk + 4 * q, k = kt * GK + kk;
      const bool kin = k < P.K;
      cp_async16_zfill(&As[buf][kk][lr], a_src + (size_t)(kin ? k : 0) * P.lda, kin && a_row);
      cp_async16_zfill(&Bs[buf][kk][lr], b_src + (size_t)(kin ? k : 0) * P.ldb, kin && b_row);
    }
  };
  const int nk = (P.K + GK - 1) / GK;
#pragma unroll
  for (int s = 0; s < GP_STAGES - 1; s++) {
    if (s < nk) load_stage(s, s);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int kt = 0; kt < nk; kt++) {
    asm volatile("cp.async.wait_group %0;" ::"n"(GP_STAGES - 2) : "memory");  // K tile kt has landed (this thread's part)
    __syncthreads();  // ... everybody's part; and the buffer refilled below was read for the last time in iteration kt - 1
    if (kt + GP_STAGES - 1 < nk) load_stage(kt + GP_STAGES - 1, (kt + GP_STAGES - 1) % GP_STAGES);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const int buf = kt % GP_STAGES;
#pragma unroll
    for (int ks = 0; ks < GK; ks += 4) {
      const int kk = ks + (lane & 3), rr = lane >> 2;
      cplx av[2], bv[4];
#pragma unroll
      for (int a = 0; a < 2; a++) av[a] = As[buf][kk][wm + a * 8 + rr];
#pragma unroll
      for (int b = 0; b < 4; b++) bv[b] = Bs[buf][kk][wn + b * 8 + rr];
#pragma unroll
      for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
          dmma(acc_r[a][b][0], acc_r[a][b][1], av[a].x, bv[b].x);
          dmma(acc_r[a][b][0], acc_r[a][b][1], -av[a].y, bv[b].y);
          dmma(acc_i[a][b][0], acc_i[a][b][1], av[a].x, bv[b].y);
          dmma(acc_i[a][b][0], acc_i[a][b][1], av[a].y, bv[b].x);
        }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
  for (int a = 0; a < 2; a++)
#pragma unroll
    for (int b = 0; b < 4; b++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        int row = m0 + wm + a * 8 + (lane >> 2), col = n0 + wn + b * 8 + (lane & 3) * 2 + e;
        if (row < P.m && col < P.n) {
          cplx* c = P.C + row + (size_t)col * P.ldc;
          cplx v = P.beta0 ? make_double2(0.0, 0.0) : *c;
          v.x -= acc_r[a][b][e];
          v.y -= acc_i[a][b][e];
          *c = v;
        }
      }
}

// mode 0: Lp trailing pivot columns, mode 1: Up trailing pivot columns, mode 2: Schur complement.
// Modes 0/1 update C = X[c0:ld, c0:min(s,cap)] -= X[c0:ld, k0:k0+kw] * Y[c0:min(s,cap), k0:k0+kw]^T  (X,Y = Lp,Up or Up,Lp).
// Two-level blocking: inside an outer block of NBO columns the NB-wide steps only touch the columns of that outer block
// (kw = NB, cap = end of the outer block); the rest of the pivot block is updated once per outer block with kw = NBO.
template <int PIPE>  // 0: single-buffered tile (the measured kernel), 1: cp.async ring (opt-in, WAE_LU_GEMM=2)
__global__ void __launch_bounds__(256) lu_gemm_kernel(LuDev D, const int32_t* __restrict__ list, int mode, int k0, int kw, int c0, int cap,
                                                      cplx* __restrict__ upd, int sym) {
  const int sn = list[blockIdx.z];
  SnView S = sn_view(D, sn);
  GemmProblem P;
  P.beta0 = false;
  if (mode == 2) {
    P.beta0 = (sym & 4) != 0;
    if (S.r == 0) return;
    if ((sym & 1) && blockIdx.y > blockIdx.x) return;  // symmetric Schur complement: lower tiles only
    P.m = P.n = S.r; P.K = S.s;
    P.lda = P.ldb = S.ld; P.ldc = S.r;
    P.A = S.lp + S.s; P.B = S.up + S.s;
    P.C = upd + D.upd_off[sn];
  } else {
    const int cend = min(S.s, cap);
    if (c0 >= cend) return;  // no trailing pivot columns in range
    // sym bit 1 (WAE_LU_SKIP_UPPER, opt-in): tiles strictly above the diagonal of the pivot block are never read again -- the U part of
    // the pivot block lives (transposed) in the other panel and the NB x NB diagonal blocks sit in the tiles with equal row and column index
    // (both tile grids start at c0, a multiple of NB, and GT = 2 NB) -- so they need no update: half of the pivot-square GEMM work
    if ((sym & 2) && blockIdx.x < blockIdx.y) return;
    P.m = S.ld - c0; P.n = cend - c0; P.K = kw;
    P.lda = P.ldb = P.ldc = S.ld;
    cplx* X = mode == 0 ? S.lp : S.up;
    cplx* Y = mode == 0 ? S.up : S.lp;
    P.A = X + c0 + (size_t)k0 * S.ld;
    P.B = Y + c0 + (size_t)k0 * S.ld;
    P.C = X + c0 + (size_t)c0 * S.ld;
  }
  if ((int)(blockIdx.x * GT) >= P.m || (int)(blockIdx.y * GT) >= P.n) return;
  if constexpr (PIPE == 0)
    zgemm_nt_tile(P, blockIdx.x, blockIdx.y);
  else
    zgemm_nt_tile_async(P, blockIdx.x, blockIdx.y);
}

// ---- triangular solves ---------------------------------------------------------------------------------
// Forward substitution with a lower-trapezoidal panel P (Lp: unit diagonal, Up: general diagonal), backward substitution with its
// transpose.  The pivot block of a supernode is cut into windows of LU_SOLVE_W columns; per window two kernels run:
//   tri     (one CTA of 32 warps per supernode and right-hand side): the W x W triangle of the window, blocked by NB.  The
//           diagonal blocks were inverted at factorisation time, so a block step is a 32x32 matrix-vector product (warp w = row w,
//           shuffle reduction) followed by the rank-NB update of the remaining rows of the window -- no serial substitution chain.
//   update  (many CTAs per supernode, NR right-hand sides per pass over the panel): all rows below the window -- the rest of the
//           pivot block and the structure rows -- i.e. the bulk of the factor, streamed once at full width.
// Large supernodes (the top of the assembly tree, where a level has fewer supernodes than the GPU has SMs) are thereby spread
// over the whole GPU instead of one CTA each.
//   use_up = 0: T_kk^-1 = L_kk^-1            use_up = 1: T_kk = U_kk^T  ->  T_kk^-1 = (U_kk^-1)^T
#define LU_SOLVE_W 256

// Opt-in (WAE_LU_SOLVE_PF=1, bit 1 of the kernels' use_up argument; written without GPU access, timed and residual-checked by
// tools/bench_lu_knobs.py): the block steps of a window are a serial chain in which every step waits for one load of its inverse diagonal
// block and one of its panel rows; with the hint the lower triangle of the window's panel (<= 1 MB) and its inverse blocks (<= 128 KB) are
// requested into L2 at kernel start, so that the steps wait for L2 instead of HBM.  Pure hints on addresses the kernel reads anyway.
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ void tri_window_prefetch(const cplx* P, int ld, const cplx* dinv, int c_lo, int c_hi) {
  const int w = c_hi - c_lo;
  const int col = threadIdx.x >> 2;  // four threads per column of the window, 128-byte lines (8 complex) from the diagonal down
  if (col < w) {
    const cplx* cp = P + (size_t)(c_lo + col) * ld + c_lo;
    for (int r = (col & ~7) + 8 * (threadIdx.x & 3); r < w; r += 32) prefetch_l2(cp + r);
  }
  const int nblk = (w + NB - 1) / NB;  // NB * NB complex per inverse block = 128 lines
  for (int l = threadIdx.x; l < nblk * 128; l += blockDim.x) prefetch_l2(dinv + (size_t)(c_lo / NB + (l >> 7)) * 2 * NB * NB + (l & 127) * 8);
}

__global__ void __launch_bounds__(1024) lu_fwd_tri_kernel(LuDev D, const int32_t* __restrict__ list, int use_up, int c_lo, int64_t n, cplx* __restrict__ x) {
  const int sn = list[blockIdx.x];
  SnView S = sn_view(D, sn);
  if (c_lo >= S.s) return;
  const int c_hi = min(c_lo + LU_SOLVE_W, S.s);
  const int pf = use_up >> 1;
  use_up &= 1;
  const cplx* P = use_up ? S.up : S.lp;
  const cplx* dinv = D.dinv + D.dinv_off[sn] + (use_up ? NB * NB : 0);
  if (pf) tri_window_prefetch(P, S.ld, dinv, c_lo, c_hi);
  __shared__ cplx yk[NB];
  __shared__ cplx xw[LU_SOLVE_W];  // the window of x: the block steps work on shared memory only
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nrhs = gridDim.y;  // one CTA per (supernode, right-hand side); x is stored position-major: x[pos * nrhs + rhs]
  cplx* xg = x + (size_t)S.first * nrhs + blockIdx.y;
  if (threadIdx.x < c_hi - c_lo) xw[threadIdx.x] = xg[(size_t)(c_lo + threadIdx.x) * nrhs];
  __syncthreads();
  cplx* xs = xw - c_lo;  // xs[i] = entry i of the pivot block, valid for c_lo <= i < c_hi
  for (int c0 = c_lo, kb = c_lo / NB; c0 < c_hi; c0 += NB, kb++) {
    const int nb = min(NB, S.s - c0);
    {
      const cplx* Tinv = dinv + (size_t)kb * 2 * NB * NB;
      // y[warp] = sum_lane Tinv(warp, lane) x[lane]   (Tinv(r,c) stored at r + c*NB; transposed access for use_up)
      cplx m = use_up ? Tinv[lane + warp * NB] : Tinv[warp + lane * NB];
      cplx v = lane < nb ? xs[c0 + lane] : make_double2(0.0, 0.0);
      double sr = m.x * v.x - m.y * v.y, si = m.x * v.y + m.y * v.x;
#pragma unroll
      for (int off = 16; off; off >>= 1) {
        sr += __shfl_xor_sync(0xffffffffu, sr, off);
        si += __shfl_xor_sync(0xffffffffu, si, off);
      }
      if (lane == 0) yk[warp] = make_double2(sr, si);
    }
    __syncthreads();
    if (threadIdx.x < nb) xs[c0 + threadIdx.x] = yk[threadIdx.x];
    {
      // rank-NB update of the remaining rows of the window: 4 threads per row, 8 columns each (the window has at most 256 rows)
      const int i = c0 + nb + (threadIdx.x >> 2), part = (threadIdx.x & 3) * (NB / 4);
      cplx acc = make_double2(0.0, 0.0);
      if (i < c_hi) {
        const cplx* row = P + i + (size_t)(c0 + part) * S.ld;
#pragma unroll
        for (int j = 0; j < NB / 4; j++)
          if (part + j < nb) {
            const cplx a = row[(size_t)j * S.ld], v = yk[part + j];
            acc.x += a.x * v.x - a.y * v.y;
            acc.y += a.x * v.y + a.y * v.x;
          }
      }
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 1);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 1);
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 2);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 2);
      if (i < c_hi && (threadIdx.x & 3) == 0) xs[i] = csub(xs[i], acc);
    }
    __syncthreads();
  }
  if (threadIdx.x < c_hi - c_lo) xg[(size_t)(c_lo + threadIdx.x) * nrhs] = xw[threadIdx.x];
}

// x[rows below the window] -= P[rows, window] * y_window.  grid.x row chunks of 64, grid.y supernode, grid.z groups of NR
// right-hand sides; a CTA is 64 rows x 4 column quarters (the quarters are summed through shared memory), so a level with few, large
// supernodes still keeps many loads in flight.  Pivot rows are owned by one thread; structure rows are shared with sibling
// supernodes -> atomic.
template <int NR>
__global__ void __launch_bounds__(256) lu_fwd_update_kernel(LuDev D, const int32_t* __restrict__ list, int use_up, int c_lo, int nrhs, int64_t n,
                                                            cplx* __restrict__ x) {
  SnView S = sn_view(D, list[blockIdx.y]);
  if (c_lo >= S.s) return;
  const int c_hi = min(c_lo + LU_SOLVE_W, S.s);
  const int nrows = S.ld - c_hi;
  if (blockIdx.x * 64 >= nrows) return;
  const int r = threadIdx.x & 63, quarter = threadIdx.x >> 6;
  const int i = blockIdx.x * 64 + r;
  const int rhs0 = blockIdx.z * NR, nr = min(NR, nrhs - rhs0);
  const cplx* P = (use_up ? S.up : S.lp) + c_hi;
  const int32_t* st = D.struct_idx + D.struct_ptr[list[blockIdx.y]];
  __shared__ cplx ys[NR][LU_SOLVE_W];  // window part of x, later the partial sums of quarters 1..3 ([q][quarter-1][row])
  const int ncol = c_hi - c_lo;
  for (int j = threadIdx.x; j < ncol; j += 256)
#pragma unroll
    for (int q = 0; q < NR; q++)
      if (q < nr) ys[q][j] = x[(size_t)(S.first + c_lo + j) * nrhs + rhs0 + q];
  __syncthreads();
  cplx acc[NR];
#pragma unroll
  for (int q = 0; q < NR; q++) acc[q] = make_double2(0.0, 0.0);
  if (i < nrows) {
    const int j0 = quarter * (LU_SOLVE_W / 4), j1 = min(ncol, j0 + LU_SOLVE_W / 4);
    const cplx* row = P + i + (size_t)c_lo * S.ld;
#pragma unroll 8
    for (int j = j0; j < j1; j++) {
      const cplx a = row[(size_t)j * S.ld];
#pragma unroll
      for (int q = 0; q < NR; q++) {
        acc[q].x += a.x * ys[q][j].x - a.y * ys[q][j].y;
        acc[q].y += a.x * ys[q][j].y + a.y * ys[q][j].x;
      }
    }
  }
  __syncthreads();
  if (quarter)
#pragma unroll
    for (int q = 0; q < NR; q++) ys[q][(quarter - 1) * 64 + r] = acc[q];
  __syncthreads();
  if (quarter == 0 && i < nrows) {
    const int g = c_hi + i;
#pragma unroll
    for (int q = 0; q < NR; q++)
      if (q < nr) {
        cplx t = acc[q];
#pragma unroll
        for (int u = 0; u < 3; u++) {
          t.x += ys[q][u * 64 + r].x;
          t.y += ys[q][u * 64 + r].y;
        }
        if (g < S.s) {
          cplx* d = x + (size_t)(S.first + g) * nrhs + rhs0 + q;
          *d = csub(*d, t);
        } else
          catomic_sub(x + (size_t)st[g - S.s] * nrhs + rhs0 + q, t);
      }
  }
}

// backward: x_S[c] -= sum_{rows below the window} P[row, c] * x[row]   for the columns c of the window.
// grid.x = (column groups of 8 = 8 warps) x (row chunks), grid.y supernode, grid.z groups of NR right-hand sides
template <int NR>
__global__ void __launch_bounds__(256) lu_bwd_update_kernel(LuDev D, const int32_t* __restrict__ list, int use_up, int c_lo, int ncg, int row_chunk,
                                                            int nrhs, int64_t n, cplx* __restrict__ x) {
  SnView S = sn_view(D, list[blockIdx.y]);
  if (c_lo >= S.s) return;
  const int c_hi = min(c_lo + LU_SOLVE_W, S.s);
  const int nrows = S.ld - c_hi;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = c_lo + (blockIdx.x % ncg) * 8 + warp;  // ncg = column groups of the widest window on the level
  const int r_begin = (blockIdx.x / ncg) * row_chunk;
  if (c >= c_hi || r_begin >= nrows) return;
  const int r_end = min(nrows, r_begin + row_chunk);
  const int rhs0 = blockIdx.z * NR, nr = min(NR, nrhs - rhs0);
  const cplx* col = (use_up ? S.up : S.lp) + c_hi + (size_t)c * S.ld;
  const int32_t* st = D.struct_idx + D.struct_ptr[list[blockIdx.y]];
  double sr[NR], si[NR];
#pragma unroll
  for (int q = 0; q < NR; q++) sr[q] = si[q] = 0.0;
  for (int i = r_begin + lane; i < r_end; i += 32) {
    const cplx a = col[i];
    const int g = c_hi + i;
    const int64_t idx = g < S.s ? (int64_t)S.first + g : (int64_t)st[g - S.s];
#pragma unroll
    for (int q = 0; q < NR; q++)
      if (q < nr) {
        const cplx v = x[(size_t)idx * nrhs + rhs0 + q];
        sr[q] += a.x * v.x - a.y * v.y;
        si[q] += a.x * v.y + a.y * v.x;
      }
  }
#pragma unroll
  for (int q = 0; q < NR; q++) {
#pragma unroll
    for (int off = 16; off; off >>= 1) {
      sr[q] += __shfl_xor_sync(0xffffffffu, sr[q], off);
      si[q] += __shfl_xor_sync(0xffffffffu, si[q], off);
    }
    if (lane == 0 && q < nr) catomic_sub(x + (size_t)(S.first + c) * nrhs + rhs0 + q, make_double2(sr[q], si[q]));
  }
}

// backward, W x W triangle of the window (one CTA per supernode and right-hand side, 32 warps): blocks from last to first.
//   use_up = 1 (A x = b): x_k = U_kk^-1 (...)        use_up = 0 (A^T x = b): x_k = (L_kk^-1)^T (...)
__global__ void __launch_bounds__(1024) lu_bwd_tri_kernel(LuDev D, const int32_t* __restrict__ list, int use_up, int c_lo, int64_t n, cplx* __restrict__ x) {
  const int sn = list[blockIdx.x];
  SnView S = sn_view(D, sn);
  if (c_lo >= S.s) return;
  const int c_hi = min(c_lo + LU_SOLVE_W, S.s);
  const int pf = use_up >> 1;
  use_up &= 1;
  const cplx* P = use_up ? S.up : S.lp;
  const cplx* dinv = D.dinv + D.dinv_off[sn] + (use_up ? NB * NB : 0);
  if (pf) tri_window_prefetch(P, S.ld, dinv, c_lo, c_hi);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ cplx yk[NB];
  __shared__ cplx xw[LU_SOLVE_W];  // the window of x in shared memory
  const int nrhs = gridDim.y;
  cplx* xg = x + (size_t)S.first * nrhs + blockIdx.y;
  if (threadIdx.x < c_hi - c_lo) xw[threadIdx.x] = xg[(size_t)(c_lo + threadIdx.x) * nrhs];
  __syncthreads();
  cplx* xs = xw - c_lo;
  for (int kb = (c_hi - 1) / NB; kb >= c_lo / NB; kb--) {
    const int c0 = kb * NB, nb = min(NB, S.s - c0);
    // column c0+warp: subtract the contribution of the already solved rows of the window below this block
    {
      double sr = 0.0, si = 0.0;
      if (warp < nb) {
        const cplx* col = P + (size_t)(c0 + warp) * S.ld;
        for (int i = c0 + nb + lane; i < c_hi; i += 32) {
          cplx a = col[i], v = xs[i];
          sr += a.x * v.x - a.y * v.y;
          si += a.x * v.y + a.y * v.x;
        }
      }
#pragma unroll
      for (int off = 16; off; off >>= 1) {
        sr += __shfl_xor_sync(0xffffffffu, sr, off);
        si += __shfl_xor_sync(0xffffffffu, si, off);
      }
      if (lane == 0) yk[warp] = warp < nb ? make_double2(xs[c0 + warp].x - sr, xs[c0 + warp].y - si) : make_double2(0.0, 0.0);
    }
    __syncthreads();
    {
      // x[warp] = sum_lane M(warp, lane) y[lane], M = U_kk^-1 (use_up) or (L_kk^-1)^T
      const cplx* Tinv = dinv + (size_t)kb * 2 * NB * NB;
      cplx m = use_up ? Tinv[warp + lane * NB] : Tinv[lane + warp * NB];
      cplx v = yk[lane];
      double sr = m.x * v.x - m.y * v.y, si = m.x * v.y + m.y * v.x;
#pragma unroll
      for (int off = 16; off; off >>= 1) {
        sr += __shfl_xor_sync(0xffffffffu, sr, off);
        si += __shfl_xor_sync(0xffffffffu, si, off);
      }
      if (lane == 0 && warp < nb) xs[c0 + warp] = make_double2(sr, si);
    }
    __syncthreads();
  }
  if (threadIdx.x < c_hi - c_lo) xg[(size_t)(c_lo + threadIdx.x) * nrhs] = xw[threadIdx.x];
}

// ---- round 2: window inverses ---------------------------------------------------------------------------------
// The W x W lower triangle T of a window (blocks T_ik = P[block i, block k], k <= i) is inverted ONCE per factorisation, so that a
// window step of the solves is a matrix-vector product instead of a chain of W / NB dependent block steps (round 1: 45 us per window on
// one CTA, 40 % of a solve at the top of the tree).  The diagonal blocks of the inverse are the NB x NB inverses the diagonal-block kernel
// already keeps (dinv); block (i, j), i > j, of the inverse,
//     Inv_ij = -T_ii^-1  sum_{k = j}^{i-1}  T_ik Inv_kj ,
// is stored TRANSPOSED at block (j, i) of the same panel -- the strictly upper block triangle of the pivot block is unused by the
// factorisation (L lives below the block diagonal of Lp, U^T below the block diagonal of Up) -- so no storage is added and T itself stays
// intact.  In that position the forward product reads COLUMNS of the panel (y_rho = column rho above its block . x: contiguous) and the
// backward (transposed) product reads ROWS (thread per row: coalesced), like the panel below the window.
#define LU_WBLK (LU_SOLVE_W / NB)
#define LU_TILE_BYTES (NB * (NB + 1) * (int)sizeof(cplx))

// items: (supernode, block j) pairs; the CTA computes Inv_ij for i = j+1 .. end of j's window (exactly nI blocks), blockIdx.y = 0 Lp, 1 Up
__global__ void __launch_bounds__(NB * NB) lu_wininv_kernel(LuDev D, const int32_t* __restrict__ items, int nI, int sym_dual) {
  extern __shared__ __align__(16) unsigned char wininv_smem[];
  typedef cplx Tile[NB][NB + 1];
  Tile* Is = reinterpret_cast<Tile*>(wininv_smem);  // Is[m][b][c] = Inv_{j+m, j}[c, b], m = 0 .. nI
  Tile& Ts = Is[nI + 1];                            // Ts[c][a] = left operand [a, c]
  Tile& Ss = Is[nI + 2];                            // Ss[b][c] = sum [c, b]
  const int sn = items[2 * blockIdx.x], j = items[2 * blockIdx.x + 1];
  SnView S = sn_view(D, sn);
  const bool up = blockIdx.y == 1;
  cplx* P = up ? S.up : S.lp;
  const cplx* dinv = D.dinv + D.dinv_off[sn] + (up ? NB * NB : 0);
  const int nblk = (S.s + NB - 1) / NB;
  const int wend = min((j / LU_WBLK + 1) * LU_WBLK, nblk);
  const int tx = threadIdx.x, ty = threadIdx.y;
  const cplx zero = make_double2(0.0, 0.0);
  {
    // T_jj^-1: Lp: L^-1[tx, ty] is entry [c = tx, b = ty]; Up: T_jj^-1 = (U^-1)^T, U^-1[tx, ty] is entry [c = ty, b = tx]
    const cplx v = dinv[(size_t)j * 2 * NB * NB + tx + ty * NB];
    if (up) Is[0][tx][ty] = v; else Is[0][ty][tx] = v;
  }
  for (int i = j + 1; i < wend; i++) {
    cplx acc = zero;
    const int row = NB * i + tx;
    for (int k = j; k < i; k++) {
      __syncthreads();  // Ts free again, Is[k - j] complete
      Ts[ty][tx] = row < S.s ? P[row + (size_t)(NB * k + ty) * S.ld] : zero;  // T_ik[a = tx, c = ty]; rows past the pivot block are not part of T
      __syncthreads();
      const cplx* ik = &Is[k - j][ty][0];
#pragma unroll 8
      for (int c = 0; c < NB; c++) {
        const cplx a = Ts[c][tx], b = ik[c];
        acc.x += a.x * b.x - a.y * b.y;
        acc.y += a.x * b.y + a.y * b.x;
      }
    }
    __syncthreads();
    Ss[ty][tx] = acc;
    {
      // T_ii^-1[a, c]: Lp: L^-1[a, c] at a + c NB (a = tx, c = ty); Up: U^-1[c, a] at c + a NB (c = tx, a = ty)
      const cplx v = dinv[(size_t)i * 2 * NB * NB + tx + ty * NB];
      if (up) Ts[tx][ty] = v; else Ts[ty][tx] = v;
    }
    __syncthreads();
    cplx r = zero;
    const cplx* sb = &Ss[ty][0];
#pragma unroll 8
    for (int c = 0; c < NB; c++) {
      const cplx a = Ts[c][tx], b = sb[c];
      r.x -= a.x * b.x - a.y * b.y;
      r.y -= a.x * b.y + a.y * b.x;
    }
    Is[i - j][ty][tx] = r;  // Inv_ij[a = tx, b = ty]
  }
  __syncthreads();
  // Inv_ij[a, b] -> P[(NB j + b) + (NB i + a) ld]: b = tx (contiguous), a = ty
  // sym_dual (symmetric elimination, U = D L^T, launched for the L panel only): the inverse of the U^T window is D^-1 times the inverse of
  // the L window, i.e. the stored (transposed) block with column rho divided by the pivot U_rho,rho -- written here instead of computed
  for (int i = j + 1; i < wend; i++) {
    const int col = NB * i + ty;
    if (col < S.s) {
      const cplx v = Is[i - j][tx][ty];
      P[(NB * j + tx) + (size_t)col * S.ld] = v;
      if (sym_dual) {
        const cplx di = D.dinv[D.dinv_off[sn] + (size_t)i * 2 * NB * NB + NB * NB + ty + ty * NB];  // 1 / U_col,col
        S.up[(NB * j + tx) + (size_t)col * S.ld] = cmul(v, di);
      }
    }
  }
}

// forward product of one window: ys[q][rho] = sum_{m < NB (rho / NB)} P[c_lo + m, c_lo + rho] xs[q][m] + (T_ii^-1 xs[q][block i])[rho]
// (warp per output row, lanes along the contiguous column segment).  xs / ys: [NR][stride] in shared memory, w = width of the window.
template <int NR>
__device__ __forceinline__ void win_fwd_matvec(const cplx* __restrict__ P, int ld, const cplx* __restrict__ dinv, int use_up, int c_lo, int w,
                                               const cplx* xs, cplx* ys, int stride, int nwarps) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int rho = warp; rho < w; rho += nwarps) {
    const int i = rho / NB, a = rho % NB;
    const cplx* col = P + c_lo + (size_t)(c_lo + rho) * ld;
    double sr[NR], si[NR];
#pragma unroll
    for (int q = 0; q < NR; q++) sr[q] = si[q] = 0.0;
    for (int m = lane; m < NB * i; m += 32) {
      const cplx v = col[m];
#pragma unroll
      for (int q = 0; q < NR; q++) {
        const cplx xv = xs[q * stride + m];
        sr[q] += v.x * xv.x - v.y * xv.y;
        si[q] += v.x * xv.y + v.y * xv.x;
      }
    }
    {
      const cplx* Tinv = dinv + (size_t)(c_lo / NB + i) * 2 * NB * NB;
      const int m = NB * i + lane;
      if (m < w) {  // the inverse block is zero-padded, but the padded entries of xs are not initialised
        const cplx v = use_up ? Tinv[lane + a * NB] : Tinv[a + lane * NB];
#pragma unroll
        for (int q = 0; q < NR; q++) {
          const cplx xv = xs[q * stride + m];
          sr[q] += v.x * xv.x - v.y * xv.y;
          si[q] += v.x * xv.y + v.y * xv.x;
        }
      }
    }
#pragma unroll
    for (int q = 0; q < NR; q++) {
#pragma unroll
      for (int off = 16; off; off >>= 1) {
        sr[q] += __shfl_xor_sync(0xffffffffu, sr[q], off);
        si[q] += __shfl_xor_sync(0xffffffffu, si[q], off);
      }
      if (lane == 0) ys[q * stride + rho] = make_double2(sr[q], si[q]);
    }
  }
}

// backward (transposed) product of one window, the part one thread owns: row rho of the panel over the columns kappa = k0, k0 + kstep, ...
// of the blocks behind rho's block, plus (first == true) the diagonal block:  sum_a T_jj^-1[a, b] v[NB j + a]
template <int NR>
__device__ __forceinline__ void win_bwd_row(const cplx* __restrict__ P, int ld, const cplx* __restrict__ dinv, int use_up, int c_lo, int w, int rho,
                                            int kpart, int kstep, const cplx* vs, int stride, cplx* acc) {
  const int j = rho / NB, b = rho % NB;
  const cplx* row = P + (c_lo + rho) + (size_t)c_lo * ld;
#pragma unroll 8
  for (int kappa = NB * (j + 1) + kpart; kappa < w; kappa += kstep) {
    const cplx v = row[(size_t)kappa * ld];
#pragma unroll
    for (int q = 0; q < NR; q++) {
      const cplx xv = vs[q * stride + kappa];
      acc[q].x += v.x * xv.x - v.y * xv.y;
      acc[q].y += v.x * xv.y + v.y * xv.x;
    }
  }
  const cplx* Tinv = dinv + (size_t)(c_lo / NB + j) * 2 * NB * NB;
  const int a1 = min(NB, w - NB * j);
#pragma unroll 8
  for (int a = kpart; a < a1; a += kstep) {
    const cplx v = use_up ? Tinv[b + a * NB] : Tinv[a + b * NB];
#pragma unroll
    for (int q = 0; q < NR; q++) {
      const cplx xv = vs[q * stride + NB * j + a];
      acc[q].x += v.x * xv.x - v.y * xv.y;
      acc[q].y += v.x * xv.y + v.y * xv.x;
    }
  }
}

// All CTAs of a cluster have read the window of x before any of them overwrites its part of it
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
#define LU_WIN_CLUSTER 8

// window step of the forward sweep with the inverted window: a cluster of 8 CTAs per (supernode, group of NR <= 4 right-hand sides), one
// output row per warp (CTA c, warp p: row c + 8 p, so that every CTA holds rows of every block) -- the <= 8 loads of a row are in flight
// together, serve all right-hand sides of the group, and the step costs one memory latency instead of the 8 rows per warp a single CTA needs
template <int NR>
__global__ void __cluster_dims__(LU_WIN_CLUSTER, 1, 1) __launch_bounds__(1024)
    lu_fwd_win_kernel(LuDev D, const int32_t* __restrict__ list, int use_up, int c_lo, int nrhs, cplx* __restrict__ x) {
  const int sn = list[blockIdx.x / LU_WIN_CLUSTER], rank = blockIdx.x % LU_WIN_CLUSTER;
  SnView S = sn_view(D, sn);
  if (c_lo >= S.s) return;  // the whole cluster leaves
  const int w = min(c_lo + LU_SOLVE_W, S.s) - c_lo;
  const int rhs0 = blockIdx.y * NR, nr = min(NR, nrhs - rhs0);
  __shared__ cplx xs[NR][LU_SOLVE_W];
  cplx* xg = x + (size_t)(S.first + c_lo) * nrhs + rhs0;  // x is stored position-major: x[pos * nrhs + rhs]
  for (int e = threadIdx.x; e < w * NR; e += 1024) {
    const int m = e / NR, q = e - m * NR;
    xs[q][m] = q < nr ? xg[(size_t)m * nrhs + q] : make_double2(0.0, 0.0);
  }
  __syncthreads();
  cluster_sync_all();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rho = rank + LU_WIN_CLUSTER * warp;
  if (rho >= w) return;
  const cplx* P = use_up ? S.up : S.lp;
  const cplx* dinv = D.dinv + D.dinv_off[sn] + (use_up ? NB * NB : 0);
  const int i = rho / NB, a = rho % NB;
  const cplx* col = P + c_lo + (size_t)(c_lo + rho) * S.ld;
  cplx v[LU_WBLK];
#pragma unroll
  for (int t = 0; t < LU_WBLK - 1; t++) v[t] = t < i ? col[lane + 32 * t] : make_double2(0.0, 0.0);
  {
    const cplx* Tinv = dinv + (size_t)(c_lo / NB + i) * 2 * NB * NB;
    v[LU_WBLK - 1] = NB * i + lane < w ? (use_up ? Tinv[lane + a * NB] : Tinv[a + lane * NB]) : make_double2(0.0, 0.0);
  }
#pragma unroll
  for (int q = 0; q < NR; q++) {
    if (q >= nr) break;
    double sr = 0.0, si = 0.0;
#pragma unroll
    for (int t = 0; t < LU_WBLK - 1; t++)
      if (t < i) {
        const cplx xv = xs[q][lane + 32 * t];
        sr += v[t].x * xv.x - v[t].y * xv.y;
        si += v[t].x * xv.y + v[t].y * xv.x;
      }
    if (NB * i + lane < w) {
      const cplx xv = xs[q][NB * i + lane];
      sr += v[LU_WBLK - 1].x * xv.x - v[LU_WBLK - 1].y * xv.y;
      si += v[LU_WBLK - 1].x * xv.y + v[LU_WBLK - 1].y * xv.x;
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) {
      sr += __shfl_xor_sync(0xffffffffu, sr, off);
      si += __shfl_xor_sync(0xffffffffu, si, off);
    }
    if (lane == 0) xg[(size_t)rho * nrhs + q] = make_double2(sr, si);
  }
}

// window step of the backward sweep: cluster of 8 CTAs, CTA j owns block row j of the window: thread = (row, one of 32 column parts)
template <int NR>
__global__ void __cluster_dims__(LU_WIN_CLUSTER, 1, 1) __launch_bounds__(1024)
    lu_bwd_win_kernel(LuDev D, const int32_t* __restrict__ list, int use_up, int c_lo, int nrhs, cplx* __restrict__ x) {
  const int sn = list[blockIdx.x / LU_WIN_CLUSTER], rank = blockIdx.x % LU_WIN_CLUSTER;
  SnView S = sn_view(D, sn);
  if (c_lo >= S.s) return;
  const int w = min(c_lo + LU_SOLVE_W, S.s) - c_lo;
  const int rhs0 = blockIdx.y * NR, nr = min(NR, nrhs - rhs0);
  __shared__ cplx vs[NR][LU_SOLVE_W], part[32][NB + 1];
  cplx* xg = x + (size_t)(S.first + c_lo) * nrhs + rhs0;
  for (int e = threadIdx.x; e < w * NR; e += 1024) {
    const int m = e / NR, q = e - m * NR;
    vs[q][m] = q < nr ? xg[(size_t)m * nrhs + q] : make_double2(0.0, 0.0);
  }
  __syncthreads();
  cluster_sync_all();
  const int rr = threadIdx.x & 31, p = threadIdx.x >> 5;
  const int rho = NB * rank + rr;
  cplx acc[NR];
#pragma unroll
  for (int q = 0; q < NR; q++) acc[q] = make_double2(0.0, 0.0);
  if (rho < w) win_bwd_row<NR>(use_up ? S.up : S.lp, S.ld, D.dinv + D.dinv_off[sn] + (use_up ? NB * NB : 0), use_up, c_lo, w, rho, p, 32, &vs[0][0], LU_SOLVE_W, acc);
#pragma unroll
  for (int q = 0; q < NR; q++) {  // the 32 column parts of a row are summed through shared memory, one right-hand side after the other
    part[p][rr] = acc[q];
    __syncthreads();
    if (p == 0 && rho < w && q < nr) {
      double tr = 0.0, ti = 0.0;
#pragma unroll 8
      for (int u = 0; u < 32; u++) {
        tr += part[u][rr].x;
        ti += part[u][rr].y;
      }
      xg[(size_t)rho * nrhs + q] = make_double2(tr, ti);
    }
    if (q + 1 < NR) __syncthreads();
  }
}

// x[rows below the window] -= P[rows, window] * y_window with PARTS column parts per row (64 rows per CTA): PARTS = 16 keeps a level with
// few supernodes (the top of the tree: 7 windows per supernode, one after the other) short -- 16 loads per thread, all in flight at once.
template <int NR, int PARTS>
__global__ void __launch_bounds__(64 * PARTS) lu_fwd_update2_kernel(LuDev D, const int32_t* __restrict__ list, int use_up, int c_lo, int nrhs, int64_t n,
                                                                   cplx* __restrict__ x) {
  SnView S = sn_view(D, list[blockIdx.y]);
  if (c_lo >= S.s) return;
  const int c_hi = min(c_lo + LU_SOLVE_W, S.s);
  const int nrows = S.ld - c_hi;
  if (blockIdx.x * 64 >= nrows) return;
  const int r = threadIdx.x & 63, part = threadIdx.x >> 6;
  const int i = blockIdx.x * 64 + r;
  const int rhs0 = blockIdx.z * NR, nr = min(NR, nrhs - rhs0);
  const cplx* P = (use_up ? S.up : S.lp) + c_hi;
  const int32_t* st = D.struct_idx + D.struct_ptr[list[blockIdx.y]];
  __shared__ cplx ys[NR][LU_SOLVE_W];
  __shared__ cplx red[NR][PARTS - 1][64];
  const int ncol = c_hi - c_lo;
  for (int j = threadIdx.x; j < ncol; j += 64 * PARTS)
#pragma unroll
    for (int q = 0; q < NR; q++)
      if (q < nr) ys[q][j] = x[(size_t)(S.first + c_lo + j) * nrhs + rhs0 + q];
  __syncthreads();
  cplx acc[NR];
#pragma unroll
  for (int q = 0; q < NR; q++) acc[q] = make_double2(0.0, 0.0);
  if (i < nrows) {
    constexpr int CP = LU_SOLVE_W / PARTS;
    const int j0 = part * CP, j1 = min(ncol, j0 + CP);
    const cplx* row = P + i + (size_t)c_lo * S.ld;
#pragma unroll 16
    for (int j = j0; j < j1; j++) {
      const cplx a = row[(size_t)j * S.ld];
#pragma unroll
      for (int q = 0; q < NR; q++) {
        acc[q].x += a.x * ys[q][j].x - a.y * ys[q][j].y;
        acc[q].y += a.x * ys[q][j].y + a.y * ys[q][j].x;
      }
    }
  }
  if (part)
#pragma unroll
    for (int q = 0; q < NR; q++) red[q][part - 1][r] = acc[q];
  __syncthreads();
  if (part == 0 && i < nrows) {
    const int g = c_hi + i;
#pragma unroll
    for (int q = 0; q < NR; q++)
      if (q < nr) {
        cplx t = acc[q];
#pragma unroll
        for (int u = 0; u < PARTS - 1; u++) {
          t.x += red[q][u][r].x;
          t.y += red[q][u][r].y;
        }
        if (g < S.s) {
          cplx* d = x + (size_t)(S.first + g) * nrhs + rhs0 + q;
          *d = csub(*d, t);
        } else
          catomic_sub(x + (size_t)st[g - S.s] * nrhs + rhs0 + q, t);
      }
  }
}

// CW adjacent columns per warp, lanes along the rows [r_begin, r_end) below row c_hi of the panel: sums[cw][q] = sum_i P[c_hi + i, c + cw] x_q[idx(i)].
// One gather of x serves CW columns (round 1: one gather per column -- as many bytes from L2 as the panel brings from HBM).
template <int NR, int CW>
__device__ __forceinline__ void col_dots(const cplx* __restrict__ colbase, int ld, int ncw, int r_begin, int r_end, int c_hi, int s, int first,
                                         const int32_t* __restrict__ st, const cplx* __restrict__ x, int nrhs, int nr, double (*sr)[NR], double (*si)[NR]) {
  const int lane = threadIdx.x & 31;
  for (int i = r_begin + lane; i < r_end; i += 32) {
    const int g = c_hi + i;
    const int64_t idx = g < s ? (int64_t)first + g : (int64_t)st[g - s];
    cplx a[CW];
#pragma unroll
    for (int cw = 0; cw < CW; cw++) a[cw] = cw < ncw ? colbase[i + (size_t)cw * ld] : make_double2(0.0, 0.0);
#pragma unroll
    for (int q = 0; q < NR; q++)
      if (q < nr) {
        const cplx v = x[(size_t)idx * nrhs + q];
#pragma unroll
        for (int cw = 0; cw < CW; cw++) {
          sr[cw][q] += a[cw].x * v.x - a[cw].y * v.y;
          si[cw][q] += a[cw].x * v.y + a[cw].y * v.x;
        }
      }
  }
#pragma unroll
  for (int cw = 0; cw < CW; cw++)
#pragma unroll
    for (int q = 0; q < NR; q++)
#pragma unroll
      for (int off = 16; off; off >>= 1) {
        sr[cw][q] += __shfl_xor_sync(0xffffffffu, sr[cw][q], off);
        si[cw][q] += __shfl_xor_sync(0xffffffffu, si[cw][q], off);
      }
}

// columns per warp of the backward kernels: 4, or 2 when 8 right-hand sides already fill the register file
#define LU_CW(NR) ((NR) >= 8 ? 2 : 4)

// backward update, CW columns per warp: grid.x = (column groups of 8 warps x CW) x (row chunks), grid.y supernode, grid.z groups of NR
template <int NR>
__global__ void __launch_bounds__(256, NR <= 2 ? 4 : 2) lu_bwd_update2_kernel(LuDev D, const int32_t* __restrict__ list, int use_up, int c_lo, int ncg, int row_chunk,
                                                             int nrhs, int64_t n, cplx* __restrict__ x) {
  SnView S = sn_view(D, list[blockIdx.y]);
  if (c_lo >= S.s) return;
  const int c_hi = min(c_lo + LU_SOLVE_W, S.s);
  const int nrows = S.ld - c_hi;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int CW = LU_CW(NR);
  const int c = c_lo + (blockIdx.x % ncg) * (8 * CW) + warp * CW;
  const int r_begin = (blockIdx.x / ncg) * row_chunk;
  if (c >= c_hi || r_begin >= nrows) return;
  const int r_end = min(nrows, r_begin + row_chunk);
  const int rhs0 = blockIdx.z * NR, nr = min(NR, nrhs - rhs0);
  double sr[CW][NR], si[CW][NR];
#pragma unroll
  for (int cw = 0; cw < CW; cw++)
#pragma unroll
    for (int q = 0; q < NR; q++) sr[cw][q] = si[cw][q] = 0.0;
  const int ncw = min(CW, c_hi - c);
  col_dots<NR, CW>((use_up ? S.up : S.lp) + c_hi + (size_t)c * S.ld, S.ld, ncw, r_begin, r_end, c_hi, S.s, S.first,
                  D.struct_idx + D.struct_ptr[list[blockIdx.y]], x + rhs0, nrhs, nr, sr, si);
  if (lane == 0)
#pragma unroll
    for (int cw = 0; cw < CW; cw++)
      if (cw < ncw)
#pragma unroll
        for (int q = 0; q < NR; q++)
          if (q < nr) catomic_sub(x + (size_t)(S.first + c + cw) * nrhs + rhs0 + q, make_double2(sr[cw][q], si[cw][q]));
}

// ---- fused sweeps of the deep levels: one CTA per supernode (pivot block within one window of LU_FW columns) ----------------------
// Thousands of small fronts per level: one launch per level and direction instead of a triangle and an update launch, the panel read once.
#define LU_FW 128
template <int NR>
__global__ void __launch_bounds__(256) lu_fwd_fused_kernel(LuDev D, const int32_t* __restrict__ list, int use_up, int nrhs, int64_t n, cplx* __restrict__ x) {
  const int sn = list[blockIdx.x];
  SnView S = sn_view(D, sn);
  const int rhs0 = blockIdx.y * NR, nr = min(NR, nrhs - rhs0);
  const cplx* P = use_up ? S.up : S.lp;
  __shared__ cplx xs[NR][LU_FW], ys[NR][LU_FW];
  cplx* xg = x + (size_t)S.first * nrhs + rhs0;  // x[pos * nrhs + rhs]
  for (int e = threadIdx.x; e < NR * S.s; e += 256) {
    const int m = e / NR, q = e - m * NR;
    xs[q][m] = q < nr ? xg[(size_t)m * nrhs + q] : make_double2(0.0, 0.0);
  }
  __syncthreads();
  win_fwd_matvec<NR>(P, S.ld, D.dinv + D.dinv_off[sn] + (use_up ? NB * NB : 0), use_up, 0, S.s, &xs[0][0], &ys[0][0], LU_FW, 8);
  __syncthreads();
  for (int e = threadIdx.x; e < NR * S.s; e += 256) {
    const int m = e / NR, q = e - m * NR;
    if (q < nr) xg[(size_t)m * nrhs + q] = ys[q][m];
  }
  // structure rows: one row per thread over all columns of the pivot block
  const int32_t* st = D.struct_idx + D.struct_ptr[sn];
  for (int i = threadIdx.x; i < S.r; i += 256) {
    const cplx* row = P + S.s + i;
    cplx acc[NR];
#pragma unroll
    for (int q = 0; q < NR; q++) acc[q] = make_double2(0.0, 0.0);
#pragma unroll 8
    for (int j = 0; j < S.s; j++) {
      const cplx a = row[(size_t)j * S.ld];
#pragma unroll
      for (int q = 0; q < NR; q++) {
        acc[q].x += a.x * ys[q][j].x - a.y * ys[q][j].y;
        acc[q].y += a.x * ys[q][j].y + a.y * ys[q][j].x;
      }
    }
    const int64_t g = st[i];
#pragma unroll
    for (int q = 0; q < NR; q++)
      if (q < nr) catomic_sub(x + (size_t)g * nrhs + rhs0 + q, acc[q]);
  }
}

template <int NR>
__global__ void __launch_bounds__(256, NR <= 2 ? 4 : 2) lu_bwd_fused_kernel(LuDev D, const int32_t* __restrict__ list, int use_up, int nrhs, int64_t n, cplx* __restrict__ x) {
  const int sn = list[blockIdx.x];
  SnView S = sn_view(D, sn);
  const int rhs0 = blockIdx.y * NR, nr = min(NR, nrhs - rhs0);
  const cplx* P = use_up ? S.up : S.lp;
  __shared__ cplx vs[NR][LU_FW];
  cplx* xg = x + (size_t)S.first * nrhs + rhs0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int32_t* st = D.struct_idx + D.struct_ptr[sn];
  // v_c = x_c - sum_{structure rows i} P[s + i, c] x[st[i]]: CW columns per warp
  constexpr int CW = LU_CW(NR);
  for (int c = warp * CW; c < S.s; c += 8 * CW) {
    double sr[CW][NR], si[CW][NR];
#pragma unroll
    for (int cw = 0; cw < CW; cw++)
#pragma unroll
      for (int q = 0; q < NR; q++) sr[cw][q] = si[cw][q] = 0.0;
    const int ncw = min(CW, S.s - c);
    col_dots<NR, CW>(P + S.s + (size_t)c * S.ld, S.ld, ncw, 0, S.r, S.s, S.s, S.first, st, x + rhs0, nrhs, nr, sr, si);
    if (lane < ncw) {
#pragma unroll
      for (int q = 0; q < NR; q++) {
        double tr = 0.0, ti = 0.0;
#pragma unroll
        for (int cw = 0; cw < CW; cw++)
          if (cw == lane) {
            tr = sr[cw][q];
            ti = si[cw][q];
          }
        const cplx xv = q < nr ? xg[(size_t)(c + lane) * nrhs + q] : make_double2(0.0, 0.0);
        vs[q][c + lane] = make_double2(xv.x - tr, xv.y - ti);
      }
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < S.s) {
    cplx acc[NR];
#pragma unroll
    for (int q = 0; q < NR; q++) acc[q] = make_double2(0.0, 0.0);
    win_bwd_row<NR>(P, S.ld, D.dinv + D.dinv_off[sn] + (use_up ? NB * NB : 0), use_up, 0, S.s, threadIdx.x, 0, 1, &vs[0][0], LU_FW, acc);
#pragma unroll
    for (int q = 0; q < NR; q++)
      if (q < nr) xg[(size_t)threadIdx.x * nrhs + q] = acc[q];
  }
}

// ---- backward kernels for 4 ... 8 right-hand sides: the rows of x a CTA needs are STAGED in shared memory ---------------------------
// With NR right-hand sides every lane of col_dots gathers NR values of x per row and CW columns -- for NR = 8, CW = 2 that is four times
// the bytes of the panel itself, from L2 (measured: the backward half of an 8-rhs sweep pair took twice the forward half).  Here a
// chunk of LU_XCH rows of x is gathered once per CTA (all its warps and columns use it), position-major -> [rhs][row] in shared memory.
#define LU_XCH 128
template <int NR>
__device__ __forceinline__ void stage_x_rows(cplx (*xs)[LU_XCH + 1], int base, int r_end, int c_hi, int s, int first, const int32_t* __restrict__ st,
                                             const cplx* __restrict__ x /* + rhs0 */, int nrhs, int nr) {
  for (int e = threadIdx.x; e < LU_XCH * NR; e += blockDim.x) {
    const int i = e / NR, q = e - i * NR, row = base + i;
    if (row < r_end && q < nr) {
      const int g = c_hi + row;
      const int64_t idx = g < s ? (int64_t)first + g : (int64_t)st[g - s];
      xs[q][i] = x[(size_t)idx * nrhs + q];
    }
  }
}
// partial sums of CW columns over the staged chunk: lanes along the rows
template <int NR, int CW>
__device__ __forceinline__ void col_dots_staged(const cplx* __restrict__ colbase /* row `base` of the first column */, int ld, int ncw, int nrows_chunk,
                                                const cplx (*xs)[LU_XCH + 1], int nr, double (*sr)[NR], double (*si)[NR]) {
  const int lane = threadIdx.x & 31;
  for (int i = lane; i < nrows_chunk; i += 32) {
    cplx a[CW];
#pragma unroll
    for (int cw = 0; cw < CW; cw++) a[cw] = cw < ncw ? colbase[i + (size_t)cw * ld] : make_double2(0.0, 0.0);
#pragma unroll
    for (int q = 0; q < NR; q++)
      if (q < nr) {
        const cplx v = xs[q][i];
#pragma unroll
        for (int cw = 0; cw < CW; cw++) {
          sr[cw][q] += a[cw].x * v.x - a[cw].y * v.y;
          si[cw][q] += a[cw].x * v.y + a[cw].y * v.x;
        }
      }
  }
}

template <int NR>
__global__ void __launch_bounds__(256, 2) lu_bwd_update3_kernel(LuDev D, const int32_t* __restrict__ list, int use_up, int c_lo, int ncg, int row_chunk,
                                                                int nrhs, int64_t n, cplx* __restrict__ x) {
  SnView S = sn_view(D, list[blockIdx.y]);
  if (c_lo >= S.s) return;
  const int c_hi = min(c_lo + LU_SOLVE_W, S.s);
  const int nrows = S.ld - c_hi;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int CW = LU_CW(NR);
  const int c = c_lo + (blockIdx.x % ncg) * (8 * CW) + warp * CW;
  const int r_begin = (blockIdx.x / ncg) * row_chunk;
  if (r_begin >= nrows || c_lo + (int)(blockIdx.x % ncg) * (8 * CW) >= c_hi) return;  // CTA-uniform
  const int r_end = min(nrows, r_begin + row_chunk);
  const int rhs0 = blockIdx.z * NR, nr = min(NR, nrhs - rhs0);
  __shared__ cplx xs[NR][LU_XCH + 1];
  const bool active = c < c_hi;
  const int ncw = active ? min(CW, c_hi - c) : 0;
  const int32_t* st = D.struct_idx + D.struct_ptr[list[blockIdx.y]];
  const cplx* col0 = (use_up ? S.up : S.lp) + c_hi + (size_t)(active ? c : c_lo) * S.ld;
  double sr[CW][NR], si[CW][NR];
#pragma unroll
  for (int cw = 0; cw < CW; cw++)
#pragma unroll
    for (int q = 0; q < NR; q++) sr[cw][q] = si[cw][q] = 0.0;
  for (int base = r_begin; base < r_end; base += LU_XCH) {
    __syncthreads();
    stage_x_rows<NR>(xs, base, r_end, c_hi, S.s, S.first, st, x + rhs0, nrhs, nr);
    __syncthreads();
    if (active) col_dots_staged<NR, CW>(col0 + base, S.ld, ncw, min(LU_XCH, r_end - base), xs, nr, sr, si);
  }
  if (!active) return;
#pragma unroll
  for (int cw = 0; cw < CW; cw++)
#pragma unroll
    for (int q = 0; q < NR; q++)
#pragma unroll
      for (int off = 16; off; off >>= 1) {
        sr[cw][q] += __shfl_xor_sync(0xffffffffu, sr[cw][q], off);
        si[cw][q] += __shfl_xor_sync(0xffffffffu, si[cw][q], off);
      }
  if (lane == 0)
#pragma unroll
    for (int cw = 0; cw < CW; cw++)
      if (cw < ncw)
#pragma unroll
        for (int q = 0; q < NR; q++)
          if (q < nr) catomic_sub(x + (size_t)(S.first + c + cw) * nrhs + rhs0 + q, make_double2(sr[cw][q], si[cw][q]));
}

template <int NR>
__global__ void __launch_bounds__(256, 2) lu_bwd_fused2_kernel(LuDev D, const int32_t* __restrict__ list, int use_up, int nrhs, int64_t n, cplx* __restrict__ x) {
  const int sn = list[blockIdx.x];
  SnView S = sn_view(D, sn);
  const int rhs0 = blockIdx.y * NR, nr = min(NR, nrhs - rhs0);
  const cplx* P = use_up ? S.up : S.lp;
  __shared__ cplx vs[NR][LU_FW];
  __shared__ cplx xs[NR][LU_XCH + 1];
  cplx* xg = x + (size_t)S.first * nrhs + rhs0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int32_t* st = D.struct_idx + D.struct_ptr[sn];
  constexpr int CW = LU_CW(NR);
  for (int e = threadIdx.x; e < NR * S.s; e += 256) {
    const int m = e / NR, q = e - m * NR;
    vs[q][m] = q < nr ? xg[(size_t)m * nrhs + q] : make_double2(0.0, 0.0);
  }
  // v_c = x_c - sum_{structure rows i} P[s + i, c] x[st[i]], chunk of rows by chunk; a warp owns its columns of vs
  for (int base = 0; base < S.r; base += LU_XCH) {
    __syncthreads();
    stage_x_rows<NR>(xs, base, S.r, S.s, S.s, S.first, st, x + rhs0, nrhs, nr);
    __syncthreads();
    const int nch = min(LU_XCH, S.r - base);
    for (int c = warp * CW; c < S.s; c += 8 * CW) {
      double sr[CW][NR], si[CW][NR];
#pragma unroll
      for (int cw = 0; cw < CW; cw++)
#pragma unroll
        for (int q = 0; q < NR; q++) sr[cw][q] = si[cw][q] = 0.0;
      const int ncw = min(CW, S.s - c);
      col_dots_staged<NR, CW>(P + S.s + base + (size_t)c * S.ld, S.ld, ncw, nch, xs, nr, sr, si);
#pragma unroll
      for (int cw = 0; cw < CW; cw++)
#pragma unroll
        for (int q = 0; q < NR; q++)
#pragma unroll
          for (int off = 16; off; off >>= 1) {
            sr[cw][q] += __shfl_xor_sync(0xffffffffu, sr[cw][q], off);
            si[cw][q] += __shfl_xor_sync(0xffffffffu, si[cw][q], off);
          }
      if (lane < ncw) {
#pragma unroll
        for (int q = 0; q < NR; q++) {
          double tr = 0.0, ti = 0.0;
#pragma unroll
          for (int cw = 0; cw < CW; cw++)
            if (cw == lane) {
              tr = sr[cw][q];
              ti = si[cw][q];
            }
          vs[q][c + lane].x -= tr;
          vs[q][c + lane].y -= ti;
        }
      }
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < S.s) {
    cplx acc[NR];
#pragma unroll
    for (int q = 0; q < NR; q++) acc[q] = make_double2(0.0, 0.0);
    win_bwd_row<NR>(P, S.ld, D.dinv + D.dinv_off[sn] + (use_up ? NB * NB : 0), use_up, 0, S.s, threadIdx.x, 0, 1, &vs[0][0], LU_FW, acc);
#pragma unroll
    for (int q = 0; q < NR; q++)
      if (q < nr) xg[(size_t)threadIdx.x * nrhs + q] = acc[q];
  }
}

// ---- vector utilities -------------------------------------------------------------------------------------
// y[pos] = d[perm[pos]] * b[perm[pos]] (optionally conjugated)      /      x[perm[pos]] = d[perm[pos]] * y[pos]
__global__ void lu_permute_in_kernel(const cplx* __restrict__ b, const int32_t* __restrict__ perm, const double* __restrict__ d,
                                     int64_t n, int nrhs, int conj, cplx* __restrict__ y) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  int32_t o = perm[p];
  double s = d[o];
  for (int r = 0; r < nrhs; r++) {
    cplx v = b[(size_t)r * n + o];
    y[(size_t)p * nrhs + r] = make_double2(v.x * s, conj ? -v.y * s : v.y * s);
  }
}
__global__ void lu_permute_out_kernel(const cplx* __restrict__ y, const int32_t* __restrict__ perm, const double* __restrict__ d,
                                      int64_t n, int nrhs, int conj, int accumulate, cplx* __restrict__ x) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  int32_t o = perm[p];
  double s = d[o];
  for (int r = 0; r < nrhs; r++) {
    cplx v = y[(size_t)p * nrhs + r];
    v = make_double2(v.x * s, conj ? -v.y * s : v.y * s);
    cplx* dst = x + (size_t)r * n + o;
    if (accumulate) {
      v.x += dst->x;
      v.y += dst->y;
    }
    *dst = v;
  }
}
// r = b - r
__global__ void lu_residual_kernel(const cplx* __restrict__ b, int64_t total, cplx* __restrict__ r) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  r[i] = make_double2(b[i].x - r[i].x, b[i].y - r[i].y);
}

// ============================================================================================================
static LuDev make_dev(LuSolver& S) {
  LuDev D;
  D.sn_first = S.d_sn_first.p;
  D.struct_ptr = S.d_struct_ptr.p;
  D.struct_idx = S.d_struct_idx.p;
  D.rel_idx = S.d_rel_idx.p;
  D.sn_parent = S.d_sn_parent.p;
  D.lp_off = S.d_lp_off.p;
  D.up_off = S.d_up_off.p;
  D.upd_off = S.d_upd_off.p;
  D.dinv_off = S.d_dinv_off.p;
  D.fac = S.d_fac.p;
  D.dinv = S.d_dinv.p;
  return D;
}

void wae_lu_setup_device(wae_ctx* h, LuSolver& S) {
  LuSymbolic& Y = S.sym;
  cudaStream_t st = h->stream;
  S.d_perm.upload(Y.perm, st);
  S.d_iperm.upload(Y.iperm, st);
  S.d_sn_first.upload(Y.sn_first, st);
  S.d_sn_parent.upload(Y.sn_parent, st);
  S.d_struct_ptr.upload(Y.struct_ptr, st);
  S.d_struct_idx.upload(Y.struct_idx.empty() ? std::vector<int32_t>(1, 0) : Y.struct_idx, st);
  S.d_rel_idx.upload(Y.rel_idx.empty() ? std::vector<int32_t>(1, 0) : Y.rel_idx, st);
  S.d_diagpos.upload(Y.diagpos, st);
  S.d_lp_off.upload(Y.lp_off, st);
  S.d_up_off.upload(Y.up_off, st);
  S.d_upd_off.upload(Y.upd_off, st);
  S.d_dinv_off.upload(Y.dinv_off, st);
  S.d_dinv.alloc((size_t)std::max<int64_t>(Y.dinv_size, 1));
  S.d_amap.upload(Y.amap, st);
  // per-depth lists, largest pivot block first (so that "supernodes with more than k blocks" is a prefix)
  S.d_level.resize(Y.levels.size());
  for (size_t d = 0; d < Y.levels.size(); d++) {
    std::vector<int32_t>& L = Y.levels[d];
    std::stable_sort(L.begin(), L.end(), [&](int32_t a, int32_t b) {
      return Y.sn_first[a + 1] - Y.sn_first[a] > Y.sn_first[b + 1] - Y.sn_first[b];
    });
    S.d_level[d].upload(L, st);
  }
  for (int g = 0; g < 2; g++) {  // two front groups per depth (both stay sorted by pivot size)
    S.level_half[g].clear();
    S.level_half_ptr[g].assign(Y.levels.size() + 1, 0);
    for (size_t d = 0; d < Y.levels.size(); d++) {
      for (size_t i = g; i < Y.levels[d].size(); i += 2) S.level_half[g].push_back(Y.levels[d][i]);
      S.level_half_ptr[g][d + 1] = (int64_t)S.level_half[g].size();
    }
    S.d_level_half[g].upload(S.level_half[g].empty() ? std::vector<int32_t>(1, 0) : S.level_half[g], st);
  }
  S.d_xa_tile_ptr.resize(Y.levels.size());
  S.xa_tiles.assign(Y.levels.size(), 0);
  for (size_t d = 0; d < Y.levels.size(); d++) {
    const std::vector<int32_t>& C = Y.levels[d];
    std::vector<int32_t> tile_ptr(C.size() + 1, 0);
    for (size_t i = 0; i < C.size(); i++) {
      int64_t r = Y.struct_ptr[C[i] + 1] - Y.struct_ptr[C[i]];
      int64_t nt = (r + 31) / 32;
      tile_ptr[i + 1] = tile_ptr[i] + (int32_t)(nt * nt);
    }
    S.xa_tiles[d] = tile_ptr.back();
    S.d_xa_tile_ptr[d].upload(tile_ptr, st);
  }
  // rounds of the extend-add: sibling rank of every supernode among the children of its parent
  {
    std::vector<int32_t> nchild_seen(Y.nsn, 0), rank(Y.nsn, 0);
    for (int k = 0; k < Y.nsn; k++)
      if (Y.sn_parent[k] >= 0) rank[k] = nchild_seen[Y.sn_parent[k]]++;
    S.xa_rounds.assign(Y.levels.size(), {});
    S.d_xa_round_children.resize(Y.levels.size());
    S.d_xa_round_tile_ptr.resize(Y.levels.size());
    for (size_t d = 0; d < Y.levels.size(); d++) {
      const std::vector<int32_t>& C = Y.levels[d];
      int nround = 0;
      for (int32_t k : C)
        if (Y.sn_parent[k] >= 0) nround = std::max(nround, rank[k] + 1);
      std::vector<int32_t> order, tptr;
      for (int r = 0; r < nround; r++) {
        LuSolver::XaRound R;
        R.off = (int32_t)order.size();
        const size_t t0 = tptr.size();
        tptr.push_back(0);
        for (int32_t k : C) {
          if (Y.sn_parent[k] < 0 || rank[k] != r) continue;
          const int64_t rr = Y.struct_ptr[k + 1] - Y.struct_ptr[k], nt = (rr + 31) / 32;
          order.push_back(k);
          tptr.push_back(tptr.back() + (int32_t)(nt * nt));
        }
        R.count = (int32_t)order.size() - R.off;
        R.tiles = tptr.back();
        (void)t0;
        S.xa_rounds[d].push_back(R);
      }
      S.d_xa_round_children[d].upload(order.empty() ? std::vector<int32_t>(1, 0) : order, st);
      S.d_xa_round_tile_ptr[d].upload(tptr.empty() ? std::vector<int32_t>(1, 0) : tptr, st);
    }
  }
  // work items of the window inverses: every block j that is not the last one of its window, grouped by the blocks below it
  {
    std::vector<std::vector<int32_t>> cls(LU_WBLK);
    for (int d = 0; d < (int)Y.levels.size(); d++)  // top of the tree (the long items) first
      for (int32_t k : Y.levels[d]) {
        const int nblk = (Y.sn_first[k + 1] - Y.sn_first[k] + NB - 1) / NB;
        for (int j = 0; j < nblk; j++) {
          const int wend = std::min((j / LU_WBLK + 1) * LU_WBLK, nblk);
          const int below = wend - 1 - j;
          if (below > 0) {
            cls[below].push_back(k);
            cls[below].push_back(j);
          }
        }
      }
    std::vector<int32_t> items;
    S.wininv_ptr[0] = S.wininv_ptr[1] = 0;
    for (int c = 1; c < LU_WBLK; c++) {
      items.insert(items.end(), cls[c].begin(), cls[c].end());
      S.wininv_ptr[c + 1] = (int)(items.size() / 2);
    }
    if (items.empty()) items.assign(2, 0);
    S.d_wininv_items.upload(items, st);
  }
  // column index of every nonzero of A
  {
    Pattern& U = h->pat(h->fam(S.fam).pattern);
    std::vector<int32_t> col(U.nnz);
    for (int64_t j = 0; j < U.dim; j++)
      for (int64_t k = U.colptr[j]; k < U.colptr[j + 1]; k++) col[k] = (int32_t)j;
    S.d_colidx_nz.upload(col, st);
    S.d_Aval.alloc((size_t)U.nnz);
  }
  int64_t e0 = 0, e1 = 0;
  for (size_t d = 0; d < Y.level_upd_size.size(); d++) (d & 1 ? e1 : e0) = std::max<int64_t>(d & 1 ? e1 : e0, Y.level_upd_size[d]);
  S.d_upd[0].alloc((size_t)std::max<int64_t>(e0, 1));
  S.d_upd[1].alloc((size_t)std::max<int64_t>(e1, 1));
  S.d_fac.alloc((size_t)std::max<int64_t>(Y.fac_size, 1));
  S.d_scale.alloc(Y.n);
  S.d_flag.alloc(2);
  CUDA_CHECK(cudaStreamSynchronize(st));
  std::vector<int64_t>().swap(Y.amap);  // large and only needed on the device
}

void wae_lu_factor_device(wae_ctx* h, LuSolver& S, const cplx* d_Aval, const cplx* d_full, int sym) {
  LuSymbolic& Y = S.sym;
  cudaStream_t st = h->stream;
  Family& F = h->fam(S.fam);
  Pattern& U = h->pat(F.pattern);
  wae_family_ensure_csr(h, F);
  CUDA_CHECK(cudaMemcpyAsync(S.d_Aval.p, d_full ? d_full : d_Aval, (size_t)U.nnz * sizeof(cplx), cudaMemcpyDeviceToDevice, st));
  S.sym_mode = sym != 0;
  S.aval_csr_valid = false;
  LuDev D = make_dev(S);
  CUDA_CHECK(cudaMemsetAsync(S.d_fac.p, 0, (size_t)Y.fac_size * sizeof(cplx), st));
  CUDA_CHECK(cudaMemsetAsync(S.d_flag.p, 0, 2 * sizeof(int32_t), st));
  lu_scale_kernel<<<(unsigned)((Y.n + 255) / 256), 256, 0, st>>>(d_Aval, S.d_diagpos.p, Y.n, S.d_scale.p);
  lu_scatter_kernel<<<(unsigned)((U.nnz + 255) / 256), 256, 0, st>>>(d_Aval, S.d_amap.p, U.d_rowval.p, S.d_colidx_nz.p, S.d_scale.p, U.nnz, S.d_fac.p);
  h->launches += 2;
  const int maxd = (int)Y.levels.size() - 1;
  int nbo_blocks = 4;  // outer block = 4 * NB = 128 columns
  if (const char* env = getenv("WAE_LU_NBO")) nbo_blocks = std::max(1, atoi(env) / NB);
  const bool gemm_ring = !(getenv("WAE_LU_GEMM") && atoi(getenv("WAE_LU_GEMM")) == 1);  // default: cp.async ring (2); 1 = single-buffered tile
  if (gemm_ring) {  // 99 KB of dynamic shared memory per CTA, two CTAs per SM
    CUDA_CHECK(cudaFuncSetAttribute(lu_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GP_SMEM));
    if (cudaFuncSetAttribute(lu_gemm_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess)
      cudaGetLastError();  // a refused hint must not surface as a launch error
  }
  auto gemm = [&](cudaStream_t s, dim3 g, const int32_t* lst, int mode, int k0, int kw, int c0, int cap, cplx* upd_, int flag) {
    if (gemm_ring)
      lu_gemm_kernel<1><<<g, 256, GP_SMEM, s>>>(D, lst, mode, k0, kw, c0, cap, upd_, flag);
    else
      lu_gemm_kernel<0><<<g, 256, 0, s>>>(D, lst, mode, k0, kw, c0, cap, upd_, flag);
  };
  const int upd_flag = (sym ? 1 : 0) | ((getenv("WAE_LU_SKIP_UPPER") && !atoi(getenv("WAE_LU_SKIP_UPPER"))) ? 0 : 2);  // pivot-block updates only; default: skip
  // symmetric elimination: U^T panel = L panel * D inside the panel kernel (WAE_LU_SYM_PANEL=0: round-1 path, copy + second solve)
  const bool sym_panel = !(getenv("WAE_LU_SYM_PANEL") && !atoi(getenv("WAE_LU_SYM_PANEL")));
  // round-2 kernels, each with its switch for A/B runs: one-warp diagonal blocks, panel rows times the explicit inverse, Schur complement
  // written instead of accumulated
  // (measured on config 2: the one-warp diagonal block wins where a launch holds >= 1024 blocks -- 1.07 vs 1.70 ms on the 8000-front levels --
  // and loses where a launch is one block on one warp, 58 vs 34 us; the inverse-times-row panel wins only on launches of one or two fronts)
  const int diag_warp_min = getenv("WAE_LU_DIAG") ? atoi(getenv("WAE_LU_DIAG")) : 1024;   // 0: never
  const int panel_inv_max = (!sym || sym_panel) ? (getenv("WAE_LU_PANEL") ? atoi(getenv("WAE_LU_PANEL")) : (1 << 30)) : 0;  // 0: never
  const bool late_xadd = !(getenv("WAE_LU_LATE_XADD") && !atoi(getenv("WAE_LU_LATE_XADD")));
  const bool two_groups = !(getenv("WAE_LU_GROUPS") && atoi(getenv("WAE_LU_GROUPS")) == 1);
  // extend-add in rounds of sibling rank without atomics (WAE_LU_XADD_ROUNDS=0: one launch with fp64 atomics)
  const bool xadd_rounds = !(getenv("WAE_LU_XADD_ROUNDS") && !atoi(getenv("WAE_LU_XADD_ROUNDS")));
  const int prof_depth = getenv("WAE_LU_PROFILE_DEPTH") ? atoi(getenv("WAE_LU_PROFILE_DEPTH")) : -1;
  if (two_groups && !h->aux_stream[0]) {
    // group 0 on a high-priority stream, group 1 on a low-priority one: CTAs are dispatched kernel by kernel, so with equal priorities the
    // small kernels of one group would queue behind the whole GEMM grid of the other; with priorities group 0 never waits and group 1's
    // GEMMs fill the gaps its serial steps leave (WAE_LU_GROUPS=3: equal priorities)
    int prio_lo = 0, prio_hi = 0;
    CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    const bool equal_prio = getenv("WAE_LU_GROUPS") && atoi(getenv("WAE_LU_GROUPS")) == 3;
    for (int g = 0; g < 2; g++) {
      CUDA_CHECK(cudaStreamCreateWithPriority(&h->aux_stream[g], cudaStreamNonBlocking, equal_prio ? prio_lo : (g == 0 ? prio_hi : prio_lo)));
      CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_join[g], cudaEventDisableTiming));
    }
    CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  }
  // WAE_LU_TRACE=1 (diagnostic): every launch class is timed with its own pair of events (serialising the stream) and summed per tree
  // depth; the table goes to stderr and the totals to wae_last_ms("lu_trace_<class>")
  const bool trace = getenv("WAE_LU_TRACE") != nullptr;
  enum { T_MEMSET, T_XADD, T_DIAG, T_COPY, T_PANEL, T_GIN, T_GOUT, T_SCHUR, T_N };
  static const char* tname[T_N] = {"memset", "extend_add", "diag", "sym_copy", "panel", "gemm_inner", "gemm_outer", "gemm_schur"};
  std::vector<std::array<double, T_N>> tms(maxd + 1);
  for (auto& a : tms) a.fill(0.0);
  cudaEvent_t te0 = nullptr, te1 = nullptr;
  if (trace) {
    CUDA_CHECK(cudaEventCreate(&te0));
    CUDA_CHECK(cudaEventCreate(&te1));
  }
  int cur_d = maxd;
  auto timed = [&](int cls, auto&& launch) {
    if (!trace) {
      launch();
      return;
    }
    cudaEventRecord(te0, st);
    launch();
    cudaEventRecord(te1, st);
    cudaEventSynchronize(te1);
    float ms = 0;
    cudaEventElapsedTime(&ms, te0, te1);
    tms[cur_d][cls] += ms;
  };
  for (int d = maxd; d >= 0; d--) {
    cur_d = d;
    const std::vector<int32_t>& L = Y.levels[d];
    const int nl = (int)L.size();
    cplx* upd = S.d_upd[d & 1].p;
    // late_xadd (default): the Schur complement is WRITTEN (C = -A B^T, no memset and no read of C) and the children's contributions to
    // the update matrix are added after it; only their contributions to the pivot panels have to be there before the factorisation
    if (Y.level_upd_size[d] && !late_xadd) timed(T_MEMSET, [&] { CUDA_CHECK(cudaMemsetAsync(upd, 0, (size_t)Y.level_upd_size[d] * sizeof(cplx), st)); });
    const bool has_children = d < maxd && S.xa_tiles[d + 1] > 0;
    auto xadd = [&](int part) {
      if (xadd_rounds) {  // one launch per sibling rank, plain read-modify-writes
        int tp = 0;  // round r's tile prefix starts after the (count + 1) entries of every earlier round
        for (const LuSolver::XaRound& R : S.xa_rounds[d + 1]) {
          if (R.tiles > 0) {
            timed(T_XADD, [&] {
              lu_extend_add_kernel<<<R.tiles, dim3(32, 8), 0, st>>>(D, S.d_xa_round_children[d + 1].p + R.off, S.d_xa_round_tile_ptr[d + 1].p + tp, R.count,
                                                                    S.d_upd[(d + 1) & 1].p, upd, sym, part, 0);
            });
            h->launches++;
          }
          tp += R.count + 1;
        }
        return;
      }
      timed(T_XADD, [&] {
        lu_extend_add_kernel<<<S.xa_tiles[d + 1], dim3(32, 8), 0, st>>>(D, S.d_level[d + 1].p, S.d_xa_tile_ptr[d + 1].p,
                                                                          (int)Y.levels[d + 1].size(), S.d_upd[(d + 1) & 1].p, upd, sym, part, 1);
      });
      h->launches++;
    };
    if (has_children) xadd(late_xadd ? 1 : 0);
    // blocked partial factorisation + Schur complement of a group of fronts of this depth (host list Lh / device list Ld, size-sorted) on stream s_
    auto front_group = [&](const int32_t* Lh, int nl_, const int32_t* Ld, cudaStream_t s_) {
    int max_s = 0, max_ld = 0, max_r = 0;
    for (int i_ = 0; i_ < nl_; i_++) {
      const int32_t k = Lh[i_];
      int s = Y.sn_first[k + 1] - Y.sn_first[k], r = (int)(Y.struct_ptr[k + 1] - Y.struct_ptr[k]);
      max_s = std::max(max_s, s);
      max_ld = std::max(max_ld, s + r);
      max_r = std::max(max_r, r);
    }
    const int nsteps = (max_s + NB - 1) / NB;
    for (int k = 0; k < nsteps; k++) {
      // supernodes with more than k blocks form a prefix of the (size-sorted) list
      int cnt = 0;
      while (cnt < nl_ && Y.sn_first[Lh[cnt] + 1] - Y.sn_first[Lh[cnt]] > k * NB) cnt++;
      if (!cnt) break;
      for (int z0 = 0; z0 < cnt; z0 += 32768) {
        int zc = std::min(32768, cnt - z0);
        const int32_t* lst = Ld + z0;
        timed(T_DIAG, [&] {
          if (diag_warp_min > 0 && zc >= diag_warp_min)
            lu_diag_warp_kernel<<<(zc + LU_DIAG_WARPS - 1) / LU_DIAG_WARPS, 32 * LU_DIAG_WARPS, 0, s_>>>(D, lst, zc, k, S.pivot_eps, S.d_flag.p);
          else
            lu_diag_kernel<<<zc, dim3(NB, NB), 0, s_>>>(D, lst, k, S.pivot_eps, S.d_flag.p);
        });
        int rows = max_ld - k * NB - 1;  // upper bound of ld - (c0 + nb) over the batch (nb >= 1)
        if (rows > 0) {
          if (sym && !sym_panel) {
            timed(T_COPY, [&] { lu_sym_copy_kernel<<<dim3((rows + 127) / 128, 1, zc), 128, 0, s_>>>(D, lst, k); });
            h->launches++;
          }
          const int py = (sym && sym_panel) ? 1 : 2;
          timed(T_PANEL, [&] {
            if (zc <= panel_inv_max)
              lu_panel_inv_kernel<<<dim3((rows + 127) / 128, py, zc), 128, 0, s_>>>(D, lst, k, py == 1);
            else
              lu_panel_kernel<<<dim3((rows + 127) / 128, py, zc), 128, 0, s_>>>(D, lst, k, py == 1);
          });
          // inner update: columns of the current outer block only
          const int c0 = (k + 1) * NB;
          const int oend = (k / nbo_blocks + 1) * nbo_blocks * NB;  // end column of the outer block
          int tn = std::min(max_s, oend) - c0;
          if (tn > 0) {
            dim3 g((max_ld - c0 + GT - 1) / GT, (tn + GT - 1) / GT, zc);
            timed(T_GIN, [&] {
              gemm(s_, g, lst, 0, k * NB, NB, c0, oend, nullptr, upd_flag);
              if (!sym) gemm(s_, g, lst, 1, k * NB, NB, c0, oend, nullptr, upd_flag);
            });
            h->launches += sym ? 1 : 2;
          }
          // outer update once the outer block is complete
          if (c0 == oend && max_s > oend) {
            const int o0 = oend - nbo_blocks * NB;
            dim3 g((max_ld - oend + GT - 1) / GT, (max_s - oend + GT - 1) / GT, zc);
            timed(T_GOUT, [&] {
              gemm(s_, g, lst, 0, o0, oend - o0, oend, 1 << 30, nullptr, upd_flag);
              if (!sym) gemm(s_, g, lst, 1, o0, oend - o0, oend, 1 << 30, nullptr, upd_flag);
            });
            h->launches += sym ? 1 : 2;
          }
          h->launches++;
        }
        h->launches++;
      }
    }
    if (max_r > 0) {
      for (int z0 = 0; z0 < nl_; z0 += 32768) {
        int zc = std::min(32768, nl_ - z0);
        dim3 g((max_r + GT - 1) / GT, (max_r + GT - 1) / GT, zc);
        // WAE_LU_PROFILE_DEPTH=d (diagnostic): the profiler range covers the Schur-complement launches of depth d (ncu --profile-from-start off)
        const bool prof = prof_depth == d;
        if (prof) {
          cudaStreamSynchronize(s_);
          cudaProfilerStart();
        }
        timed(T_SCHUR, [&] { gemm(s_, g, Ld + z0, 2, 0, 0, 0, 0, upd, sym | (late_xadd ? 4 : 0)); });
        if (prof) {
          cudaStreamSynchronize(s_);
          cudaProfilerStop();
        }
        h->launches++;
      }
    }
    };
    // two groups of fronts on two auxiliary streams (levels with 2 ... 1023 fronts of more than one block step): the serial chain of small
    // kernels of one group (diagonal block -> panel -> K = 32 update, 53 times on the top levels) runs under the GEMMs of the other
    int lvl_max_s = 0, lvl_max_r = 0;
    for (int32_t k : L) {
      lvl_max_s = std::max(lvl_max_s, (int)(Y.sn_first[k + 1] - Y.sn_first[k]));
      lvl_max_r = std::max(lvl_max_r, (int)(Y.struct_ptr[k + 1] - Y.struct_ptr[k]));
    }
    if (two_groups && !trace && nl >= 2 && nl < 1024 && lvl_max_s > NB) {
      CUDA_CHECK(cudaEventRecord(h->ev_fork, st));
      for (int g = 0; g < 2; g++) {
        CUDA_CHECK(cudaStreamWaitEvent(h->aux_stream[g], h->ev_fork, 0));
        const int64_t o = S.level_half_ptr[g][d];
        front_group(S.level_half[g].data() + o, (int)(S.level_half_ptr[g][d + 1] - o), S.d_level_half[g].p + o, h->aux_stream[g]);
        CUDA_CHECK(cudaEventRecord(h->ev_join[g], h->aux_stream[g]));
        CUDA_CHECK(cudaStreamWaitEvent(st, h->ev_join[g], 0));
      }
    } else {
      front_group(L.data(), nl, S.d_level[d].p, st);
    }
    if (lvl_max_r > 0) {
      if (late_xadd && has_children) xadd(2);
    }
  }
  // window inverses for the solves (WAE_LU_WININV=0: round-1 solve kernels, block-by-block substitution inside a window)
  S.wininv = !(getenv("WAE_LU_WININV") && !atoi(getenv("WAE_LU_WININV")));
  if (S.wininv) {
    CUDA_CHECK(cudaFuncSetAttribute(lu_wininv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (LU_WBLK + 2) * LU_TILE_BYTES));
    cudaEvent_t w0 = nullptr, w1 = nullptr;
    if (trace) {
      CUDA_CHECK(cudaEventCreate(&w0));
      CUDA_CHECK(cudaEventCreate(&w1));
      cudaEventRecord(w0, st);
    }
    const int dual = sym && sym_panel;  // the U^T panel is exactly L D there
    for (int c = LU_WBLK - 1; c >= 1; c--) {
      const int cnt = S.wininv_ptr[c + 1] - S.wininv_ptr[c];
      for (int z0 = 0; z0 < cnt; z0 += 1 << 20) {
        const int zc = std::min(1 << 20, cnt - z0);
        lu_wininv_kernel<<<dim3(zc, dual ? 1 : 2), dim3(NB, NB), (c + 3) * LU_TILE_BYTES, st>>>(D, S.d_wininv_items.p + 2 * (size_t)(S.wininv_ptr[c] + z0), c, dual);
        h->launches++;
      }
    }
    if (trace) {
      cudaEventRecord(w1, st);
      cudaEventSynchronize(w1);
      float ms = 0;
      cudaEventElapsedTime(&ms, w0, w1);
      fprintf(stderr, "[wae lu trace] window inverses %.3f ms\n", ms);
      h->last_ms["lu_trace_wininv"] = ms;
      cudaEventDestroy(w0);
      cudaEventDestroy(w1);
    }
  }
  if (trace) {
    std::array<double, T_N> tot;
    tot.fill(0.0);
    fprintf(stderr, "[wae lu trace] depth  fronts  max_s  max_r ");
    for (int c = 0; c < T_N; c++) fprintf(stderr, " %10s", tname[c]);
    fprintf(stderr, "   (ms)\n");
    for (int d = maxd; d >= 0; d--) {
      int max_s = 0, max_r = 0;
      for (int32_t k : Y.levels[d]) {
        max_s = std::max(max_s, (int)(Y.sn_first[k + 1] - Y.sn_first[k]));
        max_r = std::max(max_r, (int)(Y.struct_ptr[k + 1] - Y.struct_ptr[k]));
      }
      fprintf(stderr, "[wae lu trace] %5d %7d %6d %6d ", d, (int)Y.levels[d].size(), max_s, max_r);
      for (int c = 0; c < T_N; c++) {
        fprintf(stderr, " %10.3f", tms[d][c]);
        tot[c] += tms[d][c];
      }
      // useful work of the depth (8 real flops per complex multiply-add; symmetric mode: lower halves only) and the rate it ran at
      double f_schur = 0.0, f_piv = 0.0;
      for (int32_t k : Y.levels[d]) {
        const double s = Y.sn_first[k + 1] - Y.sn_first[k], r = (double)(Y.struct_ptr[k + 1] - Y.struct_ptr[k]);
        f_schur += 8.0 * r * r * s * (sym ? 0.5 : 1.0);
        f_piv += 8.0 * (s * s * s / 3.0 + r * s * s * (sym ? 0.5 : 1.0));  // LU of the pivot block + the panel updates
      }
      const double t_piv = tms[d][T_DIAG] + tms[d][T_COPY] + tms[d][T_PANEL] + tms[d][T_GIN] + tms[d][T_GOUT];
      fprintf(stderr, "   schur %6.1f GF %5.1f TF/s   pivot block %6.1f GF %5.1f TF/s\n", f_schur / 1e9, tms[d][T_SCHUR] > 0 ? f_schur / tms[d][T_SCHUR] / 1e9 : 0.0,
              f_piv / 1e9, t_piv > 0 ? f_piv / t_piv / 1e9 : 0.0);
    }
    fprintf(stderr, "[wae lu trace] total                      ");
    for (int c = 0; c < T_N; c++) {
      fprintf(stderr, " %10.3f", tot[c]);
      h->last_ms[std::string("lu_trace_") + tname[c]] = tot[c];
    }
    fprintf(stderr, "\n");
    cudaEventDestroy(te0);
    cudaEventDestroy(te1);
  }
  CUDA_CHECK(cudaGetLastError());
  int32_t flag[2] = {0, 0};
  CUDA_CHECK(cudaMemcpyAsync(flag, S.d_flag.p, sizeof(flag), cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  S.factored = true;
  if (flag[0] & 2) {
    S.factored = false;
    WAE_THROW(WAE_E_SINGULAR, "non-finite pivot in the numeric factorisation");
  }
  h->last_ms["static_pivots"] = flag[1];
  h->last_ms["zero_pivots"] = (flag[0] & 4) ? 1.0 : 0.0;
  // UMFPACK (the reference's lu / factorize) raises SingularException on an exactly singular matrix; here an exactly zero pivot is the
  // case the static perturbation cannot repair (the solves would return garbage).  perturb's lu(...; check=false) of the deliberately
  // singular L(0,0) (perturbation.jl:329) goes through wae_lu_factor_ex(check = 0).
  if ((flag[0] & 4) && S.check_singular) {
    S.factored = false;
    WAE_THROW(WAE_E_SINGULAR, "exactly zero pivot in the numeric factorisation (%d perturbed pivots): the matrix is singular for an LU without row exchanges", flag[1]);
  }
}

template <int NR>
static void lu_sweeps_nr(wae_ctx* h, LuSolver& S, int trans_t, int nrhs, cplx* y) {
  LuSymbolic& Y = S.sym;
  cudaStream_t st = h->stream;
  LuDev D = make_dev(S);
  const int maxd = (int)Y.levels.size() - 1;
  const int fwd_up = trans_t ? 1 : 0;  // A: forward with Lp, backward with Up ; A^T: forward with Up, backward with Lp
  const int tri_pf = (getenv("WAE_LU_SOLVE_PF") && atoi(getenv("WAE_LU_SOLVE_PF"))) ? 2 : 0;  // bit 1 of the tri kernels' use_up: L2 prefetch hint
  const int zr = (nrhs + NR - 1) / NR;
  const int W = LU_SOLVE_W;
  constexpr int WNR = NR > 4 ? 4 : NR;  // right-hand sides per cluster of the window kernels (shared memory: 2 x WNR x 256 complex)
  // round-2 kernels (each can be switched off for A/B runs): fused deep levels (needs the window inverses), 16-part forward update and
  // 4-columns-per-warp backward update
  const bool fused_ok = S.wininv && !(getenv("WAE_LU_SOLVE_FUSED") && !atoi(getenv("WAE_LU_SOLVE_FUSED")));
  const bool upd2_ok = !(getenv("WAE_LU_SOLVE_UPD2") && !atoi(getenv("WAE_LU_SOLVE_UPD2")));
  const bool stage_ok = !(getenv("WAE_LU_SOLVE_STAGE") && !atoi(getenv("WAE_LU_SOLVE_STAGE")));  // 4 ... 8 rhs: backward kernels with staged x rows
  // a level goes through the fused kernels when it has at least this many supernodes (one CTA each must fill the GPU); tests set 1
  const int fused_min = getenv("WAE_LU_SOLVE_FUSED_MIN") ? atoi(getenv("WAE_LU_SOLVE_FUSED_MIN")) : 2 * h->sm_count;
  // WAE_LU_TRACE=2 (diagnostic): the four kernel classes of a sweep pair timed per tree depth with their own events (serialising)
  const bool trace = getenv("WAE_LU_TRACE") && atoi(getenv("WAE_LU_TRACE")) == 2;
  std::vector<std::array<double, 4>> tms(maxd + 1);
  for (auto& a : tms) a.fill(0.0);
  cudaEvent_t te0 = nullptr, te1 = nullptr;
  if (trace) {
    CUDA_CHECK(cudaEventCreate(&te0));
    CUDA_CHECK(cudaEventCreate(&te1));
  }
  int cur_d = 0;
  auto timed = [&](int cls, auto&& launch) {
    if (!trace) {
      launch();
      return;
    }
    cudaEventRecord(te0, st);
    launch();
    cudaEventRecord(te1, st);
    cudaEventSynchronize(te1);
    float ms = 0;
    cudaEventElapsedTime(&ms, te0, te1);
    tms[cur_d][cls] += ms;
  };
  for (int d = maxd; d >= 0; d--) {
    cur_d = d;
    const std::vector<int32_t>& L = Y.levels[d];
    int max_s = 0, max_ld = 0;
    for (int32_t k : L) {
      const int s = Y.sn_first[k + 1] - Y.sn_first[k];
      max_s = std::max(max_s, s);
      max_ld = std::max(max_ld, s + (int)(Y.struct_ptr[k + 1] - Y.struct_ptr[k]));
    }
    for (int z0 = 0; z0 < (int)L.size(); z0 += 32768) {
      const int zc = std::min<int>(32768, (int)L.size() - z0);
      const int32_t* lst = S.d_level[d].p + z0;
      if (fused_ok && max_s <= LU_FW && (int)L.size() >= fused_min) {  // deep level: one CTA per supernode does triangle + update
        timed(0, [&] { lu_fwd_fused_kernel<NR><<<dim3(zc, zr), 256, 0, st>>>(D, lst, fwd_up, nrhs, Y.n, y); });
        h->launches++;
        continue;
      }
      for (int c_lo = 0; c_lo < max_s; c_lo += W) {
        timed(0, [&] {
          if (S.wininv)
            lu_fwd_win_kernel<WNR><<<dim3(zc * LU_WIN_CLUSTER, (nrhs + WNR - 1) / WNR), 1024, 0, st>>>(D, lst, fwd_up, c_lo, nrhs, y);
          else
            lu_fwd_tri_kernel<<<dim3(zc, nrhs), 1024, 0, st>>>(D, lst, fwd_up | tri_pf, c_lo, Y.n, y);
        });
        h->launches++;
        const int rows = max_ld - std::min(c_lo + W, max_s);  // upper bound of the rows below the window over the level
        if (max_ld > c_lo + 1 && rows + W > 0) {
          const int rmax = max_ld - c_lo;  // a supernode with a short pivot block has its window end (and first row) earlier
          const dim3 g((rmax + 63) / 64, zc, zr);
          timed(1, [&] {
            if constexpr (NR <= 2) {
              if (upd2_ok && (int)(g.x * g.y) < 2 * h->sm_count) {  // few CTAs: 16 column parts per row instead of 4
                lu_fwd_update2_kernel<NR, 16><<<g, 1024, 0, st>>>(D, lst, fwd_up, c_lo, nrhs, Y.n, y);
                return;
              }
            }
            lu_fwd_update_kernel<NR><<<g, 256, 0, st>>>(D, lst, fwd_up, c_lo, nrhs, Y.n, y);
          });
          h->launches++;
        }
      }
    }
  }
  for (int d = 0; d <= maxd; d++) {
    cur_d = d;
    const std::vector<int32_t>& L = Y.levels[d];
    int max_s = 0, max_ld = 0;
    for (int32_t k : L) {
      const int s = Y.sn_first[k + 1] - Y.sn_first[k];
      max_s = std::max(max_s, s);
      max_ld = std::max(max_ld, s + (int)(Y.struct_ptr[k + 1] - Y.struct_ptr[k]));
    }
    for (int z0 = 0; z0 < (int)L.size(); z0 += 32768) {
      const int zc = std::min<int>(32768, (int)L.size() - z0);
      const int32_t* lst = S.d_level[d].p + z0;
      if (fused_ok && max_s <= LU_FW && (int)L.size() >= fused_min) {
        timed(3, [&] {
          if constexpr (NR >= 4) {
            if (stage_ok) {
              lu_bwd_fused2_kernel<NR><<<dim3(zc, zr), 256, 0, st>>>(D, lst, !fwd_up, nrhs, Y.n, y);
              return;
            }
          }
          lu_bwd_fused_kernel<NR><<<dim3(zc, zr), 256, 0, st>>>(D, lst, !fwd_up, nrhs, Y.n, y);
        });
        h->launches++;
        continue;
      }
      for (int c_lo = ((max_s - 1) / W) * W; c_lo >= 0; c_lo -= W) {
        const int rmax = max_ld - c_lo;
        if (rmax > 1) {
          // few supernodes on the level: cut the rows into chunks so that the level still fills the GPU
          const int cpc = upd2_ok ? 8 * LU_CW(NR) : 8;  // columns per CTA: 8 warps x CW columns (round 2) or 8 warps x 1
          const int ncg = (std::min(W, max_s - c_lo) + cpc - 1) / cpc;
          int row_chunk = rmax;
          if (zc * ncg < 4 * h->sm_count) row_chunk = std::max(256, (int)((int64_t)rmax * zc * ncg / (4 * h->sm_count)) / 32 * 32 + 32);
          const int nrc = (rmax + row_chunk - 1) / row_chunk;
          timed(2, [&] {
            if constexpr (NR >= 4) {
              if (upd2_ok && stage_ok) {
                lu_bwd_update3_kernel<NR><<<dim3(ncg * nrc, zc, zr), 256, 0, st>>>(D, lst, !fwd_up, c_lo, ncg, row_chunk, nrhs, Y.n, y);
                return;
              }
            }
            if (upd2_ok)
              lu_bwd_update2_kernel<NR><<<dim3(ncg * nrc, zc, zr), 256, 0, st>>>(D, lst, !fwd_up, c_lo, ncg, row_chunk, nrhs, Y.n, y);
            else
              lu_bwd_update_kernel<NR><<<dim3(ncg * nrc, zc, zr), 256, 0, st>>>(D, lst, !fwd_up, c_lo, ncg, row_chunk, nrhs, Y.n, y);
          });
          h->launches++;
        }
        timed(3, [&] {
          if (S.wininv)
            lu_bwd_win_kernel<WNR><<<dim3(zc * LU_WIN_CLUSTER, (nrhs + WNR - 1) / WNR), 1024, 0, st>>>(D, lst, (int)!fwd_up, c_lo, nrhs, y);
          else
            lu_bwd_tri_kernel<<<dim3(zc, nrhs), 1024, 0, st>>>(D, lst, (int)!fwd_up | tri_pf, c_lo, Y.n, y);
        });
        h->launches++;
      }
    }
  }
  if (trace) {
    double tot[4] = {0, 0, 0, 0};
    fprintf(stderr, "[wae solve trace] nrhs %d  depth  fronts  max_s  max_ld    fwd_tri   fwd_upd   bwd_upd   bwd_tri (ms)\n", nrhs);
    for (int d = maxd; d >= 0; d--) {
      int max_s = 0, max_ld = 0;
      for (int32_t k : Y.levels[d]) {
        const int s = Y.sn_first[k + 1] - Y.sn_first[k];
        max_s = std::max(max_s, s);
        max_ld = std::max(max_ld, s + (int)(Y.struct_ptr[k + 1] - Y.struct_ptr[k]));
      }
      double mb = 0.0;  // one panel (L or U^T) of every supernode of the depth: what one sweep reads
      for (int32_t k : Y.levels[d]) {
        const double s = Y.sn_first[k + 1] - Y.sn_first[k];
        mb += (s + (double)(Y.struct_ptr[k + 1] - Y.struct_ptr[k])) * s * sizeof(cplx) / 1e6;
      }
      fprintf(stderr, "[wae solve trace]         %5d %7d %6d %7d  %9.3f %9.3f %9.3f %9.3f   %8.1f MB  fwd %5.0f bwd %5.0f GB/s\n", d, (int)Y.levels[d].size(), max_s,
              max_ld, tms[d][0], tms[d][1], tms[d][2], tms[d][3], mb, mb / (tms[d][0] + tms[d][1]), mb / (tms[d][2] + tms[d][3]));
      for (int c = 0; c < 4; c++) tot[c] += tms[d][c];
    }
    fprintf(stderr, "[wae solve trace] total                              %9.3f %9.3f %9.3f %9.3f\n", tot[0], tot[1], tot[2], tot[3]);
    cudaEventDestroy(te0);
    cudaEventDestroy(te1);
  }
}

static void lu_sweeps_direct(wae_ctx* h, LuSolver& S, int trans_t, int nrhs, cplx* y) {
  if (nrhs >= 8) lu_sweeps_nr<8>(h, S, trans_t, nrhs, y);
  else if (nrhs >= 4) lu_sweeps_nr<4>(h, S, trans_t, nrhs, y);
  else if (nrhs >= 2) lu_sweeps_nr<2>(h, S, trans_t, nrhs, y);
  else lu_sweeps_nr<1>(h, S, trans_t, nrhs, y);
}

LuSolver::~LuSolver() {
  for (auto& kv : sweep_graphs)
    if (kv.second.exec) cudaGraphExecDestroy((cudaGraphExec_t)kv.second.exec);
}

// A sweep pair is ~190 small, dependent launches (two per window of every top level); its sequence is a function of the symbolic
// structure and of (transposition, right-hand sides, work vector, kernel switches) alone.  It is captured into a CUDA graph the first
// time a key is seen and replayed afterwards: the dependent launches then follow each other without the per-launch stream overhead.
// Not under WAE_LU_TRACE, and WAE_LU_GRAPH=0 switches it off; any failure of the capture falls back to the direct launches.
static void lu_sweeps(wae_ctx* h, LuSolver& S, int trans_t, int nrhs, cplx* y) {
  cudaStream_t st = h->stream;
  const bool off = (getenv("WAE_LU_GRAPH") && !atoi(getenv("WAE_LU_GRAPH"))) || getenv("WAE_LU_TRACE");
  if (off) {
    lu_sweeps_direct(h, S, trans_t, nrhs, y);
    return;
  }
  auto envi = [](const char* n, int64_t dflt) { return getenv(n) ? (int64_t)atoll(getenv(n)) : dflt; };
  const std::array<int64_t, 8> key = {trans_t, nrhs, (int64_t)(uintptr_t)y, S.wininv ? 1 : 0, envi("WAE_LU_SOLVE_FUSED", 1), envi("WAE_LU_SOLVE_UPD2", 1),
                                      envi("WAE_LU_SOLVE_FUSED_MIN", -1), envi("WAE_LU_SOLVE_PF", 0) | (envi("WAE_LU_SOLVE_STAGE", 1) << 4)};
  auto it = S.sweep_graphs.find(key);
  if (it == S.sweep_graphs.end()) {
    LuSolver::SweepGraph g;
    const int64_t l0 = h->launches;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    // captured on the context's own auxiliary stream (the context stream may be the legacy default stream, which cannot capture); the
    // graph is then launched on the context stream like any other work
    if (!h->cap_stream && cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking) != cudaSuccess) h->cap_stream = nullptr;
    bool ok = h->cap_stream && cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeRelaxed) == cudaSuccess;
    if (ok) {
      h->stream = h->cap_stream;
      try {
        lu_sweeps_direct(h, S, trans_t, nrhs, y);  // recorded, not executed
      } catch (...) {
        h->stream = st;
        cudaStreamEndCapture(h->cap_stream, &graph);
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        throw;
      }
      h->stream = st;
      ok = cudaStreamEndCapture(h->cap_stream, &graph) == cudaSuccess && graph != nullptr;
    }
    if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
    if (graph) cudaGraphDestroy(graph);
    g.launches = h->launches - l0;
    h->launches = l0;
    if (!ok) {
      cudaGetLastError();  // the failed capture must not surface as a launch error
      g.exec = nullptr;
    } else
      g.exec = exec;
    it = S.sweep_graphs.emplace(key, g).first;
  }
  if (!it->second.exec) {  // capture failed once for this key: direct launches from now on
    lu_sweeps_direct(h, S, trans_t, nrhs, y);
    return;
  }
  CUDA_CHECK(cudaGraphLaunch((cudaGraphExec_t)it->second.exec, st));
  h->launches += it->second.launches;
}

// ---- rank-k (Sherman-Morrison-Woodbury) correction on top of the symmetric factorisation ------------------------------
//   A = S + Sm F Gm^T :  A^-1 b = y - Z  (F^-1 + Gm^T Z )^-1 Gm^T y,   y = S^-1 b, Z  = S^-1 Sm
//   A^T                : A^-T b = y - Zt (F^-1 + Sm^T Zt)^-1 Sm^T y,   Zt = S^-1 Gm
// t[j + k*r] += sum_p W[p + n*j] * X[p + n*r]   (plain transpose, no conjugation); grid (chunks, k, nrhs)
__global__ void __launch_bounds__(256) r1_dots_kernel(const cplx* __restrict__ W, int64_t n, int k, const cplx* __restrict__ X, cplx* __restrict__ t) {
  const cplx* w = W + (size_t)blockIdx.y * n;
  const cplx* x = X + (size_t)blockIdx.z * n;
  double sr = 0.0, si = 0.0;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    cplx a = w[p], b = x[p];
    sr += a.x * b.x - a.y * b.y;
    si += a.x * b.y + a.y * b.x;
  }
#pragma unroll
  for (int off = 16; off; off >>= 1) {
    sr += __shfl_xor_sync(0xffffffffu, sr, off);
    si += __shfl_xor_sync(0xffffffffu, si, off);
  }
  if ((threadIdx.x & 31) == 0 && (sr != 0.0 || si != 0.0)) {
    cplx* d = t + blockIdx.y + (size_t)k * blockIdx.z;
    atomicAdd(&d->x, sr);
    atomicAdd(&d->y, si);
  }
}
// X[:, r] -= Zx * (Kinv * t[:, r])
__global__ void r1_apply_kernel(const cplx* __restrict__ Zx, const cplx* __restrict__ Kinv, const cplx* __restrict__ t, int64_t n, int k,
                                int nrhs, cplx* __restrict__ X) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  for (int r = 0; r < nrhs; r++) {
    cplx acc = X[p + (size_t)r * n];
    for (int j = 0; j < k; j++) {
      cplx c = make_double2(0.0, 0.0);
      for (int i = 0; i < k; i++) {
        cplx m = Kinv[j + i * k], ti = t[i + (size_t)k * r];
        c.x += m.x * ti.x - m.y * ti.y;
        c.y += m.x * ti.y + m.y * ti.x;
      }
      cplx z = Zx[p + (size_t)j * n];
      acc.x -= z.x * c.x - z.y * c.y;
      acc.y -= z.x * c.y + z.y * c.x;
    }
    X[p + (size_t)r * n] = acc;
  }
}
__global__ void conj_kernel(int64_t total, cplx* __restrict__ x) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) x[i].y = -x[i].y;
}
__global__ void add_kernel(const cplx* __restrict__ a, int64_t total, cplx* __restrict__ x) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) x[i] = make_double2(x[i].x + a[i].x, x[i].y + a[i].y);
}

// X <- (LU)^-1 X or (LU)^-T X in the original ordering (factors only: no rank-k correction, no refinement)
void wae_lu_base_solve(wae_ctx* h, LuSolver& S, int tt, int nrhs, cplx* d_X) {
  const int64_t n = S.sym.n;
  cudaStream_t st = h->stream;
  if (S.work_nrhs < nrhs) {
    S.d_work.alloc((size_t)3 * n * nrhs);
    S.work_nrhs = nrhs;
  }
  cplx* y = S.d_work.p;
  unsigned gb = (unsigned)((n + 255) / 256);
  lu_permute_in_kernel<<<gb, 256, 0, st>>>(d_X, S.d_perm.p, S.d_scale.p, n, nrhs, 0, y);
  lu_sweeps(h, S, tt, nrhs, y);
  lu_permute_out_kernel<<<gb, 256, 0, st>>>(y, S.d_perm.p, S.d_scale.p, n, nrhs, 0, 0, d_X);
  h->launches += 2;
}

// X <- op(A)^-1 X with the rank-k correction (trans: 0 N, 1 T, 2 C)
static void lu_apply_inverse(wae_ctx* h, LuSolver& S, int trans, int nrhs, cplx* d_X) {
  const int64_t n = S.sym.n;
  cudaStream_t st = h->stream;
  const unsigned gt = (unsigned)(((size_t)n * nrhs + 255) / 256);
  if (trans == 2) conj_kernel<<<gt, 256, 0, st>>>(n * nrhs, d_X);  // A^H x = b  <=>  A^T conj(x) = conj(b)
  const int tt = trans != 0;
  wae_lu_base_solve(h, S, tt, nrhs, d_X);
  if (S.r1_k > 0) {
    const int k = S.r1_k;
    S.d_r1_t.reserve((size_t)k * nrhs);
    CUDA_CHECK(cudaMemsetAsync(S.d_r1_t.p, 0, (size_t)k * nrhs * sizeof(cplx), st));
    int chunks = (int)std::min<int64_t>((n + 255) / 256, 64);
    r1_dots_kernel<<<dim3(chunks, k, nrhs), 256, 0, st>>>(tt ? S.d_r1_Sm.p : S.d_r1_Gm.p, n, k, d_X, S.d_r1_t.p);
    r1_apply_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(tt ? S.d_r1_Zt.p : S.d_r1_Z.p, tt ? S.d_r1_KinvT.p : S.d_r1_Kinv.p, S.d_r1_t.p,
                                                                 n, k, nrhs, d_X);
    h->launches += 2;
  }
  if (trans == 2) conj_kernel<<<gt, 256, 0, st>>>(n * nrhs, d_X);
}

// X2 (n x 2, device): column 0 <- A^{-1} x0, column 1 <- A^{-H} x1 in ONE pass over the factor.  Only with the symmetric elimination:
// the factorised part S is complex symmetric, so A^{-H} x = conj(A^{-T} conj(x)) needs the same S^{-1} sweeps as A^{-1} x (S^{-T} = S^{-1});
// the two columns differ in the rank-k correction alone (flame terms: Woodbury with the (Gm, Z) resp. (Sm, Zt) vectors).  No refinement.
void wae_lu_solve_pair_device(wae_ctx* h, LuSolver& S, cplx* d_X2) {
  if (!S.factored || !S.sym_mode) WAE_THROW(WAE_E_INVALID, "paired solve needs a symmetric-mode factorisation");
  const int64_t n = S.sym.n;
  cudaStream_t st = h->stream;
  const unsigned gn = (unsigned)((n + 255) / 256);
  conj_kernel<<<gn, 256, 0, st>>>(n, d_X2 + n);
  wae_lu_base_solve(h, S, 0, 2, d_X2);
  if (S.r1_k > 0) {
    const int k = S.r1_k;
    const int chunks = (int)std::min<int64_t>((n + 255) / 256, 64);
    S.d_r1_t.reserve((size_t)k * 2);
    for (int tt = 0; tt < 2; tt++) {
      cplx* x = d_X2 + (size_t)tt * n;
      CUDA_CHECK(cudaMemsetAsync(S.d_r1_t.p, 0, (size_t)k * sizeof(cplx), st));
      r1_dots_kernel<<<dim3(chunks, k, 1), 256, 0, st>>>(tt ? S.d_r1_Sm.p : S.d_r1_Gm.p, n, k, x, S.d_r1_t.p);
      r1_apply_kernel<<<gn, 256, 0, st>>>(tt ? S.d_r1_Zt.p : S.d_r1_Z.p, tt ? S.d_r1_KinvT.p : S.d_r1_Kinv.p, S.d_r1_t.p, n, k, 1, x);
    }
    h->launches += 4;
  }
  conj_kernel<<<gn, 256, 0, st>>>(n, d_X2 + n);
  h->launches += 2;
  CUDA_CHECK(cudaGetLastError());
}

// x <- x + r, and with it the Beyn moments A_p += w z^p x (the epilogue of the last refinement step of a quadrature node's solve)
__global__ void add_moment_kernel(const cplx* __restrict__ r, int64_t total, cplx* __restrict__ x, int n_mom, cplx w, cplx z, cplx* __restrict__ A) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  cplx v = x[i];
  if (r) {
    v.x += r[i].x;
    v.y += r[i].y;
    x[i] = v;
  }
  cplx f = w;
  for (int p = 0; p < n_mom; p++) {
    cplx* a = A + (size_t)p * total + i;
    a->x += f.x * v.x - f.y * v.y;
    a->y += f.x * v.y + f.y * v.x;
    f = make_double2(f.x * z.x - f.y * z.y, f.x * z.y + f.y * z.x);
  }
}

void wae_lu_solve_device(wae_ctx* h, LuSolver& S, int trans, int nrhs, cplx* d_X, int refine, const LuMomentEpilogue* ep) {
  Family& F = h->fam(S.fam);
  if (!S.factored) WAE_THROW(WAE_E_INVALID, "wae_lu_factor has not been called (or failed)");
  const int64_t n = S.sym.n;
  cudaStream_t st = h->stream;
  if (S.work_nrhs < nrhs) {
    S.d_work.alloc((size_t)3 * n * nrhs);
    S.work_nrhs = nrhs;
  }
  cplx* b0 = S.d_work.p + (size_t)n * nrhs;   // copy of the right-hand side
  cplx* res = S.d_work.p + (size_t)2 * n * nrhs;
  const unsigned gt = (unsigned)(((size_t)n * nrhs + 255) / 256);
  if (refine > 0) CUDA_CHECK(cudaMemcpyAsync(b0, d_X, (size_t)n * nrhs * sizeof(cplx), cudaMemcpyDeviceToDevice, st));
  lu_apply_inverse(h, S, trans, nrhs, d_X);
  for (int it = 0; it < refine; it++) {
    // res = b - op(A) x ; x += op(A)^{-1} res
    if (trans == 0) {  // y = A x walks the CSR view: gather the values into that order once per factorisation, stream them afterwards
      if (!S.aval_csr_valid) {
        S.d_Aval_csr.reserve(S.d_Aval.n);
        wae_values_to_csr(h, F, S.d_Aval.p, S.d_Aval_csr.p);
        S.aval_csr_valid = true;
      }
      wae_spmm_values_csr(h, F, S.d_Aval_csr.p, nrhs, d_X, res);
    } else
      wae_spmm_values(h, F, S.d_Aval.p, trans, nrhs, d_X, res);
    lu_residual_kernel<<<gt, 256, 0, st>>>(b0, n * nrhs, res);
    lu_apply_inverse(h, S, trans, nrhs, res);
    if (ep && it == refine - 1)
      add_moment_kernel<<<gt, 256, 0, st>>>(res, n * nrhs, d_X, ep->n_mom, ep->w, ep->z, ep->A);
    else
      add_kernel<<<gt, 256, 0, st>>>(res, n * nrhs, d_X);
    h->launches += 2;
  }
  if (ep && refine <= 0) {
    add_moment_kernel<<<gt, 256, 0, st>>>(nullptr, n * nrhs, d_X, ep->n_mom, ep->w, ep->z, ep->A);
    h->launches++;
  }
  CUDA_CHECK(cudaGetLastError());
}
