// Operator family kernels: L(z) = sum_t f_t(z) A_t on the shared (union) pattern in one
// vectorised complex-fp64 pass (replaces the allocating sparse adds of
// src/NLEVP/LinOpFam.jl:499-522), and complex CSC/CSR SpMM for M*x, L'(z)*v.
#include <cuda_runtime.h>

#include <algorithm>

#include "wae_internal.h"

#define WAE_MAX_FUSED 8
struct CombineArgs {
  const double* val[WAE_MAX_FUSED];
  const int32_t* inv[WAE_MAX_FUSED];  // NULL: the term lives on the union pattern; else union nz -> term nz (-1: absent)
  double cr[WAE_MAX_FUSED], ci[WAE_MAX_FUSED];
  int is_complex[WAE_MAX_FUSED];
  int n;
};

// out[k] = (accumulate ? out[k] : 0) + sum_t coef_t * val_t[k]   for terms stored on the union pattern
__global__ void __launch_bounds__(256) combine_identity_kernel(CombineArgs a, int64_t nnz, int accumulate, double2* __restrict__ out) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride) {
    double re = 0.0, im = 0.0;
    int64_t qprev = k;
    if (accumulate) {
      double2 o = out[k];
      re = o.x;
      im = o.y;
    }
#pragma unroll
    for (int t = 0; t < WAE_MAX_FUSED; t++) {
      if (t < a.n) {
        // consecutive terms on the same sub-pattern (M, K) share the inverse map: one load serves both
        int64_t q = k;
        if (a.inv[t]) q = (t > 0 && a.inv[t] == a.inv[t - 1]) ? qprev : (int64_t)a.inv[t][k];
        qprev = q;
        if (q < 0) continue;
        if (a.is_complex[t]) {
          double2 v = reinterpret_cast<const double2*>(a.val[t])[q];
          re += a.cr[t] * v.x - a.ci[t] * v.y;
          im += a.cr[t] * v.y + a.ci[t] * v.x;
        } else {
          double v = a.val[t][q];
          re += a.cr[t] * v;
          im += a.ci[t] * v;
        }
      }
    }
    out[k] = make_double2(re, im);
  }
}

// out[map[k]] += coef * val[k]   (map injective inside one term: no conflicts)
__global__ void __launch_bounds__(256) combine_mapped_kernel(const double* __restrict__ val, int is_complex, double cr, double ci,
                                                             const int32_t* __restrict__ map, int64_t nnz, double2* __restrict__ out) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  double vr, vi = 0.0;
  if (is_complex) {
    double2 v = reinterpret_cast<const double2*>(val)[k];
    vr = v.x;
    vi = v.y;
  } else
    vr = val[k];
  double2 o = out[map[k]];
  o.x += cr * vr - ci * vi;
  o.y += cr * vi + ci * vr;
  out[map[k]] = o;
}

void wae_combine_device(wae_ctx* h, Family& F, const double* coeffs, int slot) {
  if (slot < 0 || slot >= WAE_FAMILY_SLOTS) WAE_THROW(WAE_E_INVALID, "slot %d out of range", slot);
  Pattern& U = h->pat(F.pattern);
  if (!F.slot[slot].p) F.slot[slot].alloc(2 * (size_t)U.nnz);
  F.slot_coeffs[slot].assign(coeffs, coeffs + 2 * F.n_terms);
  double2* out = (double2*)F.slot[slot].p;
  int blocks = (int)std::min<int64_t>((U.nnz + 255) / 256, (int64_t)h->sm_count * 8);
  if (blocks < 1) blocks = 1;
  CombineArgs a;
  a.n = 0;
  int accumulate = 0;
  bool any = false;
  auto flush = [&]() {
    combine_identity_kernel<<<blocks, 256, 0, h->stream>>>(a, U.nnz, accumulate, out);
    h->launches++;
    accumulate = 1;
    a.n = 0;
    any = true;
  };
  for (int t = 0; t < F.n_terms; t++) {
    double cr = coeffs[2 * t], ci = coeffs[2 * t + 1];
    if ((!F.identity[t] && F.inv_of[t] < 0) || (cr == 0.0 && ci == 0.0)) continue;
    Matrix& M = h->mat(F.mats[t]);
    a.val[a.n] = M.d_val.p;
    a.inv[a.n] = F.identity[t] ? nullptr : F.d_inv[F.inv_of[t]].p;
    a.cr[a.n] = cr;
    a.ci[a.n] = ci;
    a.is_complex[a.n] = M.is_complex;
    if (++a.n == WAE_MAX_FUSED) flush();
  }
  if (a.n || !any) flush();  // also zero-fills when no identity term is active
  for (int t = 0; t < F.n_terms; t++) {
    double cr = coeffs[2 * t], ci = coeffs[2 * t + 1];
    if (F.identity[t] || F.inv_of[t] >= 0 || (cr == 0.0 && ci == 0.0)) continue;
    Matrix& M = h->mat(F.mats[t]);
    int64_t nnz = h->pat(M.pattern).nnz;
    if (!nnz) continue;
    combine_mapped_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, h->stream>>>(M.d_val.p, M.is_complex, cr, ci, F.d_map[F.map_of[t]].p, nnz, out);
    h->launches++;
  }
  CUDA_CHECK(cudaGetLastError());
}

// ---- SpMM --------------------------------------------------------------------------------
// Column gather (CSC): y_j = sum_i op(A_ij) x_i  -> computes A^T x (trans=1) or A^H x (trans=2).
// Row gather through the CSR view (rowptr, colidx, perm into the CSC values): y = A x (trans=0).
// One warp per output entry, lanes over the nonzeros, shuffle reduction.
__global__ void __launch_bounds__(256) spmm_gather_kernel(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                                          const int32_t* __restrict__ perm, const double2* __restrict__ val, int conj,
                                                          int64_t dim, int nrhs, const double2* __restrict__ X, double2* __restrict__ Y) {
  int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (w >= dim) return;
  int64_t b = ptr[w], e = ptr[w + 1];
  for (int r = 0; r < nrhs; r++) {
    const double2* x = X + (size_t)r * dim;
    double sr = 0.0, si = 0.0;
    for (int64_t k = b + lane; k < e; k += 32) {
      double2 a = val[perm ? perm[k] : k];
      if (conj) a.y = -a.y;
      double2 xv = x[idx[k]];
      sr += a.x * xv.x - a.y * xv.y;
      si += a.x * xv.y + a.y * xv.x;
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) {
      sr += __shfl_xor_sync(0xffffffffu, sr, off);
      si += __shfl_xor_sync(0xffffffffu, si, off);
    }
    if (lane == 0) Y[(size_t)r * dim + w] = make_double2(sr, si);
  }
}

void wae_family_ensure_csr(wae_ctx* h, Family& F) {
  if (F.tr_built) return;
  Pattern& U = h->pat(F.pattern);
  std::vector<int64_t> rowptr(U.dim + 1, 0);
  for (int64_t k = 0; k < U.nnz; k++) rowptr[U.rowval[k] + 1]++;
  for (int64_t i = 0; i < U.dim; i++) rowptr[i + 1] += rowptr[i];
  std::vector<int32_t> colidx(U.nnz), perm(U.nnz);
  std::vector<int64_t> pos(rowptr.begin(), rowptr.end() - 1);
  for (int64_t j = 0; j < U.dim; j++)
    for (int64_t k = U.colptr[j]; k < U.colptr[j + 1]; k++) {
      int64_t q = pos[U.rowval[k]]++;
      colidx[q] = (int32_t)j;
      perm[q] = (int32_t)k;
    }
  F.d_rowptr.upload(rowptr, h->stream);
  F.d_colidx.upload(colidx, h->stream);
  F.d_tr_perm.upload(perm, h->stream);
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  F.tr_built = true;
}

void wae_spmm_device(wae_ctx* h, Family& F, int slot, int trans, int nrhs, const cplx* X, cplx* Y) {
  wae_spmm_values(h, F, (const cplx*)F.slot[slot].p, trans, nrhs, X, Y);
}

// val_csr[q] = val[perm[q]]: the values of the union pattern in the order of the CSR view (one gather, so that repeated products y = A x with
// the same values -- the refinement steps of the solves -- stream them instead of gathering them)
__global__ void permute_values_kernel(const double2* __restrict__ val, const int32_t* __restrict__ perm, int64_t nnz, double2* __restrict__ out) {
  int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < nnz) out[q] = val[perm[q]];
}
void wae_values_to_csr(wae_ctx* h, Family& F, const cplx* val, cplx* val_csr) {
  Pattern& U = h->pat(F.pattern);
  wae_family_ensure_csr(h, F);
  permute_values_kernel<<<(unsigned)((U.nnz + 255) / 256), 256, 0, h->stream>>>(val, F.d_tr_perm.p, U.nnz, val_csr);
  h->launches++;
  CUDA_CHECK(cudaGetLastError());
}
// Y = A X with the values already in CSR order (wae_values_to_csr)
void wae_spmm_values_csr(wae_ctx* h, Family& F, const cplx* val_csr, int nrhs, const cplx* X, cplx* Y) {
  Pattern& U = h->pat(F.pattern);
  wae_family_ensure_csr(h, F);
  spmm_gather_kernel<<<(unsigned)((U.dim * 32 + 255) / 256), 256, 0, h->stream>>>(F.d_rowptr.p, F.d_colidx.p, nullptr, val_csr, 0, U.dim, nrhs, X, Y);
  h->launches++;
  CUDA_CHECK(cudaGetLastError());
}

void wae_spmm_values(wae_ctx* h, Family& F, const cplx* val, int trans, int nrhs, const cplx* X, cplx* Y) {
  Pattern& U = h->pat(F.pattern);
  unsigned blocks = (unsigned)((U.dim * 32 + 255) / 256);
  if (trans == 0) {
    wae_family_ensure_csr(h, F);
    spmm_gather_kernel<<<blocks, 256, 0, h->stream>>>(F.d_rowptr.p, F.d_colidx.p, F.d_tr_perm.p, val, 0, U.dim, nrhs, X, Y);
  } else {
    spmm_gather_kernel<<<blocks, 256, 0, h->stream>>>(U.d_colptr.p, U.d_rowval.p, nullptr, val, trans == 2, U.dim, nrhs, X, Y);
  }
  h->launches++;
  CUDA_CHECK(cudaGetLastError());
}

// out += (cr + i ci) * A_t  for one term of the family (out lives on the union pattern)
__global__ void __launch_bounds__(256) axpy_identity_kernel(const double* __restrict__ val, int is_complex, double cr, double ci, int64_t nnz,
                                                            double2* __restrict__ out) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  double vr, vi = 0.0;
  if (is_complex) {
    double2 v = reinterpret_cast<const double2*>(val)[k];
    vr = v.x;
    vi = v.y;
  } else
    vr = val[k];
  double2 o = out[k];
  o.x += cr * vr - ci * vi;
  o.y += cr * vi + ci * vr;
  out[k] = o;
}

void wae_axpy_term(wae_ctx* h, Family& F, int t, double cr, double ci, cplx* out) {
  Matrix& M = h->mat(F.mats[t]);
  int64_t nnz = h->pat(M.pattern).nnz;
  if (!nnz) return;
  unsigned blocks = (unsigned)((nnz + 255) / 256);
  if (F.identity[t])
    axpy_identity_kernel<<<blocks, 256, 0, h->stream>>>(M.d_val.p, M.is_complex, cr, ci, nnz, out);
  else
    combine_mapped_kernel<<<blocks, 256, 0, h->stream>>>(M.d_val.p, M.is_complex, cr, ci, F.d_map[F.map_of[t]].p, nnz, out);
  h->launches++;
}
