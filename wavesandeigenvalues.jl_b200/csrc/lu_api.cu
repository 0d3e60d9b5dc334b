// C-ABI entry points for the sparse LU, the shift-invert Arnoldi eigensolver and the Beyn
// moment accumulation (see include/wae_b200.h).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "lu.h"

typedef std::complex<double> zc;

#define WAE_API_BEGIN \
  if (!h) return WAE_E_INVALID; \
  try {
#define WAE_API_END                    \
  }                                    \
  catch (const WaeError& e) {          \
    h->err = e.msg;                    \
    return e.code;                     \
  }                                    \
  catch (const std::bad_alloc&) {      \
    h->err = "host allocation failed"; \
    return WAE_E_NOMEM;                \
  }                                    \
  catch (const std::exception& e) {    \
    h->err = e.what();                 \
    return WAE_E_INVALID;              \
  }                                    \
  return WAE_OK;

int wae_lu_family_of(const LuSolver& S) { return S.fam; }

static LuSolver& get_lu(wae_ctx* h, int id) {
  if (id < 0 || id >= (int)h->lus.size() || !h->lus[id]) WAE_THROW(WAE_E_INVALID, "unknown LU id %d", id);
  return *h->lus[id];
}

// coordinates of every DOF (vertices, P2: edge midpoints) when the family lives on the context mesh
static bool dof_coords(wae_ctx* h, int64_t dim, std::vector<double>& xyz) {
  if (!h->order || h->dim != dim || h->n_tet == 0) return false;
  xyz.assign((size_t)3 * dim, 0.0);
  std::copy(h->xyz.begin(), h->xyz.end(), xyz.begin());
  if (h->order == 2) {
    static const int ea[6] = {0, 0, 0, 1, 1, 2}, eb[6] = {1, 2, 3, 2, 3, 3};
    for (int64_t e = 0; e < h->n_tet; e++) {
      const uint32_t* d = h->tets.data() + (size_t)e * 10;
      for (int k = 0; k < 6; k++)
        for (int r = 0; r < 3; r++) xyz[3 * (size_t)d[4 + k] + r] = 0.5 * (h->xyz[3 * (size_t)d[ea[k]] + r] + h->xyz[3 * (size_t)d[eb[k]] + r]);
    }
  }
  return true;
}

// ---- small dense complex eigenproblem (upper Hessenberg, QR algorithm) ------------------------------------
// H (n x n, row-major, destroyed) -> eigenvalues w and eigenvectors X (columns, row-major n x n)
static void hessenberg_eig(int n, std::vector<zc>& H, std::vector<zc>& w, std::vector<zc>& X) {
  auto A = [&](int i, int j) -> zc& { return H[(size_t)i * n + j]; };
  std::vector<zc> Z((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) Z[(size_t)i * n + i] = 1.0;
  const double eps = 2.2e-16;
  int hi = n - 1, iter = 0;
  while (hi > 0) {
    int l = hi;
    for (; l > 0; l--) {
      double sd = std::abs(A(l - 1, l - 1)) + std::abs(A(l, l));
      if (sd == 0.0) sd = 1.0;
      if (std::abs(A(l, l - 1)) <= eps * sd) {
        A(l, l - 1) = 0.0;
        break;
      }
    }
    if (l == hi) {
      hi--;
      iter = 0;
      continue;
    }
    // Wilkinson shift
    zc a = A(hi - 1, hi - 1), b = A(hi - 1, hi), c = A(hi, hi - 1), d = A(hi, hi);
    zc tr = a + d, det = a * d - b * c;
    zc disc = std::sqrt(tr * tr - 4.0 * det);
    zc m1 = 0.5 * (tr + disc), m2 = 0.5 * (tr - disc);
    zc mu = std::abs(m1 - d) < std::abs(m2 - d) ? m1 : m2;
    if (iter % 11 == 10) mu = d + std::abs(c);  // exceptional shift
    if (++iter > 30 * n) break;
    for (int i = l; i <= hi; i++) A(i, i) -= mu;
    std::vector<double> cs(n);
    std::vector<zc> sn(n);
    for (int k = l; k < hi; k++) {
      zc f = A(k, k), g = A(k + 1, k);
      double nr = std::sqrt(std::norm(f) + std::norm(g));
      double cc = 1.0;
      zc ss = 0.0;
      if (nr > 0.0) {
        if (std::abs(f) == 0.0) {
          cc = 0.0;
          ss = 1.0;
        } else {
          cc = std::abs(f) / nr;
          ss = (f / std::abs(f)) * std::conj(g) / nr;
        }
      }
      cs[k] = cc;
      sn[k] = ss;
      for (int j = k; j < n; j++) {  // rows k,k+1 <- G * rows
        zc x = A(k, j), y = A(k + 1, j);
        A(k, j) = cc * x + ss * y;
        A(k + 1, j) = -std::conj(ss) * x + cc * y;
      }
    }
    for (int k = l; k < hi; k++) {  // columns k,k+1 <- columns * G^H
      double cc = cs[k];
      zc ss = sn[k];
      for (int i = 0; i <= std::min(k + 2, hi); i++) {
        zc x = A(i, k), y = A(i, k + 1);
        A(i, k) = cc * x + std::conj(ss) * y;
        A(i, k + 1) = -ss * x + cc * y;
      }
      for (int i = 0; i < n; i++) {
        zc x = Z[(size_t)i * n + k], y = Z[(size_t)i * n + k + 1];
        Z[(size_t)i * n + k] = cc * x + std::conj(ss) * y;
        Z[(size_t)i * n + k + 1] = -ss * x + cc * y;
      }
    }
    for (int i = l; i <= hi; i++) A(i, i) += mu;
  }
  w.resize(n);
  for (int i = 0; i < n; i++) w[i] = A(i, i);
  // eigenvectors of the triangular factor, back-transformed with Z
  X.assign((size_t)n * n, 0.0);
  double tn = 0;
  for (int i = 0; i < n; i++)
    for (int j = i; j < n; j++) tn = std::max(tn, std::abs(A(i, j)));
  std::vector<zc> y(n);
  for (int k = 0; k < n; k++) {
    std::fill(y.begin(), y.end(), 0.0);
    y[k] = 1.0;
    for (int i = k - 1; i >= 0; i--) {
      zc s = 0.0;
      for (int j = i + 1; j <= k; j++) s += A(i, j) * y[j];
      zc den = A(i, i) - A(k, k);
      if (std::abs(den) < eps * std::max(tn, 1e-300)) den = eps * std::max(tn, 1e-300);
      y[i] = -s / den;
    }
    double nrm = 0;
    std::vector<zc> x(n, 0.0);
    for (int i = 0; i < n; i++) {
      zc s = 0.0;
      for (int j = 0; j <= k; j++) s += Z[(size_t)i * n + j] * y[j];
      x[i] = s;
      nrm += std::norm(s);
    }
    nrm = std::sqrt(nrm);
    for (int i = 0; i < n; i++) X[(size_t)i * n + k] = x[i] / nrm;
  }
}

// ---- device helpers for the Arnoldi process -----------------------------------------------------------------
// out[i] += sum_p conj(V_i[p]) w[p]   for i < nv   (grid.x chunks, grid.y = i)
__global__ void __launch_bounds__(256) multi_dot_kernel(const cplx* __restrict__ V, int64_t n, const cplx* __restrict__ w, double* __restrict__ out) {
  const cplx* v = V + (size_t)blockIdx.y * n;
  double sr = 0.0, si = 0.0;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    cplx a = v[p], b = w[p];
    sr += a.x * b.x + a.y * b.y;
    si += a.x * b.y - a.y * b.x;
  }
  __shared__ double sh[2][8];
#pragma unroll
  for (int off = 16; off; off >>= 1) {
    sr += __shfl_xor_sync(0xffffffffu, sr, off);
    si += __shfl_xor_sync(0xffffffffu, si, off);
  }
  if ((threadIdx.x & 31) == 0) {
    sh[0][threadIdx.x >> 5] = sr;
    sh[1][threadIdx.x >> 5] = si;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; k++) {
      sr += sh[0][k];
      si += sh[1][k];
    }
    atomicAdd(out + 2 * blockIdx.y, sr);
    atomicAdd(out + 2 * blockIdx.y + 1, si);
  }
}
// w -= sum_i c_i V_i
__global__ void multi_axpy_kernel(const cplx* __restrict__ V, int64_t n, int nv, const cplx* __restrict__ c, cplx* __restrict__ w) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  cplx a = w[p];
  for (int i = 0; i < nv; i++) {
    cplx ci = c[i], v = V[(size_t)i * n + p];
    a.x -= ci.x * v.x - ci.y * v.y;
    a.y -= ci.x * v.y + ci.y * v.x;
  }
  w[p] = a;
}
// out = sum_i c_i V_i
__global__ void lincomb_kernel(const cplx* __restrict__ V, int64_t n, int nv, const cplx* __restrict__ c, cplx* __restrict__ out) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  cplx a = make_double2(0.0, 0.0);
  for (int i = 0; i < nv; i++) {
    cplx ci = c[i], v = V[(size_t)i * n + p];
    a.x += ci.x * v.x - ci.y * v.y;
    a.y += ci.x * v.y + ci.y * v.x;
  }
  out[p] = a;
}
__global__ void scale_copy_kernel(const cplx* __restrict__ in, int64_t n, double s, cplx* __restrict__ out) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) out[p] = make_double2(in[p].x * s, in[p].y * s);
}
// X[:, c] = e_c for c < l
// start vector of the Arnoldi process: v += amp * r_i with a deterministic pseudo-random r_i in (-1, 1) (splitmix-style hash of i).
// ARPACK recovers eigenvectors the start vector is exactly orthogonal to (v0 = ones against an antisymmetric mode) from round-off
// during its >= ncv steps; this process stops as soon as the wanted Ritz pairs have converged, so the missing components are
// seeded explicitly instead.
__global__ void perturb_start_kernel(cplx* __restrict__ v, int64_t n, double amp) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t z = (uint64_t)i + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  const double a = (double)(z >> 11) * (1.0 / 9007199254740992.0), b = (double)((z * 0x9E3779B97F4A7C15ull) >> 11) * (1.0 / 9007199254740992.0);
  v[i].x += amp * (2.0 * a - 1.0);
  v[i].y += amp * (2.0 * b - 1.0);
}

__global__ void identity_cols_kernel(int64_t n, int l, cplx* __restrict__ X) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n * l) return;
  X[p] = make_double2((p % n) == (p / n) ? 1.0 : 0.0, 0.0);
}
// A_p[:, :] += w z^p X  for p < n_mom
__global__ void moment_accum_kernel(const cplx* __restrict__ X, int64_t total, int n_mom, cplx w, cplx z, cplx* __restrict__ A) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  cplx x = X[i];
  cplx f = w;
  for (int p = 0; p < n_mom; p++) {
    cplx* a = A + (size_t)p * total + i;
    a->x += f.x * x.x - f.y * x.y;
    a->y += f.x * x.y + f.y * x.x;
    f = make_double2(f.x * z.x - f.y * z.y, f.x * z.y + f.y * z.x);
  }
}

// out[idx[i] + n*col] = vals[i]  (dense complex copy of a sparse real vector)
__global__ void scatter_real_kernel(const double* __restrict__ vals, const int32_t* __restrict__ idx, int cnt, cplx* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cnt) out[idx[i]] = make_double2(vals[i], 0.0);
}
__global__ void __launch_bounds__(256) r1_gram_kernel(const cplx* __restrict__ W, int64_t n, int k, const cplx* __restrict__ Z, double* __restrict__ out) {
  // out[(i + k*j)] += sum_p W[p + n*i] * Z[p + n*j]   (plain transpose); grid (chunks, k, k)
  const cplx* w = W + (size_t)blockIdx.y * n;
  const cplx* z = Z + (size_t)blockIdx.z * n;
  double sr = 0.0, si = 0.0;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    cplx a = w[p], b = z[p];
    sr += a.x * b.x - a.y * b.y;
    si += a.x * b.y + a.y * b.x;
  }
#pragma unroll
  for (int off = 16; off; off >>= 1) {
    sr += __shfl_xor_sync(0xffffffffu, sr, off);
    si += __shfl_xor_sync(0xffffffffu, si, off);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out + 2 * (blockIdx.y + (size_t)k * blockIdx.z), sr);
    atomicAdd(out + 2 * (blockIdx.y + (size_t)k * blockIdx.z) + 1, si);
  }
}

// Numeric factorisation of a family slot.  If every active term is complex symmetric (M, K, C by construction) or a rank-1
// flame operator S (x) G, the symmetric part is eliminated LDL^T-style (half of the GEMM work) and the flame terms enter
// through the Sherman-Morrison-Woodbury formula; otherwise (user matrices, Bloch terms) the general LU is used.
static void factor_slot(wae_ctx* h, LuSolver& S, Family& F, int slot) {
  Pattern& U = h->pat(F.pattern);
  const cplx* A = (const cplx*)F.slot[slot].p;
  cudaStream_t st = h->stream;
  const char* env = getenv("WAE_LU_SYM");
  bool ok = !(env && atoi(env) == 0);
  const std::vector<double>& cf = F.slot_coeffs[slot];
  if ((int)cf.size() != 2 * F.n_terms) ok = false;
  std::vector<int> r1;
  for (int t = 0; ok && t < F.n_terms; t++) {
    if (cf[2 * t] == 0.0 && cf[2 * t + 1] == 0.0) continue;
    Matrix& M = h->mat(F.mats[t]);
    if (M.rank1) r1.push_back(t);
    else if (!M.symmetric) ok = false;
  }
  if (r1.size() > 8) ok = false;
  S.r1_k = 0;
  h->last_ms["factor_sym"] = ok ? 1.0 : 0.0;
  if (!ok) {
    wae_lu_factor_device(h, S, A, nullptr, 0);
    return;
  }
  const cplx* Sv = A;
  const int k = (int)r1.size();
  if (k) {
    S.d_Sval.reserve((size_t)U.nnz);
    CUDA_CHECK(cudaMemcpyAsync(S.d_Sval.p, A, (size_t)U.nnz * sizeof(cplx), cudaMemcpyDeviceToDevice, st));
    for (int t : r1) wae_axpy_term(h, F, t, -cf[2 * t], -cf[2 * t + 1], S.d_Sval.p);
    Sv = S.d_Sval.p;
  }
  wae_lu_factor_device(h, S, Sv, A, 1);
  if (!k) return;
  const int64_t n = S.sym.n;
  S.d_r1_Sm.reserve((size_t)n * k);
  S.d_r1_Gm.reserve((size_t)n * k);
  S.d_r1_Z.reserve((size_t)n * k);
  S.d_r1_Zt.reserve((size_t)n * k);
  S.d_r1_Kinv.reserve((size_t)k * k);
  S.d_r1_KinvT.reserve((size_t)k * k);
  CUDA_CHECK(cudaMemsetAsync(S.d_r1_Sm.p, 0, (size_t)n * k * sizeof(cplx), st));
  CUDA_CHECK(cudaMemsetAsync(S.d_r1_Gm.p, 0, (size_t)n * k * sizeof(cplx), st));
  for (int j = 0; j < k; j++) {
    Matrix& Q = h->mat(F.mats[r1[j]]);
    int nr = (int)Q.r1_rows.size(), nc = (int)Q.r1_cols.size();
    scatter_real_kernel<<<(nr + 255) / 256, 256, 0, st>>>(Q.d_r1_S.p, Q.d_r1_rows.p, nr, S.d_r1_Sm.p + (size_t)j * n);
    scatter_real_kernel<<<(nc + 255) / 256, 256, 0, st>>>(Q.d_r1_G.p, Q.d_r1_cols.p, nc, S.d_r1_Gm.p + (size_t)j * n);
    h->launches += 2;
  }
  // Z = S^-1 Sm and Zt = S^-1 Gm (S symmetric: S^-T = S^-1) as the 2 k right-hand sides of ONE pass over the factor
  S.d_r1_ZZ.reserve((size_t)n * 2 * k);
  CUDA_CHECK(cudaMemcpyAsync(S.d_r1_ZZ.p, S.d_r1_Sm.p, (size_t)n * k * sizeof(cplx), cudaMemcpyDeviceToDevice, st));
  CUDA_CHECK(cudaMemcpyAsync(S.d_r1_ZZ.p + (size_t)n * k, S.d_r1_Gm.p, (size_t)n * k * sizeof(cplx), cudaMemcpyDeviceToDevice, st));
  wae_lu_base_solve(h, S, 0, 2 * k, S.d_r1_ZZ.p);
  CUDA_CHECK(cudaMemcpyAsync(S.d_r1_Z.p, S.d_r1_ZZ.p, (size_t)n * k * sizeof(cplx), cudaMemcpyDeviceToDevice, st));
  CUDA_CHECK(cudaMemcpyAsync(S.d_r1_Zt.p, S.d_r1_ZZ.p + (size_t)n * k, (size_t)n * k * sizeof(cplx), cudaMemcpyDeviceToDevice, st));
  // K = F^-1 + Gm^T Z  (k x k), inverted on the host
  S.d_arn_dots.reserve(2 * (size_t)k * k + 64);
  CUDA_CHECK(cudaMemsetAsync(S.d_arn_dots.p, 0, 2 * (size_t)k * k * sizeof(double), st));
  int chunks = (int)std::min<int64_t>((n + 255) / 256, 64);
  r1_gram_kernel<<<dim3(chunks, k, k), 256, 0, st>>>(S.d_r1_Gm.p, n, k, S.d_r1_Z.p, S.d_arn_dots.p);
  h->launches++;
  std::vector<zc> K((size_t)k * k), Kinv((size_t)k * k, 0.0), KinvT((size_t)k * k);
  CUDA_CHECK(cudaMemcpyAsync(K.data(), S.d_arn_dots.p, 2 * (size_t)k * k * sizeof(double), cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  for (int j = 0; j < k; j++) K[j + (size_t)k * j] += 1.0 / zc(cf[2 * r1[j]], cf[2 * r1[j] + 1]);
  // Gauss-Jordan with partial pivoting (column-major k x k)
  for (int j = 0; j < k; j++) Kinv[j + (size_t)k * j] = 1.0;
  for (int c = 0; c < k; c++) {
    int piv = c;
    for (int r = c + 1; r < k; r++)
      if (std::abs(K[r + (size_t)k * c]) > std::abs(K[piv + (size_t)k * c])) piv = r;
    if (std::abs(K[piv + (size_t)k * c]) == 0.0) WAE_THROW(WAE_E_SINGULAR, "singular capacitance matrix in the rank-%d update", k);
    for (int q = 0; q < k; q++) {
      std::swap(K[c + (size_t)k * q], K[piv + (size_t)k * q]);
      std::swap(Kinv[c + (size_t)k * q], Kinv[piv + (size_t)k * q]);
    }
    zc d = 1.0 / K[c + (size_t)k * c];
    for (int q = 0; q < k; q++) {
      K[c + (size_t)k * q] *= d;
      Kinv[c + (size_t)k * q] *= d;
    }
    for (int r = 0; r < k; r++) {
      if (r == c) continue;
      zc f = K[r + (size_t)k * c];
      for (int q = 0; q < k; q++) {
        K[r + (size_t)k * q] -= f * K[c + (size_t)k * q];
        Kinv[r + (size_t)k * q] -= f * Kinv[c + (size_t)k * q];
      }
    }
  }
  for (int a = 0; a < k; a++)
    for (int b = 0; b < k; b++) KinvT[a + (size_t)k * b] = Kinv[b + (size_t)k * a];
  CUDA_CHECK(cudaMemcpyAsync(S.d_r1_Kinv.p, Kinv.data(), (size_t)k * k * sizeof(zc), cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaMemcpyAsync(S.d_r1_KinvT.p, KinvT.data(), (size_t)k * k * sizeof(zc), cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  S.r1_k = k;
}

struct ArnoldiWork {
  DevBuf<cplx> V, w, t, c;
  DevBuf<double> dots;
};

extern "C" {

// Host-only diagnostic (no GPU needed): statistics of the symbolic phase for a CSC pattern.
// out[0]=supernodes, [1]=factor nnz, [2]=flops, [3]=max pivot block, [4]=max structure, [5]=depth,
// [6]=factor array entries, [7]=max update-buffer entries of one depth
int32_t wae_lu_symbolic_stats(int64_t n, const int64_t* colptr, const int64_t* rowval, const double* coords, int32_t leaf, double* out) {
  try {
    std::vector<int32_t> rv(colptr[n]);
    for (int64_t k = 0; k < colptr[n]; k++) rv[k] = (int32_t)rowval[k];
    LuSymbolic S;
    wae_lu_symbolic(n, colptr, rv.data(), coords, leaf > 0 ? leaf : 64, S);
    out[0] = S.nsn;
    out[1] = (double)S.factor_nnz;
    out[2] = S.flops;
    out[3] = S.max_s;
    out[4] = S.max_r;
    out[5] = (double)S.levels.size();
    out[6] = (double)S.fac_size;
    out[7] = (double)*std::max_element(S.level_upd_size.begin(), S.level_upd_size.end());
    // verify that perm is a permutation
    std::vector<char> seen(n, 0);
    for (int64_t p = 0; p < n; p++) {
      if (seen[S.perm[p]]) return WAE_E_INVALID;
      seen[S.perm[p]] = 1;
    }
    return WAE_OK;
  } catch (...) {
    return WAE_E_INVALID;
  }
}

int32_t wae_lu_analyze(wae_ctx* h, int32_t fam_id, int32_t* lu_id, int64_t* factor_nnz, double* factor_flops) {
  WAE_API_BEGIN
  CUDA_CHECK(cudaSetDevice(h->device));
  Family& F = h->fam(fam_id);
  Pattern& U = h->pat(F.pattern);
  std::shared_ptr<LuSolver> S(new LuSolver());
  S->fam = fam_id;
  std::vector<double> xyz;
  bool have = dof_coords(h, U.dim, xyz);
  int leaf = 64;
  if (const char* env = getenv("WAE_LU_LEAF")) leaf = atoi(env);
  if (const char* env = getenv("WAE_LU_PIVOT_EPS")) S->pivot_eps = atof(env);
  if (const char* env = getenv("WAE_LU_REFINE")) S->refine_steps = atoi(env);
  if (const char* env = getenv("WAE_EIGS_REFINE")) S->eigs_refine = atoi(env);
  wae_lu_symbolic(U.dim, U.colptr.data(), U.rowval.data(), have ? xyz.data() : nullptr, leaf, S->sym);
  wae_lu_setup_device(h, *S);
  h->lus.push_back(S);
  if (lu_id) *lu_id = (int)h->lus.size() - 1;
  if (factor_nnz) *factor_nnz = S->sym.factor_nnz;
  if (factor_flops) *factor_flops = S->sym.flops;
  WAE_API_END
}

// Releases the factor storage (the largest allocation of the path: 23 GB at config 2), the symbolic data and the solve / Arnoldi
// workspaces of one analysis.
int32_t wae_lu_free(wae_ctx* h, int32_t lu_id) {
  WAE_API_BEGIN
  get_lu(h, lu_id);
  CUDA_CHECK(cudaSetDevice(h->device));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  h->lus[lu_id].reset();
  WAE_API_END
}

int32_t wae_lu_factor_ex(wae_ctx* h, int32_t lu_id, int32_t slot, int32_t check) {
  WAE_API_BEGIN
  CUDA_CHECK(cudaSetDevice(h->device));
  LuSolver& S = get_lu(h, lu_id);
  Family& F = h->fam(S.fam);
  if (slot < 0 || slot >= WAE_FAMILY_SLOTS || !F.slot[slot].p) WAE_THROW(WAE_E_INVALID, "family slot %d is empty", slot);
  struct Restore {  // the flag is a property of this call only
    LuSolver& S;
    ~Restore() { S.check_singular = true; }
  } restore{S};
  S.check_singular = check != 0;
  PhaseTimer t(h, "factor");
  factor_slot(h, S, F, slot);
  t.stop();
  // Safety net of the static pivoting (only when pivots WERE perturbed, and not for lu(...; check = false)): UMFPACK would have exchanged
  // rows; here a perturbed pivot that mattered shows as a solve whose refinement does not contract.  One probe solve A x = (1, ..., 1)
  // with one refinement step; a relative residual above 1e-6 raises WAE_E_SINGULAR instead of returning garbage from later solves.
  if (check != 0 && h->last_ms["static_pivots"] > 0.0) {
    const int64_t n = S.sym.n;
    cudaStream_t st = h->stream;
    S.d_io.reserve((size_t)3 * n);
    cplx *x = S.d_io.p, *r = S.d_io.p + n, *b = S.d_io.p + 2 * n;
    std::vector<zc> ones((size_t)n, zc(1.0, 0.0));
    CUDA_CHECK(cudaMemcpyAsync(b, ones.data(), (size_t)n * sizeof(cplx), cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaMemcpyAsync(x, b, (size_t)n * sizeof(cplx), cudaMemcpyDeviceToDevice, st));
    wae_lu_solve_device(h, S, 0, 1, x, 1);
    wae_spmm_values(h, F, S.d_Aval.p, 0, 1, x, r);  // r = A x
    S.d_arn_dots.reserve(64);
    auto norm2 = [&](const cplx* u, const cplx* v) {  // |u - v|^2 via three dots
      double hd[6];
      CUDA_CHECK(cudaMemsetAsync(S.d_arn_dots.p, 0, 6 * sizeof(double), st));
      const int blocks = (int)std::min<int64_t>((n + 255) / 256, 64);
      multi_dot_kernel<<<dim3(blocks, 1), 256, 0, st>>>(u, n, u, S.d_arn_dots.p);
      multi_dot_kernel<<<dim3(blocks, 1), 256, 0, st>>>(v, n, v, S.d_arn_dots.p + 2);
      multi_dot_kernel<<<dim3(blocks, 1), 256, 0, st>>>(u, n, v, S.d_arn_dots.p + 4);
      h->launches += 3;
      CUDA_CHECK(cudaMemcpyAsync(hd, S.d_arn_dots.p, sizeof(hd), cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaStreamSynchronize(st));
      return hd[0] + hd[2] - 2.0 * hd[4];
    };
    const double res2 = norm2(r, b);
    const double rel = std::sqrt(std::max(0.0, res2) / (double)n);
    h->last_ms["factor_probe_residual"] = rel;
    if (!(rel <= 1e-6)) {
      S.factored = false;
      WAE_THROW(WAE_E_SINGULAR,
                "%d pivots were perturbed and a probe solve does not reach its right-hand side (relative residual %.2e): the matrix is singular or needs row "
                "exchanges this factorisation does not do",
                (int)h->last_ms["static_pivots"], rel);
    }
  }
  WAE_API_END
}

int32_t wae_lu_factor(wae_ctx* h, int32_t lu_id, int32_t slot) { return wae_lu_factor_ex(h, lu_id, slot, 1); }

int32_t wae_lu_solve(wae_ctx* h, int32_t lu_id, int32_t trans, int32_t nrhs, double* X) {
  WAE_API_BEGIN
  CUDA_CHECK(cudaSetDevice(h->device));
  LuSolver& S = get_lu(h, lu_id);
  if (trans < 0 || trans > 2 || nrhs <= 0 || !X) WAE_THROW(WAE_E_INVALID, "bad solve arguments");
  int64_t n = S.sym.n;
  DevBuf<cplx>& dX = S.d_io;
  dX.upload((const cplx*)X, (size_t)n * nrhs, h->stream);
  PhaseTimer t(h, "solve");
  wae_lu_solve_device(h, S, trans, nrhs, dX.p, S.refine_steps);
  t.stop();
  CUDA_CHECK(cudaMemcpyAsync(X, dX.p, (size_t)n * nrhs * sizeof(cplx), cudaMemcpyDeviceToHost, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  WAE_API_END
}

// One shift-invert Arnoldi process of wae_eigs_si_pair as a state machine: input() names the basis vector the operator has to be applied
// to, consume() takes the result (left in `w` by the caller) through Gram-Schmidt, the Ritz test and -- after ncv steps without
// convergence -- the explicit restart.  Same arithmetic, tolerances and stopping rule as the loop of wae_eigs_si.
namespace {
struct ArnoldiProc {
  wae_ctx* h;
  int64_t n;
  int m, nev;
  cplx *V, *w, *c;  // basis n x (m+1), operator result (a column of the pair buffer), m+1 coefficients
  double* dots;     // 2 (m+2)
  std::vector<zc> H, Hs, theta, Y, hcol, h2;
  std::vector<int> order;
  int restart = 0, j = 0, jdim = 0;
  bool converged = false, done = false;
  double last_res = 1e300;  // worst relative residual of the wanted Ritz pairs of the LAST step: the ones result() returns
  static constexpr int max_restart = 15;
  static constexpr double tol = 1e-13;

  void dot(const cplx* Vb, int nv, const cplx* x, std::vector<zc>& out) {
    cudaStream_t st = h->stream;
    const int dot_blocks = (int)std::min<int64_t>((n + 255) / 256, 64);
    CUDA_CHECK(cudaMemsetAsync(dots, 0, 2 * nv * sizeof(double), st));
    multi_dot_kernel<<<dim3(dot_blocks, nv), 256, 0, st>>>(Vb, n, x, dots);
    h->launches++;
    out.resize(nv);
    CUDA_CHECK(cudaMemcpyAsync(out.data(), dots, 2 * nv * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
  }
  void upload_c(const std::vector<zc>& cc) { CUDA_CHECK(cudaMemcpyAsync(c, cc.data(), cc.size() * sizeof(zc), cudaMemcpyHostToDevice, h->stream)); }
  unsigned gb() const { return (unsigned)((n + 255) / 256); }

  void start(const double* v0_host) {  // w is free at this point: it stages the start vector
    cudaStream_t st = h->stream;
    CUDA_CHECK(cudaMemcpyAsync(w, v0_host, n * sizeof(cplx), cudaMemcpyHostToDevice, st));
    dot(w, 1, w, hcol);
    double nrm = std::sqrt(hcol[0].real());
    if (!(nrm > 0.0) || !std::isfinite(nrm)) WAE_THROW(WAE_E_INVALID, "start vector is zero or not finite");
    perturb_start_kernel<<<gb(), 256, 0, st>>>(w, n, 1e-6 * nrm / std::sqrt((double)n));
    h->launches++;
    dot(w, 1, w, hcol);
    nrm = std::sqrt(hcol[0].real());
    scale_copy_kernel<<<gb(), 256, 0, st>>>(w, n, 1.0 / nrm, V);
    h->launches++;
    H.assign((size_t)(m + 1) * m, 0.0);
  }
  const cplx* input() const { return V + (size_t)j * n; }
  void consume() {
    cudaStream_t st = h->stream;
    dot(V, j + 1, w, hcol);  // classical Gram-Schmidt with one reorthogonalisation
    upload_c(hcol);
    multi_axpy_kernel<<<gb(), 256, 0, st>>>(V, n, j + 1, c, w);
    dot(V, j + 1, w, h2);
    upload_c(h2);
    multi_axpy_kernel<<<gb(), 256, 0, st>>>(V, n, j + 1, c, w);
    h->launches += 2;
    for (int i = 0; i <= j; i++) H[(size_t)i * m + j] = hcol[i] + h2[i];
    std::vector<zc> nn;
    dot(w, 1, w, nn);
    const double beta = std::sqrt(std::max(0.0, nn[0].real()));
    if (!std::isfinite(beta)) WAE_THROW(WAE_E_SINGULAR, "non-finite Krylov vector (singular factorisation)");
    H[(size_t)(j + 1) * m + j] = beta;
    jdim = j + 1;
    Hs.assign((size_t)jdim * jdim, 0.0);
    for (int a = 0; a < jdim; a++)
      for (int b = 0; b < jdim; b++) Hs[(size_t)a * jdim + b] = H[(size_t)a * m + b];
    hessenberg_eig(jdim, Hs, theta, Y);
    order.resize(jdim);
    for (int i = 0; i < jdim; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return std::abs(theta[a]) > std::abs(theta[b]); });
    bool ok = jdim >= nev;
    const bool ok_all_evaluated = ok;
    double worst = 0;
    for (int i = 0; i < nev && jdim >= nev; i++) {  // all wanted pairs: `worst` is the residual of what result() would return
      const int q = order[i];
      const double rel = beta * std::abs(Y[(size_t)(jdim - 1) * jdim + q]) / std::max(std::abs(theta[q]), 1e-300);
      worst = std::max(worst, rel);
      if (!(rel <= tol)) ok = false;
    }
    last_res = (jdim >= nev && ok_all_evaluated) ? worst : 1e300;
    if (ok || beta <= 1e-300) {
      converged = done = true;
      return;
    }
    if (j + 1 < m) {
      scale_copy_kernel<<<gb(), 256, 0, st>>>(w, n, 1.0 / beta, V + (size_t)(j + 1) * n);
      h->launches++;
      j++;
      return;
    }
    if (restart == max_restart) {  // ncv steps of the last restart are used up
      done = true;
      return;
    }
    std::vector<zc> cc(jdim, 0.0);  // explicit restart with the sum of the wanted Ritz vectors
    for (int i = 0; i < nev; i++)
      for (int a = 0; a < jdim; a++) cc[a] += Y[(size_t)a * jdim + order[i]];
    upload_c(cc);
    lincomb_kernel<<<gb(), 256, 0, st>>>(V, n, jdim, c, w);
    dot(w, 1, w, nn);
    scale_copy_kernel<<<gb(), 256, 0, st>>>(w, n, 1.0 / std::sqrt(nn[0].real()), V);
    h->launches += 2;
    std::fill(H.begin(), H.end(), 0.0);
    restart++;
    j = 0;
  }
  // Ritz vectors -> host (through w); lambda = 1 / theta
  void result(double* lam, double* Vout) {
    cudaStream_t st = h->stream;
    // accepted without convergence to 1e-13 only if the pairs that are RETURNED (last step) are good to 1e-8 (ARPACK would throw)
    if (!converged && !(last_res <= 1e-8)) WAE_THROW(WAE_E_NOCONV, "Arnoldi did not converge (relative Ritz residual of the last step %.3e)", last_res);
    h->last_ms["eigs_residual"] = last_res < 1e300 ? last_res : -1.0;
    for (int i = 0; i < nev; i++) {
      const int q = order[i];
      std::vector<zc> cc(jdim);
      for (int a = 0; a < jdim; a++) cc[a] = Y[(size_t)a * jdim + q];
      upload_c(cc);
      lincomb_kernel<<<gb(), 256, 0, st>>>(V, n, jdim, c, w);
      h->launches++;
      CUDA_CHECK(cudaMemcpyAsync(Vout + 2 * (size_t)i * n, w, n * sizeof(cplx), cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaStreamSynchronize(st));  // w is reused by the next vector
      const zc l = 1.0 / theta[q];
      lam[2 * i] = l.real();
      lam[2 * i + 1] = l.imag();
    }
  }
};
}  // namespace

// The two shift-invert eigenproblems of one householder / mslp iteration -- eigs(A, M) and eigs(A', M') (Householder.jl:100-101,
// iterative_solvers.jl:132-133) -- advanced TOGETHER: every step applies A^{-1} M to the direct basis vector and A^{-H} M^H to the adjoint
// one as the two right-hand sides of one pass over the factor (wae_lu_solve_pair_device).  Needs the symmetric-mode factorisation; the
// caller falls back to two wae_eigs_si calls otherwise.  Opt-in from the host mirror (WAE_EIGS_PAIRED=1); written without GPU access.
int32_t wae_eigs_si_pair(wae_ctx* h, int32_t lu_id, int32_t fam_id, int32_t m_slot, int32_t nev, const double* v0, const double* v0_adj,
                         double* lam, double* Vout, double* lam_adj, double* Vout_adj, int32_t* n_solves) {
  WAE_API_BEGIN
  CUDA_CHECK(cudaSetDevice(h->device));
  LuSolver& S = get_lu(h, lu_id);
  Family& F = h->fam(fam_id);
  if (fam_id != S.fam) WAE_THROW(WAE_E_INVALID, "the LU belongs to another family");
  if (m_slot < 0 || m_slot >= WAE_FAMILY_SLOTS || !F.slot[m_slot].p) WAE_THROW(WAE_E_INVALID, "family slot %d (M) is empty", m_slot);
  if (!S.factored || !S.sym_mode) WAE_THROW(WAE_E_INVALID, "wae_eigs_si_pair needs a symmetric-mode factorisation (use wae_eigs_si twice)");
  const int64_t n = S.sym.n;
  if (nev < 1 || nev >= n || !v0 || !v0_adj || !lam || !Vout || !lam_adj || !Vout_adj) WAE_THROW(WAE_E_INVALID, "bad eigs arguments");
  const int m = (int)std::min<int64_t>(std::max(20, 2 * nev + 1), n);
  PhaseTimer timer(h, "eigs");
  S.d_arn_V.reserve((size_t)n * (m + 1));
  S.d_arn2_V.reserve((size_t)n * (m + 1));
  S.d_arn_c.reserve(m + 1);
  S.d_arn2_c.reserve(m + 1);
  S.d_arn_dots.reserve(2 * (m + 2));
  S.d_arn2_dots.reserve(2 * (m + 2));
  S.d_arn_pair.reserve((size_t)2 * n);
  ArnoldiProc P[2];
  for (int q = 0; q < 2; q++) {
    P[q].h = h;
    P[q].n = n;
    P[q].m = m;
    P[q].nev = nev;
    P[q].w = S.d_arn_pair.p + (size_t)q * n;
  }
  P[0].V = S.d_arn_V.p;  P[0].c = S.d_arn_c.p;  P[0].dots = S.d_arn_dots.p;
  P[1].V = S.d_arn2_V.p; P[1].c = S.d_arn2_c.p; P[1].dots = S.d_arn2_dots.p;
  P[0].start(v0);
  P[1].start(v0_adj);
  int solves = 0;
  while (!P[0].done || !P[1].done) {
    if (!P[0].done && !P[1].done) {
      wae_spmm_device(h, F, m_slot, 0, 1, P[0].input(), P[0].w);
      wae_spmm_device(h, F, m_slot, 2, 1, P[1].input(), P[1].w);
      wae_lu_solve_pair_device(h, S, S.d_arn_pair.p);
      solves += 2;
      P[0].consume();
      P[1].consume();
    } else {  // one process has converged: the other one finishes alone
      const int q = P[0].done ? 1 : 0;
      wae_spmm_device(h, F, m_slot, q ? 2 : 0, 1, P[q].input(), P[q].w);
      wae_lu_solve_device(h, S, q ? 2 : 0, 1, P[q].w, S.eigs_refine);
      solves++;
      P[q].consume();
    }
  }
  P[0].result(lam, Vout);
  P[1].result(lam_adj, Vout_adj);
  timer.stop();
  if (n_solves) *n_solves = solves;
  WAE_API_END
}

int32_t wae_eigs_si(wae_ctx* h, int32_t lu_id, int32_t fam_id, int32_t m_slot, int32_t trans, int32_t nev, const double* v0,
                    double* lam, double* Vout, int32_t* n_solves) {
  WAE_API_BEGIN
  CUDA_CHECK(cudaSetDevice(h->device));
  LuSolver& S = get_lu(h, lu_id);
  Family& F = h->fam(fam_id);
  if (fam_id != S.fam) WAE_THROW(WAE_E_INVALID, "the LU belongs to another family");
  if (m_slot < 0 || m_slot >= WAE_FAMILY_SLOTS || !F.slot[m_slot].p) WAE_THROW(WAE_E_INVALID, "family slot %d (M) is empty", m_slot);
  if (trans != 0 && trans != 2) WAE_THROW(WAE_E_INVALID, "trans must be 0 (A,M) or 2 (A',M')");
  const int64_t n = S.sym.n;
  if (nev < 1 || nev >= n || !v0 || !lam || !Vout) WAE_THROW(WAE_E_INVALID, "bad eigs arguments");
  const int m = (int)std::min<int64_t>(std::max(20, 2 * nev + 1), n);  // ncv as in Arpack.jl
  cudaStream_t st = h->stream;
  PhaseTimer timer(h, "eigs");
  struct { DevBuf<cplx>&V, &w, &t, &c; DevBuf<double>& dots; } W{S.d_arn_V, S.d_arn_w, S.d_arn_t, S.d_arn_c, S.d_arn_dots};
  W.V.reserve((size_t)n * (m + 1));
  W.w.reserve(n);
  W.t.reserve(n);
  W.c.reserve(m + 1);
  W.dots.reserve(2 * (m + 2));
  const unsigned gb = (unsigned)((n + 255) / 256);
  const int dot_blocks = (int)std::min<int64_t>((n + 255) / 256, 64);
  auto dots = [&](const cplx* Vb, int nv, const cplx* w, std::vector<zc>& out) {
    CUDA_CHECK(cudaMemsetAsync(W.dots.p, 0, 2 * nv * sizeof(double), st));
    multi_dot_kernel<<<dim3(dot_blocks, nv), 256, 0, st>>>(Vb, n, w, W.dots.p);
    h->launches++;
    out.resize(nv);
    CUDA_CHECK(cudaMemcpyAsync(out.data(), W.dots.p, 2 * nv * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
  };
  auto upload_c = [&](const std::vector<zc>& c) {
    CUDA_CHECK(cudaMemcpyAsync(W.c.p, c.data(), c.size() * sizeof(zc), cudaMemcpyHostToDevice, st));
  };
  int solves = 0;
  auto apply_op = [&](const cplx* x, cplx* y) {  // y = op(A)^{-1} op(M) x
    wae_spmm_device(h, F, m_slot, trans, 1, x, y);
    wae_lu_solve_device(h, S, trans, 1, y, S.eigs_refine);
    solves++;
  };
  // start vector
  CUDA_CHECK(cudaMemcpyAsync(W.w.p, v0, n * sizeof(cplx), cudaMemcpyHostToDevice, st));
  std::vector<zc> hcol, h2;
  dots(W.w.p, 1, W.w.p, hcol);
  double nrm = std::sqrt(hcol[0].real());
  if (!(nrm > 0.0) || !std::isfinite(nrm)) WAE_THROW(WAE_E_INVALID, "start vector is zero or not finite");
  perturb_start_kernel<<<gb, 256, 0, st>>>(W.w.p, n, 1e-6 * nrm / std::sqrt((double)n));
  h->launches++;
  dots(W.w.p, 1, W.w.p, hcol);
  nrm = std::sqrt(hcol[0].real());
  scale_copy_kernel<<<gb, 256, 0, st>>>(W.w.p, n, 1.0 / nrm, W.V.p);
  const double tol = 1e-13;
  const int max_restart = 15;
  std::vector<zc> H((size_t)(m + 1) * m, 0.0), Hs, theta, Y;
  std::vector<int> order;
  bool converged = false;
  double last_res = 1e300;  // worst relative residual of the wanted Ritz pairs of the last step (the ones that are returned)
  int jdim = 0;
  for (int restart = 0; restart <= max_restart && !converged; restart++) {
    std::fill(H.begin(), H.end(), 0.0);
    for (int j = 0; j < m; j++) {
      apply_op(W.V.p + (size_t)j * n, W.w.p);
      // classical Gram-Schmidt with one reorthogonalisation
      dots(W.V.p, j + 1, W.w.p, hcol);
      upload_c(hcol);
      multi_axpy_kernel<<<gb, 256, 0, st>>>(W.V.p, n, j + 1, W.c.p, W.w.p);
      dots(W.V.p, j + 1, W.w.p, h2);
      upload_c(h2);
      multi_axpy_kernel<<<gb, 256, 0, st>>>(W.V.p, n, j + 1, W.c.p, W.w.p);
      h->launches += 2;
      for (int i = 0; i <= j; i++) H[(size_t)i * m + j] = hcol[i] + h2[i];
      std::vector<zc> nn;
      dots(W.w.p, 1, W.w.p, nn);
      double beta = std::sqrt(std::max(0.0, nn[0].real()));
      if (!std::isfinite(beta)) WAE_THROW(WAE_E_SINGULAR, "non-finite Krylov vector (singular factorisation)");
      H[(size_t)(j + 1) * m + j] = beta;
      jdim = j + 1;
      // Ritz pairs of the leading jdim x jdim block
      Hs.assign((size_t)jdim * jdim, 0.0);
      for (int a = 0; a < jdim; a++)
        for (int b = 0; b < jdim; b++) Hs[(size_t)a * jdim + b] = H[(size_t)a * m + b];
      hessenberg_eig(jdim, Hs, theta, Y);
      order.resize(jdim);
      for (int i = 0; i < jdim; i++) order[i] = i;
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return std::abs(theta[a]) > std::abs(theta[b]); });
      bool ok = jdim >= nev;
      const bool ok_all_evaluated = ok;
      double worst = 0;
      for (int i = 0; i < nev && jdim >= nev; i++) {  // all wanted pairs: `worst` is the residual of what is returned
        int q = order[i];
        double res = beta * std::abs(Y[(size_t)(jdim - 1) * jdim + q]);
        double rel = res / std::max(std::abs(theta[q]), 1e-300);
        worst = std::max(worst, rel);
        if (!(rel <= tol)) ok = false;
      }
      last_res = (jdim >= nev && ok_all_evaluated) ? worst : 1e300;
      if (ok || beta <= 1e-300) {
        converged = true;
        break;
      }
      scale_copy_kernel<<<gb, 256, 0, st>>>(W.w.p, n, 1.0 / beta, W.V.p + (size_t)(j + 1) * n);
      h->launches++;
    }
    if (!converged && restart < max_restart) {
      // explicit restart with the sum of the wanted Ritz vectors
      std::vector<zc> c(jdim, 0.0);
      for (int i = 0; i < nev; i++)
        for (int a = 0; a < jdim; a++) c[a] += Y[(size_t)a * jdim + order[i]];
      upload_c(c);
      lincomb_kernel<<<gb, 256, 0, st>>>(W.V.p, n, jdim, W.c.p, W.w.p);
      std::vector<zc> nn;
      dots(W.w.p, 1, W.w.p, nn);
      scale_copy_kernel<<<gb, 256, 0, st>>>(W.w.p, n, 1.0 / std::sqrt(nn[0].real()), W.V.p);
      h->launches += 2;
    }
  }
  if (!converged && !(last_res <= 1e-8)) WAE_THROW(WAE_E_NOCONV, "Arnoldi did not converge (relative Ritz residual of the last step %.3e)", last_res);
  h->last_ms["eigs_residual"] = last_res < 1e300 ? last_res : -1.0;
  // Ritz vectors -> host; lambda = 1/theta
  for (int i = 0; i < nev; i++) {
    int q = order[i];
    std::vector<zc> c(jdim);
    for (int a = 0; a < jdim; a++) c[a] = Y[(size_t)a * jdim + q];
    upload_c(c);
    lincomb_kernel<<<gb, 256, 0, st>>>(W.V.p, n, jdim, W.c.p, W.t.p);
    h->launches++;
    CUDA_CHECK(cudaMemcpyAsync(Vout + 2 * (size_t)i * n, W.t.p, n * sizeof(cplx), cudaMemcpyDeviceToHost, st));
    zc l = 1.0 / theta[q];
    lam[2 * i] = l.real();
    lam[2 * i + 1] = l.imag();
  }
  timer.stop();
  CUDA_CHECK(cudaStreamSynchronize(st));
  if (n_solves) *n_solves = solves;
  WAE_API_END
}

// The node loop of one device: combine -> factorise -> solve l right-hand sides -> moments (in the solve's last kernel).  Throws.
static void beyn_moments_device(wae_ctx* h, int32_t fam_id, int32_t lu_id, int32_t n_nodes, const double* z, const double* w,
                                const double* coeffs, int32_t l, int32_t n_mom, const double* V, cplx* A_out) {
  CUDA_CHECK(cudaSetDevice(h->device));
  LuSolver& S = get_lu(h, lu_id);
  Family& F = h->fam(fam_id);
  if (fam_id != S.fam) WAE_THROW(WAE_E_INVALID, "the LU belongs to another family");
  const int64_t n = S.sym.n;
  if (n_nodes < 0 || (n_nodes && (!z || !w || !coeffs)) || l < 1 || l > n || n_mom < 1 || !A_out) WAE_THROW(WAE_E_INVALID, "bad Beyn arguments");
  cudaStream_t st = h->stream;
  DevBuf<cplx>& X = S.d_io;
  X.reserve((size_t)n * l);
  const int slot = WAE_FAMILY_SLOTS - 1;
  if (V) S.d_arn_V.upload((const cplx*)V, (size_t)n * l, st);  // the probing matrix is kept in the (otherwise idle) Arnoldi workspace
  double t_fac = 0, t_sol = 0;
  for (int j = 0; j < n_nodes; j++) {
    wae_combine_device(h, F, coeffs + 2 * (size_t)j * F.n_terms, slot);
    {
      PhaseTimer t(h, "factor");
      factor_slot(h, S, F, slot);
      t.stop();
      t_fac += h->last_ms["factor"];
    }
    PhaseTimer t(h, "solve");
    if (V)
      CUDA_CHECK(cudaMemcpyAsync(X.p, S.d_arn_V.p, (size_t)n * l * sizeof(cplx), cudaMemcpyDeviceToDevice, st));
    else
      identity_cols_kernel<<<(unsigned)(((size_t)n * l + 255) / 256), 256, 0, st>>>(n, l, X.p);
    const LuMomentEpilogue ep{n_mom, make_double2(w[2 * j], w[2 * j + 1]), make_double2(z[2 * j], z[2 * j + 1]), A_out};
    wae_lu_solve_device(h, S, 0, l, X.p, S.refine_steps, &ep);
    h->launches += 1;
    t.stop();
    t_sol += h->last_ms["solve"];
  }
  h->last_ms["beyn_factor_total"] = t_fac;
  h->last_ms["beyn_solve_total"] = t_sol;
  CUDA_CHECK(cudaStreamSynchronize(st));
}

int32_t wae_beyn_moments(wae_ctx* h, int32_t fam_id, int32_t lu_id, int32_t n_nodes, const double* z, const double* w,
                         const double* coeffs, int32_t l, int32_t n_mom, const double* V, void* A_out) {
  WAE_API_BEGIN
  beyn_moments_device(h, fam_id, lu_id, n_nodes, z, w, coeffs, l, n_mom, V, (cplx*)A_out);
  WAE_API_END
}

// ---- all GPUs of the box from ONE host process (the reference's beyn is one process, one loop over all nodes: beyn.jl:34-74,112-138)
// NCCL is loaded at run time (libnccl.so.2, or the library named by WAE_NCCL_LIB): libwae_b200.so has no link-time dependency on it, and
// a single-GPU caller never touches it.
namespace {
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::map<std::vector<int>, std::vector<ncclComm_t>> comms;  // one communicator set per device list, kept for the process lifetime
  std::mutex mu;
};
NcclApi& nccl_api() {
  static NcclApi api;
  std::lock_guard<std::mutex> lock(api.mu);
  if (api.lib) return api;
  const char* names[] = {getenv("WAE_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* nm : names)
    if (nm && *nm && (api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL))) break;
  if (!api.lib) WAE_THROW(WAE_E_CUDA, "NCCL not found (libnccl.so.2; set WAE_NCCL_LIB): %s", dlerror());
  auto sym = [&](const char* nm) {
    void* f = dlsym(api.lib, nm);
    if (!f) WAE_THROW(WAE_E_CUDA, "NCCL symbol %s missing", nm);
    return f;
  };
  api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
  api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
  api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
  api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
  api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
  api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  return api;
}
#define NCCL_CHECK(api, x)                                                                                   \
  do {                                                                                                       \
    ncclResult_t _r = (x);                                                                                   \
    if (_r != ncclSuccess) WAE_THROW(WAE_E_CUDA, "%s failed: %s", #x, (api).GetErrorString(_r));             \
  } while (0)
}  // namespace

int32_t wae_beyn_moments_multi(int32_t n_ctx, wae_ctx* const* hs, const int32_t* fam_ids, const int32_t* lu_ids, int32_t n_nodes,
                               const double* z, const double* w, const double* coeffs, int32_t l, int32_t n_mom, const double* V,
                               double* A_out) {
  if (n_ctx < 1 || !hs || !hs[0]) return WAE_E_INVALID;
  wae_ctx* h = hs[0];
  WAE_API_BEGIN
  if (!fam_ids || !lu_ids || n_nodes < 0 || !A_out) WAE_THROW(WAE_E_INVALID, "bad Beyn arguments");
  int64_t n = -1;
  int n_terms = -1;
  std::vector<int> devs(n_ctx);
  for (int r = 0; r < n_ctx; r++) {
    if (!hs[r]) WAE_THROW(WAE_E_INVALID, "context %d is NULL", r);
    LuSolver& S = get_lu(hs[r], lu_ids[r]);
    Family& F = hs[r]->fam(fam_ids[r]);
    if (r == 0) n = S.sym.n, n_terms = F.n_terms;
    if (S.sym.n != n || F.n_terms != n_terms) WAE_THROW(WAE_E_INVALID, "context %d holds a different family (dimension / number of terms)", r);
    devs[r] = hs[r]->device;
    for (int q = 0; q < r; q++)
      if (devs[q] == devs[r]) WAE_THROW(WAE_E_INVALID, "contexts %d and %d sit on the same device", q, r);
  }
  const size_t total = (size_t)n * l * n_mom;
  // quadrature node j -> context j mod n_ctx, every context accumulates its own partial moments
  std::vector<std::vector<double>> zr(n_ctx), wr(n_ctx), cr(n_ctx);
  for (int j = 0; j < n_nodes; j++) {
    const int r = j % n_ctx;
    zr[r].insert(zr[r].end(), z + 2 * j, z + 2 * j + 2);
    wr[r].insert(wr[r].end(), w + 2 * j, w + 2 * j + 2);
    cr[r].insert(cr[r].end(), coeffs + 2 * (size_t)j * n_terms, coeffs + 2 * (size_t)(j + 1) * n_terms);
  }
  std::vector<DevBuf<cplx>> A(n_ctx);
  std::vector<WaeError> errs(n_ctx, WaeError{WAE_OK, ""});
  {
    std::vector<std::thread> th;
    for (int r = 0; r < n_ctx; r++)
      th.emplace_back([&, r]() {
        try {
          CUDA_CHECK(cudaSetDevice(hs[r]->device));
          A[r].alloc(total);
          CUDA_CHECK(cudaMemsetAsync(A[r].p, 0, total * sizeof(cplx), hs[r]->stream));
          beyn_moments_device(hs[r], fam_ids[r], lu_ids[r], (int32_t)(zr[r].size() / 2), zr[r].data(), wr[r].data(), cr[r].data(), l, n_mom, V, A[r].p);
        } catch (const WaeError& e) {
          errs[r] = e;
        } catch (const std::exception& e) {
          errs[r] = WaeError{WAE_E_INVALID, e.what()};
        }
      });
    for (auto& t : th) t.join();
  }
  for (int r = 0; r < n_ctx; r++)
    if (errs[r].code != WAE_OK) {
      hs[r]->err = errs[r].msg;
      throw WaeError{errs[r].code, "device " + std::to_string(devs[r]) + ": " + errs[r].msg};
    }
  if (n_ctx > 1) {  // one all-reduce of the 2K x l x d complex moment tensor over NVLink
    NcclApi& api = nccl_api();
    std::vector<ncclComm_t>* comm;
    {
      std::lock_guard<std::mutex> lock(api.mu);
      auto it = api.comms.find(devs);
      if (it == api.comms.end()) {
        std::vector<ncclComm_t> c(n_ctx);
        NCCL_CHECK(api, api.CommInitAll(c.data(), n_ctx, devs.data()));
        it = api.comms.emplace(devs, c).first;
      }
      comm = &it->second;
    }
    NCCL_CHECK(api, api.GroupStart());
    for (int r = 0; r < n_ctx; r++)
      NCCL_CHECK(api, api.AllReduce(A[r].p, A[r].p, 2 * total, ncclDouble, ncclSum, (*comm)[r], hs[r]->stream));
    NCCL_CHECK(api, api.GroupEnd());
    for (int r = 0; r < n_ctx; r++) {
      CUDA_CHECK(cudaSetDevice(hs[r]->device));
      CUDA_CHECK(cudaStreamSynchronize(hs[r]->stream));
    }
  }
  CUDA_CHECK(cudaSetDevice(h->device));
  CUDA_CHECK(cudaMemcpyAsync(A_out, A[0].p, total * sizeof(cplx), cudaMemcpyDeviceToHost, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  for (int r = 0; r < n_ctx; r++) {  // device buffers are released on their own device
    cudaSetDevice(hs[r]->device);
    A[r].release();
  }
  cudaSetDevice(h->device);
  WAE_API_END
}

}  // extern "C"
