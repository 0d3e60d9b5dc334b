// Element kernels: P1/P2 tetrahedral mass/stiffness, boundary-triangle mass, flame source.
// One element per thread; closed-form element matrices (fem_gen.h, exact integrals of the
// Lagrange bases == the reference's tables src/FEM/FEM.jl:435-450,704-738,1745-1874).
//
// Two generations of the tetrahedral M/K kernel live here:
//   * assemble_tet_atomic : scatter through a precomputed slot map with fp64 RED atomics
//     (any operator, any c; used for small sub-domain operators and as cross-check)
//   * assemble_tet_gather : owner-computes patches, element matrices staged in shared memory,
//     every nonzero written exactly once in a fixed summation order (the production path
//     for the big M+K pass; see assembly_symbolic.cpp for the gather program).
#include <cuda_runtime.h>

#include "fem_gen.h"
#include "wae_internal.h"

// ---- geometry -----------------------------------------------------------------------
// CooTrafo (FEM.jl:2-21): J[:,k] = X_k - X_4, inverse rows = grad(lambda_k), |det J|.
struct TetGeom {
  double G[4][3];  // gradients of the four barycentric coordinates
  double adet;     // |det J|
};

__device__ __forceinline__ void tet_geom(const double* __restrict__ xyz, const uint32_t v[4], TetGeom& t) {
  double x3 = xyz[3 * (size_t)v[3]], y3 = xyz[3 * (size_t)v[3] + 1], z3 = xyz[3 * (size_t)v[3] + 2];
  double a[3][3];  // a[r][k] = component r of edge k
#pragma unroll
  for (int k = 0; k < 3; k++) {
    a[0][k] = xyz[3 * (size_t)v[k]] - x3;
    a[1][k] = xyz[3 * (size_t)v[k] + 1] - y3;
    a[2][k] = xyz[3 * (size_t)v[k] + 2] - z3;
  }
  // cofactors
  double c00 = a[1][1] * a[2][2] - a[1][2] * a[2][1];
  double c01 = a[1][2] * a[2][0] - a[1][0] * a[2][2];
  double c02 = a[1][0] * a[2][1] - a[1][1] * a[2][0];
  double det = a[0][0] * c00 + a[0][1] * c01 + a[0][2] * c02;
  double id = 1.0 / det;
  // inverse = adj/det ; row k of the inverse is grad(lambda_k)
  t.G[0][0] = c00 * id;
  t.G[0][1] = (a[0][2] * a[2][1] - a[0][1] * a[2][2]) * id;
  t.G[0][2] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) * id;
  t.G[1][0] = c01 * id;
  t.G[1][1] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) * id;
  t.G[1][2] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) * id;
  t.G[2][0] = c02 * id;
  t.G[2][1] = (a[0][1] * a[2][0] - a[0][0] * a[2][1]) * id;
  t.G[2][2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) * id;
#pragma unroll
  for (int d = 0; d < 3; d++) t.G[3][d] = -(t.G[0][d] + t.G[1][d] + t.G[2][d]);
  t.adet = fabs(det);
}

// g[q] = grad(l_a).grad(l_b) * s for the 10 pairs a<=b
__device__ __forceinline__ void tet_gram(const TetGeom& t, double s, double g[10]) {
  int q = 0;
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = a; b < 4; b++) g[q++] = (t.G[a][0] * t.G[b][0] + t.G[a][1] * t.G[b][1] + t.G[a][2] * t.G[b][2]) * s;
}

template <int NLOC>
struct Tab;
template <>
struct Tab<4> {
  __device__ static const double* mass() { return WAE_P1_TET_MASS; }
  __device__ static const double* src() { return WAE_P1_TET_SRC; }
  __device__ static const double* stiffcc() { return WAE_P1_TET_STIFFCC; }
  __device__ static void stiff(const double* g, double* K) { wae_p1_tet_stiff(g, K); }
};
template <>
struct Tab<10> {
  __device__ static const double* mass() { return WAE_P2_TET_MASS; }
  __device__ static const double* src() { return WAE_P2_TET_SRC; }
  __device__ static const double* stiffcc() { return WAE_P2_TET_STIFFCC; }
  __device__ static void stiff(const double* g, double* K) { wae_p2_tet_stiff(g, K); }
};

// ---- generation 1: atomic scatter -------------------------------------------------------
// mode bit 0: mass -> out_m, bit 1: stiffness -> out_k
template <int NLOC>
__global__ void __launch_bounds__(128) assemble_tet_atomic(const double* __restrict__ xyz, const uint32_t* __restrict__ conn,
                                                           const int32_t* __restrict__ elems, int64_t n_elem,
                                                           const int32_t* __restrict__ slotmap, const double* __restrict__ c,
                                                           int c_per_elem, int mode, double mass_scale,
                                                           double* __restrict__ out_m, double* __restrict__ out_k) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_elem) return;
  const uint32_t* d = conn + (size_t)elems[e] * NLOC;
  uint32_t v[4] = {d[0], d[1], d[2], d[3]};
  TetGeom t;
  tet_geom(xyz, v, t);
  const int32_t* sm = slotmap + (size_t)e * NLOC * NLOC;
  if (mode & 1) {
    const double* Tm = Tab<NLOC>::mass();
    double s = t.adet * mass_scale;
#pragma unroll 4
    for (int k = 0; k < NLOC * NLOC; k++) atomicAdd(out_m + sm[k], s * Tm[k]);
  }
  if (mode & 2) {
    double K[NLOC * NLOC];
    if (c_per_elem == 1) {
      double cc = c[e];
      double g[10];
      tet_gram(t, -cc * cc * t.adet, g);
      Tab<NLOC>::stiff(g, K);
    } else {  // linear c: K_ab = - sum_q g_q sum_{k<=l} c_k c_l T[a][b][q][kl]   (FEM.jl:2283-2424)
      double g[10], cc[10];
      tet_gram(t, -t.adet, g);
      const double* cv = c + 4 * e;
      int q = 0;
      for (int k = 0; k < 4; k++)
        for (int l = k; l < 4; l++) cc[q++] = cv[k] * cv[l];
      const double* T = Tab<NLOC>::stiffcc();
      for (int ab = 0; ab < NLOC * NLOC; ab++) {
        double acc = 0;
        for (int qq = 0; qq < 10; qq++) {
          double w = 0;
          for (int kl = 0; kl < 10; kl++) w += cc[kl] * T[(ab * 10 + qq) * 10 + kl];
          acc += g[qq] * w;
        }
        K[ab] = acc;
      }
    }
#pragma unroll 4
    for (int k = 0; k < NLOC * NLOC; k++) atomicAdd(out_k + sm[k], K[k]);
  }
}

// boundary triangles: C = -i * c * int phi_i phi_j   (Helmholtz.jl:151-171,446-463; FEM.jl:435-450,469-525)
template <int NLOC3>
__global__ void __launch_bounds__(128) assemble_tri_atomic(const double* __restrict__ xyz, const uint32_t* __restrict__ conn,
                                                           const int32_t* __restrict__ elems, int64_t n_elem,
                                                           const int32_t* __restrict__ slotmap, const double* __restrict__ c,
                                                           int c_per_elem, double scale, double* __restrict__ out /* complex */) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_elem) return;
  const uint32_t* d = conn + (size_t)elems[e] * NLOC3;
  double X[3][3];
  for (int k = 0; k < 3; k++)
    for (int r = 0; r < 3; r++) X[k][r] = xyz[3 * (size_t)d[k] + r];
  double e1[3], e2[3];
  for (int r = 0; r < 3; r++) {
    e1[r] = X[0][r] - X[2][r];
    e2[r] = X[1][r] - X[2][r];
  }
  double nx = e1[1] * e2[2] - e1[2] * e2[1], ny = e1[2] * e2[0] - e1[0] * e2[2], nz = e1[0] * e2[1] - e1[1] * e2[0];
  double adet = sqrt(nx * nx + ny * ny + nz * nz);  // |det [e1 e2 n^]| = 2 * area
  const double* Tm = NLOC3 == 3 ? WAE_P1_TRI_MASS : WAE_P2_TRI_MASS;
  const double* Tc = NLOC3 == 3 ? WAE_P1_TRI_MASSC : WAE_P2_TRI_MASSC;
  const int32_t* sm = slotmap + (size_t)e * NLOC3 * NLOC3;
  for (int k = 0; k < NLOC3 * NLOC3; k++) {
    double m;
    if (c_per_elem == 1)
      m = c[e] * Tm[k];
    else
      m = c[3 * e] * Tc[3 * k] + c[3 * e + 1] * Tc[3 * k + 1] + c[3 * e + 2] * Tc[3 * k + 2];
    // value = -i * m * adet * scale  -> imaginary part only
    atomicAdd(out + 2 * (size_t)sm[k] + 1, -m * adet * scale);
  }
}

// ---- flame: S_i = sum_t |det_t| src[loc] over flame tets, G_j on the reference tet ------
template <int NLOC>
__global__ void flame_src_kernel(const double* __restrict__ xyz, const uint32_t* __restrict__ conn,
                                 const int32_t* __restrict__ tets, int64_t n, const int32_t* __restrict__ rowpos,
                                 double* __restrict__ S, double* __restrict__ vol) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const uint32_t* d = conn + (size_t)tets[e] * NLOC;
  uint32_t v[4] = {d[0], d[1], d[2], d[3]};
  TetGeom t;
  tet_geom(xyz, v, t);
  const double* src = Tab<NLOC>::src();
  for (int k = 0; k < NLOC; k++) atomicAdd(S + rowpos[e * NLOC + k], t.adet * src[k]);
  atomicAdd(vol, t.adet / 6.0);
}

// grad phi_j(x_ref).n_ref for the reference tet (FEM.jl:2442-2484), times `fac`; one thread.
template <int NLOC>
__global__ void flame_grad_kernel(const double* __restrict__ xyz, const uint32_t* __restrict__ conn, int64_t ref_tet,
                                  double xr, double yr, double zr, double n0, double n1, double n2, double fac,
                                  double* __restrict__ Gout) {
  if (threadIdx.x || blockIdx.x) return;
  const uint32_t* d = conn + (size_t)ref_tet * NLOC;
  uint32_t v[4] = {d[0], d[1], d[2], d[3]};
  TetGeom t;
  tet_geom(xyz, v, t);
  double gn[4], lam[4];
  double dx = xr - xyz[3 * (size_t)v[3]], dy = yr - xyz[3 * (size_t)v[3] + 1], dz = zr - xyz[3 * (size_t)v[3] + 2];
  for (int a = 0; a < 4; a++) gn[a] = t.G[a][0] * n0 + t.G[a][1] * n1 + t.G[a][2] * n2;
  for (int a = 0; a < 3; a++) lam[a] = t.G[a][0] * dx + t.G[a][1] * dy + t.G[a][2] * dz;
  lam[3] = 1.0 - lam[0] - lam[1] - lam[2];
  if (NLOC == 4) {
    for (int a = 0; a < 4; a++) Gout[a] = fac * gn[a];
  } else {
    for (int a = 0; a < 4; a++) Gout[a] = fac * (4.0 * lam[a] - 1.0) * gn[a];
    int q = 4;
    for (int a = 0; a < 4; a++)
      for (int b = a + 1; b < 4; b++) Gout[q++] = fac * 4.0 * (lam[a] * gn[b] + lam[b] * gn[a]);
  }
}

// Q[r, c] = S[r] * G[colsrc[c]]  (dense block, column-major == CSC order of the flame pattern)
__global__ void flame_outer_kernel(const double* __restrict__ S, const double* __restrict__ G,
                                   const int32_t* __restrict__ colsrc, int64_t nrows, int ncols, double* __restrict__ out) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nrows * ncols) return;
  out[k] = S[k % nrows] * G[colsrc[k / nrows]];
}

// host-side wrappers ---------------------------------------------------------------------
void wae_launch_assemble_atomic(wae_ctx* h, Pattern& P, int kind, const double* d_c, int c_per_elem, double scale,
                                double* d_out_a, double* d_out_b) {
  int64_t ne = (int64_t)P.elems.size();
  if (ne == 0) return;
  int threads = 128;
  unsigned blocks = (unsigned)((ne + threads - 1) / threads);
  if (P.elem_kind == 3) {
    int mode = kind;  // 1 mass, 2 stiff, 3 both
    if (h->nloc == 4)
      assemble_tet_atomic<4><<<blocks, threads, 0, h->stream>>>(h->d_xyz.p, h->d_tets.p, P.d_elems.p, ne, P.d_slotmap.p, d_c,
                                                                c_per_elem, mode, scale, d_out_a, d_out_b);
    else
      assemble_tet_atomic<10><<<blocks, threads, 0, h->stream>>>(h->d_xyz.p, h->d_tets.p, P.d_elems.p, ne, P.d_slotmap.p, d_c,
                                                                 c_per_elem, mode, scale, d_out_a, d_out_b);
  } else {
    if (h->nloc3 == 3)
      assemble_tri_atomic<3><<<blocks, threads, 0, h->stream>>>(h->d_xyz.p, h->d_tris.p, P.d_elems.p, ne, P.d_slotmap.p, d_c,
                                                                c_per_elem, scale, d_out_a);
    else
      assemble_tri_atomic<6><<<blocks, threads, 0, h->stream>>>(h->d_xyz.p, h->d_tris.p, P.d_elems.p, ne, P.d_slotmap.p, d_c,
                                                                c_per_elem, scale, d_out_a);
  }
  h->launches++;
  CUDA_CHECK(cudaGetLastError());
}

void wae_launch_flame(wae_ctx* h, const int32_t* d_tets, int64_t n, const int32_t* d_rowpos, double* d_S, double* d_vol,
                      int64_t ref_tet, const double* x_ref, const double* n_ref, double fac, double* d_G,
                      const int32_t* d_colsrc, int64_t nrows, int ncols, double* d_out) {
  unsigned blocks = (unsigned)((n + 127) / 128);
  if (h->nloc == 4) {
    flame_src_kernel<4><<<blocks, 128, 0, h->stream>>>(h->d_xyz.p, h->d_tets.p, d_tets, n, d_rowpos, d_S, d_vol);
    flame_grad_kernel<4><<<1, 32, 0, h->stream>>>(h->d_xyz.p, h->d_tets.p, ref_tet, x_ref[0], x_ref[1], x_ref[2], n_ref[0],
                                                  n_ref[1], n_ref[2], fac, d_G);
  } else {
    flame_src_kernel<10><<<blocks, 128, 0, h->stream>>>(h->d_xyz.p, h->d_tets.p, d_tets, n, d_rowpos, d_S, d_vol);
    flame_grad_kernel<10><<<1, 32, 0, h->stream>>>(h->d_xyz.p, h->d_tets.p, ref_tet, x_ref[0], x_ref[1], x_ref[2], n_ref[0],
                                                   n_ref[1], n_ref[2], fac, d_G);
  }
  int64_t tot = nrows * ncols;
  flame_outer_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(d_S, d_G, d_colsrc, nrows, ncols, d_out);
  h->launches += 3;
  CUDA_CHECK(cudaGetLastError());
}

// ---- generation 2: owner-computes gather ------------------------------------------------
// One CTA per patch.  Phase A: one staged element per thread -> packed symmetric stiffness
// (NSYM entries, already scaled by -c^2 |det|) and |det| into shared memory.  Phase B: one warp
// per owned column; lanes walk the column's nonzeros, a warp scan of the per-slot source counts
// gives each lane its offset into the packed source list; K = sum of staged entries,
// M = table(sym) * sum of |det| (the mass entry of a DOF pair depends only on the pair's type,
// which is the same in every element that contains the pair).  Every output is written once.
template <int NLOC>
struct SymTab;
template <>
struct SymTab<4> {
  static constexpr int NSYM = 10;
  __device__ static const double* mass() { return WAE_P1_TET_MASS_SYM; }
  __device__ static void stiff(const double* g, double* K) { wae_p1_tet_stiff_sym(g, K); }
};
template <>
struct SymTab<10> {
  static constexpr int NSYM = 55;
  __device__ static const double* mass() { return WAE_P2_TET_MASS_SYM; }
  __device__ static void stiff(const double* g, double* K) { wae_p2_tet_stiff_sym(g, K); }
};

// Shared-memory staging layout: 64 doubles per staged element (NSYM stiffness entries, |det| at logical index 63), XOR-swizzled
// with the element index so that lanes reading the same logical entry of different elements hit different banks:
//   physical index of (element t, entry e) = t*64 + (e ^ (t & 63)).   A source code is t*64 + e, so the address of the
// stiffness entry is one LOP3 away from the code and the |det| entry is (code | 63) ^ (t & 63).  Element index WAE_GATHER_PAD
// is a block of zeros: padding sources point there, which removes every branch from the inner loop.
__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));  // not volatile: smem is read-only after the barrier
  return v;
}

template <int NLOC>
__global__ void __launch_bounds__(512, 2) assemble_tet_gather(
    const double* __restrict__ xyz, const uint32_t* __restrict__ conn, const int32_t* __restrict__ elems,
    const double* __restrict__ c, const int64_t* __restrict__ patch_tet_ptr, const int32_t* __restrict__ patch_tets,
    const int64_t* __restrict__ patch_grp_ptr, const int64_t* __restrict__ patch_src_ptr, const uint32_t* __restrict__ grp,
    const int32_t* __restrict__ out_idx, const uint16_t* __restrict__ src, int pad_elem, double mass_scale, double* __restrict__ out_m,
    double* __restrict__ out_k, int dbg) {
  constexpr int NSYM = SymTab<NLOC>::NSYM;
  constexpr int SL = NLOC == 4 ? 4 : 6;     // log2 of the staging stride (16 / 64 doubles per element)
  constexpr unsigned SM_ = (1u << SL) - 1;  // mask of the entry index; |det| lives at logical entry SM_
  extern __shared__ double sm[];
  __shared__ double s_mass[NSYM];
  __shared__ int s_next;
  const int p = blockIdx.x;
  const int64_t t0 = patch_tet_ptr[p];
  const int nt = (int)(patch_tet_ptr[p + 1] - t0);
  if (threadIdx.x < NSYM) s_mass[threadIdx.x] = SymTab<NLOC>::mass()[threadIdx.x] * mass_scale;
  if (threadIdx.x == 0) s_next = blockDim.x >> 5;
  if (threadIdx.x <= SM_) sm[(pad_elem << SL) + threadIdx.x] = 0.0;
  // phase A: one staged element per thread
  for (int t = threadIdx.x; t < nt && dbg != 2; t += blockDim.x) {
    int32_t e = patch_tets[t0 + t];
    const uint32_t* d = conn + (size_t)elems[e] * NLOC;
    uint32_t v[4] = {d[0], d[1], d[2], d[3]};
    TetGeom tg;
    tet_geom(xyz, v, tg);
    double* dst = sm + ((size_t)t << SL);
    const int sw = t & SM_;
    if (out_k) {
      double cc = c[e];
      double g[10];
      tet_gram(tg, -cc * cc * tg.adet, g);
      double K[NSYM];
      SymTab<NLOC>::stiff(g, K);
#pragma unroll
      for (int k = 0; k < NSYM; k++) dst[k ^ sw] = K[k];
    }
    dst[SM_ ^ sw] = tg.adet;
  }
  __syncthreads();
  if (dbg == 1) return;
  // phase B: one group of WAE_GATHER_GROUP owned nonzeros per warp iteration (GS/32 per lane); the
  // nonzeros of a patch are sorted by source count, so all lanes run (almost) the same trip count, branch-free
  const int lane = threadIdx.x & 31;
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm);
  const int g0 = (int)patch_grp_ptr[p], ng = (int)(patch_grp_ptr[p + 1] - patch_grp_ptr[p]);
  const uint16_t* psrc = src + patch_src_ptr[p] + lane;
  constexpr int GS = WAE_GATHER_GROUP;
  int g = threadIdx.x >> 5;
  while (g < ng) {
    const uint32_t hdr = grp[g0 + g];
    const int niter = hdr & 255;
    const uint16_t* sp = psrc + (hdr >> 8);
    const int32_t* op = out_idx + (size_t)(g0 + g) * GS + lane;
    constexpr int NJ = GS / 32;  // nonzeros per lane and group
    int oi[NJ];
#pragma unroll
    for (int j = 0; j < NJ; j++) oi[j] = op[j * 32];
    double ak[NJ], ad[NJ];
    unsigned first[NJ];
#pragma unroll
    for (int j = 0; j < NJ; j++) {
      ak[j] = ad[j] = 0.0;
      first[j] = niter ? sp[j * 32] : 0u;
    }
#pragma unroll 4
    for (int k = 0; k < niter; k++) {
      unsigned cd[NJ];
#pragma unroll
      for (int j = 0; j < NJ; j++) cd[j] = sp[k * GS + j * 32];
#pragma unroll
      for (int j = 0; j < NJ; j++) {
        const unsigned sw = (cd[j] >> SL) & SM_;
        if (out_k) ak[j] += lds_f64(sbase + ((cd[j] ^ sw) << 3));
        ad[j] += lds_f64(sbase + (((cd[j] | SM_) ^ sw) << 3));
      }
    }
    if (dbg == 3) {  // timing experiment: no scattered stores (one dependent store per lane and group keeps the work alive)
      double z = 0.0;
      for (int j = 0; j < NJ; j++) z += ak[j] + ad[j];
      if (z == 1.2345e-300 && out_k) out_k[oi[0] & 1023] = z;
    } else
#pragma unroll
    for (int j = 0; j < NJ; j++)
      if (oi[j] >= 0) {
        if (out_k) out_k[oi[j]] = ak[j];
        if (out_m) out_m[oi[j]] = s_mass[first[j] & SM_] * ad[j];
      }
    if (lane == 0) g = atomicAdd(&s_next, 1);
    g = __shfl_sync(0xffffffffu, g, 0);
  }
}

void wae_launch_assemble_gather(wae_ctx* h, Pattern& P, const double* d_c, double* d_mass, double* d_stiff, double mass_scale) {
  auto& G = P.gather;
  if (!G.built || G.n_patch == 0) return;
  const int nsym = h->nloc == 4 ? 10 : 55;
  size_t smem = (size_t)(G.max_tets + 1) * (h->nloc == 4 ? 16 : 64) * sizeof(double);  // +1: the zero block the padding sources point to
  static bool attr_set[2] = {false, false};
  const int dbg = getenv("WAE_GATHER_DBG") ? atoi(getenv("WAE_GATHER_DBG")) : 0;
  if (h->nloc == 4) {
    if (!attr_set[0]) {
      CUDA_CHECK(cudaFuncSetAttribute(assemble_tet_gather<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
      attr_set[0] = true;
    }
    assemble_tet_gather<4><<<G.n_patch, 512, smem, h->stream>>>(h->d_xyz.p, h->d_tets.p, P.d_elems.p, d_c, G.d_patch_tet_ptr.p, G.d_patch_tets.p,
                                                                G.d_patch_grp_ptr.p, G.d_patch_src_ptr.p, G.d_grp.p, G.d_out_idx.p, G.d_src.p,
                                                                G.max_tets, mass_scale, d_mass, d_stiff, dbg);
  } else {
    if (!attr_set[1]) {
      CUDA_CHECK(cudaFuncSetAttribute(assemble_tet_gather<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
      attr_set[1] = true;
    }
    assemble_tet_gather<10><<<G.n_patch, 512, smem, h->stream>>>(h->d_xyz.p, h->d_tets.p, P.d_elems.p, d_c, G.d_patch_tet_ptr.p, G.d_patch_tets.p,
                                                                 G.d_patch_grp_ptr.p, G.d_patch_src_ptr.p, G.d_grp.p, G.d_out_idx.p, G.d_src.p,
                                                                 G.max_tets, mass_scale, d_mass, d_stiff, dbg);
  }
  h->launches++;
  CUDA_CHECK(cudaGetLastError());
}

void wae_ensure_gather(wae_ctx* h, Pattern& P) {
  auto& G = P.gather;
  if (G.built) return;
  // staged elements per patch: bounded by shared memory; WAE_GATHER_TETS overrides for tuning
  const int stride = h->nloc == 4 ? 16 : 64;  // doubles per staged element
  int target = h->nloc == 4 ? 840 : 208;      // 2 CTAs per SM
  if (const char* env = getenv("WAE_GATHER_TETS")) target = atoi(env);
  int cap = (int)((226 * 1024 - 1024) / (stride * sizeof(double))) - 1;
  if (cap > 1022) cap = 1022;
  if (target > cap) target = cap;
  if (target < 64) target = 64;
  GatherHost GH;
  wae_build_gather(h->xyz.data(), h->tets.data(), h->nloc, P, target, GH);
  G.n_patch = (int)GH.patch_tet_ptr.size() - 1;
  G.max_tets = GH.max_tets;
  G.n_src = (int64_t)GH.src.size();
  G.n_staged = (int64_t)GH.patch_tets.size();
  G.d_patch_tet_ptr.upload(GH.patch_tet_ptr, h->stream);
  G.d_patch_tets.upload(GH.patch_tets, h->stream);
  G.d_patch_grp_ptr.upload(GH.patch_grp_ptr, h->stream);
  G.d_patch_src_ptr.upload(GH.patch_src_ptr, h->stream);
  G.d_grp.upload(GH.grp, h->stream);
  G.d_out_idx.upload(GH.out_idx, h->stream);
  G.d_src.upload(GH.src, h->stream);
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  G.built = true;
}
