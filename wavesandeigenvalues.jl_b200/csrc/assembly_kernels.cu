// Element kernels: P1/P2 tetrahedral mass/stiffness, boundary-triangle mass, flame source.
// One element per thread; closed-form element matrices (fem_gen.h, exact integrals of the
// Lagrange bases == the reference's tables src/FEM/FEM.jl:435-450,704-738,1745-1874).
//
// Two generations of the tetrahedral M/K kernel live here:
//   * assemble_tet_atomic : scatter through a precomputed slot map with fp64 RED atomics
//     (any operator, any c; used for small sub-domain operators and as cross-check)
//   * assemble_tet_pairs  : owner-computes patches, coordinates and program staged in shared memory by bulk
//     copies, every nonzero written exactly once in a fixed summation order (the production path
//     for the big M+K pass; see assembly_symbolic.cpp for the pair program).
#include <cuda_runtime.h>

#include <algorithm>

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

// speaker source vector (Helmholtz.jl:193-210, 489-497; FEM.jl:2557-2589): m_i = -i * sum_tri c |det| int phi_i  (c constant per
// triangle) or -i * sum_tri |det| sum_k c_k int phi_i lambda_k (c linear); out is a dense complex vector over the DOFs
template <int NLOC3>
__global__ void __launch_bounds__(128) wallsrc_kernel(const double* __restrict__ xyz, const uint32_t* __restrict__ conn,
                                                      const int32_t* __restrict__ elems, int64_t n_elem, const double* __restrict__ c,
                                                      int c_per_elem, double* __restrict__ out /* complex */) {
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
  double adet = sqrt(nx * nx + ny * ny + nz * nz);
  const double* Ts = NLOC3 == 3 ? WAE_P1_TRI_SRC : WAE_P2_TRI_SRC;
  const double* Tc = NLOC3 == 3 ? WAE_P1_TRI_SRCC : WAE_P2_TRI_SRCC;
  for (int k = 0; k < NLOC3; k++) {
    double m;
    if (c_per_elem == 1)
      m = c[e] * Ts[k];
    else
      m = c[3 * e] * Tc[3 * k] + c[3 * e + 1] * Tc[3 * k + 1] + c[3 * e + 2] * Tc[3 * k + 2];
    if (m != 0.0) atomicAdd(out + 2 * (size_t)d[k] + 1, -m * adet);  // V ./= 1im  ->  imaginary part only
  }
}

void wae_launch_wallsrc(wae_ctx* h, const int32_t* d_elems, int64_t n, const double* d_c, int c_per_elem, double* d_out) {
  if (n == 0) return;
  unsigned blocks = (unsigned)((n + 127) / 128);
  if (h->nloc3 == 3)
    wallsrc_kernel<3><<<blocks, 128, 0, h->stream>>>(h->d_xyz.p, h->d_tris.p, d_elems, n, d_c, c_per_elem, d_out);
  else
    wallsrc_kernel<6><<<blocks, 128, 0, h->stream>>>(h->d_xyz.p, h->d_tris.p, d_elems, n, d_c, c_per_elem, d_out);
  h->launches++;
  CUDA_CHECK(cudaGetLastError());
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

// ---- generation 2: owner-computes pair program ---------------------------------------------
// Persistent kernel, one CTA per SM, patches round-robin (see assembly_symbolic.cpp for the program).  Per patch:
//   fetch        the program blob (local vertex numbers, element ids, group words, counts, store chunks) and the patch's vertex
//                coordinates arrive in shared memory by one bulk copy each (TMA, mbarrier), double-buffered: the copies for the
//                next patch run while the current one is processed; its slot words / store program are prefetched into L2.
//   element pass one staged element per lane.  Coordinates come from shared memory; the gram entries stay in registers; the packed
//                upper triangle of the stiffness (already times -c^2 |det|) and mass matrix is visited entry by entry (compile-time
//                index s) and every entry the patch owns is stored as a (K, M) pair into its shared-memory slot.
//   summation    one group of 32 units per warp step; lane l adds the slots base + 32 k + (l xor (k mod 8)), k < count -- 16-byte
//                loads, conflict-free -- and leaves the sum in slot base + l.
//   store pass   32 consecutive nonzeros of a column per warp step, each fetched from the slot the program names: full-line stores.
// Every nonzero is written exactly once, no atomics, fixed summation order (bit-reproducible), no memset of the outputs.
template <int NLOC>
struct SymVisit;
template <>
struct SymVisit<4> {
  template <class F>
  __device__ __forceinline__ static void run(const double* g, F&& f) { wae_p1_tet_sym_visit(g, f); }
};
template <>
struct SymVisit<10> {
  template <class F>
  __device__ __forceinline__ static void run(const double* g, F&& f) { wae_p2_tet_sym_visit(g, f); }
};

__device__ __forceinline__ void sts_pair(uint32_t addr, double k, double m) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(k), "d"(m) : "memory");
}

// ---- bulk copy (TMA, 1-D) + mbarrier helpers ----------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(mbar), "r"(parity)
                 : "memory");
  } while (!ok);
}

// vertex coordinates in patch order: pxyz[3 i .. 3 i + 2] = xyz of mesh vertex gvtx[i]
__global__ void gather_patch_xyz_kernel(const double* __restrict__ xyz, const uint32_t* __restrict__ gvtx, int64_t n, double* __restrict__ pxyz) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * n) return;
  pxyz[i] = xyz[3 * (size_t)gvtx[i / 3] + (i % 3)];
}

// CooTrafo from coordinates staged in shared memory (3 doubles per patch-local vertex)
__device__ __forceinline__ void tet_geom_smem(const double* __restrict__ px, const uint2 lv, TetGeom& t) {
  const double* p0 = px + 3 * (lv.x & 0xFFFFu);
  const double* p1 = px + 3 * (lv.x >> 16);
  const double* p2 = px + 3 * (lv.y & 0xFFFFu);
  const double* p3 = px + 3 * (lv.y >> 16);
  const double x3 = p3[0], y3 = p3[1], z3 = p3[2];
  double a[3][3];  // a[r][k] = component r of edge k
  a[0][0] = p0[0] - x3; a[1][0] = p0[1] - y3; a[2][0] = p0[2] - z3;
  a[0][1] = p1[0] - x3; a[1][1] = p1[1] - y3; a[2][1] = p1[2] - z3;
  a[0][2] = p2[0] - x3; a[1][2] = p2[1] - y3; a[2][2] = p2[2] - z3;
  double c00 = a[1][1] * a[2][2] - a[1][2] * a[2][1];
  double c01 = a[1][2] * a[2][0] - a[1][0] * a[2][2];
  double c02 = a[1][0] * a[2][1] - a[1][1] * a[2][0];
  double det = a[0][0] * c00 + a[0][1] * c01 + a[0][2] * c02;
  double id = 1.0 / det;
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

// entries s in [LO, HI) of one staged element: the element's gram entries from the staged coordinates, then the owned (K, M) pairs
template <int NLOC, int LO, int HI>
__device__ __forceinline__ void element_part(const double* __restrict__ px, const uint2 lv, double cc, double mass_scale,
                                             const uint32_t* __restrict__ dp, uint32_t sbase) {
  TetGeom tg;
  tet_geom_smem(px, lv, tg);
  double g[10];
  tet_gram(tg, -cc * cc * tg.adet, g);
  const double md = tg.adet * mass_scale;
  uint32_t w = 0;
  SymVisit<NLOC>::run(g, [&](auto S, double k, double m) {
    constexpr int s = decltype(S)::value;
    if constexpr (s >= LO && s < HI) {
      if (!(s & 1) || s == LO) w = dp[(s >> 1) * 32];  // the compiler hoists these independent loads as far as registers allow
      const uint32_t sl = (s & 1) ? (w >> 16) : (w & 0xFFFFu);
      if (sl != 0xFFFFu) sts_pair(sbase + (sl << 4), k, m * md);
    }
  });
}

__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// Persistent: CTA b works on patches b, b + gridDim.x, ...  While patch i is processed, the program blob and the staged
// coordinates of patch i+1 arrive in the other shared-memory buffer (bulk copy + mbarrier), its slot words and store program
// are prefetched into L2, and the descriptor of patch i+2 is fetched.
// MODE bit 0: mass -> out_m, bit 1: stiffness -> out_k
// VAR selects variants prepared from the phase shares of profiles/r01_ncu_assembly_phase_shares.txt (WAE_ASM_VARIANT, default 0 =
// the measured kernel).  bit 0: summation with four 16-byte loads in flight per warp -- the first four sources under predicates, the rest
// unrolled by four (the pass is latency-bound: the rolled loop has one load in flight per warp);
// bit 1: the element pass of a P2 patch split into three parts per element (a patch stages ~350 elements, i.e. 11 of the 32 warps
// had work; every part recomputes the geometry and visits a third of the packed triangle).  Same sums in the same order.
template <int NLOC, int MODE, int VAR>
__global__ void __launch_bounds__(1024, 1) assemble_tet_pairs(const int64_t* __restrict__ desc, int n_patch, const uint8_t* __restrict__ blob,
                                                               const double* __restrict__ pxyz, const double* __restrict__ c,
                                                               const uint32_t* __restrict__ dest, const uint16_t* __restrict__ res,
                                                               int slot_bytes, int pxyz_bytes, int buf_bytes, double mass_scale,
                                                               double* __restrict__ out_m, double* __restrict__ out_k, int dbg) {
  constexpr int NSYM = NLOC * (NLOC + 1) / 2, NPK = (NSYM + 1) / 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long mbar_storage[2];
  __shared__ __align__(16) long long sdesc[3][8];
  double2* slots = reinterpret_cast<double2*>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(slots);
  const uint32_t mbar0 = (uint32_t)__cvta_generic_to_shared(&mbar_storage[0]);
  const uint32_t lane16 = (uint32_t)lane << 4;
  const int G = gridDim.x;
  int p = blockIdx.x;
  if (p >= n_patch) return;
  // issue the bulk copies of the patch described by sdesc[ds] into buffer b (one thread)
  auto fetch = [&](int ds, int b) {
    const long long* D = sdesc[ds];
    const int* I = reinterpret_cast<const int*>(D + 4);
    const uint32_t mb = mbar0 + 8u * b, dst = sbase + slot_bytes + (uint32_t)b * buf_bytes;
    mbar_expect_tx(mb, (uint32_t)I[4] + (uint32_t)I[1] * 24u);
    bulk_g2s(dst + pxyz_bytes, blob + D[0], (uint32_t)I[4], mb);
    bulk_g2s(dst, pxyz + D[1], (uint32_t)I[1] * 24u, mb);
    bulk_prefetch_l2(dest + (size_t)D[2] * NPK * 32, (uint32_t)((I[0] + 31) >> 5) * NPK * 128u);
    bulk_prefetch_l2(res + (size_t)D[3] * 32, (uint32_t)I[3] * 64u);
  };
  if (threadIdx.x < 8) {
    sdesc[0][threadIdx.x] = desc[(size_t)p * 8 + threadIdx.x];
    if (p + G < n_patch) sdesc[1][threadIdx.x] = desc[(size_t)(p + G) * 8 + threadIdx.x];
  }
  if (threadIdx.x == 0) {
    mbar_init(mbar0, 1);
    mbar_init(mbar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) fetch(0, 0);
  for (int i = 0; p < n_patch; i++, p += G) {
    const int b = i & 1, ds = i % 3;
    if (threadIdx.x == 0 && p + G < n_patch) fetch((i + 1) % 3, b ^ 1);
    long long dnext = 0;
    if (threadIdx.x < 8 && p + 2 * G < n_patch) dnext = desc[(size_t)(p + 2 * G) * 8 + threadIdx.x];
    const longlong2 d23 = *reinterpret_cast<const longlong2*>(&sdesc[ds][2]);
    const int4 i03 = *reinterpret_cast<const int4*>(&sdesc[ds][4]);
    const int4 i47 = *reinterpret_cast<const int4*>(&sdesc[ds][6]);
    const int nt = i03.x, ng = i03.z, nc = i03.w;
    const unsigned char* buf = smem_raw + slot_bytes + (size_t)b * buf_bytes;
    const double* px = reinterpret_cast<const double*>(buf);
    const unsigned char* pb = buf + pxyz_bytes;
    const uint2* lvtx = reinterpret_cast<const uint2*>(pb);
    const int32_t* tets = reinterpret_cast<const int32_t*>(pb + i47.y);
    const uint32_t* grp = reinterpret_cast<const uint32_t*>(pb + i47.z);
    const unsigned char* cnt = pb + i47.w;
    const uint2* chunk = reinterpret_cast<const uint2*>(pb + i47.w + 32 * ng);
    const int nwb = (nt + 31) >> 5;
    mbar_wait(mbar0 + 8u * b, (uint32_t)(i >> 1) & 1u);
    // ---- element pass: one staged element per lane, 32 elements per warp step
    if constexpr ((VAR & 2) && NLOC == 10) {
      constexpr int C1 = 18, C2 = 36;  // even cuts: a part starts on a fresh slot word
      for (int job = warp; job < 3 * nwb && dbg != 2; job += nwarp) {
        const int wb = job / 3, part = job - 3 * wb;
        const int t = wb * 32 + lane;
        if (t >= nt) continue;
        const uint32_t* dp = dest + ((size_t)(d23.x + wb) * NPK) * 32 + lane;
        const double cc = (MODE & 2) ? c[tets[t]] : 0.0;
        if (part == 0)
          element_part<NLOC, 0, C1>(px, lvtx[t], cc, mass_scale, dp, sbase);
        else if (part == 1)
          element_part<NLOC, C1, C2>(px, lvtx[t], cc, mass_scale, dp, sbase);
        else
          element_part<NLOC, C2, NSYM>(px, lvtx[t], cc, mass_scale, dp, sbase);
      }
    } else {
      for (int wb = warp; wb < nwb && dbg != 2; wb += nwarp) {
        const int t = wb * 32 + lane;
        if (t >= nt) continue;
        const uint32_t* dp = dest + ((size_t)(d23.x + wb) * NPK) * 32 + lane;
        const double cc = (MODE & 2) ? c[tets[t]] : 0.0;
        element_part<NLOC, 0, NSYM>(px, lvtx[t], cc, mass_scale, dp, sbase);
      }
    }
    __syncthreads();
    // first batch of the store pass: issued here so that its latency hides behind the summation pass
    const uint16_t* rp = res + (size_t)d23.y * 32 + lane;
    uint32_t r[4];
#pragma unroll
    for (int u = 0; u < 4; u++) r[u] = warp + u * nwarp < nc ? rp[(size_t)(warp + u * nwarp) * 32] : 0;
    // ---- summation pass: lane l of group g adds slots base + 32 k + (l xor (k mod 8)), k < count; the sum stays in slot base + l
    for (int g = warp; g < ng && dbg != 1; g += nwarp) {
      const uint32_t hdr = grp[g];
      const int cn = cnt[g * 32 + lane];
      const uint32_t base = sbase + ((hdr >> 8) << 4);
      double ak = 0.0, am = 0.0;
      auto lds = [&](int k, double& vx, double& vy) {
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(vx), "=d"(vy) : "r"(base + (k << 9) + (lane16 ^ ((k & 7) << 4))) : "memory");
      };
      int k = 0;
      if constexpr (VAR & 1) {
        // most units have two to four sources (74 sources on 28 units per P2 tetrahedron): the first four loads are issued back to back
        // under predicates, missing sources contribute +0.0 (bitwise neutral after the first addition to the +0.0 start value)
        double x0 = 0.0, y0 = 0.0, x1 = 0.0, y1 = 0.0, x2 = 0.0, y2 = 0.0, x3 = 0.0, y3 = 0.0;
        if (cn > 0) lds(0, x0, y0);
        if (cn > 1) lds(1, x1, y1);
        if (cn > 2) lds(2, x2, y2);
        if (cn > 3) lds(3, x3, y3);
        ak = (((ak + x0) + x1) + x2) + x3;
        am = (((am + y0) + y1) + y2) + y3;
        k = cn < 4 ? cn : 4;
#pragma unroll 1
        for (; k + 4 <= cn; k += 4) {  // four loads in flight, then the same left-to-right sum as the rolled loop
          double x0, y0, x1, y1, x2, y2, x3, y3;
          lds(k, x0, y0);
          lds(k + 1, x1, y1);
          lds(k + 2, x2, y2);
          lds(k + 3, x3, y3);
          ak = (((ak + x0) + x1) + x2) + x3;
          am = (((am + y0) + y1) + y2) + y3;
        }
      }
#pragma unroll 1
      for (; k < cn; k++) {
        double vx, vy;
        lds(k, vx, vy);
        ak += vx;
        am += vy;
      }
      if (cn) sts_pair(base + lane16, ak, am);
    }
    __syncthreads();
    // ---- store pass: at most 32 consecutive nonzeros (one 256-byte line of each value array) per step, four steps in flight
    for (int ch = warp; ch < nc && dbg != 1 && dbg != 3; ch += 4 * nwarp) {
      uint2 h[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int cu = ch + u * nwarp;
        h[u] = cu < nc ? chunk[cu] : make_uint2(0, 0);
        if (ch != warp) r[u] = cu < nc ? rp[(size_t)cu * 32] : 0;
      }
#pragma unroll
      for (int u = 0; u < 4; u++)
        if (lane < (int)h[u].y) {
          const double2 v = slots[r[u]];
          if (MODE & 2) out_k[(size_t)h[u].x + lane] = v.x;
          if (MODE & 1) out_m[(size_t)h[u].x + lane] = v.y;
        }
    }
    if (threadIdx.x < 8 && p + 2 * G < n_patch) sdesc[(i + 2) % 3][threadIdx.x] = dnext;
    __syncthreads();  // slots, the program buffer and the descriptor ring are free for the next patches
  }
}

void wae_launch_assemble_gather(wae_ctx* h, Pattern& P, const double* d_c, double* d_mass, double* d_stiff, double mass_scale) {
  auto& G = P.gather;
  if (!G.built || G.n_patch == 0) return;
  if (G.xyz_version != h->xyz_version) {  // refresh the patch-ordered coordinates
    gather_patch_xyz_kernel<<<(unsigned)((3 * G.n_pv + 255) / 256), 256, 0, h->stream>>>(h->d_xyz.p, G.d_gvtx.p, G.n_pv, G.d_pxyz.p);
    G.xyz_version = h->xyz_version;
    h->launches++;
  }
  const int slot_bytes = G.max_slots * 16, pxyz_bytes = (G.max_nv * 24 + 15) & ~15, buf_bytes = pxyz_bytes + ((G.max_blob + 15) & ~15);
  const size_t smem = (size_t)slot_bytes + 2 * (size_t)buf_bytes;
  const int dbg = getenv("WAE_GATHER_DBG") ? atoi(getenv("WAE_GATHER_DBG")) : 0;
  int threads = smem > 113 * 1024 ? 1024 : 512;  // one or two CTAs per SM, 64 registers per thread either way
  int ctas = smem > 113 * 1024 ? 1 : 2;
  if (const char* env = getenv("WAE_GATHER_CTAS")) {  // tuning: more, smaller CTAs per SM (with WAE_GATHER_SLOTS small enough to fit)
    const int fit = (int)((227 * 1024) / (smem + 1024));
    ctas = std::max(1, std::min(atoi(env), fit));
    threads = std::max(128, (1024 / ctas) & ~31);
  }
  if (const char* env = getenv("WAE_GATHER_THREADS")) threads = std::max(32, std::min(1024, atoi(env) & ~31));
  const int grid = std::min(G.n_patch, h->sm_count * ctas);
  auto launch = [&](auto kern) {
    // per device, and cheap: set on every launch (a process may hold contexts on several GPUs)
    CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
    // several CTAs per SM only co-reside if the L1 / shared-memory split leaves room for all of them (a hint; the one-CTA layout needs the
    // maximum anyway)
    if (cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess)
      cudaGetLastError();  // a refused hint must not surface as the launch's error
    kern<<<grid, threads, smem, h->stream>>>(G.d_desc.p, G.n_patch, G.d_blob.p, G.d_pxyz.p, d_c, G.d_dest.p, G.d_res.p, slot_bytes, pxyz_bytes,
                                             buf_bytes, mass_scale, d_mass, d_stiff, dbg);
  };
  const bool both = d_stiff != nullptr;
  const int var = getenv("WAE_ASM_VARIANT") ? atoi(getenv("WAE_ASM_VARIANT")) & 3 : 0;
  auto pick = [&](auto nloc_c, auto mode_c) {
    constexpr int NL = decltype(nloc_c)::value, MD = decltype(mode_c)::value;
    switch (var) {
      case 1: launch(assemble_tet_pairs<NL, MD, 1>); break;
      case 2: launch(assemble_tet_pairs<NL, MD, 2>); break;
      case 3: launch(assemble_tet_pairs<NL, MD, 3>); break;
      default: launch(assemble_tet_pairs<NL, MD, 0>); break;
    }
  };
  if (h->nloc == 4) {
    if (both) pick(WaeIdx<4>{}, WaeIdx<3>{}); else pick(WaeIdx<4>{}, WaeIdx<1>{});
  } else {
    if (both) pick(WaeIdx<10>{}, WaeIdx<3>{}); else pick(WaeIdx<10>{}, WaeIdx<1>{});
  }
  h->launches++;
  CUDA_CHECK(cudaGetLastError());
}

void wae_ensure_gather(wae_ctx* h, Pattern& P) {
  auto& G = P.gather;
  if (G.built) return;
  if (P.elems.empty() || P.nnz == 0) {  // nothing to assemble: no program, no launch
    G.n_patch = 0;
    G.built = true;
    return;
  }
  // shared memory = slots (16 B each) + two buffers (staged coordinates + program blob of the current and the next patch);
  // the loop below shrinks the slot count until everything fits one SM (one CTA per SM); WAE_GATHER_SLOTS overrides for tuning
  int slot_cap = 12288;
  if (const char* env = getenv("WAE_GATHER_SLOTS")) slot_cap = atoi(env);
  slot_cap = std::max(1024, std::min(slot_cap, (226 * 1024) / 16));
  GatherHost GH;
  const size_t smem_max = 226 * 1024;
  for (;;) {
    wae_build_gather(h->xyz.data(), h->tets.data(), h->nloc, P, slot_cap, GH);
    const size_t need = (size_t)GH.max_slots * 16 + 2 * ((((size_t)GH.max_nv * 24 + 15) & ~(size_t)15) + (((size_t)GH.max_blob + 15) & ~(size_t)15));
    if (need <= smem_max) break;
    slot_cap -= (int)((need - smem_max) / 16) + 256;
    if (slot_cap < 1024) WAE_THROW(WAE_E_INVALID, "pair program does not fit shared memory");
  }
  G.n_patch = GH.n_patch;
  G.max_slots = GH.max_slots;
  G.max_blob = GH.max_blob;
  G.max_nv = GH.max_nv;
  G.n_pairs = GH.n_pairs;
  G.n_sources = GH.n_sources;
  G.n_staged = GH.n_staged;
  G.n_pv = (int64_t)GH.gvtx.size();
  G.d_desc.upload(GH.desc, h->stream);
  G.d_blob.upload(GH.blob.data(), GH.blob.size(), h->stream);
  G.d_gvtx.upload(GH.gvtx.data(), GH.gvtx.size(), h->stream);
  G.d_pxyz.alloc((size_t)G.n_pv * 3 + 2);
  G.d_dest.upload(GH.dest.data(), GH.dest.size(), h->stream);
  G.d_res.upload(GH.res.data(), GH.res.size(), h->stream);
  G.xyz_version = 0;
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  G.built = true;
}
