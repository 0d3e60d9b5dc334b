// Generation 3 of the tetrahedral M/K assembly: the STAR kernel (program: assembly_star_symbolic.cpp).
//
// Reference semantics: the element loops + sparse() of src/Helmholtz.jl:405-445,515 with the P1/P2 tables of
// src/FEM/FEM.jl:704-738,1745-1874 (M = int phi_i phi_j, K = -c^2 int grad phi_i . grad phi_j, c constant per element).
//
// Persistent CTAs, patches round-robin.  Per patch:
//   geometry pass  one staged element per thread: CooTrafo (FEM.jl:2-21) from the patch's vertex coordinates in shared memory,
//                  w * grad(l_i).grad(l_j) (4x4, w = -c^2 |det|) and |det| -> 17 doubles per element in shared memory.
//   fetch          the whole program of a patch arrives in shared memory by bulk copies (TMA, mbarrier): blob A (local vertex numbers,
//                  element ids, group headers, lane counts, star sources) + the patch's vertex coordinates one patch ahead, while the
//                  previous patch is stored; blob B (store chunks + codes) while the patch's own geometry / star passes run.  No pass
//                  waits on a dependent global load; the only scattered global read is c[element].
//   star pass      one sub-simplex (vertex / edge / face / tetrahedron) per lane, 32 of one type per warp step: the gram entries of
//                  the simplex's own vertices and |det| are summed over the star IN REGISTERS, the nonzeros of the simplex ("roles", fem_gen.h: wae_p*_star_*) are formed from the
//                  sums and written to the group's record block in shared memory, entry-major (row j: role j of the 32 simplices, then
//                  one row per mass class = coefficient * sum |det| -- conflict-free stores).  A vertex star (24 elements on a Kuhn
//                  mesh) is split over four lanes whose partial sums are combined by shuffles in a fixed order.
//   store pass     32 consecutive nonzeros of an owned column per warp step: a 16-bit code names the record entry of K and the row
//                  distance to M, both come straight from the record.  Full-line streaming stores, every nonzero written exactly once.
// No atomics, no memset of the outputs, fixed summation order (Morton order of the star's elements): bit-reproducible.
// Shared-memory traffic per P2 tetrahedron ~2 KB (generation 2: ~3 KB at twice the instruction count per byte), 15 star sources
// instead of 74 slot sources, no per-entry predicates.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <numeric>

#include "assembly_star.h"
#include "fem_gen.h"
#include "wae_internal.h"

namespace {
constexpr int GS = 17;  // doubles per staged element: 4x4 gram (row-major) + |det|

// ---- arithmetic shared by the kernel and its host replay (wae_star_program_check) -------------------------------------------
WAE_HD inline void star_geometry(const double* p0, const double* p1, const double* p2, const double* p3, double cc, bool with_k,
                                 double* out) {
  const double x3 = p3[0], y3 = p3[1], z3 = p3[2];
  double a[3][3];  // a[r][k] = component r of edge k
  a[0][0] = p0[0] - x3; a[1][0] = p0[1] - y3; a[2][0] = p0[2] - z3;
  a[0][1] = p1[0] - x3; a[1][1] = p1[1] - y3; a[2][1] = p1[2] - z3;
  a[0][2] = p2[0] - x3; a[1][2] = p2[1] - y3; a[2][2] = p2[2] - z3;
  const double c00 = a[1][1] * a[2][2] - a[1][2] * a[2][1];
  const double c01 = a[1][2] * a[2][0] - a[1][0] * a[2][2];
  const double c02 = a[1][0] * a[2][1] - a[1][1] * a[2][0];
  const double det = a[0][0] * c00 + a[0][1] * c01 + a[0][2] * c02;
  const double adet = fabs(det);
  out[16] = adet;
  if (!with_k) return;
  const double id = 1.0 / det;
  double G[4][3];  // rows of the inverse = grad(lambda_k); the fourth is minus their sum
  G[0][0] = c00 * id;
  G[0][1] = (a[0][2] * a[2][1] - a[0][1] * a[2][2]) * id;
  G[0][2] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) * id;
  G[1][0] = c01 * id;
  G[1][1] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) * id;
  G[1][2] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) * id;
  G[2][0] = c02 * id;
  G[2][1] = (a[0][1] * a[2][0] - a[0][0] * a[2][1]) * id;
  G[2][2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) * id;
  for (int d = 0; d < 3; d++) G[3][d] = -(G[0][d] + G[1][d] + G[2][d]);
  const double w = -cc * cc * adet;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int i = 0; i < 4; i++)
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int j = i; j < 4; j++) {
      const double g = (G[i][0] * G[j][0] + G[i][1] * G[j][1] + G[i][2] * G[j][2]) * w;
      out[4 * i + j] = g;
      if (i != j) out[4 * j + i] = g;
    }
}

// Sums of one sub-simplex (or of one slice of a vertex star) over its sources: S = gram entries of the simplex's own vertices,
// W = |det|.  sp[32 k] is source word k; Gs the staged gram blocks.  NV = number of vertices of the simplex.  The sums run in
// source order (k = 0, 1, ...), two sources in flight.
template <int NLOC, int NV>
struct StarNS {
  static constexpr int value = NV == 1 ? 1 : NV == 2 ? (NLOC == 4 ? 1 : 3) : 6;
};
template <int NLOC, int MODE, int NV>
WAE_HD inline void star_sums(int cnt, const uint16_t* sp, const double* Gs, double* S, double& W) {
  constexpr bool WK = (MODE & 2) != 0;
  constexpr int NS = StarNS<NLOC, NV>::value;
  W = 0.0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int i = 0; i < NS; i++) S[i] = 0.0;
  auto term = [&](uint32_t w, double* s, double& det) {
    const double* g = Gs + (size_t)(NV < 4 ? w >> (2 * NV) : w) * GS;
    det = g[16];
    if (!WK) return;
    const int ia = w & 3, ib = (w >> 2) & 3, ic = (w >> 4) & 3;
    if (NV == 1) {
      s[0] = g[5 * ia];
    } else if (NV == 2) {
      if (NLOC == 4) {
        s[0] = g[4 * ia + ib];
      } else {
        s[0] = g[5 * ia];
        s[1] = g[5 * ib];
        s[2] = g[4 * ia + ib];
      }
    } else if (NV == 3) {
      s[0] = g[5 * ia];
      s[1] = g[5 * ib];
      s[2] = g[5 * ic];
      s[3] = g[4 * ia + ib];
      s[4] = g[4 * ia + ic];
      s[5] = g[4 * ib + ic];
    } else {
      s[0] = g[1];   // the element's own frame: g01, g02, g03, g12, g13, g23
      s[1] = g[2];
      s[2] = g[3];
      s[3] = g[6];
      s[4] = g[7];
      s[5] = g[11];
    }
  };
  int k = 0;
  for (; k + 2 <= cnt; k += 2) {
    double s0[NS], s1[NS], d0, d1;
    term(sp[32 * k], s0, d0);
    term(sp[32 * k + 32], s1, d1);
    W = (W + d0) + d1;
    if (WK) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
      for (int i = 0; i < NS; i++) S[i] = (S[i] + s0[i]) + s1[i];
    }
  }
  if (k < cnt) {
    double s0[NS], d0;
    term(sp[32 * k], s0, d0);
    W += d0;
    if (WK) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
      for (int i = 0; i < NS; i++) S[i] += s0[i];
    }
  }
}

// Roles of one sub-simplex from its sums -> the simplex's column of the group's record block (rec[WAE_STAR_RS r] = row r, assembly_star.h)
template <int NLOC, int MODE, int NV>
WAE_HD inline void star_finish(const double* S, double W, double mass_scale, double* rec) {
  constexpr bool WK = (MODE & 2) != 0;
  double m[16];
  if (NLOC == 4) wae_p1_star_mass(m); else wae_p2_star_mass(m);
  if (NLOC == 4) {
    if (WK) {
      double K[1];
      if (NV == 1) wae_p1_star_vert(S, K); else wae_p1_star_edge(S, K);
      rec[0] = K[0];
    }
    rec[1 * WAE_STAR_RS] = W * (m[NV == 1 ? 0 : 1] * mass_scale);
  } else if (NV == 1) {
    if (WK) {
      double K[1];
      wae_p2_star_vert(S, K);
      rec[0] = K[0];
    }
    rec[1 * WAE_STAR_RS] = W * (m[0] * mass_scale);
  } else if (NV == 2) {
    if (WK) {
      double K[4];
      wae_p2_star_edge(S, K);
      rec[0] = K[0]; rec[1 * WAE_STAR_RS] = K[1]; rec[2 * WAE_STAR_RS] = K[2]; rec[3 * WAE_STAR_RS] = K[3];
    }
    rec[4 * WAE_STAR_RS] = W * (m[1] * mass_scale);
    rec[5 * WAE_STAR_RS] = W * (m[2] * mass_scale);
    rec[6 * WAE_STAR_RS] = W * (m[4] * mass_scale);
  } else if (NV == 3) {
    if (WK) {
      double K[6];
      wae_p2_star_face(S, K);
      rec[0] = K[0]; rec[1 * WAE_STAR_RS] = K[1]; rec[2 * WAE_STAR_RS] = K[2]; rec[3 * WAE_STAR_RS] = K[3]; rec[4 * WAE_STAR_RS] = K[4]; rec[5 * WAE_STAR_RS] = K[5];
    }
    rec[6 * WAE_STAR_RS] = W * (m[5] * mass_scale);
    rec[7 * WAE_STAR_RS] = W * (m[8] * mass_scale);
  } else {
    if (WK) {
      double K[3];
      wae_p2_star_tet(S, K);
      rec[0] = K[0]; rec[1 * WAE_STAR_RS] = K[1]; rec[2 * WAE_STAR_RS] = K[2];
    }
    rec[3 * WAE_STAR_RS] = W * (m[11] * mass_scale);
  }
}

// edge / face / tetrahedron: one simplex per lane
template <int NLOC, int MODE, int NV>
WAE_HD inline void star_simplex(int cnt, const uint16_t* sp, const double* Gs, double mass_scale, double* rec) {
  double S[StarNS<NLOC, NV>::value], W;
  star_sums<NLOC, MODE, NV>(cnt, sp, Gs, S, W);
  star_finish<NLOC, MODE, NV>(S, W, mass_scale, rec);
}

// ---- device helpers: bulk copy (TMA, 1-D) + mbarrier -----------------------------------------------------------------------------
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

// MODE bit 0: mass -> out_m, bit 1: stiffness -> out_k
template <int NLOC, int MODE>
__global__ void __launch_bounds__(1024, 1) assemble_tet_stars(const StarDesc* __restrict__ desc, int n_patch, const uint8_t* __restrict__ blobA,
                                                               const uint8_t* __restrict__ blobB, const double* __restrict__ pxyz,
                                                               const double* __restrict__ c, int off_rec, int off_a, int off_px, int off_b,
                                                               double mass_scale, double* __restrict__ out_m,
                                                               double* __restrict__ out_k, int dbg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(16) StarDesc sdesc[2];
  __shared__ __align__(8) unsigned long long mbar_storage[2];
  double* Gs = reinterpret_cast<double*>(smem_raw);
  double* rec = reinterpret_cast<double*>(smem_raw + off_rec);
  const unsigned char* bufA = smem_raw + off_a;
  const double* px = reinterpret_cast<const double*>(smem_raw + off_px);
  const unsigned char* bufB = smem_raw + off_b;
  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = T >> 5, G = gridDim.x;
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem_raw);
  const uint32_t mbarA = (uint32_t)__cvta_generic_to_shared(&mbar_storage[0]), mbarB = mbarA + 8;
  int p = blockIdx.x;
  if (p >= n_patch) return;
  if (tid < 8) {
    reinterpret_cast<long long*>(&sdesc[0])[tid] = reinterpret_cast<const long long*>(desc + p)[tid];
    if (p + G < n_patch) reinterpret_cast<long long*>(&sdesc[1])[tid] = reinterpret_cast<const long long*>(desc + p + G)[tid];
  }
  if (tid == 0) {
    mbar_init(mbarA, 1);
    mbar_init(mbarB, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto fetchA = [&](const StarDesc& D) {  // one thread
    mbar_expect_tx(mbarA, (uint32_t)D.bytesA + (uint32_t)D.nv * 24u);
    bulk_g2s(sbase + off_a, blobA + D.blobA, (uint32_t)D.bytesA, mbarA);
    bulk_g2s(sbase + off_px, pxyz + D.pxyz, (uint32_t)D.nv * 24u, mbarA);
  };
  auto fetchB = [&](const StarDesc& D) {
    mbar_expect_tx(mbarB, (uint32_t)D.bytesB);
    bulk_g2s(sbase + off_b, blobB + D.blobB, (uint32_t)D.bytesB, mbarB);
  };
  if (tid == 0) fetchA(sdesc[0]);
  for (int i = 0; p < n_patch; i++, p += G) {
    const StarDesc& D = sdesc[i & 1];
    const int nt = D.nt, ng = D.ng, nc = D.nc;
    const StarBlob L(nt, ng, D.nsrc, nc, D.ncode);
    const uint2* lvtx = reinterpret_cast<const uint2*>(bufA);
    const int32_t* tets = reinterpret_cast<const int32_t*>(bufA + L.o_tets);
    const uint2* grp = reinterpret_cast<const uint2*>(bufA + L.o_grp);
    const unsigned char* cnt = bufA + L.o_cnt;
    const uint16_t* ssrc = reinterpret_cast<const uint16_t*>(bufA + L.o_src);
    const uint2* chunk = reinterpret_cast<const uint2*>(bufB);
    const uint16_t* codes = reinterpret_cast<const uint16_t*>(bufB + L.o_code);
    const uint32_t par = (uint32_t)i & 1u;
    if (tid == 0) fetchB(D);  // the store program of this patch arrives while its geometry and star passes run
    mbar_wait(mbarA, par);
    // ---- geometry pass
    for (int t = tid; t < nt && !(dbg & 1); t += T) {
      const uint2 lv = lvtx[t];
      const double cc = (MODE & 2) ? __ldg(c + tets[t]) : 0.0;
      star_geometry(px + 3 * (lv.x & 0xFFFFu), px + 3 * (lv.x >> 16), px + 3 * (lv.y & 0xFFFFu), px + 3 * (lv.y >> 16), cc, (MODE & 2) != 0,
                    Gs + (size_t)t * GS);
    }
    __syncthreads();
    // ---- star pass: one simplex per lane (a vertex star: WAE_STAR_VSPLIT lanes), groups of 32 lanes of one type
    for (int g = warp; g < ng && !(dbg & 2); g += nwarp) {
      const uint2 h = grp[g];
      const int n = cnt[g * 32 + lane];
      const int type = (h.y >> 16) & 15, r0 = (int)(h.y & 0xFFFFu) * WAE_STAR_RS;
      const uint16_t* sp = ssrc + h.x + lane;
      double* rc = rec + r0 + lane;
      if (type == 0) {  // all lanes take part in the combination of the partial sums
        double S[1], W;
        star_sums<NLOC, MODE, 1>(n, sp, Gs, S, W);
        S[0] += __shfl_xor_sync(0xffffffffu, S[0], 1);
        W += __shfl_xor_sync(0xffffffffu, W, 1);
        S[0] += __shfl_xor_sync(0xffffffffu, S[0], 2);
        W += __shfl_xor_sync(0xffffffffu, W, 2);
        if (n && !(lane & (WAE_STAR_VSPLIT - 1))) star_finish<NLOC, MODE, 1>(S, W, mass_scale, rc);
      } else if (n) {
        if (type == 1)
          star_simplex<NLOC, MODE, 2>(n, sp, Gs, mass_scale, rc);
        else if constexpr (NLOC == 10) {
          if (type == 2)
            star_simplex<NLOC, MODE, 3>(n, sp, Gs, mass_scale, rc);
          else
            star_simplex<NLOC, MODE, 4>(n, sp, Gs, mass_scale, rc);
        }
      }
    }
    // descriptor of the patch after the next one: into the ring slot this patch's descriptor frees at the end of the iteration
    long long dnext = 0;
    if (tid < 8 && p + 2 * G < n_patch) dnext = reinterpret_cast<const long long*>(desc + p + 2 * G)[tid];
    __syncthreads();  // records complete; blob A and the coordinates are free
    if (tid == 0 && p + G < n_patch) fetchA(sdesc[(i + 1) & 1]);  // next patch's blob A arrives while this patch is stored
    mbar_wait(mbarB, par);
    // ---- store pass: at most 32 consecutive nonzeros per warp step; a warp takes four consecutive chunks at a time (the chunk
    // list is padded to a multiple of four with empty chunks).  Lanes past the end of a chunk read the chunk's first code: no
    // divergent regions, only the stores are predicated.
    for (int c4 = 4 * warp; c4 < nc && !(dbg & 4); c4 += 4 * nwarp) {
      const uint4 h01 = *reinterpret_cast<const uint4*>(chunk + c4), h23 = *reinterpret_cast<const uint4*>(chunk + c4 + 2);
      const uint32_t z0[4] = {h01.x, h01.z, h23.x, h23.z}, hy[4] = {h01.y, h01.w, h23.y, h23.w};
      uint32_t cw[4];
#pragma unroll
      for (int u = 0; u < 4; u++) cw[u] = codes[(hy[u] >> 6) + (lane < (int)(hy[u] & 63u) ? lane : 0)];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const double* r = rec + (cw[u] >> 3);  // the K entry; the M entry sits (cw & 7) rows further
        const double kv = (MODE & 2) ? r[0] : 0.0, mv = r[WAE_STAR_RS * (cw[u] & 7)];
        if (lane < (int)(hy[u] & 63u)) {
          const uint32_t z = z0[u] + (uint32_t)lane;
          if (MODE & 2) __stcs(out_k + z, kv);
          if (MODE & 1) __stcs(out_m + z, mv);
        }
      }
    }
    __syncthreads();  // records, group table, blob B and this patch's descriptor slot are free
    if (tid < 8 && p + 2 * G < n_patch) reinterpret_cast<long long*>(&sdesc[i & 1])[tid] = dnext;
    // (the write above is ordered before its first readers by the barrier after the next patch's geometry pass; thread 0, the
    //  only earlier reader -- fetchB / fetchA of that slot happen two iterations later -- reads it after a barrier as well)
  }
}

__global__ void star_patch_xyz_kernel(const double* __restrict__ xyz, const uint32_t* __restrict__ gvtx, int64_t n, double* __restrict__ pxyz) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * n) return;
  pxyz[i] = xyz[3 * (size_t)gvtx[i / 3] + (i % 3)];
}

}  // namespace

void wae_launch_assemble_star(wae_ctx* h, Pattern& P, const double* d_c, double* d_mass, double* d_stiff, double mass_scale) {
  auto& G = P.star;
  if (!G.built || G.n_patch == 0) return;
  if (G.xyz_version != h->xyz_version) {
    star_patch_xyz_kernel<<<(unsigned)((3 * G.n_pv + 255) / 256), 256, 0, h->stream>>>(h->d_xyz.p, G.d_gvtx.p, G.n_pv, G.d_pxyz.p);
    G.xyz_version = h->xyz_version;
    h->launches++;
  }
  const StarLayout L(G.max_nt, G.max_rows, G.max_a, G.max_b, G.max_nv);
  const size_t smem = (size_t)L.total;
  int ctas = std::max(1, (int)((228 * 1024) / (smem + 1024 + WAE_STAR_STATIC_SMEM)));
  int threads = ctas >= 4 ? 256 : ctas >= 2 ? 512 : 1024;
  if (const char* env = getenv("WAE_STAR_THREADS")) threads = std::max(64, std::min(1024, atoi(env) & ~31));
  ctas = std::min(ctas, std::max(1, 1024 / threads));  // 64 registers per thread
  if (const char* env = getenv("WAE_STAR_CTAS")) ctas = std::max(1, std::min(ctas, atoi(env)));
  const int grid = std::min(G.n_patch, h->sm_count * ctas);
  // WAE_STAR_DBG (phase timing; results are then wrong by design): bit 0 skips the geometry pass, bit 1 the star pass, bit 2 the store pass
  const int dbg = getenv("WAE_STAR_DBG") ? atoi(getenv("WAE_STAR_DBG")) : 0;
  auto launch = [&](auto kern) {
    CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
    if (cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess)
      cudaGetLastError();
    kern<<<grid, threads, smem, h->stream>>>(reinterpret_cast<const StarDesc*>(G.d_desc.p), G.n_patch, G.d_blob.p, G.d_blobB.p, G.d_pxyz.p, d_c,
                                             L.off_rec, L.off_a, L.off_px, L.off_b, mass_scale, d_mass, d_stiff, dbg);
  };
  const bool both = d_stiff != nullptr;
  if (h->nloc == 4) {
    if (both) launch(assemble_tet_stars<4, 3>); else launch(assemble_tet_stars<4, 1>);
  } else {
    if (both) launch(assemble_tet_stars<10, 3>); else launch(assemble_tet_stars<10, 1>);
  }
  h->launches++;
  CUDA_CHECK(cudaGetLastError());
  // layout of this launch, readable through wae_last_ms (diagnostics of tools/sweep_star.py and bench.py)
  h->last_ms["star_patches"] = G.n_patch;
  h->last_ms["star_staged"] = (double)G.n_staged;
  h->last_ms["star_simplices"] = (double)G.n_entities;
  h->last_ms["star_sources"] = (double)G.n_sources;
  h->last_ms["star_program_bytes"] = (double)G.program_bytes + 24.0 * (double)G.n_pv;
  h->last_ms["star_smem"] = (double)smem;
  h->last_ms["star_threads"] = threads;
  h->last_ms["star_ctas_per_sm"] = ctas;
}

// Shared-memory budget of one CTA: WAE_STAR_SMEM (bytes) or, by default, what lets two CTAs share an SM (one CTA stores while
// the other computes).
static int64_t star_budget() {
  int64_t b = 112 * 1024;
  if (const char* env = getenv("WAE_STAR_SMEM")) b = atoll(env);
  return std::max<int64_t>(16 * 1024, std::min<int64_t>(b, 225 * 1024));
}

bool wae_ensure_star(wae_ctx* h, Pattern& P) {
  auto& G = P.star;
  const int64_t budget = star_budget();
  if (G.built && G.budget == budget) return !G.failed;
  G = Pattern::Star();
  G.budget = budget;
  G.built = true;
  if (P.elems.empty() || P.nnz == 0) return true;  // nothing to assemble: no program, no launch
  StarHost SH;
  try {
    wae_build_star(h->xyz.data(), h->tets.data(), h->nloc, P, budget, SH);
  } catch (const WaeError&) {  // e.g. a vertex shared by more than 255 elements: the caller uses the pair program / atomics
    G.failed = true;
    return false;
  }
  G.n_patch = SH.n_patch;
  G.max_nt = SH.max_nt;
  G.max_nv = SH.max_nv;
  G.max_rows = SH.max_rows;
  G.max_ng = SH.max_ng;
  G.max_a = SH.max_a;
  G.max_b = SH.max_b;
  G.max_smem = SH.max_smem;
  G.n_staged = SH.n_staged;
  G.n_entities = SH.n_entities;
  G.n_sources = SH.n_sources;
  G.n_chunks = SH.n_chunks;
  G.n_pv = (int64_t)SH.gvtx.size();
  G.program_bytes = (int64_t)(SH.desc.size() * 8 + SH.blob.size() + SH.blobB.size() + SH.gvtx.size() * 4);
  G.d_desc.upload(SH.desc, h->stream);
  G.d_blob.upload(SH.blob.data(), SH.blob.size(), h->stream);
  G.d_blobB.upload(SH.blobB.data(), SH.blobB.size(), h->stream);
  G.d_gvtx.upload(SH.gvtx.data(), SH.gvtx.size(), h->stream);
  G.d_pxyz.alloc((size_t)G.n_pv * 3 + 2);
  G.xyz_version = 0;
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  return true;
}

// ---- host replay (no GPU, no context): pattern + star program of a tetrahedral mesh, then the three passes of
// assemble_tet_stars with the kernel's own arithmetic (the functions above compiled for the host).  See include/wae_b200.h.
static int64_t g_sim[8], g_simtype[4][2];
// shared-memory wavefronts of one 8-byte warp load: per half-warp, the largest number of distinct addresses in one bank pair
static int wavefronts64(const int64_t* addr, const bool* act, int* ideal) {
  int tot = 0;
  for (int h = 0; h < 2; h++) {
    int64_t seen[16][16];
    int n[16] = {0};
    int worst = 0;
    for (int l = 16 * h; l < 16 * h + 16; l++) {
      if (!act[l]) continue;
      const int b = (int)(addr[l] & 15);
      bool dup = false;
      for (int q = 0; q < n[b]; q++) dup |= seen[b][q] == addr[l];
      if (!dup) seen[b][n[b]++] = addr[l];
      worst = std::max(worst, n[b]);
    }
    tot += worst;
    *ideal += worst ? 1 : 0;
  }
  return tot;
}

template <int NLOC>
static void star_replay(const StarHost& G, const double* xyz, const double* c, double mass_scale, double* val_m, double* val_k,
                        std::vector<int>& written, int64_t& bad, int64_t* wf /* star gathers, ideal, store gathers, ideal */) {
  std::vector<double> Gs((size_t)G.max_nt * GS), rec((size_t)G.max_rows * WAE_STAR_RS), px((size_t)G.max_nv * 3);
  for (int p = 0; p < G.n_patch; p++) {
    const StarDesc& D = *reinterpret_cast<const StarDesc*>(&G.desc[(size_t)p * 8]);
    const StarBlob L(D.nt, D.ng, D.nsrc, D.nc, D.ncode);
    if (D.blobA % 16 || D.blobB % 16 || (D.pxyz * 8) % 16 || D.nv % 2 || D.bytesA != L.bytesA || D.bytesB != L.bytesB || D.bytesA > G.max_a ||
        D.bytesB > G.max_b || D.nt > G.max_nt || D.nv > G.max_nv || D.ng > G.max_ng)
      bad++;
    const uint8_t* A = G.blob.data() + D.blobA;
    const uint8_t* B = G.blobB.data() + D.blobB;
    const uint16_t* lv = reinterpret_cast<const uint16_t*>(A);
    const int32_t* tets = reinterpret_cast<const int32_t*>(A + L.o_tets);
    const uint32_t* grp = reinterpret_cast<const uint32_t*>(A + L.o_grp);
    const uint8_t* cnt = A + L.o_cnt;
    const uint16_t* src = reinterpret_cast<const uint16_t*>(A + L.o_src);
    const uint32_t* chunk = reinterpret_cast<const uint32_t*>(B);
    const uint16_t* codes = reinterpret_cast<const uint16_t*>(B + L.o_code);
    for (int i = 0; i < D.nv; i++)
      for (int r = 0; r < 3; r++) px[3 * i + r] = xyz[3 * (size_t)G.gvtx[D.pxyz / 3 + i] + r];
    std::fill(rec.begin(), rec.end(), std::nan(""));
    for (int t = 0; t < D.nt; t++)
      star_geometry(&px[3 * lv[4 * t]], &px[3 * lv[4 * t + 1]], &px[3 * lv[4 * t + 2]], &px[3 * lv[4 * t + 3]], c[tets[t]], true, &Gs[(size_t)t * GS]);
    for (int g = 0; g < D.ng; g++) {
      const uint32_t so = grp[2 * g], hy = grp[2 * g + 1];
      const int type = (hy >> 16) & 15, niter = hy >> 24;
      const size_t r0 = (size_t)(hy & 0xFFFFu) * WAE_STAR_RS;
      if (r0 + (size_t)star_rows(NLOC, type) * WAE_STAR_RS > rec.size() || (NLOC == 4 && type > 1)) { bad++; continue; }
      // bank conflicts of the gram gathers: load j of source k, all lanes (the kernel's access pattern)
      for (int k = 0; k < niter; k++) {
        static const int nload[4] = {2, NLOC == 4 ? 2 : 4, 7, 7};
        for (int j = 0; j < nload[type]; j++) {
          int64_t addr[32], tt[32], ii[32];
          bool act[32];
          for (int l = 0; l < 32; l++) {
            act[l] = k < cnt[g * 32 + l];
            const uint32_t w = src[so + 32 * k + l];
            const int ia = w & 3, ib = (w >> 2) & 3, ic = (w >> 4) & 3;
            const int64_t t = type < 3 ? w >> (2 * (type + 1)) : w;
            int idx = 16;
            if (type == 0) idx = j == 0 ? 5 * ia : 16;
            else if (type == 1 && NLOC == 4) idx = j == 0 ? 4 * ia + ib : 16;
            else if (type == 1) idx = j == 0 ? 5 * ia : j == 1 ? 5 * ib : j == 2 ? 4 * ia + ib : 16;
            else if (type == 2) { const int q[7] = {5 * ia, 5 * ib, 5 * ic, 4 * ia + ib, 4 * ia + ic, 4 * ib + ic, 16}; idx = q[j]; }
            else { const int q[7] = {1, 2, 3, 6, 7, 11, 16}; idx = q[j]; }
            addr[l] = t * GS + idx;
            tt[l] = t;
            ii[l] = idx;
          }
          int id = 0;
          const int w0 = wavefronts64(addr, act, &id);
          wf[0] += w0;
          wf[1] += id;
          g_simtype[type][0] += w0;
          g_simtype[type][1] += id;
          if (getenv("WAE_STAR_SIM")) {  // layout experiments: other element strides, entry-major (SoA) storage
            static const int S[6] = {17, 19, 21, 23, 25, 27};
            for (int q = 0; q < 6; q++) {
              for (int l = 0; l < 32; l++) addr[l] = tt[l] * S[q] + ii[l];
              g_sim[q] += wavefronts64(addr, act, &id);
            }
            const int64_t ntp = (D.nt + 15) & ~15;
            for (int l = 0; l < 32; l++) addr[l] = ii[l] * ntp + tt[l];
            g_sim[6] += wavefronts64(addr, act, &id);
            for (int l = 0; l < 32; l++) addr[l] = ii[l] * (ntp + 1) + tt[l];
            g_sim[7] += wavefronts64(addr, act, &id);
          }
        }
      }
      double pS[32], pW[32];
      for (int l = 0; l < 32; l++) {
        const int n = cnt[g * 32 + l];
        if (n > niter || (int)so + 32 * niter > D.nsrc) bad++;
        const uint16_t* sp = src + so + l;
        double* rc = &rec[r0 + l];
        if (type == 0)
          star_sums<NLOC, 3, 1>(n, sp, Gs.data(), &pS[l], pW[l]);
        else if (!n)
          continue;
        else if (type == 1)
          star_simplex<NLOC, 3, 2>(n, sp, Gs.data(), mass_scale, rc);
        else if constexpr (NLOC == 10) {
          if (type == 2)
            star_simplex<NLOC, 3, 3>(n, sp, Gs.data(), mass_scale, rc);
          else
            star_simplex<NLOC, 3, 4>(n, sp, Gs.data(), mass_scale, rc);
        }
      }
      if (type == 0)  // the shuffle tree of the kernel: (p0 + p1) + (p2 + p3)
        for (int l = 0; l < 32; l += WAE_STAR_VSPLIT) {
          if (!cnt[g * 32 + l]) continue;
          const double S = (pS[l] + pS[l + 1]) + (pS[l + 2] + pS[l + 3]), W = (pW[l] + pW[l + 1]) + (pW[l + 2] + pW[l + 3]);
          star_finish<NLOC, 3, 1>(&S, W, mass_scale, &rec[r0 + l]);
        }
    }
    if (D.nc % 4) bad++;
    for (int ch = 0; ch < D.nc; ch++) {
      const uint32_t z0 = chunk[2 * ch], len = chunk[2 * ch + 1] & 63u, co = chunk[2 * ch + 1] >> 6;
      if (len == 0 && co == 0) continue;  // padding of the chunk list
      if (len < 1 || len > 32 || (z0 >> 5) != ((z0 + len - 1) >> 5) || (int)(co + len) > D.ncode) { bad++; continue; }
      int64_t ak[32], am[32];
      bool act[32];
      for (uint32_t l = 0; l < 32; l++) {  // lanes past the end read the chunk's first code
        const uint16_t cw = codes[co + (l < len ? l : 0)];
        act[l] = true;
        ak[l] = cw >> 3;
        am[l] = (cw >> 3) + WAE_STAR_RS * (cw & 7);
      }
      int id = 0;
      wf[2] += wavefronts64(ak, act, &id) + wavefronts64(am, act, &id);
      wf[3] += id;
      for (uint32_t l = 0; l < len; l++) {
        const uint16_t cw = codes[co + l];
        if ((size_t)(cw >> 3) + WAE_STAR_RS * (cw & 7) >= rec.size() || !(cw & 7)) { bad++; continue; }
        val_k[z0 + l] = rec[cw >> 3];
        val_m[z0 + l] = rec[(size_t)(cw >> 3) + WAE_STAR_RS * (cw & 7)];
        written[z0 + l]++;
      }
    }
  }
}

extern "C" int32_t wae_star_program_check(int32_t order, int64_t n_pts, const double* xyz, int64_t n_tet, const uint32_t* tets, const double* c,
                                          int64_t smem_budget, int64_t nnz_cap, int64_t* colptr, int32_t* rowval, double* val_m, double* val_k,
                                          double* stats) {
  try {
    (void)n_pts;
    const int nloc = order == 1 ? 4 : 10;
    int64_t dim = 0;
    for (int64_t k = 0; k < n_tet * nloc; k++) dim = std::max<int64_t>(dim, (int64_t)tets[k] + 1);
    Pattern P;
    P.elem_kind = 3;
    P.elems.resize(n_tet);
    std::iota(P.elems.begin(), P.elems.end(), 0);
    wae_build_pattern_from_elements(tets, nloc, P.elems, dim, P);
    if (P.nnz > nnz_cap) return WAE_E_INVALID;
    StarHost G;
    wae_build_star(xyz, tets, nloc, P, smem_budget, G);
    std::vector<int> written(P.nnz, 0);
    int64_t bad = 0, wf[4] = {0, 0, 0, 0};
    if (nloc == 4)
      star_replay<4>(G, xyz, c, 1.0, val_m, val_k, written, bad, wf);
    else
      star_replay<10>(G, xyz, c, 1.0, val_m, val_k, written, bad, wf);
    for (int64_t k = 0; k < P.nnz; k++) bad += written[k] != 1;
    std::copy(P.colptr.begin(), P.colptr.end(), colptr);
    std::copy(P.rowval.begin(), P.rowval.end(), rowval);
    const StarLayout L(G.max_nt, G.max_rows, G.max_a, G.max_b, G.max_nv);
    stats[0] = (double)P.nnz;
    stats[1] = G.n_patch;
    stats[2] = (double)G.n_staged;
    stats[3] = (double)G.n_entities;
    stats[4] = (double)G.n_sources;
    stats[5] = (double)L.total + WAE_STAR_STATIC_SMEM;
    stats[6] = (double)(G.desc.size() * 8 + G.blob.size() + G.blobB.size() + G.gvtx.size() * 4);
    stats[7] = (double)bad;
    if (getenv("WAE_STAR_SIM")) {
      fprintf(stderr, "[star sim] gram gathers, wavefronts per tetrahedron: strides 17 19 21 23 25 27 | SoA | SoA+1:");
      for (int q = 0; q < 8; q++) fprintf(stderr, " %.2f", (double)g_sim[q] / (double)n_tet), g_sim[q] = 0;
      fprintf(stderr, "\n[star sim] gram gathers by type (wavefronts / conflict-free per tetrahedron):");
      for (int q = 0; q < 4; q++) fprintf(stderr, "  %.2f / %.2f", (double)g_simtype[q][0] / (double)n_tet, (double)g_simtype[q][1] / (double)n_tet), g_simtype[q][0] = g_simtype[q][1] = 0;
      fprintf(stderr, "\n");
    }
    for (int i = 0; i < 4; i++) stats[8 + i] = (double)wf[i];  // shared-memory wavefronts of the gathers (simulated) and their conflict-free count
    return WAE_OK;
  } catch (const WaeError& e) {
    if (getenv("WAE_SYMB_TIMING")) fprintf(stderr, "[wae_star_program_check] %s\n", e.msg.c_str());
    return e.code;
  } catch (...) {
    return WAE_E_INVALID;
  }
}
