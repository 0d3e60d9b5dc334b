// Generation 3 of the tetrahedral M/K assembly: the STAR kernel (program: assembly_star_symbolic.cpp).
//
// Reference semantics: the element loops + sparse() of src/Helmholtz.jl:405-445,515 with the P1/P2 tables of
// src/FEM/FEM.jl:704-738,1745-1874 (M = int phi_i phi_j, K = -c^2 int grad phi_i . grad phi_j, c constant per element).
//
// Persistent CTAs, patches round-robin.  Per patch:
//   geometry pass  one staged element per thread: CooTrafo (FEM.jl:2-21) from the patch's vertex coordinates in shared memory,
//                  w * grad(l_i).grad(l_j) (4x4, w = -c^2 |det|) and |det| -> 17 doubles per element in shared memory.
//   star pass      one sub-simplex (vertex / edge / face / tetrahedron) per lane, 32 of one type per warp step: the gram entries of
//                  the simplex's own vertices and |det| are summed over the star IN REGISTERS (source words staged in shared memory
//                  by asynchronous copies), the nonzeros of the simplex ("roles", fem_gen.h: wae_p*_star_*) are formed from the
//                  sums and written to the group's record block in shared memory, entry-major (row 0: the 32 sums of |det|, row 1 + j:
//                  role j of the 32 simplices -- conflict-free stores).
//   store pass     32 consecutive nonzeros of an owned column per warp step: a 16-bit code names group, lane and role, K comes
//                  straight from the record, M = m_role * sum |det|.  Full-line streaming stores, every nonzero written exactly once.
// No atomics, no memset of the outputs, fixed summation order (Morton order of the star's elements): bit-reproducible.
// Shared-memory traffic per P2 tetrahedron ~2 KB (generation 2: ~3 KB at twice the instruction count per byte), 15 star sources
// instead of 74 slot sources, no per-entry predicates.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <numeric>

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

// One sub-simplex: sum over its star, form the roles, write the record (rec[32 j] = entry j).  sp[32 k] is source word k; Gs the
// staged gram blocks.
// NV = number of vertices of the simplex.  The sums run in source order (k = 0, 1, ...), two sources in flight.
template <int NLOC, int MODE, int NV>
WAE_HD inline void star_simplex(int cnt, const uint16_t* sp, const double* Gs, double* rec) {
  constexpr bool WK = (MODE & 2) != 0;
  constexpr int NS = NV == 1 ? 1 : NV == 2 ? (NLOC == 4 ? 1 : 3) : 6;
  double S[NS], W = 0.0;
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
  rec[0] = W;
  if (!WK) return;
  if (NLOC == 4) {
    double K[1];
    if (NV == 1) wae_p1_star_vert(S, K); else wae_p1_star_edge(S, K);
    rec[32] = K[0];
  } else if (NV == 1) {
    double K[1];
    wae_p2_star_vert(S, K);
    rec[32] = K[0];
  } else if (NV == 2) {
    double K[4];
    wae_p2_star_edge(S, K);
    rec[32] = K[0]; rec[64] = K[1]; rec[96] = K[2]; rec[128] = K[3];
  } else if (NV == 3) {
    double K[6];
    wae_p2_star_face(S, K);
    rec[32] = K[0]; rec[64] = K[1]; rec[96] = K[2]; rec[128] = K[3]; rec[160] = K[4]; rec[192] = K[5];
  } else {
    double K[3];
    wae_p2_star_tet(S, K);
    rec[32] = K[0]; rec[64] = K[1]; rec[96] = K[2];
  }
}

template <int NLOC, int MODE>
WAE_HD inline void star_group(int type, int cnt, const uint16_t* sp, const double* Gs, double* rec) {
  if (type == 0)
    star_simplex<NLOC, MODE, 1>(cnt, sp, Gs, rec);
  else if (type == 1)
    star_simplex<NLOC, MODE, 2>(cnt, sp, Gs, rec);
  else if constexpr (NLOC == 10) {
    if (type == 2)
      star_simplex<NLOC, MODE, 3>(cnt, sp, Gs, rec);
    else
      star_simplex<NLOC, MODE, 4>(cnt, sp, Gs, rec);
  }
}

template <int NLOC>
struct StarRec {
  // row of a role's K entry in its group's record block: vertex 1 | edge 1..4 | face 1..6 | tet 1..3
  WAE_HD static int koff(int role) { return NLOC == 4 ? 1 : (int)((0x32165432143211ULL >> (4 * role)) & 7); }
};

// ---- device helpers ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

struct StarDesc {
  long long blob, pxyz, src, code0;
  int nt, nv, ng, nc, nsrc, ncode, o_grp, o_cnt;
};

// MODE bit 0: mass -> out_m, bit 1: stiffness -> out_k
template <int NLOC, int MODE>
__global__ void __launch_bounds__(1024, 1) assemble_tet_stars(const StarDesc* __restrict__ desc, int n_patch, const uint8_t* __restrict__ blob,
                                                               const double* __restrict__ pxyz, const double* __restrict__ c,
                                                               const uint16_t* __restrict__ src, const uint16_t* __restrict__ code,
                                                               int off_rec, int off_src, int off_px, int off_gt, double mass_scale,
                                                               double* __restrict__ out_m, double* __restrict__ out_k) {
  using R = StarRec<NLOC>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double mlut[16];
  double* Gs = reinterpret_cast<double*>(smem_raw);
  double* rec = reinterpret_cast<double*>(smem_raw + off_rec);
  uint16_t* ssrc = reinterpret_cast<uint16_t*>(smem_raw + off_src);
  double* px = reinterpret_cast<double*>(smem_raw + off_px);
  int* gtab = reinterpret_cast<int*>(smem_raw + off_gt);  // group -> first double of its record block
  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = T >> 5;
  if (tid == 0) {
    double m[16];
    if (NLOC == 4) wae_p1_star_mass(m); else wae_p2_star_mass(m);
    for (int i = 0; i < (NLOC == 4 ? WAE_P1_STAR_ROLES : WAE_P2_STAR_ROLES); i++) mlut[i] = m[i] * mass_scale;
  }
  // asynchronous staging of a patch's source words and vertex coordinates (16-byte pieces; both are padded to multiples of 16)
  auto stage = [&](const StarDesc& D) {
    const uint4* gs = reinterpret_cast<const uint4*>(src + D.src);
    for (int i = tid; i < D.nsrc / 8; i += T) cp_async16(reinterpret_cast<uint4*>(ssrc) + i, gs + i);
    const uint4* gp = reinterpret_cast<const uint4*>(pxyz + D.pxyz);
    const int np = (D.nv * 24) / 16;
    for (int i = tid; i < np; i += T) cp_async16(reinterpret_cast<uint4*>(px) + i, gp + i);
  };
  int p = blockIdx.x;
  if (p < n_patch) stage(desc[p]);
  for (; p < n_patch; p += gridDim.x) {
    const StarDesc D = desc[p];
    const unsigned char* pb = blob + D.blob;
    const uint2* lvtx = reinterpret_cast<const uint2*>(pb);
    const int32_t* tets = reinterpret_cast<const int32_t*>(pb + ((8 * D.nt + 15) & ~15));
    const uint2* grp = reinterpret_cast<const uint2*>(pb + D.o_grp);
    const unsigned char* cnt = pb + D.o_cnt;
    const uint2* chunk = reinterpret_cast<const uint2*>(pb + D.o_cnt + 32 * D.ng);
    const uint16_t* cd = code + D.code0;
    // the store pass's codes: into L2 while the geometry and star passes run
    for (int i = tid; i < (D.ncode + 63) / 64; i += T) prefetch_l2(cd + (size_t)i * 64);
    cp_async_wait_all();
    __syncthreads();  // coordinates and source words of this patch have landed; mlut is set
    // ---- geometry pass
    for (int t = tid; t < D.nt; t += T) {
      const uint2 lv = __ldg(lvtx + t);
      const double cc = (MODE & 2) ? __ldg(c + __ldg(tets + t)) : 0.0;
      star_geometry(px + 3 * (lv.x & 0xFFFFu), px + 3 * (lv.x >> 16), px + 3 * (lv.y & 0xFFFFu), px + 3 * (lv.y >> 16), cc, (MODE & 2) != 0,
                    Gs + (size_t)t * GS);
    }
    __syncthreads();
    // ---- star pass: one simplex per lane, groups of 32 of one type
    for (int g = warp; g < D.ng; g += nwarp) {
      const uint2 h = __ldg(grp + g);
      const int n = __ldg(cnt + g * 32 + lane);
      const int type = (h.y >> 16) & 15, r0 = (int)(h.y & 0xFFFFu) * 32;
      if (lane == 0) gtab[g] = r0;
      if (n) star_group<NLOC, MODE>(type, n, ssrc + h.x + lane, Gs, rec + r0 + lane);
    }
    __syncthreads();
    // the next patch's sources and coordinates arrive while this patch's nonzeros are stored
    if (p + (int)gridDim.x < n_patch) stage(desc[p + gridDim.x]);
    // ---- store pass: at most 32 consecutive nonzeros per warp step, four steps in flight
    for (int ch = warp; ch < D.nc; ch += 4 * nwarp) {
      uint2 hd[4];
      uint32_t cw[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int cu = ch + u * nwarp;
        hd[u] = cu < D.nc ? __ldg(chunk + cu) : make_uint2(0, 0);
        cw[u] = lane < (int)(hd[u].y & 63u) ? __ldg(cd + (hd[u].y >> 6) + lane) : 0u;
      }
#pragma unroll
      for (int u = 0; u < 4; u++)
        if (lane < (int)(hd[u].y & 63u)) {
          const double* r = rec + gtab[cw[u] >> 9] + ((cw[u] >> 4) & 31);
          const int role = cw[u] & 15;
          if (MODE & 2) __stcs(out_k + (size_t)hd[u].x + lane, r[32 * R::koff(role)]);
          if (MODE & 1) __stcs(out_m + (size_t)hd[u].x + lane, r[0] * mlut[role]);
        }
    }
    // (the barrier at the top of the next iteration separates this store pass from the next geometry / star pass)
  }
  cp_async_wait_all();
}

__global__ void star_patch_xyz_kernel(const double* __restrict__ xyz, const uint32_t* __restrict__ gvtx, int64_t n, double* __restrict__ pxyz) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * n) return;
  pxyz[i] = xyz[3 * (size_t)gvtx[i / 3] + (i % 3)];
}

struct StarLayout {
  int off_rec, off_src, off_px, off_gt;
  size_t total;
};
StarLayout star_layout(int max_nt, int max_rows, int max_src, int max_nv, int max_ng) {
  auto pad = [](size_t x) { return (x + 15) & ~(size_t)15; };
  StarLayout L;
  size_t o = pad((size_t)max_nt * GS * 8);
  L.off_rec = (int)o;
  o += (size_t)max_rows * 256;
  L.off_src = (int)o;
  o += pad((size_t)max_src * 2);
  L.off_px = (int)o;
  o += pad((size_t)max_nv * 24);
  L.off_gt = (int)o;
  o += pad((size_t)max_ng * 4);
  L.total = o;
  return L;
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
  const StarLayout L = star_layout(G.max_nt, G.max_rows, G.max_src, G.max_nv, G.max_ng);
  const size_t smem = L.total;
  int ctas = std::max(1, (int)((227 * 1024) / (smem + 1024 + 256)));
  int threads = ctas >= 4 ? 256 : ctas >= 2 ? 512 : 1024;
  if (const char* env = getenv("WAE_STAR_THREADS")) threads = std::max(64, std::min(1024, atoi(env) & ~31));
  ctas = std::min(ctas, std::max(1, 1024 / threads));  // 64 registers per thread
  if (const char* env = getenv("WAE_STAR_CTAS")) ctas = std::max(1, std::min(ctas, atoi(env)));
  const int grid = std::min(G.n_patch, h->sm_count * ctas);
  auto launch = [&](auto kern) {
    CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
    if (cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess)
      cudaGetLastError();
    kern<<<grid, threads, smem, h->stream>>>(reinterpret_cast<const StarDesc*>(G.d_desc.p), G.n_patch, G.d_blob.p, G.d_pxyz.p, d_c, G.d_src.p,
                                             G.d_code.p, L.off_rec, L.off_src, L.off_px, L.off_gt, mass_scale, d_mass, d_stiff);
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
  G.max_src = SH.max_src;
  G.max_smem = SH.max_smem;
  G.n_staged = SH.n_staged;
  G.n_entities = SH.n_entities;
  G.n_sources = SH.n_sources;
  G.n_chunks = SH.n_chunks;
  G.n_pv = (int64_t)SH.gvtx.size();
  G.program_bytes = (int64_t)(SH.desc.size() * 8 + SH.blob.size() + SH.gvtx.size() * 4 + SH.src.size() * 2 + SH.code.size() * 2);
  G.d_desc.upload(SH.desc, h->stream);
  G.d_blob.upload(SH.blob.data(), SH.blob.size(), h->stream);
  G.d_gvtx.upload(SH.gvtx.data(), SH.gvtx.size(), h->stream);
  G.d_pxyz.alloc((size_t)G.n_pv * 3 + 2);
  G.d_src.upload(SH.src.data(), SH.src.size(), h->stream);
  G.d_code.upload(SH.code.data(), SH.code.size(), h->stream);
  G.xyz_version = 0;
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  return true;
}

// ---- host replay (no GPU, no context): pattern + star program of a tetrahedral mesh, then the three passes of
// assemble_tet_stars with the kernel's own arithmetic (the functions above compiled for the host).  See include/wae_b200.h.
template <int NLOC>
static void star_replay(const StarHost& G, const double* xyz, const double* c, double mass_scale, double* val_m, double* val_k,
                        std::vector<int>& written, int64_t& bad) {
  using R = StarRec<NLOC>;
  double mlut[16];
  if (NLOC == 4) wae_p1_star_mass(mlut); else wae_p2_star_mass(mlut);
  for (double& m : mlut) m *= mass_scale;
  std::vector<double> Gs((size_t)G.max_nt * GS), rec((size_t)G.max_rows * 32), px((size_t)G.max_nv * 3);
  std::vector<int> gtab(G.max_ng);
  for (int p = 0; p < G.n_patch; p++) {
    const StarDesc& D = *reinterpret_cast<const StarDesc*>(&G.desc[(size_t)p * 8]);
    const uint8_t* pb = G.blob.data() + D.blob;
    if (D.blob % 16 || (D.pxyz * 8) % 16 || (D.src * 2) % 16 || D.nsrc % 8 || D.nv % 2) bad++;
    const uint16_t* lv = reinterpret_cast<const uint16_t*>(pb);
    const int32_t* tets = reinterpret_cast<const int32_t*>(pb + ((8 * D.nt + 15) & ~15));
    const uint32_t* grp = reinterpret_cast<const uint32_t*>(pb + D.o_grp);
    const uint8_t* cnt = pb + D.o_cnt;
    const uint32_t* chunk = reinterpret_cast<const uint32_t*>(pb + D.o_cnt + 32 * D.ng);
    for (int i = 0; i < D.nv; i++)
      for (int r = 0; r < 3; r++) px[3 * i + r] = xyz[3 * (size_t)G.gvtx[D.pxyz / 3 + i] + r];
    std::fill(rec.begin(), rec.end(), std::nan(""));
    for (int t = 0; t < D.nt; t++)
      star_geometry(&px[3 * lv[4 * t]], &px[3 * lv[4 * t + 1]], &px[3 * lv[4 * t + 2]], &px[3 * lv[4 * t + 3]], c[tets[t]], true, &Gs[(size_t)t * GS]);
    for (int g = 0; g < D.ng; g++) {
      const uint32_t so = grp[2 * g], hy = grp[2 * g + 1];
      const int type = (hy >> 16) & 15, niter = hy >> 24;
      const size_t r0 = (size_t)(hy & 0xFFFFu) * 32;
      if (g >= (int)gtab.size() || r0 + (size_t)wae_star_record_rows(NLOC, type) * 32 > rec.size()) { bad++; continue; }
      gtab[g] = (int)r0;
      for (int l = 0; l < 32; l++) {
        const int n = cnt[g * 32 + l];
        if (n > niter || (int)so + 32 * niter > D.nsrc) bad++;
        if (n) star_group<NLOC, 3>(type, n, G.src.data() + D.src + so + l, Gs.data(), &rec[r0 + l]);
      }
    }
    for (int ch = 0; ch < D.nc; ch++) {
      const uint32_t z0 = chunk[2 * ch], len = chunk[2 * ch + 1] & 63u, co = chunk[2 * ch + 1] >> 6;
      if (len < 1 || len > 32 || (z0 >> 5) != ((z0 + len - 1) >> 5) || (int)(co + len) > D.ncode) { bad++; continue; }
      for (uint32_t l = 0; l < len; l++) {
        const uint16_t cw = G.code[(size_t)D.code0 + co + l];
        if ((cw >> 9) >= D.ng) { bad++; continue; }
        const double* r = &rec[(size_t)gtab[cw >> 9] + ((cw >> 4) & 31)];
        val_k[z0 + l] = r[32 * R::koff(cw & 15)];
        val_m[z0 + l] = r[0] * mlut[cw & 15];
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
    int64_t bad = 0;
    if (nloc == 4)
      star_replay<4>(G, xyz, c, 1.0, val_m, val_k, written, bad);
    else
      star_replay<10>(G, xyz, c, 1.0, val_m, val_k, written, bad);
    for (int64_t k = 0; k < P.nnz; k++) bad += written[k] != 1;
    std::copy(P.colptr.begin(), P.colptr.end(), colptr);
    std::copy(P.rowval.begin(), P.rowval.end(), rowval);
    const StarLayout L = star_layout(G.max_nt, G.max_rows, G.max_src, G.max_nv, G.max_ng);
    stats[0] = (double)P.nnz;
    stats[1] = G.n_patch;
    stats[2] = (double)G.n_staged;
    stats[3] = (double)G.n_entities;
    stats[4] = (double)G.n_sources;
    stats[5] = (double)L.total;
    stats[6] = (double)(G.desc.size() * 8 + G.blob.size() + G.gvtx.size() * 4 + G.src.size() * 2 + G.code.size() * 2);
    stats[7] = (double)bad;
    return WAE_OK;
  } catch (...) {
    return WAE_E_INVALID;
  }
}
