// Host-side symbolic phase of the assembly: sparsity pattern, scatter map and the
// owner-computes gather program.  Replaces the triplet growth + SparseArrays.sparse()
// pattern merge of the reference (src/Helmholtz.jl:406-417,515; src/FEM/FEM.jl:22-32)
// with sort-based O(n log n) set construction.  The resulting (colptr,rowval) is the
// pattern sparse(I,J,V,dim,dim) would produce from the element triplets (explicit
// zeros kept, rows sorted inside columns).
#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>

#include "wae_internal.h"

namespace {
template <typename F>
void parallel_for(int64_t n, F f) {
  unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  if (n < 4096) nt = 1;
  std::vector<std::thread> th;
  int64_t chunk = (n + nt - 1) / nt;
  for (unsigned t = 0; t < nt; t++) {
    int64_t a = t * chunk, b = std::min<int64_t>(n, a + chunk);
    if (a >= b) break;
    th.emplace_back([=]() { f(a, b); });
  }
  for (auto& x : th) x.join();
}
}  // namespace

// node -> incident elements (positions in `elems`), CSR
static void node_to_elem(const uint32_t* conn, int nloc, const std::vector<int64_t>& elems, int64_t dim,
                         std::vector<int64_t>& ptr, std::vector<int32_t>& adj) {
  ptr.assign(dim + 1, 0);
  for (size_t e = 0; e < elems.size(); e++) {
    const uint32_t* d = conn + (size_t)elems[e] * nloc;
    for (int k = 0; k < nloc; k++) ptr[d[k] + 1]++;
  }
  for (int64_t i = 0; i < dim; i++) ptr[i + 1] += ptr[i];
  adj.resize(ptr[dim]);
  std::vector<int64_t> pos(ptr.begin(), ptr.end() - 1);
  for (size_t e = 0; e < elems.size(); e++) {
    const uint32_t* d = conn + (size_t)elems[e] * nloc;
    for (int k = 0; k < nloc; k++) adj[pos[d[k]]++] = (int32_t)e;
  }
}

void wae_build_pattern_from_elements(const uint32_t* conn, int nloc, const std::vector<int64_t>& elems,
                                     int64_t dim, Pattern& P) {
  std::vector<int64_t> nptr;
  std::vector<int32_t> nadj;
  node_to_elem(conn, nloc, elems, dim, nptr, nadj);
  // pass 1: column counts, pass 2: fill.  Column j holds the union of the DOFs of its elements.
  std::vector<int64_t> cnt(dim + 1, 0);
  std::vector<std::vector<int32_t>> cols;  // only used transiently per thread
  P.dim = dim;
  P.colptr.assign(dim + 1, 0);
  auto column = [&](int64_t j, std::vector<int32_t>& buf) {
    buf.clear();
    for (int64_t q = nptr[j]; q < nptr[j + 1]; q++) {
      const uint32_t* d = conn + (size_t)elems[nadj[q]] * nloc;
      for (int k = 0; k < nloc; k++) buf.push_back((int32_t)d[k]);
    }
    std::sort(buf.begin(), buf.end());
    buf.erase(std::unique(buf.begin(), buf.end()), buf.end());
  };
  parallel_for(dim, [&](int64_t a, int64_t b) {
    std::vector<int32_t> buf;
    for (int64_t j = a; j < b; j++) {
      column(j, buf);
      cnt[j + 1] = (int64_t)buf.size();
    }
  });
  for (int64_t j = 0; j < dim; j++) P.colptr[j + 1] = P.colptr[j] + cnt[j + 1];
  P.nnz = P.colptr[dim];
  if (P.nnz >= (int64_t)1 << 31) WAE_THROW(WAE_E_INVALID, "pattern has %lld nonzeros (>= 2^31)", (long long)P.nnz);
  P.rowval.resize(P.nnz);
  parallel_for(dim, [&](int64_t a, int64_t b) {
    std::vector<int32_t> buf;
    for (int64_t j = a; j < b; j++) {
      column(j, buf);
      std::memcpy(P.rowval.data() + P.colptr[j], buf.data(), buf.size() * sizeof(int32_t));
    }
  });
}

// slot of (row i, col j) by binary search in column j
static inline int32_t find_slot(const Pattern& P, int32_t i, int32_t j) {
  const int32_t* b = P.rowval.data() + P.colptr[j];
  const int32_t* e = P.rowval.data() + P.colptr[j + 1];
  const int32_t* it = std::lower_bound(b, e, i);
  return (int32_t)(it - P.rowval.data());
}

// slotmap[e*nloc*nloc + a*nloc + b] = nz index of (row dof[a], col dof[b])
void wae_build_slotmap(const uint32_t* conn, int nloc, const Pattern& P, std::vector<int32_t>& slotmap) {
  int64_t ne = (int64_t)P.elems.size();
  slotmap.resize((size_t)ne * nloc * nloc);
  parallel_for(ne, [&](int64_t a0, int64_t b0) {
    for (int64_t e = a0; e < b0; e++) {
      const uint32_t* d = conn + (size_t)P.elems[e] * nloc;
      int32_t* s = slotmap.data() + (size_t)e * nloc * nloc;
      for (int a = 0; a < nloc; a++)
        for (int b = 0; b < nloc; b++) s[a * nloc + b] = find_slot(P, (int32_t)d[a], (int32_t)d[b]);
    }
  });
}

// ------------------------------------------------------------------------------------
// Owner-computes gather program (tetrahedral patterns, symmetric operators M and K).
//
// The elements of the pattern are cut into spatially compact patches; every DOF column is
// owned by exactly one patch, which stages the element matrices of ALL elements touching
// its owned columns in shared memory and then produces each owned nonzero by summing its
// sources in a fixed order -- no atomics, every output written exactly once, bit-reproducible.
// Element order inside the pattern is expected to be spatially coherent (the host mirror
// sorts elements along a Morton curve before calling; any order is correct, only the halo
// factor suffers).
//
// Per patch p:
//   patch_tets [patch_tet_ptr[p]..)   positions (in P.elems) of the staged elements
//   patch_rows [patch_row_ptr[p]..)   owned DOF columns
//   for each owned column, for each of its nonzeros (in pattern order):
//       slot_cnt (u8)  number of sources,  src (u16) = tet_local*128 + a*nloc+b ... packed below
// A source is (tet_local, sym) with sym = index of (min(a,b), max(a,b)) in the packed upper
// triangle of the (symmetric) element matrix, a/b = local indices of the row / owned column
// DOF, packed as tet_local * 64 + sym (sym < 55), tet_local < 1024, so that it fits 16 bits.
// ------------------------------------------------------------------------------------
static inline uint64_t spread21(uint64_t v) {  // interleave helper: 21 bits -> every third bit
  v &= 0x1fffff;
  v = (v | v << 32) & 0x1f00000000ffffULL;
  v = (v | v << 16) & 0x1f0000ff0000ffULL;
  v = (v | v << 8) & 0x100f00f00f00f00fULL;
  v = (v | v << 4) & 0x10c30c30c30c30c3ULL;
  v = (v | v << 2) & 0x1249249249249249ULL;
  return v;
}

void wae_build_gather(const double* xyz, const uint32_t* conn, int nloc, const Pattern& P, int target_tets, GatherHost& G) {
  const int64_t ne = (int64_t)P.elems.size();
  const int64_t dim = P.dim;
  const int max_tets = 1022 < target_tets ? 1022 : target_tets;
  std::vector<int64_t> nptr;
  std::vector<int32_t> nadj;
  node_to_elem(conn, nloc, P.elems, dim, nptr, nadj);
  // Morton rank of every element (centroid on a 2^21 grid over the bounding box)
  std::vector<int32_t> rank(ne);
  {
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    std::vector<double> cen(3 * ne);
    for (int64_t e = 0; e < ne; e++) {
      const uint32_t* d = conn + (size_t)P.elems[e] * nloc;
      for (int r = 0; r < 3; r++) {
        double c = 0.25 * (xyz[3 * (size_t)d[0] + r] + xyz[3 * (size_t)d[1] + r] + xyz[3 * (size_t)d[2] + r] + xyz[3 * (size_t)d[3] + r]);
        cen[3 * e + r] = c;
        lo[r] = std::min(lo[r], c);
        hi[r] = std::max(hi[r], c);
      }
    }
    double ext = std::max({hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2], 1e-300});
    std::vector<std::pair<uint64_t, int32_t>> key(ne);
    parallel_for(ne, [&](int64_t a, int64_t b) {
      for (int64_t e = a; e < b; e++) {
        uint64_t k = 0;
        for (int r = 0; r < 3; r++) {
          uint64_t q = (uint64_t)std::min(2097151.0, (cen[3 * e + r] - lo[r]) / ext * 2097152.0);
          k |= spread21(q) << r;
        }
        key[e] = {k, (int32_t)e};
      }
    });
    std::sort(key.begin(), key.end());
    for (int64_t i = 0; i < ne; i++) rank[key[i].second] = (int32_t)i;
  }
  // owner of a DOF = its incident element of lowest Morton rank; DOFs processed in owner order
  std::vector<int32_t> order;
  {
    std::vector<std::pair<int32_t, int32_t>> key;
    key.reserve(dim);
    for (int64_t i = 0; i < dim; i++) {
      if (nptr[i + 1] == nptr[i]) continue;
      int32_t best = rank[nadj[nptr[i]]];
      for (int64_t q = nptr[i] + 1; q < nptr[i + 1]; q++) best = std::min(best, rank[nadj[q]]);
      key.emplace_back(best, (int32_t)i);
    }
    std::sort(key.begin(), key.end());
    order.reserve(key.size());
    for (auto& k : key) order.push_back(k.second);
  }
  G = GatherHost();
  G.patch_row_ptr.push_back(0);
  G.patch_tet_ptr.push_back(0);
  // pass 1: cut the owner-ordered DOF list into patches (serial, cheap)
  std::vector<int32_t> mark(ne, -1);
  std::vector<int32_t> cur;
  size_t pos = 0;
  while (pos < order.size()) {
    cur.clear();
    while (pos < order.size()) {
      int32_t dof = order[pos];
      int nnew = 0;
      for (int64_t q = nptr[dof]; q < nptr[dof + 1]; q++) nnew += mark[nadj[q]] < 0;
      if (!cur.empty() && (int)cur.size() + nnew > max_tets) break;
      if (nnew > max_tets) WAE_THROW(WAE_E_INVALID, "a DOF is shared by %d elements; gather patches hold at most %d", nnew, max_tets);
      for (int64_t q = nptr[dof]; q < nptr[dof + 1]; q++) {
        int32_t e = nadj[q];
        if (mark[e] < 0) {
          mark[e] = (int32_t)cur.size();
          cur.push_back(e);
        }
      }
      G.patch_rows.push_back(dof);
      pos++;
    }
    // stage elements in Morton order inside the patch (coherent coordinate reads)
    std::sort(cur.begin(), cur.end(), [&](int32_t a, int32_t b) { return rank[a] < rank[b]; });
    for (int32_t e : cur) {
      G.patch_tets.push_back(e);
      mark[e] = -1;
    }
    G.max_tets = std::max<int>(G.max_tets, (int)cur.size());
    G.patch_row_ptr.push_back((int64_t)G.patch_rows.size());
    G.patch_tet_ptr.push_back((int64_t)G.patch_tets.size());
  }
  // pass 2: bucketed gather program (parallel over patches).  All nonzeros owned by a patch are sorted by their number
  // of sources (descending) and cut into groups of 32 (one per lane); a group stores its sources TRANSPOSED -- for
  // iteration k the 32 codes of the 32 lanes are contiguous (0xFFFF = no source) -- so that every lane of a warp runs the
  // same trip count and the source codes are read with one coalesced load per iteration.
  const int64_t npatch = (int64_t)G.patch_row_ptr.size() - 1;
  const int code_shift = nloc == 4 ? 4 : 6;  // staging stride of the kernel: 16 (P1) / 64 (P2) doubles per element
  const uint16_t pad_code = (uint16_t)(G.max_tets << code_shift);  // element index max_tets = block of zeros in shared memory
  std::vector<std::vector<uint16_t>> psrc(npatch);
  std::vector<std::vector<int32_t>> pout(npatch);
  std::vector<std::vector<uint32_t>> pgrp(npatch);
  std::atomic<int> too_many(0);
  parallel_for(npatch, [&](int64_t pa, int64_t pb) {
    std::vector<std::pair<int32_t, int32_t>> lmap;      // (element, staged index) sorted by element
    std::vector<std::pair<int32_t, uint16_t>> contrib;  // (row, code) of one column
    struct Slot { int32_t out; int32_t first; int32_t cnt; };
    std::vector<Slot> slots;
    std::vector<uint16_t> codes;
    for (int64_t p = pa; p < pb; p++) {
      lmap.clear();
      for (int64_t t = G.patch_tet_ptr[p]; t < G.patch_tet_ptr[p + 1]; t++)
        lmap.emplace_back(G.patch_tets[t], (int32_t)(t - G.patch_tet_ptr[p]));
      std::sort(lmap.begin(), lmap.end());
      slots.clear();
      codes.clear();
      for (int64_t r = G.patch_row_ptr[p]; r < G.patch_row_ptr[p + 1]; r++) {
        int32_t col = G.patch_rows[r];
        contrib.clear();
        for (int64_t q = nptr[col]; q < nptr[col + 1]; q++) {
          int32_t e = nadj[q];
          int32_t tl = std::lower_bound(lmap.begin(), lmap.end(), std::make_pair(e, (int32_t)-1))->second;
          const uint32_t* d = conn + (size_t)P.elems[e] * nloc;
          int b = 0;
          for (int k = 0; k < nloc; k++)
            if ((int32_t)d[k] == col) b = k;
          for (int a = 0; a < nloc; a++) {
            int lo_ = a < b ? a : b, hi_ = a < b ? b : a;
            int sym = lo_ * nloc - lo_ * (lo_ - 1) / 2 + (hi_ - lo_);
            contrib.emplace_back((int32_t)d[a], (uint16_t)((tl << code_shift) + sym));
          }
        }
        std::sort(contrib.begin(), contrib.end());  // fixed summation order: by row, then by staged position
        const int32_t* rows = P.rowval.data() + P.colptr[col];
        int64_t len = P.colptr[col + 1] - P.colptr[col];
        size_t ci = 0;
        for (int64_t sidx = 0; sidx < len; sidx++) {
          Slot sl{(int32_t)(P.colptr[col] + sidx), (int32_t)codes.size(), 0};
          while (ci < contrib.size() && contrib[ci].first == rows[sidx]) {
            codes.push_back(contrib[ci].second);
            ci++;
            sl.cnt++;
          }
          slots.push_back(sl);
        }
      }
      std::stable_sort(slots.begin(), slots.end(), [](const Slot& x, const Slot& y) { return x.cnt > y.cnt; });
      // groups of GS nonzeros: lane l of the warp owns nonzeros j*32 + l (j < GS/32) of the group;
      // entry (iteration k, nonzero q) of a group is src[k*GS + q]
      const size_t GS = WAE_GATHER_GROUP;
      const size_t ng = (slots.size() + GS - 1) / GS;
      pgrp[p].resize(ng);
      pout[p].assign(ng * GS, -1);
      size_t off = 0;
      for (size_t g = 0; g < ng; g++) {
        int niter = slots[g * GS].cnt;
        if (niter > 255) too_many = niter;
        if (off >= ((size_t)1 << 24)) too_many = 1 << 24;
        pgrp[p][g] = (uint32_t)(off << 8) | (uint32_t)(niter & 255);
        psrc[p].resize(off + (size_t)niter * GS, pad_code);
        for (size_t q = 0; q < GS && g * GS + q < slots.size(); q++) {
          const Slot& sl = slots[g * GS + q];
          pout[p][g * GS + q] = sl.out;
          for (int k = 0; k < sl.cnt; k++) psrc[p][off + (size_t)k * GS + q] = codes[sl.first + k];
        }
        off += (size_t)niter * GS;
      }
    }
  });
  if (too_many) WAE_THROW(WAE_E_INVALID, "gather program overflow (%d)", (int)too_many);
  G.patch_grp_ptr.assign(npatch + 1, 0);
  G.patch_src_ptr.assign(npatch + 1, 0);
  for (int64_t p = 0; p < npatch; p++) {
    G.patch_grp_ptr[p + 1] = G.patch_grp_ptr[p] + (int64_t)pgrp[p].size();
    G.patch_src_ptr[p + 1] = G.patch_src_ptr[p] + (int64_t)psrc[p].size();
  }
  G.grp.resize(G.patch_grp_ptr[npatch]);
  G.out_idx.resize(G.patch_grp_ptr[npatch] * WAE_GATHER_GROUP);
  G.src.resize(G.patch_src_ptr[npatch]);
  parallel_for(npatch, [&](int64_t pa, int64_t pb) {
    for (int64_t p = pa; p < pb; p++) {
      std::copy(pgrp[p].begin(), pgrp[p].end(), G.grp.begin() + G.patch_grp_ptr[p]);
      std::copy(pout[p].begin(), pout[p].end(), G.out_idx.begin() + G.patch_grp_ptr[p] * WAE_GATHER_GROUP);
      std::copy(psrc[p].begin(), psrc[p].end(), G.src.begin() + G.patch_src_ptr[p]);
    }
  });
}

// ------------------------------------------------------------------------------------
// Bloch-periodic assembly (src/Bloch.jl:4-112): every element contribution (e, a, b) lands on the folded unit-cell DOFs
// (dof_new) in one of n_class matrices chosen by the reference's rule -- plain if row and column are both (or neither)
// images of the Bloch plane, "+" if only the column is, "-" if only the row is; with axis DOFs (n_class == 6) the entries
// touching an axis DOF go to the three axis variants.  n_class == 1 sums everything into one matrix (the blochified
// weighting matrix of Helmholtz.jl:545-549).  Patterns are built from the sorted unique (class, col, row) keys;
// slotmap[e*nloc^2 + a*nloc + b] = offset of the contribution in the concatenation of the class value arrays.
// ------------------------------------------------------------------------------------
void wae_build_bloch(const uint32_t* conn, int nloc, const std::vector<int64_t>& elems, int64_t dim_red, const int64_t* dof_new,
                     const uint8_t* dof_flag, int n_class, std::vector<Pattern>& P, std::vector<int32_t>& slotmap,
                     std::vector<int64_t>& class_base) {
  const int64_t ne = (int64_t)elems.size();
  const size_t nc = (size_t)ne * nloc * nloc;
  struct Key { uint64_t k; uint32_t src; };
  std::vector<Key> keys(nc);
  const uint64_t D = (uint64_t)dim_red;
  parallel_for(ne, [&](int64_t a0, int64_t b0) {
    for (int64_t e = a0; e < b0; e++) {
      const uint32_t* d = conn + (size_t)elems[e] * nloc;
      for (int a = 0; a < nloc; a++)
        for (int b = 0; b < nloc; b++) {
          uint8_t fa = dof_flag[d[a]], fb = dof_flag[d[b]];
          int cls = 0;
          if (n_class > 1) {
            bool ic = fa & 1, jc = fb & 1;
            cls = ic == jc ? 0 : (jc ? 1 : 2);
            if (n_class == 6 && ((fa | fb) & 2)) cls += 3;
          }
          uint64_t row = (uint64_t)dof_new[d[a]], col = (uint64_t)dof_new[d[b]];
          size_t q = ((size_t)e * nloc + a) * nloc + b;
          keys[q].k = ((uint64_t)cls * D + col) * D + row;
          keys[q].src = (uint32_t)q;
        }
    }
  });
  if (nc >= ((size_t)1 << 32)) WAE_THROW(WAE_E_INVALID, "Bloch assembly: too many element contributions");
  std::sort(keys.begin(), keys.end(), [](const Key& x, const Key& y) { return x.k < y.k; });
  P.clear();
  P.resize(n_class);
  for (auto& p : P) {
    p.dim = dim_red;
    p.colptr.assign(dim_red + 1, 0);
  }
  slotmap.resize(nc);
  class_base.assign(n_class + 1, 0);
  // first pass: count unique keys per class / column
  uint64_t prev = ~0ull;
  for (size_t i = 0; i < nc; i++) {
    if (keys[i].k == prev) continue;
    prev = keys[i].k;
    int cls = (int)(prev / (D * D));
    int64_t col = (int64_t)((prev / D) % D);
    P[cls].colptr[col + 1]++;
  }
  for (int c = 0; c < n_class; c++) {
    for (int64_t j = 0; j < dim_red; j++) P[c].colptr[j + 1] += P[c].colptr[j];
    P[c].nnz = P[c].colptr[dim_red];
    P[c].rowval.resize(P[c].nnz);
    class_base[c + 1] = class_base[c] + P[c].nnz;
  }
  if (class_base[n_class] >= ((int64_t)1 << 31)) WAE_THROW(WAE_E_INVALID, "Bloch assembly: pattern too large");
  std::vector<int64_t> fill(n_class, 0);
  prev = ~0ull;
  int64_t cur = -1;
  for (size_t i = 0; i < nc; i++) {
    if (keys[i].k != prev) {
      prev = keys[i].k;
      int cls = (int)(prev / (D * D));
      P[cls].rowval[fill[cls]] = (int32_t)(prev % D);
      cur = class_base[cls] + fill[cls];
      fill[cls]++;
    }
    slotmap[keys[i].src] = (int32_t)cur;
  }
}
