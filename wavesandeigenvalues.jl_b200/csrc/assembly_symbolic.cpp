// Host-side symbolic phase of the assembly: sparsity pattern, scatter map and the
// owner-computes gather program.  Replaces the triplet growth + SparseArrays.sparse()
// pattern merge of the reference (src/Helmholtz.jl:406-417,515; src/FEM/FEM.jl:22-32)
// with sort-based O(n log n) set construction.  The resulting (colptr,rowval) is the
// pattern sparse(I,J,V,dim,dim) would produce from the element triplets (explicit
// zeros kept, rows sorted inside columns).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "wae_internal.h"

#include "host_parallel.h"

namespace {
struct PhaseClock {  // WAE_SYMB_TIMING=1 prints the wall time of the symbolic phases to stderr
  std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
  bool on = getenv("WAE_SYMB_TIMING") != nullptr;
  void tick(const char* what) {
    auto n = std::chrono::steady_clock::now();
    if (on) fprintf(stderr, "[wae symbolic] %-34s %8.3f s\n", what, std::chrono::duration<double>(n - t).count());
    t = n;
  }
};
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
  // Column j holds the union of the DOFs of its elements.  One pass: every thread builds the sorted, uniquified rows of a
  // contiguous range of columns into its own buffer; the buffers are then copied behind each other (column ranges are
  // contiguous, so each buffer is one contiguous piece of rowval).
  P.dim = dim;
  P.colptr.assign(dim + 1, 0);
  const unsigned nthr = dim < 4096 ? 1u : std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  std::vector<std::vector<int32_t>> rows(nthr);
  std::vector<int64_t> cnt(dim + 1, 0);
  {
    std::vector<std::thread> th;
    for (unsigned q = 0; q < nthr; q++)
      th.emplace_back([&, q]() {
        const int64_t a = dim * q / nthr, b = dim * (q + 1) / nthr;
        std::vector<int32_t> buf;
        std::vector<int32_t>& out = rows[q];
        out.reserve((size_t)(nptr[b] - nptr[a]) * 3);
        for (int64_t j = a; j < b; j++) {
          buf.clear();
          for (int64_t r = nptr[j]; r < nptr[j + 1]; r++) {
            const uint32_t* d = conn + (size_t)elems[nadj[r]] * nloc;
            for (int k = 0; k < nloc; k++) buf.push_back((int32_t)d[k]);
          }
          std::sort(buf.begin(), buf.end());
          buf.erase(std::unique(buf.begin(), buf.end()), buf.end());
          cnt[j + 1] = (int64_t)buf.size();
          out.insert(out.end(), buf.begin(), buf.end());
        }
      });
    for (auto& x : th) x.join();
  }
  for (int64_t j = 0; j < dim; j++) P.colptr[j + 1] = P.colptr[j] + cnt[j + 1];
  P.nnz = P.colptr[dim];
  if (P.nnz >= (int64_t)1 << 31) WAE_THROW(WAE_E_INVALID, "pattern has %lld nonzeros (>= 2^31)", (long long)P.nnz);
  P.rowval.resize(P.nnz);
  {
    std::vector<std::thread> th;
    for (unsigned q = 0; q < nthr; q++)
      th.emplace_back([&, q]() {
        const int64_t a = dim * q / nthr;
        if (!rows[q].empty()) std::memcpy(P.rowval.data() + P.colptr[a], rows[q].data(), rows[q].size() * sizeof(int32_t));
        std::vector<int32_t>().swap(rows[q]);
      });
    for (auto& x : th) x.join();
  }
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
// Owner-computes pair program (tetrahedral patterns, symmetric operators M and K).
//
// Every entry (i, j) of M and K is the sum, over the elements that contain both DOFs, of one entry of the packed upper
// triangle of the (symmetric) element matrix.  No atomics: every nonzero is written exactly once, sums run in a fixed
// order (bit-reproducible), and all global stores are full, consecutive runs of the value arrays.
//
//  * elements are ranked along a Morton curve; a DOF is owned by its incident element of lowest rank; DOFs are put in
//    owner order ("positions") and cut into patches of consecutive positions.  A patch owns the COLUMNS of its DOFs and
//    stages every element touching one of them, so it sees all sources of its nonzeros.
//  * a source is (staged element t, packed index s of a local DOF pair).  If both DOFs of the pair are owned by the patch
//    the source feeds nz(i,j) and nz(j,i) at once (one summation unit), otherwise only the entry in the owned column.
//  * the units of a patch are sorted by their number of sources and cut into groups of 32 (one per lane); the k-th source
//    of lane l of a group lives in shared-memory slot  group_base + 32 k + (l xor (k mod 8)), so the summation pass reads its slots with
//    conflict-free, fully coalesced 16-byte loads and all lanes of a warp run (almost) the same trip count.  The sum is
//    left in slot group_base + l.
//  * the element pass needs, per staged element and local pair, the slot (0xFFFF: not owned): two 16-bit slots per word,
//    stored (32-element block, word, lane)-major so that a warp reads them with coalesced 128-byte loads.
//  * the store pass walks the owned columns in DOF order, 32 consecutive nonzeros per warp step: 2 bytes per nonzero
//    name the slot that holds its sum.
// Program size per tetrahedron of a P2 Kuhn mesh: 112 B per STAGED element (2.0-2.5 staged per tetrahedron), 2 B per
// nonzero, 1 B per unit -- about 350 B next to 671 B of algorithmic traffic.
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

// Morton rank of every element, DOF owner (incident element of lowest rank) and the owner order of the DOFs: shared by the pair
// program (generation 2) and the star program (generation 3, assembly_star_symbolic.cpp).
void wae_build_owner_order(const double* xyz, const uint32_t* conn, int nloc, const Pattern& P, OwnerOrder& O) {
  const int64_t ne = (int64_t)P.elems.size();
  const int64_t dim = P.dim;
  PhaseClock clk;
  std::vector<int64_t>& nptr = O.nptr;
  std::vector<int32_t>& nadj = O.nadj;
  node_to_elem(conn, nloc, P.elems, dim, nptr, nadj);
  clk.tick("node -> element adjacency");
  // Morton rank of every element (centroid on a 2^21 grid over the bounding box)
  std::vector<int32_t>& rank = O.rank;
  rank.resize(ne);
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
    parallel_sort(key);
    for (int64_t i = 0; i < ne; i++) rank[key[i].second] = (int32_t)i;
  }
  clk.tick("Morton ranks");
  // owner of a DOF = its incident element of lowest Morton rank; position = index in owner order
  std::vector<int32_t>& order = O.order;  // position -> DOF
  std::vector<int32_t>& pos = O.pos;      // DOF -> position (-1: DOF not touched by the pattern's elements)
  order.clear();
  pos.assign(dim, -1);
  int max_inc = 1;
  {
    std::vector<std::pair<int32_t, int32_t>> key;
    key.reserve(dim);
    for (int64_t i = 0; i < dim; i++) {
      if (nptr[i + 1] == nptr[i]) continue;
      int32_t best = rank[nadj[nptr[i]]];
      for (int64_t q = nptr[i] + 1; q < nptr[i + 1]; q++) best = std::min(best, rank[nadj[q]]);
      key.emplace_back(best, (int32_t)i);
      max_inc = std::max<int>(max_inc, (int)(nptr[i + 1] - nptr[i]));
    }
    parallel_sort(key);
    order.reserve(key.size());
    for (auto& k : key) {
      pos[k.second] = (int32_t)order.size();
      order.push_back(k.second);
    }
  }
  O.max_inc = max_inc;
  clk.tick("DOF owner order");
}

void wae_build_gather(const double* xyz, const uint32_t* conn, int nloc, const Pattern& P, int slot_cap, GatherHost& G) {
  const int64_t ne = (int64_t)P.elems.size();
  (void)ne;
  const int nsym = nloc * (nloc + 1) / 2;
  PhaseClock clk;
  OwnerOrder OO;
  wae_build_owner_order(xyz, conn, nloc, P, OO);
  const std::vector<int64_t>& nptr = OO.nptr;
  const std::vector<int32_t>& nadj = OO.nadj;
  const std::vector<int32_t>& rank = OO.rank;
  const std::vector<int32_t>& order = OO.order;
  const std::vector<int32_t>& pos = OO.pos;
  const int max_inc = OO.max_inc;
  if (max_inc > 255) WAE_THROW(WAE_E_INVALID, "a DOF is shared by %d elements; the pair program holds at most 255 sources per entry", max_inc);
  const int64_t npos = (int64_t)order.size();
  // ---- cut the positions into patches by their exact number of sources --------------------------------------------------
  // A patch [lo, hi) owns the columns of its positions.  Its sources are the (element, local pair {a,b}) with at least one of the
  // two DOFs owned; a pair with BOTH DOFs owned is one source feeding nz(i,j) and nz(j,i).  Adding position q to the patch
  // [lo, q) therefore adds the (element, a) combinations of q's column whose other DOF is not an earlier position of the
  // patch.  The padding of the count-sorted groups stays below 32 * (largest source count + 1) slots except for degenerate
  // patches; on overflow the cut is repeated with a wider margin.
  struct PatchOut {
    std::vector<int32_t> tets;
    std::vector<uint32_t> dest;    // per 32-element block: npk x 32 words, two 16-bit slots each (0xFFFF = not owned)
    std::vector<uint32_t> grp;     // (first slot << 8) | iterations
    std::vector<uint8_t> cnt;      // per group lane
    std::vector<uint16_t> lvtx;    // four patch-local vertex numbers per staged element
    std::vector<uint32_t> gv;      // patch-local vertex -> mesh vertex (even count: padded with a repeat of the last one)
    std::vector<uint16_t> res;     // per chunk lane: slot that holds the sum of that nonzero after the summation pass
    std::vector<uint32_t> chunk;   // pairs of words: first global nonzero, length (<= 32, inside one 32-nonzero line)
    int slots = 0;
    int64_t units = 0, sources = 0;
  };
  const int npk = (nsym + 1) / 2;  // packed slot words per element
  std::vector<int64_t> cut;
  std::vector<PatchOut> po;
  int64_t npatch = 0;
  std::atomic<int> bad(0);
  for (int attempt = 1; attempt <= 4; attempt++) {
    const int64_t raw_cap = (int64_t)slot_cap - 32 * (int64_t)(max_inc + 1) * attempt;
    if (raw_cap < 4 * (int64_t)max_inc * nloc) WAE_THROW(WAE_E_INVALID, "pair program: slot capacity %d is too small for this mesh", slot_cap);
    cut.assign(1, 0);
    {
      int64_t acc = 0;
      int32_t lo = 0;
      for (int64_t q = 0; q < npos; q++) {
        const int32_t j = order[q];
        int full = 0, back = 0;  // all (element, a) of the column / those whose other DOF is an earlier position of the patch
        for (int64_t r = nptr[j]; r < nptr[j + 1]; r++) {
          const uint32_t* d = conn + (size_t)P.elems[nadj[r]] * nloc;
          for (int a = 0; a < nloc; a++) {
            const int32_t o = pos[d[a]];
            full++;
            back += o >= lo && o < q;
          }
        }
        if (acc + full - back > raw_cap && acc > 0) {
          cut.push_back(q);
          acc = 0;
          lo = (int32_t)q;
          back = 0;
        }
        acc += full - back;
      }
      cut.push_back(npos);
    }
    npatch = (int64_t)cut.size() - 1;
    clk.tick("patch cut");
    po.clear();
    po.resize(npatch);
    bad = 0;
    parallel_for(npatch, [&](int64_t pa, int64_t pb) {
      struct Src { uint64_t key; int32_t t; int32_t s; };
      struct Unit { uint64_t key; int32_t first; int32_t cnt; int32_t slot; };
      std::vector<std::pair<int32_t, int32_t>> st;  // (rank, element)
      std::vector<Src> src, tmp;
      std::vector<Unit> units;
      std::vector<int32_t> by_cnt, cols, colptr, fill, ucol;
      for (int64_t p = pa; p < pb; p++) {
        PatchOut& O = po[p];
        const int32_t lo = (int32_t)cut[p], hi = (int32_t)cut[p + 1];
        st.clear();
        for (int32_t q = lo; q < hi; q++) {
          const int32_t j = order[q];
          for (int64_t r = nptr[j]; r < nptr[j + 1]; r++) st.emplace_back(rank[nadj[r]], nadj[r]);
        }
        std::sort(st.begin(), st.end());
        st.erase(std::unique(st.begin(), st.end()), st.end());
        const int nt = (int)st.size();
        // stage the elements by the number of local DOFs the patch owns (interior elements first, then Morton rank): the lanes
        // of a 32-element block then own (almost) the same entries, so the predicated slot stores are either full or empty
        {
          std::vector<std::pair<int32_t, std::pair<int32_t, int32_t>>> key(nt);
          for (int t = 0; t < nt; t++) {
            const uint32_t* d = conn + (size_t)P.elems[st[t].second] * nloc;
            int32_t own = 0;
            for (int a = 0; a < nloc; a++) own |= (pos[d[a]] >= lo && pos[d[a]] < hi) << a;
            key[t] = {-(int32_t)__builtin_popcount(own) * 1024 - own, st[t]};
          }
          std::sort(key.begin(), key.end());
          for (int t = 0; t < nt; t++) st[t] = key[t].second;
        }
        O.tets.resize(nt);
        O.lvtx.resize((size_t)nt * 4);
        O.gv.clear();
        for (int t = 0; t < nt; t++) {
          const uint32_t* d = conn + (size_t)P.elems[st[t].second] * nloc;
          for (int a = 0; a < 4; a++) O.gv.push_back(d[a]);
        }
        std::sort(O.gv.begin(), O.gv.end());
        O.gv.erase(std::unique(O.gv.begin(), O.gv.end()), O.gv.end());
        if (O.gv.size() >= 0xFFFF) bad = 1 << 28;
        // sources: key = (local position of the column, row DOF).  Both DOFs owned -> the column is the one of lower position.
        src.clear();
        for (int t = 0; t < nt; t++) {
          const int32_t e = st[t].second;
          O.tets[t] = e;
          const uint32_t* d = conn + (size_t)P.elems[e] * nloc;
          for (int a = 0; a < 4; a++)
            O.lvtx[(size_t)t * 4 + a] = (uint16_t)(std::lower_bound(O.gv.begin(), O.gv.end(), d[a]) - O.gv.begin());
          int32_t o[10];
          for (int a = 0; a < nloc; a++) o[a] = pos[d[a]];
          int s = 0;
          for (int a = 0; a < nloc; a++)
            for (int b = a; b < nloc; b++, s++) {
              const bool ia = o[a] >= lo && o[a] < hi, ib = o[b] >= lo && o[b] < hi;
              if (!ia && !ib) continue;
              int32_t cpos;
              uint32_t row;
              if (ia && ib) {
                cpos = std::min(o[a], o[b]);
                row = o[a] <= o[b] ? d[b] : d[a];
              } else if (ia) {
                cpos = o[a];
                row = d[b];
              } else {
                cpos = o[b];
                row = d[a];
              }
              src.push_back(Src{((uint64_t)(uint32_t)(cpos - lo) << 32) | row, t, s});
            }
        }
        O.sources = (int64_t)src.size();
        // order by (column, row, staged element): the sources were generated in ascending t, so a stable counting sort by column
        // followed by a stable sort of every (short) column bucket by row is all that is needed
        {
          const int ncol = hi - lo;
          colptr.assign(ncol + 1, 0);
          for (const Src& x : src) colptr[(x.key >> 32) + 1]++;
          for (int cc = 0; cc < ncol; cc++) colptr[cc + 1] += colptr[cc];
          tmp.resize(src.size());
          fill.assign(colptr.begin(), colptr.end() - 1);
          for (const Src& x : src) tmp[fill[x.key >> 32]++] = x;
          for (int cc = 0; cc < ncol; cc++)
            std::stable_sort(tmp.begin() + colptr[cc], tmp.begin() + colptr[cc + 1], [](const Src& x, const Src& y) { return x.key < y.key; });
          src.swap(tmp);
        }
        units.clear();
        for (size_t i = 0; i < src.size();) {
          size_t j = i;
          while (j < src.size() && src[j].key == src[i].key) j++;
          units.push_back(Unit{src[i].key, (int32_t)i, (int32_t)(j - i), -1});
          i = j;
        }
        O.units = (int64_t)units.size();
        ucol.assign(hi - lo + 1, 0);  // units are sorted by key: first unit of every column
        for (const Unit& u : units) ucol[(u.key >> 32) + 1]++;
        for (int cc = 0; cc < hi - lo; cc++) ucol[cc + 1] += ucol[cc];
        by_cnt.resize(units.size());
        for (size_t i = 0; i < units.size(); i++) by_cnt[i] = (int32_t)i;
        std::stable_sort(by_cnt.begin(), by_cnt.end(), [&](int32_t x, int32_t y) { return units[x].cnt > units[y].cnt; });
        const size_t ng = (units.size() + 31) / 32;
        O.grp.resize(ng);
        O.cnt.assign(ng * 32, 0);
        const int nwb = (nt + 31) / 32;
        O.dest.assign((size_t)nwb * npk * 32, 0xFFFFFFFFu);
        int sb = 0;
        for (size_t g = 0; g < ng; g++) {
          const int niter = units[by_cnt[g * 32]].cnt;
          O.grp[g] = ((uint32_t)sb << 8) | (uint32_t)niter;
          for (size_t l = 0; l < 32 && g * 32 + l < units.size(); l++) {
            Unit& u = units[by_cnt[g * 32 + l]];
            u.slot = sb + (int)l;
            O.cnt[g * 32 + l] = (uint8_t)u.cnt;
            for (int k = 0; k < u.cnt; k++) {
              const Src& sc = src[u.first + k];
              const uint32_t slot = (uint32_t)(sb + 32 * k + ((int)l ^ (k & 7)));  // the sources of one unit sit in 8 different bank groups
              uint32_t& w = O.dest[((size_t)(sc.t >> 5) * npk + (sc.s >> 1)) * 32 + (sc.t & 31)];
              w = (sc.s & 1) ? ((w & 0x0000FFFFu) | (slot << 16)) : ((w & 0xFFFF0000u) | slot);
            }
          }
          sb += 32 * niter;
        }
        O.slots = sb;
        if (sb > slot_cap || sb >= 0xFFFF) bad = sb;
        // store program: owned columns by DOF id (adjacent DOFs have adjacent nonzero ranges), cut into chunks of at most 32
        // consecutive nonzeros that do not straddle a 32-nonzero (256-byte) boundary of the value arrays
        cols.resize(hi - lo);
        for (int32_t q = lo; q < hi; q++) cols[q - lo] = order[q];
        std::sort(cols.begin(), cols.end());
        O.res.clear();
        O.chunk.clear();
        auto find_unit = [&](uint64_t key) -> int32_t {
          size_t a0 = (size_t)ucol[key >> 32], b0 = (size_t)ucol[(key >> 32) + 1];
          const size_t end = b0;
          while (a0 < b0) {
            size_t m = (a0 + b0) >> 1;
            if (units[m].key < key) a0 = m + 1; else b0 = m;
          }
          return (a0 < end && units[a0].key == key) ? units[a0].slot : -1;
        };
        for (int32_t j : cols) {
          const int32_t qj = pos[j];
          for (int64_t z = P.colptr[j]; z < P.colptr[j + 1]; z++) {
            const int32_t i = P.rowval[z];
            const int32_t qi = pos[i];
            const bool mirror = qi >= lo && qi < hi && qi < qj;  // summed as the pair (column i, row j)
            const int32_t sl = mirror ? find_unit(((uint64_t)(uint32_t)(qi - lo) << 32) | (uint32_t)j)
                                      : find_unit(((uint64_t)(uint32_t)(qj - lo) << 32) | (uint32_t)i);
            if (sl < 0) bad = 1 << 30;
            const size_t nc = O.chunk.size();
            if (nc && O.chunk[nc - 2] + O.chunk[nc - 1] == (uint32_t)z && (z & 31) != 0)
              O.chunk[nc - 1]++;
            else {
              O.chunk.push_back((uint32_t)z);
              O.chunk.push_back(1u);
              O.res.resize(O.res.size() + 32, 0);
            }
            O.res[O.res.size() - 32 + (O.chunk.back() - 1)] = (uint16_t)sl;
          }
        }
      }
    });
    clk.tick("per-patch programs");
    if (!bad) break;
  }
  if (bad) WAE_THROW(WAE_E_INVALID, "pair program overflow (%d; slot capacity %d)", (int)bad, slot_cap);
  // ---- pack: one descriptor + one contiguous, 16-byte aligned blob per patch (fetched with a single bulk copy) -----------
  //   blob = [ lvtx: nt x 4 u16 | tets: nt i32 | grp: ng u32 | cnt: ng x 32 u8 | chunk: nc x (u32 first, u32 length) ]
  auto pad16 = [](int64_t x) { return (x + 15) & ~(int64_t)15; };
  G = GatherHost();
  G.npk = npk;
  G.n_patch = (int)npatch;
  G.desc.assign((size_t)npatch * 8, 0);
  std::vector<int64_t> tet_ptr(npatch + 1, 0);
  int64_t blob_total = 0, pv_total = 0, wb_total = 0, chunk_total = 0;
  for (int64_t p = 0; p < npatch; p++) {
    PatchOut& O = po[p];
    if (O.gv.size() & 1) O.gv.push_back(O.gv.back());
    const int64_t nt = (int64_t)O.tets.size(), ng = (int64_t)O.grp.size(), nc = (int64_t)O.chunk.size() / 2, nv = (int64_t)O.gv.size();
    const int64_t o_tets = pad16(8 * nt), o_grp = o_tets + pad16(4 * nt), o_cnt = o_grp + pad16(4 * ng), o_chunk = o_cnt + 32 * ng;
    const int64_t bytes = o_chunk + pad16(8 * nc);
    int64_t* D = &G.desc[(size_t)p * 8];
    D[0] = blob_total;
    D[1] = pv_total * 3;  // doubles
    D[2] = wb_total;
    D[3] = chunk_total;
    int32_t* I = reinterpret_cast<int32_t*>(D + 4);
    I[0] = (int32_t)nt; I[1] = (int32_t)nv; I[2] = (int32_t)ng; I[3] = (int32_t)nc;
    I[4] = (int32_t)bytes; I[5] = (int32_t)o_tets; I[6] = (int32_t)o_grp; I[7] = (int32_t)o_cnt;
    tet_ptr[p + 1] = tet_ptr[p] + nt;
    blob_total += bytes;
    pv_total += nv;
    wb_total += (nt + 31) / 32;
    chunk_total += nc;
    G.max_slots = std::max(G.max_slots, O.slots);
    G.max_blob = std::max<int>(G.max_blob, (int)bytes);
    G.max_nv = std::max<int>(G.max_nv, (int)nv);
    G.n_pairs += O.units;
    G.n_sources += O.sources;
  }
  G.n_staged = tet_ptr[npatch];
  G.blob.resize((size_t)blob_total);  // (padding bytes between the sections stay uninitialised: never read)
  G.gvtx.resize(pv_total);
  G.dest.resize((size_t)wb_total * npk * 32);
  G.res.resize((size_t)chunk_total * 32);
  parallel_for(npatch, [&](int64_t pa, int64_t pb) {
    for (int64_t p = pa; p < pb; p++) {
      PatchOut& O = po[p];
      const int64_t* D = &G.desc[(size_t)p * 8];
      const int32_t* I = reinterpret_cast<const int32_t*>(D + 4);
      uint8_t* B = G.blob.data() + D[0];
      std::memcpy(B, O.lvtx.data(), O.lvtx.size() * 2);
      std::memcpy(B + I[5], O.tets.data(), O.tets.size() * 4);
      std::memcpy(B + I[6], O.grp.data(), O.grp.size() * 4);
      std::memcpy(B + I[7], O.cnt.data(), O.cnt.size());
      std::memcpy(B + I[7] + 32 * (int64_t)I[2], O.chunk.data(), O.chunk.size() * 4);
      std::copy(O.gv.begin(), O.gv.end(), G.gvtx.begin() + D[1] / 3);
      std::copy(O.dest.begin(), O.dest.end(), G.dest.begin() + (size_t)D[2] * npk * 32);
      std::copy(O.res.begin(), O.res.end(), G.res.begin() + (size_t)D[3] * 32);
      O = PatchOut();
    }
  });
  clk.tick("pack");
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
