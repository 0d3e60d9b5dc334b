// Host-side symbolic phase of the third-generation assembly: the STAR PROGRAM.
//
// Replaces, like the pair program before it, the triplet growth + SparseArrays.sparse() duplicate summation of the reference
// (src/Helmholtz.jl:405-445,515; src/FEM/FEM.jl:22-43,704-738,1745-1874) for the symmetric tetrahedral operators M and K.
//
// Idea.  Every nonzero (p, q) of M and K belongs to the sub-simplex sigma spanned by the mesh vertices of its two DOFs: a vertex,
// an edge, a face or the tetrahedron itself (P1: a vertex or an edge).  Its sources are exactly the elements of the star of sigma,
// and -- because the Lagrange bases are invariant under vertex permutations -- the element entry K_e[p, q] is the same linear
// combination of the element's gram entries  g_xy = -c^2 |det| grad(l_x).grad(l_y)  (x, y vertices of sigma) in every element of the
// star, and M_e[p, q] = m_role |det|.  So instead of moving 100 element entries per P2 tetrahedron through shared memory
// (generation 2), the kernel sums g_xy and |det| ONCE per sub-simplex over its star, in registers, and forms all nonzeros of the
// sub-simplex ("roles": 1 per vertex, 4 per edge, 6 per face, 3 per tetrahedron, each written to (p, q) and (q, p)) from those sums.
// Per P2 tetrahedron of a Kuhn mesh that is 15 star sources instead of 74 slot sources, and one 8-byte record entry per role.
//
//  * elements are ranked on a Morton curve, a DOF is owned by its incident element of lowest rank, DOFs in owner order are cut
//    into patches (as in generation 2).  A patch owns the COLUMNS of its DOFs, stages every element touching one of them
//    (geometry pass: gram matrix + |det| per staged element in shared memory) and runs every sub-simplex that has at least one
//    nonzero in an owned column; all elements of such a star are staged, because they all contain the owned DOF.
//  * a source word (16 bits) is (staged element << 2 n) | the n local vertex numbers of the sub-simplex in canonical order (vertices
//    sorted by mesh number, so every element of the star agrees on the frame of the roles); a tetrahedron is its own star and uses
//    the element's local frame (word = staged element).  Patches stage fewer than 1024 (P2) / 4096 (P1) elements.
//  * the sub-simplices of a patch are sorted by type and source count and cut into groups of 32 (one per lane); source k of lane l
//    is word  group_base + 32 k + l  (coalesced).  The records of a group are stored entry-major: row 0 holds the 32 sums of |det|,
//    row 1 + j the 32 values of role j (conflict-free 8-byte stores).
//  * the store pass walks the owned columns in DOF order, 32 consecutive nonzeros per warp step; 2 bytes per nonzero name the
//    record entry of its K value (13 bits: row * 33 + lane; a patch has at most 248 record rows = 64 KB) and the distance to the row
//    of its M value (3 bits); codes are stored back to back, a chunk header holds the first nonzero, the length and the offset of
//    its codes.  Every nonzero is written exactly once, in a fixed summation order.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>

#include "assembly_star.h"
#include "wae_internal.h"

#include "host_parallel.h"

namespace {
struct Src {
  uint64_t a, b;  // canonical vertex numbers: a = v0 << 32 | v1, b = v2 << 32 | v3 (unused = 0xFFFFFFFF)
  int32_t type;   // number of vertices - 1
  uint32_t word;
};
struct Ent {
  int32_t type, first, cnt, slot;  // slot = first record row of the group * 32 + lane
};
struct PatchOut {
  std::vector<int32_t> tets;
  std::vector<uint16_t> lvtx;
  std::vector<uint32_t> gv;
  std::vector<uint32_t> grp;   // two words per group
  std::vector<uint8_t> cnt;    // 32 per group
  std::vector<uint16_t> src;   // group-major source words
  std::vector<uint32_t> chunk; // pairs: first global nonzero, length | code offset << 6
  std::vector<uint16_t> code;  // one per owned nonzero, in chunk order
  int rows = 0;  // record rows (32 doubles each)
  int64_t entities = 0, sources = 0;
  int bad = 0;
};
const int EIDX[4][4] = {{-1, 0, 1, 2}, {0, -1, 3, 4}, {1, 3, -1, 5}, {2, 4, 5, -1}};  // local edge number of two local vertices
}  // namespace

// a source word has 16 bits: staged element << 2 n | n local vertex numbers (n <= 3)
int wae_star_max_staged(int nloc) { return nloc == 4 ? 4095 : 1023; }

static void build_patch(const uint32_t* conn, int nloc, const Pattern& P, const OwnerOrder& OO, int32_t lo, int32_t hi, PatchOut& O) {
  const std::vector<int64_t>& nptr = OO.nptr;
  const std::vector<int32_t>& nadj = OO.nadj;
  const std::vector<int32_t>& rank = OO.rank;
  const std::vector<int32_t>& order = OO.order;
  const std::vector<int32_t>& pos = OO.pos;
  const bool p2 = nloc == 10;
  auto owned = [&](uint32_t dof) { const int32_t o = pos[dof]; return o >= lo && o < hi; };
  // staged elements: everything touching an owned DOF, in Morton order
  std::vector<std::pair<int32_t, int32_t>> st;
  for (int32_t q = lo; q < hi; q++) {
    const int32_t j = order[q];
    for (int64_t r = nptr[j]; r < nptr[j + 1]; r++) st.emplace_back(rank[nadj[r]], nadj[r]);
  }
  std::sort(st.begin(), st.end());
  st.erase(std::unique(st.begin(), st.end()), st.end());
  const int nt = (int)st.size();
  O.tets.resize(nt);
  O.lvtx.resize((size_t)nt * 4);
  O.gv.clear();
  for (int t = 0; t < nt; t++) {
    const uint32_t* d = conn + (size_t)P.elems[st[t].second] * nloc;
    for (int a = 0; a < 4; a++) O.gv.push_back(d[a]);
  }
  std::sort(O.gv.begin(), O.gv.end());
  O.gv.erase(std::unique(O.gv.begin(), O.gv.end()), O.gv.end());
  if (O.gv.size() >= 0xFFFF || nt > wae_star_max_staged(nloc)) O.bad |= 1;
  // sources of every sub-simplex with a nonzero in an owned column
  std::vector<Src> src;
  src.reserve((size_t)nt * 15);
  for (int t = 0; t < nt; t++) {
    const int32_t e = st[t].second;
    O.tets[t] = e;
    const uint32_t* d = conn + (size_t)P.elems[e] * nloc;
    for (int a = 0; a < 4; a++) O.lvtx[(size_t)t * 4 + a] = (uint16_t)(std::lower_bound(O.gv.begin(), O.gv.end(), d[a]) - O.gv.begin());
    bool ow[10];
    for (int a = 0; a < nloc; a++) ow[a] = owned(d[a]);
    for (int m = 1; m < 16; m++) {
      int lv[4], n = 0;
      for (int a = 0; a < 4; a++)
        if (m >> a & 1) lv[n++] = a;
      if (!p2 && n > 2) continue;
      bool need = false;
      if (n < 4)
        for (int i = 0; i < n; i++) need |= ow[lv[i]];
      if (p2)
        for (int i = 0; i < n; i++)
          for (int j = i + 1; j < n; j++) need |= ow[4 + EIDX[lv[i]][lv[j]]];
      if (!need) continue;
      // canonical frame: vertices by mesh number (a tetrahedron has one source: the element's own frame)
      uint32_t v[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
      uint32_t w = (uint32_t)t;
      if (n < 4) {
        for (int i = 1; i < n; i++)
          for (int j = i; j > 0 && d[lv[j]] < d[lv[j - 1]]; j--) std::swap(lv[j], lv[j - 1]);
        w <<= 2 * n;
        for (int i = 0; i < n; i++) w |= (uint32_t)lv[i] << (2 * i);
      }
      for (int i = 0; i < n; i++) v[i] = d[lv[i]];
      if (n == 4) std::sort(v, v + 4);  // key only
      src.push_back(Src{((uint64_t)v[0] << 32) | v[1], ((uint64_t)v[2] << 32) | v[3], n - 1, w});
    }
  }
  O.sources = (int64_t)src.size();
  // generated in ascending staged element: a stable sort by (type, vertices) keeps the summation order = Morton order
  std::stable_sort(src.begin(), src.end(), [](const Src& x, const Src& y) {
    if (x.type != y.type) return x.type < y.type;
    if (x.a != y.a) return x.a < y.a;
    return x.b < y.b;
  });
  std::vector<Ent> ent;
  for (size_t i = 0; i < src.size();) {
    size_t j = i;
    while (j < src.size() && src[j].type == src[i].type && src[j].a == src[i].a && src[j].b == src[i].b) j++;
    if (j - i > 255) O.bad |= 2;
    ent.push_back(Ent{src[i].type, (int32_t)i, (int32_t)(j - i), -1});
    i = j;
  }
  O.entities = (int64_t)ent.size();
  std::vector<int32_t> by(ent.size());
  for (size_t i = 0; i < ent.size(); i++) by[i] = (int32_t)i;
  // by type, then by source count (lanes of a group run the same trip count), then by the first source element: neighbouring lanes
  // read neighbouring gram blocks
  auto first_elem = [&](const Ent& E) { return E.type == 3 ? src[E.first].word : src[E.first].word >> (2 * (E.type + 1)); };
  std::stable_sort(by.begin(), by.end(), [&](int32_t x, int32_t y) {
    if (ent[x].type != ent[y].type) return ent[x].type < ent[y].type;
    if (ent[x].cnt != ent[y].cnt) return ent[x].cnt > ent[y].cnt;
    return first_elem(ent[x]) < first_elem(ent[y]);
  });
  // ---- store program first (it depends on the owned columns only): owned columns by DOF number, chunks of at most 32 consecutive
  // nonzeros that do not straddle a 32-nonzero (256-byte) boundary of the value arrays; hc[i] = half-warp (2 * chunk + lane / 16) that
  // will read the record entry of local nonzero i
  std::vector<int32_t> cols(hi - lo);
  for (int32_t q = lo; q < hi; q++) cols[q - lo] = order[q];
  std::sort(cols.begin(), cols.end());
  std::vector<int64_t> coloff(cols.size() + 1, 0);
  for (size_t c = 0; c < cols.size(); c++) coloff[c + 1] = coloff[c] + (P.colptr[cols[c] + 1] - P.colptr[cols[c]]);
  const size_t nnz_own = (size_t)coloff.back();
  std::vector<int32_t> hc(nnz_own);
  O.chunk.clear();
  for (size_t c = 0; c < cols.size(); c++) {
    const int32_t j = cols[c];
    for (int64_t z = P.colptr[j]; z < P.colptr[j + 1]; z++) {
      const size_t nc = O.chunk.size();
      if (nc && O.chunk[nc - 2] + (O.chunk[nc - 1] & 63u) == (uint32_t)z && (z & 31) != 0)
        O.chunk[nc - 1]++;
      else {
        O.chunk.push_back((uint32_t)z);
        O.chunk.push_back(1u | ((uint32_t)(coloff[c] + (z - P.colptr[j])) << 6));  // codes are stored in the same order
      }
      const size_t ch = O.chunk.size() / 2 - 1;
      hc[(size_t)coloff[c] + (z - P.colptr[j])] = (int32_t)(2 * ch + (((O.chunk[2 * ch + 1] & 63u) - 1) >> 4));
    }
  }
  const size_t n_half = O.chunk.size();  // 2 per chunk
  // ---- the nonzeros every sub-simplex feeds: (local nonzero, role), owned columns only
  std::vector<int32_t> enz_ptr(ent.size() + 1, 0);
  std::vector<std::pair<int32_t, int32_t>> enz;
  enz.reserve(nnz_own);
  {
    auto put = [&](uint32_t prow, uint32_t qcol, int role) {
      if (!owned(qcol)) return;
      const size_t c = std::lower_bound(cols.begin(), cols.end(), (int32_t)qcol) - cols.begin();
      const int32_t* b = P.rowval.data() + P.colptr[qcol];
      const int32_t* e = P.rowval.data() + P.colptr[qcol + 1];
      const int32_t* it = std::lower_bound(b, e, (int32_t)prow);
      if (it == e || *it != (int32_t)prow) { O.bad |= 16; return; }
      enz.emplace_back((int32_t)(coloff[c] + (it - b)), role);
    };
    auto put2 = [&](uint32_t p_, uint32_t q_, int role) {
      put(p_, q_, role);
      if (p_ != q_) put(q_, p_, role);
    };
    for (size_t ei = 0; ei < ent.size(); ei++) {
      const Ent& E = ent[ei];
      const Src& s0 = src[E.first];
      const int n = E.type + 1;
      const int t = (int)(n < 4 ? s0.word >> (2 * n) : s0.word);
      const uint32_t* d = conn + (size_t)P.elems[st[t].second] * nloc;
      int l[4] = {0, 1, 2, 3};
      if (n < 4)
        for (int i = 0; i < n; i++) l[i] = (s0.word >> (2 * i)) & 3;
      const uint32_t va = d[l[0]];
      if (E.type == 0) {
        put2(va, va, 0);
      } else if (E.type == 1) {
        const uint32_t vb = d[l[1]];
        put2(va, vb, 1);
        if (p2) {
          const uint32_t eab = d[4 + EIDX[l[0]][l[1]]];
          put2(va, eab, 2);
          put2(vb, eab, 3);
          put2(eab, eab, 4);
        }
      } else if (E.type == 2) {
        const uint32_t vb = d[l[1]], vc = d[l[2]];
        const uint32_t eab = d[4 + EIDX[l[0]][l[1]]], eac = d[4 + EIDX[l[0]][l[2]]], ebc = d[4 + EIDX[l[1]][l[2]]];
        put2(vc, eab, 5);
        put2(vb, eac, 6);
        put2(va, ebc, 7);
        put2(eab, eac, 8);
        put2(eab, ebc, 9);
        put2(eac, ebc, 10);
      } else {
        const uint32_t eab = d[4 + EIDX[l[0]][l[1]]], eac = d[4 + EIDX[l[0]][l[2]]], ead = d[4 + EIDX[l[0]][l[3]]];
        const uint32_t ebc = d[4 + EIDX[l[1]][l[2]]], ebd = d[4 + EIDX[l[1]][l[3]]], ecd = d[4 + EIDX[l[2]][l[3]]];
        put2(eab, ecd, 11);
        put2(eac, ebd, 12);
        put2(ead, ebc, 13);
      }
      enz_ptr[ei + 1] = (int32_t)enz.size();
    }
  }
  // ---- groups of 32 lanes of one type.  Which 16 simplices share a half-warp is fixed by the order above; their lanes inside the
  // half-warp are chosen greedily so that the record entries one half-warp of the store pass reads (hc) fall into different
  // shared-memory bank pairs: occ[half][bank] counts the entries placed so far.  WAE_STAR_NO_LANE_OPT=1 keeps the plain order.
  static const bool lane_opt = getenv("WAE_STAR_NO_LANE_OPT") == nullptr;
  std::vector<uint8_t> occ(n_half * 32, 0);  // [half][0..15] K entries, [half][16..31] M entries
  O.grp.clear();
  O.cnt.clear();
  O.src.clear();
  int rows = 0;
  for (size_t i = 0; i < by.size();) {
    const int type = ent[by[i]].type;
    const int split = type == 0 ? WAE_STAR_VSPLIT : 1, per = 32 / split;  // simplices per group
    size_t j = i;
    while (j < by.size() && j < i + per && ent[by[j]].type == type) j++;
    const int niter = (ent[by[i]].cnt + split - 1) / split;
    const size_t sb = O.src.size();
    O.src.resize(sb + (size_t)32 * niter, (uint16_t)0);
    O.cnt.resize(O.cnt.size() + 32, 0);
    O.grp.push_back((uint32_t)sb);
    O.grp.push_back((uint32_t)rows | ((uint32_t)type << 16) | ((uint32_t)niter << 24));
    for (int half = 0; half < 2; half++) {
      const size_t h0 = i + (size_t)half * (per / 2), h1 = std::min(j, h0 + per / 2);
      bool used[16] = {false};
      // simplices with many nonzeros choose first
      std::vector<size_t> ord;
      for (size_t x = h0; x < h1; x++) ord.push_back(x);
      std::stable_sort(ord.begin(), ord.end(), [&](size_t x, size_t y) {
        return enz_ptr[by[x] + 1] - enz_ptr[by[x]] > enz_ptr[by[y] + 1] - enz_ptr[by[y]];
      });
      for (size_t x : ord) {
        const int ei = by[x];
        Ent& E = ent[ei];
        int best = -1;
        long bestc = 0;
        for (int r = 0; r < 16; r += split) {
          if (used[r]) continue;
          if (!lane_opt) {
            if (r == (int)((x - h0) * split)) best = r;
            continue;
          }
          long cst = 0;
          for (int32_t q = enz_ptr[ei]; q < enz_ptr[ei + 1]; q++) {
            const int role = enz[q].second;
            const uint8_t* oc = &occ[(size_t)hc[enz[q].first] * 32];
            cst += oc[((rows + star_krow(nloc, role)) * WAE_STAR_RS + r) & 15] + oc[16 + (((rows + star_mrow(nloc, role)) * WAE_STAR_RS + r) & 15)];
          }
          if (best < 0 || cst < bestc) best = r, bestc = cst;
        }
        used[best] = true;
        const int lane0 = 16 * half + best;
        E.slot = rows * WAE_STAR_RS + lane0;
        for (int32_t q = enz_ptr[ei]; q < enz_ptr[ei + 1]; q++) {
          const int role = enz[q].second;
          uint8_t* oc = &occ[(size_t)hc[enz[q].first] * 32];
          oc[((rows + star_krow(nloc, role)) * WAE_STAR_RS + lane0) & 15]++;
          oc[16 + (((rows + star_mrow(nloc, role)) * WAE_STAR_RS + lane0) & 15)]++;
        }
        const int q = (E.cnt + split - 1) / split;  // sources per lane of a split star
        for (int sl = 0; sl < split; sl++) {
          const int k0 = sl * q, k1 = std::min(E.cnt, k0 + q), lane = lane0 + sl;
          O.cnt[O.cnt.size() - 32 + lane] = (uint8_t)std::max(0, k1 - k0);
          for (int k = k0; k < k1; k++) {
            if (src[E.first + k].word > 0xFFFFu) O.bad |= 1;
            O.src[sb + (size_t)32 * (k - k0) + lane] = (uint16_t)src[E.first + k].word;
          }
        }
      }
    }
    // ---- order of the sources inside every lane: in iteration k the 16 lanes of a half-warp issue the same loads (one per gram entry of
    // the sub-simplex type) at data-dependent addresses; two lanes that hit the same bank pair with different addresses cost a second
    // wavefront.  The sum over a star does not depend on the order, so every lane walks its sources in the order that a greedy schedule
    // (iteration by iteration, lane by lane, cheapest remaining source first) finds -- fixed once per pattern, so the summation order is
    // still fixed.  WAE_STAR_NO_SRC_OPT=1 keeps the element order.
    static const bool src_opt = getenv("WAE_STAR_NO_SRC_OPT") == nullptr;
    if (src_opt && niter > 1) {
      const uint8_t* cn = &O.cnt[O.cnt.size() - 32];
      for (int half = 0; half < 2; half++) {
        std::vector<uint16_t> rem[16];
        for (int l = 0; l < 16; l++)
          for (int k = 0; k < cn[16 * half + l]; k++) rem[l].push_back(O.src[sb + (size_t)32 * k + 16 * half + l]);
        for (int k = 0; k < niter; k++) {
          // occupancy of this iteration: per load j and bank pair the distinct addresses placed so far
          int nadr[7][16] = {{0}}, adr[7][16][16], worst[7] = {0};
          // lanes with the fewest remaining choices first (they cannot avoid anything later)
          int lanes[16], nl = 0;
          for (int l = 0; l < 16; l++)
            if (!rem[l].empty()) lanes[nl++] = l;
          std::stable_sort(lanes, lanes + nl, [&](int a, int b) { return rem[a].size() < rem[b].size(); });
          for (int q = 0; q < nl; q++) {
            const int l = lanes[q];
            int best = -1;
            long bestc = 0;
            for (size_t cand = 0; cand < rem[l].size(); cand++) {
              int wd[7];
              const int n = star_load_words(nloc, type, rem[l][cand], wd);
              long cst = 0;
              for (int jj = 0; jj < n; jj++) {
                const int b = wd[jj] & 15;
                bool dup = false;
                for (int z = 0; z < nadr[jj][b]; z++) dup |= adr[jj][b][z] == wd[jj];
                if (dup) continue;
                cst += nadr[jj][b] + 1 > worst[jj] ? 1000 : 0;  // a new wavefront of load jj
                cst += nadr[jj][b];                              // otherwise prefer the emptier bank pair
              }
              if (best < 0 || cst < bestc) best = (int)cand, bestc = cst;
            }
            const uint16_t wsel = rem[l][best];
            rem[l].erase(rem[l].begin() + best);
            O.src[sb + (size_t)32 * k + 16 * half + l] = wsel;
            int wd[7];
            const int n = star_load_words(nloc, type, wsel, wd);
            for (int jj = 0; jj < n; jj++) {
              const int b = wd[jj] & 15;
              bool dup = false;
              for (int z = 0; z < nadr[jj][b]; z++) dup |= adr[jj][b][z] == wd[jj];
              if (!dup) {
                adr[jj][b][nadr[jj][b]++] = wd[jj];
                worst[jj] = std::max(worst[jj], nadr[jj][b]);
              }
            }
          }
        }
      }
    }
    rows += star_rows(nloc, type);
    i = j;
  }
  O.rows = rows;
  if (rows * WAE_STAR_RS > 8192) O.bad |= 8;  // 13 bits of a code word name the record entry
  // ---- codes, in the order of the owned nonzeros
  O.code.assign(nnz_own, (uint16_t)0xFFFF);
  for (size_t ei = 0; ei < ent.size(); ei++)
    for (int32_t q = enz_ptr[ei]; q < enz_ptr[ei + 1]; q++) {
      const int role = enz[q].second;
      uint16_t& f = O.code[enz[q].first];
      if (f != 0xFFFF) O.bad |= 32;
      f = (uint16_t)(((ent[ei].slot + WAE_STAR_RS * star_krow(nloc, role)) << 3) | (star_mrow(nloc, role) - star_krow(nloc, role)));
    }
  for (uint16_t f : O.code)
    if (f == 0xFFFF) O.bad |= 64;
  while ((O.chunk.size() / 2) & 3) {  // the store pass takes four chunks per warp step
    O.chunk.push_back(0u);
    O.chunk.push_back(0u);
  }
}

static inline int64_t pad16(int64_t x) { return (x + 15) & ~(int64_t)15; }

// shared memory one CTA needs for a patch: gram + |det| of the staged elements (17 doubles), the records, blob A + the patch's vertex
// coordinates, blob B, the group table
static int64_t patch_smem(int nloc, const PatchOut& O) {
  (void)nloc;
  const StarBlob L((int)O.tets.size(), (int)O.grp.size() / 2, (int)O.src.size(), (int)O.chunk.size() / 2, (int)O.code.size());
  return pad16((int64_t)O.tets.size() * 17 * 8) + pad16((int64_t)O.rows * WAE_STAR_RS * 8) + L.bytesA + L.bytesB + pad16(((int64_t)O.gv.size() + 1) * 24) +
         512;  // + descriptor ring, mbarriers
}

void wae_build_star(const double* xyz, const uint32_t* conn, int nloc, const Pattern& P, int64_t smem_budget, StarHost& G) {
  const bool timing = getenv("WAE_SYMB_TIMING") != nullptr;
  auto t0 = std::chrono::steady_clock::now();
  auto tick = [&](const char* what) {
    auto n = std::chrono::steady_clock::now();
    if (timing) fprintf(stderr, "[wae symbolic] %-34s %8.3f s\n", what, std::chrono::duration<double>(n - t0).count());
    t0 = n;
  };
  OwnerOrder OO;
  wae_build_owner_order(xyz, conn, nloc, P, OO);
  const int64_t ne = (int64_t)P.elems.size(), npos = (int64_t)OO.order.size();
  if (OO.max_inc > 255) WAE_THROW(WAE_E_INVALID, "a DOF is shared by %d elements; the star program holds at most 255 sources per star", OO.max_inc);
  // ---- two ways to cut the DOFs into patches:
  //  (M, default) consecutive ranges of the Morton owner order of generation 2, cut by their exact number of staged elements: uniform use
  //      of the shared-memory budget; compact when the cells line up with the curve's power-of-two boxes (64^3 grid: 2.35 staged elements per
  //      tetrahedron), ragged otherwise (203^3: 3.2);
  //  (B, WAE_STAR_CUT=bisect) leaves of a recursive coordinate bisection of the DOF positions (vertices; P2: edge midpoints): a range of k
  //      leaves is split at the k/2 : k - k/2 quantile of its longest axis -- box-shaped clusters of equal DOF count whatever the mesh (203^3:
  //      2.86 staged), but leaves of equal DOF count need very different amounts of shared memory (44 .. 114 KB), so the patches are smaller
  //      on average and the kernel is not faster for it (config 5: 14.4 ms against 14.0 ms) while the host phase takes three times as long.
  //  WAE_STAR_CUT=auto builds both cuts and keeps the one that stages fewer elements per owned DOF on a sample of patches.
  std::vector<int64_t> cut;
  const std::vector<int32_t> order_m = OO.order, pos_m = OO.pos;
  const char* cut_env = getenv("WAE_STAR_CUT");
  std::vector<float> dpos;
  if (cut_env && cut_env[0] != 'm') {
    dpos.assign((size_t)3 * P.dim, 0.0f);
    static const int EVL[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
    parallel_for(ne, [&](int64_t e0, int64_t e1) {
      for (int64_t e = e0; e < e1; e++) {
        const uint32_t* d = conn + (size_t)P.elems[e] * nloc;
        for (int a2 = 0; a2 < 4; a2++)
          for (int r = 0; r < 3; r++) dpos[3 * (size_t)d[a2] + r] = (float)xyz[3 * (size_t)d[a2] + r];
        if (nloc == 10)
          for (int k = 0; k < 6; k++)
            for (int r = 0; r < 3; r++)
              dpos[3 * (size_t)d[4 + k] + r] = (float)(0.5 * (xyz[3 * (size_t)d[EVL[k][0]] + r] + xyz[3 * (size_t)d[EVL[k][1]] + r]));
      }
    });
  }
  std::function<void(int32_t*, int64_t, int64_t, int)> bisect = [&](int32_t* idx, int64_t n, int64_t k, int depth) {
    if (k <= 1 || n <= 1) return;
    float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
    for (int64_t i = 0; i < n; i++)
      for (int r = 0; r < 3; r++) {
        const float v = dpos[3 * (size_t)idx[i] + r];
        lo[r] = std::min(lo[r], v);
        hi[r] = std::max(hi[r], v);
      }
    int ax = 0;
    for (int r = 1; r < 3; r++)
      if (hi[r] - lo[r] > hi[ax] - lo[ax]) ax = r;
    const int64_t kl = k / 2, nl = n * kl / k;
    std::nth_element(idx, idx + nl, idx + n, [&](int32_t x, int32_t y) {
      const float a2 = dpos[3 * (size_t)x + ax], b2 = dpos[3 * (size_t)y + ax];
      return a2 < b2 || (a2 == b2 && x < y);
    });
    if (depth < 4 && n > 200000) {  // the top of the tree runs on sixteen threads
      std::thread t([&] { bisect(idx, nl, kl, depth + 1); });
      bisect(idx + nl, n - nl, k - kl, depth + 1);
      t.join();
    } else {
      bisect(idx, nl, kl, depth + 1);
      bisect(idx + nl, n - nl, k - kl, depth + 1);
    }
  };
  std::vector<int32_t> stamp;
  // cap: staged elements per patch (Morton) / DOFs per patch (bisection)
  auto do_cut = [&](bool morton, int64_t cap) {
    cut.assign(1, 0);
    if (!morton) {
      const int64_t k = std::max<int64_t>(1, (npos + cap - 1) / cap);
      OO.order = order_m;
      bisect(OO.order.data(), npos, k, 0);
      for (int64_t q = 0; q < npos; q++) OO.pos[OO.order[q]] = (int32_t)q;
      std::function<void(int64_t, int64_t, int64_t)> bounds = [&](int64_t lo, int64_t n, int64_t kk) {  // the recursion's own splits
        if (kk <= 1 || n <= 1) {
          if (n > 0) cut.push_back(lo + n);
          return;
        }
        const int64_t kl = kk / 2, nl = n * kl / kk;
        bounds(lo, nl, kl);
        bounds(lo + nl, n - nl, kk - kl);
      };
      bounds(0, npos, k);
      return;
    }
    OO.order = order_m;
    OO.pos = pos_m;
    cap = std::min<int64_t>(cap, wae_star_max_staged(nloc));
    stamp.assign(ne, -1);
    int64_t staged = 0;
    int32_t cur = 0;
    for (int64_t q = 0; q < npos; q++) {
      const int32_t j = OO.order[q];
      int64_t fresh = 0;
      for (int64_t r = OO.nptr[j]; r < OO.nptr[j + 1]; r++) fresh += stamp[OO.nadj[r]] != cur;
      if (staged + fresh > cap && staged > 0) {
        cut.push_back(q);
        cur++;
        staged = 0;
        fresh = OO.nptr[j + 1] - OO.nptr[j];
      }
      for (int64_t r = OO.nptr[j]; r < OO.nptr[j + 1]; r++) stamp[OO.nadj[r]] = cur;
      staged += fresh;
    }
    cut.push_back(npos);
  };
  // a sample of patches of the current cut: largest shared memory per unit of the cut, staged elements per owned DOF
  auto sample = [&](bool morton, double& per_unit, double& staged_per_dof) {
    const int64_t np = (int64_t)cut.size() - 1, ns = std::min<int64_t>(np, 32);
    std::vector<double> pers(ns, 0.0), st(ns, 0.0), dofs(ns, 0.0);
    parallel_for(ns, [&](int64_t a2, int64_t b2) {
      for (int64_t i = a2; i < b2; i++) {
        const int64_t p = i * np / ns;
        PatchOut O;
        build_patch(conn, nloc, P, OO, (int32_t)cut[p], (int32_t)cut[p + 1], O);
        const double unit = morton ? (double)std::max<size_t>(1, O.tets.size()) : (double)std::max<int64_t>(1, cut[p + 1] - cut[p]);
        pers[i] = (double)patch_smem(nloc, O) / unit;
        st[i] = (double)O.tets.size();
        dofs[i] = (double)(cut[p + 1] - cut[p]);
      }
    }, 2);
    per_unit = 0;
    double s1 = 0, s2 = 0;
    for (int64_t i = 0; i < ns; i++) {
      per_unit = std::max(per_unit, pers[i]);
      s1 += st[i];
      s2 += dofs[i];
    }
    staged_per_dof = s1 / std::max(s2, 1.0);
  };
  const int64_t cap_min_of[2] = {2 * (int64_t)OO.max_inc, 4};
  int64_t cap_of[2] = {0, 0};
  double ratio_of[2] = {1e300, 1e300};
  const char* force = getenv("WAE_STAR_CUT");
  if (!force) force = "morton";
  for (int mode = 0; mode < 2; mode++) {  // 0: Morton, 1: bisection
    if (force[0] != 'a' && ((mode == 0) != (force[0] == 'm'))) continue;
    const bool morton = mode == 0;
    const double guess = morton ? (nloc == 4 ? 220.0 : 420.0) : (nloc == 4 ? 2600.0 : 1150.0);  // shared memory per unit, first guess
    double per = 0, ratio = 0;
    do_cut(morton, std::max<int64_t>(cap_min_of[mode], (int64_t)(smem_budget / guess)));
    sample(morton, per, ratio);
    cap_of[mode] = std::max<int64_t>(cap_min_of[mode], (int64_t)(0.95 * smem_budget / std::max(per, 1.0)));
    do_cut(morton, cap_of[mode]);
    sample(morton, per, ratio);
    ratio_of[mode] = ratio;
  }
  const bool morton = ratio_of[0] <= ratio_of[1];
  int64_t cap_nt = cap_of[morton ? 0 : 1];
  const int64_t cap_min = cap_min_of[morton ? 0 : 1];
  if (timing && force[0] == 'a')
    fprintf(stderr, "[wae symbolic] star: staged elements per owned DOF on the samples: Morton %.3f, bisection %.3f -> %s\n", ratio_of[0], ratio_of[1],
            morton ? "Morton" : "bisection");
  tick("star: cut + calibration");
  std::vector<PatchOut> po;
  int64_t npatch = 0;
  for (int attempt = 0;; attempt++) {
    do_cut(morton, cap_nt);
    npatch = (int64_t)cut.size() - 1;
    po.clear();
    po.resize(npatch);
    parallel_for(npatch, [&](int64_t a, int64_t b) {
      for (int64_t p = a; p < b; p++) build_patch(conn, nloc, P, OO, (int32_t)cut[p], (int32_t)cut[p + 1], po[p]);
    }, 16);
    int bad = 0;
    int64_t worst = 0;
    int worst_rows = 0;
    {  // the kernel's buffers are sized by the largest patch of each section
      int m_nt = 0, m_a = 0, m_b = 0, m_nv = 0;
      for (auto& O : po) {
        bad |= O.bad & ~8;
        const StarBlob L((int)O.tets.size(), (int)O.grp.size() / 2, (int)O.src.size(), (int)O.chunk.size() / 2, (int)O.code.size());
        m_nt = std::max(m_nt, (int)O.tets.size());
        m_a = std::max(m_a, L.bytesA);
        m_b = std::max(m_b, L.bytesB);
        m_nv = std::max(m_nv, (int)O.gv.size() + 1);
        worst_rows = std::max(worst_rows, O.rows);
      }
      worst = StarLayout(m_nt, worst_rows, m_a, m_b, m_nv).total + WAE_STAR_STATIC_SMEM;
    }
    if (bad) WAE_THROW(WAE_E_INVALID, "star program: inconsistent program (flags %d)", bad);
    if (worst <= smem_budget && worst_rows * WAE_STAR_RS <= 8192) break;
    if (attempt >= 8 || cap_nt <= cap_min) WAE_THROW(WAE_E_INVALID, "star program does not fit %lld bytes of shared memory", (long long)smem_budget);
    const double f = std::min(0.9, std::min((double)smem_budget / (double)worst, (8192.0 / WAE_STAR_RS) / (double)std::max(worst_rows, 1)) * 0.97);
    cap_nt = std::max<int64_t>(cap_min, (int64_t)(cap_nt * f));
  }
  tick("star: per-patch programs");
  if (timing) {
    double mn = 1e300, mx = 0, sum = 0, dmn = 1e300, dmx = 0;
    for (int64_t p = 0; p < npatch; p++) {
      const double b2 = (double)patch_smem(nloc, po[p]), nd = (double)(cut[p + 1] - cut[p]);
      mn = std::min(mn, b2); mx = std::max(mx, b2); sum += b2;
      dmn = std::min(dmn, b2 / nd); dmx = std::max(dmx, b2 / nd);
    }
    fprintf(stderr, "[wae symbolic] star: %lld patches, shared memory per patch min %.0f mean %.0f max %.0f (budget %lld), per owned DOF %.0f .. %.0f, cap %lld\n",
            (long long)npatch, mn, sum / (double)npatch, mx, (long long)smem_budget, dmn, dmx, (long long)cap_nt);
  }
  // ---- pack: descriptor + blob A + blob B + patch vertices per patch (assembly_star.h) -----------------------------------------------
  G = StarHost();
  G.nloc = nloc;
  G.n_patch = (int)npatch;
  G.desc.assign((size_t)npatch * 8, 0);
  int64_t a_total = 0, b_total = 0, pv_total = 0, chunk_total = 0;
  for (int64_t p = 0; p < npatch; p++) {
    PatchOut& O = po[p];
    if (O.gv.size() & 1) O.gv.push_back(O.gv.back());
    const int nt = (int)O.tets.size(), ng = (int)O.grp.size() / 2, nc = (int)O.chunk.size() / 2, nv = (int)O.gv.size();
    const StarBlob L(nt, ng, (int)O.src.size(), nc, (int)O.code.size());
    StarDesc& D = *reinterpret_cast<StarDesc*>(&G.desc[(size_t)p * 8]);
    D.blobA = a_total;
    D.pxyz = pv_total * 3;
    D.blobB = b_total;
    D.nt = nt; D.nv = nv; D.ng = ng; D.nc = nc;
    D.nsrc = (int)O.src.size(); D.ncode = (int)O.code.size(); D.bytesA = L.bytesA; D.bytesB = L.bytesB;
    a_total += L.bytesA;
    b_total += L.bytesB;
    pv_total += nv;
    chunk_total += nc;
    G.max_nt = std::max(G.max_nt, nt);
    G.max_nv = std::max(G.max_nv, nv);
    G.max_rows = std::max(G.max_rows, O.rows);
    G.max_ng = std::max(G.max_ng, ng);
    G.max_a = std::max(G.max_a, L.bytesA);
    G.max_b = std::max(G.max_b, L.bytesB);
    G.max_smem = std::max(G.max_smem, patch_smem(nloc, O));
    G.n_staged += nt;
    G.n_entities += O.entities;
    G.n_sources += O.sources;
  }
  G.n_chunks = chunk_total;
  G.blob.resize((size_t)a_total);
  G.blobB.resize((size_t)b_total);
  G.gvtx.resize(pv_total);
  parallel_for(npatch, [&](int64_t pa, int64_t pb) {
    for (int64_t p = pa; p < pb; p++) {
      PatchOut& O = po[p];
      const StarDesc& D = *reinterpret_cast<const StarDesc*>(&G.desc[(size_t)p * 8]);
      const StarBlob L(D.nt, D.ng, D.nsrc, D.nc, D.ncode);
      uint8_t* A = G.blob.data() + D.blobA;
      std::memset(A, 0, (size_t)L.bytesA);
      std::memcpy(A, O.lvtx.data(), O.lvtx.size() * 2);
      std::memcpy(A + L.o_tets, O.tets.data(), O.tets.size() * 4);
      std::memcpy(A + L.o_grp, O.grp.data(), O.grp.size() * 4);
      std::memcpy(A + L.o_cnt, O.cnt.data(), O.cnt.size());
      std::memcpy(A + L.o_src, O.src.data(), O.src.size() * 2);
      uint8_t* B = G.blobB.data() + D.blobB;
      std::memset(B, 0, (size_t)L.bytesB);
      std::memcpy(B, O.chunk.data(), O.chunk.size() * 4);
      std::memcpy(B + L.o_code, O.code.data(), O.code.size() * 2);
      std::copy(O.gv.begin(), O.gv.end(), G.gvtx.begin() + D.pxyz / 3);
      O = PatchOut();
    }
  }, 16);
  tick("star: pack");
}
