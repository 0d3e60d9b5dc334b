// Host symbolic phase of the sparse LU: nested-dissection ordering, supernodes (= separators
// and leaf sub-domains), row structures, assembly tree, frontal layout and scatter maps.
// The sparsity pattern of the operator family does not depend on omega, so this runs once per
// family (the reference re-runs UMFPACK's symbolic+numeric phases on every lu() call).
//
// Ordering: recursive bisection along the longest coordinate axis of the current node set (DOF
// coordinates come from the mesh; without coordinates a BFS level structure from a pseudo-
// peripheral node provides the 1-D key), then the minimum vertex separator contained in the
// edge cut, obtained as a minimum vertex cover of the bipartite cut graph (Hopcroft-Karp matching
// + Koenig's theorem).  Separator nodes are ordered last and form one supernode.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <queue>

#include "lu.h"

namespace {

struct Graph {
  int64_t n;
  std::vector<int64_t> xadj;
  std::vector<int32_t> adj;
};

// adjacency of A + A^T without the diagonal
void build_graph(int64_t n, const int64_t* colptr, const int32_t* rowval, Graph& g) {
  g.n = n;
  std::vector<int64_t> deg(n + 1, 0);
  for (int64_t j = 0; j < n; j++)
    for (int64_t k = colptr[j]; k < colptr[j + 1]; k++) {
      int32_t i = rowval[k];
      if (i != j) {
        deg[i + 1]++;
        deg[j + 1]++;
      }
    }
  for (int64_t i = 0; i < n; i++) deg[i + 1] += deg[i];
  std::vector<int32_t> tmp(deg[n]);
  std::vector<int64_t> pos(deg.begin(), deg.end() - 1);
  for (int64_t j = 0; j < n; j++)
    for (int64_t k = colptr[j]; k < colptr[j + 1]; k++) {
      int32_t i = rowval[k];
      if (i != j) {
        tmp[pos[i]++] = (int32_t)j;
        tmp[pos[j]++] = i;
      }
    }
  g.xadj.assign(n + 1, 0);
  g.adj.clear();
  g.adj.reserve(deg[n] / 2 + n);
  for (int64_t i = 0; i < n; i++) {
    int32_t* b = tmp.data() + deg[i];
    int32_t* e = tmp.data() + deg[i + 1];
    std::sort(b, e);
    e = std::unique(b, e);
    g.adj.insert(g.adj.end(), b, e);
    g.xadj[i + 1] = (int64_t)g.adj.size();
  }
}

struct NDNode {
  std::vector<int32_t> verts;
  std::vector<int> children;
};

struct ND {
  const Graph& g;
  const double* coords;
  int leaf;
  std::vector<int32_t> mark;   // membership stamp
  std::vector<int8_t> side;
  std::vector<int32_t> lid;    // local ids for the matching
  std::vector<NDNode> tree;
  int stamp = 0;
  ND(const Graph& g_, const double* c, int leaf_) : g(g_), coords(c), leaf(leaf_), mark(g_.n, -1), side(g_.n, 0), lid(g_.n, -1) {}

  // 1-D key for the bisection when no coordinates are available: BFS order from a pseudo-peripheral node
  void bfs_keys(const std::vector<int32_t>& S, int st, std::vector<double>& key) {
    std::vector<int32_t> order;
    order.reserve(S.size());
    std::vector<int32_t>& seen = lid;  // reuse as visit flag (reset afterwards)
    auto bfs = [&](int32_t root, std::vector<int32_t>& out) {
      size_t head = out.size();
      out.push_back(root);
      seen[root] = 1;
      while (head < out.size()) {
        int32_t v = out[head++];
        for (int64_t q = g.xadj[v]; q < g.xadj[v + 1]; q++) {
          int32_t u = g.adj[q];
          if (mark[u] == st && seen[u] < 0) {
            seen[u] = 1;
            out.push_back(u);
          }
        }
      }
    };
    // pseudo-peripheral start: two sweeps inside the component of S[0]
    std::vector<int32_t> tmp;
    bfs(S[0], tmp);
    int32_t far = tmp.back();
    for (int32_t v : tmp) seen[v] = -1;
    tmp.clear();
    bfs(far, tmp);
    far = tmp.back();
    for (int32_t v : tmp) seen[v] = -1;
    bfs(far, order);
    for (int32_t v : S)
      if (seen[v] < 0) bfs(v, order);  // remaining components
    key.resize(S.size());
    for (size_t i = 0; i < order.size(); i++) seen[order[i]] = (int32_t)i;
    for (size_t i = 0; i < S.size(); i++) key[i] = (double)seen[S[i]];
    for (int32_t v : S) seen[v] = -1;
  }

  // returns the roots of the elimination sub-forest of S
  std::vector<int> dissect(std::vector<int32_t>& S) {
    std::vector<int> roots;
    if (S.empty()) return roots;
    if ((int)S.size() <= leaf) {
      tree.emplace_back();
      tree.back().verts.swap(S);
      roots.push_back((int)tree.size() - 1);
      return roots;
    }
    const int st = ++stamp;
    for (int32_t v : S) mark[v] = st;
    // split key
    std::vector<double> key(S.size());
    if (coords) {
      double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
      for (int32_t v : S)
        for (int r = 0; r < 3; r++) {
          lo[r] = std::min(lo[r], coords[3 * (size_t)v + r]);
          hi[r] = std::max(hi[r], coords[3 * (size_t)v + r]);
        }
      int ax = 0;
      for (int r = 1; r < 3; r++)
        if (hi[r] - lo[r] > hi[ax] - lo[ax]) ax = r;
      for (size_t i = 0; i < S.size(); i++) key[i] = coords[3 * (size_t)S[i] + ax];
    } else {
      bfs_keys(S, st, key);
    }
    std::vector<int32_t> idx(S.size());
    for (size_t i = 0; i < S.size(); i++) idx[i] = (int32_t)i;
    size_t half = S.size() / 2;
    auto less = [&](int32_t a, int32_t b) { return key[a] < key[b] || (key[a] == key[b] && S[a] < S[b]); };
    if (coords && getenv("WAE_LU_NO_SNAP") == nullptr) {
      // The DOFs of a (nearly) structured mesh sit on planes: a split at the exact median cuts through the plane that holds the
      // median, the cut zigzags and its vertex cover is up to 1.5x a plane.  Snap the split to the widest gap of the key among
      // the middle 20 % of the nodes (between two planes): the halves stay balanced within 40/60 and the separator is one plane.
      const size_t w0 = S.size() * 2 / 5, w1 = S.size() - w0;
      std::nth_element(idx.begin(), idx.begin() + w0, idx.end(), less);
      std::nth_element(idx.begin() + w0, idx.begin() + w1, idx.end(), less);
      std::sort(idx.begin() + w0, idx.begin() + w1, less);
      double best = -1.0;
      for (size_t i = w0; i + 1 < w1; i++) {
        const double gap = key[idx[i + 1]] - key[idx[i]];
        // prefer the gap closest to the median among (almost) equally wide ones
        const double score = gap * (1.0 - 0.25 * std::fabs((double)(i + 1) - 0.5 * (double)S.size()) / (double)(w1 - w0));
        if (score > best) {
          best = score;
          half = i + 1;
        }
      }
    } else {
      std::nth_element(idx.begin(), idx.begin() + half, idx.end(), less);
    }
    for (size_t i = 0; i < S.size(); i++) side[S[idx[i]]] = i < half ? 0 : 1;
    // bipartite cut graph
    std::vector<int32_t> L, R;  // boundary nodes of side 0 / side 1
    std::vector<int64_t> lptr(1, 0);
    std::vector<int32_t> ladj;
    for (int32_t v : S) {
      if (side[v] != 0) continue;
      bool any = false;
      for (int64_t q = g.xadj[v]; q < g.xadj[v + 1]; q++) {
        int32_t u = g.adj[q];
        if (mark[u] == st && side[u] == 1) {
          if (lid[u] < 0) {
            lid[u] = (int32_t)R.size();
            R.push_back(u);
          }
          ladj.push_back(lid[u]);
          any = true;
        }
      }
      if (any) {
        L.push_back(v);
        lptr.push_back((int64_t)ladj.size());
      }
    }
    for (int32_t u : R) lid[u] = -1;
    const int nl = (int)L.size(), nr = (int)R.size();
    // Hopcroft-Karp
    std::vector<int32_t> ml(nl, -1), mr(nr, -1), dist(nl);
    auto bfs = [&]() {
      std::queue<int32_t> q;
      bool found = false;
      for (int i = 0; i < nl; i++) {
        if (ml[i] < 0) {
          dist[i] = 0;
          q.push(i);
        } else
          dist[i] = -1;
      }
      while (!q.empty()) {
        int32_t a = q.front();
        q.pop();
        for (int64_t e = lptr[a]; e < lptr[a + 1]; e++) {
          int32_t b = mr[ladj[e]];
          if (b < 0)
            found = true;
          else if (dist[b] < 0) {
            dist[b] = dist[a] + 1;
            q.push(b);
          }
        }
      }
      return found;
    };
    std::vector<int64_t> it(nl);
    std::function<bool(int32_t)> dfs = [&](int32_t a) -> bool {
      for (int64_t& e = it[a]; e < lptr[a + 1]; e++) {
        int32_t rb = ladj[e];
        int32_t b = mr[rb];
        if (b < 0 || (dist[b] == dist[a] + 1 && dfs(b))) {
          ml[a] = rb;
          mr[rb] = a;
          return true;
        }
      }
      dist[a] = -1;
      return false;
    };
    while (bfs()) {
      for (int i = 0; i < nl; i++) it[i] = lptr[i];
      for (int i = 0; i < nl; i++)
        if (ml[i] < 0) dfs(i);
    }
    // Koenig: Z = reachable from unmatched left nodes by alternating paths; cover = (L \ Z) u (R n Z)
    std::vector<int8_t> zl(nl, 0), zr(nr, 0);
    {
      std::vector<int32_t> stack;
      for (int i = 0; i < nl; i++)
        if (ml[i] < 0) {
          zl[i] = 1;
          stack.push_back(i);
        }
      while (!stack.empty()) {
        int32_t a = stack.back();
        stack.pop_back();
        for (int64_t e = lptr[a]; e < lptr[a + 1]; e++) {
          int32_t rb = ladj[e];
          if (ml[a] == rb || zr[rb]) continue;
          zr[rb] = 1;
          int32_t b = mr[rb];
          if (b >= 0 && !zl[b]) {
            zl[b] = 1;
            stack.push_back(b);
          }
        }
      }
    }
    std::vector<int32_t> sep;
    for (int i = 0; i < nl; i++)
      if (!zl[i]) {
        sep.push_back(L[i]);
        side[L[i]] = 2;
      }
    for (int i = 0; i < nr; i++)
      if (zr[i]) {
        sep.push_back(R[i]);
        side[R[i]] = 2;
      }
    std::vector<int32_t> A, B;
    A.reserve(half);
    B.reserve(S.size() - half);
    for (int32_t v : S) {
      if (side[v] == 0) A.push_back(v);
      else if (side[v] == 1) B.push_back(v);
    }
    std::vector<int32_t>().swap(S);
    std::vector<int> ra = dissect(A);
    std::vector<int> rb = dissect(B);
    ra.insert(ra.end(), rb.begin(), rb.end());
    if (sep.empty()) return ra;  // the halves were not connected
    tree.emplace_back();
    tree.back().verts.swap(sep);
    tree.back().children = ra;
    roots.push_back((int)tree.size() - 1);
    return roots;
  }
};

}  // namespace

void wae_lu_symbolic(int64_t n, const int64_t* colptr, const int32_t* rowval, const double* coords, int leaf_size, LuSymbolic& S) {
  Graph g;
  build_graph(n, colptr, rowval, g);
  ND nd(g, coords, leaf_size);
  std::vector<int32_t> all(n);
  for (int64_t i = 0; i < n; i++) all[i] = (int32_t)i;
  std::vector<int> roots = nd.dissect(all);
  // postorder numbering of the ND forest
  S = LuSymbolic();
  S.n = n;
  S.perm.reserve(n);
  S.sn_first.push_back(0);
  {
    std::vector<std::pair<int, size_t>> stack;
    for (auto it = roots.rbegin(); it != roots.rend(); ++it) stack.emplace_back(*it, 0);
    while (!stack.empty()) {
      auto& top = stack.back();
      NDNode& nd_ = nd.tree[top.first];
      if (top.second < nd_.children.size()) {
        int c = nd_.children[top.second++];
        stack.emplace_back(c, 0);
      } else {
        S.perm.insert(S.perm.end(), nd_.verts.begin(), nd_.verts.end());
        S.sn_first.push_back((int32_t)S.perm.size());
        std::vector<int32_t>().swap(nd_.verts);
        stack.pop_back();
      }
    }
  }
  if ((int64_t)S.perm.size() != n) WAE_THROW(WAE_E_INVALID, "internal: ordering covers %zu of %lld nodes", S.perm.size(), (long long)n);
  S.nsn = (int)S.sn_first.size() - 1;
  S.iperm.resize(n);
  for (int64_t p = 0; p < n; p++) S.iperm[S.perm[p]] = (int32_t)p;
  std::vector<int32_t> sn_of(n);
  for (int k = 0; k < S.nsn; k++)
    for (int32_t p = S.sn_first[k]; p < S.sn_first[k + 1]; p++) sn_of[p] = k;
  // row structures, bottom-up; true assembly-tree parent = supernode of the first structure row
  S.sn_parent.assign(S.nsn, -1);
  S.struct_ptr.assign(S.nsn + 1, 0);
  std::vector<std::vector<int32_t>> pending(S.nsn);
  std::vector<int32_t> seen(n, -1), cur;
  for (int k = 0; k < S.nsn; k++) {
    const int32_t f = S.sn_first[k], l = S.sn_first[k + 1];
    cur.clear();
    for (int32_t p = f; p < l; p++) {
      int32_t v = S.perm[p];
      for (int64_t q = g.xadj[v]; q < g.xadj[v + 1]; q++) {
        int32_t pq = S.iperm[g.adj[q]];
        if (pq >= l && seen[pq] != k) {
          seen[pq] = k;
          cur.push_back(pq);
        }
      }
    }
    for (int32_t pq : pending[k])
      if (pq >= l && seen[pq] != k) {
        seen[pq] = k;
        cur.push_back(pq);
      }
    std::vector<int32_t>().swap(pending[k]);
    std::sort(cur.begin(), cur.end());
    S.struct_idx.insert(S.struct_idx.end(), cur.begin(), cur.end());
    S.struct_ptr[k + 1] = (int64_t)S.struct_idx.size();
    if (!cur.empty()) {
      int par = sn_of[cur[0]];
      S.sn_parent[k] = par;
      pending[par].insert(pending[par].end(), cur.begin(), cur.end());
    }
  }
  // depth, levels
  S.sn_depth.assign(S.nsn, 0);
  int maxd = 0;
  for (int k = S.nsn - 1; k >= 0; k--) {
    if (S.sn_parent[k] >= 0) S.sn_depth[k] = S.sn_depth[S.sn_parent[k]] + 1;
    maxd = std::max(maxd, (int)S.sn_depth[k]);
  }
  S.levels.assign(maxd + 1, {});
  for (int k = 0; k < S.nsn; k++) S.levels[S.sn_depth[k]].push_back(k);
  // layout + statistics
  S.lp_off.resize(S.nsn);
  S.up_off.resize(S.nsn);
  S.upd_off.resize(S.nsn);
  S.dinv_off.resize(S.nsn);
  int64_t doff = 0;
  S.level_upd_size.assign(maxd + 1, 0);
  int64_t off = 0;
  for (int k = 0; k < S.nsn; k++) {
    int64_t s = S.sn_first[k + 1] - S.sn_first[k], r = S.struct_ptr[k + 1] - S.struct_ptr[k];
    S.lp_off[k] = off;
    off += (s + r) * s;
    S.up_off[k] = off;
    off += (s + r) * s;
    S.dinv_off[k] = doff;
    doff += 2 * (int64_t)WAE_LU_NB * WAE_LU_NB * ((s + WAE_LU_NB - 1) / WAE_LU_NB);
    S.upd_off[k] = S.level_upd_size[S.sn_depth[k]];
    S.level_upd_size[S.sn_depth[k]] += r * r;
    S.factor_nnz += s * s + 2 * s * r;
    S.max_s = std::max<int>(S.max_s, (int)s);
    S.max_r = std::max<int>(S.max_r, (int)r);
    double m = (double)(s + r);
    // sum_{j=0}^{s-1} (m-1-j)^2 + (m-1-j) complex multiply-adds
    double ds = (double)s;
    double sum1 = ds * (m - 1) - ds * (ds - 1) / 2;
    double sum2 = ds * (m - 1) * (m - 1) - (m - 1) * ds * (ds - 1) + (ds - 1) * ds * (2 * ds - 1) / 6;
    S.flops += 8.0 * (sum1 + sum2);
  }
  S.fac_size = off;
  S.dinv_size = doff;
  // relative indices of every structure row in the parent's front
  S.rel_ptr = S.struct_ptr;
  S.rel_idx.resize(S.struct_idx.size());
  for (int k = 0; k < S.nsn; k++) {
    int par = S.sn_parent[k];
    if (par < 0) continue;
    const int32_t pf = S.sn_first[par], pl = S.sn_first[par + 1];
    const int32_t* pb = S.struct_idx.data() + S.struct_ptr[par];
    const int32_t* pe = S.struct_idx.data() + S.struct_ptr[par + 1];
    for (int64_t q = S.struct_ptr[k]; q < S.struct_ptr[k + 1]; q++) {
      int32_t p = S.struct_idx[q];
      if (p < pl)
        S.rel_idx[q] = p - pf;
      else {
        const int32_t* it = std::lower_bound(pb, pe, p);
        if (it == pe || *it != p) WAE_THROW(WAE_E_INVALID, "internal: structure of supernode %d not contained in its parent", k);
        S.rel_idx[q] = (pl - pf) + (int32_t)(it - pb);
      }
    }
  }
  // scatter map of A into the fronts
  int64_t nnz = colptr[n];
  S.amap.resize(nnz);
  S.diagpos.assign(n, -1);
  for (int64_t j = 0; j < n; j++) {
    for (int64_t k = colptr[j]; k < colptr[j + 1]; k++) {
      int32_t i = rowval[k];
      if (i == j) S.diagpos[j] = (int32_t)k;
      int32_t pi = S.iperm[i], pj = S.iperm[j];
      int own = sn_of[std::min(pi, pj)];
      int32_t f = S.sn_first[own];
      int64_t s = S.sn_first[own + 1] - f, r = S.struct_ptr[own + 1] - S.struct_ptr[own], ld = s + r;
      const int32_t* sb = S.struct_idx.data() + S.struct_ptr[own];
      const int32_t* se = sb + r;
      auto sidx = [&](int32_t p) -> int64_t {
        const int32_t* it = std::lower_bound(sb, se, p);
        if (it == se || *it != p) WAE_THROW(WAE_E_INVALID, "internal: entry (%d,%d) outside the symbolic structure", (int)i, (int)j);
        return s + (it - sb);
      };
      if (pj < f + s) {
        int64_t b = pj - f;
        int64_t a = pi < f + s ? pi - f : sidx(pi);
        if (a < s && a / WAE_LU_NB < b / WAE_LU_NB)
          S.amap[k] = S.up_off[own] + b + a * ld;  // upper block of the pivot block, stored transposed
        else
          S.amap[k] = S.lp_off[own] + a + b * ld;
      } else {
        int64_t a = pi - f, c = sidx(pj);
        S.amap[k] = S.up_off[own] + c + a * ld;
      }
    }
  }
}
