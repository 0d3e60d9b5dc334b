// Host-side mesh numbering (no GPU, no context): the reference's simplex ordering rule, sort-based and thread-parallel.
//
// The reference keeps lines / triangles / tetrahedra as lists sorted ascending in the DESCENDING-sorted vertex tuple
// (src/Mesh/sorter.jl:9-31), inserts simplex by simplex (insert_smplx!, :141-150: O(n^2)) and drops repeated vertex sets, the first
// occurrence keeping its vertex order (Meshutils.jl:92-165; collect_lines! :831-840 does the same for the six edges of every
// tetrahedron).  Here: one key per simplex (its descending vertex tuple packed into 128 bits), one parallel sort of (key, input row),
// group starts = the unique simplices, smallest input row of a group = its representative.  Identical numbering, O(n log n).
#include <atomic>
#include <cstdint>
#include <cstring>

#include "../../include/wae_b200.h"
#include "host_parallel.h"

namespace {
struct Rec2 {  // edges: two 32-bit vertex ids in one word
  uint64_t key;
  int64_t row;
  bool operator<(const Rec2& o) const { return key < o.key || (key == o.key && row < o.row); }
};
struct Rec4 {  // triangles / tetrahedra
  uint64_t hi, lo;
  int64_t row;
  bool operator<(const Rec4& o) const { return hi < o.hi || (hi == o.hi && (lo < o.lo || (lo == o.lo && row < o.row))); }
};
inline bool same_key(const Rec2& a, const Rec2& b) { return a.key == b.key; }
inline bool same_key(const Rec4& a, const Rec4& b) { return a.hi == b.hi && a.lo == b.lo; }

template <class R>
int64_t finish(std::vector<R>& rec, int64_t* first, int64_t* inv) {
  const int64_t n = (int64_t)rec.size();
  parallel_sort(rec);
  // unique index of every record = number of group starts before it: per-chunk counts, exclusive scan, second pass
  // (same chunking as parallel_for: chunk index = range start / chunk)
  const unsigned nt = n < 4096 ? 1u : std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  const int64_t chunk = std::max<int64_t>(1, (n + nt - 1) / nt);
  std::vector<int64_t> starts(nt + 1, 0);
  parallel_for(n, [&](int64_t a, int64_t b) {
    int64_t c = 0;
    for (int64_t i = a; i < b; i++) c += (i == 0 || !same_key(rec[i - 1], rec[i]));
    starts[a / chunk + 1] = c;
  });
  for (unsigned t = 0; t < nt; t++) starts[t + 1] += starts[t];
  parallel_for(n, [&](int64_t a, int64_t b) {
    int64_t u = starts[a / chunk] - 1;  // a chunk that starts inside a group continues the previous chunk's last unique index
    for (int64_t i = a; i < b; i++) {
      if (i == 0 || !same_key(rec[i - 1], rec[i])) first[++u] = rec[i].row;  // (key, row) order: the group start has the smallest row
      inv[rec[i].row] = u;
    }
  });
  return starts[nt];
}
}  // namespace

extern "C" int32_t wae_sorted_unique_simplices(int64_t n, int32_t k, const int64_t* simp, int64_t* first, int64_t* inv, int64_t* n_unique) {
  if (n < 0 || k < 2 || k > 4 || (n && (!simp || !first || !inv)) || !n_unique) return WAE_E_INVALID;
  try {
    std::atomic<bool> bad{false};
    if (k == 2) {
      std::vector<Rec2> rec((size_t)n);
      parallel_for(n, [&](int64_t a, int64_t b) {
        for (int64_t i = a; i < b; i++) {
          const int64_t x = simp[2 * i], y = simp[2 * i + 1];
          if ((x | y) < 0 || (x | y) > 0xFFFFFFFFll) bad = true;
          const uint64_t h = (uint64_t)std::max(x, y), l = (uint64_t)std::min(x, y);
          rec[i] = Rec2{(h << 32) | l, i};
        }
      });
      if (bad) return WAE_E_INVALID;
      *n_unique = finish(rec, first, inv);
    } else {
      std::vector<Rec4> rec((size_t)n);
      parallel_for(n, [&](int64_t a, int64_t b) {
        for (int64_t i = a; i < b; i++) {
          int64_t v[4] = {0, 0, 0, 0};
          for (int q = 0; q < k; q++) {
            v[q] = simp[(size_t)k * i + q];
            if (v[q] < 0 || v[q] > 0xFFFFFFFFll) bad = true;
          }
          // descending sort of the k ids (insertion sort); a missing 4th id of a triangle is the lowest word and stays 0 for all
          for (int p = 1; p < k; p++)
            for (int q = p; q > 0 && v[q] > v[q - 1]; q--) std::swap(v[q], v[q - 1]);
          rec[i] = Rec4{((uint64_t)v[0] << 32) | (uint64_t)v[1], ((uint64_t)v[2] << 32) | (uint64_t)v[3], i};
        }
      });
      if (bad) return WAE_E_INVALID;
      *n_unique = finish(rec, first, inv);
    }
    return WAE_OK;
  } catch (const std::bad_alloc&) {
    return WAE_E_NOMEM;
  } catch (...) {
    return WAE_E_INVALID;
  }
}
