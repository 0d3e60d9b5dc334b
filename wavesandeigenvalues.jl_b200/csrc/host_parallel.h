// Host-side helpers shared by the symbolic phases (pattern / pair program / mesh numbering): thread-parallel loops and sorts.
#pragma once
#include <algorithm>
#include <cstdint>
#include <thread>
#include <vector>

namespace {
// min_parallel: below this many items the loop runs on the calling thread (cheap bodies); loops over patches pass a small value
template <typename F>
void parallel_for(int64_t n, F f, int64_t min_parallel = 4096) {
  unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  if (n < min_parallel) nt = 1;
  std::vector<std::thread> th;
  int64_t chunk = (n + nt - 1) / nt;
  for (unsigned t = 0; t < nt; t++) {
    int64_t a = t * chunk, b = std::min<int64_t>(n, a + chunk);
    if (a >= b) break;
    th.emplace_back([=]() { f(a, b); });
  }
  for (auto& x : th) x.join();
}
// sort with the host threads: sorted chunks, then pairwise merges (stable enough for unique keys)
template <typename T>
void parallel_sort(std::vector<T>& v) {
  const size_t n = v.size();
  unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  if (n < (size_t)1 << 16 || nt == 1) {
    std::sort(v.begin(), v.end());
    return;
  }
  unsigned parts = 1;
  while (parts * 2 <= nt) parts *= 2;
  std::vector<size_t> b(parts + 1);
  for (unsigned q = 0; q <= parts; q++) b[q] = n * q / parts;
  {
    std::vector<std::thread> th;
    for (unsigned q = 0; q < parts; q++) th.emplace_back([&, q]() { std::sort(v.begin() + b[q], v.begin() + b[q + 1]); });
    for (auto& x : th) x.join();
  }
  for (unsigned w = 1; w < parts; w *= 2) {
    std::vector<std::thread> th;
    for (unsigned q = 0; q + w < parts; q += 2 * w)
      th.emplace_back([&, q, w]() { std::inplace_merge(v.begin() + b[q], v.begin() + b[q + w], v.begin() + b[std::min(parts, q + 2 * w)]); });
    for (auto& x : th) x.join();
  }
}
}  // namespace
