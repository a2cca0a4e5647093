// std::partition, replayed in parallel with the SAME resulting permutation (host only).
//
// The reference builds its BVH with std::partition (split_middle, yocto_pathtrace.cpp:564-595), and the ORDER
// it leaves the primitives in decides which of two primitives hit at exactly the same distance its traversal
// reports (yh_scene.cuh "tie order").  That order has to be reproduced, not just the partition point, and on a
// million primitives the top splits are a serial bottleneck of the scene set-up.
//
// libstdc++'s std::partition for bidirectional iterators is Hoare's scheme: the k-th element from the LEFT that
// fails the predicate is swapped with the k-th element from the RIGHT that satisfies it, for as long as the
// former lies left of the latter.  With m = number of elements satisfying the predicate, that is exactly: the
// misplaced "false" elements of [0, m), in increasing index order, are exchanged with the misplaced "true"
// elements of [m, n), in decreasing index order; everything else stays where it is.  Both lists have the same
// length K, so the permutation can be formed by counting (parallel), ranking (prefix sums) and K independent
// swaps.  csrc/tools/partition_host.cpp + tests/test_partition_host.py compare it with std::partition itself.
#pragma once
#include <algorithm>
#include <cstddef>
#include <future>
#include <utility>
#include <vector>

namespace yh {

template <class T, class Pred>
size_t partition_replay(T* a, size_t n, Pred pred, int threads) {
  if (threads <= 1 || n < (size_t)threads * 1024) return (size_t)(std::partition(a, a + n, pred) - a);
  const int nt = threads;
  auto lo = [&](int t) { return n * (size_t)t / nt; };
  auto run = [&](auto&& body) {
    std::vector<std::future<void>> jobs;
    for (int t = 0; t < nt; t++) jobs.push_back(std::async(std::launch::async, [&body, t]() { body(t); }));
    for (auto& j : jobs) j.get();
  };
  // 1. m = number of "true" elements
  std::vector<size_t> trues(nt, 0);
  run([&](int t) {
    size_t c = 0;
    for (size_t i = lo(t); i < lo(t + 1); i++) c += pred(a[i]) ? 1 : 0;
    trues[t] = c;
  });
  size_t m = 0;
  for (size_t c : trues) m += c;
  // 2. misplaced elements per chunk: "false" in [0, m), "true" in [m, n)
  std::vector<size_t> nf(nt, 0), ntr(nt, 0);
  run([&](int t) {
    size_t f = 0, r = 0;
    for (size_t i = lo(t); i < lo(t + 1); i++) {
      bool p = pred(a[i]);
      if (i < m) f += p ? 0 : 1;
      else r += p ? 1 : 0;
    }
    nf[t] = f, ntr[t] = r;
  });
  size_t K = 0;
  std::vector<size_t> f0(nt, 0), r0(nt, 0);  // rank of a chunk's first misplaced false (from the left) / true (from the right)
  for (int t = 0; t < nt; t++) f0[t] = K, K += nf[t];
  size_t R = 0;
  for (int t = nt - 1; t >= 0; t--) r0[t] = R, R += ntr[t];
  if (K == 0) return m;
  // 3. index lists by rank
  std::vector<size_t> fi(K), ti(K);
  run([&](int t) {
    size_t f = f0[t];
    for (size_t i = lo(t); i < lo(t + 1) && i < m; i++)
      if (!pred(a[i])) fi[f++] = i;
    size_t r = r0[t] + ntr[t];  // ranks grow towards the left inside the chunk
    for (size_t i = std::max(lo(t), m); i < lo(t + 1); i++)
      if (pred(a[i])) ti[--r] = i;
  });
  // 4. the K swaps
  run([&](int t) {
    for (size_t k = K * (size_t)t / nt; k < K * (size_t)(t + 1) / nt; k++) std::swap(a[fi[k]], a[ti[k]]);
  });
  return m;
}

}  // namespace yh
