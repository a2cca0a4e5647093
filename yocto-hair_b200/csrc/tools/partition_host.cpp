// Host check of csrc/yh_partition.h: partition_replay() must leave an array in exactly the state libstdc++'s
// std::partition leaves it in (tests/test_partition_host.py).  Returns the number of differing positions
// (+ 1 if the partition points differ).
#include <cstdint>
#include <vector>

#include "yh_partition.h"

struct item { float key[3]; int id; };

extern "C" long yh_host_partition_check(long n, unsigned long long seed, int threads, int axis, float middle_fraction, int duplicates) {
  std::vector<item> a((size_t)n);
  unsigned long long s = seed * 6364136223846793005ULL + 1442695040888963407ULL;
  auto next = [&]() { s = s * 6364136223846793005ULL + 1442695040888963407ULL; return (unsigned)(s >> 33); };
  for (long i = 0; i < n; i++) {
    for (int c = 0; c < 3; c++) {
      unsigned r = next();
      a[(size_t)i].key[c] = duplicates ? (float)(r % (unsigned)duplicates) : (float)r / 2147483648.0f;
    }
    a[(size_t)i].id = (int)i;
  }
  float middle = duplicates ? middle_fraction * (float)duplicates : middle_fraction;
  auto  pred = [axis, middle](const item& p) { return p.key[axis] < middle; };
  std::vector<item> b = a;
  size_t m0 = (size_t)(std::partition(a.begin(), a.end(), pred) - a.begin());
  size_t m1 = yh::partition_replay(b.data(), b.size(), pred, threads);
  long diff = m0 != m1 ? 1 : 0;
  for (size_t i = 0; i < a.size(); i++) diff += a[i].id != b[i].id ? 1 : 0;
  return diff;
}
