// PCG32 (XSH-RR), integer-exact restatement of yocto_math.h:1396-1442.
#pragma once
#include "yh_math.cuh"

namespace yh {

struct rng_state {
  uint64_t state, inc;
};

YH_HD uint32_t advance_rng(rng_state& rng) {  // :1410-1416
  uint64_t oldstate   = rng.state;
  rng.state           = oldstate * 6364136223846793005ULL + rng.inc;
  uint32_t xorshifted = (uint32_t)(((oldstate >> 18u) ^ oldstate) >> 27u);
  uint32_t rot        = (uint32_t)(oldstate >> 59u);
  return (xorshifted >> rot) | (xorshifted << ((0u - rot) & 31));
}
YH_HD rng_state make_rng(uint64_t seed, uint64_t seq = 1) {  // :1419-1427
  rng_state rng;
  rng.state = 0U;
  rng.inc   = (seq << 1u) | 1u;
  advance_rng(rng);
  rng.state += seed;
  advance_rng(rng);
  return rng;
}
YH_HD int   rand1i(rng_state& rng, int n) { return (int)(advance_rng(rng) % (uint32_t)n); }  // :1430
YH_HD float rand1f(rng_state& rng) {                                                         // :1431-1441
  return u2f((advance_rng(rng) >> 9) | 0x3f800000u) - 1.0f;
}
YH_HD vec2f rand2f(rng_state& rng) {  // :1442-1447, x first
  float x = rand1f(rng);
  float y = rand1f(rng);
  return {x, y};
}

}  // namespace yh
