// climt_b200 -- host-side Mersenne-twister mask generation for McICA bit parity with the reference's default RNG.
// MersenneTwister module: climt/_lib/rrtmg_lw/mcica_random_numbers.f90:60-300 (initialize_scalar :169-185, nextState
// :116-133, temper :154-165, getRandomReal :276-295 with its default-real numerator for negative integers).
#pragma once
#include <cstdint>
#include <vector>

#include "mcica_core.cuh"

namespace cb {
namespace mcica {

struct MT19937 {
  uint32_t state[624];
  int cur;
  explicit MT19937(int32_t seed) {
    state[0] = (uint32_t)seed;
    for (int i = 1; i < 624; ++i) state[i] = 1812433253u * (state[i - 1] ^ (state[i - 1] >> 30)) + (uint32_t)i;
    cur = 624;
  }
  static uint32_t twist(uint32_t u, uint32_t v) {
    return (((u & 0x80000000u) | (v & 0x7fffffffu)) >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
  }
  void next_state() {
    for (int k = 0; k < 624 - 397; ++k) state[k] = state[k + 397] ^ twist(state[k], state[k + 1]);
    for (int k = 624 - 397; k < 623; ++k) state[k] = state[k + 397 - 624] ^ twist(state[k], state[k + 1]);
    state[623] = state[396] ^ twist(state[623], state[0]);
    cur = 0;
  }
  int32_t next_int() {
    if (cur >= 624) next_state();
    uint32_t x = state[cur++];
    x ^= x >> 11;
    x ^= (x << 7) & 0x9d2c5680u;
    x ^= (x << 15) & 0xefc60000u;
    x ^= x >> 18;
    return (int32_t)x;
  }
  double next_real() {
    const int32_t i = next_int();
    if (i < 0) {
      const float num = (float)i + 4294967296.0f;  // integer + default real -> default real
      return (double)num / (4294967296.0 - 1.0);
    }
    return (double)i / (4294967296.0 - 1.0);
  }
};

// Whole-call mask in the reference's stream order (sub-column, column, layer).  mask: [nlay][nwords][ncol] words.
inline void mask_mt_host(const double* cldfr /*(nlay, ncol)*/, int ncol, int nlay, int nsub, int nwords, int icld,
                         int seed, std::vector<unsigned>& mask) {
  mask.assign((size_t)nlay * nwords * ncol, 0u);
  if (icld == 0) return;
  MT19937 mt(seed);
  const double cldmin = 1.0e-20;
  ColumnMasker m;
  m.icld = icld;
  for (int s = 0; s < nsub; ++s)
    for (int i = 0; i < ncol; ++i) {
      m.begin_subcolumn();
      double r3 = 0.;
      if (icld == 3) r3 = mt.next_real();
      double cf_below = 0.;
      for (int l = 0; l < nlay; ++l) {
        double cf = cldfr[(size_t)l * ncol + i];
        if (cf < cldmin) cf = 0.;
        const double r = icld == 3 ? r3 : mt.next_real();
        if (m.step(r, l, cf, cf_below)) mask[((size_t)l * nwords + (s >> 5)) * ncol + i] |= 1u << (s & 31);
        cf_below = cf;
      }
    }
}

}  // namespace mcica
}  // namespace cb
