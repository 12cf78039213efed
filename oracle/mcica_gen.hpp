// ORACLE — TEST INFRASTRUCTURE ONLY.  McICA sub-column generator shared by the LW and SW restatements.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace orc {

// MersenneTwister (mcica_random_numbers.f90:60-300), signed 32-bit arithmetic as in the Fortran
struct MT19937 {
  int32_t state[624];
  int cur;
  static int32_t shr(int32_t x, int n) { return (int32_t)((uint32_t)x >> n); }
  static int32_t shl(int32_t x, int n) { return (int32_t)((uint32_t)x << n); }
  explicit MT19937(int32_t seed) {  // initialize_scalar :169-185
    state[0] = seed;
    for (int i = 1; i < 624; ++i) {
      uint32_t prev = (uint32_t)state[i - 1];
      state[i] = (int32_t)(1812433253u * (prev ^ (prev >> 30)) + (uint32_t)i);
    }
    cur = 624;
  }
  static int32_t mixbits(int32_t u, int32_t v) { return (int32_t)(((uint32_t)u & 0x80000000u) | ((uint32_t)v & 0x7fffffffu)); }
  static int32_t twist(int32_t u, int32_t v) {
    const int32_t t_matrix[2] = {0, (int32_t)0x9908b0dfu};
    return shr(mixbits(u, v), 1) ^ t_matrix[v & 1];
  }
  void nextState() {  // :116-133
    const int M = 397, N = 624;
    for (int k = 0; k <= N - M - 1; ++k) state[k] = state[k + M] ^ twist(state[k], state[k + 1]);
    for (int k = N - M; k <= N - 2; ++k) state[k] = state[k + M - N] ^ twist(state[k], state[k + 1]);
    state[N - 1] = state[M - 1] ^ twist(state[N - 1], state[0]);
    cur = 0;
  }
  static int32_t temper(int32_t y) {  // :154-165
    int32_t x = y ^ shr(y, 11);
    x = x ^ (shl(x, 7) & (int32_t)0x9d2c5680u);
    x = x ^ (shl(x, 15) & (int32_t)0xefc60000u);
    return x ^ shr(x, 18);
  }
  int32_t getRandomInt() {
    if (cur >= 624) nextState();
    return temper(state[cur++]);
  }
  double getRandomReal() {  // :276-295 -- note the default-real (float32) numerator for negative integers
    int32_t localInt = getRandomInt();
    if (localInt < 0) {
      float num = (float)localInt + 4294967296.0f;
      return (double)num / (4294967296.0 - 1.0);
    }
    return (double)localInt / (4294967296.0 - 1.0);
  }
};

// kissvec (mcica_subcol_gen_lw.f90:530-562) for one column, wrap-around int32 arithmetic
struct Kiss {
  int32_t s1, s2, s3, s4;
  static int32_t m(int32_t k, int n) {
    uint32_t u = (uint32_t)k;
    uint32_t sh = n >= 0 ? (u << n) : (u >> (-n));
    return (int32_t)(u ^ sh);
  }
  double next() {
    s1 = (int32_t)(69069u * (uint32_t)s1 + 1327217885u);
    s2 = m(m(m(s2, 13), -17), 5);
    s3 = (int32_t)(18000u * ((uint32_t)s3 & 65535u) + ((uint32_t)s3 >> 16));
    s4 = (int32_t)(30903u * ((uint32_t)s4 & 65535u) + ((uint32_t)s4 >> 16));
    int32_t kiss = (int32_t)((uint32_t)s1 + (uint32_t)s2 + ((uint32_t)s3 << 16) + (uint32_t)s4);
    return kiss * 2.328306e-10 + 0.5;
  }
};

// generate_stochastic_clouds (mcica_subcol_gen_lw.f90:156-522).  Arrays: cld/clwp/ciwp (ncol,nlay) column-fastest,
// tauc (nb, ncol, nlay); outputs (nsub, ncol, nlay) with nsub fastest.  ngb maps sub-column -> band (1-based).
inline int generate_stochastic_clouds(int ncol, int nlay, int nsub, int icld, int irng, const double* pmid,
                                      const double* cld, const double* clwp, const double* ciwp, const double* tauc,
                                      int nb, const int* ngb, int bandoff, double* cld_stoch, double* clwp_stoch,
                                      double* ciwp_stoch, double* tauc_stoch, int changeSeed, std::string& err,
                                      // shortwave extras (mcica_subcol_gen_sw.f90:172-555), null for the longwave
                                      const double* ssac, const double* asmc, const double* fsfc, double* ssac_stoch,
                                      double* asmc_stoch, double* fsfc_stoch) {
  const double cldmin = 1.0e-20;
  if (irng != 0) irng = 1;
  auto C2 = [&](const double* a, int i, int l) { return a[(size_t)i + (size_t)ncol * l]; };  // 0-based
  std::vector<double> cldf((size_t)ncol * nlay);
  for (int l = 0; l < nlay; ++l)
    for (int i = 0; i < ncol; ++i) {
      double v = C2(cld, i, l);
      cldf[(size_t)i + (size_t)ncol * l] = v < cldmin ? 0. : v;
    }
  std::vector<double> CDF((size_t)nsub * ncol * nlay);
  auto X = [&](int s, int i, int l) -> double& { return CDF[(size_t)s + (size_t)nsub * ((size_t)i + (size_t)ncol * l)]; };
  std::vector<Kiss> ks;
  MT19937 mt(changeSeed);
  if (irng == 0) {
    ks.resize(ncol);
    for (int i = 0; i < ncol; ++i) {
      if (nlay < 4 || C2(pmid, i, 0) < C2(pmid, i, 1)) { err = "MCICA_SUBCOL: KISSVEC SEED GENERATOR REQUIRES PMID FROM BOTTOM FOUR LAYERS."; return 1; }
      auto sd = [&](int l) { double p = C2(pmid, i, l); return (int32_t)((p - (double)(int)p) * 1000000000.0); };
      ks[i].s1 = sd(0); ks[i].s2 = sd(1); ks[i].s3 = sd(2); ks[i].s4 = sd(3);
    }
    for (int k = 1; k <= changeSeed; ++k)
      for (int i = 0; i < ncol; ++i) ks[i].next();
  }
  if (icld == 1 || icld == 2) {
    if (irng == 0) {
      for (int s = 0; s < nsub; ++s)
        for (int l = 0; l < nlay; ++l)
          for (int i = 0; i < ncol; ++i) X(s, i, l) = ks[i].next();
    } else {
      for (int s = 0; s < nsub; ++s)
        for (int i = 0; i < ncol; ++i)
          for (int l = 0; l < nlay; ++l) X(s, i, l) = mt.getRandomReal();
    }
    if (icld == 2)
      for (int l = 1; l < nlay; ++l)
        for (int i = 0; i < ncol; ++i)
          for (int s = 0; s < nsub; ++s) {
            double cf = cldf[(size_t)i + (size_t)ncol * (l - 1)];
            if (X(s, i, l - 1) > 1. - cf) X(s, i, l) = X(s, i, l - 1);
            else X(s, i, l) = X(s, i, l) * (1. - cf);
          }
  } else if (icld == 3) {
    if (irng == 0) {
      for (int s = 0; s < nsub; ++s)
        for (int i = 0; i < ncol; ++i) {
          double r = ks[i].next();
          for (int l = 0; l < nlay; ++l) X(s, i, l) = r;
        }
    } else {
      for (int s = 0; s < nsub; ++s)
        for (int i = 0; i < ncol; ++i) {
          double r = mt.getRandomReal();
          for (int l = 0; l < nlay; ++l) X(s, i, l) = r;
        }
    }
  }
  for (int l = 0; l < nlay; ++l)
    for (int i = 0; i < ncol; ++i)
      for (int s = 0; s < nsub; ++s) {
        size_t o = (size_t)s + (size_t)nsub * ((size_t)i + (size_t)ncol * l);
        bool cloudy = X(s, i, l) >= 1. - cldf[(size_t)i + (size_t)ncol * l];
        if (cloudy) {
          cld_stoch[o] = 1.;
          clwp_stoch[o] = C2(clwp, i, l);
          ciwp_stoch[o] = C2(ciwp, i, l);
          int n = ngb[s] - bandoff;  // 1-based band within the tauc array
          const size_t ob = (size_t)(n - 1) + (size_t)nb * ((size_t)i + (size_t)ncol * l);
          tauc_stoch[o] = tauc[ob];
          if (ssac) { ssac_stoch[o] = ssac[ob]; asmc_stoch[o] = asmc[ob]; fsfc_stoch[o] = fsfc[ob]; }
        } else {
          cld_stoch[o] = 0.; clwp_stoch[o] = 0.; ciwp_stoch[o] = 0.; tauc_stoch[o] = 0.;
          if (ssac) { ssac_stoch[o] = 1.; asmc_stoch[o] = 0.; fsfc_stoch[o] = 0.; }
        }
      }
  return 0;
}


}  // namespace orc
