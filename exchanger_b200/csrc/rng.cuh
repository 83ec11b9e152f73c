// Counter-based randomness shared by host and device code of the sampler.
//
// Every draw of the sweep is a pure function of (seed; id, iteration, phase, attribute,
// slot), so results do not depend on the launch shape, the batch structure or the number of
// GPUs (DESIGN.md "sampling specification").  Philox4x32-10 (Salmon et al., SC'11) with
//   key     = (seed_lo, seed_hi)
//   counter = (id, iteration, phase << 16 | attribute, slot)
// and two 52-bit uniforms in (0,1) per call.  It replaces R's global Mersenne-Twister that
// the reference consumes in program order (R::unif_rand at src/alias_sampler.cpp:46,
// src/links.cpp:311, src/records.cpp:99,112,132, src/clust_params.cpp:56,88,112).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define EXB_HD __host__ __device__ __forceinline__
#else
#define EXB_HD inline
#endif

namespace exb {

enum Phase : uint32_t {
  PH_LINK_DECIDE = 1,  // id = record: (new vs existing, choice among existing)
  PH_LINK_NEWATTR = 2, // id = record, attr: attribute of the entity the record would found
  PH_ENT_ATTR = 3,     // id = entity, attr: entity attribute resampling
  PH_DISTORT = 4,      // id = record, attr: (indicator, auxiliary)
  PH_CLUST_U = 5,      // id = i: Pitman-Yor u_i
  PH_CLUST_V = 6,      // id = entity, slot = j: per-cluster v_j
  PH_HOST = 7,         // id = counter: host scalar draws (Beta / Gamma / NegBinom)
  PH_CONC_ZETA = 8,    // id = entity, attr, slot = member index: auxiliary of the DP concentration
  PH_CONC_BETA = 9     // id = entity, attr: the cluster's Beta variate (drawn on the host)
};

EXB_HD void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                          uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 52 random bits -> (0,1): (x + 0.5) / 2^52 is exact in binary64 and never 0 or 1
EXB_HD double u52(uint32_t hi, uint32_t lo) {
  const uint64_t x = ((uint64_t)(hi >> 6) << 26) | (uint64_t)(lo >> 6);
  return ((double)x + 0.5) * (1.0 / 4503599627370496.0);
}

struct U2 { double a, b; };

EXB_HD U2 uniforms(uint64_t seed, uint32_t id, uint32_t iter, uint32_t phase, uint32_t attr, uint32_t slot) {
  uint32_t o[4];
  philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), id, iter, (phase << 16) | (attr & 0xFFFFu), slot, o);
  U2 u;
  u.a = u52(o[0], o[1]);
  u.b = u52(o[2], o[3]);
  return u;
}

// One-uniform draw from an alias table (role of AliasSampler::draw, src/alias_sampler.cpp:44-49)
EXB_HD int alias_draw(const double* __restrict__ prob, const int* __restrict__ alias, int n, double u) {
  const double ru = u * (double)n;
  int k = (int)ru;
  if (k >= n) k = n - 1;
  return (ru - (double)k) < prob[k] ? k : alias[k];
}

} // namespace exb
