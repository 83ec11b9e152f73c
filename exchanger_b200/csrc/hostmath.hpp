// Host-side scalar work of the sweep: the Beta / Gamma / negative-binomial draws that the
// reference takes from R's nmath (R::rbeta, R::rgamma, R::rnbinom at
// src/clust_params.cpp:45,63,70,95,100,119 and src/distortion_probs.cpp:43) and the alias
// table construction of src/alias_sampler.cpp:3-42.  These stay scalar host code by design
// (BASELINE.json north_star): a few draws per sweep fed by device reductions.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

#include "rng.cuh"

namespace exb {

// Sequential view of the keyed generator for host draws: uniform number j of iteration `it`
// is component (j & 1) of uniforms(seed, j >> 1, it, PH_HOST, 0, 0).
struct HostStream {
  uint64_t seed = 0;
  uint32_t iteration = 0;
  uint32_t counter = 0;
  void begin_sweep(uint32_t it) { iteration = it; counter = 0; }
  double next() {
    const uint32_t j = counter++;
    const U2 u = uniforms(seed, j >> 1, iteration, PH_HOST, 0, 0);
    return (j & 1u) ? u.b : u.a;
  }
  double normal() { // Box-Muller, one variate per pair
    const double u1 = next(), u2 = next();
    return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586476925286766559 * u2);
  }
  // Marsaglia & Tsang (2000); shape < 1 through the U^(1/shape) boost
  double gamma(double shape, double scale) {
    if (!(shape > 0.0)) return 0.0;
    if (shape < 1.0) {
      const double u = next();
      return gamma(shape + 1.0, scale) * std::pow(u, 1.0 / shape);
    }
    const double d = shape - 1.0 / 3.0, c = 1.0 / std::sqrt(9.0 * d);
    for (;;) {
      const double x = normal();
      double v = 1.0 + c * x;
      if (v <= 0.0) continue;
      v = v * v * v;
      const double u = next(), x2 = x * x;
      if (u < 1.0 - 0.0331 * x2 * x2) return d * v * scale;
      if (std::log(u) < 0.5 * x2 + d * (1.0 - v + std::log(v))) return d * v * scale;
    }
  }
  // R semantics at the edges: rbeta(a, 0) = 1, rbeta(0, b) = 0
  double beta(double a, double b) {
    if (b == 0.0) return 1.0;
    if (a == 0.0) return 0.0;
    const double x = gamma(a, 1.0), y = gamma(b, 1.0);
    return x / (x + y);
  }
  double poisson(double lam) {
    if (!(lam > 0.0)) return 0.0;
    if (lam < 10.0) { // multiplication method
      const double L = std::exp(-lam);
      double p = 1.0;
      long k = 0;
      do { k++; p *= next(); } while (p > L);
      return (double)(k - 1);
    }
    // transformed rejection (Hoermann 1993, PTRS)
    const double slam = std::sqrt(lam), loglam = std::log(lam);
    const double b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
    const double invalpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2.0);
    for (;;) {
      const double U = next() - 0.5, V = next(), us = 0.5 - std::fabs(U);
      const double k = std::floor((2.0 * a / us + b) * U + lam + 0.43);
      if (us >= 0.07 && V <= vr) return k;
      if (k < 0.0 || (us < 0.013 && V > us)) continue;
      if (std::log(V) + std::log(invalpha) - std::log(a / (us * us) + b) <=
          -lam + k * loglam - std::lgamma(k + 1.0))
        return k;
    }
  }
  double negbinom(double size, double prob) { // gamma-Poisson mixture
    if (prob >= 1.0) return 0.0;
    return poisson(gamma(size, (1.0 - prob) / prob));
  }
};

// Gamma / Beta variates from keyed uniforms: one stream per (id, attribute), the slot advances.
// Used for the per-cluster Beta(b, n) variates of DistortDistConcs::update
// (src/distort_dist_concs.cpp:87), which must not depend on the order the clusters are visited in.
struct KeyedStream {
  uint64_t seed; uint32_t id, iter, phase, attr;
  uint32_t slot = 0; bool half = false; double ub = 0.0;
  double next() {
    if (half) { half = false; return ub; }
    const U2 u = uniforms(seed, id, iter, phase, attr, slot++);
    ub = u.b; half = true;
    return u.a;
  }
  double normal() {
    const double u1 = next(), u2 = next();
    return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586476925286766559 * u2);
  }
  double gamma(double shape) {
    if (!(shape > 0.0)) return 0.0;
    if (shape < 1.0) { const double u = next(); return gamma(shape + 1.0) * std::pow(u, 1.0 / shape); }
    const double d = shape - 1.0 / 3.0, c = 1.0 / std::sqrt(9.0 * d);
    for (;;) {
      const double x = normal();
      double v = 1.0 + c * x;
      if (v <= 0.0) continue;
      v = v * v * v;
      const double u = next(), x2 = x * x;
      if (u < 1.0 - 0.0331 * x2 * x2) return d * v;
      if (std::log(u) < 0.5 * x2 + d * (1.0 - v + std::log(v))) return d * v;
    }
  }
  double beta(double a, double b) {
    if (b == 0.0) return 1.0;
    if (a == 0.0) return 0.0;
    const double x = gamma(a), y = gamma(b);
    return x / (x + y);
  }
};

// Vose's alias construction.  Returns false for the weights AliasSampler rejects
// (src/alias_sampler.h:76-97: negative / non-finite weight, non-positive total).
inline bool alias_build(const double* w, int n, double* prob, int* alias) {
  double total = 0.0;
  for (int i = 0; i < n; i++) {
    if (!(w[i] >= 0.0) || !std::isfinite(w[i])) return false;
    total += w[i];
  }
  if (!(total > 0.0) || !std::isfinite(total)) return false;
  std::vector<double> p(n);
  std::vector<int> small, large;
  small.reserve(n);
  large.reserve(n);
  for (int i = 0; i < n; i++) {
    p[i] = w[i] * (double)n / total;
    (p[i] < 1.0 ? small : large).push_back(i);
  }
  while (!small.empty() && !large.empty()) {
    const int s = small.back(), l = large.back();
    small.pop_back();
    large.pop_back();
    prob[s] = p[s];
    alias[s] = l;
    p[l] = (p[l] + p[s]) - 1.0;
    (p[l] < 1.0 ? small : large).push_back(l);
  }
  for (int l : large) { prob[l] = 1.0; alias[l] = l; }
  for (int s : small) { prob[s] = 1.0; alias[s] = s; }
  return true;
}

} // namespace exb
