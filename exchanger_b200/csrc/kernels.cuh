// sm_100a kernels of the Gibbs sweep.  Each kernel names the reference function it replaces.
#pragma once
#include <cooperative_groups.h>
#include <cooperative_groups/scan.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "device_state.cuh"
#include "rng.cuh"

namespace cg = cooperative_groups;

namespace exb {

#define FULL 0xFFFFFFFFu
constexpr int TPB = 256;

__device__ __forceinline__ void dev_fail(const DevState& S, int code, int id) {
  if (atomicCAS(&S.dev_err[0], 0, code) == 0) S.dev_err[1] = id;
}

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// bucket of an attribute tuple y under the key mask of table t (mixed-radix key over the key
// attributes in the mask; hashed when the key space is larger than the bucket array)
// (attributes `ea` / `eb`, if any, take the values `ev` / `fv` instead of y[.])
__device__ __forceinline__ uint32_t bucket_of(const DevState& S, const DevTable& t, const int* y, int ea = -1, int ev = 0,
                                              int eb = -1, int fv = 0) {
  uint64_t key = 0;
  for (uint32_t m = t.mask; m; m &= m - 1) {
    const int a = S.key_attr[__ffs(m) - 1];
    key = key * (uint64_t)S.attrs[a].D + (uint64_t)(a == ea ? ev : (a == eb ? fv : y[a]));
  }
  return t.exact ? (uint32_t)key : (uint32_t)(mix64(key) & (uint64_t)(t.B - 1));
}
// A record may be served by a table with ONE key attribute it cannot pin (missing or distorted):
// it then probes the bucket of every value of that attribute (a few dozen for the categorical
// ones), which saves building a table of its own for a rare key mask.  Returns that attribute or -1.
// The values of an unpinned attribute whose buckets the record probes: when the record's value is
// observed but distorted and the attribute has a neighbourhood index (string attribute, infinite
// concentration), only the entity values v with p(x|v) > 0 can give a possible link - the
// neighbourhood row of x (what getPossibleLinks unions, links.cpp:132-171); otherwise every value of
// the (small) domain.
struct ProbeValues {
  const int* col; // row entries (nullptr: the whole domain 0..n-1)
  int n;
  __device__ __forceinline__ int value(int j) const { return col ? col[j] : j; }
};
__device__ __forceinline__ ProbeValues probe_values(const DevState& S, int ea, const int* x, uint32_t zm) {
  const DevAttr& t = S.attrs[ea];
  ProbeValues p;
  if (x[ea] >= 0 && ((zm >> ea) & 1u) && !t.is_const && !((S.dp_mask >> ea) & 1u)) {
    p.col = t.col + t.row[x[ea]];
    p.n = t.row[x[ea] + 1] - t.row[x[ea]];
  } else { p.col = nullptr; p.n = t.D; }
  return p;
}
// The probes of one record in one table: up to two unpinned key attributes, every combination of
// their possible values (n_un = 3: more than two, or too many combinations - the table cannot
// serve the record and its possible links are listed by the batched scan over all the items).
constexpr int MAX_PROBES = 16384;
constexpr int FAST_PROBES = 32;  // probes one thread of the fast candidate kernel makes for a record
constexpr int PROBE_NOT_SERVED = 4; // Probe::n_un: more than three unpinned key attributes, or too many combinations
struct Pv { int v1, v2, v3; };      // the probed values of the (up to three) unpinned key attributes
struct Probe {
  int n_un, a1, a2, a3, count;
  ProbeValues p1, p2, p3;
  __device__ __forceinline__ Pv values(int j) const {
    Pv v;
    v.v1 = p1.value(j % p1.n);
    v.v2 = n_un >= 2 ? p2.value((j / p1.n) % p2.n) : 0;
    v.v3 = n_un >= 3 ? p3.value(j / (p1.n * p2.n)) : 0;
    return v;
  }
  __device__ __forceinline__ int value_of(int a, const int* x, const Pv& v) const { // x with the probed values put in
    return a == a1 ? v.v1 : (n_un >= 2 && a == a2 ? v.v2 : (n_un >= 3 && a == a3 ? v.v3 : x[a]));
  }
  __device__ __forceinline__ uint32_t bucket(const DevState& S, const DevTable& t, const int* x, const Pv& v) const {
    uint64_t key = 0;
    for (uint32_t m = t.mask; m; m &= m - 1) {
      const int a = S.key_attr[__ffs(m) - 1];
      key = key * (uint64_t)S.attrs[a].D + (uint64_t)value_of(a, x, v);
    }
    return t.exact ? (uint32_t)key : (uint32_t)(mix64(key) & (uint64_t)(t.B - 1));
  }
  __device__ __forceinline__ bool own_probe(const int* y, const Pv& v) const { // hashed buckets: an item counts in the probe of its own values only
    return y[a1] == v.v1 && (n_un < 2 || y[a2] == v.v2) && (n_un < 3 || y[a3] == v.v3);
  }
};
__device__ __forceinline__ Probe make_probe(const DevState& S, uint32_t mask, const int* x, uint32_t zm) {
  Probe p;
  p.n_un = 0; p.a1 = p.a2 = p.a3 = -1; p.count = 1;
  p.p1.col = p.p2.col = p.p3.col = nullptr; p.p1.n = p.p2.n = p.p3.n = 1;
  for (uint32_t m = mask; m; m &= m - 1) {
    const int a = S.key_attr[__ffs(m) - 1];
    if (x[a] >= 0 && !((zm >> a) & 1u)) continue;
    if (p.n_un == 0) { p.a1 = a; p.p1 = probe_values(S, a, x, zm); }
    else if (p.n_un == 1) { p.a2 = a; p.p2 = probe_values(S, a, x, zm); }
    else if (p.n_un == 2) { p.a3 = a; p.p3 = probe_values(S, a, x, zm); }
    p.n_un++;
  }
  if (p.n_un > 3) p.n_un = PROBE_NOT_SERVED;
  else if (p.n_un >= 1) {
    const long long c = (long long)p.p1.n * (long long)p.p2.n * (long long)p.p3.n;
    if (c > MAX_PROBES) p.n_un = PROBE_NOT_SERVED; else p.count = (int)c;
  }
  return p;
}

// The same for the listing kernel of the serial path (k_wide_emit), which spreads the probes of one
// record over several blocks and can therefore afford many more of them: up to four unpinned key
// attributes (every key attribute of the largest tables), the count in 64 bits.
constexpr int BIG_UNPINNED = 4;
struct BigProbe {
  int n_un;           // unpinned key attributes (BIG_UNPINNED + 1: not served)
  int a[BIG_UNPINNED];
  ProbeValues p[BIG_UNPINNED];
  long long count;
  __device__ __forceinline__ void values(long long j, int* v) const {
    for (int k = 0; k < n_un; k++) { v[k] = p[k].value((int)(j % p[k].n)); j /= p[k].n; }
  }
  __device__ __forceinline__ int value_of(int attr, const int* x, const int* v) const {
    for (int k = 0; k < n_un; k++) if (a[k] == attr) return v[k];
    return x[attr];
  }
  __device__ __forceinline__ uint32_t bucket(const DevState& S, const DevTable& t, const int* x, const int* v) const {
    uint64_t key = 0;
    for (uint32_t m = t.mask; m; m &= m - 1) {
      const int at = S.key_attr[__ffs(m) - 1];
      key = key * (uint64_t)S.attrs[at].D + (uint64_t)value_of(at, x, v);
    }
    return t.exact ? (uint32_t)key : (uint32_t)(mix64(key) & (uint64_t)(t.B - 1));
  }
  __device__ __forceinline__ uint32_t cell(const DevState& S, const DevTable& t, const int* x, const int* v) const {
    unsigned long long K = 0;
    for (uint32_t m = t.mask; m; m &= m - 1) {
      const int at = S.key_attr[__ffs(m) - 1];
      K |= (unsigned long long)(uint32_t)value_of(at, x, v) << S.pk_shift[at];
    }
    return (uint32_t)(mix64(K & t.kmask) & (unsigned long long)(t.cells - 1u));
  }
  __device__ __forceinline__ bool own_probe(const int* y, const int* v) const { // an item counts in the probe of its own values only
    bool ok = true;
    for (int k = 0; k < n_un; k++) ok &= y[a[k]] == v[k];
    return ok;
  }
};
__device__ __forceinline__ BigProbe make_big_probe(const DevState& S, uint32_t mask, const int* x, uint32_t zm) {
  BigProbe p;
  p.n_un = 0; p.count = 1;
  for (int k = 0; k < BIG_UNPINNED; k++) { p.a[k] = -1; p.p[k].col = nullptr; p.p[k].n = 1; }
  for (uint32_t m = mask; m; m &= m - 1) {
    const int a = S.key_attr[__ffs(m) - 1];
    if (x[a] >= 0 && !((zm >> a) & 1u)) continue;
    if (p.n_un < BIG_UNPINNED) {
      p.a[p.n_un] = a; p.p[p.n_un] = probe_values(S, a, x, zm);
      p.count = p.count > (1ll << 40) ? p.count : p.count * (long long)p.p[p.n_un].n;
    }
    p.n_un++;
  }
  return p;
}
// What listing the possible links of record x through table t would cost (about one unit per cell or
// bucket item read), for the choice of a table for a record that its own table does not serve.
__device__ __forceinline__ double big_probe_cost(const DevState& S, const DevTable& t, const int* x, uint32_t zm) {
  const BigProbe p = make_big_probe(S, t.mask, x, zm);
  if (p.n_un > BIG_UNPINNED) return CUDART_INF;
  if (t.chain) return 2.0 * (double)p.count;
  return (double)p.count * (1.0 + 1.2 * (double)S.N / (double)t.B);
}

// ---- packed attribute tuples (fast candidate path; S.pk_ok: the domains fit 64 bits and A <= 8)
__device__ __forceinline__ unsigned long long pk_field(const DevState& S, int a) {
  return ((1ull << S.pk_bits[a]) - 1ull) << S.pk_shift[a];
}
__device__ __forceinline__ unsigned long long pack_tuple(const DevState& S, const int* y) {
  unsigned long long P = 0;
  for (int a = 0; a < S.A; a++) P |= (unsigned long long)(uint32_t)y[a] << S.pk_shift[a];
  return P;
}
__device__ __forceinline__ int unpack_attr(const DevState& S, unsigned long long P, int a) {
  return (int)((P >> S.pk_shift[a]) & ((1ull << S.pk_bits[a]) - 1ull));
}
// fields (M) and values (V) of the attributes a record pins: an item with packed tuple P agrees with
// the record on every observed non-distorted attribute (links.cpp:122-131) iff ((P ^ V) & M) == 0
__device__ __forceinline__ void record_fields(const DevState& S, const int* x, uint32_t zm, unsigned long long& M,
                                              unsigned long long& V) {
  M = 0; V = 0;
  for (int a = 0; a < S.A; a++)
    if (x[a] >= 0 && !((zm >> a) & 1u)) { M |= pk_field(S, a); V |= (unsigned long long)(uint32_t)x[a] << S.pk_shift[a]; }
}
// cell of a chain table for key fields K (the packed tuple restricted to the table's key attributes)
__device__ __forceinline__ uint32_t chain_cell(const DevTable& t, unsigned long long K) {
  return (uint32_t)(mix64(K & t.kmask) & (unsigned long long)(t.cells - 1u));
}
// ... for a record x whose unpinned key attributes take the probed values v
__device__ __forceinline__ uint32_t chain_cell_probe(const DevState& S, const DevTable& t, const int* x, const Probe& pr, const Pv& v) {
  unsigned long long K = 0;
  for (uint32_t m = t.mask; m; m &= m - 1) {
    const int a = S.key_attr[__ffs(m) - 1];
    K |= (unsigned long long)(uint32_t)pr.value_of(a, x, v) << S.pk_shift[a];
  }
  return chain_cell(t, K);
}
__device__ __forceinline__ ChainNode load_node(const ChainNode* p) {
  const int4 v = *reinterpret_cast<const int4*>(p);
  ChainNode n;
  n.P = ((unsigned long long)(uint32_t)v.y << 32) | (unsigned long long)(uint32_t)v.x;
  n.next = v.z; n.flags = (uint32_t)v.w;
  return n;
}

// log p(w | v) = log get_distortion_prob(v, w)   (src/attribute_index.cpp:36-47, 77-92)
__device__ __forceinline__ double log_distortion_prob(const DevAttr& t, int v, int w) {
  if (t.is_const) {
    if (t.exclude) return v == w ? -CUDART_INF : t.logpsi[w] - t.log1mpsi[v];
    return t.logpsi[w];
  }
  if (v == w) return t.exclude ? -CUDART_INF : 0.0;
  int lo = t.row[w], hi = t.row[w + 1];
  const int end = hi;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (t.col[mid] < v) lo = mid + 1; else hi = mid;
  }
  return (lo < end && t.col[lo] == v) ? t.logp[lo] : -CUDART_INF;
}

// p(w | v) itself (only used with a finite concentration, which implies exclude_entity_value)
__device__ __forceinline__ double distortion_prob(const DevAttr& t, int v, int w) {
  if (v == w) return t.exclude ? 0.0 : (t.is_const ? t.psi[w] : 1.0);
  if (t.is_const) return t.exclude ? t.psi[w] / t.om[v] : t.psi[w];
  int lo = t.row[w], hi = t.row[w + 1];
  const int end = hi;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (t.col[mid] < v) lo = mid + 1; else hi = mid;
  }
  return (lo < end && t.col[lo] == v) ? t.p[lo] : 0.0;
}

__device__ __forceinline__ double prior_existing(const ClustDev& c, int n) { // clust_params.cpp:4-27
  if (c.kind == EXB200_CLUST_PITMAN_YOR) return fabs((double)n - c.d);
  if (c.kind == EXB200_CLUST_GEN_COUPON) return (double)n + c.kappa;
  return 1.0;
}
__device__ __forceinline__ double log_prior_existing(const ClustDev& c, int n) {
  return n < LPE_MAX ? c.logpe[n] : log(prior_existing(c, n));
}
__device__ __forceinline__ double prior_new(const ClustDev& c, int k) { // clust_params.cpp:9-31
  if (c.kind == EXB200_CLUST_PITMAN_YOR) return c.alpha + c.d * (double)k;
  if (c.kind == EXB200_CLUST_GEN_COUPON) return c.kappa * fabs(c.m - (double)k);
  return fabs(c.m - (double)k);
}

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------
// theta-dependent log tables (per sweep).  lt_diff[f][v] = log(theta[f,a] mef(v)) and
// lt_same[f][v] = log(1 - theta mef(v))  or  log(1 - eff + eff p(v|v)): the two per-record
// factors of logWeightEntityValue (src/entities.cpp:264-274).
// ------------------------------------------------------------------------------------------
__global__ void k_theta_tables(DevState S, int a) {
  const DevAttr& t = S.attrs[a];
  const int n = S.F * t.D;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int f = i / t.D, v = i - f * t.D;
    const double eff = S.theta[f * S.A + a] * t.mef[v];
    t.lt_diff[i] = log(eff);
    t.lt_same[i] = t.exclude ? log(1.0 - eff) : log(1.0 - eff + eff * (t.is_const ? t.psi[v] : 1.0));
    if (t.exclude) {
      // An entity with one observed member value x (a singleton above all): the two weights its draw
      // compares - the very expressions of draw_entity_value for nobs = 1, so that k_entities can take
      // "keep x" from this table and leave everything else to that function.
      const double th = S.theta[f * S.A + a];
      double keep, other;
      if (t.is_const) {
        double Cc = 1.0, ww = t.phi[v], rest = t.G[1];
        Cc *= th * t.psi[v];
        ww *= 1.0 - th;
        rest -= t.g[(size_t)t.D + v];
        if (rest < 0.0) rest = 0.0;
        keep = ww; other = Cc * rest;
      } else {
        double thp = 1.0, wx = t.phi[v];
        thp *= th;
        wx *= 1.0 - th * t.mef[v];
        const int lo = t.row[v], hi = t.row[v + 1];
        const double rsum = hi > lo ? t.cs[hi - 1] : 0.0;
        keep = wx; other = thp * rsum;
      }
      t.sg[2 * (size_t)i] = keep; t.sg[2 * (size_t)i + 1] = other;
    }
  }
}

// ------------------------------------------------------------------------------------------
// K1a  per-record preparation of the links update.
//   * attributes of the entity the record would found (src/links.cpp:424-468) -> slot N + r
//   * logLikelihoodWeightNew (src/links.cpp:316-347)
//   * choice of the blocking key: the two observed non-distorted attributes whose record values
//     are rarest, or all of them when that pair is expected to match many items (replaces the
//     set-size ordering of getPossibleLinks, links.cpp:187-189)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB, 4) k_prep(DevState S) {
  extern __shared__ int s_prep[]; // [2][2^n_key] usage counts, [2^n_key] expected visits (float)
  const int AA = 1 << S.n_key;
  int* s_usage = s_prep;
  float* s_est = reinterpret_cast<float*>(s_prep + 2 * AA);
  for (int i = threadIdx.x; i < AA; i += blockDim.x) { s_usage[i] = 0; s_usage[AA + i] = 0; s_est[i] = 0.f; }
  __syncthreads();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < S.N) {
    const int* x = S.rec_x + (size_t)r * S.RS;
    const uint32_t zm = S.rec_z[r];
    int* yn = S.ent_y + (size_t)(S.N + r) * S.ES;
    double L = 0.0;
    int b1 = -1, b2 = -1;        // key positions of the most selective observed non-distorted attributes
    uint32_t nd_mask = 0;
    double s1 = 2.0, s2 = 2.0;
    bool pinned_string = false;  // some string attribute is observed and not distorted
    int soft = -1, soft_len = 0; // key position of the distorted string attribute with the shortest neighbourhood row
    for (int a = 0; a < S.A; a++) {
      const DevAttr& t = S.attrs[a];
      const int xv = x[a];
      const bool zz = (zm >> a) & 1u;
      int yv;
      if (xv < 0) { // unobserved: prior draw (links.cpp:434-437)
        const U2 u = uniforms(S.seed, r, S.iter, PH_LINK_NEWATTR, a, 0);
        yv = alias_draw(t.aprob, t.aalias, t.D, u.a);
      } else {
        L += zz ? t.logE[xv] : t.logphi[xv];
        if (!zz) { // links.cpp:440-442
          yv = xv;
          if (S.key_pos[a] >= 0) {
            pinned_string |= !t.is_const;
            nd_mask |= 1u << S.key_pos[a];
            const double sel = t.psi[xv];
            if (sel < s1) { s2 = s1; b2 = b1; s1 = sel; b1 = S.key_pos[a]; }
            else if (sel < s2) { s2 = sel; b2 = S.key_pos[a]; }
          }
        } else if (t.is_const) {
          if (t.exclude) { // law of rejectionSampleConstant (links.cpp:290-314): phi/(1-psi), v != x
            yv = xv;
            for (int k = 0; k < REJECT_CAP; k++) {
              const U2 u = uniforms(S.seed, r, S.iter, PH_LINK_NEWATTR, a, k);
              yv = alias_draw(t.aprob + t.D, t.aalias + t.D, t.D, u.a);
              if (yv != xv) break;
            }
            if (yv == xv) dev_fail(S, EXB200_ERR_WEIGHTS, r);
          } else { // links.cpp:449-452
            const U2 u = uniforms(S.seed, r, S.iter, PH_LINK_NEWATTR, a, 0);
            yv = alias_draw(t.pprob, t.palias, t.D, u.a);
          }
        } else { // links.cpp:454-461: phi(v) p(x | v) over the neighbourhood of x (static prefix sums)
          const int lo = t.row[xv], hi = t.row[xv + 1];
          if (S.key_pos[a] >= 0 && !((S.dp_mask >> a) & 1u) && (soft < 0 || hi - lo < soft_len)) { soft = S.key_pos[a]; soft_len = hi - lo; }
          const double tot = hi > lo ? t.csn[hi - 1] : 0.0;
          yv = xv;
          if (tot > 0.0) {
            const U2 u = uniforms(S.seed, r, S.iter, PH_LINK_NEWATTR, a, 0);
            const double tt = u.a * tot;
            int a0 = lo, b0 = hi; // first j with tt < csn[j]
            while (a0 < b0) { const int mid = (a0 + b0) >> 1; if (tt < t.csn[mid]) b0 = mid; else a0 = mid + 1; }
            if (a0 == hi) { // rounding: the last entry of positive weight
              a0 = lo; b0 = hi - 1;
              while (a0 < b0) { const int mid = (a0 + b0) >> 1; if (t.csn[mid] >= tot) b0 = mid; else a0 = mid + 1; }
            }
            yv = t.col[a0];
          }
        }
      }
      yn[a] = yv;
    }
    S.Lnew[r] = L;
    S.ent_size[S.N + r] = 0;
    S.ent_head[S.N + r] = -1;
    S.fin[r] = 0;
    // blocking table
    uint32_t tab = 0, pair = 0;
    if (b1 >= 0 && !pinned_string && soft >= 0 && soft_len <= 1024) {
      // No pinned name, but a distorted one: only entity values in its neighbourhood row can be
      // linked (links.cpp:132-171), so the record probes those buckets of the (name, best pinned
      // attribute) table instead of scanning the long buckets of the date-only tables.
      pair = (1u << b1) | (1u << soft);
      tab = pair;
      atomicAdd(&s_usage[AA + pair], 1);
    } else if (b1 >= 0) {
      pair = (1u << b1) | (b2 >= 0 ? 1u << b2 : 0u);
      const double est = 2.0 * (double)S.N * s1 * (b2 >= 0 ? s2 : 1.0);
      tab = est > S.est_threshold ? nd_mask : pair;
      atomicAdd(&s_usage[AA + pair], 1);
      if (tab != pair) { atomicAdd(&s_usage[tab], 1); atomicAdd(&s_est[tab], (float)est); }
    }
    S.rec_tab[r] = (uint16_t)tab; // the tables actually built are chosen on the host (tab_remap)
    S.rec_alt[r] = (uint16_t)pair;
    // twin: r alone in entity slot r and its potential entity would have the same attributes
    bool tw = S.lambda[r] == r && S.ent_size[r] == 1;
    if (tw) {
      const int* ye = S.ent_y + (size_t)r * S.ES;
      for (int a = 0; a < S.A; a++) tw &= ye[a] == yn[a];
    }
    S.twin[r] = tw;
    S.live0[r] = S.ent_size[r] > 0;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < AA; i += blockDim.x) {
    if (s_usage[i]) atomicAdd(&S.tab_usage[i], s_usage[i]);
    if (s_usage[AA + i]) atomicAdd(&S.tab_usage[N_TABLES + i], s_usage[AA + i]);
    if (s_est[i] != 0.f) atomicAdd(&S.tab_est[i], s_est[i]);
  }
}

// ------------------------------------------------------------------------------------------
// K1b  blocking index over the items (entity slots) of the sweep: live entities [0,N) and the
// N potential entities [N,2N).  Replaces Entities::inverse_index_ (src/entities.h:51): instead
// of one hash set per (attribute, value) kept up to date by every change, CSR buckets over
// attribute pairs are rebuilt once per sweep and stay fixed during the links update.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool item_indexed(const DevState& S, int i) {
  return i < S.N ? S.ent_size[i] > 0 : !S.twin[i - S.N];
}
// The tuple of item i for the index builders, false if the item is not indexed: from the packed tuples
// when they are there (S.pk: 9 bytes per item, coalesced, instead of the 37 of size / twin flag / row)
__device__ __forceinline__ bool load_item(const DevState& S, int i, int* ybuf, const int*& y) {
  if (S.pk) {
    if (!(S.pkfl[i] & 1)) return false;
    const unsigned long long P = S.pk[i];
    for (int a = 0; a < S.A; a++) ybuf[a] = (int)((P >> S.pk_shift[a]) & ((1ull << S.pk_bits[a]) - 1ull));
    y = ybuf;
    return true;
  }
  if (!(i < S.N ? S.ent_size[i] > 0 : !S.twin[i - S.N])) return false;
  y = S.ent_y + (size_t)i * S.ES;
  return true;
}
__device__ __forceinline__ const int* load_item_ptr(const DevState& S, int i, int* ybuf) { // nullptr: not indexed
  const int* y = nullptr;
  return load_item(S, i, ybuf, y) ? y : nullptr;
}
__global__ void __launch_bounds__(TPB) k_index_count(DevState S, const int* __restrict__ active, int n_active) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * S.N) return;
  int ybuf[8];
  const int* y;
  if (!load_item(S, i, ybuf, y)) return;
  for (int k = 0; k < n_active; k++) {
    const DevTable& t = S.tables[active[k]];
    atomicAdd(&t.ptr[bucket_of(S, t, y)], 1);
  }
}
__global__ void __launch_bounds__(TPB) k_index_fill(DevState S, const int* __restrict__ active, int n_active) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * S.N) return;
  int ybuf[8];
  const int* y;
  if (!load_item(S, i, ybuf, y)) return;
  for (int k = 0; k < n_active; k++) {
    const DevTable& t = S.tables[active[k]];
    const int pos = atomicAdd(&t.ptr[bucket_of(S, t, y)], 1);
    t.ids[pos] = i;
  }
}

// Tables with few buckets (date pairs and triples, single attributes): all items fall on a few
// thousand counters, so the counts are aggregated per block in shared memory.  phase 0 adds the
// block histogram to ptr[]; after the exclusive scan, phase 1 reserves one slice per (block, bucket)
// with a single atomicAdd on ptr[b] (which thereby ends as the bucket's END offset) and places the ids.
constexpr int SMALL_B_MAX = 8192;
constexpr int SMALL_TABS_MAX = 16;      // tables per launch ...
constexpr int SMALL_SMEM_INTS = 49152;  // ... and their buckets together (192 KB of shared memory)
struct SmallTabs { int n; int tab[SMALL_TABS_MAX]; int off[SMALL_TABS_MAX + 1]; }; // off: start of each table's counters
// All the given tables in one pass over the item tuples (every item is read once per phase, whatever
// the number of tables).
__global__ void __launch_bounds__(TPB) k_index_small(DevState S, SmallTabs T, int phase) {
  extern __shared__ int s_dyn[]; // [off[n]]
  const int total = T.off[T.n];
  const int n = 2 * S.N;
  const int per = (n + gridDim.x - 1) / gridDim.x;
  const int i0 = blockIdx.x * per, i1 = min(n, i0 + per);
  for (int b = threadIdx.x; b < total; b += TPB) s_dyn[b] = 0;
  __syncthreads();
  for (int i = i0 + threadIdx.x; i < i1; i += TPB)
    if (int ybuf[8]; const int* y = load_item_ptr(S, i, ybuf)) {
      for (int k = 0; k < T.n; k++) atomicAdd(&s_dyn[T.off[k] + bucket_of(S, S.tables[T.tab[k]], y)], 1);
    }
  __syncthreads();
  for (int q = threadIdx.x; q < total; q += TPB)
    if (s_dyn[q]) {
      int k = 0;
      while (q >= T.off[k + 1]) k++;
      s_dyn[q] = atomicAdd(&S.tables[T.tab[k]].ptr[q - T.off[k]], s_dyn[q]); // phase 0: value unused; phase 1: slice base
    }
  if (phase == 0) return;
  __syncthreads();
  for (int i = i0 + threadIdx.x; i < i1; i += TPB)
    if (int ybuf[8]; const int* y = load_item_ptr(S, i, ybuf)) {
      for (int k = 0; k < T.n; k++) {
        const DevTable& t = S.tables[T.tab[k]];
        t.ids[atomicAdd(&s_dyn[T.off[k] + bucket_of(S, t, y)], 1)] = i;
      }
    }
}

// ------------------------------------------------------------------------------------------
// K1b'  partitioned build of the large blocking tables.  Random global atomics (one per item,
// table and pass) bound k_index_count / k_index_fill; for tables with many buckets the CSR is
// built like a two-level bucket sort instead:
//   pass A  every (item, table) pair goes to partition p = bucket >> PART_BITS.  Blocks own
//           contiguous item ranges, histogram their range per table in shared memory and reserve
//           one contiguous slice per (block, partition) - one global atomic per non-empty bin.
//   pass B  one block per (table, partition): the 2^PART_BITS buckets of the partition are
//           counted, scanned and filled entirely in shared memory; ptr / ids are written once.
// The order of the ids inside a bucket is arbitrary (candidate lists are sorted later).
// ------------------------------------------------------------------------------------------
constexpr int PART_BITS = 14;
constexpr int PART_SIZE = 1 << PART_BITS;
constexpr int MAX_PARTS = 1024;          // partitions per table (bucket arrays up to 2^24)
// (pass A and the tiny-table builder run on eight blocks per multiprocessor: the same grid in both phases)
constexpr int PA_MAX_TABS = 24;          // tables per pass-A launch (shared memory: 8 KB each)

struct PartBuild {
  int n_tab;                 // large tables of this launch
  const int* tabs;           // their key masks
  int* part_cnt;             // [n_tab][MAX_PARTS] pairs per partition, then running cursors
  int* part_start;           // [n_tab][MAX_PARTS + 1] exclusive prefix of part_cnt
  unsigned long long* pairs; // [n_tab][cap] (item << PART_BITS | low bucket bits), grouped by partition
  long long cap;             // pairs per table
};


// phase 0: partition histogram; phase 1: scatter of the pairs.  Every item tuple is read once per
// phase for all tables; shared memory holds one histogram (and one slice base) per table.
__global__ void __launch_bounds__(TPB) k_part_a(DevState S, PartBuild P, int phase) {
  extern __shared__ int s_cnt[]; // [n_tab][MAX_PARTS]: counts, then the running position inside the block's slice
  const int n = 2 * S.N;
  const int per = (n + gridDim.x - 1) / gridDim.x;
  const int i0 = blockIdx.x * per, i1 = min(n, i0 + per);
  const int nbins = P.n_tab * MAX_PARTS;
  for (int q = threadIdx.x; q < nbins; q += TPB) s_cnt[q] = 0;
  __syncthreads();
  for (int i = i0 + threadIdx.x; i < i1; i += TPB)
    if (int ybuf[8]; const int* y = load_item_ptr(S, i, ybuf)) {
      for (int k = 0; k < P.n_tab; k++)
        atomicAdd(&s_cnt[k * MAX_PARTS + (bucket_of(S, S.tables[P.tabs[k]], y) >> PART_BITS)], 1);
    }
  __syncthreads();
  if (phase == 0) {
    for (int q = threadIdx.x; q < nbins; q += TPB)
      if (s_cnt[q]) atomicAdd(&P.part_cnt[q], s_cnt[q]);
    return;
  }
  // reserve this block's slice of every partition (one global atomic per non-empty bin), then place
  // the pairs: the shared counter now walks through the slice
  for (int q = threadIdx.x; q < nbins; q += TPB)
    if (s_cnt[q]) s_cnt[q] = atomicAdd(&P.part_cnt[q], s_cnt[q]);
  __syncthreads();
  for (int i = i0 + threadIdx.x; i < i1; i += TPB)
    if (int ybuf[8]; const int* y = load_item_ptr(S, i, ybuf)) {
      for (int k = 0; k < P.n_tab; k++) {
        const uint32_t b = bucket_of(S, S.tables[P.tabs[k]], y);
        const int pos = atomicAdd(&s_cnt[k * MAX_PARTS + (b >> PART_BITS)], 1);
        P.pairs[(size_t)k * P.cap + pos] =
            ((unsigned long long)(uint32_t)i << PART_BITS) | (unsigned long long)(b & (PART_SIZE - 1));
      }
    }
}

// exclusive prefix of the partition counts of every large table; the counts become cursors
__global__ void __launch_bounds__(TPB) k_part_scan(PartBuild P) {
  __shared__ int s_sum[TPB];
  const int k = blockIdx.x;
  int* cnt = P.part_cnt + k * MAX_PARTS;
  int* start = P.part_start + k * (MAX_PARTS + 1);
  constexpr int per = MAX_PARTS / TPB;
  int t = 0;
  for (int j = 0; j < per; j++) t += cnt[threadIdx.x * per + j];
  s_sum[threadIdx.x] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int q = 0; q < TPB; q++) { const int x = s_sum[q]; s_sum[q] = run; run += x; }
    start[MAX_PARTS] = run;
  }
  __syncthreads();
  int run = s_sum[threadIdx.x];
  for (int j = 0; j < per; j++) {
    const int p = threadIdx.x * per + j;
    const int x = cnt[p];
    start[p] = run;
    cnt[p] = run; // cursor of pass A's scatter
    run += x;
  }
}

// pass B: one block per (table, partition); the PART_SIZE bucket counters live in shared memory
__global__ void __launch_bounds__(TPB) k_part_b(DevState S, PartBuild P) {
  extern __shared__ int s_cnt[]; // [PART_SIZE]
  __shared__ int s_w[TPB];
  const int k = blockIdx.y, p = blockIdx.x;
  const DevTable& t = S.tables[P.tabs[k]];
  const int np = (int)((t.B + PART_SIZE - 1) >> PART_BITS);
  if (p >= np) return;
  const int* start = P.part_start + k * (MAX_PARTS + 1);
  const int lo = start[p], hi = start[p + 1];
  const unsigned long long* pairs = P.pairs + (size_t)k * P.cap;
  for (int b = threadIdx.x; b < PART_SIZE; b += TPB) s_cnt[b] = 0;
  __syncthreads();
  // four loads in flight per thread: the kernel waits on these loads and on nothing else
  for (int q0 = lo; q0 < hi; q0 += 4 * TPB) {
    unsigned long long v[4];
#pragma unroll
    for (int u = 0; u < 4; u++) { const int q = q0 + u * TPB + threadIdx.x; v[u] = q < hi ? pairs[q] : ~0ull; }
#pragma unroll
    for (int u = 0; u < 4; u++) if (q0 + u * TPB + (int)threadIdx.x < hi) atomicAdd(&s_cnt[(int)(v[u] & (PART_SIZE - 1))], 1);
  }
  __syncthreads();
  // exclusive scan of the counters, TPB consecutive ones at a time (thread = counter: no bank
  // conflicts), with a running carry; s_cnt[b] becomes the start of bucket b, ptr[b] its end
  const uint32_t b0 = (uint32_t)p << PART_BITS;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int carry = lo;
  for (int c0 = 0; c0 < PART_SIZE; c0 += TPB) {
    const int b = c0 + threadIdx.x;
    const int c = s_cnt[b];
    int x = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULL, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_w[w] = x;
    __syncthreads();
    int off = 0, total = 0;
#pragma unroll
    for (int k2 = 0; k2 < TPB / 32; k2++) { const int sv = s_w[k2]; if (k2 < w) off += sv; total += sv; }
    const int end = carry + off + x;
    s_cnt[b] = end - c;       // start of the bucket: cursor of the fill below
    if (b0 + b < t.B) t.ptr[b0 + b] = end; // END offset of the bucket
    carry += total;
    __syncthreads();
  }
  for (int q0 = lo; q0 < hi; q0 += 4 * TPB) {
    unsigned long long v[4];
#pragma unroll
    for (int u = 0; u < 4; u++) { const int q = q0 + u * TPB + threadIdx.x; v[u] = q < hi ? pairs[q] : ~0ull; }
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (q0 + u * TPB + (int)threadIdx.x < hi) t.ids[atomicAdd(&s_cnt[(int)(v[u] & (PART_SIZE - 1))], 1)] = (int)(v[u] >> PART_BITS);
  }
}

// ------------------------------------------------------------------------------------------
// exclusive prefix sum of n integers (three passes).  T_IN is int or int8_t.
// ------------------------------------------------------------------------------------------
constexpr int SCAN_TILE = TPB * 8;

template <class T_IN>
__global__ void __launch_bounds__(TPB) k_scan_reduce(const T_IN* __restrict__ in, int n, int* __restrict__ bsum) {
  __shared__ int s_w[TPB / 32];
  const int base = blockIdx.x * SCAN_TILE;
  int v = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const int i = base + k * TPB + threadIdx.x;
    if (i < n) v += (int)in[i];
  }
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < TPB / 32; w++) t += s_w[w];
    bsum[blockIdx.x] = t;
  }
}
// one block: exclusive scan of the block sums in place
__global__ void __launch_bounds__(1024) k_scan_top(int* __restrict__ bsum, int nb) {
  __shared__ int s_part[1024];
  const int per = (nb + 1023) / 1024;
  const int lo = threadIdx.x * per, hi = min(nb, lo + per);
  int t = 0;
  for (int i = lo; i < hi; i++) t += bsum[i];
  s_part[threadIdx.x] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int i = 0; i < 1024; i++) { const int x = s_part[i]; s_part[i] = run; run += x; }
  }
  __syncthreads();
  int run = s_part[threadIdx.x];
  for (int i = lo; i < hi; i++) { const int x = bsum[i]; bsum[i] = run; run += x; }
}
template <class T_IN>
__global__ void __launch_bounds__(TPB) k_scan_apply(const T_IN* __restrict__ in, int* __restrict__ out, int n,
                                                    const int* __restrict__ bsum) {
  __shared__ int s_w[TPB / 32];
  const int base = blockIdx.x * SCAN_TILE + threadIdx.x * 8;
  int v[8];
  int t = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    v[k] = base + k < n ? (int)in[base + k] : 0;
    t += v[k];
  }
  // exclusive scan of t over the block
  int incl = t;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(FULL, incl, o);
    if ((threadIdx.x & 31) >= o) incl += y;
  }
  if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
  __syncthreads();
  int off = bsum[blockIdx.x];
  for (int w = 0; w < (int)(threadIdx.x >> 5); w++) off += s_w[w];
  int run = off + incl - t;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    if (base + k < n) out[base + k] = run;
    run += v[k];
  }
}

// ------------------------------------------------------------------------------------------
// K1c  candidate generation: a group of lanes per record - eight in the first launch, a whole warp
// for the records listed for the second (replaces getPossibleLinks, src/links.cpp:90-212, and the
// per-candidate likelihood logLikelihoodWeightExisting, src/links.cpp:216-288).
// The group streams the record's bucket(s), gathers each item's attribute tuple (one 32-byte sector),
// keeps the items that agree with the record on every observed non-distorted attribute
// (links.cpp:122-131) and stores them, sorted by id, with the log-likelihood of the distorted
// attributes.  Attribute tuples do not change during the links update, so this list is computed
// once per sweep; only cluster sizes are read again when the record's turn comes.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool item_compatible(const DevState& S, const int* x, uint32_t zm, const int* y) {
  bool ok = true;
  for (int a = 0; a < S.A; a++) {
    const int xv = x[a];
    ok &= (xv < 0) | ((zm >> a) & 1u) | (y[a] == xv);
  }
  return ok;
}
__device__ __forceinline__ double item_loglik(const DevState& S, const int* x, uint32_t zm, const int* y) {
  double l = 0.0;
  // attributes with a finite concentration depend on the entity's other members: dp_loglik
  const uint32_t zs = zm & ~S.dp_mask;
  for (int a = 0; a < S.A; a++) {
    const int xv = x[a];
    if (xv >= 0 && ((zs >> a) & 1u)) l += log_distortion_prob(S.attrs[a], y[a], xv);
  }
  return l;
}
// The member-dependent part of logLikelihoodWeightExisting for attributes with a finite
// concentration b (links.cpp:251-279): the other members of e whose value is distorted and observed
// form the urn.  Member lists and the members' (x, z) cannot change while r holds e.
__device__ double dp_loglik(const DevState& S, int r, int e) {
  const int* x = S.rec_x + (size_t)r * S.RS;
  const uint32_t zm = S.rec_z[r];
  const int* y = S.ent_y + (size_t)e * S.ES;
  double l = 0.0;
  for (int a = 0; a < S.A; a++) {
    const int xv = x[a];
    if (!(xv >= 0 && ((zm >> a) & 1u)) || isinf(S.conc[a])) continue;
    int c_dist = 0, c_same = 0;
    for (int q = S.ent_head[e]; q >= 0; q = S.rec_next[q]) {
      if (q == r) continue;
      const int xq = S.rec_x[(size_t)q * S.RS + a];
      if (xq >= 0 && ((S.rec_z[q] >> a) & 1u)) { c_dist++; c_same += xq == xv; }
    }
    l += log((double)c_same + S.conc[a] * distortion_prob(S.attrs[a], y[a], xv));
    l += -log((double)c_dist + S.conc[a]);
  }
  return l;
}
// log-weight of candidate e (live size sz) for record r from the stored static likelihood
// (DP: some attribute has a finite concentration - the b = inf kernels carry none of that code)
template <bool DP>
__device__ __forceinline__ double cand_lw(const DevState& S, int r, int e, int sz, double ll) {
  double lw = log_prior_existing(S.cl, sz) + ll;
  if (DP) lw += dp_loglik(S, r, e);
  return lw;
}

// finish one record's candidate list: bounds of its change of K and registration as a long record
__device__ __forceinline__ void finish_candidates(const DevState& S, int r, uint32_t off, int cnt, bool other) {
  S.cand_off[r] = off;
  S.cand_cnt[r] = cnt;
  S.dlo[r] = other ? -1 : 0; // while its outcome is unknown the record may lower K only if it can join another entity
  S.dhi[r] = 1;
  if (cnt == -1) S.wide_list[atomicAdd(&S.counters[C_NWIDE], 1)] = r;
  else if (cnt == -2) { atomicAdd(&S.counters[C_NFENCE], 1); atomicAdd(&S.counters[C_NOLIST], 1); }
  else if (cnt == -3) { S.unk_list[atomicAdd(&S.counters[C_NUNK], 1)] = r; atomicAdd(&S.counters[C_NOLIST], 1); }
  else if (cnt > S.long_min) S.actl_in[atomicAdd(&S.counters[C_NACTL_INIT], 1)] = r;
}

constexpr int WARP_BUCKET_MAX = 8192; // buckets up to this length are scanned by one warp
// Eight lanes per record (four records per warp): candidate lists are short (1.2-1.4 entries on
// average), so narrow groups keep four times as many independent gather chains in flight.
// Records served by a table with an unpinned key attribute probe one bucket per value of that
// attribute: they are listed (probe_list) and taken by a second launch with a whole warp per record
// (GL = 32, `list` = probe_list), so that their dozens of probes do not hold up a block of ordinary
// records.
// Small blocks (BT threads) for the first launch: a block lives as long as its slowest record.
template <int GL, int BT>
__global__ void __launch_bounds__(BT) k_cand(DevState S, const int* __restrict__ list, int n_rec) {
  __shared__ int s_id[BT / GL][CAND_SMEM];
  __shared__ double s_ll[BT / GL][CAND_SMEM];
  __shared__ unsigned s_stat[2], s_done; // candidates stored / bucket items visited by this block; warps finished
  if (threadIdx.x < 2) s_stat[threadIdx.x] = 0;
  if (threadIdx.x == 2) s_done = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, gl = lane & (GL - 1), gsh = lane & ~(GL - 1);
  constexpr unsigned GM = GL == 32 ? 0xFFFFFFFFu : (1u << GL) - 1u;
  const unsigned gmask = GM << gsh;
  const int gib = threadIdx.x / GL;
  const int gi = blockIdx.x * (BT / GL) + gib;
  const int r = gi < n_rec ? (list ? list[gi] : gi) : S.N;
  if (r < S.N) {
    int* ids = s_id[gib];
    double* lls = s_ll[gib];
    const int* x = S.rec_x + (size_t)r * S.RS;
    const uint32_t zm = S.rec_z[r];
    int tab = S.tab_remap[S.rec_tab[r]];
    if (!tab) tab = S.tab_remap[N_TABLES + S.rec_alt[r]];
    const int own = S.lambda[r];
    int cnt = 0;
    bool wide = tab == 0, spill = false, overflow = false;
    int lo = 0, hi = 0;
    bool deferred = false;
    Probe pr;
    pr.n_un = 0;
    if (!wide) {
      const DevTable& t = S.tables[tab];
      pr = make_probe(S, t.mask, x, zm); // n_un >= 1: the buckets of the unpinned attributes' possible values are probed
      if (pr.n_un == PROBE_NOT_SERVED) { wide = true; pr.n_un = 0; } // not served by this table after all: batched scan over all items
      else if (pr.n_un == 0) {
        const uint32_t b = bucket_of(S, t, x);
        if (gl == 0) S.rec_bkt[r] = b;
        lo = b ? t.ptr[b - 1] : 0;
        hi = t.ptr[b];
        // a bucket longer than bucket_cap is left to a whole warp (second launch), a very long one
        // to a whole block (k_cand_long)
        if (GL == 32) spill = hi - lo > WARP_BUCKET_MAX;
        else if (hi - lo > S.bucket_cap) {
          if (hi - lo <= WARP_BUCKET_MAX && S.bucket_cap > 0) {
            if (gl == 0) S.probe_list[atomicAdd(&S.counters[C_NPROBE], 1)] = r;
            deferred = true;
          } else spill = true;
        }
      }
    }
    // stage the group's hits of one step (lane: item / ok / ll), twins expanded
    auto stage = [&](int item, bool ok, double ll) {
      // an indexed slot whose potential entity is its twin stands for both items
      const bool tw = ok && item < S.N && item != r && S.twin[item];
      const unsigned m = (__ballot_sync(gmask, ok) >> gsh) & GM, m2 = (__ballot_sync(gmask, tw) >> gsh) & GM;
      const unsigned below = (1u << gl) - 1u;
      const int pos = cnt + __popc(m & below) + __popc(m2 & below);
      const int add = __popc(m) + __popc(m2);
      if (cnt + add > S.cand_cap) spill = true;
      else if (ok) {
        ids[pos] = item; lls[pos] = ll;
        if (tw) { ids[pos + 1] = S.N + item; lls[pos + 1] = ll; }
      }
      cnt += add;
    };
    if (!wide && !spill && !deferred && pr.n_un == 0) {
      const int* bucket = S.tables[tab].ids;
      for (int base = lo; base < hi && !spill; base += GL) {
        const int q = base + gl;
        int item = -1;
        bool ok = false;
        double ll = 0.0;
        if (q < hi) {
          item = bucket[q];
          if (item != S.N + r) {
            const int* y = S.ent_y + (size_t)item * S.ES;
            ok = item_compatible(S, x, zm, y);
            if (ok) {
              ll = item_loglik(S, x, zm, y);
              ok = ll > -CUDART_INF; // zero weight whatever its size: never chosen, adds 0 to every sum
            }
          }
        }
        stage(item, ok, ll);
      }
      if (gl == 0) atomicAdd(&s_stat[1], (unsigned)(hi - lo));
    } else if (!wide && pr.n_un > 0 && GL != 32) {
      if (pr.count > 512) spill = true; // thousands of probes: a whole block, one thread per probe
      else {
        if (gl == 0) S.probe_list[atomicAdd(&S.counters[C_NPROBE], 1)] = r; // a warp of the second launch takes it
        deferred = true;
      }
    } else if (!wide && pr.n_un > 0) {
      // lanes take the probed values round-robin and walk their own (short) buckets - or, for a
      // table in chain form, the chains of their cells (the node holds the item's packed tuple)
      const DevTable& t = S.tables[tab];
      const bool chain = t.chain != 0;
      int j_next = gl, q = 0, q_end = 0, visited = 0, c_item = -1;
      Pv v_cur = {0, 0, 0};
      // chain form: a node is compared as a packed word (fields Mc / values Vc of the pinned and the
      // probed attributes) and only unpacked when it agrees
      unsigned long long M0 = 0, V0 = 0, Mc = 0, Vc = 0;
      if (chain) record_fields(S, x, zm, M0, V0);
      for (;;) {
        while ((chain ? c_item < 0 : q >= q_end) && j_next < pr.count) {
          v_cur = pr.values(j_next);
          if (chain) {
            Mc = M0 | pk_field(S, pr.a1); Vc = V0 | ((unsigned long long)(uint32_t)v_cur.v1 << S.pk_shift[pr.a1]);
            if (pr.n_un >= 2) { Mc |= pk_field(S, pr.a2); Vc |= (unsigned long long)(uint32_t)v_cur.v2 << S.pk_shift[pr.a2]; }
            if (pr.n_un >= 3) { Mc |= pk_field(S, pr.a3); Vc |= (unsigned long long)(uint32_t)v_cur.v3 << S.pk_shift[pr.a3]; }
            c_item = t.head[chain_cell(t, Vc)];
          } else {
            const uint32_t b = pr.bucket(S, t, x, v_cur);
            q = b ? t.ptr[b - 1] : 0;
            q_end = t.ptr[b];
          }
          j_next += GL;
        }
        const bool active = chain ? c_item >= 0 : q < q_end;
        if (spill || !__any_sync(gmask, active)) break;
        int item = -1;
        bool ok = false;
        double ll = 0.0;
        if (active) {
          int yv[8];
          const int* y = yv;
          bool agrees = true;
          if (chain) {
            item = c_item;
            const ChainNode nd = load_node(&t.node[item]);
            c_item = nd.next;
            agrees = ((nd.P ^ Vc) & Mc) == 0;
            if (agrees) for (int a = 0; a < S.A; a++) yv[a] = unpack_attr(S, nd.P, a);
          } else {
            item = t.ids[q++];
            y = S.ent_y + (size_t)item * S.ES;
          }
          visited++;
          if (item != S.N + r && agrees) {
            // hashed buckets: an item counts in the probe of its own value only
            ok = pr.own_probe(y, v_cur) && item_compatible(S, x, zm, y);
            if (ok) {
              ll = item_loglik(S, x, zm, y);
              ok = ll > -CUDART_INF;
            }
          }
        }
        stage(item, ok, ll);
      }
      atomicAdd(&s_stat[1], (unsigned)visited);
    }
    if (deferred) {
    } else if (spill) {
      // long bucket or more candidates than a group can stage: a whole block takes the record
      if (gl == 0) S.long_todo[atomicAdd(&S.counters[C_NLONG_TODO], 1)] = r;
    } else {
      uint32_t off = 0;
      if (!wide && cnt > 0 && cnt <= S.long_cap) {
        // sort the staged candidates by id (bitonic, eight lanes)
        int n2 = 1;
        while (n2 < cnt) n2 <<= 1;
        for (int i = cnt + gl; i < n2; i += GL) { ids[i] = 0x7FFFFFFF; lls[i] = 0.0; }
        __syncwarp(gmask);
        for (int k = 2; k <= n2; k <<= 1)
          for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = gl; i < n2; i += GL) {
              const int l = i ^ j;
              if (l > i) {
                const bool up = (i & k) == 0;
                const int a0 = ids[i], a1 = ids[l];
                if ((a0 > a1) == up) {
                  ids[i] = a1; ids[l] = a0;
                  const double t0 = lls[i]; lls[i] = lls[l]; lls[l] = t0;
                }
              }
            }
            __syncwarp(gmask);
          }
        // lists of up to CAND_INLINE entries live in the record's own slot of the pool; only longer
        // ones take space from the shared cursor (one global atomic per record stalled this kernel)
        unsigned long long o = (unsigned long long)CAND_INLINE * (unsigned long long)r;
        if (cnt > CAND_INLINE) {
          if (gl == 0) o = atomicAdd(S.pool_cursor, (unsigned long long)cnt);
          o = __shfl_sync(gmask, o, gsh);
        }
        if (o + (unsigned long long)cnt > S.pool_cap) overflow = true;
        else {
          off = (uint32_t)o;
          for (int i = gl; i < cnt; i += GL) {
            S.cand_id[o + i] = ids[i]; S.cand_ll[o + i] = lls[i];
            if (ids[i] != own) S.referenced[ids[i]] = 1;
          }
        }
      }
      if (gl == 0) {
        // no usable bucket: the host counts this record's possible links to classify it (-3);
        // candidate pool exhausted: serial turn inside the rounds (-2)
        const int code = wide ? -3 : (cnt > S.long_cap ? -1 : (overflow ? -2 : cnt));
        bool other = code < 0;
        if (code >= 0) for (int i = 0; i < cnt; i++) other |= ids[i] != own;
        finish_candidates(S, r, off, code, other);
        if (code >= 0) atomicAdd(&s_stat[0], (unsigned)cnt);
      }
    }
  }
  // the last warp to finish adds the block's statistics to the global counters (no barrier: the
  // other warps are gone by then)
  __syncwarp();
  if (lane == 0) {
    __threadfence_block();
    if (atomicAdd(&s_done, 1u) == BT / 32 - 1) {
      for (int k = 0; k < 2; k++) {
        const unsigned v = atomicAdd(&s_stat[k], 0u);
        if (v) atomicAdd((unsigned long long*)&S.counters64[k], (unsigned long long)v);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// K1b''  chain form of the large blocking tables (key spaces about as large as the item set):
// every indexed item is pushed on the chain of its cell with ONE atomicExch per table; its node
// (packed tuple, next, twin flag) is written at the item's own index - coalesced.  Replaces the
// partition / count / scan / fill passes of the CSR form for these tables.
// ------------------------------------------------------------------------------------------
// pass 1: packed tuple of every indexed item (~0: not indexed), bit 63 of the word below it: twin flag
__global__ void __launch_bounds__(TPB) k_chain_pack(DevState S, unsigned long long* __restrict__ P, uint8_t* __restrict__ fl) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * S.N) return;
  const bool idx = item_indexed(S, i);
  fl[i] = (uint8_t)(idx ? ((i < S.N && S.twin[i]) ? 3 : 1) : 0); // bit 0: indexed, bit 1: twin
  if (idx) {
    const int4* yp = reinterpret_cast<const int4*>(S.ent_y + (size_t)i * 8); // ES == 8 on this path
    const int4 y0 = yp[0], y1 = yp[1];
    const int y[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
    P[i] = pack_tuple(S, y);
  }
}
// pass 2, once per table (its head array stays in the L2 during the pass)
__global__ void __launch_bounds__(TPB) k_chain_build(DevState S, int tab, const unsigned long long* __restrict__ P,
                                                     const uint8_t* __restrict__ fl) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * S.N) return;
  const uint8_t f = fl[i];
  if (!(f & 1)) return;
  const DevTable& t = S.tables[tab];
  const unsigned long long p = P[i];
  const int prev = atomicExch(&t.head[chain_cell(t, p)], i);
  *reinterpret_cast<int4*>(&t.node[i]) = make_int4((int)(uint32_t)p, (int)(uint32_t)(p >> 32), prev, (int)(f >> 1));
}

// ------------------------------------------------------------------------------------------
// K1c'  candidate generation, ONE THREAD per record, for records whose table is in chain form and
// whose key attributes are all pinned (19 of 20 records at 10 M): the thread walks the chain of its
// cell, compares packed tuples (the node it has to read anyway holds the item's attributes: no
// gather of the entity table, no bucket bounds), keeps up to FAST_CAND candidates sorted by id in
// local memory and stores them.  A list that is just the record's own entity with log-likelihood 0
// (7 of 8 records) is not stored at all (CAND_OWN_ONLY).  Everything else is handed on:
//   * records that probe several cells (an unpinned key attribute)    -> probe_list (warp / record)
//   * records of tables in CSR form, records without a table, chains longer than hop_cap or more
//     than FAST_CAND candidates                                         -> slow_list (k_cand<8>)
// Same candidate sets, same order, same log-likelihoods as k_cand.
// ------------------------------------------------------------------------------------------
// Two passes: the first over all records does the single-cell ones and lists the records that probe
// several cells (`multi_list`); the second runs over that list, so that a warp is not held up by
// the one lane in thirty that makes a hundred probes.  No block barrier: warps retire on their own.
constexpr int FAST_TPB = 128;
__global__ void __launch_bounds__(FAST_TPB) k_cand_fast(DevState S, const int* __restrict__ list, int n_rec, int* __restrict__ multi_list) {
  const int gi = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = gi < n_rec ? (list ? list[gi] : gi) : S.N;
  unsigned n_cand = 0, n_hops = 0;
  if (r < S.N) {
    int x[8];
    {
      const int4* xp = reinterpret_cast<const int4*>(S.rec_x + (size_t)r * 8); // RS == 8 on this path
      const int4 x0 = xp[0], x1 = xp[1];
      x[0] = x0.x; x[1] = x0.y; x[2] = x0.z; x[3] = x0.w; x[4] = x1.x; x[5] = x1.y; x[6] = x1.z; x[7] = x1.w;
    }
    const uint32_t zm = S.rec_z[r];
    int tab = S.tab_remap[S.rec_tab[r]];
    if (!tab) tab = S.tab_remap[N_TABLES + S.rec_alt[r]];
    bool slow = tab == 0;
    if (!slow) {
      const DevTable& t = S.tables[tab];
      unsigned long long M, V;
      record_fields(S, x, zm, M, V);
      Probe pr;
      pr.n_un = 0; pr.count = 1; pr.a1 = pr.a2 = pr.a3 = -1;
      if (!t.chain) slow = true;
      else if ((M & t.kmask) != t.kmask) {
        // an unpinned key attribute: the cells of its possible values are probed - by this thread if
        // they are few (a distorted name: its neighbourhood row), else by a warp or a block
        pr = make_probe(S, t.mask, x, zm);
        if (pr.n_un == PROBE_NOT_SERVED) slow = true;  // not served by this table after all (k_cand marks it)
        else if (pr.count > 512) { S.long_todo[atomicAdd(&S.counters[C_NLONG_TODO], 1)] = r; pr.n_un = -1; } // a block, one thread per probe
        else if (pr.count > FAST_PROBES) { S.probe_list[atomicAdd(&S.counters[C_NPROBE], 1)] = r; pr.n_un = -1; }
        else if (multi_list) { multi_list[atomicAdd(&S.counters[C_NMULTI], 1)] = r; pr.n_un = -1; } // second pass
      }
      if (!slow && pr.n_un >= 0) {
        const int own = S.lambda[r];
        const uint32_t zs = zm & ~S.dp_mask;
        int ids[FAST_CAND];
        double lls[FAST_CAND];
        int cnt = 0, hops = 0;
        bool over = false;
        auto insert = [&](int id, double ll) {
          if (cnt == FAST_CAND) { over = true; return; }
          int j = cnt++;
          while (j > 0 && ids[j - 1] > id) { ids[j] = ids[j - 1]; lls[j] = lls[j - 1]; j--; }
          ids[j] = id; lls[j] = ll;
        };
        for (int pj = 0; pj < pr.count && !over; pj++) {
          // fields / values the items of this probe must match: the pinned attributes, and the probed
          // values of the unpinned key attributes (an item sits in the cell of its own key only)
          unsigned long long Mp = M, Vp = V;
          if (pr.n_un > 0) {
            const Pv v = pr.values(pj);
            Mp |= pk_field(S, pr.a1); Vp |= (unsigned long long)(uint32_t)v.v1 << S.pk_shift[pr.a1];
            if (pr.n_un >= 2) { Mp |= pk_field(S, pr.a2); Vp |= (unsigned long long)(uint32_t)v.v2 << S.pk_shift[pr.a2]; }
            if (pr.n_un >= 3) { Mp |= pk_field(S, pr.a3); Vp |= (unsigned long long)(uint32_t)v.v3 << S.pk_shift[pr.a3]; }
          }
          int item = t.head[chain_cell(t, Vp)];
          while (item >= 0 && !over) {
            if (++hops > S.hop_cap) { over = true; break; }
            const ChainNode nd = load_node(&t.node[item]);
            if (item != S.N + r && ((nd.P ^ Vp) & Mp) == 0) {
              double ll = 0.0; // item_loglik on the packed tuple
              for (int a = 0; a < S.A; a++)
                if (x[a] >= 0 && ((zs >> a) & 1u)) ll += log_distortion_prob(S.attrs[a], unpack_attr(S, nd.P, a), x[a]);
              if (ll > -CUDART_INF) {
                insert(item, ll);
                if ((nd.flags & 1u) && item != r) insert(S.N + item, ll); // the twin potential entity of that slot
              }
            }
            item = nd.next;
          }
        }
        n_hops = (unsigned)hops;
        if (over) { if (pr.n_un == 0) S.tab_over[tab] = 1; slow = true; } // (probing records need no CSR form: k_cand<32> walks chains)
        else {
          uint32_t off = 0;
          int code = cnt;
          bool other = false;
          for (int j = 0; j < cnt; j++) other |= ids[j] != own;
          if (cnt == 1 && !other && lls[0] == 0.0 && S.long_min >= 1) off = CAND_OWN_ONLY;
          else if (cnt > 0) {
            unsigned long long o = (unsigned long long)CAND_INLINE * (unsigned long long)r;
            if (cnt > CAND_INLINE) o = atomicAdd(S.pool_cursor, (unsigned long long)cnt);
            if (o + (unsigned long long)cnt > S.pool_cap) { code = -2; other = true; }
            else {
              off = (uint32_t)o;
              for (int j = 0; j < cnt; j++) {
                S.cand_id[o + j] = ids[j]; S.cand_ll[o + j] = lls[j];
                if (ids[j] != own) S.referenced[ids[j]] = 1;
              }
            }
          }
          finish_candidates(S, r, off, code, other);
          if (code >= 0) n_cand = (unsigned)cnt;
        }
      }
    }
    if (slow) S.slow_list[atomicAdd(&S.counters[C_NSLOW], 1)] = r;
  }
  // statistics: per warp, spread over 32 slots (one hot address would serialise 300 k warps)
  n_cand = (unsigned)warp_sum((int)n_cand); n_hops = (unsigned)warp_sum((int)n_hops);
  if ((threadIdx.x & 31) == 0) {
    long long* slot = S.counters64 + 2 + 2 * ((gi >> 5) & 31);
    if (n_cand) atomicAdd((unsigned long long*)&slot[0], (unsigned long long)n_cand);
    if (n_hops) atomicAdd((unsigned long long*)&slot[1], (unsigned long long)n_hops);
  }
}

__host__ __device__ __forceinline__ int pow2_ceil(int n) { int p = 1; while (p < n) p <<= 1; return p; }

// Block-wide tail of candidate staging: sort `cnt` staged (id, ll) pairs by id, append them to the
// pool and register the record (cnt > cap: more than wide_threshold possible links).
__device__ void stage_sort_store(const DevState& S, int r, int own, int* ids, double* lls, int cnt, int cap,
                                 unsigned long long* s_off) {
  bool wide = cnt > cap;
  if (!wide) {
    int n2 = 1;
    while (n2 < cnt) n2 <<= 1;
    for (int i = cnt + threadIdx.x; i < n2; i += blockDim.x) { ids[i] = 0x7FFFFFFF; lls[i] = 0.0; }
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = threadIdx.x; i < n2; i += blockDim.x) {
          const int l = i ^ j;
          if (l > i) {
            const bool up = (i & k) == 0;
            const int a0 = ids[i], a1 = ids[l];
            if ((a0 > a1) == up) {
              ids[i] = a1; ids[l] = a0;
              const double t0 = lls[i]; lls[i] = lls[l]; lls[l] = t0;
            }
          }
        }
        __syncthreads();
      }
    if (threadIdx.x == 0) *s_off = atomicAdd(S.pool_cursor, (unsigned long long)cnt);
    __syncthreads();
    const unsigned long long o = *s_off;
    const bool overflow = o + (unsigned long long)cnt > S.pool_cap;
    if (!overflow) for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
      S.cand_id[o + i] = ids[i]; S.cand_ll[o + i] = lls[i];
      if (ids[i] != own) S.referenced[ids[i]] = 1;
    }
    if (threadIdx.x == 0) {
      const bool other = overflow || cnt > 1 || (cnt == 1 && ids[0] != own);
      finish_candidates(S, r, overflow ? 0u : (uint32_t)o, overflow ? -2 : cnt, other);
      if (!overflow) atomicAdd((unsigned long long*)&S.counters64[0], (unsigned long long)cnt);
    }
  } else if (threadIdx.x == 0) {
    finish_candidates(S, r, 0u, -1, true); // more than wide_threshold possible links: its turn comes first
  }
  __syncthreads();
}

// One BLOCK per record whose bucket is long or whose candidates overflowed a group's staging area:
// same scan, staged in dynamic shared memory (wide_threshold entries), sorted block-wide.
// Buckets longer than huge_min are left to the grid-wide k_huge_scan.
__global__ void __launch_bounds__(TPB) k_cand_long(DevState S, int n_todo, int cap) {
  extern __shared__ unsigned char s_raw[];
  double* lls = reinterpret_cast<double*>(s_raw);
  int* ids = reinterpret_cast<int*>(lls + pow2_ceil(cap)); // the sort pads the list to a power of two
  __shared__ int s_cnt;
  __shared__ unsigned long long s_off;
  __shared__ int s_next;
  for (;;) { // records are handed out one at a time: their costs differ by orders of magnitude
    if (threadIdx.x == 0) { s_next = atomicAdd(&S.counters[C_LONG_NEXT], 1); s_cnt = 0; }
    __syncthreads();
    const int w = s_next;
    if (w >= n_todo) break;
    const int r = S.long_todo[w];
    const int* x = S.rec_x + (size_t)r * S.RS;
    const uint32_t zm = S.rec_z[r];
    int tab = S.tab_remap[S.rec_tab[r]];
    if (!tab) tab = S.tab_remap[N_TABLES + S.rec_alt[r]];
    const DevTable& t = S.tables[tab];
    const int own = S.lambda[r];
    const Probe pr = make_probe(S, t.mask, x, zm); // (n_un = 3 never gets here: k_cand sends those to the batched scan)
    if (pr.n_un == 0) {
      const uint32_t b = S.rec_bkt[r];
      const int lo = b ? t.ptr[b - 1] : 0, hi = t.ptr[b];
      if (hi - lo > S.huge_min) {
        if (threadIdx.x == 0) S.huge_list[atomicAdd(&S.counters[C_NHUGE], 1)] = r;
        __syncthreads();
        continue;
      }
    }
    auto visit = [&](int item, const Pv& v, const int* y) {
      if (item == S.N + r) return;
      if (!y) y = S.ent_y + (size_t)item * S.ES;
      if (pr.n_un > 0 && !pr.own_probe(y, v)) return;
      if (!item_compatible(S, x, zm, y)) return;
      const double ll = item_loglik(S, x, zm, y);
      if (!(ll > -CUDART_INF)) return;
      const bool tw = item < S.N && item != r && S.twin[item];
      const int pos = atomicAdd(&s_cnt, tw ? 2 : 1);
      if (pos < cap) { ids[pos] = item; lls[pos] = ll; }
      if (tw && pos + 1 < cap) { ids[pos + 1] = S.N + item; lls[pos + 1] = ll; }
    };
    if (pr.n_un > 0 && t.chain) {
      // chain form, one thread per probe, four probes of a thread in flight: most cells of a probed
      // value combination are empty, so the time goes into the reads of the cell heads
      // (half of the probed cells hold an item of some other key: the first node of every chain is read
      // with the others in flight too, and a node is compared as a packed word - fields M / values V of
      // the pinned and the probed attributes - before anything is unpacked)
      unsigned long long M, V;
      record_fields(S, x, zm, M, V);
      for (int j0 = threadIdx.x; j0 < pr.count; j0 += 4 * blockDim.x) {
        Pv v[4];
        int head[4];
        unsigned long long Mp[4], Vp[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int j = j0 + u * (int)blockDim.x;
          head[u] = -1;
          if (j < pr.count) {
            v[u] = pr.values(j);
            Mp[u] = M | pk_field(S, pr.a1); Vp[u] = V | ((unsigned long long)(uint32_t)v[u].v1 << S.pk_shift[pr.a1]);
            if (pr.n_un >= 2) { Mp[u] |= pk_field(S, pr.a2); Vp[u] |= (unsigned long long)(uint32_t)v[u].v2 << S.pk_shift[pr.a2]; }
            if (pr.n_un >= 3) { Mp[u] |= pk_field(S, pr.a3); Vp[u] |= (unsigned long long)(uint32_t)v[u].v3 << S.pk_shift[pr.a3]; }
            head[u] = t.head[chain_cell(t, Vp[u])];
          }
        }
        ChainNode nd[4];
#pragma unroll
        for (int u = 0; u < 4; u++) if (head[u] >= 0) nd[u] = load_node(&t.node[head[u]]);
#pragma unroll
        for (int u = 0; u < 4; u++)
          for (int item = head[u]; item >= 0 && s_cnt <= cap;) {
            if (item != head[u]) nd[u] = load_node(&t.node[item]);
            if (((nd[u].P ^ Vp[u]) & Mp[u]) == 0) {
              int yv[8];
              for (int a = 0; a < S.A; a++) yv[a] = unpack_attr(S, nd[u].P, a);
              visit(item, v[u], yv);
            }
            item = nd[u].next;
          }
      }
    } else if (pr.n_un > 0) {
      // one thread per probe: the buckets are short, the probes independent
      for (int j = threadIdx.x; j < pr.count; j += blockDim.x) {
        const Pv v = pr.values(j);
        const uint32_t b = pr.bucket(S, t, x, v);
        const int lo = b ? t.ptr[b - 1] : 0, hi = t.ptr[b];
        for (int q = lo; q < hi && s_cnt <= cap; q++) visit(t.ids[q], v, nullptr);
      }
    } else if (S.pk) {
      // the bucket against the packed tuples: four items per thread in flight, one masked compare per
      // item (the attribute row of an item is only unpacked for the few that agree)
      const uint32_t b = S.rec_bkt[r];
      const int lo = b ? t.ptr[b - 1] : 0, hi = t.ptr[b];
      unsigned long long M, V;
      record_fields(S, x, zm, M, V);
      for (int q0 = lo; q0 < hi; q0 += 4 * blockDim.x) {
        if (s_cnt > cap) break; // already too many: wide (benign race: any exit is a valid exit)
        int it[4];
        unsigned long long P[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { const int q = q0 + u * (int)blockDim.x + (int)threadIdx.x; it[u] = q < hi ? t.ids[q] : -1; }
#pragma unroll
        for (int u = 0; u < 4; u++) P[u] = it[u] >= 0 ? S.pk[it[u]] : ~V;
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (it[u] >= 0 && ((P[u] ^ V) & M) == 0) {
            int yv[8];
            for (int a = 0; a < S.A; a++) yv[a] = unpack_attr(S, P[u], a);
            visit(it[u], Pv{0, 0, 0}, yv);
          }
      }
    } else {
      const uint32_t b = S.rec_bkt[r];
      const int lo = b ? t.ptr[b - 1] : 0, hi = t.ptr[b];
      for (int q0 = lo; q0 < hi; q0 += blockDim.x) {
        if (s_cnt > cap) break; // already too many: wide (benign race: any exit is a valid exit)
        const int q = q0 + threadIdx.x;
        if (q < hi) visit(t.ids[q], Pv{0, 0, 0}, nullptr);
      }
    }
    __syncthreads();
    stage_sort_store(S, r, own, ids, lls, s_cnt, cap, &s_off);
  }
}

// Huge buckets (longer than huge_min) are scanned by many blocks at once, HUGE_BATCH records per
// launch: k_huge_plan cuts every record's bucket into tiles, k_huge_scan takes one tile per block
// and appends the possible links to the record's scratch list (unordered), k_huge_finish sorts and
// stores each list with one block per record (or finds the record wide).
constexpr int HUGE_TILE = 8192;
__device__ __forceinline__ void huge_bucket(const DevState& S, int r, const DevTable*& t, int& lo, int& hi) {
  int tab = S.tab_remap[S.rec_tab[r]];
  if (!tab) tab = S.tab_remap[N_TABLES + S.rec_alt[r]];
  t = &S.tables[tab];
  const uint32_t b = S.rec_bkt[r];
  lo = b ? t->ptr[b - 1] : 0;
  hi = t->ptr[b];
}
// tile_start[i] = first tile of record i of the batch (exclusive scan, one block), [n] = total
__global__ void __launch_bounds__(1024) k_huge_plan(DevState S, const int* __restrict__ recs, int n, int* __restrict__ tile_start,
                                                    int* __restrict__ cnt) {
  __shared__ int s_w[32];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    int tiles = 0;
    if (i < n) {
      const DevTable* t; int lo, hi;
      huge_bucket(S, recs[i], t, lo, hi);
      tiles = (hi - lo + HUGE_TILE - 1) / HUGE_TILE;
      cnt[i] = 0;
    }
    int x = tiles;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULL, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_w[w] = x;
    __syncthreads();
    if (w == 0) {
      int z = s_w[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULL, z, o); if (lane >= o) z += y; }
      s_w[lane] = z;
    }
    __syncthreads();
    const int excl = s_carry + x - tiles + (w ? s_w[w - 1] : 0);
    if (i < n) tile_start[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = excl + tiles;
    __syncthreads();
  }
  if (threadIdx.x == 0) tile_start[n] = s_carry;
}
__global__ void __launch_bounds__(TPB) k_huge_scan(DevState S, const int* __restrict__ recs, int n, const int* __restrict__ tile_start,
                                                   int* __restrict__ cnt, int cap) {
  // which record does this tile belong to
  int a0 = 0, b0 = n; // last i with tile_start[i] <= blockIdx.x
  while (b0 - a0 > 1) { const int mid = (a0 + b0) >> 1; if (tile_start[mid] <= (int)blockIdx.x) a0 = mid; else b0 = mid; }
  const int i = a0, r = recs[i];
  const DevTable* t; int lo, hi;
  huge_bucket(S, r, t, lo, hi);
  const int q_lo = lo + ((int)blockIdx.x - tile_start[i]) * HUGE_TILE, q_hi = min(hi, q_lo + HUGE_TILE);
  const int* x = S.rec_x + (size_t)r * S.RS;
  const uint32_t zm = S.rec_z[r];
  int* ids = S.huge_id + (size_t)i * (cap + 2);
  double* lls = S.huge_ll + (size_t)i * (cap + 2);
  for (int q = q_lo + threadIdx.x; q < q_hi; q += TPB) {
    if (cnt[i] > cap) return; // already wide (benign race: any exit is a valid exit)
    const int item = t->ids[q];
    if (item == S.N + r) continue;
    const int* y = S.ent_y + (size_t)item * S.ES;
    if (!item_compatible(S, x, zm, y)) continue;
    const double ll = item_loglik(S, x, zm, y);
    if (!(ll > -CUDART_INF)) continue;
    const bool tw = item < S.N && item != r && S.twin[item];
    const int pos = atomicAdd(&cnt[i], tw ? 2 : 1);
    if (pos < cap) { ids[pos] = item; lls[pos] = ll; }
    if (tw && pos + 1 < cap) { ids[pos + 1] = S.N + item; lls[pos + 1] = ll; }
  }
}
__global__ void __launch_bounds__(TPB) k_huge_finish(DevState S, const int* __restrict__ recs, const int* __restrict__ cnt, int cap) {
  extern __shared__ unsigned char s_raw[];
  double* lls = reinterpret_cast<double*>(s_raw);
  int* ids = reinterpret_cast<int*>(lls + pow2_ceil(cap)); // the sort pads the list to a power of two
  __shared__ unsigned long long s_off;
  const int i = blockIdx.x, r = recs[i];
  const int n = cnt[i];
  const int* gid = S.huge_id + (size_t)i * (cap + 2);
  const double* gll = S.huge_ll + (size_t)i * (cap + 2);
  if (n <= cap)
    for (int k = threadIdx.x; k < n; k += blockDim.x) { ids[k] = gid[k]; lls[k] = gll[k]; }
  __syncthreads();
  stage_sort_store(S, r, S.lambda[r], ids, lls, n, cap, &s_off);
}

// ------------------------------------------------------------------------------------------
// K1d  dependency rounds.  The links update must equal the sequential scan of
// Links::update (src/links.cpp:485-488).  Every unfinalised record reserves (atomicMin of its id)
// all the items its conditional reads or may write: its entity, its potential entity and its
// candidates.  A record whose reservations all carry its own id has no earlier unfinalised
// record that could change what it reads, so its conditional is already the one the sequential
// scan would see.  A record whose list did not fit the candidate pool (cand_cnt = -2) acts as a
// fence: it takes its turn serially when every earlier record is final.
// Records with short lists take one thread each; records with more than long_min candidates take
// a warp each (k_reserve_w / k_commit_w) - same arithmetic in the same order.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long resv_key(uint32_t round, int r) {
  return ((unsigned long long)(0xFFFFFFFFu - round) << 32) | (unsigned long long)(uint32_t)r;
}

// reservations of one record with a short list (thread) ...
__device__ __forceinline__ void reserve_short(const DevState& S, int r, int cnt, unsigned long long key) {
  atomicMin(&S.owner[S.lambda[r]], key);
  atomicMin(&S.owner[S.N + r], key);
  if (S.cand_off[r] == CAND_OWN_ONLY) return; // its only candidate is its own entity
  const int* ids = S.cand_id + S.cand_off[r];
  for (int j = 0; j < cnt; j++) atomicMin(&S.owner[ids[j]], key);
}
// ... and of one record with a long list (all threads of the block)
__device__ __forceinline__ void reserve_long(const DevState& S, int r, unsigned long long key) {
  const int cnt = S.cand_cnt[r];
  const int* ids = S.cand_id + S.cand_off[r];
  if (threadIdx.x == 0) {
    atomicMin(&S.owner[S.lambda[r]], key);
    atomicMin(&S.owner[S.N + r], key);
  }
  for (int j = threadIdx.x; j < cnt; j += blockDim.x) atomicMin(&S.owner[ids[j]], key);
}

__global__ void __launch_bounds__(TPB) k_reserve(DevState S, const int* __restrict__ act, int n_act, uint32_t round) {
  __shared__ int s_min;
  if (threadIdx.x == 0) s_min = 0x7FFFFFFF;
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int mine = 0x7FFFFFFF;
  if (i < n_act) {
    const int r = act ? act[i] : i;
    const int cnt = S.fin[r] == 1 ? 0x7FFFFFFF : S.cand_cnt[r]; // (final already: wide records, round 0)
    if (cnt <= S.long_min) { // long records: k_reserve_w
      mine = r;
      const unsigned long long key = resv_key(round, r);
      if (cnt < 0) atomicMin(S.fence, key);
      else reserve_short(S, r, cnt, key);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mine = min(mine, __shfl_xor_sync(FULL, mine, o));
  if ((threadIdx.x & 31) == 0) atomicMin(&s_min, mine);
  __syncthreads();
  if (threadIdx.x == 0 && s_min != 0x7FFFFFFF) atomicMin(&S.counters[C_MIN_ACTIVE], s_min);
}

__global__ void __launch_bounds__(TPB) k_reserve_w(DevState S, const int* __restrict__ act, int n_act, uint32_t round) {
  const int r = act[blockIdx.x]; // one block per record with a long list
  reserve_long(S, r, resv_key(round, r));
  if (threadIdx.x == 0) atomicMin(&S.counters[C_MIN_ACTIVE], r);
}

// move record r from entity `own` to `target`, keeping member lists sorted (the role of
// Links::set_link, src/links.cpp:6-39, and Entities::add / remove, src/entities.cpp:18-69)
__device__ __forceinline__ void apply_move(const DevState& S, int r, int own, int target) {
  if (target != own) {
    int h = S.ent_head[own];
    if (h == r) S.ent_head[own] = S.rec_next[r];
    else {
      while (S.rec_next[h] != r) h = S.rec_next[h];
      S.rec_next[h] = S.rec_next[r];
    }
    S.ent_size[own] -= 1;
    h = S.ent_head[target];
    if (h < 0 || r < h) { S.rec_next[r] = h; S.ent_head[target] = r; }
    else {
      int nx = S.rec_next[h];
      while (nx >= 0 && nx < r) { h = nx; nx = S.rec_next[h]; }
      S.rec_next[r] = nx;
      S.rec_next[h] = r;
    }
    S.ent_size[target] += 1;
    S.lambda[r] = target;
  }
}

// decision "found a new entity" given the number of entities k seen by prior_weight_new
__device__ __forceinline__ bool decide_new(const DevState& S, int r, int k, double m1, double srel, double u1,
                                           bool& invalid) {
  const double lwn = log(prior_new(S.cl, k)) + S.Lnew[r];
  const double m2 = m1 > lwn ? m1 : lwn;
  invalid = !(m2 > -CUDART_INF);
  const double wa = exp(lwn - m2);
  const double wb = m1 > -CUDART_INF ? srel * exp(m1 - m2) : 0.0;
  return u1 * (wa + wb) < wa;
}

// Given the existing-candidate summary (m1, srel) decide "new" for the interval of K that the
// unfinalised earlier records leave open.  Returns 0: existing, 1: new, 2: cannot decide yet.
__device__ __forceinline__ int decide_with_bounds(const DevState& S, int r, bool single, double m1, double srel,
                                                  double u1) {
  // with r taken out there are between 0 and N-1 entities, whatever the earlier records do
  const int klo = max(0, S.K0 + (S.use_kbounds ? S.plo[r] : 0) - (single ? 1 : 0));
  const int khi = min(S.N - 1, S.K0 + (S.use_kbounds ? S.phi_[r] : 0) - (single ? 1 : 0));
  bool inv_lo = false, inv_hi = false;
  const bool new_lo = decide_new(S, r, klo, m1, srel, u1, inv_lo);
  const bool new_hi = klo == khi ? new_lo : decide_new(S, r, khi, m1, srel, u1, inv_hi);
  // prior_weight_new is monotone in K (alpha + d K; kappa |m - K|) unless the interval straddles
  // the kink of |m - K|; only then the end points do not decide
  const bool kink = S.cl.kind != EXB200_CLUST_PITMAN_YOR && (double)klo < S.cl.m && S.cl.m < (double)khi;
  if (new_lo != new_hi || kink) return 2;
  if (inv_lo || inv_hi) dev_fail(S, EXB200_ERR_WEIGHTS, r); // "no valid links for record" (links.cpp:415-419)
  return new_lo ? 1 : 0;
}

// A record without any live candidate founds an entity whatever the uniform says (u1 wa < wa with
// wb = 0), provided its new-entity weight is valid at both ends of the K interval: seven records in
// eight, for which the logarithms, exponentials and the Philox call of the general decision are
// skipped.  Returns true if that shortcut applies (the general path handles everything else,
// including the errors).
__device__ __forceinline__ bool founds_for_sure(const DevState& S, int r, bool single) {
  const int klo = max(0, S.K0 + (S.use_kbounds ? S.plo[r] : 0) - (single ? 1 : 0));
  const int khi = min(S.N - 1, S.K0 + (S.use_kbounds ? S.phi_[r] : 0) - (single ? 1 : 0));
  const bool kink = S.cl.kind != EXB200_CLUST_PITMAN_YOR && (double)klo < S.cl.m && S.cl.m < (double)khi;
  const double Ln = S.Lnew[r], p_lo = prior_new(S.cl, klo), p_hi = prior_new(S.cl, khi);
  return !kink && Ln > -CUDART_INF && Ln < CUDART_INF && p_lo > 0.0 && p_lo < CUDART_INF && p_hi > 0.0 && p_hi < CUDART_INF;
}

__device__ __forceinline__ void finalize_record(const DevState& S, int r, int own, int target, bool single) {
  const bool is_new = target == S.N + r;
  const int dk = (is_new ? 1 : 0) - (single ? 1 : 0);
  if (!(target == own || (single && is_new))) atomicAdd(&S.counters[C_CHANGED], 1);
  apply_move(S, r, own, target);
  S.fin[r] = 1;
  S.dlo[r] = (int8_t)dk;
  S.dhi[r] = (int8_t)dk;
  if (dk) atomicAdd(&S.counters[C_DK_TOTAL], dk);
}

__device__ __forceinline__ int live_size(const DevState& S, int e, int own, bool single) {
  const int sz = S.ent_size[e];
  return e == own ? (single ? 0 : sz - 1) : sz; // links.cpp:359, :375
}

constexpr int SHORT_MAX = 32; // log-weights a thread keeps for one record (the default long_min)
// Turn of one record with a short list, if it owns everything it reserved (thread per record).
// Returns true if the record stays unfinalised.
template <bool DP>
__device__ __forceinline__ bool commit_short(const DevState& S, int r, int cnt, unsigned long long key) {
  const int own = S.lambda[r];
  const bool own_only = S.cand_off[r] == CAND_OWN_ONLY;
  const int* ids = own_only ? &own : S.cand_id + S.cand_off[r];
  bool mine = S.owner[own] == key && S.owner[S.N + r] == key;
  if (!own_only) for (int j = 0; j < cnt && mine; j++) mine = S.owner[ids[j]] == key;
  if (!mine) return true;
  const double zero_ll = 0.0;
  const double* lls = own_only ? &zero_ll : S.cand_ll + S.cand_off[r];
  const bool single = S.ent_size[own] == 1;
  // weights of the live candidates (links.cpp:372-388).  The sizes are gathered once (independent
  // loads, four in flight) and the log-weights kept: the three passes of the sequential conditional
  // (maximum, sum, inverse CDF) then run on the same values without touching memory again.
  double lw[SHORT_MAX];
  double m1 = -CUDART_INF;
  if (cnt <= SHORT_MAX) {
    for (int j0 = 0; j0 < cnt; j0 += 4) {
      int sz[4];
#pragma unroll
      for (int u = 0; u < 4; u++) sz[u] = j0 + u < cnt ? live_size(S, ids[j0 + u], own, single) : 0;
#pragma unroll
      for (int u = 0; u < 4; u++)
        if (j0 + u < cnt) {
          const double v = sz[u] > 0 ? cand_lw<DP>(S, r, ids[j0 + u], sz[u], lls[j0 + u]) : -CUDART_INF;
          lw[j0 + u] = v;
          m1 = v > m1 ? v : m1;
        }
    }
  } else { // (long_min raised above SHORT_MAX by an option: nothing is kept)
    for (int j = 0; j < cnt; j++) {
      const int sz = live_size(S, ids[j], own, single);
      if (sz > 0) { const double v = cand_lw<DP>(S, r, ids[j], sz, lls[j]); m1 = v > m1 ? v : m1; }
    }
  }
  auto lw_of = [&](int j) -> double {
    if (cnt <= SHORT_MAX) return lw[j];
    const int sz = live_size(S, ids[j], own, single);
    return sz > 0 ? cand_lw<DP>(S, r, ids[j], sz, lls[j]) : -CUDART_INF;
  };
  double srel = 0.0;
  if (m1 > -CUDART_INF)
    for (int j = 0; j < cnt; j++) {
      const double v = lw_of(j);
      if (v != -CUDART_INF) srel += exp(v - m1); // (a candidate that is not live, or of zero weight, adds nothing)
    }
  if (!(m1 > -CUDART_INF) && founds_for_sure(S, r, single)) {
    finalize_record(S, r, own, S.N + r, single);
    return false;
  }
  const U2 u = uniforms(S.seed, r, S.iter, PH_LINK_DECIDE, 0, 0);
  const int dec = decide_with_bounds(S, r, single, m1, srel, u.a);
  if (dec == 2) {
    // entity-wise ready, only K is still uncertain: tighten this record's bounds
    S.dlo[r] = single ? -1 : 0;
    S.dhi[r] = single ? 0 : 1;
    return true;
  }
  int target;
  if (dec == 1) target = S.N + r;
  else {
    const double tt = u.b * srel;
    double c = 0.0;
    int pick = -1, lastpos = -1;
    for (int j = 0; j < cnt; j++) {
      const double v = lw_of(j);
      if (v == -CUDART_INF) continue; // not live
      const double w = exp(v - m1);
      if (w > 0.0) lastpos = ids[j];
      c += w;
      if (tt < c) { pick = ids[j]; break; }
    }
    target = pick >= 0 ? pick : lastpos;
    if (target < 0) { dev_fail(S, EXB200_ERR_WEIGHTS, r); target = own; }
  }
  finalize_record(S, r, own, target, single);
  return false;
}

template <bool DP>
__global__ void __launch_bounds__(TPB) k_commit(DevState S, const int* __restrict__ act, int n_act, uint32_t round) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool stay = false;
  int r = -1;
  if (i < n_act) {
    r = act ? act[i] : i;
    const int cnt = S.fin[r] == 1 ? 0x7FFFFFFF : S.cand_cnt[r]; // (final already: wide records, round 0)
    stay = cnt <= S.long_min; // long records live in the other list
    const unsigned long long key = resv_key(round, r);
    const unsigned long long fence = *S.fence;
    const bool fenced = (fence >> 32) == (key >> 32) && (uint32_t)fence < (uint32_t)r;
    if (!stay) {
    } else if (cnt < 0) {
      // serial record: its turn comes when every earlier record is final
      if (S.counters[C_MIN_ACTIVE] == r) S.counters[C_WIDE_READY] = r;
    } else if (!fenced) stay = commit_short<DP>(S, r, cnt, key);
  }
  // append the records that stay to the next active list
  const unsigned m = __ballot_sync(FULL, stay);
  if (m) {
    int base = 0;
    const int lane = threadIdx.x & 31;
    if (lane == (__ffs(m) - 1)) base = atomicAdd(&S.counters[C_NACT_OUT], __popc(m));
    base = __shfl_sync(FULL, base, __ffs(m) - 1);
    if (stay) S.act_out[base + __popc(m & ((1u << lane) - 1u))] = r;
  }
}

// Turn of one record with a long candidate list, taken by a whole BLOCK (any block size that is a
// multiple of 32, up to 1024 threads).  All threads gather sizes and evaluate weights into shared
// memory (s_w: one double per candidate); one thread then accumulates them in candidate order, i.e.
// with exactly the association of the sequential loop in commit_short and in the oracle.
// Returns true (to every thread) if the record stays unfinalised.
struct CommitLongShared {
  double red[32];
  double m1, srel;
  int dec, pick, last;
};
template <bool DP>
__device__ bool commit_long(const DevState& S, int r, unsigned long long key, double* s_w, CommitLongShared& sh) {
  const int cnt = S.cand_cnt[r];
  const int own = S.lambda[r];
  const int* ids = S.cand_id + S.cand_off[r];
  const double* lls = S.cand_ll + S.cand_off[r];
  const int nthr = blockDim.x;
  // a record that owned all its items once owns them until its turn (earlier unfinalised records
  // only disappear) and nothing it reads can change: its weights are summed once
  const bool ready = S.fin[r] == 2;
  int mine = 1;
  if (!ready) {
    mine = S.owner[own] == key && S.owner[S.N + r] == key;
    for (int j = threadIdx.x; j < cnt && mine; j += nthr) mine = S.owner[ids[j]] == key;
    mine = __syncthreads_and(mine);
  }
  if (!mine) return true;
  const bool single = S.ent_size[own] == 1;
  // s_w[j] = running sum (candidate order, summed by one thread: the association of the
  // sequential loop in commit_short and in the oracle) of exp(lw_j - m1), 0 for candidates that are
  // not live; sh.last = last candidate of positive weight; sh.srel = the total
  auto fill_weights = [&](bool need_max) {
    double m = -CUDART_INF;
    for (int j = threadIdx.x; j < cnt; j += nthr) {
      const int sz = live_size(S, ids[j], own, single);
      const double lw = sz > 0 ? cand_lw<DP>(S, r, ids[j], sz, lls[j]) : -CUDART_INF;
      s_w[j] = lw;
      m = lw > m ? lw : m;
    }
    if (threadIdx.x == 0) sh.last = -1;
    if (need_max) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { const double y = __shfl_xor_sync(FULL, m, o); m = y > m ? y : m; }
      if ((threadIdx.x & 31) == 0) sh.red[threadIdx.x >> 5] = m;
      __syncthreads();
      if (threadIdx.x == 0) {
        double mm = -CUDART_INF;
        for (int k = 0; k < nthr / 32; k++) mm = sh.red[k] > mm ? sh.red[k] : mm;
        sh.m1 = mm;
      }
    }
    __syncthreads();
    const double m1 = sh.m1;
    int last = -1;
    for (int j = threadIdx.x; j < cnt; j += nthr) {
      const double w = s_w[j] > -CUDART_INF ? exp(s_w[j] - m1) : 0.0;
      s_w[j] = w;
      if (w > 0.0) last = j;
    }
    if (last >= 0) atomicMax(&sh.last, last);
    __syncthreads();
    if (threadIdx.x == 0) {
      double c = 0.0;
      if (m1 > -CUDART_INF) for (int j = 0; j < cnt; j++) { c += s_w[j]; s_w[j] = c; } // dead ones add 0
      sh.srel = c;
    }
    __syncthreads();
  };
  if (ready) { if (threadIdx.x == 0) { sh.m1 = S.rdy_m1[r]; sh.srel = S.rdy_srel[r]; } __syncthreads(); }
  else fill_weights(true);
  const U2 u = uniforms(S.seed, r, S.iter, PH_LINK_DECIDE, 0, 0);
  if (threadIdx.x == 0) sh.dec = decide_with_bounds(S, r, single, sh.m1, sh.srel, u.a);
  __syncthreads();
  const int dec = sh.dec;
  if (dec == 2) {
    if (threadIdx.x == 0) {
      S.dlo[r] = single ? -1 : 0;
      S.dhi[r] = single ? 0 : 1;
      if (!ready) { S.fin[r] = 2; S.rdy_m1[r] = sh.m1; S.rdy_srel[r] = sh.srel; }
    }
    __syncthreads();
    return true;
  }
  if (dec == 0) {
    if (ready) { const double keep = sh.srel; fill_weights(false); if (threadIdx.x == 0) sh.srel = keep; __syncthreads(); } // the sums were not kept
    // the first candidate whose running sum exceeds the threshold (it has a positive weight:
    // a zero weight leaves the sum where it was); the running sums never decrease
    if (threadIdx.x == 0) sh.pick = 0x7FFFFFFF;
    __syncthreads();
    const double tt = u.b * sh.srel;
    for (int j = threadIdx.x; j < cnt; j += nthr)
      if (tt < s_w[j] && (j == 0 || !(tt < s_w[j - 1]))) atomicMin(&sh.pick, j);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    int target;
    if (dec == 1) target = S.N + r;
    else {
      const int j = sh.pick != 0x7FFFFFFF ? sh.pick : sh.last;
      target = j >= 0 ? ids[j] : -1;
      if (target < 0) { dev_fail(S, EXB200_ERR_WEIGHTS, r); target = own; }
    }
    finalize_record(S, r, own, target, single);
  }
  __syncthreads();
  return false;
}

template <bool DP>
__global__ void __launch_bounds__(TPB) k_commit_w(DevState S, const int* __restrict__ act, int n_act, uint32_t round) {
  extern __shared__ double s_w[]; // [cnt]
  __shared__ CommitLongShared sh;
  const int r = act[blockIdx.x];
  const unsigned long long key = resv_key(round, r);
  const unsigned long long fence = *S.fence;
  const bool fenced = (fence >> 32) == (key >> 32) && (uint32_t)fence < (uint32_t)r;
  bool stay = true;
  if (!fenced) stay = commit_long<DP>(S, r, key, s_w, sh);
  if (stay && threadIdx.x == 0) S.actl_out[atomicAdd(&S.counters[C_NACTL_OUT], 1)] = r;
}

// ------------------------------------------------------------------------------------------
// K1d'  the dependency rounds after the first one, in ONE persistent cooperative kernel.
// After round 1 a tenth of the records is left and every further round shrinks the set; launched
// one by one, rounds 5+ cost a few microseconds of work and ~0.3 ms of launches, two N-wide scans
// of the K bounds and a host synchronisation each.  Here the unfinalised records sit in a list
// SORTED by record id (short and long lists together) and every round is three grid-wide steps:
//   A  reservations (atomicMin of (round, id) on every item the record reads or may write)
//   B  turns: records that own all they reserved decide and move (commit_short / commit_long);
//      every position notes whether its record stays and by how much its K bounds moved
//   C  one scan over the LIST of (stay, d dlo, d dhi): stable compaction into the next list, and
//      plo / phi of the records that stay move by the prefix of the deltas before them - all the
//      records whose bounds changed are in the list, and the list is sorted, so this is the
//      exclusive prefix sum over all records that the per-round N-wide scans used to recompute.
// Same arithmetic per record as k_commit / k_commit_w, so the chain is the same bit for bit.
// ------------------------------------------------------------------------------------------
struct TailState {
  int* list[2];     // ping-pong lists of unfinalised records, ascending id
  int* n;           // [0] length of the current list (in: from the selection after round 1), [1] rounds run,
                    // [2] status (1: a round made no progress - internal error), [3] list length at exit,
                    // [4] next tile of 32 positions to hand out, [5] long records of the round, [6] next of them
  uint8_t* flag;    // [n] per position: stay | (d dlo + 2) << 1 | (d dhi + 2) << 4
  int* blk_tot;     // [grid][3] per chunk of positions: records that stay, sum of d dlo, sum of d dhi
  int* long_pos;    // [n] positions holding a record with a long list (this round)
  unsigned long long* dbg; // trace: [round][4] = (records, ns after A, after B, after C) of the first 64 rounds (may be null)
  unsigned long long* bar; // arrivals at the kernel's own barrier (zero at launch)
};
// Barrier over the first `n_blocks` blocks of the grid (the others have left the kernel): every block
// adds one to a counter that only grows and waits until it holds all the arrivals so far.  A round
// over a few thousand records is run by a few blocks, and their barrier costs a fraction of one
// over the whole grid.
__device__ __forceinline__ void tail_barrier(unsigned long long* counter, unsigned long long& expected, int n_blocks) {
  __syncthreads();
  expected += (unsigned long long)n_blocks;
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(counter), "l"(1ull) : "memory");
    unsigned long long v;
    do { asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(counter) : "memory"); } while (v < expected);
  }
  __syncthreads();
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
constexpr int TAIL_TPB = 512;

// Steps A and B hand out tiles of 32 positions to warps from a shared cursor and have no block
// barrier: the cost of a record's turn varies by orders of magnitude (a chain of dependent loads per
// candidate), and a barrier per tile made every tile as slow as its slowest record.  Records with
// long lists are reserved by their warp (lanes strided over the list) and listed; their turns are
// taken by whole blocks, again handed out one at a time, while other blocks still work on tiles -
// turns of one round touch disjoint items, so the order does not matter.
template <bool DP>
__global__ void __launch_bounds__(TAIL_TPB, 2) k_rounds_tail(DevState S, TailState T, uint32_t round0, int max_rounds) {
  extern __shared__ double s_w[]; // [long_cap] weights of a long record
  __shared__ CommitLongShared sh;
  __shared__ int s_red[3][TAIL_TPB / 32];
  __shared__ int s_carry[3];
  __shared__ int s_next;
  const int b = blockIdx.x, lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  int cur = 0, rounds = 0;
  int n = T.n[0];
  int n_long_prev = gridDim.x; // long records of the previous round (they take a block each)
  unsigned long long expected = 0; // arrivals at the barrier so far
  if (T.dbg && b == 0 && threadIdx.x == 0) T.dbg[256] = global_ns();
  while (n > 0 && rounds < max_rounds) {
    // The blocks that run this round: about one record per thread plus a block per long record.  The
    // list only shrinks, so a block that is not needed now never is again and leaves.
    const int nb = (int)min((long long)gridDim.x, (long long)(n + TAIL_TPB - 1) / TAIL_TPB + (long long)n_long_prev);
    if (b >= nb) break;
    const int gw = b * (TAIL_TPB / 32) + wp, n_warps = nb * (TAIL_TPB / 32);
    const uint32_t round = round0 + (uint32_t)rounds;
    const int* in = T.list[cur];
    int* out = T.list[cur ^ 1];
    // chunks of positions (a multiple of 32) whose totals step C scans: chunk c = positions [c per, (c+1) per)
    const int per = ((n + nb - 1) / nb + 31) & ~31;
    // ---- A: reservations (tiles assigned round-robin: the work per record is a few atomics)
    if (threadIdx.x < 3) T.blk_tot[3 * b + threadIdx.x] = 0;
    for (int t0 = gw * 32; t0 < n; t0 += n_warps * 32) {
      const int i = t0 + lane;
      int r = -1, cnt = 0;
      if (i < n) { r = in[i]; cnt = S.cand_cnt[r]; }
      const bool is_long = r >= 0 && cnt > S.long_min;
      if (r >= 0 && !is_long) reserve_short(S, r, cnt, resv_key(round, r));
      unsigned lm = __ballot_sync(FULL, is_long);
      if (lm) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&T.n[5], __popc(lm));
        base = __shfl_sync(FULL, base, 0);
        if (is_long) T.long_pos[base + __popc(lm & ((1u << lane) - 1u))] = i;
        for (; lm; lm &= lm - 1) { // the warp reserves the items of each long record
          const int src = __ffs(lm) - 1;
          const int r2 = __shfl_sync(FULL, r, src), c2 = __shfl_sync(FULL, cnt, src);
          const unsigned long long key = resv_key(round, r2);
          const int* ids = S.cand_id + S.cand_off[r2];
          if (lane == 0) { atomicMin(&S.owner[S.lambda[r2]], key); atomicMin(&S.owner[S.N + r2], key); }
          for (int j = lane; j < c2; j += 128) { // four independent loads in flight per lane
            int id[4];
#pragma unroll
            for (int u = 0; u < 4; u++) id[u] = j + 32 * u < c2 ? ids[j + 32 * u] : -1;
#pragma unroll
            for (int u = 0; u < 4; u++) if (id[u] >= 0) atomicMin(&S.owner[id[u]], key);
          }
        }
      }
    }
    tail_barrier(T.bar, expected, nb);
    if (T.dbg && b == 0 && threadIdx.x == 0 && rounds < 64) { T.dbg[4 * rounds] = (unsigned long long)n; T.dbg[4 * rounds + 1] = global_ns(); }
    // ---- B: turns.  Short lists: warps take tiles from the cursor ...
    for (;;) {
      int t0 = 0;
      if (lane == 0) t0 = atomicAdd(&T.n[4], 32);
      t0 = __shfl_sync(FULL, t0, 0);
      if (t0 >= n) break;
      const int i = t0 + lane;
      int stay = 0, dl = 0, dh = 0;
      if (i < n) {
        const int r = in[i];
        const int cnt = S.cand_cnt[r];
        if (cnt <= S.long_min) {
          const int lo0 = S.dlo[r], hi0 = S.dhi[r];
          stay = commit_short<DP>(S, r, cnt, resv_key(round, r)) ? 1 : 0;
          dl = (int)S.dlo[r] - lo0; dh = (int)S.dhi[r] - hi0;
          T.flag[i] = (uint8_t)(stay | ((dl + 2) << 1) | ((dh + 2) << 4));
        }
      }
      stay = warp_sum(stay); dl = warp_sum(dl); dh = warp_sum(dh);
      if (lane == 0) { // a tile lies inside one chunk
        int* tot = T.blk_tot + 3 * (t0 / per);
        if (stay) atomicAdd(&tot[0], stay);
        if (dl) atomicAdd(&tot[1], dl);
        if (dh) atomicAdd(&tot[2], dh);
      }
    }
    __syncthreads();
    // ... long lists: blocks take them one at a time
    const int n_long = T.n[5];
    n_long_prev = n_long;
    for (;;) {
      if (threadIdx.x == 0) s_next = atomicAdd(&T.n[6], 1);
      __syncthreads();
      const int k = s_next;
      if (k >= n_long) break;
      const int i2 = T.long_pos[k], r = in[i2];
      const int lo0 = S.dlo[r], hi0 = S.dhi[r];
      const bool stay = commit_long<DP>(S, r, resv_key(round, r), s_w, sh); // ends with a barrier
      if (threadIdx.x == 0) {
        const int dl = (int)S.dlo[r] - lo0, dh = (int)S.dhi[r] - hi0;
        T.flag[i2] = (uint8_t)((stay ? 1 : 0) | ((dl + 2) << 1) | ((dh + 2) << 4));
        int* tot = T.blk_tot + 3 * (i2 / per);
        if (stay) atomicAdd(&tot[0], 1);
        if (dl) atomicAdd(&tot[1], dl);
        if (dh) atomicAdd(&tot[2], dh);
      }
    }
    tail_barrier(T.bar, expected, nb);
    if (T.dbg && b == 0 && threadIdx.x == 0 && rounds < 64) T.dbg[4 * rounds + 2] = global_ns();
    // ---- C: prefix over the chunks before this block's, then a scan of the chunk's own positions
    {
      if (b == 0 && threadIdx.x == 0) { T.n[4] = 0; T.n[5] = 0; T.n[6] = 0; } // cursors of the next round
      const int i0 = min(n, b * per), i1 = min(n, i0 + per);
      int a0 = 0, a1 = 0, a2 = 0, tot = 0;
      for (int k = threadIdx.x; k < nb; k += TAIL_TPB) {
        const int st = T.blk_tot[3 * k];
        tot += st;
        if (k < b) { a0 += st; a1 += T.blk_tot[3 * k + 1]; a2 += T.blk_tot[3 * k + 2]; }
      }
      a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); tot = warp_sum(tot);
      if (lane == 0) { s_red[0][wp] = a0; s_red[1][wp] = a1; s_red[2][wp] = a2; }
      __syncthreads();
      if (threadIdx.x < 3) {
        int t = 0;
        for (int k = 0; k < TAIL_TPB / 32; k++) t += s_red[threadIdx.x][k];
        s_carry[threadIdx.x] = t;
      }
      __syncthreads();
      if (lane == 0) s_red[0][wp] = tot; // every block computes the same total: the next n
      __syncthreads();
      int n_next = 0;
      for (int k = 0; k < TAIL_TPB / 32; k++) n_next += s_red[0][k];
      for (int t0 = i0; t0 < i1; t0 += TAIL_TPB) {
        const int i = t0 + threadIdx.x;
        int st = 0, dl = 0, dh = 0, r = -1;
        if (i < i1) {
          const unsigned f = T.flag[i];
          st = f & 1u; dl = (int)((f >> 1) & 7u) - 2; dh = (int)((f >> 4) & 7u) - 2;
          r = in[i];
        }
        int x0 = st, x1 = dl, x2 = dh;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int y0 = __shfl_up_sync(FULL, x0, o), y1 = __shfl_up_sync(FULL, x1, o), y2 = __shfl_up_sync(FULL, x2, o);
          if (lane >= o) { x0 += y0; x1 += y1; x2 += y2; }
        }
        __syncthreads(); // s_red reads of the previous step are done
        if (lane == 31) { s_red[0][wp] = x0; s_red[1][wp] = x1; s_red[2][wp] = x2; }
        __syncthreads();
        int o0 = s_carry[0], o1 = s_carry[1], o2 = s_carry[2], t0s = 0, t1s = 0, t2s = 0;
        for (int k = 0; k < TAIL_TPB / 32; k++) {
          const int v0 = s_red[0][k], v1 = s_red[1][k], v2 = s_red[2][k];
          if (k < wp) { o0 += v0; o1 += v1; o2 += v2; }
          t0s += v0; t1s += v1; t2s += v2;
        }
        if (st) {
          out[o0 + x0 - st] = r;
          if (S.use_kbounds) { S.plo[r] += o1 + x1 - dl; S.phi_[r] += o2 + x2 - dh; }
        }
        __syncthreads();
        if (threadIdx.x == 0) { s_carry[0] += t0s; s_carry[1] += t1s; s_carry[2] += t2s; }
        __syncthreads();
      }
      if (T.dbg && b == 0 && threadIdx.x == 0 && rounds < 64) T.dbg[4 * rounds + 3] = global_ns();
      rounds++;
      if (n_next >= n) { // no record took its turn: cannot happen (the smallest unfinalised record always can)
        if (b == 0 && threadIdx.x == 0) T.n[2] = 1;
        n = 0;
      } else n = n_next;
      cur ^= 1;
    }
    tail_barrier(T.bar, expected, nb);
  }
  if (b == 0 && threadIdx.x == 0) { T.n[1] = rounds; T.n[3] = n; }
}

// ------------------------------------------------------------------------------------------
// serial path (records without a stored candidate list) and test hook: the conditional of ONE record
// against every item, in item order.
// ------------------------------------------------------------------------------------------
// test hook: lwbuf[i] = log prior_weight_existing + logLikelihoodWeightExisting, NaN if no candidate
__global__ void __launch_bounds__(TPB) k_eval_all(DevState S, int r, int n_items) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_items) return;
  double out = CUDART_NAN;
  const int own = S.lambda[r];
  int sz = S.ent_size[i];
  if (i == own) sz = sz == 1 ? 0 : sz - 1;
  if (sz > 0 && i != S.N + r) {
    const int* x = S.rec_x + (size_t)r * S.RS;
    const uint32_t zm = S.rec_z[r];
    const int* y = S.ent_y + (size_t)i * S.ES;
    if (item_compatible(S, x, zm, y)) out = S.has_dp ? cand_lw<true>(S, r, i, sz, item_loglik(S, x, zm, y)) : cand_lw<false>(S, r, i, sz, item_loglik(S, x, zm, y));
  }
  S.lwbuf[i] = out;
}

// Records without a stored candidate list (more than wide_threshold possible links, or no usable
// blocking bucket): their possible links among ALL the items of the sweep are listed here, up to
// WIDE_NB records per pass over the item tuples.  The item range is cut into WIDE_PIECES contiguous
// pieces, one warp each; a counting pass, a scan per record and a filling pass leave every record's
// (item, log-likelihood) list in ascending item order - no sort.  Items are static during the links
// update; sizes are read when the record's turn comes (k_wide_reduce / k_wide_commit).
constexpr int WIDE_PIECES = WIDE_BLOCKS * (TPB / 32);
constexpr int WIDE_Q = 256; // a group of four records adds at most 128 pairs to fewer than 32 waiting ones
__device__ __forceinline__ void wide_piece(long long n, int n_pieces, int p, long long& lo, long long& hi) {
  const long long per = (n + n_pieces - 1) / n_pieces;
  lo = min(n, (long long)p * per);
  hi = min(n, lo + per);
}
// FILL = false: wl_cnt[b][piece] = possible links of record list[b] in the piece
// FILL = true : entries of the records with a base (wl_base[b] != ~0) written at
//               wl_base[b] + (scanned) wl_cnt[b][piece]
// Slots [0,N) that held no entity when the sweep started (live0) are left out: that is the
// specification's count of possible links, and it does not change while records move between the
// two passes (sizes are checked again at the record's turn).
// W-bit hash of an attribute value; W = 64 / (attributes in the packed word), at most 16
__device__ __forceinline__ unsigned long long hash_w(int v, int W) { return (unsigned long long)(((uint32_t)v * 0x9E3779B1u) >> (32 - W)); }
template <bool FILL>
__global__ void __launch_bounds__(TPB) k_wide_scan(DevState S, const int* __restrict__ list, int nb) {
  // records taking part in this pass, compacted: k -> batch position s_b[k]
  __shared__ ulonglong2 s_mv[WIDE_NB];       // field a of .x: all ones if attribute a is pinned; of .y: hash of the pinned value
  __shared__ int s_x[WIDE_NB][8];            // pinned value of the first 8 attributes (-1: free)
  __shared__ int s_r[WIDE_NB], s_b[WIDE_NB];
  __shared__ uint8_t s_more[WIDE_NB];        // the record pins an attribute beyond the eighth
  __shared__ int s_run[TPB / 32][WIDE_NB];
  __shared__ uint32_t s_q[TPB / 32][WIDE_Q]; // per-warp queue of (record << 24 | item - first item of the piece) pairs waiting for their likelihood
  __shared__ int s_n;
  if (threadIdx.x == 0) {
    int n = 0;
    for (int b = 0; b < nb; b++) if (!FILL || S.wl_base[b] != ~0ull) s_b[n++] = b;
    s_n = n;
  }
  __syncthreads();
  const int n_act = s_n;
  const int W = min(16, 64 / min(S.A, 8));
  const unsigned long long field = (1ull << W) - 1ull;
  for (int k = threadIdx.x; k < WIDE_NB; k += TPB) {
    if (k >= n_act) { s_mv[k] = make_ulonglong2(~0ull, 0ull); continue; } // padding of the last group of four: matches (almost) nothing

    const int r = list[s_b[k]];
    s_r[k] = r;
    const uint32_t zm = S.rec_z[r];
    unsigned long long M = 0, V = 0;
    for (int a = 0; a < 8; a++) {
      const int xv = a < S.A ? S.rec_x[(size_t)r * S.RS + a] : -1;
      const bool pinned = xv >= 0 && !((zm >> a) & 1u);
      s_x[k][a] = pinned ? xv : -1;
      if (pinned) { M |= field << (W * a); V |= hash_w(xv, W) << (W * a); }
    }
    s_mv[k] = make_ulonglong2(M, V);
    bool more = false;
    for (int a = 8; a < S.A; a++) more |= S.rec_x[(size_t)r * S.RS + a] >= 0 && !((zm >> a) & 1u);
    s_more[k] = more;
  }
  for (int q = threadIdx.x; q < (TPB / 32) * WIDE_NB; q += TPB) (&s_run[0][0])[q] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5, p = blockIdx.x * (TPB / 32) + wp;
  long long lo, hi;
  wide_piece(2LL * S.N, WIDE_PIECES, p, lo, hi);
  int qh = 0, qn = 0; // queue head and length (warp-uniform)
  // The counting pass logs its hits per piece (record, item, log-likelihood; in the order they are
  // found), so that the filling pass only replays the log (k_wide_fill_log); a piece whose log
  // overflows is filled by repeating the scan.
  int lg = 0;
  if (FILL && S.wl_logn[p] >= 0) hi = lo; // logged piece: nothing to do here
  for (long long base = lo; base < hi; base += 32) {
    const long long i = base + lane;
    bool live = i < hi;
    int y[8];
    const int* yp = S.ent_y + (size_t)(live ? i : lo) * S.ES;
    {
      const int4 y0 = *reinterpret_cast<const int4*>(yp), y1 = *reinterpret_cast<const int4*>(yp + 4);
      y[0] = y0.x; y[1] = y0.y; y[2] = y0.z; y[3] = y0.w; y[4] = y1.x; y[5] = y1.y; y[6] = y1.z; y[7] = y1.w;
    }
    if (live && i < S.N) live = S.live0[i] != 0;
    unsigned long long P = 0;
#pragma unroll
    for (int a = 0; a < 8; a++) if (a * W < 64) P |= hash_w(y[a], W) << (W * a);
    // Pairs that pass the pinned attributes are queued per warp and their likelihoods evaluated 32
    // at a time, one pair per lane: the evaluation is a chain of dependent loads (neighbourhood
    // search), far too slow to run with a lane or two active.  The queue is filled in (step,
    // record, item) order and emptied in that order, so every record sees its items in ascending order.
    auto drain = [&](int cnt) {
      int k = -1 - lane, rank = 0; // inactive lanes: distinct keys that match nobody
      long long it = 0;
      double ll = 0.0;
      bool hit = false;
      if (lane < cnt) {
        const unsigned e = s_q[wp][(qh + lane) & (WIDE_Q - 1)];
        k = (int)(e >> 24);
        it = lo + (long long)(e & 0xFFFFFFu);
        const int r = k < n_act ? s_r[k] : 0;
        hit = k < n_act && it != (long long)S.N + r;
        if (hit) {
          const int* x = S.rec_x + (size_t)r * S.RS;
          const uint32_t zm = S.rec_z[r];
          const int* yq = S.ent_y + (size_t)it * S.ES;
          { // the hashed pre-test let it through: now the pinned attributes themselves
            const int4 q0 = *reinterpret_cast<const int4*>(yq), q1 = *reinterpret_cast<const int4*>(yq + 4);
            const int yv[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
            for (int a = 0; a < 8; a++) { const int xv = s_x[k][a]; hit &= (xv < 0) | (yv[a] == xv); }
          }
          if (hit && s_more[k]) hit = item_compatible(S, x, zm, yq);
          if (hit) { ll = item_loglik(S, x, zm, yq); hit = ll > -CUDART_INF; }
        }
      }
      const unsigned hm = __ballot_sync(FULL, hit);
      if (hm) {
        const unsigned peers = __match_any_sync(FULL, k) & hm;
        if (hit) {
          rank = __popc(peers & ((1u << lane) - 1u));
          if (FILL) {
            const int b = s_b[k];
            const unsigned long long pos = S.wl_base[b] + (unsigned long long)S.wl_cnt[(size_t)b * WIDE_PIECES + p] +
                                           (unsigned long long)(s_run[wp][k] + rank);
            S.wl_id[pos] = (int)it;
            S.wl_ll[pos] = ll;
          }
        }
        if (!FILL) {
          const int lpos = lg + __popc(hm & ((1u << lane) - 1u));
          if (hit && lpos < S.wl_logcap) {
            const size_t o = (size_t)p * S.wl_logcap + lpos;
            S.wl_log_k[o] = (uint8_t)s_b[k]; S.wl_log_item[o] = (int)it; S.wl_log_ll[o] = ll;
          }
          lg += __popc(hm);
        }
        __syncwarp();
        if (hit && rank == 0) s_run[wp][k] += __popc(peers);
      }
      __syncwarp();
      qh += cnt; qn -= cnt;
    };
    for (int k4 = 0; k4 < n_act; k4 += 4) {
      // hashed pre-test of four records at once: exact for nearly every pair
      const ulonglong2 mv0 = s_mv[k4], mv1 = s_mv[k4 + 1], mv2 = s_mv[k4 + 2], mv3 = s_mv[k4 + 3];
      const bool h0 = ((P ^ mv0.y) & mv0.x) == 0, h1 = ((P ^ mv1.y) & mv1.x) == 0;
      const bool h2 = ((P ^ mv2.y) & mv2.x) == 0, h3 = ((P ^ mv3.y) & mv3.x) == 0;
      if (!__any_sync(FULL, live && (h0 | h1 | h2 | h3))) continue;
      const unsigned lt = (1u << lane) - 1u;
      const uint32_t off = (uint32_t)(i - lo); // (a piece holds far fewer than 2^24 items)
      const unsigned m0 = __ballot_sync(FULL, live && h0), m1 = __ballot_sync(FULL, live && h1);
      const unsigned m2 = __ballot_sync(FULL, live && h2), m3 = __ballot_sync(FULL, live && h3);
      // (the padding records of the last group match next to nothing and are dropped in drain)
      if (m0) { if (live && h0) s_q[wp][(qh + qn + __popc(m0 & lt)) & (WIDE_Q - 1)] = ((uint32_t)k4 << 24) | off; qn += __popc(m0); }
      if (m1) { if (live && h1) s_q[wp][(qh + qn + __popc(m1 & lt)) & (WIDE_Q - 1)] = ((uint32_t)(k4 + 1) << 24) | off; qn += __popc(m1); }
      if (m2) { if (live && h2) s_q[wp][(qh + qn + __popc(m2 & lt)) & (WIDE_Q - 1)] = ((uint32_t)(k4 + 2) << 24) | off; qn += __popc(m2); }
      if (m3) { if (live && h3) s_q[wp][(qh + qn + __popc(m3 & lt)) & (WIDE_Q - 1)] = ((uint32_t)(k4 + 3) << 24) | off; qn += __popc(m3); }
      __syncwarp();
      while (qn >= 32) drain(32);
    }
    // (what is left waits for the pairs of the next 32 items: a drain costs a chain of dependent loads
    // whatever the number of lanes that take part in it)
    if (qn && base + 32 >= hi) drain(qn);
  }
  if (!FILL) {
    __syncwarp();
    for (int k = lane; k < n_act; k += 32) S.wl_cnt[(size_t)s_b[k] * WIDE_PIECES + p] = s_run[wp][k];
    if (lane == 0) S.wl_logn[p] = lg <= S.wl_logcap ? lg : -1;
  }
}
// filling pass from the logs: one warp per piece replays its hits in order
__global__ void __launch_bounds__(TPB) k_wide_fill_log(DevState S, int nb) {
  __shared__ int s_run[TPB / 32][WIDE_NB];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5, p = blockIdx.x * (TPB / 32) + wp;
  for (int b = lane; b < nb; b += 32) s_run[wp][b] = 0;
  __syncwarp();
  const int n = S.wl_logn[p];
  for (int e0 = 0; e0 < n; e0 += 32) {
    const int e = e0 + lane;
    int b = -1 - lane, item = 0;
    double ll = 0.0;
    bool act = false;
    if (e < n) {
      const size_t o = (size_t)p * S.wl_logcap + e;
      b = (int)S.wl_log_k[o]; item = S.wl_log_item[o]; ll = S.wl_log_ll[o];
      act = S.wl_base[b] != ~0ull;
      if (!act) b = -1 - lane;
    }
    const unsigned am = __ballot_sync(FULL, act);
    if (am) {
      const unsigned peers = __match_any_sync(FULL, b) & am;
      int rank = 0;
      if (act) {
        rank = __popc(peers & ((1u << lane) - 1u));
        const unsigned long long pos = S.wl_base[b] + (unsigned long long)S.wl_cnt[(size_t)b * WIDE_PIECES + p] +
                                       (unsigned long long)(s_run[wp][b] + rank);
        S.wl_id[pos] = item;
        S.wl_ll[pos] = ll;
      }
      __syncwarp();
      if (act && rank == 0) s_run[wp][b] += __popc(peers);
    }
    __syncwarp();
  }
}
// one block per record of the batch: exclusive scan of its piece counts, total in wl_tot
__global__ void __launch_bounds__(1024) k_wide_offsets(DevState S) {
  __shared__ int s_w[32];
  int* c = S.wl_cnt + (size_t)blockIdx.x * WIDE_PIECES;
  constexpr int PER = WIDE_PIECES / 1024;
  int v[PER], t = 0;
#pragma unroll
  for (int k = 0; k < PER; k++) { v[k] = c[threadIdx.x * PER + k]; t += v[k]; }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int x = t;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULL, x, o); if (lane >= o) x += y; }
  if (lane == 31) s_w[w] = x;
  __syncthreads();
  if (w == 0) {
    int z = s_w[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULL, z, o); if (lane >= o) z += y; }
    s_w[lane] = z;
  }
  __syncthreads();
  int excl = x - t + (w ? s_w[w - 1] : 0);
#pragma unroll
  for (int k = 0; k < PER; k++) { c[threadIdx.x * PER + k] = excl; excl += v[k]; }
  if (threadIdx.x == 1023) S.wl_tot[blockIdx.x] = excl;
}
// records that turned out to have at most wide_threshold possible links keep their list like
// everybody else (one block per record; the list is sorted already)
__global__ void __launch_bounds__(TPB) k_wide_store(DevState S, const int* __restrict__ list, const int* __restrict__ sel, int cap) {
  __shared__ unsigned long long s_off;
  const int b = sel[blockIdx.x], r = list[b];
  const int cnt = S.wl_tot[b];
  const unsigned long long base = S.wl_base[b];
  const int own = S.lambda[r];
  if (threadIdx.x == 0) s_off = atomicAdd(S.pool_cursor, (unsigned long long)cnt);
  __syncthreads();
  const unsigned long long o = s_off;
  const bool overflow = o + (unsigned long long)cnt > S.pool_cap;
  if (!overflow) for (int i = threadIdx.x; i < cnt; i += TPB) {
    const int id = S.wl_id[base + i];
    S.cand_id[o + i] = id; S.cand_ll[o + i] = S.wl_ll[base + i];
    if (id != own) S.referenced[id] = 1;
  }
  if (threadIdx.x == 0) {
    const bool other = overflow || cnt > 1 || (cnt == 1 && S.wl_id[base] != own);
    S.cand_off[r] = overflow ? 0u : (uint32_t)o;
    S.cand_cnt[r] = overflow ? -2 : cnt;
    S.dlo[r] = other ? -1 : 0;
    S.dhi[r] = 1;
    if (overflow) atomicAdd(&S.counters[C_NFENCE], 1);
    else {
      atomicAdd((unsigned long long*)&S.counters64[0], (unsigned long long)cnt);
      if (cnt > S.long_min) S.actl_in[atomicAdd(&S.counters[C_NACTL_INIT], 1)] = r;
    }
  }
}

// Turn of record r, listed at wl_base = `base` with `len` entries: every warp of the grid reduces a
// contiguous piece (lanes strided: coalesced) to (max, sum of exp(lw - max)) in wide_part[2p], [2p+1].
// (src = 1: the list is in wlb_id / wlb_ll - possible links listed from the tables)
// `vblock` of `n_vblocks`: the block's place in the grid of the stand-alone kernel (the persistent
// kernel k_wide_turns gives every block several).
// A turn is a chain of dependent memory accesses, so what does not depend on the turns before it
// (the list entries - WideTile; the record's own entity) is read before the wait for the previous
// turn, and the sums take two passes (maximum, then independent exponentials) instead of a running
// rescaled sum whose every step waits for the last one.
struct WideTile { int id[8]; double l[8]; };
__device__ __forceinline__ void wide_piece_of(int len, int n_vblocks, int vblock, long long& lo, long long& hi) {
  wide_piece(len, n_vblocks * (TPB / 32), vblock * (TPB / 32) + (threadIdx.x >> 5), lo, hi);
}
__device__ __forceinline__ void wide_load_tile(const int* ids, const double* ll, long long i0, long long hi, WideTile& t) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const long long i = i0 + j * 32 + lane;
    t.id[j] = i < hi ? ids[i] : -1;
    t.l[j] = i < hi ? ll[i] : 0.0;
  }
}
// sizes are read past the L1 (they change between the turns of one kernel)
__device__ __forceinline__ int live_size_cg(const DevState& S, int e, int own, bool single) {
  const int sz = __ldcg(&S.ent_size[e]);
  return e == own ? (single ? 0 : sz - 1) : sz;
}
template <bool DP>
__device__ __forceinline__ void wide_reduce_block(const DevState& S, int r, unsigned long long base, int len, int src, int vblock,
                                                  int n_vblocks, int own, bool single, const WideTile* first) {
  const int* ids = (src ? S.wlb_id : S.wl_id) + base;
  const double* ll = (src ? S.wlb_ll : S.wl_ll) + base;
  const int lane = threadIdx.x & 31, p = vblock * (TPB / 32) + (threadIdx.x >> 5);
  long long lo, hi;
  wide_piece_of(len, n_vblocks, vblock, lo, hi);
  double m = -CUDART_INF, sum = 0.0;
  for (long long i0 = lo; i0 < hi; i0 += 256) { // 8 entries per lane in flight (a piece has at most 256 unless the grid is capped)
    WideTile t;
    if (first && i0 == lo) t = *first; else wide_load_tile(ids, ll, i0, hi, t);
    int sz[8];
#pragma unroll
    for (int j = 0; j < 8; j++) sz[j] = t.id[j] >= 0 ? live_size_cg(S, t.id[j], own, single) : 0;
    double v[8], mt = -CUDART_INF;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      v[j] = sz[j] > 0 ? cand_lw<DP>(S, r, t.id[j], sz[j], t.l[j]) : -CUDART_INF;
      mt = v[j] > mt ? v[j] : mt;
    }
    if (!(mt > -CUDART_INF)) continue;
    const double mn = mt > m ? mt : m;
    double st = 0.0;
#pragma unroll
    for (int j = 0; j < 8; j++) if (v[j] > -CUDART_INF) st += exp(v[j] - mn);
    sum = (m > -CUDART_INF ? sum * exp(m - mn) : 0.0) + st;
    m = mn;
  }
  // the warp's maximum first, then one rescaling per lane and a plain sum
  double mw = m;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const double m2 = __shfl_xor_sync(FULL, mw, o); mw = m2 > mw ? m2 : mw; }
  double sw = m > -CUDART_INF ? sum * exp(m - mw) : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sw += __shfl_xor_sync(FULL, sw, o);
  if (lane == 0) { S.wide_part[2 * p] = mw; S.wide_part[2 * p + 1] = sw; }
}

template <bool DP>
__global__ void __launch_bounds__(TPB) k_wide_reduce(DevState S, int r, unsigned long long base, int len, int src) {
  const int own = S.lambda[r];
  wide_reduce_block<DP>(S, r, base, len, src, (int)blockIdx.x, (int)gridDim.x, own, S.ent_size[own] == 1, nullptr);
}

// decision "found a new entity" with the record's new-entity log-likelihood at hand (decide_new reads it)
__device__ __forceinline__ bool decide_new_l(const DevState& S, double lnew, int k, double m1, double srel, double u1, bool& invalid) {
  const double lwn = log(prior_new(S.cl, k)) + lnew;
  const double m2 = m1 > lwn ? m1 : lwn;
  invalid = !(m2 > -CUDART_INF);
  const double wa = exp(lwn - m2);
  const double wb = m1 > -CUDART_INF ? srel * exp(m1 - m2) : 0.0;
  return u1 * (wa + wb) < wa;
}

// decision and commit of record r (one block; n_blocks = grid of the k_wide_reduce before it):
// scale the piece sums to the global max in parallel, walk block sums and piece sums in order,
// then one warp scans the chosen piece in order.  `u`, `lnew`: the record's uniforms and new-entity
// log-likelihood (static: the caller has them before the sums are ready).
template <bool DP>
__device__ void wide_commit_block(const DevState& S, int r, unsigned long long base, int len, int n_blocks, int src, int own, bool single,
                                  U2 u, double lnew) {
  __shared__ double s_bsum[WIDE_BLOCKS];
  __shared__ double s_red[TPB / 32];
  __shared__ double s_tt, s_c;
  __shared__ int s_piece, s_target, s_tail;
  const int* ids = (src ? S.wlb_id : S.wl_id) + base;
  const double* ll = (src ? S.wlb_ll : S.wl_ll) + base;
  const int n_pieces = n_blocks * (TPB / 32);
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  double* scaled = S.wide_part + 2 * WIDE_PIECES;
  double m = -CUDART_INF;
  for (int p = threadIdx.x; p < n_pieces; p += TPB) { const double mp = __ldcg(&S.wide_part[2 * p]); m = mp > m ? mp : m; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const double m2 = __shfl_xor_sync(FULL, m, o); m = m2 > m ? m2 : m; }
  if (lane == 0) s_red[wp] = m;
  __syncthreads();
  double m1 = s_red[0];
#pragma unroll
  for (int k = 1; k < TPB / 32; k++) m1 = s_red[k] > m1 ? s_red[k] : m1;
  for (int p = threadIdx.x; p < n_pieces; p += TPB) {
    const double mp = __ldcg(&S.wide_part[2 * p]);
    scaled[p] = mp > -CUDART_INF ? __ldcg(&S.wide_part[2 * p + 1]) * exp(mp - m1) : 0.0;
  }
  __syncthreads();
  for (int b = threadIdx.x; b < n_blocks; b += TPB) {
    double t = 0.0;
    for (int w = 0; w < TPB / 32; w++) t += scaled[b * (TPB / 32) + w];
    s_bsum[b] = t;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double srel = 0.0;
    for (int b = 0; b < n_blocks; b++) srel += s_bsum[b];
    // (read past the L1: the counter is moved by atomics, which a cached copy of this very block would not see)
    const int k = S.K0 + __ldcg(&S.counters[C_DK_TOTAL]) - (single ? 1 : 0); // every earlier record is final
    bool invalid = false;
    const bool is_new = decide_new_l(S, lnew, k, m1, srel, u.a, invalid);
    if (invalid) dev_fail(S, EXB200_ERR_WEIGHTS, r);
    s_piece = -1; s_tail = 0; s_target = is_new ? S.N + r : -1;
    if (!is_new) {
      const double tt = u.b * srel;
      double c = 0.0;
      int blk = -1, lastb = -1;
      for (int b = 0; b < n_blocks; b++) {
        if (s_bsum[b] > 0.0) lastb = b;
        if (tt < c + s_bsum[b]) { blk = b; break; }
        c += s_bsum[b];
      }
      bool tail = false;
      if (blk < 0) { blk = lastb; tail = true; } // rounding: the last positive item
      if (blk >= 0) {
        int pc = -1, lastp = -1;
        for (int w = 0; w < TPB / 32; w++) {
          const int p = blk * (TPB / 32) + w;
          if (scaled[p] > 0.0) lastp = p;
          if (!tail && tt < c + scaled[p]) { pc = p; break; }
          if (!tail) c += scaled[p];
        }
        if (pc < 0) { pc = lastp; tail = true; }
        s_piece = pc;
      }
      s_tt = tt; s_c = c; s_tail = tail;
    }
  }
  __syncthreads();
  if (s_piece >= 0 && threadIdx.x < 32) {
    long long lo, hi;
    wide_piece(len, n_pieces, s_piece, lo, hi);
    const double tt = s_tt;
    const bool tail = s_tail != 0;
    double c = s_c;
    int target = -1, lastpos = -1;
    for (long long i0 = lo; i0 < hi && target < 0; i0 += 256) {
      WideTile t;
      wide_load_tile(ids, ll, i0, hi, t);
      int sz[8];
#pragma unroll
      for (int j = 0; j < 8; j++) sz[j] = t.id[j] >= 0 ? live_size_cg(S, t.id[j], own, single) : 0;
      double wj[8];
#pragma unroll
      for (int j = 0; j < 8; j++) wj[j] = sz[j] > 0 ? exp(cand_lw<DP>(S, r, t.id[j], sz[j], t.l[j]) - m1) : 0.0;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        if (target >= 0 || i0 + j * 32 >= hi) break; // warp-uniform
        const double w = wj[j];
        double x = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const double y = __shfl_up_sync(FULL, x, o); if (lane >= o) x = x + y; }
        const unsigned pos = __ballot_sync(FULL, w > 0.0);
        if (pos) lastpos = __shfl_sync(FULL, t.id[j], 31 - __clz(pos));
        const unsigned hit = __ballot_sync(FULL, !tail && w > 0.0 && tt < c + x);
        if (hit) target = __shfl_sync(FULL, t.id[j], __ffs(hit) - 1);
        c += __shfl_sync(FULL, x, 31);
      }
    }
    if (lane == 0) s_target = target >= 0 ? target : lastpos;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int target = s_target;
    if (target < 0) { dev_fail(S, EXB200_ERR_WEIGHTS, r); target = own; }
    finalize_record(S, r, own, target, single);
  }
}
template <bool DP>
__global__ void __launch_bounds__(TPB) k_wide_commit(DevState S, int r, unsigned long long base, int len, int n_blocks, int src) {
  const int own = S.lambda[r];
  wide_commit_block<DP>(S, r, base, len, n_blocks, src, own, S.ent_size[own] == 1, uniforms(S.seed, r, S.iter, PH_LINK_DECIDE, 0, 0), S.Lnew[r]);
}

// The turns of many wide records, one after the other, in ONE persistent kernel: same pieces, same
// sums as the two kernels.  Only the blocks that hold pieces of a record's list take part in its
// turn, and they synchronise among themselves through two words per turn (arrivals after the
// reduction, a flag once block 0 has committed) instead of two barriers over the whole grid: a turn
// over a list of a few thousand entries costs a few microseconds whatever the longest list of the
// stretch is.  Launched cooperatively, so that every block is resident while others spin on it.
struct WideTurn { int r, len, src, pad; unsigned long long base; };
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
template <bool DP>
__global__ void __launch_bounds__(TPB) k_wide_turns(DevState S, const WideTurn* turns, int n, unsigned* __restrict__ sync,
                                                    unsigned long long* __restrict__ dbg) { // dbg (trace, may be null): [n][4] ns of block 0
  unsigned* arrived = sync;      // [n] blocks that have reduced their pieces of turn t
  unsigned* committed = sync + n; // [n] 1: turn t is committed
  int waited = -1;                // last turn this block knows to be committed
  for (int t = 0; t < n; t++) {
    const WideTurn w = turns[t];
    const int n_blocks = max(1, min(WIDE_BLOCKS, (w.len + 2047) / 2048));
    const int taking_part = min(n_blocks, (int)gridDim.x);
    if ((int)blockIdx.x >= taking_part) continue;
    // what does not depend on the turns before this one: the record's own entity (its link moves at
    // its own turn only), the first tile of this warp's piece, and - block 0 - the record's draws
    const int own = S.lambda[w.r];
    WideTile first;
    {
      long long lo, hi;
      wide_piece_of(w.len, n_blocks, (int)blockIdx.x, lo, hi);
      wide_load_tile((w.src ? S.wlb_id : S.wl_id) + w.base, (w.src ? S.wlb_ll : S.wl_ll) + w.base, lo, hi, first);
    }
    U2 u = {0.0, 0.0};
    double lnew = 0.0;
    if (blockIdx.x == 0) { u = uniforms(S.seed, w.r, S.iter, PH_LINK_DECIDE, 0, 0); lnew = S.Lnew[w.r]; }
    if (t > 0 && waited < t - 1) { // sizes and member lists as the previous turn left them
      if (threadIdx.x == 0) while (ld_acquire_u32(&committed[t - 1]) == 0u) {}
      __syncthreads();
      waited = t - 1;
    }
    const bool single = __ldcg(&S.ent_size[own]) == 1;
    if (dbg && blockIdx.x == 0 && threadIdx.x == 0) dbg[4 * t] = global_ns();
    for (int vb = blockIdx.x; vb < n_blocks; vb += gridDim.x)
      wide_reduce_block<DP>(S, w.r, w.base, w.len, w.src, vb, n_blocks, own, single, vb == (int)blockIdx.x ? &first : nullptr);
    if (dbg && blockIdx.x == 0 && threadIdx.x == 0) dbg[4 * t + 1] = global_ns();
    if (taking_part > 1) {
      __syncthreads();
      if (threadIdx.x == 0) red_release_add_u32(&arrived[t], 1u);
    }
    if (blockIdx.x == 0) {
      if (taking_part > 1) {
        if (threadIdx.x == 0) while (ld_acquire_u32(&arrived[t]) < (unsigned)taking_part) {}
      }
      __syncthreads();
      if (dbg && threadIdx.x == 0) dbg[4 * t + 2] = global_ns();
      wide_commit_block<DP>(S, w.r, w.base, w.len, n_blocks, w.src, own, single, u, lnew);
      __syncthreads();
      if (threadIdx.x == 0) red_release_add_u32(&committed[t], 1u);
      if (dbg && threadIdx.x == 0) dbg[4 * t + 3] = global_ns();
      waited = t;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Possible links of the wide records that HAVE a bucket (their candidate scan found more than
// wide_threshold possible links): instead of scanning all 2N items for them once more
// (k_wide_scan), their own bucket / cells are enumerated again and every possible link is emitted as
// (key = batch position << 32 | item, log-likelihood), in any order; one radix sort by key then
// leaves every record's list contiguous and in ascending item order (k_wide_split cuts it up).
// WE_BLOCKS blocks per record share its tiles of bucket items (or its probes).
// ------------------------------------------------------------------------------------------
constexpr int WE_BLOCKS = 16;
constexpr int WE_TILE = 2048;
// A record that its own table does not serve (cand_cnt = -3: no table for its key mask, or too many
// probe combinations for the candidate kernels) is listed through the cheapest of the tables built
// this sweep - chain tables are probed with up to four unpinned key attributes - and `served[i]`
// says whether there was one within emit_cost_cap (else: batched scan over all the items).
__global__ void __launch_bounds__(TPB) k_wide_emit(DevState S, const int* __restrict__ recs, int n, unsigned long long* __restrict__ keys,
                                                   double* __restrict__ vals, unsigned long long* __restrict__ cursor,
                                                   unsigned long long cap, int* __restrict__ served) {
  const int i = blockIdx.x / WE_BLOCKS, jb = blockIdx.x % WE_BLOCKS;
  if (i >= n) return;
  const int r = recs[i];
  const int* x = S.rec_x + (size_t)r * S.RS;
  const uint32_t zm = S.rec_z[r];
  const bool unserved = S.cand_cnt[r] == -3;
  int tab = 0;
  if (!unserved) {
    tab = S.tab_remap[S.rec_tab[r]];
    if (!tab) tab = S.tab_remap[N_TABLES + S.rec_alt[r]];
  } else {
    double best = S.emit_cost_cap;
    for (int k = 0; k < S.n_act_tabs; k++) {
      const double c = big_probe_cost(S, S.tables[S.act_tabs[k]], x, zm);
      if (c < best) { best = c; tab = S.act_tabs[k]; }
    }
  }
  if (jb == 0 && threadIdx.x == 0) served[i] = tab != 0;
  if (!tab) return;
  const DevTable& t = S.tables[tab];
  const BigProbe pr = make_big_probe(S, t.mask, x, zm);
  if (pr.n_un > BIG_UNPINNED) { if (jb == 0 && threadIdx.x == 0) served[i] = 0; return; } // (cannot happen for a record with a stored bucket)
  // one possible link (and its twin) of record i; the threads that get here together share one atomic
  auto emit = [&](int item, double ll, bool twin) {
    cg::coalesced_group g = cg::coalesced_threads();
    const int mine = twin ? 2 : 1;
    const int before = cg::exclusive_scan(g, mine);
    unsigned long long base = 0;
    if (g.thread_rank() == g.size() - 1) base = atomicAdd(cursor, (unsigned long long)(before + mine));
    base = g.shfl(base, g.size() - 1) + (unsigned long long)before;
    if (base + (unsigned long long)mine <= cap) {
      keys[base] = ((unsigned long long)(uint32_t)i << 32) | (unsigned long long)(uint32_t)item; vals[base] = ll;
      if (twin) { keys[base + 1] = ((unsigned long long)(uint32_t)i << 32) | (unsigned long long)(uint32_t)(S.N + item); vals[base + 1] = ll; }
    }
  };
  auto visit = [&](int item, const int* v, const int* y) {
    if (item == S.N + r) return;
    if (!y) y = S.ent_y + (size_t)item * S.ES;
    if (pr.n_un > 0 && !pr.own_probe(y, v)) return;
    if (!item_compatible(S, x, zm, y)) return;
    const double ll = item_loglik(S, x, zm, y);
    if (!(ll > -CUDART_INF)) return;
    emit(item, ll, item < S.N && item != r && S.twin[item]);
  };
  if (pr.n_un == 0 && t.csr) {
    const uint32_t b = unserved ? bucket_of(S, t, x) : S.rec_bkt[r];
    const int lo = b ? t.ptr[b - 1] : 0, hi = t.ptr[b];
    const int none[1] = {0};
    for (int q0 = lo + jb * WE_TILE; q0 < hi; q0 += WE_BLOCKS * WE_TILE)
      for (int q = q0 + threadIdx.x; q < min(hi, q0 + WE_TILE); q += TPB) visit(t.ids[q], none, nullptr);
  } else if (!t.chain && 1.2 * (double)S.N >= 64.0 * (double)t.B) {
    // CSR table with long buckets: the blocks take the probes in turn, all threads stride over the bucket
    for (long long j = jb; j < pr.count; j += WE_BLOCKS) {
      int v[BIG_UNPINNED];
      pr.values(j, v);
      const uint32_t b = pr.bucket(S, t, x, v);
      const int lo = b ? t.ptr[b - 1] : 0, hi = t.ptr[b];
      for (int q = lo + threadIdx.x; q < hi; q += TPB) visit(t.ids[q], v, nullptr);
    }
  } else {
    for (long long j = (long long)jb * TPB + threadIdx.x; j < pr.count; j += WE_BLOCKS * TPB) {
      int v[BIG_UNPINNED];
      pr.values(j, v);
      if (t.chain) {
        for (int item = t.head[pr.cell(S, t, x, v)]; item >= 0;) {
          const ChainNode nd = load_node(&t.node[item]);
          int yv[8];
          for (int a = 0; a < S.A; a++) yv[a] = unpack_attr(S, nd.P, a);
          // (n_un = 0: the cell may hold items of other keys too - the pinned attributes are compared in visit)
          visit(item, v, yv);
          item = nd.next;
        }
      } else {
        const uint32_t b = pr.bucket(S, t, x, v);
        const int lo = b ? t.ptr[b - 1] : 0, hi = t.ptr[b];
        for (int q = lo; q < hi; q++) visit(t.ids[q], v, nullptr);
      }
    }
  }
}
// keys sorted by (batch position, item): lists into wlb_id / wlb_ll, start and length of every record's list
__global__ void __launch_bounds__(TPB) k_wide_split(DevState S, const unsigned long long* __restrict__ keys, const double* __restrict__ vals,
                                                    long long n, int* __restrict__ seg_base, int* __restrict__ seg_len) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long k = keys[i];
  const int pos = (int)(k >> 32);
  S.wlb_id[i] = (int)(uint32_t)k;
  S.wlb_ll[i] = vals[i];
  if (i == 0 || (int)(keys[i - 1] >> 32) != pos) seg_base[pos] = (int)i;
  if (i == n - 1 || (int)(keys[i + 1] >> 32) != pos) seg_len[pos] = (int)i + 1; // end for now: the host subtracts the start
}

// records listed by k_wide_emit that turned out to have at most wide_threshold possible links keep their
// list like everybody else (the counterpart of k_wide_store; one block per record, list sorted already)
__global__ void __launch_bounds__(TPB) k_wide_store_b(DevState S, const int* __restrict__ recs, const int* __restrict__ sel,
                                                      const int* __restrict__ seg_base, const int* __restrict__ seg_end) {
  __shared__ unsigned long long s_off;
  const int i = sel[blockIdx.x], r = recs[i];
  const int base = seg_base[i];
  const int cnt = seg_end[i] > 0 ? seg_end[i] - base : 0;
  const int own = S.lambda[r];
  if (threadIdx.x == 0) s_off = cnt > 0 ? atomicAdd(S.pool_cursor, (unsigned long long)cnt) : 0ull;
  __syncthreads();
  const unsigned long long o = s_off;
  const bool overflow = o + (unsigned long long)cnt > S.pool_cap;
  if (!overflow) for (int k = threadIdx.x; k < cnt; k += TPB) {
    const int id = S.wlb_id[base + k];
    S.cand_id[o + k] = id; S.cand_ll[o + k] = S.wlb_ll[base + k];
    if (id != own) S.referenced[id] = 1;
  }
  if (threadIdx.x == 0) {
    const bool other = overflow || cnt > 1 || (cnt == 1 && S.wlb_id[base] != own);
    S.cand_off[r] = overflow ? 0u : (uint32_t)o;
    S.cand_cnt[r] = overflow ? -2 : cnt;
    S.dlo[r] = other ? -1 : 0;
    S.dhi[r] = 1;
    if (overflow) atomicAdd(&S.counters[C_NFENCE], 1);
    else {
      atomicAdd((unsigned long long*)&S.counters64[0], (unsigned long long)cnt);
      if (cnt > S.long_min) S.actl_in[atomicAdd(&S.counters[C_NACTL_INIT], 1)] = r;
    }
  }
}

// Upper bound of an unfinalised record's change of K: it can only add an entity (+1) if it shares
// its entity at its turn, i.e. if the entity has other members now or some other record may join
// it first (the entity is on somebody's candidate list, or a record without a list is pending).
// "Round 0" (sure_ok: prior_weight_new is positive and finite for every possible K): a record that
// is alone in its entity, has no candidate but that entity, and whose entity and potential entity
// are on nobody's stored list founds its entity whatever happens before its turn - nothing it reads
// can change and nothing it writes is read by anybody else - so it is final here and now, without
// reservations (seven records in eight at 10 M).  Same outcome as its turn in round 1.
__global__ void __launch_bounds__(TPB) k_bounds_init(DevState S, int any_nolist, int sure_ok) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= S.N || S.fin[r] == 1) return;
  const int own = S.lambda[r];
  const int sz = S.ent_size[own];
  const bool shared = any_nolist || sz > 1 || S.referenced[own];
  if (S.use_kbounds) S.dhi[r] = shared ? 1 : 0;
  if (sure_ok && !shared && sz == 1 && S.cand_off[r] == CAND_OWN_ONLY && S.cand_cnt[r] == 1 && !S.referenced[S.N + r]) {
    const double Ln = S.Lnew[r];
    if (Ln > -CUDART_INF && Ln < CUDART_INF) finalize_record(S, r, own, S.N + r, true);
  }
}

// wide records are final before the rounds start and their change of K is folded into K0
__global__ void __launch_bounds__(TPB) k_clear_bounds(DevState S, const int* __restrict__ list, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { S.dlo[list[i]] = 0; S.dhi[list[i]] = 0; }
}

// drop finalised records from an active list
__global__ void __launch_bounds__(TPB) k_compact_active(DevState S, const int* __restrict__ act, int n_act) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = i < n_act ? (act ? act[i] : i) : -1;
  const bool stay = r >= 0 && S.fin[r] != 1;
  const unsigned m = __ballot_sync(FULL, stay);
  if (m) {
    int base = 0;
    const int lane = threadIdx.x & 31;
    if (lane == (__ffs(m) - 1)) base = atomicAdd(&S.counters[C_NACT_OUT], __popc(m));
    base = __shfl_sync(FULL, base, __ffs(m) - 1);
    if (stay) S.act_out[base + __popc(m & ((1u << lane) - 1u))] = r;
  }
}


// ------------------------------------------------------------------------------------------
// K1e  rename every entity to its smallest member record (the head of its sorted member list).
// Slots [0,N) of the "next" arrays become the entity table of the following phases.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_canon(DevState S) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  int rep = 0;
  if (r < S.N) {
    const int e = S.lambda[r];
    const int c = S.ent_head[e];
    S.lambda[r] = c;
    rep = c == r;
    S.ent_size_nx[r] = rep ? S.ent_size[e] : 0;
    S.ent_head_nx[r] = rep ? r : -1;
  }
  rep = warp_sum(rep);
  if ((threadIdx.x & 31) == 0 && rep) atomicAdd(&S.counters[C_K], rep);
}

// ------------------------------------------------------------------------------------------
// K2  entity attribute resampling (replaces Entities::update_attributes, logWeightEntityValue,
// drawEntityValueSequential and drawEntityValueIndex, src/entities.cpp:236-469).
// One thread per entity slot: >= 65% of the entities are singletons whose only member is the
// record with the same index, so member reads are coalesced; larger clusters walk their
// sorted member list.  Persistent distributions come from device alias tables.
// ------------------------------------------------------------------------------------------
// `row_first` / `row_j`: when v was taken from position row_j of the neighbourhood row of value
// row_first, log p(row_first | v) is that row entry - no search for members with this value
template <bool DP>
__device__ double log_weight_entity_value(const DevState& S, const DevAttr& t, int a, int v, int head,
                                          int row_first = -1, int row_j = -1) {
  double lw = t.logphi[v];
  const double b = S.conc[a];
  int ctr_disagree = 0;
  for (int r = head; r >= 0; r = S.rec_next[r]) {
    const int* x = S.rec_x + (size_t)r * S.RS;
    const int xv = x[a];
    if (xv < 0) continue;
    const int f = x[S.RS - 1];
    if (xv == v) lw += t.lt_same[f * t.D + v];
    else {
      lw += t.lt_diff[f * t.D + v];
      if (!DP || isinf(b)) lw += (xv == row_first && row_j >= 0) ? t.logp[row_j] : log_distortion_prob(t, v, xv);
      else { // entities.cpp:275-288: Chinese-restaurant terms of the disagreeing values
        ctr_disagree++;
        lw += -log((double)ctr_disagree - 1.0 + b);
        int seen = 0; // earlier members with the same disagreeing value
        for (int q = head; q != r; q = S.rec_next[q]) seen += S.rec_x[(size_t)q * S.RS + a] == xv;
        const double bp = b * distortion_prob(t, v, xv);
        lw += seen ? log((double)seen + bp) : log(bp);
      }
    }
  }
  return lw;
}

// `xs`: the record rows (8 ints each) of the entity's `n_xs` members in ascending id, read already (a
// singleton's only row, the few rows of a small cluster); nullptr: the members are read from the
// sorted member list starting at `head`
__device__ __forceinline__ int draw_entity_value(const DevState& S, int e, int a, int head, const int* xs = nullptr, int n_xs = 1) {
  // fn(row) for every member row in ascending record id
  auto each = [&](auto&& fn) {
    if (xs) for (int k = 0; k < n_xs; k++) fn(xs + 8 * k);
    else for (int r = head; r >= 0; r = S.rec_next[r]) fn(S.rec_x + (size_t)r * S.RS);
  };
  const DevAttr& t = S.attrs[a];
  int nobs = 0, first = -1;
  each([&](const int* x) {
    const int xv = x[a];
    if (xv >= 0) { if (first < 0) first = xv; nobs++; }
  });
  const U2 u0 = uniforms(S.seed, e, S.iter, PH_ENT_ATTR, a, 0);
  if (nobs == 0) return alias_draw(t.aprob, t.aalias, t.D, u0.a); // entities.cpp:387-389

  if (t.is_const && t.exclude && nobs <= NMAX_FAST && (nobs == 1 || !((S.dp_mask >> a) & 1u))) {
    // whole-domain weights of drawEntityValueSequential (entities.cpp:298-348) in closed form:
    // explicit for the members' values, C phi(v)/(1-psi(v))^nobs for every other value
    int vals[NMAX_FAST];
    double w[NMAX_FAST];
    int nv = 0;
    double Cc = 1.0;
    each([&](const int* x) {
      const int xv = x[a];
      if (xv < 0) return;
      Cc *= S.theta[x[S.RS - 1] * S.A + a] * t.psi[xv];
      bool seen = false;
      for (int j = 0; j < nv; j++) seen |= vals[j] == xv;
      if (!seen) vals[nv++] = xv;
    });
    const double* g = t.g + (size_t)nobs * t.D;
    double sumw = 0.0, rest = t.G[nobs];
    for (int j = 0; j < nv; j++) {
      const int v = vals[j];
      double ww = t.phi[v];
      each([&](const int* x) {
        const int xv = x[a];
        if (xv < 0) return;
        const double th = S.theta[x[S.RS - 1] * S.A + a];
        ww *= xv == v ? 1.0 - th : th * t.psi[xv] / t.om[v];
      });
      w[j] = ww;
      sumw += ww;
      rest -= g[v];
    }
    if (rest < 0.0) rest = 0.0;
    const double wrest = Cc * rest, total = sumw + wrest;
    if (!(total > 0.0) || !isfinite(total)) { dev_fail(S, EXB200_ERR_WEIGHTS, e); return first; }
    const double tt = u0.a * total;
    double c = 0.0;
    for (int j = 0; j < nv; j++) { c += w[j]; if (tt < c) return vals[j]; }
    if (!(wrest > 0.0))
      for (int j = nv - 1; j >= 0; j--) if (w[j] > 0.0) return vals[j];
    const double* ap = t.aprob + (size_t)nobs * t.D;
    const int* aa = t.aalias + (size_t)nobs * t.D;
    for (int k = 1; k < REJECT_CAP; k++) {
      const U2 u = uniforms(S.seed, e, S.iter, PH_ENT_ATTR, a, k);
      const int v = alias_draw(ap, aa, t.D, u.a);
      bool hit = false;
      for (int j = 0; j < nv; j++) hit |= vals[j] == v;
      if (!hit) return v;
    }
    dev_fail(S, EXB200_ERR_WEIGHTS, e);
    return first;
  }

  bool same = true; // every observed member value equals the first one
  if (nobs == 2)
    each([&](const int* x) {
      const int xv = x[a];
      same &= xv < 0 || xv == first;
    });
  if (!t.is_const && t.exclude && (nobs == 1 || (nobs == 2 && same && !((S.dp_mask >> a) & 1u)))) {
    // one observed member value x, held by one member (singletons above all) or by two: every
    // v != x has weight (prod_m theta_m) phi(v) (mef(v) p(x|v))^nobs whose row sums are static, so
    // drawEntityValueIndex (entities.cpp:351-431) reduces to "keep x" versus a search in the
    // row's prefix sums
    double th = 1.0, wx = t.phi[first];
    each([&](const int* x) {
      if (x[a] >= 0) {
        const double thm = S.theta[x[S.RS - 1] * S.A + a];
        th *= thm;
        wx *= 1.0 - thm * t.mef[first];
      }
    });
    const double* cs = nobs == 1 ? t.cs : t.cs2;
    const int lo = t.row[first], hi = t.row[first + 1];
    const double rsum = hi > lo ? cs[hi - 1] : 0.0;
    const double total = wx + th * rsum;
    if (!(total > 0.0) || !isfinite(total)) { dev_fail(S, EXB200_ERR_WEIGHTS, e); return first; }
    const double tt = u0.a * total;
    if (tt < wx || hi == lo) return first;
    const double t2 = tt - wx;
    int a0 = lo, b0 = hi; // first j with t2 < th * cs[j]
    while (a0 < b0) { const int mid = (a0 + b0) >> 1; if (t2 < th * cs[mid]) b0 = mid; else a0 = mid + 1; }
    if (a0 == hi) { // rounding: the first entry that reaches the row total
      a0 = lo; b0 = hi - 1;
      while (a0 < b0) { const int mid = (a0 + b0) >> 1; if (cs[mid] >= rsum) b0 = mid; else a0 = mid + 1; }
    }
    return t.col[a0];
  }

  return -1; // general scheme: left to the warp-cooperative kernel below
}

// General scheme, one WARP per (entity, attribute): explicit support - the whole domain for a
// constant attribute (entities.cpp:315-329), else the first observed member value followed by its
// neighbourhood row (a subset of entities.cpp:361-383 that holds all the mass).  Lanes take
// consecutive support entries; weights exp(lw - max) are summed per chunk of 32 with the warp's
// shuffle scan and chunk totals sequentially - the oracle uses the same association order.
// logWeightEntityValue for a warp that holds the entity's members in its lanes (lane k: value and
// file of the k-th member, n_mem <= 32; infinite concentration): the terms of
// log_weight_entity_value in the same order, without walking the member list again for every
// candidate value.  Called by all 32 lanes.
__device__ __forceinline__ double log_weight_members(const DevAttr& t, int v, bool valid, int n_mem, int mx, int mf,
                                                     int row_first, int row_j) {
  double lw = valid ? t.logphi[v] : -CUDART_INF;
  for (int k = 0; k < n_mem; k++) {
    const int xv = __shfl_sync(FULL, mx, k), f = __shfl_sync(FULL, mf, k);
    if (xv < 0 || !valid) continue;
    if (xv == v) lw += t.lt_same[f * t.D + v];
    else {
      lw += t.lt_diff[f * t.D + v];
      lw += (xv == row_first && row_j >= 0) ? t.logp[row_j] : log_distortion_prob(t, v, xv);
    }
  }
  return lw;
}

template <bool DP>
__device__ int draw_entity_value_warp(const DevState& S, int e, int a, int head) {
  const DevAttr& t = S.attrs[a];
  const int lane = threadIdx.x & 31;
  // members into lanes (every lane walks the list: the loads are broadcasts)
  int first = -1, n_mem = 0, mx = -1, mf = 0;
  for (int r = head; r >= 0; r = S.rec_next[r], n_mem++) {
    if ((n_mem & 31) == lane || first < 0) {
      const int* x = S.rec_x + (size_t)r * S.RS;
      const int xv = x[a];
      if (first < 0 && xv >= 0) first = xv;
      if ((n_mem & 31) == lane) { mx = xv; mf = x[S.RS - 1]; }
    }
  }
  const bool in_lanes = n_mem <= 32 && !(DP && isfinite(S.conc[a])); // a finite concentration needs the urn terms: list walk
  const int lo = t.is_const ? 0 : t.row[first];
  const int ns = t.is_const ? t.D : 1 + (t.row[first + 1] - lo);
  const int nchunk = (ns + 31) >> 5;
  auto weight_log = [&](int j) -> double { // called by all lanes
    const bool valid = j < ns;
    int v = 0;
    if (valid) v = t.is_const ? j : (j == 0 ? first : t.col[lo + j - 1]);
    const bool dup = valid && !t.is_const && j > 0 && v == first;
    const int rf = t.is_const ? -1 : first, rj = (t.is_const || j == 0) ? -1 : lo + j - 1;
    if (in_lanes) {
      return log_weight_members(t, v, valid && !dup, n_mem, mx, mf, rf, rj);
    }
    if (!valid || dup) return -CUDART_INF;
    return log_weight_entity_value<DP>(S, t, a, v, head, rf, rj);
  };
  // the log-weights of the first chunks stay in registers for the later passes
  constexpr int NC = 8;
  double lwc[NC];
  double m = -CUDART_INF;
  for (int c = 0; c < nchunk; c++) {
    const double lw = weight_log(32 * c + lane);
    if (c < NC) lwc[c] = lw;
    m = lw > m ? lw : m;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const double y = __shfl_xor_sync(FULL, m, o); m = y > m ? y : m; }
  if (!(m > -CUDART_INF)) { if (lane == 0) dev_fail(S, EXB200_ERR_WEIGHTS, e); return first; }
  auto chunk_lw = [&](int c) -> double {
    double lw = 0.0;
    bool cached = false;
#pragma unroll
    for (int k = 0; k < NC; k++) if (c == k) { lw = lwc[k]; cached = true; }
    return cached ? lw : weight_log(32 * c + lane);
  };
  // chunk totals (lane c % 32 keeps the total of chunk c), summed in chunk order
  double total = 0.0, tcs = 0.0;
  int lastpos = -1;
  for (int c = 0; c < nchunk; c++) {
    const double lw = chunk_lw(c);
    const double w = lw > -CUDART_INF ? exp(lw - m) : 0.0;
    double x = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const double y = __shfl_up_sync(FULL, x, o); if (lane >= o) x = x + y; }
    const double tc = __shfl_sync(FULL, x, 31);
    if ((c & 31) == lane && c < 32) tcs = tc;
    total += tc;
    const unsigned pos = __ballot_sync(FULL, w > 0.0);
    if (pos) lastpos = 32 * c + (31 - __clz(pos));
  }
  const U2 u0 = uniforms(S.seed, e, S.iter, PH_ENT_ATTR, a, 0);
  const double tt = u0.a * total;
  double run = 0.0;
  for (int c = 0; c < nchunk; c++) {
    double tc;
    bool look = true;
    if (c < 32) { tc = __shfl_sync(FULL, tcs, c); look = tt < run + tc; }
    if (look) { // the chunk that holds the draw (beyond 32 chunks: every chunk is looked at again)
      const int j = 32 * c + lane;
      const double lw = chunk_lw(c);
      const double w = lw > -CUDART_INF ? exp(lw - m) : 0.0;
      double x = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const double y = __shfl_up_sync(FULL, x, o); if (lane >= o) x = x + y; }
      tc = __shfl_sync(FULL, x, 31);
      if (tt < run + tc) {
        const unsigned hit = __ballot_sync(FULL, j < ns && tt < run + x);
        if (hit) {
          const int jj = 32 * c + (__ffs(hit) - 1);
          return t.is_const ? jj : (jj == 0 ? first : t.col[lo + jj - 1]);
        }
      }
    }
    run += tc;
  }
  if (lastpos < 0) return first;
  return t.is_const ? lastpos : (lastpos == 0 ? first : t.col[lo + lastpos - 1]);
}

// (entity slots [e_lo, e_hi): the whole range on one GPU, a slice under within-chain sharding)
// gen_list holds two lists of (entity << 8 | attribute) pairs: from the front (gen_count[0]) the pairs
// left to the general scheme (k_entities_general), from the back (gen_count[1], last entry at
// gen_cap - 1) the draws of singletons that did not keep their member's value (k_entities_slow).
__global__ void __launch_bounds__(TPB) k_entities(DevState S, unsigned long long* __restrict__ gen_list, int* __restrict__ gen_count,
                                                  long long gen_cap, int e_lo, int e_hi, int* __restrict__ multi_list) {
  __shared__ int s_warp[TPB / 32]; // (launched with blocks of up to TPB threads)
  __shared__ int s_base;
  const int e = e_lo + blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t need = 0, slow = 0;
  bool deferred = false;
  const int sz = e < e_hi ? S.ent_size[e] : 0;
  if (sz == 1 && S.RS == 8 && S.ES == 8) {
    // a singleton: its only member is record e (entities are named after their smallest member), whose
    // row is read once - coalesced - and whose new attribute row is written once
    int x[8], yn[8];
    {
      const int4* xp = reinterpret_cast<const int4*>(S.rec_x + (size_t)e * 8);
      const int4 x0 = xp[0], x1 = xp[1];
      x[0] = x0.x; x[1] = x0.y; x[2] = x0.z; x[3] = x0.w; x[4] = x1.x; x[5] = x1.y; x[6] = x1.z; x[7] = x1.w;
    }
#pragma unroll
    for (int a = 0; a < 8; a++) yn[a] = 0;
    for (int a = 0; a < S.A; a++) {
      // "keep the member's value" straight from the per-sweep table (nineteen draws in twenty).  The
      // others are listed for k_entities_slow - one thread per draw through the general function,
      // which recomputes the same weights from the same uniform - instead of being made here by one
      // lane or two while the rest of the warp waits.
      const DevAttr& t = S.attrs[a];
      const int xv = x[a];
      int v = -2;
      if (xv >= 0 && t.exclude) {
        const double2 kw = *reinterpret_cast<const double2*>(t.sg + 2 * ((size_t)x[7] * t.D + xv));
        const double total = kw.x + kw.y;
        if (total > 0.0 && isfinite(total)) {
          const U2 u0 = uniforms(S.seed, e, S.iter, PH_ENT_ATTR, a, 0);
          if (u0.a * total < kw.x) v = xv;
        }
      }
      if (v >= 0) yn[a] = v; else slow |= 1u << a;
    }
    int* y = S.ent_y + (size_t)e * 8;
    if (slow) { for (int a = 0; a < S.A; a++) if (!((slow >> a) & 1u)) y[a] = yn[a]; }
    else {
      int4* yp = reinterpret_cast<int4*>(y);
      yp[0] = make_int4(yn[0], yn[1], yn[2], yn[3]);
      yp[1] = make_int4(yn[4], yn[5], yn[6], yn[7]);
    }
  } else if (sz > 0 && multi_list) {
    // more than one member: listed for k_entities_multi, so that a warp of singletons does not wait
    // for the lane that walks a member list
    deferred = true;
  } else if (sz > 0) {
    const int head = S.ent_head[e];
    int* y = S.ent_y + (size_t)e * S.ES;
    for (int a = 0; a < S.A; a++) {
      const int v = draw_entity_value(S, e, a, head);
      if (v >= 0) y[a] = v; else need |= 1u << a;
    }
  }
  {
    const unsigned dm = __ballot_sync(FULL, deferred);
    if (dm) {
      const int ln = threadIdx.x & 31;
      int base = 0;
      if (ln == __ffs(dm) - 1) base = atomicAdd(gen_count + 2, __popc(dm));
      base = __shfl_sync(FULL, base, __ffs(dm) - 1);
      if (deferred) multi_list[base + __popc(dm & ((1u << ln) - 1u))] = e;
    }
  }
  // both kinds of pairs are listed with one atomic per block and list
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int which = 0; which < 2; which++) {
    const uint32_t mask = which ? slow : need;
    const int mine = __popc(mask);
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += y; }
    __syncthreads(); // (s_warp / s_base of the first list have been read)
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int k = 0; k < (int)blockDim.x / 32; k++) { const int c = s_warp[k]; s_warp[k] = t; t += c; }
      s_base = t ? atomicAdd(gen_count + which, t) : 0;
    }
    __syncthreads();
    long long pos = (long long)s_base + s_warp[w] + incl - mine;
    for (uint32_t m = mask; m; m &= m - 1) {
      const unsigned long long pr = ((unsigned long long)(uint32_t)e << 8) | (unsigned long long)(__ffs(m) - 1);
      gen_list[which ? gen_cap - 1 - pos : pos] = pr;
      pos++;
    }
  }
}

// the entities with several members that k_entities listed (gen_count[2] of them): one thread per
// entity; the rows of up to ENT_ROWS members are read once and kept for all the attributes
constexpr int ENT_ROWS = 4;
__global__ void __launch_bounds__(128) k_entities_multi(DevState S, const int* __restrict__ multi_list, unsigned long long* __restrict__ gen_list,
                                                        int* __restrict__ gen_count) {
  const int n = gen_count[2];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int e = multi_list[i];
    const int head = S.ent_head[e], sz = S.ent_size[e];
    int xm[ENT_ROWS * 8];
    const bool rows = sz <= ENT_ROWS && S.RS == 8;
    if (rows) {
      int k = 0;
      for (int r = head; r >= 0 && k < ENT_ROWS; r = S.rec_next[r], k++) {
        const int4* xp = reinterpret_cast<const int4*>(S.rec_x + (size_t)r * 8);
        const int4 x0 = xp[0], x1 = xp[1];
        int* x = xm + 8 * k;
        x[0] = x0.x; x[1] = x0.y; x[2] = x0.z; x[3] = x0.w; x[4] = x1.x; x[5] = x1.y; x[6] = x1.z; x[7] = x1.w;
      }
    }
    int* y = S.ent_y + (size_t)e * S.ES;
    for (int a = 0; a < S.A; a++) {
      const int v = rows ? draw_entity_value(S, e, a, head, xm, sz) : draw_entity_value(S, e, a, head);
      if (v >= 0) y[a] = v;
      else gen_list[atomicAdd(gen_count, 1)] = ((unsigned long long)(uint32_t)e << 8) | (unsigned long long)a;
    }
  }
}

// the singleton draws k_entities listed: one thread per (entity, attribute) pair
__global__ void __launch_bounds__(TPB) k_entities_slow(DevState S, unsigned long long* __restrict__ gen_list, int* __restrict__ gen_count,
                                                       long long gen_cap) {
  const int n = gen_count[1];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned long long p = gen_list[gen_cap - 1 - i];
    const int e = (int)(p >> 8), a = (int)(p & 0xFFu);
    int x[8];
    {
      const int4* xp = reinterpret_cast<const int4*>(S.rec_x + (size_t)e * 8);
      const int4 x0 = xp[0], x1 = xp[1];
      x[0] = x0.x; x[1] = x0.y; x[2] = x0.z; x[3] = x0.w; x[4] = x1.x; x[5] = x1.y; x[6] = x1.z; x[7] = x1.w;
    }
    const int v = draw_entity_value(S, e, a, e, x);
    if (v >= 0) S.ent_y[(size_t)e * 8 + a] = v;
    else gen_list[gen_cap - 1 - i] = p | (1ull << 63); // the general scheme after all: flagged where it stands
  }
}

// one warp per listed (entity, attribute) pair, whatever the launch shape
template <bool DP>
__global__ void __launch_bounds__(TPB) k_entities_general(DevState S, const unsigned long long* __restrict__ gen_list,
                                                          const int* __restrict__ gen_count, long long gen_cap) {
  const int lane = threadIdx.x & 31;
  const int n = gen_count[0], n_back = gen_count[1]; // the front list, and the flagged entries of the back list
  const int n_warps = gridDim.x * (TPB / 32);
  for (int i = blockIdx.x * (TPB / 32) + (threadIdx.x >> 5); i < n + n_back; i += n_warps) {
    unsigned long long p = i < n ? gen_list[i] : gen_list[gen_cap - 1 - (i - n)];
    if (i >= n) { if (!(p >> 63)) continue; p &= ~(1ull << 63); }
    const int e = (int)(p >> 8), a = (int)(p & 0xFFu);
    const int v = draw_entity_value_warp<DP>(S, e, a, S.ent_head[e]);
    if (lane == 0) S.ent_y[(size_t)e * S.ES + a] = v;
  }
}

// ------------------------------------------------------------------------------------------
// Auxiliary variables of DistortDistConcs::update (src/distort_dist_concs.cpp:56-88) for attribute a:
// per entity the number of disagreeing observed members (for the host's Beta(b, n) variate) and
// the Bernoulli sum zeta, with the value counts kept per cluster.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_conc_aux(DevState S, int a, int* __restrict__ n_disagree,
                                                  unsigned long long* __restrict__ sum_zeta) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  int zeta = 0;
  if (e < S.N) {
    int ctr = 0;
    if (S.ent_size[e] > 0) {
      const DevAttr& t = S.attrs[a];
      const int head = S.ent_head[e], yv = S.ent_y[(size_t)e * S.ES + a];
      int idx = 0;
      for (int r = head; r >= 0; r = S.rec_next[r], idx++) {
        const int xv = S.rec_x[(size_t)r * S.RS + a];
        if (xv < 0 || xv == yv) continue;
        int seen = 0;
        for (int q = head; q != r; q = S.rec_next[q]) seen += S.rec_x[(size_t)q * S.RS + a] == xv;
        double prob = 1.0;
        if (seen) { const double zz = S.conc[a] * distortion_prob(t, yv, xv); prob = zz / (zz + (double)seen); }
        const U2 u = uniforms(S.seed, e, S.iter, PH_CONC_ZETA, a, idx);
        zeta += u.a < prob;
        ctr++;
      }
    }
    n_disagree[e] = ctr;
  }
  zeta = warp_sum(zeta);
  if ((threadIdx.x & 31) == 0 && zeta) atomicAdd(sum_zeta, (unsigned long long)zeta);
}

// histogram of the entity values of attribute a (Entities::update_distributions, entities.cpp:140-148)
__global__ void __launch_bounds__(TPB) k_ent_hist(DevState S, int a, int* __restrict__ hist) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < S.N && S.ent_size[e] > 0) atomicAdd(&hist[S.ent_y[(size_t)e * S.ES + a]], 1);
}

// ------------------------------------------------------------------------------------------
// K3  distortion indicators (replaces Records::update_distortion, src/records.cpp:70-138):
// elementwise over records x attributes, counts reduced per (file, attribute) with warp
// ballots into shared memory and one global atomic per block and bin.
// ------------------------------------------------------------------------------------------
// (records [r_lo, r_hi): the whole range on one GPU, a slice under within-chain sharding)
__global__ void __launch_bounds__(TPB) k_distort(DevState S, int use_smem, int r_lo, int r_hi) {
  extern __shared__ int s_cnt[]; // [2][F*A] when use_smem
  const int FA = S.F * S.A;
  if (use_smem) {
    for (int i = threadIdx.x; i < 2 * FA; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
  }
  const int r = r_lo + blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = r < r_hi;
  const int* x = S.rec_x + (size_t)(live ? r : 0) * S.RS;
  const int f = live ? x[S.RS - 1] : -1;
  const int* y = S.ent_y + (size_t)(live ? S.lambda[r] : 0) * S.ES;
  const unsigned peers = __match_any_sync(FULL, f);
  const bool leader = live && (threadIdx.x & 31) == (__ffs(peers) - 1);
  uint32_t zm = 0;
  for (int a = 0; a < S.A; a++) {
    bool zz = false, cor = false;
    if (live) {
      const DevAttr& t = S.attrs[a];
      const int xv = x[a], yv = y[a];
      const double th = S.theta[f * S.A + a];
      const bool need_u = xv < 0 || (xv == yv && !t.exclude) || t.mef[yv] != 1.0;
      U2 u;
      u.a = 0.0; u.b = 0.0;
      if (need_u) u = uniforms(S.seed, r, S.iter, PH_DISTORT, a, 0);
      if (xv < 0) zz = u.a < th;                       // records.cpp:95-100
      else if (xv == yv) {
        if (!t.exclude) {                               // records.cpp:109-112
          const double pr1 = th * (t.is_const ? t.psi[xv] : 1.0), pr0 = 1.0 - th;
          zz = u.a < pr1 / (pr1 + pr0);
        }
      } else zz = true;                                 // records.cpp:114-116
      double pr1 = t.mef[yv];                           // records.cpp:122-133
      bool alt;
      if (pr1 == 1.0) alt = true;
      else {
        const double pr0 = 1.0 - pr1;
        pr1 *= zz ? th : 1.0 - th;
        alt = u.b < pr1 / (pr1 + pr0);
      }
      cor = alt && !zz;                                 // records.cpp:134
      zm |= (uint32_t)zz << a;
    }
    const unsigned bz = __ballot_sync(FULL, zz), bc = __ballot_sync(FULL, cor);
    if (leader) {
      const int nz = __popc(bz & peers), nc = __popc(bc & peers);
      if (use_smem) {
        if (nz) atomicAdd(&s_cnt[f * S.A + a], nz);
        if (nc) atomicAdd(&s_cnt[FA + f * S.A + a], nc);
      } else {
        if (nz) atomicAdd(&S.ndist[f * S.A + a], nz);
        if (nc) atomicAdd(&S.corr[f * S.A + a], nc);
      }
    }
  }
  if (live) S.rec_z[r] = zm;
  if (use_smem) {
    __syncthreads();
    for (int i = threadIdx.x; i < FA; i += blockDim.x) {
      if (s_cnt[i]) atomicAdd(&S.ndist[i], s_cnt[i]);
      if (s_cnt[FA + i]) atomicAdd(&S.corr[i], s_cnt[FA + i]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Bernoulli sums of ClustParams::update (src/clust_params.cpp:52-58, 83-90, 108-114); the scalar
// Beta / Gamma / NegBinom draws that consume them stay on the host.
// out[0] += #{ i in [1, K-1] : u_i < alpha / (alpha + d i) }
// out[1] += sum over clusters, j in [1, size-1]: PY: u >= (j-1)/(j-d);  coupon: u < kappa/(kappa+j)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_clust_counts(DevState S, int K, int do_u, int do_v, unsigned long long* out) {
  int cu = 0, cv = 0;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (do_u && i >= 1 && i <= K - 1) {
    const U2 u = uniforms(S.seed, i, S.iter, PH_CLUST_U, 0, 0);
    cu = u.a < S.cl.alpha / (S.cl.alpha + S.cl.d * (double)i);
  }
  if (do_v && i < S.N) {
    const int sz = S.ent_size[i];
    for (int j = 1; j <= sz - 1; j++) {
      const U2 u = uniforms(S.seed, i, S.iter, PH_CLUST_V, 0, j);
      if (S.cl.kind == EXB200_CLUST_PITMAN_YOR) cv += u.a >= ((double)j - 1.0) / ((double)j - S.cl.d);
      else cv += u.a < S.cl.kappa / (S.cl.kappa + (double)j);
    }
  }
  cu = warp_sum(cu);
  cv = warp_sum(cv);
  if ((threadIdx.x & 31) == 0) {
    if (cu) atomicAdd(&out[0], (unsigned long long)cu);
    if (cv) atomicAdd(&out[1], (unsigned long long)cv);
  }
}

} // namespace exb
