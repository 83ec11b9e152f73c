/*
 * exoracle.c - CPU restatement of exchanger's partially-collapsed Gibbs sweep.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under exchanger_b200/ may include, link or call
 * this file; it is the checker the CUDA sampler is compared with (tests/,
 * __graft_entry__.smoke(), bench.py's cpu_baseline leg).
 *
 * What it restates (reference = /root/reference, cleanzr/exchanger v0.3.0):
 *   sweep order ............ src/state.h:60-75
 *   links update ........... src/links.cpp:90-212 (candidates), :216-288, :316-347
 *                            (likelihoods), :350-475 (update_link), :477-489 (loop)
 *   entity attributes ...... src/entities.cpp:236-296, :298-348, :351-431, :434-469
 *   distortion indicators .. src/records.cpp:70-138
 *   attribute tables ....... src/attribute_index.cpp:26-116
 *   cluster prior .......... src/clust_params.cpp:4-31, :37-72, :74-121
 *   distortion probs ....... src/distortion_probs.cpp:30-47
 *   DP concentrations ...... src/distort_dist_concs.cpp:32-94 (per-cluster value counts; the
 *                            reference's map that is never cleared between clusters, :53 vs :56,
 *                            is NOT reproduced - SURVEY.md Appendix B.7)
 *   entity distributions ... src/entities.cpp:133-156, src/dirichlet.h:12-34
 *
 * The reference draws from R's Mersenne-Twister in program order and iterates
 * std::unordered_* containers, so its draws are not reproducible by any parallel
 * program.  This restatement keeps every conditional distribution of the reference but
 * fixes what the reference leaves to chance, so that integer outputs are comparable
 * bit for bit (the same rules are implemented by the CUDA sampler; DESIGN.md "sampling
 * specification"):
 *   - randomness is Philox4x32-10 keyed by (seed; id, iteration, phase, attribute, slot);
 *   - records are visited in ascending id (as the reference does, links.cpp:485-488), except
 *     that records with more than `wide_threshold` possible links (compatible items of non-zero
 *     likelihood, counted over all entity slots of the sweep, live or not) take their turn first,
 *     in ascending id: a fixed systematic-scan order that depends only on quantities the links
 *     update does not change.  Candidate entities and entity members in ascending id;
 *   - one-shot discrete draws use the inverse CDF in that order instead of a throw-away
 *     alias table (links.cpp:421-422, entities.cpp:346-347); persistent distributions
 *     (phi, psi, phi/(1-psi)^n) use alias tables (alias_sampler.cpp:3-49);
 *   - a new entity founded by record r during sweep t is called N + r until the end of
 *     the links update, then every entity is renamed to its smallest member record;
 *   - categorical attributes with exclude_entity_value use the algebraically equal
 *     "explicit values + remainder" form of the domain-wide weights.
 * Parity with the reference itself is pinned through oracle/_ref (the reference's own
 * sources compiled against stub headers): per-record log-weights agree to rounding on
 * identical states (tests/test_oracle_vs_reference.py) and posterior summaries agree
 * within Monte-Carlo error.
 */
#include "../include/exb200.h"

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define NMAX_FAST 8 /* categorical fast scheme covers up to this many observed members */
#define REJECT_CAP 100000

enum { PH_LINK_DECIDE = 1, PH_LINK_NEWATTR = 2, PH_ENT_ATTR = 3, PH_DISTORT = 4, PH_CLUST_U = 5,
       PH_CLUST_V = 6, PH_HOST = 7, PH_CONC_ZETA = 8, PH_CONC_BETA = 9 };

/* ------------------------------------------------------------------ Philox4x32-10 */
static void philox4x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                       uint32_t out[4]) {
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static double u52(uint32_t hi, uint32_t lo) {
  uint64_t x = ((uint64_t)(hi >> 6) << 26) | (uint64_t)(lo >> 6);
  return ((double)x + 0.5) * (1.0 / 4503599627370496.0);
}
static void rng2(uint64_t seed, uint32_t id, uint32_t iter, uint32_t phase, uint32_t attr, uint32_t slot,
                 double *ua, double *ub) {
  uint32_t o[4];
  philox4x32((uint32_t)seed, (uint32_t)(seed >> 32), id, iter, (phase << 16) | (attr & 0xFFFFu), slot, o);
  *ua = u52(o[0], o[1]);
  if (ub) *ub = u52(o[2], o[3]);
}
void exo_philox(uint32_t k0, uint32_t k1, const uint32_t c[4], uint32_t out[4]) {
  philox4x32(k0, k1, c[0], c[1], c[2], c[3], out);
}
void exo_uniforms(uint64_t seed, uint32_t id, uint32_t iter, uint32_t phase, uint32_t attr, uint32_t slot,
                  double out[2]) {
  rng2(seed, id, iter, phase, attr, slot, &out[0], &out[1]);
}

/* ------------------------------------------------------------------ alias tables */
typedef struct { int n; double *prob; int *alias; } alias_t;

/* Vose's construction; the role of AliasSampler::build_tables (alias_sampler.cpp:3-42). */
static int alias_build(alias_t *t, const double *w, int n) {
  double total = 0.0;
  for (int i = 0; i < n; i++) {
    if (!(w[i] >= 0.0) || !isfinite(w[i])) return -1; /* alias_sampler.h:76-90 */
    total += w[i];
  }
  if (!(total > 0.0) || !isfinite(total)) return -1;  /* alias_sampler.h:91-95 */
  t->n = n;
  t->prob = (double *)malloc(sizeof(double) * n);
  t->alias = (int *)malloc(sizeof(int) * n);
  double *p = (double *)malloc(sizeof(double) * n);
  int *small = (int *)malloc(sizeof(int) * n), *large = (int *)malloc(sizeof(int) * n);
  int ns = 0, nl = 0;
  for (int i = 0; i < n; i++) {
    p[i] = w[i] * (double)n / total;
    if (p[i] < 1.0) small[ns++] = i; else large[nl++] = i;
  }
  while (ns > 0 && nl > 0) {
    int s = small[--ns], l = large[--nl];
    t->prob[s] = p[s];
    t->alias[s] = l;
    p[l] = (p[l] + p[s]) - 1.0;
    if (p[l] < 1.0) small[ns++] = l; else large[nl++] = l;
  }
  while (nl > 0) { int l = large[--nl]; t->prob[l] = 1.0; t->alias[l] = l; }
  while (ns > 0) { int s = small[--ns]; t->prob[s] = 1.0; t->alias[s] = s; }
  free(p); free(small); free(large);
  return 0;
}
/* one-uniform draw; the role of AliasSampler::draw (alias_sampler.cpp:44-49) */
static int alias_draw(const alias_t *t, double u) {
  double ru = u * (double)t->n;
  int k = (int)ru;
  if (k >= t->n) k = t->n - 1;
  return (ru - (double)k) < t->prob[k] ? k : t->alias[k];
}
static void alias_free(alias_t *t) { free(t->prob); free(t->alias); t->prob = NULL; t->alias = NULL; }

/* ------------------------------------------------------------------ state */
typedef struct {
  int D, is_const, exclude;
  double *psi, *phi;              /* empirical pmf, entity-value pmf */
  double *logpsi, *log1mpsi, *om; /* log psi, log(1-psi), 1-psi */
  double *logphi, *logE;          /* log phi; log E_{v~phi} p(x|v) (attribute_index.cpp:55-62,94-101) */
  double Z;                       /* sum_v phi/(1-psi) (attribute_index.cpp:26-34) */
  int64_t *row;                   /* inverted neighbourhood: row w -> (v, p(w|v)), v ascending */
  int *col; double *p, *logp;
  double *cs;                     /* per row: inclusive prefix sums of phi(v) mef(v) p(w|v) */
  double *cs2;                    /* ... and of phi(v) (mef(v) p(w|v))^2: two members with the same value */
  double *mef;                    /* max_exp_factor (1 for constant, attribute_index.cpp:49) */
  double **g; double G[NMAX_FAST + 1]; alias_t Tg[NMAX_FAST + 1]; /* phi/(1-psi)^n */
  alias_t Tpsi;
  double s1, s2;                  /* Beta prior on theta (<=0: fixed) */
  double conc;                    /* DP concentration b (INFINITY: none) */
  double conc_shape, conc_rate;   /* Gamma prior on b (<= 0: fixed) */
  int dirichlet; double *dir_alpha; /* Dirichlet prior on phi */
} attr_t;

typedef struct exo {
  int N, A, F, K, iteration;
  uint64_t seed;
  int *x, *z, *fid, *lambda;      /* [N*A] record-major, NA = -1 */
  int *ey, *esize;                /* entities: [2N*A], [2N] */
  double *theta; int *ndist;      /* [F*A] index f*A + a */
  attr_t *at;
  exb200_clust cl;
  uint32_t host_ctr;
  int err_rec; char err[256];
  /* statistics of the last sweep */
  long long n_cand; int n_changed, n_wide;
  /* scratch for the links update */
  double *Lnew; int **bptr, **bitem; /* per attribute: items bucketed by their value */
  int *mhead, *mnext;             /* member lists (ascending record id) during the links update */
  int brute, has_dp;
  int wide_threshold;             /* scan order: records with more possible links go first */
  int *corr;                      /* [F*A] corrected non-distorted counts of the last sweep */
} exo;

static double hostu(exo *s) {
  uint32_t j = s->host_ctr++;
  double a, b;
  rng2(s->seed, j >> 1, (uint32_t)s->iteration, PH_HOST, 0, 0, &a, &b);
  return (j & 1u) ? b : a;
}
static double rnorm_(exo *s) {
  double u1 = hostu(s), u2 = hostu(s);
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
}
/* Marsaglia-Tsang; stands in for R::rgamma(shape, scale) */
static double rgamma_(exo *s, double a, double scale) {
  if (!(a > 0.0)) return 0.0;
  if (a < 1.0) {
    double u = hostu(s);
    return rgamma_(s, a + 1.0, scale) * pow(u, 1.0 / a);
  }
  double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  for (;;) {
    double x = rnorm_(s), v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    double u = hostu(s), x2 = x * x;
    if (u < 1.0 - 0.0331 * x2 * x2) return d * v * scale;
    if (log(u) < 0.5 * x2 + d * (1.0 - v + log(v))) return d * v * scale;
  }
}
/* stands in for R::rbeta; rbeta(a,0)=1 and rbeta(0,b)=0 as in R's nmath */
static double rbeta_(exo *s, double a, double b) {
  if (b == 0.0) return 1.0;
  if (a == 0.0) return 0.0;
  double x = rgamma_(s, a, 1.0), y = rgamma_(s, b, 1.0);
  return x / (x + y);
}
static double rpois_(exo *s, double lam) {
  if (!(lam > 0.0)) return 0.0;
  if (lam < 10.0) {
    double L = exp(-lam), p = 1.0; long k = 0;
    do { k++; p *= hostu(s); } while (p > L);
    return (double)(k - 1);
  }
  double slam = sqrt(lam), loglam = log(lam), b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
  double invalpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2.0);
  for (;;) {
    double U = hostu(s) - 0.5, V = hostu(s), us = 0.5 - fabs(U);
    double k = floor((2.0 * a / us + b) * U + lam + 0.43);
    if (us >= 0.07 && V <= vr) return k;
    if (k < 0.0 || (us < 0.013 && V > us)) continue;
    if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -lam + k * loglam - lgamma(k + 1.0)) return k;
  }
}
/* gamma / beta variates from keyed uniforms (one stream per (id, attr): slot advances) */
typedef struct { uint64_t seed; uint32_t id, iter, phase, attr, slot; int half; double ub; } kstream;
static double ks_u(kstream *k) {
  if (k->half) { k->half = 0; return k->ub; }
  double a, b;
  rng2(k->seed, k->id, k->iter, k->phase, k->attr, k->slot++, &a, &b);
  k->ub = b; k->half = 1;
  return a;
}
static double ks_norm(kstream *k) {
  double u1 = ks_u(k), u2 = ks_u(k);
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
}
static double ks_gamma(kstream *k, double a) {
  if (!(a > 0.0)) return 0.0;
  if (a < 1.0) { double u = ks_u(k); return ks_gamma(k, a + 1.0) * pow(u, 1.0 / a); }
  double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  for (;;) {
    double x = ks_norm(k), v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    double u = ks_u(k), x2 = x * x;
    if (u < 1.0 - 0.0331 * x2 * x2) return d * v;
    if (log(u) < 0.5 * x2 + d * (1.0 - v + log(v))) return d * v;
  }
}
static double ks_beta(kstream *k, double a, double b) {
  if (b == 0.0) return 1.0;
  if (a == 0.0) return 0.0;
  double x = ks_gamma(k, a), y = ks_gamma(k, b);
  return x / (x + y);
}

/* stands in for R::rnbinom(size, prob): gamma-Poisson mixture */
static double rnbinom_(exo *s, double size, double prob) {
  if (prob >= 1.0) return 0.0;
  return rpois_(s, rgamma_(s, size, (1.0 - prob) / prob));
}

/* ------------------------------------------------------------------ tables */
static double prior_existing(const exb200_clust *c, int n) { /* clust_params.cpp:4-27 */
  switch (c->kind) {
    case EXB200_CLUST_PITMAN_YOR: return fabs((double)n - c->d);
    case EXB200_CLUST_GEN_COUPON: return (double)n + c->kappa;
    default: return 1.0;
  }
}
static double prior_new(const exb200_clust *c, int k) { /* clust_params.cpp:9-31 */
  switch (c->kind) {
    case EXB200_CLUST_PITMAN_YOR: return c->alpha + c->d * (double)k;
    case EXB200_CLUST_GEN_COUPON: return c->kappa * fabs(c->m - (double)k);
    default: return fabs(c->m - (double)k);
  }
}

/* p(w | v) = get_distortion_prob(v, w), as a log (attribute_index.cpp:36-47, 77-92) */
static double log_distortion_prob(const attr_t *t, int v, int w) {
  if (t->is_const) {
    if (t->exclude) return v == w ? -INFINITY : t->logpsi[w] - t->log1mpsi[v];
    return t->logpsi[w];
  }
  if (v == w) return t->exclude ? -INFINITY : 0.0;
  int64_t lo = t->row[w], hi = t->row[w + 1];
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (t->col[mid] < v) lo = mid + 1; else hi = mid;
  }
  return (lo < t->row[w + 1] && t->col[lo] == v) ? t->logp[lo] : -INFINITY;
}

static double distortion_prob(const attr_t *t, int v, int w) {
  double l = log_distortion_prob(t, v, w);
  if (!(l > -INFINITY)) return 0.0;
  if (t->is_const) return t->exclude ? t->psi[w] / t->om[v] : t->psi[w];
  if (v == w) return 1.0;
  int64_t lo = t->row[w], hi = t->row[w + 1];
  while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (t->col[mid] < v) lo = mid + 1; else hi = mid; }
  return t->p[lo];
}

/* everything that depends on phi: rebuilt whenever the entity-value pmf changes
   (Cache::update, cache.cpp:20-27) */
static int attr_refresh(attr_t *t) {
  int D = t->D;
  t->Z = 0.0;
  for (int v = 0; v < D; v++) {
    t->logphi[v] = log(t->phi[v]);
    t->Z += t->phi[v] / t->om[v];
  }
  for (int w = 0; w < D; w++) {
    double e;
    if (t->is_const) {
      e = t->exclude ? t->Z * t->psi[w] - t->phi[w] * t->psi[w] / t->om[w] : t->psi[w];
    } else {
      e = 0.0;
      for (int64_t j = t->row[w]; j < t->row[w + 1]; j++) e += t->phi[t->col[j]] * t->p[j];
    }
    t->logE[w] = log(e);
  }
  if (!t->is_const)
    for (int w = 0; w < D; w++) {
      double c = 0.0, c2 = 0.0;
      for (int64_t j = t->row[w]; j < t->row[w + 1]; j++) {
        const double base = t->phi[t->col[j]] * t->mef[t->col[j]] * t->p[j];
        c += base;
        t->cs[j] = c;
        c2 += base * (t->mef[t->col[j]] * t->p[j]);
        t->cs2[j] = c2;
      }
    }
  for (int n = 0; n <= NMAX_FAST; n++) {
    if (t->Tg[n].prob) alias_free(&t->Tg[n]);
    double G = 0.0;
    for (int v = 0; v < D; v++) {
      t->g[n][v] = n == 0 ? t->phi[v] : t->g[n - 1][v] / t->om[v];
      G += t->g[n][v];
    }
    t->G[n] = G;
    if (alias_build(&t->Tg[n], t->g[n], D)) return -1;
    if (!t->is_const) break; /* string attributes only need phi itself */
  }
  return 0;
}

/* ------------------------------------------------------------------ create / destroy */
static void set_err(exo *s, int rec, const char *msg) {
  if (s->err[0]) return;
  s->err_rec = rec;
  snprintf(s->err, sizeof s->err, "%s (record/entity %d)", msg, rec);
}
const char *exo_last_error(exo *s) { return s->err; }

void exo_destroy(exo *s);
static void phase_entity_dists(exo *s);

exo *exo_create(const exb200_model *m) {
  exo *s = (exo *)calloc(1, sizeof(exo));
  int N = m->n_records, A = m->n_attrs, F = m->n_files;
  s->N = N; s->A = A; s->F = F; s->iteration = m->iteration; s->seed = m->seed; s->cl = m->clust;
  s->x = (int *)malloc(sizeof(int) * N * A);
  s->z = (int *)malloc(sizeof(int) * N * A);
  s->fid = (int *)malloc(sizeof(int) * N);
  s->lambda = (int *)malloc(sizeof(int) * N);
  s->ey = (int *)calloc((size_t)2 * N * A, sizeof(int));
  s->esize = (int *)calloc((size_t)2 * N, sizeof(int));
  s->theta = (double *)malloc(sizeof(double) * F * A);
  s->ndist = (int *)calloc(F * A, sizeof(int));
  s->Lnew = (double *)malloc(sizeof(double) * N);
  s->mhead = (int *)malloc(sizeof(int) * 2 * N); s->mnext = (int *)malloc(sizeof(int) * N);
  s->corr = (int *)calloc(F * A, sizeof(int));
  s->wide_threshold = 4096;
  s->at = (attr_t *)calloc(A, sizeof(attr_t));
  for (int i = 0; i < N * A; i++) {
    s->x[i] = m->rec_attrs[i] < 0 ? -1 : m->rec_attrs[i];
    s->z[i] = m->rec_distortions[i] != 0;
  }
  for (int r = 0; r < N; r++) s->fid[r] = m->file_ids[r];
  for (int f = 0; f < F; f++)
    for (int a = 0; a < A; a++) s->theta[f * A + a] = m->distort_probs[f + F * a];
  for (int a = 0; a < A; a++) {
    const exb200_attr *in = &m->attrs[a];
    attr_t *t = &s->at[a];
    int D = in->domain_size;
    t->D = D; t->is_const = in->is_constant; t->exclude = in->exclude_entity_value;
    t->s1 = in->distort_prob_shape1; t->s2 = in->distort_prob_shape2;
    t->conc = in->distort_dist_conc; t->conc_shape = in->conc_prior_shape; t->conc_rate = in->conc_prior_rate;
    t->dirichlet = in->entity_dist_kind == EXB200_ENTDIST_DIRICHLET;
    if (isfinite(t->conc) && !t->exclude) set_err(s, a, "a finite concentration requires exclude_entity_value");
    if (t->dirichlet) {
      t->dir_alpha = (double *)malloc(sizeof(double) * D);
      for (int v = 0; v < D; v++) t->dir_alpha[v] = in->entity_dist[v];
    }
    t->psi = (double *)malloc(sizeof(double) * D); t->phi = (double *)malloc(sizeof(double) * D);
    t->logpsi = (double *)malloc(sizeof(double) * D); t->log1mpsi = (double *)malloc(sizeof(double) * D);
    t->om = (double *)malloc(sizeof(double) * D); t->logphi = (double *)malloc(sizeof(double) * D);
    t->logE = (double *)malloc(sizeof(double) * D); t->mef = (double *)malloc(sizeof(double) * D);
    t->g = (double **)calloc(NMAX_FAST + 1, sizeof(double *));
    for (int n = 0; n <= NMAX_FAST; n++) t->g[n] = (double *)malloc(sizeof(double) * D);
    for (int v = 0; v < D; v++) {
      t->psi[v] = in->probs[v];
      t->phi[v] = (in->entity_dist && !t->dirichlet) ? in->entity_dist[v] : in->probs[v];
      t->om[v] = 1.0 - t->psi[v];
      t->logpsi[v] = log(t->psi[v]);
      t->log1mpsi[v] = log(t->om[v]);
      t->mef[v] = in->is_constant ? 1.0 : in->max_exp_factor[v];
    }
    if (!in->is_constant) { /* invert close_values: attribute_index.cpp:125-140 */
      int64_t nnz = in->close_row_ptr[D];
      t->row = (int64_t *)calloc(D + 1, sizeof(int64_t));
      t->col = (int *)malloc(sizeof(int) * (nnz + 1)); t->p = (double *)malloc(sizeof(double) * (nnz + 1));
      t->logp = (double *)malloc(sizeof(double) * (nnz + 1));
      t->cs = (double *)malloc(sizeof(double) * (nnz + 1));
      t->cs2 = (double *)malloc(sizeof(double) * (nnz + 1));
      for (int64_t j = 0; j < nnz; j++) t->row[in->close_val_ids[j] + 1]++;
      for (int w = 0; w < D; w++) t->row[w + 1] += t->row[w];
      int64_t *cur = (int64_t *)malloc(sizeof(int64_t) * D);
      memcpy(cur, t->row, sizeof(int64_t) * D);
      for (int v = 0; v < D; v++)
        for (int64_t j = in->close_row_ptr[v]; j < in->close_row_ptr[v + 1]; j++) {
          int w = in->close_val_ids[j];
          int64_t q = cur[w]++;
          t->col[q] = v; t->p[q] = in->close_probs[j]; t->logp[q] = log(in->close_probs[j]);
        }
      free(cur);
    }
    if (alias_build(&t->Tpsi, t->psi, D) || attr_refresh(t)) set_err(s, a, "invalid pmf");
  }
  s->bptr = (int **)malloc(sizeof(int *) * A);
  s->bitem = (int **)malloc(sizeof(int *) * A);
  for (int a = 0; a < A; a++) {
    s->bptr[a] = (int *)malloc(sizeof(int) * (s->at[a].D + 2));
    s->bitem[a] = (int *)malloc(sizeof(int) * 2 * N);
  }
  /* labels -> smallest member record (entities without records are dropped) */
  {
    int maxlab = 0;
    for (int e = 0; e < m->n_entities; e++) if (m->ent_ids[e] > maxlab) maxlab = m->ent_ids[e];
    for (int r = 0; r < N; r++) if (m->links[r] > maxlab) maxlab = m->links[r];
    int *slot = (int *)malloc(sizeof(int) * (maxlab + 1)), *rep = (int *)malloc(sizeof(int) * (maxlab + 1));
    for (int i = 0; i <= maxlab; i++) { slot[i] = -1; rep[i] = -1; }
    for (int e = 0; e < m->n_entities; e++) slot[m->ent_ids[e]] = e;
    for (int r = N - 1; r >= 0; r--) {
      if (m->links[r] < 0 || slot[m->links[r]] < 0) { set_err(s, r, "link to unknown entity"); continue; }
      rep[m->links[r]] = r;
    }
    for (int r = 0; r < N && !s->err[0]; r++) {
      int lab = m->links[r], c = rep[lab];
      s->lambda[r] = c;
      if (s->esize[c]++ == 0)
        for (int a = 0; a < A; a++) s->ey[c * A + a] = m->ent_attrs[a + (size_t)A * slot[lab]];
    }
    free(slot); free(rep);
  }
  s->K = 0;
  for (int e = 0; e < N; e++) s->K += s->esize[e] > 0;
  s->host_ctr = 0x40000000u; /* the constructor's draw has its own counter range */
  phase_entity_dists(s);
  for (int r = 0; r < N; r++)
    for (int a = 0; a < A; a++) s->ndist[s->fid[r] * A + a] += s->z[r * A + a];
  return s;
}

void exo_destroy(exo *s) {
  if (!s) return;
  for (int a = 0; a < s->A; a++) {
    attr_t *t = &s->at[a];
    free(t->psi); free(t->phi); free(t->logpsi); free(t->log1mpsi); free(t->om); free(t->logphi);
    free(t->logE); free(t->mef); free(t->row); free(t->col); free(t->p); free(t->logp); free(t->cs); free(t->cs2);
    for (int n = 0; n <= NMAX_FAST; n++) { free(t->g[n]); if (t->Tg[n].prob) alias_free(&t->Tg[n]); }
    free(t->g); if (t->Tpsi.prob) alias_free(&t->Tpsi);
  }
  free(s->x); free(s->z); free(s->fid); free(s->lambda); free(s->ey); free(s->esize); free(s->theta);
  free(s->ndist); free(s->corr); free(s->mhead); free(s->mnext); free(s->Lnew);
  for (int a = 0; s->bptr && a < s->A; a++) { free(s->bptr[a]); free(s->bitem[a]); }
  free(s->bitem); free(s->bptr); free(s->at); free(s);
}

/* ------------------------------------------------------------------ 1. cluster prior */
/* ClustParams::update (clust_params.cpp:37-72, 74-121).  The Bernoulli sums are keyed per
   index / per (entity, j) so that they can be evaluated in any order. */
static void phase_clust(exo *s) {
  exb200_clust *c = &s->cl;
  const int N = s->N, K = s->K;
  uint32_t it = (uint32_t)s->iteration;
  double u;
  if (c->kind == EXB200_CLUST_PITMAN_YOR) {
    int has_a = c->alpha_prior_shape > 0, has_d = c->d_prior_shape1 > 0;
    if (!has_a && !has_d) return;
    long succ = 0;
    for (int i = 1; i <= K - 1; i++) { /* :83-90 */
      rng2(s->seed, (uint32_t)i, it, PH_CLUST_U, 0, 0, &u, NULL);
      succ += u < c->alpha / (c->alpha + c->d * (double)i);
    }
    long fail = (K - 1) - succ;
    long vfail = 0;
    if (has_d) /* :104-115, with the d and alpha of the start of the update */
      for (int e = 0; e < N; e++)
        for (int j = 1; j <= s->esize[e] - 1; j++) {
          rng2(s->seed, (uint32_t)e, it, PH_CLUST_V, 0, (uint32_t)j, &u, NULL);
          vfail += u >= ((double)j - 1.0) / ((double)j - c->d);
        }
    if (has_a) { /* :92-101 */
      double w = rbeta_(s, c->alpha + 1.0, (double)N - 1.0);
      double shape = c->alpha_prior_shape + (double)succ, rate = c->alpha_prior_rate - log(w);
      c->alpha = rgamma_(s, shape, 1.0 / rate);
    }
    if (has_d) c->d = rbeta_(s, c->d_prior_shape1 + (double)fail, c->d_prior_shape2 + (double)vfail); /* :117-120 */
  } else if (c->kind == EXB200_CLUST_GEN_COUPON) {
    int has_k = c->kappa_prior_shape > 0, has_m = c->m_prior_size > 0;
    if (!has_k && !has_m) return;
    double w = rbeta_(s, c->m * c->kappa + 1.0, (double)N - 1.0); /* :45 */
    if (has_k) {
      long succ = 0;
      for (int e = 0; e < N; e++)
        for (int j = 1; j <= s->esize[e] - 1; j++) { /* :52-58 */
          rng2(s->seed, (uint32_t)e, it, PH_CLUST_V, 0, (uint32_t)j, &u, NULL);
          succ += u < c->kappa / (c->kappa + (double)j);
        }
      double shape = c->kappa_prior_shape + (double)K - 1.0 + (double)succ;
      double rate = c->kappa_prior_rate - c->m * log(w);
      c->kappa = rgamma_(s, shape, 1.0 / rate); /* :61-63 */
    }
    if (has_m) { /* :66-71; m_ is an int in the reference */
      double sz = c->m_prior_size + (double)K - 1.0;
      double pb = 1.0 - (1.0 - c->m_prior_prob) * pow(w, c->kappa);
      c->m = (double)(int)(rnbinom_(s, sz, pb) + (double)K);
    }
  }
}

/* ------------------------------------------------------------------ 2. links */
/* attributes a new entity founded by record r would get (links.cpp:424-468), and
   logLikelihoodWeightNew (links.cpp:316-347) */
static void links_prepare(exo *s) {
  const int N = s->N, A = s->A;
  uint32_t it = (uint32_t)s->iteration;
  for (int r = 0; r < N; r++) {
    int *yn = &s->ey[(size_t)(N + r) * A];
    double L = 0.0;
    s->esize[N + r] = 0;
    for (int a = 0; a < A; a++) {
      const attr_t *t = &s->at[a];
      int xv = s->x[r * A + a], zz = s->z[r * A + a];
      double u;
      if (xv < 0) { /* :434-437 prior draw */
        rng2(s->seed, (uint32_t)r, it, PH_LINK_NEWATTR, (uint32_t)a, 0, &u, NULL);
        yn[a] = alias_draw(&t->Tg[0], u);
        continue;
      }
      L += zz ? t->logE[xv] : t->logphi[xv];
      if (!zz) { yn[a] = xv; continue; } /* :440-442 */
      if (t->is_const) {
        if (t->exclude) { /* target of rejectionSampleConstant (:290-314): phi(v)/(1-psi(v)), v != x */
          int v = xv;
          for (uint32_t k = 0; k < REJECT_CAP; k++) {
            rng2(s->seed, (uint32_t)r, it, PH_LINK_NEWATTR, (uint32_t)a, k, &u, NULL);
            v = alias_draw(&t->Tg[1], u);
            if (v != xv) break;
          }
          if (v == xv) set_err(s, r, "rejection sampler did not terminate");
          yn[a] = v;
        } else { /* :449-452 empirical distribution */
          rng2(s->seed, (uint32_t)r, it, PH_LINK_NEWATTR, (uint32_t)a, 0, &u, NULL);
          yn[a] = alias_draw(&t->Tpsi, u);
        }
      } else { /* :454-461 phi(v) p(x|v) over the neighbourhood of x */
        double tot = 0.0;
        for (int64_t j = t->row[xv]; j < t->row[xv + 1]; j++) tot += t->phi[t->col[j]] * t->p[j];
        yn[a] = xv;
        if (tot > 0.0) {
          rng2(s->seed, (uint32_t)r, it, PH_LINK_NEWATTR, (uint32_t)a, 0, &u, NULL);
          /* first entry whose cumulative weight exceeds u * total, else the last positive one */
          double tt = u * tot, c = 0.0;
          int pick = -1, lastpos = xv;
          for (int64_t j = t->row[xv]; j < t->row[xv + 1]; j++) {
            double w = t->phi[t->col[j]] * t->p[j];
            if (w > 0.0) lastpos = t->col[j];
            c += w;
            if (tt < c) { pick = t->col[j]; break; }
          }
          yn[a] = pick >= 0 ? pick : lastpos;
        }
      }
    }
    s->Lnew[r] = L;
  }
  /* blocking buckets, one index per attribute: items (entity slots 0..2N-1) grouped by their value
     of the attribute, ascending id inside a bucket.  A record searches the shortest bucket among
     its observed non-distorted attributes (pick_bucket) - the role of the set-size ordering of
     getPossibleLinks (links.cpp:187-189); whichever bucket is used, the compatible items come out
     in ascending id. */
  for (int a0 = 0; a0 < A; a0++) {
    int D = s->at[a0].D, *bp = s->bptr[a0], *bi = s->bitem[a0];
    memset(bp, 0, sizeof(int) * (D + 2));
    for (int e = 0; e < 2 * N; e++) bp[s->ey[(size_t)e * A + a0] + 1]++;
    for (int v = 0; v < D; v++) bp[v + 1] += bp[v];
    int *cur = (int *)malloc(sizeof(int) * (D + 1));
    memcpy(cur, bp, sizeof(int) * (D + 1));
    for (int e = 0; e < 2 * N; e++) bi[cur[s->ey[(size_t)e * A + a0]]++] = e;
    free(cur);
  }
}

/* the shortest bucket among the record's observed non-distorted attributes: items[lo..hi), or
   every item (items = NULL) when there is none or in brute mode */
static void pick_bucket(const exo *s, int r, const int **items, int *lo, int *hi) {
  const int A = s->A;
  *items = NULL; *lo = 0; *hi = 2 * s->N;
  if (s->brute) return;
  for (int a = 0; a < A; a++) {
    int xv = s->x[r * A + a];
    if (xv < 0 || s->z[r * A + a]) continue;
    int b0 = s->bptr[a][xv], b1 = s->bptr[a][xv + 1];
    if (b1 - b0 < *hi - *lo) { *items = s->bitem[a]; *lo = b0; *hi = b1; }
  }
}

/* compatible(r, e): e agrees with r on every observed, non-distorted attribute of r - the
   intersection getPossibleLinks builds at links.cpp:122-131 (the distorted-attribute sets
   at :132-171 only drop zero-weight entities) */
static int compatible(const exo *s, int r, int e) {
  const int A = s->A;
  for (int a = 0; a < A; a++) {
    int xv = s->x[r * A + a];
    if (xv >= 0 && !s->z[r * A + a] && s->ey[(size_t)e * A + a] != xv) return 0;
  }
  return 1;
}
/* logLikelihoodWeightExisting (links.cpp:216-288), the part that does not depend on the other
   members of e: attributes with infinite concentration */
static double loglik_existing(const exo *s, int r, int e) {
  const int A = s->A;
  double l = 0.0;
  for (int a = 0; a < A; a++) {
    int xv = s->x[r * A + a];
    if (xv >= 0 && s->z[r * A + a] && !isfinite(s->at[a].conc))
      l += log_distortion_prob(&s->at[a], s->ey[(size_t)e * A + a], xv);
  }
  return l;
}
/* ... and the part that does (finite concentration b, links.cpp:251-279): the other members of e
   whose value of the attribute is distorted and observed form the urn */
static double loglik_existing_dp(const exo *s, int r, int e) {
  const int A = s->A;
  double l = 0.0;
  for (int a = 0; a < A; a++) {
    const attr_t *t = &s->at[a];
    int xv = s->x[r * A + a];
    if (!(xv >= 0 && s->z[r * A + a] && isfinite(t->conc))) continue;
    int c_dist = 0, c_same = 0;
    for (int q = s->mhead[e]; q >= 0; q = s->mnext[q]) {
      if (q == r) continue;
      int xq = s->x[q * A + a];
      if (xq >= 0 && s->z[q * A + a]) { c_dist++; c_same += xq == xv; }
    }
    l += log((double)c_same + t->conc * distortion_prob(t, s->ey[(size_t)e * A + a], xv));
    l += -log((double)c_dist + t->conc);
  }
  return l;
}
static int any_finite_conc(const exo *s) {
  for (int a = 0; a < s->A; a++) if (isfinite(s->at[a].conc)) return 1;
  return 0;
}

/* Enumerate the conditional of record r (links.cpp:361-399) in the current state:
   live compatible entities in ascending id with log(prior) + log-likelihood.
   `removed` tells whether r's own singleton entity counts as removed (links.cpp:358-359). */
static int link_candidates(const exo *s, int r, int **ids, double **lw, int *cap) {
  int own = s->lambda[r], single = s->esize[own] == 1;
  int n = 0;
  const int *items; int lo, hi;
  pick_bucket(s, r, &items, &lo, &hi);
  for (int q = lo; q < hi; q++) {
    int e = items ? items[q] : q;
    int sz = s->esize[e];
    if (sz <= 0) continue;
    if (e == own) { if (single) continue; sz -= 1; } /* :359, :375 */
    if (!compatible(s, r, e)) continue;
    if (n == *cap) {
      *cap = *cap ? *cap * 2 : 64;
      *ids = (int *)realloc(*ids, sizeof(int) * *cap);
      *lw = (double *)realloc(*lw, sizeof(double) * *cap);
    }
    (*ids)[n] = e;
    (*lw)[n] = log(prior_existing(&s->cl, sz)) + loglik_existing(s, r, e);
    if (s->has_dp) (*lw)[n] += loglik_existing_dp(s, r, e);
    n++;
  }
  return n;
}

/* number of possible links of record r: compatible items of non-zero likelihood among all 2N
   entity slots of the sweep (whatever their current size), its own potential entity excluded */
static int count_possible_links(const exo *s, int r, int stop_after) {
  const int N = s->N;
  const int *items; int lo, hi;
  pick_bucket(s, r, &items, &lo, &hi);
  int n = 0;
  for (int q = lo; q < hi && n <= stop_after; q++) {
    int e = items ? items[q] : q;
    if (e == N + r) continue;
    if (e < N && s->esize[e] <= 0) continue; /* slots [0,N) without an entity at the start of the sweep */
    if (!compatible(s, r, e)) continue;
    if (loglik_existing(s, r, e) > -INFINITY) n++;
  }
  return n;
}

static void phase_links(exo *s) {
  const int N = s->N;
  uint32_t it = (uint32_t)s->iteration;
  int *ids = NULL, cap = 0; double *lw = NULL;
  links_prepare(s);
  s->n_cand = 0; s->n_changed = 0;
  s->has_dp = any_finite_conc(s);
  for (int e = 0; e < 2 * N; e++) s->mhead[e] = -1;
  for (int r = N - 1; r >= 0; r--) { s->mnext[r] = s->mhead[s->lambda[r]]; s->mhead[s->lambda[r]] = r; }
  /* scan order: records with more than wide_threshold possible links first, then the others */
  int *order = (int *)malloc(sizeof(int) * N), no = 0;
  char *is_wide = (char *)calloc(N, 1);
  for (int r = 0; r < N; r++) is_wide[r] = count_possible_links(s, r, s->wide_threshold) > s->wide_threshold;
  for (int r = 0; r < N; r++) if (is_wide[r]) order[no++] = r;
  s->n_wide = no;
  for (int r = 0; r < N; r++) if (!is_wide[r]) order[no++] = r;
  free(is_wide);
  for (int oi = 0; oi < N; oi++) {
    int r = order[oi];
    int own = s->lambda[r], single = s->esize[own] == 1;
    int n = link_candidates(s, r, &ids, &lw, &cap);
    s->n_cand += n;
    int Know = s->K - single; /* ents.n_entities() after the removal, links.cpp:390 */
    double lwn = log(prior_new(&s->cl, Know)) + s->Lnew[r];
    double m1 = -INFINITY;
    for (int j = 0; j < n; j++) if (lw[j] > m1) m1 = lw[j];
    double srel = 0.0;
    if (m1 > -INFINITY) for (int j = 0; j < n; j++) srel += exp(lw[j] - m1);
    double m2 = m1 > lwn ? m1 : lwn;
    if (!(m2 > -INFINITY)) { set_err(s, r, "no valid links for record"); continue; } /* :415-419 */
    double wa = exp(lwn - m2), wb = m1 > -INFINITY ? srel * exp(m1 - m2) : 0.0;
    double u1, u2;
    rng2(s->seed, (uint32_t)r, it, PH_LINK_DECIDE, 0, 0, &u1, &u2);
    int target;
    if (u1 * (wa + wb) < wa) {
      target = N + r;
    } else {
      double tt = u2 * srel, c = 0.0;
      int pick = -1, lastpos = -1;
      for (int j = 0; j < n; j++) {
        double w = exp(lw[j] - m1);
        if (w > 0.0) lastpos = j;
        c += w;
        if (tt < c) { pick = j; break; }
      }
      if (pick < 0) pick = lastpos;
      target = ids[pick];
    }
    /* set_link + entity bookkeeping (links.cpp:6-39, entities.cpp:18-69) */
    s->esize[own]--;
    if (single) s->K--;
    if (target == N + r) s->K++;
    s->esize[target]++;
    s->lambda[r] = target;
    if (target != own) { /* member lists stay sorted by record id */
      int h = s->mhead[own];
      if (h == r) s->mhead[own] = s->mnext[r];
      else { while (s->mnext[h] != r) h = s->mnext[h]; s->mnext[h] = s->mnext[r]; }
      h = s->mhead[target];
      if (h < 0 || r < h) { s->mnext[r] = h; s->mhead[target] = r; }
      else { int nx = s->mnext[h]; while (nx >= 0 && nx < r) { h = nx; nx = s->mnext[h]; } s->mnext[r] = nx; s->mnext[h] = r; }
    }
    if (!(target == own || (single && target == N + r))) s->n_changed++;
  }
  free(ids); free(lw); free(order);
}

/* rename every entity to its smallest member record; attributes move along so that the
   state stays complete between phases (the next phase overwrites them anyway) */
static void phase_canon(exo *s) {
  const int N = s->N, A = s->A;
  int *rep = (int *)malloc(sizeof(int) * 2 * N);
  for (int e = 0; e < 2 * N; e++) rep[e] = -1;
  for (int r = N - 1; r >= 0; r--) rep[s->lambda[r]] = r;
  int *ny = (int *)calloc((size_t)2 * N * A, sizeof(int)), *ns = (int *)calloc((size_t)2 * N, sizeof(int));
  for (int e = 0; e < 2 * N; e++)
    if (rep[e] >= 0) {
      ns[rep[e]] = s->esize[e];
      memcpy(&ny[(size_t)rep[e] * A], &s->ey[(size_t)e * A], sizeof(int) * A);
    }
  for (int r = 0; r < N; r++) s->lambda[r] = rep[s->lambda[r]];
  free(s->ey); free(s->esize); free(rep);
  s->ey = ny; s->esize = ns;
}

/* ------------------------------------------------------------------ 3. entity attributes */
/* inclusive scan of 32 values in the association order of a warp's shuffle scan:
   for o = 1, 2, 4, 8, 16: x[i] += x[i - o] (all lanes at once) */
static void ks_scan32(double x[32]) {
  for (int o = 1; o < 32; o <<= 1) {
    double y[32];
    for (int i = 0; i < 32; i++) y[i] = i >= o ? x[i] + x[i - o] : x[i];
    memcpy(x, y, sizeof y);
  }
}

/* logWeightEntityValue (entities.cpp:236-296) */
static double log_weight_entity_value(const exo *s, int v, int a, const int *mem, int n) {
  const attr_t *t = &s->at[a];
  const int A = s->A;
  double lw = t->logphi[v], mefv = t->mef[v];
  int ctr_disagree = 0;
  for (int i = 0; i < n; i++) {
    int r = mem[i], xv = s->x[r * A + a];
    if (xv < 0) continue;
    double eff = s->theta[s->fid[r] * A + a] * mefv; /* :264 */
    if (xv == v) {
      if (t->exclude) lw += log(1.0 - eff);
      else lw += log(1.0 - eff + eff * (t->is_const ? t->psi[v] : 1.0)); /* :270, attribute_index.cpp:45,82 */
    } else {
      lw += log(eff);
      if (isfinite(t->conc)) { /* :275-288: Chinese-restaurant terms of the disagreeing values */
        double b = t->conc;
        ctr_disagree++;
        lw += -log((double)ctr_disagree - 1.0 + b);
        int seen = 0;
        for (int q = 0; q < i; q++) {
          int xq = s->x[mem[q] * A + a];
          if (xq >= 0 && xq != v && xq == xv) seen++;
        }
        if (seen) lw += log((double)seen + b * distortion_prob(t, v, xv));
        else lw += log(b * distortion_prob(t, v, xv));
      } else {
        lw += log_distortion_prob(t, v, xv); /* :290 */
      }
    }
  }
  return lw;
}

static int draw_entity_value(exo *s, int e, int a, const int *mem, int n) {
  const attr_t *t = &s->at[a];
  const int A = s->A;
  uint32_t it = (uint32_t)s->iteration;
  double u;
  int nobs = 0, first = -1;
  for (int i = 0; i < n; i++) {
    int xv = s->x[mem[i] * A + a];
    if (xv >= 0) { if (first < 0) first = xv; nobs++; }
  }
  rng2(s->seed, (uint32_t)e, it, PH_ENT_ATTR, (uint32_t)a, 0, &u, NULL);
  if (nobs == 0) return alias_draw(&t->Tg[0], u); /* entities.cpp:387-389 */

  if (t->is_const && t->exclude && nobs <= NMAX_FAST && (nobs == 1 || !isfinite(t->conc))) {
    /* drawEntityValueSequential (entities.cpp:298-348) without touching the whole domain:
       w(v) for v among the members' values explicitly, C * phi(v)/(1-psi(v))^nobs elsewhere */
    int vals[NMAX_FAST]; double w[NMAX_FAST]; int nv = 0;
    double Cc = 1.0;
    for (int i = 0; i < n; i++) {
      int r = mem[i], xv = s->x[r * A + a];
      if (xv < 0) continue;
      Cc *= s->theta[s->fid[r] * A + a] * t->psi[xv];
      int seen = 0;
      for (int j = 0; j < nv; j++) seen |= vals[j] == xv;
      if (!seen) vals[nv++] = xv;
    }
    double sumw = 0.0, rest = t->G[nobs];
    for (int j = 0; j < nv; j++) {
      int v = vals[j];
      double ww = t->phi[v];
      for (int i = 0; i < n; i++) {
        int r = mem[i], xv = s->x[r * A + a];
        if (xv < 0) continue;
        double th = s->theta[s->fid[r] * A + a];
        ww *= xv == v ? 1.0 - th : th * t->psi[xv] / t->om[v];
      }
      w[j] = ww; sumw += ww;
      rest -= t->g[nobs][v];
    }
    if (rest < 0.0) rest = 0.0;
    double wrest = Cc * rest, total = sumw + wrest;
    if (!(total > 0.0) || !isfinite(total)) { set_err(s, e, "no valid entity values for attribute"); return first; }
    double tt = u * total, c = 0.0;
    for (int j = 0; j < nv; j++) { c += w[j]; if (tt < c) return vals[j]; }
    if (!(wrest > 0.0)) { /* numerically empty remainder: fall back on the last positive value */
      for (int j = nv - 1; j >= 0; j--) if (w[j] > 0.0) return vals[j];
    }
    for (uint32_t k = 1; k < REJECT_CAP; k++) {
      rng2(s->seed, (uint32_t)e, it, PH_ENT_ATTR, (uint32_t)a, k, &u, NULL);
      int v = alias_draw(&t->Tg[nobs], u), hit = 0;
      for (int j = 0; j < nv; j++) hit |= vals[j] == v;
      if (!hit) return v;
    }
    set_err(s, e, "rejection sampler did not terminate");
    return first;
  }

  int same = 1; /* every observed member value equals the first one */
  for (int i = 0; i < n; i++) { int xv = s->x[mem[i] * A + a]; if (xv >= 0 && xv != first) same = 0; }
  if (!t->is_const && t->exclude && (nobs == 1 || (nobs == 2 && same && !isfinite(t->conc)))) {
    /* one observed member value x, held by one member (singletons above all) or by two: every
       v != x has weight (prod_m theta_m) phi(v) (mef(v) p(x|v))^nobs, whose row sums are static, so
       drawEntityValueIndex (entities.cpp:351-431) reduces to "keep x" versus a search in the row's
       prefix sums */
    double th = 1.0, wx = t->phi[first];
    for (int i = 0; i < n; i++)
      if (s->x[mem[i] * A + a] >= 0) {
        double thm = s->theta[s->fid[mem[i]] * A + a];
        th *= thm;
        wx *= 1.0 - thm * t->mef[first];
      }
    const double *cs = nobs == 1 ? t->cs : t->cs2;
    int64_t lo = t->row[first], hi = t->row[first + 1];
    double rsum = hi > lo ? cs[hi - 1] : 0.0;
    double total = wx + th * rsum;
    if (!(total > 0.0) || !isfinite(total)) { set_err(s, e, "no valid entity values for attribute"); return first; }
    double tt = u * total;
    if (tt < wx || hi == lo) return first;
    double t2 = tt - wx;
    int64_t a0 = lo, b0 = hi; /* first j with t2 < th * cs[j] */
    while (a0 < b0) { int64_t mid = (a0 + b0) >> 1; if (t2 < th * cs[mid]) b0 = mid; else a0 = mid + 1; }
    if (a0 == hi) { /* rounding: the first entry that reaches the row total */
      a0 = lo; b0 = hi - 1;
      while (a0 < b0) { int64_t mid = (a0 + b0) >> 1; if (cs[mid] >= rsum) b0 = mid; else a0 = mid + 1; }
    }
    return t->col[a0];
  }

  /* general scheme: explicit support.  Constant attribute: the whole domain
     (entities.cpp:315-329).  Otherwise the first observed member value followed by its
     neighbourhood row - every other value has p(x_first | v) = 0 (entities.cpp:361-383 builds a
     superset); a row entry equal to x_first (possible when the value is not excluded) gets
     weight 0 so that it is not counted twice.
     The weights exp(lw - max) are summed in chunks of 32 consecutive support entries with the
     Kogge-Stone order of a warp scan (ks_scan32), chunk totals sequentially: the order the CUDA
     kernel uses, so both sides round identically. */
  int64_t lo = t->is_const ? 0 : t->row[first];
  int ns = t->is_const ? t->D : 1 + (int)(t->row[first + 1] - lo);
  double *lw = (double *)malloc(sizeof(double) * ns), m = -INFINITY;
  int *sup = (int *)malloc(sizeof(int) * ns);
  for (int j = 0; j < ns; j++) {
    int v = t->is_const ? j : (j == 0 ? first : t->col[lo + j - 1]);
    sup[j] = v;
    lw[j] = (!t->is_const && j > 0 && v == first) ? -INFINITY : log_weight_entity_value(s, v, a, mem, n);
    if (lw[j] > m) m = lw[j];
  }
  int pick = -1;
  if (!(m > -INFINITY)) {
    set_err(s, e, "no valid entity values for attribute");
    pick = 0;
  } else {
    int nchunk = (ns + 31) / 32;
    double total = 0.0, sc[32];
    for (int c = 0; c < nchunk; c++) {
      for (int i = 0; i < 32; i++) sc[i] = 32 * c + i < ns ? exp(lw[32 * c + i] - m) : 0.0;
      ks_scan32(sc);
      total += sc[31];
    }
    double tt = u * total, run = 0.0;
    for (int c = 0; c < nchunk && pick < 0; c++) {
      for (int i = 0; i < 32; i++) sc[i] = 32 * c + i < ns ? exp(lw[32 * c + i] - m) : 0.0;
      ks_scan32(sc);
      if (tt < run + sc[31])
        for (int i = 0; i < 32 && 32 * c + i < ns; i++)
          if (tt < run + sc[i]) { pick = 32 * c + i; break; }
      run += sc[31];
    }
    if (pick < 0)
      for (int j = ns - 1; j >= 0; j--) if (exp(lw[j] - m) > 0.0) { pick = j; break; }
  }
  int v = sup[pick];
  free(lw); free(sup);
  return v;
}

static void phase_entities(exo *s) { /* Entities::update_attributes, entities.cpp:434-469 */
  const int N = s->N, A = s->A;
  int *head = (int *)malloc(sizeof(int) * N), *next = (int *)malloc(sizeof(int) * N);
  for (int e = 0; e < N; e++) head[e] = -1;
  for (int r = N - 1; r >= 0; r--) { next[r] = head[s->lambda[r]]; head[s->lambda[r]] = r; }
  int *mem = (int *)malloc(sizeof(int) * N);
  for (int e = 0; e < N; e++) {
    if (s->esize[e] <= 0) continue;
    int n = 0;
    for (int r = head[e]; r >= 0; r = next[r]) mem[n++] = r;
    for (int a = 0; a < A; a++) s->ey[(size_t)e * A + a] = draw_entity_value(s, e, a, mem, n);
  }
  free(head); free(next); free(mem);
}

/* ------------------------------------------------------------------ 3b. entity distributions */
/* Entities::update_distributions (entities.cpp:133-156) with rdirichlet (dirichlet.h:12-34):
   phi_a ~ Dirichlet(count_a(v) + alpha_v), then the tables that depend on phi (Cache::update) */
static void phase_entity_dists(exo *s) {
  const int N = s->N, A = s->A;
  for (int a = 0; a < A; a++) {
    attr_t *t = &s->at[a];
    if (!t->dirichlet) continue;
    double *g = (double *)malloc(sizeof(double) * t->D), total = 0.0;
    int *cnt = (int *)calloc(t->D, sizeof(int));
    for (int e = 0; e < N; e++) if (s->esize[e] > 0) cnt[s->ey[(size_t)e * A + a]]++;
    for (int v = 0; v < t->D; v++) { g[v] = rgamma_(s, (double)cnt[v] + t->dir_alpha[v], 1.0); total += g[v]; }
    for (int v = 0; v < t->D; v++) t->phi[v] = g[v] / total;
    free(g); free(cnt);
    if (attr_refresh(t)) set_err(s, a, "invalid entity distribution");
  }
}

/* ------------------------------------------------------------------ 3c. DP concentrations */
/* DistortDistConcs::update (distort_dist_concs.cpp:32-94), value counts per cluster.  Entities in
   ascending id, members in ascending id; the Bernoulli of member number i of entity e is keyed
   (e, attribute, i), the cluster's Beta variate has its own keyed stream; log t_e is accumulated in
   fixed point (2^-40) so that the sum does not depend on the order. */
static void phase_concs(exo *s) {
  const int N = s->N, A = s->A;
  uint32_t it = (uint32_t)s->iteration;
  int any = 0;
  for (int a = 0; a < A; a++) any |= s->at[a].conc_shape > 0;
  if (!any) return;
  int *head = (int *)malloc(sizeof(int) * N), *next = (int *)malloc(sizeof(int) * N);
  for (int e = 0; e < N; e++) head[e] = -1;
  for (int r = N - 1; r >= 0; r--) { next[r] = head[s->lambda[r]]; head[s->lambda[r]] = r; }
  for (int a = 0; a < A; a++) {
    attr_t *t = &s->at[a];
    if (!(t->conc_shape > 0)) continue;
    long long sum_zeta = 0, sum_logt_fx = 0;
    for (int e = 0; e < N; e++) {
      if (s->esize[e] <= 0) continue;
      int yv = s->ey[(size_t)e * A + a], ctr = 0, idx = 0;
      for (int r = head[e]; r >= 0; r = next[r], idx++) {
        int xv = s->x[r * A + a];
        if (xv < 0 || xv == yv) continue;
        int seen = 0; /* earlier members of this cluster with the same disagreeing value */
        for (int q = head[e]; q != r; q = next[q]) seen += s->x[q * A + a] == xv;
        double prob = 1.0;
        if (seen) { double zz = t->conc * distortion_prob(t, yv, xv); prob = zz / (zz + (double)seen); }
        double u;
        rng2(s->seed, (uint32_t)e, it, PH_CONC_ZETA, (uint32_t)a, (uint32_t)idx, &u, NULL);
        sum_zeta += u < prob;
        ctr++;
      }
      if (ctr > 0) { /* rbeta(b, 0) = 1 otherwise: log 1 = 0 */
        kstream k = {s->seed, (uint32_t)e, it, PH_CONC_BETA, (uint32_t)a, 0, 0, 0.0};
        double te = ks_beta(&k, t->conc, (double)ctr);
        if (te < 1e-300) te = 1e-300; /* a Beta(b, n) variate can underflow for tiny b */
        sum_logt_fx += llround(log(te) * 1099511627776.0);
      }
    }
    double shape = t->conc_shape + (double)sum_zeta;
    double rate = t->conc_rate - (double)sum_logt_fx / 1099511627776.0;
    t->conc = rgamma_(s, shape, 1.0 / rate);
  }
  free(head); free(next);
}

/* ------------------------------------------------------------------ 4. distortion */
static void phase_distort(exo *s, int *corr) { /* Records::update_distortion, records.cpp:70-138 */
  const int N = s->N, A = s->A, F = s->F;
  uint32_t it = (uint32_t)s->iteration;
  memset(corr, 0, sizeof(int) * F * A);
  memset(s->ndist, 0, sizeof(int) * F * A);
  for (int r = 0; r < N; r++) {
    int f = s->fid[r], e = s->lambda[r];
    for (int a = 0; a < A; a++) {
      const attr_t *t = &s->at[a];
      int xv = s->x[r * A + a], yv = s->ey[(size_t)e * A + a];
      double th = s->theta[f * A + a], ua, ub;
      rng2(s->seed, (uint32_t)r, it, PH_DISTORT, (uint32_t)a, 0, &ua, &ub);
      int zz;
      if (xv < 0) zz = ua < th; /* :95-100 */
      else if (xv == yv) {
        if (t->exclude) zz = 0; /* :106-107 */
        else { /* :109-112 */
          double pr1 = th * (t->is_const ? t->psi[xv] : 1.0), pr0 = 1.0 - th;
          zz = ua < pr1 / (pr1 + pr0);
        }
      } else zz = 1; /* :114-116 */
      s->z[r * A + a] = zz;
      s->ndist[f * A + a] += zz;
      double pr1 = t->mef[yv]; /* :122-133 */
      int alt;
      if (pr1 == 1.0) alt = 1;
      else {
        double pr0 = 1.0 - pr1;
        pr1 *= zz ? th : 1.0 - th;
        alt = ub < pr1 / (pr1 + pr0);
      }
      corr[f * A + a] += alt & !zz; /* :134 */
    }
  }
}

static void phase_theta(exo *s, const int *corr) { /* DistortProbs::update, distortion_probs.cpp:30-47 */
  for (int f = 0; f < s->F; f++)
    for (int a = 0; a < s->A; a++) {
      const attr_t *t = &s->at[a];
      if (t->s1 > 0)
        s->theta[f * s->A + a] = rbeta_(s, (double)s->ndist[f * s->A + a] + t->s1,
                                        (double)corr[f * s->A + a] + t->s2);
    }
}

/* ------------------------------------------------------------------ sweep */
/* State::update (state.h:60-75).  phases: bit0 clust, bit1 links(+rename), bit2 entity
   attributes, bit3 distortion indicators, bit4 distortion probs; the iteration advances
   when bit4 is set. */
int exo_sweep_phases(exo *s, int phases) {
  int *corr = s->corr;
  if (s->err[0]) return EXB200_ERR_WEIGHTS;
  if (phases & 1) { s->host_ctr = 0; phase_clust(s); }
  if (phases & 2) { phase_links(s); phase_canon(s); }
  if (phases & 4) { phase_entities(s); phase_entity_dists(s); phase_concs(s); }
  if (phases & 8) phase_distort(s, corr);
  if (phases & 16) { phase_theta(s, corr); s->iteration++; }
  return s->err[0] ? EXB200_ERR_WEIGHTS : EXB200_OK;
}
int exo_sweeps(exo *s, int n) {
  for (int i = 0; i < n; i++) {
    int rc = exo_sweep_phases(s, 31);
    if (rc) return rc;
  }
  return EXB200_OK;
}
void exo_set_brute(exo *s, int on) { s->brute = on; }
void exo_set_wide_threshold(exo *s, int w) { s->wide_threshold = w; }
int exo_n_wide(exo *s) { return s->n_wide; }

/* same layout as exb200_get_state */
void exo_get_state(exo *s, exb200_state *out) {
  const int N = s->N, A = s->A, F = s->F;
  out->iteration = s->iteration;
  int k = 0;
  for (int e = 0; e < N; e++)
    if (s->esize[e] > 0) {
      if (out->ent_ids) out->ent_ids[k] = e;
      if (out->ent_attrs) for (int a = 0; a < A; a++) out->ent_attrs[a + (size_t)A * k] = s->ey[(size_t)e * A + a];
      k++;
    }
  out->n_entities = k;
  if (out->links) memcpy(out->links, s->lambda, sizeof(int) * N);
  if (out->rec_distortions) for (int i = 0; i < N * A; i++) out->rec_distortions[i] = s->z[i];
  for (int f = 0; f < F; f++)
    for (int a = 0; a < A; a++) {
      if (out->distort_probs) out->distort_probs[f + F * a] = s->theta[f * A + a];
      if (out->distort_counts) out->distort_counts[f + F * a] = s->ndist[f * A + a];
    }
  if (out->distort_dist_concs) for (int a = 0; a < A; a++) out->distort_dist_concs[a] = s->at[a].conc;
  out->clust = s->cl;
}
void exo_get_stats(exo *s, long long *n_cand, int *n_changed, int *K) {
  *n_cand = s->n_cand; *n_changed = s->n_changed; *K = s->K;
}

/* the conditional of one link in the current state (same contract as
   exb200_link_conditional); the new-entity slots of the current sweep are ignored */
int exo_link_conditional(exo *s, int rid, int cap, int32_t *cand, double *lwout, double *lw_new) {
  int *ids = NULL, c = 0; double *lw = NULL;
  const int N = s->N, A = s->A;
  int save = s->brute; s->brute = 1;
  for (int r = 0; r < N; r++) s->esize[N + r] = 0;
  s->has_dp = any_finite_conc(s);
  for (int e = 0; e < 2 * N; e++) s->mhead[e] = -1;
  for (int r = N - 1; r >= 0; r--) { s->mnext[r] = s->mhead[s->lambda[r]]; s->mhead[s->lambda[r]] = r; }
  double L = 0.0;
  for (int a = 0; a < A; a++) {
    int xv = s->x[rid * A + a];
    if (xv >= 0) L += s->z[rid * A + a] ? s->at[a].logE[xv] : s->at[a].logphi[xv];
  }
  int n = link_candidates(s, rid, &ids, &lw, &c);
  s->brute = save;
  for (int j = 0; j < n && j < cap; j++) { cand[j] = ids[j]; lwout[j] = lw[j]; }
  int single = s->esize[s->lambda[rid]] == 1;
  *lw_new = log(prior_new(&s->cl, s->K - single)) + L;
  free(ids); free(lw);
  return n;
}
/* logWeightEntityValue of entity `eid` (canonical label) for value v of attribute a */
double exo_log_weight_entity_value(exo *s, int eid, int a, int v) {
  int *mem = (int *)malloc(sizeof(int) * s->N), n = 0;
  for (int r = 0; r < s->N; r++) if (s->lambda[r] == eid) mem[n++] = r;
  double lw = log_weight_entity_value(s, v, a, mem, n);
  free(mem);
  return lw;
}
/* host scalar samplers, exposed so that their moments can be tested */
void exo_test_draws(uint64_t seed, int kind, double p1, double p2, int n, double *out) {
  exo tmp; memset(&tmp, 0, sizeof tmp); tmp.seed = seed;
  for (int i = 0; i < n; i++) {
    switch (kind) {
      case 0: out[i] = hostu(&tmp); break;
      case 1: out[i] = rnorm_(&tmp); break;
      case 2: out[i] = rgamma_(&tmp, p1, p2); break;
      case 3: out[i] = rbeta_(&tmp, p1, p2); break;
      case 4: out[i] = rpois_(&tmp, p1); break;
      default: out[i] = rnbinom_(&tmp, p1, p2); break;
    }
  }
}
/* entity-value pmf of attribute a (moves with a Dirichlet prior, entities.cpp:133-156) */
void exo_get_phi(exo *s, int a, double *out) { memcpy(out, s->at[a].phi, sizeof(double) * s->at[a].D); }
/* Entities::update_distributions only */
void exo_update_distributions(exo *s) { phase_entity_dists(s); }

