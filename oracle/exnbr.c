/*
 * exnbr.c - thresholded-Levenshtein value neighbourhoods on the host cores.
 *
 * TEST / BENCHMARK INFRASTRUCTURE, like everything under oracle/: only tests/, the smoke test and
 * the CPU legs of bench.py (cpu_baseline, --impl reference) may load it; the product never does.
 *
 * Restates compute_close_values (reference R/AttributeIndex.R:117-137) for
 * dist_fn = transform_dist_fn(comparator::Levenshtein(), threshold) (README.md:82-84,
 * R/transform_dist_fn.R:27-40): for every domain value v the values w with unit-cost edit distance
 * d(v, w) <= threshold (v itself only if include_self), p(w | v) = exp(-d) / sum, and
 * max_exp_factor[v] = exp(-0.5 min_d / max_d_overall) (0 without a neighbour).  The R code makes
 * D^2 distance calls; this makes them in C on all host cores so that the reference arm of the
 * benchmark can prepare a 10^5..10^6-record model without the GPU.  Checked against the Python
 * restatement (exchanger_b200/model.py:compute_close_values) in tests/test_oracle.py.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define STRIDE 32 /* strings of at most 31 bytes, packed with a stride of 32 (as exb200_neighbours_create) */

typedef struct {
  int n, threshold, include_self;
  const uint8_t *chars;
  const int *lens;
  uint64_t *mask; /* presence of byte value (c & 63): one edit changes at most two bits */
  int64_t *row_ptr;
  int32_t **cols; /* per row */
  uint8_t **dist;
  int *cnt;
} nbr_t;

typedef struct { nbr_t *nb; int t, T; } job_t;

/* edit distance capped at thr + 1 (band of width 2 thr + 1) */
static int lev_capped(const uint8_t *a, int la, const uint8_t *b, int lb, int thr) {
  int prev[STRIDE + 1], cur[STRIDE + 1];
  if (abs(la - lb) > thr) return thr + 1;
  for (int j = 0; j <= lb; j++) prev[j] = j;
  for (int i = 1; i <= la; i++) {
    int lo = i - thr > 1 ? i - thr : 1, hi = i + thr < lb ? i + thr : lb;
    cur[lo - 1] = lo - 1 == 0 ? i : thr + 1;
    int best = cur[lo - 1]; /* column 0 (deleting the whole prefix) is part of the band while i <= thr */
    for (int j = lo; j <= hi; j++) {
      int sub = prev[j - 1] + (a[i - 1] != b[j - 1]);
      int del = (j <= i - 1 + thr ? prev[j] : thr + 1) + 1;
      int ins = cur[j - 1] + 1;
      int v = sub < del ? sub : del;
      v = v < ins ? v : ins;
      cur[j] = v;
      if (v < best) best = v;
    }
    if (hi < lb) cur[hi + 1] = thr + 1;
    if (best > thr) return thr + 1;
    memcpy(prev, cur, sizeof(int) * (lb + 1));
  }
  return prev[lb] > thr ? thr + 1 : prev[lb];
}

static void *worker(void *arg) {
  job_t *jb = (job_t *)arg;
  nbr_t *nb = jb->nb;
  const int n = nb->n, thr = nb->threshold;
  for (int i = jb->t; i < n; i += jb->T) { /* interleaved rows: even load */
    int cap = 16, c = 0;
    int32_t *cols = (int32_t *)malloc(sizeof(int32_t) * cap);
    uint8_t *dist = (uint8_t *)malloc(cap);
    const uint8_t *a = nb->chars + (size_t)i * STRIDE;
    const int la = nb->lens[i];
    const uint64_t ma = nb->mask[i];
    for (int j = 0; j < n; j++) {
      if (j == i && !nb->include_self) continue;
      if (abs(la - nb->lens[j]) > thr) continue;
      if (__builtin_popcountll(ma ^ nb->mask[j]) > 2 * thr) continue;
      int d = lev_capped(a, la, nb->chars + (size_t)j * STRIDE, nb->lens[j], thr);
      if (d > thr) continue;
      if (c == cap) { cap *= 2; cols = (int32_t *)realloc(cols, sizeof(int32_t) * cap); dist = (uint8_t *)realloc(dist, cap); }
      cols[c] = j; dist[c] = (uint8_t)d; c++;
    }
    nb->cols[i] = cols; nb->dist[i] = dist; nb->cnt[i] = c;
  }
  return NULL;
}

nbr_t *exnbr_create(const uint8_t *chars, const int *lens, int n, int threshold, int include_self, int n_threads) {
  if (!chars || !lens || n < 1 || threshold < 0 || threshold > 8) return NULL;
  nbr_t *nb = (nbr_t *)calloc(1, sizeof(nbr_t));
  nb->n = n; nb->threshold = threshold; nb->include_self = include_self; nb->chars = chars; nb->lens = lens;
  nb->mask = (uint64_t *)calloc(n, sizeof(uint64_t));
  nb->cols = (int32_t **)calloc(n, sizeof(int32_t *));
  nb->dist = (uint8_t **)calloc(n, sizeof(uint8_t *));
  nb->cnt = (int *)calloc(n, sizeof(int));
  nb->row_ptr = (int64_t *)calloc((size_t)n + 1, sizeof(int64_t));
  for (int i = 0; i < n; i++)
    for (int k = 0; k < lens[i]; k++) nb->mask[i] |= 1ull << (chars[(size_t)i * STRIDE + k] & 63);
  int T = n_threads < 1 ? 1 : (n_threads > 64 ? 64 : n_threads);
  pthread_t th[64];
  job_t jobs[64];
  for (int t = 0; t < T; t++) { jobs[t].nb = nb; jobs[t].t = t; jobs[t].T = T; pthread_create(&th[t], NULL, worker, &jobs[t]); }
  for (int t = 0; t < T; t++) pthread_join(th[t], NULL);
  for (int i = 0; i < n; i++) nb->row_ptr[i + 1] = nb->row_ptr[i] + nb->cnt[i];
  return nb;
}

int64_t exnbr_nnz(const nbr_t *nb) { return nb ? nb->row_ptr[nb->n] : -1; }

/* row_ptr [n+1], val_ids [nnz], probs [nnz], max_exp_factor [n] */
void exnbr_get(const nbr_t *nb, int64_t *row_ptr, int32_t *val_ids, double *probs, double *mef) {
  const int n = nb->n;
  memcpy(row_ptr, nb->row_ptr, sizeof(int64_t) * ((size_t)n + 1));
  double max_d = 0.0;
  for (int i = 0; i < n; i++)
    for (int k = 0; k < nb->cnt[i]; k++) if ((double)nb->dist[i][k] > max_d) max_d = (double)nb->dist[i][k];
  for (int i = 0; i < n; i++) {
    double sum = 0.0, min_d = INFINITY;
    int64_t o = nb->row_ptr[i];
    for (int k = 0; k < nb->cnt[i]; k++) {
      val_ids[o + k] = nb->cols[i][k];
      probs[o + k] = exp(-(double)nb->dist[i][k]);
      sum += probs[o + k];
      if ((double)nb->dist[i][k] < min_d) min_d = (double)nb->dist[i][k];
    }
    for (int k = 0; k < nb->cnt[i]; k++) probs[o + k] /= sum;
    mef[i] = isfinite(min_d) ? exp(-0.5 * min_d / max_d) : 0.0; /* AttributeIndex.R:135 */
  }
}

void exnbr_destroy(nbr_t *nb) {
  if (!nb) return;
  for (int i = 0; i < nb->n; i++) { free(nb->cols[i]); free(nb->dist[i]); }
  free(nb->cols); free(nb->dist); free(nb->cnt); free(nb->mask); free(nb->row_ptr); free(nb);
}
