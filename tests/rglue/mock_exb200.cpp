// Test double of libexb200.so for the CPU test of the R glue (tests/test_rglue.py): records what the
// glue hands over, follows the save schedule of src/sample.cpp:171-197, fills the history with
// recognisable values and polls the callbacks like the real library.  No sampling happens here.
#include <cstring>
#include <string>
#include <vector>

#include "exb200.h"

struct exb200_handle {
  int N, A, F, K, iteration;
  exb200_clust cl;
  std::vector<int> links, ent_ids, ent_attrs, rec_dist;
  std::vector<double> theta, concs;
  std::string err;
};

static exb200_model g_seen;            // what exb200_create received (scalars; pointers are not kept)
static std::vector<int> g_seen_rec_attrs, g_seen_file_ids, g_seen_close_ids;
static std::vector<exb200_attr> g_seen_attrs;
static std::string g_err;

extern "C" {

int exb200_abi_version(void) { return EXB200_ABI_VERSION; }
const char* exb200_last_error(const exb200_handle* h) { return h ? h->err.c_str() : g_err.c_str(); }

int exb200_create(const exb200_model* m, int, exb200_handle** out) {
  if (m->n_records < 2) { g_err = "mock: need two records"; return EXB200_ERR_INVALID; }
  g_seen = *m;
  g_seen_rec_attrs.assign(m->rec_attrs, m->rec_attrs + (size_t)m->n_attrs * m->n_records);
  g_seen_file_ids.assign(m->file_ids, m->file_ids + m->n_records);
  g_seen_attrs.assign(m->attrs, m->attrs + m->n_attrs);
  g_seen_close_ids.clear();
  for (int a = 0; a < m->n_attrs; a++)
    if (!m->attrs[a].is_constant)
      g_seen_close_ids.insert(g_seen_close_ids.end(), m->attrs[a].close_val_ids,
                              m->attrs[a].close_val_ids + m->attrs[a].close_row_ptr[m->attrs[a].domain_size]);
  exb200_handle* h = new exb200_handle();
  h->N = m->n_records; h->A = m->n_attrs; h->F = m->n_files; h->K = m->n_entities; h->iteration = m->iteration;
  h->cl = m->clust;
  h->links.assign(m->links, m->links + h->N);
  h->ent_ids.assign(m->ent_ids, m->ent_ids + h->K);
  h->ent_attrs.assign(m->ent_attrs, m->ent_attrs + (size_t)h->A * h->K);
  h->rec_dist.assign(m->rec_distortions, m->rec_distortions + (size_t)h->A * h->N);
  h->theta.assign(m->distort_probs, m->distort_probs + (size_t)h->F * h->A);
  for (int a = 0; a < h->A; a++) h->concs.push_back(m->attrs[a].distort_dist_conc);
  *out = h;
  return EXB200_OK;
}

int exb200_run(exb200_handle* h, int32_t n_samples, int32_t thin, int32_t burnin, exb200_history* out,
               int (*abort_cb)(void*), void (*progress_cb)(void*, int32_t), void* ctx) {
  const int N = h->N, FA = h->F * h->A, initial = h->iteration;
  int sample = 0;
  out->n_saved = 0;
  while (sample < n_samples) {
    h->iteration++;                                     // one "sweep"
    const int completed = h->iteration - initial;
    if (completed >= burnin && (completed - burnin) % thin == 0) {
      for (int r = 0; r < N; r++) out->links[(size_t)sample * N + r] = r % 7;           // 0-based labels
      out->n_linked_ents[sample] = 1000 + completed;
      for (int k = 0; k < FA; k++) { out->distort_probs[(size_t)sample * FA + k] = completed + 0.001 * k; out->distort_counts[(size_t)sample * FA + k] = 10 * completed + k; }
      for (int a = 0; a < h->A; a++) out->distort_dist_concs[(size_t)sample * h->A + a] = h->concs[a];
      out->clust_params[2 * sample] = 0.5 + completed; out->clust_params[2 * sample + 1] = 0.25 + completed;
      sample++;
      out->n_saved = sample;
    }
    if (abort_cb && abort_cb(ctx)) { h->err = "aborted by the caller"; return EXB200_ERR_ABORTED; }
    if (progress_cb) progress_cb(ctx, completed);
  }
  return EXB200_OK;
}

int exb200_get_state(exb200_handle* h, exb200_state* out) {
  out->iteration = h->iteration;
  out->n_entities = h->K;
  std::memcpy(out->links, h->links.data(), sizeof(int) * h->N);
  std::memcpy(out->ent_ids, h->ent_ids.data(), sizeof(int) * h->K);
  std::memcpy(out->ent_attrs, h->ent_attrs.data(), sizeof(int) * (size_t)h->A * h->K);
  std::memcpy(out->rec_distortions, h->rec_dist.data(), sizeof(int) * (size_t)h->A * h->N);
  std::memcpy(out->distort_probs, h->theta.data(), sizeof(double) * (size_t)h->F * h->A);
  for (int k = 0; k < h->F * h->A; k++) out->distort_counts[k] = k;
  std::memcpy(out->distort_dist_concs, h->concs.data(), sizeof(double) * h->A);
  out->clust = h->cl;
  return EXB200_OK;
}

void exb200_destroy(exb200_handle* h) { delete h; }

// what the glue handed to exb200_create, for the test
const exb200_model* mock_seen_model(void) { return &g_seen; }
const exb200_attr* mock_seen_attrs(void) { return g_seen_attrs.data(); }
const int* mock_seen_rec_attrs(void) { return g_seen_rec_attrs.data(); }
const int* mock_seen_file_ids(void) { return g_seen_file_ids.data(); }
const int* mock_seen_close_ids(void) { return g_seen_close_ids.data(); }

} // extern "C"
