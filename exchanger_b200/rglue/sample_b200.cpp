// sample_b200.cpp - the R-side binding of libexb200.so: a drop-in replacement for the reference's
// src/sample.cpp.  Same exported name and signature (`.sample`, R/RcppExports.R:12-14,
// src/RcppExports.cpp:35-48), same `ExchangERFit` S4 result, so exchanger(), run_inference(),
// extract(), smp_clusters() ... of the package work unchanged.
//
//   readS4State        (src/sample.cpp:71-105)  -> flatten_model(): S4 slots -> exb200_model (no copies of
//                                                  the big matrices: R's column-major layout IS the ABI's)
//   the sampling loop  (src/sample.cpp:171-202) -> exb200_run() with the same save schedule; abort_cb /
//                                                  progress_cb play Progress::check_abort / p.increment
//   history matrices   (src/sample.cpp:133-166, 208-216) -> history_to_R(): same element names, order,
//                                                  column names ("attr[file]" in f + a F order)
//   buildFinalS4State  (src/sample.cpp:14-63)   -> final_state_to_R()
//
// Built in the package with  PKG_CPPFLAGS = -I<repo>/include,  PKG_LIBS = -L<repo>/exchanger_b200 -lexb200
// (INTEGRATION.md).  R is not in the build image of this repository: the file is compiled there against
// the stand-in headers of oracle/ref_shim/ together with the reference's own src/sample.cpp, and
// tests/test_rglue.py checks slot for slot that both produce the same ExchangERFit structure.
#include <RcppArmadillo.h>
#include <progress.hpp>
#include <progress_bar.hpp>

#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

#include "exb200.h"

namespace {

// Everything the flat model points to that is not an R vector of the S4 object itself.
struct FlatModel {
  exb200_model m;
  std::vector<exb200_attr> attrs;
  std::vector<std::vector<int64_t>> row_ptr;
  std::vector<std::vector<int32_t>> val_ids;
  std::vector<std::vector<double>> close_probs;
  std::vector<std::vector<double>> dirichlet_alpha;
  std::vector<Rcpp::NumericVector> keep_num; // R vectors whose storage the model points into
  Rcpp::IntegerMatrix rec_attrs, rec_distortions, ent_attrs;
  Rcpp::IntegerVector file_ids, links, ent_ids;
  Rcpp::NumericMatrix distort_probs;
  bool kappa_random = false, m_random = false, alpha_random = false, d_random = false;
};

template <class V> auto *data_of(V &v) { return v.size() ? &v[0] : nullptr; }

// S4 slots -> POD model (what readS4State reads, src/sample.cpp:71-105)
void flatten_model(const Rcpp::S4 &init_state, FlatModel &fm) {
  Rcpp::List attr_params = init_state.slot("attr_params"), attr_indices = init_state.slot("attr_indices");
  Rcpp::IntegerMatrix rec_attrs_R = init_state.slot("rec_attrs");       // A x N, 1-based, NA
  fm.rec_attrs = rec_attrs_R - 1;                                      // NA stays NA_INTEGER (< 0), sample.cpp:92
  fm.rec_distortions = init_state.slot("rec_distortions");
  Rcpp::IntegerMatrix ent_attrs_R = init_state.slot("ent_attrs");
  fm.ent_attrs = ent_attrs_R - 1;                                      // sample.cpp:93
  fm.ent_ids = init_state.slot("ent_ids");
  fm.links = init_state.slot("links");
  Rcpp::IntegerVector file_ids_R = init_state.slot("file_ids");
  fm.file_ids = file_ids_R - 1;                                        // sample.cpp:90
  fm.distort_probs = init_state.slot("distort_probs");                 // F x A
  Rcpp::NumericVector concs = init_state.slot("distort_dist_concs");
  const int A = fm.rec_attrs.nrow(), N = fm.rec_attrs.ncol(), F = fm.distort_probs.nrow();

  fm.attrs.assign(A, exb200_attr());
  fm.row_ptr.resize(A); fm.val_ids.resize(A); fm.close_probs.resize(A); fm.dirichlet_alpha.resize(A);
  for (int a = 0; a < A; a++) {
    Rcpp::S4 spec = attr_params[a], idx = attr_indices[a];
    exb200_attr &t = fm.attrs[a];
    Rcpp::NumericVector probs = idx.slot("probs");
    fm.keep_num.push_back(probs);
    t.domain_size = probs.size();
    t.probs = data_of(probs);
    t.exclude_entity_value = (bool)spec.slot("exclude_entity_value") ? 1 : 0;
    Rcpp::List cv = idx.slot("close_values");                          // R orientation: row = entity value
    t.is_constant = cv.size() == 0;                                    // readAttributeIndex, attribute_index.cpp:118-149
    if (!t.is_constant) {
      fm.row_ptr[a].push_back(0);
      for (int v = 0; v < cv.size(); v++) {
        Rcpp::List e = cv[v];
        Rcpp::IntegerVector ids = e["valIds"];
        Rcpp::NumericVector p = e["probs"];
        for (int j = 0; j < ids.size(); j++) { fm.val_ids[a].push_back(ids[j] - 1); fm.close_probs[a].push_back(p[j]); }
        fm.row_ptr[a].push_back((int64_t)fm.val_ids[a].size());
      }
      t.close_row_ptr = fm.row_ptr[a].data();
      t.close_val_ids = fm.val_ids[a].data();
      t.close_probs = fm.close_probs[a].data();
      Rcpp::NumericVector mef = idx.slot("max_exp_factor");
      fm.keep_num.push_back(mef);
      t.max_exp_factor = data_of(mef);
    }
    // entity-value distribution prior: NULL / numeric pmf / DirichletRV with scalar alpha (entities.cpp:181-211)
    SEXP edp = spec.slot("entity_dist_prior");
    t.entity_dist = nullptr;
    t.entity_dist_kind = EXB200_ENTDIST_FIXED;
    if (TYPEOF(edp) == REALSXP) {
      Rcpp::NumericVector v(edp);
      fm.keep_num.push_back(v);
      t.entity_dist = data_of(v);
    } else if (TYPEOF(edp) == S4SXP) {
      Rcpp::S4 dr(edp);
      if (!dr.is("DirichletRV")) Rcpp::stop("Invalid entity_dist_prior");       // entities.cpp:195
      Rcpp::NumericVector alpha = dr.slot("alpha");
      if (alpha.size() != 1) Rcpp::stop("Invalid entity_dist_prior");            // vector alpha: entities.cpp:192-196
      fm.dirichlet_alpha[a].assign((size_t)t.domain_size, alpha[0]);
      t.entity_dist = fm.dirichlet_alpha[a].data();
      t.entity_dist_kind = EXB200_ENTDIST_DIRICHLET;
    }
    // distortion probability prior: BetaRV or a fixed probability (distortion_probs.cpp:4-21)
    SEXP dpp = spec.slot("distort_prob_prior");
    t.distort_prob_shape1 = t.distort_prob_shape2 = 0.0;
    if (TYPEOF(dpp) == S4SXP) {
      Rcpp::S4 b(dpp);
      t.distort_prob_shape1 = b.slot("shape1");
      t.distort_prob_shape2 = b.slot("shape2");
    }
    // distortion distribution prior: DirichletProcess(alpha = number | GammaRV) (distort_dist_concs.cpp:3-19)
    Rcpp::S4 ddp = spec.slot("distort_dist_prior");
    SEXP alpha = ddp.slot("alpha");
    t.distort_dist_conc = concs[a];
    t.conc_prior_shape = t.conc_prior_rate = 0.0;
    if (TYPEOF(alpha) == S4SXP) {
      Rcpp::S4 g(alpha);
      t.conc_prior_shape = g.slot("shape");
      t.conc_prior_rate = g.slot("rate");
    }
  }

  exb200_model &m = fm.m;
  m = exb200_model();
  m.n_records = N; m.n_attrs = A; m.n_files = F; m.n_entities = fm.ent_ids.size();
  m.iteration = init_state.slot("iteration");
  m.rec_attrs = data_of(fm.rec_attrs);
  m.rec_distortions = data_of(fm.rec_distortions);
  m.file_ids = data_of(fm.file_ids);
  m.links = data_of(fm.links);
  m.ent_ids = data_of(fm.ent_ids);
  m.ent_attrs = data_of(fm.ent_attrs);
  m.distort_probs = data_of(fm.distort_probs);
  m.attrs = fm.attrs.data();

  // cluster prior and current parameter values (read_clust_params, clust_params.cpp:259-281)
  Rcpp::S4 prior = init_state.slot("clust_prior"), params = init_state.slot("clust_params");
  exb200_clust &c = m.clust;
  if (prior.is("PitmanYorRP")) {
    c.kind = EXB200_CLUST_PITMAN_YOR;
    c.alpha = params.slot("alpha");
    c.d = params.slot("d");
    SEXP a = prior.slot("alpha"), d = prior.slot("d");
    if (TYPEOF(a) == S4SXP) { Rcpp::S4 g(a); c.alpha_prior_shape = g.slot("shape"); c.alpha_prior_rate = g.slot("rate"); fm.alpha_random = true; }
    if (TYPEOF(d) == S4SXP) { Rcpp::S4 b(d); c.d_prior_shape1 = b.slot("shape1"); c.d_prior_shape2 = b.slot("shape2"); fm.d_random = true; }
  } else if (prior.is("GeneralizedCouponRP")) {
    c.m = params.slot("m");
    c.kappa = params.slot("kappa");
    SEXP mp = prior.slot("m"), kp = prior.slot("kappa");
    if (TYPEOF(kp) == S4SXP) { Rcpp::S4 g(kp); c.kappa_prior_shape = g.slot("shape"); c.kappa_prior_rate = g.slot("rate"); fm.kappa_random = true; }
    if (TYPEOF(mp) == S4SXP) { Rcpp::S4 nb(mp); c.m_prior_size = nb.slot("size"); c.m_prior_prob = nb.slot("prob"); fm.m_random = true; }
    // kappa = Inf is the plain coupon-collector prior (CouponClustParams, clust_params.cpp:262-266)
    c.kind = (!fm.kappa_random && !R_FINITE(c.kappa)) ? EXB200_CLUST_COUPON : EXB200_CLUST_GEN_COUPON;
  } else {
    Rcpp::stop("Unrecognized clust_prior");                                      // clust_params.cpp:279
  }
  // set.seed() keeps fixing the chain: the Philox key is drawn from R's RNG (RNGScope is active in
  // the generated wrapper, src/RcppExports.cpp:40)
  const uint64_t hi = (uint64_t)(R::unif_rand() * 4294967296.0), lo = (uint64_t)(R::unif_rand() * 4294967296.0);
  m.seed = (hi << 32) | (lo & 0xFFFFFFFFull);
}

struct HistoryBuffers { // sample-major rows, as the ABI fills them
  std::vector<int32_t> links, counts, n_linked;
  std::vector<double> theta, concs, clust;
};

// history list of the ExchangERFit (src/sample.cpp:133-166, 208-216)
Rcpp::List history_to_R(const Rcpp::S4 &init_state, const FlatModel &fm, const HistoryBuffers &hb, int n_samples, int n_saved) {
  const int N = fm.m.n_records, A = fm.m.n_attrs, F = fm.m.n_files, FA = F * A;
  Rcpp::CharacterVector rec_ids = init_state.slot("rec_ids");
  const Rcpp::IntegerMatrix &rec_attrs = init_state.slot("rec_attrs");
  Rcpp::CharacterVector attrs_names = rownames(rec_attrs);
  const Rcpp::IntegerVector &file_ids = init_state.slot("file_ids");
  Rcpp::CharacterVector file_ids_names = file_ids.attr("levels");

  Rcpp::IntegerMatrix hist_links(n_samples, N);
  colnames(hist_links) = rec_ids;
  Rcpp::CharacterVector probs_names(FA);
  for (int i = 0; i < F; ++i)
    for (int j = 0; j < A; ++j)
      probs_names[i + j * F] = Rcpp::as<std::string>(attrs_names[j]) + "[" + Rcpp::as<std::string>(file_ids_names[i]) + "]";
  Rcpp::NumericMatrix hist_distort_probs(n_samples, FA);
  colnames(hist_distort_probs) = probs_names;
  Rcpp::NumericMatrix hist_distort_dist_concs(n_samples, A);
  colnames(hist_distort_dist_concs) = attrs_names;
  Rcpp::IntegerMatrix hist_n_linked_ents(n_samples, 1);
  colnames(hist_n_linked_ents) = Rcpp::CharacterVector::create("n_linked_ents");
  Rcpp::IntegerMatrix hist_distort_counts(n_samples, FA);
  colnames(hist_distort_counts) = probs_names;

  // only the random cluster parameters, named as ClustParams::to_R_vec does (clust_params.cpp:158-211)
  std::vector<std::string> cp_names;
  std::vector<int> cp_cols;
  if (fm.m.clust.kind == EXB200_CLUST_PITMAN_YOR) {
    if (fm.alpha_random) { cp_names.push_back("alpha"); cp_cols.push_back(0); }
    if (fm.d_random) { cp_names.push_back("d"); cp_cols.push_back(1); }
  } else {
    if (fm.kappa_random) { cp_names.push_back("kappa"); cp_cols.push_back(0); }
    if (fm.m_random) { cp_names.push_back("m"); cp_cols.push_back(1); }
  }
  Rcpp::NumericMatrix hist_clust_params;
  if (!cp_names.empty()) {
    hist_clust_params = Rcpp::NumericMatrix(n_samples, (int)cp_names.size());
    colnames(hist_clust_params) = Rcpp::CharacterVector(cp_names.begin(), cp_names.end());
  }

  // rows that were never reached (interrupt) stay zero, as the reference leaves them
  for (int s = 0; s < n_saved; s++) {
    // labels come back as "smallest member record, 0-based": + 1 makes them R-style ids (only the
    // partition matters downstream, src/point_estimate.cpp:26-40)
    for (int r = 0; r < N; r++) hist_links(s, r) = hb.links[(size_t)s * N + r] + 1;
    hist_n_linked_ents(s, 0) = hb.n_linked[s];
    for (int k = 0; k < FA; k++) {
      hist_distort_probs(s, k) = hb.theta[(size_t)s * FA + k];
      hist_distort_counts(s, k) = hb.counts[(size_t)s * FA + k];
    }
    for (int a = 0; a < A; a++) hist_distort_dist_concs(s, a) = hb.concs[(size_t)s * A + a];
    for (size_t k = 0; k < cp_cols.size(); k++) hist_clust_params(s, (int)k) = hb.clust[2 * (size_t)s + cp_cols[k]];
  }

  Rcpp::List history;
  history["links"] = hist_links;
  history["distort_probs"] = hist_distort_probs;
  history["distort_counts"] = hist_distort_counts;
  history["distort_dist_concs"] = hist_distort_dist_concs;
  history["n_linked_ents"] = hist_n_linked_ents;
  if (!cp_names.empty()) history["clust_params"] = hist_clust_params;
  return history;
}

// final `ExchangERModel` (buildFinalS4State, src/sample.cpp:14-63)
Rcpp::S4 final_state_to_R(const Rcpp::S4 &init_state, const FlatModel &fm, exb200_handle *h) {
  const int N = fm.m.n_records, A = fm.m.n_attrs, F = fm.m.n_files;
  std::vector<int32_t> links(N), ent_ids(N), ent_attrs((size_t)A * N), rec_dist((size_t)A * N), counts((size_t)F * A);
  std::vector<double> theta((size_t)F * A), concs(A);
  exb200_state st = exb200_state();
  st.links = links.data(); st.ent_ids = ent_ids.data(); st.ent_attrs = ent_attrs.data();
  st.rec_distortions = rec_dist.data(); st.distort_probs = theta.data(); st.distort_counts = counts.data();
  st.distort_dist_concs = concs.data();
  if (exb200_get_state(h, &st) != EXB200_OK) Rcpp::stop(exb200_last_error(h));
  const int K = st.n_entities;

  Rcpp::S4 final_state("ExchangERModel");
  // some slots are unchanged from init_state (sample.cpp:18-24)
  final_state.slot("attr_params") = init_state.slot("attr_params");
  final_state.slot("rec_attrs") = init_state.slot("rec_attrs");
  final_state.slot("file_ids") = init_state.slot("file_ids");
  final_state.slot("rec_ids") = init_state.slot("rec_ids");
  final_state.slot("attr_indices") = init_state.slot("attr_indices");
  final_state.slot("clust_prior") = init_state.slot("clust_prior");
  const Rcpp::IntegerMatrix &rec_attrs = final_state.slot("rec_attrs");
  const Rcpp::CharacterVector &rec_ids = final_state.slot("rec_ids");
  const Rcpp::IntegerVector &file_ids = init_state.slot("file_ids");

  final_state.slot("iteration") = (int)st.iteration;

  Rcpp::IntegerMatrix rec_distortions(A, N);
  for (size_t k = 0; k < (size_t)A * N; k++) rec_distortions[k] = rec_dist[k];
  rownames(rec_distortions) = rownames(rec_attrs);
  colnames(rec_distortions) = rec_ids;
  final_state.slot("rec_distortions") = rec_distortions;

  Rcpp::IntegerMatrix ent_attrs_R(A, K);                                 // 1-based again (entities.cpp:124)
  for (size_t k = 0; k < (size_t)A * K; k++) ent_attrs_R[k] = ent_attrs[k] + 1;
  rownames(ent_attrs_R) = rownames(rec_attrs);
  final_state.slot("ent_attrs") = ent_attrs_R;

  Rcpp::IntegerVector ent_ids_R(K);
  for (int e = 0; e < K; e++) ent_ids_R[e] = ent_ids[e] + 1;
  final_state.slot("ent_ids") = ent_ids_R;

  Rcpp::IntegerVector links_R(N);
  for (int r = 0; r < N; r++) links_R[r] = links[r] + 1;
  links_R.attr("names") = rec_ids;
  final_state.slot("links") = links_R;

  Rcpp::NumericMatrix distort_probs(F, A);
  for (size_t k = 0; k < (size_t)F * A; k++) distort_probs[k] = theta[k];
  colnames(distort_probs) = rownames(rec_attrs);
  const Rcpp::CharacterVector &file_names = file_ids.attr("levels");
  rownames(distort_probs) = file_names;
  final_state.slot("distort_probs") = distort_probs;

  Rcpp::NumericVector distort_dist_concs(A);
  for (int a = 0; a < A; a++) distort_dist_concs[a] = concs[a];
  distort_dist_concs.attr("names") = rownames(rec_attrs);
  final_state.slot("distort_dist_concs") = distort_dist_concs;

  // ClustParams::to_R (clust_params.cpp:135-156)
  if (st.clust.kind == EXB200_CLUST_PITMAN_YOR) {
    Rcpp::S4 cp("PitmanYorRP");
    cp.slot("alpha") = st.clust.alpha;
    cp.slot("d") = st.clust.d;
    final_state.slot("clust_params") = cp;
  } else {
    Rcpp::S4 cp("GeneralizedCouponRP");
    cp.slot("m") = (int)st.clust.m;                                   // m_ is an int in the reference (clust_params.cpp:138,146)
    cp.slot("kappa") = st.clust.kind == EXB200_CLUST_COUPON ? R_PosInf : st.clust.kappa;
    final_state.slot("clust_params") = cp;
  }
  return final_state;
}

int abort_cb(void *) { return Progress::check_abort() ? 1 : 0; }                  // src/sample.cpp:200
void progress_cb(void *ctx, int32_t) { static_cast<Progress *>(ctx)->increment(); } // src/sample.cpp:201

} // namespace

//[[Rcpp::export(.sample)]]
Rcpp::S4 sample(Rcpp::S4 init_state, int n_samples, int thin_interval = 1, int burnin_interval = 0) {
  FlatModel fm;
  flatten_model(init_state, fm);

  exb200_handle *h = nullptr;
  if (exb200_create(&fm.m, /*device=*/0, &h) != EXB200_OK) Rcpp::stop(exb200_last_error(nullptr));

  const int N = fm.m.n_records, A = fm.m.n_attrs, FA = fm.m.n_files * A;
  HistoryBuffers hb;
  hb.links.assign((size_t)n_samples * N, 0); hb.counts.assign((size_t)n_samples * FA, 0); hb.n_linked.assign(n_samples, 0);
  hb.theta.assign((size_t)n_samples * FA, 0.0); hb.concs.assign((size_t)n_samples * A, 0.0); hb.clust.assign((size_t)n_samples * 2, 0.0);
  exb200_history out = exb200_history();
  out.links = hb.links.data(); out.distort_probs = hb.theta.data(); out.distort_counts = hb.counts.data();
  out.distort_dist_concs = hb.concs.data(); out.n_linked_ents = hb.n_linked.data(); out.clust_params = hb.clust.data();

  Progress p(burnin_interval + thin_interval * n_samples, true);             // src/sample.cpp:130-131
  const int rc = exb200_run(h, n_samples, thin_interval, burnin_interval, &out, abort_cb, progress_cb, &p);
  if (rc != EXB200_OK && rc != EXB200_ERR_ABORTED) {                         // an interrupt returns the partial history
    const std::string msg = exb200_last_error(h);
    exb200_destroy(h);
    Rcpp::stop(msg);
  }

  Rcpp::S4 result("ExchangERFit");
  Rcpp::List history;
  Rcpp::S4 final_state;
  try {
    history = history_to_R(init_state, fm, hb, n_samples, out.n_saved);
    final_state = final_state_to_R(init_state, fm, h);
  } catch (...) {
    exb200_destroy(h);
    throw;
  }
  exb200_destroy(h);
  result.slot("history") = history;
  result.slot("state") = final_state;
  return result;
}
