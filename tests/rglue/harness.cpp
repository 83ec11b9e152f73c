// Harness of the R-glue test (tests/test_rglue.py).  Compiled against the stand-in Rcpp headers of
// oracle/ref_shim/ together with
//   * the reference's OWN src/sample.cpp (its entry renamed with -Dsample=reference_sample) and the
//     ten hot-path sources it needs, and
//   * exchanger_b200/rglue/sample_b200.cpp, linked with a test double of libexb200 (CPU test) or
//     with the real library (GPU test).
// It builds the `ExchangERModel` S4 object exchanger() would hand to .sample() from a flat model,
// runs either entry, and dumps the returned `ExchangERFit` as JSON so that the test can compare the
// two objects slot for slot.
#include <RcppArmadillo.h>
#include <progress.hpp>

#include <cstdint>
#include <cstdio>
#include <string>

#include "exb200.h"
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
static void segv_handler(int) { void* bt[40]; int n = backtrace(bt, 40); backtrace_symbols_fd(bt, n, 2); _exit(3); }

using namespace Rcpp;

Rcpp::S4 reference_sample(Rcpp::S4 init_state, int n_samples, int thin_interval, int burnin_interval); // src/sample.cpp
Rcpp::S4 sample(Rcpp::S4 init_state, int n_samples, int thin_interval, int burnin_interval);           // the glue
extern "C" void exref_seed(uint64_t seed);                                                             // oracle/ref_shim/exref.cpp

static SEXP rv(const char* klass, const char* a, double x, const char* b, double y) {
  S4 o(klass); o.slot(a) = x; o.slot(b) = y; return o.node;
}
static CharacterVector numbered(const char* prefix, int n) {
  CharacterVector v(n);
  for (int i = 0; i < n; i++) v[i] = std::string(prefix) + std::to_string(i + 1);
  return v;
}

// the S4 object of R/ExchangERModel.R:166-182 (1-based ids, NA_INTEGER for missing values, dimnames)
static S4 make_model(const exb200_model* m) {
  const int N = m->n_records, A = m->n_attrs, F = m->n_files, K = m->n_entities;
  List attr_params, attr_indices;
  CharacterVector attr_names = numbered("attr", A);
  for (int a = 0; a < A; a++) {
    const exb200_attr& at = m->attrs[a];
    const int D = at.domain_size;
    S4 spec(at.is_constant ? "CategoricalAttribute" : "Attribute");
    spec.slot("exclude_entity_value") = make_lgl(at.exclude_entity_value != 0);
    if (at.entity_dist_kind == EXB200_ENTDIST_DIRICHLET) {
      S4 dr("DirichletRV"); dr.slot("alpha") = make_real({at.entity_dist[0]});
      spec.slot("entity_dist_prior") = dr.node;
    } else if (at.entity_dist) spec.slot("entity_dist_prior") = make_real(std::vector<double>(at.entity_dist, at.entity_dist + D));
    else spec.slot("entity_dist_prior") = SEXP();
    if (at.distort_prob_shape1 > 0) spec.slot("distort_prob_prior") = rv("BetaRV", "shape1", at.distort_prob_shape1, "shape2", at.distort_prob_shape2);
    else spec.slot("distort_prob_prior") = make_real({m->distort_probs[(size_t)F * a]});
    S4 dp("DirichletProcess");
    if (at.conc_prior_shape > 0) dp.slot("alpha") = rv("GammaRV", "shape", at.conc_prior_shape, "rate", at.conc_prior_rate);
    else dp.slot("alpha") = make_real({at.distort_dist_conc});
    spec.slot("distort_dist_prior") = dp.node;
    attr_params.push_back(spec.node, attr_names[a]);

    S4 idx("AttributeIndex");
    idx.slot("domain") = numbered("v", D);
    idx.slot("probs") = make_real(std::vector<double>(at.probs, at.probs + D));
    List cvs;
    if (!at.is_constant) {
      for (int v = 0; v < D; v++) {
        List cv;
        std::vector<int> ids; std::vector<double> pr;
        for (int64_t j = at.close_row_ptr[v]; j < at.close_row_ptr[v + 1]; j++) { ids.push_back(at.close_val_ids[j] + 1); pr.push_back(at.close_probs[j]); }
        cv.push_back(make_int(ids), "valIds");
        cv.push_back(make_real(pr), "probs");
        cvs.push_back(cv.node);
      }
      idx.slot("max_exp_factor") = make_real(std::vector<double>(at.max_exp_factor, at.max_exp_factor + D));
    } else idx.slot("max_exp_factor") = make_real({});
    idx.slot("close_values") = cvs.node;
    attr_indices.push_back(idx.node, attr_names[a]);
  }
  CharacterVector rec_ids = numbered("rec", N);
  IntegerMatrix rec_attrs(A, N), rec_dist(A, N), ent_attrs(A, K);
  for (size_t k = 0; k < (size_t)A * N; k++) {
    rec_attrs[k] = m->rec_attrs[k] < 0 ? NA_INTEGER : m->rec_attrs[k] + 1;
    rec_dist[k] = m->rec_distortions[k];
  }
  for (size_t k = 0; k < (size_t)A * K; k++) ent_attrs[k] = m->ent_attrs[k] + 1;
  rownames(rec_attrs) = attr_names; colnames(rec_attrs) = rec_ids;
  rownames(rec_dist) = attr_names; colnames(rec_dist) = rec_ids;
  rownames(ent_attrs) = attr_names;
  IntegerVector file_ids(N), ent_ids(K), links(N);
  for (int r = 0; r < N; r++) { file_ids[r] = m->file_ids[r] + 1; links[r] = m->links[r] + 1; }
  for (int e = 0; e < K; e++) ent_ids[e] = m->ent_ids[e] + 1;
  file_ids.attr("levels") = numbered("file", F);
  links.attr("names") = rec_ids;
  NumericMatrix theta(F, A);
  for (int k = 0; k < F * A; k++) theta[k] = m->distort_probs[k];
  rownames(theta) = numbered("file", F); colnames(theta) = attr_names;
  NumericVector concs(A);
  for (int a = 0; a < A; a++) concs[a] = m->attrs[a].distort_dist_conc;
  concs.attr("names") = attr_names;

  const exb200_clust& c = m->clust;
  const bool py = c.kind == EXB200_CLUST_PITMAN_YOR;
  S4 prior(py ? "PitmanYorRP" : "GeneralizedCouponRP"), params(py ? "PitmanYorRP" : "GeneralizedCouponRP");
  if (py) {
    params.slot("alpha") = c.alpha; params.slot("d") = c.d;
    if (c.alpha_prior_shape > 0) prior.slot("alpha") = rv("GammaRV", "shape", c.alpha_prior_shape, "rate", c.alpha_prior_rate); else prior.slot("alpha") = c.alpha;
    if (c.d_prior_shape1 > 0) prior.slot("d") = rv("BetaRV", "shape1", c.d_prior_shape1, "shape2", c.d_prior_shape2); else prior.slot("d") = c.d;
  } else {
    params.slot("m") = c.m;
    params.slot("kappa") = c.kind == EXB200_CLUST_COUPON ? R_PosInf : c.kappa;
    if (c.m_prior_size > 0) prior.slot("m") = rv("ShiftedNegBinomRV", "size", c.m_prior_size, "prob", c.m_prior_prob); else prior.slot("m") = c.m;
    if (c.kind == EXB200_CLUST_COUPON) prior.slot("kappa") = R_PosInf;
    else if (c.kappa_prior_shape > 0) prior.slot("kappa") = rv("GammaRV", "shape", c.kappa_prior_shape, "rate", c.kappa_prior_rate);
    else prior.slot("kappa") = c.kappa;
  }

  S4 model("ExchangERModel");
  model.slot("iteration") = (int)m->iteration;
  model.slot("attr_params") = attr_params;
  model.slot("file_ids") = file_ids;
  model.slot("rec_ids") = rec_ids;
  model.slot("rec_attrs") = rec_attrs;
  model.slot("rec_distortions") = rec_dist;
  model.slot("ent_attrs") = ent_attrs;
  model.slot("ent_ids") = ent_ids;
  model.slot("links") = links;
  model.slot("distort_probs") = theta;
  model.slot("distort_dist_concs") = concs;
  model.slot("clust_params") = params;
  model.slot("clust_prior") = prior;
  model.slot("attr_indices") = attr_indices;
  return model;
}

// ---- JSON dump of an R object tree (plain string building: no iostreams across library boundaries)
typedef std::string Out;
static void put(Out& o, long long v) { char b[32]; snprintf(b, sizeof b, "%lld", v); o += b; }
static void put(Out& o, double v) { char b[40]; snprintf(b, sizeof b, "%.17g", v); o += b; }
static void dump_strs(Out& o, const std::vector<std::string>& v, size_t n) {
  o += "[";
  for (size_t i = 0; i < n; i++) { o += i ? ",\"" : "\""; o += v[i]; o += "\""; }
  o += "]";
}
static void dump(Out& o, const SEXP& n, int max_values) {
  if (!n) { o += "null"; return; }
  const char* tname = n->type == S4SXP ? "S4" : n->type == LGLSXP ? "logical" : n->type == INTSXP ? "integer"
                      : n->type == REALSXP ? "double" : n->type == STRSXP ? "character" : n->type == VECSXP ? "list" : "NULL";
  o += "{\"type\":\""; o += tname; o += "\"";
  if (n->type == S4SXP) { o += ",\"class\":\""; o += n->klass; o += "\""; }
  const size_t len = n->type == INTSXP || n->type == LGLSXP ? n->ints.size() : n->type == REALSXP ? n->reals.size()
                     : n->type == STRSXP ? n->strs.size() : n->items.size();
  if (n->type != S4SXP) { o += ",\"len\":"; put(o, (long long)len); }
  if (n->nrow || n->ncol) { o += ",\"dim\":["; put(o, (long long)n->nrow); o += ","; put(o, (long long)n->ncol); o += "]"; }
  if (n->type == VECSXP) {
    o += ",\"names\":"; dump_strs(o, n->names, n->names.size());
    o += ",\"items\":[";
    for (size_t i = 0; i < n->items.size(); i++) { if (i) o += ","; dump(o, n->items[i], max_values); }
    o += "]";
  }
  const size_t shown = (long long)len <= max_values ? len : (size_t)max_values;
  if (n->type == INTSXP || n->type == LGLSXP) {
    o += ",\"values\":[";
    for (size_t i = 0; i < shown; i++) { if (i) o += ","; if (n->ints[i] == NA_INTEGER) o += "null"; else put(o, (long long)n->ints[i]); }
    o += "]";
  } else if (n->type == REALSXP) {
    o += ",\"values\":[";
    for (size_t i = 0; i < shown; i++) {
      if (i) o += ",";
      if (std::isfinite(n->reals[i])) put(o, n->reals[i]); else o += n->reals[i] > 0 ? "\"Inf\"" : "\"-Inf\"";
    }
    o += "]";
  } else if (n->type == STRSXP) { o += ",\"values\":"; dump_strs(o, n->strs, shown); }
  if (!n->attrs.empty()) {
    o += n->type == S4SXP ? ",\"slots\":{" : ",\"attrs\":{";
    bool first = true;
    for (auto& kv : n->attrs) { o += first ? "\"" : ",\""; o += kv.first; o += "\":"; dump(o, kv.second, max_values); first = false; }
    o += "}";
  }
  o += "}";
}

static std::string g_out;

extern "C" {

// which = 0: the reference's sample(); 1: the glue.  abort_after >= 0: Progress::check_abort() answers
// true at its (abort_after + 1)-th call.  Returns the ExchangERFit as JSON, or {"error": ...}.
const char* rglue_run(const exb200_model* m, int n_samples, int thin, int burnin, int which, int abort_after,
                      uint64_t seed, int max_values) {
  Out o;
  signal(SIGSEGV, segv_handler);
  try {
    exref_seed(seed);
    Progress::abort_after() = abort_after;
    Progress::increments() = 0;
    S4 model = make_model(m);
    S4 fit = which == 0 ? reference_sample(model, n_samples, thin, burnin) : sample(model, n_samples, thin, burnin);
    o += "{\"increments\":"; put(o, (long long)Progress::increments()); o += ",\"fit\":";
    dump(o, fit.node, max_values);
    o += "}";
  } catch (std::exception& e) {
    o = "{\"error\":\""; o += e.what(); o += "\"}";
  }
  Progress::abort_after() = -1;
  g_out = o;
  return g_out.c_str();
}

// the Philox key the glue derives from R's RNG after set.seed (here: exref_seed(seed))
uint64_t rglue_seed_for(uint64_t seed) {
  exref_seed(seed);
  const uint64_t hi = (uint64_t)(R::unif_rand() * 4294967296.0), lo = (uint64_t)(R::unif_rand() * 4294967296.0);
  return (hi << 32) | (lo & 0xFFFFFFFFull);
}

} // extern "C"
