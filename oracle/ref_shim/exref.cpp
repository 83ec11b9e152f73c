// C harness around the UNMODIFIED reference sources (/root/reference/src): builds the
// reference's own `State` (src/state.h:42-54) from the flat model of include/exb200.h in the
// order readS4State uses (src/sample.cpp:97-104) and exposes State::update() plus the
// externally linked weight functions (src/links.cpp:90,216,316; src/entities.cpp:236).
// Test infrastructure: it is the live reference for log-weight parity, posterior summaries
// and the CPU baseline ("kind": "reference").  R's RNG hooks (R::unif_rand, rgamma, rbeta,
// rnbinom) are provided here with std::mt19937_64, so streams differ from R's.
#include "state.h"
#include "../../include/exb200.h"
#include <chrono>
#include <random>

using namespace Rcpp;

// ---- functions with external linkage in the reference that have no header declaration
std::vector<ent_id> getPossibleLinks(rec_id rid, const Records &recs, const Entities &ents,
    const std::vector<std::shared_ptr<AbstractAttributeIndex> > &attr_indices);
double logLikelihoodWeightExisting(rec_id rid, ent_id eid, const std::unordered_set<rec_id> &linked_rids,
    const Records &recs, const Entities &ents, const DistortDistConcs &distort_dist_concs,
    const std::vector<std::shared_ptr<AbstractAttributeIndex> > &attr_indices);
double logLikelihoodWeightNew(rec_id rid, const Records &recs,
    const std::vector<std::shared_ptr<AbstractAttributeIndex> > &attr_indices, const Entities &ents);
double logWeightEntityValue(val_id e_vid, attr_id aid, const Records &recs, const DistortProbs &distort_probs,
    const DistortDistConcs &distort_dist_concs, const std::unordered_set<rec_id> &linked_rids,
    const AbstractAttributeIndex *index, IndexNonUniformDiscreteDist *distribution);

// ---- R's RNG, supplied by the harness
static thread_local std::mt19937_64 g_rng(12345);
namespace R {
double unif_rand() {
  // (0,1) open interval like R's unif_rand
  double u;
  do { u = (g_rng() >> 11) * (1.0 / 9007199254740992.0); } while (u <= 0.0);
  return u;
}
double rgamma(double shape, double scale) {
  if (shape <= 0.0) return 0.0;
  std::gamma_distribution<double> g(shape, scale);
  return g(g_rng);
}
double rbeta(double a, double b) {
  if (b == 0.0) return 1.0; // R: point mass at 1
  if (a == 0.0) return 0.0;
  double x = rgamma(a, 1.0), y = rgamma(b, 1.0);
  return x / (x + y);
}
double rnbinom(double size, double prob) {
  if (prob >= 1.0) return 0.0;
  double lam = rgamma(size, (1.0 - prob) / prob);
  if (lam <= 0.0) return 0.0;
  std::poisson_distribution<long long> p(lam);
  return (double)p(g_rng);
}
}

static SEXP beta_rv(double s1, double s2) {
  S4 o("BetaRV"); o.slot("shape1") = s1; o.slot("shape2") = s2; return o.node;
}
static SEXP gamma_rv(double shape, double rate) {
  S4 o("GammaRV"); o.slot("shape") = shape; o.slot("rate") = rate; return o.node;
}

struct ExRef {
  std::unique_ptr<State> state;
  int N, A, F;
};

extern "C" {

void exref_seed(uint64_t seed) { g_rng.seed(seed); }

void *exref_create(const exb200_model *m) {
  try {
    const int N = m->n_records, A = m->n_attrs, F = m->n_files, K = m->n_entities;
    List attr_params, attr_indices;
    for (int a = 0; a < A; a++) {
      const exb200_attr &at = m->attrs[a];
      const int D = at.domain_size;
      S4 spec(at.is_constant ? "CategoricalAttribute" : "Attribute");
      spec.slot("exclude_entity_value") = make_lgl(at.exclude_entity_value != 0);
      if (at.entity_dist_kind == EXB200_ENTDIST_DIRICHLET) {
        S4 dr("DirichletRV");
        dr.slot("alpha") = make_real({at.entity_dist[0]});
        spec.slot("entity_dist_prior") = dr.node;
      } else if (at.entity_dist) {
        spec.slot("entity_dist_prior") = make_real(std::vector<double>(at.entity_dist, at.entity_dist + D));
      } else {
        spec.slot("entity_dist_prior") = SEXP();
      }
      if (at.distort_prob_shape1 > 0)
        spec.slot("distort_prob_prior") = beta_rv(at.distort_prob_shape1, at.distort_prob_shape2);
      else
        spec.slot("distort_prob_prior") = make_real({0.0});
      S4 dp("DirichletProcess");
      if (at.conc_prior_shape > 0) dp.slot("alpha") = gamma_rv(at.conc_prior_shape, at.conc_prior_rate);
      else dp.slot("alpha") = make_real({at.distort_dist_conc});
      spec.slot("distort_dist_prior") = dp.node;
      attr_params.push_back(spec.node);

      S4 idx("AttributeIndex");
      idx.slot("domain") = make_str(std::vector<std::string>(D));
      idx.slot("probs") = make_real(std::vector<double>(at.probs, at.probs + D));
      List cvs;
      if (!at.is_constant) {
        for (int v = 0; v < D; v++) {
          List cv;
          std::vector<int> ids; std::vector<double> pr;
          for (int64_t j = at.close_row_ptr[v]; j < at.close_row_ptr[v + 1]; j++) {
            ids.push_back(at.close_val_ids[j] + 1); // R ids are 1-based
            pr.push_back(at.close_probs[j]);
          }
          cv.push_back(make_int(ids), "valIds");
          cv.push_back(make_real(pr), "probs");
          cvs.push_back(cv.node);
        }
        idx.slot("max_exp_factor") = make_real(std::vector<double>(at.max_exp_factor, at.max_exp_factor + D));
      } else {
        idx.slot("max_exp_factor") = make_real({});
      }
      idx.slot("close_values") = cvs.node;
      attr_indices.push_back(idx.node);
    }

    arma::imat rec_attrs(A, N), rec_dist(A, N), ent_attrs(A, K);
    for (int r = 0; r < N; r++)
      for (int a = 0; a < A; a++) {
        int v = m->rec_attrs[a + (size_t)A * r];
        rec_attrs(a, r) = v < 0 ? NA_INTEGER : v;
        rec_dist(a, r) = m->rec_distortions[a + (size_t)A * r];
      }
    for (int e = 0; e < K; e++)
      for (int a = 0; a < A; a++) ent_attrs(a, e) = m->ent_attrs[a + (size_t)A * e];
    arma::ivec fids(N), fsizes(F);
    for (int r = 0; r < N; r++) { fids[r] = m->file_ids[r]; fsizes[m->file_ids[r]] += 1; }
    arma::mat theta(F, A);
    for (int i = 0; i < F * A; i++) theta.d[i] = m->distort_probs[i];
    arma::vec concs(A);
    for (int a = 0; a < A; a++) concs[a] = m->attrs[a].distort_dist_conc;
    IntegerVector ent_ids(m->ent_ids, m->ent_ids + K), links(m->links, m->links + N);

    const exb200_clust &c = m->clust;
    S4 prior(c.kind == EXB200_CLUST_PITMAN_YOR ? "PitmanYorRP" : "GeneralizedCouponRP");
    S4 params(c.kind == EXB200_CLUST_PITMAN_YOR ? "PitmanYorRP" : "GeneralizedCouponRP");
    if (c.kind == EXB200_CLUST_PITMAN_YOR) {
      params.slot("alpha") = c.alpha; params.slot("d") = c.d;
      if (c.alpha_prior_shape > 0) prior.slot("alpha") = gamma_rv(c.alpha_prior_shape, c.alpha_prior_rate);
      else prior.slot("alpha") = c.alpha;
      if (c.d_prior_shape1 > 0) prior.slot("d") = beta_rv(c.d_prior_shape1, c.d_prior_shape2);
      else prior.slot("d") = c.d;
    } else {
      params.slot("m") = c.m; params.slot("kappa") = c.kappa;
      if (c.m_prior_size > 0) {
        S4 nb("ShiftedNegBinomRV"); nb.slot("size") = c.m_prior_size; nb.slot("prob") = c.m_prior_prob;
        prior.slot("m") = nb.node;
      } else prior.slot("m") = c.m;
      if (c.kind == EXB200_CLUST_COUPON) prior.slot("kappa") = R_PosInf;
      else if (c.kappa_prior_shape > 0) prior.slot("kappa") = gamma_rv(c.kappa_prior_shape, c.kappa_prior_rate);
      else prior.slot("kappa") = c.kappa;
    }
    std::shared_ptr<ClustParams> cp = read_clust_params(params, prior);

    ExRef *h = new ExRef();
    h->N = N; h->A = A; h->F = F;
    h->state.reset(new State(m->iteration,
        Entities(ent_ids, ent_attrs, attr_params, attr_indices),
        DistortProbs(theta, attr_params),
        DistortDistConcs(concs, attr_params),
        Links(links),
        Records(rec_attrs, fids, rec_dist, fsizes),
        Cache(attr_indices, attr_params),
        std::move(cp)));
    return h;
  } catch (std::exception &e) {
    std::cerr << "exref_create: " << e.what() << std::endl;
    return nullptr;
  }
}

void exref_destroy(void *p) { delete (ExRef *)p; }

// n sweeps of State::update(); returns 0, or -1 when the reference threw
int exref_update(void *p, int n) {
  ExRef *h = (ExRef *)p;
  try { for (int i = 0; i < n; i++) h->state->update(); }
  catch (std::exception &e) { std::cerr << "exref_update: " << e.what() << std::endl; return -1; }
  return 0;
}

// one sweep, phase by phase, with wall-clock seconds per phase:
// [0] clust params, [1] links, [2] entity attrs, [3] entity dists + cache + DP concs,
// [4] distortion indicators, [5] distortion probs   (order of src/state.h:60-75)
int exref_update_timed(void *p, double *secs) {
  ExRef *h = (ExRef *)p; State &s = *h->state;
  typedef std::chrono::steady_clock clk;
  auto t = [](clk::time_point a) { return std::chrono::duration<double>(clk::now() - a).count(); };
  try {
    auto t0 = clk::now(); s.clust_params_->update(s.links_); secs[0] += t(t0);
    t0 = clk::now(); s.links_.update(s.ents_, s.recs_, s.distort_dist_concs_, s.clust_params_, s.cache_.attr_indices_); secs[1] += t(t0);
    t0 = clk::now(); s.ents_.update_attributes(s.links_, s.recs_, s.distort_probs_, s.distort_dist_concs_, s.cache_); secs[2] += t(t0);
    t0 = clk::now(); s.ents_.update_distributions(); s.cache_.update(s.ents_);
    s.distort_dist_concs_.update(s.links_, s.recs_, s.ents_, s.cache_); secs[3] += t(t0);
    t0 = clk::now(); arma::imat corr = s.recs_.update_distortion(s.ents_, s.links_, s.distort_probs_, s.cache_); secs[4] += t(t0);
    t0 = clk::now(); s.distort_probs_.update(s.recs_, corr); secs[5] += t(t0);
    s.set_iteration(s.get_iteration() + 1);
  } catch (std::exception &e) { std::cerr << "exref_update_timed: " << e.what() << std::endl; return -1; }
  return 0;
}

int exref_n_entities(void *p) { return ((ExRef *)p)->state->links_.n_linked_ents(); }
int exref_iteration(void *p) { return ((ExRef *)p)->state->get_iteration(); }

// state read-back; ent_ids/ent_attrs need capacity N
void exref_get_state(void *p, int32_t *links, int32_t *n_ent, int32_t *ent_ids, int32_t *ent_attrs,
                     int32_t *rec_dist, double *theta, int32_t *dist_counts, double *clust2) {
  ExRef *h = (ExRef *)p; State &s = *h->state;
  for (int r = 0; r < h->N; r++) links[r] = s.links_.get_entity_id(r);
  std::vector<ent_id> ids = s.ents_.get_entity_ids();
  std::sort(ids.begin(), ids.end());
  *n_ent = (int)ids.size();
  for (size_t i = 0; i < ids.size(); i++) {
    ent_ids[i] = ids[i];
    for (int a = 0; a < h->A; a++) ent_attrs[a + (size_t)h->A * i] = s.ents_.get_attribute_value(ids[i], a);
  }
  for (int r = 0; r < h->N; r++)
    for (int a = 0; a < h->A; a++) rec_dist[a + (size_t)h->A * r] = s.recs_.get_attribute_distortion(r, a);
  for (int f = 0; f < h->F; f++)
    for (int a = 0; a < h->A; a++) {
      theta[f + h->F * a] = s.distort_probs_.get(f, a);
      dist_counts[f + h->F * a] = s.recs_.n_distorted(f, a);
    }
  NumericVector cv = s.clust_params_->to_R_vec(true);
  clust2[0] = cv.size() > 0 ? cv[0] : 0.0; // alpha | kappa
  clust2[1] = cv.size() > 1 ? cv[1] : 0.0; // d | m
}

// candidate entities of record rid in the current state (src/links.cpp:90-212), sorted.
// The record's own singleton entity is NOT removed first (update_link does that at :359).
int exref_possible_links(void *p, int rid, int32_t *out, int cap) {
  ExRef *h = (ExRef *)p; State &s = *h->state;
  std::vector<ent_id> c = getPossibleLinks(rid, s.recs_, s.ents_, s.cache_.attr_indices_);
  std::sort(c.begin(), c.end());
  int n = 0;
  for (ent_id e : c) if (n < cap) out[n++] = e;
  return (int)c.size();
}
double exref_loglik_existing(void *p, int rid, int eid) {
  ExRef *h = (ExRef *)p; State &s = *h->state;
  return logLikelihoodWeightExisting(rid, eid, s.links_.get_record_ids(eid), s.recs_, s.ents_,
                                     s.distort_dist_concs_, s.cache_.attr_indices_);
}
double exref_loglik_new(void *p, int rid) {
  ExRef *h = (ExRef *)p; State &s = *h->state;
  return logLikelihoodWeightNew(rid, s.recs_, s.cache_.attr_indices_, s.ents_);
}
double exref_prior_existing(void *p, int n) { return ((ExRef *)p)->state->clust_params_->prior_weight_existing(n); }
double exref_prior_new(void *p, int k) { return ((ExRef *)p)->state->clust_params_->prior_weight_new(k); }
int exref_cluster_size(void *p, int eid) { return ((ExRef *)p)->state->links_.n_records(eid); }
double exref_log_weight_entity_value(void *p, int eid, int aid, int vid) {
  ExRef *h = (ExRef *)p; State &s = *h->state;
  return logWeightEntityValue(vid, aid, s.recs_, s.distort_probs_, s.distort_dist_concs_,
                              s.links_.get_record_ids(eid), s.cache_.attr_indices_[aid].get(),
                              s.ents_.get_distribution(aid));
}
double exref_distortion_prob(void *p, int aid, int v, int w) {
  return ((ExRef *)p)->state->cache_.attr_indices_[aid]->get_distortion_prob(v, w);
}
double exref_exp_distortion_prob(void *p, int aid, int w) {
  ExRef *h = (ExRef *)p; State &s = *h->state;
  return s.cache_.attr_indices_[aid]->get_exp_distortion_prob(w, s.ents_.get_distribution(aid));
}
} // extern "C"
extern "C" void exref_get_concs(void *p, double *out) {
  ExRef *h = (ExRef *)p;
  for (int a = 0; a < h->A; a++) out[a] = h->state->distort_dist_concs_.get(a);
}
// ---- single phases of State::update, for the tests of the draws' laws
// Links::update only (src/links.cpp:477-489)
extern "C" int exref_update_links(void *p) {
  ExRef *h = (ExRef *)p; State &s = *h->state;
  try { s.links_.update(s.ents_, s.recs_, s.distort_dist_concs_, s.clust_params_, s.cache_.attr_indices_); }
  catch (std::exception &e) { std::cerr << "exref_update_links: " << e.what() << std::endl; return -1; }
  return 0;
}
// Entities::update_distributions only (src/entities.cpp:133-156)
extern "C" int exref_update_distributions(void *p) {
  ExRef *h = (ExRef *)p;
  try { h->state->ents_.update_distributions(); }
  catch (std::exception &e) { std::cerr << "exref_update_distributions: " << e.what() << std::endl; return -1; }
  return 0;
}
// entity-value pmf of attribute aid
extern "C" void exref_get_phi(void *p, int aid, int D, double *out) {
  ExRef *h = (ExRef *)p;
  IndexNonUniformDiscreteDist *d = h->state->ents_.get_distribution(aid);
  for (int v = 0; v < D; v++) out[v] = d->get_probability(v);
}

