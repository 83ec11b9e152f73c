/*
 * exb200.h - C ABI of the B200-native sampler for exchanger's partially-collapsed
 * Gibbs sweep.
 *
 * This header is the drop-in boundary.  It replaces the Rcpp entry
 *     Rcpp::S4 sample(Rcpp::S4 init_state, int n_samples, int thin_interval,
 *                     int burnin_interval)            (reference src/sample.cpp:118-219)
 * which R reaches through .Call('_exchanger_sample', ...) (R/RcppExports.R:12-14,
 * src/RcppExports.cpp:35-48).  The reference unmarshals an `ExchangERModel` S4
 * object into its C++ `State` (src/sample.cpp:71-105), runs `State::update()`
 * (src/state.h:60-75) in a loop and marshals the history and the final state
 * back (src/sample.cpp:14-63,171-218).  Here the glue flattens the same S4 slots
 * into the POD structs below; no R, Rcpp or torch type crosses this boundary.
 *
 * Conventions
 *   - every id is int32 and 0-based (the glue applies the same "-1" shift as
 *     src/sample.cpp:89-93); a missing record value is any negative number
 *     (R's NA_INTEGER = INT32_MIN is accepted as is);
 *   - matrices are column-major exactly as R stores them;
 *   - the caller owns every buffer; the library copies what it needs during the
 *     call and keeps no pointer afterwards;
 *   - every function returns EXB200_OK or a negative status; the message is
 *     available from exb200_last_error() (no exception crosses the ABI; the
 *     reference reports the same conditions through Rcpp::stop / std::runtime_error
 *     translated by BEGIN_RCPP/END_RCPP, src/RcppExports.cpp:38-47);
 *   - there is no CPU fallback: without a CUDA device exb200_create fails with
 *     EXB200_ERR_NO_DEVICE.
 */
#ifndef EXB200_H
#define EXB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 2: exb200_stats grew (n_probe, n_huge); slice calls (exb200_shard_*), device buffers, point
   estimates (exb200_pe_*) added
   3: exb200_pe_overflowed; exb200_stats grew (ms_cand_fast, n_slow)
   4: exb200_stats grew (n_listed, n_scanned, ms_tail, ms_entities_single); exb200_trim */
#define EXB200_ABI_VERSION 4

enum {
  EXB200_OK = 0,
  EXB200_ERR_INVALID = -1,     /* malformed model (sizes, ids out of range, bad pmf) */
  EXB200_ERR_UNSUPPORTED = -2, /* a model feature this build does not implement     */
  EXB200_ERR_NO_DEVICE = -3,   /* no CUDA device / CUDA runtime failure at start    */
  EXB200_ERR_CUDA = -4,        /* CUDA error during a call                           */
  EXB200_ERR_WEIGHTS = -5,     /* a conditional had no valid outcome (the reference's
                                  CHECK_WEIGHTS / AliasSampler errors,
                                  src/links.cpp:415-419, src/alias_sampler.h:76-97)  */
  EXB200_ERR_ABORTED = -6,     /* abort callback asked to stop (partial history kept) */
  EXB200_ERR_OOM = -7
};

/* Cluster prior families: reference src/clust_params.cpp:259-281 (read_clust_params). */
enum {
  EXB200_CLUST_PITMAN_YOR = 0, /* PitmanYorRP; EwensRP is d = 0 (R/PitmanYorRP.R:107-109) */
  EXB200_CLUST_GEN_COUPON = 1, /* GeneralizedCouponRP with finite kappa                   */
  EXB200_CLUST_COUPON = 2      /* GeneralizedCouponRP with kappa = Inf                    */
};

/* Entity-value distribution prior: reference src/entities.cpp:181-211. */
enum {
  EXB200_ENTDIST_FIXED = 0,    /* NULL (empirical probs) or a fixed numeric pmf */
  EXB200_ENTDIST_DIRICHLET = 1 /* DirichletRV with scalar alpha                 */
};

/*
 * One attribute: the `Attribute` spec (R/Attribute.R:112-119) together with its
 * `AttributeIndex` (R/AttributeIndex.R:68-88).  Replaces what
 * readAttributeIndex (src/attribute_index.cpp:118-149), Entities::Entities
 * (src/entities.cpp:163-212), DistortProbs::init_prior
 * (src/distortion_probs.cpp:4-21) and DistortDistConcs::init_prior
 * (src/distort_dist_concs.cpp:3-19) read from the S4 object.
 */
typedef struct exb200_attr {
  int32_t domain_size;          /* length(index@domain)                                 */
  int32_t is_constant;          /* 1: length(index@close_values) == 0 (categorical)     */
  int32_t exclude_entity_value; /* spec@exclude_entity_value                            */
  int32_t entity_dist_kind;     /* EXB200_ENTDIST_*                                     */
  const double* probs;          /* [D] index@probs (smoothed empirical pmf)             */
  const double* entity_dist;    /* [D] fixed pmf of the entity value, NULL = probs;
                                   for DIRICHLET: the concentration vector              */
  /* index@close_values in R's orientation: row v (an ENTITY value) lists the record
     values w with p(w | v) > 0; the library inverts it like
     src/attribute_index.cpp:125-140.  All NULL / 0 for a constant attribute.           */
  const int64_t* close_row_ptr; /* [D + 1]                                              */
  const int32_t* close_val_ids; /* [nnz] 0-based record values w                        */
  const double* close_probs;    /* [nnz] p(w | v)                                       */
  const double* max_exp_factor; /* [D] index@max_exp_factor (NULL if constant)          */
  double distort_prob_shape1;   /* BetaRV prior on theta[.,a]; <= 0: theta stays fixed  */
  double distort_prob_shape2;
  double distort_dist_conc;     /* distort_dist_concs[a]; +Inf for DirichletProcess(Inf) */
  double conc_prior_shape;      /* GammaRV prior on the DP concentration; <= 0: none    */
  double conc_prior_rate;
} exb200_attr;

/* Cluster prior and current parameter values (clust_prior / clust_params slots). */
typedef struct exb200_clust {
  int32_t kind; /* EXB200_CLUST_* */
  double alpha, d; /* Pitman-Yor values                                 */
  double m, kappa; /* coupon values (m is integer valued)               */
  double alpha_prior_shape, alpha_prior_rate; /* GammaRV, <= 0: fixed  */
  double d_prior_shape1, d_prior_shape2;      /* BetaRV,  <= 0: fixed  */
  double kappa_prior_shape, kappa_prior_rate; /* GammaRV, <= 0: fixed  */
  double m_prior_size, m_prior_prob;          /* ShiftedNegBinomRV, size <= 0: fixed */
} exb200_clust;

/* The `ExchangERModel` slots (R/ExchangERModel.R:166-182) as flat arrays. */
typedef struct exb200_model {
  int32_t n_records;              /* N */
  int32_t n_attrs;                /* A */
  int32_t n_files;                /* F */
  int32_t n_entities;             /* K = length(ent_ids) */
  int32_t iteration;              /* @iteration */
  const int32_t* rec_attrs;       /* [A x N] @rec_attrs - 1, NA kept negative           */
  const int32_t* rec_distortions; /* [A x N] @rec_distortions                           */
  const int32_t* file_ids;        /* [N] @file_ids - 1                                  */
  const int32_t* links;           /* [N] @links (labels taken from ent_ids)             */
  const int32_t* ent_ids;         /* [K] @ent_ids                                       */
  const int32_t* ent_attrs;       /* [A x K] @ent_attrs - 1                             */
  const double* distort_probs;    /* [F x A] @distort_probs                             */
  const exb200_attr* attrs;       /* [A] */
  exb200_clust clust;
  uint64_t seed;                  /* Philox key; the glue derives it from R's RNG so
                                     set.seed() still fixes the chain                   */
} exb200_model;

/*
 * Caller-allocated history, filled with the save schedule of
 * src/sample.cpp:171-197.  Row s of every array is sample s (sample-major; the R
 * glue transposes into R's column-major n_samples x p matrices).  Any pointer may
 * be NULL to skip that series.
 */
typedef struct exb200_history {
  int32_t* links;             /* [n_samples][N]  hist_links                           */
  double* distort_probs;      /* [n_samples][F*A] column f + a*F (sample.cpp:144-149) */
  int32_t* distort_counts;    /* [n_samples][F*A] same column order                   */
  double* distort_dist_concs; /* [n_samples][A]                                       */
  int32_t* n_linked_ents;     /* [n_samples]                                          */
  double* clust_params;       /* [n_samples][2]: (alpha, d) or (kappa, m); the glue
                                 keeps the random ones (clust_params.cpp:161-216)      */
  int32_t n_saved;            /* out: rows actually written                           */
} exb200_history;

/* Caller-allocated final state (what buildFinalS4State writes, sample.cpp:14-63). */
typedef struct exb200_state {
  int32_t iteration;
  int32_t n_entities;          /* out: K */
  int32_t* links;              /* [N]                                 */
  int32_t* ent_ids;            /* [N] capacity, first K filled        */
  int32_t* ent_attrs;          /* [A x N] capacity, first K columns   */
  int32_t* rec_distortions;    /* [A x N]                             */
  double* distort_probs;       /* [F x A]                             */
  int32_t* distort_counts;     /* [F x A] Records::distorted_counts_  */
  double* distort_dist_concs;  /* [A]                                 */
  exb200_clust clust;          /* current values, priors echoed       */
} exb200_state;

/* Per-sweep counters of the last completed sweep (for measurement and tests). */
typedef struct exb200_stats {
  int64_t n_candidates;   /* sum over records of compatible candidate items        */
  int64_t n_bucket_items; /* items visited while generating candidates             */
  int32_t n_rounds;       /* dependency rounds of the links update                 */
  int32_t n_wide;         /* records that took the serial wide path                */
  int32_t n_entities;     /* K after the sweep                                     */
  int32_t n_links_changed;/* records whose cluster changed                         */
  int32_t n_tables;       /* blocking tables built                                 */
  int32_t n_long;         /* records whose turn is taken by a block (lists longer than long_min) */
  float ms_clust, ms_prep, ms_index, ms_cand, ms_rounds, ms_canon, ms_entities,
      ms_distort, ms_total, ms_cand_long, ms_wide, ms_index_count, ms_index_fill; /* CUDA-event times */
  int32_t n_probe;        /* records taken by a whole warp in the second candidate launch (several
                             buckets to probe, or one bucket of 513..8192 items)                  */
  int32_t n_huge;         /* records whose bucket was cut into tiles (longer than huge_min)       */
  float ms_cand_fast;     /* the thread-per-record candidate kernel alone (part of ms_cand)       */
  int32_t n_slow;         /* records that kernel left to the bucket-scanning kernels              */
  int32_t n_listed;       /* records no table served directly whose possible links were listed through
                             the cheapest table built this sweep (k_wide_emit)                    */
  int32_t n_scanned;      /* records whose possible links were listed by the scan over all items  */
  float ms_tail;          /* the persistent kernel of the dependency rounds alone (part of ms_rounds; 0: not run) */
  float ms_entities_single; /* the first entity kernel alone: singletons (part of ms_entities)     */
} exb200_stats;

typedef struct exb200_handle exb200_handle;

/* returns EXB200_ABI_VERSION */
int exb200_abi_version(void);

/* Device slabs and pinned staging rows of destroyed handles are kept for the next handle of the process
   (cudaMalloc / cudaFree of a chain's memory cost 0.05-0.5 s per handle); this gives them back to the
   driver.  Safe to call at any time; handles that exist are not affected. */
void exb200_trim(void);

/* number of visible CUDA devices (0 without a driver/GPU) */
int exb200_device_count(void);

/* Build the device state from a host model on CUDA device `device`.
   Replaces readS4State + the State constructor (sample.cpp:71-105, state.h:42-54). */
int exb200_create(const exb200_model* model, int device, exb200_handle** out);

/* Run the sampling loop of sample() (sample.cpp:171-202): sweeps until n_samples
   are saved.  abort_cb (may be NULL) replaces Progress::check_abort (sample.cpp:200);
   progress_cb (may be NULL) replaces p.increment() (sample.cpp:201).
   A saved links row leaves the device behind the next sweep (snapshot -> pinned buffer on a copy
   stream -> the caller's array, written by a helper thread of the library); every row has landed
   when the call returns, and out->n_saved tells how many did if it returns an error.  The
   callbacks are only ever called on the calling thread. */
int exb200_run(exb200_handle* h, int32_t n_samples, int32_t thin_interval,
               int32_t burnin_interval, exb200_history* out,
               int (*abort_cb)(void*), void (*progress_cb)(void*, int32_t), void* ctx);

/* n_sweeps applications of State::update() with nothing copied back
   (device-resident timing; state.h:60-75). */
int exb200_sweeps(exb200_handle* h, int32_t n_sweeps);

/* exb200_sweeps bracketed by CUDA events on the library's own stream (torch.cuda.Event only sees
   torch's streams): elapsed device milliseconds and the number of kernels launched. */
int exb200_sweeps_timed(exb200_handle* h, int32_t n_sweeps, float* device_ms, int64_t* launches);

/* Copy the current state to the caller (buildFinalS4State, sample.cpp:14-63). */
int exb200_get_state(exb200_handle* h, exb200_state* out);

int exb200_get_stats(exb200_handle* h, exb200_stats* out);

/*
 * Within-chain sharding over several GPUs of one node.  Every GPU holds the
 * whole state and runs the links update (steps 1-2 of State::update, state.h:60-67) identically -
 * the randomness is keyed, so the replicas agree bit for bit.  The entity attributes
 * (Entities::update_attributes, state.h:68) and the distortion indicators
 * (Records::update_distortion, state.h:72) are computed on a slice of the entity slots / records
 * per GPU; the host exchanges the slices between the calls (all-gather of the device buffers
 * "ent_y" rows [lo, hi) and "rec_z" entries [lo, hi), sum of the two count matrices):
 *
 *   exb200_shard_links(h);
 *   exb200_shard_entities(h, lo, hi);   exb200_stream_sync(h);   all-gather "ent_y"
 *   exb200_shard_distort(h, lo, hi, ndist, corr);                all-gather "rec_z", sum counts
 *   exb200_shard_finish(h, ndist_sum, corr_sum);
 *
 * Models with a DirichletRV entity prior or a GammaRV concentration prior (steps 4-6) are not
 * supported in this mode (EXB200_ERR_UNSUPPORTED).  Count matrices are [F x A], file-major
 * (f * A + a).  One call sequence with lo = 0, hi = n_records equals exb200_sweeps(h, 1).
 */
int exb200_shard_links(exb200_handle* h);
int exb200_shard_entities(exb200_handle* h, int32_t lo, int32_t hi);
int exb200_shard_distort(exb200_handle* h, int32_t lo, int32_t hi, int32_t* ndist, int32_t* corr);
int exb200_shard_finish(exb200_handle* h, const int32_t* ndist, const int32_t* corr);
/* device address and size of a state buffer: "ent_y" ([2N][row] int32, row = *row_bytes / 4
   values per entity slot), "rec_z" ([N] uint32, one distortion bit per attribute) */
int exb200_device_buffer(exb200_handle* h, const char* name, void** ptr, int64_t* bytes, int32_t* row_bytes);
/* wait until the library's stream has finished everything queued so far */
int exb200_stream_sync(exb200_handle* h);

/*
 * Test hook: the conditional of one record's link in the CURRENT state, without
 * changing it: what Links::update_link computes at src/links.cpp:361-399.
 * cand_ids/log_weights receive up to `cap` existing candidates in ascending entity
 * id (log prior_weight_existing + logLikelihoodWeightExisting); *log_weight_new is
 * log prior_weight_new + logLikelihoodWeightNew.  Returns the candidate count.
 */
int exb200_link_conditional(exb200_handle* h, int32_t rid, int32_t cap, int32_t* cand_ids,
                            double* log_weights, double* log_weight_new);

/* Options (name, value); unknown names return EXB200_ERR_INVALID.
   One option is part of the sampling specification (it fixes the systematic-scan order of the
   links update) and therefore changes the chain:
     "wide_threshold" (default 4096, <= 16384): records with more possible links than this
                 take their turn first, in ascending id, before all other records
   The others change no result, only how the work is scheduled:
     "cand_cap"  candidates a group of lanes stages per record (<= 64); more go to a block-wide kernel
     "long_min"  list length above which a record's turn is taken by a block instead of a thread
     "bucket_cap" bucket length above which a whole warp (up to 8192 items) or a block scans the
                 bucket; 0 sends every record to the block-wide kernel
     "huge_min"  bucket length above which the bucket is cut into tiles scanned by many blocks
     "est_threshold" expected bucket size above which the blocking key uses every usable attribute
     "table_cost_permille" assumed cost of building one blocking table, in bucket visits per item
                 x 1000 (default 1000); 0 builds every table that some record wants
     "table_min_usage" (default 16) a key mask wanted by at most this many records gets no table of
                 its own: another table serves them, or the batched scan over all items
     "index_partitioned" 0 builds every table with global atomics (debugging)
     "fast_path" (default 1) candidate lists by one thread per record over tables in chain form
                 (needs attribute domains that pack into 64 bits); 0: bucket-scanning kernels only
     "hop_cap"   (default 1024) chain items a thread walks for one record before the record is left
                 to the bucket-scanning kernels (whose CSR form of that table is then built too)
     "wide_buckets" (default 1) the possible links of a wide record are listed from the bucket its
                 candidate scan found too long, and sorted; 0: by one more scan over all items
     "tiny_tables_always" (default 0) build a wanted table with at most 8192 buckets however few
                 records want it (otherwise "table_min_usage" decides)
     "rounds_tail" (default 1) the dependency rounds after the first run in one persistent
                 cooperative kernel; 0 launches every round from the host
     "check"     recompute sizes / member lists / K from the links after every sweep and compare
     "trace"     print the phases of every sweep to stderr (also: environment EXB200_TRACE) */
int exb200_set_option(exb200_handle* h, const char* name, int64_t value);

void exb200_destroy(exb200_handle* h);

/*
 * Point estimates on the device from the stream of saved link vectors, without the n_samples x N
 * history: what mp_clusters / smp_clusters compute from it (src/point_estimate.cpp:58-136,
 * R/analysis.R:53-186).  Samples are added one at a time, from a host row of the history or
 * straight from a chain's current links.  Labels must lie in [0, 2 * n_records).
 *   smp_membership[r] : smallest record of r's shared most probable cluster (records with equal
 *                       values form one cluster of smp_clusters)
 *   mp_prob[r]        : relative frequency of r's most probable cluster (mp_clusters()$probabilities)
 *   mp_sample[r]      : index of the first added sample in which that cluster appeared (its members
 *                       are the records linked like r in that sample)
 *   tied[r]           : 1 if another cluster of r is as frequent (the reference then takes the
 *                       lexicographically smallest member set; here the smaller fingerprint)
 *   *n_overflow       : occurrences not counted because a record was seen in more than `slots`
 *                       distinct clusters
 * Any output pointer may be NULL.
 */
typedef struct exb200_pe exb200_pe;
int exb200_pe_create(int32_t n_records, int32_t slots, int device, exb200_pe** out);
int exb200_pe_add(exb200_pe* pe, const int32_t* links);
int exb200_pe_add_chain(exb200_pe* pe, exb200_handle* h);
int exb200_pe_finish(exb200_pe* pe, int32_t* smp_membership, double* mp_prob, int32_t* mp_sample,
                     int32_t* tied, int64_t* n_overflow);
/* flags[r] = 1 for every record that was seen in more than `slots` distinct clusters: some of its
   occurrences were not counted, so its estimate is unreliable and has to be redone from the link
   history (analysis.DevicePointEstimate.finish does) or with more slots.  `slots` >= the number of
   samples added can never overflow. */
int exb200_pe_overflowed(exb200_pe* pe, uint8_t* flags);
/* Record pairs ([n_pairs][2], 0-based) whose coreference is counted in every sample added from
   now on: counts[p] = samples in which both records had the same entity, the numerator of
   coreference_probs (R/coreference_probs.R:54-82); *n_samples = samples added since the pairs
   were set. */
int exb200_pe_set_pairs(exb200_pe* pe, const int32_t* pairs, int64_t n_pairs);
int exb200_pe_pair_counts(exb200_pe* pe, int32_t* counts, int32_t* n_samples);
void exb200_pe_destroy(exb200_pe* pe);
/* Every sample exb200_run saves from now on is also added to `pe` (NULL: detach), on the device;
   with exb200_history.links = NULL the link vectors then never leave the GPU. */
int exb200_attach_point_estimate(exb200_handle* h, exb200_pe* pe);

/*
 * One-off model preparation on the GPU: the value neighbourhoods of a string attribute under
 * transform_dist_fn(Levenshtein(), threshold) - what compute_close_values builds in R
 * (R/AttributeIndex.R:117-137; the result fills exb200_attr.close_* and max_exp_factor).
 * `chars` holds n strings of at most 31 bytes each, packed with a stride of 32 bytes;
 * distances are unit-cost edit distances over bytes; pairs farther apart than `threshold`
 * (an integer, <= 8) are not neighbours; a value is its own neighbour only if include_self
 * (i.e. exclude_entity_value = FALSE).  Not part of the sweep.
 */
typedef struct exb200_neighbours exb200_neighbours;
int exb200_neighbours_create(const uint8_t* chars, const int32_t* lens, int32_t n, int32_t threshold,
                             int32_t include_self, int device, exb200_neighbours** out);
int64_t exb200_neighbours_nnz(const exb200_neighbours* nb);
/* row_ptr [n+1], val_ids [nnz], probs [nnz] (exp(-d) normalised per row), max_exp_factor [n] */
int exb200_neighbours_get(const exb200_neighbours* nb, int64_t* row_ptr, int32_t* val_ids, double* probs,
                          double* max_exp_factor);
void exb200_neighbours_destroy(exb200_neighbours* nb);

/* Message of the last failing call on this handle (or of exb200_create when h is NULL). */
const char* exb200_last_error(const exb200_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* EXB200_H */
