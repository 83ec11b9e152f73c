// Device-resident layout of the sampler state and of the per-sweep scratch.
//
// Records (static) ..... rec_x   int32 [N][RS]   x[0..A-1], file id at [RS-1]; RS = 8 for A <= 7 so
//                                                that one record is one 32-byte sector
// Indicators ........... rec_z   uint32 [N]      bit a = z[a, r]            (Records::attr_distortions_)
// Links ................ lambda  int32 [N]       entity slot of record r     (Links::forward_index_)
// Entities ............. ent_y   int32 [2N][ES]  attribute values; slots [0,N) hold the entities that
//                                                exist at the start of a sweep (slot = smallest member
//                                                record), slot N + r is the entity record r would found
//                                                during the sweep (Entities::forward_index_)
//                        ent_size int32 [2N]     cluster sizes              (Links::inverse_index_ sizes)
//                        ent_head int32 [2N], rec_next int32 [N]  member lists in ascending record id
// The reference keeps the same information in unordered_map / unordered_set containers
// (src/links.h:23-26, src/entities.h:50-51, src/records.h:20-24).
#pragma once
#include <stdint.h>

#include "../../include/exb200.h"

namespace exb {

constexpr int NMAX_FAST = 8;        // categorical fast scheme: up to this many observed members
constexpr int REJECT_CAP = 100000;  // safety bound of the rejection loops
constexpr int CAND_SMEM = 64;       // compatible candidates an 8-lane group can stage for one record
constexpr int LPE_MAX = 32;         // host-computed log prior_weight_existing(n), n < LPE_MAX
constexpr int MAX_ATTRS = 32;
constexpr int KEY_ATTRS = 12;       // attributes (largest domains first) that may form blocking keys
constexpr int N_TABLES = 1 << KEY_ATTRS;
constexpr int WIDE_BLOCKS = 1024;   // partition of the item range in the wide path
constexpr int CAND_INLINE = 4;      // candidate lists up to this length are stored at pool[CAND_INLINE * record]
constexpr int FAST_CAND = 16;       // candidates the thread-per-record kernel keeps for one record
constexpr uint32_t CAND_OWN_ONLY = 0xFFFFFFFFu; // cand_off value: the list is {the record's own entity}, log-likelihood 0
constexpr int WIDE_NB = 128;        // records whose possible links are listed in one pass over the items
constexpr int ACT_TABS_MAX = 48;    // built tables a record without a serving table chooses from
constexpr int HUGE_BATCH = 512;     // records with a huge bucket scanned per launch

// Per-attribute tables: the device form of ConstantAttributeIndex / NonConstantAttributeIndex
// (src/attribute_index.h:12-146) plus the alias tables that replace IndexNonUniformDiscreteDist
// (src/discrete_dist.h:197-294).  All logarithms of static quantities are computed on the host.
struct DevAttr {
  int D, is_const, exclude, pad;
  const double *psi, *phi, *om;            // empirical pmf, entity-value pmf, 1 - psi
  const double *logpsi, *log1mpsi, *logphi;
  const double *logE;                      // log E_{v~phi} p(x | v)  (attribute_index.cpp:55-62,94-101)
  const double *mef;                       // max_exp_factor (all 1 for a constant attribute)
  const int *row, *col;                    // inverted neighbourhood CSR: row w -> entity values v, ascending
  const double *p, *logp;                  // p(w | v) and its log
  const double *cs;                        // per row: inclusive prefix sums of phi(v) mef(v) p(w|v)
  const double *cs2;                       // ... and of phi(v) (mef(v) p(w|v))^2 (two members with the same value)
  const double *csn;                       // per row: inclusive prefix sums of phi(v) p(w|v) (new-entity draw)
  const double *g;                         // [(NMAX_FAST+1)][D]  phi(v) / (1-psi(v))^n
  const double *aprob; const int *aalias;  // alias tables of g_n, same shape
  const double *pprob; const int *palias;  // alias table of psi
  double G[NMAX_FAST + 1];                 // sum_v g_n(v)
  double *lt_diff, *lt_same;               // per sweep, [F][D]: log(theta mef(v)); log(1 - theta mef(v))
                                           // or log(1 - eff + eff p(v|v)) when the value is not excluded
  double *sg;                              // per sweep, [F][D][2], exclude_entity_value only: for an entity whose ONE observed
                                           // member (file f) has value x, the weight of keeping x and of all other values
};

// One blocking table: items (entity slots) bucketed by their values on a set of key attributes
// (bit k of `mask` = key attribute k).  CSR over buckets built per sweep by count / scan / fill.
// A table comes in two forms.  CSR (count / scan / fill): contiguous buckets that many lanes can scan -
// needed for the few key masks with long buckets.  CHAIN (key spaces about as large as the item
// set): head[cell] -> item, node[item] = (packed attribute tuple, next item, flags) - one atomicExch
// per item to build, no bucket bounds to fetch, and the tuple a probe has to compare sits in the
// node it reads anyway (no gather of the entity table).
struct ChainNode {
  unsigned long long P; // packed attribute tuple of the item (pack_tuple)
  int next;             // next item of the cell's chain (-1: end)
  uint32_t flags;       // bit 0: the item is an entity slot whose potential entity is its twin
};
struct DevTable {
  uint32_t mask;
  int exact;         // 1: bucket = mixed-radix key; 0: bucket = mix(key) & (B - 1)
  uint32_t B;        // number of buckets
  int chain;         // 1: the chain form is built this sweep (probes walk head / node)
  int *ptr;          // [B]  after fill: END offset of bucket (start = ptr[b-1], 0 for b = 0)
  int *ids;          // [n_items]
  unsigned long long kmask; // fields of the packed tuple that form the key
  uint32_t cells;    // chain form: number of cells (power of two)
  int csr;           // 1: the CSR form is built this sweep
  int *head;         // [cells]
  ChainNode *node;   // [2N]
};

struct ClustDev {
  int kind;
  double alpha, d, m, kappa;
  double logpe[LPE_MAX];
};

struct DevState {
  int N, A, F, RS, ES;
  uint32_t iter;
  uint64_t seed;
  // model state
  const int *rec_x;
  uint32_t *rec_z;
  int *lambda;
  int *ent_y;
  int *ent_size, *ent_head;     // current arrays [2N]
  int *ent_size_nx, *ent_head_nx; // next arrays (written by the rename kernel)
  int *rec_next;
  const double *theta;          // [F*A] index f*A + a
  const DevAttr *attrs;
  // links-update scratch
  double *Lnew;                 // [N] logLikelihoodWeightNew
  uint16_t *rec_tab;            // [N] key mask the record would like: every usable key attribute when its best
                                //     pair is expected to match many items, else that pair (0: none)
  uint16_t *rec_alt;            // [N] key mask of its best pair
  uint8_t *twin;                // [N] 1: potential entity N+r has the same attributes as entity slot r (a
                                //     singleton of r); it is then not indexed but implied by slot r
  uint8_t *live0;               // [N] 1: entity slot r held an entity when the sweep started
  uint32_t *rec_bkt;            // [N] bucket inside that table
  const DevTable *tables;       // [2^n_key] indexed by key mask
  int *tab_usage;               // [2][2^n_key] records wanting each full mask / each pair mask
  float *tab_est;               // [2^n_key] sum over the records wanting a full mask of the expected size
                                //           of their pair bucket (what falling back to the pair would cost)
  const uint16_t *tab_remap;    // [2][2^n_key] wanted full mask / pair mask -> mask of a built table (0: none)
  int pk_ok;                    // the attribute domains fit a 64-bit packed tuple (fast candidate path)
  const unsigned long long *pk; // [2N] packed tuples of the indexed items of this sweep (k_chain_pack; null: not made)
  const uint8_t *pkfl;          // [2N] bit 0: indexed, bit 1: twin
  uint8_t pk_shift[MAX_ATTRS], pk_bits[MAX_ATTRS]; // field of attribute a in the packed tuple
  int hop_cap;                  // chain items a thread walks for one record before the record is left to the CSR path
  int *slow_list;               // [N] records the fast candidate kernel leaves to the bucket-scanning kernels
  int *tab_over;                // [2^n_key] chain tables in which some record exceeded hop_cap / FAST_CAND
  int n_key;                    // number of key attributes
  int n_act_tabs;               // blocking tables built this sweep ...
  uint16_t act_tabs[ACT_TABS_MAX]; // ... and their key masks (a record no table serves picks the cheapest of them, k_wide_emit)
  double emit_cost_cap;         // ... unless even that one costs more cell / item reads than this (then: scan over all items)
  int key_attr[KEY_ATTRS];      // key position -> attribute
  int8_t key_pos[MAX_ATTRS];    // attribute -> key position (-1: not a key attribute)
  uint32_t *cand_off;           // [N]
  int *cand_cnt;                // [N] >= 0: stored list; -1: wide (turn taken first); -2: list did not fit the pool
                                //     (serial turn inside the rounds); -3: no usable bucket yet (batched scan)
  int *cand_id; double *cand_ll;// candidate pool
  unsigned long long *pool_cursor; unsigned long long pool_cap;
  unsigned long long *owner;    // [2N] reservation words
  unsigned long long *fence;    // reservation word of the wide records
  uint8_t *fin;                 // [N] 1 once the record's link is final; 2: ready (owns its items, weights cached)
  double *rdy_m1, *rdy_srel;    // [N] cached max log-weight and relative sum of a ready record
  uint8_t *referenced;          // [2N] 1: the item is on some record's stored candidate list
  int8_t *dlo, *dhi;            // [N] bounds of the record's change of K
  int *plo, *phi_;              // [N] exclusive prefix sums of dlo / dhi
  int *act_in, *act_out;        // active records with short candidate lists (thread per record)
  int *actl_in, *actl_out;      // active records with long candidate lists (warp per record)
  int *long_todo;               // records whose candidates did not fit a warp's staging area
  int *probe_list;              // records that probe several buckets of their table (warp per record)
  int *wide_list;               // records with more than wide_threshold possible links (taken first)
  int *unk_list;                // records without a usable blocking bucket: classified one by one
  int *huge_list;               // records whose bucket is scanned by the whole grid
  int *huge_id; double *huge_ll;// [HUGE_BATCH][wide_threshold + 2] scratch lists of that scan
  int *huge_tiles, *huge_cnt;   // [HUGE_BATCH + 1] first tile of each record of the batch, [HUGE_BATCH] list lengths
  int *batch_list;              // [WIDE_NB] records of the current k_wide_scan
  int *wl_sel;                  // [WIDE_NB] batch positions handed to k_wide_store
  int *wl_cnt;                  // [WIDE_NB][WIDE_PIECES] possible links per record and piece, then their scan
  int *wl_tot;                  // [WIDE_NB] possible links per record
  unsigned long long *wl_base;  // [WIDE_NB] start of the record's list in wl_id / wl_ll (~0: not listed)
  int *wl_id; double *wl_ll;    // [wl_cap] lists of the current batch (carved out of lwbuf)
  int *wlb_id; double *wlb_ll;  // lists of the wide records that have a bucket, sorted by (record, item) (k_wide_emit / sort / k_wide_split)
  uint8_t *wl_log_k; int *wl_log_item; double *wl_log_ll; // [WIDE_PIECES][wl_logcap] hits of the counting pass, per piece
  int *wl_logn;                 // [WIDE_PIECES] entries logged (-1: the piece's log overflowed)
  int wl_logcap;
  unsigned long long wl_cap;
  int huge_min;                 // bucket length above which the bucket is cut into tiles (k_huge_*)
  double *wide_part;            // [WIDE_BLOCKS * 257 + WIDE_BLOCKS] partial maxima / sums of the wide path
  int *counters;                // see CounterSlot
  long long *counters64;        // [0] candidates, [1] bucket items visited
  int *dev_err;                 // [0] status, [1] id
  double *lwbuf;                // [16N] scratch: lists of the serial path / test hook (2N values)
  int *ndist, *corr;            // [F*A] distorted counts, corrected non-distorted counts
  int bucket_cap, cand_cap, long_cap, long_min; // thresholds: see exb200_set_option
  double est_threshold;         // expected bucket size above which the full key is used
  int K0;                       // number of entities at the start of the links update
  int use_kbounds;              // 0 when prior_weight_new does not depend on K
  ClustDev cl;
  double conc[MAX_ATTRS];       // DP concentration per attribute (infinite: none)
  int has_dp;                   // some concentration is finite
  uint32_t dp_mask;             // bit a: attribute a has a finite concentration
};

enum CounterSlot {
  C_NACT_OUT = 0,   // size of act_out
  C_MIN_ACTIVE = 1, // smallest unfinalised record of the round
  C_WIDE_READY = 2, // wide record whose turn has come (-1: none)
  C_CHANGED = 3,    // records whose cluster changed
  C_DK_TOTAL = 4,   // sum of dK over finalised records
  C_NWIDE = 5,      // wide records of the sweep
  C_K = 6,          // entities counted by the rename kernel
  C_NACTL_OUT = 7,  // size of actl_out
  C_NLONG_TODO = 8, // size of long_todo
  C_NACTL_INIT = 9, // long records found by candidate generation
  C_NUNK = 10,      // size of unk_list
  C_NFENCE = 11,    // records that take their turn serially inside the rounds
  C_LONG_NEXT = 12, // next entry of long_todo to hand out (k_cand_long)
  C_NHUGE = 13,     // size of huge_list
  C_NOLIST = 14,    // records without a stored list after candidate generation (diagnostic)
  C_NPROBE = 15,    // size of probe_list
  C_NSLOW = 16,     // size of slow_list
  C_NMULTI = 17,    // records the first pass of k_cand_fast left to its second pass
  C_NSLOTS = 18
};

} // namespace exb
