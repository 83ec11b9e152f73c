// Host side of the C ABI (include/exb200.h): device state construction, the sweep driver that
// fixes the partially-collapsed order of State::update (src/state.h:60-75), the scalar
// hyper-parameter updates, history streaming and state read-back.
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <type_traits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "hostmath.hpp"
#include "kernels.cuh"
#include "neighbours.cuh"
#include "setup.cuh"
#include "point_estimate.cuh"

using namespace exb;

namespace {

thread_local std::string g_create_error;

// array of n values left uninitialised (the neighbourhood-sized host arrays are written in full by
// several threads: no single-threaded zero fill, pages are first touched in parallel)
template <class T> struct RawVec {
  std::unique_ptr<T[]> p;
  size_t n = 0;
  void resize(size_t m) { p.reset(m ? new T[m] : nullptr); n = m; }
  size_t size() const { return n; }
  bool empty() const { return n == 0; }
  T* data() { return p.get(); }
  const T* data() const { return p.get(); }
  T& operator[](size_t i) { return p[i]; }
  const T& operator[](size_t i) const { return p[i]; }
};

struct HostAttr {
  int D = 0, is_const = 1, exclude = 1;
  double s1 = 0, s2 = 0; // Beta prior on theta
  double conc_shape = 0, conc_rate = 0; // Gamma prior on the DP concentration (<= 0: fixed)
  bool dirichlet = false;               // Dirichlet prior on the entity-value distribution
  std::vector<double> psi, phi, mef, om, dir_alpha;
  std::vector<int> row;                 // inverted neighbourhood CSR (host copy)
  RawVec<int> col;
  RawVec<double> p;
};

struct TableHost {
  DevTable d{};
  bool allocated = false;
  unsigned long long keys = 1; // size of the key space (product of the domain sizes)
};

} // namespace

struct exb200_handle {
  int device = 0;
  cudaStream_t stream = nullptr;
  int N = 0, A = 0, F = 0, RS = 8, ES = 8;
  int iteration = 0;
  int K = 0;
  uint64_t seed = 0;
  exb200_clust cl{};
  HostStream hs;
  std::vector<HostAttr> attrs;
  std::vector<double> theta;   // [F*A] f*A + a
  double conc[MAX_ATTRS];      // DP concentrations (infinite: none)
  int* d_hist = nullptr;       // [max D] entity-value histogram of one attribute
  std::vector<int> ndist, corr;
  DevState S{};                // device pointers (by-value kernel argument)
  // device memory comes from a few large slabs (bump allocation, freed together: a handle makes
  // about a hundred allocations and cudaFree is slow per call)
  struct Slab { char* base; size_t size, used; };
  std::vector<Slab> slabs;
  std::mutex alloc_mu;          // attribute tables are built by concurrent host threads at create time
  DevAttr* d_attrs = nullptr;
  std::vector<DevAttr> h_attrs;
  std::vector<TableHost> tables; // [2^n_key] indexed by key mask
  DevTable* d_tables = nullptr;
  int* d_active = nullptr;
  uint16_t* d_remap = nullptr;
  PartBuild pb{};              // partitioned build of the large tables
  int* d_part_tabs = nullptr;
  int* d_chain_tabs = nullptr;
  // scratch of wide_bucket_lists
  unsigned long long* we_keys[2] = {nullptr, nullptr}; double* we_vals[2] = {nullptr, nullptr};
  unsigned long long* we_cursor = nullptr; unsigned long long we_cap = 0;
  int* we_recs = nullptr; int* we_seg = nullptr; void* we_tmp = nullptr; size_t we_tmp_bytes = 0;
  int opt_wide_buckets = 1;
  long long emit_scratch = 0; // entries of the listing scratch (0: sized by the record count); set before the first sweep
  int emit_chunk = 65536; // records listed by one k_wide_emit pass at most (halved when their lists do not fit the scratch)
  int opt_round0 = 1; // records that found an entity whatever happens are final before round 1 (k_bounds_init)
  double emit_cost_cap = 262144.0; // cell / item reads k_wide_emit may spend on a record no table serves (else: scan over all items)
  int opt_tiny_always = 0;     // build every wanted table with few buckets, however few records want it (measured: slower)
  WideTurn* d_turns = nullptr; size_t turns_cap = 0; int turns_grid = 0;
  unsigned long long* d_turn_dbg = nullptr; size_t turn_dbg_cap = 0; // (trace)
  unsigned* d_turn_sync = nullptr; // [2 * turns_cap] arrivals / commit flags of k_wide_turns
  unsigned long long* d_pk = nullptr; uint8_t* d_pkfl = nullptr; // packed tuples / flags of the items (chain build)
  int opt_fast = 1;            // thread-per-record candidate kernel over chain tables (needs packed tuples)
  int hop_cap = 1024;
  int pb_tables_cap = 0;
  int huge_cap_alloc = -1;     // wide_threshold the huge-bucket scratch lists were sized for
  int* d_bsum = nullptr; size_t bsum_cap = 0;
  unsigned long long* d_cl_counts = nullptr;
  double* d_theta = nullptr;
  unsigned long long* d_gen_list = nullptr; // (entity << 8 | attribute) pairs that need the general scheme
  int* d_gen_count = nullptr;
  uint32_t round_ctr = 1;
  int opt_tail = 1;            // rounds after the first in one persistent kernel (k_rounds_tail)
  TailState tail{};            // its lists (allocated when first used)
  void* d_select_tmp = nullptr; size_t select_tmp_bytes = 0;
  int tail_grid = 0;           // co-resident blocks of k_rounds_tail for the current wide_threshold
  int tail_grid_cap = -1;      // ... and the wide_threshold it was computed for
  int opt_check = 0;
  int trace = std::getenv("EXB200_TRACE") ? 1 : 0;
  int bucket_cap = 512, cand_cap = CAND_SMEM, long_cap = 4096, long_min = 32;
  double est_threshold = 16.0;
  int huge_min = 65536;
  int index_partitioned = 1;
  double table_full_cost = 1.0, table_pair_cost = 0.25; // cost of building a table, in bucket visits per item
  int table_min_usage = 16;     // a key mask wanted by at most this many records never gets a table of its own
  int n_tab = 0; // 2^n_key
  exb200_stats stats{};
  cudaEvent_t ev[20]{};
  bool tail_ran = false, ent_single_ran = false;
  std::string err;
  bool poisoned = false;
  bool has_index_events = false;
  bool fast_ran = false;       // k_cand_fast ran this sweep (events 3 / 13 bracket it)
  long long n_launches = 0; // kernels launched by this handle
  struct exb200_pe* pe = nullptr; // point estimate fed by exb200_run (not owned)
  // saved samples leave the device behind the next sweep: snapshot (D2D) -> pinned buffer on the
  // copy stream -> the caller's array by a host thread
  cudaStream_t copy_stream = nullptr;
  bool pipeline_ready = false;  // every buffer and event of the sample pipeline exists
  int n_sm = 148;               // multiprocessors of the device (grid sizes of the persistent kernels)
  int* d_snap[2] = {nullptr, nullptr};
  int* pin_snap[2] = {nullptr, nullptr};
  cudaEvent_t ev_snap[2]{}, ev_d2h[2]{};
};

namespace {

int fail(exb200_handle* h, int code, const std::string& msg) {
  static std::mutex mu; // attribute tables are built by concurrent threads
  std::lock_guard<std::mutex> lk(mu);
  if (h) { h->err = msg; }
  else g_create_error = msg;
  return code;
}

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      h->poisoned = true;                                                                          \
      return fail(h, EXB200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));        \
    }                                                                                              \
  } while (0)

// fn(t, T) on T host threads (table construction at create time)
template <class F> void host_parallel(size_t work, F fn) {
  // (a share of the cores: with one process per GPU every rank builds its tables at the same time)
  static const int T_max = [] {
    if (const char* e = std::getenv("EXB200_HOST_THREADS")) return std::max(1, std::atoi(e));
    int n_dev = 1;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess) { cudaGetLastError(); n_dev = 1; }
    const int hw = (int)std::max(1u, std::thread::hardware_concurrency());
    return std::max(2, std::min(16, hw / std::max(1, n_dev)));
  }();
  int T = T_max;
  if (work < (1u << 16)) T = 1;
  if (T == 1) { fn(0, 1); return; }
  std::vector<std::thread> th;
  for (int t = 1; t < T; t++) th.emplace_back([=] { fn(t, T); });
  fn(0, T);
  for (auto& x : th) x.join();
}

// Slabs (and the pinned staging rows of the history) of destroyed handles are kept for the next handle
// of the process on the same device: cudaMalloc / cudaFree of a chain's ten gigabytes cost 0.05-0.5 s
// per handle, which is what a caller that runs one chain after the other (run_inference in a loop)
// would pay every time.  exb200_trim() gives everything back to the driver.
struct MemPool {
  struct Dev { int device; char* base; size_t size; };
  struct Pin { void* p; size_t size; };
  std::mutex mu;
  std::vector<Dev> dev;
  std::vector<Pin> pin;
  size_t dev_bytes = 0;
  static constexpr size_t MAX_DEV_BYTES = (size_t)64 << 30;
  char* take_dev(int device, size_t want, size_t at_least, size_t* size) { // smallest slab >= want, else >= at_least
    std::lock_guard<std::mutex> lk(mu);
    int best = -1;
    for (int pass = 0; pass < 2 && best < 0; pass++)
      for (int i = 0; i < (int)dev.size(); i++)
        if (dev[i].device == device && dev[i].size >= (pass ? at_least : want) && (best < 0 || dev[i].size < dev[best].size)) best = i;
    if (best < 0) return nullptr;
    char* q = dev[best].base;
    *size = dev[best].size;
    dev_bytes -= dev[best].size;
    dev.erase(dev.begin() + best);
    return q;
  }
  bool give_dev(int device, char* base, size_t size) {
    std::lock_guard<std::mutex> lk(mu);
    if (dev_bytes + size > MAX_DEV_BYTES) return false;
    dev.push_back({device, base, size});
    dev_bytes += size;
    return true;
  }
  void* take_pin(size_t size) {
    std::lock_guard<std::mutex> lk(mu);
    for (size_t i = 0; i < pin.size(); i++)
      if (pin[i].size >= size && pin[i].size <= 2 * size) { void* q = pin[i].p; pin.erase(pin.begin() + (long)i); return q; }
    return nullptr;
  }
  void give_pin(void* p, size_t size) {
    std::lock_guard<std::mutex> lk(mu);
    if (pin.size() < 8) pin.push_back({p, size}); else cudaFreeHost(p);
  }
  void trim() {
    std::lock_guard<std::mutex> lk(mu);
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto& d : dev) { cudaSetDevice(d.device); cudaFree(d.base); }
    for (auto& q : pin) cudaFreeHost(q.p);
    cudaSetDevice(cur);
    dev.clear(); pin.clear(); dev_bytes = 0;
  }
};
MemPool g_pool;

template <class T> int dalloc(exb200_handle* h, T** p, size_t n) {
  const size_t bytes = (std::max<size_t>(n, 1) * sizeof(T) + 255) & ~(size_t)255;
  std::lock_guard<std::mutex> lk(h->alloc_mu);
  if (h->slabs.empty() || h->slabs.back().size - h->slabs.back().used < bytes) {
    // slabs double from 64 MiB to 2 GiB; a larger request gets a slab of its own size
    size_t want = h->slabs.empty() ? ((size_t)64 << 20) : std::min<size_t>(h->slabs.back().size * 2, (size_t)2 << 30);
    want = std::max(want, bytes);
    size_t got = 0;
    void* q = g_pool.take_dev(h->device, want, bytes, &got);
    if (q) {
      // (a slab that comes back from an earlier handle starts out zero like a fresh one usually does)
      // (on the handle's own stream - it does not synchronise with the default stream - so that the
      // clearing is ordered before everything the handle does with the slab)
      cudaError_t e = cudaMemsetAsync(q, 0, got, h->stream);
      if (e == cudaSuccess && !h->stream) e = cudaStreamSynchronize(0);
      if (e != cudaSuccess) { cudaGetLastError(); cudaFree(q); q = nullptr; }
      else want = got;
    }
    if (!q) {
      cudaError_t e = cudaMalloc(&q, want);
      if (e != cudaSuccess && want > bytes) { want = bytes; e = cudaMalloc(&q, want); }
      if (e != cudaSuccess) { // the pool may be holding what is missing
        cudaGetLastError();
        g_pool.trim();
        e = cudaMalloc(&q, want);
      }
      if (e != cudaSuccess) return fail(h, EXB200_ERR_OOM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    }
    h->slabs.push_back({(char*)q, want, 0});
  }
  exb200_handle::Slab& sl = h->slabs.back();
  *p = (T*)(sl.base + sl.used);
  sl.used += bytes;
  return EXB200_OK;
}
template <class T> int upload(exb200_handle* h, T** p, const T* src, size_t n) {
  int rc = dalloc(h, p, n);
  if (rc) return rc;
  if (n) CK(cudaMemcpyAsync(*p, src, n * sizeof(T), cudaMemcpyHostToDevice, h->stream));
  return EXB200_OK;
}
// kernel with a compile-time switch for finite concentrations, chosen by the model's state
#define KLAUNCH_DP(h, kern, grid, block, smem, ...)                   \
  do {                                                                 \
    if ((h)->S.has_dp) KLAUNCH(h, kern<true>, grid, block, smem, __VA_ARGS__);  \
    else KLAUNCH(h, kern<false>, grid, block, smem, __VA_ARGS__);      \
  } while (0)
#define KLAUNCH(h, kern, grid, block, smem, ...)                      \
  do {                                                                 \
    kern<<<(grid), (block), (smem), (h)->stream>>>(__VA_ARGS__);       \
    (h)->n_launches++;                                                 \
  } while (0)

inline int nblk(long long n, int tpb = TPB) { return (int)std::max<long long>(1, (n + tpb - 1) / tpb); }

// exclusive prefix sum on the handle's stream
template <class T_IN> int scan_exclusive(exb200_handle* h, const T_IN* in, int* out, int n) {
  const int nb = nblk(n, SCAN_TILE);
  if ((size_t)nb > h->bsum_cap) {
    int rc = dalloc(h, &h->d_bsum, (size_t)nb * 2);
    if (rc) return rc;
    h->bsum_cap = (size_t)nb * 2;
  }
  KLAUNCH(h, k_scan_reduce<T_IN>, nb, TPB, 0, in, n, h->d_bsum);
  KLAUNCH(h, k_scan_top, 1, 1024, 0, h->d_bsum, nb);
  KLAUNCH(h, k_scan_apply<T_IN>, nb, TPB, 0, in, out, n, h->d_bsum);
  return EXB200_OK;
}


// ---- large copies between the caller's pageable arrays and the device.  cudaMemcpyAsync from / to
// pageable memory goes through the driver's own staging at ~4 GB/s (0.2 s for the 0.9 GB model of
// 10 M records, 0.1 s for the state on the way back); here a few host threads move the data through a
// process-wide ring of pinned buffers, memcpy and DMA overlapped, at several times that.
struct Stager {
  static constexpr size_t CHUNK = (size_t)8 << 20;
  static constexpr int THREADS = 4, NB = 2 * THREADS;
  char* buf[NB] = {};
  cudaEvent_t ev[NB] = {};
  bool ready = false, failed = false;
  std::mutex mu;
  bool init() {
    if (ready || failed) return ready;
    for (int b = 0; b < NB; b++)
      if (cudaHostAlloc((void**)&buf[b], CHUNK, cudaHostAllocPortable) != cudaSuccess ||
          cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); failed = true; return false; }
    ready = true;
    return true;
  }
};
Stager g_stager;

// to_device: dst is device memory, src the caller's array; else the other way round
cudaError_t staged_copy(void* dst, const void* src, size_t bytes, bool to_device, int device, cudaStream_t st) {
  const cudaMemcpyKind kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
  if (bytes < 4 * Stager::CHUNK) return cudaMemcpyAsync(dst, src, bytes, kind, st);
  std::lock_guard<std::mutex> lk(g_stager.mu); // one large copy at a time through the ring
  if (!g_stager.init()) return cudaMemcpyAsync(dst, src, bytes, kind, st);
  const size_t n_chunks = (bytes + Stager::CHUNK - 1) / Stager::CHUNK;
  cudaError_t err[Stager::THREADS];
  auto work = [&](int t) {
    cudaSetDevice(device);
    cudaError_t e = cudaSuccess;
    size_t pending_off[2] = {0, 0}, pending_n[2] = {0, 0}; // device-to-host: chunks in flight per buffer
    int k = 0;
    for (size_t c = (size_t)t; c < n_chunks && e == cudaSuccess; c += Stager::THREADS, k ^= 1) {
      const int b = 2 * t + k;
      const size_t off = c * Stager::CHUNK, n = std::min(Stager::CHUNK, bytes - off);
      e = cudaEventSynchronize(g_stager.ev[b]); // the buffer's previous transfer is done
      if (e != cudaSuccess) break;
      if (to_device) {
        std::memcpy(g_stager.buf[b], (const char*)src + off, n);
        e = cudaMemcpyAsync((char*)dst + off, g_stager.buf[b], n, kind, st);
      } else {
        if (pending_n[k]) std::memcpy((char*)dst + pending_off[k], g_stager.buf[b], pending_n[k]);
        e = cudaMemcpyAsync(g_stager.buf[b], (const char*)src + off, n, kind, st);
        pending_off[k] = off; pending_n[k] = n;
      }
      if (e == cudaSuccess) e = cudaEventRecord(g_stager.ev[b], st);
    }
    for (int q = 0; q < 2 && e == cudaSuccess; q++) { // drain
      e = cudaEventSynchronize(g_stager.ev[2 * t + q]);
      if (e == cudaSuccess && !to_device && pending_n[q]) std::memcpy((char*)dst + pending_off[q], g_stager.buf[2 * t + q], pending_n[q]);
    }
    err[t] = e;
  };
  std::vector<std::thread> th;
  for (int t = 1; t < Stager::THREADS; t++) th.emplace_back(work, t);
  work(0);
  for (auto& x : th) x.join();
  for (int t = 0; t < Stager::THREADS; t++) if (err[t] != cudaSuccess) return err[t];
  return cudaSuccess;
}

double prior_existing_host(const exb200_clust& c, int n) {
  if (c.kind == EXB200_CLUST_PITMAN_YOR) return std::fabs((double)n - c.d);
  if (c.kind == EXB200_CLUST_GEN_COUPON) return (double)n + c.kappa;
  return 1.0;
}

void refresh_conc_dev(exb200_handle* h) {
  h->S.has_dp = 0;
  h->S.dp_mask = 0;
  for (int a = 0; a < MAX_ATTRS; a++) {
    h->S.conc[a] = a < h->A ? h->conc[a] : std::numeric_limits<double>::infinity();
    if (a < h->A && std::isfinite(h->conc[a])) { h->S.has_dp = 1; h->S.dp_mask |= 1u << a; }
  }
}

void refresh_clust_dev(exb200_handle* h) {
  ClustDev& c = h->S.cl;
  c.kind = h->cl.kind; c.alpha = h->cl.alpha; c.d = h->cl.d; c.m = h->cl.m; c.kappa = h->cl.kappa;
  for (int n = 0; n < LPE_MAX; n++) c.logpe[n] = std::log(prior_existing_host(h->cl, n));
  h->S.use_kbounds = !(h->cl.kind == EXB200_CLUST_PITMAN_YOR && h->cl.d == 0.0);
}

struct NotFinal { // records whose link is not final yet (selection of the tail's first list)
  const uint8_t* fin;
  __device__ __forceinline__ bool operator()(int r) const { return fin[r] != 1; }
};

int check_device_error(exb200_handle* h) {
  int e[2] = {0, 0};
  CK(cudaMemcpyAsync(e, h->S.dev_err, sizeof e, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (e[0]) {
    h->poisoned = true;
    return fail(h, e[0], "no valid outcome for record/entity " + std::to_string(e[1]) +
                             " (all weights zero or a rejection sampler did not terminate)");
  }
  return EXB200_OK;
}

// ---- per-attribute tables (ConstantAttributeIndex / NonConstantAttributeIndex + alias tables)
// Everything that depends on the entity-value pmf phi of attribute a - what Cache::update
// (cache.cpp:20-27) refreshes plus the alias tables and prefix sums of the O(1) schemes - computed
// on the host and copied into the attribute's device buffers.  Called at creation and, with a
// Dirichlet prior (Entities::update_distributions, entities.cpp:133-156), after every sweep's draw.
int refresh_phi(exb200_handle* h, int a) {
  HostAttr& ha = h->attrs[a];
  DevAttr& da = h->h_attrs[a];
  const int D = ha.D;
  std::vector<double> logphi(D), logE(D);
  double Z = 0.0;
  for (int v = 0; v < D; v++) {
    if (!(ha.phi[v] >= 0.0)) return fail(h, EXB200_ERR_INVALID, "attribute " + std::to_string(a) + ": invalid pmf");
    logphi[v] = std::log(ha.phi[v]);
    Z += ha.phi[v] / ha.om[v]; // ConstantAttributeIndex::update, attribute_index.cpp:26-34
  }
  RawVec<double> cs, cs2, csn;
  cs.resize(ha.p.size()); cs2.resize(ha.p.size()); csn.resize(ha.p.size());
  host_parallel(ha.p.size() + (size_t)D, [&](int t, int T) {
    const int w_lo = (int)((int64_t)D * t / T), w_hi = (int)((int64_t)D * (t + 1) / T);
    for (int w = w_lo; w < w_hi && !ha.is_const; w++) {
      double c = 0.0, c2 = 0.0, cn = 0.0;
      for (int j = ha.row[w]; j < ha.row[w + 1]; j++) {
        const double base = ha.phi[ha.col[j]] * ha.mef[ha.col[j]] * ha.p[j];
        c += base;
        cs[j] = c;
        c2 += base * (ha.mef[ha.col[j]] * ha.p[j]);
        cs2[j] = c2;
        cn += ha.phi[ha.col[j]] * ha.p[j];
        csn[j] = cn;
      }
    }
    for (int w = w_lo; w < w_hi; w++) {
      double e;
      if (ha.is_const) // get_exp_distortion_prob, attribute_index.cpp:55-62
        e = ha.exclude ? Z * ha.psi[w] - ha.phi[w] * ha.psi[w] / ha.om[w] : ha.psi[w];
      else {           // attribute_index.cpp:94-101
        e = 0.0;
        for (int j = ha.row[w]; j < ha.row[w + 1]; j++) e += ha.phi[ha.col[j]] * ha.p[j];
      }
      logE[w] = std::log(e);
    }
  });
  const int nt = ha.is_const ? NMAX_FAST + 1 : 1;
  std::vector<double> g((size_t)nt * D), aprob((size_t)nt * D);
  std::vector<int> aalias((size_t)nt * D);
  for (int n = 0; n < nt; n++) {
    double G = 0.0;
    for (int v = 0; v < D; v++) {
      g[(size_t)n * D + v] = n == 0 ? ha.phi[v] : g[(size_t)(n - 1) * D + v] / ha.om[v];
      G += g[(size_t)n * D + v];
    }
    da.G[n] = G;
    if (!alias_build(&g[(size_t)n * D], D, &aprob[(size_t)n * D], &aalias[(size_t)n * D]))
      return fail(h, EXB200_ERR_INVALID, "attribute " + std::to_string(a) + ": invalid entity distribution");
  }
  cudaStream_t st = h->stream;
#define PUT(field, vec) if (!(vec).empty()) CK(cudaMemcpyAsync(const_cast<void*>((const void*)da.field), (vec).data(), sizeof((vec)[0]) * (vec).size(), cudaMemcpyHostToDevice, st));
  PUT(phi, ha.phi) PUT(logphi, logphi) PUT(logE, logE) PUT(cs, cs) PUT(cs2, cs2) PUT(csn, csn) PUT(g, g) PUT(aprob, aprob) PUT(aalias, aalias)
#undef PUT
  if (h->d_attrs) CK(cudaMemcpyAsync(h->d_attrs + a, &da, sizeof(DevAttr), cudaMemcpyHostToDevice, st)); // G[] changed
  CK(cudaStreamSynchronize(st)); // the host vectors go out of scope
  return EXB200_OK;
}

// ---- per-attribute tables (ConstantAttributeIndex / NonConstantAttributeIndex + alias tables)
int build_attr(exb200_handle* h, const exb200_attr& in, int a) {
  HostAttr& ha = h->attrs[a];
  DevAttr& da = h->h_attrs[a];
  const int D = in.domain_size;
  if (D < 1) return fail(h, EXB200_ERR_INVALID, "attribute " + std::to_string(a) + ": empty domain");
  if (std::isfinite(in.distort_dist_conc) && !in.exclude_entity_value) // comment at links.cpp:251
    return fail(h, EXB200_ERR_INVALID, "attribute " + std::to_string(a) + ": a finite distort_dist_prior requires exclude_entity_value");
  if (!(in.distort_dist_conc > 0.0)) return fail(h, EXB200_ERR_INVALID, "attribute " + std::to_string(a) + ": concentration must be positive");
  ha.D = D; ha.is_const = in.is_constant != 0; ha.exclude = in.exclude_entity_value != 0;
  ha.s1 = in.distort_prob_shape1; ha.s2 = in.distort_prob_shape2;
  ha.conc_shape = in.conc_prior_shape; ha.conc_rate = in.conc_prior_rate;
  h->conc[a] = in.distort_dist_conc;
  ha.dirichlet = in.entity_dist_kind == EXB200_ENTDIST_DIRICHLET;
  if (ha.dirichlet) {
    if (!in.entity_dist) return fail(h, EXB200_ERR_INVALID, "Invalid entity_dist_prior"); // entities.cpp:195
    ha.dir_alpha.assign(in.entity_dist, in.entity_dist + D);
    for (double al : ha.dir_alpha)
      if (!(al > 0.0)) return fail(h, EXB200_ERR_INVALID, "rdirichlet: concentration parameter must be positive"); // dirichlet.h:21
  }
  ha.psi.assign(in.probs, in.probs + D);
  const double* phi0 = (in.entity_dist && !ha.dirichlet) ? in.entity_dist : in.probs;
  ha.phi.assign(phi0, phi0 + D);
  ha.mef.assign(D, 1.0);
  if (!ha.is_const) ha.mef.assign(in.max_exp_factor, in.max_exp_factor + D);
  ha.om.resize(D);
  std::vector<double> logpsi(D), log1mpsi(D);
  for (int v = 0; v < D; v++) {
    if (!(ha.psi[v] > 0.0 && ha.psi[v] < 1.0))
      return fail(h, EXB200_ERR_INVALID, "attribute " + std::to_string(a) + ": invalid pmf");
    ha.om[v] = 1.0 - ha.psi[v];
    logpsi[v] = std::log(ha.psi[v]);
    log1mpsi[v] = std::log(ha.om[v]);
  }
  // inverted neighbourhood CSR (attribute_index.cpp:125-140)
  ha.row.assign(D + 1, 0);
  RawVec<double> logp;
  if (!ha.is_const) {
    const int64_t nnz = in.close_row_ptr[D];
    if (nnz >= (int64_t)std::numeric_limits<int>::max())
      return fail(h, EXB200_ERR_INVALID, "neighbourhood index too large");
    // every thread owns a range of output rows w and scans the whole input in order, so the
    // entries of a row stay in ascending v whatever the number of threads
    std::vector<int> bad(64, 0);
    host_parallel((size_t)nnz, [&](int t, int T) {
      const int w_lo = (int)((int64_t)D * t / T), w_hi = (int)((int64_t)D * (t + 1) / T);
      for (int64_t j = 0; j < nnz; j++) {
        const int w = in.close_val_ids[j];
        if (w < 0 || w >= D) { bad[t] = 1; continue; }
        if (w >= w_lo && w < w_hi) ha.row[w + 1]++;
      }
    });
    for (int b : bad) if (b) return fail(h, EXB200_ERR_INVALID, "close_val_ids out of range");
    for (int w = 0; w < D; w++) ha.row[w + 1] += ha.row[w];
    ha.col.resize(nnz); ha.p.resize(nnz); logp.resize(nnz);
    std::vector<int> cur(ha.row.begin(), ha.row.end() - 1);
    host_parallel((size_t)nnz, [&](int t, int T) {
      const int w_lo = (int)((int64_t)D * t / T), w_hi = (int)((int64_t)D * (t + 1) / T);
      for (int v = 0; v < D; v++)
        for (int64_t j = in.close_row_ptr[v]; j < in.close_row_ptr[v + 1]; j++) {
          const int w = in.close_val_ids[j];
          if (w < w_lo || w >= w_hi) continue;
          const int q = cur[w]++;
          ha.col[q] = v; ha.p[q] = in.close_probs[j]; logp[q] = std::log(in.close_probs[j]);
        }
    });
  }
  std::vector<double> pprob(D);
  std::vector<int> palias(D);
  if (!alias_build(ha.psi.data(), D, pprob.data(), palias.data()))
    return fail(h, EXB200_ERR_INVALID, "attribute " + std::to_string(a) + ": invalid probs");
  da.D = D; da.is_const = ha.is_const; da.exclude = ha.exclude; da.pad = 0;
  const int nt = ha.is_const ? NMAX_FAST + 1 : 1;
  int rc = 0;
  double* dp; int* ip;
#define UPD(field, vec) rc = upload(h, &dp, (vec).data(), (vec).size()); if (rc) return rc; da.field = dp;
#define UPI(field, vec) rc = upload(h, &ip, (vec).data(), (vec).size()); if (rc) return rc; da.field = ip;
#define NEWD(field, n) rc = dalloc(h, &dp, (size_t)(n)); if (rc) return rc; da.field = dp;
  UPD(psi, ha.psi) UPD(om, ha.om) UPD(logpsi, logpsi) UPD(log1mpsi, log1mpsi) UPD(mef, ha.mef)
  UPI(row, ha.row) UPI(col, ha.col) UPD(p, ha.p) UPD(logp, logp) UPD(pprob, pprob) UPI(palias, palias)
  NEWD(phi, D) NEWD(logphi, D) NEWD(logE, D) NEWD(cs, ha.p.size()) NEWD(cs2, ha.p.size()) NEWD(csn, ha.p.size())
  NEWD(g, (size_t)nt * D) NEWD(aprob, (size_t)nt * D)
  rc = dalloc(h, &ip, (size_t)nt * D); if (rc) return rc; da.aalias = ip;
#undef UPD
#undef UPI
#undef NEWD
  rc = dalloc(h, &da.lt_diff, (size_t)h->F * D); if (rc) return rc;
  rc = dalloc(h, &da.lt_same, (size_t)h->F * D); if (rc) return rc;
  rc = dalloc(h, &da.sg, (size_t)2 * h->F * D); if (rc) return rc;
  CK(cudaStreamSynchronize(h->stream)); // the host vectors go out of scope
  return refresh_phi(h, a);
}

// phi_a ~ Dirichlet(count_a(v) + alpha_v) (entities.cpp:133-156, dirichlet.h:12-34): counts from the
// device, Gamma variates from the host stream
int draw_entity_dist(exb200_handle* h, int a, const std::vector<int>& counts) {
  HostAttr& ha = h->attrs[a];
  double total = 0.0;
  std::vector<double> g(ha.D);
  for (int v = 0; v < ha.D; v++) { g[v] = h->hs.gamma((double)counts[v] + ha.dir_alpha[v], 1.0); total += g[v]; }
  for (int v = 0; v < ha.D; v++) ha.phi[v] = g[v] / total;
  return refresh_phi(h, a);
}

// Entities::update_distributions and DistortDistConcs::update (distort_dist_concs.cpp:32-94; value
// counts per cluster) after the entity attributes of the sweep are drawn
int update_dists_and_concs(exb200_handle* h) {
  DevState& S = h->S;
  const int N = h->N;
  cudaStream_t st = h->stream;
  for (int a = 0; a < h->A; a++) {
    if (!h->attrs[a].dirichlet) continue;
    std::vector<int> counts(h->attrs[a].D);
    CK(cudaMemsetAsync(h->d_hist, 0, sizeof(int) * counts.size(), st));
    KLAUNCH(h, k_ent_hist, nblk(N), TPB, 0, S, a, h->d_hist);
    CK(cudaMemcpyAsync(counts.data(), h->d_hist, sizeof(int) * counts.size(), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    int rc = draw_entity_dist(h, a, counts); if (rc) return rc;
  }
  bool changed = false;
  for (int a = 0; a < h->A; a++) {
    const HostAttr& ha = h->attrs[a];
    if (!(ha.conc_shape > 0)) continue;
    std::vector<int> nd(N);
    unsigned long long zeta = 0;
    CK(cudaMemsetAsync(h->d_cl_counts, 0, sizeof(unsigned long long), st));
    KLAUNCH(h, k_conc_aux, nblk(N), TPB, 0, S, a, S.plo, h->d_cl_counts); // plo is free outside the links update
    CK(cudaMemcpyAsync(nd.data(), S.plo, sizeof(int) * N, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&zeta, h->d_cl_counts, sizeof zeta, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    long long sum_logt_fx = 0; // fixed point (2^-40): independent of the summation order
    for (int e = 0; e < N; e++)
      if (nd[e] > 0) { // rbeta(b, 0) = 1 otherwise: log 1 = 0
        KeyedStream k{h->seed, (uint32_t)e, (uint32_t)h->iteration, PH_CONC_BETA, (uint32_t)a};
        double te = k.beta(h->conc[a], (double)nd[e]);
        if (te < 1e-300) te = 1e-300;
        sum_logt_fx += std::llround(std::log(te) * 1099511627776.0);
      }
    const double shape = ha.conc_shape + (double)zeta;
    const double rate = ha.conc_rate - (double)sum_logt_fx / 1099511627776.0;
    h->conc[a] = h->hs.gamma(shape, 1.0 / rate);
    changed = true;
  }
  if (changed) refresh_conc_dev(h);
  return EXB200_OK;
}

int upload_theta(exb200_handle* h) {
  CK(cudaMemcpyAsync(h->d_theta, h->theta.data(), sizeof(double) * h->F * h->A, cudaMemcpyHostToDevice, h->stream));
  for (int a = 0; a < h->A; a++)
    KLAUNCH(h, k_theta_tables, nblk((long long)h->F * h->attrs[a].D), TPB, 0, h->S, a);
  return EXB200_OK;
}

// ---- 1. cluster-prior hyper-parameters: ClustParams::update (clust_params.cpp:37-72, 74-121)
int update_clust(exb200_handle* h) {
  exb200_clust& c = h->cl;
  const int N = h->N, K = h->K;
  const bool py = c.kind == EXB200_CLUST_PITMAN_YOR, gc = c.kind == EXB200_CLUST_GEN_COUPON;
  const bool has_a = py && c.alpha_prior_shape > 0, has_d = py && c.d_prior_shape1 > 0;
  const bool has_k = gc && c.kappa_prior_shape > 0, has_m = gc && c.m_prior_size > 0;
  if (!(has_a || has_d || has_k || has_m)) return EXB200_OK;
  unsigned long long cnt[2] = {0, 0};
  if (py || has_k) {
    CK(cudaMemsetAsync(h->d_cl_counts, 0, sizeof cnt, h->stream));
    KLAUNCH(h, k_clust_counts, nblk(std::max(N, K)), TPB, 0, h->S, K, py ? 1 : 0, (has_d || has_k) ? 1 : 0,
                                                                 h->d_cl_counts);
    CK(cudaMemcpyAsync(cnt, h->d_cl_counts, sizeof cnt, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  if (py) {
    const double succ = (double)cnt[0], failn = (double)(K - 1) - succ, vfail = (double)cnt[1];
    if (has_a) { // :92-101
      const double w = h->hs.beta(c.alpha + 1.0, (double)N - 1.0);
      const double shape = c.alpha_prior_shape + succ, rate = c.alpha_prior_rate - std::log(w);
      c.alpha = h->hs.gamma(shape, 1.0 / rate);
    }
    if (has_d) c.d = h->hs.beta(c.d_prior_shape1 + failn, c.d_prior_shape2 + vfail); // :117-120
  } else {
    const double w = h->hs.beta(c.m * c.kappa + 1.0, (double)N - 1.0); // :45
    if (has_k) { // :61-63
      const double shape = c.kappa_prior_shape + (double)K - 1.0 + (double)cnt[1];
      const double rate = c.kappa_prior_rate - c.m * std::log(w);
      c.kappa = h->hs.gamma(shape, 1.0 / rate);
    }
    if (has_m) { // :66-71 (m_ is an int in the reference)
      const double sz = c.m_prior_size + (double)K - 1.0;
      const double pb = 1.0 - (1.0 - c.m_prior_prob) * std::pow(w, c.kappa);
      c.m = (double)(int)(h->hs.negbinom(sz, pb) + (double)K);
    }
  }
  return EXB200_OK;
}

int ensure_table(exb200_handle* h, int tab) {
  TableHost& t = h->tables[tab];
  if (t.allocated) return EXB200_OK;
  int rc = dalloc(h, &t.d.ptr, (size_t)t.d.B + 1); if (rc) return rc;
  rc = dalloc(h, &t.d.ids, (size_t)2 * h->N); if (rc) return rc;
  t.allocated = true;
  CK(cudaMemcpyAsync(h->d_tables + tab, &t.d, sizeof(DevTable), cudaMemcpyHostToDevice, h->stream));
  return EXB200_OK;
}

// CSR form of the given blocking tables (count / scan / fill)
int build_csr_tables(exb200_handle* h, const std::vector<int>& tabs) {
  DevState& S = h->S;
  const int N = h->N;
  cudaStream_t st = h->stream;
  if (tabs.empty()) return EXB200_OK;
  for (int t : tabs) { int rc = ensure_table(h, t); if (rc) return rc; }
  // tables with many buckets are built by partitioning (k_part_a / k_part_b); the few small ones
  // (their bucket arrays stay in the L2) by global atomics (k_index_count / k_index_fill)
  std::vector<int> large, small, tiny;
  for (int t : tabs) {
    const uint32_t B = h->tables[t].d.B;
    const bool part = h->index_partitioned && B >= 16u * PART_SIZE && B <= (uint32_t)MAX_PARTS * PART_SIZE && (int)large.size() < PA_MAX_TABS;
    if (part) large.push_back(t);
    else if (h->index_partitioned && B <= (uint32_t)SMALL_B_MAX) { tiny.push_back(t); small.push_back(t); }
    else small.push_back(t);
  }
  // `small` keeps the tiny tables for the memset / scan; they are removed from the atomic kernels' list
  std::vector<int> atomic_tabs;
  for (int t : small) if (std::find(tiny.begin(), tiny.end(), t) == tiny.end()) atomic_tabs.push_back(t);
  for (int t : small) CK(cudaMemsetAsync(h->tables[t].d.ptr, 0, sizeof(int) * ((size_t)h->tables[t].d.B + 1), st));
  if (!large.empty()) {
    const long long cap = 2LL * N;
    if ((int)large.size() > h->pb_tables_cap) {
      const int want = std::min(PA_MAX_TABS, ((int)large.size() + 7) / 8 * 8);
      int rc = dalloc(h, &h->pb.pairs, (size_t)want * cap); if (rc) return rc;
      h->pb_tables_cap = want;
    }
    PartBuild& P = h->pb;
    P.n_tab = (int)large.size(); P.tabs = h->d_part_tabs; P.cap = cap;
    CK(cudaMemcpyAsync(h->d_part_tabs, large.data(), sizeof(int) * large.size(), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(P.part_cnt, 0, sizeof(int) * (size_t)P.n_tab * MAX_PARTS, st));
  }
  // the tiny tables in groups that share one pass over the items (their counters fit shared memory together)
  std::vector<SmallTabs> tiny_groups;
  for (int t : tiny) {
    const int B = (int)h->tables[t].d.B;
    if (tiny_groups.empty() || tiny_groups.back().n == SMALL_TABS_MAX || tiny_groups.back().off[tiny_groups.back().n] + B > SMALL_SMEM_INTS) {
      SmallTabs g{}; g.n = 0; g.off[0] = 0;
      tiny_groups.push_back(g);
    }
    SmallTabs& g = tiny_groups.back();
    g.tab[g.n] = t; g.off[g.n + 1] = g.off[g.n] + B; g.n++;
  }
  CK(cudaMemcpyAsync(h->d_active, atomic_tabs.data(), sizeof(int) * atomic_tabs.size(), cudaMemcpyHostToDevice, st));
  if (!atomic_tabs.empty()) KLAUNCH(h, k_index_count, nblk(2LL * N), TPB, 0, S, h->d_active, (int)atomic_tabs.size());
  for (const SmallTabs& g : tiny_groups) KLAUNCH(h, k_index_small, h->n_sm * 8, TPB, sizeof(int) * g.off[g.n], S, g, 0);
  if (!large.empty()) {
    KLAUNCH(h, k_part_a, h->n_sm * 8, TPB, sizeof(int) * large.size() * MAX_PARTS, S, h->pb, 0);
    KLAUNCH(h, k_part_scan, (int)large.size(), TPB, 0, h->pb);
  }
  for (int t : small) {
    int rc = scan_exclusive<int>(h, h->tables[t].d.ptr, h->tables[t].d.ptr, (int)h->tables[t].d.B);
    if (rc) return rc;
  }
  if (!atomic_tabs.empty()) KLAUNCH(h, k_index_fill, nblk(2LL * N), TPB, 0, S, h->d_active, (int)atomic_tabs.size());
  for (const SmallTabs& g : tiny_groups) KLAUNCH(h, k_index_small, h->n_sm * 8, TPB, sizeof(int) * g.off[g.n], S, g, 1);
  if (!large.empty()) {
    KLAUNCH(h, k_part_a, h->n_sm * 8, TPB, sizeof(int) * large.size() * MAX_PARTS, S, h->pb, 1);
    k_part_b<<<dim3(MAX_PARTS, (unsigned)large.size()), TPB, sizeof(int) * PART_SIZE, st>>>(S, h->pb);
    h->n_launches++;
  }
  return EXB200_OK;
}

// serial path: possible links of up to WIDE_NB records among all the items (counting pass)
int wide_count(exb200_handle* h, const int* recs, int nb, int* totals) {
  DevState& S = h->S;
  if (!S.wl_log_k) { // the hit logs of the counting pass, allocated when the serial path is first used
    const long long per_piece = (2LL * h->N + WIDE_PIECES - 1) / WIDE_PIECES;
    S.wl_logcap = (int)std::min<long long>(4096, std::max<long long>(64, 8 * per_piece));
    const size_t n = (size_t)WIDE_PIECES * S.wl_logcap;
    int rc = dalloc(h, &S.wl_log_k, n); if (rc) return rc;
    rc = dalloc(h, &S.wl_log_item, n); if (rc) return rc;
    rc = dalloc(h, &S.wl_log_ll, n); if (rc) return rc;
    rc = dalloc(h, &S.wl_logn, (size_t)WIDE_PIECES); if (rc) return rc;
  }
  CK(cudaMemcpyAsync(S.batch_list, recs, sizeof(int) * nb, cudaMemcpyHostToDevice, h->stream));
  KLAUNCH(h, k_wide_scan<false>, WIDE_BLOCKS, TPB, 0, S, S.batch_list, nb);
  KLAUNCH(h, k_wide_offsets, nb, 1024, 0, S);
  CK(cudaMemcpyAsync(totals, S.wl_tot, sizeof(int) * nb, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return EXB200_OK;
}
// ... and the turn of one listed record
void wide_turn(exb200_handle* h, int r, unsigned long long base, int len, int src = 0) {
  const int n_blocks = std::max(1, std::min(WIDE_BLOCKS, (len + 2047) / 2048));
  KLAUNCH_DP(h, k_wide_reduce, n_blocks, TPB, 0, h->S, r, base, len, src);
  KLAUNCH_DP(h, k_wide_commit, 1, TPB, 0, h->S, r, base, len, n_blocks, src);
}

// ... and the turns of several listed records, in the order given, in one persistent kernel
int wide_turns(exb200_handle* h, const std::vector<WideTurn>& turns) {
  if (turns.empty()) return EXB200_OK;
  DevState& S = h->S;
  cudaStream_t st = h->stream;
  if (turns.size() > h->turns_cap) {
    h->turns_cap = std::max<size_t>(1024, 2 * turns.size());
    int rc = dalloc(h, &h->d_turns, h->turns_cap); if (rc) return rc;
    rc = dalloc(h, &h->d_turn_sync, 2 * h->turns_cap); if (rc) return rc;
  }
  if (!h->turns_grid) {
    int per_sm = 0;
    CK(S.has_dp ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_wide_turns<true>, TPB, 0)
                : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_wide_turns<false>, TPB, 0));
    h->turns_grid = std::max(1, std::min(WIDE_BLOCKS, per_sm * h->n_sm));
  }
  CK(cudaMemcpyAsync(h->d_turns, turns.data(), sizeof(WideTurn) * turns.size(), cudaMemcpyHostToDevice, st));
  int n = (int)turns.size();
  // as many blocks as the longest list has pieces of 2048 entries (a grid-wide barrier costs more with more blocks)
  int want = 1;
  for (const WideTurn& w : turns) want = std::max(want, (w.len + 2047) / 2048);
  const int grid = std::min(h->turns_grid, want);
  CK(cudaMemsetAsync(h->d_turn_sync, 0, sizeof(unsigned) * 2 * (size_t)n, st));
  unsigned long long* dbg = nullptr;
  if (h->trace) { // per-turn timers of block 0 (trace only)
    if (h->turn_dbg_cap < (size_t)n) {
      int rc = dalloc(h, &h->d_turn_dbg, 4 * h->turns_cap); if (rc) return rc;
      h->turn_dbg_cap = h->turns_cap;
    }
    dbg = h->d_turn_dbg;
  }
  void* args[] = {(void*)&S, (void*)&h->d_turns, (void*)&n, (void*)&h->d_turn_sync, (void*)&dbg};
  CK(cudaLaunchCooperativeKernel(S.has_dp ? (const void*)k_wide_turns<true> : (const void*)k_wide_turns<false>,
                                 dim3((unsigned)grid), dim3(TPB), args, 0, st));
  h->n_launches++;
  CK(cudaStreamSynchronize(st)); // `turns` and `n` live on the caller's stack
  if (dbg) {
    std::vector<unsigned long long> d(4 * (size_t)n);
    CK(cudaMemcpy(d.data(), dbg, sizeof(unsigned long long) * d.size(), cudaMemcpyDeviceToHost));
    double red = 0, wait = 0, com = 0, gap = 0;
    for (int t = 0; t < n; t++) {
      red += (double)(d[4 * t + 1] - d[4 * t]); wait += (double)(d[4 * t + 2] - d[4 * t + 1]); com += (double)(d[4 * t + 3] - d[4 * t + 2]);
      if (t) gap += (double)(d[4 * t] - d[4 * t - 1]);
    }
    std::fprintf(stderr, "[exb200]   per turn (block 0, us): reduce %.2f, wait for the others %.2f, commit %.2f, to the next turn %.2f; kernel %.1f us for %d turns on %d blocks\n",
                 red / n * 1e-3, wait / n * 1e-3, com / n * 1e-3, gap / std::max(1, n - 1) * 1e-3, (double)(d[4 * (size_t)n - 1] - d[0]) * 1e-3, n, grid);
    for (int t = 0; t < std::min(n, 12); t++)
      std::fprintf(stderr, "[exb200]     turn %d len %d: reduce %.2f wait %.2f commit %.2f\n", t, turns[t].len, (double)(d[4 * t + 1] - d[4 * t]) * 1e-3,
                   (double)(d[4 * t + 2] - d[4 * t + 1]) * 1e-3, (double)(d[4 * t + 3] - d[4 * t + 2]) * 1e-3);
  }
  return EXB200_OK;
}

// Possible links of the wide records that have a bucket, listed from the bucket itself (k_wide_emit),
// sorted by (record, item) and cut into one list per record.  `recs` ascending; on return
// base[i] / len[i] locate record i's list in wlb_id / wlb_ll.  Returns EXB200_OK with *ok = false
// when the lists do not fit the scratch (the caller then scans all items for these records too).
constexpr int WIDE_BUCKET_MAX = 65536;
int wide_bucket_lists(exb200_handle* h, const std::vector<int>& recs, std::vector<int>& base, std::vector<int>& len, std::vector<int>& served, bool* ok) {
  DevState& S = h->S;
  cudaStream_t st = h->stream;
  const int n = (int)recs.size();
  *ok = false;
  if (n > WIDE_BUCKET_MAX) return EXB200_OK;
  const auto t_0 = std::chrono::steady_clock::now();
  if (!h->we_keys[0]) {
    h->we_cap = std::min<unsigned long long>(2ull * h->N + (1ull << 20), 48ull << 20); // (10.4 M entries at 10 M records, sweep 200)
    if (h->emit_scratch > 0) h->we_cap = (unsigned long long)h->emit_scratch; // (tests: lists that do not fit)
    int rc = 0;
    for (int k = 0; k < 2 && !rc; k++) { rc = dalloc(h, &h->we_keys[k], (size_t)h->we_cap); if (!rc) rc = dalloc(h, &h->we_vals[k], (size_t)h->we_cap); }
    if (!rc) rc = dalloc(h, &S.wlb_id, (size_t)h->we_cap);
    if (!rc) rc = dalloc(h, &S.wlb_ll, (size_t)h->we_cap);
    if (!rc) rc = dalloc(h, &h->we_cursor, 1);
    if (!rc) rc = dalloc(h, &h->we_recs, (size_t)WIDE_BUCKET_MAX);
    if (!rc) rc = dalloc(h, &h->we_seg, (size_t)4 * WIDE_BUCKET_MAX); // start, end, served, selection
    if (rc) return rc;
    cub::DeviceRadixSort::SortPairs(nullptr, h->we_tmp_bytes, h->we_keys[0], h->we_keys[1],
                                    reinterpret_cast<unsigned long long*>(h->we_vals[0]), reinterpret_cast<unsigned long long*>(h->we_vals[1]),
                                    (long long)h->we_cap, 0, 64, st);
    char* tmp = nullptr;
    rc = dalloc(h, &tmp, h->we_tmp_bytes); if (rc) return rc;
    h->we_tmp = tmp;
  }
  CK(cudaMemcpyAsync(h->we_recs, recs.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(h->we_cursor, 0, sizeof(unsigned long long), st));
  CK(cudaMemsetAsync(h->we_seg, 0, sizeof(int) * 3 * (size_t)n, st));
  KLAUNCH(h, k_wide_emit, n * WE_BLOCKS, TPB, 0, S, h->we_recs, n, h->we_keys[0], h->we_vals[0], h->we_cursor, h->we_cap, h->we_seg + 2 * n);
  unsigned long long n_emit = 0;
  CK(cudaMemcpyAsync(&n_emit, h->we_cursor, sizeof n_emit, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (h->trace) std::fprintf(stderr, "[exb200]   k_wide_emit: %.2f ms, %llu entries\n",
                             std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_0).count(), n_emit);
  if (n_emit > h->we_cap) return EXB200_OK; // does not fit: the batched scan takes these records
  base.assign(n, 0); len.assign(n, 0); served.assign(n, 0);
  if (n_emit == 0) {
    CK(cudaMemcpyAsync(served.data(), h->we_seg + 2 * n, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  } else {
    int top = 32; // key bits that matter: the item and the batch position
    while ((1ll << (top - 32)) < (long long)n && top < 64) top++;
    size_t bytes = h->we_tmp_bytes;
    CK(cub::DeviceRadixSort::SortPairs(h->we_tmp, bytes, h->we_keys[0], h->we_keys[1],
                                       reinterpret_cast<unsigned long long*>(h->we_vals[0]), reinterpret_cast<unsigned long long*>(h->we_vals[1]),
                                       (long long)n_emit, 0, top, st));
    h->n_launches += 4;
    KLAUNCH(h, k_wide_split, nblk((long long)n_emit), TPB, 0, S, h->we_keys[1], h->we_vals[1], (long long)n_emit, h->we_seg, h->we_seg + n);
    std::vector<int> seg(3 * (size_t)n);
    CK(cudaMemcpyAsync(seg.data(), h->we_seg, sizeof(int) * seg.size(), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (int i = 0; i < n; i++) { base[i] = seg[i]; len[i] = seg[n + i] > 0 ? seg[n + i] - seg[i] : 0; served[i] = seg[2 * n + i]; }
  }
  *ok = true;
  return EXB200_OK;
}

// ---- 2. links: Links::update (links.cpp:477-489)
int update_links(exb200_handle* h) {
  DevState& S = h->S;
  const int N = h->N, NT = h->n_tab;
  cudaStream_t st = h->stream;
  S.K0 = h->K;
  h->tail_ran = false; h->ent_single_ran = false;
  CK(cudaMemsetAsync(S.tab_usage, 0, sizeof(int) * 2 * N_TABLES, st));
  CK(cudaMemsetAsync(S.tab_est, 0, sizeof(float) * N_TABLES, st));
  CK(cudaMemsetAsync(S.counters, 0, sizeof(int) * C_NSLOTS, st));
  CK(cudaMemsetAsync(S.counters64, 0, sizeof(long long) * 66, st));
  {
    const unsigned long long shared_from = (unsigned long long)CAND_INLINE * (unsigned long long)N; // after the per-record slots
    CK(cudaMemcpyAsync(S.pool_cursor, &shared_from, sizeof shared_from, cudaMemcpyHostToDevice, st));
  }
  CK(cudaMemsetAsync(S.referenced, 0, (size_t)2 * N, st));
  CK(cudaEventRecord(h->ev[1], st));
  KLAUNCH(h, k_prep, nblk(N), TPB, sizeof(int) * 3 * ((size_t)1 << S.n_key), S);
  std::vector<int> usage(2 * (size_t)N_TABLES);
  std::vector<float> est(N_TABLES);
  CK(cudaMemcpyAsync(usage.data(), S.tab_usage, sizeof(int) * usage.size(), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(est.data(), S.tab_est, sizeof(float) * est.size(), cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(h->ev[2], st));
  CK(cudaStreamSynchronize(st));
  // Which blocking tables to build.  A table costs atomics for every item, so
  //  * a full key mask gets a table only if sending its records to their pair tables instead would
  //    cost more bucket visits (est) than building it;
  //  * a pair mask wanted by few records is served by the most selective table that is built anyway
  //    and whose mask is a subset of it (bigger buckets, same candidates after the filter).
  std::vector<int> active;
  std::vector<uint16_t> remap(2 * (size_t)N_TABLES, 0);
  {
    const double items = 1.2 * (double)N;
    std::vector<char> built(NT, 0);
    auto keys_of = [&](int t) { return (double)h->tables[t].keys; };
    // a table whose key space is much larger than the item set is built in chain form (one atomicExch
    // per item, tiny chains): cheap to build, and its probes cost one node per item
    auto chainable = [&](int t) { return S.pk_ok && h->opt_fast && h->tables[t].d.B > (uint32_t)SMALL_B_MAX && keys_of(t) >= 16.0 * items; };
    for (int t = 1; t < NT; t++)
      if (usage[t] > 0 && (double)est[t] >= (chainable(t) ? 0.05 : 1.0) * h->table_full_cost * items) { built[t] = 1; active.push_back(t); remap[t] = (uint16_t)t; }
    std::vector<int> wanted;
    for (int t = 1; t < NT; t++) if (usage[N_TABLES + t] > 0) wanted.push_back(t);
    std::sort(wanted.begin(), wanted.end(), [&](int a, int b) { return usage[N_TABLES + a] > usage[N_TABLES + b]; });
    // A built table b serves the pair mask t directly if b is a subset of t.  It also serves it if it
    // has extra key attributes: a record that cannot pin one of them probes the bucket of every
    // possible value (make_probe).  CSR form: ONE extra attribute with a small domain, short buckets.
    // Chain form: up to two extra attributes (a record rarely lacks them: most probes are single).
    // Returns the expected bucket visits per record (-1: not served).
    auto visits_via = [&](int b, int t) -> double {
      const int extra = b & ~t;
      const double per_bucket = std::max(1.0, items / keys_of(b));
      if (extra == 0) return per_bucket;
      if ((b & t) == 0) return -1.0; // shares no pinned attribute with the pair
      double probes = 1.0, expected = 1.0;
      int n_extra = 0;
      for (int pos = 0; pos < S.n_key; pos++)
        if ((extra >> pos) & 1) {
          const double D = (double)h->attrs[S.key_attr[pos]].D;
          probes *= D; n_extra++;
          expected *= 1.0 + 0.05 * D; // a record lacks an extra attribute now and then (missing or distorted)
        }
      if (chainable(b)) return n_extra <= 2 && probes <= 4096.0 ? per_bucket * expected : -1.0; // (make_probe: at most two, MAX_PROBES)
      // (one thread walks each probed bucket: short buckets only)
      return n_extra == 1 && probes <= 256.0 && per_bucket <= 32.0 ? probes * per_bucket : -1.0;
    };
    for (int t : wanted) {
      int best = 0;
      double best_visits = 0.0;
      for (int b : active) {
        const double v = visits_via(b, t);
        if (v > 0.0 && (best == 0 || v < best_visits)) { best_visits = v; best = b; }
      }
      // expected bucket visits through the substitute versus the cost of a table of its own
      const bool too_slow = !best || (double)usage[N_TABLES + t] * best_visits > (chainable(t) ? 0.2 : 1.0) * h->table_pair_cost * items;
      // (a table with few buckets is cheap - two passes with shared-memory histograms - and always
      // cheaper than listing its records' possible links by a scan over all items)
      const bool tiny = h->opt_tiny_always && h->index_partitioned && h->tables[t].d.B <= (uint32_t)SMALL_B_MAX;
      if (too_slow && usage[N_TABLES + t] <= h->table_min_usage && !built[t] && !tiny) {
        // a handful of records: no table for them (remap 0 if nothing serves them: their possible
        // links are then listed by the batched scan over all the items)
        remap[N_TABLES + t] = (uint16_t)(best && best_visits <= 65536.0 ? best : 0);
        continue;
      }
      if (too_slow) { best = t; if (!built[t]) { built[t] = 1; active.push_back(t); } }
      remap[N_TABLES + t] = (uint16_t)best;
    }
  }
  if (h->trace) {
    for (int t = 1; t < NT; t++)
      if (usage[t] > 0 || usage[N_TABLES + t] > 0)
        std::fprintf(stderr, "[exb200] key mask %2d: wanted as full key by %8d (pair cost %.3g visits), as pair by %8d -> served by %d / %d%s, %u buckets\n",
                     t, usage[t], (double)est[t], usage[N_TABLES + t], (int)remap[t], (int)remap[N_TABLES + t],
                     std::find(active.begin(), active.end(), t) != active.end() ? " (built)" : "", h->tables[t].d.B);
  }
  CK(cudaMemcpyAsync(h->d_remap, remap.data(), sizeof(uint16_t) * remap.size(), cudaMemcpyHostToDevice, st));
  h->stats.n_tables = (int)active.size();
  S.n_act_tabs = (int)std::min<size_t>(active.size(), ACT_TABS_MAX);
  for (int k = 0; k < S.n_act_tabs; k++) S.act_tabs[k] = (uint16_t)active[k];
  S.emit_cost_cap = h->emit_cost_cap;
  // Form of every table of this sweep: the chain form for key spaces about as large as the item set
  // (one atomicExch per item, probes walk a short chain of nodes holding packed tuples), the CSR
  // form for the few masks with long buckets (scanned by many lanes).
  std::vector<int> csr_tabs, chain_tabs;
  {
    const double items = 1.2 * (double)N;
    for (int t : active) {
      TableHost& T = h->tables[t];
      const bool chain = S.pk_ok && h->opt_fast && T.d.B > (uint32_t)SMALL_B_MAX && (double)T.keys >= 16.0 * items;
      if (chain) {
        if (!T.d.head) {
          uint32_t cells = 1024;
          while ((double)cells < std::min(1.2 * (double)N, (double)T.keys)) cells <<= 1; // about one cell per indexed item
          T.d.cells = cells;
          unsigned long long km = 0;
          for (int k = 0; k < S.n_key; k++)
            if (t >> k & 1) km |= ((1ull << S.pk_bits[S.key_attr[k]]) - 1ull) << S.pk_shift[S.key_attr[k]];
          T.d.kmask = km;
          int rc = dalloc(h, &T.d.head, (size_t)cells); if (rc) return rc;
          rc = dalloc(h, &T.d.node, (size_t)2 * N); if (rc) return rc;
        }
        chain_tabs.push_back(t);
      } else csr_tabs.push_back(t);
      if (T.d.chain != (chain ? 1 : 0) || T.d.csr != (chain ? 0 : 1)) {
        T.d.chain = chain ? 1 : 0; T.d.csr = chain ? 0 : 1;
        CK(cudaMemcpyAsync(h->d_tables + t, &T.d, sizeof(DevTable), cudaMemcpyHostToDevice, st));
      }
    }
  }
  CK(cudaEventRecord(h->ev[12], st));
  h->has_index_events = !active.empty();
  // packed tuples of the indexed items: what the chain tables store, and what every index builder reads
  S.pk = nullptr; S.pkfl = nullptr;
  if (S.pk_ok && S.ES == 8 && !active.empty()) {
    if (!h->d_pk) {
      int rc = dalloc(h, &h->d_pk, (size_t)2 * N); if (rc) return rc;
      rc = dalloc(h, &h->d_pkfl, (size_t)2 * N); if (rc) return rc;
    }
    KLAUNCH(h, k_chain_pack, nblk(2LL * N), TPB, 0, S, h->d_pk, h->d_pkfl);
    S.pk = h->d_pk; S.pkfl = h->d_pkfl;
  }
  { int rc = build_csr_tables(h, csr_tabs); if (rc) return rc; }
  CK(cudaEventRecord(h->ev[14], st));
  if (!chain_tabs.empty()) {
    for (int t : chain_tabs) CK(cudaMemsetAsync(h->tables[t].d.head, 0xFF, sizeof(int) * (size_t)h->tables[t].d.cells, st));
    CK(cudaMemsetAsync(S.tab_over, 0, sizeof(int) * N_TABLES, st));
    for (int t : chain_tabs) KLAUNCH(h, k_chain_build, nblk(2LL * N), TPB, 0, S, t, h->d_pk, h->d_pkfl);
  }
  CK(cudaEventRecord(h->ev[3], st));
  int hc[C_NSLOTS];
  h->fast_ran = !chain_tabs.empty();
  if (chain_tabs.empty()) KLAUNCH(h, (k_cand<8, 64>), nblk((long long)N, 64 / 8), 64, 0, S, (const int*)nullptr, N);
  else {
    // thread per record over the chain tables; what it cannot finish is listed for the bucket kernels
    const auto t_f = std::chrono::steady_clock::now();
    KLAUNCH(h, k_cand_fast, nblk(N, FAST_TPB), FAST_TPB, 0, S, (const int*)nullptr, N, S.huge_list); // (huge_list is free until k_cand_long)
    CK(cudaEventRecord(h->ev[13], st));
    CK(cudaMemcpyAsync(hc, S.counters, sizeof hc, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (hc[C_NMULTI] > 0) { // second pass: the records that probe several cells, one thread each
      KLAUNCH(h, k_cand_fast, nblk(hc[C_NMULTI], FAST_TPB), FAST_TPB, 0, S, (const int*)S.huge_list, hc[C_NMULTI], (int*)nullptr);
      CK(cudaMemcpyAsync(hc, S.counters, sizeof hc, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
    }
    if (h->trace) std::fprintf(stderr, "[exb200] index + thread-per-record candidates (host lap) %8.2f ms; %d records for warps, %d for blocks, %d for the bucket kernels, %d in the second pass\n",
                               std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_f).count(), hc[C_NPROBE], hc[C_NLONG_TODO], hc[C_NSLOW], hc[C_NMULTI]);
    if (hc[C_NSLOW] > 0) {
      // chains longer than hop_cap / more than FAST_CAND candidates: those tables get their CSR form too
      std::vector<int> over(N_TABLES), lazy;
      CK(cudaMemcpy(over.data(), S.tab_over, sizeof(int) * N_TABLES, cudaMemcpyDeviceToHost));
      for (int t : chain_tabs)
        if (over[t]) {
          lazy.push_back(t);
          h->tables[t].d.csr = 1;
          int rc = ensure_table(h, t); if (rc) return rc;
          CK(cudaMemcpyAsync(h->d_tables + t, &h->tables[t].d, sizeof(DevTable), cudaMemcpyHostToDevice, st));
        }
      int rc = build_csr_tables(h, lazy); if (rc) return rc;
      if (h->trace) std::fprintf(stderr, "[exb200] fast path left %d records to the bucket kernels (%d chain tables also built as CSR)\n",
                                 hc[C_NSLOW], (int)lazy.size());
      KLAUNCH(h, (k_cand<8, 64>), nblk((long long)hc[C_NSLOW], 64 / 8), 64, 0, S, (const int*)S.slow_list, hc[C_NSLOW]);
    }
  }
  CK(cudaEventRecord(h->ev[10], st));
  CK(cudaMemcpyAsync(hc, S.counters, sizeof hc, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  auto t_tr = std::chrono::steady_clock::now();
  auto lap = [&](const char* what, int n) { // EXB200_TRACE: host-side laps between synchronisation points
    if (!h->trace) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[exb200] %-28s %6d records %8.2f ms\n", what, n, std::chrono::duration<double, std::milli>(now - t_tr).count());
    t_tr = now;
  };
  h->stats.n_probe = hc[C_NPROBE];
  h->stats.n_slow = h->fast_ran ? hc[C_NSLOW] : 0;
  h->stats.n_huge = 0;
  if (hc[C_NPROBE] > 0) { // records that probe several buckets: a warp each
    KLAUNCH(h, (k_cand<32, 128>), nblk((long long)hc[C_NPROBE], 128 / 32), 128, 0, S, (const int*)S.probe_list, hc[C_NPROBE]);
    CK(cudaMemcpyAsync(hc, S.counters, sizeof hc, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    lap("multi-bucket probes", hc[C_NPROBE]);
  }
  if (hc[C_NLONG_TODO] > 0) { // records whose candidates need a whole block to stage and sort
    const size_t smem = (size_t)pow2_ceil(std::max(h->long_cap, 1)) * (sizeof(double) + sizeof(int)); // padded for the bitonic sort
    const int n_todo = hc[C_NLONG_TODO];
    KLAUNCH(h, k_cand_long, std::min(hc[C_NLONG_TODO], h->n_sm * 4), TPB, smem, S, hc[C_NLONG_TODO], h->long_cap);
    CK(cudaMemcpyAsync(hc, S.counters, sizeof hc, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    lap("block-staged candidate lists", n_todo);
    if (hc[C_NHUGE] > 0) { // buckets so long that many blocks scan them, HUGE_BATCH records per launch
      const size_t n_huge = (size_t)hc[C_NHUGE];
      h->stats.n_huge = hc[C_NHUGE];
      if (h->huge_cap_alloc < h->long_cap) { // scratch lists sized by the current wide_threshold
        int rc = dalloc(h, &S.huge_id, (size_t)HUGE_BATCH * (h->long_cap + 2)); if (rc) return rc;
        rc = dalloc(h, &S.huge_ll, (size_t)HUGE_BATCH * (h->long_cap + 2)); if (rc) return rc;
        h->huge_cap_alloc = h->long_cap;
      }
      for (size_t g = 0; g < n_huge; g += HUGE_BATCH) {
        const int nb = (int)std::min<size_t>(HUGE_BATCH, n_huge - g);
        int n_tiles = 0;
        KLAUNCH(h, k_huge_plan, 1, 1024, 0, S, S.huge_list + g, nb, S.huge_tiles, S.huge_cnt);
        CK(cudaMemcpyAsync(&n_tiles, S.huge_tiles + nb, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (n_tiles > 0) KLAUNCH(h, k_huge_scan, n_tiles, TPB, 0, S, S.huge_list + g, nb, S.huge_tiles, S.huge_cnt, h->long_cap);
        KLAUNCH(h, k_huge_finish, nb, TPB, smem, S, S.huge_list + g, S.huge_cnt, h->long_cap);
      }
      CK(cudaMemcpyAsync(hc, S.counters, sizeof hc, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      lap("grid-scanned huge buckets", (int)n_huge);
    }
  }
  // Records without a stored list: no usable bucket (their possible links still have to be counted)
  // or more than wide_threshold possible links in their bucket.  Their possible links among all the
  // items are listed in batches (k_wide_scan).  Scan order: those with more than
  // wide_threshold take their turn first, in ascending id, one at a time; the others keep their list
  // like everybody else and wait for the rounds.
  std::vector<int> unk; // no stored list and no count of their possible links yet
  if (hc[C_NUNK] > 0) {
    unk.resize(hc[C_NUNK]);
    CK(cudaMemcpyAsync(unk.data(), S.unk_list, sizeof(int) * unk.size(), cudaMemcpyDeviceToHost, st));
  }
  std::vector<int> known; // wide records found by the candidate scan of their own bucket
  if (hc[C_NWIDE] > 0) {
    known.resize(hc[C_NWIDE]);
    CK(cudaMemcpyAsync(known.data(), S.wide_list, sizeof(int) * known.size(), cudaMemcpyDeviceToHost, st));
  }
  if (!unk.empty() || !known.empty()) CK(cudaStreamSynchronize(st));
  std::sort(unk.begin(), unk.end());
  std::sort(known.begin(), known.end());
  CK(cudaEventRecord(h->ev[4], st));
  // Their possible links are listed from the blocking tables (k_wide_emit): the wide records from the
  // bucket they were found in, the others through the cheapest table built this sweep - instead of a
  // scan over all the items, which is left to the records no table can serve.
  std::vector<int> all_listed(known.size() + unk.size()); // both kinds, ascending
  std::merge(known.begin(), known.end(), unk.begin(), unk.end(), all_listed.begin());
  auto is_unk = [&](int r) { return std::binary_search(unk.begin(), unk.end(), r); };
  std::vector<int> wide; // the records that took their turn first
  int dk_wide = 0;
  h->stats.n_listed = 0; h->stats.n_scanned = 0;
  // The listed records are taken in ascending id, in chunks whose lists fit the scratch together (all of
  // them at once unless the chain has drifted into a state with tens of thousands of such records): a
  // chunk is listed, the records that keep their list are stored, the turns of the others are taken,
  // then the next chunk follows - the order of the serial turns is the ascending id as before.
  size_t chunk = std::max<size_t>(1, std::min<size_t>((size_t)h->emit_chunk, (size_t)WIDE_BUCKET_MAX));
  for (size_t c0 = 0; c0 < all_listed.size();) {
  const size_t c1 = h->opt_wide_buckets ? std::min(all_listed.size(), c0 + chunk) : all_listed.size();
  const std::vector<int> listed(all_listed.begin() + (long)c0, all_listed.begin() + (long)c1);
  int n_known = 0;
  for (int r : listed) n_known += !is_unk(r);
  std::vector<int> kb_base, kb_len, kb_served;
  bool from_buckets = false;
  if (h->opt_wide_buckets) {
    const auto t_a = std::chrono::steady_clock::now();
    int rc = wide_bucket_lists(h, listed, kb_base, kb_len, kb_served, &from_buckets); if (rc) return rc;
    if (!from_buckets && listed.size() > 1) { // the lists of this chunk do not fit the scratch together: a smaller chunk
      chunk = std::max<size_t>(1, listed.size() / 2);
      if (h->trace) std::fprintf(stderr, "[exb200] the lists of %d records do not fit the scratch: chunks of %d\n", (int)listed.size(), (int)chunk);
      continue;
    }
    if (h->trace) {
      long long sum = 0;
      int n_served = 0;
      for (size_t i = 0; i < kb_len.size(); i++) { sum += kb_len[i]; n_served += kb_served[i]; }
      std::fprintf(stderr, "[exb200] %d wide records and %d without a serving table listed from the tables: %.2f ms, %lld possible links, %d left to the scan%s\n",
                   n_known, (int)listed.size() - n_known, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_a).count(), sum,
                   from_buckets ? (int)listed.size() - n_served : (int)listed.size() - n_known, from_buckets ? "" : " (did not fit: batched scan)");
    }
  }
  std::vector<int> serial; // (record << 1) | no count of its possible links yet (listed by the batched scan, then classified)
  std::vector<char> listed_wide_unk(listed.size(), 0); // listed, and found to have more than wide_threshold possible links
  {
    std::vector<int> keep; // positions in `listed`: at most wide_threshold possible links - they keep their list and wait for the rounds
    for (size_t i = 0; i < listed.size(); i++) {
      const int r = listed[i];
      const bool u = is_unk(r);
      if (!from_buckets) serial.push_back((r << 1) | (u ? 1 : 0)); // everything by the batched scan
      else if (!u) serial.push_back(r << 1);
      else if (!kb_served[i]) serial.push_back((r << 1) | 1);
      else if (kb_len[i] > h->long_cap) { listed_wide_unk[i] = 1; serial.push_back(r << 1); }
      else keep.push_back((int)i);
    }
    for (size_t i = 0; i < listed.size(); i++) {
      const bool u = is_unk(listed[i]);
      if (from_buckets && u && kb_served[i]) h->stats.n_listed++;
      else if (!from_buckets || u) h->stats.n_scanned++;
    }
    if (!keep.empty()) {
      const int n = (int)listed.size();
      CK(cudaMemcpyAsync(h->we_seg + 3 * n, keep.data(), sizeof(int) * keep.size(), cudaMemcpyHostToDevice, st));
      KLAUNCH(h, k_wide_store_b, (int)keep.size(), TPB, 0, S, h->we_recs, h->we_seg + 3 * n, h->we_seg, h->we_seg + n);
      CK(cudaStreamSynchronize(st)); // `keep` is on the stack
    }
  }
  // The serial list is taken in ascending id, in stretches that hold at most WIDE_NB records whose
  // possible links have to be listed by the batched scan over all items; the turns of a stretch are
  // then taken in order, whichever way a record's list was made (lists are static, sizes are not).
  size_t g = 0;
  while (g < serial.size()) {
    int recs[WIDE_NB], tot[WIDE_NB], nb = 0;
    size_t g_end = g, entry_of_slot[WIDE_NB];
    std::vector<int> slot; // per entry of the stretch: its scan slot, or -1 - (position in `known`)
    for (; g_end < serial.size(); g_end++) {
      const int r = serial[g_end] >> 1;
      const bool scan = (serial[g_end] & 1) || !from_buckets;
      if (scan && nb == WIDE_NB) break;
      if (scan) { recs[nb] = r; entry_of_slot[nb] = g_end; slot.push_back(nb++); }
      else slot.push_back(-1 - (int)(std::lower_bound(listed.begin(), listed.end(), r) - listed.begin()));
    }
    if (nb > 0) {
      const auto t_a = std::chrono::steady_clock::now();
      int rc = wide_count(h, recs, nb, tot); if (rc) return rc;
      if (h->trace) {
        long long sum = 0;
        for (int q = 0; q < nb; q++) sum += tot[q];
        std::fprintf(stderr, "[exb200] serial batch of %d: count pass %.2f ms, %lld possible links\n", nb,
                     std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_a).count(), sum);
      }
    }
    size_t e = 0; // entries of the stretch handled so far
    int b = 0;    // scan slots whose lists have been filled
    while (e < slot.size()) {
      // the next group of scan slots whose lists fit the scratch together
      unsigned long long bases[WIDE_NB], used = 0;
      int sel[WIDE_NB], n_sel = 0;
      for (int q = 0; q < nb; q++) bases[q] = ~0ull;
      const int b0 = b;
      for (; b < nb; b++) {
        if (used + (unsigned long long)tot[b] > S.wl_cap) break;
        bases[b] = used; used += (unsigned long long)tot[b];
        if (tot[b] <= h->long_cap && (serial[entry_of_slot[b]] & 1)) sel[n_sel++] = b; // keeps its list
      }
      if (b == b0 && b < nb) return fail(h, EXB200_ERR_CUDA, "serial path: list larger than its scratch (internal error)");
      if (b > b0) {
        CK(cudaMemcpyAsync(S.wl_base, bases, sizeof(unsigned long long) * nb, cudaMemcpyHostToDevice, st));
        KLAUNCH(h, k_wide_fill_log, WIDE_BLOCKS, TPB, 0, S, nb);
        KLAUNCH(h, k_wide_scan<true>, WIDE_BLOCKS, TPB, 0, S, S.batch_list, nb); // pieces whose log overflowed
        if (n_sel > 0) {
          CK(cudaMemcpyAsync(S.wl_sel, sel, sizeof(int) * n_sel, cudaMemcpyHostToDevice, st));
          KLAUNCH(h, k_wide_store, n_sel, TPB, 0, S, S.batch_list, S.wl_sel, h->long_cap);
        }
      }
      // turns, in order, up to the first entry whose list is not filled yet
      std::vector<WideTurn> turns;
      for (; e < slot.size(); e++) {
        const int sl = slot[e], entry = serial[g + e], r = entry >> 1;
        if (sl >= b) break;
        if (sl < 0) { // listed from the tables
          const int k = -1 - sl;
          if (listed_wide_unk[k]) { // more than wide_threshold possible links, as it turned out
            static const int code_wide = -1;
            CK(cudaMemcpyAsync(S.cand_cnt + r, &code_wide, sizeof(int), cudaMemcpyHostToDevice, st));
          }
          turns.push_back(WideTurn{r, kb_len[k], 1, 0, (unsigned long long)kb_base[k]});
          wide.push_back(r);
          continue;
        }
        if ((entry & 1) && tot[sl] <= h->long_cap) continue; // keeps its list and waits for the rounds
        if (entry & 1) { // no bucket and more than wide_threshold possible links
          static const int code_wide = -1;
          CK(cudaMemcpyAsync(S.cand_cnt + r, &code_wide, sizeof(int), cudaMemcpyHostToDevice, st));
        }
        turns.push_back(WideTurn{r, tot[sl], 0, 0, bases[sl]});
        wide.push_back(r);
      }
      const auto t_t = std::chrono::steady_clock::now();
      { int rc = wide_turns(h, turns); if (rc) return rc; }
      CK(cudaStreamSynchronize(st)); // bases / sel are on the stack
      if (h->trace && !turns.empty())
        std::fprintf(stderr, "[exb200] %d serial turns: %.2f ms\n", (int)turns.size(),
                     std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_t).count());
    }
    g = g_end;
  }
  c0 = c1;
  } // chunks of the listed records
  if (!all_listed.empty()) {
    CK(cudaMemcpyAsync(hc, S.counters, sizeof hc, cudaMemcpyDeviceToHost, st)); // C_NACTL_INIT, C_NFENCE moved (k_wide_store, k_wide_store_b)
    CK(cudaStreamSynchronize(st));
  }
  const int n_fence = hc[C_NFENCE]; // candidate pool exhausted: these take their turn serially inside the rounds
  h->stats.n_long = hc[C_NACTL_INIT];
  h->stats.n_wide = (int)wide.size() + n_fence;
  if (!wide.empty()) {
    CK(cudaMemcpyAsync(&dk_wide, S.counters + C_DK_TOTAL, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(S.act_out, wide.data(), sizeof(int) * wide.size(), cudaMemcpyHostToDevice, st));
    KLAUNCH(h, k_clear_bounds, nblk((long long)wide.size()), TPB, 0, S, S.act_out, (int)wide.size());
    CK(cudaMemsetAsync(S.counters + C_DK_TOTAL, 0, sizeof(int), st));
    CK(cudaStreamSynchronize(st));
    S.K0 = h->K + dk_wide; // what the remaining records see as the starting number of entities
  }
  // records whose list did not fit the candidate pool (n_fence) are pending too
  bool round0_ran = false;
  {
    // round 0 needs prior_weight_new(K) positive and finite whatever K is when the record's turn comes
    // (clust_params.cpp:9-31: alpha + d K; kappa |m - K|, |m - K| - zero only at K = m)
    const exb200_clust& c = h->cl;
    bool sure = h->opt_round0 != 0;
    if (c.kind == EXB200_CLUST_PITMAN_YOR) sure = sure && c.alpha > 0.0 && c.d >= 0.0 && std::isfinite(c.alpha + c.d * (double)N);
    else sure = sure && c.m > (double)(N - 1) && std::isfinite(c.m) /* K without the record is at most N - 1 */ && (c.kind != EXB200_CLUST_GEN_COUPON || (c.kappa > 0.0 && std::isfinite(c.kappa * c.m)));
    if (S.use_kbounds || sure) KLAUNCH(h, k_bounds_init, nblk(N), TPB, 0, S, n_fence > 0 ? 1 : 0, sure ? 1 : 0);
    round0_ran = sure && n_fence == 0;
  }
  CK(cudaEventRecord(h->ev[11], st));

  // dependency rounds
  int n_act = N, n_actl = hc[C_NACTL_INIT];
  const int* act = nullptr; // round 1: every record (the short kernels skip the long ones)
  int rounds = 0;
  // Round 1 runs over all records as separate launches; the rounds after it in ONE persistent
  // cooperative kernel over the sorted list of the records still open (k_rounds_tail) - unless a
  // record without a stored list has to take a serial turn inside the rounds (pool exhausted).
  const bool use_tail = h->opt_tail && n_fence == 0;
  if (use_tail && !h->tail.n) {
    int rc = dalloc(h, &h->tail.list[0], (size_t)N); if (rc) return rc;
    rc = dalloc(h, &h->tail.list[1], (size_t)N); if (rc) return rc;
    rc = dalloc(h, &h->tail.flag, (size_t)N); if (rc) return rc;
    rc = dalloc(h, &h->tail.blk_tot, (size_t)3 * 4096); if (rc) return rc;
    cub::DeviceSelect::If(nullptr, h->select_tmp_bytes, thrust::counting_iterator<int>(0), h->tail.list[0], h->tail.n, N,
                          NotFinal{S.fin}, st);
    char* tmp = nullptr;
    rc = dalloc(h, &tmp, h->select_tmp_bytes); if (rc) return rc;
    h->d_select_tmp = tmp;
    rc = dalloc(h, &h->tail.n, 8); if (rc) return rc;
    rc = dalloc(h, &h->tail.long_pos, (size_t)N); if (rc) return rc;
    rc = dalloc(h, &h->tail.bar, 1); if (rc) return rc;
  }
  if (use_tail && h->trace && !h->tail.dbg) {
    int rc = dalloc(h, &h->tail.dbg, 257); if (rc) return rc;
  }
  // After round 0 one record in eight is left: no separate launches for round 1 either - the
  // persistent kernel starts from the compacted list of the open records (a thread per record over all
  // N would keep nearly every warp busy with its one or two open records).
  const bool skip_round1 = use_tail && round0_ran;
  while (n_act + n_actl > 0) {
    if (!skip_round1) rounds++;
    const uint32_t round = h->round_ctr++;
    if (h->round_ctr >= 0x7FFFFFF0u) { // reservation words would wrap: start over
      CK(cudaMemsetAsync(S.owner, 0xFF, sizeof(unsigned long long) * 2 * N, st));
      CK(cudaMemsetAsync(S.fence, 0xFF, sizeof(unsigned long long), st));
      h->round_ctr = 1;
    }
    const int init[3] = {0, 0x7FFFFFFF, -1}; // C_NACT_OUT, C_MIN_ACTIVE, C_WIDE_READY
    CK(cudaMemcpyAsync(S.counters, init, sizeof init, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(S.counters + C_NACTL_OUT, 0, sizeof(int), st));
    if (!skip_round1) {
      if (n_act > 0) KLAUNCH(h, k_reserve, nblk(n_act), TPB, 0, S, act, n_act, round);
      if (n_actl > 0) KLAUNCH(h, k_reserve_w, n_actl, TPB, 0, S, S.actl_in, n_actl, round);
      if (S.use_kbounds) {
        int rc = scan_exclusive<int8_t>(h, S.dlo, S.plo, N); if (rc) return rc;
        rc = scan_exclusive<int8_t>(h, S.dhi, S.phi_, N); if (rc) return rc;
      }
      if (n_act > 0) KLAUNCH_DP(h, k_commit, nblk(n_act), TPB, 0, S, act, n_act, round);
      if (n_actl > 0) KLAUNCH_DP(h, k_commit_w, n_actl, TPB, sizeof(double) * (size_t)std::max(h->long_cap, 1), S, S.actl_in, n_actl, round);
    }
    if (use_tail) {
      // everything that is left, in ascending id; K bounds from the outcome of round 1
      if (S.use_kbounds) {
        int rc = scan_exclusive<int8_t>(h, S.dlo, S.plo, N); if (rc) return rc;
        rc = scan_exclusive<int8_t>(h, S.dhi, S.phi_, N); if (rc) return rc;
      }
      CK(cudaMemsetAsync(h->tail.n, 0, sizeof(int) * 8, st));
      CK(cudaMemsetAsync(h->tail.bar, 0, sizeof(unsigned long long), st));
      CK(cub::DeviceSelect::If(h->d_select_tmp, h->select_tmp_bytes, thrust::counting_iterator<int>(0), h->tail.list[0],
                               h->tail.n, N, NotFinal{S.fin}, st));
      h->n_launches += 2;
      const size_t smem = sizeof(double) * (size_t)std::max(h->long_cap, 1);
      if (h->tail_grid_cap != h->long_cap) {
        int per_sm = 0;
        CK(S.has_dp ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_rounds_tail<true>, TAIL_TPB, smem)
                    : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_rounds_tail<false>, TAIL_TPB, smem));
        if (per_sm < 1) return fail(h, EXB200_ERR_CUDA, "k_rounds_tail does not fit a multiprocessor");
        h->tail_grid = std::min(4096, per_sm * h->n_sm);
        h->tail_grid_cap = h->long_cap;
      }
      uint32_t round0 = h->round_ctr;
      int max_rounds = (int)std::min<long long>(4LL * N + 16, 0x30000000LL);
      void* args[] = {(void*)&S, (void*)&h->tail, (void*)&round0, (void*)&max_rounds};
      CK(cudaEventRecord(h->ev[15], st));
      CK(cudaLaunchCooperativeKernel(S.has_dp ? (const void*)k_rounds_tail<true> : (const void*)k_rounds_tail<false>,
                                     dim3((unsigned)h->tail_grid), dim3(TAIL_TPB), args, smem, st));
      CK(cudaEventRecord(h->ev[16], st));
      h->tail_ran = true;
      h->n_launches++;
      int tn[4];
      CK(cudaMemcpyAsync(tn, h->tail.n, sizeof tn, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      if (h->trace) std::fprintf(stderr, "[exb200] round 1: %d records, then %d open; %d more rounds in the persistent kernel (%d blocks)\n",
                                 n_act, tn[0], tn[1], h->tail_grid);
      if (h->trace && h->tail.dbg) {
        unsigned long long dbg[257];
        CK(cudaMemcpy(dbg, h->tail.dbg, sizeof dbg, cudaMemcpyDeviceToHost));
        unsigned long long prev = dbg[256];
        for (int k = 0; k < std::min(tn[1], 64); k++) {
          std::fprintf(stderr, "[exb200]   tail round %2d: %8llu records  reserve %7.1f us  turns %7.1f us  scan %7.1f us\n", k + 2, dbg[4 * k],
                       (dbg[4 * k + 1] - prev) * 1e-3, (dbg[4 * k + 2] - dbg[4 * k + 1]) * 1e-3, (dbg[4 * k + 3] - dbg[4 * k + 2]) * 1e-3);
          prev = dbg[4 * k + 3];
        }
      }
      h->round_ctr += (uint32_t)tn[1];
      rounds += tn[1];
      if (tn[2] || tn[3] > 0) return fail(h, EXB200_ERR_CUDA, "links update made no progress (internal error)");
      break;
    }
    CK(cudaMemcpyAsync(hc, S.counters, sizeof hc, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (h->trace) std::fprintf(stderr, "[exb200] round %d: short %d long %d -> %d %d%s\n", rounds, n_act, n_actl,
                               hc[C_NACT_OUT], hc[C_NACTL_OUT], hc[C_WIDE_READY] >= 0 ? " (serial record)" : "");
    const int* next_act = S.act_out;
    int n_next = hc[C_NACT_OUT];
    const int n_nextl = hc[C_NACTL_OUT];
    std::swap(S.act_in, S.act_out);
    std::swap(S.actl_in, S.actl_out);
    if (hc[C_WIDE_READY] >= 0) {
      // serial path for a record without a stored candidate list
      const int w = hc[C_WIDE_READY];
      int tot = 0;
      int rc = wide_count(h, &w, 1, &tot); if (rc) return rc;
      const unsigned long long zero_base = 0;
      CK(cudaMemcpyAsync(S.wl_base, &zero_base, sizeof zero_base, cudaMemcpyHostToDevice, st));
      KLAUNCH(h, k_wide_fill_log, WIDE_BLOCKS, TPB, 0, S, 1);
      KLAUNCH(h, k_wide_scan<true>, WIDE_BLOCKS, TPB, 0, S, S.batch_list, 1); // pieces whose log overflowed
      wide_turn(h, w, 0, tot);
      const int zero = 0;
      CK(cudaMemcpyAsync(S.counters + C_NACT_OUT, &zero, sizeof zero, cudaMemcpyHostToDevice, st));
      KLAUNCH(h, k_compact_active, nblk(n_next), TPB, 0, S, next_act, n_next);
      CK(cudaMemcpyAsync(hc, S.counters, sizeof hc, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      next_act = S.act_out;
      n_next = hc[C_NACT_OUT];
      std::swap(S.act_in, S.act_out);
    } else if (n_next + n_nextl >= n_act + n_actl && act != nullptr) {
      return fail(h, EXB200_ERR_CUDA, "links update made no progress (internal error)");
    }
    act = next_act;
    n_act = n_next;
    n_actl = n_nextl;
    if (rounds > 4 * N + 16) return fail(h, EXB200_ERR_CUDA, "links update did not terminate");
  }
  CK(cudaEventRecord(h->ev[5], st));
  h->stats.n_rounds = rounds;
  // rename entities to their smallest member; slots [0,N) of the next arrays become current
  const int zero = 0;
  CK(cudaMemcpyAsync(S.counters + C_K, &zero, sizeof zero, cudaMemcpyHostToDevice, st));
  KLAUNCH(h, k_canon, nblk(N), TPB, 0, S);
  std::swap(S.ent_size, S.ent_size_nx);
  std::swap(S.ent_head, S.ent_head_nx);
  CK(cudaMemcpyAsync(hc, S.counters, sizeof hc, cudaMemcpyDeviceToHost, st));
  long long c64[66];
  CK(cudaMemcpyAsync(c64, S.counters64, sizeof c64, cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(h->ev[6], st));
  CK(cudaStreamSynchronize(st));
  { // a record without a valid link (EXB200_ERR_WEIGHTS) also upsets the counts below: report it first
    int rc = check_device_error(h); if (rc) return rc;
  }
  if (hc[C_K] != h->K + dk_wide + hc[C_DK_TOTAL])
    return fail(h, EXB200_ERR_CUDA, "entity count mismatch after the links update: " + std::to_string(hc[C_K]) +
                                        " vs " + std::to_string(h->K + dk_wide + hc[C_DK_TOTAL]));
  h->K = hc[C_K];
  h->stats.n_entities = h->K;
  h->stats.n_links_changed = hc[C_CHANGED];
  for (int k = 0; k < 32; k++) { c64[0] += c64[2 + 2 * k]; c64[1] += c64[3 + 2 * k]; } // k_cand_fast spreads its counts
  h->stats.n_candidates = c64[0];
  h->stats.n_bucket_items = c64[1];
  return EXB200_OK;
}

// debugging aid ("check" option): recompute sizes / member lists from the links on the host
int check_state(exb200_handle* h) {
  const int N = h->N;
  std::vector<int> lam(N), size(2 * (size_t)N), head(2 * (size_t)N), next(N);
  CK(cudaMemcpy(lam.data(), h->S.lambda, sizeof(int) * N, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(size.data(), h->S.ent_size, sizeof(int) * 2 * N, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(head.data(), h->S.ent_head, sizeof(int) * 2 * N, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(next.data(), h->S.rec_next, sizeof(int) * N, cudaMemcpyDeviceToHost));
  std::vector<int> cnt(N, 0);
  int K = 0;
  for (int r = 0; r < N; r++) {
    if (lam[r] < 0 || lam[r] >= N) return fail(h, EXB200_ERR_CUDA, "check: link out of range");
    cnt[lam[r]]++;
  }
  for (int e = 0; e < N; e++) {
    if (cnt[e] != size[e]) return fail(h, EXB200_ERR_CUDA, "check: size mismatch at entity " + std::to_string(e));
    if (!cnt[e]) continue;
    K++;
    int n = 0, prev = -1;
    for (int r = head[e]; r >= 0; r = next[r]) {
      if (lam[r] != e || r <= prev) return fail(h, EXB200_ERR_CUDA, "check: member list broken at entity " + std::to_string(e));
      prev = r;
      if (++n > cnt[e]) break;
    }
    if (n != cnt[e] || head[e] != e)
      return fail(h, EXB200_ERR_CUDA, "check: member list / label mismatch at entity " + std::to_string(e));
  }
  if (K != h->K) return fail(h, EXB200_ERR_CUDA, "check: K mismatch");
  return EXB200_OK;
}

// ---- one application of State::update (state.h:60-75)
// One sweep = State::update() (src/state.h:60-75), in the pieces within-chain sharding needs
// (DESIGN.md section 7): the links update runs on every GPU, the entity attributes and the
// distortion indicators on a slice of the entity slots / records.
int sweep_links(exb200_handle* h) {
  h->S.iter = (uint32_t)h->iteration;
  h->hs.begin_sweep((uint32_t)h->iteration);
  CK(cudaEventRecord(h->ev[0], h->stream));
  int rc = update_clust(h); if (rc) return rc;          // clust_params_->update(links_)
  refresh_clust_dev(h);
  return update_links(h);                               // links_.update(...)
}
int sweep_entities(exb200_handle* h, int lo, int hi) {
  DevState& S = h->S;
  CK(cudaMemsetAsync(h->d_gen_count, 0, sizeof(int) * 3, h->stream));
  const long long gen_cap = (long long)h->N * h->A;
  if (hi > lo) { // ents_.update_attributes(...)
    int* multi = S.act_in; // [N] (free outside the links update)
    CK(cudaEventRecord(h->ev[17], h->stream));
    KLAUNCH(h, k_entities, nblk(hi - lo, 64), 64, 0, S, h->d_gen_list, h->d_gen_count, gen_cap, lo, hi, multi);
    CK(cudaEventRecord(h->ev[18], h->stream));
    h->ent_single_ran = true;
    KLAUNCH(h, k_entities_slow, h->n_sm * 8, TPB, 0, S, h->d_gen_list, h->d_gen_count, gen_cap);
    KLAUNCH(h, k_entities_multi, h->n_sm * 16, 128, 0, S, (const int*)multi, h->d_gen_list, h->d_gen_count);
  }
  KLAUNCH_DP(h, k_entities_general, h->n_sm * 8, TPB, 0, S, h->d_gen_list, h->d_gen_count, gen_cap);
  return EXB200_OK;
}
int sweep_distort(exb200_handle* h, int lo, int hi) {
  DevState& S = h->S;
  cudaStream_t st = h->stream;
  const int FA = h->F * h->A;
  CK(cudaEventRecord(h->ev[7], st));
  CK(cudaMemsetAsync(S.ndist, 0, sizeof(int) * FA, st));
  CK(cudaMemsetAsync(S.corr, 0, sizeof(int) * FA, st));
  const int use_smem = 2 * FA * (int)sizeof(int) <= 32768;
  if (hi > lo) KLAUNCH(h, k_distort, nblk(hi - lo), TPB, use_smem ? 2 * FA * sizeof(int) : 0, S, use_smem, lo, hi); // recs_.update_distortion
  CK(cudaMemcpyAsync(h->ndist.data(), S.ndist, sizeof(int) * FA, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(h->corr.data(), S.corr, sizeof(int) * FA, cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(h->ev[8], st));
  return check_device_error(h);                         // also synchronises
}
int sweep_finish(exb200_handle* h);

int sweep(exb200_handle* h) {
  int rc = sweep_links(h); if (rc) return rc;
  rc = sweep_entities(h, 0, h->N); if (rc) return rc;
  rc = update_dists_and_concs(h); if (rc) return rc;    // ents_.update_distributions(); cache_.update(); distort_dist_concs_.update()
  rc = sweep_distort(h, 0, h->N); if (rc) return rc;
  return sweep_finish(h);
}

// distort_probs_.update from the (global) counts in h->ndist / h->corr, end of the sweep
int sweep_finish(exb200_handle* h) {
  cudaStream_t st = h->stream;
  int rc = 0;
  // distort_probs_.update (distortion_probs.cpp:30-47): file-major, attribute-minor order
  for (int f = 0; f < h->F; f++)
    for (int a = 0; a < h->A; a++)
      if (h->attrs[a].s1 > 0)
        h->theta[f * h->A + a] = h->hs.beta((double)h->ndist[f * h->A + a] + h->attrs[a].s1,
                                            (double)h->corr[f * h->A + a] + h->attrs[a].s2);
  rc = upload_theta(h); if (rc) return rc;
  h->iteration++;
  CK(cudaEventRecord(h->ev[9], st));
  CK(cudaStreamSynchronize(st));
  auto ms = [&](int a, int b) { float t = 0; cudaEventElapsedTime(&t, h->ev[a], h->ev[b]); return t; };
  h->stats.ms_clust = ms(0, 1); h->stats.ms_prep = ms(1, 2); h->stats.ms_index = ms(2, 3);
  h->stats.ms_index_count = h->has_index_events ? ms(12, 14) : 0.f; // CSR tables
  h->stats.ms_index_fill = h->has_index_events ? ms(14, 3) : 0.f;   // chain tables
  h->stats.ms_cand = ms(3, 10); h->stats.ms_cand_long = ms(10, 4); h->stats.ms_wide = ms(4, 11); h->stats.ms_rounds = ms(11, 5); h->stats.ms_canon = ms(5, 6);
  h->stats.ms_cand_fast = h->fast_ran ? ms(3, 13) : 0.f;
  h->stats.ms_tail = h->tail_ran ? ms(15, 16) : 0.f;
  h->stats.ms_entities_single = h->ent_single_ran ? ms(17, 18) : 0.f;
  h->stats.ms_entities = ms(6, 7); h->stats.ms_distort = ms(7, 8); h->stats.ms_total = ms(0, 9);
  if (h->opt_check) { rc = check_state(h); if (rc) return rc; }
  return EXB200_OK;
}

} // namespace

// =============================================================================================
extern "C" {

int exb200_abi_version(void) { return EXB200_ABI_VERSION; }

int exb200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

void exb200_trim(void) { g_pool.trim(); }

const char* exb200_last_error(const exb200_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

void exb200_destroy(exb200_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
  for (auto& sl : h->slabs) if (!g_pool.give_dev(h->device, sl.base, sl.size)) cudaFree(sl.base); // (kept for the next handle)
  for (int b = 0; b < 2; b++) {
    if (h->pin_snap[b]) g_pool.give_pin(h->pin_snap[b], sizeof(int) * (size_t)h->N);
    if (h->ev_snap[b]) cudaEventDestroy(h->ev_snap[b]);
    if (h->ev_d2h[b]) cudaEventDestroy(h->ev_d2h[b]);
  }
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  for (auto& e : h->ev) if (e) cudaEventDestroy(e);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int exb200_create(const exb200_model* m, int device, exb200_handle** out) {
  if (!m || !out) return fail(nullptr, EXB200_ERR_INVALID, "null argument");
  *out = nullptr;
  const int N = m->n_records, A = m->n_attrs, F = m->n_files;
  if (N < 2 || A < 1 || A > MAX_ATTRS || F < 1 || m->n_entities < 1)
    return fail(nullptr, EXB200_ERR_INVALID, "need n_records >= 2, 1 <= n_attrs <= 32, n_files >= 1");
  if ((long long)N * 2 >= (1LL << 31)) return fail(nullptr, EXB200_ERR_INVALID, "n_records too large");
  int ndev = exb200_device_count();
  if (ndev <= 0) return fail(nullptr, EXB200_ERR_NO_DEVICE, "no CUDA device is visible; this library has no CPU path");
  if (device < 0 || device >= ndev) return fail(nullptr, EXB200_ERR_NO_DEVICE, "invalid device index");
  if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, EXB200_ERR_NO_DEVICE, "cudaSetDevice failed");

  exb200_handle* h = new exb200_handle();
  // the attribute tables are built by host threads while the records are loaded on the device
  std::vector<std::thread> attr_threads;
  std::vector<int> attr_rc(A, 0);
  auto join_attr = [&] { for (auto& t : attr_threads) if (t.joinable()) t.join(); };
  auto bail = [&](int rc) { join_attr(); g_create_error = h->err; exb200_destroy(h); return rc; };
  auto t_stage = std::chrono::steady_clock::now();
  auto stage = [&](const char* what) { // EXB200_TRACE: where the time of create goes
    if (!h->trace) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[exb200] create: %-18s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t_stage).count());
    t_stage = now;
  };
  h->device = device; h->N = N; h->A = A; h->F = F;
  h->RS = ((A + 1 + 7) / 8) * 8; h->ES = ((A + 7) / 8) * 8;
  h->iteration = m->iteration; h->seed = m->seed; h->cl = m->clust; h->hs.seed = m->seed;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess)
    return bail(fail(h, EXB200_ERR_NO_DEVICE, "cudaStreamCreate failed"));
  if (cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || h->n_sm < 1) h->n_sm = 148;
  for (auto& e : h->ev) cudaEventCreate(&e);
  const exb200_clust& c = m->clust;
  if (c.kind < 0 || c.kind > 2) return bail(fail(h, EXB200_ERR_INVALID, "Unrecognized clust_prior"));
  if (c.kind != EXB200_CLUST_PITMAN_YOR && c.m < (double)N)
    return bail(fail(h, EXB200_ERR_INVALID, "Coupon collecting parameter m < n_records is currently not supported."));

  stage("context/stream");
  // ---- device state.  The caller's arrays are uploaded in their R layout and turned into the
  // resident layout on the device (setup.cuh): entity label = smallest member record, member lists
  // in ascending record id.
  DevState& S = h->S;
  S.N = N; S.A = A; S.F = F; S.RS = h->RS; S.ES = h->ES; S.seed = h->seed;
  int rc = 0;
#define TRY(x) do { rc = (x); if (rc) return bail(rc); } while (0)
  h->attrs.resize(A); h->h_attrs.resize(A);
  for (int a = 0; a < MAX_ATTRS; a++) h->conc[a] = std::numeric_limits<double>::infinity();
  // one host thread per attribute (each spreads its neighbourhood work over more threads)
  for (int a = 0; a < A; a++)
    attr_threads.emplace_back([&, a] { cudaSetDevice(device); attr_rc[a] = build_attr(h, m->attrs[a], a); });
  {
    // labels: small non-negative integers (what exchanger() and this library produce) index a
    // direct table; anything else is first renamed to the column of ent_attrs
    const int nE = m->n_entities;
    int max_label = -1;
    bool dense = true;
    for (int e = 0; e < nE && dense; e++) {
      if (m->ent_ids[e] < 0 || m->ent_ids[e] > 8LL * N + 1024) dense = false;
      else max_label = std::max(max_label, m->ent_ids[e]);
    }
    std::vector<int> links_slot, ids_slot;
    const int* links_src = m->links;
    const int* ids_src = m->ent_ids;
    if (!dense) {
      std::unordered_map<int, int> slot_m;
      slot_m.reserve((size_t)nE * 2);
      for (int e = 0; e < nE; e++) slot_m[m->ent_ids[e]] = e;
      links_slot.resize(N);
      for (int r = 0; r < N; r++) {
        auto it = slot_m.find(m->links[r]);
        if (it == slot_m.end()) return bail(fail(h, EXB200_ERR_INVALID, "links refers to an unknown entity"));
        links_slot[r] = it->second;
      }
      ids_slot.resize(nE);
      for (int e = 0; e < nE; e++) ids_slot[e] = slot_m[m->ent_ids[e]]; // a repeated label keeps its last column
      links_src = links_slot.data();
      ids_src = ids_slot.data();
      max_label = nE - 1;
    }
    SetupDomains dom{};
    for (int a = 0; a < A; a++) dom.D[a] = m->attrs[a].domain_size;
    int *raw_x = nullptr, *raw_z = nullptr, *raw_f = nullptr, *raw_l = nullptr, *raw_ids = nullptr, *raw_y = nullptr;
    int *lab2slot = nullptr, *rep = nullptr, *d_misc = nullptr;
    unsigned long long *keys = nullptr, *keys2 = nullptr;
    void* cub_tmp = nullptr;
    std::vector<void*> tmp;
    auto talloc = [&](auto** p, size_t n) {
      void* q = nullptr;
      if (cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(**p)) != cudaSuccess) return false;
      tmp.push_back(q);
      *p = (std::remove_reference_t<decltype(**p)>*)q;
      return true;
    };
    auto release = [&] { for (void* q : tmp) cudaFree(q); tmp.clear(); };
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, cub_bytes, keys, keys2, N, 0, 64, h->stream);
    const bool ok = talloc(&raw_x, (size_t)N * A) && talloc(&raw_z, (size_t)N * A) && talloc(&raw_f, (size_t)N) &&
                    talloc(&raw_l, (size_t)N) && talloc(&raw_ids, (size_t)nE) && talloc(&raw_y, (size_t)nE * A) &&
                    talloc(&lab2slot, (size_t)max_label + 1) && talloc(&rep, (size_t)max_label + 1) &&
                    talloc(&d_misc, 2) && talloc(&keys, (size_t)N) && talloc(&keys2, (size_t)N) &&
                    cudaMalloc(&cub_tmp, std::max<size_t>(cub_bytes, 1)) == cudaSuccess;
    if (cub_tmp) tmp.push_back(cub_tmp);
    if (!ok) { release(); return bail(fail(h, EXB200_ERR_OOM, "cudaMalloc failed while loading the model")); }
    int* ip; uint32_t* up;
    auto tbail = [&](int code) { release(); return bail(code); };
#define TRYT(x) do { rc = (x); if (rc) return tbail(rc); } while (0)
    TRYT(dalloc(h, &ip, (size_t)N * h->RS)); S.rec_x = ip;
    TRYT(dalloc(h, &up, (size_t)N)); S.rec_z = up;
    TRYT(dalloc(h, &S.lambda, (size_t)N));
    TRYT(dalloc(h, &S.ent_y, 2 * (size_t)N * h->ES));
    TRYT(dalloc(h, &S.ent_size, 2 * (size_t)N));
    TRYT(dalloc(h, &S.ent_head, 2 * (size_t)N));
    TRYT(dalloc(h, &S.ent_size_nx, 2 * (size_t)N));
    TRYT(dalloc(h, &S.ent_head_nx, 2 * (size_t)N));
    TRYT(dalloc(h, &S.rec_next, (size_t)N));
    TRYT(dalloc(h, &S.ndist, (size_t)F * A));
    TRYT(dalloc(h, &S.corr, (size_t)F * A));
#undef TRYT
    cudaStream_t st = h->stream;
    cudaError_t ce = cudaSuccess;
    auto H2D = [&](void* d, const void* s, size_t bytes) { if (ce == cudaSuccess && bytes) ce = staged_copy(d, s, bytes, true, device, st); };
    auto SET = [&](void* d, int v, size_t bytes) { if (ce == cudaSuccess && bytes) ce = cudaMemsetAsync(d, v, bytes, st); };
    H2D(raw_x, m->rec_attrs, sizeof(int) * (size_t)N * A);
    H2D(raw_z, m->rec_distortions, sizeof(int) * (size_t)N * A);
    H2D(raw_f, m->file_ids, sizeof(int) * (size_t)N);
    H2D(raw_l, links_src, sizeof(int) * (size_t)N);
    H2D(raw_ids, ids_src, sizeof(int) * (size_t)nE);
    H2D(raw_y, m->ent_attrs, sizeof(int) * (size_t)nE * A);
    SET(lab2slot, 0xFF, sizeof(int) * ((size_t)max_label + 1));
    SET(rep, 0x7F, sizeof(int) * ((size_t)max_label + 1));
    SET(d_misc, 0, sizeof(int) * 2);
    SET(S.ndist, 0, sizeof(int) * (size_t)F * A);
    SET(S.ent_y, 0, sizeof(int) * 2 * (size_t)N * h->ES);
    SET(S.ent_size, 0, sizeof(int) * 2 * (size_t)N);
    SET(S.ent_head, 0xFF, sizeof(int) * 2 * (size_t)N);
    if (ce == cudaSuccess) {
      KLAUNCH(h, k_setup_records, nblk(N), TPB, 0, S, dom, raw_x, raw_z, raw_f, d_misc);
      KLAUNCH(h, k_setup_slots, nblk(nE), TPB, 0, raw_ids, nE, lab2slot);
      KLAUNCH(h, k_setup_rep, nblk(N), TPB, 0, N, raw_l, lab2slot, max_label, rep, d_misc);
    }
    int misc[2] = {0, 0};
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(misc, d_misc, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    if (ce == cudaSuccess && misc[0] == 0) { // every link is valid: entity tables and member lists
      KLAUNCH(h, k_setup_links, nblk(N), TPB, 0, S, dom, raw_l, lab2slot, rep, raw_y, keys, d_misc + 1, d_misc);
      ce = cub::DeviceRadixSort::SortKeys(cub_tmp, cub_bytes, keys, keys2, N, 0, 64, st);
      if (ce == cudaSuccess) KLAUNCH(h, k_setup_next, nblk(N), TPB, 0, N, keys2, S.rec_next);
      if (ce == cudaSuccess) ce = cudaMemcpyAsync(S.ent_size_nx, S.ent_size, sizeof(int) * 2 * (size_t)N, cudaMemcpyDeviceToDevice, st);
      if (ce == cudaSuccess) ce = cudaMemcpyAsync(S.ent_head_nx, S.ent_head, sizeof(int) * 2 * (size_t)N, cudaMemcpyDeviceToDevice, st);
      h->ndist.assign((size_t)F * A, 0); h->corr.assign((size_t)F * A, 0);
      if (ce == cudaSuccess) ce = cudaMemcpyAsync(h->ndist.data(), S.ndist, sizeof(int) * (size_t)F * A, cudaMemcpyDeviceToHost, st);
      if (ce == cudaSuccess) ce = cudaMemcpyAsync(misc, d_misc, sizeof(int) * 2, cudaMemcpyDeviceToHost, st);
      if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    }
    release();
    if (ce != cudaSuccess) return bail(fail(h, EXB200_ERR_CUDA, std::string("loading the model: ") + cudaGetErrorString(ce)));
    static const char* const what[] = {"", "rec_attrs out of range", "file_ids out of range",
                                       "links refers to an unknown entity", "ent_attrs out of range"};
    if (misc[0]) return bail(fail(h, EXB200_ERR_INVALID, what[std::min(misc[0], 4)]));
    h->K = misc[1];
  }
  h->theta.resize((size_t)F * A);
  for (int f = 0; f < F; f++)
    for (int a = 0; a < A; a++) h->theta[f * A + a] = m->distort_probs[f + (size_t)F * a];
  stage("state on device");
  TRY(dalloc(h, &h->d_theta, (size_t)F * A)); S.theta = h->d_theta;
  int maxD = 1;
  join_attr();
  for (int a = 0; a < A; a++) {
    if (attr_rc[a]) return bail(attr_rc[a]);
    maxD = std::max(maxD, m->attrs[a].domain_size);
  }
  stage("attribute tables");
  TRY(dalloc(h, &h->d_hist, (size_t)maxD));
  refresh_conc_dev(h);
  // the reference draws phi in the Entities constructor (entities.cpp:233): same here, from the
  // entity values of the initial state, with a counter range of its own
  h->hs.iteration = (uint32_t)m->iteration; h->hs.counter = 0x40000000u;
  for (int a = 0; a < A; a++)
    if (h->attrs[a].dirichlet) {
      std::vector<int> counts(h->attrs[a].D, 0);
      cudaMemsetAsync(h->d_hist, 0, sizeof(int) * counts.size(), h->stream);
      KLAUNCH(h, k_ent_hist, nblk(N), TPB, 0, S, a, h->d_hist);
      cudaMemcpyAsync(counts.data(), h->d_hist, sizeof(int) * counts.size(), cudaMemcpyDeviceToHost, h->stream);
      if (cudaStreamSynchronize(h->stream) != cudaSuccess) return bail(fail(h, EXB200_ERR_CUDA, "entity-value histogram failed"));
      TRY(draw_entity_dist(h, a, counts));
    }
  TRY(upload(h, &h->d_attrs, h->h_attrs.data(), (size_t)A)); S.attrs = h->d_attrs;
  TRY(dalloc(h, &S.Lnew, (size_t)N));
  TRY(dalloc(h, &S.rec_tab, (size_t)N));
  TRY(dalloc(h, &S.rec_bkt, (size_t)N));
  TRY(dalloc(h, &S.tab_usage, (size_t)2 * N_TABLES));
  TRY(dalloc(h, &S.tab_est, (size_t)N_TABLES));
  TRY(dalloc(h, &S.rec_alt, (size_t)N));
  TRY(dalloc(h, &S.twin, (size_t)N));
  TRY(dalloc(h, &S.live0, (size_t)N));
  TRY(dalloc(h, &S.cand_off, (size_t)N));
  TRY(dalloc(h, &S.cand_cnt, (size_t)N));
  S.pool_cap = std::min<unsigned long long>((8ull + CAND_INLINE) * N + (1ull << 20), 0xFFFFFFF0ull);
  TRY(dalloc(h, &S.cand_id, (size_t)S.pool_cap));
  TRY(dalloc(h, &S.cand_ll, (size_t)S.pool_cap));
  TRY(dalloc(h, &S.pool_cursor, 1));
  TRY(dalloc(h, &S.owner, (size_t)2 * N));
  TRY(dalloc(h, &S.fence, 1));
  TRY(dalloc(h, &S.fin, (size_t)N));
  TRY(dalloc(h, &S.rdy_m1, (size_t)N));
  TRY(dalloc(h, &S.rdy_srel, (size_t)N));
  TRY(dalloc(h, &S.referenced, (size_t)2 * N));
  TRY(dalloc(h, &S.dlo, (size_t)N));
  TRY(dalloc(h, &S.dhi, (size_t)N));
  TRY(dalloc(h, &S.plo, (size_t)N));
  TRY(dalloc(h, &S.phi_, (size_t)N));
  TRY(dalloc(h, &S.act_in, (size_t)N));
  TRY(dalloc(h, &S.act_out, (size_t)N));
  TRY(dalloc(h, &S.actl_in, (size_t)N));
  TRY(dalloc(h, &S.actl_out, (size_t)N));
  TRY(dalloc(h, &S.long_todo, (size_t)N));
  TRY(dalloc(h, &S.probe_list, (size_t)N));
  TRY(dalloc(h, &S.wide_list, (size_t)N));
  TRY(dalloc(h, &S.unk_list, (size_t)N));
  TRY(dalloc(h, &S.wide_part, (size_t)3 * WIDE_PIECES));
  TRY(dalloc(h, &S.counters, (size_t)C_NSLOTS));
  TRY(dalloc(h, &S.counters64, 66));
  TRY(dalloc(h, &S.slow_list, (size_t)N));
  TRY(dalloc(h, &S.tab_over, (size_t)N_TABLES));
  TRY(dalloc(h, &h->d_chain_tabs, (size_t)N_TABLES));
  TRY(dalloc(h, &S.dev_err, 2));
  TRY(dalloc(h, &S.lwbuf, (size_t)16 * N));
  S.wl_cap = 10ull * N; // 10N doubles + 10N ints inside the 16N doubles of lwbuf
  S.wl_ll = S.lwbuf; S.wl_id = reinterpret_cast<int*>(S.lwbuf + S.wl_cap);
  TRY(dalloc(h, &S.wl_cnt, (size_t)WIDE_NB * WIDE_PIECES));
  TRY(dalloc(h, &S.wl_tot, (size_t)WIDE_NB));
  TRY(dalloc(h, &S.wl_base, (size_t)WIDE_NB));
  TRY(dalloc(h, &S.wl_sel, (size_t)WIDE_NB));
  TRY(dalloc(h, &S.huge_list, (size_t)N));
  TRY(dalloc(h, &S.huge_tiles, (size_t)HUGE_BATCH + 1));
  TRY(dalloc(h, &S.huge_cnt, (size_t)HUGE_BATCH));
  TRY(dalloc(h, &S.batch_list, (size_t)WIDE_NB));
  TRY(dalloc(h, &h->d_active, (size_t)N_TABLES));
  TRY(dalloc(h, &h->d_remap, (size_t)2 * N_TABLES)); S.tab_remap = h->d_remap;
  TRY(dalloc(h, &h->d_part_tabs, (size_t)N_TABLES));
  TRY(dalloc(h, &h->pb.part_cnt, (size_t)PA_MAX_TABS * MAX_PARTS));
  TRY(dalloc(h, &h->pb.part_start, (size_t)PA_MAX_TABS * (MAX_PARTS + 1)));
  TRY(dalloc(h, &h->d_cl_counts, 2));
  TRY(dalloc(h, &h->d_gen_list, (size_t)N * (size_t)A)); // worst case (include mode): every entity, every attribute
  TRY(dalloc(h, &h->d_gen_count, 3));
  cudaMemsetAsync(S.owner, 0xFF, sizeof(unsigned long long) * 2 * N, h->stream);
  cudaMemsetAsync(S.fence, 0xFF, sizeof(unsigned long long), h->stream);
  cudaMemsetAsync(S.dev_err, 0, sizeof(int) * 2, h->stream);
  // blocking-table descriptors (storage is allocated when a table is first used).  Key attributes:
  // up to KEY_ATTRS attributes with the largest domains; a table is identified by its key mask.
  {
    std::vector<int> order(A);
    for (int a = 0; a < A; a++) order[a] = a;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return m->attrs[x].domain_size > m->attrs[y].domain_size; });
    S.n_key = std::min(A, KEY_ATTRS);
    for (int a = 0; a < MAX_ATTRS; a++) S.key_pos[a] = -1;
    std::vector<int> keys(order.begin(), order.begin() + S.n_key);
    std::sort(keys.begin(), keys.end());
    for (int k = 0; k < S.n_key; k++) { S.key_attr[k] = keys[k]; S.key_pos[keys[k]] = (int8_t)k; }
    h->n_tab = 1 << S.n_key;
  }
  { // packed attribute tuples: one field per attribute, if they fit 64 bits (and a record is one 32-byte row)
    int shift = 0;
    S.pk_ok = h->RS == 8 && h->ES == 8;
    for (int a = 0; a < A; a++) {
      int bits = 1;
      while ((1ll << bits) < (long long)m->attrs[a].domain_size) bits++;
      if (shift + bits > 64) { S.pk_ok = 0; break; }
      S.pk_shift[a] = (uint8_t)shift; S.pk_bits[a] = (uint8_t)bits;
      shift += bits;
    }
    S.hop_cap = h->hop_cap;
  }
  h->tables.resize((size_t)h->n_tab);
  uint32_t Bhash = 1024; // about one bucket per indexed item (twins are not indexed)
  while ((double)Bhash < 1.2 * (double)N) Bhash <<= 1;
  for (int t = 1; t < h->n_tab; t++) {
    DevTable& d = h->tables[t].d;
    d.mask = (uint32_t)t;
    unsigned long long keys = 1;
    for (int k = 0; k < S.n_key; k++)
      if (t >> k & 1) keys = std::min<unsigned long long>(keys * (unsigned long long)m->attrs[S.key_attr[k]].domain_size, 1ull << 40);
    h->tables[t].keys = keys;
    d.exact = keys <= (unsigned long long)Bhash;
    d.B = d.exact ? (uint32_t)keys : Bhash;
    d.ptr = nullptr; d.ids = nullptr; d.chain = 0; d.csr = 1; d.kmask = 0; d.cells = 0; d.head = nullptr; d.node = nullptr;
  }
  {
    std::vector<DevTable> td((size_t)h->n_tab);
    for (int t = 0; t < h->n_tab; t++) td[t] = h->tables[t].d;
    TRY(upload(h, &h->d_tables, td.data(), td.size())); S.tables = h->d_tables;
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) return bail(fail(h, EXB200_ERR_CUDA, "upload failed"));
  }
  if (cudaFuncSetAttribute(k_rounds_tail<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024) != cudaSuccess ||
      cudaFuncSetAttribute(k_rounds_tail<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024) != cudaSuccess ||
      cudaFuncSetAttribute(k_commit_w<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024) != cudaSuccess ||
      cudaFuncSetAttribute(k_commit_w<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024) != cudaSuccess ||
      cudaFuncSetAttribute(k_index_small, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
      cudaFuncSetAttribute(k_part_a, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
      cudaFuncSetAttribute(k_part_b, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024) != cudaSuccess ||
      cudaFuncSetAttribute(k_huge_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
      cudaFuncSetAttribute(k_cand_long, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
    return bail(fail(h, EXB200_ERR_CUDA, "cannot reserve shared memory for k_cand_long"));
#undef TRY
  S.bucket_cap = h->bucket_cap; S.cand_cap = h->cand_cap; S.long_cap = h->long_cap; S.long_min = h->long_min;
  S.est_threshold = h->est_threshold; S.huge_min = h->huge_min;
  refresh_clust_dev(h);
  rc = upload_theta(h);
  if (rc) return bail(rc);
  if (cudaStreamSynchronize(h->stream) != cudaSuccess)
    return bail(fail(h, EXB200_ERR_CUDA, std::string("device initialisation failed: ") + cudaGetErrorString(cudaGetLastError())));
  stage("scratch/finish");
  *out = h;
  return EXB200_OK;
}

int exb200_set_option(exb200_handle* h, const char* name, int64_t value) {
  if (!h || !name) return EXB200_ERR_INVALID;
  const std::string n(name);
  if (n == "check") h->opt_check = value != 0;
  else if (n == "trace") h->trace = value != 0;
  else if (n == "bucket_cap") h->bucket_cap = (int)std::max<int64_t>(0, value);
  else if (n == "cand_cap") h->cand_cap = (int)std::min<int64_t>(CAND_SMEM, std::max<int64_t>(0, value));
  else if (n == "wide_threshold") h->long_cap = (int)std::min<int64_t>(16384, std::max<int64_t>(0, value));
  else if (n == "long_min") h->long_min = (int)std::max<int64_t>(0, value);
  else if (n == "est_threshold") h->est_threshold = (double)value;
  else if (n == "huge_min") h->huge_min = (int)std::max<int64_t>(0, value);
  else if (n == "index_partitioned") h->index_partitioned = value != 0;
  else if (n == "rounds_tail") h->opt_tail = value != 0;
  else if (n == "fast_path") h->opt_fast = value != 0;
  else if (n == "wide_buckets") h->opt_wide_buckets = value != 0;
  else if (n == "round0") h->opt_round0 = value != 0;
  else if (n == "emit_cost_cap") h->emit_cost_cap = (double)value;
  else if (n == "emit_chunk") h->emit_chunk = (int)std::max<int64_t>(1, value);
  else if (n == "emit_scratch") h->emit_scratch = (long long)std::max<int64_t>(0, value);
  else if (n == "tiny_tables_always") h->opt_tiny_always = value != 0;
  else if (n == "hop_cap") h->hop_cap = (int)std::max<int64_t>(0, value);
  else if (n == "table_min_usage") h->table_min_usage = (int)std::max<int64_t>(0, value);
  else if (n == "table_cost_permille") { h->table_full_cost = (double)value * 1e-3; h->table_pair_cost = 0.25 * (double)value * 1e-3; }
  else return fail(h, EXB200_ERR_INVALID, "unknown option " + n);
  h->S.bucket_cap = h->bucket_cap; h->S.cand_cap = h->cand_cap; h->S.long_cap = h->long_cap;
  h->S.long_min = h->long_min; h->S.est_threshold = h->est_threshold; h->S.huge_min = h->huge_min;
  h->S.hop_cap = h->hop_cap;
  return EXB200_OK;
}

int exb200_sweeps(exb200_handle* h, int32_t n_sweeps) {
  if (!h) return EXB200_ERR_INVALID;
  if (h->poisoned) return fail(h, EXB200_ERR_CUDA, "handle unusable after an earlier error: " + h->err);
  cudaSetDevice(h->device);
  for (int i = 0; i < n_sweeps; i++) {
    int rc = sweep(h);
    if (rc) return rc;
  }
  return EXB200_OK;
}

int exb200_sweeps_timed(exb200_handle* h, int32_t n_sweeps, float* device_ms, int64_t* launches) {
  if (!h) return EXB200_ERR_INVALID;
  if (h->poisoned) return fail(h, EXB200_ERR_CUDA, "handle unusable after an earlier error: " + h->err);
  cudaSetDevice(h->device);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const long long l0 = h->n_launches;
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaEventRecord(e0, h->stream));
  int rc = EXB200_OK;
  for (int i = 0; i < n_sweeps && !rc; i++) rc = sweep(h);
  if (!rc) {
    CK(cudaEventRecord(e1, h->stream));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (device_ms) *device_ms = ms;
    if (launches) *launches = h->n_launches - l0;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return rc;
}

int pe_add_chain_links(struct exb200_pe* pe, exb200_handle* h); // point estimates below

namespace {
// Host side of the sample pipeline of exb200_run: waits for a pinned buffer to be filled and copies
// it into the caller's history row while the device runs the next sweep.
struct RowCopier {
  exb200_handle* h;
  std::mutex mu;
  std::condition_variable cv;
  std::deque<std::pair<int, int*>> jobs; // (buffer, destination row)
  bool busy[2] = {false, false};
  bool stop = false, failed = false;
  std::thread th;
  explicit RowCopier(exb200_handle* hh) : h(hh) {
    th = std::thread([this] {
      cudaSetDevice(h->device);
      for (;;) {
        std::pair<int, int*> job;
        {
          std::unique_lock<std::mutex> lk(mu);
          cv.wait(lk, [&] { return stop || !jobs.empty(); });
          if (jobs.empty()) return;
          job = jobs.front();
          jobs.pop_front();
        }
        // (a null destination only releases the buffer: the enqueue of this row failed)
        const bool ok = !job.second || cudaEventSynchronize(h->ev_d2h[job.first]) == cudaSuccess;
        if (ok && job.second) { // the destination is usually untouched memory: several threads take the page faults
          const size_t n = (size_t)h->N;
          const int parts = n >= (1u << 20) ? 4 : 1;
          std::vector<std::thread> th;
          for (int q = 1; q < parts; q++)
            th.emplace_back([=] { std::memcpy(job.second + n * q / parts, h->pin_snap[job.first] + n * q / parts, sizeof(int) * (n * (q + 1) / parts - n * q / parts)); });
          std::memcpy(job.second, h->pin_snap[job.first], sizeof(int) * (n / parts));
          for (auto& t : th) t.join();
        }
        {
          std::lock_guard<std::mutex> lk(mu);
          busy[job.first] = false;
          failed |= !ok;
        }
        cv.notify_all();
      }
    });
  }
  void acquire(int b) { // blocks until the previous use of buffer b has reached the caller's array
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return !busy[b]; });
    busy[b] = true;
  }
  void submit(int b, int* dst) {
    { std::lock_guard<std::mutex> lk(mu); jobs.emplace_back(b, dst); }
    cv.notify_all();
  }
  bool drain() { // true if every row arrived
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return jobs.empty() && !busy[0] && !busy[1]; });
    return !failed;
  }
  ~RowCopier() {
    { std::lock_guard<std::mutex> lk(mu); stop = true; }
    cv.notify_all();
    th.join();
  }
};

int ensure_sample_pipeline(exb200_handle* h) {
  if (h->pipeline_ready) return EXB200_OK;
  // (a failed earlier attempt leaves what it allocated in place: only the missing pieces are made)
  if (!h->copy_stream) CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  for (int b = 0; b < 2; b++) {
    if (!h->d_snap[b]) { int rc = dalloc(h, &h->d_snap[b], (size_t)h->N); if (rc) return rc; }
    if (!h->pin_snap[b]) h->pin_snap[b] = (int*)g_pool.take_pin(sizeof(int) * (size_t)h->N);
    if (!h->pin_snap[b] && cudaMallocHost((void**)&h->pin_snap[b], sizeof(int) * (size_t)h->N) != cudaSuccess) {
      h->pin_snap[b] = nullptr;
      cudaGetLastError();
      return fail(h, EXB200_ERR_OOM, "cannot allocate pinned staging for the history");
    }
    if (!h->ev_snap[b]) CK(cudaEventCreateWithFlags(&h->ev_snap[b], cudaEventDisableTiming));
    if (!h->ev_d2h[b]) CK(cudaEventCreateWithFlags(&h->ev_d2h[b], cudaEventDisableTiming));
  }
  h->pipeline_ready = true;
  return EXB200_OK;
}
} // namespace

int exb200_run(exb200_handle* h, int32_t n_samples, int32_t thin, int32_t burnin, exb200_history* out,
               int (*abort_cb)(void*), void (*progress_cb)(void*, int32_t), void* ctx) {
  if (!h || !out) return EXB200_ERR_INVALID;
  if (n_samples < 0 || thin < 1 || burnin < 0) return fail(h, EXB200_ERR_INVALID, "need n_samples >= 0, thin_interval >= 1, burnin_interval >= 0");
  if (h->poisoned) return fail(h, EXB200_ERR_CUDA, "handle unusable after an earlier error: " + h->err);
  cudaSetDevice(h->device);
  const int N = h->N, FA = h->F * h->A;
  out->n_saved = 0;
  if (out->links && n_samples > 0) { int rc = ensure_sample_pipeline(h); if (rc) return rc; }
  RowCopier copier(h);
  auto finish = [&](int rc, int sample) { // rows still in flight land before the call returns
    if (!copier.drain() && !rc) { h->poisoned = true; rc = fail(h, EXB200_ERR_CUDA, "copying a saved sample failed"); }
    out->n_saved = sample;
    return rc;
  };
  int sample = 0;
  double wait_ms = 0.0; // time spent waiting for a staging buffer (EXB200_TRACE)
  const int initial = h->iteration;
  while (sample < n_samples) { // src/sample.cpp:171-202
    int rc = sweep(h);
    if (rc) return finish(rc, sample);
    const int completed = h->iteration - initial;
    if (completed >= burnin && (completed - burnin) % thin == 0) {
      if (out->links) {
        const int b = sample & 1;
        const auto t_w = std::chrono::steady_clock::now();
        copier.acquire(b);
        wait_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_w).count();
        cudaError_t e = cudaMemcpyAsync(h->d_snap[b], h->S.lambda, sizeof(int) * N, cudaMemcpyDeviceToDevice, h->stream);
        if (e == cudaSuccess) e = cudaEventRecord(h->ev_snap[b], h->stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(h->copy_stream, h->ev_snap[b], 0);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h->pin_snap[b], h->d_snap[b], sizeof(int) * N, cudaMemcpyDeviceToHost, h->copy_stream);
        if (e == cudaSuccess) e = cudaEventRecord(h->ev_d2h[b], h->copy_stream);
        // a failed enqueue releases the buffer without copying stale pinned data into the caller's row
        copier.submit(b, e == cudaSuccess ? out->links + (size_t)sample * N : nullptr);
        if (e != cudaSuccess) {
          h->poisoned = true;
          return finish(fail(h, EXB200_ERR_CUDA, std::string("history copy: ") + cudaGetErrorString(e)), sample);
        }
      }
      if (h->pe) { const int rc2 = pe_add_chain_links(h->pe, h); if (rc2) return finish(rc2, sample); }
      if (out->n_linked_ents) out->n_linked_ents[sample] = h->K;
      for (int f = 0; f < h->F; f++)
        for (int a = 0; a < h->A; a++) {
          if (out->distort_probs) out->distort_probs[(size_t)sample * FA + f + a * h->F] = h->theta[f * h->A + a];
          if (out->distort_counts) out->distort_counts[(size_t)sample * FA + f + a * h->F] = h->ndist[f * h->A + a];
        }
      if (out->distort_dist_concs)
        for (int a = 0; a < h->A; a++) out->distort_dist_concs[(size_t)sample * h->A + a] = h->conc[a];
      if (out->clust_params) {
        const bool py = h->cl.kind == EXB200_CLUST_PITMAN_YOR;
        out->clust_params[2 * (size_t)sample] = py ? h->cl.alpha : h->cl.kappa;
        out->clust_params[2 * (size_t)sample + 1] = py ? h->cl.d : h->cl.m;
      }
      sample++;
    }
    if (abort_cb && abort_cb(ctx)) return finish(fail(h, EXB200_ERR_ABORTED, "aborted by the caller"), sample);
    if (progress_cb) progress_cb(ctx, completed);
  }
  if (h->trace) std::fprintf(stderr, "[exb200] run: %d samples, %.1f ms waiting for staging buffers\n", sample, wait_ms);
  return finish(EXB200_OK, sample);
}

int exb200_get_state(exb200_handle* h, exb200_state* out) {
  if (!h || !out) return EXB200_ERR_INVALID;
  cudaSetDevice(h->device);
  const int N = h->N, A = h->A, F = h->F;
  cudaStream_t st = h->stream;
  DevState& S = h->S;
  // packed on the device in the R layout, then straight into the caller's arrays (the links-update
  // scratch is free between sweeps: plo = flags, phi_ = positions, cand_id = packed values)
  out->iteration = h->iteration;
  out->n_entities = h->K;
  int* pack = S.cand_id;
  void* big = nullptr;
  if ((unsigned long long)N * (unsigned long long)A > S.pool_cap) { // many attributes: a buffer of its own
    if (cudaMalloc(&big, sizeof(int) * (size_t)N * A) != cudaSuccess) return fail(h, EXB200_ERR_OOM, "cudaMalloc failed in get_state");
    pack = (int*)big;
  }
  struct Release { void* p; ~Release() { if (p) cudaFree(p); } } release{big};
  if (out->links) CK(staged_copy(out->links, S.lambda, sizeof(int) * (size_t)N, false, h->device, st));
  if (out->ent_ids || out->ent_attrs) {
    KLAUNCH(h, k_pack_flags, nblk(N), TPB, 0, S, S.plo);
    int rc = scan_exclusive(h, S.plo, S.phi_, N); if (rc) return rc;
    int* ids = S.act_in;      // [N]
    int* attrs = pack;        // [>= N * A]
    KLAUNCH(h, k_pack_entities, nblk(N), TPB, 0, S, S.phi_, out->ent_ids ? ids : nullptr, out->ent_attrs ? attrs : nullptr);
    if (out->ent_ids) CK(staged_copy(out->ent_ids, ids, sizeof(int) * (size_t)h->K, false, h->device, st));
    if (out->ent_attrs) CK(staged_copy(out->ent_attrs, attrs, sizeof(int) * (size_t)h->K * A, false, h->device, st));
    CK(cudaStreamSynchronize(st)); // the packing buffer is reused below
  }
  if (out->rec_distortions) {
    KLAUNCH(h, k_pack_distortions, nblk((long long)N * A), TPB, 0, S, pack);
    CK(staged_copy(out->rec_distortions, pack, sizeof(int) * (size_t)N * A, false, h->device, st));
  }
  CK(cudaStreamSynchronize(st));
  for (int f = 0; f < F; f++)
    for (int a = 0; a < A; a++) {
      if (out->distort_probs) out->distort_probs[f + (size_t)F * a] = h->theta[f * A + a];
      if (out->distort_counts) out->distort_counts[f + (size_t)F * a] = h->ndist[f * A + a];
    }
  if (out->distort_dist_concs) for (int a = 0; a < A; a++) out->distort_dist_concs[a] = h->conc[a];
  out->clust = h->cl;
  return EXB200_OK;
}

namespace {
int shard_ready(exb200_handle* h) {
  if (!h) return EXB200_ERR_INVALID;
  if (h->poisoned) return fail(h, EXB200_ERR_CUDA, "handle unusable after an earlier error: " + h->err);
  cudaSetDevice(h->device);
  for (int a = 0; a < h->A; a++)
    if (h->attrs[a].dirichlet || h->attrs[a].conc_shape > 0)
      return fail(h, EXB200_ERR_UNSUPPORTED, "within-chain sharding: DirichletRV entity priors and GammaRV concentration priors are not supported");
  return EXB200_OK;
}
} // namespace

int exb200_shard_links(exb200_handle* h) {
  int rc = shard_ready(h); if (rc) return rc;
  return sweep_links(h);
}
int exb200_shard_entities(exb200_handle* h, int32_t lo, int32_t hi) {
  int rc = shard_ready(h); if (rc) return rc;
  if (lo < 0 || hi > h->N || lo > hi) return fail(h, EXB200_ERR_INVALID, "slice out of range");
  return sweep_entities(h, lo, hi);
}
int exb200_shard_distort(exb200_handle* h, int32_t lo, int32_t hi, int32_t* ndist, int32_t* corr) {
  int rc = shard_ready(h); if (rc) return rc;
  if (lo < 0 || hi > h->N || lo > hi || !ndist || !corr) return fail(h, EXB200_ERR_INVALID, "slice out of range");
  rc = sweep_distort(h, lo, hi); if (rc) return rc;
  std::memcpy(ndist, h->ndist.data(), sizeof(int) * h->ndist.size());
  std::memcpy(corr, h->corr.data(), sizeof(int) * h->corr.size());
  return EXB200_OK;
}
int exb200_shard_finish(exb200_handle* h, const int32_t* ndist, const int32_t* corr) {
  int rc = shard_ready(h); if (rc) return rc;
  if (!ndist || !corr) return EXB200_ERR_INVALID;
  std::memcpy(h->ndist.data(), ndist, sizeof(int) * h->ndist.size());
  std::memcpy(h->corr.data(), corr, sizeof(int) * h->corr.size());
  // the device copies of the counts are only read back by the host; keep them in step anyway
  CK(cudaMemcpyAsync(h->S.ndist, h->ndist.data(), sizeof(int) * h->ndist.size(), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->S.corr, h->corr.data(), sizeof(int) * h->corr.size(), cudaMemcpyHostToDevice, h->stream));
  return sweep_finish(h);
}
int exb200_device_buffer(exb200_handle* h, const char* name, void** ptr, int64_t* bytes, int32_t* row_bytes) {
  if (!h || !name || !ptr || !bytes || !row_bytes) return EXB200_ERR_INVALID;
  const std::string n(name);
  if (n == "ent_y") { *ptr = h->S.ent_y; *row_bytes = (int32_t)(sizeof(int) * h->ES); *bytes = (int64_t)2 * h->N * *row_bytes; }
  else if (n == "rec_z") { *ptr = const_cast<uint32_t*>(h->S.rec_z); *row_bytes = (int32_t)sizeof(uint32_t); *bytes = (int64_t)h->N * *row_bytes; }
  else return fail(h, EXB200_ERR_INVALID, "unknown buffer " + n);
  return EXB200_OK;
}
int exb200_stream_sync(exb200_handle* h) {
  if (!h) return EXB200_ERR_INVALID;
  cudaSetDevice(h->device);
  CK(cudaStreamSynchronize(h->stream));
  return EXB200_OK;
}

int exb200_get_stats(exb200_handle* h, exb200_stats* out) {
  if (!h || !out) return EXB200_ERR_INVALID;
  *out = h->stats;
  return EXB200_OK;
}

int exb200_link_conditional(exb200_handle* h, int32_t rid, int32_t cap, int32_t* cand_ids, double* log_weights,
                            double* log_weight_new) {
  if (!h || rid < 0 || rid >= h->N) return EXB200_ERR_INVALID;
  cudaSetDevice(h->device);
  const int N = h->N;
  refresh_clust_dev(h);
  h->S.iter = (uint32_t)h->iteration;
  KLAUNCH(h, k_eval_all, nblk(N), TPB, 0, h->S, rid, N); // between sweeps only slots [0,N) are live
  std::vector<double> lw(N);
  std::vector<int> x(h->RS);
  uint32_t zm = 0;
  int own = 0, own_size = 0;
  CK(cudaMemcpyAsync(lw.data(), h->S.lwbuf, sizeof(double) * N, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(x.data(), h->S.rec_x + (size_t)rid * h->RS, sizeof(int) * h->RS, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(&zm, h->S.rec_z + rid, sizeof zm, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(&own, h->S.lambda + rid, sizeof own, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaMemcpy(&own_size, h->S.ent_size + own, sizeof own_size, cudaMemcpyDeviceToHost));
  int n = 0;
  for (int e = 0; e < N; e++)
    if (lw[e] == lw[e]) {
      if (n < cap) { cand_ids[n] = e; log_weights[n] = lw[e]; }
      n++;
    }
  // logLikelihoodWeightNew from the host copies of the tables (links.cpp:316-347)
  double L = 0.0;
  for (int a = 0; a < h->A; a++) {
    if (x[a] < 0) continue;
    const HostAttr& t = h->attrs[a];
    double v;
    if ((zm >> a) & 1u) {
      std::vector<double> logE(t.D);
      CK(cudaMemcpy(logE.data(), h->h_attrs[a].logE, sizeof(double) * t.D, cudaMemcpyDeviceToHost));
      v = logE[x[a]];
    } else v = std::log(t.phi[x[a]]);
    L += v;
  }
  const int k = h->K - (own_size == 1 ? 1 : 0);
  double pn;
  if (h->cl.kind == EXB200_CLUST_PITMAN_YOR) pn = h->cl.alpha + h->cl.d * (double)k;
  else if (h->cl.kind == EXB200_CLUST_GEN_COUPON) pn = h->cl.kappa * std::fabs(h->cl.m - (double)k);
  else pn = std::fabs(h->cl.m - (double)k);
  *log_weight_new = std::log(pn) + L;
  return n;
}

} // extern "C"

// =============================================================================================
// value neighbourhoods (one-off model preparation; R/AttributeIndex.R:117-137)
struct exb200_neighbours {
  std::vector<int64_t> row_ptr;
  std::vector<int32_t> cols;
  std::vector<double> probs, mef;
};

extern "C" {

int exb200_neighbours_create(const uint8_t* chars, const int32_t* lens, int32_t n, int32_t threshold,
                             int32_t include_self, int device, exb200_neighbours** out) {
  if (!chars || !lens || !out || n < 1 || threshold < 0 || threshold > 8)
    return fail(nullptr, EXB200_ERR_INVALID, "neighbours: bad arguments (threshold must be an integer in [0, 8])");
  for (int i = 0; i < n; i++)
    if (lens[i] < 0 || lens[i] >= NB_STRIDE) return fail(nullptr, EXB200_ERR_INVALID, "neighbours: strings must be shorter than 32 bytes");
  if (exb200_device_count() <= 0) return fail(nullptr, EXB200_ERR_NO_DEVICE, "no CUDA device is visible; this library has no CPU path");
  if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, EXB200_ERR_NO_DEVICE, "cudaSetDevice failed");
  uint8_t *d_str = nullptr, *d_dist = nullptr;
  int *d_len = nullptr, *d_cnt = nullptr, *d_col = nullptr;
  long long* d_row = nullptr;
  auto cleanup = [&]() { cudaFree(d_str); cudaFree(d_dist); cudaFree(d_len); cudaFree(d_cnt); cudaFree(d_col); cudaFree(d_row); };
#define NCK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(); return fail(nullptr, EXB200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); } } while (0)
  NCK(cudaMalloc(&d_str, (size_t)n * NB_STRIDE));
  NCK(cudaMalloc(&d_len, sizeof(int) * (size_t)n));
  NCK(cudaMalloc(&d_cnt, sizeof(int) * (size_t)n));
  NCK(cudaMalloc(&d_row, sizeof(long long) * ((size_t)n + 1)));
  NCK(cudaMemcpy(d_str, chars, (size_t)n * NB_STRIDE, cudaMemcpyHostToDevice));
  NCK(cudaMemcpy(d_len, lens, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice));
  const int nb = (n + NB_TILE - 1) / NB_TILE;
  k_neighbours<<<nb, NB_TILE>>>(d_str, d_len, n, threshold, include_self, 0, d_cnt, nullptr, nullptr, nullptr);
  std::vector<int> cnt(n);
  NCK(cudaMemcpy(cnt.data(), d_cnt, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost));
  exb200_neighbours* r = new exb200_neighbours();
  r->row_ptr.assign((size_t)n + 1, 0);
  for (int i = 0; i < n; i++) r->row_ptr[i + 1] = r->row_ptr[i] + cnt[i];
  const int64_t nnz = r->row_ptr[n];
  std::vector<uint8_t> dist((size_t)nnz);
  r->cols.resize((size_t)nnz);
  if (nnz > 0) {
    static_assert(sizeof(long long) == sizeof(int64_t), "row pointer width");
    cudaError_t e1 = cudaMalloc(&d_col, sizeof(int) * (size_t)nnz), e2 = cudaMalloc(&d_dist, (size_t)nnz);
    if (e1 != cudaSuccess || e2 != cudaSuccess) { delete r; cleanup(); return fail(nullptr, EXB200_ERR_OOM, "neighbours: out of device memory"); }
    cudaMemcpy(d_row, r->row_ptr.data(), sizeof(long long) * ((size_t)n + 1), cudaMemcpyHostToDevice);
    k_neighbours<<<nb, NB_TILE>>>(d_str, d_len, n, threshold, include_self, 1, nullptr, d_row, d_col, d_dist);
    cudaError_t e3 = cudaMemcpy(r->cols.data(), d_col, sizeof(int) * (size_t)nnz, cudaMemcpyDeviceToHost);
    cudaError_t e4 = cudaMemcpy(dist.data(), d_dist, (size_t)nnz, cudaMemcpyDeviceToHost);
    if (e3 != cudaSuccess || e4 != cudaSuccess) { delete r; cleanup(); return fail(nullptr, EXB200_ERR_CUDA, std::string("neighbours: ") + cudaGetErrorString(e3 != cudaSuccess ? e3 : e4)); }
  }
  cleanup();
#undef NCK
  // probabilities exp(-d) / sum and max_exp_factor = exp(-0.5 min_d / max_d) (AttributeIndex.R:128-135)
  r->probs.resize((size_t)nnz);
  r->mef.assign((size_t)n, 0.0);
  double max_d = 0.0;
  for (int64_t j = 0; j < nnz; j++) max_d = std::max(max_d, (double)dist[j]);
  for (int i = 0; i < n; i++) {
    double sum = 0.0, min_d = std::numeric_limits<double>::infinity();
    for (int64_t j = r->row_ptr[i]; j < r->row_ptr[i + 1]; j++) {
      r->probs[j] = std::exp(-(double)dist[j]);
      sum += r->probs[j];
      min_d = std::min(min_d, (double)dist[j]);
    }
    for (int64_t j = r->row_ptr[i]; j < r->row_ptr[i + 1]; j++) r->probs[j] /= sum;
    r->mef[i] = std::isfinite(min_d) ? std::exp(-0.5 * min_d / max_d) : 0.0;
  }
  *out = r;
  return EXB200_OK;
}

// ---------------------------------------------------------------- point estimates on the device
struct exb200_pe {
  int device = 0, N = 0, slots = 8, n_samples = 0;
  cudaStream_t stream = nullptr;
  PeSlot* table = nullptr;
  unsigned long long *fp1 = nullptr, *fp2 = nullptr;
  int* links = nullptr;
  int* misc = nullptr; // [0] label out of range, [1] overflow count
  uint8_t* over_flag = nullptr; // [N] 1: the record was seen in more than `slots` distinct clusters
  int* pairs = nullptr; int* pair_counts = nullptr; // coreference counting (exb200_pe_set_pairs)
  long long n_pairs = 0;
  int pair_samples = 0;
};

void exb200_pe_destroy(exb200_pe* pe) {
  if (!pe) return;
  cudaSetDevice(pe->device);
  if (pe->stream) cudaStreamSynchronize(pe->stream);
  cudaFree(pe->table); cudaFree(pe->fp1); cudaFree(pe->fp2); cudaFree(pe->links); cudaFree(pe->misc); cudaFree(pe->over_flag);
  cudaFree(pe->pairs); cudaFree(pe->pair_counts);
  if (pe->stream) cudaStreamDestroy(pe->stream);
  delete pe;
}

int exb200_pe_create(int32_t n_records, int32_t slots, int device, exb200_pe** out) {
  if (!out || n_records < 1 || slots < 1 || slots > 64) return fail(nullptr, EXB200_ERR_INVALID, "point estimate: need n_records >= 1, 1 <= slots <= 64");
  *out = nullptr;
  if (exb200_device_count() <= 0) return fail(nullptr, EXB200_ERR_NO_DEVICE, "no CUDA device is visible; this library has no CPU path");
  if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, EXB200_ERR_NO_DEVICE, "invalid device index");
  exb200_pe* pe = new exb200_pe();
  pe->device = device; pe->N = n_records; pe->slots = slots;
  const size_t n = (size_t)n_records;
  bool ok = cudaStreamCreateWithFlags(&pe->stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaMalloc((void**)&pe->table, sizeof(PeSlot) * n * slots) == cudaSuccess &&
            cudaMalloc((void**)&pe->fp1, sizeof(unsigned long long) * 2 * n) == cudaSuccess &&
            cudaMalloc((void**)&pe->fp2, sizeof(unsigned long long) * 2 * n) == cudaSuccess &&
            cudaMalloc((void**)&pe->links, sizeof(int) * n) == cudaSuccess &&
            cudaMalloc((void**)&pe->misc, sizeof(int) * 2) == cudaSuccess &&
            cudaMalloc((void**)&pe->over_flag, n) == cudaSuccess;
  if (ok) ok = cudaMemsetAsync(pe->table, 0, sizeof(PeSlot) * n * slots, pe->stream) == cudaSuccess &&
               cudaMemsetAsync(pe->over_flag, 0, n, pe->stream) == cudaSuccess &&
               cudaMemsetAsync(pe->misc, 0, sizeof(int) * 2, pe->stream) == cudaSuccess &&
               cudaStreamSynchronize(pe->stream) == cudaSuccess;
  if (!ok) { exb200_pe_destroy(pe); return fail(nullptr, EXB200_ERR_OOM, "point estimate: device allocation failed"); }
  *out = pe;
  return EXB200_OK;
}

namespace {
int pe_add_device(exb200_pe* pe, const int* d_links) {
  const int N = pe->N;
  cudaStream_t st = pe->stream;
  cudaMemsetAsync(pe->fp1, 0, sizeof(unsigned long long) * 2 * (size_t)N, st);
  cudaMemsetAsync(pe->fp2, 0, sizeof(unsigned long long) * 2 * (size_t)N, st);
  k_pe_fingerprint<<<nblk(N), TPB, 0, st>>>(N, 2 * N, d_links, pe->fp1, pe->fp2, pe->misc);
  int bad = 0;
  cudaMemcpyAsync(&bad, pe->misc, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (cudaStreamSynchronize(st) != cudaSuccess) return fail(nullptr, EXB200_ERR_CUDA, "point estimate: fingerprint kernel failed");
  if (bad) {
    cudaMemsetAsync(pe->misc, 0, sizeof(int), st);
    return fail(nullptr, EXB200_ERR_INVALID, "point estimate: labels must lie in [0, 2 * n_records)");
  }
  k_pe_count<<<nblk(N), TPB, 0, st>>>(N, pe->slots, pe->n_samples, d_links, pe->fp1, pe->fp2, pe->table, pe->misc + 1, pe->over_flag);
  if (pe->n_pairs > 0) {
    k_pe_pairs<<<nblk(pe->n_pairs), TPB, 0, st>>>(pe->n_pairs, pe->pairs, d_links, pe->pair_counts);
    pe->pair_samples++;
  }
  if (cudaStreamSynchronize(st) != cudaSuccess) return fail(nullptr, EXB200_ERR_CUDA, "point estimate: counting kernel failed");
  pe->n_samples++;
  return EXB200_OK;
}
} // namespace

int exb200_pe_set_pairs(exb200_pe* pe, const int32_t* pairs, int64_t n_pairs) {
  if (!pe || n_pairs < 0 || (n_pairs > 0 && !pairs)) return EXB200_ERR_INVALID;
  cudaSetDevice(pe->device);
  for (int64_t i = 0; i < 2 * n_pairs; i++)
    if (pairs[i] < 0 || pairs[i] >= pe->N) return fail(nullptr, EXB200_ERR_INVALID, "point estimate: `pairs` refers to records that are not there");
  cudaStreamSynchronize(pe->stream);
  cudaFree(pe->pairs); cudaFree(pe->pair_counts);
  pe->pairs = pe->pair_counts = nullptr; pe->n_pairs = 0; pe->pair_samples = 0;
  if (n_pairs == 0) return EXB200_OK;
  if (cudaMalloc((void**)&pe->pairs, sizeof(int) * 2 * (size_t)n_pairs) != cudaSuccess ||
      cudaMalloc((void**)&pe->pair_counts, sizeof(int) * (size_t)n_pairs) != cudaSuccess)
    return fail(nullptr, EXB200_ERR_OOM, "point estimate: device allocation failed");
  cudaMemcpyAsync(pe->pairs, pairs, sizeof(int) * 2 * (size_t)n_pairs, cudaMemcpyHostToDevice, pe->stream);
  cudaMemsetAsync(pe->pair_counts, 0, sizeof(int) * (size_t)n_pairs, pe->stream);
  if (cudaStreamSynchronize(pe->stream) != cudaSuccess) return fail(nullptr, EXB200_ERR_CUDA, "point estimate: upload failed");
  pe->n_pairs = n_pairs;
  return EXB200_OK;
}

int exb200_pe_pair_counts(exb200_pe* pe, int32_t* counts, int32_t* n_samples) {
  if (!pe || !counts) return EXB200_ERR_INVALID;
  cudaSetDevice(pe->device);
  if (pe->n_pairs > 0 && (cudaMemcpyAsync(counts, pe->pair_counts, sizeof(int) * (size_t)pe->n_pairs, cudaMemcpyDeviceToHost, pe->stream) != cudaSuccess ||
                          cudaStreamSynchronize(pe->stream) != cudaSuccess))
    return fail(nullptr, EXB200_ERR_CUDA, "point estimate: download failed");
  if (n_samples) *n_samples = pe->pair_samples;
  return EXB200_OK;
}

int exb200_pe_add(exb200_pe* pe, const int32_t* links) {
  if (!pe || !links) return EXB200_ERR_INVALID;
  cudaSetDevice(pe->device);
  if (cudaMemcpyAsync(pe->links, links, sizeof(int) * (size_t)pe->N, cudaMemcpyHostToDevice, pe->stream) != cudaSuccess)
    return fail(nullptr, EXB200_ERR_CUDA, "point estimate: upload failed");
  return pe_add_device(pe, pe->links);
}

int pe_add_chain_links(exb200_pe* pe, exb200_handle* h) { return exb200_pe_add_chain(pe, h); }

int exb200_attach_point_estimate(exb200_handle* h, exb200_pe* pe) {
  if (!h) return EXB200_ERR_INVALID;
  if (pe && (pe->N != h->N || pe->device != h->device)) return fail(h, EXB200_ERR_INVALID, "point estimate of another size or device");
  h->pe = pe;
  return EXB200_OK;
}

int exb200_pe_add_chain(exb200_pe* pe, exb200_handle* h) {
  if (!pe || !h) return EXB200_ERR_INVALID;
  if (h->N != pe->N || h->device != pe->device) return fail(nullptr, EXB200_ERR_INVALID, "point estimate: chain of another size or device");
  cudaSetDevice(pe->device);
  if (cudaStreamSynchronize(h->stream) != cudaSuccess) return fail(nullptr, EXB200_ERR_CUDA, "point estimate: chain stream failed");
  return pe_add_device(pe, h->S.lambda);
}

int exb200_pe_finish(exb200_pe* pe, int32_t* smp_membership, double* mp_prob, int32_t* mp_sample, int32_t* tied,
                     int64_t* n_overflow) {
  if (!pe) return EXB200_ERR_INVALID;
  if (pe->n_samples < 1) return fail(nullptr, EXB200_ERR_INVALID, "point estimate: no sample added");
  cudaSetDevice(pe->device);
  const int N = pe->N;
  const size_t n = (size_t)N;
  cudaStream_t st = pe->stream;
  unsigned long long *k1 = nullptr, *k2 = nullptr, *k1s = nullptr;
  int *cnt = nullptr, *smp = nullptr, *tie = nullptr, *ids = nullptr, *ids_s = nullptr, *head = nullptr, *memb = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k1, k1s, ids, ids_s, N, 0, 64, st);
  std::vector<void*> held;
  auto dal = [&](auto** p, size_t bytes) { void* q = nullptr; if (cudaMalloc(&q, std::max<size_t>(bytes, 1)) != cudaSuccess) return false; held.push_back(q); *p = (std::remove_reference_t<decltype(**p)>*)q; return true; };
  auto release = [&] { for (void* q : held) cudaFree(q); };
  bool ok = dal(&k1, 8 * n) && dal(&k2, 8 * n) && dal(&k1s, 8 * n) && dal(&cnt, 4 * n) && dal(&smp, 4 * n) && dal(&tie, 4 * n) &&
            dal(&ids, 4 * n) && dal(&ids_s, 4 * n) && dal(&head, 4 * n) && dal(&memb, 4 * n);
  if (ok) { ok = cudaMalloc(&tmp, std::max<size_t>(tmp_bytes, 1)) == cudaSuccess; if (ok) held.push_back(tmp); }
  if (!ok) { release(); return fail(nullptr, EXB200_ERR_OOM, "point estimate: device allocation failed"); }
  k_pe_best<<<nblk(N), TPB, 0, st>>>(N, pe->slots, pe->table, k1, k2, cnt, smp, tie, ids);
  cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k1, k1s, ids, ids_s, N, 0, 64, st);
  if (e == cudaSuccess) {
    k_pe_heads<<<nblk(N), TPB, 0, st>>>(N, k1s, ids_s, k2, head);
    k_pe_membership<<<nblk(N), TPB, 0, st>>>(N, ids_s, head, memb);
  }
  std::vector<int> h_cnt;
  if (e == cudaSuccess && smp_membership) e = cudaMemcpyAsync(smp_membership, memb, 4 * n, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && mp_sample) e = cudaMemcpyAsync(mp_sample, smp, 4 * n, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && tied) e = cudaMemcpyAsync(tied, tie, 4 * n, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && mp_prob) { h_cnt.resize(n); e = cudaMemcpyAsync(h_cnt.data(), cnt, 4 * n, cudaMemcpyDeviceToHost, st); }
  int over = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&over, pe->misc + 1, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  release();
  if (e != cudaSuccess) return fail(nullptr, EXB200_ERR_CUDA, std::string("point estimate: ") + cudaGetErrorString(e));
  if (mp_prob) for (size_t r = 0; r < n; r++) mp_prob[r] = 1.0 / (double)pe->n_samples * (double)h_cnt[r]; // point_estimate.cpp:97
  if (n_overflow) *n_overflow = over;
  return EXB200_OK;
}

int exb200_pe_overflowed(exb200_pe* pe, uint8_t* flags) {
  if (!pe || !flags) return EXB200_ERR_INVALID;
  cudaSetDevice(pe->device);
  if (cudaMemcpyAsync(flags, pe->over_flag, (size_t)pe->N, cudaMemcpyDeviceToHost, pe->stream) != cudaSuccess ||
      cudaStreamSynchronize(pe->stream) != cudaSuccess)
    return fail(nullptr, EXB200_ERR_CUDA, "point estimate: download failed");
  return EXB200_OK;
}

int64_t exb200_neighbours_nnz(const exb200_neighbours* nb) { return nb ? (int64_t)nb->cols.size() : -1; }

int exb200_neighbours_get(const exb200_neighbours* nb, int64_t* row_ptr, int32_t* val_ids, double* probs,
                          double* max_exp_factor) {
  if (!nb) return EXB200_ERR_INVALID;
  if (row_ptr) std::memcpy(row_ptr, nb->row_ptr.data(), sizeof(int64_t) * nb->row_ptr.size());
  if (val_ids) std::memcpy(val_ids, nb->cols.data(), sizeof(int32_t) * nb->cols.size());
  if (probs) std::memcpy(probs, nb->probs.data(), sizeof(double) * nb->probs.size());
  if (max_exp_factor) std::memcpy(max_exp_factor, nb->mef.data(), sizeof(double) * nb->mef.size());
  return EXB200_OK;
}

void exb200_neighbours_destroy(exb200_neighbours* nb) { delete nb; }

} // extern "C"
