// Device side of exb200_create / exb200_get_state: the caller's R-layout arrays are uploaded as
// they are and turned into the resident layout here (and back), so that the host never walks the
// records.  Layout conventions of the inputs: src/state.cpp:20-60 (init_state) and R/exchanger.R.
#pragma once
#include <cub/device/device_radix_sort.cuh>

#include "device_state.cuh"

namespace exb {

struct SetupDomains { int D[MAX_ATTRS]; };

enum SetupError { SE_REC_ATTRS = 1, SE_FILE_IDS = 2, SE_LINKS = 3, SE_ENT_ATTRS = 4 };

// records: rec_x rows (attribute codes, -1 = missing, file id in the last column), distortion bits
// and the per-(file, attribute) distortion counts
__global__ void __launch_bounds__(TPB) k_setup_records(DevState S, SetupDomains dom, const int* __restrict__ raw_x,
                                                       const int* __restrict__ raw_z, const int* __restrict__ file_ids,
                                                       int* __restrict__ err) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = r < S.N;
  int f = 0;
  uint32_t zm = 0;
  if (live) {
    int* x = const_cast<int*>(S.rec_x) + (size_t)r * S.RS;
    for (int a = 0; a < S.A; a++) {
      const int v = raw_x[a + (size_t)S.A * r];
      if (v >= dom.D[a]) err[0] = SE_REC_ATTRS;
      x[a] = v < 0 ? -1 : v;
      if (raw_z[a + (size_t)S.A * r]) zm |= 1u << a;
    }
    for (int a = S.A; a < S.RS - 1; a++) x[a] = -1;
    f = file_ids[r];
    if (f < 0 || f >= S.F) { err[0] = SE_FILE_IDS; f = 0; }
    x[S.RS - 1] = f;
    const_cast<uint32_t*>(S.rec_z)[r] = zm;
  }
  // counts: lanes of one warp mostly share the file, so reduce per warp when they do
  for (int a = 0; a < S.A; a++) {
    const int bit = live ? (int)((zm >> a) & 1u) : 0;
    const int f0 = __shfl_sync(FULL, f, 0);
    if (__all_sync(FULL, !live || f == f0)) {
      const int c = __popc(__ballot_sync(FULL, bit));
      if ((threadIdx.x & 31) == 0 && c) atomicAdd(&S.ndist[f0 * S.A + a], c);
    } else if (bit) atomicAdd(&S.ndist[f * S.A + a], 1);
  }
}

// label -> column of ent_attrs (the last column wins for a repeated label, as a sequential fill would)
__global__ void __launch_bounds__(TPB) k_setup_slots(const int* __restrict__ ent_ids, int n_ent, int* __restrict__ lab2slot) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n_ent) atomicMax(&lab2slot[ent_ids[e]], e);
}

// label -> smallest member record
__global__ void __launch_bounds__(TPB) k_setup_rep(int N, const int* __restrict__ links, const int* __restrict__ lab2slot,
                                                   int max_label, int* __restrict__ rep, int* __restrict__ err) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= N) return;
  const int lab = links[r];
  if (lab < 0 || lab > max_label || lab2slot[lab] < 0) { err[0] = SE_LINKS; return; }
  atomicMin(&rep[lab], r);
}

// canonical links, sizes, heads, entity attributes and the (entity, record) sort keys
__global__ void __launch_bounds__(TPB) k_setup_links(DevState S, SetupDomains dom, const int* __restrict__ links,
                                                     const int* __restrict__ lab2slot, const int* __restrict__ rep,
                                                     const int* __restrict__ ent_attrs, unsigned long long* __restrict__ keys,
                                                     int* __restrict__ n_linked, int* __restrict__ err) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  int first = 0;
  if (r < S.N) {
    const int lab = links[r];
    const int c = rep[lab];
    S.lambda[r] = c;
    atomicAdd(&S.ent_size[c], 1);
    keys[r] = ((unsigned long long)(unsigned)c << 32) | (unsigned)r;
    if (c == r) {
      first = 1;
      S.ent_head[c] = r;
      const int col = lab2slot[lab];
      int* y = S.ent_y + (size_t)c * S.ES;
      for (int a = 0; a < S.A; a++) {
        const int v = ent_attrs[a + (size_t)S.A * col];
        if (v < 0 || v >= dom.D[a]) { err[0] = SE_ENT_ATTRS; y[a] = 0; }
        else y[a] = v;
      }
    }
  }
  const int c = __popc(__ballot_sync(FULL, first));
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(n_linked, c);
}

// member lists in ascending record id from the sorted keys
__global__ void __launch_bounds__(TPB) k_setup_next(int N, const unsigned long long* __restrict__ keys, int* __restrict__ rec_next) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const unsigned long long k = keys[i];
  const unsigned long long k2 = i + 1 < N ? keys[i + 1] : ~0ull;
  rec_next[(int)(unsigned)k] = (k2 >> 32) == (k >> 32) ? (int)(unsigned)k2 : -1;
}

// ---- the way back (exb200_get_state): live entities in ascending id, R layout
__global__ void __launch_bounds__(TPB) k_pack_flags(DevState S, int* __restrict__ flag) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < S.N) flag[e] = S.ent_size[e] > 0;
}
__global__ void __launch_bounds__(TPB) k_pack_entities(DevState S, const int* __restrict__ pos, int* __restrict__ ent_ids,
                                                       int* __restrict__ ent_attrs) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= S.N || S.ent_size[e] <= 0) return;
  const int k = pos[e];
  if (ent_ids) ent_ids[k] = e;
  if (ent_attrs) {
    const int* y = S.ent_y + (size_t)e * S.ES;
    for (int a = 0; a < S.A; a++) ent_attrs[a + (size_t)S.A * k] = y[a];
  }
}
__global__ void __launch_bounds__(TPB) k_pack_distortions(DevState S, int* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)S.N * S.A) return;
  const int r = (int)(i / S.A), a = (int)(i % S.A);
  out[i] = (S.rec_z[r] >> a) & 1u;
}

} // namespace exb
