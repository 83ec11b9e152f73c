// Point estimates from the stream of saved link vectors, on the device:
// most probable cluster of every record and the shared most probable clusters
// (mp_clusters / smp_clusters, src/point_estimate.cpp:58-136) without keeping the n_samples x N
// history anywhere.  A cluster is identified by a 128-bit fingerprint of its member set (sum over
// the members of two independent 64-bit mixes: order-free, label-free); every record keeps the
// clusters it has been seen in, with their frequencies, in a small table.
#pragma once
#include <cub/device/device_radix_sort.cuh>

#include "device_state.cuh"

namespace exb {

struct PeSlot { unsigned long long k1, k2; int count, first_sample; };

__device__ __forceinline__ unsigned long long pe_mix(unsigned long long z, unsigned long long salt) {
  z += salt;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
constexpr unsigned long long PE_SALT1 = 0x9E3779B97F4A7C15ull, PE_SALT2 = 0xD1B54A32D192ED03ull;

// fingerprints of the clusters of one sample: fp[label] += mix(record)
__global__ void __launch_bounds__(TPB) k_pe_fingerprint(int n, int n_labels, const int* __restrict__ links,
                                                        unsigned long long* __restrict__ fp1, unsigned long long* __restrict__ fp2,
                                                        int* __restrict__ err) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int lab = links[r];
  if (lab < 0 || lab >= n_labels) { *err = 1; return; }
  atomicAdd(&fp1[lab], pe_mix((unsigned long long)r, PE_SALT1));
  atomicAdd(&fp2[lab], pe_mix((unsigned long long)r, PE_SALT2));
}
// every record counts the cluster it is in (its own table: no atomics)
__global__ void __launch_bounds__(TPB) k_pe_count(int n, int slots, int sample, const int* __restrict__ links,
                                                  const unsigned long long* __restrict__ fp1, const unsigned long long* __restrict__ fp2,
                                                  PeSlot* __restrict__ table, int* __restrict__ overflow,
                                                  uint8_t* __restrict__ over_flag) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int lab = links[r];
  const unsigned long long k1 = fp1[lab], k2 = fp2[lab];
  PeSlot* t = table + (size_t)r * slots;
  for (int s = 0; s < slots; s++) {
    if (t[s].count == 0) { t[s].k1 = k1; t[s].k2 = k2; t[s].count = 1; t[s].first_sample = sample; return; }
    if (t[s].k1 == k1 && t[s].k2 == k2) { t[s].count++; return; }
  }
  atomicAdd(overflow, 1); // more distinct clusters than slots: this occurrence is not counted ...
  over_flag[r] = 1;       // ... and the record is flagged: its estimate has to be redone from the history
}
// most frequent cluster of every record; ties go to the smaller fingerprint (the reference breaks
// them by the lexicographic order of the member sets, which a fingerprint cannot reproduce: the
// host wrapper settles tied records from the history when it has one)
__global__ void __launch_bounds__(TPB) k_pe_best(int n, int slots, const PeSlot* __restrict__ table,
                                                 unsigned long long* __restrict__ best_k1, unsigned long long* __restrict__ best_k2,
                                                 int* __restrict__ best_count, int* __restrict__ best_sample, int* __restrict__ tied,
                                                 int* __restrict__ ids) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const PeSlot* t = table + (size_t)r * slots;
  int b = -1, ties = 0;
  for (int s = 0; s < slots && t[s].count > 0; s++) {
    if (b < 0 || t[s].count > t[b].count) { b = s; ties = 0; }
    else if (t[s].count == t[b].count) {
      ties = 1;
      if (t[s].k1 < t[b].k1 || (t[s].k1 == t[b].k1 && t[s].k2 < t[b].k2)) b = s;
    }
  }
  best_k1[r] = b >= 0 ? t[b].k1 : 0ull;
  best_k2[r] = b >= 0 ? t[b].k2 : 0ull;
  best_count[r] = b >= 0 ? t[b].count : 0;
  best_sample[r] = b >= 0 ? t[b].first_sample : -1;
  tied[r] = ties;
  ids[r] = r;
}
// records sorted by the fingerprint of their most probable cluster (stable: ascending record id
// inside a group): a group starts where the fingerprint changes and is named after its first record
__global__ void __launch_bounds__(TPB) k_pe_heads(int n, const unsigned long long* __restrict__ k1_sorted, const int* __restrict__ ids_sorted,
                                                  const unsigned long long* __restrict__ best_k2, int* __restrict__ head_flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  head_flag[i] = i == 0 || k1_sorted[i] != k1_sorted[i - 1] || best_k2[ids_sorted[i]] != best_k2[ids_sorted[i - 1]];
}
__global__ void __launch_bounds__(TPB) k_pe_membership(int n, const int* __restrict__ ids_sorted, const int* __restrict__ head_flag,
                                                       int* __restrict__ membership) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int j = i; // walk back to the head of the group (groups are clusters: short)
  while (!head_flag[j]) j--;
  membership[ids_sorted[i]] = ids_sorted[j];
}

// coreference counts of given record pairs (numerator of coreference_probs, R/coreference_probs.R:77-79)
__global__ void __launch_bounds__(TPB) k_pe_pairs(long long n_pairs, const int* __restrict__ pairs, const int* __restrict__ links,
                                                  int* __restrict__ counts) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n_pairs && links[pairs[2 * p]] == links[pairs[2 * p + 1]]) counts[p]++;
}

} // namespace exb
