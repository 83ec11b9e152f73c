// Thresholded-Levenshtein value neighbourhoods on the GPU: the device form of
// compute_close_values (R/AttributeIndex.R:117-137) for dist_fn =
// transform_dist_fn(Levenshtein(), threshold) (R/transform_dist_fn.R:27-40).  The reference
// evaluates the D x D distance matrix through R-level calls (optionally future_lapply); at the
// domain sizes of a 10 M-record model (D ~ 1e5) that is 1e10 distance evaluations, which is why
// this one-off precompute gets a kernel.  Not part of the timed sweep.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace exb {

constexpr int NB_STRIDE = 32;  // bytes per packed string (length <= 31)
constexpr int NB_TILE = 256;

// unit-cost edit distance of a (length la) and b (length lb), exact up to `thr`, thr+1 otherwise.
// Banded dynamic programme (|i - j| <= thr), rows kept in registers.
__device__ __forceinline__ int lev_banded(const uint8_t* a, int la, const uint8_t* b, int lb, int thr) {
  if (abs(la - lb) > thr) return thr + 1;
  int prev[NB_STRIDE], cur[NB_STRIDE];
#pragma unroll
  for (int j = 0; j < NB_STRIDE; j++) prev[j] = j;
  for (int i = 1; i <= la; i++) {
    const int jlo = max(1, i - thr), jhi = min(lb, i + thr);
    int best = thr + 1;
    cur[jlo - 1] = (jlo == 1) ? i : thr + 1;
    const uint8_t ca = a[i - 1];
    for (int j = jlo; j <= jhi; j++) {
      const int sub = prev[j - 1] + (ca != b[j - 1]);
      const int del = (j <= i - 1 + thr) ? prev[j] + 1 : thr + 1;
      const int ins = cur[j - 1] + 1;
      const int v = min(sub, min(del, ins));
      cur[j] = v;
      best = min(best, v);
    }
    if (jlo == 1) best = min(best, i);
    if (best > thr) return thr + 1;
#pragma unroll
    for (int j = 0; j < NB_STRIDE; j++) prev[j] = cur[j];
    if (jhi < lb) prev[jhi + 1] = thr + 1;
  }
  return min(prev[lb], thr + 1);
}

// pass 0: counts per row; pass 1: fill (cols, dists) at row offsets.  One thread per value i,
// the other values j stream through shared memory in tiles.
__global__ void __launch_bounds__(NB_TILE) k_neighbours(const uint8_t* __restrict__ strs, const int* __restrict__ lens,
                                                        int D, int thr, int include_self, int pass,
                                                        int* __restrict__ counts, const long long* __restrict__ row_ptr,
                                                        int* __restrict__ cols, uint8_t* __restrict__ dists) {
  __shared__ uint8_t s_str[NB_TILE][NB_STRIDE];
  __shared__ int s_len[NB_TILE];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  uint8_t a[NB_STRIDE];
  int la = 0;
  if (i < D) {
    la = lens[i];
    const uint4* src = reinterpret_cast<const uint4*>(strs + (size_t)i * NB_STRIDE);
    *reinterpret_cast<uint4*>(a) = src[0];
    *reinterpret_cast<uint4*>(a + 16) = src[1];
  }
  int n = 0;
  long long out = (pass == 1 && i < D) ? row_ptr[i] : 0;
  for (int base = 0; base < D; base += NB_TILE) {
    __syncthreads();
    const int j = base + threadIdx.x;
    if (j < D) {
      const uint4* src = reinterpret_cast<const uint4*>(strs + (size_t)j * NB_STRIDE);
      *reinterpret_cast<uint4*>(s_str[threadIdx.x]) = src[0];
      *reinterpret_cast<uint4*>(s_str[threadIdx.x] + 16) = src[1];
      s_len[threadIdx.x] = lens[j];
    }
    __syncthreads();
    if (i < D) {
      const int lim = min(NB_TILE, D - base);
      for (int k = 0; k < lim; k++) {
        const int jj = base + k;
        if (jj == i && !include_self) continue;
        const int d = lev_banded(a, la, s_str[k], s_len[k], thr);
        if (d <= thr) {
          if (pass == 1) { cols[out] = jj; dists[out] = (uint8_t)d; out++; }
          n++;
        }
      }
    }
  }
  if (pass == 0 && i < D) counts[i] = n;
}

} // namespace exb
