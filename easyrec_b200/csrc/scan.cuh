// Device-wide exclusive prefix sum over int32, reduce-then-scan in three
// launches (tile sums -> scan of tile sums -> tile scan + offset).  Inputs on
// this path are tiny next to HBM bandwidth (<= a few million int32), so the
// design goal is "no host sync, graph capturable, deterministic", not a
// decoupled look-back.
#pragma once
#include "common.cuh"

namespace er {
namespace scan {

constexpr int kThreads = 256;
constexpr int kItems = 8;
constexpr int kTile = kThreads * kItems;  // 2048 elements per CTA

inline int64_t num_tiles(int64_t n) { return n > 0 ? ceil_div(n, kTile) : 1; }
// workspace: tile sums (int32 each), 256-byte aligned by the caller's carve-up
inline size_t workspace_bytes(int64_t n) { return (size_t)num_tiles(n) * sizeof(int32_t) + 256; }

__device__ __forceinline__ int warp_incl_scan(int v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) >= o) v += t;
  }
  return v;
}

// exclusive scan of one int per thread across the CTA; returns the exclusive
// prefix and (optionally) the CTA total.
__device__ __forceinline__ int block_excl_scan(int v, int* total) {
  __shared__ int warp_tot[kThreads / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = warp_incl_scan(v);
  if (lane == 31) warp_tot[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int t = (lane < kThreads / 32) ? warp_tot[lane] : 0;
    int ti = warp_incl_scan(t);
    if (lane < kThreads / 32) warp_tot[lane] = ti - t;  // exclusive warp offsets
    if (lane == kThreads / 32 - 1 && total) *total = ti;
  }
  __syncthreads();
  int r = warp_tot[wid] + incl - v;
  __syncthreads();
  return r;
}

template <class In>
__global__ void __launch_bounds__(kThreads) tile_sums_kernel(In in, int64_t n, int32_t* tile_sums) {
  __shared__ int s_total;
  const int64_t base = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * kItems;
  int acc = 0;
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    int64_t j = base + i;
    if (j < n) acc += in(j);
  }
  block_excl_scan(acc, &s_total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = s_total;
}

// single CTA: exclusive scan of tile sums in place; writes grand total to total_out[0] if set.
static __global__ void __launch_bounds__(kThreads) scan_tile_sums_kernel(int32_t* tile_sums, int64_t n_tiles,
                                                                 int32_t* total_out) {
  __shared__ int s_total;
  int carry = 0;
  for (int64_t base = 0; base < n_tiles; base += kThreads) {
    int64_t j = base + threadIdx.x;
    int v = (j < n_tiles) ? tile_sums[j] : 0;
    int ex = block_excl_scan(v, &s_total);
    if (j < n_tiles) tile_sums[j] = carry + ex;
    carry += s_total;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

template <class In, class Out>
__global__ void __launch_bounds__(kThreads)
    tile_scan_kernel(In in, int64_t n, const int32_t* tile_sums, Out out) {
  const int64_t base = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * kItems;
  int v[kItems];
  int acc = 0;
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    int64_t j = base + i;
    v[i] = (j < n) ? in(j) : 0;
    acc += v[i];
  }
  int ex = block_excl_scan(acc, nullptr) + tile_sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    int64_t j = base + i;
    if (j < n) out(j, ex, v[i]);
    ex += v[i];
  }
}

// out(j, exclusive_prefix, value) is called for every j < n; total_out[0] = sum.
template <class In, class Out>
inline void exclusive_scan(In in, Out out, int64_t n, int32_t* total_out, void* ws,
                           cudaStream_t st) {
  int32_t* tile_sums = reinterpret_cast<int32_t*>(ws);
  const int64_t nt = num_tiles(n);
  tile_sums_kernel<<<(unsigned)nt, kThreads, 0, st>>>(in, n, tile_sums);
  scan_tile_sums_kernel<<<1, kThreads, 0, st>>>(tile_sums, nt, total_out);
  tile_scan_kernel<<<(unsigned)nt, kThreads, 0, st>>>(in, n, tile_sums, out);
}

}  // namespace scan
}  // namespace er
