// Stable LSD radix sort of (row, lookup-position) pairs: the deterministic core
// of the gradient dedup.  It replaces TF's Unique + UnsortedSegmentSum
// (_deduplicate_indexed_slices, applied before compat/adam_s.py:185-213 and
// every TF sparse apply; and array_ops.unique at
// compat/feature_column/feature_column.py:263).
//
// Keys are arena rows (< 2^32 per GPU: 180 GB / 16 B minimum row), values the
// lookup positions; dropped lookups (row < 0) get the sentinel key `n_rows`
// and therefore sort behind every valid row.  8-bit digits; only
// ceil(bits(n_rows)/8) passes are run (3 for a 10M-row arena).  Working set at
// the benchmark shapes (<= 1M pairs x 8 B x 2 buffers) is L2 resident on B200,
// so the passes are latency-, not HBM-bound; each pass is hist -> scan -> scatter.
#pragma once
#include "common.cuh"
#include "scan.cuh"

namespace er {
namespace rsort {

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kItems = 4;
constexpr int kTile = kThreads * kItems;  // 1024 pairs per CTA

inline int64_t num_tiles(int64_t n) { return n > 0 ? ceil_div(n, kTile) : 1; }
inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct Workspace {
  uint32_t* keys_tmp;
  uint32_t* vals_tmp;
  int32_t* hist;   // [kRadix][n_tiles], scanned in place
  void* scan_ws;
};

inline size_t workspace_bytes(int64_t n) {
  const int64_t nt = num_tiles(n);
  return align256((size_t)n * 4) * 2 + align256((size_t)kRadix * nt * 4) +
         align256(scan::workspace_bytes((int64_t)kRadix * nt)) + 256;
}

inline Workspace carve(void* ws, int64_t n) {
  const int64_t nt = num_tiles(n);
  char* p = reinterpret_cast<char*>(align256(reinterpret_cast<size_t>(ws)));
  Workspace w;
  w.keys_tmp = reinterpret_cast<uint32_t*>(p);
  p += align256((size_t)n * 4);
  w.vals_tmp = reinterpret_cast<uint32_t*>(p);
  p += align256((size_t)n * 4);
  w.hist = reinterpret_cast<int32_t*>(p);
  p += align256((size_t)kRadix * nt * 4);
  w.scan_ws = p;
  return w;
}

inline int num_passes(int64_t n_rows) {
  int bits = 1;
  while (bits < 32 && (1ULL << bits) <= (uint64_t)n_rows) ++bits;  // sentinel key == n_rows
  return (bits + kRadixBits - 1) / kRadixBits;
}

static __global__ void __launch_bounds__(256)
    init_pairs_kernel(const int64_t* __restrict__ rows, int64_t cap, const int32_t* __restrict__ n_dev,
                      uint32_t sentinel, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int64_t n = n_dev ? (int64_t)(*n_dev < cap ? *n_dev : cap) : cap;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = (i < n) ? rows[i] : -1;
    keys[i] = (r < 0 || r >= (int64_t)sentinel) ? sentinel : (uint32_t)r;
    vals[i] = (uint32_t)i;
  }
}

static __global__ void __launch_bounds__(kThreads)
    hist_kernel(const uint32_t* __restrict__ keys, int64_t n, int shift, int32_t* __restrict__ hist,
                int64_t n_tiles) {
  __shared__ int s_hist[kRadix];
  s_hist[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kTile;
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    int64_t idx = base + i * kThreads + threadIdx.x;
    if (idx < n) atomicAdd(&s_hist[(keys[idx] >> shift) & (kRadix - 1)], 1);
  }
  __syncthreads();
  hist[(int64_t)threadIdx.x * n_tiles + blockIdx.x] = s_hist[threadIdx.x];
}

struct HistIn {
  const int32_t* h;
  __device__ int operator()(int64_t j) const { return h[j]; }
};
struct HistOut {
  int32_t* h;
  __device__ void operator()(int64_t j, int ex, int) const { h[j] = ex; }
};

static __global__ void __launch_bounds__(kThreads)
    scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                   uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int64_t n,
                   int shift, const int32_t* __restrict__ offsets, int64_t n_tiles) {
  __shared__ int s_warp_hist[kWarps][kRadix];
  for (int i = threadIdx.x; i < kWarps * kRadix; i += kThreads) (&s_warp_hist[0][0])[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int64_t base = (int64_t)blockIdx.x * kTile + (int64_t)w * (32 * kItems);
  uint32_t key[kItems], val[kItems];
  int rank[kItems];
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    const int64_t idx = base + i * 32 + lane;
    const bool valid = idx < n;
    key[i] = valid ? keys_in[idx] : 0u;
    val[i] = valid ? vals_in[idx] : 0u;
    const int d = (int)((key[i] >> shift) & (kRadix - 1));
    // lanes past the end never match a real digit
    const unsigned peers = __match_any_sync(0xffffffffu, valid ? d : (kRadix + lane));
    const int r = __popc(peers & lt_mask);
    int prev = 0;
    if (valid) prev = s_warp_hist[w][d];
    __syncwarp();
    if (valid && r == 0) s_warp_hist[w][d] = prev + __popc(peers);
    __syncwarp();
    rank[i] = prev + r;
  }
  __syncthreads();
  {
    // digit threadIdx.x: exclusive prefix over warps + this tile's global offset
    int run = offsets[(int64_t)threadIdx.x * n_tiles + blockIdx.x];
#pragma unroll
    for (int ww = 0; ww < kWarps; ++ww) {
      int t = s_warp_hist[ww][threadIdx.x];
      s_warp_hist[ww][threadIdx.x] = run;
      run += t;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    const int64_t idx = base + i * 32 + lane;
    if (idx < n) {
      const int d = (int)((key[i] >> shift) & (kRadix - 1));
      const int pos = s_warp_hist[w][d] + rank[i];
      keys_out[pos] = key[i];
      vals_out[pos] = val[i];
    }
  }
}

// Sort `cap` pairs; result lands in keys_out / vals_out.  hist_kernel reads tiles in a
// strided (coalesced) order, scatter_kernel in a warp-blocked order: both cover the same
// [tile*kTile, (tile+1)*kTile) range, so the per-tile digit counts agree.
inline void sort_rows(const int64_t* rows, int64_t cap, const int32_t* n_dev, int64_t n_rows,
                      uint32_t* keys_out, uint32_t* vals_out, void* ws, cudaStream_t st) {
  Workspace w = carve(ws, cap);
  const int passes = num_passes(n_rows);
  const int64_t nt = num_tiles(cap);
  uint32_t* ka = (passes % 2 == 0) ? keys_out : w.keys_tmp;
  uint32_t* va = (passes % 2 == 0) ? vals_out : w.vals_tmp;
  uint32_t* kb = (passes % 2 == 0) ? w.keys_tmp : keys_out;
  uint32_t* vb = (passes % 2 == 0) ? w.vals_tmp : vals_out;
  init_pairs_kernel<<<grid_for(cap, 256, 8), 256, 0, st>>>(rows, cap, n_dev, (uint32_t)n_rows, ka, va);
  count_launches(1);
  for (int p = 0; p < passes; ++p) {
    const int shift = p * kRadixBits;
    hist_kernel<<<(unsigned)nt, kThreads, 0, st>>>(ka, cap, shift, w.hist, nt);
    scan::exclusive_scan(HistIn{w.hist}, HistOut{w.hist}, (int64_t)kRadix * nt, nullptr, w.scan_ws, st);
    scatter_kernel<<<(unsigned)nt, kThreads, 0, st>>>(ka, va, kb, vb, cap, shift, w.hist, nt);
    count_launches(5);
    uint32_t* t = ka; ka = kb; kb = t;
    t = va; va = vb; vb = t;
  }
}

}  // namespace rsort
}  // namespace er
