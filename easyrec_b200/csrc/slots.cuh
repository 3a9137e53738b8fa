// Per-CTA shared-memory view of the er_slot_t plan.
//
// Every sparse kernel maps a segment number to its slot (which table rule, which output
// matrix/column).  Profiling the first version showed those lookups -- a binary search plus a
// 48-byte descriptor read per segment -- made K2/K7 instruction-issue bound, not HBM bound.  Here
// each CTA stages a 16-byte SlotLite per slot in shared memory once, and when all slots have the
// same number of segments laid out back to back (the packed feature-major batch: seg = f*B + b)
// the slot is `seg / n_seg` by a multiply-shift, so the per-lookup cost is one LDS.128.
#pragma once
#include "common.cuh"

namespace er {

struct SlotLite {
  int32_t seg_begin;
  int32_t out_stride;
  int32_t out_col;
  int32_t misc;  // out_buf | combiner << 8
};

struct FastDiv {  // exact unsigned 32-bit division by a runtime constant (Granlund-Montgomery)
  uint32_t m, s1, s2;
};

__device__ __forceinline__ FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  if (d <= 1) {
    f.m = 0; f.s1 = 0; f.s2 = 0;
    return f;
  }
  const uint32_t ell = 32 - __clz(d - 1);  // ceil(log2 d)
  f.m = (uint32_t)((((uint64_t)1 << 32) * (((uint64_t)1 << ell) - d)) / d + 1);
  f.s1 = 1;
  f.s2 = ell - 1;
  return f;
}
__device__ __forceinline__ uint32_t fastdiv(uint32_t n, const FastDiv& f) {
  const uint32_t t = __umulhi(f.m, n);
  return (t + ((n - t) >> f.s1)) >> f.s2;
}

struct SlotView {
  const SlotLite* tab;  // shared memory, n_slots entries
  int n_slots;
  int uniform_nseg;     // > 0: slot f owns segments [f*uniform_nseg, (f+1)*uniform_nseg)
  FastDiv div;
};

// Shared-memory bytes a kernel must reserve (dynamic smem) for the slot table.
__host__ __device__ inline size_t slot_smem_bytes(int n_slots) { return (size_t)n_slots * sizeof(SlotLite) + 16; }

// All threads of the CTA call this once; `smem` is 16-byte aligned dynamic shared memory.
__device__ __forceinline__ SlotView load_slots(void* smem, const er_slot_t* __restrict__ slots,
                                               int n_slots) {
  SlotLite* tab = reinterpret_cast<SlotLite*>(smem);

  const int nseg0 = slots[0].n_seg;
  int ok = 1;
  for (int i = threadIdx.x; i < n_slots; i += blockDim.x) {
    const er_slot_t s = slots[i];
    SlotLite l;
    l.seg_begin = s.seg_begin;
    l.out_stride = s.out_stride;
    l.out_col = s.out_col;
    l.misc = (s.out_buf & 0xff) | (s.combiner << 8);   // combiner incl. the ER_COMBINER_UNIT_WEIGHTS flag
    tab[i] = l;
    if (s.n_seg != nseg0 || s.seg_begin != i * nseg0) ok = 0;
  }
  ok = __syncthreads_and(ok);

  SlotView v;
  v.tab = tab;
  v.n_slots = n_slots;
  v.uniform_nseg = ok ? nseg0 : 0;
  v.div = make_fastdiv((uint32_t)(nseg0 > 0 ? nseg0 : 1));
  return v;
}

__device__ __forceinline__ int slot_of(const SlotView& v, int32_t s) {
  if (v.uniform_nseg > 0) return (int)fastdiv((uint32_t)s, v.div);
  int lo = 0, hi = v.n_slots;  // invariant: tab[lo].seg_begin <= s < tab[hi].seg_begin
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (v.tab[mid].seg_begin <= s)
      lo = mid;
    else
      hi = mid;
  }
  return lo;
}

__device__ __forceinline__ int slot_comb(const SlotLite& l) { return (l.misc >> 8) & 0xf; }
__device__ __forceinline__ bool slot_unit_weights(const SlotLite& l) { return (l.misc >> 12) & 1; }

__device__ __forceinline__ SlotLite slot_lite(const SlotView& v, int f) {
  const int4 q = *reinterpret_cast<const int4*>(v.tab + f);
  SlotLite l;
  l.seg_begin = q.x;
  l.out_stride = q.y;
  l.out_col = q.z;
  l.misc = q.w;
  return l;
}

}  // namespace er
