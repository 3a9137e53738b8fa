// Fingerprint64 (FarmHash farmhashna::Hash64, the hash behind TensorFlow's
// StringToHashBucketFast) for the strings EasyRec feeds it on the hot path:
// tf.as_string(int64) decimal text, 1..20 bytes
// (reference call sites: feature_column_v2.py:3915-3921, input/input.py:356-376).
//
// Device side: the decimal text never touches memory -- it is built in three
// 64-bit registers (little-endian byte order, byte i of the string in bits
// 8*(i%8) of word i/8) and the FarmHash fetches become funnel shifts.
// Host side: er_fingerprint64_host() handles arbitrary byte strings.
#pragma once
#include <stdint.h>

namespace er {
namespace farm {

constexpr uint64_t k0 = 0xc3a5c85c97cb3127ULL;
constexpr uint64_t k1 = 0xb492b66fbe98f273ULL;
constexpr uint64_t k2 = 0x9ae16a3b2f90404fULL;

__host__ __device__ __forceinline__ uint64_t rot(uint64_t v, int s) {
  return (v >> s) | (v << (64 - s));
}
__host__ __device__ __forceinline__ uint64_t shift_mix(uint64_t v) { return v ^ (v >> 47); }
__host__ __device__ __forceinline__ uint64_t hash_len16(uint64_t u, uint64_t v, uint64_t mul) {
  uint64_t a = (u ^ v) * mul;
  a ^= (a >> 47);
  uint64_t b = (v ^ a) * mul;
  b ^= (b >> 47);
  return b * mul;
}

struct Dec {
  uint64_t w0, w1, w2;
  int len;
};

// tf.as_string(int64): "%lld" text.
__device__ __forceinline__ Dec to_decimal(int64_t v) {
  const bool neg = v < 0;
  uint64_t mag = neg ? (0ULL - (uint64_t)v) : (uint64_t)v;
  int nd = 1;
  {
    uint64_t p = 10ULL;
#pragma unroll
    for (int i = 1; i < 20; ++i) {
      nd += (mag >= p) ? 1 : 0;
      p = (i < 19) ? p * 10ULL : p;
    }
    // the 20th power of ten overflows: 10^19 is the largest threshold that matters
    if (nd > 20) nd = 20;
  }
  Dec d;
  d.w0 = neg ? (uint64_t)'-' : 0ULL;
  d.w1 = 0ULL;
  d.w2 = 0ULL;
  d.len = nd + (neg ? 1 : 0);
  for (int p = d.len - 1; p >= (neg ? 1 : 0); --p) {
    uint64_t q = mag / 10ULL;
    uint64_t c = (uint64_t)('0' + (uint32_t)(mag - q * 10ULL));
    mag = q;
    uint64_t c64 = c << ((p & 7) * 8);
    if (p < 8)
      d.w0 |= c64;
    else if (p < 16)
      d.w1 |= c64;
    else
      d.w2 |= c64;
  }
  return d;
}

__device__ __forceinline__ uint64_t fetch64(const Dec& d, int off) {
  const int k = off >> 3;
  const int r = (off & 7) * 8;
  uint64_t lo = (k == 0) ? d.w0 : ((k == 1) ? d.w1 : d.w2);
  uint64_t hi = (k == 0) ? d.w1 : ((k == 1) ? d.w2 : 0ULL);
  return r ? ((lo >> r) | (hi << (64 - r))) : lo;
}
__device__ __forceinline__ uint64_t fetch32(const Dec& d, int off) {
  return fetch64(d, off) & 0xffffffffULL;
}
__device__ __forceinline__ uint32_t byte_at(const Dec& d, int off) {
  return (uint32_t)(fetch64(d, off) & 0xffULL);
}

// farmhashna::Hash64 restricted to len <= 32 (decimal int64 text is <= 20 bytes).
__device__ __forceinline__ uint64_t fingerprint64_dec(const Dec& d) {
  const uint64_t len = (uint64_t)d.len;
  if (d.len <= 16) {
    if (d.len >= 8) {
      uint64_t mul = k2 + len * 2;
      uint64_t a = fetch64(d, 0) + k2;
      uint64_t b = fetch64(d, d.len - 8);
      uint64_t c = rot(b, 37) * mul + a;
      uint64_t e = (rot(a, 25) + b) * mul;
      return hash_len16(c, e, mul);
    }
    if (d.len >= 4) {
      uint64_t mul = k2 + len * 2;
      uint64_t a = fetch32(d, 0);
      return hash_len16(len + (a << 3), fetch32(d, d.len - 4), mul);
    }
    if (d.len > 0) {
      uint32_t a = byte_at(d, 0);
      uint32_t b = byte_at(d, d.len >> 1);
      uint32_t c = byte_at(d, d.len - 1);
      uint32_t y = a + (b << 8);
      uint32_t z = (uint32_t)d.len + (c << 2);
      return shift_mix((uint64_t)y * k2 ^ (uint64_t)z * k0) * k2;
    }
    return k2;
  }
  // 17..32
  uint64_t mul = k2 + len * 2;
  uint64_t a = fetch64(d, 0) * k1;
  uint64_t b = fetch64(d, 8);
  uint64_t c = fetch64(d, d.len - 8) * mul;
  uint64_t e = fetch64(d, d.len - 16) * k2;
  return hash_len16(rot(a + b, 43) + rot(c, 30) + e, a + rot(b + k2, 18) + c, mul);
}

}  // namespace farm
}  // namespace er
