// Library-level entry points of liber_b200.so: version, per-thread error text,
// and the host-side Fingerprint64 for arbitrary byte strings.
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "common.cuh"
#include "hash.cuh"

namespace er {
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
static std::atomic<unsigned long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add((unsigned long long)n); }
bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("ER_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}
}  // namespace er

extern "C" int er_abi_version(void) { return ER_B200_ABI_VERSION; }

extern "C" uint64_t er_launch_count(void) { return er::g_launches.load(); }

extern "C" const char* er_last_error(void) { return er::g_last_error.c_str(); }

namespace {
using namespace er::farm;

inline uint64_t f64(const char* p) {
  uint64_t r;
  memcpy(&r, p, 8);
  return r;  // x86-64 / aarch64 little endian
}
inline uint64_t f32(const char* p) {
  uint32_t r;
  memcpy(&r, p, 4);
  return r;
}
struct P128 {
  uint64_t first, second;
};
inline P128 weak32(uint64_t w, uint64_t x, uint64_t y, uint64_t z, uint64_t a, uint64_t b) {
  a += w;
  b = rot(b + a + z, 21);
  uint64_t c = a;
  a += x;
  a += y;
  b += rot(a, 44);
  return {a + z, b + c};
}
inline P128 weak32(const char* s, uint64_t a, uint64_t b) {
  return weak32(f64(s), f64(s + 8), f64(s + 16), f64(s + 24), a, b);
}
}  // namespace

// farmhashna::Hash64 == farmhash::Fingerprint64 == tensorflow::Fingerprint64.
extern "C" uint64_t er_fingerprint64_host(const char* s, size_t len) {
  if (len <= 16) {
    if (len >= 8) {
      uint64_t mul = k2 + len * 2;
      uint64_t a = f64(s) + k2;
      uint64_t b = f64(s + len - 8);
      uint64_t c = rot(b, 37) * mul + a;
      uint64_t d = (rot(a, 25) + b) * mul;
      return hash_len16(c, d, mul);
    }
    if (len >= 4) {
      uint64_t mul = k2 + len * 2;
      uint64_t a = f32(s);
      return hash_len16(len + (a << 3), f32(s + len - 4), mul);
    }
    if (len > 0) {
      uint8_t a = (uint8_t)s[0], b = (uint8_t)s[len >> 1], c = (uint8_t)s[len - 1];
      uint32_t y = (uint32_t)a + ((uint32_t)b << 8);
      uint32_t z = (uint32_t)len + ((uint32_t)c << 2);
      return shift_mix((uint64_t)y * k2 ^ (uint64_t)z * k0) * k2;
    }
    return k2;
  }
  if (len <= 32) {
    uint64_t mul = k2 + len * 2;
    uint64_t a = f64(s) * k1;
    uint64_t b = f64(s + 8);
    uint64_t c = f64(s + len - 8) * mul;
    uint64_t d = f64(s + len - 16) * k2;
    return hash_len16(rot(a + b, 43) + rot(c, 30) + d, a + rot(b + k2, 18) + c, mul);
  }
  if (len <= 64) {
    uint64_t mul = k2 + len * 2;
    uint64_t a = f64(s) * k2;
    uint64_t b = f64(s + 8);
    uint64_t c = f64(s + len - 8) * mul;
    uint64_t d = f64(s + len - 16) * k2;
    uint64_t y = rot(a + b, 43) + rot(c, 30) + d;
    uint64_t z = hash_len16(y, a + rot(b + k2, 18) + c, mul);
    uint64_t e = f64(s + 16) * mul;
    uint64_t f = f64(s + 24);
    uint64_t g = (y + f64(s + len - 32)) * mul;
    uint64_t h = (z + f64(s + len - 24)) * mul;
    return hash_len16(rot(e + f, 43) + rot(g, 30) + h, e + rot(f + a, 18) + g, mul);
  }
  const uint64_t seed = 81;
  uint64_t x = seed;
  uint64_t y = seed * k1 + 113;
  uint64_t z = shift_mix(y * k2 + 113) * k2;
  P128 v = {0, 0}, w = {0, 0};
  x = x * k2 + f64(s);
  const char* end = s + ((len - 1) / 64) * 64;
  const char* last64 = end + ((len - 1) & 63) - 63;
  do {
    x = rot(x + y + v.first + f64(s + 8), 37) * k1;
    y = rot(y + v.second + f64(s + 48), 42) * k1;
    x ^= w.second;
    y += v.first + f64(s + 40);
    z = rot(z + w.first, 33) * k1;
    v = weak32(s, v.second * k1, x + w.first);
    w = weak32(s + 32, z + w.second, y + f64(s + 16));
    uint64_t t = z;
    z = x;
    x = t;
    s += 64;
  } while (s != end);
  uint64_t mul = k1 + ((z & 0xff) << 1);
  s = last64;
  w.first += ((len - 1) & 63);
  v.first += w.first;
  w.first += v.first;
  x = rot(x + y + v.first + f64(s + 8), 37) * mul;
  y = rot(y + v.second + f64(s + 48), 42) * mul;
  x ^= w.second * 9;
  y += v.first * 9 + f64(s + 40);
  z = rot(z + w.first, 33) * mul;
  v = weak32(s, v.second * mul, x + w.first);
  w = weak32(s + 32, z + w.second, y + f64(s + 16));
  {
    uint64_t t = z;
    z = x;
    x = t;
  }
  return hash_len16(hash_len16(v.first, w.first, mul) + shift_mix(y) * k0 + z,
                    hash_len16(v.second, w.second, mul) + x, mul);
}
