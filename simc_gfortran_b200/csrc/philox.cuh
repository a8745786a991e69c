// Counter-based per-event random stream (replaces the sequential RANLUX of
// call_ranlux.f / cern/ranlux.f).  Philox4x32-10; draw d of try t = 52-bit uniform built
// from words (0,1) [d even] or (2,3) [d odd] of block (d/2, stream, t_lo, t_hi) under key
// (seed_lo, seed_hi):  u = (k + 1/2) * 2^-52, 0 < u < 1.
#pragma once
#include <stdint.h>

namespace simc {

struct DevRng {
  uint32_t k0, k1, t0, t1, stream;
  uint32_t draw;
  uint32_t w2, w3;        // second half of the cached block
  uint32_t cached;        // block index whose second half is cached (0xffffffff = none)

  __device__ __forceinline__ void init(unsigned long long seed, unsigned long long try_index, uint32_t stream_id,
                                       uint32_t first_draw) {
    k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32);
    t0 = (uint32_t)try_index; t1 = (uint32_t)(try_index >> 32);
    stream = stream_id; draw = first_draw; cached = 0xffffffffu; w2 = w3 = 0;
  }

  __device__ __forceinline__ void block(uint32_t b, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) const {
    uint32_t c0 = b, c1 = stream, c2 = t0, c3 = t1, ka = k0, kb = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      const uint32_t n0 = hi1 ^ c1 ^ ka, n2 = hi0 ^ c3 ^ kb;
      c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
      ka += 0x9E3779B9u; kb += 0xBB67AE85u;
    }
    r0 = c0; r1 = c1; r2 = c2; r3 = c3;
  }

  __device__ __noinline__ double uniform() {
    const uint32_t b = draw >> 1;
    uint32_t lo, hi;
    if ((draw & 1u) && cached == b) {
      lo = w2; hi = w3;
    } else {
      uint32_t r0, r1, r2, r3;
      block(b, r0, r1, r2, r3);
      if (draw & 1u) { lo = r2; hi = r3; }
      else { lo = r0; hi = r1; w2 = r2; w3 = r3; cached = b; }
    }
    ++draw;
    const unsigned long long k = (((unsigned long long)hi << 32) | lo) >> 12;
    return ((double)k + 0.5) * (1.0 / 4503599627370496.0);
  }
};

}  // namespace simc
