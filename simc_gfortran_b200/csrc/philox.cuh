// Counter-based per-event random stream (replaces the sequential RANLUX of
// call_ranlux.f / cern/ranlux.f).  Philox4x32-10; draw d of try t = 52-bit uniform built
// from words (0,1) [d even] or (2,3) [d odd] of block (d/2, stream, t_lo, t_hi) under key
// (seed_lo, seed_hi):  u = (k + 1/2) * 2^-52, 0 < u < 1.
#pragma once
#include <stdint.h>

namespace simc {

__device__ __forceinline__ void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2,
                                              uint32_t c3, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  r0 = c0; r1 = c1; r2 = c2; r3 = c3;
}
__device__ __forceinline__ double philox_to_unit(uint32_t lo, uint32_t hi) {
  const unsigned long long k = (((unsigned long long)hi << 32) | lo) >> 12;
  return ((double)k + 0.5) * (1.0 / 4503599627370496.0);
}
// Draw `draw` of the stream: a pure function of its arguments (registers in, register out), so callers
// keep their generator state in registers across the call.
static __device__ __noinline__ double philox_uniform(uint32_t k0, uint32_t k1, uint32_t t0, uint32_t t1, uint32_t stream,
                                              uint32_t draw) {
  uint32_t r0, r1, r2, r3;
  philox4x32_10(k0, k1, draw >> 1, stream, t0, t1, r0, r1, r2, r3);
  return (draw & 1u) ? philox_to_unit(r2, r3) : philox_to_unit(r0, r1);
}

struct DevRng {
  uint32_t k0, k1, t0, t1, stream;
  uint32_t draw;

  __device__ __forceinline__ void init(unsigned long long seed, unsigned long long try_index, uint32_t stream_id,
                                       uint32_t first_draw) {
    k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32);
    t0 = (uint32_t)try_index; t1 = (uint32_t)(try_index >> 32);
    stream = stream_id; draw = first_draw;
  }
  __device__ __forceinline__ double uniform() {
    const double u = philox_uniform(k0, k1, t0, t1, stream, draw);
    ++draw;
    return u;
  }
};

}  // namespace simc
