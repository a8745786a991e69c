// Counter-based per-event random stream (replaces the sequential RANLUX of
// call_ranlux.f / cern/ranlux.f).  Philox4x32-10; draw d of try t = 52-bit uniform built
// from words (0,1) [d even] or (2,3) [d odd] of block (d/2, stream, t_lo, t_hi) under key
// (seed_lo, seed_hi):  u = (k + 1/2) * 2^-52, 0 < u < 1.
#pragma once
#include <stdint.h>

namespace simc {

// The ten round keys of the run's seed (k0 + r*0x9E3779B9, k1 + r*0xBB67AE85) sit in constant memory: a round is
// then two wide multiplies and two three-input XORs with a constant-bank operand, without the two per-thread key
// additions (a third of the generator's instructions).  The host writes them before the first launch with a new
// seed (ensure_rng_key in kernels.cu).  One copy per translation unit (strict / fast).
static __constant__ uint32_t c_philox_rk[20];

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t& r0, uint32_t& r1,
                                              uint32_t& r2, uint32_t& r3) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ c_philox_rk[2 * r], n2 = hi0 ^ c3 ^ c_philox_rk[2 * r + 1];
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
  }
  r0 = c0; r1 = c1; r2 = c2; r3 = c3;
}
__device__ __forceinline__ double philox_to_unit(uint32_t lo, uint32_t hi) {
  const unsigned long long k = (((unsigned long long)hi << 32) | lo) >> 12;
  return ((double)k + 0.5) * (1.0 / 4503599627370496.0);
}
// 2u - 1 for the same draw, as gauss1 forms it (gauss1.f: v = 2.*grnd() - 1.): with k the 52-bit integer,
// 2u - 1 = k*2^-51 + 2^-52 - 1.  The double with mantissa k in [2,4) is 2 + k*2^-51; minus 3 is exact (a multiple of
// 2^-51 below 1 in size), and adding 2^-52 rounds the same real number the reference rounds: bit-identical to
// 2.0 * philox_to_unit() - 1.0 with two additions instead of a conversion, an addition, two multiplies and a subtraction.
__device__ __forceinline__ double philox_to_pm1(uint32_t lo, uint32_t hi) {
  const unsigned long long k = (((unsigned long long)hi << 32) | lo) >> 12;
  const double d = __longlong_as_double((long long)(k | 0x4000000000000000ULL));
  return (d - 3.0) + 2.220446049250313e-16;
}
// A block gives two draws: the even draw d of a try is words (0,1) of block d/2, the odd draw d+1 words (2,3).
// The generator keeps the upper half of the block of its last even draw (h2, h3), so INVARIANT: whenever `draw` is
// odd, (h2, h3) are words (2,3) of block draw/2 -- an odd draw costs no Philox evaluation, and gauss1 (which takes its
// uniforms in pairs) starts an odd-aligned pair without recomputing the block the pair begins in.
struct PhiloxHalf { uint32_t lo0, lo1, hi0, hi1; };
static __device__ __noinline__ PhiloxHalf philox_block(uint32_t t0, uint32_t t1, uint32_t stream, uint32_t block) {
  PhiloxHalf r;
  philox4x32_10(block, stream, t0, t1, r.lo0, r.lo1, r.hi0, r.hi1);
  return r;
}

struct DevRng {
  uint32_t t0, t1, stream;
  uint32_t draw;
  uint32_t h2, h3;

  // (the seed itself is in c_philox_rk)
  __device__ __forceinline__ void init(unsigned long long try_index, uint32_t stream_id, uint32_t first_draw) {
    t0 = (uint32_t)try_index; t1 = (uint32_t)(try_index >> 32);
    stream = stream_id; draw = first_draw;
    h2 = 0u; h3 = 0u;
    if (first_draw & 1u) {                       // resuming in the middle of a block
      const PhiloxHalf r = philox_block(t0, t1, stream, first_draw >> 1);
      h2 = r.hi0; h3 = r.hi1;
    }
  }
  __device__ __forceinline__ double uniform() {
    double u;
    if (draw & 1u) {
      u = philox_to_unit(h2, h3);
    } else {
      const PhiloxHalf r = philox_block(t0, t1, stream, draw >> 1);
      u = philox_to_unit(r.lo0, r.lo1);
      h2 = r.hi0; h3 = r.hi1;
    }
    ++draw;
    return u;
  }
};

}  // namespace simc
