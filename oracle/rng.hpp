// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into libsimc_b200.so.
// PARITY UNPINNED: the reference ships no golden vectors and cannot be built here
// (no Fortran compiler); see DESIGN.md "Oracle".
//
// Random-number back-ends for the CPU restatement of SIMC's event loop.
//   * Ranlux : the reference's stream -- RANLUX luxury level 3 (cern/ranlux.f:3-312)
//              served through the 1000-deep buffer of grnd() (call_ranlux.f:58-74).
//   * Philox : the counter-based stream the B200 path uses, keyed (seed, try, draw).
//   * Tape   : pre-drawn uniforms replayed in call order.
#pragma once
#include <cstdint>
#include <vector>
#include <stdexcept>

namespace simc_oracle {

// ---- Philox4x32-10 (Salmon et al., SC'11), the published algorithm -------------
struct Philox4x32 {
  static inline void round(uint32_t c[4], uint32_t k0, uint32_t k1) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
  }
  static inline void block(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
      round(c, k0, k1);
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
  }
};

// One uniform = 52 random bits, u = (k + 1/2) * 2^-52, so 0 < u < 1 strictly
// (the reference's RANLUX also excludes both end points, cern/ranlux.f:112-129).
// Draw d of try t uses block d/2 of counter (d/2, stream, t_lo, t_hi), words
// (0,1) for even d and (2,3) for odd d.
struct Rng {
  enum Mode { PHILOX, RANLUX, TAPE } mode = PHILOX;
  // philox
  uint32_t key[2] = {0, 0};
  uint32_t ctr_try[2] = {0, 0};
  uint32_t stream = 0;
  uint32_t draw = 0;
  uint32_t cache[4];
  uint32_t cached_block = 0xFFFFFFFFu;
  // tape
  const double* tape = nullptr; size_t tape_len = 0; size_t tape_pos = 0;
  // ranlux
  struct RanluxState* rl = nullptr;

  void seed_philox(uint64_t seed, uint64_t try_index, uint32_t stream_id = 0, uint32_t first_draw = 0) {
    mode = PHILOX;
    key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
    ctr_try[0] = (uint32_t)try_index; ctr_try[1] = (uint32_t)(try_index >> 32);
    stream = stream_id; draw = first_draw; cached_block = 0xFFFFFFFFu;
  }
  void set_tape(const double* t, size_t n) { mode = TAPE; tape = t; tape_len = n; tape_pos = 0; draw = 0; }

  double grnd();
};

// ---- RANLUX, luxury level 3, restated from cern/ranlux.f ---------------------
struct RanluxState {
  static constexpr int    maxlev = 4;
  static constexpr int    igiga = 1000000000, jsdflt = 314159265, itwo24 = 1 << 24, icons = 2147483563;
  // Implicitly typed REALs of ranlux.f are 8 bytes under the reference's
  // -fdefault-real-8 (Makefile:63); every value is a multiple of 2^-24 (2^-48 in the
  // small-number padding branch), so double arithmetic is exact here.
  double seeds[25];   // 1-based
  int    next[25];
  int    i24 = 24, j24 = 10, in24 = 0, kount = 0, mkount = 0;
  int    luxlev = 3, nskip = 0, inseed = 0;
  double carry = 0., twom24 = 1., twom12 = 0.;
  bool   notyet = true;
  // call_ranlux.f:58-74 buffer
  double rvec[1001]; int latest = 0;

  void rluxgo(int lux, int ins, int k1, int k2);   // cern/ranlux.f:209-283
  void ranlux(double* out, int lenv);              // cern/ranlux.f:76-136 (rvec is real*8 in grnd)
  double grnd() {
    if (latest <= 0 || latest >= 1000) { ranlux(rvec + 1, 1000); latest = 1; }
    return rvec[latest++];
  }
};

inline double Rng::grnd() {
  if (mode == PHILOX) {
    const uint32_t blk = draw >> 1;
    if (blk != cached_block) {
      const uint32_t c[4] = {blk, stream, ctr_try[0], ctr_try[1]};
      Philox4x32::block(c, key, cache);
      cached_block = blk;
    }
    const uint32_t lo = cache[(draw & 1) * 2], hi = cache[(draw & 1) * 2 + 1];
    ++draw;
    const uint64_t k = (((uint64_t)hi << 32) | lo) >> 12;
    return ((double)k + 0.5) * (1.0 / 4503599627370496.0);
  } else if (mode == TAPE) {
    if (tape_pos >= tape_len) throw std::runtime_error("oracle: random tape exhausted");
    ++draw;
    return tape[tape_pos++];
  } else {
    ++draw;
    return rl->grnd();
  }
}

}  // namespace simc_oracle
