// Device side of single-arm transport + reconstruction: one thread per event interprets
// the arm program (arm_program.h).  Replaces, per event, mc_hms/mc_shms/... with their hut
// and recon routines and the shared primitives project/musc/musc_ext/rotate_*/transp/lfit.
//
// SIMC_STRICT=1 (compiled with -fmad=false): every product and sum is formed in the
// reference's order with separate multiply and add, so the COSY sums are bit-identical to a
// no-FMA x86-64 build; only libm calls (log, log10, acos, sin, cos) can differ in the last ulp.
// SIMC_STRICT=0: in the COSY polynomials only, the monomial products are re-associated (suffix
// product per group) and explicit fused multiply-adds are used; results agree to ~1e-15 relative.
// Everything else is the same no-FMA arithmetic (the whole library is built with -fmad=false).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "arm_program.h"
#include "philox.cuh"
#include "target.cuh"

#ifndef SIMC_STRICT
#define SIMC_STRICT 1
#endif
#ifndef SIMC_VARIANT_NS
#define SIMC_VARIANT_NS strict
#endif

namespace simc {
namespace SIMC_VARIANT_NS {

constexpr int kBlock = 128;               // threads per CTA for the event kernels
static_assert(kBlock == kArmBlockThreads, "the record offsets are compiled for this CTA size");
// shared power table of a COSY map: [kPolyEntries][kBlock] doubles, dynamic shared memory (49 KB).  The hut has
// no map before its reconstruction, so the same columns hold each thread's queue of Gaussians there
// (GaussQueue below), which needs one row more: kPowRows rows of kBlock doubles.
#ifndef SIMC_GAUSS_Q
#define SIMC_GAUSS_Q 16
#endif
constexpr int kGaussQ = SIMC_GAUSS_Q;            // Gaussians a thread draws ahead (16: the queue of a CTA is 40 KB, five CTAs per SM)
constexpr int kPowRows = kPolyEntries + 1;
static_assert(2 * kGaussQ + (kGaussQ + 1) / 2 <= kPowRows, "the Gaussian queue must fit the thread's column");
constexpr int kPowDoubles = kPowRows * kBlock;
constexpr size_t kPowBytes = sizeof(double) * kPowDoubles;
// dynamic shared memory of a kernel that only needs the queue (the hut in front of a compiled reconstruction map)
constexpr size_t kQueueBytes = sizeof(double) * (size_t)(2 * kGaussQ + (kGaussQ + 1) / 2) * kBlock;

struct ArmDev {                           // lives in global memory, read through warp-uniform loads
  ArmTablesDev tab;
  ArmOp ops[kMaxArmOps];
};

struct TrackDev {
  double xs, ys, dxdzs, dydzs, dpps;      // COMMON /track/ (spectrometers.inc:47-60)
  double p, m2, pathlen;
  double decdist, mh2_final, ctau;        // simulate.inc:183, :92, :153
  double mc1, mbeta2;                     // cached Es/p/beta and beta^2 of musc (valid while p,m2 unchanged)
  double p_spec;                          // the arm's central momentum (argument of mc_hms / mc_hms_coll)
  bool dflag;
};

struct ArmResult {
  double x_fp, dx_fp, y_fp, dy_fp;
  double dpp_rec, dth_rec, dph_rec, y_rec;
  double resmult;
  int stop_code;
  bool reached_hut;
  bool ok;
};

// constants.inc:17-42 (decay kinematics)
#define SIMC_MPI 139.57018
#define SIMC_MK 493.677
#define SIMC_MMU 105.6583755
#define SIMC_PI 3.141592653589793

// gauss1.f:1-30, warp-synchronous: the lanes that called together iterate together, so they leave
// the function converged.  The polar rejection (s > 1) costs two uniforms per attempt and is
// looped on its own; log/sqrt/div run once for all lanes, and only a lane whose |g| exceeds
// nsigmax (never for 99, 0.3 % for 3) goes round again -- the same draws, in the same order.
// Everything goes in and out by value (registers), so the caller's generator never touches memory.
struct GaussOut { double g; uint32_t draw, h2, h3; };
__device__ __noinline__ GaussOut gauss1v(uint32_t t0, uint32_t t1, uint32_t stream, uint32_t draw, uint32_t h2, uint32_t h3,
                                         double nsigmax) {
  const unsigned mask = __activemask();
  // a pair (u1,u2) is one Philox block when `draw` is even; when it is odd the pair starts with the half block the
  // generator holds (DevRng's invariant) and ends in the lower half of the next block, whose upper half is kept
  double g = 0.0;
  bool need = true;
  while (__any_sync(mask, need)) {
    double v1 = 0.0, s = 1.0;
    bool open = need;
    while (__any_sync(mask, open)) {
      if (open) {
        uint32_t r0, r1, r2, r3, w0, w1, w2, w3;          // (w0,w1) = words of u1, (w2,w3) = words of u2
        const uint32_t b = draw >> 1;
        if (!(draw & 1u)) {
          philox4x32_10(b, stream, t0, t1, r0, r1, r2, r3);
          w0 = r0; w1 = r1; w2 = r2; w3 = r3;
        } else {
          w0 = h2; w1 = h3;
          philox4x32_10(b + 1u, stream, t0, t1, r0, r1, r2, r3);
          w2 = r0; w3 = r1;
          h2 = r2; h3 = r3;
        }
        draw += 2u;
        v1 = philox_to_pm1(w0, w1);                        // 2.*grnd() - 1., bit for bit (philox.cuh)
        const double v2 = philox_to_pm1(w2, w3);
        s = v1 * v1 + v2 * v2;
        open = (s > 1. || s == 0.);
      }
    }
    if (need) {
      g = v1 * sqrt(-2. * m::log(s) / s);
      need = fabs(g) > nsigmax;
    }
  }
  GaussOut o;
  o.g = g; o.draw = draw; o.h2 = h2; o.h3 = h3;
  return o;
}
__device__ __forceinline__ double gauss1(DevRng& r, double nsigmax) {
  const GaussOut o = gauss1v(r.t0, r.t1, r.stream, r.draw, r.h2, r.h3, nsigmax);
  r.draw = o.draw; r.h2 = o.h2; r.h3 = o.h3;
  return o.g;
}

// Two consecutive gauss1(99.) calls -- every Gaussian of the hut comes in such pairs (musc, musc_ext, the two
// smearings of a chamber plane).  A lane that has accepted its first pair of uniforms goes straight on to draw for
// the second while slower lanes still work on their first, so the warp makes max(n1 + n2) trips through the
// rejection loop instead of max(n1) + max(n2); each lane consumes exactly the draws of the two sequential calls.
// |g| <= 12 for the smallest s the 52-bit uniforms can produce, so the nsigmax = 99 test can never fire.
struct Gauss2Out { double g1, g2; uint32_t draw, h2, h3; };
__device__ __noinline__ Gauss2Out gauss2v(uint32_t t0, uint32_t t1, uint32_t stream, uint32_t draw, uint32_t h2, uint32_t h3) {
  const unsigned mask = __activemask();
  double va = 0.0, sa = 1.0, vb = 0.0, sb = 1.0;
  int k = 0;
  while (__any_sync(mask, k < 2)) {
    if (k < 2) {
      uint32_t r0, r1, r2, r3, w0, w1, w2, w3;
      const uint32_t b = draw >> 1;
      if (!(draw & 1u)) {
        philox4x32_10(b, stream, t0, t1, r0, r1, r2, r3);
        w0 = r0; w1 = r1; w2 = r2; w3 = r3;
      } else {
        w0 = h2; w1 = h3;
        philox4x32_10(b + 1u, stream, t0, t1, r0, r1, r2, r3);
        w2 = r0; w3 = r1;
        h2 = r2; h3 = r3;
      }
      draw += 2u;
      const double v1 = philox_to_pm1(w0, w1);
      const double v2 = philox_to_pm1(w2, w3);
      const double s = v1 * v1 + v2 * v2;
      if (!(s > 1. || s == 0.)) {
        if (k == 0) { va = v1; sa = s; } else { vb = v1; sb = s; }
        ++k;
      }
    }
  }
  Gauss2Out o;
  o.g1 = va * sqrt(-2. * m::log(sa) / sa);
  o.g2 = vb * sqrt(-2. * m::log(sb) / sb);
  o.draw = draw; o.h2 = h2; o.h3 = h3;
  return o;
}
__device__ __forceinline__ void gauss2(DevRng& r, double& g1, double& g2) {
  const Gauss2Out o = gauss2v(r.t0, r.t1, r.stream, r.draw, r.h2, r.h3);
  r.draw = o.draw; r.h2 = o.h2; r.h3 = o.h3;
  g1 = o.g1; g2 = o.g2;
}

__device__ __forceinline__ void musc_refresh(TrackDev& t) {
  const double beta = t.p / sqrt(t.m2 + t.p * t.p);
  t.mc1 = 13.6 / t.p / beta;
  t.mbeta2 = beta * beta;
}

// loren.f:1-26
__device__ __forceinline__ void loren(double gam, double bx, double by, double bz, double e, double x, double y,
                                      double z, double& pxf, double& pyf, double& pzf, double& pf1) {
  const double gam1 = gam * gam / (1. + gam);
  pxf = (1 + gam1 * bx * bx) * x + gam1 * bx * (by * y + bz * z) - gam * bx * e;
  pyf = (1 + gam1 * by * by) * y + gam1 * by * (bx * x + bz * z) - gam * by * e;
  pzf = (1 + gam1 * bz * bz) * z + gam1 * bz * (by * y + bx * x) - gam * bz * e;
  pf1 = sqrt(pxf * pxf + pyf * pyf + pzf * pzf);
}

// decay kinematics common to project.f:67-110 and transp.f:147-186 / :236-275
__device__ __forceinline__ void decay_in_flight(TrackDev& t, DevRng& r, double p_spec, double beta, double gamma,
                                             double kaon_pipi_mfinal) {
  const double rph = r.uniform() * 2. * SIMC_PI;
  const double rth1 = r.uniform() * 2. - 1.;
  const double rth = m::acos(rth1);
  double pr = 0.;
  double m_final = SIMC_MMU;
  const double m = sqrt(t.m2);
  if (fabs(m - SIMC_MPI) < 2) pr = 29.783;
  if (fabs(m - SIMC_MK) < 2) {
    if (r.uniform() < 0.7) {
      pr = 235.5;
    } else {
      pr = sqrt(SIMC_MK * SIMC_MK / 4. - SIMC_MPI * SIMC_MPI);
      m_final = kaon_pipi_mfinal;
    }
  }
  // pr == 0 is a fatal `stop` in the reference; the host refuses decay_flag for other masses.
  const double er = sqrt(m_final * m_final + pr * pr);
  const double pxr = pr * m::sin(rth) * m::cos(rph);
  const double pyr = pr * m::sin(rth) * m::sin(rph);
  const double pzr = pr * m::cos(rth);
  const double nrm = sqrt(1. + t.dxdzs * t.dxdzs + t.dydzs * t.dydzs);
  const double bx = -beta * t.dxdzs / nrm;
  const double by = -beta * t.dydzs / nrm;
  const double bz = -beta * 1. / nrm;
  double pxf, pyf, pzf, pf;
  loren(gamma, bx, by, bz, er, pxr, pyr, pzr, pxf, pyf, pzf, pf);
  t.dxdzs = pxf / pzf;
  t.dydzs = pyf / pzf;
  t.dpps = 100. * (pf / p_spec - 1.);
  t.p = pf;
  t.m2 = m_final * m_final;
  t.mh2_final = t.m2;
  musc_refresh(t);
}

// shared/project.f:43-119, the branch that tests for a decay
__device__ __forceinline__ void project_decay(TrackDev& t, DevRng& r, double z_drift) {
  const double p_spec = t.p / (1. + t.dpps / 100.);
  const double beta = t.p / sqrt(t.p * t.p + t.m2);
  const double gamma = 1. / sqrt(1. - beta * beta);
  const double dlen = t.ctau * beta * gamma;
  const double z_decay = -1. * dlen * m::log(1 - r.uniform());
  if (z_decay > z_drift * sqrt(1 + t.dxdzs * t.dxdzs + t.dydzs * t.dydzs)) {
    t.decdist = t.decdist + z_drift * sqrt(1 + t.dxdzs * t.dxdzs + t.dydzs * t.dydzs);
    t.pathlen = t.pathlen + z_drift * sqrt(1 + t.dxdzs * t.dxdzs + t.dydzs * t.dydzs);
    t.xs = t.xs + t.dxdzs * z_drift;
    t.ys = t.ys + t.dydzs * z_drift;
  } else {
    t.dflag = true;
    t.decdist = t.decdist + z_decay;
    t.pathlen = t.pathlen + z_decay;
    t.xs = t.xs + t.dxdzs * z_decay / sqrt(1 + t.dxdzs * t.dxdzs + t.dydzs * t.dydzs);
    t.ys = t.ys + t.dydzs * z_decay / sqrt(1 + t.dxdzs * t.dxdzs + t.dydzs * t.dydzs);
    decay_in_flight(t, r, p_spec, beta, gamma, SIMC_MPI);
    const double tmpdrift = z_drift - z_decay / sqrt(1 + t.dxdzs * t.dxdzs + t.dydzs * t.dydzs);
    t.pathlen = t.pathlen + tmpdrift * sqrt(1 + t.dxdzs * t.dxdzs + t.dydzs * t.dydzs);
    t.xs = t.xs + t.dxdzs * tmpdrift;
    t.ys = t.ys + t.dydzs * tmpdrift;
  }
}

// shared/project.f:1-122
__device__ __forceinline__ void project(TrackDev& t, DevRng& r, double z_drift, bool decay_flag) {
  if (!decay_flag || t.dflag) {
    t.pathlen = t.pathlen + z_drift * sqrt(1 + t.dxdzs * t.dxdzs + t.dydzs * t.dydzs);
    t.xs = t.xs + t.dxdzs * z_drift;
    t.ys = t.ys + t.dydzs * z_drift;
  } else {
    project_decay(t, r, z_drift);
  }
}

// ---- COSY polynomial: branch-free term stream --------------------------------------------
// The thread's column of the CTA's shared power table holds kPolyEntries doubles: x^a*theta^b for
// a+b <= 6, then y^e, phi^e, delta^e for e = 0..6 (arm_program.h).  A term is four table reads,
// three multiplies in the reference's left-to-right order (a unit factor multiplies exactly), and one
// multiply + add per output with the coefficient read straight from the record, zeros included:
// the same operations as shared/transp.f:205-214, with no data-dependent branch in the loop.
// The records are warp-uniform: the warp copies them chunk by chunk (24 lanes x 16 bytes, one
// coalesced load) into its own double-buffered ring in shared memory, one chunk ahead of the
// arithmetic, and every lane reads them back as broadcasts.  ALL 32 LANES MUST CALL eval_poly.
__device__ __forceinline__ double lds_f64(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f64(unsigned addr, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void lds128(unsigned addr, unsigned long long& a, double& b) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=d"(b) : "r"(addr));
}
__device__ __forceinline__ void lds128d(unsigned addr, double& a, double& b) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}
__device__ __forceinline__ void sts128d(unsigned addr, double a, double b) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void ldg128(const double* p, double& a, double& b) {
  asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "l"(p));
}

// ---- Gaussians drawn ahead -------------------------------------------------------------------
// Inside the hut every random number is a gauss1(99.) (musc, musc_ext, the chamber resolutions; hundreds per
// track), and the sequence is fixed by the arm program.  Drawn one call at a time, the warp waits for its slowest
// lane in every call: ~2.5 trips through the polar rejection loop per Gaussian where a lane alone needs 4/pi.
// Here every lane draws its next `n` Gaussians in one go, at its own pace, into its column of the shared power
// table (free until the reconstruction map): the warp waits for max over lanes of the SUM of the trips
// (~1.5 per Gaussian for n = 20), and log/sqrt/div then run for all lanes with no divergence.  Each lane consumes
// exactly the uniforms of the sequential calls, in the same order.  Rows of the column: [0,Q) v1 then g,
// [Q,2Q) s, then the draw counter after each Gaussian as 32-bit words (a track that stops with Gaussians still
// queued must report the number of draws it really consumed).
struct GaussQueue {
  unsigned base;          // shared address of row 0 of this thread's column
  int pos, avail;         // warp-uniform: next entry, entries left
  int ts_pos;             // warp-uniform: next multiple-scattering width of the chunk (rows kGaussQ + m)
  uint32_t draw_before;   // draw counter before entry 0
};
constexpr unsigned kRowBytes = kBlock * 8u;
__device__ __forceinline__ void gq_init(GaussQueue& q, const double* pw) {
  q.base = (unsigned)__cvta_generic_to_shared(pw); q.pos = 0; q.avail = 0; q.ts_pos = 0; q.draw_before = 0u;
}
__device__ __forceinline__ unsigned gq_draw_addr(const GaussQueue& q, int k) {
  // 32-bit words behind the 2*Q double rows, two per 8-byte slot of the thread's OWN column (other threads'
  // columns may hold a power table at the same time: warps do not walk the program in step)
  return q.base + (unsigned)(2 * kGaussQ + (k >> 1)) * kRowBytes + (unsigned)(k & 1) * 4u;
}
// Gaussians an op takes from the queue (warp-uniform: flags and op constants only)
__device__ __forceinline__ int gauss_need(const ArmOp* o, bool ms_flag, bool wcs_flag) {
  const int op = o->op;
  if (op == OP_DC_PLANE) return wcs_flag ? 2 : 0;
  if (op == OP_MUSC) return (ms_flag && o->a != 0.) ? 2 : 0;
  if (op == OP_MUSC_EXT) return (ms_flag && o->a != 0.) ? 4 : 0;
  return 0;
}
// musc.f:52 / musc_ext.f:45: theta_sigma = Es/p/beta * sqrt(radw) * (1 + 0.088*log10(radw/beta**2)), with
// mc1 = 13.6/p/beta and mbeta2 = beta**2 of the track (musc_refresh) and b = sqrt(radw) from the host
__device__ __forceinline__ double musc_width(double mc1, double mbeta2, double a, double b) {
  return mc1 * b * (1 + 0.088 * fastlog::log10(a / mbeta2));
}
// One chunk of the queue.  All 32 lanes call, with an EMPTY queue (chunks end on op boundaries, so the queue runs dry
// exactly where the next gauss op starts).  The chunk = the ops from `pc` on whose Gaussians fit kGaussQ entries, up to
// the end of the stretch (`rem` Gaussians away, optics_host.cpp).  `live` lanes
//   1. draw the chunk's Gaussians at their own pace (polar rejection, gauss1.f),
//   2. transform them two at a time -- log, divide and square root inlined, two independent chains per warp,
//   3. evaluate the multiple-scattering width of every musc / musc_ext op of the chunk, two at a time as well (they
//      depend on the op's constants and on the track's p and beta only, which nothing in a hut changes): the walk
//      through the ops afterwards calls no library routine.
// Queue and generator go in and out by value, so that the caller's copies stay in registers.
struct GqOut { uint32_t draw, h2, h3, draw_before; int n_new; };
__device__ __noinline__ GqOut gq_fill_v(const ArmDev* arm, GaussQueue q, DevRng rng, int pc, int op_end, int rem, bool ms_flag,
                                        bool wcs_flag, double mc1, double mbeta2, bool live) {
  const unsigned mask = 0xffffffffu;
  // ---- the chunk: whole ops only
  int n_new = 0, pc_end = pc;
  for (; pc_end < op_end && n_new < rem; ++pc_end) {
    const int nd = gauss_need(&arm->ops[pc_end], ms_flag, wcs_flag);
    if (n_new + nd > kGaussQ) break;
    n_new += nd;
  }
  q.draw_before = rng.draw;
  const int target = live ? n_new : 0;
  int k = 0;
  uint32_t draw = rng.draw, h2 = rng.h2, h3 = rng.h3;
  const uint32_t t0 = rng.t0, t1 = rng.t1, stream = rng.stream;
  while (__any_sync(mask, k < target)) {
    if (k < target) {
      uint32_t r0, r1, r2, r3, w0, w1, w2, w3;
      const uint32_t b = draw >> 1;
      if (!(draw & 1u)) {
        philox4x32_10(b, stream, t0, t1, r0, r1, r2, r3);
        w0 = r0; w1 = r1; w2 = r2; w3 = r3;
      } else {
        w0 = h2; w1 = h3;
        philox4x32_10(b + 1u, stream, t0, t1, r0, r1, r2, r3);
        w2 = r0; w3 = r1;
        h2 = r2; h3 = r3;
      }
      draw += 2u;
      const double v1 = philox_to_pm1(w0, w1);
      const double v2 = philox_to_pm1(w2, w3);
      const double sq = v1 * v1 + v2 * v2;
      if (!(sq > 1. || sq == 0.)) {
        sts_f64(q.base + (unsigned)k * kRowBytes, v1);
        sts_f64(q.base + (unsigned)(kGaussQ + k) * kRowBytes, sq);
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(gq_draw_addr(q, k)), "r"(draw) : "memory");
        ++k;
      }
    }
  }
  // gauss1.f: g = v1*sqrt(-2.*log(s)/s); |g| <= 12 for the smallest s the 52-bit uniforms can give, so the
  // nsigmax = 99 test of these calls can never fire
  if (live) {
    int j = 0;
    for (; j + 1 < n_new; j += 2) {
      const double va = lds_f64(q.base + (unsigned)j * kRowBytes), vb = lds_f64(q.base + (unsigned)(j + 1) * kRowBytes);
      const double sa = lds_f64(q.base + (unsigned)(kGaussQ + j) * kRowBytes), sb = lds_f64(q.base + (unsigned)(kGaussQ + j + 1) * kRowBytes);
      const double ga = va * sqrt(-2. * fastlog::log(sa) / sa);
      const double gb = vb * sqrt(-2. * fastlog::log(sb) / sb);
      sts_f64(q.base + (unsigned)j * kRowBytes, ga);
      sts_f64(q.base + (unsigned)(j + 1) * kRowBytes, gb);
    }
    if (j < n_new) {
      const double va = lds_f64(q.base + (unsigned)j * kRowBytes), sa = lds_f64(q.base + (unsigned)(kGaussQ + j) * kRowBytes);
      sts_f64(q.base + (unsigned)j * kRowBytes, va * sqrt(-2. * fastlog::log(sa) / sa));
    }
    // the widths of the chunk's multiple-scattering ops go where the s values were
    if (ms_flag) {
      int m = 0;
      bool have = false;
      double a0 = 0., b0 = 0.;
      for (int p = pc; p < pc_end; ++p) {
        const ArmOp* o = &arm->ops[p];
        if (!((o->op == OP_MUSC || o->op == OP_MUSC_EXT) && o->a != 0.)) continue;
        if (!have) { a0 = o->a; b0 = o->b; have = true; continue; }
        const double w0 = musc_width(mc1, mbeta2, a0, b0);
        const double w1 = musc_width(mc1, mbeta2, o->a, o->b);
        sts_f64(q.base + (unsigned)(kGaussQ + m) * kRowBytes, w0);
        sts_f64(q.base + (unsigned)(kGaussQ + m + 1) * kRowBytes, w1);
        m += 2; have = false;
      }
      if (have) sts_f64(q.base + (unsigned)(kGaussQ + m) * kRowBytes, musc_width(mc1, mbeta2, a0, b0));
    }
  }
  __syncwarp();
  GqOut o;
  o.draw = draw; o.h2 = h2; o.h3 = h3; o.draw_before = q.draw_before; o.n_new = n_new;
  return o;
}
__device__ __forceinline__ void gq_fill(const ArmDev* arm, GaussQueue& q, DevRng& rng, int pc, int op_end, int rem, bool ms_flag,
                                        bool wcs_flag, double mc1, double mbeta2, bool live) {
  const GqOut o = gq_fill_v(arm, q, rng, pc, op_end, rem, ms_flag, wcs_flag, mc1, mbeta2, live);
  rng.draw = o.draw; rng.h2 = o.h2; rng.h3 = o.h3;
  q.draw_before = o.draw_before; q.avail = o.n_new; q.pos = 0; q.ts_pos = 0;
}
__device__ __forceinline__ double gq_at(const GaussQueue& q, int k) { return lds_f64(q.base + (unsigned)k * kRowBytes); }
__device__ __forceinline__ double gq_ts(const GaussQueue& q, int m) { return lds_f64(q.base + (unsigned)(kGaussQ + m) * kRowBytes); }
// The draw counter of a track that stops now, with q.avail Gaussians drawn ahead but not consumed
__device__ __forceinline__ uint32_t gq_draw_consumed(const GaussQueue& q) {
  if (q.pos == 0) return q.draw_before;
  uint32_t d;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(d) : "r"(gq_draw_addr(q, q.pos - 1)));
  return d;
}

constexpr unsigned kChunkBytes = kRecChunk * kRecWords * 8u;        // 384
constexpr unsigned kRingBytesPerWarp = 2u * kChunkBytes;            // double buffer
constexpr size_t kArmSmemBytes = kPowBytes + (kBlock / 32) * kRingBytesPerWarp;

// Evaluates one compiled map at v = (v1..v5).  pw = this thread's column of the shared power table,
// ring = shared address of this warp's record ring.  Lanes without a live track pass anything finite.
template <int NOUT>
__device__ __forceinline__ void eval_poly(const PolyClass& pc, const double* __restrict__ recs, const double (&v)[5],
                                       double* pw, unsigned ring, double (&sum)[NOUT]) {
  const unsigned base = (unsigned)__cvta_generic_to_shared(pw);
  constexpr unsigned S = kBlock * 8u;
  const unsigned lane = threadIdx.x & 31u;
  const bool loader = lane < kChunkBytes / 16u;
  const double* g = recs + pc.rec_begin + 2 * lane;
  const int nch = pc.n_chunks;
  double n0 = 0.0, n1 = 0.0;
  if (loader && nch > 0) ldg128(g, n0, n1);
  {
    double xp[7], tp[7];
    // libgcc __powidf2 association: x^3 = x*(x*x), x^5 = x*(x^2)^2, x^6 = x^2*x^4
    xp[0] = 1.0; xp[1] = v[0]; xp[2] = v[0] * v[0]; xp[3] = v[0] * xp[2]; xp[4] = xp[2] * xp[2];
    xp[5] = v[0] * xp[4]; xp[6] = xp[2] * xp[4];
    tp[0] = 1.0; tp[1] = v[1]; tp[2] = v[1] * v[1]; tp[3] = v[1] * tp[2]; tp[4] = tp[2] * tp[2];
    tp[5] = v[1] * tp[4]; tp[6] = tp[2] * tp[4];
#pragma unroll
    for (int a = 0; a < 7; ++a)
#pragma unroll
      for (int b = 0; b + a < 7; ++b)
        sts_f64(base + (unsigned)poly_xt_index(a, b) * S, a == 0 ? tp[b] : b == 0 ? xp[a] : xp[a] * tp[b]);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double a = v[2 + j], a2 = a * a, a4 = a2 * a2;
      const unsigned q = base + (unsigned)(28 + 7 * j) * S;
      sts_f64(q, 1.0); sts_f64(q + S, a); sts_f64(q + 2 * S, a2); sts_f64(q + 3 * S, a * a2); sts_f64(q + 4 * S, a4);
      sts_f64(q + 5 * S, a * a4); sts_f64(q + 6 * S, a2 * a4);
    }
  }
  double acc[5];
#pragma unroll
  for (int o = 0; o < 5; ++o) acc[o] = 0.0;
  for (int c = 0; c < nch; ++c) {
    const unsigned buf = ring + (unsigned)(c & 1) * kChunkBytes;
    if (loader) sts128d(buf + lane * 16u, n0, n1);
    if (loader && c + 1 < nch) ldg128(g + (size_t)(c + 1) * (kRecChunk * kRecWords), n0, n1);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < kRecChunk; ++k) {
      const unsigned rb = buf + (unsigned)k * (kRecWords * 8u);
      unsigned long long d;
      double c0, c1, c2, c3, c4;
      lds128(rb, d, c0);
      lds128d(rb + 16u, c1, c2);
      lds128d(rb + 32u, c3, c4);
      double t = lds_f64(base + ((unsigned)d & 0xffffu));
      const double s3 = lds_f64(base + ((unsigned)(d >> 16) & 0xffffu));
      const double s4 = lds_f64(base + ((unsigned)(d >> 32) & 0xffffu));
      const double s5 = lds_f64(base + (unsigned)(d >> 48));
      t = t * s3;
      t = t * s4;
      t = t * s5;
#if SIMC_STRICT
      acc[0] = acc[0] + t * c0; acc[1] = acc[1] + t * c1; acc[2] = acc[2] + t * c2; acc[3] = acc[3] + t * c3;
      if (NOUT == 5) acc[4] = acc[4] + t * c4;
#else
      acc[0] = fma(t, c0, acc[0]); acc[1] = fma(t, c1, acc[1]); acc[2] = fma(t, c2, acc[2]); acc[3] = fma(t, c3, acc[3]);
      if (NOUT == 5) acc[4] = fma(t, c4, acc[4]);
#endif
    }
  }
  __syncwarp();
#pragma unroll
  for (int o = 0; o < NOUT; ++o) sum[o] = acc[o];
}

// shared/transp.f:134-279
// Called by all 32 lanes (eval_poly is warp-cooperative); `live` = this lane has a track.
__device__ __forceinline__ void transp(const ArmDev* arm, TrackDev& t, DevRng& r, int klass, double zd,
                                    bool decay_flag, double* pw, unsigned ring, bool live) {
  double p_spec = 0, beta = 0, gamma = 0, z_decay = 0;
  const bool check = live && decay_flag && !t.dflag;
  if (check) {
    p_spec = t.p / (1. + t.dpps / 100.);
    beta = t.p / sqrt(t.p * t.p + t.m2);
    gamma = 1. / sqrt(1. - beta * beta);
    const double dlen = t.ctau * beta * gamma;
    z_decay = -1. * dlen * m::log(1 - r.uniform());
    if (z_decay <= zd / 2) {
      t.dflag = true;
      t.decdist = t.decdist + z_decay;
      decay_in_flight(t, r, p_spec, beta, gamma, SIMC_MK);      // m_final = Mk as written, transp.f:158
    }
  }
  const double ray[5] = {t.xs, t.dxdzs * 1000., t.ys, t.dydzs * 1000., t.dpps};
  double sum[5];
  eval_poly<5>(arm->tab.fwd[klass - 1], arm->tab.recs, ray, pw, ring, sum);
  if (!live) return;
  t.xs = sum[0];
  t.dxdzs = sum[1] / 1000.;
  t.ys = sum[2];
  t.dydzs = sum[3] / 1000.;
  const double delta_z = -sum[4];
  if (decay_flag && !t.dflag) {
    if (z_decay > zd + delta_z) {
      t.decdist = t.decdist + (zd + delta_z);
    } else {
      t.dflag = true;
      t.decdist = t.decdist + z_decay;
      decay_in_flight(t, r, p_spec, beta, gamma, SIMC_MPI);
    }
  }
  t.pathlen = t.pathlen + (zd + delta_z);
}

// hms/mc_hms.f:445-492 with hms/apertures_hms.inc
__device__ __forceinline__ bool hms_hit_dipole(double x, double y) {
  const double xl = fabs(x), yl = fabs(y);
  const bool c1 = (xl <= 34.29) && (yl <= 12.07);
  const bool c2 = (xl <= 27.94) && (yl <= 18.42);
  const bool c3 = (xl <= 13.97) && (yl <= 18.95);
  const bool c4 = (xl <= 1.956) && (yl <= 20.32);
  const bool c5 = ((xl - 27.94) * (xl - 27.94) + (yl - 12.065) * (yl - 12.065)) <= 6.35 * 6.35;
  const bool c6 = (xl >= 1.956) && (xl <= 13.97) && ((yl - (-0.114) * xl - 20.54) <= 0.0);
  return !(c1 || c2 || c3 || c4 || c5 || c6);
}

// cern/lfit.f:11-61 with KEY=0: REAL*4 points, sums in 8-byte reals
__device__ __noinline__ void lfit12(const float* z, const float* y, float& a, float& b) {
  a = 0.f; b = 0.f;
  double count = 0., sumx = 0., sumy = 0., sumxy = 0., sumxx = 0.;
  for (int j = 0; j < 12; ++j) {
    if (y[j] == 0.f) continue;
    sumx = sumx + (double)z[j];
    sumy = sumy + (double)y[j];
    count = count + 1.0;
  }
  if (count <= 1.) return;
  const double ymed = sumy / count, xmed = sumx / count;
  for (int j = 0; j < 12; ++j) {
    if (y[j] == 0.f) continue;
    const double sx = (double)z[j] - xmed, sy = (double)y[j] - ymed;
    sumxy = sumxy + sx * sy;
    sumxx = sumxx + sx * sx;
  }
  if (sumxx == 0.) return;
  a = (float)(sumxy / sumxx);
  b = (float)(ymed - (double)a * xmed);
}

// Adds the number of currently converged lanes to *counter with one atomic (all of them are at
// the same warp-uniform place of the arm program).
__device__ __forceinline__ void warp_count(unsigned* counter) {
  const unsigned m = __activemask();
  if ((threadIdx.x & 31u) == (unsigned)(__ffs(m) - 1)) atomicAdd(counter, (unsigned)__popc(m));
}
// Same with an explicit predicate, for places where dead lanes are converged with live ones.
__device__ __forceinline__ void warp_count_if_alive(unsigned* counter, bool alive) {
  const unsigned m = __ballot_sync(0xffffffffu, alive);
  if ((threadIdx.x & 31u) == 0 && m) atomicAdd(counter, (unsigned)__popc(m));
}
// Histogram increment aggregated over the lanes that hit the same bin (bin < 0: no increment).
__device__ __forceinline__ void warp_hist_add(unsigned* hist, int bin) {
  const unsigned act = __activemask();
  const unsigned peers = __match_any_sync(act, bin);
  if (bin >= 0 && (threadIdx.x & 31u) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&hist[bin], (unsigned)__popc(peers));
}

struct ArmFlags {
  bool ms_flag, wcs_flag, decay_flag, using_coll;
};

// Hut state that has to survive from the drift-chamber planes to the fit (REAL*4 in the reference)
struct HutState {
  float xdc[12], ydc[12];
  int scincount;
};

__device__ __forceinline__ void arm_result_clear(ArmResult& res) {
  res.ok = false; res.stop_code = 0; res.reached_hut = false; res.resmult = 0.0;
  res.x_fp = res.dx_fp = res.y_fp = res.dy_fp = 0.0;
  res.dpp_rec = res.dth_rec = res.dph_rec = res.y_rec = 0.0;
}

// hms/pion_coll_absorb.f:1-95: transmission of a pion through `thick` cm of the collimator alloy
__device__ const double kCollT[14] = {85.0, 125.0, 165.0, 205.0, 245.0, 315.0, 584.02, 711.95, 870.12, 1227.57, 1446.58,
                                      1865.29, 2858.0, 4159.0};
__device__ const double kCollSig[14] = {26.03, 84.47, 117.3, 117.4, 101.9, 69.58, 42.5, 44.7, 47.9, 46.5, 45.2, 39.6, 35.34,
                                        33.15};
__device__ const double kCollQ[14] = {0.948, 0.659, 0.56, 0.5342, 0.5452, 0.60796, 0.699, 0.689, 0.679, 0.683, 0.688, 0.704,
                                      0.7483, 0.7705};
__device__ __noinline__ double pion_coll_absorb(double ppi, double thick) {
  const double* T = kCollT;
  const double* sigreac = kCollSig;
  const double* qreac = kCollQ;
  const double Navagadro = 6.0221367e+23, mpi = 139.56995, mate_dens = 17.0, mate_A = 171.57;
  const double Tpi = sqrt(ppi * ppi + mpi * mpi) - mpi;
  double sigA = 0.0;
  for (int i = 1; i <= 13; ++i) {
    if ((Tpi > T[i - 1]) && (Tpi <= T[i])) {
      const double Thi = T[i], Tlo = T[i - 1];
      const double sigAhi = sigreac[i] * m::pow(mate_A, qreac[i]);
      const double sigAlo = sigreac[i - 1] * m::pow(mate_A, qreac[i - 1]);
      sigA = (sigAlo * (Thi - Tpi) + sigAhi * (Tpi - Tlo)) / (Thi - Tlo);
      sigA = sigA * 1.e-27;
    }
  }
  if (Tpi > T[13]) sigA = (sigreac[13] * m::pow(mate_A, qreac[13])) * 1.e-27;
  const double lambdai = mate_dens * Navagadro * sigA / mate_A;
  return m::exp(-(0.0 + thick * lambdai));
}

__device__ __forceinline__ bool collimator_steps_body(const ArmOp* o, TrackDev& t, DevRng& rng, bool decay_flag, unsigned* stop_counts);
// mc_hms_coll (hms/mc_hms_coll.f:1-147) / mc_shms_coll: 20 slices through the collimator; in the material the
// particle may be absorbed (pions), scatters, loses energy (sampled), and decays in flight.  `o` is the OP_COLL
// op, followed by its two data ops.  Returns false when the particle is lost.  The slit STOP counters are bumped
// once per slice spent in the material, survivors included (mc_hms_coll.f:95-115).
// Track and generator go in and out by value, so the caller's copies stay in registers (this code only runs for
// decks with using_HMScoll / using_SHMScoll).
struct CollOut { TrackDev t; uint32_t draw, h2, h3; bool ok; };
__device__ __noinline__ CollOut collimator_steps(const ArmOp* o, TrackDev t, DevRng rng, bool decay_flag, unsigned* stop_counts) {
  CollOut R;
  R.ok = collimator_steps_body(o, t, rng, decay_flag, stop_counts);
  R.t = t; R.draw = rng.draw; R.h2 = rng.h2; R.h3 = rng.h3;
  return R;
}
__device__ __forceinline__ bool collimator_steps_body(const ArmOp* o, TrackDev& t, DevRng& rng, bool decay_flag, unsigned* stop_counts) {
  const ArmOp* d1 = o + 1;
  const ArmOp* d2 = o + 2;
  MatConst mc;
  mc.rho = d1->c; mc.I = 0.; mc.CO = d1->d; mc.co27 = d1->e; mc.ln10 = d2->a; mc.log_me_I2 = d2->b; mc.p_mp = d2->c;
  mc.p_log = d2->d; mc.p_chsi = d2->e;
  const int nstep = 20;
  const double step_size = d1->a, radl = d1->b, y_off = o->e;
  const double p_spec = t.p_spec;
  double h_step = o->a, v_step = o->b;
  for (int n = 1; n <= nstep; ++n) {
    const bool hor = fabs(t.ys - y_off) > h_step;
    const bool ver = fabs(t.xs - 0.000) > v_step;
    const bool oct = fabs(t.xs - 0.000) > (-v_step / h_step * fabs(t.ys - y_off) + 3 * v_step / 2);
    if (stop_counts) {
      if (hor) atomicAdd(&stop_counts[2 + o->i1], 1u);
      if (ver) atomicAdd(&stop_counts[2 + o->i1 + 1], 1u);
      if (oct) atomicAdd(&stop_counts[2 + o->i1 + 2], 1u);
    }
    if (hor || ver || oct) {
      const double thick = step_size;
      if (t.m2 > 12000. && t.m2 < 20000.) {          // pions only: hadronic interaction
        const double trans = pion_coll_absorb(t.p, thick);
        if (rng.uniform() > trans) return false;
      }
      const double radw = thick / radl;
      const double ts = t.mc1 * sqrt(radw) * (1 + 0.088 * m::log10(radw / t.mbeta2));      // musc(m2,p,radw,dydzs,dxdzs)
      t.dydzs = t.dydzs + ts * gauss1(rng, 99.0);
      t.dxdzs = t.dxdzs + ts * gauss1(rng, 99.0);
      double epart = sqrt(t.p * t.p + t.m2);
      const double mpart = sqrt(t.m2);
      const ParticleKin k = particle_kin(epart, mpart, mc.ln10);
      const double x = fabs(gauss1(rng, 10.0));
      const double eloss = enerloss_material(k, thick, mc, x);
      epart = epart - eloss;
      if (epart < mpart) return false;
      t.p = sqrt(epart * epart - t.m2);
      t.dpps = 100. * (t.p / p_spec - 1.);
      musc_refresh(t);
    }
    project(t, rng, step_size, decay_flag);
    h_step = h_step + (o->c - o->a) / nstep;
    v_step = v_step + (o->d - o->b) / nstep;
  }
  return true;
}

// Interprets ops [op_begin, op_end) of the arm program.  EVERY lane of the warp must call this
// (lanes without an event pass alive = false): the warp walks the program in lock step and
// re-converges before each op, so a lane only idles while other lanes still have work in the same
// op.  `alive` comes back false when the event stopped (res.stop_code) or finished (res.ok).
// `pw` = this thread's column of the CTA's shared power table.
// WITH_COLL: the kernel was built with the collimator stepping of OP_COLL (only launched for decks that ask for
// it, so that the other kernels do not carry its call).
// NO_MAPS: the op range holds no COSY map (the hut in front of a compiled reconstruction map): their code, and the
// registers it needs, stay out of the kernel.
template <bool WITH_COLL = false, bool NO_MAPS = false>
__device__ __forceinline__ void run_arm(const ArmDev* arm, TrackDev& t, DevRng& rng, const ArmFlags f,
                                        double fry, double* pw, unsigned ring, ArmResult& res, HutState& hs,
                                        bool& alive, int op_begin, int op_end, unsigned* call_counts = nullptr,
                                        unsigned* stop_counts = nullptr) {
  double xt = 0., yt = 0.;
  int skip_until = 0;          // this lane ignores ops before this index (set by OP_COLL)
  // Gaussians are drawn ahead (GaussQueue) unless decays in flight put other draws between them
  const bool use_q = !f.decay_flag;
  GaussQueue gq;
  gq_init(gq, pw);
  for (int pc = op_begin; pc < op_end; ++pc) {
    __syncwarp();
    if (!__any_sync(0xffffffffu, alive)) break;
    const ArmOp* o = &arm->ops[pc];
    const int op = o->op;
    if (op == OP_END) break;
    const double a = o->a, b = o->b, c = o->c, d = o->d;
    // the two map evaluations are warp-cooperative: every lane goes in, dead ones compute on stale values
    if (!NO_MAPS && op == OP_TRANSP) {
      const int klass = o->i0;
      if (call_counts) warp_count_if_alive(&call_counts[klass - 1], alive);
      transp(arm, t, rng, klass, a, f.decay_flag, pw, ring, alive);
      continue;
    }
    if (!NO_MAPS && op == OP_RECON) {     // mc_hms.f:419-437 + mc_hms_recon.f:104-137
      double hut[5];
      hut[0] = res.x_fp / 100.;
      hut[1] = res.dx_fp;
      hut[2] = res.y_fp / 100.;
      hut[3] = res.dy_fp;
      hut[4] = fry / 100.;
      if (fabs(hut[4]) <= 1.e-30) hut[4] = 1.e-30;
      if (o->i0) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (fabs(hut[i]) <= 1.e-30) hut[i] = 1.e-30;
      }
      double sum[4];
      if (call_counts) warp_count_if_alive(&call_counts[47], alive);
      eval_poly<4>(arm->tab.rec, arm->tab.recs, hut, pw, ring, sum);
      if (alive) {
        if (a != 0.) res.y_fp = res.y_fp - a;
        res.dph_rec = sum[0];
        res.y_rec = sum[1] * 100.;
        res.dth_rec = sum[2];
        res.dpp_rec = sum[3] * 100.;
        res.ok = true;
        alive = false;
      }
      continue;
    }
    // this op's Gaussians: entries [qoff, qoff + need) of the queue, and its width if it is a multiple-scattering op;
    // pos / avail / ts_pos stay warp-uniform
    int qoff = 0, tsoff = 0;
    if (use_q && (op == OP_MUSC || op == OP_MUSC_EXT || op == OP_DC_PLANE)) {
      const int need = gauss_need(o, f.ms_flag, f.wcs_flag);
      if (need > gq.avail) {          // (avail == 0: chunks end on op boundaries)
        const unsigned code = (unsigned)o->code;      // Gaussians from this op to the end of the stretch (optics_host.cpp)
        const int rem = (f.ms_flag ? (int)(code & 0xffffu) : 0) + (f.wcs_flag ? (int)(code >> 16) : 0);
        gq_fill(arm, gq, rng, pc, op_end, rem, f.ms_flag, f.wcs_flag, t.mc1, t.mbeta2, alive);
      }
      qoff = gq.pos;
      gq.pos += need; gq.avail -= need;
      if (op != OP_DC_PLANE && need > 0) tsoff = gq.ts_pos++;
    }
    if (!alive || pc < skip_until) continue;
    bool stop = false;
    switch (op) {
      case OP_PROJECT: project(t, rng, a, f.decay_flag); break;
      case OP_CUT_R2: stop = (t.xs * t.xs + t.ys * t.ys) > a; break;
      case OP_CUT_ABS_Y: stop = fabs(t.ys - a) > b; break;
      case OP_CUT_ABS_X: stop = fabs(t.xs - a) > b; break;
      case OP_CUT_OCT: stop = fabs(t.xs - a) > (c * fabs(t.ys - b) + d); break;
      case OP_CUT_OFF_R2: stop = ((t.xs - a) * (t.xs - a) + (t.ys - b) * (t.ys - b)) > c; break;
      case OP_ROT_H: {   // rotate_haxis.f:46-62
        const double alpha = t.dxdzs, beta = t.dydzs;
        const double alpha_p = (alpha + a) / (1. - alpha * a);
        const double beta_p = beta / (c - alpha * b);
        const double xi = t.xs;
        xt = xi * (c + alpha_p * b);
        yt = t.ys + xi * beta_p * b;
        xt = xt + d;
        break;
      }
      case OP_ROT_V: {   // rotate_vaxis.f:42-58
        const double alpha = t.dydzs, beta = t.dxdzs;
        const double alpha_p = (alpha + a) / (1. - alpha * a);
        const double beta_p = beta / (c - alpha * b);
        const double yi = t.ys;
        yt = yi * (c + alpha_p * b);
        xt = t.xs + yi * beta_p * b;
        yt = yt + d;
        break;
      }
      case OP_CUT_T_R2: stop = (xt * xt + yt * yt) > a; break;
      case OP_CUT_HB: stop = (xt * xt > a) || (yt > b) || (yt < c); break;
      case OP_CUT_HMS_DIPOLE: stop = hms_hit_dipole(xt, yt); break;
      case OP_CUT_HMS_PIPE: stop = (((xt - a) * (xt - a) + (yt - b) * (yt - b)) > c) || (fabs(yt - b) > d); break;
      case OP_MARK_HUT: res.reached_hut = true; break;
      case OP_RESMULT_DRAW: res.resmult = (rng.uniform() < a) ? 2.0 : 1.0; break;
      case OP_RESMULT_ONE: res.resmult = 1.0; break;
      case OP_MUSC:        // musc.f:46-55, called as musc(m2,p,radw,dydzs,dxdzs)
        if (f.ms_flag && a != 0.) {
          const double ts = use_q ? gq_ts(gq, tsoff) : t.mc1 * b * (1 + 0.088 * m::log10(a / t.mbeta2));
          double g1, g2;
          if (use_q) { g1 = gq_at(gq, qoff); g2 = gq_at(gq, qoff + 1); }
          else gauss2(rng, g1, g2);
          t.dydzs = t.dydzs + ts * g1;
          t.dxdzs = t.dxdzs + ts * g2;
        }
        break;
      case OP_MUSC_EXT:    // musc_ext.f:37-51, called as musc_ext(m2,p,radw,drift,dydzs,dxdzs,ys,xs)
        if (f.ms_flag && a != 0.) {
          const double ts = use_q ? gq_ts(gq, tsoff) : t.mc1 * b * (1 + 0.088 * m::log10(a / t.mbeta2));
          double g1, g2;
          if (use_q) { g1 = gq_at(gq, qoff); g2 = gq_at(gq, qoff + 1); }
          else gauss2(rng, g1, g2);
          t.dxdzs = t.dxdzs + ts * g1;
          t.xs = t.xs + ts * c * g2 / 3.4641016151377544 + ts * c * g1 / 2.;
          if (use_q) { g1 = gq_at(gq, qoff + 2); g2 = gq_at(gq, qoff + 3); }
          else gauss2(rng, g1, g2);
          t.dydzs = t.dydzs + ts * g1;
          t.ys = t.ys + ts * c * g2 / 3.4641016151377544 + ts * c * g1 / 2.;
        }
        break;
      case OP_DC_PLANE: {  // mc_hms_hut.f:351-364
        double r1 = 0., r2 = 0.;
        if (f.wcs_flag) {
          if (use_q) { r1 = gq_at(gq, qoff); r2 = gq_at(gq, qoff + 1); }
          else gauss2(rng, r1, r2);
        }
        const int ip = o->i0;
        if (o->i1) { hs.ydc[ip] = (float)(t.ys + a * r2 * res.resmult); hs.xdc[ip] = 0.f; }
        else { hs.xdc[ip] = (float)(t.xs + a * r1 * res.resmult); hs.ydc[ip] = 0.f; }
        break;
      }
      case OP_CUT_BOX: stop = (t.xs > a) || (t.xs < b) || (t.ys > c) || (t.ys < d); break;
      case OP_CUT_R: stop = sqrt(t.xs * t.xs + t.ys * t.ys) > a; break;
      case OP_CUT_T_ABSX: stop = fabs(xt - a) > b; break;
      case OP_CUT_T_TRAP: stop = (fabs(yt) + a * xt) > b; break;
      case OP_CUT_T_RECT: stop = (xt > a) || (xt < b) || (yt > c) || (yt < d); break;
      case OP_CUT_T_BOX: stop = (yt > a) || (-yt > b) || (-xt > c) || (-xt < d); break;
      case OP_CUT_SOS_EXIT: {
        const double tmpwidth = a + b * (t.xs + c);
        stop = (fabs(t.xs) > c) || (fabs(t.ys) > tmpwidth);
        break;
      }
      case OP_SHIFT:
        t.xs = t.xs + a * t.dxdzs;
        t.ys = t.ys + a * t.dydzs;
        break;
      case OP_SCIN_COUNT: if (t.ys < a && t.ys > b && t.xs < c && t.xs > d) ++hs.scincount; break;
      case OP_SCIN_TRIG: stop = hs.scincount < o->i0; break;
      case OP_LFIT: {      // mc_hms_hut.f:438-455
        float zdc[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) {
          const int iplane = (j % 6) + 1;
          zdc[j] = (float)((j < 6 ? a : b) + (iplane - 0.5 - 0.5 * 6) * c);
        }
        float dx4, x4, dy4, y4;
        lfit12(zdc, hs.xdc, dx4, x4);
        lfit12(zdc, hs.ydc, dy4, y4);
        res.x_fp = (double)x4; res.dx_fp = (double)dx4; res.y_fp = (double)y4; res.dy_fp = (double)dy4;
        break;
      }
      case OP_CUT_FP_CAL: {   // mc_shms_hut.f:413-424
        const double e = o->e;
        const double xcal = res.x_fp + res.dx_fp * a;
        const double ycal = res.y_fp + res.dy_fp * a;
        stop = (ycal > b) || (ycal < c) || (xcal > d) || (xcal < e);
        break;
      }
      case OP_COLL:
        // mc_hms.f:206 / mc_shms.f:503: pions and muons are stepped through the collimator material
        if (WITH_COLL && f.using_coll && (t.m2 > 100.0 * 100.0) && (t.m2 < 200.0 * 200.0)) {
          const CollOut co = collimator_steps(o, t, rng, f.decay_flag, stop_counts);
          t = co.t; rng.draw = co.draw; rng.h2 = co.h2; rng.h3 = co.h3;
          stop = !co.ok;
          skip_until = pc + 3 + o->i0;          // the plain aperture checks belong to the other branch
        }
        break;
      case OP_COLL_DATA: break;
      default: break;
    }
    if (stop) {
      res.stop_code = o->code; alive = false;
      if (use_q && gq.avail > 0) rng.draw = gq_draw_consumed(gq);      // give back the Gaussians drawn ahead
    }
  }
  __syncwarp();
}

}  // namespace SIMC_VARIANT_NS
}  // namespace simc
