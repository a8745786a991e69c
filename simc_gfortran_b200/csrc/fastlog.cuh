// Double-precision natural logarithm and log10 for the event kernels: table-driven, fma-based, < 0.52 ulp.
//
// Why not the CUDA math library's log(): it is ~65 instructions on its common path (range reduction by a division,
// a long odd polynomial), and the loop calls a logarithm ~280 times per track in a detector hut (one per Gaussian of
// gauss1.f, one log10 per multiple-scattering step of musc.f) and ~130 times per try in generation (enerloss_new.f,
// brem.f).  This one is ~35: x = 2^k z with z in [0.6875, 1.375); step i of 128 in that range gives invc ~ 1/z and
// -log(invc) as a double-double from a 4 KB table (tools/gen_fastlog_table.py); r = z*invc - 1 is formed exactly (as a
// double-double, with one fma) and |r| < 2^-8, so log1p(r) needs the series to r^7 only; the sum k ln2 + logc + r is carried as (hi, lo).
// In the two steps around 1.0 the table has invc = 1, logc = 0: log(z) = log1p(z - 1) with the series to r^10,
// which keeps the RELATIVE accuracy where the result goes through zero.
// Accuracy (tests/test_fastlog_cpu.py against libquadmath on 4e7 arguments): max 0.52 ulp, i.e. correctly rounded
// but for rare cases, like glibc's log which the reference build calls; the CUDA library's is 1 ulp.
// Zero, negative, subnormal, infinite and NaN arguments go to the library's log.
#pragma once
#include <math.h>
#include "fastlog_table.h"

namespace simc {
namespace fastlog {

#if defined(__CUDACC__)
#define SIMC_FL_HD __host__ __device__ __forceinline__
__device__ const double kTab[128 * 4] = {SIMC_FASTLOG_TABLE};
#else
#define SIMC_FL_HD inline
#endif
static const double kTabHost[128 * 4] = {SIMC_FASTLOG_TABLE};

// log(x) = hi + lo for a positive normal x; returns false for anything else (hi, lo untouched)
SIMC_FL_HD bool log_dd(double x, double& hi, double& lo) {
  long long ix;
#if defined(__CUDA_ARCH__)
  ix = __double_as_longlong(x);
#else
  memcpy(&ix, &x, sizeof(ix));
#endif
  const unsigned top = (unsigned)((unsigned long long)ix >> 52);
  if (top - 0x001u >= 0x7ffu - 0x001u) return false;       // zero, subnormal, negative, inf, nan
  const long long tmp = ix - 0x3fe6000000000000LL;
  const int i = (int)((tmp >> 45) & 127);
  const int k = (int)(tmp >> 52);
  const long long iz = ix - (tmp & (long long)0xfff0000000000000ULL);
  double z;
#if defined(__CUDA_ARCH__)
  z = __longlong_as_double(iz);
  double invc, lchi, lclo, pad;
  asm("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(invc), "=d"(lchi) : "l"(kTab + 4 * i));
  asm("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(lclo), "=d"(pad) : "l"(kTab + 4 * i + 2));
#else
  memcpy(&z, &iz, sizeof(z));
  const double invc = kTabHost[4 * i], lchi = kTabHost[4 * i + 1], lclo = kTabHost[4 * i + 2];
#endif
  // r = z*invc - 1 as (r, rl): the product's high part is within 2^-7 of 1, so the subtraction is exact, and the fma
  // gives the product's rounding error exactly
  const double ph = z * invc;
  const double rl = fma(z, invc, -ph);
  const double r = ph - 1.0;
  const double kd = (double)k;
  const double t1 = kd * SIMC_LN2HI;                       // exact: SIMC_LN2HI has 42 significant bits
  const double w = t1 + lchi;                              // |t1| >= |lchi| or t1 == 0: fast two-sum
  const double e = (t1 - w) + lchi;
  hi = w + r;                                              // |w| >= |r| or w == 0
  const double r2 = r * r;
  double p;
  if (i == 79 || i == 80) {
    // log1p(r) - r for |r| < 2^-7, series to r^10
    p = fma(r, -1.0 / 10.0, 1.0 / 9.0);
    p = fma(r, p, -1.0 / 8.0);
    p = fma(r, p, 1.0 / 7.0);
    p = fma(r, p, -1.0 / 6.0);
    p = fma(r, p, 1.0 / 5.0);
    p = fma(r, p, -1.0 / 4.0);
    p = fma(r, p, 1.0 / 3.0);
    p = fma(r, p, -0.5);
    p = p * r2;
  } else {
    // log1p(r) - r for |r| < 2^-8, series to r^7 (Estrin)
    const double a = fma(r, 1.0 / 3.0, -0.5);
    const double b = fma(r, 1.0 / 5.0, -1.0 / 4.0);
    const double c = fma(r, 1.0 / 7.0, -1.0 / 6.0);
    p = fma(r2, c, b);
    p = fma(r2, p, a);
    p = p * r2;
  }
  // log1p(r + rl) = log1p(r) + rl / (1 + r) + ...: the low part enters with the derivative, rl * (1 - r)
  lo = ((w - hi) + r) + ((fma(kd, SIMC_LN2LO, lclo) + e) + fma(-rl, r, rl)) + p;
  return true;
}

SIMC_FL_HD double log(double x) {
  double hi, lo;
  if (!log_dd(x, hi, lo)) return ::log(x);
  return hi + lo;
}

// log10(x) = log(x) / ln 10 with the product carried in double-double
SIMC_FL_HD double log10(double x) {
  double hi, lo;
  if (!log_dd(x, hi, lo)) return ::log10(x);
  const double s = hi + lo;                                // renormalise
  const double t = (hi - s) + lo;
  const double ph = s * SIMC_ILN10_HI;
  const double pl = fma(s, SIMC_ILN10_HI, -ph) + fma(s, SIMC_ILN10_LO, t * SIMC_ILN10_HI);
  return ph + pl;
}

}  // namespace fastlog
}  // namespace simc
