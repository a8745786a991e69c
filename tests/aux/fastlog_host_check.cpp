#include <cstdio>
#include <cstring>
#include <cmath>
#include <cstdint>
#include <random>
#include <quadmath.h>
#include "fastlog.cuh"
static double ulp_err(double got, __float128 want) {
  double w = (double)want;
  int e; frexp(w, &e);
  double ulp = ldexp(1.0, e - 53);
  if (w == 0) return got == 0 ? 0 : 1e9;
  return (double)fabsq(((__float128)got - want) / ulp);
}
int main(int argc, char** argv) {
  long n = argc > 1 ? atol(argv[1]) : 10000000;
  std::mt19937_64 g(12345);
  double m1 = 0, m2 = 0, m3 = 0, mg = 0; long bad1 = 0, badg = 0;
  auto check = [&](double x) {
    double a = simc::fastlog::log(x), b = simc::fastlog::log10(x), c = std::log(x);
    __float128 q = logq((__float128)x), q10 = log10q((__float128)x);
    double e1 = ulp_err(a, q), e2 = ulp_err(b, q10), eg = ulp_err(c, q);
    if (e1 > m1) m1 = e1; if (e2 > m2) m2 = e2; if (eg > mg) mg = eg;
    if (e1 > 0.5) ++bad1; if (eg > 0.5) ++badg;
  };
  std::uniform_real_distribution<double> u01(0.0, 1.0);
  for (long i = 0; i < n; ++i) {
    check(u01(g));                                   // the Gaussians' argument
    check(0.99 + 0.02 * u01(g));                     // through zero
    check(std::exp(40.0 * (u01(g) - 0.5)));          // wide range
    uint64_t b = (g() & 0x7fefffffffffffffULL) | 0x0010000000000000ULL; double x; memcpy(&x, &b, 8); check(x);   // any positive normal
  }
  check(1.0); check(0.6875); check(1.375); check(nextafter(1.0, 0.0)); check(nextafter(1.0, 2.0)); check(2.0); check(0.5);
  printf("max ulp: log %.4f log10 %.4f (glibc log %.4f); not correctly rounded: fastlog %ld glibc %ld of %ld\n", m1, m2, mg, bad1, badg, 4 * n);
  printf("special: %g %g %g %g %g\n", simc::fastlog::log(0.0), simc::fastlog::log(-1.0), simc::fastlog::log(INFINITY), simc::fastlog::log(NAN), simc::fastlog::log(5e-324));
  return 0;
}
