// ORACLE -- TEST INFRASTRUCTURE ONLY (see rng.hpp).
// RANLUX restated from cern/ranlux.f (James 1994; Luescher 1994), as built by the
// reference's flags: default REAL is 8 bytes, so SEEDS/UNI/CARRY/RVEC are doubles.
#include "rng.hpp"

namespace simc_oracle {

static const int ndskip[5] = {0, 24, 73, 199, 365};   // cern/ranlux.f:66

// cern/ranlux.f:209-283 (ENTRY RLUXGO); sgrnd() calls it with lux=3,k1=k2=0 (call_ranlux.f:6-17)
void RanluxState::rluxgo(int lux, int ins, int k1, int k2) {
  if (lux < 0) luxlev = 3;
  else if (lux <= maxlev) luxlev = lux;
  else if (lux < 24 || lux > 2000) luxlev = maxlev;
  else { luxlev = lux; for (int ilx = 0; ilx <= maxlev; ++ilx) if (lux == ndskip[ilx] + 24) luxlev = ilx; }
  nskip = (luxlev <= maxlev) ? ndskip[luxlev] : luxlev - 24;
  in24 = 0;
  int jseed = (ins > 0) ? ins : jsdflt;
  inseed = jseed;
  notyet = false;
  twom24 = 1.;
  int iseeds[25];
  for (int i = 1; i <= 24; ++i) {
    twom24 *= 0.5;
    const int k = jseed / 53668;
    jseed = 40014 * (jseed - k * 53668) - k * 12211;
    if (jseed < 0) jseed += icons;
    iseeds[i] = jseed % itwo24;
  }
  twom12 = twom24 * 4096.;
  for (int i = 1; i <= 24; ++i) { seeds[i] = (double)iseeds[i] * twom24; next[i] = i - 1; }
  next[1] = 24; i24 = 24; j24 = 10; carry = 0.;
  if (seeds[24] == 0.) carry = twom24;
  kount = k1; mkount = k2;
  if (k1 + k2 != 0) {
    for (int iouter = 1; iouter <= k2 + 1; ++iouter) {
      const int inner = (iouter == k2 + 1) ? k1 : igiga;
      for (int isk = 1; isk <= inner; ++isk) {
        double uni = seeds[j24] - seeds[i24] - carry;
        if (uni < 0.) { uni += 1.0; carry = twom24; } else carry = 0.;
        seeds[i24] = uni; i24 = next[i24]; j24 = next[j24];
      }
    }
    in24 = kount % (nskip + 24);
    if (mkount > 0) { const int izip = igiga % (nskip + 24); in24 = (mkount * izip + in24) % (nskip + 24); }
    if (in24 > 23) in24 = 0;
  }
  latest = 0;
}

// cern/ranlux.f:76-136
void RanluxState::ranlux(double* out, int lenv) {
  if (notyet) rluxgo(3, 0, 0, 0);           // default initialisation, ranlux.f:76-108
  for (int ivec = 0; ivec < lenv; ++ivec) {
    double uni = seeds[j24] - seeds[i24] - carry;
    if (uni < 0.) { uni += 1.0; carry = twom24; } else carry = 0.;
    seeds[i24] = uni; i24 = next[i24]; j24 = next[j24];
    out[ivec] = uni;
    if (uni < twom12) {                      // small numbers get extra bits, ranlux.f:121-124
      out[ivec] += twom24 * seeds[j24];
      if (out[ivec] == 0.) out[ivec] = twom24 * twom24;
    }
    if (++in24 == 24) {
      in24 = 0; kount += nskip;
      for (int isk = 1; isk <= nskip; ++isk) {
        double u2 = seeds[j24] - seeds[i24] - carry;
        if (u2 < 0.) { u2 += 1.0; carry = twom24; } else carry = 0.;
        seeds[i24] = u2; i24 = next[i24]; j24 = next[j24];
      }
    }
  }
  kount += lenv;
  if (kount >= igiga) { ++mkount; kount -= igiga; }
}

}  // namespace simc_oracle
