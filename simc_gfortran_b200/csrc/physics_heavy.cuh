// A(e,e'p) weights on the device (and on the host for the weight scale): Benhar spectral-function
// lookup (sf_lookup.f:85-170) and the deForest off-shell cross sections sigma_cc1 / sigma_cc2
// (physics_proton.f:23-135).  The table is the caller's (simc_b200_set_sf_table), normalised to
// sum 1 on the host like sf_lookup_init (sf_lookup.f:64-78).
#pragma once
#include "target.cuh"

namespace simc {

struct SfDev {                // device pointers; val[iPm * n_em + iEm]
  const double* pm;
  const double* em;
  const double* val;
  int n_pm, n_em;
  const double* dem;          // widths of the Em bins (generate_em only; null unless set)
};

// generate_em (sf_lookup.f:181-245): missing energy drawn from the spectral function's Em distribution at fixed
// Pm.  The reference calls sf_lookup at every grid energy; at a grid point the interpolation returns the grid
// value (the slope term is 0 * finite), except at the last one, where no branch of sf_lookup matches and its
// SAVEd interval of the previous call is reused: sf(n-1) + (E_n - E_n-1) * (sf(n) - sf(n-1)) / (E_n - E_n-1).
// u1, u2: the two uniforms generate_em draws.  Two passes over the column pair instead of a 200-entry array.
SIMC_HD_CALL double generate_em(const SfDev T, double Pm, double u1, double u2) {   // by value: see loop.cuh
  const int numPm = T.n_pm, numEm = T.n_em;
  int iPm;
  double w1, w2;
  if (Pm >= T.pm[numPm - 1]) { iPm = numPm - 1; w1 = 0; w2 = 1; }
  else if (Pm <= T.pm[0]) { iPm = 1; w1 = 1; w2 = 0; }
  else {
    int ind = 1;
    while (Pm > T.pm[ind - 1]) ind = ind + 1;
    iPm = ind - 1;
    w2 = (Pm - T.pm[iPm - 1]) / (T.pm[iPm] - T.pm[iPm - 1]);
    w1 = (T.pm[iPm] - Pm) / (T.pm[iPm] - T.pm[iPm - 1]);
  }
  const double* c1 = T.val + (size_t)(iPm - 1) * numEm;      // sfval(:, iPm), sfval(:, iPm+1)
  const double* c2 = c1 + numEm;
  auto yval = [&](int iEm) {                                 // what sf_lookup returns for Em = Emval(iEm)
    double v = w1 * c1[iEm - 1] + w2 * c2[iEm - 1];
    if (iEm == numEm) {
      const double sf1 = w1 * c1[numEm - 2] + w2 * c2[numEm - 2];
      const double Em1 = T.em[numEm - 2], Em2 = T.em[numEm - 1];
      v = (sf1 + (Em2 - Em1) * (v - sf1) / (Em2 - Em1));
    }
    if (v < 1.e-20) v = 0;
    return v;
  };
  double total = yval(1);
  for (int iEm = 2; iEm <= numEm; ++iEm) total = yval(iEm) + total;
  double cum = yval(1);
  int ind = 1;
  while (u1 > cum / total) {           // NaN (no strength at this Pm) compares false: ind = 1, as in the reference
    ind = ind + 1;
    cum = yval(ind) + cum;
  }
  return T.em[ind - 1] + T.dem[ind - 1] * (u2 - 0.5);
}

// Independent-particle spectral function of COMMON /theory/ (simulate.inc:116-131) after theory_init
// (init.f:828-905).  buf: 8 doubles per shell { nprot*absorption, Em, Emsig, Em_int, Pm min, Pm bin, n, offset of
// rho in buf }, then the distributions rho_i(Pm) / bs_norm_i.
struct TheoryDev {
  const double* buf;
  int nrho;
  double e_fermi;
};
// event.f:1402-1428: linear interpolation of rho_i(Pm); Lorentzian in Em above E_Fermi for A > 2
SIMC_HD_CALL double theory_sf_weight(const TheoryDev T, bool heavy, double Em, double Pm) {
  const double pi = 3.141592653589793;
  double SF_weight = 0.0;
  for (int i = 0; i < T.nrho; ++i) {
    const double* sh = T.buf + 8 * i;
    const int n = (int)sh[6];
    double weight = 0.0;
    const double r = (Pm - sh[4]) / sh[5];
    if (r >= 0 && r <= n) {
      int iPm1 = (int)round(r);                 // nint: half away from zero
      if (iPm1 == 0) iPm1 = 1;
      if (iPm1 == n) iPm1 = n - 1;
      const double frac = r + 0.5 - (double)iPm1;
      const double* rho = T.buf + (long long)sh[7];
      const double b = rho[iPm1 - 1];
      const double a = rho[iPm1] - b;
      weight = a * frac + b;
    }
    if (heavy) {
      const double width = sh[2] / 2.0;
      if (Em < T.e_fermi) weight = 0.0;
      const double dE = Em - sh[1];
      weight = weight / pi / sh[3] * width / (dE * dE + width * width);
    }
    SF_weight = SF_weight + weight * sh[0];
  }
  return SF_weight;
}

// sf_lookup.f:97-170 (1-based indices of the Fortran kept in the helpers)
SIMC_HD_CALL double sf_lookup(const SfDev T, double Em, double Pm, bool& bad) {
  const int numPm = T.n_pm, numEm = T.n_em;
#define SF_PM(i) T.pm[(i) - 1]
#define SF_EM(i) T.em[(i) - 1]
#define SF_VAL(iEm, iPm) T.val[(size_t)((iPm) - 1) * numEm + ((iEm) - 1)]
  int iPm;
  double w1, w2;
  if (Pm >= SF_PM(numPm)) {
    iPm = numPm - 1; w1 = 0; w2 = 1;
  } else if (Pm <= SF_PM(1)) {
    iPm = 1; w1 = 1; w2 = 0;
  } else if (Pm > SF_PM(1)) {
    int ind = 1;
    while (Pm > SF_PM(ind)) ind = ind + 1;
    iPm = ind - 1;
    w2 = (Pm - SF_PM(iPm)) / (SF_PM(iPm + 1) - SF_PM(iPm));
    w1 = (SF_PM(iPm + 1) - Pm) / (SF_PM(iPm + 1) - SF_PM(iPm));
    if (fabs(w1 * SF_PM(iPm) + w2 * SF_PM(iPm + 1) - Pm) > 0.0001) bad = true;     // `stop` in the reference
  } else {                    // NaN: the reference's do-while would run off the table
    bad = true;
    return 0.0;
  }
  if (fabs(w1 + w2 - 1) > 0.0001) bad = true;
  double Em1 = 0, Em2 = 1, sf1 = 0, sf2 = 0;
  if (Em <= SF_EM(1)) {
    Em1 = SF_EM(1); Em2 = SF_EM(2);
    sf1 = w1 * SF_VAL(1, iPm) + w2 * SF_VAL(1, iPm + 1);
    sf2 = w1 * SF_VAL(2, iPm) + w2 * SF_VAL(2, iPm + 1);
  } else if (Em > SF_EM(numEm)) {
    Em1 = SF_EM(numEm - 1); Em2 = SF_EM(numEm);
    sf1 = w1 * SF_VAL(numEm - 1, iPm) + w2 * SF_VAL(numEm - 1, iPm + 1);
    sf2 = w1 * SF_VAL(numEm, iPm) + w2 * SF_VAL(numEm, iPm + 1);
  } else {
    for (int iEm = 1; iEm <= numEm - 1; ++iEm) {
      if (Em >= SF_EM(iEm) && Em < SF_EM(iEm + 1)) {
        Em1 = SF_EM(iEm); Em2 = SF_EM(iEm + 1);
        sf1 = w1 * SF_VAL(iEm, iPm) + w2 * SF_VAL(iEm, iPm + 1);
        sf2 = w1 * SF_VAL(iEm + 1, iPm) + w2 * SF_VAL(iEm + 1, iPm + 1);
      }
    }
  }
#undef SF_PM
#undef SF_EM
#undef SF_VAL
  double SF = (sf1 + (Em - Em1) * (sf2 - sf1) / (Em2 - Em1));
  if (SF < 1.e-20) SF = 0;
  return SF;
}

// sf_lookup.f:85-95
SIMC_HD_CALL double sf_lookup_diff(const SfDev T, double Em, double Pm, bool& bad) {
  const double SF = sf_lookup(T, Em, Pm, bad);
  return SF / 4 / 3.1415926535 / (Pm * Pm) / 5.0 / 20.0;
}

struct HeavyEv {              // what deForest reads from an `event` record
  double Q2, q, nu, Pm, Pmx, Pmy, Pmz, pE, pP, eE, etheta;
  double uqx, uqy, uqz, upx, upy, upz;
};

// physics_proton.f:23-135
SIMC_HD_CALL double deForest(const HeavyEv& ev, double Mh2, int deForest_flag) {
  const double hbarc = 197.327053, Mp2 = 938.27231 * 938.27231, alpha = 1. / 137.0359895;
  const double q4sq = -ev.Q2;
  const double q2 = ev.q * ev.q;
  double ebar, qbsq;
  if (deForest_flag >= 0) {
    ebar = sqrt(ev.Pm * ev.Pm + Mh2);
    qbsq = (ev.pE - ebar) * (ev.pE - ebar) - q2;
  } else {
    ebar = ev.pE - ev.nu;
    qbsq = q4sq;
  }
  const double cg = ev.uqx * ev.upx + ev.uqy * ev.upy + ev.uqz * ev.upz;
  double sin_gamma = 1. - cg * cg;
  if (sin_gamma < 0) sin_gamma = 0.0;
  sin_gamma = sqrt(sin_gamma);
  double cos_phi = 0.0;
  if (sin_gamma != 0)
    cos_phi = (ev.uqy * (ev.uqy * ev.upz - ev.uqz * ev.upy) - ev.uqx * (ev.uqz * ev.upx - ev.uqx * ev.upz)) / sin_gamma /
              sqrt(1. - ev.uqz * ev.uqz);
  if (fabs(cos_phi) > 1.) cos_phi = copysign(1.0, cos_phi);
  // fofa_best_fit(q4sq/hbarc**2), physics_proton.f:137-172
  const double qsquar = q4sq / (hbarc * hbarc);
  const double Q2g = -qsquar * (hbarc * hbarc) * 1.e-6;        // hbarc**2. : pow(x,2.) == x*x
  const double Q = sqrt(fmax(Q2g, 0.e0));
  const double Q3 = m::pow(Q, 3.), Q4 = m::pow(Q, 4.), Q5 = m::pow(Q, 5.);
  double denom = 1. + 0.62 * Q + 0.68 * Q2g + 2.8 * Q3 + 0.83 * Q4;
  const double GE = 1. / denom;
  denom = 1. + 0.35 * Q + 2.44 * Q2g + 0.5 * Q3 + 1.04 * Q4 + 0.34 * Q5;
  const double GM = 2.793 / denom;
  const double qmu4mp = q4sq / 4. / Mp2;
  const double f1 = (GE - GM * qmu4mp) / (1.0 - qmu4mp);
  const double kf2 = (GM - GE) / (1.0 - qmu4mp);
  const double f1sq = f1 * f1;
  const double kf2_over_2m_allsq = kf2 * kf2 / 4. / Mh2;
  const double th2 = m::tan(ev.etheta / 2.);
  const double qq = q4sq / q2;
  const double termC = qq * qq;
  const double termT = th2 * th2 - q4sq / 2. / q2;
  const double termS = th2 * th2 - (q4sq / q2) * (cos_phi * cos_phi);
  const double termI = (-q4sq / q2) * sqrt(th2 * th2 - q4sq / q2) * cos_phi;
  double WC, WT, WS, WI;
  if (deForest_flag <= 0) {
    const double sumFF1 = (f1 + kf2) * (f1 + kf2);
    const double sumFF2 = f1sq - qbsq * kf2 * kf2 / 4. / Mh2;
    WC = ((ebar + ev.pE) * (ebar + ev.pE)) * sumFF2 - q2 * sumFF1;
    WT = -2 * qbsq * sumFF1;
    WS = 4 * (ev.pP * ev.pP) * (sin_gamma * sin_gamma) * sumFF2;
    WI = -4 * (ebar + ev.pE) * ev.pP * sin_gamma * sumFF2;
  } else {
    const double pbarp = ebar * ev.pE - ev.pP * (ev.upx * ev.Pmx + ev.upy * ev.Pmy + ev.upz * ev.Pmz);
    const double pbarq = ebar * ev.nu - ev.q * (ev.uqx * ev.Pmx + ev.uqy * ev.Pmy + ev.uqz * ev.Pmz);
    const double pq = ev.pE * ev.nu - ev.pP * ev.q * (ev.upx * ev.uqx + ev.upy * ev.uqy + ev.upz * ev.uqz);
    const double qbarq = (ev.pE - ebar) * ev.nu - q2;
    WC = (ebar * ev.pE + (-pbarp + Mh2) / 2.) * f1sq - q2 * f1 * kf2 / 2. -
         ((-pbarq * ev.pE - pq * ebar) * ev.nu + ebar * ev.pE * q4sq + pbarq * pq - (-pbarp - Mh2) / 2. * q2) *
             kf2_over_2m_allsq;
    WT = -(-pbarp + Mh2) * f1sq - qbarq * f1 * kf2 + (2. * pbarq * pq + (-pbarp - Mh2) * q4sq) * kf2_over_2m_allsq;
    WS = (ev.pP * sin_gamma) * (ev.pP * sin_gamma) * (f1sq - q4sq * kf2_over_2m_allsq);
    WI = ev.pP * sin_gamma * (-(ebar + ev.pE) * f1sq + ((-pbarq - pq) * ev.nu + (ebar + ev.pE) * q4sq) * kf2_over_2m_allsq);
  }
  double allsum = termC * WC + termT * WT + termS * WS + termI * WI;
  if (deForest_flag <= 0) allsum = allsum / 4.0;
  const double mott = 2. * alpha * hbarc * ev.eE * m::cos(ev.etheta / 2.) / ev.Q2;
  const double sigMott = (mott * mott) * 1.e4;
  return sigMott * ev.pP * allsum / ebar;
}

}  // namespace simc
