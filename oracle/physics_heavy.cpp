// ORACLE -- TEST INFRASTRUCTURE ONLY (see event.hpp).  PARITY UNPINNED.
// A(e,e'p) weights: Benhar spectral-function lookup (sf_lookup.f:85-170) and the deForest
// off-shell cross sections sigma_cc1 / sigma_cc2 (physics_proton.f:23-135).
#include <cmath>
#include <stdexcept>

#include "event.hpp"

namespace simc_oracle {

// sf_lookup.f:97-170.  1-based indices like the Fortran.
double sf_lookup(const SfTable& T, double Em, double Pm) {
  SfLookupState st;
  return sf_lookup_state(T, Em, Pm, st);
}
double sf_lookup_state(const SfTable& T, double Em, double Pm, SfLookupState& st) {
  const int numPm = T.numPm, numEm = T.numEm;
  auto Pmval = [&](int i) { return T.Pmval[i - 1]; };
  auto Emval = [&](int i) { return T.Emval[i - 1]; };
  auto sfval = [&](int iEm, int iPm) { return T.sfval[(size_t)(iPm - 1) * numEm + (iEm - 1)]; };
  int iPm;
  double w1, w2;
  if (Pm >= Pmval(numPm)) {
    iPm = numPm - 1; w1 = 0; w2 = 1;
  } else if (Pm <= Pmval(1)) {
    iPm = 1; w1 = 1; w2 = 0;
  } else {
    int ind = 1;
    while (Pm > Pmval(ind)) ind = ind + 1;
    iPm = ind - 1;
    w2 = (Pm - Pmval(iPm)) / (Pmval(iPm + 1) - Pmval(iPm));
    w1 = (Pmval(iPm + 1) - Pm) / (Pmval(iPm + 1) - Pmval(iPm));
    if (std::fabs(w1 * Pmval(iPm) + w2 * Pmval(iPm + 1) - Pm) > 0.0001) throw std::runtime_error("sf_lookup: bad Pm weights");
  }
  if (std::fabs(w1 + w2 - 1) > 0.0001) throw std::runtime_error("sf_lookup: w1+w2 != 1");
  double &Em1 = st.Em1, &Em2 = st.Em2, &sf1 = st.sf1, &sf2 = st.sf2;
  if (Em <= Emval(1)) {
    Em1 = Emval(1); Em2 = Emval(2);
    sf1 = w1 * sfval(1, iPm) + w2 * sfval(1, iPm + 1);
    sf2 = w1 * sfval(2, iPm) + w2 * sfval(2, iPm + 1);
  } else if (Em > Emval(numEm)) {
    Em1 = Emval(numEm - 1); Em2 = Emval(numEm);
    sf1 = w1 * sfval(numEm - 1, iPm) + w2 * sfval(numEm - 1, iPm + 1);
    sf2 = w1 * sfval(numEm, iPm) + w2 * sfval(numEm, iPm + 1);
  } else {
    for (int iEm = 1; iEm <= numEm - 1; ++iEm) {
      if (Em >= Emval(iEm) && Em < Emval(iEm + 1)) {
        Em1 = Emval(iEm); Em2 = Emval(iEm + 1);
        sf1 = w1 * sfval(iEm, iPm) + w2 * sfval(iEm, iPm + 1);
        sf2 = w1 * sfval(iEm + 1, iPm) + w2 * sfval(iEm + 1, iPm + 1);
      }
    }
  }
  const double logsf = (sf1 + (Em - Em1) * (sf2 - sf1) / (Em2 - Em1));
  double SF = logsf;
  if (SF < 1.e-20) SF = 0;
  return SF;
}

// generate_em, sf_lookup.f:181-245: missing energy from the spectral function's Em distribution at fixed Pm
double generate_em(const SfTable& T, Rng& rng, double Pm) {
  const int numEm = T.numEm;
  if ((int)T.dEm.size() != numEm) throw std::runtime_error("oracle: generate_em needs the Em bin widths");
  std::vector<double> y(numEm + 1);
  SfLookupState st;
  y[1] = sf_lookup_state(T, T.Emval[0], Pm, st);
  for (int iEm = 2; iEm <= numEm; ++iEm) {
    y[iEm] = sf_lookup_state(T, T.Emval[iEm - 1], Pm, st);
    y[iEm] = y[iEm] + y[iEm - 1];
  }
  for (int iEm = 1; iEm <= numEm; ++iEm) y[iEm] = y[iEm] / y[numEm];
  const double ranprob = rng.grnd();
  int ind = 1;
  while (ranprob > y[ind]) ind = ind + 1;
  return T.Emval[ind - 1] + T.dEm[ind - 1] * (rng.grnd() - 0.5);
}

// sf_lookup.f:85-95
double sf_lookup_diff(const SfTable& T, double Em, double Pm) {
  const double SF = sf_lookup(T, Em, Pm);
  return SF / 4 / 3.1415926535 / (Pm * Pm) / 5.0 / 20.0;
}

namespace {
void fofa_best_fit(double qsquar, double& GE, double& GM) {     // physics_proton.f:137-172
  const double mu_p = 2.793;
  const double Q2 = -qsquar * std::pow(K::hbarc, 2.) * 1.e-6;
  const double Q = std::sqrt(std::max(Q2, 0.e0));
  const double Q3 = std::pow(Q, 3.), Q4 = std::pow(Q, 4.), Q5 = std::pow(Q, 5.);
  double denom = 1. + 0.62 * Q + 0.68 * Q2 + 2.8 * Q3 + 0.83 * Q4;
  GE = 1. / denom;
  denom = 1. + 0.35 * Q + 2.44 * Q2 + 0.5 * Q3 + 1.04 * Q4 + 0.34 * Q5;
  GM = mu_p / denom;
}
double sigMott(double e0, double theta, double Q2) {            // physics_proton.f:176-190
  const double sig = powi(2. * K::alpha * K::hbarc * e0 * std::cos(theta / 2.) / Q2, 2);
  return sig * 1.e4;
}
}  // namespace

// physics_proton.f:23-135
double deForest(const simc_run_config& cfg, const Event& ev) {
  const double Mh2 = cfg.Mh2;
  const int deForest_flag = cfg.deForest_flag;
  const double q4sq = -ev.Q2;
  const double q2 = ev.q * ev.q;
  double ebar, qbsq;
  if (deForest_flag >= 0) {
    ebar = std::sqrt(ev.Pm * ev.Pm + Mh2);
    qbsq = powi(ev.p.E - ebar, 2) - q2;
  } else {
    ebar = ev.p.E - ev.nu;
    qbsq = q4sq;
  }
  double sin_gamma = 1. - powi(ev.uq.x * ev.up.x + ev.uq.y * ev.up.y + ev.uq.z * ev.up.z, 2);
  if (sin_gamma < 0) sin_gamma = 0.0;
  sin_gamma = std::sqrt(sin_gamma);
  double cos_phi = 0.0;
  if (sin_gamma != 0)
    cos_phi = (ev.uq.y * (ev.uq.y * ev.up.z - ev.uq.z * ev.up.y) - ev.uq.x * (ev.uq.z * ev.up.x - ev.uq.x * ev.up.z)) /
              sin_gamma / std::sqrt(1. - ev.uq.z * ev.uq.z);
  if (std::fabs(cos_phi) > 1.) cos_phi = std::copysign(1.0, cos_phi);
  double GE, GM;
  fofa_best_fit(q4sq / (K::hbarc * K::hbarc), GE, GM);
  const double qmu4mp = q4sq / 4. / K::Mp2;
  const double f1 = (GE - GM * qmu4mp) / (1.0 - qmu4mp);
  const double kf2 = (GM - GE) / (1.0 - qmu4mp);
  const double f1sq = f1 * f1;
  const double kf2_over_2m_allsq = kf2 * kf2 / 4. / Mh2;
  const double th2 = std::tan(ev.e.theta / 2.);
  const double termC = powi(q4sq / q2, 2);
  const double termT = powi(th2, 2) - q4sq / 2. / q2;
  const double termS = powi(th2, 2) - (q4sq / q2) * (cos_phi * cos_phi);
  const double termI = (-q4sq / q2) * std::sqrt(powi(th2, 2) - q4sq / q2) * cos_phi;
  double WC, WT, WS, WI;
  if (deForest_flag <= 0) {
    const double sumFF1 = powi(f1 + kf2, 2);
    const double sumFF2 = f1sq - qbsq * kf2 * kf2 / 4. / Mh2;
    WC = (powi(ebar + ev.p.E, 2)) * sumFF2 - q2 * sumFF1;
    WT = -2 * qbsq * sumFF1;
    WS = 4 * (ev.p.P * ev.p.P) * (sin_gamma * sin_gamma) * sumFF2;
    WI = -4 * (ebar + ev.p.E) * ev.p.P * sin_gamma * sumFF2;
  } else {
    const double pbarp = ebar * ev.p.E - ev.p.P * (ev.up.x * ev.Pmx + ev.up.y * ev.Pmy + ev.up.z * ev.Pmz);
    const double pbarq = ebar * ev.nu - ev.q * (ev.uq.x * ev.Pmx + ev.uq.y * ev.Pmy + ev.uq.z * ev.Pmz);
    const double pq = ev.p.E * ev.nu - ev.p.P * ev.q * (ev.up.x * ev.uq.x + ev.up.y * ev.uq.y + ev.up.z * ev.uq.z);
    const double qbarq = (ev.p.E - ebar) * ev.nu - q2;
    WC = (ebar * ev.p.E + (-pbarp + Mh2) / 2.) * f1sq - q2 * f1 * kf2 / 2. -
         ((-pbarq * ev.p.E - pq * ebar) * ev.nu + ebar * ev.p.E * q4sq + pbarq * pq - (-pbarp - Mh2) / 2. * q2) *
             kf2_over_2m_allsq;
    WT = -(-pbarp + Mh2) * f1sq - qbarq * f1 * kf2 + (2. * pbarq * pq + (-pbarp - Mh2) * q4sq) * kf2_over_2m_allsq;
    WS = powi(ev.p.P * sin_gamma, 2) * (f1sq - q4sq * kf2_over_2m_allsq);
    WI = ev.p.P * sin_gamma *
         (-(ebar + ev.p.E) * f1sq + ((-pbarq - pq) * ev.nu + (ebar + ev.p.E) * q4sq) * kf2_over_2m_allsq);
  }
  double allsum = termC * WC + termT * WT + termS * WS + termI * WI;
  if (deForest_flag <= 0) allsum = allsum / 4.0;
  return sigMott(ev.e.E, ev.e.theta, ev.Q2) * ev.p.P * allsum / ebar;
}

}  // namespace simc_oracle
