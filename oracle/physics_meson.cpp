// ORACLE -- TEST INFRASTRUCTURE ONLY (see event.hpp).  PARITY UNPINNED.
// Meson electroproduction weights: transform_to_cm (jacobians.f:1-282), peepi + sig_param_2021 +
// exclfit (physics_pion.f:1-130, 738-854), peeK + sig_factorized (physics_kaon.f:1-236).
// Hydrogen targets: pfer = 0, pferx/y/z = 0, efer = Mtar_struck (event.f:330-335); the Fermi
// terms are kept in the formulas so that the arithmetic is the reference's.
//
// The MAID-2007 branch of peepi for W < 2 GeV (sigmaid, physics_pion.f:131-154, 577-728) needs the caller's
// table; without it such events are counted in simc_accum.unsupported.  The Saghai model eekeek / eekeeks of peeK
// (physics_kaon.f:241-489 with CERNLIB's fint, cern/fint.f) only fills the ntuple column sigcm1, never the weight
// (physics_kaon.f:100-115); it is evaluated when its tables are set (oracle_set_saghai_table).
//
// Types of fint under the reference's flags (Makefile:63, -fdefault-real-8): ARG, ENT and TABLE are REAL*4 as
// declared; WEIGHT, X, H, ETA and the function result are default REAL, i.e. 8 bytes.  The restatement follows the
// source as written.  (As BUILT, eekeek declares `real*4 fint` while the function returns a default REAL: on x86-64
// the caller then reads the low half of a double as a float.  Like the HRS REAL(16) literals of SURVEY A.5 this is
// recorded, not reproduced: the column is a diagnostic and the as-written value is the meaningful one.)
#include <algorithm>
#include <complex>
#include <stdexcept>

#include "event.hpp"

namespace simc_oracle {

namespace {
struct Fermi { double pfer = 0.0, pferx = 0.0, pfery = 0.0, pferz = 0.0, efer = 0.0; };

struct CmFrame {
  double gstar, bstar, bstarx, bstary, bstarz;
  double nustar, qstar, qstarx, qstary, qstarz;
  double ehadcm, phadcm, phadcmx, phadcmy, phadcmz;
  double ebeamcm, pbeamcm, pbeamcmx, pbeamcmy, pbeamcmz;
  double etarcm, ptarcm, ptarcmx, ptarcmy, ptarcmz;
  double thetacm, phicm, phiqn, jacobian, jac_old;
};

// jacobians.f:1-282
void transform_to_cm(const Event& vertex, const EventMain& main, const Fermi& F, CmFrame& C) {
  const double pi = K::pi;
  const double pfer = F.pfer, pferx = F.pferx, pfery = F.pfery, pferz = F.pferz, efer = F.efer;
  double tcos = vertex.up.x * vertex.uq.x + vertex.up.y * vertex.uq.y + vertex.up.z * vertex.uq.z;
  if (tcos - 1. > 0. && tcos - 1. < 1.e-8) tcos = 1.0;
  const double tsin = sqrt(1. - tcos * tcos);
  double tfcos = pferx * vertex.uq.x + pfery * vertex.uq.y + pferz * vertex.uq.z;
  if (tfcos - 1. > 0. && tfcos - 1. < 1.e-8) tfcos = 1.0;
  const double tfsin = sqrt(1. - tfcos * tfcos);
  const double cospq = cos(main.phi_pq), sinpq = sin(main.phi_pq);
  const double qx = -vertex.uq.y, qy = vertex.uq.x, qz = vertex.uq.z;
  const double px = -pfery, py = pferx, pz = pferz;
  double dummy = sqrt((qx * qx + qy * qy) * (qx * qx + qy * qy + qz * qz));
  const double tmp_x_x = -qx * qz / dummy, tmp_x_y = -qy * qz / dummy, tmp_x_z = (qx * qx + qy * qy) / dummy;
  dummy = sqrt(qx * qx + qy * qy);
  const double tmp_y_x = qy / dummy, tmp_y_y = -qx / dummy, tmp_y_z = 0.0;
  const double p_tmp_x = pfer * (px * tmp_x_x + py * tmp_x_y + pz * tmp_x_z);
  const double p_tmp_y = pfer * (px * tmp_y_x + py * tmp_y_y + pz * tmp_y_z);
  if (p_tmp_x == 0.) C.phiqn = 0.;
  else C.phiqn = atan2(p_tmp_y, p_tmp_x);
  if (C.phiqn < 0.) C.phiqn = C.phiqn + 2. * pi;
  const double phiqn = C.phiqn;

  const double pbeam = vertex.Ein;
  const double beam_tmpx = pbeam * tmp_x_z, beam_tmpy = pbeam * tmp_y_z, beam_tmpz = pbeam * vertex.uq.z;
  C.bstar = sqrt(powi(vertex.q + pfer * tfcos, 2) + powi(pfer * tfsin, 2)) / (efer + vertex.nu);
  C.gstar = 1. / sqrt(1. - C.bstar * C.bstar);
  C.bstarz = (vertex.q + pfer * tfcos) / (efer + vertex.nu);
  C.bstarx = p_tmp_x / (efer + vertex.nu);
  C.bstary = p_tmp_y / (efer + vertex.nu);
  const double gstar = C.gstar, bstar = C.bstar, bstarx = C.bstarx, bstary = C.bstary, bstarz = C.bstarz;

  loren(gstar, bstarx, bstary, bstarz, vertex.Ein, beam_tmpx, beam_tmpy, beam_tmpz, C.ebeamcm, C.pbeamcmx, C.pbeamcmy,
        C.pbeamcmz, C.pbeamcm);
  const double zero = 0.e0;
  loren(gstar, bstarx, bstary, bstarz, vertex.nu, zero, zero, vertex.q, C.nustar, C.qstarx, C.qstary, C.qstarz, C.qstar);
  const double phadz = vertex.p.P * tcos, phadx = vertex.p.P * tsin * cospq, phady = vertex.p.P * tsin * sinpq;
  loren(gstar, bstarx, bstary, bstarz, vertex.p.E, phadx, phady, phadz, C.ehadcm, C.phadcmx, C.phadcmy, C.phadcmz,
        C.phadcm);
  C.thetacm = acos((C.phadcmx * C.qstarx + C.phadcmy * C.qstary + C.phadcmz * C.qstarz) / C.phadcm / C.qstar);
  const double ptarz = pfer * tfcos, ptarx = p_tmp_x, ptary = p_tmp_y;
  loren(gstar, bstarx, bstary, bstarz, efer, ptarx, ptary, ptarz, C.etarcm, C.ptarcmx, C.ptarcmy, C.ptarcmz, C.ptarcm);
  const double qstarx = C.qstarx, qstary = C.qstary, qstarz = C.qstarz, qstar = C.qstar;
  const double pbeamcmx = C.pbeamcmx, pbeamcmy = C.pbeamcmy, pbeamcmz = C.pbeamcmz;
  const double phadcmx = C.phadcmx, phadcmy = C.phadcmy, phadcmz = C.phadcmz;

  dummy = sqrt(powi(qstary * pbeamcmz - qstarz * pbeamcmy, 2) + powi(qstarz * pbeamcmx - qstarx * pbeamcmz, 2) +
               powi(qstarx * pbeamcmy - qstary * pbeamcmx, 2));
  const double tmp2_y_x = (qstary * pbeamcmz - qstarz * pbeamcmy) / dummy;
  const double tmp2_y_y = (qstarz * pbeamcmx - qstarx * pbeamcmz) / dummy;
  const double tmp2_y_z = (qstarx * pbeamcmy - qstary * pbeamcmx) / dummy;
  dummy = sqrt(powi(tmp2_y_y * qstarz - tmp2_y_z * qstary, 2) + powi(tmp2_y_z * qstarx - tmp2_y_x * qstarz, 2) +
               powi(tmp2_y_x * qstary - tmp2_y_y * qstarx, 2));
  const double tmp2_x_x = (tmp2_y_y * qstarz - tmp2_y_z * qstary) / dummy;
  const double tmp2_x_y = (tmp2_y_z * qstarx - tmp2_y_x * qstarz) / dummy;
  const double tmp2_x_z = (tmp2_y_x * qstary - tmp2_y_y * qstarx) / dummy;
  const double tmp2_z_x = qstarx / qstar, tmp2_z_y = qstary / qstar, tmp2_z_z = qstarz / qstar;
  const double phadcm_tmp2x = phadcmx * tmp2_x_x + phadcmy * tmp2_x_y + phadcmz * tmp2_x_z;
  const double phadcm_tmp2y = phadcmx * tmp2_y_x + phadcmy * tmp2_y_y + phadcmz * tmp2_y_z;
  C.phicm = atan2(phadcm_tmp2y, phadcm_tmp2x);
  if (C.phicm < 0.) C.phicm = 2. * pi + C.phicm;

  // Jacobian dt dphi_cm -> dOmega_lab, jacobians.f:180-262
  const double P = vertex.p.P, E = vertex.p.E;
  const double psign = cos(phiqn) * cospq + sin(phiqn) * sinpq;
  const double square_root = vertex.q + pfer * tfcos - P * tcos;
  const double dp_dcos_num = P + (P * P * tcos - psign * pfer * P * tfsin * tcos / tsin) / square_root;
  const double dp_dcos_den =
      ((vertex.nu + efer - E) * P / E + P * tsin * tsin - psign * pfer * tfsin * tsin) / square_root - tcos;
  const double dp_dcos = dp_dcos_num / dp_dcos_den;
  const double dp_dphi_num = pfer * P * tsin * tfsin * (cos(phiqn) * sinpq - sin(phiqn) * cospq) / square_root;
  const double dp_dphi_den =
      tcos + (pfer * tsin * tfsin * psign - P * tsin * tsin - (vertex.nu + efer - E) * P / E) / square_root;
  const double dp_dphi = dp_dphi_num / dp_dphi_den;
  const double dt_dcos_lab = 2. * (vertex.q * P + (vertex.q * tcos - vertex.nu * P / E) * dp_dcos);
  const double dt_dphi_lab = 2. * (vertex.q * tcos - vertex.nu * P / E) * dp_dphi;

  const double dpxdphi = P * tsin * (-sinpq + (gstar - 1.) * bstarx / (bstar * bstar) * (bstary * cospq - bstarx * sinpq)) +
                         ((phadcmx + gstar * bstarx * E) / P - gstar * bstarx * P / E) * dp_dphi;
  const double dpydphi = P * tsin * (cospq + (gstar - 1.) * bstary / (bstar * bstar) * (bstary * cospq - bstarx * sinpq)) +
                         ((phadcmy + gstar * bstary * E) / P - gstar * bstary * P / E) * dp_dphi;
  const double dpzdphi = P * (gstar - 1.) / (bstar * bstar) * bstarz * tsin * (bstary * cospq - bstarx * sinpq) +
                         ((phadcmz + gstar * bstarz * E) / P - gstar * bstarz * P / E) * dp_dphi;
  const double dpxdcos =
      -P * tcos / tsin *
          (cospq + (gstar - 1.) * bstarx / (bstar * bstar) * (bstarx * cospq + bstary * sinpq - bstarz * tsin / tcos)) +
      ((phadcmx + gstar * bstarx * E) / P - gstar * bstarx * P / E) * dp_dcos;
  const double dpydcos =
      -P * tcos / tsin *
          (sinpq + (gstar - 1.) * bstary / (bstar * bstar) * (bstarx * cospq + bstary * sinpq - bstarz * tsin / tcos)) +
      ((phadcmy + gstar * bstary * E) / P - gstar * bstary * P / E) * dp_dcos;
  const double dpzdcos =
      P * (1. - (gstar - 1.) / (bstar * bstar) * bstarz * tcos / tsin *
                    (bstarx * cospq + bstary * sinpq - tsin / tcos * bstarz)) +
      ((phadcmz + gstar * bstarz * E) / P - gstar * bstarz * P / E) * dp_dcos;

  const double dpxnewdphi = dpxdphi * tmp2_x_x + dpydphi * tmp2_x_y + dpzdphi * tmp2_x_z;
  const double dpynewdphi = dpxdphi * tmp2_y_x + dpydphi * tmp2_y_y + dpzdphi * tmp2_y_z;
  const double dphicmdphi = (dpynewdphi * phadcm_tmp2x - phadcm_tmp2y * dpxnewdphi) /
                            (phadcm_tmp2x * phadcm_tmp2x + phadcm_tmp2y * phadcm_tmp2y);
  const double dpxnewdcos = dpxdcos * tmp2_x_x + dpydcos * tmp2_x_y + dpzdcos * tmp2_x_z;
  const double dpynewdcos = dpxdcos * tmp2_y_x + dpydcos * tmp2_y_y + dpzdcos * tmp2_y_z;
  const double dphicmdcos = (dpynewdcos * phadcm_tmp2x - phadcm_tmp2y * dpxnewdcos) /
                            (phadcm_tmp2x * phadcm_tmp2x + phadcm_tmp2y * phadcm_tmp2y);
  C.jacobian = fabs(dt_dcos_lab * dphicmdphi - dt_dphi_lab * dphicmdcos);
  C.jac_old = 2 * (efer - 2 * pferz * pfer * E / P * tcos) * (vertex.q + pferz * pfer) * P /
                  (efer + vertex.nu - (vertex.q + pferz * pfer) * E / P * tcos) -
              2 * P * pfer;
}

// physics_pion.f:807-854
double exclfit(double t, double thetacm, double phicm, double q2_gev, double s_gev, double eps, const double* pp,
               double fpifact) {
  const double* p = pp - 1;   // 1-based like the Fortran array
  const double mtar_gev = 0.938;
  const double fpi = fpifact / (1.0 + p[1] * q2_gev + p[2] * q2_gev * q2_gev);
  const double q2fpi2 = q2_gev * (fpi * fpi);
  double sigL = (p[3] + p[15] / q2_gev) * fabs(t) / powi(fabs(t) + 0.02, 2) * q2fpi2 * exp(p[4] * fabs(t));
  sigL = sigL / (pow(s_gev, p[11]) + pow(sqrt(s_gev), p[17]));
  double sigT = p[5] / q2_gev * exp(p[6] * (q2_gev * q2_gev));
  sigT = sigT / (pow(s_gev, p[12]) + pow(sqrt(s_gev), p[16]));
  sigT = sigT * exp(p[14] * fabs(t));
  double sigLT = (p[7] / (1.0 + p[10] * q2_gev)) * exp(p[8] * fabs(t)) * sin(thetacm);
  sigLT = sigLT / pow(s_gev, p[13]);
  const double sigTT = (p[9] / (1. + 1.0 * q2_gev)) * exp(-7.0 * fabs(t)) * powi(sin(thetacm), 2);
  const double sig219 =
      (sigT + eps * sigL + eps * cos(2.0 * phicm) * sigTT + sqrt(2.0 * eps * (1.0 + eps)) * cos(phicm) * sigLT) / 1.0;
  double sig = sig219 * 8.539 / powi(s_gev - mtar_gev * mtar_gev, 2);
  sig = sig / 2.0 / 3.1415928 / 1.0e+06;
  return sig;
}

// physics_pion.f:738-805 (charged pions)
double sig_param_2021(double thcm, double phicm, double t, double q2, double wsq, double eps, int which_pion) {
  static const double pp[17] = {1.60077, -0.01523, 37.08142, -4.11060, 23.26192, 0.00983, 0.87073, -5.77115, -271.08678,
                                0.13766, -0.00855, 0.27885,  -1.13212, -1.50415, -6.34766, 0.55769, -0.01709};
  static const double pm[17] = {1.75169, 0.11144, 47.35877, -4.69434, 1.60552, 0.00800, 0.44194, -2.29188, -41.67194,
                                0.69475, 0.02527, -0.50178, -1.22825, -1.16878, 5.75825, -1.00355, 0.05055};
  if (which_pion == 1 || which_pion == 11 || which_pion == 3) return exclfit(t, thcm, phicm, q2, wsq, eps, pm, 1.0);
  return exclfit(t, thcm, phicm, q2, wsq, eps, pp, 1.0);
}

// physics_kaon.f:175-236
double sig_factorized(double q2, double w, double t, double pk, double mrec) {
  const double Mp = K::Mp, Mk2 = K::Mk2;
  const double nu = (w * w + q2 - Mp * Mp) / 2. / Mp;
  const double q = sqrt(q2 + nu * nu);
  const double qcm = q * (Mp / w);
  const double nucm = sqrt(qcm * qcm - q2);
  const double tmin = -1. * (Mk2 - q2 - 2 * nucm * sqrt(pk * pk + Mk2) + 2 * qcm * pk);
  const double q2val = q2 / 1.e6, w2val = w * w / 1.e6, pkval = pk / 1000., tval = t / 1.e6, tminval = tmin / 1.e6;
  double fact_q, fact_t, fact_w = 0.0;
  if (mrec < 1150.) {
    fact_q = 1. / powi(q2val + 2.67, 2);
    fact_t = exp(-2.1 * (tval - tminval));
    if (w2val != 0) {
      fact_w = 0.959 * 4.1959 * pkval / (sqrt(w2val) * (w2val - 0.93827 * 0.93827));
      fact_w = fact_w + (0.18 * (1.72 * 1.72) * (0.10 * 0.10)) /
                            (powi(w2val - 1.72 * 1.72, 2) + (1.72 * 1.72) * (0.10 * 0.10));
    }
  } else {
    fact_q = 1. / powi(q2val + 0.79, 2);
    fact_t = exp(-1.0 * (tval - tminval));
    if (w2val != 0) fact_w = 0.959 * 4.1959 * pkval / (sqrt(w2val) * (w2val - 0.93827 * 0.93827));
  }
  return fact_q * fact_t * fact_w;
}

// COMMON /pfermi_stuff/ as generate left it (event.f:327-373): at rest for hydrogen (efer = Mtar_struck)
Fermi fermi_of(const Sim& s) {
  Fermi F;
  F.pfer = s.pfer; F.pferx = s.pferx; F.pfery = s.pfery; F.pferz = s.pferz; F.efer = s.efer;
  return F;
}
}  // namespace

// sigmaid, physics_pion.f:577-728: nearest-bin lookup in the MAID-2007 table; only sig0 is used by peepi
double sigmaid_sig0(const MaidTable& M, int ipi, double q2, double w, double e0, double costh, double phi) {
  static const double cthmin[6] = {-0.20, 0.20, 0.44, 0.63, 0.78, 0.90};
  static const double cthmax[6] = {0.20, 0.44, 0.63, 0.78, 0.90, 1.0};
  const double am = 0.9383;
  const std::vector<double>& T = M.tbl[ipi - 3];
  double sig0 = 0.;
  if (w < 1.08) return sig0;
  const double nu = (w * w - am * am + q2) / 2. / am;
  if (nu > e0) return sig0;
  const double ep = e0 - nu;
  const double sin2 = q2 / 4. / e0 / ep;
  if (sin2 <= 0.0 || sin2 > 1.) return sig0;
  const double eps = 1. / (1. + 2. * (1. + nu * nu / q2) * sin2 / (1. - sin2));
  int iq = (int)((q2 + 0.1) / 0.2);
  iq = std::min(25, std::max(1, iq));
  int iw = (int)((w - 1.090) / 0.020);
  iw = std::min(46, std::max(1, iw));
  int ith = 0;
  for (int i = 1; i <= 6; ++i)
    if (costh >= cthmin[i - 1] && costh <= cthmax[i - 1]) ith = i;
  ith = std::min(6, std::max(1, ith));
  double wfact = 1.;
  if (w > 1.232) wfact = (w - 1.132) / 0.100;
  const double* row = &T[(((size_t)(iq - 1) * 46 + (iw - 1)) * 6 + (ith - 1)) * 4];
  const double ST = row[0] / std::max(0.2, q2) / wfact;
  const double SL = row[1] * ST;
  const double STL = row[2] * ST;
  const double STT = row[3] * ST;
  const double CSF = cos(phi);
  const double CS2F = cos(2. * phi);
  sig0 = ST + eps * SL + sqrt(2. * eps * (1. + eps)) * CSF * STL + eps * CS2F * STT;
  return sig0;
}

// physics_pion.f:1-130
double peepi(Sim& s, const Event& vertex, EventMain& main) {
  const simc_run_config& cfg = *s.cfg;
  const simc_target& targ = cfg.targ;
  const Fermi F = fermi_of(s);
  CmFrame C;
  transform_to_cm(vertex, main, F, C);
  main.thetacm = C.thetacm;
  main.phicm = C.phicm;
  main.pcm = C.phadcm;
  main.davejac = C.jacobian;
  main.johnjac = C.jac_old;
  double tfcos = F.pferx * vertex.uq.x + F.pfery * vertex.uq.y + F.pferz * vertex.uq.z;
  if (tfcos - 1. > 0. && tfcos - 1. < 1.e-8) tfcos = 1.0;
  const double tfsin = sqrt(1. - tfcos * tfcos);
  const double sgev = powi(vertex.nu + F.efer, 2) - powi(vertex.q + F.pfer * tfcos, 2) - powi(F.pfer * tfsin, 2);
  main.wcm = sqrt(sgev);
  const double k_eq = (main.wcm * main.wcm - targ.Mtar_struck * targ.Mtar_struck) / 2. / targ.Mtar_struck;
  s.ntup.sigcm1 =
      sig_param_2021(C.thetacm, C.phicm, main.t / 1.e6, vertex.Q2 / 1.e6, sgev / 1.e6, main.epsilon, cfg.which_pion);
  double sigma_eepi = s.ntup.sigcm1;
  if (main.wcm < 2000) {                    // physics_pion.f:131-154: blend with MAID-2007 below W = 2 GeV
    const int ipi = (cfg.which_pion == 1 || cfg.which_pion == 11 || cfg.which_pion == 3) ? 4 : 3;
    if (s.maid && !s.maid->tbl[ipi - 3].empty()) {
      const double Q2gev = vertex.Q2 / 1.e6, Wgev = main.wcm / 1000.0, cthcm = cos(C.thetacm), E0 = vertex.Ein / 1000.0;
      const double sig0 = sigmaid_sig0(*s.maid, ipi, Q2gev, Wgev, E0, cthcm, C.phicm);
      s.ntup.sigcm2 = sig0 / C.phadcm / C.qstar / 2.;
      const double fac1 = std::min(1., std::max(0., (Wgev - 1.5) / 0.4));
      sigma_eepi = s.ntup.sigcm1 * fac1 + s.ntup.sigcm2 * (1 - fac1);
    } else {
      s.low_w = true;                       // table not provided: parametrisation alone, counted in `unsupported`
    }
  }
  s.ntup.sigcm = sigma_eepi;
  const double fac = 1. / (1. - F.pferz * F.pfer / F.efer) * targ.Mtar_struck / F.efer;
  const double gtpr = K::alpha / 2. / (K::pi * K::pi) * vertex.e.E / vertex.Ein * k_eq / vertex.Q2 / (1. - main.epsilon);
  return sigma_eepi * C.jacobian * (gtpr * fac);
}

// physics_kaon.f:1-171
// cern/fint.f:10-76: multilinear interpolation (with linear extrapolation outside the grid) in up to 5 arguments.
// Indices are kept 1-based like the source.
double fint(int narg, const float* arg, const int* nent, const float* ent, const float* table) {
  double result = 0.;
  if (narg < 1 || narg > 5) throw std::runtime_error("fint: narg not within range");
  int index[33];
  double weight[33];
  int lmax = 0, istep = 1, knots = 1;
  index[1] = 1;
  weight[1] = 1.;
  for (int n = 1; n <= narg; ++n) {
    const double x = arg[n - 1];
    const int ndim = nent[n - 1];
    int loca = lmax;
    const int lmin = lmax + 1;
    lmax = lmax + ndim;
    int ishift = 0;
    double eta = 0.;
    bool on_node = false;                       // labels 20 / 21: the argument sits on a grid point
    if (ndim > 2) {
      int locb = lmax + 1, locc = 0;
      bool hit = false;
      do {                                      // label 11
        locc = (loca + locb) / 2;
        const double d = x - (double)ent[locc - 1];
        if (d < 0.) locb = locc;
        else if (d == 0.) { hit = true; break; }
        else loca = locc;
      } while (locb - loca > 1);
      if (hit) {
        ishift = (locc - lmin) * istep;
        on_node = true;
      } else {
        loca = std::min(std::max(loca, lmin), lmax - 1);
        ishift = (loca - lmin) * istep;
        eta = (x - (double)ent[loca - 1]) / (double)(float)(ent[loca] - ent[loca - 1]);    // REAL*4 difference
      }
    } else {
      if (ndim == 1) continue;                  // label 100 directly: istep is not advanced (as written)
      const double h = x - (double)ent[lmin - 1];
      if (h == 0.) { istep = istep * ndim; continue; }
      ishift = istep;
      if (x - (double)ent[lmin] == 0.) on_node = true;
      else {
        ishift = 0;
        eta = h / (double)(float)(ent[lmin] - ent[lmin - 1]);
      }
    }
    if (on_node) {
      for (int k = 1; k <= knots; ++k) index[k] = index[k] + ishift;
    } else {
      for (int k = 1; k <= knots; ++k) {
        index[k] = index[k] + ishift;
        index[k + knots] = index[k] + istep;
        weight[k + knots] = weight[k] * eta;
        weight[k] = weight[k] - weight[k + knots];
      }
      knots = 2 * knots;
    }
    istep = istep * ndim;
  }
  for (int k = 1; k <= knots; ++k) result = result + weight[k] * (double)table[index[k] - 1];
  return result;
}

// physics_kaon.f:241-355 (eekeek: K+ Lambda) and 357-489 (eekeeks: K+ Sigma0); the two differ in their grids only
double eekeek(const SaghaiTable& T, bool lambda, double mrec_struck, double ss, double q22, double angl, double theta,
              double phi, double epsi) {
  using std::complex;
  const std::vector<float>& tab = lambda ? T.proton : T.sigma0;
  const double w = sqrt(ss) * 1000.;
  double skc2 = powi(w * w - K::Mk2 - mrec_struck * mrec_struck, 2) - 4. * K::Mk2 * (mrec_struck * mrec_struck);
  skc2 = std::max(skc2, 0.);
  const double skc = sqrt(skc2) / 2. / w;
  const double q0 = -(-q22 - w * w + K::Mp2) / 2. / K::Mp;
  const double q0c = (-q22 + q0 * K::Mp) / w;
  const double qr = sqrt(q22) / q0c;
  const double aflx = skc / 2. / w / (w * w - K::Mp2) * (K::hbarc * K::hbarc) * 10000.;
  const double aflxl = aflx * (qr * qr);
  const double an = angl * 180. / K::pi;
  const double x = cos(angl), sx = sin(angl);
  float px[3], pa[50];
  int pna[3];
  px[0] = (float)ss;
  px[1] = (float)(q22 / 1.e+06);
  px[2] = (float)an;
  size_t n_tab;
  if (lambda) {
    pna[0] = 10; pna[1] = 11; pna[2] = 19;
    double ps = 2.6, qs = 0.0, as = 0.0;
    for (int i = 1; i <= 10; ++i) { pa[i - 1] = (float)ps; ps = ps + 0.3; }
    for (int i = 11; i <= 21; ++i) { pa[i - 1] = (float)qs; qs = qs + 0.2; }
    for (int i = 22; i <= 40; ++i) { pa[i - 1] = (float)as; as = as + 10.; }
    n_tab = 10 * 11 * 19;
  } else {
    pna[0] = 20; pna[1] = 10; pna[2] = 19;
    static const double grid[30] = {2.851, 2.898, 2.945, 2.991, 3.038, 3.085, 3.132, 3.320, 3.507, 3.695,
                                    3.883, 4.070, 4.258, 4.446, 4.633, 4.821, 5.009, 5.196, 5.384, 5.572,
                                    0.0,   0.250, 0.376, 0.520, 0.750, 1.000, 1.250, 1.500, 1.750, 2.000};
    for (int i = 0; i < 30; ++i) pa[i] = (float)grid[i];
    double as = 0.0;
    for (int i = 31; i <= 49; ++i) { pa[i - 1] = (float)as; as = as + 10.; }
    n_tab = 20 * 10 * 19;
  }
  if (tab.size() != 12 * n_tab) throw std::runtime_error("eekeek: Saghai table not set");
  double zf[2][6];
  for (int k = 0; k < 6; ++k) {
    zf[0][k] = fint(3, px, pna, pa, tab.data() + (size_t)k * n_tab);
    zf[1][k] = fint(3, px, pna, pa, tab.data() + (size_t)(6 + k) * n_tab);
  }
  const complex<double> z1a(zf[0][0], zf[1][0]), z2a(zf[0][1], zf[1][1]), z3a(zf[0][2], zf[1][2]),
      z4a(zf[0][3], zf[1][3]), z7a(zf[0][4], zf[1][4]), z8a(zf[0][5], zf[1][5]);
  auto abs2 = [](const complex<double>& z) { const double a = std::abs(z); return a * a; };   // abs(z)**2
  const double dsigt00 =
      aflx * (abs2(z1a) + abs2(z2a) + 2. * std::real(std::conj(z1a) * z2a) * x +
              0.5 * (sx * sx) * (abs2(z3a) + abs2(z4a) +
                                 2. * std::real(std::conj(z1a) * z4a - std::conj(z2a) * z3a + std::conj(z3a) * z4a * x)));
  const double dsigl00 = aflxl * epsi * (abs2(z7a) + abs2(z8a) + 2. * std::real(std::conj(z7a) * z8a) * x);
  const double dsigp00 = aflx * epsi * powi(sin(theta), 2) * cos(2. * phi) *
                         (0.5 * abs2(z3a) + 0.5 * abs2(z4a) +
                          std::real(std::conj(z1a) * z4a - std::conj(z2a) * z3a + std::conj(z3a) * z4a * x));
  const double dsigi00 = aflx * sqrt(2. * (qr * qr) * epsi * (1. + epsi)) * sin(theta) * cos(phi) *
                         std::real(z7a * (std::conj(z3a) - std::conj(z2a) + std::conj(z4a) * x) +
                                   z8a * (std::conj(z1a) + std::conj(z3a) * x + std::conj(z4a)));
  return dsigt00 + dsigl00 + dsigp00 + dsigi00;
}

double peeK(Sim& s, const Event& vertex, EventMain& main, double& survivalprob) {
  const simc_run_config& cfg = *s.cfg;
  const simc_target& targ = cfg.targ;
  const Fermi F = fermi_of(s);
  CmFrame C;
  transform_to_cm(vertex, main, F, C);
  double jacobian = C.jacobian / (2. * C.phadcm * C.qstar);
  double jac_old = C.jac_old / (2. * C.phadcm * C.qstar);
  main.thetacm = C.thetacm;
  main.phicm = C.phicm;
  main.pcm = C.phadcm;
  main.davejac = jacobian;
  main.johnjac = jac_old;
  double tfcos = F.pferx * vertex.uq.x + F.pfery * vertex.uq.y + F.pferz * vertex.uq.z;
  if (tfcos - 1. > 0. && tfcos - 1. < 1.e-8) tfcos = 1.0;
  const double tfsin = sqrt(1. - tfcos * tfcos);
  const double sgev = powi(vertex.nu + F.efer, 2) - powi(vertex.q + F.pfer * tfcos, 2) - powi(F.pfer * tfsin, 2);
  main.wcm = sqrt(sgev);
  // physics_kaon.f:100-108.  `phi` is a local of peeK that nothing assigns; with -fno-automatic it is static
  // storage, i.e. zero ("WE ARE ALWAYS CALCULATING FOR PHI=0", physics_kaon.f:97-98).
  {
    const SaghaiTable* T = saghai_tables();
    const bool lambda = targ.Mrec_struck < 1150.;
    const double phi = 0.0;
    if (!(lambda ? T->proton : T->sigma0).empty())
      s.ntup.sigcm1 = eekeek(*T, lambda, targ.Mrec_struck, sgev / 1.e6, vertex.Q2, main.thetacm, main.theta_pq, phi, main.epsilon);
  }
  s.ntup.sigcm2 = sig_factorized(vertex.Q2, main.wcm, main.t, C.phadcm, targ.Mrec_struck);
  const double sigma_eek = s.ntup.sigcm2;
  s.ntup.sigcm = sigma_eek;
  const double k_eq = (main.wcm * main.wcm - targ.Mtar_struck * targ.Mtar_struck) / 2. / targ.Mtar_struck;
  const double fac = 1. / (1. - F.pferz * F.pfer / F.efer) * targ.Mtar_struck / F.efer;
  const double gtpr = K::alpha / 2. / (K::pi * K::pi) * vertex.e.E / vertex.Ein * k_eq / vertex.Q2 / (1. - main.epsilon);
  const double result = sigma_eek * jacobian * (gtpr * fac);
  survivalprob = 1.0;
  if (!cfg.doing_decay) {
    double zaero = 0.;
    if (cfg.hadron_arm == 1) zaero = 0.;
    else if (cfg.hadron_arm == 2) zaero = -82.8;
    else if (cfg.hadron_arm == 3) zaero = -183.;
    else if (cfg.hadron_arm == 4) zaero = -183.;
    const double pathlen = main.FP_p.path + zaero * (1 + main.FP_p.dx * main.FP_p.dx + main.FP_p.dy * main.FP_p.dy);
    const double betak = cfg.spec_p.P / sqrt(cfg.spec_p.P * cfg.spec_p.P + cfg.Mh2);
    const double gammak = 1. / sqrt(1. - betak * betak);
    survivalprob = 1. / exp(pathlen / (cfg.ctau * betak * gammak));
    s.trk.decdist = survivalprob;
  }
  return result;
}

// physics_pion.f:404-465
double sig_blok(double thetacm, double phicm, double t, double q2_gev, double s_gev, double eps, double mtar_gev, int which_pion) {
  double sigl = 27.8 * exp(-11.5 * fabs(t));
  double sigt = 10.0 * (5. * fabs(t)) * exp(-5. * fabs(t));
  const double siglt = 0.0 * sin(thetacm);
  double sigtt = -(4.0 * sigl + 0.5 * sigt) * powi(sin(thetacm), 2);
  if (which_pion == 1 || which_pion == 11 || which_pion == 3) {
    sigt = sigt * 0.25 * (1. + 3. * exp(-10. * fabs(t)));
    sigtt = sigtt * 0.25 * (1. + 3. * exp(-10. * fabs(t)));
  }
  const double fpi = 1. / (1. + 1.65 * q2_gev + 0.5 * powi(q2_gev, 2));
  const double fpi2 = powi(fpi, 2);
  sigl = sigl * (fpi2 * q2_gev) / 0.1215;
  sigt = sigt / (0.3 + q2_gev);
  sigtt = sigtt / (0.3 + q2_gev);
  const double sig219 = (sigt + eps * sigl + eps * cos(2. * phicm) * sigtt + sqrt(2.0 * eps * (1. + eps)) * cos(phicm) * siglt) / 1.e0;
  double sig = sig219 * 15.333 / powi(s_gev - powi(mtar_gev, 2), 2);
  sig = sig / 2. / K::pi / 1.e+06;
  return sig;
}

// physics_delta.f:1-135
double peedelta(Sim& s, const Event& vertex, EventMain& main) {
  const simc_run_config& cfg = *s.cfg;
  const simc_target& targ = cfg.targ;
  const Fermi F = fermi_of(s);
  CmFrame C;
  transform_to_cm(vertex, main, F, C);
  main.thetacm = C.thetacm;
  main.phicm = C.phicm;
  main.pcm = C.phadcm;
  main.davejac = C.jacobian;
  main.johnjac = C.jac_old;
  double tfcos = F.pferx * vertex.uq.x + F.pfery * vertex.uq.y + F.pferz * vertex.uq.z;
  if (tfcos - 1. > 0. && tfcos - 1. < 1.e-8) tfcos = 1.0;
  const double tfsin = sqrt(1. - tfcos * tfcos);
  const double sgev = powi(vertex.nu + F.efer, 2) - powi(vertex.q + F.pfer * tfcos, 2) - powi(F.pfer * tfsin, 2);
  main.wcm = sqrt(sgev);
  const double k_eq = (main.wcm * main.wcm - targ.Mtar_struck * targ.Mtar_struck) / 2. / targ.Mtar_struck;
  s.ntup.sigcm1 = sig_blok(C.thetacm, C.phicm, main.t / 1.e6, vertex.Q2 / 1.e6, sgev / 1.e6, main.epsilon,
                           targ.Mtar_struck / 1000., cfg.which_pion);
  s.ntup.sigcm = s.ntup.sigcm1;
  const double fac = 1. / (1. - F.pferz * F.pfer / F.efer) * targ.Mtar_struck / F.efer;
  const double gtpr = K::alpha / 2. / (K::pi * K::pi) * vertex.e.E / vertex.Ein * k_eq / vertex.Q2 / (1. - main.epsilon);
  return 1.0 * C.jacobian * (gtpr * fac);
}

// rho_physics.f:1-402: p(e,e'rho)p after PYTHIA / the HERMES Monte Carlo.  sigma_T from a fit to photoproduction,
// R = sigma_L/sigma_T and the Q2 dependence from HERMES, exponential t' slope b(c*delta_tau), virtual-photon flux.
// Overwrites main%W (GeV) and main%t (GeV^2), as the reference does.
double peerho(Sim& s, const Event& vertex, EventMain& main) {
  const simc_run_config& cfg = *s.cfg;
  const simc_target& targ = cfg.targ;
  const double pi = K::pi;
  const double Q2_g = vertex.Q2 / 1000000.;
  const double phipq = main.phi_pq;
  const double cospq = cos(phipq), sinpq = sin(phipq);
  const double pfer = s.pfer, pferx = s.pferx, pfery = s.pfery, pferz = s.pferz;
  s.efer = sqrt(pfer * pfer + targ.Mtar_struck * targ.Mtar_struck);
  if (cfg.doing_deutpi || cfg.doing_hepi) {
    s.efer = targ.M - sqrt(K::Mn * K::Mn + pfer * pfer);
    if (cfg.doing_hepi) s.efer = s.efer - K::Mp;
  }
  const double efer = s.efer;
  double tcos = vertex.up.x * vertex.uq.x + vertex.up.y * vertex.uq.y + vertex.up.z * vertex.uq.z;
  if (tcos - 1. > 0. && tcos - 1. < 1.e-8) tcos = 1.0;
  const double tsin = sqrt(1. - tcos * tcos);
  double tfcos = pferx * vertex.uq.x + pfery * vertex.uq.y + pferz * vertex.uq.z;
  if (tfcos - 1. > 0. && tfcos - 1. < 1.e-8) tfcos = 1.0;
  const double tfsin = sqrt(1. - tfcos * tfcos);
  const double epsi = 1. / (1. + 2 * (1. + vertex.nu * vertex.nu / vertex.Q2) * powi(tan(vertex.e.theta / 2.), 2));
  double ss = powi(vertex.nu + efer, 2) - powi(vertex.q + pfer * tfcos, 2) - powi(pfer * tfsin, 2);
  ss = ss / 1.e6;
  main.W = sqrt(ss);
  double t = vertex.Q2 - K::Mrho2 + 2. * vertex.nu * vertex.p.E - 2. * vertex.p.P * vertex.q * tcos;
  t = t / 1.e6;
  main.t = t;

  const double qx = -vertex.uq.y, qy = vertex.uq.x, qz = vertex.uq.z;
  const double px = -pfery, py = pferx, pz = pferz;
  double dummy = sqrt((qx * qx + qy * qy) * (qx * qx + qy * qy + qz * qz));
  double new_x_x = -qx * qz / dummy, new_x_y = -qy * qz / dummy, new_x_z = (qx * qx + qy * qy) / dummy;
  dummy = sqrt(qx * qx + qy * qy);
  double new_y_x = qy / dummy, new_y_y = -qx / dummy, new_y_z = 0.0;
  const double p_new_x = pfer * (px * new_x_x + py * new_x_y + pz * new_x_z);
  const double p_new_y = pfer * (px * new_y_x + py * new_y_y + pz * new_y_z);
  double phiqn;
  if (p_new_x == 0.) phiqn = 0.;
  else phiqn = atan2(p_new_y, p_new_x);
  if (phiqn < 0.) phiqn = phiqn + 2. * pi;

  const double pbeam = sqrt(vertex.Ein * vertex.Ein - K::Me * K::Me);
  const double beam_newx = pbeam * new_x_z, beam_newy = pbeam * new_y_z, beam_newz = pbeam * vertex.uq.z;
  const double bstar = sqrt(powi(vertex.q + pfer * tfcos, 2) + powi(pfer * tfsin, 2)) / (efer + vertex.nu);
  const double gstar = 1. / sqrt(1. - bstar * bstar);
  const double bstarz = (vertex.q + pfer * tfcos) / (efer + vertex.nu);
  const double bstarx = p_new_x / (efer + vertex.nu);
  const double bstary = p_new_y / (efer + vertex.nu);
  const double zero = 0.e0;
  double nustar, qstarx, qstary, qstarz, qstar;
  loren(gstar, bstarx, bstary, bstarz, vertex.nu, zero, zero, vertex.q, nustar, qstarx, qstary, qstarz, qstar);
  const double ppiz = vertex.p.P * tcos, ppix = vertex.p.P * tsin * cospq, ppiy = vertex.p.P * tsin * sinpq;
  double epicm, ppicmx, ppicmy, ppicmz, ppicm;
  loren(gstar, bstarx, bstary, bstarz, vertex.p.E, ppix, ppiy, ppiz, epicm, ppicmx, ppicmy, ppicmz, ppicm);
  const double thetacm = acos((ppicmx * qstarx + ppicmy * qstary + ppicmz * qstarz) / ppicm / qstar);
  main.pcm = ppicm;
  double ebeamcm, pbeamcmx, pbeamcmy, pbeamcmz, pbeamcm;
  loren(gstar, bstarx, bstary, bstarz, vertex.Ein, beam_newx, beam_newy, beam_newz, ebeamcm, pbeamcmx, pbeamcmy, pbeamcmz,
        pbeamcm);
  dummy = sqrt(powi(qstary * pbeamcmz - qstarz * pbeamcmy, 2) + powi(qstarz * pbeamcmx - qstarx * pbeamcmz, 2) +
               powi(qstarx * pbeamcmy - qstary * pbeamcmx, 2));
  new_y_x = (qstary * pbeamcmz - qstarz * pbeamcmy) / dummy;
  new_y_y = (qstarz * pbeamcmx - qstarx * pbeamcmz) / dummy;
  new_y_z = (qstarx * pbeamcmy - qstary * pbeamcmx) / dummy;
  dummy = sqrt(powi(new_y_y * qstarz - new_y_z * qstary, 2) + powi(new_y_z * qstarx - new_y_x * qstarz, 2) +
               powi(new_y_x * qstary - new_y_y * qstarx, 2));
  new_x_x = (new_y_y * qstarz - new_y_z * qstary) / dummy;
  new_x_y = (new_y_z * qstarx - new_y_x * qstarz) / dummy;
  new_x_z = (new_y_x * qstary - new_y_y * qstarx) / dummy;
  const double new_z_x = qstarx / qstar, new_z_y = qstary / qstar, new_z_z = qstarz / qstar;
  const double ppicm_newx = ppicmx * new_x_x + ppicmy * new_x_y + ppicmz * new_x_z;
  const double ppicm_newy = ppicmx * new_y_x + ppicmy * new_y_y + ppicmz * new_y_z;
  double phicm = atan2(ppicm_newy, ppicm_newx);
  if (phicm < 0.) phicm = 2. * 3.141592654 + phicm;
  main.thetacm = thetacm;
  main.phicm = phicm;

  const double mt = targ.Mtar_struck / 1000.;
  const double tmin = -(powi((-Q2_g - K::Mrho2 / 1.e6 - mt * mt + mt * mt) / (2. * sqrt(ss)), 2) - powi((qstar - ppicm) / 1000., 2));
  const double tprime = fabs(t - tmin);
  const double sig0 = 41.263 / std::pow(vertex.nu / 1000.0, 0.4765);
  double R = 0.33 * std::pow(vertex.Q2 / K::Mrho2, 0.61);
  if (R < 0.) R = 0.;
  double sigt = sig0 * (1.0 + epsi * R) * std::pow(K::Mrho2 / (vertex.Q2 + K::Mrho2), 2.575);
  if (sigt < 0.) sigt = 0.;
  const double cdeltatau = K::hbarc / (sqrt(vertex.nu * vertex.nu + vertex.Q2 + K::Mrho2) - vertex.nu);
  double brho;
  if (cdeltatau < 2.0) {
    brho = 4.4679 + 8.6106 * log10(cdeltatau);
    if (brho < 1.0) brho = 1.0;
  } else {
    brho = 7.0;
  }
  const double sig219 = sigt * brho * exp(-brho * tprime) / 2.0 / pi;
  double sig = sig219 / 1.e+06;
  sig = sig * 2. * qstar * ppicm;
  double gtpr = K::alpha / 2. / (pi * pi) * vertex.e.E / vertex.Ein * (ss - mt * mt) / 2. / ((efer - pfer * tfcos) / 1000.) / Q2_g /
                (1. - epsi);
  if (gtpr <= 0.) gtpr = 0.;

  // the "full blown Jacobian" of rho_physics.f:281-351: main%davejac only (the weight does not use it)
  const double psign = cos(phiqn) * cospq + sin(phiqn) * sinpq;
  const double P = vertex.p.P, E = vertex.p.E;
  const double square_root = vertex.q + pfer * tfcos - P * tcos;
  const double dp_dcos_num = P + (P * P * tcos - psign * pfer * P * tfsin * tcos / tsin) / square_root;
  const double dp_dcos_den = ((vertex.nu + efer - E) * P / E + P * tsin * tsin - psign * pfer * tfsin * tsin) / square_root - tcos;
  const double dp_dcos = dp_dcos_num / dp_dcos_den;
  const double dp_dphi_num = pfer * P * tsin * tfsin * (cos(phiqn) * sinpq - sin(phiqn) * cospq) / square_root;
  const double dp_dphi_den = tcos + (pfer * tsin * tfsin * psign - P * tsin * tsin - (vertex.nu + efer - E) * P / E) / square_root;
  const double dp_dphi = dp_dphi_num / dp_dphi_den;
  const double dt_dcos_lab = 2. * (vertex.q * P + (vertex.q * tcos - vertex.nu * P / E) * dp_dcos);
  const double dt_dphi_lab = 2. * (vertex.q * tcos - vertex.nu * P / E) * dp_dphi;
  const double g1b2 = (gstar - 1.) / (bstar * bstar);
  const double dpxdphi = P * tsin * (-sinpq + (gstar - 1.) * bstarx / (bstar * bstar) * (bstary * cospq - bstarx * sinpq)) +
                         ((ppicmx + gstar * bstarx * E) / P - gstar * bstarx * P / E) * dp_dphi;
  const double dpydphi = P * tsin * (cospq + (gstar - 1.) * bstary / (bstar * bstar) * (bstary * cospq - bstarx * sinpq)) +
                         ((ppicmy + gstar * bstary * E) / P - gstar * bstary * P / E) * dp_dphi;
  const double dpzdphi = P * (gstar - 1.) / (bstar * bstar) * bstarz * tsin * (bstary * cospq - bstarx * sinpq) +
                         ((ppicmz + gstar * bstarz * E) / P - gstar * bstarz * P / E) * dp_dphi;
  const double dpxdcos = -P * tcos / tsin * (cospq + (gstar - 1.) * bstarx / (bstar * bstar) *
                                                         (bstarx * cospq + bstary * sinpq - bstarz * tsin / tcos)) +
                         ((ppicmx + gstar * bstarx * E) / P - gstar * bstarx * P / E) * dp_dcos;
  const double dpydcos = -P * tcos / tsin * (sinpq + (gstar - 1.) * bstary / (bstar * bstar) *
                                                         (bstarx * cospq + bstary * sinpq - bstarz * tsin / tcos)) +
                         ((ppicmy + gstar * bstary * E) / P - gstar * bstary * P / E) * dp_dcos;
  const double dpzdcos = P * (1. - g1b2 * bstarz * tcos / tsin * (bstarx * cospq + bstary * sinpq - tsin / tcos * bstarz)) +
                         ((ppicmz + gstar * bstarz * E) / P - gstar * bstarz * P / E) * dp_dcos;
  const double dpxnewdphi = dpxdphi * new_x_x + dpydphi * new_x_y + dpzdphi * new_x_z;
  const double dpynewdphi = dpxdphi * new_y_x + dpydphi * new_y_y + dpzdphi * new_y_z;
  const double dphicmdphi = (dpynewdphi * ppicm_newx - ppicm_newy * dpxnewdphi) / (ppicm_newx * ppicm_newx + ppicm_newy * ppicm_newy);
  const double dpxnewdcos = dpxdcos * new_x_x + dpydcos * new_x_y + dpzdcos * new_x_z;
  const double dpynewdcos = dpxdcos * new_y_x + dpydcos * new_y_y + dpzdcos * new_y_z;
  const double dphicmdcos = (dpynewdcos * ppicm_newx - ppicm_newy * dpxnewdcos) / (ppicm_newx * ppicm_newx + ppicm_newy * ppicm_newy);
  main.davejac = fabs(dt_dcos_lab * dphicmdphi - dt_dphi_lab * dphicmdcos);
  main.johnjac = 2 * (efer - 2 * pferz * pfer * E / P * tcos) * (vertex.q + pferz * pfer) * P /
                     (efer + vertex.nu - (vertex.q + pferz * pfer) * E / P * tcos) -
                 2 * P * pfer;
  (void)new_z_x; (void)new_z_y; (void)new_z_z; (void)ebeamcm; (void)epicm; (void)nustar;
  const double davesig = gtpr * sig;
  double sigma_eerho = davesig / 1.e3;
  if (sigma_eerho > 1.E10 || sigma_eerho < 0.) sigma_eerho = 0.;
  s.ntup.sigcm = sig;
  return sigma_eerho;
}

}  // namespace simc_oracle
