// Meson electroproduction weights on the device (and on the host for the central weight):
// transform_to_cm (jacobians.f:1-282), peepi with sig_param_2021/exclfit (physics_pion.f:1-130,
// 738-854), peeK with sig_factorized (physics_kaon.f:1-236).  The struck nucleon's momentum and energy
// (pfer, efer) come with the vertex: at rest for hydrogen (pfer = 0, efer = Mtar_struck, event.f:330-335),
// thrown from the deuteron's momentum distribution for D(e,e'pi/K) (event.f:337-367).
//
// The MAID-2007 branch of peepi below W = 2 GeV (sigmaid, physics_pion.f:131-154, 577-728) reads the caller's
// table (simc_b200_set_maid_table); without it such events take the parametrisation alone and are counted in
// simc_accum.unsupported.  The Saghai model of peeK (eekeek / eekeeks, physics_kaon.f:241-489, with CERNLIB's fint)
// feeds the ntuple column sigcm1 only; it is evaluated where rows are produced, when the caller has set its tables
// (simc_b200_set_saghai_table).
#pragma once
#include "target.cuh"

namespace simc {

struct MesonVertex {          // what the weights read from `vertex` / `main`
  double Ein, eE, nu, q, Q2, pP, pE;
  double uqx, uqy, uqz, upx, upy, upz;
  double phi_pq, t, epsilon;
  // COMMON /pfermi_stuff/ (simulate.inc:212-217) as generate left it: zero and Mtar_struck for hydrogen
  double pfer, pferx, pfery, pferz, efer;
};
struct MesonCm {
  double thetacm, phicm, pcm, qstar, jacobian, jac_old, wcm, sgev;
};

namespace mesondetail {
struct MV4 { double e, x, y, z, p; };
// loren.f:1-26
SIMC_HD MV4 loren(double gam, double bx, double by, double bz, double e, double x, double y, double z) {
  MV4 r;
  const double gam1 = gam * gam / (1. + gam);
  r.e = gam * (e - bx * x - by * y - bz * z);
  r.x = (1 + gam1 * bx * bx) * x + gam1 * bx * (by * y + bz * z) - gam * bx * e;
  r.y = (1 + gam1 * by * by) * y + gam1 * by * (bx * x + bz * z) - gam * by * e;
  r.z = (1 + gam1 * bz * bz) * z + gam1 * bz * (by * y + bx * x) - gam * bz * e;
  r.p = sqrt(r.x * r.x + r.y * r.y + r.z * r.z);
  return r;
}
SIMC_HD double msq(double x) { return x * x; }
}  // namespace mesondetail

// jacobians.f:1-282
SIMC_HD_CALL void transform_to_cm(const MesonVertex& v, MesonCm& C) {
  using namespace mesondetail;
  const double pi = 3.141592653589793;
  const double pfer = v.pfer, pferx = v.pferx, pfery = v.pfery, pferz = v.pferz, efer = v.efer;
  double tcos = v.upx * v.uqx + v.upy * v.uqy + v.upz * v.uqz;
  if (tcos - 1. > 0. && tcos - 1. < 1.e-8) tcos = 1.0;
  const double tsin = sqrt(1. - tcos * tcos);
  double tfcos = pferx * v.uqx + pfery * v.uqy + pferz * v.uqz;
  if (tfcos - 1. > 0. && tfcos - 1. < 1.e-8) tfcos = 1.0;
  const double tfsin = sqrt(1. - tfcos * tfcos);
  const double cospq = m::cos(v.phi_pq), sinpq = m::sin(v.phi_pq);
  const double qx = -v.uqy, qy = v.uqx, qz = v.uqz;
  const double px = -pfery, py = pferx, pz = pferz;
  double dummy = sqrt((qx * qx + qy * qy) * (qx * qx + qy * qy + qz * qz));
  const double tmp_x_x = -qx * qz / dummy, tmp_x_y = -qy * qz / dummy, tmp_x_z = (qx * qx + qy * qy) / dummy;
  dummy = sqrt(qx * qx + qy * qy);
  const double tmp_y_x = qy / dummy, tmp_y_y = -qx / dummy, tmp_y_z = 0.0;
  const double p_tmp_x = pfer * (px * tmp_x_x + py * tmp_x_y + pz * tmp_x_z);
  const double p_tmp_y = pfer * (px * tmp_y_x + py * tmp_y_y + pz * tmp_y_z);
  double phiqn;
  if (p_tmp_x == 0.) phiqn = 0.;
  else phiqn = m::atan2(p_tmp_y, p_tmp_x);
  if (phiqn < 0.) phiqn = phiqn + 2. * pi;
  const double cosqn = m::cos(phiqn), sinqn = m::sin(phiqn);

  const double pbeam = v.Ein;
  const double beam_tmpx = pbeam * tmp_x_z, beam_tmpy = pbeam * tmp_y_z, beam_tmpz = pbeam * v.uqz;
  const double bstar = sqrt(msq(v.q + pfer * tfcos) + msq(pfer * tfsin)) / (efer + v.nu);
  const double gstar = 1. / sqrt(1. - bstar * bstar);
  const double bstarz = (v.q + pfer * tfcos) / (efer + v.nu);
  const double bstarx = p_tmp_x / (efer + v.nu);
  const double bstary = p_tmp_y / (efer + v.nu);
  const MV4 beam = loren(gstar, bstarx, bstary, bstarz, v.Ein, beam_tmpx, beam_tmpy, beam_tmpz);
  const MV4 qs = loren(gstar, bstarx, bstary, bstarz, v.nu, 0.e0, 0.e0, v.q);
  const double phadz = v.pP * tcos, phadx = v.pP * tsin * cospq, phady = v.pP * tsin * sinpq;
  const MV4 had = loren(gstar, bstarx, bstary, bstarz, v.pE, phadx, phady, phadz);
  C.thetacm = m::acos((had.x * qs.x + had.y * qs.y + had.z * qs.z) / had.p / qs.p);
  C.pcm = had.p;
  C.qstar = qs.p;

  dummy = sqrt(msq(qs.y * beam.z - qs.z * beam.y) + msq(qs.z * beam.x - qs.x * beam.z) + msq(qs.x * beam.y - qs.y * beam.x));
  const double tmp2_y_x = (qs.y * beam.z - qs.z * beam.y) / dummy;
  const double tmp2_y_y = (qs.z * beam.x - qs.x * beam.z) / dummy;
  const double tmp2_y_z = (qs.x * beam.y - qs.y * beam.x) / dummy;
  dummy = sqrt(msq(tmp2_y_y * qs.z - tmp2_y_z * qs.y) + msq(tmp2_y_z * qs.x - tmp2_y_x * qs.z) +
               msq(tmp2_y_x * qs.y - tmp2_y_y * qs.x));
  const double tmp2_x_x = (tmp2_y_y * qs.z - tmp2_y_z * qs.y) / dummy;
  const double tmp2_x_y = (tmp2_y_z * qs.x - tmp2_y_x * qs.z) / dummy;
  const double tmp2_x_z = (tmp2_y_x * qs.y - tmp2_y_y * qs.x) / dummy;
  const double had2x = had.x * tmp2_x_x + had.y * tmp2_x_y + had.z * tmp2_x_z;
  const double had2y = had.x * tmp2_y_x + had.y * tmp2_y_y + had.z * tmp2_y_z;
  C.phicm = m::atan2(had2y, had2x);
  if (C.phicm < 0.) C.phicm = 2. * pi + C.phicm;

  // dt dphi_cm -> dOmega_lab, jacobians.f:180-262
  const double P = v.pP, E = v.pE;
  const double psign = cosqn * cospq + sinqn * sinpq;
  const double square_root = v.q + pfer * tfcos - P * tcos;
  const double dp_dcos_num = P + (P * P * tcos - psign * pfer * P * tfsin * tcos / tsin) / square_root;
  const double dp_dcos_den = ((v.nu + efer - E) * P / E + P * tsin * tsin - psign * pfer * tfsin * tsin) / square_root - tcos;
  const double dp_dcos = dp_dcos_num / dp_dcos_den;
  const double dp_dphi_num = pfer * P * tsin * tfsin * (cosqn * sinpq - sinqn * cospq) / square_root;
  const double dp_dphi_den = tcos + (pfer * tsin * tfsin * psign - P * tsin * tsin - (v.nu + efer - E) * P / E) / square_root;
  const double dp_dphi = dp_dphi_num / dp_dphi_den;
  const double dt_dcos_lab = 2. * (v.q * P + (v.q * tcos - v.nu * P / E) * dp_dcos);
  const double dt_dphi_lab = 2. * (v.q * tcos - v.nu * P / E) * dp_dphi;
  const double b2 = bstar * bstar;
  const double kx = (had.x + gstar * bstarx * E) / P - gstar * bstarx * P / E;
  const double ky = (had.y + gstar * bstary * E) / P - gstar * bstary * P / E;
  const double kz = (had.z + gstar * bstarz * E) / P - gstar * bstarz * P / E;
  const double dpxdphi = P * tsin * (-sinpq + (gstar - 1.) * bstarx / b2 * (bstary * cospq - bstarx * sinpq)) + kx * dp_dphi;
  const double dpydphi = P * tsin * (cospq + (gstar - 1.) * bstary / b2 * (bstary * cospq - bstarx * sinpq)) + ky * dp_dphi;
  const double dpzdphi = P * (gstar - 1.) / b2 * bstarz * tsin * (bstary * cospq - bstarx * sinpq) + kz * dp_dphi;
  const double dpxdcos =
      -P * tcos / tsin * (cospq + (gstar - 1.) * bstarx / b2 * (bstarx * cospq + bstary * sinpq - bstarz * tsin / tcos)) +
      kx * dp_dcos;
  const double dpydcos =
      -P * tcos / tsin * (sinpq + (gstar - 1.) * bstary / b2 * (bstarx * cospq + bstary * sinpq - bstarz * tsin / tcos)) +
      ky * dp_dcos;
  const double dpzdcos =
      P * (1. - (gstar - 1.) / b2 * bstarz * tcos / tsin * (bstarx * cospq + bstary * sinpq - tsin / tcos * bstarz)) +
      kz * dp_dcos;
  const double dpxnewdphi = dpxdphi * tmp2_x_x + dpydphi * tmp2_x_y + dpzdphi * tmp2_x_z;
  const double dpynewdphi = dpxdphi * tmp2_y_x + dpydphi * tmp2_y_y + dpzdphi * tmp2_y_z;
  const double den = had2x * had2x + had2y * had2y;
  const double dphicmdphi = (dpynewdphi * had2x - had2y * dpxnewdphi) / den;
  const double dpxnewdcos = dpxdcos * tmp2_x_x + dpydcos * tmp2_x_y + dpzdcos * tmp2_x_z;
  const double dpynewdcos = dpxdcos * tmp2_y_x + dpydcos * tmp2_y_y + dpzdcos * tmp2_y_z;
  const double dphicmdcos = (dpynewdcos * had2x - had2y * dpxnewdcos) / den;
  C.jacobian = fabs(dt_dcos_lab * dphicmdphi - dt_dphi_lab * dphicmdcos);
  C.jac_old = 2 * (efer - 2 * pferz * pfer * E / P * tcos) * (v.q + pferz * pfer) * P /
                  (efer + v.nu - (v.q + pferz * pfer) * E / P * tcos) -
              2 * P * pfer;
  // s of the photon-nucleon system, physics_pion.f:60-66 / physics_kaon.f:70-75
  C.sgev = msq(v.nu + efer) - msq(v.q + pfer * tfcos) - msq(pfer * tfsin);
  C.wcm = sqrt(C.sgev);
}

// physics_pion.f:807-854; p is the 1-based parameter array of the fit
SIMC_HD_CALL double exclfit(double t, double thetacm, double phicm, double q2_gev, double s_gev, double eps,
                            const double* pp, double fpifact) {
  using mesondetail::msq;
  const double* p = pp - 1;
  const double mtar_gev = 0.938;
  const double fpi = fpifact / (1.0 + p[1] * q2_gev + p[2] * (q2_gev * q2_gev));
  const double q2fpi2 = q2_gev * (fpi * fpi);
  const double at = fabs(t);
  double sigL = (p[3] + p[15] / q2_gev) * at / msq(at + 0.02) * q2fpi2 * m::exp(p[4] * at);
  sigL = sigL / (m::pow(s_gev, p[11]) + m::pow(sqrt(s_gev), p[17]));
  double sigT = p[5] / q2_gev * m::exp(p[6] * (q2_gev * q2_gev));
  sigT = sigT / (m::pow(s_gev, p[12]) + m::pow(sqrt(s_gev), p[16]));
  sigT = sigT * m::exp(p[14] * at);
  const double sth = m::sin(thetacm);
  double sigLT = (p[7] / (1.0 + p[10] * q2_gev)) * m::exp(p[8] * at) * sth;
  sigLT = sigLT / m::pow(s_gev, p[13]);
  const double sigTT = (p[9] / (1. + 1.0 * q2_gev)) * m::exp(-7.0 * at) * (sth * sth);
  const double sig219 = (sigT + eps * sigL + eps * m::cos(2.0 * phicm) * sigTT +
                         sqrt(2.0 * eps * (1.0 + eps)) * m::cos(phicm) * sigLT) / 1.0;
  double sig = sig219 * 8.539 / msq(s_gev - mtar_gev * mtar_gev);
  sig = sig / 2.0 / 3.1415928 / 1.0e+06;
  return sig;
}

// physics_pion.f:738-805, charged pions
SIMC_HD_CALL double sig_param_2021(double thcm, double phicm, double t, double q2, double wsq, double eps, int which_pion) {
  const double pp[17] = {1.60077, -0.01523, 37.08142, -4.11060, 23.26192, 0.00983, 0.87073, -5.77115, -271.08678,
                         0.13766, -0.00855, 0.27885,  -1.13212, -1.50415, -6.34766, 0.55769, -0.01709};
  const double pm[17] = {1.75169, 0.11144, 47.35877, -4.69434, 1.60552, 0.00800, 0.44194, -2.29188, -41.67194,
                         0.69475, 0.02527, -0.50178, -1.22825, -1.16878, 5.75825, -1.00355, 0.05055};
  if (which_pion == 1 || which_pion == 11 || which_pion == 3) return exclfit(t, thcm, phicm, q2, wsq, eps, pm, 1.0);
  return exclfit(t, thcm, phicm, q2, wsq, eps, pp, 1.0);
}

// physics_kaon.f:175-236
SIMC_HD_CALL double sig_factorized(double q2, double w, double t, double pk, double mrec) {
  using mesondetail::msq;
  const double Mp = 938.27231, Mk2 = 493.677 * 493.677;
  const double nu = (w * w + q2 - Mp * Mp) / 2. / Mp;
  const double q = sqrt(q2 + nu * nu);
  const double qcm = q * (Mp / w);
  const double nucm = sqrt(qcm * qcm - q2);
  const double tmin = -1. * (Mk2 - q2 - 2 * nucm * sqrt(pk * pk + Mk2) + 2 * qcm * pk);
  const double q2val = q2 / 1.e6, w2val = w * w / 1.e6, pkval = pk / 1000., tval = t / 1.e6, tminval = tmin / 1.e6;
  double fact_q, fact_t, fact_w = 0.0;
  if (mrec < 1150.) {
    fact_q = 1. / msq(q2val + 2.67);
    fact_t = m::exp(-2.1 * (tval - tminval));
    if (w2val != 0) {
      fact_w = 0.959 * 4.1959 * pkval / (sqrt(w2val) * (w2val - 0.93827 * 0.93827));
      fact_w = fact_w + (0.18 * (1.72 * 1.72) * (0.10 * 0.10)) / (msq(w2val - 1.72 * 1.72) + (1.72 * 1.72) * (0.10 * 0.10));
    }
  } else {
    fact_q = 1. / msq(q2val + 0.79);
    fact_t = m::exp(-1.0 * (tval - tminval));
    if (w2val != 0) fact_w = 0.959 * 4.1959 * pkval / (sqrt(w2val) * (w2val - 0.93827 * 0.93827));
  }
  return fact_q * fact_t * fact_w;
}

// maidtbl of sigmaid (physics_pion.f:596), the slice sig0 reads: [25 Q2][46 W][6 cos(theta*)][4 columns] for the
// charge state of this run (ipi = 3: pi+ n, 4: pi- p); null = table not provided
struct MaidDev { const double* tbl; };

// sigmaid, physics_pion.f:577-728: nearest-bin lookup in the MAID-2007 table; peepi only uses sig0
SIMC_HD_CALL double sigmaid_sig0(const MaidDev M, double q2, double w, double e0, double costh, double phi) {
  const double am = 0.9383;
  if (w < 1.08) return 0.;
  const double nu = (w * w - am * am + q2) / 2. / am;
  if (nu > e0) return 0.;
  const double ep = e0 - nu;
  const double sin2 = q2 / 4. / e0 / ep;
  if (sin2 <= 0.0 || sin2 > 1.) return 0.;
  const double eps = 1. / (1. + 2. * (1. + nu * nu / q2) * sin2 / (1. - sin2));
  int iq = (int)((q2 + 0.1) / 0.2);
  iq = min(25, max(1, iq));
  int iw = (int)((w - 1.090) / 0.020);
  iw = min(46, max(1, iw));
  // cthmin / cthmax of physics_pion.f:601-602: the last bin whose closed interval holds costh, else the first
  int ith = 1;
  if (costh >= -0.20 && costh <= 0.20) ith = 1;
  if (costh >= 0.20 && costh <= 0.44) ith = 2;
  if (costh >= 0.44 && costh <= 0.63) ith = 3;
  if (costh >= 0.63 && costh <= 0.78) ith = 4;
  if (costh >= 0.78 && costh <= 0.90) ith = 5;
  if (costh >= 0.90 && costh <= 1.0) ith = 6;
  double wfact = 1.;
  if (w > 1.232) wfact = (w - 1.132) / 0.100;
  const double* row = M.tbl + (((long long)(iq - 1) * 46 + (iw - 1)) * 6 + (ith - 1)) * 4;
  const double ST = row[0] / fmax(0.2, q2) / wfact;
  const double SL = row[1] * ST;
  const double STL = row[2] * ST;
  const double STT = row[3] * ST;
  return ST + eps * SL + sqrt(2. * eps * (1. + eps)) * m::cos(phi) * STL + eps * m::cos(2. * phi) * STT;
}

struct MesonWeight {
  double sigcc, sigcm, thetacm, phicm, pcm, wcm, davejac, johnjac;
  double sigcm1;              // peeK only: the Saghai model (ntup%sigcm1), 0 unless asked for
  double t_gev;               // peerho only: the t it leaves in main%t (GeV^2)
  bool low_w;                 // W < 2 GeV: the reference would blend in the MAID table here
};

// physics_pion.f:1-130
SIMC_HD_CALL MesonWeight peepi(const simc_run_config& cfg, const MaidDev maid, const MesonVertex& v) {
  const double pi = 3.141592653589793, alpha = 1. / 137.0359895;
  const double Mtar = cfg.targ.Mtar_struck, efer = v.efer, pfer = v.pfer, pferz = v.pferz;
  MesonCm C;
  transform_to_cm(v, C);
  MesonWeight w;
  w.thetacm = C.thetacm; w.phicm = C.phicm; w.pcm = C.pcm; w.davejac = C.jacobian; w.johnjac = C.jac_old; w.wcm = C.wcm;
  const double k_eq = (C.wcm * C.wcm - Mtar * Mtar) / 2. / Mtar;
  const double sigcm1 = sig_param_2021(C.thetacm, C.phicm, v.t / 1.e6, v.Q2 / 1.e6, C.sgev / 1.e6, v.epsilon, cfg.which_pion);
  double sigma_eepi = sigcm1;
  w.low_w = false;
  if (C.wcm < 2000) {                       // physics_pion.f:131-154: blend with MAID-2007 below W = 2 GeV
    if (maid.tbl) {
      const double Wgev = C.wcm / 1000.0;
      const double sig0 = sigmaid_sig0(maid, v.Q2 / 1.e6, Wgev, v.Ein / 1000.0, m::cos(C.thetacm), C.phicm);
      const double sigcm2 = sig0 / C.pcm / C.qstar / 2.;
      const double fac1 = fmin(1., fmax(0., (Wgev - 1.5) / 0.4));
      sigma_eepi = sigcm1 * fac1 + sigcm2 * (1 - fac1);
    } else {
      w.low_w = true;                       // table not provided: parametrisation alone, counted in `unsupported`
    }
  }
  w.sigcm = sigma_eepi;
  const double fac = 1. / (1. - pferz * pfer / efer) * Mtar / efer;
  const double gtpr = alpha / 2. / (pi * pi) * v.eE / v.Ein * k_eq / v.Q2 / (1. - v.epsilon);
  w.sigcc = sigma_eepi * C.jacobian * (gtpr * fac);
  return w;
}

// physics_pion.f:404-465: fit to Brauel et al.; dsigma/dt/dphi_cm in ub/MeV^2/rad
SIMC_HD double sig_blok(double thetacm, double phicm, double t, double q2_gev, double s_gev, double eps, double mtar_gev,
                        int which_pion) {
  const double pi = 3.141592653589793;
  double sigl = 27.8 * m::exp(-11.5 * fabs(t));
  double sigt = 10.0 * (5. * fabs(t)) * m::exp(-5. * fabs(t));
  const double siglt = 0.0 * m::sin(thetacm);
  const double sth = m::sin(thetacm);
  double sigtt = -(4.0 * sigl + 0.5 * sigt) * (sth * sth);
  if (which_pion == 1 || which_pion == 11 || which_pion == 3) {
    sigt = sigt * 0.25 * (1. + 3. * m::exp(-10. * fabs(t)));
    sigtt = sigtt * 0.25 * (1. + 3. * m::exp(-10. * fabs(t)));
  }
  const double fpi = 1. / (1. + 1.65 * q2_gev + 0.5 * (q2_gev * q2_gev));
  const double fpi2 = fpi * fpi;
  sigl = sigl * (fpi2 * q2_gev) / 0.1215;
  sigt = sigt / (0.3 + q2_gev);
  sigtt = sigtt / (0.3 + q2_gev);
  const double sig219 = (sigt + eps * sigl + eps * m::cos(2. * phicm) * sigtt + sqrt(2.0 * eps * (1. + eps)) * m::cos(phicm) * siglt) / 1.e0;
  const double d = s_gev - mtar_gev * mtar_gev;
  double sig = sig219 * 15.333 / (d * d);
  sig = sig / 2. / pi / 1.e+06;
  return sig;
}

// physics_delta.f:1-135: H(e,e'p)pi0.  The weight is phase space times the virtual-photon flux (the model value
// multiplies by 1.0, physics_delta.f:131); sig_blok only feeds the ntuple's sigcm column.
SIMC_HD_CALL MesonWeight peedelta(const simc_run_config& cfg, const MesonVertex& v) {
  const double pi = 3.141592653589793, alpha = 1. / 137.0359895;
  const double Mtar = cfg.targ.Mtar_struck, efer = v.efer, pfer = v.pfer, pferz = v.pferz;
  MesonCm C;
  transform_to_cm(v, C);
  MesonWeight w;
  w.thetacm = C.thetacm; w.phicm = C.phicm; w.pcm = C.pcm; w.davejac = C.jacobian; w.johnjac = C.jac_old; w.wcm = C.wcm;
  w.low_w = false;
  const double k_eq = (C.wcm * C.wcm - Mtar * Mtar) / 2. / Mtar;
  w.sigcm = sig_blok(C.thetacm, C.phicm, v.t / 1.e6, v.Q2 / 1.e6, C.sgev / 1.e6, v.epsilon, Mtar / 1000., cfg.which_pion);
  const double fac = 1. / (1. - pferz * pfer / efer) * Mtar / efer;
  const double gtpr = alpha / 2. / (pi * pi) * v.eE / v.Ein * k_eq / v.Q2 / (1. - v.epsilon);
  w.sigcc = 1.0 * C.jacobian * (gtpr * fac);
  return w;
}

// rho_physics.f:1-402: p(e,e'rho)p in the form PYTHIA uses with the HERMES modifications.  sigma_T from a fit to
// photoproduction data, R = sigma_L/sigma_T = 0.33 (Q2/Mrho2)^0.61 and the (Mrho2/(Q2+Mrho2))^2.575 dependence from
// HERMES, an exponential t' slope b(c delta tau), times the virtual-photon flux.  `v` holds the RHO's momentum,
// energy and direction (vertex%p, vertex%up); theta_e is vertex%e%theta.  main%davejac ("full blown Jacobian",
// :281-351) and main%johnjac only reach the ntuple.
SIMC_HD_CALL MesonWeight peerho(const simc_run_config& cfg, const MesonVertex& v, double theta_e) {
  using namespace mesondetail;
  const double pi = 3.141592653589793, alpha = 1. / 137.0359895, hbarc = 197.327053, Me = 0.51099906;
  const double Mrho2 = 769.3 * 769.3;
  const double Mtar = cfg.targ.Mtar_struck;
  const double Q2_g = v.Q2 / 1000000.;
  const double cospq = m::cos(v.phi_pq), sinpq = m::sin(v.phi_pq);
  const double pfer = v.pfer, pferx = v.pferx, pfery = v.pfery, pferz = v.pferz;
  // rho_physics.f:93-98: on-shell struck nucleon (the off-shell forms belong to doing_deutpi / doing_hepi, never set here)
  const double efer = sqrt(pfer * pfer + Mtar * Mtar);
  double tcos = v.upx * v.uqx + v.upy * v.uqy + v.upz * v.uqz;
  if (tcos - 1. > 0. && tcos - 1. < 1.e-8) tcos = 1.0;
  const double tsin = sqrt(1. - tcos * tcos);
  double tfcos = pferx * v.uqx + pfery * v.uqy + pferz * v.uqz;
  if (tfcos - 1. > 0. && tfcos - 1. < 1.e-8) tfcos = 1.0;
  const double tfsin = sqrt(1. - tfcos * tfcos);
  const double th2 = m::tan(theta_e / 2.);
  const double epsi = 1. / (1. + 2 * (1. + v.nu * v.nu / v.Q2) * (th2 * th2));
  double ss = msq(v.nu + efer) - msq(v.q + pfer * tfcos) - msq(pfer * tfsin);
  ss = ss / 1.e6;
  double t = v.Q2 - Mrho2 + 2. * v.nu * v.pE - 2. * v.pP * v.q * tcos;
  t = t / 1.e6;

  const double qx = -v.uqy, qy = v.uqx, qz = v.uqz;
  const double px = -pfery, py = pferx, pz = pferz;
  double dummy = sqrt((qx * qx + qy * qy) * (qx * qx + qy * qy + qz * qz));
  double new_x_x = -qx * qz / dummy, new_x_y = -qy * qz / dummy, new_x_z = (qx * qx + qy * qy) / dummy;
  dummy = sqrt(qx * qx + qy * qy);
  double new_y_x = qy / dummy, new_y_y = -qx / dummy, new_y_z = 0.0;
  const double p_new_x = pfer * (px * new_x_x + py * new_x_y + pz * new_x_z);
  const double p_new_y = pfer * (px * new_y_x + py * new_y_y + pz * new_y_z);
  double phiqn;
  if (p_new_x == 0.) phiqn = 0.;
  else phiqn = m::atan2(p_new_y, p_new_x);
  if (phiqn < 0.) phiqn = phiqn + 2. * pi;
  const double cosqn = m::cos(phiqn), sinqn = m::sin(phiqn);

  const double pbeam = sqrt(v.Ein * v.Ein - Me * Me);
  const double beam_newx = pbeam * new_x_z, beam_newy = pbeam * new_y_z, beam_newz = pbeam * v.uqz;
  const double bstar = sqrt(msq(v.q + pfer * tfcos) + msq(pfer * tfsin)) / (efer + v.nu);
  const double gstar = 1. / sqrt(1. - bstar * bstar);
  const double bstarz = (v.q + pfer * tfcos) / (efer + v.nu);
  const double bstarx = p_new_x / (efer + v.nu);
  const double bstary = p_new_y / (efer + v.nu);
  const MV4 qs = loren(gstar, bstarx, bstary, bstarz, v.nu, 0.e0, 0.e0, v.q);
  const double ppiz = v.pP * tcos, ppix = v.pP * tsin * cospq, ppiy = v.pP * tsin * sinpq;
  const MV4 had = loren(gstar, bstarx, bstary, bstarz, v.pE, ppix, ppiy, ppiz);
  const double thetacm = m::acos((had.x * qs.x + had.y * qs.y + had.z * qs.z) / had.p / qs.p);
  const MV4 beam = loren(gstar, bstarx, bstary, bstarz, v.Ein, beam_newx, beam_newy, beam_newz);
  dummy = sqrt(msq(qs.y * beam.z - qs.z * beam.y) + msq(qs.z * beam.x - qs.x * beam.z) + msq(qs.x * beam.y - qs.y * beam.x));
  new_y_x = (qs.y * beam.z - qs.z * beam.y) / dummy;
  new_y_y = (qs.z * beam.x - qs.x * beam.z) / dummy;
  new_y_z = (qs.x * beam.y - qs.y * beam.x) / dummy;
  dummy = sqrt(msq(new_y_y * qs.z - new_y_z * qs.y) + msq(new_y_z * qs.x - new_y_x * qs.z) + msq(new_y_x * qs.y - new_y_y * qs.x));
  new_x_x = (new_y_y * qs.z - new_y_z * qs.y) / dummy;
  new_x_y = (new_y_z * qs.x - new_y_x * qs.z) / dummy;
  new_x_z = (new_y_x * qs.y - new_y_y * qs.x) / dummy;
  const double ppicm_newx = had.x * new_x_x + had.y * new_x_y + had.z * new_x_z;
  const double ppicm_newy = had.x * new_y_x + had.y * new_y_y + had.z * new_y_z;
  double phicm = m::atan2(ppicm_newy, ppicm_newx);
  if (phicm < 0.) phicm = 2. * 3.141592654 + phicm;

  const double mt = Mtar / 1000.;
  const double tmin = -(msq((-Q2_g - Mrho2 / 1.e6 - mt * mt + mt * mt) / (2. * sqrt(ss))) - msq((qs.p - had.p) / 1000.));
  const double tprime = fabs(t - tmin);
  const double sig0 = 41.263 / m::pow(v.nu / 1000.0, 0.4765);
  double R = 0.33 * m::pow(v.Q2 / Mrho2, 0.61);
  if (R < 0.) R = 0.;
  double sigt = sig0 * (1.0 + epsi * R) * m::pow(Mrho2 / (v.Q2 + Mrho2), 2.575);
  if (sigt < 0.) sigt = 0.;
  const double cdeltatau = hbarc / (sqrt(v.nu * v.nu + v.Q2 + Mrho2) - v.nu);
  double brho;
  if (cdeltatau < 2.0) {
    brho = 4.4679 + 8.6106 * m::log10(cdeltatau);
    if (brho < 1.0) brho = 1.0;
  } else {
    brho = 7.0;
  }
  const double sig219 = sigt * brho * m::exp(-brho * tprime) / 2.0 / pi;
  double sig = sig219 / 1.e+06;
  sig = sig * 2. * qs.p * had.p;
  double gtpr = alpha / 2. / (pi * pi) * v.eE / v.Ein * (ss - mt * mt) / 2. / ((efer - pfer * tfcos) / 1000.) / Q2_g / (1. - epsi);
  if (gtpr <= 0.) gtpr = 0.;

  const double psign = cosqn * cospq + sinqn * sinpq;
  const double P = v.pP, E = v.pE;
  const double square_root = v.q + pfer * tfcos - P * tcos;
  const double dp_dcos_num = P + (P * P * tcos - psign * pfer * P * tfsin * tcos / tsin) / square_root;
  const double dp_dcos_den = ((v.nu + efer - E) * P / E + P * tsin * tsin - psign * pfer * tfsin * tsin) / square_root - tcos;
  const double dp_dcos = dp_dcos_num / dp_dcos_den;
  const double dp_dphi_num = pfer * P * tsin * tfsin * (cosqn * sinpq - sinqn * cospq) / square_root;
  const double dp_dphi_den = tcos + (pfer * tsin * tfsin * psign - P * tsin * tsin - (v.nu + efer - E) * P / E) / square_root;
  const double dp_dphi = dp_dphi_num / dp_dphi_den;
  const double dt_dcos_lab = 2. * (v.q * P + (v.q * tcos - v.nu * P / E) * dp_dcos);
  const double dt_dphi_lab = 2. * (v.q * tcos - v.nu * P / E) * dp_dphi;
  const double b2 = bstar * bstar;
  const double dpxdphi = P * tsin * (-sinpq + (gstar - 1.) * bstarx / b2 * (bstary * cospq - bstarx * sinpq)) +
                         ((had.x + gstar * bstarx * E) / P - gstar * bstarx * P / E) * dp_dphi;
  const double dpydphi = P * tsin * (cospq + (gstar - 1.) * bstary / b2 * (bstary * cospq - bstarx * sinpq)) +
                         ((had.y + gstar * bstary * E) / P - gstar * bstary * P / E) * dp_dphi;
  const double dpzdphi = P * (gstar - 1.) / b2 * bstarz * tsin * (bstary * cospq - bstarx * sinpq) +
                         ((had.z + gstar * bstarz * E) / P - gstar * bstarz * P / E) * dp_dphi;
  const double dpxdcos = -P * tcos / tsin * (cospq + (gstar - 1.) * bstarx / b2 * (bstarx * cospq + bstary * sinpq - bstarz * tsin / tcos)) +
                         ((had.x + gstar * bstarx * E) / P - gstar * bstarx * P / E) * dp_dcos;
  const double dpydcos = -P * tcos / tsin * (sinpq + (gstar - 1.) * bstary / b2 * (bstarx * cospq + bstary * sinpq - bstarz * tsin / tcos)) +
                         ((had.y + gstar * bstary * E) / P - gstar * bstary * P / E) * dp_dcos;
  const double dpzdcos = P * (1. - (gstar - 1.) / b2 * bstarz * tcos / tsin * (bstarx * cospq + bstary * sinpq - tsin / tcos * bstarz)) +
                         ((had.z + gstar * bstarz * E) / P - gstar * bstarz * P / E) * dp_dcos;
  const double dpxnewdphi = dpxdphi * new_x_x + dpydphi * new_x_y + dpzdphi * new_x_z;
  const double dpynewdphi = dpxdphi * new_y_x + dpydphi * new_y_y + dpzdphi * new_y_z;
  const double den = ppicm_newx * ppicm_newx + ppicm_newy * ppicm_newy;
  const double dphicmdphi = (dpynewdphi * ppicm_newx - ppicm_newy * dpxnewdphi) / den;
  const double dpxnewdcos = dpxdcos * new_x_x + dpydcos * new_x_y + dpzdcos * new_x_z;
  const double dpynewdcos = dpxdcos * new_y_x + dpydcos * new_y_y + dpzdcos * new_y_z;
  const double dphicmdcos = (dpynewdcos * ppicm_newx - ppicm_newy * dpxnewdcos) / den;

  MesonWeight w;
  w.thetacm = thetacm; w.phicm = phicm; w.pcm = had.p; w.wcm = 0.0; w.sigcm1 = 0.0; w.low_w = false;
  w.davejac = fabs(dt_dcos_lab * dphicmdphi - dt_dphi_lab * dphicmdcos);
  w.johnjac = 2 * (efer - 2 * pferz * pfer * E / P * tcos) * (v.q + pferz * pfer) * P /
                  (efer + v.nu - (v.q + pferz * pfer) * E / P * tcos) -
              2 * P * pfer;
  w.t_gev = t;
  double sigma_eerho = gtpr * sig / 1.e3;
  if (sigma_eerho > 1.E10 || sigma_eerho < 0.) sigma_eerho = 0.;
  w.sigcc = sigma_eerho;
  w.sigcm = sig;
  return w;
}

// Saghai amplitude tables on the device: buf = [50 grid values (the `pa` array of eekeek / eekeeks) | 14 pad |
// 12 tables of n1*n2*n3 REAL*4 in Fortran storage order: zrff1..6 then ziff1..6].  Null unless set.
struct SaghaiDev { const float* buf; int n1, n2, n3; };

// One argument of CERNLIB's fint (cern/fint.f:24-53) on a grid of more than two points: the cell and the
// interpolation fraction, or the node itself when the argument sits on one (no split of the knots then).
// X, ETA are default REAL (8 bytes under the reference's -fdefault-real-8), ENT is REAL*4 and the grid spacing is a
// REAL*4 difference.
struct FintAxis { int cell; double eta; bool split; };
SIMC_HD FintAxis fint_axis(const float* ent, int n, float arg) {
  const double x = arg;
  FintAxis a;
  int lo = -1, hi = n;                    // bisection like the source: ent[lo] < x < ent[hi] (virtual ends)
  while (hi - lo > 1) {
    const int mid = (lo + hi) / 2;
    const double d = x - (double)ent[mid];
    if (d == 0.) { a.cell = mid; a.eta = 0.; a.split = false; return a; }
    if (d < 0.) hi = mid; else lo = mid;
  }
  lo = lo < 0 ? 0 : (lo > n - 2 ? n - 2 : lo);
  a.cell = lo;
  a.eta = (x - (double)ent[lo]) / (double)(ent[lo + 1] - ent[lo]);
  a.split = true;
  return a;
}

// physics_kaon.f:241-355 (K+ Lambda) / 357-489 (K+ Sigma0): the two routines differ in their grids only, which the
// host wrote in front of the tables.  All twelve fint calls share their arguments, so cell and weights are made once;
// the knots are visited in fint's order (first argument fastest) with fint's weights w - w*eta and w*eta.
SIMC_HD_CALL double saghai_sigma(const SaghaiDev S, double mrec_struck, double ss, double q22, double angl, double theta,
                                 double phi, double epsi) {
  const double pi = 3.141592653589793, Mk2 = 493.677 * 493.677, Mp = 938.27231, Mp2 = Mp * Mp, hbarc = 197.327053;
  const double w = sqrt(ss) * 1000.;
  double skc2 = mesondetail::msq(w * w - Mk2 - mrec_struck * mrec_struck) - 4. * Mk2 * (mrec_struck * mrec_struck);
  skc2 = skc2 > 0. ? skc2 : 0.;
  const double skc = sqrt(skc2) / 2. / w;
  const double q0 = -(-q22 - w * w + Mp2) / 2. / Mp;
  const double q0c = (-q22 + q0 * Mp) / w;
  const double qr = sqrt(q22) / q0c;
  const double aflx = skc / 2. / w / (w * w - Mp2) * (hbarc * hbarc) * 10000.;
  const double aflxl = aflx * (qr * qr);
  const double an = angl * 180. / pi;
  const double x = m::cos(angl), sx = m::sin(angl);
  const float* ent = S.buf;
  const FintAxis a1 = fint_axis(ent, S.n1, (float)ss);
  const FintAxis a2 = fint_axis(ent + S.n1, S.n2, (float)(q22 / 1.e+06));
  const FintAxis a3 = fint_axis(ent + S.n1 + S.n2, S.n3, (float)an);
  int idx[8];
  double wt[8];
  int knots = 1, istep = 1;
  idx[0] = 0; wt[0] = 1.;
  const FintAxis ax[3] = {a1, a2, a3};
  const int nd[3] = {S.n1, S.n2, S.n3};
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const int shift = ax[d].cell * istep;
    if (!ax[d].split) {
      for (int k = 0; k < knots; ++k) idx[k] += shift;
    } else {
      for (int k = 0; k < knots; ++k) {
        idx[k] += shift;
        idx[k + knots] = idx[k] + istep;
        wt[k + knots] = wt[k] * ax[d].eta;
        wt[k] = wt[k] - wt[k + knots];
      }
      knots *= 2;
    }
    istep *= nd[d];
  }
  const size_t n_tab = (size_t)S.n1 * S.n2 * S.n3;
  const float* tab = S.buf + 64;
  double zr[6], zi[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    double r = 0., i = 0.;
    for (int j = 0; j < knots; ++j) {
      r = r + wt[j] * (double)tab[(size_t)k * n_tab + idx[j]];
      i = i + wt[j] * (double)tab[(size_t)(6 + k) * n_tab + idx[j]];
    }
    zr[k] = r; zi[k] = i;
  }
  // amplitudes z1, z2, z3, z4, z7, z8 = slots 0..5; abs(z)**2, real(conjg(a)*b) written out
  auto abs2 = [](double re, double im) { const double a = hypot(re, im); return a * a; };
  auto rcab = [](double ar, double ai, double br, double bi) { return ar * br + ai * bi; };          // real(conjg(a)*b)
  const double a1_ = abs2(zr[0], zi[0]), a2_ = abs2(zr[1], zi[1]), a3_ = abs2(zr[2], zi[2]), a4_ = abs2(zr[3], zi[3]);
  const double a7_ = abs2(zr[4], zi[4]), a8_ = abs2(zr[5], zi[5]);
  // real(conjg(z1)*z4 - conjg(z2)*z3 + conjg(z3)*z4*x), left to right like the complex expression
  const double mix = (rcab(zr[0], zi[0], zr[3], zi[3]) - rcab(zr[1], zi[1], zr[2], zi[2])) + rcab(zr[2], zi[2], zr[3], zi[3]) * x;
  const double dsigt00 = aflx * (a1_ + a2_ + 2. * rcab(zr[0], zi[0], zr[1], zi[1]) * x + 0.5 * (sx * sx) * (a3_ + a4_ + 2. * mix));
  const double dsigl00 = aflxl * epsi * (a7_ + a8_ + 2. * rcab(zr[4], zi[4], zr[5], zi[5]) * x);
  const double sth = m::sin(theta);
  const double dsigp00 = aflx * epsi * (sth * sth) * m::cos(2. * phi) * (0.5 * a3_ + 0.5 * a4_ + mix);
  // real(z7*(conjg(z3) - conjg(z2) + conjg(z4)*x) + z8*(conjg(z1) + conjg(z3)*x + conjg(z4)))
  const double br = (zr[2] - zr[1]) + zr[3] * x, bi = (-zi[2] + zi[1]) + (-zi[3]) * x;
  const double cr = (zr[0] + zr[2] * x) + zr[3], ci = (-zi[0] + (-zi[2]) * x) + (-zi[3]);
  const double inter = (zr[4] * br - zi[4] * bi) + (zr[5] * cr - zi[5] * ci);
  const double dsigi00 = aflx * sqrt(2. * (qr * qr) * epsi * (1. + epsi)) * sth * m::cos(phi) * inter;
  return dsigt00 + dsigl00 + dsigp00 + dsigi00;
}

// physics_kaon.f:1-171 (without the survival probability, which needs the focal-plane track)
// saghai.buf != null: also evaluate the Saghai model at main%theta_pq (physics_kaon.f:100-108; `phi` is an
// unassigned static local there, i.e. zero: "WE ARE ALWAYS CALCULATING FOR PHI=0")
SIMC_HD_CALL MesonWeight peeK(const simc_run_config& cfg, const MesonVertex& v, const SaghaiDev saghai = SaghaiDev{nullptr, 0, 0, 0},
                              double theta_pq = 0.) {
  const double pi = 3.141592653589793, alpha = 1. / 137.0359895;
  const double Mtar = cfg.targ.Mtar_struck, efer = v.efer, pfer = v.pfer, pferz = v.pferz;
  MesonCm C;
  transform_to_cm(v, C);
  MesonWeight w;
  const double jacobian = C.jacobian / (2. * C.pcm * C.qstar);
  const double jac_old = C.jac_old / (2. * C.pcm * C.qstar);
  w.thetacm = C.thetacm; w.phicm = C.phicm; w.pcm = C.pcm; w.davejac = jacobian; w.johnjac = jac_old; w.wcm = C.wcm;
  w.low_w = false;
  w.sigcm1 = 0.;
  if (saghai.buf) w.sigcm1 = saghai_sigma(saghai, cfg.targ.Mrec_struck, C.sgev / 1.e6, v.Q2, C.thetacm, theta_pq, 0.0, v.epsilon);
  const double sigcm2 = sig_factorized(v.Q2, C.wcm, v.t, C.pcm, cfg.targ.Mrec_struck);
  w.sigcm = sigcm2;
  const double k_eq = (C.wcm * C.wcm - Mtar * Mtar) / 2. / Mtar;
  const double fac = 1. / (1. - pferz * pfer / efer) * Mtar / efer;
  const double gtpr = alpha / 2. / (pi * pi) * v.eE / v.Ein * k_eq / v.Q2 / (1. - v.epsilon);
  w.sigcc = sigcm2 * jacobian * (gtpr * fac);
  return w;
}

// physics_kaon.f:148-165: survival probability when decay is not simulated
SIMC_HD_CALL double kaon_survival(const simc_run_config& cfg, double fp_path, double fp_dx, double fp_dy) {
  double zaero = 0.;
  if (cfg.hadron_arm == 2) zaero = -82.8;
  else if (cfg.hadron_arm == 3 || cfg.hadron_arm == 4) zaero = -183.;
  const double pathlen = fp_path + zaero * (1 + fp_dx * fp_dx + fp_dy * fp_dy);
  const double betak = cfg.spec_p.P / sqrt(cfg.spec_p.P * cfg.spec_p.P + cfg.Mh2);
  const double gammak = 1. / sqrt(1. - betak * betak);
  return 1. / m::exp(pathlen / (cfg.ctau * betak * gammak));
}

}  // namespace simc
