// Radiative corrections on the device: per-event constants (radc_init_ev, basicrad_init_ev,
// init.f:655-813), soft-photon factors (bremos / inter / inter_prime / spence, brem.f:216-596),
// photon-energy sampling and weights (basicrad, peaked_rad_weight, extrad_phi, gamma,
// lambda_dave; radc.f), plus the option branches no shipped deck selects: the on-shell brem (brem.f:6-214,
// use_offshell_rad = 0), schwinger (radc.f:711, intcor_mode = 0), the Friedrich form of extrad_phi
// (extrad_flag = 3) and the (Egamma1, Egamma2, Egamma3) basis of rad_flag = 2, 3.  Those live in out-of-line
// functions behind run-constant branches, so the default path (rad_flag 0, extrad_flag 2, intcor_mode 1,
// use_offshell_rad 1) executes what it executed before.
// The reference is built with -fno-automatic (Makefile:63): a local that an option setting never assigns
// (dsoft_prime under intcor_mode = 0; dsoft_intmin/max under rad_flag = 1 with extrad_flag = 3) is static and
// reads 0.0.
#pragma once
#include <math.h>
#include "target.cuh"

namespace simc {

struct RadEvDev {                         // per-event block of COMMON /radccom/ (radc.inc:5-19)
  double bt[2], lambda[3], frac[3];
  double g[5], c4, c_ext0, g_ext, hardcorfac;
  double Egamma_used[3];
  int ntail;
  bool rad_proton_this_ev;
};

struct V4 { double e, x, y, z; };

SIMC_HD double sq(double x) { return x * x; }

// brem.f:286-311 (the "saves a WHALE of CPU" approximation)
SIMC_HD double spence_approx(double ax) {
  const double bx = fabs(ax);
  if (bx <= 1) return 0.0;
  const double l = m::log(bx);
  return -0.5 * (l * l);
}

// One interference term: inter (brem.f:216-240) and inter_prime (brem.f:581-596) share
// amult and the two log-ratio factors.  WHAT selects the outputs a caller reads (the arithmetic of each
// is the reference's, untouched): kBremPrime = d(b)/dE only (radc_init_ev reads nothing else of bremos
// besides bhard), kBremSoft = b only (peaked_rad_weight), kBremAll = both.
constexpr int kBremAll = 0, kBremPrime = 1, kBremSoft = 2;
template <int WHAT>
SIMC_HD_CALL void inter_pair(double alpha, double ar1, double ar2, double e1, double e2, double de, double& val,
                        double& prime) {
  const double pi = 3.141592653589793;
  const double de2 = e1 - e2;
  const double amult = -1. / (alpha * (ar1 - ar2));
  const double l1 = m::log(fabs((ar1 - 1.) / ar1));
  const double l2 = m::log(fabs((ar2 - 1.) / ar2));
  val = 0.0; prime = 0.0;
  if (WHAT != kBremPrime) {
    double v = m::log(fabs((e2 / de) + ar1 * (de2 / de))) * l1 - m::log(fabs((e2 / de) + ar2 * (de2 / de))) * l2;
    const double arg1 = (de2 / (e2 + ar1 * de2)) * (ar1 - 1.);
    const double arg2 = (de2 / (e2 + ar1 * de2)) * (ar1);
    const double arg3 = (de2 / (e2 + ar2 * de2)) * (ar2 - 1.);
    const double arg4 = (de2 / (e2 + ar2 * de2)) * (ar2);
    v = v - spence_approx(arg1) + spence_approx(arg2) + spence_approx(arg3) - spence_approx(arg4);
    val = v * amult / (pi);
  }
  if (WHAT != kBremSoft) prime = (-1. / de) * amult / pi * (l1 - l2);
}

SIMC_HD double dot4(const V4& a, const V4& b) { return a.e * b.e - a.x * b.x - a.y * b.y - a.z * b.z; }

// One of the six interference contributions of bremos (brem.f:420-491): masses m1 (the one whose
// square enters the numerator of ar1/ar2) and m2.
template <int WHAT>
SIMC_HD void interference(double aprod, const V4& a, const V4& b, double m_num, double m_oth, double de, double& bsum,
                          double& dbsum) {
  const double adot = dot4(a, b);
  const double alpha = m_num * m_num + m_oth * m_oth - 2. * adot;
  const double mm = m_oth * m_num;   // (ame*ami) etc.: the reference always writes (small*large)
  const double root = sqrt(4. * adot * adot - 4. * (mm * mm));
  const double ar1 = (2. * m_num * m_num - 2. * adot + root) / (2. * alpha);
  const double ar2 = (2. * m_num * m_num - 2. * adot - root) / (2. * alpha);
  double v, p;
  inter_pair<WHAT>(alpha, ar1, ar2, a.e, b.e, de, v, p);
  bsum = aprod * adot * v;
  dbsum = aprod * adot * p;
}

// bremos, brem.f:344-577, with the electron along +z before scattering (k_i = (0,0,Ein)) and the
// initial proton at rest, which is how both call sites use it (init.f:706, radc.f:600-623).
// Energies in MeV on input.  Returns bsoft, bhard and d(bsoft)/dE (per MeV).
template <int WHAT = kBremAll>
SIMC_HD_CALL void bremos(double egamma, double ein, double kfx, double kfy, double kfz, double pfx, double pfy, double pfz,
                    double pfe, bool radiate_proton, double& bsoft, double& bhard, double& dbsoft) {
  const double pi = 3.141592653589793, twopi = 2. * pi, ame = .00051099906, e2 = 1. / 137.0359895, mp = .93827231;
  const double de = egamma / 1000.;
  V4 k_i, k_f, p_i, p_f;
  k_i.x = 0. / 1000.; k_i.y = 0. / 1000.; k_i.z = ein / 1000.;
  k_f.x = kfx / 1000.; k_f.y = kfy / 1000.; k_f.z = kfz / 1000.;
  p_i.x = 0.; p_i.y = 0.; p_i.z = 0.;
  p_f.e = pfe / 1000.; p_f.x = pfx / 1000.; p_f.y = pfy / 1000.; p_f.z = pfz / 1000.;
  // (...)**0.5 in the reference: m::pow(x,0.5) and sqrt(x) agree except for rare last-bit cases
  k_i.e = sqrt(k_i.x * k_i.x + k_i.y * k_i.y + k_i.z * k_i.z + ame * ame);
  k_f.e = sqrt(k_f.x * k_f.x + k_f.y * k_f.y + k_f.z * k_f.z + ame * ame);
  p_i.e = mp;
  const double q2 = -1. * (sq(k_f.e - k_i.e) - sq(k_f.x - k_i.x) - sq(k_f.y - k_i.y) - sq(k_f.z - k_i.z));
  const double ami = mp;
  const double amf = sqrt(sq(p_f.e) - sq(p_f.x) - sq(p_f.y) - sq(p_f.z));
  const double bei = WHAT == kBremPrime ? 0.0 : 1.e0 * (-1. / twopi) * m::log(k_i.e / de);
  const double dbei = 1.e0 * (-1. / twopi) * (-1. / de);
  const double bef = WHAT == kBremPrime ? 0.0 : 1.e0 * (-1. / twopi) * m::log(k_f.e / de);
  const double dbef = 1.e0 * (-1. / twopi) * (-1. / de);
  double bee, dbee;
  {   // e-e interference, brem.f:411-418 (its own form of ar1/ar2)
    const double adot = dot4(k_i, k_f);
    const double alpha = 2. * ame * ame - 2. * adot;
    const double ame2 = ame * ame;
    const double root = sqrt(4. * adot * adot - 4. * (ame2 * ame2));
    const double ar1 = 0.5 + root / (2. * alpha);
    const double ar2 = 0.5 - root / (2. * alpha);
    double v, p;
    inter_pair<WHAT>(alpha, ar1, ar2, k_i.e, k_f.e, de, v, p);
    bee = -1.e0 * adot * v;
    dbee = -1.e0 * adot * p;
  }
  const double b = 2. * e2 * (bei + bef + bee);
  const double db = 2. * e2 * (dbei + dbef + dbee);
  double bz = 0.0, bzz = 0.0, dbz = 0.0, dbzz = 0.0;
  if (radiate_proton) {
    const double bpi = WHAT == kBremPrime ? 0.0 : 1.e0 * (-1. / twopi) * m::log(p_i.e / de);
    const double dbpi = 1.e0 * (-1. / twopi) * (-1. / de);
    const double bpf = WHAT == kBremPrime ? 0.0 : 1.e0 * (-1. / twopi) * m::log(p_f.e / de);
    const double dbpf = 1.e0 * (-1. / twopi) * (-1. / de);
    double bpp, dbpp, bepii, dbepii, bepff, dbepff, bepif, dbepif, bepfi, dbepfi;
    // alpha = ami^2+amf^2-2adot, ar uses 2 amf^2, root uses (ami*amf)^2           brem.f:430-437
    {
      const double adot = dot4(p_i, p_f);
      const double alpha = ami * ami + amf * amf - 2. * adot;
      const double mm = ami * amf;
      const double root = sqrt(4. * adot * adot - 4. * (mm * mm));
      const double ar1 = (2. * amf * amf - 2. * adot + root) / (2. * alpha);
      const double ar2 = (2. * amf * amf - 2. * adot - root) / (2. * alpha);
      double v, p;
      inter_pair<WHAT>(alpha, ar1, ar2, p_i.e, p_f.e, de, v, p);
      bpp = -1.e0 * adot * v; dbpp = -1.e0 * adot * p;
    }
    interference<WHAT>(-1.e0, k_i, p_i, ami, ame, de, bepii, dbepii);   // ei-pi  brem.f:439-446
    interference<WHAT>(-1.e0, k_f, p_f, amf, ame, de, bepff, dbepff);   // ef-pf  brem.f:448-455
    interference<WHAT>(1.e0, k_i, p_f, amf, ame, de, bepif, dbepif);    // ei-pf  brem.f:457-464
    interference<WHAT>(1.e0, k_f, p_i, ami, ame, de, bepfi, dbepfi);    // ef-pi  brem.f:466-473
    bzz = 2. * e2 * (bpi + bpf + bpp);
    bz = 2. * e2 * (bepii + bepff + bepif + bepfi);
    dbzz = 2. * e2 * (dbpi + dbpf + dbpp);
    dbz = 2. * e2 * (dbepii + dbepff + dbepif + dbepfi);
  }
  bsoft = b + bz + bzz;
  bhard = WHAT == kBremSoft ? 0.0 : -1. * (e2 / pi) * (-28 / 9. + 13. / 6. * m::log(q2 / (ame * ame)));
  dbsoft = db + dbz + dbzz;
  dbsoft = dbsoft / 1000.;
}

// brem, brem.f:6-214: the on-shell soft-photon calculation (elastic e-p kinematics rebuilt from ein, eout).
// Same WHAT convention as bremos.  include_hard = calculate_spence = .true. (init.f:646-647).
template <int WHAT>
SIMC_HD_CALL void brem_onshell(double ein, double eout, double egamma, bool radiate_proton, double& bsoft, double& bhard,
                               double& dbsoft) {
  const double pi = 3.141592653589793, am = .93827231, ame = .00051099906, e2 = 1. / 137.0359895;
  const double ak = ein / 1000., akp = eout / 1000., de = egamma / 1000.;
  // (...)**0.5 in the reference: sqrt, as in bremos above (glibc pow(x, 0.5) is correctly rounded but for rare arguments)
  const double eang = 2. * m::asin(sqrt(am / (2. * ak) * (ak / akp - 1.)));
  const double sh = m::sin(eang / 2.);
  const double q2 = 4. * ak * akp * (sh * sh);
  const double ape = am + ak - akp;
  const double ap = sqrt(ape * ape - am * am);
  const double pang = m::acos((ak - akp * m::cos(eang)) / ap);
  const double ame2 = ame * ame, ame4 = ame2 * ame2, am2 = am * am, am4 = am2 * am2;
  // one interference term: value through inter_pair (= inter, brem.f:216-240); the derivative as brem writes it,
  // aprod*adot/(pi*alpha*(ar1-ar2)*de)*(log((ar1-1)/ar1)-log((ar2-1)/ar2)), without inter_prime's abs()
  auto term = [&](double aprod, double adot, double alpha, double ar1, double ar2, double e1, double e2_, double& b,
                  double& db) {
    double v = 0.0, p = 0.0;
    if (WHAT != kBremPrime) inter_pair<kBremSoft>(alpha, ar1, ar2, e1, e2_, de, v, p);
    b = aprod * adot * v;
    db = 0.0;
    if (WHAT != kBremSoft)
      db = aprod * adot / (pi * alpha * (ar1 - ar2) * de) * (m::log((ar1 - 1) / ar1) - m::log((ar2 - 1) / ar2));
  };
  const double bei = WHAT == kBremPrime ? 0.0 : 1.e0 * (-1. / (2. * pi)) * m::log(ak / de);
  const double dbei = 1.e0 * (1. / (2. * pi * de));
  const double bef = WHAT == kBremPrime ? 0.0 : 1.e0 * (-1. / (2. * pi)) * m::log(akp / de);
  const double dbef = 1.e0 * (1. / (2. * pi * de));
  double bee, dbee;
  {
    const double adot = ak * akp * (1. - m::cos(eang));
    const double alpha = 2. * ame2 - 2. * adot;
    const double root = sqrt(adot * adot - ame4);
    term(-1.e0, adot, alpha, 0.5 + root / alpha, 0.5 - root / alpha, ak, akp, bee, dbee);
  }
  double bz = 0.0, bzz = 0.0, dbz = 0.0, dbzz = 0.0;
  if (radiate_proton) {
    const double bpi = WHAT == kBremPrime ? 0.0 : 1.e0 * (-1. / (2. * pi)) * m::log(am / de);
    const double dbpi = 1.e0 * (1. / (2. * pi * de));
    const double bpf = WHAT == kBremPrime ? 0.0 : 1.e0 * (-1. / (2. * pi)) * m::log(ape / de);
    const double dbpf = 1.e0 * (1 / (2. * pi * de));
    double bpp, dbpp, bepii, dbepii, bepff, dbepff, bepif, dbepif, bepfi, dbepfi;
    {
      const double adot = am * ape;
      const double alpha = 2. * am2 - 2. * adot;
      const double root = sqrt(adot * adot - am4);
      term(-1.e0, adot, alpha, 0.5 + root / alpha, 0.5 - root / alpha, am, ape, bpp, dbpp);
    }
    const double mm = ame * am, mm2 = mm * mm;
    auto ep = [&](double aprod, double adot, double e1, double e2_, double& b, double& db) {
      const double alpha = am2 + ame2 - 2. * adot;
      const double root = sqrt(adot * adot - mm2);
      term(aprod, adot, alpha, (am2 - adot + root) / alpha, (am2 - adot - root) / alpha, e1, e2_, b, db);
    };
    ep(-1.e0, ak * am, ak, am, bepii, dbepii);
    ep(-1.e0, akp * ape - akp * ap * m::cos(eang + pang), akp, ape, bepff, dbepff);
    ep(1.e0, ak * ape - ak * ap * m::cos(pang), ak, ape, bepif, dbepif);
    ep(1.e0, akp * am, akp, am, bepfi, dbepfi);
    bzz = 2. * e2 * (bpi + bpf + bpp);
    bz = 2. * e2 * (bepii + bepff + bepif + bepfi);
    dbzz = 2. * e2 * (dbpi + dbpf + dbpp);
    dbz = 2. * e2 * (dbepii + dbepff + dbepif + dbepfi);
  }
  const double b = 2. * e2 * (bei + bef + bee);
  const double db = 2. * e2 * (dbei + dbef + dbee);
  bsoft = b + bz + bzz;
  bhard = -1. * (e2 / pi) * (-28 / 9. + 13. / 6. * m::log(q2 / ame2));
  dbsoft = db + dbz + dbzz;
  dbsoft = dbsoft / 1000.;
}

// spen, radc.f:746-764 (Abramowitz & Stegun 27.7.2)
SIMC_HD_CALL double spen(double x) {
  double y = 1.0, s = 0.0;
  int i = 0;
  while (i <= 100 && fabs(y) > fabs(s) * 1.e-4) {
    i = i + 1;
    y = x * y;
    s = s + y / (double)(i * i);
  }
  return s;
}

// schwinger, radc.f:711-742 (etta = 1, init.f:688).  The function value itself is read by no caller.
SIMC_HD_CALL void schwinger(const simc_run_config& cfg, double Ecutoff, double Ein, double eE, double etheta, double Q2,
                            double nu, double& dsoft, double& dhard) {
  const double alpi = (1. / 137.0359895) / 3.141592653589793, Me = 0.51099906, amu = 931.49432, etta = 1.0;
  const double lq = m::log(Q2 / (Me * Me)) - 1.0;
  const double sh = m::sin(etheta / 2.);
  const double s2 = sh * sh;
  const double b = 1. + 2. * nu * s2 / (cfg.targ.A * amu);
  const double spence = spen(1. - s2) - 2.5893784;
  dsoft = alpi * lq * m::log(Ein / (etta * etta) * eE * b / (Ecutoff * Ecutoff));
  const double lr = m::log(Ein / eE);
  dhard = -alpi * (2.166666 * lq + spence - (lr * lr) / 2.0);
}

// extrad_friedrich, radc.f:650-664
SIMC_HD_CALL void extrad_friedrich(double etatzai, double Ei, double Ecutoff, double trad, double& dbrem, double& dbrem_prime) {
  const double x = Ecutoff / Ei;
  dbrem = trad * (-(etatzai - 0.5) - etatzai * m::log(x) + etatzai * x - 0.5 * (x * x));
  dbrem_prime = -trad / Ei * (etatzai / x - etatzai + x);
}

// radc.f:92-116
SIMC_HD_CALL double gamma_fn(double x) {
  double g = 1.0;
  const int n = (int)round((x - 1) - 0.5);
  const double y = x - 1 - n;
  if (n > 0) {
    for (int i = 1; i <= n; ++i) g = g * (y + 1 + i);
  } else if (n < 0) {
    for (int i = -1; i >= n; --i) g = g * (1.0 / (y + 1 + i));
  }
  const double y2 = y * y, y3 = y * y2, y4 = y2 * y2, y5 = y * y4;
  g = g * (1. - 0.5748646 * y + 0.9512363 * y2 - 0.6998588 * y3 + 0.4245549 * y4 - 0.1010678 * y5);
  return g;
}

// radc.f:768-821
SIMC_HD_CALL double lambda_dave(int itail, bool doing_proton, double e1, double e2, double e3, double p3, double th) {
  const double alpi = (1. / 137.0359895) / 3.141592653589793, Me = 0.51099906;
  double plus_term = 0.0;
  if (itail < 3) {
    plus_term = m::log((1. - m::cos(th)) / 2.);
    if (doing_proton) plus_term = plus_term + 2. * m::log(e1 / e2);
  }
  if (itail == 1) return alpi * (2. * m::log(2. * e1 / Me) - 1. + plus_term);
  if (itail == 2) return alpi * (2. * m::log(2. * e2 / Me) - 1. + plus_term);
  if (doing_proton) {
    const double v = alpi * ((e3 / p3) * m::log((e3 + p3) / (e3 - p3)) - 2.);
    return v < 0 ? 0.0 : v;
  }
  return 0.0;
}

struct VertexKin {     // what the radiative routines read from `vertex`
  double Ein, eE, eP, etheta, pE, pP;
  double uex, uey, uez, upx, upy, upz;
};

// radc_init_ev + basicrad_init_ev, init.f:655-813.  Only the constants that the live branches
// read afterwards are kept: g(0..4), c(4), c_ext(0), g_ext, hardcorfac, frac.
SIMC_HD_CALL void radc_init_ev(const simc_run_config& cfg, const VertexKin& v, double teff1, double teff2, RadEvDev& R) {
  const double Mp = 938.27231, euler = 0.577215665;
  R.bt[0] = cfg.etatzai * teff1;
  R.bt[1] = cfg.etatzai * teff2;
  const bool dp = cfg.doing_tail[2] != 0;
  R.lambda[0] = lambda_dave(1, dp, v.Ein, v.eE, v.pE, v.pP, v.etheta);
  R.lambda[1] = lambda_dave(2, dp, v.Ein, v.eE, v.pE, v.pP, v.etheta);
  R.lambda[2] = lambda_dave(3, dp, v.Ein, v.eE, v.pE, v.pP, v.etheta);
  R.rad_proton_this_ev = R.lambda[2] > 0;
  const double Ecutoff = 450.;
  double dsoft, dhard, dsoft_prime = 0.0;       // (static local of the reference: schwinger leaves it at zero)
  if (cfg.intcor_mode == 0)
    schwinger(cfg, Ecutoff, v.Ein, v.eE, v.etheta, 2 * v.Ein * v.eE * (1. - v.uez), v.Ein - v.eE, dsoft, dhard);
  else if (!cfg.use_offshell_rad)
    brem_onshell<kBremPrime>(v.Ein, v.eE, Ecutoff, R.rad_proton_this_ev, dsoft, dhard, dsoft_prime);
  else
    bremos<kBremPrime>(Ecutoff, v.Ein, v.eP * v.uex, v.eP * v.uey, v.eP * v.uez, v.pP * v.upx, v.pP * v.upy, v.pP * v.upz, v.pE,
           R.rad_proton_this_ev, dsoft, dhard, dsoft_prime);
  R.hardcorfac = 1. / (1. - dhard);
  R.g[4] = -dsoft_prime * Ecutoff + R.bt[0] + R.bt[1];
  // basicrad_init_ev(e1=Ein, e2=e.E, e3=p.E)
  const double e1 = v.Ein, e2 = v.eE, e3 = v.pE;
  R.g[1] = R.lambda[0] + R.bt[0];
  R.g[2] = R.lambda[1] + R.bt[1];
  R.g[3] = R.lambda[2];
  R.g[0] = R.g[1] + R.g[2] + R.g[3];
  double c_ext1 = R.bt[0] / m::pow(e1, R.bt[0]) / gamma_fn(1. + R.bt[0]);
  double c_ext2 = R.bt[1] / m::pow(e2, R.bt[1]) / gamma_fn(1. + R.bt[1]);
  R.g_ext = R.bt[0] + R.bt[1];
  double c_ext0 = c_ext1 * c_ext2 * R.g_ext / R.bt[0] / R.bt[1];
  c_ext0 = c_ext0 * gamma_fn(1. + R.bt[0]) * gamma_fn(1. + R.bt[1]) / gamma_fn(1. + R.g_ext);
  R.c_ext0 = c_ext0;
  double c4 = R.g[4] / m::pow(e1 * e2, R.g[4]) / gamma_fn(1. + R.g[4]);
  if (R.g[3] > 0) c4 = c4 / m::pow(e3, R.g[4]);
  R.c4 = c4;
  (void)Mp; (void)euler;
  R.frac[0] = R.g[1] / R.g[0];
  R.frac[1] = R.g[2] / R.g[0];
  R.frac[2] = R.g[3] / R.g[0];
}

// basicrad_init_ev's constants of the (Egamma1, Egamma2, Egamma3) basis, init.f:755-806: c_int(i), c(i), and the
// combined c_int(0), c(0), g_int.  Only rad_flag >= 2 reads c(1..3); they are rebuilt on demand from lambda, bt, g
// and the energies radc_init_ev was called with (the event record does not carry them).
struct BasisConst { double c[4], c_int[4], c_ext[4], c_int0, g_int; };
SIMC_HD_CALL BasisConst basis_constants(const RadEvDev& R, double e1, double e2, double e3) {
  const double Mp = 938.27231, euler = 0.577215665, one = 1.;
  const double e[4] = {0, e1, e2, e3};
  const double* lambda = R.lambda - 1;
  const double* bt = R.bt - 1;
  const double* g = R.g;
  double c_int[4], c_ext[4];
  c_int[1] = lambda[1] / m::pow(e[1] * e[2], lambda[1] / 2.);
  c_int[2] = lambda[2] / m::pow(e[1] * e[2], lambda[2] / 2.);
  c_int[3] = lambda[3] / m::pow(Mp * e[3], lambda[3] / 2.);
  for (int i = 1; i <= 3; ++i) c_int[i] = c_int[i] * m::exp(-euler * lambda[i]) / gamma_fn(one + lambda[i]);
  BasisConst B;
  B.g_int = lambda[1] + lambda[2] + lambda[3];
  c_int[0] = c_int[1] * c_int[2] * B.g_int / lambda[1] / lambda[2];
  if (lambda[3] > 0) c_int[0] = c_int[0] * c_int[3] / lambda[3];
  c_int[0] = c_int[0] * gamma_fn(one + lambda[1]) * gamma_fn(one + lambda[2]) * gamma_fn(one + lambda[3]) / gamma_fn(one + B.g_int);
  B.c_int0 = c_int[0];
  for (int i = 1; i <= 2; ++i) c_ext[i] = bt[i] / m::pow(e[i], bt[i]) / gamma_fn(one + bt[i]);
  c_ext[3] = 0.0;
  c_ext[0] = c_ext[1] * c_ext[2] * (bt[1] + bt[2]) / bt[1] / bt[2];
  c_ext[0] = c_ext[0] * gamma_fn(one + bt[1]) * gamma_fn(one + bt[2]) / gamma_fn(one + (bt[1] + bt[2]));
  for (int i = 0; i < 4; ++i) { B.c_int[i] = c_int[i]; B.c_ext[i] = c_ext[i]; }
  for (int i = 1; i <= 2; ++i)
    B.c[i] = c_int[i] * c_ext[i] * g[i] / lambda[i] / bt[i] * gamma_fn(one + lambda[i]) * gamma_fn(one + bt[i]) / gamma_fn(one + g[i]);
  B.c[3] = c_int[3];
  B.c[0] = B.c[1] * B.c[2] * g[0] / g[1] / g[2];
  if (g[3] > 0) B.c[0] = B.c[0] * B.c[3] / g[3];
  B.c[0] = B.c[0] * gamma_fn(one + g[1]) * gamma_fn(one + g[2]) * gamma_fn(one + g[3]) / gamma_fn(one + g[0]);
  return B;
}

// basicrad for one tail of that basis (itail = 1..3), radc.f:3-88
template <class RNG>
SIMC_HD_CALL void basicrad_tail(double g, double c, RNG& rng, double Egamma_lo, double Egamma_hi, double& Egamma, double& weight) {
  Egamma = 0.0; weight = 0.0;
  if (g <= 0) { weight = 1.0; return; }
  if (Egamma_hi <= Egamma_lo || Egamma_hi <= 0) return;
  const double power_hi = m::pow(Egamma_hi, g);
  double power_lo = 0.0;
  if (Egamma_lo > 0) power_lo = m::pow(Egamma_lo, g);
  const double ymin = power_lo / power_hi;
  const double y = ymin + rng.uniform() * (1. - ymin);
  const double x = m::pow(y, 1. / g);
  Egamma = x * Egamma_hi;
  weight = c / g * (power_hi - power_lo);
}

// extrad_phi, radc.f:668-707.  itail = 0 with extrad_flag = 3 is a `stop` in the reference: create() refuses the
// only setting that reaches it (rad_flag <= 1 never calls it with extrad_flag = 3, rad_flag >= 2 never with itail 0).
SIMC_HD_CALL double extrad_phi(const simc_run_config& cfg, const RadEvDev& R, int itail, double E1, double E2, double Egamma) {
  const double E[3] = {0, E1, E2};
  double phi = 1.0;
  if (cfg.extrad_flag == 2) {
    if (itail == 0) phi = 1. - (R.bt[0] / E[1] + R.bt[1] / E[2]) / (R.g[1] + R.g[2]) * Egamma;
    else if (itail == 1 || itail == 2) phi = 1. - R.bt[itail - 1] / E[itail] / R.g[itail] * Egamma;
  } else if (cfg.extrad_flag == 3) {
    if (itail == 1 || itail == 2) {
      const double x = Egamma / E[itail];
      const double t = R.bt[itail - 1] / cfg.etatzai;
      phi = phi * (1. - x + (x * x) / cfg.etatzai) * m::exp(t * ((cfg.etatzai - 0.5) - cfg.etatzai * x + (x * x) / 2.)) *
            gamma_fn(1. + R.bt[itail - 1]);
    }
  }
  return phi;
}

// basicrad with itail=0 -> 4 (peaked basis), radc.f:3-88.  u is the uniform it draws; it is
// only consumed when the function gets past its early returns (`drew`).
template <class RNG>
SIMC_HD_CALL void basicrad4(const RadEvDev& R, RNG& rng, double Egamma_lo, double Egamma_hi, double& Egamma,
                       double& weight) {
  Egamma = 0.0; weight = 0.0;
  const double g = R.g[4];
  if (g <= 0) { weight = 1.0; return; }
  if (Egamma_hi <= Egamma_lo || Egamma_hi <= 0) return;
  const double power_hi = m::pow(Egamma_hi, g);
  double power_lo = 0.0;
  if (Egamma_lo > 0) power_lo = m::pow(Egamma_lo, g);
  const double ymin = power_lo / power_hi;
  const double y = ymin + rng.uniform() * (1. - ymin);
  const double x = m::pow(y, 1. / g);
  Egamma = x * Egamma_hi;
  weight = R.c4 / g * (power_hi - power_lo);
}

// peaked_rad_weight, radc.f:523-646.  What the external branches compute besides phi_ext (dsoft_ext*, the Friedrich
// terms of extrad_flag = 3) is never read by the reference and is left out.
SIMC_HD_CALL double peaked_rad_weight(const simc_run_config& cfg, const RadEvDev& R, const VertexKin& v, double Egamma,
                                 double emin, double emax, double basicrad_weight) {
  const double eul = 0.577215665;
  if (cfg.extrad_flag <= 2 && cfg.rad_flag == 1) return basicrad_weight * extrad_phi(cfg, R, 0, v.Ein, v.eE, Egamma);
  double dsoft_intmin = 0.0, dsoft_intmax = 0.0, dhard, dprime;      // static locals of the reference: zero until assigned
  if (cfg.rad_flag == 0) {
    dsoft_intmin = 1.0;
    if (!cfg.use_offshell_rad) {
      if (emin > 0) brem_onshell<kBremSoft>(v.Ein, v.eE, emin, R.rad_proton_this_ev, dsoft_intmin, dhard, dprime);
      brem_onshell<kBremSoft>(v.Ein, v.eE, emax, R.rad_proton_this_ev, dsoft_intmax, dhard, dprime);
    } else {
      if (emin > 0)
        bremos<kBremSoft>(emin, v.Ein, v.eE * v.uex, v.eE * v.uey, v.eE * v.uez, v.pP * v.upx, v.pP * v.upy, v.pP * v.upz, v.pE,
               R.rad_proton_this_ev, dsoft_intmin, dhard, dprime);
      bremos<kBremSoft>(emax, v.Ein, v.eE * v.uex, v.eE * v.uey, v.eE * v.uez, v.pP * v.upx, v.pP * v.upy, v.pP * v.upz, v.pE,
             R.rad_proton_this_ev, dsoft_intmax, dhard, dprime);
    }
  }
  double w;
  if (emin > 0)
    w = R.c_ext0 / R.g_ext * (m::exp(-dsoft_intmax) * m::pow(emax, R.g_ext) - m::exp(-dsoft_intmin) * m::pow(emin, R.g_ext));
  else
    w = R.c_ext0 / R.g_ext * (m::exp(-dsoft_intmax) * m::pow(emax, R.g_ext));
  w = w * m::exp(-eul * R.g[4]) / gamma_fn(1. + R.g[4]) * gamma_fn(1. + R.g[4] - R.bt[0] - R.bt[1]) *
      gamma_fn(1. + R.bt[0]) * gamma_fn(1. + R.bt[1]) / gamma_fn(1. + R.g[4]);
  if (w < 0) w = 0;
  return w;
}

}  // namespace simc
