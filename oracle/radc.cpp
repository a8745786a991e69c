// ORACLE -- TEST INFRASTRUCTURE ONLY (see event.hpp).
// brem.f (bremos, inter, inter_prime, spence), radc.f (basicrad, gamma, generate_rad,
// peaked_rad_weight, extrad_phi, lambda_dave), init.f:655-813 (radc_init_ev, basicrad_init_ev).
// Every option branch is restated: rad_flag 0..3, extrad_flag 1..3, intcor_mode 0/1 (schwinger),
// use_offshell_rad 0/1 (brem / bremos).  The reference is built with -fno-automatic (Makefile:63): locals are
// static and start at zero, so a local that an option setting never assigns (dsoft_prime under intcor_mode=0,
// dsoft_intmin/max under rad_flag=1 with extrad_flag=3) reads as 0.0 -- restated as zero-initialised locals.
#include <cmath>
#include "event.hpp"

namespace simc_oracle {

using std::fabs;
using std::log;
using std::sqrt;

// brem.f:286-311
static double spence(double ax) {
  const double bx = fabs(ax);
  if (bx <= 1) return 0.0;
  return -0.5 * powi(log(bx), 2);
}

// brem.f:216-240
static double inter(bool calculate_spence, double alpha, double ar1, double ar2, double e1, double e2, double de) {
  const double pi = 3.141592653589793;
  const double de2 = e1 - e2;
  const double amult = -1. / (alpha * (ar1 - ar2));
  double v = log(fabs((e2 / de) + ar1 * (de2 / de))) * log(fabs((ar1 - 1.) / ar1)) -
             log(fabs((e2 / de) + ar2 * (de2 / de))) * log(fabs((ar2 - 1.) / ar2));
  if (calculate_spence) {
    const double arg1 = (de2 / (e2 + ar1 * de2)) * (ar1 - 1.);
    const double arg2 = (de2 / (e2 + ar1 * de2)) * (ar1);
    const double arg3 = (de2 / (e2 + ar2 * de2)) * (ar2 - 1.);
    const double arg4 = (de2 / (e2 + ar2 * de2)) * (ar2);
    v = v - spence(arg1) + spence(arg2) + spence(arg3) - spence(arg4);
  }
  return v * amult / (pi);
}

// brem.f:581-596
static double inter_prime(double alpha, double ar1, double ar2, double de) {
  const double pi = 3.141592653589793;
  const double amult = -1. / (alpha * (ar1 - ar2));
  return (-1. / de) * amult / pi * (log(fabs((ar1 - 1.) / ar1)) - log(fabs((ar2 - 1.) / ar2)));
}

// brem.f:6-214: the on-shell calculation (elastic e-p kinematics from ein, eout alone).
// include_hard = calculate_spence = .true. (init.f:646-647); produce_output off.
double brem(double ein, double eout, double egamma, bool radiate_proton, bool exponentiate, double& bsoft,
            double& bhard, double& dbsoft) {
  const double pi = 3.141592653589793, am = .93827231, ame = .00051099906, e2 = 1. / 137.0359895;
  const bool calculate_spence = true, include_hard = true;
  const double ak = ein / 1000., akp = eout / 1000., de = egamma / 1000.;
  const double eang = 2. * std::asin(std::pow(am / (2. * ak) * (ak / akp - 1.), 0.5));
  const double q2 = 4. * ak * akp * powi(std::sin(eang / 2.), 2);
  const double ape = am + ak - akp;
  const double ap = sqrt(powi(ape, 2) - powi(am, 2));
  const double pang = std::acos((ak - akp * std::cos(eang)) / ap);
  double aprod, adot, alpha, ar1, ar2;
  aprod = 1.e0;
  const double bei = aprod * (-1. / (2. * pi)) * log(ak / de);
  const double dbei = aprod * (1. / (2. * pi * de));
  aprod = 1.e0;
  const double bef = aprod * (-1. / (2. * pi)) * log(akp / de);
  const double dbef = aprod * (1. / (2. * pi * de));
  // e-e interference
  aprod = -1.e0;
  adot = ak * akp * (1. - std::cos(eang));
  alpha = 2. * powi(ame, 2) - 2. * adot;
  ar1 = 0.5 + sqrt(powi(adot, 2) - powi(ame, 4)) / alpha;
  ar2 = 0.5 - sqrt(powi(adot, 2) - powi(ame, 4)) / alpha;
  const double bee = aprod * adot * inter(calculate_spence, alpha, ar1, ar2, ak, akp, de);
  const double dbee = aprod * adot / (pi * alpha * (ar1 - ar2) * de) * (log((ar1 - 1) / ar1) - log((ar2 - 1) / ar2));
  double bpi = 0, bpf = 0, bpp = 0, bepii = 0, bepff = 0, bepif = 0, bepfi = 0;
  double dbpi = 0, dbpf = 0, dbpp = 0, dbepii = 0, dbepff = 0, dbepif = 0, dbepfi = 0;
  if (radiate_proton) {
    aprod = 1.e0;
    bpi = aprod * (-1. / (2. * pi)) * log(am / de);
    dbpi = aprod * (1. / (2. * pi * de));
    aprod = 1.e0;
    bpf = aprod * (-1. / (2. * pi)) * log(ape / de);
    dbpf = aprod * (1 / (2. * pi * de));
    // p-p
    aprod = -1.e0;
    adot = am * ape;
    alpha = 2. * powi(am, 2) - 2. * adot;
    ar1 = 0.5 + sqrt(powi(adot, 2) - powi(am, 4)) / alpha;
    ar2 = 0.5 - sqrt(powi(adot, 2) - powi(am, 4)) / alpha;
    bpp = aprod * adot * inter(calculate_spence, alpha, ar1, ar2, am, ape, de);
    dbpp = aprod * adot / (pi * alpha * (ar1 - ar2) * de) * (log((ar1 - 1) / ar1) - log((ar2 - 1) / ar2));
    // ei-pi
    aprod = -1.e0;
    adot = ak * am;
    alpha = powi(am, 2) + powi(ame, 2) - 2. * adot;
    ar1 = (powi(am, 2) - adot + sqrt(powi(adot, 2) - powi(ame * am, 2))) / alpha;
    ar2 = (powi(am, 2) - adot - sqrt(powi(adot, 2) - powi(ame * am, 2))) / alpha;
    bepii = aprod * adot * inter(calculate_spence, alpha, ar1, ar2, ak, am, de);
    dbepii = aprod * adot / (pi * alpha * (ar1 - ar2) * de) * (log((ar1 - 1) / ar1) - log((ar2 - 1) / ar2));
    // ef-pf
    aprod = -1.e0;
    adot = akp * ape - akp * ap * std::cos(eang + pang);
    alpha = powi(am, 2) + powi(ame, 2) - 2. * adot;
    ar1 = (powi(am, 2) - adot + sqrt(powi(adot, 2) - powi(ame * am, 2))) / alpha;
    ar2 = (powi(am, 2) - adot - sqrt(powi(adot, 2) - powi(ame * am, 2))) / alpha;
    bepff = aprod * adot * inter(calculate_spence, alpha, ar1, ar2, akp, ape, de);
    dbepff = aprod * adot / (pi * alpha * (ar1 - ar2) * de) * (log((ar1 - 1) / ar1) - log((ar2 - 1) / ar2));
    // ei-pf
    aprod = 1.e0;
    adot = ak * ape - ak * ap * std::cos(pang);
    alpha = powi(am, 2) + powi(ame, 2) - 2. * adot;
    ar1 = (powi(am, 2) - adot + sqrt(powi(adot, 2) - powi(ame * am, 2))) / alpha;
    ar2 = (powi(am, 2) - adot - sqrt(powi(adot, 2) - powi(ame * am, 2))) / alpha;
    bepif = aprod * adot * inter(calculate_spence, alpha, ar1, ar2, ak, ape, de);
    dbepif = aprod * adot / (pi * alpha * (ar1 - ar2) * de) * (log((ar1 - 1) / ar1) - log((ar2 - 1) / ar2));
    // ef-pi
    aprod = 1.e0;
    adot = akp * am;
    alpha = powi(am, 2) + powi(ame, 2) - 2. * adot;
    ar1 = (powi(am, 2) - adot + sqrt(powi(adot, 2) - powi(ame * am, 2))) / alpha;
    ar2 = (powi(am, 2) - adot - sqrt(powi(adot, 2) - powi(ame * am, 2))) / alpha;
    bepfi = aprod * adot * inter(calculate_spence, alpha, ar1, ar2, akp, am, de);
    dbepfi = aprod * adot / (pi * alpha * (ar1 - ar2) * de) * (log((ar1 - 1) / ar1) - log((ar2 - 1) / ar2));
  }
  const double b = 2. * e2 * (bei + bef + bee);
  double bz, bzz;
  if (radiate_proton) {
    bzz = 2. * e2 * (bpi + bpf + bpp);
    bz = 2. * e2 * (bepii + bepff + bepif + bepfi);
  } else {
    bzz = 0.0; bz = 0.0;
  }
  bsoft = b + bz + bzz;
  bhard = -1. * (e2 / pi) * (-28 / 9. + 13. / 6. * log(q2 / powi(ame, 2)));
  const double db = 2. * e2 * (dbei + dbef + dbee);
  double dbz, dbzz;
  if (radiate_proton) {
    dbzz = 2. * e2 * (dbpi + dbpf + dbpp);
    dbz = 2. * e2 * (dbepii + dbepff + dbepif + dbepfi);
  } else {
    dbzz = 0.0; dbz = 0.0;
  }
  dbsoft = db + dbz + dbzz;
  dbsoft = dbsoft / 1000.;
  double r;
  if (exponentiate) r = -dbsoft / std::exp(bsoft);
  else r = 1. - dbsoft;
  if (include_hard) r = r * (1. - bhard);
  return r;
}

// radc.f:746-764 (Abramowitz & Stegun 27.7.2 series)
double spen(double x) {
  double y = 1.0, s = 0.0;
  int i = 0;
  while (i <= 100 && fabs(y) > fabs(s) * 1.e-4) {
    i = i + 1;
    y = x * y;
    s = s + y / (double)(i * i);
  }
  return s;
}

// radc.f:711-742
double schwinger(const simc_run_config& cfg, double etta, double Ecutoff, const Event& vertex, bool include_hard,
                 double& dsoft, double& dhard) {
  const double lq = log(vertex.Q2 / powi(K::Me, 2)) - 1.0;
  const double s2 = powi(std::sin(vertex.e.theta / 2.), 2);
  const double b = 1. + 2. * vertex.nu * s2 / (cfg.targ.A * K::amu);
  const double spence_ = spen(1. - s2) - 2.5893784;
  dsoft = K::alpi * lq * log(vertex.Ein / powi(etta, 2) * vertex.e.E * b / powi(Ecutoff, 2));
  dhard = -K::alpi * (2.166666 * lq + spence_ - powi(log(vertex.Ein / vertex.e.E), 2) / 2.0);
  double r;
  if (cfg.use_expon == 0) r = std::exp(dsoft);
  else r = 1. + dsoft;
  if (include_hard) r = r / (1. - dhard);
  return r;
}

// radc.f:650-664
void extrad_friedrich(double etatzai, double Ei, double Ecutoff, double trad, double& dbrem, double& dbrem_prime) {
  const double x = Ecutoff / Ei;
  dbrem = trad * (-(etatzai - 0.5) - etatzai * log(x) + etatzai * x - 0.5 * powi(x, 2));
  dbrem_prime = -trad / Ei * (etatzai / x - etatzai + x);
}

// brem.f:344-577.  include_hard = calculate_spence = .true. (init.f:646-647)
double bremos(double egamma, double k_ix, double k_iy, double k_iz, double k_fx, double k_fy, double k_fz,
              double p_ix, double p_iy, double p_iz, double p_fx, double p_fy, double p_fz, double p_fe,
              bool radiate_proton, bool exponentiate, double& bsoft, double& bhard, double& dbsoft) {
  const double pi = 3.141592653589793, twopi = 2. * pi, ame = .00051099906, e2 = 1. / 137.0359895, mp = .93827231;
  const bool calculate_spence = true, include_hard = true;
  struct V4 { double e, x, y, z; } k_i, k_f, p_i, p_f;
  const double de = egamma / 1000.;
  k_i.x = k_ix / 1000.; k_i.y = k_iy / 1000.; k_i.z = k_iz / 1000.;
  k_f.x = k_fx / 1000.; k_f.y = k_fy / 1000.; k_f.z = k_fz / 1000.;
  p_i.x = p_ix / 1000.; p_i.y = p_iy / 1000.; p_i.z = p_iz / 1000.;
  p_f.e = p_fe / 1000.; p_f.x = p_fx / 1000.; p_f.y = p_fy / 1000.; p_f.z = p_fz / 1000.;
  k_i.e = std::pow(k_i.x * k_i.x + k_i.y * k_i.y + k_i.z * k_i.z + ame * ame, 0.5);
  k_f.e = std::pow(k_f.x * k_f.x + k_f.y * k_f.y + k_f.z * k_f.z + ame * ame, 0.5);
  p_i.e = mp;
  const double q2 = -1. * ((k_f.e - k_i.e) * (k_f.e - k_i.e) - (k_f.x - k_i.x) * (k_f.x - k_i.x) -
                           (k_f.y - k_i.y) * (k_f.y - k_i.y) - (k_f.z - k_i.z) * (k_f.z - k_i.z));
  const double ami = mp;
  const double amf = std::pow((p_f.e) * (p_f.e) - (p_f.x) * (p_f.x) - (p_f.y) * (p_f.y) - (p_f.z) * (p_f.z), 0.5);
  double aprod, adot, alpha, ar1, ar2;
  // electron terms
  aprod = 1.e0;
  const double bei = aprod * (-1. / twopi) * log(k_i.e / de);
  const double dbei = aprod * (-1. / twopi) * (-1. / de);
  aprod = 1.e0;
  const double bef = aprod * (-1. / twopi) * log(k_f.e / de);
  const double dbef = aprod * (-1. / twopi) * (-1. / de);
  aprod = -1.e0;
  adot = k_i.e * k_f.e - k_i.x * k_f.x - k_i.y * k_f.y - k_i.z * k_f.z;
  alpha = 2. * ame * ame - 2. * adot;
  ar1 = 0.5 + sqrt(4. * adot * adot - 4. * powi(ame, 4)) / (2. * alpha);
  ar2 = 0.5 - sqrt(4. * adot * adot - 4. * powi(ame, 4)) / (2. * alpha);
  const double bee = aprod * adot * inter(calculate_spence, alpha, ar1, ar2, k_i.e, k_f.e, de);
  const double dbee = aprod * adot * inter_prime(alpha, ar1, ar2, de);
  double bpi = 0, bpf = 0, bpp = 0, bepii = 0, bepff = 0, bepif = 0, bepfi = 0;
  double dbpi = 0, dbpf = 0, dbpp = 0, dbepii = 0, dbepff = 0, dbepif = 0, dbepfi = 0;
  if (radiate_proton) {
    aprod = 1.e0;
    bpi = aprod * (-1. / twopi) * log(p_i.e / de);
    dbpi = aprod * (-1. / twopi) * (-1. / de);
    aprod = 1.e0;
    bpf = aprod * (-1. / twopi) * log(p_f.e / de);
    dbpf = aprod * (-1. / twopi) * (-1. / de);
    // p-p interference
    aprod = -1.e0;
    adot = p_i.e * p_f.e - p_i.x * p_f.x - p_i.y * p_f.y - p_i.z * p_f.z;
    alpha = ami * ami + amf * amf - 2. * adot;
    ar1 = (2. * amf * amf - 2. * adot + sqrt(4. * adot * adot - 4. * powi(ami * amf, 2))) / (2. * alpha);
    ar2 = (2. * amf * amf - 2. * adot - sqrt(4. * adot * adot - 4. * powi(ami * amf, 2))) / (2. * alpha);
    bpp = aprod * adot * inter(calculate_spence, alpha, ar1, ar2, p_i.e, p_f.e, de);
    dbpp = aprod * adot * inter_prime(alpha, ar1, ar2, de);
    // ei-pi
    aprod = -1.e0;
    adot = k_i.e * p_i.e - k_i.x * p_i.x - k_i.y * p_i.y - k_i.z * p_i.z;
    alpha = ami * ami + ame * ame - 2. * adot;
    ar1 = (2. * ami * ami - 2. * adot + sqrt(4. * adot * adot - 4. * powi(ame * ami, 2))) / (2. * alpha);
    ar2 = (2. * ami * ami - 2. * adot - sqrt(4. * adot * adot - 4. * powi(ame * ami, 2))) / (2. * alpha);
    bepii = aprod * adot * inter(calculate_spence, alpha, ar1, ar2, k_i.e, p_i.e, de);
    dbepii = aprod * adot * inter_prime(alpha, ar1, ar2, de);
    // ef-pf
    aprod = -1.e0;
    adot = k_f.e * p_f.e - k_f.x * p_f.x - k_f.y * p_f.y - k_f.z * p_f.z;
    alpha = amf * amf + ame * ame - 2. * adot;
    ar1 = (2. * amf * amf - 2. * adot + sqrt(4. * adot * adot - 4. * powi(ame * amf, 2))) / (2. * alpha);
    ar2 = (2. * amf * amf - 2. * adot - sqrt(4. * adot * adot - 4. * powi(ame * amf, 2))) / (2. * alpha);
    bepff = aprod * adot * inter(calculate_spence, alpha, ar1, ar2, k_f.e, p_f.e, de);
    dbepff = aprod * adot * inter_prime(alpha, ar1, ar2, de);
    // ei-pf
    aprod = 1.e0;
    adot = k_i.e * p_f.e - k_i.x * p_f.x - k_i.y * p_f.y - k_i.z * p_f.z;
    alpha = amf * amf + ame * ame - 2. * adot;
    ar1 = (2. * amf * amf - 2. * adot + sqrt(4. * adot * adot - 4. * powi(ame * amf, 2))) / (2. * alpha);
    ar2 = (2. * amf * amf - 2. * adot - sqrt(4. * adot * adot - 4. * powi(ame * amf, 2))) / (2. * alpha);
    bepif = aprod * adot * inter(calculate_spence, alpha, ar1, ar2, k_i.e, p_f.e, de);
    dbepif = aprod * adot * inter_prime(alpha, ar1, ar2, de);
    // ef-pi
    aprod = 1.e0;
    adot = k_f.e * p_i.e - k_f.x * p_i.x - k_f.y * p_i.y - k_f.z * p_i.z;
    alpha = ami * ami + ame * ame - 2. * adot;
    ar1 = (2. * ami * ami - 2. * adot + sqrt(4. * adot * adot - 4. * powi(ame * ami, 2))) / (2. * alpha);
    ar2 = (2. * ami * ami - 2. * adot - sqrt(4. * adot * adot - 4. * powi(ame * ami, 2))) / (2. * alpha);
    bepfi = aprod * adot * inter(calculate_spence, alpha, ar1, ar2, k_f.e, p_i.e, de);
    dbepfi = aprod * adot * inter_prime(alpha, ar1, ar2, de);
  }
  const double b = 2. * e2 * (bei + bef + bee);
  double bz, bzz;
  if (radiate_proton) {
    bzz = 2. * e2 * (bpi + bpf + bpp);
    bz = 2. * e2 * (bepii + bepff + bepif + bepfi);
  } else {
    bzz = 0.0; bz = 0.0;
  }
  bsoft = b + bz + bzz;
  bhard = -1. * (e2 / pi) * (-28 / 9. + 13. / 6. * log(q2 / (ame * ame)));
  const double db = 2. * e2 * (dbei + dbef + dbee);
  double dbz, dbzz;
  if (radiate_proton) {
    dbzz = 2. * e2 * (dbpi + dbpf + dbpp);
    dbz = 2. * e2 * (dbepii + dbepff + dbepif + dbepfi);
  } else {
    dbzz = 0.0; dbz = 0.0;
  }
  dbsoft = db + dbz + dbzz;
  dbsoft = dbsoft / 1000.;
  double r;
  if (exponentiate) r = -dbsoft / std::exp(bsoft);
  else r = 1. - dbsoft;
  if (include_hard) r = r * (1. - bhard);
  return r;
}

// radc.f:92-116
double gamma_fn(double x) {
  double g = 1.0;
  const int n = (int)std::lround((x - 1) - 0.5);     // Fortran nint: half away from zero
  const double y = x - 1 - n;
  if (n != 0) {
    const int sgn = n > 0 ? 1 : -1;
    for (int i = sgn; sgn > 0 ? i <= n : i >= n; i += sgn) g = g * powi(y + 1 + i, sgn);
  }
  g = g * (1. - 0.5748646 * y + 0.9512363 * powi(y, 2) - 0.6998588 * powi(y, 3) + 0.4245549 * powi(y, 4) -
           0.1010678 * powi(y, 5));
  return g;
}

// radc.f:768-821
static double lambda_dave(int itail, int plus_flag, bool doing_proton, double e1, double e2, double e3, double p3,
                          double th) {
  double plus_term = 0.0;
  if (plus_flag == 1 && itail < 3) {
    plus_term = log((1. - std::cos(th)) / 2.);
    if (doing_proton) plus_term = plus_term + 2. * log(e1 / e2);
  }
  if (itail == 1) return K::alpi * (2. * log(2. * e1 / K::Me) - 1. + plus_term);
  if (itail == 2) return K::alpi * (2. * log(2. * e2 / K::Me) - 1. + plus_term);
  if (itail == 3) {
    if (doing_proton) {
      double v = K::alpi * ((e3 / p3) * log((e3 + p3) / (e3 - p3)) - 2.);
      if (v < 0) v = 0.0;
      return v;
    }
    return 0.0;
  }
  return 0.0;
}

// init.f:732-813
static void basicrad_init_ev(Sim& s, double e1, double e2, double e3) {
  RadEv& R = s.rad;
  const double one = 1.;
  const double e[4] = {0, e1, e2, e3};
  double* g = R.g; double* c = R.c; double* c_int = R.c_int; double* c_ext = R.c_ext;
  const double* lambda = R.lambda - 1;   // 1-based
  const double* bt = R.bt - 1;
  g[1] = lambda[1] + bt[1];
  g[2] = lambda[2] + bt[2];
  g[3] = lambda[3];
  g[0] = g[1] + g[2] + g[3];
  c_int[1] = lambda[1] / std::pow(e[1] * e[2], lambda[1] / 2.);
  c_int[2] = lambda[2] / std::pow(e[1] * e[2], lambda[2] / 2.);
  c_int[3] = lambda[3] / std::pow(K::Mp * e[3], lambda[3] / 2.);
  for (int i = 1; i <= 3; ++i) c_int[i] = c_int[i] * std::exp(-K::euler * lambda[i]) / gamma_fn(one + lambda[i]);
  R.g_int = lambda[1] + lambda[2] + lambda[3];
  c_int[0] = c_int[1] * c_int[2] * R.g_int / lambda[1] / lambda[2];
  if (lambda[3] > 0) c_int[0] = c_int[0] * c_int[3] / lambda[3];
  c_int[0] = c_int[0] * gamma_fn(one + lambda[1]) * gamma_fn(one + lambda[2]) * gamma_fn(one + lambda[3]) /
             gamma_fn(one + R.g_int);
  for (int i = 1; i <= 2; ++i) c_ext[i] = bt[i] / std::pow(e[i], bt[i]) / gamma_fn(one + bt[i]);
  c_ext[3] = 0.0;
  R.g_ext = bt[1] + bt[2];
  c_ext[0] = c_ext[1] * c_ext[2] * R.g_ext / bt[1] / bt[2];
  c_ext[0] = c_ext[0] * gamma_fn(one + bt[1]) * gamma_fn(one + bt[2]) / gamma_fn(one + R.g_ext);
  for (int i = 1; i <= 2; ++i)
    c[i] = c_int[i] * c_ext[i] * g[i] / lambda[i] / bt[i] * gamma_fn(one + lambda[i]) * gamma_fn(one + bt[i]) /
           gamma_fn(one + g[i]);
  c[3] = c_int[3];
  c[0] = c[1] * c[2] * g[0] / g[1] / g[2];
  if (g[3] > 0) c[0] = c[0] * c[3] / g[3];
  c[0] = c[0] * gamma_fn(one + g[1]) * gamma_fn(one + g[2]) * gamma_fn(one + g[3]) / gamma_fn(one + g[0]);
  c[4] = g[4] / std::pow(e1 * e2, g[4]) / gamma_fn(one + g[4]);
  if (g[3] > 0) c[4] = c[4] / std::pow(e3, g[4]);
}

// init.f:655-728
void radc_init_ev(Sim& s, EventMain& main, Event& vertex) {
  const simc_run_config& cfg = *s.cfg;
  RadEv& R = s.rad;
  R.etta = 1.0;
  for (int i = 0; i < 2; ++i) R.bt[i] = cfg.etatzai * main.target.teff[i];
  for (int i = 1; i <= 3; ++i)
    R.lambda[i - 1] = lambda_dave(i, 1, cfg.doing_tail[2] != 0, vertex.Ein, vertex.e.E, vertex.p.E, vertex.p.P,
                                  vertex.e.theta);
  R.rad_proton_this_ev = R.lambda[2] > 0;
  const double Ecutoff = 450.;
  // dsoft_prime is a static local (-fno-automatic) that schwinger never assigns: it reads 0.0 under intcor_mode=0
  double dsoft = 0, dhard = 0, dsoft_prime = 0;
  if (cfg.intcor_mode == 0)
    schwinger(cfg, R.etta, Ecutoff, vertex, true, dsoft, dhard);
  else if (!cfg.use_offshell_rad)
    brem(vertex.Ein, vertex.e.E, Ecutoff, R.rad_proton_this_ev, cfg.use_expon == 1, dsoft, dhard, dsoft_prime);
  else
    bremos(Ecutoff, 0., 0., vertex.Ein, vertex.e.P * vertex.ue.x, vertex.e.P * vertex.ue.y, vertex.e.P * vertex.ue.z,
           0., 0., 0., vertex.p.P * vertex.up.x, vertex.p.P * vertex.up.y, vertex.p.P * vertex.up.z, vertex.p.E,
           R.rad_proton_this_ev, cfg.use_expon == 1, dsoft, dhard, dsoft_prime);
  R.hardcorfac = 1. / (1. - dhard);
  R.g[4] = -dsoft_prime * Ecutoff + R.bt[0] + R.bt[1];
  basicrad_init_ev(s, vertex.Ein, vertex.e.E, vertex.p.E);
  for (int i = 1; i <= 3; ++i) R.frac[i - 1] = R.g[i] / R.g[0];
}

// radc.f:3-88
static void basicrad(Sim& s, int itail, double Egamma_lo, double Egamma_hi, double& Egamma, double& weight,
                     double& val_reciprocal) {
  RadEv& R = s.rad;
  Egamma = 0.0; weight = 0.0; val_reciprocal = 0.0;
  if (itail == 0) itail = 4;
  if (R.g[itail] <= 0) { weight = 1.0; return; }
  if (Egamma_hi <= Egamma_lo || Egamma_hi <= 0) return;
  const double power_hi = std::pow(Egamma_hi, R.g[itail]);
  double power_lo = 0.0;
  if (Egamma_lo > 0) power_lo = std::pow(Egamma_lo, R.g[itail]);
  const double ymin = power_lo / power_hi;
  const double y = ymin + s.rng->grnd() * (1. - ymin);
  const double x = std::pow(y, 1. / R.g[itail]);
  Egamma = x * Egamma_hi;
  if (Egamma > 0) val_reciprocal = std::pow(Egamma, 1. - R.g[itail]) * (power_hi - power_lo) / R.g[itail];
  weight = R.c[itail] / R.g[itail] * (power_hi - power_lo);
}

// radc.f:668-707
static double extrad_phi(Sim& s, int itail, double E1, double E2, double Egamma) {
  const RadEv& R = s.rad;
  const double E[3] = {0, E1, E2};
  double phi = 1.0;
  if (s.cfg->extrad_flag == 2) {
    if (itail == 0) phi = 1. - (R.bt[0] / E[1] + R.bt[1] / E[2]) / (R.g[1] + R.g[2]) * Egamma;
    else if (itail == 1 || itail == 2) phi = 1. - R.bt[itail - 1] / E[itail] / R.g[itail] * Egamma;
  } else if (s.cfg->extrad_flag == 3) {
    if (itail == 0) throw std::runtime_error("Idiot! a multiplicative factor EXTRAD_PHI is not defined for peaking approx and EXTRAD_FLAG>2!");
    if (itail == 1 || itail == 2) {
      const double etatzai = s.cfg->etatzai;
      const double x = Egamma / E[itail];
      const double t = R.bt[itail - 1] / etatzai;
      phi = phi * (1. - x + powi(x, 2) / etatzai) * std::exp(t * ((etatzai - 0.5) - etatzai * x + powi(x, 2) / 2.)) *
            gamma_fn(1. + R.bt[itail - 1]);
    }
  }
  return phi;
}

// radc.f:523-646
static double peaked_rad_weight(Sim& s, const Event& vertex, double Egamma, double emin, double emax,
                                double basicrad_val_reciprocal, double basicrad_weight) {
  const RadEv& R = s.rad;
  const simc_run_config& cfg = *s.cfg;
  const double ein = vertex.Ein, eout = vertex.e.E, eul = 0.577215665;
  (void)basicrad_val_reciprocal;
  // static locals of the reference (zero until an executed branch assigns them)
  double dsoft_intmin = 0.0, dsoft_intmax = 0.0, dhard = 0.0, dprime = 0.0;
  double phi_ext = 1.0;
  if (cfg.extrad_flag <= 2) {
    phi_ext = extrad_phi(s, 0, ein, eout, Egamma);
    if (cfg.rad_flag == 1) return basicrad_weight * phi_ext;
    // dsoft_extmin/max and dsoft_ext_prime (radc.f:574-578) are computed and never read
  } else {
    // Friedrich prescription: computed and never read (radc.f:580-585)
    const double t1 = R.bt[0] / cfg.etatzai, t2 = R.bt[1] / cfg.etatzai;
    double d1, d1p, d2, d2p;
    extrad_friedrich(cfg.etatzai, ein, Egamma, t1, d1, d1p);
    extrad_friedrich(cfg.etatzai, eout, Egamma, t2, d2, d2p);
  }
  if (cfg.rad_flag == 0) {
    if (!cfg.use_offshell_rad) {
      if (emin > 0) brem(ein, eout, emin, R.rad_proton_this_ev, cfg.use_expon == 1, dsoft_intmin, dhard, dprime);
      else dsoft_intmin = 1.0;
      brem(ein, eout, emax, R.rad_proton_this_ev, cfg.use_expon == 1, dsoft_intmax, dhard, dprime);
    } else {
      if (emin > 0)
        bremos(emin, 0., 0., ein, vertex.e.E * vertex.ue.x, vertex.e.E * vertex.ue.y, vertex.e.E * vertex.ue.z, 0., 0.,
               0., vertex.p.P * vertex.up.x, vertex.p.P * vertex.up.y, vertex.p.P * vertex.up.z, vertex.p.E,
               R.rad_proton_this_ev, cfg.use_expon == 1, dsoft_intmin, dhard, dprime);
      else
        dsoft_intmin = 1.0;
      bremos(emax, 0., 0., ein, vertex.e.E * vertex.ue.x, vertex.e.E * vertex.ue.y, vertex.e.E * vertex.ue.z, 0., 0., 0.,
             vertex.p.P * vertex.up.x, vertex.p.P * vertex.up.y, vertex.p.P * vertex.up.z, vertex.p.E,
             R.rad_proton_this_ev, cfg.use_expon == 1, dsoft_intmax, dhard, dprime);
    }
  }
  // (rad_flag = 1 with extrad_flag = 3 gets here with dsoft_int* never assigned: radc.f:627-631 sets other locals)
  double w;
  if (emin > 0)
    w = R.c_ext[0] / R.g_ext *
        (std::exp(-dsoft_intmax) * std::pow(emax, R.g_ext) - std::exp(-dsoft_intmin) * std::pow(emin, R.g_ext));
  else
    w = R.c_ext[0] / R.g_ext * (std::exp(-dsoft_intmax) * std::pow(emax, R.g_ext));
  w = w * std::exp(-eul * R.g[4]) / gamma_fn(1. + R.g[4]) * gamma_fn(1. + R.g[4] - R.bt[0] - R.bt[1]) *
      gamma_fn(1. + R.bt[0]) * gamma_fn(1. + R.bt[1]) / gamma_fn(1. + R.g[4]);
  if (w < 0) w = 0;
  return w;
}

double peaked_rad_weight_public(Sim& s, const Event& vertex, double Egamma, double emin, double emax) {
  return peaked_rad_weight(s, vertex, Egamma, emin, emax, 0.0, 1.0);
}
double extrad_phi_public(Sim& s, int itail, double E1, double E2, double Egamma) { return extrad_phi(s, itail, E1, E2, Egamma); }

// radc.f:120-519
bool generate_rad(Sim& s, EventMain& main, Event& vertex, Event& orig) {
  const simc_run_config& cfg = *s.cfg;
  RadEv& R = s.rad;
  double rad_weight = 1;
  for (int i = 0; i < 3; ++i) R.Egamma_used[i] = 0.0;
  s.ntup.radphot = 0.; s.ntup.radarm = 0.;
  int peaked_basis_flag = 1;
  if (cfg.rad_flag <= 1) {
    peaked_basis_flag = 0;
    const double x = s.rng->grnd();
    if (x >= R.frac[0] + R.frac[1]) R.ntail = 3;
    else if (x >= R.frac[0]) R.ntail = 2;
    else R.ntail = 1;
  } else if (cfg.rad_flag == 2) {
    R.ntail = (int)(s.rng->grnd() * 3.) + 1;
    if (R.ntail == 4) R.ntail = 3;
  } else if (cfg.rad_flag == 3) {
    R.ntail = 0;
  } else {
    throw std::runtime_error("Idiot! rad_flag is set stupidly");
  }
  const int ntail = R.ntail;
  const double max_delta_Trec =
      std::max((vertex.Trec - cfg.VERTEXedge.Trec.min), (cfg.VERTEXedge.Trec.max - vertex.Trec));
  double basicrad_weight, basicrad_val_reciprocal;
  double* Egamma_min = R.Egamma_min - 1; double* Egamma_max = R.Egamma_max - 1; double* Egamma_used = R.Egamma_used - 1;

  // tail 1: incoming electron
  if (cfg.doing_tail[0] && (ntail == 0 || ntail == 1)) {
    if (cfg.doing_heavy) {
      Egamma_min[1] = vertex.Em - cfg.VERTEXedge.Em.max - max_delta_Trec;
      Egamma_max[1] = vertex.Em - cfg.VERTEXedge.Em.min + max_delta_Trec;
      if (ntail != 0) Egamma_max[1] = std::min(Egamma_max[1], vertex.Em - cfg.edge.Em.min + max_delta_Trec);
    } else if (cfg.doing_hyd_elast) {
      double ebeam_max = K::Mp * cfg.edge.e.E.max / (K::Mp - cfg.edge.e.E.max * (1. - vertex.ue.z));
      if (ebeam_max < 0) ebeam_max = 1.e10;
      const double ebeam_min = K::Mp * cfg.edge.e.E.min / (K::Mp - cfg.edge.e.E.min * (1. - vertex.ue.z));
      Egamma_min[1] = vertex.Ein - ebeam_max;
      Egamma_max[1] = vertex.Ein - ebeam_min;
      Egamma_max[1] = std::min(Egamma_max[1], cfg.edge.Em.max);
    } else if (cfg.doing_deuterium) {
      Egamma_max[1] = std::min(cfg.Egamma1_max, cfg.gen.sumEgen.max - vertex.e.E);
      if (ntail != 0) Egamma_min[1] = cfg.gen.sumEgen.min - vertex.e.E;
      // ntail = 0: Egamma_min(1) keeps the value the previous event left in COMMON /radccom/, which only ever
      // decreases from its initial zero (dE_edge_test is subtracted below): any value <= 0 gives the same
      // basicrad, so zero stands for it here
      else Egamma_min[1] = 0.0;
    } else if (cfg.doing_pion || cfg.doing_kaon || cfg.doing_rho || cfg.doing_semi) {
      Egamma_min[1] = 0.;
      Egamma_max[1] = cfg.gen.sumEgen.max - vertex.e.E;
    }
    Egamma_max[1] = std::min(Egamma_max[1], cfg.Egamma1_max);
    Egamma_min[1] = Egamma_min[1] - cfg.dE_edge_test;
    Egamma_max[1] = Egamma_max[1] + cfg.dE_edge_test;
    if (cfg.hardwired_rad) Egamma_max[1] = cfg.Egamma_gen_max;
    basicrad(s, 1 * peaked_basis_flag, Egamma_min[1], Egamma_max[1], Egamma_used[1], basicrad_weight,
             basicrad_val_reciprocal);
    if (basicrad_weight <= 0) return false;
    vertex.Ein = vertex.Ein - Egamma_used[1];
    if (!complete_ev(s, main, vertex)) return false;
    rad_weight = rad_weight * basicrad_weight;
    if (cfg.rad_flag <= 1)
      rad_weight = peaked_rad_weight(s, vertex, Egamma_used[1], Egamma_min[1], Egamma_max[1], basicrad_val_reciprocal,
                                     basicrad_weight);
    else
      rad_weight = rad_weight * extrad_phi(s, 1, vertex.Ein, vertex.e.E, Egamma_used[1]);
  }
  if (cfg.doing_heavy) {
    if (vertex.Em < cfg.VERTEXedge.Em.min || vertex.Em > cfg.VERTEXedge.Em.max || vertex.Pm < cfg.VERTEXedge.Pm.min ||
        vertex.Pm > cfg.VERTEXedge.Pm.max)
      return false;
  }
  // tail 2: scattered electron
  if (cfg.doing_tail[1] && (ntail == 0 || ntail == 2)) {
    Egamma_min[2] = vertex.e.E - cfg.edge.e.E.max;
    Egamma_max[2] = vertex.e.E - cfg.edge.e.E.min;
    if (cfg.doing_eep) {
      Egamma_max[2] = std::min(Egamma_max[2], (cfg.edge.Em.max - vertex.Em) - Egamma_used[1] + max_delta_Trec);
      if (ntail != 0 || !R.rad_proton_this_ev)
        Egamma_min[2] = std::max(Egamma_min[2], (cfg.edge.Em.min - vertex.Em) - Egamma_used[1] - max_delta_Trec);
    }
    Egamma_max[2] = std::min(Egamma_max[2], cfg.Egamma_tot_max - Egamma_used[1]);
    Egamma_min[2] = Egamma_min[2] - cfg.dE_edge_test;
    Egamma_max[2] = Egamma_max[2] + cfg.dE_edge_test;
    if (cfg.hardwired_rad) Egamma_max[2] = cfg.Egamma_gen_max;
    basicrad(s, 2 * peaked_basis_flag, Egamma_min[2], Egamma_max[2], Egamma_used[2], basicrad_weight,
             basicrad_val_reciprocal);
    if (basicrad_weight <= 0) return false;
    rad_weight = rad_weight * basicrad_weight;
    if (cfg.rad_flag <= 1)
      rad_weight = peaked_rad_weight(s, vertex, Egamma_used[2], Egamma_min[2], Egamma_max[2], basicrad_val_reciprocal,
                                     basicrad_weight);
    else
      rad_weight = rad_weight * extrad_phi(s, 2, vertex.Ein, vertex.e.E, Egamma_used[2]);
  }
  // tail 3: hadron
  if (R.rad_proton_this_ev && (ntail == 0 || ntail == 3)) {
    Egamma_min[3] = vertex.p.E - cfg.edge.p.E.max;
    Egamma_max[3] = vertex.p.E - cfg.edge.p.E.min;
    if (cfg.doing_eep) {
      Egamma_max[3] =
          std::min(Egamma_max[3], (cfg.edge.Em.max - vertex.Em) - Egamma_used[1] - Egamma_used[2] + max_delta_Trec);
      Egamma_min[3] =
          std::max(Egamma_min[3], (cfg.edge.Em.min - vertex.Em) - Egamma_used[1] - Egamma_used[2] - max_delta_Trec);
    }
    Egamma_max[3] = std::min(Egamma_max[3], cfg.Egamma_tot_max - Egamma_used[1] - Egamma_used[2]);
    Egamma_min[3] = Egamma_min[3] - cfg.dE_edge_test;
    Egamma_max[3] = Egamma_max[3] + cfg.dE_edge_test;
    if (cfg.hardwired_rad) Egamma_max[3] = cfg.Egamma_gen_max;
    basicrad(s, 3 * peaked_basis_flag, Egamma_min[3], Egamma_max[3], Egamma_used[3], basicrad_weight,
             basicrad_val_reciprocal);
    if (basicrad_weight <= 0) return false;
    rad_weight = rad_weight * basicrad_weight;
    if (cfg.rad_flag <= 1)
      rad_weight = peaked_rad_weight(s, vertex, Egamma_used[3], Egamma_min[3], Egamma_max[3], basicrad_val_reciprocal,
                                     basicrad_weight);
    else
      rad_weight = rad_weight * extrad_phi(s, 3, vertex.Ein, vertex.e.E, Egamma_used[3]);
  }
  // orig = vertex + radiation, radc.f:476-515
  orig = vertex;
  orig.Ein = vertex.Ein + Egamma_used[1];
  orig.e.E = vertex.e.E - Egamma_used[2];
  if (orig.e.E <= 0e0) return false;
  orig.e.P = orig.e.E;
  orig.e.delta = (orig.e.P - cfg.spec_e.P) / cfg.spec_e.P * 100.;
  orig.p.E = vertex.p.E - Egamma_used[3];
  if (orig.p.E <= s.Mh) return false;
  orig.p.P = sqrt(orig.p.E * orig.p.E - s.Mh2);
  orig.p.delta = (orig.p.P - cfg.spec_p.P) / cfg.spec_p.P * 100.;
  s.ntup.radphot = Egamma_used[1] + Egamma_used[2] + Egamma_used[3];
  s.ntup.radarm = ntail;
  main.gen_weight = main.gen_weight * rad_weight / R.hardcorfac;
  return true;
}

}  // namespace simc_oracle
