// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into libsimc_b200.so.
// PARITY UNPINNED (no golden vectors in the reference, no Fortran compiler here).
//
// CPU restatement of SIMC's numeric primitives and shared transport routines:
//   gauss1.f, cern/lfit.f, loren.f, shared/project.f, shared/musc.f,
//   shared/musc_ext.f, shared/rotate_haxis.f, shared/rotate_vaxis.f,
//   shared/transp.f (evaluation + transp_init parser).
// Plain scalar C++, glibc libm, compiled with -O2 -ffp-contract=off so that no FMA
// is formed (the reference is an x86-64 SSE2 build).
#pragma once
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>
#include "rng.hpp"

namespace simc_oracle {

// constants.inc:17-46 (literals are doubles under -fdefault-real-8)
namespace K {
constexpr double Me = 0.51099906, Me2 = Me * Me;
constexpr double Mp = 938.27231, Mp2 = Mp * Mp;
constexpr double Mn = 939.56563;
constexpr double Mpi = 139.57018, Mpi2 = Mpi * Mpi;
constexpr double Mmu = 105.6583755;
constexpr double Mpi0 = 134.9766;
constexpr double Mk = 493.677, Mk2 = Mk * Mk;
constexpr double Mrho = 769.3, Mrho2 = Mrho * Mrho;
constexpr double amu = 931.49432;
constexpr double hbarc = 197.327053;
constexpr double pi = 3.141592653589793;
constexpr double alpha = 1. / 137.0359895;
constexpr double alpi = alpha / pi;
constexpr double euler = 0.577215665;
}  // namespace K

// libgcc __powidf2: what gfortran emits for real**integer
static inline double powi(double x, int m) {
  unsigned n = m < 0 ? 0u - (unsigned)m : (unsigned)m;
  double y = (n & 1) ? x : 1.0;
  while (n >>= 1) {
    x = x * x;
    if (n & 1) y *= x;
  }
  return m < 0 ? 1.0 / y : y;
}

// COMMON /track/ (spectrometers.inc:47-60) plus the few globals the single-arm
// code touches: ctau (simulate.inc:153), Mh2_final, decdist (simulate.inc:183).
struct Track {
  double xs = 0, ys = 0, zs = 0, dxdzs = 0, dydzs = 0, dpps = 0;
  double ctau = 0, Mh2_final = 0, decdist = 0;
  Rng* rng = nullptr;
  long long* calls = nullptr;   // work counters (transp calls per class, recon in slot 47), not in the reference
};

// gauss1.f:1-30
static inline double gauss1(Rng& r, double nsigmax) {
  for (;;) {
    const double u1 = r.grnd();
    const double u2 = r.grnd();
    const double v1 = 2.0 * u1 - 1.0;
    const double v2 = 2.0 * u2 - 1.0;
    const double s = v1 * v1 + v2 * v2;
    if (s > 1. || s == 0) continue;
    const double g = v1 * std::sqrt(-2. * std::log(s) / s);
    if (std::fabs(g) > nsigmax) continue;
    return g;
  }
}

// cern/lfit.f:11-61, KEY=0.  X,Y,A,B are REAL*4; the sums are implicitly typed
// REAL, i.e. 8 bytes under -fdefault-real-8.
static inline void lfit(const float* x, const float* y, int l, float& a, float& b) {
  a = 0.f; b = 0.f;
  if (l < 2) return;
  double count = 0., sumx = 0., sumy = 0., sumxy = 0., sumxx = 0., sumyy = 0.;
  for (int j = 0; j < l; ++j) {
    if (y[j] == 0.f) continue;
    sumx = sumx + x[j];
    sumy = sumy + y[j];
    count = count + 1.0;
  }
  if (count <= 1.) return;
  const double ymed = sumy / count, xmed = sumx / count;
  for (int j = 0; j < l; ++j) {
    if (y[j] == 0.f) continue;
    const double scartx = x[j] - xmed, scarty = y[j] - ymed;
    sumxy = sumxy + scartx * scarty;
    sumxx = sumxx + scartx * scartx;
    sumyy = sumyy + scarty * scarty;
  }
  if (sumxx == 0.) return;
  a = (float)(sumxy / sumxx);
  b = (float)(ymed - (double)a * xmed);
}

// loren.f:1-26
static inline void loren(double gam, double bx, double by, double bz, double e, double x, double y, double z,
                         double& ef1, double& pxf, double& pyf, double& pzf, double& pf1) {
  const double gam1 = gam * gam / (1. + gam);
  ef1 = gam * (e - bx * x - by * y - bz * z);
  pxf = (1 + gam1 * bx * bx) * x + gam1 * bx * (by * y + bz * z) - gam * bx * e;
  pyf = (1 + gam1 * by * by) * y + gam1 * by * (bx * x + bz * z) - gam * by * e;
  pzf = (1 + gam1 * bz * bz) * z + gam1 * bz * (by * y + bx * x) - gam * bz * e;
  pf1 = std::sqrt(pxf * pxf + pyf * pyf + pzf * pzf);
}

// Decay kinematics shared by project.f:67-110 and transp.f:147-186,236-275.
// `kaon_pipi_mfinal` is the mass given to the K->pi pi branch: Mpi everywhere
// except the first-half branch of transp (Mk, transp.f:158).
static inline void decay_in_flight(Track& t, double& m2, double& ph, double p_spec, double beta, double gamma,
                                   double kaon_pipi_mfinal) {
  Rng& r = *t.rng;
  const double rph = r.grnd() * 2. * K::pi;
  const double rth1 = r.grnd() * 2. - 1.;
  const double rth = std::acos(rth1);
  double pr = 0.;
  double m_final = K::Mmu;
  if (std::fabs(std::sqrt(m2) - K::Mpi) < 2) pr = 29.783;
  if (std::fabs(std::sqrt(m2) - K::Mk) < 2) {
    if (r.grnd() < 0.7) {
      pr = 235.5;
    } else {
      pr = std::sqrt(K::Mk * K::Mk / 4. - K::Mpi * K::Mpi);
      m_final = kaon_pipi_mfinal;
    }
  }
  if (pr == 0.) throw std::runtime_error("error, cannot decay particle");
  const double er = std::sqrt(m_final * m_final + pr * pr);
  const double pxr = pr * std::sin(rth) * std::cos(rph);
  const double pyr = pr * std::sin(rth) * std::sin(rph);
  const double pzr = pr * std::cos(rth);
  const double nrm = std::sqrt(1. + t.dxdzs * t.dxdzs + t.dydzs * t.dydzs);
  const double bx = -beta * t.dxdzs / nrm;
  const double by = -beta * t.dydzs / nrm;
  const double bz = -beta * 1. / nrm;
  double ef, pxf, pyf, pzf, pf;
  loren(gamma, bx, by, bz, er, pxr, pyr, pzr, ef, pxf, pyf, pzf, pf);
  t.dxdzs = pxf / pzf;
  t.dydzs = pyf / pzf;
  t.dpps = 100. * (pf / p_spec - 1.);
  ph = pf;
  m2 = m_final * m_final;
  t.Mh2_final = m2;
}

// shared/project.f:1-122 (x_new,y_new are always the COMMON xs,ys in every caller)
static inline void project(Track& t, double z_drift, bool decay_flag, bool& dflag, double& m2, double& ph,
                           double& pathlen) {
  if (!decay_flag || dflag) {
    pathlen = pathlen + z_drift * std::sqrt(1 + t.dxdzs * t.dxdzs + t.dydzs * t.dydzs);
    t.xs = t.xs + t.dxdzs * z_drift;
    t.ys = t.ys + t.dydzs * z_drift;
    return;
  }
  const double p_spec = ph / (1. + t.dpps / 100.);
  const double beta = ph / std::sqrt(ph * ph + m2);
  const double gamma = 1. / std::sqrt(1. - beta * beta);
  const double dlen = t.ctau * beta * gamma;
  const double z_decay = -1. * dlen * std::log(1 - t.rng->grnd());
  if (z_decay > z_drift * std::sqrt(1 + t.dxdzs * t.dxdzs + t.dydzs * t.dydzs)) {
    t.decdist = t.decdist + z_drift * std::sqrt(1 + t.dxdzs * t.dxdzs + t.dydzs * t.dydzs);
    pathlen = pathlen + z_drift * std::sqrt(1 + t.dxdzs * t.dxdzs + t.dydzs * t.dydzs);
    t.xs = t.xs + t.dxdzs * z_drift;
    t.ys = t.ys + t.dydzs * z_drift;
  } else {
    dflag = true;
    t.decdist = t.decdist + z_decay;
    pathlen = pathlen + z_decay;
    t.xs = t.xs + t.dxdzs * z_decay / std::sqrt(1 + t.dxdzs * t.dxdzs + t.dydzs * t.dydzs);
    t.ys = t.ys + t.dydzs * z_decay / std::sqrt(1 + t.dxdzs * t.dxdzs + t.dydzs * t.dydzs);
    decay_in_flight(t, m2, ph, p_spec, beta, gamma, K::Mpi);
    const double tmpdrift = z_drift - z_decay / std::sqrt(1 + t.dxdzs * t.dxdzs + t.dydzs * t.dydzs);
    pathlen = pathlen + tmpdrift * std::sqrt(1 + t.dxdzs * t.dxdzs + t.dydzs * t.dydzs);
    t.xs = t.xs + t.dxdzs * tmpdrift;
    t.ys = t.ys + t.dydzs * tmpdrift;
  }
}

// shared/musc.f:1-58
static inline void musc(Rng& r, double m2, double p, double rad_len, double& dth, double& dph) {
  if (rad_len == 0) return;
  const double beta = p / std::sqrt(m2 + p * p);
  const double theta_sigma = 13.6 / p / beta * std::sqrt(rad_len) * (1 + 0.088 * std::log10(rad_len / (beta * beta)));
  dth = dth + theta_sigma * gauss1(r, 99.0);
  dph = dph + theta_sigma * gauss1(r, 99.0);
}

// shared/musc_ext.f:1-54; dummy order (dph,dth,y,x)
static inline void musc_ext(Rng& r, double m2, double p, double rad_len, double x_len, double& dph, double& dth,
                            double& y, double& x) {
  if (rad_len == 0) return;
  if (x_len <= 0 || rad_len < 0) throw std::runtime_error("x_len or rad_len < 0 in musc_ext");
  const double beta = p / std::sqrt(m2 + p * p);
  const double theta_sigma = 13.6 / p / beta * std::sqrt(rad_len) * (1 + 0.088 * std::log10(rad_len / (beta * beta)));
  double g1 = gauss1(r, 99.0);
  double g2 = gauss1(r, 99.0);
  dth = dth + theta_sigma * g1;
  x = x + theta_sigma * x_len * g2 / std::sqrt(12.) + theta_sigma * x_len * g1 / 2.;
  g1 = gauss1(r, 99.0);
  g2 = gauss1(r, 99.0);
  dph = dph + theta_sigma * g1;
  y = y + theta_sigma * x_len * g2 / std::sqrt(12.) + theta_sigma * x_len * g1 / 2.;
}

// shared/rotate_haxis.f:1-64
static inline void rotate_haxis(const Track& t, double rotang, double& xp0, double& yp0) {
  const double rotang_rad = rotang * 0.017453292;
  const double tan_th = std::tan(rotang_rad), sin_th = std::sin(rotang_rad), cos_th = std::cos(rotang_rad);
  const double alpha = t.dxdzs, beta = t.dydzs;
  const double alpha_p = (alpha + tan_th) / (1. - alpha * tan_th);
  const double beta_p = beta / (cos_th - alpha * sin_th);
  const double xi = xp0;
  xp0 = xi * (cos_th + alpha_p * sin_th);
  yp0 = yp0 + xi * beta_p * sin_th;
}

// shared/rotate_vaxis.f:1-60
static inline void rotate_vaxis(const Track& t, double rotang, double& xp0, double& yp0) {
  const double rotang_rad = rotang * 0.017453292;
  const double tan_th = std::tan(rotang_rad), sin_th = std::sin(rotang_rad), cos_th = std::cos(rotang_rad);
  const double alpha = t.dydzs, beta = t.dxdzs;
  const double alpha_p = (alpha + tan_th) / (1. - alpha * tan_th);
  const double beta_p = beta / (cos_th - alpha * sin_th);
  const double yi = yp0;
  yp0 = yi * (cos_th + alpha_p * sin_th);
  xp0 = xp0 + yi * beta_p * sin_th;
}

// ---- COSY forward maps: tables of shared/transp.f:85-89 -------------------------
struct CosyClass {
  std::vector<double> coeff;   // [n][5]
  std::vector<int8_t> expon;   // [n][5] : x, theta, y, phi, delta
  int n_terms = 0;
  double length = 0.;          // !LENGTH: comment, cm (0 if absent)
  bool adrift = true;
  double driftdist = 0.;
};
struct CosyForward {
  std::vector<CosyClass> cls;  // cls[k-1] = class k
  int n_classes() const { return (int)cls.size(); }
  // transp_init, shared/transp.f:294-474
  void load(const std::string& path);
};
struct CosyRecon {
  std::vector<double> coeff;   // [n][4]
  std::vector<int8_t> expon;   // [n][5] : x, x', y, y', fry
  int n_terms = 0;
  // hms/mc_hms_recon.f:70-102
  void load(const std::string& path);
  // hms/mc_hms_recon.f:104-137
  // clamp_all: sos/mc_sos_recon.f:79-81 moves every |hut(i)| <= 1e-30 to 1e-30, the other arms only hut(5)
  void eval(const Track& t, double fry, double& delta_p, double& delta_t, double& delta_phi, double& y_tgt,
            bool clamp_all = false) const;
};

// shared/transp.f:134-279
void transp(Track& t, const CosyForward& f, int klass, bool decay_flag, bool& dflag, double& m2, double& ph,
            double zd, double& pathlen);

}  // namespace simc_oracle
