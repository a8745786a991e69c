// Target geometry, energy loss and target multiple scattering, host+device.
// Replaces trip_thru_target / target_musc (target.f:1-306, 548-577) and enerloss_new
// (enerloss_new.f:1-85).  The per-particle part of enerloss_new (beta, gamma, the density
// correction inputs) is the same for every material of one trip and is evaluated once.
#pragma once
#include <math.h>
#include <cmath>
#include "../../include/simc_b200.h"

#if defined(__CUDACC__)
#define SIMC_HD __host__ __device__ __forceinline__
// Big routines are real functions on the device: the generation kernel inlined to 1.4 MB of SASS
// and stalled on instruction fetch (profiles/r1_notes.md); called code keeps it inside the i-cache.
#define SIMC_HD_CALL static __host__ __device__ __noinline__
#else
#define SIMC_HD inline
#define SIMC_HD_CALL static inline
#endif

// libm entry points as out-of-line device functions (each inlined pow/sin/cos/acos is 1-3 KB of SASS)
#include "fastlog.cuh"
namespace simc {
namespace m {
#if defined(__CUDA_ARCH__)
#define SIMC_MATH1(name) static __device__ __noinline__ double name(double x) { return ::name(x); }
// the logarithms are the loop's most frequent library calls: table-driven versions, 0.51 ulp (fastlog.cuh)
static __device__ __noinline__ double log(double x) { return fastlog::log(x); }
static __device__ __noinline__ double log10(double x) { return fastlog::log10(x); }
#else
#define SIMC_MATH1(name) static inline double name(double x) { return std::name(x); }
SIMC_MATH1(log) SIMC_MATH1(log10)
#endif
SIMC_MATH1(exp) SIMC_MATH1(sin) SIMC_MATH1(cos) SIMC_MATH1(tan)
SIMC_MATH1(acos) SIMC_MATH1(atan) SIMC_MATH1(asin)
#undef SIMC_MATH1
#if defined(__CUDA_ARCH__)
static __device__ __noinline__ double pow(double x, double y) { return ::pow(x, y); }
static __device__ __noinline__ double atan2(double y, double x) { return ::atan2(y, x); }
#else
static inline double pow(double x, double y) { return std::pow(x, y); }
static inline double atan2(double y, double x) { return std::atan2(y, x); }
#endif
}  // namespace m
}  // namespace simc

namespace simc {

// target.inc:5-32
struct Material { double rho, Z, A, X0_cm; };
#define SIMC_MAT_AL     Material{2.70, 13., 26.98, 24.01 / 2.70}
#define SIMC_MAT_MYLAR  Material{1.39, 4.545, 8.735, 39.95 / 1.39}
#define SIMC_MAT_KEVLAR Material{0.74, 2.67, 4.67, 55.2 / 0.74}
#define SIMC_MAT_AIR    Material{0.00121, 7.2, 14.4, 36.66 / 0.00121}

// What enerloss_new derives from (epart, mpart) alone, enerloss_new.f:26-28,38
struct ParticleKin {
  double epart, mpart, gamma, beta, beta2, log10bg, two_log_gb;
};
SIMC_HD ParticleKin particle_kin(double epart, double mpart, double ln10) {
  ParticleKin k;
  k.epart = epart; k.mpart = mpart;
  k.gamma = epart / mpart;
  k.beta = sqrt(1. - 1. / (k.gamma * k.gamma));
  k.beta2 = k.beta * k.beta;
  k.log10bg = m::log(k.beta * k.gamma) / ln10;
  k.two_log_gb = 2. * m::log(k.gamma * k.beta);
  return k;
}

// What enerloss_new computes from the material alone (enerloss_new.f:47-58,69-74): mean excitation
// energy, plasma term, and the leading factors of the three Z/A products (each is the first operation
// of its expression in the reference, so hoisting it does not change the rounding).  Evaluated once
// per run on the host; the event loop reads the table from kernel parameters.
struct MatConst { double rho, I, CO, co27, ln10, log_me_I2, p_mp, p_log, p_chsi; };
struct MatTable { MatConst targ, al, air, kevlar, mylar; };

SIMC_HD MatConst make_mat(double dens, double zeff, double aeff) {
  const double me = 0.51099906;
  MatConst c;
  c.rho = dens;
  if (zeff == 1) c.I = 21.8e-06;
  else c.I = (16. * m::pow(zeff, 0.9)) * 1.0e-06;
  const double hnup = 28.816e-06 * sqrt(dens * zeff / aeff);
  c.CO = m::log(hnup) - m::log(c.I) + 0.5;
  c.co27 = fabs(c.CO / 27.);
  c.ln10 = m::log(10.);
  c.log_me_I2 = m::log(me / (c.I * c.I));
  c.p_mp = 0.1536e-03 * zeff / aeff;
  c.p_log = 0.1536 * zeff / aeff;
  c.p_chsi = 0.307075 / 2. * zeff / aeff;
  return c;
}
SIMC_HD MatTable make_mat_table(const simc_target& targ) {
  MatTable t;
  t.targ = make_mat(targ.rho, targ.Z, targ.A);
  const Material al = SIMC_MAT_AL, air = SIMC_MAT_AIR, kev = SIMC_MAT_KEVLAR, myl = SIMC_MAT_MYLAR;
  t.al = make_mat(al.rho, al.Z, al.A);
  t.air = make_mat(air.rho, air.Z, air.A);
  t.kevlar = make_mat(kev.rho, kev.Z, kev.A);
  t.mylar = make_mat(myl.rho, myl.Z, myl.A);
  return t;
}

// enerloss_new.f:30-85.  x = |gauss1(10)| for typeflag 1 (drawn by the caller, only when
// thick > 0 -- the reference draws inside the thick>0 branch), 3 / 0.0067 / 1 for 2 / 3 / 4.
SIMC_HD_CALL double enerloss_material(const ParticleKin& k, double len, const MatConst& mc, double x) {
  const double thick = len * mc.rho;
  double eloss;
  if (thick <= 0.) {
    eloss = 0.;
  } else {
    const double CO = mc.CO;
    double denscorr;
    if (k.log10bg < 0.) denscorr = 0.;
    else if (k.log10bg < 3.) {
      const double d = 3. - k.log10bg;
      denscorr = CO + mc.ln10 * k.log10bg + mc.co27 * (d * (d * d));
    } else if (k.log10bg < 4.7) denscorr = CO + mc.ln10 * k.log10bg;
    else denscorr = CO + mc.ln10 * 4.7;
    const double eloss_mp_new = mc.p_mp * thick / k.beta2 *
                                (mc.log_me_I2 + 1.063 + k.two_log_gb + m::log(mc.p_log * thick / k.beta2) - k.beta2 -
                                 denscorr);
    const double eloss_mp = eloss_mp_new * 1000.;
    const double chsi = mc.p_chsi * thick / k.beta2;
    double lambda;
    if (x > 0.0) lambda = -2.0 * m::log(x);
    else lambda = 100000.;
    eloss = lambda * chsi + eloss_mp;
  }
  if (eloss > (k.epart - k.mpart)) eloss = (k.epart - k.mpart) - 0.0000001;
  return eloss;
}

SIMC_HD double typeflag_x(int typeflag) { return typeflag == 2 ? 3. : typeflag == 3 ? 0.0067 : 1.; }

// Window thicknesses in front of a spectrometer, target.f:66-101 (=184-219)
struct ArmWindows { double s_Al, s_air, s_kevlar, s_mylar; bool plus_angle; };
SIMC_HD ArmWindows arm_windows(int arm) {
  const double inch_cm = 2.54;
  ArmWindows w;
  w.plus_angle = false;
  if (arm == 1) { w.s_Al = 0.016 * inch_cm; w.s_air = 15; w.s_kevlar = 0.015 * inch_cm; w.s_mylar = 0.005 * inch_cm; w.plus_angle = true; }
  else if (arm == 2) { w.s_Al = 0.008 * inch_cm; w.s_air = 15; w.s_kevlar = 0.005 * inch_cm; w.s_mylar = 0.003 * inch_cm; }
  else if (arm == 3 || arm == 4) { w.s_Al = 0.013 * inch_cm; w.s_air = 15; w.s_kevlar = 0. * inch_cm; w.s_mylar = 0.010 * inch_cm; }
  // calorimeter arms (7 on the HMS side, 8 on the other): the reference defines no windows for them (target.f:73-101,
  // 188-222; its SAVEd locals keep what the previous call left): no window material, the target and its can only
  else if (arm == 7 || arm == 8) { w.s_Al = 0.0; w.s_air = 0.0; w.s_kevlar = 0.0; w.s_mylar = 0.0; w.plus_angle = arm == 7; }
  else { w.s_Al = (0.02 + 0.01) * inch_cm; w.s_air = 57.27; w.s_kevlar = 0.0; w.s_mylar = 0.0; }
  return w;
}

// Path lengths of trip_thru_target for an outgoing particle (narm = 2 or 3): target.f:102-168
SIMC_HD_CALL void outgoing_paths(const simc_target& targ, const ArmWindows& w, double zpos, double theta, double& s_target,
                            double& s_Al) {
  const double inch_cm = 2.54, target_pi = 3.14159265358979;
  s_Al = w.s_Al;
  const double forward_path =
      (targ.length / 2. - zpos) / fabs(m::cos(w.plus_angle ? theta + targ.angle : theta - targ.angle));
  s_target = forward_path;
  if (targ.Z < 2.4) {
    if (targ.can == 1) {
      const double side_path = 1.325 * inch_cm / fabs(m::sin(theta));
      if (forward_path < side_path) {
        s_Al = s_Al + 0.005 * inch_cm / fabs(m::cos(theta));
      } else {
        s_target = side_path;
        s_Al = s_Al + 0.005 * inch_cm / fabs(m::sin(theta));
      }
    } else if (targ.can == 2) {
      const double tt = m::tan(theta);
      const double t = tt * tt;
      const double atmp = 1 + t;
      const double btmp = -2 * zpos * t;
      const double hl = targ.length / 2.;
      const double ctmp = zpos * zpos * t - hl * hl;
      const double z_can = (-btmp + sqrt(btmp * btmp - 4. * atmp * ctmp)) / 2. / atmp;
      s_target = (z_can - zpos) / fabs(m::cos(theta));
      const double costmp = z_can / (targ.length / 2.);
      double th_can = 0.;
      if (fabs(costmp) <= 1) th_can = m::acos(z_can / (targ.length / 2.));
      s_Al = s_Al + 0.0050 * inch_cm / fabs(m::sin(target_pi / 2 - (theta - th_can)));
    } else if (targ.can == 3) {
      const double ecir = 1.315 * 2.54;
      const double ecor = (1.315 + 0.0071) * 2.54;
      const double entec = targ.length - ecir;
      const double twall = ecor - ecir;
      const double tcm = zpos + targ.length / 2.0;
      double tliquid, tal;
      if ((tcm + ecir / m::tan(theta)) < entec) {
        tliquid = ecir / m::sin(theta);
        tal = twall / m::sin(theta);
      } else {
        const double u = (targ.length - ecir - tcm) * m::sin(theta);
        tliquid = (sqrt(ecir * ecir - u * u) + (targ.length - ecir - tcm) * m::cos(theta));
        tal = +(sqrt(ecor * ecor - u * u) - sqrt(ecir * ecir - u * u)) * twall / (ecor - ecir);
      }
      s_Al = s_Al + tal;
      s_target = tliquid;
    }
  }
}

// Entrance side (narm = 1): target.f:34-52
SIMC_HD void incoming_paths(const simc_target& targ, double zpos, double& s_target, double& s_Al) {
  const double inch_cm = 2.54;
  s_Al = 0.0;
  s_target = (targ.length / 2. + zpos) / fabs(m::cos(targ.angle));
  if (targ.Z < 2.4) {
    if (targ.can == 1) s_Al = s_Al + 0.0028 * inch_cm;
    else if (targ.can == 2) s_Al = s_Al + 0.0050 * inch_cm;
    else if (targ.can == 3) s_Al = s_Al + 0.013;
  }
}

// trip_thru_target with a fixed typeflag 2/3/4 (no random numbers): used for the most-probable
// energy-loss correction (simc.f:1637-1645) and by the host-side init (init.f:44-56,
// target.f:310-544).  arm = spectrometer id for narm 2/3, ignored for narm 1.
SIMC_HD_CALL void trip_thru_target_fixed(const simc_target& targ, const MatTable& mt, int narm, int arm, double zpos,
                                    double energy, double theta, double mass, int typeflag, double& Eloss,
                                    double& radlen) {
  const Material al = SIMC_MAT_AL;
  const ParticleKin k = particle_kin(energy, mass, mt.targ.ln10);
  const double x = typeflag_x(typeflag);
  double s_target, s_Al;
  if (narm == 1) {
    incoming_paths(targ, zpos, s_target, s_Al);
    radlen = s_target / targ.X0_cm + s_Al / al.X0_cm;
    Eloss = enerloss_material(k, s_target, mt.targ, x) + enerloss_material(k, s_Al, mt.al, x);
    return;
  }
  const Material air = SIMC_MAT_AIR, kev = SIMC_MAT_KEVLAR, myl = SIMC_MAT_MYLAR;
  const ArmWindows w = arm_windows(arm);
  outgoing_paths(targ, w, zpos, theta, s_target, s_Al);
  radlen = s_target / targ.X0_cm + s_Al / al.X0_cm + w.s_air / air.X0_cm + w.s_kevlar / kev.X0_cm +
           w.s_mylar / myl.X0_cm;
  const double e1 = enerloss_material(k, s_target, mt.targ, x);
  const double e2 = enerloss_material(k, s_Al, mt.al, x);
  const double e3 = enerloss_material(k, w.s_air, mt.air, x);
  const double e4 = enerloss_material(k, w.s_kevlar, mt.kevlar, x);
  const double e5 = enerloss_material(k, w.s_mylar, mt.mylar, x);
  Eloss = e1 + e2 + e3 + e4 + e5;
}

}  // namespace simc
