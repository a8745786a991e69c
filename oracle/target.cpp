// ORACLE -- TEST INFRASTRUCTURE ONLY (see event.hpp).
// target.f (trip_thru_target, target_musc), enerloss_new.f.
#include <cmath>
#include "event.hpp"

namespace simc_oracle {

namespace T {   // target.inc:5-32
constexpr double rho_Al = 2.70, Z_Al = 13., A_Al = 26.98, X0_Al = 24.01, X0_cm_Al = X0_Al / rho_Al;
constexpr double rho_mylar = 1.39, Z_mylar = 4.545, A_mylar = 8.735, X0_mylar = 39.95, X0_cm_mylar = X0_mylar / rho_mylar;
constexpr double rho_kevlar = 0.74, Z_kevlar = 2.67, A_kevlar = 4.67, X0_kevlar = 55.2,
                 X0_cm_kevlar = X0_kevlar / rho_kevlar;
constexpr double rho_air = 0.00121, Z_air = 7.2, A_air = 14.4, X0_air = 36.66, X0_cm_air = X0_air / rho_air;
constexpr double inch_cm = 2.54, target_pi = 3.14159265358979;
}  // namespace T

// enerloss_new.f:1-85
void enerloss_new(Sim& s, double len, double dens, double zeff, double aeff, double epart, double mpart, int typeflag,
                  double& Eloss) {
  const double me = 0.51099906;
  const double thick = len * dens;
  const double gamma = epart / mpart;
  const double beta = std::sqrt(1. - 1. / (gamma * gamma));
  double I;
  if (zeff == 1) I = 21.8e-06;
  else I = (16. * std::pow(zeff, 0.9)) * 1.0e-06;
  const double hnup = 28.816e-06 * std::sqrt(dens * zeff / aeff);
  const double log10bg = std::log(beta * gamma) / std::log(10.);
  const double CO = std::log(hnup) - std::log(I) + 0.5;
  double denscorr;
  if (log10bg < 0.) denscorr = 0.;
  else if (log10bg < 3.) denscorr = CO + std::log(10.) * log10bg + std::fabs(CO / 27.) * powi(3. - log10bg, 3);
  else if (log10bg < 4.7) denscorr = CO + std::log(10.) * log10bg;
  else denscorr = CO + std::log(10.) * 4.7;
  if (thick <= 0.) {
    Eloss = 0.;
  } else {
    const double Eloss_mp_new =
        0.1536e-03 * zeff / aeff * thick / (beta * beta) *
        (std::log(me / (I * I)) + 1.063 + 2. * std::log(gamma * beta) +
         std::log(0.1536 * zeff / aeff * thick / (beta * beta)) - beta * beta - denscorr);
    const double Eloss_mp = Eloss_mp_new * 1000.;
    const double chsi = 0.307075 / 2. * zeff / aeff * thick / (beta * beta);
    double x = 0.;
    if (typeflag == 1) x = std::fabs(gauss1(*s.rng, 10.0));
    else if (typeflag == 2) x = 3;
    else if (typeflag == 3) x = 0.0067;
    else if (typeflag == 4) x = 1;
    double lambda;
    if (x > 0.0) lambda = -2.0 * std::log(x);
    else lambda = 100000.;
    Eloss = lambda * chsi + Eloss_mp;
  }
  if (Eloss > (epart - mpart)) Eloss = (epart - mpart) - 0.0000001;
}

// Material budget in front of one spectrometer, target.f:66-101 / :184-219
static bool arm_windows(int arm, double& s_Al, double& s_air, double& s_kevlar, double& s_mylar, bool& plus_angle) {
  using namespace T;
  plus_angle = false;
  if (arm == 1) { s_Al = 0.016 * inch_cm; s_air = 15; s_kevlar = 0.015 * inch_cm; s_mylar = 0.005 * inch_cm; plus_angle = true; }
  else if (arm == 2) { s_Al = 0.008 * inch_cm; s_air = 15; s_kevlar = 0.005 * inch_cm; s_mylar = 0.003 * inch_cm; }
  else if (arm == 3 || arm == 4) { s_Al = 0.013 * inch_cm; s_air = 15; s_kevlar = 0. * inch_cm; s_mylar = 0.010 * inch_cm; }
  else if (arm == 5 || arm == 6) { s_Al = (0.02 + 0.01) * inch_cm; s_air = 57.27; s_kevlar = 0.0; s_mylar = 0.0; }
  // Calorimeter arms: the reference has no branch for them (target.f:73-101, 188-222), so its SAVEd locals keep what
  // the previous call left (the other arm's windows, with the wall term added once more per call).  Restated as
  // what the source defines for them: no window material at all, the target and its can only; arm 7 sits on the HMS
  // side (theta + angle), arm 8 on the other.
  else if (arm == 7 || arm == 8) { s_Al = 0.0; s_air = 0.0; s_kevlar = 0.0; s_mylar = 0.0; plus_angle = arm == 7; }
  else return false;
  return true;
}

// target.f:1-306
void trip_thru_target(Sim& s, int narm, double zpos, double energy, double theta, double& Eloss, double& radlen,
                      double mass, int typeflag) {
  using namespace T;
  const simc_target& targ = s.cfg->targ;
  double s_Al = 0.0, s_target, s_air = 0, s_kevlar = 0, s_mylar = 0;
  double Eloss_target, Eloss_Al, Eloss_air, Eloss_kevlar, Eloss_mylar;
  const bool liquid = targ.Z < 2.4;
  if (narm == 1) {   // incoming electron, :34-52
    s_target = (targ.length / 2. + zpos) / std::fabs(std::cos(targ.angle));
    if (liquid) {
      if (targ.can == 1) s_Al = s_Al + 0.0028 * inch_cm;
      else if (targ.can == 2) s_Al = s_Al + 0.0050 * inch_cm;
      else if (targ.can == 3) s_Al = s_Al + 0.013;
    }
    radlen = s_target / targ.X0_cm + s_Al / X0_cm_Al;
    enerloss_new(s, s_target, targ.rho, targ.Z, targ.A, energy, mass, typeflag, Eloss_target);
    enerloss_new(s, s_Al, rho_Al, Z_Al, A_Al, energy, mass, typeflag, Eloss_Al);
    Eloss = Eloss_target + Eloss_Al;
    return;
  }
  // scattered electron (narm=2, :66-180) and hadron (narm=3, :184-305): same geometry code
  const int arm = (narm == 2) ? s.cfg->electron_arm : s.cfg->hadron_arm;
  bool plus_angle;
  if (!arm_windows(arm, s_Al, s_air, s_kevlar, s_mylar, plus_angle))
    throw std::runtime_error("trip_thru_target: unknown spectrometer");
  const double forward_path =
      (targ.length / 2. - zpos) / std::fabs(std::cos(plus_angle ? theta + targ.angle : theta - targ.angle));
  s_target = forward_path;
  if (liquid) {
    if (targ.can == 1) {          // beer can
      const double side_path = 1.325 * inch_cm / std::fabs(std::sin(theta));
      if (forward_path < side_path) {
        s_Al = s_Al + 0.005 * inch_cm / std::fabs(std::cos(theta));
      } else {
        s_target = side_path;
        s_Al = s_Al + 0.005 * inch_cm / std::fabs(std::sin(theta));
      }
    } else if (targ.can == 2) {   // pudding can
      const double t = powi(std::tan(theta), 2);
      const double atmp = 1 + t;
      const double btmp = -2 * zpos * t;
      const double ctmp = zpos * zpos * t - powi(targ.length / 2., 2);
      const double z_can = (-btmp + std::sqrt(btmp * btmp - 4. * atmp * ctmp)) / 2. / atmp;
      const double side_path = (z_can - zpos) / std::fabs(std::cos(theta));
      s_target = side_path;
      const double costmp = z_can / (targ.length / 2.);
      double th_can = 0.;   // static local in the reference: keeps its last value in the "else" case
      if (std::fabs(costmp) <= 1) th_can = std::acos(z_can / (targ.length / 2.));
      else if (std::fabs(costmp - 1.) <= 0.000001) th_can = 0.;
      s_Al = s_Al + 0.0050 * inch_cm / std::fabs(std::sin(target_pi / 2 - (theta - th_can)));
    } else if (targ.can == 3) {   // 2017 10 cm cryo cells
      const double ecir = 1.315 * 2.54;
      const double ecor = (1.315 + 0.0071) * 2.54;
      const double entec = targ.length - ecir;
      const double twall = ecor - ecir;
      const double tcm = zpos + targ.length / 2.0;
      double tliquid, tal;
      if ((tcm + ecir / std::tan(theta)) < entec) {
        tliquid = ecir / std::sin(theta);
        tal = twall / std::sin(theta);
      } else {
        tliquid = (std::sqrt(ecir * ecir - powi((targ.length - ecir - tcm) * std::sin(theta), 2)) +
                   (targ.length - ecir - tcm) * std::cos(theta));
        tal = +(std::sqrt(ecor * ecor - powi((targ.length - ecir - tcm) * std::sin(theta), 2)) -
                std::sqrt(ecir * ecir - powi((targ.length - ecir - tcm) * std::sin(theta), 2))) *
              twall / (ecor - ecir);
      }
      s_Al = s_Al + tal;
      s_target = tliquid;
    }
  }
  radlen = s_target / targ.X0_cm + s_Al / X0_cm_Al + s_air / X0_cm_air + s_kevlar / X0_cm_kevlar +
           s_mylar / X0_cm_mylar;
  enerloss_new(s, s_target, targ.rho, targ.Z, targ.A, energy, mass, typeflag, Eloss_target);
  enerloss_new(s, s_Al, rho_Al, Z_Al, A_Al, energy, mass, typeflag, Eloss_Al);
  enerloss_new(s, s_air, rho_air, Z_air, A_air, energy, mass, typeflag, Eloss_air);
  enerloss_new(s, s_kevlar, rho_kevlar, Z_kevlar, A_kevlar, energy, mass, typeflag, Eloss_kevlar);
  enerloss_new(s, s_mylar, rho_mylar, Z_mylar, A_mylar, energy, mass, typeflag, Eloss_mylar);
  Eloss = Eloss_target + Eloss_Al + Eloss_air + Eloss_kevlar + Eloss_mylar;
}

// target.f:548-577
void target_musc(Sim& s, double p, double beta, double teff, double dangles[2]) {
  const double theta_sigma = 13.6 / p / beta * std::sqrt(teff) * (1 + 0.088 * std::log10(teff / (beta * beta)));
  dangles[0] = theta_sigma * gauss1(*s.rng, 3.5);
  dangles[1] = theta_sigma * gauss1(*s.rng, 3.5);
}

}  // namespace simc_oracle
