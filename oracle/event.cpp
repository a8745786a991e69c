// ORACLE -- TEST INFRASTRUCTURE ONLY (see event.hpp).
// event.f (generate, complete_ev, complete_recon_ev, complete_main, physics_angles,
// spectrometer_angles), physics_proton.f (sigep, fofa_best_fit, sigMott), simc.f (montecarlo,
// loop body).  Reaction coverage: H(e,e'p) complete; the D/A(e,e'p) and hydrogen pi/K branches of
// generate/complete_ev are restated where they share code, their cross sections are not (yet).
#include <cmath>
#include "event.hpp"
#include "field.hpp"

namespace simc_oracle {

using std::acos;
using std::asin;
using std::atan;
using std::atan2;
using std::cos;
using std::fabs;
using std::sin;
using std::sqrt;
using std::tan;

// event.f:1572-1614
void physics_angles(double theta0, double phi0, double dx, double dy, double& theta, double& phi) {
  const double costh = cos(theta0), sinth = sin(theta0), sinph = sin(phi0);
  const double r = sqrt(1. + dx * dx + dy * dy);
  theta = acos((costh - dy * sinth * sinph) / r);
  if (dx != 0.0) {
    phi = atan((dy * costh + sinth * sinph) / dx);
    if (phi <= 0) phi = phi + K::pi;
    if (sinph < 0.) phi = phi + K::pi;
  } else {
    phi = phi0;
  }
}

// event.f:1618-1648
void spectrometer_angles(double theta0, double phi0, double& dx, double& dy, double theta, double phi) {
  const double x = sin(theta) * cos(phi), y = sin(theta) * sin(phi), z = cos(theta);
  const double x0 = sin(theta0) * cos(phi0), y0 = sin(theta0) * sin(phi0), z0 = cos(theta0);
  const double cos_dtheta = x * x0 + y * y0 + z * z0;
  dx = x / cos_dtheta;
  dy = sqrt(1 / (cos_dtheta * cos_dtheta) - 1. - dx * dx);
  const double y_event = y / cos_dtheta;
  if (y_event < y0) dy = -dy;
}

// physics_proton.f:137-172
static void fofa_best_fit(double qsquar, double& GE, double& GM) {
  const double mu_p = 2.793;
  const double Q2 = -qsquar * std::pow(K::hbarc, 2.) * 1.e-6;
  const double Q = sqrt(std::max(Q2, 0.e0));
  const double Q3 = std::pow(Q, 3.), Q4 = std::pow(Q, 4.), Q5 = std::pow(Q, 5.);
  double denom = 1. + 0.62 * Q + 0.68 * Q2 + 2.8 * Q3 + 0.83 * Q4;
  GE = 1. / denom;
  denom = 1. + 0.35 * Q + 2.44 * Q2 + 0.5 * Q3 + 1.04 * Q4 + 0.34 * Q5;
  GM = mu_p / denom;
}
// physics_proton.f:176-190
static double sigMott(double e0, double theta, double Q2) {
  const double sig = powi(2. * K::alpha * K::hbarc * e0 * cos(theta / 2.) / Q2, 2);
  return sig * 1.e4;
}
// physics_proton.f:1-22
double sigep(const Event& vertex) {
  const double q4sq = vertex.Q2;
  double GE, GM;
  fofa_best_fit(-q4sq / (K::hbarc * K::hbarc), GE, GM);
  const double qmu4mp = q4sq / 4. / K::Mp2;
  const double W1p = GM * GM * qmu4mp;
  const double W2p = (GE * GE + GM * GM * qmu4mp) / (1.0 + qmu4mp);
  const double Wp = W2p + 2. * W1p * powi(tan(vertex.e.theta / 2.), 2);
  return sigMott(vertex.e.E, vertex.e.theta, vertex.Q2) * vertex.e.E / vertex.Ein * Wp;
}

// generate_rho.f:1-131: the rho is thrown flat in cos(theta), phi in the virtual photon - nucleon centre of mass with a
// Breit-Wigner mass, then boosted to the lab.  Three random numbers; replaces COMMON Mh, Mh2.
bool generate_rho(Sim& s, Event& vertex) {
  const simc_run_config& cfg = *s.cfg;
  Rng& rng = *s.rng;
  const double nu = vertex.nu, Q2 = vertex.Q2, q = vertex.q;
  const double qux = vertex.uq.x, quy = vertex.uq.y, quz = vertex.uq.z;
  const double bx = -(q * qux + s.pferx * s.pfer) / (nu + s.efer);
  const double by = -(q * quy + s.pfery * s.pfer) / (nu + s.efer);
  const double bz = -(q * quz + s.pferz * s.pfer) / (nu + s.efer);
  const double betacm = sqrt(bx * bx + by * by + bz * bz);
  if (betacm > 1.0) return false;
  const double gammacm = 1. / sqrt(1.0 - betacm * betacm);
  const double ss = -Q2 + s.efer * s.efer + 2. * nu * s.efer;
  s.Mh = K::Mrho;
  s.Mh = s.Mh + 0.5 * 150.2 * tan((2. * rng.grnd() - 1.) * atan(2. * 500. / 150.2));
  s.Mh2 = s.Mh * s.Mh;
  s.ntup.rhomass = s.Mh;
  s.trk.Mh2_final = s.Mh;                       // as written (generate_rho.f:82); rho_decay overwrites it
  const double Erhocm = (ss + s.Mh2 - cfg.targ.Mrec_struck * cfg.targ.Mrec_struck) / 2. / sqrt(ss);
  if (Erhocm < s.Mh) return false;
  const double Prhocm = sqrt(Erhocm * Erhocm - s.Mh2);
  const double rph = rng.grnd() * 2. * K::pi;
  const double rth1 = rng.grnd() * 2. - 1.;
  const double rth = acos(rth1);
  const double pxr = Prhocm * sin(rth) * cos(rph);
  const double pyr = Prhocm * sin(rth) * sin(rph);
  const double pzr = Prhocm * cos(rth);
  const double er = Erhocm;
  double ef, pxf, pyf, pzf, pf;
  loren(gammacm, bx, by, bz, er, pxr, pyr, pzr, ef, pxf, pyf, pzf, pf);
  vertex.p.delta = 100. * (pf / cfg.spec_p.P - 1.);
  vertex.p.P = pf;
  vertex.p.E = ef;
  vertex.up.x = pxf / pf;
  vertex.up.y = pyf / pf;
  vertex.up.z = pzf / pf;
  vertex.p.xptar = acos(pzf / pf);              // "not really used for anything" (generate_rho.f:123-126)
  vertex.p.yptar = atan2(pyf, pxf);
  return true;
}

// rho_decay.f:1-163: rho0 -> pi+ pi-, the detected pion flat in phi and ~ sin^2 + 2 eps R cos^2 in the rho rest frame
// (rejection loop), boosted to the lab; `orig` becomes the pion that enters the hadron arm and COMMON Mh = Mpi.
bool rho_decay(Sim& s, Event& orig, double p_spec, double epsilon) {
  const simc_run_config& cfg = *s.cfg;
  Rng& rng = *s.rng;
  const double R_rho = 0.33 * std::pow(orig.Q2 / K::Mrho2, 0.61);
  const double ph = orig.p.P;
  const double beta = ph / sqrt(ph * ph + s.Mh2);
  const double gamma = 1. / sqrt(1. - beta * beta);
  const double rph = rng.grnd() * 2. * K::pi;
  double rth;
  for (;;) {
    const double rth1 = rng.grnd() * 2. - 1.;
    rth = acos(rth1);
    const double norm = (1.0 + 2.0 * epsilon * R_rho) * rng.grnd();
    const double dist = powi(sin(rth), 2) + 2.0 * epsilon * R_rho * powi(cos(rth), 2);
    if (!(dist < norm)) break;
  }
  s.ntup.rhotheta = rth;
  const double er = s.ntup.rhomass / 2.0;
  if (er < K::Mpi) return false;
  const double pr = sqrt(er * er - K::Mpi2);
  const double pxr = pr * sin(rth) * cos(rph);
  const double pyr = pr * sin(rth) * sin(rph);
  const double pzr = pr * cos(rth);
  const double bx = -beta * orig.up.x, by = -beta * orig.up.y, bz = -beta * orig.up.z;
  double ef, pxf, pyf, pzf, pf;
  loren(gamma, bx, by, bz, er, pxr, pyr, pzr, ef, pxf, pyf, pzf, pf);
  const int arm = cfg.hadron_arm;
  const double th0 = cfg.spec_p.theta;
  double pzprime;
  if (arm == 1 || arm == 3) pzprime = pzf * cos(th0) - pyf * sin(th0);
  else if (arm == 2 || arm == 4 || arm == 5) pzprime = pzf * cos(th0) + pyf * sin(th0);
  else throw std::runtime_error("Unknown spectrometer setup dude!");
  if (pzprime < 0.0) return false;
  const double th_oop = asin(pxf / pf);
  const double cos_th_inp = pzf / pf / cos(th_oop);
  double th_inp = acos(cos_th_inp);
  if (arm == 1 || arm == 3) {
    if (pyf < 0.0) th_inp = th0 - th_inp;
    else th_inp = th0 + th_inp;
  } else {
    if (pyf > 0.0) th_inp = th_inp - th0;
    else th_inp = th_inp + th0;
  }
  if ((th_inp > K::pi / 2.) || (th_oop > K::pi / 2.)) return false;
  orig.up.x = pxf / pf;
  orig.up.y = pyf / pf;
  orig.up.z = pzf / pf;
  orig.p.xptar = tan(th_oop);
  orig.p.yptar = tan(th_inp);
  orig.p.delta = 100. * (pf / p_spec - 1.);
  orig.p.P = pf;
  s.Mh = K::Mpi;
  s.Mh2 = K::Mpi2;
  s.trk.Mh2_final = s.Mh2;
  orig.p.E = sqrt(orig.p.P * orig.p.P + s.Mh2);
  physics_angles(cfg.spec_p.theta, cfg.spec_p.phi, orig.p.xptar, orig.p.yptar, orig.p.theta, orig.p.phi);
  return true;
}

// The polarised-target block both complete_ev (event.f:796-895) and complete_recon_ev (:1187-1263) carry: angle of the
// target polarisation about q relative to the scattering plane (phi_targ) and to the reaction plane (beta), the
// "Sivers" and "Collins" combinations with phi_pq, and the polar angle between q and the polarisation.
static void poltarg_block(const simc_run_config& cfg, const Vec3& uq, const Vec3& up, double phi_pq, double& phi_targ,
                          double& beta, double& phi_s, double& phi_c, double& theta_tarq) {
  double qx = -uq.y, qy = uq.x, qz = uq.z;
  double targx = -cfg.targ_pol * sin(fabs(cfg.targ_Bangle));
  double targy = 0.0;
  double targz = cfg.targ_pol * cos(fabs(cfg.targ_Bangle));
  double dummy = sqrt((qx * qx + qy * qy) * (qx * qx + qy * qy + qz * qz));
  double new_x_x = -qx * qz / dummy;
  double new_x_y = -qy * qz / dummy;
  double new_x_z = (qx * qx + qy * qy) / dummy;
  dummy = sqrt(qx * qx + qy * qy);
  double new_y_x = qy / dummy;
  double new_y_y = -qx / dummy;
  double new_y_z = 0.0;
  const double p_new_x = targx * new_x_x + targy * new_x_y + targz * new_x_z;
  const double p_new_y = targx * new_y_x + targy * new_y_y + targz * new_y_z;
  phi_targ = atan2(p_new_y, p_new_x);
  if (phi_targ < 0.) phi_targ = 2. * K::pi + phi_targ;
  const double px = -up.y, py = up.x, pz = up.z;
  dummy = sqrt(powi(qy * pz - qz * py, 2) + powi(qz * px - qx * pz, 2) + powi(qx * py - qy * px, 2));
  new_y_x = (qy * pz - qz * py) / dummy;
  new_y_y = (qz * px - qx * pz) / dummy;
  new_y_z = (qx * py - qy * px) / dummy;
  dummy = sqrt(powi(new_y_y * qz - new_y_z * qy, 2) + powi(new_y_z * qx - new_y_x * qz, 2) + powi(new_y_x * qy - new_y_y * qx, 2));
  new_x_x = (new_y_y * qz - new_y_z * qy) / dummy;
  new_x_y = (new_y_z * qx - new_y_x * qz) / dummy;
  new_x_z = (new_y_x * qy - new_y_y * qx) / dummy;
  const double targ_new_x = targx * new_x_x + targy * new_x_y + targz * new_x_z;
  const double targ_new_y = targx * new_y_x + targy * new_y_y + targz * new_y_z;
  beta = atan2(targ_new_y, targ_new_x);
  if (beta < 0.) beta = 2 * K::pi + beta;
  phi_s = phi_pq - phi_targ;
  if (phi_s < 0.) phi_s = 2 * K::pi + phi_s;
  phi_c = phi_pq + phi_targ;
  if (phi_c > 2. * K::pi) phi_c = phi_c - 2 * K::pi;
  if (phi_c < 0.0) phi_c = 2 * K::pi + phi_c;
  dummy = sqrt((qx * qx + qy * qy + qz * qz)) * sqrt((targx * targx + targy * targy + targz * targz));
  theta_tarq = acos((qx * targx + qy * targy + qz * targz) / dummy);
}

// pizero_decay.f:1-97: pi0 -> gamma gamma, flat in cos(theta) and phi in the pi0 rest frame, boosted to the lab.  Two
// random numbers.  (The routine's own energy / momentum sums only print a warning.)
void pizero_decay(Sim& s, const Event& vertex) {
  Rng& rng = *s.rng;
  const double Mgamma = 0.0;
  const double ph = vertex.p.P;
  const double eh = sqrt(ph * ph + s.Mh * s.Mh);
  const double beta = ph / eh;
  const double gamma = 1. / sqrt(1. - beta * beta);
  const double rph = rng.grnd() * 2. * K::pi;
  const double rth1 = rng.grnd() * 2. - 1.;
  const double rth = acos(rth1);
  const double er = s.Mh / 2.0;
  const double pr = sqrt(er * er - Mgamma * Mgamma);
  const double pxr1 = pr * sin(rth) * cos(rph), pyr1 = pr * sin(rth) * sin(rph), pzr1 = pr * cos(rth);
  const double pxr2 = -pxr1, pyr2 = -pyr1, pzr2 = -pzr1;
  const double bx = -beta * vertex.up.x, by = -beta * vertex.up.y, bz = -beta * vertex.up.z;
  double pf1, pf2;
  double* g1 = s.ntup.gamma1;
  double* g2 = s.ntup.gamma2;
  loren(gamma, bx, by, bz, er, pxr1, pyr1, pzr1, g1[0], g1[1], g1[2], g1[3], pf1);
  loren(gamma, bx, by, bz, er, pxr2, pyr2, pzr2, g2[0], g2[1], g2[2], g2[3], pf2);
}

// event.f:432-1052
bool complete_ev(Sim& s, EventMain& main, Event& vertex) {
  const simc_run_config& cfg = *s.cfg;
  const simc_target& targ = cfg.targ;
  double Mh = s.Mh, Mh2 = s.Mh2;          // COMMON: generate_rho replaces them (rho production only)
  main.jacobian = 1.0;
  vertex.ue.x = sin(vertex.e.theta) * cos(vertex.e.phi);
  vertex.ue.y = sin(vertex.e.theta) * sin(vertex.e.phi);
  vertex.ue.z = cos(vertex.e.theta);
  if (!cfg.doing_hyd_elast && !cfg.doing_rho) {
    vertex.up.x = sin(vertex.p.theta) * cos(vertex.p.phi);
    vertex.up.y = sin(vertex.p.theta) * sin(vertex.p.phi);
    vertex.up.z = cos(vertex.p.theta);
  }
  if (cfg.doing_hyd_elast) {
    vertex.e.E = vertex.Ein * Mh / (Mh + vertex.Ein * (1. - vertex.ue.z));
    if (vertex.e.E > vertex.Ein) return false;
    vertex.e.P = vertex.e.E;
    vertex.e.delta = (vertex.e.P - cfg.spec_e.P) * 100. / cfg.spec_e.P;
  }
  vertex.nu = vertex.Ein - vertex.e.E;
  vertex.Q2 = 2 * vertex.Ein * vertex.e.E * (1. - vertex.ue.z);
  vertex.q = sqrt(vertex.Q2 + vertex.nu * vertex.nu);
  vertex.xbj = vertex.Q2 / 2. / K::Mp / vertex.nu;
  vertex.uq.x = -vertex.e.P * vertex.ue.x / vertex.q;
  vertex.uq.y = -vertex.e.P * vertex.ue.y / vertex.q;
  vertex.uq.z = (vertex.Ein - vertex.e.P * vertex.ue.z) / vertex.q;
  if (fabs(vertex.uq.x * vertex.uq.x + vertex.uq.y * vertex.uq.y + vertex.uq.z * vertex.uq.z - 1) > 0.01)
    throw std::runtime_error("Error in q vector normalization");

  if (cfg.doing_hyd_elast) {   // :545-562
    vertex.Em = 0.0;
    vertex.Pm = 0.0;
    vertex.Mrec = 0.0;
    vertex.up.x = vertex.uq.x;
    vertex.up.y = vertex.uq.y;
    vertex.up.z = vertex.uq.z;
    vertex.p.P = vertex.q;
    vertex.p.theta = acos(vertex.up.z);
    vertex.p.phi = atan2(vertex.up.y, vertex.up.x);
    if (vertex.p.phi < 0.) vertex.p.phi = vertex.p.phi + 2. * K::pi;
    spectrometer_angles(cfg.spec_p.theta, cfg.spec_p.phi, vertex.p.xptar, vertex.p.yptar, vertex.p.theta,
                        vertex.p.phi);
    vertex.p.E = sqrt(vertex.p.P * vertex.p.P + Mh2);
    vertex.p.delta = (vertex.p.P - cfg.spec_p.P) * 100. / cfg.spec_p.P;
  } else if (cfg.doing_deuterium) {   // :564-607
    vertex.Em = targ.Mtar_struck + targ.Mrec - targ.M;
    vertex.Mrec = targ.M - targ.Mtar_struck + vertex.Em;
    const double a = -1. * vertex.q * (vertex.uq.x * vertex.up.x + vertex.uq.y * vertex.up.y + vertex.uq.z * vertex.up.z);
    const double b = vertex.q * vertex.q;
    const double c = vertex.nu + targ.M;
    const double t = c * c - b + Mh2 - vertex.Mrec * vertex.Mrec;
    const double QA = 4. * (a * a - c * c);
    const double QB = 4. * c * t;
    const double QC = -4. * a * a * Mh2 - t * t;
    const double radical = QB * QB - 4. * QA * QC;
    if (radical < 0) return false;
    vertex.p.E = (-QB - sqrt(radical)) / 2. / QA;
    if (vertex.p.E <= Mh) return false;
    vertex.p.P = sqrt(vertex.p.E * vertex.p.E - Mh2);
    vertex.p.delta = (vertex.p.P - cfg.spec_p.P) * 100. / cfg.spec_p.P;
    main.jacobian = (t * (c - vertex.p.E) + 2 * c * vertex.p.E * (vertex.p.E - c)) /
                    (2 * (a * a - c * c) * vertex.p.E + c * t);
    main.jacobian = fabs(main.jacobian);
  } else if (cfg.doing_pion || cfg.doing_kaon || cfg.doing_delta) {   // :609-698, hydrogen targets only here
    vertex.Pm = s.pfer;
    vertex.Mrec = targ.M - targ.Mtar_struck + vertex.Em;
    double a = -1. * vertex.q * (vertex.uq.x * vertex.up.x + vertex.uq.y * vertex.up.y + vertex.uq.z * vertex.up.z);
    double b = vertex.q * vertex.q;
    double c = vertex.nu + targ.M;
    if (cfg.doing_deutpi || cfg.doing_deutkaon || cfg.doing_hepi || cfg.doing_hekaon) {      // :646-654 Fermi motion and binding
      a = a - fabs(s.pfer) * (s.pferx * vertex.up.x + s.pfery * vertex.up.y + s.pferz * vertex.up.z);
      b = b + s.pfer * s.pfer +
          2 * vertex.q * fabs(s.pfer) * (s.pferx * vertex.uq.x + s.pfery * vertex.uq.y + s.pferz * vertex.uq.z);
      c = vertex.nu + s.efer;
    }
    const double t = c * c - b + Mh2 - targ.Mrec_struck * targ.Mrec_struck;
    const double QA = 4. * (a * a - c * c);
    const double QB = 4. * c * t;
    const double QC = -4. * a * a * Mh2 - t * t;
    const double radical = QB * QB - 4. * QA * QC;
    if (radical < 0) return false;
    vertex.p.E = (-QB - sqrt(radical)) / 2. / QA;
    if (vertex.p.E < 0.0) return false;
    const double Ehad2 = (-QB + sqrt(radical)) / 2. / QA;
    if (cfg.doing_delta) {                       // :682-684 choose one of the two solutions
      if (s.rng->grnd() > 0.5) vertex.p.E = Ehad2;
    }
    const double E_rec = c - vertex.p.E;
    if (E_rec <= targ.Mrec_struck) return false;
    if (vertex.p.E <= Mh) return false;
    vertex.p.P = sqrt(vertex.p.E * vertex.p.E - Mh2);
    vertex.p.delta = (vertex.p.P - cfg.spec_p.P) * 100. / cfg.spec_p.P;
  } else if (cfg.doing_rho) {                    // :701-708: the rho is thrown in 4 pi in the photon-nucleon c.m.
    if (!generate_rho(s, vertex)) return false;
    Mh = s.Mh; Mh2 = s.Mh2;
  } else if (cfg.doing_phsp) {
    vertex.p.P = cfg.spec_p.P;
    vertex.p.E = sqrt(Mh2 + vertex.p.P * vertex.p.P);
    vertex.p.delta = (vertex.p.P - cfg.spec_p.P) * 100. / cfg.spec_p.P;
  } else if (cfg.doing_heavy || cfg.doing_semi) {
    // nothing: E and P of both arms were generated
  } else {
    throw std::runtime_error("oracle: reaction not restated in complete_ev");
  }

  if (cfg.doing_pion || cfg.doing_kaon || cfg.doing_delta || cfg.doing_rho || cfg.doing_semi) {   // :707-771
    const double W2 = targ.Mtar_struck * targ.Mtar_struck + 2. * targ.Mtar_struck * vertex.nu - vertex.Q2;
    main.W = sqrt(fabs(W2)) * W2 / fabs(W2);
    main.epsilon = 1. / (1. + 2. * (1 + vertex.nu * vertex.nu / vertex.Q2) * powi(tan(vertex.e.theta / 2.), 2));
    main.theta_pq = acos(vertex.up.x * vertex.uq.x + vertex.up.y * vertex.uq.y + vertex.up.z * vertex.uq.z);
    main.t = vertex.Q2 - Mh2 + 2 * vertex.nu * vertex.p.E - 2 * vertex.p.P * vertex.q * cos(main.theta_pq);
    main.tmin = vertex.Q2 - Mh2 + 2 * vertex.p.E * vertex.nu - 2 * vertex.p.P * vertex.q;
    main.q2 = vertex.Q2;
    const double qx = -vertex.uq.y, qy = vertex.uq.x, qz = vertex.uq.z;
    const double px = -vertex.up.y, py = vertex.up.x, pz = vertex.up.z;
    double dummy = sqrt((qx * qx + qy * qy) * (qx * qx + qy * qy + qz * qz));
    const double new_x_x = -qx * qz / dummy, new_x_y = -qy * qz / dummy, new_x_z = (qx * qx + qy * qy) / dummy;
    dummy = sqrt(qx * qx + qy * qy);
    const double new_y_x = qy / dummy, new_y_y = -qx / dummy, new_y_z = 0.0;
    const double p_new_x = px * new_x_x + py * new_x_y + pz * new_x_z;
    const double p_new_y = px * new_y_x + py * new_y_y + pz * new_y_z;
    main.phi_pq = atan2(p_new_y, p_new_x);
    if (main.phi_pq < 0.e0) main.phi_pq = main.phi_pq + 2. * K::pi;
    if (cfg.using_tgt_field)                              // :796-895
      poltarg_block(cfg, vertex.uq, vertex.up, main.phi_pq, main.phi_targ, main.beta, vertex.phi_s, vertex.phi_c, main.theta_tarq);
    if (cfg.doing_pizero) pizero_decay(s, vertex);        // :899-901
  }

  // :880-955 missing momentum
  vertex.Pmx = vertex.p.P * vertex.up.x - vertex.q * vertex.uq.x;
  vertex.Pmy = vertex.p.P * vertex.up.y - vertex.q * vertex.uq.y;
  vertex.Pmz = vertex.p.P * vertex.up.z - vertex.q * vertex.uq.z;
  vertex.Pmiss = sqrt(vertex.Pmx * vertex.Pmx + vertex.Pmy * vertex.Pmy + vertex.Pmz * vertex.Pmz);
  vertex.Emiss = vertex.nu + targ.M - vertex.p.E;
  const double oop_x = -vertex.uq.y, oop_y = vertex.uq.x;
  vertex.PmPar = (vertex.Pmx * vertex.uq.x + vertex.Pmy * vertex.uq.y + vertex.Pmz * vertex.uq.z);
  vertex.PmOop = (vertex.Pmx * oop_x + vertex.Pmy * oop_y) / sqrt(oop_x * oop_x + oop_y * oop_y);
  vertex.PmPer = sqrt(std::max(0.e0, vertex.Pm * vertex.Pm - vertex.PmPar * vertex.PmPar - vertex.PmOop * vertex.PmOop));
  if (cfg.doing_hyd_elast) {
    vertex.Trec = 0.0;
  } else if (cfg.doing_deuterium) {
    vertex.Pm = vertex.Pmiss;
    vertex.Trec = sqrt(vertex.Mrec * vertex.Mrec + vertex.Pm * vertex.Pm) - vertex.Mrec;
  } else if (cfg.doing_heavy) {
    vertex.Pm = vertex.Pmiss;
    vertex.Mrec = sqrt(vertex.Emiss * vertex.Emiss - vertex.Pmiss * vertex.Pmiss);
    vertex.Em = targ.Mtar_struck + vertex.Mrec - targ.M;
    vertex.Trec = sqrt(vertex.Mrec * vertex.Mrec + vertex.Pm * vertex.Pm) - vertex.Mrec;
  } else if (cfg.doing_hydpi || cfg.doing_hydkaon || cfg.doing_delta || (cfg.doing_rho && std::lround(targ.A) == 1)) {
    vertex.Trec = 0.0;                         // doing_hyddelta, doing_hydrho: hydrogen only here
  } else if (cfg.doing_deutpi || cfg.doing_deutkaon || cfg.doing_hepi || cfg.doing_hekaon) {
    vertex.Trec = sqrt(vertex.Mrec * vertex.Mrec + vertex.Pm * vertex.Pm) - vertex.Mrec;
  } else if (cfg.doing_semi) {
    vertex.Pm = vertex.Pmiss;
    vertex.Em = vertex.Emiss;
  }
  s.ntup.krel = 0.0;
  if (cfg.doing_semi) {   // :979-996
    if ((powi(targ.Mtar_struck + vertex.nu - vertex.p.E, 2) - vertex.Pmiss * vertex.Pmiss) < powi(K::Mp + K::Mpi0, 2))
      return false;
    vertex.zhad = vertex.p.E / vertex.nu;
    vertex.pt2 = vertex.p.P * vertex.p.P * (1.0 - powi(cos(main.theta_pq), 2));
    if (vertex.zhad > 1.0) return false;
  }

  // :1013-1023 Jacobian of (xptar,yptar) -> solid angle
  double r = sqrt(1. + vertex.e.yptar * vertex.e.yptar + vertex.e.xptar * vertex.e.xptar);
  main.jacobian = main.jacobian / powi(r, 3);
  if (cfg.doing_heavy || cfg.doing_pion || cfg.doing_kaon || cfg.doing_delta || cfg.doing_semi) {
    r = sqrt(1. + vertex.p.yptar * vertex.p.yptar + vertex.p.xptar * vertex.p.xptar);
    main.jacobian = main.jacobian / powi(r, 3);
  }
  // :1031-1040 energy loss of the outgoing particles
  trip_thru_target(s, 2, main.target.z - targ.zoffset, vertex.e.E, vertex.e.theta, main.target.Eloss[1],
                   main.target.teff[1], K::Me, 1);
  trip_thru_target(s, 3, main.target.z - targ.zoffset, vertex.p.E, vertex.p.theta, main.target.Eloss[2],
                   main.target.teff[2], Mh, 1);
  if (!cfg.using_Eloss) {
    main.target.Eloss[1] = 0.0;
    main.target.Eloss[2] = 0.0;
  }
  radc_init_ev(s, main, vertex);
  return true;
}

// event.f:126-428
bool generate(Sim& s, EventMain& main, Event& vertex, Event& orig) {
  const simc_run_config& cfg = *s.cfg;
  const simc_target& targ = cfg.targ;
  Rng& rng = *s.rng;
  const simc_gen_limits& gen = cfg.gen;
  main.target.x = gauss1(rng, 3.0) * gen.xwid + targ.xoffset;
  main.target.y = gauss1(rng, 3.0) * gen.ywid + targ.yoffset;
  double t3, t4, t5, t6;
  if (targ.fr_pattern == 1) {
    t3 = rng.grnd() * K::pi;
    t4 = rng.grnd() * K::pi;
    t5 = cos(t3) * targ.fr1;
    t6 = cos(t4) * targ.fr2;
  } else if (targ.fr_pattern == 2) {
    t3 = rng.grnd() * 2. * K::pi;
    t4 = sqrt(rng.grnd()) * (targ.fr2 - targ.fr1) + targ.fr1;
    t5 = cos(t3) * t4;
    t6 = sin(t3) * t4;
  } else if (targ.fr_pattern == 3) {
    t3 = 2. * rng.grnd() - 1.0;
    t4 = 2. * rng.grnd() - 1.0;
    t5 = targ.fr1 * t3;
    t6 = targ.fr2 * t4;
  } else {
    t5 = 0.0;
    t6 = 0.0;
  }
  main.target.x = main.target.x + t5;
  main.target.y = main.target.y + t6;
  main.target.z = (0.5 - rng.grnd()) * targ.length + targ.zoffset;
  main.target.rastery = t6;
  main.target.rasterx = t5;
  trip_thru_target(s, 1, main.target.z - targ.zoffset, cfg.Ebeam, 0.0e0, main.target.Eloss[0], main.target.teff[0],
                   K::Me, 1);
  if (!cfg.using_Eloss) main.target.Eloss[0] = 0.0;
  if (cfg.using_Coulomb) main.target.Coulomb = targ.Coulomb_constant;
  else main.target.Coulomb = 0.0;
  vertex.Ein = cfg.Ebeam + (rng.grnd() - 0.5) * cfg.dEbeam + main.target.Coulomb - main.target.Eloss[0];
  main.Ein_shift = vertex.Ein - cfg.Ebeam_vertex_ave;
  main.Ee_shift = main.target.Coulomb - targ.Coulomb_ave;
  main.gen_weight = 1.0;

  vertex.e.yptar = gen.e.yptar.min + rng.grnd() * (gen.e.yptar.max - gen.e.yptar.min);
  vertex.e.xptar = gen.e.xptar.min + rng.grnd() * (gen.e.xptar.max - gen.e.xptar.min);
  if (cfg.doing_deuterium || cfg.doing_heavy || cfg.doing_pion || cfg.doing_kaon || cfg.doing_delta || cfg.doing_semi) {
    vertex.p.yptar = gen.p.yptar.min + rng.grnd() * (gen.p.yptar.max - gen.p.yptar.min);
    vertex.p.xptar = gen.p.xptar.min + rng.grnd() * (gen.p.xptar.max - gen.p.xptar.min);
  }
  if (cfg.doing_heavy || cfg.doing_semi) {
    const double Emin = std::max(gen.p.E.min, gen.sumEgen.min - gen.e.E.max);
    const double Emax = std::min(gen.p.E.max, gen.sumEgen.max - gen.e.E.min);
    if (Emin > Emax) return false;
    main.gen_weight = main.gen_weight * (Emax - Emin) / (gen.p.E.max - gen.p.E.min);
    vertex.p.E = Emin + rng.grnd() * (Emax - Emin);
    vertex.p.P = sqrt(vertex.p.E * vertex.p.E - s.Mh2);
    vertex.p.delta = 100. * (vertex.p.P - cfg.spec_p.P) / cfg.spec_p.P;
  }
  if (cfg.doing_deuterium || cfg.doing_heavy || cfg.doing_pion || cfg.doing_kaon || cfg.doing_delta || cfg.doing_rho ||
      cfg.doing_semi) {
    double Emin = gen.e.E.min, Emax = gen.e.E.max;
    if (cfg.doing_deuterium || cfg.doing_pion || cfg.doing_kaon || cfg.doing_delta || cfg.doing_rho) {
      Emin = std::max(Emin, gen.sumEgen.min);
      Emax = std::min(Emax, gen.sumEgen.max);
    } else if (cfg.doing_heavy) {
      Emin = std::max(Emin, gen.sumEgen.min - vertex.p.E);
      Emax = std::min(Emax, gen.sumEgen.max - vertex.p.E);
    }
    if (Emin > Emax) return false;
    main.gen_weight = main.gen_weight * (Emax - Emin) / (gen.e.E.max - gen.e.E.min);
    vertex.e.E = Emin + rng.grnd() * (Emax - Emin);
    vertex.e.P = vertex.e.E;
    vertex.e.delta = 100. * (vertex.e.P - cfg.spec_e.P) / cfg.spec_e.P;
  }
  physics_angles(cfg.spec_e.theta, cfg.spec_e.phi, vertex.e.xptar, vertex.e.yptar, vertex.e.theta, vertex.e.phi);
  physics_angles(cfg.spec_p.theta, cfg.spec_p.phi, vertex.p.xptar, vertex.p.yptar, vertex.p.theta, vertex.p.phi);
  // :327-373 Fermi momentum (drawn for deuterium semi-inclusive production whether or not do_fermi is set)
  s.pfer = 0.0; s.pferx = 0.0; s.pfery = 0.0; s.pferz = 0.0;
  vertex.Em = 0.0;
  s.efer = targ.Mtar_struck;
  if (cfg.doing_deutsemi || cfg.doing_deutpi || cfg.doing_deutkaon || cfg.doing_hepi || cfg.doing_hekaon) {
    if (!s.pfermi || s.pfermi->pval.empty()) throw std::runtime_error("oracle: momentum distribution not set");
    const std::vector<double>& pval = s.pfermi->pval;
    const std::vector<double>& mprob = s.pfermi->mprob;
    const int nump = (int)pval.size();
    const double ranprob = rng.grnd();
    int ii = 1;
    while (ranprob > mprob[ii - 1] && ii < nump) ii = ii + 1;
    double pferlo, pferhi;
    if (ii == 1) pferlo = 0;
    else pferlo = (pval[ii - 2] + pval[ii - 1]) / 2;
    if (ii == nump) pferhi = pval[nump - 1];
    else pferhi = (pval[ii - 1] + pval[ii]) / 2;
    s.pfer = pferlo + (pferhi - pferlo) * rng.grnd();
    const double ranth1 = rng.grnd() * 2. - 1.0;
    const double ranth = acos(ranth1);
    const double ranph = rng.grnd() * 2. * K::pi;
    s.pferx = sin(ranth) * cos(ranph);
    s.pfery = sin(ranth) * sin(ranph);
    s.pferz = cos(ranth);
    if (cfg.doing_hepi || cfg.doing_hekaon) {      // :368-372
      if (!s.sf) throw std::runtime_error("oracle: spectral-function table not set");
      vertex.Em = generate_em(*s.sf, rng, s.pfer);
    } else {
      vertex.Em = K::Mp + K::Mn - targ.M;
    }
    const double m_spec = targ.M - targ.Mtar_struck + vertex.Em;
    s.efer = targ.M - sqrt(m_spec * m_spec + s.pfer * s.pfer);
  }
  if (!complete_ev(s, main, vertex)) return false;
  main.sigcc = 1.0;
  main.Trec = vertex.Trec;
  bool success;
  if (cfg.using_rad) {
    success = generate_rad(s, main, vertex, orig);
  } else {
    success = true;
    if (cfg.doing_heavy)
      success = (vertex.Em >= cfg.VERTEXedge.Em.min && vertex.Em <= cfg.VERTEXedge.Em.max &&
                 vertex.Pm >= cfg.VERTEXedge.Pm.min && vertex.Pm <= cfg.VERTEXedge.Pm.max);
    if (success) orig = vertex;
  }
  // :420-422.  The reference calls rho_decay whatever generate_rad returned; for a try that already failed it would
  // only burn random numbers of a stream nobody reads again (and `orig` is then the previous event's): skipped.
  if (cfg.doing_rho && success) success = rho_decay(s, orig, cfg.spec_p.P, main.epsilon);
  return success;
}

// One arm of montecarlo: dispatch of simc.f:1463-1488 / :1716-1745
static void run_arm(Sim& s, int arm_id, const ArmOptics* o, ArmCall& a) {
  if (!o) throw std::runtime_error("oracle: optics not set for an arm in use");
  if (arm_id == 1) mc_hms(s.trk, *o, a);
  else if (arm_id == 5) mc_shms(s.trk, *o, a);
  else if (arm_id == 2) mc_sos(s.trk, *o, a);
  else if (arm_id == 3 || arm_id == 4) mc_hrs(s.trk, *o, a, arm_id == 3);
  else throw std::runtime_error("oracle: spectrometer not restated");
}

// calo/mc_calo.f:1-191: the calorimeter arm.  A field-free drift to the front face and the two half-size cuts (the
// NPS: 30 x 36 blocks of 2.05 cm); dpp, y and the slopes are returned as they came in.
namespace calo_stop { enum { OK = 0, SLIT_HOR, SLIT_VERT }; }
static void mc_calo(Track& t, ArmCall& a, double drift_to_cal) {
  const double h_entr = 30.75, v_entr = 36.9;
  a.ok_spec = false;
  bool dflag = false;
  t.xs = a.x; t.ys = a.y; t.zs = a.z; t.dxdzs = a.dxdz; t.dydzs = a.dydz; t.dpps = a.dpp;
  double p = a.p_spec * (1. + t.dpps / 100.);
  project(t, drift_to_cal, a.decay_flag, dflag, a.m2, p, a.pathlen);
  if (fabs(t.ys) > h_entr) { a.stop_code = calo_stop::SLIT_HOR; return; }
  if (fabs(t.xs) > v_entr) { a.stop_code = calo_stop::SLIT_VERT; return; }
  a.x_fp = t.xs; a.y_fp = t.ys; a.dx_fp = t.dxdzs; a.dy_fp = t.dydzs;
  a.ok_spec = true;
}

// simc.f:1310-1852 (using_tgt_field = .false., no calorimeter arms)
bool montecarlo(Sim& s, Event& orig, EventMain& main, Event& recon) {
  const simc_run_config& cfg = *s.cfg;
  const double Mh2 = s.Mh2, Mh = s.Mh;      // rho production: the decay pion's (rho_decay.f:150-153)
  s.ntup.resfac = 0.0;
  double fry;
  if (cfg.correct_raster) fry = -main.target.rastery;
  else fry = 0.0;
  double dang_in[2], dangles[2];
  if (cfg.mc_smear) target_musc(s, orig.Ein, 1., main.target.teff[0], dang_in);
  else { dang_in[0] = 0.0; dang_in[1] = 0.0; }

  // ---- P arm
  if (cfg.using_Eloss)
    main.SP_p.delta = (sqrt(fabs(powi(orig.p.E - main.target.Eloss[2], 2) - Mh2)) - cfg.spec_p.P) / cfg.spec_p.P * 100.;
  else
    main.SP_p.delta = orig.p.delta;
  if (cfg.mc_smear) {
    const double beta = orig.p.P / orig.p.E;
    target_musc(s, orig.p.P, beta, main.target.teff[2], dangles);
  } else { dangles[0] = 0.0; dangles[1] = 0.0; }
  main.SP_p.yptar = orig.p.yptar + dangles[0] + dang_in[0];
  main.SP_p.xptar = orig.p.xptar + dangles[1] + dang_in[1] * cfg.spec_p.cos_th;
  if (cfg.using_P_arm_montecarlo) {
    double x_P_arm = -main.target.y;
    double y_P_arm = -main.target.x * cfg.spec_p.cos_th - main.target.z * cfg.spec_p.sin_th * sin(cfg.spec_p.phi);
    double z_P_arm = main.target.z * cfg.spec_p.cos_th + main.target.x * cfg.spec_p.sin_th * sin(cfg.spec_p.phi);
    x_P_arm = x_P_arm - cfg.spec_p.off_x;
    y_P_arm = y_P_arm - cfg.spec_p.off_y;
    z_P_arm = z_P_arm - cfg.spec_p.off_z;
    double dx_P_arm = main.SP_p.xptar - cfg.spec_p.off_xptar;
    double dy_P_arm = main.SP_p.yptar - cfg.spec_p.off_yptar;
    if (cfg.using_tgt_field)                          // simc.f:1425-1432 (its ok is overwritten by the arm's own)
      track_from_tgt(*s.field, x_P_arm, y_P_arm, z_P_arm, dx_P_arm, dy_P_arm,
                     cfg.sign_hadron * cfg.spec_p.P * (1 + main.SP_p.delta / 100.), Mh, 1);
    x_P_arm = x_P_arm - z_P_arm * dx_P_arm;
    y_P_arm = y_P_arm - z_P_arm * dy_P_arm;
    z_P_arm = 0.0;
    const double xtar_init_P = x_P_arm;
    main.SP_p.z = y_P_arm;
    ArmCall a;
    a.p_spec = cfg.spec_p.P; a.th_spec = cfg.spec_p.theta; a.dpp = main.SP_p.delta;
    a.x = x_P_arm; a.y = y_P_arm; a.z = z_P_arm; a.dxdz = dx_P_arm; a.dydz = dy_P_arm;
    a.m2 = Mh2; a.ms_flag = cfg.mc_smear; a.wcs_flag = cfg.mc_smear; a.decay_flag = cfg.doing_decay;
    a.pathlen = 0.0;
    const int arm = cfg.hadron_arm;
    a.fry = (arm == 1 || arm == 5 || arm == 6) ? xtar_init_P : fry;
    a.using_coll = (arm == 1) ? cfg.using_HMScoll : (arm == 5 ? cfg.using_SHMScoll : 0);
    s.trk.calls = s.calls[1];
    if (arm == 7 || arm == 8) {                      // simc.f:1489-1564
      if (cfg.doing_pizero) {                        // one call per photon
        const double* g1 = s.ntup.gamma1;
        const double* g2 = s.ntup.gamma2;
        const double cth = cos(cfg.spec_p.theta), sth = sin(cfg.spec_p.theta);
        double exrot1, eyrot1, ezrot1, exrot2, eyrot2, ezrot2;
        if (arm == 8) {
          exrot1 = g1[1]; eyrot1 = g1[2] * cth - g1[3] * sth; ezrot1 = g1[2] * sth + g1[3] * cth;
          exrot2 = g2[1]; eyrot2 = g2[2] * cth - g2[3] * sth; ezrot2 = g2[2] * sth + g2[3] * cth;
        } else {
          exrot1 = g1[1]; eyrot1 = g1[2] * cth + g1[3] * sth; ezrot1 = -g1[2] * sth + g1[3] * cth;
          exrot2 = g2[1]; eyrot2 = g2[2] * cth + g2[3] * sth; ezrot2 = -g2[2] * sth + g2[3] * cth;
        }
        a.dxdz = exrot1 / ezrot1; a.dydz = eyrot1 / ezrot1;
        mc_calo(s.trk, a, cfg.drift_to_cal);
        const bool ok_gamma1 = a.ok_spec;
        const int stop1 = a.stop_code;
        s.ntup.xcal_gamma1 = -1.0e10; s.ntup.ycal_gamma1 = -1.0e10;
        if (ok_gamma1) { s.ntup.xcal_gamma1 = a.x_fp; s.ntup.ycal_gamma1 = a.y_fp; }
        a.dxdz = exrot2 / ezrot2; a.dydz = eyrot2 / ezrot2;
        a.stop_code = 0;
        mc_calo(s.trk, a, cfg.drift_to_cal);
        const bool ok_gamma2 = a.ok_spec;
        s.ntup.xcal_gamma2 = -1.0e10; s.ntup.ycal_gamma2 = -1.0e10;
        if (ok_gamma2) { s.ntup.xcal_gamma2 = a.x_fp; s.ntup.ycal_gamma2 = a.y_fp; }
        a.ok_spec = false;
        if (cfg.pizero_ngamma == 2) a.ok_spec = ok_gamma1 && ok_gamma2;
        else if (cfg.pizero_ngamma == 1) a.ok_spec = ok_gamma1 || ok_gamma2;
        else throw std::runtime_error("pizero_ngamma not set correctly (should be 1 or 2), stopping");
        if (!a.ok_spec && a.stop_code == 0) a.stop_code = stop1;     // our bookkeeping: where the first lost photon stopped
      } else {
        mc_calo(s.trk, a, cfg.drift_to_cal);
      }
    } else {
      run_arm(s, arm, s.optics_p, a);
    }
    s.ntup.resfac = a.resmult;
    for (int k = 0; k < 3; ++k) s.coll_steps[1][k] = a.coll_steps[k];
    s.stop_p = a.ok_spec ? 0 : a.stop_code;
    s.hut_p = a.reached_hut;
    if (cfg.using_tgt_field && a.ok_spec) {           // simc.f:1573-1587 (for a rejected track it changes nothing)
      const double frx = cfg.correct_raster ? -main.target.rasterx : 0.0;
      const double phad = cfg.spec_p.P * (1. + a.dpp / 100.0);
      const double ctheta = cos(cfg.spec_p.theta), stheta = sin(cfg.spec_p.theta);
      const bool right = arm == 1 || arm == 3;
      if (!right && !(arm == 2 || arm == 4 || arm == 5))
        throw std::runtime_error("Target field reconstruction not set up for your spectrometer");
      const ArmOptics& o = *s.optics_p;
      auto recon = [&](double& delta, double& dy, double& dx, double& y, double xxd) {
        o.rec.eval(s.trk, xxd, delta, dy, dx, y, /*clamp_all=*/arm == 2);
      };
      a.ok_spec = track_to_tgt(*s.field, a.dpp, a.y, a.dxdz, a.dydz, frx, -fry, cfg.sign_hadron * phad, sqrt(a.m2), ctheta,
                               right ? stheta : -stheta, 1, a.ok_spec, recon);
      s.field_fail_p = !a.ok_spec;
    }
    if (!a.ok_spec) return false;
    main.RECON_p.delta = a.dpp;
    main.RECON_p.yptar = a.dydz;
    main.RECON_p.xptar = a.dxdz;
    main.RECON_p.z = a.y;
    main.FP_p.x = a.x_fp; main.FP_p.dx = a.dx_fp; main.FP_p.y = a.y_fp; main.FP_p.dy = a.dy_fp;
    main.FP_p.path = a.pathlen;
    if (cfg.doing_pizero) {                          // :1605-1609: "no pi0 reconstruction yet"
      main.RECON_p.delta = main.SP_p.delta;
      main.RECON_p.yptar = main.SP_p.yptar;
      main.RECON_p.xptar = main.SP_p.xptar;
    }
  } else {
    main.RECON_p.delta = main.SP_p.delta;
    main.RECON_p.yptar = main.SP_p.yptar;
    main.RECON_p.xptar = main.SP_p.xptar;
  }
  recon.p.delta = main.RECON_p.delta;
  recon.p.yptar = main.RECON_p.yptar;
  recon.p.xptar = main.RECON_p.xptar;
  recon.p.z = main.RECON_p.z;
  recon.p.P = cfg.spec_p.P * (1. + recon.p.delta / 100.);
  recon.p.E = sqrt(recon.p.P * recon.p.P + Mh2);
  double dx_tmp = recon.p.xptar + cfg.spec_p.off_xptar;
  double dy_tmp = recon.p.yptar + cfg.spec_p.off_yptar;
  physics_angles(cfg.spec_p.theta, cfg.spec_p.phi, dx_tmp, dy_tmp, recon.p.theta, recon.p.phi);
  if (cfg.correct_Eloss) {
    double eloss_P_arm, r;
    trip_thru_target(s, 3, 0.0, recon.p.E, recon.p.theta, eloss_P_arm, r, Mh, 4);
    recon.p.E = recon.p.E + eloss_P_arm;
    recon.p.E = std::max(recon.p.E, sqrt(Mh2 + 0.000001));
    recon.p.P = sqrt(recon.p.E * recon.p.E - Mh2);
  }

  // ---- E arm
  main.SP_e.delta = 100 * (orig.e.E - main.target.Eloss[1] - main.target.Coulomb - cfg.spec_e.P) / cfg.spec_e.P;
  if (cfg.mc_smear) target_musc(s, orig.e.P, 1., main.target.teff[1], dangles);
  else { dangles[0] = 0.0; dangles[1] = 0.0; }
  main.SP_e.yptar = orig.e.yptar + dangles[0] + dang_in[0];
  main.SP_e.xptar = orig.e.xptar + dangles[1] + dang_in[1] * cfg.spec_e.cos_th;
  if (cfg.using_E_arm_montecarlo) {
    double x_E_arm = -main.target.y;
    double y_E_arm = -main.target.x * cfg.spec_e.cos_th - main.target.z * cfg.spec_e.sin_th * sin(cfg.spec_e.phi);
    double z_E_arm = main.target.z * cfg.spec_e.cos_th + main.target.x * cfg.spec_e.sin_th * sin(cfg.spec_e.phi);
    x_E_arm = x_E_arm - cfg.spec_e.off_x;
    y_E_arm = y_E_arm - cfg.spec_e.off_y;
    z_E_arm = z_E_arm - cfg.spec_e.off_z;
    double dx_E_arm = main.SP_e.xptar - cfg.spec_e.off_xptar;
    double dy_E_arm = main.SP_e.yptar - cfg.spec_e.off_yptar;
    if (cfg.using_tgt_field)                          // simc.f:1693-1700
      track_from_tgt(*s.field, x_E_arm, y_E_arm, z_E_arm, dx_E_arm, dy_E_arm, -cfg.spec_e.P * (1 + main.SP_e.delta / 100.),
                     K::Me, -1);
    x_E_arm = x_E_arm - z_E_arm * dx_E_arm;
    y_E_arm = y_E_arm - z_E_arm * dy_E_arm;
    z_E_arm = 0.0;
    const double xtar_init_E = x_E_arm;
    main.SP_e.z = y_E_arm;
    ArmCall a;
    a.p_spec = cfg.spec_e.P; a.th_spec = cfg.spec_e.theta; a.dpp = main.SP_e.delta;
    a.x = x_E_arm; a.y = y_E_arm; a.z = z_E_arm; a.dxdz = dx_E_arm; a.dydz = dy_E_arm;
    a.m2 = K::Me2; a.ms_flag = cfg.mc_smear; a.wcs_flag = cfg.mc_smear; a.decay_flag = false;
    a.pathlen = 0.0;
    const int arm = cfg.electron_arm;
    a.fry = (arm == 1 || arm == 5 || arm == 6) ? xtar_init_E : fry;
    a.using_coll = (arm == 1) ? cfg.using_HMScoll : (arm == 5 ? cfg.using_SHMScoll : 0);
    s.trk.calls = s.calls[0];
    run_arm(s, arm, s.optics_e, a);
    s.ntup.resfac = s.ntup.resfac + a.resmult;
    for (int k = 0; k < 3; ++k) s.coll_steps[0][k] = a.coll_steps[k];
    s.stop_e = a.ok_spec ? 0 : a.stop_code;
    s.hut_e = a.reached_hut;
    if (cfg.using_tgt_field && a.ok_spec) {           // simc.f:1777-1790
      const double frx = cfg.correct_raster ? -main.target.rasterx : 0.0;
      const double pelec = cfg.spec_e.P * (1. + a.dpp / 100.0);
      const double ctheta = cos(cfg.spec_e.theta), stheta = sin(cfg.spec_e.theta);
      const bool right = arm == 1 || arm == 3;
      if (!right && !(arm == 2 || arm == 4 || arm == 5))
        throw std::runtime_error("Target field reconstruction not set up for your spectrometer");
      const ArmOptics& o = *s.optics_e;
      auto recon = [&](double& delta, double& dy, double& dx, double& y, double xxd) {
        o.rec.eval(s.trk, xxd, delta, dy, dx, y, /*clamp_all=*/arm == 2);
      };
      a.ok_spec = track_to_tgt(*s.field, a.dpp, a.y, a.dxdz, a.dydz, frx, -fry, -1.0 * pelec, sqrt(K::Me2), ctheta,
                               right ? stheta : -stheta, -1, a.ok_spec, recon);
    }
    if (!a.ok_spec) return false;
    main.RECON_e.delta = a.dpp;
    main.RECON_e.yptar = a.dydz;
    main.RECON_e.xptar = a.dxdz;
    main.RECON_e.z = a.y;
    main.FP_e.x = a.x_fp; main.FP_e.dx = a.dx_fp; main.FP_e.y = a.y_fp; main.FP_e.dy = a.dy_fp;
    main.FP_e.path = a.pathlen;
  } else {
    main.RECON_e.delta = main.SP_e.delta;
    main.RECON_e.yptar = main.SP_e.yptar;
    main.RECON_e.xptar = main.SP_e.xptar;
  }
  recon.e.delta = main.RECON_e.delta;
  recon.e.yptar = main.RECON_e.yptar;
  recon.e.xptar = main.RECON_e.xptar;
  recon.e.z = main.RECON_e.z;
  recon.e.P = cfg.spec_e.P * (1. + recon.e.delta / 100.);
  recon.e.E = recon.e.P;
  dx_tmp = recon.e.xptar + cfg.spec_e.off_xptar;
  dy_tmp = recon.e.yptar + cfg.spec_e.off_yptar;
  physics_angles(cfg.spec_e.theta, cfg.spec_e.phi, dx_tmp, dy_tmp, recon.e.theta, recon.e.phi);
  if (cfg.correct_Eloss) {
    double eloss_E_arm, r;
    trip_thru_target(s, 2, 0.0, recon.e.E, recon.e.theta, eloss_E_arm, r, K::Me, 4);
    recon.e.E = recon.e.E + eloss_E_arm;
  }
  recon.e.P = recon.e.E;
  return true;
}

// event.f:1056-1359 (using_tgt_field = .false.)
bool complete_recon_ev(Sim& s, Event& recon) {
  const simc_run_config& cfg = *s.cfg;
  const simc_target& targ = cfg.targ;
  const double Mh2 = s.Mh2;
  recon.Ein = cfg.Ebeam_vertex_ave - targ.Coulomb_ave;
  recon.ue.x = sin(recon.e.theta) * cos(recon.e.phi);
  recon.ue.y = sin(recon.e.theta) * sin(recon.e.phi);
  recon.ue.z = cos(recon.e.theta);
  recon.up.x = sin(recon.p.theta) * cos(recon.p.phi);
  recon.up.y = sin(recon.p.theta) * sin(recon.p.phi);
  recon.up.z = cos(recon.p.theta);
  recon.nu = recon.Ein - recon.e.E;
  recon.Q2 = 2 * recon.Ein * recon.e.E * (1 - recon.ue.z);
  recon.q = sqrt(recon.Q2 + recon.nu * recon.nu);
  recon.uq.x = -recon.e.P * recon.ue.x / recon.q;
  recon.uq.y = -recon.e.P * recon.ue.y / recon.q;
  recon.uq.z = (recon.Ein - recon.e.P * recon.ue.z) / recon.q;
  const double W2 = K::Mp * K::Mp + 2. * K::Mp * recon.nu - recon.Q2;
  recon.W = sqrt(fabs(W2)) * W2 / fabs(W2);
  recon.xbj = recon.Q2 / 2. / K::Mp / recon.nu;
  if (cfg.doing_phsp) {
    recon.p.P = cfg.spec_p.P;
    recon.p.E = sqrt(Mh2 + recon.p.P * recon.p.P);
  }
  recon.epsilon = 1. / (1. + 2. * (1 + recon.nu * recon.nu / recon.Q2) * powi(tan(recon.e.theta / 2.), 2));
  recon.theta_pq = acos(std::min(1.0, recon.up.x * recon.uq.x + recon.up.y * recon.uq.y + recon.up.z * recon.uq.z));
  const double qx = -recon.uq.y, qy = recon.uq.x, qz = recon.uq.z;
  const double px = -recon.up.y, py = recon.up.x, pz = recon.up.z;
  double dummy = sqrt((qx * qx + qy * qy) * (qx * qx + qy * qy + qz * qz));
  const double new_x_x = -qx * qz / dummy, new_x_y = -qy * qz / dummy, new_x_z = (qx * qx + qy * qy) / dummy;
  dummy = sqrt(qx * qx + qy * qy);
  const double new_y_x = qy / dummy, new_y_y = -qx / dummy, new_y_z = 0.0;
  const double p_new_x = px * new_x_x + py * new_x_y + pz * new_x_z;
  const double p_new_y = px * new_y_x + py * new_y_y + pz * new_y_z;
  if ((p_new_x * p_new_x + p_new_y * p_new_y) == 0.) recon.phi_pq = 0.0;
  else recon.phi_pq = acos(p_new_x / sqrt(p_new_x * p_new_x + p_new_y * p_new_y));
  if (p_new_y < 0.) recon.phi_pq = 2 * K::pi - recon.phi_pq;
  if (cfg.using_tgt_field)                                // :1187-1263
    poltarg_block(cfg, recon.uq, recon.up, recon.phi_pq, recon.phi_targ, recon.beta, recon.phi_s, recon.phi_c, recon.theta_tarq);
  recon.Pmx = recon.p.P * recon.up.x - recon.q * recon.uq.x;
  recon.Pmy = recon.p.P * recon.up.y - recon.q * recon.uq.y;
  recon.Pmz = recon.p.P * recon.up.z - recon.q * recon.uq.z;
  recon.Pm = sqrt(recon.Pmx * recon.Pmx + recon.Pmy * recon.Pmy + recon.Pmz * recon.Pmz);
  const double oop_x = -recon.uq.y, oop_y = recon.uq.x;
  recon.PmPar = (recon.Pmx * recon.uq.x + recon.Pmy * recon.uq.y + recon.Pmz * recon.uq.z);
  recon.PmOop = (recon.Pmx * oop_x + recon.Pmy * oop_y) / sqrt(oop_x * oop_x + oop_y * oop_y);
  recon.PmPer = sqrt(std::max(0.e0, recon.Pm * recon.Pm - recon.PmPar * recon.PmPar - recon.PmOop * recon.PmOop));
  if (cfg.doing_pion || cfg.doing_kaon || cfg.doing_delta || cfg.doing_rho || cfg.doing_semi) {
    recon.Em = recon.nu + targ.Mtar_struck - recon.p.E;
    const double mm2 = recon.Em * recon.Em - recon.Pm * recon.Pm;
    s.ntup.mm = sqrt(fabs(mm2)) * fabs(mm2) / mm2;
    const double mmA2 = powi(recon.nu + targ.M - recon.p.E, 2) - recon.Pm * recon.Pm;
    s.ntup.mmA = sqrt(fabs(mmA2)) * fabs(mmA2) / mmA2;
    s.ntup.t = recon.Q2 - Mh2 + 2 * (recon.nu * recon.p.E - recon.p.P * recon.q * cos(recon.theta_pq));
  }
  if (cfg.doing_semi || cfg.doing_rho) {   // :1336-1339
    recon.zhad = recon.p.E / recon.nu;
    recon.pt2 = recon.p.P * recon.p.P * (1.0 - powi(cos(recon.theta_pq), 2));
  }
  if (cfg.doing_hyd_elast) {
    recon.Trec = 0.0;
    recon.Em = recon.nu + targ.M - recon.p.E - recon.Trec;
  } else if (cfg.doing_deuterium || cfg.doing_heavy) {
    recon.Trec = sqrt(recon.Pm * recon.Pm + targ.Mrec * targ.Mrec) - targ.Mrec;
    recon.Em = recon.nu + targ.Mtar_struck - recon.p.E - recon.Trec;
  } else if (cfg.doing_pion || cfg.doing_kaon || cfg.doing_delta || cfg.doing_rho) {
    recon.Em = recon.nu + targ.Mtar_struck - recon.p.E;
  }
  return true;
}

// event.f:1402-1428: linear interpolation of rho_i(Pm), Lorentzian in Em for A > 2
double theory_sf_weight(const simc_run_config& cfg, const TheoryTable& T, double Em, double Pm) {
  double SF_weight = 0.0;
  for (int i = 0; i < T.nrhoPm; ++i) {
    double weight = 0.0;
    const double r = (Pm - T.pm_min[i]) / T.pm_bin[i];
    if (r >= 0 && r <= T.n[i]) {
      int iPm1 = (int)std::lround(r);          // nint
      if (iPm1 == 0) iPm1 = 1;
      if (iPm1 == T.n[i]) iPm1 = T.n[i] - 1;
      const double frac = r + 0.5 - (double)iPm1;
      const double b = T.rho[i][iPm1 - 1];
      const double a = T.rho[i][iPm1] - b;
      weight = a * frac + b;
    }
    if (cfg.doing_heavy) {
      const double width = T.Emsig[i] / 2.0;
      if (Em < T.E_Fermi) weight = 0.0;
      weight = weight / K::pi / T.Em_int[i] * width / (powi(Em - T.Em[i], 2) + width * width);
    }
    SF_weight = SF_weight + weight * T.nprot[i];
  }
  return SF_weight;
}

// event.f:1363-1569
bool complete_main(Sim& s, bool force_sigcc, EventMain& main, Event& vertex, Event& recon) {
  const simc_run_config& cfg = *s.cfg;
  if (cfg.doing_hyd_elast || cfg.doing_pion || cfg.doing_kaon || cfg.doing_delta || cfg.doing_phsp || cfg.doing_rho ||
      cfg.doing_semi) {
    main.SF_weight = 1.0;
  } else if (cfg.use_benhar_sf && cfg.doing_heavy) {
    if (!s.sf) throw std::runtime_error("oracle: spectral-function table not set");
    const double weight = sf_lookup_diff(*s.sf, vertex.Em, vertex.Pm);
    main.SF_weight = cfg.targ.Z * cfg.transparency * weight;
  } else if (cfg.doing_deuterium || (cfg.doing_heavy && !cfg.use_benhar_sf)) {
    if (!s.theory) throw std::runtime_error("oracle: theory table not set");
    main.SF_weight = theory_sf_weight(cfg, *s.theory, vertex.Em, vertex.Pm);
  } else {
    throw std::runtime_error("oracle: no spectral-function weight for this reaction");
  }
  if (main.SF_weight <= 0 && !force_sigcc) return false;
  double tgtweight = 1.0, survivalprob = 1.0;
  if (cfg.doing_phsp) {
    main.sigcc = 1.0;
    main.sigcc_recon = 1.0;
  } else if (cfg.doing_hyd_elast) {
    main.sigcc = sigep(vertex);
    main.sigcc_recon = sigep(recon);
  } else if (cfg.doing_deuterium || cfg.doing_heavy) {
    main.sigcc = deForest(cfg, vertex);
    main.sigcc_recon = deForest(cfg, recon);
  } else if (cfg.doing_pion) {
    main.sigcc = peepi(s, vertex, main);
    if (cfg.which_pion == 2) {                    // :1464-1476
      if (cfg.doing_hydpi) main.sigcc = cfg.doing_pizero ? 0.55 * main.sigcc : 0.4 * main.sigcc;
      else if (cfg.doing_deutpi) main.sigcc = cfg.doing_pizero ? 0.55 * main.sigcc : 0.4 * main.sigcc + 0.8 * main.sigcc;
    } else if (cfg.which_pion == 3) {             // :1477-1491
      if (cfg.doing_hydpi) main.sigcc = cfg.doing_pizero ? 0 : 0.55 * main.sigcc;
      else if (cfg.doing_deutpi) main.sigcc = cfg.doing_pizero ? 0.99 * main.sigcc : 0.55 * main.sigcc + 0.99 * main.sigcc;
    }
    main.sigcc_recon = 1.0;
    if (cfg.which_pion == 1 || cfg.which_pion == 11) tgtweight = cfg.targ.N;
    else tgtweight = cfg.targ.Z;
  } else if (cfg.doing_delta) {                 // :1511-1513
    main.sigcc = peedelta(s, vertex, main);
    main.sigcc_recon = 1.0;
  } else if (cfg.doing_rho) {                   // :1515-1518
    main.sigcc = peerho(s, vertex, main);
    main.sigcc_recon = 1.0;
    tgtweight = cfg.targ.Z + cfg.targ.N;
  } else if (cfg.doing_kaon) {
    main.sigcc = peeK(s, vertex, main, survivalprob);
    main.sigcc_recon = 1.0;
    if (cfg.which_kaon == 2 || cfg.which_kaon == 12) tgtweight = cfg.targ.N;
    else tgtweight = cfg.targ.Z;
  } else if (cfg.doing_semi) {
    // NB peepiX reads vertex%theta_pq (semi_physics.f:243), which nothing ever assigns (complete_ev fills
    // main%theta_pq, event.f:727): it is zero, so the reference's jacobian has cos(theta_pq) = 1.  Reproduced.
    main.sigcc = peepiX(s, vertex, main, survivalprob);
    main.sigcc_recon = 1.0;
  } else {
    throw std::runtime_error("oracle: cross section of this reaction not restated yet");
  }
  if (cfg.using_Coulomb) main.sigcc = main.sigcc * powi(1.0 + cfg.targ.Coulomb_ave / cfg.Ebeam, 2);
  main.weight = main.SF_weight * main.jacobian * main.gen_weight * main.sigcc;
  main.weight = main.weight * tgtweight;
  if ((cfg.doing_kaon || cfg.doing_semika) && !cfg.doing_decay) main.weight = main.weight * survivalprob;
  s.ntup.survivalprob = survivalprob;
  return true;
}

// loop body, simc.f:169-246, first half: generate + montecarlo
bool try_until_recon(Sim& s, EventMain& main, Event& vertex, Event& orig, Event& recon, TryResult& r) {
  const simc_run_config& cfg = *s.cfg;
  s.trk.rng = s.rng;
  s.trk.ctau = cfg.ctau;
  s.trk.Mh2_final = cfg.Mh2;
  s.Mh = cfg.Mh; s.Mh2 = cfg.Mh2;
  s.trk.decdist = 0.0;
  s.ntup = NtupVars();
  s.stop_e = -1; s.stop_p = -1; s.hut_e = s.hut_p = false;
  bool success = generate(s, main, vertex, orig);
  r.gen_success = success;
  r.stage = 0;
  s.field_fail_p = false;
  if (success) { success = montecarlo(s, orig, main, recon); r.stage = success ? 3 : ((s.stop_p > 0 || s.field_fail_p) ? 1 : 2); }
  return success;
}

// second half: complete_recon_ev, complete_main, pass_cuts, hard cuts (simc.f:219-246)
void finish_try(Sim& s, EventMain& main, Event& vertex, Event& recon, bool success, TryResult& r) {
  const simc_run_config& cfg = *s.cfg;
  if (success) success = complete_recon_ev(s, recon);
  if (success) success = complete_main(s, false, main, vertex, recon);
  // NB the p-arm upper delta edge uses SPedge%e%delta%max (simc.f:237)
  r.pass_cuts = !(recon.e.delta <= (cfg.SPedge_e.delta.min + cfg.slop_MC_e_used[0]) ||
                  recon.e.delta >= (cfg.SPedge_e.delta.max - cfg.slop_MC_e_used[0]) ||
                  recon.e.yptar <= (cfg.SPedge_e.yptar.min + cfg.slop_MC_e_used[1]) ||
                  recon.e.yptar >= (cfg.SPedge_e.yptar.max - cfg.slop_MC_e_used[1]) ||
                  recon.e.xptar <= (cfg.SPedge_e.xptar.min + cfg.slop_MC_e_used[2]) ||
                  recon.e.xptar >= (cfg.SPedge_e.xptar.max - cfg.slop_MC_e_used[2]) ||
                  recon.p.delta <= (cfg.SPedge_p.delta.min + cfg.slop_MC_p_used[0]) ||
                  recon.p.delta >= (cfg.SPedge_e.delta.max - cfg.slop_MC_p_used[0]) ||
                  recon.p.yptar <= (cfg.SPedge_p.yptar.min + cfg.slop_MC_p_used[1]) ||
                  recon.p.yptar >= (cfg.SPedge_p.yptar.max - cfg.slop_MC_p_used[1]) ||
                  recon.p.xptar <= (cfg.SPedge_p.xptar.min + cfg.slop_MC_p_used[2]) ||
                  recon.p.xptar >= (cfg.SPedge_p.xptar.max - cfg.slop_MC_p_used[2]));
  if (cfg.hard_cuts) {
    if (!r.pass_cuts) success = false;
    if (cfg.doing_eep && (recon.Em > cfg.cuts_Em.max)) success = false;
  }
  r.success = success;
  if (success) r.stage = 4;
}

TryResult one_try(Sim& s, EventMain& main, Event& vertex, Event& orig, Event& recon) {
  TryResult r;
  const bool success = try_until_recon(s, main, vertex, orig, recon, r);
  finish_try(s, main, vertex, recon, success, r);
  return r;
}

}  // namespace simc_oracle
