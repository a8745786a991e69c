// Event generation, vertex / reconstructed kinematics and weights on the device.
// Replaces generate + complete_ev (event.f:126-1052), generate_rad (radc.f:120-519),
// the target-to-spectrometer part of montecarlo (simc.f:1365-1443, 1623-1645, 1655-1846),
// complete_recon_ev (event.f:1056-1359), complete_main (event.f:1363-1569) and sigep
// (physics_proton.f:1-190).  Reaction coverage of this build: H(e,e'p), H(e,e'pi+-) and H(e,e'K+);
// the host refuses the other reaction flags at create().
#pragma once
#include "radc.cuh"
#include "physics_meson.cuh"
#include "physics_heavy.cuh"
#include "physics_semi.cuh"

namespace simc {

#define SIMC_PI_D 3.141592653589793
#define SIMC_ME 0.51099906
#define SIMC_MP 938.27231
#define SIMC_MPI 139.57018
#define SIMC_MRHO 769.3
#define SIMC_HBARC 197.327053
#define SIMC_ALPHA (1. / 137.0359895)

// event.f:1572-1614
SIMC_HD_CALL void physics_angles(double theta0, double phi0, double dx, double dy, double& theta, double& phi) {
  const double costh = m::cos(theta0), sinth = m::sin(theta0), sinph = m::sin(phi0);
  const double r = sqrt(1. + dx * dx + dy * dy);
  theta = m::acos((costh - dy * sinth * sinph) / r);
  if (dx != 0.0) {
    phi = m::atan((dy * costh + sinth * sinph) / dx);
    if (phi <= 0) phi = phi + SIMC_PI_D;
    if (sinph < 0.) phi = phi + SIMC_PI_D;
  } else {
    phi = phi0;
  }
}

// event.f:1618-1648
SIMC_HD_CALL void spectrometer_angles(double theta0, double phi0, double& dx, double& dy, double theta, double phi) {
  const double x = m::sin(theta) * m::cos(phi), y = m::sin(theta) * m::sin(phi), z = m::cos(theta);
  const double x0 = m::sin(theta0) * m::cos(phi0), y0 = m::sin(theta0) * m::sin(phi0), z0 = m::cos(theta0);
  const double cos_dtheta = x * x0 + y * y0 + z * z0;
  dx = x / cos_dtheta;
  dy = sqrt(1 / (cos_dtheta * cos_dtheta) - 1. - dx * dx);
  const double y_event = y / cos_dtheta;
  if (y_event < y0) dy = -dy;
}

// sigep, fofa_best_fit, sigMott: physics_proton.f:1-22,137-190
SIMC_HD_CALL double sigep(double Ein, double eE, double etheta, double Q2v) {
  const double mu_p = 2.793;
  const double qsquar = -Q2v / (SIMC_HBARC * SIMC_HBARC);
  const double Q2 = -qsquar * (SIMC_HBARC * SIMC_HBARC) * 1.e-6;   // hbarc**2. : m::pow(x,2.) == x*x
  const double Q = sqrt(fmax(Q2, 0.e0));
  const double Q3 = m::pow(Q, 3.), Q4 = m::pow(Q, 4.), Q5 = m::pow(Q, 5.);
  double denom = 1. + 0.62 * Q + 0.68 * Q2 + 2.8 * Q3 + 0.83 * Q4;
  const double GE = 1. / denom;
  denom = 1. + 0.35 * Q + 2.44 * Q2 + 0.5 * Q3 + 1.04 * Q4 + 0.34 * Q5;
  const double GM = mu_p / denom;
  const double qmu4mp = Q2v / 4. / (SIMC_MP * SIMC_MP);
  const double W1p = GM * GM * qmu4mp;
  const double W2p = (GE * GE + GM * GM * qmu4mp) / (1.0 + qmu4mp);
  const double th2 = m::tan(etheta / 2.);
  const double Wp = W2p + 2. * W1p * (th2 * th2);
  const double mott = 2. * SIMC_ALPHA * SIMC_HBARC * eE * m::cos(etheta / 2.) / Q2v;
  const double sigMott = (mott * mott) * 1.e4;
  return sigMott * eE / Ein * Wp;
}

// Azimuthal angles of a polarised target (using_tgt_field): event.f:796-895 for the vertex, :1187-1263 for the
// reconstructed event -- the same construction from the unit vectors of q and of the hadron and phi_pq.  The target
// polarisation lies in the horizontal plane at |targ_Bangle| to the beam ("replay" coordinates: x right, y down).
struct PolTargAngles { double phi_targ, beta, phi_s, phi_c, theta_tarq; };
SIMC_HD_CALL PolTargAngles poltarg_angles(double targ_pol, double targ_Bangle, double uqx, double uqy, double uqz, double upx,
                                          double upy, double upz, double phi_pq) {
  PolTargAngles A;
  const double qx = -uqy, qy = uqx, qz = uqz;
  const double targx = -targ_pol * m::sin(fabs(targ_Bangle)), targy = 0.0, targz = targ_pol * m::cos(fabs(targ_Bangle));
  double dummy = sqrt((qx * qx + qy * qy) * (qx * qx + qy * qy + qz * qz));
  double new_x_x = -qx * qz / dummy, new_x_y = -qy * qz / dummy, new_x_z = (qx * qx + qy * qy) / dummy;
  dummy = sqrt(qx * qx + qy * qy);
  double new_y_x = qy / dummy, new_y_y = -qx / dummy, new_y_z = 0.0;
  const double p_new_x = targx * new_x_x + targy * new_x_y + targz * new_x_z;
  const double p_new_y = targx * new_y_x + targy * new_y_y + targz * new_y_z;
  A.phi_targ = m::atan2(p_new_y, p_new_x);
  if (A.phi_targ < 0.) A.phi_targ = 2. * SIMC_PI_D + A.phi_targ;
  const double px = -upy, py = upx, pz = upz;
  dummy = sqrt((qy * pz - qz * py) * (qy * pz - qz * py) + (qz * px - qx * pz) * (qz * px - qx * pz) +
               (qx * py - qy * px) * (qx * py - qy * px));
  new_y_x = (qy * pz - qz * py) / dummy;
  new_y_y = (qz * px - qx * pz) / dummy;
  new_y_z = (qx * py - qy * px) / dummy;
  dummy = sqrt((new_y_y * qz - new_y_z * qy) * (new_y_y * qz - new_y_z * qy) + (new_y_z * qx - new_y_x * qz) * (new_y_z * qx - new_y_x * qz) +
               (new_y_x * qy - new_y_y * qx) * (new_y_x * qy - new_y_y * qx));
  new_x_x = (new_y_y * qz - new_y_z * qy) / dummy;
  new_x_y = (new_y_z * qx - new_y_x * qz) / dummy;
  new_x_z = (new_y_x * qy - new_y_y * qx) / dummy;
  const double targ_new_x = targx * new_x_x + targy * new_x_y + targz * new_x_z;
  const double targ_new_y = targx * new_y_x + targy * new_y_y + targz * new_y_z;
  A.beta = m::atan2(targ_new_y, targ_new_x);
  if (A.beta < 0.) A.beta = 2 * SIMC_PI_D + A.beta;
  A.phi_s = phi_pq - A.phi_targ;
  if (A.phi_s < 0.) A.phi_s = 2 * SIMC_PI_D + A.phi_s;
  A.phi_c = phi_pq + A.phi_targ;
  if (A.phi_c > 2. * SIMC_PI_D) A.phi_c = A.phi_c - 2 * SIMC_PI_D;
  if (A.phi_c < 0.0) A.phi_c = 2 * SIMC_PI_D + A.phi_c;
  dummy = sqrt((qx * qx + qy * qy + qz * qz)) * sqrt((targx * targx + targy * targy + targz * targz));
  A.theta_tarq = m::acos((qx * targx + qy * targy + qz * targz) / dummy);
  return A;
}

// State of one try between the stages of the loop: the parts of `main`, `vertex`, `orig`
// and /radccom/ that later stages read.
struct EventState {
  // main%target
  double tx, ty, tz, rastery, rasterx, Eloss[3], teff[3], Coulomb;
  // main
  double gen_weight, jacobian, Ein_shift, Ee_shift, Trec;
  // vertex
  double v_Ein, v_eE, v_edelta, v_eyptar, v_exptar, v_etheta, v_ephi;
  double v_pE, v_pP, v_pdelta, v_pyptar, v_pxptar, v_ptheta, v_pphi;
  double v_Q2, v_Em, v_Pm, v_Trec;
  double uex, uey, uez, upx, upy, upz;
  // meson production only: vertex%nu, q, uq and main%epsilon, theta_pq, phi_pq, t, tmin, W
  double v_nu, v_q, uqx, uqy, uqz, m_eps, m_thpq, m_phipq, m_t, m_tmin, m_W;
  // semi-inclusive production: vertex%zhad, pt2 and COMMON /pfermi_stuff/ (simulate.inc:212-217)
  double v_zhad, v_pt2, pfer, pferx, pfery, pferz, efer;
  // orig (fields that differ from vertex)
  double o_Ein, o_eE, o_edelta, o_pE, o_pP, o_pdelta;
  // rho production only: the decay pion's angles in `orig` (rho_decay.f:142-146) and ntup%rhomass, ntup%rhotheta.
  // With doing_pizero the last two hold rph and rth of pizero_decay (pizero_decay.f:43-45) instead: k_calo rebuilds
  // the two photons from them.
  double o_pyptar, o_pxptar, rho_mass, rho_theta;
  RadEvDev rad;
};

// Mass of the particle that enters the hadron arm and is reconstructed: COMMON Mh after generate.  A run constant for
// every reaction but rho production, where rho_decay leaves the pion mass behind (rho_decay.f:150-153).
SIMC_HD double detected_Mh(const simc_run_config& cfg) { return cfg.doing_rho ? SIMC_MPI : cfg.Mh; }
SIMC_HD double detected_Mh2(const simc_run_config& cfg) { return cfg.doing_rho ? SIMC_MPI * SIMC_MPI : cfg.Mh2; }

// trip_thru_target with typeflag = 1 (sampled energy loss): one |gauss1(10)| per material with
// thick > 0, in the reference's order target, Al, air, kevlar, mylar (target.f:46-52,170-180).
template <class RNG, class GAUSS>
SIMC_HD_CALL void trip_thru_target_sampled(const simc_run_config& cfg, const MatTable& mt, RNG& rng, GAUSS gauss, int narm,
                                      double zpos, double energy, double theta, double mass, double& Eloss,
                                      double& radlen) {
  const simc_target& targ = cfg.targ;
  const Material al = SIMC_MAT_AL;
  const ParticleKin k = particle_kin(energy, mass, mt.targ.ln10);
  double s_target, s_Al;
  auto one = [&](double len, const MatConst& mc) {
    double x = 0.;
    if (len * mc.rho > 0.) x = fabs(gauss(rng, 10.0));
    return enerloss_material(k, len, mc, x);
  };
  if (narm == 1) {
    incoming_paths(targ, zpos, s_target, s_Al);
    radlen = s_target / targ.X0_cm + s_Al / al.X0_cm;
    const double e1 = one(s_target, mt.targ);
    const double e2 = one(s_Al, mt.al);
    Eloss = e1 + e2;
    return;
  }
  const Material air = SIMC_MAT_AIR, kev = SIMC_MAT_KEVLAR, myl = SIMC_MAT_MYLAR;
  const ArmWindows w = arm_windows(narm == 2 ? cfg.electron_arm : cfg.hadron_arm);
  outgoing_paths(targ, w, zpos, theta, s_target, s_Al);
  radlen = s_target / targ.X0_cm + s_Al / al.X0_cm + w.s_air / air.X0_cm + w.s_kevlar / kev.X0_cm +
           w.s_mylar / myl.X0_cm;
  const double e1 = one(s_target, mt.targ);
  const double e2 = one(s_Al, mt.al);
  const double e3 = one(w.s_air, mt.air);
  const double e4 = one(w.s_kevlar, mt.kevlar);
  const double e5 = one(w.s_mylar, mt.mylar);
  Eloss = e1 + e2 + e3 + e4 + e5;
}

// The generation code is one long straight line (no hot loop): if the warps of an SM drift apart
// in it, each of them streams >100 KB of SASS through the instruction cache on its own and the
// kernel stalls on instruction fetch (profiles/r1_notes.md).  SIMC_PHASE() re-aligns the warps of
// the CTA between phases, so every thread of the CTA MUST walk through generate_hyd_elast /
// complete_ev_hyd_elast, whether it has a live event or not (`run` guards the work).
#if defined(__CUDA_ARCH__)
#define SIMC_PHASE() __syncthreads()
#else
#define SIMC_PHASE() ((void)0)
#endif

// complete_ev for H(e,e'p), event.f:432-1052.  Needs v_Ein, v_eyptar/xptar/theta/phi, tz; fills
// the rest of the vertex, the jacobian, Eloss/teff(2:3) and the radiative constants.
template <class RNG, class GAUSS>
SIMC_HD bool complete_ev_hyd_elast(const simc_run_config& cfg, const MatTable& mt, RNG& rng, GAUSS gauss, EventState& s, bool run) {
  const double Mh = cfg.Mh, Mh2 = cfg.Mh2;
  if (run) {
    s.jacobian = 1.0;
    s.uex = m::sin(s.v_etheta) * m::cos(s.v_ephi);
    s.uey = m::sin(s.v_etheta) * m::sin(s.v_ephi);
    s.uez = m::cos(s.v_etheta);
    s.v_eE = s.v_Ein * Mh / (Mh + s.v_Ein * (1. - s.uez));
    if (s.v_eE > s.v_Ein) run = false;
  }
  if (run) {
    const double eP = s.v_eE;
    s.v_edelta = (eP - cfg.spec_e.P) * 100. / cfg.spec_e.P;
    const double nu = s.v_Ein - s.v_eE;
    s.v_Q2 = 2 * s.v_Ein * s.v_eE * (1. - s.uez);
    const double q = sqrt(s.v_Q2 + nu * nu);
    const double uqx = -eP * s.uex / q;
    const double uqy = -eP * s.uey / q;
    const double uqz = (s.v_Ein - eP * s.uez) / q;
    // |uq|^2-1 > 0.01 is a fatal `stop` in the reference (event.f:538); it cannot trigger here
    s.v_Em = 0.0;
    s.v_Pm = 0.0;
    s.upx = uqx; s.upy = uqy; s.upz = uqz;
    s.v_pP = q;
    s.v_ptheta = m::acos(s.upz);
    s.v_pphi = m::atan2(s.upy, s.upx);
    if (s.v_pphi < 0.) s.v_pphi = s.v_pphi + 2. * SIMC_PI_D;
    spectrometer_angles(cfg.spec_p.theta, cfg.spec_p.phi, s.v_pxptar, s.v_pyptar, s.v_ptheta, s.v_pphi);
    s.v_pE = sqrt(s.v_pP * s.v_pP + Mh2);
    s.v_pdelta = (s.v_pP - cfg.spec_p.P) * 100. / cfg.spec_p.P;
    s.v_Trec = 0.0;
    const double r = sqrt(1. + s.v_eyptar * s.v_eyptar + s.v_exptar * s.v_exptar);
    s.jacobian = s.jacobian / (r * (r * r));
  }
  const double zpos = s.tz - cfg.targ.zoffset;
  SIMC_PHASE();
  if (run) trip_thru_target_sampled(cfg, mt, rng, gauss, 2, zpos, s.v_eE, s.v_etheta, SIMC_ME, s.Eloss[1], s.teff[1]);
  SIMC_PHASE();
  if (run) trip_thru_target_sampled(cfg, mt, rng, gauss, 3, zpos, s.v_pE, s.v_ptheta, Mh, s.Eloss[2], s.teff[2]);
  SIMC_PHASE();
  if (run) {
    if (!cfg.using_Eloss) { s.Eloss[1] = 0.0; s.Eloss[2] = 0.0; }
    VertexKin v;
    v.Ein = s.v_Ein; v.eE = s.v_eE; v.eP = s.v_eE; v.etheta = s.v_etheta; v.pE = s.v_pE; v.pP = s.v_pP;
    v.uex = s.uex; v.uey = s.uey; v.uez = s.uez; v.upx = s.upx; v.upy = s.upy; v.upz = s.upz;
    radc_init_ev(cfg, v, s.teff[0], s.teff[1], s.rad);
  }
  SIMC_PHASE();
  return run;
}

// What generate_rad leaves for later: which tail radiates (0: none), the photon-energy limits and the
// basicrad weight.  peaked_rad_weight (radc.f:523-646) draws no random number and only scales
// main%gen_weight, so it is evaluated for the tries that survive both spectrometers (k_radw, loop.cuh).
struct GenRad { double emin, emax, eg, bw; int which; };

// orig = vertex (+) radiation, radc.f:476-515.  main%gen_weight keeps its pre-radiation value here:
// gen_weight * rad_weight / hardcorfac (radc.f:518) is applied by k_radw.
SIMC_HD bool generate_finalize(const simc_run_config& cfg, EventState& s, bool ok) {
  if (!cfg.using_rad) {
    if (ok) {
      s.o_Ein = s.v_Ein; s.o_eE = s.v_eE; s.o_edelta = s.v_edelta; s.o_pE = s.v_pE; s.o_pP = s.v_pP;
      s.o_pdelta = s.v_pdelta;
    }
    return ok;
  }
  const RadEvDev& R = s.rad;
  const double Mh = cfg.doing_rho ? s.rho_mass : cfg.Mh;               // COMMON Mh: generate_rho's for rho production
  const double Mh2 = cfg.doing_rho ? s.rho_mass * s.rho_mass : cfg.Mh2;
  if (ok) {
    s.o_Ein = s.v_Ein + R.Egamma_used[0];
    s.o_eE = s.v_eE - R.Egamma_used[1];
    if (s.o_eE <= 0e0) ok = false;
  }
  if (ok) {
    s.o_edelta = (s.o_eE - cfg.spec_e.P) / cfg.spec_e.P * 100.;
    s.o_pE = s.v_pE - R.Egamma_used[2];
    if (s.o_pE <= Mh) ok = false;
  }
  if (ok) {
    s.o_pP = sqrt(s.o_pE * s.o_pE - Mh2);
    s.o_pdelta = (s.o_pP - cfg.spec_p.P) / cfg.spec_p.P * 100.;
  }
  return ok;
}

// ---- generate_rad in the (Egamma1, Egamma2, Egamma3) basis, radc.f:120-519 with rad_flag = 2 (one tail, chosen with
// equal probability) or 3 (all tails radiate).  Each tail draws its photon energy from BASICRAD with that tail's own
// (g, c) and the weight is the product of the BASICRAD weights and extrad_phi(itail).  Cut where the reference
// re-enters complete_ev (radc.f:324): *_first ends with tail 1, *_rest does the tails behind it (in k_regen for the
// tries whose tail 1 was active, at once for the others).  gr.which = 4 marks the finished weight in gr.bw.
template <class RNG>
SIMC_HD_CALL bool generate_rad_basis_rest(const simc_run_config& cfg, RNG& rng, EventState& s, GenRad& gr, bool after_tail1) {
  RadEvDev& R = s.rad;
  const int ntail = R.ntail;
  const simc_edge& VE = cfg.VERTEXedge;
  double rad_weight = 1;
  if (after_tail1) {
    rad_weight = rad_weight * gr.bw;
    rad_weight = rad_weight * extrad_phi(cfg, R, 1, s.v_Ein, s.v_eE, R.Egamma_used[0]);
  } else if (cfg.doing_heavy) {     // radc.f:340-350 (the second pass makes this test itself)
    if (s.v_Em < VE.Em.min || s.v_Em > VE.Em.max || s.v_Pm < VE.Pm.min || s.v_Pm > VE.Pm.max) return false;
  }
  // max_delta_Trec is formed before tail 1 from the vertex%Trec of the first pass, which main%Trec still holds
  const double max_delta_Trec = fmax((s.Trec - VE.Trec.min), (VE.Trec.max - s.Trec));
  const bool eep = cfg.doing_eep != 0;
  const bool tail2 = cfg.doing_tail[1] && (ntail == 0 || ntail == 2);
  const bool tail3 = R.rad_proton_this_ev && (ntail == 0 || ntail == 3);
  BasisConst B;
  if (tail2 || tail3) B = basis_constants(R, s.v_Ein, s.v_eE, s.v_pE);
  if (tail2) {                      // radc.f:354-404
    double emin = s.v_eE - cfg.edge.e.E.max;
    double emax = s.v_eE - cfg.edge.e.E.min;
    if (eep) {
      emax = fmin(emax, (cfg.edge.Em.max - s.v_Em) - R.Egamma_used[0] + max_delta_Trec);
      if (ntail != 0 || !R.rad_proton_this_ev) emin = fmax(emin, (cfg.edge.Em.min - s.v_Em) - R.Egamma_used[0] - max_delta_Trec);
    }
    emax = fmin(emax, cfg.Egamma_tot_max - R.Egamma_used[0]);
    emin = emin - cfg.dE_edge_test;
    emax = emax + cfg.dE_edge_test;
    if (cfg.hardwired_rad) emax = cfg.Egamma_gen_max;
    double eg, bw;
    basicrad_tail(R.g[2], B.c[2], rng, emin, emax, eg, bw);
    if (bw <= 0) return false;
    R.Egamma_used[1] = eg;
    rad_weight = rad_weight * bw;
    rad_weight = rad_weight * extrad_phi(cfg, R, 2, s.v_Ein, s.v_eE, eg);
  }
  if (tail3) {                      // radc.f:408-455
    double emin = s.v_pE - cfg.edge.p.E.max;
    double emax = s.v_pE - cfg.edge.p.E.min;
    if (eep) {
      emax = fmin(emax, (cfg.edge.Em.max - s.v_Em) - R.Egamma_used[0] - R.Egamma_used[1] + max_delta_Trec);
      emin = fmax(emin, (cfg.edge.Em.min - s.v_Em) - R.Egamma_used[0] - R.Egamma_used[1] - max_delta_Trec);
    }
    emax = fmin(emax, cfg.Egamma_tot_max - R.Egamma_used[0] - R.Egamma_used[1]);
    emin = emin - cfg.dE_edge_test;
    emax = emax + cfg.dE_edge_test;
    if (cfg.hardwired_rad) emax = cfg.Egamma_gen_max;
    double eg, bw;
    basicrad_tail(R.g[3], B.c[3], rng, emin, emax, eg, bw);
    if (bw <= 0) return false;
    R.Egamma_used[2] = eg;
    rad_weight = rad_weight * bw;
    rad_weight = rad_weight * extrad_phi(cfg, R, 3, s.v_Ein, s.v_eE, eg);
  }
  gr.which = 4; gr.bw = rad_weight;
  return true;
}
template <class RNG>
SIMC_HD_CALL bool generate_rad_basis_first(const simc_run_config& cfg, RNG& rng, EventState& s, GenRad& gr) {
  RadEvDev& R = s.rad;
  const simc_edge& VE = cfg.VERTEXedge;
  if (cfg.rad_flag == 2) {          // radc.f:206-208
    R.ntail = (int)(rng.uniform() * 3.) + 1;
    if (R.ntail == 4) R.ntail = 3;
  } else {
    R.ntail = 0;
  }
  const int ntail = R.ntail;
  if (cfg.doing_tail[0] && (ntail == 0 || ntail == 1)) {
    const double max_delta_Trec = fmax((s.v_Trec - VE.Trec.min), (VE.Trec.max - s.v_Trec));
    double emin = 0.0, emax = 0.0;
    if (cfg.doing_heavy) {          // radc.f:249-255
      emin = s.v_Em - VE.Em.max - max_delta_Trec;
      emax = s.v_Em - VE.Em.min + max_delta_Trec;
      if (ntail != 0) emax = fmin(emax, s.v_Em - cfg.edge.Em.min + max_delta_Trec);
    } else if (cfg.doing_hyd_elast) {   // radc.f:264-271
      double ebeam_max = SIMC_MP * cfg.edge.e.E.max / (SIMC_MP - cfg.edge.e.E.max * (1. - s.uez));
      if (ebeam_max < 0) ebeam_max = 1.e10;
      const double ebeam_min = SIMC_MP * cfg.edge.e.E.min / (SIMC_MP - cfg.edge.e.E.min * (1. - s.uez));
      emin = s.v_Ein - ebeam_max;
      emax = s.v_Ein - ebeam_min;
      emax = fmin(emax, cfg.edge.Em.max);
    } else if (cfg.doing_deuterium) {   // radc.f:275-277
      emax = fmin(cfg.Egamma1_max, cfg.gen.sumEgen.max - s.v_eE);
      // ntail = 0: Egamma_min(1) keeps what the previous event left in COMMON /radccom/, which starts at zero and only
      // decreases (dE_edge_test is subtracted every event); any value <= 0 gives the same BASICRAD: zero stands for it
      emin = ntail != 0 ? cfg.gen.sumEgen.min - s.v_eE : 0.0;
    } else {                        // pion / kaon / semi-inclusive, radc.f:283-291
      emin = 0.;
      emax = cfg.gen.sumEgen.max - s.v_eE;
    }
    emax = fmin(emax, cfg.Egamma1_max);
    emin = emin - cfg.dE_edge_test;
    emax = emax + cfg.dE_edge_test;
    if (cfg.hardwired_rad) emax = cfg.Egamma_gen_max;
    const BasisConst B = basis_constants(R, s.v_Ein, s.v_eE, s.v_pE);
    double eg, bw;
    basicrad_tail(R.g[1], B.c[1], rng, emin, emax, eg, bw);
    if (bw <= 0) return false;
    R.Egamma_used[0] = eg;
    s.v_Ein = s.v_Ein - eg;
    gr.emin = emin; gr.emax = emax; gr.eg = eg; gr.bw = bw; gr.which = 1;      // on to complete_ev (k_regen)
    return true;
  }
  return generate_rad_basis_rest(cfg, rng, s, gr, false);
}

// generate + generate_rad for H(e,e'p): event.f:126-428, radc.f:120-519.  `ok` in: the thread has
// a try to generate; returns success.  The work is cut where generate_rad re-enters complete_ev
// (radc.f:324, the tries whose incoming electron radiates): *_first runs up to the photon energy of the
// chosen tail; tries with gr.which == 1 then go through *_second on compacted warps (k_regen), the others
// are complete.  Every try consumes its random numbers in the reference's order.
template <class RNG, class GAUSS>
SIMC_HD bool generate_hyd_elast_first(const simc_run_config& cfg, const MatTable& mt, RNG& rng, GAUSS gauss, EventState& s, bool ok,
                                      GenRad& gr) {
  const simc_target& targ = cfg.targ;
  if (ok) {
    s.tx = gauss(rng, 3.0) * cfg.gen.xwid + targ.xoffset;
    s.ty = gauss(rng, 3.0) * cfg.gen.ywid + targ.yoffset;
    double t3, t4, t5, t6;
    if (targ.fr_pattern == 1) {
      t3 = rng.uniform() * SIMC_PI_D;
      t4 = rng.uniform() * SIMC_PI_D;
      t5 = m::cos(t3) * targ.fr1;
      t6 = m::cos(t4) * targ.fr2;
    } else if (targ.fr_pattern == 2) {
      t3 = rng.uniform() * 2. * SIMC_PI_D;
      t4 = sqrt(rng.uniform()) * (targ.fr2 - targ.fr1) + targ.fr1;
      t5 = m::cos(t3) * t4;
      t6 = m::sin(t3) * t4;
    } else if (targ.fr_pattern == 3) {
      t3 = 2. * rng.uniform() - 1.0;
      t4 = 2. * rng.uniform() - 1.0;
      t5 = targ.fr1 * t3;
      t6 = targ.fr2 * t4;
    } else {
      t5 = 0.0; t6 = 0.0;
    }
    s.tx = s.tx + t5;
    s.ty = s.ty + t6;
    s.tz = (0.5 - rng.uniform()) * targ.length + targ.zoffset;
    s.rastery = t6;
    s.rasterx = t5;
    trip_thru_target_sampled(cfg, mt, rng, gauss, 1, s.tz - targ.zoffset, cfg.Ebeam, 0.0, SIMC_ME, s.Eloss[0], s.teff[0]);
    if (!cfg.using_Eloss) s.Eloss[0] = 0.0;
    s.Coulomb = cfg.using_Coulomb ? targ.Coulomb_constant : 0.0;
    s.v_Ein = cfg.Ebeam + (rng.uniform() - 0.5) * cfg.dEbeam + s.Coulomb - s.Eloss[0];
    s.Ein_shift = s.v_Ein - cfg.Ebeam_vertex_ave;
    s.Ee_shift = s.Coulomb - targ.Coulomb_ave;
    s.gen_weight = 1.0;
    s.v_eyptar = cfg.gen.e.yptar.min + rng.uniform() * (cfg.gen.e.yptar.max - cfg.gen.e.yptar.min);
    s.v_exptar = cfg.gen.e.xptar.min + rng.uniform() * (cfg.gen.e.xptar.max - cfg.gen.e.xptar.min);
    physics_angles(cfg.spec_e.theta, cfg.spec_e.phi, s.v_exptar, s.v_eyptar, s.v_etheta, s.v_ephi);
    // (the reference also converts the not-yet-known proton angles here, event.f:325; the result is
    //  overwritten in complete_ev before anyone reads it)
    s.v_Em = 0.0;
    s.rad.Egamma_used[0] = s.rad.Egamma_used[1] = s.rad.Egamma_used[2] = 0.0;
    s.rad.ntail = 0;
  }
  SIMC_PHASE();
  ok = complete_ev_hyd_elast(cfg, mt, rng, gauss, s, ok);
  if (ok) s.Trec = s.v_Trec;
  gr.emin = 0.0; gr.emax = 0.0; gr.eg = 0.0; gr.bw = 0.0; gr.which = 0;
  if (!cfg.using_rad) return ok;
  if (cfg.rad_flag >= 2) return ok ? generate_rad_basis_first(cfg, rng, s, gr) : false;
  // ---- generate_rad, peaked basis: exactly one tail radiates (radc.f:198-208).
  RadEvDev& R = s.rad;
  double bw = 0, emin = 0.0, emax = 0.0, eg = 0.0;
  int which = 0;
  if (ok) {
    const double x = rng.uniform();
    if (x >= R.frac[0] + R.frac[1]) R.ntail = 3;
    else if (x >= R.frac[0]) R.ntail = 2;
    else R.ntail = 1;
    const int ntail = R.ntail;
    const double max_delta_Trec = fmax((s.v_Trec - cfg.VERTEXedge.Trec.min), (cfg.VERTEXedge.Trec.max - s.v_Trec));
    if (cfg.doing_tail[0] && ntail == 1) {          // hydrogen elastic limits, radc.f:264-271
      double ebeam_max = SIMC_MP * cfg.edge.e.E.max / (SIMC_MP - cfg.edge.e.E.max * (1. - s.uez));
      if (ebeam_max < 0) ebeam_max = 1.e10;
      const double ebeam_min = SIMC_MP * cfg.edge.e.E.min / (SIMC_MP - cfg.edge.e.E.min * (1. - s.uez));
      emin = s.v_Ein - ebeam_max;
      emax = s.v_Ein - ebeam_min;
      emax = fmin(emax, cfg.edge.Em.max);
      emax = fmin(emax, cfg.Egamma1_max);
      which = 1;
    } else if (cfg.doing_tail[1] && ntail == 2) {   // radc.f:358-374
      emin = s.v_eE - cfg.edge.e.E.max;
      emax = s.v_eE - cfg.edge.e.E.min;
      emax = fmin(emax, (cfg.edge.Em.max - s.v_Em) - R.Egamma_used[0] + max_delta_Trec);
      emin = fmax(emin, (cfg.edge.Em.min - s.v_Em) - R.Egamma_used[0] - max_delta_Trec);   // ntail != 0
      emax = fmin(emax, cfg.Egamma_tot_max - R.Egamma_used[0]);
      which = 2;
    } else if (R.rad_proton_this_ev && ntail == 3) {   // radc.f:409-425
      emin = s.v_pE - cfg.edge.p.E.max;
      emax = s.v_pE - cfg.edge.p.E.min;
      emax = fmin(emax, (cfg.edge.Em.max - s.v_Em) - R.Egamma_used[0] - R.Egamma_used[1] + max_delta_Trec);
      emin = fmax(emin, (cfg.edge.Em.min - s.v_Em) - R.Egamma_used[0] - R.Egamma_used[1] - max_delta_Trec);
      emax = fmin(emax, cfg.Egamma_tot_max - R.Egamma_used[0] - R.Egamma_used[1]);
      which = 3;
    }
    if (which) {
      emin = emin - cfg.dE_edge_test;
      emax = emax + cfg.dE_edge_test;
      if (cfg.hardwired_rad) emax = cfg.Egamma_gen_max;
      basicrad4(R, rng, emin, emax, eg, bw);
      if (bw <= 0) ok = false;
      else {
        R.Egamma_used[which - 1] = eg;
        if (which == 1) s.v_Ein = s.v_Ein - eg;
      }
    }
  }
  gr.emin = emin; gr.emax = emax; gr.eg = eg; gr.bw = bw; gr.which = which;
  return ok;
}
// second pass through complete_ev for the tries whose incoming electron radiated (radc.f:324)
template <class RNG, class GAUSS>
SIMC_HD bool generate_hyd_elast_second(const simc_run_config& cfg, const MatTable& mt, RNG& rng, GAUSS gauss, EventState& s, bool run) {
  return complete_ev_hyd_elast(cfg, mt, rng, gauss, s, run);
}

// ---- H(e,e'rho0): generate_rho.f:1-131.  The rho is thrown flat in cos(theta), phi in the virtual photon - nucleon
// centre of mass with a Breit-Wigner mass and boosted to the lab: three random numbers, the rho's lab momentum,
// energy and direction, and COMMON Mh for the rest of complete_ev (s.rho_mass).  Hydrogen: the nucleon is at rest
// (pfer = 0, efer = Mtar_struck); the Fermi terms of the boost vanish identically.
template <class RNG>
SIMC_HD_CALL bool generate_rho(const simc_run_config& cfg, RNG& rng, EventState& s) {
  using namespace mesondetail;
  const double nu = s.v_nu, Q2 = s.v_Q2, q = s.v_q;
  const double bx = -(q * s.uqx + s.pferx * s.pfer) / (nu + s.efer);
  const double by = -(q * s.uqy + s.pfery * s.pfer) / (nu + s.efer);
  const double bz = -(q * s.uqz + s.pferz * s.pfer) / (nu + s.efer);
  const double betacm = sqrt(bx * bx + by * by + bz * bz);
  if (betacm > 1.0) return false;
  const double gammacm = 1. / sqrt(1.0 - betacm * betacm);
  const double ss = -Q2 + s.efer * s.efer + 2. * nu * s.efer;
  double Mh = SIMC_MRHO;
  Mh = Mh + 0.5 * 150.2 * m::tan((2. * rng.uniform() - 1.) * m::atan(2. * 500. / 150.2));
  const double Mh2 = Mh * Mh;
  s.rho_mass = Mh;
  const double Erhocm = (ss + Mh2 - cfg.targ.Mrec_struck * cfg.targ.Mrec_struck) / 2. / sqrt(ss);
  if (Erhocm < Mh) return false;
  const double Prhocm = sqrt(Erhocm * Erhocm - Mh2);
  const double rph = rng.uniform() * 2. * SIMC_PI_D;
  const double rth1 = rng.uniform() * 2. - 1.;
  const double rth = m::acos(rth1);
  const double srth = m::sin(rth);
  const double pxr = Prhocm * srth * m::cos(rph);
  const double pyr = Prhocm * srth * m::sin(rph);
  const double pzr = Prhocm * m::cos(rth);
  const MV4 f = loren(gammacm, bx, by, bz, Erhocm, pxr, pyr, pzr);
  s.v_pdelta = 100. * (f.p / cfg.spec_p.P - 1.);
  s.v_pP = f.p;
  s.v_pE = f.e;
  s.upx = f.x / f.p; s.upy = f.y / f.p; s.upz = f.z / f.p;
  s.v_pxptar = m::acos(f.z / f.p);              // lab theta and phi of the rho, "not really used for anything"
  s.v_pyptar = m::atan2(f.y, f.x);              // (generate_rho.f:123-126): they do reach the gen histograms
  return true;
}

// rho_decay.f:1-163: rho0 -> pi+ pi- before the spectrometer.  The detected pion is thrown flat in phi and like
// sin^2 + 2 eps R cos^2 (rejection loop, R = sigma_L/sigma_T of HERMES) in the rho rest frame and boosted to the lab;
// `orig` becomes that pion and COMMON Mh the pion mass.  False: the pion cannot reach the hadron arm.
template <class RNG>
SIMC_HD_CALL bool rho_decay(const simc_run_config& cfg, RNG& rng, EventState& s) {
  using namespace mesondetail;
  const double epsilon = s.m_eps;
  const double R_rho = 0.33 * m::pow(s.v_Q2 / (SIMC_MRHO * SIMC_MRHO), 0.61);
  const double ph = s.o_pP;
  const double beta = ph / sqrt(ph * ph + s.rho_mass * s.rho_mass);
  const double gamma = 1. / sqrt(1. - beta * beta);
  const double rph = rng.uniform() * 2. * SIMC_PI_D;
  double rth, srth, crth;
  for (;;) {
    const double rth1 = rng.uniform() * 2. - 1.;
    rth = m::acos(rth1);
    const double norm = (1.0 + 2.0 * epsilon * R_rho) * rng.uniform();
    srth = m::sin(rth); crth = m::cos(rth);
    const double dist = srth * srth + 2.0 * epsilon * R_rho * (crth * crth);
    if (!(dist < norm)) break;
  }
  s.rho_theta = rth;
  const double er = s.rho_mass / 2.0;
  if (er < SIMC_MPI) return false;
  const double pr = sqrt(er * er - SIMC_MPI * SIMC_MPI);
  const double pxr = pr * srth * m::cos(rph);
  const double pyr = pr * srth * m::sin(rph);
  const double pzr = pr * crth;
  const MV4 f = loren(gamma, -beta * s.upx, -beta * s.upy, -beta * s.upz, er, pxr, pyr, pzr);
  const int arm = cfg.hadron_arm;
  const bool right = arm == 1 || arm == 3;      // HMS, HRS-R; the others (SOS, HRS-L, SHMS) sit on the left
  const double th0 = cfg.spec_p.theta;
  const double pzprime = right ? f.z * m::cos(th0) - f.y * m::sin(th0) : f.z * m::cos(th0) + f.y * m::sin(th0);
  if (pzprime < 0.0) return false;
  const double th_oop = m::asin(f.x / f.p);
  const double cos_th_inp = f.z / f.p / m::cos(th_oop);
  double th_inp = m::acos(cos_th_inp);
  if (right) th_inp = f.y < 0.0 ? th0 - th_inp : th0 + th_inp;
  else th_inp = f.y > 0.0 ? th_inp - th0 : th_inp + th0;
  if ((th_inp > SIMC_PI_D / 2.) || (th_oop > SIMC_PI_D / 2.)) return false;
  s.o_pxptar = m::tan(th_oop);
  s.o_pyptar = m::tan(th_inp);
  s.o_pdelta = 100. * (f.p / cfg.spec_p.P - 1.);
  s.o_pP = f.p;
  s.o_pE = sqrt(s.o_pP * s.o_pP + SIMC_MPI * SIMC_MPI);
  return true;
}

// ---- H(e,e'pi) / H(e,e'K): complete_ev, event.f:432-1052 with doing_hydpi / doing_hydkaon ----------
// Needs v_Ein, v_eE, the electron and hadron angles and tz; fills the hadron energy from the
// two-body quadratic (event.f:634-698), W, epsilon, theta_pq, phi_pq, t (event.f:707-771), the
// jacobian, Eloss/teff(2:3) and the radiative constants.
template <class RNG, class GAUSS>
SIMC_HD bool complete_ev_meson(const simc_run_config& cfg, const MatTable& mt, RNG& rng, GAUSS gauss, EventState& s, bool run) {
  double Mh = cfg.Mh, Mh2 = cfg.Mh2;              // rho production: replaced by the mass generate_rho draws
  const simc_target& targ = cfg.targ;
  if (run) {
    s.jacobian = 1.0;
    s.uex = m::sin(s.v_etheta) * m::cos(s.v_ephi);
    s.uey = m::sin(s.v_etheta) * m::sin(s.v_ephi);
    s.uez = m::cos(s.v_etheta);
    if (!cfg.doing_rho) {                           // event.f:507-511
      s.upx = m::sin(s.v_ptheta) * m::cos(s.v_pphi);
      s.upy = m::sin(s.v_ptheta) * m::sin(s.v_pphi);
      s.upz = m::cos(s.v_ptheta);
    }
    const double eP = s.v_eE;
    s.v_nu = s.v_Ein - s.v_eE;
    s.v_Q2 = 2 * s.v_Ein * s.v_eE * (1. - s.uez);
    s.v_q = sqrt(s.v_Q2 + s.v_nu * s.v_nu);
    s.uqx = -eP * s.uex / s.v_q;
    s.uqy = -eP * s.uey / s.v_q;
    s.uqz = (s.v_Ein - eP * s.uez) / s.v_q;
    if (cfg.doing_rho) {                            // event.f:701-708: momentum, energy and direction of the rho
      run = generate_rho(cfg, rng, s);
      Mh = s.rho_mass; Mh2 = Mh * Mh;
    } else if (cfg.doing_deuterium) {                      // event.f:565-607: E_p from the two-body quadratic, |dEp'/dEm|
      s.v_Em = targ.Mtar_struck + targ.Mrec - targ.M;
      const double Mrec = targ.M - targ.Mtar_struck + s.v_Em;
      const double a = -1. * s.v_q * (s.uqx * s.upx + s.uqy * s.upy + s.uqz * s.upz);
      const double b = s.v_q * s.v_q;
      const double c = s.v_nu + targ.M;
      const double t = c * c - b + Mh2 - Mrec * Mrec;
      const double QA = 4. * (a * a - c * c);
      const double QB = 4. * c * t;
      const double QC = -4. * (a * a) * Mh2 - t * t;
      const double radical = QB * QB - 4. * QA * QC;
      if (radical < 0) run = false;
      if (run) {
        s.v_pE = (-QB - sqrt(radical)) / 2. / QA;
        if (s.v_pE <= Mh) run = false;
      }
      if (run) {
        s.jacobian = fabs((t * (c - s.v_pE) + 2 * c * s.v_pE * (s.v_pE - c)) / (2 * (a * a - c * c) * s.v_pE + c * t));
        s.v_pP = sqrt(s.v_pE * s.v_pE - Mh2);
        s.v_pdelta = (s.v_pP - cfg.spec_p.P) * 100. / cfg.spec_p.P;
        // event.f:880-886, 939-941
        const double Pmx = s.v_pP * s.upx - s.v_q * s.uqx;
        const double Pmy = s.v_pP * s.upy - s.v_q * s.uqy;
        const double Pmz = s.v_pP * s.upz - s.v_q * s.uqz;
        s.v_Pm = sqrt(Pmx * Pmx + Pmy * Pmy + Pmz * Pmz);
        s.v_Trec = sqrt(Mrec * Mrec + s.v_Pm * s.v_Pm) - Mrec;
      }
    } else if (!cfg.doing_semi && !cfg.doing_rho) { // semi-inclusive: the hadron energy was thrown (event.f:289-297)
      s.v_Pm = s.pfer;                              // event.f:633 (zero for hydrogen)
      double a = -1. * s.v_q * (s.uqx * s.upx + s.uqy * s.upy + s.uqz * s.upz);
      double b = s.v_q * s.v_q;
      double c = s.v_nu + targ.M;
      if (cfg.doing_deutpi || cfg.doing_deutkaon || cfg.doing_hepi || cfg.doing_hekaon) {   // event.f:646-654: Fermi motion and binding
        a = a - fabs(s.pfer) * (s.pferx * s.upx + s.pfery * s.upy + s.pferz * s.upz);
        b = b + s.pfer * s.pfer + 2 * s.v_q * fabs(s.pfer) * (s.pferx * s.uqx + s.pfery * s.uqy + s.pferz * s.uqz);
        c = s.v_nu + s.efer;
      }
      const double t = c * c - b + Mh2 - targ.Mrec_struck * targ.Mrec_struck;
      const double QA = 4. * (a * a - c * c);
      const double QB = 4. * c * t;
      const double QC = -4. * (a * a) * Mh2 - t * t;
      const double radical = QB * QB - 4. * QA * QC;
      if (radical < 0) run = false;
      if (run) {
        s.v_pE = (-QB - sqrt(radical)) / 2. / QA;
        if (s.v_pE < 0.0) run = false;
        if (run && cfg.doing_delta) {               // event.f:680-684: one of the two solutions, by a coin toss
          const double Ehad2 = (-QB + sqrt(radical)) / 2. / QA;
          if (rng.uniform() > 0.5) s.v_pE = Ehad2;
        }
        if (!run) {}
        else if (c - s.v_pE <= targ.Mrec_struck) run = false;
        else if (s.v_pE <= Mh) run = false;
      }
    }
  }
  if (run && cfg.doing_deuterium) {
    // event.f:1013-1016: only the electron's (xptar, yptar) -> solid angle factor (doing_deuterium is not in
    // the list of event.f:1019-1020)
    const double r = sqrt(1. + s.v_eyptar * s.v_eyptar + s.v_exptar * s.v_exptar);
    s.jacobian = s.jacobian / (r * (r * r));
  }
  if (run && !cfg.doing_deuterium) {
    if (!cfg.doing_semi && !cfg.doing_rho) {
      s.v_pP = sqrt(s.v_pE * s.v_pE - Mh2);
      s.v_pdelta = (s.v_pP - cfg.spec_p.P) * 100. / cfg.spec_p.P;
    }
    // event.f:707-771
    const double W2 = targ.Mtar_struck * targ.Mtar_struck + 2. * targ.Mtar_struck * s.v_nu - s.v_Q2;
    s.m_W = sqrt(fabs(W2)) * W2 / fabs(W2);
    const double th2 = m::tan(s.v_etheta / 2.);
    s.m_eps = 1. / (1. + 2. * (1 + s.v_nu * s.v_nu / s.v_Q2) * (th2 * th2));
    s.m_thpq = m::acos(s.upx * s.uqx + s.upy * s.uqy + s.upz * s.uqz);
    s.m_t = s.v_Q2 - Mh2 + 2 * s.v_nu * s.v_pE - 2 * s.v_pP * s.v_q * m::cos(s.m_thpq);
    s.m_tmin = s.v_Q2 - Mh2 + 2 * s.v_pE * s.v_nu - 2 * s.v_pP * s.v_q;
    const double qx = -s.uqy, qy = s.uqx, qz = s.uqz;
    const double px = -s.upy, py = s.upx, pz = s.upz;
    double dummy = sqrt((qx * qx + qy * qy) * (qx * qx + qy * qy + qz * qz));
    const double new_x_x = -qx * qz / dummy, new_x_y = -qy * qz / dummy, new_x_z = (qx * qx + qy * qy) / dummy;
    dummy = sqrt(qx * qx + qy * qy);
    const double new_y_x = qy / dummy, new_y_y = -qx / dummy, new_y_z = 0.0;
    const double p_new_x = px * new_x_x + py * new_x_y + pz * new_x_z;
    const double p_new_y = px * new_y_x + py * new_y_y + pz * new_y_z;
    s.m_phipq = m::atan2(p_new_y, p_new_x);
    if (s.m_phipq < 0.e0) s.m_phipq = s.m_phipq + 2. * SIMC_PI_D;
    if (cfg.doing_pizero) {                         // event.f:899-901: pizero_decay's two random numbers
      s.rho_mass = rng.uniform() * 2. * SIMC_PI_D;
      const double rth1 = rng.uniform() * 2. - 1.;
      s.rho_theta = m::acos(rth1);
    }
    s.v_Trec = 0.0;
    if (cfg.doing_deutpi || cfg.doing_deutkaon || cfg.doing_hepi || cfg.doing_hekaon) {   // event.f:945-947: recoil of the spectator system
      const double Mrec = targ.M - targ.Mtar_struck + s.v_Em;
      s.v_Trec = sqrt(Mrec * Mrec + s.v_Pm * s.v_Pm) - Mrec;
    }
    if (cfg.doing_semi) {     // event.f:880-886, 952-955, 979-996: Pm, Em of the undetected system; z and pt^2
      const double Pmx = s.v_pP * s.upx - s.v_q * s.uqx;
      const double Pmy = s.v_pP * s.upy - s.v_q * s.uqy;
      const double Pmz = s.v_pP * s.upz - s.v_q * s.uqz;
      const double Pmiss = sqrt(Pmx * Pmx + Pmy * Pmy + Pmz * Pmz);
      s.v_Pm = Pmiss;
      s.v_Em = s.v_nu + targ.M - s.v_pE;
      const double e_x = targ.Mtar_struck + s.v_nu - s.v_pE;
      const double thr = SIMC_MP + 134.9766;
      if ((e_x * e_x - Pmiss * Pmiss) < thr * thr) run = false;
      if (run) {
        s.v_zhad = s.v_pE / s.v_nu;
        const double cth = m::cos(s.m_thpq);
        s.v_pt2 = s.v_pP * s.v_pP * (1.0 - cth * cth);
        if (s.v_zhad > 1.0) run = false;
      }
    }
  }
  if (run && !cfg.doing_deuterium) {
    // event.f:1013-1023: both arms' angles were generated
    double r = sqrt(1. + s.v_eyptar * s.v_eyptar + s.v_exptar * s.v_exptar);
    s.jacobian = s.jacobian / (r * (r * r));
    if (!cfg.doing_rho) {                           // "we generate rho's in 4pi ... no stinkin' Jacobian" (event.f:1017-1023)
      r = sqrt(1. + s.v_pyptar * s.v_pyptar + s.v_pxptar * s.v_pxptar);
      s.jacobian = s.jacobian / (r * (r * r));
    }
  }
  const double zpos = s.tz - targ.zoffset;
  SIMC_PHASE();
  if (run) trip_thru_target_sampled(cfg, mt, rng, gauss, 2, zpos, s.v_eE, s.v_etheta, SIMC_ME, s.Eloss[1], s.teff[1]);
  SIMC_PHASE();
  if (run) trip_thru_target_sampled(cfg, mt, rng, gauss, 3, zpos, s.v_pE, s.v_ptheta, Mh, s.Eloss[2], s.teff[2]);
  SIMC_PHASE();
  if (run) {
    if (!cfg.using_Eloss) { s.Eloss[1] = 0.0; s.Eloss[2] = 0.0; }
    VertexKin v;
    v.Ein = s.v_Ein; v.eE = s.v_eE; v.eP = s.v_eE; v.etheta = s.v_etheta; v.pE = s.v_pE; v.pP = s.v_pP;
    v.uex = s.uex; v.uey = s.uey; v.uez = s.uez; v.upx = s.upx; v.upy = s.upy; v.upz = s.upz;
    radc_init_ev(cfg, v, s.teff[0], s.teff[1], s.rad);
  }
  SIMC_PHASE();
  return run;
}

// generate + generate_rad for hydrogen meson production: event.f:126-428 (both arms' angles and the
// electron energy are thrown, :283-318), radc.f:120-519 with the doing_pion/doing_kaon photon-energy
// limits (:289-294) and no Em constraints on tails 2 and 3 (doing_eep = .false.).
template <class RNG, class GAUSS>
SIMC_HD bool generate_meson_first(const simc_run_config& cfg, const MatTable& mt, const PfermiDev pfm, const SfDev sf, RNG& rng,
                                  GAUSS gauss, EventState& s, bool ok, GenRad& gr) {
  const simc_target& targ = cfg.targ;
  const simc_gen_limits& gen = cfg.gen;
  if (ok) {
    s.tx = gauss(rng, 3.0) * gen.xwid + targ.xoffset;
    s.ty = gauss(rng, 3.0) * gen.ywid + targ.yoffset;
    double t3, t4, t5, t6;
    if (targ.fr_pattern == 1) {
      t3 = rng.uniform() * SIMC_PI_D;
      t4 = rng.uniform() * SIMC_PI_D;
      t5 = m::cos(t3) * targ.fr1;
      t6 = m::cos(t4) * targ.fr2;
    } else if (targ.fr_pattern == 2) {
      t3 = rng.uniform() * 2. * SIMC_PI_D;
      t4 = sqrt(rng.uniform()) * (targ.fr2 - targ.fr1) + targ.fr1;
      t5 = m::cos(t3) * t4;
      t6 = m::sin(t3) * t4;
    } else if (targ.fr_pattern == 3) {
      t3 = 2. * rng.uniform() - 1.0;
      t4 = 2. * rng.uniform() - 1.0;
      t5 = targ.fr1 * t3;
      t6 = targ.fr2 * t4;
    } else {
      t5 = 0.0; t6 = 0.0;
    }
    s.tx = s.tx + t5;
    s.ty = s.ty + t6;
    s.tz = (0.5 - rng.uniform()) * targ.length + targ.zoffset;
    s.rastery = t6;
    s.rasterx = t5;
    trip_thru_target_sampled(cfg, mt, rng, gauss, 1, s.tz - targ.zoffset, cfg.Ebeam, 0.0, SIMC_ME, s.Eloss[0], s.teff[0]);
    if (!cfg.using_Eloss) s.Eloss[0] = 0.0;
    s.Coulomb = cfg.using_Coulomb ? targ.Coulomb_constant : 0.0;
    s.v_Ein = cfg.Ebeam + (rng.uniform() - 0.5) * cfg.dEbeam + s.Coulomb - s.Eloss[0];
    s.Ein_shift = s.v_Ein - cfg.Ebeam_vertex_ave;
    s.Ee_shift = s.Coulomb - targ.Coulomb_ave;
    s.gen_weight = 1.0;
    s.v_eyptar = gen.e.yptar.min + rng.uniform() * (gen.e.yptar.max - gen.e.yptar.min);
    s.v_exptar = gen.e.xptar.min + rng.uniform() * (gen.e.xptar.max - gen.e.xptar.min);
    if (!cfg.doing_rho) {   // event.f:277-283.  Rho production throws no hadron angles: the reference converts whatever
      // the previous event left in vertex%p%xptar/yptar; a try starts from zero here (DESIGN, known deviations)
      s.v_pyptar = gen.p.yptar.min + rng.uniform() * (gen.p.yptar.max - gen.p.yptar.min);
      s.v_pxptar = gen.p.xptar.min + rng.uniform() * (gen.p.xptar.max - gen.p.xptar.min);
    }
    if (cfg.doing_semi) {   // hadron energy, event.f:289-297
      const double Emin = fmax(gen.p.E.min, gen.sumEgen.min - gen.e.E.max);
      const double Emax = fmin(gen.p.E.max, gen.sumEgen.max - gen.e.E.min);
      if (Emin > Emax) ok = false;
      if (ok) {
        s.gen_weight = s.gen_weight * (Emax - Emin) / (gen.p.E.max - gen.p.E.min);
        s.v_pE = Emin + rng.uniform() * (Emax - Emin);
        s.v_pP = sqrt(s.v_pE * s.v_pE - cfg.Mh2);
        s.v_pdelta = 100. * (s.v_pP - cfg.spec_p.P) / cfg.spec_p.P;
      }
    }
    // event.f:296-318 (semi-inclusive: the plain electron-arm limits)
    const double Emin = cfg.doing_semi ? gen.e.E.min : fmax(gen.e.E.min, gen.sumEgen.min);
    const double Emax = cfg.doing_semi ? gen.e.E.max : fmin(gen.e.E.max, gen.sumEgen.max);
    if (Emin > Emax) ok = false;
    if (ok) {
      s.gen_weight = s.gen_weight * (Emax - Emin) / (gen.e.E.max - gen.e.E.min);
      s.v_eE = Emin + rng.uniform() * (Emax - Emin);
      s.v_edelta = 100. * (s.v_eE - cfg.spec_e.P) / cfg.spec_e.P;
      physics_angles(cfg.spec_e.theta, cfg.spec_e.phi, s.v_exptar, s.v_eyptar, s.v_etheta, s.v_ephi);
      physics_angles(cfg.spec_p.theta, cfg.spec_p.phi, s.v_pxptar, s.v_pyptar, s.v_ptheta, s.v_pphi);
      s.v_Em = 0.0;
      s.rad.Egamma_used[0] = s.rad.Egamma_used[1] = s.rad.Egamma_used[2] = 0.0;
      s.rad.ntail = 0;
      // event.f:327-373: nucleon momentum in the deuteron (thrown whether or not do_fermi uses it)
      s.pfer = 0.0; s.pferx = 0.0; s.pfery = 0.0; s.pferz = 0.0;
      s.efer = targ.Mtar_struck;
      if (cfg.doing_deutsemi || cfg.doing_deutpi || cfg.doing_deutkaon || cfg.doing_hepi || cfg.doing_hekaon) {
        const double ranprob = rng.uniform();
        const int nump = pfm.nump;
        // first ii (1-based) with ranprob <= mprob(ii), capped at nump: the reference's linear scan
        // (event.f:341-343) on a non-decreasing table
        int lo = 0, hi = nump - 1;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (ranprob > pfm.mprob[mid]) lo = mid + 1; else hi = mid;
        }
        const int ii = lo + 1;
        const double pferlo = ii == 1 ? 0.0 : (pfm.pval[ii - 2] + pfm.pval[ii - 1]) / 2;
        const double pferhi = ii == nump ? pfm.pval[nump - 1] : (pfm.pval[ii - 1] + pfm.pval[ii]) / 2;
        s.pfer = pferlo + (pferhi - pferlo) * rng.uniform();
        const double ranth1 = rng.uniform() * 2. - 1.0;
        const double ranth = m::acos(ranth1);
        const double ranph = rng.uniform() * 2. * SIMC_PI_D;
        s.pferx = m::sin(ranth) * m::cos(ranph);
        s.pfery = m::sin(ranth) * m::sin(ranph);
        s.pferz = m::cos(ranth);
        if (cfg.doing_hepi || cfg.doing_hekaon) {      // event.f:368-372
          const double u1 = rng.uniform();
          const double u2 = rng.uniform();
          s.v_Em = generate_em(sf, s.pfer, u1, u2);
        } else {
          s.v_Em = SIMC_MP + 939.56563 - targ.M;
        }
        const double m_spec = targ.M - targ.Mtar_struck + s.v_Em;
        s.efer = targ.M - sqrt(m_spec * m_spec + s.pfer * s.pfer);
      }
    }
  }
  SIMC_PHASE();
  ok = complete_ev_meson(cfg, mt, rng, gauss, s, ok);
  if (ok) s.Trec = s.v_Trec;
  gr.emin = 0.0; gr.emax = 0.0; gr.eg = 0.0; gr.bw = 0.0; gr.which = 0;
  if (!cfg.using_rad) return ok;
  if (cfg.rad_flag >= 2) return ok ? generate_rad_basis_first(cfg, rng, s, gr) : false;
  RadEvDev& R = s.rad;
  double bw = 0, emin = 0.0, emax = 0.0, eg = 0.0;
  int which = 0;
  if (ok) {
    const double x = rng.uniform();
    if (x >= R.frac[0] + R.frac[1]) R.ntail = 3;
    else if (x >= R.frac[0]) R.ntail = 2;
    else R.ntail = 1;
    const int ntail = R.ntail;
    const bool eep = cfg.doing_eep != 0;            // D(e,e'p): the measured-Em clauses of radc.f:369-374, 420-425
    const double max_delta_Trec = fmax((s.v_Trec - cfg.VERTEXedge.Trec.min), (cfg.VERTEXedge.Trec.max - s.v_Trec));
    if (cfg.doing_tail[0] && ntail == 1) {          // radc.f:280-294
      if (cfg.doing_deuterium) {
        emax = fmin(cfg.Egamma1_max, gen.sumEgen.max - s.v_eE);
        emin = gen.sumEgen.min - s.v_eE;            // ntail != 0
      } else {
        emin = 0.;
        emax = gen.sumEgen.max - s.v_eE;
      }
      emax = fmin(emax, cfg.Egamma1_max);
      which = 1;
    } else if (cfg.doing_tail[1] && ntail == 2) {   // radc.f:358-374
      emin = s.v_eE - cfg.edge.e.E.max;
      emax = s.v_eE - cfg.edge.e.E.min;
      if (eep) {
        emax = fmin(emax, (cfg.edge.Em.max - s.v_Em) - R.Egamma_used[0] + max_delta_Trec);
        emin = fmax(emin, (cfg.edge.Em.min - s.v_Em) - R.Egamma_used[0] - max_delta_Trec);
      }
      emax = fmin(emax, cfg.Egamma_tot_max - R.Egamma_used[0]);
      which = 2;
    } else if (R.rad_proton_this_ev && ntail == 3) {   // radc.f:409-425
      emin = s.v_pE - cfg.edge.p.E.max;
      emax = s.v_pE - cfg.edge.p.E.min;
      if (eep) {
        emax = fmin(emax, (cfg.edge.Em.max - s.v_Em) - R.Egamma_used[0] - R.Egamma_used[1] + max_delta_Trec);
        emin = fmax(emin, (cfg.edge.Em.min - s.v_Em) - R.Egamma_used[0] - R.Egamma_used[1] - max_delta_Trec);
      }
      emax = fmin(emax, cfg.Egamma_tot_max - R.Egamma_used[0] - R.Egamma_used[1]);
      which = 3;
    }
    if (which) {
      emin = emin - cfg.dE_edge_test;
      emax = emax + cfg.dE_edge_test;
      if (cfg.hardwired_rad) emax = cfg.Egamma_gen_max;
      basicrad4(R, rng, emin, emax, eg, bw);
      if (bw <= 0) ok = false;
      else {
        R.Egamma_used[which - 1] = eg;
        if (which == 1) s.v_Ein = s.v_Ein - eg;
      }
    }
  }
  gr.emin = emin; gr.emax = emax; gr.eg = eg; gr.bw = bw; gr.which = which;
  return ok;
}
template <class RNG, class GAUSS>
SIMC_HD bool generate_meson_second(const simc_run_config& cfg, const MatTable& mt, RNG& rng, GAUSS gauss, EventState& s, bool run) {
  return complete_ev_meson(cfg, mt, rng, gauss, s, run);
}

// ---- A(e,e'p) on a nucleus (doing_heavy): complete_ev, event.f:432-1052 --------------------------
// Both energies and both directions are thrown; the kinematics only derive q, the missing
// momentum/energy and the recoil system (event.f:914-955).
template <class RNG, class GAUSS>
SIMC_HD bool complete_ev_heavy(const simc_run_config& cfg, const MatTable& mt, RNG& rng, GAUSS gauss, EventState& s,
                               bool run) {
  const simc_target& targ = cfg.targ;
  if (run) {
    s.jacobian = 1.0;
    s.uex = m::sin(s.v_etheta) * m::cos(s.v_ephi);
    s.uey = m::sin(s.v_etheta) * m::sin(s.v_ephi);
    s.uez = m::cos(s.v_etheta);
    s.upx = m::sin(s.v_ptheta) * m::cos(s.v_pphi);
    s.upy = m::sin(s.v_ptheta) * m::sin(s.v_pphi);
    s.upz = m::cos(s.v_ptheta);
    const double eP = s.v_eE;
    s.v_nu = s.v_Ein - s.v_eE;
    s.v_Q2 = 2 * s.v_Ein * s.v_eE * (1. - s.uez);
    s.v_q = sqrt(s.v_Q2 + s.v_nu * s.v_nu);
    s.uqx = -eP * s.uex / s.v_q;
    s.uqy = -eP * s.uey / s.v_q;
    s.uqz = (s.v_Ein - eP * s.uez) / s.v_q;
    // event.f:880-955
    const double Pmx = s.v_pP * s.upx - s.v_q * s.uqx;
    const double Pmy = s.v_pP * s.upy - s.v_q * s.uqy;
    const double Pmz = s.v_pP * s.upz - s.v_q * s.uqz;
    const double Pmiss = sqrt(Pmx * Pmx + Pmy * Pmy + Pmz * Pmz);
    const double Emiss = s.v_nu + targ.M - s.v_pE;
    s.v_Pm = Pmiss;
    const double Mrec = sqrt(Emiss * Emiss - Pmiss * Pmiss);
    s.v_Em = targ.Mtar_struck + Mrec - targ.M;
    s.v_Trec = sqrt(Mrec * Mrec + s.v_Pm * s.v_Pm) - Mrec;
    double r = sqrt(1. + s.v_eyptar * s.v_eyptar + s.v_exptar * s.v_exptar);
    s.jacobian = s.jacobian / (r * (r * r));
    r = sqrt(1. + s.v_pyptar * s.v_pyptar + s.v_pxptar * s.v_pxptar);
    s.jacobian = s.jacobian / (r * (r * r));
  }
  const double zpos = s.tz - targ.zoffset;
  SIMC_PHASE();
  if (run) trip_thru_target_sampled(cfg, mt, rng, gauss, 2, zpos, s.v_eE, s.v_etheta, SIMC_ME, s.Eloss[1], s.teff[1]);
  SIMC_PHASE();
  if (run) trip_thru_target_sampled(cfg, mt, rng, gauss, 3, zpos, s.v_pE, s.v_ptheta, cfg.Mh, s.Eloss[2], s.teff[2]);
  SIMC_PHASE();
  if (run) {
    if (!cfg.using_Eloss) { s.Eloss[1] = 0.0; s.Eloss[2] = 0.0; }
    VertexKin v;
    v.Ein = s.v_Ein; v.eE = s.v_eE; v.eP = s.v_eE; v.etheta = s.v_etheta; v.pE = s.v_pE; v.pP = s.v_pP;
    v.uex = s.uex; v.uey = s.uey; v.uez = s.uez; v.upx = s.upx; v.upy = s.upy; v.upz = s.upz;
    radc_init_ev(cfg, v, s.teff[0], s.teff[1], s.rad);
  }
  SIMC_PHASE();
  return run;
}

// generate + generate_rad for A(e,e'p): event.f:126-428 (hadron energy :283-294, electron energy
// :296-318), radc.f:120-519 with the doing_heavy photon-energy limits (:249-263) and the Em/Pm window
// test after the first tail (:340-350).
template <class RNG, class GAUSS>
SIMC_HD bool generate_heavy_first(const simc_run_config& cfg, const MatTable& mt, RNG& rng, GAUSS gauss, EventState& s, bool ok,
                                  GenRad& gr) {
  const simc_target& targ = cfg.targ;
  const simc_gen_limits& gen = cfg.gen;
  if (ok) {
    s.tx = gauss(rng, 3.0) * gen.xwid + targ.xoffset;
    s.ty = gauss(rng, 3.0) * gen.ywid + targ.yoffset;
    double t3, t4, t5, t6;
    if (targ.fr_pattern == 1) {
      t3 = rng.uniform() * SIMC_PI_D;
      t4 = rng.uniform() * SIMC_PI_D;
      t5 = m::cos(t3) * targ.fr1;
      t6 = m::cos(t4) * targ.fr2;
    } else if (targ.fr_pattern == 2) {
      t3 = rng.uniform() * 2. * SIMC_PI_D;
      t4 = sqrt(rng.uniform()) * (targ.fr2 - targ.fr1) + targ.fr1;
      t5 = m::cos(t3) * t4;
      t6 = m::sin(t3) * t4;
    } else if (targ.fr_pattern == 3) {
      t3 = 2. * rng.uniform() - 1.0;
      t4 = 2. * rng.uniform() - 1.0;
      t5 = targ.fr1 * t3;
      t6 = targ.fr2 * t4;
    } else {
      t5 = 0.0; t6 = 0.0;
    }
    s.tx = s.tx + t5;
    s.ty = s.ty + t6;
    s.tz = (0.5 - rng.uniform()) * targ.length + targ.zoffset;
    s.rastery = t6;
    s.rasterx = t5;
    trip_thru_target_sampled(cfg, mt, rng, gauss, 1, s.tz - targ.zoffset, cfg.Ebeam, 0.0, SIMC_ME, s.Eloss[0], s.teff[0]);
    if (!cfg.using_Eloss) s.Eloss[0] = 0.0;
    s.Coulomb = cfg.using_Coulomb ? targ.Coulomb_constant : 0.0;
    s.v_Ein = cfg.Ebeam + (rng.uniform() - 0.5) * cfg.dEbeam + s.Coulomb - s.Eloss[0];
    s.Ein_shift = s.v_Ein - cfg.Ebeam_vertex_ave;
    s.Ee_shift = s.Coulomb - targ.Coulomb_ave;
    s.gen_weight = 1.0;
    s.v_eyptar = gen.e.yptar.min + rng.uniform() * (gen.e.yptar.max - gen.e.yptar.min);
    s.v_exptar = gen.e.xptar.min + rng.uniform() * (gen.e.xptar.max - gen.e.xptar.min);
    s.v_pyptar = gen.p.yptar.min + rng.uniform() * (gen.p.yptar.max - gen.p.yptar.min);
    s.v_pxptar = gen.p.xptar.min + rng.uniform() * (gen.p.xptar.max - gen.p.xptar.min);
    {   // hadron energy, event.f:283-294
      const double Emin = fmax(gen.p.E.min, gen.sumEgen.min - gen.e.E.max);
      const double Emax = fmin(gen.p.E.max, gen.sumEgen.max - gen.e.E.min);
      if (Emin > Emax) ok = false;
      if (ok) {
        s.gen_weight = s.gen_weight * (Emax - Emin) / (gen.p.E.max - gen.p.E.min);
        s.v_pE = Emin + rng.uniform() * (Emax - Emin);
        s.v_pP = sqrt(s.v_pE * s.v_pE - cfg.Mh2);
        s.v_pdelta = 100. * (s.v_pP - cfg.spec_p.P) / cfg.spec_p.P;
      }
    }
    if (ok) {   // electron energy, event.f:296-318
      const double Emin = fmax(gen.e.E.min, gen.sumEgen.min - s.v_pE);
      const double Emax = fmin(gen.e.E.max, gen.sumEgen.max - s.v_pE);
      if (Emin > Emax) ok = false;
      if (ok) {
        s.gen_weight = s.gen_weight * (Emax - Emin) / (gen.e.E.max - gen.e.E.min);
        s.v_eE = Emin + rng.uniform() * (Emax - Emin);
        s.v_edelta = 100. * (s.v_eE - cfg.spec_e.P) / cfg.spec_e.P;
        physics_angles(cfg.spec_e.theta, cfg.spec_e.phi, s.v_exptar, s.v_eyptar, s.v_etheta, s.v_ephi);
        physics_angles(cfg.spec_p.theta, cfg.spec_p.phi, s.v_pxptar, s.v_pyptar, s.v_ptheta, s.v_pphi);
        s.v_Em = 0.0;
        s.rad.Egamma_used[0] = s.rad.Egamma_used[1] = s.rad.Egamma_used[2] = 0.0;
        s.rad.ntail = 0;
      }
    }
  }
  SIMC_PHASE();
  ok = complete_ev_heavy(cfg, mt, rng, gauss, s, ok);
  if (ok) s.Trec = s.v_Trec;
  const simc_edge& VE = cfg.VERTEXedge;
  gr.emin = 0.0; gr.emax = 0.0; gr.eg = 0.0; gr.bw = 0.0; gr.which = 0;
  if (!cfg.using_rad) {
    if (ok) ok = (s.v_Em >= VE.Em.min && s.v_Em <= VE.Em.max && s.v_Pm >= VE.Pm.min && s.v_Pm <= VE.Pm.max);
    return ok;
  }
  if (cfg.rad_flag >= 2) return ok ? generate_rad_basis_first(cfg, rng, s, gr) : false;
  RadEvDev& R = s.rad;
  double bw = 0, emin = 0.0, emax = 0.0, eg = 0.0, max_delta_Trec = 0.0;
  int which = 0, ntail = 0;
  if (ok) {
    const double x = rng.uniform();
    if (x >= R.frac[0] + R.frac[1]) R.ntail = 3;
    else if (x >= R.frac[0]) R.ntail = 2;
    else R.ntail = 1;
    ntail = R.ntail;
    max_delta_Trec = fmax((s.v_Trec - VE.Trec.min), (VE.Trec.max - s.v_Trec));
    if (cfg.doing_tail[0] && ntail == 1) {          // radc.f:249-263
      emin = s.v_Em - VE.Em.max - max_delta_Trec;
      emax = s.v_Em - VE.Em.min + max_delta_Trec;
      emax = fmin(emax, s.v_Em - cfg.edge.Em.min + max_delta_Trec);       // ntail != 0
      emax = fmin(emax, cfg.Egamma1_max);
      emin = emin - cfg.dE_edge_test;
      emax = emax + cfg.dE_edge_test;
      if (cfg.hardwired_rad) emax = cfg.Egamma_gen_max;
      which = 1;
      basicrad4(R, rng, emin, emax, eg, bw);
      if (bw <= 0) ok = false;
      else { R.Egamma_used[0] = eg; s.v_Ein = s.v_Ein - eg; }
    }
  }
  // tries with which == 1 re-enter complete_ev (radc.f:324) in generate_heavy_second; the others go on here
  if (ok && !which) {
    // radc.f:340-350: the vertex must sit inside the spectral function's window
    if (s.v_Em < VE.Em.min || s.v_Em > VE.Em.max || s.v_Pm < VE.Pm.min || s.v_Pm > VE.Pm.max) ok = false;
  }
  if (ok && !which) {
    if (cfg.doing_tail[1] && ntail == 2) {          // radc.f:358-374
      emin = s.v_eE - cfg.edge.e.E.max;
      emax = s.v_eE - cfg.edge.e.E.min;
      emax = fmin(emax, (cfg.edge.Em.max - s.v_Em) - R.Egamma_used[0] + max_delta_Trec);
      emin = fmax(emin, (cfg.edge.Em.min - s.v_Em) - R.Egamma_used[0] - max_delta_Trec);
      emax = fmin(emax, cfg.Egamma_tot_max - R.Egamma_used[0]);
      which = 2;
    } else if (R.rad_proton_this_ev && ntail == 3) {   // radc.f:409-425
      emin = s.v_pE - cfg.edge.p.E.max;
      emax = s.v_pE - cfg.edge.p.E.min;
      emax = fmin(emax, (cfg.edge.Em.max - s.v_Em) - R.Egamma_used[0] - R.Egamma_used[1] + max_delta_Trec);
      emin = fmax(emin, (cfg.edge.Em.min - s.v_Em) - R.Egamma_used[0] - R.Egamma_used[1] - max_delta_Trec);
      emax = fmin(emax, cfg.Egamma_tot_max - R.Egamma_used[0] - R.Egamma_used[1]);
      which = 3;
    }
    if (which) {
      emin = emin - cfg.dE_edge_test;
      emax = emax + cfg.dE_edge_test;
      if (cfg.hardwired_rad) emax = cfg.Egamma_gen_max;
      basicrad4(R, rng, emin, emax, eg, bw);
      if (bw <= 0) ok = false;
      else R.Egamma_used[which - 1] = eg;
    }
  }
  gr.emin = emin; gr.emax = emax; gr.eg = eg; gr.bw = bw; gr.which = which;
  return ok;
}
// second pass through complete_ev for the tries whose incoming electron radiated (radc.f:324), then the
// spectral-function window of radc.f:340-350
template <class RNG, class GAUSS>
SIMC_HD bool generate_heavy_second(const simc_run_config& cfg, const MatTable& mt, RNG& rng, GAUSS gauss, EventState& s, bool run) {
  bool ok = complete_ev_heavy(cfg, mt, rng, gauss, s, run);
  const simc_edge& VE = cfg.VERTEXedge;
  if (ok) {
    if (s.v_Em < VE.Em.min || s.v_Em > VE.Em.max || s.v_Pm < VE.Pm.min || s.v_Pm > VE.Pm.max) ok = false;
  }
  return ok;
}

// Target-to-spectrometer transformation of one arm, simc.f:1379-1443 (P) / :1655-1706 (E)
struct ArmEntry {
  double sp_delta, sp_yptar, sp_xptar, sp_z;      // main%SP
  double x, y, dx, dy;                            // TRANSPORT coordinates at z = 0
};
SIMC_HD void arm_entry(const simc_spectrometer& sp, double tx, double ty, double tz, double sp_delta, double sp_yptar,
                       double sp_xptar, ArmEntry& a) {
  a.sp_delta = sp_delta; a.sp_yptar = sp_yptar; a.sp_xptar = sp_xptar;
  double x_arm = -ty;
  double y_arm = -tx * sp.cos_th - tz * sp.sin_th * m::sin(sp.phi);
  double z_arm = tz * sp.cos_th + tx * sp.sin_th * m::sin(sp.phi);
  x_arm = x_arm - sp.off_x;
  y_arm = y_arm - sp.off_y;
  z_arm = z_arm - sp.off_z;
  a.dx = sp_xptar - sp.off_xptar;
  a.dy = sp_yptar - sp.off_yptar;
  a.x = x_arm - z_arm * a.dx;
  a.y = y_arm - z_arm * a.dy;
  a.sp_z = a.y;
}

}  // namespace simc
