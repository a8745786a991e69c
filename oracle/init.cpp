// ORACLE -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (no reference-held golden values exist).
// The reference's one-time setup restated a second time, independently of the product's run_init.cu and on the
// oracle's own trip_thru_target / enerloss_new (oracle/target.cpp): the post-processing of dbase_read
// (dbase.f:119-553), target_init (init.f:1-87), limits_init (init.f:91-572) with extreme_trip_thru_target
// (target.f:310-544), and radc_init (init.f:576-651).  Input: the deck as "key = value" lines (keys lower case; the
// deck text is split by the tests, not by the product's reader) plus what the reference reads from data files at
// this point (x_pm_theory_absmax, x_e_fermi: theory_init; x_pval_last: deut.dat / he3.dat; x_emval_last: the
// spectral function).  tests/test_init_cpu.py compares the result with simc_b200_config_from_deck field by field.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include "event.hpp"

namespace simc_oracle {
namespace {

struct KV {
  std::map<std::string, std::string> m;
  explicit KV(const char* text) {
    std::istringstream in(text);
    std::string line;
    while (std::getline(in, line)) {
      const size_t eq = line.find('=');
      if (eq == std::string::npos) continue;
      auto trim = [](std::string s) {
        while (!s.empty() && std::isspace((unsigned char)s.front())) s.erase(s.begin());
        while (!s.empty() && std::isspace((unsigned char)s.back())) s.pop_back();
        return s;
      };
      m[trim(line.substr(0, eq))] = trim(line.substr(eq + 1));
    }
  }
  bool has(const std::string& k) const { return m.count(k) != 0; }
  double d(const std::string& k, double def = 0.0) const {
    auto it = m.find(k);
    if (it == m.end() || it->second.empty()) return def;
    std::string v = it->second;
    for (char& c : v) if (c == 'd' || c == 'D') c = 'e';
    return std::strtod(v.c_str(), nullptr);
  }
  long i(const std::string& k, long def = 0) const { return std::lround(d(k, (double)def)); }
  bool flag(const std::string& k, long def = 0) const { return i(k, def) > 0; }      // convert_to_logical, dbase.f:1128-1165
};

const double Me = 0.51099906, Mp = 938.27231, Mn = 939.56563, Mpi = 139.57018, Mk = 493.677, Md = 1875.613,
             Mlambda = 1115.68, Msigma0 = 1192.64, Msigma_minus = 1197.45, MDelta = 1232.0, Mrho = 769.3,
             amu = 931.49432, hbarc = 197.327053, pi = 3.141592653589793, alpha = 1. / 137.0359895, degrad = 180. / pi;

struct Lim { double min, max; };

struct TargX { double Eloss_ave[3], Eloss_min[3], Eloss_max[3], teff_ave[3], teff_min[3], teff_max[3], musc_max[3]; };

void trip(simc_run_config& c, int narm, double z, double energy, double theta, double& Eloss, double& radlen, double mass,
          int typeflag) {
  Sim s;
  s.cfg = &c;
  trip_thru_target(s, narm, z, energy, theta, Eloss, radlen, mass, typeflag);
}

// target.f:569-577 (entry extreme_target_musc): 3.5 sigma of the Lynch-Dahl width
double extreme_musc(double p, double beta, double teff) {
  const double Es = 13.6, epsilon = 0.088, nsig_max = 3.5;
  const double theta_sigma = Es / p / beta * std::sqrt(teff) * (1 + epsilon * std::log10(teff / (beta * beta)));
  return theta_sigma * nsig_max;
}

// target.f:310-544
void extreme_trip(simc_run_config& c, TargX& x, double ebeam, Lim the, Lim thp, Lim pe, Lim pp, Lim z, double m) {
  const simc_target& targ = c.targ;
  const double inch_cm = 2.54, target_pi = 3.14159265358979;
  const bool liquid = targ.Z < 2.4;
  double E1, E2, E3, E4, t1, t2, t3, t4, zz = 0, th1 = 0, th2 = 0;
  trip(c, 1, z.max, ebeam, 0.0, x.Eloss_max[0], x.teff_max[0], Me, 3);
  trip(c, 1, z.min, ebeam, 0.0, x.Eloss_min[0], x.teff_min[0], Me, 2);
  double energymin = pe.min, energymax = pe.max;
  auto corner_shot = [&](const Lim& th) {
    double th_corner_max;
    if (z.max >= targ.length / 2.) th_corner_max = target_pi / 2.;
    else th_corner_max = std::atan(1.25 * inch_cm / (targ.length / 2. - z.max));
    const double th_corner_min = std::atan(1.25 * inch_cm / (targ.length / 2. - z.min));
    if (th_corner_min <= th.max && th_corner_max >= th.min) {
      const double th_corner = std::max(th_corner_min, th.min);
      zz = targ.length / 2. - 1.25 * inch_cm / std::tan(th_corner);
      th1 = th_corner - .0001;
      th2 = th_corner + .0001;
    } else {
      zz = z.min; th1 = th.min; th2 = th.max;
    }
  };
  if (!liquid) {
    trip(c, 2, z.min, energymax, the.max, x.Eloss_max[1], x.teff_max[1], Me, 3);
    trip(c, 2, z.max, energymin, the.min, x.Eloss_min[1], x.teff_min[1], Me, 2);
  } else if (targ.can == 1) {
    corner_shot(the);
    trip(c, 2, zz, energymax, th1, E1, t1, Me, 3);
    trip(c, 2, zz, energymax, th2, E2, t2, Me, 3);
    x.Eloss_max[1] = std::max(E1, E2);
    x.teff_max[1] = std::max(t1, t2);
  } else if (targ.can == 2 || targ.can == 3) {
    zz = -(targ.length / 2.) / std::tan(the.min);
    zz = std::max(zz, (-targ.length / 2.));
    trip(c, 2, zz, energymax, the.min, x.Eloss_max[1], x.teff_max[1], Me, 3);
  }
  if (liquid) {
    x.Eloss_min[1] = 1.e10;
    for (int i = 0; i <= 3; ++i) {
      trip(c, 2, z.min + (i / 2) * (z.max - z.min), energymin, the.min + (i % 2) * (the.max - the.min), E1, t1, Me, 2);
      if (E1 < x.Eloss_min[1]) { x.Eloss_min[1] = E1; x.teff_min[1] = t1; }
    }
  }
  energymin = std::sqrt(pp.min * pp.min + m * m);
  energymax = std::sqrt(pp.max * pp.max + m * m);
  const double betap_min = pp.min / std::sqrt(pp.min * pp.min + m * m);
  if (!liquid) {
    trip(c, 3, z.min, energymin, thp.max, E1, t1, m, 3);
    trip(c, 3, z.min, energymax, thp.max, E2, t2, m, 3);
    x.Eloss_max[2] = std::max(E1, E2);
    x.teff_max[2] = std::max(t1, t2);
    trip(c, 3, z.max, energymin, thp.min, E1, t1, m, 2);
    trip(c, 3, z.max, energymax, thp.min, E2, t2, m, 2);
    x.Eloss_min[2] = std::min(E1, E2);
    x.teff_min[2] = std::min(t1, t2);
  } else {
    if (targ.can == 1) {
      corner_shot(thp);
      trip(c, 3, zz, energymin, th1, E1, t1, m, 3);
      trip(c, 3, zz, energymin, th2, E2, t2, m, 3);
      trip(c, 3, zz, energymax, th1, E3, t3, m, 3);
      trip(c, 3, zz, energymax, th2, E4, t4, m, 3);
      x.Eloss_max[2] = std::max(std::max(E1, E2), std::max(E3, E4));
      x.teff_max[2] = std::max(std::max(t1, t2), std::max(t3, t4));
    } else {
      zz = -(targ.length / 2.) / std::tan(the.min);          // the%min, as written (target.f:513)
      zz = std::max(zz, (-targ.length / 2.));
      trip(c, 3, zz, energymin, thp.min, E1, t1, m, 3);
      trip(c, 3, zz, energymax, thp.min, E2, t2, m, 3);
      x.Eloss_max[2] = std::max(E1, E2);
      x.teff_max[2] = std::max(t1, t2);
    }
    x.Eloss_min[2] = 1.e10;
    for (int i = 0; i <= 3; ++i) {
      const double zi = z.min + (i / 2) * (z.max - z.min), thi = thp.min + (i % 2) * (thp.max - thp.min);
      trip(c, 3, zi, energymin, thi, E1, t1, m, 2);
      if (E1 < x.Eloss_min[2]) { x.Eloss_min[2] = E1; x.teff_min[2] = t1; zz = zi; th1 = thi; }
    }
    trip(c, 3, zz, energymax, th1, E1, t1, m, 2);
    x.Eloss_min[2] = std::min(x.Eloss_min[2], E1);
  }
  x.musc_max[0] = extreme_musc(ebeam, 1.e0, x.teff_max[0]);
  x.musc_max[1] = extreme_musc(pe.min, 1.e0, x.teff_max[1]);
  x.musc_max[2] = extreme_musc(pp.min, betap_min, x.teff_max[2]);
}

}  // namespace

void init_from_kv(const char* text, simc_run_config& c) {
  const KV D(text);
  std::memset(&c, 0, sizeof(c));
  c.abi_version = SIMC_B200_ABI_VERSION;
  simc_target& targ = c.targ;
  // ---- what the deck sets (regallvars, dbase.f:967-1122) and convert_to_logical (dbase.f:1128-1165)
  c.Ebeam = D.d("ebeam"); c.dEbeam = D.d("debeam");
  c.electron_arm = (int)D.i("electron_arm"); c.hadron_arm = (int)D.i("hadron_arm");
  c.spec_e.P = D.d("spec%e%p"); c.spec_e.theta = D.d("spec%e%theta");
  c.spec_p.P = D.d("spec%p%p"); c.spec_p.theta = D.d("spec%p%theta");
  targ.A = D.d("targ%a"); targ.Z = D.d("targ%z"); targ.mass_amu = D.d("targ%mass_amu"); targ.mrec_amu = D.d("targ%mrec_amu");
  targ.rho = D.d("targ%rho"); targ.thick = D.d("targ%thick"); targ.angle = D.d("targ%angle"); targ.abundancy = D.d("targ%abundancy");
  targ.can = (int)D.i("targ%can"); targ.fr_pattern = (int)D.i("targ%fr_pattern"); targ.fr1 = D.d("targ%fr1"); targ.fr2 = D.d("targ%fr2");
  targ.xoffset = D.d("targ%xoffset"); targ.yoffset = D.d("targ%yoffset"); targ.zoffset = D.d("targ%zoffset");
  c.gen.xwid = D.d("gen%xwid"); c.gen.ywid = D.d("gen%ywid");
  c.spec_e.off_x = D.d("spec%e%offset%x"); c.spec_e.off_y = D.d("spec%e%offset%y"); c.spec_e.off_z = D.d("spec%e%offset%z");
  c.spec_e.off_xptar = D.d("spec%e%offset%xptar"); c.spec_e.off_yptar = D.d("spec%e%offset%yptar");
  c.spec_p.off_x = D.d("spec%p%offset%x"); c.spec_p.off_y = D.d("spec%p%offset%y"); c.spec_p.off_z = D.d("spec%p%offset%z");
  c.spec_p.off_xptar = D.d("spec%p%offset%xptar"); c.spec_p.off_yptar = D.d("spec%p%offset%yptar");
  simc_arm_cuts* sp[2] = {&c.SPedge_e, &c.SPedge_p};
  const char* an[2] = {"e", "p"};
  for (int k = 0; k < 2; ++k) {
    const std::string b = std::string("spedge%") + an[k] + "%";
    sp[k]->delta.min = D.d(b + "delta%min"); sp[k]->delta.max = D.d(b + "delta%max");
    sp[k]->yptar.min = D.d(b + "yptar%min"); sp[k]->yptar.max = D.d(b + "yptar%max");
    sp[k]->xptar.min = D.d(b + "xptar%min"); sp[k]->xptar.max = D.d(b + "xptar%max");
    sp[k]->z.min = -1.0e10; sp[k]->z.max = 1.0e10;                  // min_max_init, init.f:921-1190
  }
  c.doing_phsp = D.flag("doing_phsp"); c.doing_kaon = D.flag("doing_kaon"); c.doing_pion = D.flag("doing_pion");
  c.doing_delta = D.flag("doing_delta"); c.doing_semi = D.flag("doing_semi"); c.doing_rho = D.flag("doing_rho");
  c.doing_hplus = D.flag("doing_hplus", 1); c.doing_decay = D.flag("doing_decay"); c.do_fermi = D.flag("do_fermi");
  c.which_pion = (int)D.i("which_pion"); c.which_kaon = (int)D.i("which_kaon");
  c.doing_pizero = D.flag("doing_pizero"); c.pizero_ngamma = (int)D.i("pizero_ngamma"); c.drift_to_cal = D.d("drift_to_cal");
  c.using_tgt_field = D.flag("using_tgt_field"); c.targ_pol = D.d("targ_pol");
  c.targ_Bangle = D.d("targ_bangle") / degrad; c.targ_Bphi = D.d("targ_bphi") / degrad;          // dbase.f:454-455
  c.ctau = D.d("ctau"); c.transparency = D.d("transparency"); c.use_benhar_sf = D.flag("use_benhar_sf");
  c.hard_cuts = D.flag("hard_cuts"); c.using_rad = D.flag("using_rad"); c.use_expon = (int)D.i("use_expon");
  c.intcor_mode = (int)D.i("intcor_mode"); c.mc_smear = D.flag("mc_smear");
  c.using_Eloss = D.flag("using_eloss"); c.correct_Eloss = D.flag("correct_eloss"); c.correct_raster = D.flag("correct_raster");
  c.using_HMScoll = D.flag("using_hmscoll"); c.using_SHMScoll = D.flag("using_shmscoll");
  c.deForest_flag = (int)D.i("deforest_flag"); c.rad_flag = (int)D.i("rad_flag"); c.extrad_flag = (int)D.i("extrad_flag");
  c.using_Coulomb = D.flag("using_coulomb"); c.dE_edge_test = D.d("de_edge_test"); c.use_offshell_rad = D.flag("use_offshell_rad");
  c.Egamma_gen_max = D.d("egamma_gen_max");
  const long one_tail = D.i("one_tail"), spect_mode = D.i("spect_mode");
  c.cuts_Em.min = D.has("cuts%em%min") ? D.d("cuts%em%min") : -1.0e10;
  c.cuts_Em.max = D.has("cuts%em%max") ? D.d("cuts%em%max") : 1.0e10;
  c.cuts_Pm.min = -1.0e10; c.cuts_Pm.max = 1.0e10;

  // ---- dbase.f:123-222: reaction
  bool semipi = false, semika = false;
  if (c.doing_pion && c.doing_semi) { semipi = true; c.doing_pion = 0; }
  if (c.doing_kaon && c.doing_semi) { semika = true; c.doing_kaon = 0; }
  const long nA = std::lround(targ.A);
  if (c.doing_pion) {
    c.Mh = Mpi;
    if (c.doing_pizero) c.Mh = 134.9766;          // dbase.f:140
    c.doing_hydpi = nA == 1; c.doing_deutpi = nA == 2; c.doing_hepi = nA >= 3;
    if (c.which_pion >= 10) { c.doing_hydpi = 1; c.doing_deutpi = 0; c.doing_hepi = 0; }
  } else if (c.doing_kaon) {
    c.Mh = Mk;
    c.doing_hydkaon = nA == 1; c.doing_deutkaon = nA == 2; c.doing_hekaon = nA >= 3;
    if (c.which_kaon >= 10) { c.doing_hydkaon = 1; c.doing_deutkaon = 0; c.doing_hekaon = 0; }
  } else if (c.doing_delta) {
    c.Mh = Mp;
  } else if (c.doing_semi) {
    c.Mh = semipi ? Mpi : Mk;
    c.doing_semipi = semipi; c.doing_semika = semika;
    c.doing_hydsemi = nA == 1; c.doing_deutsemi = nA == 2;
    if (c.doing_hydsemi && c.do_fermi) c.do_fermi = 0;
  } else if (c.doing_rho) {
    c.Mh = Mrho;
  } else {
    c.Mh = Mp;
    c.doing_eep = 1;
    c.doing_hyd_elast = nA == 1; c.doing_deuterium = nA == 2; c.doing_heavy = nA >= 3;
  }
  c.Mh2 = c.Mh * c.Mh;
  if (c.doing_phsp) { c.rad_flag = 0; c.doing_eep = 0; c.doing_pion = 0; c.doing_kaon = 0; c.doing_delta = 0; c.doing_rho = 0; }
  // ---- dbase.f:226-252: kinematics
  c.dEbeam = c.Ebeam * c.dEbeam / 100.;
  c.spec_e.theta = std::fabs(c.spec_e.theta) / degrad;
  c.spec_e.cos_th = std::cos(c.spec_e.theta); c.spec_e.sin_th = std::sin(c.spec_e.theta);
  c.spec_p.theta = std::fabs(c.spec_p.theta) / degrad;
  c.spec_p.cos_th = std::cos(c.spec_p.theta); c.spec_p.sin_th = std::sin(c.spec_p.theta);
  auto phi_of = [&](int arm) {
    if (arm == 1 || arm == 3 || arm == 7) return 3 * pi / 2.;
    if (arm == 2 || arm == 4 || arm == 5 || arm == 6 || arm == 8) return pi / 2.;
    throw std::runtime_error("I dont know what phi should be for this arm");
  };
  c.spec_e.phi = phi_of(c.electron_arm);
  c.spec_p.phi = phi_of(c.hadron_arm);
  // ---- dbase.f:256-447: target and struck / recoil masses
  targ.N = targ.A - targ.Z;
  targ.M = targ.mass_amu * amu;
  targ.Mrec = targ.mrec_amu * amu;
  double Mrec_guess = 0;
  if (nA == 1) { targ.M = Mp; targ.Mrec = 0.; }
  else if (nA == 2) { targ.M = Md; }
  else {
    Mrec_guess = targ.M - Mp;
    if (std::fabs(targ.Mrec - Mrec_guess) > 100.) targ.Mrec = Mrec_guess;
  }
  // sign_hadron, dbase.f:296-423: the charge of the detected hadron (target-field tracking only)
  c.sign_hadron = 1.0;
  if (c.doing_semi || c.doing_rho) c.sign_hadron = c.doing_hplus ? 1.0 : -1.0;
  else if (c.doing_pion) c.sign_hadron = (c.which_pion == 1 || c.which_pion == 3 || c.which_pion == 11) ? -1.0 : 1.0;
  if (c.doing_eep) { targ.Mtar_struck = Mp; targ.Mrec_struck = 0.0; }
  else if (c.doing_delta) { targ.Mtar_struck = Mp; targ.Mrec_struck = Mpi; }
  else if (c.doing_semi) { targ.Mtar_struck = Mp; targ.Mrec_struck = Mp; }
  else if (c.doing_rho) { targ.Mtar_struck = Mp; targ.Mrec_struck = Mp; }
  else if (c.doing_pion && c.doing_pizero) {       // dbase.f:330-347
    switch (c.which_pion) {
      case 0: targ.Mtar_struck = Mp; targ.Mrec_struck = Mp; break;
      case 1: targ.Mtar_struck = Mn; targ.Mrec_struck = Mn; break;
      case 2: targ.Mtar_struck = Mp; targ.Mrec_struck = MDelta; break;
      case 3: targ.Mtar_struck = Mn; targ.Mrec_struck = MDelta; break;
      default: throw std::runtime_error("Bad value for which_pion");
    }
  } else if (c.doing_pion) {
    switch (c.which_pion) {
      case 0: targ.Mtar_struck = Mp; targ.Mrec_struck = Mn; break;
      case 1: targ.Mtar_struck = Mn; targ.Mrec_struck = Mp; break;
      case 2: targ.Mtar_struck = Mp; targ.Mrec_struck = MDelta; break;
      case 3: targ.Mtar_struck = Mp; targ.Mrec_struck = MDelta; break;
      case 10: targ.Mtar_struck = targ.M; targ.Mrec_struck = targ.Mrec; Mrec_guess = targ.M - Mp + Mn; break;
      case 11: targ.Mtar_struck = targ.M; targ.Mrec_struck = targ.Mrec; Mrec_guess = targ.M - Mn + Mp; break;
      default: throw std::runtime_error("Bad value for which_pion");
    }
    if (c.which_pion >= 10) {
      if (std::fabs(targ.Mrec_struck - Mrec_guess) > 100.) targ.Mrec_struck = Mrec_guess;
      targ.Mrec = 0.;
    }
  } else if (c.doing_kaon) {
    switch (c.which_kaon) {
      case 0: targ.Mtar_struck = Mp; targ.Mrec_struck = Mlambda; break;
      case 1: targ.Mtar_struck = Mp; targ.Mrec_struck = Msigma0; break;
      case 2: targ.Mtar_struck = Mn; targ.Mrec_struck = Msigma_minus; break;
      case 10: targ.Mtar_struck = targ.M; targ.Mrec_struck = targ.Mrec; Mrec_guess = targ.M - Mp + Mlambda; break;
      case 11: targ.Mtar_struck = targ.M; targ.Mrec_struck = targ.Mrec; Mrec_guess = targ.M - Mp + Msigma0; break;
      case 12: targ.Mtar_struck = targ.M; targ.Mrec_struck = targ.Mrec; Mrec_guess = targ.M - Mn + Msigma_minus; break;
      default: throw std::runtime_error("Bad value for which_kaon");
    }
    if (c.which_kaon >= 10) {
      if (std::fabs(targ.Mrec_struck - Mrec_guess) > 100.) targ.Mrec_struck = Mrec_guess;
      targ.Mrec = 0.;
    }
  }
  if (nA == 2) targ.Mrec = Mp + Mn - targ.Mtar_struck;
  targ.thick = targ.thick / 1000.;
  targ.length = targ.thick / targ.rho;
  targ.angle = targ.angle / degrad;
  if (targ.Z < 2.4) {
    if (std::fabs(targ.angle) > 0.0001) targ.angle = 0.0;
    if (targ.can != 1 && targ.can != 2 && targ.can != 3) throw std::runtime_error("bad targ.can value");
  }
  // ---- dbase.f:468-553: offsets, acceptance, simulate block
  c.spec_e.off_xptar /= 1000.; c.spec_e.off_yptar /= 1000.; c.spec_p.off_xptar /= 1000.; c.spec_p.off_yptar /= 1000.;
  if (c.SPedge_e.delta.min <= -100.0) c.SPedge_e.delta.min = -99.99;
  if (c.SPedge_p.delta.min <= -100.0) c.SPedge_p.delta.min = -99.99;
  for (int k = 0; k < 2; ++k) {
    sp[k]->yptar.min /= 1000.; sp[k]->yptar.max /= 1000.; sp[k]->xptar.min /= 1000.; sp[k]->xptar.max /= 1000.;
  }
  c.doing_tail[0] = one_tail == 0 || one_tail == 1 || one_tail == -2 || one_tail == -3;
  c.doing_tail[1] = one_tail == 0 || one_tail == 2 || one_tail == -3 || one_tail == -1;
  c.doing_tail[2] = one_tail == 0 || one_tail == 3 || one_tail == -1 || one_tail == -2;
  if (!c.using_rad) c.doing_tail[0] = c.doing_tail[1] = c.doing_tail[2] = 0;
  c.hardwired_rad = c.Egamma_gen_max > 0.01;
  c.using_E_arm_montecarlo = spect_mode != 1 && spect_mode != -1;
  c.using_P_arm_montecarlo = spect_mode != 1 && spect_mode != -2;
  if (c.doing_pion || c.doing_kaon || c.doing_delta || (c.cuts_Em.min == c.cuts_Em.max)) { c.cuts_Em.min = -1.e6; c.cuts_Em.max = 1.e6; }
  if (c.correct_Eloss && !c.using_Eloss) c.correct_Eloss = 0;
  if (std::lround(targ.Z) == 1) c.using_Coulomb = 0;

  // ---- target_init, init.f:1-87
  TargX X;
  std::memset(&X, 0, sizeof(X));
  targ.L1 = std::log(184.15) - std::log(targ.Z) / 3.0;
  targ.L2 = std::log(1194.) - 2. * std::log(targ.Z) / 3.0;
  if (targ.Z == 1) { targ.L1 = 5.31; targ.L2 = 6.144; }
  {
    const double za2 = (targ.Z * alpha) * (targ.Z * alpha);
    const double fc = za2 * (1.202 + za2 * (-1.0369 + za2 * 1.008 / (za2 + 1)));
    if (nA == 1) targ.X0 = 61.28;
    else if (nA == 2) targ.X0 = 122.4;
    else if (nA == 4) targ.X0 = 94.32;
    else targ.X0 = 716.405 * targ.A / targ.Z / (targ.Z * (targ.L1 - fc) + targ.L2);
    targ.X0_cm = targ.X0 / targ.rho;
  }
  trip(c, 1, 0.0, c.Ebeam, 0.0, X.Eloss_ave[0], X.teff_ave[0], Me, 4);
  trip(c, 2, 0.0, c.spec_e.P, c.spec_e.theta, X.Eloss_ave[1], X.teff_ave[1], Me, 4);
  trip(c, 3, 0.0, std::sqrt(c.spec_p.P * c.spec_p.P + c.Mh2), c.spec_p.theta, X.Eloss_ave[2], X.teff_ave[2], std::sqrt(c.Mh2), 4);
  if (!c.using_Eloss) X.Eloss_ave[0] = X.Eloss_ave[1] = X.Eloss_ave[2] = 0.0;
  if (c.using_Coulomb) {
    targ.Coulomb_ave = 0.75 * 1.5 * (targ.Z - 1.) * alpha * hbarc / (1.1 * std::pow(targ.A, 1. / 3.) + 0.86 * std::pow(targ.A, -1. / 3.));
    targ.Coulomb_constant = targ.Coulomb_ave;
    targ.Coulomb_min = targ.Coulomb_constant;
    targ.Coulomb_max = targ.Coulomb_constant;
  }

  // ---- limits_init, init.f:91-517
  auto slop_of = [](int arm, double* s3) {
    if (arm == 2) { s3[0] = 1.0; s3[1] = 0.008; s3[2] = 0.008; }
    else if (arm == 1 || arm == 3 || arm == 4 || arm == 5 || arm == 6) { s3[0] = 0.5; s3[1] = 0.005; s3[2] = 0.005; }
  };
  if (c.using_E_arm_montecarlo) slop_of(c.electron_arm, c.slop_MC_e_used);
  if (c.using_P_arm_montecarlo) slop_of(c.hadron_arm, c.slop_MC_p_used);
  c.SPedge_e.delta.min -= c.slop_MC_e_used[0]; c.SPedge_e.delta.max += c.slop_MC_e_used[0];
  c.SPedge_e.yptar.min -= c.slop_MC_e_used[1]; c.SPedge_e.yptar.max += c.slop_MC_e_used[1];
  c.SPedge_e.xptar.min -= c.slop_MC_e_used[2]; c.SPedge_e.xptar.max += c.slop_MC_e_used[2];
  c.SPedge_p.delta.min -= c.slop_MC_p_used[0]; c.SPedge_p.delta.max += c.slop_MC_p_used[0];
  c.SPedge_p.yptar.min -= c.slop_MC_p_used[1]; c.SPedge_p.yptar.max += c.slop_MC_p_used[1];
  c.SPedge_p.xptar.min -= c.slop_MC_p_used[2]; c.SPedge_p.xptar.max += c.slop_MC_p_used[2];
  simc_edge& edge = c.edge;
  simc_edge& V = c.VERTEXedge;
  // min_max_init: wide-open defaults of everything limits_init does not assign
  for (simc_edge* e : {&edge, &V}) {
    simc_cut* all[] = {&e->e.E, &e->e.yptar, &e->e.xptar, &e->p.E, &e->p.yptar, &e->p.xptar, &e->Em, &e->Pm, &e->Mrec, &e->Trec, &e->Trec_struck};
    for (simc_cut* q : all) { q->min = -1.0e10; q->max = 1.0e10; }
  }
  edge.e.E.min = (1. + c.SPedge_e.delta.min / 100.) * c.spec_e.P + targ.Coulomb_min - c.dE_edge_test;
  edge.e.E.max = (1. + c.SPedge_e.delta.max / 100.) * c.spec_e.P + targ.Coulomb_max + c.dE_edge_test;
  Lim pp;
  pp.min = (1. + c.SPedge_p.delta.min / 100.) * c.spec_p.P - c.dE_edge_test;
  pp.max = (1. + c.SPedge_p.delta.max / 100.) * c.spec_p.P + c.dE_edge_test;
  pp.min = std::max(0.001e0, pp.min);
  edge.p.E.min = std::sqrt(pp.min * pp.min + c.Mh2);
  edge.p.E.max = std::sqrt(pp.max * pp.max + c.Mh2);
  Lim the, thp, z;
  the.max = std::acos((c.spec_e.cos_th - c.spec_e.sin_th * c.SPedge_e.yptar.max) /
                      std::sqrt(1. + c.SPedge_e.yptar.max * c.SPedge_e.yptar.max + c.SPedge_e.xptar.max * c.SPedge_e.xptar.max));
  the.min = std::acos((c.spec_e.cos_th - c.spec_e.sin_th * c.SPedge_e.yptar.min) / std::sqrt(1. + c.SPedge_e.yptar.min * c.SPedge_e.yptar.min));
  thp.max = std::acos((c.spec_p.cos_th - c.spec_p.sin_th * c.SPedge_p.yptar.max) /
                      std::sqrt(1. + c.SPedge_p.yptar.max * c.SPedge_p.yptar.max + c.SPedge_p.xptar.max * c.SPedge_p.xptar.max));
  thp.min = std::acos((c.spec_p.cos_th - c.spec_p.sin_th * c.SPedge_p.yptar.min) / std::sqrt(1. + c.SPedge_p.yptar.min * c.SPedge_p.yptar.min));
  z.min = -0.5 * targ.length;
  z.max = 0.5 * targ.length;
  Lim pe = {edge.e.E.min, edge.e.E.max};
  extreme_trip(c, X, c.Ebeam, the, thp, pe, pp, z, c.Mh);
  if (!c.using_Eloss) for (int i = 0; i < 3; ++i) { X.Eloss_min[i] = 0.0; X.Eloss_max[i] = 0.0; }
  if (!c.mc_smear) X.musc_max[0] = X.musc_max[1] = X.musc_max[2] = 0.;
  edge.e.E.min += X.Eloss_min[1]; edge.e.E.max += X.Eloss_max[1];
  edge.p.E.min += X.Eloss_min[2]; edge.p.E.max += X.Eloss_max[2];
  edge.e.yptar.min = c.SPedge_e.yptar.min - X.musc_max[1]; edge.e.yptar.max = c.SPedge_e.yptar.max + X.musc_max[1];
  edge.e.xptar.min = c.SPedge_e.xptar.min - X.musc_max[1]; edge.e.xptar.max = c.SPedge_e.xptar.max + X.musc_max[1];
  edge.p.yptar.min = c.SPedge_p.yptar.min - X.musc_max[2]; edge.p.yptar.max = c.SPedge_p.yptar.max + X.musc_max[2];
  edge.p.xptar.min = c.SPedge_p.xptar.min - X.musc_max[2]; edge.p.xptar.max = c.SPedge_p.xptar.max + X.musc_max[2];
  c.Ebeam_vertex_ave = c.Ebeam + targ.Coulomb_ave - X.Eloss_ave[0];
  const double Ebeam_max = c.Ebeam + c.dEbeam / 2. - X.Eloss_min[0] + targ.Coulomb_max;
  const double Ebeam_min = c.Ebeam - c.dEbeam / 2. - X.Eloss_max[0] + targ.Coulomb_min;
  const double slop_Coulomb = targ.Coulomb_max - targ.Coulomb_ave;
  double slop_Ebeam = c.dEbeam / 2. + slop_Coulomb;
  double slop_Ee = c.slop_MC_e_used[0] / 100. * c.spec_e.P + slop_Coulomb;
  const double r = std::sqrt(edge.p.E.max * edge.p.E.max - c.Mh2);
  const double rs = r + c.slop_MC_p_used[0] / 100. * c.spec_p.P;
  double slop_Ep = std::sqrt(rs * rs + c.Mh2) - edge.p.E.max;
  slop_Ebeam += (X.Eloss_max[0] - X.Eloss_min[0]);
  slop_Ee += (X.Eloss_max[1] - X.Eloss_min[1]);
  slop_Ep += (X.Eloss_max[2] - X.Eloss_min[2]);
  if (c.doing_heavy) {
    const double slop_Em = slop_Ebeam + slop_Ee + slop_Ep + c.dE_edge_test;
    edge.Em.min = c.cuts_Em.min - slop_Em;
    edge.Em.max = c.cuts_Em.max + slop_Em;
    edge.Em.min = std::max(0.e0, edge.Em.min);
  }
  const double pm_theory = D.d("x_pm_theory_absmax"), e_fermi = D.d("x_e_fermi"), pval_last = D.d("x_pval_last"), emval_last = D.d("x_emval_last");
  const bool hyd_meson = c.doing_hydpi || c.doing_hydkaon || (c.doing_delta && nA == 1) || (c.doing_rho && nA == 1);
  const bool deut_meson = c.doing_deutpi || c.doing_deutkaon || (c.doing_delta && nA == 2) || (c.doing_rho && nA == 2);
  const bool he_meson = c.doing_hepi || c.doing_hekaon || (c.doing_delta && nA >= 3) || (c.doing_rho && nA == 3);
  if (c.doing_hyd_elast) { V.Em.min = V.Em.max = V.Pm.min = V.Pm.max = 0.0; }
  else if (c.doing_deuterium) { V.Em.min = V.Em.max = Mp + Mn - targ.M; V.Pm.min = 0.0; V.Pm.max = pm_theory; }
  else if (c.doing_heavy) {
    V.Pm.min = 0.0;
    V.Pm.max = c.use_benhar_sf ? 790.0 : std::max(0.0, pm_theory);
    V.Em.min = e_fermi;
    V.Em.max = 1000.;
  } else if (hyd_meson) { V.Em.min = V.Em.max = V.Pm.min = V.Pm.max = 0.0; }
  else if (deut_meson) { V.Em.min = V.Em.max = Mp + Mn - targ.M; V.Pm.min = 0.0; V.Pm.max = pval_last; }
  else if (he_meson) { V.Em.min = targ.Mtar_struck + targ.Mrec - targ.M; V.Em.max = emval_last; V.Pm.min = 0.0; V.Pm.max = pval_last; }
  else if (c.doing_semi) { V.Em.min = V.Em.max = V.Pm.min = V.Pm.max = 0.0; }
  if (c.doing_hyd_elast || hyd_meson || c.doing_semi) { V.Mrec.min = V.Mrec.max = V.Trec.min = V.Trec.max = 0.0; }
  else {
    V.Mrec.min = targ.M - targ.Mtar_struck + V.Em.min;
    V.Mrec.max = targ.M - targ.Mtar_struck + V.Em.max;
    V.Trec.min = std::sqrt(V.Mrec.max * V.Mrec.max + V.Pm.min * V.Pm.min) - V.Mrec.max;
    V.Trec.max = std::sqrt(V.Mrec.min * V.Mrec.min + V.Pm.max * V.Pm.max) - V.Mrec.min;
  }
  if (c.doing_eep || c.doing_semi) { V.Trec_struck.min = 0.; V.Trec_struck.max = 0.; }
  else {
    V.Trec_struck.min = 0.;
    V.Trec_struck.max = Ebeam_max + targ.Mtar_struck - targ.Mrec_struck - edge.e.E.min - edge.p.E.min - V.Em.min - V.Trec.min;
  }
  c.Egamma_tot_max = Ebeam_max + targ.Mtar_struck - targ.Mrec_struck - edge.e.E.min - edge.p.E.min - V.Em.min - V.Trec.min - V.Trec_struck.min;
  if (c.doing_heavy) {
    const double t2 = (edge.Em.max - V.Em.min) + (V.Trec.max - V.Trec.min);
    c.Egamma_tot_max = std::min(c.Egamma_tot_max, t2);
  }
  if (c.hardwired_rad) c.Egamma_tot_max = c.Egamma_gen_max;
  if (!c.using_rad) c.Egamma_tot_max = 0.0;
  if (c.doing_tail[0]) c.Egamma1_max = c.Egamma_tot_max;
  if (c.doing_tail[1]) c.Egamma2_max = c.Egamma_tot_max;
  if (c.doing_tail[2]) c.Egamma3_max = c.Egamma_tot_max;
  if (c.doing_heavy) {
    V.Em.min = std::max(V.Em.min, edge.Em.min - c.Egamma_tot_max);
    V.Em.max = std::min(V.Em.max, edge.Em.max);
  }
  simc_gen_limits& gen = c.gen;
  gen.Trec.min = -1.0e10; gen.Trec.max = 1.0e10;          // min_max_init (init.f:921-1190); limits_init never narrows it
  if (c.doing_hyd_elast) { gen.sumEgen.min = 0.0; gen.sumEgen.max = 0.0; }
  else if (c.doing_heavy) {
    gen.sumEgen.max = Ebeam_max + targ.Mtar_struck - V.Trec.min - V.Em.min;
    gen.sumEgen.min = Ebeam_min + targ.Mtar_struck - V.Trec.max - V.Em.max - c.Egamma1_max;
    gen.sumEgen.max = std::min(gen.sumEgen.max, edge.e.E.max + edge.p.E.max + c.Egamma_tot_max);
    gen.sumEgen.min = std::max(gen.sumEgen.min, edge.e.E.min + edge.p.E.min);
  } else if (c.doing_semi) {
    gen.sumEgen.max = Ebeam_max + targ.Mtar_struck - targ.Mrec_struck;
    gen.sumEgen.min = edge.e.E.min + edge.p.E.min;
  } else {
    gen.sumEgen.max = Ebeam_max + targ.Mtar_struck - targ.Mrec_struck - edge.p.E.min - V.Em.min - V.Trec.min - V.Trec_struck.min;
    gen.sumEgen.min = Ebeam_min + targ.Mtar_struck - targ.Mrec_struck - edge.p.E.max - V.Em.max - V.Trec.max - V.Trec_struck.max - c.Egamma_tot_max;
    gen.sumEgen.max = std::min(gen.sumEgen.max, edge.e.E.max + c.Egamma2_max);
    gen.sumEgen.min = std::max(gen.sumEgen.min, edge.e.E.min);
  }
  gen.sumEgen.min -= c.dE_edge_test;
  gen.sumEgen.max += c.dE_edge_test;
  gen.sumEgen.min = std::max(0.e0, gen.sumEgen.min);
  if (c.doing_hyd_elast) { gen.e.E.min = edge.e.E.min; gen.e.E.max = edge.e.E.max + c.Egamma2_max; }
  else if (c.doing_deuterium || c.doing_pion || c.doing_kaon || c.doing_rho || c.doing_delta) { gen.e.E.min = gen.sumEgen.min; gen.e.E.max = gen.sumEgen.max; }
  else if (c.doing_heavy || c.doing_semi) { gen.e.E.min = gen.sumEgen.min - edge.p.E.max - c.Egamma3_max; gen.e.E.max = gen.sumEgen.max - edge.p.E.min; }
  gen.e.E.min = std::max(gen.e.E.min, edge.e.E.min);
  gen.e.E.max = std::min(gen.e.E.max, edge.e.E.max + c.Egamma2_max);
  gen.e.delta.min = (gen.e.E.min / c.spec_e.P - 1.) * 100.;
  gen.e.delta.max = (gen.e.E.max / c.spec_e.P - 1.) * 100.;
  gen.e.yptar = edge.e.yptar; gen.e.xptar = edge.e.xptar;
  if (c.doing_hyd_elast || c.doing_deuterium || c.doing_pion || c.doing_kaon || c.doing_rho || c.doing_delta) {
    gen.p.E.min = edge.p.E.min; gen.p.E.max = edge.p.E.max + c.Egamma3_max;
  } else if (c.doing_heavy || c.doing_semi) {
    gen.p.E.min = gen.sumEgen.min - edge.e.E.max - c.Egamma2_max; gen.p.E.max = gen.sumEgen.max - edge.e.E.min;
  }
  gen.p.E.min = std::max(gen.p.E.min, edge.p.E.min);
  gen.p.E.max = std::min(gen.p.E.max, edge.p.E.max + c.Egamma3_max);
  gen.p.delta.min = (std::sqrt(gen.p.E.min * gen.p.E.min - c.Mh2) / c.spec_p.P - 1.) * 100.;
  gen.p.delta.max = (std::sqrt(gen.p.E.max * gen.p.E.max - c.Mh2) / c.spec_p.P - 1.) * 100.;
  gen.p.yptar = edge.p.yptar; gen.p.xptar = edge.p.xptar;
  // histogram axes, init.f:519-569 (nHbins = 50); the three sets share them
  const double nb = (double)(float)SIMC_NHIST;
  simc_axis ax[SIMC_H_PER_SET];
  ax[SIMC_H_E_DELTA] = {gen.e.delta.min, (gen.e.delta.max - gen.e.delta.min) / nb};
  ax[SIMC_H_E_YPTAR] = {gen.e.yptar.min, (gen.e.yptar.max - gen.e.yptar.min) / nb};
  ax[SIMC_H_E_XPTAR] = {-gen.e.xptar.max, (gen.e.xptar.max - gen.e.xptar.min) / nb};
  ax[SIMC_H_P_DELTA] = {gen.p.delta.min, (gen.p.delta.max - gen.p.delta.min) / nb};
  ax[SIMC_H_P_YPTAR] = {gen.p.yptar.min, (gen.p.yptar.max - gen.p.yptar.min) / nb};
  ax[SIMC_H_P_XPTAR] = {-gen.p.xptar.max, (gen.p.xptar.max - gen.p.xptar.min) / nb};
  ax[SIMC_H_EM] = {V.Em.min, (std::max(100.e0, V.Em.max) - V.Em.min) / nb};
  ax[SIMC_H_PM] = {V.Pm.min, (std::max(100.e0, V.Pm.max) - V.Pm.min) / nb};
  for (int s = 0; s < 3; ++s) for (int k = 0; k < SIMC_H_PER_SET; ++k) c.hist_axis[s][k] = ax[k];

  // ---- radc_init, init.f:576-651
  if (c.extrad_flag == 0) c.extrad_flag = c.rad_flag == 0 ? 3 : 1;
  c.etatzai = (12.0 + (targ.Z + 1.) / (targ.Z * targ.L1 + targ.L2)) / 9.0;
}

}  // namespace simc_oracle

extern "C" void oracle_set_error(const char* m);
extern "C" int oracle_init_from_kv(const char* text, simc_run_config* out) {
  try {
    simc_oracle::init_from_kv(text, *out);
    return 0;
  } catch (const std::exception& e) {
    oracle_set_error(e.what());
    return -1;
  }
}
