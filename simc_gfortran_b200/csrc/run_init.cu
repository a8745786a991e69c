// Host-side run initialisation: CTP deck reader and the one-time setup the reference does
// before the loop -- dbase_read post-processing (dbase.f:119-553), target_init (init.f:1-87),
// limits_init (init.f:91-572) with extreme_trip_thru_target (target.f:310-544), radc_init
// (init.f:576-651) -- producing the simc_run_config the loop consumes.  In a drop-in
// deployment the Fortran driver fills simc_run_config from its COMMON blocks instead
// (INTEGRATION.md); this file exists so the library can run a deck without any Fortran.
// Host code only (compiled by nvcc for the shared SIMC_HD helpers of target.cuh).
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include "../../include/simc_b200.h"
#include "event.cuh"
#include "tables_host.h"

namespace simc {
namespace {

// ---- CTP "parm" deck: `begin parm <name>` ... `key = value ; comment` ... `end parm` -------------
// Keys are matched case-insensitively; the pre-gfortran dotted form (targ.A) is accepted as an
// alias of the % form (infiles/convert_inputfile.sh does that conversion for the reference).
struct Deck {
  std::map<std::string, std::string> kv;
  static std::string norm(std::string k) {
    std::string r;
    for (char c : k) {
      if (std::isspace((unsigned char)c)) continue;
      r.push_back(c == '.' ? '%' : (char)std::tolower((unsigned char)c));
    }
    return r;
  }
  void load(const std::string& path) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("Loading problem!  cannot open deck " + path);
    std::string line;
    bool in_parm = false;
    while (std::getline(in, line)) {
      const size_t sc = line.find(';');
      if (sc != std::string::npos) line.erase(sc);
      std::string t = line;
      t.erase(0, t.find_first_not_of(" \t\r"));
      if (t.empty()) continue;
      std::string low = t;
      std::transform(low.begin(), low.end(), low.begin(), [](unsigned char c) { return (char)std::tolower(c); });
      if (low.compare(0, 10, "begin parm") == 0) { in_parm = true; continue; }
      if (low.compare(0, 8, "end parm") == 0) { in_parm = false; continue; }
      if (!in_parm) continue;
      const size_t eq = t.find('=');
      if (eq == std::string::npos) continue;
      std::string key = norm(t.substr(0, eq));
      std::string val = t.substr(eq + 1);
      val.erase(0, val.find_first_not_of(" \t"));
      while (!val.empty() && std::isspace((unsigned char)val.back())) val.pop_back();
      if (!val.empty() && val.front() == '\'') {
        const size_t q = val.find('\'', 1);
        val = val.substr(1, q == std::string::npos ? std::string::npos : q - 1);
      }
      kv[key] = val;
    }
  }
  bool has(const char* k) const { return kv.count(norm(k)) != 0; }
  double d(const char* k, double def = 0.0) const {
    auto it = kv.find(norm(k));
    if (it == kv.end() || it->second.empty()) return def;
    std::string v = it->second;
    for (char& c : v) if (c == 'd' || c == 'D') c = 'e';
    return std::strtod(v.c_str(), nullptr);
  }
  int i(const char* k, int def = 0) const { return (int)std::lround(d(k, (double)def)); }
  std::string s(const char* k) const { auto it = kv.find(norm(k)); return it == kv.end() ? std::string() : it->second; }
};

constexpr double Me = 0.51099906, Mp = 938.27231, Mn = 939.56563, Mpi = 139.57018, Mpi0 = 134.9766, Mk = 493.677,
                 Md = 1875.613, amu = 931.49432, hbarc = 197.327053, pi = 3.141592653589793, alpha = 1. / 137.0359895,
                 degrad = 180. / pi;

struct TargetExt {          // target_info members only the init needs (target.inc:44-47)
  double Eloss_ave[3], Eloss_min[3], Eloss_max[3], teff_ave[3], teff_min[3], teff_max[3], musc_max[3];
};

void trip(const simc_run_config& c, int narm, double zpos, double energy, double theta, double mass, int typeflag,
          double& Eloss, double& radlen) {
  const MatTable mt = make_mat_table(c.targ);
  trip_thru_target_fixed(c.targ, mt, narm, narm == 2 ? c.electron_arm : c.hadron_arm, zpos, energy, theta, mass, typeflag,
                         Eloss, radlen);
}

// target.f:310-544
void extreme_trip_thru_target(const simc_run_config& c, TargetExt& x, double ebeam, simc_cut the, simc_cut thp,
                              simc_cut pe, simc_cut pp, simc_cut z, double m) {
  const simc_target& targ = c.targ;
  const double inch_cm = 2.54, target_pi = 3.14159265358979;
  const bool liquid = targ.Z < 2.4;
  trip(c, 1, z.max, ebeam, 0.0, Me, 3, x.Eloss_max[0], x.teff_max[0]);
  trip(c, 1, z.min, ebeam, 0.0, Me, 2, x.Eloss_min[0], x.teff_min[0]);
  double energymin = pe.min, energymax = pe.max, E1, E2, E3, E4, t1, t2, t3, t4, zz, th1, th2;
  auto corner = [&](const simc_cut& th, double& zz_, double& th1_, double& th2_) {
    double th_corner_max;
    if (z.max >= targ.length / 2.) th_corner_max = target_pi / 2.;
    else th_corner_max = std::atan(1.25 * inch_cm / (targ.length / 2. - z.max));
    const double th_corner_min = std::atan(1.25 * inch_cm / (targ.length / 2. - z.min));
    if (th_corner_min <= th.max && th_corner_max >= th.min) {
      const double th_corner = std::max(th_corner_min, th.min);
      zz_ = targ.length / 2. - 1.25 * inch_cm / std::tan(th_corner);
      th1_ = th_corner - .0001;
      th2_ = th_corner + .0001;
    } else {
      zz_ = z.min; th1_ = th.min; th2_ = th.max;
    }
  };
  if (!liquid) {
    trip(c, 2, z.min, energymax, the.max, Me, 3, x.Eloss_max[1], x.teff_max[1]);
    trip(c, 2, z.max, energymin, the.min, Me, 2, x.Eloss_min[1], x.teff_min[1]);
  } else {
    if (targ.can == 1) {
      corner(the, zz, th1, th2);
      trip(c, 2, zz, energymax, th1, Me, 3, E1, t1);
      trip(c, 2, zz, energymax, th2, Me, 3, E2, t2);
      x.Eloss_max[1] = std::max(E1, E2);
      x.teff_max[1] = std::max(t1, t2);
    } else {
      zz = -(targ.length / 2.) / std::tan(the.min);
      zz = std::max(zz, (-targ.length / 2.));
      trip(c, 2, zz, energymax, the.min, Me, 3, x.Eloss_max[1], x.teff_max[1]);
    }
    x.Eloss_min[1] = 1.e10;
    for (int i = 0; i <= 3; ++i) {
      trip(c, 2, z.min + (i / 2) * (z.max - z.min), energymin, the.min + (i % 2) * (the.max - the.min), Me, 2, E1, t1);
      if (E1 < x.Eloss_min[1]) { x.Eloss_min[1] = E1; x.teff_min[1] = t1; }
    }
  }
  energymin = std::sqrt(pp.min * pp.min + m * m);
  energymax = std::sqrt(pp.max * pp.max + m * m);
  if (!liquid) {
    trip(c, 3, z.min, energymin, thp.max, m, 3, E1, t1);
    trip(c, 3, z.min, energymax, thp.max, m, 3, E2, t2);
    x.Eloss_max[2] = std::max(E1, E2); x.teff_max[2] = std::max(t1, t2);
    trip(c, 3, z.max, energymin, thp.min, m, 2, E1, t1);
    trip(c, 3, z.max, energymax, thp.min, m, 2, E2, t2);
    x.Eloss_min[2] = std::min(E1, E2); x.teff_min[2] = std::min(t1, t2);
  } else {
    if (targ.can == 1) {
      corner(thp, zz, th1, th2);
      trip(c, 3, zz, energymin, th1, m, 3, E1, t1);
      trip(c, 3, zz, energymin, th2, m, 3, E2, t2);
      trip(c, 3, zz, energymax, th1, m, 3, E3, t3);
      trip(c, 3, zz, energymax, th2, m, 3, E4, t4);
      x.Eloss_max[2] = std::max(std::max(E1, E2), std::max(E3, E4));
      x.teff_max[2] = std::max(std::max(t1, t2), std::max(t3, t4));
    } else {
      zz = -(targ.length / 2.) / std::tan(the.min);       // the%min as written (target.f:513)
      zz = std::max(zz, (-targ.length / 2.));
      trip(c, 3, zz, energymin, thp.min, m, 3, E1, t1);
      trip(c, 3, zz, energymax, thp.min, m, 3, E2, t2);
      x.Eloss_max[2] = std::max(E1, E2); x.teff_max[2] = std::max(t1, t2);
    }
    x.Eloss_min[2] = 1.e10;
    for (int i = 0; i <= 3; ++i) {
      const double zi = z.min + (i / 2) * (z.max - z.min), thi = thp.min + (i % 2) * (thp.max - thp.min);
      trip(c, 3, zi, energymin, thi, m, 2, E1, t1);
      if (E1 < x.Eloss_min[2]) { x.Eloss_min[2] = E1; x.teff_min[2] = t1; zz = zi; th1 = thi; }
    }
    trip(c, 3, zz, energymax, th1, m, 2, E1, t1);
    x.Eloss_min[2] = std::min(x.Eloss_min[2], E1);
  }
  // extreme multiple scattering, target.f:528-544 + target_musc entry (target.f:569-577)
  auto ext_musc = [](double p, double beta, double teff) {
    const double ts = 13.6 / p / beta * std::sqrt(teff) * (1 + 0.088 * std::log10(teff / (beta * beta)));
    return ts * 3.5;
  };
  x.musc_max[0] = ext_musc(ebeam, 1., x.teff_max[0]);
  x.musc_max[1] = ext_musc(pe.min, 1., x.teff_max[1]);
  const double betap_min = pp.min / std::sqrt(pp.min * pp.min + m * m);
  x.musc_max[2] = ext_musc(pp.min, betap_min, x.teff_max[2]);
}

void set_axis(simc_axis& a, double mn, double bin) { a.min = mn; a.bin = bin; }

}  // namespace

void config_from_deck(const std::string& path, const std::string& extra_dir, const std::string& data_dir,
                      simc_run_config& c, int* ngen, double* charge_mC, simc_report_info* info = nullptr) {
  double slop_total_Em_used = 0.0;
  Deck D;
  D.load(path);
  const std::string extra = D.s("extra_dbase_file");
  if (!extra.empty()) {
    std::string p = extra;
    if (p.find('/') == std::string::npos) p = (extra_dir.empty() ? std::string("") : extra_dir + "/") + p;
    if (p.find('.', p.find_last_of('/') == std::string::npos ? 0 : p.find_last_of('/')) == std::string::npos) p += ".inp";
    D.load(p);
  }
  std::memset(&c, 0, sizeof(c));
  c.abi_version = SIMC_B200_ABI_VERSION;
  if (ngen) *ngen = D.i("ngen");
  if (charge_mC) *charge_mC = D.d("EXPER%charge");
  // ---- regallvars + convert_to_logical, dbase.f:967-1165
  c.mc_smear = D.i("mc_smear") > 0;
  c.electron_arm = D.i("electron_arm");
  c.hadron_arm = D.i("hadron_arm");
  c.transparency = D.d("transparency");
  c.use_benhar_sf = D.i("use_benhar_sf") > 0;
  c.Ebeam = D.d("Ebeam");
  c.dEbeam = D.d("dEbeam");
  c.doing_kaon = D.i("doing_kaon") > 0; c.which_kaon = D.i("which_kaon");
  c.doing_pion = D.i("doing_pion") > 0; c.which_pion = D.i("which_pion");
  c.doing_delta = D.i("doing_delta") > 0;
  c.doing_semi = D.i("doing_semi") > 0;
  c.doing_hplus = D.i("doing_hplus", 1) > 0;
  c.doing_rho = D.i("doing_rho") > 0;
  c.doing_decay = D.i("doing_decay") > 0;
  c.doing_pizero = D.i("doing_pizero") > 0; c.pizero_ngamma = D.i("pizero_ngamma"); c.drift_to_cal = D.d("drift_to_cal");
  c.doing_phsp = D.i("doing_phsp") > 0;
  c.ctau = D.d("ctau");
  simc_target& targ = c.targ;
  targ.A = D.d("targ%A"); targ.Z = D.d("targ%Z"); targ.mass_amu = D.d("targ%mass_amu"); targ.mrec_amu = D.d("targ%mrec_amu");
  targ.rho = D.d("targ%rho"); targ.thick = D.d("targ%thick"); targ.xoffset = D.d("targ%xoffset");
  targ.yoffset = D.d("targ%yoffset"); targ.fr_pattern = D.i("targ%fr_pattern"); targ.fr1 = D.d("targ%fr1");
  targ.fr2 = D.d("targ%fr2"); targ.zoffset = D.d("targ%zoffset"); targ.angle = D.d("targ%angle");
  targ.abundancy = D.d("targ%abundancy"); targ.can = D.i("targ%can");
  c.spec_e.P = D.d("spec%e%P"); c.spec_e.theta = D.d("spec%e%theta");
  c.spec_p.P = D.d("spec%p%P"); c.spec_p.theta = D.d("spec%p%theta");
  c.gen.ywid = D.d("gen%ywid"); c.gen.xwid = D.d("gen%xwid");
  c.spec_e.off_x = D.d("spec%e%offset%x"); c.spec_e.off_y = D.d("spec%e%offset%y"); c.spec_e.off_z = D.d("spec%e%offset%z");
  c.spec_e.off_xptar = D.d("spec%e%offset%xptar"); c.spec_e.off_yptar = D.d("spec%e%offset%yptar");
  c.spec_p.off_x = D.d("spec%p%offset%x"); c.spec_p.off_y = D.d("spec%p%offset%y"); c.spec_p.off_z = D.d("spec%p%offset%z");
  c.spec_p.off_xptar = D.d("spec%p%offset%xptar"); c.spec_p.off_yptar = D.d("spec%p%offset%yptar");
  c.use_expon = D.i("use_expon");
  const int one_tail = D.i("one_tail");
  c.intcor_mode = D.i("intcor_mode");
  c.hard_cuts = D.i("hard_cuts") > 0;
  c.using_rad = D.i("using_rad") > 0;
  const int spect_mode = D.i("spect_mode");
  // min_max_init defaults (init.f:921-1190) for what the deck does not set
  c.cuts_Em.min = D.has("cuts%Em%min") ? D.d("cuts%Em%min") : -1.0e10;
  c.cuts_Em.max = D.has("cuts%Em%max") ? D.d("cuts%Em%max") : 1.0e10;
  c.cuts_Pm.min = -1.0e10; c.cuts_Pm.max = 1.0e10;
  c.using_Eloss = D.i("using_Eloss") > 0;
  c.correct_Eloss = D.i("correct_Eloss") > 0;
  c.correct_raster = D.i("correct_raster") > 0;
  c.using_HMScoll = D.i("using_HMScoll") > 0;
  c.using_SHMScoll = D.i("using_SHMScoll") > 0;
  const int deForest_flag = D.i("deForest_flag");
  c.rad_flag = D.i("rad_flag");
  c.extrad_flag = D.i("extrad_flag");
  c.using_Coulomb = D.i("using_Coulomb") > 0;
  c.dE_edge_test = D.d("dE_edge_test");
  c.use_offshell_rad = D.i("use_offshell_rad") > 0;
  c.Egamma_gen_max = D.d("Egamma_gen_max");
  c.using_tgt_field = D.i("using_tgt_field") > 0; c.targ_pol = D.d("targ_pol");
  c.targ_Bangle = D.d("targ_Bangle") / degrad; c.targ_Bphi = D.d("targ_Bphi") / degrad;        // dbase.f:454-455
  auto spedge = [&](simc_arm_cuts& a, const char* arm) {
    auto key = [&](const char* q, const char* mm) { return std::string("SPedge%") + arm + "%" + q + "%" + mm; };
    a.delta.min = D.d(key("delta", "min").c_str()); a.delta.max = D.d(key("delta", "max").c_str());
    a.yptar.min = D.d(key("yptar", "min").c_str()); a.yptar.max = D.d(key("yptar", "max").c_str());
    a.xptar.min = D.d(key("xptar", "min").c_str()); a.xptar.max = D.d(key("xptar", "max").c_str());
    a.z.min = -1.0e10; a.z.max = 1.0e10;
  };
  spedge(c.SPedge_e, "e");
  spedge(c.SPedge_p, "p");

  // ---- dbase_read, dbase.f:122-553
  bool doing_semipi = false, doing_semika = false;
  if (c.doing_pion && c.doing_semi) { doing_semipi = true; c.doing_pion = 0; }
  if (c.doing_kaon && c.doing_semi) { doing_semika = true; c.doing_kaon = 0; }
  const int nA = (int)std::lround(targ.A);
  if (c.doing_pion) {
    c.Mh = Mpi;
    if (c.doing_pizero) c.Mh = 134.9766;          // dbase.f:140
    if (nA == 1 && c.which_pion == 1 && !c.doing_pizero) throw std::runtime_error("Pi- production from Hydrogen not allowed!");
    if (nA <= 2 && c.which_pion >= 10) throw std::runtime_error("Coherent production from Hydrogen/Deuterium not allowed!");
    if (nA == 3 && c.which_pion == 11) throw std::runtime_error("Coherent Pi- production from 3He not allowed!");
    if (nA == 4 && c.which_pion >= 10) throw std::runtime_error("Coherent production from 4He not allowed!");
    c.doing_hydpi = nA == 1; c.doing_deutpi = nA == 2; c.doing_hepi = nA >= 3;
    if (c.which_pion >= 10) { c.doing_hydpi = 1; c.doing_deutpi = 0; c.doing_hepi = 0; }
  } else if (c.doing_kaon) {
    c.Mh = Mk;
    if (nA == 1 && c.which_kaon == 2) throw std::runtime_error("Sigma- production from Hydrogen not allowed!");
    if (nA <= 2 && c.which_kaon >= 10) throw std::runtime_error("Coherent production from Hydrogen/Deuterium not allowed!");
    c.doing_hydkaon = nA == 1; c.doing_deutkaon = nA == 2; c.doing_hekaon = nA >= 3;
    if (c.which_kaon >= 10) { c.doing_hydkaon = 1; c.doing_deutkaon = 0; c.doing_hekaon = 0; }
  } else if (c.doing_delta) {                    // dbase.f:182-190: the reference warns for A >= 2 ("only set up for proton target")
    c.Mh = Mp;
    if (nA != 1) throw std::runtime_error("Delta production (doing_delta) is only set up for a proton target");
  } else if (c.doing_semi) {
    c.Mh = doing_semika ? Mk : Mpi;
    c.doing_hydsemi = nA == 1; c.doing_deutsemi = nA == 2;
    c.doing_semipi = doing_semipi; c.doing_semika = doing_semika;
    c.do_fermi = D.i("do_fermi") > 0;
    if (c.doing_hydsemi && c.do_fermi) c.do_fermi = 0;      // 'Cannot do Fermi motion for Hydrogen!' (dbase.f:202-206)
    if (nA >= 3) throw std::runtime_error("semi-inclusive production from A >= 3 (doing_hesemi) is not implemented in this build");
  } else if (c.doing_rho) {                      // dbase.f:209-214, 861-873
    c.Mh = 769.3;
    if (nA >= 3) throw std::runtime_error("A(e,e'rho): not yet implemented (dbase.f:864)");
    if (nA != 1) throw std::runtime_error("rho production (doing_rho) is only set up for a proton target: the reference "
                                          "reads no momentum distribution for D(e,e'rho) (dbase.f:563)");
  } else {
    c.Mh = Mp;
    c.doing_eep = 1;
    c.doing_hyd_elast = nA == 1; c.doing_deuterium = nA == 2; c.doing_heavy = nA >= 3;
  }
  c.Mh2 = c.Mh * c.Mh;
  if (c.doing_phsp) { c.rad_flag = 0; c.doing_eep = 0; c.doing_pion = 0; c.doing_kaon = 0; c.doing_delta = 0; c.doing_rho = 0; }
  c.dEbeam = c.Ebeam * c.dEbeam / 100.;
  for (simc_spectrometer* sp : {&c.spec_e, &c.spec_p}) {
    sp->theta = std::fabs(sp->theta) / degrad;
    sp->cos_th = std::cos(sp->theta);
    sp->sin_th = std::sin(sp->theta);
  }
  auto arm_phi = [&](int arm, const char* what) {
    if (arm == 1 || arm == 3 || arm == 7) return 3 * pi / 2.;
    if (arm == 2 || arm == 4 || arm == 5 || arm == 6 || arm == 8) return pi / 2.;
    throw std::runtime_error(std::string("I dont know what phi should be for the ") + what + " arm");
  };
  c.spec_e.phi = arm_phi(c.electron_arm, "electron");
  c.spec_p.phi = arm_phi(c.hadron_arm, "hadron");
  targ.N = targ.A - targ.Z;
  targ.M = targ.mass_amu * amu;
  targ.Mrec = targ.mrec_amu * amu;
  if (nA == 1) { targ.M = Mp; targ.Mrec = 0.; }
  else if (nA == 2) targ.M = Md;
  else {
    const double Mrec_guess = targ.M - Mp;
    if (std::fabs(targ.Mrec - Mrec_guess) > 100.) targ.Mrec = Mrec_guess;
  }
  // sign_hadron, dbase.f:296-423: the charge of the detected hadron (target-field tracking only)
  c.sign_hadron = 1.0;
  if (c.doing_semi || c.doing_rho) c.sign_hadron = c.doing_hplus ? 1.0 : -1.0;
  else if (c.doing_pion) c.sign_hadron = (c.which_pion == 1 || c.which_pion == 3 || c.which_pion == 11) ? -1.0 : 1.0;
  if (c.doing_eep) { targ.Mtar_struck = Mp; targ.Mrec_struck = 0.0; }
  else if (c.doing_delta) { targ.Mtar_struck = Mp; targ.Mrec_struck = Mpi; }
  else if (c.doing_semi || c.doing_rho) { targ.Mtar_struck = Mp; targ.Mrec_struck = Mp; }
  else if (c.doing_pion && c.doing_pizero) {       // dbase.f:330-347: the nucleon keeps its charge
    if (c.which_pion == 0) { targ.Mtar_struck = Mp; targ.Mrec_struck = Mp; }
    else if (c.which_pion == 1) { targ.Mtar_struck = Mn; targ.Mrec_struck = Mn; }
    else if (c.which_pion == 2) { targ.Mtar_struck = Mp; targ.Mrec_struck = 1232.0; }
    else if (c.which_pion == 3) { targ.Mtar_struck = Mn; targ.Mrec_struck = 1232.0; }
    else throw std::runtime_error("Bad value for which_pion");
  } else if (c.doing_pion) {
    if (c.which_pion == 0) { targ.Mtar_struck = Mp; targ.Mrec_struck = Mn; }
    else if (c.which_pion == 1) { targ.Mtar_struck = Mn; targ.Mrec_struck = Mp; }
    else if (c.which_pion == 2 || c.which_pion == 3) { targ.Mtar_struck = Mp; targ.Mrec_struck = 1232.0; }
    else if (c.which_pion == 10 || c.which_pion == 11) {        // dbase.f:365-390: A(e,e'pi)A', production from a heavy "proton"
      targ.Mtar_struck = targ.M; targ.Mrec_struck = targ.Mrec;
      const double Mrec_guess = c.which_pion == 10 ? targ.M - Mp + Mn : targ.M - Mn + Mp;
      if (std::fabs(targ.Mrec_struck - Mrec_guess) > 100.) targ.Mrec_struck = Mrec_guess;
      targ.Mrec = 0.;
    } else throw std::runtime_error("Bad value for which_pion");
  } else if (c.doing_kaon) {
    if (c.which_kaon == 0) { targ.Mtar_struck = Mp; targ.Mrec_struck = 1115.68; }
    else if (c.which_kaon == 1) { targ.Mtar_struck = Mp; targ.Mrec_struck = 1192.64; }
    else if (c.which_kaon == 2) { targ.Mtar_struck = Mn; targ.Mrec_struck = 1197.45; }
    else if (c.which_kaon >= 10 && c.which_kaon <= 12) {        // dbase.f:409-438: bound hypernucleus in the final state
      targ.Mtar_struck = targ.M; targ.Mrec_struck = targ.Mrec;
      const double Mrec_guess = c.which_kaon == 10 ? targ.M - Mp + 1115.68 : c.which_kaon == 11 ? targ.M - Mp + 1192.64
                                                                                                 : targ.M - Mn + 1197.45;
      if (std::fabs(targ.Mrec_struck - Mrec_guess) > 100.) targ.Mrec_struck = Mrec_guess;
      targ.Mrec = 0.;
    } else throw std::runtime_error("Bad value for which_kaon");
  }
  if (nA == 2) targ.Mrec = Mp + Mn - targ.Mtar_struck;
  targ.thick = targ.thick / 1000.;
  targ.length = targ.thick / targ.rho;
  targ.angle = targ.angle / degrad;
  if (targ.Z < 2.4) {
    if (std::fabs(targ.angle) > 0.0001) targ.angle = 0.0;
    if (targ.can != 1 && targ.can != 2 && targ.can != 3) throw std::runtime_error("bad targ.can value");
  }
  if (std::sin(targ.angle) > 0.85) throw std::runtime_error("BAD targ.angle");
  for (simc_spectrometer* sp : {&c.spec_e, &c.spec_p}) { sp->off_xptar /= 1000.; sp->off_yptar /= 1000.; }
  for (simc_arm_cuts* a : {&c.SPedge_e, &c.SPedge_p}) {
    if (a->delta.min <= -100.0) a->delta.min = -99.99;
    a->yptar.min /= 1000.; a->yptar.max /= 1000.; a->xptar.min /= 1000.; a->xptar.max /= 1000.;
  }
  c.doing_tail[0] = one_tail == 0 || one_tail == 1 || one_tail == -2 || one_tail == -3;
  c.doing_tail[1] = one_tail == 0 || one_tail == 2 || one_tail == -3 || one_tail == -1;
  c.doing_tail[2] = one_tail == 0 || one_tail == 3 || one_tail == -1 || one_tail == -2;
  if (std::abs(one_tail) > 3 && c.using_rad)
    throw std::runtime_error("Moron! one_tail>3 turns radiation off, but using_rad wants it on.");
  if (!c.using_rad) c.doing_tail[0] = c.doing_tail[1] = c.doing_tail[2] = 0;
  c.hardwired_rad = c.Egamma_gen_max > 0.01;
  c.using_E_arm_montecarlo = spect_mode != 1 && spect_mode != -1;
  c.using_P_arm_montecarlo = spect_mode != 1 && spect_mode != -2;
  if (c.doing_pion || c.doing_kaon || c.doing_delta || (c.cuts_Em.min == c.cuts_Em.max)) {
    c.cuts_Em.min = -1.e6; c.cuts_Em.max = 1.e6;
  }
  if (std::abs(deForest_flag) > 1) throw std::runtime_error("Idiot! check setting of deForest_flag");
  c.deForest_flag = deForest_flag;
  if (c.correct_Eloss && !c.using_Eloss) c.correct_Eloss = 0;
  if ((int)std::lround(targ.Z) == 1) c.using_Coulomb = 0;

  // ---- target_init, init.f:1-87
  TargetExt X;
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
  }
  targ.X0_cm = targ.X0 / targ.rho;
  trip(c, 1, 0.0, c.Ebeam, 0.0, Me, 4, X.Eloss_ave[0], X.teff_ave[0]);
  trip(c, 2, 0.0, c.spec_e.P, c.spec_e.theta, Me, 4, X.Eloss_ave[1], X.teff_ave[1]);
  trip(c, 3, 0.0, std::sqrt(c.spec_p.P * c.spec_p.P + c.Mh2), c.spec_p.theta, std::sqrt(c.Mh2), 4, X.Eloss_ave[2],
       X.teff_ave[2]);
  if (!c.using_Eloss) X.Eloss_ave[0] = X.Eloss_ave[1] = X.Eloss_ave[2] = 0.0;
  if (c.using_Coulomb) {
    targ.Coulomb_ave = 0.75 * 1.5 * (targ.Z - 1.) * alpha * hbarc /
                       (1.1 * std::pow(targ.A, 1. / 3.) + 0.86 * std::pow(targ.A, -1. / 3.));
    targ.Coulomb_constant = targ.Coulomb_ave;
    targ.Coulomb_min = targ.Coulomb_constant;
    targ.Coulomb_max = targ.Coulomb_constant;
  }
  // theory_init (init.f:828-905): the deck setup needs the Pm range of the distributions and E_Fermi
  double theory_pm_max = 0.0, E_Fermi = 0.0;
  if (c.doing_deuterium || (c.doing_heavy && !c.use_benhar_sf)) {
    if (data_dir.empty())
      throw std::runtime_error(std::string("this deck needs ") + theory_file_for(nA) +
                               ": use simc_b200_config_from_deck_data with the directory that holds it");
    const TheoryFile T = read_theory_file(data_dir + "/" + theory_file_for(nA));
    E_Fermi = T.e_fermi;
    const bool deut = T.n_shells == 1 && T.e_fermi < 1.0;              // init.f:889
    if (deut != (c.doing_deuterium != 0))
      throw std::runtime_error("theory file and target disagree on doing_deuterium (init.f:889)");
    for (int i = 0; i < (c.doing_deuterium ? 1 : T.n_shells); ++i) {
      const double mn = T.pm_first[i] - T.pm_bin[i] / 2., mx = T.pm_last[i] + T.pm_bin[i] / 2.;
      theory_pm_max = std::max(theory_pm_max, std::max(std::fabs(mn), std::fabs(mx)));
    }
  }

  // momentum distribution of dbase.f:563-587 and, for A > 2, the spectral function of dbase.f:594-621: the deck
  // setup needs pval(nump) and Emval(numEm) (init.f:348-357)
  double pfermi_max = 0.0, sf_em_max = 0.0;
  if (c.doing_deutpi || c.doing_deutkaon || c.doing_hepi || c.doing_hekaon) {
    const bool he = c.doing_hepi || c.doing_hekaon;
    const char* pfile = !he ? "deut.dat" : nA == 3 ? "he3.dat" : nA == 4 ? "he4.dat" : "c12.dat";
    if (data_dir.empty())
      throw std::runtime_error(std::string("this deck needs ") + pfile + ": use simc_b200_config_from_deck_data with the directory that holds it");
    std::vector<double> pval, mprob;
    read_pfermi_file(data_dir + "/" + pfile, pval, mprob);
    pfermi_max = pval.back();
    if (he) {
      const char* sfile = nA == 3 ? "benharsf_3mod.dat" : nA == 4 ? "benharsf_4.dat" : nA == 56 ? "benharsf_56.dat"
                          : nA == 197 ? "benharsf_197.dat" : "benharsf_12.dat";
      const std::string path = data_dir + "/" + sfile;
      FILE* f = std::fopen(path.c_str(), "r");
      if (!f) throw std::runtime_error("cannot open spectral function file " + path);
      int n_pm = 0, n_em = 0;
      bool ok = std::fscanf(f, "%d %d", &n_pm, &n_em) == 2 && n_em >= 2 && n_em <= 200;
      for (int j = 0; ok && j < n_em; ++j) {            // first Pm row: Em of every bin
        double v[6];
        ok = std::fscanf(f, "%lf %lf %lf %lf %lf %lf", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5]) == 6;
        sf_em_max = v[1];
      }
      std::fclose(f);
      if (!ok) throw std::runtime_error("spectral function file " + path + ": bad header or short read");
    }
  }

  // ---- limits_init, init.f:91-572
  auto slop_for = [](int arm, double* used) {
    if (arm == 2) { used[0] = 1.0; used[1] = 0.008; used[2] = 0.008; }
    else if (arm >= 1 && arm <= 6) { used[0] = 0.5; used[1] = 0.005; used[2] = 0.005; }          // simulate.inc:20-44
    // calorimeter arms (7, 8): init.f:157-204 has no branch for them, the slop stays zero
  };
  if (c.using_E_arm_montecarlo) slop_for(c.electron_arm, c.slop_MC_e_used);
  if (c.using_P_arm_montecarlo) slop_for(c.hadron_arm, c.slop_MC_p_used);
  auto widen = [](simc_arm_cuts& a, const double* used) {
    a.delta.min -= used[0]; a.delta.max += used[0];
    a.yptar.min -= used[1]; a.yptar.max += used[1];
    a.xptar.min -= used[2]; a.xptar.max += used[2];
  };
  widen(c.SPedge_e, c.slop_MC_e_used);
  widen(c.SPedge_p, c.slop_MC_p_used);
  simc_edge& edge = c.edge;
  simc_edge& VE = c.VERTEXedge;
  for (simc_edge* e : {&edge, &VE}) {          // min_max_init
    for (simc_cut* q : {&e->e.E, &e->e.yptar, &e->e.xptar, &e->p.E, &e->p.yptar, &e->p.xptar, &e->Em, &e->Pm, &e->Mrec,
                        &e->Trec, &e->Trec_struck}) { q->min = -1.0e10; q->max = 1.0e10; }
  }
  edge.e.E.min = (1. + c.SPedge_e.delta.min / 100.) * c.spec_e.P + targ.Coulomb_min - c.dE_edge_test;
  edge.e.E.max = (1. + c.SPedge_e.delta.max / 100.) * c.spec_e.P + targ.Coulomb_max + c.dE_edge_test;
  simc_cut pp;
  pp.min = (1. + c.SPedge_p.delta.min / 100.) * c.spec_p.P - c.dE_edge_test;
  pp.max = (1. + c.SPedge_p.delta.max / 100.) * c.spec_p.P + c.dE_edge_test;
  pp.min = std::max(0.001e0, pp.min);
  edge.p.E.min = std::sqrt(pp.min * pp.min + c.Mh2);
  edge.p.E.max = std::sqrt(pp.max * pp.max + c.Mh2);
  simc_cut the_phys, thp_phys, z;
  the_phys.max = std::acos((c.spec_e.cos_th - c.spec_e.sin_th * c.SPedge_e.yptar.max) /
                           std::sqrt(1. + c.SPedge_e.yptar.max * c.SPedge_e.yptar.max + c.SPedge_e.xptar.max * c.SPedge_e.xptar.max));
  the_phys.min = std::acos((c.spec_e.cos_th - c.spec_e.sin_th * c.SPedge_e.yptar.min) /
                           std::sqrt(1. + c.SPedge_e.yptar.min * c.SPedge_e.yptar.min));
  thp_phys.max = std::acos((c.spec_p.cos_th - c.spec_p.sin_th * c.SPedge_p.yptar.max) /
                           std::sqrt(1. + c.SPedge_p.yptar.max * c.SPedge_p.yptar.max + c.SPedge_p.xptar.max * c.SPedge_p.xptar.max));
  thp_phys.min = std::acos((c.spec_p.cos_th - c.spec_p.sin_th * c.SPedge_p.yptar.min) /
                           std::sqrt(1. + c.SPedge_p.yptar.min * c.SPedge_p.yptar.min));
  z.min = -0.5 * targ.length;
  z.max = 0.5 * targ.length;
  extreme_trip_thru_target(c, X, c.Ebeam, the_phys, thp_phys, edge.e.E, pp, z, c.Mh);
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
  if (c.doing_heavy) {          // 'reconstructed' Em cuts, init.f:296-302
    const double slop_Coulomb = targ.Coulomb_max - targ.Coulomb_ave;
    double slop_Ebeam = c.dEbeam / 2. + slop_Coulomb;
    double slop_Ee = c.slop_MC_e_used[0] / 100. * c.spec_e.P + slop_Coulomb;
    const double r = std::sqrt(edge.p.E.max * edge.p.E.max - c.Mh2);
    double slop_Ep = std::sqrt((r + c.slop_MC_p_used[0] / 100. * c.spec_p.P) * (r + c.slop_MC_p_used[0] / 100. * c.spec_p.P) +
                               c.Mh2) - edge.p.E.max;
    slop_Ebeam = slop_Ebeam + (X.Eloss_max[0] - X.Eloss_min[0]);
    slop_Ee = slop_Ee + (X.Eloss_max[1] - X.Eloss_min[1]);
    slop_Ep = slop_Ep + (X.Eloss_max[2] - X.Eloss_min[2]);
    const double slop_total_Em = slop_Ebeam + slop_Ee + slop_Ep + c.dE_edge_test;
    slop_total_Em_used = slop_total_Em;
    edge.Em.min = c.cuts_Em.min - slop_total_Em;
    edge.Em.max = c.cuts_Em.max + slop_total_Em;
    edge.Em.min = std::max(0.e0, edge.Em.min);
  }
  if (c.doing_hyd_elast || c.doing_hydpi || c.doing_hydkaon || c.doing_delta || c.doing_rho || c.doing_semi) {   // doing_delta, doing_rho: hydrogen only here
    VE.Em.min = 0.0; VE.Em.max = 0.0; VE.Pm.min = 0.0; VE.Pm.max = 0.0;
    VE.Mrec.min = 0.0; VE.Mrec.max = 0.0; VE.Trec.min = 0.0; VE.Trec.max = 0.0;
  } else if (c.doing_hepi || c.doing_hekaon) {        // init.f:353-357,379-386
    VE.Em.min = targ.Mtar_struck + targ.Mrec - targ.M; VE.Em.max = sf_em_max;
    VE.Pm.min = 0.0; VE.Pm.max = pfermi_max;
    VE.Mrec.min = targ.M - targ.Mtar_struck + VE.Em.min;
    VE.Mrec.max = targ.M - targ.Mtar_struck + VE.Em.max;
    VE.Trec.min = std::sqrt(VE.Mrec.max * VE.Mrec.max + VE.Pm.min * VE.Pm.min) - VE.Mrec.max;
    VE.Trec.max = std::sqrt(VE.Mrec.min * VE.Mrec.min + VE.Pm.max * VE.Pm.max) - VE.Mrec.min;
  } else if (c.doing_deutpi || c.doing_deutkaon) {    // init.f:348-352,379-386
    VE.Em.min = Mp + Mn - targ.M; VE.Em.max = Mp + Mn - targ.M;
    VE.Pm.min = 0.0; VE.Pm.max = pfermi_max;
    VE.Mrec.min = targ.M - targ.Mtar_struck + VE.Em.min;
    VE.Mrec.max = targ.M - targ.Mtar_struck + VE.Em.max;
    VE.Trec.min = std::sqrt(VE.Mrec.max * VE.Mrec.max + VE.Pm.min * VE.Pm.min) - VE.Mrec.max;
    VE.Trec.max = std::sqrt(VE.Mrec.min * VE.Mrec.min + VE.Pm.max * VE.Pm.max) - VE.Mrec.min;
  } else if (c.doing_deuterium) {                     // init.f:326-330,379-386
    VE.Em.min = Mp + Mn - targ.M; VE.Em.max = Mp + Mn - targ.M;
    VE.Pm.min = 0.0; VE.Pm.max = theory_pm_max;
    VE.Mrec.min = targ.M - targ.Mtar_struck + VE.Em.min;
    VE.Mrec.max = targ.M - targ.Mtar_struck + VE.Em.max;
    VE.Trec.min = std::sqrt(VE.Mrec.max * VE.Mrec.max + VE.Pm.min * VE.Pm.min) - VE.Mrec.max;
    VE.Trec.max = std::sqrt(VE.Mrec.min * VE.Mrec.min + VE.Pm.max * VE.Pm.max) - VE.Mrec.min;
  } else if (c.doing_heavy) {                         // init.f:331-343,379-386
    VE.Pm.min = 0.0; VE.Pm.max = c.use_benhar_sf ? 790.0 : theory_pm_max;
    VE.Em.min = E_Fermi;                              // only theory_init sets it (init.f:856): 0 with use_benhar_sf
    VE.Em.max = 1000.;
    VE.Mrec.min = targ.M - targ.Mtar_struck + VE.Em.min;
    VE.Mrec.max = targ.M - targ.Mtar_struck + VE.Em.max;
    VE.Trec.min = std::sqrt(VE.Mrec.max * VE.Mrec.max + VE.Pm.min * VE.Pm.min) - VE.Mrec.max;
    VE.Trec.max = std::sqrt(VE.Mrec.min * VE.Mrec.min + VE.Pm.max * VE.Pm.max) - VE.Mrec.min;
  } else {
    throw std::runtime_error("limits for this target/reaction are not implemented in this build");
  }
  if (c.doing_eep || c.doing_semi) {
    VE.Trec_struck.min = 0.; VE.Trec_struck.max = 0.;
  } else {
    VE.Trec_struck.min = 0.;
    VE.Trec_struck.max = Ebeam_max + targ.Mtar_struck - targ.Mrec_struck - edge.e.E.min - edge.p.E.min - VE.Em.min -
                         VE.Trec.min;
  }
  c.Egamma_tot_max = Ebeam_max + targ.Mtar_struck - targ.Mrec_struck - edge.e.E.min - edge.p.E.min - VE.Em.min -
                     VE.Trec.min - VE.Trec_struck.min;
  if (c.doing_heavy) {
    const double t2 = (edge.Em.max - VE.Em.min) + (VE.Trec.max - VE.Trec.min);
    c.Egamma_tot_max = std::min(c.Egamma_tot_max, t2);
  }
  if (c.hardwired_rad) c.Egamma_tot_max = c.Egamma_gen_max;
  if (!c.using_rad) c.Egamma_tot_max = 0.0;
  if (c.doing_tail[0]) c.Egamma1_max = c.Egamma_tot_max;
  if (c.doing_tail[1]) c.Egamma2_max = c.Egamma_tot_max;
  if (c.doing_tail[2]) c.Egamma3_max = c.Egamma_tot_max;
  if (c.doing_heavy) {           // init.f:419-422
    VE.Em.min = std::max(VE.Em.min, edge.Em.min - c.Egamma_tot_max);
    VE.Em.max = std::min(VE.Em.max, edge.Em.max);
  }
  simc_gen_limits& gen = c.gen;
  if (c.doing_hyd_elast) {
    gen.sumEgen.min = 0.0; gen.sumEgen.max = 0.0;
  } else if (c.doing_heavy) {
    gen.sumEgen.max = Ebeam_max + targ.Mtar_struck - VE.Trec.min - VE.Em.min;
    gen.sumEgen.min = Ebeam_min + targ.Mtar_struck - VE.Trec.max - VE.Em.max - c.Egamma1_max;
    gen.sumEgen.max = std::min(gen.sumEgen.max, edge.e.E.max + edge.p.E.max + c.Egamma_tot_max);
    gen.sumEgen.min = std::max(gen.sumEgen.min, edge.e.E.min + edge.p.E.min);
  } else if (c.doing_semi) {
    gen.sumEgen.max = Ebeam_max + targ.Mtar_struck - targ.Mrec_struck;
    gen.sumEgen.min = edge.e.E.min + edge.p.E.min;
  } else {
    gen.sumEgen.max = Ebeam_max + targ.Mtar_struck - targ.Mrec_struck - edge.p.E.min - VE.Em.min - VE.Trec.min -
                      VE.Trec_struck.min;
    gen.sumEgen.min = Ebeam_min + targ.Mtar_struck - targ.Mrec_struck - edge.p.E.max - VE.Em.max - VE.Trec.max -
                      VE.Trec_struck.max - c.Egamma_tot_max;
    gen.sumEgen.max = std::min(gen.sumEgen.max, edge.e.E.max + c.Egamma2_max);
    gen.sumEgen.min = std::max(gen.sumEgen.min, edge.e.E.min);
  }
  gen.sumEgen.min -= c.dE_edge_test;
  gen.sumEgen.max += c.dE_edge_test;
  gen.sumEgen.min = std::max(0.e0, gen.sumEgen.min);
  if (c.doing_hyd_elast) {
    gen.e.E.min = edge.e.E.min;
    gen.e.E.max = edge.e.E.max + c.Egamma2_max;
  } else if (c.doing_deuterium || c.doing_pion || c.doing_kaon || c.doing_rho || c.doing_delta) {
    gen.e.E.min = gen.sumEgen.min;
    gen.e.E.max = gen.sumEgen.max;
  } else {
    gen.e.E.min = gen.sumEgen.min - edge.p.E.max - c.Egamma3_max;
    gen.e.E.max = gen.sumEgen.max - edge.p.E.min;
  }
  gen.e.E.min = std::max(gen.e.E.min, edge.e.E.min);
  gen.e.E.max = std::min(gen.e.E.max, edge.e.E.max + c.Egamma2_max);
  gen.e.delta.min = (gen.e.E.min / c.spec_e.P - 1.) * 100.;
  gen.e.delta.max = (gen.e.E.max / c.spec_e.P - 1.) * 100.;
  gen.e.yptar = edge.e.yptar;
  gen.e.xptar = edge.e.xptar;
  if (c.doing_hyd_elast || c.doing_deuterium || c.doing_pion || c.doing_kaon || c.doing_rho || c.doing_delta) {
    gen.p.E.min = edge.p.E.min;
    gen.p.E.max = edge.p.E.max + c.Egamma3_max;
  } else {
    gen.p.E.min = gen.sumEgen.min - edge.e.E.max - c.Egamma2_max;
    gen.p.E.max = gen.sumEgen.max - edge.e.E.min;
  }
  gen.p.E.min = std::max(gen.p.E.min, edge.p.E.min);
  gen.p.E.max = std::min(gen.p.E.max, edge.p.E.max + c.Egamma3_max);
  gen.p.delta.min = (std::sqrt(gen.p.E.min * gen.p.E.min - c.Mh2) / c.spec_p.P - 1.) * 100.;
  gen.p.delta.max = (std::sqrt(gen.p.E.max * gen.p.E.max - c.Mh2) / c.spec_p.P - 1.) * 100.;
  gen.p.yptar = edge.p.yptar;
  gen.p.xptar = edge.p.xptar;
  gen.Trec.min = -1.0e10; gen.Trec.max = 1.0e10;
  // histogram axes, init.f:519-569 (the three sets share them)
  const double nb = (double)SIMC_NHIST;
  for (int s = 0; s < 3; ++s) {
    simc_axis* ax = c.hist_axis[s];
    set_axis(ax[SIMC_H_E_DELTA], gen.e.delta.min, (gen.e.delta.max - gen.e.delta.min) / nb);
    set_axis(ax[SIMC_H_E_YPTAR], gen.e.yptar.min, (gen.e.yptar.max - gen.e.yptar.min) / nb);
    set_axis(ax[SIMC_H_E_XPTAR], -gen.e.xptar.max, (gen.e.xptar.max - gen.e.xptar.min) / nb);
    set_axis(ax[SIMC_H_P_DELTA], gen.p.delta.min, (gen.p.delta.max - gen.p.delta.min) / nb);
    set_axis(ax[SIMC_H_P_YPTAR], gen.p.yptar.min, (gen.p.yptar.max - gen.p.yptar.min) / nb);
    set_axis(ax[SIMC_H_P_XPTAR], -gen.p.xptar.max, (gen.p.xptar.max - gen.p.xptar.min) / nb);
    set_axis(ax[SIMC_H_EM], VE.Em.min, (std::max(100.e0, VE.Em.max) - VE.Em.min) / nb);
    set_axis(ax[SIMC_H_PM], VE.Pm.min, (std::max(100.e0, VE.Pm.max) - VE.Pm.min) / nb);
  }
  // ---- radc_init, init.f:576-651
  if (c.extrad_flag == 0) {
    if (c.rad_flag == 0) c.extrad_flag = 3;
    else if (c.rad_flag >= 1 && c.rad_flag <= 3) c.extrad_flag = 1;
  } else if (c.extrad_flag < 0) {
    throw std::runtime_error("Imbecile! check your stupid setting of EXTRAD_FLAG");
  }
  c.etatzai = (12.0 + (targ.Z + 1.) / (targ.Z * targ.L1 + targ.L2)) / 9.0;
  // ---- scale of the weights: the central-kinematics cross section (cf. calculate_central, simc.f:1143)
  c.w_ref = 1.0;
  if (c.doing_semi) {
    // semi-inclusive cross sections are a few nb/GeV/sr^2 = 1e-9 ub/MeV/sr^2; the weight needs the parton
    // tables, which a deck does not carry, so the scale is nominal
    c.w_ref = 1.0e-9;
  } else if (c.doing_rho) {
    c.w_ref = 1.0e-7;          // peerho at JLab kinematics: ~1e-7 ub/MeV/sr^2 (4 pi generation, steep t' slope): nominal
  } else if (c.doing_deuterium) {
    c.w_ref = 1.0e-5;          // sigma_cc1 (~1e2 ub/sr) x rho(Pm ~ 100 MeV/c) (~1e-7 MeV^-3): nominal
  } else if (c.doing_hyd_elast) {
    const double Ein = c.Ebeam_vertex_ave, uez = std::cos(c.spec_e.theta);
    const double eE = Ein * c.Mh / (c.Mh + Ein * (1. - uez));
    const double w = sigep(Ein, eE, c.spec_e.theta, 2 * Ein * eE * (1. - uez));
    if (w > 0 && std::isfinite(w)) c.w_ref = w;
  } else if (c.doing_heavy) {
    // central kinematics x a typical spectral-function density: S ~ 1/3200 per bin of 5 MeV x 20 MeV/c at Pm ~ 200 MeV/c
    EventState s{};
    s.v_Ein = c.Ebeam_vertex_ave; s.v_eE = c.spec_e.P; s.v_pP = c.spec_p.P; s.v_pE = std::sqrt(c.spec_p.P * c.spec_p.P + c.Mh2);
    s.v_etheta = c.spec_e.theta; s.v_ephi = c.spec_e.phi; s.v_ptheta = c.spec_p.theta; s.v_pphi = c.spec_p.phi;
    s.tz = targ.zoffset;
    simc_run_config quiet = c;
    quiet.using_Eloss = 0;
    struct NoRng { double uniform() { return 0.5; } } rng;
    auto nogauss = [](NoRng&, double) { return 0.0; };
    const MatTable mt = make_mat_table(c.targ);
    if (complete_ev_heavy(quiet, mt, rng, nogauss, s, true)) {
      HeavyEv ev;
      ev.Q2 = s.v_Q2; ev.q = s.v_q; ev.nu = s.v_nu; ev.Pm = s.v_Pm; ev.pE = s.v_pE; ev.pP = s.v_pP; ev.eE = s.v_eE;
      ev.etheta = s.v_etheta; ev.uqx = s.uqx; ev.uqy = s.uqy; ev.uqz = s.uqz; ev.upx = s.upx; ev.upy = s.upy; ev.upz = s.upz;
      ev.Pmx = ev.pP * ev.upx - ev.q * ev.uqx; ev.Pmy = ev.pP * ev.upy - ev.q * ev.uqy; ev.Pmz = ev.pP * ev.upz - ev.q * ev.uqz;
      const double w = deForest(ev, c.Mh2, c.deForest_flag) * targ.Z * c.transparency / 3200. / (4. * 3.14159265 * 200. * 200. * 100.);
      if (w > 0 && std::isfinite(w)) c.w_ref = w;
    }
  } else if (c.doing_pion || c.doing_kaon || c.doing_delta) {
    // central event: both particles along their spectrometer axes, electron at the central momentum, nucleon at rest
    EventState s{};
    s.efer = targ.Mtar_struck;
    if (!(c.doing_hydpi || c.doing_hydkaon || c.doing_delta)) {
      s.v_Em = (c.doing_hepi || c.doing_hekaon) ? targ.Mtar_struck + targ.Mrec - targ.M : Mp + Mn - targ.M;
      s.efer = targ.M - (targ.M - targ.Mtar_struck + s.v_Em);
    }
    s.v_Ein = c.Ebeam_vertex_ave; s.v_eE = c.spec_e.P;
    s.v_etheta = c.spec_e.theta; s.v_ephi = c.spec_e.phi; s.v_ptheta = c.spec_p.theta; s.v_pphi = c.spec_p.phi;
    s.tz = targ.zoffset;
    simc_run_config quiet = c;                 // kinematics only: no energy loss draws, no radiative constants
    quiet.using_Eloss = 0;
    struct NoRng { double uniform() { return 0.5; } } rng;
    auto nogauss = [](NoRng&, double) { return 0.0; };
    const MatTable mt = make_mat_table(c.targ);
    if (complete_ev_meson(quiet, mt, rng, nogauss, s, true)) {
      MesonVertex mv;
      mv.Ein = s.v_Ein; mv.eE = s.v_eE; mv.nu = s.v_nu; mv.q = s.v_q; mv.Q2 = s.v_Q2; mv.pP = s.v_pP; mv.pE = s.v_pE;
      mv.uqx = s.uqx; mv.uqy = s.uqy; mv.uqz = s.uqz; mv.upx = s.upx; mv.upy = s.upy; mv.upz = s.upz;
      mv.phi_pq = s.m_phipq; mv.t = s.m_t; mv.epsilon = s.m_eps;
      mv.pfer = 0.0; mv.pferx = 0.0; mv.pfery = 0.0; mv.pferz = 0.0; mv.efer = targ.Mtar_struck;
      const MesonWeight w = c.doing_pion ? peepi(c, MaidDev{nullptr}, mv) : c.doing_delta ? peedelta(c, mv) : peeK(c, mv);
      if (w.sigcc > 0 && std::isfinite(w.sigcc)) c.w_ref = w.sigcc;
    }
  }
  if (info) {         // what subroutine report (simc.f:644-1139) prints beyond the run constants
    std::memset(info, 0, sizeof(*info));
    info->ngen = D.i("ngen"); info->random_seed = D.i("random_seed"); info->one_tail = one_tail;
    info->doing_pizero = D.i("doing_pizero") > 0; info->pizero_ngamma = D.i("pizero_ngamma");
    info->use_first_cer = D.i("use_first_cer", 1) > 0; info->using_tgt_field = D.i("using_tgt_field") > 0;
    info->charge_mC = D.d("EXPER%charge");
    for (int i = 0; i < 3; ++i) {
      info->Eloss_ave[i] = X.Eloss_ave[i]; info->Eloss_min[i] = X.Eloss_min[i]; info->Eloss_max[i] = X.Eloss_max[i];
      info->teff_ave[i] = X.teff_ave[i]; info->teff_min[i] = X.teff_min[i]; info->teff_max[i] = X.teff_max[i];
      info->musc_max[i] = X.musc_max[i];
    }
    info->musc_nsig_max = 3.5;                           // target.f:569-577 (extreme_target_musc's nsig_max)
    info->slop_total_Em_used = slop_total_Em_used;
    const int nA_ = (int)std::lround(c.targ.A);
    if (c.doing_deuterium || c.doing_heavy) std::snprintf(info->theory_file, sizeof info->theory_file, "%s", theory_file_for(nA_));
  }
}

}  // namespace simc

extern "C" int simc_b200_config_from_deck(const char* deck_path, const char* extra_deck_dir, simc_run_config* out,
                                          int32_t* ngen, double* charge_mC, char* err, int errlen) {
  return simc_b200_config_from_deck_data(deck_path, extra_deck_dir, nullptr, out, ngen, charge_mC, err, errlen);
}

extern "C" int simc_b200_config_from_deck_data(const char* deck_path, const char* extra_deck_dir, const char* data_dir,
                                               simc_run_config* out, int32_t* ngen, double* charge_mC, char* err,
                                               int errlen) {
  if (!deck_path || !out) return SIMC_ERR_ARG;
  try {
    int ng = 0;
    double q = 0;
    simc::config_from_deck(deck_path, extra_deck_dir ? extra_deck_dir : "", data_dir ? data_dir : "", *out, &ng, &q);
    if (ngen) *ngen = ng;
    if (charge_mC) *charge_mC = q;
    return SIMC_OK;
  } catch (const std::exception& e) {
    if (err && errlen > 0) { std::strncpy(err, e.what(), errlen - 1); err[errlen - 1] = 0; }
    return SIMC_ERR_IO;
  }
}
