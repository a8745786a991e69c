// End-of-run text files of `program simc` in the reference's own layout, so that existing scripts that parse them
// keep working without the Fortran driver:
//   <base>.geni  STOP counters of the two spectrometers            simc.f:446-537   (list-directed writes)
//   <base>.gen   the 24 acceptance histograms, 50 bins each         simc.f:539-612   (format 3(1x,2(e11.4,1x)))
//   <base>.hist  subroutine report                                  simc.f:644-1139
// and the "central event" the report prints (calculate_central, simc.f:1143-1306).
// Host code; only simc_b200_central_event touches the GPU (central%sigcc comes from the same weight kernel as every
// event's, through simc_b200_weight_batch, so that it uses the run's tables).
// Fortran edit descriptors are reproduced by hand: Fw.d, Ew.d (0.dddE+ee form), Iw, Lw, Aw (right-justified), Tn, and
// gfortran's list-directed output for character and INTEGER*4 items (a leading blank; I12 per integer).
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/simc_b200.h"
#include "event.cuh"
#include "optics_host.h"

namespace simc {
void config_from_deck(const std::string& path, const std::string& extra_dir, const std::string& data_dir, simc_run_config& c,
                      int* ngen, double* charge_mC, simc_report_info* info);
}

namespace {

using std::string;

double fixed_value(const simc_fixed128& f) {
  const long double v = (long double)f.hi * 18446744073709551616.0L + (long double)f.lo;
  return (double)std::ldexp(v, f.qexp);
}

// ---- Fortran edit descriptors -------------------------------------------------------------------------------
string stars(int w) { return string((size_t)w, '*'); }
string fmt_f(double v, int w, int d) {                     // Fw.d
  char b[400];
  std::snprintf(b, sizeof b, "%.*f", d, v);
  string s = b;
  if ((int)s.size() > w) {
    if (s.compare(0, 2, "0.") == 0) s = s.substr(1);                // the optional leading zero goes first
    else if (s.compare(0, 3, "-0.") == 0) s = "-" + s.substr(2);
  }
  if ((int)s.size() > w) return stars(w);
  return string((size_t)(w - (int)s.size()), ' ') + s;
}
string fmt_e(double v, int w, int d) {                     // Ew.d: [-]0.dddE+ee
  if (std::isnan(v)) { string s = "NaN"; return (int)s.size() > w ? stars(w) : string((size_t)(w - 3), ' ') + s; }
  if (std::isinf(v)) { string s = v > 0 ? "Infinity" : "-Infinity"; if ((int)s.size() > w) s = v > 0 ? "Inf" : "-Inf";
                       return (int)s.size() > w ? stars(w) : string((size_t)(w - (int)s.size()), ' ') + s; }
  char b[400];
  std::snprintf(b, sizeof b, "%.*E", d - 1, std::fabs(v));          // D.DDDDE+XX with d significant digits
  string m = b;
  const size_t e = m.find('E');
  int ex = std::atoi(m.c_str() + e + 1);
  string digits;
  for (size_t i = 0; i < e; ++i) if (m[i] != '.') digits += m[i];
  if (v != 0.0) ex += 1;
  char eb[16];
  if (std::abs(ex) <= 99) std::snprintf(eb, sizeof eb, "E%c%02d", ex < 0 ? '-' : '+', std::abs(ex));
  else std::snprintf(eb, sizeof eb, "%c%03d", ex < 0 ? '-' : '+', std::abs(ex));
  string s = string(std::signbit(v) ? "-" : "") + "0." + digits + eb;
  if ((int)s.size() > w) s = string(std::signbit(v) ? "-" : "") + "." + digits + eb;
  if ((int)s.size() > w) return stars(w);
  return string((size_t)(w - (int)s.size()), ' ') + s;
}
string fmt_i(long long v, int w) {                         // Iw
  char b[64];
  std::snprintf(b, sizeof b, "%lld", v);
  string s = b;
  if ((int)s.size() > w) return stars(w);
  return string((size_t)(w - (int)s.size()), ' ') + s;
}
string fmt_l(bool v, int w) { return string((size_t)(w - 1), ' ') + (v ? "T" : "F"); }      // Lw
string fmt_a(const string& t, int w) {                     // Aw: right-justified, or the leftmost w characters
  if ((int)t.size() >= w) return t.substr(0, (size_t)w);
  return string((size_t)(w - (int)t.size()), ' ') + t;
}
string sp(int n) { return string((size_t)n, ' '); }
void tab_to(string& line, int col) {                       // Tn: next character goes to column n (1-based)
  if ((int)line.size() < col - 1) line += sp(col - 1 - (int)line.size());
  else line.resize((size_t)(col - 1));
}

struct Out {
  FILE* f;
  void line(const string& s) { std::fputs(s.c_str(), f); std::fputc('\n', f); }
  void list(const string& s) { line(" " + s); }           // write(iun,*) 'text'
};

// position of a named STOP counter (optics_host.cpp: stop_name) in simc_accum.stop[arm]
long long stop_of(const simc_accum& a, int which, int arm, const char* name) {
  for (int c = 1; 2 + c < SIMC_NSTOP; ++c)
    if (std::strcmp(simc::stop_name(arm, c), name) == 0) return (long long)a.stop[which][2 + c];
  return 0;
}

}  // namespace

extern "C" {

// Fw.d / Ew.d as the writers produce them (check entry point for the format helpers): kind = 'F' or 'E'
int simc_b200_format_real(int kind, int w, int d, double v, char* out, int outlen) {
  if (!out || outlen <= w || w <= 0 || d < 0) return SIMC_ERR_ARG;
  const string s = kind == 'E' ? fmt_e(v, w, d) : fmt_f(v, w, d);
  std::snprintf(out, (size_t)outlen, "%s", s.c_str());
  return SIMC_OK;
}

int simc_b200_report_info_from_deck(const char* deck_path, const char* extra_deck_dir, const char* data_dir,
                                    simc_report_info* out, char* err, int errlen) {
  if (!deck_path || !out) return SIMC_ERR_ARG;
  try {
    simc_run_config c;
    int ng = 0;
    double q = 0;
    std::memset(out, 0, sizeof(*out));
    simc::config_from_deck(deck_path, extra_deck_dir ? extra_deck_dir : "", data_dir ? data_dir : "", c, &ng, &q, out);
    return SIMC_OK;
  } catch (const std::exception& e) {
    if (err && errlen > 0) { std::strncpy(err, e.what(), errlen - 1); err[errlen - 1] = 0; }
    return SIMC_ERR_IO;
  }
}

// calculate_central, simc.f:1143-1306: the event with both particles on their spectrometer axes at the central
// momenta goes through complete_recon_ev (event.f:1056-1359), radc_init_ev and complete_main(force_sigcc = .true.).
// main0 is a zero-initialised local of the reference (-fno-automatic), so main%epsilon, phi_pq, t and the jacobian
// the cross-section routines read are zero: "may give non-physical kinematics", as the report itself says.
int simc_b200_central_event(simc_handle* h, const simc_run_config* cfg, const simc_report_info* info, simc_central* out) {
  if (!cfg || !info || !out) return SIMC_ERR_ARG;
  using namespace simc;
  const simc_run_config& c = *cfg;
  const simc_target& targ = c.targ;
  std::memset(out, 0, sizeof(*out));
  // ---- complete_recon_ev(vertex0)
  const double Ein = c.Ebeam_vertex_ave - targ.Coulomb_ave;
  const double eth = c.spec_e.theta, eph = c.spec_e.phi, pth = c.spec_p.theta, pph = c.spec_p.phi;
  const double eP = c.spec_e.P * (1. + 0.0 / 100.), eE = eP;
  double pP = c.spec_p.P * (1. + 0.0 / 100.);
  double pE = std::sqrt(pP * pP + c.Mh2);
  const double uex = std::sin(eth) * std::cos(eph), uey = std::sin(eth) * std::sin(eph), uez = std::cos(eth);
  const double upx = std::sin(pth) * std::cos(pph), upy = std::sin(pth) * std::sin(pph), upz = std::cos(pth);
  const double nu = Ein - eE;
  const double Q2 = 2 * Ein * eE * (1 - uez);
  const double q = std::sqrt(Q2 + nu * nu);
  const double uqx = -eP * uex / q, uqy = -eP * uey / q, uqz = (Ein - eP * uez) / q;
  const double W2 = SIMC_MP * SIMC_MP + 2. * SIMC_MP * nu - Q2;
  const double W = std::sqrt(std::fabs(W2)) * W2 / std::fabs(W2);
  if (c.doing_phsp) { pP = c.spec_p.P; pE = std::sqrt(c.Mh2 + pP * pP); }
  const double theta_pq = std::acos(std::fmin(1.0, upx * uqx + upy * uqy + upz * uqz));
  const double Pmx = pP * upx - q * uqx, Pmy = pP * upy - q * uqy, Pmz = pP * upz - q * uqz;
  const double Pm = std::sqrt(Pmx * Pmx + Pmy * Pmy + Pmz * Pmz);
  double Em = 0.0, zhad = 0.0, pt2 = 0.0;
  if (c.doing_pion || c.doing_kaon || c.doing_delta || c.doing_rho || c.doing_semi) Em = nu + targ.Mtar_struck - pE;
  if (c.doing_semi || c.doing_rho) { zhad = pE / nu; const double ct = std::cos(theta_pq); pt2 = pP * pP * (1.0 - ct * ct); }
  if (c.doing_hyd_elast) Em = nu + targ.M - pE - 0.0;
  else if (c.doing_deuterium || c.doing_heavy) Em = nu + targ.Mtar_struck - pE - (std::sqrt(Pm * Pm + targ.Mrec * targ.Mrec) - targ.Mrec);
  out->e_delta = 0.0; out->e_xptar = 0.0; out->e_yptar = 0.0; out->p_delta = 0.0; out->p_xptar = 0.0; out->p_yptar = 0.0;
  out->Q2 = Q2; out->q = q; out->nu = nu; out->Em = Em; out->Pm = Pm; out->W = W;
  out->MM = (Em * Em - Pm * Pm < 0) ? -std::sqrt(std::fabs(Em * Em - Pm * Pm)) : std::sqrt(Em * Em - Pm * Pm);
  // ---- radc_init_ev(main0, vertex0) with the average thicknesses (simc.f:1254-1274)
  if (c.using_rad) {
    VertexKin v;
    v.Ein = Ein; v.eE = eE; v.eP = eP; v.etheta = eth; v.pE = pE; v.pP = pP;
    v.uex = uex; v.uey = uey; v.uez = uez; v.upx = upx; v.upy = upy; v.upz = upz;
    RadEvDev R;
    radc_init_ev(c, v, info->teff_ave[0], info->teff_ave[1], R);
    const BasisConst B = basis_constants(R, Ein, eE, pE);
    out->hardcorfac = R.hardcorfac; out->etatzai = c.etatzai; out->g_int = B.g_int; out->g_ext = R.g_ext;
    for (int i = 0; i < 3; ++i) { out->frac[i] = R.frac[i]; out->lambda[i] = R.lambda[i]; }
    out->bt[0] = R.bt[0]; out->bt[1] = R.bt[1];
    for (int i = 0; i < 4; ++i) { out->c_int[i] = B.c_int[i]; out->c_ext[i] = B.c_ext[i]; out->c[i] = B.c[i]; out->g[i] = R.g[i]; }
  }
  // ---- complete_main(.true., main0, vertex0, vertex0, recon0): main0%sigcc through the weight kernel
  if (h) {
    double in[SIMC_WEIGHT_NIN], res[SIMC_WEIGHT_NOUT];
    for (double& x : in) x = 0.0;
    in[0] = eE; in[1] = eth; in[2] = eph; in[3] = pP; in[4] = pE; in[5] = pth; in[6] = pph;
    in[7] = Ein; in[8] = eE; in[9] = eth; in[10] = Q2; in[11] = nu; in[12] = q; in[13] = pE; in[14] = pP;
    in[15] = uqx; in[16] = uqy; in[17] = uqz; in[18] = upx; in[19] = upy; in[20] = upz; in[21] = Em; in[22] = Pm;
    in[23] = 0.0; in[24] = 0.0; in[25] = 0.0;            // main0%phi_pq, t, epsilon: never assigned
    in[26] = 0.0; in[27] = 1.0;                           // main0%jacobian (never assigned), gen_weight
    in[28] = zhad; in[29] = pt2;
    in[30] = 0.0; in[31] = 0.0; in[32] = 0.0; in[33] = 0.0;      // pfer, pferx, pfery, pferz (simc.f:1192-1195)
    in[34] = targ.M - std::sqrt((targ.M - targ.Mtar_struck + (targ.Mtar_struck + targ.Mrec - targ.M)) *
                                (targ.M - targ.Mtar_struck + (targ.Mtar_struck + targ.Mrec - targ.M)) + 0.0);     // efer
    const int rc = simc_b200_weight_batch(h, 1, in, res);
    if (rc) return rc;
    out->sigcc = res[3];
  }
  return SIMC_OK;
}

int simc_b200_write_geni(const char* path, const simc_run_config* cfg, const simc_accum* acc) {
  if (!path || !cfg || !acc) return SIMC_ERR_ARG;
  FILE* f = std::fopen(path, "w");
  if (!f) return SIMC_ERR_IO;
  Out o{f};
  const simc_accum& a = *acc;
  auto I = [](long long v) { return fmt_i(v, 12); };     // list-directed INTEGER*4
  for (int arm_id : {1, 2, 3, 4, 5}) {
    int w = -1;
    if (cfg->electron_arm == arm_id || (arm_id == 5 && cfg->electron_arm == 6)) w = 0;
    if (cfg->hadron_arm == arm_id || (arm_id == 5 && cfg->hadron_arm == 6)) w = w < 0 ? 1 : w;
    if (w < 0) continue;
    // the reference keeps ONE set of counters per spectrometer type; with the same type on both sides they add up
    auto S = [&](const char* nm) {
      long long v = 0;
      if (cfg->electron_arm == arm_id) v += stop_of(a, 0, arm_id, nm);
      if (cfg->hadron_arm == arm_id) v += stop_of(a, 1, arm_id, nm);
      return v;
    };
    auto T = [&](int slot) {
      long long v = 0;
      if (cfg->electron_arm == arm_id) v += a.stop[0][slot];
      if (cfg->hadron_arm == arm_id) v += a.stop[1][slot];
      return v;
    };
    if (arm_id == 1) {
      o.list("HMS Trials:           " + I(T(0)));
      o.list("Slit hor/vert/corners " + I(S("slit_hor")) + I(S("slit_vert")) + I(S("slit_oct")));
      o.list("Q1 entrance/mid/exit  " + I(S("Q1_in")) + I(S("Q1_mid")) + I(S("Q1_out")));
      o.list("Q2 entrance/mid/exit  " + I(S("Q2_in")) + I(S("Q2_mid")) + I(S("Q2_out")));
      o.list("Q3 entrance/mid/exit  " + I(S("Q3_in")) + I(S("Q3_mid")) + I(S("Q3_out")));
      o.list("Dipole entrance/exit  " + I(S("D1_in")) + I(S("D1_out")));
      o.list("Events reaching hut   " + I(T(2)));
      o.list("DC1, DC2, Scin, Cal   " + I(S("dc1")) + I(S("dc2")) + I(S("scin")) + I(S("cal")));
      o.list("Successes             " + I(T(1)));
      o.line("");
    } else if (arm_id == 2) {
      o.list("SOS Trials:           " + I(T(0)));
      o.list("Slit hor/vert/corners " + I(S("slit_vert")) + I(S("slit_hor")) + I(S("slit_oct")));      // order as written, simc.f:476
      o.list("Quad entrance/mid/exit" + I(S("quad_in")) + I(S("quad_mid")) + I(S("quad_out")));
      o.list("D1 entrance/exit      " + I(S("bm01_in")) + I(S("bm01_out")));
      o.list("D2 entrance/exit      " + I(S("bm02_in")) + I(S("bm02_out")));
      o.list("Vacuum exit           " + I(S("exit")));
      o.list("Events reaching hut   " + I(T(2)));
      o.list("DC1, DC2, Scin, Cal   " + I(S("dc1")) + I(S("dc2")) + I(S("scin")) + I(0));
      o.list("Successes             " + I(T(1)));
      o.line("");
    } else if (arm_id == 3 || arm_id == 4) {
      o.list(string(arm_id == 3 ? "HRSr Trials:          " : "HRSl Trials:          ") + I(T(0)));
      o.list("Slit hor/vert         " + I(S("slit_vert")) + I(S("slit_hor")));
      o.list("Q1 entrance/mid/exit  " + I(S("Q1_in")) + I(S("Q1_mid")) + I(S("Q1_out")));
      o.list("Q2 entrance/mid/exit  " + I(S("Q2_in")) + I(S("Q2_mid")) + I(S("Q2_out")));
      o.list("Dipole entrance/exit  " + I(S("D1_in")) + I(S("D1_out")));
      o.list("Q3 entrance/mid/exit  " + I(S("Q3_in")) + I(S("Q3_mid")) + I(S("Q3_out")));
      o.list("Events reaching hut   " + I(T(2)));
      o.list("VDC1, VDC2            " + I(S("dc1")) + I(S("dc2")));
      o.list("S1, S2, Cal\t    " + I(S("s1")) + I(S("s2")) + I(0));                 // the literal has a tab, simc.f:496
      if (arm_id == 3) o.line("");                                                   // (none after the HRSl block)
    } else {
      o.list("SHMS Trials:          " + I(T(0)));
      o.list("HB phys entrance/mag entr/mag exit/phys exit  " + I(S("HB_in")) + I(S("HB_men")) + I(S("HB_mex")) + I(S("HB_out")));
      o.list("Slit hor/vert/corners " + I(S("slit_hor")) + I(S("slit_vert")) + I(S("slit_oct")));
      o.list("Q1 phys entrance/mag entr/mid/mag exit/phys exit  " + I(S("Q1_in")) + I(S("Q1_men")) + I(S("Q1_mid")) + I(S("Q1_mex")) + I(S("Q1_out")));
      o.list("Q2 phys entrance/mag entr/mid/mag exit/phys exit  " + I(S("Q2_in")) + I(S("Q2_men")) + I(S("Q2_mid")) + I(S("Q2_mex")) + I(S("Q2_out")));
      // simc.f:520 prints shmsSTOP_q2_mid in the Q3 line, as written
      o.list("Q3 phys entrance/mag entr/mid/mag exit/phys exit       " + I(S("Q3_in")) + I(S("Q3_men")) + I(S("Q2_mid")) + I(S("Q3_mex")) + I(S("Q3_out")));
      o.list("D1 entrance/flare/mid 1-2   " + I(S("D1_in")) + I(S("D1_flr")) + I(S("D1_mid1")) + I(S("D1_mid2")));
      o.list("D1 mid 3-5            " + I(S("D1_mid3")) + I(S("D1_mid4")) + I(S("D1_mid5")));
      o.list("D1 mid 6-7/mag exit/phys exit         " + I(S("D1_mid6")) + I(S("D1_mid7")) + I(S("D1_mex")) + I(S("D1_out")));
      o.list("Events reaching hut   " + I(T(2)));
      o.list("DC1, DC2, Scin, Cal   " + I(S("dc1")) + I(S("dc2")));
      // mc_shms_hut.f:298,313 both count in s1; :359 (2x plane) in s3, :374 (2y plane) in s2; :396,439 in cal
      o.list("S1, S2, S3, Cal       " + I(S("s1x") + S("s1y")) + I(S("s2y")) + I(S("s2x")) + I(S("cal") + S("cal_fid")));
      o.list("Successes             " + I(T(1)));
      o.line("");
    }
  }
  return std::fclose(f) == 0 ? SIMC_OK : SIMC_ERR_IO;
}

int simc_b200_write_gen(const char* path, const simc_run_config* cfg, const simc_accum* acc) {
  if (!path || !cfg || !acc) return SIMC_ERR_ARG;
  FILE* f = std::fopen(path, "w");
  if (!f) return SIMC_ERR_IO;
  Out o{f};
  const simc_accum& a = *acc;
  auto pairs = [&](const std::vector<std::pair<double, double>>& v) {      // 3(1x,2(e11.4,1x))
    string s;
    for (const auto& p : v) s += " " + fmt_e(p.first, 11, 4) + " " + fmt_e(p.second, 11, 4) + " ";
    o.line(s);
  };
  auto centre = [&](int set, int k, int i) { const simc_axis& ax = cfg->hist_axis[set][k]; return ax.min + (i + 1 - 0.5) * ax.bin; };
  auto head = [&](const char* c1, const char* t1, const char* c2, const char* t2, const char* c3, const char* t3) {
    o.line(fmt_a(c1, 12) + fmt_a(t1, 12) + fmt_a(c2, 12) + fmt_a(t2, 12) + fmt_a(c3, 12) + fmt_a(t3, 12));
  };
  auto recon = [&](int k, int i) { return fixed_value(a.hist_w[k][i]); };
  auto cnt = [&](int set, int k, int i) { return (double)a.hist_n[set][k][i]; };
  const int E = SIMC_H_E_DELTA, P = SIMC_H_P_DELTA;
  o.list("E arm Experimental Target Distributions:");
  head("delta", "EXPERIM", "yptar", "EXPERIM", "xptar", "EXPERIM");
  for (int i = 0; i < SIMC_NHIST; ++i)
    pairs({{centre(0, E, i), recon(E, i)}, {centre(0, E + 1, i), recon(E + 1, i)}, {centre(0, E + 2, i), recon(E + 2, i)}});
  o.list("P arm Experimental Target Distributions:");
  head("delta", "EXPERIM", "yptar", "EXPERIM", "xptar", "EXPERIM");
  for (int i = 0; i < SIMC_NHIST; ++i)
    pairs({{centre(0, P, i), recon(P, i)}, {centre(0, P + 1, i), recon(P + 1, i)}, {centre(0, P + 2, i), recon(P + 2, i)}});
  o.list("Distributions of Contributing E arm Events:");
  head("delta", "CONTRIB", "yuptar", "CONTRIB", "xptar", "CONTRIB");
  for (int i = 0; i < SIMC_NHIST; ++i)
    pairs({{centre(1, E, i), cnt(1, E, i)}, {centre(1, E + 1, i), cnt(1, E + 1, i)}, {centre(1, E + 2, i), cnt(1, E + 2, i)}});
  o.list("Distributions of Contributing P arm Events:");
  head("delta", "CONTRIB", "yptar", "CONTRIB", "xptar", "CONTRIB");
  for (int i = 0; i < SIMC_NHIST; ++i)
    pairs({{centre(1, P, i), cnt(1, P, i)}, {centre(1, P + 1, i), cnt(1, P + 1, i)}, {centre(1, P + 2, i), cnt(1, P + 2, i)}});
  // simc.f:582-589: the "ORIGIN" E-arm block prints the gen axes and, for delta, the gen buffer (as written)
  o.list("Original E arm Events:");
  head("delta", "ORIGIN", "yptar", "ORIGIN", "xptar", "ORIGIN");
  for (int i = 0; i < SIMC_NHIST; ++i)
    pairs({{centre(1, E, i), cnt(1, E, i)}, {centre(1, E + 1, i), cnt(2, E + 1, i)}, {centre(1, E + 2, i), cnt(2, E + 2, i)}});
  o.list("Original P arm Events:");
  head("delta", "ORIGIN", "yptar", "ORIGIN", "xptar", "ORIGIN");
  for (int i = 0; i < SIMC_NHIST; ++i)
    pairs({{centre(2, P, i), cnt(2, P, i)}, {centre(2, P + 1, i), cnt(2, P + 1, i)}, {centre(2, P + 2, i), cnt(2, P + 2, i)}});
  o.list("Original Em/Pm distributions:");
  o.line(fmt_a("Em", 12) + fmt_a("ORIGIN", 12) + fmt_a("Pm", 12) + fmt_a("ORIGIN", 12));
  for (int i = 0; i < SIMC_NHIST; ++i) {                 // format (3(1x,4(e11.4,1x))) with four items
    string s = " ";
    for (double v : {centre(2, SIMC_H_EM, i), cnt(2, SIMC_H_EM, i), centre(2, SIMC_H_PM, i), cnt(2, SIMC_H_PM, i)}) s += fmt_e(v, 11, 4) + " ";
    o.line(s);
  }
  return std::fclose(f) == 0 ? SIMC_OK : SIMC_ERR_IO;
}

int simc_b200_write_hist(const char* path, const simc_run_config* cfg, const simc_report_info* info, const simc_central* central,
                         const simc_accum* acc, const simc_results* res, const char* timestring1, const char* timestring2) {
  if (!path || !cfg || !info || !central || !acc || !res) return SIMC_ERR_ARG;
  FILE* f = std::fopen(path, "w");
  if (!f) return SIMC_ERR_IO;
  Out o{f};
  const simc_run_config& c = *cfg;
  const simc_target& targ = c.targ;
  const simc_accum& a = *acc;
  const double degrad = 180. / 3.141592653589793, hbarc = 197.327053;
  auto A30 = [](const char* t) { string s = t ? t : ""; while (!s.empty() && (s.back() == '\n' || s.back() == '\r')) s.pop_back();
                                 s.resize(30, ' '); return s; };
  o.line("");                                                                   // '(/1x,...'
  o.line(" BEGIN Time: " + A30(timestring1));
  o.line(" END Time:   " + A30(timestring2));
  o.list("KINEMATICS:");
  const int nA = (int)std::lround(targ.A);
  if (c.doing_eep) {
    if (c.doing_hyd_elast) o.list("              ****--------  H(e,e'p)  --------****");
    else if (c.doing_deuterium) o.list("              ****--------  D(e,e'p)  --------****");
    else if (c.doing_heavy) o.list("              ****--------  A(e,e'p)  --------****");
  } else if (c.doing_semi) {
    const char* tgt = targ.A == 1 ? "H" : targ.A == 2 ? "D" : "A";
    const string had = c.doing_semipi ? (c.doing_hplus ? "pi+" : "pi-") : (c.doing_hplus ? "k+" : "k-");
    o.list(string(" ****--------  ") + tgt + "(e,e'" + had + ")X  --------****");
  } else if (c.doing_rho) {
    if (targ.A == 1) o.list("              ****--------  H(e,e'rho)  --------****");
    else o.list("I am not set up for anything else yet!");
  } else if (c.doing_delta) {
    // (the reference writes this banner to unit 6, simc.f:733-739: nothing goes to the file)
  } else if (c.doing_pion) {
    if (c.doing_hydpi) {
      if (targ.A == 1) o.list("              ****--------  H(e,e'pi)  --------****");
      else if (targ.A >= 3) o.list("              ****--------  A(e,e'pi)  --------****");
    } else if (c.doing_deutpi) o.list("              ****--------  D(e,e'pi)  --------****");
    else if (c.doing_hepi) o.list("              ****--------  A(e,e'pi)  --------****");
    if (c.which_pion == 0 || c.which_pion == 1) o.list("              ****----  Default Final State ----****");
    else if (c.which_pion == 10) o.list("              ****----  Final State is A + pi ----****");
    else if (c.which_pion == 2 || c.which_pion == 3) o.list("              ****----  Final State is pi + Delta ----****");
  } else if (c.doing_kaon) {
    if (c.doing_hydkaon) {
      if (targ.A == 1) o.list("              ****--------  H(e,e'K)  --------****");
      else if (targ.A >= 3) o.list("              ****--------  A(e,e'K)  --------****");
    } else if (c.doing_deutkaon) o.list("              ****--------  D(e,e'K)  --------****");
    else if (c.doing_hekaon) o.list("              ****--------  A(e,e'K)  --------****");
    static const char* const prod[3] = {"producing a LAMBDA", "producing a SIGMA0", "producing a SIGMA-"};
    static const char* const bound[3] = {"WITH BOUND LAMBDA ", "WITH BOUND SIGMA0 ", "WITH BOUND SIGMA- "};
    if (c.which_kaon >= 0 && c.which_kaon <= 2) o.list(string("              ****---- ") + prod[c.which_kaon] + " ----****");
    else if (c.which_kaon >= 10 && c.which_kaon <= 12) o.list(string("              ****---- ") + bound[c.which_kaon - 10] + " ----****");
  } else if (c.doing_phsp) {
    o.list("              ****--- PHASE SPACE - NO physics, NO radiation (may not work)---****");
  }
  (void)nA;
  auto kv = [&](const char* name, double v, const char* unit) { o.line(sp(9) + fmt_a(name, 12) + " = " + fmt_f(v, 15, 4) + sp(2) + fmt_a(unit, 10)); };
  kv("Ebeam", c.Ebeam, "MeV");
  kv("(dE/E)beam", c.dEbeam / c.Ebeam, "(full wid)");
  kv("x-width", c.gen.xwid, "cm");
  kv("y-width", c.gen.ywid, "cm");
  o.line(sp(9) + fmt_a("fr_pattern", 12) + " = " + fmt_i(targ.fr_pattern, 15) + sp(2) + fmt_a("1=square,2=circ", 16));
  kv("fr1", targ.fr1, "cm");
  kv("fr2", targ.fr2, "cm");
  o.list(" ");
  o.line(sp(9) + sp(18) + fmt_a("____E arm____", 15) + sp(2) + fmt_a("____P arm____", 15));      // (a trailing 2x writes nothing)
  auto kv2 = [&](const char* name, double v1, double v2, const char* unit) {
    o.line(sp(9) + fmt_a(name, 12) + " = " + fmt_f(v1, 15, 4) + sp(2) + fmt_f(v2, 15, 4) + sp(2) + sp(2) + fmt_a(unit, 5));
  };
  kv2("angle", c.spec_e.theta * degrad, c.spec_p.theta * degrad, "deg");
  kv2("momentum", c.spec_e.P, c.spec_p.P, "MeV/c");
  kv2("x offset", c.spec_e.off_x, c.spec_p.off_x, "cm");
  kv2("y offset", c.spec_e.off_y, c.spec_p.off_y, "cm");
  kv2("z offset", c.spec_e.off_z, c.spec_p.off_z, "cm");
  kv2("xptar offset", c.spec_e.off_xptar, c.spec_p.off_xptar, "mr");
  kv2("yptar offset", c.spec_e.off_yptar, c.spec_p.off_yptar, "mr");
  o.list("                      VALUES FOR \"CENTRAL\" EVENT:");
  kv2("delta", central->e_delta, central->p_delta, "%");
  kv2("xptar", central->e_xptar, central->p_xptar, "mr");
  kv2("yptar", central->e_yptar, central->p_yptar, "mr");
  auto kc = [&](const char* name, double v, const char* unit) { o.line(sp(17) + fmt_a(name, 10) + " = " + fmt_f(v, 15, 4) + sp(2) + fmt_a(unit, 9)); };
  kc("Q2", central->Q2 / 1.e6, "(GeV/c)^2");
  kc("q", central->q, "MeV/c");
  kc("nu", central->nu, "MeV");
  kc("recon Em", central->Em, "MeV");
  kc("recon Pm", central->Pm, "MeV/c");
  kc("recon  W", central->W, "MeV/c");
  kc("recon MM", central->MM, "MeV/c");
  // ---- target
  o.list("TARGET specs:");
  auto t2 = [&](const char* n1, double v1, const char* u1, const char* n2, double v2, const char* u2) {      // 9911
    o.line(sp(2) + sp(5) + fmt_a(n1, 10) + " = " + fmt_e(v1, 12, 6) + " " + fmt_a(u1, 5) + sp(5) + fmt_a(n2, 10) + " = " + fmt_e(v2, 12, 6) +
           " " + fmt_a(u2, 5));
  };
  t2("A", targ.A, " ", "Z", targ.Z, " ");
  t2("mass", targ.mass_amu, "amu", "mass", targ.M, "MeV");
  t2("Mrec", targ.mrec_amu, "amu", "Mrec", targ.Mrec, "MeV");
  t2("Mtar_struc", targ.Mtar_struck, "MeV", "Mrec_struc", targ.Mrec_struck, "MeV");
  t2("rho", targ.rho, "g/cm3", "thick", targ.thick, "g/cm2");
  t2("angle", targ.angle * degrad, "deg", "abundancy", targ.abundancy, "%");
  t2("X0", targ.X0, "g/cm2", "X0_cm", targ.X0_cm, "cm");
  t2("length", targ.length, "cm", "zoffset", targ.zoffset, "cm");
  t2("xoffset", targ.xoffset, "cm", "yoffset", targ.yoffset, "cm");
  { string s; tab_to(s, 12); s += fmt_a("__ave__", 15) + fmt_a("__lo__", 15) + fmt_a("__hi__", 15); o.line(s); }
  auto t3 = [&](const char* n, double a1, double a2, double a3, const char* u) {      // 9912
    o.line(" " + fmt_a(n, 15) + fmt_f(a1, 15, 5) + fmt_f(a2, 15, 5) + fmt_f(a3, 15, 5) + sp(2) + fmt_a(u, 6));
  };
  t3("Coulomb", targ.Coulomb_ave, targ.Coulomb_min, targ.Coulomb_max, "MeV");
  static const char* const en[3] = {"Eloss_beam", "Eloss_e", "Eloss_p"};
  static const char* const tn[3] = {"teff_beam", "teff_e", "teff_p"};
  for (int i = 0; i < 3; ++i) t3(en[i], info->Eloss_ave[i], info->Eloss_min[i], info->Eloss_max[i], "MeV");
  for (int i = 0; i < 3; ++i) t3(tn[i], info->teff_ave[i], info->teff_min[i], info->teff_max[i], "radlen");
  auto t4 = [&](const char* n, double v, const char* u) {      // 9913: 1x,a15,t25,f15.5,2x,a6
    string s = " " + fmt_a(n, 15); tab_to(s, 25); s += fmt_f(v, 15, 5) + sp(2) + fmt_a(u, 6); o.line(s);
  };
  t4("musc_nsig_max", info->musc_nsig_max, " ");
  t4("musc_max_beam", info->musc_max[0] * 1000., "mr");
  t4("musc_max_e", info->musc_max[1] * 1000., "mr");
  t4("musc_max_p", info->musc_max[2] * 1000., "mr");
  // ---- flags
  o.list("FLAGS:");
  auto L = [&](const char* n, int v) { return sp(2) + fmt_a(n, 19) + "=" + fmt_l(v != 0, 2); };
  auto I2 = [&](const char* n, int v) { return sp(2) + fmt_a(n, 19) + "=" + fmt_i(v, 2); };
  o.line(sp(5) + L("doing_eep", c.doing_eep) + L("doing_kaon", c.doing_kaon) + L("doing_pion", c.doing_pion));
  o.line(sp(5) + L("doing_semi", c.doing_semi) + L("doing_rho", c.doing_rho) + L("doing_hplus", c.doing_hplus));
  o.line(sp(5) + L("doing_semipi", c.doing_semipi) + L("doing_semika", c.doing_semika) + L("doing_pizero", info->doing_pizero));
  o.line(sp(5) + L("doing_delta", c.doing_delta) + L("doing_phsp", c.doing_phsp));
  o.line(sp(5) + I2("which_pion", c.which_pion) + I2("which_kaon", c.which_kaon) + I2("pizero_ngamma", info->pizero_ngamma));
  o.line(sp(5) + L("doing_hyd_elast", c.doing_hyd_elast) + L("doing_deuterium", c.doing_deuterium) + L("doing_heavy", c.doing_heavy));
  o.line(sp(5) + L("doing_hydpi", c.doing_hydpi) + L("doing_deutpi", c.doing_deutpi) + L("doing_hepi", c.doing_hepi));
  o.line(sp(5) + L("doing_hydkaon", c.doing_hydkaon) + L("doing_deutkaon", c.doing_deutkaon) + L("doing_hekaon", c.doing_hekaon));
  o.line(sp(5) + L("doing_hydsemi", c.doing_hydsemi) + L("doing_deutsemi", c.doing_deutsemi) + L("do_fermi", c.do_fermi));
  o.line(sp(5) + L("doing_hydrho", info->doing_hydrho) + L("doing_deutrho", info->doing_deutrho) + L("doing_herho", info->doing_herho));
  o.line(sp(5) + L("mc_smear", c.mc_smear) + I2("electron_arm", c.electron_arm) + I2("hadron_arm", c.hadron_arm));
  o.line(sp(5) + L("using_Eloss", c.using_Eloss) + L("using_Coulomb", c.using_Coulomb) + L("deForest_flag", c.deForest_flag));
  o.line(sp(5) + L("correct_Eloss", c.correct_Eloss) + L("correct_raster", c.correct_raster) + L("doing_decay", c.doing_decay));
  o.line(sp(5) + L("using_E_arm_montecarlo", c.using_E_arm_montecarlo) + L("using_P_arm_montecarlo", c.using_P_arm_montecarlo) +
         L("use_benhar_sf", c.use_benhar_sf));
  if (c.electron_arm == 5 || c.hadron_arm == 5 || c.electron_arm == 6 || c.hadron_arm == 6)
    o.line(sp(7) + fmt_a("use_first_cer", 19) + "=" + fmt_l(info->use_first_cer != 0, 2));
  o.line(sp(7) + fmt_a("ctau", 11) + "=" + fmt_f(c.ctau, 10, 3) + fmt_a("cm", 4));
  if (c.use_benhar_sf) o.line(sp(7) + fmt_a("transparency", 12) + "=" + fmt_f(c.transparency, 8, 4));
  // ---- counters
  o.list("COUNTERS:");
  o.line(sp(12) + "Ngen (request) = " + fmt_i(info->ngen, 10));
  o.line(sp(12) + "Ntried         = " + fmt_i(a.ntried, 10));
  o.line(sp(12) + "Ncontribute    = " + fmt_i(a.ncontribute, 10));
  o.line(sp(12) + "Nco_no_rad_prot= " + fmt_i(a.ncontribute_no_rad_proton, 10));
  o.line(sp(12) + "-> %no_rad_prot= " + fmt_f(100. * (double)a.ncontribute_no_rad_proton / std::fmax((double)a.ncontribute, 0.1e0), 10, 3));
  o.line("");
  o.line(" INTEGRATED WEIGHTS (number of counts in delta/Em cuts!):");
  // simc.f:399,921: wtcontribute has been multiplied by normfac; the report divides it by nevent
  o.line("               MeV: wtcontr= " + fmt_e(res->yield / (double)res->nevent, 16, 8));
  // ---- radiative corrections
  o.list("RADIATIVE CORRECTIONS:");
  auto L3 = [&](const char* n, int v) { return " " + fmt_a(n, 14) + "=" + fmt_l(v != 0, 3); };
  auto I3 = [&](const char* n, int v) { return " " + fmt_a(n, 14) + "=" + fmt_i(v, 3); };
  if (!c.using_rad) {
    o.line(L3("using_rad", 0));
  } else {
    o.line(L3("use_expon", c.use_expon) + L3("include_hard", 1) + L3("calc_spence", 1));
    o.line(L3("using_rad", c.using_rad) + L3("use_offshell_rad", c.use_offshell_rad));
    o.line(I3("rad_flag", c.rad_flag) + I3("extrad_flag", c.extrad_flag) + I3("one_tail", info->one_tail) + I3("intcor_mode", c.intcor_mode));
    o.line(" " + fmt_a("dE_edge_test", 14) + "=" + fmt_f(c.dE_edge_test, 11, 3));
    o.line(" " + fmt_a("Egamma_max", 14) + "=" + fmt_f(c.Egamma_tot_max, 11, 3));
    auto r4 = [&](const char* n, const double* v, int k) {      // 9914: 1x,a18,' = ',4f11.3
      string s = " " + fmt_a(n, 18) + " = ";
      for (int i = 0; i < k; ++i) s += fmt_f(v[i], 11, 3);
      o.line(s);
    };
    o.list("Central Values:");
    r4("hardcorfac", &central->hardcorfac, 1);
    r4("etatzai", &central->etatzai, 1);
    r4("frac(1:3)", central->frac, 3);
    r4("lambda(1:3)", central->lambda, 3);
    r4("bt(1:2)", central->bt, 2);
    r4("c_int(0:3)", central->c_int, 4);
    r4("c_ext(0:3)", central->c_ext, 4);
    r4("c(0:3)", central->c, 4);
    r4("g_int", &central->g_int, 1);
    r4("g_ext", &central->g_ext, 1);
    r4("g(0:3)", central->g, 4);
  }
  // ---- miscellaneous
  o.list("MISCELLANEOUS:");
  auto m1 = [&](const char* n, double v, const char* u) { o.line(sp(12) + fmt_a(n, 14) + " = " + fmt_e(v, 16, 6) + " " + fmt_a(u, 6)); };      // 9915
  o.list("Note that central.sigcc is for central delta,theta,phi in both spectrometers");
  o.list(" and may give non-physical kinematics, esp. for Hydrogen");
  o.list("Note also that AVE.sigcc is really AVER.weight (the two arenot exactly equal)");
  const double targetfac = targ.mass_amu / 3.75914e+6 / (targ.abundancy / 100.) * std::fabs(std::cos(targ.angle)) / (targ.thick * 1000.);
  m1("CENTRAL.sigcc", central->sigcc, " ");
  m1("AVERAGE.sigcc", res->central_sigcc_ave, " ");
  m1("charge", info->charge_mC, "mC");
  m1("targetfac", targetfac, " ");
  m1("luminosity", res->luminosity, "ub^-1");
  m1("luminosity", res->luminosity * (hbarc / 100000.) * (hbarc / 100000.), "GeV^2");
  m1("genvol", res->genvol, " ");
  m1("normfac", res->normfac, " ");
  if (c.doing_heavy) o.line(sp(12) + "Theory file:  " + string(info->theory_file));
  if (info->random_seed != 0) o.line(sp(15) + "Random Seed = " + fmt_i(info->random_seed, 10));
  // ---- resolution summary
  o.list("RECON SUMMARY:            Ave.Error   Resolution");
  auto rs = [&](const char* n, double a1, double a2, const char* u) { o.line(sp(2) + fmt_a(n, 22) + fmt_f(a1, 12, 5) + fmt_f(a2, 12, 5) + sp(2) + u); };
  rs("Electron arm: delta =", 10. * res->aveerr[0], 10. * res->resol[0], "x10^-3");
  rs("xptar =", 1000. * res->aveerr[1], 1000. * res->resol[1], "mr");
  rs("yptar =", 1000. * res->aveerr[2], 1000. * res->resol[2], "mr");
  rs("ytar  =", 10. * res->aveerr[3], 10. * res->resol[3], "mm");
  rs("Hadron arm:   delta =", 10. * res->aveerr[4], 10. * res->resol[4], "x10-3");
  rs("xptar =", 1000. * res->aveerr[5], 1000. * res->resol[5], "mr");
  rs("yptar =", 1000. * res->aveerr[6], 1000. * res->resol[6], "mr");
  rs("ytar  =", 10. * res->aveerr[7], 10. * res->resol[7], "mm");
  // ---- limits
  o.list("Input Spectrometer Limits:");
  auto sl = [&](const char* n, double v1, double v2, const char* u) { o.line(sp(9) + fmt_a(n, 25) + " = " + sp(2) + fmt_f(v1, 15, 4) + sp(2) + fmt_f(v2, 15, 4) + fmt_a(u, 5)); };
  sl("SPedge.e.delta.min/max", c.SPedge_e.delta.min, c.SPedge_e.delta.max, "%");
  sl("SPedge.e.yptar.min/max", c.SPedge_e.yptar.min, c.SPedge_e.yptar.max, "rad");
  sl("SPedge.e.xptar.min/max", c.SPedge_e.xptar.min, c.SPedge_e.xptar.max, "rad");
  sl("SPedge.p.delta.min/max", c.SPedge_p.delta.min, c.SPedge_p.delta.max, "%");
  sl("SPedge.p.yptar.min/max", c.SPedge_p.yptar.min, c.SPedge_p.yptar.max, "rad");
  sl("SPedge.p.xptar.min/max", c.SPedge_p.xptar.min, c.SPedge_p.xptar.max, "rad");
  auto used_found = [&]() {
    { string s; tab_to(s, 25); s += sp(2) + fmt_a("______used______", 16) + sp(2); tab_to(s, 50); s += sp(2) + fmt_a("_____found______", 16); o.line(s); }
    { string s; tab_to(s, 25); s += fmt_a("min", 10) + fmt_a("max", 10); tab_to(s, 50); s += fmt_a("lo", 10) + fmt_a("hi", 10); o.line(s); }
  };
  auto lim = [&](const char* n, double u1, double u2, double lo, double hi, const char* u) {      // 9917
    string s = " " + fmt_a(n, 18); tab_to(s, 21); s += fmt_f(u1, 12, 3) + fmt_f(u2, 12, 3); tab_to(s, 50);
    s += fmt_f(lo, 10, 3) + fmt_f(hi, 10, 3) + sp(2) + fmt_a(u, 5); o.line(s);
  };
  const simc_range* C = a.contrib;
  o.list("Limiting VERTEX values (vertex.e/p.*,Em,Pm,Trec)");
  o.list("   USED limits are gen.e/p.*, and VERTEXedge.Em,Pm,Trec");
  used_found();
  lim("E arm  delta", c.gen.e.delta.min, c.gen.e.delta.max, C[0].lo, C[0].hi, "%");
  lim("E arm  yptar", c.gen.e.yptar.min * 1000., c.gen.e.yptar.max * 1000., C[1].lo * 1000., C[1].hi * 1000., "mr");
  lim("E arm  xptar", c.gen.e.xptar.min * 1000., c.gen.e.xptar.max * 1000., C[2].lo * 1000., C[2].hi * 1000., "mr");
  lim("P arm  delta", c.gen.p.delta.min, c.gen.p.delta.max, C[3].lo, C[3].hi, "%");
  lim("P arm  yptar", c.gen.p.yptar.min * 1000., c.gen.p.yptar.max * 1000., C[4].lo * 1000., C[4].hi * 1000., "mr");
  lim("P arm  xptar", c.gen.p.xptar.min * 1000., c.gen.p.xptar.max * 1000., C[5].lo * 1000., C[5].hi * 1000., "mr");
  lim("sumEgen", c.gen.sumEgen.min, c.gen.sumEgen.max, C[7].lo, C[7].hi, "MeV");
  lim("Trec", c.VERTEXedge.Trec.min, c.VERTEXedge.Trec.max, C[23].lo, C[23].hi, "MeV");
  lim("Em", c.VERTEXedge.Em.min, c.VERTEXedge.Em.max, C[24].lo, C[24].hi, "MeV");
  lim("Pm", c.VERTEXedge.Pm.min, c.VERTEXedge.Pm.max, C[25].lo, C[25].hi, "MeV/c");
  if ((c.doing_deuterium || c.doing_pion || c.doing_kaon || c.doing_delta) && c.using_rad)
    o.list("      *** NOTE: sumEgen.min only used in GENERATE_RAD");
  o.list("Limiting ORIGINAL values: orig.e/p.*,Em,Pm,Trec (no edge.* limits for Pm,Trec)");
  used_found();
  lim("E arm   E", c.edge.e.E.min, c.edge.e.E.max, C[8].lo, C[8].hi, "MeV");
  lim("E arm  yptar", c.edge.e.yptar.min * 1000., c.edge.e.yptar.max * 1000., C[10].lo * 1000., C[10].hi * 1000., "mr");
  lim("E arm  xptar", c.edge.e.xptar.min * 1000., c.edge.e.xptar.max * 1000., C[9].lo * 1000., C[9].hi * 1000., "mr");
  lim("P arm      E", c.edge.p.E.min, c.edge.p.E.max, C[11].lo, C[11].hi, "MeV");
  lim("P arm  yptar", c.edge.p.yptar.min * 1000., c.edge.p.yptar.max * 1000., C[12].lo * 1000., C[12].hi * 1000., "mr");
  lim("P arm  xptar", c.edge.p.xptar.min * 1000., c.edge.p.xptar.max * 1000., C[13].lo * 1000., C[13].hi * 1000., "mr");
  lim("Em", std::fmax(-999999.999e0, c.edge.Em.min), std::fmin(999999.999e0, c.edge.Em.max), C[14].lo, C[14].hi, "MeV");
  lim("Pm", 0., 0., C[15].lo, C[15].hi, "MeV");
  lim("Trec", 0., 0., C[16].lo, C[16].hi, "MeV");
  o.list("Limiting SPECTROMETER values");
  used_found();
  lim("E arm delta", c.SPedge_e.delta.min, c.SPedge_e.delta.max, C[17].lo, C[17].hi, "%");
  lim("E arm  yptar", c.SPedge_e.yptar.min * 1000., c.SPedge_e.yptar.max * 1000., C[18].lo * 1000., C[18].hi * 1000., "mr");
  lim("E arm  xptar", c.SPedge_e.xptar.min * 1000., c.SPedge_e.xptar.max * 1000., C[19].lo * 1000., C[19].hi * 1000., "mr");
  lim("P arm  delta", c.SPedge_p.delta.min, c.SPedge_p.delta.max, C[20].lo, C[20].hi, "%");
  lim("P arm  yptar", c.SPedge_p.yptar.min * 1000., c.SPedge_p.yptar.max * 1000., C[21].lo * 1000., C[21].hi * 1000., "mr");
  lim("P arm  xptar", c.SPedge_p.xptar.min * 1000., c.SPedge_p.xptar.max * 1000., C[22].lo * 1000., C[22].hi * 1000., "mr");
  if (c.using_rad) {
    o.list("Limiting RADIATION values CONTRIBUTING to the (Em,Pm) distributions:");
    used_found();
    lim("Egamma(1)", 0., c.Egamma1_max, C[26].lo, C[26].hi, "MeV");
    lim("Egamma(2)", 0., c.Egamma2_max, C[27].lo, C[27].hi, "MeV");
    lim("Egamma(3)", 0., c.Egamma3_max, C[28].lo, C[28].hi, "MeV");
    lim("Egamma_total", 0., c.Egamma_tot_max, C[29].lo, C[29].hi, "MeV");
  }
  o.list("ACTUAL and LIMITING SLOP values used/obtained:");
  { string s; tab_to(s, 25); s += fmt_a("__used__", 10) + fmt_a("__min__", 10) + fmt_a("__max__", 10); o.line(s); }
  auto slp = [&](const char* n1, const char* n2, double used, double lo, double hi, const char* u) {      // 9918
    string s = " " + fmt_a(n1, 10) + fmt_a(n2, 12); tab_to(s, 25); s += fmt_f(used, 10, 3) + fmt_f(lo, 10, 3) + fmt_f(hi, 10, 3) + sp(2) + fmt_a(u, 5);
    o.line(s);
  };
  const simc_range* SL = a.slop;
  slp("slop.MC  ", "E arm delta", c.slop_MC_e_used[0], SL[0].lo, SL[0].hi, "%");
  slp(" ", "E arm yptar", c.slop_MC_e_used[1] * 1000., SL[1].lo * 1000., SL[1].hi * 1000., "mr");
  slp(" ", "E arm xptar", c.slop_MC_e_used[2] * 1000., SL[2].lo * 1000., SL[2].hi * 1000., "mr");
  slp(" ", "P arm delta", c.slop_MC_p_used[0], SL[3].lo, SL[3].hi, "%");
  slp(" ", "P arm yptar", c.slop_MC_p_used[1] * 1000., SL[4].lo * 1000., SL[4].hi * 1000., "mr");
  slp(" ", "P arm xptar", c.slop_MC_p_used[2] * 1000., SL[5].lo * 1000., SL[5].hi * 1000., "mr");
  slp("slop.total", "Em", info->slop_total_Em_used, SL[6].lo, SL[6].hi, "MeV");
  slp(" ", "Pm", 0., SL[7].lo, SL[7].hi, "MeV/c");
  o.line("");
  o.line("");                                                                   // '(/)'
  return std::fclose(f) == 0 ? SIMC_OK : SIMC_ERR_IO;
}

}  // extern "C"
