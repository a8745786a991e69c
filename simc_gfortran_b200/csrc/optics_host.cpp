// COSY table reader + compiler + arm programs (host).  See optics_host.h / arm_program.h.
#include "optics_host.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <stdexcept>

namespace simc {

// ------------------------------------------------------------------------------------------
// File reader.  The forward format is Fortran (1x,5g14.7,1x,6i1): fixed columns that may
// touch, so fields are cut by column (shared/transp.f:1200 format).
// ------------------------------------------------------------------------------------------
namespace {

struct LineReader {
  std::ifstream in;
  std::string cur;
  bool ok = false;
  explicit LineReader(const std::string& path) : in(path) {}
  bool next() {
    std::string raw;
    if (!std::getline(in, raw)) { ok = false; return false; }
    while (!raw.empty() && (raw.back() == '\r' || raw.back() == '\n')) raw.pop_back();
    if (raw.size() < 132) raw.resize(132, ' ');
    cur.swap(raw);
    ok = true;
    return true;
  }
  bool starts(const char* s) const { return cur.compare(0, std::strlen(s), s) == 0; }
  bool blank() const { return cur.find_first_not_of(' ') == std::string::npos; }
};

double field_real(const std::string& s, int col0, int width) {
  char buf[40];
  int n = 0;
  for (int i = 0; i < width && n < 39; ++i) {
    char ch = s[col0 + i];
    if (ch == ' ') continue;              // blanks are null in Fortran numeric input
    if (ch == 'D' || ch == 'd') ch = 'E';
    buf[n++] = ch;
  }
  buf[n] = 0;
  return n ? std::strtod(buf, nullptr) : 0.0;
}
int field_digit(const std::string& s, int col) { return s[col] == ' ' ? 0 : s[col] - '0'; }
double length_comment_cm(const std::string& s) { return 100.0 * std::strtod(s.c_str() + 9, nullptr); }

}  // namespace

void classify_drifts(ForwardMaps& f) {
  const double tiny = 1.0e-14;             // coeff_min, transp.f:66
  const size_t nc = f.cls.size();
  f.adrift.assign(nc, 1);
  f.driftdist_cm.assign(nc, 0.0);
  for (size_t k = 0; k < nc; ++k) {
    const CosyTerms& t = f.cls[k];
    int drift = 1;
    double dd = 0.0;
    for (int i = 0; i < t.n() && drift; ++i) {
      const double* c = &t.coef[5 * i];
      const int8_t* e = &t.expo[5 * i];
      const int order = e[0] + e[1] + e[2] + e[3];
      const double want[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
      if (order == 1) {
        int v = e[0] == 1 ? 0 : e[1] == 1 ? 1 : e[2] == 1 ? 2 : 3;
        if (v == 1) dd = 1000.0 * c[0];                       // <x|theta> carries the length
        for (int o = 0; o < 4; ++o) {
          if (v == 1 && o == 0) continue;
          if (v == 3 && o == 2) { if (std::fabs(dd - 1000.0 * c[2]) > tiny) drift = 0; continue; }
          if (std::fabs(c[o] - want[v][o]) > tiny) drift = 0;
        }
      } else if (std::fabs(c[0]) + std::fabs(c[1]) + std::fabs(c[2]) + std::fabs(c[3]) > tiny) {
        drift = 0;
      }
    }
    f.adrift[k] = drift;
    f.driftdist_cm[k] = dd;
  }
}

ForwardMaps read_forward_maps(const std::string& path) {
  LineReader r(path);
  if (!r.in) throw std::runtime_error("TRANSP_INIT: cannot open " + path);
  ForwardMaps f;
  double pending_len = 0.0;
  // header: leading '!' lines (a !LENGTH: here belongs to class 1)
  do {
    if (!r.next()) throw std::runtime_error("TRANSP_INIT: no data in " + path);
    if (r.starts("!LENGTH:")) pending_len = length_comment_cm(r.cur);
  } while (r.cur[0] == '!');
  bool in_class = true;
  CosyTerms cur;
  while (true) {
    if (in_class) {
      if (r.starts(" ---")) {
        if ((int)f.cls.size() >= kMaxClasses) throw std::runtime_error("TRANSP_INIT: too many transformations!");
        f.cls.push_back(cur);
        f.length_cm.push_back(pending_len);
        cur = CosyTerms();
        pending_len = 0.0;
        in_class = false;
      } else {
        double c[5];
        for (int i = 0; i < 5; ++i) c[i] = field_real(r.cur, 1 + 14 * i, 14);
        const int tof = field_digit(r.cur, 76);
        if (tof != 0) {
          if (c[0] != 0 || c[1] != 0 || c[2] != 0 || c[3] != 0)
            throw std::runtime_error("TRANSP_INIT: non-zero TOF terms!");
        } else {
          if (cur.n() >= 1000) throw std::runtime_error("TRANSP_INIT: too many COSY terms!");
          cur.coef.insert(cur.coef.end(), c, c + 5);
          for (int j = 0; j < 4; ++j) cur.expo.push_back((int8_t)field_digit(r.cur, 72 + j));
          cur.expo.push_back((int8_t)field_digit(r.cur, 77));
        }
      }
      if (!r.next()) {
        if (in_class) throw std::runtime_error("TRANSP_INIT: file ends inside a class: " + path);
        break;
      }
    } else {
      // between classes: comments, separators and blank lines
      if (!r.next()) break;
      if (r.starts("!LENGTH:")) pending_len = length_comment_cm(r.cur);
      if (r.cur[0] == '!' || r.starts(" ---") || r.blank()) continue;
      in_class = true;
    }
  }
  classify_drifts(f);
  return f;
}

CosyTerms read_recon_map(const std::string& path) {
  LineReader r(path);
  if (!r.in) throw std::runtime_error("MC_RECON: cannot open " + path);
  CosyTerms t;
  t.nout = 4;
  do {
    if (!r.next()) throw std::runtime_error("MC_RECON: no data in " + path);
  } while (r.cur[0] == '!');
  while (!r.starts(" ---")) {
    if (t.n() >= 1000) throw std::runtime_error("WCRECON: too many COSY terms!");
    for (int i = 0; i < 4; ++i) t.coef.push_back(field_real(r.cur, 1 + 16 * i, 16));
    for (int j = 0; j < 5; ++j) t.expo.push_back((int8_t)field_digit(r.cur, 66 + j));
    if (!r.next()) throw std::runtime_error("MC_RECON: missing terminator in " + path);
  }
  return t;
}

// ------------------------------------------------------------------------------------------
// Group compiler
// ------------------------------------------------------------------------------------------
namespace {

PolyClass compile_terms(const CosyTerms& t, std::vector<double>& recs, long long* nonzero, int block_threads) {
  PolyClass pc{};
  pc.rec_begin = (int)recs.size();
  pc.n_terms = t.n();
  const uint32_t stride = (uint32_t)block_threads * 8u;          // bytes between two table entries of one thread
  if ((uint64_t)(kPolyEntries - 1) * stride > 0xffffu) throw std::runtime_error("power-table offsets exceed 16 bits");
  auto push = [&](int e1, int e2, int e3, int e4, int e5, const double* c) {
    const uint64_t d = (uint64_t)(poly_xt_index(e1, e2) * stride) | ((uint64_t)((28 + e3) * stride) << 16) |
                       ((uint64_t)((35 + e4) * stride) << 32) | ((uint64_t)((42 + e5) * stride) << 48);
    double w;
    std::memcpy(&w, &d, sizeof(w));
    recs.push_back(w);
    for (int o = 0; o < 5; ++o) recs.push_back(c && o < t.nout ? c[o] : 0.0);
  };
  int n_rec = 0;
  for (int i = 0; i < t.n(); ++i) {
    const int8_t* e = &t.expo[5 * i];
    for (int j = 0; j < 5; ++j)
      if (e[j] < 0 || e[j] > 6) throw std::runtime_error("COSY exponent outside 0..6");
    if (e[0] + e[1] > 6) throw std::runtime_error("COSY term of degree > 6 in (x,theta)");
    int nz = 0;
    for (int o = 0; o < t.nout; ++o)
      if (t.coef[t.nout * i + o] != 0.0) ++nz;
    if (nz == 0) continue;                         // adds exact zeros in the reference
    if (nonzero) *nonzero += nz;
    push(e[0], e[1], e[2], e[3], e[4], &t.coef[t.nout * i]);
    ++n_rec;
  }
  while (n_rec % kRecChunk) { push(0, 0, 0, 0, 0, nullptr); ++n_rec; }     // 1.0 * 0.0
  pc.n_chunks = n_rec / kRecChunk;
  return pc;
}

// ---- op list helpers -------------------------------------------------------------------
}  // namespace
}  // namespace simc
#include "target.cuh"
namespace simc {
namespace {
struct Prog {
  std::vector<ArmOp> ops;
  ArmOp& add(int op, int code = 0) {
    ArmOp o{};
    o.op = op; o.code = code;
    ops.push_back(o);
    return ops.back();
  }
  void project(double z) { add(OP_PROJECT).a = z; }
  // mc_hms_coll.f:17-47 / mc_shms_coll.f:17-47: geometry, 20 steps, tungsten-alloy constants
  void collimator(double h_entr, double v_entr, double h_exit, double v_exit, double y_off, double thick, double radl,
                  int n_skip, int slit_hor_code, int coll_code) {
    ArmOp& o = add(OP_COLL, coll_code);
    o.a = h_entr; o.b = v_entr; o.c = h_exit; o.d = v_exit; o.e = y_off; o.i0 = n_skip; o.i1 = slit_hor_code;
    const MatConst mc = make_mat(17.0, 69.45, 171.56797);
    ArmOp& d1 = add(OP_COLL_DATA);
    d1.a = thick / 20; d1.b = radl; d1.c = mc.rho; d1.d = mc.CO; d1.e = mc.co27;
    ArmOp& d2 = add(OP_COLL_DATA);
    d2.a = mc.ln10; d2.b = mc.log_me_I2; d2.c = mc.p_mp; d2.d = mc.p_log; d2.e = mc.p_chsi;
  }
  void project_dd(int cls, double plus) { ArmOp& o = add(OP_PROJECT_DD); o.i0 = cls; o.a = plus; }
  void transp(int cls, double zd) { ArmOp& o = add(OP_TRANSP); o.i0 = cls; o.a = zd; }
  void cut_r2(double r, int code) { add(OP_CUT_R2, code).a = r * r; }
  void musc(double radw) { ArmOp& o = add(OP_MUSC); o.a = radw; o.b = std::sqrt(radw); }
  void musc_ext(double radw, double len) { ArmOp& o = add(OP_MUSC_EXT); o.a = radw; o.b = std::sqrt(radw); o.c = len; }
  void cut_box(double xhi, double xlo, double yhi, double ylo, int code) {
    ArmOp& o = add(OP_CUT_BOX, code); o.a = xhi; o.b = xlo; o.c = yhi; o.d = ylo;
  }
  void rot(int op, double deg, double add_after) {
    const double rad = deg * 0.017453292;            // raddeg, rotate_haxis.f:41
    ArmOp& o = add(op); o.a = std::tan(rad); o.b = std::sin(rad); o.c = std::cos(rad); o.d = add_after;
  }
  void octagon(double xoff, double yoff, double h, double v, int c_hor, int c_vert, int c_oct) {
    ArmOp& a1 = add(OP_CUT_ABS_Y, c_hor); a1.a = yoff; a1.b = h;
    ArmOp& a2 = add(OP_CUT_ABS_X, c_vert); a2.a = xoff; a2.b = v;
    ArmOp& a3 = add(OP_CUT_OCT, c_oct); a3.a = xoff; a3.b = yoff; a3.c = -v / h; a3.d = 3 * v / 2;
  }
  // the 2 x 6 drift-chamber planes, identical in mc_hms_hut.f:331-433 and mc_shms_hut.f:158-285
  // yplanes: bit (ip-1) set when plane ip measures y (its x entry is zeroed); windows = false for
  // the SOS chambers, which have no entrance/exit foils (sos/mc_sos_hut.f:303-343).
  void chamber(int jchamber, double entr_radw, double cath_radw, double gas_thick, double gas_radlen,
               double cath_thick, double wire_thick, double wire_radw, double exit_radw, double sigma,
               unsigned yplanes = 0x12u, bool windows = true) {
    if (windows) musc(entr_radw);
    for (int ip = 1; ip <= 6; ++ip) {
      musc(cath_radw);
      double drift = 0.5 * gas_thick;
      musc(drift / gas_radlen);
      drift = drift + cath_thick;
      project(drift);
      musc(wire_radw);
      ArmOp& o = add(OP_DC_PLANE);
      o.i0 = (jchamber - 1) * 6 + ip - 1; o.i1 = (yplanes >> (ip - 1)) & 1u; o.a = sigma;
      drift = 0.5 * gas_thick;
      musc(drift / gas_radlen);
      drift = drift + wire_thick;
      project(drift);
    }
    if (windows) musc(exit_radw);
  }
  int split = -1;                      // explicit survivor-compaction point (else: after the last octagon cut)
  std::vector<int> mids;               // further compaction points (where a sizeable fraction of tracks has died)
  void mid_here() { mids.push_back((int)ops.size()); }
  void cut_r(double r, int code) { add(OP_CUT_R, code).a = r; }
  void cut_abs_xy(double xmax, double ymax, int code) { cut_box(xmax, -xmax, ymax, -ymax, code); }
  // "drift the remaining distance of map class cls": length = b_target - (-(driftdist(cls) - ztmp))
  void project_to_hut(int cls, double ztmp, double target_z) {
    ArmOp& o = add(OP_PROJECT_DD); o.i0 = cls; o.a = -ztmp; o.b = target_z; o.i1 = 1;
  }
  // musc_ext over the drift just resolved from a map class (radw = drift / radlen)
  void musc_ext_prev(double radlen) { ArmOp& o = add(OP_MUSC_EXT); o.i0 = 1; o.a = radlen; }
};

const double kInf = std::numeric_limits<double>::infinity();

// ---- HMS: hms/mc_hms.f:185-437 + hms/mc_hms_hut.f:292-601 ---------------------------------
namespace hmsc {
enum { OK = 0, SLIT_HOR, SLIT_VERT, SLIT_OCT, Q1_IN, Q1_MID, Q1_OUT, Q2_IN, Q2_MID, Q2_OUT, Q3_IN, Q3_MID, Q3_OUT,
       D1_IN, D1_OUT, DC1, DC2, SCIN, CAL, COLL, N };
const char* names[] = {"ok", "slit_hor", "slit_vert", "slit_oct", "Q1_in", "Q1_mid", "Q1_out", "Q2_in", "Q2_mid",
                       "Q2_out", "Q3_in", "Q3_mid", "Q3_out", "D1_in", "D1_out", "dc1", "dc2", "scin", "cal", "coll"};
}
void build_hms(Prog& P) {
  using namespace hmsc;
  // mc_hms.f:55-95
  const double h_entr = 4.575, v_entr = 11.646, h_exit = 4.759, v_exit = 12.114;
  const double x_off = +0.000, y_off = +0.028, z_off = +40.17;
  const double z_entr = 126.2e0 + z_off, z_exit = z_entr + 6.3e0;
  const double z_dip1 = 64.77e0, z_dip2 = z_dip1 + 297.18e0, z_dip3 = z_dip2 + 115.57e0;
  const double xop = 2.8, yop = 0.0;
  const double r_Q1 = 20.50, r_Q2 = 30.22, r_Q3 = 30.22;       // apertures_hms.inc
  P.project(z_entr);
  {   // mc_hms.f:204-254: pions / muons step through the collimator when using_HMScoll, else two aperture checks
    const size_t at = P.ops.size();
    P.collimator(h_entr, v_entr, h_exit, v_exit, y_off, 6.30, 0.41753, 0, SLIT_HOR, COLL);
    const size_t first = P.ops.size();
    P.octagon(x_off, y_off, h_entr, v_entr, SLIT_HOR, SLIT_VERT, SLIT_OCT);
    P.project(z_exit - z_entr);
    P.octagon(x_off, y_off, h_exit, v_exit, SLIT_HOR, SLIT_VERT, SLIT_OCT);
    P.ops[at].i0 = (int)(P.ops.size() - first);
  }
  P.project_dd(1, -z_exit);           P.cut_r2(r_Q1, Q1_IN);
  P.transp(2, 125.233e0);             P.cut_r2(r_Q1, Q1_MID);
  P.transp(3, 62.617e0);              P.cut_r2(r_Q1, Q1_OUT);
  P.project_dd(4, 0.0);               P.cut_r2(r_Q2, Q2_IN);
  P.transp(5, 143.90e0);              P.cut_r2(r_Q2, Q2_MID);
  P.transp(6, 71.95e0);               P.cut_r2(r_Q2, Q2_OUT);
  P.project_dd(7, 0.0);               P.cut_r2(r_Q3, Q3_IN);
  P.transp(8, 143.8e0);               P.cut_r2(r_Q3, Q3_MID);
  P.transp(9, 71.9e0);                P.cut_r2(r_Q3, Q3_OUT);
  P.project_dd(10, 0.0);
  P.rot(OP_ROT_H, -6.0e0, 0.0);       P.add(OP_CUT_HMS_DIPOLE, D1_IN);
  P.transp(11, 526.053e0);
  P.rot(OP_ROT_H, 6.0e0, 0.0);        P.add(OP_CUT_HMS_DIPOLE, D1_OUT);
  { ArmOp& o = P.add(OP_CUT_HMS_PIPE, D1_OUT); o.a = xop; o.b = yop; o.c = 30.48 * 30.48; o.d = 20.5232; }
  P.project(z_dip1);
  { ArmOp& o = P.add(OP_CUT_OFF_R2, D1_OUT); o.a = xop; o.b = yop; o.c = 1145.518; }
  P.project(z_dip2 - z_dip1);
  { ArmOp& o = P.add(OP_CUT_OFF_R2, D1_OUT); o.a = xop; o.b = yop; o.c = 1512.2299; }
  P.project(z_dip3 - z_dip2);
  { ArmOp& o = P.add(OP_CUT_OFF_R2, D1_OUT); o.a = xop; o.b = yop; o.c = 2162.9383; }
  P.add(OP_MARK_HUT);

  // ---- hut, mc_hms_hut.f ----
  const double hfoil_exit_radlen = 8.90, hfoil_exit_thick = 0.011 * 2.54, hair_radlen = 30420.;
  const double hdc_entr_radlen = 28.7, hdc_entr_thick = 0.001 * 2.54, hdc_radlen = 16700.0, hdc_thick = 1.8;
  const double hdc_wire_radlen = 0.35, hdc_wire_thick = 0.0000049, hdc_cath_radlen = 7.2, hdc_cath_thick = 0.000177;
  const double hdc_exit_radlen = 28.7, hdc_exit_thick = 0.001 * 2.54;
  const double haer_entr_radlen = 8.90, haer_entr_thick = 0.15, haer_radlen = 150.0, haer_thick = 9.0;
  const double haer_air_radlen = 30420.0, haer_air_thick = 16.0, haer_exit_radlen = 8.90, haer_exit_thick = 0.1;
  const double hscin_radlen = 42.4, hcer_entr_radlen = 8.90, hcer_entr_thick = 0.040 * 2.54, hcer_radlen = 9620.0;
  const double hcer_mir_radlen = 400.0, hcer_mir_thick = 2.0, hcer_exit_radlen = 8.90, hcer_exit_thick = 0.040 * 2.54;
  const int hdc_nr_plan = 6;
  const double hdc_1_zpos = -52.1084, hdc_2_zpos = 29.2608;
  const double hdc_del_plane = hdc_thick + hdc_wire_thick + hdc_cath_thick;
  const double hdc_1_left = 26.0, hdc_1_right = -26.0, hdc_1y_offset = 1.443, hdc_1_top = -56.5, hdc_1_bot = 56.5,
               hdc_1x_offset = 1.670;
  const double hdc_2_left = 26.0, hdc_2_right = -26.0, hdc_2y_offset = 2.753, hdc_2_top = -56.5, hdc_2_bot = 56.5,
               hdc_2x_offset = 2.758;
  const double haer_zentrance = 35.699, haer_zexit = 60.949;
  const double hscin_1x_zpos = 77.830, hscin_1y_zpos = 97.520, hscin_2x_zpos = 298.820, hscin_2y_zpos = 318.510;
  const double hscin_thick = 1.067;
  const double hscin_1x_left = 37.75, hscin_1x_right = -37.75, hscin_1x_offset = -1.3;
  const double hscin_1y_top = -60.25, hscin_1y_bot = 60.25, hscin_1y_offset = -1.3;
  const double hcer_zentrance = 110.000, hcer_zmirror = 230.000, hcer_zexit = 265.000;
  const double hcal_4ta_zpos = 371.69;

  P.add(OP_RESMULT_DRAW).a = 0.15;
  // mc_hms.f:411-413: the hut starts at zinit = -(driftdist(12) - z_dip3)
  { ArmOp& o = P.add(OP_PROJECT_DD); o.i0 = 12; o.a = -z_dip3; o.b = (hdc_1_zpos - 25.000); o.i1 = 1; }
  P.musc(hfoil_exit_thick / hfoil_exit_radlen);
  double drift = (hdc_1_zpos - 0.5 * hdc_nr_plan * hdc_del_plane) - (hdc_1_zpos - 25.000);
  P.project(drift);
  P.musc_ext(drift / hair_radlen, drift);
  P.chamber(1, hdc_entr_thick / hdc_entr_radlen, hdc_cath_thick / hdc_cath_radlen, hdc_thick, hdc_radlen,
            hdc_cath_thick, hdc_wire_thick, hdc_wire_thick / hdc_wire_radlen, hdc_exit_thick / hdc_exit_radlen, 0.030);
  P.cut_box(hdc_1_bot - hdc_1x_offset, hdc_1_top - hdc_1x_offset, hdc_1_left - hdc_1y_offset,
            hdc_1_right - hdc_1y_offset, DC1);
  P.musc(hdc_cath_thick / hdc_cath_radlen);
  drift = (hdc_2_zpos - 0.5 * hdc_nr_plan * hdc_del_plane) - (hdc_1_zpos + 0.5 * hdc_nr_plan * hdc_del_plane);
  P.project(drift);
  P.musc_ext(drift / hair_radlen, drift);
  P.chamber(2, hdc_entr_thick / hdc_entr_radlen, hdc_cath_thick / hdc_cath_radlen, hdc_thick, hdc_radlen,
            hdc_cath_thick, hdc_wire_thick, hdc_wire_thick / hdc_wire_radlen, hdc_exit_thick / hdc_exit_radlen, 0.030);
  P.cut_box(hdc_2_bot - hdc_2x_offset, hdc_2_top - hdc_2x_offset, hdc_2_left - hdc_2y_offset,
            hdc_2_right - hdc_2y_offset, DC2);
  P.musc(hdc_cath_thick / hdc_cath_radlen);
  { ArmOp& o = P.add(OP_LFIT); o.a = hdc_1_zpos; o.b = hdc_2_zpos; o.c = hdc_del_plane; }
  drift = haer_zentrance - hdc_2_zpos - 0.5 * hdc_nr_plan * hdc_del_plane;
  P.project(drift);
  P.musc_ext(drift / hair_radlen, drift);
  P.musc(haer_entr_thick / haer_entr_radlen);
  drift = haer_thick;
  P.project(drift);
  P.musc_ext(drift / haer_radlen, drift);
  drift = haer_air_thick;
  P.project(drift);
  P.musc_ext(drift / haer_air_radlen, drift);
  P.musc(haer_exit_thick / haer_exit_radlen);
  auto scin = [&]() {   // every plane uses the 1x/1y sizes (mc_hms_hut.f:537-551 as written)
    ArmOp& o = P.add(OP_SCIN_COUNT);
    o.a = hscin_1x_left + hscin_1y_offset; o.b = hscin_1x_right + hscin_1y_offset;
    o.c = hscin_1y_bot + hscin_1x_offset;  o.d = hscin_1y_top + hscin_1x_offset;
  };
  drift = hscin_1x_zpos - haer_zexit;
  P.project(drift); P.musc_ext(drift / hair_radlen, drift); scin(); P.musc(hscin_thick / hscin_radlen);
  drift = hscin_1y_zpos - hscin_1x_zpos;
  P.project(drift); P.musc_ext(drift / hair_radlen, drift); scin(); P.musc(hscin_thick / hscin_radlen);
  drift = hcer_zentrance - hscin_1y_zpos;
  P.project(drift); P.musc_ext(drift / hair_radlen, drift);
  P.musc(hcer_entr_thick / hcer_entr_radlen);
  drift = hcer_zmirror - hcer_zentrance;
  P.project(drift); P.musc_ext(drift / hcer_radlen, drift);
  P.musc(hcer_mir_thick / hcer_mir_radlen);
  drift = hcer_zexit - hcer_zmirror;
  P.musc(hcer_exit_thick / hcer_exit_radlen);
  P.project(drift);
  P.musc(hcer_exit_thick / hcer_exit_radlen);
  drift = hscin_2x_zpos - hcer_zexit;
  P.project(drift); P.musc_ext(drift / hair_radlen, drift); scin(); P.musc(hscin_thick / hscin_radlen);
  drift = hscin_2y_zpos - hscin_2x_zpos;
  P.project(drift); P.musc_ext(drift / hair_radlen, drift); scin(); P.musc(hscin_thick / hscin_radlen);
  P.add(OP_SCIN_TRIG, SCIN).i0 = 3;
  drift = hcal_4ta_zpos - hscin_2y_zpos;
  P.project(drift); P.musc_ext(drift / hair_radlen, drift);
  P.add(OP_RECON);
  P.add(OP_END);
}

// ---- SHMS: shms/mc_shms.f:414-1099 + shms/mc_shms_hut.f:73-447 + shms/hut.inc ---------------
namespace shmsc {
enum { OK = 0, HB_IN, HB_MEN, HB_MEX, HB_OUT, SLIT_HOR, SLIT_VERT, SLIT_OCT, Q1_IN, Q1_MEN, Q1_MID, Q1_MEX, Q1_OUT,
       Q2_IN, Q2_MEN, Q2_MID, Q2_MEX, Q2_OUT, Q3_IN, Q3_MEN, Q3_MID, Q3_MEX, Q3_OUT, D1_IN, D1_FLR, D1_MEN,
       D1_MID1, D1_MID2, D1_MID3, D1_MID4, D1_MID5, D1_MID6, D1_MID7, D1_MEX, D1_OUT, DC1, DC2, S1X, S1Y, S2X, S2Y,
       CAL, CAL_FID, COLL, N };
const char* names[] = {"ok", "HB_in", "HB_men", "HB_mex", "HB_out", "slit_hor", "slit_vert", "slit_oct", "Q1_in",
                       "Q1_men", "Q1_mid", "Q1_mex", "Q1_out", "Q2_in", "Q2_men", "Q2_mid", "Q2_mex", "Q2_out",
                       "Q3_in", "Q3_men", "Q3_mid", "Q3_mex", "Q3_out", "D1_in", "D1_flr", "D1_men", "D1_mid1",
                       "D1_mid2", "D1_mid3", "D1_mid4", "D1_mid5", "D1_mid6", "D1_mid7", "D1_mex", "D1_out", "dc1",
                       "dc2", "s1x", "s1y", "s2x", "s2y", "cal", "cal_fid", "coll"};
}
void build_shms(Prog& P) {
  using namespace shmsc;
  const double r_HBx = 11.2, r_HBfym = -4.13, r_HBfyp = 11.75, r_HBmenym = -5.45, r_HBmenyp = 11.74;
  const double r_HBmexym = -10.25, r_HBmexyp = 11.71, r_HBbym = -11.71, r_HBbyp = 11.70;
  const double r_Q1 = 20.00, r_Q2 = 30.00, r_Q3 = 30.00, r_D1 = 30.00;
  const double h_entr = 8.5, v_entr = 12.5, h_exit = 8.65, v_exit = 12.85, x_off = +0.00, y_off = +0.00;
  const double zd_hbin = 118.39, zd_hbmen = 17.61, zd_hbmex = 80.0, zd_hbout = 17.61, z_entr = 25.189, z_thick = 6.35;
  const double zd_q1in = 58.39, zd_q1men = 28.35, zd_q1mid = 93.65, zd_q1mex = 93.65, zd_q1out = 28.35;
  const double zd_q2in = 25.55, zd_q2men = 39.1, zd_q2mid = 79.35, zd_q2mex = 79.35, zd_q2out = 39.1;
  const double zd_q3in = 28.10, zd_q3men = 39.1, zd_q3mid = 79.35, zd_q3mex = 79.35, zd_q3out = 39.1;
  const double zd_q3d1trans = 18.00, zd_d1flare = 30.10, zd_d1men = 39.47, zd_d1mid = 36.406263,
               zd_d1mex = 36.406263, zd_d1out = 60.68, zd_fp = 307.95;
  auto hb = [&](double yadd, double yp, double ym, double deg, int code) {
    P.rot(OP_ROT_V, deg, yadd);
    ArmOp& o = P.add(OP_CUT_HB, code); o.a = r_HBx * r_HBx; o.b = yp; o.c = ym;
  };
  P.project(zd_hbin);              hb(1.51, r_HBfyp, r_HBfym, 1.5, HB_IN);
  P.project(zd_hbmen);             hb(0.98, r_HBmenyp, r_HBmenym, 1.5, HB_MEN);
  P.transp(3, zd_hbmex);           hb(0.98, r_HBmexyp, r_HBmexym, -1.5, HB_MEX);
  P.project(zd_hbout);             hb(1.51, r_HBbyp, r_HBbym, -1.5, HB_OUT);
  P.project(z_entr);               // first statement of both branches of shms/mc_shms.f:503-560
  {   // pions / muons step through the collimator when using_SHMScoll (mc_shms_coll), else two aperture checks
    const size_t at = P.ops.size();
    P.collimator(8.50, 12.50, 8.65, 12.85, 0.000, 6.35, 0.42084, 0, SLIT_HOR, COLL);
    const size_t first = P.ops.size();
    P.octagon(x_off, y_off, h_entr, v_entr, SLIT_HOR, SLIT_VERT, SLIT_OCT);
    P.project(z_thick);
    // exit side is written ((-v_exit)/(h_exit)*|y| + 3*(v_exit)/2): same arithmetic as the entrance form
    P.octagon(x_off, y_off, h_exit, v_exit, SLIT_HOR, SLIT_VERT, SLIT_OCT);
    P.ops[at].i0 = (int)(P.ops.size() - first);
  }
  P.project(zd_q1in - z_entr - z_thick); P.cut_r2(r_Q1, Q1_IN);
  P.project(zd_q1men);             P.cut_r2(r_Q1, Q1_MEN);
  P.transp(7, zd_q1mid);           P.cut_r2(r_Q1, Q1_MID);
  P.transp(8, zd_q1mex);           P.cut_r2(r_Q1, Q1_MEX);
  P.project(zd_q1out);             P.cut_r2(r_Q1, Q1_OUT);
  P.mid_here();                    // a quarter of the tracks that pass the slit still die inside Q1 (C1: 14 k of 60 k)
  P.project(zd_q2in);              P.cut_r2(r_Q2, Q2_IN);
  P.project(zd_q2men);             P.cut_r2(r_Q2, Q2_MEN);
  P.transp(12, zd_q2mid);          P.cut_r2(r_Q2, Q2_MID);
  P.transp(13, zd_q2mex);          P.cut_r2(r_Q2, Q2_MEX);
  P.project(zd_q2out);             P.cut_r2(r_Q2, Q2_OUT);
  P.project(zd_q3in);              P.cut_r2(r_Q3, Q3_IN);
  P.project(zd_q3men);             P.cut_r2(r_Q3, Q3_MEN);
  P.transp(17, zd_q3mid);          P.cut_r2(r_Q3, Q3_MID);
  P.transp(18, zd_q3mex);          P.cut_r2(r_Q3, Q3_MEX);
  P.project(zd_q3out);             P.cut_r2(r_Q3, Q3_OUT);
  P.project(zd_q3d1trans);         P.cut_r2(r_D1, D1_IN);
  auto tilted = [&](double deg, double xadd, int code) {
    P.rot(OP_ROT_H, deg, xadd);
    P.add(OP_CUT_T_R2, code).a = r_D1 * r_D1;
  };
  P.project(zd_d1flare);           tilted(9.200, -3.5, D1_FLR);
  P.project(zd_d1men);             tilted(9.200, 2.82, D1_MEN);
  const double ang[8] = {6.9, 4.6, 2.3, 0.0, -2.3, -4.6, -6.9, -9.2};
  const double off[8] = {8.05, 11.75, 13.96, 14.70, 13.96, 11.75, 8.05, 2.82};
  for (int k = 0; k < 8; ++k) { P.transp(23 + k, k < 7 ? zd_d1mid : zd_d1mex); tilted(ang[k], off[k], D1_MID1 + k); }
  P.project(zd_d1out);             tilted(-9.20, -6.88, D1_OUT);

  // hut.inc
  const double hfoil_exit_radlen = 8.89, hfoil_exit_thick = 0.020 * 2.54, hair_radlen = 30420.;
  const double hdc_entr_radlen = 28.7, hdc_entr_thick = 0.001 * 2.54, hdc_radlen = 16700.0, hdc_thick = 0.125 * 2.54;
  const double hdc_wire_radlen = 0.35, hdc_wire_thick = 0.0000354, hdc_cath_radlen = 28.6, hdc_cath_thick = 0.001 * 2.54;
  const double hdc_exit_radlen = 28.7, hdc_exit_thick = 0.001 * 2.54, hscin_radlen = 42.4;
  const double hcer_entr_radlen = 19.63, hcer_entr_thick = 0.002 * 2.54, hcer_1_radlen = 11700.0;
  const double hcer_mirglass_radlen = 12.29, hcer_mirglass_thick = 0.3, hcer_exit_radlen = 19.63,
               hcer_exit_thick = 0.002 * 2.54;
  const double hcer_2_entr_radlen = 8.90, hcer_2_entr_thick = 0.040 * 2.54, hcer_2_radlen = 1202.5;
  const double hcer_mir_radlen = 400., hcer_mir_thick = 2.00, hcer_2_exit_radlen = 8.90, hcer_2_exit_thick = 0.040 * 2.54;
  const int hdc_nr_plan = 6;
  const double hdc_1_zpos = -40.656, hdc_2_zpos = 39.332;
  const double hdc_1_left = 40.0, hdc_1_right = -40.0, hdc_1y_offset = 0.0, hdc_1_top = -40., hdc_1_bot = 40.,
               hdc_1x_offset = 0.0;
  const double hdc_2_left = 40.0, hdc_2_right = -40.0, hdc_2y_offset = 0.0, hdc_2_top = -40., hdc_2_bot = 40.,
               hdc_2x_offset = 0.;
  const double hscin_1x_zpos = 52.1, hscin_1y_zpos = 61.7, hscin_2x_zpos = 271.4, hscin_2y_zpos = 282.4;
  const double hscin_thick = 1.000 * 1.067;
  const double hscin_1x_left = 50., hscin_1x_right = -50., hscin_1x_offset = 0.0;
  const double hscin_1y_top = -45., hscin_1y_bot = 45., hscin_1y_offset = 0.0;
  const double hscin_2x_left = 55., hscin_2x_right = -55., hscin_2x_offset = 0.;
  const double hscin_2y_top = -62.5, hscin_2y_bot = 62.5, hscin_2y_offset = 0;
  const double hcer_1_zentrance = -291.700, hcer_1_zmirror = -84.900, hcer_1_zexit = -61.700;
  const double hcer_2_zentrance = 72.600, hcer_2_zmirror = 179.400, hcer_2_zexit = 202.600;
  const double hcal_4ta_zpos = 341.0, hcal_left = 63.00, hcal_right = -63.00, hcal_top = -70.00, hcal_bottom = 70.00;
  const double hdc_del_plane = hdc_thick + hdc_wire_thick + hdc_cath_thick;

  P.mid_here();                    // 10-25 % more are lost in the dipole: the hut (2/3 of the remaining work) runs on full warps
  P.project(zd_fp + hcer_1_zentrance);                     // cer_flag = .true. (mc_shms.f:352,1045)
  P.add(OP_MARK_HUT);
  P.add(OP_RESMULT_ONE);
  P.musc(hfoil_exit_thick / hfoil_exit_radlen);
  P.musc(hcer_entr_thick / hcer_entr_radlen);
  double drift = hcer_1_zmirror - hcer_1_zentrance - hcer_mirglass_thick / 2;
  P.project(drift); P.musc_ext(drift / hcer_1_radlen, drift);
  P.musc(hcer_mirglass_thick / hcer_mirglass_radlen);
  drift = hcer_1_zexit - hcer_1_zmirror - hcer_mirglass_thick / 2;
  P.project(drift); P.musc_ext(drift / hcer_1_radlen, drift);
  P.musc(hcer_exit_thick / hcer_exit_radlen);
  drift = (hdc_1_zpos - 0.5 * hdc_nr_plan * hdc_del_plane) - hcer_1_zexit;
  P.project(drift); P.musc_ext(drift / hair_radlen, drift);
  P.chamber(1, hdc_entr_thick / hdc_entr_radlen, hdc_cath_thick / hdc_cath_radlen, hdc_thick, hdc_radlen,
            hdc_cath_thick, hdc_wire_thick, hdc_wire_thick / hdc_wire_radlen, hdc_exit_thick / hdc_exit_radlen, 0.020);
  P.cut_box(hdc_1_bot - hdc_1x_offset, hdc_1_top - hdc_1x_offset, hdc_1_left - hdc_1y_offset,
            hdc_1_right - hdc_1y_offset, DC1);
  P.musc(hdc_cath_thick / hdc_cath_radlen);
  drift = hdc_2_zpos - hdc_1_zpos - hdc_nr_plan * hdc_del_plane;
  P.project(drift); P.musc_ext(drift / hair_radlen, drift);
  P.chamber(2, hdc_entr_thick / hdc_entr_radlen, hdc_cath_thick / hdc_cath_radlen, hdc_thick, hdc_radlen,
            hdc_cath_thick, hdc_wire_thick, hdc_wire_thick / hdc_wire_radlen, hdc_exit_thick / hdc_exit_radlen, 0.020);
  P.cut_box(hdc_2_bot - hdc_2x_offset, hdc_2_top - hdc_2x_offset, hdc_2_left - hdc_2y_offset,
            hdc_2_right - hdc_2y_offset, DC2);
  P.musc(hdc_cath_thick / hdc_cath_radlen);
  drift = hscin_1x_zpos - hdc_2_zpos - 0.5 * hdc_nr_plan * hdc_del_plane;
  P.project(drift); P.musc_ext(drift / hair_radlen, drift);
  P.cut_box(kInf, -kInf, hscin_1x_left + hscin_1y_offset, hscin_1x_right + hscin_1y_offset, S1X);
  P.musc(hscin_thick / hscin_radlen);
  drift = hscin_1y_zpos - hscin_1x_zpos;
  P.project(drift); P.musc_ext(drift / hair_radlen, drift);
  P.cut_box(hscin_1y_bot + hscin_1x_offset, hscin_1y_top + hscin_1x_offset, kInf, -kInf, S1Y);
  P.musc(hscin_thick / hscin_radlen);
  drift = hcer_2_zentrance - hscin_1y_zpos;
  P.project(drift); P.musc_ext(drift / hair_radlen, drift);
  P.musc(hcer_2_entr_thick / hcer_2_entr_radlen);
  drift = hcer_2_zmirror - hcer_2_zentrance;
  P.project(drift); P.musc_ext(drift / hcer_2_radlen, drift);
  P.musc(hcer_mir_thick / hcer_mir_radlen);
  drift = hcer_2_zexit - hcer_2_zmirror;
  P.project(drift); P.musc_ext(drift / hcer_2_radlen, drift);
  P.musc(hcer_2_exit_thick / hcer_2_exit_radlen);
  drift = hscin_2x_zpos - hcer_2_zexit;
  P.project(drift); P.musc_ext(drift / hair_radlen, drift);
  P.cut_box(kInf, -kInf, hscin_2x_left + hscin_2y_offset, hscin_2x_right + hscin_2y_offset, S2X);
  P.musc(hscin_thick / hscin_radlen);
  drift = hscin_2y_zpos - hscin_2x_zpos;
  P.project(drift); P.musc_ext(drift / hair_radlen, drift);
  P.cut_box(hscin_2y_bot + hscin_2x_offset, hscin_2y_top + hscin_2x_offset, kInf, -kInf, S2Y);
  P.musc(hscin_thick / hscin_radlen);
  drift = hcal_4ta_zpos - hscin_2y_zpos;
  P.project(drift); P.musc_ext(drift / hair_radlen, drift);
  P.cut_box(hcal_bottom, hcal_top, hcal_left, hcal_right, CAL);
  { ArmOp& o = P.add(OP_LFIT); o.a = hdc_1_zpos; o.b = hdc_2_zpos; o.c = hdc_del_plane; }
  { ArmOp& o = P.add(OP_CUT_FP_CAL, CAL_FID); o.a = hcal_4ta_zpos; o.b = hcal_left - 5.0; o.c = hcal_right + 5.0;
    o.d = hcal_bottom - 5.0; o.e = hcal_top + 5.0; }
  P.add(OP_RECON);
  P.add(OP_END);
}

// ---- SOS: sos/mc_sos.f:115-391 + sos/mc_sos_hut.f:251-545 + sos/apertures_sos.inc -----------
namespace sosc {
enum { OK = 0, SLIT_HOR, SLIT_VERT, SLIT_OCT, QUAD_IN, QUAD_MID, QUAD_OUT, BM01_IN, BM01_OUT, BM02_IN, BM02_OUT, EXIT,
       DC1, DC2, SCIN, N };
const char* names[] = {"ok", "slit_hor", "slit_vert", "slit_oct", "quad_in", "quad_mid", "quad_out", "bm01_in",
                       "bm01_out", "bm02_in", "bm02_out", "exit", "dc1", "dc2", "scin"};
}
void build_sos(Prog& P) {
  using namespace sosc;
  const double r2_quad = 163.84, w_bm01 = 8.0, w_bm02 = 8.0;
  const double t_bm01_in = 41.32, b_bm01_in = -30.52, t_bm01_out = 53.05, b_bm01_out = -65.07;
  const double t_bm02_in = 51.73, b_bm02_in = -66.38, t_bm02_out = 51.52, b_bm02_out = -55.85;
  const double w_exit = 8.57, t_exit = 51.52, b_exit = -55.85;
  const double h_entr = 7.201, v_entr = 4.696, h_exit = 7.567, v_exit = 4.935;
  const double z_entr = 126.3e0, z_exit = z_entr + 6.3e0;
  P.project(z_entr);
  P.octagon(0.0, 0.0, h_entr, v_entr, SLIT_HOR, SLIT_VERT, SLIT_OCT);   // |xs - 0.0| == |xs| exactly
  P.project(z_exit - z_entr);
  P.octagon(0.0, 0.0, h_exit, v_exit, SLIT_HOR, SLIT_VERT, SLIT_OCT);
  P.project_dd(1, -z_exit);           P.add(OP_CUT_R2, QUAD_IN).a = r2_quad;
  P.transp(2, 35.0e0);                P.add(OP_CUT_R2, QUAD_MID).a = r2_quad;
  P.transp(3, 35.0e0);                P.add(OP_CUT_R2, QUAD_OUT).a = r2_quad;
  auto tbox = [&](double deg, double yhi, double ylo, double top, double bot, int code) {
    P.rot(OP_ROT_H, deg, 0.0);
    ArmOp& o = P.add(OP_CUT_T_BOX, code); o.a = yhi; o.b = ylo; o.c = top; o.d = bot;
  };
  // |yt| > w  <=>  yt > w or -yt > w
  P.transp(4, 80.0e0);                tbox(-45.0e0, w_bm01, w_bm01 - 0.05 * 2.54, t_bm01_in, b_bm01_in, BM01_IN);
  P.transp(5, 169.52e0);              tbox(45.0e0, w_bm01, w_bm01, t_bm01_out, b_bm01_out, BM01_OUT);
  P.transp(6, 80.80e0);               tbox(49.0e0, w_bm02, w_bm02, t_bm02_in, b_bm02_in, BM02_IN);
  P.transp(7, 77.06e0);               tbox(57.0e0, w_bm02, w_bm02, t_bm02_out, b_bm02_out, BM02_OUT);
  P.transp(8, 43.82e0);               tbox(45.0e0, w_exit, w_exit, t_exit, b_exit, EXIT);
  P.transp(9, 44.34e0);
  { ArmOp& o = P.add(OP_CUT_SOS_EXIT, EXIT); o.a = 10.998; o.b = 0.10209; o.c = 37.694; }
  P.add(OP_SHIFT).a = 7.62;
  P.cut_abs_xy(38.1, 12.7, EXIT);
  P.add(OP_MARK_HUT);

  // ---- hut ----
  const double sfoil_exit_radlen = 53.3, sfoil_exit_thick = 0.020 * 2.54, sfoil_exit_zpos = -3.22;
  const double sair_radlen = 30420., sdc_radlen = 16700.0, sdc_thick = 0.61775;
  const double sdc_wire_radlen = 0.35, sdc_wire_thick = 0.0000354, sdc_cath_radlen = 28.7, sdc_cath_thick = 0.0005 * 2.54;
  const double sscin_radlen = 42.4, scer_entr_radlen = 8.90, scer_entr_thick = 0.050, scer_radlen = 4810.0;
  const double scer_mir_radlen = 400.0, scer_mir_thick = 2.0, scer_exit_radlen = 8.90, scer_exit_thick = 0.050;
  const int sdc_nr_plan = 6;
  const double sdc_1_zpos = 6.25, sdc_2_zpos = 55.77;
  const double sdc_del_plane = sdc_thick + sdc_wire_thick + sdc_cath_thick;
  const double sdc_1_left = 24.0, sdc_1_right = -24.0, sdc_1y_offset = -1.822, sdc_1_top = -32.0, sdc_1_bot = 32.0,
               sdc_1x_offset = 8.649;
  const double sdc_2_left = 24.0, sdc_2_right = -24.0, sdc_2y_offset = -1.976, sdc_2_top = -32.0, sdc_2_bot = 32.0,
               sdc_2x_offset = -1.532;
  const double sscin_1y_zpos = 73.61, sscin_1x_zpos = 97.11, sscin_2y_zpos = 249.51, sscin_2x_zpos = 290.81;
  const double sscin_1x_thick = 1.040, sscin_1y_thick = 1.098, sscin_2x_thick = 1.040, sscin_2y_thick = 1.098;
  const double sscin_1_left = 18.25, sscin_1_right = -18.25, sscin_1x_offset = 2.8, sscin_1_top = -31.75,
               sscin_1_bot = 31.75, sscin_1y_offset = 2.25;
  const double sscin_2_left = 18.25, sscin_2_right = -18.25, sscin_2x_offset = 4.9, sscin_2_top = -56.25,
               sscin_2_bot = 56.25, sscin_2y_offset = 2.9;
  const double scer_zentrance = 130.000, scer_zmirror = 155.000, scer_zexit = 160.000, scal_4ta_zpos = 346.01;
  const double zinit = -3.206267e0;                     // mc_sos.f:369

  P.add(OP_RESMULT_DRAW).a = 0.15;
  // mc_sos_hut.f:282-287: the foil is untilted here (sfoil_exit_ang = 0), so xt*tan(ang) adds 0
  const double foil_z = sfoil_exit_zpos + 0.0;
  double drift = foil_z - zinit;
  if (drift <= 0.001) drift = 0.001;
  P.project(drift);
  P.musc(sfoil_exit_thick / sfoil_exit_radlen / 1.0);
  drift = (sdc_1_zpos - 0.5 * sdc_nr_plan * sdc_del_plane) - foil_z;
  P.project(drift); P.musc_ext(drift / sair_radlen, drift);
  P.chamber(1, 0., sdc_cath_thick / sdc_cath_radlen, sdc_thick, sdc_radlen, sdc_cath_thick, sdc_wire_thick,
            sdc_wire_thick / sdc_wire_radlen, 0., 0.030, 0x15u, false);
  P.cut_box(sdc_1_bot - sdc_1x_offset, sdc_1_top - sdc_1x_offset, sdc_1_left - sdc_1y_offset,
            sdc_1_right - sdc_1y_offset, DC1);
  P.musc(sdc_cath_thick / sdc_cath_radlen);
  drift = sdc_2_zpos - sdc_1_zpos - sdc_nr_plan * sdc_del_plane;
  P.project(drift); P.musc_ext(drift / sair_radlen, drift);
  P.chamber(2, 0., sdc_cath_thick / sdc_cath_radlen, sdc_thick, sdc_radlen, sdc_cath_thick, sdc_wire_thick,
            sdc_wire_thick / sdc_wire_radlen, 0., 0.030, 0x15u, false);
  P.cut_box(sdc_2_bot - sdc_2x_offset, sdc_2_top - sdc_2x_offset, sdc_2_left - sdc_2y_offset,
            sdc_2_right - sdc_2y_offset, DC2);
  P.musc(sdc_cath_thick / sdc_cath_radlen);
  { ArmOp& o = P.add(OP_LFIT); o.a = sdc_1_zpos; o.b = sdc_2_zpos; o.c = sdc_del_plane; }
  auto scin = [&](int plane) {
    ArmOp& o = P.add(OP_SCIN_COUNT);
    if (plane == 1) {
      o.a = sscin_1_left + sscin_1y_offset; o.b = sscin_1_right + sscin_1y_offset;
      o.c = sscin_1_bot + sscin_1x_offset;  o.d = sscin_1_top + sscin_1x_offset;
    } else {
      o.a = sscin_2_left + sscin_2y_offset; o.b = sscin_2_right + sscin_2y_offset;
      o.c = sscin_2_bot + sscin_2x_offset;  o.d = sscin_2_top + sscin_2x_offset;
    }
  };
  drift = sscin_1y_zpos - sdc_2_zpos - 0.5 * sdc_nr_plan * sdc_del_plane;
  P.project(drift); P.musc_ext(drift / sair_radlen, drift); scin(1); P.musc(sscin_1y_thick / sscin_radlen);
  drift = sscin_1x_zpos - sscin_1y_zpos;
  P.project(drift); P.musc_ext(drift / sair_radlen, drift); scin(1); P.musc(sscin_1x_thick / sscin_radlen);
  drift = scer_zentrance - sscin_1x_zpos;
  P.project(drift); P.musc_ext(drift / sair_radlen, drift);
  P.musc(scer_entr_thick / scer_entr_radlen);
  drift = scer_zmirror - scer_zentrance;
  P.project(drift); P.musc_ext(drift / scer_radlen, drift);
  P.musc(scer_mir_thick / scer_mir_radlen);
  drift = scer_zexit - scer_zmirror;
  P.project(drift); P.musc_ext(drift / scer_radlen, drift);
  P.musc(scer_exit_thick / scer_exit_radlen);
  drift = sscin_2y_zpos - scer_zexit;
  P.project(drift); P.musc_ext(drift / sair_radlen, drift); scin(2); P.musc(sscin_2y_thick / sscin_radlen);
  drift = sscin_2x_zpos - sscin_2y_zpos;
  P.project(drift); P.musc_ext(drift / sair_radlen, drift); scin(2); P.musc(sscin_2x_thick / sscin_radlen);
  P.add(OP_SCIN_TRIG, SCIN).i0 = 3;
  drift = scal_4ta_zpos - sscin_2x_zpos;
  P.project(drift); P.musc_ext(drift / sair_radlen, drift);
  P.add(OP_RECON).i0 = 1;
  P.add(OP_END);
}

// ---- HRS: hrsl/mc_hrsl.f:114-546 (hrsr/mc_hrsr.f) + hrs?/mc_hrs?_hut.f -----------------------
namespace hrsc {
enum { OK = 0, SLIT_HOR, SLIT_VERT, Q1_IN, Q1_MID, Q1_OUT, Q2_IN, Q2_MID, Q2_OUT, D1_IN, D1_OUT, Q3_IN, Q3_MID, Q3_OUT,
       DC1, DC2, S1, S2, N };
const char* names[] = {"ok", "slit_hor", "slit_vert", "Q1_in", "Q1_mid", "Q1_out", "Q2_in", "Q2_mid", "Q2_out",
                       "D1_in", "D1_out", "Q3_in", "Q3_mid", "Q3_out", "dc1", "dc2", "s1", "s2"};
}
void build_hrs(Prog& P, bool right) {
  using namespace hrsc;
  const double r_Q1 = 15.0, r_Q2 = 30.22, r_Q3 = 30.22;
  const double h_entr = 3.145, v_entr = 6.090, h_exit = right ? 3.340 : 3.335, v_exit = 6.485;
  const double y_off = 0.0, x_off = 0.0, z_off = 0.0;
  const double z_entr = (right ? 110.0 : 110.9) + z_off, z_exit = z_entr + 8.0;
  double zdrift, ztmp;
  auto slit = [&](double h, double v) {
    ArmOp& a1 = P.add(OP_CUT_ABS_Y, SLIT_HOR); a1.a = y_off; a1.b = h;
    ArmOp& a2 = P.add(OP_CUT_ABS_X, SLIT_VERT); a2.a = x_off; a2.b = v;
  };
  zdrift = 65.686; ztmp = zdrift;
  P.project(zdrift);                  P.cut_r(7.3787, SLIT_HOR);
  zdrift = 80.436 - ztmp; ztmp = 80.436;
  P.project(zdrift);                  P.cut_r(7.4092, SLIT_HOR);
  zdrift = z_entr - ztmp;
  P.project(zdrift);                  slit(h_entr, v_entr);
  zdrift = z_exit - z_entr;
  P.project(zdrift);                  slit(h_exit, v_exit);
  P.split = (int)P.ops.size();
  ztmp = 135.064; zdrift = ztmp - z_exit;
  P.project(zdrift);                  P.cut_r(12.5222, Q1_IN);
  P.project_dd(1, -ztmp);             P.cut_r2(r_Q1, Q1_IN);
  P.transp(2, 62.75333333);           P.cut_r2(r_Q1, Q1_MID);
  P.transp(3, 31.37666667);           P.cut_r2(r_Q1, Q1_OUT);
  zdrift = 300.464 - 253.16; ztmp = zdrift;
  P.project(zdrift);                  P.cut_r(14.9225, Q1_OUT);
  zdrift = 314.464 - 300.464; ztmp = ztmp + zdrift;
  P.project(zdrift);                  P.cut_r(20.9550, Q2_IN);
  P.project_dd(4, -ztmp);             P.cut_r2(r_Q2, Q2_IN);
  P.transp(5, 121.77333333);          P.cut_r2(r_Q2, Q2_MID);
  P.transp(6, 60.88666667);           P.cut_r2(r_Q2, Q2_OUT);
  zdrift = 609.664 - 553.020; ztmp = zdrift;
  P.project(zdrift);                  P.cut_r(30.0073, Q2_OUT);
  zdrift = 641.800 - 609.664; ztmp = ztmp + zdrift;
  P.project(zdrift);                  P.cut_r(30.0073, Q2_OUT);
  zdrift = 819.489 - 641.800; ztmp = ztmp + zdrift;
  P.project(zdrift);                  P.cut_abs_xy(50.0, 15.0, D1_IN);
  P.project_dd(7, -ztmp);
  auto dipole_face = [&](double deg, int code) {
    P.rot(OP_ROT_H, deg, 0.0);
    { ArmOp& o = P.add(OP_CUT_T_ABSX, code); o.a = 2.500; o.b = 52.5; }
    { ArmOp& o = P.add(OP_CUT_T_TRAP, code); o.a = 0.01861; o.b = 12.5; }
  };
  dipole_face(-30.0, D1_IN);
  P.transp(8, 659.73445725);          dipole_face(30.0, D1_OUT);
  zdrift = 1745.33546 - 1655.83446; ztmp = zdrift;
  P.project(zdrift);                  P.cut_r(30.3276, D1_OUT); P.cut_abs_xy(50.0, 15.0, D1_OUT);
  P.mid_here();                       // C5: 2/3 of the tracks that pass the slit are gone behind the dipole
  zdrift = 1759.00946 - 1745.33546; ztmp = ztmp + zdrift;
  P.project(zdrift);                  P.cut_r(30.3276, Q3_IN);
  P.project_dd(9, -ztmp);             P.cut_r2(r_Q3, Q3_IN);
  P.transp(10, 121.7866667);          P.cut_r2(r_Q3, Q3_MID);
  P.transp(11, 60.89333333);          P.cut_r2(r_Q3, Q3_OUT);
  zdrift = 2080.38746 - 1997.76446; ztmp = zdrift;
  P.project(zdrift);                  P.cut_abs_xy(35.56, 17.145, Q3_OUT);
  zdrift = 2327.47246 - 2080.38746; ztmp = ztmp + zdrift;
  P.project(zdrift);                  P.cut_abs_xy(99.76635, 17.145, Q3_OUT);
  P.mid_here();                       // and half of the rest in Q3
  P.add(OP_MARK_HUT);

  // ---- hut ----
  const double hfoil_exit_radlen = 3.56, hfoil_exit_thick = 0.01, hair_radlen = 30420.;
  const double hdc_entr_radlen = 34.4, hdc_entr_thick = 0.00018 * 2.54, hdc_radlen = 16700.0, hdc_thick = 1.5;
  const double hdc_wire_radlen = 0.35, hdc_wire_thick = 0.0000049, hdc_cath_radlen = 7.2, hdc_cath_thick = 0.000177;
  const double hdc_exit_radlen = 34.4, hdc_exit_thick = 0.00018 * 2.54, hscin_radlen = 42.4;
  const double hcer_entr_radlen = 8.90, hcer_entr_thick = 0.040 * 2.54, hcer_radlen = 36620.0;
  const double hcer_mir_radlen = 400.0, hcer_mir_thick = 2.0, hcer_exit_radlen = 8.90, hcer_exit_thick = 0.040 * 2.54;
  const int hdc_nr_plan = 6;
  const double hdc_1_zpos = -25.0 + 25.0, hdc_2_zpos = 25.0 + 25.0;
  const double hdc_del_plane = hdc_thick + hdc_wire_thick + hdc_cath_thick;
  const double hdc_left = 14.4, hdc_right = -14.4, hdc_y_offset = 0.000, hdc_top = -105.6, hdc_bot = 105.6,
               hdc_x_offset = 0.000;
  const double hscin_1x_zpos = 95.0 + 25.0, hscin_2x_zpos = 288.3 + 25.0;
  const double hscin_1x_thick = 0.5 * 1.067, hscin_2x_thick = 0.5 * 1.067;
  const double hscin_1x_left = 18.0, hscin_1x_right = -18.0, hscin_2x_left = 30.0, hscin_2x_right = -30.0;
  const double hcer_zentrance = 137.0 + 25.0, hcer_zmirror = 197.0 + 25.0, hcer_zexit = 237.0 + 25.0;
  const double hcal_4ta_zpos = 407.3 + 25.0;

  P.add(OP_RESMULT_ONE);
  P.musc(hfoil_exit_thick / hfoil_exit_radlen);
  // mc_hrsl.f:507-509: zinit = -(driftdist(12) - ztmp); drift = (first VDC plane) - zinit
  P.project_to_hut(12, ztmp, hdc_1_zpos - 0.5 * hdc_nr_plan * hdc_del_plane);
  P.musc_ext_prev(hair_radlen);
  auto vdc_frame = [&](int code) {    // tested in the chamber plane, rotated by 45 degrees
    P.rot(OP_ROT_H, 45.0, 0.0);
    ArmOp& o = P.add(OP_CUT_T_RECT, code);
    o.a = hdc_bot - hdc_x_offset; o.b = hdc_top - hdc_x_offset; o.c = hdc_left - hdc_y_offset; o.d = hdc_right - hdc_y_offset;
  };
  P.chamber(1, hdc_entr_thick / hdc_entr_radlen, hdc_cath_thick / hdc_cath_radlen, hdc_thick, hdc_radlen,
            hdc_cath_thick, hdc_wire_thick, hdc_wire_thick / hdc_wire_radlen, hdc_exit_thick / hdc_exit_radlen, 0.0225,
            0x15u);
  vdc_frame(DC1);
  P.musc(hdc_cath_thick / hdc_cath_radlen);
  double drift = hdc_2_zpos - hdc_1_zpos - hdc_nr_plan * hdc_del_plane;
  P.project(drift); P.musc_ext(drift / hair_radlen, drift);
  P.chamber(2, hdc_entr_thick / hdc_entr_radlen, hdc_cath_thick / hdc_cath_radlen, hdc_thick, hdc_radlen,
            hdc_cath_thick, hdc_wire_thick, hdc_wire_thick / hdc_wire_radlen, hdc_exit_thick / hdc_exit_radlen, 0.0225,
            0x15u);
  vdc_frame(DC2);
  P.musc(hdc_cath_thick / hdc_cath_radlen);
  { ArmOp& o = P.add(OP_LFIT); o.a = hdc_1_zpos; o.b = hdc_2_zpos; o.c = hdc_del_plane; }
  drift = hscin_1x_zpos - hdc_2_zpos - 0.5 * hdc_nr_plan * hdc_del_plane;
  P.project(drift); P.musc_ext(drift / hair_radlen, drift);
  P.cut_box(kInf, -kInf, hscin_1x_left, hscin_1x_right, S1);
  P.musc(hscin_1x_thick / hscin_radlen);
  drift = hcer_zentrance - hscin_1x_zpos;
  P.project(drift); P.musc_ext(drift / hair_radlen, drift);
  P.musc(hcer_entr_thick / hcer_entr_radlen);
  drift = hcer_zmirror - hcer_zentrance;
  P.project(drift); P.musc_ext(drift / hcer_radlen, drift);
  P.musc(hcer_mir_thick / hcer_mir_radlen);
  drift = hcer_zexit - hcer_zmirror;
  P.project(drift); P.musc_ext(drift / hcer_radlen, drift);
  P.musc(hcer_exit_thick / hcer_exit_radlen);
  drift = hscin_2x_zpos - hcer_zexit;
  P.project(drift); P.musc_ext(drift / hair_radlen, drift);
  P.cut_box(kInf, -kInf, hscin_2x_left, hscin_2x_right, S2);
  P.musc(hscin_2x_thick / hscin_radlen);
  drift = hcal_4ta_zpos - hscin_2x_zpos;
  P.project(drift); P.musc_ext(drift / hair_radlen, drift);
  P.add(OP_RECON).a = right ? 0.48 : 0.78;
  P.add(OP_END);
}

}  // namespace

CompiledArm compile_arm(int arm_id, const ForwardMaps& fwd, const CosyTerms& rec) {
  CompiledArm A;
  const int want = (arm_id == 1 || arm_id == 3 || arm_id == 4) ? 12 : arm_id == 2 ? 10 : arm_id == 5 ? 32 : -1;
  if (want < 0) throw std::runtime_error("unknown spectrometer id");
  if ((int)fwd.cls.size() != want) {
    throw std::runtime_error(arm_id == 1 ? "MC_HMS, wrong number of transport classes"
                             : arm_id == 2 ? "MC_SOS, wrong number of transport classes"
                             : arm_id == 3 ? "MC_HRSR, wrong number of transport classes"
                             : arm_id == 4 ? "MC_HRSL, wrong number of transport classes"
                                           : "Bender-SHMS, wrong number of transport classes");
  }
  A.fwd = fwd; A.rec = rec; A.arm_id = arm_id;
  std::memset(&A.tab, 0, sizeof(A.tab));
  A.tab.n_classes = (int)fwd.cls.size();
  for (size_t k = 0; k < fwd.cls.size(); ++k) {
    PolyClass pc = compile_terms(fwd.cls[k], A.recs, &A.fwd_nonzero, kArmBlockThreads);
    pc.length_cm = k < fwd.length_cm.size() ? fwd.length_cm[k] : 0.0;
    pc.adrift = fwd.adrift[k];
    pc.driftdist_cm = fwd.driftdist_cm[k];
    A.tab.fwd[k] = pc;
    A.fwd_terms += fwd.cls[k].n();
  }
  if (rec.nout != 4) throw std::runtime_error("reconstruction map must have 4 outputs");
  A.tab.rec = compile_terms(rec, A.recs, nullptr, kArmBlockThreads);
  A.rec_terms = rec.n();
  Prog P;
  if (arm_id == 1) build_hms(P);
  else if (arm_id == 5) build_shms(P);
  else if (arm_id == 2) build_sos(P);
  else build_hrs(P, arm_id == 3);
  if ((int)P.ops.size() > kMaxArmOps) throw std::runtime_error("arm program too long");
  // Drift lengths taken from the maps (driftdist(spectr,k), transp.f:419) are known now:
  // fold them into plain drifts with the reference's arithmetic.
  for (size_t k = 0; k < P.ops.size(); ++k) {
    ArmOp& o = P.ops[k];
    if (o.op != OP_PROJECT_DD) continue;
    const double dd = fwd.driftdist_cm.at(o.i0 - 1);
    if (!fwd.adrift.at(o.i0 - 1))
      std::fprintf(stderr, "Transformation #%d is NOT a drift\n", o.i0);      // mc_hms.f:257
    if (o.i1) {                      // mc_hms.f:411-413 + mc_hms_hut.f:315
      const double zdrift = dd + o.a;
      const double zinit = -zdrift;
      o.a = o.b - zinit;
    } else {
      o.a = dd + o.a;
    }
    o.op = OP_PROJECT; o.b = 0; o.i0 = 0; o.i1 = 0;
    if (k + 1 < P.ops.size() && P.ops[k + 1].op == OP_MUSC_EXT && P.ops[k + 1].i0 == 1) {
      ArmOp& m = P.ops[k + 1];           // radw = drift/radlen, mc_hrsl_hut.f:197-200
      const double radw = o.a / m.a;
      m.a = radw; m.b = std::sqrt(radw); m.c = o.a; m.i0 = 0;
    }
  }
  A.ops = P.ops;
  A.tab.n_ops = (int)P.ops.size();
  // the loop compacts survivors after the entrance collimator: split behind its last cut
  A.tab.split_op = 0;
  for (int k = 0; k < (int)P.ops.size(); ++k)
    if (P.ops[k].op == OP_CUT_OCT) A.tab.split_op = k + 1;
  if (P.split >= 0) A.tab.split_op = P.split;
  A.tab.n_mid = 0;
  for (int m : P.mids)
    if (A.tab.n_mid < 3 && m > (A.tab.n_mid ? A.tab.mid_op[A.tab.n_mid - 1] : A.tab.split_op) && m < (int)P.ops.size())
      A.tab.mid_op[A.tab.n_mid++] = m;
  // Gaussian look-ahead (transport.cuh: GaussQueue): every op that draws gauss1(99.) pairs carries, in its unused
  // `code` word, how many Gaussians the program draws from this op up to the next point where the stream of
  // Gaussians is interrupted -- another kind of draw (the resmult uniform, the collimator stepping) or a
  // compaction point of the loop, where the kernel ends.  Low half: multiple-scattering draws (ms_flag), high half:
  // chamber-resolution draws (wcs_flag).
  {
    unsigned rem_ms = 0, rem_wcs = 0;
    for (int k = (int)A.ops.size() - 1; k >= 0; --k) {
      ArmOp& o = A.ops[k];
      if (o.op == OP_MUSC && o.a != 0.) rem_ms += 2;
      else if (o.op == OP_MUSC_EXT && o.a != 0.) rem_ms += 4;
      else if (o.op == OP_DC_PLANE) rem_wcs += 2;
      if (o.op == OP_MUSC || o.op == OP_MUSC_EXT || o.op == OP_DC_PLANE) {
        if (rem_ms > 0xffffu || rem_wcs > 0xffffu) throw std::runtime_error("arm program: too many Gaussian draws");
        o.code = (int32_t)(rem_ms | (rem_wcs << 16));
      }
      bool barrier = (o.op == OP_RESMULT_DRAW || o.op == OP_COLL || o.op == OP_COLL_DATA || k == A.tab.split_op ||
                      o.op == OP_TRANSP || o.op == OP_RECON);        // the maps use the queue's shared memory
      for (int m = 0; m < A.tab.n_mid; ++m) barrier = barrier || (k == A.tab.mid_op[m]);
      if (barrier) {
        // ops before k count up to here only; op k itself starts a new stretch (its own draws were added above)
        if (o.op == OP_MUSC || o.op == OP_MUSC_EXT || o.op == OP_DC_PLANE) { /* keeps its count */ }
        rem_ms = 0; rem_wcs = 0;
      }
    }
  }
  return A;
}

const char* stop_name(int arm_id, int code) {
  if (arm_id == 1 && code >= 0 && code < hmsc::N) return hmsc::names[code];
  if (arm_id == 5 && code >= 0 && code < shmsc::N) return shmsc::names[code];
  if (arm_id == 2 && code >= 0 && code < sosc::N) return sosc::names[code];
  if ((arm_id == 3 || arm_id == 4) && code >= 0 && code < hrsc::N) return hrsc::names[code];
  return "?";
}
int n_stop_codes(int arm_id) {
  return arm_id == 1 ? hmsc::N : arm_id == 5 ? shmsc::N : arm_id == 2 ? sosc::N : (arm_id == 3 || arm_id == 4) ? hrsc::N : 0;
}

}  // namespace simc
