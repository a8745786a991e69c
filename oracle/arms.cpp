// ORACLE -- TEST INFRASTRUCTURE ONLY (see arms.hpp).  Literal, sequential restatement;
// every block cites the reference lines it follows.
#include "arms.hpp"


// the reference's build flags (Makefile:63, -fdefault-real-8 without -fdefault-double-8) turn every
// `...d0` literal into a REAL(16) constant; hrsl/mc_hrsl.f:248-466 and hrsl/mc_hrsl_hut.f:293,356 (same in hrsr)
// pass such literals by reference to REAL*8 dummies, which then read the low 8 bytes of the 16-byte value:
// 62.75333333d0 -> -1.3e-41, 121.77333333d0 -> -3.1e+232, 45.0d0 -> 0.0, ...  `as_written` (default, what the
// product implements: the intended lengths and the 45 degree chamber tilt) or `as_built` (that reinterpretation,
// reproduced from a libquadmath probe) is chosen with oracle_set_hrs_literals(); tests/test_oracle_hrs_literals.py shows
// what changes (path length, hence the kaon survival probability; the un-rotated VDC frame).
#include <cstdint>
#include <cstring>
static bool g_hrs_as_built = false;
extern "C" void oracle_set_hrs_literals(int as_built) { g_hrs_as_built = as_built != 0; }
static double hrs_lit(double as_written, const char* decimal) {
  if (!g_hrs_as_built) return as_written;
  // low 8 bytes of the binary128 value of each literal, printed by tools/hrs_literal_probe.cpp (libquadmath)
  static const struct { const char* lit; uint64_t lo; } kTable[] = {
      {"62.75333333", 0xb7729d815ab37fcaULL},    // -1.3355790489890208e-41
      {"31.37666667", 0xa595a644f8ad7b4eULL},    // -1.2493079844832989e-127
      {"121.77333333", 0xf03430085b6e3ac6ULL},   // -3.1341656965427691e+232
      {"60.88666667", 0x6745b46a2a6b3888ULL},    //  3.0220523071482735e+189
      {"659.73445725", 0xb139c94f69ca9ef5ULL},   // -1.4594567076032913e-71
      {"121.7866667", 0xb27100f41bc62f24ULL},    // -1.0091251298040999e-65
      {"60.89333333", 0xd62aef6cdfd2381cULL},    // -1.235519403111592e+107
      {"45.0", 0x0000000000000000ULL},           //  0
  };
  for (const auto& e : kTable)
    if (std::strcmp(e.lit, decimal) == 0) {
      double d;
      std::memcpy(&d, &e.lo, sizeof(d));
      return d;
    }
  return as_written;
}

namespace simc_oracle {

// =====================================================================================
// HMS
// =====================================================================================
namespace {
namespace hms {
// hms/apertures_hms.inc
constexpr double r_Q1 = 20.50, r_Q2 = 30.22, r_Q3 = 30.22;
constexpr double x_d1 = 34.29, y_d1 = 12.07, x_d2 = 27.94, y_d2 = 18.42, x_d3 = 13.97, y_d3 = 18.95;
constexpr double x_d4 = 1.956, y_d4 = 20.32, x_d5 = 27.94, y_d5 = 12.065, r_d5 = 6.35, a_d6 = -0.114, b_d6 = 20.54;
// hms/mc_hms.f:28-95
constexpr double x_offset_pipes = 2.8, y_offset_pipes = 0.0;
constexpr double h_entr = 4.575, v_entr = 11.646, h_exit = 4.759, v_exit = 12.114;
constexpr double x_off = +0.000, y_off = +0.028, z_off = +40.17;
constexpr double z_entr = 126.2e0 + z_off, z_exit = z_entr + 6.3e0;
constexpr double z_dip1 = 64.77e0, z_dip2 = z_dip1 + 297.18e0, z_dip3 = z_dip2 + 115.57e0;

// hms/mc_hms.f:445-492
bool hit_dipole(double x, double y) {
  const double x_local = std::fabs(x), y_local = std::fabs(y);
  const bool check1 = (x_local <= x_d1) && (y_local <= y_d1);
  const bool check2 = (x_local <= x_d2) && (y_local <= y_d2);
  const bool check3 = (x_local <= x_d3) && (y_local <= y_d3);
  const bool check4 = (x_local <= x_d4) && (y_local <= y_d4);
  const bool check5 = ((x_local - x_d5) * (x_local - x_d5) + (y_local - y_d5) * (y_local - y_d5)) <= r_d5 * r_d5;
  const bool check6 = (x_local >= x_d4) && (x_local <= x_d3) && ((y_local - a_d6 * x_local - b_d6) <= 0.0);
  return !(check1 || check2 || check3 || check4 || check5 || check6);
}

// hms/mc_hms_hut.f:27-255 (geometry parameters)
constexpr double hfoil_exit_radlen = 8.90, hfoil_exit_thick = 0.011 * 2.54;
constexpr double hair_radlen = 30420.;
constexpr double hdc_entr_radlen = 28.7, hdc_entr_thick = 0.001 * 2.54;
constexpr double hdc_radlen = 16700.0, hdc_thick = 1.8;
constexpr double hdc_wire_radlen = 0.35, hdc_wire_thick = 0.0000049;
constexpr double hdc_cath_radlen = 7.2, hdc_cath_thick = 0.000177;
constexpr double hdc_exit_radlen = 28.7, hdc_exit_thick = 0.001 * 2.54;
constexpr double haer_entr_radlen = 8.90, haer_entr_thick = 0.15;
constexpr double haer_radlen = 150.0, haer_thick = 9.0;
constexpr double haer_air_radlen = 30420.0, haer_air_thick = 16.0;
constexpr double haer_exit_radlen = 8.90, haer_exit_thick = 0.1;
constexpr double hscin_radlen = 42.4;
constexpr double hcer_entr_radlen = 8.90, hcer_entr_thick = 0.040 * 2.54;
constexpr double hcer_radlen = 9620.0;
constexpr double hcer_mir_radlen = 400.0, hcer_mir_thick = 2.0;
constexpr double hcer_exit_radlen = 8.90, hcer_exit_thick = 0.040 * 2.54;
constexpr double hdc_sigma = 0.030;
constexpr int hdc_nr_cham = 2, hdc_nr_plan = 6;
constexpr double hdc_1_zpos = -52.1084, hdc_2_zpos = 29.2608;
constexpr double hdc_del_plane = hdc_thick + hdc_wire_thick + hdc_cath_thick;
constexpr double hdc_1_left = 26.0, hdc_1_right = -26.0, hdc_1y_offset = 1.443, hdc_1_top = -56.5, hdc_1_bot = 56.5,
                 hdc_1x_offset = 1.670;
constexpr double hdc_2_left = 26.0, hdc_2_right = -26.0, hdc_2y_offset = 2.753, hdc_2_top = -56.5, hdc_2_bot = 56.5,
                 hdc_2x_offset = 2.758;
constexpr double haer_zentrance = 35.699, haer_zexit = 60.949;
constexpr double hscin_1x_zpos = 77.830, hscin_1y_zpos = 97.520, hscin_2x_zpos = 298.820, hscin_2y_zpos = 318.510;
constexpr double hscin_1x_thick = 1.067, hscin_1y_thick = 1.067, hscin_2x_thick = 1.067, hscin_2y_thick = 1.067;
constexpr double hscin_1x_left = 37.75, hscin_1x_right = -37.75, hscin_1x_offset = -1.3;
constexpr double hscin_1y_top = -60.25, hscin_1y_bot = 60.25, hscin_1y_offset = -1.3;
constexpr double hcer_zentrance = 110.000, hcer_zmirror = 230.000, hcer_zexit = 265.000;
constexpr double hcal_4ta_zpos = 371.69;
constexpr int scintrig = 3;

// mc_hms_hut, hms/mc_hms_hut.f:1-608
bool hut(Track& t, ArmCall& a, double& m2, double& p, bool& dflag, double zinit) {
  Rng& r = *t.rng;
  const bool ms = a.ms_flag, wcs = a.wcs_flag, dec = a.decay_flag;
  double radw, drift;
  float xdc[12], ydc[12], zdc[12];
  // :292-298  15% of events get doubled DC resolution
  const double tmpran = r.grnd();
  a.resmult = (tmpran < 0.15) ? 2.0 : 1.0;
  for (int i = 0; i < 12; ++i) { xdc[i] = 0.f; ydc[i] = 0.f; }
  int scincount = 0;

  // :315-317 drift to a point 25 cm in front of DC1, exit foil
  drift = (hdc_1_zpos - 25.000) - zinit;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  radw = hfoil_exit_thick / hfoil_exit_radlen;
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
  // :321-326 air to DC1
  drift = (hdc_1_zpos - 0.5 * hdc_nr_plan * hdc_del_plane) - (hdc_1_zpos - 25.000);
  radw = drift / hair_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);

  for (int jchamber = 1; jchamber <= 2; ++jchamber) {
    // :331-336 / :385-390 entrance window
    radw = hdc_entr_thick / hdc_entr_radlen;
    if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
    const int npl_off = (jchamber - 1) * hdc_nr_plan;
    for (int iplane = 1; iplane <= hdc_nr_plan; ++iplane) {   // :341-370 / :395-425
      radw = hdc_cath_thick / hdc_cath_radlen;
      if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
      drift = 0.5 * hdc_thick;
      radw = drift / hdc_radlen;
      if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
      drift = drift + hdc_cath_thick;
      project(t, drift, dec, dflag, m2, p, a.pathlen);
      radw = hdc_wire_thick / hdc_wire_radlen;
      if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
      double tmpran1 = 0., tmpran2 = 0.;
      if (wcs) { tmpran1 = gauss1(r, 99.0); tmpran2 = gauss1(r, 99.0); }
      xdc[npl_off + iplane - 1] = (float)(t.xs + hdc_sigma * tmpran1 * a.resmult);
      ydc[npl_off + iplane - 1] = (float)(t.ys + hdc_sigma * tmpran2 * a.resmult);
      if (iplane == 2 || iplane == 5) xdc[npl_off + iplane - 1] = 0.f;
      else ydc[npl_off + iplane - 1] = 0.f;
      drift = 0.5 * hdc_thick;
      radw = drift / hdc_radlen;
      if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
      drift = drift + hdc_wire_thick;
      project(t, drift, dec, dflag, m2, p, a.pathlen);
    }
    radw = hdc_exit_thick / hdc_exit_radlen;
    if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
    if (jchamber == 1) {
      if (t.xs > (hdc_1_bot - hdc_1x_offset) || t.xs < (hdc_1_top - hdc_1x_offset) ||
          t.ys > (hdc_1_left - hdc_1y_offset) || t.ys < (hdc_1_right - hdc_1y_offset)) {
        a.stop_code = hms_stop::DC1;
        return false;
      }
      radw = hdc_cath_thick / hdc_cath_radlen;
      if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
      // :379-384 air to DC2
      drift = (hdc_2_zpos - 0.5 * hdc_nr_plan * hdc_del_plane) - (hdc_1_zpos + 0.5 * hdc_nr_plan * hdc_del_plane);
      radw = drift / hair_radlen;
      project(t, drift, dec, dflag, m2, p, a.pathlen);
      if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
    } else {
      if (t.xs > (hdc_2_bot - hdc_2x_offset) || t.xs < (hdc_2_top - hdc_2x_offset) ||
          t.ys > (hdc_2_left - hdc_2y_offset) || t.ys < (hdc_2_right - hdc_2y_offset)) {
        a.stop_code = hms_stop::DC2;
        return false;
      }
      radw = hdc_cath_thick / hdc_cath_radlen;
      if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
    }
  }
  // :438-455 fit the track through REAL*4 arrays
  for (int jchamber = 1; jchamber <= hdc_nr_cham; ++jchamber) {
    const int npl_off = (jchamber - 1) * hdc_nr_plan;
    for (int iplane = 1; iplane <= hdc_nr_plan; ++iplane) {
      const double z0 = (jchamber == 1) ? hdc_1_zpos : hdc_2_zpos;
      zdc[npl_off + iplane - 1] = (float)(z0 + (iplane - 0.5 - 0.5 * hdc_nr_plan) * hdc_del_plane);
    }
  }
  float dxfp4, xfp4, dyfp4, yfp4;
  lfit(zdc, xdc, 12, dxfp4, xfp4);
  lfit(zdc, ydc, 12, dyfp4, yfp4);
  a.x_fp = (double)xfp4;
  a.y_fp = (double)yfp4;
  a.dx_fp = (double)dxfp4;
  a.dy_fp = (double)dyfp4;

  // :460-483 aerogel
  drift = haer_zentrance - hdc_2_zpos - 0.5 * hdc_nr_plan * hdc_del_plane;
  radw = drift / hair_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  radw = haer_entr_thick / haer_entr_radlen;
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
  drift = haer_thick;
  radw = drift / haer_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  drift = haer_air_thick;
  radw = drift / haer_air_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  radw = haer_exit_thick / haer_exit_radlen;
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);

  // All four scintillator planes are tested against the 1x/1y dimensions (A.7, :537-551).
  auto in_scin = [&]() {
    return t.ys < (hscin_1x_left + hscin_1y_offset) && t.ys > (hscin_1x_right + hscin_1y_offset) &&
           t.xs < (hscin_1y_bot + hscin_1x_offset) && t.xs > (hscin_1y_top + hscin_1x_offset);
  };
  // :487-498 S1X
  drift = hscin_1x_zpos - haer_zexit;
  radw = drift / hair_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  if (in_scin()) scincount++;
  radw = hscin_1x_thick / hscin_radlen;
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
  // :500-510 S1Y
  drift = hscin_1y_zpos - hscin_1x_zpos;
  radw = drift / hair_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  if (in_scin()) scincount++;
  radw = hscin_1y_thick / hscin_radlen;
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
  // :514-535 Cherenkov
  drift = hcer_zentrance - hscin_1y_zpos;
  radw = drift / hair_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  radw = hcer_entr_thick / hcer_entr_radlen;
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
  drift = hcer_zmirror - hcer_zentrance;
  radw = drift / hcer_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  radw = hcer_mir_thick / hcer_mir_radlen;
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
  drift = hcer_zexit - hcer_zmirror;
  radw = hcer_exit_thick / hcer_exit_radlen;
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  radw = hcer_exit_thick / hcer_exit_radlen;
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
  // :537-551 S2X, S2Y
  drift = hscin_2x_zpos - hcer_zexit;
  radw = drift / hair_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  if (in_scin()) scincount++;
  radw = hscin_2x_thick / hscin_radlen;
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
  drift = hscin_2y_zpos - hscin_2x_zpos;
  radw = drift / hair_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  if (in_scin()) scincount++;
  radw = hscin_2y_thick / hscin_radlen;
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
  if (scincount < scintrig) {
    a.stop_code = hms_stop::SCIN;
    return false;
  }
  // :580-590 drift to the calorimeter (no cut applied)
  drift = hcal_4ta_zpos - hscin_2y_zpos;
  radw = drift / hair_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  return true;
}
}  // namespace hms
}  // namespace

// mc_hms_coll (hms/mc_hms_coll.f:1-147) / mc_shms_coll (shms/mc_shms_coll.f): a pion or muon is stepped through
// the 6.3 cm tungsten collimator in 20 slices; inside the material it may be absorbed (pions, pion_coll_absorb),
// scatters (musc) and loses energy (enerloss_new with a sampled fluctuation), and decays in flight (project).
namespace {
struct CollGeom { double h_entr, v_entr, h_exit, v_exit, x_off, y_off, thick, radl; };

// hms/pion_coll_absorb.f:1-95
double pion_coll_absorb(double ppi, double thick) {
  static const double T[14] = {85.0, 125.0, 165.0, 205.0, 245.0, 315.0, 584.02, 711.95, 870.12, 1227.57, 1446.58, 1865.29,
                               2858.0, 4159.0};
  static const double sigreac[14] = {26.03, 84.47, 117.3, 117.4, 101.9, 69.58, 42.5, 44.7, 47.9, 46.5, 45.2, 39.6, 35.34, 33.15};
  static const double qreac[14] = {0.948, 0.659, 0.56, 0.5342, 0.5452, 0.60796, 0.699, 0.689, 0.679, 0.683, 0.688, 0.704,
                                   0.7483, 0.7705};
  const double Navagadro = 6.0221367e+23, mpi = 139.56995, mate_dens = 17.0, mate_A = 171.57;
  const double Epi = std::sqrt(ppi * ppi + mpi * mpi);
  const double Tpi = Epi - mpi;
  double sigA = 0.0;
  for (int i = 1; i <= 13; ++i) {
    if ((Tpi > T[i - 1]) && (Tpi <= T[i])) {
      const double Thi = T[i], Tlo = T[i - 1];
      const double sigAhi = sigreac[i] * std::pow(mate_A, qreac[i]);
      const double sigAlo = sigreac[i - 1] * std::pow(mate_A, qreac[i - 1]);
      sigA = (sigAlo * (Thi - Tpi) + sigAhi * (Tpi - Tlo)) / (Thi - Tlo);
      sigA = sigA * 1.e-27;
    }
  }
  if (Tpi > T[13]) {
    sigA = sigreac[13] * std::pow(mate_A, qreac[13]);
    sigA = sigA * 1.e-27;
  }
  const double lambdai = mate_dens * Navagadro * sigA / mate_A;
  return std::exp(-(0.0 + thick * lambdai));
}

// enerloss_new.f:1-85 with typeflag = 1 (same arithmetic as oracle/target.cpp, on the arm's generator)
double coll_enerloss(Track& t, double len, double dens, double zeff, double aeff, double epart, double mpart) {
  const double me = 0.51099906;
  const double thick = len * dens;
  const double gamma = epart / mpart;
  const double beta = std::sqrt(1. - 1. / (gamma * gamma));
  const double I = zeff == 1 ? 21.8e-06 : (16. * std::pow(zeff, 0.9)) * 1.0e-06;
  const double hnup = 28.816e-06 * std::sqrt(dens * zeff / aeff);
  const double log10bg = std::log(beta * gamma) / std::log(10.);
  const double CO = std::log(hnup) - std::log(I) + 0.5;
  double denscorr;
  if (log10bg < 0.) denscorr = 0.;
  else if (log10bg < 3.) denscorr = CO + std::log(10.) * log10bg + std::fabs(CO / 27.) * powi(3. - log10bg, 3);
  else if (log10bg < 4.7) denscorr = CO + std::log(10.) * log10bg;
  else denscorr = CO + std::log(10.) * 4.7;
  double Eloss;
  if (thick <= 0.) {
    Eloss = 0.;
  } else {
    const double Eloss_mp_new = 0.1536e-03 * zeff / aeff * thick / (beta * beta) *
                                (std::log(me / (I * I)) + 1.063 + 2. * std::log(gamma * beta) +
                                 std::log(0.1536 * zeff / aeff * thick / (beta * beta)) - beta * beta - denscorr);
    const double Eloss_mp = Eloss_mp_new * 1000.;
    const double chsi = 0.307075 / 2. * zeff / aeff * thick / (beta * beta);
    const double x = std::fabs(gauss1(*t.rng, 10.0));
    const double lambda = x > 0.0 ? -2.0 * std::log(x) : 100000.;
    Eloss = lambda * chsi + Eloss_mp;
  }
  if (Eloss > (epart - mpart)) Eloss = (epart - mpart) - 0.0000001;
  return Eloss;
}

bool mc_coll(Track& t, const CollGeom& G, ArmCall& a, double& m2, double& p, bool decay_flag, bool& dflag) {
  const int nstep = 20;
  const double coll_dens = 17.0, zcoll = 69.45, acoll = 171.56797;
  const double step_size = G.thick / nstep;
  double h_step = G.h_entr, v_step = G.v_entr;
  double epart = std::sqrt(p * p + m2);
  double thick_temp = 0.0;
  for (int n = 1; n <= nstep; ++n) {
    bool step_flag = false;
    if (std::fabs(t.ys - G.y_off) > h_step) { step_flag = true; a.coll_steps[0]++; thick_temp = step_size; }
    if (std::fabs(t.xs - G.x_off) > v_step) { step_flag = true; a.coll_steps[1]++; thick_temp = step_size; }
    if (std::fabs(t.xs - G.x_off) > (-v_step / h_step * std::fabs(t.ys - G.y_off) + 3 * v_step / 2)) {
      step_flag = true; a.coll_steps[2]++; thick_temp = step_size;
    }
    const double thick = thick_temp;
    if (step_flag) {
      if (m2 > 12000. && m2 < 20000.) {            // pions only: hadronic interaction
        const double trans = pion_coll_absorb(p, thick);
        const double rantemp = t.rng->grnd();
        if (rantemp > trans) return false;
      }
      const double coll_radw = thick / G.radl;
      musc(*t.rng, m2, p, coll_radw, t.dydzs, t.dxdzs);
      epart = std::sqrt(p * p + m2);
      const double eloss = coll_enerloss(t, thick, coll_dens, zcoll, acoll, epart, std::sqrt(m2));
      epart = epart - eloss;
      if (epart < std::sqrt(m2)) return false;
      p = std::sqrt(epart * epart - m2);
      t.dpps = 100. * (p / a.p_spec - 1.);
    }
    project(t, step_size, decay_flag, dflag, m2, p, a.pathlen);
    h_step = h_step + (G.h_exit - G.h_entr) / nstep;
    v_step = v_step + (G.v_exit - G.v_entr) / nstep;
  }
  return true;
}
}  // namespace

// mc_hms, hms/mc_hms.f:1-441
void mc_hms(Track& t, const ArmOptics& o, ArmCall& a) {
  using namespace hms;
  const int spectr_classes = o.fwd.n_classes();
  if (spectr_classes != 12) throw std::runtime_error("MC_HMS, wrong number of transport classes");
  const bool dec = a.decay_flag;
  a.ok_spec = false;
  a.stop_code = 0;
  a.reached_hut = false;
  bool dflag = false;
  t.xs = a.x; t.ys = a.y; t.zs = a.z; t.dxdzs = a.dxdz; t.dydzs = a.dydz;
  t.dpps = a.dpp;
  double p = a.p_spec * (1. + t.dpps / 100.);
  double& m2 = a.m2;
  double xt, yt, zdrift;
  auto stop = [&](int code) { a.stop_code = code; };

  // :201-254 collimator
  zdrift = z_entr;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if (a.using_coll && (m2 > 100.0 * 100.0) && (m2 < 200.0 * 200.0)) {
    const CollGeom G{4.575, 11.646, 4.759, 12.114, 0.000, +0.028, 6.30, 0.41753};      // hms/mc_hms_coll.f:17-34
    if (!mc_coll(t, G, a, m2, p, dec, dflag)) return stop(hms_stop::COLL);
  } else {
    if (std::fabs(t.ys - y_off) > h_entr) return stop(hms_stop::SLIT_HOR);
    if (std::fabs(t.xs - x_off) > v_entr) return stop(hms_stop::SLIT_VERT);
    if (std::fabs(t.xs - x_off) > (-v_entr / h_entr * std::fabs(t.ys - y_off) + 3 * v_entr / 2))
      return stop(hms_stop::SLIT_OCT);
    zdrift = z_exit - z_entr;
    project(t, zdrift, dec, dflag, m2, p, a.pathlen);
    if (std::fabs(t.ys - y_off) > h_exit) return stop(hms_stop::SLIT_HOR);
    if (std::fabs(t.xs - x_off) > v_exit) return stop(hms_stop::SLIT_VERT);
    if (std::fabs(t.xs - x_off) > (-v_exit / h_exit * std::fabs(t.ys - y_off) + 3 * v_exit / 2))
      return stop(hms_stop::SLIT_OCT);
  }
  // :258-280 Q1
  zdrift = o.fwd.cls[0].driftdist - z_exit;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if ((t.xs * t.xs + t.ys * t.ys) > r_Q1 * r_Q1) return stop(hms_stop::Q1_IN);
  transp(t, o.fwd, 2, dec, dflag, m2, p, 125.233e0, a.pathlen);
  if ((t.xs * t.xs + t.ys * t.ys) > r_Q1 * r_Q1) return stop(hms_stop::Q1_MID);
  transp(t, o.fwd, 3, dec, dflag, m2, p, 62.617e0, a.pathlen);
  if ((t.xs * t.xs + t.ys * t.ys) > r_Q1 * r_Q1) return stop(hms_stop::Q1_OUT);
  // :284-306 Q2
  zdrift = o.fwd.cls[3].driftdist;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if ((t.xs * t.xs + t.ys * t.ys) > r_Q2 * r_Q2) return stop(hms_stop::Q2_IN);
  transp(t, o.fwd, 5, dec, dflag, m2, p, 143.90e0, a.pathlen);
  if ((t.xs * t.xs + t.ys * t.ys) > r_Q2 * r_Q2) return stop(hms_stop::Q2_MID);
  transp(t, o.fwd, 6, dec, dflag, m2, p, 71.95e0, a.pathlen);
  if ((t.xs * t.xs + t.ys * t.ys) > r_Q2 * r_Q2) return stop(hms_stop::Q2_OUT);
  // :310-332 Q3
  zdrift = o.fwd.cls[6].driftdist;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if ((t.xs * t.xs + t.ys * t.ys) > r_Q3 * r_Q3) return stop(hms_stop::Q3_IN);
  transp(t, o.fwd, 8, dec, dflag, m2, p, 143.8e0, a.pathlen);
  if ((t.xs * t.xs + t.ys * t.ys) > r_Q3 * r_Q3) return stop(hms_stop::Q3_MID);
  transp(t, o.fwd, 9, dec, dflag, m2, p, 71.9e0, a.pathlen);
  if ((t.xs * t.xs + t.ys * t.ys) > r_Q3 * r_Q3) return stop(hms_stop::Q3_OUT);
  // :337-396 dipole
  zdrift = o.fwd.cls[9].driftdist;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  xt = t.xs; yt = t.ys;
  rotate_haxis(t, -6.0e0, xt, yt);
  if (hit_dipole(xt, yt)) return stop(hms_stop::D1_IN);
  transp(t, o.fwd, 11, dec, dflag, m2, p, 526.053e0, a.pathlen);
  xt = t.xs; yt = t.ys;
  rotate_haxis(t, 6.0e0, xt, yt);
  if (hit_dipole(xt, yt)) return stop(hms_stop::D1_OUT);
  if ((((xt - x_offset_pipes) * (xt - x_offset_pipes) + (yt - y_offset_pipes) * (yt - y_offset_pipes)) >
       30.48 * 30.48) ||
      (std::fabs((yt - y_offset_pipes)) > 20.5232))
    return stop(hms_stop::D1_OUT);
  zdrift = z_dip1;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if (((t.xs - x_offset_pipes) * (t.xs - x_offset_pipes) + (t.ys - y_offset_pipes) * (t.ys - y_offset_pipes)) >
      1145.518)
    return stop(hms_stop::D1_OUT);
  zdrift = z_dip2 - z_dip1;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if (((t.xs - x_offset_pipes) * (t.xs - x_offset_pipes) + (t.ys - y_offset_pipes) * (t.ys - y_offset_pipes)) >
      1512.2299)
    return stop(hms_stop::D1_OUT);
  zdrift = z_dip3 - z_dip2;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if (((t.xs - x_offset_pipes) * (t.xs - x_offset_pipes) + (t.ys - y_offset_pipes) * (t.ys - y_offset_pipes)) >
      2162.9383)
    return stop(hms_stop::D1_OUT);

  // :400-417 hut
  a.reached_hut = true;
  zdrift = o.fwd.cls[11].driftdist - z_dip3;
  if (!hut(t, a, m2, p, dflag, -zdrift)) return;
  // :419-437 reconstruct from the fitted focal-plane track
  t.xs = a.x_fp; t.ys = a.y_fp; t.dxdzs = a.dx_fp; t.dydzs = a.dy_fp;
  double dpp_recon, dth_recon, dph_recon, y_recon;
  if (t.calls) t.calls[47]++;
  o.rec.eval(t, a.fry, dpp_recon, dth_recon, dph_recon, y_recon);
  a.dpp = dpp_recon;
  a.dxdz = dph_recon;
  a.dydz = dth_recon;
  a.y = y_recon;
  a.ok_spec = true;
}

// =====================================================================================
// SHMS
// =====================================================================================
namespace {
namespace shms {
// shms/apertures_shms.inc
constexpr double r_HBx = 11.2, r_HBfym = -4.13, r_HBfyp = 11.75, r_HBmenym = -5.45, r_HBmenyp = 11.74;
constexpr double r_HBmexym = -10.25, r_HBmexyp = 11.71, r_HBbym = -11.71, r_HBbyp = 11.70;
constexpr double r_Q1 = 20.00, r_Q2 = 30.00, r_Q3 = 30.00, r_D1 = 30.00;
// shms/mc_shms.f:233-318
constexpr double h_entr = 8.5, v_entr = 12.5, h_exit = 8.65, v_exit = 12.85, x_off = +0.00, y_off = +0.00;
constexpr double zd_hbin = 118.39, zd_hbmen = 17.61, zd_hbmex = 80.0, zd_hbout = 17.61;
constexpr double z_entr = 25.189, z_thick = 6.35;
constexpr double zd_q1in = 58.39, zd_q1men = 28.35, zd_q1mid = 93.65, zd_q1mex = 93.65, zd_q1out = 28.35;
constexpr double zd_q2in = 25.55, zd_q2men = 39.1, zd_q2mid = 79.35, zd_q2mex = 79.35, zd_q2out = 39.1;
constexpr double zd_q3in = 28.10, zd_q3men = 39.1, zd_q3mid = 79.35, zd_q3mex = 79.35, zd_q3out = 39.1;
constexpr double zd_q3d1trans = 18.00, zd_d1flare = 30.10, zd_d1men = 39.47;
constexpr double zd_d1mid = 36.406263, zd_d1mex = 36.406263, zd_d1out = 60.68, zd_fp = 307.95;
// shms/hut.inc
constexpr double hfoil_exit_radlen = 8.89, hfoil_exit_thick = 0.020 * 2.54;
constexpr double hair_radlen = 30420.;
constexpr double hdc_entr_radlen = 28.7, hdc_entr_thick = 0.001 * 2.54;
constexpr double hdc_radlen = 16700.0, hdc_thick = 0.125 * 2.54;
constexpr double hdc_wire_radlen = 0.35, hdc_wire_thick = 0.0000354;
constexpr double hdc_cath_radlen = 28.6, hdc_cath_thick = 0.001 * 2.54;
constexpr double hdc_exit_radlen = 28.7, hdc_exit_thick = 0.001 * 2.54;
constexpr double hscin_radlen = 42.4;
constexpr double hcer_entr_radlen = 19.63, hcer_entr_thick = 0.002 * 2.54;
constexpr double hcer_1_radlen = 11700.0;
constexpr double hcer_mirglass_radlen = 12.29, hcer_mirglass_thick = 0.3;
constexpr double hcer_exit_radlen = 19.63, hcer_exit_thick = 0.002 * 2.54;
constexpr double hcer_2_entr_radlen = 8.90, hcer_2_entr_thick = 0.040 * 2.54;
constexpr double hcer_2_radlen = 1202.5;
constexpr double hcer_mir_radlen = 400., hcer_mir_thick = 2.00;
constexpr double hcer_2_exit_radlen = 8.90, hcer_2_exit_thick = 0.040 * 2.54;
constexpr double hdc_sigma = 0.020;
constexpr int hdc_nr_cham = 2, hdc_nr_plan = 6;
constexpr double hdc_1_zpos = -40.656, hdc_2_zpos = 39.332;
constexpr double hdc_1_left = 40.0, hdc_1_right = -40.0, hdc_1y_offset = 0.0, hdc_1_top = -40., hdc_1_bot = 40.,
                 hdc_1x_offset = 0.0;
constexpr double hdc_2_left = 40.0, hdc_2_right = -40.0, hdc_2y_offset = 0.0, hdc_2_top = -40., hdc_2_bot = 40.,
                 hdc_2x_offset = 0.;
constexpr double hscin_1x_zpos = 52.1, hscin_1y_zpos = 61.7, hscin_2x_zpos = 271.4, hscin_2y_zpos = 282.4;
constexpr double hscin_1x_thick = 1.000 * 1.067, hscin_1y_thick = 1.000 * 1.067, hscin_2x_thick = 1.000 * 1.067,
                 hscin_2y_thick = 1.000 * 1.067;
constexpr double hscin_1x_left = 50., hscin_1x_right = -50., hscin_1x_offset = 0.0;
constexpr double hscin_1y_top = -45., hscin_1y_bot = 45., hscin_1y_offset = 0.0;
constexpr double hscin_2x_left = 55., hscin_2x_right = -55., hscin_2x_offset = 0.;
constexpr double hscin_2y_top = -62.5, hscin_2y_bot = 62.5, hscin_2y_offset = 0;
constexpr double hcer_1_zentrance = -291.700, hcer_1_zmirror = -84.900, hcer_1_zexit = -61.700;
constexpr double hcer_2_zentrance = 72.600, hcer_2_zmirror = 179.400, hcer_2_zexit = 202.600;
constexpr double hcal_4ta_zpos = 341.0;
constexpr double hcal_left = 63.00, hcal_right = -63.00, hcal_top = -70.00, hcal_bottom = 70.00;

// mc_shms_hut, shms/mc_shms_hut.f:1-458 with cer_flag=.true., vac_flag=.false.
// (hard-wired in shms/mc_shms.f:352-353)
bool hut(Track& t, ArmCall& a, double& m2, double& p, bool& dflag) {
  Rng& r = *t.rng;
  const bool ms = a.ms_flag, wcs = a.wcs_flag, dec = a.decay_flag;
  double radw, drift;
  float xdc[12], ydc[12], zdc[12];
  const double hdc_del_plane = hdc_thick + hdc_wire_thick + hdc_cath_thick;   // run-time sum, :73
  a.resmult = 1.0;
  for (int i = 0; i < 12; ++i) { xdc[i] = 0.f; ydc[i] = 0.f; }

  // :89-118 noble-gas Cherenkov
  radw = hfoil_exit_thick / hfoil_exit_radlen;
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
  radw = hcer_entr_thick / hcer_entr_radlen;
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
  drift = hcer_1_zmirror - hcer_1_zentrance - hcer_mirglass_thick / 2;
  radw = drift / hcer_1_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  radw = hcer_mirglass_thick / hcer_mirglass_radlen;
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
  drift = hcer_1_zexit - hcer_1_zmirror - hcer_mirglass_thick / 2;
  radw = drift / hcer_1_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  radw = hcer_exit_thick / hcer_exit_radlen;
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);

  // :150-156 air to DC1
  drift = (hdc_1_zpos - 0.5 * hdc_nr_plan * hdc_del_plane) - hcer_1_zexit;
  radw = drift / hair_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);

  for (int jchamber = 1; jchamber <= 2; ++jchamber) {
    radw = hdc_entr_thick / hdc_entr_radlen;
    if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
    const int npl_off = (jchamber - 1) * hdc_nr_plan;
    for (int iplane = 1; iplane <= hdc_nr_plan; ++iplane) {   // :163-201 / :231-266
      radw = hdc_cath_thick / hdc_cath_radlen;
      if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
      drift = 0.5 * hdc_thick;
      radw = drift / hdc_radlen;
      if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
      drift = drift + hdc_cath_thick;
      project(t, drift, dec, dflag, m2, p, a.pathlen);
      radw = hdc_wire_thick / hdc_wire_radlen;
      if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
      double tmpran1 = 0., tmpran2 = 0.;
      if (wcs) { tmpran1 = gauss1(r, 99.0); tmpran2 = gauss1(r, 99.0); }
      xdc[npl_off + iplane - 1] = (float)(t.xs + hdc_sigma * tmpran1 * a.resmult);
      ydc[npl_off + iplane - 1] = (float)(t.ys + hdc_sigma * tmpran2 * a.resmult);
      if (iplane == 2 || iplane == 5) xdc[npl_off + iplane - 1] = 0.f;
      else ydc[npl_off + iplane - 1] = 0.f;
      drift = 0.5 * hdc_thick;
      radw = drift / hdc_radlen;
      if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
      drift = drift + hdc_wire_thick;
      project(t, drift, dec, dflag, m2, p, a.pathlen);
    }
    radw = hdc_exit_thick / hdc_exit_radlen;
    if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
    if (jchamber == 1) {
      if (t.xs > (hdc_1_bot - hdc_1x_offset) || t.xs < (hdc_1_top - hdc_1x_offset) ||
          t.ys > (hdc_1_left - hdc_1y_offset) || t.ys < (hdc_1_right - hdc_1y_offset)) {
        a.stop_code = shms_stop::DC1;
        return false;
      }
      radw = hdc_cath_thick / hdc_cath_radlen;
      if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
      // :222-226 air to DC2
      drift = hdc_2_zpos - hdc_1_zpos - hdc_nr_plan * hdc_del_plane;
      project(t, drift, dec, dflag, m2, p, a.pathlen);
      radw = drift / hair_radlen;
      if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
    } else {
      if (t.xs > (hdc_2_bot - hdc_2x_offset) || t.xs < (hdc_2_top - hdc_2x_offset) ||
          t.ys > (hdc_2_left - hdc_2y_offset) || t.ys < (hdc_2_right - hdc_2y_offset)) {
        a.stop_code = shms_stop::DC2;
        return false;
      }
      radw = hdc_cath_thick / hdc_cath_radlen;
      if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
    }
  }
  // :290-318 S1X, S1Y
  drift = hscin_1x_zpos - hdc_2_zpos - 0.5 * hdc_nr_plan * hdc_del_plane;
  radw = drift / hair_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  if (t.ys > (hscin_1x_left + hscin_1y_offset) || t.ys < (hscin_1x_right + hscin_1y_offset)) {
    a.stop_code = shms_stop::S1X;
    return false;
  }
  radw = hscin_1x_thick / hscin_radlen;
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
  drift = hscin_1y_zpos - hscin_1x_zpos;
  radw = drift / hair_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  if (t.xs > (hscin_1y_bot + hscin_1x_offset) || t.xs < (hscin_1y_top + hscin_1x_offset)) {
    a.stop_code = shms_stop::S1Y;
    return false;
  }
  radw = hscin_1y_thick / hscin_radlen;
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
  // :322-346 heavy-gas Cherenkov
  drift = hcer_2_zentrance - hscin_1y_zpos;
  radw = drift / hair_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  radw = hcer_2_entr_thick / hcer_2_entr_radlen;
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
  drift = hcer_2_zmirror - hcer_2_zentrance;
  radw = drift / hcer_2_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  radw = hcer_mir_thick / hcer_mir_radlen;
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
  drift = hcer_2_zexit - hcer_2_zmirror;
  radw = drift / hcer_2_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  radw = hcer_2_exit_thick / hcer_2_exit_radlen;
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
  // :350-380 S2X, S2Y
  drift = hscin_2x_zpos - hcer_2_zexit;
  radw = drift / hair_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  if (t.ys > (hscin_2x_left + hscin_2y_offset) || t.ys < (hscin_2x_right + hscin_2y_offset)) {
    a.stop_code = shms_stop::S2X;
    return false;
  }
  radw = hscin_2x_thick / hscin_radlen;
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
  drift = hscin_2y_zpos - hscin_2x_zpos;
  radw = drift / hair_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  if (t.xs > (hscin_2y_bot + hscin_2x_offset) || t.xs < (hscin_2y_top + hscin_2x_offset)) {
    a.stop_code = shms_stop::S2Y;
    return false;
  }
  radw = hscin_2y_thick / hscin_radlen;
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
  // :384-398 calorimeter
  drift = hcal_4ta_zpos - hscin_2y_zpos;
  radw = drift / hair_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  if (t.ys > hcal_left || t.ys < hcal_right || t.xs > hcal_bottom || t.xs < hcal_top) {
    a.stop_code = shms_stop::CAL;
    return false;
  }
  // :402-424 track fit (REAL*4) and calorimeter fiducial cut on the fitted track
  for (int jchamber = 1; jchamber <= hdc_nr_cham; ++jchamber) {
    const int npl_off = (jchamber - 1) * hdc_nr_plan;
    for (int iplane = 1; iplane <= hdc_nr_plan; ++iplane) {
      const double z0 = (jchamber == 1) ? hdc_1_zpos : hdc_2_zpos;
      zdc[npl_off + iplane - 1] = (float)(z0 + (iplane - 0.5 - 0.5 * hdc_nr_plan) * hdc_del_plane);
    }
  }
  float dx, x0, dy, y0;
  lfit(zdc, xdc, 12, dx, x0);
  lfit(zdc, ydc, 12, dy, y0);
  a.x_fp = x0; a.y_fp = y0; a.dx_fp = dx; a.dy_fp = dy;
  const double xcal = a.x_fp + a.dx_fp * hcal_4ta_zpos;
  const double ycal = a.y_fp + a.dy_fp * hcal_4ta_zpos;
  if (ycal > (hcal_left - 5.0) || ycal < (hcal_right + 5.0) || xcal > (hcal_bottom - 5.0) ||
      xcal < (hcal_top + 5.0)) {
    a.stop_code = shms_stop::CAL_FID;
    return false;
  }
  return true;
}
}  // namespace shms
}  // namespace

// mc_shms, shms/mc_shms.f:1-1110 (use_sieve=.false., use_coll=.true., skip_hb=.false.)
void mc_shms(Track& t, const ArmOptics& o, ArmCall& a) {
  using namespace shms;
  if (o.fwd.n_classes() != 32) throw std::runtime_error("Bender-SHMS, wrong number of transport classes");
  const bool dec = a.decay_flag;
  a.ok_spec = false;
  a.stop_code = 0;
  a.reached_hut = false;
  bool dflag = false;
  t.dpps = a.dpp;
  double p = a.p_spec * (1. + t.dpps / 100.);
  t.xs = a.x; t.ys = a.y; t.zs = a.z; t.dxdzs = a.dxdz; t.dydzs = a.dydz;
  double& m2 = a.m2;
  double xt, yt, zdrift;
  auto stop = [&](int code) { a.stop_code = code; };
  auto r2 = [&]() { return t.xs * t.xs + t.ys * t.ys; };

  // :414-492 horizontal bender
  zdrift = zd_hbin;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  xt = t.xs; yt = t.ys;
  rotate_vaxis(t, 1.5, xt, yt);
  yt = yt + 1.51;
  if ((xt * xt > r_HBx * r_HBx) || (yt > r_HBfyp) || (yt < r_HBfym)) return stop(shms_stop::HB_IN);
  zdrift = zd_hbmen;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  xt = t.xs; yt = t.ys;
  rotate_vaxis(t, 1.5, xt, yt);
  yt = yt + 0.98;
  if ((xt * xt > r_HBx * r_HBx) || (yt > r_HBmenyp) || (yt < r_HBmenym)) return stop(shms_stop::HB_MEN);
  transp(t, o.fwd, 3, dec, dflag, m2, p, zd_hbmex, a.pathlen);
  xt = t.xs; yt = t.ys;
  rotate_vaxis(t, -1.5, xt, yt);
  yt = yt + 0.98;
  if ((xt * xt > r_HBx * r_HBx) || (yt > r_HBmexyp) || (yt < r_HBmexym)) return stop(shms_stop::HB_MEX);
  zdrift = zd_hbout;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  xt = t.xs; yt = t.ys;
  rotate_vaxis(t, -1.5, xt, yt);
  yt = yt + 1.51;
  if ((xt * xt > r_HBx * r_HBx) || (yt > r_HBbyp) || (yt < r_HBbym)) return stop(shms_stop::HB_OUT);

  // :496-560 collimator
  if (a.using_coll && (m2 > 100.0 * 100.0) && (m2 < 200.0 * 200.0)) {
    zdrift = z_entr;
    project(t, zdrift, dec, dflag, m2, p, a.pathlen);
    const CollGeom G{8.50, 12.50, 8.65, 12.85, 0.000, +0.000, 6.35, 0.42084};          // shms/mc_shms_coll.f:17-34
    if (!mc_coll(t, G, a, m2, p, dec, dflag)) return stop(shms_stop::COLL);
  } else {
    zdrift = z_entr;
    project(t, zdrift, dec, dflag, m2, p, a.pathlen);
    xt = t.xs; yt = t.ys;
    if (std::fabs(yt - y_off) > h_entr) return stop(shms_stop::SLIT_HOR);
    if (std::fabs(xt - x_off) > v_entr) return stop(shms_stop::SLIT_VERT);
    if (std::fabs(xt - x_off) > (-v_entr / h_entr * std::fabs(yt - y_off) + 3 * v_entr / 2))
      return stop(shms_stop::SLIT_OCT);
    zdrift = z_thick;
    project(t, zdrift, dec, dflag, m2, p, a.pathlen);
    xt = t.xs; yt = t.ys;
    if (std::fabs(yt - y_off) > (h_exit)) return stop(shms_stop::SLIT_HOR);
    if (std::fabs(xt - x_off) > (v_exit)) return stop(shms_stop::SLIT_VERT);
    if (std::fabs(xt - x_off) > ((-v_exit) / (h_exit)*std::fabs(yt - y_off) + 3 * (v_exit) / 2))
      return stop(shms_stop::SLIT_OCT);
  }
  // :566-656 Q1
  zdrift = zd_q1in - z_entr - z_thick;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if (r2() > r_Q1 * r_Q1) return stop(shms_stop::Q1_IN);
  zdrift = zd_q1men;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if (r2() > r_Q1 * r_Q1) return stop(shms_stop::Q1_MEN);
  transp(t, o.fwd, 7, dec, dflag, m2, p, zd_q1mid, a.pathlen);
  if (r2() > r_Q1 * r_Q1) return stop(shms_stop::Q1_MID);
  transp(t, o.fwd, 8, dec, dflag, m2, p, zd_q1mex, a.pathlen);
  if (r2() > r_Q1 * r_Q1) return stop(shms_stop::Q1_MEX);
  zdrift = zd_q1out;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if (r2() > r_Q1 * r_Q1) return stop(shms_stop::Q1_OUT);
  // :660-745 Q2
  zdrift = zd_q2in;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if (r2() > r_Q2 * r_Q2) return stop(shms_stop::Q2_IN);
  zdrift = zd_q2men;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if (r2() > r_Q2 * r_Q2) return stop(shms_stop::Q2_MEN);
  transp(t, o.fwd, 12, dec, dflag, m2, p, zd_q2mid, a.pathlen);
  if (r2() > r_Q2 * r_Q2) return stop(shms_stop::Q2_MID);
  transp(t, o.fwd, 13, dec, dflag, m2, p, zd_q2mex, a.pathlen);
  if (r2() > r_Q2 * r_Q2) return stop(shms_stop::Q2_MEX);
  zdrift = zd_q2out;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if (r2() > r_Q2 * r_Q2) return stop(shms_stop::Q2_OUT);
  // :749-834 Q3
  zdrift = zd_q3in;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if (r2() > r_Q3 * r_Q3) return stop(shms_stop::Q3_IN);
  zdrift = zd_q3men;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if (r2() > r_Q3 * r_Q3) return stop(shms_stop::Q3_MEN);
  transp(t, o.fwd, 17, dec, dflag, m2, p, zd_q3mid, a.pathlen);
  if (r2() > r_Q3 * r_Q3) return stop(shms_stop::Q3_MID);
  transp(t, o.fwd, 18, dec, dflag, m2, p, zd_q3mex, a.pathlen);
  if (r2() > r_Q3 * r_Q3) return stop(shms_stop::Q3_MEX);
  zdrift = zd_q3out;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if (r2() > r_Q3 * r_Q3) return stop(shms_stop::Q3_OUT);
  // :838-1040 dipole: entrance, flare, magnetic entrance, 7 mid planes, magnetic exit, exit
  zdrift = zd_q3d1trans;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if (r2() > r_D1 * r_D1) return stop(shms_stop::D1_IN);
  zdrift = zd_d1flare;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  xt = t.xs; yt = t.ys;
  rotate_haxis(t, 9.200, xt, yt);
  xt = xt - 3.5;
  if ((xt * xt + yt * yt) > r_D1 * r_D1) return stop(shms_stop::D1_FLR);
  zdrift = zd_d1men;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  xt = t.xs; yt = t.ys;
  rotate_haxis(t, 9.200, xt, yt);
  xt = xt + 2.82;
  if ((xt * xt + yt * yt) > r_D1 * r_D1) return stop(shms_stop::D1_MEN);
  static const double mid_ang[8] = {6.9, 4.6, 2.3, 0.0, -2.3, -4.6, -6.9, -9.2};
  static const double mid_off[8] = {8.05, 11.75, 13.96, 14.70, 13.96, 11.75, 8.05, 2.82};
  for (int k = 0; k < 8; ++k) {   // classes 23..30: mid1..mid7, mex
    transp(t, o.fwd, 23 + k, dec, dflag, m2, p, (k < 7) ? zd_d1mid : zd_d1mex, a.pathlen);
    xt = t.xs; yt = t.ys;
    rotate_haxis(t, mid_ang[k], xt, yt);
    xt = xt + mid_off[k];
    if ((xt * xt + yt * yt) > r_D1 * r_D1) return stop(shms_stop::D1_MID1 + k);
  }
  zdrift = zd_d1out;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  xt = t.xs; yt = t.ys;
  rotate_haxis(t, -9.20, xt, yt);
  xt = xt - 6.88;
  if ((xt * xt + yt * yt) > r_D1 * r_D1) return stop(shms_stop::D1_OUT);
  // :1044-1060 drift to the Cherenkov entrance (cer_flag=.true.)
  zdrift = zd_fp + hcer_1_zentrance;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  a.reached_hut = true;
  if (!hut(t, a, m2, p, dflag)) return;
  // :1076-1094 recon
  t.xs = a.x_fp; t.ys = a.y_fp; t.dxdzs = a.dx_fp; t.dydzs = a.dy_fp;
  double dpp_recon, dth_recon, dph_recon, y_recon;
  if (t.calls) t.calls[47]++;
  o.rec.eval(t, a.fry, dpp_recon, dth_recon, dph_recon, y_recon);
  a.dpp = dpp_recon;
  a.dxdz = dph_recon;
  a.dydz = dth_recon;
  a.y = y_recon;
  a.ok_spec = true;
}


// =====================================================================================
// SOS
// =====================================================================================
namespace {
namespace sos {
// sos/apertures_sos.inc
constexpr double r2_quad = 163.84, w_bm01 = 8.0, w_bm02 = 8.0;
constexpr double t_bm01_in = 41.32, b_bm01_in = -30.52, t_bm01_out = 53.05, b_bm01_out = -65.07;
constexpr double t_bm02_in = 51.73, b_bm02_in = -66.38, t_bm02_out = 51.52, b_bm02_out = -55.85;
constexpr double w_exit = 8.57, t_exit = 51.52, b_exit = -55.85;
// sos/mc_sos.f:59-67
constexpr double h_entr = 7.201, v_entr = 4.696, h_exit = 7.567, v_exit = 4.935;
constexpr double z_entr = 126.3e0, z_exit = z_entr + 6.3e0;
// sos/mc_sos_hut.f:45-191
constexpr double sfoil_exit_radlen = 53.3, sfoil_exit_thick = 0.020 * 2.54, sfoil_exit_zpos = -3.22;
constexpr double hut_pi = 3.141592654, hut_d_r = hut_pi / 180.;
constexpr double sfoil_exit_ang = 0. * hut_d_r;
constexpr double sair_radlen = 30420.;
constexpr double sdc_radlen = 16700.0, sdc_thick = 0.61775;
constexpr double sdc_wire_radlen = 0.35, sdc_wire_thick = 0.0000354;
constexpr double sdc_cath_radlen = 28.7, sdc_cath_thick = 0.0005 * 2.54;
constexpr double sscin_radlen = 42.4;
constexpr double scer_entr_radlen = 8.90, scer_entr_thick = 0.050, scer_radlen = 4810.0;
constexpr double scer_mir_radlen = 400.0, scer_mir_thick = 2.0, scer_exit_radlen = 8.90, scer_exit_thick = 0.050;
constexpr double sdc_sigma = 0.030;
constexpr int sdc_nr_cham = 2, sdc_nr_plan = 6;
constexpr double sdc_1_zpos = 6.25, sdc_2_zpos = 55.77;
constexpr double sdc_del_plane = sdc_thick + sdc_wire_thick + sdc_cath_thick;
constexpr double sdc_1_left = 24.0, sdc_1_right = -24.0, sdc_1y_offset = -1.822, sdc_1_top = -32.0, sdc_1_bot = 32.0,
                 sdc_1x_offset = 8.649;
constexpr double sdc_2_left = 24.0, sdc_2_right = -24.0, sdc_2y_offset = -1.976, sdc_2_top = -32.0, sdc_2_bot = 32.0,
                 sdc_2x_offset = -1.532;
constexpr double sscin_1y_zpos = 73.61, sscin_1x_zpos = 97.11, sscin_2y_zpos = 249.51, sscin_2x_zpos = 290.81;
constexpr double sscin_1x_thick = 1.040, sscin_1y_thick = 1.098, sscin_2x_thick = 1.040, sscin_2y_thick = 1.098;
constexpr double sscin_1_left = 18.25, sscin_1_right = -18.25, sscin_1x_offset = 2.8, sscin_1_top = -31.75,
                 sscin_1_bot = 31.75, sscin_1y_offset = 2.25;
constexpr double sscin_2_left = 18.25, sscin_2_right = -18.25, sscin_2x_offset = 4.9, sscin_2_top = -56.25,
                 sscin_2_bot = 56.25, sscin_2y_offset = 2.9;
constexpr double scer_zentrance = 130.000, scer_zmirror = 155.000, scer_zexit = 160.000;
constexpr double scal_4ta_zpos = 346.01;
constexpr int scintrig = 3;

// mc_sos_hut, sos/mc_sos_hut.f:1-548
bool hut(Track& t, ArmCall& a, double& m2, double& p, bool& dflag, double zinit) {
  Rng& r = *t.rng;
  const bool ms = a.ms_flag, wcs = a.wcs_flag, dec = a.decay_flag;
  double radw, drift;
  float xdc[12], ydc[12], zdc[12];
  // :251-257
  const double tmpran = r.grnd();
  a.resmult = (tmpran < 0.15) ? 2.0 : 1.0;
  for (int i = 0; i < 12; ++i) { xdc[i] = 0.f; ydc[i] = 0.f; }
  int scincount = 0;
  // :282-287 exit foil (the tilt angle is zero in this version: tan = 0, cos = 1)
  const double xt = t.xs + t.dxdzs * sfoil_exit_zpos;
  drift = (sfoil_exit_zpos + xt * std::tan(sfoil_exit_ang)) - zinit;
  if (drift <= 0.001) drift = 0.001;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  radw = sfoil_exit_thick / sfoil_exit_radlen / std::cos(sfoil_exit_ang);
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
  // :297-301 air to the first chamber
  drift = (sdc_1_zpos - 0.5 * sdc_nr_plan * sdc_del_plane) - (sfoil_exit_zpos + xt * std::tan(sfoil_exit_ang));
  radw = drift / sair_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  for (int jchamber = 1; jchamber <= 2; ++jchamber) {
    const int npl_off = (jchamber - 1) * sdc_nr_plan;
    for (int iplane = 1; iplane <= sdc_nr_plan; ++iplane) {   // :305-334 / :354-384
      radw = sdc_cath_thick / sdc_cath_radlen;
      if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
      drift = 0.5 * sdc_thick;
      radw = drift / sdc_radlen;
      if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
      drift = drift + sdc_cath_thick;
      project(t, drift, dec, dflag, m2, p, a.pathlen);
      radw = sdc_wire_thick / sdc_wire_radlen;
      if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
      double tmpran1 = 0., tmpran2 = 0.;
      if (wcs) { tmpran1 = gauss1(r, 99.0); tmpran2 = gauss1(r, 99.0); }
      xdc[npl_off + iplane - 1] = (float)(t.xs + sdc_sigma * tmpran1 * a.resmult);
      ydc[npl_off + iplane - 1] = (float)(t.ys + sdc_sigma * tmpran2 * a.resmult);
      if (iplane == 1 || iplane == 3 || iplane == 5) xdc[npl_off + iplane - 1] = 0.f;
      else ydc[npl_off + iplane - 1] = 0.f;
      drift = 0.5 * sdc_thick;
      radw = drift / sdc_radlen;
      if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
      drift = drift + sdc_wire_thick;
      project(t, drift, dec, dflag, m2, p, a.pathlen);
    }
    if (jchamber == 1) {   // :335-350
      if (t.xs > (sdc_1_bot - sdc_1x_offset) || t.xs < (sdc_1_top - sdc_1x_offset) ||
          t.ys > (sdc_1_left - sdc_1y_offset) || t.ys < (sdc_1_right - sdc_1y_offset)) {
        a.stop_code = sos_stop::DC1;
        return false;
      }
      radw = sdc_cath_thick / sdc_cath_radlen;
      if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
      drift = sdc_2_zpos - sdc_1_zpos - sdc_nr_plan * sdc_del_plane;
      radw = drift / sair_radlen;
      project(t, drift, dec, dflag, m2, p, a.pathlen);
      if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
    } else {               // :385-393
      if (t.xs > (sdc_2_bot - sdc_2x_offset) || t.xs < (sdc_2_top - sdc_2x_offset) ||
          t.ys > (sdc_2_left - sdc_2y_offset) || t.ys < (sdc_2_right - sdc_2y_offset)) {
        a.stop_code = sos_stop::DC2;
        return false;
      }
      radw = sdc_cath_thick / sdc_cath_radlen;
      if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
    }
  }
  // :399-414
  for (int jchamber = 1; jchamber <= sdc_nr_cham; ++jchamber) {
    const int npl_off = (jchamber - 1) * sdc_nr_plan;
    for (int iplane = 1; iplane <= sdc_nr_plan; ++iplane) {
      const double z0 = (jchamber == 1) ? sdc_1_zpos : sdc_2_zpos;
      zdc[npl_off + iplane - 1] = (float)(z0 + (iplane - 0.5 - 0.5 * sdc_nr_plan) * sdc_del_plane);
    }
  }
  float dxfp4, xfp4, dyfp4, yfp4;
  lfit(zdc, xdc, 12, dxfp4, xfp4);
  lfit(zdc, ydc, 12, dyfp4, yfp4);
  a.x_fp = (double)xfp4;
  a.y_fp = (double)yfp4;
  a.dx_fp = (double)dxfp4;
  a.dy_fp = (double)dyfp4;
  auto in_s1 = [&]() {
    return t.ys < (sscin_1_left + sscin_1y_offset) && t.ys > (sscin_1_right + sscin_1y_offset) &&
           t.xs < (sscin_1_bot + sscin_1x_offset) && t.xs > (sscin_1_top + sscin_1x_offset);
  };
  auto in_s2 = [&]() {
    return t.ys < (sscin_2_left + sscin_2y_offset) && t.ys > (sscin_2_right + sscin_2y_offset) &&
           t.xs < (sscin_2_bot + sscin_2x_offset) && t.xs > (sscin_2_top + sscin_2x_offset);
  };
  // :418-438 S1Y, S1X
  drift = sscin_1y_zpos - sdc_2_zpos - 0.5 * sdc_nr_plan * sdc_del_plane;
  radw = drift / sair_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  if (in_s1()) scincount++;
  radw = sscin_1y_thick / sscin_radlen;
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
  drift = sscin_1x_zpos - sscin_1y_zpos;
  radw = drift / sair_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  if (in_s1()) scincount++;
  radw = sscin_1x_thick / sscin_radlen;
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
  // :442-464 Cherenkov
  drift = scer_zentrance - sscin_1x_zpos;
  radw = drift / sair_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  radw = scer_entr_thick / scer_entr_radlen;
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
  drift = scer_zmirror - scer_zentrance;
  radw = drift / scer_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  radw = scer_mir_thick / scer_mir_radlen;
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
  drift = scer_zexit - scer_zmirror;
  radw = drift / scer_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  radw = scer_exit_thick / scer_exit_radlen;
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
  // :468-488 S2Y, S2X
  drift = sscin_2y_zpos - scer_zexit;
  radw = drift / sair_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  if (in_s2()) scincount++;
  radw = sscin_2y_thick / sscin_radlen;
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
  drift = sscin_2x_zpos - sscin_2y_zpos;
  radw = drift / sair_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  if (in_s2()) scincount++;
  radw = sscin_2x_thick / sscin_radlen;
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
  if (scincount < scintrig) {
    a.stop_code = sos_stop::SCIN;
    return false;
  }
  // :500-503 calorimeter (no cut)
  drift = scal_4ta_zpos - sscin_2x_zpos;
  radw = drift / sair_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  return true;
}
}  // namespace sos
}  // namespace

// mc_sos, sos/mc_sos.f:1-394 (use_sieve = .false.)
void mc_sos(Track& t, const ArmOptics& o, ArmCall& a) {
  using namespace sos;
  if (o.fwd.n_classes() != 10) throw std::runtime_error("MC_SOS, wrong number of transport classes");
  const bool dec = a.decay_flag;
  a.ok_spec = false;
  a.stop_code = 0;
  a.reached_hut = false;
  bool dflag = false;
  t.xs = a.x; t.ys = a.y; t.zs = a.z; t.dxdzs = a.dxdz; t.dydzs = a.dydz;
  t.dpps = a.dpp;
  double p = a.p_spec * (1. + t.dpps / 100.);
  double& m2 = a.m2;
  double xt, yt, zdrift;
  auto stop = [&](int code) { a.stop_code = code; };
  // :167-197 slit
  zdrift = z_entr;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if (std::fabs(t.ys) > h_entr) return stop(sos_stop::SLIT_HOR);
  if (std::fabs(t.xs) > v_entr) return stop(sos_stop::SLIT_VERT);
  if (std::fabs(t.xs) > (-v_entr / h_entr * std::fabs(t.ys) + 3 * v_entr / 2)) return stop(sos_stop::SLIT_OCT);
  zdrift = z_exit - z_entr;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if (std::fabs(t.ys) > h_exit) return stop(sos_stop::SLIT_HOR);
  if (std::fabs(t.xs) > v_exit) return stop(sos_stop::SLIT_VERT);
  if (std::fabs(t.xs) > (-v_exit / h_exit * std::fabs(t.ys) + 3 * v_exit / 2)) return stop(sos_stop::SLIT_OCT);
  // :202-224 quad
  zdrift = o.fwd.cls[0].driftdist - z_exit;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if ((t.xs * t.xs + t.ys * t.ys) > r2_quad) return stop(sos_stop::QUAD_IN);
  transp(t, o.fwd, 2, dec, dflag, m2, p, 35.0e0, a.pathlen);
  if ((t.xs * t.xs + t.ys * t.ys) > r2_quad) return stop(sos_stop::QUAD_MID);
  transp(t, o.fwd, 3, dec, dflag, m2, p, 35.0e0, a.pathlen);
  if ((t.xs * t.xs + t.ys * t.ys) > r2_quad) return stop(sos_stop::QUAD_OUT);
  // :228-277 the two bending magnets
  transp(t, o.fwd, 4, dec, dflag, m2, p, 80.0e0, a.pathlen);
  xt = t.xs; yt = t.ys;
  rotate_haxis(t, -45.0e0, xt, yt);
  if ((yt > w_bm01) || (-yt > (w_bm01 - 0.05 * 2.54)) || (-xt > t_bm01_in) || (-xt < b_bm01_in))
    return stop(sos_stop::BM01_IN);
  transp(t, o.fwd, 5, dec, dflag, m2, p, 169.52e0, a.pathlen);
  xt = t.xs; yt = t.ys;
  rotate_haxis(t, 45.0e0, xt, yt);
  if ((std::fabs(yt) > w_bm01) || (-xt > t_bm01_out) || (-xt < b_bm01_out)) return stop(sos_stop::BM01_OUT);
  transp(t, o.fwd, 6, dec, dflag, m2, p, 80.80e0, a.pathlen);
  xt = t.xs; yt = t.ys;
  rotate_haxis(t, 49.0e0, xt, yt);
  if ((std::fabs(yt) > w_bm02) || (-xt > t_bm02_in) || (-xt < b_bm02_in)) return stop(sos_stop::BM02_IN);
  transp(t, o.fwd, 7, dec, dflag, m2, p, 77.06e0, a.pathlen);
  xt = t.xs; yt = t.ys;
  rotate_haxis(t, 57.0e0, xt, yt);
  if ((std::fabs(yt) > w_bm02) || (-xt > t_bm02_out) || (-xt < b_bm02_out)) return stop(sos_stop::BM02_OUT);
  // :307-347 exit flange, new exit aperture, extension box
  transp(t, o.fwd, 8, dec, dflag, m2, p, 43.82e0, a.pathlen);
  xt = t.xs; yt = t.ys;
  rotate_haxis(t, 45.0e0, xt, yt);
  if ((std::fabs(yt) > w_exit) || (-xt > t_exit) || (-xt < b_exit)) return stop(sos_stop::EXIT);
  transp(t, o.fwd, 9, dec, dflag, m2, p, 44.34e0, a.pathlen);
  const double tmpwidth = 10.998 + 0.10209 * (t.xs + 37.694);
  if ((std::fabs(t.xs) > 37.694) || std::fabs(t.ys) > tmpwidth) return stop(sos_stop::EXIT);
  t.xs = t.xs + 7.62 * t.dxdzs;
  t.ys = t.ys + 7.62 * t.dydzs;
  if ((std::fabs(t.xs) > 38.1) || std::fabs(t.ys) > 12.7) return stop(sos_stop::EXIT);
  // :364-370 hut
  a.reached_hut = true;
  if (!hut(t, a, m2, p, dflag, -3.206267e0)) return;
  // :373-388 recon
  t.xs = a.x_fp; t.ys = a.y_fp; t.dxdzs = a.dx_fp; t.dydzs = a.dy_fp;
  double dpp_recon, dth_recon, dph_recon, y_recon;
  if (t.calls) t.calls[47]++;
  o.rec.eval(t, a.fry, dpp_recon, dth_recon, dph_recon, y_recon, /*clamp_all=*/true);
  a.dpp = dpp_recon;
  a.dxdz = dph_recon;
  a.dydz = dth_recon;
  a.y = y_recon;
  a.ok_spec = true;
}

// =====================================================================================
// HRS (left = spectrometer 4, right = spectrometer 3)
// =====================================================================================
namespace {
namespace hrs {
// hrsl/apertures_hrsl.inc == hrsr/apertures_hrsr.inc
constexpr double r_Q1 = 15.0, r_Q2 = 30.22, r_Q3 = 30.22;
// hrsl/mc_hrsl_hut.f:28-140 == hrsr/mc_hrsr_hut.f
constexpr double hfoil_exit_radlen = 3.56, hfoil_exit_thick = 0.01;
constexpr double hair_radlen = 30420.;
constexpr double hdc_entr_radlen = 34.4, hdc_entr_thick = 0.00018 * 2.54;
constexpr double hdc_radlen = 16700.0, hdc_thick = 1.5;
constexpr double hdc_wire_radlen = 0.35, hdc_wire_thick = 0.0000049;
constexpr double hdc_cath_radlen = 7.2, hdc_cath_thick = 0.000177;
constexpr double hdc_exit_radlen = 34.4, hdc_exit_thick = 0.00018 * 2.54;
constexpr double hscin_radlen = 42.4;
constexpr double hcer_entr_radlen = 8.90, hcer_entr_thick = 0.040 * 2.54, hcer_radlen = 36620.0;
constexpr double hcer_mir_radlen = 400.0, hcer_mir_thick = 2.0, hcer_exit_radlen = 8.90, hcer_exit_thick = 0.040 * 2.54;
constexpr double hdc_sigma = 0.0225;
constexpr int hdc_nr_cham = 2, hdc_nr_plan = 6;
constexpr double hdc_1_zpos = -25.0 + 25.0, hdc_2_zpos = 25.0 + 25.0;
constexpr double hdc_del_plane = hdc_thick + hdc_wire_thick + hdc_cath_thick;
constexpr double hdc_1_left = 14.4, hdc_1_right = -14.4, hdc_1y_offset = 0.000, hdc_1_top = -105.6, hdc_1_bot = 105.6,
                 hdc_1x_offset = 0.000;
constexpr double hdc_2_left = 14.4, hdc_2_right = -14.4, hdc_2y_offset = 0.000, hdc_2_top = -105.6, hdc_2_bot = 105.6,
                 hdc_2x_offset = 0.000;
constexpr double hscin_1x_zpos = 95.0 + 25.0, hscin_2x_zpos = 288.3 + 25.0;
constexpr double hscin_1x_thick = 0.5 * 1.067, hscin_2x_thick = 0.5 * 1.067;
constexpr double hscin_1x_left = 18.0, hscin_1x_right = -18.0, hscin_2x_left = 30.0, hscin_2x_right = -30.0;
constexpr double hcer_zentrance = 137.0 + 25.0, hcer_zmirror = 197.0 + 25.0, hcer_zexit = 237.0 + 25.0;
constexpr double hcal_4ta_zpos = 407.3 + 25.0;

// mc_hrsl_hut / mc_hrsr_hut, hrsl/mc_hrsl_hut.f:1-470
bool hut(Track& t, ArmCall& a, double& m2, double& p, bool& dflag, double zinit) {
  Rng& r = *t.rng;
  const bool ms = a.ms_flag, wcs = a.wcs_flag, dec = a.decay_flag;
  double radw, drift, xt, yt;
  float xdc[12], ydc[12], zdc[12];
  for (int i = 0; i < 12; ++i) { xdc[i] = 0.f; ydc[i] = 0.f; }
  a.resmult = 1.0;                                            // :184
  // :190-200 exit foil, air to the first VDC
  radw = hfoil_exit_thick / hfoil_exit_radlen;
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
  drift = (hdc_1_zpos - 0.5 * hdc_nr_plan * hdc_del_plane) - zinit;
  radw = drift / hair_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  for (int jchamber = 1; jchamber <= 2; ++jchamber) {
    radw = hdc_entr_thick / hdc_entr_radlen;
    if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
    const int npl_off = (jchamber - 1) * hdc_nr_plan;
    for (int iplane = 1; iplane <= hdc_nr_plan; ++iplane) {
      radw = hdc_cath_thick / hdc_cath_radlen;
      if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
      drift = 0.5 * hdc_thick;
      radw = drift / hdc_radlen;
      if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
      drift = drift + hdc_cath_thick;
      project(t, drift, dec, dflag, m2, p, a.pathlen);
      radw = hdc_wire_thick / hdc_wire_radlen;
      if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
      double tmpran1 = 0., tmpran2 = 0.;
      if (wcs) { tmpran1 = gauss1(r, 99.0); tmpran2 = gauss1(r, 99.0); }
      xdc[npl_off + iplane - 1] = (float)(t.xs + hdc_sigma * tmpran1 * a.resmult);
      ydc[npl_off + iplane - 1] = (float)(t.ys + hdc_sigma * tmpran2 * a.resmult);
      if (iplane == 1 || iplane == 3 || iplane == 5) xdc[npl_off + iplane - 1] = 0.f;
      else ydc[npl_off + iplane - 1] = 0.f;
      drift = 0.5 * hdc_thick;
      radw = drift / hdc_radlen;
      if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
      drift = drift + hdc_wire_thick;
      project(t, drift, dec, dflag, m2, p, a.pathlen);
    }
    radw = hdc_exit_thick / hdc_exit_radlen;
    if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
    // the VDC frame is tested in the chamber plane, tilted 45 degrees
    xt = t.xs; yt = t.ys;
    rotate_haxis(t, hrs_lit(45.0, "45.0"), xt, yt);          // 45.0d0 in the reference (mc_hrsl_hut.f:293,356)
    if (jchamber == 1) {
      if (xt > (hdc_1_bot - hdc_1x_offset) || xt < (hdc_1_top - hdc_1x_offset) ||
          yt > (hdc_1_left - hdc_1y_offset) || yt < (hdc_1_right - hdc_1y_offset)) {
        a.stop_code = hrs_stop::DC1;
        return false;
      }
      radw = hdc_cath_thick / hdc_cath_radlen;
      if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
      drift = hdc_2_zpos - hdc_1_zpos - hdc_nr_plan * hdc_del_plane;
      radw = drift / hair_radlen;
      project(t, drift, dec, dflag, m2, p, a.pathlen);
      if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
    } else {
      if (xt > (hdc_2_bot - hdc_2x_offset) || xt < (hdc_2_top - hdc_2x_offset) ||
          yt > (hdc_2_left - hdc_2y_offset) || yt < (hdc_2_right - hdc_2y_offset)) {
        a.stop_code = hrs_stop::DC2;
        return false;
      }
      radw = hdc_cath_thick / hdc_cath_radlen;
      if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
    }
  }
  for (int jchamber = 1; jchamber <= hdc_nr_cham; ++jchamber) {
    const int npl_off = (jchamber - 1) * hdc_nr_plan;
    for (int iplane = 1; iplane <= hdc_nr_plan; ++iplane) {
      const double z0 = (jchamber == 1) ? hdc_1_zpos : hdc_2_zpos;
      zdc[npl_off + iplane - 1] = (float)(z0 + (iplane - 0.5 - 0.5 * hdc_nr_plan) * hdc_del_plane);
    }
  }
  float dxfp4, xfp4, dyfp4, yfp4;
  lfit(zdc, xdc, 12, dxfp4, xfp4);
  lfit(zdc, ydc, 12, dyfp4, yfp4);
  a.x_fp = (double)xfp4;
  a.y_fp = (double)yfp4;
  a.dx_fp = (double)dxfp4;
  a.dy_fp = (double)dyfp4;
  // S1
  drift = hscin_1x_zpos - hdc_2_zpos - 0.5 * hdc_nr_plan * hdc_del_plane;
  radw = drift / hair_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  if (t.ys > hscin_1x_left || t.ys < hscin_1x_right) {
    a.stop_code = hrs_stop::S1;
    return false;
  }
  radw = hscin_1x_thick / hscin_radlen;
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
  // Cherenkov
  drift = hcer_zentrance - hscin_1x_zpos;
  radw = drift / hair_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  radw = hcer_entr_thick / hcer_entr_radlen;
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
  drift = hcer_zmirror - hcer_zentrance;
  radw = drift / hcer_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  radw = hcer_mir_thick / hcer_mir_radlen;
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
  drift = hcer_zexit - hcer_zmirror;
  radw = drift / hcer_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  radw = hcer_exit_thick / hcer_exit_radlen;
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
  // S2
  drift = hscin_2x_zpos - hcer_zexit;
  radw = drift / hair_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  if (t.ys > hscin_2x_left || t.ys < hscin_2x_right) {
    a.stop_code = hrs_stop::S2;
    return false;
  }
  radw = hscin_2x_thick / hscin_radlen;
  if (ms) musc(r, m2, p, radw, t.dydzs, t.dxdzs);
  // calorimeter (no cut)
  drift = hcal_4ta_zpos - hscin_2x_zpos;
  radw = drift / hair_radlen;
  project(t, drift, dec, dflag, m2, p, a.pathlen);
  if (ms) musc_ext(r, m2, p, radw, drift, t.dydzs, t.dxdzs, t.ys, t.xs);
  return true;
}
}  // namespace hrs
}  // namespace

// mc_hrsl (hrsl/mc_hrsl.f:1-551) and mc_hrsr (hrsr/mc_hrsr.f); they differ in the slit exit
// half-gap (:58), the slit distance (:67) and the focal-plane y offset (:525 / :524).
// The Makefile's -fdefault-real-8 makes the bare literals of rotate_haxis(-30.0,..) 8-byte reals.
void mc_hrs(Track& t, const ArmOptics& o, ArmCall& a, bool right) {
  using namespace hrs;
  if (o.fwd.n_classes() != 12)
    throw std::runtime_error(right ? "MC_HRSR, wrong number of transport classes"
                                   : "MC_HRSL, wrong number of transport classes");
  const double h_entr = 3.145, v_entr = 6.090, h_exit = right ? 3.340 : 3.335, v_exit = 6.485;
  const double y_off = 0.0, x_off = 0.0, z_off = 0.0;
  const double z_entr = (right ? 110.0 : 110.9) + z_off, z_exit = z_entr + 8.0;
  const bool dec = a.decay_flag;
  a.ok_spec = false;
  a.stop_code = 0;
  a.reached_hut = false;
  bool dflag = false;
  t.xs = a.x; t.ys = a.y; t.zs = a.z; t.dxdzs = a.dxdz; t.dydzs = a.dydz;
  t.dpps = a.dpp;
  double p = a.p_spec * (1. + t.dpps / 100.);
  double& m2 = a.m2;
  double xt, yt, zdrift, ztmp;
  auto stop = [&](int code) { a.stop_code = code; };
  auto rad = [&]() { return std::sqrt(t.xs * t.xs + t.ys * t.ys); };
  auto r2 = [&]() { return t.xs * t.xs + t.ys * t.ys; };
  // :159-179 scattering-chamber exit pipes
  zdrift = 65.686;
  ztmp = zdrift;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if (rad() > 7.3787) return stop(hrs_stop::SLIT_HOR);
  zdrift = 80.436 - ztmp;
  ztmp = 80.436;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if (rad() > 7.4092) return stop(hrs_stop::SLIT_HOR);
  // :184-218 rectangular collimator
  zdrift = z_entr - ztmp;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if (std::fabs(t.ys - y_off) > h_entr) return stop(hrs_stop::SLIT_HOR);
  if (std::fabs(t.xs - x_off) > v_entr) return stop(hrs_stop::SLIT_VERT);
  zdrift = z_exit - z_entr;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if (std::fabs(t.ys - y_off) > h_exit) return stop(hrs_stop::SLIT_HOR);
  if (std::fabs(t.xs - x_off) > v_exit) return stop(hrs_stop::SLIT_VERT);
  // :222-279 Q1
  ztmp = 135.064;
  zdrift = ztmp - z_exit;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if (rad() > 12.5222) return stop(hrs_stop::Q1_IN);
  zdrift = o.fwd.cls[0].driftdist - ztmp;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if (r2() > r_Q1 * r_Q1) return stop(hrs_stop::Q1_IN);
  transp(t, o.fwd, 2, dec, dflag, m2, p, hrs_lit(62.75333333, "62.75333333"), a.pathlen);
  if (r2() > r_Q1 * r_Q1) return stop(hrs_stop::Q1_MID);
  transp(t, o.fwd, 3, dec, dflag, m2, p, hrs_lit(31.37666667, "31.37666667"), a.pathlen);
  if (r2() > r_Q1 * r_Q1) return stop(hrs_stop::Q1_OUT);
  zdrift = 300.464 - 253.16;
  ztmp = zdrift;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if (rad() > 14.9225) return stop(hrs_stop::Q1_OUT);
  // :281-349 Q2
  zdrift = 314.464 - 300.464;
  ztmp = ztmp + zdrift;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if (rad() > 20.9550) return stop(hrs_stop::Q2_IN);
  zdrift = o.fwd.cls[3].driftdist - ztmp;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if (r2() > r_Q2 * r_Q2) return stop(hrs_stop::Q2_IN);
  transp(t, o.fwd, 5, dec, dflag, m2, p, hrs_lit(121.77333333, "121.77333333"), a.pathlen);
  if (r2() > r_Q2 * r_Q2) return stop(hrs_stop::Q2_MID);
  transp(t, o.fwd, 6, dec, dflag, m2, p, hrs_lit(60.88666667, "60.88666667"), a.pathlen);
  if (r2() > r_Q2 * r_Q2) return stop(hrs_stop::Q2_OUT);
  zdrift = 609.664 - 553.020;
  ztmp = zdrift;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if (rad() > 30.0073) return stop(hrs_stop::Q2_OUT);
  zdrift = 641.800 - 609.664;
  ztmp = ztmp + zdrift;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if (rad() > 30.0073) return stop(hrs_stop::Q2_OUT);
  // :351-406 dipole
  zdrift = 819.489 - 641.800;
  ztmp = ztmp + zdrift;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if (std::fabs(t.xs) > 50.0 || std::fabs(t.ys) > 15.0) return stop(hrs_stop::D1_IN);
  zdrift = o.fwd.cls[6].driftdist - ztmp;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  xt = t.xs; yt = t.ys;
  rotate_haxis(t, -30.0, xt, yt);
  if (std::fabs(xt - 2.500) > 52.5) return stop(hrs_stop::D1_IN);
  if ((std::fabs(yt) + 0.01861 * xt) > 12.5) return stop(hrs_stop::D1_IN);
  transp(t, o.fwd, 8, dec, dflag, m2, p, hrs_lit(659.73445725, "659.73445725"), a.pathlen);
  xt = t.xs; yt = t.ys;
  rotate_haxis(t, 30.0, xt, yt);
  if (std::fabs(xt - 2.500) > 52.5) return stop(hrs_stop::D1_OUT);
  if ((std::fabs(yt) + 0.01861 * xt) > 12.5) return stop(hrs_stop::D1_OUT);
  zdrift = 1745.33546 - 1655.83446;
  ztmp = zdrift;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if (rad() > 30.3276) return stop(hrs_stop::D1_OUT);
  if (std::fabs(t.xs) > 50.0 || std::fabs(t.ys) > 15.0) return stop(hrs_stop::D1_OUT);
  // :429-499 Q3
  zdrift = 1759.00946 - 1745.33546;
  ztmp = ztmp + zdrift;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if (rad() > 30.3276) return stop(hrs_stop::Q3_IN);
  zdrift = o.fwd.cls[8].driftdist - ztmp;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if (r2() > r_Q3 * r_Q3) return stop(hrs_stop::Q3_IN);
  transp(t, o.fwd, 10, dec, dflag, m2, p, hrs_lit(121.7866667, "121.7866667"), a.pathlen);
  if (r2() > r_Q3 * r_Q3) return stop(hrs_stop::Q3_MID);
  transp(t, o.fwd, 11, dec, dflag, m2, p, hrs_lit(60.89333333, "60.89333333"), a.pathlen);
  if (r2() > r_Q3 * r_Q3) return stop(hrs_stop::Q3_OUT);
  zdrift = 2080.38746 - 1997.76446;
  ztmp = zdrift;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if (std::fabs(t.xs) > 35.56 || std::fabs(t.ys) > 17.145) return stop(hrs_stop::Q3_OUT);
  zdrift = 2327.47246 - 2080.38746;
  ztmp = ztmp + zdrift;
  project(t, zdrift, dec, dflag, m2, p, a.pathlen);
  if (std::fabs(t.xs) > 99.76635 || std::fabs(t.ys) > 17.145) return stop(hrs_stop::Q3_OUT);
  // :503-510 hut
  a.reached_hut = true;
  zdrift = o.fwd.cls[11].driftdist - ztmp;
  if (!hut(t, a, m2, p, dflag, -zdrift)) return;
  // :513-539 recon sees the fitted track; the returned y_fp is then shifted to the VDC centre
  t.xs = a.x_fp; t.ys = a.y_fp; t.dxdzs = a.dx_fp; t.dydzs = a.dy_fp;
  a.y_fp = a.y_fp - (right ? 0.48 : 0.78);
  double dpp_recon, dth_recon, dph_recon, y_recon;
  if (t.calls) t.calls[47]++;
  o.rec.eval(t, a.fry, dpp_recon, dth_recon, dph_recon, y_recon);
  a.dpp = dpp_recon;
  a.dxdz = dph_recon;
  a.dydz = dth_recon;
  a.y = y_recon;
  a.ok_spec = true;
}

}  // namespace simc_oracle
