// ORACLE -- TEST INFRASTRUCTURE ONLY (see event.hpp).
// Loop bookkeeping of simc.f:229-350 (pass_cuts, inc, counters, sumerr, limits_update) on top
// of one_try(), accumulating into the same exact integer representation the B200 path uses
// (include/simc_b200.h: simc_accum), so that the two can be compared bit for bit.
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>
#include "event.hpp"
#include "field.hpp"

namespace simc_oracle {

typedef __int128 i128;

static inline i128 to_fixed(double x, int qexp) {
  const double sc = std::ldexp(x, -qexp);
  if (!(std::fabs(sc) < 4.611686018427387904e18)) {        // >= 2^62: already an integer
    if (!(std::fabs(sc) < 1.0e38)) return 0;               // inf/nan/overflow: dropped
    const double hi = std::floor(sc / 18446744073709551616.0);
    const double lo = sc - hi * 18446744073709551616.0;
    return ((i128)(long long)hi << 64) + (i128)(unsigned long long)lo;
  }
  return (i128)std::llrint(sc);                            // round to nearest even
}
static inline void add_fixed(simc_fixed128& f, double x) {
  i128 v = ((i128)f.hi << 64) | (i128)f.lo;
  v += to_fixed(x, f.qexp);
  f.lo = (uint64_t)v;
  f.hi = (int64_t)(v >> 64);
}
static inline void upd(simc_range& r, double v) { r.lo = std::min(r.lo, v); r.hi = std::max(r.hi, v); }

static int weight_qexp(const simc_run_config& cfg) {
  const double w = cfg.w_ref > 0 ? cfg.w_ref : 1.0;
  return std::ilogb(w) - 64;
}

void accum_clear(const simc_run_config& cfg, simc_accum& a) {
  std::memset(&a, 0, sizeof(a));
  const int qw = weight_qexp(cfg);
  a.wtcontribute.qexp = qw;
  a.sum_sigcc.qexp = qw;
  for (int i = 0; i < 8; ++i) { a.sumerr[i].qexp = -80; a.sumerr2[i].qexp = -80; }
  for (int h = 0; h < 6; ++h) for (int b = 0; b < SIMC_NHIST; ++b) a.hist_w[h][b].qexp = qw;
  for (int i = 0; i < 32; ++i) { a.contrib[i].lo = 1.0e10; a.contrib[i].hi = -1.0e10; }   // init.f:921-1190
  for (int i = 0; i < 8; ++i) { a.slop[i].lo = 1.0e10; a.slop[i].hi = -1.0e10; }
}

// simc.f:624-640
static inline int hist_bin(const simc_axis& ax, double val) {
  const double r = std::round(0.5 + (val - ax.min) / ax.bin);     // nint: half away from zero
  if (!(r >= 1.0 && r <= (double)SIMC_NHIST)) return -1;
  return (int)r - 1;
}

void accumulate(const Sim& s, const TryResult& r, const EventMain& main, const Event& vertex, const Event& orig,
                const Event& recon, simc_accum& a) {
  const simc_run_config& cfg = *s.cfg;
  a.ntried++;
  // STOP counters of the arms that were entered
  auto stops = [&](int which, int code, bool hut) {
    if (code < 0) return;
    a.stop[which][0]++;
    if (code == 0) a.stop[which][1]++;
    if (hut) a.stop[which][2]++;
    if (code > 0 && 2 + code < SIMC_NSTOP) a.stop[which][2 + code]++;
  };
  for (int w = 0; w < 2; ++w) for (int k = 0; k < 48; ++k) a.transp_calls[w][k] += s.calls[w][k];
  stops(1, s.stop_p, s.hut_p);
  stops(0, s.stop_e, s.hut_e);
  // steps spent in the collimator material bump the slit counters (mc_hms_coll.f:95-115), survivors included
  for (int w = 0; w < 2; ++w) {
    const int arm = w == 0 ? cfg.electron_arm : cfg.hadron_arm;
    const int first = arm == 1 ? hms_stop::SLIT_HOR : arm == 5 ? shms_stop::SLIT_HOR : -1;
    if (first < 0) continue;
    for (int k = 0; k < 3; ++k) a.stop[w][2 + first + k] += s.coll_steps[w][k];
  }
  if (r.success) add_fixed(a.sum_sigcc, main.sigcc);
  // geni: every try (simc.f:253-262)
  const double geni_vals[8] = {vertex.e.delta, vertex.e.yptar, -vertex.e.xptar, vertex.p.delta,
                               vertex.p.yptar, -vertex.p.xptar, vertex.Em, vertex.Pm};
  for (int k = 0; k < 8; ++k) {
    const int b = hist_bin(cfg.hist_axis[2][k], geni_vals[k]);
    if (b >= 0) a.hist_n[2][k][b]++;
  }
  if (!r.success) return;
  a.nsuccess++;
  // RECON (weighted) and gen histograms, simc.f:272-286
  const double rec_vals[6] = {main.RECON_e.delta, main.RECON_e.yptar, main.RECON_e.xptar,
                              main.RECON_p.delta, main.RECON_p.yptar, main.RECON_p.xptar};
  for (int k = 0; k < 6; ++k) {
    const int b = hist_bin(cfg.hist_axis[0][k], rec_vals[k]);
    if (b >= 0) add_fixed(a.hist_w[k][b], main.weight);
  }
  { int b = hist_bin(cfg.hist_axis[0][SIMC_H_EM], recon.Em); if (b >= 0) a.hist_n[0][SIMC_H_EM][b]++; }
  { int b = hist_bin(cfg.hist_axis[0][SIMC_H_PM], recon.Pm); if (b >= 0) a.hist_n[0][SIMC_H_PM][b]++; }
  for (int k = 0; k < 7; ++k) {     // H%gen has no Pm increment
    const int b = hist_bin(cfg.hist_axis[1][k], geni_vals[k]);
    if (b >= 0) a.hist_n[1][k][b]++;
  }
  a.ncontribute++;
  if (s.low_w) a.unsupported++;
  {   // weights the fixed-point sums drop (NaN, inf, out of range): counted, like the product does
    const int qw = a.wtcontribute.qexp;
    if (!(std::fabs(std::ldexp(main.weight, -qw)) < 1.0e38) || !(std::fabs(std::ldexp(main.sigcc, -qw)) < 1.0e38)) a.nonfinite++;
  }
  if (!s.rad.rad_proton_this_ev) a.ncontribute_no_rad_proton++;
  if (r.pass_cuts) {
    a.npasscuts++;
    add_fixed(a.wtcontribute, main.weight);
    const double err[8] = {recon.e.delta - main.SP_e.delta, recon.e.xptar - vertex.e.xptar,
                           recon.e.yptar - vertex.e.yptar, recon.e.z - main.SP_e.z,
                           recon.p.delta - main.SP_p.delta, recon.p.xptar - vertex.p.xptar,
                           recon.p.yptar - vertex.p.yptar, recon.p.z - main.SP_p.z};
    for (int k = 0; k < 8; ++k) { add_fixed(a.sumerr[k], err[k]); add_fixed(a.sumerr2[k], err[k] * err[k]); }
  }
  // limits_update, event.f:1-90
  simc_range* c = a.contrib;
  upd(c[0], vertex.e.delta); upd(c[1], vertex.e.yptar); upd(c[2], vertex.e.xptar);
  upd(c[3], vertex.p.delta); upd(c[4], vertex.p.yptar); upd(c[5], vertex.p.xptar);
  upd(c[6], main.Trec);
  if (cfg.doing_deuterium || cfg.doing_pion || cfg.doing_kaon || cfg.doing_delta || cfg.doing_rho)
    upd(c[7], vertex.e.E - main.Ein_shift);
  else
    upd(c[7], vertex.e.E + vertex.p.E - main.Ein_shift);
  upd(c[8], orig.e.E - main.Ee_shift); upd(c[9], orig.e.xptar); upd(c[10], orig.e.yptar);
  upd(c[11], orig.p.E); upd(c[12], orig.p.yptar); upd(c[13], orig.p.xptar);
  upd(c[14], orig.Em - main.Ein_shift + main.Ee_shift); upd(c[15], orig.Pm); upd(c[16], orig.Trec);
  upd(c[17], main.SP_e.delta); upd(c[18], main.SP_e.yptar); upd(c[19], main.SP_e.xptar);
  upd(c[20], main.SP_p.delta); upd(c[21], main.SP_p.yptar); upd(c[22], main.SP_p.xptar);
  upd(c[23], vertex.Trec); upd(c[24], vertex.Em); upd(c[25], vertex.Pm);
  for (int i = 0; i < 3; ++i) upd(c[26 + i], s.rad.Egamma_used[i]);
  upd(c[29], s.rad.Egamma_used[0] + s.rad.Egamma_used[1] + s.rad.Egamma_used[2]);
  simc_range* sl = a.slop;
  upd(sl[0], main.RECON_e.delta - main.SP_e.delta); upd(sl[1], main.RECON_e.yptar - main.SP_e.yptar);
  upd(sl[2], main.RECON_e.xptar - main.SP_e.xptar);
  upd(sl[3], main.RECON_p.delta - main.SP_p.delta); upd(sl[4], main.RECON_p.yptar - main.SP_p.yptar);
  upd(sl[5], main.RECON_p.xptar - main.SP_p.xptar);
  upd(sl[6], recon.Em - (orig.Em - main.Ein_shift + main.Ee_shift));
  upd(sl[7], std::fabs(recon.Pm) - std::fabs(orig.Pm));
}

void merge_accum(simc_accum& a, const simc_accum& b) {
  a.ntried += b.ntried; a.nsuccess += b.nsuccess; a.ncontribute += b.ncontribute; a.npasscuts += b.npasscuts;
  a.ncontribute_no_rad_proton += b.ncontribute_no_rad_proton;
  a.unsupported += b.unsupported;
  a.nonfinite += b.nonfinite;
  auto addf = [](simc_fixed128& x, const simc_fixed128& y) {
    i128 v = (((i128)x.hi << 64) | (i128)x.lo) + (((i128)y.hi << 64) | (i128)y.lo);
    x.lo = (uint64_t)v; x.hi = (int64_t)(v >> 64);
  };
  addf(a.wtcontribute, b.wtcontribute); addf(a.sum_sigcc, b.sum_sigcc);
  for (int i = 0; i < 8; ++i) { addf(a.sumerr[i], b.sumerr[i]); addf(a.sumerr2[i], b.sumerr2[i]); }
  for (int h = 0; h < 6; ++h) for (int k = 0; k < SIMC_NHIST; ++k) addf(a.hist_w[h][k], b.hist_w[h][k]);
  for (int s = 0; s < 3; ++s) for (int h = 0; h < SIMC_H_PER_SET; ++h) for (int k = 0; k < SIMC_NHIST; ++k)
    a.hist_n[s][h][k] += b.hist_n[s][h][k];
  for (int i = 0; i < 32; ++i) { a.contrib[i].lo = std::min(a.contrib[i].lo, b.contrib[i].lo); a.contrib[i].hi = std::max(a.contrib[i].hi, b.contrib[i].hi); }
  for (int i = 0; i < 8; ++i) { a.slop[i].lo = std::min(a.slop[i].lo, b.slop[i].lo); a.slop[i].hi = std::max(a.slop[i].hi, b.slop[i].hi); }
  for (int w = 0; w < 2; ++w) for (int i = 0; i < SIMC_NSTOP; ++i) a.stop[w][i] += b.stop[w][i];
  for (int w = 0; w < 2; ++w) for (int i = 0; i < 48; ++i) a.transp_calls[w][i] += b.transp_calls[w][i];
}

void fill_record(const Sim& s, const TryResult& r, const EventMain& main, const Event& vertex, const Event& orig,
                 const Event& recon, double* rec, int64_t n, int64_t i) {
  double v[SIMC_EVENT_NREC] = {
      (double)r.stage, (double)r.pass_cuts, (double)s.rng->draw, (double)s.stop_p, (double)s.stop_e,
      main.weight, main.sigcc, main.gen_weight, main.jacobian, main.sigcc_recon,
      vertex.Ein, vertex.e.E, vertex.e.delta, vertex.e.yptar, vertex.e.xptar,
      vertex.p.E, vertex.p.delta, vertex.p.yptar, vertex.p.xptar, vertex.Q2,
      orig.e.E, orig.p.E, s.rad.Egamma_used[0], s.rad.Egamma_used[1], s.rad.Egamma_used[2], (double)s.rad.ntail,
      main.target.x, main.target.y, main.target.z, main.target.Eloss[0], main.target.Eloss[1], main.target.Eloss[2],
      main.SP_e.delta, main.SP_e.yptar, main.SP_e.xptar, main.SP_p.delta, main.SP_p.yptar, main.SP_p.xptar,
      recon.e.delta, recon.e.yptar, recon.e.xptar, recon.p.delta, recon.p.yptar, recon.p.xptar,
      recon.Em, recon.Pm, recon.W, s.rad.hardcorfac,
      main.thetacm, main.phicm, s.ntup.sigcm, main.davejac, s.ntup.survivalprob, s.ntup.mm, main.wcm, main.t,
      s.cfg->doing_rho ? orig.p.yptar : 0.0, s.cfg->doing_rho ? orig.p.xptar : 0.0, s.ntup.rhomass, s.ntup.rhotheta};
  for (int k = 0; k < SIMC_EVENT_NREC; ++k) rec[k * n + i] = v[k];
}

// results_ntu_write, results_write.f:1-269
int fill_ntuple(const Sim& s, const EventMain& main, const Event& vertex, const Event& orig, const Event& recon,
                double* out) {
  const simc_run_config& cfg = *s.cfg;
  double ntu[81] = {0};
  double corrsing = 0, Pm_Heepx = 0, Pm_Heepy = 0, Pm_Heepz = 0;
  const bool eep = cfg.doing_hyd_elast || cfg.doing_deuterium || cfg.doing_heavy;
  if (eep) {
    const double poftheta = K::Mp * cfg.Ebeam / (2 * cfg.Ebeam * powi(std::sin(recon.e.theta / 2.), 2) + K::Mp);
    corrsing = recon.e.P - poftheta;
    Pm_Heepz = -(recon.Pmy * recon.uq.y + recon.Pmz * recon.uq.z) / std::sqrt(recon.uq.y * recon.uq.y + recon.uq.z * recon.uq.z);
    Pm_Heepy = (recon.Pmz * recon.uq.y - recon.Pmy * recon.uq.z) / std::sqrt(recon.uq.y * recon.uq.y + recon.uq.z * recon.uq.z);
    Pm_Heepx = -recon.Pmx;
  }
  const int ea = cfg.electron_arm;
  const bool e_right = (ea == 1 || ea == 3 || ea == 7);
  const ArmFull& r1 = e_right ? recon.e : recon.p;
  const ArmFull& r2 = e_right ? recon.p : recon.e;
  const ArmFP& f1 = e_right ? main.FP_e : main.FP_p;
  const ArmFP& f2 = e_right ? main.FP_p : main.FP_e;
  const ArmFull& o1 = e_right ? orig.e : orig.p;
  const ArmFull& o2 = e_right ? orig.p : orig.e;
  const double s1 = e_right ? cfg.spec_e.sin_th : cfg.spec_p.sin_th;
  const double s2 = e_right ? cfg.spec_p.sin_th : cfg.spec_e.sin_th;
  ntu[1] = r1.delta; ntu[2] = r1.yptar; ntu[3] = r1.xptar; ntu[4] = r1.z;
  ntu[5] = f1.x; ntu[6] = f1.dx; ntu[7] = f1.y; ntu[8] = f1.dy;
  ntu[9] = o1.delta; ntu[10] = o1.yptar; ntu[11] = o1.xptar; ntu[12] = main.target.z * s1;
  ntu[13] = r2.delta; ntu[14] = r2.yptar; ntu[15] = r2.xptar; ntu[16] = r2.z;
  ntu[17] = f2.x; ntu[18] = f2.dx; ntu[19] = f2.y; ntu[20] = f2.dy;
  ntu[21] = o2.delta; ntu[22] = o2.yptar; ntu[23] = o2.xptar; ntu[24] = -main.target.z * s2;
  ntu[25] = recon.q / 1000.; ntu[26] = recon.nu / 1000.; ntu[27] = recon.Q2 / 1.e6; ntu[28] = recon.W / 1000.;
  ntu[29] = recon.epsilon; ntu[30] = recon.Em / 1000.; ntu[31] = recon.Pm / 1000.; ntu[32] = recon.theta_pq;
  ntu[33] = recon.phi_pq;
  int ncol;
  if (cfg.doing_pion || cfg.doing_kaon || cfg.doing_delta) {
    ntu[34] = s.ntup.mm / 1000.; ntu[35] = s.ntup.mmA / 1000.; ntu[36] = recon.p.P / 1000.; ntu[37] = s.ntup.t / 1.e6;
    ntu[38] = recon.PmPar / 1000.; ntu[39] = recon.PmPer / 1000.; ntu[40] = recon.PmOop / 1000.;
    ntu[41] = -main.target.rastery; ntu[42] = s.ntup.radphot / 1000.;
    const double pfer = s.pfer, pferx = s.pferx, pfery = s.pfery, pferz = s.pferz;     // zero for hydrogen, event.f:330-335
    double dummy = pferx * vertex.uq.x + pfery * vertex.uq.y + pferz * vertex.uq.z;
    if (dummy == 0) dummy = 1.e-20;
    ntu[43] = pfer / 1000. * std::fabs(dummy) / dummy;
    ntu[44] = main.sigcc; ntu[45] = s.ntup.sigcm; ntu[46] = main.weight; ntu[47] = s.trk.decdist;
    ntu[48] = std::sqrt(s.trk.Mh2_final); ntu[49] = pfer / 1000. * dummy; ntu[50] = vertex.Q2 / 1.e6;
    ntu[51] = main.W / 1.e3; ntu[52] = main.t / 1.e6; ntu[53] = main.phi_pq;
    ncol = 53;
    if (cfg.using_tgt_field) {                                 // results_write.f:154-166
      ntu[54] = recon.theta_tarq; ntu[55] = recon.phi_targ; ntu[56] = recon.beta; ntu[57] = recon.phi_s; ntu[58] = recon.phi_c;
      ntu[59] = main.beta; ntu[60] = vertex.phi_s; ntu[61] = vertex.phi_c;
      ncol = 61;
      if (cfg.doing_kaon) { ntu[62] = s.ntup.sigcm1; ntu[63] = s.ntup.sigcm2; ncol = 63; }
    } else if (cfg.doing_kaon) { ntu[54] = s.ntup.sigcm1; ntu[55] = s.ntup.sigcm2; ncol = 55; }
    if (cfg.doing_pizero) {                                    // results_write.f:167-180 (no target field)
      ntu[54] = s.ntup.xcal_gamma1; ntu[55] = s.ntup.ycal_gamma1;
      for (int k = 0; k < 4; ++k) ntu[56 + k] = s.ntup.gamma1[k];
      ntu[60] = s.ntup.xcal_gamma2; ntu[61] = s.ntup.ycal_gamma2;
      for (int k = 0; k < 4; ++k) ntu[62 + k] = s.ntup.gamma2[k];
      ncol = 65;
    }
  } else if (cfg.doing_semi || cfg.doing_rho) {      // results_write.f:187-213
    ntu[34] = s.ntup.mm / 1000.; ntu[35] = recon.p.P / 1000.; ntu[36] = s.ntup.t / 1.e6;
    ntu[37] = -main.target.rastery; ntu[38] = s.ntup.radphot / 1000.; ntu[39] = main.sigcc; ntu[40] = main.sigcent;
    ntu[41] = main.weight; ntu[42] = s.trk.decdist; ntu[43] = std::sqrt(s.trk.Mh2_final);
    ntu[44] = recon.zhad; ntu[45] = vertex.zhad; ntu[46] = recon.pt2 / 1.e06; ntu[47] = vertex.pt2 / 1.e06;
    ntu[48] = recon.xbj; ntu[49] = vertex.xbj; ntu[50] = std::acos(vertex.uq.z); ntu[51] = s.ntup.sigcm;
    ntu[52] = main.davejac; ntu[53] = main.johnjac;
    const double dummy = s.pferx * vertex.uq.x + s.pfery * vertex.uq.y + s.pferz * vertex.uq.z;
    ntu[54] = s.pfer / 1000. * std::fabs(dummy) / dummy;      // NaN for hydrogen (0/0), as in the reference
    ntu[55] = s.ntup.xfermi; ntu[56] = main.phi_pq;
    ncol = 56;
    if (cfg.using_tgt_field) {                                 // results_write.f:212-225
      ntu[57] = recon.theta_tarq; ntu[58] = recon.phi_targ; ntu[59] = recon.beta; ntu[60] = recon.phi_s; ntu[61] = recon.phi_c;
      ntu[62] = main.beta; ntu[63] = vertex.phi_s; ntu[64] = vertex.phi_c;
      ncol = 64;
      if (cfg.doing_rho) { ntu[65] = s.ntup.rhomass; ntu[66] = s.ntup.rhotheta; ncol = 67; }      // 67 tags, the last never filled
    } else if (cfg.doing_rho) {                                // results_write.f:226-230
      ntu[57] = s.ntup.rhomass; ntu[58] = s.ntup.rhotheta; ntu[59] = s.ntup.mmA / 1000.;
      ncol = 59;
    }
  } else if (eep) {
    ntu[34] = corrsing / 1000.; ntu[35] = Pm_Heepx / 1000.; ntu[36] = Pm_Heepy / 1000.; ntu[37] = Pm_Heepz / 1000.;
    ntu[38] = recon.PmPar / 1000.; ntu[39] = recon.PmPer / 1000.; ntu[40] = recon.PmOop / 1000.;
    ntu[41] = -main.target.rastery; ntu[42] = s.ntup.radphot / 1000.; ntu[43] = main.sigcc; ntu[44] = main.weight;
    ntu[45] = recon.e.theta; ntu[46] = recon.p.theta;
    ncol = 46;
  } else {
    throw std::runtime_error("oracle: ntuple layout of this reaction not restated");
  }
  for (int i = 0; i < ncol; ++i) out[i] = ntu[i + 1];
  return ncol;
}

// tries [first, first+n) of stream `seed`; rec/status may be null
void run_range(const simc_run_config& cfg, const ArmOptics* oe, const ArmOptics* op, int64_t first, int64_t n,
               uint64_t seed, simc_accum* acc, double* rec, int32_t* status, int64_t rec_stride, int64_t rec_off,
               RanluxState* ranlux, const SfTable* sf, double* ntu_rows, int64_t* n_rows, int* n_cols,
               int64_t* try_of_row, const PfermiTable* pfermi, const Cteq5Table* pdf, const TheoryTable* theory, const MaidTable* maid, const FdssTable* fdss) {
  TrgField field_run;
  const TrgField* field = nullptr;
  if (cfg.using_tgt_field) {                    // simc.f:120-156: trgInit with the angles of this run's spectrometers
    const TrgField* map = field_map();
    if (!map->set) throw std::runtime_error("oracle: field map not set");
    field_run = *map;
    double ae, ap;
    field_arm_angles(cfg.targ_Bangle, cfg.targ_Bphi, cfg.spec_e.theta, cfg.spec_e.phi, cfg.spec_p.theta, cfg.spec_p.phi, ae, ap);
    const double pi180 = 3.141592653 / 180.;
    field_run.B_stheta[0] = std::sin(ae * pi180); field_run.B_ctheta[0] = std::cos(ae * pi180);
    field_run.B_stheta[1] = std::sin(ap * pi180); field_run.B_ctheta[1] = std::cos(ap * pi180);
    field = &field_run;
  }
  for (int64_t i = 0; i < n; ++i) {
    Rng rng;
    if (ranlux) { rng.mode = Rng::RANLUX; rng.rl = ranlux; rng.draw = 0; }   // the reference's sequential stream
    else rng.seed_philox(seed, (uint64_t)(first + i));
    Sim s;
    s.cfg = &cfg; s.optics_e = oe; s.optics_p = op; s.rng = &rng; s.sf = sf; s.pfermi = pfermi; s.pdf = pdf; s.theory = theory; s.maid = maid; s.fdss = fdss;
    s.field = field;
    EventMain main;
    Event vertex, orig, recon;
    const TryResult r = one_try(s, main, vertex, orig, recon);
    if (acc) accumulate(s, r, main, vertex, orig, recon, *acc);
    if (rec) fill_record(s, r, main, vertex, orig, recon, rec, rec_stride, rec_off + i);
    if (status) status[rec_off + i] = r.stage;
    if (ntu_rows && r.success) {
      const int nc = fill_ntuple(s, main, vertex, orig, recon, ntu_rows + (*n_rows) * SIMC_NTUPLE_MAXCOL);
      if (n_cols) *n_cols = nc;
      if (try_of_row) try_of_row[*n_rows] = first + i;
      ++*n_rows;
    }
  }
}

}  // namespace simc_oracle
