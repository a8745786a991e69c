// ORACLE -- TEST INFRASTRUCTURE ONLY (see event.hpp).
// Loop bookkeeping of simc.f:229-350 (pass_cuts, inc, counters, sumerr, limits_update) on top
// of one_try(), accumulating into the same exact integer representation the B200 path uses
// (include/simc_b200.h: simc_accum), so that the two can be compared bit for bit.
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>
#include "event.hpp"

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
      main.thetacm, main.phicm, s.ntup.sigcm, main.davejac, s.ntup.survivalprob, s.ntup.mm, main.wcm, main.t};
  for (int k = 0; k < SIMC_EVENT_NREC; ++k) rec[k * n + i] = v[k];
}

// tries [first, first+n) of stream `seed`; rec/status may be null
void run_range(const simc_run_config& cfg, const ArmOptics* oe, const ArmOptics* op, int64_t first, int64_t n,
               uint64_t seed, simc_accum* acc, double* rec, int32_t* status, int64_t rec_stride, int64_t rec_off,
               RanluxState* ranlux, const SfTable* sf) {
  for (int64_t i = 0; i < n; ++i) {
    Rng rng;
    if (ranlux) { rng.mode = Rng::RANLUX; rng.rl = ranlux; rng.draw = 0; }   // the reference's sequential stream
    else rng.seed_philox(seed, (uint64_t)(first + i));
    Sim s;
    s.cfg = &cfg; s.optics_e = oe; s.optics_p = op; s.rng = &rng; s.sf = sf;
    EventMain main;
    Event vertex, orig, recon;
    const TryResult r = one_try(s, main, vertex, orig, recon);
    if (acc) accumulate(s, r, main, vertex, orig, recon, *acc);
    if (rec) fill_record(s, r, main, vertex, orig, recon, rec, rec_stride, rec_off + i);
    if (status) status[rec_off + i] = r.stage;
  }
}

}  // namespace simc_oracle
