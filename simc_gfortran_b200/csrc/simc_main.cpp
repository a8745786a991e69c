// simc_b200: stand-alone driver over the C ABI -- what `program simc` (simc.f) does around its event loop,
// for machines without the Fortran driver.  Reads a CTP deck, loads the optics and physics tables from a copy of
// the reference's working directory (hms/forward_cosy.dat, benharsf_12.dat, deut.dat, ...), runs the loop on one
// GPU, normalises (simc.f:366-432) and writes
//   <out>.hist  subroutine report of the reference (simc.f:644-1139), same formats
//   <out>.gen   the acceptance histograms (simc.f:539-612)
//   <out>.geni  the STOP counters of the spectrometers (simc.f:446-537)
//   <out>.bin   ntuple in the reference's unformatted layout (with --ntuple 1; the deck's Nntu is not read)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>
#include "../../include/simc_b200.h"

namespace {
struct ArmFiles { const char* fwd; const char* rec; };
ArmFiles arm_files(int arm) {
  switch (arm) {
    case SIMC_ARM_HMS: return {"hms/forward_cosy.dat", "hms/recon_cosy.dat"};
    case SIMC_ARM_SOS: return {"sos/forward_cosy.dat", "sos/recon_cosy.dat"};
    case SIMC_ARM_HRSR: return {"hrsr/hrs_forward_cosy.dat", "hrsr/hrs_recon_cosy.dat"};
    case SIMC_ARM_HRSL: return {"hrsl/hrs_forward_cosy.dat", "hrsl/hrs_recon_cosy.dat"};
    case SIMC_ARM_SHMS: return {"shms/shms_forward.dat", "shms/shms_recon.dat"};
  }
  return {nullptr, nullptr};
}
double fx(const simc_fixed128& f) {
  const long double v = (long double)f.hi * 18446744073709551616.0L + (long double)f.lo;
  return (double)std::ldexp(v, f.qexp);
}
int die(simc_handle* h, const char* what, int rc) {
  std::fprintf(stderr, "simc_b200: %s failed (%d): %s\n", what, rc, simc_b200_last_error(h));
  return 1;
}
}  // namespace

int main(int argc, char** argv) {
  std::string deck, data = ".", out = "simc_b200_run";
  long long seed = 1, chunk = 1 << 22;
  int device = 0, want_ntuple = 0;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto next = [&]() { return i + 1 < argc ? std::string(argv[++i]) : std::string(); };
    if (a == "--data") data = next();
    else if (a == "--out") out = next();
    else if (a == "--seed") seed = std::atoll(next().c_str());
    else if (a == "--device") device = std::atoi(next().c_str());
    else if (a == "--ntuple") want_ntuple = std::atoi(next().c_str());
    else if (a == "--chunk") chunk = std::atoll(next().c_str());
    else if (a[0] != '-') deck = a;
    else { std::fprintf(stderr, "usage: simc_b200 deck.inp [--data DIR] [--out PREFIX] [--seed N] [--device D] [--ntuple 0|1] [--chunk TRIES]\n"); return 2; }
  }
  if (deck.empty()) { std::fprintf(stderr, "simc_b200: no deck given\n"); return 2; }
  simc_run_config cfg;
  int32_t ngen = 0;
  double charge = 0;
  char err[512] = "";
  const std::string deck_dir = deck.find('/') == std::string::npos ? "." : deck.substr(0, deck.find_last_of('/'));
  int rc = simc_b200_config_from_deck_data(deck.c_str(), deck_dir.c_str(), data.c_str(), &cfg, &ngen, &charge, err, sizeof err);
  if (rc) { std::fprintf(stderr, "simc_b200: %s\n", err); return 1; }
  simc_handle* h = nullptr;
  rc = simc_b200_create(&cfg, device, &h);
  if (rc) return die(nullptr, "simc_b200_create", rc);
  for (int arm : {cfg.electron_arm, cfg.hadron_arm}) {
    if (arm == cfg.hadron_arm && (arm == SIMC_ARM_CALO_RIGHT || arm == SIMC_ARM_CALO_LEFT)) continue;     // a calorimeter has no maps
    const ArmFiles f = arm_files(arm);
    if (!f.fwd) { std::fprintf(stderr, "simc_b200: spectrometer %d has no single-arm Monte Carlo here\n", arm); return 1; }
    rc = simc_b200_load_optics(h, arm, (data + "/" + f.fwd).c_str(), (data + "/" + f.rec).c_str());
    if (rc) return die(h, "simc_b200_load_optics", rc);
  }
  // tables by reaction: the files dbase.f / theory_init / semi_physics.f / physics_pion.f open
  const int nA = (int)std::lround(cfg.targ.A);
  const bool he_meson = cfg.doing_hepi || cfg.doing_hekaon;
  if ((cfg.doing_heavy && cfg.use_benhar_sf) || he_meson) {
    const char* f = nA == 3 ? "benharsf_3mod.dat" : nA == 4 ? "benharsf_4.dat" : nA == 56 ? "benharsf_56.dat" : nA == 197 ? "benharsf_197.dat" : "benharsf_12.dat";
    const int proton = std::fabs(cfg.targ.Mtar_struck - 938.27231) <= 1.e-6 || std::fabs(cfg.targ.Mtar_struck - 939.56563) > 1.e-6;
    if ((rc = simc_b200_load_sf_file(h, (data + "/" + f).c_str(), proton))) return die(h, "simc_b200_load_sf_file", rc);
  }
  if (cfg.doing_deuterium || (cfg.doing_heavy && !cfg.use_benhar_sf)) {
    const char* f = nA == 2 ? "h2.theory" : nA == 56 ? "fe56.theory" : nA == 197 ? "au197.theory" : "c12.theory";
    if ((rc = simc_b200_load_theory_file(h, (data + "/" + f).c_str()))) return die(h, "simc_b200_load_theory_file", rc);
  }
  if (cfg.doing_deutpi || cfg.doing_deutkaon || cfg.doing_deutsemi || he_meson) {
    const char* f = !he_meson ? "deut.dat" : nA == 3 ? "he3.dat" : nA == 4 ? "he4.dat" : "c12.dat";
    if ((rc = simc_b200_load_pfermi_file(h, (data + "/" + f).c_str()))) return die(h, "simc_b200_load_pfermi_file", rc);
  }
  if (cfg.doing_semi) {
    if ((rc = simc_b200_load_cteq5_file(h, (data + "/cteq5/cteq5m.tbl").c_str()))) return die(h, "simc_b200_load_cteq5_file", rc);
    if (cfg.doing_semika && (rc = simc_b200_load_fdss_file(h, (data + "/fdss/KANLO.GRID").c_str()))) return die(h, "simc_b200_load_fdss_file", rc);
  }
  if (cfg.doing_pion) {       // optional: without it events below W = 2 GeV are counted as unsupported
    const bool piminus = cfg.which_pion == 1 || cfg.which_pion == 11 || cfg.which_pion == 3;
    const std::string f = data + (piminus ? "/maidpimp.dat" : "/maidpipn.dat");
    if (FILE* t = std::fopen(f.c_str(), "r")) { std::fclose(t); if ((rc = simc_b200_load_maid_file(h, piminus ? 4 : 3, f.c_str()))) return die(h, "simc_b200_load_maid_file", rc); }
  }

  if (cfg.using_tgt_field) {  // trgInit (simc.f:154): tgt_field_file of the deck is trg_field_map.dat in every shipped deck
    if ((rc = simc_b200_load_field_file(h, (data + "/trg_field_map.dat").c_str()))) return die(h, "simc_b200_load_field_file", rc);
  }
  if (cfg.doing_kaon) {       // optional: the Saghai model only fills ntuple column 54 (zero without its tables)
    if (FILE* t = std::fopen((data + "/saghai_proton.dat").c_str(), "r")) {
      std::fclose(t);
      if ((rc = simc_b200_load_saghai_files(h, data.c_str()))) return die(h, "simc_b200_load_saghai_files", rc);
    }
  }

  const std::time_t t_begin = std::time(nullptr);
  simc_accum acc;
  if ((rc = simc_b200_accum_clear(h, &acc))) return die(h, "simc_b200_accum_clear", rc);
  simc_ntuple_file* nt = nullptr;
  if (want_ntuple && (rc = simc_b200_ntuple_open(&cfg, (out + ".bin").c_str(), &nt))) return die(h, "simc_b200_ntuple_open", rc);
  std::vector<double> rows;
  if (nt) rows.resize((size_t)chunk * SIMC_NTUPLE_MAXCOL);
  // ngen < 0: that many tries; ngen > 0: until that many successes (simc.f:346-350), try by try reproducible
  const long long want_tries = ngen < 0 ? -(long long)ngen : -1, want_success = ngen > 0 ? ngen : -1;
  long long first = 0;
  while (true) {
    long long n = chunk;
    if (want_tries >= 0) n = std::min(n, want_tries - first);
    if (n <= 0) break;
    if (want_success >= 0) {
      // size the chunk from the acceptance seen so far so that the last one does not overshoot by much
      const double accf = acc.ntried > 0 ? std::max(1e-6, (double)acc.nsuccess / (double)acc.ntried) : 0.05;
      n = std::min<long long>(chunk, std::max<long long>(1024, (long long)((want_success - acc.nsuccess) / accf * 1.05)));
    }
    simc_accum before = acc;
    if ((rc = simc_b200_run(h, first, n, (uint64_t)seed, &acc))) return die(h, "simc_b200_run", rc);
    if (want_success >= 0 && acc.nsuccess >= want_success) {
      // the reference stops at the try that yields the ngen-th success (simc.f:346-350): bisect the range for the
      // smallest number of tries with that many successes (try t is reproducible), so that no failed try behind
      // the last success is counted in ntried, the geni histograms or the STOP counters
      long long lo = 0, hi = n;
      while (hi - lo > 1) {
        const long long mid = (lo + hi) / 2;
        simc_accum t = before;
        if ((rc = simc_b200_run(h, first, mid, (uint64_t)seed, &t))) return die(h, "simc_b200_run", rc);
        if (t.nsuccess >= want_success) hi = mid; else lo = mid;
      }
      acc = before;
      n = hi;
      if ((rc = simc_b200_run(h, first, n, (uint64_t)seed, &acc))) return die(h, "simc_b200_run", rc);
    }
    if (nt) {
      int32_t n_cols = 0;
      int64_t n_rows = 0;
      if ((rc = simc_b200_ntuple_batch(h, first, n, (uint64_t)seed, rows.data(), &n_cols, &n_rows, nullptr))) return die(h, "simc_b200_ntuple_batch", rc);
      if ((rc = simc_b200_ntuple_append(nt, rows.data(), n_rows))) return die(h, "simc_b200_ntuple_append", rc);
    }
    first += n;
    if (want_success >= 0 && acc.nsuccess >= want_success) break;
  }
  if (nt) simc_b200_ntuple_close(nt);
  simc_results res;
  simc_b200_normalise(&cfg, &acc, ngen, charge, &res);

  // the reference's three text files (simc.f:446-1139)
  simc_report_info info;
  if ((rc = simc_b200_report_info_from_deck(deck.c_str(), deck_dir.c_str(), data.c_str(), &info, err, sizeof err))) {
    std::fprintf(stderr, "simc_b200: %s\n", err); return 1;
  }
  info.random_seed = (int32_t)seed;
  simc_central central;
  if ((rc = simc_b200_central_event(h, &cfg, &info, &central))) return die(h, "simc_b200_central_event", rc);
  const std::time_t t_end = std::time(nullptr);
  char ts1[64], ts2[64];
  std::snprintf(ts1, sizeof ts1, "%s", std::ctime(&t_begin));
  std::snprintf(ts2, sizeof ts2, "%s", std::ctime(&t_end));
  if ((rc = simc_b200_write_geni((out + ".geni").c_str(), &cfg, &acc)) || (rc = simc_b200_write_gen((out + ".gen").c_str(), &cfg, &acc)) ||
      (rc = simc_b200_write_hist((out + ".hist").c_str(), &cfg, &info, &central, &acc, &res, ts1, ts2))) {
    std::fprintf(stderr, "simc_b200: cannot write %s.{geni,gen,hist}\n", out.c_str()); return 1;
  }
  std::printf("simc_b200: %lld tries, %lld successes, normalised yield %.6g for %.4g mC -> %s.hist%s\n", (long long)acc.ntried,
              (long long)acc.nsuccess, res.yield, charge, out.c_str(), want_ntuple ? " + .bin" : "");
  simc_b200_destroy(h);
  return 0;
}
