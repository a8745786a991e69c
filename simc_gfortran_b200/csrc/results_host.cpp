// End-of-run host code of the C ABI: normalisation (simc.f:94-101, 366-432) and the ntuple file writer
// (NtupleInit.f, results_write.f:264-266).  No device code.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include "../../include/simc_b200.h"

namespace {

double fixed_value(const simc_fixed128& f) {
  const long double v = (long double)f.hi * 18446744073709551616.0L + (long double)f.lo;
  return (double)std::ldexp(v, f.qexp);
}

// NtupleInit.f:33-343 (no target field, no pi0, no rho)
const char* const kCommon[33] = {"hsdelta", "hsyptar", "hsxptar", "hsytar", "hsxfp", "hsxpfp", "hsyfp", "hsypfp", "hsdeltai",
                                 "hsyptari", "hsxptari", "hsytari", "ssdelta", "ssyptar", "ssxptar", "ssytar", "ssxfp",
                                 "ssxpfp", "ssyfp", "ssypfp", "ssdeltai", "ssyptari", "ssxptari", "ssytari", "q", "nu",
                                 "Q2", "W", "epsilon", "Em", "Pm", "thetapq", "phipq"};
const char* const kMeson[22] = {"missmass", "mmnuc", "phad", "t", "pmpar", "pmper", "pmoop", "fry", "radphot", "pfermi",
                                "siglab", "sigcm", "Weight", "decdist", "Mhadron", "pdotqhat", "Q2i", "Wi", "ti", "phipqi",
                                "saghai", "factor"};
const char* const kSemi[26] = {"missmass", "ppi", "t", "fry", "radphot", "siglab", "sigcent", "Weight", "decdist", "Mhadron",
                               "z", "zi", "pt2", "pt2i", "xbj", "xbji", "thqi", "sighad", "jacobian", "centjac", "pfermi",
                               "xfermi", "phipqi", "Mrho", "Thrho", "mmnuc"};     // the last three: doing_rho (NtupleInit.f:257-264)
const char* const kEep[13] = {"corrsing", "Pmx", "Pmy", "Pmz", "PmPar", "PmPer", "PmOop", "fry", "radphot", "sigcc", "Weight",
                              "theta_e", "theta_p"};

}  // namespace

struct simc_ntuple_file {
  FILE* f;
  int n_cols;
};

extern "C" {

int simc_b200_normalise(const simc_run_config* cfg, const simc_accum* acc, int32_t ngen, double charge_mC, simc_results* out) {
  if (!cfg || !acc || !out) return SIMC_ERR_ARG;
  std::memset(out, 0, sizeof(*out));
  const simc_target& targ = cfg->targ;
  // simc.f:94-101 (targ%thick is in g/cm^2 at this point, dbase.f:496)
  const double targetfac = targ.mass_amu / 3.75914e+6 / (targ.abundancy / 100.) * std::fabs(std::cos(targ.angle)) / (targ.thick * 1000.);
  out->luminosity = charge_mC / targetfac;
  const simc_gen_limits& gen = cfg->gen;
  // simc.f:346-350: with ngen < 0 every try counts as an event (nevent = ntried when the loop ends), with
  // ngen > 0 only the successes do; simc.f:368: normfac = luminosity/ntried*nevent
  const double nevent = ngen < 0 ? (double)acc->ntried : (double)acc->nsuccess;
  out->nevent = (int64_t)nevent;
  double normfac = acc->ntried > 0 ? out->luminosity / (double)acc->ntried * nevent : 0.0;
  const double domega_e = (gen.e.yptar.max - gen.e.yptar.min) * (gen.e.xptar.max - gen.e.xptar.min);
  const double domega_p = cfg->doing_rho ? 4. * 3.141592653589793 : (gen.p.yptar.max - gen.p.yptar.min) * (gen.p.xptar.max - gen.p.xptar.min);
  double genvol = domega_e;
  if (cfg->doing_deuterium || cfg->doing_heavy || cfg->doing_pion || cfg->doing_kaon || cfg->doing_delta || cfg->doing_rho ||
      cfg->doing_semi)
    genvol = genvol * domega_p * (gen.e.E.max - gen.e.E.min);
  if (cfg->doing_heavy || cfg->doing_semi) genvol = genvol * (gen.p.E.max - gen.p.E.min);
  normfac = normfac * genvol;
  if (cfg->doing_phsp) normfac = 1.0;
  out->genvol = genvol;
  out->normfac = normfac;
  out->yield = fixed_value(acc->wtcontribute) * normfac;
  out->central_sigcc_ave = nevent > 0 ? fixed_value(acc->sum_sigcc) / nevent : 0.0;      // simc.f:959
  if (acc->npasscuts > 1) {
    const double tmpnum = (double)acc->npasscuts;
    for (int k = 0; k < 8; ++k) {
      const double s1 = fixed_value(acc->sumerr[k]), s2 = fixed_value(acc->sumerr2[k]);
      out->aveerr[k] = s1 / tmpnum;
      out->resol[k] = std::sqrt(std::fmax(0., (s2 / tmpnum) - (s1 / tmpnum) * (s1 / tmpnum)));
    }
  }
  return SIMC_OK;
}

int simc_b200_accum_merge(simc_accum* into, const simc_accum* from) {
  if (!into || !from) return SIMC_ERR_ARG;
  typedef __int128 i128;
  auto addf = [](simc_fixed128& a, const simc_fixed128& b) -> bool {
    if (a.qexp != b.qexp) {
      const bool a0 = a.lo == 0 && a.hi == 0, b0 = b.lo == 0 && b.hi == 0;
      if (b0) return true;                               // nothing to add
      if (!a0) return false;                             // two non-empty sums on different quanta cannot be added exactly
      a.qexp = b.qexp;
    }
    const i128 v = (((i128)a.hi << 64) | (i128)a.lo) + (((i128)b.hi << 64) | (i128)b.lo);
    a.lo = (uint64_t)v; a.hi = (int64_t)(v >> 64);
    return true;
  };
  bool ok = true;
  into->ntried += from->ntried; into->nsuccess += from->nsuccess; into->ncontribute += from->ncontribute;
  into->npasscuts += from->npasscuts; into->ncontribute_no_rad_proton += from->ncontribute_no_rad_proton;
  into->unsupported += from->unsupported; into->nonfinite += from->nonfinite;
  ok &= addf(into->wtcontribute, from->wtcontribute);
  ok &= addf(into->sum_sigcc, from->sum_sigcc);
  for (int i = 0; i < 8; ++i) { ok &= addf(into->sumerr[i], from->sumerr[i]); ok &= addf(into->sumerr2[i], from->sumerr2[i]); }
  for (int k = 0; k < 6; ++k) for (int b = 0; b < SIMC_NHIST; ++b) ok &= addf(into->hist_w[k][b], from->hist_w[k][b]);
  for (int s = 0; s < 3; ++s) for (int k = 0; k < SIMC_H_PER_SET; ++k) for (int b = 0; b < SIMC_NHIST; ++b)
    into->hist_n[s][k][b] += from->hist_n[s][k][b];
  for (int i = 0; i < 32; ++i) {
    if (from->contrib[i].lo < into->contrib[i].lo) into->contrib[i].lo = from->contrib[i].lo;
    if (from->contrib[i].hi > into->contrib[i].hi) into->contrib[i].hi = from->contrib[i].hi;
  }
  for (int i = 0; i < 8; ++i) {
    if (from->slop[i].lo < into->slop[i].lo) into->slop[i].lo = from->slop[i].lo;
    if (from->slop[i].hi > into->slop[i].hi) into->slop[i].hi = from->slop[i].hi;
  }
  for (int w = 0; w < 2; ++w) for (int i = 0; i < SIMC_NSTOP; ++i) into->stop[w][i] += from->stop[w][i];
  for (int w = 0; w < 2; ++w) for (int i = 0; i < 48; ++i) into->transp_calls[w][i] += from->transp_calls[w][i];
  return ok ? SIMC_OK : SIMC_ERR_ARG;
}

int simc_b200_ntuple_tags(const simc_run_config* cfg, char (*tags)[17], int max_tags) {
  if (!cfg || !tags) return SIMC_ERR_ARG;
  const char* const* tail;
  int n_tail;
  if (cfg->doing_pion || cfg->doing_kaon || cfg->doing_delta) { tail = kMeson; n_tail = cfg->doing_kaon ? 22 : 20; }
  else if (cfg->doing_semi || cfg->doing_rho) { tail = kSemi; n_tail = cfg->doing_rho ? 26 : 23; }
  else if (cfg->doing_hyd_elast || cfg->doing_deuterium || cfg->doing_heavy) { tail = kEep; n_tail = 13; }
  else return SIMC_ERR_ARG;
  // polarised target: eight tags behind "phipqi" (the 20th meson / 23rd semi-inclusive tag), NtupleInit.f:142-160, 239-256
  static const char* const kPol[8] = {"th_tarq", "phitarq", "beta", "phis", "phic", "betai", "phisi", "phici"};
  const bool pol = cfg->using_tgt_field && tail != kEep;
  const int pol_at = tail == kMeson ? 20 : 23;
  // pi0 -> gamma gamma: twelve tags behind the meson ones (NtupleInit.f:163-187)
  static const char* const kPi0[12] = {"xcal_gamma1", "ycal_gamma1", "Egamma1", "Pgamma1x", "Pgamma1y", "Pgamma1z",
                                       "xcal_gamma2", "ycal_gamma2", "Egamma2", "Pgamma2x", "Pgamma2y", "Pgamma2z"};
  const bool pi0 = cfg->doing_pizero && tail == kMeson;
  const int n = 33 + n_tail + (pol ? 8 : 0) + (pi0 ? 12 : 0);
  if (n > max_tags) return SIMC_ERR_ARG;
  for (int i = 0; i < n; ++i) {
    std::memset(tags[i], 0, 17);
    const char* t;
    if (i < 33) t = kCommon[i];
    else {
      const int j = i - 33;
      if (pi0 && j >= n_tail + (pol ? 8 : 0)) t = kPi0[j - n_tail - (pol ? 8 : 0)];
      else if (!pol || j < pol_at) t = tail[j];
      else if (j < pol_at + 8) t = kPol[j - pol_at];
      else t = tail[j - 8];
    }
    std::strncpy(tags[i], t, 16);
  }
  return n;
}

static bool put_record(FILE* f, const void* p, int32_t len) {
  return std::fwrite(&len, 4, 1, f) == 1 && std::fwrite(p, 1, (size_t)len, f) == (size_t)len && std::fwrite(&len, 4, 1, f) == 1;
}

int simc_b200_ntuple_open(const simc_run_config* cfg, const char* path, simc_ntuple_file** out) {
  if (!cfg || !path || !out) return SIMC_ERR_ARG;
  *out = nullptr;
  char tags[80][17];
  const int n = simc_b200_ntuple_tags(cfg, tags, 80);
  if (n < 0) return n;
  FILE* f = std::fopen(path, "wb");
  if (!f) return SIMC_ERR_IO;
  const int32_t size = n;
  bool ok = put_record(f, &size, 4);                     // write(NtupleIO) NtupleSize
  for (int i = 0; i < n && ok; ++i) {
    char rec[16];
    std::memset(rec, ' ', 16);                           // character*16, blank padded
    std::memcpy(rec, tags[i], std::strlen(tags[i]));
    ok = put_record(f, rec, 16);
  }
  if (!ok) { std::fclose(f); return SIMC_ERR_IO; }
  simc_ntuple_file* h = new (std::nothrow) simc_ntuple_file{f, n};
  if (!h) { std::fclose(f); return SIMC_ERR_IO; }
  *out = h;
  return SIMC_OK;
}

int simc_b200_ntuple_append(simc_ntuple_file* h, const double* rows, int64_t n_rows) {
  if (!h || (n_rows > 0 && !rows)) return SIMC_ERR_ARG;
  for (int64_t r = 0; r < n_rows; ++r)
    for (int k = 0; k < h->n_cols; ++k)                  // do i=1,NtupleSize: write(NtupleIO) ntu(i)
      if (!put_record(h->f, &rows[r * SIMC_NTUPLE_MAXCOL + k], 8)) return SIMC_ERR_IO;
  return SIMC_OK;
}

int simc_b200_ntuple_close(simc_ntuple_file* h) {
  if (!h) return SIMC_ERR_ARG;
  const int rc = std::fclose(h->f) == 0 ? SIMC_OK : SIMC_ERR_IO;
  delete h;
  return rc;
}

}  // extern "C"
