// ORACLE -- TEST INFRASTRUCTURE ONLY.  C entry points for tests/ (ctypes), smoke() and the
// cpu_baseline leg of bench.py.  Never linked into, loaded by, or called from libsimc_b200.so.
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>
#include "event.hpp"
#include "field.hpp"

using namespace simc_oracle;

static std::map<int, ArmOptics> g_optics;
static SfTable g_sf;
static PfermiTable g_pfermi;
static Cteq5Table g_pdf;
static TheoryTable g_theory;
static MaidTable g_maid;
static FdssTable g_fdss;
static SaghaiTable g_saghai;
namespace simc_oracle { const SaghaiTable* saghai_tables() { return &g_saghai; } }
static std::string g_err;

static simc_oracle::TrgField g_field;          // COMMON /trgFieldStrength/ as oracle_set_field_map left it
namespace simc_oracle { const TrgField* field_map() { return &g_field; } }

extern "C" {

const char* oracle_last_error() { return g_err.c_str(); }
void oracle_set_error(const char* m) { g_err = m ? m : ""; }

int oracle_load_optics(int arm, const char* fwd_path, const char* rec_path) {
  try {
    ArmOptics o;
    o.fwd.load(fwd_path);
    o.rec.load(rec_path);
    g_optics[arm] = std::move(o);
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

// Table export: sizes first (pass null arrays), then data.  Layout = simc_b200_set_optics.
int oracle_optics_sizes(int arm, int* n_classes, int* n_fwd_terms, int* n_rec) {
  auto it = g_optics.find(arm);
  if (it == g_optics.end()) { g_err = "optics not loaded"; return -1; }
  *n_classes = it->second.fwd.n_classes();
  int n = 0;
  for (auto& c : it->second.fwd.cls) n += c.n_terms;
  *n_fwd_terms = n;
  *n_rec = it->second.rec.n_terms;
  return 0;
}
int oracle_optics_export(int arm, int32_t* class_start, double* fwd_coeff, int8_t* fwd_expon, double* length_cm,
                         int8_t* adrift, double* driftdist, double* rec_coeff, int8_t* rec_expon) {
  auto it = g_optics.find(arm);
  if (it == g_optics.end()) { g_err = "optics not loaded"; return -1; }
  const ArmOptics& o = it->second;
  int pos = 0, k = 0;
  for (auto& c : o.fwd.cls) {
    class_start[k] = pos;
    length_cm[k] = c.length;
    adrift[k] = c.adrift;
    driftdist[k] = c.driftdist;
    std::memcpy(fwd_coeff + 5 * pos, c.coeff.data(), sizeof(double) * 5 * c.n_terms);
    std::memcpy(fwd_expon + 5 * pos, c.expon.data(), 5 * c.n_terms);
    pos += c.n_terms;
    ++k;
  }
  class_start[k] = pos;
  std::memcpy(rec_coeff, o.rec.coeff.data(), sizeof(double) * 4 * o.rec.n_terms);
  std::memcpy(rec_expon, o.rec.expon.data(), 5 * o.rec.n_terms);
  return 0;
}
int oracle_set_optics(int arm, int n_classes, const int32_t* class_start, const double* fwd_coeff,
                      const int8_t* fwd_expon, const double* length_cm, const int8_t* adrift, const double* driftdist,
                      int n_rec, const double* rec_coeff, const int8_t* rec_expon) {
  ArmOptics o;
  for (int k = 0; k < n_classes; ++k) {
    CosyClass c;
    c.n_terms = class_start[k + 1] - class_start[k];
    c.coeff.assign(fwd_coeff + 5 * class_start[k], fwd_coeff + 5 * class_start[k + 1]);
    c.expon.assign(fwd_expon + 5 * class_start[k], fwd_expon + 5 * class_start[k + 1]);
    c.length = length_cm[k];
    c.adrift = adrift[k];
    c.driftdist = driftdist[k];
    o.fwd.cls.push_back(std::move(c));
  }
  o.rec.n_terms = n_rec;
  o.rec.coeff.assign(rec_coeff, rec_coeff + 4 * n_rec);
  o.rec.expon.assign(rec_expon, rec_expon + 5 * n_rec);
  g_optics[arm] = std::move(o);
  return 0;
}

// sf_lookup_init (sf_lookup.f:1-80) from arrays: sf[iPm][iEm] in file order; normalised to sum 1 here.
int oracle_set_sf_table(int n_pm, int n_em, const double* pm, const double* em, const double* sf) {
  g_sf.numPm = n_pm; g_sf.numEm = n_em;
  g_sf.Pmval.assign(pm, pm + n_pm);
  g_sf.Emval.assign(em, em + n_em);
  g_sf.sfval.assign(sf, sf + (size_t)n_pm * n_em);
  double sftotnorm = 0.0;
  for (int iPm = 0; iPm < n_pm; ++iPm)
    for (int iEm = 0; iEm < n_em; ++iEm) sftotnorm = sftotnorm + g_sf.sfval[(size_t)iPm * n_em + iEm];
  for (double& v : g_sf.sfval) v = v / sftotnorm;
  g_sf.dEm.clear();
  return 0;
}
int oracle_set_sf_em_widths(int n_em, const double* dem) {
  g_sf.dEm.assign(dem, dem + n_em);
  return 0;
}
int oracle_generate_em_batch(uint64_t seed, int64_t n, const double* pm, double* out) {
  try {
    for (int64_t i = 0; i < n; ++i) {
      Rng rng; rng.seed_philox(seed, (uint64_t)i);
      out[i] = generate_em(g_sf, rng, pm[i]);
    }
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
// sf_lookup_diff and deForest on dumped vectors: in [Em, Pm] x n -> out[n]
int oracle_sf_batch(int64_t n, const double* em, const double* pm, double* out) {
  try {
    for (int64_t i = 0; i < n; ++i) out[i] = sf_lookup_diff(g_sf, em[i], pm[i]);
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

// Batch form of mc_hms / mc_shms; same row layout as simc_b200_transport_batch.
int oracle_transport_batch(int arm, int64_t n, const double* in, uint64_t seed, int ms_flag, int wcs_flag,
                           int decay_flag, int using_coll, double ctau, double* out, int32_t* flags) {
  auto it = g_optics.find(arm);
  if (it == g_optics.end()) { g_err = "optics not loaded"; return -1; }
  const ArmOptics& o = it->second;
  try {
    for (int64_t i = 0; i < n; ++i) {
      Rng rng;
      rng.seed_philox(seed, (uint64_t)i);
      Track t;
      t.rng = &rng;
      t.ctau = ctau;
      ArmCall a;
      a.dpp = in[0 * n + i]; a.x = in[1 * n + i]; a.y = in[2 * n + i]; a.z = in[3 * n + i];
      a.dxdz = in[4 * n + i]; a.dydz = in[5 * n + i]; a.m2 = in[6 * n + i]; a.p_spec = in[7 * n + i];
      a.fry = in[8 * n + i];
      a.ms_flag = ms_flag; a.wcs_flag = wcs_flag; a.decay_flag = decay_flag; a.using_coll = using_coll;
      t.Mh2_final = a.m2;
      if (arm == 1) mc_hms(t, o, a);
      else if (arm == 5) mc_shms(t, o, a);
      else if (arm == 2) mc_sos(t, o, a);
      else if (arm == 3 || arm == 4) mc_hrs(t, o, a, arm == 3);
      else { g_err = "oracle: unknown spectrometer"; return -1; }
      out[0 * n + i] = a.dpp; out[1 * n + i] = a.dxdz; out[2 * n + i] = a.dydz; out[3 * n + i] = a.y;
      out[4 * n + i] = a.x_fp; out[5 * n + i] = a.dx_fp; out[6 * n + i] = a.y_fp; out[7 * n + i] = a.dy_fp;
      out[8 * n + i] = a.pathlen; out[9 * n + i] = a.m2; out[10 * n + i] = a.resmult;
      out[11 * n + i] = (double)rng.draw;
      flags[i] = a.ok_spec ? 0 : a.stop_code;
    }
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

// The loop: tries [first, first+n) of stream `seed`, `threads` host threads (each owns a
// contiguous range and its own accumulator; integer accumulators make the merge exact).
int oracle_accum_clear(const simc_run_config* cfg, simc_accum* acc) {
  simc_oracle::accum_clear(*cfg, *acc);
  return 0;
}
// rng_mode 0: the counter-based stream (same events as the B200 path); 1: RANLUX luxury 3 through
// grnd()'s buffer, one sequential generator per thread seeded seed+t -- the reference's own
// generator and its only parallel mode (several processes with different random_seed).
int oracle_run_rng(const simc_run_config* cfg, int64_t first, int64_t n, uint64_t seed, int threads, int rng_mode,
                   simc_accum* acc);
int oracle_run(const simc_run_config* cfg, int64_t first, int64_t n, uint64_t seed, int threads, simc_accum* acc) {
  return oracle_run_rng(cfg, first, n, seed, threads, 0, acc);
}
int oracle_run_rng(const simc_run_config* cfg, int64_t first, int64_t n, uint64_t seed, int threads, int rng_mode,
                   simc_accum* acc) {
  auto ie = g_optics.find(cfg->electron_arm), ip = g_optics.find(cfg->hadron_arm);
  const ArmOptics* oe = ie == g_optics.end() ? nullptr : &ie->second;
  const ArmOptics* op = ip == g_optics.end() ? nullptr : &ip->second;
  if (threads < 1) threads = 1;
  std::vector<simc_accum> part(threads);
  std::vector<std::string> errs(threads);
  std::vector<std::thread> th;
  for (int t = 0; t < threads; ++t) {
    accum_clear(*cfg, part[t]);
    const int64_t b = first + n * t / threads, e = first + n * (t + 1) / threads;
    th.emplace_back([&, t, b, e]() {
      try {
        RanluxState st;
        if (rng_mode == 1) st.rluxgo(3, (int)(seed + 1 + t), 0, 0);
        run_range(*cfg, oe, op, b, e - b, seed, &part[t], nullptr, nullptr, 0, 0, rng_mode == 1 ? &st : nullptr,
                  g_sf.numPm ? &g_sf : nullptr, nullptr, nullptr, nullptr, nullptr,
                  g_pfermi.pval.empty() ? nullptr : &g_pfermi, g_pdf.Nx ? &g_pdf : nullptr,
              g_theory.nrhoPm ? &g_theory : nullptr, &g_maid, g_fdss.set ? &g_fdss : nullptr);
      }
      catch (const std::exception& ex) { errs[t] = ex.what(); }
    });
  }
  for (auto& x : th) x.join();
  for (int t = 0; t < threads; ++t) {
    if (!errs[t].empty()) { g_err = errs[t]; return -1; }
    merge_accum(*acc, part[t]);
  }
  return 0;
}
int oracle_event_batch(const simc_run_config* cfg, int64_t first, int64_t n, uint64_t seed, double* rec,
                       int32_t* status) {
  auto ie = g_optics.find(cfg->electron_arm), ip = g_optics.find(cfg->hadron_arm);
  try {
    run_range(*cfg, ie == g_optics.end() ? nullptr : &ie->second, ip == g_optics.end() ? nullptr : &ip->second, first,
              n, seed, nullptr, rec, status, n, 0, nullptr, g_sf.numPm ? &g_sf : nullptr, nullptr, nullptr, nullptr,
              nullptr, g_pfermi.pval.empty() ? nullptr : &g_pfermi, g_pdf.Nx ? &g_pdf : nullptr,
              g_theory.nrhoPm ? &g_theory : nullptr, &g_maid, g_fdss.set ? &g_fdss : nullptr);
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

int oracle_ntuple_batch(const simc_run_config* cfg, int64_t first, int64_t n, uint64_t seed, double* rows,
                        int32_t* n_cols, int64_t* n_rows, int64_t* try_of_row) {
  auto ie = g_optics.find(cfg->electron_arm), ip = g_optics.find(cfg->hadron_arm);
  try {
    int nc = 0;
    *n_rows = 0;
    run_range(*cfg, ie == g_optics.end() ? nullptr : &ie->second, ip == g_optics.end() ? nullptr : &ip->second, first,
              n, seed, nullptr, nullptr, nullptr, n, 0, nullptr, g_sf.numPm ? &g_sf : nullptr, rows, n_rows, &nc,
              try_of_row, g_pfermi.pval.empty() ? nullptr : &g_pfermi, g_pdf.Nx ? &g_pdf : nullptr,
              g_theory.nrhoPm ? &g_theory : nullptr, &g_maid, g_fdss.set ? &g_fdss : nullptr);
    *n_cols = nc;
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

// radc_init_ev + peaked_rad_weight + sigep on dumped vertex vectors (layout: simc_b200_radc_batch)
int oracle_radc_batch(const simc_run_config* cfg, int64_t n, const double* in, double* out) {
  try {
    for (int64_t i = 0; i < n; ++i) {
      Sim s; s.cfg = cfg;
      EventMain main; Event v;
      v.Ein = in[0 * n + i]; v.e.E = in[1 * n + i]; v.e.P = v.e.E; v.e.theta = in[2 * n + i];
      v.ue.x = in[3 * n + i]; v.ue.y = in[4 * n + i]; v.ue.z = in[5 * n + i];
      v.p.E = in[6 * n + i]; v.p.P = in[7 * n + i];
      v.up.x = in[8 * n + i]; v.up.y = in[9 * n + i]; v.up.z = in[10 * n + i];
      main.target.teff[0] = in[11 * n + i]; main.target.teff[1] = in[12 * n + i];
      v.Q2 = 2 * v.Ein * v.e.E * (1. - v.ue.z);
      v.nu = v.Ein - v.e.E;
      radc_init_ev(s, main, v);
      const double w = peaked_rad_weight_public(s, v, in[13 * n + i], in[14 * n + i], in[15 * n + i]);
      const RadEv& R = s.rad;
      out[0 * n + i] = R.bt[0]; out[1 * n + i] = R.bt[1];
      out[2 * n + i] = R.lambda[0]; out[3 * n + i] = R.lambda[1]; out[4 * n + i] = R.lambda[2];
      out[5 * n + i] = R.g[4]; out[6 * n + i] = R.hardcorfac; out[7 * n + i] = R.c[4]; out[8 * n + i] = R.c_ext[0];
      out[9 * n + i] = w;
      out[10 * n + i] = sigep(v);
      // the constants of the (Egamma1, Egamma2, Egamma3) basis and the pieces of the other option branches
      out[11 * n + i] = R.c[1]; out[12 * n + i] = R.c[2]; out[13 * n + i] = R.c[3]; out[14 * n + i] = R.c[0];
      out[15 * n + i] = R.c_int[0]; out[16 * n + i] = R.g_int;
      const int eflag = cfg->extrad_flag;
      out[17 * n + i] = extrad_phi_public(s, 1, v.Ein, v.e.E, in[13 * n + i]);
      out[18 * n + i] = extrad_phi_public(s, 2, v.Ein, v.e.E, in[13 * n + i]);
      (void)eflag;
      double ds, dh;
      schwinger(*cfg, 1.0, 450., v, true, ds, dh);
      out[19 * n + i] = ds; out[20 * n + i] = dh;
      double db, dbp;
      extrad_friedrich(cfg->etatzai, v.Ein, in[13 * n + i], R.bt[0] / cfg->etatzai, db, dbp);
      out[21 * n + i] = db; out[22 * n + i] = dbp;
      double bs, bh, dbs;
      brem(v.Ein, v.e.E, 450., R.rad_proton_this_ev, false, bs, bh, dbs);
      out[23 * n + i] = bs; out[24 * n + i] = bh; out[25 * n + i] = dbs;
    }
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

// fDSS tables from the rows of a *.GRID file (layout of simc_b200_set_fdss_table)
int oracle_set_fdss_table(const double* parton) {
  g_fdss.init(parton);
  return 0;
}
int oracle_fdss_batch(int ic, int64_t n, const double* x, const double* q2, double* out) {
  for (int64_t i = 0; i < n; ++i)
    fDSS(g_fdss, ic, x[i], q2[i], out[0 * n + i], out[1 * n + i], out[2 * n + i], out[3 * n + i], out[4 * n + i], out[5 * n + i]);
  return 0;
}

// Saghai amplitude tables (simc_b200_set_saghai_table): which = 0 K+ Lambda (12 x 10*11*19), 1 K+ Sigma0
// (12 x 20*10*19); n = 0 clears
int oracle_set_saghai_table(int which, int64_t n, const float* tbl) {
  if (which < 0 || which > 1) return -1;
  (which == 0 ? g_saghai.proton : g_saghai.sigma0).assign(tbl, tbl + n);
  return 0;
}
// eekeek / eekeeks on arrays (ss in GeV^2, q22 in MeV^2, angles in rad)
int oracle_saghai_batch(int lambda, double mrec_struck, int64_t n, const double* ss, const double* q22, const double* angl,
                        const double* theta, const double* phi, const double* epsi, double* out) {
  try {
    for (int64_t i = 0; i < n; ++i)
      out[i] = simc_oracle::eekeek(g_saghai, lambda != 0, mrec_struck, ss[i], q22[i], angl[i], theta[i], phi[i], epsi[i]);
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
// cern/fint.f on one point
double oracle_fint(int narg, const float* arg, const int* nent, const float* ent, const float* table) {
  return simc_oracle::fint(narg, arg, nent, ent, table);
}
// maidtbl slice of simc_b200_set_maid_table; n = 0 clears it
int oracle_set_maid_table(int ipi, int64_t n, const double* tbl) {
  if (ipi != 3 && ipi != 4) { g_err = "ipi must be 3 or 4"; return -1; }
  g_maid.tbl[ipi - 3].assign(tbl, tbl + n);
  return 0;
}
int oracle_sigmaid_batch(int ipi, int64_t n, const double* q2, const double* w, const double* e0, const double* costh,
                         const double* phi, double* out) {
  for (int64_t i = 0; i < n; ++i) out[i] = sigmaid_sig0(g_maid, ipi, q2[i], w[i], e0[i], costh[i], phi[i]);
  return 0;
}

// theory_init (init.f:828-905) from arrays: layout of simc_b200_set_theory_table
int oracle_set_theory_table(int doing_heavy, int n_shells, double absorption, double e_fermi, const double* nprot,
                            const double* em, const double* emsig, const double* bs_norm, const int32_t* n_pm,
                            const double* pm_first, const double* pm_bin, const double* rho) {
  TheoryTable T;
  T.nrhoPm = n_shells; T.E_Fermi = e_fermi;
  size_t pos = 0;
  for (int m = 0; m < n_shells; ++m) {
    T.nprot.push_back(nprot[m] * absorption);
    T.Em.push_back(em[m]); T.Emsig.push_back(emsig[m]); T.bs_norm.push_back(bs_norm[m]);
    T.n.push_back(n_pm[m]);
    T.pm_bin.push_back(pm_bin[m]);
    T.pm_min.push_back(pm_first[m] - pm_bin[m] / 2.);
    std::vector<double> r(rho + pos, rho + pos + n_pm[m]);
    for (double& v : r) v = v / bs_norm[m];
    T.rho.push_back(std::move(r));
    pos += n_pm[m];
    double em_int = 1.;
    if (doing_heavy) em_int = (K::pi / 2. + std::atan((em[m] - e_fermi) / (0.5 * emsig[m]))) / K::pi;
    T.Em_int.push_back(em_int);
  }
  g_theory = std::move(T);
  return 0;
}
int oracle_theory_batch(const simc_run_config* cfg, int64_t n, const double* em, const double* pm, double* out) {
  for (int64_t i = 0; i < n; ++i) out[i] = theory_sf_weight(*cfg, g_theory, em[i], pm[i]);
  return 0;
}

// dbase.f:563-587: momentum distribution, cumulative probability normalised to its last entry
int oracle_set_pfermi_table(int n, const double* pval, const double* mprob) {
  g_pfermi.pval.assign(pval, pval + n);
  g_pfermi.mprob.assign(mprob, mprob + n);
  for (int ii = 0; ii < n; ++ii) g_pfermi.mprob[ii] = g_pfermi.mprob[ii] / mprob[n - 1];
  return 0;
}
// ReadTbl, cteq5/Ctq5Pdf.f:239-281, from arrays
int oracle_set_cteq5_table(int nx, int nt, int nfmx, double al, double qini, double qmax, double xmin,
                           const double* xv, const double* qv, const double* upd) {
  g_pdf.set(nx, nt, nfmx, al, qini, qmax, xmin, xv, qv, upd);
  return 0;
}
// Ctq5Pdf on dumped (x, Q) pairs
int oracle_ctq5pdf_batch(int iparton, int64_t n, const double* x, const double* q, double* out) {
  try {
    for (int64_t i = 0; i < n; ++i) { double Q = q[i]; out[i] = Ctq5Pdf(g_pdf, iparton, x[i], Q); }
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
// SF of F1F2IN21 on dumped (W2, Q2) pairs: out[6][n] = f1p, fLp, f2p, f1n, fLn, f2n
int oracle_christy_batch(int64_t n, const double* w2, const double* q2, double* out) {
  for (int64_t i = 0; i < n; ++i)
    christy_sf(w2[i], q2[i], out[0 * n + i], out[1 * n + i], out[2 * n + i], out[3 * n + i], out[4 * n + i], out[5 * n + i]);
  return 0;
}
// peepiX on dumped vertex vectors (layout: simc_b200_semi_batch)
int oracle_semi_batch(const simc_run_config* cfg, int64_t n, const double* in, double* out) {
  try {
    for (int64_t i = 0; i < n; ++i) {
      Sim s; s.cfg = cfg; s.pdf = g_pdf.Nx ? &g_pdf : nullptr; s.fdss = g_fdss.set ? &g_fdss : nullptr;
      EventMain main; Event v;
      v.Ein = in[0 * n + i]; v.e.E = in[1 * n + i]; v.nu = in[2 * n + i]; v.Q2 = in[3 * n + i]; v.q = in[4 * n + i];
      v.uq.x = in[5 * n + i]; v.uq.y = in[6 * n + i]; v.uq.z = in[7 * n + i];
      v.pt2 = in[8 * n + i]; v.zhad = in[9 * n + i]; v.theta_pq = in[10 * n + i];
      s.pfer = in[11 * n + i]; s.pferx = in[12 * n + i]; s.pfery = in[13 * n + i]; s.pferz = in[14 * n + i];
      s.efer = in[15 * n + i];
      double surv = 1.0;
      simc_run_config c2 = *cfg;
      c2.doing_decay = 1;                 // no survival probability here (needs the focal-plane track)
      s.cfg = &c2;
      SemiDebug d;
      const double sig = peepiX(s, v, main, surv, &d);
      const double o[SIMC_SEMI_NOUT] = {sig, s.ntup.sigcm, main.davejac, d.xbj, d.u, d.ubar, d.d, d.dbar, d.s, d.sbar,
                                        d.F1p, d.F2p, d.F1n, d.F2n, d.sige, 0.0};
      for (int k = 0; k < SIMC_SEMI_NOUT; ++k) out[k * n + i] = o[k];
    }
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

// ---- end of the loop body on dumped vectors (twin of simc_b200_weight_batch, include/simc_b200.h) ----------
static void weight_tables(Sim& s) {
  s.sf = g_sf.numPm ? &g_sf : nullptr; s.pfermi = g_pfermi.pval.empty() ? nullptr : &g_pfermi; s.pdf = g_pdf.Nx ? &g_pdf : nullptr;
  s.theory = g_theory.nrhoPm ? &g_theory : nullptr; s.maid = &g_maid; s.fdss = g_fdss.set ? &g_fdss : nullptr;
}
// Dumps, for tries [first, first+n), what complete_recon_ev and complete_main read once montecarlo has returned
// (valid[i] = 0: the try did not get that far).
int oracle_weight_inputs(const simc_run_config* cfg, int64_t first, int64_t n, uint64_t seed, double* in, int32_t* valid) {
  auto ie = g_optics.find(cfg->electron_arm), ip = g_optics.find(cfg->hadron_arm);
  try {
    for (int64_t i = 0; i < n; ++i) {
      Rng rng;
      rng.seed_philox(seed, (uint64_t)(first + i));
      Sim s;
      s.cfg = cfg; s.optics_e = ie == g_optics.end() ? nullptr : &ie->second; s.optics_p = ip == g_optics.end() ? nullptr : &ip->second;
      s.rng = &rng;
      weight_tables(s);
      EventMain main;
      Event vertex, orig, recon;
      TryResult r;
      const bool ok = try_until_recon(s, main, vertex, orig, recon, r);
      valid[i] = ok ? 1 : 0;
      const double v[SIMC_WEIGHT_NIN] = {
          recon.e.E, recon.e.theta, recon.e.phi, recon.p.P, recon.p.E, recon.p.theta, recon.p.phi,
          vertex.Ein, vertex.e.E, vertex.e.theta, vertex.Q2, vertex.nu, vertex.q, vertex.p.E, vertex.p.P,
          vertex.uq.x, vertex.uq.y, vertex.uq.z, vertex.up.x, vertex.up.y, vertex.up.z, vertex.Em, vertex.Pm,
          main.phi_pq, main.t, main.epsilon, main.jacobian, main.gen_weight, vertex.zhad, vertex.pt2,
          s.pfer, s.pferx, s.pfery, s.pferz, s.efer, main.FP_p.path, main.FP_p.dx, main.FP_p.dy,
          recon.e.delta, recon.e.yptar, recon.e.xptar, recon.p.delta, recon.p.yptar, recon.p.xptar};
      for (int k = 0; k < SIMC_WEIGHT_NIN; ++k) in[k * n + i] = ok ? v[k] : 0.0;
    }
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
int oracle_weight_batch(const simc_run_config* cfg, int64_t n, const double* in, double* out) {
  try {
    for (int64_t i = 0; i < n; ++i) {
      Sim s;
      s.cfg = cfg;
      weight_tables(s);
      EventMain main;
      Event vertex, recon;
      auto I = [&](int k) { return in[k * n + i]; };
      recon.e.E = I(0); recon.e.P = recon.e.E; recon.e.theta = I(1); recon.e.phi = I(2);
      recon.p.P = I(3); recon.p.E = I(4); recon.p.theta = I(5); recon.p.phi = I(6);
      vertex.Ein = I(7); vertex.e.E = I(8); vertex.e.P = vertex.e.E; vertex.e.theta = I(9); vertex.Q2 = I(10); vertex.nu = I(11);
      vertex.q = I(12); vertex.p.E = I(13); vertex.p.P = I(14);
      vertex.uq.x = I(15); vertex.uq.y = I(16); vertex.uq.z = I(17); vertex.up.x = I(18); vertex.up.y = I(19); vertex.up.z = I(20);
      vertex.Em = I(21); vertex.Pm = I(22);
      // event.f:880-886: the vector the cross sections of (e,e'p) read
      vertex.Pmx = vertex.p.P * vertex.up.x - vertex.q * vertex.uq.x;
      vertex.Pmy = vertex.p.P * vertex.up.y - vertex.q * vertex.uq.y;
      vertex.Pmz = vertex.p.P * vertex.up.z - vertex.q * vertex.uq.z;
      main.phi_pq = I(23); main.t = I(24); main.epsilon = I(25); main.jacobian = I(26); main.gen_weight = I(27);
      vertex.zhad = I(28); vertex.pt2 = I(29);
      s.pfer = I(30); s.pferx = I(31); s.pfery = I(32); s.pferz = I(33); s.efer = I(34);
      main.FP_p.path = I(35); main.FP_p.dx = I(36); main.FP_p.dy = I(37);
      recon.e.delta = I(38); recon.e.yptar = I(39); recon.e.xptar = I(40);
      recon.p.delta = I(41); recon.p.yptar = I(42); recon.p.xptar = I(43);
      TryResult r;
      finish_try(s, main, vertex, recon, true, r);
      const double o[SIMC_WEIGHT_NOUT] = {r.success ? 1.0 : 0.0, r.pass_cuts ? 1.0 : 0.0, main.weight, main.sigcc, main.sigcc_recon,
                                          recon.Em, recon.Pm, recon.W, main.thetacm, main.phicm, s.ntup.sigcm, main.davejac,
                                          s.ntup.survivalprob, s.ntup.mm, main.wcm};
      for (int k = 0; k < SIMC_WEIGHT_NOUT; ++k) out[k * n + i] = o[k];
    }
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

// RANLUX known-answer access: n uniforms from grnd() after sgrnd(seed)
int oracle_ranlux(int seed, int lux, int64_t n, double* out) {
  RanluxState st;
  st.rluxgo(lux, seed, 0, 0);
  for (int64_t i = 0; i < n; ++i) st.ranlux(out + i, 1);
  return 0;
}
int oracle_philox_block(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
  Philox4x32::block(ctr, key, out);
  return 0;
}
int oracle_philox_uniforms(uint64_t seed, uint64_t try_index, int64_t n, double* out) {
  Rng r; r.seed_philox(seed, try_index);
  for (int64_t i = 0; i < n; ++i) out[i] = r.grnd();
  return 0;
}

// ---- trg_track.f: field of the polarised target -------------------------------------------------------------
// bz, br: 51 x 51 nodes in the file's reading order, or both null for the uniform test field; the angles as trgInit takes them
int oracle_set_field_map(const double* bz, const double* br, double theta_e_deg, double theta_p_deg) {
  g_field.init(bz, br, theta_e_deg, theta_p_deg);
  return 0;
}
// track_from_tgt on rows (x, y, z, dx, dy, mom, mass) -> (x, y, z, dx, dy, ok)
int oracle_field_batch(int spect, int64_t n, const double* in, double* out) {
  if (!g_field.set) return -1;
  for (int64_t i = 0; i < n; ++i) {
    double x = in[0 * n + i], y = in[1 * n + i], z = in[2 * n + i], dx = in[3 * n + i], dy = in[4 * n + i];
    const bool ok = simc_oracle::track_from_tgt(g_field, x, y, z, dx, dy, in[5 * n + i], in[6 * n + i], spect);
    out[0 * n + i] = x; out[1 * n + i] = y; out[2 * n + i] = z; out[3 * n + i] = dx; out[4 * n + i] = dy;
    out[5 * n + i] = ok ? 1.0 : 0.0;
  }
  return 0;
}
// trgField on points (x, y, z) -> (Bx, By, Bz)
int oracle_field_at(int spect, int64_t n, const double* xyz, double* b) {
  if (!g_field.set) return -1;
  for (int64_t i = 0; i < n; ++i) {
    const double x[3] = {xyz[0 * n + i], xyz[1 * n + i], xyz[2 * n + i]};
    double B[3];
    simc_oracle::trgField(g_field, x, B, spect);
    b[0 * n + i] = B[0]; b[1 * n + i] = B[1]; b[2 * n + i] = B[2];
  }
  return 0;
}
// n_steps Runge-Kutta steps of length dl (cm) from each state (x, y, z, vx, vy, vz): the state after every step,
// traj[(step * 6 + k) * n + i].  E = signed energy (MeV).
int oracle_field_steps(int spect, int64_t n, const double* u0, double E, double dl, int n_steps, double* traj) {
  if (!g_field.set) return -1;
  for (int64_t i = 0; i < n; ++i) {
    double u[9] = {0}, u1[9] = {0};
    for (int k = 0; k < 6; ++k) u[k] = u0[k * n + i];
    const double ts = -dl / std::sqrt(u[3] * u[3] + u[4] * u[4] + u[5] * u[5]);      // as trgTrackToPlane starts out
    for (int s = 0; s < n_steps; ++s) {
      simc_oracle::trgRK4(g_field, 90. / E, u, u1, -ts, spect);
      for (int k = 0; k < 6; ++k) { u[k] = u1[k]; traj[((int64_t)s * 6 + k) * n + i] = u[k]; }
    }
  }
  return 0;
}

}  // extern "C"
