// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into libsimc_b200.so.
// PARITY UNPINNED.  Event-level restatement: records of modules.f, COMMON /radccom/, and the
// routines of event.f, radc.f, brem.f, init.f (per-event part), target.f, enerloss_new.f,
// physics_proton.f, simc.f (montecarlo + loop body).
#pragma once
#include "../include/simc_b200.h"
#include <vector>
#include "arms.hpp"

namespace simc_oracle {

// modules.f:108-115 arm_full
struct ArmFull { double delta = 0, xptar = 0, yptar = 0, z = 0, theta = 0, phi = 0, E = 0, P = 0; };
struct Vec3 { double x = 0, y = 0, z = 0; };

// modules.f:117-130 type event
struct Event {
  double Ein = 0, Em = 0, Pm = 0, Emiss = 0, Pmiss = 0, Pmx = 0, Pmy = 0, Pmz = 0, PmPar = 0, PmPer = 0, PmOop = 0;
  double nu = 0, q = 0, Q2 = 0, Trec = 0, W = 0, Mrec = 0, epsilon = 0, theta_pq = 0, theta_tarq = 0, phi_pq = 0,
         phi_targ = 0;
  double beta = 0, phi_s = 0, phi_c = 0, zhad = 0, pt2 = 0, xbj = 0;
  ArmFull e, p;
  Vec3 ue, up, uq;
};

struct ArmSP { double delta = 0, yptar = 0, xptar = 0, z = 0; };
struct ArmFP { double x = 0, dx = 0, y = 0, dy = 0, path = 0; };

// modules.f:133-153 event_target + event_main
struct EventMain {
  double weight = 0, SF_weight = 0, gen_weight = 0, jacobian = 0, Ein_shift = 0, Ee_shift = 0;
  double sigcc = 0, sigcc_recon = 0, sigcent = 0;
  double epsilon = 0, theta_pq = 0, theta_tarq = 0, phi_pq = 0, phi_targ = 0, beta = 0, W = 0, t = 0, tmin = 0, q2 = 0;
  double pcm = 0, thetacm = 0, phicm = 0, wcm = 0, davejac = 0, johnjac = 0;
  struct { double x = 0, y = 0, z = 0, rasterx = 0, rastery = 0, teff[3] = {0, 0, 0}, Eloss[3] = {0, 0, 0}, Coulomb = 0; } target;
  ArmSP SP_e, SP_p, RECON_e, RECON_p;
  ArmFP FP_e, FP_p;
  double Trec = 0;
};

// per-event part of COMMON /radccom/ (radc.inc:5-19)
struct RadEv {
  double etta = 0, frac[3] = {0, 0, 0}, lambda[3] = {0, 0, 0}, bt[2] = {0, 0}, hardcorfac = 0;
  double c_int[4] = {0, 0, 0, 0}, c_ext[4] = {0, 0, 0, 0}, c[5] = {0, 0, 0, 0, 0}, g_int = 0, g_ext = 0,
         g[5] = {0, 0, 0, 0, 0};
  double Egamma_used[3] = {0, 0, 0}, Egamma_min[3] = {0, 0, 0}, Egamma_max[3] = {0, 0, 0};
  int ntail = 0;
  bool rad_proton_this_ev = false;
};

// simulate.inc:161-176 (the fields the loop touches)
struct NtupVars {
  double radphot = 0, radarm = 0, resfac = 0, sigcm = 0, sigcm1 = 0, sigcm2 = 0, krel = 0, mm = 0, mmA = 0, t = 0;
  double xfermi = 0;
  double rhomass = 0, rhotheta = 0;        // generate_rho.f:81, rho_decay.f:62
  // pi0 -> gamma gamma into a calorimeter arm: (E, px, py, pz) of the photons in the lab (pizero_decay.f:66-70) and
  // where they hit the calorimeter front (simc.f:1524-1546)
  double gamma1[4] = {0, 0, 0, 0}, gamma2[4] = {0, 0, 0, 0};
  double xcal_gamma1 = 0, ycal_gamma1 = 0, xcal_gamma2 = 0, ycal_gamma2 = 0;
  double survivalprob = 1.0;       // local of complete_main (event.f:1373), kept for the parity records
};

// COMMON /sftable/ (sf_lookup.inc) after sf_lookup_init: normalised to sum 1
struct SfTable {
  int numPm = 0, numEm = 0;
  std::vector<double> Pmval, Emval, sfval;   // sfval[iPm * numEm + iEm]
  std::vector<double> dEm;                   // width of each Em bin (generate_em only)
};
// SAVEd locals of sf_lookup (-fno-automatic): a call whose Em matches no branch (Em == Emval(numEm)) reuses
// the interval of the previous call, which generate_em relies on (sf_lookup.f:139-160, 196-203)
struct SfLookupState { double Em1 = 0, Em2 = 1, sf1 = 0, sf2 = 0; };
double sf_lookup_state(const SfTable& T, double Em, double Pm, SfLookupState& st);
double generate_em(const SfTable& T, class Rng& rng, double Pm);     // sf_lookup.f:181-245
// COMMON /theory/ after theory_init (init.f:828-905): independent-particle spectral function
struct TheoryTable {
  int nrhoPm = 0;
  double E_Fermi = 0;
  std::vector<double> nprot, Em, Emsig, bs_norm, Em_int;      // per shell
  std::vector<int> n;                                          // Pm_theory(i)%n
  std::vector<double> pm_min, pm_bin;                          // Pm_theory(i)%min, %bin
  std::vector<std::vector<double>> rho;                        // theory(i, 1:n), already divided by bs_norm
};
// maidtbl of sigmaid (physics_pion.f:596): [ipi 3,4][25][46][ith 1..6][columns 1..4], the part sig0 reads
struct MaidTable {
  std::vector<double> tbl[2];           // [0]: pi+ n (ipi = 3), [1]: pi- p (ipi = 4); empty = not set
};
double sigmaid_sig0(const MaidTable& M, int ipi, double q2, double w, double e0, double costh, double phi);
// Saghai amplitude tables of eekeek / eekeeks (simulate.inc:188-195, read by dbase.f:644-679), REAL*4: twelve tables
// per channel in the order zrff1..6, ziff1..6 (real then imaginary parts), each in Fortran storage order
// (iread fastest, then iq2, then iang).  proton: (10,11,19), sigma0: (20,10,19).  Empty = not set.
struct SaghaiTable {
  std::vector<float> proton, sigma0;
};
const SaghaiTable* saghai_tables();           // what oracle_set_saghai_table stored (capi.cpp); never null
double fint(int narg, const float* arg, const int* nent, const float* ent, const float* table);       // cern/fint.f:10-76
// physics_kaon.f:241-355 (lambda = true) / 357-489: sigma_eep of the Saghai model
double eekeek(const SaghaiTable& T, bool lambda, double mrec_struck, double ss, double q22, double angl, double theta,
              double phi, double epsi);
// momentum distribution of dbase.f:563-587 (deut.dat ...): mprob normalised to mprob(nump) = 1
struct PfermiTable {
  std::vector<double> pval, mprob;
};
// COMMON /CtqPar1/ /CtqPar2/ /XQrange/ /QCDtable/ after ReadTbl (cteq5/Ctq5Pdf.f:239-281)
struct Cteq5Table {
  int Nx = 0, Nt = 0, NfMx = 0;
  double Al = 0, Alambda = 0, Qini = 0, Qmax = 0, Xmin = 0;
  std::vector<double> XV, QL, UPD;
  void set(int nx, int nt, int nfmx, double al, double qini, double qmax, double xmin, const double* xv,
           const double* qv, const double* upd);
};
double Ctq5Pdf(const Cteq5Table& T, int Iparton, double X, double& Q);          // Ctq5Pdf.f:69
void christy_sf(double w2, double q2, double& f1p, double& fLp, double& f2p, double& f1n, double& fLn,
                double& f2n);                                                    // F1F2IN21_v1.0.f:2345
// SAVEd tables of fDSS after its first call (fdss/fdss.f:106-125): XUTOTF, XDTOTF, XSTOTF, XUVALF, XDVALF,
// XSVALF as (NX=35, NQ=24) column-major arrays, and ARRF = log of the grids
struct FdssTable {
  bool set = false;
  double tab[6][35 * 24];
  double arrf[35 + 24];
  void init(const double* parton);           // parton[34][24][9], the file's reading order
};
void fDSS(const FdssTable& T, int IC, double X, double Q2, double& U, double& UB, double& D, double& DB, double& S,
          double& SB);                       // fdss/fdss.f:1-215 for IH = 2, IO = 1
struct SemiDebug {
  double xbj = 0, u = 0, ubar = 0, d = 0, dbar = 0, s = 0, sbar = 0, F1p = 0, F2p = 0, F1n = 0, F2n = 0, sighad = 0,
         sige = 0;
};
double sf_lookup(const SfTable& T, double Em, double Pm);            // sf_lookup.f:97-170
double sf_lookup_diff(const SfTable& T, double Em, double Pm);       // sf_lookup.f:85-95
double deForest(const simc_run_config& cfg, const struct Event& ev); // physics_proton.f:23-135

// Everything one try needs: the run constants, both arms' optics, the RNG and the scratch
// COMMON state.  One instance per try in counter-based mode (state starts from zero, see
// DESIGN.md on the reference's stale-state quirks).
struct Sim {
  const simc_run_config* cfg = nullptr;
  const ArmOptics* optics_e = nullptr;
  const ArmOptics* optics_p = nullptr;
  const SfTable* sf = nullptr;
  const PfermiTable* pfermi = nullptr;
  const Cteq5Table* pdf = nullptr;
  const TheoryTable* theory = nullptr;
  const MaidTable* maid = nullptr;
  const FdssTable* fdss = nullptr;
  const struct TrgField* field = nullptr;       // using_tgt_field: map + this run's angles (field.hpp)
  bool field_fail_p = false;                    // track_to_tgt rejected the hadron arm's track (no STOP counter has it)
  Rng* rng = nullptr;
  double pfer = 0, pferx = 0, pfery = 0, pferz = 0, efer = 0;   // COMMON /pfermi_stuff/ (simulate.inc:212-217)
  // COMMON Mh, Mh2 (simulate.inc:91-92): run constants for every reaction but rho production, where generate_rho draws
  // the rho mass of the event and rho_decay leaves the pion mass behind; set from cfg by try_until_recon
  double Mh = 0, Mh2 = 0;
  RadEv rad;
  NtupVars ntup;
  Track trk;                     // COMMON /track/ + decdist, Mh2_final
  int stop_e = 0, stop_p = 0;    // stop code of each arm (0 = ok), -1 = arm not entered
  bool hut_e = false, hut_p = false;
  bool low_w = false;            // peepi wanted the MAID table (W < 2 GeV), see physics_meson.cpp
  long long calls[2][48] = {};   // [0] electron arm, [1] hadron arm
  int coll_steps[2][3] = {};     // slit STOP counter increments made inside mc_*_coll (hor, vert, oct), per arm
};

// Result of one pass through the loop body, simc.f:169-351
struct TryResult {
  bool gen_success = false;      // after generate
  bool success = false;          // after complete_main (+ hard cuts)
  bool pass_cuts = false;
  int stage = 0;                 // 0 failed in generate, 1 failed in P arm, 2 failed in E arm, 3 recon/main, 4 success
};

void physics_angles(double theta0, double phi0, double dx, double dy, double& theta, double& phi);      // event.f:1572
void spectrometer_angles(double theta0, double phi0, double& dx, double& dy, double theta, double phi); // event.f:1618
void trip_thru_target(Sim& s, int narm, double zpos, double energy, double theta, double& Eloss, double& radlen,
                      double mass, int typeflag);                                                       // target.f:1
void enerloss_new(Sim& s, double len, double dens, double zeff, double aeff, double epart, double mpart, int typeflag,
                  double& Eloss);                                                                       // enerloss_new.f:1
void target_musc(Sim& s, double p, double beta, double teff, double dangles[2]);                         // target.f:548
double bremos(double egamma, double k_ix, double k_iy, double k_iz, double k_fx, double k_fy, double k_fz,
              double p_ix, double p_iy, double p_iz, double p_fx, double p_fy, double p_fz, double p_fe,
              bool radiate_proton, bool exponentiate, double& bsoft, double& bhard, double& dbsoft);     // brem.f:344
double brem(double ein, double eout, double egamma, bool radiate_proton, bool exponentiate, double& bsoft,
            double& bhard, double& dbsoft);                                                                       // brem.f:6
double spen(double x);                                                                                     // radc.f:746
double schwinger(const simc_run_config& cfg, double etta, double Ecutoff, const Event& vertex, bool include_hard,
                 double& dsoft, double& dhard);                                                              // radc.f:711
void extrad_friedrich(double etatzai, double Ei, double Ecutoff, double trad, double& dbrem, double& dbrem_prime);   // radc.f:650
double extrad_phi_public(Sim& s, int itail, double E1, double E2, double Egamma);                           // radc.f:668
double gamma_fn(double x);                                                                               // radc.f:92
void radc_init_ev(Sim& s, EventMain& main, Event& vertex);                                               // init.f:655
bool complete_ev(Sim& s, EventMain& main, Event& vertex);                                                // event.f:432
bool generate_rad(Sim& s, EventMain& main, Event& vertex, Event& orig);                                  // radc.f:120
bool generate(Sim& s, EventMain& main, Event& vertex, Event& orig);                                      // event.f:126
bool montecarlo(Sim& s, Event& orig, EventMain& main, Event& recon);                                     // simc.f:1310
bool complete_recon_ev(Sim& s, Event& recon);                                                            // event.f:1056
bool complete_main(Sim& s, bool force_sigcc, EventMain& main, Event& vertex, Event& recon);              // event.f:1363
double sigep(const Event& vertex);
double peepi(Sim& s, const Event& vertex, EventMain& main);                                             // physics_pion.f:1
double peedelta(Sim& s, const Event& vertex, EventMain& main);                                          // physics_delta.f:1
double peerho(Sim& s, const Event& vertex, EventMain& main);                                            // rho_physics.f:1
void pizero_decay(Sim& s, const Event& vertex);                                                         // pizero_decay.f:1
bool generate_rho(Sim& s, Event& vertex);                                                                // generate_rho.f:1
bool rho_decay(Sim& s, Event& orig, double p_spec, double epsilon);                                      // rho_decay.f:1
double peeK(Sim& s, const Event& vertex, EventMain& main, double& survivalprob);                        // physics_kaon.f:1
double peepiX(Sim& s, const Event& vertex, EventMain& main, double& survivalprob, SemiDebug* dbg = nullptr); // semi_physics.f:1
double peaked_rad_weight_public(Sim& s, const Event& vertex, double Egamma, double emin, double emax);  // radc.f:523                                                                       // physics_proton.f:1

// One try of the loop (simc.f:169-351) in counter-based mode.
TryResult one_try(Sim& s, EventMain& main, Event& vertex, Event& orig, Event& recon);
bool try_until_recon(Sim& s, EventMain& main, Event& vertex, Event& orig, Event& recon, TryResult& r);   // generate + montecarlo
void finish_try(Sim& s, EventMain& main, Event& vertex, Event& recon, bool success, TryResult& r);       // the rest of the loop body

// results_ntu_write, results_write.f:1-269 (no target field, no pi0): returns the number of columns
int fill_ntuple(const Sim& s, const EventMain& main, const Event& vertex, const Event& orig, const Event& recon,
                double* ntu);
void accum_clear(const simc_run_config& cfg, simc_accum& a);
void merge_accum(simc_accum& a, const simc_accum& b);
void run_range(const simc_run_config& cfg, const ArmOptics* oe, const ArmOptics* op, int64_t first, int64_t n,
               uint64_t seed, simc_accum* acc, double* rec, int32_t* status, int64_t rec_stride, int64_t rec_off,
               RanluxState* ranlux = nullptr, const SfTable* sf = nullptr, double* ntu_rows = nullptr,
               int64_t* n_rows = nullptr, int* n_cols = nullptr, int64_t* try_of_row = nullptr,
               const PfermiTable* pfermi = nullptr, const Cteq5Table* pdf = nullptr,
               const TheoryTable* theory = nullptr, const MaidTable* maid = nullptr,
               const FdssTable* fdss = nullptr);
double theory_sf_weight(const simc_run_config& cfg, const TheoryTable& T, double Em, double Pm);     // event.f:1402-1428

}  // namespace simc_oracle
