/*
 * simc_b200.h -- C ABI of libsimc_b200.so, the B200 implementation of SIMC's
 * per-event Monte Carlo loop.
 *
 * The reference (JeffersonLab/simc_gfortran) has no plugin / FFI interface: the
 * seam is the body of the event loop in `program simc` (simc.f:169-351), i.e.
 *
 *     call generate(main,vertex,orig,success)                       ! event.f:126
 *     call montecarlo(orig,main,recon,success)                      ! simc.f:1310
 *     call complete_recon_ev(recon,success)                         ! event.f:1056
 *     call complete_main(.false.,main,vertex,vertex0,recon,success) ! event.f:1363
 *     call inc(...) / counters / limits_update                      ! simc.f:229-336
 *
 * plus the "structure-free" single-arm entry points mc_hms / mc_shms / mc_sos /
 * mc_hrsl / mc_hrsr (hms/mc_hms.f:1-4, shms/mc_shms.f:1-4, simulate.inc:177-179).
 * A per-event call into a GPU is useless, so every entry point here is
 * batch-level.  An unchanged Fortran driver keeps simc.f:1-165 (deck reading,
 * init, calculate_central) and simc.f:354-620 (normalisation, reports) and
 * replaces the loop by simc_b200_run(); see INTEGRATION.md for the
 * ISO_C_BINDING shim.
 *
 * Conventions: plain C, caller-owned memory, HOST pointers unless a function
 * says "device", int status return (0 = ok, <0 = error; text through
 * simc_b200_last_error), no exceptions, no torch types.  All energies MeV,
 * lengths cm, angles rad, deltas percent -- the reference's units
 * (constants.inc:3-10).
 *
 * Threads: a handle is used by one thread at a time (the reference is not
 * re-entrant at all: COMMON /track/, SAVEd tables).  Several handles may run
 * from several threads; the random generator's round keys of the seed sit in
 * device-wide constant memory, so two handles on the SAME device running with
 * DIFFERENT seeds at the same time take turns (a change of seed waits for the
 * work in flight on that device).  One process per GPU, the layout of
 * bench.py and multi.py, never meets this.
 */
#ifndef SIMC_B200_H
#define SIMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIMC_B200_ABI_VERSION 3

/* ---- spectrometer ids: electron_arm / hadron_arm numbering of dbase.f:247-263 */
enum {
  SIMC_ARM_HMS  = 1,
  SIMC_ARM_SOS  = 2,
  SIMC_ARM_HRSR = 3,
  SIMC_ARM_HRSL = 4,
  SIMC_ARM_SHMS = 5,
  SIMC_ARM_CALO_RIGHT = 7,   /* calorimeter on the HMS side (calo/mc_calo.f; dbase.f:247-259) */
  SIMC_ARM_CALO_LEFT  = 8    /* calorimeter on the SOS / SHMS side */
};

/* ---- status codes */
enum {
  SIMC_OK            =  0,
  SIMC_ERR_ARG       = -1,   /* bad argument / unsupported flag combination */
  SIMC_ERR_IO        = -2,   /* optics / table file unreadable or malformed (reference: `stop` in transp.f:319-389) */
  SIMC_ERR_CUDA      = -3,   /* CUDA runtime error; there is NO CPU fallback */
  SIMC_ERR_STATE     = -4    /* call order (e.g. run before load_optics) */
};

/* ---- basic records: field order follows modules.f so a Fortran `sequence`
 *      type can be passed by reference where the layouts coincide. */
typedef struct { double min, max; } simc_cut;                      /* modules.f:5  cutstype   */
typedef struct { double lo, hi; } simc_range;                      /* modules.f:10 rangetype  */
typedef struct { simc_cut delta, yptar, xptar, z; } simc_arm_cuts; /* modules.f:35 arm_cuts   */
typedef struct { simc_cut delta, yptar, xptar, E; } simc_arm_limits; /* modules.f:202 arm_limits */
typedef struct { simc_cut E, yptar, xptar; } simc_edge_arm;        /* modules.f:170 edge_arm  */
typedef struct {                                                   /* modules.f:178 edge_true */
  simc_edge_arm e, p;
  simc_cut Em, Pm, Mrec, Trec, Trec_struck;
} simc_edge;
typedef struct {                                                   /* modules.f:210 gen_limits */
  simc_arm_limits e, p;
  simc_cut sumEgen, Trec;
  double xwid, ywid;
} simc_gen_limits;
typedef struct {                                                   /* modules.f:150-165 spectrometer */
  double P, theta, cos_th, sin_th, phi;
  double off_x, off_y, off_z, off_xptar, off_yptar;
} simc_spectrometer;
typedef struct { double min, bin; } simc_axis;                     /* modules.f:195 axis (n is always 50: init.f:519-569) */

/* target_info, target.inc:37-49 (only what the loop reads) */
typedef struct {
  double A, Z, N, mass_amu, M, mrec_amu, Mrec, rho, thick, angle, abundancy;
  double length, zoffset, X0, X0_cm, L1, L2, fr1, fr2, xoffset, yoffset;
  double Coulomb_ave, Coulomb_min, Coulomb_max, Coulomb_constant;
  double Mtar_struck, Mrec_struck;
  int32_t fr_pattern, can;
} simc_target;

#define SIMC_NHIST 50            /* bins per histogram, init.f:519-569 */

/* Histogram slots of one set (simc.f:253-286).  The x' histograms of the gen
 * and geni sets are filled with -xptar, as in the reference. */
enum {
  SIMC_H_E_DELTA = 0, SIMC_H_E_YPTAR, SIMC_H_E_XPTAR,
  SIMC_H_P_DELTA,     SIMC_H_P_YPTAR, SIMC_H_P_XPTAR,
  SIMC_H_EM,          SIMC_H_PM,
  SIMC_H_PER_SET
};

/* Run constants: everything the loop body reads from /gnrl/ (simulate.inc:91-113),
 * /radccom/ run-level part (radc.inc:13-19), /target_info/ (target.inc:52-53),
 * /decd/ (simulate.inc:153-158), and the histogram axes H%*%min/bin.  Filled by
 * the Fortran shim after calculate_central (simc.f:89), or by
 * simc_b200_config_from_deck() for standalone runs. */
typedef struct {
  int32_t abi_version;                       /* = SIMC_B200_ABI_VERSION */

  /* reaction flags, dbase.f:138-222 */
  int32_t doing_phsp, doing_hyd_elast, doing_deuterium, doing_heavy, doing_eep;
  int32_t doing_pion, doing_kaon, doing_delta, doing_rho, doing_semi;
  int32_t doing_hydpi, doing_deutpi, doing_hepi;
  int32_t doing_hydkaon, doing_deutkaon, doing_hekaon;
  int32_t doing_hydsemi, doing_deutsemi;
  int32_t doing_semipi, doing_semika;        /* dbase.f:125-134: doing_semi splits off doing_pion / doing_kaon */
  int32_t do_fermi;                          /* dbase.f:1148: Fermi motion in semi-inclusive D(e,e'h)X */
  int32_t doing_hplus, doing_decay;
  int32_t which_pion, which_kaon;
  /* switches */
  int32_t using_rad, using_Eloss, using_Coulomb, correct_Eloss, correct_raster;
  int32_t mc_smear, hard_cuts;
  int32_t using_E_arm_montecarlo, using_P_arm_montecarlo;
  int32_t electron_arm, hadron_arm;          /* SIMC_ARM_* */
  int32_t using_HMScoll, using_SHMScoll, use_benhar_sf;
  /* radiative flags, radc.inc:3-4 (after radc_init, init.f:621-629) */
  int32_t rad_flag, extrad_flag, intcor_mode, use_expon, use_offshell_rad;
  int32_t doing_tail[3];
  int32_t hardwired_rad;
  int32_t deForest_flag;                     /* 0 = sigcc1, 1 = sigcc2, -1 = sigcc1 on shell (physics_proton.f:39-45) */
  int32_t doing_pizero, pizero_ngamma;       /* dbase.f:1006-1007: pi0 -> gamma gamma into a calorimeter arm (SIMC_ARM_CALO_*) */
  int32_t using_tgt_field, pad_flags;        /* dbase.f:1102: track both arms through the polarised target's field (trg_track.f) */

  /* /gnrl/ scalars */
  double Mh, Mh2, Ebeam, dEbeam, Ebeam_vertex_ave;
  double dE_edge_test, Egamma_gen_max, ctau, transparency;
  double drift_to_cal;                       /* cm from the target to the calorimeter front (dbase.f:1104, calo/mc_calo.f:139) */
  /* polarised target (simulate.inc:47): direction of the field axis (rad, after dbase.f:454-455), polarisation, and the
   * charge sign of the detected hadron (dbase.f:296-423) */
  double targ_Bangle, targ_Bphi, targ_pol, sign_hadron;
  /* /radccom/ run-level */
  double etatzai, Egamma_tot_max, Egamma1_max, Egamma2_max, Egamma3_max, Egamma_res_limit;

  simc_gen_limits   gen;
  simc_spectrometer spec_e, spec_p;
  simc_cut          cuts_Em, cuts_Pm;
  simc_edge         edge, VERTEXedge;
  simc_arm_cuts     SPedge_e, SPedge_p;
  double            slop_MC_e_used[3], slop_MC_p_used[3];   /* delta, yptar, xptar (modules.f:297-327) */
  simc_target       targ;

  /* histogram axes, 3 sets x 8 (histograms_module.f:6-31): [set][slot], set 0 = RECON, 1 = gen, 2 = geni */
  simc_axis         hist_axis[3][SIMC_H_PER_SET];

  /* Reference weight (e.g. central%sigcc): fixes the fixed-point quantum of the
   * deterministic weight accumulators, quantum = 2^(ilogb(w_ref)-64). <=0 -> 1. */
  double w_ref;
} simc_run_config;

/* Number of STOP slots per arm in simc_accum.stop[][] (hms/struct_hms.inc,
 * shms/struct_shms.inc, ...).  Slot 0 = trials, slot 1 = successes, slot 2 =
 * events reaching the hut, slots 3.. = 2 + stop code (see simc_b200_stop_name). */
#define SIMC_NSTOP 64

/* 128-bit two's-complement fixed-point sum: value = (hi*2^64 + lo) * 2^qexp */
typedef struct { uint64_t lo; int64_t hi; int32_t qexp; int32_t pad; } simc_fixed128;

/* What the loop leaves behind (simc.f:229-350); everything is an exact
 * integer or a min/max, hence independent of event order and GPU count. */
typedef struct {
  int64_t ntried, nsuccess, ncontribute, npasscuts, ncontribute_no_rad_proton; /* simulate.inc:60-61 */
  simc_fixed128 wtcontribute;                  /* simc.f:304 */
  simc_fixed128 sum_sigcc;                     /* simc.f:248 */
  simc_fixed128 sumerr[8], sumerr2[8];         /* e: delta,xptar,yptar,ytar; p: same (simc.f:305-322) */
  simc_fixed128 hist_w[6][SIMC_NHIST];         /* H%RECON e/p delta,yptar,xptar: sum of weights */
  int64_t       hist_n[3][SIMC_H_PER_SET][SIMC_NHIST]; /* counts: [0]=RECON (only Em,Pm used) [1]=gen [2]=geni */
  simc_range    contrib[32];                   /* limits_update order, event.f:19-72 */
  simc_range    slop[8];                       /* MC e/p delta,yptar,xptar; total Em,Pm (event.f:75-87) */
  int64_t       stop[2][SIMC_NSTOP];           /* [0] = electron arm, [1] = hadron arm */
  /* Work counters (not in the reference): calls of transp() per forward class, [arm][class-1], and
   * calls of the reconstruction map in slot 47.  bench.py turns them into algorithmic FLOPs. */
  int64_t       transp_calls[2][48];
  /* Contributing events whose weight needed a model this build lacks: peepi below W = 2 GeV blends in
   * the MAID-2007 table (physics_pion.f:88-107); they are weighted with the parametrisation alone. */
  int64_t       unsupported;
  /* Contributing events whose weight or cross section is NaN, infinite or beyond the range of the fixed-point sums
   * (|w| >= 1e38 quanta, i.e. ~2^62 * w_ref): they are counted in ncontribute / npasscuts like in the reference but
   * add nothing to wtcontribute, sum_sigcc and the weighted histograms -- the reference's sums would turn NaN.
   * Non-zero means: look at the run (e.g. semi-inclusive events with x clipped to 1). */
  int64_t       nonfinite;
} simc_accum;

typedef struct simc_handle simc_handle;

/* Host-only (no GPU needed): reads a CTP deck (`begin parm ... end parm`, the decks under infiles/) and
 * performs the reference's one-time setup -- dbase_read post-processing (dbase.f:119-553),
 * target_init (init.f:1-87), limits_init (init.f:91-572), radc_init (init.f:576-651) -- to fill
 * *out.  extra_deck_dir: where `extra_dbase_file` is looked up (the reference uses infiles/).
 * ngen / charge_mC return the deck's `ngen` and `EXPER%charge`.  err receives the message. */
int simc_b200_config_from_deck(const char* deck_path, const char* extra_deck_dir, simc_run_config* out,
                               int32_t* ngen, double* charge_mC, char* err, int errlen);

/* Same, for decks whose setup reads a data file of the reference's working directory: D(e,e'p) and
 * A(e,e'p) without use_benhar_sf take VERTEXedge%Pm and E_Fermi from h2.theory / c12.theory / fe56.theory /
 * au197.theory (init.f:326-343, 838-856), looked up in data_dir.  data_dir may be NULL for all other decks. */
int simc_b200_config_from_deck_data(const char* deck_path, const char* extra_deck_dir, const char* data_dir,
                                    simc_run_config* out, int32_t* ngen, double* charge_mC, char* err, int errlen);

/* lifecycle ------------------------------------------------------------- */
int  simc_b200_abi_version(void);
int  simc_b200_create(const simc_run_config* cfg, int device, simc_handle** out);
void simc_b200_destroy(simc_handle* h);
const char* simc_b200_last_error(const simc_handle* h);   /* h may be NULL: last create() error */
int64_t simc_b200_sizeof(int which);                      /* 0: simc_run_config, 1: simc_accum (binding self-check) */
/* Arithmetic variant.  1 (default) = "strict": separate multiply/add in the reference's
 * order (an x86-64 gfortran -O build forms no FMA, Makefile:63), COSY sums bit-identical to
 * such a build.  0 = "fast": explicit fused multiply-add and re-associated monomials in the COSY
 * polynomials only (~1e-15 relative); generation, radiation and weights are identical in both.
 * The environment variable SIMC_B200_MODE=strict|fast sets the default at create(). */
int simc_b200_set_mode(simc_handle* h, int strict_mode);
int simc_b200_sync(simc_handle* h);                       /* waits for the handle's stream */

/* optics --------------------------------------------------------------- *
 * Replaces transp_init (shared/transp.f:294-474) and the first-call loaders of
 * mc_*_recon (hms/mc_hms_recon.f:70-102): the reference's tables are SAVEd
 * locals and cannot be handed over, so the library parses the COSY files. */
int simc_b200_load_optics(simc_handle* h, int arm_id, const char* forward_path, const char* recon_path);

/* Same tables passed as arrays (used by tests and by callers that have them in
 * memory).  fwd_coeff[n_fwd_terms][5], fwd_expon[n_fwd_terms][5] (x,theta,y,phi,delta;
 * TOF lines already dropped), fwd_class_start[n_classes+1], fwd_length_cm[n_classes]
 * (the !LENGTH: comment x100, 0 if absent); rec_coeff[n_rec][4], rec_expon[n_rec][5]. */
int simc_b200_set_optics(simc_handle* h, int arm_id,
                         int n_classes, const int32_t* fwd_class_start,
                         const double* fwd_coeff, const int8_t* fwd_expon, const double* fwd_length_cm,
                         int n_rec, const double* rec_coeff, const int8_t* rec_expon);

/* Compiled maps.  The RNG-free stretches of an arm program -- magnet apertures, drifts and the COSY forward maps
 * between them (shared/transp.f:134-279 without decay) -- run as straight-line kernels generated from the loaded
 * tables (csrc/mapgen.h) and compiled for sm_100a with NVRTC; cubins are cached in <library dir>/jit_cache (or
 * $SIMC_B200_JIT_CACHE).  Arms with decay in flight or collimator stepping use the record interpreter, which is also
 * what on = 0 selects for everything (parity twin; default on, $SIMC_B200_COMPILED_MAPS=0 turns it off). */
int simc_b200_set_compiled_maps(simc_handle* h, int on);
/* Device-free: generate + compile the stretches of one set of optics tables (same arrays as simc_b200_set_optics) and
 * leave the cubin in cache_dir (NULL: the default cache).  info4 = stretches, source bytes, cubin bytes, 1 if it was
 * cached already; msg receives the compiler log on failure. */
int simc_b200_precompile_optics(int arm_id, int n_classes, const int32_t* fwd_class_start, const double* fwd_coeff,
                                const int8_t* fwd_expon, const double* fwd_length_cm, int n_rec, const double* rec_coeff,
                                const int8_t* rec_expon, int strict_mode, const char* cache_dir, const char* dump_source_path,
                                int64_t* info4, char* msg, int msg_len);

/* Device-free: the library's COSY file readers (the semantics of transp_init, shared/transp.f:294-474, and of the
 * first call of mc_*_recon, hms/mc_hms_recon.f:70-102) on one forward / reconstruction pair, as the arrays
 * simc_b200_set_optics takes; fwd_class_start needs 42 entries.  n_out[3] = classes, forward terms, recon terms. */
int simc_b200_read_optics_files(const char* forward_path, const char* recon_path, int32_t max_fwd_terms, int32_t max_rec_terms,
                                int32_t* fwd_class_start, double* fwd_coeff, int8_t* fwd_expon, double* fwd_length_cm,
                                int32_t* fwd_adrift, double* fwd_driftdist_cm, double* rec_coeff, int8_t* rec_expon,
                                int32_t* n_out, char* msg, int msg_len);

/* info[0..6] = n_classes, forward terms, non-zero forward coefficients, recon terms,
 * compiled groups, packed coefficients, ops in the arm program */
int simc_b200_optics_info(simc_handle* h, int arm_id, int64_t* info8);

/* Ntuple rows (replaces results_ntu_write, results_write.f:1-269): runs tries [first_try, first_try+n) like
 * simc_b200_run -- without touching the accumulators -- and writes one row per contributing event, in try
 * order, row-major rows[row][col] with *n_cols columns in the reference's order (NtupleInit.f:33-343;
 * the left/right column swap of results_write.f:64-118 included).  rows must hold n * SIMC_NTUPLE_MAXCOL
 * doubles.  try_of_row (may be NULL) receives the try index of each row. */
int simc_b200_ntuple_batch(simc_handle* h, int64_t first_try, int64_t n, uint64_t seed, double* rows, int32_t* n_cols,
                           int64_t* n_rows, int64_t* try_of_row);

/* Benhar-type spectral function S(Em,Pm) for A(e,e'p) with use_benhar_sf (replaces sf_lookup_init,
 * sf_lookup.f:1-80, called from dbase.f:594-621).  pm[n_pm], em[n_em] are the bin centres and
 * sf[n_pm][n_em] (Em fastest, the file's order) the proton or neutron column the caller picked; the
 * library normalises the sum to one like the reference.  load_sf_file reads benharsf_*.dat itself. */
int simc_b200_set_sf_table(simc_handle* h, int n_pm, int n_em, const double* pm, const double* em, const double* sf);
int simc_b200_load_sf_file(simc_handle* h, const char* path, int proton_flag);
/* Widths dEm(1:numEm) of the table's Em bins (the file's last column): only generate_em (sf_lookup.f:181-245)
 * reads them, i.e. pion/kaon production from A > 2, where the spectral function also supplies the missing
 * energy of the struck nucleon.  load_sf_file sets them itself; after set_sf_table call this one. */
int simc_b200_set_sf_em_widths(simc_handle* h, int n_em, const double* dem);

/* Independent-particle spectral function for D(e,e'p) and A(e,e'p) without use_benhar_sf: replaces
 * theory_init (init.f:828-905) and its COMMON /theory/ (simulate.inc:116-131).  One momentum distribution
 * rho_i(Pm) per shell i < n_shells (<= 21): nprot[i] protons, mean removal energy em[i], Lorentzian width
 * emsig[i], normalisation bs_norm[i] (the four columns of the file's shell lines); n_pm[i] equidistant
 * points (<= 500) starting at pm_first[i] with spacing pm_bin[i], values rho[] concatenated shell after
 * shell.  absorption and e_fermi are the file's first line.  The library scales nprot by the absorption,
 * divides rho by bs_norm and integrates the Lorentzians above e_fermi like the reference.
 * load_theory_file reads h2.theory / c12.theory / fe56.theory / au197.theory itself. */
int simc_b200_set_theory_table(simc_handle* h, int n_shells, double absorption, double e_fermi, const double* nprot,
                               const double* em, const double* emsig, const double* bs_norm, const int32_t* n_pm,
                               const double* pm_first, const double* pm_bin, const double* rho);
int simc_b200_load_theory_file(simc_handle* h, const char* path);

/* Nucleon momentum distribution of the deuteron (or 3He/4He/C) for Fermi-smeared meson production:
 * replaces the read of deut.dat / he3.dat / he4.dat / c12.dat in dbase.f:563-587.  pval[n] (MeV/c) and the
 * cumulative probability mprob[n]; the library divides by mprob[n-1] like the reference.  n <= 2000. */
int simc_b200_set_pfermi_table(simc_handle* h, int n, const double* pval, const double* mprob);
int simc_b200_load_pfermi_file(simc_handle* h, const char* path);

/* CTEQ5 parton distributions for the semi-inclusive weight peepiX (semi_physics.f:226-264): replaces
 * SetCtq5 / ReadTbl (cteq5/Ctq5Pdf.f:193-281).  xv[nx+1], qv[nt+1] (GeV, as in the file: the library
 * takes Log(Q/Lambda) like ReadTbl), upd[(nx+1)*(nt+1)*(nfmx+3)] in the file's order.
 * load_cteq5_file reads a cteq5*.tbl itself. */
int simc_b200_set_cteq5_table(simc_handle* h, int nx, int nt, int nfmx, double lambda, double qini, double qmax,
                              double xmin, const double* xv, const double* qv, const double* upd);
int simc_b200_load_cteq5_file(simc_handle* h, const char* path);

/* MAID-2007 table of peepi's low-W branch (sigmaid, physics_pion.f:577-640): below W = 2 GeV the pion
 * weight blends the parametrisation with MAID (physics_pion.f:131-154).  ipi = 3: pi+ n (maidpipn.dat),
 * ipi = 4: pi- p (maidpimp.dat).  tbl[25][46][6][4]: Q2 bin, W bin, the six cos(theta*) bins sigmaid uses
 * (cthmin/cthmax, physics_pion.f:601-602) and the columns sigma_T, L/T, LT/T, TT/T that enter sig0.
 * load_maid_file reads the reference's file ('(f11.6,18f8.4)', 23 angle rows per (Q2, W) bin) itself.
 * Without the table, contributing events below W = 2 GeV take the parametrisation alone and are counted in
 * simc_accum.unsupported. */
int simc_b200_set_maid_table(simc_handle* h, int ipi, const double* tbl);
int simc_b200_load_maid_file(simc_handle* h, int ipi, const char* path);

/* Saghai amplitude tables of peeK's ntuple column sigcm1 (eekeek / eekeeks, physics_kaon.f:241-489; COMMON arrays of
 * simulate.inc:188-195 filled by dbase.f:644-679).  which = 0: K+ Lambda, tbl[12][10*11*19] = zrff1..6 then ziff1..6;
 * which = 1: K+ Sigma0, tbl[12][20*10*19] = zsrff1..6 then zsiff1..6; each table in Fortran storage order (s index
 * fastest, then Q2, then angle), REAL*4 like the reference's.  The model never enters the weight (physics_kaon.f:112);
 * without the tables the column is zero.  load_saghai_files reads saghai_proton.dat and saghai_sigma0.dat from dir;
 * read_saghai_file (device-free) parses one of them into tbl. */
int simc_b200_set_saghai_table(simc_handle* h, int which, const float* tbl);
int simc_b200_load_saghai_files(simc_handle* h, const char* dir);
int simc_b200_read_saghai_file(const char* path, int which, float* tbl, char* msg, int msg_len);

/* DSS fragmentation functions for semi-inclusive kaon production (doing_semi with doing_kaon): replaces the
 * first-call table read of fDSS (fdss/fdss.f:60-125; peepiX asks for kaons at NLO, fdss/KANLO.GRID).
 * parton[34][24][9]: the file's rows in reading order (x index slowest, then Q2 index), nine columns
 * (u+ubar, d+dbar, s+sbar, c, b, gluon, u-ubar, d-dbar, s-sbar, each times z).  The library divides out the
 * (1-z)^4 z^0.5 shape like the reference.  load_fdss_file reads a *.GRID file ('9(1PE10.3)') itself. */
int simc_b200_set_fdss_table(simc_handle* h, const double* parton);
int simc_b200_load_fdss_file(simc_handle* h, const char* path);

/* Field of the polarised target (using_tgt_field): replaces trgInit (trg_track.f:243-347, called from simc.f:154) and
 * COMMON /trgFieldStrength/.  bz, br: B_field_z(iz, ir) and B_field_r(iz, ir) in T on 51 x 51 nodes 2 cm apart, in the
 * file's reading order (ir outer, iz inner); both NULL: the uniform 5 T test field trgInit builds for a blank file
 * name.  load_field_file reads trg_field_map.dat itself ("0": no field, blank: the test field). */
int simc_b200_set_field_map(simc_handle* h, const double* bz, const double* br);
int simc_b200_load_field_file(simc_handle* h, const char* path);
/* stage-level parity entry point: track_from_tgt (trg_track.f:591-672; Runge-Kutta tracking from the vertex to the
 * field-free plane z = 100 cm) on dumped vectors.  spect = -1 (electron arm) or +1; theta_deg: angle between the field
 * axis and that spectrometer as handed to trgInit.  in[k*n+i], k = 0..6: { x, y, z, dx, dy (TRANSPORT coordinates,
 * cm and slopes), mom (MeV/c, negative for a negative particle), mass (MeV) }; out[k*n+i], k = 0..5: { x, y, z, dx,
 * dy of the image track, ok }. */
#define SIMC_FIELD_NIN  7
#define SIMC_FIELD_NOUT 6
int simc_b200_field_batch(simc_handle* h, int spect, double theta_deg, int64_t n, const double* in_soa, double* out_soa);

/* stage-level parity entry point for the semi-inclusive weight: peepiX (semi_physics.f:1-617) with
 * Ctq5Pdf, the Bosted fragmentation fit and F1F2IN21 on dumped vertex vectors.  in[k*n+i], k = 0..15:
 * { Ein, e.E, nu, Q2, q, uq.x, uq.y, uq.z, pt2, zhad, theta_pq, pfer, pferx, pfery, pferz, efer };
 * out[k*n+i], k = 0..15: { sigma_eepiX, sighad (ntup%sigcm), davejac, xbj used, u, ubar, d, dbar, s, sbar,
 * F1p, F2p, F1n, F2n, sige, 0 }. */
#define SIMC_SEMI_NIN  16
#define SIMC_SEMI_NOUT 16
int simc_b200_semi_batch(simc_handle* h, int64_t n, const double* in_soa, double* out_soa);

/* the loop -------------------------------------------------------------- *
 * Replaces simc.f:169-351 for tries first_try .. first_try+n_tries-1 of the
 * counter-based stream `seed` (try t always sees the same random numbers, on
 * any GPU).  Adds into *acc (zero it with simc_b200_accum_clear first). */
int simc_b200_accum_clear(simc_handle* h, simc_accum* acc);
int simc_b200_run(simc_handle* h, int64_t first_try, int64_t n_tries, uint64_t seed, simc_accum* acc);
/* tries per pass of the stage pipeline (default 2^20) */
int simc_b200_set_batch(simc_handle* h, int64_t tries_per_batch);
/* Per-stage device time of the loop, measured with CUDA events on the handle's stream while
 * enabled: ms[0..3] = generate, hadron arm, electron arm, finish (sums since the last call);
 * launches[0..3] = launches of each.  enable = 1/0 switches the event recording. */
int simc_b200_stage_times(simc_handle* h, int enable, double* ms4, int64_t* launches4);
/* FP64 pipe microbenchmark on this device: dependent-chain-free DFMA and DMUL+DADD loops.
 * Returns TFLOP/s (FMA = 2 flops) for the roofline denominator. */
int simc_b200_fp64_peak(simc_handle* h, double* tflops_fma, double* tflops_muladd);
/* The device's double-precision log and log10 (csrc/fastlog.cuh; the loop's most frequent library calls, where the
 * reference calls libm: gauss1.f:24, musc.f:52, enerloss_new.f, brem.f) on n host values: a check entry point. */
int simc_b200_log_batch(simc_handle* h, int64_t n, const double* x, double* out_log, double* out_log10);

/* Asynchronous pieces of simc_b200_run for callers that overlap or time the
 * device work themselves (bench.py): launch on the handle's stream, then fetch. */
int simc_b200_run_async(simc_handle* h, int64_t first_try, int64_t n_tries, uint64_t seed);
int simc_b200_fetch(simc_handle* h, simc_accum* acc);          /* syncs, adds device accumulators into *acc, clears them */
int simc_b200_device_accum(simc_handle* h, void** dev_ptr, int64_t* n_int64, void** dev_minmax, int64_t* n_minmax);
/* Multi-GPU end of run: one process per GPU, each on its own range of the try index; the only exchange of
 * the whole path is ONE all-gather of the device accumulator blocks (simc_b200_device_accum: n_int64 words per rank,
 * e.g. ncclAllGather / torch.distributed.all_gather_into_tensor on the handle's stream), after which every rank folds
 * the n_ranks blocks into its own with simc_b200_reduce_gathered (device pointer, rank after rank; a kernel on the
 * handle's stream: counters add, 128-bit sums add with carry, range keys take min / max -- integers, so all ranks hold
 * identical bits) and reads the total with simc_b200_fetch.  simc_b200_accum_merge is the host form of the same fold for
 * accumulators already fetched (CPU ranks, separate runs of one deck): exact, SIMC_ERR_ARG if two non-empty
 * fixed-point sums sit on different quanta (different w_ref). */
int simc_b200_reduce_gathered(simc_handle* h, const void* d_gathered, int n_ranks);
int simc_b200_accum_merge(simc_accum* into, const simc_accum* from);
void* simc_b200_stream(simc_handle* h);                        /* cudaStream_t */
int64_t simc_b200_launch_count(const simc_handle* h);          /* kernels launched so far */

/* single-arm parity entry point ---------------------------------------- *
 * Batch form of mc_hms / mc_shms / ... (hms/mc_hms.f:1-4).  Row i of in[] is
 * { dpp(%), x, y, z, dxdz, dydz, m2, p_spec, fry } (SoA: in[k*n+i]); the random
 * stream is (seed, try=i).  out[k*n+i], k = 0..11:
 * { dpp_recon, dxdz_recon(=dph), dydz_recon(=dth), y_recon, x_fp, dx_fp, y_fp, dy_fp,
 *   pathlen, m2_final, resmult, n_draws } -- the in/out convention of
 * mc_hms.f:428-431; rows that stop keep their pre-hut values where the
 * reference would.  flags[i] = 0 if ok_spec else the stop code. */
#define SIMC_TRANSPORT_NIN  9
#define SIMC_TRANSPORT_NOUT 12
int simc_b200_transport_batch(simc_handle* h, int arm_id, int64_t n,
                              const double* in_soa, uint64_t seed,
                              int ms_flag, int wcs_flag, int decay_flag, int using_coll,
                              double* out_soa, int32_t* flags);
/* Same with in/out already resident in device memory (HBM); used by bench.py
 * for the HBM-resident timing leg.  Asynchronous on the handle's stream. */
int simc_b200_transport_batch_device(simc_handle* h, int arm_id, int64_t n,
                              const double* d_in_soa, uint64_t seed,
                              int ms_flag, int wcs_flag, int decay_flag, int using_coll,
                              double* d_out_soa, int32_t* d_flags);

/* whole-event parity entry point: per-try records instead of accumulators.
 * rec[k*n+i], k = 0..SIMC_EVENT_NREC-1 (see simc_b200_event_field_name). */
#define SIMC_EVENT_NREC 60
/* Columns of one ntuple row, results_ntu_write (results_write.f:1-269): 46 for (e,e'p), 53 for pion, 55
 * for kaon and 56 for semi-inclusive production (no target field). */
#define SIMC_NTUPLE_MAXCOL 68      /* widest rows: rho production 59, pi0 -> gamma gamma 65 / 68 columns (NtupleInit.f:101-343) */
int simc_b200_event_batch(simc_handle* h, int64_t first_try, int64_t n, uint64_t seed,
                          double* rec_soa, int32_t* status);
const char* simc_b200_event_field_name(int k);

/* stage-level parity entry point for the radiative corrections and the cross-section weight ---- *
 * radc_init_ev + basicrad_init_ev (init.f:655-813), peaked_rad_weight (radc.f:523-646) and sigep
 * (physics_proton.f:1-22) on dumped per-event vertex vectors.  in[k*n+i], k = 0..15:
 * { Ein, e.E, e.theta, ue.x, ue.y, ue.z, p.E, p.P, up.x, up.y, up.z, teff(1), teff(2), Egamma, Emin, Emax };
 * out[k*n+i], k = 0..25: { bt(1), bt(2), lambda(1..3), g(4), hardcorfac, c(4), c_ext(0),
 * peaked_rad_weight(basicrad_weight = 1), sigep,
 * c(1), c(2), c(3), c(0), c_int(0), g_int                       (basicrad_init_ev, init.f:732-813),
 * extrad_phi(1,Ein,e.E,Egamma), extrad_phi(2,...)               (radc.f:668-707, the handle's extrad_flag),
 * schwinger's dsoft, dhard at Ecutoff = 450                     (radc.f:711-742),
 * extrad_friedrich(Ein, Egamma, bt(1)/etatzai): dbrem, dbrem'   (radc.f:650-664),
 * brem(Ein, e.E, 450, rad_proton_this_ev): bsoft, bhard, dbsoft (brem.f:6-214) }.
 * Every radiative option of the handle's run constants is honoured (rad_flag, extrad_flag, intcor_mode,
 * use_offshell_rad). */
#define SIMC_RADC_NIN  16
#define SIMC_RADC_NOUT 26
int simc_b200_radc_batch(simc_handle* h, int64_t n, const double* in_soa, double* out_soa);

/* Stage-level parity entry point for the end of the loop body: complete_recon_ev (event.f:1056-1359), complete_main
 * (event.f:1363-1569: spectral-function weight, sigep / deForest / peepi / peedelta / peeK / peepiX, Coulomb factor,
 * final weight), pass_cuts and the hard cuts (simc.f:219-246) on dumped per-event vectors -- what those routines read
 * from `recon`, `vertex`, `main` and COMMON /pfermi_stuff/ after montecarlo returned.  in[k*n+i], k =
 *  0 recon.e.E   1 recon.e.theta  2 recon.e.phi    3 recon.p.P     4 recon.p.E    5 recon.p.theta  6 recon.p.phi
 *  7 vertex.Ein  8 vertex.e.E     9 vertex.e.theta 10 vertex.Q2   11 vertex.nu   12 vertex.q
 * 13 vertex.p.E 14 vertex.p.P    15-17 vertex.uq   18-20 vertex.up 21 vertex.Em  22 vertex.Pm
 * 23 main.phi_pq 24 main.t       25 main.epsilon   26 main.jacobian 27 main.gen_weight
 * 28 vertex.zhad 29 vertex.pt2   30 pfer  31-33 pferx,y,z  34 efer
 * 35 main.FP.p.path  36 main.FP.p.dx  37 main.FP.p.dy
 * 38-40 recon.e.delta,yptar,xptar   41-43 recon.p.delta,yptar,xptar
 * out[k*n+i], k = 0 success  1 pass_cuts  2 main.weight  3 main.sigcc  4 main.sigcc_recon  5 recon.Em  6 recon.Pm
 *  7 recon.W  8 main.thetacm  9 main.phicm  10 ntup.sigcm  11 main.davejac  12 survivalprob  13 ntup.mm  14 main.wcm
 * The run's tables (spectral function, theory, CTEQ5, DSS, MAID) must be set as for simc_b200_run. */
#define SIMC_WEIGHT_NIN  44
#define SIMC_WEIGHT_NOUT 15
int simc_b200_weight_batch(simc_handle* h, int64_t n, const double* in_soa, double* out_soa);

const char* simc_b200_stop_name(int arm_id, int code);

/* end of run ------------------------------------------------------------- *
 * Host-only.  Replaces the normalisation and resolution block of program simc (simc.f:94-101, 366-432):
 * luminosity from the charge and the target, the generation volume of the reaction, normfac, the normalised
 * yield wtcontribute*normfac, and mean / rms of the reconstruction errors. */
typedef struct {
  double luminosity;          /* ub^-1, simc.f:94-101 */
  double genvol;              /* product of the generated ranges, simc.f:376-396 */
  double normfac;             /* luminosity / ntried * nevent * genvol (1 for doing_phsp), simc.f:368-398 */
  double yield;               /* wtcontribute * normfac: counts for EXPER%charge, simc.f:399 */
  double central_sigcc_ave;   /* sum_sigcc / nevent, simc.f:959 */
  int64_t nevent;             /* simc.f:346-350: ntried when ngen < 0 (every try counts), the successes when ngen > 0 */
  double aveerr[8], resol[8]; /* e: delta, xptar, yptar, ytar; p: same (simc.f:406-431); 0 if npasscuts <= 1 */
} simc_results;
/* ngen: the deck's ngen (its sign selects what nevent counts). */
int simc_b200_normalise(const simc_run_config* cfg, const simc_accum* acc, int32_t ngen, double charge_mC, simc_results* out);

/* ---- the reference's end-of-run text files (simc.f:446-1139), byte layout of the Fortran formats ---------------- */
/* Init-only values subroutine report prints beyond simc_run_config (targ%Eloss / teff / musc_max extremes of
 * limits_init, slop%total%Em%used, deck switches the loop does not read). */
typedef struct {
  int32_t ngen, random_seed, one_tail;
  int32_t doing_pizero, pizero_ngamma, use_first_cer, using_tgt_field;
  int32_t doing_hyddelta, doing_deutdelta, doing_hedelta, doing_hydrho, doing_deutrho, doing_herho;
  int32_t pad;
  double charge_mC;
  double Eloss_ave[3], Eloss_min[3], Eloss_max[3];      /* targ%Eloss(1:3): beam, e, hadron */
  double teff_ave[3], teff_min[3], teff_max[3];
  double musc_max[3], musc_nsig_max;
  double slop_total_Em_used;
  char theory_file[128];
} simc_report_info;
int simc_b200_report_info_from_deck(const char* deck_path, const char* extra_deck_dir, const char* data_dir,
                                    simc_report_info* out, char* err, int errlen);
/* event_central (modules.f:156-168) as calculate_central fills it (simc.f:1143-1306): the event with both particles on
 * the spectrometer axes through complete_recon_ev, radc_init_ev and complete_main(force_sigcc).  Kinematics and
 * radiative constants are host arithmetic; sigcc goes through simc_b200_weight_batch on h (pass h = NULL to skip it). */
typedef struct {
  double e_delta, e_xptar, e_yptar, p_delta, p_xptar, p_yptar;
  double Q2, q, nu, Em, Pm, W, MM, sigcc;
  double hardcorfac, etatzai, frac[3], lambda[3], bt[2], c_int[4], c_ext[4], c[4], g_int, g_ext, g[4];
} simc_central;
int simc_b200_central_event(simc_handle* h, const simc_run_config* cfg, const simc_report_info* info, simc_central* out);
/* <base>.geni (STOP counters, simc.f:446-537), <base>.gen (histograms, simc.f:539-612), <base>.hist (subroutine
 * report, simc.f:644-1139; timestring = ctime() text of the loop's start and end). */
/* The Fortran edit descriptors Fw.d (kind 'F') and Ew.d (kind 'E', 0.dddE+ee form) as the writers produce them. */
int simc_b200_format_real(int kind, int w, int d, double v, char* out, int outlen);
int simc_b200_write_geni(const char* path, const simc_run_config* cfg, const simc_accum* acc);
int simc_b200_write_gen(const char* path, const simc_run_config* cfg, const simc_accum* acc);
int simc_b200_write_hist(const char* path, const simc_run_config* cfg, const simc_report_info* info, const simc_central* central,
                         const simc_accum* acc, const simc_results* res, const char* timestring1, const char* timestring2);

/* Ntuple file in the reference's layout (NtupleInit.f:32,352-355; results_write.f:264-266): a Fortran
 * unformatted sequential file -- record "NtupleSize" (int32), one 16-character record per tag, then one
 * 8-byte record per column of every row; each record sits between two 4-byte length markers.  This is what
 * util/root_tree/make_root_tree.f and util/ntuple read.  ntuple_tags fills tags[n][17] (NUL-terminated) for
 * the reaction of cfg and returns n (46 / 53 / 55 / 56), or < 0. */
int simc_b200_ntuple_tags(const simc_run_config* cfg, char (*tags)[17], int max_tags);
typedef struct simc_ntuple_file simc_ntuple_file;
int simc_b200_ntuple_open(const simc_run_config* cfg, const char* path, simc_ntuple_file** out);
int simc_b200_ntuple_append(simc_ntuple_file* f, const double* rows, int64_t n_rows);   /* rows[n_rows][SIMC_NTUPLE_MAXCOL] */
int simc_b200_ntuple_close(simc_ntuple_file* f);

#ifdef __cplusplus
}
#endif
#endif /* SIMC_B200_H */
