// The event loop on the device (replaces simc.f:169-351): four stage kernels over a batch of
// tries, with survivors compacted between stages so that warps stay full.
//
//   k_generate : one thread per try.  generate + complete_ev + generate_rad; geni histograms.
//                Survivors get a slot in the SoA state buffer and an entry in list 0.
//   k_arm<P>   : hadron arm.  target multiple scattering, SP quantities, TRANSPORT
//                coordinates, single-arm Monte Carlo (transport.cuh), recon of the arm,
//                most-probable energy-loss correction.  Survivors -> list 1.
//   k_arm<E>   : electron arm, same.  Survivors -> list 2.
//   k_finish   : complete_recon_ev + complete_main + pass_cuts + histogram / counter / sum /
//                range accumulation into exact integer accumulators (DevAccum).
//
// Nothing here depends on the order in which slots are handed out: the random stream is keyed by
// the try index, and every accumulator is an integer sum or a min/max.
#pragma once
#include "event.cuh"
#include "transport.cuh"
#include "field.cuh"
#include "kernels.h"

namespace simc {
namespace SIMC_VARIANT_NS {

// ---- state buffer ------------------------------------------------------------------------
enum StateField : int {
  // ---- the block k_generate / k_regen write in one go (store_event / load_event): F_GEN_END doubles in groups of
  // four (one 32-byte sector each), grouped by the kernels that read them later
  F_TRY = 0, F_DRAW, F_NTAIL, F_RADP,
  F_TX, F_TY, F_TZ, F_RASTERY,
  F_ELOSS0, F_ELOSS1, F_ELOSS2, F_COULOMB,
  F_TEFF0, F_TEFF1, F_TEFF2, F_GENW,
  F_OEIN, F_OEE, F_OEDELTA, F_OPE,
  F_OPP, F_OPDELTA, F_VPYP, F_VPXP,
  F_VEYP, F_VEXP, F_VEDELTA, F_VPDELTA,
  F_VEIN, F_VEE, F_VETHETA, F_VQ2,
  F_VPE, F_VPP, F_VEM, F_VPM,
  F_UEX, F_UEY, F_UEZ, F_HARDCOR,
  F_UPX, F_UPY, F_UPZ, F_VTREC,
  F_RC_CEXT0, F_RC_GEXT, F_RC_G1, F_RC_G2,
  F_RC_G4, F_RC_BT0, F_RC_BT1, F_RD_WHICH,
  F_RD_EMIN, F_RD_EMAX, F_RD_EG, F_RD_BW,
  F_JAC, F_EINSHIFT, F_EESHIFT, F_MTREC,
  F_EG0, F_EG1, F_EG2, F_VEPHI,
  F_VPTHETA, F_VPPHI, F_VNU, F_VQ,
  F_UQX, F_UQY, F_UQZ, F_MEPS,
  F_MTHPQ, F_MPHIPQ, F_MT, F_MW,
  F_ZHAD, F_PT2, F_PFER, F_EFER,
  F_PFERX, F_PFERY, F_PFERZ, F_RASTERX,         // F_RASTERX: main%target%rasterx, read by the target-field tracking only
  F_OPYP, F_OPXP, F_RHOMASS, F_RHOTHETA,        // rho production only: orig%p%yptar/xptar of the decay pion, ntup%rhomass/rhotheta
  F_GEN_END,
  // ---- written by the later stages
  F_STAGE = F_GEN_END, F_STOP_P, F_STOP_E, F_RESFAC,
  F_DANG0, F_DANG1, F_FPP_PATH, F_FPE_PATH,
  F_SPP_D, F_SPP_Y, F_SPP_X, F_SPP_Z, F_RCP_D, F_RCP_Y, F_RCP_X, F_RCP_Z,
  F_SPE_D, F_SPE_Y, F_SPE_X, F_SPE_Z, F_RCE_D, F_RCE_Y, F_RCE_X, F_RCE_Z,
  F_RP_P, F_RP_E, F_RP_TH, F_RP_PH, F_RE_E, F_RE_TH, F_RE_PH, F_FPP_DX,
  F_FPP_DY, F_WEIGHT, F_SIGCC, F_SIGCC_RECON, F_PASSCUTS, F_REM, F_RPM, F_RW,
  // track of the arm in flight, between the segments of its program (two sectors per slot)
  F_TK_XS, F_TK_YS, F_TK_DX, F_TK_DY, F_TK_DPP, F_TK_P, F_TK_M2, F_TK_PATH, F_TK_DECD, F_TK_DFLAG, F_TK_FRY, F_TK_PAD,
  // results of peepi / peeK / peepiX (record mode)
  F_THCM, F_PHICM, F_SIGCM, F_DAVEJAC, F_SURV, F_MM, F_WCM, F_XFERMI,
  // ntuple rows (record mode only): focal-plane positions, decay bookkeeping, then the row itself
  F_FPP_X, F_FPP_Y, F_FPE_X, F_FPE_DX, F_FPE_Y, F_FPE_DY, F_DECDIST, F_MH2FINAL,
  F_NTU0, F_NTU_LAST = F_NTU0 + SIMC_NTUPLE_MAXCOL - 1,
  F_NFIELDS
};
static_assert(F_GEN_END % 4 == 0 && F_TK_XS % 4 == 0 && F_SPP_D % 4 == 0 && F_SPE_D % 4 == 0 && F_RCP_D % 4 == 0 &&
              F_RCE_D % 4 == 0 && F_RP_P % 4 == 0 && F_TX % 4 == 0 && F_ELOSS0 % 4 == 0 && F_TEFF0 % 4 == 0 && F_OEIN % 4 == 0 &&
              F_OPP % 4 == 0 && F_VEYP % 4 == 0, "32-byte groups");

// the compiled stretches (mapgen.h) address the track by row offsets from F_TK_XS
static_assert(F_TK_YS == F_TK_XS + 1 && F_TK_DX == F_TK_XS + 2 && F_TK_DY == F_TK_XS + 3 && F_TK_DPP == F_TK_XS + 4 &&
              F_TK_P == F_TK_XS + 5 && F_TK_M2 == F_TK_XS + 6 && F_TK_PATH == F_TK_XS + 7, "track rows: mapgen.h");

// Layout of the state buffer.  SIMC_STATE_AOS = 1 (default): one record of kStateStride doubles per slot
// (base[slot * kStateStride + field]).  From the first aperture on, the survivors are a shrinking, scattered subset of
// the slots (a fifth of them reach k_finish), so a warp's 32 loads of one field touch 32 different 32-byte sectors
// whatever the layout; with records, the other three doubles of each sector are the neighbouring fields of the same
// event, which the kernel reads next, instead of three events that died.  SIMC_STATE_AOS = 0: struct of arrays
// (base[field * cap + slot]), the round-1 layout, coalesced only while every slot is alive.
#ifndef SIMC_STATE_AOS
#define SIMC_STATE_AOS 1
#endif
constexpr int kStateStride = (F_NFIELDS + 15) / 16 * 16;      // records start on 128-byte lines
#ifndef SIMC_STATE_ST_HINT
#define SIMC_STATE_ST_HINT 0
#endif
#ifndef SIMC_STATE_LD_HINT
#define SIMC_STATE_LD_HINT 0
#endif
// 32-byte accesses (sm_100: ld/st.global.v4.f64), for groups of four fields that start at a multiple of four
__device__ __forceinline__ void st_v4(double* p, double a, double b, double c, double d) {
#if SIMC_STATE_ST_HINT == 1
  asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
#elif SIMC_STATE_ST_HINT == 2
  asm volatile("st.global.L2::evict_first.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
#else
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
#endif
}
__device__ __forceinline__ void ld_v4(const double* p, double& a, double& b, double& c, double& d) {
#if SIMC_STATE_LD_HINT == 1
  asm volatile("ld.global.cs.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
#elif SIMC_STATE_LD_HINT == 2
  asm volatile("ld.global.L1::no_allocate.L2::evict_first.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
#else
  asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
#endif
}
struct StateBuf {
  double* base;
  long long cap;
#if SIMC_STATE_AOS
  __device__ __forceinline__ double ld(int f, long long slot) const { return base[slot * kStateStride + f]; }
  __device__ __forceinline__ void st(int f, long long slot, double v) const { base[slot * kStateStride + f] = v; }
  __device__ __forceinline__ void ld4(int f, long long slot, double& a, double& b, double& c, double& d) const {
    ld_v4(base + slot * kStateStride + f, a, b, c, d);
  }
  __device__ __forceinline__ void st4(int f, long long slot, double a, double b, double c, double d) const {
    st_v4(base + slot * kStateStride + f, a, b, c, d);
  }
#else
  __device__ __forceinline__ double ld(int f, long long slot) const { return base[(long long)f * cap + slot]; }
  __device__ __forceinline__ void st(int f, long long slot, double v) const { base[(long long)f * cap + slot] = v; }
  __device__ __forceinline__ void ld4(int f, long long slot, double& a, double& b, double& c, double& d) const {
    a = ld(f, slot); b = ld(f + 1, slot); c = ld(f + 2, slot); d = ld(f + 3, slot);
  }
  __device__ __forceinline__ void st4(int f, long long slot, double a, double b, double c, double d) const {
    st(f, slot, a); st(f + 1, slot, b); st(f + 2, slot, c); st(f + 3, slot, d);
  }
#endif
};
// strides (in doubles) from field to field and from slot to slot, for code that addresses the buffer itself (mapgen.h)
__host__ __device__ inline long long state_field_stride(long long cap) { return SIMC_STATE_AOS ? 1 : cap; }
__host__ __device__ inline long long state_slot_stride() { return SIMC_STATE_AOS ? kStateStride : 1; }

// ---- exact accumulators --------------------------------------------------------------------
struct DevAccum {
  unsigned long long counters[8];                 // ntried, nsuccess, ncontribute, npasscuts, nco_no_rad_proton, unsupported, nonfinite
  unsigned long long wt[2], sigcc[2];             // 128-bit two's complement (lo, hi)
  unsigned long long sumerr[8][2], sumerr2[8][2];
  unsigned long long hist_w[6][SIMC_NHIST][2];
  unsigned long long hist_n[3][SIMC_H_PER_SET][SIMC_NHIST];
  long long contrib_lo[32], contrib_hi[32], slop_lo[8], slop_hi[8];   // order-preserving keys of doubles
  unsigned long long stop[2][SIMC_NSTOP];
  unsigned long long transp_calls[2][48];
};

struct LoopArgs {
  const simc_run_config* cfg;      // device copy
  MatTable mt;                     // per-material energy-loss constants (target.cuh), made on the host
  SfDev sf;                        // Benhar spectral function (A(e,e'p) only)
  Cteq5Dev pdf;                    // CTEQ5 parton distributions (semi-inclusive production only)
  PfermiDev pfm;                   // nucleon momentum distribution (deuterium semi-inclusive production only)
  FdssDev fdss;                    // DSS fragmentation functions (semi-inclusive kaons only)
  MaidDev maid;                    // MAID-2007 slice of peepi's low-W branch (null unless set)
  SaghaiDev saghai;                // Saghai amplitude tables of peeK's ntuple column sigcm1 (null unless set)
  TheoryDev theory;                // independent-particle spectral function (D(e,e'p), A(e,e'p) without use_benhar_sf)
  FieldDev field;                  // field of the polarised target (using_tgt_field only; map null otherwise)
  StateBuf st;
  unsigned* lists;                 // [kLoopLists][cap] (kernels.h)
  unsigned* counts;                // [0] slots handed out, [1 + l] length of list l
  int op_begin, op_end;            // ops of the arm program a k_arm launch runs (kernels.h: ArmStage)
  int in_idx, out_idx;             // the list it reads, the list its survivors go to
  DevAccum* acc;
  long long first_try, n_tries;
  unsigned long long seed;
  int qexp_w;                      // quantum exponent of the weight sums
  int record_mode;                 // every try gets a slot (parity entry point)
};

__device__ __forceinline__ long long dkey(double d) {          // monotonic double -> int64
  const long long i = __double_as_longlong(d);
  return i >= 0 ? i : (i ^ 0x7fffffffffffffffLL);
}

// double -> 128-bit fixed point with quantum 2^qexp, round to nearest even (same rule as the
// oracle's to_fixed); values >= 2^62 quanta are already integers.
__device__ __forceinline__ void to_fixed(double x, int qexp, unsigned long long& lo, unsigned long long& hi) {
  const double sc = ldexp(x, -qexp);
  if (!(fabs(sc) < 4.611686018427387904e18)) {
    if (!(fabs(sc) < 1.0e38)) { lo = 0; hi = 0; return; }
    const double h = floor(sc / 18446744073709551616.0);
    const double l = sc - h * 18446744073709551616.0;
    hi = (unsigned long long)(long long)h;
    lo = (unsigned long long)l;
    return;
  }
  const long long v = __double2ll_rn(sc);
  lo = (unsigned long long)v;
  hi = v < 0 ? ~0ULL : 0ULL;
}
__device__ __forceinline__ void add128(unsigned long long* dst, double x, int qexp) {
  unsigned long long lo, hi;
  to_fixed(x, qexp, lo, hi);
  const unsigned long long old = atomicAdd(&dst[0], lo);
  const unsigned long long carry = (old + lo < old) ? 1ULL : 0ULL;
  if (hi + carry) atomicAdd(&dst[1], hi + carry);
}
__device__ __forceinline__ void upd_range(long long* lo, long long* hi, double v) {
  const long long k = dkey(v);
  atomicMin(lo, k);
  atomicMax(hi, k);
}
// simc.f:624-640: ibin = nint(0.5+(val-min)/bin), kept if 1..50
__device__ __forceinline__ int hist_bin(const simc_axis& ax, double val) {
  const double r = round(0.5 + (val - ax.min) / ax.bin);
  if (!(r >= 1.0 && r <= (double)SIMC_NHIST)) return -1;
  return (int)r - 1;
}

// Warp-aggregated append: returns this lane's position in a list whose length is *counter.
__device__ __forceinline__ unsigned warp_append(unsigned* counter, bool take) {
  const unsigned mask = __ballot_sync(__activemask(), take);
  if (!take) return 0u;
  const unsigned lane = threadIdx.x & 31u;
  const int leader = __ffs(mask) - 1;
  unsigned base = 0;
  if ((int)lane == leader) base = atomicAdd(counter, __popc(mask));
  base = __shfl_sync(mask, base, leader);
  return base + __popc(mask & ((1u << lane) - 1u));
}

struct GaussFn {
  __device__ __forceinline__ double operator()(DevRng& r, double nsig) const { return gauss1(r, nsig); }
};

// ---- stage 1: generation -------------------------------------------------------------------
// CTA size of the generation kernels: the code is one long straight line, and every CTA resident on an
// SM streams it through the instruction cache at its own position; fewer, larger CTAs (whose warps the
// phase barriers keep together) mean fewer streams.  Registers are capped at 65536 / (block * min blocks).
#ifndef SIMC_GEN_BLOCK
#define SIMC_GEN_BLOCK 512
#endif
#ifndef SIMC_GEN_MIN_BLOCKS
#define SIMC_GEN_MIN_BLOCKS 2
#endif
constexpr int kGenBlock = SIMC_GEN_BLOCK;
// Run constants in shared memory (bit 0: generation kernels, bit 1: k_radw / k_finish, bit 2: k_arm): the routines read
// simc_run_config through a reference all along their straight-line code; in global memory those warp-uniform loads
// share the L1 with the local frames of up to 1024 threads and miss it a third of the time.
#ifndef SIMC_CFG_SMEM
#define SIMC_CFG_SMEM 0
#endif
static_assert(sizeof(simc_run_config) % 8 == 0, "copied as 8-byte words");
// (the caller's next __syncthreads() publishes the copy)
__device__ __forceinline__ void cfg_to_shared(simc_run_config& dst, const simc_run_config* src, int nthreads) {
  for (int i = threadIdx.x; i < (int)(sizeof(simc_run_config) / 8); i += nthreads)
    ((unsigned long long*)&dst)[i] = ((const unsigned long long*)src)[i];
}
// launch shape of the end-of-loop kernels (k_radw, k_finish): straight-line code on compacted survivors as well
#ifndef SIMC_FIN_BLOCK
#define SIMC_FIN_BLOCK 256
#endif
#ifndef SIMC_FIN_MIN_BLOCKS
#define SIMC_FIN_MIN_BLOCKS 4
#endif
constexpr int kFinBlock = SIMC_FIN_BLOCK;
constexpr int kRegenList = kRegenListIdx;

struct GenFlags { bool semi, fermi, meson, heavy, rho, xtra, field; };
__device__ __forceinline__ GenFlags gen_flags(const simc_run_config& cfg) {
  GenFlags g;
  g.semi = cfg.doing_semi != 0;
  g.fermi = cfg.doing_deutsemi || cfg.doing_deutpi || cfg.doing_deutkaon || cfg.doing_hepi || cfg.doing_hekaon;   // nucleon momentum thrown
  g.rho = cfg.doing_rho != 0;                   // the rho is thrown in the photon-nucleon c.m. inside complete_ev
  g.meson = cfg.doing_pion || cfg.doing_kaon || cfg.doing_delta || g.semi || cfg.doing_deuterium || g.rho;   // hadron energy from two-body kinematics (or thrown: semi)
  g.heavy = cfg.doing_heavy != 0;
  g.field = cfg.using_tgt_field != 0;           // the raster's x position travels with the record (track_to_tgt)
  g.xtra = g.rho || cfg.doing_pizero;           // the record's last group is in use (rho: mass, decay angle; pi0: decay angles)
  return g;
}

// geni histograms: every try, from the vertex values it ended with (simc.f:253-262)
__device__ __forceinline__ void geni_hist(const simc_run_config& cfg, unsigned (*h_geni)[SIMC_NHIST], const EventState& s) {
  const double gv[8] = {s.v_edelta, s.v_eyptar, -s.v_exptar, s.v_pdelta, s.v_pyptar, -s.v_pxptar, s.v_Em, s.v_Pm};
#pragma unroll
  for (int k = 0; k < 8; ++k) warp_hist_add(h_geni[k], hist_bin(cfg.hist_axis[2][k], gv[k]));
}

// The try's record in the state buffer: the first F_GEN_END fields, written (and, for the second pass of a try whose
// incoming electron radiated, read back) as 32-byte groups.  Fields a reaction does not use hold zeros.
__device__ __forceinline__ void store_event(const LoopArgs& A, const GenFlags& g, unsigned slot, long long i, unsigned draw,
                                            const EventState& s, const GenRad& gr, bool ok, bool carry) {
  const StateBuf& S = A.st;
  (void)carry;
  double r[F_GEN_END];
  r[F_TRY] = (double)i; r[F_DRAW] = (double)draw; r[F_NTAIL] = (double)s.rad.ntail; r[F_RADP] = s.rad.rad_proton_this_ev ? 1.0 : 0.0;
  r[F_TX] = s.tx; r[F_TY] = s.ty; r[F_TZ] = s.tz; r[F_RASTERY] = s.rastery;
  r[F_ELOSS0] = s.Eloss[0]; r[F_ELOSS1] = s.Eloss[1]; r[F_ELOSS2] = s.Eloss[2]; r[F_COULOMB] = s.Coulomb;
  r[F_TEFF0] = s.teff[0]; r[F_TEFF1] = s.teff[1]; r[F_TEFF2] = s.teff[2]; r[F_GENW] = s.gen_weight;
  r[F_OEIN] = s.o_Ein; r[F_OEE] = s.o_eE; r[F_OEDELTA] = s.o_edelta; r[F_OPE] = s.o_pE;
  r[F_OPP] = s.o_pP; r[F_OPDELTA] = s.o_pdelta; r[F_VPYP] = s.v_pyptar; r[F_VPXP] = s.v_pxptar;
  r[F_VEYP] = s.v_eyptar; r[F_VEXP] = s.v_exptar; r[F_VEDELTA] = s.v_edelta; r[F_VPDELTA] = s.v_pdelta;
  r[F_VEIN] = s.v_Ein; r[F_VEE] = s.v_eE; r[F_VETHETA] = s.v_etheta; r[F_VQ2] = s.v_Q2;
  r[F_VPE] = s.v_pE; r[F_VPP] = s.v_pP; r[F_VEM] = s.v_Em; r[F_VPM] = s.v_Pm;
  r[F_UEX] = s.uex; r[F_UEY] = s.uey; r[F_UEZ] = s.uez; r[F_HARDCOR] = s.rad.hardcorfac;
  r[F_UPX] = s.upx; r[F_UPY] = s.upy; r[F_UPZ] = s.upz; r[F_VTREC] = s.v_Trec;
  r[F_RC_CEXT0] = s.rad.c_ext0; r[F_RC_GEXT] = s.rad.g_ext; r[F_RC_G1] = s.rad.g[1]; r[F_RC_G2] = s.rad.g[2];
  r[F_RC_G4] = s.rad.g[4]; r[F_RC_BT0] = s.rad.bt[0]; r[F_RC_BT1] = s.rad.bt[1]; r[F_RD_WHICH] = (double)gr.which;
  r[F_RD_EMIN] = gr.emin; r[F_RD_EMAX] = gr.emax; r[F_RD_EG] = gr.eg; r[F_RD_BW] = gr.bw;
  r[F_JAC] = s.jacobian; r[F_EINSHIFT] = s.Ein_shift; r[F_EESHIFT] = s.Ee_shift; r[F_MTREC] = s.Trec;
  r[F_EG0] = s.rad.Egamma_used[0]; r[F_EG1] = s.rad.Egamma_used[1]; r[F_EG2] = s.rad.Egamma_used[2]; r[F_VEPHI] = s.v_ephi;
  r[F_VPTHETA] = s.v_ptheta; r[F_VPPHI] = s.v_pphi;
  const bool hm = g.heavy || g.meson;
  r[F_VNU] = hm ? s.v_nu : 0.0; r[F_VQ] = hm ? s.v_q : 0.0;
  r[F_UQX] = hm ? s.uqx : 0.0; r[F_UQY] = hm ? s.uqy : 0.0; r[F_UQZ] = hm ? s.uqz : 0.0; r[F_MEPS] = g.meson ? s.m_eps : 0.0;
  r[F_MTHPQ] = g.meson ? s.m_thpq : 0.0; r[F_MPHIPQ] = g.meson ? s.m_phipq : 0.0; r[F_MT] = g.meson ? s.m_t : 0.0;
  r[F_MW] = g.meson ? s.m_W : 0.0;
  const bool fm = g.meson && (g.semi || g.fermi);
  r[F_ZHAD] = (g.meson && g.semi) ? s.v_zhad : 0.0; r[F_PT2] = (g.meson && g.semi) ? s.v_pt2 : 0.0;
  r[F_PFER] = fm ? s.pfer : 0.0; r[F_EFER] = fm ? s.efer : 0.0;
  r[F_PFERX] = fm ? s.pferx : 0.0; r[F_PFERY] = fm ? s.pfery : 0.0; r[F_PFERZ] = fm ? s.pferz : 0.0; r[F_RASTERX] = s.rasterx;
  r[F_OPYP] = g.rho ? s.o_pyptar : 0.0; r[F_OPXP] = g.rho ? s.o_pxptar : 0.0;
  r[F_RHOMASS] = g.xtra ? s.rho_mass : 0.0; r[F_RHOTHETA] = g.xtra ? s.rho_theta : 0.0;
  // the groups past F_VPPHI only matter to the reactions that fill them
  const int n_groups = g.xtra ? F_GEN_END / 4 : (hm || g.field) ? F_OPYP / 4 : (F_VQ + 1) / 4;
#if SIMC_STATE_AOS
  double* rec = S.base + (long long)slot * kStateStride;
#pragma unroll
  for (int k = 0; k < F_GEN_END / 4; ++k)
    if (k < n_groups) st_v4(rec + 4 * k, r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
#else
#pragma unroll
  for (int k = 0; k < F_GEN_END; ++k)
    if (k < 4 * n_groups) S.st(k, slot, r[k]);
#endif
  if (A.record_mode) {      // only the per-try records read these
    S.st(F_STAGE, slot, ok ? 1.0 : 0.0); S.st(F_STOP_P, slot, -1.0); S.st(F_STOP_E, slot, -1.0);
  }
}

// What a try brings into its second pass through complete_ev: everything the first pass left in `vertex`,
// `main` and /radccom/ (complete_ev overwrites what it derives; a try that fails half-way keeps the rest for
// the geni histograms, as in the one-pass reference).
__device__ __forceinline__ void load_event(const LoopArgs& A, const GenFlags& g, unsigned slot, EventState& s, GenRad& gr) {
  const StateBuf& S = A.st;
  double r[F_GEN_END];
  const bool hm = g.heavy || g.meson;
  const int n_groups = g.xtra ? F_GEN_END / 4 : (hm || g.field) ? F_OPYP / 4 : (F_VQ + 1) / 4;
#if SIMC_STATE_AOS
  const double* rec = S.base + (long long)slot * kStateStride;
#pragma unroll
  for (int k = 0; k < F_GEN_END / 4; ++k) {
    if (k < n_groups) ld_v4(rec + 4 * k, r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
    else { r[4 * k] = 0.0; r[4 * k + 1] = 0.0; r[4 * k + 2] = 0.0; r[4 * k + 3] = 0.0; }
  }
#else
#pragma unroll
  for (int k = 0; k < F_GEN_END; ++k) r[k] = k < 4 * n_groups ? S.ld(k, slot) : 0.0;
#endif
  s.tx = r[F_TX]; s.ty = r[F_TY]; s.tz = r[F_TZ]; s.rastery = r[F_RASTERY];
  s.Eloss[0] = r[F_ELOSS0]; s.Eloss[1] = r[F_ELOSS1]; s.Eloss[2] = r[F_ELOSS2];
  s.teff[0] = r[F_TEFF0]; s.teff[1] = r[F_TEFF1]; s.teff[2] = r[F_TEFF2];
  s.Coulomb = r[F_COULOMB];
  s.gen_weight = r[F_GENW]; s.jacobian = r[F_JAC]; s.Ein_shift = r[F_EINSHIFT];
  s.Ee_shift = r[F_EESHIFT]; s.Trec = r[F_MTREC];
  s.v_Ein = r[F_VEIN]; s.v_eE = r[F_VEE]; s.v_edelta = r[F_VEDELTA];
  s.v_eyptar = r[F_VEYP]; s.v_exptar = r[F_VEXP]; s.v_etheta = r[F_VETHETA];
  s.v_ephi = r[F_VEPHI];
  s.v_pE = r[F_VPE]; s.v_pP = r[F_VPP]; s.v_pdelta = r[F_VPDELTA];
  s.v_pyptar = r[F_VPYP]; s.v_pxptar = r[F_VPXP]; s.v_ptheta = r[F_VPTHETA];
  s.v_pphi = r[F_VPPHI]; s.v_Q2 = r[F_VQ2];
  s.v_Em = r[F_VEM]; s.v_Pm = r[F_VPM]; s.v_Trec = r[F_VTREC];
  s.uex = r[F_UEX]; s.uey = r[F_UEY]; s.uez = r[F_UEZ];
  s.upx = r[F_UPX]; s.upy = r[F_UPY]; s.upz = r[F_UPZ];
  s.rad.Egamma_used[0] = r[F_EG0]; s.rad.Egamma_used[1] = r[F_EG1]; s.rad.Egamma_used[2] = r[F_EG2];
  s.rad.ntail = (int)r[F_NTAIL];
  s.rad.rad_proton_this_ev = r[F_RADP] != 0.0; s.rad.hardcorfac = r[F_HARDCOR];
  s.rad.c_ext0 = r[F_RC_CEXT0]; s.rad.g_ext = r[F_RC_GEXT]; s.rad.g[1] = r[F_RC_G1];
  s.rad.g[2] = r[F_RC_G2]; s.rad.g[4] = r[F_RC_G4]; s.rad.bt[0] = r[F_RC_BT0];
  s.rad.bt[1] = r[F_RC_BT1];
  gr.emin = r[F_RD_EMIN]; gr.emax = r[F_RD_EMAX]; gr.eg = r[F_RD_EG]; gr.bw = r[F_RD_BW];
  gr.which = (int)r[F_RD_WHICH];
  s.v_nu = r[F_VNU]; s.v_q = r[F_VQ]; s.uqx = r[F_UQX]; s.uqy = r[F_UQY]; s.uqz = r[F_UQZ];
  s.m_eps = r[F_MEPS]; s.m_thpq = r[F_MTHPQ]; s.m_phipq = r[F_MPHIPQ]; s.m_t = r[F_MT]; s.m_W = r[F_MW]; s.m_tmin = 0;
  s.v_zhad = r[F_ZHAD]; s.v_pt2 = r[F_PT2];
  s.pfer = r[F_PFER]; s.pferx = r[F_PFERX]; s.pfery = r[F_PFERY]; s.pferz = r[F_PFERZ];
  s.efer = (g.meson && (g.semi || g.fermi)) ? r[F_EFER] : A.cfg->targ.Mtar_struck;
  s.rasterx = r[F_RASTERX];
  s.o_pyptar = r[F_OPYP]; s.o_pxptar = r[F_OPXP]; s.rho_mass = r[F_RHOMASS]; s.rho_theta = r[F_RHOTHETA];
}

__device__ __forceinline__ void gen_shared_init(const LoopArgs& A, unsigned (*h_geni)[SIMC_NHIST], MatTable& mt_s) {
  for (int i = threadIdx.x; i < SIMC_H_PER_SET * SIMC_NHIST; i += kGenBlock) (&h_geni[0][0])[i] = 0u;
  for (int i = threadIdx.x; i < (int)(sizeof(MatTable) / sizeof(double)); i += kGenBlock)
    ((double*)&mt_s)[i] = ((const double*)&A.mt)[i];
  __syncthreads();
}
__device__ __forceinline__ void gen_shared_flush(const LoopArgs& A, unsigned (*h_geni)[SIMC_NHIST]) {
  __syncthreads();
  for (int i = threadIdx.x; i < SIMC_H_PER_SET * SIMC_NHIST; i += kGenBlock) {
    const unsigned v = (&h_geni[0][0])[i];
    if (v) atomicAdd(&A.acc->hist_n[2][0][0] + i, (unsigned long long)v);
  }
}

// First pass: one thread per try.  generate, complete_ev, and generate_rad up to the photon energy of the tail
// that radiates.  A try whose incoming electron radiated (a third of them) goes to list kRegenList for its
// second pass through complete_ev on compacted warps; all others are complete here.
__global__ void __launch_bounds__(kGenBlock, SIMC_GEN_MIN_BLOCKS) k_generate(LoopArgs A) {
  __shared__ unsigned h_geni[SIMC_H_PER_SET][SIMC_NHIST];
  // the out-of-line energy-loss routines take the material table by reference: give them shared memory,
  // not a generic pointer into the kernel-parameter bank
  __shared__ MatTable mt_s;
#if SIMC_CFG_SMEM & 1
  __shared__ simc_run_config cfg_s;
  cfg_to_shared(cfg_s, A.cfg, kGenBlock);
  const simc_run_config& cfg = cfg_s;
#else
  const simc_run_config& cfg = *A.cfg;
#endif
  gen_shared_init(A, h_geni, mt_s);
  const GenFlags g = gen_flags(cfg);
  const long long stride = (long long)gridDim.x * kGenBlock;
  for (long long i0 = (long long)blockIdx.x * kGenBlock; i0 < A.n_tries; i0 += stride) {
    const long long i = i0 + threadIdx.x;
    const bool active = i < A.n_tries;
    bool ok = false;
    EventState s;
    GenRad gr;
    DevRng rng;
    rng.init((unsigned long long)(A.first_try + (active ? i : 0)), 0u, 0u);
    s.v_pdelta = 0; s.v_pyptar = 0; s.v_pxptar = 0; s.v_edelta = 0; s.v_Pm = 0; s.v_Em = 0;
    s.v_eyptar = 0; s.v_exptar = 0; s.tz = 0;
    // every thread of the CTA walks through the generation code (SIMC_PHASE); the reaction is a run constant
    // (fields only meson / nuclear reactions use: H(e,e'p) neither reads them nor stores them, store_event, and
    //  leaving them untouched keeps a quarter of the local frame's lines clean)
    if (g.meson || g.heavy) {
      s.pfer = 0; s.pferx = 0; s.pfery = 0; s.pferz = 0; s.efer = cfg.targ.Mtar_struck; s.v_zhad = 0; s.v_pt2 = 0;
      s.m_eps = 0; s.m_thpq = 0; s.m_phipq = 0; s.m_t = 0; s.m_W = 0; s.m_tmin = 0;
      s.o_pyptar = 0; s.o_pxptar = 0; s.rho_mass = 0; s.rho_theta = 0;
    }
    s.rasterx = 0;
    // (tables by value: a reference into the kernel parameters handed to an out-of-line function would make the
    //  compiler copy all of LoopArgs to every thread's stack -- +600 bytes of frame, +10 % kernel time)
    if (g.meson) ok = generate_meson_first(cfg, mt_s, A.pfm, A.sf, rng, GaussFn(), s, active, gr);
    else if (g.heavy) ok = generate_heavy_first(cfg, mt_s, rng, GaussFn(), s, active, gr);
    else ok = generate_hyd_elast_first(cfg, mt_s, rng, GaussFn(), s, active, gr);
    const bool carry = active && ok && gr.which == 1;       // radc.f:324: complete_ev once more
    if (!carry) ok = generate_finalize(cfg, s, ok);
    // event.f:420-422.  (The reference calls rho_decay whatever generate_rad returned; for a failed try that only
    // burns random numbers of a stream nobody reads again.)
    if (g.rho && !carry && ok) ok = rho_decay(cfg, rng, s);
    if (active && !carry) geni_hist(cfg, h_geni, s);
    const bool want_slot = active && (ok || carry || A.record_mode);
    const unsigned slot = warp_append(&A.counts[0], want_slot);
    if (want_slot) store_event(A, g, slot, i, rng.draw, s, gr, ok && !carry, carry);
    const unsigned pos = warp_append(&A.counts[1], active && ok && !carry);
    if (active && ok && !carry) A.lists[0 * A.st.cap + pos] = slot;
    const unsigned pos2 = warp_append(&A.counts[1 + kRegenList], carry);
    if (carry) A.lists[(long long)kRegenList * A.st.cap + pos2] = slot;
  }
  gen_shared_flush(A, h_geni);
  if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(&A.acc->counters[0], (unsigned long long)A.n_tries);
}

// Second pass (generate_rad, radc.f:324): the tries of list kRegenList go through complete_ev again with the
// radiated beam energy, on full warps.
__global__ void __launch_bounds__(kGenBlock, SIMC_GEN_MIN_BLOCKS) k_regen(LoopArgs A) {
  __shared__ unsigned h_geni[SIMC_H_PER_SET][SIMC_NHIST];
  __shared__ MatTable mt_s;
#if SIMC_CFG_SMEM & 1
  __shared__ simc_run_config cfg_s;
  cfg_to_shared(cfg_s, A.cfg, kGenBlock);
  const simc_run_config& cfg = cfg_s;
#else
  const simc_run_config& cfg = *A.cfg;
#endif
  gen_shared_init(A, h_geni, mt_s);
  const GenFlags g = gen_flags(cfg);
  const StateBuf& S = A.st;
  const unsigned n_in = A.counts[1 + kRegenList];
  const unsigned* in_list = A.lists + (long long)kRegenList * A.st.cap;
  const long long stride = (long long)gridDim.x * kGenBlock;
  for (long long i0 = (long long)blockIdx.x * kGenBlock; i0 < n_in; i0 += stride) {
    const long long i = i0 + threadIdx.x;
    const bool active = i < n_in;
    const unsigned slot = active ? in_list[i] : 0u;
    const long long itry = (long long)S.ld(F_TRY, slot);
    EventState s;
    GenRad gr;
    load_event(A, g, slot, s, gr);
    DevRng rng;
    rng.init((unsigned long long)(A.first_try + itry), 0u, (unsigned)S.ld(F_DRAW, slot));
    bool ok;
    if (g.meson) ok = generate_meson_second(cfg, mt_s, rng, GaussFn(), s, active);
    else if (g.heavy) ok = generate_heavy_second(cfg, mt_s, rng, GaussFn(), s, active);
    else ok = generate_hyd_elast_second(cfg, mt_s, rng, GaussFn(), s, active);
    // rad_flag = 2, 3: the tails behind tail 1 (radc.f:354-455)
    if (cfg.rad_flag >= 2) ok = (active && ok) ? generate_rad_basis_rest(cfg, rng, s, gr, true) : false;
    ok = generate_finalize(cfg, s, ok);
    if (g.rho && ok) ok = rho_decay(cfg, rng, s);
    if (active) {
      geni_hist(cfg, h_geni, s);
      store_event(A, g, slot, itry, rng.draw, s, gr, ok, false);
    }
    const unsigned pos = warp_append(&A.counts[1], active && ok);
    if (active && ok) A.lists[0 * A.st.cap + pos] = slot;
  }
  gen_shared_flush(A, h_geni);
}

// peaked_rad_weight (radc.f:523-646) and main%gen_weight = gen_weight * rad_weight / hardcorfac (radc.f:518) for
// the tries of list `list_idx`: the survivors of both arms in a run, every generated try in record mode.
__global__ void __launch_bounds__(kFinBlock, SIMC_FIN_MIN_BLOCKS) k_radw(LoopArgs A, int list_idx) {
#if SIMC_CFG_SMEM & 2
  __shared__ simc_run_config cfg_s;
  cfg_to_shared(cfg_s, A.cfg, kFinBlock);
  __syncthreads();
  const simc_run_config& cfg = cfg_s;
#else
  const simc_run_config& cfg = *A.cfg;
#endif
  const StateBuf& S = A.st;
  const unsigned n_in = A.counts[1 + list_idx];
  const unsigned* in_list = A.lists + (long long)list_idx * A.st.cap;
  for (long long i = (long long)blockIdx.x * kFinBlock + threadIdx.x; i < n_in; i += (long long)gridDim.x * kFinBlock) {
    const unsigned slot = in_list[i];
    const int which = (int)S.ld(F_RD_WHICH, slot);
    double rad_weight = 1;
    if (which == 4) {               // rad_flag = 2, 3: the weight was finished at generation (generate_rad_basis_*)
      rad_weight = S.ld(F_RD_BW, slot);
    } else if (which) {
      RadEvDev R;
      R.c_ext0 = S.ld(F_RC_CEXT0, slot); R.g_ext = S.ld(F_RC_GEXT, slot); R.g[1] = S.ld(F_RC_G1, slot);
      R.g[2] = S.ld(F_RC_G2, slot); R.g[4] = S.ld(F_RC_G4, slot); R.bt[0] = S.ld(F_RC_BT0, slot); R.bt[1] = S.ld(F_RC_BT1, slot);
      R.rad_proton_this_ev = S.ld(F_RADP, slot) != 0.0;
      VertexKin v;
      v.Ein = S.ld(F_VEIN, slot); v.eE = S.ld(F_VEE, slot); v.eP = v.eE; v.etheta = S.ld(F_VETHETA, slot);
      v.pE = S.ld(F_VPE, slot); v.pP = S.ld(F_VPP, slot);
      v.uex = S.ld(F_UEX, slot); v.uey = S.ld(F_UEY, slot); v.uez = S.ld(F_UEZ, slot);
      v.upx = S.ld(F_UPX, slot); v.upy = S.ld(F_UPY, slot); v.upz = S.ld(F_UPZ, slot);
      rad_weight = peaked_rad_weight(cfg, R, v, S.ld(F_RD_EG, slot), S.ld(F_RD_EMIN, slot), S.ld(F_RD_EMAX, slot), S.ld(F_RD_BW, slot));
    }
    S.st(F_GENW, slot, S.ld(F_GENW, slot) * rad_weight / S.ld(F_HARDCOR, slot));
  }
}

#ifndef SIMC_ARM_MIN_BLOCKS
#define SIMC_ARM_MIN_BLOCKS 4
#endif
// ---- stages 2,3: the two arms ----------------------------------------------------------------
// WHICH = 1: hadron arm (simc.f:1374-1645), WHICH = 0: electron arm (simc.f:1647-1846).
// Each arm runs as a chain of kernels with the survivors compacted in between (kernels.h: ArmSchedule; the host
// cuts the program): SEG 0 = target multiple scattering, SP quantities, TRANSPORT coordinates, then ops
// [A.op_begin, A.op_end) (the entrance apertures, where most rejected tracks die after almost no arithmetic; empty
// when a compiled stretch takes over from op 0); SEG 2 = the interpreter on a stretch of ops; SEG 1 = the rest of the
// program (always the whole hut), reconstruction, and the arm's recon quantities.  Compiled stretches (mapgen.h)
// run between them on the same track fields (F_TK_*).
// SEG_ = 3: SEG 0 built with the collimator stepping (using_HMScoll / using_SHMScoll).
// track_to_tgt (trg_track.f:738-877; field.cuh has the single-track form) for the lanes of a warp at once: the arm's
// reconstruction map is evaluated by the whole warp (eval_poly), so every lane walks the iteration as long as one
// lane needs it.  ALL 32 LANES MUST CALL; `on`: the lane has a reconstructed track.  `orec` is the arm's OP_RECON op.
__device__ __forceinline__ bool track_to_tgt_warp(const FieldDev& F, int k, const ArmDev* arm, const ArmOp* orec, const ArmResult& res,
                                                  double* pw, unsigned ring, bool on, double& delta, double& y, double& dx, double& dy,
                                                  double frx, double fry, double mom, double mass, double ctheta, double stheta) {
  using namespace fielddetail;
  const double cc = 29.9792458;
  bool ok = on;
  double xx = -fry;
  double vel = fabs(mom) / sqrt(mom * mom + mass * mass) * cc;
  double eng = sign1(mom) * sqrt(mom * mom + mass * mass);
  const double mom_0 = mom / (1.e0 + delta / 100.e0);
  FieldState vT, vTx;
  vT.x = vT.y = vT.z = vT.vx = vT.vy = 0.0; vT.vz = 1.0;
  auto start = [&](double x0) {
    vT.x = x0 + 100. * dx;
    vT.y = y + 100. * dy;
    vT.z = 100.;
    vT.vz = vel / sqrt(1 + dy * dy + dx * dx);
    vT.vx = dx * vT.vz;
    vT.vy = dy * vT.vz;
  };
  if (on) {
    start(-fry);
    ok = track_to_plane(F, k, vT, eng, 1., 0., -ctheta, stheta, frx, ok);
  }
  int n = 0;
  double delx = 1.;
  // the focal-plane track the reconstruction sees (mc_hms.f:421-424); the HRS shift of y_fp is applied afterwards
  const double yfp = res.y_fp + orec->a;
  for (;;) {
    const bool go = on && (delx > .0001) && (n < 10) && ok;
    if (!__any_sync(0xffffffffu, go)) break;
    if (go) {
      delx = fabs(-fry - vT.x);
      vTx = vT;
      vTx.x = -fry;
      ok = track_to_plane(F, k, vT, eng, 1., 0., 0., 1., 0., ok);
      ok = track_to_plane(F, k, vTx, eng, 1., 0., 0., 1., 0., ok);
      xx = xx + fmin(1., fmax(-1., (vTx.x - vT.x)));
    }
    double hut[5];
    hut[0] = res.x_fp / 100.; hut[1] = res.dx_fp; hut[2] = yfp / 100.; hut[3] = res.dy_fp; hut[4] = xx / 100.;
    if (fabs(hut[4]) <= 1.e-30) hut[4] = 1.e-30;
    if (orec->i0) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (fabs(hut[i]) <= 1.e-30) hut[i] = 1.e-30;
    }
    double sum[4];
    eval_poly<4>(arm->tab.rec, arm->tab.recs, hut, pw, ring, sum);
    if (go) {
      dx = sum[0]; y = sum[1] * 100.; dy = sum[2]; delta = sum[3] * 100.;
      mom = mom_0 * (1.e0 + delta / 100.e0);
      vel = fabs(mom) / sqrt(mom * mom + mass * mass) * cc;
      eng = sign1(mom) * sqrt(mom * mom + mass * mass);
      start(xx);
      ok = track_to_plane(F, k, vT, eng, 1., 0., -ctheta, stheta, frx, ok);
      n = n + 1;
    }
  }
  if (on) {
    if (delx > .2) ok = false;
    dy = vT.vy / vT.vz;
    dx = vT.vx / vT.vz;
    y = vT.y;
  }
  return ok;
}

// SEG_ = 4, 5: SEG 0 / SEG 1 built with the tracking through the polarised target's field (using_tgt_field):
// track_from_tgt in front of the arm (simc.f:1425-1432, 1693-1700), track_to_tgt behind its reconstruction (:1573-1587).
#ifndef SIMC_HUT_MIN_BLOCKS
#define SIMC_HUT_MIN_BLOCKS 5
#endif
// SEG_ = 8: SEG 0 of a schedule whose first stretch is compiled (no op of the program runs here: target multiple
// scattering, SP quantities, TRANSPORT coordinates only).  It and the tail (7) need neither the power table nor the
// record ring, are bound by the latency of their gathers, and are launched without dynamic shared memory at
// SIMC_LIGHT_MIN_BLOCKS CTAs per SM.
#ifndef SIMC_LIGHT_MIN_BLOCKS
#define SIMC_LIGHT_MIN_BLOCKS 6
#endif
template <int WHICH, int SEG_>
__global__ void __launch_bounds__(kBlock, SEG_ == 6 ? SIMC_HUT_MIN_BLOCKS : (SEG_ == 7 || SEG_ == 8) ? SIMC_LIGHT_MIN_BLOCKS : SIMC_ARM_MIN_BLOCKS)
k_arm(LoopArgs A, const __grid_constant__ ArmDev arm_c) {
  constexpr int SEG = (SEG_ == 3 || SEG_ == 4 || SEG_ == 8) ? 0 : (SEG_ == 5 || SEG_ == 6 || SEG_ == 7) ? 1 : SEG_;
  constexpr bool kNoOps = SEG_ == 8;
  constexpr bool kColl = SEG_ == 3;
  constexpr bool kField = SEG_ == 4 || SEG_ == 5;
  // SEG_ = 6, 7: SEG 1 in two kernels around the compiled reconstruction map (kernels.h: ARM_STAGE_HUT / _TAIL).
  // 6 runs the hut up to OP_RECON and leaves the fitted focal-plane track in the track rows; the generated kernel
  // turns them into the reconstructed target quantities (mapgen.cpp: OP_RECON); 7 makes the arm's recon quantities.
  constexpr bool kHutOnly = SEG_ == 6;
  constexpr bool kTail = SEG_ == 7;
  extern __shared__ double pw_s[];
  __shared__ unsigned s_stop[SIMC_NSTOP];
  __shared__ unsigned s_calls[48];
  __shared__ MatTable mt_s;
  for (int i = threadIdx.x; i < (int)(sizeof(MatTable) / sizeof(double)); i += kBlock) ((double*)&mt_s)[i] = ((const double*)&A.mt)[i];
  for (int i = threadIdx.x; i < SIMC_NSTOP; i += kBlock) s_stop[i] = 0u;
  for (int i = threadIdx.x; i < 48; i += kBlock) s_calls[i] = 0u;
#if SIMC_CFG_SMEM & 4
  __shared__ simc_run_config cfg_s;
  cfg_to_shared(cfg_s, A.cfg, kBlock);
  __syncthreads();
  const simc_run_config& cfg = cfg_s;
#else
  __syncthreads();
  const simc_run_config& cfg = *A.cfg;
#endif
  const unsigned ring = (unsigned)__cvta_generic_to_shared(pw_s) + (unsigned)kPowBytes + (threadIdx.x >> 5) * kRingBytesPerWarp;
  const StateBuf& S = A.st;
  const unsigned n_in = A.counts[1 + A.in_idx];
  const unsigned* in_list = A.lists + (long long)A.in_idx * A.st.cap;
  unsigned* out_list = A.lists + (long long)A.out_idx * A.st.cap;
  unsigned* out_count = &A.counts[1 + A.out_idx];
  const simc_spectrometer& sp = WHICH == 1 ? cfg.spec_p : cfg.spec_e;
  const ArmDev* arm = &arm_c;        // program + map directory live in the kernel's constant bank
  const int arm_id = WHICH == 1 ? cfg.hadron_arm : cfg.electron_arm;
  const bool use_mc = WHICH == 1 ? cfg.using_P_arm_montecarlo != 0 : cfg.using_E_arm_montecarlo != 0;
  const double Mh2 = detected_Mh2(cfg);
  ArmFlags f;
  f.ms_flag = cfg.mc_smear != 0; f.wcs_flag = cfg.mc_smear != 0;
  f.decay_flag = WHICH == 1 ? cfg.doing_decay != 0 : false;
  f.using_coll = arm_id == 1 ? cfg.using_HMScoll != 0 : (arm_id == 5 ? cfg.using_SHMScoll != 0 : false);
  // op range of this launch (the host cuts the program into stages, kernels.h: ArmSchedule)
  const int op_begin = use_mc ? A.op_begin : 0, op_end = use_mc ? A.op_end : 0;
  const long long stride = (long long)gridDim.x * kBlock;
  for (long long i0 = (long long)blockIdx.x * kBlock; i0 < n_in; i0 += stride) {
    const long long i = i0 + threadIdx.x;
    const bool active = i < n_in;
    const unsigned slot = active ? in_list[i] : 0u;
    DevRng rng;
    rng.init((unsigned long long)(A.first_try + (long long)S.ld(F_TRY, slot)), 0u, (unsigned)S.ld(F_DRAW, slot));
    TrackDev t;
    ArmResult res;
    arm_result_clear(res);
    HutState hs;
#pragma unroll
    for (int k = 0; k < 12; ++k) { hs.xdc[k] = 0.f; hs.ydc[k] = 0.f; }
    hs.scincount = 0;
    double fry = 0.0;
    bool alive = active;
    bool ok = false;
    if (SEG == 0) {
      ArmEntry en;
      en.sp_delta = en.sp_yptar = en.sp_xptar = en.sp_z = en.x = en.y = en.dx = en.dy = 0.0;
      if (active) {
        // the generation record comes in 32-byte groups (StateField): target position, energy losses, thicknesses,
        // the `orig` energies and the generated angles
        double tx, ty, tz, rastery, el0, el1, el2, coul, tf0, tf1, tf2, genw, oEin, oeE, oed, opE, opP, opd, vpy, vpx;
        S.ld4(F_TX, slot, tx, ty, tz, rastery);
        S.ld4(F_ELOSS0, slot, el0, el1, el2, coul);
        S.ld4(F_TEFF0, slot, tf0, tf1, tf2, genw);
        S.ld4(F_OEIN, slot, oEin, oeE, oed, opE);
        double dang0, dang1, sp_delta, ang0 = 0.0, ang1 = 0.0, o_yptar, o_xptar;
        if (WHICH == 1) {
          S.ld4(F_OPP, slot, opP, opd, vpy, vpx);
          o_yptar = vpy; o_xptar = vpx;
          if (cfg.doing_rho) { o_yptar = S.ld(F_OPYP, slot); o_xptar = S.ld(F_OPXP, slot); }      // the decay pion's
          // beam multiple scattering (simc.f:1365), then the hadron's (simc.f:1379-1399)
          if (cfg.mc_smear) {
            const double teff = tf0, p = oEin;
            const double ts = 13.6 / p / 1. * sqrt(teff) * (1 + 0.088 * m::log10(teff / (1. * 1.)));
            dang0 = ts * gauss1(rng, 3.5);
            dang1 = ts * gauss1(rng, 3.5);
          } else { dang0 = 0.0; dang1 = 0.0; }
          S.st(F_DANG0, slot, dang0); S.st(F_DANG1, slot, dang1);
          if (cfg.using_Eloss) {
            const double d = opE - el2;
            sp_delta = (sqrt(fabs(d * d - Mh2)) - sp.P) / sp.P * 100.;
          } else sp_delta = opd;
          if (cfg.mc_smear) {
            const double beta = opP / opE, teff = tf2;
            const double ts = 13.6 / opP / beta * sqrt(teff) * (1 + 0.088 * m::log10(teff / (beta * beta)));
            ang0 = ts * gauss1(rng, 3.5);
            ang1 = ts * gauss1(rng, 3.5);
          }
        } else {
          double ved, vpd;
          S.ld4(F_VEYP, slot, o_yptar, o_xptar, ved, vpd);
          dang0 = S.ld(F_DANG0, slot); dang1 = S.ld(F_DANG1, slot);
          sp_delta = 100 * (oeE - el1 - coul - sp.P) / sp.P;
          if (cfg.mc_smear) {
            const double teff = tf1;
            const double ts = 13.6 / oeE / 1. * sqrt(teff) * (1 + 0.088 * m::log10(teff / (1. * 1.)));
            ang0 = ts * gauss1(rng, 3.5);
            ang1 = ts * gauss1(rng, 3.5);
          }
        }
        arm_entry(sp, tx, ty, tz, sp_delta, o_yptar + ang0 + dang0, o_xptar + ang1 + dang1 * sp.cos_th, en);
        if (kField) {
          // the position before the drift to z = 0 (simc.f:1405-1421), through the field to the field-free image track
          // (track_from_tgt), then the drift back (:1437-1442); its `ok` is overwritten by the arm's own (:1463)
          double x = -ty, y = -tx * sp.cos_th - tz * sp.sin_th * m::sin(sp.phi), z = tz * sp.cos_th + tx * sp.sin_th * m::sin(sp.phi);
          x = x - sp.off_x; y = y - sp.off_y; z = z - sp.off_z;
          double dx = en.dx, dy = en.dy;
          const double mom = (WHICH == 1 ? cfg.sign_hadron : -1.0) * sp.P * (1 + sp_delta / 100.);
          track_from_tgt(A.field, WHICH == 1 ? 1 : 0, x, y, z, dx, dy, mom, WHICH == 1 ? detected_Mh(cfg) : SIMC_ME);
          x = x - z * dx; y = y - z * dy;
          en.x = x; en.y = y; en.dx = dx; en.dy = dy; en.sp_z = y;
        }
        S.st4(WHICH == 1 ? F_SPP_D : F_SPE_D, slot, en.sp_delta, en.sp_yptar, en.sp_xptar, en.sp_z);
        const double fry_raster = cfg.correct_raster ? -rastery : 0.0;
        fry = (arm_id == 1 || arm_id == 5) ? en.x : fry_raster;    // xtar_init, simc.f:1441,1463
      }
      t.dpps = en.sp_delta; t.xs = en.x; t.ys = en.y; t.dxdzs = en.dx; t.dydzs = en.dy;
      t.m2 = WHICH == 1 ? Mh2 : SIMC_ME * SIMC_ME;
      t.p = sp.P * (1. + t.dpps / 100.);
      t.p_spec = sp.P;
      t.pathlen = 0.0; t.decdist = 0.0; t.mh2_final = Mh2; t.ctau = cfg.ctau;
      t.dflag = false;
      musc_refresh(t);
      if (use_mc) {
        if (active) warp_count(&s_stop[0]);
        if (!kNoOps) run_arm<kColl>(arm, t, rng, f, fry, pw_s + threadIdx.x, ring, res, hs, alive, op_begin, op_end, s_calls, s_stop);
        ok = alive;
      } else {
        ok = active;
      }
      if (active) {
        S.st(F_DRAW, slot, (double)rng.draw);
        if (ok) {
          S.st4(F_TK_XS, slot, t.xs, t.ys, t.dxdzs, t.dydzs);
          S.st4(F_TK_DPP, slot, t.dpps, t.p, t.m2, t.pathlen);
          S.st4(F_TK_DECD, slot, t.decdist, t.dflag ? 1.0 : 0.0, fry, 0.0);
        } else {
          if (A.record_mode) S.st(WHICH == 1 ? F_STOP_P : F_STOP_E, slot, (double)res.stop_code);
          warp_hist_add(s_stop, 2 + res.stop_code < SIMC_NSTOP ? 2 + res.stop_code : -1);
        }
      }
    } else if (SEG == 2) {
      if (active) {
        double dfl, pad;
        S.ld4(F_TK_XS, slot, t.xs, t.ys, t.dxdzs, t.dydzs);
        S.ld4(F_TK_DPP, slot, t.dpps, t.p, t.m2, t.pathlen);
        S.ld4(F_TK_DECD, slot, t.decdist, dfl, fry, pad);
        t.p_spec = sp.P; t.dflag = dfl != 0.0;
      } else {
        t.xs = t.ys = t.dxdzs = t.dydzs = t.dpps = 0.0; t.p = sp.P; t.p_spec = sp.P; t.m2 = Mh2; t.pathlen = 0.0; t.decdist = 0.0; t.dflag = false;
      }
      t.mh2_final = (WHICH == 1) ? t.m2 : Mh2; t.ctau = cfg.ctau;
      musc_refresh(t);
      run_arm(arm, t, rng, f, fry, pw_s + threadIdx.x, ring, res, hs, alive, op_begin, op_end, s_calls);
      ok = alive;
      if (active) {
        S.st(F_DRAW, slot, (double)rng.draw);
        if (ok) {
          S.st4(F_TK_XS, slot, t.xs, t.ys, t.dxdzs, t.dydzs);
          S.st4(F_TK_DPP, slot, t.dpps, t.p, t.m2, t.pathlen);
          S.st4(F_TK_DECD, slot, t.decdist, t.dflag ? 1.0 : 0.0, fry, 0.0);
        } else {
          if (A.record_mode) S.st(WHICH == 1 ? F_STOP_P : F_STOP_E, slot, (double)res.stop_code);
          warp_hist_add(s_stop, 2 + res.stop_code < SIMC_NSTOP ? 2 + res.stop_code : -1);
        }
      }
    } else {
      double rc_delta = 0, rc_yptar = 0, rc_xptar = 0, rc_z = 0.0, path = 0.0, resmult = 0.0;
      if (active) {
        double dfl, pad;
        S.ld4(F_TK_XS, slot, t.xs, t.ys, t.dxdzs, t.dydzs);
        S.ld4(F_TK_DPP, slot, t.dpps, t.p, t.m2, t.pathlen);
        S.ld4(F_TK_DECD, slot, t.decdist, dfl, fry, pad);
        t.p_spec = sp.P; t.dflag = dfl != 0.0;
      } else {
        t.xs = t.ys = t.dxdzs = t.dydzs = t.dpps = 0.0; t.p = sp.P; t.p_spec = sp.P; t.m2 = Mh2; t.pathlen = 0.0; t.decdist = 0.0; t.dflag = false;
      }
      // a hadron that decayed before the collimator carries its daughter's mass (Mh2_final, simulate.inc:92)
      t.mh2_final = (WHICH == 1) ? t.m2 : Mh2; t.ctau = cfg.ctau;
      musc_refresh(t);
      if (kTail) {
        // the generated reconstruction kernel left: xs = y_tgt, dxdzs = dph, dydzs = dth, dpps = delta (mapgen.cpp)
        ok = active;
        rc_delta = t.dpps; rc_yptar = t.dydzs; rc_xptar = t.dxdzs; rc_z = t.xs; path = t.pathlen;
        if (active) warp_hist_add(s_stop, 1);
      } else if (kHutOnly) {
        run_arm<false, true>(arm, t, rng, f, fry, pw_s + threadIdx.x, ring, res, hs, alive, op_begin, op_end, s_calls);
        ok = active && alive;          // reached OP_RECON
        if (active) {
          if (res.reached_hut) warp_count(&s_stop[2]);
          if (!ok) warp_hist_add(s_stop, 2 + res.stop_code < SIMC_NSTOP ? 2 + res.stop_code : -1);
          if (A.record_mode) S.st(WHICH == 1 ? F_STOP_P : F_STOP_E, slot, (double)(ok ? 0 : res.stop_code));
          S.st(F_DRAW, slot, (double)rng.draw);
        }
        if (ok) {
          // the fitted track for the reconstruction map, path length and decay bookkeeping for the rows behind it
          S.st4(F_TK_XS, slot, res.x_fp, res.y_fp, res.dx_fp, res.dy_fp);
          S.st4(F_TK_DPP, slot, t.dpps, t.p, t.m2, t.pathlen);
          const ArmOp* orec = &arm->ops[op_end];
          const double yfp_out = orec->a != 0. ? res.y_fp - orec->a : res.y_fp;      // hrsl/mc_hrsl.f:525: shifted afterwards
          if (WHICH == 1) { S.st(F_FPP_DX, slot, res.dx_fp); S.st(F_FPP_DY, slot, res.dy_fp); }
          if (A.record_mode) {
            if (WHICH == 1) {
              S.st(F_FPP_X, slot, res.x_fp); S.st(F_FPP_Y, slot, yfp_out);
              S.st(F_DECDIST, slot, t.decdist); S.st(F_MH2FINAL, slot, t.mh2_final);
            } else {
              S.st(F_FPE_X, slot, res.x_fp); S.st(F_FPE_DX, slot, res.dx_fp); S.st(F_FPE_Y, slot, yfp_out);
              S.st(F_FPE_DY, slot, res.dy_fp);
            }
          }
        }
      } else if (use_mc) {
        run_arm(arm, t, rng, f, fry, pw_s + threadIdx.x, ring, res, hs, alive, op_begin, op_end, s_calls);
        ok = active && res.ok;
        rc_delta = res.dpp_rec; rc_yptar = res.dth_rec; rc_xptar = res.dph_rec; rc_z = res.y_rec;
        if (kField) {          // simc.f:1573-1587 / :1777-1790
          const ArmOp* orec = arm->ops;
          while (orec->op != OP_RECON && orec->op != OP_END) ++orec;
          double frx = 0.0, fry_r = 0.0;
          if (ok && cfg.correct_raster) { fry_r = -S.ld(F_RASTERY, slot); frx = -S.ld(F_RASTERX, slot); }
          const double pmag = sp.P * (1. + rc_delta / 100.0);
          const double mom = (WHICH == 1 ? cfg.sign_hadron : -1.0) * pmag;
          const double mass = WHICH == 1 ? sqrt(t.m2) : sqrt(SIMC_ME * SIMC_ME);
          const bool right = arm_id == 1 || arm_id == 3;          // HMS, HRS-R; the left arms take -stheta
          const double ctheta = m::cos(sp.theta), stheta = right ? m::sin(sp.theta) : -m::sin(sp.theta);
          ok = track_to_tgt_warp(A.field, WHICH == 1 ? 1 : 0, arm, orec, res, pw_s + threadIdx.x, ring, ok, rc_delta, rc_z, rc_xptar,
                                 rc_yptar, frx, -fry_r, mom, mass, ctheta, stheta);
        }
        path = t.pathlen; resmult = res.resmult;
        if (active) {
          if (res.reached_hut) warp_count(&s_stop[2]);
          // (a try the arm accepted counts as one of its successes even when track_to_tgt then fails: simc.f:1573-1592)
          warp_hist_add(s_stop, res.ok ? 1 : (2 + res.stop_code < SIMC_NSTOP ? 2 + res.stop_code : -1));
        }
      } else {
        ok = active;
        rc_delta = S.ld(WHICH == 1 ? F_SPP_D : F_SPE_D, slot); rc_yptar = S.ld(WHICH == 1 ? F_SPP_Y : F_SPE_Y, slot);
        rc_xptar = S.ld(WHICH == 1 ? F_SPP_X : F_SPE_X, slot);
      }
      if (active && !kHutOnly && !kTail) {
        if (A.record_mode) S.st(WHICH == 1 ? F_STOP_P : F_STOP_E, slot, (double)(ok ? 0 : res.stop_code));
        S.st(F_DRAW, slot, (double)rng.draw);
      }
      if (ok && !kHutOnly) {
        S.st4(WHICH == 1 ? F_RCP_D : F_RCE_D, slot, rc_delta, rc_yptar, rc_xptar, rc_z);
        S.st(WHICH == 1 ? F_FPP_PATH : F_FPE_PATH, slot, path);
        if (WHICH == 1 && !kTail) { S.st(F_FPP_DX, slot, res.dx_fp); S.st(F_FPP_DY, slot, res.dy_fp); }
        if (A.record_mode && !kTail) {
          if (WHICH == 1) {
            S.st(F_FPP_X, slot, res.x_fp); S.st(F_FPP_Y, slot, res.y_fp);
            S.st(F_DECDIST, slot, t.decdist); S.st(F_MH2FINAL, slot, t.mh2_final);
          } else {
            S.st(F_FPE_X, slot, res.x_fp); S.st(F_FPE_DX, slot, res.dx_fp); S.st(F_FPE_Y, slot, res.y_fp);
            S.st(F_FPE_DY, slot, res.dy_fp);
          }
        }
        // recon quantities of this arm, simc.f:1623-1645 / :1820-1846
        double rP = sp.P * (1. + rc_delta / 100.);
        double rE = WHICH == 1 ? sqrt(rP * rP + Mh2) : rP;
        double rth, rph;
        physics_angles(sp.theta, sp.phi, rc_xptar + sp.off_xptar, rc_yptar + sp.off_yptar, rth, rph);
        if (cfg.correct_Eloss) {
          double el, rl;
          trip_thru_target_fixed(cfg.targ, mt_s, WHICH == 1 ? 3 : 2, arm_id, 0.0, rE, rth, WHICH == 1 ? detected_Mh(cfg) : SIMC_ME, 4,
                                 el, rl);
          rE = rE + el;
          if (WHICH == 1) {
            rE = fmax(rE, sqrt(Mh2 + 0.000001));
            rP = sqrt(rE * rE - Mh2);
          }
        }
        if (WHICH == 1) {
          S.st4(F_RP_P, slot, rP, rE, rth, rph);
          if (A.record_mode) S.st(F_STAGE, slot, 2.0);
        } else {
          S.st(F_RE_E, slot, rE); S.st(F_RE_TH, slot, rth); S.st(F_RE_PH, slot, rph);
          if (A.record_mode) S.st(F_STAGE, slot, 3.0);
        }
      }
    }
    __syncwarp();
    const unsigned pos = warp_append(out_count, active && ok);
    if (active && ok) out_list[pos] = slot;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < SIMC_NSTOP; i += kBlock)
    if (s_stop[i]) atomicAdd(&A.acc->stop[WHICH][i], (unsigned long long)s_stop[i]);
  for (int i = threadIdx.x; i < 48; i += kBlock)
    if (s_calls[i]) atomicAdd(&A.acc->transp_calls[WHICH][i], (unsigned long long)s_calls[i]);
}

// ---- calorimeter as the hadron arm (SIMC_ARM_CALO_*): simc.f:1374-1443, 1489-1564, 1595-1645 with calo/mc_calo.f ----
// One kernel does what the three segments of k_arm do for a magnetic spectrometer: target multiple scattering, SP
// quantities, TRANSPORT coordinates at z = 0, then mc_calo -- a field-free drift to the front face and the two
// half-size cuts of the NPS -- once for the hadron or, with doing_pizero, once per decay photon (pizero_decay.f: the
// photons are rebuilt from the two numbers complete_ev drew), and the arm's recon quantities.  mc_calo returns dpp,
// y and the slopes as they came in.  Stop codes: 1 = slit_hor, 2 = slit_vert (caloSTOP_*, calo/struct_calo.inc).
__global__ void __launch_bounds__(kBlock, 4) k_calo(LoopArgs A) {
  using namespace mesondetail;
  __shared__ unsigned s_stop[SIMC_NSTOP];
  __shared__ MatTable mt_s;
  for (int i = threadIdx.x; i < (int)(sizeof(MatTable) / sizeof(double)); i += kBlock) ((double*)&mt_s)[i] = ((const double*)&A.mt)[i];
  for (int i = threadIdx.x; i < SIMC_NSTOP; i += kBlock) s_stop[i] = 0u;
  __syncthreads();
  const simc_run_config& cfg = *A.cfg;
  const StateBuf& S = A.st;
  const unsigned n_in = A.counts[1 + A.in_idx];
  const unsigned* in_list = A.lists + (long long)A.in_idx * A.st.cap;
  unsigned* out_list = A.lists + (long long)A.out_idx * A.st.cap;
  unsigned* out_count = &A.counts[1 + A.out_idx];
  const simc_spectrometer& sp = cfg.spec_p;
  const int arm_id = cfg.hadron_arm;
  const double Mh2 = detected_Mh2(cfg);
  const double h_entr = 30.75, v_entr = 36.9;         // half width / height of the NPS: 30 x 36 blocks of 2.05 cm
  const long long stride = (long long)gridDim.x * kBlock;
  for (long long i0 = (long long)blockIdx.x * kBlock; i0 < n_in; i0 += stride) {
    const long long i = i0 + threadIdx.x;
    const bool active = i < n_in;
    bool ok = false;
    unsigned slot = 0u;
    if (active) {
      slot = in_list[i];
      DevRng rng;
      rng.init((unsigned long long)(A.first_try + (long long)S.ld(F_TRY, slot)), 0u, (unsigned)S.ld(F_DRAW, slot));
      double tx, ty, tz, rastery, el0, el1, el2, coul, tf0, tf1, tf2, genw, oEin, oeE, oed, opE, opP, opd, vpy, vpx;
      S.ld4(F_TX, slot, tx, ty, tz, rastery);
      S.ld4(F_ELOSS0, slot, el0, el1, el2, coul);
      S.ld4(F_TEFF0, slot, tf0, tf1, tf2, genw);
      S.ld4(F_OEIN, slot, oEin, oeE, oed, opE);
      S.ld4(F_OPP, slot, opP, opd, vpy, vpx);
      double dang0 = 0.0, dang1 = 0.0, ang0 = 0.0, ang1 = 0.0, sp_delta;
      if (cfg.mc_smear) {                               // simc.f:1365
        const double teff = tf0, p = oEin;
        const double ts = 13.6 / p / 1. * sqrt(teff) * (1 + 0.088 * m::log10(teff / (1. * 1.)));
        dang0 = ts * gauss1(rng, 3.5);
        dang1 = ts * gauss1(rng, 3.5);
      }
      S.st(F_DANG0, slot, dang0); S.st(F_DANG1, slot, dang1);
      if (cfg.using_Eloss) {
        const double d = opE - el2;
        sp_delta = (sqrt(fabs(d * d - Mh2)) - sp.P) / sp.P * 100.;
      } else sp_delta = opd;
      if (cfg.mc_smear) {
        const double beta = opP / opE, teff = tf2;
        const double ts = 13.6 / opP / beta * sqrt(teff) * (1 + 0.088 * m::log10(teff / (beta * beta)));
        ang0 = ts * gauss1(rng, 3.5);
        ang1 = ts * gauss1(rng, 3.5);
      }
      ArmEntry en;
      arm_entry(sp, tx, ty, tz, sp_delta, vpy + ang0 + dang0, vpx + ang1 + dang1 * sp.cos_th, en);
      S.st4(F_SPP_D, slot, en.sp_delta, en.sp_yptar, en.sp_xptar, en.sp_z);
      warp_hist_add(s_stop, 0);                         // caloSTOP_trials, once per try as for the other arms
      TrackDev t;
      t.dpps = en.sp_delta; t.m2 = Mh2; t.p = sp.P * (1. + t.dpps / 100.); t.p_spec = sp.P;
      t.pathlen = 0.0; t.decdist = 0.0; t.mh2_final = Mh2; t.ctau = cfg.ctau; t.dflag = false;
      double x_fp = 0.0, y_fp = 0.0, dx_fp = 0.0, dy_fp = 0.0;
      int stop_code = 0;
      // one call of mc_calo (calo/mc_calo.f:120-152) for the slopes (dx, dy); false = it stopped on a slit
      auto mc_calo = [&](double dx, double dy, int& code) {
        t.xs = en.x; t.ys = en.y; t.dxdzs = dx; t.dydzs = dy;
        project(t, rng, cfg.drift_to_cal, cfg.doing_decay != 0);
        if (fabs(t.ys) > h_entr) { code = 1; return false; }
        if (fabs(t.xs) > v_entr) { code = 2; return false; }
        x_fp = t.xs; y_fp = t.ys; dx_fp = t.dxdzs; dy_fp = t.dydzs;
        return true;
      };
      double rc_yptar = en.dy, rc_xptar = en.dx;        // what mc_calo hands back: its inputs
      if (cfg.doing_pizero) {
        // pizero_decay.f:37-70 from the two numbers complete_ev drew (rph, rth), then simc.f:1494-1556
        const double Mh = cfg.Mh;
        const double ph = S.ld(F_VPP, slot);
        const double eh = sqrt(ph * ph + Mh * Mh);
        const double beta = ph / eh;
        const double gamma = 1. / sqrt(1. - beta * beta);
        const double rph = S.ld(F_RHOMASS, slot), rth = S.ld(F_RHOTHETA, slot);
        const double er = Mh / 2.0;
        const double pr = sqrt(er * er - 0.0 * 0.0);
        const double pxr1 = pr * m::sin(rth) * m::cos(rph), pyr1 = pr * m::sin(rth) * m::sin(rph), pzr1 = pr * m::cos(rth);
        const double upx = S.ld(F_UPX, slot), upy = S.ld(F_UPY, slot), upz = S.ld(F_UPZ, slot);
        const double bx = -beta * upx, by = -beta * upy, bz = -beta * upz;
        const MV4 g1 = mesondetail::loren(gamma, bx, by, bz, er, pxr1, pyr1, pzr1);
        const MV4 g2 = mesondetail::loren(gamma, bx, by, bz, er, -pxr1, -pyr1, -pzr1);
        const double cth = m::cos(sp.theta), sth = m::sin(sp.theta);
        double ey1, ez1, ey2, ez2;
        if (arm_id == 8) {
          ey1 = g1.y * cth - g1.z * sth; ez1 = g1.y * sth + g1.z * cth;
          ey2 = g2.y * cth - g2.z * sth; ez2 = g2.y * sth + g2.z * cth;
        } else {
          ey1 = g1.y * cth + g1.z * sth; ez1 = -g1.y * sth + g1.z * cth;
          ey2 = g2.y * cth + g2.z * sth; ez2 = -g2.y * sth + g2.z * cth;
        }
        int c1 = 0, c2 = 0;
        const bool ok1 = mc_calo(g1.x / ez1, ey1 / ez1, c1);
        const double xc1 = ok1 ? x_fp : -1.0e10, yc1 = ok1 ? y_fp : -1.0e10;
        const bool ok2 = mc_calo(g2.x / ez2, ey2 / ez2, c2);
        const double xc2 = ok2 ? x_fp : -1.0e10, yc2 = ok2 ? y_fp : -1.0e10;
        ok = cfg.pizero_ngamma == 2 ? (ok1 && ok2) : (ok1 || ok2);
        stop_code = c2 ? c2 : c1;                       // our bookkeeping of a lost pair: the second photon's slit, else the first's
        rc_yptar = ey2 / ez2; rc_xptar = g2.x / ez2;    // dy_P_arm, dx_P_arm of the last call (replaced just below)
        if (A.record_mode) {                            // ntuple columns 54-65 (results_write.f:167-180)
          const double v[12] = {xc1, yc1, g1.e, g1.x, g1.y, g1.z, xc2, yc2, g2.e, g2.x, g2.y, g2.z};
#pragma unroll
          for (int k = 0; k < 12; ++k) S.st(F_NTU0 + 53 + k, slot, v[k]);
        }
      } else {
        ok = mc_calo(en.dx, en.dy, stop_code);
      }
      S.st(F_DRAW, slot, (double)rng.draw);
      if (A.record_mode) S.st(F_STOP_P, slot, (double)(ok ? 0 : stop_code));
      warp_hist_add(s_stop, ok ? 1 : 2 + stop_code);
      if (ok) {
        double rc_delta = en.sp_delta, rc_z = en.y;
        if (cfg.doing_pizero) { rc_yptar = en.sp_yptar; rc_xptar = en.sp_xptar; }       // simc.f:1605-1609
        S.st4(F_RCP_D, slot, rc_delta, rc_yptar, rc_xptar, rc_z);
        S.st(F_FPP_PATH, slot, t.pathlen);
        S.st(F_FPP_DX, slot, dx_fp); S.st(F_FPP_DY, slot, dy_fp);
        if (A.record_mode) {
          S.st(F_FPP_X, slot, x_fp); S.st(F_FPP_Y, slot, y_fp);
          S.st(F_DECDIST, slot, t.decdist); S.st(F_MH2FINAL, slot, t.mh2_final);
          S.st(F_RESFAC, slot, 0.0);
        }
        double rP = sp.P * (1. + rc_delta / 100.);
        double rE = sqrt(rP * rP + Mh2);
        double rth, rph;
        physics_angles(sp.theta, sp.phi, rc_xptar + sp.off_xptar, rc_yptar + sp.off_yptar, rth, rph);
        if (cfg.correct_Eloss) {
          double el, rl;
          trip_thru_target_fixed(cfg.targ, mt_s, 3, arm_id, 0.0, rE, rth, detected_Mh(cfg), 4, el, rl);
          rE = rE + el;
          rE = fmax(rE, sqrt(Mh2 + 0.000001));
          rP = sqrt(rE * rE - Mh2);
        }
        S.st4(F_RP_P, slot, rP, rE, rth, rph);
        if (A.record_mode) S.st(F_STAGE, slot, 2.0);
      }
    }
    __syncwarp();
    const unsigned pos = warp_append(out_count, active && ok);
    if (active && ok) out_list[pos] = slot;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < SIMC_NSTOP; i += kBlock)
    if (s_stop[i]) atomicAdd(&A.acc->stop[1][i], (unsigned long long)s_stop[i]);
}

// ---- stage 4: recon kinematics, weight, accumulation --------------------------------------------
// Per-CTA accumulators in shared memory: every quantity is first reduced over the warp with
// shuffles (integer adds / min / max: exact and order-free), then lane 0 adds it to the CTA's
// copy; the CTA flushes once to the global DevAccum at the end of the kernel.
struct BlockAcc {
  unsigned long long sums[18][2];                    // wt, sigcc, sumerr[8], sumerr2[8]
  unsigned long long hist_w[6][SIMC_NHIST][2];
  unsigned hist_n[9][SIMC_NHIST];                    // gen (7), RECON Em, RECON Pm
  unsigned counters[6];                              // nsuccess, ncontribute, npasscuts, nco_no_rad_proton, unsupported, nonfinite
  long long mins[40], maxs[40];                      // contrib (30 used of 32) + slop (8)
};

__device__ __forceinline__ void warp_add128(unsigned long long* dst, unsigned long long lo, unsigned long long hi) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const unsigned long long olo = __shfl_down_sync(0xffffffffu, lo, off);
    const unsigned long long ohi = __shfl_down_sync(0xffffffffu, hi, off);
    lo += olo;
    hi += ohi + (lo < olo ? 1ULL : 0ULL);
  }
  if ((threadIdx.x & 31u) == 0 && (lo | hi)) {
    const unsigned long long old = atomicAdd(&dst[0], lo);
    const unsigned long long carry = (old + lo < old) ? 1ULL : 0ULL;
    if (hi + carry) atomicAdd(&dst[1], hi + carry);
  }
}
__device__ __forceinline__ void warp_sum128(unsigned long long* dst, bool valid, double x, int qexp) {
  unsigned long long lo = 0, hi = 0;
  if (valid) to_fixed(x, qexp, lo, hi);
  warp_add128(dst, lo, hi);
}
__device__ __forceinline__ void warp_minmax(long long* mn, long long* mx, bool valid, double v) {
  long long lo = valid ? dkey(v) : 0x7fffffffffffffffLL;
  long long hi = valid ? dkey(v) : (long long)0x8000000000000000ULL;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    lo = min(lo, __shfl_down_sync(0xffffffffu, lo, off));
    hi = max(hi, __shfl_down_sync(0xffffffffu, hi, off));
  }
  if ((threadIdx.x & 31u) == 0) { atomicMin(mn, lo); atomicMax(mx, hi); }
}
__device__ __forceinline__ void warp_count_if(unsigned* dst, bool flag) {
  const unsigned m = __ballot_sync(0xffffffffu, flag);
  if ((threadIdx.x & 31u) == 0 && m) atomicAdd(dst, (unsigned)__popc(m));
}

__global__ void __launch_bounds__(kFinBlock, SIMC_FIN_MIN_BLOCKS) k_finish(LoopArgs A) {
  __shared__ BlockAcc B;
  {
    unsigned long long* w = (unsigned long long*)&B;
    for (int i = threadIdx.x; i < (int)(sizeof(BlockAcc) / 8); i += kFinBlock) w[i] = 0ULL;
    __syncthreads();
    for (int i = threadIdx.x; i < 40; i += kFinBlock) { B.mins[i] = 0x7fffffffffffffffLL; B.maxs[i] = (long long)0x8000000000000000ULL; }
  }
#if SIMC_CFG_SMEM & 2
  __shared__ simc_run_config cfg_s;
  cfg_to_shared(cfg_s, A.cfg, kFinBlock);
  const simc_run_config& cfg = cfg_s;
#else
  const simc_run_config& cfg = *A.cfg;
#endif
  __syncthreads();
  const StateBuf& S = A.st;
  DevAccum* acc = A.acc;
  const unsigned n_in = A.counts[1 + 2 * kArmLists];
  const unsigned* in_list = A.lists + (long long)(2 * kArmLists) * A.st.cap;
  const long long stride = (long long)gridDim.x * kFinBlock;
  for (long long i0 = (long long)blockIdx.x * kFinBlock; i0 < n_in; i0 += stride) {
    const long long i = i0 + threadIdx.x;
    const bool active = i < n_in;
    bool success = false, pass_cuts = false, no_rad_p = false, low_w = false;
    double weight = 0, sigcc = 0, rEm = 0, rPm = 0, sigcm1 = 0, johnjac = 0;
    double rec_vals[6] = {0, 0, 0, 0, 0, 0}, gen_vals[7] = {0, 0, 0, 0, 0, 0, 0}, err[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double cv[30], sv[8];
#pragma unroll
    for (int k = 0; k < 30; ++k) cv[k] = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) sv[k] = 0;
    if (active) {
      const unsigned slot = in_list[i];
      // complete_recon_ev for H(e,e'p), event.f:1056-1359
      const double r_Ein = cfg.Ebeam_vertex_ave - cfg.targ.Coulomb_ave;
      const double reE = S.ld(F_RE_E, slot), reth = S.ld(F_RE_TH, slot), reph = S.ld(F_RE_PH, slot);
      const double rpP = S.ld(F_RP_P, slot), rpE = S.ld(F_RP_E, slot), rpth = S.ld(F_RP_TH, slot), rpph = S.ld(F_RP_PH, slot);
      const double reP = reE;
      const double uex = sin(reth) * cos(reph), uey = sin(reth) * sin(reph), uez = cos(reth);
      const double upx = sin(rpth) * cos(rpph), upy = sin(rpth) * sin(rpph), upz = cos(rpth);
      const double nu = r_Ein - reE;
      const double Q2 = 2 * r_Ein * reE * (1 - uez);
      const double q = sqrt(Q2 + nu * nu);
      const double uqx = -reP * uex / q, uqy = -reP * uey / q, uqz = (r_Ein - reP * uez) / q;
      const double W2 = SIMC_MP * SIMC_MP + 2. * SIMC_MP * nu - Q2;
      const double rW = sqrt(fabs(W2)) * W2 / fabs(W2);
      const double Pmx = rpP * upx - q * uqx, Pmy = rpP * upy - q * uqy, Pmz = rpP * upz - q * uqz;
      rPm = sqrt(Pmx * Pmx + Pmy * Pmy + Pmz * Pmz);
      const bool semi = cfg.doing_semi != 0;
      const bool fermi = cfg.doing_deutsemi || cfg.doing_deutpi || cfg.doing_deutkaon || cfg.doing_hepi || cfg.doing_hekaon;
      const bool rho = cfg.doing_rho != 0;
      const bool meson = cfg.doing_pion || cfg.doing_kaon || cfg.doing_delta || semi || rho;
      const bool deut = cfg.doing_deuterium != 0;
      const double Mh2_det = detected_Mh2(cfg);       // COMMON Mh2 when complete_recon_ev runs
      const bool heavy = cfg.doing_heavy != 0 || deut;          // (e,e'p) from a nucleus: deForest, A-1 recoil
      const double rTrec = heavy ? sqrt(rPm * rPm + cfg.targ.Mrec * cfg.targ.Mrec) - cfg.targ.Mrec : 0.0;   // event.f:1349
      // complete_main, event.f:1363-1569
      const double v_Ein = S.ld(F_VEIN, slot), v_eE = S.ld(F_VEE, slot), v_eth = S.ld(F_VETHETA, slot), v_Q2 = S.ld(F_VQ2, slot);
      double sigcc_recon, tgtweight = 1.0, survivalprob = 1.0, SF_weight = 1.0;
      double mv_pfer[4] = {0.0, 0.0, 0.0, 0.0};           // pfer, pferx, pfery, pferz of this event
      if (fermi) {
        mv_pfer[0] = S.ld(F_PFER, slot); mv_pfer[1] = S.ld(F_PFERX, slot); mv_pfer[2] = S.ld(F_PFERY, slot);
        mv_pfer[3] = S.ld(F_PFERZ, slot);
      }
      if (heavy) {
        rEm = nu + cfg.targ.Mtar_struck - rpE - rTrec;
        bool bad = false;
        if (cfg.use_benhar_sf && !deut)
          SF_weight = cfg.targ.Z * cfg.transparency * sf_lookup_diff(A.sf, S.ld(F_VEM, slot), S.ld(F_VPM, slot), bad);
        else                                           // event.f:1402-1428
          SF_weight = theory_sf_weight(A.theory, !deut, S.ld(F_VEM, slot), S.ld(F_VPM, slot));
        low_w = bad;                                   // counted in `unsupported`: the reference would `stop`
        HeavyEv ev;
        ev.Q2 = v_Q2; ev.q = S.ld(F_VQ, slot); ev.nu = S.ld(F_VNU, slot); ev.Pm = S.ld(F_VPM, slot);
        ev.pE = S.ld(F_VPE, slot); ev.pP = S.ld(F_VPP, slot); ev.eE = v_eE; ev.etheta = v_eth;
        ev.uqx = S.ld(F_UQX, slot); ev.uqy = S.ld(F_UQY, slot); ev.uqz = S.ld(F_UQZ, slot);
        ev.upx = S.ld(F_UPX, slot); ev.upy = S.ld(F_UPY, slot); ev.upz = S.ld(F_UPZ, slot);
        ev.Pmx = ev.pP * ev.upx - ev.q * ev.uqx; ev.Pmy = ev.pP * ev.upy - ev.q * ev.uqy; ev.Pmz = ev.pP * ev.upz - ev.q * ev.uqz;
        sigcc = deForest(ev, cfg.Mh2, cfg.deForest_flag);
        HeavyEv rv;
        rv.Q2 = Q2; rv.q = q; rv.nu = nu; rv.Pm = rPm; rv.pE = rpE; rv.pP = rpP; rv.eE = reE; rv.etheta = reth;
        rv.uqx = uqx; rv.uqy = uqy; rv.uqz = uqz; rv.upx = upx; rv.upy = upy; rv.upz = upz;
        rv.Pmx = Pmx; rv.Pmy = Pmy; rv.Pmz = Pmz;
        sigcc_recon = deForest(rv, cfg.Mh2, cfg.deForest_flag);
      } else if (!meson) {
        rEm = nu + cfg.targ.M - rpE - rTrec;
        sigcc = sigep(v_Ein, v_eE, v_eth, v_Q2);
        sigcc_recon = sigep(r_Ein, reE, reth, Q2);
      } else {
        // event.f:1306-1345: missing mass of the undetected system
        rEm = nu + cfg.targ.Mtar_struck - rpE;
        const double mm2 = rEm * rEm - rPm * rPm;
        if (A.record_mode) S.st(F_MM, slot, sqrt(fabs(mm2)) * fabs(mm2) / mm2);
        MesonVertex mv;
        mv.Ein = v_Ein; mv.eE = v_eE; mv.nu = S.ld(F_VNU, slot); mv.q = S.ld(F_VQ, slot); mv.Q2 = v_Q2;
        mv.pP = S.ld(F_VPP, slot); mv.pE = S.ld(F_VPE, slot);
        mv.uqx = S.ld(F_UQX, slot); mv.uqy = S.ld(F_UQY, slot); mv.uqz = S.ld(F_UQZ, slot);
        mv.upx = S.ld(F_UPX, slot); mv.upy = S.ld(F_UPY, slot); mv.upz = S.ld(F_UPZ, slot);
        mv.phi_pq = S.ld(F_MPHIPQ, slot); mv.t = S.ld(F_MT, slot); mv.epsilon = S.ld(F_MEPS, slot);
        mv.pfer = 0.0; mv.pferx = 0.0; mv.pfery = 0.0; mv.pferz = 0.0; mv.efer = cfg.targ.Mtar_struck;
        if (fermi) {
          mv.pfer = S.ld(F_PFER, slot); mv.pferx = S.ld(F_PFERX, slot); mv.pfery = S.ld(F_PFERY, slot);
          mv.pferz = S.ld(F_PFERZ, slot); mv.efer = S.ld(F_EFER, slot);
        }
        MesonWeight mw;
        if (semi) {
          // peepiX reads vertex%theta_pq, which nothing assigns (complete_ev fills main%theta_pq): zero,
          // i.e. cos(theta_pq) = 1 in the reference's jacobian (semi_physics.f:243).  Reproduced.
          SemiVertex sv_;
          sv_.Ein = v_Ein; sv_.eE = v_eE; sv_.nu = mv.nu; sv_.Q2 = v_Q2; sv_.q = mv.q;
          sv_.uqx = mv.uqx; sv_.uqy = mv.uqy; sv_.uqz = mv.uqz;
          sv_.pt2 = S.ld(F_PT2, slot); sv_.zhad = S.ld(F_ZHAD, slot); sv_.theta_pq = 0.0;
          sv_.pfer = mv.pfer; sv_.pferx = mv.pferx; sv_.pfery = mv.pfery; sv_.pferz = mv.pferz; sv_.efer = mv.efer;
          const SemiWeight w_ = peepiX(cfg, A.pdf, A.fdss, sv_, nullptr);
          mw.sigcc = w_.sigcc; mw.sigcm = w_.sighad; mw.davejac = w_.davejac; mw.low_w = w_.bad;
          mw.thetacm = 0.0; mw.phicm = 0.0; mw.wcm = 0.0;
          if (A.record_mode) S.st(F_XFERMI, slot, w_.xfermi);      // ntup%xfermi, semi_physics.f:250
          if (!cfg.doing_decay && !w_.early) survivalprob = semi_survival(cfg, S.ld(F_FPP_PATH, slot), S.ld(F_FPP_DX, slot), S.ld(F_FPP_DY, slot));
        } else if (cfg.doing_pion) {
          mw = peepi(cfg, A.maid, mv);
          // event.f:1464-1490: Delta final states scale the pi-N cross section (empirical coefficients of 2021/2023)
          if (cfg.which_pion == 2) {
            if (cfg.doing_hydpi) mw.sigcc = cfg.doing_pizero ? 0.55 * mw.sigcc : 0.4 * mw.sigcc;
            else if (cfg.doing_deutpi) mw.sigcc = cfg.doing_pizero ? 0.55 * mw.sigcc : 0.4 * mw.sigcc + 0.8 * mw.sigcc;
          } else if (cfg.which_pion == 3) {
            if (cfg.doing_hydpi) mw.sigcc = cfg.doing_pizero ? 0 : 0.55 * mw.sigcc;
            else if (cfg.doing_deutpi) mw.sigcc = cfg.doing_pizero ? 0.99 * mw.sigcc : 0.55 * mw.sigcc + 0.99 * mw.sigcc;
          }
          tgtweight = (cfg.which_pion == 1 || cfg.which_pion == 11) ? cfg.targ.N : cfg.targ.Z;
        } else if (cfg.doing_delta) {
          mw = peedelta(cfg, mv);                      // event.f:1511-1513; tgtweight stays 1
        } else if (rho) {                              // event.f:1515-1518
          mw = peerho(cfg, mv, v_eth);
          johnjac = mw.johnjac;
          tgtweight = cfg.targ.Z + cfg.targ.N;
          if (A.record_mode) S.st(F_MT, slot, mw.t_gev);      // peerho leaves its own t (GeV^2) in main%t
        } else {
          // the Saghai model only feeds ntuple column 54: evaluated where rows / records are produced
          mw = peeK(cfg, mv, A.record_mode ? A.saghai : SaghaiDev{nullptr, 0, 0, 0}, S.ld(F_MTHPQ, slot));
          sigcm1 = mw.sigcm1;
          tgtweight = (cfg.which_kaon == 2 || cfg.which_kaon == 12) ? cfg.targ.N : cfg.targ.Z;
          if (!cfg.doing_decay) survivalprob = kaon_survival(cfg, S.ld(F_FPP_PATH, slot), S.ld(F_FPP_DX, slot), S.ld(F_FPP_DY, slot));
        }
        low_w = mw.low_w;
        sigcc = mw.sigcc;
        sigcc_recon = 1.0;
        if (A.record_mode) {
          S.st(F_THCM, slot, mw.thetacm); S.st(F_PHICM, slot, mw.phicm); S.st(F_SIGCM, slot, mw.sigcm);
          S.st(F_DAVEJAC, slot, mw.davejac); S.st(F_SURV, slot, survivalprob); S.st(F_WCM, slot, mw.wcm);
        }
      }
      if (cfg.using_Coulomb) { const double c = 1.0 + cfg.targ.Coulomb_ave / cfg.Ebeam; sigcc = sigcc * (c * c); }
      weight = SF_weight * S.ld(F_JAC, slot) * S.ld(F_GENW, slot) * sigcc;
      weight = weight * tgtweight;
      if ((cfg.doing_kaon || cfg.doing_semika) && !cfg.doing_decay) weight = weight * survivalprob;
      // pass_cuts, simc.f:229-241 (p-arm upper delta edge uses SPedge%e%delta%max, as written)
      const double red = S.ld(F_RCE_D, slot), rey = S.ld(F_RCE_Y, slot), rex = S.ld(F_RCE_X, slot), rez = S.ld(F_RCE_Z, slot);
      const double rpd = S.ld(F_RCP_D, slot), rpy = S.ld(F_RCP_Y, slot), rpx = S.ld(F_RCP_X, slot), rpz = S.ld(F_RCP_Z, slot);
      pass_cuts = !(red <= (cfg.SPedge_e.delta.min + cfg.slop_MC_e_used[0]) ||
                    red >= (cfg.SPedge_e.delta.max - cfg.slop_MC_e_used[0]) ||
                    rey <= (cfg.SPedge_e.yptar.min + cfg.slop_MC_e_used[1]) ||
                    rey >= (cfg.SPedge_e.yptar.max - cfg.slop_MC_e_used[1]) ||
                    rex <= (cfg.SPedge_e.xptar.min + cfg.slop_MC_e_used[2]) ||
                    rex >= (cfg.SPedge_e.xptar.max - cfg.slop_MC_e_used[2]) ||
                    rpd <= (cfg.SPedge_p.delta.min + cfg.slop_MC_p_used[0]) ||
                    rpd >= (cfg.SPedge_e.delta.max - cfg.slop_MC_p_used[0]) ||
                    rpy <= (cfg.SPedge_p.yptar.min + cfg.slop_MC_p_used[1]) ||
                    rpy >= (cfg.SPedge_p.yptar.max - cfg.slop_MC_p_used[1]) ||
                    rpx <= (cfg.SPedge_p.xptar.min + cfg.slop_MC_p_used[2]) ||
                    rpx >= (cfg.SPedge_p.xptar.max - cfg.slop_MC_p_used[2]));
      success = !(SF_weight <= 0.0);        // event.f:1435: no spectral strength here -> the try does not count
      if (cfg.hard_cuts) {
        if (!pass_cuts) success = false;
        if (cfg.doing_eep && (rEm > cfg.cuts_Em.max)) success = false;
      }
      if (A.record_mode) {      // only the per-try records and the ntuple rows read these
        S.st(F_WEIGHT, slot, weight); S.st(F_SIGCC, slot, sigcc); S.st(F_SIGCC_RECON, slot, sigcc_recon);
        S.st(F_PASSCUTS, slot, pass_cuts ? 1.0 : 0.0); S.st(F_REM, slot, rEm); S.st(F_RPM, slot, rPm); S.st(F_RW, slot, rW);
        S.st(F_STAGE, slot, success ? 4.0 : 3.0);
      }
      if (success && A.record_mode) {
        // ---- ntuple row, results_ntu_write (results_write.f:1-269); complete_recon_ev's remaining
        // quantities (event.f:1150-1300) are only needed here
        const double th2 = m::tan(reth / 2.);
        const double r_eps = 1. / (1. + 2. * (1 + nu * nu / Q2) * (th2 * th2));
        const double r_thpq = m::acos(fmin(1.0, upx * uqx + upy * uqy + upz * uqz));
        const double qx = -uqy, qy = uqx, qz = uqz, px = -upy, py = upx, pz = upz;
        double dummy = sqrt((qx * qx + qy * qy) * (qx * qx + qy * qy + qz * qz));
        const double new_x_x = -qx * qz / dummy, new_x_y = -qy * qz / dummy, new_x_z = (qx * qx + qy * qy) / dummy;
        dummy = sqrt(qx * qx + qy * qy);
        const double new_y_x = qy / dummy, new_y_y = -qx / dummy, new_y_z = 0.0;
        const double p_new_x = px * new_x_x + py * new_x_y + pz * new_x_z;
        const double p_new_y = px * new_y_x + py * new_y_y + pz * new_y_z;
        double r_phipq;
        if ((p_new_x * p_new_x + p_new_y * p_new_y) == 0.) r_phipq = 0.0;
        else r_phipq = m::acos(p_new_x / sqrt(p_new_x * p_new_x + p_new_y * p_new_y));
        if (p_new_y < 0.) r_phipq = 2 * SIMC_PI_D - r_phipq;
        const double oop_x = -uqy, oop_y = uqx;
        const double PmPar = (Pmx * uqx + Pmy * uqy + Pmz * uqz);
        const double PmOop = (Pmx * oop_x + Pmy * oop_y) / sqrt(oop_x * oop_x + oop_y * oop_y);
        const double PmPer = sqrt(fmax(0.e0, rPm * rPm - PmPar * PmPar - PmOop * PmOop));
        double ntu[SIMC_NTUPLE_MAXCOL + 1];
#pragma unroll
        for (int k = 0; k <= SIMC_NTUPLE_MAXCOL; ++k) ntu[k] = 0.0;
        const int ea = cfg.electron_arm;
        const bool e_right = (ea == 1 || ea == 3 || ea == 7);
        const double tz = S.ld(F_TZ, slot);
        const double eb[12] = {red, rey, rex, rez, S.ld(F_FPE_X, slot), S.ld(F_FPE_DX, slot), S.ld(F_FPE_Y, slot),
                               S.ld(F_FPE_DY, slot), S.ld(F_OEDELTA, slot), S.ld(F_VEYP, slot), S.ld(F_VEXP, slot), cfg.spec_e.sin_th};
        const double pb[12] = {rpd, rpy, rpx, rpz, S.ld(F_FPP_X, slot), S.ld(F_FPP_DX, slot), S.ld(F_FPP_Y, slot),
                               S.ld(F_FPP_DY, slot), S.ld(F_OPDELTA, slot), S.ld(rho ? F_OPYP : F_VPYP, slot), S.ld(rho ? F_OPXP : F_VPXP, slot), cfg.spec_p.sin_th};
#pragma unroll
        for (int k = 0; k < 11; ++k) {
          ntu[1 + k] = e_right ? eb[k] : pb[k];
          ntu[13 + k] = e_right ? pb[k] : eb[k];
        }
        ntu[12] = tz * (e_right ? eb[11] : pb[11]);
        ntu[24] = -tz * (e_right ? pb[11] : eb[11]);
        ntu[25] = q / 1000.; ntu[26] = nu / 1000.; ntu[27] = Q2 / 1.e6; ntu[28] = rW / 1000.; ntu[29] = r_eps;
        ntu[30] = rEm / 1000.; ntu[31] = rPm / 1000.; ntu[32] = r_thpq; ntu[33] = r_phipq;
        const double radphot = S.ld(F_EG0, slot) + S.ld(F_EG1, slot) + S.ld(F_EG2, slot);
        const double wfinal = weight;                  // survival probability already applied above
        if (semi || rho) {   // results_write.f:187-230
          const double mm2 = rEm * rEm - rPm * rPm;
          ntu[34] = (sqrt(fabs(mm2)) * fabs(mm2) / mm2) / 1000.;
          ntu[35] = rpP / 1000.;
          ntu[36] = (Q2 - Mh2_det + 2 * (nu * rpE - rpP * q * m::cos(r_thpq))) / 1.e6;
          ntu[37] = -S.ld(F_RASTERY, slot); ntu[38] = radphot / 1000.; ntu[39] = sigcc; ntu[40] = 0.0; ntu[41] = wfinal;
          ntu[42] = (semi && !cfg.doing_decay) ? survivalprob : S.ld(F_DECDIST, slot);
          ntu[43] = sqrt(S.ld(F_MH2FINAL, slot));
          const double cth = m::cos(r_thpq);
          ntu[44] = rpE / nu; ntu[45] = S.ld(F_ZHAD, slot);
          ntu[46] = (rpP * rpP * (1.0 - cth * cth)) / 1.e06; ntu[47] = S.ld(F_PT2, slot) / 1.e06;
          ntu[48] = Q2 / 2. / SIMC_MP / nu; ntu[49] = v_Q2 / 2. / SIMC_MP / S.ld(F_VNU, slot);
          ntu[50] = m::acos(S.ld(F_UQZ, slot)); ntu[51] = S.ld(F_SIGCM, slot); ntu[52] = S.ld(F_DAVEJAC, slot); ntu[53] = johnjac;
          const double dummy = S.ld(F_PFERX, slot) * S.ld(F_UQX, slot) + S.ld(F_PFERY, slot) * S.ld(F_UQY, slot) +
                               S.ld(F_PFERZ, slot) * S.ld(F_UQZ, slot);
          ntu[54] = S.ld(F_PFER, slot) / 1000. * fabs(dummy) / dummy;     // NaN for hydrogen (0/0), as in the reference
          ntu[55] = semi ? S.ld(F_XFERMI, slot) : 0.0; ntu[56] = S.ld(F_MPHIPQ, slot);
          if (cfg.using_tgt_field) {       // results_write.f:212-225
            const PolTargAngles ra = poltarg_angles(cfg.targ_pol, cfg.targ_Bangle, uqx, uqy, uqz, upx, upy, upz, r_phipq);
            const PolTargAngles va = poltarg_angles(cfg.targ_pol, cfg.targ_Bangle, S.ld(F_UQX, slot), S.ld(F_UQY, slot), S.ld(F_UQZ, slot),
                                                    S.ld(F_UPX, slot), S.ld(F_UPY, slot), S.ld(F_UPZ, slot), S.ld(F_MPHIPQ, slot));
            ntu[57] = ra.theta_tarq; ntu[58] = ra.phi_targ; ntu[59] = ra.beta; ntu[60] = ra.phi_s; ntu[61] = ra.phi_c;
            ntu[62] = va.beta; ntu[63] = va.phi_s; ntu[64] = va.phi_c;
            if (rho) { ntu[65] = S.ld(F_RHOMASS, slot); ntu[66] = S.ld(F_RHOTHETA, slot); }       // (the 67th tag, mmnuc, is never filled)
          } else if (rho) {  // results_write.f:226-230
            const double e_A = nu + cfg.targ.M - rpE;
            const double mmA2 = e_A * e_A - rPm * rPm;
            ntu[57] = S.ld(F_RHOMASS, slot); ntu[58] = S.ld(F_RHOTHETA, slot);
            ntu[59] = (sqrt(fabs(mmA2)) * fabs(mmA2) / mmA2) / 1000.;
          }
        } else if (meson) {
          const double mm2 = rEm * rEm - rPm * rPm;
          const double e_A = nu + cfg.targ.M - rpE;
          const double mmA2 = e_A * e_A - rPm * rPm;
          ntu[34] = (sqrt(fabs(mm2)) * fabs(mm2) / mm2) / 1000.;
          ntu[35] = (sqrt(fabs(mmA2)) * fabs(mmA2) / mmA2) / 1000.;
          ntu[36] = rpP / 1000.;
          ntu[37] = (Q2 - Mh2_det + 2 * (nu * rpE - rpP * q * m::cos(r_thpq))) / 1.e6;
          ntu[38] = PmPar / 1000.; ntu[39] = PmPer / 1000.; ntu[40] = PmOop / 1000.;
          ntu[41] = -S.ld(F_RASTERY, slot); ntu[42] = radphot / 1000.;
          double pdot = mv_pfer[1] * S.ld(F_UQX, slot) + mv_pfer[2] * S.ld(F_UQY, slot) + mv_pfer[3] * S.ld(F_UQZ, slot);
          if (pdot == 0) pdot = 1.e-20;
          ntu[43] = mv_pfer[0] / 1000. * fabs(pdot) / pdot;
          ntu[44] = sigcc; ntu[45] = S.ld(F_SIGCM, slot); ntu[46] = wfinal;
          ntu[47] = (cfg.doing_kaon && !cfg.doing_decay) ? survivalprob : S.ld(F_DECDIST, slot);
          ntu[48] = sqrt(S.ld(F_MH2FINAL, slot));
          ntu[49] = mv_pfer[0] / 1000. * pdot;
          ntu[50] = v_Q2 / 1.e6; ntu[51] = S.ld(F_MW, slot) / 1.e3; ntu[52] = S.ld(F_MT, slot) / 1.e6; ntu[53] = S.ld(F_MPHIPQ, slot);
          if (cfg.using_tgt_field) {       // results_write.f:154-166
            const PolTargAngles ra = poltarg_angles(cfg.targ_pol, cfg.targ_Bangle, uqx, uqy, uqz, upx, upy, upz, r_phipq);
            const PolTargAngles va = poltarg_angles(cfg.targ_pol, cfg.targ_Bangle, S.ld(F_UQX, slot), S.ld(F_UQY, slot), S.ld(F_UQZ, slot),
                                                    S.ld(F_UPX, slot), S.ld(F_UPY, slot), S.ld(F_UPZ, slot), S.ld(F_MPHIPQ, slot));
            ntu[54] = ra.theta_tarq; ntu[55] = ra.phi_targ; ntu[56] = ra.beta; ntu[57] = ra.phi_s; ntu[58] = ra.phi_c;
            ntu[59] = va.beta; ntu[60] = va.phi_s; ntu[61] = va.phi_c;
            if (cfg.doing_kaon) { ntu[62] = sigcm1; ntu[63] = S.ld(F_SIGCM, slot); }
          } else if (cfg.doing_kaon) { ntu[54] = sigcm1; ntu[55] = S.ld(F_SIGCM, slot); }
        } else {
          const double sh = m::sin(reth / 2.);
          const double poftheta = SIMC_MP * cfg.Ebeam / (2 * cfg.Ebeam * (sh * sh) + SIMC_MP);
          const double nrm = sqrt(uqy * uqy + uqz * uqz);
          ntu[34] = (reP - poftheta) / 1000.;
          ntu[35] = (-Pmx) / 1000.;
          ntu[36] = ((Pmz * uqy - Pmy * uqz) / nrm) / 1000.;
          ntu[37] = (-(Pmy * uqy + Pmz * uqz) / nrm) / 1000.;
          ntu[38] = PmPar / 1000.; ntu[39] = PmPer / 1000.; ntu[40] = PmOop / 1000.;
          ntu[41] = -S.ld(F_RASTERY, slot); ntu[42] = radphot / 1000.; ntu[43] = sigcc; ntu[44] = wfinal;
          ntu[45] = reth; ntu[46] = rpth;
        }
#pragma unroll
        for (int k = 0; k < SIMC_NTUPLE_MAXCOL; ++k)
          if (!(cfg.doing_pizero && k >= 53)) S.st(F_NTU0 + k, slot, ntu[k + 1]);      // columns 54-65 of a pi0 row: k_calo wrote them
      }
      if (success) {
        no_rad_p = S.ld(F_RADP, slot) == 0.0;
        rec_vals[0] = red; rec_vals[1] = rey; rec_vals[2] = rex; rec_vals[3] = rpd; rec_vals[4] = rpy; rec_vals[5] = rpx;
        const double v_ed = S.ld(F_VEDELTA, slot), v_ey = S.ld(F_VEYP, slot), v_ex = S.ld(F_VEXP, slot);
        const double v_pd = S.ld(F_VPDELTA, slot), v_py = S.ld(F_VPYP, slot), v_px = S.ld(F_VPXP, slot);
        const double v_Em = S.ld(F_VEM, slot), v_Pm = S.ld(F_VPM, slot), v_Trec = S.ld(F_VTREC, slot);
        gen_vals[0] = v_ed; gen_vals[1] = v_ey; gen_vals[2] = -v_ex; gen_vals[3] = v_pd; gen_vals[4] = v_py;
        gen_vals[5] = -v_px; gen_vals[6] = v_Em;
        const double spe_d = S.ld(F_SPE_D, slot), spe_y = S.ld(F_SPE_Y, slot), spe_x = S.ld(F_SPE_X, slot), spe_z = S.ld(F_SPE_Z, slot);
        const double spp_d = S.ld(F_SPP_D, slot), spp_y = S.ld(F_SPP_Y, slot), spp_x = S.ld(F_SPP_X, slot), spp_z = S.ld(F_SPP_Z, slot);
        err[0] = red - spe_d; err[1] = rex - v_ex; err[2] = rey - v_ey; err[3] = rez - spe_z;
        err[4] = rpd - spp_d; err[5] = rpx - v_px; err[6] = rpy - v_py; err[7] = rpz - spp_z;
        // limits_update, event.f:1-90
        const double Ein_shift = S.ld(F_EINSHIFT, slot), Ee_shift = S.ld(F_EESHIFT, slot);
        const double v_pE = S.ld(F_VPE, slot);
        const double o_eE = S.ld(F_OEE, slot), o_pE = S.ld(F_OPE, slot);
        const double o_py = rho ? S.ld(F_OPYP, slot) : v_py, o_px = rho ? S.ld(F_OPXP, slot) : v_px;      // the decay pion's
        const double o_Em = v_Em, o_Pm = v_Pm, o_Trec = v_Trec;     // orig = vertex for these (radc.f:476)
        const double eg0 = S.ld(F_EG0, slot), eg1 = S.ld(F_EG1, slot), eg2 = S.ld(F_EG2, slot);
        const double sumEgen = ((meson && !semi) || deut) ? v_eE - Ein_shift : v_eE + v_pE - Ein_shift;     // event.f:28-32
        const double c_[30] = {v_ed, v_ey, v_ex, v_pd, v_py, v_px, S.ld(F_MTREC, slot), sumEgen,
                               o_eE - Ee_shift, v_ex, v_ey, o_pE, o_py, o_px, o_Em - Ein_shift + Ee_shift, o_Pm, o_Trec,
                               spe_d, spe_y, spe_x, spp_d, spp_y, spp_x, v_Trec, v_Em, v_Pm, eg0, eg1, eg2, eg0 + eg1 + eg2};
#pragma unroll
        for (int k = 0; k < 30; ++k) cv[k] = c_[k];
        sv[0] = red - spe_d; sv[1] = rey - spe_y; sv[2] = rex - spe_x; sv[3] = rpd - spp_d; sv[4] = rpy - spp_y;
        sv[5] = rpx - spp_x; sv[6] = rEm - (o_Em - Ein_shift + Ee_shift); sv[7] = fabs(rPm) - fabs(o_Pm);
      }
    }
    // ---- accumulation, simc.f:248-336: whole warps take part, invalid lanes contribute the identity
    if (!__any_sync(0xffffffffu, success)) continue;
    warp_sum128(B.sums[1], success, sigcc, A.qexp_w);
    warp_count_if(&B.counters[0], success);
    warp_count_if(&B.counters[1], success);
    warp_count_if(&B.counters[3], success && no_rad_p);
    warp_count_if(&B.counters[4], success && low_w);
    // weights the fixed-point sums cannot hold (NaN, inf, |w| >= 1e38 quanta): to_fixed drops them, count them
    warp_count_if(&B.counters[5], success && (!(fabs(ldexp(weight, -A.qexp_w)) < 1.0e38) || !(fabs(ldexp(sigcc, -A.qexp_w)) < 1.0e38)));
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const int b = success ? hist_bin(cfg.hist_axis[0][k], rec_vals[k]) : -1;
      if (b >= 0) {          // bins differ between lanes: CTA-level 64-bit atomics with carry
        unsigned long long lo, hi;
        to_fixed(weight, A.qexp_w, lo, hi);
        const unsigned long long old = atomicAdd(&B.hist_w[k][b][0], lo);
        const unsigned long long carry = (old + lo < old) ? 1ULL : 0ULL;
        if (hi + carry) atomicAdd(&B.hist_w[k][b][1], hi + carry);
      }
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 7; ++k) warp_hist_add(B.hist_n[k], success ? hist_bin(cfg.hist_axis[1][k], gen_vals[k]) : -1);
    warp_hist_add(B.hist_n[7], success ? hist_bin(cfg.hist_axis[0][SIMC_H_EM], rEm) : -1);
    warp_hist_add(B.hist_n[8], success ? hist_bin(cfg.hist_axis[0][SIMC_H_PM], rPm) : -1);
    const bool pc = success && pass_cuts;
    warp_count_if(&B.counters[2], pc);
    if (__any_sync(0xffffffffu, pc)) {
      warp_sum128(B.sums[0], pc, weight, A.qexp_w);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        warp_sum128(B.sums[2 + k], pc, err[k], -80);
        warp_sum128(B.sums[10 + k], pc, err[k] * err[k], -80);
      }
    }
#pragma unroll
    for (int k = 0; k < 30; ++k) warp_minmax(&B.mins[k], &B.maxs[k], success, cv[k]);
#pragma unroll
    for (int k = 0; k < 8; ++k) warp_minmax(&B.mins[32 + k], &B.maxs[32 + k], success, sv[k]);
  }
  __syncthreads();
  // ---- flush the CTA's accumulators
  for (int k = threadIdx.x; k < 18; k += kFinBlock) {
    unsigned long long* dst = k == 0 ? acc->wt : k == 1 ? acc->sigcc : k < 10 ? acc->sumerr[k - 2] : acc->sumerr2[k - 10];
    const unsigned long long lo = B.sums[k][0], hi = B.sums[k][1];
    if (lo | hi) {
      const unsigned long long old = atomicAdd(&dst[0], lo);
      const unsigned long long carry = (old + lo < old) ? 1ULL : 0ULL;
      if (hi + carry) atomicAdd(&dst[1], hi + carry);
    }
  }
  for (int k = threadIdx.x; k < 6 * SIMC_NHIST; k += kFinBlock) {
    const unsigned long long lo = (&B.hist_w[0][0][0])[2 * k], hi = (&B.hist_w[0][0][0])[2 * k + 1];
    if (lo | hi) {
      unsigned long long* dst = &acc->hist_w[0][0][0] + 2 * k;
      const unsigned long long old = atomicAdd(&dst[0], lo);
      const unsigned long long carry = (old + lo < old) ? 1ULL : 0ULL;
      if (hi + carry) atomicAdd(&dst[1], hi + carry);
    }
  }
  for (int k = threadIdx.x; k < 9 * SIMC_NHIST; k += kFinBlock) {
    const unsigned v = (&B.hist_n[0][0])[k];
    if (!v) continue;
    const int h = k / SIMC_NHIST, b = k % SIMC_NHIST;
    unsigned long long* dst = h < 7 ? &acc->hist_n[1][h][b] : h == 7 ? &acc->hist_n[0][SIMC_H_EM][b] : &acc->hist_n[0][SIMC_H_PM][b];
    atomicAdd(dst, (unsigned long long)v);
  }
  if (threadIdx.x < 6 && B.counters[threadIdx.x]) atomicAdd(&acc->counters[1 + threadIdx.x], (unsigned long long)B.counters[threadIdx.x]);
  for (int k = threadIdx.x; k < 40; k += kFinBlock) {
    if (B.mins[k] == 0x7fffffffffffffffLL) continue;
    if (k < 32) { atomicMin(&acc->contrib_lo[k], B.mins[k]); atomicMax(&acc->contrib_hi[k], B.maxs[k]); }
    else { atomicMin(&acc->slop_lo[k - 32], B.mins[k]); atomicMax(&acc->slop_hi[k - 32], B.maxs[k]); }
  }
}

// ---- parity entry point: the end of the loop body on dumped vectors (simc_b200_weight_batch) ----------
// Row i of the input becomes slot i of the state buffer, the list of survivors is 0..n-1, k_finish runs in record
// mode, and k_wb_store collects what it left.  Column order: include/simc_b200.h.
__global__ void k_wb_load(LoopArgs A, long long n, const double* __restrict__ in) {
  const StateBuf& S = A.st;
  const int fields[SIMC_WEIGHT_NIN] = {
      F_RE_E, F_RE_TH, F_RE_PH, F_RP_P, F_RP_E, F_RP_TH, F_RP_PH, F_VEIN, F_VEE, F_VETHETA, F_VQ2, F_VNU, F_VQ, F_VPE, F_VPP,
      F_UQX, F_UQY, F_UQZ, F_UPX, F_UPY, F_UPZ, F_VEM, F_VPM, F_MPHIPQ, F_MT, F_MEPS, F_JAC, F_GENW, F_ZHAD, F_PT2,
      F_PFER, F_PFERX, F_PFERY, F_PFERZ, F_EFER, F_FPP_PATH, F_FPP_DX, F_FPP_DY, F_RCE_D, F_RCE_Y, F_RCE_X, F_RCP_D, F_RCP_Y, F_RCP_X};
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) A.counts[1 + 2 * kArmLists] = (unsigned)n;
  if (i >= n) return;
  for (int k = 0; k < SIMC_WEIGHT_NIN; ++k) S.st(fields[k], i, in[(long long)k * n + i]);
  S.st(F_TRY, i, (double)i);
  A.lists[(long long)(2 * kArmLists) * A.st.cap + i] = (unsigned)i;
}
__global__ void k_wb_store(LoopArgs A, long long n, double* __restrict__ out) {
  const StateBuf& S = A.st;
  const int fields[SIMC_WEIGHT_NOUT] = {F_STAGE, F_PASSCUTS, F_WEIGHT, F_SIGCC, F_SIGCC_RECON, F_REM, F_RPM, F_RW, F_THCM, F_PHICM,
                                        F_SIGCM, F_DAVEJAC, F_SURV, F_MM, F_WCM};
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int k = 0; k < SIMC_WEIGHT_NOUT; ++k) out[(long long)k * n + i] = S.ld(fields[k], i);
  out[0 * n + i] = S.ld(F_STAGE, i) == 4.0 ? 1.0 : 0.0;
}

// ---- parity entry point: per-try records (include/simc_b200.h: simc_b200_event_batch) ----------
__global__ void k_records(LoopArgs A, double* __restrict__ rec, int* __restrict__ status, long long n) {
  const StateBuf& S = A.st;
  const unsigned n_slots = A.counts[0];
  for (long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x; slot < n_slots;
       slot += (long long)gridDim.x * blockDim.x) {
    const long long i = (long long)S.ld(F_TRY, slot);
    const int stage = (int)S.ld(F_STAGE, slot);
    // stage codes of the oracle: 0 generate failed, 1 P arm failed, 2 E arm failed, 3 failed later, 4 success.
    // F_STAGE holds how far the try got: 0 gen failed, 1 gen ok (P arm failed), 2 P ok (E arm failed), 3, 4.
    const int fields[SIMC_EVENT_NREC] = {
        F_STAGE, F_PASSCUTS, F_DRAW, F_STOP_P, F_STOP_E, F_WEIGHT, F_SIGCC, F_GENW, F_JAC, F_SIGCC_RECON,
        F_VEIN, F_VEE, F_VEDELTA, F_VEYP, F_VEXP, F_VPE, F_VPDELTA, F_VPYP, F_VPXP, F_VQ2,
        F_OEE, F_OPE, F_EG0, F_EG1, F_EG2, F_NTAIL, F_TX, F_TY, F_TZ, F_ELOSS0, F_ELOSS1, F_ELOSS2,
        F_SPE_D, F_SPE_Y, F_SPE_X, F_SPP_D, F_SPP_Y, F_SPP_X, F_RCE_D, F_RCE_Y, F_RCE_X, F_RCP_D, F_RCP_Y, F_RCP_X,
        F_REM, F_RPM, F_RW, F_HARDCOR, F_THCM, F_PHICM, F_SIGCM, F_DAVEJAC, F_SURV, F_MM, F_WCM, F_MT,
        F_OPYP, F_OPXP, F_RHOMASS, F_RHOTHETA};
    if (A.record_mode == 2) {        // ntuple columns instead of the parity record
      for (int k = 0; k < SIMC_NTUPLE_MAXCOL; ++k) rec[(long long)k * n + i] = S.ld(F_NTU0 + k, slot);
    } else {
      const int n_fields = A.cfg->doing_rho ? SIMC_EVENT_NREC : SIMC_EVENT_NREC - 4;     // the last four: rho production only
      for (int k = 0; k < SIMC_EVENT_NREC; ++k) rec[(long long)k * n + i] = k < n_fields ? S.ld(fields[k], slot) : 0.0;
      rec[0 * n + i] = (double)stage;
    }
    status[i] = stage;
  }
}

}  // namespace SIMC_VARIANT_NS
}  // namespace simc
