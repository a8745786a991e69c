// CUDA kernels of libsimc_b200 (sm_100a).  Compiled twice: -DSIMC_STRICT=1 -fmad=false
// (namespace simc::strict) and -DSIMC_STRICT=0 -fmad=false (namespace simc::fast: explicit fma()
// and re-associated monomials in the COSY polynomials only).
#if SIMC_STRICT
#define SIMC_VARIANT_NS strict
#else
#define SIMC_VARIANT_NS fast
#endif
#include "transport.cuh"
#include "kernels.h"
#include "loop.cuh"
#include "jit.h"
#include <string.h>
#include <stddef.h>

#include <mutex>
#include <algorithm>

namespace simc {
namespace SIMC_VARIANT_NS {

// Writes the Philox round keys of `seed` into this variant's constant memory on the current device, if they are not
// there already.  The constant is device-wide, so a change of seed waits for whatever is in flight on the device
// first (another handle may still be running with the old one); launches with the seed already loaded cost a compare.
// The caller holds rng_key_mutex() until its kernels are enqueued, so that wait sees them.
static std::mutex& rng_key_mutex() {
  static std::mutex mu;
  return mu;
}
static cudaError_t ensure_rng_key_locked(unsigned long long seed) {
  static unsigned long long cur[64];
  static bool valid[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (valid[dev] && cur[dev] == seed) return cudaSuccess;
  uint32_t rk[20];
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) { rk[2 * r] = k0; rk[2 * r + 1] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return e;
  e = cudaMemcpyToSymbol(c_philox_rk, rk, sizeof(rk));
  if (e != cudaSuccess) return e;
  cur[dev] = seed; valid[dev] = true;
  return cudaSuccess;
}

// Batch form of mc_hms / mc_shms / ... (hms/mc_hms.f:1-4): one thread per row.
// tk == nullptr: the whole program from the input rows.  tk != nullptr (compiled path): the rows of `list` resume
// at op_begin from the track the generated stretches left in tk.
template <bool WITH_COLL>
__global__ void __launch_bounds__(kBlock)
k_transport_batch(const __grid_constant__ ArmDev arm_c, long long n, const double* __restrict__ in,
                  unsigned long long seed, ArmFlags f, double ctau, double* __restrict__ out,
                  int* __restrict__ flags, int op_begin, const double* __restrict__ tk, const unsigned* __restrict__ list,
                  const unsigned* __restrict__ count) {
  extern __shared__ double pw_s[];
  const ArmDev* arm = &arm_c;
  const long long j = (long long)blockIdx.x * kBlock + threadIdx.x;
  const long long n_rows = tk ? (long long)*count : n;
  bool alive = j < n_rows;
  const long long i = tk ? (alive ? (long long)list[j] : 0) : j;
  const long long ii = alive ? i : 0;
  TrackDev t;
  t.dpps = in[0 * n + ii];
  t.xs = in[1 * n + ii];
  t.ys = in[2 * n + ii];
  // in[3] = z is carried by the reference (zs) but never read by any single-arm routine
  t.dxdzs = in[4 * n + ii];
  t.dydzs = in[5 * n + ii];
  t.m2 = in[6 * n + ii];
  const double p_spec = in[7 * n + ii];
  const double fry = in[8 * n + ii];
  t.p = p_spec * (1. + t.dpps / 100.);     // mc_hms.f:181
  t.p_spec = p_spec;
  t.pathlen = 0.0;
  t.decdist = 0.0;
  t.mh2_final = t.m2;
  t.ctau = ctau;
  const double dpp_in = t.dpps, y_in = t.ys, dxdz_in = t.dxdzs, dydz_in = t.dydzs;
  if (tk) {        // rows of tk: xs, ys, dxdzs, dydzs, dpps, p, m2, pathlen (loop.cuh F_TK_*)
    t.xs = tk[0 * n + ii]; t.ys = tk[1 * n + ii]; t.dxdzs = tk[2 * n + ii]; t.dydzs = tk[3 * n + ii];
    t.dpps = tk[4 * n + ii]; t.pathlen = tk[7 * n + ii];
  }
  DevRng rng;
  rng.init((unsigned long long)i, 0u, 0u);
  ArmResult res;
  arm_result_clear(res);
  HutState hs;
#pragma unroll
  for (int k = 0; k < 12; ++k) { hs.xdc[k] = 0.f; hs.ydc[k] = 0.f; }
  hs.scincount = 0;
  t.dflag = false;
  musc_refresh(t);
  const unsigned ring = (unsigned)__cvta_generic_to_shared(pw_s) + (unsigned)kPowBytes + (threadIdx.x >> 5) * kRingBytesPerWarp;
  run_arm<WITH_COLL>(arm, t, rng, f, fry, pw_s + threadIdx.x, ring, res, hs, alive, op_begin, arm->tab.n_ops);
  if (j >= n_rows) return;
  out[0 * n + i] = res.ok ? res.dpp_rec : dpp_in;
  out[1 * n + i] = res.ok ? res.dph_rec : dxdz_in;
  out[2 * n + i] = res.ok ? res.dth_rec : dydz_in;
  out[3 * n + i] = res.ok ? res.y_rec : y_in;
  out[4 * n + i] = res.x_fp;
  out[5 * n + i] = res.dx_fp;
  out[6 * n + i] = res.y_fp;
  out[7 * n + i] = res.dy_fp;
  out[8 * n + i] = t.pathlen;
  out[9 * n + i] = t.m2;
  out[10 * n + i] = res.resmult;
  out[11 * n + i] = (double)rng.draw;
  flags[i] = res.ok ? 0 : res.stop_code;
}

// compiled path, first kernel: the rows become tracks (mc_hms.f:181), every row is on list 0
__global__ void k_tb_load(long long n, const double* __restrict__ in, double* __restrict__ tk, unsigned* __restrict__ list0,
                          unsigned* __restrict__ counts, int n_counts, unsigned long long* __restrict__ sink, int n_sink) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_counts) counts[i] = i == 0 ? (unsigned)n : 0u;
  if (i < n_sink) sink[i] = 0ULL;
  if (i >= n) return;
  const double dpp = in[0 * n + i], p_spec = in[7 * n + i];
  tk[0 * n + i] = in[1 * n + i]; tk[1 * n + i] = in[2 * n + i]; tk[2 * n + i] = in[4 * n + i]; tk[3 * n + i] = in[5 * n + i];
  tk[4 * n + i] = dpp; tk[5 * n + i] = p_spec * (1. + dpp / 100.); tk[6 * n + i] = in[6 * n + i]; tk[7 * n + i] = 0.0;
  tk[11 * n + i] = -1.0;            // stop code of the compiled stretches (none yet)
  list0[i] = (unsigned)i;
}
// compiled path, last kernel: the rows a compiled stretch stopped (mc_hms returns its inputs untouched for them)
__global__ void k_tb_stopped(long long n, const double* __restrict__ in, const double* __restrict__ tk, double* __restrict__ out,
                             int* __restrict__ flags) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double code = tk[11 * n + i];
  if (code < 0.0) return;
  out[0 * n + i] = in[0 * n + i]; out[1 * n + i] = in[4 * n + i]; out[2 * n + i] = in[5 * n + i]; out[3 * n + i] = in[2 * n + i];
  out[4 * n + i] = 0.0; out[5 * n + i] = 0.0; out[6 * n + i] = 0.0; out[7 * n + i] = 0.0;
  out[8 * n + i] = tk[7 * n + i]; out[9 * n + i] = in[6 * n + i]; out[10 * n + i] = 0.0; out[11 * n + i] = 0.0;
  flags[i] = (int)code;
}

// > 48 KB of dynamic shared memory needs an opt-in per function AND per device (the attribute lives in the context):
// done once for every device a handle runs on.  The caller holds rng_key_mutex().
static cudaError_t ensure_smem_opt_in_locked() {
  static bool done[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (done[dev]) return cudaSuccess;
  const int bytes = (int)kArmSmemBytes;
  e = cudaFuncSetAttribute(k_transport_batch<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_transport_batch<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_arm<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_arm<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_arm<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_arm<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_arm<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_arm<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_arm<1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_arm<0, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_arm<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_arm<0, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_arm<1, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_arm<0, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_arm<1, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_arm<0, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_arm<1, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_arm<0, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return e;
  done[dev] = true;
  return cudaSuccess;
}

cudaError_t launch_transport_batch(const TransportBatchArgs& a, cudaStream_t s) {
  if (a.n <= 0) return cudaSuccess;
  const long long blocks = (a.n + kBlock - 1) / kBlock;
  ArmFlags f;
  f.ms_flag = a.ms_flag != 0; f.wcs_flag = a.wcs_flag != 0; f.decay_flag = a.decay_flag != 0;
  f.using_coll = a.using_coll != 0;
  std::lock_guard<std::mutex> key_lock(rng_key_mutex());
  {
    cudaError_t e = ensure_smem_opt_in_locked();
    if (e == cudaSuccess) e = ensure_rng_key_locked(a.seed);
    if (e != cudaSuccess) return e;
  }
  if (a.n_stretch > 0 && !f.using_coll && !f.decay_flag) {
    // compiled path: rows -> tracks, the generated stretches with compaction in between, the interpreter from the
    // hut on for the survivors, then the rows that stopped on the way
    const int n_sink = SIMC_NSTOP + 48;
    const long long nb = (std::max<long long>(a.n, n_sink) + 255) / 256;
    k_tb_load<<<(unsigned)nb, 256, 0, s>>>(a.n, a.in, a.tk, a.lists, a.counts, a.n_stretch + 1, a.sink, n_sink);
    for (int k = 0; k < a.n_stretch; ++k) {
      double* tk = a.tk;
      long long fs = a.n, ss = 1;          // this entry point's own scratch rows: [field][row]
      const unsigned* in_list = a.lists + (long long)k * a.n;
      const unsigned* in_count = a.counts + k;
      unsigned* out_list = a.lists + (long long)(k + 1) * a.n;
      unsigned* out_count = a.counts + k + 1;
      unsigned long long* stop_acc = a.sink;
      unsigned long long* calls_acc = a.sink + SIMC_NSTOP;
      double* stop_field = a.tk + 11 * a.n;
      void* args[] = {&tk, &fs, &ss, &in_list, &in_count, &out_list, &out_count, &stop_acc, &calls_acc, &stop_field};
      const long long g = std::min<long long>((a.n + a.stretch_block - 1) / a.stretch_block, a.stretch_grid);
      if (jit_launch(a.stretch_fn[k], (unsigned)g, (unsigned)a.stretch_block, (void*)s, args) != 0) return cudaErrorLaunchFailure;
    }
    k_transport_batch<false><<<(unsigned)blocks, kBlock, kArmSmemBytes, s>>>(*(const ArmDev*)a.arm, a.n, a.in, a.seed, f, a.ctau, a.out,
                                                                 a.flags, a.hut_begin, a.tk, a.lists + (long long)a.n_stretch * a.n,
                                                                 a.counts + a.n_stretch);
    k_tb_stopped<<<(unsigned)((a.n + 255) / 256), 256, 0, s>>>(a.n, a.in, a.tk, a.out, a.flags);
    return cudaGetLastError();
  }
  if (f.using_coll)
    k_transport_batch<true><<<(unsigned)blocks, kBlock, kArmSmemBytes, s>>>(*(const ArmDev*)a.arm, a.n, a.in, a.seed, f, a.ctau, a.out,
                                                                a.flags, 0, nullptr, nullptr, nullptr);
  else
    k_transport_batch<false><<<(unsigned)blocks, kBlock, kArmSmemBytes, s>>>(*(const ArmDev*)a.arm, a.n, a.in, a.seed, f, a.ctau, a.out,
                                                                 a.flags, 0, nullptr, nullptr, nullptr);
  return cudaGetLastError();
}

// Stage-level parity entry point: radc_init_ev + peaked_rad_weight on dumped vertex inputs.
__global__ void k_radc_batch(const simc_run_config* __restrict__ cfg, long long n, const double* __restrict__ in,
                             double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  VertexKin v;
  v.Ein = in[0 * n + i]; v.eE = in[1 * n + i]; v.eP = v.eE; v.etheta = in[2 * n + i];
  v.uex = in[3 * n + i]; v.uey = in[4 * n + i]; v.uez = in[5 * n + i];
  v.pE = in[6 * n + i]; v.pP = in[7 * n + i];
  v.upx = in[8 * n + i]; v.upy = in[9 * n + i]; v.upz = in[10 * n + i];
  RadEvDev R;
  radc_init_ev(*cfg, v, in[11 * n + i], in[12 * n + i], R);
  const double w = peaked_rad_weight(*cfg, R, v, in[13 * n + i], in[14 * n + i], in[15 * n + i], 1.0);
  out[0 * n + i] = R.bt[0]; out[1 * n + i] = R.bt[1];
  out[2 * n + i] = R.lambda[0]; out[3 * n + i] = R.lambda[1]; out[4 * n + i] = R.lambda[2];
  out[5 * n + i] = R.g[4]; out[6 * n + i] = R.hardcorfac; out[7 * n + i] = R.c4; out[8 * n + i] = R.c_ext0;
  out[9 * n + i] = w;
  out[10 * n + i] = sigep(v.Ein, v.eE, v.etheta, 2 * v.Ein * v.eE * (1. - v.uez));
  // the (Egamma1, Egamma2, Egamma3) basis and the pieces of the other option branches (include/simc_b200.h)
  const BasisConst B = basis_constants(R, v.Ein, v.eE, v.pE);
  out[11 * n + i] = B.c[1]; out[12 * n + i] = B.c[2]; out[13 * n + i] = B.c[3]; out[14 * n + i] = B.c[0];
  out[15 * n + i] = B.c_int0; out[16 * n + i] = B.g_int;
  out[17 * n + i] = extrad_phi(*cfg, R, 1, v.Ein, v.eE, in[13 * n + i]);
  out[18 * n + i] = extrad_phi(*cfg, R, 2, v.Ein, v.eE, in[13 * n + i]);
  double ds, dh;
  schwinger(*cfg, 450., v.Ein, v.eE, v.etheta, 2 * v.Ein * v.eE * (1. - v.uez), v.Ein - v.eE, ds, dh);
  out[19 * n + i] = ds; out[20 * n + i] = dh;
  double db, dbp;
  extrad_friedrich(cfg->etatzai, v.Ein, in[13 * n + i], R.bt[0] / cfg->etatzai, db, dbp);
  out[21 * n + i] = db; out[22 * n + i] = dbp;
  double bs, bh, dbs;
  brem_onshell<kBremAll>(v.Ein, v.eE, 450., R.rad_proton_this_ev, bs, bh, dbs);
  out[23 * n + i] = bs; out[24 * n + i] = bh; out[25 * n + i] = dbs;
}
// Stage-level parity entry point: track_from_tgt (trg_track.f:591-672) on dumped vectors (simc_b200_field_batch).
// in [7][n]: x, y, z, dx, dy (TRANSPORT coordinates of the arm), mom (MeV/c, signed by the charge), mass;
// out [6][n]: x, y, z, dx, dy of the image track at z = 100 cm, ok.
__global__ void k_field_batch(FieldDev F, int k, long long n, const double* __restrict__ in, double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x = in[0 * n + i], y = in[1 * n + i], z = in[2 * n + i], dx = in[3 * n + i], dy = in[4 * n + i];
  const bool ok = track_from_tgt(F, k, x, y, z, dx, dy, in[5 * n + i], in[6 * n + i]);
  out[0 * n + i] = x; out[1 * n + i] = y; out[2 * n + i] = z; out[3 * n + i] = dx; out[4 * n + i] = dy;
  out[5 * n + i] = ok ? 1.0 : 0.0;
}
cudaError_t launch_field_batch(const double* map, double theta_deg, int spect, long long n, const double* in, double* out,
                               cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  FieldDev F;
  F.map = map;
  const double pi180 = 3.141592653 / 180.;            // trgInit's constant, trg_track.f:275
  const int k = spect == -1 ? 0 : 1;
  for (int j = 0; j < 2; ++j) { F.stht[j] = std::sin(theta_deg * pi180); F.ctht[j] = std::cos(theta_deg * pi180); }
  k_field_batch<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(F, k, n, in, out);
  return cudaGetLastError();
}

cudaError_t launch_radc_batch(const void* cfg, long long n, const double* in, double* out, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  k_radc_batch<<<(unsigned)((n + 127) / 128), 128, 0, s>>>((const simc_run_config*)cfg, n, in, out);
  return cudaGetLastError();
}

// Stage-level parity entry point: peepiX on dumped vertex vectors (include/simc_b200.h: simc_b200_semi_batch)
static Cteq5Dev make_pdf(const LoopLaunch& a) {
  Cteq5Dev T;
  T.xv = a.pdf_buf; T.ql = a.pdf_buf ? a.pdf_buf + (a.pdf_nx + 1) : nullptr;
  T.upd = a.pdf_buf ? a.pdf_buf + (a.pdf_nx + 1) + (a.pdf_nt + 1) : nullptr;
  T.nx = a.pdf_nx; T.nt = a.pdf_nt; T.nfmx = a.pdf_nfmx; T.al = a.pdf_al;
  return T;
}
__global__ void k_semi_batch(const simc_run_config* __restrict__ cfg, Cteq5Dev T, FdssDev F, long long n, const double* __restrict__ in,
                             double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  SemiVertex v;
  v.Ein = in[0 * n + i]; v.eE = in[1 * n + i]; v.nu = in[2 * n + i]; v.Q2 = in[3 * n + i]; v.q = in[4 * n + i];
  v.uqx = in[5 * n + i]; v.uqy = in[6 * n + i]; v.uqz = in[7 * n + i]; v.pt2 = in[8 * n + i]; v.zhad = in[9 * n + i];
  v.theta_pq = in[10 * n + i]; v.pfer = in[11 * n + i]; v.pferx = in[12 * n + i]; v.pfery = in[13 * n + i];
  v.pferz = in[14 * n + i]; v.efer = in[15 * n + i];
  double dbg[11];
#pragma unroll
  for (int k = 0; k < 11; ++k) dbg[k] = 0.0;
  const SemiWeight w = peepiX(*cfg, T, F, v, dbg);
  out[0 * n + i] = w.sigcc; out[1 * n + i] = w.sighad; out[2 * n + i] = w.davejac; out[3 * n + i] = w.xbj;
#pragma unroll
  for (int k = 0; k < 11; ++k) out[(4 + k) * n + i] = dbg[k];
  out[15 * n + i] = w.bad ? 1.0 : 0.0;
}
cudaError_t launch_semi_batch(const void* cfg, const LoopLaunch& tables, long long n, const double* in, double* out,
                              cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  k_semi_batch<<<(unsigned)((n + 127) / 128), 128, 0, s>>>((const simc_run_config*)cfg, make_pdf(tables), FdssDev{tables.fdss_buf}, n, in,
                                                            out);
  return cudaGetLastError();
}

size_t arm_dev_bytes() { return sizeof(ArmDev); }
size_t dev_accum_bytes() { return sizeof(DevAccum); }
int n_state_fields() { return SIMC_STATE_AOS ? kStateStride : (int)F_NFIELDS; }

}  // namespace SIMC_VARIANT_NS
}  // namespace simc
#if SIMC_STRICT
// byte offset of the min/max key block inside DevAccum (same in both variants)
size_t simc_dev_accum_minmax_offset() { return offsetof(simc::strict::DevAccum, contrib_lo); }

// Multi-GPU end of run: `gathered` holds one DevAccum per rank (what a single all-gather of the
// accumulator blocks delivers); every rank folds them into its own block.  One thread per 64-bit word: counters and
// count histograms add, 128-bit fixed-point sums add with the carry of their low word, range keys take min / max.
// Integer arithmetic only, so every rank ends with the same bits whatever the order of the ranks.
namespace {
__global__ void k_reduce_gathered(const unsigned long long* __restrict__ gathered, int n_ranks, unsigned long long* __restrict__ out) {
  typedef simc::strict::DevAccum A;
  constexpr unsigned kWords = sizeof(A) / 8;
  constexpr unsigned kPairLo = offsetof(A, wt) / 8, kPairHi = offsetof(A, hist_n) / 8;      // (lo, hi) pairs
  constexpr unsigned kMinMax = offsetof(A, contrib_lo) / 8, kAfterMinMax = offsetof(A, stop) / 8;
  const unsigned w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= kWords) return;
  if (w >= kPairLo && w < kPairHi) {
    if ((w - kPairLo) & 1u) return;                        // the thread of the low word does the pair
    unsigned long long lo = 0, hi = 0;
    for (int r = 0; r < n_ranks; ++r) {
      const unsigned long long a = gathered[(size_t)r * kWords + w], b = gathered[(size_t)r * kWords + w + 1];
      const unsigned long long s = lo + a;
      hi += b + (s < lo ? 1ull : 0ull);
      lo = s;
    }
    out[w] = lo; out[w + 1] = hi;
    return;
  }
  if (w >= kMinMax && w < kAfterMinMax) {
    const unsigned k = w - kMinMax;                        // contrib_lo[32] contrib_hi[32] slop_lo[8] slop_hi[8]
    const bool is_min = k < 32 || (k >= 64 && k < 72);
    long long v = (long long)gathered[w];
    for (int r = 1; r < n_ranks; ++r) {
      const long long x = (long long)gathered[(size_t)r * kWords + w];
      v = is_min ? (x < v ? x : v) : (x > v ? x : v);
    }
    out[w] = (unsigned long long)v;
    return;
  }
  unsigned long long s = 0;
  for (int r = 0; r < n_ranks; ++r) s += gathered[(size_t)r * kWords + w];
  out[w] = s;
}
}  // namespace
cudaError_t simc_launch_reduce_gathered(const void* gathered, int n_ranks, void* dev_accum, cudaStream_t s) {
  const unsigned words = sizeof(simc::strict::DevAccum) / 8;
  k_reduce_gathered<<<(words + 255) / 256, 256, 0, s>>>((const unsigned long long*)gathered, n_ranks, (unsigned long long*)dev_accum);
  return cudaGetLastError();
}
#endif
namespace simc {
namespace SIMC_VARIANT_NS {
static inline double key_to_double(long long k) {
  const long long i = k >= 0 ? k : (k ^ 0x7fffffffffffffffLL);
  double d;
  memcpy(&d, &i, sizeof(d));
  return d;
}

// Adds a host copy of the device accumulators into the public simc_accum.
void accum_to_host(const void* dev_copy, void* out_v, int qexp_w) {
  const DevAccum& d = *(const DevAccum*)dev_copy;
  simc_accum& o = *(simc_accum*)out_v;
  typedef __int128 i128;
  auto addf = [](simc_fixed128& f, const unsigned long long* w, int qexp) {
    i128 v = (((i128)f.hi << 64) | (i128)f.lo) + (((i128)(long long)w[1] << 64) | (i128)w[0]);
    f.lo = (uint64_t)v; f.hi = (int64_t)(v >> 64); f.qexp = qexp;
  };
  o.ntried += (int64_t)d.counters[0]; o.nsuccess += (int64_t)d.counters[1]; o.ncontribute += (int64_t)d.counters[2];
  o.npasscuts += (int64_t)d.counters[3]; o.ncontribute_no_rad_proton += (int64_t)d.counters[4];
  o.unsupported += (int64_t)d.counters[5];
  o.nonfinite += (int64_t)d.counters[6];
  addf(o.wtcontribute, d.wt, qexp_w);
  addf(o.sum_sigcc, d.sigcc, qexp_w);
  for (int i = 0; i < 8; ++i) { addf(o.sumerr[i], d.sumerr[i], -80); addf(o.sumerr2[i], d.sumerr2[i], -80); }
  for (int h = 0; h < 6; ++h) for (int b = 0; b < SIMC_NHIST; ++b) addf(o.hist_w[h][b], d.hist_w[h][b], qexp_w);
  for (int s = 0; s < 3; ++s) for (int h = 0; h < SIMC_H_PER_SET; ++h) for (int b = 0; b < SIMC_NHIST; ++b)
    o.hist_n[s][h][b] += (int64_t)d.hist_n[s][h][b];
  for (int i = 0; i < 32; ++i) {
    const double lo = key_to_double(d.contrib_lo[i]), hi = key_to_double(d.contrib_hi[i]);
    if (lo < o.contrib[i].lo) o.contrib[i].lo = lo;
    if (hi > o.contrib[i].hi) o.contrib[i].hi = hi;
  }
  for (int i = 0; i < 8; ++i) {
    const double lo = key_to_double(d.slop_lo[i]), hi = key_to_double(d.slop_hi[i]);
    if (lo < o.slop[i].lo) o.slop[i].lo = lo;
    if (hi > o.slop[i].hi) o.slop[i].hi = hi;
  }
  for (int w = 0; w < 2; ++w) for (int i = 0; i < SIMC_NSTOP; ++i) o.stop[w][i] += (int64_t)d.stop[w][i];
  for (int w = 0; w < 2; ++w) for (int i = 0; i < 48; ++i) o.transp_calls[w][i] += (int64_t)d.transp_calls[w][i];
}

// ---- FP64 pipe microbenchmark (roofline denominator) ---------------------------------------------
template <bool FMA>
__global__ void k_fp64_peak(double* out, int iters) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    if (FMA) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    } else {
      a0 = __dadd_rn(__dmul_rn(a0, m), c); a1 = __dadd_rn(__dmul_rn(a1, m), c); a2 = __dadd_rn(__dmul_rn(a2, m), c);
      a3 = __dadd_rn(__dmul_rn(a3, m), c); a4 = __dadd_rn(__dmul_rn(a4, m), c); a5 = __dadd_rn(__dmul_rn(a5, m), c);
      a6 = __dadd_rn(__dmul_rn(a6, m), c); a7 = __dadd_rn(__dmul_rn(a7, m), c);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
// check entry point of the device logarithms (fastlog.cuh): out[0..n) = log(x), out[n..2n) = log10(x)
__global__ void k_log_batch(long long n, const double* __restrict__ x, double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = m::log(x[i]);
  out[n + i] = m::log10(x[i]);
}
cudaError_t launch_log_batch(long long n, const double* x, double* out, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  k_log_batch<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, x, out);
  return cudaGetLastError();
}

cudaError_t launch_fp64_peak(double* scratch, int blocks, int threads, int iters, int fma, cudaStream_t s) {
  if (fma) k_fp64_peak<true><<<blocks, threads, 0, s>>>(scratch, iters);
  else k_fp64_peak<false><<<blocks, threads, 0, s>>>(scratch, iters);
  return cudaGetLastError();
}

cudaError_t launch_loop_stage(const LoopLaunch& a, int stage, cudaStream_t s) {
  std::lock_guard<std::mutex> key_lock(rng_key_mutex());
  {
    const cudaError_t e = ensure_rng_key_locked(a.seed);
    if (e != cudaSuccess) return e;
  }
  LoopArgs A;
  A.cfg = (const simc_run_config*)a.cfg;
  static const ArmDev no_arm = {};
  const ArmDev& arm_e = a.arm_e ? *(const ArmDev*)a.arm_e : no_arm;
  const ArmDev& arm_p = a.arm_p ? *(const ArmDev*)a.arm_p : no_arm;
  A.st.base = a.state; A.st.cap = a.cap;
  A.lists = a.lists; A.counts = a.counts; A.acc = (DevAccum*)a.acc;
  A.first_try = a.first_try; A.n_tries = a.n_tries; A.seed = a.seed; A.qexp_w = a.qexp_w;
  A.record_mode = a.record_mode;
  A.op_begin = 0; A.op_end = 0; A.in_idx = 0; A.out_idx = 0;
  static_assert(sizeof(MatTable) == sizeof(a.mats), "MatTable layout");
  memcpy(&A.mt, a.mats, sizeof(MatTable));
  A.sf.pm = a.sf_pm; A.sf.em = a.sf_em; A.sf.val = a.sf_val; A.sf.n_pm = a.sf_npm; A.sf.n_em = a.sf_nem; A.sf.dem = a.sf_dem;
  A.pdf = make_pdf(a);
  A.maid.tbl = a.maid_buf;
  A.saghai = SaghaiDev{a.saghai_buf, a.saghai_n[0], a.saghai_n[1], a.saghai_n[2]};
  A.fdss.buf = a.fdss_buf;
  A.theory.buf = a.theory_buf; A.theory.nrho = a.theory_nrho; A.theory.e_fermi = a.theory_efermi;
  A.field.map = a.field_map;
  {
    const double pi180 = 3.141592653 / 180.;          // trgInit's constant (trg_track.f:275)
    A.field.stht[0] = std::sin(a.field_theta_e_deg * pi180); A.field.ctht[0] = std::cos(a.field_theta_e_deg * pi180);
    A.field.stht[1] = std::sin(a.field_theta_p_deg * pi180); A.field.ctht[1] = std::cos(a.field_theta_p_deg * pi180);
  }
  const bool field = a.field_map != nullptr;
  A.pfm.pval = a.pfm_buf; A.pfm.mprob = a.pfm_buf ? a.pfm_buf + a.pfm_n : nullptr; A.pfm.nump = a.pfm_n;
  const long long need = (a.n_tries + kBlock - 1) / kBlock;
  const unsigned grid = (unsigned)(need < a.grid_blocks ? need : a.grid_blocks);
  {
    const cudaError_t e = ensure_smem_opt_in_locked();
    if (e != cudaSuccess) return e;
  }
  if (stage == 0) {
    cudaError_t e = cudaMemsetAsync(a.counts, 0, kLoopCounts * sizeof(unsigned), s);
    if (e != cudaSuccess) return e;
    {
      const long long gneed = (a.n_tries + kGenBlock - 1) / kGenBlock;
      const long long gmax = (long long)a.grid_blocks / 4 * SIMC_GEN_MIN_BLOCKS;     // one wave of resident CTAs (grid_blocks = 4 per SM)
      k_generate<<<(unsigned)(gneed < gmax ? gneed : gmax), kGenBlock, 0, s>>>(A);
      if (a.using_rad) k_regen<<<(unsigned)(gneed < gmax ? gneed : gmax), kGenBlock, 0, s>>>(A);
    }
  } else if (stage == 1 || stage == 2) {
    // one spectrometer: the chain of stages of its schedule, survivors handed on from list to list
    const bool hadron = stage == 1;
    const ArmSchedule& sc = hadron ? a.sched_p : a.sched_e;
    const ArmDev& arm = hadron ? arm_p : arm_e;
    const int coll = hadron ? a.coll_p : a.coll_e;
    const int base = hadron ? 0 : kArmLists;
    int in = base, out = base + 1;
    for (int k = 0; k < sc.n; ++k) {
      const ArmStage& st = sc.st[k];
      if (k == sc.n - 1) out = base + kArmLists;
      A.op_begin = st.begin; A.op_end = st.end; A.in_idx = in; A.out_idx = out;
      // kernels that run no op of the program (loop.cuh: SEG_ 7, 8): no dynamic shared memory, more CTAs per SM
      const unsigned lgrid = (unsigned)std::min<long long>(need, (long long)a.grid_blocks / 4 * SIMC_LIGHT_MIN_BLOCKS);
      if (st.kind == ARM_STAGE_CALO) {
        k_calo<<<grid, kBlock, 0, s>>>(A);
      } else if (st.kind == ARM_STAGE_ENTRY) {
        if (hadron) {
          if (field) k_arm<1, 4><<<grid, kBlock, kArmSmemBytes, s>>>(A, arm);
          else if (coll) k_arm<1, 3><<<grid, kBlock, kArmSmemBytes, s>>>(A, arm);
          else if (st.begin == st.end) k_arm<1, 8><<<lgrid, kBlock, 0, s>>>(A, arm);
          else k_arm<1, 0><<<grid, kBlock, kArmSmemBytes, s>>>(A, arm);
        } else {
          if (field) k_arm<0, 4><<<grid, kBlock, kArmSmemBytes, s>>>(A, arm);
          else if (coll) k_arm<0, 3><<<grid, kBlock, kArmSmemBytes, s>>>(A, arm);
          else if (st.begin == st.end) k_arm<0, 8><<<lgrid, kBlock, 0, s>>>(A, arm);
          else k_arm<0, 0><<<grid, kBlock, kArmSmemBytes, s>>>(A, arm);
        }
      } else if (st.kind == ARM_STAGE_COMPILED) {
        // generated straight-line kernel (mapgen.h); ABI in mapgen.h
        long long fs = state_field_stride(a.cap), ss = state_slot_stride();
        double* tk = a.state + (long long)F_TK_XS * fs;
        const unsigned* in_list = a.lists + (long long)in * a.cap;
        const unsigned* in_count = a.counts + 1 + in;
        unsigned* out_list = a.lists + (long long)out * a.cap;
        unsigned* out_count = a.counts + 1 + out;
        DevAccum* acc = (DevAccum*)a.acc;
        unsigned long long* stop_acc = &acc->stop[hadron ? 1 : 0][0];
        unsigned long long* calls_acc = &acc->transp_calls[hadron ? 1 : 0][0];
        double* stop_field = a.record_mode ? a.state + (long long)(hadron ? F_STOP_P : F_STOP_E) * fs : nullptr;
        void* args[] = {&tk, &fs, &ss, &in_list, &in_count, &out_list, &out_count, &stop_acc, &calls_acc, &stop_field};
        const int rc = jit_launch(st.fn, (unsigned)st.grid, (unsigned)st.block, (void*)s, args);
        if (rc != 0) return cudaErrorLaunchFailure;
      } else if (st.kind == ARM_STAGE_HUT) {
        // five CTAs per SM (96 registers, 40 KB of queue each): a grid of 5 x SMs x ... keeps the persistent loop balanced
        const unsigned hgrid = (unsigned)std::min<long long>(need, (long long)a.grid_blocks / 4 * SIMC_HUT_MIN_BLOCKS);
        if (hadron) k_arm<1, 6><<<hgrid, kBlock, kQueueBytes, s>>>(A, arm);
        else k_arm<0, 6><<<hgrid, kBlock, kQueueBytes, s>>>(A, arm);
      } else if (st.kind == ARM_STAGE_TAIL) {
        if (hadron) k_arm<1, 7><<<lgrid, kBlock, 0, s>>>(A, arm);
        else k_arm<0, 7><<<lgrid, kBlock, 0, s>>>(A, arm);
      } else if (st.kind == ARM_STAGE_MIDDLE) {
        if (hadron) k_arm<1, 2><<<grid, kBlock, kArmSmemBytes, s>>>(A, arm);
        else k_arm<0, 2><<<grid, kBlock, kArmSmemBytes, s>>>(A, arm);
      } else {
        if (hadron) { if (field) k_arm<1, 5><<<grid, kBlock, kArmSmemBytes, s>>>(A, arm); else k_arm<1, 1><<<grid, kBlock, kArmSmemBytes, s>>>(A, arm); }
        else { if (field) k_arm<0, 5><<<grid, kBlock, kArmSmemBytes, s>>>(A, arm); else k_arm<0, 1><<<grid, kBlock, kArmSmemBytes, s>>>(A, arm); }
      }
      in = out; out = out + 1;
    }
  } else if (stage == 3) {
    const long long fneed = (a.n_tries + kFinBlock - 1) / kFinBlock;
    const unsigned fgrid = (unsigned)std::min<long long>(fneed, (long long)a.grid_blocks / 4 * SIMC_FIN_MIN_BLOCKS);
    if (a.using_rad) k_radw<<<fgrid, kFinBlock, 0, s>>>(A, a.record_mode ? 0 : 2 * kArmLists);
    k_finish<<<fgrid, kFinBlock, 0, s>>>(A);
  }
  else if (stage == 4 && a.record_mode && a.rec) k_records<<<grid, kBlock, 0, s>>>(A, a.rec, a.status, a.n_tries);
  else if (stage == 5) {      // simc_b200_weight_batch: a.rec = input rows, a.wb_out = output rows, a.n_tries rows
    const unsigned nb = (unsigned)((a.n_tries + 255) / 256);
    k_wb_load<<<nb, 256, 0, s>>>(A, a.n_tries, a.rec);
    const long long fneed = (a.n_tries + kFinBlock - 1) / kFinBlock;
    const unsigned fgrid = (unsigned)std::min<long long>(fneed, (long long)a.grid_blocks / 4 * SIMC_FIN_MIN_BLOCKS);
    k_finish<<<fgrid, kFinBlock, 0, s>>>(A);
    k_wb_store<<<nb, 256, 0, s>>>(A, a.n_tries, a.wb_out);
  }
  return cudaGetLastError();
}

int launches_of_stage(const LoopLaunch& a, int stage) {
  if (stage == 0 || stage == 3) return a.using_rad ? 2 : 1;
  if (stage == 1) return a.sched_p.n;
  if (stage == 2) return a.sched_e.n;
  return 1;
}

}  // namespace SIMC_VARIANT_NS
}  // namespace simc
