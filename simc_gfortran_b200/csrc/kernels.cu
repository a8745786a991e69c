// CUDA kernels of libsimc_b200 (sm_100a).  Compiled twice: -DSIMC_STRICT=1 -fmad=false
// (namespace simc::strict) and -DSIMC_STRICT=0 -fmad=true (namespace simc::fast).
#if SIMC_STRICT
#define SIMC_VARIANT_NS strict
#else
#define SIMC_VARIANT_NS fast
#endif
#include "transport.cuh"
#include "kernels.h"

namespace simc {
namespace SIMC_VARIANT_NS {

// Batch form of mc_hms / mc_shms / ... (hms/mc_hms.f:1-4): one thread per row.
__global__ void __launch_bounds__(kBlock)
k_transport_batch(const ArmDev* __restrict__ arm, long long n, const double* __restrict__ in,
                  unsigned long long seed, ArmFlags f, double ctau, double* __restrict__ out,
                  int* __restrict__ flags) {
  __shared__ double pw_s[kPowDoubles];
  const long long i = (long long)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  TrackDev t;
  t.dpps = in[0 * n + i];
  t.xs = in[1 * n + i];
  t.ys = in[2 * n + i];
  // in[3] = z is carried by the reference (zs) but never read by any single-arm routine
  t.dxdzs = in[4 * n + i];
  t.dydzs = in[5 * n + i];
  t.m2 = in[6 * n + i];
  const double p_spec = in[7 * n + i];
  const double fry = in[8 * n + i];
  t.p = p_spec * (1. + t.dpps / 100.);     // mc_hms.f:181
  t.pathlen = 0.0;
  t.decdist = 0.0;
  t.mh2_final = t.m2;
  t.ctau = ctau;
  const double dpp_in = t.dpps, y_in = t.ys, dxdz_in = t.dxdzs, dydz_in = t.dydzs;
  DevRng rng;
  rng.init(seed, (unsigned long long)i, 0u, 0u);
  ArmResult res;
  run_arm(arm, t, rng, f, fry, pw_s + threadIdx.x, res);
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

cudaError_t launch_transport_batch(const TransportBatchArgs& a, cudaStream_t s) {
  if (a.n <= 0) return cudaSuccess;
  const long long blocks = (a.n + kBlock - 1) / kBlock;
  ArmFlags f;
  f.ms_flag = a.ms_flag != 0; f.wcs_flag = a.wcs_flag != 0; f.decay_flag = a.decay_flag != 0;
  f.using_coll = a.using_coll != 0;
  k_transport_batch<<<(unsigned)blocks, kBlock, 0, s>>>((const ArmDev*)a.arm, a.n, a.in, a.seed, f, a.ctau, a.out,
                                                        a.flags);
  return cudaGetLastError();
}

size_t arm_dev_bytes() { return sizeof(ArmDev); }

}  // namespace SIMC_VARIANT_NS
}  // namespace simc
