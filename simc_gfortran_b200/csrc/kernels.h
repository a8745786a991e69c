// Host-callable launchers of the two kernel variants (strict = -fmad=false, reference
// association; fast = FMA + re-associated monomials).  Implemented in kernels.cu, which is
// compiled twice.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace simc {

struct TransportBatchArgs {
  const void* arm;          // ArmDev* (device)
  long long n;
  const double* in;         // [9][n] device
  unsigned long long seed;
  int ms_flag, wcs_flag, decay_flag, using_coll;
  double ctau;
  double* out;              // [12][n] device
  int* flags;               // [n] device
};

namespace strict { cudaError_t launch_transport_batch(const TransportBatchArgs& a, cudaStream_t s); size_t arm_dev_bytes(); }
namespace fast   { cudaError_t launch_transport_batch(const TransportBatchArgs& a, cudaStream_t s); size_t arm_dev_bytes(); }

}  // namespace simc
