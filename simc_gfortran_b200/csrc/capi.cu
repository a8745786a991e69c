// C ABI of libsimc_b200.so (include/simc_b200.h).  Host code only: owns the handle, the device
// tables and the stream, and launches the kernels of kernels.cu.  There is no CPU fallback:
// every compute entry point fails with SIMC_ERR_CUDA when no usable GPU is present.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/simc_b200.h"
#include "kernels.h"
#include "optics_host.h"

using namespace simc;

namespace {

struct ArmSlot {
  bool loaded = false;
  CompiledArm host;
  void* d_arm = nullptr;                 // ArmDev
  unsigned long long* d_hdr = nullptr;
  double* d_coef = nullptr;
};

std::string g_create_error;

}  // namespace

struct simc_handle {
  simc_run_config cfg;
  int device = 0;
  int strict = 1;
  cudaStream_t stream = nullptr;
  std::map<int, ArmSlot> arms;
  std::string err;
  long long launches = 0;
  // scratch for the host-pointer entry points
  double* d_in = nullptr; double* d_out = nullptr; int* d_flags = nullptr; long long scratch_n = 0;
};

namespace {

int fail(simc_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg; else g_create_error = msg;
  return code;
}
int cuda_fail(simc_handle* h, cudaError_t e, const char* what) {
  return fail(h, SIMC_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(h, call)                                                   \
  do {                                                                \
    cudaError_t e_ = (call);                                          \
    if (e_ != cudaSuccess) return cuda_fail(h, e_, #call);            \
  } while (0)

void free_arm(ArmSlot& s) {
  if (s.d_arm) cudaFree(s.d_arm);
  if (s.d_hdr) cudaFree(s.d_hdr);
  if (s.d_coef) cudaFree(s.d_coef);
  s = ArmSlot();
}

int upload_arm(simc_handle* h, int arm_id, CompiledArm&& ca) {
  CU(h, cudaSetDevice(h->device));
  ArmSlot& s = h->arms[arm_id];
  free_arm(s);
  s.host = std::move(ca);
  CU(h, cudaMalloc(&s.d_hdr, s.host.hdr.size() * sizeof(unsigned long long)));
  CU(h, cudaMalloc(&s.d_coef, s.host.coef.size() * sizeof(double)));
  CU(h, cudaMemcpy(s.d_hdr, s.host.hdr.data(), s.host.hdr.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice));
  CU(h, cudaMemcpy(s.d_coef, s.host.coef.data(), s.host.coef.size() * sizeof(double), cudaMemcpyHostToDevice));
  // ArmDev = { ArmTablesDev tab; ArmOp ops[kMaxArmOps]; } -- identical layout in both variants
  const size_t bytes = strict::arm_dev_bytes();
  std::vector<unsigned char> img(bytes, 0);
  ArmTablesDev tab = s.host.tab;
  tab.hdr = s.d_hdr;
  tab.coef = s.d_coef;
  std::memcpy(img.data(), &tab, sizeof(tab));
  std::memcpy(img.data() + sizeof(ArmTablesDev), s.host.ops.data(), s.host.ops.size() * sizeof(ArmOp));
  CU(h, cudaMalloc(&s.d_arm, bytes));
  CU(h, cudaMemcpy(s.d_arm, img.data(), bytes, cudaMemcpyHostToDevice));
  s.loaded = true;
  return SIMC_OK;
}

int ensure_scratch(simc_handle* h, long long n) {
  if (n <= h->scratch_n) return SIMC_OK;
  if (h->d_in) cudaFree(h->d_in);
  if (h->d_out) cudaFree(h->d_out);
  if (h->d_flags) cudaFree(h->d_flags);
  h->d_in = nullptr; h->d_out = nullptr; h->d_flags = nullptr; h->scratch_n = 0;
  CU(h, cudaMalloc(&h->d_in, sizeof(double) * SIMC_TRANSPORT_NIN * n));
  CU(h, cudaMalloc(&h->d_out, sizeof(double) * SIMC_TRANSPORT_NOUT * n));
  CU(h, cudaMalloc(&h->d_flags, sizeof(int) * n));
  h->scratch_n = n;
  return SIMC_OK;
}

}  // namespace

extern "C" {

int simc_b200_abi_version(void) { return SIMC_B200_ABI_VERSION; }

int simc_b200_create(const simc_run_config* cfg, int device, simc_handle** out) {
  if (!out) return fail(nullptr, SIMC_ERR_ARG, "simc_b200_create: out is NULL");
  *out = nullptr;
  if (cfg && cfg->abi_version != SIMC_B200_ABI_VERSION)
    return fail(nullptr, SIMC_ERR_ARG, "simc_b200_create: simc_run_config.abi_version mismatch");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, SIMC_ERR_CUDA, std::string("simc_b200_create: no CUDA device (") +
                                            (e != cudaSuccess ? cudaGetErrorString(e) : "count = 0") +
                                            "); libsimc_b200 has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(nullptr, SIMC_ERR_ARG, "simc_b200_create: bad device index");
  simc_handle* h = new (std::nothrow) simc_handle();
  if (!h) return fail(nullptr, SIMC_ERR_ARG, "out of memory");
  if (cfg) h->cfg = *cfg; else std::memset(&h->cfg, 0, sizeof(h->cfg));
  h->device = device;
  const char* mode = std::getenv("SIMC_B200_MODE");
  h->strict = !(mode && std::strcmp(mode, "fast") == 0);
  if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) {
    std::string m = std::string("simc_b200_create: ") + cudaGetErrorString(e);
    delete h;
    return fail(nullptr, SIMC_ERR_CUDA, m);
  }
  *out = h;
  return SIMC_OK;
}

void simc_b200_destroy(simc_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  for (auto& kv : h->arms) free_arm(kv.second);
  if (h->d_in) cudaFree(h->d_in);
  if (h->d_out) cudaFree(h->d_out);
  if (h->d_flags) cudaFree(h->d_flags);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

const char* simc_b200_last_error(const simc_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int simc_b200_set_mode(simc_handle* h, int strict_mode) {
  if (!h) return SIMC_ERR_ARG;
  h->strict = strict_mode ? 1 : 0;
  return SIMC_OK;
}

int simc_b200_load_optics(simc_handle* h, int arm_id, const char* forward_path, const char* recon_path) {
  if (!h || !forward_path || !recon_path) return fail(h, SIMC_ERR_ARG, "simc_b200_load_optics: NULL argument");
  try {
    ForwardMaps f = read_forward_maps(forward_path);
    CosyTerms r = read_recon_map(recon_path);
    return upload_arm(h, arm_id, compile_arm(arm_id, f, r));
  } catch (const std::exception& e) {
    return fail(h, SIMC_ERR_IO, e.what());
  }
}

int simc_b200_set_optics(simc_handle* h, int arm_id, int n_classes, const int32_t* fwd_class_start,
                         const double* fwd_coeff, const int8_t* fwd_expon, const double* fwd_length_cm, int n_rec,
                         const double* rec_coeff, const int8_t* rec_expon) {
  if (!h || !fwd_class_start || !fwd_coeff || !fwd_expon || !rec_coeff || !rec_expon || n_classes <= 0 || n_rec <= 0)
    return fail(h, SIMC_ERR_ARG, "simc_b200_set_optics: bad argument");
  try {
    ForwardMaps f;
    for (int k = 0; k < n_classes; ++k) {
      CosyTerms t;
      const int b = fwd_class_start[k], e = fwd_class_start[k + 1];
      if (e < b) throw std::runtime_error("fwd_class_start must be non-decreasing");
      t.coef.assign(fwd_coeff + 5 * (size_t)b, fwd_coeff + 5 * (size_t)e);
      t.expo.assign(fwd_expon + 5 * (size_t)b, fwd_expon + 5 * (size_t)e);
      f.cls.push_back(std::move(t));
      f.length_cm.push_back(fwd_length_cm ? fwd_length_cm[k] : 0.0);
    }
    classify_drifts(f);
    CosyTerms r;
    r.nout = 4;
    r.coef.assign(rec_coeff, rec_coeff + 4 * (size_t)n_rec);
    r.expo.assign(rec_expon, rec_expon + 5 * (size_t)n_rec);
    return upload_arm(h, arm_id, compile_arm(arm_id, f, r));
  } catch (const std::exception& e) {
    return fail(h, SIMC_ERR_IO, e.what());
  }
}

int simc_b200_optics_info(simc_handle* h, int arm_id, int64_t* info8) {
  if (!h || !info8) return SIMC_ERR_ARG;
  auto it = h->arms.find(arm_id);
  if (it == h->arms.end() || !it->second.loaded) return fail(h, SIMC_ERR_STATE, "optics not loaded for this arm");
  const CompiledArm& c = it->second.host;
  info8[0] = c.tab.n_classes; info8[1] = c.fwd_terms; info8[2] = c.fwd_nonzero; info8[3] = c.rec_terms;
  info8[4] = (int64_t)c.hdr.size(); info8[5] = (int64_t)c.coef.size(); info8[6] = (int64_t)c.ops.size(); info8[7] = 0;
  return SIMC_OK;
}

int simc_b200_transport_batch_device(simc_handle* h, int arm_id, int64_t n, const double* d_in_soa, uint64_t seed,
                                     int ms_flag, int wcs_flag, int decay_flag, int using_coll, double* d_out_soa,
                                     int32_t* d_flags) {
  if (!h) return SIMC_ERR_ARG;
  if (n < 0 || (n > 0 && (!d_in_soa || !d_out_soa || !d_flags)))
    return fail(h, SIMC_ERR_ARG, "simc_b200_transport_batch: bad argument");
  auto it = h->arms.find(arm_id);
  if (it == h->arms.end() || !it->second.loaded)
    return fail(h, SIMC_ERR_STATE, "simc_b200_transport_batch: optics not loaded for this arm");
  if (using_coll) return fail(h, SIMC_ERR_ARG, "collimator stepping (mc_hms_coll/mc_shms_coll) is not implemented");
  if (n == 0) return SIMC_OK;
  CU(h, cudaSetDevice(h->device));
  TransportBatchArgs a;
  a.arm = it->second.d_arm; a.n = n; a.in = d_in_soa; a.seed = seed;
  a.ms_flag = ms_flag; a.wcs_flag = wcs_flag; a.decay_flag = decay_flag; a.using_coll = using_coll;
  a.ctau = h->cfg.ctau; a.out = d_out_soa; a.flags = d_flags;
  cudaError_t e = h->strict ? strict::launch_transport_batch(a, h->stream) : fast::launch_transport_batch(a, h->stream);
  if (e != cudaSuccess) return cuda_fail(h, e, "k_transport_batch launch");
  h->launches += 1;
  return SIMC_OK;
}

int simc_b200_transport_batch(simc_handle* h, int arm_id, int64_t n, const double* in_soa, uint64_t seed, int ms_flag,
                              int wcs_flag, int decay_flag, int using_coll, double* out_soa, int32_t* flags) {
  if (!h) return SIMC_ERR_ARG;
  if (n < 0 || (n > 0 && (!in_soa || !out_soa || !flags)))
    return fail(h, SIMC_ERR_ARG, "simc_b200_transport_batch: bad argument");
  if (n == 0) {
    // still validate the state so that empty batches report the same errors
    auto it = h->arms.find(arm_id);
    if (it == h->arms.end() || !it->second.loaded)
      return fail(h, SIMC_ERR_STATE, "simc_b200_transport_batch: optics not loaded for this arm");
    return SIMC_OK;
  }
  CU(h, cudaSetDevice(h->device));
  int rc = ensure_scratch(h, n);
  if (rc) return rc;
  CU(h, cudaMemcpyAsync(h->d_in, in_soa, sizeof(double) * SIMC_TRANSPORT_NIN * n, cudaMemcpyHostToDevice, h->stream));
  rc = simc_b200_transport_batch_device(h, arm_id, n, h->d_in, seed, ms_flag, wcs_flag, decay_flag, using_coll,
                                        h->d_out, h->d_flags);
  if (rc) return rc;
  CU(h, cudaMemcpyAsync(out_soa, h->d_out, sizeof(double) * SIMC_TRANSPORT_NOUT * n, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaMemcpyAsync(flags, h->d_flags, sizeof(int) * n, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return SIMC_OK;
}

void* simc_b200_stream(simc_handle* h) { return h ? (void*)h->stream : nullptr; }
int64_t simc_b200_launch_count(const simc_handle* h) { return h ? h->launches : 0; }
int simc_b200_sync(simc_handle* h) {
  if (!h) return SIMC_ERR_ARG;
  CU(h, cudaSetDevice(h->device));
  CU(h, cudaStreamSynchronize(h->stream));
  return SIMC_OK;
}

const char* simc_b200_stop_name(int arm_id, int code) { return stop_name(arm_id, code); }

int64_t simc_b200_sizeof(int which) {
  return which == 0 ? (int64_t)sizeof(simc_run_config) : which == 1 ? (int64_t)sizeof(simc_accum) : -1;
}

}  // extern "C"

// ---- the event loop (implemented in loop.cu once generation/weights are in) ---------------
extern "C" {
#ifndef SIMC_HAVE_LOOP
int simc_b200_accum_clear(simc_handle* h, simc_accum* acc) {
  if (!h || !acc) return SIMC_ERR_ARG;
  std::memset(acc, 0, sizeof(*acc));
  return SIMC_OK;
}
int simc_b200_run(simc_handle* h, int64_t, int64_t, uint64_t, simc_accum*) {
  return fail(h, SIMC_ERR_STATE, "simc_b200_run: event loop not built into this library yet");
}
int simc_b200_run_async(simc_handle* h, int64_t, int64_t, uint64_t) {
  return fail(h, SIMC_ERR_STATE, "simc_b200_run_async: event loop not built into this library yet");
}
int simc_b200_fetch(simc_handle* h, simc_accum*) {
  return fail(h, SIMC_ERR_STATE, "simc_b200_fetch: event loop not built into this library yet");
}
int simc_b200_device_accum(simc_handle* h, void**, int64_t*, void**, int64_t*) {
  return fail(h, SIMC_ERR_STATE, "simc_b200_device_accum: event loop not built into this library yet");
}
int simc_b200_event_batch(simc_handle* h, int64_t, int64_t, uint64_t, double*, int32_t*) {
  return fail(h, SIMC_ERR_STATE, "simc_b200_event_batch: event loop not built into this library yet");
}
const char* simc_b200_event_field_name(int) { return ""; }
#endif
}
