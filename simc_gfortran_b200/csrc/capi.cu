// C ABI of libsimc_b200.so (include/simc_b200.h).  Host code only: owns the handle, the device
// tables and the stream, and launches the kernels of kernels.cu.  There is no CPU fallback:
// every compute entry point fails with SIMC_ERR_CUDA when no usable GPU is present.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/simc_b200.h"
#include "kernels.h"
#include "mapgen.h"
#include "jit.h"
#include "optics_host.h"
#include "target.cuh"
#include "tables_host.h"
#include "physics_semi.cuh"

using namespace simc;

size_t simc_dev_accum_minmax_offset();      // kernels.cu

namespace {

struct ArmSlot {
  bool loaded = false;
  CompiledArm host;
  std::vector<unsigned char> img;        // ArmDev image (host): kernels take it by value
  double* d_recs = nullptr;
  // compiled stretches of the program (mapgen.h): one generated kernel per RNG-free stretch
  std::vector<StretchSpec> stretches;
  int hut_begin = 0;                     // first op that is not static
  int recon_pc = -1;                     // the program's OP_RECON (-1: none); compiled as one more kernel, jit.fns[stretches.size()]
  JitModule jit;
  bool jit_ready = false;
  int jit_strict = -1;                   // arithmetic variant the module was generated for
};

std::string g_create_error;

}  // namespace

struct simc_handle {
  simc_run_config cfg;
  int device = 0;
  int strict = 1;
  int compiled_maps = 1;               // run RNG-free stretches of the arm programs as generated kernels (mapgen.h)
  int compiled_recon = 1;              // ... and the reconstruction map behind the hut (SIMC_B200_COMPILED_RECON=0: interpreter)
  cudaStream_t stream = nullptr;
  std::map<int, ArmSlot> arms;
  std::string err;
  long long launches = 0;
  // scratch for the host-pointer entry points
  double* d_in = nullptr; double* d_out = nullptr; int* d_flags = nullptr; long long scratch_n = 0;
  // scratch of the compiled path of transport_batch: track rows, survivor lists, counters (kernels.h)
  double* tb_tk = nullptr; unsigned* tb_lists = nullptr; unsigned* tb_counts = nullptr; unsigned long long* tb_sink = nullptr;
  long long tb_n = 0;
  // event loop
  simc_run_config* d_cfg = nullptr;
  double* d_state = nullptr; unsigned* d_lists = nullptr; unsigned* d_counts = nullptr; void* d_acc = nullptr;
  long long loop_cap = 0;
  long long batch = 1 << 20;           // tries per batch of the stage pipeline
  int grid_blocks = 0;
  int qexp_w = -64;
  std::vector<unsigned char> acc_host;
  double* d_rec = nullptr; int* d_status = nullptr; long long rec_n = 0;
  double* d_sf = nullptr; int sf_npm = 0, sf_nem = 0;      // Benhar spectral function: [pm | em | val]
  double* d_sf_dem = nullptr;                              // widths of its Em bins (generate_em)
  double* d_pdf = nullptr; int pdf_nx = 0, pdf_nt = 0, pdf_nfmx = 0; double pdf_al = 0;   // CTEQ5: [xv | ql | upd]
  double* d_pfm = nullptr; int pfm_n = 0;                  // momentum distribution: [pval | mprob]
  double* d_fdss = nullptr;                                // fDSS tables (physics_semi.cuh: FdssDev)
  float* d_saghai[2] = {nullptr, nullptr};                 // Saghai amplitude tables: [0] K+ Lambda, [1] K+ Sigma0
  double* d_field = nullptr;                               // target field map (field.cuh: FieldDev), [Bz | Br] of 51 x 51 nodes
  double* d_maid[2] = {nullptr, nullptr};                  // MAID-2007 slices: [0] pi+ n (ipi 3), [1] pi- p (ipi 4)
  double* d_theory = nullptr; int theory_nrho = 0; double theory_efermi = 0;   // physics_heavy.cuh: TheoryDev
  // optional per-stage timing
  int timing = 0;
  std::vector<cudaEvent_t> ev;                 // 5 events per batch, recycled
  std::vector<int> ev_used;                    // number of batches whose events are pending
  double stage_ms[4] = {0, 0, 0, 0};
  long long stage_launches[4] = {0, 0, 0, 0};
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
  if (s.d_recs) cudaFree(s.d_recs);
  jit_unload(s.jit);
  s = ArmSlot();
}

// Cuts of the arm program for the compiled path: RNG-free stretches from op 0 up to the first op that needs the
// interpreter (the hut), split at the loop's compaction points.
void plan_stretches(ArmSlot& s) {
  const std::vector<ArmOp>& ops = s.host.ops;
  int hut = 0;
  while (hut < (int)ops.size() && op_is_static(ops[hut].op)) ++hut;
  s.hut_begin = hut;
  s.stretches.clear();
  std::vector<int> cuts;
  cuts.push_back(0);
  const ArmTablesDev& tab = s.host.tab;
  if (tab.split_op > 0 && tab.split_op < hut) cuts.push_back(tab.split_op);
  for (int k = 0; k < tab.n_mid; ++k)
    if (tab.mid_op[k] > cuts.back() && tab.mid_op[k] < hut) cuts.push_back(tab.mid_op[k]);
  cuts.push_back(hut);
  for (size_t k = 0; k + 1 < cuts.size(); ++k)
    if (cuts[k + 1] > cuts[k]) s.stretches.push_back(StretchSpec{cuts[k], cuts[k + 1]});
  s.recon_pc = -1;
  for (int pc = hut; pc < (int)ops.size(); ++pc)
    if (ops[pc].op == OP_RECON) { s.recon_pc = pc; break; }
}
// What goes into one generated translation unit: the stretches, then the reconstruction map as a kernel of its own
std::vector<StretchSpec> source_specs(const ArmSlot& s) {
  std::vector<StretchSpec> v = s.stretches;
  if (s.recon_pc >= 0) v.push_back(StretchSpec{s.recon_pc, s.recon_pc + 1});
  return v;
}

// Launch shape of the generated kernels: two CTAs of 256 threads per SM (128 registers per thread).  The warps of a
// CTA enter every map together (a __syncthreads in front of each), so an SM streams the straight-line code through
// its instruction cache twice, not once per warp.  Measured on C1 (profiles/README.md, step 14): 128x3, 192x2,
// 384x1, 512x1 within 1 %; 256x1 and 128x2 (fewer warps) 3 % slower.
int map_block_threads() {
  if (const char* e = std::getenv("SIMC_B200_MAP_BLOCK")) { const int v = std::atoi(e); if (v >= 64 && v <= 1024 && v % 32 == 0) return v; }
  return 256;
}
int map_min_blocks() {
  if (const char* e = std::getenv("SIMC_B200_MAP_MINBLOCKS")) { const int v = std::atoi(e); if (v >= 1 && v <= 8) return v; }
  return 2;
}

int upload_arm(simc_handle* h, int arm_id, CompiledArm&& ca) {
  CU(h, cudaSetDevice(h->device));
  ArmSlot& s = h->arms[arm_id];
  free_arm(s);
  s.host = std::move(ca);
  CU(h, cudaMalloc(&s.d_recs, s.host.recs.size() * sizeof(double)));       // cudaMalloc: 256-byte aligned
  CU(h, cudaMemcpy(s.d_recs, s.host.recs.data(), s.host.recs.size() * sizeof(double), cudaMemcpyHostToDevice));
  // ArmDev = { ArmTablesDev tab; ArmOp ops[kMaxArmOps]; } -- identical layout in both variants
  const size_t bytes = strict::arm_dev_bytes();
  std::vector<unsigned char>& img = s.img;
  img.assign(bytes, 0);
  ArmTablesDev tab = s.host.tab;
  tab.recs = s.d_recs;
  tab.pad_ptr = nullptr;
  std::memcpy(img.data(), &tab, sizeof(tab));
  std::memcpy(img.data() + sizeof(ArmTablesDev), s.host.ops.data(), s.host.ops.size() * sizeof(ArmOp));
  s.loaded = true;
  return SIMC_OK;
}

int ensure_compiled(simc_handle* h, ArmSlot& s);

int ensure_tb_scratch(simc_handle* h, long long n) {
  if (n <= h->tb_n) return SIMC_OK;
  if (h->tb_tk) cudaFree(h->tb_tk);
  if (h->tb_lists) cudaFree(h->tb_lists);
  h->tb_tk = nullptr; h->tb_lists = nullptr; h->tb_n = 0;
  CU(h, cudaMalloc(&h->tb_tk, sizeof(double) * 12 * (size_t)n));
  CU(h, cudaMalloc(&h->tb_lists, sizeof(unsigned) * (kArmLists + 1) * (size_t)n));
  if (!h->tb_counts) CU(h, cudaMalloc(&h->tb_counts, sizeof(unsigned) * (kArmLists + 1)));
  if (!h->tb_sink) CU(h, cudaMalloc(&h->tb_sink, sizeof(unsigned long long) * (SIMC_NSTOP + 48)));
  h->tb_n = n;
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
  const char* cm = std::getenv("SIMC_B200_COMPILED_MAPS");
  h->compiled_maps = !(cm && std::strcmp(cm, "0") == 0);
  const char* cr = std::getenv("SIMC_B200_COMPILED_RECON");
  h->compiled_recon = !(cr && std::strcmp(cr, "0") == 0);
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
  if (h->d_cfg) cudaFree(h->d_cfg);
  if (h->d_state) cudaFree(h->d_state);
  if (h->d_lists) cudaFree(h->d_lists);
  if (h->tb_tk) cudaFree(h->tb_tk);
  if (h->tb_lists) cudaFree(h->tb_lists);
  if (h->tb_counts) cudaFree(h->tb_counts);
  if (h->tb_sink) cudaFree(h->tb_sink);
  if (h->d_counts) cudaFree(h->d_counts);
  if (h->d_acc) cudaFree(h->d_acc);
  if (h->d_rec) cudaFree(h->d_rec);
  if (h->d_status) cudaFree(h->d_status);
  if (h->d_sf) cudaFree(h->d_sf);
  if (h->d_sf_dem) cudaFree(h->d_sf_dem);
  if (h->d_pdf) cudaFree(h->d_pdf);
  if (h->d_pfm) cudaFree(h->d_pfm);
  if (h->d_theory) cudaFree(h->d_theory);
  for (double* p : h->d_maid) if (p) cudaFree(p);
  if (h->d_field) cudaFree(h->d_field);
  for (float* p : h->d_saghai) if (p) cudaFree(p);
  if (h->d_fdss) cudaFree(h->d_fdss);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

// sf_lookup_init (sf_lookup.f:1-80) from arrays
int simc_b200_set_sf_table(simc_handle* h, int n_pm, int n_em, const double* pm, const double* em, const double* sf) {
  if (!h || !pm || !em || !sf) return SIMC_ERR_ARG;
  if (n_pm < 2 || n_em < 2 || n_pm > 100 || n_em > 200)
    return fail(h, SIMC_ERR_ARG, "spectral function: 2..100 Pm bins and 2..200 Em bins (sf_lookup.inc)");
  CU(h, cudaSetDevice(h->device));
  std::vector<double> img((size_t)n_pm + n_em + (size_t)n_pm * n_em);
  std::copy(pm, pm + n_pm, img.begin());
  std::copy(em, em + n_em, img.begin() + n_pm);
  double sftotnorm = 0.0;
  for (size_t i = 0; i < (size_t)n_pm * n_em; ++i) sftotnorm = sftotnorm + sf[i];
  if (!(sftotnorm > 0)) return fail(h, SIMC_ERR_ARG, "spectral function sums to zero");
  for (size_t i = 0; i < (size_t)n_pm * n_em; ++i) img[(size_t)n_pm + n_em + i] = sf[i] / sftotnorm;
  if (h->d_sf) { cudaFree(h->d_sf); h->d_sf = nullptr; }
  CU(h, cudaMalloc(&h->d_sf, img.size() * sizeof(double)));
  CU(h, cudaMemcpy(h->d_sf, img.data(), img.size() * sizeof(double), cudaMemcpyHostToDevice));
  h->sf_npm = n_pm; h->sf_nem = n_em;
  if (h->d_sf_dem) { cudaFree(h->d_sf_dem); h->d_sf_dem = nullptr; }       // belongs to the previous table
  return SIMC_OK;
}

int simc_b200_set_sf_em_widths(simc_handle* h, int n_em, const double* dem) {
  if (!h || !dem) return SIMC_ERR_ARG;
  if (!h->d_sf || n_em != h->sf_nem) return fail(h, SIMC_ERR_STATE, "set_sf_em_widths: set the spectral function first (same number of Em bins)");
  CU(h, cudaSetDevice(h->device));
  if (!h->d_sf_dem) CU(h, cudaMalloc(&h->d_sf_dem, sizeof(double) * n_em));
  CU(h, cudaMemcpy(h->d_sf_dem, dem, sizeof(double) * n_em, cudaMemcpyHostToDevice));
  return SIMC_OK;
}

// The reference's file format (benharsf_*.dat): "numPm numEm", then numPm*numEm rows
// "Pm Em S_proton S_neutron dPm dEm", Em running fastest (sf_lookup.f:17-49).
int simc_b200_load_sf_file(simc_handle* h, const char* path, int proton_flag) {
  if (!h || !path) return SIMC_ERR_ARG;
  FILE* f = std::fopen(path, "r");
  if (!f) return fail(h, SIMC_ERR_IO, std::string("cannot open spectral function file ") + path);
  int n_pm = 0, n_em = 0;
  if (std::fscanf(f, "%d %d", &n_pm, &n_em) != 2 || n_pm < 2 || n_em < 2 || n_pm > 100 || n_em > 200) {
    std::fclose(f);
    return fail(h, SIMC_ERR_IO, "spectral function file: bad header");
  }
  std::vector<double> pm(n_pm), em(n_em), sf((size_t)n_pm * n_em), dem(n_em);
  for (int i = 0; i < n_pm; ++i)
    for (int j = 0; j < n_em; ++j) {
      double tPm, tEm, sp, sn, dPm, dEm;
      if (std::fscanf(f, "%lf %lf %lf %lf %lf %lf", &tPm, &tEm, &sp, &sn, &dPm, &dEm) != 6) {
        std::fclose(f);
        return fail(h, SIMC_ERR_IO, "spectral function file: short read");
      }
      sf[(size_t)i * n_em + j] = proton_flag ? sp : sn;
      if (j == 0) pm[i] = tPm;
      if (i == 0) { em[j] = tEm; dem[j] = dEm; }
    }
  std::fclose(f);
  const int rc = simc_b200_set_sf_table(h, n_pm, n_em, pm.data(), em.data(), sf.data());
  if (rc) return rc;
  return simc_b200_set_sf_em_widths(h, n_em, dem.data());
}

// First-call initialisation of fDSS (fdss/fdss.f:96-125) from the rows of a *.GRID file
int simc_b200_set_fdss_table(simc_handle* h, const double* parton) {
  if (!h || !parton) return SIMC_ERR_ARG;
  static const double QS[24] = {1., 1.25, 1.5, 2.5, 4.0, 6.4, 1.0e1, 1.5e1, 2.5e1, 4.0e1, 6.4e1, 1.0e2, 1.8e2, 3.2e2,
                                5.8e2, 1.0e3, 1.8e3, 3.2e3, 5.8e3, 1.0e4, 1.8e4, 3.2e4, 5.8e4, 1.0e5};
  static const double XB[35] = {0.01, 0.02, 0.03, 0.04, 0.05, 0.06, 0.07, 0.08, 0.09, 0.095, 0.1, 0.125, 0.15, 0.175,
                                0.2, 0.225, 0.25, 0.275, 0.3, 0.325, 0.35, 0.375, 0.4, 0.45, 0.5, 0.55, 0.6, 0.65,
                                0.7, 0.75, 0.8, 0.85, 0.9, 0.93, 1.0};
  const int NX = 35, NQ = 24;
  const int col[6] = {0, 1, 2, 6, 7, 8};       // UTOT, DTOT, STOT, UVAL, DVAL, SVAL
  std::vector<double> img(kFdssTab0 + 6 * 840, 0.0);
  for (int ix = 0; ix < NX; ++ix) img[ix] = std::log(XB[ix]);
  for (int iq = 0; iq < NQ; ++iq) img[NX + iq] = std::log(QS[iq]);
  for (int iq = 0; iq < NQ; ++iq)
    for (int ix = 0; ix < NX - 1; ++ix) {
      const double XB0 = XB[ix], XB1 = 1. - XB[ix];
      const double x2 = XB1 * XB1;
      for (int k = 0; k < 6; ++k)
        img[kFdssTab0 + k * 840 + iq * NX + ix] = parton[((size_t)ix * NQ + iq) * 9 + col[k]] / ((x2 * x2) * std::pow(XB0, 0.5));
    }
  CU(h, cudaSetDevice(h->device));
  if (!h->d_fdss) CU(h, cudaMalloc(&h->d_fdss, img.size() * sizeof(double)));
  CU(h, cudaMemcpy(h->d_fdss, img.data(), img.size() * sizeof(double), cudaMemcpyHostToDevice));
  return SIMC_OK;
}

// fdss/*.GRID: 34 x 24 rows of '9(1PE10.3)' (fdss/fdss.f:96-104)
int simc_b200_load_fdss_file(simc_handle* h, const char* path) {
  if (!h || !path) return SIMC_ERR_ARG;
  FILE* f = std::fopen(path, "r");
  if (!f) return fail(h, SIMC_ERR_IO, std::string("cannot open fragmentation-function grid ") + path);
  std::vector<double> parton((size_t)34 * 24 * 9);
  char line[256];
  bool ok = true;
  for (int r = 0; r < 34 * 24 && ok; ++r) {
    if (!std::fgets(line, sizeof line, f) || std::strlen(line) < 90) { ok = false; break; }
    for (int k = 0; k < 9; ++k) {
      char field[16];
      std::memcpy(field, line + 10 * k, 10);
      field[10] = 0;
      parton[(size_t)r * 9 + k] = std::atof(field);
    }
  }
  std::fclose(f);
  if (!ok) return fail(h, SIMC_ERR_IO, "fragmentation-function grid: short or malformed file");
  return simc_b200_set_fdss_table(h, parton.data());
}

// Saghai amplitude tables (simulate.inc:188-195): which = 0: zrff1..6, ziff1..6 of K+ Lambda, each (10,11,19);
// which = 1: zsrff1..6, zsiff1..6 of K+ Sigma0, each (20,10,19); Fortran storage order.  The grids eekeek / eekeeks
// build in their `pa` arrays (physics_kaon.f:285-304, 398-431: REAL*8 sums stored as REAL*4) go in front.
int simc_b200_set_saghai_table(simc_handle* h, int which, const float* tbl) {
  if (!h || !tbl || which < 0 || which > 1) return SIMC_ERR_ARG;
  const int n1 = which ? 20 : 10, n2 = which ? 10 : 11, n3 = 19;
  const size_t n_tab = (size_t)n1 * n2 * n3;
  std::vector<float> buf(64 + 12 * n_tab, 0.f);
  if (which == 0) {
    double ps = 2.6, qs = 0.0, as = 0.0;
    for (int i = 0; i < 10; ++i) { buf[i] = (float)ps; ps = ps + 0.3; }
    for (int i = 10; i < 21; ++i) { buf[i] = (float)qs; qs = qs + 0.2; }
    for (int i = 21; i < 40; ++i) { buf[i] = (float)as; as = as + 10.; }
  } else {
    static const double grid[30] = {2.851, 2.898, 2.945, 2.991, 3.038, 3.085, 3.132, 3.320, 3.507, 3.695,
                                    3.883, 4.070, 4.258, 4.446, 4.633, 4.821, 5.009, 5.196, 5.384, 5.572,
                                    0.0,   0.250, 0.376, 0.520, 0.750, 1.000, 1.250, 1.500, 1.750, 2.000};
    for (int i = 0; i < 30; ++i) buf[i] = (float)grid[i];
    double as = 0.0;
    for (int i = 30; i < 49; ++i) { buf[i] = (float)as; as = as + 10.; }
  }
  std::memcpy(buf.data() + 64, tbl, sizeof(float) * 12 * n_tab);
  CU(h, cudaSetDevice(h->device));
  float*& d = h->d_saghai[which];
  if (!d) CU(h, cudaMalloc(&d, sizeof(float) * buf.size()));
  CU(h, cudaMemcpy(d, buf.data(), sizeof(float) * buf.size(), cudaMemcpyHostToDevice));
  return SIMC_OK;
}

// saghai_proton.dat / saghai_sigma0.dat as dbase.f:644-679 reads them: per (iread, iq2) of the Lambda file one
// header line, then per angle a line of five kinematic numbers and two lines '(6e12.4)' of (re, im) pairs; the
// Sigma0 file repeats the header line for every angle.
namespace {
bool read_saghai_file(const std::string& path, int which, std::vector<float>& tbl, std::string& err) {
  const int n1 = which ? 20 : 10, n2 = which ? 10 : 11, n3 = 19;
  const size_t n_tab = (size_t)n1 * n2 * n3;
  tbl.assign(12 * n_tab, 0.f);
  FILE* f = std::fopen(path.c_str(), "r");
  if (!f) { err = "cannot open " + path; return false; }
  char line[256];
  // read(3,*) dum1,dum2: list-directed, two numbers from the next line
  auto next_list2 = [&]() -> bool {
    if (!std::fgets(line, sizeof(line), f)) return false;
    char* p = line;
    for (int got = 0; got < 2; ++got) {
      char* e = nullptr;
      std::strtod(p, &e);
      if (e == p) return false;
      p = e;
    }
    return true;
  };
  // read(3,'(6e12.4)'): n fields of twelve columns from the next line; a short line is padded with blanks and a
  // blank field is zero.  (saghai_sigma0.dat starts with a fragment of a line, so each of the reference's reads of
  // that file sits one line early -- its amplitude reads pick up the kinematic line and the first amplitude line.
  // The tables are what the reference's own read statements make of the file.)
  auto next = [&](double* v, int n) -> bool {
    if (!std::fgets(line, sizeof(line), f)) return false;
    size_t len = std::strlen(line);
    while (len > 0 && (line[len - 1] == '\n' || line[len - 1] == '\r')) line[--len] = 0;
    for (int k = 0; k < n; ++k) {
      char field[13];
      for (int c = 0; c < 12; ++c) { const size_t at = (size_t)12 * k + c; field[c] = at < len ? line[at] : ' '; }
      field[12] = 0;
      char* e = nullptr;
      const double x = std::strtod(field, &e);
      if (e == field) {
        for (const char* q = field; *q; ++q) if (*q != ' ') return false;
        v[k] = 0.;
      } else {
        for (const char* q = e; *q; ++q) if (*q != ' ') return false;
        v[k] = x;
      }
    }
    return true;
  };
  bool ok = true;
  double v[6];
  for (int ir = 0; ir < n1 && ok; ++ir)
    for (int iq = 0; iq < n2 && ok; ++iq) {
      if (which == 0) ok = next_list2();
      for (int ia = 0; ia < n3 && ok; ++ia) {
        if (which == 1) ok = next_list2();
        ok = ok && next(v, 5);
        const size_t at = (size_t)ir + (size_t)n1 * ((size_t)iq + (size_t)n2 * ia);
        // zrff1, ziff1, zrff2, ziff2, zrff3, ziff3 / zrff4 ... ziff6; REAL*4 variables
        ok = ok && next(v, 6);
        if (ok) for (int k = 0; k < 3; ++k) { tbl[(size_t)k * n_tab + at] = (float)v[2 * k]; tbl[(size_t)(6 + k) * n_tab + at] = (float)v[2 * k + 1]; }
        ok = ok && next(v, 6);
        if (ok) for (int k = 0; k < 3; ++k) { tbl[(size_t)(3 + k) * n_tab + at] = (float)v[2 * k]; tbl[(size_t)(9 + k) * n_tab + at] = (float)v[2 * k + 1]; }
      }
    }
  std::fclose(f);
  if (!ok) err = path + ": malformed Saghai table";
  return ok;
}
}  // namespace

int simc_b200_read_saghai_file(const char* path, int which, float* tbl, char* msg, int msg_len) {
  if (!path || !tbl || which < 0 || which > 1) return SIMC_ERR_ARG;
  std::vector<float> t;
  std::string err;
  if (!read_saghai_file(path, which, t, err)) {
    if (msg && msg_len > 0) std::snprintf(msg, msg_len, "%s", err.c_str());
    return SIMC_ERR_IO;
  }
  std::memcpy(tbl, t.data(), sizeof(float) * t.size());
  return SIMC_OK;
}

int simc_b200_load_saghai_files(simc_handle* h, const char* dir) {
  if (!h || !dir) return SIMC_ERR_ARG;
  const char* names[2] = {"saghai_proton.dat", "saghai_sigma0.dat"};
  for (int which = 0; which < 2; ++which) {
    std::vector<float> t;
    std::string err;
    if (!read_saghai_file(std::string(dir) + "/" + names[which], which, t, err)) return fail(h, SIMC_ERR_IO, err);
    const int rc = simc_b200_set_saghai_table(h, which, t.data());
    if (rc) return rc;
  }
  return SIMC_OK;
}

// trgInit (trg_track.f:243-347): the field map of the polarised target.  bz, br: B_field_z(iz, ir), B_field_r(iz, ir)
// in the file's reading order (ir outer, iz inner), 51 x 51 nodes 2 cm apart; both null: the uniform 5 T test field
// (26 cm in z, 16 cm in r) trgInit builds for a blank file name.
int simc_b200_set_field_map(simc_handle* h, const double* bz, const double* br) {
  if (!h || ((bz == nullptr) != (br == nullptr))) return SIMC_ERR_ARG;
  const int N = 51;
  std::vector<double> img(2 * (size_t)N * N);
  for (int ir = 0; ir < N; ++ir)
    for (int iz = 0; iz < N; ++iz) {
      const size_t k = (size_t)ir * N + iz;
      if (bz) { img[k] = bz[k]; img[(size_t)N * N + k] = br[k]; }
      else {
        const double rr = 2. * (double)ir, zz = 2. * (double)iz;
        img[k] = (rr <= 16. && zz <= 26.) ? 5.0 : 0.0;
        img[(size_t)N * N + k] = 0.;
      }
    }
  CU(h, cudaSetDevice(h->device));
  if (!h->d_field) CU(h, cudaMalloc(&h->d_field, img.size() * sizeof(double)));
  CU(h, cudaMemcpy(h->d_field, img.data(), img.size() * sizeof(double), cudaMemcpyHostToDevice));
  return SIMC_OK;
}

// tgt_field_file (simc.f:154, trg_track.f:291-320): "0" = no field, blank = the uniform test field, else the map:
// 2601 list-directed rows "z r Bz Br ..." with z running fastest.
int simc_b200_load_field_file(simc_handle* h, const char* path) {
  if (!h || !path) return SIMC_ERR_ARG;
  std::string name(path);
  while (!name.empty() && name.back() == ' ') name.pop_back();
  if (name.empty()) return simc_b200_set_field_map(h, nullptr, nullptr);
  const int N = 51;
  std::vector<double> bz((size_t)N * N, 0.0), br((size_t)N * N, 0.0);
  if (name != "0") {
    FILE* f = std::fopen(name.c_str(), "r");
    if (!f) return fail(h, SIMC_ERR_IO, "cannot open target field map " + name);
    bool ok = true;
    for (size_t k = 0; k < bz.size() && ok; ++k) {
      double v[7];
      ok = std::fscanf(f, "%lf %lf %lf %lf %lf %lf %lf", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6]) == 7;
      bz[k] = v[2]; br[k] = v[3];
    }
    std::fclose(f);
    if (!ok) return fail(h, SIMC_ERR_IO, "target field map " + name + ": fewer than 51 x 51 rows of seven numbers");
  }
  return simc_b200_set_field_map(h, bz.data(), br.data());
}

// maidtbl of sigmaid (physics_pion.f:596-625): the slice sig0 reads
int simc_b200_set_maid_table(simc_handle* h, int ipi, const double* tbl) {
  if (!h || !tbl) return SIMC_ERR_ARG;
  if (ipi != 3 && ipi != 4) return fail(h, SIMC_ERR_ARG, "MAID table: ipi = 3 (pi+ n) or 4 (pi- p); neutral pions are out of scope");
  CU(h, cudaSetDevice(h->device));
  double*& d = h->d_maid[ipi - 3];
  const size_t bytes = sizeof(double) * 25 * 46 * 6 * 4;
  if (!d) CU(h, cudaMalloc(&d, bytes));
  CU(h, cudaMemcpy(d, tbl, bytes, cudaMemcpyHostToDevice));
  return SIMC_OK;
}

// maidpipn.dat / maidpimp.dat: 25 x 46 x 23 rows of '(f11.6,18f8.4)' (physics_pion.f:611-625); sigmaid only
// ever indexes the first six angle rows (ith = 1..6) and sig0 the first four columns
int simc_b200_load_maid_file(simc_handle* h, int ipi, const char* path) {
  if (!h || !path) return SIMC_ERR_ARG;
  FILE* f = std::fopen(path, "r");
  if (!f) return fail(h, SIMC_ERR_IO, std::string("cannot open MAID table ") + path);
  std::vector<double> tbl((size_t)25 * 46 * 6 * 4);
  char line[512];
  bool ok = true;
  for (int iq = 0; iq < 25 && ok; ++iq)
    for (int iw = 0; iw < 46 && ok; ++iw)
      for (int ith = 0; ith < 23 && ok; ++ith) {
        if (!std::fgets(line, sizeof line, f)) { ok = false; break; }
        if (ith >= 6) continue;
        const size_t len = std::strlen(line);
        const int start[4] = {0, 11, 19, 27}, width[4] = {11, 8, 8, 8};
        for (int j = 0; j < 4; ++j) {
          if (len < (size_t)(start[j] + width[j])) { ok = false; break; }
          char field[16];
          std::memcpy(field, line + start[j], width[j]);
          field[width[j]] = 0;
          tbl[(((size_t)iq * 46 + iw) * 6 + ith) * 4 + j] = std::atof(field);
        }
      }
  std::fclose(f);
  if (!ok) return fail(h, SIMC_ERR_IO, "MAID table: short or malformed file");
  return simc_b200_set_maid_table(h, ipi, tbl.data());
}

// theory_init (init.f:828-905) from arrays
int simc_b200_set_theory_table(simc_handle* h, int n_shells, double absorption, double e_fermi, const double* nprot,
                               const double* em, const double* emsig, const double* bs_norm, const int32_t* n_pm,
                               const double* pm_first, const double* pm_bin, const double* rho) {
  if (!h || !nprot || !em || !emsig || !bs_norm || !n_pm || !pm_first || !pm_bin || !rho) return SIMC_ERR_ARG;
  if (n_shells < 1 || n_shells > 21) return fail(h, SIMC_ERR_ARG, "theory table: 1..21 momentum distributions (simulate.inc:118)");
  size_t total = 0;
  for (int m = 0; m < n_shells; ++m) {
    if (n_pm[m] < 2 || n_pm[m] > 500) return fail(h, SIMC_ERR_ARG, "theory table: 2..500 points per distribution (simulate.inc:117)");
    if (!(pm_bin[m] > 0) || !(bs_norm[m] != 0)) return fail(h, SIMC_ERR_ARG, "theory table: bad bin width or normalisation");
    total += (size_t)n_pm[m];
  }
  const bool heavy = h->cfg.doing_heavy != 0;
  const double pi = 3.141592653589793;
  std::vector<double> img(8 * (size_t)n_shells + total);
  size_t pos = 8 * (size_t)n_shells, src = 0;
  for (int m = 0; m < n_shells; ++m) {
    double* sh = img.data() + 8 * m;
    sh[0] = nprot[m] * absorption;
    sh[1] = em[m]; sh[2] = emsig[m];
    sh[3] = heavy ? (pi / 2. + std::atan((em[m] - e_fermi) / (0.5 * emsig[m]))) / pi : 1.;      // init.f:896-901
    sh[4] = pm_first[m] - pm_bin[m] / 2.;
    sh[5] = pm_bin[m];
    sh[6] = (double)n_pm[m];
    sh[7] = (double)pos;
    for (int k = 0; k < n_pm[m]; ++k) img[pos + k] = rho[src + k] / bs_norm[m];
    pos += n_pm[m]; src += n_pm[m];
  }
  CU(h, cudaSetDevice(h->device));
  if (h->d_theory) { cudaFree(h->d_theory); h->d_theory = nullptr; }
  CU(h, cudaMalloc(&h->d_theory, img.size() * sizeof(double)));
  CU(h, cudaMemcpy(h->d_theory, img.data(), img.size() * sizeof(double), cudaMemcpyHostToDevice));
  h->theory_nrho = n_shells; h->theory_efermi = e_fermi;
  return SIMC_OK;
}

int simc_b200_load_theory_file(simc_handle* h, const char* path) {
  if (!h || !path) return SIMC_ERR_ARG;
  try {
    const TheoryFile T = read_theory_file(path);
    std::vector<int32_t> n(T.n_pm.begin(), T.n_pm.end());
    return simc_b200_set_theory_table(h, T.n_shells, T.absorption, T.e_fermi, T.nprot.data(), T.em.data(), T.emsig.data(),
                                      T.bs_norm.data(), n.data(), T.pm_first.data(), T.pm_bin.data(), T.rho.data());
  } catch (const std::exception& e) {
    return fail(h, SIMC_ERR_IO, e.what());
  }
}

// dbase.f:563-587 from arrays: cumulative probability divided by its last entry
int simc_b200_set_pfermi_table(simc_handle* h, int n, const double* pval, const double* mprob) {
  if (!h || !pval || !mprob) return SIMC_ERR_ARG;
  if (n < 2 || n > 2000) return fail(h, SIMC_ERR_ARG, "momentum distribution: 2..2000 rows (dbase.f:581)");
  std::vector<double> img(2 * (size_t)n);
  for (int i = 0; i < n; ++i) { img[i] = pval[i]; img[n + i] = mprob[i] / mprob[n - 1]; }
  for (int i = 1; i < n; ++i)
    if (!(img[n + i] >= img[n + i - 1]) || !(pval[i] > pval[i - 1]))
      return fail(h, SIMC_ERR_ARG, "momentum distribution: pval must increase and mprob must be a cumulative probability");
  CU(h, cudaSetDevice(h->device));
  if (h->d_pfm) { cudaFree(h->d_pfm); h->d_pfm = nullptr; }
  CU(h, cudaMalloc(&h->d_pfm, img.size() * sizeof(double)));
  CU(h, cudaMemcpy(h->d_pfm, img.data(), img.size() * sizeof(double), cudaMemcpyHostToDevice));
  h->pfm_n = n;
  return SIMC_OK;
}

// The reference's file format (deut.dat, he3.dat, ...): rows "p  cumulative probability", list-directed
// (Fortran `d` exponents), at most 2000 rows (dbase.f:581-584)
int simc_b200_load_pfermi_file(simc_handle* h, const char* path) {
  if (!h || !path) return SIMC_ERR_ARG;
  try {
    std::vector<double> pval, mprob;
    read_pfermi_file(path, pval, mprob);
    return simc_b200_set_pfermi_table(h, (int)pval.size(), pval.data(), mprob.data());
  } catch (const std::exception& e) {
    return fail(h, SIMC_ERR_IO, e.what());
  }
}

// ReadTbl (cteq5/Ctq5Pdf.f:239-281) from arrays
int simc_b200_set_cteq5_table(simc_handle* h, int nx, int nt, int nfmx, double lambda, double qini, double qmax,
                              double xmin, const double* xv, const double* qv, const double* upd) {
  if (!h || !xv || !qv || !upd) return SIMC_ERR_ARG;
  (void)qini; (void)qmax; (void)xmin;        // only used for warnings in the reference
  if (nx < 2 || nx > 105 || nt < 2 || nt > 25 || nfmx < 3 || nfmx > 6 || !(lambda > 0))
    return fail(h, SIMC_ERR_ARG, "CTEQ5 table: 2 <= Nx <= 105, 2 <= Nt <= 25, 3 <= NfMx <= 6 (Ctq5Pdf.f:117)");
  const size_t npts = (size_t)(nx + 1) * (nt + 1) * (nfmx + 3);
  std::vector<double> img((size_t)(nx + 1) + (nt + 1) + npts);
  for (int i = 0; i <= nx; ++i) img[i] = xv[i];
  for (int i = 0; i <= nt; ++i) img[(size_t)nx + 1 + i] = std::log(qv[i] / lambda);
  std::copy(upd, upd + npts, img.begin() + (nx + 1) + (nt + 1));
  for (int i = 1; i <= nx; ++i) if (!(xv[i] > xv[i - 1])) return fail(h, SIMC_ERR_ARG, "CTEQ5 table: x grid must increase");
  for (int i = 1; i <= nt; ++i) if (!(qv[i] > qv[i - 1])) return fail(h, SIMC_ERR_ARG, "CTEQ5 table: Q grid must increase");
  CU(h, cudaSetDevice(h->device));
  if (h->d_pdf) { cudaFree(h->d_pdf); h->d_pdf = nullptr; }
  CU(h, cudaMalloc(&h->d_pdf, img.size() * sizeof(double)));
  CU(h, cudaMemcpy(h->d_pdf, img.data(), img.size() * sizeof(double), cudaMemcpyHostToDevice));
  h->pdf_nx = nx; h->pdf_nt = nt; h->pdf_nfmx = nfmx; h->pdf_al = lambda;
  return SIMC_OK;
}

// cteq5*.tbl: comment lines alternate with list-directed numeric blocks (Ctq5Pdf.f:250-279)
int simc_b200_load_cteq5_file(simc_handle* h, const char* path) {
  if (!h || !path) return SIMC_ERR_ARG;
  FILE* f = std::fopen(path, "r");
  if (!f) return fail(h, SIMC_ERR_IO, std::string("cannot open CTEQ5 table ") + path);
  char line[1024];
  auto skip = [&]() { return std::fgets(line, sizeof line, f) != nullptr; };
  auto read_n = [&](std::vector<double>& v, size_t n) {
    v.resize(n);
    for (size_t i = 0; i < n; ++i) if (std::fscanf(f, "%lf", &v[i]) != 1) return false;
    int c;
    while ((c = std::fgetc(f)) != EOF && c != '\n') {}       // rest of the last numeric line
    return true;
  };
  std::vector<double> head, dims, q, x, upd;
  bool ok = skip() && skip() && read_n(head, 9) && skip() && read_n(dims, 3);
  int nx = 0, nt = 0, nfmx = 0;
  if (ok) {
    nx = (int)dims[0]; nt = (int)dims[1]; nfmx = (int)dims[2];
    ok = nx >= 2 && nx <= 105 && nt >= 2 && nt <= 25 && nfmx >= 3 && nfmx <= 6;
  }
  ok = ok && skip() && read_n(q, 2 + (size_t)nt + 1) && skip() && read_n(x, 1 + (size_t)nx + 1) && skip() &&
       read_n(upd, (size_t)(nx + 1) * (nt + 1) * (nfmx + 3));
  std::fclose(f);
  if (!ok) return fail(h, SIMC_ERR_IO, "CTEQ5 table: malformed file");
  return simc_b200_set_cteq5_table(h, nx, nt, nfmx, head[2], q[0], q[1], x[0], x.data() + 1, q.data() + 2, upd.data());
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

int simc_b200_set_compiled_maps(simc_handle* h, int on) {
  if (!h) return SIMC_ERR_ARG;
  h->compiled_maps = on ? 1 : 0;
  return SIMC_OK;
}

// Device-free: runs the map compiler on a set of optics tables and leaves the cubin in the cache directory, so that
// a later run with the same tables loads it without calling NVRTC.  __graft_entry__.build() does this for the
// shipped optics.  info4: stretches, source bytes, cubin bytes, 1 if the cubin was already cached.
int simc_b200_precompile_optics(int arm_id, int n_classes, const int32_t* fwd_class_start, const double* fwd_coeff,
                                const int8_t* fwd_expon, const double* fwd_length_cm, int n_rec, const double* rec_coeff,
                                const int8_t* rec_expon, int strict_mode, const char* cache_dir, const char* dump_source_path,
                                int64_t* info4, char* msg, int msg_len) {
  auto say = [&](const std::string& m) { if (msg && msg_len > 0) { std::snprintf(msg, (size_t)msg_len, "%s", m.c_str()); } };
  if (!fwd_class_start || !fwd_coeff || !fwd_expon || !rec_coeff || !rec_expon || n_classes <= 0 || n_rec <= 0) {
    say("bad argument");
    return SIMC_ERR_ARG;
  }
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
    ArmSlot s;
    s.host = compile_arm(arm_id, f, r);
    plan_stretches(s);
    const std::string src = generate_stretch_source(s.host, source_specs(s), strict_mode != 0, map_block_threads(), map_min_blocks());
    if (dump_source_path && *dump_source_path) {
      FILE* fp = std::fopen(dump_source_path, "w");
      if (fp) { std::fwrite(src.data(), 1, src.size(), fp); std::fclose(fp); }
    }
    std::string cubin, err;
    bool cached = false;
    if (!jit_compile_cubin(src, cache_dir ? std::string(cache_dir) : jit_default_cache_dir(), cubin, &cached, err)) {
      say(err);
      return SIMC_ERR_STATE;
    }
    if (info4) { info4[0] = (int64_t)s.stretches.size(); info4[1] = (int64_t)src.size(); info4[2] = (int64_t)cubin.size(); info4[3] = cached ? 1 : 0; }
    return SIMC_OK;
  } catch (const std::exception& e) {
    say(e.what());
    return SIMC_ERR_IO;
  }
}

// Device-free: the product's COSY readers (transp_init / mc_*_recon first-call semantics, optics_host.cpp) on a pair
// of files, returned as the arrays simc_b200_set_optics takes.  Capacities in terms; counts come back in n_out[3] =
// {classes, forward terms, recon terms}.
int simc_b200_read_optics_files(const char* forward_path, const char* recon_path, int32_t max_fwd_terms, int32_t max_rec_terms,
                                int32_t* fwd_class_start, double* fwd_coeff, int8_t* fwd_expon, double* fwd_length_cm,
                                int32_t* fwd_adrift, double* fwd_driftdist_cm, double* rec_coeff, int8_t* rec_expon,
                                int32_t* n_out, char* msg, int msg_len) {
  auto say = [&](const std::string& m) { if (msg && msg_len > 0) std::snprintf(msg, (size_t)msg_len, "%s", m.c_str()); };
  if (!forward_path || !recon_path || !fwd_class_start || !fwd_coeff || !fwd_expon || !rec_coeff || !rec_expon || !n_out) {
    say("NULL argument");
    return SIMC_ERR_ARG;
  }
  try {
    const ForwardMaps f = read_forward_maps(forward_path);
    const CosyTerms r = read_recon_map(recon_path);
    int32_t pos = 0;
    for (size_t k = 0; k < f.cls.size(); ++k) {
      fwd_class_start[k] = pos;
      const int n = f.cls[k].n();
      if (pos + n > max_fwd_terms) { say("forward capacity too small"); return SIMC_ERR_ARG; }
      std::memcpy(fwd_coeff + 5 * (size_t)pos, f.cls[k].coef.data(), sizeof(double) * 5 * (size_t)n);
      std::memcpy(fwd_expon + 5 * (size_t)pos, f.cls[k].expo.data(), 5 * (size_t)n);
      if (fwd_length_cm) fwd_length_cm[k] = f.length_cm[k];
      if (fwd_adrift) fwd_adrift[k] = f.adrift[k];
      if (fwd_driftdist_cm) fwd_driftdist_cm[k] = f.driftdist_cm[k];
      pos += n;
    }
    fwd_class_start[f.cls.size()] = pos;
    if (r.n() > max_rec_terms) { say("recon capacity too small"); return SIMC_ERR_ARG; }
    std::memcpy(rec_coeff, r.coef.data(), sizeof(double) * 4 * (size_t)r.n());
    std::memcpy(rec_expon, r.expo.data(), 5 * (size_t)r.n());
    n_out[0] = (int32_t)f.cls.size(); n_out[1] = pos; n_out[2] = r.n();
    return SIMC_OK;
  } catch (const std::exception& e) {
    say(e.what());
    return SIMC_ERR_IO;
  }
}

int simc_b200_optics_info(simc_handle* h, int arm_id, int64_t* info8) {
  if (!h || !info8) return SIMC_ERR_ARG;
  auto it = h->arms.find(arm_id);
  if (it == h->arms.end() || !it->second.loaded) return fail(h, SIMC_ERR_STATE, "optics not loaded for this arm");
  const CompiledArm& c = it->second.host;
  info8[0] = c.tab.n_classes; info8[1] = c.fwd_terms; info8[2] = c.fwd_nonzero; info8[3] = c.rec_terms;
  info8[4] = (int64_t)c.recs.size(); info8[5] = (int64_t)c.fwd_nonzero; info8[6] = (int64_t)c.ops.size(); info8[7] = 0;
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
  if (using_coll && arm_id != SIMC_ARM_HMS && arm_id != SIMC_ARM_SHMS)
    return fail(h, SIMC_ERR_ARG, "collimator stepping exists for the HMS and the SHMS only (mc_hms_coll / mc_shms_coll)");
  if (n == 0) return SIMC_OK;
  CU(h, cudaSetDevice(h->device));
  TransportBatchArgs a;
  a.arm = it->second.img.data(); a.n = n; a.in = d_in_soa; a.seed = seed;
  a.ms_flag = ms_flag; a.wcs_flag = wcs_flag; a.decay_flag = decay_flag; a.using_coll = using_coll;
  a.ctau = h->cfg.ctau; a.out = d_out_soa; a.flags = d_flags;
  a.n_stretch = 0; a.hut_begin = 0; a.tk = nullptr; a.lists = nullptr; a.counts = nullptr; a.sink = nullptr;
  if (h->compiled_maps && !decay_flag && !using_coll) {
    // the RNG-free stretches of the program run as generated kernels (mapgen.h), as in the event loop
    ArmSlot& slot = it->second;
    int rc = ensure_compiled(h, slot);
    if (rc) return rc;
    if (!slot.stretches.empty()) {
      rc = ensure_tb_scratch(h, n);
      if (rc) return rc;
      a.n_stretch = (int)slot.stretches.size();
      for (int k = 0; k < a.n_stretch; ++k) a.stretch_fn[k] = slot.jit.fns[k];
      a.hut_begin = slot.hut_begin;
      int sms = 148;
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
      a.stretch_block = map_block_threads(); a.stretch_grid = sms * map_min_blocks();
      a.tk = h->tb_tk; a.lists = h->tb_lists; a.counts = h->tb_counts; a.sink = h->tb_sink;
    }
  }
  cudaError_t e = h->strict ? strict::launch_transport_batch(a, h->stream) : fast::launch_transport_batch(a, h->stream);
  if (e != cudaSuccess) return cuda_fail(h, e, "k_transport_batch launch");
  h->launches += a.n_stretch > 0 ? 3 + a.n_stretch : 1;
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


// ---- the event loop ----------------------------------------------------------------------------
namespace {

int weight_qexp(const simc_run_config& cfg) {
  const double w = cfg.w_ref > 0 ? cfg.w_ref : 1.0;
  return std::ilogb(w) - 64;
}

// Which settings this build of the loop implements; everything else is refused loudly.
int validate_loop_config(simc_handle* h, bool need_optics = true) {
  const simc_run_config& c = h->cfg;
  const bool meson = ((c.doing_hydpi || c.doing_deutpi || c.doing_hepi) && c.doing_pion) ||
                     ((c.doing_hydkaon || c.doing_deutkaon || c.doing_hekaon) && c.doing_kaon);
  const bool he_meson = meson && (c.doing_hepi || c.doing_hekaon);
  const bool heavy = c.doing_heavy && c.doing_eep && !c.doing_deuterium;
  const bool deut = c.doing_deuterium && c.doing_eep && !c.doing_heavy;
  const bool semi = c.doing_semi && (c.doing_semipi || c.doing_semika) && (c.doing_hydsemi || c.doing_deutsemi) &&
                    !c.doing_pion && !c.doing_kaon;
  const bool delta = c.doing_delta && !c.doing_pion && !c.doing_kaon && !c.doing_semi && !c.doing_eep &&
                     std::lround(c.targ.A) == 1;
  // H(e,e'rho0): the reference reads no momentum distribution for D(e,e'rho) (dbase.f:563 leaves doing_deutrho out, so
  // pfer is always 0 there) and stops on A >= 3 (dbase.f:861-866): hydrogen only
  const bool rho = c.doing_rho && !c.doing_pion && !c.doing_kaon && !c.doing_semi && !c.doing_eep && !c.doing_delta &&
                   std::lround(c.targ.A) == 1;
  if (!(c.doing_hyd_elast || meson || heavy || deut || semi || delta || rho) || (c.doing_delta && !delta) ||
      (c.doing_rho && !rho) || (c.doing_semi && !semi) || c.doing_phsp)
    return fail(h, SIMC_ERR_ARG,
                "this build of the event loop implements H(e,e'p), D(e,e'p), A(e,e'p) with a Benhar or an "
                "independent-particle spectral function, H/D/A(e,e'pi+-), H/D/A(e,e'K+), H(e,e'p)pi0, H(e,e'rho0) and "
                "semi-inclusive H/D(e,e'pi+-/K+-)X");
  // using_tgt_field (trg_track.f): both arms tracked through the target's field, magnetic spectrometers only
  if (c.using_tgt_field) {
    if (!h->d_field)
      return fail(h, SIMC_ERR_STATE, "using_tgt_field: set the field map first (simc_b200_set_field_map / load_field_file)");
    if (c.electron_arm < 1 || c.electron_arm > 5 || c.hadron_arm < 1 || c.hadron_arm > 5)
      return fail(h, SIMC_ERR_ARG, "Target field reconstruction not set up for your spectrometer (simc.f:1583-1586)");
    if (c.using_HMScoll || c.using_SHMScoll)
      return fail(h, SIMC_ERR_ARG, "using_tgt_field with collimator stepping (using_HMScoll / using_SHMScoll) is not built");
    if (!c.using_E_arm_montecarlo || !c.using_P_arm_montecarlo)
      return fail(h, SIMC_ERR_ARG, "using_tgt_field needs both spectrometer Monte Carlos (spect_mode = 0)");
  }
  // calorimeter arms (calo/mc_calo.f): as the hadron arm only; with doing_pizero both decay photons are tracked
  const bool calo_p = c.hadron_arm == SIMC_ARM_CALO_RIGHT || c.hadron_arm == SIMC_ARM_CALO_LEFT;
  if (c.electron_arm == SIMC_ARM_CALO_RIGHT || c.electron_arm == SIMC_ARM_CALO_LEFT)
    return fail(h, SIMC_ERR_ARG, "a calorimeter as the electron arm is not built (hadron arm only)");
  if (c.doing_pizero && !(c.doing_pion && calo_p && (c.doing_hydpi || c.doing_deutpi) && c.which_pion <= 3))
    return fail(h, SIMC_ERR_ARG, "doing_pizero: built for exclusive H/D(e,e'pi0) with a calorimeter as the hadron arm (hadron_arm = 7, 8)");
  if (c.doing_pizero && c.pizero_ngamma != 1 && c.pizero_ngamma != 2)
    return fail(h, SIMC_ERR_ARG, "pizero_ngamma not set correctly (should be 1 or 2)");
  if (calo_p && (c.using_HMScoll || c.using_SHMScoll || c.doing_rho))
    return fail(h, SIMC_ERR_ARG, "calorimeter hadron arm: no collimator stepping, no rho decay (rho_decay.f:96-104)");
  if (rho && (c.hadron_arm < 1 || c.hadron_arm > 5))
    return fail(h, SIMC_ERR_ARG, "doing_rho: rho_decay knows the hadron arms 1..5 only (rho_decay.f:96-104)");
  if (delta && c.using_rad)
    return fail(h, SIMC_ERR_ARG,
                "doing_delta with using_rad: the reference sets no photon-energy limits for this reaction "
                "(radc.f:249-294 has no doing_delta case), so the radiated run is undefined; set using_rad = 0");
  if ((deut || (heavy && !c.use_benhar_sf)) && !h->d_theory)
    return fail(h, SIMC_ERR_STATE, "simc_b200_run: this reaction needs the theory table (simc_b200_set_theory_table / load_theory_file) first");
  if (semi && c.doing_semika && !h->d_fdss)
    return fail(h, SIMC_ERR_STATE, "simc_b200_run: semi-inclusive kaon production needs the DSS grid (simc_b200_set_fdss_table) first");
  if (semi && !h->d_pdf)
    return fail(h, SIMC_ERR_STATE, "simc_b200_run: semi-inclusive production needs the CTEQ5 table (simc_b200_set_cteq5_table) first");
  if (((semi && c.doing_deutsemi) || c.doing_deutpi || c.doing_deutkaon || he_meson) && !h->d_pfm)
    return fail(h, SIMC_ERR_STATE, "simc_b200_run: production from a nucleus needs the momentum distribution (simc_b200_set_pfermi_table) first");
  if (he_meson && (!h->d_sf || !h->d_sf_dem))
    return fail(h, SIMC_ERR_STATE, "simc_b200_run: pion/kaon production from A > 2 needs the spectral function with its Em bin widths "
                                   "(simc_b200_load_sf_file, or set_sf_table + set_sf_em_widths) first");
  if (heavy && c.use_benhar_sf && !h->d_sf)
    return fail(h, SIMC_ERR_STATE, "simc_b200_run: A(e,e'p) needs the spectral function (simc_b200_set_sf_table) first");
  // every radiative option branch is built (radc.cuh); what is left is what the reference itself stops on
  // (radc.f:211 'rad_flag is set stupidly', init.f:627 extrad_flag < 0; extrad_flag = 0 is resolved by radc_init)
  if (c.using_rad && (c.rad_flag < 0 || c.rad_flag > 3 || c.extrad_flag < 1 || c.extrad_flag > 3))
    return fail(h, SIMC_ERR_ARG, "radiative options: rad_flag must be 0..3 and extrad_flag 1..3 (after radc_init)");
  if (need_optics) for (int arm : {c.electron_arm, c.hadron_arm}) {
    const bool used = (arm == c.electron_arm) ? c.using_E_arm_montecarlo : c.using_P_arm_montecarlo;
    if (!used) continue;
    if (arm == c.hadron_arm && (arm == SIMC_ARM_CALO_RIGHT || arm == SIMC_ARM_CALO_LEFT)) continue;     // a calorimeter has no maps
    auto it = h->arms.find(arm);
    if (it == h->arms.end() || !it->second.loaded)
      return fail(h, SIMC_ERR_STATE, "simc_b200_run: load the optics of both spectrometers first");
  }
  return SIMC_OK;
}

int clear_dev_accum(simc_handle* h);

int ensure_loop_buffers(simc_handle* h, long long cap) {
  CU(h, cudaSetDevice(h->device));
  if (!h->d_cfg) {
    CU(h, cudaMalloc(&h->d_cfg, sizeof(simc_run_config)));
    CU(h, cudaMemcpy(h->d_cfg, &h->cfg, sizeof(simc_run_config), cudaMemcpyHostToDevice));
    h->qexp_w = weight_qexp(h->cfg);
    cudaDeviceProp prop;
    CU(h, cudaGetDeviceProperties(&prop, h->device));
    h->grid_blocks = prop.multiProcessorCount * 4;          // persistent stage kernels: 4 CTAs of 128 per SM
    const size_t ab = strict::dev_accum_bytes();
    CU(h, cudaMalloc(&h->d_acc, ab));
    h->acc_host.assign(ab, 0);
    CU(h, cudaMalloc(&h->d_counts, kLoopCounts * sizeof(unsigned)));
    // freshly allocated accumulators hold the identity element, whichever entry point made them
    const int rc = clear_dev_accum(h);
    if (rc) return rc;
  }
  if (cap > h->loop_cap) {
    if (h->d_state) cudaFree(h->d_state);
    if (h->d_lists) cudaFree(h->d_lists);
    h->d_state = nullptr; h->d_lists = nullptr; h->loop_cap = 0;
    CU(h, cudaMalloc(&h->d_state, sizeof(double) * (size_t)strict::n_state_fields() * (size_t)cap));
    CU(h, cudaMalloc(&h->d_lists, sizeof(unsigned) * kLoopLists * (size_t)cap));
    h->loop_cap = cap;
  }
  return SIMC_OK;
}

// device accumulators <- identity element (zeros; min/max keys of +-1e10)
int clear_dev_accum(simc_handle* h) {
  const size_t ab = strict::dev_accum_bytes();
  std::vector<unsigned char> img(ab, 0);
  // order-preserving keys of +1e10 / -1e10
  auto key = [](double d) { long long i; std::memcpy(&i, &d, 8); return i >= 0 ? i : (i ^ 0x7fffffffffffffffLL); };
  // layout knowledge lives in kernels.cu; ask it to translate an "empty" simc_accum instead
  // (contrib/slop arrays sit at the tail of DevAccum: write them by offset helper)
  const size_t off = simc_dev_accum_minmax_offset();
  long long* mm = (long long*)(img.data() + off);
  for (int i = 0; i < 32; ++i) mm[i] = key(1.0e10);          // contrib_lo
  for (int i = 0; i < 32; ++i) mm[32 + i] = key(-1.0e10);    // contrib_hi
  for (int i = 0; i < 8; ++i) mm[64 + i] = key(1.0e10);      // slop_lo
  for (int i = 0; i < 8; ++i) mm[72 + i] = key(-1.0e10);     // slop_hi
  CU(h, cudaMemcpyAsync(h->d_acc, img.data(), ab, cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return SIMC_OK;
}

// Generated kernels of one arm: source from the map compiler, cubin from the cache or NVRTC, loaded once.
int ensure_compiled(simc_handle* h, ArmSlot& s) {
  if (s.jit_ready && s.jit_strict == h->strict) return SIMC_OK;
  jit_unload(s.jit);
  s.jit_ready = false;
  s.jit_strict = h->strict;
  plan_stretches(s);
  if (s.stretches.empty()) { s.jit_ready = true; return SIMC_OK; }
  std::string src, cubin, err;
  try {
    src = generate_stretch_source(s.host, source_specs(s), h->strict != 0, map_block_threads(), map_min_blocks());
  } catch (const std::exception& e) {
    return fail(h, SIMC_ERR_STATE, std::string("map compiler: ") + e.what());
  }
  bool cached = false;
  if (!jit_compile_cubin(src, jit_default_cache_dir(), cubin, &cached, err)) return fail(h, SIMC_ERR_STATE, err);
  std::vector<std::string> names;
  for (size_t k = 0; k < source_specs(s).size(); ++k) names.push_back("seg_" + std::to_string(k));
  CU(h, cudaSetDevice(h->device));
  CU(h, cudaFree(0));                                    // the primary context is current for the driver API
  if (!jit_load(cubin, names, s.jit, err)) return fail(h, SIMC_ERR_CUDA, err);
  s.jit.from_cache = cached;
  s.jit_ready = true;
  return SIMC_OK;
}

// The chain of kernels of one spectrometer (kernels.h: ArmSchedule).
int build_schedule(simc_handle* h, int arm_id, bool use_mc, bool decay, bool coll, ArmSchedule& sc) {
  sc.n = 0;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
  const int mblock = map_block_threads(), mgrid = sms * map_min_blocks();
  auto push = [&](int kind, int b, int e, void* fn) { sc.st[sc.n++] = ArmStage{kind, b, e, fn, mblock, mgrid}; };
  auto it = h->arms.find(arm_id);
  if (use_mc && (arm_id == SIMC_ARM_CALO_RIGHT || arm_id == SIMC_ARM_CALO_LEFT)) {     // no optics: one kernel does the arm
    push(ARM_STAGE_CALO, 0, 0, nullptr);
    return SIMC_OK;
  }
  if (!use_mc || it == h->arms.end() || !it->second.loaded) {
    push(ARM_STAGE_ENTRY, 0, 0, nullptr);
    push(ARM_STAGE_LAST, 0, 0, nullptr);
    return SIMC_OK;
  }
  ArmSlot& s = it->second;
  const ArmTablesDev& tab = s.host.tab;
  const int n_ops = tab.n_ops;
  const bool compiled = h->compiled_maps && !decay && !coll;
  int pos = 0;
  if (compiled) {
    const int rc = ensure_compiled(h, s);
    if (rc) return rc;
  }
  if (compiled && !s.stretches.empty()) {
    push(ARM_STAGE_ENTRY, 0, 0, nullptr);
    for (size_t k = 0; k < s.stretches.size(); ++k) push(ARM_STAGE_COMPILED, s.stretches[k].begin, s.stretches[k].end, s.jit.fns[k]);
    pos = s.hut_begin;
  } else {
    push(ARM_STAGE_ENTRY, 0, tab.split_op, nullptr);
    pos = tab.split_op;
  }
  for (int k = 0; k < tab.n_mid; ++k) {
    if (tab.mid_op[k] <= pos || tab.mid_op[k] >= n_ops) continue;
    if (sc.n >= kArmLists - 1) break;
    push(ARM_STAGE_MIDDLE, pos, tab.mid_op[k], nullptr);
    pos = tab.mid_op[k];
  }
  // the reconstruction map as a compiled kernel of its own: hut (interpreter, up to OP_RECON), map, recon quantities
  const bool split_recon = compiled && !s.stretches.empty() && s.recon_pc > pos && !h->cfg.using_tgt_field &&
                           sc.n + 3 <= kArmLists && (int)s.jit.fns.size() > (int)s.stretches.size() && h->compiled_recon;
  if (split_recon) {
    push(ARM_STAGE_HUT, pos, s.recon_pc, nullptr);
    push(ARM_STAGE_COMPILED, s.recon_pc, s.recon_pc + 1, s.jit.fns[s.stretches.size()]);
    push(ARM_STAGE_TAIL, 0, 0, nullptr);
    return SIMC_OK;
  }
  push(ARM_STAGE_LAST, pos, n_ops, nullptr);
  return SIMC_OK;
}

// Everything a loop kernel needs besides the range of tries: buffers, tables, the stage schedules of both arms.
int prepare_launch(simc_handle* h, LoopLaunch& a, uint64_t seed, int record, bool with_arms) {
  int rc = SIMC_OK;
  a = LoopLaunch{};
  a.cfg = h->d_cfg;
  a.arm_e = h->arms.count(h->cfg.electron_arm) && h->arms[h->cfg.electron_arm].loaded ? h->arms[h->cfg.electron_arm].img.data() : nullptr;
  a.arm_p = h->arms.count(h->cfg.hadron_arm) && h->arms[h->cfg.hadron_arm].loaded ? h->arms[h->cfg.hadron_arm].img.data() : nullptr;
  a.state = h->d_state; a.cap = h->loop_cap; a.lists = h->d_lists; a.counts = h->d_counts; a.acc = h->d_acc;
  a.seed = seed; a.qexp_w = h->qexp_w; a.record_mode = record;
  a.grid_blocks = h->grid_blocks;
  auto coll = [&](int arm) { return arm == SIMC_ARM_HMS ? h->cfg.using_HMScoll : arm == SIMC_ARM_SHMS ? h->cfg.using_SHMScoll : 0; };
  a.coll_e = coll(h->cfg.electron_arm); a.coll_p = coll(h->cfg.hadron_arm);
  a.using_rad = h->cfg.using_rad;
  a.field_map = nullptr; a.field_theta_e_deg = 0.0; a.field_theta_p_deg = 0.0;
  if (h->cfg.using_tgt_field) {          // simc.f:120-156: the angle between the field axis and each arm, then trgInit
    const simc_run_config& c = h->cfg;
    const double degrad = 180. / 3.141592653589793;
    double ang[2] = {0.0, 0.0};
    for (int w = 0; w < 2; ++w) {
      const simc_spectrometer& sp = w == 0 ? c.spec_e : c.spec_p;
      if (degrad * std::fabs(c.targ_Bphi - sp.phi) < .01) {
        if (c.targ_Bangle >= sp.theta) ang[w] = -1 * std::sin(sp.phi) * (c.targ_Bangle - sp.theta);
        else ang[w] = w == 0 ? +1 * std::sin(sp.phi) * (sp.theta - c.targ_Bangle)      // as written: the two arms differ here
                             : +1 * std::sin(sp.phi) * (c.targ_Bangle - sp.theta);    // (simc.f:125 against :139)
      } else if (degrad * std::fabs(c.targ_Bphi - sp.phi) - 180.0 < .01) {
        ang[w] = +1 * std::sin(sp.phi) * (c.targ_Bangle + sp.theta);
      }     // else: the reference prints an error and goes on with the angle it had (zero)
    }
    a.field_map = h->d_field;
    a.field_theta_e_deg = ang[0] * degrad; a.field_theta_p_deg = ang[1] * degrad;
  }
  if (with_arms) {
    rc = build_schedule(h, h->cfg.hadron_arm, h->cfg.using_P_arm_montecarlo != 0, h->cfg.doing_decay != 0, a.coll_p != 0, a.sched_p);
    if (rc) return rc;
    rc = build_schedule(h, h->cfg.electron_arm, h->cfg.using_E_arm_montecarlo != 0, false, a.coll_e != 0, a.sched_e);
    if (rc) return rc;
  }
  a.sf_pm = h->d_sf; a.sf_em = h->d_sf ? h->d_sf + h->sf_npm : nullptr;
  a.sf_val = h->d_sf ? h->d_sf + h->sf_npm + h->sf_nem : nullptr;
  a.sf_npm = h->sf_npm; a.sf_nem = h->sf_nem; a.sf_dem = h->d_sf_dem;
  a.pdf_buf = h->d_pdf; a.pdf_nx = h->pdf_nx; a.pdf_nt = h->pdf_nt; a.pdf_nfmx = h->pdf_nfmx; a.pdf_al = h->pdf_al;
  a.pfm_buf = h->d_pfm; a.pfm_n = h->pfm_n;
  a.fdss_buf = h->d_fdss;
  {
    const int which = h->cfg.targ.Mrec_struck < 1150. ? 0 : 1;       // physics_kaon.f:100: Lambda below, Sigma0 above
    a.saghai_buf = h->cfg.doing_kaon ? h->d_saghai[which] : nullptr;
    a.saghai_n[0] = which ? 20 : 10; a.saghai_n[1] = which ? 10 : 11; a.saghai_n[2] = 19;
  }
  a.maid_buf = h->d_maid[(h->cfg.which_pion == 1 || h->cfg.which_pion == 11 || h->cfg.which_pion == 3) ? 1 : 0];
  a.theory_buf = h->d_theory; a.theory_nrho = h->theory_nrho; a.theory_efermi = h->theory_efermi;
  {
    const MatTable mt = make_mat_table(h->cfg.targ);          // host libm, once per call
    static_assert(sizeof(mt) == sizeof(a.mats), "MatTable layout");
    std::memcpy(a.mats, &mt, sizeof(mt));
  }
  return SIMC_OK;
}

// rows of the device buffer simc_b200_event_batch and simc_b200_ntuple_batch share
constexpr int kRecRows = SIMC_EVENT_NREC > SIMC_NTUPLE_MAXCOL ? SIMC_EVENT_NREC : SIMC_NTUPLE_MAXCOL;
int run_batches(simc_handle* h, int64_t first_try, int64_t n_tries, uint64_t seed, int record, double* d_rec,
                int* d_status) {
  int rc = validate_loop_config(h);
  if (rc) return rc;
  const long long cap = record ? n_tries : std::min<long long>(h->batch, n_tries);
  rc = ensure_loop_buffers(h, std::max<long long>(cap, 1));
  if (rc) return rc;
  LoopLaunch a;
  rc = prepare_launch(h, a, seed, record, true);
  if (rc) return rc;
  a.rec = d_rec; a.status = d_status;
  size_t ev_pos = 5 * h->ev_used.size();
  for (int64_t done = 0; done < n_tries;) {
    const int64_t nb = std::min<int64_t>(n_tries - done, cap);
    a.first_try = first_try + done;
    a.n_tries = nb;
    const int n_stage = record ? 5 : 4;
    if (h->timing) {
      while (h->ev.size() < ev_pos + 5) { cudaEvent_t e; CU(h, cudaEventCreate(&e)); h->ev.push_back(e); }
      CU(h, cudaEventRecord(h->ev[ev_pos], h->stream));
    }
    for (int st = 0; st < n_stage; ++st) {
      cudaError_t e = h->strict ? strict::launch_loop_stage(a, st, h->stream) : fast::launch_loop_stage(a, st, h->stream);
      if (e != cudaSuccess) return cuda_fail(h, e, "event-loop kernel launch");
      h->launches += h->strict ? strict::launches_of_stage(a, st) : fast::launches_of_stage(a, st);
      if (h->timing && st < 4) CU(h, cudaEventRecord(h->ev[ev_pos + 1 + st], h->stream));
    }
    if (h->timing) { ev_pos += 5; h->ev_used.push_back(1); }
    done += nb;
  }
  return SIMC_OK;
}

}  // namespace

extern "C" {

int simc_b200_accum_clear(simc_handle* h, simc_accum* acc) {
  if (!h || !acc) return SIMC_ERR_ARG;
  std::memset(acc, 0, sizeof(*acc));
  const int qw = weight_qexp(h->cfg);
  acc->wtcontribute.qexp = qw;
  acc->sum_sigcc.qexp = qw;
  for (int i = 0; i < 8; ++i) { acc->sumerr[i].qexp = -80; acc->sumerr2[i].qexp = -80; }
  for (int k = 0; k < 6; ++k) for (int b = 0; b < SIMC_NHIST; ++b) acc->hist_w[k][b].qexp = qw;
  for (int i = 0; i < 32; ++i) { acc->contrib[i].lo = 1.0e10; acc->contrib[i].hi = -1.0e10; }
  for (int i = 0; i < 8; ++i) { acc->slop[i].lo = 1.0e10; acc->slop[i].hi = -1.0e10; }
  return SIMC_OK;
}

int simc_b200_run_async(simc_handle* h, int64_t first_try, int64_t n_tries, uint64_t seed) {
  if (!h) return SIMC_ERR_ARG;
  if (n_tries < 0) return fail(h, SIMC_ERR_ARG, "simc_b200_run: n_tries < 0");
  if (n_tries == 0) return validate_loop_config(h);
  return run_batches(h, first_try, n_tries, seed, 0, nullptr, nullptr);
}

int simc_b200_fetch(simc_handle* h, simc_accum* acc) {
  if (!h || !acc) return SIMC_ERR_ARG;
  if (!h->d_acc) return SIMC_OK;                       // nothing was run
  CU(h, cudaSetDevice(h->device));
  CU(h, cudaMemcpyAsync(h->acc_host.data(), h->d_acc, h->acc_host.size(), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  if (h->strict) strict::accum_to_host(h->acc_host.data(), acc, h->qexp_w);
  else fast::accum_to_host(h->acc_host.data(), acc, h->qexp_w);
  return clear_dev_accum(h);
}

int simc_b200_run(simc_handle* h, int64_t first_try, int64_t n_tries, uint64_t seed, simc_accum* acc) {
  if (!h || !acc) return SIMC_ERR_ARG;
  int rc = simc_b200_run_async(h, first_try, n_tries, seed);
  if (rc) return rc;
  return simc_b200_fetch(h, acc);
}

int simc_b200_device_accum(simc_handle* h, void** dev_ptr, int64_t* n_int64, void** dev_minmax, int64_t* n_minmax) {
  if (!h) return SIMC_ERR_ARG;
  if (!h->d_acc) {
    int rc = validate_loop_config(h);
    if (rc) return rc;
    rc = ensure_loop_buffers(h, 1);
    if (rc) return rc;
    rc = clear_dev_accum(h);
    if (rc) return rc;
  }
  const size_t off = simc_dev_accum_minmax_offset();
  // [0, off): unsigned 64-bit sums (add across ranks as two's complement int64 pairs -- see DESIGN.md);
  // [off, off+80*8): min/max keys: lo keys take MIN, hi keys take MAX; the STOP counters follow.
  if (dev_ptr) *dev_ptr = h->d_acc;
  if (n_int64) *n_int64 = (int64_t)(strict::dev_accum_bytes() / 8);
  if (dev_minmax) *dev_minmax = (unsigned char*)h->d_acc + off;
  if (n_minmax) *n_minmax = 80;
  return SIMC_OK;
}

int simc_b200_reduce_gathered(simc_handle* h, const void* d_gathered, int n_ranks) {
  if (!h || !d_gathered || n_ranks < 1) return SIMC_ERR_ARG;
  if (!h->d_acc) return fail(h, SIMC_ERR_STATE, "simc_b200_reduce_gathered: no device accumulators (call simc_b200_device_accum first)");
  CU(h, cudaSetDevice(h->device));
  const cudaError_t e = simc_launch_reduce_gathered(d_gathered, n_ranks, h->d_acc, h->stream);
  if (e != cudaSuccess) return cuda_fail(h, e, "k_reduce_gathered launch");
  h->launches += 1;
  return SIMC_OK;
}

int simc_b200_stage_times(simc_handle* h, int enable, double* ms4, int64_t* launches4) {
  if (!h) return SIMC_ERR_ARG;
  CU(h, cudaSetDevice(h->device));
  CU(h, cudaStreamSynchronize(h->stream));
  for (size_t b = 0; b < h->ev_used.size(); ++b) {
    for (int st = 0; st < 4; ++st) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, h->ev[5 * b + st], h->ev[5 * b + st + 1]) == cudaSuccess) {
        h->stage_ms[st] += ms;
        h->stage_launches[st] += 1;
      }
    }
  }
  h->ev_used.clear();
  for (int st = 0; st < 4; ++st) {
    if (ms4) ms4[st] = h->stage_ms[st];
    if (launches4) launches4[st] = h->stage_launches[st];
    h->stage_ms[st] = 0; h->stage_launches[st] = 0;
  }
  h->timing = enable ? 1 : 0;
  return SIMC_OK;
}

int simc_b200_fp64_peak(simc_handle* h, double* tflops_fma, double* tflops_muladd) {
  if (!h) return SIMC_ERR_ARG;
  CU(h, cudaSetDevice(h->device));
  cudaDeviceProp prop;
  CU(h, cudaGetDeviceProperties(&prop, h->device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 16;
  double* scratch = nullptr;
  CU(h, cudaMalloc(&scratch, sizeof(double) * (size_t)blocks * threads));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double res[2] = {0, 0};
  for (int fma = 1; fma >= 0; --fma) {
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0, h->stream);
      cudaError_t e = strict::launch_fp64_peak(scratch, blocks, threads, iters, fma, h->stream);
      cudaEventRecord(e1, h->stream);
      cudaStreamSynchronize(h->stream);
      if (e != cudaSuccess) { cudaFree(scratch); return cuda_fail(h, e, "fp64 peak kernel"); }
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      const double flops = 2.0 * 8.0 * (double)iters * (double)blocks * threads;
      if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    res[fma ? 0 : 1] = best;
    h->launches += 4;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(scratch);
  if (tflops_fma) *tflops_fma = res[0];
  if (tflops_muladd) *tflops_muladd = res[1];
  return SIMC_OK;
}

int simc_b200_log_batch(simc_handle* h, int64_t n, const double* x, double* out_log, double* out_log10) {
  if (!h) return SIMC_ERR_ARG;
  if (n < 0 || (n > 0 && (!x || !out_log || !out_log10))) return fail(h, SIMC_ERR_ARG, "simc_b200_log_batch: bad argument");
  if (n == 0) return SIMC_OK;
  CU(h, cudaSetDevice(h->device));
  double *d_x = nullptr, *d_o = nullptr;
  CU(h, cudaMalloc(&d_x, sizeof(double) * (size_t)n));
  cudaError_t e = cudaMalloc(&d_o, sizeof(double) * 2 * (size_t)n);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_x, x, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = h->strict ? strict::launch_log_batch(n, d_x, d_o, h->stream) : fast::launch_log_batch(n, d_x, d_o, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_log, d_o, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_log10, d_o + n, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d_x); cudaFree(d_o);
  h->launches += 1;
  if (e != cudaSuccess) return cuda_fail(h, e, "simc_b200_log_batch");
  return SIMC_OK;
}

int simc_b200_set_batch(simc_handle* h, int64_t tries_per_batch) {
  if (!h || tries_per_batch < 128) return SIMC_ERR_ARG;
  h->batch = tries_per_batch;
  return SIMC_OK;
}

int simc_b200_event_batch(simc_handle* h, int64_t first_try, int64_t n, uint64_t seed, double* rec_soa,
                          int32_t* status) {
  if (!h) return SIMC_ERR_ARG;
  if (n < 0 || (n > 0 && (!rec_soa || !status))) return fail(h, SIMC_ERR_ARG, "simc_b200_event_batch: bad argument");
  if (n == 0) return validate_loop_config(h);
  CU(h, cudaSetDevice(h->device));
  if (n > h->rec_n) {
    if (h->d_rec) cudaFree(h->d_rec);
    if (h->d_status) cudaFree(h->d_status);
    h->d_rec = nullptr; h->d_status = nullptr; h->rec_n = 0;
    CU(h, cudaMalloc(&h->d_rec, sizeof(double) * kRecRows * (size_t)n));
    CU(h, cudaMalloc(&h->d_status, sizeof(int) * (size_t)n));
    h->rec_n = n;
  }
  int rc = validate_loop_config(h);
  if (rc) return rc;
  rc = ensure_loop_buffers(h, n);
  if (rc) return rc;
  // records of stages that were not reached read as zero
  CU(h, cudaMemsetAsync(h->d_state, 0, sizeof(double) * (size_t)strict::n_state_fields() * (size_t)h->loop_cap, h->stream));
  CU(h, cudaMemsetAsync(h->d_rec, 0, sizeof(double) * kRecRows * (size_t)n, h->stream));
  rc = run_batches(h, first_try, n, seed, 1, h->d_rec, h->d_status);
  if (rc) return rc;
  CU(h, cudaMemcpyAsync(rec_soa, h->d_rec, sizeof(double) * SIMC_EVENT_NREC * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaMemcpyAsync(status, h->d_status, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  // the accumulators also saw these tries: a parity dump is not part of a run, drop them
  return clear_dev_accum(h);
}

// results_ntu_write for a range of tries: the loop in record mode 2, then the contributing rows in try order
int simc_b200_ntuple_batch(simc_handle* h, int64_t first_try, int64_t n, uint64_t seed, double* rows, int32_t* n_cols,
                           int64_t* n_rows, int64_t* try_of_row) {
  if (!h || !n_cols || !n_rows) return SIMC_ERR_ARG;
  if (n < 0 || (n > 0 && !rows)) return fail(h, SIMC_ERR_ARG, "simc_b200_ntuple_batch: bad argument");
  int rc = validate_loop_config(h);
  if (rc) return rc;
  const simc_run_config& c = h->cfg;
  *n_cols = c.doing_pizero ? 65 : c.doing_rho ? 59 : c.doing_semi ? 56 : (c.doing_pion || c.doing_kaon || c.doing_delta) ? (c.doing_kaon ? 55 : 53) : 46;      // NtupleInit.f:33-343
  // eight more columns for a polarised target in the meson layouts (NtupleInit.f:142-160, 239-256); rho's third tag
  // (mmnuc) stays behind them
  if (c.using_tgt_field && *n_cols != 46) *n_cols += 8;
  *n_rows = 0;
  if (n == 0) return SIMC_OK;
  CU(h, cudaSetDevice(h->device));
  if (n > h->rec_n) {
    if (h->d_rec) cudaFree(h->d_rec);
    if (h->d_status) cudaFree(h->d_status);
    h->d_rec = nullptr; h->d_status = nullptr; h->rec_n = 0;
    CU(h, cudaMalloc(&h->d_rec, sizeof(double) * kRecRows * (size_t)n));
    CU(h, cudaMalloc(&h->d_status, sizeof(int) * (size_t)n));
    h->rec_n = n;
  }
  rc = ensure_loop_buffers(h, n);
  if (rc) return rc;
  // the accumulators of a run in progress must survive: park them
  std::vector<unsigned char> saved(h->acc_host.size());
  CU(h, cudaMemcpyAsync(saved.data(), h->d_acc, saved.size(), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaMemsetAsync(h->d_rec, 0, sizeof(double) * kRecRows * (size_t)n, h->stream));
  rc = run_batches(h, first_try, n, seed, 2, h->d_rec, h->d_status);
  if (rc) {      // put the parked accumulators back before reporting the failure
    cudaMemcpyAsync(h->d_acc, saved.data(), saved.size(), cudaMemcpyHostToDevice, h->stream);
    cudaStreamSynchronize(h->stream);
    return rc;
  }
  std::vector<double> soa((size_t)SIMC_NTUPLE_MAXCOL * (size_t)n);
  std::vector<int> status((size_t)n);
  CU(h, cudaMemcpyAsync(soa.data(), h->d_rec, sizeof(double) * soa.size(), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaMemcpyAsync(status.data(), h->d_status, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaMemcpyAsync(h->d_acc, saved.data(), saved.size(), cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  int64_t r = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (status[i] != 4) continue;
    for (int k = 0; k < *n_cols; ++k) rows[r * SIMC_NTUPLE_MAXCOL + k] = soa[(size_t)k * n + i];
    if (try_of_row) try_of_row[r] = first_try + i;
    ++r;
  }
  *n_rows = r;
  return SIMC_OK;
}

int simc_b200_radc_batch(simc_handle* h, int64_t n, const double* in_soa, double* out_soa) {
  if (!h) return SIMC_ERR_ARG;
  if (n < 0 || (n > 0 && (!in_soa || !out_soa))) return fail(h, SIMC_ERR_ARG, "simc_b200_radc_batch: bad argument");
  if (n == 0) return SIMC_OK;
  const simc_run_config& c = h->cfg;
  if (c.rad_flag < 0 || c.rad_flag > 3 || c.extrad_flag < 1 || c.extrad_flag > 3)
    return fail(h, SIMC_ERR_ARG, "simc_b200_radc_batch: rad_flag must be 0..3 and extrad_flag 1..3 (after radc_init)");
  CU(h, cudaSetDevice(h->device));
  int rc = ensure_loop_buffers(h, 1);
  if (rc) return rc;
  double *d_in = nullptr, *d_out = nullptr;
  CU(h, cudaMalloc(&d_in, sizeof(double) * SIMC_RADC_NIN * (size_t)n));
  CU(h, cudaMalloc(&d_out, sizeof(double) * SIMC_RADC_NOUT * (size_t)n));
  cudaError_t e = cudaMemcpyAsync(d_in, in_soa, sizeof(double) * SIMC_RADC_NIN * (size_t)n, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess)
    e = h->strict ? strict::launch_radc_batch(h->d_cfg, n, d_in, d_out, h->stream)
                  : fast::launch_radc_batch(h->d_cfg, n, d_in, d_out, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_soa, d_out, sizeof(double) * SIMC_RADC_NOUT * (size_t)n, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d_in); cudaFree(d_out);
  h->launches += 1;
  if (e != cudaSuccess) return cuda_fail(h, e, "simc_b200_radc_batch");
  return SIMC_OK;
}

int simc_b200_field_batch(simc_handle* h, int spect, double theta_deg, int64_t n, const double* in_soa, double* out_soa) {
  if (!h) return SIMC_ERR_ARG;
  if (n < 0 || (n > 0 && (!in_soa || !out_soa)) || (spect != 1 && spect != -1))
    return fail(h, SIMC_ERR_ARG, "simc_b200_field_batch: bad argument");
  if (!h->d_field) return fail(h, SIMC_ERR_STATE, "simc_b200_field_batch: set the field map first (simc_b200_set_field_map / load_field_file)");
  if (n == 0) return SIMC_OK;
  CU(h, cudaSetDevice(h->device));
  double *d_in = nullptr, *d_out = nullptr;
  CU(h, cudaMalloc(&d_in, sizeof(double) * SIMC_FIELD_NIN * (size_t)n));
  CU(h, cudaMalloc(&d_out, sizeof(double) * SIMC_FIELD_NOUT * (size_t)n));
  cudaError_t e = cudaMemcpyAsync(d_in, in_soa, sizeof(double) * SIMC_FIELD_NIN * (size_t)n, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess)
    e = h->strict ? strict::launch_field_batch(h->d_field, theta_deg, spect, n, d_in, d_out, h->stream)
                  : fast::launch_field_batch(h->d_field, theta_deg, spect, n, d_in, d_out, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_soa, d_out, sizeof(double) * SIMC_FIELD_NOUT * (size_t)n, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d_in); cudaFree(d_out);
  h->launches += 1;
  if (e != cudaSuccess) return cuda_fail(h, e, "simc_b200_field_batch");
  return SIMC_OK;
}

int simc_b200_semi_batch(simc_handle* h, int64_t n, const double* in_soa, double* out_soa) {
  if (!h) return SIMC_ERR_ARG;
  if (n < 0 || (n > 0 && (!in_soa || !out_soa))) return fail(h, SIMC_ERR_ARG, "simc_b200_semi_batch: bad argument");
  if (!h->d_pdf) return fail(h, SIMC_ERR_STATE, "simc_b200_semi_batch: set the CTEQ5 table first");
  if (h->cfg.doing_semika && !h->d_fdss) return fail(h, SIMC_ERR_STATE, "simc_b200_semi_batch: set the DSS grid first");
  if (n == 0) return SIMC_OK;
  CU(h, cudaSetDevice(h->device));
  int rc = ensure_loop_buffers(h, 1);
  if (rc) return rc;
  LoopLaunch a{};
  a.pdf_buf = h->d_pdf; a.pdf_nx = h->pdf_nx; a.pdf_nt = h->pdf_nt; a.pdf_nfmx = h->pdf_nfmx; a.pdf_al = h->pdf_al;
  a.fdss_buf = h->d_fdss;
  double *d_in = nullptr, *d_out = nullptr;
  CU(h, cudaMalloc(&d_in, sizeof(double) * SIMC_SEMI_NIN * (size_t)n));
  CU(h, cudaMalloc(&d_out, sizeof(double) * SIMC_SEMI_NOUT * (size_t)n));
  cudaError_t e = cudaMemcpyAsync(d_in, in_soa, sizeof(double) * SIMC_SEMI_NIN * (size_t)n, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess)
    e = h->strict ? strict::launch_semi_batch(h->d_cfg, a, n, d_in, d_out, h->stream)
                  : fast::launch_semi_batch(h->d_cfg, a, n, d_in, d_out, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_soa, d_out, sizeof(double) * SIMC_SEMI_NOUT * (size_t)n, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d_in); cudaFree(d_out);
  h->launches += 1;
  if (e != cudaSuccess) return cuda_fail(h, e, "simc_b200_semi_batch");
  return SIMC_OK;
}

int simc_b200_weight_batch(simc_handle* h, int64_t n, const double* in_soa, double* out_soa) {
  if (!h) return SIMC_ERR_ARG;
  if (n < 0 || (n > 0 && (!in_soa || !out_soa))) return fail(h, SIMC_ERR_ARG, "simc_b200_weight_batch: bad argument");
  int rc = validate_loop_config(h, false);
  if (rc) return rc;
  if (n == 0) return SIMC_OK;
  CU(h, cudaSetDevice(h->device));
  rc = ensure_loop_buffers(h, n);
  if (rc) return rc;
  // the accumulators of a run in progress must survive: park them
  std::vector<unsigned char> saved(h->acc_host.size());
  CU(h, cudaMemcpyAsync(saved.data(), h->d_acc, saved.size(), cudaMemcpyDeviceToHost, h->stream));
  double *d_in = nullptr, *d_out = nullptr;
  CU(h, cudaMalloc(&d_in, sizeof(double) * SIMC_WEIGHT_NIN * (size_t)n));
  CU(h, cudaMalloc(&d_out, sizeof(double) * SIMC_WEIGHT_NOUT * (size_t)n));
  LoopLaunch a;
  rc = prepare_launch(h, a, 0, 1, false);
  cudaError_t e = cudaSuccess;
  if (!rc) {
    a.rec = d_in; a.wb_out = d_out; a.n_tries = n; a.first_try = 0;
    // what the end of the loop body does not write for this reaction reads as zero
    e = cudaMemsetAsync(h->d_state, 0, sizeof(double) * (size_t)strict::n_state_fields() * (size_t)h->loop_cap, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_in, in_soa, sizeof(double) * SIMC_WEIGHT_NIN * (size_t)n, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = h->strict ? strict::launch_loop_stage(a, 5, h->stream) : fast::launch_loop_stage(a, 5, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_soa, d_out, sizeof(double) * SIMC_WEIGHT_NOUT * (size_t)n, cudaMemcpyDeviceToHost, h->stream);
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(h->d_acc, saved.data(), saved.size(), cudaMemcpyHostToDevice, h->stream);
  const cudaError_t e2 = cudaStreamSynchronize(h->stream);
  cudaFree(d_in); cudaFree(d_out);
  h->launches += 3;
  if (rc) return rc;
  if (e != cudaSuccess || e2 != cudaSuccess) return cuda_fail(h, e != cudaSuccess ? e : e2, "simc_b200_weight_batch");
  return SIMC_OK;
}

static const char* kEventFields[SIMC_EVENT_NREC] = {
    "stage", "pass_cuts", "n_draws", "stop_p", "stop_e", "weight", "sigcc", "gen_weight", "jacobian", "sigcc_recon",
    "vertex.Ein", "vertex.e.E", "vertex.e.delta", "vertex.e.yptar", "vertex.e.xptar", "vertex.p.E", "vertex.p.delta",
    "vertex.p.yptar", "vertex.p.xptar", "vertex.Q2", "orig.e.E", "orig.p.E", "Egamma_used1", "Egamma_used2",
    "Egamma_used3", "ntail", "target.x", "target.y", "target.z", "Eloss1", "Eloss2", "Eloss3", "SP.e.delta",
    "SP.e.yptar", "SP.e.xptar", "SP.p.delta", "SP.p.yptar", "SP.p.xptar", "recon.e.delta", "recon.e.yptar",
    "recon.e.xptar", "recon.p.delta", "recon.p.yptar", "recon.p.xptar", "recon.Em", "recon.Pm", "recon.W",
    "hardcorfac", "main.thetacm", "main.phicm", "ntup.sigcm", "main.davejac", "survivalprob", "ntup.mm", "main.wcm",
    "main.t", "orig.p.yptar", "orig.p.xptar", "ntup.rhomass", "ntup.rhotheta"};       // the last four: rho production only
const char* simc_b200_event_field_name(int k) { return (k >= 0 && k < SIMC_EVENT_NREC) ? kEventFields[k] : ""; }

}  // extern "C"
