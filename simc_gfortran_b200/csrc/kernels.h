// Host-callable launchers of the two kernel variants (strict = reference association,
// no FMA; fast = explicit FMA + re-associated monomials in the COSY polynomials only).  Implemented in kernels.cu, which is
// compiled twice.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace simc {

// Survivor lists of the event loop (loop.cuh: LoopArgs::lists): 0 = generated; 1..8 = hadron-arm stages (8 = the arm's
// survivors); 9..16 = electron-arm stages (16 = survivors of both arms); 17 = tries on their way to k_regen.
constexpr int kArmLists = 8;
constexpr int kLoopLists = 2 * kArmLists + 2;
constexpr int kRegenListIdx = 2 * kArmLists + 1;
constexpr int kLoopCounts = 32;                 // [0] slots handed out, [1 + l] length of list l

// How one spectrometer's program is cut into kernels.  ENTRY: k_arm<W,0> (target multiple scattering, SP quantities,
// TRANSPORT coordinates, then ops [begin,end)); COMPILED: a generated straight-line kernel (mapgen.h) for an RNG-free
// stretch; MIDDLE: k_arm<W,2>, the interpreter on a stretch; LAST: k_arm<W,1>, the interpreter up to the end of the
// program, reconstruction and the arm's recon quantities.  Survivors are compacted between stages.
enum ArmStageKind : int { ARM_STAGE_ENTRY = 0, ARM_STAGE_COMPILED = 1, ARM_STAGE_MIDDLE = 2, ARM_STAGE_LAST = 3,
                          ARM_STAGE_CALO = 4,       // the whole arm in one kernel: a calorimeter as the hadron arm (k_calo)
                          // the last stage in three: HUT = interpreter up to (not including) OP_RECON, the map as a
                          // compiled kernel (ARM_STAGE_COMPILED), TAIL = the arm's recon quantities
                          ARM_STAGE_HUT = 5, ARM_STAGE_TAIL = 6 };
struct ArmStage { int kind; int begin, end; void* fn; int block, grid; };   // block / grid: launch shape of a compiled stage
struct ArmSchedule { int n; ArmStage st[kArmLists]; };

struct TransportBatchArgs {
  const void* arm;          // ArmDev image on the HOST: passed to the kernel by value (constant bank)
  long long n;
  const double* in;         // [9][n] device
  unsigned long long seed;
  int ms_flag, wcs_flag, decay_flag, using_coll;
  double ctau;
  double* out;              // [12][n] device
  int* flags;               // [n] device
  // compiled path (mapgen.h): the RNG-free stretches [0, hut_begin) run as generated kernels on a scratch track
  // buffer, the interpreter takes over at hut_begin for the survivors.  n_stretch = 0: interpreter only.
  int n_stretch;
  void* stretch_fn[kArmLists];
  int stretch_block, stretch_grid;   // launch shape the stretches were generated for
  int hut_begin;
  double* tk;               // [12][n] device scratch: 11 track rows (loop.cuh F_TK_*) + the stop code
  unsigned* lists;          // [n_stretch + 1][n] device scratch
  unsigned* counts;         // [n_stretch + 1] device scratch
  unsigned long long* sink; // [SIMC_NSTOP + 48] device scratch for the stop / call counters the kernels keep
};

// Opaque to the host code: built and consumed inside kernels.cu (loop.cuh: LoopArgs).
struct LoopLaunch {
  const void* cfg;            // simc_run_config* (device)
  const void* arm_e;          // ArmDev images on the HOST (null if the arm's Monte Carlo is off): each arm
  const void* arm_p;          // kernel gets its program and map directory by value, in the constant bank
  double* state;              // [n_state_fields][cap]
  long long cap;
  unsigned* lists;            // [kLoopLists][cap] (loop.cuh: LoopArgs)
  unsigned* counts;           // [16]
  void* acc;                  // DevAccum*
  long long first_try, n_tries;
  unsigned long long seed;
  int qexp_w;
  int record_mode;
  double* wb_out;             // simc_b200_weight_batch (stage 5): [SIMC_WEIGHT_NOUT][n_tries] device; its input rows are `rec`
  double* rec;                // [SIMC_EVENT_NREC][n_tries] device, record mode only
  int* status;
  int grid_blocks;            // persistent grid for the stage kernels
  int coll_e, coll_p;         // the arm steps pions through its collimator (using_HMScoll / using_SHMScoll)
  int using_rad;              // radiative corrections on: second generation pass (k_regen) and k_radw are launched
  // using_tgt_field: the map (device, field.cuh: FieldDev) and trgInit's angles (degrees) between the field axis and
  // the electron / hadron spectrometer (simc.f:120-156); field_map null = no field tracking
  const double* field_map;
  double field_theta_e_deg, field_theta_p_deg;
  ArmSchedule sched_e, sched_p;
  double mats[45];            // MatTable (target.cuh): 5 materials x 9 energy-loss constants, made on the host
  const double* sf_pm;        // Benhar spectral function (device): Pm axis, Em axis, values [n_pm][n_em]
  const double* sf_em;
  const double* sf_val;
  int sf_npm, sf_nem;
  const double* sf_dem;       // widths of the Em bins (generate_em); null unless set
  const double* pdf_buf;      // CTEQ5 table (device): [xv(nx+1) | ql(nt+1) | upd]; null unless set
  int pdf_nx, pdf_nt, pdf_nfmx;
  double pdf_al;
  const double* fdss_buf;     // fDSS tables (device), physics_semi.cuh: FdssDev; null unless set
  const double* maid_buf;     // MAID-2007 slice [25][46][6][4] of this run's charge state (device); null unless set
  const float* saghai_buf;    // Saghai tables of this run's hyperon (device), physics_meson.cuh: SaghaiDev; null unless set
  int saghai_n[3];
  const double* theory_buf;   // independent-particle spectral function (device), physics_heavy.cuh: TheoryDev
  int theory_nrho;
  double theory_efermi;
  const double* pfm_buf;      // momentum distribution (device): [pval(n) | mprob(n)]; null unless set
  int pfm_n;
};

namespace strict {
cudaError_t launch_radc_batch(const void* cfg, long long n, const double* in, double* out, cudaStream_t s);
cudaError_t launch_semi_batch(const void* cfg, const LoopLaunch& tables, long long n, const double* in, double* out, cudaStream_t s);
cudaError_t launch_loop_stage(const LoopLaunch& a, int stage, cudaStream_t s);   // 0 gen, 1 P arm, 2 E arm, 3 finish, 4 records
cudaError_t launch_fp64_peak(double* scratch, int blocks, int threads, int iters, int fma, cudaStream_t s);
cudaError_t launch_log_batch(long long n, const double* x, double* out, cudaStream_t s);
cudaError_t launch_field_batch(const double* map, double theta_deg, int spect, long long n, const double* in, double* out,
                               cudaStream_t s);
size_t dev_accum_bytes();
int n_state_fields();
void accum_to_host(const void* dev_accum_host_copy, void* simc_accum_out, int qexp_w);
int launches_of_stage(const LoopLaunch& a, int stage);
}
namespace fast {
cudaError_t launch_radc_batch(const void* cfg, long long n, const double* in, double* out, cudaStream_t s);
cudaError_t launch_semi_batch(const void* cfg, const LoopLaunch& tables, long long n, const double* in, double* out, cudaStream_t s);
cudaError_t launch_loop_stage(const LoopLaunch& a, int stage, cudaStream_t s);   // 0 gen, 1 P arm, 2 E arm, 3 finish, 4 records
cudaError_t launch_fp64_peak(double* scratch, int blocks, int threads, int iters, int fma, cudaStream_t s);
cudaError_t launch_log_batch(long long n, const double* x, double* out, cudaStream_t s);
cudaError_t launch_field_batch(const double* map, double theta_deg, int spect, long long n, const double* in, double* out,
                               cudaStream_t s);
size_t dev_accum_bytes();
int n_state_fields();
void accum_to_host(const void* dev_accum_host_copy, void* simc_accum_out, int qexp_w);
int launches_of_stage(const LoopLaunch& a, int stage);
}

}  // namespace simc
// variant-independent (compiled once, in the strict translation unit)
cudaError_t simc_launch_reduce_gathered(const void* gathered, int n_ranks, void* dev_accum, cudaStream_t s);
namespace simc {

namespace strict { cudaError_t launch_transport_batch(const TransportBatchArgs& a, cudaStream_t s); size_t arm_dev_bytes(); }
namespace fast   { cudaError_t launch_transport_batch(const TransportBatchArgs& a, cudaStream_t s); size_t arm_dev_bytes(); }

}  // namespace simc
