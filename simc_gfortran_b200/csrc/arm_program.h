// Arm programs: the single-arm Monte Carlos of the reference (hms/mc_hms.f + mc_hms_hut.f,
// shms/mc_shms.f + mc_shms_hut.f, ...) expressed as a flat list of warp-uniform operations
// that one CUDA thread per event interprets.  Every geometric constant is evaluated on the
// host in IEEE double exactly as the Fortran parameter expression it comes from, so the
// device compares against bit-identical limits.  POD only: shared by host and device code.
#pragma once
#include <stdint.h>

namespace simc {

enum ArmOpCode : int32_t {
  OP_END = 0,
  OP_PROJECT,        // a = z_drift                                   shared/project.f
  OP_PROJECT_DD,     // z_drift = driftdist(class i0) + a             (mc_hms.f:258,284,310,337,411)
  OP_TRANSP,         // i0 = class (1-based), a = zd                  shared/transp.f
  OP_CUT_R2,         // stop if xs^2+ys^2 > a          (a = r*r)      e.g. mc_hms.f:261
  OP_CUT_ABS_Y,      // stop if |ys-a| > b                            mc_hms.f:222
  OP_CUT_ABS_X,      // stop if |xs-a| > b                            mc_hms.f:226
  OP_CUT_OCT,        // stop if |xs-a| > c*|ys-b| + d                 mc_hms.f:230
  OP_CUT_OFF_R2,     // stop if (xs-a)^2+(ys-b)^2 > c                 mc_hms.f:373
  OP_ROT_H,          // (xt,yt) = rotate_haxis(xs,ys); a=tan,b=sin,c=cos; then xt += d
  OP_ROT_V,          // (xt,yt) = rotate_vaxis(xs,ys); a=tan,b=sin,c=cos; then yt += d
  OP_CUT_T_R2,       // stop if xt^2+yt^2 > a                         shms/mc_shms.f:866
  OP_CUT_HB,         // stop if xt^2 > a or yt > b or yt < c          shms/mc_shms.f:425
  OP_CUT_HMS_DIPOLE, // stop if hit_dipole(xt,yt)                     mc_hms.f:445-492
  OP_CUT_HMS_PIPE,   // stop if (xt-a)^2+(yt-b)^2 > c or |yt-b| > d   mc_hms.f:361
  OP_MARK_HUT,       // STOP_hut counter
  OP_RESMULT_DRAW,   // one uniform; resmult = 2 if u < a else 1      mc_hms_hut.f:292-298
  OP_RESMULT_ONE,    // resmult = 1                                   shms/mc_shms_hut.f:80
  OP_MUSC,           // a = radw, b = sqrt(radw)                      shared/musc.f
  OP_MUSC_EXT,       // a = radw, b = sqrt(radw), c = x_len           shared/musc_ext.f
  OP_DC_PLANE,       // i0 = plane (0..11), i1 = 1 for a y plane, a = sigma   mc_hms_hut.f:351-364
  OP_CUT_BOX,        // stop if xs > a or xs < b or ys > c or ys < d  mc_hms_hut.f:374
  OP_SCIN_COUNT,     // count++ if ys < a and ys > b and xs < c and xs > d   mc_hms_hut.f:492
  OP_SCIN_TRIG,      // stop if count < i0                            mc_hms_hut.f:570
  OP_LFIT,           // fit focal-plane track through REAL*4 arrays   mc_hms_hut.f:438-455
  OP_CUT_FP_CAL,     // xcal=x_fp+dx_fp*a ...; stop if ycal>b or ycal<c or xcal>d or xcal<e  mc_shms_hut.f:415
  OP_RECON,          // mc_*_recon + overwrite dpp,y,dxdz,dydz        mc_hms.f:419-437
                     //   i0 = 1: clamp every |hut(i)| <= 1e-30 (sos/mc_sos_recon.f:79-81);
                     //   a = shift taken off the returned y_fp afterwards (hrsl/mc_hrsl.f:525)
  OP_CUT_R,          // stop if sqrt(xs^2+ys^2) > a                   hrsl/mc_hrsl.f:162
  OP_CUT_T_ABSX,     // stop if |xt-a| > b                            hrsl/mc_hrsl.f:370
  OP_CUT_T_TRAP,     // stop if |yt| + a*xt > b                       hrsl/mc_hrsl.f:377
  OP_CUT_T_RECT,     // stop if xt > a or xt < b or yt > c or yt < d  hrsl/mc_hrsl_hut.f:286
  OP_CUT_T_BOX,      // stop if yt > a or -yt > b or -xt > c or -xt < d   sos/mc_sos.f:234
  OP_CUT_SOS_EXIT,   // w = a + b*(xs+c); stop if |xs| > c or |ys| > w    sos/mc_sos.f:328
  OP_SHIFT,          // xs += a*dxdzs, ys += a*dydzs (no path length) sos/mc_sos.f:339
  OP_COLL,           // mc_hms_coll / mc_shms_coll: pions and muons step through the collimator when using_coll.
                     //   a..d = h_entr, v_entr, h_exit, v_exit, e = y_off, i0 = ops to skip when taken (the plain
                     //   aperture checks), i1 = stop code of SLIT_HOR (VERT, OCT follow); the next two ops are data
  OP_COLL_DATA       // [1]: a = step size, b = radiation length, c..e = rho, CO, |CO/27| of the collimator material;
                     // [2]: a..e = ln10, log(me/I^2), and the three Z/A products (target.cuh: MatConst)
};

struct ArmOp {
  int32_t op;
  int32_t code;      // stop code recorded when this op rejects the event
  int32_t i0, i1;
  double a, b, c, d, e;
};

constexpr int kMaxArmOps = 320;
constexpr int kArmBlockThreads = 128;  // threads per CTA of every kernel that evaluates COSY maps (= kBlock)
constexpr int kMaxClasses = 41;       // max_class, spectrometers.inc:6

// One COSY map compiled into term records.  The reference evaluates every term as
//   term = x^e1 * theta^e2 * y^e3 * phi^e4 * delta^e5   (left to right, unit factors included)
//   sum(1:5) = sum(1:5) + term * coeff(1:5,i)           (all five outputs, zero coefficients included)
// (shared/transp.f:205-214).  A record keeps exactly that shape so that the device loop has no
// data-dependent branch: one 8-byte word with four 16-bit byte offsets into the thread's shared power
// table (x^a*theta^b | y^e | phi^e | delta^e), then the term's coefficients, zeros included
// (forward maps 5, reconstruction maps 4 + one pad word): 6 words = 48 bytes.
// Terms whose coefficients are all zero add exact zeros in the reference and are dropped.  Records
// come in chunks of kRecChunk; the last chunk of a map is filled with records that multiply 1.0 by
// zero coefficients (they add +0.0).  A warp stages one chunk at a time in shared memory.
constexpr int kPolyEntries = 28 + 3 * 7;      // (a,b) with a+b <= 6, then 7 powers of each slow variable
constexpr int kRecWords = 6;
constexpr int kRecChunk = 8;                  // records per staged chunk: 8 * 48 bytes = 24 lanes * 16 bytes
// index of x^a*theta^b in the power table (a + b <= 6)
#if defined(__CUDACC__)
__host__ __device__
#endif
inline constexpr int poly_xt_index(int a, int b) { return a * 7 - (a * (a - 1)) / 2 + b; }

struct PolyClass {
  int32_t rec_begin;                // into recs[], in 8-byte words (chunks are 16-byte aligned)
  int32_t n_chunks;                 // chunks of kRecChunk records (terms with a non-zero coefficient + fill)
  int32_t n_terms;                  // terms in the file
  int32_t adrift;
  double length_cm;                 // !LENGTH: comment (0 if none)
  double driftdist_cm;              // extracted drift length if a pure drift
};

struct ArmTablesDev {
  const double* recs;               // term records, forward classes then recon
  const double* pad_ptr;
  PolyClass fwd[kMaxClasses];       // 1-based class k -> fwd[k-1]
  PolyClass rec;
  int32_t n_classes;
  int32_t n_ops;
  int32_t split_op;     // ops [0,split_op): entrance apertures (cheap, most rejections); [split_op,n_ops): the rest
  int32_t n_mid;        // further compaction points behind split_op (0..3), increasing op indices:
  int32_t mid_op[3];    // survivors are compacted again before ops [mid_op[k], ...)
  int32_t pad;
};

}  // namespace simc
