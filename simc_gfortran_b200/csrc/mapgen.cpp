// Map compiler (mapgen.h): CUDA source for the RNG-free stretches of an arm program.
// Every expression below restates, operation for operation, what transport.cuh: run_arm / transp / project do
// for the same op in the strict variant; constants are printed as hexadecimal floating literals (exact).
#include "mapgen.h"
#include "../../include/simc_b200.h"
#include <cstdio>
#include <stdexcept>

namespace simc {

bool op_is_static(int op) {
  switch (op) {
    case OP_PROJECT: case OP_TRANSP: case OP_CUT_R2: case OP_CUT_ABS_Y: case OP_CUT_ABS_X: case OP_CUT_OCT:
    case OP_CUT_OFF_R2: case OP_ROT_H: case OP_ROT_V: case OP_CUT_T_R2: case OP_CUT_HB: case OP_CUT_HMS_DIPOLE:
    case OP_CUT_HMS_PIPE: case OP_CUT_BOX: case OP_CUT_R: case OP_CUT_T_ABSX: case OP_CUT_T_TRAP: case OP_CUT_T_RECT:
    case OP_CUT_T_BOX: case OP_CUT_SOS_EXIT: case OP_SHIFT:
    case OP_COLL: case OP_COLL_DATA:       // the stepping only exists in the interpreter; compiled stretches run without it
      return true;
    default:
      return false;
  }
}

namespace {

std::string lit(double v) {
  char b[64];
  std::snprintf(b, sizeof(b), "%a", v);
  return std::string("(") + b + ")";
}

struct Out {
  std::string s;
  void line(const std::string& l) { s += l; s += '\n'; }
};

const char* kVar[5] = {"x", "t", "y", "p", "d"};

std::string power_name(int var, int e) { return e == 1 ? std::string(kVar[var]) : std::string(kVar[var]) + std::to_string(e); }

// one COSY map written out.  Inputs: doubles named x,t,y,p,d in scope; outputs s0..s{nout-1}.
void emit_map(Out& o, const CosyTerms& m, bool strict) {
  int maxe[5] = {0, 0, 0, 0, 0};
  for (int i = 0; i < m.n(); ++i) {
    bool nz = false;
    for (int q = 0; q < m.nout; ++q) nz = nz || m.coef[(size_t)m.nout * i + q] != 0.0;
    if (!nz) continue;
    for (int j = 0; j < 5; ++j) {
      const int e = m.expo[5 * (size_t)i + j];
      if (e < 0 || e > 6) throw std::runtime_error("COSY exponent outside 0..6");
      if (e > maxe[j]) maxe[j] = e;
    }
  }
  // powers with libgcc __powidf2's association (the reference's real**integer): v^3 = v*v^2, v^4 = (v^2)^2,
  // v^5 = v*v^4, v^6 = v^2*v^4
  for (int j = 0; j < 5; ++j) {
    const std::string v = kVar[j];
    if (maxe[j] >= 2) o.line("      const double " + v + "2 = " + v + " * " + v + ";");
    if (maxe[j] >= 3) o.line("      const double " + v + "3 = " + v + " * " + v + "2;");
    if (maxe[j] >= 4) o.line("      const double " + v + "4 = " + v + "2 * " + v + "2;");
    if (maxe[j] >= 5) o.line("      const double " + v + "5 = " + v + " * " + v + "4;");
    if (maxe[j] >= 6) o.line("      const double " + v + "6 = " + v + "2 * " + v + "4;");
  }
  bool started[5] = {false, false, false, false, false};
  for (int i = 0; i < m.n(); ++i) {
    const double* c = &m.coef[(size_t)m.nout * i];
    bool nz = false;
    for (int q = 0; q < m.nout; ++q) nz = nz || c[q] != 0.0;
    if (!nz) continue;
    std::string mono;
    for (int j = 0; j < 5; ++j) {
      const int e = m.expo[5 * (size_t)i + j];
      if (e == 0) continue;
      mono = mono.empty() ? power_name(j, e) : "(" + mono + " * " + power_name(j, e) + ")";
    }
    if (mono.empty()) mono = "1.0";
    o.line("      { const double m = " + mono + ";");
    for (int q = 0; q < m.nout; ++q) {
      if (c[q] == 0.0) continue;
      const std::string s = "s" + std::to_string(q);
      if (strict) o.line("        " + s + " = " + s + " + m * " + lit(c[q]) + ";");
      else o.line("        " + s + " = fma(m, " + lit(c[q]) + ", " + s + ");");
      started[q] = true;
    }
    o.line("      }");
  }
  (void)started;
}

void emit_op(Out& o, const CompiledArm& arm, const ArmOp& op, bool strict) {
  const std::string a = lit(op.a), b = lit(op.b), c = lit(op.c), d = lit(op.d);
  const std::string stop = " { stop_code = " + std::to_string(op.code) + "; alive = false; }";
  switch (op.op) {
    case OP_PROJECT:       // shared/project.f, no decay
      o.line("      path = path + " + a + " * sqrt(1 + dx * dx + dy * dy);");
      o.line("      xs = xs + dx * " + a + ";");
      o.line("      ys = ys + dy * " + a + ";");
      break;
    case OP_TRANSP: {      // shared/transp.f:134-279, no decay
      const int klass = op.i0;
      const CosyTerms& m = arm.fwd.cls.at(klass - 1);
      o.line("      {   // transp class " + std::to_string(klass) + ": " + std::to_string(m.n()) + " terms");
      o.line("      count_call(s_calls, " + std::to_string(klass - 1) + ");");
      o.line("      const double x = xs, t = dx * 1000., y = ys, p = dy * 1000., d = dpp;");
      o.line("      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0;");
      emit_map(o, m, strict);
      o.line("      xs = s0;");
      o.line("      dx = s1 / 1000.;");
      o.line("      ys = s2;");
      o.line("      dy = s3 / 1000.;");
      o.line("      const double delta_z = -s4;");
      o.line("      path = path + (" + a + " + delta_z);");
      o.line("      }");
      break;
    }
    case OP_CUT_R2: o.line("      if ((xs * xs + ys * ys) > " + a + ")" + stop); break;
    case OP_CUT_ABS_Y: o.line("      if (fabs(ys - " + a + ") > " + b + ")" + stop); break;
    case OP_CUT_ABS_X: o.line("      if (fabs(xs - " + a + ") > " + b + ")" + stop); break;
    case OP_CUT_OCT: o.line("      if (fabs(xs - " + a + ") > (" + c + " * fabs(ys - " + b + ") + " + d + "))" + stop); break;
    case OP_CUT_OFF_R2:
      o.line("      if (((xs - " + a + ") * (xs - " + a + ") + (ys - " + b + ") * (ys - " + b + ")) > " + c + ")" + stop);
      break;
    case OP_ROT_H:         // rotate_haxis.f:46-62
      o.line("      { const double alpha = dx, beta = dy;");
      o.line("        const double alpha_p = (alpha + " + a + ") / (1. - alpha * " + a + ");");
      o.line("        const double beta_p = beta / (" + c + " - alpha * " + b + ");");
      o.line("        const double xi = xs;");
      o.line("        xt = xi * (" + c + " + alpha_p * " + b + ");");
      o.line("        yt = ys + xi * beta_p * " + b + ";");
      o.line("        xt = xt + " + d + "; (void)alpha_p; }");
      break;
    case OP_ROT_V:         // rotate_vaxis.f:42-58
      o.line("      { const double alpha = dy, beta = dx;");
      o.line("        const double alpha_p = (alpha + " + a + ") / (1. - alpha * " + a + ");");
      o.line("        const double beta_p = beta / (" + c + " - alpha * " + b + ");");
      o.line("        const double yi = ys;");
      o.line("        yt = yi * (" + c + " + alpha_p * " + b + ");");
      o.line("        xt = xs + yi * beta_p * " + b + ";");
      o.line("        yt = yt + " + d + "; }");
      break;
    case OP_CUT_T_R2: o.line("      if ((xt * xt + yt * yt) > " + a + ")" + stop); break;
    case OP_CUT_HB: o.line("      if ((xt * xt > " + a + ") || (yt > " + b + ") || (yt < " + c + "))" + stop); break;
    case OP_CUT_HMS_DIPOLE: o.line("      if (hms_hit_dipole(xt, yt))" + stop); break;
    case OP_CUT_HMS_PIPE:
      o.line("      if ((((xt - " + a + ") * (xt - " + a + ") + (yt - " + b + ") * (yt - " + b + ")) > " + c + ") || (fabs(yt - " + b + ") > " + d + "))" + stop);
      break;
    case OP_CUT_BOX: o.line("      if ((xs > " + a + ") || (xs < " + b + ") || (ys > " + c + ") || (ys < " + d + "))" + stop); break;
    case OP_CUT_R: o.line("      if (sqrt(xs * xs + ys * ys) > " + a + ")" + stop); break;
    case OP_CUT_T_ABSX: o.line("      if (fabs(xt - " + a + ") > " + b + ")" + stop); break;
    case OP_CUT_T_TRAP: o.line("      if ((fabs(yt) + " + a + " * xt) > " + b + ")" + stop); break;
    case OP_CUT_T_RECT: o.line("      if ((xt > " + a + ") || (xt < " + b + ") || (yt > " + c + ") || (yt < " + d + "))" + stop); break;
    case OP_CUT_T_BOX: o.line("      if ((yt > " + a + ") || (-yt > " + b + ") || (-xt > " + c + ") || (-xt < " + d + "))" + stop); break;
    case OP_CUT_SOS_EXIT:
      o.line("      { const double tmpwidth = " + a + " + " + b + " * (xs + " + c + ");");
      o.line("        if ((fabs(xs) > " + c + ") || (fabs(ys) > tmpwidth))" + stop + " }");
      break;
    case OP_SHIFT:
      o.line("      xs = xs + " + a + " * dx;");
      o.line("      ys = ys + " + a + " * dy;");
      break;
    case OP_COLL: case OP_COLL_DATA:       // using_HMScoll / using_SHMScoll off: the plain aperture checks follow
      break;
    case OP_RECON: {       // mc_hms.f:419-437 + mc_hms_recon.f:104-137 on the fitted focal-plane track, which the hut
                           // kernel left in the track rows (xs, dx, ys, dy = x_fp, dx_fp, y_fp, dy_fp); out: the
                           // reconstructed target quantities in the same rows (xs = y_tgt, dx = dph, dy = dth, dpp = delta)
      const CosyTerms& m = arm.rec;
      o.line("      {   // reconstruction map: " + std::to_string(m.n()) + " terms");
      o.line("      count_call(s_calls, 47);");
      o.line("      double h0 = xs / 100., h1 = dx, h2 = ys / 100., h3 = dy, h4 = fry / 100.;");
      o.line("      if (fabs(h4) <= 1.e-30) h4 = 1.e-30;");
      if (op.i0) {         // sos/mc_sos_recon.f:79-81: every variable is kept off zero
        o.line("      if (fabs(h0) <= 1.e-30) h0 = 1.e-30;");
        o.line("      if (fabs(h1) <= 1.e-30) h1 = 1.e-30;");
        o.line("      if (fabs(h2) <= 1.e-30) h2 = 1.e-30;");
        o.line("      if (fabs(h3) <= 1.e-30) h3 = 1.e-30;");
      }
      o.line("      const double x = h0, t = h1, y = h2, p = h3, d = h4;");
      o.line("      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;");
      emit_map(o, m, strict);
      o.line("      dx = s0;");
      o.line("      xs = s1 * 100.;");
      o.line("      dy = s2;");
      o.line("      dpp = s3 * 100.;");
      o.line("      }");
      break;
    }
    default:
      throw std::runtime_error("map compiler: op " + std::to_string(op.op) + " is not static");
  }
}

const char* kPreamble = R"SRC(
// generated by simc_b200 (mapgen.cpp): straight-line arm-program stretches
#define NSTOP @NSTOP@
__device__ __forceinline__ bool hms_hit_dipole(double x, double y) {   // hms/mc_hms.f:445-492
  const double xl = fabs(x), yl = fabs(y);
  const bool c1 = (xl <= 34.29) && (yl <= 12.07);
  const bool c2 = (xl <= 27.94) && (yl <= 18.42);
  const bool c3 = (xl <= 13.97) && (yl <= 18.95);
  const bool c4 = (xl <= 1.956) && (yl <= 20.32);
  const bool c5 = ((xl - 27.94) * (xl - 27.94) + (yl - 12.065) * (yl - 12.065)) <= 6.35 * 6.35;
  const bool c6 = (xl >= 1.956) && (xl <= 13.97) && ((yl - (-0.114) * xl - 20.54) <= 0.0);
  return !(c1 || c2 || c3 || c4 || c5 || c6);
}
__device__ __forceinline__ void count_call(unsigned* s_calls, int k) {
  const unsigned m = __activemask();
  if ((threadIdx.x & 31u) == (unsigned)(__ffs(m) - 1)) atomicAdd(&s_calls[k], (unsigned)__popc(m));
}
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
)SRC";

}  // namespace

std::string generate_stretch_source(const CompiledArm& arm, const std::vector<StretchSpec>& segs, bool strict, int block_threads,
                                    int min_blocks) {
  Out o;
  std::string pre = kPreamble;
  const std::string key = "@NSTOP@";
  pre.replace(pre.find(key), key.size(), std::to_string(SIMC_NSTOP));
  pre += "#define BLOCK " + std::to_string(block_threads) + "\n";
  o.s += pre;
  for (size_t k = 0; k < segs.size(); ++k) {
    o.line("extern \"C\" __global__ void __launch_bounds__(BLOCK, " + std::to_string(min_blocks) + ") seg_" + std::to_string(k) +
           "(double* __restrict__ tk, long long fs, long long ss, const unsigned* __restrict__ in_list, const unsigned* __restrict__ in_count,");
    o.line("    unsigned* __restrict__ out_list, unsigned* __restrict__ out_count, unsigned long long* __restrict__ stop_acc,");
    o.line("    unsigned long long* __restrict__ calls_acc, double* __restrict__ stop_field) {");
    o.line("  __shared__ unsigned s_stop[NSTOP];");
    o.line("  __shared__ unsigned s_calls[48];");
    o.line("  for (int i = threadIdx.x; i < NSTOP; i += BLOCK) s_stop[i] = 0u;");
    o.line("  for (int i = threadIdx.x; i < 48; i += BLOCK) s_calls[i] = 0u;");
    o.line("  __syncthreads();");
    o.line("  const unsigned n_in = *in_count;");
    o.line("  for (long long i0 = (long long)blockIdx.x * BLOCK; i0 < n_in; i0 += (long long)gridDim.x * BLOCK) {");
    o.line("    const long long i = i0 + threadIdx.x;");
    o.line("    const bool active = i < n_in;");
    o.line("    bool alive = active;");
    o.line("    unsigned slot = 0u;");
    o.line("    double xs = 0., ys = 0., dx = 0., dy = 0., dpp = 0., path = 0., xt = 0., yt = 0., fry = 0.;");
    o.line("    int stop_code = -1;");
    o.line("    if (active) {");
    o.line("      slot = in_list[i];");
    o.line("      double* const t = tk + (long long)slot * ss;");
    o.line("      xs = t[0 * fs]; ys = t[1 * fs]; dx = t[2 * fs]; dy = t[3 * fs];");
    o.line("      dpp = t[4 * fs]; path = t[7 * fs];");
    if (segs[k].end - segs[k].begin == 1 && arm.ops.at(segs[k].begin).op == OP_RECON) o.line("      fry = t[10 * fs];");
    o.line("    }");
    for (int pc = segs[k].begin; pc < segs[k].end; ++pc) {
      const ArmOp& op = arm.ops.at(pc);
      if (op.op == OP_COLL || op.op == OP_COLL_DATA) continue;
      // the warps of the CTA enter every map together: its code is tens of KB of straight line, and warps that
      // drift apart in it stall on instruction fetch (each streams the code through the cache on its own)
      if (op.op == OP_TRANSP || op.op == OP_RECON) o.line("    __syncthreads();");
      o.line("    if (alive) {");
      emit_op(o, arm, op, strict);
      o.line("    }");
    }
    o.line("    (void)xt; (void)yt; (void)fry;");
    o.line("    const bool ok = alive;");
    o.line("    if (active) {");
    o.line("      double* const t = tk + (long long)slot * ss;");
    o.line("      t[7 * fs] = path;");
    o.line("      if (stop_code >= 0) {");
    o.line("        if (stop_field) stop_field[(long long)slot * ss] = (double)stop_code;");
    o.line("        if (2 + stop_code < NSTOP) atomicAdd(&s_stop[2 + stop_code], 1u);");
    o.line("      } else {");
    o.line("        t[0 * fs] = xs; t[1 * fs] = ys; t[2 * fs] = dx; t[3 * fs] = dy;");
    o.line("        t[4 * fs] = dpp;");
    o.line("      }");
    o.line("    }");
    o.line("    __syncwarp();");
    o.line("    const unsigned pos = warp_append(out_count, ok);");
    o.line("    if (ok) out_list[pos] = slot;");
    o.line("  }");
    o.line("  __syncthreads();");
    o.line("  for (int i = threadIdx.x; i < NSTOP; i += BLOCK) if (s_stop[i]) atomicAdd(&stop_acc[i], (unsigned long long)s_stop[i]);");
    o.line("  for (int i = threadIdx.x; i < 48; i += BLOCK) if (s_calls[i]) atomicAdd(&calls_acc[i], (unsigned long long)s_calls[i]);");
    o.line("}");
  }
  return o.s;
}

}  // namespace simc
