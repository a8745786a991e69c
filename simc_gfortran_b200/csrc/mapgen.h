// Map compiler: turns the RNG-free stretches of an arm program (arm_program.h) -- magnet apertures, drifts and
// the COSY forward maps between them -- into straight-line CUDA source, one kernel per stretch.  A COSY class
// becomes its terms written out: the monomial of each term with its unit factors dropped (x**0 = 1.0 multiplies
// exactly), then one multiply + add per NON-ZERO coefficient in file order (a zero coefficient adds an exact
// +-0.0 to a sum that is never -0.0), i.e. the arithmetic of shared/transp.f:205-214 without the operations that
// cannot change a bit.  The compiler shares equal left-to-right prefixes between terms.  No tables, no shared
// memory: ~5 FP64 instructions per term instead of ~28 instructions and 11 shared-memory wavefronts for the
// record interpreter (transport.cuh: eval_poly), which stays as the general path (decay in flight, collimator
// stepping) and as the parity twin.
#pragma once
#include <string>
#include <vector>
#include "optics_host.h"

namespace simc {

// ops a compiled stretch may contain (no random numbers, no cross-lane work) when decay is off
bool op_is_static(int op);

struct StretchSpec { int begin, end; };      // ops [begin, end)

// One translation unit with kernels "seg_0" ... "seg_{n-1}".  Kernel ABI (all device pointers):
//   seg_k(double* tk, long long fs, long long ss, const unsigned* in_list, const unsigned* in_count, unsigned* out_list,
//         unsigned* out_count, unsigned long long* stop_acc, unsigned long long* calls_acc, double* stop_field)
// tk = field F_TK_XS of slot 0 in the state buffer (fields: xs, ys, dxdzs, dydzs, dpps, p, m2, pathlen, ...; loop.cuh);
// field k of slot s is tk[k * fs + s * ss] (records: fs = 1, ss = record length; rows: fs = row length, ss = 1);
// survivors are appended to out_list; a stopped track leaves its path length in tk, its stop code in
// stop_field[slot * ss] (if not null) and in stop_acc[2 + code]; calls_acc[class - 1] counts the map evaluations.
// block_threads / min_blocks: CTA size and minimum resident CTAs per SM the kernels are built for (launch bounds).
std::string generate_stretch_source(const CompiledArm& arm, const std::vector<StretchSpec>& segs, bool strict, int block_threads,
                                    int min_blocks);

}  // namespace simc
