// Host side of the optics: COSY file reader, group compiler, arm-program builders.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "arm_program.h"

namespace simc {

struct CosyTerms {                       // one map, file order, TOF terms removed
  std::vector<double> coef;              // [n][nout]
  std::vector<int8_t> expo;              // [n][5]
  int nout = 5;
  int n() const { return nout ? (int)(coef.size() / nout) : 0; }
};

struct ForwardMaps {
  std::vector<CosyTerms> cls;            // class k -> cls[k-1]
  std::vector<double> length_cm;         // from !LENGTH: comments
  std::vector<int> adrift;
  std::vector<double> driftdist_cm;
};

// Reads a forward file with the semantics of transp_init (shared/transp.f:294-474).
// Throws std::runtime_error with the reference's `stop` text on malformed input.
ForwardMaps read_forward_maps(const std::string& path);
// Reads a reconstruction file like mc_hms_recon's first call (hms/mc_hms_recon.f:70-102).
CosyTerms read_recon_map(const std::string& path);
// Fills adrift/driftdist for maps supplied as arrays (same test as transp.f:399-438).
void classify_drifts(ForwardMaps& f);

struct CompiledArm {
  std::vector<double> recs;              // term records (arm_program.h), 8-byte words
  ArmTablesDev tab;                      // recs pointer left null (the device address is set by the caller)
  std::vector<ArmOp> ops;
  long long fwd_terms = 0, fwd_nonzero = 0, rec_terms = 0;
  ForwardMaps fwd;                       // the maps as read (file order), for the map compiler (mapgen.h)
  CosyTerms rec;
  int arm_id = 0;
};

// Compiles the maps into groups and appends the arm's op list.  using_coll selects the
// pion collimator stepping branch (mc_hms.f:209) when the particle mass calls for it.
CompiledArm compile_arm(int arm_id, const ForwardMaps& fwd, const CosyTerms& rec);

const char* stop_name(int arm_id, int code);
int n_stop_codes(int arm_id);

}  // namespace simc
