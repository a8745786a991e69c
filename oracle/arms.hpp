// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into libsimc_b200.so.
// PARITY UNPINNED.  Single-arm Monte Carlos restated from hms/mc_hms*.f,
// shms/mc_shms*.f, sos/mc_sos*.f, hrsl|hrsr/mc_hrs?*.f.
#pragma once
#include "track.hpp"

namespace simc_oracle {

struct ArmOptics {
  CosyForward fwd;
  CosyRecon rec;
};

// In/out record of a single-arm call: the argument list of mc_hms (hms/mc_hms.f:1-4).
struct ArmCall {
  double p_spec = 0, th_spec = 0;
  double dpp = 0, x = 0, y = 0, z = 0, dxdz = 0, dydz = 0;   // in; dpp,y,dxdz,dydz overwritten with recon on success
  double x_fp = 0, dx_fp = 0, y_fp = 0, dy_fp = 0;           // out
  double m2 = 0;                                             // in/out (decay)
  bool ms_flag = true, wcs_flag = true, decay_flag = false;
  double resmult = 0;                                        // out
  double fry = 0;                                            // in
  bool ok_spec = false;                                      // out
  double pathlen = 0;                                        // in/out
  bool using_coll = false;                                   // using_HMScoll / using_SHMScoll
  int stop_code = 0;                                         // 0 = ok, else where it stopped (our enumeration)
  bool reached_hut = false;
  // mc_hms_coll / mc_shms_coll bump the slit STOP counters once per 0.3 cm step spent in the material, whether
  // or not the pion survives (hms/mc_hms_coll.f:95-115): hor, vert, oct
  int coll_steps[3] = {0, 0, 0};
};

// stop codes, shared with include/simc_b200.h (simc_b200_stop_name)
namespace hms_stop {
enum { OK = 0, SLIT_HOR, SLIT_VERT, SLIT_OCT, Q1_IN, Q1_MID, Q1_OUT, Q2_IN, Q2_MID, Q2_OUT, Q3_IN, Q3_MID, Q3_OUT,
       D1_IN, D1_OUT, DC1, DC2, SCIN, CAL, COLL };
}
namespace shms_stop {
enum { OK = 0, HB_IN, HB_MEN, HB_MEX, HB_OUT, SLIT_HOR, SLIT_VERT, SLIT_OCT, Q1_IN, Q1_MEN, Q1_MID, Q1_MEX, Q1_OUT,
       Q2_IN, Q2_MEN, Q2_MID, Q2_MEX, Q2_OUT, Q3_IN, Q3_MEN, Q3_MID, Q3_MEX, Q3_OUT, D1_IN, D1_FLR, D1_MEN,
       D1_MID1, D1_MID2, D1_MID3, D1_MID4, D1_MID5, D1_MID6, D1_MID7, D1_MEX, D1_OUT, DC1, DC2, S1X, S1Y, S2X, S2Y,
       CAL, CAL_FID, COLL };
}

namespace sos_stop {
enum { OK = 0, SLIT_HOR, SLIT_VERT, SLIT_OCT, QUAD_IN, QUAD_MID, QUAD_OUT, BM01_IN, BM01_OUT, BM02_IN, BM02_OUT, EXIT,
       DC1, DC2, SCIN };
}
namespace hrs_stop {
enum { OK = 0, SLIT_HOR, SLIT_VERT, Q1_IN, Q1_MID, Q1_OUT, Q2_IN, Q2_MID, Q2_OUT, D1_IN, D1_OUT, Q3_IN, Q3_MID, Q3_OUT,
       DC1, DC2, S1, S2 };
}

void mc_hms(Track& t, const ArmOptics& o, ArmCall& a);
void mc_sos(Track& t, const ArmOptics& o, ArmCall& a);
// right = true: mc_hrsr (spectrometer 3), false: mc_hrsl (spectrometer 4)
void mc_hrs(Track& t, const ArmOptics& o, ArmCall& a, bool right);
void mc_shms(Track& t, const ArmOptics& o, ArmCall& a);

}  // namespace simc_oracle
