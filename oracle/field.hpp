// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into libsimc_b200.so.  PARITY UNPINNED (KAT: the circle of a
// charged particle in the uniform test field, tests/test_oracle_field.py).
// trg_track.f: tracking of charged particles through the field of the polarised target (G. Warren 1999, after
// M. Muehlbauer's replay-engine code): field map and interpolation, Runge-Kutta steps, tracking to a plane,
// track_from_tgt, track_to_tgt.
#pragma once
#include <functional>
#include <string>

namespace simc_oracle {

// COMMON /trgFieldStrength/ and the two /trgFieldAngles_x/ blocks after trgInit (trg_track.f:243-347)
struct TrgField {
  static const int nz = 51, nr = 51;
  double B_field_z[nr][nz];       // Fortran B_field_z(iz, ir): [ir][iz] here
  double B_field_r[nr][nz];
  double zz[nz], rr[nr];
  double B_stheta[2], B_ctheta[2];        // [0]: electron spectrometer (spect = -1), [1]: hadron (spect = +1)
  bool set = false;
  // map: 51*51 rows of the file in reading order (ir outer, iz inner): Bz, Br per row.  null = uniform 5 T test field.
  void init(const double* bz, const double* br, double theta_e_deg, double theta_p_deg);
};

const TrgField* field_map();               // what oracle_set_field_map stored (capi.cpp); never null, `set` says if filled
// simc.f:120-156: the angles (degrees) between the field axis and the electron / hadron arm that trgInit is given
void field_arm_angles(double targ_Bangle, double targ_Bphi, double theta_e, double phi_e, double theta_p, double phi_p,
                      double& ang_e_deg, double& ang_p_deg);

void trgField(const TrgField& F, const double x_[3], double B_[3], int spect);                 // trg_track.f:350-447
void trgRK4(const TrgField& F, double factor, const double u0[9], double u1[9], double h, int spect);      // :492-533
bool trgTrackToPlane(const TrgField& F, double u[9], double E, double dl, double a, double b, double c, double d, bool ok,
                     int spect);                                                            // :154-237
// trg_track.f:591-672; returns ok of the second trgTrackToPlane
bool track_from_tgt(const TrgField& F, double& x, double& y, double& z, double& dx, double& dy, double mom, double mass,
                    int spect);
// trg_track.f:738-877.  recon(delta, dy, dx, y, fry) is the arm's mc_*_recon on the focal-plane track it holds.
bool track_to_tgt(const TrgField& F, double& delta, double& y, double& dx, double& dy, double frx, double fry, double mom,
                  double mass, double ctheta, double stheta, int spect, bool ok,
                  const std::function<void(double&, double&, double&, double&, double)>& recon);

}  // namespace simc_oracle
