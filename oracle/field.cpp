// ORACLE -- TEST INFRASTRUCTURE ONLY (see field.hpp).  trg_track.f restated.
#include "field.hpp"

#include <cmath>

namespace simc_oracle {

static inline double sign1(double b) { return std::signbit(b) ? -1. : 1.; }      // Fortran SIGN(1., b)

// trg_track.f:243-347
void TrgField::init(const double* bz, const double* br, double theta_e_deg, double theta_p_deg) {
  const double pi180 = 3.141592653 / 180.;
  B_stheta[0] = std::sin(theta_e_deg * pi180); B_ctheta[0] = std::cos(theta_e_deg * pi180);
  B_stheta[1] = std::sin(theta_p_deg * pi180); B_ctheta[1] = std::cos(theta_p_deg * pi180);
  for (int ir = 1; ir <= nr; ++ir) {
    rr[ir - 1] = 2. * (double)(ir - 1);
    zz[ir - 1] = 2. * (double)(ir - 1);
  }
  for (int ir = 1; ir <= nr; ++ir)
    for (int iz = 1; iz <= nz; ++iz) {
      if (bz) {                       // READ (1,*) xx, xx, B_field_z(iz,ir), B_field_r(iz,ir), xx, xx, xx
        B_field_z[ir - 1][iz - 1] = bz[(ir - 1) * nz + (iz - 1)];
        B_field_r[ir - 1][iz - 1] = br[(ir - 1) * nz + (iz - 1)];
      } else {                        // uniform 5 T over 26 cm in z and 16 cm in r
        B_field_r[ir - 1][iz - 1] = 0.;
        B_field_z[ir - 1][iz - 1] = (rr[ir - 1] <= 16. && zz[iz - 1] <= 26.) ? 5.0 : 0.0;
      }
    }
  set = true;
}

// simc.f:120-156 (angles in radians in, degrees out; the two arms differ in their second branch, as written)
void field_arm_angles(double targ_Bangle, double targ_Bphi, double theta_e, double phi_e, double theta_p, double phi_p,
                      double& ang_e_deg, double& ang_p_deg) {
  const double pi = 3.141592653589793, degrad = 180. / pi;
  double ang_targ_earm = 0.0, ang_targ_parm = 0.0;
  if (degrad * std::fabs(targ_Bphi - phi_e) < .01) {
    if (targ_Bangle >= theta_e) ang_targ_earm = -1 * std::sin(phi_e) * (targ_Bangle - theta_e);
    else ang_targ_earm = +1 * std::sin(phi_e) * (theta_e - targ_Bangle);
  } else if (degrad * std::fabs(targ_Bphi - phi_e) - 180.0 < .01) {
    ang_targ_earm = +1 * std::sin(phi_e) * (targ_Bangle + theta_e);
  }
  if (degrad * std::fabs(targ_Bphi - phi_p) < .01) {
    if (targ_Bangle >= theta_p) ang_targ_parm = -1 * std::sin(phi_p) * (targ_Bangle - theta_p);
    else ang_targ_parm = +1 * std::sin(phi_p) * (targ_Bangle - theta_p);
  } else if (degrad * std::fabs(targ_Bphi - phi_p) - 180.0 < .01) {
    ang_targ_parm = +1 * std::sin(phi_p) * (targ_Bangle + theta_p);
  }
  ang_e_deg = ang_targ_earm * degrad;
  ang_p_deg = ang_targ_parm * degrad;
}

// trg_track.f:350-447
void trgField(const TrgField& F, const double x_[3], double B_[3], int spect) {
  const int k = spect == -1 ? 0 : 1;
  const double B_stht = F.B_stheta[k], B_ctht = F.B_ctheta[k];
  double x[3], B[3];
  x[0] = x_[0];
  x[1] = B_stht * x_[2] + B_ctht * x_[1];
  x[2] = B_ctht * x_[2] - B_stht * x_[1];
  const double z = std::fabs(x[2]);
  const double r = std::sqrt(x[0] * x[0] + x[1] * x[1]);
  const int nz = TrgField::nz, nr = TrgField::nr;
  const int i = (int)((z - F.zz[0]) / (F.zz[1] - F.zz[0])) + 1;
  const int j = (int)((r - F.rr[0]) / (F.rr[1] - F.rr[0])) + 1;
  if ((i + 1 > nz) || (i < 1) || (j + 1 > nr) || (j < 1)) {
    B_[0] = 0.; B_[1] = 0.; B_[2] = 0.;
    return;
  }
  auto Bz = [&](int ii, int jj) { return F.B_field_z[jj - 1][ii - 1]; };
  auto Br = [&](int ii, int jj) { return F.B_field_r[jj - 1][ii - 1]; };
  const double az = ((z - F.zz[i - 1]) / (F.zz[1] - F.zz[0]));
  const double ar = ((r - F.rr[j - 1]) / (F.rr[1] - F.rr[0]));
  double a0 = az * (Bz(i + 1, j) - Bz(i, j)) + Bz(i, j);
  double a1 = az * (Bz(i + 1, j + 1) - Bz(i, j + 1)) + Bz(i, j + 1);
  B[2] = (ar * (a1 - a0) + a0);
  if (r > 0.) {
    a0 = az * (Br(i + 1, j) - Br(i, j)) + Br(i, j);
    a1 = az * (Br(i + 1, j + 1) - Br(i, j + 1)) + Br(i, j + 1);
    B[1] = (ar * (a1 - a0) + a0) / r;
    if (x[2] < 0.) B[1] = -B[1];
    B[0] = B[1] * x[0];
    B[1] = B[1] * x[1];
    B_[0] = B[0];
    B_[1] = -B_stht * B[2] + B_ctht * B[1];
    B_[2] = B_ctht * B[2] + B_stht * B[1];
  } else {
    B_[0] = 0.;
    B_[1] = -B_stht * B[2];
    B_[2] = B_ctht * B[2];
  }
}

// trg_track.f:452-490
static void trgDeriv(const TrgField& F, double factor, const double u[9], double dudt[9], int spect) {
  double B[3];
  trgField(F, u, B, spect);
  dudt[0] = u[3]; dudt[1] = u[4]; dudt[2] = u[5];
  dudt[6] = u[4] * B[2] - u[5] * B[1];
  dudt[7] = u[5] * B[0] - u[3] * B[2];
  dudt[8] = u[3] * B[1] - u[4] * B[0];
  dudt[3] = dudt[6] * factor;
  dudt[4] = dudt[7] * factor;
  dudt[5] = dudt[8] * factor;
}

// trg_track.f:492-533 (components 1..6)
void trgRK4(const TrgField& F, double factor, const double u0[9], double u1[9], double h, int spect) {
  double ut[9] = {0}, dudt[9], dut[9], dum[9];
  const double hh = h * 0.5, h6 = h / 6.;
  trgDeriv(F, factor, u0, dudt, spect);
  for (int i = 0; i < 6; ++i) ut[i] = u0[i] + hh * dudt[i];
  trgDeriv(F, factor, ut, dut, spect);
  for (int i = 0; i < 6; ++i) ut[i] = u0[i] + hh * dut[i];
  trgDeriv(F, factor, ut, dum, spect);
  for (int i = 0; i < 6; ++i) {
    ut[i] = u0[i] + h * dum[i];
    dum[i] = dut[i] + dum[i];
  }
  trgDeriv(F, factor, ut, dut, spect);
  for (int i = 0; i < 6; ++i) u1[i] = u0[i] + h6 * (dudt[i] + dut[i] + 2. * dum[i]);
}

// trg_track.f:154-237
bool trgTrackToPlane(const TrgField& F, double u[9], double E, double dl, double a, double b, double c, double d, bool ok,
                     int spect) {
  if (!ok) return ok;
  const double n = 1 / std::sqrt(a * a + b * b + c * c);
  const double an = a * n, bn = b * n, cn = c * n, dn = d * n;
  const double factor = 90. / E;
  double ts = -dl / std::sqrt(u[3] * u[3] + u[4] * u[4] + u[5] * u[5]);
  double dist0 = u[0] * an + u[1] * bn + u[2] * cn + dn;
  const double maxdist = std::max(std::fabs(dist0) * 4., 1.0);
  double u0[9] = {0}, u1[9] = {0};
  trgRK4(F, factor, u, u1, ts, spect);
  double dist1 = u1[0] * an + u1[1] * bn + u1[2] * cn + dn;
  if ((sign1(dist0) == sign1(dist1)) && (std::fabs(dist0) < std::fabs(dist1))) ts = -ts;
  int steps = 0;
  const int max_steps = (int)(std::max(dist0, 10. * dl) / dl) * 10;
  if (sign1(dist0) == sign1(dist1)) {
    dist1 = dist0;
    while ((sign1(dist0) == sign1(dist1)) && ok) {
      trgRK4(F, factor, u1, u0, ts, spect);
      dist0 = u0[0] * an + u0[1] * bn + u0[2] * cn + dn;
      if (sign1(dist0) == sign1(dist1)) {
        trgRK4(F, factor, u0, u1, ts, spect);
        dist1 = u1[0] * an + u1[1] * bn + u1[2] * cn + dn;
      }
      ok = (std::fabs(dist1) < maxdist) && steps < max_steps;
      steps = steps + 1;
    }
  } else {
    for (int i = 0; i < 6; ++i) u0[i] = u[i];
  }
  if (ok)
    for (int i = 0; i < 6; ++i) u[i] = u0[i] + (u1[i] - u0[i]) * dist0 / (dist0 - dist1);
  return ok;
}

// trg_track.f:591-672
bool track_from_tgt(const TrgField& F, double& x, double& y, double& z, double& dx, double& dy, double mom, double mass,
                    int spect) {
  const double cc = 29.9792458;
  const double vel = std::fabs(mom) / std::sqrt(mom * mom + mass * mass) * cc;
  const double eng = sign1(mom) * std::sqrt(mom * mom + mass * mass);
  double vT[9] = {0};
  vT[0] = x; vT[1] = y; vT[2] = z;
  vT[5] = vel / std::sqrt(1 + dx * dx + dy * dy);
  vT[3] = dx * vT[5];
  vT[4] = dy * vT[5];
  bool ok = true;
  trgTrackToPlane(F, vT, eng, 1., 0., 0., 1., 0., ok, spect);         // "for debugging, run track first to z=0"
  ok = true;
  ok = trgTrackToPlane(F, vT, eng, 1., 0., 0., 1., -100., ok, spect);
  x = vT[0]; y = vT[1]; z = vT[2];
  dx = vT[3] / vT[5];
  dy = vT[4] / vT[5];
  return ok;
}

// trg_track.f:738-877
bool track_to_tgt(const TrgField& F, double& delta, double& y, double& dx, double& dy, double frx, double fry, double mom,
                  double mass, double ctheta, double stheta, int spect, bool ok,
                  const std::function<void(double&, double&, double&, double&, double)>& recon) {
  const double cc = 29.9792458;
  double vT[9] = {0}, vTx[9] = {0};
  double xx = -fry;
  double vel = std::fabs(mom) / std::sqrt(mom * mom + mass * mass) * cc;
  double eng = sign1(mom) * std::sqrt(mom * mom + mass * mass);
  const double mom_0 = mom / (1.e0 + delta / 100.e0);
  vT[0] = -fry + 100. * dx;
  vT[1] = y + 100. * dy;
  vT[2] = 100.;
  vT[5] = vel / std::sqrt(1 + dy * dy + dx * dx);
  vT[3] = dx * vT[5];
  vT[4] = dy * vT[5];
  ok = trgTrackToPlane(F, vT, eng, 1., 0., -ctheta, stheta, frx, ok, spect);
  int n = 0;
  double delx = 1.;
  while ((delx > .0001) && (n < 10) && ok) {
    delx = std::fabs(-fry - vT[0]);
    vTx[0] = -fry;
    for (int i = 1; i < 6; ++i) vTx[i] = vT[i];
    ok = trgTrackToPlane(F, vT, eng, 1., 0., 0., 1., 0., ok, spect);
    ok = trgTrackToPlane(F, vTx, eng, 1., 0., 0., 1., 0., ok, spect);
    xx = xx + std::min(1., std::max(-1., (vTx[0] - vT[0])));
    const double xxd = xx;
    recon(delta, dy, dx, y, xxd);
    mom = mom_0 * (1.e0 + delta / 100.e0);
    vel = std::fabs(mom) / std::sqrt(mom * mom + mass * mass) * cc;
    eng = sign1(mom) * std::sqrt(mom * mom + mass * mass);
    vT[0] = xx + 100. * dx;
    vT[1] = y + 100. * dy;
    vT[2] = 100.;
    vT[5] = vel / std::sqrt(1 + dy * dy + dx * dx);
    vT[3] = dx * vT[5];
    vT[4] = dy * vT[5];
    ok = trgTrackToPlane(F, vT, eng, 1., 0., -ctheta, stheta, frx, ok, spect);
    n = n + 1;
  }
  if (delx > .2) ok = false;
  dy = vT[4] / vT[5];
  dx = vT[3] / vT[5];
  y = vT[1];
  return ok;
}

}  // namespace simc_oracle
