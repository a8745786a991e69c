// Tracking of charged particles through the field of the polarised target (trg_track.f; using_tgt_field):
// bilinear interpolation of the (z, r) field map in the frame of the field axis, fourth-order Runge-Kutta steps of
// 1 cm, tracking to a plane with the reference's stepping and interpolation rules (trgTrackToPlane), and the two
// hooks of montecarlo: track_from_tgt (vertex -> field-free image track, simc.f:1425-1432) and track_to_tgt (iterated
// reconstruction against the field, simc.f:1573-1587).  Positions in cm, velocities in cm/ns, fields in T.
#pragma once
#include "target.cuh"

namespace simc {

constexpr int kFieldN = 51;                       // nz = nr = 51 nodes, 2 cm apart (trg_track.f:258-259, 312-313)
// map: [Bz(iz, ir) | Br(iz, ir)] of COMMON /trgFieldStrength/ in the file's reading order (index ir * 51 + iz); null
// unless set.  stht / ctht: sine and cosine of the angle between the field axis and each spectrometer
// (COMMON /trgFieldAngles_e/, /trgFieldAngles_p/): [0] electron arm (spect = -1), [1] hadron arm (spect = +1).
struct FieldDev { const double* map; double stht[2], ctht[2]; };

struct FieldState { double x, y, z, vx, vy, vz; };

namespace fielddetail {
SIMC_HD double sign1(double b) { return (b < 0. || (b == 0. && 1. / b < 0.)) ? -1. : 1.; }      // Fortran SIGN(1., b)

// trgField, trg_track.f:350-447 (B_phi is always treated as 0 there)
SIMC_HD void field_at(const FieldDev& F, int k, double px, double py, double pz, double& bx, double& by, double& bz) {
  const double stht = F.stht[k], ctht = F.ctht[k];
  const double x1 = px;
  const double x2 = stht * pz + ctht * py;
  const double x3 = ctht * pz - stht * py;
  const double z = fabs(x3);
  const double r = sqrt(x1 * x1 + x2 * x2);
  // zz(1) = rr(1) = 0, zz(2) - zz(1) = rr(2) - rr(1) = 2: the node spacing of trgInit
  const int i = (int)((z - 0.) / (2. - 0.)) + 1;
  const int j = (int)((r - 0.) / (2. - 0.)) + 1;
  if ((i + 1 > kFieldN) || (i < 1) || (j + 1 > kFieldN) || (j < 1)) { bx = 0.; by = 0.; bz = 0.; return; }
  const double* Bz = F.map;
  const double* Br = F.map + kFieldN * kFieldN;
  const int n00 = (j - 1) * kFieldN + (i - 1);         // (i, j), then (i+1, j), (i, j+1), (i+1, j+1)
  const double az = ((z - 2. * (double)(i - 1)) / (2. - 0.));
  const double ar = ((r - 2. * (double)(j - 1)) / (2. - 0.));
  double a0 = az * (Bz[n00 + 1] - Bz[n00]) + Bz[n00];
  double a1 = az * (Bz[n00 + kFieldN + 1] - Bz[n00 + kFieldN]) + Bz[n00 + kFieldN];
  const double B3 = (ar * (a1 - a0) + a0);
  if (r > 0.) {
    a0 = az * (Br[n00 + 1] - Br[n00]) + Br[n00];
    a1 = az * (Br[n00 + kFieldN + 1] - Br[n00 + kFieldN]) + Br[n00 + kFieldN];
    double B2 = (ar * (a1 - a0) + a0) / r;
    if (x3 < 0.) B2 = -B2;
    const double B1 = B2 * x1;
    B2 = B2 * x2;
    bx = B1;
    by = -stht * B3 + ctht * B2;
    bz = ctht * B3 + stht * B2;
  } else {
    bx = 0.;
    by = -stht * B3;
    bz = ctht * B3;
  }
}

struct Deriv { double dx, dy, dz, ax, ay, az; };
// trgDeriv, trg_track.f:452-490: velocity and (v x B) * factor
SIMC_HD Deriv deriv(const FieldDev& F, int k, double factor, const FieldState& u) {
  double bx, by, bz;
  field_at(F, k, u.x, u.y, u.z, bx, by, bz);
  Deriv d;
  d.dx = u.vx; d.dy = u.vy; d.dz = u.vz;
  const double c1 = u.vy * bz - u.vz * by;
  const double c2 = u.vz * bx - u.vx * bz;
  const double c3 = u.vx * by - u.vy * bx;
  d.ax = c1 * factor; d.ay = c2 * factor; d.az = c3 * factor;
  return d;
}
SIMC_HD FieldState advance(const FieldState& u0, double h, const Deriv& d) {
  FieldState u;
  u.x = u0.x + h * d.dx; u.y = u0.y + h * d.dy; u.z = u0.z + h * d.dz;
  u.vx = u0.vx + h * d.ax; u.vy = u0.vy + h * d.ay; u.vz = u0.vz + h * d.az;
  return u;
}
// trgRK4, trg_track.f:492-533 (Numerical Recipes' rk4)
SIMC_HD_CALL FieldState rk4(const FieldDev& F, int k, double factor, const FieldState& u0, double h) {
  const double hh = h * 0.5, h6 = h / 6.;
  const Deriv dudt = deriv(F, k, factor, u0);
  FieldState ut = advance(u0, hh, dudt);
  Deriv dut = deriv(F, k, factor, ut);
  ut = advance(u0, hh, dut);
  Deriv dum = deriv(F, k, factor, ut);
  ut = advance(u0, h, dum);
  dum.dx = dut.dx + dum.dx; dum.dy = dut.dy + dum.dy; dum.dz = dut.dz + dum.dz;
  dum.ax = dut.ax + dum.ax; dum.ay = dut.ay + dum.ay; dum.az = dut.az + dum.az;
  dut = deriv(F, k, factor, ut);
  FieldState u1;
  u1.x = u0.x + h6 * (dudt.dx + dut.dx + 2. * dum.dx);
  u1.y = u0.y + h6 * (dudt.dy + dut.dy + 2. * dum.dy);
  u1.z = u0.z + h6 * (dudt.dz + dut.dz + 2. * dum.dz);
  u1.vx = u0.vx + h6 * (dudt.ax + dut.ax + 2. * dum.ax);
  u1.vy = u0.vy + h6 * (dudt.ay + dut.ay + 2. * dum.ay);
  u1.vz = u0.vz + h6 * (dudt.az + dut.az + 2. * dum.az);
  return u1;
}
}  // namespace fielddetail

// trgTrackToPlane, trg_track.f:154-237: steps of dl along the track until the plane a x + b y + c z + d = 0 is crossed
// (two steps per trip of the loop, the direction chosen by a trial step), then a linear interpolation between the
// last two points.  ok false on entry: nothing happens.
SIMC_HD_CALL bool track_to_plane(const FieldDev& F, int k, FieldState& u, double E, double dl, double a, double b, double c, double d,
                                 bool ok) {
  using namespace fielddetail;
  if (!ok) return ok;
  const double n = 1 / sqrt(a * a + b * b + c * c);
  const double an = a * n, bn = b * n, cn = c * n, dn = d * n;
  const double factor = 90. / E;
  double ts = -dl / sqrt(u.vx * u.vx + u.vy * u.vy + u.vz * u.vz);
  double dist0 = u.x * an + u.y * bn + u.z * cn + dn;
  const double maxdist = fmax(fabs(dist0) * 4., 1.0);
  FieldState u1 = rk4(F, k, factor, u, ts), u0 = u;
  double dist1 = u1.x * an + u1.y * bn + u1.z * cn + dn;
  if ((sign1(dist0) == sign1(dist1)) && (fabs(dist0) < fabs(dist1))) ts = -ts;
  int steps = 0;
  const int max_steps = (int)(fmax(dist0, 10. * dl) / dl) * 10;
  if (sign1(dist0) == sign1(dist1)) {
    dist1 = dist0;
    while ((sign1(dist0) == sign1(dist1)) && ok) {
      u0 = rk4(F, k, factor, u1, ts);
      dist0 = u0.x * an + u0.y * bn + u0.z * cn + dn;
      if (sign1(dist0) == sign1(dist1)) {
        u1 = rk4(F, k, factor, u0, ts);
        dist1 = u1.x * an + u1.y * bn + u1.z * cn + dn;
      }
      ok = (fabs(dist1) < maxdist) && steps < max_steps;
      steps = steps + 1;
    }
  }
  if (ok) {
    const double f = dist0 / (dist0 - dist1);
    u.x = u0.x + (u1.x - u0.x) * f; u.y = u0.y + (u1.y - u0.y) * f; u.z = u0.z + (u1.z - u0.z) * f;
    u.vx = u0.vx + (u1.vx - u0.vx) * f; u.vy = u0.vy + (u1.vy - u0.vy) * f; u.vz = u0.vz + (u1.vz - u0.vz) * f;
  }
  return ok;
}

// track_from_tgt, trg_track.f:591-672: from the vertex (TRANSPORT coordinates of the arm) through the field to the
// plane z = 100 cm; the caller drifts the image track back to z = 0.  mom < 0 for negative particles.
SIMC_HD bool track_from_tgt(const FieldDev& F, int k, double& x, double& y, double& z, double& dx, double& dy, double mom, double mass) {
  using namespace fielddetail;
  const double cc = 29.9792458;
  const double vel = fabs(mom) / sqrt(mom * mom + mass * mass) * cc;
  const double eng = sign1(mom) * sqrt(mom * mom + mass * mass);
  FieldState v;
  v.x = x; v.y = y; v.z = z;
  v.vz = vel / sqrt(1 + dx * dx + dy * dy);
  v.vx = dx * v.vz;
  v.vy = dy * v.vz;
  track_to_plane(F, k, v, eng, 1., 0., 0., 1., 0., true);          // "for debugging, run track first to z=0"
  const bool ok = track_to_plane(F, k, v, eng, 1., 0., 0., 1., -100., true);
  x = v.x; y = v.y; z = v.z;
  dx = v.vx / v.vz;
  dy = v.vy / v.vz;
  return ok;
}

// track_to_tgt, trg_track.f:738-877: the arm's reconstruction repeated (at most ten times) with the vertical offset
// that makes the track, followed back through the field, meet the beam at the raster position.  RECON is the arm's
// mc_*_recon on the focal-plane track it holds: recon(delta, dy, dx, y, fry).
template <class RECON>
SIMC_HD bool track_to_tgt(const FieldDev& F, int k, double& delta, double& y, double& dx, double& dy, double frx, double fry, double mom,
                          double mass, double ctheta, double stheta, bool ok, RECON recon) {
  using namespace fielddetail;
  const double cc = 29.9792458;
  double xx = -fry;
  double vel = fabs(mom) / sqrt(mom * mom + mass * mass) * cc;
  double eng = sign1(mom) * sqrt(mom * mom + mass * mass);
  const double mom_0 = mom / (1.e0 + delta / 100.e0);
  FieldState vT, vTx;
  auto start = [&](double x0) {
    vT.x = x0 + 100. * dx;
    vT.y = y + 100. * dy;
    vT.z = 100.;
    vT.vz = vel / sqrt(1 + dy * dy + dx * dx);
    vT.vx = dx * vT.vz;
    vT.vy = dy * vT.vz;
  };
  start(-fry);
  ok = track_to_plane(F, k, vT, eng, 1., 0., -ctheta, stheta, frx, ok);
  int n = 0;
  double delx = 1.;
  while ((delx > .0001) && (n < 10) && ok) {
    delx = fabs(-fry - vT.x);
    vTx = vT;
    vTx.x = -fry;
    ok = track_to_plane(F, k, vT, eng, 1., 0., 0., 1., 0., ok);
    ok = track_to_plane(F, k, vTx, eng, 1., 0., 0., 1., 0., ok);
    xx = xx + fmin(1., fmax(-1., (vTx.x - vT.x)));
    recon(delta, dy, dx, y, xx);
    mom = mom_0 * (1.e0 + delta / 100.e0);
    vel = fabs(mom) / sqrt(mom * mom + mass * mass) * cc;
    eng = sign1(mom) * sqrt(mom * mom + mass * mass);
    start(xx);
    ok = track_to_plane(F, k, vT, eng, 1., 0., -ctheta, stheta, frx, ok);
    n = n + 1;
  }
  if (delx > .2) ok = false;
  dy = vT.vy / vT.vz;
  dx = vT.vx / vT.vz;
  y = vT.y;
  return ok;
}

}  // namespace simc
