// Semi-inclusive meson electroproduction weight A(e,e'pi+-)X on the device (and on the host for tests
// of the table readers): replaces peepiX / slacemcfit / Rhad_global (semi_physics.f:1-680), the CTEQ5
// parton distributions Ctq5Pdf / PartonX / POLINT (cteq5/Ctq5Pdf.f:69-191, 308-352) and the free-nucleon
// branch of the Christy 2021 inclusive fit F1F2IN21 -> SF -> rescsp / rescsn -> RESMODP / RESMODN
// (F1F2IN21_v1.0.f:24-91, 189-250, 325-960, 2345-2375).
//
// Layout: the CTEQ5 table is one device buffer [xv | ql | upd] (82 KB for CTEQ5M, L2-resident; a thread
// touches 6 flavours x 9 values of it).  The six flavours of one event share the (x, Q) cell, so the two
// binary searches run once per event instead of once per call.  The reference evaluates the nucleon fit
// twice per event (F1F2IN21 for the proton, then for the neutron, each computing both); here once.
// Kaons (doing_semika) take the DSS fragmentation functions fDSS / fFINT (fdss/fdss.f) from a second table.
// Not built: doing_pizero, the "central" cross section (never requested: event.f:1521-1523).
#pragma once
#include "target.cuh"

namespace simc {

struct Cteq5Dev {               // device pointers into one buffer
  const double* xv;             // XV(0:Nx)
  const double* ql;             // Log(Q/Al), QL(0:Nt)
  const double* upd;            // UPD(1:(Nx+1)*(Nt+1)*(NfMx+3))
  int nx, nt, nfmx;
  double al;                    // Lambda = Alambda
};

struct PfermiDev {              // dbase.f:563-587: pval(1:nump), mprob(1:nump) with mprob(nump) = 1
  const double* pval;
  const double* mprob;
  int nump;
};

// SAVEd tables of fDSS after its first call (fdss/fdss.f:96-125), one device buffer:
// [ARRF: log x (35), log Q2 (24) | pad | XUTOTF, XDTOTF, XSTOTF, XUVALF, XDVALF, XSVALF as (35, 24) column-major]
struct FdssDev { const double* buf; };
constexpr int kFdssTab0 = 60;                 // offset of the first table in the buffer

// fFINT (fdss/fdss.f:218-259) for two arguments: the interval search of both axes, shared by the six tables
struct FdssCell { double d0, d1; int kd; };
SIMC_HD FdssCell fdss_cell(const FdssDev& T, double lx, double lq) {
  FdssCell c;
  int j = 1;                                  // first J in 1..35 with ARG <= ENT(J), else 35; never the first point
  while (j < 35 && !(lx <= T.buf[j - 1])) ++j;
  if (j == 1) j = 2;
  c.d0 = (T.buf[j - 1] - lx) / (T.buf[j - 1] - T.buf[j - 2]);
  int k = 1;
  while (k < 24 && !(lq <= T.buf[35 + k - 1])) ++k;
  if (k == 1) k = 2;
  c.d1 = (T.buf[35 + k - 1] - lq) / (T.buf[35 + k - 1] - T.buf[35 + k - 2]);
  c.kd = (j - 1) + (k - 1) * 35;              // 0-based index of TABLE(KD)
  return c;
}
SIMC_HD double fdss_interp(const double* tab, const FdssCell& c) {
  double r = 0.;
  r = r + ((1. * (1. - c.d0)) * (1. - c.d1)) * tab[c.kd];
  r = r + ((1. * (1. - c.d0)) * c.d1) * tab[c.kd - 35];
  r = r + ((1. * c.d0) * (1. - c.d1)) * tab[c.kd - 1];
  r = r + ((1. * c.d0) * c.d1) * tab[c.kd - 1 - 35];
  return r;
}
// fDSS (fdss/fdss.f:1-215) for kaons at NLO: z D(z, Q2) of u, ubar, d, dbar, s, sbar into a hadron of charge ic
SIMC_HD_CALL void fDSS(const FdssDev T, int ic, double X, double Q2, double* out6) {
  const FdssCell c = fdss_cell(T, m::log(X), m::log(Q2));
  const double x1 = 1. - X;
  const double x1s = x1 * x1;
  const double shape = (x1s * x1s) * m::pow(X, 0.5);
  const double* t0 = T.buf + kFdssTab0;
  const double UTOT = fdss_interp(t0, c) * shape, DTOT = fdss_interp(t0 + 840, c) * shape;
  const double STOT = fdss_interp(t0 + 2 * 840, c) * shape, UVAL = fdss_interp(t0 + 3 * 840, c) * shape;
  const double DVAL = fdss_interp(t0 + 4 * 840, c) * shape, SVAL = fdss_interp(t0 + 5 * 840, c) * shape;
  const double Up = (UTOT + UVAL) / 2., UBp = (UTOT - UVAL) / 2.;
  const double Dp = (DTOT + DVAL) / 2., DBp = (DTOT - DVAL) / 2.;
  const double Sp = (STOT + SVAL) / 2., SBp = (STOT - SVAL) / 2.;
  if (ic == 1) { out6[0] = Up; out6[1] = UBp; out6[2] = Dp; out6[3] = DBp; out6[4] = Sp; out6[5] = SBp; }
  else { out6[0] = UBp; out6[1] = Up; out6[2] = DBp; out6[3] = Dp; out6[4] = SBp; out6[5] = Sp; }
}

// POLINT with N = 3 (cteq5/Ctq5Pdf.f:308-352): Neville's scheme exactly as the reference walks it
SIMC_HD double polint3(const double xa0, const double xa1, const double xa2, const double ya0, const double ya1,
                       const double ya2, const double x) {
  int ns = 1;
  double dif = fabs(x - xa0);
  double dift = fabs(x - xa1);
  if (dift < dif) { ns = 2; dif = dift; }
  dift = fabs(x - xa2);
  if (dift < dif) { ns = 3; dif = dift; }
  double c1 = ya0, c2 = ya1, c3 = ya2, d1 = ya0, d2 = ya1, d3 = ya2;
  double y = ns == 1 ? ya0 : ns == 2 ? ya1 : ya2;
  ns = ns - 1;
  // M = 1
  {
    double ho = xa0 - x, hp = xa1 - x, w = c2 - d1, den = w / (ho - hp);
    d1 = hp * den; c1 = ho * den;
    ho = xa1 - x; hp = xa2 - x; w = c3 - d2; den = w / (ho - hp);
    d2 = hp * den; c2 = ho * den;
  }
  double dy;
  if (2 * ns < 2) { dy = ns == 0 ? c1 : c2; }
  else { dy = ns == 1 ? d1 : d2; ns = ns - 1; }
  y = y + dy;
  // M = 2
  {
    const double ho = xa0 - x, hp = xa2 - x, w = c2 - d1, den = w / (ho - hp);
    d1 = hp * den; c1 = ho * den;
  }
  if (2 * ns < 1) dy = c1;           // ns == 0 -> C(1)
  else dy = d1;                      // ns == 1 -> D(1)
  (void)c3; (void)d3;
  y = y + dy;
  return y;
}

// The (x, Q) cell of PartonX (Ctq5Pdf.f:125-166): shared by the six flavours of one event
struct PdfCell { int jx, jq; double qg; };
SIMC_HD PdfCell pdf_cell(const Cteq5Dev& T, double X, double Q) {
  PdfCell c;
  c.qg = m::log(Q / T.al);
  int jl = -1, ju = T.nx + 1;
  while (ju - jl > 1) {
    const int jm = (ju + jl) / 2;
    if (X > T.xv[jm]) jl = jm; else ju = jm;
  }
  c.jx = jl;                         // M = 2: Jx = JL - (M-1)/2 = JL
  if (c.jx < 0) c.jx = 0;
  else if (c.jx > T.nx - 2) c.jx = T.nx - 2;
  jl = -1; ju = T.nt + 1;
  while (ju - jl > 1) {
    const int jm = (ju + jl) / 2;
    if (c.qg > T.ql[jm]) jl = jm; else ju = jm;
  }
  c.jq = jl;
  if (c.jq < 0) c.jq = 0;
  else if (c.jq > T.nt - 2) c.jq = T.nt - 2;
  return c;
}
// Ctq5Pdf (Ctq5Pdf.f:69-105) for one flavour in a located cell; negative values are clamped to zero
SIMC_HD double ctq5pdf_cell(const Cteq5Dev& T, const PdfCell& c, int iparton, double X) {
  const int ip = iparton >= 3 ? -iparton : iparton;
  const int jfl = ip + T.nfmx;
  const double* u = T.upd + ((long long)(jfl * (T.nt + 1) + c.jq) * (T.nx + 1) + c.jx);
  const double x0 = T.xv[c.jx], x1 = T.xv[c.jx + 1], x2 = T.xv[c.jx + 2];
  const int row = T.nx + 1;
  const double f0 = polint3(x0, x1, x2, u[0], u[1], u[2], X);
  const double f1 = polint3(x0, x1, x2, u[row], u[row + 1], u[row + 2], X);
  const double f2 = polint3(x0, x1, x2, u[2 * row], u[2 * row + 1], u[2 * row + 2], X);
  const double v = polint3(T.ql[c.jq], T.ql[c.jq + 1], T.ql[c.jq + 2], f0, f1, f2, c.qg);
  return v < 0. ? 0. : v;
}

// ---- Christy 2021 free-nucleon fit -----------------------------------------------------------------
// data xval of rescsp (F1F2IN21_v1.0.f:201-222) and xvaln of rescsn (:332-353); the longitudinal set
// takes entries 1-12, 47 and 48 from the transverse one (:224-234).  i is the reference's 1-based index.
template <bool NEUTRON, int SF>
SIMC_HD double christy_xval(int i) {
  static constexpr double kP[100] = {
    0.12291E+01, 0.15173E+01, 0.15044E+01, 0.17100E+01, 0.16801E+01,
    0.14312E+01, 0.12616E+00, 0.23000E+00, 0.92594E-01, 0.90606E-01,
    0.75000E-01, 0.35067E+00, 0.75729E+01, 0.56091E+01, 0.94606E+01,
    0.20156E+01, 0.66190E+01, 0.41732E+00, 0.23980E-01, 0.53136E+01,
    0.63752E+00, 0.11484E+02, 0.69949E-01, 0.26191E+01, 0.53603E-01,
    0.65000E+02, 0.15351E+00, 0.20624E+01, 0.23408E+01, 0.16100E+02,
    0.62414E+02, 0.17201E+01, 0.23261E+00, 0.65000E+02, 0.23292E+01,
    0.14980E+01, 0.23000E+00, 0.63385E+00, 0.19093E-01, 0.61061E-01,
    0.29146E-02, 0.54388E+00, 0.77997E+00, 0.28783E+00, 0.10605E+01,
    0.69793E+00, 0.20009E+01, 0.57000E+00, 0.41632E+01, 0.38427E+00,
    0.10000E+01, 0.99842E+00, 0.98719E+00, 0.10168E+01, 0.98945E+00,
    0.99594E+00, 0.98799E+00, 0.10271E+01, 0.10650E+01, 0.97920E+00,
    0.10152E+01, 0.99622E+00, 0.81011E+01, 0.10070E-02, 0.14857E+01,
    0.33445E+01, 0.31641E-09, 0.69755E+02, 0.55228E+01, 0.14438E+00,
    0.60474E+01, 0.65395E-07, 0.14129E+01, 0.58609E+00, 0.36220E+01,
    0.92699E+00, 0.14418E+01, 0.86403E-02, 0.10001E-03, 0.75106E+00,
    0.76077E+00, 0.42272E+00, 0.55511E-11, 0.52486E+00, 0.58153E+00,
    0.15798E+01, 0.50105E+00, 0.89149E+02, 0.72789E+00, 0.24813E-01,
    -0.61906E+00, 0.10000E+01, 0.00000E+00, 0.00000E+00, 0.68158E+03,
    0.12429E+01, 0.00000E+00, 0.00000E+00, 0.00000E+00, 0.10000E-05,
  };
  static constexpr double kN[100] = {
    0.12291E+01, 0.15173E+01, 0.15044E+01, 0.17100E+01, 0.16801E+01,
    0.14312E+01, 0.12616E+00, 0.23000E+00, 0.92594E-01, 0.90606E-01,
    0.75000E-01, 0.35067E+00, 0.69500E+01, 0.86633E+01, 0.11557E+02,
    0.22138E+01, 0.44886E+01, 0.20500E+03, 0.84433E+03, 0.31167E+01,
    0.96301E+00, 0.14956E+00, 0.20761E-07, 0.10440E+01, 0.40143E-03,
    0.90028E+02, 0.75248E-01, 0.20532E+00, 0.12444E-01, 0.34469E+03,
    0.19948E+00, 0.26925E+01, 0.48635E+01, 0.86000E+02, 0.67813E+04,
    0.44281E+02, 0.29548E+00, 0.65421E+00, 0.23787E-09, 0.51967E-01,
    0.39926E-08, 0.29960E+00, 0.97516E+00, 0.46934E-01, 0.14246E+03,
    0.55801E+00, 0.19349E+01, 0.27400E+00, 0.38891E+00, 0.40000E-02,
    0.10108E+01, 0.97020E+00, 0.98248E+00, 0.97768E+00, 0.10425E+01,
    0.10198E+01, 0.97822E+00, 0.98239E+00, 0.10103E+01, 0.10076E+01,
    0.10044E+01, 0.99687E+00, 0.16696E+01, 0.10721E-06, 0.54114E+00,
    0.11923E+04, 0.55938E+02, 0.95000E+03, 0.39840E+02, 0.22026E+03,
    0.30498E+01, 0.24459E+00, 0.95574E+00, 0.35596E+00, 0.21228E-05,
    0.96696E+01, 0.27563E+01, 0.93027E-01, 0.33559E+02, 0.31207E-01,
    0.29020E+02, 0.86417E+00, 0.36471E-08, 0.99167E+00, 0.68124E+00,
    0.10000E-01, 0.90227E-01, 0.40115E+01, 0.29915E+01, 0.45929E-01,
    -0.16758E+01, 0.78493E+01, 0.78184E+01, 0.42074E+01, 0.41179E-05,
    0.80597E+00, 0.00000E+00, 0.00000E+00, 0.10045E+01, 0.62364E+00,
  };
  const int j = (SF == 1 || i <= 12 || i == 47 || i == 48) ? i - 1 : 50 + i - 1;
  return NEUTRON ? kN[j] : kP[j];
}

// RESMODP (F1F2IN21_v1.0.f:373-669) / RESMODN (:672-958): seven Breit-Wigners with energy-dependent
// widths plus a non-resonant background.  W2, Q2 in GeV^2.  Everything that depends on the fit parameters
// alone folds at compile time (the loops are unrolled over constant indices).
template <bool NEUTRON, int SF>
SIMC_HD_CALL double resmod(double w2, double q2) {
#define XV(i) christy_xval<NEUTRON, SF>(i)
  const double mp = NEUTRON ? 0.939565 : 0.9382727;
  const double mpi = 0.134977;
  const double meta = 0.547862;
  const double mp2 = mp * mp;
  const double w = sqrt(w2);
  const double wdif1 = w - (mp + mpi);
  const double q20 = XV(50);
  const double mon = 1. / (1. + q2 / 1.5);
  const double xb = q2 / (q2 + w2 - mp2);
  double xpr1 = 1.00 + (w2 - (mp + mpi) * (mp + mpi)) / (q2 + q20);
  xpr1 = 1. / xpr1;
  double xpr2 = 1. + (w2 - (mp + meta) * (mp + meta)) / (q2 + q20);
  xpr2 = 1. / xpr2;
  if (w <= (mp + mpi)) xpr1 = 1.0;
  if (w <= (mp + meta)) xpr2 = 1.0;
  const double k = (w2 - mp2) / 2. / mp;
  const double kcm = (w2 - mp2) / 2. / w;
  const double epicm = (w2 + mpi * mpi - mp2) / 2. / w;
  const double ppicm = sqrt(fmax(0.0, (epicm * epicm - mpi * mpi)));
  const double epi2cm = (w2 + (2. * mpi) * (2. * mpi) - mp2) / 2. / w;
  const double ppi2cm = sqrt(fmax(0.0, (epi2cm * epi2cm - (2. * mpi) * (2. * mpi))));
  const double eetacm = (w2 + meta * meta - mp2) / 2. / w;
  const double petacm = sqrt(fmax(0.0, (eetacm * eetacm - meta * meta)));
  double sig_res = 0.0;
#pragma unroll
  for (int i = 1; i <= 7; ++i) {
    const double mass = i <= 6 ? XV(i) : XV(47);
    const double intwidth = i <= 6 ? XV(6 + i) : XV(48);
    const double br1 = i == 1 ? 1.00 : i == 2 ? 0.45 : i == 3 ? 0.60 : i == 4 ? 0.65 : i == 5 ? 0.60 : i == 6 ? 0.65 : 0.60;
    const double br3 = i == 2 ? 0.40 : i == 3 ? 0.08 : i == 5 ? 0.20 : 0.0;
    const double br2 = 1. - br1 - br3;
    const double ang = i == 1 ? 1. : i == 2 ? 0. : i == 3 ? 2. : i == 4 ? 3. : i == 5 ? 0. : i == 6 ? 1. : 3.;
    const double x0 = (SF == 2 && i == 1) ? 0.07 : 0.160;
    const double kr = (mass * mass - mp2) / 2. / mp;
    const double kcmr = (mass * mass - mp2) / 2. / mass;
    const double epicmr = (mass * mass + mpi * mpi - mp2) / 2. / mass;
    const double ppicmr = sqrt(fmax(0.0, (epicmr * epicmr - mpi * mpi)));
    const double epi2cmr = (mass * mass + (2. * mpi) * (2. * mpi) - mp2) / 2. / mass;
    const double ppi2cmr = sqrt(fmax(0.0, (epi2cmr * epi2cmr - (2. * mpi) * (2. * mpi))));
    const double eetacmr = (mass * mass + meta * meta - mp2) / 2. / mass;
    const double petacmr = sqrt(fmax(0.0, (eetacmr * eetacmr - meta * meta)));
    const double pwid1 = intwidth * m::pow(ppicm / ppicmr, 2. * ang + 1.) *
                         m::pow((ppicmr * ppicmr + x0 * x0) / (ppicm * ppicm + x0 * x0), ang);
    double pwid2 = intwidth * m::pow(ppi2cm / ppi2cmr, 2. * ang + 4.) *
                   m::pow((ppi2cmr * ppi2cmr + x0 * x0) / (ppi2cm * ppi2cm + x0 * x0), ang + 2);
    pwid2 = w / mass * pwid2;
    double pwid3 = 0.;
    if (i == 2 || i == 5) {
      pwid3 = intwidth * m::pow(petacm / petacmr, 2. * ang + 1.) *
              m::pow((petacmr * petacmr + x0 * x0) / (petacm * petacm + x0 * x0), ang);
    }
    double pgam = (kcm / kcmr) * (kcm / kcmr) * (kcmr * kcmr + x0 * x0) / (kcm * kcm + x0 * x0);
    pgam = intwidth * pgam;
    const double width = br1 * pwid1 + br2 * pwid2 + br3 * pwid3;
    double height;
    if (i <= 6) {
      const int b = 12 + 4 * (i - 1);          // rescoef(i,1..4) = xval(13 + 4(i-1) ...)
      if (SF == 1) height = XV(b + 1) * (1. + XV(b + 2) * q2 / (1. + XV(b + 3) * q2)) * m::pow(mon, XV(b + 4));
      else height = (XV(b + 1) + XV(b + 2) * q2) * m::exp(-1. * XV(b + 3) * q2);
    } else if (SF == 2) {
      if (NEUTRON) height = (XV(44) + XV(45) * q2) * m::exp(-1.0 * XV(46) * q2);
      else height = (XV(16) + XV(20) * q2) * m::exp(-1.0 * XV(24) * q2);
    } else {
      if (NEUTRON) height = XV(49) * mon;
      else height = XV(49) * m::pow(mon, XV(45));
    }
    height = height * height;
    const double dm = w2 - mass * mass;
    const double mw = mass * width;
    double sigr = width * pgam / (dm * dm + mw * mw);
    sigr = height * kr / k * kcmr / kcm * sigr / intwidth;
    sig_res = sig_res + sigr;
  }
  sig_res = sig_res * w;
  if (SF == 2) sig_res = sig_res * q2;
  double sig_nr = 0.;
  if (SF == 1 && xpr1 < 1.0) {
    const double A0 = XV(37) / m::pow(1.0 + q2 / XV(42), XV(43));
    double t1;
    if (NEUTRON) t1 = XV(38) * m::log(1.05 + q2) + XV(39) / (1.05 + q2);
    else t1 = XV(38) * m::log(1.06 + q2) + XV(39) / m::log(1.06 + q2);
    const double t2 = XV(40) * m::pow(1.0 + q2 / XV(41), XV(44));
    if (xpr1 <= 1.0) sig_nr = 389.4 * A0 * m::pow(1. - xpr1, t1) * m::pow(xpr1, t2);
    if (xpr2 <= 1.0) sig_nr = sig_nr + XV(46) * 389.4 * A0 * m::pow(1. - xpr2, t1) * m::pow(xpr2, t2);
  } else if (SF == 2 && xpr1 < 1.0) {
    const double d = 1.0 + q2 / XV(39);
    const double A0 = XV(37) / (d * d);
    const double t1 = XV(38) / (1.0 + q2 / (XV(40))) + XV(32) * m::log(q2 + XV(36));
    double t2;
    if (NEUTRON) t2 = XV(41) / m::pow(1.00 + q2 / XV(42), XV(43));
    else t2 = XV(41);
    if (xpr1 <= 1.0) sig_nr = sig_nr + 389.4 * A0 * xb * m::pow(1. - xpr1, t1) * m::pow(xpr1, t2);
  }
  double sig = sig_res + sig_nr;
  if ((w - mp) < wdif1) sig = 0.0;
  return sig;
#undef XV
}

// SF (F1F2IN21_v1.0.f:2345-2375): F1, F2 of the free proton and neutron
struct NucleonSF { double f1p, f2p, f1n, f2n; };
SIMC_HD_CALL NucleonSF christy_sf(double w2, double q2) {
  const double mp = 0.938272;
  const double mp2 = mp * mp;
  const double pi = 3.14159;
  const double pi2 = pi * pi;
  const double alpha = 1 / 137.03599;
  const double x = q2 / (q2 + w2 - mp2);
  const double sigTp = resmod<false, 1>(w2, q2), sigLp = resmod<false, 2>(w2, q2);
  const double sigTn = resmod<true, 1>(w2, q2), sigLn = resmod<true, 2>(w2, q2);
  NucleonSF r;
  r.f1p = sigTp / 0.3894e3 / pi2 / alpha / 8.0 * fabs(w2 - mp2);
  r.f1n = sigTn / 0.3894e3 / pi2 / alpha / 8.0 * fabs(w2 - mp2);
  const double fLp = sigLp * 2.0 * x / 0.3894e3 / pi2 / alpha / 8.0 * fabs(w2 - mp2);
  const double fLn = sigLn * 2.0 * x / 0.3894e3 / pi2 / alpha / 8.0 * fabs(w2 - mp2);
  r.f2p = (2. * x * r.f1p + fLp) / (1. + 4. * mp2 * x * x / q2);
  r.f2n = (2. * x * r.f1n + fLn) / (1. + 4. * mp2 * x * x / q2);
  return r;
}

// semi_physics.f:621-639 (x**n with integer n: libgcc's __powidf2 association)
SIMC_HD double slacemcfit(double A, double x) {
  if (!(A > 2.0)) return 1.0;
  const double x2 = x * x, x3 = x * x2, x4 = x2 * x2, x5 = x * x4, x6 = x2 * x4, x7 = (x * x2) * x4, x8 = x4 * x4;
  const double alpha = -0.070 + 2.189 * x - 24.667 * x2 + 145.291 * x3 - 497.237 * x4 + 1013.129 * x5 - 1208.393 * x6 +
                       775.767 * x7 - 205.872 * x8;
  const double lx = m::log(x);
  const double C = m::exp(0.017 + 0.018 * lx + 0.005 * (lx * lx));
  return C * m::pow(A, alpha);
}
// semi_physics.f:641-680
SIMC_HD double Rhad_global(double A, double z) {
  if (A == 1.0) return 1.0;
  const double Nzero = 0.98883 - 0.0038309 * A + 0.10841E-4 * (A * A);
  const double alphah = 0.31953E-01 - 0.18659E-02 * A + 0.51747E-05 * (A * A);
  const double Atmp = A < 83.8 ? A : 83.8;
  const double betah = 0.85475E-02 + 0.12763E-02 * Atmp - 0.24451E-05 * (Atmp * Atmp);
  return Nzero * m::pow(z, alphah) * m::pow(1 - z, betah);
}

struct SemiVertex {            // what peepiX reads from `vertex`, `main` and COMMON /pfermi_stuff/
  double Ein, eE, nu, Q2, q, uqx, uqy, uqz, pt2, zhad, theta_pq;
  double pfer, pferx, pfery, pferz, efer;
};
struct SemiWeight {
  double sigcc, sighad, davejac, xbj, xfermi;
  bool bad;                    // x < 0 (possible with do_fermi): a `Stop` in Ctq5Pdf, Ctq5Pdf.f:80-83
  bool early;                  // returned zero before the parton densities (threshold, semi_physics.f:283-287)
};

// peepiX with doing_cent = .false. (semi_physics.f:1-617), pions.  dbg (may be null) receives
// { u, ubar, d, dbar, s, sbar, F1p, F2p, F1n, F2n, sige } for the stage-level parity entry point.
SIMC_HD_CALL SemiWeight peepiX(const simc_run_config& cfg, const Cteq5Dev T, const FdssDev F, const SemiVertex& v,
                               double* dbg) {
  const double pf[12] = {1.0424, -0.1714, 1.8960, -0.0307, 0.1636, -0.1272, -4.2093, 5.0103, 2.7406, -0.5778, 3.5292, 7.3910};
  const double pu[12] = {0.7840, 0.2369, 1.4238, 0.1484, 0.1518, -1.2923, -1.5710, 3.0305, 1.1995, 1.3553, 2.5868, 8.0666};
  const double Mpi = 139.57018, Mp = 938.27231, hbarc = 197.327053, alpha = 1. / 137.0359895, pi = 3.141592653589793;
  const double qu = 2. / 3., qd = -1. / 3., qs = -1. / 3.;
  SemiWeight R;
  R.sigcc = 0.0; R.sighad = 0.0; R.davejac = 0.0; R.xbj = 0.0; R.xfermi = 0.0; R.bad = false; R.early = true;
  const simc_target& targ = cfg.targ;
  const double targA = targ.A, targZ = targ.Z, targN = targA - targZ;
  const double Mpi_gev = Mpi / 1000.0, Mp_gev = Mp / 1000.0;
  const double nu = v.nu, Q2 = v.Q2, Eb = v.Ein, Eprime = v.eE, pt2 = v.pt2, zhad = v.zhad;
  const double qx = v.uqx * v.q, qy = v.uqy * v.q, qz = v.uqz * v.q;
  const double mhad = cfg.doing_semika ? 493.677 : Mpi;
  const double mtar = targ.Mtar_struck;
  const double Ehad = zhad * nu;
  const double phad = sqrt(Ehad * Ehad - mhad * mhad);
  const double cthpq = m::cos(v.theta_pq);
  double xbj;
  if (cfg.do_fermi) xbj = Q2 / 2. / (v.efer * nu - fabs(v.pfer) * (v.pferx * qx + v.pfery * qy + v.pferz * qz));
  else xbj = Q2 / 2. / mtar / nu;
  if (cfg.do_fermi) R.xfermi = xbj;          // ntup%xfermi is stored before the clip (semi_physics.f:249-258)
  if (xbj > 1.0) xbj = 1.0;
  R.xbj = xbj;
  const double Q2gev = Q2 / 1.e6;
  double Qgev = sqrt(Q2gev);
  const double pt2gev = pt2 / 1.e6;
  const double wsq = Mp_gev * Mp_gev + Q2gev * (1. / xbj - 1.);
  const double w = sqrt(wsq);
  const double mtargev = mtar / 1000.;
  const double nugev = nu / 1000.;
  const double mmpi2 = mtargev * mtargev + 2. * mtargev * nugev * (1 - zhad) * (1 - pt2gev);
  {
    const double thr = mtargev + mhad / 1000.;
    if (mmpi2 < thr * thr) return R;
  }
  if (!(xbj >= 0.)) { R.bad = true; return R; }
  if (Qgev < T.al) Qgev = T.al;
  const PdfCell cell = pdf_cell(T, xbj, Qgev);
  const double u = ctq5pdf_cell(T, cell, 1, xbj);
  const double ubar = ctq5pdf_cell(T, cell, -1, xbj);
  const double d = ctq5pdf_cell(T, cell, 2, xbj);
  const double dbar = ctq5pdf_cell(T, cell, -2, xbj);
  const double sq = ctq5pdf_cell(T, cell, 3, xbj);
  const double sbar = ctq5pdf_cell(T, cell, -3, xbj);
  const double uA = targZ * u + targN * d;
  const double ubarA = targZ * ubar + targN * dbar;
  const double dA = targZ * d + targN * u;
  const double dbarA = targZ * dbar + targN * ubar;
  const double sA = targZ * sq + targN * sq;
  const double sbarA = targZ * sbar + targN * sbar;
  const double sum_sq = qu * qu * (uA + ubarA) + qd * qd * (dA + dbarA) + qs * qs * (sA + sbarA);
  double u1, d1, ub, db, s1, sb;
  if (cfg.doing_semipi) {
    // Bosted's fragmentation fit of 9/20/2021 in the modified scaling variable zp, semi_physics.f:464-495
    const double xp = 2. * xbj / (1. + sqrt(1. + 4. * (xbj * xbj) * (Mp_gev * Mp_gev) / Q2gev));
    const double zp = (zhad / 2.) * (xp / xbj) *
                      (1. + sqrt(1 - 4 * (xbj * xbj) * (Mp_gev * Mp_gev) * (Mpi_gev * Mpi_gev + pt2gev) / (zhad * zhad) /
                                         (Q2gev * Q2gev)));
    const double sv = m::log(Q2gev / 2.);
    const double zp3 = zp * (zp * zp);
    double yf = pf[0] * m::pow(zp, pf[1] + pf[3] * sv + pf[8] / w) * m::pow(1. - zp, pf[2] + pf[4] * sv + pf[9] / w);
    yf = yf * (1. + pf[5] * zp + pf[6] * (zp * zp) + pf[7] * zp3) * (1. + pf[10] / w + pf[11] / (w * w));
    double yu = pu[0] * m::pow(zp, pu[1] + pu[3] * sv + pu[8] / w) * m::pow(1. - zp, pu[2] + pu[4] * sv + pu[9] / w);
    yu = yu * (1. + pu[5] * zp + pu[6] * (zp * zp) + pu[7] * zp3) * (1. + pu[10] / w + pu[11] / (w * w));
    if (cfg.doing_hplus) { u1 = yf; d1 = yu; }
    else { u1 = yu; d1 = yf; }
    ub = d1; db = u1; s1 = yu; sb = s1;
  } else {
    // kaons: DSS fragmentation functions at NLO, semi_physics.f:496-506
    double ff[6];
    fDSS(F, cfg.doing_hplus ? 1 : -1, zhad, Q2gev, ff);
    u1 = ff[0]; ub = ff[1]; d1 = ff[2]; db = ff[3]; s1 = ff[4]; sb = ff[5];
  }
  const double dsigdz = (qu * qu * uA * u1 + qu * qu * ubarA * ub + qd * qd * dA * d1 + qd * qd * dbarA * db +
                         qs * qs * sA * s1 + qs * qs * sbarA * sb) / sum_sq / zhad;
  const double b = 1. / (0.120 * (zhad * zhad) + 0.200);
  const double sighad = Rhad_global(targA, zhad) * dsigdz * b * m::exp(-b * pt2gev) / 2. / pi;
  const NucleonSF N = christy_sf(wsq, Q2gev);
  const double emc = slacemcfit(targA, xbj);
  const double F1 = (targZ * N.f1p + targN * N.f1n) * emc;
  const double F2 = (targZ * N.f2p + targN * N.f2n) * emc;
  const double W1 = F1 / (mtar / 1000.);
  const double W2 = F2 / (nu / 1000.);
  const double sin2th2 = Q2 / 4. / Eb / Eprime;
  const double cos2th2 = 1. - sin2th2;
  const double ep = Eprime / 1000;
  const double sige = 4. * (alpha * alpha) * (ep * ep) / (Q2gev * Q2gev) * (W2 * cos2th2 + 2. * W1 * sin2th2);
  const double hb = hbarc / 1000.;
  const double sigsemi = sige * sighad * (hb * hb) * 10000.0;
  const double pg = phad / 1000.;
  const double jacobian = 1. / (nu / 1000.) * 2. * (pg * pg) * cthpq;
  double sigma = sigsemi * jacobian / 1.e6;
  double fac = 1.0;
  if (cfg.do_fermi) fac = 1. / (1. - v.pferz * v.pfer / v.efer) * mtar / v.efer;
  sigma = sigma * fac;
  R.early = false;
  R.sigcc = sigma;
  R.sighad = sighad;
  R.davejac = jacobian * 1000.0;
  if (dbg) {
    dbg[0] = u; dbg[1] = ubar; dbg[2] = d; dbg[3] = dbar; dbg[4] = sq; dbg[5] = sbar;
    dbg[6] = N.f1p; dbg[7] = N.f2p; dbg[8] = N.f1n; dbg[9] = N.f2n; dbg[10] = sige;
  }
  return R;
}

// Survival probability to the aerogel when decay is off, semi_physics.f:593-612 (zaero keeps 0 for arms
// without a branch: -fno-automatic locals start at zero)
SIMC_HD double semi_survival(const simc_run_config& cfg, double fp_path, double fp_dx, double fp_dy) {
  double zaero = 0.;
  if (cfg.hadron_arm == 1) zaero = -331.491;
  else if (cfg.hadron_arm == 2) zaero = -82.8;
  else if (cfg.hadron_arm == 3) zaero = -183.;
  else if (cfg.hadron_arm == 4) zaero = -183.;
  const double pathlen = fp_path + zaero * (1 + fp_dx * fp_dx + fp_dy * fp_dy);
  const double betak = cfg.spec_p.P / sqrt(cfg.spec_p.P * cfg.spec_p.P + cfg.Mh2);
  const double gammak = 1. / sqrt(1. - betak * betak);
  return 1. / m::exp(pathlen / (cfg.ctau * betak * gammak));
}

}  // namespace simc
