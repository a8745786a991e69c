// ORACLE -- TEST INFRASTRUCTURE ONLY (see event.hpp).  PARITY UNPINNED.
// Semi-inclusive meson electroproduction weight, A(e,e'pi+-)X: peepiX, slacemcfit, Rhad_global
// (semi_physics.f:1-680), the CTEQ5 parton distributions Ctq5Pdf / PartonX / POLINT / ReadTbl
// (cteq5/Ctq5Pdf.f:69-352) and the free-nucleon branch of the Christy 2021 inclusive fit
// F1F2IN21 -> SF -> rescsp / rescsn -> RESMODP / RESMODN (F1F2IN21_v1.0.f:24-91, 189-250, 325-960,
// 2345-2375).  The reference is built with -fdefault-real-8 (Makefile:63), so every literal and every
// implicitly typed variable of these files is a double.
//
// Not restated: the kaon fragmentation functions fDSS (fdss/fdss.f; doing_semika), the "central"
// cross section (doing_cent = .true. is never passed: event.f:1521-1523), doing_pizero.
#include <cmath>
#include <stdexcept>

#include "event.hpp"

namespace simc_oracle {

using std::exp;
using std::log;
using std::pow;
using std::sqrt;

// ---- CTEQ5 ---------------------------------------------------------------------------------------
// ReadTbl, Ctq5Pdf.f:239-281: the Q grid is stored as Log(Q/Al)
void Cteq5Table::set(int nx, int nt, int nfmx, double al, double qini, double qmax, double xmin, const double* xv,
                     const double* qv, const double* upd) {
  Nx = nx; Nt = nt; NfMx = nfmx; Al = al; Alambda = al; Qini = qini; Qmax = qmax; Xmin = xmin;
  XV.assign(xv, xv + nx + 1);
  QL.resize(nt + 1);
  for (int iq = 0; iq <= nt; ++iq) QL[iq] = log(qv[iq] / al);
  const size_t npts = (size_t)(nx + 1) * (nt + 1) * (nfmx + 3);
  UPD.assign(upd, upd + npts);
}

// POLINT, Ctq5Pdf.f:308-352 ("Numerical Recipes"), arrays 1-based in the reference
static void polint(const double* XA, const double* YA, int N, double X, double& Y, double& DY) {
  double C[11], D[11];
  int NS = 1;
  double DIF = std::fabs(X - XA[0]);
  for (int I = 1; I <= N; ++I) {
    const double DIFT = std::fabs(X - XA[I - 1]);
    if (DIFT < DIF) { NS = I; DIF = DIFT; }
    C[I] = YA[I - 1];
    D[I] = YA[I - 1];
  }
  Y = YA[NS - 1];
  NS = NS - 1;
  for (int M = 1; M <= N - 1; ++M) {
    for (int I = 1; I <= N - M; ++I) {
      const double HO = XA[I - 1] - X;
      const double HP = XA[I + M - 1] - X;
      const double W = C[I + 1] - D[I];
      double DEN = HO - HP;
      if (DEN == 0.) throw std::runtime_error("POLINT: PAUSE (two equal abscissae)");
      DEN = W / DEN;
      D[I] = HP * DEN;
      C[I] = HO * DEN;
    }
    if (2 * NS < N - M) {
      DY = C[NS + 1];
    } else {
      DY = D[NS];
      NS = NS - 1;
    }
    Y = Y + DY;
  }
}

// PartonX, Ctq5Pdf.f:107-191 (M = 2: three-point interpolation in x, then in Log(Q/Lambda)).
// The reference clamps Jx < 0 only on its first call with X < Xmin (a SAVEd flag); X <= XV(0) = 0
// cannot occur here (Ctq5Pdf stops for X < 0, and X = Q2/(2 M nu) > 0), so the stateless clamp below
// is the same function.
static double PartonX(const Cteq5Table& T, int IPRTN, double X, double Q) {
  const int M = 2, M1 = M + 1;
  const int Nx = T.Nx, NT = T.Nt;
  double Fq[3], Df[3];
  const double QG = log(Q / T.Al);
  int JL = -1, JU = Nx + 1;
  while (JU - JL > 1) {
    const int JM = (JU + JL) / 2;
    if (X > T.XV[JM]) JL = JM; else JU = JM;
  }
  int Jx = JL - (M - 1) / 2;
  if (Jx < 0) Jx = 0;
  else if (Jx > Nx - M) Jx = Nx - M;
  JL = -1; JU = NT + 1;
  while (JU - JL > 1) {
    const int JM = (JU + JL) / 2;
    if (QG > T.QL[JM]) JL = JM; else JU = JM;
  }
  int Jq = JL - (M - 1) / 2;
  if (Jq < 0) Jq = 0;
  else if (Jq > NT - M) Jq = NT - M;
  const int Ip = IPRTN >= 3 ? -IPRTN : IPRTN;
  const int JFL = Ip + T.NfMx;
  const int J0 = (JFL * (NT + 1) + Jq) * (Nx + 1) + Jx;
  for (int Iq = 1; Iq <= M1; ++Iq) {
    const int J1 = J0 + (Nx + 1) * (Iq - 1) + 1;
    polint(&T.XV[Jx], &T.UPD[J1 - 1], M1, X, Fq[Iq - 1], Df[Iq - 1]);
  }
  double Ftmp, Ddf;
  polint(&T.QL[Jq], Fq, M1, QG, Ftmp, Ddf);
  return Ftmp;
}

// Ctq5Pdf, Ctq5Pdf.f:69-105.  Q is an in/out argument in the reference (raised to Alambda).
double Ctq5Pdf(const Cteq5Table& T, int Iparton, double X, double& Q) {
  if (X < 0. || X > 1.) throw std::runtime_error("X out of range in Ctq5Pdf");     // `Stop`
  if (Q < T.Alambda) Q = T.Alambda;
  if (Iparton < -T.NfMx || Iparton > T.NfMx) return 0.;
  double v = PartonX(T, Iparton, X, Q);
  if (v < 0.) v = 0.;
  return v;
}

// ---- DSS fragmentation functions ----------------------------------------------------------------------
static const double kFdssQS[24] = {1., 1.25, 1.5, 2.5, 4.0, 6.4, 1.0e1, 1.5e1, 2.5e1, 4.0e1, 6.4e1, 1.0e2, 1.8e2, 3.2e2,
                                   5.8e2, 1.0e3, 1.8e3, 3.2e3, 5.8e3, 1.0e4, 1.8e4, 3.2e4, 5.8e4, 1.0e5};
static const double kFdssXB[35] = {0.01, 0.02, 0.03, 0.04, 0.05, 0.06, 0.07, 0.08, 0.09, 0.095, 0.1, 0.125, 0.15, 0.175,
                                   0.2, 0.225, 0.25, 0.275, 0.3, 0.325, 0.35, 0.375, 0.4, 0.45, 0.5, 0.55, 0.6, 0.65,
                                   0.7, 0.75, 0.8, 0.85, 0.9, 0.93, 1.0};
// fdss/fdss.f:96-125: PARTON(k,N,M) / ((1-x)^4 x^0.5), zero at x = 1; ARRF = log of the two grids
void FdssTable::init(const double* parton) {
  const int NX = 35, NQ = 24;
  const int col[6] = {0, 1, 2, 6, 7, 8};       // UTOT, DTOT, STOT, UVAL, DVAL, SVAL
  for (int iq = 0; iq < NQ; ++iq) {
    for (int ix = 0; ix < NX - 1; ++ix) {
      const double XB0 = kFdssXB[ix], XB1 = 1. - kFdssXB[ix];
      for (int k = 0; k < 6; ++k)
        tab[k][iq * NX + ix] = parton[((size_t)ix * NQ + iq) * 9 + col[k]] / (powi(XB1, 4) * pow(XB0, 0.5));
    }
    for (int k = 0; k < 6; ++k) tab[k][iq * NX + NX - 1] = 0.;
  }
  for (int ix = 0; ix < NX; ++ix) arrf[ix] = log(kFdssXB[ix]);
  for (int iq = 0; iq < NQ; ++iq) arrf[NX + iq] = log(kFdssQS[iq]);
  set = true;
}
// fFINT (fdss/fdss.f:218-259) for NARG = 2: bilinear interpolation, linear extrapolation outside the grid
static double fFINT2(const double* ARG, const double* ENT, const double* TABLE) {
  const int NENT[2] = {35, 24};
  double D[2];
  int IENT[2];
  int KD = 1, M = 1, JA = 1;
  for (int I = 1; I <= 2; ++I) {
    const int JB = JA - 1 + NENT[I - 1];
    int J;
    for (J = JA; J <= JB; ++J)
      if (ARG[I - 1] <= ENT[J - 1]) break;
    if (J > JB) J = JB;
    if (J == JA) J = J + 1;
    const int JR = J - 1;
    D[I - 1] = (ENT[J - 1] - ARG[I - 1]) / (ENT[J - 1] - ENT[JR - 1]);
    IENT[I - 1] = J - JA;
    KD = KD + IENT[I - 1] * M;
    M = M * NENT[I - 1];
    JA = JB + 1;
  }
  double r = 0.;
  // NCOMB runs (1,1), (1,0), (0,1), (0,0); FAC multiplies the first argument's factor first
  r = r + ((1. * (1. - D[0])) * (1. - D[1])) * TABLE[KD - 1];
  r = r + ((1. * (1. - D[0])) * D[1]) * TABLE[KD - 35 - 1];
  r = r + ((1. * D[0]) * (1. - D[1])) * TABLE[KD - 1 - 1];
  r = r + ((1. * D[0]) * D[1]) * TABLE[KD - 1 - 35 - 1];
  return r;
}
// fDSS, fdss/fdss.f:1-215 (kaons at NLO; the charm, bottom and gluon functions are not used by peepiX).
// The reference forms (1.D0-X)**4 in REAL(16) (no -fdefault-double-8); the double product below
// differs from that in the last bit at most.
void fDSS(const FdssTable& T, int IC, double X, double Q2, double& U, double& UB, double& D, double& DB, double& S,
          double& SB) {
  const double XT[2] = {log(X), log(Q2)};
  const double shape = powi(1. - X, 4) * pow(X, 0.5);
  const double UTOT = fFINT2(XT, T.arrf, T.tab[0]) * shape;
  const double DTOT = fFINT2(XT, T.arrf, T.tab[1]) * shape;
  const double STOT = fFINT2(XT, T.arrf, T.tab[2]) * shape;
  const double UVAL = fFINT2(XT, T.arrf, T.tab[3]) * shape;
  const double DVAL = fFINT2(XT, T.arrf, T.tab[4]) * shape;
  const double SVAL = fFINT2(XT, T.arrf, T.tab[5]) * shape;
  const double Up = (UTOT + UVAL) / 2., UBp = (UTOT - UVAL) / 2.;
  const double Dp = (DTOT + DVAL) / 2., DBp = (DTOT - DVAL) / 2.;
  const double Sp = (STOT + SVAL) / 2., SBp = (STOT - SVAL) / 2.;
  if (IC == 1) { U = Up; UB = UBp; D = Dp; DB = DBp; S = Sp; SB = SBp; }
  else if (IC == -1) { U = UBp; UB = Up; D = DBp; DB = Dp; S = SBp; SB = Sp; }
  else throw std::runtime_error("fDSS: WRONG CHARGE");
}

// ---- Christy 2021 free-nucleon fit ----------------------------------------------------------------
// data xval of rescsp (F1F2IN21_v1.0.f:201-222) and xvaln of rescsn (:332-353)
static const double kXvalP[100] = {
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
static const double kXvalN[100] = {
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

// RESMODP (F1F2IN21_v1.0.f:373-669) and RESMODN (:672-958); the two differ in the nucleon mass, the
// seventh resonance's height, and the exponents of the non-resonant background.  xval is 1-based.
static double resmod(bool neutron, int sf, double w2, double q2, const double* xv) {
  auto xval = [&](int i) { return xv[i - 1]; };
  double mass[8], width[8], height[8], rescoef[7][5], nr_coef[4][5], sigr[8], wdif[3], intwidth[8];
  double kr[8], kcmr[8], ppicmr[8], ppi2cmr[8], petacmr[8], epicmr[8], epi2cmr[8], eetacmr[8];
  double br[8][4], ang[8], pgam[8], pwid[8][4], x0[8], xpr[3];
  const double mp = neutron ? 0.939565 : 0.9382727;
  const double mpi = 0.134977;
  const double meta = 0.547862;
  const double mp2 = mp * mp;
  const double W = sqrt(w2);
  const double w = W;
  wdif[1] = w - (mp + mpi);
  wdif[2] = w - (mp + meta);
  const double q20 = xval(50);
  br[1][1] = 1.00; br[2][1] = 0.45; br[3][1] = 0.60; br[4][1] = 0.65; br[5][1] = 0.60; br[6][1] = 0.65; br[7][1] = 0.60;
  br[1][3] = 0.0; br[2][3] = 0.40; br[3][3] = 0.08; br[4][3] = 0.0; br[5][3] = 0.20; br[6][3] = 0.0; br[7][3] = 0.0;
  for (int i = 1; i <= 7; ++i) br[i][2] = 1. - br[i][1] - br[i][3];
  ang[1] = 1.; ang[2] = 0.; ang[3] = 2.; ang[4] = 3.; ang[5] = 0.; ang[6] = 1.; ang[7] = 3.;
  for (int i = 1; i <= 7; ++i) x0[i] = 0.160;
  if (sf == 2) x0[1] = 0.07;
  const double mon = 1. / (1. + q2 / 1.5);      // (...)**1.
  const double xb = q2 / (q2 + w2 - mp2);
  xpr[1] = 1.00 + (w2 - (mp + mpi) * (mp + mpi)) / (q2 + q20);
  xpr[1] = 1. / xpr[1];
  xpr[2] = 1. + (w2 - (mp + meta) * (mp + meta)) / (q2 + q20);
  xpr[2] = 1. / xpr[2];
  if (w <= (mp + mpi)) xpr[1] = 1.0;
  if (w <= (mp + meta)) xpr[2] = 1.0;
  const double k = (w2 - mp2) / 2. / mp;
  const double kcm = (w2 - mp2) / 2. / w;
  const double epicm = (w2 + mpi * mpi - mp2) / 2. / w;
  const double ppicm = sqrt(std::max(0.0, (epicm * epicm - mpi * mpi)));
  const double epi2cm = (w2 + (2. * mpi) * (2. * mpi) - mp2) / 2. / w;
  const double ppi2cm = sqrt(std::max(0.0, (epi2cm * epi2cm - (2. * mpi) * (2. * mpi))));
  const double eetacm = (w2 + meta * meta - mp2) / 2. / w;
  const double petacm = sqrt(std::max(0.0, (eetacm * eetacm - meta * meta)));
  int num = 0;
  for (int i = 1; i <= 6; ++i) { num = num + 1; mass[i] = xval(i); }
  for (int i = 1; i <= 6; ++i) { num = num + 1; intwidth[i] = xval(num); width[i] = intwidth[i]; }
  mass[7] = xval(47);
  intwidth[7] = xval(48);
  width[7] = intwidth[7];
  for (int i = 1; i <= 7; ++i) {
    kr[i] = (mass[i] * mass[i] - mp2) / 2. / mp;
    kcmr[i] = (mass[i] * mass[i] - mp2) / 2. / mass[i];
    epicmr[i] = (mass[i] * mass[i] + mpi * mpi - mp2) / 2. / mass[i];
    ppicmr[i] = sqrt(std::max(0.0, (epicmr[i] * epicmr[i] - mpi * mpi)));
    epi2cmr[i] = (mass[i] * mass[i] + (2. * mpi) * (2. * mpi) - mp2) / 2. / mass[i];
    ppi2cmr[i] = sqrt(std::max(0.0, (epi2cmr[i] * epi2cmr[i] - (2. * mpi) * (2. * mpi))));
    eetacmr[i] = (mass[i] * mass[i] + meta * meta - mp2) / 2. / mass[i];
    petacmr[i] = sqrt(std::max(0.0, (eetacmr[i] * eetacmr[i] - meta * meta)));
    pwid[i][1] = intwidth[i] * pow(ppicm / ppicmr[i], 2. * ang[i] + 1.) *
                 pow((ppicmr[i] * ppicmr[i] + x0[i] * x0[i]) / (ppicm * ppicm + x0[i] * x0[i]), ang[i]);
    pwid[i][2] = intwidth[i] * pow(ppi2cm / ppi2cmr[i], 2. * ang[i] + 4.) *
                 pow((ppi2cmr[i] * ppi2cmr[i] + x0[i] * x0[i]) / (ppi2cm * ppi2cm + x0[i] * x0[i]), ang[i] + 2);
    pwid[i][2] = W / mass[i] * pwid[i][2];
    pwid[i][3] = 0.;
    if (i == 2 || i == 5) {
      pwid[i][3] = intwidth[i] * pow(petacm / petacmr[i], 2. * ang[i] + 1.) *
                   pow((petacmr[i] * petacmr[i] + x0[i] * x0[i]) / (petacm * petacm + x0[i] * x0[i]), ang[i]);
    }
    pgam[i] = (kcm / kcmr[i]) * (kcm / kcmr[i]) * (kcmr[i] * kcmr[i] + x0[i] * x0[i]) / (kcm * kcm + x0[i] * x0[i]);
    pgam[i] = intwidth[i] * pgam[i];
    width[i] = br[i][1] * pwid[i][1] + br[i][2] * pwid[i][2] + br[i][3] * pwid[i][3];
  }
  for (int i = 1; i <= 6; ++i) {
    for (int j = 1; j <= 4; ++j) { num = num + 1; rescoef[i][j] = xval(num); }
    if (sf == 1) {
      height[i] = rescoef[i][1] * (1. + rescoef[i][2] * q2 / (1. + rescoef[i][3] * q2)) * pow(mon, rescoef[i][4]);
    } else {
      height[i] = (rescoef[i][1] + rescoef[i][2] * q2) * exp(-1. * rescoef[i][3] * q2);
    }
    height[i] = height[i] * height[i];
  }
  if (sf == 2) {
    if (neutron) height[7] = (xval(44) + xval(45) * q2) * exp(-1.0 * xval(46) * q2);
    else height[7] = (xval(16) + xval(20) * q2) * exp(-1.0 * xval(24) * q2);
  } else {
    if (neutron) height[7] = xval(49) * mon;
    else height[7] = xval(49) * pow(mon, xval(45));
  }
  height[7] = height[7] * height[7];
  for (int i = 1; i <= 3; ++i)
    for (int j = 1; j <= 4; ++j) { num = num + 1; nr_coef[i][j] = xval(num); }
  double sig_res = 0.0;
  for (int i = 1; i <= 7; ++i) {
    const double dm = w2 - mass[i] * mass[i];
    const double mw = mass[i] * width[i];
    sigr[i] = width[i] * pgam[i] / (dm * dm + mw * mw);
    sigr[i] = height[i] * kr[i] / k * kcmr[i] / kcm * sigr[i] / intwidth[i];
    sig_res = sig_res + sigr[i];
  }
  sig_res = sig_res * w;
  if (sf == 2) sig_res = sig_res * q2;
  double sig_nr = 0.;
  if (sf == 1 && xpr[1] < 1.0) {
    const double A0 = xval(37) / pow(1.0 + q2 / xval(42), xval(43));
    double t1;
    if (neutron) t1 = xval(38) * log(1.05 + q2) + xval(39) / (1.05 + q2);
    else t1 = xval(38) * log(1.06 + q2) + xval(39) / log(1.06 + q2);
    const double t2 = xval(40) * pow(1.0 + q2 / xval(41), xval(44));
    if (xpr[1] <= 1.0) sig_nr = 389.4 * A0 * pow(1. - xpr[1], t1) * pow(xpr[1], t2);
    if (xpr[2] <= 1.0) sig_nr = sig_nr + xval(46) * 389.4 * A0 * pow(1. - xpr[2], t1) * pow(xpr[2], t2);
  } else if (sf == 2 && xpr[1] < 1.0) {
    const double d = 1.0 + q2 / xval(39);
    const double A0 = xval(37) / (d * d);
    const double t1 = xval(38) / (1.0 + q2 / (xval(40))) + xval(32) * log(q2 + xval(36));
    double t2;
    if (neutron) t2 = xval(41) / pow(1.00 + q2 / xval(42), xval(43));
    else t2 = xval(41);
    if (xpr[1] <= 1.0) sig_nr = sig_nr + 389.4 * A0 * xb * pow(1. - xpr[1], t1) * pow(xpr[1], t2);
  }
  double sig = sig_res + sig_nr;
  if ((w - mp) < wdif[1]) sig = 0.0;
  (void)nr_coef; (void)wdif;
  return sig;
}

// rescsp / rescsn: the parameter split of F1F2IN21_v1.0.f:224-234 and :355-361
static void rescs(bool neutron, double w2, double q2, double& sigT, double& sigL) {
  const double* xval = neutron ? kXvalN : kXvalP;
  double xval1[50], xvalL[50];
  for (int i = 1; i <= 50; ++i) {
    xval1[i - 1] = xval[i - 1];
    xvalL[i - 1] = xval[50 + i - 1];
    if (i <= 12) xvalL[i - 1] = xval1[i - 1];
    if (i == 47 || i == 48) xvalL[i - 1] = xval1[i - 1];
  }
  sigT = resmod(neutron, 1, w2, q2, xval1);
  sigL = resmod(neutron, 2, w2, q2, xvalL);
}

// SF, F1F2IN21_v1.0.f:2345-2375
void christy_sf(double w2, double q2, double& f1p, double& fLp, double& f2p, double& f1n, double& fLn, double& f2n) {
  const double mp = 0.938272;
  const double mp2 = mp * mp;
  const double pi = 3.14159;
  const double pi2 = pi * pi;
  const double alpha = 1 / 137.03599;
  const double x = q2 / (q2 + w2 - mp2);
  double sigTp, sigLp, sigTn, sigLn;
  rescs(false, w2, q2, sigTp, sigLp);
  rescs(true, w2, q2, sigTn, sigLn);
  f1p = sigTp / 0.3894e3 / pi2 / alpha / 8.0 * std::fabs(w2 - mp2);
  f1n = sigTn / 0.3894e3 / pi2 / alpha / 8.0 * std::fabs(w2 - mp2);
  fLp = sigLp * 2.0 * x / 0.3894e3 / pi2 / alpha / 8.0 * std::fabs(w2 - mp2);
  fLn = sigLn * 2.0 * x / 0.3894e3 / pi2 / alpha / 8.0 * std::fabs(w2 - mp2);
  f2p = (2. * x * f1p + fLp) / (1. + 4. * mp2 * x * x / q2);
  f2n = (2. * x * f1n + fLn) / (1. + 4. * mp2 * x * x / q2);
}

// F1F2IN21, F1F2IN21_v1.0.f:24-91, free nucleons only (IA < 2)
static void F1F2IN21(double Z, double A, double QSQ, double WSQ, double& F1, double& F2) {
  const int IA = (int)A, IZ = (int)Z;
  if (IA >= 2) throw std::runtime_error("oracle: F1F2IN21 restated for free nucleons only");
  double F1p, FLp, F2p, F1n, FLn, F2n;
  christy_sf(WSQ, QSQ, F1p, FLp, F2p, F1n, FLn, F2n);
  if (IZ < 1) { F1 = F1n; F2 = F2n; }
  else { F1 = F1p; F2 = F2p; }
}

// semi_physics.f:621-639
static double slacemcfit(double A, double x) {
  double r = 1.0;
  if (A > 2.0) {
    const double Atmp = A;
    const double alpha = -0.070 + 2.189 * x - 24.667 * powi(x, 2) + 145.291 * powi(x, 3) - 497.237 * powi(x, 4) +
                         1013.129 * powi(x, 5) - 1208.393 * powi(x, 6) + 775.767 * powi(x, 7) - 205.872 * powi(x, 8);
    const double C = exp(0.017 + 0.018 * log(x) + 0.005 * powi(log(x), 2));
    r = C * pow(Atmp, alpha);
  }
  return r;
}
// semi_physics.f:641-680
static double Rhad_global(double A, double z) {
  const double Ahyd = 1.0;
  if (A == Ahyd) return 1.0;
  const double Nzero = 0.98883 - 0.0038309 * A + 0.10841E-4 * powi(A, 2);
  const double alphah = 0.31953E-01 - 0.18659E-02 * A + 0.51747E-05 * powi(A, 2);
  const double Atmp = A < 83.8 ? A : 83.8;
  const double betah = 0.85475E-02 + 0.12763E-02 * Atmp - 0.24451E-05 * powi(Atmp, 2);
  return Nzero * pow(z, alphah) * pow(1 - z, betah);
}

// peepiX with doing_cent = .false., semi_physics.f:1-617
double peepiX(Sim& s, const Event& vertex, EventMain& main, double& survivalprob, SemiDebug* dbg) {
  const simc_run_config& cfg = *s.cfg;
  const simc_target& targ = cfg.targ;
  if (cfg.doing_semika && !s.fdss) throw std::runtime_error("oracle: fDSS table not set");
  if (!s.pdf) throw std::runtime_error("oracle: CTEQ5 table not set");
  static const double pf[12] = {1.0424, -0.1714, 1.8960, -0.0307, 0.1636, -0.1272, -4.2093, 5.0103, 2.7406, -0.5778, 3.5292, 7.3910};
  static const double pu[12] = {0.7840, 0.2369, 1.4238, 0.1484, 0.1518, -1.2923, -1.5710, 3.0305, 1.1995, 1.3553, 2.5868, 8.0666};
  const double qu = 2. / 3., qd = -1. / 3., qs = -1. / 3.;
  const double targA = targ.A, targZ = targ.Z, targN = targA - targZ;
  const double Mpi_gev = K::Mpi / 1000.0;
  const double Mp_gev = K::Mp / 1000.0;
  const double nu = vertex.nu;
  const double qx = vertex.uq.x * vertex.q, qy = vertex.uq.y * vertex.q, qz = vertex.uq.z * vertex.q;
  const double Q2 = vertex.Q2;
  const double Eb = vertex.Ein;
  const double Eprime = vertex.e.E;
  const double pt2 = vertex.pt2;
  const double zhad = vertex.zhad;
  const double mhad = cfg.doing_semika ? K::Mk : K::Mpi;
  const double mtar = targ.Mtar_struck;
  const double Ehad = zhad * nu;
  const double phad = sqrt(Ehad * Ehad - mhad * mhad);
  const double cthpq = cos(vertex.theta_pq);
  double xbj;
  if (cfg.do_fermi) {
    xbj = Q2 / 2. / (s.efer * nu - std::fabs(s.pfer) * (s.pferx * qx + s.pfery * qy + s.pferz * qz));
    s.ntup.xfermi = xbj;
  } else {
    xbj = Q2 / 2. / mtar / nu;
  }
  if (xbj > 1.0) xbj = 1.0;        // 'XBj is too large!'
  const double Q2gev = Q2 / 1.e6;
  double Qgev = sqrt(Q2gev);
  const double pt2gev = pt2 / 1.e6;
  double wsq = Mp_gev * Mp_gev + Q2gev * (1. / xbj - 1.);
  const double w = sqrt(wsq);
  const double mtargev = mtar / 1000.;
  const double nugev = nu / 1000.;
  const double mmpi2 = mtargev * mtargev + 2. * mtargev * nugev * (1 - zhad) * (1 - pt2gev);
  if (dbg) { *dbg = SemiDebug(); dbg->xbj = xbj; }
  if (mmpi2 < powi(mtargev + mhad / 1000., 2)) return 0.0;      // returns before davejac / sigcm are set
  const Cteq5Table& T = *s.pdf;
  // With do_fermi a hard nucleon moving along q can make P.q, and with it x, negative; Ctq5Pdf then `Stop`s the
  // reference (Ctq5Pdf.f:80-83).  Here the event is counted in simc_accum.unsupported and weighted zero.
  if (!(xbj >= 0.)) { s.low_w = true; return 0.0; }
  const double u = Ctq5Pdf(T, 1, xbj, Qgev);
  const double ubar = Ctq5Pdf(T, -1, xbj, Qgev);
  const double d = Ctq5Pdf(T, 2, xbj, Qgev);
  const double dbar = Ctq5Pdf(T, -2, xbj, Qgev);
  const double sq = Ctq5Pdf(T, 3, xbj, Qgev);
  const double sbar = Ctq5Pdf(T, -3, xbj, Qgev);
  const double uA = targZ * u + targN * d;
  const double ubarA = targZ * ubar + targN * dbar;
  const double dA = targZ * d + targN * u;
  const double dbarA = targZ * dbar + targN * ubar;
  const double sA = targZ * sq + targN * sq;
  const double sbarA = targZ * sbar + targN * sbar;
  const double sum_sq = qu * qu * (uA + ubarA) + qd * qd * (dA + dbarA) + qs * qs * (sA + sbarA);
  double u1, d1, ub, db, s1, sb;
  if (cfg.doing_semipi) {
    // Peter Bosted's fit of 9/20/2021, semi_physics.f:464-495
    const double xp = 2. * xbj / (1. + sqrt(1. + 4. * (xbj * xbj) * (Mp_gev * Mp_gev) / Q2gev));
    const double zp = (zhad / 2.) * (xp / xbj) *
                      (1. + sqrt(1 - 4 * (xbj * xbj) * (Mp_gev * Mp_gev) * (Mpi_gev * Mpi_gev + pt2gev) / (zhad * zhad) /
                                         (Q2gev * Q2gev)));
    const double sv = log(Q2gev / 2.);
    double yf = pf[0] * pow(zp, pf[1] + pf[3] * sv + pf[8] / w) * pow(1. - zp, pf[2] + pf[4] * sv + pf[9] / w);
    yf = yf * (1. + pf[5] * zp + pf[6] * (zp * zp) + pf[7] * powi(zp, 3)) * (1. + pf[10] / w + pf[11] / (w * w));
    double yu = pu[0] * pow(zp, pu[1] + pu[3] * sv + pu[8] / w) * pow(1. - zp, pu[2] + pu[4] * sv + pu[9] / w);
    yu = yu * (1. + pu[5] * zp + pu[6] * (zp * zp) + pu[7] * powi(zp, 3)) * (1. + pu[10] / w + pu[11] / (w * w));
    if (cfg.doing_hplus) { u1 = yf; d1 = yu; }
    else { u1 = yu; d1 = yf; }
    ub = d1; db = u1; s1 = yu; sb = s1;
  } else {
    // kaons: DSS fragmentation functions at NLO, semi_physics.f:496-506
    fDSS(*s.fdss, cfg.doing_hplus ? 1 : -1, zhad, Q2gev, u1, ub, d1, db, s1, sb);
  }
  const double dsigdz = (qu * qu * uA * u1 + qu * qu * ubarA * ub + qd * qd * dA * d1 + qd * qd * dbarA * db +
                         qs * qs * sA * s1 + qs * qs * sbarA * sb) / sum_sq / zhad;
  const double b = 1. / (0.120 * (zhad * zhad) + 0.200);
  const double sighad = Rhad_global(targA, zhad) * dsigdz * b * exp(-b * pt2gev) / 2. / K::pi;
  wsq = Mp_gev * Mp_gev + Q2gev * (1. / xbj - 1.);
  double F1, F2;
  F1F2IN21(1.0, 1.0, Q2gev, wsq, F1, F2);
  const double F1p = F1, F2p = F2;
  F1F2IN21(0.0, 1.0, Q2gev, wsq, F1, F2);
  const double F1n = F1, F2n = F2;
  F1 = (targZ * F1p + targN * F1n) * slacemcfit(targA, xbj);
  F2 = (targZ * F2p + targN * F2n) * slacemcfit(targA, xbj);
  const double W1 = F1 / (mtar / 1000.);
  const double W2 = F2 / (nu / 1000.);
  const double sin2th2 = Q2 / 4. / Eb / Eprime;
  const double cos2th2 = 1. - sin2th2;
  const double W2coeff = cos2th2;
  const double sige = 4. * (K::alpha * K::alpha) * powi(Eprime / 1000, 2) / (Q2gev * Q2gev) * (W2 * W2coeff + 2. * W1 * sin2th2);
  const double sigsemi = sige * sighad * powi(K::hbarc / 1000., 2) * 10000.0;
  const double jacobian = 1. / (nu / 1000.) * 2. * powi(phad / 1000., 2) * cthpq;
  double sigma_eepiX = sigsemi * jacobian / 1.e6;
  double fac;
  if (cfg.do_fermi) fac = 1. / (1. - s.pferz * s.pfer / s.efer) * mtar / s.efer;
  else fac = 1.0;
  sigma_eepiX = sigma_eepiX * fac;
  main.davejac = jacobian * 1000.0;
  s.ntup.sigcm = sighad;
  if (dbg) {
    dbg->xbj = xbj; dbg->u = u; dbg->ubar = ubar; dbg->d = d; dbg->dbar = dbar; dbg->s = sq; dbg->sbar = sbar;
    dbg->F1p = F1p; dbg->F2p = F2p; dbg->F1n = F1n; dbg->F2n = F2n; dbg->sighad = sighad; dbg->sige = sige;
  }
  // survival probability when decay is off, semi_physics.f:593-612 (zaero keeps 0 for arms without a branch)
  if (!cfg.doing_decay) {
    double zaero = 0.;
    if (cfg.hadron_arm == 1) zaero = -331.491;
    else if (cfg.hadron_arm == 2) zaero = -82.8;
    else if (cfg.hadron_arm == 3) zaero = -183.;
    else if (cfg.hadron_arm == 4) zaero = -183.;
    const double pathlen = main.FP_p.path + zaero * (1 + main.FP_p.dx * main.FP_p.dx + main.FP_p.dy * main.FP_p.dy);
    const double betak = cfg.spec_p.P / sqrt(cfg.spec_p.P * cfg.spec_p.P + cfg.Mh2);
    const double gammak = 1. / sqrt(1. - betak * betak);
    survivalprob = 1. / exp(pathlen / (cfg.ctau * betak * gammak));
    s.trk.decdist = survivalprob;
  }
  return sigma_eepiX;
}

}  // namespace simc_oracle
