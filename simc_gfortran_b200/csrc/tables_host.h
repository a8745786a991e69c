// Host-side readers of the reference's text tables that more than one translation unit needs.
// read_theory_file: theory_init's file format (init.f:838-880), used by simc_b200_load_theory_file and by
// the deck setup (VERTEXedge%Pm and E_Fermi come from the same file, init.f:326-343).
#pragma once
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

namespace simc {

struct TheoryFile {
  int n_shells = 0;
  double absorption = 0, e_fermi = 0;
  std::vector<double> nprot, em, emsig, bs_norm;      // per shell, as in the file (nprot not yet scaled)
  std::vector<int> n_pm;
  std::vector<double> pm_first, pm_bin, pm_last;
  std::vector<double> rho;                             // concatenated, as in the file (not yet normalised)
};

// "nrhoPm absorption E_Fermi", nrhoPm lines "nprot Em Emsig bs_norm", then rows "Pm rho": a shell ends where
// Pm stops increasing (init.f:860-868).  The reference insists on equidistant points (init.f:871-880).
inline TheoryFile read_theory_file(const std::string& path) {
  FILE* f = std::fopen(path.c_str(), "r");
  if (!f) throw std::runtime_error("cannot open theory file " + path);
  TheoryFile T;
  auto fail = [&](const char* what) { std::fclose(f); throw std::runtime_error("theory_init failed to read " + path + ": " + what); };
  if (std::fscanf(f, "%d %lf %lf", &T.n_shells, &T.absorption, &T.e_fermi) != 3) fail("header");
  if (T.n_shells < 1 || T.n_shells > 21) fail("1..21 momentum distributions (simulate.inc:118)");
  T.nprot.resize(T.n_shells); T.em.resize(T.n_shells); T.emsig.resize(T.n_shells); T.bs_norm.resize(T.n_shells);
  for (int m = 0; m < T.n_shells; ++m)
    if (std::fscanf(f, "%lf %lf %lf %lf", &T.nprot[m], &T.em[m], &T.emsig[m], &T.bs_norm[m]) != 4) fail("shell line");
  std::vector<double> pm, rho;
  double a, b;
  while (std::fscanf(f, "%lf %lf", &a, &b) == 2) { pm.push_back(a); rho.push_back(b); }
  std::fclose(f);
  size_t pos = 0;
  for (int m = 0; m < T.n_shells; ++m) {
    if (pos + 1 >= pm.size()) throw std::runtime_error("theory file " + path + ": too few rows");
    size_t end = pos + 1;
    while (end < pm.size() && pm[end] > pm[end - 1]) ++end;
    const int n = (int)(end - pos);
    if (n < 2 || n > 500) throw std::runtime_error("theory file " + path + ": 2..500 points per distribution (simulate.inc:117)");
    const double bin = pm[pos + 1] - pm[pos];
    const double mn = pm[pos] - bin / 2., mx = pm[end - 1] + bin / 2.;
    if (std::fabs(mn + bin * n - mx) > 0.1)
      throw std::runtime_error("theory_init found unequal Pm bins in distribution number " + std::to_string(m + 1));
    T.n_pm.push_back(n); T.pm_first.push_back(pm[pos]); T.pm_bin.push_back(bin); T.pm_last.push_back(pm[end - 1]);
    T.rho.insert(T.rho.end(), rho.begin() + pos, rho.begin() + end);
    pos = end;
  }
  return T;
}

// deut.dat / he3.dat / ...: rows "p  cumulative probability" with Fortran `d` exponents, at most 2000
// (dbase.f:581-584)
inline void read_pfermi_file(const std::string& path, std::vector<double>& pval, std::vector<double>& mprob) {
  FILE* f = std::fopen(path.c_str(), "r");
  if (!f) throw std::runtime_error("cannot open momentum distribution file " + path);
  char line[512];
  while (pval.size() < 2000 && std::fgets(line, sizeof line, f)) {
    for (char* c = line; *c; ++c) if (*c == 'd' || *c == 'D') *c = 'e';
    double p, q;
    if (std::sscanf(line, "%lf %lf", &p, &q) == 2) { pval.push_back(p); mprob.push_back(q); }
  }
  std::fclose(f);
  if (pval.size() < 2) throw std::runtime_error("momentum distribution file " + path + ": fewer than two rows");
}

// theory_file of init.f:838-851
inline const char* theory_file_for(int nA) {
  return nA == 2 ? "h2.theory" : nA == 12 ? "c12.theory" : nA == 56 ? "fe56.theory" : nA == 197 ? "au197.theory" : "c12.theory";
}

}  // namespace simc
