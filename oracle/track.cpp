// ORACLE -- TEST INFRASTRUCTURE ONLY (see track.hpp).
#include "track.hpp"
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <stdexcept>

namespace simc_oracle {

static std::string pad_line(const std::string& s, size_t n) {
  std::string r = s;
  if (!r.empty() && r.back() == '\r') r.pop_back();
  if (r.size() < n) r.append(n - r.size(), ' ');
  return r;
}

// Fortran Gw.d input of a fixed-width field: blanks are ignored, the text carries its
// own decimal point in every shipped file.
static double read_real_field(const std::string& line, size_t pos, size_t width) {
  std::string f;
  for (size_t i = pos; i < pos + width; ++i)
    if (line[i] != ' ') f.push_back(line[i] == 'D' || line[i] == 'd' ? 'E' : line[i]);
  if (f.empty()) return 0.0;
  return std::strtod(f.c_str(), nullptr);
}
static int read_digit(const std::string& line, size_t pos) {
  const char c = line[pos];
  return (c == ' ') ? 0 : (c - '0');
}

// transp_init, shared/transp.f:294-474.  Line format 1200: (1x,5g14.7,1x,6i1).
void CosyForward::load(const std::string& path) {
  std::ifstream in(path);
  if (!in) throw std::runtime_error("TRANSP_INIT: cannot open " + path);
  const double coeff_min = 1.0e-14;
  cls.clear();
  std::string raw, line;
  double next_length = 0.;
  // strip header (transp.f:323-333)
  line = "!";
  while (!line.empty() && line[0] == '!') {
    if (!std::getline(in, raw)) throw std::runtime_error("TRANSP_INIT: unexpected EOF in header");
    line = pad_line(raw, 132);
    if (line.compare(0, 8, "!LENGTH:") == 0) next_length = 100. * std::strtod(line.c_str() + 9, nullptr);
  }
  for (;;) {
    CosyClass c;
    c.length = next_length;
    next_length = 0.;
    while (line.compare(0, 4, " ---") != 0) {
      double co[5];
      for (int i = 0; i < 5; ++i) co[i] = read_real_field(line, 1 + 14 * i, 14);
      int e[4];
      for (int j = 0; j < 4; ++j) e[j] = read_digit(line, 72 + j);
      const int idummy = read_digit(line, 76);
      const int e5 = read_digit(line, 77);
      if (idummy != 0) {
        // time-of-flight term: dropped (transp.f:384-391)
        if (co[0] != 0 || co[1] != 0 || co[2] != 0 || co[3] != 0)
          throw std::runtime_error("TRANSP_INIT: non-zero TOF terms!");
      } else {
        for (int i = 0; i < 5; ++i) c.coeff.push_back(co[i]);
        for (int j = 0; j < 4; ++j) c.expon.push_back((int8_t)e[j]);
        c.expon.push_back((int8_t)e5);
        c.n_terms++;
        if (c.adrift) {   // drift detection, transp.f:399-438
          const double c1 = co[0], c2 = co[1], c3 = co[2], c4 = co[3];
          const double csum = std::fabs(c1) + std::fabs(c2) + std::fabs(c3) + std::fabs(c4);
          const int order = e[0] + e[1] + e[2] + e[3];
          if (order == 1) {
            if (e[0] == 1) {
              if (std::fabs(c1 - 1.) > coeff_min) c.adrift = false;
              if (std::fabs(c2) > coeff_min) c.adrift = false;
              if (std::fabs(c3) > coeff_min) c.adrift = false;
              if (std::fabs(c4) > coeff_min) c.adrift = false;
            } else if (e[1] == 1) {
              c.driftdist = 1000. * c1;
              if (std::fabs(c2 - 1.) > coeff_min) c.adrift = false;
              if (std::fabs(c3) > coeff_min) c.adrift = false;
              if (std::fabs(c4) > coeff_min) c.adrift = false;
            } else if (e[2] == 1) {
              if (std::fabs(c1) > coeff_min) c.adrift = false;
              if (std::fabs(c2) > coeff_min) c.adrift = false;
              if (std::fabs(c3 - 1.) > coeff_min) c.adrift = false;
              if (std::fabs(c4) > coeff_min) c.adrift = false;
            } else if (e[3] == 1) {
              if (std::fabs(c1) > coeff_min) c.adrift = false;
              if (std::fabs(c2) > coeff_min) c.adrift = false;
              if (std::fabs(c.driftdist - 1000. * c3) > coeff_min) c.adrift = false;
              if (std::fabs(c4 - 1.) > coeff_min) c.adrift = false;
            }
          } else {
            if (std::fabs(csum) > coeff_min) c.adrift = false;
          }
        }
      }
      if (!std::getline(in, raw)) throw std::runtime_error("TRANSP_INIT: EOF inside a class");
      line = pad_line(raw, 132);
    }
    cls.push_back(std::move(c));
    // skip to the next data line (transp.f:456-470)
    bool eof = false;
    for (;;) {
      if (!std::getline(in, raw)) { eof = true; break; }
      line = pad_line(raw, 132);
      if (line.compare(0, 8, "!LENGTH:") == 0) next_length = 100. * std::strtod(line.c_str() + 9, nullptr);
      const bool blank = line.find_first_not_of(' ') == std::string::npos;
      if (line[0] == '!' || line.compare(0, 4, " ---") == 0 || blank) continue;
      break;
    }
    if (eof) break;
  }
}

// shared/transp.f:134-279
void transp(Track& t, const CosyForward& f, int klass, bool decay_flag, bool& dflag, double& m2, double& ph,
            double zd, double& pathlen) {
  const CosyClass& c = f.cls.at(klass - 1);
  if (t.calls) t.calls[klass - 1]++;
  double p_spec = 0, beta = 0, gamma = 0, z_decay = 0;
  if (decay_flag && !dflag) {
    p_spec = ph / (1. + t.dpps / 100.);
    beta = ph / std::sqrt(ph * ph + m2);
    gamma = 1. / std::sqrt(1. - beta * beta);
    const double dlen = t.ctau * beta * gamma;
    z_decay = -1. * dlen * std::log(1 - t.rng->grnd());
    if (z_decay <= zd / 2) {   // decay in first half: applied BEFORE the map
      dflag = true;
      t.decdist = t.decdist + z_decay;
      decay_in_flight(t, m2, ph, p_spec, beta, gamma, K::Mk);   // m_final = Mk quirk, transp.f:158
    }
  }
  double ray[5], sum[5] = {0., 0., 0., 0., 0.};
  ray[0] = t.xs;
  ray[1] = t.dxdzs * 1000.;
  ray[2] = t.ys;
  ray[3] = t.dydzs * 1000.;
  ray[4] = t.dpps;
  for (int i = 0; i < c.n_terms; ++i) {
    double term = 1.0;
    for (int j = 0; j < 5; ++j) {
      double temp = 1.0;
      const int e = c.expon[5 * i + j];
      if (e != 0) temp = powi(ray[j], e);
      term = term * temp;
    }
    sum[0] = sum[0] + term * c.coeff[5 * i + 0];
    sum[1] = sum[1] + term * c.coeff[5 * i + 1];
    sum[2] = sum[2] + term * c.coeff[5 * i + 2];
    sum[3] = sum[3] + term * c.coeff[5 * i + 3];
    sum[4] = sum[4] + term * c.coeff[5 * i + 4];
  }
  t.xs = sum[0];
  t.dxdzs = sum[1] / 1000.;
  t.ys = sum[2];
  t.dydzs = sum[3] / 1000.;
  const double delta_z = -sum[4];
  if (decay_flag && !dflag) {   // second half: applied AFTER the map, transp.f:231-276
    if (z_decay > zd + delta_z) {
      t.decdist = t.decdist + (zd + delta_z);
    } else {
      dflag = true;
      t.decdist = t.decdist + z_decay;
      decay_in_flight(t, m2, ph, p_spec, beta, gamma, K::Mpi);
    }
  }
  pathlen = pathlen + (zd + delta_z);
}

// hms/mc_hms_recon.f:70-102, format 1200: (1x,4g16.9,1x,5i1)
void CosyRecon::load(const std::string& path) {
  std::ifstream in(path);
  if (!in) throw std::runtime_error("MC_*_RECON: cannot open " + path);
  coeff.clear(); expon.clear(); n_terms = 0;
  std::string raw, line = "!";
  while (!line.empty() && line[0] == '!') {
    if (!std::getline(in, raw)) throw std::runtime_error("recon: unexpected EOF in header");
    line = pad_line(raw, 132);
  }
  while (line.compare(0, 4, " ---") != 0) {
    for (int i = 0; i < 4; ++i) coeff.push_back(read_real_field(line, 1 + 16 * i, 16));
    for (int j = 0; j < 5; ++j) expon.push_back((int8_t)read_digit(line, 66 + j));
    n_terms++;
    if (!std::getline(in, raw)) throw std::runtime_error("recon: EOF before terminator");
    line = pad_line(raw, 132);
  }
}

// hms/mc_hms_recon.f:104-137 (identical in shms/mc_shms_recon.f, sos, hrs)
void CosyRecon::eval(const Track& t, double fry, double& delta_p, double& delta_t, double& delta_phi,
                     double& y_tgt, bool clamp_all) const {
  double sum[4] = {0., 0., 0., 0.}, hut[5];
  hut[0] = t.xs / 100.;
  hut[1] = t.dxdzs;
  hut[2] = t.ys / 100.;
  hut[3] = t.dydzs;
  hut[4] = fry / 100.;
  if (std::fabs(hut[4]) <= 1.e-30) hut[4] = 1.e-30;
  if (clamp_all)
    for (int i = 0; i < 4; ++i)
      if (std::fabs(hut[i]) <= 1.e-30) hut[i] = 1.e-30;
  for (int i = 0; i < n_terms; ++i) {
    const int8_t* e = &expon[5 * i];
    const double term = powi(hut[0], e[0]) * powi(hut[1], e[1]) * powi(hut[2], e[2]) * powi(hut[3], e[3]) *
                        powi(hut[4], e[4]);
    sum[0] = sum[0] + term * coeff[4 * i + 0];
    sum[1] = sum[1] + term * coeff[4 * i + 1];
    sum[2] = sum[2] + term * coeff[4 * i + 2];
    sum[3] = sum[3] + term * coeff[4 * i + 3];
  }
  delta_phi = sum[0];
  y_tgt = sum[1] * 100.;
  delta_t = sum[2];
  delta_p = sum[3] * 100.;
}

}  // namespace simc_oracle
