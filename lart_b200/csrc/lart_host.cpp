// lart_host.cpp — C++ mini-host (see include/lart_host.h).
//
// Restates, for the Cartesian Ly-alpha path only, the host-side steps LaRT's
// Fortran driver performs before and after `call run_simulation(grid)`:
//   read_input            src/setup.f90:4-562        (namelist subset + derived defaults)
//   setup_resonance_line  src/line_mod.f90:551-1270  (Ly-alpha branch :1241-1270)
//   setup_scattering_matrix src/setup.f90:581-649
//   grid_create           src/grid_mod_car.f90:11-1238 (uniform / sphere / analytic velocity fields)
//   car_setup_freq_grid   src/grid_mod_car.f90:1442-1511
//   observer_create_outside src/observer_rect.f90:10-300
//   output_normalize_outside src/output_sum_rect.f90:151-487
// It produces HOST arrays in the layout grid_type holds them (column-major, x
// fastest) and fills a lart_config for the C ABI.  No transport code, no GPU code.
//
// Not restated (the real host keeps them): file-based density/temperature/velocity
// inputs, symmetry-folded grids, atmospheres, clumps, AMR, HEALPix observers,
// HDF5/FITS output.  Unknown namelist keys are an error unless they are in the
// "accepted and ignored" list below (output/bookkeeping switches of the Fortran).

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/lart_host.h"

namespace {

constexpr double kPi = 3.141592653589793238462643383279502884197;
constexpr double kTwoPi = 2.0 * kPi;
constexpr double kFourPi = 4.0 * kPi;
constexpr double kDeg2Rad = kPi / 180.0;
constexpr double kRad2Deg = 180.0 / kPi;
constexpr double kNaN = std::numeric_limits<double>::quiet_NaN();
constexpr double kSpeedC = 2.99792458e5;   // km/s, define.f90:71
constexpr double kHPlanck = 6.62607004e-34;  // define.f90:72
constexpr double kUm2Km = 1.0e-9;          // define.f90:65
constexpr double kUm2M = 1.0e-6;           // define.f90:64
constexpr double kEps = std::numeric_limits<double>::epsilon();

std::string g_err;
inline bool isfin(double v) { return std::isfinite(v); }

std::string lower(std::string s) {
  for (auto &c : s) c = static_cast<char>(std::tolower(static_cast<unsigned char>(c)));
  return s;
}
std::string trim(const std::string &s) {
  size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
  return (a == std::string::npos) ? std::string() : s.substr(a, b - a + 1);
}

// params_type defaults — define.f90:209-544 (members this path reads)
struct Par {
  double no_photons = 1e5;
  int64_t nphotons = 100000;
  int64_t iseed = 0;
  double temperature = 1e4, temperature0 = -999.0, bturb = -999.0, Dfreq0 = -999.0, voigt_a0 = -999.0;
  std::string line_id = "ly_alpha";
  bool fine_structure = false;
  double taumax = -999.0, tauhomo = -999.0, tau0 = -999.0;
  double N_HImax = -999.0, N_HIhomo = -999.0, N_HI = -999.0, N_gasmax = -999.0, N_gashomo = -999.0;
  double atau3 = 0.0;
  double Vexp = 0.0, Vx = 0.0, Vy = 0.0, Vz = 0.0, rpeak = 0.0, Vrot = 0.0, rinner = 0.0;
  // clump medium — define.f90:326-352
  bool use_amr_grid = false;  // define.f90: par%use_amr_grid (leaf data handed over with lart_host_set_amr_leaves)
  bool use_clump_medium = false, clump_fully_inside = true, clump_allow_overlap = false;
  double clump_radius = -1.0, clump_N_clumps = -1.0, clump_f_vol = -1.0, clump_f_cov = -1.0, clump_tau0 = -1.0,
         clump_NHI = -1.0, clump_sigma_v = 0.0;
  bool comoving_source = true, recoil = false, core_skip = false, core_skip_global = false;
  bool xyz_symmetry = false, xy_symmetry = false, xy_periodic = false, z_symmetry = false;
  double Omega = 0.0;  // define.f90:312 — taken as the value the ray tracer adds (after grid_mod_car.f90:348-350)
  // the reference's compile-time options -DCALCJ / -DCALCP / -DCALCPnew as run-time keys of the mini-host
  bool calc_J = false, calc_P = false, calc_Pnew = false, save_all = false;
  int geometry_JPa = 2147483647;  // define.f90:269 huge(1)
  std::string geometry, velocity_type;
  int nx = 1, ny = 1, nz = 11, nr = -999;
  double xmax = 1.0, ymax = 1.0, zmax = 1.0, rmin = -999.0, rmax = -999.0, source_rmax = -999.0;
  double zmin = kNaN;  // define.f90: par%zmin, read by the plane-atmosphere geometry only
  double density_rscale = -999.9, density_zscale = -999.9, density_alpha = 0.0, velocity_alpha = 1.0;
  double xs_point = 0.0, ys_point = 0.0, zs_point = 0.0;
  double xfreq0 = 0.0, xfreq_min = kNaN, xfreq_max = kNaN;
  int nxfreq = 121;
  double velocity_min = kNaN, velocity_max = kNaN;
  int nvelocity = 0;
  double distance2cm = -999.9;
  double gaussian_sigma_vel = 12.843374, gaussian_FWHM_vel = -999.0;
  std::string distance_unit, source_geometry = "point", spectral_type = "voigt";
  bool use_reduced_wgt = false, use_stokes = false;
  bool save_Jin = true, save_Jabs = true, save_Jmu = false;
  bool continuum_normalize = true;  // define.f90:314
  double f_line = 0.0;              // define.f90:320 (internal: line fraction of 'continuum+gaussian', not on this path)
  int nmu = 11;
  double mu_min = -1.0, dmu = 0.0;
  bool save_direc0 = false, save_all_photons = false;
  bool save_peeloff = false, save_peeloff_2D = false, save_peeloff_3D = true;
  int intensity_unit = -999;
  double hgg = 0.6761, albedo = 0.3253, cext_dust = 1.6059e-21, DGR = 0.0;
  std::string scatt_mat_file;
  int nobs = 1, nxim = 0, nyim = 0;
  double distance = kNaN;
  std::vector<double> inclination_angle, position_angle, phase_angle, alpha, beta, gamma, obsx, obsy, obsz;
  double dxim = kNaN, dyim = kNaN;
  double rotation_center_x = kNaN, rotation_center_y = kNaN, rotation_center_z = kNaN;
  double nscatt_gas = 0, nscatt_dust = 0, nscatt_tot = 0;
};

// line_type — define.f90:639-680
struct Line {
  int line_type = 1;
  double wavelength0 = 0.1215668237310, damping = 6.2649e8, cross0 = 0, vtherm1 = 0.12843374;
  double E1 = 1, E2 = 0, E3 = 1, g_recoil0 = 0, DnuHK_Hz = 0;
};

}  // namespace

struct lart_host_model {
  Par par;
  Line line;
  bool is_setup = false;
  // grid arrays
  std::vector<double> xface, yface, zface, rhokap, voigt_a, Dfreq, vfx, vfy, vfz, rhokapD;
  std::vector<int8_t> mask;                                  // grid%mask (spherical atmospheres)
  std::vector<int32_t> ind_sph, ind_cyl, ncount;             // create_JPa_mem: radial bin of a cell, cells per output bin
  int jp_nr = 0;
  size_t jp_bins = 0;                                        // bins of Pa / Pnew; J has nxfreq times as many
  // scattering matrix
  std::vector<double> sm_coss, sm_S11, sm_S12, sm_S33, sm_S34, sm_pdf;
  std::vector<int32_t> sm_alias;
  std::vector<lart_observer> observers;
  // clump population + CSR acceleration grid (clump_mod.f90:30-118)
  std::vector<double> cl_x, cl_y, cl_z, cl_vx, cl_vy, cl_vz, cl_radius, cl_rhokap, cl_rhokapD, cl_voigt_a, cl_Dfreq;
  std::vector<int32_t> cg_start, cg_list;
  std::vector<double> steradian_pix;
  // octree (octree_mod.f90:19-138): leaf input as the generic reader returns it, then the flat tree + leaf physics
  std::vector<double> in_x, in_y, in_z, in_nH, in_T, in_vx, in_vy, in_vz;
  std::vector<int32_t> in_level;
  double amr_boxlen = 0.0, amr_ox = 0.0, amr_oy = 0.0, amr_oz = 0.0;
  std::vector<int32_t> a_parent, a_children, a_level, a_ileaf, a_icell_of_leaf, a_neighbor;
  std::vector<double> a_cx, a_cy, a_cz, a_ch, a_rhokap, a_voigt_a, a_Dfreq, a_vfx, a_vfy, a_vfz, a_rhokapD;
  lart_config cfg{};
  lart_host_summary sum{};
  double dwave = 0.0;
  // tallies
  lart_tallies tal{};
  std::vector<lart_observer_out> obs_out;
  std::vector<std::vector<double>> store;  // owns every tally buffer
  double vtherm_total(double T) const {   // define.f90:928-933
    double vt = line.vtherm1 * std::sqrt(T);
    if (par.bturb > 0.0) vt = std::sqrt(vt * vt + par.bturb * par.bturb);
    return vt;
  }
};

namespace {

bool parse_logical(const std::string &v, bool &out) {
  std::string s = lower(trim(v));
  if (s == ".true." || s == "t" || s == "true" || s == ".t." || s == "1") { out = true; return true; }
  if (s == ".false." || s == "f" || s == "false" || s == ".f." || s == "0") { out = false; return true; }
  return false;
}
bool parse_real(const std::string &v, double &out) {
  std::string s = lower(trim(v));
  for (auto &c : s) if (c == 'd') c = 'e';  // Fortran 1.0d4
  size_t us = s.find('_');                   // 1.0_wp
  if (us != std::string::npos) s = s.substr(0, us);
  char *end = nullptr;
  out = std::strtod(s.c_str(), &end);
  return end != s.c_str() && *end == '\0';
}
bool parse_reals(const std::string &v, std::vector<double> &out) {
  out.clear();
  std::string s = v;
  for (auto &c : s) if (c == ',') c = ' ';
  std::istringstream is(s);
  std::string tok;
  while (is >> tok) {
    double d;
    if (!parse_real(tok, d)) return false;
    out.push_back(d);
  }
  return !out.empty();
}
std::string parse_string(const std::string &v) {
  std::string s = trim(v);
  if (s.size() >= 2 && (s.front() == '\'' || s.front() == '"')) s = s.substr(1, s.size() - 2);
  return s;
}

// Keys of the Fortran namelist that only steer output files / progress printing /
// MPI dealing; they have no effect on the tallies this path returns.
const char *kIgnored[] = {"nprint", "no_print", "out_merge", "out_file", "base_name", "file_format", "out_bitpix",
                          "save_backup", "use_master_slave", "num_send_at_once", "save_sightline_tau", "save_all",
                          "save_input_grid", "luminosity", "save_radial_profile", nullptr};

int set_key(lart_host_model *m, std::string key, const std::string &value) {
  Par &p = m->par;
  key = trim(key);
  std::string lk = lower(key);
  if (lk.rfind("par%", 0) == 0) { key = key.substr(4); lk = lk.substr(4); }
#define REAL(name) if (lk == lower(#name)) { double d; if (!parse_real(value, d)) goto bad; p.name = d; return 0; }
#define INT(name)  if (lk == lower(#name)) { double d; if (!parse_real(value, d)) goto bad; p.name = static_cast<decltype(p.name)>(d); return 0; }
#define BOOL(name) if (lk == lower(#name)) { bool b; if (!parse_logical(value, b)) goto bad; p.name = b; return 0; }
#define STR(name)  if (lk == lower(#name)) { p.name = parse_string(value); return 0; }
#define VEC(name)  if (lk == lower(#name)) { if (!parse_reals(value, p.name)) goto bad; return 0; }
  REAL(no_photons) INT(iseed) REAL(temperature) REAL(temperature0) REAL(bturb) REAL(Dfreq0) REAL(voigt_a0)
  STR(line_id) BOOL(fine_structure)
  REAL(taumax) REAL(tauhomo) REAL(tau0) REAL(N_HImax) REAL(N_HIhomo) REAL(N_HI) REAL(N_gasmax) REAL(N_gashomo)
  REAL(Vexp) REAL(Vx) REAL(Vy) REAL(Vz) REAL(Vrot) REAL(rinner)
  BOOL(use_clump_medium) BOOL(clump_fully_inside) BOOL(clump_allow_overlap) REAL(clump_radius) REAL(clump_N_clumps) REAL(clump_f_vol)
  REAL(clump_f_cov) REAL(clump_tau0) REAL(clump_NHI) REAL(clump_sigma_v)
  BOOL(comoving_source) BOOL(recoil) BOOL(core_skip) BOOL(core_skip_global)
  BOOL(xyz_symmetry) BOOL(xy_symmetry) BOOL(xy_periodic) BOOL(z_symmetry)
  REAL(Omega) BOOL(calc_J) BOOL(calc_P) BOOL(calc_Pnew) BOOL(save_all) INT(geometry_JPa)
  STR(geometry) STR(velocity_type)
  INT(nx) INT(ny) INT(nz) INT(nr)
  REAL(xmax) REAL(ymax) REAL(zmax) REAL(zmin) REAL(rmin) REAL(rmax) REAL(source_rmax)
  REAL(density_rscale) REAL(density_zscale) REAL(density_alpha) REAL(velocity_alpha)
  REAL(xs_point) REAL(ys_point) REAL(zs_point)
  REAL(xfreq0) REAL(xfreq_min) REAL(xfreq_max) INT(nxfreq)
  REAL(velocity_min) REAL(velocity_max) INT(nvelocity)
  REAL(distance2cm) REAL(gaussian_sigma_vel) REAL(gaussian_FWHM_vel)
  STR(distance_unit) STR(source_geometry) STR(spectral_type)
  BOOL(use_reduced_wgt) BOOL(use_stokes) BOOL(save_Jin) BOOL(save_Jabs) BOOL(save_Jmu) BOOL(continuum_normalize) INT(nmu)
  BOOL(save_direc0) BOOL(save_all_photons) BOOL(save_peeloff) BOOL(save_peeloff_2D) BOOL(save_peeloff_3D)
  INT(intensity_unit)
  REAL(hgg) REAL(albedo) REAL(cext_dust) REAL(DGR) STR(scatt_mat_file)
  INT(nxim) INT(nyim) REAL(distance)
  VEC(inclination_angle) VEC(position_angle) VEC(phase_angle) VEC(alpha) VEC(beta) VEC(gamma)
  VEC(obsx) VEC(obsy) VEC(obsz)
  REAL(dxim) REAL(dyim) REAL(rotation_center_x) REAL(rotation_center_y) REAL(rotation_center_z)
#undef REAL
#undef INT
#undef BOOL
#undef STR
#undef VEC
  for (const char **q = kIgnored; *q; ++q) if (lk == *q) return 0;
  g_err = "lart_host_set: unknown or unsupported key '" + key + "'";
  return 1;
bad:
  g_err = "lart_host_set: cannot parse value '" + value + "' for key '" + key + "'";
  return 2;
}

inline double vec_at(const std::vector<double> &v, size_t i) { return i < v.size() ? v[i] : kNaN; }

// random_alias_setup64 — random_mt.f90:2016-2067 (1-based alias, 0 = none)
void alias_setup(std::vector<double> &probs, std::vector<int32_t> &alias) {
  int n = static_cast<int>(probs.size());
  alias.assign(n, 0);
  std::vector<int> small(n, 0), large(n, 0);
  int ns = 0, nl = 0;
  for (int i = 1; i <= n; ++i) {
    probs[i - 1] = n * probs[i - 1];
    if (probs[i - 1] < 1.0) small[ns++] = i; else large[nl++] = i;
  }
  while (ns > 0 && nl > 0) {
    int ks = small[ns - 1], kl = large[nl - 1];
    --ns; --nl;
    alias[ks - 1] = kl;
    probs[kl - 1] = probs[kl - 1] + probs[ks - 1] - 1.0;
    if (probs[kl - 1] < 1.0) small[ns++] = kl; else large[nl++] = kl;
  }
}

// setup_scattering_matrix — setup.f90:581-649
int setup_scattering_matrix(lart_host_model *m, const std::string &path) {
  std::ifstream f(path);
  if (!f) { g_err = "cannot open scatt_mat_file " + path; return 1; }
  std::string l;
  std::getline(f, l);
  double wavelength, cext, albedo, hgg;
  int n;
  if (!(f >> wavelength >> cext >> albedo >> hgg >> n)) { g_err = "bad header in " + path; return 1; }
  std::getline(f, l);
  std::getline(f, l);
  m->par.albedo = albedo; m->par.hgg = hgg; m->par.cext_dust = cext;
  m->sm_coss.resize(n); m->sm_S11.resize(n); m->sm_S12.resize(n); m->sm_S33.resize(n); m->sm_S34.resize(n);
  for (int i = 0; i < n; ++i)
    if (!(f >> m->sm_coss[i] >> m->sm_S11[i] >> m->sm_S12[i] >> m->sm_S33[i] >> m->sm_S34[i])) { g_err = "short table in " + path; return 1; }
  double norm = 0.0;  // calc_Integral — mathlib.f90:218-240
  for (int i = 1; i < n; ++i) norm += 0.5 * (m->sm_S11[i] + m->sm_S11[i - 1]) * std::fabs(m->sm_coss[i] - m->sm_coss[i - 1]);
  for (int i = 0; i < n; ++i) { m->sm_S11[i] /= norm; m->sm_S12[i] /= norm; m->sm_S33[i] /= norm; m->sm_S34[i] /= norm; }
  m->sm_pdf.resize(n - 1);
  double s = 0.0;
  for (int i = 0; i < n - 1; ++i) { m->sm_pdf[i] = (m->sm_S11[i] + m->sm_S11[i + 1]) / 2.0; s += m->sm_pdf[i]; }
  for (auto &v : m->sm_pdf) v /= s;
  alias_setup(m->sm_pdf, m->sm_alias);
  return 0;
}

// read_input's derived defaults — setup.f90:42-562 (the branches this path can reach)
int derive(lart_host_model *m) {
  Par &p = m->par;
  Line &ln = m->line;
  if (p.no_photons >= 1) p.nphotons = static_cast<int64_t>(p.no_photons);  // :42
  p.geometry = lower(p.geometry); p.source_geometry = lower(p.source_geometry);
  p.velocity_type = lower(p.velocity_type); p.distance_unit = lower(p.distance_unit);
  p.spectral_type = lower(p.spectral_type);
  if (p.geometry.empty()) p.geometry = "sphere";   // :70-75
  if (p.geometry == "box") p.geometry = "rectangle";
  if (p.geometry != "sphere" && p.geometry != "rectangle" && p.geometry != "cylinder" && p.geometry != "plane_atmosphere" &&
      p.geometry != "spherical_atmosphere") {
    g_err = "geometry '" + p.geometry + "' stays with the Fortran host (not on the GPU path)"; return 1;
  }
  // exoplanet atmospheres — setup.f90:86-113.  Upstream reads their density from a profile file (read_plane_data /
  // read_spherical_data); this mini-host builds an exponential profile from density_zscale / density_rscale instead.
  if (p.geometry == "plane_atmosphere") {
    p.xy_periodic = true; p.xyz_symmetry = false; p.xy_symmetry = false;
    p.xmax = p.zmax; p.ymax = p.zmax; p.nx = 1; p.ny = 1; p.rmax = -1.0;
  } else if (p.geometry == "spherical_atmosphere") {
    p.xy_periodic = false; p.xyz_symmetry = false;
    if (p.xy_symmetry) {
      p.nx = std::max(p.nx, p.ny); p.ny = p.nx; p.xmax = std::max(p.xmax, p.ymax); p.ymax = p.xmax;
    } else {
      p.nx = std::max({p.nx, p.ny, p.nz}); p.ny = p.nx; p.nz = p.nx;
      p.xmax = std::max({p.xmax, p.ymax, p.zmax}); p.ymax = p.xmax; p.zmax = p.xmax;
    }
    if (p.rmax <= 0.0) p.rmax = p.xmax;
  }
  if (p.Omega != 0.0 && !(p.xy_periodic && p.nx > 1 && p.ny > 1)) { g_err = "par%Omega needs an xy_periodic box with nx, ny > 1"; return 1; }
  // setup_resonance_line, Ly-alpha branch — line_mod.f90:1241-1270
  if (lower(p.line_id) != "ly_alpha" || p.fine_structure) { g_err = "only line_id='ly_alpha' without fine structure (line_type 1) is on this path"; return 1; }
  {
    const double sigma_0 = 0.026540083434, amu = 1.67262192e-24, vtherm1_amu = 0.12895319011972164, mass_amu = 1.00797;
    ln.line_type = 1; ln.DnuHK_Hz = 0.0; ln.E1 = 1.0; ln.E2 = 0.0; ln.E3 = 1.0;
    ln.wavelength0 = 0.1215668237310; ln.damping = 6.2649e8;
    ln.cross0 = sigma_0 / std::sqrt(kPi) * (0.27760 + 0.13881);
    ln.vtherm1 = vtherm1_amu / std::sqrt(mass_amu);
    // NB: amu is in grams here, as in the reference.
    ln.g_recoil0 = (kHPlanck / amu / mass_amu) / ((ln.wavelength0 * kUm2M) * (ln.wavelength0 * kUm2M));
  }
  if (p.spectral_type == "continuum") p.comoving_source = false;  // :114
  if (p.temperature0 <= 0.0) p.temperature0 = p.temperature;       // :127
  if (p.temperature <= 0.0 && p.bturb <= 0.0) { g_err = "par%temperature must be > 0 K (or set par%bturb > 0)"; return 1; }
  if (p.nx == 1 || p.ny == 1 || p.nz == 1) p.xyz_symmetry = false;  // :167
  // par%z_symmetry only changes the grid geometry (grid_mod_car.f90:135-150: z from 0, or from -dz/2, to zmax); setup.f90
  // binds no ray tracer for it, so the plain open-box routines run on the half box — here too
  if (p.z_symmetry && (p.xyz_symmetry || p.xy_symmetry)) p.z_symmetry = false;  // (the if / else-if chain of grid_mod_car.f90:92-135)
  if (p.xyz_symmetry) p.xy_symmetry = false;  // setup.f90:952-957: the xyz variant wins
  if ((p.xyz_symmetry || p.xy_symmetry) && p.xy_periodic) { g_err = "symmetry-folded and xy_periodic grids exclude each other"; return 1; }
  if (p.xy_symmetry && (p.nx == 1 || p.ny == 1)) { g_err = "xy_symmetry needs nx, ny > 1"; return 1; }
  if (!(p.save_peeloff_2D || p.save_peeloff_3D)) p.save_peeloff = false;  // :188
  if (p.nxim > 0 && p.nyim > 0) p.save_peeloff = true;                    // :189
  if (p.save_peeloff && p.xyz_symmetry) p.save_peeloff = false;  // :191-199 peeling-off is not allowed with xyz_symmetry
  if (!p.save_peeloff) { p.save_peeloff_2D = false; p.save_peeloff_3D = false; }
  if (p.tau0 > 0.0 && p.taumax < 0.0) p.taumax = p.tau0;   // :220-223
  if (p.N_HI > 0.0 && p.N_HImax < 0.0) p.N_HImax = p.N_HI;
  if (p.N_HImax > 0.0 && p.N_gasmax < 0.0) p.N_gasmax = p.N_HImax;
  if (p.N_HIhomo > 0.0 && p.N_gashomo < 0.0) p.N_gashomo = p.N_HIhomo;
  if (p.cext_dust <= 0.0) p.DGR = 0.0;  // :225-227
  if (p.DGR == 0.0) p.save_Jabs = false;
  if (p.core_skip_global) p.core_skip = true;
  if (p.save_Jmu) {  // :375-389
    if (p.nmu < 1) { g_err = "par%nmu must be >= 1 when par%save_Jmu = .true."; return 1; }
    if (p.xyz_symmetry) { p.mu_min = 0.0; p.dmu = 1.0 / p.nmu; }  // escapes are folded onto the +z hemisphere
    else { p.mu_min = -1.0; p.dmu = 2.0 / p.nmu; }
  }
  if (p.nr > 1) { p.nx = p.nr; p.ny = p.nr; if (p.geometry != "cylinder") p.nz = p.nr; }  // :392-396
  if (p.geometry == "sphere") {  // :406-418
    if (!p.xy_periodic) {
      double r0 = -1.0;
      for (double v : {p.rmax, p.xmax, p.ymax, p.zmax}) if (v > 0.0) r0 = std::max(r0, v);
      if (r0 > 0.0) { p.rmax = r0; p.xmax = r0; p.ymax = r0; p.zmax = r0; }
      if (!p.xy_symmetry) { p.nx = std::max({p.nx, p.ny, p.nz}); p.ny = p.nx; p.nz = p.nx; }
    }
  } else if (p.geometry == "cylinder") {
    double r0 = -1.0;
    for (double v : {p.rmax, p.xmax, p.ymax}) if (v > 0.0) r0 = std::max(r0, v);
    if (r0 > 0.0) { p.rmax = r0; p.xmax = r0; p.ymax = r0; }
    if (!p.xy_symmetry) { p.nx = std::max(p.nx, p.ny); p.ny = p.nx; }
  } else if (p.geometry == "rectangle") {
    p.rmax = -1.0;
  }
  if (p.source_rmax < 0.0) p.source_rmax = p.rmax;  // :432
  if (p.distance2cm < 0.0) {  // :474-489
    const double kpc2cm = 3.0856775814913673e21, pc2cm = 3.0856775814913673e18, au2cm = 1.495978707e13;
    if (p.distance_unit == "kpc") p.distance2cm = kpc2cm;
    else if (p.distance_unit == "pc") p.distance2cm = pc2cm;
    else if (p.distance_unit == "au") p.distance2cm = au2cm;
    else if (p.distance_unit.empty()) p.distance2cm = 1.0;
    else p.distance2cm = kpc2cm;
  } else {
    p.distance_unit = "user";
  }
  if (p.intensity_unit < 0) p.intensity_unit = (p.distance2cm != 1.0) ? 1 : 0;  // :492-498
  if (!p.scatt_mat_file.empty() && p.DGR > 0.0) {  // :500-504
    if (int rc = setup_scattering_matrix(m, p.scatt_mat_file)) return rc;
  } else if (p.use_stokes) {
    p.DGR = 0.0;
  }
  if (p.DGR == 0.0) p.save_Jabs = false;
  if (p.source_geometry != "point" && p.source_geometry != "uniform" && p.source_geometry != "uniform_sphere" && p.source_geometry != "sphere" &&
      p.source_geometry != "plane_illumination") {
    g_err = "source_geometry '" + p.source_geometry + "' stays with the Fortran host"; return 1;
  }
  if (p.source_geometry == "plane_illumination" && p.geometry != "plane_atmosphere" && p.geometry != "spherical_atmosphere") {
    g_err = "source_geometry 'plane_illumination' needs an atmosphere geometry (generate_photon.f90:742-778)"; return 1;
  }
  if (p.calc_J || p.calc_P || p.calc_Pnew) {  // setup.f90:438-455
    if (p.use_clump_medium || p.use_amr_grid) { g_err = "CALCJ/CALCP accumulators: Cartesian grids only on this path"; return 1; }
    if (p.rmax <= 0.0) p.geometry_JPa = 3;
    if (p.geometry_JPa < -1 || p.geometry_JPa > 3) {
      p.geometry_JPa = 1;
      if (p.save_all) p.geometry_JPa = 3;
      if (p.geometry_JPa != 3) {
        if (p.xy_periodic) p.geometry_JPa = -1;
        else if (p.xyz_symmetry) p.geometry_JPa = 1;
        else if (p.xy_symmetry) p.geometry_JPa = 2;
        else if (p.geometry == "cylinder") p.geometry_JPa = 2;
        else if (p.geometry == "spherical_atmosphere") p.geometry_JPa = 2;
      }
    }
    if (p.geometry_JPa == 0) { g_err = "par%geometry_JPa = 0 is not a geometry"; return 1; }
  }
  static const char *spec_ok[] = {"voigt", "voigt0", "continuum", "gaussian", "monochromatic", "mono", nullptr};
  bool ok = false;
  for (const char **q = spec_ok; *q; ++q) ok = ok || p.spectral_type == *q;
  if (!ok) { g_err = "spectral_type '" + p.spectral_type + "' stays with the Fortran host"; return 1; }
  return 0;
}

// init_clumps + generate_clumps + assign_clump_velocities_from_type + build_clump_csr + compute_clump_scalars —
// clump_mod.f90:646-895, 897-1150, 1153-1262, 1267-1349, 2316-2380.  The generated, uniform population only: no
// radial profiles, no bicone, no clump file, no overlap.  Positions come from MT19937-64 seeded with par%iseed
// (upstream uses rank 0's stream of the same generator, so the layout is "a" valid layout, not upstream's).
int clumps_create(lart_host_model *m) {
  Par &p = m->par;
  const Line &ln = m->line;
  lart_clumps &c = m->cfg.clumps;
  c = lart_clumps{};
  if (!p.use_clump_medium) return 0;
  if (p.rmax <= 0.0) { g_err = "par%rmax must be > 0 for clump medium"; return 1; }  // grid_mod_clump.f90:42-45
  p.xmax = p.ymax = p.zmax = p.rmax;
  if (p.nx < 1) p.nx = 11;
  if (p.ny < 1) p.ny = 11;
  if (p.nz < 1) p.nz = 11;
  p.xyz_symmetry = p.xy_symmetry = p.xy_periodic = p.z_symmetry = false;
  const double R = p.rmax, rcl = p.clump_radius;
  if (rcl <= 0.0) { g_err = "clump_radius must be > 0"; return 1; }
  const double r0 = std::max(0.0, p.rmin);
  if (r0 >= R) { g_err = "par%rmin must be < par%rmax for clump placement"; return 1; }
  const double vtherm = m->vtherm_total(p.temperature);
  const double Dref = vtherm / (ln.wavelength0 * kUm2Km), aref = (ln.damping / kFourPi) / Dref;
  auto voigt0 = [](double a) { return 1.0 + a * (-1.1283791671e+00 + a * 1.0); };
  const double shell2 = R * R + R * r0 + r0 * r0;
  int64_t N;
  if (p.clump_N_clumps > 0.0) N = static_cast<int64_t>(p.clump_N_clumps);  // :716-732
  else if (p.clump_f_vol > 0.0) N = std::llround(p.clump_f_vol * (R * R * R - r0 * r0 * r0) / (rcl * rcl * rcl));
  else if (p.clump_f_cov > 0.0) N = std::llround((4.0 / 3.0) * p.clump_f_cov * shell2 / (rcl * rcl));
  else { g_err = "specify clump_N_clumps, clump_f_vol, or clump_f_cov"; return 1; }
  if (N <= 0) N = 1;
  if (N > 200000000) { g_err = "too many clumps for the mini-host"; return 1; }
  double kap;  // :765-809
  if (p.clump_tau0 > 0.0) kap = p.clump_tau0 / (voigt0(aref) * rcl);
  else if (p.clump_NHI > 0.0) kap = p.clump_NHI * ln.cross0 / (Dref * rcl);
  else if (p.taumax > 0.0 || p.N_HImax > 0.0 || p.N_gasmax > 0.0) {
    const double GF = static_cast<double>(N) * rcl * rcl * rcl / std::max(shell2, 2.2250738585072014e-308);
    if (p.taumax > 0.0) kap = p.taumax / (GF * voigt0(aref));
    else kap = std::max(p.N_HImax, p.N_gasmax) * ln.cross0 / (GF * Dref);
  } else { g_err = "specify clump_tau0, clump_NHI, taumax, or N_HImax"; return 1; }
  const size_t n = static_cast<size_t>(N);
  m->cl_x.assign(n, 0.0); m->cl_y.assign(n, 0.0); m->cl_z.assign(n, 0.0);
  m->cl_vx.assign(n, 0.0); m->cl_vy.assign(n, 0.0); m->cl_vz.assign(n, 0.0);
  m->cl_radius.assign(n, rcl); m->cl_rhokap.assign(n, kap); m->cl_voigt_a.assign(n, aref); m->cl_Dfreq.assign(n, Dref);
  if (p.DGR > 0.0) m->cl_rhokapD.assign(n, kap * p.cext_dust * p.DGR * Dref / ln.cross0); else m->cl_rhokapD.clear();  // :859-860
  // ---- generate_clumps: random sequential addition with a linked-list grid (:897-1150, uniform path)
  std::mt19937_64 gen(static_cast<uint64_t>(p.iseed));
  auto rnd = [&]() { return (static_cast<double>(gen() >> 12) + 0.5) * (1.0 / 4503599627370496.0); };  // random_mt.f90:579-630
  bool have_g = false; double gset = 0.0;
  auto gauss = [&]() {  // rand_gauss :964-988
    if (have_g) { have_g = false; return gset; }
    double v1, v2, rsq;
    do { v1 = 2.0 * rnd() - 1.0; v2 = 2.0 * rnd() - 1.0; rsq = v1 * v1 + v2 * v2; } while (rsq >= 1.0 || rsq == 0.0);
    rsq = std::sqrt(-2.0 * std::log(rsq) / rsq);
    gset = v1 * rsq; have_g = true;
    return v2 * rsq;
  };
  int rg = std::min(512, std::max(32, static_cast<int>(std::cbrt(static_cast<double>(N))) + 1));
  double rg_cell = std::max(2.0 * R / rg, 2.0 * rcl);
  rg = std::max(2, static_cast<int>((2.0 * R) / rg_cell) + 1);
  rg_cell = (2.0 * R) / rg;
  const double min_sep2 = (2.0 * rcl) * (2.0 * rcl);
  if (p.clump_fully_inside && (rcl >= R || r0 + 2.0 * rcl > R)) { g_err = "clump_fully_inside: no clump fits inside the shell"; return 1; }
  const double rmaxc = p.clump_fully_inside ? R - rcl : R, rminc = p.clump_fully_inside ? r0 + rcl : r0;
  std::vector<int> head(static_cast<size_t>(rg) * rg * rg, -1), nxt(n, -1);
  int64_t icl = 0, attempts = 0;
  while (icl < N) {
    if (++attempts > 2000 * N + 1000000) { g_err = "clump placement does not converge (filling factor too high)"; return 1; }
    double xc, yc, zc, d2;
    do {
      xc = (2.0 * rnd() - 1.0) * rmaxc; yc = (2.0 * rnd() - 1.0) * rmaxc; zc = (2.0 * rnd() - 1.0) * rmaxc;
      d2 = xc * xc + yc * yc + zc * zc;
    } while (!(d2 <= rmaxc * rmaxc && d2 >= rminc * rminc));
    auto gi = [&](double v) { return std::min(rg - 1, std::max(0, static_cast<int>((v + R) / rg_cell))); };
    const int ig = gi(xc), jg = gi(yc), kg = gi(zc);
    bool overlap = false;
    for (int k2 = std::max(0, kg - 1); k2 <= std::min(rg - 1, kg + 1) && !overlap; ++k2)
      for (int j2 = std::max(0, jg - 1); j2 <= std::min(rg - 1, jg + 1) && !overlap; ++j2)
        for (int i2 = std::max(0, ig - 1); i2 <= std::min(rg - 1, ig + 1) && !overlap; ++i2)
          for (int jn = head[i2 + static_cast<size_t>(rg) * (j2 + static_cast<size_t>(rg) * k2)]; jn >= 0; jn = nxt[jn]) {
            double ddx = xc - m->cl_x[jn], ddy = yc - m->cl_y[jn], ddz = zc - m->cl_z[jn];
            if (ddx * ddx + ddy * ddy + ddz * ddz < min_sep2) { overlap = true; break; }
          }
    if (overlap && !p.clump_allow_overlap) continue;  // :1096
    m->cl_x[icl] = xc; m->cl_y[icl] = yc; m->cl_z[icl] = zc;
    if (p.clump_sigma_v > 0.0) {
      m->cl_vx[icl] = p.clump_sigma_v / vtherm * gauss();
      m->cl_vy[icl] = p.clump_sigma_v / vtherm * gauss();
      m->cl_vz[icl] = p.clump_sigma_v / vtherm * gauss();
    }
    size_t cell = ig + static_cast<size_t>(rg) * (jg + static_cast<size_t>(rg) * kg);
    nxt[icl] = head[cell]; head[cell] = static_cast<int>(icl);
    ++icl;
  }
  // ---- assign_clump_velocities_from_type (:1153-1262)
  const std::string &vt = p.velocity_type;
  if (!vt.empty()) {
    if (vt != "hubble" && vt != "constant_radial" && vt != "power_law" && vt != "parallel_velocity" &&
        vt != "rotating_solid_body" && vt != "rotating_galaxy_halo") {
      g_err = "velocity_type '" + vt + "' stays with the Fortran host"; return 1;
    }
    for (size_t i = 0; i < n; ++i) {
      const double xc = m->cl_x[i], yc = m->cl_y[i], zc = m->cl_z[i], rr = std::sqrt(xc * xc + yc * yc + zc * zc);
      double vx = 0.0, vy = 0.0, vz = 0.0;
      if (vt == "hubble") { vx = p.Vexp * xc / R; vy = p.Vexp * yc / R; vz = p.Vexp * zc / R; }
      else if (vt == "constant_radial") { if (rr > 0.0) { vx = p.Vexp * xc / rr; vy = p.Vexp * yc / rr; vz = p.Vexp * zc / rr; } }
      else if (vt == "power_law") { if (rr > 0.0) { double V = p.Vexp * std::pow(rr / R, p.velocity_alpha); vx = V * xc / rr; vy = V * yc / rr; vz = V * zc / rr; } }
      else if (vt == "parallel_velocity") { vx = p.Vx; vy = p.Vy; vz = p.Vz; }
      else if (vt == "rotating_solid_body") { vx = -p.Vrot * yc / R; vy = p.Vrot * xc / R; }
      else {  // rotating_galaxy_halo: flat rotation curve
        const double rc = std::sqrt(xc * xc + yc * yc);
        if (rc > 0.0) {
          if (rc < p.rinner) { vx = -p.Vrot * yc / p.rinner; vy = p.Vrot * xc / p.rinner; }
          else { vx = -p.Vrot * yc / rc; vy = p.Vrot * xc / rc; }
        }
      }
      m->cl_vx[i] += vx / vtherm; m->cl_vy[i] += vy / vtherm; m->cl_vz[i] += vz / vtherm;
    }
  }
  // ---- build_clump_csr (:1267-1349) with clump_cell_range (:1352-1366)
  const int cg = std::min(512, std::max(32, static_cast<int>(std::cbrt(static_cast<double>(N))) + 1));
  const double cmin = -(R + rcl), cd = (2.0 * (R + rcl)) / cg, cinv = 1.0 / cd;
  const size_t ncells = static_cast<size_t>(cg) * cg * cg;
  auto range = [&](double v, int &lo, int &hi) {
    lo = std::max(0, static_cast<int>((v - cmin - rcl) * cinv));
    hi = std::min(cg - 1, static_cast<int>((v - cmin + rcl) * cinv));
  };
  std::vector<int32_t> cnt(ncells, 0);
  for (size_t i = 0; i < n; ++i) {
    int i0, i1, j0, j1, k0, k1;
    range(m->cl_x[i], i0, i1); range(m->cl_y[i], j0, j1); range(m->cl_z[i], k0, k1);
    for (int k = k0; k <= k1; ++k) for (int j = j0; j <= j1; ++j) for (int ii = i0; ii <= i1; ++ii) ++cnt[ii + static_cast<size_t>(cg) * (j + static_cast<size_t>(cg) * k)];
  }
  m->cg_start.assign(ncells + 1, 0);
  m->cg_start[0] = 1;  // 1-based offsets, as upstream
  for (size_t q = 0; q < ncells; ++q) m->cg_start[q + 1] = m->cg_start[q] + cnt[q];
  m->cg_list.assign(static_cast<size_t>(m->cg_start[ncells] - 1), 0);
  std::fill(cnt.begin(), cnt.end(), 0);
  for (size_t i = 0; i < n; ++i) {
    int i0, i1, j0, j1, k0, k1;
    range(m->cl_x[i], i0, i1); range(m->cl_y[i], j0, j1); range(m->cl_z[i], k0, k1);
    for (int k = k0; k <= k1; ++k) for (int j = j0; j <= j1; ++j) for (int ii = i0; ii <= i1; ++ii) {
      size_t q = ii + static_cast<size_t>(cg) * (j + static_cast<size_t>(cg) * k);
      m->cg_list[static_cast<size_t>(m->cg_start[q] - 1) + cnt[q]] = static_cast<int32_t>(i + 1);
      ++cnt[q];
    }
  }
  // ---- compute_clump_scalars (:2316-2380)
  double v1 = 0, v2 = 0, v3 = 0, v4 = 0;
  for (size_t i = 0; i < n; ++i) {
    const double r3 = rcl * rcl * rcl;
    const double di2 = std::max(m->cl_x[i] * m->cl_x[i] + m->cl_y[i] * m->cl_y[i] + m->cl_z[i] * m->cl_z[i], rcl * rcl);
    const double wh = r3 / shell2, wm = r3 / (3.0 * di2);
    v1 += kap * voigt0(aref) * wh; v2 += kap * voigt0(aref) * wm;
    v3 += kap * Dref * wh / ln.cross0; v4 += kap * Dref * wm / ln.cross0;
  }
  p.tauhomo = v1; p.taumax = v2; p.N_gashomo = v3; p.N_gasmax = v4;
  c.n = N; c.sphere_R = R; c.Dfreq_ref = Dref;
  c.x = m->cl_x.data(); c.y = m->cl_y.data(); c.z = m->cl_z.data();
  c.vx = m->cl_vx.data(); c.vy = m->cl_vy.data(); c.vz = m->cl_vz.data();
  c.radius = m->cl_radius.data(); c.rhokap = m->cl_rhokap.data();
  c.rhokapD = m->cl_rhokapD.empty() ? nullptr : m->cl_rhokapD.data();
  c.voigt_a = m->cl_voigt_a.data(); c.Dfreq = m->cl_Dfreq.data();
  c.cgx = c.cgy = c.cgz = cg; c.has_overlap = 0;
  if (p.clump_allow_overlap) {  // check_has_overlap (:1544-1590): any pair closer than the sum of its radii
    for (size_t i = 0; i < n && !c.has_overlap; ++i) {
      int i0, i1, j0, j1, k0, k1;
      range(m->cl_x[i], i0, i1); range(m->cl_y[i], j0, j1); range(m->cl_z[i], k0, k1);
      for (int k = k0; k <= k1 && !c.has_overlap; ++k) for (int j = j0; j <= j1 && !c.has_overlap; ++j) for (int ii = i0; ii <= i1; ++ii) {
        size_t q = ii + static_cast<size_t>(cg) * (j + static_cast<size_t>(cg) * k);
        for (int32_t ip = m->cg_start[q]; ip < m->cg_start[q + 1]; ++ip) {
          const size_t jn = static_cast<size_t>(m->cg_list[ip - 1]) - 1;
          if (jn <= i) continue;
          const double ddx = m->cl_x[jn] - m->cl_x[i], ddy = m->cl_y[jn] - m->cl_y[i], ddz = m->cl_z[jn] - m->cl_z[i];
          if (ddx * ddx + ddy * ddy + ddz * ddz < (2.0 * rcl) * (2.0 * rcl)) { c.has_overlap = 1; break; }
        }
        if (c.has_overlap) break;
      }
    }
  }
  c.cg_xmin = c.cg_ymin = c.cg_zmin = cmin; c.cg_dx = c.cg_dy = c.cg_dz = cd;
  c.cg_start = m->cg_start.data(); c.cg_list = m->cg_list.data();
  return 0;
}

// grid_create — grid_mod_car.f90:11-1238 (synthetic branch: no dens/temp/velo files)
int grid_create(lart_host_model *m) {
  Par &p = m->par;
  const Line &ln = m->line;
  const int nx = p.nx, ny = p.ny, nz = p.nz;
  const size_t nc = static_cast<size_t>(nx) * ny * nz;
  // :92-118 (xyz symmetry: one octant, mirror planes at the lower faces) / :168-176 (no symmetry)
  double dx, dy, dz, xmin, ymin, zmin;
  int i0 = 0, j0 = 0, k0 = 0;
  auto axis = [](int n, double vmax, double &d, double &vmin, int &c0) {
    if ((n / 2) * 2 == n) { d = vmax / n; vmin = 0.0; c0 = 1; }
    else { d = vmax / (n - 0.5); vmin = -d / 2.0; c0 = 2; }
  };
  if (p.xyz_symmetry) {
    axis(nx, p.xmax, dx, xmin, i0); axis(ny, p.ymax, dy, ymin, j0); axis(nz, p.zmax, dz, zmin, k0);
  } else if (p.xy_symmetry) {  // :113-134 — a quadrant in x,y; the full height in z
    axis(nx, p.xmax, dx, xmin, i0); axis(ny, p.ymax, dy, ymin, j0);
    dz = 2.0 * p.zmax / nz; zmin = -p.zmax; k0 = 0;
  } else if (p.z_symmetry) {  // :135-150 — the upper half in z, the full extent in x and y; k0 is set but never read
    dx = 2.0 * p.xmax / nx; dy = 2.0 * p.ymax / ny; xmin = -p.xmax; ymin = -p.ymax;
    axis(nz, p.zmax, dz, zmin, k0);
  } else if (p.geometry == "plane_atmosphere") {  // :151-166 — the column runs from par%zmin (default 0) to zmax
    dx = 2.0 * p.xmax / nx; dy = 2.0 * p.ymax / ny; xmin = -p.xmax; ymin = -p.ymax;
    zmin = isfin(p.zmin) ? p.zmin : 0.0;
    dz = (p.zmax - zmin) / nz;
  } else {
    dx = 2.0 * p.xmax / nx; dy = 2.0 * p.ymax / ny; dz = 2.0 * p.zmax / nz;
    xmin = -p.xmax; ymin = -p.ymax; zmin = -p.zmax;
  }
  m->xface.resize(nx + 1); m->yface.resize(ny + 1); m->zface.resize(nz + 1);
  for (int i = 1; i <= nx + 1; ++i) m->xface[i - 1] = (i - 1) * dx + xmin;  // :188-190
  for (int j = 1; j <= ny + 1; ++j) m->yface[j - 1] = (j - 1) * dy + ymin;
  for (int k = 1; k <= nz + 1; ++k) m->zface[k - 1] = (k - 1) * dz + zmin;
  std::vector<double> xx(nx), yy(ny), zz(nz);
  for (int i = 0; i < nx; ++i) xx[i] = (m->xface[i] + m->xface[i + 1]) / 2.0;  // :232-240
  for (int j = 0; j < ny; ++j) yy[j] = (m->yface[j] + m->yface[j + 1]) / 2.0;
  for (int k = 0; k < nz; ++k) zz[k] = (m->zface[k] + m->zface[k + 1]) / 2.0;
  const double Dfreq_ref = m->vtherm_total(p.temperature) / (ln.wavelength0 * kUm2Km);  // :245
  m->Dfreq.assign(nc, 0.0); m->voigt_a.assign(nc, 0.0); m->rhokap.assign(nc, 0.0);
  m->vfx.assign(nc, 0.0); m->vfy.assign(nc, 0.0); m->vfz.assign(nc, 0.0);
  const bool dust = p.DGR > 0.0;
  if (dust) m->rhokapD.assign(nc, 0.0); else m->rhokapD.clear();
  // (1) uniform temperature :270-284
  {
    double vt = m->vtherm_total(p.temperature);
    double D = vt / (ln.wavelength0 * kUm2Km), a = (ln.damping / kFourPi) / D;
    std::fill(m->Dfreq.begin(), m->Dfreq.end(), D);
    std::fill(m->voigt_a.begin(), m->voigt_a.end(), a);
  }
  // (2) density :355-366 — synthetic test resets the distance unit
  if (!p.use_clump_medium) { p.distance_unit = ""; p.distance2cm = 1.0; }
  std::fill(m->rhokap.begin(), m->rhokap.end(), 1.0);
  if (dust) std::fill(m->rhokapD.begin(), m->rhokapD.end(), p.cext_dust * p.DGR);
  auto at = [&](int i, int j, int k) { return static_cast<size_t>(i) + static_cast<size_t>(nx) * (j + static_cast<size_t>(ny) * k); };
  const bool cyl = p.geometry == "cylinder";
  auto radius = [&](int i, int j, int k) {
    return cyl ? std::sqrt(xx[i] * xx[i] + yy[j] * yy[j]) : std::sqrt(xx[i] * xx[i] + yy[j] * yy[j] + zz[k] * zz[k]);
  };
  if (p.rmax > 0.0) {  // :368-394
    for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
      double rr = radius(i, j, k);
      if (rr < p.rmin || rr > p.rmax) { m->rhokap[at(i, j, k)] = 0.0; if (dust) m->rhokapD[at(i, j, k)] = 0.0; }
    }
  }
  if (p.density_rscale > 0.0)  // :414-428
    for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
      double f = std::exp(-radius(i, j, k) / p.density_rscale);
      m->rhokap[at(i, j, k)] *= f; if (dust) m->rhokapD[at(i, j, k)] *= f;
    }
  if (p.density_zscale > 0.0)  // :430-439
    for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
      double f = std::exp(-std::fabs(zz[k]) / p.density_zscale);
      m->rhokap[at(i, j, k)] *= f; if (dust) m->rhokapD[at(i, j, k)] *= f;
    }
  if (p.density_alpha != 0.0) {  // :445-466
    double rpk = (p.rmax <= 0.0) ? std::max({p.xmax, p.ymax, p.zmax}) : p.rmax;
    for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
      double rr = radius(i, j, k);
      if (rr > 0.0) { double f = std::pow(rpk / rr, p.density_alpha); m->rhokap[at(i, j, k)] *= f; if (dust) m->rhokapD[at(i, j, k)] *= f; }
    }
  }
  for (size_t c = 0; c < nc; ++c) m->rhokap[c] = m->rhokap[c] / m->Dfreq[c] * ln.cross0;  // :487-493
  double opac_length;  // :495-503
  if (p.rmax > 0.0 && p.rmin > 0.0) opac_length = p.rmax - p.rmin;
  else if (p.rmax > 0.0) opac_length = p.rmax;
  else if (p.zmax == -zmin) opac_length = (p.zmax - zmin) / 2.0;
  else opac_length = p.zmax - zmin;
  const bool sym = p.xyz_symmetry, symxy = p.xyz_symmetry || p.xy_symmetry;
  const bool zodd = (nz / 2) * 2 != nz;
  const int nxcen = symxy ? 1 : (nx + 1) / 2, nycen = symxy ? 1 : (ny + 1) / 2;  // :505-515 (1-based)
  // voigt(0,a): |x|<1 branch of voigt_seon2 at x=0 is h0(1)+a*(h1(1)+a*h2(1))
  // (voigt_mod.f90:691-700; h0(1)=1, h1(1)=-1.1283791671, h2(1)=1).
  auto voigt0 = [](double a) { return 1.0 + a * (-1.1283791671e+00 + a * 1.0); };
  auto scale_all = [&](double f) { for (auto &v : m->rhokap) v *= f; if (dust) for (auto &v : m->rhokapD) v *= f; };
  auto homo_sum = [&](auto weight, double &nopac) {  // cells cut by a mirror plane count half (:545-556)
    double s = 0.0; nopac = 0.0;
    for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
      size_t c = at(i, j, k);
      if (!(m->rhokap[c] > 0.0)) continue;
      double nadd = 1.0;
      if (symxy) {
        if (i == 0 && (nx / 2) * 2 != nx) nadd /= 2.0;
        if (j == 0 && (ny / 2) * 2 != ny) nadd /= 2.0;
        if (sym && k == 0 && zodd) nadd /= 2.0;
      }
      s += weight(c) * nadd; nopac += nadd;
    }
    return s;
  };
  // pole integral -> length factor (:524-537, :696-704): full box 2 tau = sum*dz; octant tau = (sum - first/2)*dz
  auto pole = [&](auto weight) {
    double s = 0.0;
    for (int k = 0; k < nz; ++k) s += weight(at(nxcen - 1, nycen - 1, k));
    if (sym) return (zodd ? s - weight(at(nxcen - 1, nycen - 1, 0)) / 2.0 : s) * dz;
    return (p.zmax == -zmin) ? s * dz / 2.0 : s * dz;
  };
  double nopac;
  if (p.taumax > 0.0) {  // :518-538
    scale_all(p.taumax / pole([&](size_t c) { return m->rhokap[c] * voigt0(m->voigt_a[c]); }));
  } else if (p.tauhomo > 0.0) {  // :539-566
    double s = homo_sum([&](size_t c) { return m->rhokap[c] * voigt0(m->voigt_a[c]); }, nopac);
    scale_all(p.tauhomo / (s / nopac * opac_length));
  } else if (p.N_gasmax > 0.0) {  // :567-587
    scale_all(p.N_gasmax / (pole([&](size_t c) { return m->rhokap[c] * m->Dfreq[c]; }) / ln.cross0));
  } else if (p.N_gashomo > 0.0) {  // :588-615
    double s = homo_sum([&](size_t c) { return m->rhokap[c] * m->Dfreq[c]; }, nopac);
    scale_all(p.N_gashomo / ((s / nopac / ln.cross0) * opac_length));
  }
  // diagnostics :617-743
  double s = homo_sum([&](size_t c) { return m->rhokap[c] * voigt0(m->voigt_a[c]); }, nopac);
  double tauhomo = s / nopac * opac_length;
  double taupole = pole([&](size_t c) { return m->rhokap[c] * voigt0(m->voigt_a[c]); });
  s = homo_sum([&](size_t c) { return m->rhokap[c] * m->Dfreq[c]; }, nopac);
  double N_gashomo = s / nopac / ln.cross0 * opac_length;
  double N_gaspole = pole([&](size_t c) { return m->rhokap[c] * m->Dfreq[c]; }) / ln.cross0;
  double tauhomo_dust = 0.0, taupole_dust = 0.0;
  if (dust) {
    s = homo_sum([&](size_t c) { return m->rhokapD[c]; }, nopac);
    tauhomo_dust = s / nopac * opac_length;
    taupole_dust = pole([&](size_t c) { return m->rhokapD[c]; });
  }
  if (p.taumax <= 0.0) p.taumax = taupole;  // :744-747
  if (p.tauhomo <= 0.0) p.tauhomo = tauhomo;
  if (p.N_gasmax <= 0.0) p.N_gasmax = N_gaspole;
  if (p.N_gashomo <= 0.0) p.N_gashomo = N_gashomo;
  // (3) velocity field :786-920 — analytic types only, assigned where rhokap > 0
  const std::string &vt = p.velocity_type;
  if (p.use_clump_medium) {
    // the box of a clump medium carries no bulk velocity (grid_mod_clump.f90:97-99); velocity_type belongs to the clumps
  } else if (vt == "hubble" || vt == "power_law" || vt == "constant_radial" || vt == "parallel_velocity") {
    p.rpeak = (p.rmax <= 0.0) ? std::max({p.xmax, p.ymax, p.zmax}) : p.rmax;
    const double vth = m->vtherm_total(p.temperature);
    for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
      size_t c = at(i, j, k);
      if (!(m->rhokap[c] > 0.0)) continue;
      if (vt == "hubble") {  // :786-803
        m->vfx[c] = (p.Vexp / vth) * xx[i] / p.rpeak;
        m->vfy[c] = (p.Vexp / vth) * yy[j] / p.rpeak;
        m->vfz[c] = (p.Vexp / vth) * zz[k] / p.rpeak;
      } else if (vt == "parallel_velocity") {  // :804-816
        m->vfx[c] = p.Vx / vth; m->vfy[c] = p.Vy / vth; m->vfz[c] = p.Vz / vth;
      } else {
        double rr = std::sqrt(xx[i] * xx[i] + yy[j] * yy[j] + zz[k] * zz[k]);
        if (!(rr > dz / 10.0)) continue;
        double V = (vt == "constant_radial") ? p.Vexp : p.Vexp * std::pow(rr / p.rpeak, p.velocity_alpha);  // :840-891
        m->vfx[c] = V / vth * xx[i] / rr; m->vfy[c] = V / vth * yy[j] / rr; m->vfz[c] = V / vth * zz[k] / rr;
      }
    }
  } else if (!vt.empty()) {
    g_err = "velocity_type '" + vt + "' stays with the Fortran host"; return 1;
  }
  // :1149-1155 + car_setup_freq_grid :1442-1511
  const double voigt_amean = (ln.damping / kFourPi) / Dfreq_ref;
  const double atau0 = voigt_amean * p.tauhomo;
  const double atau3 = std::pow(voigt_amean * p.tauhomo, 1.0 / 3.0);
  p.atau3 = atau3;
  const double vtherm = m->vtherm_total(p.temperature);
  if (isfin(p.velocity_min) && isfin(p.velocity_max)) {
    if (p.nvelocity == 0 && p.nxfreq > 0) p.nvelocity = p.nxfreq;
    if (p.nvelocity > 0) p.nxfreq = p.nvelocity;
    p.xfreq_min = -p.velocity_max / vtherm; p.xfreq_max = -p.velocity_min / vtherm;
  }
  if (!(isfin(p.xfreq_max) && isfin(p.xfreq_min))) {
    double xscale = (p.taumax <= 5e1) ? 25.0 : (p.taumax <= 5e2) ? 14.0 : (p.taumax <= 5e3) ? 10.0 : 5.0;
    double hk = ln.DnuHK_Hz / Dfreq_ref;
    if (p.Vexp == 0.0) {
      p.xfreq_max = std::floor(xscale * atau3) + 1; p.xfreq_min = -(std::floor(xscale * atau3 + hk) + 1);
    } else if (p.Vexp > 0.0) {
      p.xfreq_max = std::floor(xscale * atau3) + 1; p.xfreq_min = -(std::floor(xscale * atau3 + std::fabs(p.Vexp) / vtherm + hk) + 1);
    } else {
      p.xfreq_max = std::floor(xscale * atau3 + std::fabs(p.Vexp) / vtherm) + 1; p.xfreq_min = -(std::floor(xscale * atau3 + hk) + 1);
    }
    if (p.spectral_type == "continuum") {
      xscale = 4.0 * xscale;
      p.xfreq_max = std::floor(xscale * atau3 + std::fabs(p.Vexp) / vtherm) + 1;
      p.xfreq_min = -(std::floor(xscale * atau3 + std::fabs(p.Vexp) / vtherm + hk) + 1);
    }
  }
  const double dxfreq = (p.xfreq_max - p.xfreq_min) / p.nxfreq;
  m->dwave = vtherm / kSpeedC * (ln.wavelength0 * 1e4) * dxfreq;
  // global core-skip parameters :1186-1219
  double xcrit = 0.0, xcrit2 = 0.0;
  {
    double at0 = p.core_skip_global ? atau0 : atau0 / (p.xmax / dx);
    if (at0 > 1.0) {
      double xi = (at0 <= 60.0) ? 0.6 : 1.4, chi = (at0 <= 60.0) ? 1.2 : 0.6;
      xcrit = 0.02 * std::exp(xi * std::pow(std::log(at0), chi));
      xcrit2 = xcrit * xcrit;
    }
  }
  // grid%mask — grid_mod_car.f90:247-250, 320-330: cells whose centre lies within rmin belong to the planet
  m->mask.clear();
  if (p.geometry == "spherical_atmosphere") {
    m->mask.assign(nc, 0);
    for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i)
      if (std::sqrt(xx[i] * xx[i] + yy[j] * yy[j] + zz[k] * zz[k]) <= p.rmin) m->mask[at(i, j, k)] = -1;
  }
  // create_JPa_mem — grid_mod_car.f90:1242-1440: radial bin of every cell and the number of cells per output bin
  m->ind_sph.clear(); m->ind_cyl.clear(); m->ncount.clear(); m->jp_nr = 0; m->jp_bins = 0;
  if (p.calc_J || p.calc_P || p.calc_Pnew) {
    const int gj = p.geometry_JPa;
    int nr = 0;
    double rmaxJ = 0.0, dr = 0.0, roff = 0.0;
    if (gj == 2 || gj == 1) {
      nr = (gj == 2) ? std::max(nx, ny) : std::max({nx, ny, nz});
      const bool folded = (gj == 2) ? p.xy_symmetry : p.xyz_symmetry;
      if (!folded) nr = ((nr / 2) * 2 == nr) ? nr / 2 : (nr - 1) / 2 + 1;
      rmaxJ = (gj == 2) ? std::min(p.xmax, p.ymax) : std::min({p.xmax, p.ymax, p.zmax});
      if ((nr / 2) * 2 == nr) { dr = rmaxJ / nr; roff = 0.0; }   // :1281-1289 ("minor-bug fixed (2021.05.26)")
      else { dr = rmaxJ / (nr - 0.5); roff = -dr / 2.0; }
    }
    auto nadd_of = [&](int i, int j, int k, bool with_z) {
      int nadd = with_z ? 8 : 4;
      if (i == 0 && (nx / 2) * 2 != nx) nadd /= 2;
      if (j == 0 && (ny / 2) * 2 != ny) nadd /= 2;
      if (with_z && k == 0 && (nz / 2) * 2 != nz) nadd /= 2;
      return nadd;
    };
    if (gj == 1) {
      m->ind_sph.assign(nc, 0); m->ncount.assign(nr, 0); m->jp_bins = nr;
      for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
        const int ir = static_cast<int>(std::floor((std::sqrt(xx[i] * xx[i] + yy[j] * yy[j] + zz[k] * zz[k]) - roff) / dr)) + 1;
        m->ind_sph[at(i, j, k)] = ir;
        if (m->rhokap[at(i, j, k)] > 0.0 && ir >= 1 && ir <= nr) m->ncount[ir - 1] += p.xyz_symmetry ? nadd_of(i, j, k, true) : 1;
      }
    } else if (gj == 2) {
      m->ind_cyl.assign(static_cast<size_t>(nx) * ny, 0); m->ncount.assign(static_cast<size_t>(nr) * nz, 0);
      m->jp_bins = static_cast<size_t>(nr) * nz;
      for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
        const int ir = static_cast<int>(std::floor((std::sqrt(xx[i] * xx[i] + yy[j] * yy[j]) - roff) / dr)) + 1;
        m->ind_cyl[i + static_cast<size_t>(nx) * j] = ir;
        if (m->rhokap[at(i, j, k)] > 0.0 && ir >= 1 && ir <= nr)
          m->ncount[(ir - 1) + static_cast<size_t>(nr) * k] += p.xy_symmetry ? nadd_of(i, j, k, false) : 1;
      }
    } else if (gj == -1) {
      m->ncount.assign(nz, 0); m->jp_bins = nz;
      for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i)
        if (m->rhokap[at(i, j, k)] > 0.0) m->ncount[k] += 1;
    } else {
      m->jp_bins = nc;
      if (p.xyz_symmetry) {
        m->ncount.assign(nc, 0);
        for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) m->ncount[at(i, j, k)] = nadd_of(i, j, k, true);
      }
    }
    m->jp_nr = nr;
  }
  lart_grid &g = m->cfg.grid;
  g.mask = m->mask.empty() ? nullptr : m->mask.data();
  g.geometry_JPa = (p.calc_J || p.calc_P || p.calc_Pnew) ? p.geometry_JPa : 0; g.nr = m->jp_nr;
  g.ind_sph = m->ind_sph.empty() ? nullptr : m->ind_sph.data();
  g.ind_cyl = m->ind_cyl.empty() ? nullptr : m->ind_cyl.data();
  g.nx = nx; g.ny = ny; g.nz = nz; g.nxfreq = p.nxfreq;
  g.xmin = xmin; g.ymin = ymin; g.zmin = zmin; g.xmax = p.xmax; g.ymax = p.ymax; g.zmax = p.zmax;
  g.dx = dx; g.dy = dy; g.dz = dz;
  g.Dfreq_ref = Dfreq_ref; g.xfreq_min = p.xfreq_min; g.xfreq_max = p.xfreq_max; g.dxfreq = dxfreq;
  g.xcrit = xcrit; g.xcrit2 = xcrit2; g.rmax = p.rmax;
  g.i0 = i0; g.j0 = j0; g.k0 = k0; g.pad_ = 0;
  g.xface = m->xface.data(); g.yface = m->yface.data(); g.zface = m->zface.data();
  g.rhokap = m->rhokap.data(); g.voigt_a = m->voigt_a.data(); g.Dfreq = m->Dfreq.data();
  g.vfx = m->vfx.data(); g.vfy = m->vfy.data(); g.vfz = m->vfz.data();
  g.rhokapD = dust ? m->rhokapD.data() : nullptr;
  lart_host_summary &su = m->sum;
  su.voigt_a = voigt_amean; su.temperature = p.temperature; su.N_gaspole = N_gaspole; su.N_gashomo = N_gashomo;
  su.taupole = taupole; su.tauhomo = tauhomo; su.taupole_dust = taupole_dust; su.tauhomo_dust = tauhomo_dust;
  su.Dfreq_ref = Dfreq_ref; su.vtherm = vtherm; su.cross0 = ln.cross0; su.atau3 = atau3;
  su.xfreq_min = p.xfreq_min; su.xfreq_max = p.xfreq_max; su.dxfreq = dxfreq;
  su.nx = nx; su.ny = ny; su.nz = nz; su.nxfreq = p.nxfreq; su.nphotons = p.nphotons;
  su.zonly = (p.xy_periodic && nx == 1 && ny == 1) ? 1 : 0;
  return 0;
}

// observer_create_outside — observer_rect.f90:10-300
// ---------------------------------------------------------------------------
// grid_create_amr — grid_mod_amr.f90:34-526 for leaf data in the generic format (the file reader itself,
// read_generic_amr.f90, stays the reference's: the leaves arrive through lart_host_set_amr_leaves).
//   amr_build_tree      octree_mod.f90:470-575   insertion of the leaves in input order, internal nodes created on demand
//   amr_build_neighbors octree_mod.f90:619-683   same-level face neighbours, ancestors struck out
//   leaf physics        grid_mod_amr.f90:198-300 (full_neutral / global_dgr defaults, analytic velocity types refused)
//   tau normalisation   grid_mod_amr.f90:343-430 (pole traversal from the root centre along +z)
//   frequency grid      grid_mod_amr.f90:726-790 ;  global core-skip :462-473
// ---------------------------------------------------------------------------
int amr_find_cell_at_level(const lart_host_model *m, double x, double y, double z, int target_level, double xmin, double xmax,
                           double ymin, double ymax, double zmin, double zmax) {  // octree_mod.f90:833-853
  if (x < xmin || x > xmax || y < ymin || y > ymax || z < zmin || z > zmax) return 0;
  int icell = 1;
  for (;;) {
    if (m->a_level[icell - 1] >= target_level) return icell;
    if (m->a_ileaf[icell - 1] > 0) return icell;
    int ioct = 1;
    if (x >= m->a_cx[icell - 1]) ioct += 1;
    if (y >= m->a_cy[icell - 1]) ioct += 2;
    if (z >= m->a_cz[icell - 1]) ioct += 4;
    int child = m->a_children[8 * (size_t)(icell - 1) + ioct - 1];
    if (child == 0) return icell;
    icell = child;
  }
}
int amr_find_leaf_host(const lart_host_model *m, double x, double y, double z, const lart_grid &g) {  // octree_mod.f90:149-171
  if (x < g.xmin || x > g.xmax || y < g.ymin || y > g.ymax || z < g.zmin || z > g.zmax) return 0;
  int icell = 1;
  for (;;) {
    if (m->a_ileaf[icell - 1] > 0) return m->a_ileaf[icell - 1];
    int ioct = 1;
    if (x >= m->a_cx[icell - 1]) ioct += 1;
    if (y >= m->a_cy[icell - 1]) ioct += 2;
    if (z >= m->a_cz[icell - 1]) ioct += 4;
    icell = m->a_children[8 * (size_t)(icell - 1) + ioct - 1];
    if (icell == 0) return 0;
  }
}

int amr_create(lart_host_model *m) {
  Par &p = m->par;
  const Line &ln = m->line;
  const size_t nleaf = m->in_x.size();
  if (nleaf == 0) { g_err = "par%use_amr_grid: no leaf data (lart_host_set_amr_leaves)"; return 1; }
  if (p.xy_periodic || p.xyz_symmetry || p.xy_symmetry) { g_err = "AMR: periodic / mirror boundaries stay with the Fortran host"; return 1; }
  if (!p.velocity_type.empty()) { g_err = "AMR: analytic velocity types stay with the Fortran host (velocities come with the leaves)"; return 1; }
  if (!(p.distance2cm > 0.0)) { p.distance_unit = ""; p.distance2cm = 1.0; }  // no par%distance_unit: code units
  const double L = m->amr_boxlen;
  const double xmin = m->amr_ox, xmax = m->amr_ox + L, ymin = m->amr_oy, ymax = m->amr_oy + L, zmin = m->amr_oz, zmax = m->amr_oz + L;
  // ---- amr_build_tree
  auto &par_ = m->a_parent; auto &chi = m->a_children; auto &lev = m->a_level; auto &ile = m->a_ileaf;
  auto &cx = m->a_cx; auto &cy = m->a_cy; auto &cz = m->a_cz; auto &ch = m->a_ch;
  par_.assign(1, 0); chi.assign(8, 0); lev.assign(1, 0); ile.assign(1, 0);
  cx.assign(1, (xmin + xmax) * 0.5); cy.assign(1, (ymin + ymax) * 0.5); cz.assign(1, (zmin + zmax) * 0.5); ch.assign(1, (xmax - xmin) * 0.5);
  m->a_icell_of_leaf.assign(nleaf, 0);
  for (size_t il = 0; il < nleaf; ++il) {
    int icell = 1;
    for (int l = 0; l <= m->in_level[il] - 1; ++l) {
      int ioct = 1;
      if (m->in_x[il] >= cx[icell - 1]) ioct += 1;
      if (m->in_y[il] >= cy[icell - 1]) ioct += 2;
      if (m->in_z[il] >= cz[icell - 1]) ioct += 4;
      if (chi[8 * (size_t)(icell - 1) + ioct - 1] == 0) {
        const int nc = (int)par_.size() + 1;
        par_.push_back(icell); lev.push_back(l + 1); ile.push_back(0);
        for (int q = 0; q < 8; ++q) chi.push_back(0);
        const double h = ch[icell - 1] * 0.5;
        ch.push_back(h);
        const int ix = (ioct - 1) % 2, iy = ((ioct - 1) / 2) % 2, iz = (ioct - 1) / 4;
        cx.push_back(cx[icell - 1] + (double)(2 * ix - 1) * h);
        cy.push_back(cy[icell - 1] + (double)(2 * iy - 1) * h);
        cz.push_back(cz[icell - 1] + (double)(2 * iz - 1) * h);
        chi[8 * (size_t)(icell - 1) + ioct - 1] = nc;
      }
      icell = chi[8 * (size_t)(icell - 1) + ioct - 1];
    }
    if (ile[icell - 1] != 0 || [&] { for (int q = 0; q < 8; ++q) if (chi[8 * (size_t)(icell - 1) + q]) return true; return false; }()) {
      g_err = "AMR leaf data: two leaves share a cell, or a leaf sits on an internal cell"; return 1;
    }
    ile[icell - 1] = (int32_t)il + 1;
    m->a_icell_of_leaf[il] = icell;
  }
  const int ncells = (int)par_.size();
  // ---- amr_build_neighbors
  m->a_neighbor.assign(6 * (size_t)ncells, 0);
  auto is_ancestor = [&](int anc, int desc) { for (int c = desc; c > 0;) { c = par_[c - 1]; if (c == anc) return true; } return false; };
  for (int ic = 1; ic <= ncells; ++ic) {
    const double x = cx[ic - 1], y = cy[ic - 1], z = cz[ic - 1], hp = 2.0 * ch[ic - 1];
    const int l = lev[ic - 1];
    int32_t *nb = &m->a_neighbor[6 * (size_t)(ic - 1)];
    auto at = [&](double qx, double qy, double qz) { return amr_find_cell_at_level(m, qx, qy, qz, l, xmin, xmax, ymin, ymax, zmin, zmax); };
    if (x + hp <= xmax) nb[0] = at(x + hp, y, z);
    if (x - hp >= xmin) nb[1] = at(x - hp, y, z);
    if (y + hp <= ymax) nb[2] = at(x, y + hp, z);
    if (y - hp >= ymin) nb[3] = at(x, y - hp, z);
    if (z + hp <= zmax) nb[4] = at(x, y, z + hp);
    if (z - hp >= zmin) nb[5] = at(x, y, z - hp);
    for (int f = 0; f < 6; ++f) if (nb[f] > 0 && nb[f] != ic && is_ancestor(nb[f], ic)) nb[f] = 0;
  }
  // ---- leaf physics
  const double Dfreq_ref = m->vtherm_total(p.temperature) / (ln.wavelength0 * kUm2Km);
  const double voigt_amean = (ln.damping / kFourPi) / Dfreq_ref;
  const bool dust = p.DGR > 0.0;
  m->a_rhokap.assign(nleaf, 0.0); m->a_voigt_a.assign(nleaf, 0.0); m->a_Dfreq.assign(nleaf, 0.0);
  m->a_vfx.assign(nleaf, 0.0); m->a_vfy.assign(nleaf, 0.0); m->a_vfz.assign(nleaf, 0.0);
  if (dust) m->a_rhokapD.assign(nleaf, 0.0); else m->a_rhokapD.clear();
  for (size_t il = 0; il < nleaf; ++il) {
    const double T = std::max(m->in_T[il], 10.0);
    const double vth = m->vtherm_total(T);
    m->a_Dfreq[il] = vth / (ln.wavelength0 * kUm2Km);
    m->a_voigt_a[il] = (ln.damping / kFourPi) / m->a_Dfreq[il];
    m->a_rhokap[il] = m->in_nH[il] * ln.cross0 / m->a_Dfreq[il] * p.distance2cm;  // full_neutral: nHI_frac = 1
    if (dust) m->a_rhokapD[il] = m->in_nH[il] * p.cext_dust * p.DGR * p.distance2cm;
    m->a_vfx[il] = m->in_vx[il] / vth; m->a_vfy[il] = m->in_vy[il] / vth; m->a_vfz[il] = m->in_vz[il] / vth;
  }
  lart_grid &g = m->cfg.grid;
  g = lart_grid{};
  g.xmin = xmin; g.xmax = xmax; g.ymin = ymin; g.ymax = ymax; g.zmin = zmin; g.zmax = zmax;
  // ---- tauhomo and the pole traversal (from the root centre along +z; gaps are crossed with the gap cell's geometry)
  auto voigt0 = [](double a) { return 1.0 + a * (-1.1283791671e+00 + a * 1.0); };  // voigt(0,a), as in grid_create
  double opacity_sum = 0.0, nopac = 0.0;
  for (size_t il = 0; il < nleaf; ++il) { opacity_sum += m->a_rhokap[il] * voigt0(m->a_voigt_a[il]); if (m->a_rhokap[il] > 0.0) nopac += 1.0; }
  const double opac_length = L / 2.0;
  double tauhomo = nopac > 0.0 ? (opacity_sum / nopac) * opac_length : 0.0;
  double NHI_pole_raw = 0.0, tau_raw_half = 0.0;
  {
    double xc = cx[0], yc = cy[0], zc = cz[0];
    for (long niter = 0; niter < 10000000 && zc < zmax; ++niter) {
      const int il = amr_find_leaf_host(m, xc, yc, zc, g);
      int icell;
      if (il > 0) icell = m->a_icell_of_leaf[il - 1];
      else {  // amr_find_enclosing_cell, octree_mod.f90:211-237
        if (xc < xmin || xc > xmax || yc < ymin || yc > ymax || zc < zmin || zc > zmax) break;
        icell = 1;
        for (;;) {
          if (ile[icell - 1] > 0) break;
          int ioct = 1;
          if (xc >= cx[icell - 1]) ioct += 1;
          if (yc >= cy[icell - 1]) ioct += 2;
          if (zc >= cz[icell - 1]) ioct += 4;
          const int child = chi[8 * (size_t)(icell - 1) + ioct - 1];
          if (child == 0) break;
          icell = child;
        }
      }
      const double t_exit = (cz[icell - 1] + ch[icell - 1] - zc) / 1.0;  // amr_cell_exit / amr_gap_exit with k = (0,0,1)
      if (il > 0) {
        tau_raw_half += m->a_rhokap[il - 1] * voigt0(m->a_voigt_a[il - 1]) * t_exit;
        NHI_pole_raw += m->a_rhokap[il - 1] * m->a_Dfreq[il - 1] / ln.cross0 * t_exit;
      }
      zc = zc + t_exit;
    }
  }
  double taupole = tau_raw_half > 0.0 ? tau_raw_half : tauhomo;
  double opac_norm = 1.0;
  if (p.taumax > 0.0) { if (taupole > 0.0) opac_norm = p.taumax / taupole; }
  else if (p.N_HImax > 0.0) { if (NHI_pole_raw > 0.0) opac_norm = p.N_HImax / NHI_pole_raw; }
  else if (p.N_gasmax > 0.0) { if (NHI_pole_raw > 0.0) opac_norm = p.N_gasmax / NHI_pole_raw; }
  if (opac_norm != 1.0) {
    for (auto &v : m->a_rhokap) v *= opac_norm;
    for (auto &v : m->a_rhokapD) v *= opac_norm;
  }
  opacity_sum = 0.0;
  for (size_t il = 0; il < nleaf; ++il) opacity_sum += m->a_rhokap[il] * voigt0(m->a_voigt_a[il]);
  if (nopac > 0.0) tauhomo = (opacity_sum / nopac) * opac_length;
  taupole = p.taumax > 0.0 ? p.taumax : tauhomo;
  const double N_HIpole = NHI_pole_raw * opac_norm;
  if (p.N_HImax <= 0.0) p.N_HImax = N_HIpole;
  if (p.N_gasmax <= 0.0) p.N_gasmax = N_HIpole;
  p.tauhomo = tauhomo; p.taumax = taupole;
  // ---- global core skip (:462-473) and the frequency grid (amr_setup_freq_grid)
  const double atau0 = voigt_amean * p.tauhomo;
  const double atau0_cell = atau0 / (L / (2.0 * ch[0]));
  double xcrit = 0.0, xcrit2 = 0.0;
  if (atau0_cell > 1.0) {
    const double xi = atau0_cell <= 60.0 ? 0.6 : 1.4, chi_ = atau0_cell <= 60.0 ? 1.2 : 0.6;
    xcrit = 0.02 * std::exp(xi * std::pow(std::log(atau0_cell), chi_));
    xcrit2 = xcrit * xcrit;
  }
  const double vtherm = m->vtherm_total(p.temperature);
  if (isfin(p.velocity_min) && isfin(p.velocity_max)) {
    if (p.nvelocity == 0 && p.nxfreq > 0) p.nvelocity = p.nxfreq;
    if (p.nvelocity > 0) p.nxfreq = p.nvelocity;
    p.xfreq_min = -p.velocity_max / vtherm; p.xfreq_max = -p.velocity_min / vtherm;
  }
  const double atau1 = std::pow(voigt_amean * p.tauhomo, 1.0 / 3.0);
  p.atau3 = atau1;
  if (!(isfin(p.xfreq_max) && isfin(p.xfreq_min))) {
    double xscale = (p.taumax <= 5e1) ? 25.0 : (p.taumax <= 5e2) ? 14.0 : (p.taumax <= 5e3) ? 10.0 : 5.0;
    const double hk = ln.DnuHK_Hz / Dfreq_ref;
    if (p.Vexp == 0.0) { p.xfreq_max = std::floor(xscale * atau1) + 1.0; p.xfreq_min = -(std::floor(xscale * atau1 + hk) + 1.0); }
    else if (p.Vexp > 0.0) { p.xfreq_max = std::floor(xscale * atau1) + 1.0; p.xfreq_min = -(std::floor(xscale * atau1 + std::fabs(p.Vexp) / vtherm + hk) + 1.0); }
    else { p.xfreq_max = std::floor(xscale * atau1 + std::fabs(p.Vexp) / vtherm) + 1.0; p.xfreq_min = -(std::floor(xscale * atau1 + hk) + 1.0); }
    if (p.spectral_type == "continuum") {
      xscale = 4.0 * xscale;
      p.xfreq_max = std::floor(xscale * atau1 + std::fabs(p.Vexp) / vtherm) + 1.0;
      p.xfreq_min = -(std::floor(xscale * atau1 + std::fabs(p.Vexp) / vtherm + hk) + 1.0);
    }
  }
  const double dxfreq = (p.xfreq_max - p.xfreq_min) / p.nxfreq;
  m->dwave = vtherm / kSpeedC * (ln.wavelength0 * 1e4) * dxfreq;
  // 'sphere' geometry: rmax = L_box/2 (:177-180); the host's box members follow the octree (amr_sync_to_grid, :1017-1058)
  if (p.geometry == "sphere" || p.rmax > 0.0) p.rmax = L * 0.5;
  p.xmax = xmax; p.ymax = ymax; p.zmax = zmax;
  g.nx = g.ny = g.nz = 1; g.nxfreq = p.nxfreq;
  g.dx = g.dy = g.dz = L;
  g.Dfreq_ref = Dfreq_ref; g.xfreq_min = p.xfreq_min; g.xfreq_max = p.xfreq_max; g.dxfreq = dxfreq;
  g.xcrit = xcrit; g.xcrit2 = xcrit2; g.rmax = p.rmax;
  lart_amr &a = m->cfg.amr;
  a.ncells = ncells; a.nleaf = (int32_t)nleaf;
  a.children = chi.data(); a.ileaf = ile.data(); a.icell_of_leaf = m->a_icell_of_leaf.data(); a.neighbor = m->a_neighbor.data();
  a.cx = cx.data(); a.cy = cy.data(); a.cz = cz.data(); a.ch = ch.data();
  a.rhokap = m->a_rhokap.data(); a.voigt_a = m->a_voigt_a.data(); a.Dfreq = m->a_Dfreq.data();
  a.vfx = m->a_vfx.data(); a.vfy = m->a_vfy.data(); a.vfz = m->a_vfz.data();
  a.rhokapD = dust ? m->a_rhokapD.data() : nullptr;
  lart_host_summary &su = m->sum;
  su.voigt_a = voigt_amean; su.temperature = p.temperature; su.N_gaspole = N_HIpole; su.N_gashomo = N_HIpole;
  su.taupole = taupole; su.tauhomo = tauhomo; su.taupole_dust = 0.0; su.tauhomo_dust = 0.0;
  su.Dfreq_ref = Dfreq_ref; su.vtherm = vtherm; su.cross0 = ln.cross0; su.atau3 = atau1;
  su.xfreq_min = p.xfreq_min; su.xfreq_max = p.xfreq_max; su.dxfreq = dxfreq;
  su.nx = su.ny = su.nz = 1; su.nxfreq = p.nxfreq; su.nphotons = p.nphotons; su.zonly = 0;
  su.nclumps = 0;
  return 0;
}

int observer_create(lart_host_model *m) {
  Par &p = m->par;
  if (!isfin(p.rotation_center_x)) p.rotation_center_x = 0.0;
  if (!isfin(p.rotation_center_y)) p.rotation_center_y = 0.0;
  if (!isfin(p.rotation_center_z)) p.rotation_center_z = 0.0;
  const size_t NMAX = LART_MAX_OBSERVERS;
  auto pad = [&](std::vector<double> &v) { v.resize(NMAX, kNaN); };
  pad(p.alpha); pad(p.beta); pad(p.gamma); pad(p.obsx); pad(p.obsy); pad(p.obsz);
  pad(p.phase_angle); pad(p.inclination_angle); pad(p.position_angle);
  auto anyfinite = [](const std::vector<double> &v) { for (double d : v) if (isfin(d)) return true; return false; };
  if (anyfinite(p.phase_angle)) for (size_t i = 0; i < NMAX; ++i) p.alpha[i] = -p.phase_angle[i];  // :50-52
  if (anyfinite(p.inclination_angle)) for (size_t i = 0; i < NMAX; ++i) p.beta[i] = -p.inclination_angle[i];
  if (anyfinite(p.position_angle)) for (size_t i = 0; i < NMAX; ++i) p.gamma[i] = -p.position_angle[i];
  for (size_t i = 0; i < NMAX; ++i) {  // :56-57
    if (isfin(p.beta[i]) && !isfin(p.alpha[i])) p.alpha[i] = 0.0;
    if (isfin(p.alpha[i]) && !isfin(p.beta[i])) p.beta[i] = 0.0;
  }
  const double boxmax = std::max({p.xmax, p.ymax, p.zmax});
  if (!(isfin(p.alpha[0]) && isfin(p.beta[0])) && !(isfin(p.obsx[0]) && isfin(p.obsy[0]) && isfin(p.obsz[0]))) {  // :60-73
    if (!isfin(p.distance)) p.distance = boxmax * 100.0;
    p.obsx[0] = 0.0; p.obsy[0] = 0.0; p.obsz[0] = 1.0; p.alpha[0] = 0.0; p.beta[0] = 0.0;
  }
  std::vector<double> cosa, sina, cosb, sinb, cosg, sing;
  auto default_gamma = [&](size_t i) {
    if (!isfin(p.gamma[i])) p.gamma[i] = (p.beta[i] > 0.0 && p.beta[i] <= 90.0) ? 90.0 : (p.beta[i] > 90.0) ? -90.0 : 0.0;
  };
  m->observers.clear();
  if (isfin(p.alpha[0]) && isfin(p.beta[0])) {  // :75-123
    p.nobs = 0;
    for (size_t i = 0; i < NMAX; ++i) if (isfin(p.alpha[i]) && isfin(p.beta[i])) ++p.nobs;
    if (!isfin(p.distance)) p.distance = boxmax * 100.0;
    for (int i = 0; i < p.nobs; ++i) {
      default_gamma(i);
      cosa.push_back(std::cos(p.alpha[i] * kDeg2Rad)); sina.push_back(std::sin(p.alpha[i] * kDeg2Rad));
      cosb.push_back(std::cos(p.beta[i] * kDeg2Rad)); sinb.push_back(std::sin(p.beta[i] * kDeg2Rad));
      cosg.push_back(std::cos(p.gamma[i] * kDeg2Rad)); sing.push_back(std::sin(p.gamma[i] * kDeg2Rad));
      lart_observer o{};
      o.x = p.distance * cosa[i] * sinb[i] + p.rotation_center_x;
      o.y = p.distance * sina[i] * sinb[i] + p.rotation_center_y;
      o.z = p.distance * cosb[i] + p.rotation_center_z;
      m->observers.push_back(o);
    }
  } else {  // :124-199
    p.nobs = 0;
    for (size_t i = 0; i < NMAX; ++i) if (isfin(p.obsx[i]) && isfin(p.obsy[i]) && isfin(p.obsz[i])) ++p.nobs;
    if (!isfin(p.distance)) {
      p.distance = std::sqrt(p.obsx[0] * p.obsx[0] + p.obsy[0] * p.obsy[0] + p.obsz[0] * p.obsz[0]);
      if (p.distance < 10.0 * boxmax) p.distance = boxmax * 100.0;
    }
    for (int i = 0; i < p.nobs; ++i) {
      default_gamma(i);
      lart_observer o{};
      double dist_scale = p.distance / std::sqrt(p.obsx[i] * p.obsx[i] + p.obsy[i] * p.obsy[i] + p.obsz[i] * p.obsz[i]);
      if (dist_scale > 1.001) {
        o.x = p.obsx[i] * dist_scale + p.rotation_center_x; o.y = p.obsy[i] * dist_scale + p.rotation_center_y; o.z = p.obsz[i] * dist_scale + p.rotation_center_z;
      } else {
        o.x = p.obsx[i]; o.y = p.obsy[i]; o.z = p.obsz[i];
      }
      double cb = (o.z - p.rotation_center_z) / p.distance;
      if (std::fabs(cb - 1.0) < kEps) cb = 1.0;
      if (std::fabs(cb + 1.0) < kEps) cb = -1.0;
      double sb = std::sqrt(1.0 - cb * cb);
      p.beta[i] = std::atan2(sb, cb) * kRad2Deg;
      cosb.push_back(cb); sinb.push_back(sb);
      cosg.push_back(std::cos(p.gamma[i] * kDeg2Rad)); sing.push_back(std::sin(p.gamma[i] * kDeg2Rad));
      if (sb == 0.0) { cosa.push_back(1.0); sina.push_back(0.0); p.alpha[i] = 0.0; }
      else {
        double al = std::atan2(o.y - p.rotation_center_y, o.x - p.rotation_center_x);
        cosa.push_back(std::cos(al)); sina.push_back(std::sin(al)); p.alpha[i] = al * kRad2Deg;
      }
      m->observers.push_back(o);
    }
  }
  for (int i = 0; i < p.nobs; ++i) {  // :209-219; rmatrix(r,c) at [(r-1)+3*(c-1)]
    double *R = m->observers[i].rmatrix;
    R[0 + 3 * 0] = cosa[i] * cosb[i] * cosg[i] - sina[i] * sing[i];
    R[0 + 3 * 1] = sina[i] * cosb[i] * cosg[i] + cosa[i] * sing[i];
    R[0 + 3 * 2] = -sinb[i] * cosg[i];
    R[1 + 3 * 0] = -cosa[i] * cosb[i] * sing[i] - sina[i] * cosg[i];
    R[1 + 3 * 1] = -sina[i] * cosb[i] * sing[i] + cosa[i] * cosg[i];
    R[1 + 3 * 2] = sinb[i] * sing[i];
    R[2 + 3 * 0] = cosa[i] * sinb[i];
    R[2 + 3 * 1] = sina[i] * sinb[i];
    R[2 + 3 * 2] = cosb[i];
  }
  if (!(isfin(p.dxim) && isfin(p.dyim))) {  // :244-281
    if (p.geometry == "sphere") {
      p.dxim = std::asin(p.rmax / p.distance) / (p.nxim / 2.0) * kRad2Deg;
      p.dyim = std::asin(p.rmax / p.distance) / (p.nyim / 2.0) * kRad2Deg;
    } else {
      static const double vx[8] = {1, 1, 1, -1, -1, -1, 1, -1}, vy[8] = {1, 1, -1, 1, -1, 1, -1, -1}, vz[8] = {1, -1, 1, 1, 1, -1, -1, -1};
      double max_ang_x = -999.0, max_ang_y = -999.0;
      for (int i = 0; i < p.nobs; ++i) for (int iv = 0; iv < 8; ++iv) {
        const lart_observer &o = m->observers[i];
        const double *R = o.rmatrix;
        double px = o.x - vx[iv] * p.xmax, py = o.y - vy[iv] * p.ymax, pz = o.z - vz[iv] * p.zmax;
        double kx = R[0] * px + R[3] * py + R[6] * pz, ky = R[1] * px + R[4] * py + R[7] * pz, kz = R[2] * px + R[5] * py + R[8] * pz;
        max_ang_x = std::max(max_ang_x, std::fabs(std::atan2(-kx, kz)));
        max_ang_y = std::max(max_ang_y, std::fabs(std::atan2(-ky, kz)));
      }
      if (p.nxim == p.nyim) {
        p.dxim = std::max(max_ang_x, max_ang_y) / (p.nxim / 2.0) * kRad2Deg;
        p.dyim = std::max(max_ang_x, max_ang_y) / (p.nyim / 2.0) * kRad2Deg;
      } else {
        if (!isfin(p.dxim)) p.dxim = max_ang_x / (p.nxim / 2.0) * kRad2Deg;
        if (!isfin(p.dyim)) p.dyim = max_ang_y / (p.nyim / 2.0) * kRad2Deg;
      }
    }
  }
  m->steradian_pix.assign(p.nobs, p.dxim * p.dyim * (kDeg2Rad * kDeg2Rad));  // :284
  for (int i = 0; i < p.nobs; ++i) {
    lart_observer &o = m->observers[i];
    o.dxim = p.dxim; o.dyim = p.dyim; o.nxim = p.nxim; o.nyim = p.nyim;
  }
  return 0;
}

double *alloc(lart_host_model *m, size_t n) {
  m->store.emplace_back(n, 0.0);
  return m->store.back().data();
}

void alloc_tallies(lart_host_model *m) {
  const Par &p = m->par;
  m->store.clear();
  m->tal = lart_tallies{};
  const size_t nxf = p.nxfreq;
  m->tal.Jout = alloc(m, nxf);
  if (p.save_Jin) m->tal.Jin = alloc(m, nxf);
  if (p.DGR > 0.0 && p.save_Jabs) m->tal.Jabs = alloc(m, nxf);
  if (p.geometry == "plane_atmosphere" || p.geometry == "spherical_atmosphere") m->tal.Jabs2 = alloc(m, nxf);  // grid_mod_car.f90:1179-1183
  if (p.calc_J) m->tal.J = alloc(m, nxf * m->jp_bins);   // :1422-1434
  if (p.calc_P) m->tal.Pa = alloc(m, m->jp_bins);        // :1385-1395
  if (p.calc_Pnew) m->tal.Pnew = alloc(m, m->jp_bins);   // :1410-1420
  if (p.save_Jmu) m->tal.Jmu = alloc(m, nxf * p.nmu);
  m->obs_out.assign(p.save_peeloff ? p.nobs : 0, lart_observer_out{});
  for (auto &oo : m->obs_out) {  // observer_rect.f90:286-330
    const size_t n2 = static_cast<size_t>(p.nxim) * p.nyim, n3 = n2 * nxf;
    if (p.save_peeloff_2D) {
      oo.scatt_2D = alloc(m, n2); oo.direc_2D = alloc(m, n2);
      if (p.save_direc0) oo.direc0_2D = alloc(m, n2);
      if (p.use_stokes) { oo.I_2D = alloc(m, n2); oo.Q_2D = alloc(m, n2); oo.U_2D = alloc(m, n2); oo.V_2D = alloc(m, n2); }
    }
    if (p.save_peeloff_3D) {
      oo.scatt = alloc(m, n3); oo.direc = alloc(m, n3);
      if (p.save_direc0) oo.direc0 = alloc(m, n3);
      if (p.use_stokes) { oo.I = alloc(m, n3); oo.Q = alloc(m, n3); oo.U = alloc(m, n3); oo.V = alloc(m, n3); }
    }
  }
  m->tal.obs = m->obs_out.empty() ? nullptr : m->obs_out.data();
  if (p.save_all_photons) {  // grid_mod_car.f90:1130-1146
    const size_t n = static_cast<size_t>(p.nphotons);
    lart_allph_out &a = m->tal.allph;
    a.rp = alloc(m, n); a.xfreq1 = alloc(m, n); a.xfreq2 = alloc(m, n); a.nscatt_gas = alloc(m, n); a.nscatt_dust = alloc(m, n);
    if (p.source_geometry != "point") a.rp0 = alloc(m, n);
    if (p.use_stokes) { a.I = alloc(m, n); a.Q = alloc(m, n); a.U = alloc(m, n); a.V = alloc(m, n); }
  }
}

}  // namespace

extern "C" {

const char *lart_host_last_error(void) { return g_err.c_str(); }

lart_host_model *lart_host_new(void) { return new lart_host_model(); }
void lart_host_free(lart_host_model *m) { delete m; }

int lart_host_set(lart_host_model *m, const char *key, const char *value) {
  if (!m || !key || !value) { g_err = "lart_host_set: null argument"; return 1; }
  m->is_setup = false;
  return set_key(m, key, value);
}

int lart_host_read_input(lart_host_model *m, const char *path) {
  std::ifstream f(path);
  if (!f) { g_err = std::string("cannot open ") + path; return 1; }
  std::string l;
  bool in = false;
  while (std::getline(f, l)) {
    size_t ex = l.find('!');
    if (ex != std::string::npos) l = l.substr(0, ex);
    l = trim(l);
    if (l.empty()) continue;
    if (l[0] == '&') { in = true; continue; }
    if (l == "/") { in = false; continue; }
    if (!in) continue;
    size_t eq = l.find('=');
    if (eq == std::string::npos) continue;
    if (int rc = set_key(m, l.substr(0, eq), l.substr(eq + 1))) return rc;
  }
  return 0;
}

int lart_host_set_amr_leaves(lart_host_model *m, int64_t n, const double *x, const double *y, const double *z, const int32_t *level,
                             const double *nH, const double *T, const double *vx, const double *vy, const double *vz,
                             double boxlen, double origin_x, double origin_y, double origin_z) {
  if (!m || n < 1 || !x || !y || !z || !level || !nH || !T || !(boxlen > 0.0)) { g_err = "lart_host_set_amr_leaves: bad argument"; return 1; }
  m->is_setup = false;
  m->in_x.assign(x, x + n); m->in_y.assign(y, y + n); m->in_z.assign(z, z + n); m->in_level.assign(level, level + n);
  m->in_nH.assign(nH, nH + n); m->in_T.assign(T, T + n);
  if (vx && vy && vz) { m->in_vx.assign(vx, vx + n); m->in_vy.assign(vy, vy + n); m->in_vz.assign(vz, vz + n); }
  else { m->in_vx.assign(n, 0.0); m->in_vy.assign(n, 0.0); m->in_vz.assign(n, 0.0); }
  m->amr_boxlen = boxlen; m->amr_ox = origin_x; m->amr_oy = origin_y; m->amr_oz = origin_z;
  m->par.use_amr_grid = true;
  return 0;
}

int lart_host_setup(lart_host_model *m) {
  if (!m) { g_err = "lart_host_setup: null model"; return 1; }
  if (int rc = derive(m)) return rc;
  m->cfg.amr = lart_amr{};
  if (m->par.use_amr_grid) {
    if (m->par.use_clump_medium) { g_err = "use_amr_grid and use_clump_medium exclude each other"; return 1; }
    m->cfg.clumps = lart_clumps{};
    if (int rc = amr_create(m)) return rc;
  } else {
    if (int rc = clumps_create(m)) return rc;  // grid_create_clump: the population first (grid_mod_clump.f90:60-80)
    if (int rc = grid_create(m)) return rc;
  }
  Par &p = m->par;
  if (p.use_clump_medium) {  // :93-101 — the box carries no opacity, no bulk velocity, the clumps' Doppler width
    std::fill(m->rhokap.begin(), m->rhokap.end(), 0.0);
    std::fill(m->Dfreq.begin(), m->Dfreq.end(), m->cfg.clumps.Dfreq_ref);
    std::fill(m->voigt_a.begin(), m->voigt_a.end(), m->cl_voigt_a[0]);
    std::fill(m->vfx.begin(), m->vfx.end(), 0.0); std::fill(m->vfy.begin(), m->vfy.end(), 0.0); std::fill(m->vfz.begin(), m->vfz.end(), 0.0);
    m->cfg.grid.Dfreq_ref = m->cfg.clumps.Dfreq_ref;
    m->sum.nclumps = m->cfg.clumps.n;
    // the system-level scalars are the clump-derived ones (grid_mod_clump.f90:60-80), not the empty box's
    m->sum.tauhomo = p.tauhomo; m->sum.taupole = p.taumax; m->sum.N_gashomo = p.N_gashomo; m->sum.N_gaspole = p.N_gasmax;
  }
  if (p.save_peeloff) { if (int rc = observer_create(m)) return rc; } else { p.nobs = 0; m->observers.clear(); }
  if (p.nobs == 0) { p.save_peeloff = false; p.save_peeloff_2D = false; p.save_peeloff_3D = false; }
  lart_config &c = m->cfg;
  lart_params &q = c.par;
  q.nphotons = p.nphotons; q.seed = static_cast<uint64_t>(p.iseed);
  q.xfreq0 = p.xfreq0; q.xs_point = p.xs_point; q.ys_point = p.ys_point; q.zs_point = p.zs_point;
  q.source_rmax = p.source_rmax; q.DGR = p.DGR; q.albedo = p.albedo; q.hgg = p.hgg;
  // 'voigt0' source width — setup.f90:140-142 (always derived from temperature0)
  {
    double vt0 = m->vtherm_total(p.temperature0);
    p.Dfreq0 = vt0 / (m->line.wavelength0 * kUm2Km);
    p.voigt_a0 = (m->line.damping / kFourPi) / p.Dfreq0;
  }
  q.voigt_a0 = p.voigt_a0; q.Dfreq0 = p.Dfreq0;
  {
    double sig = p.gaussian_sigma_vel;
    if (p.gaussian_FWHM_vel > 0.0) sig = p.gaussian_FWHM_vel / 2.3548200450309493;  // generate_photon.f90:258-264
    q.gaussian_sigma_x = sig / m->vtherm_total(p.temperature);
  }
  q.mu_min = p.mu_min; q.dmu = p.dmu; q.nmu = p.nmu;
  q.spectral_type = (p.spectral_type == "voigt") ? LART_SPEC_VOIGT : (p.spectral_type == "voigt0") ? LART_SPEC_VOIGT0
                    : (p.spectral_type == "continuum") ? LART_SPEC_CONTINUUM : (p.spectral_type == "gaussian") ? LART_SPEC_GAUSSIAN : LART_SPEC_MONO;
  q.source_geometry = (p.source_geometry == "uniform") ? LART_SRC_UNIFORM
                      : (p.source_geometry == "uniform_sphere" || p.source_geometry == "sphere") ? LART_SRC_UNIFORM_SPHERE : LART_SRC_POINT;
  q.comoving_source = p.comoving_source; q.recoil = p.recoil; q.core_skip = p.core_skip; q.core_skip_global = p.core_skip_global;
  q.use_stokes = p.use_stokes; q.use_reduced_wgt = p.use_reduced_wgt;
  q.save_Jin = p.save_Jin; q.save_Jabs = p.save_Jabs; q.save_Jmu = p.save_Jmu;
  q.save_peeloff = p.save_peeloff; q.save_peeloff_2D = p.save_peeloff_2D; q.save_peeloff_3D = p.save_peeloff_3D; q.save_direc0 = p.save_direc0;
  if (p.source_geometry == "plane_illumination") q.source_geometry = LART_SRC_PLANE_ILLUMINATION;
  q.atmosphere = (p.geometry == "plane_atmosphere") ? LART_ATM_PLANE : (p.geometry == "spherical_atmosphere") ? LART_ATM_SPHERICAL : LART_ATM_NONE;
  q.calc_J = p.calc_J; q.calc_P = p.calc_P; q.calc_Pnew = p.calc_Pnew; q.Omega = p.Omega;
  q.save_all_photons = p.save_all_photons; q.xy_periodic = p.xy_periodic; q.xyz_symmetry = p.xyz_symmetry; q.xy_symmetry = p.xy_symmetry; q.use_clump_medium = p.use_clump_medium; q.nobs = p.nobs; q.use_amr_grid = p.use_amr_grid;
  const Line &ln = m->line;
  c.line.line_type = ln.line_type; c.line.E1 = ln.E1; c.line.E2 = ln.E2; c.line.E3 = ln.E3;
  c.line.g_recoil0 = ln.g_recoil0; c.line.DnuHK_Hz = ln.DnuHK_Hz; c.line.cross0 = ln.cross0;
  lart_scatt_mat &sm = c.scatt_mat;
  sm = lart_scatt_mat{};
  if (!m->sm_coss.empty() && p.DGR > 0.0) {
    sm.nPDF = static_cast<int32_t>(m->sm_coss.size());
    sm.coss = m->sm_coss.data(); sm.S11 = m->sm_S11.data(); sm.S12 = m->sm_S12.data(); sm.S33 = m->sm_S33.data(); sm.S34 = m->sm_S34.data();
    sm.phase_PDF = m->sm_pdf.data(); sm.alias = m->sm_alias.data();
  }
  if (p.DGR > 0.0 && p.use_stokes && sm.nPDF == 0) { g_err = "dust + Stokes needs par%scatt_mat_file"; return 1; }
  c.observers = m->observers.empty() ? nullptr : m->observers.data();
  m->sum.nobs = p.nobs; m->sum.nxim = p.nxim; m->sum.nyim = p.nyim; m->sum.dxim = p.dxim; m->sum.dyim = p.dyim;
  m->sum.distance = isfin(p.distance) ? p.distance : 0.0;
  alloc_tallies(m);
  m->is_setup = true;
  return 0;
}

const lart_config *lart_host_config(const lart_host_model *m) { return (m && m->is_setup) ? &m->cfg : nullptr; }

int lart_host_get_summary(const lart_host_model *m, lart_host_summary *out) {
  if (!m || !m->is_setup || !out) { g_err = "lart_host_get_summary: model not set up"; return 1; }
  *out = m->sum;
  return 0;
}

lart_tallies *lart_host_tallies(lart_host_model *m) { return (m && m->is_setup) ? &m->tal : nullptr; }

int lart_host_zero_tallies(lart_host_model *m) {
  if (!m || !m->is_setup) { g_err = "lart_host_zero_tallies: model not set up"; return 1; }
  for (auto &v : m->store) std::fill(v.begin(), v.end(), 0.0);
  m->tal.nscatt_gas = 0; m->tal.nscatt_dust = 0; m->tal.counters = lart_counters{};
  return 0;
}

// output_normalize_outside — output_sum_rect.f90:151-487
int lart_host_normalize(lart_host_model *m) {
  if (!m || !m->is_setup) { g_err = "lart_host_normalize: model not set up"; return 1; }
  Par &p = m->par;
  const lart_grid &g = m->cfg.grid;
  lart_tallies &t = m->tal;
  const double nph = static_cast<double>(p.nphotons);
  t.nscatt_dust /= nph; t.nscatt_gas /= nph;  // :163-165
  p.nscatt_dust = t.nscatt_dust; p.nscatt_gas = t.nscatt_gas; p.nscatt_tot = p.nscatt_gas + p.nscatt_dust;
  const double bin_unit = (p.intensity_unit == 1) ? m->dwave : g.dxfreq;  // :168-172
  double area;
  if (p.xy_periodic) area = 2.0;  // :185
  else if (p.geometry == "sphere") area = kFourPi * g.rmax * g.rmax * p.distance2cm * p.distance2cm;  // :194
  else area = (g.xmax * g.ymax + g.ymax * g.zmax + g.zmax * g.xmax) * 8.0 * p.distance2cm * p.distance2cm;
  const double den = nph * bin_unit * kTwoPi * area;
  const size_t nxf = g.nxfreq;
  for (size_t i = 0; i < nxf; ++i) {
    t.Jout[i] /= den;
    if (t.Jin) t.Jin[i] /= den;
    if (t.Jabs) t.Jabs[i] /= den;
    if (t.Jabs2) t.Jabs2[i] /= den;  // :238-248
  }
  if (t.Jmu) for (size_t i = 0; i < nxf * p.nmu; ++i) t.Jmu[i] = t.Jmu[i] * p.nmu / den;
  // CALCJ / CALCP / CALCPnew — :275-400: per-bin cell counts, cell volume, and for periodic boxes the surface area
  if (t.J || t.Pa || t.Pnew) {
    const double d2 = p.distance2cm * p.distance2cm;
    const double dVol = g.dx * g.dy * g.dz * d2;
    const double slab_area = (g.xmax - g.xmin) * (g.ymax - g.ymin) * d2;  // :177
    const int gj = p.geometry_JPa;
    const size_t nb = m->jp_bins;
    for (size_t b = 0; b < nb; ++b) {
      double fJ, fP;  // multiplicative factors of J(:,b) and P(b)
      const double cnt = m->ncount.empty() ? 1.0 : static_cast<double>(m->ncount[b]);
      if (gj == 3) {
        fJ = p.xy_periodic ? slab_area / (kFourPi * dVol * nph * bin_unit) : 1.0 / (kFourPi * dVol * nph * bin_unit);
        fP = p.xy_periodic ? slab_area / (dVol * nph) : 1.0 / (dVol * nph);
        if (p.xyz_symmetry) { fJ /= cnt; fP /= cnt; }
      } else if (gj == -1) {
        if (!(cnt > 0.0)) continue;
        fJ = 1.0 / cnt * (slab_area / (kFourPi * dVol * nph * bin_unit));
        fP = 1.0 / cnt * (slab_area / (dVol * nph));
      } else {
        if (!(cnt > 0.0)) continue;
        fJ = 1.0 / cnt / (kFourPi * dVol * nph * bin_unit);
        fP = 1.0 / cnt / (dVol * nph);
      }
      if (t.J) for (size_t i = 0; i < nxf; ++i) t.J[i + nxf * b] *= fJ;
      if (t.Pa) t.Pa[b] *= fP;
      if (t.Pnew) t.Pnew[b] *= fP;
    }
  }
  // continuum runs are expressed in units of the input continuum level — output_sum_rect.f90:252-273
  // (par%continuum_normalize defaults to .true., define.f90:314; the reference aborts without save_Jin)
  if (p.spectral_type == "continuum" && p.continuum_normalize) {
    if (!t.Jin) { g_err = "ERROR: continuum_normalize=T requires save_Jin=T"; return 1; }
    double mean = 0.0;
    for (size_t i = 0; i < nxf; ++i) mean += t.Jin[i];
    mean /= static_cast<double>(nxf);
    const double scale = (p.f_line > 0.0 && p.f_line < 1.0) ? mean * (1.0 - p.f_line) : mean;
    for (size_t i = 0; i < nxf; ++i) {
      t.Jout[i] /= scale;
      t.Jin[i] /= scale;
      if (t.Jabs) t.Jabs[i] /= scale;
      if (t.Jabs2) t.Jabs2[i] /= scale;
    }
    if (t.Jmu) for (size_t i = 0; i < nxf * p.nmu; ++i) t.Jmu[i] /= scale;
  }
  for (size_t k = 0; k < m->obs_out.size(); ++k) {  // :405-450
    lart_observer_out &oo = m->obs_out[k];
    const size_t n2 = static_cast<size_t>(p.nxim) * p.nyim, n3 = n2 * nxf;
    const double s2 = p.no_photons * m->steradian_pix[k] * p.distance2cm * p.distance2cm, s3 = s2 * bin_unit;
    auto div = [](double *a, size_t n, double s) { if (a) for (size_t i = 0; i < n; ++i) a[i] /= s; };
    div(oo.scatt_2D, n2, s2); div(oo.direc_2D, n2, s2); div(oo.direc0_2D, n2, s2);
    div(oo.I_2D, n2, s2); div(oo.Q_2D, n2, s2); div(oo.U_2D, n2, s2); div(oo.V_2D, n2, s2);
    div(oo.scatt, n3, s3); div(oo.direc, n3, s3); div(oo.direc0, n3, s3);
    div(oo.I, n3, s3); div(oo.Q, n3, s3); div(oo.U, n3, s3); div(oo.V, n3, s3);
  }
  return 0;
}

}  // extern "C"
