// lart_oracle.cpp — CPU ORACLE (test infrastructure, NOT product code).
//
// A scalar, one-photon-at-a-time FP64 restatement of the reference's Cartesian
// photon loop (LaRT v2.00) — with its folded / periodic / shearing / atmosphere boundary
// variants, sight-line maps, the CALCJ / CALCP / CALCPnew accumulators, the clump medium
// (with and without overlapping clumps) and the octree — written to be read side by side
// with the Fortran.
// Every function cites the reference file:line it follows (paths relative to
// the reference tree).  It exists so that the CUDA path has something to be
// checked against: only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library.  The product
// (lart_b200/) never links, imports or falls back to it.
//
// PARITY PINNING.  The reference ships no unit-level golden vectors and cannot
// be compiled here (no Fortran/MPI toolchain), so at function level this oracle
// is "parity unpinned": it is pinned only (a) by construction against the cited
// lines, (b) by the whole-run known answers the reference's logs hold
// (<N_scatt> = 1.7898e3 / 2.8225e4, voigt_a, N(HI)_pole — tests/test_oracle_pins.py; for the clump medium the
// population figures and <N_scatt> = 4.3454e3 of examples/clump_sphere/log_back — tests/test_oracle_clumps.py,
// tests/test_gpu_clumps.py; for the octree nleaf = 178 480, N(HI)_pole and <N_scatt> = 2.8263e4 of
// examples/amr_sphere_generic/log_amr_1M.txt — tests/test_oracle_amr.py; the documented J_out peaks of the 101^3 sphere —
// tests/test_gpu_stats.py; all transcribed into tests/golden/reference_logs.json),
// (c) by independent mathematics (Harris functions, scipy wofz, Neufeld/Dijkstra
// analytic spectra, Philox and MT19937-64 known-answer vectors, brute-force sums over all clumps for the overlap
// event walk, binning identities and Pa ~ Pnew for the accumulators, photon conservation between Jout and Jabs2).
//
// The struct layouts come from the product's public header include/lart_gpu.h
// (POD declarations only) so both sides consume identical host arrays.
//
// Build: see oracle/Makefile (g++ -O3 -ffp-contract=off; no FMA contraction so
// the DDA arithmetic is plain IEEE, matching the kernel's __dadd_rn/__dmul_rn).

#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <vector>

#include "../include/lart_gpu.h"
#include "voigt_tables.inc"

namespace {

constexpr double kPi = 3.141592653589793238462643383279502884197;      // define.f90:45
constexpr double kTwoPi = 6.283185307179586476925286766559005768394;   // define.f90:46
constexpr double kFourPi = 12.56637061435917295385057353311801153679;  // define.f90:47
constexpr double kHalfPi = kPi / 2.0;
constexpr double kRad2Deg = 180.0 / kPi;                               // define.f90:57
constexpr double kHugest = DBL_MAX;                                    // define.f90:38
constexpr double kTauHuge = 745.2;                                     // raytrace_car.f90:432

// ---------------------------------------------------------------------------
// Uniform generators.  Both map a 64-bit word w to ((w>>12)+0.5)*2^-52, an
// OPEN interval (0,1) — random_mt.f90:628-629.
// ---------------------------------------------------------------------------
inline double word_to_open01(uint64_t w) {
  return (static_cast<double>(w >> 12) + 0.5) * (1.0 / 4503599627370496.0);
}

// MT19937-64 — random_mt.f90:579-630 (generation + tempering), :486-498 (init_mt).
struct Mt64 {
  static constexpr int N = 312, M = 156;
  uint64_t mt[N];
  int mti = N + 1;
  void init(int64_t seed) {
    mt[0] = static_cast<uint64_t>(seed);
    for (int i = 1; i < N; ++i)
      mt[i] = 6364136223846793005ULL * (mt[i - 1] ^ (mt[i - 1] >> 62)) + static_cast<uint64_t>(i);
    mti = N;
  }
  uint64_t next() {
    static const uint64_t mag01[2] = {0ULL, 0xB5026F5AA96619E9ULL};
    const uint64_t UM = 0xFFFFFFFF80000000ULL, LM = 0x7FFFFFFFULL;
    if (mti >= N) {
      if (mti == N + 1) init(5489);
      int i;
      for (i = 0; i < N - M; ++i) {
        uint64_t x = (mt[i] & UM) | (mt[i + 1] & LM);
        mt[i] = mt[i + M] ^ (x >> 1) ^ mag01[x & 1ULL];
      }
      for (; i < N - 1; ++i) {
        uint64_t x = (mt[i] & UM) | (mt[i + 1] & LM);
        mt[i] = mt[i + (M - N)] ^ (x >> 1) ^ mag01[x & 1ULL];
      }
      uint64_t x = (mt[N - 1] & UM) | (mt[0] & LM);
      mt[N - 1] = mt[M - 1] ^ (x >> 1) ^ mag01[x & 1ULL];
      mti = 0;
    }
    uint64_t x = mt[mti++];
    x ^= (x >> 29) & 0x5555555555555555ULL;
    x ^= (x << 17) & 0x71D67FFFEDA60000ULL;
    x ^= (x << 37) & 0xFFF7EEE000000000ULL;
    x ^= (x >> 43);
    return x;
  }
};

// Philox4x32-10 (Salmon et al. 2011, Random123).  Not in the reference: it is
// the counter-based generator the north star prescribes for the device; the
// oracle carries it so GPU and oracle can be compared photon by photon.
//   key     = (seed_lo, seed_hi)
//   counter = (id_lo, id_hi, block_lo, block_hi)
// One block yields two uniforms: word0 = c0 | c1<<32, word1 = c2 | c3<<32.
inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * c[0];
    uint64_t p1 = static_cast<uint64_t>(0xCD9E8D57u) * c[2];
    uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = static_cast<uint32_t>(p1);
    uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = static_cast<uint32_t>(p0);
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

struct Counters {
  double n_photons_done = 0, n_scatter = 0, n_cellsteps = 0, n_peel = 0, n_rng = 0, n_reject_iter = 0;
};

// rand_number() + rand_gauss() state.  MT mode keeps the Marsaglia spare across
// photons like the Fortran `save` variable (random_mt.f90:968-987); Philox mode
// restarts stream and spare for every photon.
//
// Philox stream layout (shared with the GPU engine): every call consumes ONE
// Philox block of the photon's stream — uniform() uses its first 64-bit word,
// uniform2() both words.  Call sites that need two uniforms at the same program
// point use uniform2(); in MT mode it is simply two consecutive rand_number()
// calls, so the MT sequence is the reference's own draw order.
struct Rng {
  int mode = 0;  // 0 = MT19937-64, 1 = Philox
  Mt64 mt;
  uint64_t seed = 0, stream = 0, nblk = 0;
  bool gauss_stored = false;
  double gset = 0.0;
  Counters *cnt = nullptr;

  void start_stream(uint64_t id) {
    if (mode == 1) {
      stream = id;
      nblk = 0;
      gauss_stored = false;
    }
  }
  void block(uint64_t &w0, uint64_t &w1) {
    uint32_t c[4] = {static_cast<uint32_t>(stream), static_cast<uint32_t>(stream >> 32),
                     static_cast<uint32_t>(nblk), static_cast<uint32_t>(nblk >> 32)};
    philox4x32_10(c, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
    w0 = static_cast<uint64_t>(c[0]) | (static_cast<uint64_t>(c[1]) << 32);
    w1 = static_cast<uint64_t>(c[2]) | (static_cast<uint64_t>(c[3]) << 32);
    ++nblk;
  }
  double uniform() {
    if (cnt) cnt->n_rng += 1;
    if (mode == 0) return word_to_open01(mt.next());
    uint64_t w0, w1;
    block(w0, w1);
    return word_to_open01(w0);
  }
  void uniform2(double &a, double &b) {
    if (cnt) cnt->n_rng += 2;
    if (mode == 0) {
      a = word_to_open01(mt.next());
      b = word_to_open01(mt.next());
      return;
    }
    uint64_t w0, w1;
    block(w0, w1);
    a = word_to_open01(w0);
    b = word_to_open01(w1);
  }
  // rand_gauss1 — random_mt.f90:964-988 (Marsaglia polar, returns v2*f, stores v1*f)
  double gauss() {
    if (gauss_stored) {
      gauss_stored = false;
      return gset;
    }
    double v1, v2, rsq;
    for (;;) {
      uniform2(v1, v2);
      v1 = 2.0 * v1 - 1.0;
      v2 = 2.0 * v2 - 1.0;
      rsq = v1 * v1 + v2 * v2;
      if (cnt) cnt->n_reject_iter += 1;
      if (rsq > 0.0 && rsq < 1.0) break;
    }
    rsq = std::sqrt(-2.0 * std::log(rsq) / rsq);
    gset = v1 * rsq;
    gauss_stored = true;
    return v2 * rsq;
  }
};

// ---------------------------------------------------------------------------
// voigt_seon2 — voigt_mod.f90:541-733.  Tables are 0-based here: h(k) -> H[k-1].
// ---------------------------------------------------------------------------
double voigt_seon2(double vin, double a) {
  const double one_sqrtPI = 0.56418958354775628695;
  double v = (vin < 0.0) ? -vin : vin;  // :686-690 (no abs())
  if (v < 1.0) {                        // :691-700
    double vv = v * 20.0;
    int k1 = static_cast<int>(vv) + 1, k2 = k1 + 1;
    double y1 = ORACLE_H0[k1 - 1] + a * (ORACLE_H1[k1 - 1] + a * ORACLE_H2[k1 - 1]);
    double y2 = ORACLE_H0[k2 - 1] + a * (ORACLE_H1[k2 - 1] + a * ORACLE_H2[k2 - 1]);
    return y1 + (y2 - y1) * (vv - static_cast<double>(k1 - 1));
  } else if (v < 5.0) {                 // :701-717
    double vv = v * 20.0;
    int k1 = static_cast<int>(vv), k2 = k1 + 1, k3 = k1 + 2;
    double y1 = ORACLE_H0[k1 - 1] + a * (ORACLE_H1[k1 - 1] + a * (ORACLE_H2[k1 - 1] + a * ORACLE_H3[k1 - 1]));
    double y2 = ORACLE_H0[k2 - 1] + a * (ORACLE_H1[k2 - 1] + a * (ORACLE_H2[k2 - 1] + a * ORACLE_H3[k2 - 1]));
    double y3 = ORACLE_H0[k3 - 1] + a * (ORACLE_H1[k3 - 1] + a * (ORACLE_H2[k3 - 1] + a * ORACLE_H3[k3 - 1]));
    double u1 = static_cast<double>(k1 - 1), u2 = static_cast<double>(k2 - 1), u3 = static_cast<double>(k3 - 1);
    return 0.5 * y1 * (vv - u2) * (vv - u3) - y2 * (vv - u1) * (vv - u3) + 0.5 * y3 * (vv - u1) * (vv - u2);
  } else if (v < 10.0) {                // :718-726
    double vv = v * 20.0;
    int k1 = static_cast<int>(vv) + 1, k2 = k1 + 1;
    double a2 = a * a;
    double y1 = a * (ORACLE_H1[k1 - 1] + a2 * ORACLE_H3[k1 - 1]);
    double y2 = a * (ORACLE_H1[k2 - 1] + a2 * ORACLE_H3[k2 - 1]);
    return y1 + (y2 - y1) * (vv - static_cast<double>(k1 - 1));
  }
  double v2 = 1.0 / (v * v);            // :727-731
  return one_sqrtPI * a * v2 * (1.0 + ((1.5 - a * a) + 3.75 * v2) * v2);
}

// ---------------------------------------------------------------------------
// photon_type — define.f90:80-111 (the members this path touches)
// ---------------------------------------------------------------------------
struct Photon {
  int64_t id = 0;
  double nscatt_gas = 0, nscatt_dust = 0;
  double x = 0, y = 0, z = 0;
  double kx = 0, ky = 0, kz = 1;
  double mx = 0, my = 0, mz = 0, nx = 0, ny = 0, nz = 0;
  int icell = 1, jcell = 1, kcell = 1;  // 1-based like the reference
  int icl = 0;                          // icell_clump: 0 = vacuum, > 0 = current clump (clump medium only)
  double xfreq = 0, xfreq_ref = 0, wgt = 1;
  double vfy_shear = 0;                 // photon%vfy_shear (define.f90:100): shearing-box velocity offset picked up at x wraps
  bool inside = true;
  double I = 1, Q = 0, U = 0, V = 0;
  double E1 = 1, E2 = 0, E3 = 1;
};

// Read-only view of one run's inputs with Fortran-style 1-based accessors.
struct World {
  const lart_grid *g;
  const lart_params *par;
  const lart_line *line;
  const lart_scatt_mat *sm;
  const lart_observer *obs;
  bool zonly;
  bool sym = false;  // par%xyz_symmetry: mirror planes at the lower faces (raytrace_car.f90:584-760, 1650-1949)
  int bcxy = 0, bcz = 0;  // BC_* of the x/y axes and of the z axis (setup.f90:952-976)
  const lart_clumps *cl = nullptr;  // par%use_clump_medium (clump_mod.f90)
  const lart_amr *amr = nullptr;    // par%use_amr_grid (octree_mod.f90): the photon's icell is a LEAF index, jcell = kcell = 1
  int atm = 0;         // LART_ATM_*: par%geometry = 'plane_atmosphere' / 'spherical_atmosphere' (setup.f90:959-987)
  bool shear = false;  // raytrace_to_tau_car_xyper_shear bound (setup.f90:967-969)
  bool jp = false;     // any of the CALCJ / CALCP / CALCPnew accumulators
  // spherical atmosphere: grid%mask == -1 marks the planet's molecular layer (grid_mod_car.f90:320-330)
  inline bool masked(int i, int j, int k) const { return atm == LART_ATM_SPHERICAL && g->mask && g->mask[idx(i, j, k)] == -1; }
  inline size_t idx(int i, int j, int k) const {
    if (amr) return static_cast<size_t>(i > 0 ? i - 1 : 0);  // (a photon outside the octree reads leaf 1, never used)
    return static_cast<size_t>(i - 1) + static_cast<size_t>(g->nx) * (static_cast<size_t>(j - 1) + static_cast<size_t>(g->ny) * static_cast<size_t>(k - 1));
  }
  inline double rhokap(int i, int j, int k) const { return amr ? amr->rhokap[idx(i, j, k)] : g->rhokap[idx(i, j, k)]; }
  inline double rhokapD(int i, int j, int k) const { return amr ? amr->rhokapD[idx(i, j, k)] : g->rhokapD[idx(i, j, k)]; }
  inline double voigt_a(int i, int j, int k) const { return amr ? amr->voigt_a[idx(i, j, k)] : g->voigt_a[idx(i, j, k)]; }
  inline double Dfreq(int i, int j, int k) const { return amr ? amr->Dfreq[idx(i, j, k)] : g->Dfreq[idx(i, j, k)]; }
  inline double vdotk(int i, int j, int k, double kx, double ky, double kz) const {
    size_t c = idx(i, j, k);
    if (amr) return amr->vfx[c] * kx + amr->vfy[c] * ky + amr->vfz[c] * kz;
    return g->vfx[c] * kx + g->vfy[c] * ky + g->vfz[c] * kz;
  }
  inline double xface(int i) const { return g->xface[i - 1]; }
  inline double yface(int j) const { return g->yface[j - 1]; }
  inline double zface(int k) const { return g->zface[k - 1]; }
  // calc_voigt1 — line_mod.f90:38-47 (signed x is passed through)
  inline double calc_voigt(double xfreq, int i, int j, int k) const { return voigt_seon2(xfreq, voigt_a(i, j, k)); }
  inline bool dust() const { return par->DGR > 0.0; }
};

// Tally sinks.  Small histograms are thread-private; the big cubes are shared
// between threads and updated with an atomic add (the reference keeps one
// private copy per MPI rank and reduces — memory_mod_mpi.f90:276-292,366-458 —
// which is the same sum).
inline void atomic_add(double *p, double v) {
  if (v == 0.0) return;  // adding an exact zero never changes a sum
  std::atomic_ref<double> r(*p);
  double old = r.load(std::memory_order_relaxed);
  while (!r.compare_exchange_weak(old, old + v, std::memory_order_relaxed)) {
  }
}

struct Tally {
  std::vector<double> Jout, Jin, Jabs, Jmu, Jabs2;
  double nscatt_gas = 0, nscatt_dust = 0;
  Counters cnt;
  lart_tallies *shared;  // cubes + allph written straight into the caller's buffers
};

// ---------------------------------------------------------------------------
// setup_traversal_car — raytrace_car.f90:12-87.  Returns true when the photon
// is already leaving the grid.  `zonly_eq` selects the `== zp` test of the
// z-only to_tau variant (raytrace_car.f90:2560) instead of `<= zp` (:31,:1185).
// ---------------------------------------------------------------------------
struct Trav {
  int istep, jstep, kstep;
  double tx, ty, tz, delx, dely, delz;
};

// Boundary of one axis: open (the photon leaves), mirror plane at the lower face (xyz/xy symmetry), periodic (xy_periodic).
enum { BC_OPEN = 0, BC_MIRROR = 1, BC_PERIODIC = 2 };

inline bool axis_setup(double &k, double &p, int &cell, int n, const double *face, double d, int &step, double &t, double &del,
                       bool eq_test, int bc = BC_OPEN, int c0 = 0, bool is_tau = false) {
  if (k > 0.0) {
    if (cell > n) {
      double f = face[cell - 1];
      if (eq_test ? (f == p) : (f <= p)) {
        // the periodic to_tau variant re-enters at the first face (:2293-2297); its to_edge variant returns (:1020)
        if (bc == BC_PERIODIC && is_tau) { cell = 1; p = face[0]; }
        else return true;
      }
    }
    step = 1;
    t = (face[cell] - p) / k;
    del = d / k;
  } else if (k < 0.0) {
    step = -1;
    del = -d / k;
    if (face[cell - 1] == p) {
      if (cell > 1) cell -= 1;
      else if (bc == BC_MIRROR) { cell = c0; step = 1; k = std::fabs(k); }  // reflected at the mirror plane (:1696-1700)
      else if (bc == BC_PERIODIC) { cell = n; p = face[n]; }                // wraps to the far side (:1030-1033)
      else return true;
    }
    t = (step == 1) ? (face[cell] - p) / k : (face[cell - 1] - p) / k;
  } else {
    step = 0;
    t = kHugest;
    del = kHugest;
  }
  return false;
}

inline bool setup_traversal(const World &w, double &xp, double &yp, double &zp, double &kx, double &ky, double &kz,
                            int &ic, int &jc, int &kc, Trav &t, bool is_tau) {
  const lart_grid &g = *w.g;
  if (w.zonly) {
    // raytrace_car.f90:1184-1204 / :2559-2583 — only the z axis is walked; the to_tau variant tests `== zp`
    t.istep = t.jstep = 0;
    t.tx = t.ty = kHugest;
    t.delx = t.dely = kHugest;
    return axis_setup(kz, zp, kc, g.nz, g.zface, g.dz, t.kstep, t.tz, t.delz, is_tau);
  }
  // the to_tau variants of the folded / periodic grids test `==` (:1681, :1993, :2293), all to_edge variants `<=`
  const bool eq = (w.bcxy != BC_OPEN || w.bcz != BC_OPEN) && is_tau;
  if (axis_setup(kx, xp, ic, g.nx, g.xface, g.dx, t.istep, t.tx, t.delx, eq, w.bcxy, g.i0, is_tau)) return true;
  if (axis_setup(ky, yp, jc, g.ny, g.yface, g.dy, t.jstep, t.ty, t.dely, eq, w.bcxy, g.j0, is_tau)) return true;
  if (axis_setup(kz, zp, kc, g.nz, g.zface, g.dz, t.kstep, t.tz, t.delz, eq, w.bcz, g.k0, is_tau)) return true;
  return false;
}

// index step along one axis: reflected at a mirror plane (:1791-1799), wrapped around a periodic box (:2383-2385).
// Returns false when the photon leaves the grid.
inline bool advance_axis(int bc, int &cell, int &step, double &k, int n, int c0) {
  cell += step;
  if (cell >= 1 && cell <= n) return true;
  if (bc == BC_PERIODIC) { cell = (cell < 1) ? n : 1; return true; }
  if (bc == BC_MIRROR && cell < 1) { cell = c0; step = 1; k = -k; return true; }
  return false;
}

// minloc([tx,ty,tz],dim=1) — first minimum wins (raytrace_car.f90:476,1506)
inline int minloc3(double tx, double ty, double tz) {
  if (tx <= ty && tx <= tz) return 1;
  if (ty <= tz) return 2;
  return 3;
}

// ---------------------------------------------------------------------------
// raytrace_to_edge_car — raytrace_car.f90:410-508; _zonly :1138-1234.
// Optional trace of visited cells (0-based linear index) for parity tests.
// ---------------------------------------------------------------------------
// With a shearing box or an atmosphere model the reference leaves raytrace_to_edge bound to the plain open-box routine
// (setup.f90:947-950 is not overridden at :959-969; spherical atmospheres bind raytrace_to_edge_car_atmosphere, :984, even
// with xy_symmetry): the walk then ignores the periodic / z-only / mirror binding of the to_tau routine.
inline World edge_world(const World &w0) {
  World w = w0;
  if (w.shear || w.atm) { w.zonly = false; w.bcxy = 0; w.bcz = 0; }
  return w;
}
double raytrace_to_edge(const World &w0, const Photon &p0, Counters *cnt, int *nsteps_out = nullptr,
                        int trace_cap = 0, int32_t *trace = nullptr) {
  const World w = edge_world(w0);
  const lart_grid &g = *w.g;
  double xp = p0.x, yp = p0.y, zp = p0.z, kx = p0.kx, ky = p0.ky, kz = p0.kz;
  int ic = p0.icell, jc = p0.jcell, kc = p0.kcell;
  double tau = 0.0, d = 0.0;
  int nsteps = 0;
  Trav t;
  if (setup_traversal(w, xp, yp, zp, kx, ky, kz, ic, jc, kc, t, false)) {
    if (nsteps_out) *nsteps_out = 0;
    return tau;
  }
  int io = ic, jo = jc, ko = kc;
  double u1 = w.vdotk(io, jo, ko, kx, ky, kz);
  double xfreq = p0.xfreq;
  for (;;) {
    if (w.masked(ic, jc, kc)) { tau = std::numeric_limits<double>::infinity(); break; }  // raytrace_car.f90:3729-3733
    double rhokap = w.rhokap(ic, jc, kc) * w.calc_voigt(xfreq, ic, jc, kc);
    if (w.dust()) rhokap += w.rhokapD(ic, jc, kc);
    if (trace && nsteps < trace_cap) trace[nsteps] = static_cast<int32_t>(w.idx(ic, jc, kc));
    ++nsteps;
    int m = w.zonly ? 3 : minloc3(t.tx, t.ty, t.tz);
    if (m == 1) {
      tau += (t.tx - d) * rhokap;
      d = t.tx;
      if (!advance_axis(w.bcxy, ic, t.istep, kx, g.nx, g.i0)) break;
      t.tx += t.delx;
    } else if (m == 2) {
      tau += (t.ty - d) * rhokap;
      d = t.ty;
      if (!advance_axis(w.bcxy, jc, t.jstep, ky, g.ny, g.j0)) break;
      t.ty += t.dely;
    } else {
      tau += (t.tz - d) * rhokap;
      d = t.tz;
      if (!advance_axis(w.bcz, kc, t.kstep, kz, g.nz, g.k0)) break;
      t.tz += t.delz;
    }
    if (tau >= kTauHuge) break;
    double u2 = w.vdotk(ic, jc, kc, kx, ky, kz);
    xfreq = (xfreq + u1) * w.Dfreq(io, jo, ko) / w.Dfreq(ic, jc, kc) - u2;
    io = ic; jo = jc; ko = kc;
    u1 = u2;
  }
  if (cnt) cnt->n_cellsteps += nsteps;
  if (nsteps_out) *nsteps_out = nsteps;
  return tau;
}

// add_to_Jmu — raytrace_car.f90:4049-4064
inline int jmu_bin(const lart_params &par, double kz) {
  double mu = kz;
  if (par.xyz_symmetry) mu = std::fabs(mu);  // raytrace_car.f90:4058
  int imu = static_cast<int>(std::floor((mu - par.mu_min) / par.dmu)) + 1;
  if (imu < 1) imu = 1;
  if (imu > par.nmu) imu = par.nmu;
  return imu;
}

// ---------------------------------------------------------------------------
// CALCJ / CALCP / CALCPnew accumulators (compile-time options of the reference, run-time flags here).
// jp_bin: the bin of cell (i,j,k) in the P arrays by par%geometry_JPa — 3: the cell, 2: (ind_cyl(i,j), k), 1: ind_sph(i,j,k),
// -1: k; -1 when the cell takes no deposit (guards of raytrace_car.f90:3989-3991, :3996).  The reference does not range-check
// ind_sph; a bin outside 1..nr is skipped here instead of written out of bounds.
// ---------------------------------------------------------------------------
inline long long jp_bin(const World &w, int i, int j, int k) {
  const lart_grid &g = *w.g;
  if (!(i > 0 && i <= g.nx && j > 0 && j <= g.ny && k > 0 && k <= g.nz)) return -1;
  if (!(w.rhokap(i, j, k) > 0.0)) return -1;
  switch (g.geometry_JPa) {
    case 3: return static_cast<long long>(w.idx(i, j, k));
    case 2: {
      const int ir = g.ind_cyl[(i - 1) + static_cast<size_t>(g.nx) * (j - 1)];
      return (ir >= 1 && ir <= g.nr) ? (ir - 1) + static_cast<long long>(g.nr) * (k - 1) : -1;
    }
    case 1: {
      const int ir = g.ind_sph[w.idx(i, j, k)];
      return (ir >= 1 && ir <= g.nr) ? ir - 1 : -1;
    }
    default: return k - 1;
  }
}
// add_to_J — raytrace_car.f90:3979-4011
inline void add_to_J(const World &w, const Photon &ph, Tally *tl, int i, int j, int k, double del) {
  if (!w.par->calc_J || !tl || !tl->shared->J) return;
  const lart_grid &g = *w.g;
  if (!(i > 0 && i <= g.nx && j > 0 && j <= g.ny && k > 0 && k <= g.nz)) return;
  const double xref = ph.xfreq * (w.Dfreq(i, j, k) / g.Dfreq_ref);
  const int ix = static_cast<int>(std::floor((xref - g.xfreq_min) / g.dxfreq)) + 1;
  if (!(ix > 0 && ix <= g.nxfreq)) return;
  const long long b = jp_bin(w, i, j, k);
  if (b >= 0) atomic_add(tl->shared->J + (ix - 1) + static_cast<size_t>(g.nxfreq) * b, del * ph.wgt);
}
// add_to_Pnew — raytrace_car.f90:4015-4045
inline void add_to_Pnew(const World &w, const Photon &ph, Tally *tl, int i, int j, int k, double dtauH) {
  if (!w.par->calc_Pnew || !tl || !tl->shared->Pnew) return;
  const long long b = jp_bin(w, i, j, k);
  if (b < 0) return;
  const double rhokap = w.rhokap(i, j, k) * w.Dfreq(i, j, k) / w.line->cross0;
  atomic_add(tl->shared->Pnew + b, dtauH * ph.wgt / rhokap);
}
// add_to_Pa — scattering_car.f90:829-860
inline void add_to_Pa(const World &w, const Photon &ph, Tally &tl) {
  if (!w.par->calc_P || !tl.shared->Pa) return;
  const int i = ph.icell, j = ph.jcell, k = ph.kcell;
  const long long b = jp_bin(w, i, j, k);
  if (b < 0) return;
  const double rhokap = w.rhokap(i, j, k) * w.Dfreq(i, j, k) / w.line->cross0;
  atomic_add(tl.shared->Pa + b, ph.wgt / rhokap);
}

// ---------------------------------------------------------------------------
// raytrace_to_tau_car — raytrace_car.f90:1425-1648; _zonly :2519-2675; the shearing box _xyper_shear :2677-2954; the
// atmosphere models _zonly_atmosphere :2956-3117, _atmosphere :3119-3338, _xysym_atmosphere :3340-3663.
// `tl` may be null (unit-level batch calls: no tallies).
// ---------------------------------------------------------------------------
void raytrace_to_tau(const World &w, Photon &ph, double tau_in, Tally *tl, Counters *cnt, int *nsteps_out = nullptr) {
  const lart_grid &g = *w.g;
  double xp = ph.x, yp = ph.y, zp = ph.z, kx = ph.kx, ky = ph.ky, kz = ph.kz;
  int ic = ph.icell, jc = ph.jcell, kc = ph.kcell;
  int nsteps = 0;
  Trav t;
  if (setup_traversal(w, xp, yp, zp, kx, ky, kz, ic, jc, kc, t, true)) {
    ph.inside = false;  // :1469-1472 — returns before any tally
    if (nsteps_out) *nsteps_out = 0;
    return;
  }
  double tau = 0.0, d = 0.0, del = 0.0, dtauH = 0.0;
  int io = ic, jo = jc, ko = kc;
  // the shearing box adds photon%vfy_shear to the cell's vfy in every line-of-sight velocity (:2809, :2905, :2932)
  auto ulos = [&](int i, int j, int k) {
    if (!w.shear) return w.vdotk(i, j, k, kx, ky, kz);
    const size_t c = w.idx(i, j, k);
    return g.vfx[c] * kx + (g.vfy[c] + ph.vfy_shear) * ky + g.vfz[c] * kz;
  };
  double u1 = ulos(io, jo, ko);
  bool destroyed = false;
  while (ph.inside) {
    if (w.masked(ic, jc, kc)) { destroyed = true; break; }  // :3186-3190 the planet's molecular layer destroys the photon
    double rhokapH = w.rhokap(ic, jc, kc) * w.calc_voigt(ph.xfreq, ic, jc, kc);
    double rhokap = rhokapH;
    if (w.dust()) rhokap += w.rhokapD(ic, jc, kc);
    ++nsteps;
    int m = w.zonly ? 3 : minloc3(t.tx, t.ty, t.tz);
    double tnext = (m == 1) ? t.tx : (m == 2) ? t.ty : t.tz;
    del = tnext - d;
    tau += del * rhokap;
    dtauH = del * rhokapH;
    d = tnext;
    if (tau >= tau_in) {  // :1513-1524
      if (rhokap > 0.0) {
        double d_overshoot = (tau - tau_in) / rhokap;
        d = d - d_overshoot;
        del = del - d_overshoot;
        dtauH = del * rhokapH;
      }
      // the ORIGINAL direction even after a reflection; the point is mirrored back below (:1936-1941)
      xp = xp + d * ph.kx;
      yp = yp + d * ph.ky;
      zp = zp + d * ph.kz;
      break;
    }
    if (m == 1) {
      const int before = ic + t.istep;
      if (!advance_axis(w.bcxy, ic, t.istep, kx, g.nx, g.i0)) { ph.inside = false; break; }
      if (w.shear) {  // :2842-2850 the photon re-enters through the opposite x face of the sheared neighbour box
        if (before < 1) ph.vfy_shear = ph.vfy_shear - w.par->Omega;
        if (before > g.nx) ph.vfy_shear = ph.vfy_shear + w.par->Omega;
      }
      t.tx += t.delx;
    } else if (m == 2) {
      if (!advance_axis(w.bcxy, jc, t.jstep, ky, g.ny, g.j0)) { ph.inside = false; break; }
      t.ty += t.dely;
    } else {
      if (!advance_axis(w.bcz, kc, t.kstep, kz, g.nz, g.k0)) { ph.inside = false; break; }
      t.tz += t.delz;
    }
    if (w.jp) {  // :1579-1584
      add_to_J(w, ph, tl, io, jo, ko, del);
      add_to_Pnew(w, ph, tl, io, jo, ko, dtauH);
    }
    double u2 = ulos(ic, jc, kc);
    ph.xfreq = (ph.xfreq + u1) * w.Dfreq(io, jo, ko) / w.Dfreq(ic, jc, kc) - u2;  // :1586-1589
    io = ic; jo = jc; ko = kc;
    u1 = u2;
  }
  if (!ph.inside) { ic = io; jc = jo; kc = ko; }  // :1598-1602
  // :1604-1609 the last segment.  Deviation: for a photon destroyed by the mask the reference calls add_to_J / add_to_Pnew
  // once more with the PREVIOUS segment's del and dtauH, now credited to the masked cell (:3296-3301; the variables are
  // undefined when the first cell is masked) — a stale-variable slip that neither side of the parity tests reproduces.
  if (w.jp && !destroyed) {
    add_to_J(w, ph, tl, ic, jc, kc, del);
    add_to_Pnew(w, ph, tl, ic, jc, kc, dtauH);
  }
  if (!ph.inside || destroyed) {
    // :1613-1623 — fluid frame -> lab frame, reference Doppler units, Jout bin; a destroyed photon is binned the same way
    // into Jabs2 (:3316-3327), and so is one that leaves a plane atmosphere through its bottom cell (:3099-3107)
    double ue = ulos(ic, jc, kc);
    ph.xfreq = ph.xfreq + ue;
    ph.xfreq_ref = ph.xfreq * (w.Dfreq(ic, jc, kc) / g.Dfreq_ref);
    if (tl) {
      int ix = static_cast<int>(std::floor((ph.xfreq_ref - g.xfreq_min) / g.dxfreq)) + 1;
      if (ix >= 1 && ix <= g.nxfreq) {
        if (destroyed || (w.atm == LART_ATM_PLANE && !(kc > 1))) {
          tl->Jabs2[ix - 1] += ph.wgt;
        } else {
          tl->Jout[ix - 1] += ph.wgt;
          if (w.par->save_Jmu) tl->Jmu[(ix - 1) + static_cast<size_t>(g.nxfreq) * (jmu_bin(*w.par, ph.kz) - 1)] += ph.wgt;
        }
      }
    }
    ph.inside = false;
  }
  if (w.bcxy == BC_MIRROR) {  // :1936-1947, :2236-2245 (d is the whole path when the photon left the grid)
    if (!ph.inside) { xp = ph.x + d * ph.kx; yp = ph.y + d * ph.ky; zp = ph.z + d * ph.kz; }
    if (xp < g.xmin) xp = -xp;
    if (yp < g.ymin) yp = -yp;
    if (w.bcz == BC_MIRROR && zp < g.zmin) zp = -zp;
  } else if (w.bcxy == BC_PERIODIC) {  // folded back into the box (:2506-2508)
    const double xrange = g.xmax - g.xmin, yrange = g.ymax - g.ymin;
    xp = xp - std::floor((xp - g.xmin) / xrange) * xrange;
    yp = yp - std::floor((yp - g.ymin) / yrange) * yrange;
  }
  ph.x = xp; ph.y = yp; ph.z = zp;
  if (w.bcxy == BC_MIRROR) { ph.kx = kx; ph.ky = ky; ph.kz = kz; }
  if (!w.zonly) { ph.icell = ic; ph.jcell = jc; }
  ph.kcell = kc;
  if (cnt) cnt->n_cellsteps += nsteps;
  if (nsteps_out) *nsteps_out = nsteps;
}

// ---------------------------------------------------------------------------
// Samplers
// ===========================================================================
// Clump medium — clump_mod.f90 / raytrace_clump.f90 (non-overlapping populations)
// ===========================================================================
constexpr double kTauHugeClump = 745.2;  // raytrace_clump.f90:59

// voigt_clump / kappa_clump / ulos_clump — clump_mod.f90:130-190 (line_type 1)
inline double voigt_clump(const lart_clumps &c, double xfreq, int64_t icl) {
  double xloc = xfreq * (c.Dfreq_ref / c.Dfreq[icl - 1]);
  return voigt_seon2(xloc, c.voigt_a[icl - 1]);
}
inline double kappa_clump(const World &w, double xfreq, int64_t icl) {
  const lart_clumps &c = *w.cl;
  double kap = c.rhokap[icl - 1] * voigt_clump(c, xfreq, icl);
  if (w.par->DGR > 0.0) kap = kap + c.rhokapD[icl - 1];
  return kap;
}
inline double ulos_clump(const lart_clumps &c, int64_t icl, double kx, double ky, double kz) {
  return (c.vx[icl - 1] * kx + c.vy[icl - 1] * ky + c.vz[icl - 1] * kz) * (c.Dfreq[icl - 1] / c.Dfreq_ref);
}
// ray_sphere_isect — clump_mod.f90:1369-1390
inline bool ray_sphere_isect(const lart_clumps &c, double ox, double oy, double oz, double kx, double ky, double kz, int64_t icl,
                             double &t_entry, double &t_exit) {
  double rx = ox - c.x[icl - 1], ry = oy - c.y[icl - 1], rz = oz - c.z[icl - 1];
  double b = rx * kx + ry * ky + rz * kz;
  double disc = b * b - (rx * rx + ry * ry + rz * rz) + c.radius[icl - 1] * c.radius[icl - 1];
  if (disc < 0.0) { t_entry = 0.0; t_exit = 0.0; return false; }
  disc = std::sqrt(disc);
  t_entry = -b - disc;
  t_exit = -b + disc;
  return t_exit > 0.0;
}
// find_next_clump — clump_mod.f90:1393-1506: DDA through the CSR grid to the nearest clump hit, 0 < t <= t_max
inline bool find_next_clump(const lart_clumps &c, double xp, double yp, double zp, double kx, double ky, double kz, int64_t skip_icl,
                            double t_max, double &t_entry, double &t_exit, int64_t &icl_found, long long *ncells = nullptr) {
  const double inv_dx = 1.0 / c.cg_dx, inv_dy = 1.0 / c.cg_dy, inv_dz = 1.0 / c.cg_dz;
  double best_te = kHugest, best_tx2 = 0.0, d = 0.0;
  int64_t best_icl = 0;
  int ci = std::max(0, std::min(c.cgx - 1, static_cast<int>((xp - c.cg_xmin) * inv_dx)));
  int cj = std::max(0, std::min(c.cgy - 1, static_cast<int>((yp - c.cg_ymin) * inv_dy)));
  int ck = std::max(0, std::min(c.cgz - 1, static_cast<int>((zp - c.cg_zmin) * inv_dz)));
  int si, sj, sk;
  double tx, ty, tz, delx, dely, delz;
  auto axis = [](double k, double p, int cc, double lo, double dd, int &st, double &t, double &del) {
    if (k > 0.0) { st = 1; t = ((lo + static_cast<double>(cc + 1) * dd) - p) / k; del = dd / k; }
    else if (k < 0.0) { st = -1; t = ((lo + static_cast<double>(cc) * dd) - p) / k; del = -dd / k; }
    else { st = 0; t = kHugest; del = kHugest; }
  };
  axis(kx, xp, ci, c.cg_xmin, c.cg_dx, si, tx, delx);
  axis(ky, yp, cj, c.cg_ymin, c.cg_dy, sj, ty, dely);
  axis(kz, zp, ck, c.cg_zmin, c.cg_dz, sk, tz, delz);
  for (;;) {
    if (d > best_te || d > t_max) break;
    if (ncells) ++*ncells;
    const size_t icell = static_cast<size_t>(ci) + static_cast<size_t>(c.cgx) * (cj + static_cast<size_t>(c.cgy) * ck);
    for (int32_t ip = c.cg_start[icell]; ip < c.cg_start[icell + 1]; ++ip) {
      const int64_t icl = c.cg_list[ip - 1];
      if (icl == skip_icl) continue;
      double te, tx2;
      bool hit = ray_sphere_isect(c, xp, yp, zp, kx, ky, kz, icl, te, tx2);
      if (hit && tx2 > 0.0 && te < best_te && (te > 0.0 || icl != skip_icl)) { best_te = te; best_tx2 = tx2; best_icl = icl; }
    }
    if (tx <= ty && tx <= tz) { d = tx; ci += si; if (ci < 0 || ci >= c.cgx) break; tx += delx; }
    else if (ty <= tz) { d = ty; cj += sj; if (cj < 0 || cj >= c.cgy) break; ty += dely; }
    else { d = tz; ck += sk; if (ck < 0 || ck >= c.cgz) break; tz += delz; }
  }
  if (best_icl > 0 && best_te <= t_max) {
    t_entry = best_te; t_exit = std::min(best_tx2, t_max); icl_found = best_icl;
    return true;
  }
  return false;
}
// clump_exit_dist / sphere_exit_dist — clump_mod.f90:1509-1540
inline double clump_exit_dist(const lart_clumps &c, double xp, double yp, double zp, double kx, double ky, double kz, int64_t icl) {
  double rx = xp - c.x[icl - 1], ry = yp - c.y[icl - 1], rz = zp - c.z[icl - 1];
  double b = rx * kx + ry * ky + rz * kz;
  double disc = b * b - (rx * rx + ry * ry + rz * rz) + c.radius[icl - 1] * c.radius[icl - 1];
  if (disc < 0.0) disc = 0.0;
  return std::max(0.0, -b + std::sqrt(disc));
}
inline double sphere_exit_dist(const lart_clumps &c, double xp, double yp, double zp, double kx, double ky, double kz) {
  double b = xp * kx + yp * ky + zp * kz;
  double disc = b * b - (xp * xp + yp * yp + zp * zp) + c.sphere_R * c.sphere_R;
  if (disc < 0.0) disc = 0.0;
  return std::max(0.0, -b + std::sqrt(disc));
}
// active_set_at_point, first hit — clump_mod.f90:1595-1634
inline int64_t clump_at_point(const lart_clumps &c, double xp, double yp, double zp) {
  const double inv_dx = 1.0 / c.cg_dx, inv_dy = 1.0 / c.cg_dy, inv_dz = 1.0 / c.cg_dz;
  int ci = std::max(0, std::min(c.cgx - 1, static_cast<int>((xp - c.cg_xmin) * inv_dx)));
  int cj = std::max(0, std::min(c.cgy - 1, static_cast<int>((yp - c.cg_ymin) * inv_dy)));
  int ck = std::max(0, std::min(c.cgz - 1, static_cast<int>((zp - c.cg_zmin) * inv_dz)));
  for (int k = std::max(0, ck - 1); k <= std::min(c.cgz - 1, ck + 1); ++k)
    for (int j = std::max(0, cj - 1); j <= std::min(c.cgy - 1, cj + 1); ++j)
      for (int i = std::max(0, ci - 1); i <= std::min(c.cgx - 1, ci + 1); ++i) {
        const size_t icell = static_cast<size_t>(i) + static_cast<size_t>(c.cgx) * (j + static_cast<size_t>(c.cgy) * k);
        for (int32_t ip = c.cg_start[icell]; ip < c.cg_start[icell + 1]; ++ip) {
          const int64_t icl = c.cg_list[ip - 1];
          double rx = xp - c.x[icl - 1], ry = yp - c.y[icl - 1], rz = zp - c.z[icl - 1];
          if (rx * rx + ry * ry + rz * rz <= c.radius[icl - 1] * c.radius[icl - 1]) return icl;
        }
      }
  return 0;
}
// update_cell_idx — raytrace_clump.f90:68-75
inline void update_cell_idx(const World &w, Photon &ph) {
  const lart_grid &g = *w.g;
  ph.icell = std::max(1, std::min(g.nx, static_cast<int>(std::floor((ph.x - g.xmin) / g.dx)) + 1));
  ph.jcell = std::max(1, std::min(g.ny, static_cast<int>(std::floor((ph.y - g.ymin) / g.dy)) + 1));
  ph.kcell = std::max(1, std::min(g.nz, static_cast<int>(std::floor((ph.z - g.zmin) / g.dz)) + 1));
}
// raytrace_to_edge_clump (:205-270; tau_max <= 0) and raytrace_to_edge_clump_capped (:494-533)
double raytrace_to_edge_clump(const World &w, const Photon &p0, double tau_max, Counters *cnt, int *nclumps = nullptr) {
  const lart_clumps &c = *w.cl;
  const bool capped = tau_max > 0.0;
  double tau = 0.0;
  double xp = p0.x, yp = p0.y, zp = p0.z;
  const double kx = p0.kx, ky = p0.ky, kz = p0.kz;
  double xfreq = p0.xfreq;
  int64_t icl_cur = p0.icl;
  int ncl = 0;
  long long ncell = 0;
  auto done = [&]() { if (cnt) cnt->n_cellsteps += ncell; if (nclumps) *nclumps = ncl; return tau; };
  if (icl_cur > 0) {
    double t_seg = clump_exit_dist(c, xp, yp, zp, kx, ky, kz, icl_cur);
    double kap = kappa_clump(w, xfreq, icl_cur);
    tau = tau + kap * t_seg;
    ++ncl;
    if (capped && tau >= tau_max) return done();
    xp = xp + t_seg * kx; yp = yp + t_seg * ky; zp = zp + t_seg * kz;
    xfreq = xfreq + ulos_clump(c, icl_cur, kx, ky, kz);
    if (xp * xp + yp * yp + zp * zp >= c.sphere_R * c.sphere_R) return done();
  }
  for (;;) {
    double t_sp = sphere_exit_dist(c, xp, yp, zp, kx, ky, kz);
    if (t_sp <= 0.0) break;
    double te, tx2;
    int64_t icl_found = 0;
    if (!find_next_clump(c, xp, yp, zp, kx, ky, kz, icl_cur, t_sp, te, tx2, icl_found, &ncell)) break;
    te = std::max(0.0, te);
    xp = xp + te * kx; yp = yp + te * ky; zp = zp + te * kz;
    double u_los = ulos_clump(c, icl_found, kx, ky, kz);
    xfreq = xfreq - u_los;
    double t_seg = clump_exit_dist(c, xp, yp, zp, kx, ky, kz, icl_found);
    double kap = kappa_clump(w, xfreq, icl_found);
    tau = tau + kap * t_seg;
    ++ncl;
    if (capped && tau >= tau_max) return done();
    xp = xp + t_seg * kx; yp = yp + t_seg * ky; zp = zp + t_seg * kz;
    xfreq = xfreq + u_los;
    icl_cur = icl_found;
    if (xp * xp + yp * yp + zp * zp >= c.sphere_R * c.sphere_R) break;
  }
  return done();
}
// raytrace_to_tau_clump — raytrace_clump.f90:83-201.  `tl` may be null (unit-level calls).
// Deviation: upstream leaves photon%xfreq_ref unset on this path (allph%xfreq2 is then stale); here it is the
// escape frequency that is binned into Jout.
void raytrace_to_tau_clump(const World &w, Photon &ph, double tau_in, Tally *tl, Counters *cnt) {
  const lart_clumps &c = *w.cl;
  const lart_grid &g = *w.g;
  const double kx = ph.kx, ky = ph.ky, kz = ph.kz;
  double tau_rem = tau_in;
  int64_t last_icl = 0;
  long long ncell = 0;
  auto escape = [&]() {
    ph.inside = false;
    ph.xfreq_ref = ph.xfreq;
    if (tl) {
      int ix = static_cast<int>(std::floor((ph.xfreq - g.xfreq_min) / g.dxfreq)) + 1;
      if (ix >= 1 && ix <= g.nxfreq) {
        tl->Jout[ix - 1] += ph.wgt;
        if (w.par->save_Jmu) tl->Jmu[(ix - 1) + static_cast<size_t>(g.nxfreq) * (jmu_bin(*w.par, ph.kz) - 1)] += ph.wgt;
      }
    }
  };
  while (ph.inside) {
    if (ph.icl > 0) {
      const int64_t icl = ph.icl;
      double t_seg = clump_exit_dist(c, ph.x, ph.y, ph.z, kx, ky, kz, icl);
      double kap = kappa_clump(w, ph.xfreq, icl);
      if (tau_rem <= kap * t_seg) {  // scatter inside this clump
        double ds = tau_rem / std::max(kap, 2.2250738585072014e-308);
        ph.x = ph.x + ds * kx; ph.y = ph.y + ds * ky; ph.z = ph.z + ds * kz;
        update_cell_idx(w, ph);
        break;
      }
      tau_rem = tau_rem - kap * t_seg;
      ph.x = ph.x + t_seg * kx; ph.y = ph.y + t_seg * ky; ph.z = ph.z + t_seg * kz;
      ph.xfreq = ph.xfreq + ulos_clump(c, icl, kx, ky, kz);
      last_icl = icl;
      ph.icl = 0;
      if (ph.x * ph.x + ph.y * ph.y + ph.z * ph.z >= c.sphere_R * c.sphere_R) {
        update_cell_idx(w, ph);
        escape();
        break;
      }
    } else {
      double t_sp = sphere_exit_dist(c, ph.x, ph.y, ph.z, kx, ky, kz);
      if (t_sp <= 0.0) { escape(); break; }
      double te, tx2;
      int64_t icl_found = 0;
      if (find_next_clump(c, ph.x, ph.y, ph.z, kx, ky, kz, last_icl, t_sp, te, tx2, icl_found, &ncell)) {
        te = std::max(0.0, te);
        ph.x = ph.x + te * kx; ph.y = ph.y + te * ky; ph.z = ph.z + te * kz;
        ph.xfreq = ph.xfreq - ulos_clump(c, icl_found, kx, ky, kz);
        last_icl = 0;
        ph.icl = static_cast<int>(icl_found);
        update_cell_idx(w, ph);
      } else {
        ph.x = ph.x + t_sp * kx; ph.y = ph.y + t_sp * ky; ph.z = ph.z + t_sp * kz;
        update_cell_idx(w, ph);
        escape();
        break;
      }
    }
  }
  if (cnt) cnt->n_cellsteps += ncell;
}
// ---------------------------------------------------------------------------
// Overlapping clump populations (has_overlap; setup_clump_overlap, setup.f90:1051-1081): the event walk of
// raytrace_clump.f90:608-920.  photon%xfreq stays in the GLOBAL frame during the walk; every clump containing a point
// contributes its opacity at its own frame's frequency.
// ---------------------------------------------------------------------------
constexpr int kMaxEvt = 2048;  // MAX_EVT, raytrace_clump.f90:683
struct ClumpEvents {
  double t[kMaxEvt];
  int64_t icl[kMaxEvt];
  int type[kMaxEvt];  // +1 ENTER, -1 EXIT
  int n = 0;
};
// active_set_at_point — clump_mod.f90:1595-1634 (all clumps containing the point, first-seen order, no duplicates)
inline int active_set_at_point(const lart_clumps &c, double xp, double yp, double zp, int64_t *active, int cap) {
  const double inv_dx = 1.0 / c.cg_dx, inv_dy = 1.0 / c.cg_dy, inv_dz = 1.0 / c.cg_dz;
  int ci = std::max(0, std::min(c.cgx - 1, static_cast<int>((xp - c.cg_xmin) * inv_dx)));
  int cj = std::max(0, std::min(c.cgy - 1, static_cast<int>((yp - c.cg_ymin) * inv_dy)));
  int ck = std::max(0, std::min(c.cgz - 1, static_cast<int>((zp - c.cg_zmin) * inv_dz)));
  int na = 0;
  for (int k = std::max(0, ck - 1); k <= std::min(c.cgz - 1, ck + 1); ++k)
    for (int j = std::max(0, cj - 1); j <= std::min(c.cgy - 1, cj + 1); ++j)
      for (int i = std::max(0, ci - 1); i <= std::min(c.cgx - 1, ci + 1); ++i) {
        const size_t icell = static_cast<size_t>(i) + static_cast<size_t>(c.cgx) * (j + static_cast<size_t>(c.cgy) * k);
        for (int32_t ip = c.cg_start[icell]; ip < c.cg_start[icell + 1]; ++ip) {
          const int64_t icl = c.cg_list[ip - 1];
          double rx = xp - c.x[icl - 1], ry = yp - c.y[icl - 1], rz = zp - c.z[icl - 1];
          if (rx * rx + ry * ry + rz * rz <= c.radius[icl - 1] * c.radius[icl - 1]) {
            bool seen = false;
            for (int m = 0; m < na; ++m) seen = seen || active[m] == icl;
            if (!seen && na < cap) active[na++] = icl;
          }
        }
      }
  return na;
}
// collect_ray_events_overlap — clump_mod.f90:1639-1760: every ENTER/EXIT along the ray up to t_max, sorted by t
inline void collect_ray_events_overlap(const lart_clumps &c, double xp, double yp, double zp, double kx, double ky, double kz,
                                       double t_max, ClumpEvents &ev) {
  const double inv_dx = 1.0 / c.cg_dx, inv_dy = 1.0 / c.cg_dy, inv_dz = 1.0 / c.cg_dz;
  ev.n = 0;
  int ci = std::max(0, std::min(c.cgx - 1, static_cast<int>((xp - c.cg_xmin) * inv_dx)));
  int cj = std::max(0, std::min(c.cgy - 1, static_cast<int>((yp - c.cg_ymin) * inv_dy)));
  int ck = std::max(0, std::min(c.cgz - 1, static_cast<int>((zp - c.cg_zmin) * inv_dz)));
  int si, sj, sk;
  double tx, ty, tz, delx, dely, delz, d = 0.0;
  auto axis = [](double k, double p, int cc, double lo, double dd, int &st, double &t, double &del) {
    if (k > 0.0) { st = 1; del = dd / k; t = (lo + static_cast<double>(cc + 1) * dd - p) / k; }
    else if (k < 0.0) { st = -1; del = -dd / k; t = (lo + static_cast<double>(cc) * dd - p) / k; }
    else { st = 0; del = kHugest; t = kHugest; }
  };
  axis(kx, xp, ci, c.cg_xmin, c.cg_dx, si, tx, delx);
  axis(ky, yp, cj, c.cg_ymin, c.cg_dy, sj, ty, dely);
  axis(kz, zp, ck, c.cg_zmin, c.cg_dz, sk, tz, delz);
  for (;;) {
    if (d > t_max) break;
    if (ci < 0 || ci >= c.cgx || cj < 0 || cj >= c.cgy || ck < 0 || ck >= c.cgz) break;
    const size_t icell = static_cast<size_t>(ci) + static_cast<size_t>(c.cgx) * (cj + static_cast<size_t>(c.cgy) * ck);
    for (int32_t ip = c.cg_start[icell]; ip < c.cg_start[icell + 1]; ++ip) {
      const int64_t icl = c.cg_list[ip - 1];
      double te, tx2;
      if (!ray_sphere_isect(c, xp, yp, zp, kx, ky, kz, icl, te, tx2)) continue;
      if (tx2 <= 0.0 || te > t_max) continue;
      bool dup = false;
      for (int ie = 0; ie < ev.n; ++ie) dup = dup || ev.icl[ie] == icl;
      if (dup || ev.n + 2 > kMaxEvt) continue;
      if (te > 0.0) { ev.t[ev.n] = te; ev.icl[ev.n] = icl; ev.type[ev.n] = +1; ++ev.n; }
      ev.t[ev.n] = std::min(tx2, t_max); ev.icl[ev.n] = icl; ev.type[ev.n] = -1; ++ev.n;
    }
    if (tx <= ty && tx <= tz) { d = tx; tx += delx; ci += si; }
    else if (ty <= tz) { d = ty; ty += dely; cj += sj; }
    else { d = tz; tz += delz; ck += sk; }
  }
  for (int ie = 1; ie < ev.n; ++ie) {  // insertion sort, stable for equal t
    double tt = ev.t[ie]; int64_t ii = ev.icl[ie]; int ty_ = ev.type[ie];
    int q = ie - 1;
    while (q >= 0 && ev.t[q] > tt) { ev.t[q + 1] = ev.t[q]; ev.icl[q + 1] = ev.icl[q]; ev.type[q + 1] = ev.type[q]; --q; }
    ev.t[q + 1] = tt; ev.icl[q + 1] = ii; ev.type[q + 1] = ty_;
  }
}
inline void apply_event(int64_t icl, int type, int64_t *aset, int &na, int cap) {
  if (type == +1) { if (na < cap) aset[na++] = icl; }
  else for (int m = 0; m < na; ++m) if (aset[m] == icl) { aset[m] = aset[na - 1]; --na; return; }
}
// sum_kap_active / sample_owner_clump — raytrace_clump.f90:621-666
inline double sum_kap_active(const World &w, const int64_t *active, int na, double xfreq_g, double kx, double ky, double kz) {
  double s = 0.0;
  for (int m = 0; m < na; ++m) s = s + kappa_clump(w, xfreq_g - ulos_clump(*w.cl, active[m], kx, ky, kz), active[m]);
  return s;
}
inline int64_t sample_owner_clump(const World &w, Rng &r, const int64_t *active, int na, double xfreq_g, double kx, double ky, double kz) {
  double cumul = 0.0;
  const double rnd = r.uniform();
  int64_t owner = 0;
  for (int m = 0; m < na; ++m) {
    cumul = cumul + kappa_clump(w, xfreq_g - ulos_clump(*w.cl, active[m], kx, ky, kz), active[m]);
    owner = active[m];
    if (rnd * sum_kap_active(w, active, na, xfreq_g, kx, ky, kz) <= cumul) return owner;
  }
  return owner;
}
// raytrace_to_edge_clump_overlap (:792-855) and _overlap_capped (:858-920; tau_max > 0)
double raytrace_to_edge_clump_overlap(const World &w, const Photon &p0, double tau_max, Counters *cnt) {
  const lart_clumps &c = *w.cl;
  double tau = 0.0;
  const double kx = p0.kx, ky = p0.ky, kz = p0.kz, xg = p0.xfreq;
  const double t_sp = sphere_exit_dist(c, p0.x, p0.y, p0.z, kx, ky, kz);
  if (t_sp <= 0.0) return tau;
  static thread_local ClumpEvents ev;
  static thread_local int64_t active[kMaxEvt];
  collect_ray_events_overlap(c, p0.x, p0.y, p0.z, kx, ky, kz, t_sp, ev);
  int na = active_set_at_point(c, p0.x, p0.y, p0.z, active, kMaxEvt);
  if (cnt) cnt->n_cellsteps += ev.n;
  double t_cur = 0.0;
  for (int ie = 0; ie <= ev.n; ++ie) {
    double t_next = std::min(ie < ev.n ? ev.t[ie] : t_sp, t_sp);
    double dt = t_next - t_cur;
    if (dt > 0.0) {
      tau = tau + sum_kap_active(w, active, na, xg, kx, ky, kz) * dt;
      if (tau_max > 0.0 && tau >= tau_max) return tau;
    }
    t_cur = t_next;
    if (ie < ev.n) apply_event(ev.icl[ie], ev.type[ie], active, na, kMaxEvt);
  }
  return tau;
}
// raytrace_to_tau_clump_overlap — raytrace_clump.f90:668-790 (the owner clump is drawn with one uniform)
void raytrace_to_tau_clump_overlap(const World &w, Photon &ph, double tau_in, Rng &r, Tally *tl, Counters *cnt) {
  const lart_clumps &c = *w.cl;
  const lart_grid &g = *w.g;
  const double kx = ph.kx, ky = ph.ky, kz = ph.kz, xg = ph.xfreq;
  double tau_rem = tau_in;
  auto escape = [&]() {
    ph.inside = false;
    ph.xfreq_ref = xg;  // see raytrace_to_tau_clump
    update_cell_idx(w, ph);
    if (tl) {
      int ix = static_cast<int>(std::floor((xg - g.xfreq_min) / g.dxfreq)) + 1;
      if (ix >= 1 && ix <= g.nxfreq) {
        tl->Jout[ix - 1] += ph.wgt;
        if (w.par->save_Jmu) tl->Jmu[(ix - 1) + static_cast<size_t>(g.nxfreq) * (jmu_bin(*w.par, ph.kz) - 1)] += ph.wgt;
      }
    }
  };
  const double t_sp = sphere_exit_dist(c, ph.x, ph.y, ph.z, kx, ky, kz);
  if (t_sp <= 0.0) { escape(); return; }
  static thread_local ClumpEvents ev;
  static thread_local int64_t active[kMaxEvt];
  collect_ray_events_overlap(c, ph.x, ph.y, ph.z, kx, ky, kz, t_sp, ev);
  int na = active_set_at_point(c, ph.x, ph.y, ph.z, active, kMaxEvt);
  if (cnt) cnt->n_cellsteps += ev.n;
  double t_cur = 0.0;
  for (int ie = 0; ie <= ev.n; ++ie) {
    double t_next = std::min(ie < ev.n ? ev.t[ie] : t_sp, t_sp);
    double dt = t_next - t_cur;
    if (dt <= 0.0) {
      if (ie < ev.n) apply_event(ev.icl[ie], ev.type[ie], active, na, kMaxEvt);
      continue;
    }
    double kap_tot = sum_kap_active(w, active, na, xg, kx, ky, kz);
    if (kap_tot > 0.0 && tau_rem <= kap_tot * dt) {  // scatters inside this segment
      double ds = tau_rem / kap_tot;
      ph.x = ph.x + (t_cur + ds) * kx; ph.y = ph.y + (t_cur + ds) * ky; ph.z = ph.z + (t_cur + ds) * kz;
      ph.xfreq = xg;
      ph.icl = static_cast<int>(sample_owner_clump(w, r, active, na, xg, kx, ky, kz));
      update_cell_idx(w, ph);
      return;
    }
    tau_rem = tau_rem - kap_tot * dt;
    t_cur = t_next;
    if (ie < ev.n) apply_event(ev.icl[ie], ev.type[ie], active, na, kMaxEvt);
  }
  ph.x = ph.x + t_sp * kx; ph.y = ph.y + t_sp * ky; ph.z = ph.z + t_sp * kz;
  ph.xfreq = xg;
  ph.icl = 0;
  escape();
}

// the ray tracers the procedure pointers select (setup.f90:806-815; peel_raytrace_to_edge, peelingoff_rect.f90:894-906)
// ---------------------------------------------------------------------------
// Octree AMR — octree_mod.f90 / raytrace_amr.f90.  Cells and leaves are 1-based as upstream;
// face index: 1=+x 2=-x 3=+y 4=-y 5=+z 6=-z.
// ---------------------------------------------------------------------------
// amr_find_leaf — octree_mod.f90:149-171
int amr_find_leaf(const World &w, double x, double y, double z) {
  const lart_amr &a = *w.amr;
  const lart_grid &g = *w.g;
  if (x < g.xmin || x > g.xmax || y < g.ymin || y > g.ymax || z < g.zmin || z > g.zmax) return 0;
  int icell = 1;
  for (;;) {
    if (a.ileaf[icell - 1] > 0) return a.ileaf[icell - 1];
    int ioct = 1;
    if (x >= a.cx[icell - 1]) ioct += 1;
    if (y >= a.cy[icell - 1]) ioct += 2;
    if (z >= a.cz[icell - 1]) ioct += 4;
    icell = a.children[8 * static_cast<size_t>(icell - 1) + ioct - 1];
    if (icell == 0) return 0;
  }
}
// amr_cell_exit — octree_mod.f90:412-458 (minloc: the first minimum wins)
inline void amr_cell_exit(const lart_amr &a, int icell, double x, double y, double z, double kx, double ky, double kz,
                          double &t_exit, int &iface) {
  const double cx = a.cx[icell - 1], cy = a.cy[icell - 1], cz = a.cz[icell - 1], h = a.ch[icell - 1];
  double t[6] = {kHugest, kHugest, kHugest, kHugest, kHugest, kHugest};
  if (kx > 0.0) t[0] = (cx + h - x) / kx; else if (kx < 0.0) t[1] = (cx - h - x) / kx;
  if (ky > 0.0) t[2] = (cy + h - y) / ky; else if (ky < 0.0) t[3] = (cy - h - y) / ky;
  if (kz > 0.0) t[4] = (cz + h - z) / kz; else if (kz < 0.0) t[5] = (cz - h - z) / kz;
  iface = 1;
  for (int q = 1; q < 6; ++q) if (t[q] < t[iface - 1]) iface = q + 1;
  t_exit = t[iface - 1];
}
// amr_next_leaf — octree_mod.f90:717-757: neighbour table, then descent with the face-normal octant bit set topologically
inline int amr_next_leaf(const lart_amr &a, int icell, int iface, double x, double y, double z) {
  int ineigh = a.neighbor[6 * static_cast<size_t>(icell - 1) + iface - 1];
  if (ineigh == 0) return 0;
  while (a.ileaf[ineigh - 1] == 0) {
    int ioct = 1;
    const bool bx = x >= a.cx[ineigh - 1], by = y >= a.cy[ineigh - 1], bz = z >= a.cz[ineigh - 1];
    switch (iface) {
      case 1: if (by) ioct += 2; if (bz) ioct += 4; break;
      case 2: ioct += 1; if (by) ioct += 2; if (bz) ioct += 4; break;
      case 3: if (bx) ioct += 1; if (bz) ioct += 4; break;
      case 4: if (bx) ioct += 1; ioct += 2; if (bz) ioct += 4; break;
      case 5: if (bx) ioct += 1; if (by) ioct += 2; break;
      default: if (bx) ioct += 1; if (by) ioct += 2; ioct += 4; break;
    }
    const int child = a.children[8 * static_cast<size_t>(ineigh - 1) + ioct - 1];
    if (child == 0) break;
    ineigh = child;
  }
  return a.ileaf[ineigh - 1];
}
// opacity of leaf il at frequency x (raytrace_amr.f90:114-121, band 1, no H2)
inline double amr_opacity(const World &w, int il, double xfreq) {
  double k = w.amr->rhokap[il - 1] * voigt_seon2(xfreq, w.amr->voigt_a[il - 1]);
  if (w.dust() && w.amr->rhokapD) k = k + w.amr->rhokapD[il - 1];
  return k;
}
// raytrace_to_edge_amr — raytrace_amr.f90:265-351 (open boundaries)
double raytrace_to_edge_amr(const World &w, const Photon &p0, Counters *cnt, int *nsteps_out = nullptr) {
  const lart_amr &a = *w.amr;
  double x = p0.x, y = p0.y, z = p0.z;
  const double kx = p0.kx, ky = p0.ky, kz = p0.kz;
  int il = p0.icell, ns = 0;
  double tau = 0.0;
  if (nsteps_out) *nsteps_out = 0;
  if (il <= 0) { il = amr_find_leaf(w, x, y, z); if (il <= 0) return tau; }
  double u1 = a.vfx[il - 1] * kx + a.vfy[il - 1] * ky + a.vfz[il - 1] * kz;
  double xfreq_loc = p0.xfreq;
  for (;;) {
    const int icell = a.icell_of_leaf[il - 1];
    double t_exit; int iface;
    amr_cell_exit(a, icell, x, y, z, kx, ky, kz, t_exit, iface);
    const double rhokap = amr_opacity(w, il, xfreq_loc);
    tau = tau + t_exit * rhokap;
    ++ns;
    if (tau >= kTauHuge) break;
    x = x + t_exit * kx; y = y + t_exit * ky; z = z + t_exit * kz;
    const int il_new = amr_next_leaf(a, icell, iface, x, y, z);
    if (il_new <= 0) break;
    const double Df_old = a.Dfreq[il - 1], Df_new = a.Dfreq[il_new - 1];
    const double u2 = a.vfx[il_new - 1] * kx + a.vfy[il_new - 1] * ky + a.vfz[il_new - 1] * kz;
    xfreq_loc = (xfreq_loc + u1) * Df_old / Df_new - u2;
    u1 = u2;
    il = il_new;
  }
  if (cnt) cnt->n_cellsteps += ns;
  if (nsteps_out) *nsteps_out = ns;
  return tau;
}
// raytrace_to_tau_amr — raytrace_amr.f90:77-259 (band 1, open boundaries)
void raytrace_to_tau_amr(const World &w, Photon &ph, double tau_in, Tally *tl, Counters *cnt, int *nsteps_out = nullptr) {
  const lart_amr &a = *w.amr;
  const lart_grid &g = *w.g;
  double x = ph.x, y = ph.y, z = ph.z;
  const double kx = ph.kx, ky = ph.ky, kz = ph.kz;
  int il = ph.icell, ns = 0;
  if (nsteps_out) *nsteps_out = 0;
  if (il <= 0) {
    il = amr_find_leaf(w, x, y, z);
    if (il <= 0) { ph.inside = false; return; }
  }
  double tau = 0.0;
  double u1 = a.vfx[il - 1] * kx + a.vfy[il - 1] * ky + a.vfz[il - 1] * kz;
  while (ph.inside) {
    const int icell = a.icell_of_leaf[il - 1];
    double t_exit; int iface;
    amr_cell_exit(a, icell, x, y, z, kx, ky, kz, t_exit, iface);
    const double rhokap = amr_opacity(w, il, ph.xfreq);
    ++ns;
    if (tau + t_exit * rhokap >= tau_in) {
      const double d_step = (rhokap > 0.0) ? (tau_in - tau) / rhokap : t_exit;
      x = x + d_step * kx; y = y + d_step * ky; z = z + d_step * kz;
      tau = tau_in;
      break;
    }
    tau = tau + t_exit * rhokap;
    x = x + t_exit * kx; y = y + t_exit * ky; z = z + t_exit * kz;
    const int il_new = amr_next_leaf(a, icell, iface, x, y, z);
    if (il_new <= 0) { ph.inside = false; break; }
    const double Df_old = a.Dfreq[il - 1], Df_new = a.Dfreq[il_new - 1];
    const double u2 = a.vfx[il_new - 1] * kx + a.vfy[il_new - 1] * ky + a.vfz[il_new - 1] * kz;
    ph.xfreq = (ph.xfreq + u1) * Df_old / Df_new - u2;
    u1 = u2;
    il = il_new;
  }
  ph.x = x; ph.y = y; ph.z = z;
  ph.icell = il; ph.jcell = 1; ph.kcell = 1;
  if (cnt) cnt->n_cellsteps += ns;
  if (nsteps_out) *nsteps_out = ns;
  if (!ph.inside) {  // :226-237 — lab frame, reference Doppler units, Jout (+ Jmu)
    u1 = a.vfx[il - 1] * kx + a.vfy[il - 1] * ky + a.vfz[il - 1] * kz;
    ph.xfreq = ph.xfreq + u1;
    const double xfreq_ref = ph.xfreq * (a.Dfreq[il - 1] / g.Dfreq_ref);
    ph.xfreq_ref = xfreq_ref;
    if (tl) {
      const int ix = static_cast<int>(std::floor((xfreq_ref - g.xfreq_min) / g.dxfreq)) + 1;
      if (ix >= 1 && ix <= g.nxfreq) {
        tl->Jout[ix - 1] += ph.wgt;
        if (w.par->save_Jmu) tl->Jmu[(ix - 1) + static_cast<size_t>(g.nxfreq) * (jmu_bin(*w.par, ph.kz) - 1)] += ph.wgt;
      }
    }
  }
}
// amr_xcrit_local — octree_mod.f90:248-284
void amr_xcrit_local(const World &w, int il, double x, double y, double z, double &xc, double &xc2) {
  if (w.par->core_skip_global) { xc = w.g->xcrit; xc2 = w.g->xcrit2; return; }
  xc = 0.0; xc2 = 0.0;
  if (il <= 0) return;
  const lart_amr &a = *w.amr;
  const int icell = a.icell_of_leaf[il - 1];
  const double h = a.ch[icell - 1];
  const double dl = std::min(h - std::fabs(x - a.cx[icell - 1]), std::min(h - std::fabs(y - a.cy[icell - 1]), h - std::fabs(z - a.cz[icell - 1])));
  if (dl <= 0.0) return;
  const double atau = a.voigt_a[il - 1] * a.rhokap[il - 1] * dl;
  if (atau > 1.0) { xc = std::pow(atau, 1.0 / 3.0) / 5.0; xc2 = xc * xc; }
}

inline double edge_tau(const World &w, const Photon &p, Counters *cnt) {
  if (w.amr) return raytrace_to_edge_amr(w, p, cnt);
  if (w.cl) return w.cl->has_overlap ? raytrace_to_edge_clump_overlap(w, p, -1.0, cnt) : raytrace_to_edge_clump(w, p, -1.0, cnt);
  return raytrace_to_edge(w, p, cnt);
}
inline double peel_tau(const World &w, const Photon &p, Counters *cnt) {
  if (w.amr) return raytrace_to_edge_amr(w, p, cnt);
  if (w.cl) return w.cl->has_overlap ? raytrace_to_edge_clump_overlap(w, p, kTauHugeClump, cnt)
                                     : raytrace_to_edge_clump(w, p, kTauHugeClump, cnt);
  return raytrace_to_edge(w, p, cnt);
}
// line-of-sight fluid velocity for the peel-off frequency (peelingoff_rect.f90:186-192, 403-409, 534-540, 658-664)
inline double peel_u1(const World &w, const Photon &ph, const Photon &po) {
  if (w.cl && ph.icl > 0) return ulos_clump(*w.cl, ph.icl, po.kx, po.ky, po.kz);
  return w.vdotk(po.icell, po.jcell, po.kcell, po.kx, po.ky, po.kz);
}

// ---------------------------------------------------------------------------
// rand_resonance_vz_seon — random_mt.f90:2562-2696
// The three wing variants of the Fortran (:2605-2635, :2638-2666, :2667-2690) differ
// only in the pieces of the piecewise-constant majorant in beta = exp(-p^2/2); they
// are written here as one loop over a region table (same draws, same order).
double rand_resonance_vz(Rng &r, double x0in, double a) {
  const double x0_crit = 1.0;
  const double xc = 1.0 + std::sqrt(2.0);
  const double two_over_PI = 2.0 / kPi;
  double x0 = std::fabs(x0in);
  double vz;
  Counters *cnt = r.cnt;
  if (x0 <= x0_crit) {  // :2579-2585
    for (;;) {
      double u1, u2;
      r.uniform2(u1, u2);
      vz = x0 + a * std::tan(kPi * (u1 - 0.5));
      if (cnt) cnt->n_reject_iter += 1;
      if (u2 <= std::exp(-vz * vz)) break;
    }
    if (x0in < 0.0) vz = -vz;
    return vz;
  }
  double x0sq = x0 * x0, api = a * kPi;
  double beta0 = std::exp(-x0sq / 2.0);
  double h0_two = beta0 / a, h0 = h0_two / 2.0;
  //   piece 0: beta = beta0*sqrt(xi),  C = beta/a     (selected when rs < p0)
  //   piece 1: beta = lo1 + w1*xi,     C = c1         (rs < p01, or the only piece)
  //   piece 2: beta = lo2 + w2*xi,     C = c2
  double p0 = 0.0, p01 = 0.0, lo1 = 0.0, w1 = 0.0, c1 = 0.0, lo2 = 0.0, w2 = 0.0, c2 = 0.0;
  int mode;  // 0: single piece; 1: pieces 0+1; 2: pieces 0+1+2
  double h2 = 0.3861 / (x0sq - 1.373);
  if (x0 < xc || !(h0 < h2)) {  // :2605-2635 (h1) and :2667-2690 (max(h1,h2))
    double dbeta = std::sqrt(two_over_PI * a * (1.0 - beta0) * beta0 * x0);
    double beta1 = beta0 + dbeta, one_b1 = 1.0 - beta1;
    double pb1 = std::sqrt(-2.0 * std::log(beta1));
    double h1 = two_over_PI * beta1 * pb1 / (x0sq - pb1 * pb1);
    double hm = (x0 < xc) ? h1 : ((h1 > h2) ? h1 : h2);
    double S0 = beta0 * h0, S1 = dbeta * h0, S2 = one_b1 * hm, Stot = S0 + S1 + S2;
    (void)S1;
    mode = 2; p0 = S0 / Stot; p01 = 1.0 - S2 / Stot;
    lo1 = beta0; w1 = dbeta; c1 = h0; lo2 = beta1; w2 = one_b1; c2 = hm;
  } else if (h0_two < h2) {  // :2638-2647
    mode = 0; lo1 = 0.0; w1 = 1.0; c1 = h2;
  } else {  // :2648-2666
    double S0 = beta0 * h0, one_b0 = 1.0 - beta0, S1 = one_b0 * h2, Stot = S0 + S1;
    mode = 1; p0 = S0 / Stot; lo1 = beta0; w1 = one_b0; c1 = h2;
  }
  double t1, delt;
  for (;;) {
    double beta, Cb, ua, ub, uacc;
    r.uniform2(ua, ub);
    if (mode == 0) {
      beta = lo1 + w1 * ua; Cb = c1; uacc = ub;
    } else {
      if (ua < p0) { beta = beta0 * std::sqrt(ub); Cb = beta / a; }
      else if (mode == 1 || ua < p01) { beta = lo1 + w1 * ub; Cb = c1; }
      else { beta = lo2 + w2 * ub; Cb = c2; }
      uacc = r.uniform();
    }
    double pb = std::sqrt(-2.0 * std::log(beta));
    double t2 = std::atan((pb - x0) / a);
    t1 = std::atan((-pb - x0) / a);
    delt = t2 - t1;
    if (cnt) cnt->n_reject_iter += 1;
    if (uacc * Cb < (beta / api) * delt) break;
  }
  vz = x0 + a * std::tan(delt * r.uniform() + t1);  // :2693
  if (x0in < 0.0) vz = -vz;
  return vz;
}

// rand_resonance — random_mt.f90:2974-2993 (x**(1/3) is a pow in the Fortran)
double rand_resonance(Rng &r, double E1) {
  const double one_over_three = 1.0 / 3.0;
  if (E1 > 0.0) {
    double p2 = std::sqrt((4.0 - E1) / (3.0 * E1));
    double Q = (4.0 * r.uniform() - 2.0) / (E1 * (p2 * p2 * p2));
    double W = std::pow(Q + std::sqrt(Q * Q + 1.0), one_over_three);
    return p2 * (W - 1.0 / W);
  } else if (E1 < 0.0) {
    double p2 = std::sqrt(std::fabs((4.0 - E1) / (3.0 * E1)));
    double Q = (4.0 * r.uniform() - 2.0) / (E1 * (p2 * p2 * p2));
    return 2.0 * p2 * std::cos((std::acos(Q) + kFourPi) / 3.0);
  }
  return 2.0 * r.uniform() - 1.0;
}

// rand_henyey_greenstein — random_mt.f90:3022-3042
double rand_hg(Rng &r, double g) {
  double x = r.uniform();
  if (g == 0.0) return 2.0 * x - 1.0;
  double g2 = g * g, twog = 2.0 * g;
  double q = (1.0 - g2) / (1.0 - g + twog * x);
  return ((1.0 + g2) - q * q) / twog;
}

// rand_voigt — random_mt.f90:3062-3083
double rand_voigt(Rng &r, double a) {
  double c = std::tan(kPi * r.uniform() - kHalfPi);
  return a * c + r.gauss() * (1.0 / std::sqrt(2.0));
}

// rand_alias_linear64 — random_mt.f90:2196-2216 (arrays 1-based in the Fortran)
double rand_alias_linear(Rng &r, const lart_scatt_mat &sm) {
  int n = sm.nPDF - 1;
  double uk, ua;
  r.uniform2(uk, ua);
  int k = static_cast<int>(std::floor(n * uk)) + 1;
  int idx = (ua < sm.phase_PDF[k - 1]) ? k : sm.alias[k - 1];
  double p0 = sm.S11[idx - 1], p1 = sm.S11[idx];
  double x0 = sm.coss[idx - 1], x1 = sm.coss[idx];
  return (std::sqrt(p0 * p0 + (p1 * p1 - p0 * p0) * r.uniform()) - p0) * (x1 - x0) / (p1 - p0) + x0;
}

// interp_eq — mathlib.f90:71-106 (integer conversion truncates toward zero)
double interp_eq(const double *x, const double *y, int n, double xnew) {
  double dx = x[1] - x[0];
  int i = static_cast<int>((xnew - x[0]) / dx + 1.0);
  if (i <= 0) return y[0];
  if (i >= n) return y[n - 1];
  return y[i - 1] + (y[i] - y[i - 1]) * (xnew - x[i - 1]) / dx;
}

// car_xcrit_local — grid_mod_car.f90:1598-1629
void car_xcrit_local(const World &w, int i, int j, int k, double x, double y, double z, double &xc, double &xc2) {
  if (w.amr) { amr_xcrit_local(w, i, x, y, z, xc, xc2); return; }
  const lart_grid &g = *w.g;
  if (w.par->core_skip_global) {
    xc = g.xcrit;
    xc2 = g.xcrit2;
    return;
  }
  xc = 0.0;
  xc2 = 0.0;
  if (i < 1 || j < 1 || k < 1) return;
  double dlx = std::fmin(x - w.xface(i), w.xface(i + 1) - x);
  double dly = std::fmin(y - w.yface(j), w.yface(j + 1) - y);
  double dlz = std::fmin(z - w.zface(k), w.zface(k + 1) - z);
  double dl = std::fmin(dlx, std::fmin(dly, dlz));
  if (dl <= 0.0) return;
  double atau = w.voigt_a(i, j, k) * w.rhokap(i, j, k) * dl;
  if (atau > 1.0) {
    xc = std::pow(atau, 1.0 / 3.0) / 5.0;
    xc2 = xc * xc;
  }
}

// ---------------------------------------------------------------------------
// Peel-off (outside observer, TAN image) — peelingoff_rect.f90
// ---------------------------------------------------------------------------
struct PeelGeom {
  Photon pobs;
  double r2;
  int ix, iy;
  bool in_image;
};

// common head of every peeling routine: :44-61 / :326-357 / :595-627
inline PeelGeom peel_geometry(const Photon &ph, const lart_observer &ob) {
  PeelGeom pg;
  pg.pobs = ph;
  Photon &po = pg.pobs;
  po.kx = ob.x - ph.x;
  po.ky = ob.y - ph.y;
  po.kz = ob.z - ph.z;
  pg.r2 = po.kx * po.kx + po.ky * po.ky + po.kz * po.kz;
  double r = std::sqrt(pg.r2);
  po.kx /= r; po.ky /= r; po.kz /= r;
  const double *R = ob.rmatrix;  // R[(row-1)+3*(col-1)]
  double kx = R[0] * po.kx + R[3] * po.ky + R[6] * po.kz;
  double ky = R[1] * po.kx + R[4] * po.ky + R[7] * po.kz;
  double kz = R[2] * po.kx + R[5] * po.ky + R[8] * po.kz;
  pg.ix = static_cast<int>(std::floor(std::atan2(-kx, kz) * kRad2Deg / ob.dxim + ob.nxim / 2.0)) + 1;
  pg.iy = static_cast<int>(std::floor(std::atan2(-ky, kz) * kRad2Deg / ob.dyim + ob.nyim / 2.0)) + 1;
  pg.in_image = pg.ix >= 1 && pg.ix <= ob.nxim && pg.iy >= 1 && pg.iy <= ob.nyim;
  return pg;
}

inline size_t pix2(const lart_observer &ob, int ix, int iy) { return static_cast<size_t>(ix - 1) + static_cast<size_t>(ob.nxim) * (iy - 1); }
inline size_t pix3(const World &w, const lart_observer &ob, int ixf, int ix, int iy) {
  return static_cast<size_t>(ixf - 1) + static_cast<size_t>(w.g->nxfreq) * pix2(ob, ix, iy);
}
inline void add_to(double *arr, size_t i, double v) { if (arr) atomic_add(arr + i, v); }

// peeling_direct_outside — peelingoff_rect.f90:24-129
void peeling_direct(const World &w, const Photon &ph, Tally &tl) {
  const lart_params &par = *w.par;
  const lart_grid &g = *w.g;
  for (int i = 0; i < par.nobs; ++i) {
    const lart_observer &ob = w.obs[i];
    lart_observer_out &oo = tl.shared->obs[i];
    PeelGeom pg = peel_geometry(ph, ob);
    Photon &po = pg.pobs;
    double xfreq_ref;
    if (w.cl && ph.icl > 0) {  // :65-69 — xfreq is in the owner clump's rest frame
      xfreq_ref = ph.xfreq + ulos_clump(*w.cl, ph.icl, po.kx, po.ky, po.kz);
    } else if (!par.comoving_source) {  // :70-80
      double u1 = w.vdotk(ph.icell, ph.jcell, ph.kcell, ph.kx, ph.ky, ph.kz);
      xfreq_ref = ph.xfreq + u1;
      double u2 = w.vdotk(po.icell, po.jcell, po.kcell, po.kx, po.ky, po.kz);
      po.xfreq = xfreq_ref - u2;
    } else {  // :81-87
      double u1 = w.vdotk(po.icell, po.jcell, po.kcell, po.kx, po.ky, po.kz);
      xfreq_ref = ph.xfreq + u1;
    }
    xfreq_ref = xfreq_ref * (w.Dfreq(ph.icell, ph.jcell, ph.kcell) / g.Dfreq_ref);
    int ixf = static_cast<int>(std::floor((xfreq_ref - g.xfreq_min) / g.dxfreq)) + 1;
    if (!pg.in_image) continue;
    double tau = peel_tau(w, po, &tl.cnt);
    tl.cnt.n_peel += 1;
    double wgt0 = 1.0 / (kFourPi * pg.r2) * ph.wgt;
    double wgt = std::exp(-tau) * wgt0;
    if (par.save_peeloff_2D) {
      size_t q = pix2(ob, pg.ix, pg.iy);
      add_to(oo.direc_2D, q, wgt);
      if (par.use_stokes) add_to(oo.I_2D, q, wgt);
      if (par.save_direc0) add_to(oo.direc0_2D, q, wgt0);
    }
    if (par.save_peeloff_3D && ixf >= 1 && ixf <= g.nxfreq) {
      size_t q = pix3(w, ob, ixf, pg.ix, pg.iy);
      add_to(oo.direc, q, wgt);
      if (par.use_stokes) add_to(oo.I, q, wgt);
      if (par.save_direc0) add_to(oo.direc0, q, wgt0);
    }
  }
}

// shared tail of the two Stokes peels: rotate to the detector and deposit
// (peelingoff_rect.f90:428-479 and :243-296)
inline void deposit_stokes(const World &w, const lart_observer &ob, lart_observer_out &oo, const Photon &po, const PeelGeom &pg,
                           int ixf, double wgt, double Iobs, double Qobs, double Uobs, double Vobs) {
  const double *R = ob.rmatrix;
  double cosg = -(R[0] * po.nx + R[3] * po.ny + R[6] * po.nz);
  double sing = R[1] * po.nx + R[4] * po.ny + R[7] * po.nz;
  double cos2g = 2.0 * cosg * cosg - 1.0, sin2g = 2.0 * cosg * sing;
  double Idet = Iobs, Qdet = cos2g * Qobs + sin2g * Uobs, Udet = -sin2g * Qobs + cos2g * Uobs, Vdet = Vobs;
  if (w.par->save_peeloff_2D) {
    size_t q = pix2(ob, pg.ix, pg.iy);
    add_to(oo.scatt_2D, q, wgt * Idet);
    add_to(oo.I_2D, q, wgt * Idet);
    add_to(oo.Q_2D, q, wgt * Qdet);
    add_to(oo.U_2D, q, wgt * Udet);
    add_to(oo.V_2D, q, wgt * Vdet);
  }
  if (w.par->save_peeloff_3D && ixf >= 1 && ixf <= w.g->nxfreq) {
    size_t q = pix3(w, ob, ixf, pg.ix, pg.iy);
    add_to(oo.scatt, q, wgt * Idet);
    add_to(oo.I, q, wgt * Idet);
    add_to(oo.Q, q, wgt * Qdet);
    add_to(oo.U, q, wgt * Udet);
    add_to(oo.V, q, wgt * Vdet);
  }
}

// azimuth of the observer direction in the photon's (m,n) frame and the new
// reference normal (:364-380 / :208-224)
inline void stokes_azimuth(const Photon &ph, Photon &po, double cost, double &sint, double &cosp, double &sinp, double &cos2p, double &sin2p) {
  sint = std::sqrt(1.0 - cost * cost);
  if (sint == 0.0) {
    cosp = 1.0; sinp = 0.0; cos2p = 1.0; sin2p = 0.0;
  } else {
    cosp = (po.kx * ph.mx + po.ky * ph.my + po.kz * ph.mz) / sint;
    sinp = (po.kx * ph.nx + po.ky * ph.ny + po.kz * ph.nz) / sint;
    cos2p = 2.0 * cosp * cosp - 1.0;
    sin2p = 2.0 * cosp * sinp;
  }
  po.nx = -sinp * ph.mx + cosp * ph.nx;
  po.ny = -sinp * ph.my + cosp * ph.ny;
  po.nz = -sinp * ph.mz + cosp * ph.nz;
}

// peeling_resonance_stokes_outside — peelingoff_rect.f90:303-482
void peeling_resonance_stokes(const World &w, const Photon &ph, Tally &tl, double xfreq_atom, const double vel_atom[3]) {
  const lart_params &par = *w.par;
  const lart_grid &g = *w.g;
  for (int i = 0; i < par.nobs; ++i) {
    const lart_observer &ob = w.obs[i];
    PeelGeom pg = peel_geometry(ph, ob);
    if (!pg.in_image) continue;
    Photon &po = pg.pobs;
    double cost = ph.kx * po.kx + ph.ky * po.ky + ph.kz * po.kz;
    double sint, cosp, sinp, cos2p, sin2p;
    stokes_azimuth(ph, po, cost, sint, cosp, sinp, cos2p, sin2p);
    double cost2 = cost * cost;
    double S22 = 0.75 * ph.E1 * (cost2 + 1.0);
    double S11 = S22 + ph.E2;
    double S12 = 0.75 * ph.E1 * (cost2 - 1.0);
    double S33 = 1.5 * ph.E1 * cost;
    double S44 = 1.5 * ph.E3 * cost;
    double xfreq = xfreq_atom + (vel_atom[0] * cosp + vel_atom[1] * sinp) * sint + vel_atom[2] * cost;  // :392
    double Dcell = w.Dfreq(ph.icell, ph.jcell, ph.kcell);
    if (par.recoil) xfreq -= (w.line->g_recoil0 / Dcell) * (1.0 - cost);
    double u1 = peel_u1(w, ph, po);
    double xfreq_ref = (xfreq + u1) * (Dcell / g.Dfreq_ref);
    int ixf = static_cast<int>(std::floor((xfreq_ref - g.xfreq_min) / g.dxfreq)) + 1;
    double Q0 = cos2p * ph.Q + sin2p * ph.U, U0 = -sin2p * ph.Q + cos2p * ph.U;
    double Iobs = (S11 + S12 * Q0) / kFourPi, Qobs = (S12 + S22 * Q0) / kFourPi;
    double Uobs = (S33 * U0) / kFourPi, Vobs = (S44 * ph.V) / kFourPi;
    po.xfreq = xfreq;
    double tau = peel_tau(w, po, &tl.cnt);
    tl.cnt.n_peel += 1;
    double wgt = 1.0 / pg.r2 * std::exp(-tau) * ph.wgt;
    deposit_stokes(w, ob, tl.shared->obs[i], po, pg, ixf, wgt, Iobs, Qobs, Uobs, Vobs);
  }
}

// peeling_resonance_nostokes_outside — peelingoff_rect.f90:576-690
void peeling_resonance_nostokes(const World &w, const Photon &ph, Tally &tl, double xfreq_atom, const double vel_atom[3]) {
  const lart_params &par = *w.par;
  const lart_grid &g = *w.g;
  for (int i = 0; i < par.nobs; ++i) {
    const lart_observer &ob = w.obs[i];
    lart_observer_out &oo = tl.shared->obs[i];
    PeelGeom pg = peel_geometry(ph, ob);
    if (!pg.in_image) continue;
    Photon &po = pg.pobs;
    double cost = ph.kx * po.kx + ph.ky * po.ky + ph.kz * po.kz;
    double cost2 = cost * cost;
    double sint = std::sqrt(1.0 - cost2);
    double rho1 = std::sqrt(1.0 - ph.kz * ph.kz) * sint;
    double cosp, sinp;
    if (rho1 == 0.0) { cosp = 1.0; sinp = 0.0; }
    else {
      double rho = 1.0 / rho1;
      cosp = rho * (cost * ph.kz - po.kz);
      sinp = rho * (ph.kx * po.ky - po.kx * ph.ky);
    }
    double xfreq = xfreq_atom + (vel_atom[0] * cosp + vel_atom[1] * sinp) * sint + vel_atom[2] * cost;
    double Dcell = w.Dfreq(ph.icell, ph.jcell, ph.kcell);
    if (par.recoil) xfreq -= (w.line->g_recoil0 / Dcell) * (1.0 - cost);
    double u1 = peel_u1(w, ph, po);
    double xfreq_ref = (xfreq + u1) * (Dcell / g.Dfreq_ref);
    int ixf = static_cast<int>(std::floor((xfreq_ref - g.xfreq_min) / g.dxfreq)) + 1;
    po.xfreq = xfreq;
    double tau = peel_tau(w, po, &tl.cnt);
    tl.cnt.n_peel += 1;
    double peel = 0.75 * ph.E1 * (cost2 + 1.0) + ph.E2;
    double wgt = peel / (kFourPi * pg.r2) * std::exp(-tau) * ph.wgt;
    if (par.save_peeloff_2D) add_to(oo.scatt_2D, pix2(ob, pg.ix, pg.iy), wgt);
    if (par.save_peeloff_3D && ixf >= 1 && ixf <= g.nxfreq) add_to(oo.scatt, pix3(w, ob, ixf, pg.ix, pg.iy), wgt);
  }
}

// peeling_dust_stokes_outside — peelingoff_rect.f90:131-299
void peeling_dust_stokes(const World &w, const Photon &ph, Tally &tl) {
  const lart_params &par = *w.par;
  const lart_grid &g = *w.g;
  const lart_scatt_mat &sm = *w.sm;
  for (int i = 0; i < par.nobs; ++i) {
    const lart_observer &ob = w.obs[i];
    PeelGeom pg = peel_geometry(ph, ob);
    Photon &po = pg.pobs;
    double u1 = peel_u1(w, ph, po);
    double xfreq_ref = (ph.xfreq + u1) * (w.Dfreq(ph.icell, ph.jcell, ph.kcell) / g.Dfreq_ref);
    int ixf = static_cast<int>(std::floor((xfreq_ref - g.xfreq_min) / g.dxfreq)) + 1;
    if (!pg.in_image) continue;
    double cost = ph.kx * po.kx + ph.ky * po.ky + ph.kz * po.kz;
    double sint, cosp, sinp, cos2p, sin2p;
    stokes_azimuth(ph, po, cost, sint, cosp, sinp, cos2p, sin2p);
    double S11 = interp_eq(sm.coss, sm.S11, sm.nPDF, cost);
    double S12 = interp_eq(sm.coss, sm.S12, sm.nPDF, cost);
    double S33 = interp_eq(sm.coss, sm.S33, sm.nPDF, cost);
    double S34 = interp_eq(sm.coss, sm.S34, sm.nPDF, cost);
    double Q0 = cos2p * ph.Q + sin2p * ph.U, U0 = -sin2p * ph.Q + cos2p * ph.U;
    double Iobs = (S11 * ph.I + S12 * Q0) / kTwoPi, Qobs = (S12 * ph.I + S11 * Q0) / kTwoPi;
    double Uobs = (S33 * U0 + S34 * ph.V) / kTwoPi, Vobs = (-S34 * U0 + S33 * ph.V) / kTwoPi;
    double tau = peel_tau(w, po, &tl.cnt);  // pobs%xfreq = photon%xfreq (:195-197)
    tl.cnt.n_peel += 1;
    double wgt = 1.0 / pg.r2 * std::exp(-tau) * ph.wgt;
    deposit_stokes(w, ob, tl.shared->obs[i], po, pg, ixf, wgt, Iobs, Qobs, Uobs, Vobs);
  }
}

// peeling_dust_nostokes_outside — peelingoff_rect.f90:484-574
void peeling_dust_nostokes(const World &w, const Photon &ph, Tally &tl) {
  const lart_params &par = *w.par;
  const lart_grid &g = *w.g;
  for (int i = 0; i < par.nobs; ++i) {
    const lart_observer &ob = w.obs[i];
    lart_observer_out &oo = tl.shared->obs[i];
    PeelGeom pg = peel_geometry(ph, ob);
    if (!pg.in_image) continue;
    Photon &po = pg.pobs;
    double u1 = peel_u1(w, ph, po);
    double xfreq_ref = (ph.xfreq + u1) * (w.Dfreq(ph.icell, ph.jcell, ph.kcell) / g.Dfreq_ref);
    int ixf = static_cast<int>(std::floor((xfreq_ref - g.xfreq_min) / g.dxfreq)) + 1;
    double tau = peel_tau(w, po, &tl.cnt);
    tl.cnt.n_peel += 1;
    double cosa = ph.kx * po.kx + ph.ky * po.ky + ph.kz * po.kz;
    double hg = par.hgg;
    double peel = (1.0 - hg * hg) / std::pow((1.0 + hg * hg) - 2.0 * hg * cosa, 1.5) / kFourPi;
    double wgt = peel / pg.r2 * std::exp(-tau) * ph.wgt;
    if (par.save_peeloff_2D) add_to(oo.scatt_2D, pix2(ob, pg.ix, pg.iy), wgt);
    if (par.save_peeloff_3D && ixf >= 1 && ixf <= g.nxfreq) add_to(oo.scatt, pix3(w, ob, ixf, pg.ix, pg.iy), wgt);
  }
}

// ---------------------------------------------------------------------------
// Scattering — scattering_car.f90
// ---------------------------------------------------------------------------
// do_resonance1 — line_mod.f90:108-139
inline void do_resonance1(const World &w, Photon &ph, Rng &r, double &uz, double &xfreq_atom, double &cost, double &sint) {
  if (w.cl) {  // do_resonance1_clump — line_clump_mod.f90:29-58: sample in the clump's own Doppler units
    const lart_clumps &c = *w.cl;
    const double scale = c.Dfreq_ref / c.Dfreq[ph.icl - 1], scale_inv = 1.0 / scale;
    const double xloc = ph.xfreq * scale;
    const double uz_loc = rand_resonance_vz(r, xloc, c.voigt_a[ph.icl - 1]);
    xfreq_atom = (xloc - uz_loc) * scale_inv;
    uz = uz_loc * scale_inv;
  } else {
    uz = rand_resonance_vz(r, ph.xfreq, w.voigt_a(ph.icell, ph.jcell, ph.kcell));
    xfreq_atom = ph.xfreq - uz;
  }
  ph.E1 = w.line->E1; ph.E2 = w.line->E2; ph.E3 = w.line->E3;
  cost = rand_resonance(r, ph.E1);
  sint = std::sqrt(1.0 - cost * cost);
}

// rotate the (m,n,k) triad — scattering_car.f90:470-484 / :314-328
inline void rotate_triad(Photon &ph, double cost, double sint, double cosp, double sinp) {
  double px = cosp * ph.mx + sinp * ph.nx, py = cosp * ph.my + sinp * ph.ny, pz = cosp * ph.mz + sinp * ph.nz;
  ph.nx = cosp * ph.nx - sinp * ph.mx;
  ph.ny = cosp * ph.ny - sinp * ph.my;
  ph.nz = cosp * ph.nz - sinp * ph.mz;
  ph.mx = cost * px - sint * ph.kx;
  ph.my = cost * py - sint * ph.ky;
  ph.mz = cost * pz - sint * ph.kz;
  ph.kx = sint * px + cost * ph.kx;
  ph.ky = sint * py + cost * ph.ky;
  ph.kz = sint * pz + cost * ph.kz;
}

// new direction without a triad — scattering_car.f90:795-809 / :568-582
inline void rotate_k(Photon &ph, double cost, double sint, double cosp, double sinp) {
  if (std::fabs(ph.kz) >= 0.99999999999) {
    ph.kx = sint * cosp; ph.ky = sint * sinp; ph.kz = cost;
  } else {
    double kx1 = ph.kx, ky1 = ph.ky, kz1 = ph.kz;
    double kr = std::sqrt(kx1 * kx1 + ky1 * ky1);
    ph.kx = cost * kx1 + sint * (kz1 * kx1 * cosp - ky1 * sinp) / kr;
    ph.ky = cost * ky1 + sint * (kz1 * ky1 * cosp + kx1 * sinp) / kr;
    ph.kz = cost * kz1 - sint * cosp * kr;
  }
}

// azimuth by rejection — scattering_car.f90:364-371 / :280-287
inline double sample_phi_stokes(Rng &r, const Photon &ph, double S12overS11) {
  double phi;
  for (;;) {
    double u1, u2;
    r.uniform2(u1, u2);
    phi = kTwoPi * u1;
    double phi1 = 2.0 * phi;
    double Prand = (1.0 + std::fabs(S12overS11) * std::sqrt(ph.Q * ph.Q + ph.U * ph.U)) * u2;
    double Pcomp = 1.0 + S12overS11 * (ph.Q * std::cos(phi1) + ph.U * std::sin(phi1));
    if (r.cnt) r.cnt->n_reject_iter += 1;
    if (Prand <= Pcomp) break;
  }
  return phi;
}

// scatter_resonance_stokes — scattering_car.f90:331-486
void scatter_resonance_stokes(const World &w, Photon &ph, Rng &r, Tally &tl) {
  const lart_params &par = *w.par;
  ph.nscatt_gas += ph.wgt;
  add_to_Pa(w, ph, tl);  // :356-358
  double uz, xfreq_atom, cost, sint;
  do_resonance1(w, ph, r, uz, xfreq_atom, cost, sint);
  double cost2 = cost * cost;
  double S22 = 0.75 * ph.E1 * (cost2 + 1.0), S11 = S22 + ph.E2, S12 = 0.75 * ph.E1 * (cost2 - 1.0);
  double S33 = 1.5 * ph.E1 * cost, S44 = 1.5 * ph.E3 * cost;
  double phi = sample_phi_stokes(r, ph, S12 / S11);
  double cosp = std::cos(phi), sinp = std::sin(phi);
  double xc = 0.0, xc2 = 0.0;
  if (par.core_skip) car_xcrit_local(w, ph.icell, ph.jcell, ph.kcell, ph.x, ph.y, ph.z, xc, xc2);
  double ux, uy;
  if (par.core_skip && std::fabs(ph.xfreq) < xc) {  // :397-401
    double u1, u2;
    r.uniform2(u1, u2);
    double phi2 = kTwoPi * u1;
    double uxy = std::sqrt(xc2 - std::log(u2));
    ux = uxy * std::cos(phi2);
    uy = uxy * std::sin(phi2);
  } else {  // :413-414
    const double one_over_sqrt2 = 1.0 / std::sqrt(2.0);
    ux = r.gauss() * one_over_sqrt2;
    uy = r.gauss() * one_over_sqrt2;
  }
  if (w.cl) { const double vth_ratio = w.cl->Dfreq[ph.icl - 1] / w.cl->Dfreq_ref; ux = ux * vth_ratio; uy = uy * vth_ratio; }  // :384-388
  ph.xfreq = xfreq_atom + uz * cost + (ux * cosp + uy * sinp) * sint;
  if (par.recoil) ph.xfreq -= (w.line->g_recoil0 / (w.cl ? w.cl->Dfreq[ph.icl - 1] : w.Dfreq(ph.icell, ph.jcell, ph.kcell))) * (1.0 - cost);
  if (par.save_peeloff) {
    double va[3] = {ux, uy, uz};
    peeling_resonance_stokes(w, ph, tl, xfreq_atom, va);
  }
  double cos2p = 2.0 * cosp * cosp - 1.0, sin2p = 2.0 * sinp * cosp;
  double Q0 = cos2p * ph.Q + sin2p * ph.U, U0 = -sin2p * ph.Q + cos2p * ph.U;
  double I1 = S11 + S12 * Q0, Q1 = S12 + S22 * Q0, U1 = S33 * U0, V1 = S44 * ph.V;
  ph.I = 1.0; ph.Q = Q1 / I1; ph.U = U1 / I1; ph.V = V1 / I1;
  rotate_triad(ph, cost, sint, cosp, sinp);
}

// scatter_resonance_nostokes — scattering_car.f90:660-827
void scatter_resonance_nostokes(const World &w, Photon &ph, Rng &r, Tally &tl) {
  const lart_params &par = *w.par;
  ph.nscatt_gas += ph.wgt;
  add_to_Pa(w, ph, tl);  // :691-693
  double uz, xfreq_atom, cost, sint;
  do_resonance1(w, ph, r, uz, xfreq_atom, cost, sint);
  double phi = kTwoPi * r.uniform();
  double cosp = std::cos(phi), sinp = std::sin(phi);
  double xc = 0.0, xc2 = 0.0;
  if (par.core_skip) car_xcrit_local(w, ph.icell, ph.jcell, ph.kcell, ph.x, ph.y, ph.z, xc, xc2);
  double u1, u2;
  r.uniform2(u1, u2);
  double phi2 = kTwoPi * u1;
  double uxy = (par.core_skip && std::fabs(ph.xfreq) < xc) ? std::sqrt(xc2 - std::log(u2)) : std::sqrt(-std::log(u2));
  double ux = uxy * std::cos(phi2), uy = uxy * std::sin(phi2);
  if (w.cl) { const double vth_ratio = w.cl->Dfreq[ph.icl - 1] / w.cl->Dfreq_ref; ux = ux * vth_ratio; uy = uy * vth_ratio; }  // :715-719
  ph.xfreq = xfreq_atom + uz * cost + (ux * cosp + uy * sinp) * sint;
  if (par.recoil) ph.xfreq -= (w.line->g_recoil0 / (w.cl ? w.cl->Dfreq[ph.icl - 1] : w.Dfreq(ph.icell, ph.jcell, ph.kcell))) * (1.0 - cost);
  if (par.save_peeloff) {
    double va[3] = {ux, uy, uz};
    peeling_resonance_nostokes(w, ph, tl, xfreq_atom, va);
  }
  rotate_k(ph, cost, sint, cosp, sinp);
}

// dust absorption / weight reduction shared by both dust routines —
// scattering_car.f90:219-264 / :506-555.  Returns false when the photon died.
inline bool dust_absorb(const World &w, Photon &ph, Rng &r, Tally &tl) {
  const lart_params &par = *w.par;
  const lart_grid &g = *w.g;
  auto lab_bin = [&](double &xref) {
    double uu1 = w.vdotk(ph.icell, ph.jcell, ph.kcell, ph.kx, ph.ky, ph.kz);
    xref = (ph.xfreq + uu1) * (w.Dfreq(ph.icell, ph.jcell, ph.kcell) / g.Dfreq_ref);
    return static_cast<int>(std::floor((xref - g.xfreq_min) / g.dxfreq)) + 1;
  };
  double xref = 0.0;
  if (!par.use_reduced_wgt) {
    if (r.uniform() > par.albedo) {
      if (par.save_Jabs) {
        int ix = lab_bin(xref);
        if (ix >= 1 && ix <= g.nxfreq) tl.Jabs[ix - 1] += ph.wgt;
      }
      ph.inside = false;
      if (par.save_all_photons) {
        if (!par.save_Jabs) lab_bin(xref);
        ph.xfreq_ref = xref;
        ph.wgt = 0.0;
      }
      return false;
    }
  } else {
    if (par.save_Jabs) {
      int ix = lab_bin(xref);
      if (ix >= 1 && ix <= g.nxfreq) tl.Jabs[ix - 1] += ph.wgt * (1.0 - par.albedo);
    }
    ph.wgt = ph.wgt * par.albedo;
  }
  return true;
}

// scatter_dust_stokes — scattering_car.f90:201-329
void scatter_dust_stokes(const World &w, Photon &ph, Rng &r, Tally &tl) {
  const lart_scatt_mat &sm = *w.sm;
  ph.nscatt_dust += ph.wgt;
  if (!dust_absorb(w, ph, r, tl)) return;
  if (w.par->save_peeloff) peeling_dust_stokes(w, ph, tl);
  double cost = rand_alias_linear(r, sm);
  double sint = std::sqrt(1.0 - cost * cost);
  double S11 = interp_eq(sm.coss, sm.S11, sm.nPDF, cost), S12 = interp_eq(sm.coss, sm.S12, sm.nPDF, cost);
  double S33 = interp_eq(sm.coss, sm.S33, sm.nPDF, cost), S34 = interp_eq(sm.coss, sm.S34, sm.nPDF, cost);
  double phi = sample_phi_stokes(r, ph, S12 / S11);
  double cosp = std::cos(phi), sinp = std::sin(phi);
  double cos2p = 2.0 * cosp * cosp - 1.0, sin2p = 2.0 * sinp * cosp;
  double Q0 = cos2p * ph.Q + sin2p * ph.U, U0 = -sin2p * ph.Q + cos2p * ph.U;
  double I1 = S11 + S12 * Q0, Q1 = S12 + S11 * Q0, U1 = S33 * U0 + S34 * ph.V, V1 = -S34 * U0 + S33 * ph.V;
  ph.I = 1.0; ph.Q = Q1 / I1; ph.U = U1 / I1; ph.V = V1 / I1;
  rotate_triad(ph, cost, sint, cosp, sinp);
}

// scatter_dust_nostokes — scattering_car.f90:488-584
void scatter_dust_nostokes(const World &w, Photon &ph, Rng &r, Tally &tl) {
  ph.nscatt_dust += ph.wgt;
  if (!dust_absorb(w, ph, r, tl)) return;
  if (w.par->save_peeloff) peeling_dust_nostokes(w, ph, tl);
  double cost = rand_hg(r, w.par->hgg);
  double sint = std::sqrt(1.0 - cost * cost);
  double phi = kTwoPi * r.uniform();
  rotate_k(ph, cost, sint, std::cos(phi), std::sin(phi));
}

// scattering — scattering_car.f90:14-120 (Cartesian, no H2/AMR/clump branches)
void scattering(const World &w, Photon &ph, Rng &r, Tally &tl) {
  tl.cnt.n_scatter += 1;
  bool to_dust = false;
  if (w.cl) {  // scattering_car.f90:72-87
    if (w.dust() && w.cl->rhokapD) {
      const lart_clumps &c = *w.cl;
      double p_dust = c.rhokapD[ph.icl - 1] / (c.rhokap[ph.icl - 1] * voigt_clump(c, ph.xfreq, ph.icl) + c.rhokapD[ph.icl - 1]);
      to_dust = r.uniform() <= p_dust;
    }
  } else if (w.dust()) {
    int i = ph.icell, j = ph.jcell, k = ph.kcell;
    double p_dust = w.rhokapD(i, j, k) / (w.rhokap(i, j, k) * w.calc_voigt(ph.xfreq, i, j, k) + w.rhokapD(i, j, k));
    to_dust = r.uniform() <= p_dust;
  }
  if (to_dust) {
    if (w.par->use_stokes) scatter_dust_stokes(w, ph, r, tl);
    else scatter_dust_nostokes(w, ph, r, tl);
  } else {
    const bool ov = w.cl && w.cl->has_overlap;  // scatter_resonance_clump_* wrappers (scattering_car.f90:897-945)
    if (ov) ph.xfreq = ph.xfreq - ulos_clump(*w.cl, ph.icl, ph.kx, ph.ky, ph.kz);
    if (w.par->use_stokes) scatter_resonance_stokes(w, ph, r, tl);
    else scatter_resonance_nostokes(w, ph, r, tl);
    if (ov) ph.xfreq = ph.xfreq + ulos_clump(*w.cl, ph.icl, ph.kx, ph.ky, ph.kz);
  }
}

// ---------------------------------------------------------------------------
// generate_photon + setup_isotropic_injection — generate_photon.f90:3-339, 342-408
// ---------------------------------------------------------------------------
void generate_photon(const World &w, Photon &ph, Rng &r, Tally &tl) {
  const lart_params &par = *w.par;
  const lart_grid &g = *w.g;
  switch (par.source_geometry) {
    case LART_SRC_UNIFORM_SPHERE: {  // :34-42
      double u1, u2;
      r.uniform2(u1, u2);
      double rp = std::pow(u1, 1.0 / 3.0) * par.source_rmax;
      double cost = 2.0 * u2 - 1.0, sint = std::sqrt(1.0 - cost * cost), phi = kTwoPi * r.uniform();
      ph.x = rp * sint * std::cos(phi); ph.y = rp * sint * std::sin(phi); ph.z = rp * cost;
      break;
    }
    case LART_SRC_UNIFORM:  // :50-54
    {
      double u1, u2;
      r.uniform2(u1, u2);
      ph.x = (g.xmax - g.xmin) * u1 + g.xmin;
      ph.y = (g.ymax - g.ymin) * u2 + g.ymin;
      ph.z = (g.zmax - g.zmin) * r.uniform() + g.zmin;
    }
      break;
    case LART_SRC_PLANE_ILLUMINATION:  // random_plane_illumination :729-760
      if (w.atm == LART_ATM_PLANE) {
        ph.x = 0.0; ph.y = 0.0; ph.z = g.zmax;  // par%zmax
      } else {
        double rp = g.rmax * std::sqrt(r.uniform());
        double phi = (par.xy_symmetry ? kHalfPi : kTwoPi) * r.uniform();
        ph.x = rp * std::cos(phi); ph.y = rp * std::sin(phi); ph.z = g.zmin;
      }
      break;
    default:  // :126-131
      ph.x = par.xs_point; ph.y = par.ys_point; ph.z = par.zs_point;
  }
  // setup_isotropic_injection :342-408
  ph.wgt = 1.0;
  if (w.sym) {  // :356-360 — sources are folded into the octant
    if (ph.x < g.xmin) ph.x = -ph.x;
    if (ph.y < g.ymin) ph.y = -ph.y;
    if (ph.z < g.zmin) ph.z = -ph.z;
  }
  double cost, sint, cosp, sinp;
  if (par.source_geometry == LART_SRC_PLANE_ILLUMINATION) {  // :765-778 a parallel beam, down onto the slab or up along +z
    cost = (w.atm == LART_ATM_PLANE) ? -1.0 : 1.0;
    sint = 0.0; cosp = 1.0; sinp = 0.0;
  } else {
    double uc, up;
    r.uniform2(uc, up);
    cost = 2.0 * uc - 1.0;
    sint = std::sqrt(1.0 - cost * cost);
    double phi = kTwoPi * up;
    cosp = std::cos(phi); sinp = std::sin(phi);
  }
  ph.kx = sint * cosp; ph.ky = sint * sinp; ph.kz = cost;
  if (w.amr) {  // generate_photon.f90:375-376
    ph.icell = amr_find_leaf(w, ph.x, ph.y, ph.z); ph.jcell = 1; ph.kcell = 1;
  } else {
  ph.icell = static_cast<int>(std::floor((ph.x - g.xmin) / g.dx)) + 1;
  ph.jcell = static_cast<int>(std::floor((ph.y - g.ymin) / g.dy)) + 1;
  ph.kcell = static_cast<int>(std::floor((ph.z - g.zmin) / g.dz)) + 1;
  if (ph.kx < 0.0 && ph.icell == g.nx + 1) ph.icell = g.nx;
  if (ph.ky < 0.0 && ph.jcell == g.ny + 1) ph.jcell = g.ny;
  if (ph.kz < 0.0 && ph.kcell == g.nz + 1) ph.kcell = g.nz;
  if (ph.kx > 0.0 && ph.icell < 1) ph.icell = 1;
  if (ph.ky > 0.0 && ph.jcell < 1) ph.jcell = 1;
  if (ph.kz > 0.0 && ph.kcell < 1) ph.kcell = 1;
  }
  if (par.use_stokes) {
    ph.mx = cost * cosp; ph.my = cost * sinp; ph.mz = -sint;
    ph.nx = -sinp; ph.ny = cosp; ph.nz = 0.0;
    ph.I = 1.0; ph.Q = 0.0; ph.U = 0.0; ph.V = 0.0;
  }
  // :134-157
  ph.nscatt_gas = 0.0; ph.nscatt_dust = 0.0; ph.inside = true;
  ph.xfreq = par.xfreq0;
  double Dloc = w.Dfreq(ph.icell, ph.jcell, ph.kcell), aloc = w.voigt_a(ph.icell, ph.jcell, ph.kcell);
  // :243-300
  switch (par.spectral_type) {
    case LART_SPEC_CONTINUUM:
      ph.xfreq = r.uniform() * (g.xfreq_max - g.xfreq_min) + g.xfreq_min;
      ph.xfreq = ph.xfreq / (Dloc / g.Dfreq_ref);
      break;
    case LART_SPEC_VOIGT0:
      ph.xfreq = ph.xfreq + rand_voigt(r, par.voigt_a0) * par.Dfreq0 / Dloc;
      break;
    case LART_SPEC_VOIGT:
      ph.xfreq = ph.xfreq + rand_voigt(r, aloc);
      break;
    case LART_SPEC_GAUSSIAN:
      ph.xfreq = ph.xfreq + r.gauss() * par.gaussian_sigma_x;
      ph.xfreq = ph.xfreq / (Dloc / g.Dfreq_ref);
      break;
    default:
      break;
  }
  double u1 = w.vdotk(ph.icell, ph.jcell, ph.kcell, ph.kx, ph.ky, ph.kz);
  if (!par.comoving_source) ph.xfreq = ph.xfreq - u1;  // :303-306
  if (par.save_Jin) {                                  // :309-322
    double xlab = (ph.xfreq + u1) * (Dloc / g.Dfreq_ref);
    int ix = static_cast<int>(std::floor((xlab - g.xfreq_min) / g.dxfreq)) + 1;
    if (ix >= 1 && ix <= g.nxfreq) tl.Jin[ix - 1] += ph.wgt;
  }
  if (w.cl && !w.cl->has_overlap) {  // :325-332 — the birth clump, before the direct peel
    const int64_t icl = clump_at_point(*w.cl, ph.x, ph.y, ph.z);
    ph.icl = static_cast<int>(icl);
    if (icl > 0) ph.xfreq = ph.xfreq - ulos_clump(*w.cl, icl, ph.kx, ph.ky, ph.kz);
  }
  if (par.save_peeloff) peeling_direct(w, ph, tl);  // :334-336
}

// impact radius of the photon's line w.r.t. the origin — run_simulation_mod.f90:302-330
inline double impact_radius(const World &w, const Photon &ph, double &mx, double &my, double &mz) {
  double rmax = w.g->rmax;
  double xp = ph.x, yp = ph.y, zp = ph.z;
  if (rmax > 0.0) {
    double rr = ph.x * ph.x + ph.y * ph.y + ph.z * ph.z, dist = 0.0;
    if (rr > rmax * rmax) {
      double rk = ph.x * ph.kx + ph.y * ph.ky + ph.z * ph.kz;
      double det = rk * rk - (rr - rmax * rmax);
      dist = (det < 0.0) ? 0.0 : -rk + std::sqrt(std::fmax(0.0, det));
    }
    xp = ph.x + dist * ph.kx; yp = ph.y + dist * ph.ky; zp = ph.z + dist * ph.kz;
  }
  double rk = xp * ph.kx + yp * ph.ky + zp * ph.kz;
  mx = xp - rk * ph.kx; my = yp - rk * ph.ky; mz = zp - rk * ph.kz;
  return std::sqrt(mx * mx + my * my + mz * mz);
}

// make_all_initial_photons / make_all_photons — run_simulation_mod.f90:249-358
void record_initial(const World &w, const Photon &ph, lart_allph_out &ap) {
  if (ap.xfreq1) ap.xfreq1[ph.id - 1] = ph.xfreq;
  if (w.par->source_geometry != LART_SRC_POINT && ap.rp0) {
    double mx, my, mz;
    ap.rp0[ph.id - 1] = impact_radius(w, ph, mx, my, mz);
  }
}
void record_final(const World &w, const Photon &ph, lart_allph_out &ap) {
  double mx, my, mz;
  double mm = impact_radius(w, ph, mx, my, mz);
  size_t s = static_cast<size_t>(ph.id - 1);
  if (ap.rp) ap.rp[s] = mm;
  if (ap.xfreq2) ap.xfreq2[s] = ph.xfreq_ref;
  if (ap.nscatt_gas) ap.nscatt_gas[s] = ph.nscatt_gas;
  if (ap.nscatt_dust) ap.nscatt_dust[s] = ph.nscatt_dust;
  if (w.par->use_stokes) {
    double cos2p = 1.0, sin2p = 0.0;
    if (mm > 0.0) {
      mx /= mm; my /= mm; mz /= mm;
      double cosp = mx * ph.mx + my * ph.my + mz * ph.mz, sinp = mx * ph.nx + my * ph.ny + mz * ph.nz;
      cos2p = 2.0 * cosp * cosp - 1.0;
      sin2p = 2.0 * sinp * cosp;
    }
    if (ap.I) ap.I[s] = ph.wgt;
    if (ap.Q) ap.Q[s] = (cos2p * ph.Q + sin2p * ph.U) * ph.wgt;
    if (ap.U) ap.U[s] = (-sin2p * ph.Q + cos2p * ph.U) * ph.wgt;
    if (ap.V) ap.V[s] = ph.V * ph.wgt;
  }
}

// add_escaped_fraction_to_Jout — run_simulation_mod.f90:208-247
void add_escaped_fraction(const World &w, const Photon &ph, double tau0, Tally &tl) {
  const lart_grid &g = *w.g;
  double wgt_esc = ph.wgt * std::exp(-tau0);
  double u1 = w.vdotk(ph.icell, ph.jcell, ph.kcell, ph.kx, ph.ky, ph.kz);
  double xref = (ph.xfreq + u1) * (w.Dfreq(ph.icell, ph.jcell, ph.kcell) / g.Dfreq_ref);
  int ix = static_cast<int>(std::floor((xref - g.xfreq_min) / g.dxfreq)) + 1;
  if (ix >= 1 && ix <= g.nxfreq) {
    tl.Jout[ix - 1] += wgt_esc;
    if (w.par->save_Jmu) tl.Jmu[(ix - 1) + static_cast<size_t>(g.nxfreq) * (jmu_bin(*w.par, ph.kz) - 1)] += wgt_esc;
  }
}

// one photon, start to finish — run_simulation_mod.f90:155-196
// max_events > 0 abandons the photon after that many scatterings (the bounded runs of lart_config::max_events:
// CPU-baseline samples, and bounded parity runs at sizes where a photon needs ~1e7 scatterings to escape).  An
// abandoned photon is recorded as it stands, its current frequency in the allph%xfreq2 slot; it makes no Jout tally.
void run_photon(const World &w, int64_t id, Rng &r, Tally &tl, int64_t max_events) {
  Photon ph;
  ph.id = id;
  r.start_stream(static_cast<uint64_t>(id));
  generate_photon(w, ph, r, tl);
  if (w.par->save_all_photons) record_initial(w, ph, tl.shared->allph);
  bool first = true;
  int64_t nev = 0;
  while (ph.inside) {
    double tau;
    if (first) {  // :163-177 forced first scattering
      double tau0 = edge_tau(w, ph, &tl.cnt);
      add_escaped_fraction(w, ph, tau0, tl);
      double wgt1 = 1.0 - std::exp(-tau0);
      ph.wgt = ph.wgt * wgt1;
      tau = (tau0 > 0.0) ? -std::log(1.0 - r.uniform() * wgt1) : kHugest;
      first = false;
    } else {
      tau = -std::log(r.uniform());
    }
    if (w.cl && w.cl->has_overlap) raytrace_to_tau_clump_overlap(w, ph, tau, r, &tl, &tl.cnt);
    else if (w.cl) raytrace_to_tau_clump(w, ph, tau, &tl, &tl.cnt);
    else if (w.amr) raytrace_to_tau_amr(w, ph, tau, &tl, &tl.cnt);
    else raytrace_to_tau(w, ph, tau, &tl, &tl.cnt);
    if (ph.inside) {
      scattering(w, ph, r, tl);
      if (max_events > 0 && ++nev >= max_events) break;
    }
  }
  tl.nscatt_gas += ph.nscatt_gas;
  tl.nscatt_dust += ph.nscatt_dust;
  tl.cnt.n_photons_done += 1;
  if (max_events > 0 && ph.inside) ph.xfreq_ref = ph.xfreq;
  if (w.par->save_all_photons) record_final(w, ph, tl.shared->allph);
}

// raytrace_to_edge_car_tau_gas — raytrace_car.f90:1236-1328 (gas only, no tau cap)
double raytrace_to_edge_tau_gas(const World &w, const Photon &p0, long long *nsteps) {
  const lart_grid &g = *w.g;
  double xp = p0.x, yp = p0.y, zp = p0.z, kx = p0.kx, ky = p0.ky, kz = p0.kz;
  int ic = p0.icell, jc = p0.jcell, kc = p0.kcell;
  double tau = 0.0, d = 0.0;
  Trav t;
  if (setup_traversal(w, xp, yp, zp, kx, ky, kz, ic, jc, kc, t, false)) return tau;
  int io = ic, jo = jc, ko = kc;
  double u1 = w.vdotk(io, jo, ko, kx, ky, kz);
  double xfreq = p0.xfreq;
  for (;;) {
    if (w.masked(ic, jc, kc)) return std::numeric_limits<double>::infinity();  // _tau_gas_atmosphere, :3832-3836
    double rhokap = w.rhokap(ic, jc, kc) * w.calc_voigt(xfreq, ic, jc, kc);
    if (nsteps) ++*nsteps;
    int m = w.zonly ? 3 : minloc3(t.tx, t.ty, t.tz);
    if (m == 1) { tau += (t.tx - d) * rhokap; d = t.tx; ic += t.istep; if (ic < 1 || ic > g.nx) break; t.tx += t.delx; }
    else if (m == 2) { tau += (t.ty - d) * rhokap; d = t.ty; jc += t.jstep; if (jc < 1 || jc > g.ny) break; t.ty += t.dely; }
    else { tau += (t.tz - d) * rhokap; d = t.tz; kc += t.kstep; if (kc < 1 || kc > g.nz) break; t.tz += t.delz; }
    double u2 = w.vdotk(ic, jc, kc, kx, ky, kz);
    xfreq = (xfreq + u1) * w.Dfreq(io, jo, ko) / w.Dfreq(ic, jc, kc) - u2;
    io = ic; jo = jc; ko = kc;
    u1 = u2;
  }
  return tau;
}

// raytrace_to_edge_car_column — raytrace_car.f90:1330-1423
void raytrace_to_edge_column(const World &w, const Photon &p0, double cross0, double &N_gas, double &tau_dust, long long *nsteps) {
  const lart_grid &g = *w.g;
  double xp = p0.x, yp = p0.y, zp = p0.z, kx = p0.kx, ky = p0.ky, kz = p0.kz;
  int ic = p0.icell, jc = p0.jcell, kc = p0.kcell;
  N_gas = 0.0; tau_dust = 0.0;
  double d = 0.0;
  Trav t;
  if (setup_traversal(w, xp, yp, zp, kx, ky, kz, ic, jc, kc, t, false)) return;
  for (;;) {
    if (w.masked(ic, jc, kc)) {  // _column_atmosphere, :3933-3938
      N_gas = std::numeric_limits<double>::infinity();
      if (w.dust()) tau_dust = std::numeric_limits<double>::infinity();
      return;
    }
    double rho = w.rhokap(ic, jc, kc) * w.Dfreq(ic, jc, kc) / cross0;
    double rkD = w.dust() ? w.rhokapD(ic, jc, kc) : 0.0;
    if (nsteps) ++*nsteps;
    int m = w.zonly ? 3 : minloc3(t.tx, t.ty, t.tz);
    double tn = (m == 1) ? t.tx : (m == 2) ? t.ty : t.tz;
    double del = tn - d;
    N_gas += del * rho;
    if (w.dust()) tau_dust += del * rkD;
    d = tn;
    if (m == 1) { ic += t.istep; if (ic < 1 || ic > g.nx) break; t.tx += t.delx; }
    else if (m == 2) { jc += t.jstep; if (jc < 1 || jc > g.ny) break; t.ty += t.dely; }
    else { kc += t.kstep; if (kc < 1 || kc > g.nz) break; t.tz += t.delz; }
  }
}

// Entry point and start cell of the sight line of pixel (ix,iy) of one observer —
// sightline_tau_rect.f90:45-150.  Returns false when the line misses the grid.
bool sightline_start(const lart_grid &g, const lart_observer &ob, int ix, int iy, Photon &po) {
  double kx = std::tan((ix - (ob.nxim + 1.0) / 2.0) * ob.dxim / kRad2Deg);
  double ky = std::tan((iy - (ob.nyim + 1.0) / 2.0) * ob.dyim / kRad2Deg);
  double kz = -1.0;
  double kr = std::sqrt(kx * kx + ky * ky + kz * kz);
  kx /= kr; ky /= kr; kz /= kr;
  const double *R = ob.rmatrix;  // R[(r-1)+3*(c-1)]: transpose applied here (:57-59)
  po.kx = R[0] * kx + R[1] * ky + R[2] * kz;
  po.ky = R[3] * kx + R[4] * ky + R[5] * kz;
  po.kz = R[6] * kx + R[7] * ky + R[8] * kz;
  double delt[6];
  delt[0] = (po.kx == 0.0) ? kHugest : (g.xmax - ob.x) / po.kx;
  delt[1] = (po.kx == 0.0) ? kHugest : (g.xmin - ob.x) / po.kx;
  delt[2] = (po.ky == 0.0) ? kHugest : (g.ymax - ob.y) / po.ky;
  delt[3] = (po.ky == 0.0) ? kHugest : (g.ymin - ob.y) / po.ky;
  delt[4] = (po.kz == 0.0) ? kHugest : (g.zmax - ob.z) / po.kz;
  delt[5] = (po.kz == 0.0) ? kHugest : (g.zmin - ob.z) / po.kz;
  auto cellof = [&](double x, double y, double z, int &i, int &j, int &k) {
    i = static_cast<int>(std::floor((x - g.xmin) / g.dx)) + 1;
    j = static_cast<int>(std::floor((y - g.ymin) / g.dy)) + 1;
    k = static_cast<int>(std::floor((z - g.zmin) / g.dz)) + 1;
  };
  double dist = -999.9;
  int j0 = 0;
  for (int jj = 1; jj <= 6; ++jj) {  // the farthest boundary the ray touches (:79-106)
    double dl = delt[jj - 1];
    if (dl > 0.0 && dl < kHugest) {
      int i, j, k;
      cellof(ob.x + po.kx * dl, ob.y + po.ky * dl, ob.z + po.kz * dl, i, j, k);
      if (jj == 1) i = g.nx + 1;
      if (jj == 2) i = 1;
      if (jj == 3) j = g.ny + 1;
      if (jj == 4) j = 1;
      if (jj == 5) k = g.nz + 1;
      if (jj == 6) k = 1;
      if (i >= 1 && i <= g.nx + 1 && j >= 1 && j <= g.ny + 1 && k >= 1 && k <= g.nz + 1 && dl > dist) { dist = dl; j0 = jj; }
    }
  }
  if (!(dist > 0.0 && dist < kHugest)) return false;
  po.x = ob.x + po.kx * dist; po.y = ob.y + po.ky * dist; po.z = ob.z + po.kz * dist;
  cellof(po.x, po.y, po.z, po.icell, po.jcell, po.kcell);
  po.kx = -po.kx; po.ky = -po.ky; po.kz = -po.kz;  // toward the observer (:116-118)
  if (j0 == 1) { po.icell = g.nx + 1; po.x = g.xface[g.nx]; }
  else if (j0 == 2) { po.icell = 1; po.x = g.xface[0]; }
  else if (j0 == 3) { po.jcell = g.ny + 1; po.y = g.yface[g.ny]; }
  else if (j0 == 4) { po.jcell = 1; po.y = g.yface[0]; }
  else if (j0 == 5) { po.kcell = g.nz + 1; po.z = g.zface[g.nz]; }
  else if (j0 == 6) { po.kcell = 1; po.z = g.zface[0]; }
  if (po.icell == g.nx + 1 && po.kx < 0.0) po.icell = g.nx;
  if (po.jcell == g.ny + 1 && po.ky < 0.0) po.jcell = g.ny;
  if (po.kcell == g.nz + 1 && po.kz < 0.0) po.kcell = g.nz;
  return po.icell >= 1 && po.icell <= g.nx && po.jcell >= 1 && po.jcell <= g.ny && po.kcell >= 1 && po.kcell <= g.nz;
}

std::string g_err;

World make_world(const lart_config *cfg) {
  World w;
  w.g = &cfg->grid;
  w.par = &cfg->par;
  w.line = &cfg->line;
  w.sm = &cfg->scatt_mat;
  w.obs = cfg->observers;
  w.zonly = cfg->par.xy_periodic && cfg->grid.nx == 1 && cfg->grid.ny == 1;  // setup.f90:957-965
  w.sym = cfg->par.xyz_symmetry != 0;                                         // setup.f90:952-954
  if (w.sym) { w.bcxy = 1; w.bcz = 1; }
  else if (cfg->par.xy_symmetry) w.bcxy = 1;                                  // :955-957
  else if (cfg->par.xy_periodic && !w.zonly) w.bcxy = 2;                      // :966-975 (no shear)
  if (cfg->par.use_clump_medium && cfg->clumps.n > 0) w.cl = &cfg->clumps;    // :806-860
  if (cfg->par.use_amr_grid && cfg->amr.nleaf > 0) w.amr = &cfg->amr;         // the octree ray tracers and leaf physics
  w.shear = w.bcxy == 2 && cfg->par.Omega != 0.0;                             // :967-969
  w.atm = cfg->par.atmosphere;                                                // :959-962, :977-987
  w.jp = cfg->par.calc_J || cfg->par.calc_P || cfg->par.calc_Pnew;
  return w;
}

}  // namespace

// ===========================================================================
// C interface (ctypes) — mirrors the batched entry points of include/lart_gpu.h
// ===========================================================================
extern "C" {

const char *oracle_last_error(void) { return g_err.c_str(); }

int oracle_voigt(int64_t n, const double *x, const double *a, double *H) {
  for (int64_t i = 0; i < n; ++i) H[i] = voigt_seon2(x[i], a[i]);
  return 0;
}

int oracle_raytrace_edge(const lart_config *cfg, int64_t n, const double *x, const double *y, const double *z,
                         const double *kx, const double *ky, const double *kz, const double *xfreq,
                         const int32_t *ic, const int32_t *jc, const int32_t *kc, double *tau, int32_t *nsteps,
                         int32_t trace_cap, int32_t *trace) {
  World w = make_world(cfg);
  for (int64_t i = 0; i < n; ++i) {
    Photon p;
    p.x = x[i]; p.y = y[i]; p.z = z[i]; p.kx = kx[i]; p.ky = ky[i]; p.kz = kz[i];
    p.xfreq = xfreq[i]; p.icell = ic[i]; p.jcell = jc[i]; p.kcell = kc[i];
    int ns = 0;
    tau[i] = raytrace_to_edge(w, p, nullptr, &ns, trace_cap, trace ? trace + i * static_cast<int64_t>(trace_cap) : nullptr);
    if (nsteps) nsteps[i] = ns;
  }
  return 0;
}

int oracle_raytrace_tau(const lart_config *cfg, int64_t n, double *x, double *y, double *z, const double *kx,
                        const double *ky, const double *kz, double *xfreq, int32_t *ic, int32_t *jc, int32_t *kc,
                        const double *tau_in, int32_t *inside, double *xfreq_ref, int32_t *nsteps) {
  World w = make_world(cfg);
  for (int64_t i = 0; i < n; ++i) {
    Photon p;
    p.x = x[i]; p.y = y[i]; p.z = z[i]; p.kx = kx[i]; p.ky = ky[i]; p.kz = kz[i];
    p.xfreq = xfreq[i]; p.icell = ic[i]; p.jcell = jc[i]; p.kcell = kc[i];
    p.inside = true;
    int ns = 0;
    raytrace_to_tau(w, p, tau_in[i], nullptr, nullptr, &ns);
    x[i] = p.x; y[i] = p.y; z[i] = p.z; xfreq[i] = p.xfreq;
    ic[i] = p.icell; jc[i] = p.jcell; kc[i] = p.kcell;
    inside[i] = p.inside ? 1 : 0;
    if (xfreq_ref) xfreq_ref[i] = p.inside ? 0.0 : p.xfreq_ref;
    if (nsteps) nsteps[i] = ns;
  }
  return 0;
}

// clump-medium ray tracers, unit level (the oracle side of lart_gpu_clump_*_batch)
int oracle_clump_edge(const lart_config *cfg, int64_t n, const double *x, const double *y, const double *z, const double *kx,
                      const double *ky, const double *kz, const double *xfreq, const int32_t *icl, double tau_max, double *tau,
                      int32_t *nclumps) {
  World w = make_world(cfg);
  if (!w.cl) { g_err = "oracle_clump_edge: no clump medium"; return 1; }
  for (int64_t i = 0; i < n; ++i) {
    Photon p;
    p.x = x[i]; p.y = y[i]; p.z = z[i]; p.kx = kx[i]; p.ky = ky[i]; p.kz = kz[i]; p.xfreq = xfreq[i]; p.icl = icl[i];
    int nc = 0;
    tau[i] = w.cl->has_overlap ? raytrace_to_edge_clump_overlap(w, p, tau_max, nullptr) : raytrace_to_edge_clump(w, p, tau_max, nullptr, &nc);
    if (nclumps) nclumps[i] = nc;
  }
  return 0;
}
int oracle_clump_tau(const lart_config *cfg, int64_t n, double *x, double *y, double *z, const double *kx, const double *ky,
                     const double *kz, double *xfreq, int32_t *icl, const double *tau_in, int32_t *inside) {
  World w = make_world(cfg);
  if (!w.cl) { g_err = "oracle_clump_tau: no clump medium"; return 1; }
  for (int64_t i = 0; i < n; ++i) {
    Photon p;
    p.x = x[i]; p.y = y[i]; p.z = z[i]; p.kx = kx[i]; p.ky = ky[i]; p.kz = kz[i]; p.xfreq = xfreq[i]; p.icl = icl[i];
    p.inside = true;
    raytrace_to_tau_clump(w, p, tau_in[i], nullptr, nullptr);
    x[i] = p.x; y[i] = p.y; z[i] = p.z; xfreq[i] = p.xfreq; icl[i] = p.icl; inside[i] = p.inside ? 1 : 0;
  }
  return 0;
}
int oracle_clump_locate(const lart_config *cfg, int64_t n, const double *x, const double *y, const double *z, int32_t *icl) {
  World w = make_world(cfg);
  if (!w.cl) { g_err = "oracle_clump_locate: no clump medium"; return 1; }
  for (int64_t i = 0; i < n; ++i) icl[i] = static_cast<int32_t>(clump_at_point(*w.cl, x[i], y[i], z[i]));
  return 0;
}

// octree: raytrace_to_edge_amr / raytrace_to_tau_amr / amr_find_leaf on arrays (il = leaf index, <= 0: locate first)
int oracle_amr_edge(const lart_config *cfg, int64_t n, const double *x, const double *y, const double *z, const double *kx,
                    const double *ky, const double *kz, const double *xfreq, const int32_t *il, double *tau, int32_t *nsteps) {
  World w = make_world(cfg);
  if (!w.amr) { g_err = "oracle_amr_edge: no octree in the configuration"; return 1; }
  for (int64_t i = 0; i < n; ++i) {
    Photon p;
    p.x = x[i]; p.y = y[i]; p.z = z[i]; p.kx = kx[i]; p.ky = ky[i]; p.kz = kz[i]; p.xfreq = xfreq[i]; p.icell = il[i];
    int ns = 0;
    tau[i] = raytrace_to_edge_amr(w, p, nullptr, &ns);
    if (nsteps) nsteps[i] = ns;
  }
  return 0;
}
int oracle_amr_tau(const lart_config *cfg, int64_t n, double *x, double *y, double *z, const double *kx, const double *ky,
                   const double *kz, double *xfreq, int32_t *il, const double *tau_in, int32_t *inside, double *xfreq_ref,
                   int32_t *nsteps) {
  World w = make_world(cfg);
  if (!w.amr) { g_err = "oracle_amr_tau: no octree in the configuration"; return 1; }
  for (int64_t i = 0; i < n; ++i) {
    Photon p;
    p.x = x[i]; p.y = y[i]; p.z = z[i]; p.kx = kx[i]; p.ky = ky[i]; p.kz = kz[i]; p.xfreq = xfreq[i]; p.icell = il[i];
    int ns = 0;
    raytrace_to_tau_amr(w, p, tau_in[i], nullptr, nullptr, &ns);
    x[i] = p.x; y[i] = p.y; z[i] = p.z; xfreq[i] = p.xfreq; il[i] = p.icell; inside[i] = p.inside ? 1 : 0;
    if (xfreq_ref) xfreq_ref[i] = p.xfreq_ref;
    if (nsteps) nsteps[i] = ns;
  }
  return 0;
}
int oracle_amr_locate(const lart_config *cfg, int64_t n, const double *x, const double *y, const double *z, int32_t *il) {
  World w = make_world(cfg);
  if (!w.amr) { g_err = "oracle_amr_locate: no octree in the configuration"; return 1; }
  for (int64_t i = 0; i < n; ++i) il[i] = amr_find_leaf(w, x[i], y[i], z[i]);
  return 0;
}

int oracle_xcrit(const lart_config *cfg, int64_t n, const double *x, const double *y, const double *z,
                 const int32_t *ic, const int32_t *jc, const int32_t *kc, double *xcrit) {
  World w = make_world(cfg);
  for (int64_t i = 0; i < n; ++i) {
    double a, b;
    car_xcrit_local(w, ic[i], jc[i], kc[i], x[i], y[i], z[i], a, b);
    xcrit[i] = a;
  }
  return 0;
}

// kinds as in lart_gpu_sample_batch.  rng_mode 1: Philox stream ids[i];
// rng_mode 0: one MT19937-64 stream seeded with `seed`, elements in order.
int oracle_sample(int32_t kind, int32_t rng_mode, uint64_t seed, int64_t n, const int64_t *ids, const double *p0,
                  const double *p1, int32_t ndraw, double *out) {
  Rng r;
  r.mode = rng_mode;
  r.seed = seed;
  if (rng_mode == 0) r.mt.init(static_cast<int64_t>(seed));
  for (int64_t i = 0; i < n; ++i) {
    if (rng_mode == 1) r.start_stream(static_cast<uint64_t>(ids ? ids[i] : i));
    for (int j = 0; j < ndraw; ++j) {
      double v;
      switch (kind) {
        case 0: v = r.uniform(); break;
        case 1: v = r.gauss(); break;
        case 2: v = rand_resonance_vz(r, p0[i], p1[i]); break;
        case 3: v = rand_resonance(r, p0[i]); break;
        case 4: v = rand_hg(r, p0[i]); break;
        case 5: v = rand_voigt(r, p0[i]); break;
        default: g_err = "oracle_sample: unknown kind"; return 1;
      }
      out[i * static_cast<int64_t>(ndraw) + j] = v;
    }
  }
  return 0;
}

// raw generator words for known-answer tests
int oracle_philox_block(uint64_t seed, uint64_t stream, uint64_t block, uint32_t out[4]) {
  uint32_t c[4] = {static_cast<uint32_t>(stream), static_cast<uint32_t>(stream >> 32), static_cast<uint32_t>(block), static_cast<uint32_t>(block >> 32)};
  philox4x32_10(c, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  for (int i = 0; i < 4; ++i) out[i] = c[i];
  return 0;
}
int oracle_philox_raw(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
  philox4x32_10(c, key[0], key[1]);
  for (int i = 0; i < 4; ++i) out[i] = c[i];
  return 0;
}
int oracle_mt64_words(int64_t seed, int64_t n, uint64_t *out) {
  Mt64 m;
  m.init(seed);
  for (int64_t i = 0; i < n; ++i) out[i] = m.next();
  return 0;
}

// Whole run: photons id = first_id + t*stride ... dealt to `nthreads` threads
// the way run_equal_number deals them to ranks (run_simulation_mod.f90:150).
// rng_mode 0 = MT19937-64 seeded seed + 9999*thread (random_mt.f90:941-953);
// rng_mode 1 = Philox keyed (seed, photon id) — thread-count independent.
int oracle_run(const lart_config *cfg, int32_t rng_mode, int32_t nthreads, int64_t first_id, int64_t count, int64_t stride,
               int64_t max_events_per_photon, lart_tallies *out) {
  if (cfg->line.line_type != 1) { g_err = "oracle_run: only line_type 1"; return 1; }
  if (cfg->par.nobs > 0 && (!out->obs || !cfg->observers)) { g_err = "oracle_run: observers/outputs missing"; return 1; }
  World w = make_world(cfg);
  if (nthreads < 1) nthreads = 1;
  const int nxf = cfg->grid.nxfreq;
  std::vector<Tally> tls(nthreads);
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t) {
    Tally &tl = tls[t];
    tl.Jout.assign(nxf, 0.0);
    tl.Jin.assign(nxf, 0.0);
    tl.Jabs.assign(nxf, 0.0);
    tl.Jabs2.assign(nxf, 0.0);
    tl.Jmu.assign(static_cast<size_t>(nxf) * (cfg->par.save_Jmu ? cfg->par.nmu : 0), 0.0);
    tl.shared = out;
    th.emplace_back([&, t]() {
      Rng r;
      r.mode = rng_mode;
      r.seed = cfg->par.seed;
      r.cnt = &tls[t].cnt;
      if (rng_mode == 0) r.mt.init(static_cast<int64_t>(cfg->par.seed) + 9999 * t);
      for (int64_t q = t; q < count; q += nthreads) run_photon(w, first_id + q * stride, r, tls[t], max_events_per_photon);
    });
  }
  for (auto &x : th) x.join();
  for (int t = 0; t < nthreads; ++t) {  // reduce_mem — memory_mod_mpi.f90:366-458
    Tally &tl = tls[t];
    for (int i = 0; i < nxf; ++i) {
      if (out->Jout) out->Jout[i] += tl.Jout[i];
      if (out->Jin) out->Jin[i] += tl.Jin[i];
      if (out->Jabs) out->Jabs[i] += tl.Jabs[i];
      if (out->Jabs2) out->Jabs2[i] += tl.Jabs2[i];
    }
    if (out->Jmu) for (size_t i = 0; i < tl.Jmu.size(); ++i) out->Jmu[i] += tl.Jmu[i];
    out->nscatt_gas += tl.nscatt_gas;
    out->nscatt_dust += tl.nscatt_dust;
    out->counters.n_photons_done += tl.cnt.n_photons_done;
    out->counters.n_scatter += tl.cnt.n_scatter;
    out->counters.n_cellsteps += tl.cnt.n_cellsteps;
    out->counters.n_peel += tl.cnt.n_peel;
    out->counters.n_rng += tl.cnt.n_rng;
    out->counters.n_reject_iter += tl.cnt.n_reject_iter;
  }
  return 0;
}

// make_sightline_tau_outside — sightline_tau_rect.f90:11-190.  out[k] as lart_sightline_out.
int oracle_sightline_tau(const lart_config *cfg, double cross0, lart_sightline_out *out, double *cellsteps) {
  World w = make_world(cfg);
  const lart_grid &g = cfg->grid;
  long long ns = 0;
  for (int i = 0; i < cfg->par.nobs; ++i) {
    const lart_observer &ob = cfg->observers[i];
    for (int iy = 1; iy <= ob.nyim; ++iy)
      for (int ix = 1; ix <= ob.nxim; ++ix) {
        Photon po;
        size_t pix = static_cast<size_t>(ix - 1) + static_cast<size_t>(ob.nxim) * (iy - 1);
        if (!sightline_start(g, ob, ix, iy, po)) continue;
        double u1 = w.vdotk(po.icell, po.jcell, po.kcell, po.kx, po.ky, po.kz);
        double Dc = w.Dfreq(po.icell, po.jcell, po.kcell);
        for (int kk = 1; kk <= g.nxfreq; ++kk) {
          double xf = (kk - 0.5) * g.dxfreq + g.xfreq_min;  // grid%xfreq(kk), grid_mod_car.f90:1505
          po.xfreq = xf * g.Dfreq_ref / Dc - u1;
          out[i].tau_gas[(kk - 1) + static_cast<size_t>(g.nxfreq) * pix] = raytrace_to_edge_tau_gas(w, po, &ns);
        }
        double N, td;
        raytrace_to_edge_column(w, po, cross0, N, td, &ns);
        out[i].N_gas[pix] = N;
        if (out[i].tau_dust && w.dust()) out[i].tau_dust[pix] = td;
      }
  }
  if (cellsteps) *cellsteps = static_cast<double>(ns);
  return 0;
}

int oracle_hardware_threads(void) { return static_cast<int>(std::thread::hardware_concurrency()); }

}  // extern "C"
