// lart_device.cuh — device-side physics of the Cartesian photon loop (sm_100a).
//
// Building blocks shared by every kernel in lart_engine.cu: Philox4x32-10 stream
// per photon, voigt_seon2, the DDA cell walk, the random variates, scattering,
// emission and the peel-off descriptors.  Each block cites the reference lines it
// reproduces (paths relative to the reference tree).  All arithmetic is FP64 as in
// the reference (define.f90:25); the deterministic pieces (Voigt, DDA) use explicit
// round-to-nearest intrinsics so that no FMA contraction can change a bit relative
// to the plain IEEE evaluation order of the Fortran expressions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lart {

#define LART_DEV __device__ __forceinline__
#define DMUL(a, b) __dmul_rn((a), (b))
#define DADD(a, b) __dadd_rn((a), (b))
#define DSUB(a, b) __dsub_rn((a), (b))

constexpr double kPi = 3.141592653589793238462643383279502884197;
constexpr double kTwoPi = 6.283185307179586476925286766559005768394;
constexpr double kFourPi = 12.56637061435917295385057353311801153679;
constexpr double kHalfPi = kPi / 2.0;
constexpr double kRad2Deg = 180.0 / kPi;
constexpr double kHugest = 1.7976931348623157e308;  // huge(1d0), define.f90:38
constexpr double kTauHuge = 745.2;                  // raytrace_car.f90:432

// One grid cell, packed so that a DDA step touches two 32-byte sectors instead of
// six (SoA).  Built on the device from the host's SoA arrays at create time.
struct __align__(64) Cell {
  double rhokap, voigt_a, Dfreq, vfx, vfy, vfz, rhokapD, pad;
};

struct DevObserver {
  double x, y, z;
  double R[9];  // Fortran order: R[(r-1)+3*(c-1)]
  double dxim, dyim;
  int nxim, nyim;
};

// offsets (in doubles) inside the contiguous tally buffer; -1 = not saved
struct TallyLayout {
  long long Jout, Jin, Jabs, Jmu;
  long long obs_base, obs_stride;  // observer k starts at obs_base + k*obs_stride
  long long cube[7];               // scatt, direc, direc0, I, Q, U, V   (within an observer block)
  long long img[7];                // the same, 2-D
  long long scalars;               // nscatt_gas, nscatt_dust
  long long counters;              // C_COUNT work counters
  long long Jabs2, J, Pa, Pnew;    // atmosphere absorption spectrum; CALCJ / CALCP / CALCPnew accumulators (-1 = off)
  long long total;
};
enum { T_SCATT = 0, T_DIREC = 1, T_DIREC0 = 2, T_I = 3, T_Q = 4, T_U = 5, T_V = 6 };
enum { C_PHOTONS = 0, C_SCATTER = 1, C_CELLSTEPS = 2, C_PEEL = 3, C_RNG = 4, C_REJECT = 5, C_PEEL_BOUND = 6, C_COUNT = 8 };
// sticky device error word (lart_gpu_step / _sync / _fetch return it as an error): work that could not be queued
enum { ERR_DIRECT_QUEUE = 1, ERR_CONT_QUEUE = 2, ERR_BAD_STATE = 4 };

// Everything a kernel needs, passed by value as a __grid_constant__ parameter.
// clump medium (lart_clump.cuh): geometry record = centre + radius^2 (one 32-byte sector per ray-sphere test),
// physics record = 64 bytes read once per clump crossed
struct __align__(16) ClumpPhys { double rhokap, rhokapD, voigt_a, Dfreq, vx, vy, vz, pad_; };
struct DevClumps {
  long long n;
  double sphere_R, R2, Dfreq_ref;
  const double4 *geo;
  const double4 *geo_reg;  // geo[cg_list[ip]] for every CSR registration ip: the cell's clumps are contiguous, one load level less
  const ClumpPhys *phys;
  const int *cg_start, *cg_list;  // 1-based offsets / clump indices, as the host built them
  int cgx, cgy, cgz;
  int overlap;             // has_overlap: the event-walk ray tracers (raytrace_clump.f90:621-920), one thread per ray
  double xmin, ymin, zmin, dx, dy, dz, inv_dx, inv_dy, inv_dz;
  double *ov_t; int *ov_ev, *ov_act; long long ov_T;  // per-thread event lists and active sets of the overlap walk
};

// octree AMR: one 64-byte geometry record per LEAF (centre, half-width, the six face neighbours of its cell) —
// what a ray needs to leave the leaf — and one 64-byte record per CELL for the descent into a finer neighbour
// (octree_mod.f90:19-138).  The leaf physics sits in the packed `cells` array, indexed by leaf.
struct __align__(16) AmrGeo { double cx, cy, cz, h; int nb[6]; int icell, pad_; };
struct __align__(16) AmrCell { double cx, cy, cz; int child[8]; int ileaf, pad_; };
struct DevAmr {
  int on, ncells, nleaf, pad_;
  const AmrGeo *geo;    // [nleaf]
  const AmrCell *cell;  // [ncells]
};

// The less common bindings of the Cartesian ray tracers (setup.f90:959-987) and the mean-intensity /
// scattering-rate accumulators (8f-4).  `any` = 0 on every other run: each hook below sits behind a warp-uniform branch.
struct DevExtras {
  int any;
  int atm;            // LART_ATM_*: 1 = plane atmosphere (a photon leaving through the bottom cell goes to Jabs2),
                      // 2 = spherical atmosphere (cells with mask == -1 destroy the photon; rays through them have tau = +inf)
  int shear;          // shearing box: photon%vfy_shear changes by -+Omega at every x wrap (raytrace_car.f90:2842-2850)
  int edge_open;      // shear or atm: raytrace_to_edge stays the plain open-box routine (setup.f90:947-950, 984)
  int calc_J, calc_P, calc_Pnew, jp;  // jp = any of the three
  int geometry_JPa, nr;
  double Omega, cross0;
  const signed char *mask;   // (nx,ny,nz) Fortran order
  const int *ind_sph, *ind_cyl;
};

struct DevParams {
  // grid
  int nx, ny, nz, nxfreq;
  int nsbx, nsby, nsbz;  // 32^3 super-bricks of the packed cell array per axis
  double dx, dy, dz;
  double xmin, ymin, zmin, xmax, ymax, zmax;
  // boundary variants of the ray tracers (setup.f90:952-976): bcxy / bcz = BC_* of the x,y axes and of the z axis;
  // sym = par%xyz_symmetry (source fold, |kz| in Jmu); i0,j0,k0 = cell entered on reflection (grid_mod_car.f90:85-134)
  int sym, bcxy, bcz, i0, j0, k0;
  int clump;      // par%use_clump_medium: the ray tracers of lart_clump.cuh, photons carry their clump index
  DevAmr amr;     // par%use_amr_grid: photons carry a leaf index in `ic` (jc = kc = 1); nx = nleaf, ny = nz = 1
  DevClumps cl;
  DevExtras x;
  double Dfreq_ref, xfreq_min, xfreq_max, dxfreq, xcrit, xcrit2, rmax;
  const double *xface, *yface, *zface;
  const Cell *cells;                                            // packed records (default walk)
  const double *rhokap, *voigt_a, *Dfreq, *vfx, *vfy, *vfz, *rhokapD;  // SoA (ablation / batch API)
  const double *voigt_tab;  // [202][4] interleaved h0,h1,h2,h3
  // par / line
  unsigned long long seed;
  double xfreq0, xs, ys, zs, source_rmax, albedo, hgg, voigt_a0, Dfreq0, gaussian_sigma_x, mu_min, dmu;
  double E1, E2, E3, g_recoil0;
  double rr_p2, rr_inv;  // rand_resonance: sqrt((4-E1)/(3E1)) and 1/(E1 p2^3), E1 > 0
  int nmu, spectral_type, source_geometry;
  int zonly, dust, soa, comoving_source, recoil, core_skip, core_skip_global, use_stokes, use_reduced_wgt;
  int save_Jin, save_Jabs, save_Jmu, save_peeloff, save_peeloff_2D, save_peeloff_3D, save_direc0, save_all_photons;
  int warp_agg;
  int max_events;       // lart_config::max_events (0 = unbounded)
  unsigned int *err;    // sticky ERR_* bits
  int flags_serial_vz;  // ablation: per-lane rejection loops instead of the warp-cooperative sampler
  int local_steps;      // scatter stage resolves 1-cell peel rays / 1-cell flights itself (0: ablation)
  int nobs;
  const DevObserver *obs;
  // dust scattering matrix
  int nPDF;
  const double *sm_coss, *sm_S11, *sm_S12, *sm_S33, *sm_S34, *sm_pdf;
  const int *sm_alias;
  // outputs
  double *tally;
  TallyLayout lay;
  double *allph[10];  // rp0, rp, xfreq1, xfreq2, nscatt_gas, nscatt_dust, I, Q, U, V (NULL = off)
  long long nphotons;
};
enum { A_RP0 = 0, A_RP, A_XFREQ1, A_XFREQ2, A_NSG, A_NSD, A_I, A_Q, A_U, A_V };

// ---------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), key = seed, counter = (photon id, block).
// One block gives two 64-bit words; each word maps to the reference's open
// interval (0,1): ((w>>12)+0.5)*2^-52 — random_mt.f90:628-629.
// ---------------------------------------------------------------------------
LART_DEV void philox4x32_10_inl(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    unsigned long long p0 = (unsigned long long)0xD2511F53u * c0;  // hi|lo in one IMAD.WIDE.U32
    unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
#ifdef LART_PHILOX_NOINLINE
// One copy of the ten rounds per kernel instead of one per draw site (the scatter stage has ~10 sites, ~100 SASS
// instructions each): arguments and result travel in registers (by value), so the call costs a few moves.
__device__ __noinline__ uint4 philox4x32_10_call(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  philox4x32_10_inl(c0, c1, c2, c3, k0, k1);
  return make_uint4(c0, c1, c2, c3);
}
LART_DEV void philox4x32_10(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, uint32_t k0, uint32_t k1) {
  const uint4 r = philox4x32_10_call(c0, c1, c2, c3, k0, k1);
  c0 = r.x; c1 = r.y; c2 = r.z; c3 = r.w;
}
#else
LART_DEV void philox4x32_10(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, uint32_t k0, uint32_t k1) {
  philox4x32_10_inl(c0, c1, c2, c3, k0, k1);
}
#endif

// per-thread work counters of ONE kernel launch (32 bits are plenty; they are summed into FP64 totals)
typedef unsigned int ctr_t;
struct Counters {
  ctr_t scatter = 0, cellsteps = 0, peel = 0, rng = 0, reject = 0, photons = 0, peel_bound = 0;
};

// Stream layout (shared with the CPU oracle): every call consumes ONE Philox block of
// the photon's stream — uniform() uses its first 64-bit word, uniform2() both.  There
// is no buffered half-block, so lanes never diverge on "do I need a new block?" and
// the only generator state that travels with a photon is the block counter.
struct Rng {
  unsigned long long seed, stream, nblk;
  ctr_t nrng;  // uniforms drawn through this object (work counter)
  bool gauss_stored;
  double gset;
  LART_DEV void start(unsigned long long seed_, unsigned long long id, unsigned long long nblk_ = 0) {
    seed = seed_; stream = id; nblk = nblk_; gauss_stored = false; gset = 0.0; nrng = 0;
  }
  LART_DEV void block(unsigned long long &w0, unsigned long long &w1) {
    uint32_t c0 = (uint32_t)stream, c1 = (uint32_t)(stream >> 32), c2 = (uint32_t)nblk, c3 = (uint32_t)(nblk >> 32);
    philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
    w0 = (unsigned long long)c0 | ((unsigned long long)c1 << 32);
    w1 = (unsigned long long)c2 | ((unsigned long long)c3 << 32);
    ++nblk;
  }
  // ((w>>12)+0.5)*2^-52 without an integer->double conversion: the 52 bits become the mantissa of a
  // double in [1,2); subtracting 1-2^-53 is exact and yields (2m+1)*2^-53, the same value bit for bit.
  static LART_DEV double open01(unsigned long long w) {
    return __longlong_as_double((long long)(0x3FF0000000000000ULL | (w >> 12))) - (1.0 - 1.0 / 9007199254740992.0);
  }
  LART_DEV double uniform() {
    unsigned long long w0, w1;
    block(w0, w1);
    nrng += 1;
    return open01(w0);
  }
  LART_DEV void uniform2(double &a, double &b) {
    unsigned long long w0, w1;
    block(w0, w1);
    nrng += 2;
    a = open01(w0); b = open01(w1);
  }
  // rand_gauss1 — random_mt.f90:964-988 (Marsaglia polar; the spare deviate is kept
  // per photon stream, never shared between photons)
  LART_DEV double gauss(ctr_t &nreject) {
    if (gauss_stored) { gauss_stored = false; return gset; }
    double v1, v2, rsq;
    for (;;) {
      uniform2(v1, v2);
      v1 = 2.0 * v1 - 1.0;
      v2 = 2.0 * v2 - 1.0;
      rsq = v1 * v1 + v2 * v2;
      ++nreject;
      if (rsq > 0.0 && rsq < 1.0) break;
    }
    rsq = sqrt(-2.0 * log(rsq) / rsq);
    gset = v1 * rsq;
    gauss_stored = true;
    return v2 * rsq;
  }
};

// ---------------------------------------------------------------------------
// voigt_seon2 — voigt_mod.f90:541-733.  tab[k*4 + {0,1,2,3}] = h0,h1,h2,h3 at node k
// (0-based; h0/h2 are zero beyond node 100 and unused there).
// ---------------------------------------------------------------------------
LART_DEV double voigt_seon2(const double *__restrict__ tab, double vin, double a) {
  const double one_sqrtPI = 0.56418958354775628695;
  double v = (vin < 0.0) ? -vin : vin;
  if (v < 1.0) {  // :691-700
    double vv = DMUL(v, 20.0);
    int k1 = (int)vv;  // 0-based node
    const double *t1 = tab + 4 * k1, *t2 = t1 + 4;
    double y1 = DADD(t1[0], DMUL(a, DADD(t1[1], DMUL(a, t1[2]))));
    double y2 = DADD(t2[0], DMUL(a, DADD(t2[1], DMUL(a, t2[2]))));
    return DADD(y1, DMUL(DSUB(y2, y1), DSUB(vv, (double)k1)));
  } else if (v < 5.0) {  // :701-717
    double vv = DMUL(v, 20.0);
    int k1 = (int)vv - 1;  // 0-based node of the first of three
    const double *t1 = tab + 4 * k1, *t2 = t1 + 4, *t3 = t1 + 8;
    double y1 = DADD(t1[0], DMUL(a, DADD(t1[1], DMUL(a, DADD(t1[2], DMUL(a, t1[3]))))));
    double y2 = DADD(t2[0], DMUL(a, DADD(t2[1], DMUL(a, DADD(t2[2], DMUL(a, t2[3]))))));
    double y3 = DADD(t3[0], DMUL(a, DADD(t3[1], DMUL(a, DADD(t3[2], DMUL(a, t3[3]))))));
    double u1 = (double)k1, u2 = (double)(k1 + 1), u3 = (double)(k1 + 2);
    double d1 = DSUB(vv, u1), d2 = DSUB(vv, u2), d3 = DSUB(vv, u3);
    double A = DMUL(DMUL(DMUL(0.5, y1), d2), d3);
    double B = DMUL(DMUL(y2, d1), d3);
    double Cc = DMUL(DMUL(DMUL(0.5, y3), d1), d2);
    return DADD(DSUB(A, B), Cc);
  } else if (v < 10.0) {  // :718-726
    double vv = DMUL(v, 20.0);
    int k1 = (int)vv;
    const double *t1 = tab + 4 * k1, *t2 = t1 + 4;
    double a2 = DMUL(a, a);
    double y1 = DMUL(a, DADD(t1[1], DMUL(a2, t1[3])));
    double y2 = DMUL(a, DADD(t2[1], DMUL(a2, t2[3])));
    return DADD(y1, DMUL(DSUB(y2, y1), DSUB(vv, (double)k1)));
  }
  double v2 = 1.0 / DMUL(v, v);  // :727-731
  double inner = DADD(DSUB(1.5, DMUL(a, a)), DMUL(3.75, v2));
  return DMUL(DMUL(DMUL(one_sqrtPI, a), v2), DADD(1.0, DMUL(inner, v2)));
}

// ---------------------------------------------------------------------------
// grid access (0-based linear index; x fastest as in grid_type)
// ---------------------------------------------------------------------------
struct CellData {
  double rhokap, voigt_a, Dfreq, vfx, vfy, vfz, rhokapD;
};
LART_DEV size_t cell_index(const DevParams &P, int ic, int jc, int kc) {  // 1-based in
  return (size_t)(ic - 1) + (size_t)P.nx * ((size_t)(jc - 1) + (size_t)P.ny * (size_t)(kc - 1));
}
// Packed cell records live in a two-level tiled order: 32x32x32 super-bricks (2 MB, one TLB page) laid out
// x-fastest, and inside a super-brick the 15-bit Morton code of the local index.  A ray along ANY axis then
// stays inside one page for up to 32 steps and inside one 4 KB neighbourhood for 4; in plain Fortran order a
// walk along z jumps nx*ny*64 B (2.6 MB at 201^3) per step — measured 5.7x slower than a walk along x.
LART_DEV unsigned part1by2_5(unsigned v) {  // spread the low 5 bits: ...edcba -> e00d00c00b00a
  v = (v | (v << 8)) & 0x100Fu;
  v = (v | (v << 4)) & 0x10C3u;
  v = (v | (v << 2)) & 0x1249u;
  return v;
}
template <bool PLAIN = false>
LART_DEV size_t cell_slot(const DevParams &P, int ic, int jc, int kc) {  // 1-based in
  if (!PLAIN && P.amr.on) return (size_t)(ic - 1);  // leaf-indexed records
  const unsigned i = (unsigned)(ic - 1), j = (unsigned)(jc - 1), k = (unsigned)(kc - 1);
  const size_t sb = (size_t)(i >> 5) + (size_t)P.nsbx * ((size_t)(j >> 5) + (size_t)P.nsby * (size_t)(k >> 5));
  const unsigned m = part1by2_5(i & 31u) | (part1by2_5(j & 31u) << 1) | (part1by2_5(k & 31u) << 2);
  return (sb << 15) + m;
}
template <bool PLAIN = false>
LART_DEV void load_cell(const DevParams &P, int ic, int jc, int kc, CellData &o) {
  if (!P.soa) {
    const double2 *q = reinterpret_cast<const double2 *>(P.cells + cell_slot<PLAIN>(P, ic, jc, kc));
    double2 a = __ldg(q), b = __ldg(q + 1), d = __ldg(q + 2), e = __ldg(q + 3);
    o.rhokap = a.x; o.voigt_a = a.y; o.Dfreq = b.x; o.vfx = b.y; o.vfy = d.x; o.vfz = d.y; o.rhokapD = e.x;
  } else {
    const size_t c = cell_index(P, ic, jc, kc);
    o.rhokap = __ldg(P.rhokap + c); o.voigt_a = __ldg(P.voigt_a + c); o.Dfreq = __ldg(P.Dfreq + c);
    o.vfx = __ldg(P.vfx + c); o.vfy = __ldg(P.vfy + c); o.vfz = __ldg(P.vfz + c);
    o.rhokapD = P.dust ? __ldg(P.rhokapD + c) : 0.0;
  }
}
LART_DEV double4 ldg_d4(const void *p) {  // four doubles through the read-only path (two 16-byte loads)
  const double2 a = __ldg(reinterpret_cast<const double2 *>(p)), b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}
LART_DEV double vdotk(const CellData &c, double kx, double ky, double kz) {
  return DADD(DADD(DMUL(c.vfx, kx), DMUL(c.vfy, ky)), DMUL(c.vfz, kz));
}

// ---------------------------------------------------------------------------
// DDA — setup_traversal_car (raytrace_car.f90:12-87) and the step shared by
// raytrace_to_edge_car (:410-508, _zonly :1138-1234) and raytrace_to_tau_car
// (:1425-1648, _zonly :2519-2675).
// ---------------------------------------------------------------------------
struct Ray {
  double x0, y0, z0, kx, ky, kz;    // start point and direction (fixed)
  double tx, ty, tz, delx, dely, delz;
  double d, tau, xfreq, u1;
  CellData cell;                    // record of the current cell (ic,jc,kc)
  int ic, jc, kc;                   // 1-based
  int istep, jstep, kstep;
  int nsteps;
  int flip;                         // xyz symmetry: bit a set = direction component a reflected since the start;
                                    // bit 3 (kRayOpen): this walk ignores the periodic / z-only / mirror binding
  double vshear;                    // shearing box: the photon's vfy_shear (touched only when P.x.shear)
};
constexpr int kRayOpen = 8;

// One straight-line path for both signs of k (no divergent branch around the divide): the special cases of
// raytrace_car.f90:27-44 — already outside (k>0), sitting on the lower face (k<0) — become predicates.
// -d/k == d/|k| bit for bit; the increment itself is computed lazily by ray_advance.
LART_DEV bool axis_setup(double k, double p, int &cell, int n, const double *face, double d, int &step, double &t,
                         double &del, bool eq_test) {
  const bool pos = k > 0.0, neg = k < 0.0;
  const double flo = __ldg(face + cell - 1);
  bool leave = pos && cell > n && (eq_test ? (flo == p) : (flo <= p));
  const bool onface = neg && flo == p;
  leave = leave || (onface && cell <= 1);
  cell -= (onface && cell > 1) ? 1 : 0;
  const int fi = pos ? min(cell, n) : cell - 1;  // the face ahead (clamped: the reference reads past the end there)
  const double f = __ldg(face + fi);
  const double tt = DSUB(f, p) / k;
  const bool moving = pos || neg;
  t = moving ? tt : kHugest;
  del = moving ? -1.0 : kHugest;  // d/|k|, filled in at the first crossing of this axis
  step = pos ? 1 : (neg ? -1 : 0);
  return leave;
}

// The same for the folded and periodic grids: a mirror plane at the lower face reflects the ray into cell c0
// (raytrace_car.f90:1696-1700), a periodic axis wraps the start point to the far side (:1030-1033, :2293-2297; the
// periodic to_edge variant still returns for a photon sitting on the upper face, :1020).  `moved` = k or p changed.
enum { BC_OPEN = 0, BC_MIRROR = 1, BC_PERIODIC = 2 };
LART_DEV bool axis_setup_bc(double &k, double &p, int &cell, int n, const double *face, int &step, double &t, double &del,
                            bool is_tau, int bc, int c0, bool &moved) {
  moved = false;
  step = 0; t = kHugest; del = kHugest;
  if (k > 0.0) {
    if (cell > n) {
      const double f = __ldg(face + cell - 1);
      if (is_tau ? (f == p) : (f <= p)) {
        if (bc == BC_PERIODIC && is_tau) { cell = 1; p = __ldg(face); moved = true; }
        else return true;
      }
    }
    step = 1;
  } else if (k < 0.0) {
    step = -1;
    if (__ldg(face + cell - 1) == p) {
      if (cell > 1) cell -= 1;
      else if (bc == BC_MIRROR) { cell = c0; step = 1; k = -k; moved = true; }
      else if (bc == BC_PERIODIC) { cell = n; p = __ldg(face + n); moved = true; }
      else return true;
    }
  } else {
    return false;
  }
  t = DSUB(__ldg(face + (step == 1 ? min(cell, n) : cell - 1)), p) / k;
  del = -1.0;
  return false;
}

// line-of-sight bulk velocity of the ray's current cell; the shearing box adds the photon's vfy_shear to vfy in the
// to_tau walk (raytrace_car.f90:2809, 2905, 2932)
// PLAIN (template parameter of the walker functions below): the caller guarantees an open Cartesian box (3-D or z-only)
// without octree, mirror / periodic axes, atmosphere, shear or path-length accumulators — the binding of the three headline
// configurations.  The run-time switches of the other bindings are then compile-time false and their code is not
// instantiated (k_wf_trace / k_wf_peel are launched in this variant when is_plain(P)).
template <bool PLAIN = false>
LART_DEV double ray_ulos(const DevParams &P, const Ray &r) {
  if (!PLAIN && P.x.shear && !(r.flip & kRayOpen))
    return DADD(DADD(DMUL(r.cell.vfx, r.kx), DMUL(DADD(r.cell.vfy, r.vshear), r.ky)), DMUL(r.cell.vfz, r.kz));
  return vdotk(r.cell, r.kx, r.ky, r.kz);
}
// `here` (optional) = record of the start cell when the caller already holds it; it is used
// unless the on-face rule moved the start into a neighbouring cell.
// ---------------------------------------------------------------------------
// Octree AMR ray tracers — raytrace_amr.f90:77-351 over octree_mod.f90.  A Ray on the octree keeps its RUNNING position in
// (tx,ty,tz) (the reference advances x,y,z at every face: x = x + t_exit*kx), the current leaf in `ic` and the leaf's
// physics in `cell`; the DDA members are unused.  Every expression is written in the reference's order with explicit
// roundings, so optical depths, positions, frequencies and leaf sequences are those of the CPU oracle bit for bit.
// Open boundaries only (periodic / mirror octrees stay with the Fortran host).
// ---------------------------------------------------------------------------
// amr_find_leaf — octree_mod.f90:149-171
LART_DEV int amr_find_leaf(const DevParams &P, double x, double y, double z) {
  if (x < P.xmin || x > P.xmax || y < P.ymin || y > P.ymax || z < P.zmin || z > P.zmax) return 0;
  int icell = 1;
  for (;;) {
    const AmrCell *c = P.amr.cell + (icell - 1);
    const int il = __ldg(&c->ileaf);
    if (il > 0) return il;
    int ioct = 0;
    if (x >= __ldg(&c->cx)) ioct += 1;
    if (y >= __ldg(&c->cy)) ioct += 2;
    if (z >= __ldg(&c->cz)) ioct += 4;
    icell = __ldg(&c->child[ioct]);
    if (icell == 0) return 0;
  }
}
LART_DEV bool amr_ray_setup(const DevParams &P, Ray &r, const CellData *here) {
  r.tx = r.x0; r.ty = r.y0; r.tz = r.z0;
  r.delx = r.dely = r.delz = 0.0;
  r.istep = r.jstep = r.kstep = 0;
  r.jc = 1; r.kc = 1;
  if (r.ic <= 0) {  // raytrace_amr.f90:98-104
    r.ic = amr_find_leaf(P, r.x0, r.y0, r.z0);
    if (r.ic <= 0) return true;
    here = nullptr;
  }
  if (here) r.cell = *here;
  else load_cell(P, r.ic, 1, 1, r.cell);
  r.u1 = vdotk(r.cell, r.kx, r.ky, r.kz);
  return false;
}
// amr_cell_exit — octree_mod.f90:412-458; face 1=+x 2=-x 3=+y 4=-y 5=+z 6=-z, minloc = the first minimum
LART_DEV void amr_cell_exit(const double4 g, double x, double y, double z, double kx, double ky, double kz, double &t_exit, int &iface) {
  double t1 = kHugest, t2 = kHugest, t3 = kHugest, t4 = kHugest, t5 = kHugest, t6 = kHugest;
  if (kx > 0.0) t1 = DSUB(DADD(g.x, g.w), x) / kx; else if (kx < 0.0) t2 = DSUB(DSUB(g.x, g.w), x) / kx;
  if (ky > 0.0) t3 = DSUB(DADD(g.y, g.w), y) / ky; else if (ky < 0.0) t4 = DSUB(DSUB(g.y, g.w), y) / ky;
  if (kz > 0.0) t5 = DSUB(DADD(g.z, g.w), z) / kz; else if (kz < 0.0) t6 = DSUB(DSUB(g.z, g.w), z) / kz;
  iface = 1; t_exit = t1;
  if (t2 < t_exit) { t_exit = t2; iface = 2; }
  if (t3 < t_exit) { t_exit = t3; iface = 3; }
  if (t4 < t_exit) { t_exit = t4; iface = 4; }
  if (t5 < t_exit) { t_exit = t5; iface = 5; }
  if (t6 < t_exit) { t_exit = t6; iface = 6; }
}
// amr_next_leaf — octree_mod.f90:717-757: the neighbour of the leaf's cell across `iface`, then the descent into its
// children; the octant bit along the face normal is set from iface, the two others from the position on the face
LART_DEV int amr_next_leaf(const DevParams &P, int ineigh, int iface, double x, double y, double z) {
  if (ineigh == 0) return 0;
  for (;;) {
    const AmrCell *c = P.amr.cell + (ineigh - 1);
    const int il = __ldg(&c->ileaf);
    if (il != 0) return il;
    const bool bx = x >= __ldg(&c->cx), by = y >= __ldg(&c->cy), bz = z >= __ldg(&c->cz);
    int ioct;
    switch (iface) {
      case 1: ioct = (by ? 2 : 0) + (bz ? 4 : 0); break;
      case 2: ioct = 1 + (by ? 2 : 0) + (bz ? 4 : 0); break;
      case 3: ioct = (bx ? 1 : 0) + (bz ? 4 : 0); break;
      case 4: ioct = (bx ? 1 : 0) + 2 + (bz ? 4 : 0); break;
      case 5: ioct = (bx ? 1 : 0) + (by ? 2 : 0); break;
      default: ioct = (bx ? 1 : 0) + (by ? 2 : 0) + 4; break;
    }
    const int child = __ldg(&c->child[ioct]);
    if (child == 0) return 0;  // a gap: no leaf there (ileaf of an internal cell is 0), the ray leaves
    ineigh = child;
  }
}
// leave the current leaf: running position to the face, next leaf, frequency into its frame (:162-216 / :302-348).
// false = the ray left the grid
LART_DEV bool amr_cross(const DevParams &P, Ray &r, const AmrGeo *g, double t_exit, int iface) {
  r.tx = DADD(r.tx, DMUL(t_exit, r.kx));
  r.ty = DADD(r.ty, DMUL(t_exit, r.ky));
  r.tz = DADD(r.tz, DMUL(t_exit, r.kz));
  const int il_new = amr_next_leaf(P, __ldg(&g->nb[iface - 1]), iface, r.tx, r.ty, r.tz);
  if (il_new <= 0) return false;
  const double Dold = r.cell.Dfreq;
  load_cell(P, il_new, 1, 1, r.cell);
  const double u2 = vdotk(r.cell, r.kx, r.ky, r.kz);
  r.xfreq = DSUB(DMUL(DADD(r.xfreq, r.u1), Dold) / r.cell.Dfreq, u2);
  r.u1 = u2;
  r.ic = il_new;
  return true;
}
LART_DEV double amr_opacity(const DevParams &P, const double *vtab, const Ray &r) {
  double k = DMUL(r.cell.rhokap, voigt_seon2(vtab, r.xfreq, r.cell.voigt_a));
  if (P.dust) k = DADD(k, r.cell.rhokapD);
  return k;
}
// one leaf of raytrace_to_edge_amr (:302-348); true when the walk is finished
LART_DEV bool amr_edge_step(const DevParams &P, const double *vtab, Ray &r) {
  const AmrGeo *g = P.amr.geo + (r.ic - 1);
  const double4 gc = ldg_d4(g);
  double t_exit;
  int iface;
  amr_cell_exit(gc, r.tx, r.ty, r.tz, r.kx, r.ky, r.kz, t_exit, iface);
  const double kap = amr_opacity(P, vtab, r);
  r.tau = DADD(r.tau, DMUL(t_exit, kap));
  ++r.nsteps;
  if (r.tau >= kTauHuge) return true;
  return !amr_cross(P, r, g, t_exit, iface);
}
// one leaf of raytrace_to_tau_amr (:107-216): 0 = keep walking, 1 = reached tau_in (position in xp,yp,zp), 2 = left the grid
LART_DEV int amr_tau_step(const DevParams &P, const double *vtab, Ray &r, double tau_in, double &xp, double &yp, double &zp) {
  const AmrGeo *g = P.amr.geo + (r.ic - 1);
  const double4 gc = ldg_d4(g);
  double t_exit;
  int iface;
  amr_cell_exit(gc, r.tx, r.ty, r.tz, r.kx, r.ky, r.kz, t_exit, iface);
  const double kap = amr_opacity(P, vtab, r);
  ++r.nsteps;
  const double tnext = DADD(r.tau, DMUL(t_exit, kap));
  if (tnext >= tau_in) {
    const double d_step = (kap > 0.0) ? DSUB(tau_in, r.tau) / kap : t_exit;
    xp = DADD(r.tx, DMUL(d_step, r.kx));
    yp = DADD(r.ty, DMUL(d_step, r.ky));
    zp = DADD(r.tz, DMUL(d_step, r.kz));
    return 1;
  }
  r.tau = tnext;
  return amr_cross(P, r, g, t_exit, iface) ? 0 : 2;
}

template <bool PLAIN = false>
LART_DEV bool ray_setup(const DevParams &P, Ray &r, double x, double y, double z, double kx, double ky, double kz,
                        int ic, int jc, int kc, double xfreq, bool zonly_eq, const CellData *here = nullptr) {
  r.x0 = x; r.y0 = y; r.z0 = z; r.kx = kx; r.ky = ky; r.kz = kz;
  r.ic = ic; r.jc = jc; r.kc = kc;
  r.d = 0.0; r.tau = 0.0; r.xfreq = xfreq; r.nsteps = 0; r.flip = 0;
  if (!PLAIN && P.amr.on) return amr_ray_setup(P, r, here);
  // shearing boxes and atmospheres: raytrace_to_edge is the plain open-box routine whatever raytrace_to_tau is bound to
  const bool open = !PLAIN && P.x.edge_open && !zonly_eq;
  if (open) r.flip = kRayOpen;
  if (!PLAIN && P.bcxy && !open) {  // folded / periodic grids; their to_tau variants test `== xp` (:1681, :1993, :2293), to_edge `<= xp`
    bool mx, my, mz;
    if (axis_setup_bc(r.kx, r.x0, r.ic, P.nx, P.xface, r.istep, r.tx, r.delx, zonly_eq, P.bcxy, P.i0, mx)) return true;
    if (axis_setup_bc(r.ky, r.y0, r.jc, P.ny, P.yface, r.jstep, r.ty, r.dely, zonly_eq, P.bcxy, P.j0, my)) return true;
    if (axis_setup_bc(r.kz, r.z0, r.kc, P.nz, P.zface, r.kstep, r.tz, r.delz, zonly_eq, P.bcz, P.k0, mz)) return true;
    r.flip = (mx ? 1 : 0) | (my ? 2 : 0) | (mz ? 4 : 0);
    load_cell(P, r.ic, r.jc, r.kc, r.cell);
    r.u1 = ray_ulos<PLAIN>(P, r);
    return false;
  }
  if (P.zonly && !open) {
    r.istep = r.jstep = 0; r.tx = r.ty = kHugest; r.delx = r.dely = kHugest;
    if (axis_setup(kz, z, r.kc, P.nz, P.zface, P.dz, r.kstep, r.tz, r.delz, zonly_eq)) return true;
  } else {
    if (axis_setup(kx, x, r.ic, P.nx, P.xface, P.dx, r.istep, r.tx, r.delx, false)) return true;
    if (axis_setup(ky, y, r.jc, P.ny, P.yface, P.dy, r.jstep, r.ty, r.dely, false)) return true;
    if (axis_setup(kz, z, r.kc, P.nz, P.zface, P.dz, r.kstep, r.tz, r.delz, false)) return true;
  }
  if (here && r.ic == ic && r.jc == jc && r.kc == kc) r.cell = *here;
  else load_cell<PLAIN>(P, r.ic, r.jc, r.kc, r.cell);
  r.u1 = vdotk(r.cell, kx, ky, kz);
  return false;
}

constexpr int kFlipShift = 28, kCellMask = (1 << kFlipShift) - 1;
// Resume a suspended walk: start point and direction are the ray's own, the DDA state is what was saved.
template <bool PLAIN = false>
LART_DEV void ray_resume(const DevParams &P, Ray &r, double x, double y, double z, double kx, double ky, double kz,
                         double tx, double ty, double tz, double delx, double dely, double delz, double d, double tau,
                         double xfreq, double u1, int ic, int jc, int kc_flip, bool open = false) {
  const int kc = kc_flip & kCellMask;
  int flip = kc_flip >> kFlipShift;  // see ray_save_state
  if (PLAIN) {
    flip = 0;
  } else if (open) {            // an edge walk of a shearing box / an atmosphere: plain open box
    flip = kRayOpen;
  } else if (P.bcxy == BC_PERIODIC) {  // the start point had been wrapped to the far side
    if (flip & 1) x = __ldg(P.xface + (kx > 0.0 ? 0 : P.nx));
    if (flip & 2) y = __ldg(P.yface + (ky > 0.0 ? 0 : P.ny));
  } else {                      // reflected direction components
    if (flip & 1) kx = -kx;
    if (flip & 2) ky = -ky;
    if (flip & 4) kz = -kz;
  }
  r.flip = flip;
  r.x0 = x; r.y0 = y; r.z0 = z; r.kx = kx; r.ky = ky; r.kz = kz;
  r.tx = tx; r.ty = ty; r.tz = tz; r.delx = delx; r.dely = dely; r.delz = delz;
  r.d = d; r.tau = tau; r.xfreq = xfreq; r.u1 = u1;
  r.ic = ic; r.jc = jc; r.kc = kc; r.nsteps = 0;
  const bool zo = P.zonly && !open;
  r.istep = zo ? 0 : (kx > 0.0 ? 1 : (kx < 0.0 ? -1 : 0));
  r.jstep = zo ? 0 : (ky > 0.0 ? 1 : (ky < 0.0 ? -1 : 0));
  r.kstep = kz > 0.0 ? 1 : (kz < 0.0 ? -1 : 0);
  load_cell<PLAIN>(P, ic, jc, kc, r.cell);
}

// opacity of the current cell at the ray's current frequency (:1488-1494)
LART_DEV double ray_opacity(const DevParams &P, const double *vtab, const Ray &r) {
  double k = DMUL(r.cell.rhokap, voigt_seon2(vtab, r.xfreq, r.cell.voigt_a));
  if (P.dust) k = DADD(k, r.cell.rhokapD);
  return k;
}

// minloc([tx,ty,tz]) — first minimum wins (:476,:1506); z-only walks z
template <bool PLAIN = false>
LART_DEV int ray_axis(const DevParams &P, const Ray &r) {
  if (P.zonly && (PLAIN || !(r.flip & kRayOpen))) return 3;
  if (r.tx <= r.ty && r.tx <= r.tz) return 1;
  if (r.ty <= r.tz) return 2;
  return 3;
}

// Move the index along `axis`; false when the ray left the grid.  On success loads
// the new cell and shifts the frequency into its frame (:1586-1589 / :485-494).
LART_DEV bool ray_leave_or_turn(int bc, int &cell, int &step, double &k, int n, int c0, int &flip, int bit) {
  if (bc == BC_PERIODIC) { cell = cell < 1 ? n : 1; return false; }                                // :2383-2385
  if (bc == BC_MIRROR && cell < 1) { cell = c0; step = 1; k = -k; flip ^= bit; return false; }  // :1791-1795
  cell -= step;
  return true;
}
template <bool PLAIN = false>
LART_DEV bool ray_advance(const DevParams &P, Ray &r, int axis) {
  const bool open = PLAIN || (r.flip & kRayOpen) != 0;
  const int bcxy = open ? (int)BC_OPEN : P.bcxy, bcz = open ? (int)BC_OPEN : P.bcz;
  if (axis == 1) {
    r.ic += r.istep;
    if (r.ic < 1 || r.ic > P.nx) {
      if (!PLAIN && P.x.shear && !open) r.vshear = (r.ic < 1) ? DSUB(r.vshear, P.x.Omega) : DADD(r.vshear, P.x.Omega);  // :2842-2850
      if (ray_leave_or_turn(bcxy, r.ic, r.istep, r.kx, P.nx, P.i0, r.flip, 1)) return false;
    }
    if (r.delx < 0.0) r.delx = P.dx / fabs(r.kx);
    r.tx = DADD(r.tx, r.delx);
  } else if (axis == 2) {
    r.jc += r.jstep;
    if ((r.jc < 1 || r.jc > P.ny) && ray_leave_or_turn(bcxy, r.jc, r.jstep, r.ky, P.ny, P.j0, r.flip, 2)) return false;
    if (r.dely < 0.0) r.dely = P.dy / fabs(r.ky);
    r.ty = DADD(r.ty, r.dely);
  } else {
    r.kc += r.kstep;
    if ((r.kc < 1 || r.kc > P.nz) && ray_leave_or_turn(bcz, r.kc, r.kstep, r.kz, P.nz, P.k0, r.flip, 4)) return false;
    if (r.delz < 0.0) r.delz = P.dz / fabs(r.kz);
    r.tz = DADD(r.tz, r.delz);
  }
  return true;
}
// End point of a walk of length r.d on a folded or periodic grid.  Mirror planes: the reference advances along the
// ORIGINAL direction and mirrors the point back (raytrace_car.f90:1934-1940, :2236-2240); periodic: it is folded back
// into the box (:2506-2508).
LART_DEV void ray_endpoint_bc(const DevParams &P, const Ray &r, double &xp, double &yp, double &zp) {
  if (P.bcxy == BC_MIRROR) {
    xp = DADD(r.x0, DMUL(r.d, (r.flip & 1) ? -r.kx : r.kx));
    yp = DADD(r.y0, DMUL(r.d, (r.flip & 2) ? -r.ky : r.ky));
    zp = DADD(r.z0, DMUL(r.d, (r.flip & 4) ? -r.kz : r.kz));
    if (xp < P.xmin) xp = -xp;
    if (yp < P.ymin) yp = -yp;
    if (P.bcz == BC_MIRROR && zp < P.zmin) zp = -zp;
  } else {
    xp = DADD(r.x0, DMUL(r.d, r.kx));
    yp = DADD(r.y0, DMUL(r.d, r.ky));
    zp = DADD(r.z0, DMUL(r.d, r.kz));
  }
}
LART_DEV void fold_periodic(const DevParams &P, double &xp, double &yp) {
  const double xrange = DSUB(P.xmax, P.xmin), yrange = DSUB(P.ymax, P.ymin);
  xp = DSUB(xp, DMUL(floor(DSUB(xp, P.xmin) / xrange), xrange));
  yp = DSUB(yp, DMUL(floor(DSUB(yp, P.ymin) / yrange), yrange));
}
template <bool PLAIN = false>
LART_DEV void ray_shift(const DevParams &P, Ray &r) {
  double Dold = r.cell.Dfreq;
  load_cell<PLAIN>(P, r.ic, r.jc, r.kc, r.cell);
  double u2 = ray_ulos<PLAIN>(P, r);
  r.xfreq = DSUB(DMUL(DADD(r.xfreq, r.u1), Dold) / r.cell.Dfreq, u2);
  r.u1 = u2;
}

// spherical atmosphere: grid%mask == -1 marks the planet's molecular layer (grid_mod_car.f90:320-330)
LART_DEV bool ray_masked(const DevParams &P, const Ray &r) {
  return P.x.atm == 2 && P.x.mask[cell_index(P, r.ic, r.jc, r.kc)] == -1;
}

// CALCJ / CALCP / CALCPnew: the bin of cell (ic,jc,kc) in the P arrays by par%geometry_JPa — 3 the cell, 2 (ind_cyl, k),
// 1 ind_sph, -1 k — or -1 when the cell takes no deposit (raytrace_car.f90:3989-3991, 3996; the reference does not
// range-check ind_sph: a bin outside 1..nr is skipped instead of written out of bounds)
LART_DEV long long jp_bin(const DevParams &P, int ic, int jc, int kc, double rhokap) {
  if (!(ic > 0 && ic <= P.nx && jc > 0 && jc <= P.ny && kc > 0 && kc <= P.nz) || !(rhokap > 0.0)) return -1;
  switch (P.x.geometry_JPa) {
    case 3: return (long long)cell_index(P, ic, jc, kc);
    case 2: {
      const int ir = __ldg(P.x.ind_cyl + (ic - 1) + (size_t)P.nx * (jc - 1));
      return (ir >= 1 && ir <= P.x.nr) ? (long long)(ir - 1) + (long long)P.x.nr * (kc - 1) : -1;
    }
    case 1: {
      const int ir = __ldg(P.x.ind_sph + cell_index(P, ic, jc, kc));
      return (ir >= 1 && ir <= P.x.nr) ? (long long)(ir - 1) : -1;
    }
    default: return kc - 1;
  }
}
// One atomic per distinct address among the lanes that arrive together (north-star part 5 for the path-length tallies):
// early in a run every photon sits in the source's cell — and in a few frequency bins — so 2.4 M plain atomics per wave
// would serialise on a handful of L2 lines (measured on the tau0 = 1e7 sphere: 3.4e8 scatterings/s with plain atomics).
LART_DEV void atomic_add_grouped(double *addr, double v) {
  const unsigned m = __match_any_sync(__activemask(), (unsigned long long)addr);
  double sum = 0.0;
  for (unsigned mm = m; mm; mm &= mm - 1) sum += __shfl_sync(m, v, __ffs(mm) - 1);  // every lane of the group, same order
  if ((int)(threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(addr, sum);
}
// add_to_J + add_to_Pnew (raytrace_car.f90:3979-4045) for one path segment of a to_tau walk: del = its length,
// dtauH = its line optical depth, xfreq = the photon's frequency in the cell's frame
__device__ __noinline__ void jp_deposit(const DevParams &P, int ic, int jc, int kc, double rhokap, double Dfreq, double xfreq,
                                        double wgt, double del, double dtauH) {
  const long long b = jp_bin(P, ic, jc, kc, rhokap);
  if (b < 0) return;
  if (P.x.calc_J) {
    const double xref = DMUL(xfreq, Dfreq / P.Dfreq_ref);
    const int ix = (int)floor(DSUB(xref, P.xfreq_min) / P.dxfreq) + 1;
    if (ix > 0 && ix <= P.nxfreq) atomic_add_grouped(P.tally + P.lay.J + (ix - 1) + (long long)P.nxfreq * b, DMUL(del, wgt));
  }
  if (P.x.calc_Pnew) atomic_add_grouped(P.tally + P.lay.Pnew + b, DMUL(dtauH, wgt) / (DMUL(rhokap, Dfreq) / P.x.cross0));
}
// add_to_Pa (scattering_car.f90:829-860): one resonance scattering in cell (ic,jc,kc)
__device__ __noinline__ void jp_add_Pa(const DevParams &P, int ic, int jc, int kc, double rhokap, double Dfreq, double wgt) {
  const long long b = jp_bin(P, ic, jc, kc, rhokap);
  if (b >= 0) atomic_add_grouped(P.tally + P.lay.Pa + b, wgt / (DMUL(rhokap, Dfreq) / P.x.cross0));
}

// One cell step of raytrace_to_edge.  Returns true when the walk is finished.
template <bool PLAIN = false>
LART_DEV bool edge_step(const DevParams &P, const double *vtab, Ray &r) {
  if (!PLAIN && P.amr.on) return amr_edge_step(P, vtab, r);
  if (!PLAIN && P.x.atm && ray_masked(P, r)) { r.tau = __longlong_as_double(0x7ff0000000000000LL); return true; }  // :3729-3733 tau = +inf
  double kap = ray_opacity(P, vtab, r);
  ++r.nsteps;
  int ax = ray_axis<PLAIN>(P, r);
  double tn = (ax == 1) ? r.tx : (ax == 2) ? r.ty : r.tz;
  r.tau = DADD(r.tau, DMUL(DSUB(tn, r.d), kap));
  r.d = tn;
  if (!ray_advance<PLAIN>(P, r, ax)) return true;
  if (r.tau >= kTauHuge) return true;
  ray_shift<PLAIN>(P, r);
  return false;
}

// One cell step of raytrace_to_tau.  status: 0 = keep walking, 1 = reached tau_in
// (position in xp,yp,zp), 2 = left the grid, 3 = destroyed by the atmosphere mask (:3186-3190).
// wgt = the photon's weight (only the CALCJ / CALCPnew deposits read it).
template <bool PLAIN = false>
LART_DEV int tau_step(const DevParams &P, const double *vtab, Ray &r, double tau_in, double &xp, double &yp, double &zp,
                      double wgt = 0.0) {
  if (!PLAIN && P.amr.on) return amr_tau_step(P, vtab, r, tau_in, xp, yp, zp);
  if (!PLAIN && P.x.any) {  // the uncommon bindings: mask, path-length accumulators (one extra Voigt evaluation is not paid: kapH below)
    if (P.x.atm && ray_masked(P, r)) return 3;
    if (P.x.jp) {
      const double kapH = DMUL(r.cell.rhokap, voigt_seon2(vtab, r.xfreq, r.cell.voigt_a));
      const double kap = P.dust ? DADD(kapH, r.cell.rhokapD) : kapH;
      ++r.nsteps;
      const int ax = ray_axis(P, r);
      const double tn = (ax == 1) ? r.tx : (ax == 2) ? r.ty : r.tz;
      double del = DSUB(tn, r.d);
      r.tau = DADD(r.tau, DMUL(del, kap));
      r.d = tn;
      const bool reached = r.tau >= tau_in;
      if (reached && kap > 0.0) {
        const double over = DSUB(r.tau, tau_in) / kap;
        r.d = DSUB(r.d, over);
        del = DSUB(del, over);
      }
      // every segment is credited once to the cell it lies in (:1579-1584 in the loop, :1604-1609 for the last one)
      jp_deposit(P, r.ic, r.jc, r.kc, r.cell.rhokap, r.cell.Dfreq, r.xfreq, wgt, del, DMUL(del, kapH));
      if (reached) {
        ray_endpoint_bc(P, r, xp, yp, zp);
        if (P.bcxy == BC_PERIODIC) fold_periodic(P, xp, yp);
        return 1;
      }
      if (!ray_advance(P, r, ax)) return 2;
      ray_shift(P, r);
      return 0;
    }
  }
  double kap = ray_opacity(P, vtab, r);
  ++r.nsteps;
  int ax = ray_axis<PLAIN>(P, r);
  double tn = (ax == 1) ? r.tx : (ax == 2) ? r.ty : r.tz;
  r.tau = DADD(r.tau, DMUL(DSUB(tn, r.d), kap));
  r.d = tn;
  if (r.tau >= tau_in) {  // :1513-1524
    if (kap > 0.0) r.d = DSUB(r.d, DSUB(r.tau, tau_in) / kap);
    if (!PLAIN && P.bcxy) {
      ray_endpoint_bc(P, r, xp, yp, zp);
      if (P.bcxy == BC_PERIODIC) fold_periodic(P, xp, yp);
      return 1;
    }
    xp = DADD(r.x0, DMUL(r.d, r.kx));
    yp = DADD(r.y0, DMUL(r.d, r.ky));
    zp = DADD(r.z0, DMUL(r.d, r.kz));
    return 1;
  }
  if (!ray_advance<PLAIN>(P, r, ax)) return 2;
  ray_shift<PLAIN>(P, r);
  return 0;
}

// ---------------------------------------------------------------------------
// photon_type — define.f90:80-111 (I == 1 always; E1,E2,E3 are line constants)
// ---------------------------------------------------------------------------
enum { PH_ALIVE = 1, PH_FIRST = 2, PH_GAUSS = 4, PH_SCATTER = 8, PH_TAUPEND = 16, PH_DUSTEV = 32, PH_INFLIGHT = 64,
       PH_ABS2 = 128 /* ended in the atmosphere's molecular zone: its weight goes to Jabs2, not Jout */ };
struct Photon {
  double vshear;  // photon%vfy_shear (define.f90:100), shearing boxes only
  long long id;
  double x, y, z, kx, ky, kz, mx, my, mz, nx, ny, nz;
  double xfreq, xfreq_ref, wgt, Q, U, V, nsg, nsd;
  int ic, jc, kc;
  int flags;
};

// ---------------------------------------------------------------------------
// random variates
// ---------------------------------------------------------------------------
// rand_resonance_vz_seon — random_mt.f90:2562-2696
LART_DEV double rand_resonance_vz(Rng &r, double x0in, double a, ctr_t &nrej) {
  const double xc = 1.0 + 1.4142135623730951;
  const double two_over_PI = 2.0 / kPi;
  double x0 = fabs(x0in), vz;
  if (x0 <= 1.0) {  // :2579-2585
    for (;;) {
      double u1, u2;
      r.uniform2(u1, u2);
      vz = x0 + a * tan(kPi * (u1 - 0.5));
      ++nrej;
      if (u2 <= exp(-vz * vz)) break;
    }
    return (x0in < 0.0) ? -vz : vz;
  }
  double x0sq = x0 * x0, api = a * kPi;
  double beta0 = exp(-x0sq / 2.0);
  double h0_two = beta0 / a, h0 = h0_two / 2.0;
  // region table: up to three pieces of the piecewise-constant majorant in beta
  //   piece 0: beta = beta0*sqrt(xi),        C = beta/a
  //   piece 1: beta = lo1 + w1*xi,            C = c1
  //   piece 2: beta = lo2 + w2*xi,            C = c2
  double p0 = 0.0, p01 = 0.0;  // cumulative selection thresholds
  double lo1 = 0.0, w1 = 0.0, c1 = 0.0, lo2 = 0.0, w2 = 0.0, c2 = 0.0;
  int mode;  // 0: single piece (lo1,w1,c1); 1: pieces 0+1; 2: pieces 0+1+2
  double h2 = 0.3861 / (x0sq - 1.373);
  if (x0 < xc || !(h0 < h2)) {  // :2605-2635 and :2667-2690
    double dbeta = sqrt(two_over_PI * a * (1.0 - beta0) * beta0 * x0);
    double beta1 = beta0 + dbeta, one_b1 = 1.0 - beta1;
    double pb1 = sqrt(-2.0 * log(beta1));
    double h1 = two_over_PI * beta1 * pb1 / (x0sq - pb1 * pb1);
    double hm = (x0 < xc) ? h1 : ((h1 > h2) ? h1 : h2);
    double S0 = beta0 * h0, S1 = dbeta * h0, S2 = one_b1 * hm, Stot = S0 + S1 + S2;
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
      if (ua < p0) { beta = beta0 * sqrt(ub); Cb = beta / a; }
      else if (mode == 1 || ua < p01) { beta = lo1 + w1 * ub; Cb = c1; }
      else { beta = lo2 + w2 * ub; Cb = c2; }
      uacc = r.uniform();
    }
    double pb = sqrt(-2.0 * log(beta));
    double t2 = atan((pb - x0) / a);
    t1 = atan((-pb - x0) / a);
    delt = t2 - t1;
    ++nrej;
    if (uacc * Cb < (beta / api) * delt) break;
  }
  vz = x0 + a * tan(delt * r.uniform() + t1);  // :2693
  return (x0in < 0.0) ? -vz : vz;
}

// ---------------------------------------------------------------------------
// Warp-cooperative rand_resonance_vz (north-star part 3: vote/shuffle primitives for the
// rejection loops).  The serial sampler leaves most lanes idle: photons split between the
// |x| <= 1 proposal and the wing majorant, and every trial loop runs until its slowest lane
// accepts.  Here the warp first serves all |x| <= 1 photons, then all wing photons; in each
// phase ALL 32 lanes evaluate trials: with np photons pending every photon gets 32/np
// consecutive trials of its own Philox stream evaluated in parallel, and the FIRST accepted
// trial in stream order wins.  Trial t of a photon consumes the same Philox blocks as in the
// serial loop and the block counter advances exactly past the accepted trial, so the result
// — and every later draw of the photon — is identical to the serial algorithm's.
// ---------------------------------------------------------------------------
struct VzWarpShared {
  double x0[32], a[32];
  unsigned long long id[32], nb[32];
  double tab[32][9];  // wing majorant: beta0, p0, p01, lo1, w1, c1, lo2, w2, c2
  int mode[32];
  double r0[32], r1[32];
  int first[32];
};

LART_DEV void philox_uniform2(unsigned long long seed, unsigned long long id, unsigned long long blk, double &u1, double &u2) {
  uint32_t c0 = (uint32_t)id, c1 = (uint32_t)(id >> 32), c2 = (uint32_t)blk, c3 = (uint32_t)(blk >> 32);
  philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
  u1 = Rng::open01((unsigned long long)c0 | ((unsigned long long)c1 << 32));
  u2 = Rng::open01((unsigned long long)c2 | ((unsigned long long)c3 << 32));
}

// Must be called by all 32 lanes of a converged warp; `mine` = this lane has a photon to sample.
LART_DEV double rand_resonance_vz_warp(VzWarpShared &sh, bool mine, Rng &r, double x0in, double a, ctr_t &nrej) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const double x0 = fabs(x0in);
  const bool core = x0 <= 1.0;
  double vz = 0.0;
  // ---------------- phase A: |x| <= 1, Lorentzian proposal + exp(-u^2) acceptance (:2579-2585)
  {
    bool todo = mine && core;
    unsigned pend = __ballot_sync(FULL, todo);
    while (pend) {
      const int np = __popc(pend), rank = __popc(pend & lt);
      if (todo) { sh.x0[rank] = x0; sh.a[rank] = a; sh.id[rank] = r.stream; sh.nb[rank] = r.nblk; }
      __syncwarp();
      const int tpj = 32 / np, job = lane / tpj, t = lane - job * tpj;
      const bool work = job < np;
      const unsigned grp = (tpj == 32) ? FULL : (((1u << tpj) - 1u) << (job * tpj));
      bool acc = false;
      double v = 0.0;
      if (work) {
        double u1, u2;
        philox_uniform2(r.seed, sh.id[job], sh.nb[job] + t, u1, u2);
        v = sh.x0[job] + sh.a[job] * tan(kPi * (u1 - 0.5));
        acc = u2 <= exp(-v * v);
      }
      const unsigned accm = __ballot_sync(FULL, acc);
      if (acc && (accm & grp & lt) == 0u) { sh.r0[job] = v; sh.first[job] = t; }
      __syncwarp();
      if (todo) {
        const unsigned my = (tpj == 32) ? FULL : (((1u << tpj) - 1u) << (rank * tpj));
        int used = tpj;
        if (accm & my) { vz = sh.r0[rank]; used = sh.first[rank] + 1; todo = false; }
        r.nblk += used; r.nrng += 2 * used; nrej += used;
      }
      __syncwarp();
      pend = __ballot_sync(FULL, todo);
    }
  }
  // ---------------- phase B: wings, piecewise-constant majorant in beta (:2605-2690)
  {
    bool todo = mine && !core;
    unsigned pend = __ballot_sync(FULL, todo);
    double t1 = 0.0, delt = 0.0;
    const bool wing = todo;
    // the majorant table of my photon (same expressions as rand_resonance_vz)
    double T[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int mode = 0;
    if (todo) {
      const double xc = 1.0 + 1.4142135623730951, two_over_PI = 2.0 / kPi;
      double x0sq = x0 * x0;
      double beta0 = exp(-x0sq / 2.0);
      double h0_two = beta0 / a, h0 = h0_two / 2.0;
      double h2 = 0.3861 / (x0sq - 1.373);
      T[0] = beta0;
      if (x0 < xc || !(h0 < h2)) {
        double dbeta = sqrt(two_over_PI * a * (1.0 - beta0) * beta0 * x0);
        double beta1 = beta0 + dbeta, one_b1 = 1.0 - beta1;
        double pb1 = sqrt(-2.0 * log(beta1));
        double h1 = two_over_PI * beta1 * pb1 / (x0sq - pb1 * pb1);
        double hm = (x0 < xc) ? h1 : ((h1 > h2) ? h1 : h2);
        double S0 = beta0 * h0, S1 = dbeta * h0, S2 = one_b1 * hm, Stot = S0 + S1 + S2;
        mode = 2; T[1] = S0 / Stot; T[2] = 1.0 - S2 / Stot;
        T[3] = beta0; T[4] = dbeta; T[5] = h0; T[6] = beta1; T[7] = one_b1; T[8] = hm;
      } else if (h0_two < h2) {
        mode = 0; T[3] = 0.0; T[4] = 1.0; T[5] = h2;
      } else {
        double S0 = beta0 * h0, one_b0 = 1.0 - beta0, S1 = one_b0 * h2, Stot = S0 + S1;
        mode = 1; T[1] = S0 / Stot; T[3] = beta0; T[4] = one_b0; T[5] = h2;
      }
    }
    while (pend) {
      const int np = __popc(pend), rank = __popc(pend & lt);
      if (todo) {
        sh.x0[rank] = x0; sh.a[rank] = a; sh.id[rank] = r.stream; sh.nb[rank] = r.nblk; sh.mode[rank] = mode;
#pragma unroll
        for (int q = 0; q < 9; ++q) sh.tab[rank][q] = T[q];
      }
      __syncwarp();
      const int tpj = 32 / np, job = lane / tpj, t = lane - job * tpj;
      const bool work = job < np;
      const unsigned grp = (tpj == 32) ? FULL : (((1u << tpj) - 1u) << (job * tpj));
      bool acc = false;
      double tt1 = 0.0, td = 0.0;
      if (work) {
        const double jx0 = sh.x0[job], ja = sh.a[job];
        const double *J = sh.tab[job];
        const int jm = sh.mode[job];
        const unsigned long long jid = sh.id[job], jb = sh.nb[job] + (unsigned long long)t * (jm == 0 ? 1 : 2);
        double ua, ub, uacc, beta, Cb;
        philox_uniform2(r.seed, jid, jb, ua, ub);
        if (jm == 0) { beta = J[3] + J[4] * ua; Cb = J[5]; uacc = ub; }
        else {
          if (ua < J[1]) { beta = J[0] * sqrt(ub); Cb = beta / ja; }
          else if (jm == 1 || ua < J[2]) { beta = J[3] + J[4] * ub; Cb = J[5]; }
          else { beta = J[6] + J[7] * ub; Cb = J[8]; }
          double dummy;
          philox_uniform2(r.seed, jid, jb + 1, uacc, dummy);
        }
        double pb = sqrt(-2.0 * log(beta));
        double t2 = atan((pb - jx0) / ja);
        tt1 = atan((-pb - jx0) / ja);
        td = t2 - tt1;
        acc = uacc * Cb < (beta / (ja * kPi)) * td;
      }
      const unsigned accm = __ballot_sync(FULL, acc);
      if (acc && (accm & grp & lt) == 0u) { sh.r0[job] = tt1; sh.r1[job] = td; sh.first[job] = t; }
      __syncwarp();
      if (todo) {
        const unsigned my = (tpj == 32) ? FULL : (((1u << tpj) - 1u) << (rank * tpj));
        int used = tpj;
        if (accm & my) { t1 = sh.r0[rank]; delt = sh.r1[rank]; used = sh.first[rank] + 1; todo = false; }
        r.nblk += (unsigned long long)used * (mode == 0 ? 1 : 2);
        r.nrng += used * (mode == 0 ? 2 : 3);
        nrej += used;
      }
      __syncwarp();
      pend = __ballot_sync(FULL, todo);
    }
    if (wing) vz = x0 + a * tan(delt * r.uniform() + t1);  // :2693
  }
  return (x0in < 0.0) ? -vz : vz;
}

// rand_resonance — random_mt.f90:2974-2993
LART_DEV double rand_resonance(Rng &r, double E1) {
  if (E1 > 0.0) {
    double p2 = sqrt((4.0 - E1) / (3.0 * E1));
    double Q = (4.0 * r.uniform() - 2.0) / (E1 * (p2 * p2 * p2));
    double W = cbrt(Q + sqrt(Q * Q + 1.0));
    return p2 * (W - 1.0 / W);
  } else if (E1 < 0.0) {
    double p2 = sqrt(fabs((4.0 - E1) / (3.0 * E1)));
    double Q = (4.0 * r.uniform() - 2.0) / (E1 * (p2 * p2 * p2));
    return 2.0 * p2 * cos((acos(Q) + kFourPi) / 3.0);
  }
  return 2.0 * r.uniform() - 1.0;
}

// the E1 > 0 branch with the two per-run constants p2 and 1/(E1 p2^3) supplied by the host
LART_DEV double rand_resonance_fast(Rng &r, const DevParams &P) {
  if (!(P.E1 > 0.0)) return rand_resonance(r, P.E1);
  double Q = (4.0 * r.uniform() - 2.0) * P.rr_inv;
  double W = cbrt(Q + sqrt(Q * Q + 1.0));
  return P.rr_p2 * (W - 1.0 / W);
}

// rand_henyey_greenstein — random_mt.f90:3022-3042
LART_DEV double rand_hg(Rng &r, double g) {
  double x = r.uniform();
  if (g == 0.0) return 2.0 * x - 1.0;
  double g2 = g * g, twog = 2.0 * g;
  double q = (1.0 - g2) / (1.0 - g + twog * x);
  return ((1.0 + g2) - q * q) / twog;
}

// rand_voigt — random_mt.f90:3062-3083
LART_DEV double rand_voigt(Rng &r, double a, ctr_t &nrej) {
  double c = tan(kPi * r.uniform() - kHalfPi);
  return a * c + r.gauss(nrej) * (1.0 / 1.4142135623730951);
}

// rand_alias_linear64 — random_mt.f90:2196-2216
LART_DEV double rand_alias_linear(Rng &r, const DevParams &P) {
  int n = P.nPDF - 1;
  double uk, ua;
  r.uniform2(uk, ua);
  int k = (int)floor(n * uk) + 1;
  int idx = (ua < P.sm_pdf[k - 1]) ? k : P.sm_alias[k - 1];
  double p0 = P.sm_S11[idx - 1], p1 = P.sm_S11[idx];
  double x0 = P.sm_coss[idx - 1], x1 = P.sm_coss[idx];
  return (sqrt(p0 * p0 + (p1 * p1 - p0 * p0) * r.uniform()) - p0) * (x1 - x0) / (p1 - p0) + x0;
}

// interp_eq — mathlib.f90:71-106
LART_DEV double interp_eq(const double *x, const double *y, int n, double xnew) {
  double dx = x[1] - x[0];
  int i = (int)((xnew - x[0]) / dx + 1.0);
  if (i <= 0) return y[0];
  if (i >= n) return y[n - 1];
  return y[i - 1] + (y[i] - y[i - 1]) * (xnew - x[i - 1]) / dx;
}

// car_xcrit_local — grid_mod_car.f90:1598-1629
LART_DEV void car_xcrit_local(const DevParams &P, int i, int j, int k, double x, double y, double z, double voigt_a,
                              double rhokap, double &xc, double &xc2) {
  if (P.core_skip_global) { xc = P.xcrit; xc2 = P.xcrit2; return; }
  xc = 0.0; xc2 = 0.0;
  if (i < 1 || j < 1 || k < 1) return;
  if (P.amr.on) {  // amr_xcrit_local — octree_mod.f90:248-284
    const double4 g = ldg_d4(P.amr.geo + (i - 1));
    const double dla = fmin(g.w - fabs(x - g.x), fmin(g.w - fabs(y - g.y), g.w - fabs(z - g.z)));
    if (dla <= 0.0) return;
    const double at = voigt_a * rhokap * dla;
    if (at > 1.0) { xc = cbrt(at) / 5.0; xc2 = xc * xc; }
    return;
  }
  double dlx = fmin(x - __ldg(P.xface + i - 1), __ldg(P.xface + i) - x);
  double dly = fmin(y - __ldg(P.yface + j - 1), __ldg(P.yface + j) - y);
  double dlz = fmin(z - __ldg(P.zface + k - 1), __ldg(P.zface + k) - z);
  double dl = fmin(dlx, fmin(dly, dlz));
  if (dl <= 0.0) return;
  double atau = voigt_a * rhokap * dl;
  if (atau > 1.0) { xc = cbrt(atau) / 5.0; xc2 = xc * xc; }
}

// ---------------------------------------------------------------------------
// tallies
// ---------------------------------------------------------------------------
LART_DEV void tally_add(double *p, double v) {
  if (v != 0.0) atomicAdd(p, v);
}
LART_DEV int freq_bin(const DevParams &P, double xref) {  // 1-based; may be out of range
  return (int)floor((xref - P.xfreq_min) / P.dxfreq) + 1;
}
LART_DEV int jmu_bin(const DevParams &P, double kz) {  // add_to_Jmu — raytrace_car.f90:4049-4064
  if (P.sym) kz = fabs(kz);  // :4058
  int imu = (int)floor((kz - P.mu_min) / P.dmu) + 1;
  return imu < 1 ? 1 : (imu > P.nmu ? P.nmu : imu);
}
LART_DEV void tally_Jout(const DevParams &P, double xref, double kz, double w) {
  int ix = freq_bin(P, xref);
  if (ix >= 1 && ix <= P.nxfreq) {
    tally_add(P.tally + P.lay.Jout + ix - 1, w);
    if (P.save_Jmu) tally_add(P.tally + P.lay.Jmu + (ix - 1) + (long long)P.nxfreq * (jmu_bin(P, kz) - 1), w);
  }
}

// A peel-off ray waiting for its optical depth: everything the deposit needs.
enum { PEEL_DIRECT = 0, PEEL_STOKES = 1, PEEL_NOSTOKES = 2 };
struct __align__(16) PeelRay {  // 144 B = nine 16-byte chunks, moved with 128-bit loads/stores
  double x, y, z, kx, ky, kz, xfreq;
  double wa, wb;           // weight factors around exp(-tau): wgt = wa*exp(-tau)*wb
  double sI, sQ, sU, sV;   // detector-frame Stokes per unit weight (Stokes kinds)
  int ic, jc, kc;
  int obs, pix, ixf, kind;  // pix = (ix-1)+nxim*(iy-1); ixf 1-based or 0 when outside the cube
  int pad_[3];
};
static_assert(sizeof(PeelRay) == 144, "PeelRay must be nine 16-byte chunks");
LART_DEV void ray_store(PeelRay *dst, const PeelRay &src) {
  const uint4 *s = reinterpret_cast<const uint4 *>(&src);
  uint4 *d = reinterpret_cast<uint4 *>(dst);
#pragma unroll
  for (int i = 0; i < 9; ++i) d[i] = s[i];
}
LART_DEV void ray_load(PeelRay &dst, const PeelRay *src) {
  const uint4 *s = reinterpret_cast<const uint4 *>(src);
  uint4 *d = reinterpret_cast<uint4 *>(&dst);
#pragma unroll
  for (int i = 0; i < 9; ++i) d[i] = s[i];
}
// A peel ray suspended after its per-wave step budget: descriptor + the exact DDA state, so that the walk
// resumes with bit-identical arithmetic in a later wave.
struct __align__(16) PeelCont {  // 240 B = fifteen 16-byte chunks
  PeelRay pr;
  double tx, ty, tz, delx, dely, delz, d, tau, xfreq, u1;
  int ic, jc, kc, pad_;
};
static_assert(sizeof(PeelCont) == 240, "PeelCont must be fifteen 16-byte chunks");
LART_DEV void ray_save_state(const Ray &r, double *st10, int *c3) {
  st10[0] = r.tx; st10[1] = r.ty; st10[2] = r.tz; st10[3] = r.delx; st10[4] = r.dely; st10[5] = r.delz;
  st10[6] = r.d; st10[7] = r.tau; st10[8] = r.xfreq; st10[9] = r.u1;
  c3[0] = r.ic; c3[1] = r.jc; c3[2] = r.kc | ((r.flip & 7) << kFlipShift);
}
LART_DEV void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// atan2(y, x) for the TAN pixel (peelingoff_rect.f90:356-357).  A distant observer sees the
// whole grid within a few degrees: for x > 0 and |y/x| < 1/16 the Maclaurin series to t^15
// is exact to double rounding (next term t^17/17 < 2e-22 relative); anything else takes atan2.
LART_DEV double pixel_angle(double y, double x) {
  if (x > 0.0 && fabs(y) < 0.0625 * x) {
    double t = y / x, t2 = t * t;
    double p = fma(t2, -1.0 / 15.0, 1.0 / 13.0);
    p = fma(t2, p, -1.0 / 11.0);
    p = fma(t2, p, 1.0 / 9.0);
    p = fma(t2, p, -1.0 / 7.0);
    p = fma(t2, p, 1.0 / 5.0);
    p = fma(t2, p, -1.0 / 3.0);
    return fma(t * t2, p, t);
  }
  return atan2(y, x);
}

// common head of every peeling routine (peelingoff_rect.f90:44-61 / :326-357 / :595-627)
LART_DEV bool peel_geometry(const DevObserver &ob, const Photon &ph, PeelRay &pr, double &r2) {
  double kx = ob.x - ph.x, ky = ob.y - ph.y, kz = ob.z - ph.z;
  r2 = kx * kx + ky * ky + kz * kz;
  double ir = 1.0 / sqrt(r2);  // one reciprocal instead of three divisions
  kx *= ir; ky *= ir; kz *= ir;
  pr.x = ph.x; pr.y = ph.y; pr.z = ph.z; pr.kx = kx; pr.ky = ky; pr.kz = kz;
  pr.ic = ph.ic; pr.jc = ph.jc; pr.kc = ph.kc;
  const double *R = ob.R;
  double ox = R[0] * kx + R[3] * ky + R[6] * kz;
  double oy = R[1] * kx + R[4] * ky + R[7] * kz;
  double oz = R[2] * kx + R[5] * ky + R[8] * kz;
  int ix = (int)floor(pixel_angle(-ox, oz) * kRad2Deg / ob.dxim + ob.nxim / 2.0) + 1;
  int iy = (int)floor(pixel_angle(-oy, oz) * kRad2Deg / ob.dyim + ob.nyim / 2.0) + 1;
  pr.pix = (ix - 1) + ob.nxim * (iy - 1);
  return ix >= 1 && ix <= ob.nxim && iy >= 1 && iy <= ob.nyim;
}

// rotate Stokes to the detector axes (peelingoff_rect.f90:428-445)
LART_DEV void to_detector(const DevObserver &ob, double nx, double ny, double nz, double Qobs, double Uobs, double &Qdet,
                          double &Udet) {
  const double *R = ob.R;
  double cosg = -(R[0] * nx + R[3] * ny + R[6] * nz);
  double sing = R[1] * nx + R[4] * ny + R[7] * nz;
  double cos2g = 2.0 * cosg * cosg - 1.0, sin2g = 2.0 * cosg * sing;
  Qdet = cos2g * Qobs + sin2g * Uobs;
  Udet = -sin2g * Qobs + cos2g * Uobs;
}

// observer azimuth in the photon's (m,n) frame and the new normal (:364-380)
LART_DEV void stokes_azimuth(const Photon &ph, const PeelRay &pr, double cost, double &sint, double &cosp, double &sinp,
                             double &nx, double &ny, double &nz) {
  sint = sqrt(1.0 - cost * cost);
  if (sint == 0.0) { cosp = 1.0; sinp = 0.0; }
  else {
    double is = 1.0 / sint;
    cosp = (pr.kx * ph.mx + pr.ky * ph.my + pr.kz * ph.mz) * is;
    sinp = (pr.kx * ph.nx + pr.ky * ph.ny + pr.kz * ph.nz) * is;
  }
  nx = -sinp * ph.mx + cosp * ph.nx;
  ny = -sinp * ph.my + cosp * ph.ny;
  nz = -sinp * ph.mz + cosp * ph.nz;
}

// peeling_direct_outside — peelingoff_rect.f90:24-129.  cs = photon's cell record.
// in_clump: the photon was born inside a clump and its frequency is in that clump's frame (cs carries the clump's
// bulk velocity): only the binning frequency is shifted (peelingoff_rect.f90:65-69)
LART_DEV bool peel_direct_prepare(const DevParams &P, const DevObserver &ob, int iobs, const Photon &ph, const CellData &cs,
                                  PeelRay &pr, bool in_clump = false) {
  double r2;
  bool in_image = peel_geometry(ob, ph, pr, r2);
  double xref;
  pr.xfreq = ph.xfreq;
  if (!P.comoving_source && !in_clump) {  // :70-80
    double u1 = vdotk(cs, ph.kx, ph.ky, ph.kz);
    xref = ph.xfreq + u1;
    pr.xfreq = xref - vdotk(cs, pr.kx, pr.ky, pr.kz);
  } else {  // :81-87
    xref = ph.xfreq + vdotk(cs, pr.kx, pr.ky, pr.kz);
  }
  xref = xref * (cs.Dfreq / P.Dfreq_ref);
  int ixf = freq_bin(P, xref);
  pr.ixf = (ixf >= 1 && ixf <= P.nxfreq) ? ixf : 0;
  pr.obs = iobs; pr.kind = PEEL_DIRECT;
  pr.wa = 1.0 / (kFourPi * r2) * ph.wgt;  // wgt0
  pr.wb = 1.0;
  pr.sI = pr.sQ = pr.sU = pr.sV = 0.0;
  return in_image;
}

// peeling_resonance_stokes_outside — peelingoff_rect.f90:303-482
// `drop(pr)` is asked as soon as the ray's start, direction and frequency are known: a ray the caller can prove to
// contribute exactly zero (tau cap inside its own cell) skips the Stokes algebra.  Returns 0 = outside the image
// (no ray), 1 = descriptor complete, 2 = dropped.
struct NeverDrop { LART_DEV bool operator()(const PeelRay &) const { return false; } };
template <class Drop>
LART_DEV int peel_resonance_stokes_prepare2(const DevParams &P, const DevObserver &ob, int iobs, const Photon &ph,
                                            const CellData &cs, double xfreq_atom, double ux, double uy, double uz,
                                            PeelRay &pr, Drop &&drop) {
  double r2;
  if (!peel_geometry(ob, ph, pr, r2)) return 0;
  double cost = ph.kx * pr.kx + ph.ky * pr.ky + ph.kz * pr.kz;
  double sint, cosp, sinp, nx, ny, nz;
  stokes_azimuth(ph, pr, cost, sint, cosp, sinp, nx, ny, nz);
  double xfreq = xfreq_atom + (ux * cosp + uy * sinp) * sint + uz * cost;  // :392
  if (P.recoil) xfreq -= (P.g_recoil0 / cs.Dfreq) * (1.0 - cost);
  pr.xfreq = xfreq;
  if (drop(pr)) return 2;
  double cos2p = 1.0, sin2p = 0.0;
  if (sint != 0.0) { cos2p = 2.0 * cosp * cosp - 1.0; sin2p = 2.0 * cosp * sinp; }
  double cost2 = cost * cost;
  double S22 = 0.75 * P.E1 * (cost2 + 1.0), S11 = S22 + P.E2, S12 = 0.75 * P.E1 * (cost2 - 1.0);
  double S33 = 1.5 * P.E1 * cost, S44 = 1.5 * P.E3 * cost;
  double u1 = vdotk(cs, pr.kx, pr.ky, pr.kz);
  double xref = (xfreq + u1) * (cs.Dfreq / P.Dfreq_ref);
  int ixf = freq_bin(P, xref);
  double Q0 = cos2p * ph.Q + sin2p * ph.U, U0 = -sin2p * ph.Q + cos2p * ph.U;
  const double i4pi = 1.0 / kFourPi;
  double Iobs = (S11 + S12 * Q0) * i4pi, Qobs = (S12 + S22 * Q0) * i4pi;
  double Uobs = (S33 * U0) * i4pi, Vobs = (S44 * ph.V) * i4pi;
  pr.ixf = (ixf >= 1 && ixf <= P.nxfreq) ? ixf : 0;
  pr.obs = iobs; pr.kind = PEEL_STOKES;
  pr.wa = 1.0 / r2; pr.wb = ph.wgt;
  pr.sI = Iobs; pr.sV = Vobs;
  to_detector(ob, nx, ny, nz, Qobs, Uobs, pr.sQ, pr.sU);
  return 1;
}
LART_DEV bool peel_resonance_stokes_prepare(const DevParams &P, const DevObserver &ob, int iobs, const Photon &ph,
                                            const CellData &cs, double xfreq_atom, double ux, double uy, double uz,
                                            PeelRay &pr) {
  return peel_resonance_stokes_prepare2(P, ob, iobs, ph, cs, xfreq_atom, ux, uy, uz, pr, NeverDrop()) == 1;
}

// peeling_resonance_nostokes_outside — peelingoff_rect.f90:576-690
template <class Drop>
LART_DEV int peel_resonance_nostokes_prepare2(const DevParams &P, const DevObserver &ob, int iobs, const Photon &ph,
                                              const CellData &cs, double xfreq_atom, double ux, double uy, double uz,
                                              PeelRay &pr, Drop &&drop) {
  double r2;
  if (!peel_geometry(ob, ph, pr, r2)) return 0;
  double cost = ph.kx * pr.kx + ph.ky * pr.ky + ph.kz * pr.kz;
  double cost2 = cost * cost;
  double sint = sqrt(1.0 - cost2);
  double rho1 = sqrt(1.0 - ph.kz * ph.kz) * sint;
  double cosp = 1.0, sinp = 0.0;
  if (rho1 != 0.0) {
    double rho = 1.0 / rho1;
    cosp = rho * (cost * ph.kz - pr.kz);
    sinp = rho * (ph.kx * pr.ky - pr.kx * ph.ky);
  }
  double xfreq = xfreq_atom + (ux * cosp + uy * sinp) * sint + uz * cost;
  if (P.recoil) xfreq -= (P.g_recoil0 / cs.Dfreq) * (1.0 - cost);
  pr.xfreq = xfreq;
  if (drop(pr)) return 2;
  double u1 = vdotk(cs, pr.kx, pr.ky, pr.kz);
  double xref = (xfreq + u1) * (cs.Dfreq / P.Dfreq_ref);
  int ixf = freq_bin(P, xref);
  pr.ixf = (ixf >= 1 && ixf <= P.nxfreq) ? ixf : 0;
  pr.obs = iobs; pr.kind = PEEL_NOSTOKES;
  double peel = 0.75 * P.E1 * (cost2 + 1.0) + P.E2;
  pr.wa = peel / (kFourPi * r2); pr.wb = ph.wgt;
  pr.sI = pr.sQ = pr.sU = pr.sV = 0.0;
  return 1;
}
LART_DEV bool peel_resonance_nostokes_prepare(const DevParams &P, const DevObserver &ob, int iobs, const Photon &ph,
                                              const CellData &cs, double xfreq_atom, double ux, double uy, double uz,
                                              PeelRay &pr) {
  return peel_resonance_nostokes_prepare2(P, ob, iobs, ph, cs, xfreq_atom, ux, uy, uz, pr, NeverDrop()) == 1;
}

// peeling_dust_stokes_outside — peelingoff_rect.f90:131-299
LART_DEV bool peel_dust_stokes_prepare(const DevParams &P, const DevObserver &ob, int iobs, const Photon &ph,
                                       const CellData &cs, PeelRay &pr) {
  double r2;
  if (!peel_geometry(ob, ph, pr, r2)) return false;
  double u1 = vdotk(cs, pr.kx, pr.ky, pr.kz);
  double xref = (ph.xfreq + u1) * (cs.Dfreq / P.Dfreq_ref);
  int ixf = freq_bin(P, xref);
  double cost = ph.kx * pr.kx + ph.ky * pr.ky + ph.kz * pr.kz;
  double sint, cosp, sinp, nx, ny, nz;
  stokes_azimuth(ph, pr, cost, sint, cosp, sinp, nx, ny, nz);
  double cos2p = 1.0, sin2p = 0.0;
  if (sint != 0.0) { cos2p = 2.0 * cosp * cosp - 1.0; sin2p = 2.0 * cosp * sinp; }
  double S11 = interp_eq(P.sm_coss, P.sm_S11, P.nPDF, cost), S12 = interp_eq(P.sm_coss, P.sm_S12, P.nPDF, cost);
  double S33 = interp_eq(P.sm_coss, P.sm_S33, P.nPDF, cost), S34 = interp_eq(P.sm_coss, P.sm_S34, P.nPDF, cost);
  double Q0 = cos2p * ph.Q + sin2p * ph.U, U0 = -sin2p * ph.Q + cos2p * ph.U;
  double Iobs = (S11 + S12 * Q0) / kTwoPi, Qobs = (S12 + S11 * Q0) / kTwoPi;
  double Uobs = (S33 * U0 + S34 * ph.V) / kTwoPi, Vobs = (-S34 * U0 + S33 * ph.V) / kTwoPi;
  pr.xfreq = ph.xfreq;
  pr.ixf = (ixf >= 1 && ixf <= P.nxfreq) ? ixf : 0;
  pr.obs = iobs; pr.kind = PEEL_STOKES;
  pr.wa = 1.0 / r2; pr.wb = ph.wgt;
  pr.sI = Iobs; pr.sV = Vobs;
  to_detector(ob, nx, ny, nz, Qobs, Uobs, pr.sQ, pr.sU);
  return true;
}

// peeling_dust_nostokes_outside — peelingoff_rect.f90:484-574
LART_DEV bool peel_dust_nostokes_prepare(const DevParams &P, const DevObserver &ob, int iobs, const Photon &ph,
                                         const CellData &cs, PeelRay &pr) {
  double r2;
  if (!peel_geometry(ob, ph, pr, r2)) return false;
  double u1 = vdotk(cs, pr.kx, pr.ky, pr.kz);
  double xref = (ph.xfreq + u1) * (cs.Dfreq / P.Dfreq_ref);
  int ixf = freq_bin(P, xref);
  double cosa = ph.kx * pr.kx + ph.ky * pr.ky + ph.kz * pr.kz;
  double hg = P.hgg;
  double base = (1.0 + hg * hg) - 2.0 * hg * cosa;
  double peel = (1.0 - hg * hg) / (base * sqrt(base)) / kFourPi;
  pr.xfreq = ph.xfreq;
  pr.ixf = (ixf >= 1 && ixf <= P.nxfreq) ? ixf : 0;
  pr.obs = iobs; pr.kind = PEEL_NOSTOKES;
  pr.wa = peel / r2; pr.wb = ph.wgt;
  pr.sI = pr.sQ = pr.sU = pr.sV = 0.0;
  return true;
}

// Deposit one finished peel ray (peelingoff_rect.f90:96-126, :451-479, :674-687).
// `lanes` = mask of converged lanes calling together; with warp aggregation lanes
// that hit the same (observer, pixel, frequency bin) are summed in registers first
// and one lane issues the atomics.
LART_DEV void peel_deposit(const DevParams &P, const PeelRay &pr, double tau, unsigned lanes, bool aggregate = true) {
  double e = exp(-tau);
  double v[4] = {0.0, 0.0, 0.0, 0.0};
  int nv;
  if (pr.kind == PEEL_DIRECT) { v[0] = e * pr.wa; v[1] = pr.wa; nv = 2; }
  else if (pr.kind == PEEL_STOKES) {
    double w = pr.wa * e * pr.wb;
    v[0] = w * pr.sI; v[1] = w * pr.sQ; v[2] = w * pr.sU; v[3] = w * pr.sV; nv = 4;
  } else { v[0] = pr.wa * e * pr.wb; nv = 1; }
  // A ray that ran into the tau cap (745.2) carries exp(-tau) == 0: all its contributions are exact
  // zeros, which never change a sum (the reference adds them; the result is the same).
  const bool nonzero = (v[0] != 0.0) | (v[1] != 0.0) | (v[2] != 0.0) | (v[3] != 0.0);
  if (aggregate) lanes = __ballot_sync(lanes, nonzero);
  if (!nonzero) return;
  bool leader = true;
  if (P.warp_agg && aggregate) {
    // collision-free: ixf < 2^24, pix < 2^30, obs < 2^8 (lart_gpu_create checks the first two; LART_MAX_OBSERVERS = 181)
    unsigned long long key = ((unsigned long long)(unsigned)pr.kind << 62) | ((unsigned long long)(unsigned)pr.obs << 54) |
                             ((unsigned long long)(unsigned)pr.pix << 24) | (unsigned long long)(unsigned)pr.ixf;
    unsigned grp = __match_any_sync(lanes, key);
    int lane = threadIdx.x & 31;
    int lead = __ffs(grp) - 1;
    leader = lane == lead;
    if (__all_sync(lanes, grp == 0xffffffffu)) {
      // the whole warp hits one bin: plain butterfly
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
      }
    } else {
      unsigned rest = leader ? (grp & ~(1u << lead)) : 0u;  // lanes whose values the leader still has to add
      int maxpop = __reduce_max_sync(lanes, (unsigned)__popc(rest));
      for (int it = 0; it < maxpop; ++it) {
        bool take = rest != 0u;
        int src = take ? (__ffs(rest) - 1) : lane;
        rest &= rest - 1u;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          double o = __shfl_sync(lanes, v[q], src);
          if (take && q < nv) v[q] += o;
        }
      }
    }
  }
  if (!leader) return;
  double *base = P.tally + P.lay.obs_base + (long long)pr.obs * P.lay.obs_stride;
  const TallyLayout &L = P.lay;
  long long q3 = (long long)(pr.ixf - 1) + (long long)P.nxfreq * pr.pix;
  bool c3 = P.save_peeloff_3D && pr.ixf > 0, c2 = P.save_peeloff_2D;
  if (pr.kind == PEEL_DIRECT) {
    if (c2) {
      tally_add(base + L.img[T_DIREC] + pr.pix, v[0]);
      if (P.use_stokes) tally_add(base + L.img[T_I] + pr.pix, v[0]);
      if (P.save_direc0) tally_add(base + L.img[T_DIREC0] + pr.pix, v[1]);
    }
    if (c3) {
      tally_add(base + L.cube[T_DIREC] + q3, v[0]);
      if (P.use_stokes) tally_add(base + L.cube[T_I] + q3, v[0]);
      if (P.save_direc0) tally_add(base + L.cube[T_DIREC0] + q3, v[1]);
    }
  } else if (pr.kind == PEEL_STOKES) {
    if (c2) {
      tally_add(base + L.img[T_SCATT] + pr.pix, v[0]); tally_add(base + L.img[T_I] + pr.pix, v[0]);
      tally_add(base + L.img[T_Q] + pr.pix, v[1]); tally_add(base + L.img[T_U] + pr.pix, v[2]);
      tally_add(base + L.img[T_V] + pr.pix, v[3]);
    }
    if (c3) {
      tally_add(base + L.cube[T_SCATT] + q3, v[0]); tally_add(base + L.cube[T_I] + q3, v[0]);
      tally_add(base + L.cube[T_Q] + q3, v[1]); tally_add(base + L.cube[T_U] + q3, v[2]);
      tally_add(base + L.cube[T_V] + q3, v[3]);
    }
  } else {
    if (c2) tally_add(base + L.img[T_SCATT] + pr.pix, v[0]);
    if (c3) tally_add(base + L.cube[T_SCATT] + q3, v[0]);
  }
}

// ---------------------------------------------------------------------------
// scattering — scattering_car.f90
// ---------------------------------------------------------------------------
// rotate the (m,n,k) triad — scattering_car.f90:470-484 / :314-328
LART_DEV void rotate_triad(Photon &ph, double cost, double sint, double cosp, double sinp) {
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
LART_DEV void rotate_k(Photon &ph, double cost, double sint, double cosp, double sinp) {
  if (fabs(ph.kz) >= 0.99999999999) {
    ph.kx = sint * cosp; ph.ky = sint * sinp; ph.kz = cost;
  } else {
    double kx1 = ph.kx, ky1 = ph.ky, kz1 = ph.kz;
    double kr = sqrt(kx1 * kx1 + ky1 * ky1);
    ph.kx = cost * kx1 + sint * (kz1 * kx1 * cosp - ky1 * sinp) / kr;
    ph.ky = cost * ky1 + sint * (kz1 * ky1 * cosp + kx1 * sinp) / kr;
    ph.kz = cost * kz1 - sint * cosp * kr;
  }
}
// azimuth by rejection — scattering_car.f90:364-371 / :280-287.  cos(2phi), sin(2phi) come
// from the double-angle identities of (cos phi, sin phi) = sincospi(2 xi): one libm call per trial.
LART_DEV void sample_phi_stokes(Rng &r, const Photon &ph, double S12overS11, ctr_t &nrej, double &cosp,
                                double &sinp) {
  double env = 1.0 + fabs(S12overS11) * sqrt(ph.Q * ph.Q + ph.U * ph.U);
  for (;;) {
    double u1, u2;
    r.uniform2(u1, u2);
    sincospi(2.0 * u1, &sinp, &cosp);
    double c2 = 2.0 * cosp * cosp - 1.0, s2 = 2.0 * sinp * cosp;
    double Prand = env * u2;
    double Pcomp = 1.0 + S12overS11 * (ph.Q * c2 + ph.U * s2);
    ++nrej;
    if (Prand <= Pcomp) break;
  }
}

// ---------------------------------------------------------------------------
// Warp-cooperative versions of the two other rejection loops of a resonance scattering (north-star part 3).
// Serial loops run until the slowest lane of the warp accepts — measured on the tau0 = 1e7 sphere: 6.7 rounds of the
// azimuth loop per warp with 9 of 32 lanes active (a dipole scattering leaves the photon strongly polarised, the
// acceptance of a trial is 1/(1 + |S12/S11| p) ~ 0.5-0.9).  Here every lane first tries the next trial of its own
// photon (no exchange); the np photons still pending then share the warp — 32/np consecutive trials of each photon's
// own Philox stream are evaluated in parallel and the first accepted one in stream order wins.  The block counter
// advances exactly past the accepted trial: values and all later draws are those of the serial loop.
// Must be called by all 32 lanes of a converged warp; `mine` = this lane has a photon to sample.
// ---------------------------------------------------------------------------
// sample_phi_stokes — scattering_car.f90:364-371
LART_DEV void sample_phi_stokes_warp(VzWarpShared &sh, bool mine, Rng &r, double Q, double U, double S12overS11, ctr_t &nrej,
                                     double &cosp, double &sinp) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const double env = 1.0 + fabs(S12overS11) * sqrt(Q * Q + U * U);
  bool todo = mine;
  if (todo) {  // round 1: my own next trial
    double u1, u2;
    r.uniform2(u1, u2);
    sincospi(2.0 * u1, &sinp, &cosp);
    const double c2 = 2.0 * cosp * cosp - 1.0, s2 = 2.0 * sinp * cosp;
    ++nrej;
    todo = !(env * u2 <= 1.0 + S12overS11 * (Q * c2 + U * s2));
  }
  unsigned pend = __ballot_sync(FULL, todo);
  while (pend) {
    const int np = __popc(pend), rank = __popc(pend & lt);
    if (todo) { sh.x0[rank] = Q; sh.a[rank] = U; sh.r0[rank] = S12overS11; sh.r1[rank] = env; sh.id[rank] = r.stream; sh.nb[rank] = r.nblk; }
    __syncwarp();
    const int tpj = 32 / np, job = lane / tpj, t = lane - job * tpj;
    const bool work = job < np;
    const unsigned grp = (tpj == 32) ? FULL : (((1u << tpj) - 1u) << (job * tpj));
    bool acc = false;
    double c = 1.0, sn = 0.0;
    if (work) {
      double u1, u2;
      philox_uniform2(r.seed, sh.id[job], sh.nb[job] + t, u1, u2);
      sincospi(2.0 * u1, &sn, &c);
      const double c2 = 2.0 * c * c - 1.0, s2 = 2.0 * sn * c;
      acc = sh.r1[job] * u2 <= 1.0 + sh.r0[job] * (sh.x0[job] * c2 + sh.a[job] * s2);
    }
    const unsigned accm = __ballot_sync(FULL, acc);
    if (acc && (accm & grp & lt) == 0u) { sh.tab[job][0] = c; sh.tab[job][1] = sn; sh.first[job] = t; }
    __syncwarp();
    if (todo) {
      const unsigned my = (tpj == 32) ? FULL : (((1u << tpj) - 1u) << (rank * tpj));
      int used = tpj;
      if (accm & my) { cosp = sh.tab[rank][0]; sinp = sh.tab[rank][1]; used = sh.first[rank] + 1; todo = false; }
      r.nblk += used; r.nrng += 2 * used; nrej += used;
    }
    __syncwarp();
    pend = __ballot_sync(FULL, todo);
  }
}
// The pair of rand_gauss1 deviates of one Marsaglia polar draw (random_mt.f90:964-988), warp-cooperative:
// g_first is what the drawing call returns (v2*f), g_spare what it leaves behind for the next call (v1*f).
LART_DEV void gauss_pair_warp(VzWarpShared &sh, bool mine, Rng &r, ctr_t &nrej, double &g_first, double &g_spare) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  double v1 = 0.0, v2 = 0.0, rsq = 0.5;
  bool todo = mine;
  if (todo) {
    r.uniform2(v1, v2);
    v1 = 2.0 * v1 - 1.0; v2 = 2.0 * v2 - 1.0;
    rsq = v1 * v1 + v2 * v2;
    ++nrej;
    todo = !(rsq > 0.0 && rsq < 1.0);
  }
  unsigned pend = __ballot_sync(FULL, todo);
  while (pend) {
    const int np = __popc(pend), rank = __popc(pend & lt);
    if (todo) { sh.id[rank] = r.stream; sh.nb[rank] = r.nblk; }
    __syncwarp();
    const int tpj = 32 / np, job = lane / tpj, t = lane - job * tpj;
    const bool work = job < np;
    const unsigned grp = (tpj == 32) ? FULL : (((1u << tpj) - 1u) << (job * tpj));
    bool acc = false;
    double w1 = 0.0, w2 = 0.0, wr = 0.0;
    if (work) {
      philox_uniform2(r.seed, sh.id[job], sh.nb[job] + t, w1, w2);
      w1 = 2.0 * w1 - 1.0; w2 = 2.0 * w2 - 1.0;
      wr = w1 * w1 + w2 * w2;
      acc = wr > 0.0 && wr < 1.0;
    }
    const unsigned accm = __ballot_sync(FULL, acc);
    if (acc && (accm & grp & lt) == 0u) { sh.tab[job][0] = w1; sh.tab[job][1] = w2; sh.tab[job][2] = wr; sh.first[job] = t; }
    __syncwarp();
    if (todo) {
      const unsigned my = (tpj == 32) ? FULL : (((1u << tpj) - 1u) << (rank * tpj));
      int used = tpj;
      if (accm & my) { v1 = sh.tab[rank][0]; v2 = sh.tab[rank][1]; rsq = sh.tab[rank][2]; used = sh.first[rank] + 1; todo = false; }
      r.nblk += used; r.nrng += 2 * used; nrej += used;
    }
    __syncwarp();
    pend = __ballot_sync(FULL, todo);
  }
  const double f = sqrt(-2.0 * log(rsq) / rsq);
  g_spare = v1 * f;
  g_first = v2 * f;
}

// A resonance scattering in two halves, so that each can be its own kernel (lart_engine.cu: k_wf_draw / k_wf_apply):
//
//   draw_resonance_warp   every random variate of the event, in the reference's order of draws — u_par
//                         (rand_resonance_vz), cos(theta) (rand_resonance), the azimuth (rejection, Stokes) and the
//                         perpendicular atom velocity (two rand_gauss, or the core-skip / no-Stokes form) — with the
//                         three rejection loops served warp-cooperatively.  Needs the frequency, the cell's Voigt
//                         parameter, the photon's Q and U and (core-skip only) its position: few registers, so many
//                         warps are resident to hide the latency of the Philox and libm chains.
//   apply_resonance       the deterministic rest of scatter_resonance_stokes (:331-486) / _nostokes (:660-827): new
//                         frequency, recoil, peel-off (`peel` is invoked where the reference calls it, :446 / :788),
//                         Stokes vector and triad.  Straight-line FP64 code.
struct ScatterVariates {
  double uz, cost, cosp, sinp, ux, uy;
};
// Must be called by all 32 lanes of a converged warp; `active` = this lane has a resonance scattering to draw.
// x, a: frequency and Voigt parameter (do_resonance1, line_mod.f90:108-139); Q, U: Stokes parameters (read only with
// STOKES); xc, xc2: core-skip threshold of this photon (0 = no skip), computed by the caller.
template <bool STOKES>
LART_DEV void draw_resonance_warp(VzWarpShared &sh, bool active, const DevParams &P, Rng &r, double x, double a, double Q,
                                  double U, double xc, double xc2, Counters &cnt, ScatterVariates &v) {
  v.uz = rand_resonance_vz_warp(sh, active, r, x, a, cnt.reject);
  double S12overS11 = 0.0;
  v.cost = 0.0; v.cosp = 1.0; v.sinp = 0.0; v.ux = 0.0; v.uy = 0.0;
  if (active) {
    v.cost = rand_resonance_fast(r, P);
    const double cost2 = v.cost * v.cost;
    const double S22 = 0.75 * P.E1 * (cost2 + 1.0);
    S12overS11 = 0.75 * P.E1 * (cost2 - 1.0) / (S22 + P.E2);
  }
  if (STOKES) sample_phi_stokes_warp(sh, active, r, Q, U, S12overS11, cnt.reject, v.cosp, v.sinp);
  else if (active) sincospi(2.0 * r.uniform(), &v.sinp, &v.cosp);
  const bool skip = active && P.core_skip && fabs(x) < xc;
  if (STOKES) {  // :413-414: two rand_gauss calls = the stored spare (if any) and one polar pair
    const bool need = active && !skip;
    double g_first, g_spare;
    gauss_pair_warp(sh, need, r, cnt.reject, g_first, g_spare);
    if (need) {
      const double one_over_sqrt2 = 1.0 / 1.4142135623730951;
      if (r.gauss_stored) { v.ux = r.gset * one_over_sqrt2; v.uy = g_first * one_over_sqrt2; r.gset = g_spare; }
      else { v.ux = g_first * one_over_sqrt2; v.uy = g_spare * one_over_sqrt2; }
    }
  }
  if (active && (!STOKES || skip)) {  // :397-401 / :749-752
    double u1, u2;
    r.uniform2(u1, u2);
    const double uxy = skip ? sqrt(xc2 - log(u2)) : sqrt(-log(u2));
    double s2, c2;
    sincospi(2.0 * u1, &s2, &c2);
    v.ux = uxy * c2; v.uy = uxy * s2;
  }
}
template <bool STOKES, class PeelFn>
LART_DEV void apply_resonance(const DevParams &P, Photon &ph, const CellData &cs, const ScatterVariates &v, PeelFn &&peel) {
  ph.nsg += ph.wgt;
  if (P.x.calc_P) jp_add_Pa(P, ph.ic, ph.jc, ph.kc, cs.rhokap, cs.Dfreq, ph.wgt);  // scattering_car.f90:356-358
  const double cost = v.cost, cosp = v.cosp, sinp = v.sinp;
  const double sint = sqrt(1.0 - cost * cost);
  const double cost2 = cost * cost;
  const double S22 = 0.75 * P.E1 * (cost2 + 1.0), S11 = S22 + P.E2, S12 = 0.75 * P.E1 * (cost2 - 1.0);
  const double S33 = 1.5 * P.E1 * cost, S44 = 1.5 * P.E3 * cost;
  const double xfreq_atom = ph.xfreq - v.uz;  // do_resonance1 — line_mod.f90:108-139
  ph.xfreq = xfreq_atom + v.uz * cost + (v.ux * cosp + v.uy * sinp) * sint;
  if (P.recoil) ph.xfreq -= (P.g_recoil0 / cs.Dfreq) * (1.0 - cost);
  if (P.save_peeloff) peel(xfreq_atom, v.ux, v.uy, v.uz);
  if (STOKES) {
    const double cos2p = 2.0 * cosp * cosp - 1.0, sin2p = 2.0 * sinp * cosp;
    const double Q0 = cos2p * ph.Q + sin2p * ph.U, U0 = -sin2p * ph.Q + cos2p * ph.U;
    const double I1 = S11 + S12 * Q0, Q1 = S12 + S22 * Q0, U1 = S33 * U0, V1 = S44 * ph.V;
    const double iI = 1.0 / I1;
    ph.Q = Q1 * iI; ph.U = U1 * iI; ph.V = V1 * iI;
    rotate_triad(ph, cost, sint, cosp, sinp);
  } else {
    rotate_k(ph, cost, sint, cosp, sinp);
  }
}

// Outcome of one scattering event: what the peel stage needs.
struct ScatterOut {
  int peel_kind;  // -1 none, else PEEL_STOKES / PEEL_NOSTOKES with the fields below (resonance) or dust
  bool dust;
  double xfreq_atom, ux, uy, uz;
};

// scatter_resonance_stokes (:331-486) / _nostokes (:660-827).  The peel-off call
// sits between the frequency update and the Stokes/triad update in the reference
// (:446 / :788): `peel` is invoked at exactly that point.
// `uz` = rand_resonance_vz(x, a) has been drawn by the caller (serially, or warp-cooperatively).
// CLUMP: uz and xfreq_atom come from do_resonance1_clump (line_clump_mod.f90:29-58) and the perpendicular atom
// velocity is rescaled by vth_ratio = cl_Dfreq/cl_Dfreq_ref (scattering_car.f90:384-388, 715-719)
// STOKES: 1 / 0 = par%use_stokes known at compile time (the wavefront scatter stage is instantiated per value),
// -1 = read it from P.
template <bool CLUMP, int STOKES = -1, class PeelFn>
LART_DEV void scatter_resonance_core(const DevParams &P, Photon &ph, Rng &r, const CellData &cs, Counters &cnt, double uz,
                                     double xfreq_atom, double vth_ratio, PeelFn &&peel) {
  const bool stokes = STOKES < 0 ? (P.use_stokes != 0) : (STOKES != 0);
  ph.nsg += ph.wgt;
  if (!CLUMP && P.x.calc_P) jp_add_Pa(P, ph.ic, ph.jc, ph.kc, cs.rhokap, cs.Dfreq, ph.wgt);  // scattering_car.f90:356-358, 691-693
  double cost = rand_resonance_fast(r, P);
  double sint = sqrt(1.0 - cost * cost);
  double cost2 = cost * cost;
  double S22 = 0.75 * P.E1 * (cost2 + 1.0), S11 = S22 + P.E2, S12 = 0.75 * P.E1 * (cost2 - 1.0);
  double S33 = 1.5 * P.E1 * cost, S44 = 1.5 * P.E3 * cost;
  double cosp, sinp;
  if (stokes) sample_phi_stokes(r, ph, S12 / S11, cnt.reject, cosp, sinp);
  else sincospi(2.0 * r.uniform(), &sinp, &cosp);
  double xc = 0.0, xc2 = 0.0;
  if (P.core_skip) car_xcrit_local(P, ph.ic, ph.jc, ph.kc, ph.x, ph.y, ph.z, cs.voigt_a, cs.rhokap, xc, xc2);
  bool skip = P.core_skip && fabs(ph.xfreq) < xc;
  double ux, uy;
  if (stokes && !skip) {  // :413-414
    const double one_over_sqrt2 = 1.0 / 1.4142135623730951;
    ux = r.gauss(cnt.reject) * one_over_sqrt2;
    uy = r.gauss(cnt.reject) * one_over_sqrt2;
  } else {  // :397-401 / :749-752
    double u1, u2;
    r.uniform2(u1, u2);
    double uxy = skip ? sqrt(xc2 - log(u2)) : sqrt(-log(u2));
    double s2, c2;
    sincospi(2.0 * u1, &s2, &c2);
    ux = uxy * c2; uy = uxy * s2;
  }
  if (CLUMP) { ux = ux * vth_ratio; uy = uy * vth_ratio; }
  ph.xfreq = xfreq_atom + uz * cost + (ux * cosp + uy * sinp) * sint;
  if (P.recoil) ph.xfreq -= (P.g_recoil0 / cs.Dfreq) * (1.0 - cost);
  if (P.save_peeloff) peel(xfreq_atom, ux, uy, uz);
  if (stokes) {
    double cos2p = 2.0 * cosp * cosp - 1.0, sin2p = 2.0 * sinp * cosp;
    double Q0 = cos2p * ph.Q + sin2p * ph.U, U0 = -sin2p * ph.Q + cos2p * ph.U;
    double I1 = S11 + S12 * Q0, Q1 = S12 + S22 * Q0, U1 = S33 * U0, V1 = S44 * ph.V;
    double iI = 1.0 / I1;
    ph.Q = Q1 * iI; ph.U = U1 * iI; ph.V = V1 * iI;
    rotate_triad(ph, cost, sint, cosp, sinp);
  } else {
    rotate_k(ph, cost, sint, cosp, sinp);
  }
}

// dust absorption / weight reduction — scattering_car.f90:219-264 / :506-555
LART_DEV bool dust_absorb(const DevParams &P, Photon &ph, Rng &r, const CellData &cs) {
  double xref = (ph.xfreq + vdotk(cs, ph.kx, ph.ky, ph.kz)) * (cs.Dfreq / P.Dfreq_ref);
  int ix = freq_bin(P, xref);
  bool inb = ix >= 1 && ix <= P.nxfreq;
  if (!P.use_reduced_wgt) {
    if (r.uniform() > P.albedo) {
      if (P.save_Jabs && inb) tally_add(P.tally + P.lay.Jabs + ix - 1, ph.wgt);
      ph.flags &= ~PH_ALIVE;
      if (P.save_all_photons) { ph.xfreq_ref = xref; ph.wgt = 0.0; }
      return false;
    }
  } else {
    if (P.save_Jabs && inb) tally_add(P.tally + P.lay.Jabs + ix - 1, ph.wgt * (1.0 - P.albedo));
    ph.wgt = ph.wgt * P.albedo;
  }
  return true;
}

// scatter_dust_stokes (:201-329) / _nostokes (:488-584); `peel` is invoked where
// the reference calls peeling_dust_* (after absorption, before the new direction).
template <class PeelFn>
LART_DEV void scatter_resonance(const DevParams &P, Photon &ph, Rng &r, const CellData &cs, Counters &cnt, double uz,
                                PeelFn &&peel) {
  // do_resonance1 — line_mod.f90:108-139
  scatter_resonance_core<false, -1>(P, ph, r, cs, cnt, uz, ph.xfreq - uz, 1.0, peel);
}
template <class PeelFn>
LART_DEV void scatter_dust(const DevParams &P, Photon &ph, Rng &r, const CellData &cs, Counters &cnt, PeelFn &&peel) {
  ph.nsd += ph.wgt;
  if (!dust_absorb(P, ph, r, cs)) return;
  if (P.save_peeloff) peel();
  if (P.use_stokes) {
    double cost = rand_alias_linear(r, P);
    double sint = sqrt(1.0 - cost * cost);
    double S11 = interp_eq(P.sm_coss, P.sm_S11, P.nPDF, cost), S12 = interp_eq(P.sm_coss, P.sm_S12, P.nPDF, cost);
    double S33 = interp_eq(P.sm_coss, P.sm_S33, P.nPDF, cost), S34 = interp_eq(P.sm_coss, P.sm_S34, P.nPDF, cost);
    double cosp, sinp;
    sample_phi_stokes(r, ph, S12 / S11, cnt.reject, cosp, sinp);
    double cos2p = 2.0 * cosp * cosp - 1.0, sin2p = 2.0 * sinp * cosp;
    double Q0 = cos2p * ph.Q + sin2p * ph.U, U0 = -sin2p * ph.Q + cos2p * ph.U;
    double I1 = S11 + S12 * Q0, Q1 = S12 + S11 * Q0, U1 = S33 * U0 + S34 * ph.V, V1 = -S34 * U0 + S33 * ph.V;
    ph.Q = Q1 / I1; ph.U = U1 / I1; ph.V = V1 / I1;
    rotate_triad(ph, cost, sint, cosp, sinp);
  } else {
    double cost = rand_hg(r, P.hgg);
    double sint = sqrt(1.0 - cost * cost);
    double cosp, sinp;
    sincospi(2.0 * r.uniform(), &sinp, &cosp);
    rotate_k(ph, cost, sint, cosp, sinp);
  }
}

// ---------------------------------------------------------------------------
// generate_photon + setup_isotropic_injection — generate_photon.f90:3-339, 342-408
// (everything up to, not including, the direct peel-off call at :334-336)
// ---------------------------------------------------------------------------
LART_DEV void generate_photon(const DevParams &P, Photon &ph, Rng &r, Counters &cnt, CellData &cs) {
  if (P.source_geometry == 2) {  // uniform_sphere :34-42
    double u1, u2;
    r.uniform2(u1, u2);
    double rp = cbrt(u1) * P.source_rmax;
    double cost = 2.0 * u2 - 1.0, sint = sqrt(1.0 - cost * cost);
    double sp, cp;
    sincospi(2.0 * r.uniform(), &sp, &cp);
    ph.x = rp * sint * cp; ph.y = rp * sint * sp; ph.z = rp * cost;
  } else if (P.source_geometry == 1) {  // uniform :50-54
    double u1, u2;
    r.uniform2(u1, u2);
    ph.x = (P.xmax - P.xmin) * u1 + P.xmin;
    ph.y = (P.ymax - P.ymin) * u2 + P.ymin;
    ph.z = (P.zmax - P.zmin) * r.uniform() + P.zmin;
  } else if (P.source_geometry == 3) {  // plane_illumination — random_plane_illumination :729-760
    if (P.x.atm == 1) {
      ph.x = 0.0; ph.y = 0.0; ph.z = P.zmax;
    } else {
      const double rp = P.rmax * sqrt(r.uniform());
      const double phi = (P.bcxy == BC_MIRROR ? kHalfPi : kTwoPi) * r.uniform();
      ph.x = rp * cos(phi); ph.y = rp * sin(phi); ph.z = P.zmin;
    }
  } else {  // point :126-131
    ph.x = P.xs; ph.y = P.ys; ph.z = P.zs;
  }
  ph.wgt = 1.0;
  ph.vshear = 0.0;  // :141
  if (P.sym) {  // sources are folded into the octant (:356-360)
    if (ph.x < P.xmin) ph.x = -ph.x;
    if (ph.y < P.ymin) ph.y = -ph.y;
    if (ph.z < P.zmin) ph.z = -ph.z;
  }
  double cost, sint, cosp, sinp;
  if (P.source_geometry == 3) {  // :765-778 a parallel beam: down onto the slab, or up along +z
    cost = (P.x.atm == 1) ? -1.0 : 1.0; sint = 0.0; cosp = 1.0; sinp = 0.0;
  } else {
    double uc, up;
    r.uniform2(uc, up);
    cost = 2.0 * uc - 1.0;
    sint = sqrt(1.0 - cost * cost);
    sincospi(2.0 * up, &sinp, &cosp);
  }
  ph.kx = sint * cosp; ph.ky = sint * sinp; ph.kz = cost;
  ph.ic = (int)floor((ph.x - P.xmin) / P.dx) + 1;
  ph.jc = (int)floor((ph.y - P.ymin) / P.dy) + 1;
  ph.kc = (int)floor((ph.z - P.zmin) / P.dz) + 1;
  if (P.amr.on) {  // generate_photon.f90:375-376
    ph.ic = amr_find_leaf(P, ph.x, ph.y, ph.z); ph.jc = 1; ph.kc = 1;
  } else {
    if (ph.kx < 0.0 && ph.ic == P.nx + 1) ph.ic = P.nx;
    if (ph.ky < 0.0 && ph.jc == P.ny + 1) ph.jc = P.ny;
    if (ph.kz < 0.0 && ph.kc == P.nz + 1) ph.kc = P.nz;
    if (ph.kx > 0.0 && ph.ic < 1) ph.ic = 1;
    if (ph.ky > 0.0 && ph.jc < 1) ph.jc = 1;
    if (ph.kz > 0.0 && ph.kc < 1) ph.kc = 1;
  }
  ph.mx = cost * cosp; ph.my = cost * sinp; ph.mz = -sint;
  ph.nx = -sinp; ph.ny = cosp; ph.nz = 0.0;
  ph.Q = 0.0; ph.U = 0.0; ph.V = 0.0;
  ph.nsg = 0.0; ph.nsd = 0.0; ph.xfreq_ref = 0.0;
  ph.flags = PH_ALIVE | PH_FIRST;
  ph.xfreq = P.xfreq0;
  // a source placed on the upper boundary with k>0 keeps index n+1 in the reference and
  // reads out of bounds there; clamp the READ only (the photon leaves at its first trace)
  int ci = min(max(ph.ic, 1), P.nx), cj = min(max(ph.jc, 1), P.ny), ck = min(max(ph.kc, 1), P.nz);
  load_cell(P, ci, cj, ck, cs);
  switch (P.spectral_type) {  // :243-300
    case 3: ph.xfreq = r.uniform() * (P.xfreq_max - P.xfreq_min) + P.xfreq_min; ph.xfreq = ph.xfreq / (cs.Dfreq / P.Dfreq_ref); break;
    case 2: ph.xfreq = ph.xfreq + rand_voigt(r, P.voigt_a0, cnt.reject) * P.Dfreq0 / cs.Dfreq; break;
    case 1: ph.xfreq = ph.xfreq + rand_voigt(r, cs.voigt_a, cnt.reject); break;
    case 4: ph.xfreq = ph.xfreq + r.gauss(cnt.reject) * P.gaussian_sigma_x; ph.xfreq = ph.xfreq / (cs.Dfreq / P.Dfreq_ref); break;
    default: break;
  }
  double u1 = vdotk(cs, ph.kx, ph.ky, ph.kz);
  if (!P.comoving_source) ph.xfreq = ph.xfreq - u1;  // :303-306
  if (P.save_Jin) {  // :309-322
    double xlab = (ph.xfreq + u1) * (cs.Dfreq / P.Dfreq_ref);
    int ix = freq_bin(P, xlab);
    if (ix >= 1 && ix <= P.nxfreq) tally_add(P.tally + P.lay.Jin + ix - 1, ph.wgt);
  }
}

// impact radius of the photon's line about the origin — run_simulation_mod.f90:302-330
LART_DEV double impact_radius(const DevParams &P, const Photon &ph, double &mx, double &my, double &mz) {
  double xp = ph.x, yp = ph.y, zp = ph.z;
  if (P.rmax > 0.0) {
    double rr = ph.x * ph.x + ph.y * ph.y + ph.z * ph.z, dist = 0.0;
    if (rr > P.rmax * P.rmax) {
      double rk = ph.x * ph.kx + ph.y * ph.ky + ph.z * ph.kz;
      double det = rk * rk - (rr - P.rmax * P.rmax);
      dist = (det < 0.0) ? 0.0 : -rk + sqrt(fmax(0.0, det));
    }
    xp = ph.x + dist * ph.kx; yp = ph.y + dist * ph.ky; zp = ph.z + dist * ph.kz;
  }
  double rk = xp * ph.kx + yp * ph.ky + zp * ph.kz;
  mx = xp - rk * ph.kx; my = yp - rk * ph.ky; mz = zp - rk * ph.kz;
  return sqrt(mx * mx + my * my + mz * mz);
}
// make_all_initial_photons / make_all_photons — run_simulation_mod.f90:249-358
LART_DEV void record_initial(const DevParams &P, const Photon &ph) {
  size_t s = (size_t)(ph.id - 1);
  if (ph.id < 1 || ph.id > P.nphotons) return;
  if (P.allph[A_XFREQ1]) P.allph[A_XFREQ1][s] = ph.xfreq;
  if (P.source_geometry != 0 && P.allph[A_RP0]) {
    double mx, my, mz;
    P.allph[A_RP0][s] = impact_radius(P, ph, mx, my, mz);
  }
}
LART_DEV void record_final(const DevParams &P, const Photon &ph) {
  if (ph.id < 1 || ph.id > P.nphotons) return;
  size_t s = (size_t)(ph.id - 1);
  double mx, my, mz;
  double mm = impact_radius(P, ph, mx, my, mz);
  if (P.allph[A_RP]) P.allph[A_RP][s] = mm;
  if (P.allph[A_XFREQ2]) P.allph[A_XFREQ2][s] = ph.xfreq_ref;
  if (P.allph[A_NSG]) P.allph[A_NSG][s] = ph.nsg;
  if (P.allph[A_NSD]) P.allph[A_NSD][s] = ph.nsd;
  if (P.use_stokes) {
    double cos2p = 1.0, sin2p = 0.0;
    if (mm > 0.0) {
      mx /= mm; my /= mm; mz /= mm;
      double cosp = mx * ph.mx + my * ph.my + mz * ph.mz, sinp = mx * ph.nx + my * ph.ny + mz * ph.nz;
      cos2p = 2.0 * cosp * cosp - 1.0; sin2p = 2.0 * sinp * cosp;
    }
    if (P.allph[A_I]) P.allph[A_I][s] = ph.wgt;
    if (P.allph[A_Q]) P.allph[A_Q][s] = (cos2p * ph.Q + sin2p * ph.U) * ph.wgt;
    if (P.allph[A_U]) P.allph[A_U][s] = (-sin2p * ph.Q + cos2p * ph.U) * ph.wgt;
    if (P.allph[A_V]) P.allph[A_V][s] = ph.V * ph.wgt;
  }
}

}  // namespace lart
