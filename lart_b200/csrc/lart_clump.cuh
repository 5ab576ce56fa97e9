// Clump medium: spherical clumps in vacuum inside a sphere, found through a CSR acceleration grid.
// Device restatement of clump_mod.f90:130-190, 1369-1540, 1595-1634 and raytrace_clump.f90:68-270, 494-533
// (non-overlapping populations, line_type 1).  Every geometric expression is written with explicit roundings
// (DMUL/DADD/DSUB, no FMA contraction) in the order the reference evaluates it, so that positions, entry/exit
// distances and optical depths are bit-identical to the CPU restatement the tests check against.
#pragma once
#include "lart_device.cuh"

namespace lart {

constexpr double kTauHugeClump = 745.2;  // raytrace_clump.f90:59
constexpr double kTinyDouble = 2.2250738585072014e-308;

LART_DEV double4 ldg4(const double4 *p) {  // read-only path, two 128-bit loads
  const double2 a = __ldg(reinterpret_cast<const double2 *>(p)), b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}
// clump_exit_dist / sphere_exit_dist — clump_mod.f90:1509-1540
LART_DEV double clump_exit_dist(const DevClumps &C, double xp, double yp, double zp, double kx, double ky, double kz, int icl) {
  const double4 g = ldg4(C.geo + icl - 1);
  const double rx = DSUB(xp, g.x), ry = DSUB(yp, g.y), rz = DSUB(zp, g.z);
  const double b = DADD(DADD(DMUL(rx, kx), DMUL(ry, ky)), DMUL(rz, kz));
  double disc = DADD(DSUB(DMUL(b, b), DADD(DADD(DMUL(rx, rx), DMUL(ry, ry)), DMUL(rz, rz))), g.w);
  if (disc < 0.0) disc = 0.0;
  return fmax(0.0, DADD(-b, sqrt(disc)));
}
LART_DEV double sphere_exit_dist(const DevClumps &C, double xp, double yp, double zp, double kx, double ky, double kz) {
  const double b = DADD(DADD(DMUL(xp, kx), DMUL(yp, ky)), DMUL(zp, kz));
  double disc = DADD(DSUB(DMUL(b, b), DADD(DADD(DMUL(xp, xp), DMUL(yp, yp)), DMUL(zp, zp))), C.R2);
  if (disc < 0.0) disc = 0.0;
  return fmax(0.0, DADD(-b, sqrt(disc)));
}
LART_DEV bool outside_sphere(const DevClumps &C, double xp, double yp, double zp) {
  return DADD(DADD(DMUL(xp, xp), DMUL(yp, yp)), DMUL(zp, zp)) >= C.R2;
}
// voigt_clump / kappa_clump / ulos_clump — clump_mod.f90:130-190
LART_DEV double kappa_clump(const DevParams &P, const double *vtab, const ClumpPhys &cp, double xfreq) {
  const double xloc = DMUL(xfreq, P.cl.Dfreq_ref / cp.Dfreq);
  double kap = DMUL(cp.rhokap, voigt_seon2(vtab, xloc, cp.voigt_a));
  if (P.dust) kap = DADD(kap, cp.rhokapD);
  return kap;
}
LART_DEV double ulos_clump(const DevParams &P, const ClumpPhys &cp, double kx, double ky, double kz) {
  return DMUL(DADD(DADD(DMUL(cp.vx, kx), DMUL(cp.vy, ky)), DMUL(cp.vz, kz)), cp.Dfreq / P.cl.Dfreq_ref);
}
LART_DEV ClumpPhys load_clump(const DevClumps &C, int icl) {
  const double4 *p = reinterpret_cast<const double4 *>(C.phys + icl - 1);
  const double4 a = ldg4(p), b = ldg4(p + 1);
  ClumpPhys cp;
  cp.rhokap = a.x; cp.rhokapD = a.y; cp.voigt_a = a.z; cp.Dfreq = a.w;
  cp.vx = b.x; cp.vy = b.y; cp.vz = b.z; cp.pad_ = 0.0;
  return cp;
}
LART_DEV int cg_clamp(double p, double lo, double inv, int n) { return max(0, min(n - 1, (int)DMUL(DSUB(p, lo), inv))); }

// active_set_at_point, first hit — clump_mod.f90:1595-1634
LART_DEV int clump_at_point(const DevClumps &C, double xp, double yp, double zp) {
  const int ci = cg_clamp(xp, C.xmin, C.inv_dx, C.cgx), cj = cg_clamp(yp, C.ymin, C.inv_dy, C.cgy), ck = cg_clamp(zp, C.zmin, C.inv_dz, C.cgz);
  for (int k = max(0, ck - 1); k <= min(C.cgz - 1, ck + 1); ++k)
    for (int j = max(0, cj - 1); j <= min(C.cgy - 1, cj + 1); ++j)
      for (int i = max(0, ci - 1); i <= min(C.cgx - 1, ci + 1); ++i) {
        const size_t icell = (size_t)i + (size_t)C.cgx * ((size_t)j + (size_t)C.cgy * (size_t)k);
        const int p0 = __ldg(C.cg_start + icell), p1 = __ldg(C.cg_start + icell + 1);
        for (int ip = p0; ip < p1; ++ip) {
          const int icl = __ldg(C.cg_list + ip - 1);
          const double4 g = ldg4(C.geo + icl - 1);
          const double rx = DSUB(xp, g.x), ry = DSUB(yp, g.y), rz = DSUB(zp, g.z);
          if (DADD(DADD(DMUL(rx, rx), DMUL(ry, ry)), DMUL(rz, rz)) <= g.w) return icl;
        }
      }
  return 0;
}

// update_cell_idx — raytrace_clump.f90:68-75
LART_DEV void update_cell_idx(const DevParams &P, Photon &ph) {
  ph.ic = max(1, min(P.nx, (int)floor(DSUB(ph.x, P.xmin) / P.dx) + 1));
  ph.jc = max(1, min(P.ny, (int)floor(DSUB(ph.y, P.ymin) / P.dy) + 1));
  ph.kc = max(1, min(P.nz, (int)floor(DSUB(ph.z, P.zmin) / P.dz) + 1));
}

// ---------------------------------------------------------------------------
// The two ray tracers as one resumable state machine.  A step is either one clump segment, the set-up of a search, or
// ONE cell of the CSR walk of find_next_clump — so that the lanes of a warp, each on its own ray, stay in step and can
// be refilled one by one (k_cl_flight, k_cl_peel).  Arithmetic and its order are those of clump_mod.f90:1393-1506 and
// raytrace_clump.f90:83-270, 494-533; running the machine to completion IS raytrace_to_edge_clump / _to_tau_clump.
// ---------------------------------------------------------------------------
enum { CW_CLUMP = 1, CW_FIND = 2, CW_CELL = 3 };
struct ClumpWalk {
  double x, y, z, kx, ky, kz;  // start of the current segment or search; direction
  double xfreq;                // in the current clump's frame while inside one, else in the box frame
  double tau;                  // edge walk: accumulated depth; tau walk: depth still to go
  double t_sp;                 // distance to the bounding sphere for the current search
  double tx, ty, tz, delx, dely, delz, d, best_te;
  int ci, cj, ck, si, sj, sk;
  int best_icl, skip_icl, icl;  // icl = clump the ray is in (CW_CLUMP)
  int phase, ncells, nclumps;
};
LART_DEV void cw_start(ClumpWalk &w, double x, double y, double z, double kx, double ky, double kz, double xfreq, int icl,
                       double tau) {
  w.x = x; w.y = y; w.z = z; w.kx = kx; w.ky = ky; w.kz = kz; w.xfreq = xfreq; w.tau = tau;
  w.icl = icl; w.skip_icl = 0; w.ncells = 0; w.nclumps = 0;
  w.phase = icl > 0 ? CW_CLUMP : CW_FIND;
  // the DDA increments cg_d/|k| depend on the direction only: one divide per axis and ray, at the first crossing of that
  // axis (the reference recomputes the same quotients in every find_next_clump call; FP64 divides were 1/3 of the walkers'
  // instructions).  -1 = not yet computed.
  w.delx = (kx != 0.0) ? -1.0 : kHugest; w.dely = (ky != 0.0) ? -1.0 : kHugest; w.delz = (kz != 0.0) ? -1.0 : kHugest;
  w.si = kx > 0.0 ? 1 : (kx < 0.0 ? -1 : 0); w.sj = ky > 0.0 ? 1 : (ky < 0.0 ? -1 : 0); w.sk = kz > 0.0 ? 1 : (kz < 0.0 ? -1 : 0);
}
LART_DEV void cw_advance(ClumpWalk &w, double t) {
  w.x = DADD(w.x, DMUL(t, w.kx)); w.y = DADD(w.y, DMUL(t, w.ky)); w.z = DADD(w.z, DMUL(t, w.kz));
}
// CW_FIND: distance to the sphere, DDA set-up.  Returns false when the ray is already at the sphere (t_sp <= 0).
LART_DEV bool cw_find_begin(const DevClumps &C, ClumpWalk &w) {
  w.t_sp = sphere_exit_dist(C, w.x, w.y, w.z, w.kx, w.ky, w.kz);
  if (w.t_sp <= 0.0) return false;
  w.best_te = kHugest; w.best_icl = 0; w.d = 0.0;
  w.ci = cg_clamp(w.x, C.xmin, C.inv_dx, C.cgx); w.cj = cg_clamp(w.y, C.ymin, C.inv_dy, C.cgy); w.ck = cg_clamp(w.z, C.zmin, C.inv_dz, C.cgz);
  auto axis = [](double k, double p, int cc, double lo, double dd, int st) {  // path length to the first face ahead
    return st == 0 ? kHugest : DSUB(DADD(lo, DMUL((double)(cc + (st > 0 ? 1 : 0)), dd)), p) / k;
  };
  w.tx = axis(w.kx, w.x, w.ci, C.xmin, C.dx, w.si);
  w.ty = axis(w.ky, w.y, w.cj, C.ymin, C.dy, w.sj);
  w.tz = axis(w.kz, w.z, w.ck, C.zmin, C.dz, w.sk);
  w.phase = CW_CELL;
  return true;
}
// CW_CELL: one cell of find_next_clump.  Returns true when the search is over (w.best_icl = 0: nothing before t_sp).
LART_DEV bool cw_find_cell(const DevClumps &C, ClumpWalk &w) {
  bool over = w.d > w.best_te || w.d > w.t_sp;
  if (!over) {
    ++w.ncells;
    const size_t icell = (size_t)w.ci + (size_t)C.cgx * ((size_t)w.cj + (size_t)C.cgy * (size_t)w.ck);
    const int p0 = __ldg(C.cg_start + icell), p1 = __ldg(C.cg_start + icell + 1);
    for (int ip = p0; ip < p1; ++ip) {
      const int icl = __ldg(C.cg_list + ip - 1);
      const double4 g = ldg4(C.geo_reg + ip - 1);  // independent of the list load: offsets -> {list, geometry}
      if (icl == w.skip_icl) continue;
      const double rx = DSUB(w.x, g.x), ry = DSUB(w.y, g.y), rz = DSUB(w.z, g.z);
      const double b = DADD(DADD(DMUL(rx, w.kx), DMUL(ry, w.ky)), DMUL(rz, w.kz));
      double disc = DADD(DSUB(DMUL(b, b), DADD(DADD(DMUL(rx, rx), DMUL(ry, ry)), DMUL(rz, rz))), g.w);
      if (disc < 0.0) continue;
      disc = sqrt(disc);
      const double te = DSUB(-b, disc), tx2 = DADD(-b, disc);
      if (tx2 > 0.0 && te < w.best_te) { w.best_te = te; w.best_icl = icl; }
    }
    // cg_dx/kx for kx > 0, -cg_dx/kx for kx < 0 (clump_mod.f90:1440-1445): the same quotient, bit for bit, as cg_dx/|kx|
    if (w.tx <= w.ty && w.tx <= w.tz) {
      w.d = w.tx; w.ci += w.si;
      if (w.ci < 0 || w.ci >= C.cgx) over = true;
      else { if (w.delx < 0.0) w.delx = C.dx / fabs(w.kx); w.tx = DADD(w.tx, w.delx); }
    } else if (w.ty <= w.tz) {
      w.d = w.ty; w.cj += w.sj;
      if (w.cj < 0 || w.cj >= C.cgy) over = true;
      else { if (w.dely < 0.0) w.dely = C.dy / fabs(w.ky); w.ty = DADD(w.ty, w.dely); }
    } else {
      w.d = w.tz; w.ck += w.sk;
      if (w.ck < 0 || w.ck >= C.cgz) over = true;
      else { if (w.delz < 0.0) w.delz = C.dz / fabs(w.kz); w.tz = DADD(w.tz, w.delz); }
    }
    if (!over) prefetch_l1(C.cg_start + ((size_t)w.ci + (size_t)C.cgx * ((size_t)w.cj + (size_t)C.cgy * (size_t)w.ck)));
  }
  if (over && !(w.best_icl > 0 && w.best_te <= w.t_sp)) w.best_icl = 0;
  return over;
}
// the search found clump best_icl: move to its entry point, shift into its frame (raytrace_clump.f90:168-180, 247-254)
LART_DEV void cw_enter(const DevParams &P, ClumpWalk &w) {
  cw_advance(w, fmax(0.0, w.best_te));
  w.icl = w.best_icl;
  const ClumpPhys cp = load_clump(P.cl, w.icl);
  w.xfreq = DSUB(w.xfreq, ulos_clump(P, cp, w.kx, w.ky, w.kz));
  w.phase = CW_CLUMP;
}

// One step of raytrace_to_edge_clump(_capped).  Returns true when the walk is over; the depth is w.tau.
LART_DEV bool cw_edge_step(const DevParams &P, const double *vtab, ClumpWalk &w, double tau_max) {
  const DevClumps &C = P.cl;
  if (w.phase == CW_CLUMP) {
    const ClumpPhys cp = load_clump(C, w.icl);
    const double t_seg = clump_exit_dist(C, w.x, w.y, w.z, w.kx, w.ky, w.kz, w.icl);
    w.tau = DADD(w.tau, DMUL(kappa_clump(P, vtab, cp, w.xfreq), t_seg));
    ++w.nclumps;
    if (tau_max > 0.0 && w.tau >= tau_max) return true;
    cw_advance(w, t_seg);
    w.xfreq = DADD(w.xfreq, ulos_clump(P, cp, w.kx, w.ky, w.kz));
    w.skip_icl = w.icl;
    if (outside_sphere(C, w.x, w.y, w.z)) return true;
    w.phase = CW_FIND;
    return !cw_find_begin(C, w);  // the search set-up rides along: two kinds of step remain, clump segments and cells
  }
  if (w.phase == CW_FIND) return !cw_find_begin(C, w);
  if (cw_find_cell(C, w)) {
    if (w.best_icl == 0) return true;
    cw_enter(P, w);
  }
  return false;
}
// One step of raytrace_to_tau_clump.  0 = keep going, 1 = at the next scattering point (inside clump w.icl),
// 2 = left the sphere (w.xfreq is the box-frame frequency binned into Jout).
LART_DEV int cw_tau_step(const DevParams &P, const double *vtab, ClumpWalk &w) {
  const DevClumps &C = P.cl;
  if (w.phase == CW_CLUMP) {
    const ClumpPhys cp = load_clump(C, w.icl);
    const double t_seg = clump_exit_dist(C, w.x, w.y, w.z, w.kx, w.ky, w.kz, w.icl);
    const double kap = kappa_clump(P, vtab, cp, w.xfreq);
    if (w.tau <= DMUL(kap, t_seg)) {  // scatters inside this clump
      cw_advance(w, w.tau / fmax(kap, kTinyDouble));
      return 1;
    }
    w.tau = DSUB(w.tau, DMUL(kap, t_seg));
    cw_advance(w, t_seg);
    w.xfreq = DADD(w.xfreq, ulos_clump(P, cp, w.kx, w.ky, w.kz));
    w.skip_icl = w.icl;
    w.icl = 0;
    if (outside_sphere(C, w.x, w.y, w.z)) return 2;
    w.phase = CW_FIND;
    return cw_find_begin(C, w) ? 0 : 2;
  }
  if (w.phase == CW_FIND) return cw_find_begin(C, w) ? 0 : 2;
  if (cw_find_cell(C, w)) {
    if (w.best_icl == 0) { cw_advance(w, w.t_sp); return 2; }  // nothing ahead: to the sphere (:182-186)
    cw_enter(P, w);
    w.skip_icl = 0;
  }
  return 0;
}

// raytrace_to_edge_clump (:205-270; tau_max <= 0) and raytrace_to_edge_clump_capped (:494-533), run to completion
LART_DEV double clump_walk_edge(const DevParams &P, const double *vtab, double xp, double yp, double zp, double kx, double ky,
                                double kz, double xfreq, int icl_cur, double tau_max, int &ncells, int &nclumps) {
  ClumpWalk w;
  cw_start(w, xp, yp, zp, kx, ky, kz, xfreq, icl_cur, 0.0);
  while (!cw_edge_step(P, vtab, w, tau_max)) {}
  ncells += w.ncells; nclumps += w.nclumps;
  return w.tau;
}
// raytrace_to_tau_clump — raytrace_clump.f90:83-201, run to completion.  Returns true while the photon is inside (it then
// sits at its next scattering point, in clump icl); on escape ph.xfreq is the lab-frame frequency the caller bins into Jout.
LART_DEV void cw_finish_tau(const DevParams &P, const ClumpWalk &w, int st, Photon &ph, int &icl) {
  ph.x = w.x; ph.y = w.y; ph.z = w.z; ph.xfreq = w.xfreq; icl = (st == 1) ? w.icl : 0;
  update_cell_idx(P, ph);  // (upstream skips this on its t_sp <= 0 exit; the cell of an escaped photon is never read)
}
LART_DEV bool clump_walk_tau(const DevParams &P, const double *vtab, Photon &ph, int &icl, double tau_in, int &ncells) {
  ClumpWalk w;
  cw_start(w, ph.x, ph.y, ph.z, ph.kx, ph.ky, ph.kz, ph.xfreq, icl, tau_in);
  int st;
  while ((st = cw_tau_step(P, vtab, w)) == 0) {}
  ncells += w.ncells;
  cw_finish_tau(P, w, st, ph, icl);
  return st == 1;
}

// ---------------------------------------------------------------------------
// Overlapping populations (has_overlap; setup_clump_overlap, setup.f90:1051-1081) — the event walk of
// clump_mod.f90:1595-1760 and raytrace_clump.f90:621-920.  A ray's ENTER/EXIT events up to the bounding sphere are
// collected through the CSR grid, sorted by distance, and swept with the set of clumps the ray is inside; in an overlap
// region every clump adds its opacity at the ray's frequency in ITS frame.  The event list (up to MAX_EVT = 2048
// entries, as upstream) and the active set live in global memory, element e of thread t at [e*T + t] (a warp's lanes
// touch one line per element); one thread walks one ray.  Expression order is the reference's, so events, sums and the
// sampled owner clump coincide with the CPU restatement bit for bit.
// ---------------------------------------------------------------------------
constexpr int kMaxEvt = 2048;  // raytrace_clump.f90:680
struct OvList {
  double *t; int *ev; int *act;  // ev = +icl ENTER, -icl EXIT
  size_t T, tid;
  LART_DEV double &tt(int e) const { return t[(size_t)e * T + tid]; }
  LART_DEV int &ee(int e) const { return ev[(size_t)e * T + tid]; }
  LART_DEV int &aa(int e) const { return act[(size_t)e * T + tid]; }
};
LART_DEV OvList ov_list(const DevClumps &C) {
  OvList L;
  L.t = C.ov_t; L.ev = C.ov_ev; L.act = C.ov_act; L.T = (size_t)C.ov_T;
  L.tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  return L;
}
// active_set_at_point — clump_mod.f90:1595-1634: every clump that contains the point, first-seen order
LART_DEV int ov_active_set(const DevClumps &C, const OvList &L, double xp, double yp, double zp) {
  const int ci = cg_clamp(xp, C.xmin, C.inv_dx, C.cgx), cj = cg_clamp(yp, C.ymin, C.inv_dy, C.cgy), ck = cg_clamp(zp, C.zmin, C.inv_dz, C.cgz);
  int na = 0;
  for (int k = max(0, ck - 1); k <= min(C.cgz - 1, ck + 1); ++k)
    for (int j = max(0, cj - 1); j <= min(C.cgy - 1, cj + 1); ++j)
      for (int i = max(0, ci - 1); i <= min(C.cgx - 1, ci + 1); ++i) {
        const size_t icell = (size_t)i + (size_t)C.cgx * ((size_t)j + (size_t)C.cgy * (size_t)k);
        const int p0 = __ldg(C.cg_start + icell), p1 = __ldg(C.cg_start + icell + 1);
        for (int ip = p0; ip < p1; ++ip) {
          const int icl = __ldg(C.cg_list + ip - 1);
          const double4 g = ldg4(C.geo_reg + ip - 1);
          const double rx = DSUB(xp, g.x), ry = DSUB(yp, g.y), rz = DSUB(zp, g.z);
          if (!(DADD(DADD(DMUL(rx, rx), DMUL(ry, ry)), DMUL(rz, rz)) <= g.w)) continue;
          bool seen = false;
          for (int m = 0; m < na; ++m) seen = seen || L.aa(m) == icl;
          if (!seen && na < kMaxEvt) L.aa(na++) = icl;
        }
      }
  return na;
}
// collect_ray_events_overlap — clump_mod.f90:1639-1760: all events with 0 < t <= t_max, stably sorted by t
LART_DEV int ov_collect(const DevClumps &C, const OvList &L, double xp, double yp, double zp, double kx, double ky, double kz,
                        double t_max) {
  int n = 0;
  int ci = cg_clamp(xp, C.xmin, C.inv_dx, C.cgx), cj = cg_clamp(yp, C.ymin, C.inv_dy, C.cgy), ck = cg_clamp(zp, C.zmin, C.inv_dz, C.cgz);
  const int si = kx > 0.0 ? 1 : (kx < 0.0 ? -1 : 0), sj = ky > 0.0 ? 1 : (ky < 0.0 ? -1 : 0), sk = kz > 0.0 ? 1 : (kz < 0.0 ? -1 : 0);
  auto first = [](double k, double p, int cc, double lo, double dd, int st) {
    return st == 0 ? kHugest : DSUB(DADD(lo, DMUL((double)(cc + (st > 0 ? 1 : 0)), dd)), p) / k;
  };
  double tx = first(kx, xp, ci, C.xmin, C.dx, si), ty = first(ky, yp, cj, C.ymin, C.dy, sj), tz = first(kz, zp, ck, C.zmin, C.dz, sk);
  const double delx = si ? C.dx / fabs(kx) : kHugest, dely = sj ? C.dy / fabs(ky) : kHugest, delz = sk ? C.dz / fabs(kz) : kHugest;
  double d = 0.0;
  for (;;) {
    if (d > t_max) break;
    if (ci < 0 || ci >= C.cgx || cj < 0 || cj >= C.cgy || ck < 0 || ck >= C.cgz) break;
    const size_t icell = (size_t)ci + (size_t)C.cgx * ((size_t)cj + (size_t)C.cgy * (size_t)ck);
    const int p0 = __ldg(C.cg_start + icell), p1 = __ldg(C.cg_start + icell + 1);
    for (int ip = p0; ip < p1; ++ip) {
      const int icl = __ldg(C.cg_list + ip - 1);
      const double4 g = ldg4(C.geo_reg + ip - 1);
      const double rx = DSUB(xp, g.x), ry = DSUB(yp, g.y), rz = DSUB(zp, g.z);
      const double b = DADD(DADD(DMUL(rx, kx), DMUL(ry, ky)), DMUL(rz, kz));
      double disc = DADD(DSUB(DMUL(b, b), DADD(DADD(DMUL(rx, rx), DMUL(ry, ry)), DMUL(rz, rz))), g.w);
      if (disc < 0.0) continue;
      disc = sqrt(disc);
      const double te = DSUB(-b, disc), tx2 = DADD(-b, disc);
      if (!(tx2 > 0.0) || te > t_max) continue;
      bool dup = false;
      for (int e = 0; e < n; ++e) dup = dup || abs(L.ee(e)) == icl;
      if (dup || n + 2 > kMaxEvt) continue;
      if (te > 0.0) { L.tt(n) = te; L.ee(n) = icl; ++n; }
      L.tt(n) = fmin(tx2, t_max); L.ee(n) = -icl; ++n;
    }
    if (tx <= ty && tx <= tz) { d = tx; tx = DADD(tx, delx); ci += si; }
    else if (ty <= tz) { d = ty; ty = DADD(ty, dely); cj += sj; }
    else { d = tz; tz = DADD(tz, delz); ck += sk; }
  }
  for (int e = 1; e < n; ++e) {  // insertion sort, stable for equal t
    const double tt = L.tt(e);
    const int ii = L.ee(e);
    int q = e - 1;
    while (q >= 0 && L.tt(q) > tt) { L.tt(q + 1) = L.tt(q); L.ee(q + 1) = L.ee(q); --q; }
    L.tt(q + 1) = tt; L.ee(q + 1) = ii;
  }
  return n;
}
LART_DEV void ov_apply(const OvList &L, int ev, int &na) {
  if (ev > 0) { if (na < kMaxEvt) L.aa(na++) = ev; return; }
  for (int m = 0; m < na; ++m) if (L.aa(m) == -ev) { L.aa(m) = L.aa(na - 1); --na; return; }
}
// sum_kap_active — raytrace_clump.f90:621-640
LART_DEV double ov_sum_kap(const DevParams &P, const double *vtab, const OvList &L, int na, double xg, double kx, double ky, double kz) {
  double s = 0.0;
  for (int m = 0; m < na; ++m) {
    const ClumpPhys cp = load_clump(P.cl, L.aa(m));
    s = DADD(s, kappa_clump(P, vtab, cp, DSUB(xg, ulos_clump(P, cp, kx, ky, kz))));
  }
  return s;
}
// raytrace_to_edge_clump_overlap (:792-855; tau_max <= 0) and _overlap_capped (:858-920)
LART_DEV double clump_walk_edge_overlap(const DevParams &P, const double *vtab, double xp, double yp, double zp, double kx, double ky,
                                        double kz, double xg, double tau_max, int &nevents) {
  const DevClumps &C = P.cl;
  double tau = 0.0;
  const double t_sp = sphere_exit_dist(C, xp, yp, zp, kx, ky, kz);
  if (t_sp <= 0.0) return tau;
  const OvList L = ov_list(C);
  const int n = ov_collect(C, L, xp, yp, zp, kx, ky, kz, t_sp);
  int na = ov_active_set(C, L, xp, yp, zp);
  nevents += n;
  double t_cur = 0.0;
  for (int ie = 0; ie <= n; ++ie) {
    const double t_next = fmin(ie < n ? L.tt(ie) : t_sp, t_sp);
    const double dt = DSUB(t_next, t_cur);
    if (dt > 0.0) {
      tau = DADD(tau, DMUL(ov_sum_kap(P, vtab, L, na, xg, kx, ky, kz), dt));
      if (tau_max > 0.0 && tau >= tau_max) return tau;
    }
    t_cur = t_next;
    if (ie < n) ov_apply(L, L.ee(ie), na);
  }
  return tau;
}
// raytrace_to_tau_clump_overlap — raytrace_clump.f90:668-790.  true = the photon sits at its next scattering point, owned by
// clump icl (drawn with one uniform among the clumps that overlap there, sample_owner_clump :642-666); false = it left the
// sphere (ph.xfreq stays the global-frame frequency that is binned into Jout).
template <class RngT>
LART_DEV bool clump_walk_tau_overlap(const DevParams &P, const double *vtab, Photon &ph, int &icl, double tau_in, RngT &rng, int &nevents) {
  const DevClumps &C = P.cl;
  const double kx = ph.kx, ky = ph.ky, kz = ph.kz, xg = ph.xfreq;
  double tau_rem = tau_in;
  const double t_sp = sphere_exit_dist(C, ph.x, ph.y, ph.z, kx, ky, kz);
  if (t_sp <= 0.0) { icl = 0; update_cell_idx(P, ph); return false; }
  const OvList L = ov_list(C);
  const int n = ov_collect(C, L, ph.x, ph.y, ph.z, kx, ky, kz, t_sp);
  int na = ov_active_set(C, L, ph.x, ph.y, ph.z);
  nevents += n;
  double t_cur = 0.0;
  for (int ie = 0; ie <= n; ++ie) {
    const double t_next = fmin(ie < n ? L.tt(ie) : t_sp, t_sp);
    const double dt = DSUB(t_next, t_cur);
    if (dt <= 0.0) {
      if (ie < n) ov_apply(L, L.ee(ie), na);
      continue;
    }
    const double kap_tot = ov_sum_kap(P, vtab, L, na, xg, kx, ky, kz);
    if (kap_tot > 0.0 && tau_rem <= DMUL(kap_tot, dt)) {  // scatters inside this segment
      const double s = DADD(t_cur, tau_rem / kap_tot);
      ph.x = DADD(ph.x, DMUL(s, kx)); ph.y = DADD(ph.y, DMUL(s, ky)); ph.z = DADD(ph.z, DMUL(s, kz));
      const double rnd = rng.uniform();
      const double lim = DMUL(rnd, kap_tot);
      double cumul = 0.0;
      int owner = 0;
      for (int m = 0; m < na; ++m) {
        const ClumpPhys cp = load_clump(C, L.aa(m));
        cumul = DADD(cumul, kappa_clump(P, vtab, cp, DSUB(xg, ulos_clump(P, cp, kx, ky, kz))));
        owner = L.aa(m);
        if (lim <= cumul) break;
      }
      icl = owner;
      update_cell_idx(P, ph);
      return true;
    }
    tau_rem = DSUB(tau_rem, DMUL(kap_tot, dt));
    t_cur = t_next;
    if (ie < n) ov_apply(L, L.ee(ie), na);
  }
  ph.x = DADD(ph.x, DMUL(t_sp, kx)); ph.y = DADD(ph.y, DMUL(t_sp, ky)); ph.z = DADD(ph.z, DMUL(t_sp, kz));
  icl = 0;
  update_cell_idx(P, ph);
  return false;
}

}  // namespace lart
