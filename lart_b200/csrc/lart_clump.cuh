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

// find_next_clump — clump_mod.f90:1393-1506: Amanatides-Woo walk through the CSR grid, ray-sphere test of every
// clump registered in the cell, nearest entry wins; stops once the cell starts beyond the best hit or t_max.
LART_DEV bool find_next_clump(const DevClumps &C, double xp, double yp, double zp, double kx, double ky, double kz, int skip_icl,
                              double t_max, double &t_entry, int &icl_found, int &ncells) {
  double best_te = kHugest, d = 0.0;
  int best_icl = 0;
  int ci = cg_clamp(xp, C.xmin, C.inv_dx, C.cgx), cj = cg_clamp(yp, C.ymin, C.inv_dy, C.cgy), ck = cg_clamp(zp, C.zmin, C.inv_dz, C.cgz);
  int si, sj, sk;
  double tx, ty, tz, delx, dely, delz;
  auto axis = [](double k, double p, int cc, double lo, double dd, int &st, double &t, double &del) {
    if (k > 0.0) { st = 1; t = DSUB(DADD(lo, DMUL((double)(cc + 1), dd)), p) / k; del = dd / k; }
    else if (k < 0.0) { st = -1; t = DSUB(DADD(lo, DMUL((double)cc, dd)), p) / k; del = -dd / k; }
    else { st = 0; t = kHugest; del = kHugest; }
  };
  axis(kx, xp, ci, C.xmin, C.dx, si, tx, delx);
  axis(ky, yp, cj, C.ymin, C.dy, sj, ty, dely);
  axis(kz, zp, ck, C.zmin, C.dz, sk, tz, delz);
  for (;;) {
    if (d > best_te || d > t_max) break;
    ++ncells;
    const size_t icell = (size_t)ci + (size_t)C.cgx * ((size_t)cj + (size_t)C.cgy * (size_t)ck);
    const int p0 = __ldg(C.cg_start + icell), p1 = __ldg(C.cg_start + icell + 1);
    for (int ip = p0; ip < p1; ++ip) {
      const int icl = __ldg(C.cg_list + ip - 1);
      if (icl == skip_icl) continue;
      const double4 g = ldg4(C.geo + icl - 1);
      const double rx = DSUB(xp, g.x), ry = DSUB(yp, g.y), rz = DSUB(zp, g.z);
      const double b = DADD(DADD(DMUL(rx, kx), DMUL(ry, ky)), DMUL(rz, kz));
      double disc = DADD(DSUB(DMUL(b, b), DADD(DADD(DMUL(rx, rx), DMUL(ry, ry)), DMUL(rz, rz))), g.w);
      if (disc < 0.0) continue;
      disc = sqrt(disc);
      const double te = DSUB(-b, disc), tx2 = DADD(-b, disc);
      if (tx2 > 0.0 && te < best_te) { best_te = te; best_icl = icl; }  // (te > 0 .or. icl /= skip) holds: icl /= skip here
    }
    if (tx <= ty && tx <= tz) { d = tx; ci += si; if (ci < 0 || ci >= C.cgx) break; tx = DADD(tx, delx); }
    else if (ty <= tz) { d = ty; cj += sj; if (cj < 0 || cj >= C.cgy) break; ty = DADD(ty, dely); }
    else { d = tz; ck += sk; if (ck < 0 || ck >= C.cgz) break; tz = DADD(tz, delz); }
  }
  if (best_icl > 0 && best_te <= t_max) { t_entry = best_te; icl_found = best_icl; return true; }
  return false;
}

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

// raytrace_to_edge_clump (:205-270; tau_max <= 0) and raytrace_to_edge_clump_capped (:494-533)
LART_DEV double clump_walk_edge(const DevParams &P, const double *vtab, double xp, double yp, double zp, double kx, double ky,
                                double kz, double xfreq, int icl_cur, double tau_max, int &ncells, int &nclumps) {
  const DevClumps &C = P.cl;
  const bool capped = tau_max > 0.0;
  double tau = 0.0;
  if (icl_cur > 0) {
    const ClumpPhys cp = load_clump(C, icl_cur);
    const double t_seg = clump_exit_dist(C, xp, yp, zp, kx, ky, kz, icl_cur);
    tau = DADD(tau, DMUL(kappa_clump(P, vtab, cp, xfreq), t_seg));
    ++nclumps;
    if (capped && tau >= tau_max) return tau;
    xp = DADD(xp, DMUL(t_seg, kx)); yp = DADD(yp, DMUL(t_seg, ky)); zp = DADD(zp, DMUL(t_seg, kz));
    xfreq = DADD(xfreq, ulos_clump(P, cp, kx, ky, kz));
    if (outside_sphere(C, xp, yp, zp)) return tau;
  }
  for (;;) {
    const double t_sp = sphere_exit_dist(C, xp, yp, zp, kx, ky, kz);
    if (t_sp <= 0.0) break;
    double te;
    int icl_found = 0;
    if (!find_next_clump(C, xp, yp, zp, kx, ky, kz, icl_cur, t_sp, te, icl_found, ncells)) break;
    te = fmax(0.0, te);
    xp = DADD(xp, DMUL(te, kx)); yp = DADD(yp, DMUL(te, ky)); zp = DADD(zp, DMUL(te, kz));
    const ClumpPhys cp = load_clump(C, icl_found);
    const double u_los = ulos_clump(P, cp, kx, ky, kz);
    xfreq = DSUB(xfreq, u_los);
    const double t_seg = clump_exit_dist(C, xp, yp, zp, kx, ky, kz, icl_found);
    tau = DADD(tau, DMUL(kappa_clump(P, vtab, cp, xfreq), t_seg));
    ++nclumps;
    if (capped && tau >= tau_max) return tau;
    xp = DADD(xp, DMUL(t_seg, kx)); yp = DADD(yp, DMUL(t_seg, ky)); zp = DADD(zp, DMUL(t_seg, kz));
    xfreq = DADD(xfreq, u_los);
    icl_cur = icl_found;
    if (outside_sphere(C, xp, yp, zp)) break;
  }
  return tau;
}

// raytrace_to_tau_clump — raytrace_clump.f90:83-201.  Returns true while the photon is inside (it then sits at its
// next scattering point, in clump ph.icl); on escape ph.xfreq is the lab-frame frequency the caller bins into Jout.
LART_DEV bool clump_walk_tau(const DevParams &P, const double *vtab, Photon &ph, int &icl, double tau_in, int &ncells) {
  const DevClumps &C = P.cl;
  const double kx = ph.kx, ky = ph.ky, kz = ph.kz;
  double tau_rem = tau_in;
  int last_icl = 0;
  for (;;) {
    if (icl > 0) {
      const ClumpPhys cp = load_clump(C, icl);
      const double t_seg = clump_exit_dist(C, ph.x, ph.y, ph.z, kx, ky, kz, icl);
      const double kap = kappa_clump(P, vtab, cp, ph.xfreq);
      if (tau_rem <= DMUL(kap, t_seg)) {  // scatters inside this clump
        const double ds = tau_rem / fmax(kap, kTinyDouble);
        ph.x = DADD(ph.x, DMUL(ds, kx)); ph.y = DADD(ph.y, DMUL(ds, ky)); ph.z = DADD(ph.z, DMUL(ds, kz));
        update_cell_idx(P, ph);
        return true;
      }
      tau_rem = DSUB(tau_rem, DMUL(kap, t_seg));
      ph.x = DADD(ph.x, DMUL(t_seg, kx)); ph.y = DADD(ph.y, DMUL(t_seg, ky)); ph.z = DADD(ph.z, DMUL(t_seg, kz));
      ph.xfreq = DADD(ph.xfreq, ulos_clump(P, cp, kx, ky, kz));
      last_icl = icl;
      icl = 0;
      if (outside_sphere(C, ph.x, ph.y, ph.z)) { update_cell_idx(P, ph); return false; }
    } else {
      const double t_sp = sphere_exit_dist(C, ph.x, ph.y, ph.z, kx, ky, kz);
      if (t_sp <= 0.0) return false;
      double te;
      int icl_found = 0;
      if (find_next_clump(C, ph.x, ph.y, ph.z, kx, ky, kz, last_icl, t_sp, te, icl_found, ncells)) {
        te = fmax(0.0, te);
        ph.x = DADD(ph.x, DMUL(te, kx)); ph.y = DADD(ph.y, DMUL(te, ky)); ph.z = DADD(ph.z, DMUL(te, kz));
        const ClumpPhys cp = load_clump(C, icl_found);
        ph.xfreq = DSUB(ph.xfreq, ulos_clump(P, cp, kx, ky, kz));
        last_icl = 0;
        icl = icl_found;
        update_cell_idx(P, ph);
      } else {
        ph.x = DADD(ph.x, DMUL(t_sp, kx)); ph.y = DADD(ph.y, DMUL(t_sp, ky)); ph.z = DADD(ph.z, DMUL(t_sp, kz));
        update_cell_idx(P, ph);
        return false;
      }
    }
  }
}

}  // namespace lart
