// lart_engine.cu — the CUDA engine behind include/lart_gpu.h (sm_100a).
//
// Two drivers of the same device physics (lart_device.cuh):
//
//  * WAVEFRONT (default).  The photon pool is SoA in HBM; one "wave" advances every
//    live photon by one scattering through four stage kernels whose work lists are
//    stream-compacted on the device:
//       emit    dead slots pull photon ids from the job queue (generate_photon +
//               peeling_direct ray descriptors)
//       trace   raytrace_to_tau for every live photon (forced first scattering
//               included); escapes are tallied and retire their slot, the rest are
//               appended to the scatter list
//       scatter frequency redistribution + phase function + Stokes update; one
//               peel-ray descriptor per observer is appended to the ray queue
//       peel    raytrace_to_edge for every queued ray, exp(-tau) deposit with
//               warp-aggregated atomics
//    trace and peel are persistent kernels with per-lane refill: a lane whose ray
//    ended pulls the next work item while its neighbours keep stepping, so a warp
//    always executes the same DDA step body.
//  * MONOLITHIC (LART_FLAG_MONOLITHIC).  One thread owns one photon slot and runs
//    trace -> scatter -> peel back to back for `quantum` events.  Kept as the
//    "before compaction" baseline for the warp-efficiency evidence, and as a second
//    implementation the parity tests cross-check.
//
// Photon physics is independent of the scheduling: every photon has its own Philox
// stream (seed, photon id) and its draw counter travels with the photon, so both
// drivers — and the CPU oracle in Philox mode — produce the same photon histories.
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>
#include <nccl.h>  // types and prototypes only: the library is opened at run time (lart_gpu_comm_init), never linked

#include <algorithm>
#include <chrono>
#include <atomic>
#include <thread>
#include <cmath>
#include <map>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cerrno>
#include <functional>
#include <mutex>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/lart_gpu.h"
#include "lart_device.cuh"
#include "lart_clump.cuh"
#include "voigt_tables.cuh"


using namespace lart;

namespace {

thread_local std::string g_err;

int fail(const std::string &msg) {
  g_err = msg;
  return 1;
}
#define CUDA_OK(call)                                                                                   \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess)                                                                              \
      return fail(std::string(#call) + " failed: " + cudaGetErrorString(e_) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
  } while (0)

constexpr int kVoigtTabN = 202 * 4;
constexpr int kBlock = 256;
constexpr int64_t kTailPhotons = 131072;  // lart_gpu_run: below this many photons left, finish with k_mono (+ deferred peel rays)
constexpr int kTailQuantum = 256;

// ------------------------------- photon pool -------------------------------
enum {
  F_X, F_Y, F_Z, F_KX, F_KY, F_KZ, F_MX, F_MY, F_MZ, F_NX, F_NY, F_NZ,
  F_XFREQ, F_XREF, F_WGT, F_Q, F_U, F_V, F_NSG, F_NSD, F_GSET, F_TAU,
  F_SHEAR,  // photon%vfy_shear (shearing boxes only; the column is never touched otherwise)
  F_COUNT
};
struct Pool {
  double *f;                   // [F_COUNT][S]
  long long *id;               // [S]
  unsigned long long *ndraw;   // [S]
  int *ic, *jc, *kc, *flags;   // [S]
  double *rs;                  // [10][S] DDA state of a flight suspended at its per-wave step budget
  int *rc;                     // [3][S]  ... and its current cell
  int *nev;                    // [S] scatterings of the slot's photon so far (bounded runs, lart_config::max_events)
  double *var;                 // [6][S] variates of the slot's pending scattering: uz, cos(theta), cos(phi), sin(phi), ux, uy
  double *wtab;                // [12][S] scratch of k_wf_draw2: wing majorant tables (10), core-skip threshold, polar-pair r^2
  int *lst;                    // [S] scratch of k_wf_draw2: per warp chunk, core photons from the front, wing photons from the back
  int S;                       // slots (= SoA stride)
  int s0, n;                   // the partition [s0, s0+n) this kernel launch works on
};
struct Job {  // photon ids = first_id + j*stride, j in [0,count)
  unsigned long long next, count, done;
  long long first_id, stride;
  unsigned int err, pad_;  // sticky ERR_* bits (DevParams::err points here); read back with the job after every step
};
struct Queues {  // one per pool partition (pipeline)
  PeelRay *rays;      // [ray_cap]: S*nobs slot rays, then per partition the direct rays of this wave's emits
  unsigned int *n_direct, *head_trace, *head_peel;
  unsigned int *n_dead;   // dead slots of the partition (the emit stage returns at once when there is none)
  unsigned int direct_base, direct_cap;  // direct-ray region of this partition
  PeelCont *cont[2];  // peel rays suspended at their step budget: read from [wave&1], appended to [(wave&1)^1]
  unsigned int *n_cont;   // [2]
  unsigned int *wave;     // wave counter of this partition (its parity selects the buffers)
  unsigned int cont_cap;
};

__device__ __forceinline__ void load_trace_part(const Pool &pl, int s, Photon &ph) {
  const double *f = pl.f;
  size_t S = pl.S;
  ph.x = f[F_X * S + s]; ph.y = f[F_Y * S + s]; ph.z = f[F_Z * S + s];
  ph.kx = f[F_KX * S + s]; ph.ky = f[F_KY * S + s]; ph.kz = f[F_KZ * S + s];
  ph.xfreq = f[F_XFREQ * S + s]; ph.wgt = f[F_WGT * S + s];
  ph.id = pl.id[s]; ph.ic = pl.ic[s]; ph.jc = pl.jc[s]; ph.kc = pl.kc[s]; ph.flags = pl.flags[s];
}
__device__ __forceinline__ void load_stokes(const Pool &pl, int s, Photon &ph) {
  const double *f = pl.f;
  size_t S = pl.S;
  ph.Q = f[F_Q * S + s]; ph.U = f[F_U * S + s]; ph.V = f[F_V * S + s];
}
__device__ __forceinline__ void load_triad(const Pool &pl, int s, Photon &ph) {
  const double *f = pl.f;
  size_t S = pl.S;
  ph.mx = f[F_MX * S + s]; ph.my = f[F_MY * S + s]; ph.mz = f[F_MZ * S + s];
  ph.nx = f[F_NX * S + s]; ph.ny = f[F_NY * S + s]; ph.nz = f[F_NZ * S + s];
  ph.xfreq_ref = f[F_XREF * S + s]; ph.nsg = f[F_NSG * S + s]; ph.nsd = f[F_NSD * S + s];
}
__device__ __forceinline__ void load_rest(const Pool &pl, int s, Photon &ph) {
  load_stokes(pl, s, ph);
  load_triad(pl, s, ph);
}
__device__ __forceinline__ void store_trace_part(const Pool &pl, int s, const Photon &ph) {
  double *f = pl.f;
  size_t S = pl.S;
  f[F_X * S + s] = ph.x; f[F_Y * S + s] = ph.y; f[F_Z * S + s] = ph.z;
  f[F_XFREQ * S + s] = ph.xfreq; f[F_WGT * S + s] = ph.wgt;
  pl.ic[s] = ph.ic; pl.jc[s] = ph.jc; pl.kc[s] = ph.kc; pl.flags[s] = ph.flags;
}
__device__ __forceinline__ void store_all(const Pool &pl, int s, const Photon &ph) {
  double *f = pl.f;
  size_t S = pl.S;
  store_trace_part(pl, s, ph);
  f[F_KX * S + s] = ph.kx; f[F_KY * S + s] = ph.ky; f[F_KZ * S + s] = ph.kz;
  f[F_MX * S + s] = ph.mx; f[F_MY * S + s] = ph.my; f[F_MZ * S + s] = ph.mz;
  f[F_NX * S + s] = ph.nx; f[F_NY * S + s] = ph.ny; f[F_NZ * S + s] = ph.nz;
  f[F_XREF * S + s] = ph.xfreq_ref; f[F_Q * S + s] = ph.Q; f[F_U * S + s] = ph.U; f[F_V * S + s] = ph.V;
  f[F_NSG * S + s] = ph.nsg; f[F_NSD * S + s] = ph.nsd;
  pl.id[s] = ph.id;
}
__device__ __forceinline__ void load_rng(const DevParams &P, const Pool &pl, int s, long long id, int flags, Rng &r) {
  r.start(P.seed, (unsigned long long)id, pl.ndraw[s]);  // ndraw = Philox blocks consumed so far
  if (flags & PH_GAUSS) { r.gauss_stored = true; r.gset = pl.f[(size_t)F_GSET * pl.S + s]; }
}
__device__ __forceinline__ void store_rng(const Pool &pl, int s, const Rng &r, int &flags) {
  pl.ndraw[s] = r.nblk;
  if (r.gauss_stored) { flags |= PH_GAUSS; pl.f[(size_t)F_GSET * pl.S + s] = r.gset; }
  else flags &= ~PH_GAUSS;
}

__device__ __forceinline__ void load_vtab(const DevParams &P, double *vtab) {
  for (int i = threadIdx.x; i < kVoigtTabN; i += blockDim.x) vtab[i] = P.voigt_tab[i];
  __syncthreads();
}

// warp-reduce the work counters and add them to the tally buffer (as doubles)
// Work counters: warp shuffle sum, then a block sum through shared memory, then ONE atomic per counter
// and block.  (Per-warp atomics on six hot addresses cost ~28 us per kernel: 14 k same-address FP64 REDs
// serialise in one L2 slice — the floor of every wave when few photons are in flight.)
__device__ void flush_counters(const DevParams &P, Counters &c, ctr_t nrng) {
  constexpr int NC = 7;
  __shared__ unsigned long long part[NC][kBlock / 32];
  c.rng += nrng;
  unsigned long long v[NC] = {c.photons, c.scatter, c.cellsteps, c.peel, c.rng, c.reject, c.peel_bound};  // widened for the sums
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < NC; ++q) {
    unsigned long long x = v[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) part[q][warp] = x;
  }
  __syncthreads();
  if (threadIdx.x < NC) {
    unsigned long long x = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) x += part[threadIdx.x][w];
    if (x) atomicAdd(P.tally + P.lay.counters + threadIdx.x, (double)x);
  }
}

// full walks (monolithic driver, batch API)
template <bool PLAIN = false>
__device__ __forceinline__ double walk_edge(const DevParams &P, const double *vtab, double x, double y, double z,
                                            double kx, double ky, double kz, int ic, int jc, int kc, double xfreq,
                                            int &nsteps, int trace_cap = 0, int *trace = nullptr) {
  Ray r;
  nsteps = 0;
  if (ray_setup<PLAIN>(P, r, x, y, z, kx, ky, kz, ic, jc, kc, xfreq, false)) return 0.0;
  for (;;) {
    if (trace && r.nsteps < trace_cap) trace[r.nsteps] = (int)cell_index(P, r.ic, r.jc, r.kc);
    if (edge_step<PLAIN>(P, vtab, r)) break;
  }
  nsteps = r.nsteps;
  return r.tau;
}

// raytrace_to_tau on a photon; returns the number of cell steps.  On escape the
// photon is flagged dead and carries the lab-frame xfreq_ref (raytrace_car.f90:1598-1623);
// the Jout tally is the caller's.
// destroyed = the walk ended in a masked cell of a spherical atmosphere (:3316-3327: same binning, into Jabs2); a plane
// atmosphere absorbs what leaves through its bottom cell (:3099-3107).
template <bool PLAIN = false>
__device__ __forceinline__ int finish_escape(const DevParams &P, Photon &ph, const Ray &r, bool destroyed = false) {
  ph.flags &= ~PH_ALIVE;
  ph.xfreq = DADD(r.xfreq, r.u1);
  ph.xfreq_ref = DMUL(ph.xfreq, r.cell.Dfreq / P.Dfreq_ref);
  ph.x = r.x0; ph.y = r.y0; ph.z = r.z0;
  if (PLAIN) {
    if (!P.zonly) { ph.ic = r.ic; ph.jc = r.jc; }
    ph.kc = r.kc;
    return r.nsteps;
  }
  if (P.x.atm && (destroyed || (P.x.atm == 1 && !(r.kc > 1)))) ph.flags |= PH_ABS2;
  if (P.x.shear) ph.vshear = r.vshear;
  if (P.amr.on) { ph.x = r.tx; ph.y = r.ty; ph.z = r.tz; }  // the running position at the last face (raytrace_amr.f90:219-222)
  if (P.bcxy == BC_MIRROR) {  // raytrace_car.f90:1934-1943: the end point and the reflected direction are stored even on escape
    ray_endpoint_bc(P, r, ph.x, ph.y, ph.z);
    ph.kx = r.kx; ph.ky = r.ky; ph.kz = r.kz;
  } else if (P.bcxy == BC_PERIODIC) {  // :2506-2508
    fold_periodic(P, ph.x, ph.y);
  }
  if (!P.zonly) { ph.ic = r.ic; ph.jc = r.jc; }
  ph.kc = r.kc;
  return r.nsteps;
}
// xyz symmetry: a photon that was reflected on its way carries the reflected direction from here on (:1941-1943);
// its polarisation triad is left as it was, as in the reference.
__device__ __forceinline__ void adopt_direction(const Ray &r, Photon &ph) { ph.kx = r.kx; ph.ky = r.ky; ph.kz = r.kz; }
__device__ __forceinline__ void store_direction(const Pool &pl, int s, const Photon &ph) {
  pl.f[F_KX * pl.S + s] = ph.kx; pl.f[F_KY * pl.S + s] = ph.ky; pl.f[F_KZ * pl.S + s] = ph.kz;
}
template <bool PLAIN = false>
__device__ __forceinline__ int walk_tau(const DevParams &P, const double *vtab, Photon &ph, double tau_in, CellData &cs) {
  Ray r;
  if (!PLAIN && P.x.shear) r.vshear = ph.vshear;
  if (ray_setup<PLAIN>(P, r, ph.x, ph.y, ph.z, ph.kx, ph.ky, ph.kz, ph.ic, ph.jc, ph.kc, ph.xfreq, true)) {
    ph.flags &= ~PH_ALIVE;  // raytrace_car.f90:1469-1472: returns before any update
    return -1;
  }
  for (;;) {
    double xp, yp, zp;
    int st = tau_step<PLAIN>(P, vtab, r, tau_in, xp, yp, zp, ph.wgt);
    if (st == 1) {
      ph.x = xp; ph.y = yp; ph.z = zp; ph.xfreq = r.xfreq;
      if (!PLAIN && P.bcxy == BC_MIRROR) adopt_direction(r, ph);
      if (!P.zonly) { ph.ic = r.ic; ph.jc = r.jc; }
      ph.kc = r.kc;
      if (!PLAIN && P.x.shear) ph.vshear = r.vshear;
      cs = r.cell;
      return r.nsteps;
    }
    if (st >= 2) return finish_escape<PLAIN>(P, ph, r, st == 3);
  }
}

// photon left the system: Jout (if it was walked), allph record, per-run sums
__device__ __forceinline__ void retire_photon(const DevParams &P, const Photon &ph, bool tally_jout, Job *job, Counters &cnt) {
  if (tally_jout) {
    if (ph.flags & PH_ABS2) {  // absorbed by the atmosphere's molecular zone
      const int ix = freq_bin(P, ph.xfreq_ref);
      if (ix >= 1 && ix <= P.nxfreq) tally_add(P.tally + P.lay.Jabs2 + ix - 1, ph.wgt);
    } else {
      tally_Jout(P, ph.xfreq_ref, ph.kz, ph.wgt);
    }
  }
  if (ph.nsg != 0.0) atomicAdd(P.tally + P.lay.scalars + 0, ph.nsg);
  if (ph.nsd != 0.0) atomicAdd(P.tally + P.lay.scalars + 1, ph.nsd);
  if (P.save_all_photons) record_final(P, ph);
  atomicAdd(&job->done, 1ULL);
  cnt.photons += 1;
}

// add_escaped_fraction_to_Jout + weight/tau of the forced first scattering —
// run_simulation_mod.f90:163-177, 208-247.  cs = record of the photon's cell.
__device__ __forceinline__ double forced_first(const DevParams &P, Photon &ph, Rng &rng, const CellData &cs, double tau0) {
  double wgt_esc = ph.wgt * exp(-tau0);
  double xref = (ph.xfreq + vdotk(cs, ph.kx, ph.ky, ph.kz)) * (cs.Dfreq / P.Dfreq_ref);
  tally_Jout(P, xref, ph.kz, wgt_esc);
  double wgt1 = 1.0 - exp(-tau0);
  ph.wgt = ph.wgt * wgt1;
  ph.flags &= ~PH_FIRST;
  return (tau0 > 0.0) ? -log(1.0 - rng.uniform() * wgt1) : kHugest;
}

__device__ __forceinline__ void clamp_cell_for_read(const DevParams &P, const Photon &ph, int &ci, int &cj, int &ck) {
  ci = min(max(ph.ic, 1), P.nx); cj = min(max(ph.jc, 1), P.ny); ck = min(max(ph.kc, 1), P.nz);
}

// ------------------------------ monolithic driver ---------------------------
__device__ __forceinline__ void ray_append(const DevParams &P, const Queues &q, const PeelRay &pr);
// DEFER (the tail of lart_gpu_run): peel rays are appended to the ray queue q and walked by k_wf_peel after this launch.
// A photon's next scattering does not depend on its peel rays, but walked inline they are most of the time between two
// scatterings of a thread (tau0 = 1e4 sphere: ~17 cells per ray, 97 us per event with 8 k photons left) — and in the
// tail that latency, times the scatterings the longest-lived photon still needs, is the run time.
template <bool DEFER, bool PLAIN>
__global__ void __launch_bounds__(kBlock) k_mono(const __grid_constant__ DevParams P, Pool pl, Job *job, int quantum, Queues q) {
  __shared__ double vtab[kVoigtTabN];
  load_vtab(P, vtab);
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  Counters cnt;
  ctr_t nrng = 0;
  if (s < pl.n) {  // the monolithic driver works on the (possibly compacted) range [0, n)
    Photon ph;
    Rng rng;
    ph.flags = pl.flags[s];
    bool rng_valid = false, touched = false;
    int nev = 0;  // scatterings of this photon so far (bounded runs)
    if (ph.flags & PH_ALIVE) {
      load_trace_part(pl, s, ph);
      load_rest(pl, s, ph);
      load_rng(P, pl, s, ph.id, ph.flags, rng);
      if (!PLAIN && P.x.shear) ph.vshear = pl.f[(size_t)F_SHEAR * pl.S + s];
      nev = pl.nev[s];
      rng_valid = true;
    }
    bool at_scatter = (ph.flags & PH_ALIVE) && (ph.flags & PH_SCATTER);
    ph.flags &= ~PH_SCATTER;
    for (int ev = 0; ev < quantum; ++ev) {
      CellData cs;
      if (!at_scatter) {
        if (!(ph.flags & PH_ALIVE)) {
          if (job->next >= job->count) break;
          unsigned long long j = atomicAdd(&job->next, 1ULL);
          if (j >= job->count) break;
          touched = true;
          ph.id = job->first_id + (long long)j * job->stride;
          if (rng_valid) nrng += rng.nrng;  // draws of the previous photon in this slot
          rng.start(P.seed, (unsigned long long)ph.id);
          rng_valid = true;
          nev = 0;
          generate_photon(P, ph, rng, cnt, cs);
          if (P.save_all_photons) record_initial(P, ph);
          if (P.save_peeloff) {  // peeling_direct — generate_photon.f90:334-336
            for (int i = 0; i < P.nobs; ++i) {
              PeelRay pr;
              if (!peel_direct_prepare(P, P.obs[i], i, ph, cs, pr)) continue;
              if (DEFER) { ray_append(P, q, pr); continue; }
              int ns;
              double tau = walk_edge<PLAIN>(P, vtab, pr.x, pr.y, pr.z, pr.kx, pr.ky, pr.kz, pr.ic, pr.jc, pr.kc, pr.xfreq, ns);
              cnt.cellsteps += ns; cnt.peel += 1;
              peel_deposit(P, pr, tau, __activemask());
            }
          }
        }
        touched = true;
        double tau;
        if (ph.flags & PH_FIRST) {
          int ci, cj, ck, ns;
          clamp_cell_for_read(P, ph, ci, cj, ck);
          load_cell<PLAIN>(P, ci, cj, ck, cs);
          double tau0 = walk_edge<PLAIN>(P, vtab, ph.x, ph.y, ph.z, ph.kx, ph.ky, ph.kz, ph.ic, ph.jc, ph.kc, ph.xfreq, ns);
          cnt.cellsteps += ns;
          tau = forced_first(P, ph, rng, cs, tau0);
        } else if (ph.flags & PH_TAUPEND) {  // drawn by the wavefront scatter stage before a driver switch
          tau = pl.f[(size_t)F_TAU * pl.S + s];
          ph.flags &= ~PH_TAUPEND;
        } else {
          tau = -log(rng.uniform());
        }
        int ns = walk_tau<PLAIN>(P, vtab, ph, tau, cs);
        if (ns > 0) cnt.cellsteps += ns;
        if (!(ph.flags & PH_ALIVE)) {
          retire_photon(P, ph, ns >= 0, job, cnt);
          continue;
        }
      } else {  // the wavefront scatter stage already flew this photon to its next scattering point
        load_cell<PLAIN>(P, ph.ic, ph.jc, ph.kc, cs);
        at_scatter = false;
        touched = true;
      }
      // scattering — scattering_car.f90:14-120
      cnt.scatter += 1;
      bool to_dust = false;
      if (P.dust) {
        double pd = cs.rhokapD / (cs.rhokap * voigt_seon2(vtab, ph.xfreq, cs.voigt_a) + cs.rhokapD);
        to_dust = rng.uniform() <= pd;
      }
      auto trace_and_deposit = [&](PeelRay &pr) {
        if (DEFER) { ray_append(P, q, pr); return; }
        int ns2;
        double t = walk_edge<PLAIN>(P, vtab, pr.x, pr.y, pr.z, pr.kx, pr.ky, pr.kz, pr.ic, pr.jc, pr.kc, pr.xfreq, ns2);
        cnt.cellsteps += ns2; cnt.peel += 1;
        peel_deposit(P, pr, t, __activemask());
      };
      if (to_dust) {
        scatter_dust(P, ph, rng, cs, cnt, [&]() {
          for (int i = 0; i < P.nobs; ++i) {
            PeelRay pr;
            bool ok = P.use_stokes ? peel_dust_stokes_prepare(P, P.obs[i], i, ph, cs, pr)
                                   : peel_dust_nostokes_prepare(P, P.obs[i], i, ph, cs, pr);
            if (ok) trace_and_deposit(pr);
          }
        });
        if (!(ph.flags & PH_ALIVE)) retire_photon(P, ph, false, job, cnt);
      } else {
        const double uz_s = rand_resonance_vz(rng, ph.xfreq, cs.voigt_a, cnt.reject);
        scatter_resonance(P, ph, rng, cs, cnt, uz_s, [&](double xa, double ux, double uy, double uz) {
          for (int i = 0; i < P.nobs; ++i) {
            PeelRay pr;
            bool ok = P.use_stokes ? peel_resonance_stokes_prepare(P, P.obs[i], i, ph, cs, xa, ux, uy, uz, pr)
                                   : peel_resonance_nostokes_prepare(P, P.obs[i], i, ph, cs, xa, ux, uy, uz, pr);
            if (ok) trace_and_deposit(pr);
          }
        });
      }
      if (P.max_events > 0 && (ph.flags & PH_ALIVE) && ++nev >= P.max_events) {  // bounded run: abandon, record as it stands
        ph.flags &= ~PH_ALIVE;
        ph.xfreq_ref = ph.xfreq;
        retire_photon(P, ph, false, job, cnt);
      }
    }
    if (touched) {
      int fl = ph.flags;
      if (fl & PH_ALIVE) store_rng(pl, s, rng, fl);
      ph.flags = fl;
      store_all(pl, s, ph);
      if (!PLAIN && P.x.shear) pl.f[(size_t)F_SHEAR * pl.S + s] = ph.vshear;
      pl.nev[s] = nev;
    }
    if (rng_valid) nrng += rng.nrng;
  }
  flush_counters(P, cnt, nrng);
}

// ------------------------------ clump medium --------------------------------
// One thread per photon slot, like k_mono, with the clump ray tracers (lart_clump.cuh) in place of the cell walk.
// The slot's clump index (photon%icell_clump) lives in the first `rc` column of the pool.  Scattering, peel-off
// weights and tallies are the Cartesian routines — as upstream, which swaps only the ray tracers, do_resonance and
// the frame conversions (setup.f90:806-860).
// The walk is a chain of dependent loads (CSR offsets -> clump list -> geometry record) with little arithmetic between
// them, so resident warps matter more than registers: measured on clump_sphere_fcov5 (scatterings/s) 222 registers
// 1.33e8, 128 1.89e8, 80 2.08e8, 64 2.31e8, 40 2.32e8, 32 registers (64 warps/SM, spills served by L1) 2.75e8.
__global__ void __launch_bounds__(kBlock, 8) k_mono_clump(const __grid_constant__ DevParams P, Pool pl, Job *job, int quantum) {
  __shared__ double vtab[kVoigtTabN];
  load_vtab(P, vtab);
  const DevClumps &C = P.cl;
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  Counters cnt;
  ctr_t nrng = 0;
  const bool valid = s < pl.n;  // lanes beyond the pool run the loop idle: every lane of a warp must reach the joins
  {
    Photon ph;
    Rng rng;
    int icl = 0;
    ph.flags = valid ? pl.flags[s] : 0;
    bool rng_valid = false, touched = false;
    int nev = 0;
    if (ph.flags & PH_ALIVE) {
      load_trace_part(pl, s, ph);
      load_rest(pl, s, ph);
      load_rng(P, pl, s, ph.id, ph.flags, rng);
      icl = pl.rc[s];
      nev = pl.nev[s];
      rng_valid = true;
    }
    // a slot the wavefront flight stage already flew to its next scattering point (driver switch in lart_gpu_run)
    bool at_scatter = (ph.flags & PH_ALIVE) && (ph.flags & PH_SCATTER);
    ph.flags &= ~PH_SCATTER;
    // peel-off toward every observer from the photon's current state; csp = grid record with the clump's bulk velocity
    auto trace_and_deposit = [&](PeelRay &pr) {
      int nc = 0, ncl = 0;
      // overlapping populations: the event walk (peel_raytrace_to_edge, peelingoff_rect.f90:894-898)
      double t = C.overlap ? clump_walk_edge_overlap(P, vtab, pr.x, pr.y, pr.z, pr.kx, pr.ky, pr.kz, pr.xfreq, kTauHugeClump, nc)
                           : clump_walk_edge(P, vtab, pr.x, pr.y, pr.z, pr.kx, pr.ky, pr.kz, pr.xfreq, icl, kTauHugeClump, nc, ncl);
      cnt.cellsteps += nc; cnt.peel += 1;
      peel_deposit(P, pr, t, __activemask());
    };
    // Every lane runs the same `quantum` iterations through the same three phases and the warp is re-joined between
    // them (lanes without work idle through).  ncu: 4.95 of 32 threads active per issued instruction — the joins alone
    // do not cure that (2.96e8 -> 2.99e8 scatterings/s): most scatterings happen deep inside an opaque clump (no CSR
    // cell at all), a few lanes per iteration cross hundreds of cells.  The cure is the wavefront split with a
    // per-lane-refilled walker stepping one CSR cell at a time.
    const unsigned FULLW = 0xffffffffu;
    for (int ev = 0; ev < quantum; ++ev) {
      CellData cs;
      // ---- phase 1: refill a dead slot from the job queue; direct peel-off of the new photon
      bool fresh = false;
      CellData csp0;
      if (valid && !(ph.flags & PH_ALIVE) && job->next < job->count) {
        unsigned long long j = atomicAdd(&job->next, 1ULL);
        if (j < job->count) {
          fresh = true;
          touched = true;
          ph.id = job->first_id + (long long)j * job->stride;
          if (rng_valid) nrng += rng.nrng;
          rng.start(P.seed, (unsigned long long)ph.id);
          rng_valid = true;
          nev = 0;
          generate_photon(P, ph, rng, cnt, cs);
          // the birth clump, before the direct peel (generate_photon.f90:325-332); with overlapping populations the photon
          // stays in the global frame and owns no clump until its first scattering
          icl = C.overlap ? 0 : clump_at_point(C, ph.x, ph.y, ph.z);
          csp0 = cs;
          if (icl > 0) {
            const ClumpPhys cp = load_clump(C, icl);
            ph.xfreq = DSUB(ph.xfreq, ulos_clump(P, cp, ph.kx, ph.ky, ph.kz));
            const double ratio = cp.Dfreq / C.Dfreq_ref;
            csp0.vfx = DMUL(cp.vx, ratio); csp0.vfy = DMUL(cp.vy, ratio); csp0.vfz = DMUL(cp.vz, ratio);
          }
          if (P.save_all_photons) record_initial(P, ph);
        }
      }
      __syncwarp(FULLW);
      if (fresh && P.save_peeloff) {
        for (int i = 0; i < P.nobs; ++i) {
          PeelRay pr;
          if (!peel_direct_prepare(P, P.obs[i], i, ph, csp0, pr, icl > 0)) continue;
          trace_and_deposit(pr);
        }
      }
      __syncwarp(FULLW);
      // ---- phase 2: optical depth of the next flight, flight
      const bool live = (ph.flags & PH_ALIVE) != 0;
      bool inside = false;
      if (live && at_scatter) {
        touched = true;
        inside = true;
        at_scatter = false;
      } else if (live) {
        touched = true;
        double tau;
        if (ph.flags & PH_FIRST) {  // forced first scattering: the uncapped edge walk (setup.f90:809)
          int ci, cj, ck, nc = 0, ncl = 0;
          clamp_cell_for_read(P, ph, ci, cj, ck);
          load_cell(P, ci, cj, ck, cs);
          double tau0 = C.overlap ? clump_walk_edge_overlap(P, vtab, ph.x, ph.y, ph.z, ph.kx, ph.ky, ph.kz, ph.xfreq, -1.0, nc)
                                  : clump_walk_edge(P, vtab, ph.x, ph.y, ph.z, ph.kx, ph.ky, ph.kz, ph.xfreq, icl, -1.0, nc, ncl);
          cnt.cellsteps += nc;
          tau = forced_first(P, ph, rng, cs, tau0);
        } else {
          tau = -log(rng.uniform());
        }
        int nc = 0;
        inside = C.overlap ? clump_walk_tau_overlap(P, vtab, ph, icl, tau, rng, nc) : clump_walk_tau(P, vtab, ph, icl, tau, nc);
        cnt.cellsteps += nc;
        if (!inside) {  // left the sphere (raytrace_clump.f90:137-146, 151-160, 182-192)
          ph.flags &= ~PH_ALIVE;
          ph.xfreq_ref = ph.xfreq;  // upstream leaves photon%xfreq_ref unset on this path; Jout is binned with xfreq
          retire_photon(P, ph, true, job, cnt);
        }
      }
      __syncwarp(FULLW);
      // ---- phase 3: scattering inside clump icl — scattering_car.f90:72-87
      if (live && inside) {
        cnt.scatter += 1;
        load_cell(P, ph.ic, ph.jc, ph.kc, cs);  // the box: no opacity, no bulk velocity, reference Doppler width
        const ClumpPhys cp = load_clump(C, icl);
        const double ratio = cp.Dfreq / C.Dfreq_ref, scale = C.Dfreq_ref / cp.Dfreq;
        CellData csp = cs;
        csp.vfx = DMUL(cp.vx, ratio); csp.vfy = DMUL(cp.vy, ratio); csp.vfz = DMUL(cp.vz, ratio);
        bool to_dust = false;
        if (P.dust) {
          double pd = cp.rhokapD / (cp.rhokap * voigt_seon2(vtab, DMUL(ph.xfreq, scale), cp.voigt_a) + cp.rhokapD);
          to_dust = rng.uniform() <= pd;
        }
        if (to_dust) {
          scatter_dust(P, ph, rng, cs, cnt, [&]() {
            for (int i = 0; i < P.nobs; ++i) {
              PeelRay pr;
              bool ok = P.use_stokes ? peel_dust_stokes_prepare(P, P.obs[i], i, ph, csp, pr)
                                     : peel_dust_nostokes_prepare(P, P.obs[i], i, ph, csp, pr);
              if (ok) trace_and_deposit(pr);
            }
          });
          if (!(ph.flags & PH_ALIVE)) retire_photon(P, ph, false, job, cnt);
        } else {  // do_resonance1_clump — line_clump_mod.f90:29-58
          // overlapping populations: scatter_resonance_clump_* (scattering_car.f90:897-945) — into the owner clump's frame,
          // scatter, back to the global frame along the new direction
          if (C.overlap) ph.xfreq = DSUB(ph.xfreq, ulos_clump(P, cp, ph.kx, ph.ky, ph.kz));
          const double scale_inv = 1.0 / scale;
          const double xloc = ph.xfreq * scale;
          const double uz_loc = rand_resonance_vz(rng, xloc, cp.voigt_a, cnt.reject);
          const double xfreq_atom = (xloc - uz_loc) * scale_inv, uz = uz_loc * scale_inv;
          CellData css = cs;
          css.Dfreq = cp.Dfreq;  // recoil uses the clump's Doppler width (scattering_car.f90:428-433)
          scatter_resonance_core<true>(P, ph, rng, css, cnt, uz, xfreq_atom, ratio, [&](double xa, double ux, double uy, double uzz) {
            for (int i = 0; i < P.nobs; ++i) {
              PeelRay pr;
              bool ok = P.use_stokes ? peel_resonance_stokes_prepare(P, P.obs[i], i, ph, csp, xa, ux, uy, uzz, pr)
                                     : peel_resonance_nostokes_prepare(P, P.obs[i], i, ph, csp, xa, ux, uy, uzz, pr);
              if (ok) trace_and_deposit(pr);
            }
          });
          if (C.overlap) ph.xfreq = DADD(ph.xfreq, ulos_clump(P, cp, ph.kx, ph.ky, ph.kz));
        }
        if (P.max_events > 0 && (ph.flags & PH_ALIVE) && ++nev >= P.max_events) {
          ph.flags &= ~PH_ALIVE;
          ph.xfreq_ref = ph.xfreq;
          retire_photon(P, ph, false, job, cnt);
        }
      }
      __syncwarp(FULLW);
    }
    if (touched) {
      int fl = ph.flags;
      if (fl & PH_ALIVE) store_rng(pl, s, rng, fl);
      ph.flags = fl;
      store_all(pl, s, ph);
      pl.rc[s] = icl;
      pl.nev[s] = nev;
    }
    if (rng_valid) nrng += rng.nrng;
  }
  flush_counters(P, cnt, nrng);
}

// ------------------------------ wavefront driver ----------------------------
// Work items are pool slots (trace, scatter) and ray-queue entries (peel).  Every partition owns a region of
// 2 * n * nobs entries of the ray array (a slot can emit a photon AND scatter it in one wave); the emit and
// scatter stages append to it with warp-aggregated atomics (ray_append), the peel stage reads what was appended:
// rays that were proved dead (peel bound) or never existed cost no write and no read.  (The clump driver keeps
// the round-1 layout: the ray of slot s toward observer k at rays[s*nobs + k], kind = -1 = none, direct rays behind.)
constexpr int kRefillMin = 8;  // refill a warp when at least this many lanes are idle

// warp-aggregated reservation from a converged point: lanes with `want` get
// consecutive indices (one atomic per warp)
__device__ __forceinline__ unsigned reserve(unsigned int *ctr, bool want) {
  unsigned m = __ballot_sync(0xffffffffu, want);
  int lane = threadIdx.x & 31;
  unsigned base = 0;
  if (m) {
    int lead = __ffs(m) - 1;
    if (lane == lead) base = atomicAdd(ctr, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, lead);
  }
  return base + __popc(m & ((1u << lane) - 1u));
}

// Append a peel ray to the partition's queue.  The lanes that reach this point together (whatever subset of the warp)
// reserve consecutive entries with ONE atomic; an entry is written only by its owner, the peel stage reads [0, n_direct).
__device__ __forceinline__ void ray_append(const DevParams &P, const Queues &q, const PeelRay &pr) {
  auto g = cooperative_groups::coalesced_threads();
  unsigned base = 0;
  if (g.thread_rank() == 0) base = atomicAdd(q.n_direct, (unsigned)g.size());
  const unsigned at = g.shfl(base, 0) + g.thread_rank();
  if (at < q.direct_cap) ray_store(q.rays + q.direct_base + at, pr);
  else atomicOr(P.err, (unsigned)ERR_DIRECT_QUEUE);  // cannot happen by sizing; never drop a ray silently
}

// stage 1: refill dead slots from the job queue
__global__ void __launch_bounds__(kBlock) k_wf_emit(const __grid_constant__ DevParams P, Pool pl, Job *job, Queues q) {
  // nothing to do in most waves of an optically thick run (no photon died, or no photon is left to start): skip the scan.
  // The decision is taken once per block (other blocks change both words while this one reads them, and a block whose warps
  // disagreed would sum work counters of warps that never wrote them)
  __shared__ int go;
  if (threadIdx.x == 0) go = !(*q.n_dead == 0u || job->next >= job->count);
  __syncthreads();
  if (!go) return;
  Counters cnt;
  ctr_t nrng = 0;
  for (int s = pl.s0 + blockIdx.x * blockDim.x + threadIdx.x; s < pl.s0 + pl.n; s += gridDim.x * blockDim.x) {
    if (pl.flags[s] & PH_ALIVE) continue;
    if (job->next >= job->count) continue;
    unsigned long long j = atomicAdd(&job->next, 1ULL);
    if (j >= job->count) continue;
    {
      auto cg = cooperative_groups::coalesced_threads();
      if (cg.thread_rank() == 0) atomicSub(q.n_dead, (unsigned)cg.size());
    }
    Photon ph;
    Rng rng;
    CellData cs;
    ph.id = job->first_id + (long long)j * job->stride;
    rng.start(P.seed, (unsigned long long)ph.id);
    generate_photon(P, ph, rng, cnt, cs);
    if (P.save_all_photons) record_initial(P, ph);
    if (P.save_peeloff) {
      for (int i = 0; i < P.nobs; ++i) {
        PeelRay pr;
        if (!peel_direct_prepare(P, P.obs[i], i, ph, cs, pr)) continue;
        ray_append(P, q, pr);
      }
    }
    int fl = ph.flags;
    store_rng(pl, s, rng, fl);
    ph.flags = fl;
    store_all(pl, s, ph);
    if (P.x.shear) pl.f[(size_t)F_SHEAR * pl.S + s] = 0.0;
    pl.nev[s] = 0;
    nrng += rng.nrng;
  }
  flush_counters(P, cnt, nrng);
}

// the open-box walker instantiation applies (DevParams as created; clump runs have their own stages)
inline bool is_plain(const DevParams &P) { return !P.amr.on && !P.bcxy && !P.bcz && !P.sym && !P.x.any && !P.clump; }
// stage 2: raytrace_to_tau for every live photon, per-lane refill
#ifndef LART_TRACE_MINBLOCKS
#define LART_TRACE_MINBLOCKS 2
#endif
// PLAIN: see lart_device.cuh (ray_ulos) — the open-box instantiation of the walker
template <bool PLAIN>
__global__ void __launch_bounds__(kBlock, LART_TRACE_MINBLOCKS) k_wf_trace(const __grid_constant__ DevParams P, Pool pl, Job *job, Queues q, int budget) {
  __shared__ double vtab[kVoigtTabN];
  load_vtab(P, vtab);
  const unsigned FULL = 0xffffffffu;
  const size_t S = pl.S;
  Counters cnt;
  ctr_t nrng = 0;
  Ray r;
  Photon ph;
  Rng rng;
  double tau_in = 0.0;
  int slot = -1, mode = 0;  // mode 0: forced-first edge walk, 1: tau walk
  bool have = false, exhausted = false;
  CellData cs0;  // start cell of a first-flight photon
  for (;;) {
    // ---- refill idle lanes (converged point)
    bool need = !have && !exhausted;
    unsigned nm = __ballot_sync(FULL, need), hm = __ballot_sync(FULL, have);
    if (nm && (__popc(nm) >= kRefillMin || !hm)) {
      unsigned idx = reserve(q.head_trace, need);
      if (need) {
        if (idx >= (unsigned)pl.n) exhausted = true;
        else if ((pl.flags[pl.s0 + idx] & (PH_ALIVE | PH_SCATTER)) == PH_ALIVE) {  // alive and not already at a scattering point
          slot = pl.s0 + (int)idx;
          load_trace_part(pl, slot, ph);
          load_rng(P, pl, slot, ph.id, ph.flags, rng);
          if (ph.flags & PH_FIRST) {
            int ci, cj, ck;
            clamp_cell_for_read(P, ph, ci, cj, ck);
            load_cell<PLAIN>(P, ci, cj, ck, cs0);
          }
          if (ph.flags & PH_INFLIGHT) {  // a walk suspended at its step budget in an earlier wave: resume it exactly
            const double *st = pl.rs + slot;
            ph.flags &= ~PH_INFLIGHT;
            mode = (ph.flags & PH_FIRST) ? 0 : 1;
            tau_in = pl.f[(size_t)F_TAU * S + slot];
            ray_resume<PLAIN>(P, r, ph.x, ph.y, ph.z, ph.kx, ph.ky, ph.kz, st[0 * S], st[1 * S], st[2 * S], st[3 * S], st[4 * S],
                       st[5 * S], st[6 * S], st[7 * S], st[8 * S], st[9 * S], pl.rc[slot], pl.rc[S + slot], pl.rc[2 * S + slot],
                       !PLAIN && P.x.edge_open && mode == 0);
            if (!PLAIN && P.x.shear) r.vshear = pl.f[(size_t)F_SHEAR * S + slot];
            have = true;
          } else {
            bool leaving;
            if (ph.flags & PH_FIRST) {
              mode = 0;
              leaving = ray_setup<PLAIN>(P, r, ph.x, ph.y, ph.z, ph.kx, ph.ky, ph.kz, ph.ic, ph.jc, ph.kc, ph.xfreq, false);
              if (leaving) {  // tau0 = 0 (raytrace_to_edge returns at once)
                tau_in = forced_first(P, ph, rng, cs0, 0.0);
                mode = 1;
              }
            } else {
              if (ph.flags & PH_TAUPEND) {  // tau was drawn by the scatter stage, whose local step left the cell
                tau_in = pl.f[(size_t)F_TAU * S + slot];
                ph.flags &= ~PH_TAUPEND;
              } else {
                tau_in = -log(rng.uniform());
              }
              mode = 1;
              leaving = true;
            }
            if (!PLAIN && P.x.shear) r.vshear = pl.f[(size_t)F_SHEAR * S + slot];
            if (mode == 1 && leaving)
              leaving = ray_setup<PLAIN>(P, r, ph.x, ph.y, ph.z, ph.kx, ph.ky, ph.kz, ph.ic, ph.jc, ph.kc, ph.xfreq, true);
            if (leaving) {  // dead without tally (raytrace_car.f90:1469-1472)
              ph.flags &= ~PH_ALIVE;
              load_rest(pl, slot, ph);
              { retire_photon(P, ph, false, job, cnt); atomicAdd(q.n_dead, 1u); }
              pl.flags[slot] = ph.flags;
              nrng += rng.nrng;
            } else {
              have = true;
            }
          }
        }
      }
    }
    if (!__any_sync(FULL, have)) {
      if (!__any_sync(FULL, !exhausted)) break;
      continue;
    }
    // ---- one cell step for every lane that has a ray
    if (have) {
      if (mode == 0) {
        if (edge_step<PLAIN>(P, vtab, r)) {
          cnt.cellsteps += r.nsteps;
          tau_in = forced_first(P, ph, rng, cs0, r.tau);
          mode = 1;
          if (ray_setup<PLAIN>(P, r, ph.x, ph.y, ph.z, ph.kx, ph.ky, ph.kz, ph.ic, ph.jc, ph.kc, ph.xfreq, true)) {
            ph.flags &= ~PH_ALIVE;
            load_rest(pl, slot, ph);
            { retire_photon(P, ph, false, job, cnt); atomicAdd(q.n_dead, 1u); }
            store_trace_part(pl, slot, ph);
            nrng += rng.nrng;
            have = false;
          }
        }
      } else {
        double xp, yp, zp;
        int st = tau_step<PLAIN>(P, vtab, r, tau_in, xp, yp, zp, ph.wgt);
        if (st == 1) {
          ph.x = xp; ph.y = yp; ph.z = zp; ph.xfreq = r.xfreq;
          if (!P.zonly) { ph.ic = r.ic; ph.jc = r.jc; }
          ph.kc = r.kc;
          ph.flags |= PH_SCATTER;
          cnt.cellsteps += r.nsteps;
          if (!PLAIN && P.bcxy == BC_MIRROR && r.flip) { adopt_direction(r, ph); store_direction(pl, slot, ph); }
          if (!PLAIN && P.x.shear) pl.f[(size_t)F_SHEAR * S + slot] = r.vshear;
          store_trace_part(pl, slot, ph);
          pl.ndraw[slot] = rng.nblk;
          nrng += rng.nrng;
          have = false;
        } else if (st >= 2) {
          load_rest(pl, slot, ph);  // before finish_escape: it writes ph.xfreq_ref
          cnt.cellsteps += finish_escape<PLAIN>(P, ph, r, st == 3);
          { retire_photon(P, ph, true, job, cnt); atomicAdd(q.n_dead, 1u); }
          store_trace_part(pl, slot, ph);
          nrng += rng.nrng;
          have = false;
        }
      }
      // ---- step budget: a long walk is parked in the pool and resumed next wave, so that no wave waits for it
      if (have && r.nsteps >= budget) {
        double st10[10];
        int c3[3];
        ray_save_state(r, st10, c3);
        double *st = pl.rs + slot;
#pragma unroll
        for (int k = 0; k < 10; ++k) st[(size_t)k * S] = st10[k];
        pl.rc[slot] = c3[0]; pl.rc[S + slot] = c3[1]; pl.rc[2 * S + slot] = c3[2];
        pl.f[(size_t)F_TAU * S + slot] = tau_in;
        if (!PLAIN && P.x.shear) pl.f[(size_t)F_SHEAR * S + slot] = r.vshear;
        ph.flags |= PH_INFLIGHT;
        store_trace_part(pl, slot, ph);  // weight and flags may have changed (forced first scattering)
        pl.ndraw[slot] = rng.nblk;
        cnt.cellsteps += r.nsteps;
        nrng += rng.nrng;
        have = false;
      }
    }
  }
  flush_counters(P, cnt, nrng);
}

// stage 3: scattering for every photon flagged by the trace stage, in two kernels:
//
//   k_wf_draw   every random variate of the event (atom velocity, scattering angles), written to six pool columns.
//               The Philox rounds and the libm chains of the rejection samplers are long dependent instruction
//               sequences; as one kernel with the rest of the scattering (128 registers, 16 warps per SM) the stage
//               issued on 45 % of the cycles with "wait" and "no instruction" as top stalls (profiles/r2_*): it is
//               latency bound.  Alone, the samplers need few registers, so twice as many warps are resident.  All
//               three rejection loops are warp-cooperative (lart_device.cuh).
//   k_wf_apply  the deterministic rest: new frequency, peel-ray descriptors (one per observer), Stokes vector and
//               triad; dust events (rare) are scattered here serially.  Compile-time variants <STOKES, DUST, LOCAL>:
//               the instantiation a run uses carries only its own physics.  Only the columns a scattering changes are
//               written back (position, cell, weight and id stay as they are).
//
// A peel ray whose first cell alone is deeper than the cap ends there with a contribution of exactly zero
// (raytrace_car.f90:432,497: tau >= 745.2).  The path inside the cell is at least the distance L to the nearest face
// (|k| <= 1), and rounding is monotonic, so kappa*L >= 745.2 proves it without the DDA set-up and its three divides:
// such a ray is counted (one peel ray, one cell step, as the walk would; reported as n_peel_bound) and never written
// to the queue — the test runs as soon as the ray's frequency is known, before the Stokes algebra.  On a face L = 0.
#ifndef LART_DRAW_MINBLOCKS
#define LART_DRAW_MINBLOCKS 3
#endif
#ifndef LART_APPLY_MINBLOCKS
#define LART_APPLY_MINBLOCKS 2
#endif
#ifndef LART_PEEL_BOUND
#define LART_PEEL_BOUND 1
#endif

__device__ __forceinline__ bool peel_certainly_capped(const DevParams &P, const double *vtab, const CellData &cs, const PeelRay &pr) {
  if (P.amr.on) {  // the same bound with the leaf's faces: L = min over the axes of h - |p - c|
    if (pr.ic <= 0) return false;
    const double4 g = ldg_d4(P.amr.geo + (pr.ic - 1));
    const double La = fmin(DSUB(g.w, fabs(DSUB(pr.x, g.x))), fmin(DSUB(g.w, fabs(DSUB(pr.y, g.y))), DSUB(g.w, fabs(DSUB(pr.z, g.z)))));
    if (!(La > 0.0)) return false;
    if (!(DMUL(DADD(cs.rhokap, P.dust ? cs.rhokapD : 0.0), La) >= kTauHuge)) return false;
    double kapa = DMUL(cs.rhokap, voigt_seon2(vtab, pr.xfreq, cs.voigt_a));
    if (P.dust) kapa = DADD(kapa, cs.rhokapD);
    return DMUL(kapa, La) >= kTauHuge;
  }
  if (P.x.edge_open) return false;  // shearing boxes / atmospheres: the edge walk is not the one this bound reasons about
  double L = fmin(DSUB(pr.z, __ldg(P.zface + pr.kc - 1)), DSUB(__ldg(P.zface + pr.kc), pr.z));
  if (!P.zonly) {
    L = fmin(L, fmin(DSUB(pr.x, __ldg(P.xface + pr.ic - 1)), DSUB(__ldg(P.xface + pr.ic), pr.x)));
    L = fmin(L, fmin(DSUB(pr.y, __ldg(P.yface + pr.jc - 1)), DSUB(__ldg(P.yface + pr.jc), pr.y)));
  }
  if (!(L > 0.0)) return false;
  // H(x,a) <= H(0,a) < 1: a cell that is too thin even at line centre cannot cap any ray (no Voigt evaluation then:
  // on the tau0 = 1e4 sphere, 100 per cell, the unconditional test cost 9 %)
  if (!(DMUL(DADD(cs.rhokap, P.dust ? cs.rhokapD : 0.0), L) >= kTauHuge)) return false;
  double kap = DMUL(cs.rhokap, voigt_seon2(vtab, pr.xfreq, cs.voigt_a));
  if (P.dust) kap = DADD(kap, cs.rhokapD);
  return DMUL(kap, L) >= kTauHuge;
}

// SERIAL (LART_FLAG_SERIAL_REJECTION): per-lane rejection loops, the ablation arm of the cooperative samplers.
template <bool STOKES, bool DUST, bool SERIAL>
__global__ void __launch_bounds__(kBlock, LART_DRAW_MINBLOCKS) k_wf_draw(const __grid_constant__ DevParams P, Pool pl) {
  __shared__ double vtab[DUST ? kVoigtTabN : 1];
  __shared__ VzWarpShared vzsh[kBlock / 32];
  if (DUST) load_vtab(P, vtab);
  VzWarpShared &sh = vzsh[threadIdx.x >> 5];
  Counters cnt;
  ctr_t nrng = 0;
  const int lane = threadIdx.x & 31;
  const size_t S = pl.S;
  const int end = pl.s0 + pl.n, stride = gridDim.x * blockDim.x;
  for (int base = pl.s0 + blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < end; base += stride) {  // warp-uniform
    const int s = base + lane;
    const int fl0 = s < end ? pl.flags[s] : 0;
    const bool active = (fl0 & PH_SCATTER) != 0;
    if (!__any_sync(0xffffffffu, active)) continue;
    Rng rng;
    double x = 0.0, a = 1.0, Q = 0.0, U = 0.0, xc = 0.0, xc2 = 0.0;
    int fl = fl0;
    bool to_dust = false;
    if (active) {
      const double *f = pl.f + s;
      x = f[F_XFREQ * S];
      load_rng(P, pl, s, pl.id[s], fl0, rng);
      if (STOKES) { Q = f[F_Q * S]; U = f[F_U * S]; }
      const int ic = pl.ic[s], jc = pl.jc[s], kc = pl.kc[s];
      double rhokap = 0.0;
      if (DUST || P.soa) {
        CellData cs;
        load_cell(P, ic, jc, kc, cs);
        a = cs.voigt_a; rhokap = cs.rhokap;
        if (DUST) {  // scattering_car.f90:106-118
          const double pd = cs.rhokapD / (cs.rhokap * voigt_seon2(vtab, x, cs.voigt_a) + cs.rhokapD);
          to_dust = rng.uniform() <= pd;
        }
      } else {
        const double2 c0 = __ldg(reinterpret_cast<const double2 *>(P.cells + cell_slot(P, ic, jc, kc)));
        rhokap = c0.x; a = c0.y;
      }
      if (P.core_skip && !to_dust) car_xcrit_local(P, ic, jc, kc, f[F_X * S], f[F_Y * S], f[F_Z * S], a, rhokap, xc, xc2);
    } else {
      rng.start(P.seed, 0ULL);
    }
    const bool resonant = active && !to_dust;
    ScatterVariates v;
    if (SERIAL) {
      if (resonant) {  // the statements of scatter_resonance_core, draws only
        v.uz = rand_resonance_vz(rng, x, a, cnt.reject);
        v.cost = rand_resonance_fast(rng, P);
        const double cost2 = v.cost * v.cost, S22 = 0.75 * P.E1 * (cost2 + 1.0);
        if (STOKES) {
          Photon tmp;
          tmp.Q = Q; tmp.U = U;
          sample_phi_stokes(rng, tmp, 0.75 * P.E1 * (cost2 - 1.0) / (S22 + P.E2), cnt.reject, v.cosp, v.sinp);
        } else {
          sincospi(2.0 * rng.uniform(), &v.sinp, &v.cosp);
        }
        const bool skip = P.core_skip && fabs(x) < xc;
        if (STOKES && !skip) {
          const double one_over_sqrt2 = 1.0 / 1.4142135623730951;
          v.ux = rng.gauss(cnt.reject) * one_over_sqrt2;
          v.uy = rng.gauss(cnt.reject) * one_over_sqrt2;
        } else {
          double u1, u2;
          rng.uniform2(u1, u2);
          const double uxy = skip ? sqrt(xc2 - log(u2)) : sqrt(-log(u2));
          double s2, c2;
          sincospi(2.0 * u1, &s2, &c2);
          v.ux = uxy * c2; v.uy = uxy * s2;
        }
      }
    } else {
      draw_resonance_warp<STOKES>(sh, resonant, P, rng, x, a, Q, U, xc, xc2, cnt, v);
    }
    if (!active) continue;
    if (resonant) {
      double *o = pl.var + s;
      o[0 * S] = v.uz; o[1 * S] = v.cost; o[2 * S] = v.cosp; o[3 * S] = v.sinp; o[4 * S] = v.ux; o[5 * S] = v.uy;
    } else {
      fl |= PH_DUSTEV;  // k_wf_apply runs scatter_dust from here on the same stream
    }
    store_rng(pl, s, rng, fl);
    if (fl != fl0) pl.flags[s] = fl;
    nrng += rng.nrng;
  }
  flush_counters(P, cnt, nrng);
}

// ---- k_wf_draw2 (default): the same variates with the rejection loops COMPACTED instead of shared -----------------------
// ncu on k_wf_draw (profiles/r2_ncu_draw_split.txt): 3670 warp instructions per scattering, of which the wing sampler takes 910
// although only one photon in six is in the wings (its set-up runs with 5 of 32 lanes, its trial rounds evaluate six speculative
// trials per pending photon), and 9 Philox blocks are computed where the serial algorithm consumes 5.9.  Speculative trials buy
// latency with instructions, and this stage is issue bound (56 % of the issue slots, top stall: instruction fetch).
// Here every warp owns a contiguous CHUNK of pool slots and takes it through the event pass by pass; a pass is either straight-line
// code over the chunk (32 photons at a time) or a rejection loop with per-lane refill: a lane whose trial was accepted stores its
// result and takes the next photon of the pass's list, so every issued trial is a needed one and (almost) every lane runs one.
//   pass 0   classify: dust coin (DUST), core-skip threshold, lists of line-core (|x| <= 1) and wing photons of the chunk
//   pass A   core photons: Lorentzian proposal + exp(-u^2) acceptance                      (refill loop)
//   pass B   wing photons: majorant table (straight) / trials (refill loop) / u = x + a tan(..) (straight)
//   pass C   cos(theta) (rand_resonance); Stokes: S12/S11 and the azimuth envelope; no Stokes: the azimuth
//   pass D   Stokes: azimuth rejection                                                     (refill loop)
//   pass E   Stokes, no core skip: Marsaglia polar pair                                    (refill loop)
//   pass F   perpendicular atom velocity from the pair (or the core-skip / no-Stokes form)
// Each photon's trials are evaluated one after the other in the order of its own Philox stream, so the variates are those of the
// serial loops bit for bit (and the oracle's).  Intermediate values travel through the pool's scratch columns (var, wtab, lst):
// a chunk's few hundred slots stay in L1/L2.
#ifndef LART_DRAW2_MINBLOCKS
#define LART_DRAW2_MINBLOCKS 4
#endif

// Rejection loop over `count` work items owned by this warp.  load(i) sets the lane up for item i (false: nothing to do for it);
// trial() runs one trial of the lane's item and returns true when the item is finished (result stored).
#ifndef LART_DRAW_REFILL
#define LART_DRAW_REFILL kRefillMin
#endif
#ifndef LART_CORE_SQUEEZE
#define LART_CORE_SQUEEZE 0
#endif
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
struct NoPrefetch { __device__ __forceinline__ void operator()(int) const {} };
template <class Load, class Trial, class Pre = NoPrefetch>
__device__ __forceinline__ void warp_refill_loop(int count, Load &&load, Trial &&trial, Pre &&pre = Pre()) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  int cursor = 0;  // warp-uniform
  bool have = false;
  for (;;) {
    const unsigned nm = __ballot_sync(FULL, !have);
    if (cursor < count && (__popc(nm) >= LART_DRAW_REFILL || nm == FULL)) {
      const int idx = cursor + __popc(nm & lt);
      if (!have && idx < count) have = load(idx);
      cursor += __popc(nm);
      if (cursor + lane < count) pre(cursor + lane);  // the window the next refills will take their items from
    }
    if (!__any_sync(FULL, have)) {
      if (cursor >= count) break;
      continue;
    }
    if (have && trial()) have = false;
  }
}

template <bool STOKES, bool DUST>
__global__ void __launch_bounds__(kBlock, LART_DRAW2_MINBLOCKS) k_wf_draw2(const __grid_constant__ DevParams P, Pool pl, int chunk) {
  __shared__ double vtab[DUST ? kVoigtTabN : 1];
  if (DUST) load_vtab(P, vtab);
  const unsigned FULL = 0xffffffffu;
  Counters cnt;
  ctr_t nrng = 0;
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const size_t S = pl.S;
  const long long w = (long long)blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
  const int end = pl.s0 + pl.n;
  const int c0 = (int)min((long long)end, pl.s0 + w * chunk), c1 = min(end, c0 + chunk);
  double *const var0 = pl.var, *const var1 = pl.var + S, *const var2 = pl.var + 2 * S, *const var3 = pl.var + 3 * S,
               *const var4 = pl.var + 4 * S, *const var5 = pl.var + 5 * S;
  double *const w_xc = pl.wtab + 10 * S, *const w_rsq = pl.wtab + 11 * S;
  const double *const f = pl.f;
  const unsigned long long seed = P.seed;
  int nB = 0;  // wing photons of the chunk (warp-uniform)
  const int nC = c1 - c0;
  auto resonant = [](int fl) { return (fl & (PH_SCATTER | PH_DUSTEV)) == PH_SCATTER; };
  // ---------------- pass 0: classify
  for (int base = c0; base < c1; base += 32) {
    const int s = base + lane;
    const int fl0 = s < c1 ? pl.flags[s] : 0;
    const bool active = (fl0 & PH_SCATTER) != 0;
    if (!__any_sync(FULL, active)) continue;
    double x = 0.0;
    bool to_dust = false;
    if (active) {
      x = f[F_XFREQ * S + s];
      const int ic = pl.ic[s], jc = pl.jc[s], kc = pl.kc[s];
      double a, rhokap;
      if (DUST || P.soa) {
        CellData cs;
        load_cell(P, ic, jc, kc, cs);
        a = cs.voigt_a; rhokap = cs.rhokap;
        if (DUST) {  // scattering_car.f90:106-118
          const double pd = cs.rhokapD / (cs.rhokap * voigt_seon2(vtab, x, cs.voigt_a) + cs.rhokapD);
          double u, dummy;
          const unsigned long long nb = pl.ndraw[s];
          philox_uniform2(seed, (unsigned long long)pl.id[s], nb, u, dummy);
          pl.ndraw[s] = nb + 1;
          nrng += 1;
          to_dust = u <= pd;
          if (to_dust) pl.flags[s] = fl0 | PH_DUSTEV;  // k_wf_apply runs scatter_dust from here on the same stream
        }
      } else {
        const double2 cc = __ldg(reinterpret_cast<const double2 *>(P.cells + cell_slot(P, ic, jc, kc)));
        rhokap = cc.x; a = cc.y;
      }
      if (!to_dust) {
        var1[s] = a;
        if (P.core_skip) {
          double xc, xc2;
          car_xcrit_local(P, ic, jc, kc, f[F_X * S + s], f[F_Y * S + s], f[F_Z * S + s], a, rhokap, xc, xc2);
          w_xc[s] = xc;
        }
      }
    }
    const bool wing = active && !to_dust && !(fabs(x) <= 1.0);
    const unsigned mB = __ballot_sync(FULL, wing);
    if (wing) pl.lst[c1 - 1 - (nB + __popc(mB & lt))] = s;
    nB += __popc(mB);
  }
  __syncwarp();
  // ---------------- pass A: |x| <= 1 (random_mt.f90:2579-2585)
  {
    int slot = 0, ntr = 0;
    double x = 0.0, x0 = 0.0, a = 1.0;
    unsigned long long id = 0, nb = 0;
    warp_refill_loop(nC,
      [&](int i) {
        slot = c0 + i;
        const int fl = pl.flags[slot];
        x = f[F_XFREQ * S + slot]; a = var1[slot]; id = (unsigned long long)pl.id[slot]; nb = pl.ndraw[slot];
        x0 = fabs(x);
        return resonant(fl) && x0 <= 1.0;
      },
      [&]() {
        double u1, u2;
        philox_uniform2(seed, id, nb, u1, u2);
        ++nb; ++ntr;
        const double v = x0 + a * tan(kPi * (u1 - 0.5));
#if LART_CORE_SQUEEZE
        // squeeze around exp(-v^2): 1 - t <= exp(-t) <= 1/(1 + t) for t >= 0 decides three trials in four without the exp
        // (the bounds differ from exp by >= t^2/2; a decision could only flip against the exact test for |v| < 2e-4 AND u2
        // within 4e-16 of the bound — never in practice, and far below the CUDA-vs-glibc exp difference the parity tests live with)
        const double v2 = v * v;
        if (u2 * (1.0 + v2) > 1.0) return false;
        if (!(u2 <= 1.0 - v2) && !(u2 <= exp(-v2))) return false;
#else
        if (!(u2 <= exp(-v * v))) return false;
#endif
        var0[slot] = (x < 0.0) ? -v : v;
        pl.ndraw[slot] = nb;
        return true;
      },
      [&](int i) { const int sp = c0 + i; prefetch_l1(pl.flags + sp); prefetch_l1(f + F_XFREQ * S + sp); prefetch_l1(var1 + sp); prefetch_l1(pl.id + sp); prefetch_l1(pl.ndraw + sp); });
    cnt.reject += ntr; nrng += 2 * ntr;
  }
  // ---------------- pass B: wings, piecewise-constant majorant in beta (:2605-2690)
  if (nB > 0) {
    for (int i0 = 0; i0 < nB; i0 += 32) {  // B0: the majorant table of every wing photon (the expressions of rand_resonance_vz)
      const int i = i0 + lane;
      if (i >= nB) continue;
      const int pos = c1 - 1 - i, slot = pl.lst[pos];
      const double x0 = fabs(f[F_XFREQ * S + slot]), a = var1[slot];
      const double xc = 1.0 + 1.4142135623730951, two_over_PI = 2.0 / kPi;
      double T[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      int mode;
      const double x0sq = x0 * x0;
      const double beta0 = exp(-x0sq / 2.0);
      const double h0_two = beta0 / a, h0 = h0_two / 2.0;
      const double h2 = 0.3861 / (x0sq - 1.373);
      T[0] = beta0;
      if (x0 < xc || !(h0 < h2)) {
        const double dbeta = sqrt(two_over_PI * a * (1.0 - beta0) * beta0 * x0);
        const double beta1 = beta0 + dbeta, one_b1 = 1.0 - beta1;
        const double pb1 = sqrt(-2.0 * log(beta1));
        const double h1 = two_over_PI * beta1 * pb1 / (x0sq - pb1 * pb1);
        const double hm = (x0 < xc) ? h1 : ((h1 > h2) ? h1 : h2);
        const double S0 = beta0 * h0, S1 = dbeta * h0, S2 = one_b1 * hm, Stot = S0 + S1 + S2;
        mode = 2; T[1] = S0 / Stot; T[2] = 1.0 - S2 / Stot;
        T[3] = beta0; T[4] = dbeta; T[5] = h0; T[6] = beta1; T[7] = one_b1; T[8] = hm;
      } else if (h0_two < h2) {
        mode = 0; T[3] = 0.0; T[4] = 1.0; T[5] = h2;
      } else {
        const double S0 = beta0 * h0, one_b0 = 1.0 - beta0, S1 = one_b0 * h2, Stot = S0 + S1;
        mode = 1; T[1] = S0 / Stot; T[3] = beta0; T[4] = one_b0; T[5] = h2;
      }
#pragma unroll
      for (int q = 0; q < 9; ++q) pl.wtab[(size_t)q * S + pos] = T[q];
      pl.wtab[(size_t)9 * S + pos] = (double)mode;
    }
    __syncwarp();
    {  // B1: trials
      int slot = 0, ntr = 0, mode = 0, nu = 0;
      double x0 = 0.0, a = 1.0, T[9];
      unsigned long long id = 0, nb = 0;
      warp_refill_loop(nB,
        [&](int i) {
          const int pos = c1 - 1 - i;
          slot = pl.lst[pos]; x0 = fabs(f[F_XFREQ * S + slot]); a = var1[slot]; id = (unsigned long long)pl.id[slot]; nb = pl.ndraw[slot];
#pragma unroll
          for (int q = 0; q < 9; ++q) T[q] = pl.wtab[(size_t)q * S + pos];
          mode = (int)pl.wtab[(size_t)9 * S + pos];
          return true;
        },
        [&]() {
          double ua, ub, uacc, beta, Cb;
          philox_uniform2(seed, id, nb, ua, ub);
          ++nb; ++ntr;
          if (mode == 0) { beta = T[3] + T[4] * ua; Cb = T[5]; uacc = ub; nu += 2; }
          else {
            if (ua < T[1]) { beta = T[0] * sqrt(ub); Cb = beta / a; }
            else if (mode == 1 || ua < T[2]) { beta = T[3] + T[4] * ub; Cb = T[5]; }
            else { beta = T[6] + T[7] * ub; Cb = T[8]; }
            double dummy;
            philox_uniform2(seed, id, nb, uacc, dummy);
            ++nb; nu += 3;
          }
          const double pb = sqrt(-2.0 * log(beta));
          const double t2 = atan((pb - x0) / a), t1 = atan((-pb - x0) / a), delt = t2 - t1;
          if (!(uacc * Cb < (beta / (a * kPi)) * delt)) return false;
          var0[slot] = t1; var2[slot] = delt;
          pl.ndraw[slot] = nb;
          return true;
        });
      cnt.reject += ntr; nrng += nu;
    }
    __syncwarp();
    for (int i0 = 0; i0 < nB; i0 += 32) {  // B2: u = x + a tan(delt xi + t1)  (:2693)
      const int i = i0 + lane;
      if (i >= nB) continue;
      const int slot = pl.lst[c1 - 1 - i];
      const double x = f[F_XFREQ * S + slot], a = var1[slot];
      const unsigned long long nb = pl.ndraw[slot];
      double u, dummy;
      philox_uniform2(seed, (unsigned long long)pl.id[slot], nb, u, dummy);
      const double vz = fabs(x) + a * tan(var2[slot] * u + var0[slot]);
      var0[slot] = (x < 0.0) ? -vz : vz;
      pl.ndraw[slot] = nb + 1;
      nrng += 1;
    }
  }
  __syncwarp();
  // ---------------- pass C: cos(theta) (rand_resonance, random_mt.f90:2974-2993)
  for (int base = c0; base < c1; base += 32) {
    const int s = base + lane;
    const int fl = s < c1 ? pl.flags[s] : 0;
    if (!resonant(fl)) continue;
    Rng r;
    r.start(seed, (unsigned long long)pl.id[s], pl.ndraw[s]);
    const double cost = rand_resonance_fast(r, P);
    var1[s] = cost;
    if (STOKES) {
      const double cost2 = cost * cost, S22 = 0.75 * P.E1 * (cost2 + 1.0);
      const double S12overS11 = 0.75 * P.E1 * (cost2 - 1.0) / (S22 + P.E2);
      const double Q = f[F_Q * S + s], U = f[F_U * S + s];
      var2[s] = S12overS11;
      var3[s] = 1.0 + fabs(S12overS11) * sqrt(Q * Q + U * U);
    } else {
      double sinp, cosp;
      sincospi(2.0 * r.uniform(), &sinp, &cosp);
      var2[s] = cosp; var3[s] = sinp;
    }
    pl.ndraw[s] = r.nblk;
    nrng += r.nrng;
  }
  __syncwarp();
  if (STOKES) {
    // ---------------- pass D: azimuth by rejection (scattering_car.f90:364-371)
    {
      int slot = 0, ntr = 0;
      double Q = 0.0, U = 0.0, s12 = 0.0, env = 1.0;
      unsigned long long id = 0, nb = 0;
      warp_refill_loop(nC,
        [&](int i) {
          slot = c0 + i;
          const int fl = pl.flags[slot];
          Q = f[F_Q * S + slot]; U = f[F_U * S + slot]; s12 = var2[slot]; env = var3[slot]; id = (unsigned long long)pl.id[slot]; nb = pl.ndraw[slot];
          return resonant(fl);
        },
        [&]() {
          double u1, u2, sinp, cosp;
          philox_uniform2(seed, id, nb, u1, u2);
          ++nb; ++ntr;
          sincospi(2.0 * u1, &sinp, &cosp);
          const double c2 = 2.0 * cosp * cosp - 1.0, s2 = 2.0 * sinp * cosp;
          if (!(env * u2 <= 1.0 + s12 * (Q * c2 + U * s2))) return false;
          var2[slot] = cosp; var3[slot] = sinp;
          pl.ndraw[slot] = nb;
          return true;
        },
        [&](int i) { const int sp = c0 + i; prefetch_l1(pl.flags + sp); prefetch_l1(f + F_Q * S + sp); prefetch_l1(f + F_U * S + sp); prefetch_l1(var2 + sp); prefetch_l1(var3 + sp); prefetch_l1(pl.id + sp); prefetch_l1(pl.ndraw + sp); });
      cnt.reject += ntr; nrng += 2 * ntr;
    }
    __syncwarp();
    // ---------------- pass E: one Marsaglia polar pair per photon that is not core-skipped (:413-414; random_mt.f90:964-988)
    {
      int slot = 0, ntr = 0;
      unsigned long long id = 0, nb = 0;
      warp_refill_loop(nC,
        [&](int i) {
          slot = c0 + i;
          const int fl = pl.flags[slot];
          id = (unsigned long long)pl.id[slot]; nb = pl.ndraw[slot];
          if (P.core_skip && fabs(f[F_XFREQ * S + slot]) < w_xc[slot]) return false;
          return resonant(fl);
        },
        [&]() {
          double v1, v2;
          philox_uniform2(seed, id, nb, v1, v2);
          ++nb; ++ntr;
          v1 = 2.0 * v1 - 1.0; v2 = 2.0 * v2 - 1.0;
          const double rsq = v1 * v1 + v2 * v2;
          if (!(rsq > 0.0 && rsq < 1.0)) return false;
          var4[slot] = v1; var5[slot] = v2; w_rsq[slot] = rsq;
          pl.ndraw[slot] = nb;
          return true;
        },
        [&](int i) { const int sp = c0 + i; prefetch_l1(pl.flags + sp); prefetch_l1(pl.id + sp); prefetch_l1(pl.ndraw + sp); });
      cnt.reject += ntr; nrng += 2 * ntr;
    }
    __syncwarp();
  }
  // ---------------- pass F: perpendicular atom velocity
  for (int base = c0; base < c1; base += 32) {
    const int s = base + lane;
    const int fl = s < c1 ? pl.flags[s] : 0;
    if (!resonant(fl)) continue;
    double xc = 0.0;
    const bool skip = P.core_skip && fabs(f[F_XFREQ * S + s]) < (xc = w_xc[s]);
    double ux, uy;
    if (STOKES && !skip) {  // two rand_gauss calls = the stored spare (if any) and the pair of pass E
      const double one_over_sqrt2 = 1.0 / 1.4142135623730951;
      const double rsq = w_rsq[s];
      const double g = sqrt(-2.0 * log(rsq) / rsq), g_spare = var4[s] * g, g_first = var5[s] * g;
      if (fl & PH_GAUSS) { ux = pl.f[(size_t)F_GSET * S + s] * one_over_sqrt2; uy = g_first * one_over_sqrt2; pl.f[(size_t)F_GSET * S + s] = g_spare; }
      else { ux = g_first * one_over_sqrt2; uy = g_spare * one_over_sqrt2; }
    } else {  // :397-401 / :749-752
      const double xc2 = P.core_skip_global ? P.xcrit2 : xc * xc;
      double u1, u2, s2, c2;
      const unsigned long long nb = pl.ndraw[s];
      philox_uniform2(seed, (unsigned long long)pl.id[s], nb, u1, u2);
      pl.ndraw[s] = nb + 1;
      nrng += 2;
      const double uxy = skip ? sqrt(xc2 - log(u2)) : sqrt(-log(u2));
      sincospi(2.0 * u1, &s2, &c2);
      ux = uxy * c2; uy = uxy * s2;
    }
    var4[s] = ux; var5[s] = uy;
  }
  flush_counters(P, cnt, nrng);
}

template <bool STOKES, bool DUST, bool LOCAL>
__global__ void __launch_bounds__(kBlock, LART_APPLY_MINBLOCKS) k_wf_apply(const __grid_constant__ DevParams P, Pool pl, Job *job, Queues q) {
  __shared__ double vtab[kVoigtTabN];
  const bool bound = LART_PEEL_BOUND && !LOCAL && P.save_peeloff;
  if (LOCAL || DUST || bound) load_vtab(P, vtab);
  Counters cnt;
  ctr_t nrng = 0;
  const size_t S = pl.S;
  const int end = pl.s0 + pl.n;
  for (int s = pl.s0 + blockIdx.x * blockDim.x + threadIdx.x; s < end; s += gridDim.x * blockDim.x) {
    const int fl0 = pl.flags[s];
    if (!(fl0 & PH_SCATTER)) continue;
    Photon ph;
    CellData cs;
    const double *f = pl.f + s;
    ph.flags = fl0 & ~(PH_SCATTER | PH_DUSTEV);
    ph.id = pl.id[s];
    ph.ic = pl.ic[s]; ph.jc = pl.jc[s]; ph.kc = pl.kc[s];
    load_cell(P, ph.ic, ph.jc, ph.kc, cs);
    ph.x = f[F_X * S]; ph.y = f[F_Y * S]; ph.z = f[F_Z * S];
    ph.kx = f[F_KX * S]; ph.ky = f[F_KY * S]; ph.kz = f[F_KZ * S];
    ph.xfreq = f[F_XFREQ * S]; ph.wgt = f[F_WGT * S]; ph.nsg = f[F_NSG * S];
    ph.nsd = DUST ? f[F_NSD * S] : 0.0;
    ph.xfreq_ref = 0.0;
    if (STOKES) {
      ph.mx = f[F_MX * S]; ph.my = f[F_MY * S]; ph.mz = f[F_MZ * S];
      ph.nx = f[F_NX * S]; ph.ny = f[F_NY * S]; ph.nz = f[F_NZ * S];
      ph.Q = f[F_Q * S]; ph.U = f[F_U * S]; ph.V = f[F_V * S];
    } else {  // no triad and no Stokes vector in this variant (record_final reads them only with par%use_stokes)
      ph.mx = ph.my = ph.mz = ph.nx = ph.ny = ph.nz = 0.0; ph.Q = ph.U = ph.V = 0.0;
    }
    const double wgt_in = ph.wgt;
    cnt.scatter += 1;
    Rng rng;
    bool have_rng = false;
    bool peeled = false;
    // With local steps the ray toward observer 0 stays in registers: most of them end inside the
    // photon's own cell (tau cap) and never reach the queue.
    PeelRay pr0;
    bool have_pr0 = false;
    auto emit_ray = [&](int k, int code, const PeelRay &pr) {  // code: 0 = no ray, 1 = descriptor ready, 2 = proved dead
      if (LOCAL && k == 0) { have_pr0 = code == 1; if (have_pr0) pr0 = pr; }
      else if (code == 1) ray_append(P, q, pr);
      if (code == 2) { cnt.peel += 1; cnt.cellsteps += 1; cnt.peel_bound += 1; }
    };
    auto drop = [&](const PeelRay &pr) { return bound && peel_certainly_capped(P, vtab, cs, pr); };
    if (DUST && (fl0 & PH_DUSTEV)) {
      load_rng(P, pl, s, ph.id, ph.flags, rng);
      have_rng = true;
      scatter_dust(P, ph, rng, cs, cnt, [&]() {
        peeled = true;
        for (int k = 0; k < P.nobs; ++k) {
          PeelRay pr;
          bool ok = STOKES ? peel_dust_stokes_prepare(P, P.obs[k], k, ph, cs, pr)
                           : peel_dust_nostokes_prepare(P, P.obs[k], k, ph, cs, pr);
          emit_ray(k, (ok && drop(pr)) ? 2 : (ok ? 1 : 0), pr);
        }
      });
      if (!(ph.flags & PH_ALIVE)) { retire_photon(P, ph, false, job, cnt); atomicAdd(q.n_dead, 1u); }
    } else {
      const double *vi = pl.var + s;
      ScatterVariates v;
      v.uz = vi[0 * S]; v.cost = vi[1 * S]; v.cosp = vi[2 * S]; v.sinp = vi[3 * S]; v.ux = vi[4 * S]; v.uy = vi[5 * S];
      apply_resonance<STOKES>(P, ph, cs, v, [&](double xa, double ux, double uy, double uz) {
        peeled = true;
        for (int k = 0; k < P.nobs; ++k) {
          PeelRay pr;
          const int code = STOKES ? peel_resonance_stokes_prepare2(P, P.obs[k], k, ph, cs, xa, ux, uy, uz, pr, drop)
                                  : peel_resonance_nostokes_prepare2(P, P.obs[k], k, ph, cs, xa, ux, uy, uz, pr, drop);
          emit_ray(k, code, pr);
        }
      });
    }
    // ---- bounded runs: the photon is abandoned after max_events scatterings, recorded as it stands
    if (P.max_events > 0 && (ph.flags & PH_ALIVE)) {
      const int ne = pl.nev[s] + 1;
      pl.nev[s] = ne;
      if (ne >= P.max_events) {
        ph.flags &= ~PH_ALIVE;
        ph.xfreq_ref = ph.xfreq;
        { retire_photon(P, ph, false, job, cnt); atomicAdd(q.n_dead, 1u); }
      }
    }
    bool moved = false;  // position / cell changed (local steps only)
    if (LOCAL) {
      // ---- first cell step of the peel ray (raytrace_to_edge): in an optically thick cell it is the last one
      if (have_pr0) {
        Ray r;
        bool resolved = false;
        double tau = 0.0;
        if (ray_setup(P, r, pr0.x, pr0.y, pr0.z, pr0.kx, pr0.ky, pr0.kz, pr0.ic, pr0.jc, pr0.kc, pr0.xfreq, false, &cs)) resolved = true;
        else if (edge_step(P, vtab, r)) { resolved = true; tau = r.tau; cnt.cellsteps += r.nsteps; }
        if (resolved) { cnt.peel += 1; peel_deposit(P, pr0, tau, 0u, false); }
        else ray_append(P, q, pr0);  // the peel stage walks it (from its start)
      }
      // ---- first cell step of the next flight (raytrace_to_tau): most flights end inside the cell
      if (ph.flags & PH_ALIVE) {
        if (!have_rng) { load_rng(P, pl, s, ph.id, ph.flags, rng); have_rng = true; }
        const double tau_in = -log(rng.uniform());
        Ray r;
        if (ray_setup(P, r, ph.x, ph.y, ph.z, ph.kx, ph.ky, ph.kz, ph.ic, ph.jc, ph.kc, ph.xfreq, true, &cs)) {
          ph.flags &= ~PH_ALIVE;  // already leaving: dead without tally (raytrace_car.f90:1469-1472)
          { retire_photon(P, ph, false, job, cnt); atomicAdd(q.n_dead, 1u); }
        } else {
          double xp, yp, zp;
          const int st = tau_step(P, vtab, r, tau_in, xp, yp, zp);
          if (st == 1) {
            ph.x = xp; ph.y = yp; ph.z = zp; ph.xfreq = r.xfreq;
            if (!P.zonly) { ph.ic = r.ic; ph.jc = r.jc; }
            ph.kc = r.kc;
            ph.flags |= PH_SCATTER;
            cnt.cellsteps += r.nsteps;
            moved = true;
          } else if (st == 2) {
            cnt.cellsteps += finish_escape(P, ph, r);
            { retire_photon(P, ph, true, job, cnt); atomicAdd(q.n_dead, 1u); }
            moved = true;
          } else {  // crossed into the next cell: the trace stage walks it (from its start) with this tau
            pl.f[(size_t)F_TAU * S + s] = tau_in;
            ph.flags |= PH_TAUPEND;
          }
        }
      }
    }
    // ---- write back what a scattering changes
    {
      int fl = ph.flags;
      if (have_rng) { store_rng(pl, s, rng, fl); nrng += rng.nrng; }
      else fl |= fl0 & PH_GAUSS;  // the spare deviate k_wf_draw left stays
      pl.flags[s] = fl;
      double *o = pl.f + s;
      o[F_XFREQ * S] = ph.xfreq;
      o[F_KX * S] = ph.kx; o[F_KY * S] = ph.ky; o[F_KZ * S] = ph.kz;
      o[F_NSG * S] = ph.nsg;
      if (STOKES) {
        o[F_MX * S] = ph.mx; o[F_MY * S] = ph.my; o[F_MZ * S] = ph.mz;
        o[F_NX * S] = ph.nx; o[F_NY * S] = ph.ny; o[F_NZ * S] = ph.nz;
        o[F_Q * S] = ph.Q; o[F_U * S] = ph.U; o[F_V * S] = ph.V;
      }
      if (DUST) { o[F_NSD * S] = ph.nsd; if (ph.wgt != wgt_in) o[F_WGT * S] = ph.wgt; }
      if (LOCAL && moved) {
        o[F_X * S] = ph.x; o[F_Y * S] = ph.y; o[F_Z * S] = ph.z; o[F_XREF * S] = ph.xfreq_ref;
        pl.ic[s] = ph.ic; pl.jc[s] = ph.jc; pl.kc[s] = ph.kc;
      }
    }
  }
  flush_counters(P, cnt, nrng);
}

// stage 4: raytrace_to_edge for every queued peel ray, per-lane refill, deposit
#ifndef LART_PEEL_MINBLOCKS
#define LART_PEEL_MINBLOCKS 2
#endif
#ifndef LART_PEEL_REFILL
#define LART_PEEL_REFILL kRefillMin
#endif
template <bool PLAIN>
__global__ void __launch_bounds__(kBlock, LART_PEEL_MINBLOCKS) k_wf_peel(const __grid_constant__ DevParams P, Pool pl, Queues q, int budget, int cont_only) {
  __shared__ double vtab[kVoigtTabN];
  load_vtab(P, vtab);
  const unsigned FULL = 0xffffffffu;
  Counters cnt;
  // work items of this partition: rays suspended in the previous wave, then the rays this wave's emit and scatter stages queued
  const unsigned par = *q.wave & 1u;
  const PeelCont *cin = q.cont[par];
  PeelCont *cout = q.cont[par ^ 1u];
  const unsigned ncont = min(q.n_cont[par], q.cont_cap);
  const unsigned n = ncont + (cont_only ? 0u : min(*q.n_direct, q.direct_cap));
  Ray r;
  PeelRay pr;
  bool have = false, exhausted = false;
  for (;;) {
    bool zero_tau = false;
    bool need = !have && !exhausted;
    unsigned nm = __ballot_sync(FULL, need), hm = __ballot_sync(FULL, have);
    if (nm && (__popc(nm) >= LART_PEEL_REFILL || !hm)) {
      unsigned idx = reserve(q.head_peel, need);
      if (need) {
        if (idx >= n) exhausted = true;
        else if (idx < ncont) {  // resume a suspended ray
          const PeelCont &c = cin[idx];
          ray_load(pr, &c.pr);
          ray_resume<PLAIN>(P, r, pr.x, pr.y, pr.z, pr.kx, pr.ky, pr.kz, c.tx, c.ty, c.tz, c.delx, c.dely, c.delz, c.d, c.tau,
                     c.xfreq, c.u1, c.ic, c.jc, c.kc, !PLAIN && P.x.edge_open != 0);
          have = true;
        } else {
          ray_load(pr, q.rays + q.direct_base + (idx - ncont));
          cnt.peel += 1;
          if (ray_setup<PLAIN>(P, r, pr.x, pr.y, pr.z, pr.kx, pr.ky, pr.kz, pr.ic, pr.jc, pr.kc, pr.xfreq, false)) zero_tau = true;
          else have = true;
        }
      }
    }
    bool fin = zero_tau;
    if (have && edge_step<PLAIN>(P, vtab, r)) { fin = true; have = false; cnt.cellsteps += r.nsteps; }
    unsigned fm = __ballot_sync(FULL, fin);
    if (fin) peel_deposit(P, pr, zero_tau ? 0.0 : r.tau, fm);
    if (have && r.nsteps >= budget) {  // park the ray: the next wave continues it
      unsigned at = atomicAdd(&q.n_cont[par ^ 1u], 1u);
      if (at < q.cont_cap) {
        PeelCont &c = cout[at];
        ray_store(&c.pr, pr);
        c.tx = r.tx; c.ty = r.ty; c.tz = r.tz; c.delx = r.delx; c.dely = r.dely; c.delz = r.delz;
        c.d = r.d; c.tau = r.tau; c.xfreq = r.xfreq; c.u1 = r.u1; c.ic = r.ic; c.jc = r.jc; c.kc = r.kc | ((r.flip & 7) << kFlipShift);
        cnt.cellsteps += r.nsteps;
        have = false;
      } else {
        atomicSub(&q.n_cont[par ^ 1u], 1u);  // queue full (cannot happen by sizing): keep walking
      }
    }
    if (!__any_sync(FULL, have || !exhausted)) break;
  }
  flush_counters(P, cnt, 0);
}

__global__ void k_wf_reset(Queues q) {  // start of a wave: new parity, empty output queues
  *q.n_direct = 0; *q.head_trace = 0; *q.head_peel = 0;
  const unsigned w = *q.wave + 1u;
  *q.wave = w;
  q.n_cont[(w & 1u) ^ 1u] = 0;
}

// ------------------------------ clump medium, wavefront ---------------------
// The same four stages as the Cartesian wavefront driver.  The two walker stages (k_cl_flight, k_cl_peel) step the
// resumable clump ray tracer (lart_clump.cuh) one CSR cell at a time and refill each lane on its own, so that the rare
// ray that crosses hundreds of cells no longer idles the 31 other lanes of its warp for the whole of its walk
// (k_mono_clump: 4.95 of 32 threads active per issued instruction).  No step budgets: a walk runs to its end.
// The clump index of a slot lives in pl.rc[slot]; a peel ray carries it in PeelRay::ic.
__global__ void __launch_bounds__(kBlock) k_cl_emit(const __grid_constant__ DevParams P, Pool pl, Job *job, Queues q) {
  Counters cnt;
  ctr_t nrng = 0;
  for (int s = pl.s0 + blockIdx.x * blockDim.x + threadIdx.x; s < pl.s0 + pl.n; s += gridDim.x * blockDim.x) {
    if (pl.flags[s] & PH_ALIVE) continue;
    if (job->next >= job->count) continue;
    unsigned long long j = atomicAdd(&job->next, 1ULL);
    if (j >= job->count) continue;
    Photon ph;
    Rng rng;
    CellData cs;
    ph.id = job->first_id + (long long)j * job->stride;
    rng.start(P.seed, (unsigned long long)ph.id);
    generate_photon(P, ph, rng, cnt, cs);
    const int icl = clump_at_point(P.cl, ph.x, ph.y, ph.z);  // generate_photon.f90:325-332
    if (icl > 0) {
      const ClumpPhys cp = load_clump(P.cl, icl);
      ph.xfreq = DSUB(ph.xfreq, ulos_clump(P, cp, ph.kx, ph.ky, ph.kz));
      const double ratio = cp.Dfreq / P.cl.Dfreq_ref;
      cs.vfx = DMUL(cp.vx, ratio); cs.vfy = DMUL(cp.vy, ratio); cs.vfz = DMUL(cp.vz, ratio);
    }
    if (P.save_all_photons) record_initial(P, ph);
    if (P.save_peeloff) {
      for (int i = 0; i < P.nobs; ++i) {
        PeelRay pr;
        if (!peel_direct_prepare(P, P.obs[i], i, ph, cs, pr, icl > 0)) continue;
        pr.ic = icl;
        unsigned at = atomicAdd(q.n_direct, 1u);
        if (at < q.direct_cap) ray_store(q.rays + q.direct_base + at, pr);
        else atomicOr(P.err, (unsigned)ERR_DIRECT_QUEUE);
      }
    }
    int fl = ph.flags;
    store_rng(pl, s, rng, fl);
    ph.flags = fl;
    store_all(pl, s, ph);
    pl.rc[s] = icl;
    pl.nev[s] = 0;
    nrng += rng.nrng;
  }
  flush_counters(P, cnt, nrng);
}

// Walker scheduling: every iteration the warp votes for ONE kind of work — refill, a clump segment, or a CSR cell — the
// kind most lanes are waiting for, and only those lanes run it.  Lanes of other kinds wait their turn.  (With every
// lane doing "its" step each iteration the three code paths ran one after the other with a quarter of the lanes each:
// ncu 5.7 / 7.9 of 32 threads active per instruction in the flight / peel stage.)
__device__ __forceinline__ int cw_vote(bool need, bool have, int phase, double d, bool &any, unsigned &m_cell, unsigned &m_find) {
  const unsigned FULL = 0xffffffffu;
  const int nn = __popc(__ballot_sync(FULL, need));
  // a lane whose search cw_coop_cells has just ended (d = huge) only needs its own short step to resolve it: like CW_FIND
  m_cell = __ballot_sync(FULL, have && phase == CW_CELL && d < kHugest);
  m_find = __ballot_sync(FULL, have && (phase == CW_FIND || (phase == CW_CELL && !(d < kHugest))));
  const int nc = __popc(m_cell | m_find);
  const int nl = __popc(__ballot_sync(FULL, have && phase == CW_CLUMP));
  any = (nn | nc | nl) != 0;
  int sel = 1, mx = nc;           // 1: cell (and search set-up)
  if (nl > mx) { mx = nl; sel = 2; }  // 2: clump segment
  if (nn > mx) { sel = 0; }           // 0: refill
  return sel;
}

// When only a few lanes of a warp still walk CSR cells, the warp walks ONE of those rays together: lane j takes the ray's
// j-th cell ahead (it replays the DDA j steps from the leader's state: a few adds, no memory), the 32 cells are tested
// in one round of loads instead of 32, and the outcome is put together exactly as the sequential loop of
// find_next_clump would have produced it — cells in traversal order, the walk stops before the first cell that starts
// beyond the best hit or the sphere, or after the cell whose successor lies outside the grid; of equal entry distances
// the first in traversal order wins.  Must be called by all 32 lanes; `leader` holds a ray in phase CW_CELL.
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ void cw_coop_cells(const DevClumps &C, ClumpWalk &w, int leader) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  if (lane == leader) {  // the DDA increments are needed by every lane: compute what is still missing (once per ray)
    if (w.delx < 0.0) w.delx = C.dx / fabs(w.kx);
    if (w.dely < 0.0) w.dely = C.dy / fabs(w.ky);
    if (w.delz < 0.0) w.delz = C.dz / fabs(w.kz);
  }
  const double x = shfl_d(w.x, leader), y = shfl_d(w.y, leader), z = shfl_d(w.z, leader);
  const double kx = shfl_d(w.kx, leader), ky = shfl_d(w.ky, leader), kz = shfl_d(w.kz, leader);
  const double delx = shfl_d(w.delx, leader), dely = shfl_d(w.dely, leader), delz = shfl_d(w.delz, leader);
  const double t_sp = shfl_d(w.t_sp, leader), best_in = shfl_d(w.best_te, leader);
  double tx = shfl_d(w.tx, leader), ty = shfl_d(w.ty, leader), tz = shfl_d(w.tz, leader), d = shfl_d(w.d, leader);
  int ci = __shfl_sync(FULL, w.ci, leader), cj = __shfl_sync(FULL, w.cj, leader), ck = __shfl_sync(FULL, w.ck, leader);
  const int si = __shfl_sync(FULL, w.si, leader), sj = __shfl_sync(FULL, w.sj, leader), sk = __shfl_sync(FULL, w.sk, leader);
  const int skip = __shfl_sync(FULL, w.skip_icl, leader), icl_in = __shfl_sync(FULL, w.best_icl, leader);
  auto advance = [&]() -> bool {  // one DDA step; false when it leaves the grid (same arithmetic as cw_find_cell)
    if (tx <= ty && tx <= tz) { d = tx; ci += si; if (ci < 0 || ci >= C.cgx) return false; tx = DADD(tx, delx); }
    else if (ty <= tz) { d = ty; cj += sj; if (cj < 0 || cj >= C.cgy) return false; ty = DADD(ty, dely); }
    else { d = tz; ck += sk; if (ck < 0 || ck >= C.cgz) return false; tz = DADD(tz, delz); }
    return true;
  };
  bool alive = true;
  for (int i = 0; i < lane && alive; ++i) alive = advance();
  // my cell: nearest entry among its clumps (list order, strict <)
  const double d_mine = d;
  double te_mine = kHugest;
  int icl_mine = 0;
  if (alive) {
    const size_t icell = (size_t)ci + (size_t)C.cgx * ((size_t)cj + (size_t)C.cgy * (size_t)ck);
    const int p0 = __ldg(C.cg_start + icell), p1 = __ldg(C.cg_start + icell + 1);
    for (int ip = p0; ip < p1; ++ip) {
      const int icl = __ldg(C.cg_list + ip - 1);
      const double4 g = ldg4(C.geo_reg + ip - 1);
      if (icl == skip) continue;
      const double rx = DSUB(x, g.x), ry = DSUB(y, g.y), rz = DSUB(z, g.z);
      const double b = DADD(DADD(DMUL(rx, kx), DMUL(ry, ky)), DMUL(rz, kz));
      double disc = DADD(DSUB(DMUL(b, b), DADD(DADD(DMUL(rx, rx), DMUL(ry, ry)), DMUL(rz, rz))), g.w);
      if (disc < 0.0) continue;
      disc = sqrt(disc);
      const double te = DSUB(-b, disc), tx2 = DADD(-b, disc);
      if (tx2 > 0.0 && te < te_mine) { te_mine = te; icl_mine = icl; }
    }
  }
  const bool out_after = alive && !advance();  // the cell after mine lies outside the grid
  // inclusive prefix "first minimum" over lanes 0..j, seeded with the leader's best so far
  double pt = te_mine;
  int pi = icl_mine;
  if (lane == 0 && !(te_mine < best_in)) { pt = best_in; pi = icl_in; }
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double qt = __shfl_up_sync(FULL, pt, o);
    const int qi = __shfl_up_sync(FULL, pi, o);
    if (lane >= o && !(pt < qt)) { pt = qt; pi = qi; }  // the earlier lanes keep a tie
  }
  double et = __shfl_up_sync(FULL, pt, 1);  // exclusive prefix: best before my cell
  int ei = __shfl_up_sync(FULL, pi, 1);
  if (lane == 0) { et = best_in; ei = icl_in; }
  const bool chk = alive && (d_mine > et || d_mine > t_sp);  // the sequential loop stops before my cell
  const unsigned mc = __ballot_sync(FULL, chk), mo = __ballot_sync(FULL, out_after);
  const int jc = mc ? __ffs(mc) - 1 : 32, jo = mo ? __ffs(mo) - 1 : 32;
  int processed, src;
  bool stop, excl;
  if (jc <= jo && jc < 32) { processed = jc; src = jc; excl = true; stop = true; }
  else if (jo < 32) { processed = jo + 1; src = jo; excl = false; stop = true; }
  else { processed = 32; src = 31; excl = false; stop = false; }
  const double rt = shfl_d(excl ? et : pt, src);
  const int ri = __shfl_sync(FULL, excl ? ei : pi, src);
  // state of the cell after lane 31's (valid when !stop)
  const double ntx = shfl_d(tx, 31), nty = shfl_d(ty, 31), ntz = shfl_d(tz, 31), nd = shfl_d(d, 31);
  const int nci = __shfl_sync(FULL, ci, 31), ncj = __shfl_sync(FULL, cj, 31), nck = __shfl_sync(FULL, ck, 31);
  if (lane == leader) {
    w.best_te = rt; w.best_icl = ri; w.ncells += processed;
    if (stop) w.d = kHugest;  // the leader's next own step sees the search over and resolves it
    else { w.tx = ntx; w.ty = nty; w.tz = ntz; w.d = nd; w.ci = nci; w.cj = ncj; w.ck = nck; }
  }
}
// Walk together when no more than this many lanes of the warp are on CSR cells (0 = never).  Measured per wave of 2.4 M
// photons (clump_sphere_fcov5): flight stage 2.20 -> 1.59 ms with 12, 20 or 32 (its searches are long: the mean free
// path between clumps is ~20 cells); peel stage 3.16 ms without, 3.36 with 4, 3.48 with 12 (most of its rays end in
// their own clump, the others keep many lanes busy) — so flights only.
#ifndef LART_COOP_FLIGHT
#define LART_COOP_FLIGHT 12
#endif
#ifndef LART_COOP_PEEL
#define LART_COOP_PEEL 0
#endif

// stage 2: flights (forced first scattering included) of every alive slot that is not at a scattering point
__global__ void __launch_bounds__(kBlock, 3) k_cl_flight(const __grid_constant__ DevParams P, Pool pl, Job *job, Queues q) {
  __shared__ double vtab[kVoigtTabN];
  load_vtab(P, vtab);
  Counters cnt;
  ctr_t nrng = 0;
  ClumpWalk w;
  Photon ph;
  Rng rng;
  int slot = -1, mode = 0;  // mode 0: uncapped edge walk of the forced first scattering, 1: tau walk
  bool have = false, exhausted = false;
  CellData cs0;
  w.phase = 0; w.d = 0.0;
  for (;;) {
    const bool need = !have && !exhausted;
    bool any;
    unsigned m_cell, m_find;
    const int sel = cw_vote(need, have, w.phase, w.d, any, m_cell, m_find);
    if (!any) break;
    if (sel == 1 && m_cell && !m_find && __popc(m_cell) <= LART_COOP_FLIGHT) {  // few lanes left on cells: walk one ray together
      cw_coop_cells(P.cl, w, __ffs(m_cell) - 1);
      continue;
    }
    if (sel == 0) {
      unsigned idx = reserve(q.head_trace, need);
      if (need) {
        if (idx >= (unsigned)pl.n) exhausted = true;
        else if ((pl.flags[pl.s0 + idx] & (PH_ALIVE | PH_SCATTER)) == PH_ALIVE) {
          slot = pl.s0 + (int)idx;
          load_trace_part(pl, slot, ph);
          load_rng(P, pl, slot, ph.id, ph.flags, rng);
          const int icl = pl.rc[slot];
          if (ph.flags & PH_FIRST) {
            int ci, cj, ck;
            clamp_cell_for_read(P, ph, ci, cj, ck);
            load_cell(P, ci, cj, ck, cs0);
            mode = 0;
            cw_start(w, ph.x, ph.y, ph.z, ph.kx, ph.ky, ph.kz, ph.xfreq, icl, 0.0);
          } else {
            mode = 1;
            cw_start(w, ph.x, ph.y, ph.z, ph.kx, ph.ky, ph.kz, ph.xfreq, icl, -log(rng.uniform()));
          }
          have = true;
        }
      }
      continue;
    }
    const bool mine = have && ((sel == 2) == (w.phase == CW_CLUMP));
    if (mine) {
      if (mode == 0) {
        if (cw_edge_step(P, vtab, w, -1.0)) {  // tau0 known: escaped fraction, weight, first optical depth
          cnt.cellsteps += w.ncells;
          const double tau = forced_first(P, ph, rng, cs0, w.tau);
          mode = 1;
          cw_start(w, ph.x, ph.y, ph.z, ph.kx, ph.ky, ph.kz, ph.xfreq, pl.rc[slot], tau);
        }
      } else {
        const int st = cw_tau_step(P, vtab, w);
        if (st != 0) {
          int icl;
          cnt.cellsteps += w.ncells;
          cw_finish_tau(P, w, st, ph, icl);
          if (st == 1) {
            ph.flags |= PH_SCATTER;
          } else {
            load_rest(pl, slot, ph);
            ph.flags &= ~PH_ALIVE;
            ph.xfreq_ref = ph.xfreq;  // see k_mono_clump
            retire_photon(P, ph, true, job, cnt);
          }
          store_trace_part(pl, slot, ph);
          pl.rc[slot] = icl;
          pl.ndraw[slot] = rng.nblk;
          nrng += rng.nrng;
          have = false;
        }
      }
    }
  }
  flush_counters(P, cnt, nrng);
}

// stage 3: scattering inside the slot's clump; peel rays go to the slot's ray entries
__global__ void __launch_bounds__(kBlock, 2) k_cl_scatter(const __grid_constant__ DevParams P, Pool pl, Job *job, Queues q) {
  __shared__ double vtab[kVoigtTabN];
  __shared__ VzWarpShared vzsh[kBlock / 32];
  if (P.dust) load_vtab(P, vtab);
  VzWarpShared &sh = vzsh[threadIdx.x >> 5];
  Counters cnt;
  ctr_t nrng = 0;
  const int lane = threadIdx.x & 31;
  const int stride = gridDim.x * blockDim.x;
  for (int base = pl.s0 + blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < pl.s0 + pl.n; base += stride) {
    const int s = base + lane;
    const bool inb = s < pl.s0 + pl.n;
    const int fl0 = inb ? pl.flags[s] : 0;
    const bool active = (fl0 & PH_SCATTER) != 0;
    PeelRay *myrays = q.rays + (size_t)(inb ? s : 0) * P.nobs;
    if (inb && !active)
      for (int k = 0; k < P.nobs; ++k) myrays[k].kind = -1;
    if (!__any_sync(0xffffffffu, active)) continue;
    Photon ph;
    Rng rng;
    CellData cs, csp;
    ClumpPhys cp;
    int icl = 0;
    double scale = 1.0, ratio = 1.0, xloc = 0.0, a_loc = 1.0;
    bool to_dust = false;
    if (active) {
      load_trace_part(pl, s, ph);
      ph.flags &= ~PH_SCATTER;
      load_rng(P, pl, s, ph.id, ph.flags, rng);
      load_cell(P, ph.ic, ph.jc, ph.kc, cs);  // the box: no opacity, no bulk velocity, reference Doppler width
      icl = pl.rc[s];
      cp = load_clump(P.cl, icl);
      ratio = cp.Dfreq / P.cl.Dfreq_ref; scale = P.cl.Dfreq_ref / cp.Dfreq;
      csp = cs;
      csp.vfx = DMUL(cp.vx, ratio); csp.vfy = DMUL(cp.vy, ratio); csp.vfz = DMUL(cp.vz, ratio);
      cnt.scatter += 1;
      if (P.dust) {  // scattering_car.f90:72-87
        double pd = cp.rhokapD / (cp.rhokap * voigt_seon2(vtab, DMUL(ph.xfreq, scale), cp.voigt_a) + cp.rhokapD);
        to_dust = rng.uniform() <= pd;
      }
      xloc = ph.xfreq * scale; a_loc = cp.voigt_a;
    } else {
      rng.start(P.seed, 0ULL);
    }
    const bool resonant = active && !to_dust;
    const double uz_loc = rand_resonance_vz_warp(sh, resonant, rng, xloc, a_loc, cnt.reject);  // do_resonance1_clump
    if (!active) continue;
    load_rest(pl, s, ph);
    bool peeled = false;
    auto emit_ray = [&](int k, bool ok, PeelRay &pr) {
      if (ok) { pr.ic = icl; ray_store(myrays + k, pr); }
      else myrays[k].kind = -1;
    };
    if (to_dust) {
      scatter_dust(P, ph, rng, cs, cnt, [&]() {
        peeled = true;
        for (int k = 0; k < P.nobs; ++k) {
          PeelRay pr;
          bool ok = P.use_stokes ? peel_dust_stokes_prepare(P, P.obs[k], k, ph, csp, pr)
                                 : peel_dust_nostokes_prepare(P, P.obs[k], k, ph, csp, pr);
          emit_ray(k, ok, pr);
        }
      });
      if (!(ph.flags & PH_ALIVE)) retire_photon(P, ph, false, job, cnt);
    } else {
      const double scale_inv = 1.0 / scale;
      const double xfreq_atom = (xloc - uz_loc) * scale_inv, uz = uz_loc * scale_inv;  // line_clump_mod.f90:41-44
      CellData css = cs;
      css.Dfreq = cp.Dfreq;  // recoil uses the clump's Doppler width (scattering_car.f90:428-433)
      scatter_resonance_core<true>(P, ph, rng, css, cnt, uz, xfreq_atom, ratio, [&](double xa, double ux, double uy, double uzz) {
        peeled = true;
        for (int k = 0; k < P.nobs; ++k) {
          PeelRay pr;
          bool ok = P.use_stokes ? peel_resonance_stokes_prepare(P, P.obs[k], k, ph, csp, xa, ux, uy, uzz, pr)
                                 : peel_resonance_nostokes_prepare(P, P.obs[k], k, ph, csp, xa, ux, uy, uzz, pr);
          emit_ray(k, ok, pr);
        }
      });
    }
    if (!peeled) for (int k = 0; k < P.nobs; ++k) myrays[k].kind = -1;
    if (P.max_events > 0 && (ph.flags & PH_ALIVE)) {
      const int ne = pl.nev[s] + 1;
      pl.nev[s] = ne;
      if (ne >= P.max_events) {
        ph.flags &= ~PH_ALIVE;
        ph.xfreq_ref = ph.xfreq;
        retire_photon(P, ph, false, job, cnt);
      }
    }
    int fl = ph.flags;
    store_rng(pl, s, rng, fl);
    ph.flags = fl;
    store_all(pl, s, ph);
    nrng += rng.nrng;
  }
  flush_counters(P, cnt, nrng);
}

// stage 4: capped edge walks (peel_raytrace_to_edge, peelingoff_rect.f90:894-906) of the slot rays and the direct rays
__global__ void __launch_bounds__(kBlock, 4) k_cl_peel(const __grid_constant__ DevParams P, Pool pl, Queues q) {
  __shared__ double vtab[kVoigtTabN];
  load_vtab(P, vtab);
  const unsigned FULL = 0xffffffffu;
  Counters cnt;
  const unsigned nslot = (unsigned)pl.n * (unsigned)P.nobs, slot_lo = (unsigned)pl.s0 * (unsigned)P.nobs;
  const unsigned n = nslot + min(*q.n_direct, q.direct_cap);
  ClumpWalk w;
  PeelRay pr;
  bool have = false, exhausted = false;
  w.phase = 0; w.d = 0.0;
  for (;;) {
    const bool need = !have && !exhausted;
    bool any;
    unsigned m_cell, m_find;
    const int sel = cw_vote(need, have, w.phase, w.d, any, m_cell, m_find);
    if (!any) break;
    if (sel == 1 && m_cell && !m_find && __popc(m_cell) <= LART_COOP_PEEL) {  // few lanes left on cells: walk one ray together
      cw_coop_cells(P.cl, w, __ffs(m_cell) - 1);
      continue;
    }
    if (sel == 0) {
      unsigned idx = reserve(q.head_peel, need);
      if (need) {
        if (idx >= n) exhausted = true;
        else {
          const PeelRay *src = q.rays + (idx < nslot ? slot_lo + idx : q.direct_base + (idx - nslot));
          if (src->kind >= 0) {
            ray_load(pr, src);
            cnt.peel += 1;
            cw_start(w, pr.x, pr.y, pr.z, pr.kx, pr.ky, pr.kz, pr.xfreq, pr.ic, 0.0);
            have = true;
          }
        }
      }
      continue;
    }
    const bool mine = have && ((sel == 2) == (w.phase == CW_CLUMP));
    bool fin = false;
    if (mine && cw_edge_step(P, vtab, w, kTauHugeClump)) { fin = true; have = false; cnt.cellsteps += w.ncells; }
    const unsigned fm = __ballot_sync(FULL, fin);
    if (fin) peel_deposit(P, pr, w.tau, fm);
  }
  flush_counters(P, cnt, 0);
}

// ------------------------------ pool compaction ------------------------------
// Once the job queue is empty the pool thins out.  Compaction moves the live photons of the tail into the
// dead slots of the head [0, n_keep), so that every kernel works on a dense range again (full warps in the
// scatter stage, short scans, and a dense monolithic tail).
__global__ void k_compact_list(Pool pl, int n_keep, int *src, int *dst, unsigned int *n_src, unsigned int *n_dst) {
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < pl.n; s += gridDim.x * blockDim.x) {
    const bool alive = (pl.flags[s] & PH_ALIVE) != 0;
    if (s >= n_keep && alive) src[atomicAdd(n_src, 1u)] = s;
    if (s < n_keep && !alive) dst[atomicAdd(n_dst, 1u)] = s;
  }
}
__global__ void k_compact_move(Pool pl, const int *src, const int *dst, const unsigned int *n_src) {
  const unsigned m = *n_src;
  const size_t S = pl.S;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    const int a = src[i], b = dst[i];
    for (int c = 0; c < F_COUNT; ++c) pl.f[(size_t)c * S + b] = pl.f[(size_t)c * S + a];
    for (int c = 0; c < 10; ++c) pl.rs[(size_t)c * S + b] = pl.rs[(size_t)c * S + a];
    for (int c = 0; c < 3; ++c) pl.rc[(size_t)c * S + b] = pl.rc[(size_t)c * S + a];
    pl.id[b] = pl.id[a]; pl.ndraw[b] = pl.ndraw[a];
    pl.ic[b] = pl.ic[a]; pl.jc[b] = pl.jc[a]; pl.kc[b] = pl.kc[a];
    pl.flags[b] = pl.flags[a];
    pl.nev[b] = pl.nev[a];
    pl.flags[a] = 0;
  }
}

// ------------------------------ set-up kernels ------------------------------
__global__ void k_pack_cells(DevParams P, Cell *cells, size_t n) {
  for (size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x; c < n; c += (size_t)gridDim.x * blockDim.x) {
    Cell o;
    o.rhokap = P.rhokap[c]; o.voigt_a = P.voigt_a[c]; o.Dfreq = P.Dfreq[c];
    o.vfx = P.vfx[c]; o.vfy = P.vfy[c]; o.vfz = P.vfz[c];
    o.rhokapD = P.dust ? P.rhokapD[c] : 0.0; o.pad = 0.0;
    const int i = (int)(c % P.nx), j = (int)((c / P.nx) % P.ny), k = (int)(c / ((size_t)P.nx * P.ny));
    cells[cell_slot(P, i + 1, j + 1, k + 1)] = o;
  }
}
__global__ void k_build_vtab(double *tab) {
  for (int k = threadIdx.x; k < 202; k += blockDim.x) {
    tab[4 * k + 0] = k < 101 ? c_voigt_h0[k] : 0.0;
    tab[4 * k + 1] = c_voigt_h1[k];
    tab[4 * k + 2] = k < 101 ? c_voigt_h2[k] : 0.0;
    tab[4 * k + 3] = c_voigt_h3[k];
  }
}

// ------------------------------ batch kernels -------------------------------
__global__ void k_voigt_batch(const double *tab, long long n, const double *x, const double *a, double *H) {
  __shared__ double vtab[kVoigtTabN];
  for (int i = threadIdx.x; i < kVoigtTabN; i += blockDim.x) vtab[i] = tab[i];
  __syncthreads();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    H[i] = voigt_seon2(vtab, x[i], a[i]);
}
__global__ void k_edge_batch(const __grid_constant__ DevParams P, long long n, const double *x, const double *y,
                             const double *z, const double *kx, const double *ky, const double *kz, const double *xfreq,
                             const int *ic, const int *jc, const int *kc, double *tau, int *nsteps, int trace_cap, int *trace) {
  __shared__ double vtab[kVoigtTabN];
  load_vtab(P, vtab);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int ns;
    tau[i] = walk_edge(P, vtab, x[i], y[i], z[i], kx[i], ky[i], kz[i], ic[i], jc[i], kc[i], xfreq[i], ns, trace_cap,
                       trace ? trace + i * trace_cap : nullptr);
    if (nsteps) nsteps[i] = ns;
  }
}
__global__ void k_tau_batch(const __grid_constant__ DevParams P, long long n, double *x, double *y, double *z,
                            const double *kx, const double *ky, const double *kz, double *xfreq, int *ic, int *jc, int *kc,
                            const double *tau_in, int *inside, double *xfreq_ref, int *nsteps) {
  __shared__ double vtab[kVoigtTabN];
  load_vtab(P, vtab);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    Photon ph;
    ph.x = x[i]; ph.y = y[i]; ph.z = z[i]; ph.kx = kx[i]; ph.ky = ky[i]; ph.kz = kz[i];
    ph.xfreq = xfreq[i]; ph.ic = ic[i]; ph.jc = jc[i]; ph.kc = kc[i]; ph.flags = PH_ALIVE; ph.xfreq_ref = 0.0;
    ph.vshear = 0.0;  // a fresh photon (generate_photon.f90:141); shearing-box batches start without an offset
    ph.wgt = 0.0;     // batch walks make no tallies: with CALCJ / CALCPnew on, their deposits are exact zeros
    CellData cs;
    int ns = walk_tau(P, vtab, ph, tau_in[i], cs);
    x[i] = ph.x; y[i] = ph.y; z[i] = ph.z; xfreq[i] = ph.xfreq; ic[i] = ph.ic; jc[i] = ph.jc; kc[i] = ph.kc;
    inside[i] = (ph.flags & PH_ALIVE) ? 1 : 0;
    if (xfreq_ref) xfreq_ref[i] = (ph.flags & PH_ALIVE) ? 0.0 : ph.xfreq_ref;
    if (nsteps) nsteps[i] = ns < 0 ? 0 : ns;
  }
}
__global__ void k_clump_edge_batch(const __grid_constant__ DevParams P, long long n, const double *x, const double *y,
                                   const double *z, const double *kx, const double *ky, const double *kz, const double *xfreq,
                                   const int *icl, double tau_max, double *tau, int *nclumps, unsigned long long *ncells) {
  __shared__ double vtab[kVoigtTabN];
  load_vtab(P, vtab);
  unsigned long long mine = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int nc = 0, ncl = 0;
    tau[i] = P.cl.overlap ? clump_walk_edge_overlap(P, vtab, x[i], y[i], z[i], kx[i], ky[i], kz[i], xfreq[i], tau_max, nc)
                          : clump_walk_edge(P, vtab, x[i], y[i], z[i], kx[i], ky[i], kz[i], xfreq[i], icl[i], tau_max, nc, ncl);
    if (nclumps) nclumps[i] = ncl;
    mine += (unsigned long long)nc;
  }
  if (ncells && mine) atomicAdd(ncells, mine);
}
__global__ void k_clump_tau_batch(const __grid_constant__ DevParams P, long long n, double *x, double *y, double *z,
                                  const double *kx, const double *ky, const double *kz, double *xfreq, int *icl,
                                  const double *tau_in, int *inside) {
  __shared__ double vtab[kVoigtTabN];
  load_vtab(P, vtab);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    Photon ph;
    ph.x = x[i]; ph.y = y[i]; ph.z = z[i]; ph.kx = kx[i]; ph.ky = ky[i]; ph.kz = kz[i]; ph.xfreq = xfreq[i];
    ph.ic = ph.jc = ph.kc = 1; ph.flags = PH_ALIVE;
    int c = icl[i], nc = 0;
    const bool in = clump_walk_tau(P, vtab, ph, c, tau_in[i], nc);
    x[i] = ph.x; y[i] = ph.y; z[i] = ph.z; xfreq[i] = ph.xfreq; icl[i] = c; inside[i] = in ? 1 : 0;
  }
}
__global__ void k_clump_locate_batch(const __grid_constant__ DevParams P, long long n, const double *x, const double *y,
                                     const double *z, int *icl) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    icl[i] = clump_at_point(P.cl, x[i], y[i], z[i]);
}
__global__ void k_amr_locate_batch(const __grid_constant__ DevParams P, long long n, const double *x, const double *y,
                                   const double *z, int *il) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    il[i] = amr_find_leaf(P, x[i], y[i], z[i]);
}
// host arrays -> device records (geometry: centre + radius^2; physics: 64 bytes)
__global__ void k_clump_geo_reg(long long nreg, const int *cg_list, const double4 *geo, double4 *geo_reg) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nreg; i += (long long)gridDim.x * blockDim.x)
    geo_reg[i] = geo[cg_list[i] - 1];
}
__global__ void k_pack_clumps(long long n, const double *x, const double *y, const double *z, const double *radius,
                              const double *rhokap, const double *rhokapD, const double *voigt_a, const double *Dfreq,
                              const double *vx, const double *vy, const double *vz, double4 *geo, ClumpPhys *phys) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    geo[i] = make_double4(x[i], y[i], z[i], DMUL(radius[i], radius[i]));
    ClumpPhys cp;
    cp.rhokap = rhokap[i]; cp.rhokapD = rhokapD ? rhokapD[i] : 0.0; cp.voigt_a = voigt_a[i]; cp.Dfreq = Dfreq[i];
    cp.vx = vx[i]; cp.vy = vy[i]; cp.vz = vz[i]; cp.pad_ = 0.0;
    phys[i] = cp;
  }
}
// the scatter stage's "this peel ray certainly ends inside its own cell" bound, exposed for the parity tests
__global__ void k_peel_bound_batch(const __grid_constant__ DevParams P, long long n, const double *x, const double *y,
                                   const double *z, const double *xfreq, const int *ic, const int *jc, const int *kc, int *capped) {
  __shared__ double vtab[kVoigtTabN];
  load_vtab(P, vtab);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    CellData cs;
    load_cell(P, ic[i], jc[i], kc[i], cs);
    PeelRay pr;
    pr.x = x[i]; pr.y = y[i]; pr.z = z[i]; pr.xfreq = xfreq[i]; pr.ic = ic[i]; pr.jc = jc[i]; pr.kc = kc[i];
    capped[i] = peel_certainly_capped(P, vtab, cs, pr) ? 1 : 0;
  }
}
__global__ void k_xcrit_batch(const __grid_constant__ DevParams P, long long n, const double *x, const double *y,
                              const double *z, const int *ic, const int *jc, const int *kc, double *out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double xc = 0.0, xc2 = 0.0;
    if (P.core_skip_global) { xc = P.xcrit; }
    else if (ic[i] >= 1 && jc[i] >= 1 && kc[i] >= 1) {
      CellData cs;
      load_cell(P, ic[i], jc[i], kc[i], cs);
      car_xcrit_local(P, ic[i], jc[i], kc[i], x[i], y[i], z[i], cs.voigt_a, cs.rhokap, xc, xc2);
    }
    out[i] = xc;
  }
}
__global__ void k_sample_batch(int kind, unsigned long long seed, long long n, const long long *ids, const double *p0,
                               const double *p1, int ndraw, double *out) {
  __shared__ VzWarpShared vzsh[kBlock / 32];
  const int lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * blockDim.x;
  // warp-uniform loop: kind 6 is a warp collective
  for (long long base = blockIdx.x * (long long)blockDim.x + (threadIdx.x & ~31); base < n; base += stride) {
    const long long i = base + lane;
    const bool mine = i < n;
    Rng r;
    r.start(seed, (unsigned long long)(mine ? (ids ? ids[i] : i) : 0));
    ctr_t nrej = 0;
    const double q0 = (mine && p0) ? p0[i] : 0.0, q1 = (mine && p1) ? p1[i] : 1.0;
    for (int j = 0; j < ndraw; ++j) {
      double v = 0.0;
      if (kind == 6) v = rand_resonance_vz_warp(vzsh[threadIdx.x >> 5], mine, r, q0, q1, nrej);
      else if (mine) {
        switch (kind) {
          case 0: v = r.uniform(); break;
          case 1: v = r.gauss(nrej); break;
          case 2: v = rand_resonance_vz(r, q0, q1, nrej); break;
          case 3: v = rand_resonance(r, q0); break;
          case 4: v = rand_hg(r, q0); break;
          default: v = rand_voigt(r, q0, nrej); break;
        }
      }
      if (mine) out[i * ndraw + j] = v;
    }
  }
}

// ---- sight-line maps: raytrace_to_edge_tau_gas / _column for every pixel and frequency -------
struct SightStart {  // entry point of one pixel's sight line (host-computed, sightline_tau_rect.f90:45-150)
  double x, y, z, kx, ky, kz;
  int ic, jc, kc, valid;
};
// work item = (pixel, kk): kk < nxfreq -> tau_gas at bin centre kk; kk == nxfreq -> column density + dust tau.
// Consecutive items share the pixel, i.e. the cell sequence: warps stay converged.
__global__ void __launch_bounds__(kBlock) k_sightline(const __grid_constant__ DevParams P, long long npix, const SightStart *st,
                                                      double cross0, double *tau_gas, double *N_gas, double *tau_dust,
                                                      unsigned long long *steps) {
  __shared__ double vtab[kVoigtTabN];
  load_vtab(P, vtab);
  const long long per = P.nxfreq + 1, total = npix * per;
  unsigned long long ns = 0;
  for (long long it = blockIdx.x * (long long)blockDim.x + threadIdx.x; it < total; it += (long long)gridDim.x * blockDim.x) {
    const long long pix = it / per;
    const int kk = (int)(it - pix * per);
    const SightStart s = st[pix];
    if (!s.valid) continue;
    Ray r;
    if (kk < P.nxfreq) {
      CellData c0;
      load_cell(P, s.ic, s.jc, s.kc, c0);
      const double u1 = vdotk(c0, s.kx, s.ky, s.kz);
      const double xf = DADD(DMUL(DSUB((double)(kk + 1), 0.5), P.dxfreq), P.xfreq_min);  // grid%xfreq(kk)
      const double x0 = DSUB(DMUL(xf, P.Dfreq_ref) / c0.Dfreq, u1);
      double tau = 0.0;
      if (!ray_setup(P, r, s.x, s.y, s.z, s.kx, s.ky, s.kz, s.ic, s.jc, s.kc, x0, false)) {
        for (;;) {  // raytrace_to_edge_car_tau_gas: gas opacity only, no tau cap
          if (P.x.atm && ray_masked(P, r)) { r.tau = __longlong_as_double(0x7ff0000000000000LL); break; }  // _atmosphere :3832-3836
          const double kap = DMUL(r.cell.rhokap, voigt_seon2(vtab, r.xfreq, r.cell.voigt_a));
          ++r.nsteps;
          const int ax = ray_axis(P, r);
          const double tn = (ax == 1) ? r.tx : (ax == 2) ? r.ty : r.tz;
          r.tau = DADD(r.tau, DMUL(DSUB(tn, r.d), kap));
          r.d = tn;
          if (!ray_advance(P, r, ax)) break;
          ray_shift(P, r);
        }
        tau = r.tau;
        ns += r.nsteps;
      }
      tau_gas[(size_t)kk + (size_t)P.nxfreq * pix] = tau;
    } else {
      double N = 0.0, td = 0.0;
      if (!ray_setup(P, r, s.x, s.y, s.z, s.kx, s.ky, s.kz, s.ic, s.jc, s.kc, 0.0, false)) {
        for (;;) {  // raytrace_to_edge_car_column
          if (P.x.atm && ray_masked(P, r)) {  // _column_atmosphere :3933-3938
            N = __longlong_as_double(0x7ff0000000000000LL);
            if (P.dust) td = N;
            break;
          }
          const double rho = DMUL(r.cell.rhokap, r.cell.Dfreq) / cross0;
          const int ax = ray_axis(P, r);
          const double tn = (ax == 1) ? r.tx : (ax == 2) ? r.ty : r.tz;
          const double del = DSUB(tn, r.d);
          N = DADD(N, DMUL(del, rho));
          if (P.dust) td = DADD(td, DMUL(del, r.cell.rhokapD));
          r.d = tn;
          ++r.nsteps;
          if (!ray_advance(P, r, ax)) break;
          load_cell(P, r.ic, r.jc, r.kc, r.cell);
        }
        ns += r.nsteps;
      }
      N_gas[pix] = N;
      if (tau_dust) tau_dust[pix] = td;
    }
  }
  for (int o = 16; o > 0; o >>= 1) ns += __shfl_xor_sync(0xffffffffu, ns, o);
  if ((threadIdx.x & 31) == 0 && ns) atomicAdd(steps, ns);
}

// register-resident DFMA loop: 16 independent chains per thread
__global__ void k_dfma_peak(double *out, int iters, double a, double b) {
  double v[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) v[q] = threadIdx.x * 1e-3 + q;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = fma(v[q], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < 16; ++q) s += v[q];
  if (s == 12345.678) *out = s;
}

// shared per-device Voigt table for the handle-less batch entry points
std::mutex g_tab_mu;
double *g_tab[64] = {nullptr};
int device_vtab(int dev, double **out) {
  std::lock_guard<std::mutex> lk(g_tab_mu);
  if (dev < 0 || dev >= 64) return fail("bad device ordinal");
  if (!g_tab[dev]) {
    CUDA_OK(cudaMalloc(&g_tab[dev], sizeof(double) * kVoigtTabN));
    k_build_vtab<<<1, 256>>>(g_tab[dev]);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaDeviceSynchronize());
  }
  *out = g_tab[dev];
  return 0;
}

// Handle memory comes from a stream-ordered memory pool PRIVATE to this library (one per device), with the release
// threshold lifted while handles are alive: what a destroyed handle frees stays cached for the next lart_gpu_create of
// the process (measured: create 0.06-0.17 s, but 0.6-0.8 s when it followed the cudaFree of another handle's ~10 GB).
// The device's default pool — PyTorch's or the host program's — is never touched, and when the last handle of a
// device goes the pool is trimmed, i.e. the memory returns to the driver (LART_GPU_KEEP_POOL=1 keeps it cached).
std::mutex g_pool_mu;
cudaMemPool_t g_pool[64] = {nullptr};
int g_pool_users[64] = {0};
int pool_for(int dev, cudaMemPool_t *mp) {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  if (dev < 0 || dev >= 64) return fail("bad device ordinal");
  if (!g_pool[dev]) {
    cudaMemPoolProps props{};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    CUDA_OK(cudaMemPoolCreate(&g_pool[dev], &props));
    unsigned long long thr = ~0ULL;
    CUDA_OK(cudaMemPoolSetAttribute(g_pool[dev], cudaMemPoolAttrReleaseThreshold, &thr));
  }
  *mp = g_pool[dev];
  return 0;
}
void pool_acquire(int dev) {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  if (dev >= 0 && dev < 64) ++g_pool_users[dev];
}
void pool_release(int dev) {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  if (dev < 0 || dev >= 64 || g_pool_users[dev] <= 0) return;
  if (--g_pool_users[dev] == 0 && g_pool[dev] && !getenv("LART_GPU_KEEP_POOL")) cudaMemPoolTrimTo(g_pool[dev], 0);
}
int dev_malloc(void **p, size_t bytes) {
  int dev = 0;
  CUDA_OK(cudaGetDevice(&dev));
  cudaMemPool_t mp;
  if (int rc = pool_for(dev, &mp)) return rc;
  CUDA_OK(cudaMallocFromPoolAsync(p, bytes, mp, 0));
  CUDA_OK(cudaStreamSynchronize(0));  // the pointer is used from other streams right away
  return 0;
}
void dev_free(void *p) { cudaFreeAsync(p, 0); }

template <class T>
int upload(T **dst, const T *src, size_t n) {
  *dst = nullptr;
  if (!src || n == 0) return 0;
  if (int rc = dev_malloc((void **)dst, n * sizeof(T))) return rc;
  CUDA_OK(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}


// ---- NCCL, opened at run time
struct NcclApi {
  void *lib = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclReduce) Reduce = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
} g_nccl;
ncclComm_t g_comm = nullptr;
int g_comm_rank = 0, g_comm_size = 1, g_comm_device = -1;
int nccl_load() {
  if (g_nccl.lib) return 0;
  const char *names[] = {getenv("LART_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void *lib = nullptr;
  for (const char *n : names) if (n && *n && (lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
  if (!lib) return fail(std::string("lart_gpu: cannot open libnccl (set LART_NCCL_LIB): ") + dlerror());
#define NCCL_SYM(f) if (!(g_nccl.f = (decltype(g_nccl.f))dlsym(lib, "nccl" #f))) return fail("lart_gpu: libnccl lacks nccl" #f)
  NCCL_SYM(GetUniqueId); NCCL_SYM(CommInitRank); NCCL_SYM(CommDestroy); NCCL_SYM(Reduce); NCCL_SYM(GroupStart);
  NCCL_SYM(GroupEnd); NCCL_SYM(GetErrorString);
#undef NCCL_SYM
  g_nccl.lib = lib;
  return 0;
}
#define NCCL_OK(call)                                                                                         \
  do {                                                                                                        \
    ncclResult_t r_ = (call);                                                                                 \
    if (r_ != ncclSuccess) return fail(std::string(#call) + " failed: " + g_nccl.GetErrorString(r_));         \
  } while (0)

// ---- device tallies ADDED into caller-owned host arrays: pinned double-buffered staging, the copy of chunk i+1
// overlaps the multi-threaded add of chunk i (the caller's arrays are pageable Fortran memory)
struct Seg { double *dst; long long off, n; };
constexpr size_t kStageDoubles = 4u << 20;  // 32 MB per staging buffer
int add_device_to_host(lart_gpu_ctx *h, const double *dev, long long total, const std::vector<Seg> &segs_in);
}  // namespace

struct lart_gpu_ctx {
  double *soa_slab = nullptr;  // the host's SoA grid arrays on the device (one allocation; freed after packing unless LART_FLAG_SOA_GRID)
  long long jp_bins = 0;  // bins of the CALCP / CALCPnew arrays (CALCJ: nxfreq times as many)
  int device = 0;
  DevParams P{};
  Pool pool{};
  struct Group {  // one wave pipeline: a pool partition with its own queues, stream and events
    Pool pool;
    Queues q;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    std::vector<cudaEvent_t> tev;
  };
  std::vector<Group> groups;
  std::map<int, cudaGraphExec_t> graphs;  // one instantiated graph per wave count (quantum)
  cudaEvent_t fork = nullptr;
  PeelRay *rays = nullptr;
  Job *job = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::vector<void *> owned;  // every device allocation
  double *allph_buf = nullptr;
  long long allph_n = 0;
  int allph_slot[10];
  int quantum = 0, flags = 0, nsm = 148, nobs = 0, nxim = 0, nyim = 0;
  int budget = 32;          // cell steps a trace/peel ray may take per wave before it is parked
  bool pending_rays = false;  // suspended walks may exist (drain before switching drivers / fetching)
  unsigned int *ctr = nullptr;  // per-partition queue counters
  size_t ctr_n = 0;
  int *cmp_src = nullptr, *cmp_dst = nullptr;  // compaction work lists
  unsigned int *cmp_n = nullptr;
  unsigned long long job_next = 0;  // job queue head after the last step
  long long job_count = 0;          // ids in the job queue's current range (h->count = all ids handed to this handle so far)
  bool more_to_claim = false;       // lart_gpu_run_dealt: the node's shared counter still has photons to deal
  std::vector<DevObserver> obs_host;  // host copy of the observers (sight-line maps)
  double sight_steps = 0.0, sight_ms = 0.0;
  long long count = 0;
  double kernel_ms = 0.0;
  long long launches = 0;
  bool begun = false;
  double *pinned[2] = {nullptr, nullptr};  // pinned staging of lart_gpu_fetch / the grid upload (process-wide buffers, see host_staging)
  std::vector<cudaEvent_t> tev;  // stage-timing events of one step (monolithic driver)
  long long ray_cap = 0;  // entries of the ray array
  int draw_mode = 0;    // 0: k_wf_draw2 (compacted rejection loops), 1: serial per-lane loops, 2: warp-cooperative speculative trials
  int draw_chunk = 1024; // slots per warp of k_wf_draw2 (upper bound; LART_GPU_DRAW_CHUNK)
  double stage_ms[LART_STAGE_COUNT] = {};
  long long stage_n[LART_STAGE_COUNT] = {};
};

namespace {
void partition_pool(lart_gpu_handle h, int n);
int create_impl(const lart_config *cfg, lart_gpu_ctx *h);

// Two 32-MB page-locked staging buffers, allocated at the first use in the process and kept (page-locking costs
// ~10 ms per buffer; every handle of the process — one per run of the host program — reuses them).
std::mutex g_stage_mu;
double *g_stage[2] = {nullptr, nullptr};
int host_staging(lart_gpu_ctx *h) {
  std::lock_guard<std::mutex> lk(g_stage_mu);
  for (int k = 0; k < 2; ++k) {
    if (!g_stage[k]) CUDA_OK(cudaHostAlloc((void **)&g_stage[k], kStageDoubles * sizeof(double), cudaHostAllocPortable));
    h->pinned[k] = g_stage[k];
  }
  return 0;
}

// Host arrays -> device through the staging buffers: T host threads, each its own pipeline (copy a piece into its share
// of a staging buffer, async H2D on its own stream, two pieces in flight).  A pageable cudaMemcpy of the six 65-MB grid
// arrays of a 201^3 run moves ~10 GB/s; this keeps the DMA engine fed from page-locked memory.
struct UpJob { double *dev; const double *src; size_t n; };
int upload_pipelined(lart_gpu_ctx *h, const std::vector<UpJob> &jobs) {
  size_t total = 0;
  for (const UpJob &j : jobs) total += j.n;
  if (total < (size_t)(1 << 20)) {  // small grids: the plain copy is faster than starting threads
    for (const UpJob &j : jobs) if (j.n) CUDA_OK(cudaMemcpy(j.dev, j.src, j.n * sizeof(double), cudaMemcpyHostToDevice));
    return 0;
  }
  if (int rc = host_staging(h)) return rc;
  const int T = (int)std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
  const size_t sub = kStageDoubles / T;
  std::vector<UpJob> pieces;
  for (const UpJob &j : jobs)
    for (size_t c = 0; c < j.n; c += sub) pieces.push_back({j.dev + c, j.src + c, std::min(sub, j.n - c)});
  std::atomic<size_t> next{0};
  std::atomic<int> bad{0};
  const int dev = h->device;
  auto worker = [&](int t) {
    if (cudaSetDevice(dev) != cudaSuccess) { bad = 1; return; }
    cudaStream_t st;
    cudaEvent_t ev[2];
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) { bad = 1; return; }
    for (int k = 0; k < 2; ++k) cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming);
    bool used[2] = {false, false};
    int b = 0;
    for (size_t i; (i = next.fetch_add(1)) < pieces.size() && !bad; b ^= 1) {
      const UpJob &pc = pieces[i];
      double *stage = h->pinned[b] + (size_t)t * sub;
      if (used[b] && cudaEventSynchronize(ev[b]) != cudaSuccess) { bad = 1; break; }
      memcpy(stage, pc.src, pc.n * sizeof(double));
      if (cudaMemcpyAsync(pc.dev, stage, pc.n * sizeof(double), cudaMemcpyHostToDevice, st) != cudaSuccess) { bad = 1; break; }
      cudaEventRecord(ev[b], st);
      used[b] = true;
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) bad = 1;
    for (int k = 0; k < 2; ++k) cudaEventDestroy(ev[k]);
    cudaStreamDestroy(st);
  };
  std::vector<std::thread> th;
  for (int t = 1; t < T; ++t) th.emplace_back(worker, t);
  worker(0);
  for (auto &x : th) x.join();
  if (bad) return fail(std::string("lart_gpu_create: host-to-device copy of the grid failed: ") + cudaGetErrorString(cudaGetLastError()));
  return 0;
}

int add_device_to_host(lart_gpu_ctx *h, const double *dev, long long total, const std::vector<Seg> &segs_in) {
  if (total <= 0 || segs_in.empty()) return 0;
  std::vector<Seg> segs = segs_in;
  std::sort(segs.begin(), segs.end(), [](const Seg &a, const Seg &b) { return a.off < b.off; });
  if (int rc = host_staging(h)) return rc;
  const long long lo = segs.front().off, hi = segs.back().off + segs.back().n;
  const int nthreads = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  cudaEvent_t ev[2];
  for (int k = 0; k < 2; ++k) CUDA_OK(cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming));
  auto issue = [&](long long c0, int buf) -> int {
    const long long n = std::min<long long>((long long)kStageDoubles, hi - c0);
    CUDA_OK(cudaMemcpyAsync(h->pinned[buf], dev + c0, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaEventRecord(ev[buf], h->stream));
    return 0;
  };
  int rc = issue(lo, 0);
  int buf = 0;
  for (long long c0 = lo; c0 < hi && !rc; c0 += (long long)kStageDoubles, buf ^= 1) {
    const long long c1 = std::min<long long>(c0 + (long long)kStageDoubles, hi);
    if (c1 < hi) rc = issue(c1, buf ^ 1);
    if (cudaEventSynchronize(ev[buf]) != cudaSuccess) { rc = fail("lart_gpu_fetch: device-to-host copy failed"); break; }
    const double *src = h->pinned[buf];
    // pieces of this chunk: the intersection of [c0,c1) with every segment
    struct Piece { double *dst; const double *src; long long n; };
    std::vector<Piece> pieces;
    for (const Seg &s : segs) {
      const long long a = std::max(s.off, c0), b = std::min(s.off + s.n, c1);
      if (a < b) pieces.push_back({s.dst + (a - s.off), src + (a - c0), b - a});
    }
    long long work = 0;
    for (const Piece &p : pieces) work += p.n;
    const int nt = work < (1 << 16) ? 1 : nthreads;
    auto body = [&](int t) {
      for (const Piece &p : pieces) {
        const long long a = p.n * t / nt, b = p.n * (t + 1) / nt;
        double *d = p.dst;
        const double *q = p.src;
        for (long long i = a; i < b; ++i) d[i] += q[i];
      }
    };
    if (nt == 1) body(0);
    else {
      std::vector<std::thread> th;
      for (int t = 1; t < nt; ++t) th.emplace_back(body, t);
      body(0);
      for (auto &x : th) x.join();
    }
  }
  for (int k = 0; k < 2; ++k) cudaEventDestroy(ev[k]);
  return rc;
}
}  // namespace

namespace {
template <class T>
int dalloc(lart_gpu_ctx *h, T **p, size_t n, bool zero = true) {
  if (int rc = dev_malloc((void **)p, std::max<size_t>(n, 1) * sizeof(T))) return rc;
  if (zero) CUDA_OK(cudaMemset(*p, 0, std::max<size_t>(n, 1) * sizeof(T)));
  h->owned.push_back(*p);
  return 0;
}
template <class T>
int dupload(lart_gpu_ctx *h, const T **dst, const T *src, size_t n) {
  T *p = nullptr;
  if (int rc = upload(&p, src, n)) return rc;
  if (p) h->owned.push_back(p);
  *dst = p;
  return 0;
}

struct Scratch {  // device copies of host arrays for one batch call
  std::vector<void *> p;
  ~Scratch() { for (void *q : p) cudaFree(q); }
  template <class T>
  int in(T **d, const T *hsrc, size_t n) {
    *d = nullptr;
    if (!hsrc) return 0;
    CUDA_OK(cudaMalloc(d, std::max<size_t>(n, 1) * sizeof(T)));
    p.push_back(*d);
    CUDA_OK(cudaMemcpy(*d, hsrc, n * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
  }
  template <class T>
  int outbuf(T **d, size_t n) {
    CUDA_OK(cudaMalloc(d, std::max<size_t>(n, 1) * sizeof(T)));
    p.push_back(*d);
    return 0;
  }
};
inline int grid_for(long long n, int nsm) { return (int)std::max<long long>(1, std::min<long long>((n + kBlock - 1) / kBlock, nsm * 8LL)); }
#define D2H(dst, src, n) CUDA_OK(cudaMemcpy((dst), (src), (n) * sizeof(*(dst)), cudaMemcpyDeviceToHost))


int validate(const lart_config *c) {
  const lart_grid &g = c->grid;
  const lart_params &p = c->par;
  if (p.use_amr_grid) {  // octree: the box, Dfreq_ref and the frequency grid come from lart_grid, everything else from lart_amr
    const lart_amr &a = c->amr;
    if (a.nleaf < 1 || a.ncells < a.nleaf) return fail("lart_gpu_create: use_amr_grid needs amr.nleaf >= 1 and ncells >= nleaf");
    if (!a.children || !a.ileaf || !a.icell_of_leaf || !a.neighbor || !a.cx || !a.cy || !a.cz || !a.ch || !a.rhokap || !a.voigt_a ||
        !a.Dfreq || !a.vfx || !a.vfy || !a.vfz)
      return fail("lart_gpu_create: NULL octree array");
    if (p.DGR > 0.0 && !a.rhokapD) return fail("lart_gpu_create: DGR > 0 but amr.rhokapD is NULL");
    if (p.use_clump_medium || p.xy_periodic || p.xyz_symmetry || p.xy_symmetry)
      return fail("lart_gpu_create: clump media, periodic and mirror boundaries on an octree stay with the Fortran host");
    if (c->flags & (LART_FLAG_SOA_GRID | LART_FLAG_LOCAL_STEPS)) return fail("lart_gpu_create: LART_FLAG_SOA_GRID / LOCAL_STEPS do not apply to an octree");
    if (g.nxfreq < 1 || !(g.xmax > g.xmin)) return fail("lart_gpu_create: octree box / frequency grid missing in lart_grid");
    for (int64_t i = 0; i < (int64_t)a.ncells; ++i) {
      if (a.ileaf[i] < 0 || a.ileaf[i] > a.nleaf) return fail("lart_gpu_create: amr.ileaf out of range");
      for (int q = 0; q < 8; ++q) if (a.children[8 * i + q] < 0 || a.children[8 * i + q] > a.ncells) return fail("lart_gpu_create: amr.children out of range");
      for (int q = 0; q < 6; ++q) if (a.neighbor[6 * i + q] < 0 || a.neighbor[6 * i + q] > a.ncells) return fail("lart_gpu_create: amr.neighbor out of range");
    }
    for (int64_t i = 0; i < (int64_t)a.nleaf; ++i)
      if (a.icell_of_leaf[i] < 1 || a.icell_of_leaf[i] > a.ncells || a.ileaf[a.icell_of_leaf[i] - 1] != i + 1)
        return fail("lart_gpu_create: amr.icell_of_leaf / ileaf are not inverse to each other");
  } else {
  if (g.nx < 1 || g.ny < 1 || g.nz < 1 || g.nxfreq < 1) return fail("lart_gpu_create: grid dimensions must be >= 1");
  if (!g.xface || !g.yface || !g.zface || !g.rhokap || !g.voigt_a || !g.Dfreq || !g.vfx || !g.vfy || !g.vfz)
    return fail("lart_gpu_create: NULL grid array");
  }
  if (c->line.line_type != 1) return fail("lart_gpu_create: only line_type 1 (Ly-alpha singlet) is on the GPU path");
  if (p.DGR > 0.0 && !p.use_amr_grid && !g.rhokapD) return fail("lart_gpu_create: DGR > 0 but rhokapD is NULL");
  if (p.DGR > 0.0 && p.use_stokes && c->scatt_mat.nPDF < 2) return fail("lart_gpu_create: dust + Stokes needs scatt_mat");
  if (p.nobs < 0 || p.nobs > LART_MAX_OBSERVERS) return fail("lart_gpu_create: nobs out of range");
  if (p.save_peeloff && p.nobs > 0 && !c->observers) return fail("lart_gpu_create: observers is NULL");
  if (p.save_Jmu && (p.nmu < 1 || !(p.dmu > 0.0))) return fail("lart_gpu_create: save_Jmu needs nmu >= 1 and dmu > 0");
  if (!(g.dxfreq > 0.0)) return fail("lart_gpu_create: dxfreq must be > 0");
  if (g.nz >= (1 << kFlipShift)) return fail("lart_gpu_create: nz too large");
  if (g.nxfreq >= (1 << 24)) return fail("lart_gpu_create: nxfreq must be < 2^24 (peel-tally aggregation key)");
  if (p.save_peeloff && p.nobs > 0 && c->observers && (long long)c->observers[0].nxim * c->observers[0].nyim >= (1LL << 30))
    return fail("lart_gpu_create: nxim*nyim must be < 2^30 (peel-tally aggregation key)");
  if (c->max_events < 0) return fail("lart_gpu_create: max_events must be >= 0");
  if (p.use_clump_medium) {  // setup.f90:806-860; grid_mod_clump.f90:55-59
    const lart_clumps &cl = c->clumps;
    if (cl.n < 1 || cl.n > 2147483647LL) return fail("lart_gpu_create: use_clump_medium needs 1 <= clumps.n < 2^31");
    if (!cl.x || !cl.y || !cl.z || !cl.vx || !cl.vy || !cl.vz || !cl.radius || !cl.rhokap || !cl.voigt_a || !cl.Dfreq ||
        !cl.cg_start || !cl.cg_list)
      return fail("lart_gpu_create: NULL clump array");
    if (p.DGR > 0.0 && !cl.rhokapD) return fail("lart_gpu_create: DGR > 0 but clumps.rhokapD is NULL");
    if (cl.cgx < 1 || cl.cgy < 1 || cl.cgz < 1 || !(cl.cg_dx > 0.0) || !(cl.cg_dy > 0.0) || !(cl.cg_dz > 0.0) || !(cl.sphere_R > 0.0))
      return fail("lart_gpu_create: bad clump CSR grid");
    if (p.xyz_symmetry || p.xy_symmetry || p.xy_periodic) return fail("lart_gpu_create: the clump medium uses the plain box (grid_mod_clump.f90:55-59)");
    // the CSR arrays are indexed on the device without further checks: offsets 1-based and monotone, entries in 1..n
    const size_t ncell = (size_t)cl.cgx * cl.cgy * cl.cgz;
    if (cl.cg_start[0] != 1) return fail("lart_gpu_create: clumps.cg_start must be 1-based (cg_start[0] == 1)");
    for (size_t i = 0; i < ncell; ++i)
      if (cl.cg_start[i + 1] < cl.cg_start[i]) return fail("lart_gpu_create: clumps.cg_start is not monotone");
    const long long nreg = (long long)cl.cg_start[ncell] - 1;
    if (nreg < 1) return fail("lart_gpu_create: clumps CSR grid holds no registration");
    for (long long i = 0; i < nreg; ++i)
      if (cl.cg_list[i] < 1 || (long long)cl.cg_list[i] > cl.n) return fail("lart_gpu_create: clumps.cg_list entry outside 1..n");
  }
  if (p.xyz_symmetry || p.xy_symmetry) {  // setup.f90:167, 198-206, 952-957; grid_mod_car.f90:85-134
    if (p.xy_periodic || g.nx < 2 || g.ny < 2 || (p.xyz_symmetry && g.nz < 2))
      return fail("lart_gpu_create: xyz_symmetry / xy_symmetry need a 3-D, non-periodic grid");
    if (p.xyz_symmetry && p.save_peeloff) return fail("lart_gpu_create: peeling-off is not allowed with xyz_symmetry (setup.f90:198)");
    if (g.i0 < 1 || g.i0 > 2 || g.j0 < 1 || g.j0 > 2 || (p.xyz_symmetry && (g.k0 < 1 || g.k0 > 2)))
      return fail("lart_gpu_create: folded grids need grid.i0/j0(/k0) in {1,2} (grid_mod_car.f90:85-134)");
  }
  // ---- atmospheres, shearing boxes, CALCJ / CALCP / CALCPnew
  const bool extras = p.atmosphere || p.Omega != 0.0 || p.calc_J || p.calc_P || p.calc_Pnew;
  if (extras && (p.use_clump_medium || p.use_amr_grid))
    return fail("lart_gpu_create: atmospheres, shearing boxes and the CALCJ/CALCP accumulators are Cartesian-grid options");
  if (p.atmosphere < 0 || p.atmosphere > LART_ATM_SPHERICAL) return fail("lart_gpu_create: par.atmosphere must be a LART_ATM_* value");
  if (p.atmosphere == LART_ATM_PLANE && !(p.xy_periodic && g.nx == 1 && g.ny == 1))
    return fail("lart_gpu_create: a plane atmosphere is a 1 x 1 x nz periodic column (setup.f90:86-94)");
  if (p.atmosphere == LART_ATM_SPHERICAL) {
    if (p.xy_periodic || p.xyz_symmetry) return fail("lart_gpu_create: a spherical atmosphere excludes xy_periodic and xyz_symmetry (setup.f90:95-98)");
    if (!g.mask) return fail("lart_gpu_create: a spherical atmosphere needs grid.mask");
  }
  if (p.source_geometry == LART_SRC_PLANE_ILLUMINATION && !p.atmosphere)
    return fail("lart_gpu_create: plane illumination needs an atmosphere geometry (generate_photon.f90:742-778)");
  if (p.source_geometry < 0 || p.source_geometry > LART_SRC_PLANE_ILLUMINATION) return fail("lart_gpu_create: unknown source_geometry");
  if (p.Omega != 0.0 && !(p.xy_periodic && g.nx > 1 && g.ny > 1))
    return fail("lart_gpu_create: par.Omega (shearing box) needs xy_periodic with nx, ny > 1 (setup.f90:966-969)");
  if (p.calc_J || p.calc_P || p.calc_Pnew) {
    const int gj = g.geometry_JPa;
    if (!(gj == 3 || gj == 2 || gj == 1 || gj == -1)) return fail("lart_gpu_create: grid.geometry_JPa must be 3, 2, 1 or -1");
    if ((gj == 1 || gj == 2) && g.nr < 1) return fail("lart_gpu_create: grid.nr must be >= 1 for geometry_JPa 1 and 2");
    if (gj == 1 && !g.ind_sph) return fail("lart_gpu_create: geometry_JPa = 1 needs grid.ind_sph");
    if (gj == 2 && !g.ind_cyl) return fail("lart_gpu_create: geometry_JPa = 2 needs grid.ind_cyl");
    if ((p.calc_P || p.calc_Pnew) && !(c->line.cross0 > 0.0)) return fail("lart_gpu_create: CALCP / CALCPnew need line.cross0 > 0");
    if (c->flags & LART_FLAG_LOCAL_STEPS) return fail("lart_gpu_create: LART_FLAG_LOCAL_STEPS does not carry the CALCJ/CALCP accumulators");
  }
  return 0;
}
}  // namespace

extern "C" {

const char *lart_gpu_last_error(void) { return g_err.c_str(); }
int lart_gpu_version(void) { return 100; }

int lart_gpu_create(const lart_config *cfg, lart_gpu_handle *out) {
  if (!cfg || !out) return fail("lart_gpu_create: NULL argument");
  *out = nullptr;
  if (int rc = validate(cfg)) return rc;
  int ndev = 0;
  CUDA_OK(cudaGetDeviceCount(&ndev));
  if (cfg->device < 0 || cfg->device >= ndev) return fail("lart_gpu_create: no such CUDA device");
  CUDA_OK(cudaSetDevice(cfg->device));
  lart_gpu_ctx *h = new lart_gpu_ctx();
  h->device = cfg->device;
  pool_acquire(h->device);
  if (int rc = create_impl(cfg, h)) {  // every error path releases the half-built context
    const std::string msg = g_err;
    lart_gpu_destroy(h);
    g_err = msg;
    return rc;
  }
  *out = h;
  return 0;
}

}  // extern "C"

namespace {
int create_impl(const lart_config *cfg, lart_gpu_ctx *h) {
  auto bail = [&](int rc) { return rc; };
  const bool timing = getenv("LART_GPU_TIMING") != nullptr;  // phases of lart_gpu_create on stderr
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (!timing) return;
    cudaDeviceSynchronize();
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "lart_gpu_create: %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
    t_last = now;
  };
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, cfg->device));
  h->nsm = prop.multiProcessorCount;
  CUDA_OK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CUDA_OK(cudaEventCreate(&h->ev0));
  CUDA_OK(cudaEventCreate(&h->ev1));
  CUDA_OK(cudaEventCreateWithFlags(&h->fork, cudaEventDisableTiming));
  const lart_grid &g = cfg->grid;
  const lart_params &p = cfg->par;
  DevParams &P = h->P;
  P.nx = g.nx; P.ny = g.ny; P.nz = g.nz; P.nxfreq = g.nxfreq;
  P.dx = g.dx; P.dy = g.dy; P.dz = g.dz;
  P.xmin = g.xmin; P.ymin = g.ymin; P.zmin = g.zmin; P.xmax = g.xmax; P.ymax = g.ymax; P.zmax = g.zmax;
  P.Dfreq_ref = g.Dfreq_ref; P.xfreq_min = g.xfreq_min; P.xfreq_max = g.xfreq_max; P.dxfreq = g.dxfreq;
  P.xcrit = g.xcrit; P.xcrit2 = g.xcrit2; P.rmax = g.rmax;
  const size_t nc = (size_t)g.nx * g.ny * g.nz;
  P.dust = p.DGR > 0.0 ? 1 : 0;
  int rc = 0;
  P.amr = DevAmr{};
  const bool amr_on = p.use_amr_grid != 0;
  if (amr_on) {
    // octree: leaf physics as packed 64-byte records indexed by leaf, a 64-byte geometry record per leaf
    // (centre, half-width, face neighbours of its cell) and a 64-byte record per cell for descents
    const lart_amr &a = cfg->amr;
    P.nx = a.nleaf; P.ny = 1; P.nz = 1;
    P.dx = g.xmax - g.xmin; P.dy = g.ymax - g.ymin; P.dz = g.zmax - g.zmin;
    const double faces[6] = {g.xmin, g.xmax, g.ymin, g.ymax, g.zmin, g.zmax};
    rc = rc ? rc : dupload(h, &P.xface, faces, 2);
    rc = rc ? rc : dupload(h, &P.yface, faces + 2, 2);
    rc = rc ? rc : dupload(h, &P.zface, faces + 4, 2);
    static_assert(sizeof(Cell) == 64 && sizeof(AmrGeo) == 64 && sizeof(AmrCell) == 64, "octree records are 64 bytes");
    std::vector<Cell> hc(a.nleaf);
    std::vector<AmrGeo> hg(a.nleaf);
    std::vector<AmrCell> hx(a.ncells);
    for (int il = 0; il < a.nleaf; ++il) {
      Cell &c = hc[il];
      c.rhokap = a.rhokap[il]; c.voigt_a = a.voigt_a[il]; c.Dfreq = a.Dfreq[il];
      c.vfx = a.vfx[il]; c.vfy = a.vfy[il]; c.vfz = a.vfz[il]; c.rhokapD = P.dust ? a.rhokapD[il] : 0.0; c.pad = 0.0;
      const int ic = a.icell_of_leaf[il];
      AmrGeo &q = hg[il];
      q.cx = a.cx[ic - 1]; q.cy = a.cy[ic - 1]; q.cz = a.cz[ic - 1]; q.h = a.ch[ic - 1];
      for (int f = 0; f < 6; ++f) q.nb[f] = a.neighbor[6 * (size_t)(ic - 1) + f];
      q.icell = ic; q.pad_ = 0;
    }
    for (int ic = 0; ic < a.ncells; ++ic) {
      AmrCell &c = hx[ic];
      c.cx = a.cx[ic]; c.cy = a.cy[ic]; c.cz = a.cz[ic];
      for (int q = 0; q < 8; ++q) c.child[q] = a.children[8 * (size_t)ic + q];
      c.ileaf = a.ileaf[ic]; c.pad_ = 0;
    }
    Cell *dc = nullptr; AmrGeo *dg = nullptr; AmrCell *dx = nullptr;
    rc = rc ? rc : dalloc(h, &dc, (size_t)a.nleaf, false);
    rc = rc ? rc : dalloc(h, &dg, (size_t)a.nleaf, false);
    rc = rc ? rc : dalloc(h, &dx, (size_t)a.ncells, false);
    lap("octree allocations");
    if (!rc) rc = upload_pipelined(h, {{(double *)dc, (const double *)hc.data(), (size_t)a.nleaf * 8},
                                       {(double *)dg, (const double *)hg.data(), (size_t)a.nleaf * 8},
                                       {(double *)dx, (const double *)hx.data(), (size_t)a.ncells * 8}});
    lap("octree H2D");
    P.cells = dc;
    P.amr.on = 1; P.amr.ncells = a.ncells; P.amr.nleaf = a.nleaf; P.amr.geo = dg; P.amr.cell = dx;
    P.rhokap = P.voigt_a = P.Dfreq = P.vfx = P.vfy = P.vfz = P.rhokapD = nullptr;
  } else {
  rc = rc ? rc : dupload(h, &P.xface, g.xface, g.nx + 1);
  rc = rc ? rc : dupload(h, &P.yface, g.yface, g.ny + 1);
  rc = rc ? rc : dupload(h, &P.zface, g.zface, g.nz + 1);
  {
    const double *src[7] = {g.rhokap, g.voigt_a, g.Dfreq, g.vfx, g.vfy, g.vfz, P.dust ? g.rhokapD : nullptr};
    const double **dst[7] = {&P.rhokap, &P.voigt_a, &P.Dfreq, &P.vfx, &P.vfy, &P.vfz, &P.rhokapD};
    std::vector<UpJob> jobs;
    int narr = 0;
    for (int k = 0; k < 7; ++k) narr += src[k] ? 1 : 0;
    const size_t nc_al = (nc + 31) / 32 * 32;  // 256-byte aligned arrays inside one allocation (one driver call, not seven)
    double *slab = nullptr;
    rc = rc ? rc : dalloc(h, &slab, nc_al * narr, false);
    h->soa_slab = slab;
    for (int k = 0, at = 0; k < 7 && !rc; ++k) {
      *dst[k] = nullptr;
      if (!src[k]) continue;
      double *d = slab + nc_al * (size_t)at++;
      *dst[k] = d;
      jobs.push_back({d, src[k], nc});
    }
    lap("grid allocations");
    rc = rc ? rc : upload_pipelined(h, jobs);
    lap("grid H2D");
  }
  }
  if (rc) return bail(rc);
  double *vt = nullptr;
  if ((rc = dalloc(h, &vt, kVoigtTabN))) return bail(rc);
  k_build_vtab<<<1, 256, 0, h->stream>>>(vt);
  P.voigt_tab = vt;
  P.soa = (cfg->flags & LART_FLAG_SOA_GRID) ? 1 : 0;
  P.warp_agg = (cfg->flags & LART_FLAG_NO_WARP_AGG) ? 0 : 1;
  P.flags_serial_vz = (cfg->flags & LART_FLAG_SERIAL_REJECTION) ? 1 : 0;
  h->draw_mode = (cfg->flags & LART_FLAG_SERIAL_REJECTION) ? 1 : ((cfg->flags & LART_FLAG_SPECULATIVE_REJECTION) ? 2 : 0);
  if (const char *e = getenv("LART_GPU_DRAW_CHUNK")) h->draw_chunk = std::max(32, atoi(e) / 32 * 32);
  // which ray tracers the reference would bind (setup.f90:952-976)
  const bool zonly_grid = p.xy_periodic && g.nx == 1 && g.ny == 1;
  P.sym = p.xyz_symmetry ? 1 : 0; P.i0 = g.i0; P.j0 = g.j0; P.k0 = g.k0;
  P.bcxy = P.sym ? BC_MIRROR : (p.xy_symmetry ? BC_MIRROR : ((p.xy_periodic && !zonly_grid) ? BC_PERIODIC : BC_OPEN));
  P.bcz = P.sym ? BC_MIRROR : BC_OPEN;
  // ---- the less common bindings (setup.f90:959-987) and the CALCJ / CALCP / CALCPnew accumulators
  DevExtras &X = P.x;
  X = DevExtras{};
  X.atm = p.atmosphere;
  X.shear = (P.bcxy == BC_PERIODIC && p.Omega != 0.0) ? 1 : 0;  // :967-969
  X.Omega = p.Omega;
  X.edge_open = (X.atm || X.shear) ? 1 : 0;
  X.calc_J = p.calc_J ? 1 : 0; X.calc_P = p.calc_P ? 1 : 0; X.calc_Pnew = p.calc_Pnew ? 1 : 0;
  X.jp = (X.calc_J || X.calc_P || X.calc_Pnew) ? 1 : 0;
  X.geometry_JPa = g.geometry_JPa; X.nr = g.nr; X.cross0 = cfg->line.cross0;
  X.any = (X.atm || X.shear || X.jp) ? 1 : 0;
  if (X.atm == LART_ATM_SPHERICAL) {
    if ((rc = dupload(h, (const int8_t **)&X.mask, g.mask, nc))) return bail(rc);
  }
  if (X.jp && g.geometry_JPa == 1 && (rc = dupload(h, &X.ind_sph, g.ind_sph, nc))) return bail(rc);
  if (X.jp && g.geometry_JPa == 2 && (rc = dupload(h, &X.ind_cyl, g.ind_cyl, (size_t)g.nx * g.ny))) return bail(rc);
  P.local_steps = ((cfg->flags & LART_FLAG_LOCAL_STEPS) && !P.bcxy && !X.any) ? 1 : 0;  // the in-stage cell step knows no mirror planes or wrap-around
  // ---- clump medium: records packed on the device from the host's arrays (clump_mod.f90:30-118)
  P.clump = 0;
  P.cl = DevClumps{};
  if (p.use_clump_medium) {
    const lart_clumps &cl = cfg->clumps;
    const size_t n = (size_t)cl.n, ncell = (size_t)cl.cgx * cl.cgy * cl.cgz;
    const size_t nreg = (size_t)(cl.cg_start[ncell] - 1);
    P.clump = 1;
    P.cl.n = cl.n; P.cl.sphere_R = cl.sphere_R; P.cl.R2 = cl.sphere_R * cl.sphere_R; P.cl.Dfreq_ref = cl.Dfreq_ref;
    P.cl.cgx = cl.cgx; P.cl.cgy = cl.cgy; P.cl.cgz = cl.cgz;
    P.cl.xmin = cl.cg_xmin; P.cl.ymin = cl.cg_ymin; P.cl.zmin = cl.cg_zmin;
    P.cl.dx = cl.cg_dx; P.cl.dy = cl.cg_dy; P.cl.dz = cl.cg_dz;
    P.cl.inv_dx = 1.0 / cl.cg_dx; P.cl.inv_dy = 1.0 / cl.cg_dy; P.cl.inv_dz = 1.0 / cl.cg_dz;
    double4 *geo = nullptr;
    ClumpPhys *phys = nullptr;
    if ((rc = dalloc(h, &geo, n, false)) || (rc = dalloc(h, &phys, n, false))) return bail(rc);
    if ((rc = dupload(h, &P.cl.cg_start, cl.cg_start, ncell + 1)) || (rc = dupload(h, &P.cl.cg_list, cl.cg_list, nreg))) return bail(rc);
    {
      Scratch sc;
      double *d[11];
      const double *src[11] = {cl.x, cl.y, cl.z, cl.radius, cl.rhokap, cl.rhokapD, cl.voigt_a, cl.Dfreq, cl.vx, cl.vy, cl.vz};
      for (int k = 0; k < 11; ++k) if ((rc = sc.in(&d[k], src[k], n))) return bail(rc);
      k_pack_clumps<<<(int)std::min<size_t>((n + kBlock - 1) / kBlock, (size_t)h->nsm * 8), kBlock, 0, h->stream>>>(
          (long long)n, d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], d[8], d[9], d[10], geo, phys);
      CUDA_OK(cudaGetLastError());
      CUDA_OK(cudaStreamSynchronize(h->stream));
    }
    P.cl.geo = geo; P.cl.phys = phys;
    {
      double4 *geo_reg = nullptr;
      if ((rc = dalloc(h, &geo_reg, nreg, false))) return bail(rc);
      k_clump_geo_reg<<<(int)std::min<size_t>((nreg + kBlock - 1) / kBlock, (size_t)h->nsm * 8), kBlock, 0, h->stream>>>(
          (long long)nreg, P.cl.cg_list, geo, geo_reg);
      CUDA_OK(cudaGetLastError());
      CUDA_OK(cudaStreamSynchronize(h->stream));
      P.cl.geo_reg = geo_reg;
    }
  }
  P.nsbx = (g.nx + 31) / 32; P.nsby = (g.ny + 31) / 32; P.nsbz = (g.nz + 31) / 32;
  if (!P.soa && !amr_on) {
    Cell *cells = nullptr;
    if ((rc = dalloc(h, &cells, (size_t)P.nsbx * P.nsby * P.nsbz * 32768, false))) return bail(rc);
    k_pack_cells<<<h->nsm * 8, 256, 0, h->stream>>>(P, cells, nc);
    P.cells = cells;
    // the SoA copies have served: give their memory back to the pool, the photon pool allocated below reuses it
    CUDA_OK(cudaStreamSynchronize(h->stream));
    const double **soa[7] = {&P.rhokap, &P.voigt_a, &P.Dfreq, &P.vfx, &P.vfy, &P.vfz, &P.rhokapD};
    for (auto pp : soa) *pp = nullptr;
    if (h->soa_slab) {
      void *q = (void *)h->soa_slab;
      h->owned.erase(std::remove(h->owned.begin(), h->owned.end(), q), h->owned.end());
      dev_free(q);
      h->soa_slab = nullptr;
    }
    CUDA_OK(cudaStreamSynchronize(0));
  }
  P.seed = p.seed; P.xfreq0 = p.xfreq0; P.xs = p.xs_point; P.ys = p.ys_point; P.zs = p.zs_point;
  P.source_rmax = p.source_rmax; P.albedo = p.albedo; P.hgg = p.hgg; P.voigt_a0 = p.voigt_a0; P.Dfreq0 = p.Dfreq0;
  P.gaussian_sigma_x = p.gaussian_sigma_x; P.mu_min = p.mu_min; P.dmu = p.dmu; P.nmu = p.nmu;
  P.E1 = cfg->line.E1; P.E2 = cfg->line.E2; P.E3 = cfg->line.E3; P.g_recoil0 = cfg->line.g_recoil0;
  if (P.E1 > 0.0) {
    P.rr_p2 = std::sqrt((4.0 - P.E1) / (3.0 * P.E1));
    P.rr_inv = 1.0 / (P.E1 * (P.rr_p2 * P.rr_p2 * P.rr_p2));
  }
  P.spectral_type = p.spectral_type; P.source_geometry = p.source_geometry;
  P.zonly = (p.xy_periodic && g.nx == 1 && g.ny == 1) ? 1 : 0;
  P.comoving_source = p.comoving_source; P.recoil = p.recoil; P.core_skip = p.core_skip; P.core_skip_global = p.core_skip_global;
  P.use_stokes = p.use_stokes; P.use_reduced_wgt = p.use_reduced_wgt;
  P.save_Jin = p.save_Jin; P.save_Jabs = (p.save_Jabs && P.dust) ? 1 : 0; P.save_Jmu = p.save_Jmu;
  P.nobs = (p.save_peeloff && (p.save_peeloff_2D || p.save_peeloff_3D)) ? p.nobs : 0;
  P.save_peeloff = P.nobs > 0 ? 1 : 0;
  P.save_peeloff_2D = P.save_peeloff && p.save_peeloff_2D; P.save_peeloff_3D = P.save_peeloff && p.save_peeloff_3D;
  P.save_direc0 = p.save_direc0; P.save_all_photons = p.save_all_photons; P.nphotons = p.nphotons;
  h->nobs = P.nobs;
  if (P.nobs > 0) {
    std::vector<DevObserver> ob(P.nobs);
    for (int i = 0; i < P.nobs; ++i) {
      const lart_observer &o = cfg->observers[i];
      if (o.nxim != cfg->observers[0].nxim || o.nyim != cfg->observers[0].nyim || o.nxim < 1 || o.nyim < 1)
        return bail(fail("lart_gpu_create: observers must share one positive image size (par%nxim, par%nyim)"));
      ob[i].x = o.x; ob[i].y = o.y; ob[i].z = o.z;
      for (int k = 0; k < 9; ++k) ob[i].R[k] = o.rmatrix[k];
      ob[i].dxim = o.dxim; ob[i].dyim = o.dyim; ob[i].nxim = o.nxim; ob[i].nyim = o.nyim;
    }
    h->nxim = ob[0].nxim; h->nyim = ob[0].nyim;
    h->obs_host = ob;
    if ((rc = dupload(h, &P.obs, ob.data(), ob.size()))) return bail(rc);
  }
  const lart_scatt_mat &sm = cfg->scatt_mat;
  P.nPDF = (P.dust && P.use_stokes) ? sm.nPDF : 0;
  if (P.nPDF > 0) {
    rc = rc ? rc : dupload(h, &P.sm_coss, sm.coss, sm.nPDF);
    rc = rc ? rc : dupload(h, &P.sm_S11, sm.S11, sm.nPDF);
    rc = rc ? rc : dupload(h, &P.sm_S12, sm.S12, sm.nPDF);
    rc = rc ? rc : dupload(h, &P.sm_S33, sm.S33, sm.nPDF);
    rc = rc ? rc : dupload(h, &P.sm_S34, sm.S34, sm.nPDF);
    rc = rc ? rc : dupload(h, &P.sm_pdf, sm.phase_PDF, sm.nPDF - 1);
    rc = rc ? rc : dupload(h, &P.sm_alias, sm.alias, sm.nPDF - 1);
    if (rc) return bail(rc);
  }
  // ---- tally layout: Jout|Jin|Jabs|Jmu|observer blocks|scalars|counters
  TallyLayout &L = P.lay;
  long long off = 0;
  auto take = [&](bool on, long long n) { long long o = on ? off : -1; if (on) off += n; return o; };
  L.Jout = take(true, g.nxfreq);
  L.Jin = take(P.save_Jin, g.nxfreq);
  L.Jabs = take(P.save_Jabs, g.nxfreq);
  L.Jmu = take(P.save_Jmu, (long long)g.nxfreq * p.nmu);
  L.obs_base = off;
  {
    long long n2 = (long long)h->nxim * h->nyim, n3 = n2 * g.nxfreq, o = 0;
    bool on[7] = {true, true, P.save_direc0 != 0, P.use_stokes != 0, P.use_stokes != 0, P.use_stokes != 0, P.use_stokes != 0};
    for (int k = 0; k < 7; ++k) { L.cube[k] = (P.save_peeloff_3D && on[k]) ? o : -1; if (L.cube[k] >= 0) o += n3; }
    for (int k = 0; k < 7; ++k) { L.img[k] = (P.save_peeloff_2D && on[k]) ? o : -1; if (L.img[k] >= 0) o += n2; }
    L.obs_stride = o;
    off += o * P.nobs;
  }
  L.scalars = take(true, 2);
  L.counters = take(true, C_COUNT);
  {  // behind everything else, so that the common layouts stay as they were
    const long long nb = X.jp ? (g.geometry_JPa == 3 ? (long long)nc : g.geometry_JPa == 2 ? (long long)g.nr * g.nz
                                 : g.geometry_JPa == 1 ? (long long)g.nr : (long long)g.nz) : 0;
    h->jp_bins = nb;
    L.Jabs2 = take(X.atm != 0, g.nxfreq);
    L.J = take(X.calc_J != 0, (long long)g.nxfreq * nb);
    L.Pa = take(X.calc_P != 0, nb);
    L.Pnew = take(X.calc_Pnew != 0, nb);
  }
  lap("pack cells, clumps, observers");
  L.total = off;
  if ((rc = dalloc(h, &P.tally, (size_t)L.total))) return bail(rc);
  // ---- allph: one slot per photon id
  for (int k = 0; k < 10; ++k) { P.allph[k] = nullptr; h->allph_slot[k] = -1; }
  if (P.save_all_photons && p.nphotons > 0) {
    bool on[10] = {p.source_geometry != LART_SRC_POINT, true, true, true, true, true,
                   P.use_stokes != 0, P.use_stokes != 0, P.use_stokes != 0, P.use_stokes != 0};
    int n = 0;
    for (int k = 0; k < 10; ++k) if (on[k]) h->allph_slot[k] = n++;
    h->allph_n = (long long)n * p.nphotons;
    if ((rc = dalloc(h, &h->allph_buf, (size_t)h->allph_n))) return bail(rc);
    for (int k = 0; k < 10; ++k) if (on[k]) P.allph[k] = h->allph_buf + (long long)h->allph_slot[k] * p.nphotons;
  }
  lap("tally + allph buffers");
  // ---- photon pool
  h->flags = cfg->flags;
  // overlapping clump populations run on the one-thread-per-photon driver: the event walk keeps a per-thread event list
  const bool overlap = P.clump && cfg->clumps.has_overlap;
  if (overlap) h->flags |= LART_FLAG_MONOLITHIC;
  const bool mono = (h->flags & LART_FLAG_MONOLITHIC) != 0;
  int S = cfg->pool_slots;
  if (S <= 0 && overlap) S = h->nsm * 1024;  // 32 KB of event list per thread: 148 K photons in flight = 5 GB
  if (S <= 0) S = mono ? (P.clump ? h->nsm * 8192 : h->nsm * 2048) : h->nsm * 16384;  // k_mono_clump: 2.75e8 / 2.85e8 / 2.96e8 scatterings/s at 2048 / 4096 / 8192 per SM
  {
    // keep the ray queue below ~3 GB when many observers are configured
    long long per_slot = (long long)(2 * sizeof(PeelRay) + 4 * sizeof(PeelCont)) * std::max(1, P.nobs);
    long long cap = (8LL << 30) / per_slot;
    if (S > cap) S = (int)std::max<long long>(cap, 1024);
  }
  S = std::max(32, (S + 31) / 32 * 32);
  h->pool.S = S;
  if (overlap) {  // threads of k_mono_clump (one per slot, whole blocks) and of the batch kernels (grid_for: <= 8 blocks per SM)
    const size_t T = std::max<size_t>(((size_t)S + kBlock - 1) / kBlock * kBlock, (size_t)h->nsm * 8 * kBlock);
    P.cl.overlap = 1; P.cl.ov_T = (long long)T;
    rc = rc ? rc : dalloc(h, &P.cl.ov_t, T * kMaxEvt, false);
    rc = rc ? rc : dalloc(h, &P.cl.ov_ev, T * kMaxEvt, false);
    rc = rc ? rc : dalloc(h, &P.cl.ov_act, T * kMaxEvt, false);
    if (rc) return bail(rc);
  }
  // The pool columns, the ray queue and the continuation queues are carved from ONE device allocation: a cold handle
  // pays the driver's per-allocation cost (physical memory creation + mapping, a few ms each) once instead of 25 times —
  // this phase was 99 ms of a cold lart_gpu_create.
  P.max_events = cfg->max_events > 0 ? cfg->max_events : 0;
  h->pool.s0 = 0; h->pool.n = S;
  int G = cfg->streams > 0 ? cfg->streams : 6;
  G = std::max(1, std::min(G, std::min(16, S / 1024 > 0 ? S / 1024 : 1)));
  const long long nobs = P.nobs;
  const long long ray_cap = mono ? 0 : std::min<long long>((long long)S * std::max<long long>(nobs, 1) * 2, 0x7fffffffLL);
  // suspended peel rays (walks longer than the per-wave step budget): room for half of a wave's rays per buffer; a ray that
  // finds its queue full simply keeps walking in this wave (k_wf_peel), so the size is a performance knob, not a limit
  auto cont_cap_of = [&](long long n) { return std::max<long long>(1024, n * std::max<long long>(nobs, 1) / 2) + 32; };
  long long cont_total = 0;
  if (!mono) {
    const int per0 = ((S / G) + 31) / 32 * 32;
    for (int g = 0; g < G; ++g) {
      const int s0 = std::min(g * per0, S), n = (g == G - 1) ? S - s0 : std::min(per0, S - s0);
      cont_total += 2 * cont_cap_of(n);
    }
  }
  unsigned int *ctr = nullptr;
  PeelCont *cont = nullptr;
  auto layout = [&](char *base) -> size_t {  // base = nullptr: sizes only
    size_t off = 0;
    auto carve = [&](auto **p, size_t n) {
      using T = std::remove_pointer_t<std::remove_pointer_t<decltype(p)>>;
      *p = reinterpret_cast<T *>(base + off);
      off += (std::max<size_t>(n, 1) * sizeof(T) + 255) / 256 * 256;
    };
    carve(&h->pool.f, (size_t)F_COUNT * S);
    carve(&h->pool.id, (size_t)S);
    carve(&h->pool.ndraw, (size_t)S);
    carve(&h->pool.ic, (size_t)S);
    carve(&h->pool.jc, (size_t)S);
    carve(&h->pool.kc, (size_t)S);
    carve(&h->pool.flags, (size_t)S);
    carve(&h->pool.rs, (size_t)10 * S);
    carve(&h->pool.rc, (size_t)3 * S);
    carve(&h->pool.nev, (size_t)S);
    if (!mono) {
      carve(&h->pool.var, (size_t)6 * S);
      carve(&h->pool.wtab, (size_t)12 * S);
      carve(&h->pool.lst, (size_t)S);
    }
    carve(&h->job, (size_t)1);
    carve(&h->cmp_src, (size_t)S);
    carve(&h->cmp_dst, (size_t)S);
    carve(&h->cmp_n, (size_t)2);
    if (!mono) {
      carve(&h->rays, (size_t)ray_cap);
      carve(&ctr, 8 * (size_t)G);
      carve(&cont, (size_t)cont_total);
    }
    return off;
  };
  {
    const size_t total = layout(nullptr);
    char *slab = nullptr;
    if ((rc = dalloc(h, &slab, total, false))) return bail(rc);
    CUDA_OK(cudaMemsetAsync(slab, 0, total, h->stream));  // ~1 ms per 3 GB; flags = 0 marks every slot dead
    CUDA_OK(cudaStreamSynchronize(h->stream));
    layout(slab);
  }
  P.err = &h->job->err;
  if (!mono) {
    h->ray_cap = ray_cap;
    h->ctr = ctr; h->ctr_n = 8 * (size_t)G;
    h->groups.resize(G);
    long long cont_off = 0;
    int per = ((S / G) + 31) / 32 * 32;
    for (int g = 0; g < G && !rc; ++g) {
      lart_gpu_ctx::Group &gr = h->groups[g];
      gr.pool = h->pool;
      gr.pool.s0 = std::min(g * per, S);
      gr.pool.n = (g == G - 1) ? S - gr.pool.s0 : std::min(per, S - gr.pool.s0);
      gr.q.rays = h->rays;
      gr.q.n_direct = ctr + 8 * g; gr.q.head_trace = ctr + 8 * g + 1; gr.q.head_peel = ctr + 8 * g + 2;
      gr.q.n_cont = ctr + 8 * g + 3; gr.q.wave = ctr + 8 * g + 5; gr.q.n_dead = ctr + 8 * g + 6;
      if (P.clump) {
        gr.q.direct_base = (unsigned)((long long)S * nobs + (long long)gr.pool.s0 * nobs);
        gr.q.direct_cap = (unsigned)std::min<long long>((long long)gr.pool.n * nobs, ray_cap - gr.q.direct_base);
      } else {
        gr.q.direct_base = (unsigned)(2LL * gr.pool.s0 * nobs);
        gr.q.direct_cap = (unsigned)std::min<long long>(2LL * gr.pool.n * nobs, ray_cap - gr.q.direct_base);
      }
      if (cfg->flags & LART_FLAG_DEBUG_TINY_QUEUES) gr.q.direct_cap = std::min(gr.q.direct_cap, 1u);  // tests of the error path
      gr.q.cont_cap = (unsigned)cont_cap_of(gr.pool.n);
      gr.q.cont[0] = cont + cont_off;
      gr.q.cont[1] = gr.q.cont[0] + gr.q.cont_cap;
      cont_off += 2LL * gr.q.cont_cap;
      CUDA_OK(cudaStreamCreateWithFlags(&gr.stream, cudaStreamNonBlocking));
      CUDA_OK(cudaEventCreateWithFlags(&gr.done, cudaEventDisableTiming));
    }
  }
  if (rc) return bail(rc);
  lap("photon pool + queues");
  h->quantum = cfg->quantum > 0 ? cfg->quantum : (mono ? 32 : 8);
  h->budget = cfg->ray_budget > 0 ? cfg->ray_budget : 32;
  CUDA_OK(cudaStreamSynchronize(h->stream));
  CUDA_OK(cudaGetLastError());
  return 0;
}
}  // namespace

extern "C" {

int lart_gpu_destroy(lart_gpu_handle h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (auto &g : h->groups) if (g.stream) cudaStreamSynchronize(g.stream);
  for (void *p : h->owned) dev_free(p);
  cudaStreamSynchronize(0);
  for (cudaEvent_t e : h->tev) cudaEventDestroy(e);
  for (auto &kv : h->graphs) cudaGraphExecDestroy(kv.second);
  if (h->fork) cudaEventDestroy(h->fork);
  for (auto &g : h->groups) {
    if (g.stream) cudaStreamSynchronize(g.stream);
    for (cudaEvent_t e : g.tev) cudaEventDestroy(e);
    if (g.done) cudaEventDestroy(g.done);
    if (g.stream) cudaStreamDestroy(g.stream);
  }
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
  cudaStreamSynchronize(0);  // the cudaFreeAsync calls above are ordered on the legacy stream
  pool_release(h->device);
  delete h;
  return 0;
}

int lart_gpu_begin(lart_gpu_handle h, int64_t first_id, int64_t count, int64_t stride) {
  if (!h) return fail("lart_gpu_begin: NULL handle");
  if (count < 0 || stride < 1) return fail("lart_gpu_begin: count must be >= 0 and stride >= 1");
  CUDA_OK(cudaSetDevice(h->device));
  Job j{0ULL, (unsigned long long)count, 0ULL, (long long)first_id, (long long)stride, 0u, 0u};
  CUDA_OK(cudaMemcpyAsync(h->job, &j, sizeof(Job), cudaMemcpyHostToDevice, h->stream));
  CUDA_OK(cudaMemsetAsync(h->pool.flags, 0, sizeof(int) * h->pool.S, h->stream));  // abandon unfinished photons
  if (h->ctr) CUDA_OK(cudaMemsetAsync(h->ctr, 0, sizeof(unsigned int) * h->ctr_n, h->stream));  // ... and parked rays
  h->pending_rays = false;
  if (!h->groups.empty() && h->pool.n != h->pool.S) partition_pool(h, h->pool.S);
  for (auto &g : h->groups) {  // every slot is dead now
    const unsigned nd = (unsigned)g.pool.n;
    CUDA_OK(cudaMemcpyAsync(g.q.n_dead, &nd, sizeof(nd), cudaMemcpyHostToDevice, h->stream));
  }
  h->job_next = 0;
  CUDA_OK(cudaStreamSynchronize(h->stream));
  h->count = count;
  h->job_count = count;
  h->more_to_claim = false;
  h->begun = true;
  return 0;
}

}  // extern "C"

namespace {
int drain_peel_only(lart_gpu_handle h);
// the scatter-stage instantiations of this run: k_wf_draw<use_stokes, dust, serial rejection>, k_wf_apply<use_stokes, dust, local steps>
template <class Mark>
int launch_scatter(lart_gpu_handle h, lart_gpu_ctx::Group &g, Mark between) {
  const int nb = (g.pool.n + kBlock - 1) / kBlock;
  const int gd = std::max(1, std::min(nb, h->nsm * LART_DRAW_MINBLOCKS)), ga = std::max(1, std::min(nb, h->nsm * LART_APPLY_MINBLOCKS));
  const int st = h->P.use_stokes ? 4 : 0, du = h->P.dust ? 2 : 0;
#define LART_DR(ST, DU, SE) k_wf_draw<ST, DU, SE><<<gd, kBlock, 0, g.stream>>>(h->P, g.pool)
#define LART_AP(ST, DU, LO) k_wf_apply<ST, DU, LO><<<ga, kBlock, 0, g.stream>>>(h->P, g.pool, h->job, g.q)
#define LART_8(M, v)                                                                                              \
  switch (v) {                                                                                                    \
    case 0: M(false, false, false); break; case 1: M(false, false, true); break; case 2: M(false, true, false); break; \
    case 3: M(false, true, true); break;   case 4: M(true, false, false); break; case 5: M(true, false, true); break;  \
    case 6: M(true, true, false); break;   default: M(true, true, true); break;                                    \
  }
  if (h->draw_mode == 0) {
    // One chunk of slots per warp: as long as the partitions together offer fewer warps than the device holds, chunks stay
    // small (latency); beyond that they grow to kDrawChunkMax, so that the refill loops have many photons per lane.
    const long long G = (long long)h->groups.size(), warps_dev = (long long)h->nsm * LART_DRAW2_MINBLOCKS * (kBlock / 32);
    long long chunk = ((long long)g.pool.n * G / warps_dev + 31) / 32 * 32;
    chunk = std::max<long long>(32, std::min<long long>(chunk, h->draw_chunk));
    const long long warps = (g.pool.n + chunk - 1) / chunk;
    const int gd2 = (int)std::max<long long>(1, (warps + kBlock / 32 - 1) / (kBlock / 32));
    switch (st | du) {
      case 0: k_wf_draw2<false, false><<<gd2, kBlock, 0, g.stream>>>(h->P, g.pool, (int)chunk); break;
      case 2: k_wf_draw2<false, true><<<gd2, kBlock, 0, g.stream>>>(h->P, g.pool, (int)chunk); break;
      case 4: k_wf_draw2<true, false><<<gd2, kBlock, 0, g.stream>>>(h->P, g.pool, (int)chunk); break;
      default: k_wf_draw2<true, true><<<gd2, kBlock, 0, g.stream>>>(h->P, g.pool, (int)chunk); break;
    }
  } else {
    LART_8(LART_DR, st | du | (h->draw_mode == 1 ? 1 : 0))
  }
  if (int rc = between()) return rc;
  LART_8(LART_AP, st | du | (h->P.local_steps ? 1 : 0))
#undef LART_8
#undef LART_AP
#undef LART_DR
  return 0;
}
// sticky device error word -> error return
int check_device_error(unsigned int err) {
  if (!err) return 0;
  std::string m = "device error:";
  if (err & ERR_DIRECT_QUEUE) m += " peel-ray queue overflow (direct rays were lost);";
  if (err & ERR_CONT_QUEUE) m += " continuation queue overflow;";
  if (err & ERR_BAD_STATE) m += " impossible photon state;";
  return fail("lart_gpu: " + m + " the tallies of this run are incomplete");
}
// One step with an explicit driver choice (both drivers share the pool layout, and no
// slot is left mid-wave between steps, so they can alternate freely).
int step_impl(lart_gpu_handle h, int qn, bool mono, int64_t *in_flight, bool defer_peel = false) {
  const bool timing = (h->flags & LART_FLAG_STAGE_TIMING) != 0;
  auto mark = [&](std::vector<cudaEvent_t> &tev, size_t &ne, cudaStream_t st) -> int {  // one event between stage kernels
    if (!timing) return 0;
    if (ne == tev.size()) {
      cudaEvent_t e;
      CUDA_OK(cudaEventCreate(&e));
      tev.push_back(e);
    }
    CUDA_OK(cudaEventRecord(tev[ne++], st));
    return 0;
  };
  CUDA_OK(cudaEventRecord(h->ev0, h->stream));
  if (mono) {
    size_t ne = 0;
    if (int rc = mark(h->tev, ne, h->stream)) return rc;
    if (h->P.clump) k_mono_clump<<<(h->pool.n + kBlock - 1) / kBlock, kBlock, 0, h->stream>>>(h->P, h->pool, h->job, qn);
    else if (defer_peel) {
      // tail of a run: the ray queue is empty (drained) and all of it belongs to this launch
      Queues q = h->groups[0].q;
      q.direct_base = 0;
      q.direct_cap = (unsigned)h->ray_cap;
      k_wf_reset<<<1, 1, 0, h->stream>>>(q);
      if (is_plain(h->P)) k_mono<true, true><<<(h->pool.n + kBlock - 1) / kBlock, kBlock, 0, h->stream>>>(h->P, h->pool, h->job, qn, q);
      else k_mono<true, false><<<(h->pool.n + kBlock - 1) / kBlock, kBlock, 0, h->stream>>>(h->P, h->pool, h->job, qn, q);
      if (is_plain(h->P)) k_wf_peel<true><<<std::max(1, h->nsm * LART_PEEL_MINBLOCKS), kBlock, 0, h->stream>>>(h->P, h->pool, q, 0x7fffffff, 0);
      else k_wf_peel<false><<<std::max(1, h->nsm * LART_PEEL_MINBLOCKS), kBlock, 0, h->stream>>>(h->P, h->pool, q, 0x7fffffff, 0);
      h->launches += 2;
    } else if (is_plain(h->P)) k_mono<false, true><<<(h->pool.n + kBlock - 1) / kBlock, kBlock, 0, h->stream>>>(h->P, h->pool, h->job, qn, Queues{});
    else k_mono<false, false><<<(h->pool.n + kBlock - 1) / kBlock, kBlock, 0, h->stream>>>(h->P, h->pool, h->job, qn, Queues{});
    if (int rc = mark(h->tev, ne, h->stream)) return rc;
    h->launches += 1;
  } else {
    // Every pool partition is an independent wave pipeline on its own stream: while one partition waits for
    // the longest ray of its trace/peel kernel, the others keep the SMs busy.  The whole step (all partitions,
    // all waves: 5*G*qn launches) is captured once into a CUDA graph, so a step costs one graph launch.
    const int G = (int)h->groups.size();
    auto issue = [&](bool with_marks) -> int {
      std::vector<size_t> ne(G, 0);
      for (int w = 0; w < qn; ++w) {
        for (int gi = 0; gi < G; ++gi) {
          lart_gpu_ctx::Group &g = h->groups[gi];
          const int nb = (g.pool.n + kBlock - 1) / kBlock;
          const int grid = std::max(1, std::min(nb, h->nsm * 2));
          k_wf_reset<<<1, 1, 0, g.stream>>>(g.q);
          if (with_marks) if (int rc = mark(g.tev, ne[gi], g.stream)) return rc;
          if (h->P.clump) k_cl_emit<<<grid, kBlock, 0, g.stream>>>(h->P, g.pool, h->job, g.q);
          else k_wf_emit<<<grid, kBlock, 0, g.stream>>>(h->P, g.pool, h->job, g.q);
          if (with_marks) if (int rc = mark(g.tev, ne[gi], g.stream)) return rc;
          if (h->P.clump) k_cl_flight<<<std::max(1, std::min(nb, h->nsm * 3)), kBlock, 0, g.stream>>>(h->P, g.pool, h->job, g.q);
          else if (is_plain(h->P)) k_wf_trace<true><<<std::max(1, std::min(nb, h->nsm * LART_TRACE_MINBLOCKS)), kBlock, 0, g.stream>>>(h->P, g.pool, h->job, g.q, h->budget);
          else k_wf_trace<false><<<std::max(1, std::min(nb, h->nsm * LART_TRACE_MINBLOCKS)), kBlock, 0, g.stream>>>(h->P, g.pool, h->job, g.q, h->budget);
          if (with_marks) if (int rc = mark(g.tev, ne[gi], g.stream)) return rc;
          auto between = [&]() -> int { return with_marks ? mark(g.tev, ne[gi], g.stream) : 0; };
          if (h->P.clump) {
            k_cl_scatter<<<grid, kBlock, 0, g.stream>>>(h->P, g.pool, h->job, g.q);
            if (int rc = between()) return rc;
          } else if (int rc = launch_scatter(h, g, between)) return rc;
          if (with_marks) if (int rc = mark(g.tev, ne[gi], g.stream)) return rc;
          if (h->P.nobs == 0) {}  // no observers: nothing to peel (xyz_symmetry, plain slabs)
          else if (h->P.clump) k_cl_peel<<<std::max(1, std::min(nb, h->nsm * 4)), kBlock, 0, g.stream>>>(h->P, g.pool, g.q);
          else if (is_plain(h->P)) k_wf_peel<true><<<std::max(1, std::min(nb, h->nsm * LART_PEEL_MINBLOCKS)), kBlock, 0, g.stream>>>(h->P, g.pool, g.q, h->budget, 0);
          else k_wf_peel<false><<<std::max(1, std::min(nb, h->nsm * LART_PEEL_MINBLOCKS)), kBlock, 0, g.stream>>>(h->P, g.pool, g.q, h->budget, 0);
          if (with_marks) if (int rc = mark(g.tev, ne[gi], g.stream)) return rc;
        }
      }
      return 0;
    };
    if (timing) {  // per-stage events: plain launches
      for (auto &g : h->groups) CUDA_OK(cudaStreamWaitEvent(g.stream, h->ev0, 0));
      if (int rc = issue(true)) return rc;
      for (auto &g : h->groups) {
        CUDA_OK(cudaEventRecord(g.done, g.stream));
        CUDA_OK(cudaStreamWaitEvent(h->stream, g.done, 0));
      }
    } else {
      auto it = h->graphs.find(qn);
      if (it == h->graphs.end()) {
        cudaGraph_t graph = nullptr;
        CUDA_OK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        CUDA_OK(cudaEventRecord(h->fork, h->stream));
        for (auto &g : h->groups) CUDA_OK(cudaStreamWaitEvent(g.stream, h->fork, 0));
        int rc = issue(false);
        for (auto &g : h->groups) {
          cudaEventRecord(g.done, g.stream);
          cudaStreamWaitEvent(h->stream, g.done, 0);
        }
        cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
        if (rc) return rc;
        if (ce != cudaSuccess) return fail(std::string("cudaStreamEndCapture failed: ") + cudaGetErrorString(ce));
        cudaGraphExec_t exec = nullptr;
        CUDA_OK(cudaGraphInstantiate(&exec, graph, 0));
        CUDA_OK(cudaGraphDestroy(graph));
        it = h->graphs.emplace(qn, exec).first;
      }
      CUDA_OK(cudaGraphLaunch(it->second, h->stream));
    }
    h->launches += (h->P.nobs == 0 ? 5LL : 6LL) * G * qn;
    h->pending_rays = true;
  }
  CUDA_OK(cudaEventRecord(h->ev1, h->stream));
  Job j;
  CUDA_OK(cudaMemcpyAsync(&j, h->job, sizeof(Job), cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  CUDA_OK(cudaGetLastError());
  float ms = 0.f;
  CUDA_OK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->kernel_ms += ms;
  if (timing) {
    if (mono) {
      float t = 0.f;
      CUDA_OK(cudaEventElapsedTime(&t, h->tev[0], h->tev[1]));
      h->stage_ms[LART_STAGE_TRACE] += t;
      h->stage_n[LART_STAGE_TRACE] += 1;
    } else {
      for (auto &g : h->groups)
        for (int w = 0; w < qn; ++w)
          for (int k = 0; k < LART_STAGE_COUNT; ++k) {
            float t = 0.f;
            CUDA_OK(cudaEventElapsedTime(&t, g.tev[(LART_STAGE_COUNT + 1) * w + k], g.tev[(LART_STAGE_COUNT + 1) * w + k + 1]));
            h->stage_ms[k] += t;
            h->stage_n[k] += 1;
          }
    }
  }
  h->job_next = j.next;
  if (in_flight) *in_flight = (int64_t)h->count - (int64_t)j.done;
  return check_device_error(j.err);
}

// Split the live range [0, n) of the pool among the wave pipelines (32-slot granularity).
void partition_pool(lart_gpu_handle h, int n) {
  const int G = (int)h->groups.size();
  h->pool.s0 = 0;
  h->pool.n = n;
  const int per = ((n + G - 1) / G + 31) / 32 * 32;
  for (int g = 0; g < G; ++g) {
    Pool &p = h->groups[g].pool;
    p.s0 = std::min(g * per, n);
    p.n = std::min(per, n - p.s0);
    if (!h->P.clump) {  // the partition's region of the ray queue moves with it (between waves the queue is empty)
      Queues &q = h->groups[g].q;
      const long long nobs = h->P.nobs;
      const bool tiny = (h->flags & LART_FLAG_DEBUG_TINY_QUEUES) != 0;
      q.direct_base = (unsigned)(2LL * p.s0 * nobs);
      q.direct_cap = tiny ? 1u : (unsigned)(2LL * p.n * nobs);
    }
  }
  for (auto &kv : h->graphs) cudaGraphExecDestroy(kv.second);  // kernel arguments changed: re-capture
  h->graphs.clear();
}

// Compact the live photons into [0, alive) when the queue is empty and the pool is less than half full.
int maybe_compact(lart_gpu_handle h, int64_t alive) {
  if (h->groups.empty() || h->job_next < (unsigned long long)h->job_count || h->more_to_claim) return 0;
  if (alive * 2 > h->pool.n || h->pool.n <= 4096) return 0;
  if (int rc = drain_peel_only(h)) return rc;
  const int n_keep = (int)std::max<int64_t>(1024, (alive + 31) / 32 * 32);
  CUDA_OK(cudaMemsetAsync(h->cmp_n, 0, 2 * sizeof(unsigned int), h->stream));
  const int grid = std::max(1, std::min((h->pool.n + kBlock - 1) / kBlock, h->nsm * 8));
  k_compact_list<<<grid, kBlock, 0, h->stream>>>(h->pool, n_keep, h->cmp_src, h->cmp_dst, h->cmp_n, h->cmp_n + 1);
  k_compact_move<<<grid, kBlock, 0, h->stream>>>(h->pool, h->cmp_src, h->cmp_dst, h->cmp_n);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(h->stream));
  h->launches += 2;
  partition_pool(h, n_keep);
  return 0;
}
}  // namespace

extern "C" {

int lart_gpu_step(lart_gpu_handle h, int32_t quantum, int64_t *in_flight) {
  if (!h) return fail("lart_gpu_step: NULL handle");
  if (!h->begun) return fail("lart_gpu_step: call lart_gpu_begin first");
  CUDA_OK(cudaSetDevice(h->device));
  const bool mono = (h->flags & LART_FLAG_MONOLITHIC) != 0;
  return step_impl(h, quantum > 0 ? quantum : h->quantum, mono, in_flight);
}

}  // extern "C"

namespace {
int drain(lart_gpu_handle h, bool flights);
int drain_peel_only(lart_gpu_handle h) { return drain(h, false); }
// Finish every walk that was parked at its per-wave step budget: peel rays always; photon flights too when
// `flights` (before the monolithic kernel takes over).  One pass with an unlimited budget does it.
int drain(lart_gpu_handle h, bool flights) {
  if (!h->pending_rays || h->groups.empty()) return 0;
  if (h->P.clump) { h->pending_rays = false; return 0; }  // the clump walkers never park a ray
  const int big = 0x7fffffff;
  for (auto &g : h->groups) {
    const int nb = (g.pool.n + kBlock - 1) / kBlock;
    const int grid = std::max(1, std::min(nb, h->nsm * 2));
    k_wf_reset<<<1, 1, 0, h->stream>>>(g.q);
    if (flights) {
      if (is_plain(h->P)) k_wf_trace<true><<<grid, kBlock, 0, h->stream>>>(h->P, g.pool, h->job, g.q, big);
      else k_wf_trace<false><<<grid, kBlock, 0, h->stream>>>(h->P, g.pool, h->job, g.q, big);
    }
    if (is_plain(h->P)) k_wf_peel<true><<<grid, kBlock, 0, h->stream>>>(h->P, g.pool, g.q, big, 1);
    else k_wf_peel<false><<<grid, kBlock, 0, h->stream>>>(h->P, g.pool, g.q, big, 1);
    h->launches += flights ? 3 : 2;
  }
  CUDA_OK(cudaGetLastError());
  if (flights) h->pending_rays = false;  // (peel-only drains leave parked flights: keep the flag)
  return 0;
}
}  // namespace

extern "C" {

int lart_gpu_sync(lart_gpu_handle h) {
  if (!h) return fail("lart_gpu_sync: NULL handle");
  CUDA_OK(cudaSetDevice(h->device));
  if (int rc = drain(h, false)) return rc;  // every peel ray emitted so far is deposited
  unsigned int err = 0;
  CUDA_OK(cudaMemcpyAsync(&err, &h->job->err, sizeof(err), cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  CUDA_OK(cudaGetLastError());
  return check_device_error(err);
}

}  // extern "C"
namespace {
// the job queue's range is used up: hand the device the next range of photon ids; photons in flight are untouched
int extend_job(lart_gpu_handle h, int64_t first_id, int64_t count) {
  const unsigned long long head[2] = {0ULL, (unsigned long long)count};  // Job::next, Job::count
  const long long first = (long long)first_id, stride = 1;
  CUDA_OK(cudaMemcpyAsync(h->job, head, sizeof(head), cudaMemcpyHostToDevice, h->stream));
  CUDA_OK(cudaMemcpyAsync(&h->job->first_id, &first, sizeof(first), cudaMemcpyHostToDevice, h->stream));
  CUDA_OK(cudaMemcpyAsync(&h->job->stride, &stride, sizeof(stride), cudaMemcpyHostToDevice, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  h->job_next = 0;
  h->job_count = count;
  h->count += count;
  return 0;
}
// claim(first_id, count): the next range of photon ids for this handle, false when the node has none left (dynamic dealing);
// null: the ids queued by lart_gpu_begin are all there is
int run_to_end(lart_gpu_handle h, int64_t left, const std::function<bool(int64_t &, int64_t &)> &claim) {
  const bool mono = (h->flags & LART_FLAG_MONOLITHIC) != 0;
  long long tail_photons = kTailPhotons;
  if (const char *e = getenv("LART_GPU_TAIL_PHOTONS")) tail_photons = atoll(e);
  const char *pe = getenv("LART_GPU_PROGRESS");  // LART_GPU_PROGRESS=n: a line every n steps (default 64)
  const bool progress = pe != nullptr;
  const int pevery = pe && atoi(pe) > 0 ? atoi(pe) : 64;
  long long nstep = 0;
  h->more_to_claim = claim != nullptr;
  while (left > 0 || h->more_to_claim) {
    if (h->more_to_claim && h->job_next >= (unsigned long long)h->job_count) {  // the queue ran dry: ask for more work
      int64_t first = 0, cnt = 0;
      if (claim(first, cnt)) { if (int rc = extend_job(h, first, cnt)) return rc; left += cnt; }
      else h->more_to_claim = false;
      if (left <= 0) continue;
    }
    const bool dry = h->job_next >= (unsigned long long)h->job_count && !h->more_to_claim;
    // Heavy tail: once the queue is empty the pool thins out; keep it dense (compaction), and below
    // kTailPhotons let one thread per photon run `quantum` scatterings per launch — a wave is then bound by
    // launch latency (one scattering per ~5 launches), not by throughput.
    if (!mono) if (int rc = maybe_compact(h, left)) return rc;
    const bool tail = !mono && left < tail_photons && dry;
    if (tail) if (int rc = drain(h, true)) return rc;  // the monolithic kernel cannot resume parked walks
    // tail steps: one thread per photon runs `tq` scatterings and queues their peel rays, the peel stage walks them afterwards;
    // tq is what the ray queue holds (one ray per scattering and observer)
    int tq = kTailQuantum;
    const bool defer = tail && !h->P.clump && h->P.nobs > 0 && h->ray_cap > 0 && !getenv("LART_GPU_TAIL_INLINE");
    if (defer) tq = (int)std::max<long long>(1, std::min<long long>(kTailQuantum, h->ray_cap / ((long long)std::max(h->pool.n, 1) * h->P.nobs)));
    if (int rc = step_impl(h, tail ? tq : h->quantum, mono || tail, &left, defer)) return rc;
    if (progress && (++nstep % pevery == 0 || left == 0))  // LART_GPU_PROGRESS=1: like the reference's nprint lines
      fprintf(stderr, "lart_gpu_run: %lld of %lld photons left, pool range %d, driver %s, device time %.3f s\n", (long long)left,
              (long long)h->count, h->pool.n, (mono || tail) ? "monolithic" : "wavefront", h->kernel_ms * 1e-3);
  }
  return lart_gpu_sync(h);
}
}  // namespace
extern "C" {

int lart_gpu_run(lart_gpu_handle h, int64_t first_id, int64_t count, int64_t stride) {
  if (int rc = lart_gpu_begin(h, first_id, count, stride)) return rc;
  return run_to_end(h, count, nullptr);
}

/* ---- dynamic photon dealing on one node (src/run_simulation_mod.f90:31-128: master/worker, batches of num_send_at_once) ----
 * The node's processes share ONE 64-bit counter in POSIX shared memory; a process whose job queue runs dry claims the next
 * `batch` photon ids with an atomic fetch-add — no master rank, no messages.  Streams are keyed by photon id, so the sum over
 * the processes does not depend on who ran which photon. */
struct lart_gpu_deal {
  int fd = -1;
  volatile long long *next = nullptr;
  std::string name;
};
int lart_gpu_deal_open(const char *name, int32_t reset, lart_gpu_deal_handle *out) {
  if (!name || !out || name[0] != '/') return fail("lart_gpu_deal_open: name must be a POSIX shared-memory name ('/...')");
  *out = nullptr;
  int fd = shm_open(name, O_CREAT | O_RDWR, 0600);
  if (fd < 0) return fail(std::string("lart_gpu_deal_open: shm_open failed: ") + strerror(errno));
  if (ftruncate(fd, 64) != 0) { close(fd); return fail("lart_gpu_deal_open: ftruncate failed"); }
  void *p = mmap(nullptr, 64, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  if (p == MAP_FAILED) { close(fd); return fail("lart_gpu_deal_open: mmap failed"); }
  lart_gpu_deal *d = new lart_gpu_deal();
  d->fd = fd; d->next = (volatile long long *)p; d->name = name;
  if (reset) __atomic_store_n(d->next, 0LL, __ATOMIC_SEQ_CST);
  *out = d;
  return 0;
}
int lart_gpu_deal_close(lart_gpu_deal_handle d, int32_t unlink_name) {
  if (!d) return 0;
  if (d->next) munmap((void *)d->next, 64);
  if (d->fd >= 0) close(d->fd);
  if (unlink_name) shm_unlink(d->name.c_str());
  delete d;
  return 0;
}
int lart_gpu_deal_claim(lart_gpu_deal_handle d, int64_t nphotons, int64_t batch, int64_t *first_id, int64_t *count) {
  if (!d || !first_id || !count) return fail("lart_gpu_deal_claim: NULL argument");
  if (nphotons < 0 || batch < 1) return fail("lart_gpu_deal_claim: nphotons must be >= 0 and batch >= 1");
  const long long a = __atomic_fetch_add(d->next, (long long)batch, __ATOMIC_SEQ_CST);
  *first_id = a + 1;  // photon ids are 1-based
  *count = a >= nphotons ? 0 : std::min<int64_t>(batch, nphotons - a);
  return 0;
}
int lart_gpu_run_dealt(lart_gpu_handle h, lart_gpu_deal_handle d, int64_t nphotons, int64_t batch, int64_t *nclaimed) {
  if (!h || !d) return fail("lart_gpu_run_dealt: NULL argument");
  if (nphotons < 0 || batch < 1) return fail("lart_gpu_run_dealt: nphotons must be >= 0 and batch >= 1");
  int64_t mine = 0;
  auto claim = [&](int64_t &first, int64_t &cnt) {
    if (lart_gpu_deal_claim(d, nphotons, batch, &first, &cnt) != 0 || cnt == 0) return false;
    mine += cnt;
    return true;
  };
  if (int rc = lart_gpu_begin(h, 1, 0, 1)) return rc;  // an empty queue: the first claim fills it
  const int rc = run_to_end(h, 0, claim);
  if (nclaimed) *nclaimed = mine;
  return rc;
}

int lart_gpu_reset_tallies(lart_gpu_handle h) {
  if (!h) return fail("lart_gpu_reset_tallies: NULL handle");
  CUDA_OK(cudaSetDevice(h->device));
  CUDA_OK(cudaMemsetAsync(h->P.tally, 0, sizeof(double) * h->P.lay.total, h->stream));
  if (h->allph_buf) CUDA_OK(cudaMemsetAsync(h->allph_buf, 0, sizeof(double) * h->allph_n, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  h->kernel_ms = 0.0;
  h->launches = 0;
  for (int k = 0; k < LART_STAGE_COUNT; ++k) { h->stage_ms[k] = 0.0; h->stage_n[k] = 0; }
  return 0;
}

int lart_gpu_tally_buffer(lart_gpu_handle h, void **dev_ptr, int64_t *n_doubles) {
  if (!h || !dev_ptr || !n_doubles) return fail("lart_gpu_tally_buffer: NULL argument");
  *dev_ptr = h->P.tally;
  *n_doubles = h->P.lay.total;
  return 0;
}
int lart_gpu_allph_buffer(lart_gpu_handle h, void **dev_ptr, int64_t *n_doubles) {
  if (!h || !dev_ptr || !n_doubles) return fail("lart_gpu_allph_buffer: NULL argument");
  *dev_ptr = h->allph_buf;
  *n_doubles = h->allph_n;
  return 0;
}
int lart_gpu_stream(lart_gpu_handle h, void **stream) {
  if (!h || !stream) return fail("lart_gpu_stream: NULL argument");
  *stream = (void *)h->stream;
  return 0;
}
int lart_gpu_kernel_ms(lart_gpu_handle h, double *ms, int64_t *launches) {
  if (!h) return fail("lart_gpu_kernel_ms: NULL handle");
  if (ms) *ms = h->kernel_ms;
  if (launches) *launches = h->launches;
  return 0;
}

int lart_gpu_stage_ms(lart_gpu_handle h, double ms[LART_STAGE_COUNT], int64_t launches[LART_STAGE_COUNT]) {
  if (!h) return fail("lart_gpu_stage_ms: NULL handle");
  for (int k = 0; k < LART_STAGE_COUNT; ++k) {
    if (ms) ms[k] = h->stage_ms[k];
    if (launches) launches[k] = h->stage_n[k];
  }
  return 0;
}
int lart_gpu_pool_slots(lart_gpu_handle h, int64_t *slots) {
  if (!h || !slots) return fail("lart_gpu_pool_slots: NULL argument");
  *slots = h->pool.S;
  return 0;
}

int lart_gpu_measure_fp64(int32_t device, double *tflops) {
  if (!tflops) return fail("lart_gpu_measure_fp64: NULL argument");
  CUDA_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, device));
  double *out = nullptr;
  CUDA_OK(cudaMalloc(&out, sizeof(double)));
  cudaEvent_t e0, e1;
  CUDA_OK(cudaEventCreate(&e0));
  CUDA_OK(cudaEventCreate(&e1));
  const int iters = 4096, blocks = prop.multiProcessorCount * 8, threads = 256;
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    CUDA_OK(cudaEventRecord(e0));
    k_dfma_peak<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    CUDA_OK(cudaEventRecord(e1));
    CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    double fl = 2.0 * 16.0 * (double)iters * blocks * threads;  // 16 independent FMA chains per thread
    best = std::max(best, fl / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  CUDA_OK(cudaGetLastError());
  *tflops = best;
  return 0;
}

int lart_gpu_fetch(lart_gpu_handle h, lart_tallies *out) {
  if (!h || !out) return fail("lart_gpu_fetch: NULL argument");
  CUDA_OK(cudaSetDevice(h->device));
  const DevParams &P = h->P;
  const TallyLayout &L = P.lay;
  if (P.nobs > 0 && !out->obs) return fail("lart_gpu_fetch: out->obs is NULL but observers are configured");
  if (int rc = lart_gpu_sync(h)) return rc;  // parked peel rays are deposited; a sticky device error is an error here too
  // destination segments of the contiguous device buffer
  std::vector<Seg> segs;
  auto seg = [&](double *dst, long long off, long long n) { if (dst && off >= 0 && n > 0) segs.push_back({dst, off, n}); };
  seg(out->Jout, L.Jout, P.nxfreq);
  seg(out->Jin, L.Jin, P.nxfreq);
  seg(out->Jabs, L.Jabs, P.nxfreq);
  seg(out->Jmu, L.Jmu, (long long)P.nxfreq * P.nmu);
  const long long n2 = (long long)h->nxim * h->nyim, n3 = n2 * P.nxfreq;
  for (int k = 0; k < P.nobs; ++k) {
    lart_observer_out &o = out->obs[k];
    const long long base = L.obs_base + (long long)k * L.obs_stride;
    double *c3[7] = {o.scatt, o.direc, o.direc0, o.I, o.Q, o.U, o.V};
    double *c2[7] = {o.scatt_2D, o.direc_2D, o.direc0_2D, o.I_2D, o.Q_2D, o.U_2D, o.V_2D};
    for (int q = 0; q < 7; ++q) {
      if (L.cube[q] >= 0) seg(c3[q], base + L.cube[q], n3);
      if (L.img[q] >= 0) seg(c2[q], base + L.img[q], n2);
    }
  }
  seg(out->Jabs2, L.Jabs2, P.nxfreq);
  seg(out->J, L.J, (long long)P.nxfreq * h->jp_bins);
  seg(out->Pa, L.Pa, h->jp_bins);
  seg(out->Pnew, L.Pnew, h->jp_bins);
  double tail[2 + C_COUNT];
  segs.push_back({tail, L.scalars, 2 + C_COUNT});  // scalars and counters are adjacent in the layout
  for (double &v : tail) v = 0.0;
  if (int rc = add_device_to_host(h, P.tally, L.total, segs)) return rc;
  out->nscatt_gas += tail[0];
  out->nscatt_dust += tail[1];
  const double *c = tail + 2;
  out->counters.n_photons_done += c[C_PHOTONS]; out->counters.n_scatter += c[C_SCATTER];
  out->counters.n_cellsteps += c[C_CELLSTEPS]; out->counters.n_peel += c[C_PEEL];
  out->counters.n_rng += c[C_RNG]; out->counters.n_reject_iter += c[C_REJECT];
  out->counters.n_peel_bound += c[C_PEEL_BOUND]; out->counters.n_cellsteps_bound += c[C_PEEL_BOUND];  // one step per such ray
  if (h->allph_buf) {
    std::vector<Seg> as;
    double *dst[10] = {out->allph.rp0, out->allph.rp, out->allph.xfreq1, out->allph.xfreq2, out->allph.nscatt_gas,
                       out->allph.nscatt_dust, out->allph.I, out->allph.Q, out->allph.U, out->allph.V};
    for (int k = 0; k < 10; ++k)
      if (h->allph_slot[k] >= 0 && dst[k]) as.push_back({dst[k], (long long)h->allph_slot[k] * P.nphotons, P.nphotons});
    if (int rc = add_device_to_host(h, h->allph_buf, h->allph_n, as)) return rc;
  }
  return 0;
}

/* ---- multi-GPU: one process per GPU, one NCCL communicator per process -------------------------------------------
 * Replaces the communicator half of memory_mod_mpi.f90:366-458 / output_sum_rect.f90:13-146: ONE ncclReduce over the
 * contiguous tally buffer (and one over the allph buffer) instead of a per-array, plane-by-plane MPI_REDUCE of host
 * arrays.  libnccl is opened at run time, so the library loads (and single-GPU runs work) where NCCL is absent. */
int lart_gpu_comm_unique_id(void *id128) {
  if (!id128) return fail("lart_gpu_comm_unique_id: NULL argument");
  if (int rc = nccl_load()) return rc;
  ncclUniqueId id;
  NCCL_OK(g_nccl.GetUniqueId(&id));
  static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
  memcpy(id128, &id, sizeof(id));
  return 0;
}

int lart_gpu_comm_init(int32_t device, int32_t nranks, int32_t rank, const void *id128) {
  if (!id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail("lart_gpu_comm_init: bad argument");
  if (g_comm) return fail("lart_gpu_comm_init: the process already has a communicator (lart_gpu_comm_finalize first)");
  if (int rc = nccl_load()) return rc;
  CUDA_OK(cudaSetDevice(device));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  NCCL_OK(g_nccl.CommInitRank(&g_comm, nranks, id, rank));
  g_comm_rank = rank; g_comm_size = nranks; g_comm_device = device;
  // NCCL connects its channels at the first collective of each kind (seconds on an NVSwitch box): do that here, in the
  // MPI_INIT of the GPU path, with a latency-sized and a bandwidth-sized reduce, not inside the first output_reduce.
  if (nranks > 1) {
    double *buf = nullptr;
    const size_t big = (size_t)4 << 20;
    CUDA_OK(cudaMalloc(&buf, big * sizeof(double)));
    CUDA_OK(cudaMemset(buf, 0, big * sizeof(double)));
    cudaStream_t st;
    CUDA_OK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    NCCL_OK(g_nccl.Reduce(buf, buf, 8, ncclDouble, ncclSum, 0, g_comm, st));
    NCCL_OK(g_nccl.Reduce(buf, buf, big, ncclDouble, ncclSum, 0, g_comm, st));
    CUDA_OK(cudaStreamSynchronize(st));
    cudaStreamDestroy(st);
    cudaFree(buf);
  }
  return 0;
}

int lart_gpu_comm_finalize(void) {
  if (g_comm) { g_nccl.CommDestroy(g_comm); g_comm = nullptr; }
  g_comm_rank = 0; g_comm_size = 1; g_comm_device = -1;
  return 0;
}

int lart_gpu_comm_info(int32_t *nranks, int32_t *rank) {
  if (nranks) *nranks = g_comm ? g_comm_size : 1;
  if (rank) *rank = g_comm ? g_comm_rank : 0;
  return 0;
}

int lart_gpu_reduce(lart_gpu_handle h, int32_t root) {
  if (!h) return fail("lart_gpu_reduce: NULL handle");
  CUDA_OK(cudaSetDevice(h->device));
  if (int rc = lart_gpu_sync(h)) return rc;
  if (!g_comm || g_comm_size == 1) return root == 0 ? 0 : fail("lart_gpu_reduce: no communicator, root must be 0");
  if (root < 0 || root >= g_comm_size) return fail("lart_gpu_reduce: root outside the communicator");
  if (h->device != g_comm_device) return fail("lart_gpu_reduce: the handle lives on another device than the communicator");
  NCCL_OK(g_nccl.GroupStart());
  NCCL_OK(g_nccl.Reduce(h->P.tally, h->P.tally, (size_t)h->P.lay.total, ncclDouble, ncclSum, root, g_comm, h->stream));
  if (h->allph_buf) NCCL_OK(g_nccl.Reduce(h->allph_buf, h->allph_buf, (size_t)h->allph_n, ncclDouble, ncclSum, root, g_comm, h->stream));
  NCCL_OK(g_nccl.GroupEnd());
  if (g_comm_rank != root) {  // this rank's contribution now lives on root: the sum over ranks stays "everything so far"
    CUDA_OK(cudaMemsetAsync(h->P.tally, 0, sizeof(double) * h->P.lay.total, h->stream));
    if (h->allph_buf) CUDA_OK(cudaMemsetAsync(h->allph_buf, 0, sizeof(double) * h->allph_n, h->stream));
  }
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return 0;
}

}  // extern "C"

// ---------------------------- batched plugin points -------------------------
namespace {
}  // namespace

extern "C" {

int lart_gpu_voigt_batch(int64_t n, const double *x, const double *a, double *H) {
  if (n < 0 || (n > 0 && (!x || !a || !H))) return fail("lart_gpu_voigt_batch: bad argument");
  if (n == 0) return 0;
  int dev = 0;
  CUDA_OK(cudaGetDevice(&dev));
  double *tab;
  if (int rc = device_vtab(dev, &tab)) return rc;
  Scratch s;
  double *dx, *da, *dH;
  if (int rc = s.in(&dx, x, n)) return rc;
  if (int rc = s.in(&da, a, n)) return rc;
  if (int rc = s.outbuf(&dH, n)) return rc;
  k_voigt_batch<<<grid_for(n, 148), kBlock>>>(tab, n, dx, da, dH);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaDeviceSynchronize());
  D2H(H, dH, n);
  return 0;
}

int lart_gpu_raytrace_edge_batch(lart_gpu_handle h, int64_t n, const double *x, const double *y, const double *z,
                                 const double *kx, const double *ky, const double *kz, const double *xfreq,
                                 const int32_t *icell, const int32_t *jcell, const int32_t *kcell, double *tau,
                                 int32_t *nsteps, int32_t trace_cap, int32_t *trace_cells) {
  if (!h) return fail("lart_gpu_raytrace_edge_batch: NULL handle");
  if (n < 0 || (n > 0 && (!x || !y || !z || !kx || !ky || !kz || !xfreq || !icell || !jcell || !kcell || !tau)))
    return fail("lart_gpu_raytrace_edge_batch: bad argument");
  if (n == 0) return 0;
  CUDA_OK(cudaSetDevice(h->device));
  Scratch s;
  double *dx, *dy, *dz, *dkx, *dky, *dkz, *dxf, *dtau;
  int *dic, *djc, *dkc, *dns, *dtr = nullptr;
  int rc = 0;
  rc = rc ? rc : s.in(&dx, x, n); rc = rc ? rc : s.in(&dy, y, n); rc = rc ? rc : s.in(&dz, z, n);
  rc = rc ? rc : s.in(&dkx, kx, n); rc = rc ? rc : s.in(&dky, ky, n); rc = rc ? rc : s.in(&dkz, kz, n);
  rc = rc ? rc : s.in(&dxf, xfreq, n);
  rc = rc ? rc : s.in(&dic, icell, n); rc = rc ? rc : s.in(&djc, jcell, n); rc = rc ? rc : s.in(&dkc, kcell, n);
  rc = rc ? rc : s.outbuf(&dtau, n); rc = rc ? rc : s.outbuf(&dns, n);
  if (!rc && trace_cells && trace_cap > 0) {
    rc = s.outbuf(&dtr, (size_t)n * trace_cap);
    if (!rc) CUDA_OK(cudaMemset(dtr, 0xff, sizeof(int) * (size_t)n * trace_cap));
  }
  if (rc) return rc;
  k_edge_batch<<<grid_for(n, h->nsm), kBlock, 0, h->stream>>>(h->P, n, dx, dy, dz, dkx, dky, dkz, dxf, dic, djc, dkc, dtau, dns,
                                                             dtr ? trace_cap : 0, dtr);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(h->stream));
  D2H(tau, dtau, n);
  if (nsteps) D2H(nsteps, dns, n);
  if (dtr) D2H(trace_cells, dtr, (size_t)n * trace_cap);
  return 0;
}

int lart_gpu_raytrace_tau_batch(lart_gpu_handle h, int64_t n, double *x, double *y, double *z, const double *kx,
                                const double *ky, const double *kz, double *xfreq, int32_t *icell, int32_t *jcell,
                                int32_t *kcell, const double *tau_in, int32_t *inside, double *xfreq_ref, int32_t *nsteps) {
  if (!h) return fail("lart_gpu_raytrace_tau_batch: NULL handle");
  if (n < 0 || (n > 0 && (!x || !y || !z || !kx || !ky || !kz || !xfreq || !icell || !jcell || !kcell || !tau_in || !inside)))
    return fail("lart_gpu_raytrace_tau_batch: bad argument");
  if (n == 0) return 0;
  CUDA_OK(cudaSetDevice(h->device));
  Scratch s;
  double *dx, *dy, *dz, *dkx, *dky, *dkz, *dxf, *dti, *dxr;
  int *dic, *djc, *dkc, *dins, *dns;
  int rc = 0;
  rc = rc ? rc : s.in(&dx, (const double *)x, n); rc = rc ? rc : s.in(&dy, (const double *)y, n); rc = rc ? rc : s.in(&dz, (const double *)z, n);
  rc = rc ? rc : s.in(&dkx, kx, n); rc = rc ? rc : s.in(&dky, ky, n); rc = rc ? rc : s.in(&dkz, kz, n);
  rc = rc ? rc : s.in(&dxf, (const double *)xfreq, n); rc = rc ? rc : s.in(&dti, tau_in, n);
  rc = rc ? rc : s.in(&dic, (const int *)icell, n); rc = rc ? rc : s.in(&djc, (const int *)jcell, n); rc = rc ? rc : s.in(&dkc, (const int *)kcell, n);
  rc = rc ? rc : s.outbuf(&dins, n); rc = rc ? rc : s.outbuf(&dxr, n); rc = rc ? rc : s.outbuf(&dns, n);
  if (rc) return rc;
  k_tau_batch<<<grid_for(n, h->nsm), kBlock, 0, h->stream>>>(h->P, n, dx, dy, dz, dkx, dky, dkz, dxf, dic, djc, dkc, dti, dins, dxr, dns);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(h->stream));
  D2H(x, dx, n); D2H(y, dy, n); D2H(z, dz, n); D2H(xfreq, dxf, n);
  D2H(icell, dic, n); D2H(jcell, djc, n); D2H(kcell, dkc, n); D2H(inside, dins, n);
  if (xfreq_ref) D2H(xfreq_ref, dxr, n);
  if (nsteps) D2H(nsteps, dns, n);
  return 0;
}

int lart_gpu_xcrit_batch(lart_gpu_handle h, int64_t n, const double *x, const double *y, const double *z,
                         const int32_t *icell, const int32_t *jcell, const int32_t *kcell, double *xcrit) {
  if (!h) return fail("lart_gpu_xcrit_batch: NULL handle");
  if (n < 0 || (n > 0 && (!x || !y || !z || !icell || !jcell || !kcell || !xcrit))) return fail("lart_gpu_xcrit_batch: bad argument");
  if (n == 0) return 0;
  CUDA_OK(cudaSetDevice(h->device));
  Scratch s;
  double *dx, *dy, *dz, *dout;
  int *dic, *djc, *dkc;
  int rc = 0;
  rc = rc ? rc : s.in(&dx, x, n); rc = rc ? rc : s.in(&dy, y, n); rc = rc ? rc : s.in(&dz, z, n);
  rc = rc ? rc : s.in(&dic, icell, n); rc = rc ? rc : s.in(&djc, jcell, n); rc = rc ? rc : s.in(&dkc, kcell, n);
  rc = rc ? rc : s.outbuf(&dout, n);
  if (rc) return rc;
  k_xcrit_batch<<<grid_for(n, h->nsm), kBlock, 0, h->stream>>>(h->P, n, dx, dy, dz, dic, djc, dkc, dout);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(h->stream));
  D2H(xcrit, dout, n);
  return 0;
}

int lart_gpu_peel_bound_batch(lart_gpu_handle h, int64_t n, const double *x, const double *y, const double *z, const double *xfreq,
                              const int32_t *icell, const int32_t *jcell, const int32_t *kcell, int32_t *capped) {
  if (!h) return fail("lart_gpu_peel_bound_batch: NULL handle");
  if (n < 0 || (n > 0 && (!x || !y || !z || !xfreq || !icell || !jcell || !kcell || !capped))) return fail("lart_gpu_peel_bound_batch: bad argument");
  if (n == 0) return 0;
  for (int64_t i = 0; i < n; ++i)
    if (icell[i] < 1 || icell[i] > h->P.nx || jcell[i] < 1 || jcell[i] > h->P.ny || kcell[i] < 1 || kcell[i] > h->P.nz)
      return fail("lart_gpu_peel_bound_batch: cell index outside the grid");
  CUDA_OK(cudaSetDevice(h->device));
  Scratch s;
  double *dx, *dy, *dz, *dxf;
  int *dic, *djc, *dkc, *dout;
  int rc = 0;
  rc = rc ? rc : s.in(&dx, x, n); rc = rc ? rc : s.in(&dy, y, n); rc = rc ? rc : s.in(&dz, z, n); rc = rc ? rc : s.in(&dxf, xfreq, n);
  rc = rc ? rc : s.in(&dic, icell, n); rc = rc ? rc : s.in(&djc, jcell, n); rc = rc ? rc : s.in(&dkc, kcell, n);
  rc = rc ? rc : s.outbuf(&dout, n);
  if (rc) return rc;
  k_peel_bound_batch<<<grid_for(n, h->nsm), kBlock, 0, h->stream>>>(h->P, n, dx, dy, dz, dxf, dic, djc, dkc, dout);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(h->stream));
  D2H(capped, dout, n);
  return 0;
}

int lart_gpu_clump_edge_batch(lart_gpu_handle h, int64_t n, const double *x, const double *y, const double *z, const double *kx,
                              const double *ky, const double *kz, const double *xfreq, const int32_t *icl, double tau_max,
                              double *tau, int32_t *nclumps) {
  if (!h) return fail("lart_gpu_clump_edge_batch: NULL handle");
  if (!h->P.clump) return fail("lart_gpu_clump_edge_batch: the handle has no clump medium");
  if (n < 0 || (n > 0 && (!x || !y || !z || !kx || !ky || !kz || !xfreq || !icl || !tau))) return fail("lart_gpu_clump_edge_batch: bad argument");
  if (n == 0) return 0;
  CUDA_OK(cudaSetDevice(h->device));
  Scratch s;
  double *dx, *dy, *dz, *dkx, *dky, *dkz, *dxf, *dtau;
  int *dicl, *dncl = nullptr;
  unsigned long long *dcells;
  int rc = 0;
  rc = rc ? rc : s.in(&dx, x, n); rc = rc ? rc : s.in(&dy, y, n); rc = rc ? rc : s.in(&dz, z, n);
  rc = rc ? rc : s.in(&dkx, kx, n); rc = rc ? rc : s.in(&dky, ky, n); rc = rc ? rc : s.in(&dkz, kz, n);
  rc = rc ? rc : s.in(&dxf, xfreq, n); rc = rc ? rc : s.in(&dicl, icl, n);
  rc = rc ? rc : s.outbuf(&dtau, n); rc = rc ? rc : s.outbuf(&dcells, 1);
  if (nclumps) rc = rc ? rc : s.outbuf(&dncl, n);
  if (rc) return rc;
  CUDA_OK(cudaMemsetAsync(dcells, 0, sizeof(unsigned long long), h->stream));
  CUDA_OK(cudaEventRecord(h->ev0, h->stream));
  k_clump_edge_batch<<<grid_for(n, h->nsm), kBlock, 0, h->stream>>>(h->P, n, dx, dy, dz, dkx, dky, dkz, dxf, dicl, tau_max, dtau, dncl, dcells);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaEventRecord(h->ev1, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  float ms = 0.f;
  CUDA_OK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  unsigned long long cells = 0;
  D2H(&cells, dcells, 1);
  h->sight_ms = ms; h->sight_steps = (double)cells;  // read back through lart_gpu_sightline_stats
  D2H(tau, dtau, n);
  if (nclumps) D2H(nclumps, dncl, n);
  return 0;
}

int lart_gpu_clump_tau_batch(lart_gpu_handle h, int64_t n, double *x, double *y, double *z, const double *kx, const double *ky,
                             const double *kz, double *xfreq, int32_t *icl, const double *tau_in, int32_t *inside) {
  if (!h) return fail("lart_gpu_clump_tau_batch: NULL handle");
  if (!h->P.clump) return fail("lart_gpu_clump_tau_batch: the handle has no clump medium");
  if (h->P.cl.overlap) return fail("lart_gpu_clump_tau_batch: raytrace_to_tau_clump_overlap draws the owner clump from the photon's stream; use a run");
  if (n < 0 || (n > 0 && (!x || !y || !z || !kx || !ky || !kz || !xfreq || !icl || !tau_in || !inside))) return fail("lart_gpu_clump_tau_batch: bad argument");
  if (n == 0) return 0;
  CUDA_OK(cudaSetDevice(h->device));
  Scratch s;
  double *dx, *dy, *dz, *dkx, *dky, *dkz, *dxf, *dtau;
  int *dicl, *din;
  int rc = 0;
  rc = rc ? rc : s.in(&dx, (const double *)x, n); rc = rc ? rc : s.in(&dy, (const double *)y, n); rc = rc ? rc : s.in(&dz, (const double *)z, n);
  rc = rc ? rc : s.in(&dkx, kx, n); rc = rc ? rc : s.in(&dky, ky, n); rc = rc ? rc : s.in(&dkz, kz, n);
  rc = rc ? rc : s.in(&dxf, (const double *)xfreq, n); rc = rc ? rc : s.in(&dicl, (const int *)icl, n);
  rc = rc ? rc : s.in(&dtau, tau_in, n); rc = rc ? rc : s.outbuf(&din, n);
  if (rc) return rc;
  k_clump_tau_batch<<<grid_for(n, h->nsm), kBlock, 0, h->stream>>>(h->P, n, dx, dy, dz, dkx, dky, dkz, dxf, dicl, dtau, din);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(h->stream));
  D2H(x, dx, n); D2H(y, dy, n); D2H(z, dz, n); D2H(xfreq, dxf, n); D2H(icl, dicl, n); D2H(inside, din, n);
  return 0;
}

int lart_gpu_clump_locate_batch(lart_gpu_handle h, int64_t n, const double *x, const double *y, const double *z, int32_t *icl) {
  if (!h) return fail("lart_gpu_clump_locate_batch: NULL handle");
  if (!h->P.clump) return fail("lart_gpu_clump_locate_batch: the handle has no clump medium");
  if (n < 0 || (n > 0 && (!x || !y || !z || !icl))) return fail("lart_gpu_clump_locate_batch: bad argument");
  if (n == 0) return 0;
  CUDA_OK(cudaSetDevice(h->device));
  Scratch s;
  double *dx, *dy, *dz;
  int *dicl;
  int rc = 0;
  rc = rc ? rc : s.in(&dx, x, n); rc = rc ? rc : s.in(&dy, y, n); rc = rc ? rc : s.in(&dz, z, n); rc = rc ? rc : s.outbuf(&dicl, n);
  if (rc) return rc;
  k_clump_locate_batch<<<grid_for(n, h->nsm), kBlock, 0, h->stream>>>(h->P, n, dx, dy, dz, dicl);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(h->stream));
  D2H(icl, dicl, n);
  return 0;
}

int lart_gpu_amr_locate_batch(lart_gpu_handle h, int64_t n, const double *x, const double *y, const double *z, int32_t *il) {
  if (!h) return fail("lart_gpu_amr_locate_batch: NULL handle");
  if (!h->P.amr.on) return fail("lart_gpu_amr_locate_batch: the handle has no octree");
  if (n < 0 || (n > 0 && (!x || !y || !z || !il))) return fail("lart_gpu_amr_locate_batch: bad argument");
  if (n == 0) return 0;
  CUDA_OK(cudaSetDevice(h->device));
  Scratch s;
  double *dx, *dy, *dz;
  int *dil;
  int rc = 0;
  rc = rc ? rc : s.in(&dx, x, n); rc = rc ? rc : s.in(&dy, y, n); rc = rc ? rc : s.in(&dz, z, n); rc = rc ? rc : s.outbuf(&dil, n);
  if (rc) return rc;
  k_amr_locate_batch<<<grid_for(n, h->nsm), kBlock, 0, h->stream>>>(h->P, n, dx, dy, dz, dil);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(h->stream));
  D2H(il, dil, n);
  return 0;
}

int lart_gpu_sample_batch(int32_t kind, uint64_t seed, int64_t n, const int64_t *ids, const double *p0, const double *p1,
                          int32_t ndraw, double *out) {
  if (kind < 0 || kind > 6) return fail("lart_gpu_sample_batch: unknown kind");
  if (n < 0 || ndraw < 1 || (n > 0 && !out)) return fail("lart_gpu_sample_batch: bad argument");
  if ((kind >= 2) && n > 0 && !p0) return fail("lart_gpu_sample_batch: p0 required");
  if ((kind == 2 || kind == 6) && n > 0 && !p1) return fail("lart_gpu_sample_batch: p1 required");
  if (n == 0) return 0;
  Scratch s;
  long long *dids;
  double *dp0, *dp1, *dout;
  int rc = 0;
  rc = rc ? rc : s.in(&dids, (const long long *)ids, n);
  rc = rc ? rc : s.in(&dp0, p0, n); rc = rc ? rc : s.in(&dp1, p1, n);
  rc = rc ? rc : s.outbuf(&dout, (size_t)n * ndraw);
  if (rc) return rc;
  k_sample_batch<<<grid_for(n, 148), kBlock>>>(kind, seed, n, dids, dp0, dp1, ndraw, dout);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaDeviceSynchronize());
  D2H(out, dout, (size_t)n * ndraw);
  return 0;
}


// ---------------------------- sight-line maps -----------------------------
}  // extern "C"

namespace {
// Entry point and start cell of the sight line of pixel (ix,iy) — sightline_tau_rect.f90:45-150.
bool sightline_start(const lart_gpu_ctx *h, const std::vector<double> &xf, const std::vector<double> &yf,
                     const std::vector<double> &zf, const DevObserver &ob, int ix, int iy, SightStart &o) {
  const DevParams &P = h->P;
  const double rad2deg = 180.0 / 3.141592653589793238462643383279502884197, hugest = 1.7976931348623157e308;
  double kx = std::tan((ix - (ob.nxim + 1.0) / 2.0) * ob.dxim / rad2deg);
  double ky = std::tan((iy - (ob.nyim + 1.0) / 2.0) * ob.dyim / rad2deg);
  double kz = -1.0;
  const double kr = std::sqrt(kx * kx + ky * ky + kz * kz);
  kx /= kr; ky /= kr; kz /= kr;
  const double *R = ob.R;
  double px = R[0] * kx + R[1] * ky + R[2] * kz, py = R[3] * kx + R[4] * ky + R[5] * kz, pz = R[6] * kx + R[7] * ky + R[8] * kz;
  const double delt[6] = {px == 0.0 ? hugest : (P.xmax - ob.x) / px, px == 0.0 ? hugest : (P.xmin - ob.x) / px,
                          py == 0.0 ? hugest : (P.ymax - ob.y) / py, py == 0.0 ? hugest : (P.ymin - ob.y) / py,
                          pz == 0.0 ? hugest : (P.zmax - ob.z) / pz, pz == 0.0 ? hugest : (P.zmin - ob.z) / pz};
  auto cellof = [&](double x, double y, double z, int &i, int &j, int &k) {
    i = (int)std::floor((x - P.xmin) / P.dx) + 1;
    j = (int)std::floor((y - P.ymin) / P.dy) + 1;
    k = (int)std::floor((z - P.zmin) / P.dz) + 1;
  };
  double dist = -999.9;
  int j0 = 0;
  for (int jj = 1; jj <= 6; ++jj) {
    const double dl = delt[jj - 1];
    if (!(dl > 0.0 && dl < hugest)) continue;
    int i, j, k;
    cellof(ob.x + px * dl, ob.y + py * dl, ob.z + pz * dl, i, j, k);
    if (jj == 1) i = P.nx + 1;
    if (jj == 2) i = 1;
    if (jj == 3) j = P.ny + 1;
    if (jj == 4) j = 1;
    if (jj == 5) k = P.nz + 1;
    if (jj == 6) k = 1;
    if (i >= 1 && i <= P.nx + 1 && j >= 1 && j <= P.ny + 1 && k >= 1 && k <= P.nz + 1 && dl > dist) { dist = dl; j0 = jj; }
  }
  o.valid = 0;
  if (!(dist > 0.0 && dist < hugest)) return false;
  o.x = ob.x + px * dist; o.y = ob.y + py * dist; o.z = ob.z + pz * dist;
  cellof(o.x, o.y, o.z, o.ic, o.jc, o.kc);
  o.kx = -px; o.ky = -py; o.kz = -pz;
  if (j0 == 1) { o.ic = P.nx + 1; o.x = xf[P.nx]; }
  else if (j0 == 2) { o.ic = 1; o.x = xf[0]; }
  else if (j0 == 3) { o.jc = P.ny + 1; o.y = yf[P.ny]; }
  else if (j0 == 4) { o.jc = 1; o.y = yf[0]; }
  else if (j0 == 5) { o.kc = P.nz + 1; o.z = zf[P.nz]; }
  else if (j0 == 6) { o.kc = 1; o.z = zf[0]; }
  if (o.ic == P.nx + 1 && o.kx < 0.0) o.ic = P.nx;
  if (o.jc == P.ny + 1 && o.ky < 0.0) o.jc = P.ny;
  if (o.kc == P.nz + 1 && o.kz < 0.0) o.kc = P.nz;
  o.valid = (o.ic >= 1 && o.ic <= P.nx && o.jc >= 1 && o.jc <= P.ny && o.kc >= 1 && o.kc <= P.nz) ? 1 : 0;
  return o.valid != 0;
}
}  // namespace

extern "C" {

int lart_gpu_sightline_tau(lart_gpu_handle h, double cross0, lart_sightline_out *out) {
  if (!h || !out) return fail("lart_gpu_sightline_tau: NULL argument");
  if (h->obs_host.empty()) return fail("lart_gpu_sightline_tau: the handle has no observers (par%save_peeloff, par%nobs)");
  if (h->P.amr.on) return fail("lart_gpu_sightline_tau: sight-line maps of an octree stay with the Fortran host");
  if (h->P.zonly) return fail("lart_gpu_sightline_tau: not defined for the xy-periodic slab");
  if (h->P.bcxy) return fail("lart_gpu_sightline_tau: not defined for folded or periodic grids");
  if (!(cross0 > 0.0)) return fail("lart_gpu_sightline_tau: cross0 must be > 0");
  CUDA_OK(cudaSetDevice(h->device));
  const DevParams &P = h->P;
  std::vector<double> xf(P.nx + 1), yf(P.ny + 1), zf(P.nz + 1);
  CUDA_OK(cudaMemcpy(xf.data(), P.xface, sizeof(double) * xf.size(), cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(yf.data(), P.yface, sizeof(double) * yf.size(), cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(zf.data(), P.zface, sizeof(double) * zf.size(), cudaMemcpyDeviceToHost));
  h->sight_steps = 0.0;
  h->sight_ms = 0.0;
  for (size_t k = 0; k < h->obs_host.size(); ++k) {
    const DevObserver &ob = h->obs_host[k];
    if (!out[k].tau_gas || !out[k].N_gas) return fail("lart_gpu_sightline_tau: tau_gas / N_gas output is NULL");
    const long long npix = (long long)ob.nxim * ob.nyim;
    std::vector<SightStart> st((size_t)npix);
    for (int iy = 1; iy <= ob.nyim; ++iy)
      for (int ix = 1; ix <= ob.nxim; ++ix) sightline_start(h, xf, yf, zf, ob, ix, iy, st[(size_t)(ix - 1) + (size_t)ob.nxim * (iy - 1)]);
    Scratch s;
    SightStart *dst;
    double *dtau, *dN, *dtd = nullptr;
    unsigned long long *dsteps;
    if (int rc = s.in(&dst, st.data(), (size_t)npix)) return rc;
    if (int rc = s.outbuf(&dtau, (size_t)npix * P.nxfreq)) return rc;
    if (int rc = s.outbuf(&dN, (size_t)npix)) return rc;
    if (P.dust && out[k].tau_dust) if (int rc = s.outbuf(&dtd, (size_t)npix)) return rc;
    if (int rc = s.outbuf(&dsteps, 1)) return rc;
    CUDA_OK(cudaMemsetAsync(dtau, 0, sizeof(double) * (size_t)npix * P.nxfreq, h->stream));
    CUDA_OK(cudaMemsetAsync(dN, 0, sizeof(double) * (size_t)npix, h->stream));
    if (dtd) CUDA_OK(cudaMemsetAsync(dtd, 0, sizeof(double) * (size_t)npix, h->stream));
    CUDA_OK(cudaMemsetAsync(dsteps, 0, sizeof(unsigned long long), h->stream));
    CUDA_OK(cudaEventRecord(h->ev0, h->stream));
    k_sightline<<<h->nsm * 8, kBlock, 0, h->stream>>>(P, npix, dst, cross0, dtau, dN, dtd, dsteps);
    CUDA_OK(cudaEventRecord(h->ev1, h->stream));
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    CUDA_OK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    unsigned long long ns = 0;
    D2H(&ns, dsteps, 1);
    h->sight_ms += ms;
    h->sight_steps += (double)ns;
    h->launches += 1;
    D2H(out[k].tau_gas, dtau, (size_t)npix * P.nxfreq);
    D2H(out[k].N_gas, dN, (size_t)npix);
    if (dtd) D2H(out[k].tau_dust, dtd, (size_t)npix);
  }
  return 0;
}

/* cell steps and CUDA-event time (ms) of the last lart_gpu_sightline_tau call */
int lart_gpu_sightline_stats(lart_gpu_handle h, double *cellsteps, double *ms) {
  if (!h) return fail("lart_gpu_sightline_stats: NULL handle");
  if (cellsteps) *cellsteps = h->sight_steps;
  if (ms) *ms = h->sight_ms;
  return 0;
}
}  // extern "C"
