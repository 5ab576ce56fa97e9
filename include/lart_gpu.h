/* lart_gpu.h — C ABI of the B200-native photon-transport engine.
 *
 * This is the drop-in boundary for ONE path of LaRT v2.00: the Monte-Carlo photon
 * loop on the Cartesian grid (every binding of setup.f90:947-987), the clump medium
 * and the octree.  Every entry point below is what a Fortran
 * ISO_C_BINDING interface block (shim/lart_gpu_shim.f90) binds;
 * each cites the reference interface it replaces (paths relative to the
 * reference tree, file:line).
 *
 *   - plain C, `extern "C"`, POD structs with explicit int32/int64/double
 *     members and raw pointers; no C++/torch types cross this boundary;
 *   - every function returns int: 0 = OK, non-zero = error; the message is
 *     retrievable through lart_gpu_last_error() (reference convention: print,
 *     then MPI_ABORT — src/setup.f90:132-135; the shim does the abort);
 *   - all arrays are Fortran column-major, x fastest, exactly as grid_type
 *     holds them (src/define.f90:131-148); indices in the API are 1-based as
 *     in the reference, conversion to 0-based happens inside the library;
 *   - the library never frees or reallocates host memory; device memory is
 *     owned by the opaque handle;
 *   - tallies come back as raw, un-normalised weighted sums that are ADDED
 *     into caller-owned buffers (normalisation stays with the host,
 *     src/output_sum_rect.f90:151-487).
 *
 * Called from one host thread per process (the reference is not re-entrant
 * either: global par/observer/RNG state, src/define.f90:729-737).
 */
#ifndef LART_GPU_H
#define LART_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LART_MAX_OBSERVERS 181 /* src/define.f90:77 */

/* par%spectral_type (src/generate_photon.f90:243-300) */
enum {
  LART_SPEC_MONO = 0,      /* any other string: x = xfreq0                 */
  LART_SPEC_VOIGT = 1,     /* 'voigt'   : x0 + rand_voigt(a_cell)           */
  LART_SPEC_VOIGT0 = 2,    /* 'voigt0'  : x0 + rand_voigt(a0)*Dfreq0/Dfreq  */
  LART_SPEC_CONTINUUM = 3, /* 'continuum': uniform in [xfreq_min,xfreq_max] */
  LART_SPEC_GAUSSIAN = 4   /* 'gaussian': x0 + gauss*sigma_x (ref. units)   */
};

/* par%source_geometry (src/generate_photon.f90:33-132) */
enum {
  LART_SRC_POINT = 0,         /* default: (xs,ys,zs)_point                  */
  LART_SRC_UNIFORM = 1,       /* 'uniform': uniform in the box              */
  LART_SRC_UNIFORM_SPHERE = 2, /* 'uniform_sphere'/'sphere': r<source_rmax   */
  LART_SRC_PLANE_ILLUMINATION = 3 /* 'plane_illumination': parallel beam onto an atmosphere, random_plane_illumination
                                    (src/generate_photon.f90:729-812); needs par.atmosphere != 0 */
};

/* par%geometry of the exoplanet-atmosphere models (src/setup.f90:86-113, 959-987) */
enum {
  LART_ATM_NONE = 0,
  LART_ATM_PLANE = 1,    /* 'plane_atmosphere': 1x1xnz, raytrace_to_tau_car_zonly_atmosphere (raytrace_car.f90:2956-3117):
                            a photon that leaves through the bottom cell is absorbed into Jabs2 */
  LART_ATM_SPHERICAL = 2 /* 'spherical_atmosphere': raytrace_to_tau_car_atmosphere / _xysym_atmosphere (:3119-3663),
                            raytrace_to_edge_car_atmosphere (:3665-3770) and the two sight-line variants (:3772-3976):
                            cells with grid%mask == -1 destroy the photon (Jabs2), rays through them have tau = +inf */
};

/* grid_type scalars + arrays — src/define.f90:117-148; built by
 * src/grid_mod_car.f90:73-209,273-284,487-538,786-803. */
typedef struct lart_grid {
  int32_t nx, ny, nz;
  int32_t nxfreq;
  double xmin, ymin, zmin;
  double xmax, ymax, zmax;
  double dx, dy, dz;
  double Dfreq_ref;            /* grid%Dfreq_ref   (grid_mod_car.f90:245)   */
  double xfreq_min, xfreq_max; /* grid%xfreq_min/max (:1496-1498)           */
  double dxfreq;               /* grid%dxfreq                               */
  double xcrit, xcrit2;        /* grid%xcrit(2): global core-skip (:1186-1219) */
  double rmax;                 /* par%rmax (<=0: none); used by allph records  */
  int32_t i0, j0, k0;          /* grid%i0,j0,k0 (grid_mod_car.f90:92-118): the cell a photon re-enters when it is
                                  reflected at a lower face of an xyz-symmetric (one octant) grid; 0 otherwise */
  int32_t pad_;
  const double *xface;         /* (nx+1)  xface(i) = (i-1)*dx + xmin        */
  const double *yface;         /* (ny+1)                                    */
  const double *zface;         /* (nz+1)                                    */
  const double *rhokap;        /* (nx,ny,nz) dtau = rhokap*H(x,a)*ds        */
  const double *voigt_a;       /* (nx,ny,nz)                                */
  const double *Dfreq;         /* (nx,ny,nz)                                */
  const double *vfx, *vfy, *vfz; /* (nx,ny,nz) bulk velocity / v_th(cell)   */
  const double *rhokapD;       /* (nx,ny,nz) dust extinction, or NULL if DGR==0 */
  /* ---- next rows; all NULL / 0 when unused ---- */
  const int8_t *mask;          /* (nx,ny,nz) grid%mask (grid_mod_car.f90:247-250, 320-330): -1 = the planet's molecular
                                  layer; read only with par.atmosphere == LART_ATM_SPHERICAL */
  int32_t geometry_JPa;        /* par%geometry_JPa of the CALCJ/CALCP/CALCPnew accumulators (grid_mod_car.f90:1242-1440):
                                  3 = per cell, 2 = cylindrical (nr,nz), 1 = spherical (nr), -1 = plane parallel (nz) */
  int32_t nr;                  /* grid%nr: radial bins of geometries 1 and 2                                  */
  const int32_t *ind_sph;      /* (nx,ny,nz) grid%ind_sph, 1-based radial bin of a cell (geometry 1)          */
  const int32_t *ind_cyl;      /* (nx,ny)    grid%ind_cyl (geometry 2)                                        */
} lart_grid;

/* the members of params_type the path reads — src/define.f90:209-544 */
typedef struct lart_params {
  int64_t nphotons;  /* par%nphotons: photon ids run 1..nphotons            */
  uint64_t seed;     /* Philox key; replaces par%iseed (random_mt.f90:889-956) */
  double xfreq0;
  double xs_point, ys_point, zs_point;
  double source_rmax;
  double DGR, albedo, hgg;
  double voigt_a0, Dfreq0;   /* 'voigt0' emission                           */
  double gaussian_sigma_x;   /* 'gaussian': sigma in reference Doppler units */
  double mu_min, dmu;        /* Jmu binning (setup.f90:375-382)             */
  int32_t nmu;
  int32_t spectral_type;     /* LART_SPEC_*                                 */
  int32_t source_geometry;   /* LART_SRC_*                                  */
  int32_t comoving_source;
  int32_t recoil;
  int32_t core_skip, core_skip_global;
  int32_t use_stokes;
  int32_t use_reduced_wgt;
  int32_t save_Jin, save_Jabs, save_Jmu;
  int32_t save_peeloff, save_peeloff_2D, save_peeloff_3D, save_direc0;
  int32_t save_all_photons;
  int32_t xyz_symmetry;      /* one octant with mirror planes at the lower faces: binds the _xyzsym ray tracers
                                (setup.f90:952-954; raytrace_car.f90:584-760, 1650-1949); no peel-off     */
  int32_t xy_symmetry;       /* mirror planes at the lower x and y faces, z open: the _xysym ray tracers
                                (setup.f90:955-957; raytrace_car.f90:783-969, 1951-2250)                  */
  int32_t use_clump_medium;  /* the clump ray tracers and do_resonance1_clump (setup.f90:806-860); needs cfg.clumps */
  int32_t xy_periodic;       /* nx==ny==1: the _zonly ray tracers; otherwise the _xyper ray tracers, photons
                                wrap around in x and y (setup.f90:958-976; raytrace_car.f90:971-1136,
                                2252-2517); with par.Omega /= 0 the shearing-box variant (:2677-2954) */
  int32_t nobs;
  int32_t use_amr_grid;      /* the octree ray tracers (raytrace_amr.f90:77-351) and leaf-indexed physics; needs cfg.amr */
  int32_t atmosphere;        /* LART_ATM_*                                                                    */
  int32_t calc_J;            /* the reference's -DCALCJ build: mean intensity J(x, cell) from path lengths,
                                add_to_J (raytrace_car.f90:3979-4011), called once per cell step of raytrace_to_tau */
  int32_t calc_P;            /* -DCALCP: scattering rate per atom, add_to_Pa (scattering_car.f90:829-860), once per
                                resonance scattering                                                          */
  int32_t calc_Pnew;         /* -DCALCPnew: the same rate from path lengths, add_to_Pnew (raytrace_car.f90:4015-4045) */
  double Omega;              /* par%Omega after grid_mod_car.f90:348-350 (q*Omega*xrange in thermal-velocity units of
                                the reference frame): /= 0 with xy_periodic and nx,ny > 1 binds
                                raytrace_to_tau_car_xyper_shear (raytrace_car.f90:2677-2954); raytrace_to_edge then stays
                                the plain open-box routine, as upstream (setup.f90:967-969)                   */
} lart_params;

/* line_type members — src/define.f90:639-656; values from
 * src/line_mod.f90:1241-1270 (host owns the numbers, incl. the g_recoil0 quirk) */
typedef struct lart_line {
  int32_t line_type; /* only 1 (singlet, Ly-alpha without fine structure)  */
  int32_t pad_;
  double E1, E2, E3;
  double g_recoil0;
  double DnuHK_Hz;
  double cross0;     /* line%cross0: rhokap*Dfreq/cross0 = number density x distance2cm (add_to_Pa, add_to_Pnew) */
} lart_line;

/* observer_type members — src/define.f90:547-560; rmatrix is the Fortran
 * (3,3) array in memory order, i.e. rmatrix[(r-1) + 3*(c-1)] = rmatrix(r,c). */
typedef struct lart_observer {
  double x, y, z;
  double rmatrix[9];
  double dxim, dyim;
  int32_t nxim, nyim;
} lart_observer;

/* scattering_matrix_type — src/define.f90:616-625 (dust + Stokes only) */
typedef struct lart_scatt_mat {
  int32_t nPDF;
  int32_t pad_;
  const double *coss, *S11, *S12, *S33, *S34; /* (nPDF), S* already normalised */
  const double *phase_PDF;                   /* (nPDF-1) alias probabilities   */
  const int32_t *alias;                      /* (nPDF-1) 1-based, 0 = none     */
} lart_scatt_mat;

/* ---- next row: the clump medium — src/clump_mod.f90:30-118 -----------------------
 * N spherical clumps inside a sphere of radius sphere_R, vacuum between them (par%use_clump_medium).  The host owns
 * the population (init_clumps / generate_clumps / read_clumps_info) and the CSR acceleration grid
 * (build_clump_csr, :1267-1349); the library only reads them.  n = 0: no clump medium.  Overlapping populations
 * (has_overlap: the event-walk ray tracers raytrace_to_tau_clump_overlap / raytrace_to_edge_clump_overlap(_capped),
 * raytrace_clump.f90:621-920, and the scatter_resonance_clump_* wrappers, scattering_car.f90:897-945) run on the
 * one-thread-per-photon driver with a per-thread event list of MAX_EVT = 2048 entries, as upstream. */
typedef struct lart_clumps {
  int64_t n;                        /* N_clumps                                                     */
  double sphere_R;                  /* outer radius                                                 */
  double Dfreq_ref;                 /* cl_Dfreq_ref (= grid.Dfreq_ref in clump mode)                */
  const double *x, *y, *z;          /* cl_x/y/z (n): centres                                        */
  const double *vx, *vy, *vz;       /* cl_vx/y/z (n): bulk velocity / cl_vtherm(icl)                */
  const double *radius;             /* cl_radius (n); cl_radius2 = radius*radius                    */
  const double *rhokap;             /* cl_rhokap (n): line-centre opacity per code length           */
  const double *rhokapD;            /* cl_rhokapD (n) or NULL when DGR = 0                          */
  const double *voigt_a, *Dfreq;    /* cl_voigt_a, cl_Dfreq (n)                                     */
  int32_t cgx, cgy, cgz;            /* CSR grid cells per axis                                      */
  int32_t has_overlap;              /* clump_mod's has_overlap (setup_clump_overlap, setup.f90:1051-1081) */
  double cg_xmin, cg_ymin, cg_zmin; /* lower corner of the CSR grid                                 */
  double cg_dx, cg_dy, cg_dz;       /* its cell sizes; cg_inv_d* = 1/cg_d* is recomputed            */
  const int32_t *cg_start;          /* (cgx*cgy*cgz + 1) 1-based offsets into cg_list, as upstream  */
  const int32_t *cg_list;           /* 1-based clump indices                                        */
} lart_clumps;

/* ---- next row: the octree AMR grid — amr_grid_type, src/octree_mod.f90:19-138 ------------------
 * The host owns the flat, 1-indexed tree (amr_build_tree, :470-575), its face-neighbour table (amr_build_neighbors,
 * :619-683) and the leaf physics (grid_create_amr, src/grid_mod_amr.f90:34-526); the library only reads them.
 * With par.use_amr_grid the photon's cell is a LEAF index (photon%icell_amr), lart_grid supplies the box
 * (xmin..zmax), Dfreq_ref, the frequency grid and xcrit; its nx, ny, nz and cell arrays are not read.
 * Not on the GPU path: periodic / mirror boundaries of the octree, band-2 (Ly-beta) photons, H2, CALCJ/CALCP on leaves. */
typedef struct lart_amr {
  int32_t ncells, nleaf;            /* all cells (internal + leaf), leaves                                   */
  const int32_t *children;          /* (8,ncells) child cell or 0; octant = 1 + ix + 2*iy + 4*iz             */
  const int32_t *ileaf;             /* (ncells)   leaf index of a leaf cell, 0 for an internal cell          */
  const int32_t *icell_of_leaf;     /* (nleaf)    cell index of a leaf                                       */
  const int32_t *neighbor;          /* (6,ncells) same-level face neighbour (+x,-x,+y,-y,+z,-z) or 0         */
  const double *cx, *cy, *cz, *ch;  /* (ncells)   cell centres and half-widths                               */
  const double *rhokap, *voigt_a, *Dfreq, *vfx, *vfy, *vfz; /* (nleaf) as lart_grid's cell arrays            */
  const double *rhokapD;            /* (nleaf) or NULL when DGR = 0                                          */
} lart_amr;

typedef struct lart_config {
  lart_grid grid;
  lart_params par;
  lart_line line;
  lart_scatt_mat scatt_mat;        /* nPDF = 0 when unused                  */
  lart_clumps clumps;              /* n = 0 when unused                     */
  lart_amr amr;                    /* nleaf = 0 when unused                 */
  const lart_observer *observers;  /* par.nobs entries                      */
  int32_t device;                  /* CUDA device ordinal                   */
  int32_t pool_slots;              /* photons in flight; 0 = auto           */
  int32_t quantum;                 /* scattering events per slot per lart_gpu_step; 0 = auto */
  int32_t flags;                   /* LART_FLAG_*                           */
  int32_t streams;                 /* wave pipelines: the pool is split into this many partitions,
                                      each advanced on its own CUDA stream; 0 = auto (6) */
  int32_t ray_budget;              /* cell steps a transport or peel ray may take per wave before it is
                                      parked and resumed (exactly) in the next wave; 0 = auto (32) */
  int32_t max_events;              /* bounded runs (benchmarks and parity runs of cases where a photon needs ~1e7
                                      scatterings to escape): a photon is abandoned after this many scatterings — no Jout
                                      tally, its allph record holds the state it had (current frequency in xfreq2);
                                      0 = every photon runs until it escapes or is absorbed, as in the reference */
  int32_t pad_;
} lart_config;

enum {
  LART_FLAG_SOA_GRID = 1,   /* walk the six SoA arrays instead of packed cell records */
  LART_FLAG_NO_WARP_AGG = 2, /* plain atomics for peel tallies (ablation)      */
  LART_FLAG_MONOLITHIC = 4,  /* one thread per photon slot, no stage compaction
                                (the "before" arm of the warp-efficiency evidence) */
  LART_FLAG_STAGE_TIMING = 8, /* CUDA-event timing of every stage kernel (bench/roofline) */
  LART_FLAG_SERIAL_REJECTION = 16, /* per-lane rejection loops in the scatter stage: every warp waits for its
                                      slowest lane (ablation; same results) */
  LART_FLAG_LOCAL_STEPS = 32,      /* the scatter stage takes the first cell step of the peel ray and
                                      of the next flight itself; only longer rays reach the queues
                                      (experimental; same results) */
  LART_FLAG_SPECULATIVE_REJECTION = 128, /* round-2 first version of the scatter stage's samplers: the photons of a
                                      warp that still wait for an accepted trial share its 32 lanes, 32/np
                                      speculative trials each (ablation; same results).  Default: rejection loops
                                      compacted with per-lane refill over a chunk of slots per warp (k_wf_draw2) */
  LART_FLAG_DEBUG_TINY_QUEUES = 64 /* tests only: one direct-peel entry per pool partition, so that the
                                      queue-overflow error path (sticky device error word -> error
                                      return of lart_gpu_step / _sync / _fetch) can be exercised */
};

/* stage kernels of one wave, in launch order (index into lart_gpu_stage_ms) */
enum { LART_STAGE_EMIT = 0, LART_STAGE_TRACE = 1, LART_STAGE_DRAW = 2 /* k_wf_draw; the whole scatter stage of the clump driver */,
       LART_STAGE_APPLY = 3, LART_STAGE_PEEL = 4, LART_STAGE_COUNT = 5 };

/* per-observer output cubes (src/define.f90:561-600); frequency fastest:
 * cube(ixf,ix,iy) at [(ixf-1) + nxfreq*((ix-1) + nxim*(iy-1))]. NULL = skip. */
typedef struct lart_observer_out {
  double *scatt, *direc, *direc0, *I, *Q, *U, *V;                      /* 3-D */
  double *scatt_2D, *direc_2D, *direc0_2D, *I_2D, *Q_2D, *U_2D, *V_2D; /* 2-D */
} lart_observer_out;

/* allph record — src/define.f90:602-613; written by photon id (1-based, slot id-1) */
typedef struct lart_allph_out {
  double *rp0, *rp, *xfreq1, *xfreq2, *nscatt_gas, *nscatt_dust, *I, *Q, *U, *V;
} lart_allph_out;

/* work counters: the units the roofline is computed from */
typedef struct lart_counters {
  double n_photons_done;
  double n_scatter;     /* resonance + dust scattering events (unweighted)  */
  double n_cellsteps;   /* DDA cell steps, all ray kinds                    */
  double n_peel;        /* peel-off rays traced                             */
  double n_rng;         /* uniforms drawn                                   */
  double n_reject_iter; /* iterations of the rejection loops                */
  double n_peel_bound;  /* of n_peel: rays the scatter stage proved to end inside their own cell at the tau cap
                           (raytrace_car.f90:432,497) and therefore counted without queueing or walking them */
  double n_cellsteps_bound; /* of n_cellsteps: the one cell step each such ray would have taken (not walked) */
} lart_counters;

typedef struct lart_tallies {
  double *Jout, *Jin, *Jabs; /* (nxfreq) — grid%Jout/Jin/Jabs             */
  double *Jmu;               /* (nxfreq,nmu)                               */
  lart_observer_out *obs;    /* par.nobs entries (or NULL)                 */
  lart_allph_out allph;      /* all NULL unless save_all_photons           */
  double nscatt_gas, nscatt_dust; /* par%nscatt_* : weighted sums, ADDED   */
  lart_counters counters;    /* ADDED                                      */
  double *Jabs2;             /* (nxfreq) grid%Jabs2: photons destroyed by the atmosphere's molecular zone (NULL = skip) */
  double *J;                 /* CALCJ:    (nxfreq, [nx,ny,nz | nr,nz | nr | nz]) by geometry_JPa 3 | 2 | 1 | -1 —
                                grid%J / J2 / J1 (grid_mod_car.f90:1422-1434), raw sums of path length x weight  */
  double *Pa;                /* CALCP:    ([nx,ny,nz | nr,nz | nr | nz]) grid%Pa / P2 / P1 (:1385-1395)          */
  double *Pnew;              /* CALCPnew: the same shape, grid%Pa_new / P2_new / P1_new (:1410-1420)             */
} lart_tallies;

typedef struct lart_gpu_ctx *lart_gpu_handle;

/* ---- whole-path drop-in: one more implementation of run_sim ------------- */

/* Upload grid/par/line/observers, allocate photon pool and zeroed tallies.
 * Replaces nothing by itself; precedes the call that replaces
 * `call run_simulation(grid)` (src/main.f90:42, interface src/define.f90:832-838). */
int lart_gpu_create(const lart_config *cfg, lart_gpu_handle *out);

/* Run photons id = first_id, first_id+stride, ... (count of them) to completion:
 * generate_photon + forced first scattering + {raytrace_to_tau, scattering,
 * peeling}* — replaces the body of run_equal_number
 * (src/run_simulation_mod.f90:150-202; partition rule :150). */
int lart_gpu_run(lart_gpu_handle h, int64_t first_id, int64_t count, int64_t stride);

/* ---- dynamic photon dealing on one node: the GPU analogue of the reference's master/worker loop -------------------------
 * (src/run_simulation_mod.f90:31-128: workers ask the master for batches of par%num_send_at_once photons).  The processes
 * of a node (one per GPU) share ONE 64-bit counter in POSIX shared memory, `name` ('/lart_...'); a process whose job queue
 * runs dry claims the next `batch` photon ids with an atomic fetch-add — no master rank, no messages — and keeps its
 * photons in flight while the new ids are emitted.  Photon streams are keyed by id, so tallies summed over the processes
 * equal the statically partitioned run (lart_gpu_run) up to summation order.
 *   one rank:   lart_gpu_deal_open(name, 1, &d)   creates and zeroes the counter; then a host barrier (MPI_BARRIER)
 *   the others: lart_gpu_deal_open(name, 0, &d)
 *   every rank: lart_gpu_run_dealt(h, d, par%nphotons, batch, &mine)   mine = photons this process ran
 *   every rank: lart_gpu_deal_close(d, unlink)     unlink != 0 on one rank removes the name */
typedef struct lart_gpu_deal *lart_gpu_deal_handle;
int lart_gpu_deal_open(const char *name, int32_t reset, lart_gpu_deal_handle *out);
int lart_gpu_deal_close(lart_gpu_deal_handle d, int32_t unlink_name);
/* one claim: the next `batch` photon ids, first_id .. first_id+count-1 (count = 0: none left).  Host-only (no GPU needed);
 * lart_gpu_run_dealt makes exactly this call whenever its job queue runs dry. */
int lart_gpu_deal_claim(lart_gpu_deal_handle d, int64_t nphotons, int64_t batch, int64_t *first_id, int64_t *count);
int lart_gpu_run_dealt(lart_gpu_handle h, lart_gpu_deal_handle d, int64_t nphotons, int64_t batch, int64_t *nclaimed);

/* Bounded-work variant of the same loop (used for benchmarking heavy-tailed
 * cases): begin() queues the photon ids, each step() advances every pool slot by
 * at most `quantum` scattering events (`quantum` waves of the emit/trace/scatter/
 * peel stage kernels; finished photons are refilled from the queue);
 * *in_flight returns photons still alive or queued. */
int lart_gpu_begin(lart_gpu_handle h, int64_t first_id, int64_t count, int64_t stride);
int lart_gpu_step(lart_gpu_handle h, int32_t quantum, int64_t *in_flight);
int lart_gpu_sync(lart_gpu_handle h);

/* Add the device tallies into host buffers — the data half of output_reduce
 * (src/output_sum_rect.f90:7-149; src/memory_mod_mpi.f90:366-458). */
int lart_gpu_fetch(lart_gpu_handle h, lart_tallies *out);
int lart_gpu_reset_tallies(lart_gpu_handle h);

/* ---- multi-GPU: the communicator half of output_reduce -------------------------------------------------------------
 * One process per GPU (one MPI rank per GPU on the Fortran side).  lart_gpu_reduce sums the tallies of all ranks onto
 * `root` with ONE ncclReduce(sum, f64) over the contiguous device tally buffer (and one over the allph buffer) across
 * NVLink — it replaces reduce_mem's per-array, plane-by-plane MPI_REDUCE of host arrays
 * (src/memory_mod_mpi.f90:366-458, called from src/output_sum_rect.f90:13-146); the host then calls lart_gpu_fetch on
 * root only and skips its own reduce.  After the call the other ranks' device tallies are zero (their contribution
 * lives on root), so the sum over ranks always is "everything tallied so far".
 *   rank 0:     lart_gpu_comm_unique_id(id)      -> 128 bytes, broadcast by the host (MPI_BCAST)
 *   every rank: lart_gpu_comm_init(device, nranks, rank, id)   once per process, like MPI_INIT
 *   every rank: lart_gpu_reduce(h, root)         collective; a no-op without a communicator (single GPU)
 * libnccl.so.2 is opened at run time (LART_NCCL_LIB overrides the name); single-GPU use never needs it. */
int lart_gpu_comm_unique_id(void *id128);
int lart_gpu_comm_init(int32_t device, int32_t nranks, int32_t rank, const void *id128);
int lart_gpu_comm_info(int32_t *nranks, int32_t *rank);
int lart_gpu_comm_finalize(void);
int lart_gpu_reduce(lart_gpu_handle h, int32_t root);
int lart_gpu_destroy(lart_gpu_handle h);

/* Contiguous FP64 device buffer holding every reducible tally (Jout|Jin|Jabs|
 * Jmu|cubes|images|scalars|counters): ONE sum-reduce over it replaces the
 * reference's per-array MPI_REDUCE (src/memory_mod_mpi.f90:380-390,441-453).
 * The allph buffer is slot-per-photon-id (disjoint across ranks; also summed). */
int lart_gpu_tally_buffer(lart_gpu_handle h, void **dev_ptr, int64_t *n_doubles);
int lart_gpu_allph_buffer(lart_gpu_handle h, void **dev_ptr, int64_t *n_doubles);
/* CUDA stream (cudaStream_t) the engine launches on, for event timing. */
int lart_gpu_stream(lart_gpu_handle h, void **stream);
/* event-timed duration of the transport kernels since create/reset, ms */
int lart_gpu_kernel_ms(lart_gpu_handle h, double *ms, int64_t *launches);

/* with LART_FLAG_STAGE_TIMING: summed CUDA-event duration (ms) and launch count of
 * each stage kernel since create/reset (monolithic driver: everything in TRACE) */
int lart_gpu_stage_ms(lart_gpu_handle h, double ms[LART_STAGE_COUNT], int64_t launches[LART_STAGE_COUNT]);
/* photon slots in flight (after auto-sizing) */
int lart_gpu_pool_slots(lart_gpu_handle h, int64_t *slots);
/* FP64 FMA throughput of `device` measured with a register-resident DFMA loop,
 * TFLOP/s (2 flops per FMA): the issue-rate denominator of the FP64 roofline. */
int lart_gpu_measure_fp64(int32_t device, double *tflops);

const char *lart_gpu_last_error(void);

/* ---- unit-level batched equivalents of the finer plugin points ----------
 * (src/define.f90:741-784,820-876): arrays of photons instead of one photon.
 * All pointers are HOST pointers; n elements each. */

/* calc_voigt -> voigt_seon2 (src/line_mod.f90:38-47, src/voigt_mod.f90:541-733) */
int lart_gpu_voigt_batch(int64_t n, const double *x, const double *a, double *H);

/* raytrace_to_edge (src/raytrace_car.f90:410-508; _zonly :1138-1234).
 * icell/jcell/kcell are 1-based inputs.  Optional trace: for ray r the first
 * min(nsteps,trace_cap) visited cells, linear 0-based index, into
 * trace_cells[r*trace_cap + s] (NULL = no trace). */
int lart_gpu_raytrace_edge_batch(lart_gpu_handle h, int64_t n,
                                 const double *x, const double *y, const double *z,
                                 const double *kx, const double *ky, const double *kz,
                                 const double *xfreq,
                                 const int32_t *icell, const int32_t *jcell, const int32_t *kcell,
                                 double *tau, int32_t *nsteps,
                                 int32_t trace_cap, int32_t *trace_cells);

/* raytrace_to_tau (src/raytrace_car.f90:1425-1648; _zonly :2519-2675), without
 * the Jout tally: returns the updated photon (position, cell, frequency,
 * inside flag, and on escape the lab-frame xfreq_ref). In-place on x..kcell. */
int lart_gpu_raytrace_tau_batch(lart_gpu_handle h, int64_t n,
                                double *x, double *y, double *z,
                                const double *kx, const double *ky, const double *kz,
                                double *xfreq,
                                int32_t *icell, int32_t *jcell, int32_t *kcell,
                                const double *tau_in,
                                int32_t *inside, double *xfreq_ref, int32_t *nsteps);

/* Random variates on the path, one Philox stream per element (key = seed,
 * stream id = ids[i], starting at draw 0):
 *   kind 0: rand_number                (src/random_mt.f90:579-630 mapping)
 *   kind 1: rand_gauss                 (:964-988)
 *   kind 2: rand_resonance_vz(p0,p1)   (:2562-2696)  p0 = x, p1 = a
 *   kind 3: rand_resonance(p0)         (:2974-2993)  p0 = E1
 *   kind 4: rand_henyey_greenstein(p0) (:3022-3042)  p0 = g
 *   kind 5: rand_voigt(p0)             (:3075-3083)  p0 = a
 *   kind 6: kind 2 through the warp-cooperative sampler the scatter stage uses
 *           (same streams, same values)
 * ndraw variates per element, out[i*ndraw + j]. */
int lart_gpu_sample_batch(int32_t kind, uint64_t seed, int64_t n, const int64_t *ids,
                          const double *p0, const double *p1, int32_t ndraw, double *out);

/* car_xcrit_local (src/grid_mod_car.f90:1598-1629) */
int lart_gpu_xcrit_batch(lart_gpu_handle h, int64_t n,
                         const double *x, const double *y, const double *z,
                         const int32_t *icell, const int32_t *jcell, const int32_t *kcell,
                         double *xcrit);

/* The scatter stage's shortcut for peel-off rays, exposed for the parity tests: capped[i] = 1 when the library can
 * prove — from kappa(xfreq) times the distance to the nearest face of the ray's start cell — that raytrace_to_edge
 * would stop inside that cell with tau >= tau_huge = 745.2 (src/raytrace_car.f90:432,497), i.e. the ray's
 * contribution exp(-tau) is exactly zero whatever its direction.  Such rays are counted and not walked. */
int lart_gpu_peel_bound_batch(lart_gpu_handle h, int64_t n,
                              const double *x, const double *y, const double *z, const double *xfreq,
                              const int32_t *icell, const int32_t *jcell, const int32_t *kcell, int32_t *capped);

/* ---- next row: clump-medium ray tracers, unit level ---------------------------------
 * raytrace_to_edge_clump (src/raytrace_clump.f90:205-270; tau_max <= 0) or raytrace_to_edge_clump_capped
 * (:494-533; tau_max > 0, the peel-off call with tau_huge_clump = 745.2).  icl[i] = photon%icell_clump (0 = vacuum).
 * nclumps[i] = clumps crossed (diagnostic). */
int lart_gpu_clump_edge_batch(lart_gpu_handle h, int64_t n,
                              const double *x, const double *y, const double *z,
                              const double *kx, const double *ky, const double *kz,
                              const double *xfreq, const int32_t *icl, double tau_max,
                              double *tau, int32_t *nclumps);
/* raytrace_to_tau_clump (:83-201): x,y,z,xfreq,icl are updated in place; inside[i] = photon%inside.  No Jout tally. */
int lart_gpu_clump_tau_batch(lart_gpu_handle h, int64_t n,
                             double *x, double *y, double *z,
                             const double *kx, const double *ky, const double *kz,
                             double *xfreq, int32_t *icl, const double *tau_in, int32_t *inside);
/* active_set_at_point (src/clump_mod.f90:1595-1634), first hit: the clump a point lies in, or 0 */
int lart_gpu_clump_locate_batch(lart_gpu_handle h, int64_t n, const double *x, const double *y, const double *z,
                                int32_t *icl);

/* ---- octree AMR, unit level.  On a handle created with par.use_amr_grid the batch entry points above
 * take the LEAF index in icell (jcell = kcell = 1; icell <= 0: the leaf is located first, as raytrace_amr.f90:98-104 does):
 * lart_gpu_raytrace_edge_batch = raytrace_to_edge_amr (src/raytrace_amr.f90:265-351), lart_gpu_raytrace_tau_batch =
 * raytrace_to_tau_amr (:77-259), lart_gpu_xcrit_batch = amr_xcrit_local (src/octree_mod.f90:248-284).
 * lart_gpu_amr_locate_batch = amr_find_leaf (src/octree_mod.f90:149-171): leaf index of every point, 0 = none. */
int lart_gpu_amr_locate_batch(lart_gpu_handle h, int64_t n, const double *x, const double *y, const double *z, int32_t *il);

/* ---- next row: sight-line maps, a pure reuse of the edge walk -------------
 * make_sightline_tau_outside (src/sightline_tau_rect.f90:11-190): for every observer and detector pixel the
 * sight line enters the grid at its far boundary and runs toward the observer;
 *   tau_gas (nxfreq,nxim,nyim)  raytrace_to_edge_tau_gas  (src/raytrace_car.f90:1236-1328; gas only, no tau cap),
 *                               one walk per frequency bin centre grid%xfreq(kk) (grid_mod_car.f90:1505)
 *   N_gas   (nxim,nyim), tau_dust (nxim,nyim; DGR > 0)  raytrace_to_edge_column (:1330-1423)
 * Arrays are Fortran order (frequency fastest) and are OVERWRITTEN, as the reference assigns them; pixels whose
 * sight line misses the grid keep 0.  cross0 = line%cross0 (column density = rhokap*Dfreq/cross0 per length). */
typedef struct lart_sightline_out {
  double *tau_gas, *N_gas, *tau_dust; /* tau_dust may be NULL */
} lart_sightline_out;
int lart_gpu_sightline_tau(lart_gpu_handle h, double cross0, lart_sightline_out *out /* nobs entries */);
/* cell steps walked and CUDA-event time (ms) of the last lart_gpu_sightline_tau call */
int lart_gpu_sightline_stats(lart_gpu_handle h, double *cellsteps, double *ms);

/* library/ABI version: major*10000 + minor*100 + patch */
int lart_gpu_version(void);

#ifdef __cplusplus
}
#endif
#endif /* LART_GPU_H */
