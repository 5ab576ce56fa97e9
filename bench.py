#!/usr/bin/env python
"""bench.py — throughput of the Cartesian photon loop on B200 (and the CPU baseline beside it).

Workload (BASELINE.json configs[1], examples/sphere_peel/t4tau7.in): static uniform sphere,
201^3 cells, T = 1e4 K, tau0 = 1e7, point source, Stokes on, one observer, 201x129x129
peel-off cube.  A photon needs ~1e7 scatterings to leave this medium, so — as BASELINE.md
prescribes — both arms time a bounded photon budget and report RATES: a "step" advances every
photon slot in flight by `--quantum` scatterings (one wave = emit/trace/scatter/peel stage
kernels over the whole pool).  value = scatterings/s, whole job (all ranks).

One JSON line on stdout (rank 0).  `--impl reference` times the CPU restatement of the
reference's loop (the Fortran cannot be built in this image) on the host cores instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # BASELINE configs[1]: examples/sphere_peel/t4tau7.in
    "sphere_peel_tau1e7": dict(temperature=1e4, taumax=1e7, use_stokes=True, nx=201, ny=201, nz=201, rmax=1.0,
                               nxfreq=201, nxim=129, nyim=129, distance=1e2, save_all_photons=False),
    # the cell-by-cell core-skip variant of the same input
    "sphere_peel_tau1e7_coreskip": dict(temperature=1e4, taumax=1e7, use_stokes=True, nx=201, ny=201, nz=201, rmax=1.0,
                                        nxfreq=201, nxim=129, nyim=129, distance=1e2, core_skip=True),
    # BASELINE configs[2]: examples/vel_effect_peel/t4NHI2_20_V0200.in "with cell-by-cell core-skip" (the shipped input
    # does not set par%core_skip; BASELINE's variant adds it — SURVEY 8, config sizes) ...
    "vel_effect_peel": dict(temperature=1e4, N_HI=2e20, Vexp=200.0, velocity_type="hubble", xfreq_min=-200.0,
                            xfreq_max=40.0, nxfreq=500, use_stokes=True, comoving_source=False, nx=201, ny=201, nz=201,
                            rmax=1.0, nxim=129, nyim=129, core_skip=True),
    # ... and the input as shipped
    "vel_effect_peel_as_shipped": dict(temperature=1e4, N_HI=2e20, Vexp=200.0, velocity_type="hubble", xfreq_min=-200.0,
                                       xfreq_max=40.0, nxfreq=500, use_stokes=True, comoving_source=False, nx=201, ny=201, nz=201,
                                       rmax=1.0, nxim=129, nyim=129),
    # BASELINE configs[0]: examples/slab/t4tau7.in
    "slab_tau1e7": dict(temperature=1e4, taumax=1e7, use_stokes=True, xy_periodic=True, nx=1, ny=1, nz=201),
    # optically thinner variants of configs[1] (peel rays cross many cells: the DDA walk dominates)
    "sphere_peel_tau1e4": dict(temperature=1e4, taumax=1e4, use_stokes=True, nx=201, ny=201, nz=201, rmax=1.0,
                               nxfreq=201, nxim=129, nyim=129, distance=1e2),
    "sphere_peel_tau1e5": dict(temperature=1e4, taumax=1e5, use_stokes=True, nx=201, ny=201, nz=201, rmax=1.0,
                               nxfreq=201, nxim=129, nyim=129, distance=1e2),
    # par%xyz_symmetry: the octant of configs[1]'s sphere (101 cells with a straddling first cell = 201 across);
    # peeling-off is not allowed on folded grids (setup.f90:198), so this is the transport loop alone
    "sphere_octant_tau1e7": dict(temperature=1e4, taumax=1e7, use_stokes=True, xyz_symmetry=True, nx=101, ny=101, nz=101,
                                 rmax=1.0, nxfreq=201),
    # par%xy_symmetry: the x,y quadrant of configs[1]'s sphere, full height, with its peel-off cube
    "sphere_quadrant_tau1e7": dict(temperature=1e4, taumax=1e7, use_stokes=True, xy_symmetry=True, nx=101, ny=101, nz=201,
                                   rmax=1.0, nxfreq=201, nxim=129, nyim=129, distance=1e2),
    # par%xy_periodic with nx, ny > 1: configs[0]'s slab as a 3-D periodic box (the _xyper ray tracers)
    "box_periodic_tau1e7": dict(temperature=1e4, taumax=1e7, use_stokes=True, xy_periodic=True, geometry="rectangle",
                                nx=64, ny=64, nz=201, xmax=0.32, ymax=0.32, zmax=1.0),
    # BASELINE configs[3] family (examples/clump_sphere/clump_NHI18_fcov5.in): 7.4e6 spherical clumps of radius 1e-3 in a
    # shell 0.1 < r < 1, N_HI = 1e18 along a radial sight line, flat rotation curve, four observers' peel cubes reduced to one
    "clump_sphere_fcov5": dict(use_clump_medium=True, rmax=1.0, rmin=0.1, clump_fully_inside=False, clump_radius=0.001,
                               clump_f_cov=5.0, N_HImax=1e18, temperature=1e4, clump_sigma_v=0.0, spectral_type="monochromatic",
                               geometry="sphere", velocity_type="rotating_galaxy_halo", Vrot=300.0, rinner=0.1, nxfreq=500,
                               velocity_min=-1000.0, velocity_max=1000.0, nx=11, ny=11, nz=11, save_Jmu=True, nmu=101,
                               nxim=129, nyim=129, distance=1e4),
    # BASELINE configs[4]: examples/amr_sphere_generic (sphere_amr_inside_test1M.in / log_amr_1M.txt): octree sphere of 178 480
    # leaves (levels 3-7, the leaf list written by the reference's own generator: tests/golden/amr_sphere_l37.npz), T = 1e4 K,
    # tau_pole = 1e4, with an outside observer's peel cube instead of the HEALPix inside observer
    "amr_sphere_tau1e4": dict(amr="amr_sphere_l37", temperature=1e4, taumax=1e4, geometry="sphere", use_stokes=True, nxfreq=121,
                              nxim=100, nyim=100, distance=1e2),
    "amr_sphere_tau1e7": dict(amr="amr_sphere_l37", temperature=1e4, taumax=1e7, geometry="sphere", use_stokes=True, nxfreq=201,
                              nxim=129, nyim=129, distance=1e2),
    # ---- SURVEY 8f-1 remainder / 8f-3 / 8f-4 (round 2): the same engine behind the less common bindings
    # overlapping clump population (has_overlap): the event walk of raytrace_clump.f90:621-920 on the one-thread-per-photon driver
    "clump_overlap_fcov3": dict(use_clump_medium=True, clump_allow_overlap=True, rmax=1.0, clump_radius=0.02, clump_f_cov=3.0,
                                N_HImax=1e18, temperature=1e4, clump_sigma_v=15.0, geometry="sphere", nxfreq=201, nx=11, ny=11, nz=11,
                                velocity_min=-600.0, velocity_max=600.0, nxim=129, nyim=129, distance=1e2),
    # configs[1] with the reference's -DCALCJ -DCALCP -DCALCPnew build: J(x, r), Pa(r), Pnew(r) in 101 radial shells
    "sphere_tau1e7_calcJP_radial": dict(temperature=1e4, taumax=1e7, use_stokes=True, nx=201, ny=201, nz=201, rmax=1.0, nxfreq=201,
                                        nxim=129, nyim=129, distance=1e2, calc_J=True, calc_P=True, calc_Pnew=True, geometry_JPa=1),
    # ... and per cell: J(nxfreq, nx, ny, nz) = 3.3 GB in HBM at 51 frequency bins (13 GB at the input's 201)
    "sphere_tau1e7_calcJP_cells": dict(temperature=1e4, taumax=1e7, use_stokes=True, nx=201, ny=201, nz=201, rmax=1.0, nxfreq=51,
                                       nxim=129, nyim=129, distance=1e2, calc_J=True, calc_P=True, calc_Pnew=True, geometry_JPa=3),
    # shearing box: configs[0]'s slab as a periodic 64 x 64 x 201 box with par%Omega (raytrace_to_tau_car_xyper_shear)
    "box_shear_tau1e7": dict(temperature=1e4, taumax=1e7, use_stokes=True, xy_periodic=True, geometry="rectangle", nx=64, ny=64, nz=201,
                             xmax=0.32, ymax=0.32, zmax=1.0, Omega=2.0),
    # exoplanet atmospheres (exponential profiles; plane-parallel illumination)
    "plane_atmosphere_tau1e6": dict(geometry="plane_atmosphere", source_geometry="plane_illumination", temperature=1e4, taumax=1e6,
                                    nz=201, zmax=1.0, density_zscale=0.3, use_stokes=True, nxfreq=201),
    "spherical_atmosphere_tau1e6": dict(geometry="spherical_atmosphere", source_geometry="plane_illumination", temperature=1e4, taumax=1e6,
                                        nx=201, ny=201, nz=201, rmax=1.0, rmin=0.4, density_rscale=0.3, use_stokes=True, nxfreq=201,
                                        nxim=129, nyim=129, distance=1e2),
    # small case for smoke-testing the bench itself
    "tiny": dict(temperature=1e4, taumax=1e5, use_stokes=True, nx=41, ny=41, nz=41, rmax=1.0, nxfreq=61, nxim=33, nyim=33),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gpu", choices=["gpu", "reference"])
    ap.add_argument("--workload", default="sphere_peel_tau1e7", choices=sorted(WORKLOADS))
    ap.add_argument("--quantum", type=int, default=32, help="scatterings per photon slot per step")
    ap.add_argument("--pool-slots", type=int, default=0, help="photons in flight per GPU (0 = auto: 148*16384)")
    ap.add_argument("--flags", type=int, default=0, help="LART_FLAG_* bits (1 SoA grid, 2 no warp aggregation, 4 monolithic)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="device-resident arm only (profiling runs)")
    ap.add_argument("--seed", type=int, default=12345)
    ap.add_argument("--streams", type=int, default=0, help="wave pipelines (pool partitions on separate streams); 0 = auto")
    ap.add_argument("--complete-workload", default="sphere_peel_tau1e4", choices=sorted(WORKLOADS),
                    help="workload of the `complete_run` record: every photon to its escape through lart_gpu_run")
    ap.add_argument("--deal-batch", type=int, default=0,
                    help="complete run under --gpus N: 0 = static photon partition (ids rank+1 : N : nranks, run_simulation_mod.f90:150); "
                         "> 0 = dynamic dealing, ranks claim batches of this many ids from a shared-memory counter (lart_gpu_run_dealt, "
                         "the master/worker mode of run_simulation_mod.f90:31-128)")
    ap.add_argument("--complete-photons", type=float, default=1e6,
                    help="photons of the complete run, TOTAL over all GPUs (strong scaling under --gpus N); 0 = skip")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.stop, self.t = index, [], threading.Event(), None

    def _loop(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.splitlines()[0].split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 3 + k and r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]), "samples": len(self.rows),
                "power_w_max": max(float(r[2]) for r in self.rows), "reasons": reasons}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def build_model(args, nphotons, workload=None):
    from lart_b200 import Model
    kw = dict(WORKLOADS[workload or args.workload])
    fixture = kw.pop("amr", None)
    m = Model(no_photons=nphotons, iseed=args.seed, **kw)
    if fixture:  # octree leaves in the reference's generic format (tools/make_amr_fixtures.py)
        d = np.load(os.path.join(ROOT, "tests", "golden", fixture + ".npz"))
        lev = d["level"].astype(np.int32)
        xyz = [-1.0 + (2.0 * d[k] + 1.0) / 2.0 ** lev for k in ("ix", "iy", "iz")]
        m.set_amr_leaves(xyz[0], xyz[1], xyz[2], lev, d["dens"].astype(np.float64), kw.get("temperature", 1e4), boxlen=float(d["boxlen"]))
    return m.setup()


def workload_config(args, model):
    """`config` of the JSON line: names the workload and nothing else, so that both arms print the same object."""
    cfg = model.config.contents
    g = cfg.grid
    ncell = cfg.amr.nleaf if cfg.par.use_amr_grid else g.nx * g.ny * g.nz
    amr = cfg.amr
    return {"workload": args.workload, "grid": [g.nx, g.ny, g.nz] if not cfg.par.use_amr_grid else {"octree_leaves": amr.nleaf, "cells": amr.ncells},
            "tau0": WORKLOADS[args.workload].get("taumax"),
            "peel_cube": [g.nxfreq, cfg.observers[0].nxim, cfg.observers[0].nyim] if cfg.par.nobs else None,
            "quantum": args.quantum, "seed": args.seed,
            "step": "every photon in flight advances by `quantum` scatterings (bounded sample of the workload: a photon needs "
                    "~1e7 scatterings to leave the tau0 = 1e7 medium); value = scatterings/s",
            "l2": "working set (grid %.0f MB as the host's SoA arrays, %.0f MB packed on the device, + photon pool + ray queue + cubes) "
                  "larger than the 126 MB L2; no flush" % (48.0 * ncell / 1e6, 64.0 * ncell / 1e6),
            "parallelism": "photon ids strided over the ranks / host threads (run_simulation_mod.f90:150); grid replicated; "
                           "tallies summed once at the end"}


def cpu_sample(args, events_per_photon, seconds, threads=None):
    """Time the CPU restatement (oracle, MT19937-64 like the reference) on a bounded sample of the
    same workload: the first `events_per_photon` scatterings of n photons, n sized for ~`seconds`."""
    from oracle import oracle
    flags = oracle.use_native()
    threads = threads or oracle.hardware_threads()
    m = build_model(args, 10 ** 6)
    n = 4000 * threads  # calibration pass, long enough to amortise thread start-up
    t0 = time.perf_counter()
    oracle.run(m, rng_mode=0, nthreads=threads, first_id=1, count=n, max_events=events_per_photon, seed=args.seed)
    dt = time.perf_counter() - t0
    rate = m.counters["n_scatter"] / dt
    n = int(max(threads, min(200_000_000, rate * seconds / max(events_per_photon, 1))))
    n = (n // threads) * threads
    m.zero_tallies()
    t0 = time.perf_counter()
    oracle.run(m, rng_mode=0, nthreads=threads, first_id=1, count=n, max_events=events_per_photon, seed=args.seed)
    dt = time.perf_counter() - t0
    c = m.counters
    return {"value": c["n_scatter"] / dt, "unit": "scatterings/s", "cores": threads, "kind": "port",
            "sample": "%s: first %d scatterings of %d photons on %d threads, %.1f s (C++ restatement of the reference loop, "
                      "%s built on this host, MT19937-64; the Fortran+MPI build is impossible in this image)"
                      % (args.workload, events_per_photon, n, threads, dt, flags),
            "compiler": flags, "cellsteps_per_s": c["n_cellsteps"] / dt, "seconds": dt, "photons": n, "_model": m}


def cpu_complete(args, seconds, threads=None):
    """The CPU port on a bounded sample of the complete-run workload: n photons from emission to escape."""
    from oracle import oracle
    flags = oracle.use_native()
    threads = threads or oracle.hardware_threads()
    m = build_model(args, 10 ** 6, args.complete_workload)
    n = 8 * threads
    t0 = time.perf_counter()
    oracle.run(m, rng_mode=0, nthreads=threads, first_id=1, count=n, seed=args.seed)
    dt = time.perf_counter() - t0
    n = int(max(threads, min(10 ** 6, n * seconds / max(dt, 1e-3)))) // threads * threads
    m.zero_tallies()
    t0 = time.perf_counter()
    oracle.run(m, rng_mode=0, nthreads=threads, first_id=1, count=n, seed=args.seed)
    dt = time.perf_counter() - t0
    c = m.counters
    return {"photons": n, "wall_s": dt, "photons_per_s": n / dt, "scatterings_per_s": c["n_scatter"] / dt,
            "mean_nscatt": c["n_scatter"] / n, "cores": threads, "kind": "port", "compiler": flags}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ev = args.quantum
    per_step = max(1.0, min(20.0, 150.0 / max(args.steps + args.warmup, 1)))
    for _ in range(args.warmup):
        cpu_sample(args, ev, min(per_step, 2.0))
    tot_s, tot_t, last = 0.0, 0.0, None
    for _ in range(args.steps):
        last = cpu_sample(args, ev, per_step)
        tot_s += last["value"] * last["seconds"]
        tot_t += last["seconds"]
    v = tot_s / tot_t
    cb = dict(last)
    model = cb.pop("_model")
    cb["value"] = v
    out = {
        "impl": "reference", "metric": "scatterings_per_s", "value": v, "unit": "scatterings/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, model),
        "cpu_baseline": cb, "e2e": {"value": v, "unit": "scatterings/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}
    if args.complete_photons > 0:
        out["complete_run"] = dict(cpu_complete(args, 20.0), workload=args.complete_workload)
    print(json.dumps(out))


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from lart_b200 import Simulation, capi, measure_fp64
    from lart_b200.host import comm_init_torch, comm_finalize, photon_partition

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        comm_init_torch(local)  # the engine's own NCCL communicator (lart_gpu_reduce); torch only carries the 128-byte id

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    S_auto = args.pool_slots or 148 * 16384
    total_steps = args.warmup + args.steps
    nph = S_auto * world * 4  # far more ids than slots: the queue never runs dry, nobody finishes tau0 = 1e7 anyway
    model = build_model(args, nph)
    cfg = model.config.contents
    g = cfg.grid
    ncell = cfg.amr.nleaf if cfg.par.use_amr_grid else g.nx * g.ny * g.nz
    grid_bytes = 8 * (6 * ncell + (g.nx + g.ny + g.nz + 3)) if not cfg.par.use_amr_grid else \
        8 * 6 * cfg.amr.nleaf + cfg.amr.ncells * (4 * 8 + 4 * 15) + 4 * cfg.amr.nleaf

    # ------------------------- device-resident arm: `value`  (CUDA-graph launches, no per-stage events)
    first, count, stride = rank + 1, nph // world, world  # run_simulation_mod.f90:150 partition

    def device_arm(flags, warmup, steps, streams):
        sim = Simulation(model, device=local, pool_slots=args.pool_slots, quantum=args.quantum, flags=flags,
                         streams=streams)
        sim.begin(first, count, stride)
        for _ in range(warmup):
            sim.step(args.quantum)
        sim.sync()
        sim.reset_tallies()
        barrier()
        with ClockSampler(local) as clk_:
            t0 = time.perf_counter()
            for _ in range(steps):
                sim.step(args.quantum)
            sim.sync()
            barrier()
            wall_ = time.perf_counter() - t0
        ms_, launches_ = sim.kernel_ms()
        stage_ = sim.stage_ms()
        model.zero_tallies()
        sim._check(sim._lib.lart_gpu_fetch(sim._h, model.tallies))  # counters of the timed region only
        c_ = dict(model.counters)
        S_ = sim.pool_slots
        sim.close()
        return ms_, launches_, stage_, c_, S_, wall_, clk_

    dev_ms, launches, _, c, S, wall, clk = device_arm(args.flags, args.warmup, args.steps, args.streams)
    keys = ["n_scatter", "n_cellsteps", "n_peel", "n_photons_done", "n_rng", "n_peel_bound", "n_cellsteps_bound"]
    t = torch.tensor([dev_ms] + [c[k] for k in keys] + [float(launches)], dtype=torch.float64, device="cuda")
    tmax = t.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    max_ms = float(tmax[0])
    tot = {k: float(t[1 + i]) for i, k in enumerate(keys)}
    n_launch = float(t[-1])
    value = tot["n_scatter"] / (max_ms * 1e-3)
    # a second, short pass with CUDA events around every stage kernel feeds the roofline: plain launches on ONE
    # stream, so that every kernel is timed alone (in the value arm the partitions' kernels overlap)
    dev_ms, _, stage, c, _, _, _ = device_arm(args.flags | capi.FLAG_STAGE_TIMING, max(args.warmup, 3), min(args.steps, 4), 1)

    # ------------------------- roofline (rank 0's numbers), SURVEY.md section 8(d)
    hbm_peak, peak_src = measured_peaks()
    fp64_peak = measure_fp64(local)
    mono = bool(args.flags & 4)
    clump = bool(cfg.par.use_clump_medium)
    names = {"emit": "k_cl_emit" if clump else "k_wf_emit", "trace": "k_cl_flight" if clump else "k_wf_trace",
             "draw": "k_cl_scatter" if clump else ("k_wf_draw" if args.flags & (16 | 128) else "k_wf_draw2"), "apply": "k_wf_apply",
             "peel": "k_cl_peel" if clump else "k_wf_peel"}
    if mono:
        names["trace"] = "k_mono_clump" if clump else "k_mono"
    tot_stage = max(sum(v[0] for v in stage.values()), 1e-30)
    dom = max(stage, key=lambda k: stage[k][0])
    walked = c["n_cellsteps"] - c["n_cellsteps_bound"]  # steps a walker kernel actually took
    bcs = 62.0 if clump else (56.0 if cfg.par.DGR > 0 else 48.0)  # algorithmic bytes per cell step (8d)
    walk_ms = stage["trace"][0] + stage["peel"][0]
    walk_n = max(stage["trace"][1], 1)
    nobs = max(int(cfg.par.nobs), 0)
    tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")  # dram__bytes_read+write per launch from `ncu --set full`
    tj = json.load(open(tpath)) if os.path.exists(tpath) else {}
    same_wl = bool(tj) and tj.get("workload") == args.workload and not mono
    per_kernel = {}
    for k in stage:
        ms_k, n_k = stage[k]
        ent = {"kernel": names[k], "ms_per_launch": ms_k / max(n_k, 1), "share_of_wave": ms_k / tot_stage}
        if same_wl and names[k] in tj.get("dram_bytes_per_scattering", {}):
            ent["dram_bytes_per_scattering_ncu"] = tj["dram_bytes_per_scattering"][names[k]]
        per_kernel[k] = ent
    wave_ms = tot_stage / walk_n  # one wave = one launch of each stage kernel (serial sum)
    scat_per_wave = c["n_scatter"] / walk_n
    alg_bytes_wave = bcs * walked / walk_n
    ach = alg_bytes_wave / (wave_ms * 1e-3) / 1e9
    traffic = None
    if same_wl:
        traffic = sum(tj["dram_bytes_per_scattering"].values()) * scat_per_wave
    flops = 50.0 * walked + 300.0 * c["n_scatter"]
    roof = {
        # the contract's figure: algorithmic grid bytes of the cell walk (48 B per WALKED cell step) over the duration of one wave
        # (one launch of each stage kernel, timed alone with CUDA events) against the measured HBM copy rate
        "bound": "hbm", "kernel": "one wave: " + " + ".join(names[k] for k in stage if stage[k][0] > 0),
        "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": traffic,
        "peak_source": peak_src, "bytes_per_cellstep": bcs, "walked_cellsteps_per_wave": walked / walk_n,
        "scatterings_per_wave": scat_per_wave, "avg_launch_ms": wave_ms,
        "traffic_over_algorithmic": (traffic / alg_bytes_wave) if traffic else None,
        "note": "this path is bound by FP64/integer instruction issue and latency, not by HBM: at tau0 = 1e7 a scattering walks ~1 cell "
                "(48 B) but costs ~300 flop + 6 libm calls + ~16 Philox blocks; `traffic` is dominated by the wavefront's photon records "
                "(read+written once per stage), see `dram_traffic`",
        "dominant_stage": dom, "per_kernel": per_kernel,
        "hbm_algorithmic": {"bytes_per_cellstep": bcs, "walked_cellsteps_per_s": walked / (dev_ms * 1e-3),
                            "GBps": bcs * walked / (dev_ms * 1e-3) / 1e9, "frac_of_hbm_peak": bcs * walked / (dev_ms * 1e-3) / 1e9 / hbm_peak},
        "dda_walk": {"kernels": names["trace"] + "+" + names["peel"], "bytes_per_cellstep": bcs,
                     "walked_cellsteps": walked, "bound_skipped_cellsteps": c["n_cellsteps_bound"],
                     "achieved_GBps": bcs * walked / (walk_ms * 1e-3) / 1e9 if walk_ms > 0 else 0.0,
                     "frac_of_hbm_peak": bcs * walked / (walk_ms * 1e-3) / 1e9 / hbm_peak if walk_ms > 0 else 0.0,
                     "walked_cellsteps_per_launch": walked / walk_n, "avg_launch_ms": walk_ms / walk_n},
        "fp64_model": {"achieved_tflops": flops / (dev_ms * 1e-3) / 1e12, "peak_tflops": fp64_peak,
                       "frac": flops / (dev_ms * 1e-3) / 1e12 / fp64_peak if fp64_peak > 0 else None,
                       "model": "50 flop per walked cell step + 300 flop per scattering (SURVEY.md 8d; libm calls and the RNG not counted)",
                       "peak_source": "measured DFMA loop (lart_gpu_measure_fp64)"},
        "dram_traffic": ({"source": "profiles/r2_traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)",
                          "bytes_per_scattering": tj["dram_bytes_per_scattering"],
                          "total_bytes_per_scattering": sum(tj["dram_bytes_per_scattering"].values()),
                          "algorithmic_bytes_per_scattering": bcs * walked / max(c["n_scatter"], 1)} if same_wl else None),
    }

    # ------------------------- end-to-end arm through the public API with HOST buffers: `e2e`
    # timed: lart_gpu_create (H2D of the host grid arrays; a cold handle: the library's memory pool was trimmed when the
    # last handle closed) + begin + steps (each followed by the D2H read of its result: photons in flight) +
    # output_reduce (ONE ncclReduce of the tally buffer inside the C ABI + D2H into the host tallies).
    model.zero_tallies()
    barrier()
    t0 = time.perf_counter()
    buf_n = 0
    if not args.skip_e2e:
        sim = Simulation(model, device=local, pool_slots=args.pool_slots, quantum=args.quantum, flags=args.flags,
                         streams=args.streams)
        ta = time.perf_counter()
        sim.begin(first, count, stride)
        for _ in range(total_steps):
            sim.step(args.quantum)
        tb = time.perf_counter()
        sim.output_reduce(dst=0)
        barrier()
        tc = time.perf_counter()
        buf_n = sim.tally_buffer()[1]
        e2e_s = tc - t0  # results are in the host arrays here; releasing the device memory is not part of the job
        phases = {"create_s": ta - t0, "steps_s": tb - ta, "reduce_fetch_s": tc - tb}
        sim.close()
        print("e2e phases: create %.3f s, steps %.3f s, reduce+fetch %.3f s, destroy (untimed) %.3f s" %
              (ta - t0, tb - ta, tc - tb, time.perf_counter() - tc), file=sys.stderr)
    else:
        e2e_s = time.perf_counter() - t0
        phases = None
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    ns = torch.tensor([model.counters["n_scatter"] if rank == 0 else 0.0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.broadcast(ns, src=0)  # rank 0 holds the reduced counters of all ranks
    e2e_value = float(ns[0]) / float(te[0])

    # ------------------------- complete run: a fixed number of photons, every one to its escape, through lart_gpu_run
    # (BASELINE's other unit, photons/s).  The TOTAL is fixed: under --gpus N it is split over the ranks = strong scaling.
    complete = None
    if args.complete_photons > 0 and not args.skip_e2e:
        ntot = int(args.complete_photons)
        cm = build_model(args, ntot, args.complete_workload)
        barrier()
        t0 = time.perf_counter()
        sim = Simulation(cm, device=local, pool_slots=args.pool_slots, flags=args.flags & ~capi.FLAG_STAGE_TIMING, streams=args.streams)
        dealt = None
        if args.deal_batch > 0 and world > 1:
            deal_name = "/lart_bench_deal_%s" % os.environ.get("MASTER_PORT", "0")
            if rank == 0:  # create and zero the node's counter before anybody claims from it
                sim.open_deal(deal_name, create=True)
            barrier()
            dealt = sim.run_simulation_dealt(deal_name, ntot, batch=args.deal_batch)
        else:
            sim.run_simulation(rank, world, ntot)
        t_run = time.perf_counter() - t0
        sim.output_reduce(dst=0)
        barrier()
        t_all = time.perf_counter() - t0
        if dealt is not None and rank == 0:
            sim.unlink_deal(deal_name)
        dms, _ = sim.kernel_ms()
        sim.close()
        tt = torch.tensor([t_all, t_run, dms * 1e-3], dtype=torch.float64, device="cuda")
        tmin = tt.clone()
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        if rank == 0:
            cc = cm.counters
            cm.output_normalize()
            sm_ = cm.summary
            flux = None
            if cm.config.contents.par.nobs:
                omega = sm_.dxim * sm_.dyim * (np.pi / 180) ** 2
                flux = float((cm.observer_cube("scatt").sum() + cm.observer_cube("direc").sum()) * 4 * np.pi * omega
                             * sm_.distance ** 2 * sm_.dxfreq)
            complete = {"workload": args.complete_workload, "photons": ntot, "n_gpus": world, "scaling": "strong",
                        "wall_s": float(tt[0]), "photons_per_s": ntot / float(tt[0]), "scatterings_per_s": cc["n_scatter"] / float(tt[0]),
                        "mean_nscatt": cc["n_scatter"] / ntot, "run_s_max_rank": float(tt[1]), "run_s_min_rank": float(tmin[1]),
                        "device_s_max_rank": float(tt[2]), "flux_check": flux,
                        "photon_partition": ("dynamic: batches of %d ids claimed from a shared counter (lart_gpu_run_dealt)" % args.deal_batch)
                        if dealt is not None else "static: ids rank+1 : N : nranks",
                        "region": "lart_gpu_create (H2D grid) + lart_gpu_run to the last photon (ids rank+1 : N : nranks) + "
                                  "lart_gpu_reduce (NCCL) + lart_gpu_fetch into host arrays; wall = max over ranks"}

    if rank == 0:
        out = {
            "metric": "scatterings_per_s", "value": value, "unit": "scatterings/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, model),
            "engine": {"photons_in_flight_per_gpu": S, "driver": "monolithic" if mono else "wavefront",
                       "wave": "emit / trace / draw / apply / peel stage kernels over the photon pool; steps are CUDA-graph launches"},
            "cellsteps_per_s": (tot["n_cellsteps"] - tot["n_cellsteps_bound"]) / (max_ms * 1e-3),
            "cellsteps_bound_skipped_per_s": tot["n_cellsteps_bound"] / (max_ms * 1e-3),
            "peel_rays_per_s": tot["n_peel"] / (max_ms * 1e-3), "peel_rays_bound_skipped_per_s": tot["n_peel_bound"] / (max_ms * 1e-3),
            "photons_done": tot["n_photons_done"], "uniforms_per_s": tot["n_rng"] / (max_ms * 1e-3), "wall_s": wall,
            "gpu_launches": int(n_launch), "roofline": roof,
            "e2e": {"value": e2e_value, "unit": "scatterings/s", "h2d_bytes_per_step": grid_bytes / total_steps,
                    "d2h_bytes_per_step": (8 * buf_n + 8 * total_steps) / total_steps, "seconds": float(te[0]),
                    "phases_rank0": phases,
                    "region": "lart_gpu_create (cold handle: device allocations + H2D of the host grid) + begin + %d steps (+D2H of each "
                              "step's in-flight count) + lart_gpu_reduce (NCCL) + D2H of the tally buffer into host arrays "
                              "(lart_gpu_destroy not timed)" % total_steps},
            "clocks": clk.summary(),
        }
        if complete:
            out["complete_run"] = complete
        if not args.no_cpu_baseline and world == 1:
            cb = cpu_sample(args, args.quantum, args.cpu_seconds)
            cb.pop("_model")
            out["cpu_baseline"] = cb
            if complete:
                out["complete_run"]["cpu"] = cpu_complete(args, 10.0)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        comm_finalize()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
