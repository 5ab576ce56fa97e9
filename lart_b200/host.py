"""Host-side mirror of LaRT's driver sequence for the Cartesian photon loop.

The reference's main program (src/main.f90:19-57) runs

    read_input -> setup_procedure -> grid_create -> observer_create
    -> run_simulation(grid) -> output_reduce(grid) -> output_normalize(grid) -> write_output

`Model` covers the steps before `run_simulation` through the C++ mini-host
(liblart_host.so, include/lart_host.h); `Simulation` is the replacement for
`run_simulation` + `output_reduce`: it drives the CUDA engine through the C ABI
(liblart_gpu.so, include/lart_gpu.h) exactly as the Fortran shim
(shim/lart_gpu_shim.f90) does.  Names follow the reference (par, grid, observer,
Jout, Jin, nscatt_gas ...).  No PyTorch on this path: the multi-GPU reduce is the
C ABI's own NCCL call (lart_gpu_reduce); torch.distributed is at most the host-side
broadcast of the communicator id (comm_init_torch), as MPI_BCAST is in the shim.
"""
import ctypes as C

import numpy as np

from . import capi


class LartError(RuntimeError):
    pass


def _np(ptr, shape):
    """numpy view (no copy) of a double* owned by the mini-host; None for NULL."""
    if not ptr:
        return None
    n = int(np.prod(shape))
    return np.ctypeslib.as_array(ptr, shape=(n,)).reshape(shape, order="F")


class Model:
    """namelist -> resolved par/grid/line/observers (setup.f90 + grid_mod_car.f90 + observer_rect.f90)."""

    def __init__(self, infile=None, **par):
        self._lib = capi.load_host()
        self._m = C.c_void_p(self._lib.lart_host_new())
        self._setup = False
        if infile is not None:
            self.read_input(infile)
        for k, v in par.items():
            self.set(k, v)

    def __del__(self):
        try:
            if self._m:
                self._lib.lart_host_free(self._m)
                self._m = None
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise LartError(self._lib.lart_host_last_error().decode())

    @staticmethod
    def _fmt(v):
        if isinstance(v, bool):
            return ".true." if v else ".false."
        if isinstance(v, (list, tuple, np.ndarray)):
            return " ".join(repr(float(x)) for x in v)
        if isinstance(v, str):
            return "'%s'" % v
        return repr(v)

    def set(self, key, value):
        """`par%key = value`, as one line of the &parameters namelist (setup.f90:27-40)."""
        self._setup = False
        self._check(self._lib.lart_host_set(self._m, key.encode(), self._fmt(value).encode()))
        return self

    def read_input(self, path):
        self._setup = False
        self._check(self._lib.lart_host_read_input(self._m, str(path).encode()))
        return self

    def set_amr_leaves(self, x, y, z, level, nH, T, vx=None, vy=None, vz=None, boxlen=2.0, origin=None):
        """Leaf cells of an octree in the reference's generic AMR format (what generic_amr_read returns,
        read_generic_amr.f90:52-346); sets par%use_amr_grid.  origin = lower corner (default: box centred on 0)."""
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        x, y, z, nH = f(x), f(y), f(z), f(nH)
        T = f(np.broadcast_to(T, x.shape))
        level = np.ascontiguousarray(level, dtype=np.int32)
        v = [None if a is None else f(a) for a in (vx, vy, vz)]
        dp = lambda a: None if a is None else a.ctypes.data_as(capi.c_double_p)
        o = (-boxlen / 2.0,) * 3 if origin is None else origin
        self._setup = False
        self._check(self._lib.lart_host_set_amr_leaves(self._m, x.size, dp(x), dp(y), dp(z), level.ctypes.data_as(capi.c_int32_p),
                                                       dp(nH), dp(T), dp(v[0]), dp(v[1]), dp(v[2]), float(boxlen),
                                                       float(o[0]), float(o[1]), float(o[2])))
        return self

    def setup(self):
        self._check(self._lib.lart_host_setup(self._m))
        self._setup = True
        return self

    @property
    def config(self):
        """The lart_config handed to lart_gpu_create (pointer target owned by the model)."""
        if not self._setup:
            self.setup()
        return self._lib.lart_host_config(self._m)

    @property
    def summary(self):
        if not self._setup:
            self.setup()
        s = capi.HostSummary()
        self._check(self._lib.lart_host_get_summary(self._m, C.byref(s)))
        return s

    @property
    def tallies(self):
        if not self._setup:
            self.setup()
        return self._lib.lart_host_tallies(self._m)

    def zero_tallies(self):
        self._check(self._lib.lart_host_zero_tallies(self._m))

    def output_normalize(self):
        """output_normalize_outside (output_sum_rect.f90:151-487), in place."""
        self._check(self._lib.lart_host_normalize(self._m))

    # ---- numpy views of the grid and the tallies (Fortran order, as grid_type holds them)
    def grid_array(self, name):
        g = self.config.contents.grid
        shape = {"xface": (g.nx + 1,), "yface": (g.ny + 1,), "zface": (g.nz + 1,)}.get(name, (g.nx, g.ny, g.nz))
        return _np(getattr(g, name), shape)

    def spectrum(self, name="Jout"):
        g = self.config.contents.grid
        t = self.tallies.contents
        if name == "Jmu":
            return _np(t.Jmu, (g.nxfreq, self.config.contents.par.nmu))
        return _np(getattr(t, name), (g.nxfreq,))

    def jp_array(self, name):
        """CALCJ / CALCP / CALCPnew accumulators (grid%J|J2|J1, Pa|P2|P1, Pa_new|...; grid_mod_car.f90:1385-1434):
        name 'J' -> (nxfreq, bins...), 'Pa' / 'Pnew' -> (bins...), bins by par%geometry_JPa: 3 (nx,ny,nz), 2 (nr,nz),
        1 (nr,), -1 (nz,).  None when the accumulator is off."""
        g = self.config.contents.grid
        bins = {3: (g.nx, g.ny, g.nz), 2: (g.nr, g.nz), 1: (g.nr,), -1: (g.nz,)}.get(g.geometry_JPa)
        ptr = getattr(self.tallies.contents, name)
        if not ptr or bins is None:
            return None
        return _np(ptr, ((g.nxfreq,) + bins) if name == "J" else bins)

    def xfreq(self):
        """bin centres, grid%xfreq (grid_mod_car.f90:1505)."""
        g = self.config.contents.grid
        return (np.arange(g.nxfreq) + 0.5) * g.dxfreq + g.xfreq_min

    def observer_cube(self, name, iobs=0):
        cfg = self.config.contents
        t = self.tallies.contents
        if not t.obs:
            return None
        ob = cfg.observers[iobs]
        shape = (ob.nxim, ob.nyim) if name.endswith("_2D") else (cfg.grid.nxfreq, ob.nxim, ob.nyim)
        return _np(getattr(t.obs[iobs], name), shape)

    def allph(self, name):
        t = self.tallies.contents
        return _np(getattr(t.allph, name), (self.config.contents.par.nphotons,))

    @property
    def counters(self):
        c = self.tallies.contents.counters
        return {n: getattr(c, n) for n in capi.COUNTER_FIELDS}

    @property
    def nscatt_gas(self):
        return self.tallies.contents.nscatt_gas

    @property
    def nscatt_dust(self):
        return self.tallies.contents.nscatt_dust


def photon_partition(nphotons, rank, nproc):
    """Photon ids of `rank`: rank+1, rank+1+nproc, ... <= nphotons (run_simulation_mod.f90:150).
    Returns (first_id, count, stride)."""
    n = int(nphotons)
    count = (n - rank + nproc - 1) // nproc if n > rank else 0
    return rank + 1, count, nproc


class Simulation:
    """One more implementation of `run_sim` (define.f90:832-838) on one GPU.

    run_simulation(rank, nproc) runs photon ids rank+1, rank+1+nproc, ... as
    run_equal_number does (run_simulation_mod.f90:150); output_reduce() adds the
    device tallies into the model's host buffers (output_sum_rect.f90:7-149),
    after ONE sum-reduce over NCCL (lart_gpu_reduce) when the process has a communicator (comm_init).
    """

    def __init__(self, model, device=0, pool_slots=0, quantum=0, flags=0, seed=None, streams=0, ray_budget=0, max_events=0):
        self._lib = capi.load_gpu()  # raises if the CUDA engine is missing: no fallback
        self.model = model
        cfg = model.config.contents
        cfg.device = device
        cfg.pool_slots = pool_slots
        cfg.quantum = quantum
        cfg.flags = flags
        cfg.streams = streams
        cfg.ray_budget = ray_budget
        cfg.max_events = max_events
        if seed is not None:
            cfg.par.seed = seed
        self._h = C.c_void_p()
        self._check(self._lib.lart_gpu_create(C.byref(cfg), C.byref(self._h)))

    def _check(self, rc):
        if rc != 0:
            raise LartError(self._lib.lart_gpu_last_error().decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.lart_gpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run_simulation(self, rank=0, nproc=1, nphotons=None):
        n = int(self.model.config.contents.par.nphotons if nphotons is None else nphotons)
        first, count, stride = photon_partition(n, rank, nproc)
        self._check(self._lib.lart_gpu_run(self._h, first, count, stride))

    def open_deal(self, deal_name, create=False):
        """Create (and zero) or just touch the node's shared photon counter; the handle is closed again at once —
        the name persists until unlink_deal."""
        d = C.c_void_p()
        self._check(self._lib.lart_gpu_deal_open(deal_name.encode(), 1 if create else 0, C.byref(d)))
        self._lib.lart_gpu_deal_close(d, 0)

    def unlink_deal(self, deal_name):
        d = C.c_void_p()
        self._check(self._lib.lart_gpu_deal_open(deal_name.encode(), 0, C.byref(d)))
        self._lib.lart_gpu_deal_close(d, 1)

    def run_simulation_dealt(self, deal_name, nphotons=None, batch=65536, create=False):
        """Master/worker mode (run_simulation_mod.f90:31-128) without a master: the node's processes claim batches of
        photon ids from one shared counter (POSIX shared memory `deal_name`, '/...') whenever their queue runs dry.
        create=True on exactly one process, before the others open it (host barrier).  Returns the photons this
        process ran."""
        n = int(self.model.config.contents.par.nphotons if nphotons is None else nphotons)
        d, mine = C.c_void_p(), C.c_int64()
        self._check(self._lib.lart_gpu_deal_open(deal_name.encode(), 1 if create else 0, C.byref(d)))
        try:
            self._check(self._lib.lart_gpu_run_dealt(self._h, d, n, int(batch), C.byref(mine)))
        finally:
            self._lib.lart_gpu_deal_close(d, 0)
        return mine.value

    def begin(self, first_id, count, stride=1):
        self._check(self._lib.lart_gpu_begin(self._h, first_id, count, stride))

    def step(self, quantum=0):
        left = C.c_int64()
        self._check(self._lib.lart_gpu_step(self._h, quantum, C.byref(left)))
        return left.value

    def sync(self):
        self._check(self._lib.lart_gpu_sync(self._h))

    def reset_tallies(self):
        self._check(self._lib.lart_gpu_reset_tallies(self._h))

    def kernel_ms(self):
        ms, n = C.c_double(), C.c_int64()
        self._check(self._lib.lart_gpu_kernel_ms(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def stage_ms(self):
        """{stage: (summed ms, launches)} — needs flags & FLAG_STAGE_TIMING."""
        ms = (C.c_double * len(capi.STAGES))()
        n = (C.c_int64 * len(capi.STAGES))()
        self._check(self._lib.lart_gpu_stage_ms(self._h, ms, n))
        return {capi.STAGES[k]: (ms[k], n[k]) for k in range(len(capi.STAGES))}

    @property
    def pool_slots(self):
        n = C.c_int64()
        self._check(self._lib.lart_gpu_pool_slots(self._h, C.byref(n)))
        return n.value

    def tally_buffer(self):
        p, n = C.c_void_p(), C.c_int64()
        self._check(self._lib.lart_gpu_tally_buffer(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def allph_buffer(self):
        p, n = C.c_void_p(), C.c_int64()
        self._check(self._lib.lart_gpu_allph_buffer(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def stream(self):
        s = C.c_void_p()
        self._check(self._lib.lart_gpu_stream(self._h, C.byref(s)))
        return s.value

    def output_reduce(self, dst=0):
        """output_reduce (output_sum_rect.f90:7-149): sum the tallies of all ranks onto `dst` — ONE NCCL reduce over the
        contiguous device tally buffer inside the C ABI (lart_gpu_reduce; a no-op without a communicator, see comm_init)
        — and add them into the host arrays on `dst`."""
        self._check(self._lib.lart_gpu_reduce(self._h, dst))
        if comm_rank() == dst:
            self._check(self._lib.lart_gpu_fetch(self._h, self.model.tallies))

    def sightline_tau(self):
        """make_sightline_tau_outside (sightline_tau_rect.f90:11-190): per observer the maps
        tau_gas(nxfreq,nxim,nyim), N_gas(nxim,nyim) and, with dust, tau_dust(nxim,nyim)."""
        cfg = self.model.config.contents
        nobs, nxf = cfg.par.nobs, cfg.grid.nxfreq
        outs = (capi.SightlineOut * max(nobs, 1))()
        maps = []
        for k in range(nobs):
            ob = cfg.observers[k]
            tg = np.zeros((nxf, ob.nxim, ob.nyim), order="F")
            ng = np.zeros((ob.nxim, ob.nyim), order="F")
            td = np.zeros((ob.nxim, ob.nyim), order="F") if cfg.par.DGR > 0 else None
            outs[k].tau_gas = tg.ctypes.data_as(capi.c_double_p)
            outs[k].N_gas = ng.ctypes.data_as(capi.c_double_p)
            outs[k].tau_dust = td.ctypes.data_as(capi.c_double_p) if td is not None else None
            maps.append(dict(tau_gas=tg, N_gas=ng, tau_dust=td))
        self._check(self._lib.lart_gpu_sightline_tau(self._h, self.model.summary.cross0, outs))
        steps, ms = C.c_double(), C.c_double()
        self._check(self._lib.lart_gpu_sightline_stats(self._h, C.byref(steps), C.byref(ms)))
        self.sightline_stats = dict(cellsteps=steps.value, ms=ms.value)
        return maps

    # ---- unit-level batched plugin points (define.f90:741-784) -------------
    def raytrace_to_edge(self, x, y, z, kx, ky, kz, xfreq, icell, jcell, kcell, trace_cap=0):
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        i = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        x, y, z, kx, ky, kz, xfreq = map(f, (x, y, z, kx, ky, kz, xfreq))
        icell, jcell, kcell = map(i, (icell, jcell, kcell))
        n = x.size
        tau = np.zeros(n)
        nsteps = np.zeros(n, dtype=np.int32)
        trace = np.full((n, trace_cap), -1, dtype=np.int32) if trace_cap > 0 else None
        dp = lambda a: a.ctypes.data_as(capi.c_double_p)
        ip = lambda a: a.ctypes.data_as(capi.c_int32_p)
        self._check(self._lib.lart_gpu_raytrace_edge_batch(
            self._h, n, dp(x), dp(y), dp(z), dp(kx), dp(ky), dp(kz), dp(xfreq), ip(icell), ip(jcell), ip(kcell),
            dp(tau), ip(nsteps), trace_cap, ip(trace) if trace is not None else None))
        return tau, nsteps, trace

    def raytrace_to_tau(self, x, y, z, kx, ky, kz, xfreq, icell, jcell, kcell, tau_in):
        f = lambda a: np.array(a, dtype=np.float64, copy=True)
        i = lambda a: np.array(a, dtype=np.int32, copy=True)
        x, y, z, kx, ky, kz, xfreq, tau_in = map(f, (x, y, z, kx, ky, kz, xfreq, tau_in))
        icell, jcell, kcell = map(i, (icell, jcell, kcell))
        n = x.size
        inside = np.zeros(n, dtype=np.int32)
        nsteps = np.zeros(n, dtype=np.int32)
        xref = np.zeros(n)
        dp = lambda a: a.ctypes.data_as(capi.c_double_p)
        ip = lambda a: a.ctypes.data_as(capi.c_int32_p)
        self._check(self._lib.lart_gpu_raytrace_tau_batch(
            self._h, n, dp(x), dp(y), dp(z), dp(kx), dp(ky), dp(kz), dp(xfreq), ip(icell), ip(jcell), ip(kcell),
            dp(tau_in), ip(inside), dp(xref), ip(nsteps)))
        return dict(x=x, y=y, z=z, xfreq=xfreq, icell=icell, jcell=jcell, kcell=kcell, inside=inside,
                    xfreq_ref=xref, nsteps=nsteps)

    # ---- clump medium, unit level (raytrace_clump.f90)
    def clump_edge(self, x, y, z, kx, ky, kz, xfreq, icl, tau_max=-1.0):
        """raytrace_to_edge_clump (tau_max <= 0) / raytrace_to_edge_clump_capped; returns (tau, clumps crossed)."""
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        x, y, z, kx, ky, kz, xfreq = map(f, (x, y, z, kx, ky, kz, xfreq))
        icl = np.ascontiguousarray(icl, dtype=np.int32)
        n = x.size
        tau, ncl = np.zeros(n), np.zeros(n, dtype=np.int32)
        dp = lambda a: a.ctypes.data_as(capi.c_double_p)
        ip = lambda a: a.ctypes.data_as(capi.c_int32_p)
        self._check(self._lib.lart_gpu_clump_edge_batch(self._h, n, dp(x), dp(y), dp(z), dp(kx), dp(ky), dp(kz), dp(xfreq),
                                                        ip(icl), float(tau_max), dp(tau), ip(ncl)))
        return tau, ncl

    def clump_tau(self, x, y, z, kx, ky, kz, xfreq, icl, tau_in):
        """raytrace_to_tau_clump on copies; returns the updated photon state."""
        f = lambda a: np.array(a, dtype=np.float64, copy=True)
        x, y, z, kx, ky, kz, xfreq, tau_in = map(f, (x, y, z, kx, ky, kz, xfreq, tau_in))
        icl = np.array(icl, dtype=np.int32, copy=True)
        n = x.size
        inside = np.zeros(n, dtype=np.int32)
        dp = lambda a: a.ctypes.data_as(capi.c_double_p)
        ip = lambda a: a.ctypes.data_as(capi.c_int32_p)
        self._check(self._lib.lart_gpu_clump_tau_batch(self._h, n, dp(x), dp(y), dp(z), dp(kx), dp(ky), dp(kz), dp(xfreq),
                                                       ip(icl), dp(tau_in), ip(inside)))
        return dict(x=x, y=y, z=z, xfreq=xfreq, icl=icl, inside=inside)

    def clump_locate(self, x, y, z):
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        x, y, z = map(f, (x, y, z))
        icl = np.zeros(x.size, dtype=np.int32)
        dp = lambda a: a.ctypes.data_as(capi.c_double_p)
        self._check(self._lib.lart_gpu_clump_locate_batch(self._h, x.size, dp(x), dp(y), dp(z),
                                                          icl.ctypes.data_as(capi.c_int32_p)))
        return icl

    def amr_locate(self, x, y, z):
        """amr_find_leaf (octree_mod.f90:149-171): leaf index of every point, 0 = not covered."""
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        x, y, z = map(f, (x, y, z))
        il = np.zeros(x.size, dtype=np.int32)
        dp = lambda a: a.ctypes.data_as(capi.c_double_p)
        self._check(self._lib.lart_gpu_amr_locate_batch(self._h, x.size, dp(x), dp(y), dp(z), il.ctypes.data_as(capi.c_int32_p)))
        return il

    def batch_stats(self):
        """(cells or cell steps walked, kernel ms) of the last sight-line or clump-edge batch call."""
        steps, ms = C.c_double(0), C.c_double(0)
        self._check(self._lib.lart_gpu_sightline_stats(self._h, C.byref(steps), C.byref(ms)))
        return steps.value, ms.value

    def peel_bound(self, x, y, z, xfreq, icell, jcell, kcell):
        """1 where the scatter stage would count a peel ray from (x,y,z) at frequency xfreq instead of walking it."""
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        i = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        x, y, z, xfreq = map(f, (x, y, z, xfreq))
        icell, jcell, kcell = map(i, (icell, jcell, kcell))
        out = np.zeros(x.size, dtype=np.int32)
        dp = lambda a: a.ctypes.data_as(capi.c_double_p)
        ip = lambda a: a.ctypes.data_as(capi.c_int32_p)
        self._check(self._lib.lart_gpu_peel_bound_batch(self._h, x.size, dp(x), dp(y), dp(z), dp(xfreq), ip(icell), ip(jcell),
                                                        ip(kcell), ip(out)))
        return out

    def xcrit_local(self, x, y, z, icell, jcell, kcell):
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        i = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        x, y, z = map(f, (x, y, z))
        icell, jcell, kcell = map(i, (icell, jcell, kcell))
        out = np.zeros(x.size)
        dp = lambda a: a.ctypes.data_as(capi.c_double_p)
        ip = lambda a: a.ctypes.data_as(capi.c_int32_p)
        self._check(self._lib.lart_gpu_xcrit_batch(self._h, x.size, dp(x), dp(y), dp(z), ip(icell), ip(jcell),
                                                   ip(kcell), dp(out)))
        return out


def comm_init(device, nranks, rank, broadcast):
    """One NCCL communicator per process for lart_gpu_reduce (the MPI_INIT of the GPU path).  `broadcast(buf)` must
    hand rank 0's 128-byte `bytearray` to every rank (MPI_BCAST on the Fortran side; a torch.distributed or any other
    host-side broadcast here) and return it."""
    lib = capi.load_gpu()
    buf = (C.c_char * 128)()
    if rank == 0 and lib.lart_gpu_comm_unique_id(buf) != 0:
        raise LartError(lib.lart_gpu_last_error().decode())
    raw = broadcast(bytearray(buf.raw))
    buf = (C.c_char * 128).from_buffer_copy(bytes(raw))
    if lib.lart_gpu_comm_init(device, nranks, rank, buf) != 0:
        raise LartError(lib.lart_gpu_last_error().decode())


def comm_init_torch(device):
    """comm_init with torch.distributed (any backend) as the host-side broadcast of the NCCL id."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()

    def bcast(raw):
        objs = [bytes(raw)]
        dist.broadcast_object_list(objs, src=0)
        return objs[0]

    comm_init(device, world, rank, bcast)


def comm_rank():
    lib = capi.load_gpu()
    r = C.c_int32()
    lib.lart_gpu_comm_info(None, C.byref(r))
    return r.value


def comm_finalize():
    capi.load_gpu().lart_gpu_comm_finalize()


def measure_fp64(device=0):
    """FP64 FMA throughput of the device in TFLOP/s (register-resident DFMA loop)."""
    lib = capi.load_gpu()
    t = C.c_double()
    if lib.lart_gpu_measure_fp64(device, C.byref(t)) != 0:
        raise LartError(lib.lart_gpu_last_error().decode())
    return t.value


def calc_voigt(x, a):
    """voigt_seon2 on the GPU (voigt_mod.f90:541-733)."""
    lib = capi.load_gpu()
    x = np.ascontiguousarray(x, dtype=np.float64)
    a = np.ascontiguousarray(np.broadcast_to(a, x.shape), dtype=np.float64)
    H = np.empty_like(x)
    rc = lib.lart_gpu_voigt_batch(x.size, x.ctypes.data_as(capi.c_double_p), a.ctypes.data_as(capi.c_double_p),
                                  H.ctypes.data_as(capi.c_double_p))
    if rc != 0:
        raise LartError(lib.lart_gpu_last_error().decode())
    return H


def sample(kind, seed, ids, p0=None, p1=None, ndraw=1):
    """Random variates on the path, one Philox stream per element (see lart_gpu_sample_batch)."""
    lib = capi.load_gpu()
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    n = ids.size
    p0 = np.ascontiguousarray(np.broadcast_to(0.0 if p0 is None else p0, (n,)), dtype=np.float64)
    p1 = np.ascontiguousarray(np.broadcast_to(0.0 if p1 is None else p1, (n,)), dtype=np.float64)
    out = np.empty((n, ndraw))
    rc = lib.lart_gpu_sample_batch(kind, seed, n, ids.ctypes.data_as(capi.c_int64_p),
                                   p0.ctypes.data_as(capi.c_double_p), p1.ctypes.data_as(capi.c_double_p), ndraw,
                                   out.ctypes.data_as(capi.c_double_p))
    if rc != 0:
        raise LartError(lib.lart_gpu_last_error().decode())
    return out
