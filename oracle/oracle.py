"""ctypes wrapper of the CPU oracle (oracle/liblart_oracle.so).

TEST INFRASTRUCTURE ONLY.  Import this from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs — never from lart_b200/.
It consumes the same lart_config the product consumes (structure layouts are
shared declarations from lart_b200/capi.py) so both sides see identical inputs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from lart_b200 import capi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_NATIVE = False
PORTABLE_FLAGS = "g++ -O3 -march=x86-64-v3 -ffp-contract=off"
NATIVE_FLAGS = "g++ -O3 -march=native"


def lib_path():
    if _NATIVE:
        return os.path.join(_HERE, "_native", _cpu_tag(), "liblart_oracle_native.so")
    return os.path.join(_HERE, "liblart_oracle.so")


def _cpu_tag():
    """A -march=native build is only valid on the CPU it was made on: key it by the CPU's model and flag list."""
    import hashlib
    try:
        lines = [l for l in open("/proc/cpuinfo") if l.startswith(("model name", "flags"))][:2]
    except OSError:
        lines = []
    return hashlib.sha1("".join(lines).encode()).hexdigest()[:12]


def use_native():
    """Timing legs only (bench.py): switch to a build made ON this machine with -march=native and FMA contraction
    allowed.  Returns the compiler flags in use; falls back to the portable build (the
    parity build: -march=x86-64-v3 -ffp-contract=off) when the compile fails.  Call before the first load()."""
    global _NATIVE, _LIB
    try:
        subprocess.check_call(["make", "-C", _HERE, "-s", "native", "NATIVEDIR=_native/" + _cpu_tag()],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        _NATIVE, _LIB = True, None
        load()
        return NATIVE_FLAGS
    except Exception:
        _NATIVE, _LIB = False, None
        return PORTABLE_FLAGS


def build(force=False):
    src = os.path.join(_HERE, "lart_oracle.cpp")
    if force or not os.path.exists(lib_path()) or os.path.getmtime(lib_path()) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B" if force else "-s"])


def load():
    global _LIB
    if _LIB is None:
        if not os.path.exists(lib_path()):
            build()
        lib = C.CDLL(lib_path())
        dp, ip, lp = capi.c_double_p, capi.c_int32_p, capi.c_int64_p
        cfgp = C.POINTER(capi.Config)
        lib.oracle_last_error.restype = C.c_char_p
        lib.oracle_voigt.argtypes = [C.c_int64, dp, dp, dp]
        lib.oracle_raytrace_edge.argtypes = [cfgp, C.c_int64] + [dp] * 7 + [ip] * 3 + [dp, ip, C.c_int32, ip]
        lib.oracle_raytrace_tau.argtypes = [cfgp, C.c_int64] + [dp] * 7 + [ip] * 3 + [dp, ip, dp, ip]
        lib.oracle_xcrit.argtypes = [cfgp, C.c_int64] + [dp] * 3 + [ip] * 3 + [dp]
        lib.oracle_sample.argtypes = [C.c_int32, C.c_int32, C.c_uint64, C.c_int64, lp, dp, dp, C.c_int32, dp]
        lib.oracle_philox_block.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint32)]
        lib.oracle_philox_raw.argtypes = [C.POINTER(C.c_uint32)] * 3
        lib.oracle_mt64_words.argtypes = [C.c_int64, C.c_int64, C.POINTER(C.c_uint64)]
        lib.oracle_run.argtypes = [cfgp, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                   C.POINTER(capi.Tallies)]
        lib.oracle_hardware_threads.restype = C.c_int
        lib.oracle_sightline_tau.argtypes = [cfgp, C.c_double, C.POINTER(capi.SightlineOut), C.POINTER(C.c_double)]
        lib.oracle_clump_edge.argtypes = [cfgp, C.c_int64] + [dp] * 7 + [ip, C.c_double, dp, ip]
        lib.oracle_clump_tau.argtypes = [cfgp, C.c_int64] + [dp] * 7 + [ip, dp, ip]
        lib.oracle_clump_locate.argtypes = [cfgp, C.c_int64] + [dp] * 3 + [ip]
        lib.oracle_amr_edge.argtypes = [cfgp, C.c_int64] + [dp] * 7 + [ip, dp, ip]
        lib.oracle_amr_tau.argtypes = [cfgp, C.c_int64] + [dp] * 7 + [ip, dp, ip, dp, ip]
        lib.oracle_amr_locate.argtypes = [cfgp, C.c_int64] + [dp] * 3 + [ip]
        _LIB = lib
    return _LIB


def _check(rc):
    if rc != 0:
        raise RuntimeError(load().oracle_last_error().decode())


def _d(a):
    return a.ctypes.data_as(capi.c_double_p)


def _i(a):
    return a.ctypes.data_as(capi.c_int32_p)


def voigt(x, a):
    x = np.ascontiguousarray(x, dtype=np.float64)
    a = np.ascontiguousarray(np.broadcast_to(a, x.shape), dtype=np.float64)
    H = np.empty_like(x)
    _check(load().oracle_voigt(x.size, _d(x), _d(a), _d(H)))
    return H


def raytrace_to_edge(cfg, x, y, z, kx, ky, kz, xfreq, ic, jc, kc, trace_cap=0):
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    g = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    x, y, z, kx, ky, kz, xfreq = map(f, (x, y, z, kx, ky, kz, xfreq))
    ic, jc, kc = map(g, (ic, jc, kc))
    n = x.size
    tau = np.zeros(n)
    ns = np.zeros(n, dtype=np.int32)
    trace = np.full((n, trace_cap), -1, dtype=np.int32) if trace_cap > 0 else None
    _check(load().oracle_raytrace_edge(cfg, n, _d(x), _d(y), _d(z), _d(kx), _d(ky), _d(kz), _d(xfreq), _i(ic), _i(jc),
                                       _i(kc), _d(tau), _i(ns), trace_cap, _i(trace) if trace is not None else None))
    return tau, ns, trace


def raytrace_to_tau(cfg, x, y, z, kx, ky, kz, xfreq, ic, jc, kc, tau_in):
    f = lambda a: np.array(a, dtype=np.float64, copy=True)
    g = lambda a: np.array(a, dtype=np.int32, copy=True)
    x, y, z, kx, ky, kz, xfreq, tau_in = map(f, (x, y, z, kx, ky, kz, xfreq, tau_in))
    ic, jc, kc = map(g, (ic, jc, kc))
    n = x.size
    inside = np.zeros(n, dtype=np.int32)
    ns = np.zeros(n, dtype=np.int32)
    xref = np.zeros(n)
    _check(load().oracle_raytrace_tau(cfg, n, _d(x), _d(y), _d(z), _d(kx), _d(ky), _d(kz), _d(xfreq), _i(ic), _i(jc),
                                      _i(kc), _d(tau_in), _i(inside), _d(xref), _i(ns)))
    return dict(x=x, y=y, z=z, xfreq=xfreq, icell=ic, jcell=jc, kcell=kc, inside=inside, xfreq_ref=xref, nsteps=ns)


def clump_edge(cfg, x, y, z, kx, ky, kz, xfreq, icl, tau_max=-1.0):
    """raytrace_to_edge_clump (tau_max <= 0) / raytrace_to_edge_clump_capped."""
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    x, y, z, kx, ky, kz, xfreq = map(f, (x, y, z, kx, ky, kz, xfreq))
    icl = np.ascontiguousarray(icl, dtype=np.int32)
    n = x.size
    tau, nc = np.zeros(n), np.zeros(n, dtype=np.int32)
    _check(load().oracle_clump_edge(cfg, n, _d(x), _d(y), _d(z), _d(kx), _d(ky), _d(kz), _d(xfreq), _i(icl), float(tau_max),
                                    _d(tau), _i(nc)))
    return tau, nc


def clump_tau(cfg, x, y, z, kx, ky, kz, xfreq, icl, tau_in):
    """raytrace_to_tau_clump on copies; returns the updated photon state."""
    f = lambda a: np.array(a, dtype=np.float64, copy=True)
    x, y, z, kx, ky, kz, xfreq, tau_in = map(f, (x, y, z, kx, ky, kz, xfreq, tau_in))
    icl = np.array(icl, dtype=np.int32, copy=True)
    n = x.size
    inside = np.zeros(n, dtype=np.int32)
    _check(load().oracle_clump_tau(cfg, n, _d(x), _d(y), _d(z), _d(kx), _d(ky), _d(kz), _d(xfreq), _i(icl), _d(tau_in),
                                   _i(inside)))
    return dict(x=x, y=y, z=z, xfreq=xfreq, icl=icl, inside=inside)


def clump_locate(cfg, x, y, z):
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    x, y, z = map(f, (x, y, z))
    icl = np.zeros(x.size, dtype=np.int32)
    _check(load().oracle_clump_locate(cfg, x.size, _d(x), _d(y), _d(z), _i(icl)))
    return icl


def xcrit_local(cfg, x, y, z, ic, jc, kc):
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    g = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    x, y, z = map(f, (x, y, z))
    ic, jc, kc = map(g, (ic, jc, kc))
    out = np.zeros(x.size)
    _check(load().oracle_xcrit(cfg, x.size, _d(x), _d(y), _d(z), _i(ic), _i(jc), _i(kc), _d(out)))
    return out


def sample(kind, seed, ids, p0=None, p1=None, ndraw=1, rng_mode=1):
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    n = ids.size
    p0 = np.ascontiguousarray(np.broadcast_to(0.0 if p0 is None else p0, (n,)), dtype=np.float64)
    p1 = np.ascontiguousarray(np.broadcast_to(0.0 if p1 is None else p1, (n,)), dtype=np.float64)
    out = np.empty((n, ndraw))
    _check(load().oracle_sample(kind, rng_mode, seed, n, ids.ctypes.data_as(capi.c_int64_p), _d(p0), _d(p1), ndraw,
                                _d(out)))
    return out


def philox_block(seed, stream, block):
    out = (C.c_uint32 * 4)()
    load().oracle_philox_block(seed, stream, block, out)
    return list(out)


def philox_raw(ctr, key):
    out = (C.c_uint32 * 4)()
    load().oracle_philox_raw((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
    return list(out)


def mt64_words(seed, n):
    out = (C.c_uint64 * n)()
    load().oracle_mt64_words(seed, n, out)
    return list(out)


def hardware_threads():
    return load().oracle_hardware_threads()


def run(model, rng_mode=1, nthreads=None, first_id=1, count=None, stride=1, max_events=0, seed=None):
    """Whole run into the model's host tallies (ADDED, like the product's fetch)."""
    cfg = model.config
    if seed is not None:
        cfg.contents.par.seed = seed
    if count is None:
        count = cfg.contents.par.nphotons
    if nthreads is None:
        nthreads = hardware_threads()
    _check(load().oracle_run(cfg, rng_mode, nthreads, first_id, count, stride, max_events, model.tallies))


def sightline_tau(model):
    """make_sightline_tau_outside on the CPU: list of dict(tau_gas, N_gas, tau_dust) per observer."""
    cfg = model.config.contents
    nobs, nxf = cfg.par.nobs, cfg.grid.nxfreq
    outs = (capi.SightlineOut * max(nobs, 1))()
    maps = []
    for k in range(nobs):
        ob = cfg.observers[k]
        tg = np.zeros((nxf, ob.nxim, ob.nyim), order="F")
        ng = np.zeros((ob.nxim, ob.nyim), order="F")
        td = np.zeros((ob.nxim, ob.nyim), order="F") if cfg.par.DGR > 0 else None
        outs[k].tau_gas = _d(tg)
        outs[k].N_gas = _d(ng)
        outs[k].tau_dust = _d(td) if td is not None else None
        maps.append(dict(tau_gas=tg, N_gas=ng, tau_dust=td))
    steps = C.c_double()
    _check(load().oracle_sightline_tau(model.config, model.summary.cross0, outs, C.byref(steps)))
    return maps, steps.value


# ---- octree AMR, unit level (raytrace_amr.f90 / octree_mod.f90) ----------------------------------------------------
def amr_edge(cfg, x, y, z, kx, ky, kz, xfreq, il):
    """raytrace_to_edge_amr; returns (tau, cells crossed)."""
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    x, y, z, kx, ky, kz, xfreq = map(f, (x, y, z, kx, ky, kz, xfreq))
    il = np.ascontiguousarray(il, dtype=np.int32)
    tau, ns = np.zeros(x.size), np.zeros(x.size, dtype=np.int32)
    _check(load().oracle_amr_edge(cfg, x.size, _d(x), _d(y), _d(z), _d(kx), _d(ky), _d(kz), _d(xfreq), _i(il), _d(tau), _i(ns)))
    return tau, ns


def amr_tau(cfg, x, y, z, kx, ky, kz, xfreq, il, tau_in):
    """raytrace_to_tau_amr on copies; returns the updated photon state."""
    f = lambda a: np.array(a, dtype=np.float64, copy=True)
    x, y, z, kx, ky, kz, xfreq, tau_in = map(f, (x, y, z, kx, ky, kz, xfreq, tau_in))
    il = np.array(il, dtype=np.int32, copy=True)
    n = x.size
    inside, ns, xref = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32), np.zeros(n)
    _check(load().oracle_amr_tau(cfg, n, _d(x), _d(y), _d(z), _d(kx), _d(ky), _d(kz), _d(xfreq), _i(il), _d(tau_in), _i(inside),
                                 _d(xref), _i(ns)))
    return dict(x=x, y=y, z=z, xfreq=xfreq, il=il, inside=inside, xfreq_ref=xref, nsteps=ns)


def amr_locate(cfg, x, y, z):
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    x, y, z = map(f, (x, y, z))
    il = np.zeros(x.size, dtype=np.int32)
    _check(load().oracle_amr_locate(cfg, x.size, _d(x), _d(y), _d(z), _i(il)))
    return il
