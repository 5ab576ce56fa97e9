"""ctypes mirror of include/lart_gpu.h and include/lart_host.h.

Only declarations live here: POD structure layouts and function prototypes of the
two shared libraries the package ships

  lart_b200/liblart_gpu.so   CUDA engine behind the C ABI (include/lart_gpu.h)
  lart_b200/liblart_host.so  C++ mini-host (include/lart_host.h)

Loading the engine fails loudly when the CUDA library is missing — there is no
CPU fallback on the product path (the CPU checker used by the tests is test
infrastructure and is never imported from this package).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)

LART_MAX_OBSERVERS = 181
SPEC_MONO, SPEC_VOIGT, SPEC_VOIGT0, SPEC_CONTINUUM, SPEC_GAUSSIAN = range(5)
SRC_POINT, SRC_UNIFORM, SRC_UNIFORM_SPHERE, SRC_PLANE_ILLUMINATION = range(4)
ATM_NONE, ATM_PLANE, ATM_SPHERICAL = range(3)
FLAG_SOA_GRID = 1
FLAG_NO_WARP_AGG = 2
FLAG_MONOLITHIC = 4
FLAG_STAGE_TIMING = 8
FLAG_SERIAL_REJECTION = 16
FLAG_LOCAL_STEPS = 32
FLAG_DEBUG_TINY_QUEUES = 64
FLAG_SPECULATIVE_REJECTION = 128
STAGES = ["emit", "trace", "draw", "apply", "peel"]


class Grid(C.Structure):
    _fields_ = [
        ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("nxfreq", C.c_int32),
        ("xmin", C.c_double), ("ymin", C.c_double), ("zmin", C.c_double),
        ("xmax", C.c_double), ("ymax", C.c_double), ("zmax", C.c_double),
        ("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double),
        ("Dfreq_ref", C.c_double), ("xfreq_min", C.c_double), ("xfreq_max", C.c_double),
        ("dxfreq", C.c_double), ("xcrit", C.c_double), ("xcrit2", C.c_double), ("rmax", C.c_double),
        ("i0", C.c_int32), ("j0", C.c_int32), ("k0", C.c_int32), ("pad_", C.c_int32),
        ("xface", c_double_p), ("yface", c_double_p), ("zface", c_double_p),
        ("rhokap", c_double_p), ("voigt_a", c_double_p), ("Dfreq", c_double_p),
        ("vfx", c_double_p), ("vfy", c_double_p), ("vfz", c_double_p), ("rhokapD", c_double_p),
        ("mask", C.POINTER(C.c_int8)), ("geometry_JPa", C.c_int32), ("nr", C.c_int32),
        ("ind_sph", c_int32_p), ("ind_cyl", c_int32_p),
    ]


class Params(C.Structure):
    _fields_ = [
        ("nphotons", C.c_int64), ("seed", C.c_uint64), ("xfreq0", C.c_double),
        ("xs_point", C.c_double), ("ys_point", C.c_double), ("zs_point", C.c_double),
        ("source_rmax", C.c_double), ("DGR", C.c_double), ("albedo", C.c_double), ("hgg", C.c_double),
        ("voigt_a0", C.c_double), ("Dfreq0", C.c_double), ("gaussian_sigma_x", C.c_double),
        ("mu_min", C.c_double), ("dmu", C.c_double), ("nmu", C.c_int32),
        ("spectral_type", C.c_int32), ("source_geometry", C.c_int32), ("comoving_source", C.c_int32),
        ("recoil", C.c_int32), ("core_skip", C.c_int32), ("core_skip_global", C.c_int32),
        ("use_stokes", C.c_int32), ("use_reduced_wgt", C.c_int32),
        ("save_Jin", C.c_int32), ("save_Jabs", C.c_int32), ("save_Jmu", C.c_int32),
        ("save_peeloff", C.c_int32), ("save_peeloff_2D", C.c_int32), ("save_peeloff_3D", C.c_int32),
        ("save_direc0", C.c_int32), ("save_all_photons", C.c_int32), ("xyz_symmetry", C.c_int32), ("xy_symmetry", C.c_int32),
        ("use_clump_medium", C.c_int32), ("xy_periodic", C.c_int32),
        ("nobs", C.c_int32), ("use_amr_grid", C.c_int32),
        ("atmosphere", C.c_int32), ("calc_J", C.c_int32), ("calc_P", C.c_int32), ("calc_Pnew", C.c_int32),
        ("Omega", C.c_double),
    ]


class Line(C.Structure):
    _fields_ = [("line_type", C.c_int32), ("pad_", C.c_int32), ("E1", C.c_double), ("E2", C.c_double),
                ("E3", C.c_double), ("g_recoil0", C.c_double), ("DnuHK_Hz", C.c_double), ("cross0", C.c_double)]


class Observer(C.Structure):
    _fields_ = [("x", C.c_double), ("y", C.c_double), ("z", C.c_double), ("rmatrix", C.c_double * 9),
                ("dxim", C.c_double), ("dyim", C.c_double), ("nxim", C.c_int32), ("nyim", C.c_int32)]


class ScattMat(C.Structure):
    _fields_ = [("nPDF", C.c_int32), ("pad_", C.c_int32), ("coss", c_double_p), ("S11", c_double_p),
                ("S12", c_double_p), ("S33", c_double_p), ("S34", c_double_p), ("phase_PDF", c_double_p),
                ("alias", c_int32_p)]


class Clumps(C.Structure):
    _fields_ = [("n", C.c_int64), ("sphere_R", C.c_double), ("Dfreq_ref", C.c_double),
                ("x", c_double_p), ("y", c_double_p), ("z", c_double_p),
                ("vx", c_double_p), ("vy", c_double_p), ("vz", c_double_p),
                ("radius", c_double_p), ("rhokap", c_double_p), ("rhokapD", c_double_p),
                ("voigt_a", c_double_p), ("Dfreq", c_double_p),
                ("cgx", C.c_int32), ("cgy", C.c_int32), ("cgz", C.c_int32), ("has_overlap", C.c_int32),
                ("cg_xmin", C.c_double), ("cg_ymin", C.c_double), ("cg_zmin", C.c_double),
                ("cg_dx", C.c_double), ("cg_dy", C.c_double), ("cg_dz", C.c_double),
                ("cg_start", c_int32_p), ("cg_list", c_int32_p)]


class Amr(C.Structure):
    _fields_ = [("ncells", C.c_int32), ("nleaf", C.c_int32), ("children", c_int32_p), ("ileaf", c_int32_p),
                ("icell_of_leaf", c_int32_p), ("neighbor", c_int32_p),
                ("cx", c_double_p), ("cy", c_double_p), ("cz", c_double_p), ("ch", c_double_p),
                ("rhokap", c_double_p), ("voigt_a", c_double_p), ("Dfreq", c_double_p),
                ("vfx", c_double_p), ("vfy", c_double_p), ("vfz", c_double_p), ("rhokapD", c_double_p)]


class Config(C.Structure):
    _fields_ = [("grid", Grid), ("par", Params), ("line", Line), ("scatt_mat", ScattMat), ("clumps", Clumps), ("amr", Amr),
                ("observers", C.POINTER(Observer)), ("device", C.c_int32), ("pool_slots", C.c_int32),
                ("quantum", C.c_int32), ("flags", C.c_int32), ("streams", C.c_int32), ("ray_budget", C.c_int32),
                ("max_events", C.c_int32), ("pad_", C.c_int32)]


OBS_FIELDS = ["scatt", "direc", "direc0", "I", "Q", "U", "V",
              "scatt_2D", "direc_2D", "direc0_2D", "I_2D", "Q_2D", "U_2D", "V_2D"]
ALLPH_FIELDS = ["rp0", "rp", "xfreq1", "xfreq2", "nscatt_gas", "nscatt_dust", "I", "Q", "U", "V"]
COUNTER_FIELDS = ["n_photons_done", "n_scatter", "n_cellsteps", "n_peel", "n_rng", "n_reject_iter", "n_peel_bound",
                  "n_cellsteps_bound"]


class ObserverOut(C.Structure):
    _fields_ = [(n, c_double_p) for n in OBS_FIELDS]


class AllphOut(C.Structure):
    _fields_ = [(n, c_double_p) for n in ALLPH_FIELDS]


class Counters(C.Structure):
    _fields_ = [(n, C.c_double) for n in COUNTER_FIELDS]


class Tallies(C.Structure):
    _fields_ = [("Jout", c_double_p), ("Jin", c_double_p), ("Jabs", c_double_p), ("Jmu", c_double_p),
                ("obs", C.POINTER(ObserverOut)), ("allph", AllphOut),
                ("nscatt_gas", C.c_double), ("nscatt_dust", C.c_double), ("counters", Counters),
                ("Jabs2", c_double_p), ("J", c_double_p), ("Pa", c_double_p), ("Pnew", c_double_p)]


class SightlineOut(C.Structure):
    _fields_ = [("tau_gas", c_double_p), ("N_gas", c_double_p), ("tau_dust", c_double_p)]


class HostSummary(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ["voigt_a", "temperature", "N_gaspole", "N_gashomo", "taupole", "tauhomo", "taupole_dust",
                 "tauhomo_dust", "Dfreq_ref", "vtherm", "cross0", "atau3", "xfreq_min", "xfreq_max", "dxfreq",
                 "dxim", "dyim", "distance"]] + \
               [(n, C.c_int32) for n in ["nx", "ny", "nz", "nxfreq", "nobs", "nxim", "nyim", "zonly"]] + \
               [("nphotons", C.c_int64), ("nclumps", C.c_int64)]


# every symbol include/lart_gpu.h declares (tests check the library exports all of them)
GPU_SYMBOLS = [
    "lart_gpu_create", "lart_gpu_run", "lart_gpu_begin", "lart_gpu_step", "lart_gpu_sync", "lart_gpu_fetch",
    "lart_gpu_reset_tallies", "lart_gpu_destroy", "lart_gpu_tally_buffer", "lart_gpu_allph_buffer",
    "lart_gpu_stream", "lart_gpu_kernel_ms", "lart_gpu_last_error", "lart_gpu_voigt_batch",
    "lart_gpu_raytrace_edge_batch", "lart_gpu_raytrace_tau_batch", "lart_gpu_sample_batch",
    "lart_gpu_xcrit_batch", "lart_gpu_version", "lart_gpu_stage_ms", "lart_gpu_pool_slots", "lart_gpu_measure_fp64",
    "lart_gpu_sightline_tau", "lart_gpu_sightline_stats",
    "lart_gpu_clump_edge_batch", "lart_gpu_clump_tau_batch", "lart_gpu_clump_locate_batch", "lart_gpu_peel_bound_batch", "lart_gpu_amr_locate_batch",
    "lart_gpu_comm_unique_id", "lart_gpu_comm_init", "lart_gpu_comm_info", "lart_gpu_comm_finalize", "lart_gpu_reduce",
    "lart_gpu_deal_open", "lart_gpu_deal_close", "lart_gpu_run_dealt", "lart_gpu_deal_claim",
]
HOST_SYMBOLS = [
    "lart_host_new", "lart_host_free", "lart_host_set", "lart_host_read_input", "lart_host_setup",
    "lart_host_config", "lart_host_get_summary", "lart_host_tallies", "lart_host_zero_tallies",
    "lart_host_normalize", "lart_host_last_error", "lart_host_set_amr_leaves",
]

_host = None
_gpu = None


def host_lib_path():
    return os.path.join(_HERE, "liblart_host.so")


def gpu_lib_path():
    # LART_GPU_LIB: an alternative build of the same engine (A/B experiments); never a fallback
    return os.environ.get("LART_GPU_LIB") or os.path.join(_HERE, "liblart_gpu.so")


def load_host():
    global _host
    if _host is None:
        lib = C.CDLL(host_lib_path())
        lib.lart_host_new.restype = C.c_void_p
        lib.lart_host_free.argtypes = [C.c_void_p]
        lib.lart_host_set.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        lib.lart_host_read_input.argtypes = [C.c_void_p, C.c_char_p]
        lib.lart_host_setup.argtypes = [C.c_void_p]
        lib.lart_host_config.argtypes = [C.c_void_p]
        lib.lart_host_config.restype = C.POINTER(Config)
        lib.lart_host_get_summary.argtypes = [C.c_void_p, C.POINTER(HostSummary)]
        lib.lart_host_tallies.argtypes = [C.c_void_p]
        lib.lart_host_tallies.restype = C.POINTER(Tallies)
        lib.lart_host_zero_tallies.argtypes = [C.c_void_p]
        lib.lart_host_normalize.argtypes = [C.c_void_p]
        lib.lart_host_last_error.restype = C.c_char_p
        lib.lart_host_set_amr_leaves.argtypes = [C.c_void_p, C.c_int64] + [c_double_p] * 3 + [c_int32_p] + [c_double_p] * 5 + \
                                               [C.c_double] * 4
        _host = lib
    return _host


def load_gpu():
    """Load the CUDA engine.  Raises (never falls back) if the library is absent."""
    global _gpu
    if _gpu is None:
        path = gpu_lib_path()
        if not os.path.exists(path):
            raise RuntimeError(
                "lart_b200: CUDA engine %s not built — run `python -c 'import __graft_entry__ as g; g.build()'`; "
                "there is no CPU fallback" % path)
        lib = C.CDLL(path)
        H = C.c_void_p
        lib.lart_gpu_create.argtypes = [C.POINTER(Config), C.POINTER(H)]
        lib.lart_gpu_run.argtypes = [H, C.c_int64, C.c_int64, C.c_int64]
        lib.lart_gpu_begin.argtypes = [H, C.c_int64, C.c_int64, C.c_int64]
        lib.lart_gpu_step.argtypes = [H, C.c_int32, c_int64_p]
        lib.lart_gpu_sync.argtypes = [H]
        lib.lart_gpu_fetch.argtypes = [H, C.POINTER(Tallies)]
        lib.lart_gpu_reset_tallies.argtypes = [H]
        lib.lart_gpu_destroy.argtypes = [H]
        lib.lart_gpu_tally_buffer.argtypes = [H, C.POINTER(C.c_void_p), c_int64_p]
        lib.lart_gpu_allph_buffer.argtypes = [H, C.POINTER(C.c_void_p), c_int64_p]
        lib.lart_gpu_stream.argtypes = [H, C.POINTER(C.c_void_p)]
        lib.lart_gpu_kernel_ms.argtypes = [H, c_double_p, c_int64_p]
        lib.lart_gpu_last_error.restype = C.c_char_p
        lib.lart_gpu_voigt_batch.argtypes = [C.c_int64, c_double_p, c_double_p, c_double_p]
        lib.lart_gpu_raytrace_edge_batch.argtypes = [H, C.c_int64] + [c_double_p] * 7 + [c_int32_p] * 3 + \
            [c_double_p, c_int32_p, C.c_int32, c_int32_p]
        lib.lart_gpu_raytrace_tau_batch.argtypes = [H, C.c_int64] + [c_double_p] * 7 + [c_int32_p] * 3 + \
            [c_double_p, c_int32_p, c_double_p, c_int32_p]
        lib.lart_gpu_sample_batch.argtypes = [C.c_int32, C.c_uint64, C.c_int64, c_int64_p, c_double_p, c_double_p,
                                              C.c_int32, c_double_p]
        lib.lart_gpu_xcrit_batch.argtypes = [H, C.c_int64] + [c_double_p] * 3 + [c_int32_p] * 3 + [c_double_p]
        lib.lart_gpu_version.restype = C.c_int
        lib.lart_gpu_stage_ms.argtypes = [H, c_double_p, c_int64_p]
        lib.lart_gpu_pool_slots.argtypes = [H, c_int64_p]
        lib.lart_gpu_measure_fp64.argtypes = [C.c_int32, c_double_p]
        lib.lart_gpu_sightline_tau.argtypes = [H, C.c_double, C.POINTER(SightlineOut)]
        lib.lart_gpu_sightline_stats.argtypes = [H, c_double_p, c_double_p]
        lib.lart_gpu_clump_edge_batch.argtypes = [H, C.c_int64] + [c_double_p] * 7 + [c_int32_p, C.c_double, c_double_p, c_int32_p]
        lib.lart_gpu_clump_tau_batch.argtypes = [H, C.c_int64] + [c_double_p] * 7 + [c_int32_p, c_double_p, c_int32_p]
        lib.lart_gpu_clump_locate_batch.argtypes = [H, C.c_int64] + [c_double_p] * 3 + [c_int32_p]
        if hasattr(lib, "lart_gpu_reduce"):
            lib.lart_gpu_comm_unique_id.argtypes = [C.c_void_p]
            lib.lart_gpu_comm_init.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
            lib.lart_gpu_comm_info.argtypes = [c_int32_p, c_int32_p]
            lib.lart_gpu_reduce.argtypes = [H, C.c_int32]
        if hasattr(lib, "lart_gpu_run_dealt"):
            lib.lart_gpu_deal_open.argtypes = [C.c_char_p, C.c_int32, C.POINTER(C.c_void_p)]
            lib.lart_gpu_deal_close.argtypes = [C.c_void_p, C.c_int32]
            lib.lart_gpu_run_dealt.argtypes = [H, C.c_void_p, C.c_int64, C.c_int64, c_int64_p]
            lib.lart_gpu_deal_claim.argtypes = [C.c_void_p, C.c_int64, C.c_int64, c_int64_p, c_int64_p]
        if hasattr(lib, "lart_gpu_amr_locate_batch"):
            lib.lart_gpu_amr_locate_batch.argtypes = [H, C.c_int64] + [c_double_p] * 3 + [c_int32_p]
        if hasattr(lib, "lart_gpu_peel_bound_batch"):  # (absent from older A/B builds selected through LART_GPU_LIB)
            lib.lart_gpu_peel_bound_batch.argtypes = [H, C.c_int64] + [c_double_p] * 4 + [c_int32_p] * 4
        _gpu = lib
    return _gpu
