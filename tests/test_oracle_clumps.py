"""Clump medium: the mini-host's population + CSR grid and the oracle's clump ray tracers (CPU only).
Reference: clump_mod.f90:646-1349, 1369-1540, 2316-2380; raytrace_clump.f90:83-270, 494-533."""
import numpy as np
import pytest

from conftest import golden
from lart_b200 import Model
from oracle import oracle


def clump_model(**kw):
    par = dict(no_photons=1000, use_clump_medium=True, rmax=1.0, clump_radius=0.05, clump_f_cov=2.0, clump_tau0=3.0,
               temperature=1e4, nxfreq=81, nx=11, ny=11, nz=11, iseed=3)
    par.update(kw)
    return Model(**par).setup()


def clump_arrays(m):
    c = m.config.contents.clumps
    n = c.n
    g = lambda p, k=n, t=np.float64: np.ctypeslib.as_array(p, shape=(k,)).astype(t)
    ncell = c.cgx * c.cgy * c.cgz
    start = np.ctypeslib.as_array(c.cg_start, shape=(ncell + 1,))
    lst = np.ctypeslib.as_array(c.cg_list, shape=(int(start[-1]) - 1,))
    return dict(x=g(c.x), y=g(c.y), z=g(c.z), r=g(c.radius), kap=g(c.rhokap), a=g(c.voigt_a), D=g(c.Dfreq),
                vx=g(c.vx), vy=g(c.vy), vz=g(c.vz), start=start, list=lst, c=c)


def rays_from(rng, n, pos=None, R=1.0):
    k = rng.normal(size=(n, 3))
    k /= np.linalg.norm(k, axis=1)[:, None]
    if pos is None:
        p = rng.normal(size=(n, 3))
        p *= (R * rng.uniform(0, 1, n) ** (1 / 3) / np.linalg.norm(p, axis=1))[:, None]
    else:
        p = np.tile(np.asarray(pos, float), (n, 1))
    return p, k


def test_population_and_csr_grid():
    from scipy.spatial import cKDTree
    m = clump_model(rmin=0.2, clump_fully_inside=True)
    A = clump_arrays(m)
    c = A["c"]
    n = c.n
    assert n == round(4 / 3 * 2.0 * (1 + 0.2 + 0.04) / 0.05 ** 2) == m.summary.nclumps  # clump_mod.f90:724-727
    P = np.c_[A["x"], A["y"], A["z"]]
    rr = np.linalg.norm(P, axis=1)
    assert rr.min() >= 0.2 + 0.05 and rr.max() <= 1.0 - 0.05  # fully inside the shell
    d, _ = cKDTree(P).query(P, k=2)
    assert d[:, 1].min() >= 2 * 0.05  # random sequential addition: no overlap
    # CSR: a clump is registered in every cell its bounding box touches (clump_cell_range :1352-1366)
    assert c.cgx == c.cgy == c.cgz == 32 and A["start"][0] == 1 and c.cg_xmin == -(1.0 + 0.05)
    reg = {}
    for cell in range(c.cgx ** 3):
        for icl in A["list"][A["start"][cell] - 1:A["start"][cell + 1] - 1]:
            reg.setdefault(int(icl), []).append(cell)
    assert sorted(reg) == list(range(1, n + 1))
    inv = 1.0 / c.cg_dx
    for icl in (1, n // 2, n):
        lo = np.maximum(0, ((P[icl - 1] - c.cg_xmin - 0.05) * inv).astype(int))
        hi = np.minimum(c.cgx - 1, ((P[icl - 1] - c.cg_xmin + 0.05) * inv).astype(int))
        want = sorted(i + c.cgx * (j + c.cgx * k) for k in range(lo[2], hi[2] + 1) for j in range(lo[1], hi[1] + 1)
                      for i in range(lo[0], hi[0] + 1))
        assert sorted(reg[icl]) == want
    # system scalars: uniform, isotropic population -> tauhomo = (4/3) f_cov tau_clump (compute_clump_scalars :2316-2380)
    kap0 = 3.0 / 0.05
    assert A["kap"][0] * (1 - 1.1283791671 * A["a"][0] + A["a"][0] ** 2) == pytest.approx(kap0, rel=1e-12)
    assert m.summary.tauhomo == pytest.approx(n * 0.05 ** 3 * kap0 / 1.24, rel=1e-9)
    g = m.config.contents.grid
    assert m.grid_array("rhokap").max() == 0.0 and g.Dfreq_ref == c.Dfreq_ref and m.grid_array("vfx").max() == 0.0


def test_edge_walk_equals_brute_force_over_all_clumps():
    m = clump_model(velocity_type="hubble", Vexp=80.0, clump_sigma_v=15.0)
    A = clump_arrays(m)
    rng = np.random.default_rng(5)
    n = 3000
    p, k = rays_from(rng, n)
    xf = rng.normal(size=n) * 2
    icl = oracle.clump_locate(m.config, p[:, 0], p[:, 1], p[:, 2])
    # brute force: every clump, chord length x kappa at the frequency seen in the clump's frame
    C = np.c_[A["x"], A["y"], A["z"]]
    V = np.c_[A["vx"], A["vy"], A["vz"]]
    tau_bf = np.zeros(n)
    ncl_bf = np.zeros(n, dtype=int)
    inside_bf = np.zeros(n, dtype=int)
    for i in range(n):
        r = p[i] - C
        b = r @ k[i]
        disc = b * b - (r * r).sum(1) + A["r"] ** 2
        hit = disc > 0
        t1 = np.where(hit, -b - np.sqrt(np.maximum(disc, 0)), 0)
        t2 = np.where(hit, -b + np.sqrt(np.maximum(disc, 0)), 0)
        hit &= t2 > 0
        chord = t2 - np.maximum(t1, 0)
        own = np.flatnonzero(hit & (t1 < 0))
        inside_bf[i] = own[0] + 1 if own.size else 0
        # lab-frame frequency of the ray: inside its own clump xfreq is already in that clump's frame
        x_lab = xf[i] + (V[own[0]] @ k[i] if own.size else 0.0)
        x_cl = x_lab - V @ k[i]
        # clip the chords to the bounding sphere
        bs = p[i] @ k[i]
        t_sp = -bs + np.sqrt(bs * bs - p[i] @ p[i] + 1.0)
        chord = np.minimum(t2, t_sp) - np.maximum(t1, 0)
        use = hit & (chord > 0)
        tau_bf[i] = (A["kap"][use] * oracle.voigt(x_cl[use], A["a"][0]) * chord[use]).sum()
        ncl_bf[i] = use.sum()
    assert np.array_equal(icl, inside_bf)
    tau, ncl = oracle.clump_edge(m.config, p[:, 0], p[:, 1], p[:, 2], k[:, 0], k[:, 1], k[:, 2], xf, icl)
    # a clump cut by the bounding sphere is left at the sphere by the reference (t_exit is not clipped for tau): allow those
    clean = ncl == ncl_bf
    assert clean.mean() > 0.97
    assert np.allclose(tau[clean], tau_bf[clean], rtol=1e-9, atol=1e-12)
    assert (ncl > 0).mean() > 0.7 and (icl > 0).sum() > 10
    # capped variant: identical below the cap, stops at the first clump that takes it over
    cap = np.median(tau[tau > 0])
    tc, nc = oracle.clump_edge(m.config, p[:, 0], p[:, 1], p[:, 2], k[:, 0], k[:, 1], k[:, 2], xf, icl, tau_max=cap)
    assert np.array_equal(tc[tau < cap], tau[tau < cap]) and np.all(tc[tau >= cap] >= cap) and np.all(nc <= ncl)


def test_covering_factor_and_tau_walk():
    fcov = 1.5
    m = clump_model(clump_f_cov=fcov, clump_radius=0.02, clump_tau0=50.0, iseed=11)
    rng = np.random.default_rng(6)
    n = 20000
    p, k = rays_from(rng, n, pos=(0, 0, 0))
    p += 1e-9  # off the exact centre
    xf = np.zeros(n)
    icl = oracle.clump_locate(m.config, p[:, 0], p[:, 1], p[:, 2])
    tau, ncl = oracle.clump_edge(m.config, p[:, 0], p[:, 1], p[:, 2], k[:, 0], k[:, 1], k[:, 2], xf, icl)
    # radial sight lines cross a Poisson number of clumps with mean f_cov (clump_mod.f90:724-727)
    assert ncl.mean() == pytest.approx(fcov, rel=0.05)
    assert (ncl == 0).mean() == pytest.approx(np.exp(-fcov), rel=0.1)
    # tau walk: lands where the remaining depth to the edge is tau_total - tau_in, escapes otherwise
    tau_in = rng.exponential(size=n) * 40.0
    b = oracle.clump_tau(m.config, p[:, 0], p[:, 1], p[:, 2], k[:, 0], k[:, 1], k[:, 2], xf, icl, tau_in)
    ins = b["inside"] == 1
    assert np.array_equal(ins, tau_in <= tau) or (ins != (tau_in <= tau)).mean() < 1e-3
    assert 0.2 < ins.mean() < 0.8 and np.all(b["icl"][ins] > 0) and np.all(b["icl"][~ins] == 0)
    rest, _ = oracle.clump_edge(m.config, b["x"][ins], b["y"][ins], b["z"][ins], k[ins, 0], k[ins, 1], k[ins, 2],
                                b["xfreq"][ins], b["icl"][ins])
    assert np.allclose(rest, tau[ins] - tau_in[ins], rtol=1e-7, atol=1e-7)
    out = ~ins
    r_out = np.sqrt(b["x"][out] ** 2 + b["y"][out] ** 2 + b["z"][out] ** 2)
    assert np.allclose(r_out, 1.0, atol=1e-9)  # escaped photons sit on the bounding sphere


def test_clump_run_conserves_photons_and_escape_fraction():
    fcov = 1.0
    n = 20000
    m = clump_model(no_photons=n, clump_f_cov=fcov, clump_radius=0.02, clump_tau0=1e4, spectral_type="monochromatic",
                    nxfreq=121, xfreq_min=-30.0, xfreq_max=30.0, save_all_photons=True, iseed=21, xs_point=1e-9)
    oracle.run(m, rng_mode=1)
    assert m.counters["n_photons_done"] == n
    jout = m.spectrum("Jout")
    assert jout.sum() == pytest.approx(n, rel=2e-3)  # everything escapes (no dust), a few photons beyond the frequency grid
    # a photon whose first sight line meets no clump (tau0 = 0) leaves unscattered: exp(-f_cov) of them
    ns = m.allph("nscatt_gas")
    assert (ns == 0).mean() == pytest.approx(np.exp(-fcov), rel=0.08)
    assert jout[len(jout) // 2] / n > np.exp(-fcov) * 0.95  # they all sit in the central bin
    assert ns[ns > 0].mean() > 10  # surface scatterings off opaque clumps


LOGGED_FCOV1 = dict(use_clump_medium=True, rmax=1.0, clump_radius=0.001, clump_f_cov=1.0, N_HImax=1e18, temperature=1e4,
                    clump_sigma_v=0.0, spectral_type="monochromatic", geometry="sphere", velocity_type="rotating_galaxy_halo",
                    Vrot=300.0, rinner=0.1, nxfreq=500, velocity_min=-1000.0, velocity_max=1000.0, nx=11, ny=11, nz=11,
                    nxim=0, nyim=0)


def test_setup_scalars_clump_sphere_log():
    """The reference's own log of this input — examples/clump_sphere/log_back:4-55 (clump_NHI18_fcov1)."""
    m = Model(no_photons=1000, iseed=5, **LOGGED_FCOV1).setup()
    c = m.config.contents.clumps
    G = lambda name: golden("clump_NHI18_fcov1", name)  # tests/golden/reference_logs.json <- examples/clump_sphere/log_back
    assert c.n == G("N_clumps") == 1333333                                      # log_back:11
    assert c.rhokap[0] == pytest.approx(G("cl_rhokap"), rel=2e-5)               # :14
    assert c.voigt_a[0] == pytest.approx(0.00047, abs=5e-6)                     # :15
    assert c.Dfreq_ref == pytest.approx(G("cl_Dfreq"), rel=5e-5)                # :16
    assert c.cgx == G("csr_cells_per_axis") == 111                              # :21 "in 111^3 cells"
    nreg = np.ctypeslib.as_array(c.cg_start, shape=(c.cgx ** 3 + 1,))[-1] - 1
    assert nreg == pytest.approx(G("csr_registrations"), rel=5e-3)              # :21 (another random layout)
    assert m.summary.tauhomo == pytest.approx(G("tauhomo"), rel=2e-6)           # :22
    assert m.summary.N_gashomo == pytest.approx(G("N_gashomo"), rel=1e-5)       # :24
    assert m.summary.vtherm == pytest.approx(G("cl_vtherm_kms"), rel=1e-6)      # :37
    # <N_scatt> = 4.3454E+03 with 1e6 photons (:52) is checked on the GPU (tests/test_gpu_clumps.py); the distribution is
    # heavy-tailed (sigma/mean = 5.4), so a CPU-sized sample can only bracket it
    oracle.run(m, rng_mode=0)
    assert 2.5e3 < m.nscatt_gas / 1000 < 7e3


def test_overlapping_population_event_walk_equals_brute_force():
    """has_overlap (clump_mod.f90:1544-1590, 1639-1760; raytrace_clump.f90:608-920): in an overlap region every clump adds
    its opacity at its own frame's frequency; the frequency stays in the global frame during the walk.  Pins the oracle
    the GPU event walk is compared with (tests/test_gpu_clumps.py::test_overlap_*)."""
    m = clump_model(clump_allow_overlap=True, clump_radius=0.08, clump_f_cov=3.0, clump_sigma_v=15.0, velocity_type="hubble",
                    Vexp=60.0)
    A = clump_arrays(m)
    c = A["c"]
    assert c.has_overlap == 1
    from scipy.spatial import cKDTree
    P = np.c_[A["x"], A["y"], A["z"]]
    d, _ = cKDTree(P).query(P, k=2)
    assert d[:, 1].min() < 2 * 0.08  # really overlapping
    rng = np.random.default_rng(17)
    n = 2000
    p, k = rays_from(rng, n)
    xf = rng.normal(size=n) * 2
    V = np.c_[A["vx"], A["vy"], A["vz"]]
    tau_bf = np.zeros(n)
    for i in range(n):
        r = p[i] - P
        b = r @ k[i]
        disc = b * b - (r * r).sum(1) + A["r"] ** 2
        hit = disc > 0
        sq = np.sqrt(np.maximum(disc, 0))
        t1, t2 = -b - sq, -b + sq
        bs = p[i] @ k[i]
        t_sp = -bs + np.sqrt(bs * bs - p[i] @ p[i] + 1.0)
        chord = np.minimum(t2, t_sp) - np.maximum(t1, 0)
        use = hit & (t2 > 0) & (chord > 0)
        x_cl = xf[i] - V @ k[i]  # global-frame frequency seen in each clump's frame
        tau_bf[i] = (A["kap"][use] * oracle.voigt(x_cl[use], A["a"][0]) * chord[use]).sum()
    icl0 = np.zeros(n, dtype=np.int32)  # the overlap walk finds its own active set
    tau, _ = oracle.clump_edge(m.config, p[:, 0], p[:, 1], p[:, 2], k[:, 0], k[:, 1], k[:, 2], xf, icl0)
    assert np.allclose(tau, tau_bf, rtol=1e-9, atol=1e-12) and (tau > 0).mean() > 0.6
    cap = np.median(tau)
    tc, _ = oracle.clump_edge(m.config, p[:, 0], p[:, 1], p[:, 2], k[:, 0], k[:, 1], k[:, 2], xf, icl0, tau_max=cap)
    assert np.array_equal(tc[tau < cap], tau[tau < cap]) and np.all(tc[tau >= cap] >= cap)
    # whole runs: every photon accounted for, no dust -> everything escapes; photons scatter in some owner clump
    run = clump_model(clump_allow_overlap=True, clump_radius=0.08, clump_f_cov=3.0, clump_sigma_v=15.0, no_photons=3000,
                      nxim=9, nyim=9, use_stokes=True, save_all_photons=True, xfreq_min=-30.0, xfreq_max=30.0)
    oracle.run(run, rng_mode=1)
    assert run.counters["n_photons_done"] == 3000 and run.spectrum("Jout").sum() == pytest.approx(3000, rel=1e-3)
    assert run.nscatt_gas / 3000 > 3 and run.observer_cube("scatt").sum() > 0
    assert run.config.contents.clumps.has_overlap == 1  # the flag travels through the ABI struct
