"""GPU parity of the deterministic pieces and the samplers, through the C ABI.

Bars (BASELINE.json north_star): Voigt H(a,x) <= 1e-12 relative (we get bit-exact);
DDA cell sequences bit-exact and tau within 1e-12 (we get bit-exact); Philox uniforms
bit-exact; variates equal to the oracle's up to libm rounding.
"""
import numpy as np
import pytest

from lart_b200 import Model, Simulation, calc_voigt, sample
from oracle import oracle

pytestmark = pytest.mark.gpu


def random_rays(m, n, seed, on_faces=True):
    g = m.config.contents.grid
    rng = np.random.default_rng(seed)
    lo = np.array([g.xmin, g.ymin, g.zmin])
    d = np.array([g.dx, g.dy, g.dz])
    nn = np.array([g.nx, g.ny, g.nz])
    p = lo + rng.uniform(0, 1, (n, 3)) * d * nn
    k = rng.normal(size=(n, 3))
    k /= np.linalg.norm(k, axis=1)[:, None]
    if on_faces:  # the special cases of setup_traversal_car: on-face starts, axis-parallel, ties
        q = n // 10
        p[:q, 0] = lo[0] + rng.integers(0, nn[0] + 1, q) * d[0]       # exactly on x faces
        p[q:2 * q, 2] = lo[2] + rng.integers(0, nn[2] + 1, q) * d[2]  # exactly on z faces
        k[2 * q:3 * q] = np.eye(3)[rng.integers(0, 3, q)] * rng.choice([-1.0, 1.0], q)[:, None]  # axis parallel
        s = 1 / np.sqrt(2.0)
        k[3 * q:4 * q] = [s, s, 0.0]
        p[3 * q:4 * q] = lo + (rng.integers(0, nn, (q, 3)) + 0.5) * d  # cell centres + diagonal: exact ties
    ic = np.clip(np.floor((p - lo) / d).astype(np.int32) + 1, 1, nn)
    xf = rng.uniform(-12, 12, n)
    xf[: n // 4] = rng.normal(size=n // 4)
    return p, k, ic, xf


def test_voigt_bit_exact_on_dense_lattice():
    x = np.concatenate([np.linspace(-15, 15, 600001), [0.0, 1.0, 5.0, 10.0, -1.0, -5.0, -10.0, 1e3, 1e-300]])
    for a in (4.7186e-4, 1.4921e-2, 1e-6, 0.1):
        Hg, Ho = calc_voigt(x, a), oracle.voigt(x, a)
        assert np.array_equal(Hg, Ho)
        assert np.max(np.abs(Hg / Ho - 1)) <= 1e-12  # the stated bar
    assert calc_voigt(np.zeros(0), 1e-3).size == 0


@pytest.mark.parametrize("flags", [0, 1])
def test_raytrace_to_edge_cells_and_tau(flags):
    m = Model(no_photons=10, temperature=1e4, N_HI=2e18, nx=41, ny=41, nz=41, rmax=1.0, velocity_type="hubble",
              Vexp=200.0, nxfreq=50, xfreq_min=-40, xfreq_max=10).setup()
    sim = Simulation(m, flags=flags, pool_slots=1024)
    n = 200000
    p, k, ic, xf = random_rays(m, n, 5)
    cap = 128
    tg, ng, trg = sim.raytrace_to_edge(p[:, 0], p[:, 1], p[:, 2], k[:, 0], k[:, 1], k[:, 2], xf, ic[:, 0], ic[:, 1],
                                       ic[:, 2], trace_cap=cap)
    to, no, tro = oracle.raytrace_to_edge(m.config, p[:, 0], p[:, 1], p[:, 2], k[:, 0], k[:, 1], k[:, 2], xf, ic[:, 0],
                                          ic[:, 1], ic[:, 2], trace_cap=cap)
    assert np.array_equal(ng, no)
    assert np.array_equal(trg, tro)            # cell sequence, bit-exact
    assert np.array_equal(tg, to)              # tau, bit-exact (bar: 1e-12)
    assert ng.max() > 40 and (ng == 0).sum() > 0  # long walks and already-leaving rays both occur
    sim.close()


def test_raytrace_to_edge_tau_cap():
    m = Model(no_photons=10, temperature=1e4, taumax=1e7, nx=21, ny=21, nz=21, rmax=1.0, nxfreq=11).setup()
    sim = Simulation(m, pool_slots=1024)
    n = 20000
    p, k, ic, xf = random_rays(m, n, 9)
    xf[:] = 0.0
    tg, ng, _ = sim.raytrace_to_edge(p[:, 0], p[:, 1], p[:, 2], k[:, 0], k[:, 1], k[:, 2], xf, ic[:, 0], ic[:, 1], ic[:, 2])
    to, no, _ = oracle.raytrace_to_edge(m.config, p[:, 0], p[:, 1], p[:, 2], k[:, 0], k[:, 1], k[:, 2], xf, ic[:, 0], ic[:, 1], ic[:, 2])
    assert np.array_equal(tg, to) and np.array_equal(ng, no)
    assert (ng == 1).mean() > 0.3  # tau >= 745.2 stops the walk early (raytrace_car.f90:432,497)
    sim.close()


def test_raytrace_to_tau_photon_state():
    m = Model(no_photons=10, temperature=1e4, N_HI=2e19, nx=41, ny=41, nz=41, rmax=1.0, velocity_type="hubble",
              Vexp=200.0, nxfreq=50, xfreq_min=-40, xfreq_max=10).setup()
    sim = Simulation(m, pool_slots=1024)
    n = 200000
    p, k, ic, xf = random_rays(m, n, 6)
    tau_in = np.random.default_rng(7).exponential(size=n) * 3
    a = sim.raytrace_to_tau(p[:, 0], p[:, 1], p[:, 2], k[:, 0], k[:, 1], k[:, 2], xf, ic[:, 0], ic[:, 1], ic[:, 2], tau_in)
    b = oracle.raytrace_to_tau(m.config, p[:, 0], p[:, 1], p[:, 2], k[:, 0], k[:, 1], k[:, 2], xf, ic[:, 0], ic[:, 1], ic[:, 2], tau_in)
    for key in ("inside", "icell", "jcell", "kcell", "nsteps", "x", "y", "z", "xfreq", "xfreq_ref"):
        assert np.array_equal(a[key], b[key]), key
    frac = a["inside"].mean()
    assert 0.05 < frac < 0.95
    sim.close()


BOUNDARY_GRIDS = {
    "xyzsym_even": dict(nx=16, ny=16, nz=16, rmax=1.0, xyz_symmetry=True),
    "xyzsym_odd": dict(nx=15, ny=15, nz=15, rmax=1.0, xyz_symmetry=True),
    "xyzsym_mixed": dict(nx=16, ny=15, nz=12, rmax=1.0, xyz_symmetry=True),
    "xysym_even": dict(nx=16, ny=16, nz=24, rmax=1.0, xy_symmetry=True),
    "xysym_odd": dict(nx=15, ny=13, nz=21, rmax=1.0, xy_symmetry=True),
    "xyper_box": dict(nx=7, ny=5, nz=24, xmax=0.5, ymax=0.25, zmax=1.0, geometry="rectangle", xy_periodic=True),
}


@pytest.mark.parametrize("name", sorted(BOUNDARY_GRIDS))
def test_boundary_variant_rays_bit_exact(name):
    """The folded and periodic ray tracers — _xyzsym (raytrace_car.f90:584-760, 1650-1949), _xysym (:783-969, 1951-2250),
    _xyper (:971-1136, 2252-2517): starts exactly on the mirror / periodic faces, walks reflected or wrapped many times."""
    kw = BOUNDARY_GRIDS[name]
    m = Model(no_photons=10, temperature=1e4, N_HI=1e15, velocity_type="hubble", Vexp=100.0, nxfreq=50, xfreq_min=-40,
              xfreq_max=10, **kw).setup()
    sim = Simulation(m, pool_slots=1024)
    n = 100000
    p, k, ic, xf = random_rays(m, n, 21)
    g = m.config.contents.grid
    nn = np.array([g.nx, g.ny, g.nz])
    q = n // 10
    p[4 * q:5 * q, 0] = g.xmin  # on the lower (mirror / periodic) planes themselves
    p[5 * q:6 * q, 2] = g.zmin
    p[6 * q:7 * q, 1] = g.ymax
    k[np.abs(k[:, 2]) < 0.02, 2] = 0.3  # a horizontal ray in a periodic box never ends
    k /= np.linalg.norm(k, axis=1)[:, None]
    ic = np.floor((p - [g.xmin, g.ymin, g.zmin]) / [g.dx, g.dy, g.dz]).astype(np.int32) + 1
    if "xyper" not in name:
        ic = np.clip(ic, 1, nn)  # the periodic variants see photons in cell n+1 on the upper face, as upstream
    else:
        ic[:, 2] = np.clip(ic[:, 2], 1, g.nz)
    cols = lambda a: (a[:, 0], a[:, 1], a[:, 2])
    cap = 96
    tg, ng, trg = sim.raytrace_to_edge(*cols(p), *cols(k), xf, *cols(ic), trace_cap=cap)
    to, no, tro = oracle.raytrace_to_edge(m.config, *cols(p), *cols(k), xf, *cols(ic), trace_cap=cap)
    assert np.array_equal(ng, no) and np.array_equal(trg, tro) and np.array_equal(tg, to)
    assert ng.max() > nn.max() + 2  # reflected / wrapped walks are longer than any straight one
    tau_in = np.random.default_rng(8).exponential(size=n) * np.median(to[to > 0])
    a = sim.raytrace_to_tau(*cols(p), *cols(k), xf, *cols(ic), tau_in)
    b = oracle.raytrace_to_tau(m.config, *cols(p), *cols(k), xf, *cols(ic), tau_in)
    for key in ("inside", "icell", "jcell", "kcell", "nsteps", "x", "y", "z", "xfreq", "xfreq_ref"):
        assert np.array_equal(a[key], b[key]), key
    assert 0.05 < a["inside"].mean() < 0.95
    ins = a["inside"] == 1
    # (upstream quirk kept: a start exactly on a straddling mirror plane, x = xmin = -dx/2, is reflected into cell i0 = 2
    # with its distances still measured from -dx/2, so such a photon can land up to dx beyond xmax)
    assert a["x"][ins].min() >= g.xmin and a["y"][ins].min() >= g.ymin and a["x"][ins].max() <= g.xmax + g.dx
    sim.close()


def test_slab_zonly_rays():
    m = Model(no_photons=10, temperature=1e4, taumax=1e4, xy_periodic=True, nx=1, ny=1, nz=201, nxfreq=121).setup()
    sim = Simulation(m, pool_slots=1024)
    rng = np.random.default_rng(3)
    n = 50000
    z = rng.uniform(-1, 1, n)
    z[:100] = m.grid_array("zface")[rng.integers(0, 202, 100)]
    k = rng.normal(size=(n, 3))
    k /= np.linalg.norm(k, axis=1)[:, None]
    k[100:200] = [1.0, 0.0, 0.0]  # parallel to the slab: never leaves; tau runs into the 745.2 cap
    kc = np.clip(np.floor((z + 1) / 0.00995024875621890547).astype(np.int32) + 1, 1, 201)
    kc = np.clip(np.floor((z - (-1.0)) / m.config.contents.grid.dz).astype(np.int32) + 1, 1, 201)
    xf = rng.normal(size=n) * 3
    one = np.ones(n, dtype=np.int32)
    zero = np.zeros(n)
    sel = np.ones(n, dtype=bool)
    sel[100:200] = False  # (the parallel rays would walk forever in both implementations at tau ~ 0)
    args = lambda s: (zero[s], zero[s], z[s], k[s, 0], k[s, 1], k[s, 2], xf[s], one[s], one[s], kc[s])
    tg, ng, _ = sim.raytrace_to_edge(*args(sel))
    to, no, _ = oracle.raytrace_to_edge(m.config, *args(sel))
    assert np.array_equal(tg, to) and np.array_equal(ng, no)
    tau_in = rng.exponential(size=n) * 10
    a = sim.raytrace_to_tau(*args(sel), tau_in[sel])
    b = oracle.raytrace_to_tau(m.config, *args(sel), tau_in[sel])
    for key in ("inside", "kcell", "nsteps", "z", "x", "xfreq", "xfreq_ref"):
        assert np.array_equal(a[key], b[key]), key
    sim.close()


def test_xcrit_local():
    m = Model(no_photons=10, temperature=1e4, taumax=1e7, nx=41, ny=41, nz=41, rmax=1.0, nxfreq=11, core_skip=True).setup()
    sim = Simulation(m, pool_slots=1024)
    p, k, ic, xf = random_rays(m, 50000, 8, on_faces=False)
    a = sim.xcrit_local(p[:, 0], p[:, 1], p[:, 2], ic[:, 0], ic[:, 1], ic[:, 2])
    b = oracle.xcrit_local(m.config, p[:, 0], p[:, 1], p[:, 2], ic[:, 0], ic[:, 1], ic[:, 2])
    assert np.allclose(a, b, rtol=1e-14, atol=0)  # cbrt vs pow(.,1/3): last-bit differences only
    assert (a > 0).mean() > 0.2 and (a == 0).mean() > 0.2
    sim.close()


def test_philox_uniforms_bit_exact():
    ids = np.array([1, 2, 3, 2 ** 33 + 5, 10 ** 12], dtype=np.int64)
    for seed in (0, 12345, 2 ** 63 + 11):
        assert np.array_equal(sample(0, seed, ids, ndraw=257), oracle.sample(0, seed, ids, ndraw=257))


@pytest.mark.parametrize("kind,p0,p1", [(1, None, None), (3, 1.0, None), (3, 0.0, None), (3, -0.5, None),
                                        (4, 0.6761, None), (4, 0.0, None), (5, 4.7186e-4, None)])
def test_variates_match_oracle_streams(kind, p0, p1):
    ids = np.arange(1, 20001, dtype=np.int64)
    a = sample(kind, 42, ids, p0, p1, ndraw=8)
    b = oracle.sample(kind, 42, ids, p0, p1, ndraw=8)
    close = np.isclose(a, b, rtol=1e-11, atol=1e-13)
    assert close.mean() > 0.9999, close.mean()


@pytest.mark.parametrize("x0", [0.0, 0.3, 1.0, 1.7, 2.4142135, 2.5, 3.3, 4.5, 6.0, 12.0, -2.0, -7.0])
@pytest.mark.parametrize("a", [4.7186e-4, 1.4921e-2])
def test_rand_resonance_vz_matches_oracle_streams(x0, a):
    ids = np.arange(1, 20001, dtype=np.int64)
    g = sample(2, 7, ids, x0, a, ndraw=4)
    o = oracle.sample(2, 7, ids, x0, a, ndraw=4)
    # u = x0 + a*tan(theta) cancels when the atom is nearly at rest (|u| << x0): an ulp of x0 on |u| ~ 1e-3,
    # amplified by theta sitting ~a/x0 from the pole of tan -> absolute agreement ~1e-10, relative elsewhere
    close = np.isclose(g, o, rtol=1e-9, atol=2e-10 * max(1.0, abs(x0)))
    assert close.mean() > 0.9995, (x0, a, close.mean())
    # and the distributions agree regardless (two-sample KS against the oracle's MT19937-64 stream)
    from scipy.stats import ks_2samp
    mt = oracle.sample(2, 99, ids, x0, a, ndraw=4, rng_mode=0)
    assert ks_2samp(g.ravel(), mt.ravel()).pvalue > 1e-4


def test_warp_cooperative_sampler_equals_serial():
    """The warp-cooperative atom-velocity sampler (kind 6) evaluates trials of many photons in parallel but must
    return, photon by photon and draw by draw, exactly what the serial rejection loop (kind 2) returns."""
    rng = np.random.default_rng(12)
    n = 50000
    ids = np.arange(1, n + 1, dtype=np.int64)
    x0 = rng.normal(size=n) * 2.5          # mixes |x|<=1 photons and all three wing variants inside every warp
    x0[::7] = rng.uniform(-12, 12, x0[::7].size)
    a = np.where(rng.uniform(size=n) < 0.5, 4.7186e-4, 1.4921e-2)
    serial = sample(2, 31, ids, x0, a, ndraw=3)
    coop = sample(6, 31, ids, x0, a, ndraw=3)
    assert np.array_equal(serial, coop)
    # ragged tail: a batch that does not fill its last warp
    assert np.array_equal(sample(2, 5, ids[:45], x0[:45], a[:45], ndraw=2), sample(6, 5, ids[:45], x0[:45], a[:45], ndraw=2))


def test_sightline_maps_match_oracle_and_column_density():
    """Sight-line maps: tau_gas(nu, pixel), N_gas and tau_dust per pixel, bit for bit against the oracle;
    the central pixel's column density is 2 N_pole, and tau_gas at line centre is 2 tau0 there (static sphere)."""
    from lart_b200 import Model
    for kw in (dict(taumax=1e3), dict(N_HI=2e19, velocity_type="hubble", Vexp=200.0, xfreq_min=-40.0, xfreq_max=20.0,
                                      DGR=1.0, cext_dust=1e-21, use_stokes=False),
               dict(taumax=50.0, obsx=[0.3, -1.0], obsy=[0.2, 0.1], obsz=[1.0, 0.4], geometry="rectangle", rmax=-999.0,
                    nx=12, ny=20, nz=16, xmax=0.6, ymax=1.0, zmax=0.8)):
        par = dict(no_photons=10, temperature=1e4, nx=41, ny=41, nz=41, rmax=1.0, nxfreq=31, nxim=21, nyim=17, use_stokes=True)
        par.update(kw)
        m = Model(**par).setup()
        sim = Simulation(m, pool_slots=1024)
        g = sim.sightline_tau()
        o, osteps = oracle.sightline_tau(m)
        assert len(g) == m.config.contents.par.nobs
        for a, b in zip(g, o):
            assert np.array_equal(a["tau_gas"], b["tau_gas"])
            assert np.array_equal(a["N_gas"], b["N_gas"])
            if b["tau_dust"] is not None:
                assert np.array_equal(a["tau_dust"], b["tau_dust"]) and a["tau_dust"].max() > 0
        assert sim.sightline_stats["cellsteps"] == osteps
        sim.close()
    m = Model(no_photons=10, temperature=1e4, taumax=1e3, nx=41, ny=41, nz=41, rmax=1.0, nxfreq=31, nxim=21, nyim=21,
              use_stokes=True).setup()
    sim = Simulation(m, pool_slots=1024)
    mp = sim.sightline_tau()[0]
    s = m.summary
    assert mp["N_gas"][10, 10] == pytest.approx(2 * s.N_gaspole, rel=1e-9)
    assert mp["tau_gas"][15, 10, 10] == pytest.approx(2 * 1e3, rel=1e-9)  # bin 16 of 31 is centred on x = 0
    assert mp["N_gas"][0, 0] == 0.0 or mp["N_gas"][0, 0] < mp["N_gas"][10, 10]
    sim.close()


PEEL_BOUND_GRIDS = {
    "sphere_thick": dict(nx=31, ny=31, nz=31, rmax=1.0, taumax=1e7),
    "sphere_thick_dust": dict(nx=21, ny=21, nz=21, rmax=1.0, taumax=3e6, DGR=1.0, cext_dust=3e-17, use_stokes=False),
    "hubble_thick": dict(nx=31, ny=31, nz=31, rmax=1.0, taumax=1e7, velocity_type="hubble", Vexp=200.0),
    "xysym_thick": dict(nx=16, ny=15, nz=31, rmax=1.0, taumax=1e7, xy_symmetry=True),
    "xyper_box_thick": dict(nx=7, ny=5, nz=41, xmax=0.5, ymax=0.25, zmax=1.0, geometry="rectangle", xy_periodic=True, taumax=1e7),
    "slab_zonly_thick": dict(nx=1, ny=1, nz=201, xy_periodic=True, rmax=-999.0, taumax=1e7, nxim=0, nyim=0),
}


@pytest.mark.parametrize("name", sorted(PEEL_BOUND_GRIDS))
def test_peel_bound_implies_capped_walk(name):
    """The scatter stage counts a peel ray without walking it when kappa(x) * (distance to the nearest face of its cell)
    >= tau_huge = 745.2 (raytrace_car.f90:432,497).  Whenever that bound fires, the reference's walk must indeed stop in
    the first cell at the cap — for any direction, on faces (L = 0), with dust, on folded and periodic grids."""
    kw = dict(no_photons=10, temperature=1e4, nxfreq=21)
    kw.update(PEEL_BOUND_GRIDS[name])
    m = Model(**kw).setup()
    sim = Simulation(m, pool_slots=1024)
    g = m.config.contents.grid
    n = 120000
    rng = np.random.default_rng(3)
    lo, d, nn = np.array([g.xmin, g.ymin, g.zmin]), np.array([g.dx, g.dy, g.dz]), np.array([g.nx, g.ny, g.nz])
    ic = rng.integers(1, nn + 1, (n, 3)).astype(np.int32)
    u = rng.uniform(0, 1, (n, 3))
    q = n // 8
    zonly = g.nx == 1 and g.ny == 1
    u[:q, 2 if zonly else rng.integers(0, 3)] = 0.0  # exactly on a lower face: L = 0 (a z-only slab has z faces only)
    u[q:2 * q] = rng.choice([1e-12, 1e-9, 1e-6, 1e-3], (q, 3))  # a hair inside
    u[2 * q:3 * q, 2] = 1.0                          # on the upper z face of the cell
    faces = [m.grid_array(a) for a in ("xface", "yface", "zface")]
    p = np.stack([faces[a][ic[:, a] - 1] + u[:, a] * (faces[a][ic[:, a]] - faces[a][ic[:, a] - 1]) for a in range(3)], axis=1)
    p[2 * q:3 * q, 2] = faces[2][ic[2 * q:3 * q, 2]]
    if zonly:
        p[:, 0] = 0.0; p[:, 1] = 0.0
    k = rng.normal(size=(n, 3))
    k /= np.linalg.norm(k, axis=1)[:, None]
    # frequencies on both sides of the threshold kappa*L = 745.2: from line centre far into the wings
    xf = np.concatenate([rng.normal(size=n // 2) * 2.0, rng.uniform(-60, 60, n - n // 2)])
    capped = sim.peel_bound(p[:, 0], p[:, 1], p[:, 2], xf, ic[:, 0], ic[:, 1], ic[:, 2])
    to, no, _ = oracle.raytrace_to_edge(m.config, p[:, 0], p[:, 1], p[:, 2], k[:, 0], k[:, 1], k[:, 2], xf, ic[:, 0], ic[:, 1], ic[:, 2])
    tg, ng, _ = sim.raytrace_to_edge(p[:, 0], p[:, 1], p[:, 2], k[:, 0], k[:, 1], k[:, 2], xf, ic[:, 0], ic[:, 1], ic[:, 2])
    assert np.array_equal(tg, to) and np.array_equal(ng, no)
    c = capped == 1
    assert c.sum() > 0.02 * n and (~c).sum() > 0.1 * n, c.mean()  # both sides of the bound are sampled
    assert (to[c] >= 745.2).all()   # the walk does end at the cap ...
    assert (no[c] == 1).all()       # ... inside the first cell: one cell step, contribution exp(-tau) == 0
    assert np.exp(-to[c]).max() == 0.0
    assert (capped[:q] == 0).all()  # on a face the bound never fires
    # the bound is tight enough to matter: it catches most of the rays that do end in their first cell at the cap
    first_cell_cap = (no == 1) & (to >= 745.2)
    assert c.sum() > 0.2 * first_cell_cap.sum()
    sim.close()
