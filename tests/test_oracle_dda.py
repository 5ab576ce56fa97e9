"""Properties of the oracle's DDA restatement (CPU only): analytic optical depths, cell
sequences, the on-face / axis-parallel / tie rules of setup_traversal_car."""
import numpy as np
import pytest

from lart_b200 import Model
from oracle import oracle


def uniform_box(n=8, tau=1.0, **kw):
    return Model(no_photons=10, temperature=1e4, taumax=tau, nx=n, ny=n, nz=n, geometry="rectangle", nxfreq=11, **kw).setup()


def test_tau_is_opacity_times_path_length():
    m = uniform_box()
    rng = np.random.default_rng(0)
    n = 2000
    p = rng.uniform(-0.99, 0.99, (n, 3))
    k = rng.normal(size=(n, 3))
    k /= np.linalg.norm(k, axis=1)[:, None]
    g = m.config.contents.grid
    ic = np.floor((p - [g.xmin, g.ymin, g.zmin]) / [g.dx, g.dy, g.dz]).astype(np.int32) + 1
    xf = rng.uniform(-3, 3, n)
    tau, ns, _ = oracle.raytrace_to_edge(m.config, p[:, 0], p[:, 1], p[:, 2], k[:, 0], k[:, 1], k[:, 2], xf, ic[:, 0],
                                         ic[:, 1], ic[:, 2])
    t = np.min(np.where(k > 0, (1 - p) / k, (-1 - p) / k), axis=1)  # distance to the box
    kap = m.grid_array("rhokap")[0, 0, 0] * oracle.voigt(xf, m.summary.voigt_a)
    assert np.allclose(tau, kap * t, rtol=1e-12)
    assert ns.min() >= 1 and ns.max() <= 3 * 8


def test_axis_parallel_and_face_rules():
    m = uniform_box(n=4, tau=2.0)
    cfg = m.config
    one = lambda v: np.array([v], dtype=float)
    # +x from the centre of cell (3,3,3): crosses cells 3,4 -> 2 steps, trace = linear indices
    tau, ns, tr = oracle.raytrace_to_edge(cfg, one(0.25), one(0.25), one(0.25), one(1), one(0), one(0), one(0.0),
                                          [3], [3], [3], trace_cap=8)
    assert ns[0] == 2 and list(tr[0, :2]) == [2 + 4 * (2 + 4 * 2), 3 + 4 * (2 + 4 * 2)]
    # photon sitting exactly on the lower face of cell 3 and moving -x is moved into cell 2 (raytrace_car.f90:36-44)
    tau2, ns2, tr2 = oracle.raytrace_to_edge(cfg, one(0.0), one(0.25), one(0.25), one(-1), one(0), one(0), one(0.0),
                                             [3], [3], [3], trace_cap=8)
    assert ns2[0] == 2 and list(tr2[0, :2]) == [1 + 4 * (2 + 4 * 2), 0 + 4 * (2 + 4 * 2)]
    # on the lower face of cell 1 moving outwards: already leaving, tau = 0, no step
    tau3, ns3, _ = oracle.raytrace_to_edge(cfg, one(-1.0), one(0.25), one(0.25), one(-1), one(0), one(0), one(0.0),
                                           [1], [3], [3])
    assert tau3[0] == 0.0 and ns3[0] == 0
    # tie between x and y crossings resolves to x first (minloc)
    s = 1 / np.sqrt(2)
    _, ns4, tr4 = oracle.raytrace_to_edge(cfg, one(0.25), one(0.25), one(0.25), one(s), one(s), one(0), one(0.0),
                                          [3], [3], [3], trace_cap=8)
    c = lambda i, j, k: (i - 1) + 4 * ((j - 1) + 4 * (k - 1))
    assert list(tr4[0, :ns4[0]]) == [c(3, 3, 3), c(4, 3, 3), c(4, 4, 3)]


def test_tau_walk_lands_at_requested_depth():
    m = uniform_box(n=16, tau=50.0)
    rng = np.random.default_rng(1)
    n = 1000
    k = rng.normal(size=(n, 3))
    k /= np.linalg.norm(k, axis=1)[:, None]
    z = np.zeros(n)
    ic = np.full(n, 9, dtype=np.int32)
    tau_in = rng.exponential(size=n) * 5
    out = oracle.raytrace_to_tau(m.config, z, z, z, k[:, 0], k[:, 1], k[:, 2], z, ic, ic, ic, tau_in)
    kap = m.grid_array("rhokap")[0, 0, 0] * oracle.voigt([0.0], m.summary.voigt_a)[0]
    d = np.sqrt(out["x"] ** 2 + out["y"] ** 2 + out["z"] ** 2)
    ins = out["inside"] == 1
    assert ins.sum() > 900
    assert np.allclose(d[ins] * kap, tau_in[ins], rtol=1e-10)
    g = m.config.contents.grid
    assert np.array_equal(out["icell"][ins], np.floor((out["x"][ins] - g.xmin) / g.dx).astype(int) + 1)


def test_velocity_field_shifts_frequency():
    m = Model(no_photons=10, temperature=1e4, N_HI=1e18, nx=21, ny=21, nz=21, rmax=1.0, velocity_type="hubble",
              Vexp=100.0, nxfreq=21, xfreq_min=-20, xfreq_max=20).setup()
    one = lambda v: np.array([v], dtype=float)
    out = oracle.raytrace_to_tau(m.config, one(0), one(0), one(0), one(0), one(0), one(1), one(0.0), [11], [11], [11],
                                 one(1e9))
    assert out["inside"][0] == 0
    # escaping along +z from rest at the centre: lab frequency equals the emitted one (comoving frame at v = 0)
    assert out["xfreq_ref"][0] == pytest.approx(0.0, abs=1e-12)


def test_sightline_maps_analytic_centre():
    """Sight-line maps of the oracle (sightline_tau_rect.f90:11-190): through the centre of a static sphere the
    column is 2 N_pole and the line-centre optical depth 2 tau0; outside the projected sphere both vanish."""
    m = Model(no_photons=10, temperature=1e4, taumax=1e3, nx=41, ny=41, nz=41, rmax=1.0, nxfreq=31, nxim=21, nyim=21,
              use_stokes=True, distance=50.0).setup()
    maps, steps = oracle.sightline_tau(m)
    s = m.summary
    tg, ng = maps[0]["tau_gas"], maps[0]["N_gas"]
    assert ng[10, 10] == pytest.approx(2 * s.N_gaspole, rel=1e-12)
    assert tg[15, 10, 10] == pytest.approx(2e3, rel=1e-12)            # bin 16 of 31 is centred on x = 0
    assert np.allclose(tg[::-1], tg, rtol=1e-12, atol=0)                # static medium: symmetric in frequency
    assert np.allclose(ng, ng.T, rtol=1e-9) and np.allclose(ng, ng[::-1, :], rtol=1e-9)  # image symmetry
    assert ng[0, 0] < 0.2 * ng[10, 10] and steps > 0


@pytest.mark.parametrize("nfull,noct", [(32, 16), (29, 15)], ids=["even", "odd_straddling_cell"])
def test_xyz_symmetry_octant_unfolds_to_the_full_grid(nfull, noct):
    """par%xyz_symmetry (raytrace_car.f90:584-760, 1650-1949): a ray mirrored at the lower faces of the octant grid
    sees the same medium as the straight ray in the full grid, and lands at the mirror image of its end point."""
    kw = dict(no_photons=10, temperature=1e4, N_HI=3e14, rmax=1.0, velocity_type="hubble", Vexp=100.0, nxfreq=21)
    full = Model(nx=nfull, ny=nfull, nz=nfull, **kw).setup()
    octa = Model(nx=noct, ny=noct, nz=noct, xyz_symmetry=True, **kw).setup()
    gf, go = full.config.contents.grid, octa.config.contents.grid
    assert go.dx == pytest.approx(gf.dx, rel=1e-15) and go.i0 == (1 if noct % 2 == 0 else 2)
    assert octa.config.contents.par.xyz_symmetry == 1
    rng = np.random.default_rng(11)
    n = 20000
    p = rng.uniform(-0.999, 0.999, (n, 3))
    k = rng.normal(size=(n, 3))
    k /= np.linalg.norm(k, axis=1)[:, None]
    xf = rng.normal(size=n) * 2
    sgn = np.where(p < 0, -1.0, 1.0)
    pf, kf = np.abs(p), k * sgn
    cell = lambda q, g: np.floor((q - [g.xmin, g.ymin, g.zmin]) / [g.dx, g.dy, g.dz]).astype(np.int32) + 1
    ic, io = cell(p, gf), cell(pf, go)
    cols = lambda a: (a[:, 0], a[:, 1], a[:, 2])
    tf, nf, _ = oracle.raytrace_to_edge(full.config, *cols(p), *cols(k), xf, *cols(ic))
    to, no, _ = oracle.raytrace_to_edge(octa.config, *cols(pf), *cols(kf), xf, *cols(io))
    assert np.allclose(to, tf, rtol=1e-9, atol=1e-12)
    assert 0.05 < tf.mean() < 50 and (no != nf).mean() < 0.9
    # the tau walk: same landing point up to the mirror image, same escape decision
    tau_in = rng.exponential(size=n) * np.median(tf)
    a = oracle.raytrace_to_tau(full.config, *cols(p), *cols(k), xf, *cols(ic), tau_in)
    b = oracle.raytrace_to_tau(octa.config, *cols(pf), *cols(kf), xf, *cols(io), tau_in)
    clear = np.abs(tau_in / np.maximum(tf, 1e-300) - 1) > 1e-6  # not on the knife edge between landing and escaping
    assert np.array_equal(a["inside"][clear], b["inside"][clear])
    ins = clear & (a["inside"] == 1)
    assert 0.2 < ins.mean() < 0.8
    for key in "xyz":
        assert np.allclose(np.abs(a[key][ins]), np.abs(b[key][ins]), rtol=0, atol=1e-9), key
    assert np.all(b["x"][ins] >= go.xmin) and np.all(b["z"][ins] >= go.zmin)
    assert np.allclose(a["xfreq"][ins], b["xfreq"][ins], rtol=0, atol=1e-9)
    assert np.allclose(a["xfreq_ref"][clear & ~ins], b["xfreq_ref"][clear & ~ins], rtol=0, atol=1e-9)


def test_xyz_symmetry_run_is_statistically_the_full_run():
    """Whole photon histories in the octant (folded source, reflections, |kz| in Jmu) against the full sphere."""
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    from conftest import small_sphere
    n = 30000
    kw = dict(no_photons=n, taumax=100.0, nxim=0, nyim=0, use_stokes=False, save_Jmu=True, nmu=4, save_all_photons=False)
    full = small_sphere(nx=29, ny=29, nz=29, **kw)
    octa = small_sphere(nx=15, ny=15, nz=15, xyz_symmetry=True, iseed=1234, **kw)
    assert octa.config.contents.par.save_peeloff == 0 and octa.config.contents.par.mu_min == 0.0
    oracle.run(full, rng_mode=1)
    oracle.run(octa, rng_mode=1)
    assert octa.nscatt_gas / n == pytest.approx(full.nscatt_gas / n, rel=0.03)
    a, b = octa.spectrum("Jout"), full.spectrum("Jout")
    sel = a + b > 100
    z = (a[sel] - b[sel]) / np.sqrt(a[sel] + b[sel])
    assert sel.sum() >= 15 and (z ** 2).mean() < 1.8
    jf, jo = full.spectrum("Jmu").sum(0), octa.spectrum("Jmu").sum(0)
    assert np.allclose([jo[0] + jo[1], jo[2] + jo[3]], [jf[1] + jf[2], jf[0] + jf[3]], rtol=0.04)
    assert octa.counters["n_photons_done"] == n


@pytest.mark.parametrize("nfull,nq", [(32, 16), (29, 15)], ids=["even", "odd_straddling_cell"])
def test_xy_symmetry_quadrant_unfolds_to_the_full_grid(nfull, nq):
    """par%xy_symmetry (raytrace_car.f90:783-969, 1951-2250): mirror planes at the lower x and y faces, z open."""
    kw = dict(no_photons=10, temperature=1e4, N_HI=3e14, rmax=1.0, velocity_type="hubble", Vexp=100.0, nxfreq=21)
    full = Model(nx=nfull, ny=nfull, nz=nfull, **kw).setup()
    quad = Model(nx=nq, ny=nq, nz=nfull, xy_symmetry=True, **kw).setup()
    gf, gq = full.config.contents.grid, quad.config.contents.grid
    assert (gq.nx, gq.ny, gq.nz) == (nq, nq, nfull) and gq.dz == gf.dz and gq.zmin == gf.zmin and gq.k0 == 0
    assert gq.dx == pytest.approx(gf.dx, rel=1e-15) and quad.config.contents.par.xy_symmetry == 1
    assert np.allclose(quad.grid_array("rhokap")[gq.i0 - 1, gq.j0 - 1, :], full.grid_array("rhokap")[nfull // 2, nfull // 2, :],
                       rtol=1e-12)
    rng = np.random.default_rng(12)
    n = 20000
    p = rng.uniform(-0.999, 0.999, (n, 3))
    k = rng.normal(size=(n, 3))
    k /= np.linalg.norm(k, axis=1)[:, None]
    xf = rng.normal(size=n) * 2
    sgn = np.where(p < 0, -1.0, 1.0)
    sgn[:, 2] = 1.0
    pf, kf = p * sgn, k * sgn
    cell = lambda q, g: np.floor((q - [g.xmin, g.ymin, g.zmin]) / [g.dx, g.dy, g.dz]).astype(np.int32) + 1
    ic, iq = cell(p, gf), cell(pf, gq)
    cols = lambda a: (a[:, 0], a[:, 1], a[:, 2])
    tf, _, _ = oracle.raytrace_to_edge(full.config, *cols(p), *cols(k), xf, *cols(ic))
    tq, _, _ = oracle.raytrace_to_edge(quad.config, *cols(pf), *cols(kf), xf, *cols(iq))
    assert np.allclose(tq, tf, rtol=1e-9, atol=1e-12) and 0.05 < tf.mean() < 50
    tau_in = rng.exponential(size=n) * np.median(tf)
    a = oracle.raytrace_to_tau(full.config, *cols(p), *cols(k), xf, *cols(ic), tau_in)
    b = oracle.raytrace_to_tau(quad.config, *cols(pf), *cols(kf), xf, *cols(iq), tau_in)
    clear = np.abs(tau_in / np.maximum(tf, 1e-300) - 1) > 1e-6
    assert np.array_equal(a["inside"][clear], b["inside"][clear])
    ins = clear & (a["inside"] == 1)
    assert 0.2 < ins.mean() < 0.8
    for key in "xy":  # a point inside the straddling first cell may stay on the negative side: compare |x|
        assert np.allclose(np.abs(a[key][ins]), np.abs(b[key][ins]), atol=1e-9), key
    assert b["x"][ins].min() >= gq.xmin and b["y"][ins].min() >= gq.ymin
    assert np.allclose(a["z"][ins], b["z"][ins], atol=1e-9) and b["z"][ins].min() < -0.1  # z is not folded
    assert np.allclose(a["xfreq"][ins], b["xfreq"][ins], atol=1e-9)


def test_xy_periodic_box_is_an_infinite_slab():
    """par%xy_periodic with nx, ny > 1 (raytrace_car.f90:971-1136, 2252-2517): rays wrap around in x and y, leave through
    the z faces only, and land folded back into the box."""
    m = Model(no_photons=10, temperature=1e4, taumax=5.0, nx=4, ny=3, nz=20, xmax=0.5, ymax=0.25, zmax=1.0, geometry="rectangle",
              xy_periodic=True, nxfreq=11).setup()
    g = m.config.contents.grid
    assert m.config.contents.par.xy_periodic == 1 and (g.nx, g.ny, g.nz) == (4, 3, 20)
    rng = np.random.default_rng(13)
    n = 20000
    lo, d, nn = np.array([g.xmin, g.ymin, g.zmin]), np.array([g.dx, g.dy, g.dz]), np.array([g.nx, g.ny, g.nz])
    p = lo + rng.uniform(0, 1, (n, 3)) * d * nn
    k = rng.normal(size=(n, 3))
    k /= np.linalg.norm(k, axis=1)[:, None]
    k[np.abs(k[:, 2]) < 0.02, 2] = 0.5  # nearly horizontal rays wrap thousands of times: keep the test quick
    k /= np.linalg.norm(k, axis=1)[:, None]
    q = n // 10
    p[:q, 0] = rng.choice([g.xmin, g.xmax], q)  # on the periodic faces themselves
    p[q:2 * q, 1] = rng.choice([g.ymin, g.ymax], q)
    ic = np.floor((p - lo) / d).astype(np.int32) + 1  # a photon on the upper face sits in cell n+1, as upstream
    xf = rng.uniform(-3, 3, n)
    cols = lambda a: (a[:, 0], a[:, 1], a[:, 2])
    tau, ns, _ = oracle.raytrace_to_edge(m.config, *cols(p), *cols(k), xf, *cols(ic))
    kap = m.grid_array("rhokap")[0, 0, 0] * oracle.voigt(xf, m.summary.voigt_a)
    path = np.where(k[:, 2] > 0, (g.zmax - p[:, 2]) / k[:, 2], (g.zmin - p[:, 2]) / k[:, 2])
    upper = ((ic[:, 0] > g.nx) & (k[:, 0] > 0)) | ((ic[:, 1] > g.ny) & (k[:, 1] > 0))  # to_edge returns 0 there (:1020, :1046)
    assert np.all(tau[upper] == 0) and upper.sum() > 100
    assert np.allclose(tau[~upper], (kap * path)[~upper], rtol=1e-10)
    assert ns.max() > 3 * (g.nx + g.ny + g.nz)  # wrapped more than once
    tau_in = rng.exponential(size=n) * np.median(tau)
    b = oracle.raytrace_to_tau(m.config, *cols(p), *cols(k), xf, *cols(ic), tau_in)
    ins = b["inside"] == 1
    assert np.array_equal(ins, tau_in <= kap * path) or (ins != (tau_in <= kap * path)).mean() < 1e-3
    dist = tau_in / kap
    want = p + dist[:, None] * k
    for a, key, lo_, rng_ in ((0, "x", g.xmin, g.xmax - g.xmin), (1, "y", g.ymin, g.ymax - g.ymin)):
        folded = want[:, a] - np.floor((want[:, a] - lo_) / rng_) * rng_
        err = np.abs(b[key][ins] - folded[ins])
        assert np.all(np.minimum(err, rng_ - err) < 1e-9), key  # modulo one period for points on a face
        assert b[key][ins].min() >= lo_ and b[key][ins].max() <= lo_ + rng_
    assert np.allclose(b["z"][ins], want[ins, 2], atol=1e-9)
    assert (b["icell"][ins] >= 1).all() and (b["icell"][ins] <= g.nx).all()
