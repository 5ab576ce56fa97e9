"""Octree AMR path on the CPU (no GPU): the mini-host's grid_create_amr against the reference's own log, and the oracle's
restatement of the octree ray tracers against independent checks.  The leaf lists are written by the REFERENCE's grid generator
(python/AMR_grid/AMR_grid.py via tools/make_amr_fixtures.py); the full-size one has the 178 480 leaves the reference logs."""
import numpy as np
import pytest

from conftest import amr_sphere, golden
from oracle import oracle


def test_amr_setup_reproduces_the_reference_log():
    # examples/amr_sphere_generic/log_amr_1M.txt: nleaf 178480, voigt_a 4.7186e-4, N(HI)_pole 1.6954e17, tau_pole 1e4
    m = amr_sphere("amr_sphere_l37", taumax=1e4, nxfreq=121, nxim=0, nyim=0)
    a, s = m.config.contents.amr, m.summary
    g = lambda k: golden("amr_sphere_generic_amr_1M", k)
    assert a.nleaf == g("nleaf")
    assert s.voigt_a == pytest.approx(g("voigt_a"), rel=2e-5)
    assert s.N_gaspole == pytest.approx(g("N_HI_pole"), rel=5e-5)
    assert s.taupole == pytest.approx(g("tau_pole"), rel=1e-12)
    # the same column as the Cartesian twin's log (log_car_1M.txt: 1.695e17), and the same frequency grid
    assert s.N_gaspole == pytest.approx(golden("amr_sphere_generic_car_1M", "N_HI_pole"), rel=1e-3)
    assert (s.xfreq_min, s.xfreq_max, s.nxfreq) == (-9.0, 9.0, 121)


def test_octree_topology():
    m = amr_sphere("amr_sphere_l25")
    a = m.config.contents.amr
    nc, nl = a.ncells, a.nleaf
    children = np.ctypeslib.as_array(a.children, shape=(nc * 8,)).reshape(nc, 8)
    ileaf = np.ctypeslib.as_array(a.ileaf, shape=(nc,))
    nb = np.ctypeslib.as_array(a.neighbor, shape=(nc * 6,)).reshape(nc, 6)
    cx, cy, cz, ch = (np.ctypeslib.as_array(getattr(a, k), shape=(nc,)) for k in ("cx", "cy", "cz", "ch"))
    icl = np.ctypeslib.as_array(a.icell_of_leaf, shape=(nl,))
    assert np.array_equal(ileaf[icl - 1], np.arange(1, nl + 1))
    assert ((ileaf > 0) == (children.max(axis=1) == 0)).all()              # leaves have no children, internal cells have some
    assert np.isclose((8.0 * ch[icl - 1] ** 3).sum(), 8.0)                   # the leaves tile the box (no gaps in this grid)
    # a same-level neighbour sits exactly one cell width away; coarser neighbours contain that point
    for f, (ax, sg) in enumerate([(cx, 1), (cx, -1), (cy, 1), (cy, -1), (cz, 1), (cz, -1)]):
        has = nb[:, f] > 0
        n = nb[has, f] - 1
        d = ax[n] - ax[has]
        same = np.isclose(ch[n], ch[has])
        assert np.allclose(d[same], sg * 2.0 * ch[has][same])
        assert (ch[n] >= ch[has] - 1e-15).all()                              # never finer than the cell itself
        edge = ~has
        lim = 1.0 if sg > 0 else -1.0
        assert np.allclose(ax[edge] + sg * ch[edge], lim)                     # no neighbour only at the box faces


def test_tracers_against_direct_integration():
    """tau of raytrace_to_edge_amr = the integral of kappa along the ray, evaluated independently by locating many sample
    points (amr_find_leaf) — uniform sphere, line centre, below the tau cap."""
    m = amr_sphere("amr_sphere_l25", taumax=50.0)
    cfg, a = m.config, m.config.contents.amr
    rng = np.random.default_rng(1)
    n = 300
    p = rng.uniform(-0.6, 0.6, (n, 3))
    k = rng.normal(size=(n, 3)); k /= np.linalg.norm(k, axis=1)[:, None]
    xf = np.zeros(n)
    il = oracle.amr_locate(cfg, p[:, 0], p[:, 1], p[:, 2])
    assert (il > 0).all()
    tau, ns = oracle.amr_edge(cfg, p[:, 0], p[:, 1], p[:, 2], k[:, 0], k[:, 1], k[:, 2], xf, il)
    rk = np.ctypeslib.as_array(a.rhokap, shape=(a.nleaf,))
    va = np.ctypeslib.as_array(a.voigt_a, shape=(a.nleaf,))
    H0 = oracle.voigt(np.zeros(1), va[:1])[0]
    t = (np.arange(4000) + 0.5) / 4000 * 4.0                                  # midpoints along 4 box half-widths
    for i in range(0, n, 10):
        q = p[i][None, :] + t[:, None] * k[i][None, :]
        leaf = oracle.amr_locate(cfg, q[:, 0], q[:, 1], q[:, 2])
        kap = np.where(leaf > 0, rk[np.maximum(leaf, 1) - 1], 0.0) * H0
        assert tau[i] == pytest.approx(kap.sum() * (4.0 / 4000), rel=3e-3)
    # from the centre along +z the line-centre depth is tau_pole by construction (grid_mod_amr.f90:343-430)
    t0, _ = oracle.amr_edge(cfg, [0.0], [0.0], [0.0], [0.0], [0.0], [1.0], [0.0], [0])
    assert t0[0] == pytest.approx(50.0, rel=1e-12)
    # to_tau stops where to_edge has accumulated tau_in, in the leaf that contains the end point
    tin = 0.5 * tau
    out = oracle.amr_tau(cfg, p[:, 0], p[:, 1], p[:, 2], k[:, 0], k[:, 1], k[:, 2], xf, il, tin)
    assert (out["inside"] == 1).all()
    t2, _ = oracle.amr_edge(cfg, out["x"], out["y"], out["z"], k[:, 0], k[:, 1], k[:, 2], out["xfreq"], out["il"])
    assert np.allclose(t2, tau - tin, rtol=1e-9)
    back = oracle.amr_locate(cfg, out["x"], out["y"], out["z"])
    d = np.abs(np.stack([out["x"], out["y"], out["z"]], 1))
    interior = (np.abs(d * 32 - np.rint(d * 32)) > 1e-9).all(axis=1)          # not on a face of the finest level
    assert np.array_equal(back[interior], out["il"][interior])


def test_amr_run_known_answer_and_cartesian_twin():
    """<N_scatt> of the octree sphere = the reference's logged 2.8225e4 (log_amr_1M.txt) within Monte-Carlo error, and the
    escaping spectrum agrees with the Cartesian 64^3 twin (same tau0, same frequency grid) bin by bin."""
    n = 3000
    ma = amr_sphere("amr_sphere_l37", no_photons=n, taumax=1e4, nxfreq=121, nxim=0, nyim=0, use_stokes=False)
    oracle.run(ma, rng_mode=0, seed=21)
    ns = ma.allph("nscatt_gas")
    assert abs(ns.mean() - golden("amr_sphere_generic_amr_1M", "mean_nscatt")) < 4 * ns.std() / np.sqrt(n)
    from lart_b200 import Model
    mc = Model(no_photons=n, temperature=1e4, taumax=1e4, nx=64, ny=64, nz=64, rmax=1.0, nxfreq=121, use_stokes=False, iseed=9).setup()
    oracle.run(mc, rng_mode=0, seed=22)
    a, b = ma.spectrum("Jout"), mc.spectrum("Jout")
    sel = (a + b) >= 60
    z = (a[sel] - b[sel]) / np.sqrt(a[sel] + b[sel])
    assert sel.sum() >= 20 and (z ** 2).sum() / sel.sum() < 1 + 5 * np.sqrt(2.0 / sel.sum())
