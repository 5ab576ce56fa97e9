"""Pins of the CPU oracle (CPU only).

The reference ships no unit-level vectors and cannot be compiled here, so the oracle is
pinned by (a) the whole-run known answers the reference's own logs hold, (b) published
known-answer vectors of the generators, (c) independent mathematics.
"""
import numpy as np
import pytest

from conftest import golden
from lart_b200 import Model
from oracle import oracle


# ---- generators ------------------------------------------------------------
def test_philox4x32_10_known_answers():
    # Random123 kat_vectors (philox4x32-10)
    assert oracle.philox_raw([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox_raw([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox_raw([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_mt19937_64_known_answers():
    # std::mt19937_64 with the default seed 5489: first word and the 10000th word (C++ standard)
    w = oracle.mt64_words(5489, 10000)
    assert w[0] == 14514284786278117030
    assert w[9999] == 9981545732273789042


def test_uniform_is_open_interval_and_52_bit():
    u = oracle.sample(0, 123, np.arange(1000), ndraw=64)
    assert u.min() > 0.0 and u.max() < 1.0
    k = u * 2.0 ** 52 - 0.5  # ((w>>12)+0.5)*2^-52 — random_mt.f90:628-629
    assert np.array_equal(k, np.round(k))
    assert abs(u.mean() - 0.5) < 5e-3


# ---- Voigt -----------------------------------------------------------------
def test_voigt_against_faddeeva():
    from scipy.special import wofz
    x = np.linspace(0, 30, 6001)
    for a in (4.7186e-4, 1.4921e-2):
        H = oracle.voigt(x, a)
        exact = wofz(x + 1j * a).real
        rel = np.abs(H / exact - 1)
        assert rel.max() < 2.5e-3, (a, rel.max())  # the reference's approximation, SURVEY A8
        assert np.array_equal(oracle.voigt(-x, a), H)  # even in x


def test_voigt_nodes_and_limits():
    a = 4.7186e-4
    # x = 0: h0(1) + a*(h1(1) + a*h2(1)) with h1(1) = -2/sqrt(pi)
    assert oracle.voigt([0.0], a)[0] == pytest.approx(1 - 2 * a / np.sqrt(np.pi) + a * a, rel=1e-9)
    # asymptote beyond |x| = 10
    x = 50.0
    assert oracle.voigt([x], a)[0] == pytest.approx(a / np.sqrt(np.pi) / x ** 2 * (1 + 1.5 / x ** 2), rel=1e-5)
    # continuity across the branch boundaries 1, 5, 10 (approximation error only)
    for xb in (1.0, 5.0, 10.0):
        lo, hi = oracle.voigt([np.nextafter(xb, 0)], a)[0], oracle.voigt([xb], a)[0]
        assert abs(lo / hi - 1) < 5e-3


# ---- set-up scalars printed by the reference -------------------------------
def test_setup_scalars_sphere_peel_log():
    # examples/sphere_peel/out.txt:13-17 (t1tau3.in: T = 10 K, tau0 = 1e3, 201^3)
    m = Model(no_photons=10, temperature=10.0, taumax=1e3, use_stokes=True, nx=201, ny=201, nz=201, rmax=1.0,
              nxfreq=201, nxim=129, nyim=129).setup()
    s = m.summary
    assert float("%.3e" % s.voigt_a) == golden("sphere_peel_t1tau3", "voigt_a") == 1.492e-2
    assert float("%.3e" % s.N_gaspole) == golden("sphere_peel_t1tau3", "N_HI_pole") == 5.449e14
    assert s.taupole == pytest.approx(golden("sphere_peel_t1tau3", "tau_pole"), rel=1e-12)
    assert (s.nobs, s.nxim, s.nyim) == (1, 129, 129)
    assert s.dxim == pytest.approx(np.degrees(np.arcsin(1.0 / 100.0)) / 64.5, rel=1e-14)  # observer_rect.f90:248


def test_setup_scalars_amr_sphere_log():
    # examples/amr_sphere_generic/log_car_1M.txt:9-13 (64^3, T = 1e4 K, tau0 = 1e4)
    m = Model(no_photons=10, temperature=1e4, taumax=1e4, nx=64, ny=64, nz=64, rmax=1.0, nxfreq=121).setup()
    s = m.summary
    assert float("%.3e" % s.voigt_a) == golden("amr_sphere_generic_car_1M", "voigt_a") == 4.719e-4
    assert abs(s.voigt_a / 4.7186e-4 - 1) < 2e-5  # log_amr_1M.txt:12
    assert float("%.3e" % s.N_gaspole) == golden("amr_sphere_generic_car_1M", "N_HI_pole") == 1.695e17


def test_namelist_file_roundtrip(tmp_path):
    f = tmp_path / "t4tau4.in"
    f.write_text("""&parameters
 par%no_photons  = 1e5
 par%temperature = 1.0e4
 par%taumax      = 1.0d4     ! comment
 par%use_stokes  = .true.
 par%spectral_type = 'voigt'
 par%xy_periodic = .true.
 par%nx = 1
 par%ny = 1
 par%nz = 201
 par%nprint = 1000000
/
""")
    m = Model(str(f)).setup()
    s = m.summary
    assert (s.nx, s.ny, s.nz, s.zonly, s.nxfreq, s.nphotons) == (1, 1, 201, 1, 121, 100000)
    assert s.taupole == pytest.approx(1e4, rel=1e-12)
    from lart_b200 import LartError
    with pytest.raises(LartError):
        Model(no_such_key=1.0)
    with pytest.raises(LartError):
        Model(geometry="healpix_shell").setup()  # stays with the Fortran host
    assert Model(z_symmetry=True, nx=4, ny=4, nz=4).setup().config.contents.grid.zmin == 0.0  # geometry only, as upstream
    with pytest.raises(LartError):
        Model(xy_symmetry=True, xy_periodic=True, nx=3, ny=3).setup()
    box = Model(xy_periodic=True, nx=3, ny=3, geometry="rectangle").setup()  # 3-D periodic box: the _xyper ray tracers
    assert box.summary.zonly == 0 and box.config.contents.par.xy_periodic == 1


# ---- whole-run known answers ------------------------------------------------
def test_mean_scatterings_sphere_peel_log():
    # <N_scatt> = 1.7898e3 — examples/sphere_peel/out.txt:33
    n = 6000
    m = Model(no_photons=n, temperature=10.0, taumax=1e3, use_stokes=True, nx=201, ny=201, nz=201, rmax=1.0,
              nxfreq=201, nxim=129, nyim=129, save_all_photons=True).setup()
    oracle.run(m, rng_mode=0, seed=20240611)
    ns = m.allph("nscatt_gas")
    mean, err = ns.mean(), ns.std() / np.sqrt(n)
    assert golden("sphere_peel_t1tau3", "mean_nscatt") == 1.7898e3
    assert abs(mean - golden("sphere_peel_t1tau3", "mean_nscatt")) < 4 * err, (mean, err)
    assert m.nscatt_gas / n == pytest.approx(mean, rel=1e-12)
    # python/check_flux.py:45-52 — the peel cube integrates to ~1 (no dust)
    m.output_normalize()
    s = m.summary
    omega = s.dxim * s.dyim * (np.pi / 180) ** 2
    total = (m.observer_cube("scatt").sum() + m.observer_cube("direc").sum()) * 4 * np.pi * omega * s.distance ** 2 * s.dxfreq
    assert total == pytest.approx(1.0, abs=0.05)


def test_mean_scatterings_amr_sphere_log():
    # <N_scatt> = 2.8225e4 — examples/amr_sphere_generic/log_car_1M.txt:24
    n = 1500
    m = Model(no_photons=n, temperature=1e4, taumax=1e4, nx=64, ny=64, nz=64, rmax=1.0, nxfreq=121,
              save_all_photons=True).setup()
    oracle.run(m, rng_mode=0, seed=5)
    ns = m.allph("nscatt_gas")
    mean, err = ns.mean(), ns.std() / np.sqrt(n)
    assert golden("amr_sphere_generic_car_1M", "mean_nscatt") == 2.8225e4
    assert abs(mean - golden("amr_sphere_generic_car_1M", "mean_nscatt")) < 4 * err, (mean, err)


# ---- analytic anchors --------------------------------------------------------
def neufeld_slab(x, a, tau0):
    """Harrington (1973) / Neufeld (1990) slab with a central source; integral = 1/(4 pi)."""
    return np.sqrt(6.0) / (24.0 * np.sqrt(np.pi) * a * tau0) * x ** 2 / np.cosh(np.sqrt(np.pi ** 3 / 54.0) * np.abs(x ** 3) / (a * tau0))


def test_neufeld_formula_normalisation():
    x = np.linspace(-200, 200, 400001)
    assert np.trapezoid(neufeld_slab(x, 1.5e-2, 1e5), x) == pytest.approx(1 / (4 * np.pi), rel=1e-6)


def test_slab_against_neufeld():
    # T = 10 K, tau0 = 1e5: a*tau0 = 1.5e3 — inside the validity range of the analytic solution
    n = 400
    m = Model(no_photons=n, temperature=10.0, taumax=1e5, xy_periodic=True, nx=1, ny=1, nz=201, nxfreq=60,
              xfreq_min=-30.0, xfreq_max=30.0, spectral_type="monochromatic").setup()
    oracle.run(m, rng_mode=0, seed=11)
    m.output_normalize()
    x, J = m.xfreq(), m.spectrum("Jout")
    s = m.summary
    assert J.sum() * s.dxfreq == pytest.approx(1 / (4 * np.pi), rel=1e-9)  # every photon escapes inside the window
    # first moment of |x|: Monte Carlo vs analytic, within MC error (~ 1/sqrt(n)) + the finite-a*tau0 offset
    xa = np.linspace(-30, 30, 6001)
    mean_ana = np.trapezoid(np.abs(xa) * neufeld_slab(xa, s.voigt_a, 1e5), xa) * 4 * np.pi
    mean_mc = (np.abs(x) * J).sum() * s.dxfreq * 4 * np.pi
    assert mean_mc == pytest.approx(mean_ana, rel=0.06)


# ---- samplers: distributions -------------------------------------------------
def test_rand_resonance_vz_distribution():
    # u_par ~ exp(-u^2)/((x-u)^2+a^2): compare the sample mean/var with quadrature
    a = 4.7186e-4
    for x0 in (0.0, 0.5, 1.5, 2.4142, 3.0, 5.0, -4.0, 8.0):
        v = oracle.sample(2, 99, np.arange(20000), p0=x0, p1=a, ndraw=1, rng_mode=1).ravel()
        u = np.linspace(-8, 12, 2000001)
        w = np.exp(-u * u) / ((x0 - u) ** 2 + a * a)
        mean = np.trapezoid(u * w, u) / np.trapezoid(w, u)
        var = np.trapezoid((u - mean) ** 2 * w, u) / np.trapezoid(w, u)
        assert abs(v.mean() - mean) < 5 * np.sqrt(var / v.size) + 1e-3, (x0, v.mean(), mean)


def test_rand_resonance_phase_function():
    # P(mu) ~ 1 + mu^2 for E1 = 1 (Rayleigh); isotropic for E1 = 0
    mu = oracle.sample(3, 3, np.arange(100000), p0=1.0).ravel()
    assert abs((mu ** 2).mean() - 0.4) < 5e-3
    mu = oracle.sample(3, 3, np.arange(100000), p0=0.0).ravel()
    assert abs((mu ** 2).mean() - 1 / 3) < 5e-3
    g = 0.6761
    mu = oracle.sample(4, 3, np.arange(100000), p0=g).ravel()
    assert abs(mu.mean() - g) < 5e-3
    z = oracle.sample(1, 3, np.arange(100000)).ravel()
    assert abs(z.mean()) < 0.01 and abs(z.var() - 1) < 0.02


def test_continuum_runs_are_normalised_to_the_input_continuum_level():
    """output_normalize_outside, continuum branch (output_sum_rect.f90:252-273; par%continuum_normalize defaults to
    .true., define.f90:314): every spectrum is divided by mean(Jin), so the injected continuum sits at 1; without
    save_Jin the reference aborts."""
    from lart_b200 import LartError, Model
    kw = dict(no_photons=4000, temperature=1e4, taumax=5.0, nx=11, ny=11, nz=11, rmax=1.0, spectral_type="continuum",
              nxfreq=40, xfreq_min=-20.0, xfreq_max=20.0, save_Jmu=True, nmu=3, iseed=3)
    m = Model(**kw).setup()
    oracle.run(m, rng_mode=1, nthreads=2)
    m.output_normalize()
    assert m.spectrum("Jin").mean() == pytest.approx(1.0, rel=1e-12)
    assert m.spectrum("Jout").mean() == pytest.approx(1.0, rel=0.05)  # no dust: what goes in comes out
    assert m.spectrum("Jmu").sum(1).mean() / 3 == pytest.approx(m.spectrum("Jout").mean(), rel=1e-9)
    raw = Model(continuum_normalize=False, **kw).setup()
    oracle.run(raw, rng_mode=1, nthreads=2)
    raw.output_normalize()
    assert raw.spectrum("Jin").mean() != pytest.approx(1.0, rel=0.5)
    assert np.allclose(raw.spectrum("Jout") / raw.spectrum("Jin").mean(), m.spectrum("Jout"), rtol=1e-12)
    bad = Model(save_Jin=False, **kw).setup()
    oracle.run(bad, rng_mode=1, nthreads=2)
    with pytest.raises(LartError, match="save_Jin"):
        bad.output_normalize()
