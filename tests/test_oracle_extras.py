"""CPU pins of the less common Cartesian bindings and the CALCJ / CALCP / CALCPnew accumulators (8f-4) in
the oracle: properties that hold whatever the random stream, checked on small runs."""
import numpy as np
import pytest

from conftest import small_sphere
from lart_b200 import LartError, Model
from oracle import oracle

ATM_PLANE = dict(geometry="plane_atmosphere", source_geometry="plane_illumination", nz=40, zmax=1.0, rmax=-999.0, taumax=30.0,
                 density_zscale=0.4, nxim=0, nyim=0, spectral_type="voigt", no_photons=1500)
ATM_SPH = dict(geometry="spherical_atmosphere", source_geometry="plane_illumination", nx=21, ny=21, nz=21, rmax=1.0, rmin=0.45,
               density_rscale=0.5, taumax=20.0, spectral_type="gaussian", no_photons=1500)
SHEAR = dict(xy_periodic=True, geometry="rectangle", rmax=-999.0, nx=6, ny=5, nz=31, xmax=0.5, ymax=0.4, zmax=1.0, taumax=40.0,
             Omega=1.5, xs_point=0.2, ys_point=-0.1, nxim=0, nyim=0)


def test_accumulators_radial_and_cylindrical_bins_are_sums_of_the_cell_maps():
    # one run per geometry_JPa on the same Philox streams: every photon takes the same path, only the binning differs
    kw = dict(calc_J=True, calc_P=True, calc_Pnew=True, no_photons=400, nxim=0, nyim=0, nx=15, ny=15, nz=15, taumax=30.0)
    m3 = small_sphere(geometry_JPa=3, **kw)
    m1 = small_sphere(geometry_JPa=1, **kw)
    m2 = small_sphere(geometry_JPa=2, **kw)
    for m in (m3, m1, m2):
        oracle.run(m, rng_mode=1, nthreads=4)
    g1, g2 = m1.config.contents.grid, m2.config.contents.grid
    ind_sph = np.ctypeslib.as_array(g1.ind_sph, shape=(15 ** 3,)).reshape((15, 15, 15), order="F")
    ind_cyl = np.ctypeslib.as_array(g2.ind_cyl, shape=(15 * 15,)).reshape((15, 15), order="F")
    J3, P3, N3 = m3.jp_array("J"), m3.jp_array("Pa"), m3.jp_array("Pnew")
    assert J3.shape == (61, 15, 15, 15) and P3.sum() > 0 and N3.sum() > 0 and J3.sum() > 0
    for ir in range(1, g1.nr + 1):
        sel = ind_sph == ir
        assert m1.jp_array("Pa")[ir - 1] == pytest.approx(P3[sel].sum(), rel=1e-12, abs=1e-300)
        assert m1.jp_array("Pnew")[ir - 1] == pytest.approx(N3[sel].sum(), rel=1e-12, abs=1e-300)
        np.testing.assert_allclose(m1.jp_array("J")[:, ir - 1], J3[:, sel].sum(axis=1), rtol=1e-12, atol=1e-300)
    for ir in range(1, g2.nr + 1):
        sel = ind_cyl == ir
        np.testing.assert_allclose(m2.jp_array("Pa")[ir - 1, :], P3[sel, :].sum(axis=0), rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(m2.jp_array("J")[:, ir - 1, :], J3[:, sel, :].sum(axis=1), rtol=1e-12, atol=1e-300)


def test_scattering_rate_estimators_agree_and_count_every_scattering():
    # CALCP counts scatterings, CALCPnew integrates the line optical depth along the paths: the same expectation value.
    m = small_sphere(calc_P=True, calc_Pnew=True, geometry_JPa=1, no_photons=4000, nxim=0, nyim=0, taumax=50.0, use_stokes=False)
    oracle.run(m, rng_mode=0, nthreads=4)
    g = m.config.contents.grid
    rho = m.grid_array("rhokap") * m.grid_array("Dfreq") / m.summary.cross0
    dens = rho[rho > 0][0]  # uniform sphere
    Pa, Pn = m.jp_array("Pa") * dens, m.jp_array("Pnew") * dens  # back to weighted event counts per radial bin
    assert Pa.sum() == pytest.approx(m.nscatt_gas, rel=1e-9)  # every resonance scattering was counted once
    assert Pn.sum() == pytest.approx(Pa.sum(), rel=0.03)
    big = Pa > 2000
    assert big.sum() >= 5 and np.all(np.abs(Pn[big] / Pa[big] - 1) < 0.12)


def test_mean_intensity_integrates_to_the_path_length_scattering_rate():
    # add_to_J bins del*wgt by frequency, add_to_Pnew sums del*wgt*rhokap*H(x)/n: with fine frequency bins
    # sum_x J(x, r) * rhokap * H(x) reproduces Pnew(r) * n — the two accumulators see the same segments
    m = small_sphere(calc_J=True, calc_Pnew=True, geometry_JPa=1, no_photons=1500, nxim=0, nyim=0, taumax=30.0, nx=15, ny=15, nz=15,
                     nxfreq=4001, xfreq_min=-20.0, xfreq_max=20.0)
    oracle.run(m, rng_mode=1, nthreads=4)
    J, Pn = m.jp_array("J"), m.jp_array("Pnew")
    rk, D, a = m.grid_array("rhokap"), m.grid_array("Dfreq"), m.grid_array("voigt_a")
    inside = rk > 0
    rhokap, dens, va = rk[inside][0], (rk * D / m.summary.cross0)[inside][0], a[inside][0]
    H = oracle.voigt(m.xfreq(), va)
    lhs = (J * H[:, None]).sum(axis=0) * rhokap
    rhs = Pn * dens
    assert rhs.sum() > 100.0
    assert lhs.sum() == pytest.approx(rhs.sum(), rel=2e-3)
    big = rhs > 0.05 * rhs.max()
    np.testing.assert_allclose(lhs[big], rhs[big], rtol=5e-3)
    m.output_normalize()
    assert np.isfinite(m.jp_array("J")).all() and np.isfinite(m.jp_array("Pnew")).all()


def test_plane_atmosphere_conserves_photons_between_Jout_and_Jabs2():
    m = Model(**ATM_PLANE).setup()
    c = m.config.contents
    assert c.par.atmosphere == 1 and c.par.source_geometry == 3 and c.grid.nx == 1 and c.par.xy_periodic == 1
    oracle.run(m, rng_mode=1, nthreads=4)
    out, ab2 = m.spectrum("Jout").sum(), m.spectrum("Jabs2").sum()
    assert ab2 > 0 and out > 0   # some photons reach the bottom and are absorbed there, some are reflected
    assert out + ab2 == pytest.approx(1500.0, rel=1e-9)
    # illumination from the top, straight down: every photon starts in the top cell at z = zmax
    assert m.spectrum("Jin").sum() == pytest.approx(1500.0)


def test_spherical_atmosphere_mask_absorbs_and_blocks():
    m = Model(**ATM_SPH, nxim=9, nyim=9).setup()
    c = m.config.contents
    mask = np.ctypeslib.as_array(c.grid.mask, shape=(21 ** 3,)).reshape((21, 21, 21), order="F")
    assert mask.min() == -1 and (mask == -1).sum() > 30
    oracle.run(m, rng_mode=1, nthreads=4)
    out, ab2 = m.spectrum("Jout").sum(), m.spectrum("Jabs2").sum()
    assert ab2 > 30.0 and out > 0.1 * 1500
    assert out + ab2 == pytest.approx(1500.0, rel=1e-9)
    # a ray aimed at the planet has tau = +inf (raytrace_to_edge_car_atmosphere), one that misses it a finite tau
    z0 = np.array([-0.99, -0.99]); x0 = np.array([0.0, 0.9]); zero = np.zeros(2)
    ic = np.floor((x0 - c.grid.xmin) / c.grid.dx).astype(np.int32) + 1
    jc = np.floor((zero - c.grid.ymin) / c.grid.dy).astype(np.int32) + 1
    kc = np.floor((z0 - c.grid.zmin) / c.grid.dz).astype(np.int32) + 1
    tau, ns, _ = oracle.raytrace_to_edge(m.config, x0, zero, z0, zero, zero, zero + 1.0, zero, ic, jc, kc)
    assert np.isinf(tau[0]) and np.isfinite(tau[1]) and tau[1] > 0
    maps, _ = oracle.sightline_tau(m)
    assert np.isinf(maps[0]["N_gas"]).any() and np.isfinite(maps[0]["N_gas"]).any()


def test_shear_box_changes_frequencies_only_through_x_wraps():
    a, b = Model(**SHEAR).setup(), Model(**dict(SHEAR, Omega=0.0)).setup()
    assert a.config.contents.par.Omega == 1.5
    for m in (a, b):
        oracle.run(m, rng_mode=1, nthreads=4, count=600)
    assert a.counters["n_photons_done"] == b.counters["n_photons_done"] == 600
    assert a.spectrum("Jout").sum() == pytest.approx(600.0, rel=0.02)  # (the frequency window holds nearly all of them)
    assert np.abs(a.spectrum("Jout") - b.spectrum("Jout")).sum() > 10.0  # the shear does move the spectrum
    with pytest.raises(LartError):
        Model(**dict(SHEAR, nx=1, ny=1)).setup()


def test_z_symmetry_is_geometry_only():
    """par%z_symmetry halves the grid in z (grid_mod_car.f90:135-150) and binds no ray tracer of its own (setup.f90:947-987):
    zmin = 0 for an even nz, -dz/2 for an odd one; photons that reach the cut plane leave the grid."""
    even = small_sphere(z_symmetry=True, nx=16, ny=16, nz=16, zs_point=0.3, nxim=0, nyim=0, no_photons=500)
    odd = small_sphere(z_symmetry=True, nx=15, ny=15, nz=15, nxim=0, nyim=0, no_photons=500)
    ge, go = even.config.contents.grid, odd.config.contents.grid
    assert ge.zmin == 0.0 and ge.dz == pytest.approx(1.0 / 16) and ge.k0 == 1 and ge.xmin == -1.0 and ge.i0 == 0
    assert go.dz == pytest.approx(1.0 / 14.5) and go.zmin == pytest.approx(-go.dz / 2) and go.k0 == 2
    for m in (even, odd):
        c = m.config.contents.par
        assert c.xyz_symmetry == 0 and c.xy_symmetry == 0 and c.xy_periodic == 0
        oracle.run(m, rng_mode=1, nthreads=4)
        assert m.counters["n_photons_done"] == 500
    # the odd grid's source sits inside the straddling first cell: half of the first flights leave through the cut at once
    assert 0.2 * 500 < odd.spectrum("Jout").sum() <= 500.0
