"""The less common Cartesian bindings (shearing box, plane and spherical atmospheres — setup.f90:959-987,
raytrace_car.f90:2677-3976) and the CALCJ / CALCP / CALCPnew accumulators (8f-4: raytrace_car.f90:3979-4045,
scattering_car.f90:829-860) on the GPU against the oracle, same Philox streams, through the C ABI."""
import numpy as np
import pytest

from conftest import small_sphere
from lart_b200 import LartError, Model, Simulation, capi
from oracle import oracle
from test_gpu_runs import histories_equal, run_gpu, tallies_close
from test_oracle_extras import ATM_PLANE, ATM_SPH, SHEAR

pytestmark = pytest.mark.gpu

ALLPH = dict(save_all_photons=True, iseed=7)
CASES = {
    # accumulators, one case per par%geometry_JPa
    "calc_sphere_cells": (small_sphere, dict(calc_J=True, calc_P=True, calc_Pnew=True, geometry_JPa=3, nx=15, ny=15, nz=15, taumax=30.0,
                                             no_photons=800)),
    "calc_sphere_radial_coreskip": (small_sphere, dict(calc_J=True, calc_P=True, calc_Pnew=True, geometry_JPa=1, taumax=1e3, core_skip=True,
                                                       no_photons=500, nxim=0, nyim=0)),
    "calc_cylindrical_dust": (small_sphere, dict(calc_J=True, calc_P=True, calc_Pnew=True, geometry_JPa=2, use_stokes=False, DGR=1.0,
                                                 cext_dust=3e-17, taumax=-999.0, N_HI=2e16, no_photons=800, nxim=0, nyim=0)),
    "calc_slab_planes": (small_sphere, dict(calc_J=True, calc_P=True, calc_Pnew=True, xy_periodic=True, nx=1, ny=1, nz=101, rmax=-999.0,
                                            taumax=200.0, nxim=0, nyim=0, nxfreq=121, no_photons=800)),
    "calc_octant": (small_sphere, dict(calc_J=True, calc_Pnew=True, xyz_symmetry=True, nx=16, ny=15, nz=16, nxim=0, nyim=0, taumax=40.0,
                                       no_photons=600)),
    "calc_periodic_box_hubble": (small_sphere, dict(calc_J=True, calc_P=True, xy_periodic=True, geometry="rectangle", rmax=-999.0, nx=5, ny=4,
                                                    nz=41, xmax=0.5, ymax=0.4, zmax=1.0, taumax=50.0, nxim=0, nyim=0,
                                                    velocity_type="parallel_velocity", Vx=20.0, Vy=-10.0, Vz=5.0, no_photons=600)),
    # atmospheres
    "plane_atmosphere": (Model, dict(ATM_PLANE, **ALLPH)),
    "plane_atmosphere_stokes_calc": (Model, dict(ATM_PLANE, use_stokes=True, calc_J=True, calc_P=True, calc_Pnew=True, **ALLPH)),
    "spherical_atmosphere_peel": (Model, dict(ATM_SPH, nxim=9, nyim=9, use_stokes=True, **ALLPH)),
    "spherical_atmosphere_xysym_calc": (Model, dict(ATM_SPH, xy_symmetry=True, nx=12, ny=12, nz=23, calc_Pnew=True, calc_J=True, **ALLPH)),
    # par%z_symmetry: grid geometry only (grid_mod_car.f90:135-150), the plain open-box routines run on the upper half
    "z_symmetry_half_box_even": (small_sphere, dict(z_symmetry=True, nx=16, ny=16, nz=16, zs_point=0.3, taumax=30.0, no_photons=800)),
    "z_symmetry_half_box_odd_source_on_the_cut": (small_sphere, dict(z_symmetry=True, nx=15, ny=15, nz=15, taumax=30.0, no_photons=800,
                                                                    use_stokes=False)),
    # shearing box
    "shear_box": (Model, dict(SHEAR, no_photons=1200, **ALLPH)),
    "shear_box_peel_velocity": (Model, dict(SHEAR, no_photons=800, nxim=9, nyim=9, velocity_type="parallel_velocity", Vy=15.0, Vx=5.0,
                                            use_stokes=True, **ALLPH)),
}


def extras_close(mg, mo, same_frac):
    tol = 4 * (1 - same_frac) + 1e-9
    a, b = mg.spectrum("Jabs2"), mo.spectrum("Jabs2")
    assert (a is None) == (b is None)
    if a is not None:
        assert np.abs(a - b).sum() <= 2 * tol * max(b.sum(), 1.0) + 1e-9
    for name in ("J", "Pa", "Pnew"):
        a, b = mg.jp_array(name), mo.jp_array(name)
        assert (a is None) == (b is None), name
        if a is None:
            continue
        assert a.shape == b.shape and b.sum() > 0, name
        assert np.abs(a - b).sum() <= (2 * tol + 1e-9) * b.sum(), (name, np.abs(a - b).sum() / b.sum())


@pytest.mark.parametrize("flags", [0, capi.FLAG_MONOLITHIC], ids=["wavefront", "monolithic"])
@pytest.mark.parametrize("case", sorted(CASES))
def test_extras_match_oracle(case, flags):
    make, kw = CASES[case]
    mg, mo = make(**kw), make(**kw)
    if make is Model:
        mg.setup(); mo.setup()
    # small pools, short step budgets: walks are parked and resumed, slots are reused many times
    run_gpu(mg, flags=flags, pool_slots=1024, ray_budget=4)
    oracle.run(mo, rng_mode=1)
    vel = "velocity" in case or "hubble" in case
    same = histories_equal(mg, mo, geom_rtol=5e-3 if vel else 1e-8)
    tallies_close(mg, mo, min(same.mean(), 0.998) if vel else same.mean())
    extras_close(mg, mo, min(same.mean(), 0.998) if vel else same.mean())
    assert mg.counters["n_photons_done"] == mo.config.contents.par.nphotons


def test_atmosphere_sightline_maps_equal_oracle():
    m = Model(**ATM_SPH, nxim=9, nyim=9).setup()
    sim = Simulation(m)
    maps = sim.sightline_tau()
    ref, _ = oracle.sightline_tau(m)
    sim.close()
    assert np.isinf(ref[0]["N_gas"]).any() and np.isinf(ref[0]["tau_gas"]).any()
    np.testing.assert_array_equal(maps[0]["N_gas"], ref[0]["N_gas"])
    np.testing.assert_array_equal(maps[0]["tau_gas"], ref[0]["tau_gas"])


def test_edge_walks_of_shear_and_atmosphere_runs_are_the_open_box_routine():
    # raytrace_to_edge stays raytrace_to_edge_car when the shear / plane-atmosphere to_tau routines are bound
    rng = np.random.default_rng(5)
    for kw in (SHEAR, ATM_PLANE, dict(ATM_SPH, xy_symmetry=True, nx=12, ny=12, nz=23)):
        m = Model(**kw).setup()
        g = m.config.contents.grid
        n = 4000
        x = rng.uniform(g.xmin, g.xmax, n); y = rng.uniform(g.ymin, g.ymax, n); z = rng.uniform(g.zmin, g.zmax, n)
        mu = rng.uniform(-1, 1, n); ph = rng.uniform(0, 2 * np.pi, n)
        kx, ky, kz = np.sqrt(1 - mu * mu) * np.cos(ph), np.sqrt(1 - mu * mu) * np.sin(ph), mu
        ic = np.floor((x - g.xmin) / g.dx).astype(np.int32) + 1
        jc = np.floor((y - g.ymin) / g.dy).astype(np.int32) + 1
        kc = np.floor((z - g.zmin) / g.dz).astype(np.int32) + 1
        xf = rng.normal(0, 3, n)
        sim = Simulation(m)
        tau, ns, _ = sim.raytrace_to_edge(x, y, z, kx, ky, kz, xf, ic, jc, kc)
        to, no, _ = oracle.raytrace_to_edge(m.config, x, y, z, kx, ky, kz, xf, ic, jc, kc)
        sim.close()
        np.testing.assert_array_equal(tau, to)
        np.testing.assert_array_equal(ns, no)


def test_extras_are_refused_where_the_reference_has_no_such_binding():
    with pytest.raises(LartError):
        Model(geometry="sphere", source_geometry="plane_illumination").setup()
    m = Model(**ATM_PLANE).setup()
    m.config.contents.par.xy_periodic = 0
    with pytest.raises(LartError):
        Simulation(m)


def test_tau_walks_of_the_extra_bindings_are_bit_exact_and_leave_no_tally():
    """raytrace_to_tau through the batch API under the shear / atmosphere / accumulator bindings: photon state after the walk
    bit-identical to the oracle (a fresh photon carries vfy_shear = 0), masked cells end the walk, and a batch call leaves
    the handle's tallies — the accumulators included — untouched."""
    rng = np.random.default_rng(11)
    cases = (SHEAR, ATM_PLANE, dict(ATM_SPH, nxim=0, nyim=0), dict(ATM_SPH, xy_symmetry=True, nx=12, ny=12, nz=23, nxim=0, nyim=0))
    for kw in cases:
        m = Model(**dict(kw, calc_J=True, calc_Pnew=True)).setup()
        g = m.config.contents.grid
        n = 20000
        x = rng.uniform(g.xmin, g.xmax, n); y = rng.uniform(g.ymin, g.ymax, n); z = rng.uniform(g.zmin, g.zmax, n)
        mu = rng.uniform(-1, 1, n); ph = rng.uniform(0, 2 * np.pi, n)
        kx, ky, kz = np.sqrt(1 - mu * mu) * np.cos(ph), np.sqrt(1 - mu * mu) * np.sin(ph), mu
        ic = np.floor((x - g.xmin) / g.dx).astype(np.int32) + 1
        jc = np.floor((y - g.ymin) / g.dy).astype(np.int32) + 1
        kc = np.floor((z - g.zmin) / g.dz).astype(np.int32) + 1
        xf = rng.normal(0, 2, n)
        tau_in = rng.exponential(size=n) * 5.0
        sim = Simulation(m)
        a = sim.raytrace_to_tau(x, y, z, kx, ky, kz, xf, ic, jc, kc, tau_in)
        b = oracle.raytrace_to_tau(m.config, x, y, z, kx, ky, kz, xf, ic, jc, kc, tau_in)
        for key in ("inside", "icell", "jcell", "kcell", "nsteps", "x", "y", "z", "xfreq"):
            np.testing.assert_array_equal(a[key], b[key], err_msg=key)
        assert 0.02 < a["inside"].mean() < 0.98
        sim.output_reduce()   # fetch: nothing was tallied by the batch
        sim.close()
        assert m.jp_array("J").sum() == 0.0 and m.jp_array("Pnew").sum() == 0.0 and m.spectrum("Jout").sum() == 0.0
