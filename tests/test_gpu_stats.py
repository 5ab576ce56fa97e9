"""Statistical pins of whole runs (-m gpu).  The reference cannot be built here, so beyond photon-by-photon agreement with
the oracle these are what ties the engine to the reference itself and to the literature:

  * the peak of J_out(x) the reference documents for its own 101^3 Cartesian sphere (docs/LaRT_AMR_description.pdf, section
    11: 7.402e-3 at +-38.21 km/s for tau0 = 1e4, 6.836e-3 at +-26.75 km/s for tau0 = 1e2; transcribed into
    tests/golden/reference_logs.json by tools/extract_reference_logs.py);
  * independent generators (GPU Philox vs the reference's MT19937-64 in the CPU oracle): chi^2/dof per bin of BASELINE's slab
    (examples/slab geometry, T = 1e4 K, Stokes on) and of the peel-off cube (spectrum over the disc, radial profile) of an
    expanding sphere, with the variance of every bin measured from sub-runs.

The Neufeld slab is gated in tests/test_gpu_runs.py at a*tau0 = 1.5e3 (shape, peak, normalisation).  A per-bin chi^2 against that
curve is not a usable gate at any tractable a*tau0: tools/neufeld_slab.py measures chi^2/dof = 26 at BASELINE's T = 1e4 K,
tau0 = 1e6 (a*tau0 = 472, 3e4 photons) and 59 at a*tau0 = 1.5e3 with 1e5 photons — the analytic curve's own ~10 % error
(it is the a*tau0 -> infinity limit), identical for the CPU oracle, while GPU-vs-oracle chi^2/dof is 1 (below).
"""
import numpy as np
import pytest

from conftest import golden
from lart_b200 import Model, Simulation
from oracle import oracle

pytestmark = pytest.mark.gpu

def run_gpu(m, **kw):
    sim = Simulation(m, **kw)
    sim.run_simulation()
    sim.output_reduce()
    sim.close()


@pytest.mark.parametrize("tau0,key,n", [(1e2, "tau1e2", 1000000), (1e4, "tau1e4", 1000000)])
def test_documented_Jout_peak_of_the_cartesian_sphere(tau0, key, n):
    """101^3 sphere, T = 1e4 K, central point source, Voigt injection, 121 bins on [-9, 9], 1e6 photons as in the document."""
    m = Model(no_photons=n, temperature=1e4, taumax=tau0, nx=101, ny=101, nz=101, rmax=1.0, nxfreq=121, spectral_type="voigt",
              source_geometry="point", iseed=20260423).setup()
    run_gpu(m)
    m.output_normalize()
    s = m.summary
    v = m.xfreq() * s.vtherm  # km/s
    J = m.spectrum("Jout")
    i = int(np.argmax(J))
    v_doc, J_doc = golden("doc_car_sphere_101", "peak_kms_" + key), golden("doc_car_sphere_101", "Jout_max_" + key)
    assert abs(abs(v[i]) - v_doc) < 0.01, (v[i], v_doc)            # same bin (bins are 1.91 km/s wide)
    j = len(J) - 1 - i                                               # its mirror bin
    assert abs(J[j] / J[i] - 1) < 0.04                               # symmetric profile
    # the document's value is the maximum over the bins of a 1e6-photon run (each peak bin holds ~2e4 photons: 0.7 % noise,
    # and the larger of two noisy peaks sits ~0.4 % above their mean)
    assert J[i] == pytest.approx(J_doc, rel=0.025), (J[i], J_doc)
    assert 0.5 * (J[i] + J[j]) == pytest.approx(J_doc, rel=0.025)
    assert J.sum() * s.dxfreq * 8 * np.pi ** 2 == pytest.approx(1.0, rel=2e-3)  # output_sum_rect.f90:174-208


def test_slab_stokes_spectrum_chi2_against_mt_oracle():
    """Same physics, independent generators and implementations: GPU (Philox) vs CPU oracle (MT19937-64), slab with Stokes,
    tau0 = 1e4 at T = 1e4 K.  chi^2/dof per spectral bin ~ 1."""
    ng, no = 200000, 20000
    kw = dict(temperature=1e4, taumax=1e4, xy_periodic=True, nx=1, ny=1, nz=201, nxfreq=121, use_stokes=True)
    mg, mo = Model(no_photons=ng, iseed=3, **kw).setup(), Model(no_photons=no, iseed=4, **kw).setup()
    run_gpu(mg)
    oracle.run(mo, rng_mode=0, seed=4)
    a, b = mg.spectrum("Jout"), mo.spectrum("Jout")
    sel = (a / ng + b / no) * no >= 2 * 40
    z = (a[sel] / ng - b[sel] / no) / np.sqrt(a[sel] / ng ** 2 + b[sel] / no ** 2)
    dof = int(sel.sum())
    assert dof >= 30
    assert (z ** 2).sum() / dof < 1 + 5 * np.sqrt(2.0 / dof), ((z ** 2).sum() / dof, dof)
    assert mg.nscatt_gas / ng == pytest.approx(mo.nscatt_gas / no, rel=0.03)


def _subrun_means(run_one, nsub):
    """mean and variance-of-the-mean of every tally bin from `nsub` independent sub-runs"""
    acc = [run_one(k) for k in range(nsub)]
    out = {}
    for name in acc[0]:
        arr = np.stack([a[name] for a in acc])
        out[name] = (arr.mean(axis=0), arr.var(axis=0, ddof=1) / nsub)
    return out


def test_expanding_sphere_peel_cube_chi2_against_mt_oracle():
    """examples/vel_effect_peel geometry at reduced size (41^3, N_HI = 2e18, Hubble flow 200 km/s, lab-frame source, core skip,
    Stokes): the peel-off cube of the GPU against the MT oracle — spectrum summed over the disc and radial profile summed over
    frequency.  Peel weights are far from unit weights, so every bin's variance is measured from 8 sub-runs."""
    kw = dict(temperature=1e4, N_HI=2e18, taumax=-999.0, Vexp=200.0, velocity_type="hubble", xfreq_min=-60.0, xfreq_max=20.0,
              nxfreq=80, use_stokes=True, comoving_source=False, core_skip=True, nx=41, ny=41, nz=41, rmax=1.0, nxim=33, nyim=33)
    nsub, ng, no = 8, 40000, 6000

    def reduce_cube(m, n):
        cube = m.observer_cube("scatt") / n
        nx = cube.shape[1]
        yy, xx = np.meshgrid(np.arange(nx) - (nx - 1) / 2, np.arange(nx) - (nx - 1) / 2)
        rbin = np.minimum((np.hypot(xx, yy) / ((nx - 1) / 2) * 8).astype(int), 8)
        img = cube.sum(axis=0)
        return {"spectrum": cube.sum(axis=(1, 2)), "radial": np.bincount(rbin.ravel(), weights=img.ravel(), minlength=9)[:8],
                "Jout": m.spectrum("Jout") / n}

    def gpu_one(k):
        m = Model(no_photons=ng, iseed=100 + k, **kw).setup()
        run_gpu(m)
        return reduce_cube(m, ng)

    def cpu_one(k):
        m = Model(no_photons=no, iseed=500 + k, **kw).setup()
        oracle.run(m, rng_mode=0, seed=500 + k)
        return reduce_cube(m, no)

    g, o = _subrun_means(gpu_one, nsub), _subrun_means(cpu_one, nsub)
    for name in ("spectrum", "radial", "Jout"):
        (mg, vg), (mo, vo) = g[name], o[name]
        sel = (mo > 0.01 * mo.max()) & (vg + vo > 0)
        z2 = (mg[sel] - mo[sel]) ** 2 / (vg[sel] + vo[sel])
        dof = int(sel.sum())
        assert dof >= 6, (name, dof)
        # variances from 8 sub-runs are themselves noisy (a ratio of chi^2 variables: mean (nsub-1)/(nsub-3) = 1.4, heavy tail)
        assert z2.sum() / dof < 1.4 + 6 * np.sqrt(2.0 / dof), (name, z2.sum() / dof, dof)
