"""Small whole-path run for compute-sanitizer (memcheck): both drivers, peel-off, dust, core-skip, slab, compaction."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import small_sphere
from lart_b200 import Simulation, capi, calc_voigt
import numpy as np
calc_voigt(np.linspace(-12, 12, 1001), 4.7e-4)
for kw, flags in ((dict(no_photons=300), 0), (dict(no_photons=300), capi.FLAG_MONOLITHIC),
                  (dict(no_photons=200, use_stokes=False, DGR=1.0, cext_dust=3e-17, taumax=-999.0, N_HI=2e16, core_skip=True), 0),
                  (dict(no_photons=200, xy_periodic=True, nx=1, ny=1, nz=201, rmax=-999.0, taumax=1e2, nxim=0, nyim=0, nxfreq=121), 0),
                  (dict(no_photons=300, obsx=[0.0, 1.0], obsy=[0.0, 0.5], obsz=[1.0, 0.2], save_direc0=True, save_Jmu=True, save_peeloff_2D=True), capi.FLAG_LOCAL_STEPS)):
    m = small_sphere(**kw)
    sim = Simulation(m, flags=flags, pool_slots=128, quantum=4, ray_budget=3)
    sim.run_simulation()
    sim.output_reduce()
    sim.close()
    print("ok", kw.get("no_photons"), flags, m.counters["n_photons_done"], flush=True)
