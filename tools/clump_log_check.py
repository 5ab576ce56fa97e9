"""Run examples/clump_sphere/clump_NHI18_fcov1 (as logged in the reference's examples/clump_sphere/log_back:4-55:
N_clumps = 1333333, cl_rhokap = 4.4261E+07, tauhomo = 5.89826E+04, <N_scatt> = 4.3454E+03 with 1e6 photons) on the GPU."""
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from lart_b200 import Model, Simulation  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
kw = dict(use_clump_medium=True, rmax=1.0, clump_radius=0.001, clump_f_cov=1.0, N_HImax=1e18, temperature=1e4,
          clump_sigma_v=0.0, spectral_type="monochromatic", geometry="sphere", velocity_type="rotating_galaxy_halo", Vrot=300.0,
          rinner=0.1, nxfreq=500, velocity_min=-1000.0, velocity_max=1000.0, nx=11, ny=11, nz=11, nxim=0, nyim=0,
          save_all_photons=True)
m = Model(no_photons=n, iseed=int(sys.argv[2]) if len(sys.argv) > 2 else 2026, **kw).setup()
sim = Simulation(m)
t = time.time()
sim.run_simulation()
sim.output_reduce()
dt = time.time() - t
ns = m.allph("nscatt_gas")
print("photons %d  wall %.1f s  <N_scatt> %.4e  (log_back: 4.3454E+03)  stderr %.2f%%  scatterings/s %.3e  Jout sum %.1f"
      % (n, dt, m.nscatt_gas / n, 100 * ns.std() / ns.mean() / np.sqrt(n), m.nscatt_gas / dt, m.spectrum("Jout").sum()))
sim.close()
