"""A complete run (all photons to escape) through the public API: photons/s and scatterings/s end to end,
including the heavy tail (lart_gpu_run switches to the monolithic kernel for the last photons)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lart_b200 import Model, Simulation
tau0 = float(sys.argv[1]) if len(sys.argv) > 1 else 1e4
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10 ** 6
core_skip = len(sys.argv) > 3 and sys.argv[3] == "coreskip"
m = Model(no_photons=n, temperature=1e4, taumax=tau0, use_stokes=True, nx=201, ny=201, nz=201, rmax=1.0, nxfreq=201,
          nxim=129, nyim=129, iseed=4, core_skip=core_skip).setup()
t0 = time.perf_counter()
sim = Simulation(m)
t1 = time.perf_counter()
sim.run_simulation()
t2 = time.perf_counter()
sim.output_reduce()
t3 = time.perf_counter()
c = m.counters
dev_ms, launches = sim.kernel_ms()
sim.close()
m.output_normalize()
s = m.summary
omega = s.dxim * s.dyim * (np.pi / 180) ** 2
flux = (m.observer_cube("scatt").sum() + m.observer_cube("direc").sum()) * 4 * np.pi * omega * s.distance ** 2 * s.dxfreq
print(json.dumps({"workload": "sphere_peel 201^3 tau0=%g T=1e4 K, %d photons, Stokes, 201x129x129 cube%s" % (tau0, n, ", core_skip" if core_skip else ""),
                  "photons_per_s": n / (t3 - t0), "scatterings_per_s": c["n_scatter"] / (t3 - t0), "mean_nscatt": m.nscatt_gas,
                  "seconds": {"create": t1 - t0, "run": t2 - t1, "reduce_fetch": t3 - t2, "device_ms": dev_ms},
                  "launches": launches, "cellsteps": c["n_cellsteps"], "peel_rays": c["n_peel"], "flux_check": flux}))
