import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lart_b200 import Model, Simulation
m = Model(no_photons=10, temperature=1e4, N_HI=2e20, Vexp=200.0, velocity_type="hubble", xfreq_min=-200.0, xfreq_max=40.0, nxfreq=500,
          use_stokes=True, nx=201, ny=201, nz=201, rmax=1.0, nxim=129, nyim=129).setup()
sim = Simulation(m, pool_slots=1024)
for rep in range(2):
    t0 = time.perf_counter(); maps = sim.sightline_tau(); dt = time.perf_counter() - t0
st = sim.sightline_stats
print(json.dumps({"workload": "sight-line maps, vel_effect_peel grid 201^3, 129x129 pixels x 500 frequencies", "rays": 129 * 129 * 501,
                  "cellsteps": st["cellsteps"], "kernel_ms": st["ms"], "cellsteps_per_s": st["cellsteps"] / (st["ms"] * 1e-3),
                  "algorithmic_GBps_48B": 48 * st["cellsteps"] / (st["ms"] * 1e-3) / 1e9, "call_seconds": dt,
                  "tau_gas_max": float(maps[0]["tau_gas"].max()), "N_gas_centre": float(maps[0]["N_gas"][64, 64])}))
