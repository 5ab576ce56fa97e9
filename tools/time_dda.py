"""Throughput of the batched raytrace_to_edge kernel on long rays (cell steps/s), for two ray directions."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lart_b200 import Model, Simulation
m = Model(no_photons=10, temperature=1e4, taumax=1e2, nx=201, ny=201, nz=201, rmax=1.0, nxfreq=11).setup()
sim = Simulation(m, pool_slots=1024)
rng = np.random.default_rng(0)
n = 4_000_000
p = rng.uniform(-0.6, 0.6, (n, 3))
g = m.config.contents.grid
ic = (np.floor((p + 1.0) / g.dx).astype(np.int32) + 1)
xf = np.full(n, 6.0)
for name, k in (("warm-up", None), ("+z", [0.02, 0.03, 0.9993]), ("+x", [0.9993, 0.03, 0.02]), ("+y", [0.02, 0.9993, 0.03]), ("-z", [0.02, 0.03, -0.9993]),
                ("xz 45deg", [0.7, 0.03, 0.7]), ("random", None), ("+z again", [0.02, 0.03, 0.9993])):
    if k is None:
        kk = rng.normal(size=(n, 3)); kk /= np.linalg.norm(kk, axis=1)[:, None]
    else:
        kk = np.tile(np.array(k) / np.linalg.norm(k), (n, 1))
    for rep in range(2):
        t0 = time.perf_counter()
        tau, ns, _ = sim.raytrace_to_edge(p[:, 0], p[:, 1], p[:, 2], kk[:, 0], kk[:, 1], kk[:, 2], xf, ic[:, 0], ic[:, 1], ic[:, 2])
        dt = time.perf_counter() - t0
    print("%-20s steps %.3e mean %.1f  call %.3f s -> %.3e steps/s (incl. H2D/D2H of %d MB)" % (name, ns.sum(), ns.mean(), dt, ns.sum() / dt, n * 84 / 1e6), flush=True)
