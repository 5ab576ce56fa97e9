"""Phases of the end-to-end path through the public API (create / steps / reduce+fetch / destroy), two handles in a row;
LART_GPU_TIMING=1 prints the phases of lart_gpu_create."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lart_b200 import Model, Simulation
import bench
w = sys.argv[1] if len(sys.argv) > 1 else "sphere_peel_tau1e7"
t = time.perf_counter(); m = Model(no_photons=148 * 16384 * 4, iseed=1, **bench.WORKLOADS[w]).setup(); print("host setup %.3f" % (time.perf_counter() - t))
for rep in range(3):
    m.zero_tallies()
    t0 = time.perf_counter(); sim = Simulation(m, quantum=32); t1 = time.perf_counter()
    sim.begin(1, 148 * 16384 * 4); t2 = time.perf_counter()
    ts = []
    for _ in range(5):
        a = time.perf_counter(); sim.step(32); ts.append(time.perf_counter() - a)
    t3 = time.perf_counter(); sim.sync(); t3b = time.perf_counter(); sim.output_reduce(); t4 = time.perf_counter(); sim.close(); t5 = time.perf_counter()
    print(w, "create %.3f begin %.3f steps %s sync %.3f reduce+fetch %.3f close %.3f" % (t1 - t0, t2 - t1, ["%.3f" % x for x in ts], t3b - t3, t4 - t3b, t5 - t4), flush=True)
