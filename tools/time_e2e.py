import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lart_b200 import Model, Simulation
import bench
t=time.perf_counter(); m = Model(no_photons=4000000, iseed=1, **bench.WORKLOADS["sphere_peel_tau1e7"]).setup(); print("host setup %.3f" % (time.perf_counter()-t))
for rep in range(2):
    t0=time.perf_counter(); sim = Simulation(m); t1=time.perf_counter()
    sim.begin(1, 4000000); t2=time.perf_counter()
    for _ in range(5): sim.step(32)
    t3=time.perf_counter(); sim.output_reduce(); t4=time.perf_counter(); sim.close(); t5=time.perf_counter()
    print("create %.3f begin %.3f steps %.3f reduce+fetch %.3f close %.3f" % (t1-t0,t2-t1,t3-t2,t4-t3,t5-t4))
