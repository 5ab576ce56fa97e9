import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lart_b200 import Model, Simulation
import bench
w = sys.argv[1] if len(sys.argv) > 1 else "sphere_peel_tau1e7"
t=time.perf_counter(); m = Model(no_photons=10000000, iseed=1, **bench.WORKLOADS[w]).setup(); print("host setup %.3f" % (time.perf_counter()-t))
for rep in range(2):
    t0=time.perf_counter(); sim = Simulation(m); t1=time.perf_counter()
    sim.begin(1, 10000000); t2=time.perf_counter()
    ts=[]
    for _ in range(5):
        a=time.perf_counter(); sim.step(32); ts.append(time.perf_counter()-a)
    t3=time.perf_counter(); sim.sync(); t3b=time.perf_counter(); sim.output_reduce(); t4=time.perf_counter(); sim.close(); t5=time.perf_counter()
    print(w, "create %.3f begin %.3f steps %s sync %.3f reduce+fetch %.3f close %.3f" % (t1-t0,t2-t1,["%.3f"%x for x in ts],t3b-t3,t4-t3b,t5-t4))
