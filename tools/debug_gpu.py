"""Ad-hoc GPU debugging: one parity case per driver with timings and first mismatches."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from conftest import small_sphere
from lart_b200 import Simulation, capi, sample
from oracle import oracle
from test_gpu_runs import CASES

def one(case, flags):
    kw = CASES[case]
    t0 = time.time(); mg, mo = small_sphere(**kw), small_sphere(**kw); t1 = time.time()
    sim = Simulation(mg, flags=flags, pool_slots=4096); t2 = time.time()
    n = mg.config.contents.par.nphotons
    sim.begin(1, n)
    nstep = 0
    left = n
    while left > 0 and nstep < 400:
        left = sim.step(0); nstep += 1
    t3 = time.time()
    sim.output_reduce(); sim.close(); t4 = time.time()
    oracle.run(mo, rng_mode=1); t5 = time.time()
    ng, no = mg.allph("nscatt_gas"), mo.allph("nscatt_gas")
    same = np.isclose(ng, no, rtol=1e-12, atol=0)
    print("%-34s %-10s setup %.2f create %.2f run %.2f (%d steps, left %d) fetch %.2f oracle %.2f | same %.4f  <N> gpu %.2f ora %.2f" %
          (case, "mono" if flags & 4 else "wavefront", t1 - t0, t2 - t1, t3 - t2, nstep, left, t4 - t3, t5 - t4, same.mean(), ng.mean(), no.mean()), flush=True)
    bad = np.where(~same)[0][:5]
    for name in ("xfreq1", "xfreq2", "rp", "rp0", "Q"):
        a, b = mg.allph(name), mo.allph(name)
        if a is None or b is None: continue
        ok = np.isclose(a[same], b[same], rtol=1e-8, atol=1e-9)
        print("    %-7s close among same: %.4f   first bad ids %s" % (name, ok.mean(), (np.where(same)[0][~ok][:4] + 1).tolist()))
        for j in np.where(same)[0][~ok][:4]:
            print("        id %d %s gpu %.17g ora %.17g  nscatt %.17g / %.17g  xfreq2 %.17g / %.17g" % (j + 1, name, a[j], b[j], ng[j], no[j], mg.allph("xfreq2")[j], mo.allph("xfreq2")[j]))
    for i in bad:
        print("    id %d: nscatt gpu %g ora %g  xfreq1 %g/%g xfreq2 %g/%g" % (i + 1, ng[i], no[i], mg.allph("xfreq1")[i], mo.allph("xfreq1")[i], mg.allph("xfreq2")[i], mo.allph("xfreq2")[i]))
    cg, co = mg.counters, mo.counters
    print("    counters gpu", {k: int(v) for k, v in cg.items()})
    print("    counters ora", {k: int(v) for k, v in co.items()})
    print("    Jout sum gpu %.6f ora %.6f | nscatt_gas %.3f / %.3f" % (mg.spectrum("Jout").sum(), mo.spectrum("Jout").sum(), mg.nscatt_gas, mo.nscatt_gas))
    for nm in ("scatt", "direc", "I", "Q"):
        a, b = mg.observer_cube(nm), mo.observer_cube(nm)
        if a is not None and b is not None:
            print("    cube %-6s sum gpu %.6e ora %.6e  L1diff/L1 %.3e" % (nm, a.sum(), b.sum(), np.abs(a - b).sum() / (np.abs(b).sum() + 1e-300)))

if __name__ == "__main__":
    ids = np.arange(1, 20001, dtype=np.int64)
    for x0 in (4.5,):
        g = sample(2, 7, ids, x0, 4.7186e-4, ndraw=1).ravel(); o = oracle.sample(2, 7, ids, x0, 4.7186e-4, ndraw=1).ravel()
        d = np.abs(g - o)
        bad = np.where(~np.isclose(g, o, rtol=1e-9, atol=1e-12))[0]
        print("vz x0=%.1f: mismatches %d of %d; max |d| among close %.3e" % (x0, bad.size, g.size, d[np.isclose(g, o, rtol=1e-9, atol=1e-12)].max()))
        for i in bad[:8]:
            print("    id %d gpu %.15g ora %.15g" % (ids[i], g[i], o[i]))
    cases = sys.argv[1:] or ["sphere_stokes_peel", "off_centre_point_mono", "box_uniform_source_gaussian"]
    for c in cases:
        for fl in (capi.FLAG_MONOLITHIC, 0):
            one(c, fl)
