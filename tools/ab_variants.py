#!/usr/bin/env python3
"""Run bench.py's device-resident arm once per engine variant (lart_b200/variants/liblart_gpu_<name>.so; "default" =
the in-tree library) and print / save value and per-stage times.
usage: ab_variants.py out.json [--workload W] [--steps K] name [name ...]"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = sys.argv[1:]
out = args.pop(0)
extra = []
while args and args[0].startswith("--"):
    extra += [args.pop(0), args.pop(0)]
res = {}
for name in args:
    env = dict(os.environ)
    if name != "default":
        env["LART_GPU_LIB"] = os.path.join(ROOT, "lart_b200", "variants", "liblart_gpu_%s.so" % name)
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--skip-e2e", "--no-cpu-baseline", "--complete-photons", "0", "--steps", "10", "--warmup", "3"] + extra
    p = subprocess.run(cmd, env=env, capture_output=True, text=True)
    try:
        j = json.loads(p.stdout.strip().splitlines()[-1])
        r = j["roofline"]
        st = {k: v["ms_per_launch"] for k, v in r["per_kernel"].items()}
        res[name] = {"value": j["value"], "stage_ms_per_wave": st, "cellsteps_per_s": j.get("cellsteps_per_s")}
        print("%-14s value %.4g  stages %s" % (name, j["value"], {k: round(v, 4) for k, v in st.items()}), flush=True)
    except Exception as e:
        res[name] = {"error": str(e), "stderr": p.stderr[-2000:]}
        print(name, "FAILED", p.stderr[-1500:], flush=True)
json.dump(res, open(out, "w"), indent=1)
