#!/usr/bin/env python3
"""Join an ncu SASS source page (csv) with nvdisasm -g line info and aggregate executed
instructions / stall samples per CUDA source line.

usage: ncu_by_line.py <ncu-rep> <kernel-regex> <liblart_gpu.so> [top]
"""
import csv, re, subprocess, sys, tempfile, os, collections

rep, kre, so = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kname = rows[0][1]
hdr = rows[1]
ia, ie, it, iss = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
body = rows[2:]
base = int(body[0][ia], 16)
met = {}
for r in body:
    if len(r) <= max(ia, ie, it, iss) or not r[ia].startswith("0x"):
        continue
    met[int(r[ia], 16) - base] = (int(r[ie]), int(r[it]), int(r[iss]), r[1].strip())
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
short = re.search(r"(k_\w+)", kname).group(1)
targs = re.search(r"<([0-9, ]+)>", kname.replace("(bool)", ""))  # template instantiation, e.g. k_wf_scatter<1, 0, 0> -> k_wf_scatterILb1ELb0ELb0E
if targs:
    short += "I" + "".join("Lb%sE" % a.strip() for a in targs.group(1).split(","))
start = next(i for i, l in enumerate(dis) if l.startswith("_Z") and short in l and l.rstrip().endswith(":"))
cur = "?"
agg = collections.defaultdict(lambda: [0, 0, 0])
func = collections.defaultdict(lambda: [0, 0, 0])
for l in dis[start + 1:]:
    if l.startswith("//-----") or (l.startswith("_Z") and l.rstrip().endswith(":")):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = "%s:%s" % (os.path.basename(m.group(1)), m.group(2))
        inl = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
        if inl:
            cur += " <- " + " <- ".join("%s:%s" % (os.path.basename(a), b) for a, b in inl)
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m:
        off = int(m.group(1), 16)
        if off in met:
            e, t, s, _ = met[off]
            a = agg[cur]; a[0] += e; a[1] += t; a[2] += s
tot = [sum(v[k] for v in agg.values()) for k in range(3)]
print("kernel %s: warp-instr %d thread-instr %d (avg %.1f thr/instr) samples %d" % (short, tot[0], tot[1], tot[1] / max(tot[0], 1), tot[2]))
print("%-8s %-8s %-6s %-7s  %s" % ("instr%", "samp%", "thr", "", "source line (<- inlined at)"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%7.2f%% %7.2f%% %5.1f   %s" % (100.0 * v[0] / tot[0], 100.0 * v[2] / max(tot[2], 1), v[1] / max(v[0], 1), k))
