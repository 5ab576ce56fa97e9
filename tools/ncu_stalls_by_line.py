#!/usr/bin/env python3
"""Stall-reason samples per CUDA source line (joins the ncu SASS source page with nvdisasm -g line info).
usage: ncu_stalls_by_line.py <ncu-rep> <kernel-regex> <lib.so> [top]"""
import csv, re, subprocess, sys, tempfile, os, collections
rep, kre, so = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kname, hdr = rows[0][1], rows[1]
ia = hdr.index("Address")
reasons = ["stall_long_sb", "stall_wait", "stall_no_inst", "stall_short_sb", "stall_math", "stall_branch_resolving", "stall_lg", "stall_mio", "stall_barrier", "stall_not_selected", "stall_selected", "stall_dispatch"]
ri = [hdr.index(r) for r in reasons]
body = [r for r in rows[2:] if len(r) > max(ri) and r[ia].startswith("0x")]
base = int(body[0][ia], 16)
met = {int(r[ia], 16) - base: [int(r[i] or 0) for i in ri] for r in body}
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
short = re.search(r"(k_\w+)", kname).group(1)
targs = re.search(r"<([0-9, ]+)>", kname.replace("(bool)", ""))  # template instantiation, e.g. k_wf_scatter<1, 0, 0> -> k_wf_scatterILb1ELb0ELb0E
if targs:
    short += "I" + "".join("Lb%sE" % a.strip() for a in targs.group(1).split(","))
start = next(i for i, l in enumerate(dis) if l.startswith("_Z") and short in l and l.rstrip().endswith(":"))
cur = "?"; agg = collections.defaultdict(lambda: [0] * len(reasons))
for l in dis[start + 1:]:
    if l.startswith("//-----") or (l.startswith("_Z") and l.rstrip().endswith(":")): break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = "%s:%s" % (os.path.basename(m.group(1)), m.group(2)); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m and int(m.group(1), 16) in met:
        for k, v in enumerate(met[int(m.group(1), 16)]): agg[cur][k] += v
tot = [sum(v[k] for v in agg.values()) for k in range(len(reasons))]
allt = sum(tot)
print("kernel %s: stall samples by reason: %s" % (short, ", ".join("%s %.1f%%" % (r[6:], 100.0 * t / allt) for r, t in zip(reasons, tot))))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:top]:
    s = sum(v)
    parts = ", ".join("%s %d" % (reasons[i][6:], v[i]) for i in sorted(range(len(v)), key=lambda i: -v[i])[:3] if v[i])
    print("%6.2f%%  %-28s %s" % (100.0 * s / allt, k, parts))
