#!/usr/bin/env python3
"""Static SASS instruction count per CUDA source line for one kernel (code-size hot spots).
usage: sass_size_by_line.py <lib.so> <kernel-substring> [top]"""
import re, subprocess, sys, tempfile, os, collections
so, kern = sys.argv[1:3]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith("_Z") and kern in l and l.rstrip().endswith(":"))
cur = "?"; agg = collections.Counter(); sub = collections.Counter(); insub = None; total = 0
for l in dis[start + 1:]:
    if l.startswith("//-----"):
        break
    m = re.match(r"\s*\.weak\s+\$(\S+)|^\$(\S+):", l)
    if l.startswith("$") and l.rstrip().endswith(":"):
        insub = l.split("$")[-1].rstrip(":\n"); continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = "%s:%s" % (os.path.basename(m.group(1)), m.group(2)); continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l) and "NOP" not in l:
        total += 1
        if insub: sub[insub] += 1
        else: agg[cur] += 1
print("kernel *%s*: %d SASS instructions (%.0f KB)" % (kern, total, total * 16 / 1024))
for k, v in sub.most_common(10): print("  subroutine %-60s %6d" % (k[:60], v))
for k, v in agg.most_common(top): print("  %-40s %6d" % (k, v))
