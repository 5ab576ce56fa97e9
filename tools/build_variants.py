#!/usr/bin/env python3
"""Build A/B variants of the CUDA engine into lart_b200/variants/ (git-ignored .so files that travel to the GPU box).

usage: build_variants.py name[=REV]:-DMACRO=..,-DMACRO2=.. [...]
  name=REV builds the engine sources of git revision REV (e.g. r1=693a115) instead of the working tree.
Run them with tools/ab_variants.py (LART_GPU_LIB selects the library; never a fallback).
"""
import os, subprocess, sys, tempfile, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lart_b200.build import NVCC_FLAGS, nvcc_path

out_dir = os.path.join(ROOT, "lart_b200", "variants")
os.makedirs(out_dir, exist_ok=True)
procs = []
for spec in sys.argv[1:]:
    name, _, defs = spec.partition(":")
    name, _, rev = name.partition("=")
    flags = [d for d in defs.split(",") if d]
    src_dir = os.path.join(ROOT, "lart_b200", "csrc")
    inc = os.path.join(ROOT, "include")
    if rev:
        tmp = tempfile.mkdtemp(prefix="lart_" + name)
        for sub in ("lart_b200/csrc", "include"):
            os.makedirs(os.path.join(tmp, sub))
            for f in subprocess.check_output(["git", "ls-tree", "--name-only", rev, sub + "/"], cwd=ROOT, text=True).split():
                open(os.path.join(tmp, f), "wb").write(subprocess.check_output(["git", "show", "%s:%s" % (rev, f)], cwd=ROOT))
        src_dir = os.path.join(tmp, "lart_b200", "csrc")
    out = os.path.join(out_dir, "liblart_gpu_%s.so" % name)
    cmd = [nvcc_path()] + NVCC_FLAGS + flags + ["-o", out, os.path.join(src_dir, "lart_engine.cu")]
    print(" ".join(cmd), flush=True)
    procs.append((name, subprocess.Popen(cmd)))
bad = [n for n, p in procs if p.wait() != 0]
if bad:
    sys.exit("failed: %s" % bad)
