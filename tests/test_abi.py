"""The C-ABI library loads and exports every symbol include/*.h declares (no compute, no GPU)."""
import ctypes
import os
import re

from lart_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header, prefix):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(%s\w+)\s*\(" % prefix, text)))


def test_gpu_library_exports_every_declared_symbol():
    names = declared("lart_gpu.h", "lart_gpu_")
    assert names == sorted(capi.GPU_SYMBOLS)
    lib = ctypes.CDLL(capi.gpu_lib_path())
    for n in names:
        assert getattr(lib, n) is not None


def test_host_library_exports_every_declared_symbol():
    names = declared("lart_host.h", "lart_host_")
    assert names == sorted(capi.HOST_SYMBOLS)
    lib = ctypes.CDLL(capi.host_lib_path())
    for n in names:
        assert getattr(lib, n) is not None


def test_struct_sizes_match_the_header():
    """ctypes mirrors vs the C compiler's view of include/lart_gpu.h."""
    import subprocess
    import tempfile
    src = r'''
#include <stdio.h>
#include "lart_gpu.h"
#include "lart_host.h"
int main(void){ printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(lart_amr), sizeof(lart_clumps), sizeof(lart_grid), sizeof(lart_params), sizeof(lart_line),
  sizeof(lart_observer), sizeof(lart_scatt_mat), sizeof(lart_config), sizeof(lart_observer_out), sizeof(lart_allph_out),
  sizeof(lart_counters), sizeof(lart_tallies), sizeof(lart_host_summary)); return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, c])
        sizes = list(map(int, subprocess.check_output([exe]).split()))
    mirrors = [capi.Amr, capi.Clumps, capi.Grid, capi.Params, capi.Line, capi.Observer, capi.ScattMat, capi.Config, capi.ObserverOut,
               capi.AllphOut, capi.Counters, capi.Tallies, capi.HostSummary]
    assert sizes == [ctypes.sizeof(t) for t in mirrors]


def test_no_product_code_touches_the_oracle():
    """The oracle is test infrastructure: nothing under lart_b200/ may import, link or name it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "lart_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                if re.search(r"(^|\n)\s*(from|import)\s+oracle|liblart_oracle|oracle/", text):
                    bad.append(f)
    assert not bad, bad


def test_deal_counter_is_shared_memory_and_needs_no_gpu():
    """lart_gpu_deal_open/close (the node-wide photon counter of lart_gpu_run_dealt) are host-only calls."""
    lib = capi.load_gpu()
    name = ("/lart_abi_deal_%d" % os.getpid()).encode()
    a, b = ctypes.c_void_p(), ctypes.c_void_p()
    assert lib.lart_gpu_deal_open(name, 1, ctypes.byref(a)) == 0 and a.value
    assert lib.lart_gpu_deal_open(name, 0, ctypes.byref(b)) == 0 and b.value
    assert os.path.exists("/dev/shm" + name.decode())
    assert lib.lart_gpu_deal_close(b, 0) == 0 and lib.lart_gpu_deal_close(a, 1) == 0
    assert not os.path.exists("/dev/shm" + name.decode())
    assert lib.lart_gpu_deal_open(b"no_leading_slash", 1, ctypes.byref(a)) != 0


def test_fortran_shim_mirrors_the_header_member_for_member():
    """shim/lart_gpu_shim.f90 cannot be compiled here (no Fortran compiler in the image), so its bind(C) types are checked
    textually: every `c_lart_X` type must list the members of `lart_X` in include/lart_gpu.h in the same order, and every
    entry point it binds must be declared in the header."""
    hdr = open(os.path.join(ROOT, "include", "lart_gpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    shim = open(os.path.join(ROOT, "shim", "lart_gpu_shim.f90")).read()
    shim = re.sub(r"!.*", "", shim)
    shim = re.sub(r"&\s*\n\s*", " ", shim)

    def c_members(name):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), hdr, re.S).group(1)
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            decl = re.sub(r"^(const\s+)?(unsigned\s+)?[A-Za-z_0-9]+\s+", "", decl)  # drop the type
            for m in decl.split(","):
                m = re.sub(r"\[.*?\]", "", m).replace("*", "").replace("const", "").strip()
                out.append(m.lower())
        return out

    def f_members(name):
        body = re.search(r"type, bind\(C\) :: c_%s\b(.*?)end type" % name, shim, re.S | re.I).group(1)
        out = []
        for line in body.splitlines():
            if "::" not in line:
                continue
            for m in line.split("::", 1)[1].split(","):
                out.append(re.sub(r"\(.*?\)", "", m).strip().lower())
        return [m for m in out if m]

    for name in ("lart_grid", "lart_params", "lart_line", "lart_observer", "lart_scatt_mat", "lart_clumps", "lart_amr",
                 "lart_config", "lart_observer_out", "lart_allph_out", "lart_counters", "lart_tallies"):
        assert f_members(name) == c_members(name), name
    for fn in re.findall(r"bind\(C, name='(\w+)'\)", shim):
        assert re.search(r"\b%s\s*\(" % fn, hdr), fn
