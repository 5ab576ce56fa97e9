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
