import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Everything native is built once per session (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()


def small_sphere(**kw):
    """A small static sphere with peel-off: the work-horse configuration of the parity tests."""
    from lart_b200 import Model
    par = dict(no_photons=2000, temperature=1e4, taumax=1e2, nx=31, ny=31, nz=31, rmax=1.0, use_stokes=True,
               nxfreq=61, nxim=17, nyim=17, save_all_photons=True, iseed=7)
    par.update(kw)
    return Model(**par).setup()


def golden(case, name):
    """A value the reference printed in one of its own example logs (tests/golden/reference_logs.json, transcribed by
    tools/extract_reference_logs.py with its examples/<file>:<line> source)."""
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_logs.json")) as fh:
        return json.load(fh)[case][name]["value"]
