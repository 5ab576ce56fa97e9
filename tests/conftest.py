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


def amr_sphere(fixture="amr_sphere_l24", **kw):
    """The uniform octree sphere of examples/amr_sphere_generic (leaf lists written by the reference's own grid generator,
    tools/make_amr_fixtures.py -> tests/golden/<fixture>.npz): boxlen 2, density 1 inside r = 1, T = 1e4 K, static."""
    import numpy as np
    from lart_b200 import Model
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", fixture + ".npz"))
    lev = d["level"].astype(np.int32)
    xyz = [-1.0 + (2.0 * d[k] + 1.0) / 2.0 ** lev for k in ("ix", "iy", "iz")]
    par = dict(no_photons=2000, temperature=1e4, taumax=1e2, geometry="sphere", use_stokes=True, nxfreq=61, nxim=17, nyim=17,
               save_all_photons=True, iseed=7)
    par.update(kw)
    vel = par.pop("velocity", None)
    T = par.pop("leaf_temperature", 1.0e4)
    m = Model(**par)
    v = (None, None, None) if vel is None else vel(*xyz)
    m.set_amr_leaves(xyz[0], xyz[1], xyz[2], lev, d["dens"].astype(np.float64), T, v[0], v[1], v[2], boxlen=float(d["boxlen"]))
    return m.setup()
