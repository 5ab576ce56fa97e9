#!/usr/bin/env python3
"""Octree leaf lists from the reference's own grid generator (python/AMR_grid/AMR_grid.py, imported where
/root/reference exists) for the AMR path's tests: tests/golden/amr_sphere_l<min><max>.npz.

The grids are the uniform sphere of examples/amr_sphere_generic/make_amr_sphere_data.py (boxlen 2, radius 1, density 1
inside / 0 outside, T = 1e4 K, static, refine_boundary=True) at small level ranges; the full-size grid of that example
(level_min=3, level_max=7) is generated too, only to check its leaf count against the reference's own log
(examples/amr_sphere_generic/log_amr_1M.txt: "AMR nleaf : 178480"), and is stored as level-coded integers.
Each file holds the leaf centres, levels (generic format: level 1 = the 8 children of the root ... the reader's
convention is kept as the generator writes it), and densities, in the generator's leaf order."""
import os, sys, types
import numpy as np

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
sys.path.insert(0, os.path.join(ref, "python", "AMR_grid"))
for name in ("matplotlib", "matplotlib.pyplot"):
    try:
        __import__(name)
    except Exception:
        sys.modules.setdefault(name, types.ModuleType(name))
from AMR_grid import AMRGrid  # noqa: E402


def make(level_min, level_max):
    g = AMRGrid(2.0)
    dens_fn = lambda x, y, z: 0.0 if x * x + y * y + z * z > 1.0 else 1.0
    vel_fn = lambda x, y, z: (0.0, 0.0, 0.0)
    g.refine_sphere_by_physics(0, 0, 0, 1.0, dens_fn=dens_fn, vel_fn=vel_fn, dens_threshold=0.1, vel_threshold=0.1,
                               level_min=level_min, level_max=level_max, nprobe=2, refine_boundary=True)
    g.set_density(dens_fn)
    g.set_temperature(1.0e4)
    g.set_velocity(vel_fn)
    return g


out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
for lmin, lmax in ((2, 4), (2, 5), (3, 7)):
    g = make(lmin, lmax)
    leaves = list(g.leaves())
    cx = np.array([c.cx for c in leaves]); cy = np.array([c.cy for c in leaves]); cz = np.array([c.cz for c in leaves])
    lev = np.array([c.level for c in leaves], dtype=np.int8)
    dens = np.array([getattr(c, "dens", getattr(c, "density", 0.0)) for c in leaves])
    # integer coordinates of the centre on the leaf's own level: cx = -1 + (2 i + 1) / 2^level
    ix = np.rint(((cx + 1.0) * 2.0 ** lev - 1.0) / 2.0).astype(np.int16)
    iy = np.rint(((cy + 1.0) * 2.0 ** lev - 1.0) / 2.0).astype(np.int16)
    iz = np.rint(((cz + 1.0) * 2.0 ** lev - 1.0) / 2.0).astype(np.int16)
    assert np.array_equal(-1.0 + (2.0 * ix + 1.0) / 2.0 ** lev, cx)
    path = os.path.join(out, "amr_sphere_l%d%d.npz" % (lmin, lmax))
    np.savez_compressed(path, level=lev, ix=ix, iy=iy, iz=iz, dens=dens.astype(np.float32), boxlen=2.0)
    print(path, len(leaves), g.level_counts(), "bytes", os.path.getsize(path), "dens values", np.unique(dens)[:5])
