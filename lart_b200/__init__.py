"""lart_b200 — B200-native engine for LaRT's Cartesian Monte-Carlo photon loop.

The package holds only what that path needs: csrc/ (CUDA kernels + the C ABI of
include/lart_gpu.h, and the C++ mini-host of include/lart_host.h) and the
host-side mirror of the reference's driver sequence (host.py).
"""
from . import capi  # noqa: F401
from .host import (LartError, Model, Simulation, calc_voigt, comm_finalize, comm_init, comm_init_torch, comm_rank,  # noqa: F401
                   measure_fp64, photon_partition, sample)

__version__ = "0.1.0"
