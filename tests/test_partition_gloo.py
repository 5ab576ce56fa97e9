"""Multi-rank host logic on CPU (gloo, world_size 2): the photon-id partition of
run_equal_number (run_simulation_mod.f90:150) and the single sum-reduce of the tallies.
The oracle in Philox mode stands in for the engine: its photon histories depend only on
(seed, photon id), exactly like the GPU engine's, so the reduced spectrum of two ranks
must equal the one-rank spectrum."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


MEDIA = {
    "cartesian": dict(no_photons=301, temperature=1e4, taumax=30.0, nx=9, ny=9, nz=9, rmax=1.0, nxfreq=41, iseed=3),
    # the clump population is generated from par%iseed on every rank: identical layouts, as upstream broadcasts it
    "clumps": dict(no_photons=301, use_clump_medium=True, rmax=1.0, clump_radius=0.05, clump_f_cov=2.0, clump_tau0=5.0,
                   temperature=1e4, clump_sigma_v=10.0, nxfreq=41, nx=11, ny=11, nz=11, iseed=3),
}


def _worker(rank, world, port, q, medium):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lart_b200.host import Model, photon_partition
    from oracle import oracle
    m = Model(**MEDIA[medium]).setup()
    first, count, stride = photon_partition(301, rank, world)
    oracle.run(m, rng_mode=1, nthreads=1, first_id=first, count=count, stride=stride)
    t = torch.from_numpy(np.concatenate([m.spectrum("Jout"), m.spectrum("Jin"), [m.nscatt_gas, m.counters["n_photons_done"]]]))
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        q.put(t.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("medium", sorted(MEDIA))
def test_two_rank_partition_and_reduce_equals_one_rank(medium):
    from lart_b200.host import Model, photon_partition
    from oracle import oracle
    # partition covers ids 1..N exactly once for any world size
    for world in (1, 2, 3, 8):
        ids = []
        for r in range(world):
            f, c, s = photon_partition(301, r, world)
            ids += [f + i * s for i in range(c)]
        assert sorted(ids) == list(range(1, 302))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 7 * sorted(MEDIA).index(medium)) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, medium)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m = Model(**MEDIA[medium]).setup()
    oracle.run(m, rng_mode=1, nthreads=1)
    ref = np.concatenate([m.spectrum("Jout"), m.spectrum("Jin"), [m.nscatt_gas, m.counters["n_photons_done"]]])
    assert got[-1] == 301
    assert np.allclose(got, ref, rtol=1e-12, atol=0)


# ---- dynamic dealing (run_simulation_mod.f90:31-128 without a master): the node's shared photon counter, on CPU -------------
def _deal_worker(rank, world, port, q, name, nphotons, batch):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ctypes as C
    from lart_b200 import capi
    from lart_b200.host import Model
    from oracle import oracle
    lib = capi.load_gpu()  # the claim is a host-only entry point of the engine library: no GPU is touched
    d = C.c_void_p()
    if rank == 0:
        assert lib.lart_gpu_deal_open(name.encode(), 1, C.byref(d)) == 0   # creates and zeroes the counter ...
    dist.barrier()                                                         # ... before anybody claims
    if rank != 0:
        assert lib.lart_gpu_deal_open(name.encode(), 0, C.byref(d)) == 0
    m = Model(**MEDIA["cartesian"]).setup()
    first, count, mine = C.c_int64(), C.c_int64(), []
    while True:
        assert lib.lart_gpu_deal_claim(d, nphotons, batch, C.byref(first), C.byref(count)) == 0
        if count.value == 0:
            break
        mine += list(range(first.value, first.value + count.value))
        oracle.run(m, rng_mode=1, nthreads=1, first_id=first.value, count=count.value, stride=1)
    t = torch.from_numpy(np.concatenate([m.spectrum("Jout"), m.spectrum("Jin"), [m.nscatt_gas, m.counters["n_photons_done"]]]))
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
    ids = [None] * world
    dist.all_gather_object(ids, mine)
    if rank == 0:
        q.put((t.numpy().copy(), ids))
    dist.barrier()
    lib.lart_gpu_deal_close(d, 1 if rank == 0 else 0)
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [7, 100])
def test_two_ranks_claiming_from_the_shared_counter_cover_every_photon_once(batch):
    from lart_b200.host import Model
    from oracle import oracle
    n = 301
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 13 * batch) % 2000
    name = "/lart_gloo_deal_%d_%d" % (os.getpid(), batch)
    procs = [ctx.Process(target=_deal_worker, args=(r, 2, port, q, name, n, batch)) for r in range(2)]
    for p in procs:
        p.start()
    summed, ids = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert sorted(ids[0] + ids[1]) == list(range(1, n + 1))   # every id exactly once, whoever claimed it
    assert all(len(x) % batch in (0, n % batch) for x in ids)
    m = Model(**MEDIA["cartesian"]).setup()
    oracle.run(m, rng_mode=1, nthreads=1)
    one = np.concatenate([m.spectrum("Jout"), m.spectrum("Jin"), [m.nscatt_gas, m.counters["n_photons_done"]]])
    assert np.allclose(summed, one, rtol=1e-12, atol=1e-12)
