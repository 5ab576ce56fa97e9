"""Two ranks on two GPUs through the drop-in path: Simulation.run_simulation(rank, nproc) = run_equal_number's photon
partition (run_simulation_mod.f90:150), output_reduce = lart_gpu_reduce (ONE ncclReduce of the contiguous tally buffer and
of allph inside the C ABI, replacing memory_mod_mpi.f90:366-458 / output_sum_rect.f90:13-146) + lart_gpu_fetch on the root.
The two-rank sum must equal the one-rank run up to FP64 summation order, and a second output_reduce must not count the
other rank twice.  Needs two CUDA devices (skipped on a one-GPU box; run with `gpurun --gpus 2`)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PAR = dict(no_photons=3001, temperature=1e4, taumax=50.0, nx=21, ny=21, nz=21, rmax=1.0, use_stokes=True, nxfreq=41,
           nxim=9, nyim=9, save_all_photons=True, iseed=11)


def _ndev():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)  # host-side plumbing only: carries the 128-byte NCCL id
    from lart_b200 import Model, Simulation
    from lart_b200.host import comm_init_torch, comm_finalize
    comm_init_torch(rank)  # device = rank
    m = Model(**PAR).setup()
    sim = Simulation(m, device=rank, pool_slots=2048)
    sim.run_simulation(rank, world)
    sim.output_reduce(dst=0)
    first = None
    if rank == 0:
        first = (m.spectrum("Jout").copy(), m.observer_cube("scatt").copy(), m.observer_cube("Q").copy(),
                 m.allph("nscatt_gas").copy(), m.nscatt_gas, m.counters["n_photons_done"])
    # a second reduce + fetch right away: the root's device buffer holds the all-rank sum, rank 1's was zeroed by the first
    # reduce, so the fetch adds exactly that sum once more (not the sum plus rank 1's share again)
    sim.output_reduce(dst=0)
    if rank == 0:
        q.put((first, m.counters["n_photons_done"]))
    dist.barrier()
    sim.close()
    comm_finalize()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.skipif(_ndev() < 2, reason="needs two CUDA devices")
def test_two_gpu_reduce_inside_the_abi_equals_one_rank():
    import torch.multiprocessing as mp
    from lart_b200 import Model, Simulation
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    (jout, scatt, Q, nsc, nsg, done), done_twice = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    m = Model(**PAR).setup()
    sim = Simulation(m, device=0, pool_slots=2048)
    sim.run_simulation()
    sim.output_reduce()
    sim.close()
    n = PAR["no_photons"]
    assert done == n and m.counters["n_photons_done"] == n
    assert np.array_equal(nsc, m.allph("nscatt_gas"))                      # per-photon records: disjoint slots, exact
    assert np.allclose(jout, m.spectrum("Jout"), rtol=1e-12, atol=1e-300)
    assert np.allclose(scatt, m.observer_cube("scatt"), rtol=1e-10, atol=1e-300)
    Q1 = m.observer_cube("Q")  # signed sums: compare in norm
    assert np.abs(Q - Q1).sum() <= 1e-10 * np.abs(Q1).sum()
    assert nsg == pytest.approx(m.nscatt_gas, rel=1e-12)
    assert done_twice == 2 * n


# ---- dynamic photon dealing (run_simulation_mod.f90:31-128 without a master): processes of one node claim batches ------
def _dealt_worker(rank, name, batch, q):
    sys.path.insert(0, ROOT)
    from lart_b200 import Model, Simulation
    m = Model(**PAR).setup()
    ndev = _ndev()
    sim = Simulation(m, device=rank % max(ndev, 1), pool_slots=512, quantum=4)
    mine = sim.run_simulation_dealt(name, batch=batch)
    sim.output_reduce()  # no communicator in this process: a plain fetch into its own host arrays
    sim.close()
    q.put((rank, mine, m.spectrum("Jout").copy(), m.observer_cube("scatt").copy(), m.allph("nscatt_gas").copy(), m.nscatt_gas,
           m.counters["n_photons_done"]))


@pytest.mark.gpu
@pytest.mark.parametrize("batch", [100, 700])
def test_dynamically_dealt_photons_sum_to_the_single_run(batch):
    """Two processes share the node's photon counter (both on GPU 0 when the box has one): every photon id is run exactly once,
    by whichever process claimed its batch, and the summed tallies are those of one statically partitioned run."""
    import ctypes as C
    import torch.multiprocessing as mp
    from lart_b200 import Model, Simulation, capi
    lib = capi.load_gpu()
    name = "/lart_test_deal_%d_%d" % (os.getpid(), batch)
    d = C.c_void_p()
    assert lib.lart_gpu_deal_open(name.encode(), 1, C.byref(d)) == 0  # created and zeroed before the workers start
    try:
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        procs = [ctx.Process(target=_dealt_worker, args=(r, name, batch, q)) for r in range(2)]
        for p in procs:
            p.start()
        res = sorted(q.get(timeout=300) for _ in procs)
        for p in procs:
            p.join(timeout=120)
            assert p.exitcode == 0
    finally:
        lib.lart_gpu_deal_close(d, 1)
    n = PAR["no_photons"]
    (_, mine0, j0, s0, a0, g0, d0), (_, mine1, j1, s1, a1, g1, d1) = res
    assert mine0 + mine1 == n and d0 + d1 == n and mine0 == d0 and mine1 == d1
    assert mine0 > 0 and mine1 > 0 and mine0 % batch in (0, n % batch) or mine1 % batch in (0, n % batch)
    m = Model(**PAR).setup()
    sim = Simulation(m, device=0, pool_slots=2048)
    sim.run_simulation()
    sim.output_reduce()
    sim.close()
    ref = m.allph("nscatt_gas")
    assert np.all((a0 == 0) | (a1 == 0))                 # no photon was run twice ...
    assert np.array_equal(a0 + a1, ref)                  # ... and every one was run, with the history of the single run
    assert np.allclose(j0 + j1, m.spectrum("Jout"), rtol=1e-12, atol=1e-300)
    assert np.allclose(s0 + s1, m.observer_cube("scatt"), rtol=1e-10, atol=1e-300)
    assert g0 + g1 == pytest.approx(m.nscatt_gas, rel=1e-12)
