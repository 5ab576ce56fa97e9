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
