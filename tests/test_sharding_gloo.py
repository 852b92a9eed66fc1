"""Multi-GPU host logic on CPU.  The ensemble shards over devices / ranks by the PRODUCT's partition
(include/flint_b200.h: flint_b200_shard_*, pure host arithmetic inside libflint_b200.so -- the same functions
flint_erk_integrate_batch deals its chunks with and bench.py picks each rank's trajectories with): index space
[0, N) in chunks of 4096, chunk c -> shard c mod n_shards.  There is no data-path collective;
under torchrun the only collectives are the barrier and the MAX of the timings, exercised here with gloo at
world size 2."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_partition_is_block_cyclic_exact_and_balanced():
    import flint_b200 as F
    for N in (1, 5, 4095, 4096, 4097, 8192, 12289, 100000, 1 << 20, (1 << 20) + 17):
        for D in (1, 2, 3, 4, 8):
            for chunk in (4096, 1000, 1):
                if chunk == 1 and N > 20000:
                    continue
                parts = [F.shard_indices(N, D, d, chunk) for d in range(D)]
                counts = [F.shard_count(N, D, d, chunk) for d in range(D)]
                assert [len(p) for p in parts] == counts
                allidx = np.concatenate(parts)
                assert len(allidx) == N and np.array_equal(np.sort(allidx), np.arange(N))      # every trajectory exactly once
                for d, p in enumerate(parts):
                    if len(p) == 0:
                        continue
                    assert np.array_equal((p // chunk) % D, np.full(len(p), d))                # chunk c -> shard c mod D
                    assert (np.diff(p) > 0).all()                                              # local order = global order
                    k = np.arange(len(p))
                    assert np.array_equal(p, ((k // chunk) * D + d) * chunk + k % chunk)       # the copy engine's 2-D pattern
                assert max(counts) - min(counts) <= chunk                                      # balanced to one chunk
    lib = F.api.lib()
    assert lib.flint_b200_shard_global_index(10000, 2, 1, 4096, 0) == 4096
    assert lib.flint_b200_shard_global_index(10000, 2, 1, 4096, 4096) == -1                   # shard 1 owns chunk 1 only
    assert lib.flint_b200_shard_global_index(10000, 2, 0, 4096, 4096) == 8192
    assert F.shard_count(10000, 2, 0) == 4096 + (10000 - 8192) and F.shard_count(10000, 2, 1) == 4096
    assert F.shard_count(100, 8, 3) == 0 and F.shard_count(100, 8, 0) == 100


def test_shard_initial_conditions_are_the_ensembles():
    """A rank regenerates exactly its trajectories of the seeded ensemble from their global indices."""
    import flint_b200 as F
    from flint_b200 import problems as P
    N, D = 3 * 4096 + 123, 4
    full = P.cr3bp_ensemble(N)
    for d in range(D):
        idx = F.shard_indices(N, D, d)
        assert np.array_equal(P.cr3bp_ensemble(0, index=idx), full[:, idx])


def _worker(rank, world, port, N, q):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch
    import torch.distributed as dist
    import flint_b200 as F
    from flint_b200 import problems as P
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = F.shard_indices(N, world, rank)                               # the product's deal
    y0 = P.cr3bp_ensemble(0, index=idx)
    # stand-in for the device step (no GPU here): a per-trajectory function of the inputs, so the scatter below is checked
    yf = y0 * 2.0 + idx[None, :].astype(np.float64)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)           # the only collectives: barrier + MAX of timings
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, (rank, idx, yf))
    if rank == 0:
        q.put((float(t[0]), gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_strong_scaling_split_covers_the_ensemble():
    import torch.multiprocessing as mp
    sys.path.insert(0, HERE)
    from flint_b200 import problems as P
    world, N = 2, 5 * 4096 + 77
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, q)) for r in range(world)]
    for p in procs:
        p.start()
    tmax, gathered = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == float(world)
    out = np.full((6, N), np.nan)
    for _, idx, yf in gathered:
        out[:, idx] = yf                                               # host gather at the trajectories' own offsets
    full = P.cr3bp_ensemble(N)
    assert np.array_equal(out, full * 2.0 + np.arange(N)[None, :].astype(np.float64))
