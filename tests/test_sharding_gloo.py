"""Multi-GPU host logic on CPU: the ensemble shards over ranks with no data-path collective
(world_size 2, gloo).  Each rank regenerates ITS slice of the seeded ensemble, integrates it (here with
the CPU oracle standing in for the device step) and the host gathers; the union must equal the
single-rank run bit for bit.  This is the rank/offset/gather arithmetic bench.py uses under torchrun."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, n_per_rank, q):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch
    import torch.distributed as dist
    import cases as K
    import flint_b200 as F
    from flint_b200 import problems as P
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = K.cr3bp_case(F.ERK_DOP853, 1e-11, 0, norb=0.3)
    y0 = P.cr3bp_ensemble(n_per_rank, start=rank * n_per_rank)          # shard = index range of the seeded ensemble
    o = case.run_oracle(y0, nthreads=2)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)           # the only collectives: barrier + MAX of timings
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, (rank, o["yf"], o["counters"], o["ev_count"]))
    if rank == 0:
        q.put((float(t[0]), gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_equals_single_rank():
    import torch.multiprocessing as mp
    sys.path.insert(0, HERE)
    import cases as K
    import flint_b200 as F
    from flint_b200 import problems as P
    world, n = 2, 96
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    tmax, gathered = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == float(world)
    gathered.sort(key=lambda g: g[0])
    yf = np.concatenate([g[1] for g in gathered], axis=1)
    cnt = np.concatenate([g[2] for g in gathered], axis=1)
    case = K.cr3bp_case(F.ERK_DOP853, 1e-11, 0, norb=0.3)
    single = case.run_oracle(P.cr3bp_ensemble(world * n))
    assert np.array_equal(yf, single["yf"]) and np.array_equal(cnt, single["counters"])
