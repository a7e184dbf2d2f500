"""world_size-2 gloo test of the multi-GPU path's host logic: every rank decodes its own contiguous
shard of the codeword batch (no data-path collective), results are gathered and must equal the
single-process result byte for byte.  On this CPU box the per-shard decoder is the oracle; on the
GPU box bench.py runs the same sharding with the CUDA decoder."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from tests.util import awgn_llr


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ncw, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from myldpccppapi_b200 import shard_range
    N, K = 576, 432
    rp, ci, M = oracle.wimax_H(N, "3/4B")
    llr = awgn_llr(ncw, N, 0.6, seed=5)                      # every rank can regenerate the seeded stream
    b, e = shard_range(ncw, rank, world)
    info, iters, _, _ = oracle.Oracle(M, N, K, rp, ci).decode(llr[b:e], threads=1, want_post=False, want_hard=False)
    # timing contract of bench.py: barrier, then max over ranks of the local time
    dist.barrier()
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    pieces = [None] * world
    dist.all_gather_object(pieces, (b, e, info, iters))
    if rank == 0:
        full = np.concatenate([p[2] for p in sorted(pieces, key=lambda x: x[0])])
        its = np.concatenate([p[3] for p in sorted(pieces, key=lambda x: x[0])])
        q.put((full, its))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_sharded_decode_equals_single():
    ncw, world = 37, 2   # odd count: unequal shards
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ncw, q)) for r in range(world)]
    for p in procs:
        p.start()
    full, its = q.get(timeout=100)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    rp, ci, M = oracle.wimax_H(576, "3/4B")
    info, iters, _, _ = oracle.Oracle(M, 576, 432, rp, ci).decode(awgn_llr(ncw, 576, 0.6, seed=5), want_post=False, want_hard=False)
    assert np.array_equal(full, info) and np.array_equal(its, iters)
