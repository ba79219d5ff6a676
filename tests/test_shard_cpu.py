"""CPU: the N>1 host logic (frame sharding, max-over-ranks timing, count gather) with world_size-2 gloo."""
import os
import socket

import pytest

from rmcv_b200 import shard


def test_frame_slices_partition_the_batch():
    for total in (0, 1, 7, 1024, 1023):
        for ws in (1, 2, 3, 4, 8):
            sl = shard.all_slices(total, ws)
            assert sl[0][0] == 0 and sl[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
            sizes = [b - a for a, b in sl]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.frame_slice(10, 2, 2)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.frame_slice(11, world, rank)
    ms, units = shard.reduce_timing(10.0 + rank, hi - lo, dist)
    counts = shard.gather_counts(list(range(lo, hi)), dist)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, ms, units, counts))


def test_two_rank_gloo_reduce_and_gather():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in procs)
    [p.join(60) for p in procs]
    for rank, ms, units, counts in res:
        assert ms == 11.0 and units == 11          # MAX over ranks, SUM of units
        assert [x for part in counts for x in part] == list(range(11))   # rank-ordered concatenation
