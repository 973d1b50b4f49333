"""Multi-process (world_size 2, gloo, CPU) coverage of the data-parallel host logic: batch sharding, max-over-ranks
timing reduction and the single flat-buffer gradient all-reduce (reference: src/main.py:156,196-200)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wav2vec2 import parallel


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        batch = torch.arange(7 * 3, dtype=torch.float32).reshape(7, 3)
        mine = parallel.shard_batch(batch, rank, world)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        assert torch.equal(torch.cat(gathered), batch)                      # exact partition, order preserved
        assert parallel.max_over_ranks(10.0 + rank) == 10.0 + world - 1
        grads = {"a": torch.full((2, 3), float(rank + 1)), "b": torch.full((5,), 10.0 * (rank + 1))}
        red = parallel.allreduce_gradients(grads, ["b", "a"])
        tot = sum(range(1, world + 1))
        assert torch.equal(red["a"], torch.full((2, 3), float(tot))) and torch.equal(red["b"], torch.full((5,), 10.0 * tot))
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29533, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world))


def test_shard_bounds_cover_everything():
    for n in (1, 7, 32, 33):
        for w in (1, 2, 4, 8):
            spans = [parallel.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
