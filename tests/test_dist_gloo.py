"""N>1 host logic on CPU: two gloo ranks shard a case list, 'solve' their block, and agree on the
max-over-ranks time and the gathered, ordered results (what bench.py --gpus N and a sharded BulkRun do
with NCCL on the GPU box)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from apmapflow_b200 import dist as apm_dist


def test_shard_is_a_balanced_partition():
    items = list(range(10))
    for world in (1, 2, 3, 4, 8, 16):
        blocks = [apm_dist.shard(items, r, world) for r in range(world)]
        assert sum(blocks, []) == items
        sizes = [len(b) for b in blocks]
        assert max(sizes) - min(sizes) <= 1
    assert apm_dist.shard([], 0, 2) == []


def _worker(rank, world, port, tmpdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert apm_dist.world() == (rank, rank, world)
        cases = [("map%d" % (k // 4), 62.5 * (k % 4)) for k in range(10)]     # (map, outlet pressure) sweep
        mine = apm_dist.shard(cases, rank, world)
        results = [(name, p_out, 1000.0 - p_out) for name, p_out in mine]     # stand-in for the per-case solve
        t_local = 1.0 + rank                                                  # rank 1 is the slow one
        t = apm_dist.max_over_ranks(t_local)
        gathered = sum(apm_dist.gather_objects(results), [])
        assert t == float(world)
        assert [(g[0], g[1]) for g in gathered] == cases
        with open(os.path.join(tmpdir, "ok%d" % rank), "w") as f:
            f.write("%d %f" % (len(mine), t))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding(tmp_path):
    world, port = 2, 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    counts = [int(open(tmp_path / ("ok%d" % r)).read().split()[0]) for r in range(world)]
    assert counts == [5, 5]
