"""N>1 host-side logic on CPU: world_size-2 gloo process group, env-id sharding, max-over-ranks timing and the optional
episode-statistics all-reduce (the only collectives this design has)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from marl_llm_b200.sharding import all_reduce_stats, episode_stats, max_over_ranks, shard_range


def test_shard_ranges_partition_env_ids():
    for total, world in [(65536, 8), (10, 3), (7, 8), (1, 1), (4096, 2)]:
        seen = []
        for r in range(world):
            first, count = shard_range(total, r, world)
            seen += list(range(first, first + count))
        assert seen == list(range(total))
        counts = [shard_range(total, r, world)[1] for r in range(world)]
        assert max(counts) - min(counts) <= 1


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = shard_range(10, rank, world)
    # each rank "steps" its own envs: synthetic local rewards that depend only on the GLOBAL env id
    reward = torch.tensor([[float(e % 3 == 0)] for e in range(first, first + count)])
    in_flags = torch.ones(count, 1)
    stats = all_reduce_stats(episode_stats(reward, in_flags))
    slowest = max_over_ranks(1.0 + rank)
    out[rank] = (stats.tolist(), slowest, first, count)
    dist.destroy_process_group()


def test_two_rank_gloo_stats_and_timing(monkeypatch):
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    monkeypatch.setenv("PYTHONPATH", repo + os.pathsep + os.environ.get("PYTHONPATH", ""))   # spawned ranks re-import this module
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    expect_reward = float(sum(e % 3 == 0 for e in range(10)))
    for r in range(2):
        stats, slowest, first, count = out[r]
        assert stats == [expect_reward, 10.0, 10.0]        # identical on both ranks after the all-reduce
        assert slowest == 2.0                              # max over ranks
    assert out[0][2:] == (0, 5) and out[1][2:] == (5, 5)


def test_no_process_group_is_a_noop():
    assert max_over_ranks(3.5) == 3.5
    st = all_reduce_stats(torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64))
    assert np.allclose(st.numpy(), [1, 2, 3])
