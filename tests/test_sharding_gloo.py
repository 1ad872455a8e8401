"""CPU, world_size 2 over gloo: the multi-rank host logic — shard ranges, env-id keyed trajectories that do
not depend on the world size (checked with the oracle standing in for the device), and the episode
statistics all-gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from contracts_b200 import sharding


def test_shard_range_partitions_exactly():
    for total in (1, 7, 16, 131072, 1000003):
        for world in (1, 2, 3, 8):
            seen = 0
            for r in range(world):
                first, cnt = sharding.shard_range(total, r, world)
                assert first == seen
                seen += cnt
                for g in {first, first + cnt - 1} if cnt else ():
                    assert sharding.owner_of(g, total, world) == r
            assert seen == total
    with pytest.raises(ValueError):
        sharding.shard_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, steps, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle
    from contracts_b200.maps import CLEANUP_MAP
    first, cnt = sharding.shard_range(total, rank, world)
    o = oracle.GridOracle("cleanup", cnt, 4, CLEANUP_MAP, horizon=steps, contract="CleanupContract",
                          seed=73907, first_env_id=first)
    o.reset()
    rng = np.random.RandomState(5)
    acts = rng.randint(0, 9, size=(steps, total, 4))       # same global action table on every rank
    rew = []
    for t in range(steps):
        rew.append(o.step(acts[t, first:first + cnt], want_features=False)["rew"])
    local = sharding.local_episode_stats(torch.from_numpy(o.metrics_raw()))
    gathered = sharding.gather_episode_stats(local)
    assert gathered.shape == (world, 8)
    assert torch.equal(gathered[rank], local)
    np.save(os.path.join(out_dir, "rew_%d.npy" % rank), np.stack(rew))
    if rank == 0:
        np.save(os.path.join(out_dir, "stats.npy"), gathered.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_world2_matches_single_process(tmp_path, oracle_lib):
    total, steps, world = 11, 60, 2
    mp.spawn(_worker, args=(world, _free_port(), total, steps, str(tmp_path)), nprocs=world, join=True)
    # single-process run of the same global batch
    from contracts_b200.maps import CLEANUP_MAP
    o = oracle_lib.GridOracle("cleanup", total, 4, CLEANUP_MAP, horizon=steps, contract="CleanupContract", seed=73907)
    o.reset()
    acts = np.random.RandomState(5).randint(0, 9, size=(steps, total, 4))
    rew = np.stack([o.step(acts[t], want_features=False)["rew"] for t in range(steps)])
    sharded = np.concatenate([np.load(tmp_path / ("rew_%d.npy" % r)) for r in range(world)], axis=1)
    assert np.array_equal(sharded.view(np.uint64), rew.view(np.uint64)), "trajectories depend on the sharding"
    stats = np.load(tmp_path / "stats.npy")
    whole = sharding.local_episode_stats(torch.from_numpy(o.metrics_raw())).numpy()
    combined = sharding.combine_stats(torch.from_numpy(stats))
    assert combined["envs"] == total
    assert combined["dirt_cleaned"] == whole[3] and combined["apples_eaten"] == whole[0]
    assert abs(combined["transfers"] - whole[2]) < 1e-9
