"""Sharding of an environment batch over ranks (one process per GPU) and the only collective of the
path: an all-gather of per-rank episode statistics.

Every environment is independent: global env id g lives on rank g // envs_per_rank and
its Philox key is (seed, g), so trajectories do not depend on the world size.  Nothing is exchanged on
the step path.  The reference analogue of the statistics exchange is RLlib's MetricsCallback
aggregation over rollout workers (utils/logger_utils.py:126-151).
"""
import torch
import torch.distributed as dist

STAT_FIELDS = ("apples_eaten", "raw_env_rewards", "transfers", "dirt_cleaned", "sum_transferred_rewards",
               "sum_raw_rewards", "envs", "err_flags")


def shard_range(total_envs, rank, world_size):
    """Contiguous block partition of [0, total_envs): returns (first_env_id, num_envs) of `rank`.

    The first `total_envs % world_size` ranks get one extra env, so any total is covered exactly once.
    """
    if not 0 <= rank < world_size:
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, extra = divmod(int(total_envs), int(world_size))
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def owner_of(env_id, total_envs, world_size):
    """Rank that owns global env `env_id` under shard_range."""
    base, extra = divmod(int(total_envs), int(world_size))
    cut = extra * (base + 1)
    if env_id < cut:
        return env_id // (base + 1)
    return extra + (env_id - cut) // max(base, 1)


def local_episode_stats(metrics_raw):
    """Reduce a rank's raw metric matrix [E, 56] (ssd_get_metrics layout) to the 8-vector of STAT_FIELDS."""
    m = metrics_raw
    e = torch.tensor(float(m.shape[0]), dtype=torch.float64, device=m.device)
    return torch.stack([m[:, 0].sum(), m[:, 2].sum(), m[:, 3].sum(), m[:, 4].sum(),
                        m[:, 40:48].sum(), m[:, 24:32].sum(), e, m[:, 5].max()])


def gather_episode_stats(local_stats, group=None):
    """All-gather the per-rank statistic vectors -> [world, 8] (NCCL on GPU tensors, gloo on CPU tensors).

    Without an initialised process group this is the single-rank identity.
    """
    if not (dist.is_available() and dist.is_initialized()):
        return local_stats.unsqueeze(0)
    world = dist.get_world_size(group)
    flat = local_stats.contiguous().view(-1)
    out = torch.empty((world * flat.numel(),), dtype=flat.dtype, device=flat.device)   # flat: gloo requires it
    dist.all_gather_into_tensor(out, flat, group=group)
    return out.view((world,) + tuple(local_stats.shape))


def combine_stats(gathered):
    """Whole-job statistics from the gathered [world, 8] matrix: sums, except err_flags (max)."""
    total = gathered.sum(0)
    total[7] = gathered[:, 7].max()
    return dict(zip(STAT_FIELDS, total.tolist()))
