"""Env sharding across ranks.  Envs never interact, so there is NO collective on the step path: each
rank owns a contiguous range of global env ids and steps it with its own kernel launches.  The only exchanges are
(a) the max-over-ranks of a timed region and (b) an optional per-episode statistics all-reduce of a few floats."""
import torch
import torch.distributed as dist


def shard_range(total_envs, rank, world):
    """Contiguous, balanced partition of global env ids [0, total_envs): returns (first, count) for `rank`."""
    base, rem = divmod(int(total_envs), int(world))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def max_over_ranks(value, device="cpu"):
    """Max of a python float over all ranks (elapsed times).  No-op without an initialised process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def episode_stats(reward, in_flags):
    """Local episode statistics as a small tensor [sum_reward, n_in_shape, n_agent_steps] (device of `reward`)."""
    return torch.stack([reward.double().sum(), in_flags.double().sum(),
                        torch.tensor(float(reward.numel()), dtype=torch.float64, device=reward.device)])


def all_reduce_stats(stats):
    """Sum the statistics vector over ranks (NCCL over NVLink on GPUs, gloo on CPU); ~24 bytes, latency-bound."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats
