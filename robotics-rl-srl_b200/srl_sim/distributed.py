"""
Multi-GPU plumbing: one process per GPU (torchrun), ``torch.distributed`` for the rendezvous.

The env-step path shards trivially -- envs are independent, rank ``r`` owns the contiguous global range
``[r * n_local, (r + 1) * n_local)`` and every env's counter-based stream is keyed by its GLOBAL index, so results do
not depend on the number of GPUs.  There is NO collective on the step path.  The only exchange is the optional
all-gather of per-rank episode statistics for logging (the reference logs per-process Monitor CSVs instead,
environments/utils.py:53-54): NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
import os


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard(total_envs, rank, world):
    """Contiguous shard of ``total_envs`` for ``rank``: (global_env_offset, n_local)."""
    base, rem = divmod(int(total_envs), int(world))
    n_local = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, n_local


def allgather_episode_stats(ep_ret_sum, ep_count, device=None):
    """Sum of episode returns and number of finished episodes over all ranks -> (mean return, episodes)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(ep_ret_sum), float(ep_count)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        parts = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(parts, t)
        t = torch.stack(parts).sum(0)
    s, c = t.tolist()
    return (s / c if c > 0 else 0.0), int(c)
