"""Multi-GPU host logic: independent inversions are sharded one image at a time across ranks.

There is no data-path collective: every rank owns a full generator replica and its own (ws, camera, targets).  The only
communication is the reduction of a few counters at the end (NCCL on GPUs; gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_indices(n_items, rank, world_size):
    """Image i goes to rank i mod world_size (scripts/run_pti.py processes images independently, single_id_coach.py:30-34)."""
    if not (0 <= rank < world_size):
        raise ValueError('rank out of range')
    return list(range(rank, n_items, world_size))


def reduce_run_stats(steps, elapsed_ms, loss_sum, device='cpu'):
    """All ranks -> (total steps, max elapsed ms, sum of losses).  Max-over-ranks time is the whole-job time."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return int(steps), float(elapsed_ms), float(loss_sum)
    s = torch.tensor([float(steps), float(loss_sum)], dtype=torch.float64, device=device)
    t = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=device)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(round(s[0].item())), float(t.item()), float(s[1].item())
