"""Multi-GPU plumbing (one process per GPU, ``torch.distributed``).

The sampling path shards *independent* units -- walkers / chains -- across ranks and needs no
collective on the data path: every rank evaluates its own rows and only scalar timings
or final chains are gathered.  Training is data-parallel over the rows of each optimiser batch with
ONE all-reduce of the flat gradient per step (``train.FusedTrainer.step``).
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_rows(n, rank, world_size):
    """Contiguous [lo, hi) slice of n independent rows owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def batch_shard(idx, rank, world_size):
    """This rank's rows of one optimiser batch (strided, so every rank sees the same shuffle)."""
    return idx[rank::world_size]


def max_over_ranks(value, device=None):
    """Max of a Python float over ranks (bench timing rule: max over ranks)."""
    rank, ws = world()
    if ws == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_rows(local, device=None):
    """Concatenate per-rank row blocks (variable length) on every rank -- final chain assembly."""
    rank, ws = world()
    if ws == 1:
        return local
    sizes = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(ws)]
    dist.all_gather(sizes, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device))
    mx = int(max(int(s.item()) for s in sizes))
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    parts = [torch.zeros_like(pad) for _ in range(ws)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:int(s.item())] for p, s in zip(parts, sizes)])
