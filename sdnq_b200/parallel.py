"""Batch-sharded data parallelism for the quantized-Linear path.

The path has no cross-row interaction (activation scales are per row), so a diffusion batch shards by image with replicated
weights and no data-path collective; the only communication is one `all_gather` of the step output (NCCL over NVLink on the
GPU box, gloo in the CPU tests).  Works with any initialised `torch.distributed` backend."""
import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced split of `n_items` (images) over ranks: the first `n_items % world_size` ranks get one extra."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_batch(x: torch.Tensor, world_size: int | None = None, rank: int | None = None) -> torch.Tensor:
    """This rank's slice of a batch-first tensor."""
    world_size = dist.get_world_size() if world_size is None else world_size
    rank = dist.get_rank() if rank is None else rank
    lo, hi = shard_bounds(x.shape[0], world_size, rank)
    return x[lo:hi]


def gather_batch(y_local: torch.Tensor, n_items: int, group=None) -> torch.Tensor:
    """all_gather the per-rank outputs back into batch order (ragged shards are padded to the largest one for the collective)."""
    world = dist.get_world_size(group)
    sizes = [shard_bounds(n_items, world, r) for r in range(world)]
    longest = max(hi - lo for lo, hi in sizes)
    pad = longest - y_local.shape[0]
    if pad:
        y_local = torch.cat([y_local, y_local.new_zeros((pad, *y_local.shape[1:]))])
    parts = [torch.empty_like(y_local) for _ in range(world)]
    dist.all_gather(parts, y_local.contiguous(), group=group)
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)])


def max_over_ranks(value: float, device=None, group=None) -> float:
    """the timing rule of bench.py: a multi-GPU step takes as long as its slowest rank."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t[0])
