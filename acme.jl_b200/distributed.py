"""Multi-GPU: instances are independent, so the batch shards by contiguous instance
ranges (one process per GPU) with NO collective on the data path.  The only
collective is the optional final gather of the output shards (NCCL over NVLink on
GPUs; the same code runs on gloo/CPU tensors for the host-logic tests).

SURVEY.md section 8(e): ranges [g*B/G, (g+1)*B/G); U/Y are instance-slowest, so a
shard is one contiguous block of the (B, N, ny) tensor.
"""
from __future__ import annotations

from typing import Optional, Tuple


def shard_range(batch: int, world: int, rank: int) -> Tuple[int, int]:
    """(first, count) of rank's contiguous shard; the first `batch % world` ranks get one extra."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(batch, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


class ShardedBatchRunner:
    """This rank's shard of a batch of `batch` instances.  Per-instance arrays (params,
    overrides, init_z) are given for the FULL batch on every rank; the library copies
    only this rank's slice to its GPU."""

    def __init__(self, model, batch: int, *, rank: Optional[int] = None, world: Optional[int] = None, **kw):
        import torch.distributed as dist
        from .runner import BatchRunner
        if world is None:
            world = dist.get_world_size() if dist.is_initialized() else 1
        if rank is None:
            rank = dist.get_rank() if dist.is_initialized() else 0
        self.batch_total, self.rank, self.world = batch, rank, world
        self.first, self.count = shard_range(batch, world, rank)
        self.runner = BatchRunner(model, batch, first=self.first, count=self.count, **kw)

    def run(self, u_local, y_local=None, **kw):
        """u_local: this rank's (count, N, nu) device tensor (or a shared (N, nu) one)."""
        return self.runner.run(u_local, y_local, **kw)

    def gather(self, y_local, group=None, layout: str = "instance"):
        return gather_outputs(y_local, self.batch_total, group=group, layout=layout)


def gather_outputs(y_local, batch_total: int, group=None, layout: str = "instance"):
    """all-gather the (count_r, N, ny) output shards into the full (B, N, ny) tensor on every
    rank.  Uneven shards are padded to the largest one for the collective and trimmed after.
    ``layout="sample"``: sample-major shards (N, count_r, ny) -> (N, B, ny) (the shards interleave
    in memory, so the gathered blocks are re-assembled along the instance axis)."""
    import torch
    import torch.distributed as dist
    if layout not in ("instance", "sample"):
        raise ValueError("layout must be 'instance' or 'sample'")
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return y_local
    world = dist.get_world_size(group)
    counts = [shard_range(batch_total, world, r)[1] for r in range(world)]
    cmax = max(counts)
    if layout == "sample":
        N, cnt, ny = y_local.shape
        if cnt != counts[dist.get_rank(group)]:
            raise ValueError("y_local does not have this rank's shard size")
        padded = y_local
        if cnt != cmax:
            padded = torch.zeros((N, cmax, ny), dtype=y_local.dtype, device=y_local.device)
            padded[:, :cnt] = y_local
        out = torch.empty((world * N, cmax, ny), dtype=y_local.dtype, device=y_local.device)
        dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
        out = out.view(world, N, cmax, ny)
        return torch.cat([out[r, :, :counts[r]] for r in range(world)], dim=1)
    tail = tuple(y_local.shape[1:])
    if y_local.shape[0] != counts[dist.get_rank(group)]:
        raise ValueError("y_local does not have this rank's shard size")
    padded = y_local
    if y_local.shape[0] != cmax:
        padded = torch.zeros((cmax,) + tail, dtype=y_local.dtype, device=y_local.device)
        padded[:y_local.shape[0]] = y_local
    out = torch.empty((world * cmax,) + tail, dtype=y_local.dtype, device=y_local.device)
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    if all(c == cmax for c in counts):
        return out
    return torch.cat([out[r * cmax:r * cmax + counts[r]] for r in range(world)], dim=0)
