"""Tile sharding for full-FOV reconstructions: tiles are independent units (own objF,
pupil, intensity stack), so rank r of W owns a contiguous block of tiles and the inner loop needs
no collective; only the final results are gathered to rank 0.  torch.distributed is plumbing."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_tiles: int, rank: int, world: int):
    """Contiguous, balanced (sizes differ by <= 1) block of tiles for `rank`."""
    base, rem = divmod(n_tiles, world)
    a = rank * base + min(rank, rem)
    return a, a + base + (1 if rank < rem else 0)


def tile_grid(frame_w: int, frame_h: int, Np: int, step: int | None = None):
    """(cropX, cropY) origins of the Np x Np tiles covering a frame (the reference reconstructs one
    tile per process from `cropX,cropY,cropSizeX`, fpmMain.cpp:519,532-533)."""
    step = step or Np
    return [(x, y) for y in range(0, frame_h - Np + 1, step) for x in range(0, frame_w - Np + 1, step)]


def gather_tiles(local: torch.Tensor, n_tiles: int, rank: int, world: int):
    """Final gather of per-tile results [n_local, ...] to rank 0 (None elsewhere)."""
    if world == 1:
        return local
    shape = tuple(local.shape[1:])
    if rank == 0:
        parts = [torch.empty((shard_range(n_tiles, r, world)[1] - shard_range(n_tiles, r, world)[0],) + shape,
                             dtype=local.dtype, device=local.device) for r in range(world)]
        parts[0] = local
        reqs = [dist.irecv(parts[r], src=r) for r in range(1, world) if parts[r].numel()]
        for q in reqs:
            q.wait()
        return torch.cat(parts, 0)
    if local.numel():
        dist.send(local.contiguous(), dst=0)
    return None


def max_over_ranks(value: float, device="cpu") -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
