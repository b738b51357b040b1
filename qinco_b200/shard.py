"""Row-sharded bulk encode: the multi-GPU form of the reference's `encode_database`
(reference qinco/search/search_tasks.py:85-137: rank r encodes rows [r*floor(N/P), ...), writes `<out>.part_r.npz`, meets
the other ranks at a barrier).

Vectors are independent, so there is no data-path collective: every rank holds a full weight replica and encodes its own
contiguous rows; the final uint8 codes meet in ONE all-gather (8 bytes per vector for an 8x8 code) instead of on the
file system.  Works with any `torch.distributed` backend (NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Rows [start, end) of rank `rank`: contiguous, ceil(n / world) per rank, the tail ranks may get fewer or none."""
    per = -(-n // world) if world > 0 else n
    start = min(n, rank * per)
    return start, min(n, start + per)


@torch.no_grad()
def encode_sharded(model, x_local: torch.Tensor, n_total: int, batch: int = 1 << 20, gather: bool = True,
                   group=None) -> torch.Tensor:
    """Encode this rank's rows and return the codes of ALL rows as uint8 [n_total, M] (or just the local ones).

    `model` is a `qinco_b200.model.QINCo` (uses its uint8 fast path) or anything with the reference's call surface
    `model(x, step="encode") -> LongTensor [M, n]`.  `x_local` holds rows shard_range(n_total, rank, world).
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    start, end = shard_range(n_total, rank, world)
    assert x_local.shape[0] == end - start, f"rank {rank}: expected {end - start} local rows, got {x_local.shape[0]}"
    parts = []
    for i0 in range(0, len(x_local), batch):
        xb = x_local[i0:i0 + batch]
        if hasattr(model, "encode_u8"):
            codes, _ = model.encode_u8(xb, normalize=True, want_xhat=False)
        else:
            codes = model(xb, step="encode").t().contiguous().to(torch.uint8)
        parts.append(codes)
    M = int(model.M)
    local = torch.cat(parts) if parts else torch.empty((0, M), dtype=torch.uint8, device=x_local.device)
    if not gather or world == 1:
        return local
    per = -(-n_total // world)
    padded = torch.zeros((per, M), dtype=torch.uint8, device=local.device)
    padded[: len(local)] = local
    out = torch.empty((world * per, M), dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)      # the only collective of the path
    return out[:n_total]
