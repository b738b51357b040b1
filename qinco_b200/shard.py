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
                   group=None):
    """Encode this rank's rows and return the codes of ALL rows as uint8 [n_total, M] (or just the local ones).

    `model` is a `qinco_b200.model.QINCo` (uses its uint8 fast path) or anything with the reference's call surface
    `model(x, step="encode") -> LongTensor [M, n]`.  `x_local` holds rows shard_range(n_total, rank, world).

    IVF-QINCo models (`model.ivf_K > 0`, code matrix [M + 1, n] with the IVF code in row 0) return the pair
    `(ivf_codes int32 [n_total], codes uint8 [n_total, M])`; the IVF codes ride in the same all-gather as four extra
    byte columns, so the path still has exactly one collective.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    start, end = shard_range(n_total, rank, world)
    assert x_local.shape[0] == end - start, f"rank {rank}: expected {end - start} local rows, got {x_local.shape[0]}"
    ivf = bool(getattr(model, "ivf_K", 0))
    M = int(model.M)
    parts = []
    for i0 in range(0, len(x_local), batch):
        xb = x_local[i0:i0 + batch]
        if ivf and hasattr(model, "encode_ivf_u8"):
            iv, codes, _ = model.encode_ivf_u8(xb, normalize=True, want_xhat=False)
        elif hasattr(model, "encode_u8") and not ivf:
            iv, (codes, _) = None, model.encode_u8(xb, normalize=True, want_xhat=False)
        else:
            c = model(xb, step="encode")                       # [M (+1), n] int64
            iv = c[0].to(torch.int32).contiguous() if ivf else None
            codes = (c[1:] if ivf else c).t().contiguous().to(torch.uint8)
        if ivf:      # little-endian bytes of the int32 IVF code as columns M .. M+3
            codes = torch.cat([codes, iv.contiguous().view(torch.uint8).reshape(-1, 4)], dim=1)
        parts.append(codes)
    width = M + (4 if ivf else 0)
    local = torch.cat(parts) if parts else torch.empty((0, width), dtype=torch.uint8, device=x_local.device)
    if gather and world > 1:
        per = -(-n_total // world)
        padded = torch.zeros((per, width), dtype=torch.uint8, device=local.device)
        padded[: len(local)] = local
        out = torch.empty((world * per, width), dtype=torch.uint8, device=local.device)
        dist.all_gather_into_tensor(out, padded, group=group)      # the only collective of the path
        local = out[:n_total]
    if not ivf:
        return local
    return local[:, M:].contiguous().view(torch.int32).reshape(-1), local[:, :M].contiguous()
