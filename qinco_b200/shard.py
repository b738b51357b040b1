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


def reference_shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Rows [start, end) of rank `rank` exactly as `encode_database` cuts them (search_tasks.py:103-104): floor(n / world)
    rows per rank, the last rank also takes the remainder -- what the part files of task=encode must contain."""
    per = n // world
    return per * rank, (per * (rank + 1) if rank < world - 1 else n)


def _model_has_ivf(model) -> bool:
    """IVF-QINCo or not, for our model (`ivf_K`) and for anything with the reference's surface: the reference model
    carries it in its cfg (`cfg.ivf_in_use`, qinco_base.py:428) and has NO `ivf_K` attribute of its own."""
    if getattr(model, "ivf_K", 0):
        return True
    cfg = getattr(model, "cfg", None)
    if cfg is None:
        return False
    get = cfg.get if isinstance(cfg, dict) else (lambda k, d=None: getattr(cfg, k, d))
    try:
        return bool(get("ivf_in_use", False)) or (isinstance(cfg, dict) and bool(cfg.get("ivf_K")))
    except Exception:
        return False


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
    ivf = _model_has_ivf(model)
    M = None                      # uint8 code columns per vector (without the IVF code); known after the first batch
    parts = []
    for i0 in range(0, len(x_local), batch):
        xb = x_local[i0:i0 + batch]
        if ivf and hasattr(model, "encode_ivf_u8"):
            iv, codes, _ = model.encode_ivf_u8(xb, normalize=True, want_xhat=False)
        elif hasattr(model, "encode_u8") and not ivf:
            iv, (codes, _) = None, model.encode_u8(xb, normalize=True, want_xhat=False)
        else:
            c = model(xb, step="encode")                       # [M (+1), n] int64, row 0 = the IVF code of an IVF model
            rest = c[1:] if ivf else c
            # never truncate: a K > 256 model or an undetected IVF row must fail loudly, not wrap modulo 256
            if rest.numel() and (int(rest.min()) < 0 or int(rest.max()) > 255):
                raise ValueError("codes outside [0, 256): not a K <= 256 QINCo code matrix (is this an IVF model whose "
                                 "cfg does not say ivf_in_use?)")
            if ivf and c.shape[1] and (int(c[0].min()) < 0 or int(c[0].max()) >= 2 ** 31):
                raise ValueError("IVF codes do not fit int32")
            iv = c[0].to(torch.int32).contiguous() if ivf else None
            codes = rest.t().contiguous().to(torch.uint8)
        M = codes.shape[1] if M is None else M
        assert codes.shape[1] == M
        if ivf:      # little-endian bytes of the int32 IVF code as columns M .. M+3
            codes = torch.cat([codes, iv.contiguous().view(torch.uint8).reshape(-1, 4)], dim=1)
        parts.append(codes)
    if M is None:                 # no local rows: the width comes from the model (the reference's `model.M` is cfg._M_ivf,
        M = int(model.M) - (1 if (ivf and not hasattr(model, "ivf_K")) else 0)     # i.e. it counts the IVF row)
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
