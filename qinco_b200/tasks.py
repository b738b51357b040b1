"""`task=eval` / `task=eval_time` / `task=encode` of the reference without Hydra / accelerate.

    python -m qinco_b200.tasks eval      --model ckpt.pt --db vectors.npy [--batch 1024] [--A a] [--B b]
    python -m qinco_b200.tasks eval_time --model ckpt.pt --db vectors.npy
    torchrun --nproc-per-node N -m qinco_b200.tasks encode --model ckpt.pt --db vectors.npy --output codes.npz

* eval / eval_time follow `compute_MSE` (reference qinco/qinco_tasks.py:87-148): up to 10 warm-up batches, then per batch
  a timed `model(batch, step="encode")` and a timed `model(codes, step="decode")`, each closed by a forced host read of
  the last element; the MSE is mean_n sum_d (x - xhat)^2 (qinco/utils.py:95); the report lines have the reference's shape
  ("Test metrics: [[MSE=...]]", the code-word entropy line, and for eval_time the three timing lines in μs per vector).
* encode follows `encode_database` (reference qinco/search/search_tasks.py:85-137): one process per GPU (torchrun), rank r
  encodes rows [r * floor(N/P), ...) (the last rank takes the remainder) in batches of cfg.batch, writes
  `<output minus .npz>.part_<r>.npz {codes [n_r, M_ivf] int64}`, rank 0 writes `<output> {n_parts, K, M, D}`, with the
  reference's barriers in between.  `--gather` additionally meets the uint8 codes of all ranks in ONE all-gather
  (qinco_b200.shard.encode_sharded), which is what the bench times.

`--synthetic WORKLOAD` (a bench.py workload name) replaces --model / --db with the seeded synthetic model and data the
benchmarks use (there are no released checkpoints offline).  The model is `qinco_b200.model.QINCo`: no CPU fallback.
"""
from __future__ import annotations

import argparse
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from . import io, shard


class Timer:
    """Accumulating wall-clock timer used as a context manager (the role of qinco/metrics.py:182-230)."""

    def __init__(self):
        self.elapsed = 0.0

    def __enter__(self):
        self._t0 = time.time()
        return self

    def __exit__(self, *exc):
        self.elapsed += time.time() - self._t0

    def __str__(self):
        return f"{self.elapsed:.2f}s"


def code_entropy(counts: np.ndarray) -> np.ndarray:
    """Entropy in bits of every code position from its usage histogram [positions, K]."""
    p = counts / np.maximum(counts.sum(1, keepdims=True), 1)
    with np.errstate(divide="ignore", invalid="ignore"):
        return -np.where(p > 0, p * np.log2(p), 0.0).sum(1)


def load_model(args, device):
    """(model, vectors) from --model / --db, or from --synthetic."""
    from .model import QINCo
    if args.synthetic:
        import importlib
        import sys
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        if root not in sys.path:
            sys.path.insert(0, root)
        bench = importlib.import_module("bench")
        wl = dict(bench.WORKLOADS[args.synthetic])
        wl["cfg"] = dict(wl["cfg"], **{k: v for k, v in (("A", args.A), ("B", args.B)) if v is not None})
        cfg, w, x = bench.make_model_inputs(wl, args.limit or 10000, 0)
        return QINCo(cfg, w, device=device), x.numpy()
    cfg, sd = io.load_v2_checkpoint(args.model, dict(A=args.A, B=args.B), ivf_centroids=args.ivf_centroids)
    x = io.read_vectors(args.db)
    if args.limit:
        x = x[: args.limit]
    return QINCo(cfg, sd, device=device), x


@torch.no_grad()
def compute_MSE(model, vecs: np.ndarray, batch: int, device, timed: bool, out=print) -> dict:
    """The evaluation loop of the reference (qinco_tasks.py:87-148) on any model with its call surface."""
    batches = [torch.from_numpy(np.ascontiguousarray(vecs[i:i + batch], dtype=np.float32)).to(device)
               for i in range(0, len(vecs), batch)]
    decoded = None
    for i_batch, b in enumerate(batches):                       # warm start (:96-104)
        decoded = model(model(b, step="encode"), step="decode")
        if i_batch >= 10:
            break
    if decoded is not None:
        _ = decoded[-1][-1].item()
        out(f"Warm-start with {i_batch} batches: done")
    t_encode, t_decode = Timer(), Timer()
    n_vecs, sq_err = 0, 0.0
    usage = None
    for b in batches:
        n_vecs += len(b)
        with t_encode:
            codes = model(b, step="encode")
            _ = float(codes[-1].reshape(-1)[-1].cpu())          # completes the computation inside the timer (:112-115)
        with t_decode:
            xhat = model(codes, step="decode")
            _ = float(xhat.reshape(-1)[-1].cpu())
        assert xhat.shape == b.shape, f"{xhat.shape=} != {b.shape=}"
        sq_err += float(((b - xhat).double() ** 2).sum())
        c = codes[-model.M:].cpu().numpy()                       # the K-ary code positions (an IVF row 0 is left out)
        if usage is None:
            usage = np.zeros((c.shape[0], model.K), np.int64)
        for m in range(c.shape[0]):
            usage[m] += np.bincount(c[m], minlength=model.K)
    if hasattr(model, "synchronize"):
        model.synchronize()
    res = dict(MSE=sq_err / max(n_vecs, 1), n_vecs=n_vecs, t_encode=t_encode.elapsed, t_decode=t_decode.elapsed)
    out("Test metrics: [[" + f"MSE={res['MSE']:g}" + "]]")
    if usage is not None:
        ent = code_entropy(usage)
        out(f"test_codeword_entropy={ent.mean():g} (min={ent.min():g})")
        res["entropy"] = float(ent.mean())
    if timed and n_vecs:
        out(f"Encoding time: {t_encode} | Decoding time: {t_decode}")
        out(f"Encoding time / vector: {t_encode.elapsed / n_vecs * 1e6:.1f}μs")
        out(f"Decoding time / vector: {t_decode.elapsed / n_vecs * 1e6:.1f}μs")
    return res


@torch.no_grad()
def encode_database(model, db_vecs: np.ndarray, output: str, batch: int, device, gather: bool = False, out=print):
    """The reference's bulk encode (search_tasks.py:85-137): per-rank part files + the header file; returns this rank's
    codes [n_r, M_ivf] int64 (and, with gather, the uint8 codes of ALL rows from the one all-gather)."""
    assert output.endswith(".npz")
    base = output[:-4]
    on = dist.is_available() and dist.is_initialized()
    nproc = dist.get_world_size() if on else 1
    rank = dist.get_rank() if on else 0
    say = out if rank == 0 else (lambda *a, **k: None)
    barrier = dist.barrier if on else (lambda: None)
    barrier()
    db_size = len(db_vecs)
    say(f"Encoding {db_size} vectors using {nproc} processes")
    start, end = shard.reference_shard_range(db_size, rank, nproc)
    t_enc, t_save = Timer(), Timer()
    parts = []
    with t_enc:
        for i0 in range(start, end, batch):
            b = torch.from_numpy(np.ascontiguousarray(db_vecs[i0:min(end, i0 + batch)], dtype=np.float32)).to(device)
            parts.append(model(b, step="encode").T.cpu().numpy())
        barrier()
    say(f"Encoding done in {t_enc}")
    M_ivf = int(getattr(model, "M_ivf", getattr(model, "M", 0)))
    codes = np.concatenate(parts) if parts else np.zeros((0, M_ivf), np.int64)
    with t_save:
        if rank == 0:
            np.savez_compressed(output, n_parts=nproc, K=model.K, M=getattr(model, "M", M_ivf), D=model.D)
        np.savez_compressed(base + f".part_{rank}.npz", codes=codes)
        barrier()
    say(f"Stored codes into {output} and {nproc} associated part files [done in {t_save}]")
    if not gather:
        return codes
    # the single-collective form: uint8 codes of every rank in one all-gather (needs the ceil(N/P) row split)
    s, e = shard.shard_range(db_size, rank, nproc)
    x_local = torch.from_numpy(np.ascontiguousarray(db_vecs[s:e], dtype=np.float32)).to(device)
    return codes, shard.encode_sharded(model, x_local, db_size)


def main(argv=None):
    ap = argparse.ArgumentParser(prog="qinco_b200.tasks")
    ap.add_argument("task", choices=["eval", "eval_time", "encode"])
    ap.add_argument("--model", default=None, help="QINCo2 checkpoint (the dict written by the reference's save_model)")
    ap.add_argument("--db", default=None, help="vectors: .npy / .fvecs / .bvecs")
    ap.add_argument("--ivf_centroids", default=None)
    ap.add_argument("--output", default=None, help="encode: <name>.npz")
    ap.add_argument("--batch", type=int, default=1024, help="vectors per model call (cfg.batch of the reference)")
    ap.add_argument("--A", type=int, default=None)
    ap.add_argument("--B", type=int, default=None)
    ap.add_argument("--limit", type=int, default=0, help="use only the first LIMIT vectors (torchrun swallows a bare --n)")
    ap.add_argument("--synthetic", default=None, help="a bench.py workload name instead of --model / --db")
    ap.add_argument("--gather", action="store_true", help="encode: also all-gather the uint8 codes (one NCCL collective)")
    args = ap.parse_args(argv)
    if not args.synthetic and not (args.model and args.db):
        ap.error("--model and --db (or --synthetic) are required")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("qinco_b200.tasks needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    model, vecs = load_model(args, device)
    rank = dist.get_rank() if world > 1 else 0
    if rank == 0:
        print(f"Test set: {vecs.shape}" if args.task != "encode" else f"Database: {vecs.shape}")
    if args.task == "encode":
        assert args.output and args.output.endswith(".npz"), "encode needs --output <name>.npz"
        encode_database(model, vecs, args.output, args.batch, device, gather=args.gather)
    else:
        compute_MSE(model, vecs, args.batch, device, timed=args.task == "eval_time",
                    out=print if rank == 0 else (lambda *a, **k: None))
    model.synchronize()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
