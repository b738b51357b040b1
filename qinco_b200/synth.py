"""Seeded synthetic QINCo weights and data (numpy only).

The reference ships no checkpoints that are reachable offline and its default
initialisation zeroes `concat.mlp` and every `down_proj` (reference
qinco/model/qinco_base.py:66-69, 88-91), which turns the implicit-codebook MLP
into an identity.  Benchmarks and parity tests therefore use the weights made
here: residual k-means codebooks (same recipe as reference qinco/vrq.py:58-85,
a few Lloyd iterations per step) and dense Gaussian MLP weights scaled by
gain/sqrt(fan_in) so every layer contributes O(1) to the candidate.

State-dict key names are the reference's:
  steps.{m}.codebook.weight                      [K, D]
  steps.{m}.substep.codebook.weight              [K, D]   (m >= 1, A > 0)
  steps.{m}.concat.mlp.weight / .bias            [De, De+D] / [De]
  steps.{m}.residual_blocks.{l}.up_proj.weight   [Dh, De]
  steps.{m}.residual_blocks.{l}.down_proj.weight [De, Dh]
  steps.{m}.in_proj.weight / out_proj.weight     [De, D] / [D, De]  (De != D)
  data_mean [D], data_std []
"""
from __future__ import annotations

import hashlib

import numpy as np

# Model presets: reference config/model_args/{qinco1,qinco2-S,qinco2-M,qinco2-L}.yaml
PRESETS = {
    "qinco1": dict(L=16, de=None, dh=256, A=0, B=1, M=8, K=256, qinco1_mode=True),
    "qinco2-S": dict(L=2, de=128, dh=256, A=16, B=32, M=8, K=256, qinco1_mode=False),
    "qinco2-M": dict(L=4, de=384, dh=384, A=16, B=32, M=8, K=256, qinco1_mode=False),
    "qinco2-L": dict(L=16, de=384, dh=384, A=16, B=32, M=8, K=256, qinco1_mode=False),
}


def make_cfg(preset: str | None = None, *, D: int, **over) -> dict:
    """Model hyper-parameters as the reference's cfg names them (D,M,K,L,de,dh,A,B,qinco1_mode)."""
    cfg = dict(PRESETS[preset]) if preset else dict(L=2, de=None, dh=256, A=0, B=1, M=8, K=256, qinco1_mode=False)
    cfg.update(over)
    cfg["D"] = int(D)
    cfg["de"] = int(cfg["de"] or D)
    for k in ("L", "dh", "A", "B", "M", "K"):
        cfg[k] = int(cfg[k])
    cfg["qinco1_mode"] = bool(cfg["qinco1_mode"])
    if cfg.get("ivf_K"):        # IVF-QINCo: step 0 is an arg-min over ivf_K centroids, then M implicit-codebook steps
        cfg["ivf_K"] = int(cfg["ivf_K"])
    else:
        cfg.pop("ivf_K", None)
    return cfg


def make_data(n: int, D: int, seed: int = 1234, mean: float = 0.0, std: float = 1.0) -> np.ndarray:
    """iid Gaussian rows, the worst case for parity (incompressible, many near ties)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, D), dtype=np.float32)
    if std != 1.0:
        x *= np.float32(std)
    if mean != 0.0:
        x += np.float32(mean)
    return x


def _sqdist(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    return (a * a).sum(1)[:, None] + (b * b).sum(1)[None, :] - 2.0 * (a @ b.T)


def _kmeans(x: np.ndarray, K: int, iters: int, rng: np.random.Generator) -> np.ndarray:
    c = x[rng.choice(len(x), size=K, replace=len(x) < K)].copy()
    for _ in range(iters):
        a = _sqdist(x, c).argmin(1)
        for k in range(K):
            sel = a == k
            if sel.any():
                c[k] = x[sel].mean(0)
            else:
                c[k] = x[rng.integers(len(x))]
    return c


def make_weights(cfg: dict, seed: int = 4321, gain: float = 0.5, n_train: int = 8192,
                 kmeans_iters: int = 3, data_mean: float = 0.0, data_std: float = 1.0,
                 fp16_exact: bool = False) -> dict:
    """Synthetic state dict (numpy float32) for `cfg`.

    fp16_exact rounds the MLP matrices to fp16-representable values so the
    fp16-operand CUDA path and the fp32 oracle see identical weights.
    """
    D, M, K, L, De, Dh, A = (cfg[k] for k in ("D", "M", "K", "L", "de", "dh", "A"))
    ivf_K = int(cfg.get("ivf_K") or 0)
    n_steps = M + 1 if ivf_K else M          # reference cfg._M_ivf (qinco/qinco_tasks.py:378-383)
    rng = np.random.default_rng(seed)
    w: dict[str, np.ndarray] = {}

    def dense(out_f, in_f):
        m = (rng.standard_normal((out_f, in_f), dtype=np.float32) * np.float32(gain / np.sqrt(in_f)))
        if fp16_exact:
            m = m.astype(np.float16).astype(np.float32)
        return m

    # plain residual k-means for the explicit codebooks (and their pre-selection twins)
    xt = rng.standard_normal((n_train, D), dtype=np.float32)
    resid = xt
    for m in range(n_steps):
        if ivf_K and m == 0:                 # frozen IVF centroids (qinco_base.py:128-146), key as in IVFBook
            if ivf_K > 65536:                # billion-scale shape (2^20 centroids): k-means is pointless on synthetic data --
                cb = rng.standard_normal((ivf_K, D), dtype=np.float32) * np.float32(0.9)      # distinct Gaussian centroids
                best_d = np.full(len(resid), np.inf, np.float32)
                best_k = np.zeros(len(resid), np.int64)
                for k0 in range(0, ivf_K, 65536):                                             # arg-min in centroid slabs
                    d = _sqdist(resid, cb[k0:k0 + 65536])
                    a = d.argmin(1)
                    dm = d[np.arange(len(resid)), a]
                    upd = dm < best_d
                    best_d[upd], best_k[upd] = dm[upd], a[upd] + k0
                w["steps.0.ivf_centroids.weight"] = cb
                resid = resid - cb[best_k]
                continue
            cb = _kmeans(resid, ivf_K, kmeans_iters, rng).astype(np.float32)
            cb += rng.standard_normal(cb.shape, dtype=np.float32) * np.float32(0.02 * cb.std())
            w["steps.0.ivf_centroids.weight"] = cb
            resid = resid - cb[_sqdist(resid, cb).argmin(1)]
            continue
        cb = _kmeans(resid, K, kmeans_iters, rng).astype(np.float32)
        # empty-cluster re-seeding can duplicate a codeword, and exact duplicates are exact ties whose
        # winner is a topk implementation detail: jitter so all K codewords are distinct
        cb += rng.standard_normal(cb.shape, dtype=np.float32) * np.float32(0.02 * cb.std())
        w[f"steps.{m}.codebook.weight"] = cb
        a = _sqdist(resid, cb).argmin(1)
        resid = resid - cb[a]
        if m == 0:
            continue
        if A > 0:
            noise = rng.standard_normal(cb.shape, dtype=np.float32) * np.float32(0.05 * cb.std())
            w[f"steps.{m}.substep.codebook.weight"] = (cb + noise).astype(np.float32)
        w[f"steps.{m}.concat.mlp.weight"] = dense(De, De + D)
        w[f"steps.{m}.concat.mlp.bias"] = (rng.standard_normal(De, dtype=np.float32) * np.float32(0.05))
        for l in range(L):
            w[f"steps.{m}.residual_blocks.{l}.up_proj.weight"] = dense(Dh, De)
            w[f"steps.{m}.residual_blocks.{l}.down_proj.weight"] = dense(De, Dh)
        if De != D:
            w[f"steps.{m}.in_proj.weight"] = dense(De, D)
            w[f"steps.{m}.out_proj.weight"] = dense(D, De)
    w["data_mean"] = np.full((D,), data_mean, dtype=np.float32)
    w["data_std"] = np.array(data_std, dtype=np.float32)
    return w


def weights_digest(w: dict) -> str:
    """sha256 over the sorted tensors, to pin regenerated weights in fixtures."""
    h = hashlib.sha256()
    for k in sorted(w):
        h.update(k.encode())
        h.update(np.ascontiguousarray(w[k], dtype=np.float32).tobytes())
    return h.hexdigest()


def to_v1_state(cfg: dict, w: dict) -> dict:
    """Re-key a QINCo1-mode state dict to the v1 naming (reference qinco_v1/model_qinco.py:28-37,83-89)."""
    assert cfg["qinco1_mode"] and cfg["de"] == cfg["D"] and cfg["A"] == 0
    out = {"codebook0.weight": w["steps.0.codebook.weight"]}
    for m in range(1, cfg["M"]):
        out[f"step{m}.codebook.weight"] = w[f"steps.{m}.codebook.weight"]
        out[f"step{m}.MLPconcat.weight"] = w[f"steps.{m}.concat.mlp.weight"]
        out[f"step{m}.MLPconcat.bias"] = w[f"steps.{m}.concat.mlp.bias"]
        for l in range(cfg["L"]):
            out[f"step{m}.residual_block{l}.0.weight"] = w[f"steps.{m}.residual_blocks.{l}.up_proj.weight"]
            out[f"step{m}.residual_block{l}.2.weight"] = w[f"steps.{m}.residual_blocks.{l}.down_proj.weight"]
    return out


def from_v1_state(sd: dict) -> tuple[dict, dict]:
    """Inverse of to_v1_state: (cfg, v2-keyed weights) from a v1 state dict."""
    g = {k: np.asarray(v, dtype=np.float32) for k, v in sd.items()}
    K, D = g["codebook0.weight"].shape
    M = 1 + len({k.split(".")[0] for k in g if k.startswith("step")})
    L = len({k.split(".")[1] for k in g if k.startswith("step1.residual_block")}) if M > 1 else 0
    Dh = g["step1.residual_block0.0.weight"].shape[0] if (M > 1 and L > 0) else D
    cfg = make_cfg(None, D=D, M=M, K=K, L=L, de=D, dh=Dh, A=0, B=1, qinco1_mode=True)
    w = {"steps.0.codebook.weight": g["codebook0.weight"]}
    for m in range(1, M):
        w[f"steps.{m}.codebook.weight"] = g[f"step{m}.codebook.weight"]
        w[f"steps.{m}.concat.mlp.weight"] = g[f"step{m}.MLPconcat.weight"]
        w[f"steps.{m}.concat.mlp.bias"] = g[f"step{m}.MLPconcat.bias"]
        for l in range(L):
            w[f"steps.{m}.residual_blocks.{l}.up_proj.weight"] = g[f"step{m}.residual_block{l}.0.weight"]
            w[f"steps.{m}.residual_blocks.{l}.down_proj.weight"] = g[f"step{m}.residual_block{l}.2.weight"]
    w["data_mean"] = np.zeros((D,), np.float32)
    w["data_std"] = np.array(1.0, np.float32)
    return cfg, w


def make_pairwise_tables(D, M, K, Mt, ivf_K, seed):
    """Synthetic tables of the shapes PairwiseDecoderIVF holds after training (pairwise_decoder.py:73-86)."""
    rng = np.random.default_rng(seed)
    book = rng.standard_normal((Mt, K * K, D)).astype(np.float32)
    comb = np.stack([rng.integers(0, M + 5, Mt), rng.integers(0, M + 5, Mt)]).astype(np.int64)
    comb[:, 0] = (0, min(1, M + 4))
    if Mt > 1:
        comb[:, 1] = (M, M + 4)      # a pair made of IVF-derived codes only
    imap = rng.integers(0, K, (ivf_K, 5)).astype(np.int64)
    return book, comb, imap
