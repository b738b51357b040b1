"""Checkpoint and code-file formats of the reference, so released checkpoints run through the B200 kernels and the
outputs stay consumable by the reference's own build_index / search tasks.

* QINCo2 checkpoints: the dict written by `save_model` (reference qinco/utils.py:118-136): `model` (state dict),
  `parameters` {K, M, de, dh, L, A, B, ivf_in_use, ivf_K, qinco1_mode}, `data_dim`; read like `load_saved_model_data`
  (:140-172).  Legacy keys `residual_blocks.<i>.(in_proj|out_proj)` are rewritten like qinco/qinco_tasks.py:549-553 and
  the unused `steps.0.substep.codebook.weight` is dropped (:562-563).
* QINCo1 checkpoints: a pickled v1 `nn.Module` (qinco_v1/codec_qinco.py:112) or its state dict; keys `codebook0`,
  `step{m}.codebook`, `step{m}.MLPconcat`, `step{m}.residual_block{l}.{0,2}` (qinco_v1/model_qinco.py:28-37, 83-89).
* Encoded databases: `<out>.npz {n_parts, K, M, D}` + `<out>.part_<r>.npz {codes [n_r, M] int64}`
  (qinco/search/search_tasks.py:122-131, reader qinco/search/search_utils.py:33-78).
* v1 code files: `.npy [N, M]`, or `--raw`: M * ceil(log2 K) bits per vector, LSB first, rows padded to whole bytes —
  what `faiss.pack_bitstrings` / `unpack_bitstrings` produce (qinco_v1/codec_qinco.py:131-149; faiss is not a dependency
  here, the bit layout is restated).
"""
from __future__ import annotations

import math
import os
import re

import numpy as np

SAVED_PARAMETERS = ("K", "M", "de", "dh", "L", "A", "B", "ivf_in_use", "ivf_K", "qinco1_mode")   # qinco/utils.py:105-116


def _np(v):
    try:
        import torch
        if isinstance(v, torch.Tensor):
            return v.detach().cpu().numpy()
    except ImportError:
        pass
    return np.asarray(v)


def clean_v2_state_dict(sd: dict) -> dict:
    """Key fixes the reference applies to older checkpoints (qinco/qinco_tasks.py:549-563)."""
    out = {re.sub(r"residual_blocks.[0-9]+.(in_proj|out_proj)", r"\1", k): v for k, v in sd.items()}
    out.pop("steps.0.substep.codebook.weight", None)
    return out


def cfg_from_v2_checkpoint(ckpt: dict, overrides: dict | None = None) -> dict:
    """{D, M, K, L, de, dh, A, B, qinco1_mode} from a `save_model` dict; `overrides` win (like CLI arguments do in
    load_saved_model_data, which only fills parameters the user left unset; A > 0 over an A = 0 model is an error)."""
    overrides = {k: v for k, v in (overrides or {}).items() if v is not None}
    params = dict(ckpt.get("parameters") or {})
    if overrides.get("A", 0) and "A" in params and not params["A"]:
        raise ValueError("Can't evaluate a model trained with A=0 (no candidates pre-selection) using a non-zero A value.")
    params.update(overrides)
    sd = ckpt["model"] if "model" in ckpt else ckpt
    D = int(ckpt.get("data_dim") or _np(sd["steps.0.codebook.weight"]).shape[1])
    ivf = bool(params.get("ivf_in_use"))
    n_step = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("steps."))       # cfg._M_ivf
    M = params.get("M") or (n_step - 1 if ivf else n_step)
    K = params.get("K") or _np(sd["steps.0.codebook.weight"]).shape[0]
    L = params.get("L")
    if L is None:
        L = len({k.split(".")[3] for k in sd if k.startswith("steps.1.residual_blocks.")})
    de = params.get("de") or (_np(sd["steps.1.concat.mlp.weight"]).shape[0] if M > 1 else D)
    dh = params.get("dh") or (_np(sd["steps.1.residual_blocks.0.up_proj.weight"]).shape[0] if (M > 1 and L) else de)
    cfg = dict(D=D, M=int(M), K=int(K), L=int(L), de=int(de), dh=int(dh), A=int(params.get("A") or 0),
               B=int(params.get("B") or 1), qinco1_mode=bool(params.get("qinco1_mode") or False))
    if ivf:         # IVF-QINCo: the centroids are a separate file (cfg.ivf_centroids), see load_v2_checkpoint
        if not params.get("ivf_K"):
            raise ValueError("an IVF checkpoint needs ivf_K (or the centroid file)")
        cfg["ivf_K"] = int(params["ivf_K"])
    return cfg


def load_v2_checkpoint(path: str, overrides: dict | None = None, ivf_centroids: str | None = None):
    """-> (cfg dict, state dict of numpy arrays) for qinco_b200.model.QINCo.

    `ivf_centroids`: the .npy the reference takes as cfg.ivf_centroids ([ivf_K, D], already normalised); it becomes
    `steps.0.ivf_centroids.weight` like in qinco/qinco_tasks.py:565-569 and fixes ivf_K / D (qinco/utils.py:144-146)."""
    import torch
    ckpt = torch.load(str(path), map_location="cpu", weights_only=True)
    sd = clean_v2_state_dict(ckpt["model"] if "model" in ckpt else ckpt)
    sd = {k: _np(v) for k, v in sd.items()}
    overrides = dict(overrides or {})
    if ivf_centroids is not None:
        cent = np.load(ivf_centroids).astype(np.float32)
        sd["steps.0.ivf_centroids.weight"] = cent
        overrides.update(ivf_in_use=True, ivf_K=int(cent.shape[0]))
    cfg = cfg_from_v2_checkpoint(dict(ckpt, model=sd) if "model" in ckpt else sd, overrides)
    if cfg.get("ivf_K") and "steps.0.ivf_centroids.weight" not in sd:
        raise ValueError("IVF checkpoint: pass ivf_centroids=<npy> (the reference's cfg.ivf_centroids)")
    return cfg, sd


class _SafeV1Unpickler:
    """pickle-module stand-in for `torch.load(..., pickle_module=...)`: unpickles a whole v1 `nn.Module`
    (qinco_v1/codec_qinco.py:112 loads checkpoints that way) WITHOUT importing the reference and without executing
    arbitrary pickle code: the classes of reference qinco_v1/model_qinco.py map to inert `nn.Module` stubs, only torch /
    numpy / collections names are resolvable, everything else is refused."""
    import pickle as _pickle

    _V1_CLASSES = ("QINCo", "QINCoStep", "PQ_QINCo")
    _ALLOWED_PREFIXES = ("torch", "collections", "numpy")
    _ALLOWED_BUILTINS = ("set", "frozenset", "slice", "range", "complex", "bytearray", "list", "dict", "tuple", "int", "float", "bool", "str")

    class Unpickler(_pickle.Unpickler):
        def find_class(self, module, name):
            import builtins
            import importlib

            import torch
            if module in ("model_qinco", "qinco_v1.model_qinco") and name in _SafeV1Unpickler._V1_CLASSES:
                return type(name, (torch.nn.Module,), {"_qb_v1_kind": name})
            if module in ("builtins", "__builtin__") and name in _SafeV1Unpickler._ALLOWED_BUILTINS:
                return getattr(builtins, name)
            if module.split(".")[0] in _SafeV1Unpickler._ALLOWED_PREFIXES and not name.startswith("__"):
                return getattr(importlib.import_module(module), name)
            raise _SafeV1Unpickler._pickle.UnpicklingError(f"refusing to unpickle {module}.{name} from a v1 checkpoint")

    @staticmethod
    def load(f, **kw):
        return _SafeV1Unpickler.Unpickler(f, **kw).load()

    Pickler = _pickle.Pickler
    __name__ = "qinco_b200.io._SafeV1Unpickler"


def _load_v1_object(path: str, allow_pickled_module: bool):
    import torch
    try:
        return torch.load(str(path), map_location="cpu", weights_only=True)
    except Exception as e:
        if not allow_pickled_module:
            raise RuntimeError(
                f"{path} is not a plain state-dict file ({type(e).__name__}).  Pickled v1 nn.Module checkpoints (what the "
                "reference's codec loads) are only read on request: pass allow_pickled_module=True / --unsafe-pickle; they go "
                "through a restricted unpickler that maps model_qinco.{QINCo,QINCoStep,PQ_QINCo} to inert stubs") from e
    return torch.load(str(path), map_location="cpu", weights_only=False, pickle_module=_SafeV1Unpickler)


def load_v1_checkpoint(path: str, allow_pickled_module: bool = False):
    """-> (state dict of numpy arrays, db_scale) of a plain v1 QINCo from a state-dict file ({"state_dict", "db_scale"} or
    the bare dict) or, on request, a pickled v1 module.  PQ-QINCo and IVF v1 modules are not a single state dict: use
    `load_v1_model`, which says so."""
    obj = _load_v1_object(path, allow_pickled_module)
    if hasattr(obj, "state_dict"):
        kind = getattr(obj, "_qb_v1_kind", type(obj).__name__)
        if kind != "QINCo":
            raise ValueError(f"{path} holds a pickled {kind}; load_v1_checkpoint reads plain QINCo models (use load_v1_model)")
        sd, scale = {k: _np(v) for k, v in obj.state_dict().items()}, float(getattr(obj, "db_scale", 1.0))
    elif isinstance(obj, dict) and "state_dict" in obj:
        sd, scale = {k: _np(v) for k, v in obj["state_dict"].items()}, float(obj.get("db_scale", 1.0))
    else:
        sd, scale = {k: _np(v) for k, v in obj.items()}, 1.0
    if "codebook0.weight" not in sd:
        what = "a PQ-QINCo" if any(k.startswith("sub_quantizer_") for k in sd) else "an unsupported v1 model"
        raise ValueError(f"{path}: state dict of {what} (no codebook0.weight); use load_v1_model for PQ-QINCo")
    return sd, scale


def load_v1_model(path: str, device="cuda:0", allow_pickled_module: bool = False):
    """-> a ready `codec.QINCoV1`, or a `codec.PQQINCoV1` when the file holds a pickled PQ_QINCo (sub-quantizers with their
    own db_scale and an optional OPQ matrix, reference qinco_v1/model_qinco.py:185-201)."""
    from . import codec
    obj = _load_v1_object(path, allow_pickled_module)
    if hasattr(obj, "state_dict") and getattr(obj, "_qb_v1_kind", "") == "PQ_QINCo":
        subs = [codec.QINCoV1({k: _np(v) for k, v in q.state_dict().items()}, db_scale=float(getattr(q, "db_scale", 1.0)),
                              device=device) for q in obj.sub_quantizers]
        opq = getattr(obj, "opq_matrix", None)
        return codec.PQQINCoV1(subs, opq_matrix=None if opq is None else _np(opq))
    sd, scale = load_v1_checkpoint(path, allow_pickled_module)
    return codec.QINCoV1(sd, db_scale=scale, device=device)


# ---------------------------------------------------------------------------------------------- encoded databases (v2)
def save_encoded_db(output: str, parts, K: int, M: int, D: int) -> None:
    """`parts`: list of [n_r, M] integer arrays, one per rank (search_tasks.py:122-131).  `output` ends in .npz."""
    assert output.endswith(".npz")
    np.savez_compressed(output, n_parts=len(parts), K=K, M=M, D=D)
    for r, codes in enumerate(parts):
        np.savez_compressed(output[:-4] + f".part_{r}.npz", codes=np.asarray(codes, dtype=np.int64))


def load_encoded_db(base_path: str):
    """-> (codes [N, M] int64, info dict)   (search_utils.py:33-78, EncodedDBIterator.load_all)."""
    assert base_path.endswith(".npz")
    info = np.load(base_path)
    meta = dict(n_parts=int(info["n_parts"]), K=int(info["K"]), M=int(info["M"]), D=int(info["D"]))
    parts = [np.load(base_path[:-4] + f".part_{r}.npz")["codes"] for r in range(meta["n_parts"])]
    return np.concatenate(parts, axis=0), meta


# ------------------------------------------------------------------------------------------------- raw bit strings (v1)
def code_size_bytes(M: int, K: int) -> int:
    return (int(math.ceil(math.log2(K))) * M + 7) // 8      # codec_qinco.py:145


def pack_bitstrings(codes: np.ndarray, nbits: int) -> np.ndarray:
    """[N, M] non-negative ints -> [N, ceil(M*nbits/8)] uint8; value j occupies bits [j*nbits, (j+1)*nbits), LSB first."""
    codes = np.asarray(codes, dtype=np.uint64)
    n, M = codes.shape
    assert nbits >= 1 and (codes < (1 << nbits)).all(), "a code does not fit in nbits"
    bits = ((codes[:, :, None] >> np.arange(nbits, dtype=np.uint64)) & 1).astype(np.uint8).reshape(n, M * nbits)
    pad = (-bits.shape[1]) % 8
    if pad:
        bits = np.concatenate([bits, np.zeros((n, pad), np.uint8)], axis=1)
    return np.packbits(bits, axis=1, bitorder="little")


def unpack_bitstrings(packed: np.ndarray, nbits: int, M: int) -> np.ndarray:
    """Inverse of pack_bitstrings -> [N, M] int64."""
    packed = np.asarray(packed, dtype=np.uint8)
    n = packed.shape[0]
    bits = np.unpackbits(packed, axis=1, bitorder="little")[:, : M * nbits].reshape(n, M, nbits).astype(np.int64)
    return (bits << np.arange(nbits, dtype=np.int64)).sum(-1)


def write_raw_codes(path: str, codes: np.ndarray, K: int) -> None:
    pack_bitstrings(codes, int(math.ceil(math.log2(K)))).tofile(path)        # codec_qinco.py:131-137


def read_raw_codes(path: str, M: int, K: int) -> np.ndarray:
    nbits = int(math.ceil(math.log2(K)))
    packed = np.fromfile(path, dtype="uint8").reshape(-1, code_size_bytes(M, K))   # codec_qinco.py:141-148
    return unpack_bitstrings(packed, nbits, M)


# ------------------------------------------------------------------------------------------------------ vector files
def read_vectors(path: str) -> np.ndarray:
    """.npy, .fvecs or .bvecs -> float32 [N, D]   (codec_qinco.py:119-126; vecs layout: int32 dim header per vector)."""
    if path.endswith(".npy"):
        return np.load(path).astype(np.float32)
    if path.endswith(".fvecs"):
        a = np.fromfile(path, dtype=np.int32)
        d = int(a[0])
        return a.reshape(-1, d + 1)[:, 1:].copy().view(np.float32)
    if path.endswith(".bvecs"):
        a = np.fromfile(path, dtype=np.uint8)
        d = int(a[:4].view(np.int32)[0])
        return a.reshape(-1, d + 4)[:, 4:].astype(np.float32)
    raise RuntimeError("unrecognized format")
