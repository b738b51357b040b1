"""Mirror of the reference's v1 codec surface (reference qinco_v1/codec_qinco.py:25-46 `encode`, :54-72 `decode`)
and of the v1 model objects those functions drive (reference qinco_v1/model_qinco.py:74-130 `QINCo`, :185-234 `PQ_QINCo`).

    codes = encode(model, x_np, bs, is_float16)      # np.ndarray [N, M] int64
    x_hat = decode(model, codes_np, bs, is_float16)  # np.ndarray [N, D] float32

`model` is a `QINCoV1` or a `PQQINCoV1`.  The batch loop of the reference (H2D, model call, D2H, `.item()` per batch)
runs inside the C ABI instead (`qb_encode_host` / `qb_decode_host`: pinned staging, copies overlapped with compute);
`bs` therefore only names what the caller wants resident per step and does not change the result.
`is_float16` is accepted for signature compatibility: the kernels always use fp16 tensor-core operands with fp32
accumulation, and never cast codes to half (a latent bug of the reference, codec_qinco.py:62-63).
There is no CPU or generic-module fallback: any other model type is a TypeError.
"""
from __future__ import annotations

import time

import numpy as np
import torch

from . import synth
from .model import QINCo


def _sq_err(a: np.ndarray, b: np.ndarray) -> float:
    """sum((a - b)^2) in float64, in slabs (no full-size temporaries)."""
    tot = 0.0
    for i0 in range(0, len(a), 65536):
        d = a[i0:i0 + 65536].astype(np.float64) - b[i0:i0 + 65536]
        tot += float(np.einsum("ij,ij->", d, d))
    return tot


class QINCoV1:
    """v1-style model: `.encode(x [bs, D]) -> (codes [bs, M] int64, xhat [bs, D])`, `.decode(codes [bs, M])`, `.db_scale`.

    Same math as the v2 model in `qinco1_mode` with A=0, B=1 (reference qinco_v1/model_qinco.py:39-70, 91-118).
    """

    def __init__(self, state_dict=None, *, cfg=None, weights=None, db_scale=1.0, device="cuda:0"):
        if state_dict is not None:
            sd = {k: (v.detach().float().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v, np.float32))
                  for k, v in state_dict.items()}
            cfg, weights = synth.from_v1_state(sd)
        assert cfg is not None and weights is not None
        assert cfg["qinco1_mode"] and cfg["A"] == 0 and cfg["B"] == 1, "v1 models are qinco1_mode, A=0, B=1"
        self.db_scale = float(db_scale)
        w = dict(weights)
        w["data_mean"] = np.zeros(cfg["D"], np.float32)
        w["data_std"] = np.array(self.db_scale, np.float32)     # x / db_scale == (x - 0) / data_std
        self.cfg = cfg
        self.d, self.D, self.K, self.M, self.L, self.h = cfg["D"], cfg["D"], cfg["K"], cfg["M"], cfg["L"], cfg["dh"]
        self._m = QINCo(cfg, w, device=device)
        self.device = self._m.device

    def parameters(self):
        return self._m.parameters()

    def eval(self):
        return self

    def half(self):
        return self

    @torch.no_grad()
    def encode(self, x):
        """x [bs, D] already divided by db_scale -> (codes [bs, M] int64, xhat [bs, D])   (model_qinco.py:97-118)."""
        codes, xhat = self._m.encode_u8(x.float(), normalize=False, want_xhat=True)
        return codes.long(), xhat

    @torch.no_grad()
    def decode(self, codes):
        """codes [bs, M] -> xhat [bs, D] (normalised space)   (model_qinco.py:91-95).  Out-of-range codes surface as
        IndexError at the next `synchronize()` (QINCo._pack_codes)."""
        codes = torch.as_tensor(codes)
        assert codes.dim() == 2 and codes.shape[1] == self.M
        u8, _ = self._m._pack_codes(codes.t())
        return self._m.decode_u8(u8, denormalize=False)

    forward = encode

    def synchronize(self):
        self._m.synchronize()

    # ---- host-buffer protocol used by the module-level encode() / decode() ------------------------------------------
    def encode_host(self, x: np.ndarray):
        """data-space rows -> (codes int64 [N, M], summed squared error in data space)."""
        codes_u8, xhat = self._m._h.encode_host(x, normalize=True, want_xhat=True)      # xhat in normalised space
        s = np.float32(self.db_scale)
        return codes_u8.astype(np.int64), _sq_err(xhat, x / s) * self.db_scale ** 2

    def decode_host(self, codes: np.ndarray) -> np.ndarray:
        if codes.size and (codes.min() < 0 or codes.max() >= self.K):
            raise IndexError(f"codes out of range [0, {self.K})")
        return self._m._h.decode_host(codes.astype(np.uint8), denormalize=True)


class PQQINCoV1:
    """PQ-QINCo (reference qinco_v1/model_qinco.py:185-234): the (optionally OPQ-rotated) vector is cut into consecutive
    sub-vectors, each quantised by its own QINCo1 model in its own scale; the reconstruction is rotated back.

    Built here as a table of (column span, code span, sub-quantizer) triples fixed at construction: encode is one rotation,
    one kernel pass per span writing straight into the preallocated code / reconstruction matrices, one rotation back.
    """

    def __init__(self, sub_quantizers, opq_matrix=None):
        self.db_scale = 1                    # the scale lives in the sub-quantizers, like the reference
        self.sub_quantizers = list(sub_quantizers)
        self.device = self.sub_quantizers[0].device
        self._rot = None if opq_matrix is None else np.ascontiguousarray(np.asarray(opq_matrix, np.float32))
        self.opq_matrix = None if self._rot is None else torch.from_numpy(self._rot).to(self.device)
        dims = np.cumsum([0] + [q.d for q in self.sub_quantizers])
        cols = np.cumsum([0] + [q.M for q in self.sub_quantizers])
        self._spans = [(slice(int(dims[i]), int(dims[i + 1])), slice(int(cols[i]), int(cols[i + 1])), q)
                       for i, q in enumerate(self.sub_quantizers)]
        self.d = self.D = int(dims[-1])
        self.M = int(cols[-1])
        self.K = self.sub_quantizers[0].K

    def parameters(self):
        return self.sub_quantizers[0].parameters()

    def eval(self):
        return self

    def half(self):
        return self

    @torch.no_grad()
    def encode(self, x):
        """x [bs, D] -> (codes [bs, sum M] int64, xhat [bs, D])   (model_qinco.py:203-221)."""
        x = x.float().to(self.device)
        rotated = x if self.opq_matrix is None else x @ self.opq_matrix.T
        codes = torch.empty((len(x), self.M), dtype=torch.int64, device=self.device)
        recon = torch.empty_like(rotated)
        for span, cspan, q in self._spans:
            codes[:, cspan], part = q.encode((rotated[:, span] / q.db_scale).contiguous())
            recon[:, span] = part * q.db_scale
        return codes, (recon if self.opq_matrix is None else recon @ self.opq_matrix)

    @torch.no_grad()
    def decode(self, codes):
        """codes [bs, sum M] -> x [bs, D]   (model_qinco.py:223-234)."""
        codes = torch.as_tensor(codes).to(self.device)
        recon = torch.empty((len(codes), self.D), dtype=torch.float32, device=self.device)
        for span, cspan, q in self._spans:
            recon[:, span] = q.decode(codes[:, cspan]) * q.db_scale
        return recon if self.opq_matrix is None else recon @ self.opq_matrix

    forward = encode

    def synchronize(self):
        for q in self.sub_quantizers:
            q.synchronize()

    # ---- host-buffer protocol -------------------------------------------------------------------------------------------
    def encode_host(self, x: np.ndarray):
        rotated = x if self._rot is None else x @ self._rot.T
        codes = np.empty((len(x), self.M), np.int64)
        recon = np.empty_like(rotated)
        for span, cspan, q in self._spans:
            u8, part = q._m._h.encode_host(np.ascontiguousarray(rotated[:, span]), normalize=True, want_xhat=True)
            codes[:, cspan] = u8
            recon[:, span] = part * np.float32(q.db_scale)
        if self._rot is not None:
            recon = recon @ self._rot
        return codes, _sq_err(recon, x)

    def decode_host(self, codes: np.ndarray) -> np.ndarray:
        recon = np.empty((len(codes), self.D), np.float32)
        for span, cspan, q in self._spans:
            recon[:, span] = q.decode_host(np.ascontiguousarray(codes[:, cspan]))
        return recon if self._rot is None else recon @ self._rot


def _host_model(model):
    if not (hasattr(model, "encode_host") and hasattr(model, "decode_host")):
        raise TypeError(f"{type(model).__name__}: the B200 codec drives qinco_b200.codec.QINCoV1 / PQQINCoV1 models only "
                        "(wrap a v1 state dict with QINCoV1(state_dict, db_scale=...)); there is no generic-module fallback")
    return model


def encode(model, x, bs, is_float16=False, verbose=True):
    """numpy [N, D] float32 -> numpy [N, M] int64, reporting time and MSE like the reference (codec_qinco.py:25-46)."""
    t0 = time.time()
    x = np.ascontiguousarray(x, dtype=np.float32)
    codes, err_sum = _host_model(model).encode_host(x)
    MSE = err_sum / max(len(x), 1)
    if verbose:
        print(f"Encoding done in {time.time() - t0:.2f} s, {MSE=:g}")
    return codes


def decode(model, codes, bs, is_float16=False, verbose=True):
    """numpy [N, M] integer codes -> numpy [N, D] float32 in data space (codec_qinco.py:54-72)."""
    t0 = time.time()
    out = _host_model(model).decode_host(np.asarray(codes))
    if verbose:
        print(f"Decoding done in {time.time() - t0:.2f} s")
    return out
