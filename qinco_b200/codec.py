"""Mirror of the reference's v1 codec surface (reference qinco_v1/codec_qinco.py:25-46 `encode`, :54-72 `decode`)
and of the v1 model object those functions drive (reference qinco_v1/model_qinco.py:74-130 `QINCo`).

    codes = encode(model, x_np, bs, is_float16)      # np.ndarray [N, M] int64
    x_hat = decode(model, codes_np, bs, is_float16)  # np.ndarray [N, D] float32

With a `QINCoV1` model the batch loop (H2D, encode, D2H per batch) runs inside the C ABI (`qb_encode_host` /
`qb_decode_host`: pinned staging, copies overlapped with compute) instead of a Python loop with a `.item()` sync per
batch; `bs` then only bounds what the caller wants resident per step and does not change the result.
`is_float16` is accepted for signature compatibility: the kernels always use fp16 tensor-core operands with fp32
accumulation, and never cast codes to half (a latent bug of the reference, codec_qinco.py:62-63).
"""
from __future__ import annotations

import time

import numpy as np
import torch

from . import synth
from .model import QINCo


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
        """codes [bs, M] -> xhat [bs, D] (normalised space)   (model_qinco.py:91-95)."""
        codes = torch.as_tensor(codes).to(self.device)
        assert codes.dim() == 2 and codes.shape[1] == self.M
        if codes.numel() and (int(codes.min()) < 0 or int(codes.max()) >= self.K):
            raise IndexError(f"codes out of range [0, {self.K})")
        return self._m.decode_u8(codes.to(torch.uint8).contiguous(), denormalize=False)

    forward = encode


class PQQINCoV1:
    """PQ-QINCo (reference qinco_v1/model_qinco.py:185-234): the vector is cut into consecutive sub-vectors, each with
    its own QINCo1 quantizer and db_scale; an optional OPQ rotation is applied before / undone after.  A thin loop over
    `QINCoV1` sub-quantizers: all quantisation work stays in the CUDA kernels, the rotation is one torch matmul."""

    def __init__(self, sub_quantizers, opq_matrix=None):
        self.db_scale = 1                    # set per sub-quantizer, like the reference
        self.sub_quantizers = list(sub_quantizers)
        self.device = self.sub_quantizers[0].device
        self.opq_matrix = None if opq_matrix is None else torch.as_tensor(np.asarray(opq_matrix, np.float32)).to(self.device)
        self.d = self.D = sum(q.d for q in self.sub_quantizers)
        self.M = sum(q.M for q in self.sub_quantizers)
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
        d0, codes, xhat = 0, [], torch.zeros_like(x)
        if self.opq_matrix is not None:
            x = x @ self.opq_matrix.T
        for q in self.sub_quantizers:
            d1 = d0 + q.d
            code, xhat_sub = q.encode((x[:, d0:d1] / q.db_scale).contiguous())
            codes.append(code)
            xhat[:, d0:d1] = xhat_sub * q.db_scale
            d0 = d1
        if self.opq_matrix is not None:
            xhat = xhat @ self.opq_matrix
        return torch.cat(codes, 1), xhat

    @torch.no_grad()
    def decode(self, codes):
        """codes [bs, sum M] -> x [bs, D]   (model_qinco.py:223-234)."""
        codes = torch.as_tensor(codes).to(self.device)
        c0, xs = 0, []
        for q in self.sub_quantizers:
            c1 = c0 + q.M
            xs.append(q.decode(codes[:, c0:c1].contiguous()) * q.db_scale)
            c0 = c1
        x = torch.cat(xs, 1)
        if self.opq_matrix is not None:
            x = x @ self.opq_matrix
        return x

    forward = encode


def encode(model, x, bs, is_float16=False, verbose=True):
    """numpy [N, D] float32 -> numpy [N, M] int64; prints the reference's progress/MSE lines (codec_qinco.py:25-46)."""
    t0 = time.time()
    x = np.ascontiguousarray(x, dtype=np.float32)
    if isinstance(model, QINCoV1):
        codes_u8, xhat = model._m._h.encode_host(x, normalize=True, want_xhat=True)   # xhat in normalised space
        s = np.float32(model.db_scale)
        err_sum = 0.0
        for i0 in range(0, len(x), 65536):
            d = xhat[i0:i0 + 65536] - x[i0:i0 + 65536] / s
            err_sum += float(np.einsum("ij,ij->", d, d, dtype=np.float64)) * model.db_scale ** 2
        codes = codes_u8.astype(np.int64)
    else:   # any object with the v1 duck type: the reference's own loop
        output, err_sum = [], 0.0
        device = next(model.parameters()).device
        with torch.no_grad():
            for i0 in range(0, len(x), bs):
                batch = torch.from_numpy(x[i0:i0 + bs]).to(device) / model.db_scale
                c, recons = model.encode(batch)
                err_sum += ((recons - batch) ** 2).sum().item() * model.db_scale ** 2
                output.append(c.cpu().numpy())
        codes = np.concatenate(output) if output else np.zeros((0, model.M), np.int64)
    MSE = err_sum / max(len(x), 1)
    if verbose:
        print(f"Encoding done in {time.time() - t0:.2f} s, {MSE=:g}")
    return codes


def decode(model, codes, bs, is_float16=False, verbose=True):
    """numpy [N, M] integer codes -> numpy [N, D] float32 in data space (codec_qinco.py:54-72)."""
    t0 = time.time()
    codes = np.asarray(codes)
    if isinstance(model, QINCoV1):
        if codes.size and (codes.min() < 0 or codes.max() >= model.K):
            raise IndexError(f"codes out of range [0, {model.K})")
        out = model._m._h.decode_host(codes.astype(np.uint8), denormalize=True)
    else:
        output = []
        device = next(model.parameters()).device
        with torch.no_grad():
            for i0 in range(0, len(codes), bs):
                batch = torch.from_numpy(np.ascontiguousarray(codes[i0:i0 + bs])).to(device)
                output.append((model.decode(batch) * model.db_scale).cpu().numpy())
        out = np.concatenate(output) if output else np.zeros((0, model.D), np.float32)
    if verbose:
        print(f"Decoding done in {time.time() - t0:.2f} s")
    return out
