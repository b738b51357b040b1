"""Pairwise additive decoder, forward only — the re-ranking decoder of the reference's IVF search pipeline.

Mirror of `PairwiseDecoderIVF.forward(codes_MB, ivf_codes)` (reference qinco/search/pairwise_decoder.py:88-93 with
`map_codes` :126-130; caller qinco/search/search_tasks.py:448-471): every vector is the sum of `Mt` rows, one from each
of `Mt` tables of `K*K` rows indexed by a PAIR of its small codes (the QINCo codes plus 5 codes derived from its IVF
centroid).  Training of the tables (pairwise_decoder.py:132-205) is out of scope; the tables come from the reference's
own state dict (`codebook_MKD`, `combine_mvals_m`, `ivf_code_map`).

The work is one HBM-bound gather-accumulate CUDA kernel behind the C ABI (`qb_pairwise_*`, include/qinco_b200.h); the
additions happen in the reference's order, so the output is bit-identical to the PyTorch fp32 path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

IVF_M = 5   # PairwiseDecoderIVF.IVF_M


class PairwiseDecoderIVF:
    def __init__(self, state_dict, *, K: int, M: int, device="cuda:0"):
        """state_dict: `codebook_MKD [Mt, K*K, D]`, `combine_mvals_m [2, Mt]`, `ivf_code_map [ivf_K, 5]` (tensors or arrays);
        K, M: the base model's codebook size and codes per vector (cfg.K, cfg.M)."""
        def arr(k, dt):
            v = state_dict[k]
            v = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
            return np.ascontiguousarray(v, dtype=dt)
        book = arr("codebook_MKD", np.float32)
        comb = arr("combine_mvals_m", np.int64)
        imap = arr("ivf_code_map", np.int64)
        Mt, K2, D = book.shape
        if K2 != K * K:
            raise ValueError(f"codebook_MKD has {K2} rows per table, expected K*K = {K * K}")
        if comb.shape != (2, Mt) or imap.ndim != 2 or imap.shape[1] != IVF_M:
            raise ValueError("combine_mvals_m must be [2, Mt] and ivf_code_map [ivf_K, 5]")
        self.K_base, self.M_base, self.D = int(K), int(M), int(D)
        self.M, self.K = int(Mt), int(K2)                  # the reference renames these after pre_train_init
        self.ivf_K = int(imap.shape[0])
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("qinco_b200 has no CPU fallback: the pairwise decoder needs a CUDA (sm_100a) device")
        lib = _lib.load()
        d = _lib.QbPairwiseDesc(D=self.D, M=self.M_base, K=self.K_base, Mt=self.M, ivf_K=self.ivf_K,
                                device=self.device.index or 0,
                                codebook=book.ctypes.data_as(C.POINTER(C.c_float)),
                                combine=comb.ctypes.data_as(C.POINTER(C.c_int64)),
                                ivf_code_map=imap.ctypes.data_as(C.POINTER(C.c_int64)))
        h = C.c_void_p()
        rc = lib.qb_pairwise_create(C.byref(d), C.byref(h))
        if rc != 0:
            raise _lib.QbError(rc, lib.qb_pairwise_last_error().decode(errors="replace"))
        self._lib, self._h = lib, h
        self.table_bytes = book.nbytes

    def close(self):
        if getattr(self, "_h", None):
            self._lib.qb_pairwise_destroy(self._h)
            self._h = None

    __del__ = close

    def eval(self):
        return self

    def decode_u8(self, codes_u8: torch.Tensor, ivf_codes_i32: torch.Tensor) -> torch.Tensor:
        """codes [n, M_base] uint8 and IVF codes [n] int32 on the device -> [n, D] fp32; asynchronous."""
        assert codes_u8.dtype == torch.uint8 and codes_u8.dim() == 2 and codes_u8.shape[1] == self.M_base
        assert ivf_codes_i32.dtype == torch.int32 and ivf_codes_i32.shape == (codes_u8.shape[0],)
        codes_u8, ivf_codes_i32 = codes_u8.contiguous(), ivf_codes_i32.contiguous()
        n = codes_u8.shape[0]
        out = torch.empty((n, self.D), dtype=torch.float32, device=self.device)
        if n:
            with torch.cuda.device(self.device):
                rc = self._lib.qb_pairwise_decode(self._h, codes_u8.data_ptr(), ivf_codes_i32.data_ptr(), n, out.data_ptr(),
                                                  torch.cuda.current_stream().cuda_stream)
            if rc != 0:
                raise _lib.QbError(rc, self._lib.qb_pairwise_last_error().decode(errors="replace"))
        return out

    @torch.no_grad()
    def forward(self, codes_MB, ivf_codes=None):
        """codes_MB [M_base, n] integer, ivf_codes [n] integer -> xhat [n, D] fp32 (pairwise_decoder.py:88-93)."""
        assert ivf_codes is not None and ivf_codes.dim() == 1, "map_codes needs the IVF codes (pairwise_decoder.py:127)"
        codes_MB = torch.as_tensor(codes_MB).to(self.device)
        ivf_codes = torch.as_tensor(ivf_codes).to(self.device)
        if codes_MB.numel() and (int(codes_MB.min()) < 0 or int(codes_MB.max()) >= self.K_base):
            raise IndexError(f"codes out of range [0, {self.K_base})")
        if ivf_codes.numel() and (int(ivf_codes.min()) < 0 or int(ivf_codes.max()) >= self.ivf_K):
            raise IndexError(f"IVF codes out of range [0, {self.ivf_K})")
        return self.decode_u8(codes_MB.t().to(torch.uint8).contiguous(), ivf_codes.to(torch.int32))

    __call__ = forward

    @property
    def launch_count(self) -> int:
        return int(self._lib.qb_pairwise_launch_count(self._h))
