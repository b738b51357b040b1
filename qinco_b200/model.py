"""Host-side mirror of the reference's model surface for the encode/decode path.

`QINCo` stands in for both reference classes a caller may hold:
  * qinco.model.QINCo                    (reference qinco/model/qinco_base.py:419-549)
  * qinco.model.QINCoInferenceWrapper    (reference qinco/model/qinco_inference.py:257-353)
Same call surface — `model(x, step="encode") -> LongTensor [M, n]`, `model(codes, step="decode") -> [n, D]` in data
space, `.encode(x_norm) -> (codes [M, n], xhat [n, D])`, `.decode(codes [M, n])` in normalised space, `.data_mean`,
`.data_std`, `.load_state_dict`, `.build()`, `.built`, `.eval()`, `.to()` — but the work is done by libqinco_b200.so
(hand-written sm_100a kernels) through the C ABI; tensors cross as raw device pointers.  Training is out of scope.

Unlike the reference's GPU wrapper (which casts the whole model to fp16) the kernels keep fp32 residual streams,
tables and distances and use fp16 only for the tensor-core operands.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

_CFG_KEYS = ("M", "K", "L", "de", "dh", "A", "B", "qinco1_mode")


def _cfg_get(cfg, key, default=None):
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    try:
        v = getattr(cfg, key)
    except (AttributeError, KeyError):
        try:
            v = cfg[key]
        except Exception:
            return default
    return v


def normalize_cfg(cfg) -> dict:
    """Accept a plain dict ({D,M,K,L,de,dh,A,B,qinco1_mode}) or the reference's cfg object (cfg._D, cfg._M_ivf, ...)."""
    D = _cfg_get(cfg, "D") or _cfg_get(cfg, "_D")
    if D is None:
        raise ValueError("cfg needs D (or _D)")
    ivf_K = int(_cfg_get(cfg, "ivf_K", 0) or 0)
    if _cfg_get(cfg, "ivf_in_use", None) is False:      # the reference keeps ivf_K around when the IVF step is off
        ivf_K = 0
    M = _cfg_get(cfg, "M")                       # QINCo steps = uint8 codes per vector; cfg._M_ivf = M + 1 with an IVF step
    if M is None:
        M = int(_cfg_get(cfg, "_M_ivf")) - (1 if ivf_K else 0)
    out = dict(D=int(D), M=int(M), K=int(_cfg_get(cfg, "K", 256)), L=int(_cfg_get(cfg, "L")),
               de=int(_cfg_get(cfg, "de") or D), dh=int(_cfg_get(cfg, "dh")), A=int(_cfg_get(cfg, "A", 0) or 0),
               B=int(_cfg_get(cfg, "B", 1) or 1), qinco1_mode=bool(_cfg_get(cfg, "qinco1_mode", False)))
    if ivf_K:
        out["ivf_K"] = ivf_K                     # IVFBook first step (qinco_base.py:128-196)
    return out


def _to_numpy_state(sd) -> dict:
    out = {}
    for k, v in sd.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().float().cpu().numpy()
        out[k] = np.ascontiguousarray(np.asarray(v, dtype=np.float32))
    return out


class _Weight:
    """`.weight` holder with the look of an nn.Embedding / nn.Linear: the tensor is put on the model's device on first use."""

    def __init__(self, owner, key):
        self._owner, self._key, self._t = owner, key, None

    @property
    def weight(self):
        if self._t is None or self._t.device != self._owner.device:
            self._t = torch.from_numpy(self._owner._weights[self._key]).to(self._owner.device)
        return self._t


class _StepView:
    """What callers reach through `model.qinco_model.steps[m]`: the IVF search's re-ranking stage reads
    `steps[0].ivf_centroids.weight` (reference qinco/search/search_tasks.py:449; modules of qinco_base.py:128-146,
    :229-260).  Read-only views of the loaded state dict -- the kernels use their own packed copies."""

    _NAMES = {"codebook": "codebook.weight", "ivf_centroids": "ivf_centroids.weight", "in_proj": "in_proj.weight",
              "out_proj": "out_proj.weight"}

    def __init__(self, owner, m):
        for attr, key in self._NAMES.items():
            if f"steps.{m}.{key}" in owner._weights:
                setattr(self, attr, _Weight(owner, f"steps.{m}.{key}"))
        if f"steps.{m}.substep.codebook.weight" in owner._weights:
            self.substep = type("Substep", (), {})()
            self.substep.codebook = _Weight(owner, f"steps.{m}.substep.codebook.weight")


class QINCo:
    """B200 drop-in for the encode/decode surface of the reference model objects (see module docstring)."""

    def __init__(self, cfg, state_dict=None, device=None, plan_opts=None):
        self.cfg = normalize_cfg(cfg)
        self.D, self.M, self.K = self.cfg["D"], self.cfg["M"], self.cfg["K"]
        self.ivf_K = int(self.cfg.get("ivf_K") or 0)
        self.M_ivf = self.M + (1 if self.ivf_K else 0)      # rows of the reference's code matrix (cfg._M_ivf)
        acc = _cfg_get(cfg, "_accelerator", None)
        if device is None and acc is not None:
            device = getattr(acc, "device", None)
        self.device = torch.device(device if device is not None else "cuda:0")
        self.data_mean = torch.zeros(self.D)
        self.data_std = torch.zeros(())
        self.built = False
        self._plan_opts = plan_opts
        self._h = None
        self._weights = None
        self._ws = None
        self.qinco_model = self            # callers reach through the wrapper for `.qinco_model.steps[0]`
        self.steps = []                    # filled by load_state_dict: one _StepView per quantisation step
        if state_dict is not None:
            self.load_state_dict(state_dict)

    # ---- nn.Module-ish surface the callers use -----------------------------------------------------------------
    def eval(self):
        return self

    def train(self, mode=True):
        if mode:
            raise NotImplementedError("training is out of scope for the B200 encode/decode path")
        return self

    def to(self, device):
        device = torch.device(device)
        if device != self.device:
            self.device = device
            if self._weights is not None:
                self.build()
        return self

    def parameters(self):
        yield self.data_mean
        yield self.data_std

    def state_dict(self):
        return {k: torch.from_numpy(v.copy()) for k, v in (self._weights or {}).items()}

    def load_state_dict(self, state_dict, strict=True, **kwargs):
        """Keys are the reference's; training-only buffers (xtarget_*) are ignored."""
        w = _to_numpy_state(state_dict)
        w = {k: v for k, v in w.items() if not k.endswith(("xtarget_mean", "xtarget_var"))}
        missing = [k for k in self._expected_keys() if k not in w]
        if missing and strict:
            raise KeyError(f"missing keys in state dict: {missing[:6]}{'...' if len(missing) > 6 else ''}")
        c = self.cfg
        shapes = ({"steps.0.ivf_centroids.weight": (c["ivf_K"], c["D"])} if c.get("ivf_K") else
                  {"steps.0.codebook.weight": (c["K"], c["D"])})
        for k, shp in shapes.items():
            if tuple(w[k].shape) != shp:
                raise ValueError(f"{k}: expected shape {shp}, got {tuple(w[k].shape)}")
        w.setdefault("data_mean", np.zeros(self.D, np.float32))
        w.setdefault("data_std", np.array(1.0, np.float32))
        self._weights = w
        self.steps = [_StepView(self, m) for m in range(self.M_ivf)]
        self.build()

    def _expected_keys(self):
        c = self.cfg
        S = c["M"] + (1 if c.get("ivf_K") else 0)
        keys = [f"steps.{m}.codebook.weight" for m in range(1 if c.get("ivf_K") else 0, S)]
        if c.get("ivf_K"):
            keys.append("steps.0.ivf_centroids.weight")
        for m in range(1, S):
            keys += [f"steps.{m}.concat.mlp.weight", f"steps.{m}.concat.mlp.bias"]
            if c["A"] > 0:
                keys.append(f"steps.{m}.substep.codebook.weight")
            for l in range(c["L"]):
                keys += [f"steps.{m}.residual_blocks.{l}.up_proj.weight", f"steps.{m}.residual_blocks.{l}.down_proj.weight"]
            if c["de"] != c["D"]:
                keys += [f"steps.{m}.in_proj.weight", f"steps.{m}.out_proj.weight"]
        return keys

    def build(self):
        """Pack + upload the weights (the counterpart of QINCoInferenceWrapper.build, qinco_inference.py:290-330)."""
        if self._weights is None:
            raise RuntimeError("load_state_dict first")
        if self.device.type != "cuda":
            raise RuntimeError("qinco_b200 has no CPU path: the model must live on a CUDA (sm_100a) device")
        if not torch.cuda.is_available():
            raise RuntimeError("qinco_b200 needs a CUDA device (there is no CPU fallback)")
        if self._h is not None:
            self._h.close()
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        self._h = _lib.Handle(self.cfg, self._weights, device=idx, plan_opts=self._plan_opts)
        self.data_mean = torch.from_numpy(self._weights["data_mean"].copy()).to(self.device)
        self.data_std = torch.tensor(float(np.asarray(self._weights["data_std"]).reshape(-1)[0]), device=self.device)
        self._ws = None
        self.built = True
        return self

    # ---- raw layer: uint8 [n, M] codes, raw pointers -----------------------------------------------------------
    def _workspace(self, nbytes):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._ws

    def _check_x(self, x):
        if not isinstance(x, torch.Tensor):
            raise TypeError("x must be a torch.Tensor")
        if x.dim() != 2 or x.shape[1] != self.D:
            raise ValueError(f"x must be [n, {self.D}], got {tuple(x.shape)}")
        if x.device != self.device:
            raise ValueError(f"x is on {x.device}, the model on {self.device}")
        return x.float().contiguous()

    def encode_u8(self, x, normalize=False, want_xhat=True):
        """x [n, D] fp32 on the model's device -> (codes uint8 [n, M], xhat [n, D] or None); asynchronous."""
        if not self.built:
            raise RuntimeError("model not built")
        x = self._check_x(x)
        n = x.shape[0]
        codes = torch.empty((n, self.M), dtype=torch.uint8, device=self.device)
        xhat = torch.empty((n, self.D), dtype=torch.float32, device=self.device) if want_xhat else None
        if n:
            nbytes = self._h.encode_workspace_bytes(n)
            ws = self._workspace(nbytes)
            with torch.cuda.device(self.device):
                stream = torch.cuda.current_stream().cuda_stream
                self._h.encode(x.data_ptr(), n, normalize, codes.data_ptr(), xhat.data_ptr() if want_xhat else None,
                               ws.data_ptr(), ws.numel(), stream)
        return codes, xhat

    def decode_u8(self, codes_u8, denormalize=False):
        """codes uint8 [n, M] on the model's device -> [n, D] fp32; asynchronous."""
        if not self.built:
            raise RuntimeError("model not built")
        assert codes_u8.dtype == torch.uint8 and codes_u8.dim() == 2 and codes_u8.shape[1] == self.M
        codes_u8 = codes_u8.contiguous()
        n = codes_u8.shape[0]
        out = torch.empty((n, self.D), dtype=torch.float32, device=self.device)
        if n:
            nbytes = self._h.decode_workspace_bytes(n)
            ws = self._workspace(nbytes)
            with torch.cuda.device(self.device):
                stream = torch.cuda.current_stream().cuda_stream
                self._h.decode(codes_u8.data_ptr(), n, denormalize, out.data_ptr(), ws.data_ptr(), ws.numel(), stream)
        return out

    def encode_ivf_u8(self, x, normalize=False, want_xhat=True):
        """IVF models: x [n, D] -> (ivf codes int32 [n], codes uint8 [n, M], xhat [n, D] or None); asynchronous."""
        x = self._check_x(x)
        n = x.shape[0]
        ivf = torch.empty((n,), dtype=torch.int32, device=self.device)
        codes = torch.empty((n, self.M), dtype=torch.uint8, device=self.device)
        xhat = torch.empty((n, self.D), dtype=torch.float32, device=self.device) if want_xhat else None
        if n:
            ws = self._workspace(self._h.encode_workspace_bytes(n))
            with torch.cuda.device(self.device):
                stream = torch.cuda.current_stream().cuda_stream
                self._h.encode_ivf(x.data_ptr(), n, normalize, ivf.data_ptr(), codes.data_ptr(),
                                   xhat.data_ptr() if want_xhat else None, ws.data_ptr(), ws.numel(), stream)
        return ivf, codes, xhat

    def decode_ivf_u8(self, ivf_i32, codes_u8, denormalize=False):
        """IVF models: (ivf codes int32 [n], codes uint8 [n, M]) -> [n, D] fp32; asynchronous."""
        assert ivf_i32.dtype == torch.int32 and codes_u8.dtype == torch.uint8 and codes_u8.shape == (ivf_i32.shape[0], self.M)
        ivf_i32, codes_u8 = ivf_i32.contiguous(), codes_u8.contiguous()
        n = codes_u8.shape[0]
        out = torch.empty((n, self.D), dtype=torch.float32, device=self.device)
        if n:
            ws = self._workspace(self._h.decode_workspace_bytes(n))
            with torch.cuda.device(self.device):
                stream = torch.cuda.current_stream().cuda_stream
                self._h.decode_ivf(ivf_i32.data_ptr(), codes_u8.data_ptr(), n, denormalize, out.data_ptr(), ws.data_ptr(),
                                   ws.numel(), stream)
        return out

    def _pack_codes(self, codes_MB):
        """[M_ivf, n] integer codes (int64 / int32 / uint8, any strides; row 0 = IVF code for IVF models) ->
        (uint8 [n, M], int32 [n] or None) with ONE library kernel: no torch reductions, no host sync.  Out-of-range codes
        are flagged by that kernel and surface as IndexError from `synchronize()` (the reference's codebook lookup fails
        the same asynchronous way on a GPU)."""
        if not isinstance(codes_MB, torch.Tensor):
            codes_MB = torch.as_tensor(np.asarray(codes_MB))
        if codes_MB.dim() != 2 or codes_MB.shape[0] != self.M_ivf:
            raise AssertionError(f"codes must be [{self.M_ivf}, n], got {tuple(codes_MB.shape)}")          # qinco_base.py:449
        if codes_MB.dtype not in (torch.int64, torch.int32, torch.uint8):
            codes_MB = codes_MB.long()
        if codes_MB.device != self.device:
            codes_MB = codes_MB.to(self.device)
        n = codes_MB.shape[1]
        codes = torch.empty((n, self.M), dtype=torch.uint8, device=self.device)
        ivf = torch.empty((n,), dtype=torch.int32, device=self.device) if self.ivf_K else None
        if n:
            with torch.cuda.device(self.device):
                self._h.codes_pack(codes_MB.data_ptr(), codes_MB.element_size(), codes_MB.stride(0), codes_MB.stride(1), n,
                                   codes.data_ptr(), ivf.data_ptr() if ivf is not None else None,
                                   torch.cuda.current_stream().cuda_stream)
        return codes, ivf

    def _unpack_codes(self, codes_u8, ivf_i32=None):
        """(uint8 [n, M], int32 [n] or None) -> LongTensor [M_ivf, n], the reference's layout (qinco_base.py:480-485)."""
        n = codes_u8.shape[0]
        out = torch.empty((self.M_ivf, n), dtype=torch.int64, device=self.device)
        if n:
            with torch.cuda.device(self.device):
                self._h.codes_unpack(codes_u8.data_ptr(), ivf_i32.data_ptr() if ivf_i32 is not None else None, n,
                                     out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        return out

    # ---- the reference surface ------------------------------------------------------------------------------------
    @torch.no_grad()
    def encode(self, x_target_BD):
        """Normalised space: x [n, D] -> (codes LongTensor [M_ivf, n], xhat [n, D])   (qinco_base.py:454-485)."""
        if self.ivf_K:
            ivf, codes, xhat = self.encode_ivf_u8(x_target_BD, normalize=False, want_xhat=True)
            return self._unpack_codes(codes, ivf), xhat
        codes, xhat = self.encode_u8(x_target_BD, normalize=False, want_xhat=True)
        return self._unpack_codes(codes), xhat

    @torch.no_grad()
    def decode(self, codes_MB):
        """Normalised space: codes [M_ivf, n] (int64/int32/uint8) -> xhat [n, D] fp32   (qinco_base.py:447-452)."""
        codes, ivf = self._pack_codes(codes_MB)
        if self.ivf_K:
            return self.decode_ivf_u8(ivf, codes, denormalize=False)
        return self.decode_u8(codes, denormalize=False)

    @torch.no_grad()
    def forward(self, x_in, *args, step="train", **kwargs):
        assert step in ["train", "encode", "decode"]
        if step == "train":
            raise Exception("Don't use the B200 inference model for training!")
        assert float(self.data_std) > 0                                               # qinco_base.py:526
        if step == "encode":                                                          # :532-534
            if self.ivf_K:
                ivf, codes, _ = self.encode_ivf_u8(x_in, normalize=True, want_xhat=False)
                return self._unpack_codes(codes, ivf)
            codes, _ = self.encode_u8(x_in, normalize=True, want_xhat=False)
            return self._unpack_codes(codes)
        codes, ivf = self._pack_codes(x_in)                                           # :536-537
        if self.ivf_K:
            return self.decode_ivf_u8(ivf, codes, denormalize=True)
        return self.decode_u8(codes, denormalize=True)

    __call__ = forward

    def synchronize(self):
        """Wait for outstanding work and surface device-side failures: out-of-range codes raise IndexError (what the
        reference's codebook lookup raises), anything else the library's error."""
        torch.cuda.synchronize(self.device)
        try:
            self._h.check()
        except _lib.QbError as e:
            if "err word 0x10 " in str(e) or "err word 0x20 " in str(e):
                raise IndexError(f"codes out of range [0, {self.K})" + (f" or IVF codes out of range [0, {self.ivf_K})" if self.ivf_K else "")) from e
            raise

    @property
    def launch_count(self):
        return self._h.launch_count if self._h else 0
