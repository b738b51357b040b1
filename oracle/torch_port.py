"""CPU baseline: a PyTorch (CPU, fp32, multi-threaded) restatement of the reference's encode/decode, AS WRITTEN.

TEST / BENCH INFRASTRUCTURE ONLY (same rules as oracle/qinco_oracle.py: nothing under qinco_b200/ imports it).
It exists because the reference itself (/root/reference, Python) cannot travel to the GPU box, and the north-star
asks for "the reference's own PyTorch-CPU encode timed on the same box's host cores": this module issues the same
torch operator sequence per step as the reference — nn.Linear-shaped GEMMs over all [n*F*C, .] candidate rows without
any hoisting, `cat` for QConcat, the |a|^2+|b|^2-2ab distance via bmm, topk, gathers — so its timing stands in for
`QINCo.encode` / `QINCoInferenceWrapper.encode` on CPU:
  step MLP        qinco/model/qinco_base.py:262-280 (+ :60-64, :93-97);  qinco_inference.py:31-40
  beam step       qinco_base.py:292-374;  qinco_inference.py:89-140, 156-224
  distances       qinco/utils.py:336-346, 377-383
  encode / decode qinco_base.py:447-485;  qinco_inference.py:66-75, 239-254
It is pinned to the reference by tests/test_oracle_golden.py (identical codes on every committed fixture).

`device="cuda"` runs the SAME operator sequence through PyTorch on the GPU -- the "just run the reference on this box"
anchor of BASELINE.md section 3 -- in fp32, or with `dtype=torch.float16` in the numerics of the reference's own GPU
wrapper, which casts the whole model and the input to half (qinco_inference.py:303-317, 343).  bench.py reports it as
`torch_gpu_baseline`; it is never part of the product path.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


class TorchPort:
    def __init__(self, cfg: dict, weights: dict, threads: int | None = None, device="cpu", dtype=torch.float32):
        self.cfg = cfg
        if threads:
            torch.set_num_threads(threads)
        self.device, self.dtype = torch.device(device), dtype
        self.w = {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).to(self.device, dtype) for k, v in weights.items()}
        self.mean = self.w.get("data_mean", torch.zeros(cfg["D"], device=self.device, dtype=dtype))
        self.std = self.w.get("data_std", torch.tensor(1.0, device=self.device, dtype=dtype))

    # f_m(c, xhat): rows [..., D]
    def step_mlp(self, m, c, xhat):
        cfg, w, p = self.cfg, self.w, f"steps.{m}."
        proj = cfg["de"] != cfg["D"]
        e = F.linear(c, w[p + "in_proj.weight"]) if proj else c
        e = e + F.linear(torch.cat([e, xhat.expand(e.shape[:-1] + (cfg["D"],))], -1), w[p + "concat.mlp.weight"],
                         w[p + "concat.mlp.bias"])
        for l in range(cfg["L"]):
            e = e + F.linear(F.relu(F.linear(e, w[p + f"residual_blocks.{l}.up_proj.weight"])),
                             w[p + f"residual_blocks.{l}.down_proj.weight"])
        o = F.linear(e, w[p + "out_proj.weight"]) if proj else e
        return o if cfg["qinco1_mode"] else o + c

    @staticmethod
    def pairwise(a, b):              # [n, D] x [K, D] -> [n, K]
        return (a * a).sum(-1, keepdim=True) + (b * b).sum(-1)[None, :] - 2 * a @ b.t()

    @staticmethod
    def batch_dist(x, cand):         # [n, D], [n, R, D] -> [n, R]
        return (x * x).sum(-1, keepdim=True) + (cand * cand).sum(-1) - 2 * torch.bmm(cand, x.unsqueeze(-1)).squeeze(-1)

    @torch.no_grad()
    def decode(self, codes_MB):
        codes_MB = torch.as_tensor(codes_MB).long().to(self.device)
        ivf = bool(self.cfg.get("ivf_K"))                # IVFBook.decode (qinco_base.py:176-183): centroid lookup
        xhat = self.w["steps.0.ivf_centroids.weight" if ivf else "steps.0.codebook.weight"][codes_MB[0]].clone()
        for m in range(1, self.cfg["M"] + (1 if ivf else 0)):
            xhat = xhat + self.step_mlp(m, self.w[f"steps.{m}.codebook.weight"][codes_MB[m]], xhat)
        return xhat

    @torch.no_grad()
    def encode(self, x, max_rows=65536):
        cfg = self.cfg
        K, D, M, A, B = cfg["K"], cfg["D"], cfg["M"], cfg["A"], cfg["B"]
        x = torch.as_tensor(x).to(self.device, self.dtype)
        bs = max(1, max_rows // (B * (A or 1)))          # qinco_base.py:456-472 (enc_max_bs // (B * max(A,1)))
        if len(x) > bs:
            parts = [self.encode(x[i:i + bs], max_rows) for i in range(0, len(x), bs)]
            return torch.cat([p[0] for p in parts], 1), torch.cat([p[1] for p in parts])
        n = len(x)
        ivf = bool(cfg.get("ivf_K"))
        if ivf:                                          # IVFBook.quantize (qinco_base.py:146-163): arg-min, one beam
            M = M + 1                                    # cfg._M_ivf
            cent = self.w["steps.0.ivf_centroids.weight"]
            rows = max(1, (1 << 30) // len(cent))        # IVFBook.quantize batches rows the same way (qinco_base.py:149-156)
            c0 = torch.cat([self.pairwise(x[i:i + rows], cent).argmin(-1, keepdim=True) for i in range(0, max(n, 1), rows)])
            xhat = self.w["steps.0.ivf_centroids.weight"][c0]
        else:
            F1 = B if M > 1 else 1
            d0 = self.pairwise(x, self.w["steps.0.codebook.weight"])
            c0 = d0.topk(F1, dim=-1, largest=False).indices if F1 > 1 else d0.argmin(-1, keepdim=True)
            xhat = self.w["steps.0.codebook.weight"][c0]     # [n, F, D]
        hist = c0.unsqueeze(0)                           # [m, n, F]
        A_cfg = A
        for m in range(1, M):
            F_in, F_out = xhat.shape[1], (B if m < M - 1 else 1)
            cb = self.w[f"steps.{m}.codebook.weight"]
            A = max(A_cfg, B) if (A_cfg > 0 and ivf and m == 1) else A_cfg      # QincoSubstep._n_codes (qinco_base.py:108-112)
            if A > 0:
                r = (x.unsqueeze(1) - xhat).reshape(n * F_in, D)
                idx = self.pairwise(r, self.w[f"steps.{m}.substep.codebook.weight"]).topk(A, -1, largest=False).indices
                c = cb[idx].reshape(n, F_in, A, D)
                C = A
            else:
                idx = None
                c = cb.reshape(1, 1, K, D).expand(n, F_in, K, D)
                C = K
            xh = xhat.unsqueeze(2)
            cand = (self.step_mlp(m, c, xh) + xh).reshape(n, F_in * C, D)
            dist = self.batch_dist(x, cand)
            sel = dist.topk(F_out, -1, largest=False).indices if F_out > 1 else dist.argmin(-1, keepdim=True)
            code = idx.reshape(n, F_in * C).gather(-1, sel) if idx is not None else sel % C
            parent = sel // C
            hist = torch.cat([hist.gather(-1, parent.unsqueeze(0).expand(len(hist), n, F_out)), code.unsqueeze(0)])
            xhat = cand.gather(1, sel.unsqueeze(-1).expand(n, F_out, D))
        return hist.squeeze(-1), xhat.squeeze(1)

    def forward(self, x_in, step):
        if step == "encode":
            return self.encode((torch.as_tensor(x_in).to(self.device, self.dtype) - self.mean) / self.std)[0]
        return self.decode(x_in) * self.std + self.mean
