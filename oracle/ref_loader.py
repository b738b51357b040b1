"""Import the UNMODIFIED reference (read-only at /root/reference) in the dev container.

TEST INFRASTRUCTURE ONLY, and only usable where /root/reference exists (the dev
container).  Used by oracle/make_golden.py to produce tests/golden/*.npz and by the
optional tests in tests/test_oracle_vs_reference.py.  Nothing here runs on the GPU box.

The reference model imports `accelerate` at module level (qinco/utils.py:10,13) only
for QAccelerator, which the model never touches; a dummy module satisfies the import
(SURVEY.md section 8c).  No reference source is copied.
"""
from __future__ import annotations

import os
import sys
import types

REF_ROOT = os.environ.get("QINCO_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "qinco", "model"))


def _stub_accelerate():
    if "accelerate" in sys.modules:
        return
    acc = types.ModuleType("accelerate")
    acc.Accelerator = type("Accelerator", (object,), {})
    acc.data_loader = types.ModuleType("accelerate.data_loader")
    sys.modules["accelerate"] = acc
    sys.modules["accelerate.data_loader"] = acc.data_loader


class _FakeAccelerator:
    def __init__(self):
        import torch
        self.device = torch.device("cpu")

    def print(self, *a, **k):
        pass


def build_v2(cfg: dict, weights: dict, inference: bool = False):
    """Real reference QINCo (or QINCoInferenceWrapper) on CPU fp32 carrying `weights`."""
    import torch
    _stub_accelerate()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from qinco.model import QINCo, QINCoInferenceWrapper
    from qinco.utils import SharedCfgState

    c = SharedCfgState(dict(
        task="eval", M=cfg["M"], K=cfg["K"], L=cfg["L"], de=cfg["de"], dh=cfg["dh"], A=cfg["A"], B=cfg["B"],
        qinco1_mode=cfg["qinco1_mode"], enc_max_bs=65536, ivf_in_use=bool(cfg.get("ivf_K")), batch=64, codebook_noise_init=0.0,
        ivf_K=cfg.get("ivf_K"),
    ))
    c._accelerator = _FakeAccelerator()
    c._D = cfg["D"]
    c._M_ivf = cfg["M"] + (1 if cfg.get("ivf_K") else 0)        # qinco/qinco_tasks.py:378-383
    c._K_vals = ([cfg["ivf_K"]] if cfg.get("ivf_K") else []) + [cfg["K"]] * cfg["M"]
    c._ivf_book = None
    c._qinco_jit = False
    with torch.no_grad():
        if cfg.get("ivf_K"):
            from qinco.model.qinco_base import IVFBook
            c._ivf_book = IVFBook(c, weights["steps.0.ivf_centroids.weight"])
        model = QINCo(c)
        sd = {k: torch.from_numpy(v.copy()) for k, v in weights.items()}
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not unexpected, unexpected
        assert all(k.endswith(("xtarget_mean", "xtarget_var")) for k in missing), missing
        model.eval()
        if inference:
            model = QINCoInferenceWrapper(c, model)
            model.build()
    return model


def build_v1(cfg: dict, weights_v1: dict, db_scale: float = 1.0):
    """Real reference v1 QINCo (qinco_v1/model_qinco.py) on CPU fp32."""
    import torch
    v1 = os.path.join(REF_ROOT, "qinco_v1")
    if v1 not in sys.path:
        sys.path.insert(0, v1)
    import model_qinco  # noqa: the reference's own module name
    model = model_qinco.QINCo(cfg["D"], cfg["K"], cfg["L"], cfg["M"], cfg["dh"])
    model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in weights_v1.items()})
    model.db_scale = db_scale
    model.eval()
    return model


def v1_codec():
    v1 = os.path.join(REF_ROOT, "qinco_v1")
    if v1 not in sys.path:
        sys.path.insert(0, v1)
    import codec_qinco
    return codec_qinco


def pairwise_decoder_class():
    """The reference's PairwiseDecoderIVF class (qinco/search/pairwise_decoder.py).  Its module imports qinco.metrics,
    which subclasses torcheval.metrics.Metric at import time; torcheval is absent here and irrelevant to forward(), so a
    stub module with an empty `Metric` base class satisfies the import."""
    _stub_accelerate()
    if "torcheval" not in sys.modules:
        te, tm = types.ModuleType("torcheval"), types.ModuleType("torcheval.metrics")
        tm.Metric = type("Metric", (object,), {})
        te.metrics = tm
        sys.modules["torcheval"], sys.modules["torcheval.metrics"] = te, tm
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from qinco.search.pairwise_decoder import PairwiseDecoderIVF
    return PairwiseDecoderIVF


def build_pairwise(codebook_MKD, combine_mvals_m, ivf_code_map, K_base):
    """An instance of the reference class with just the tensors forward()/map_codes() read (no cfg, no training)."""
    import torch
    cls = pairwise_decoder_class()
    obj = cls.__new__(cls)
    torch.nn.Module.__init__(obj)
    obj.K_base = int(K_base)
    obj.codebook_MKD = torch.nn.Parameter(torch.as_tensor(codebook_MKD), requires_grad=False)
    obj.combine_mvals_m = torch.nn.Parameter(torch.as_tensor(combine_mvals_m), requires_grad=False)
    obj.ivf_code_map = torch.nn.Parameter(torch.as_tensor(ivf_code_map), requires_grad=False)
    return obj
