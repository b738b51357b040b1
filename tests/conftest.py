import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) on a machine without CUDA; the CPU tests that replay the planner need the
    built library and are skipped when it is neither present nor buildable."""
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if not has_cuda:
        skip = pytest.mark.skip(reason="needs a CUDA device (sm_100a); run with -m gpu on the B200 box")
        for it in items:
            if "gpu" in it.keywords:
                it.add_marker(skip)


def load_golden(name):
    """(cfg, weights, fixture) — weights come from the fixture or are regenerated from its seed and digest-checked."""
    from qinco_b200 import synth
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg = json.loads(str(z["cfg"]))
    w = {k[2:]: z[k] for k in z.files if k.startswith("w:")}
    if not w:
        kw = dict(seed=int(z["wseed"]), gain=float(z["gain"]), n_train=int(z["n_train"]),
                  kmeans_iters=int(z["kmeans_iters"]))
        if "data_mean" in z.files:
            kw.update(data_mean=float(z["data_mean"]), data_std=float(z["data_std"]))
        w = synth.make_weights(cfg, **kw)
    assert synth.weights_digest(w) == str(z["digest"]), f"{name}: regenerated weights drifted from the fixture"
    return cfg, w, z


GOLDEN_V2 = ["tiny_q1", "tiny_a0_b1", "tiny_a8_b4", "tiny_a0_b4", "s_a0_b1", "s_a16_b8", "s_a0_b4",
             "proj_a8_b4", "q1_l4", "l_a16_b16"]

# BASELINE configurations at full depth (L = 16, every step), 64 rows each, produced by the unmodified reference
GOLDEN_FULL = ["full_q1", "full_l_b16", "full_deep_m16", "full_contr_b32"]


# IVF-QINCo fixtures: code matrices have M + 1 rows, row 0 = the IVF code
GOLDEN_IVF = ["ivf_a8_b4", "ivf_a4_b8", "ivf_a0_b1"]


@pytest.fixture(scope="session")
def golden_loader():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_golden(name)
        return cache[name]
    return get
