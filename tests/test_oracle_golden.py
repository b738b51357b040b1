"""Pin the CPU oracle to what the unmodified reference returned (tests/golden/*.npz, made by oracle/make_golden.py)."""
import numpy as np
import pytest

from conftest import GOLDEN_FULL, GOLDEN_IVF, GOLDEN_V2
from oracle import qinco_oracle as orc


@pytest.mark.parametrize("name", GOLDEN_V2 + GOLDEN_FULL)
def test_encode_matches_reference(name, golden_loader):
    cfg, w, z = golden_loader(name)
    x = z["x"]
    xn = (x - w["data_mean"]) / np.float32(w["data_std"])
    codes, xhat = orc.encode(cfg, w, xn)
    ref_codes, ref_xhat = z["codes_ref"], z["xhat_ref"]
    assert codes.shape == ref_codes.shape == (cfg["M"], len(x))
    # integer work: bit-exact
    np.testing.assert_array_equal(codes, ref_codes)
    assert codes.min() >= 0 and codes.max() < cfg["K"]
    # fp32 rounding only (different BLAS shapes / summation order)
    scale = np.abs(ref_xhat).max()
    assert np.abs(xhat - ref_xhat).max() <= 2e-5 * scale
    if name not in GOLDEN_FULL:          # (the full-depth models take a minute per numpy pass: once is enough)
        np.testing.assert_array_equal(orc.forward(cfg, w, x, "encode"), ref_codes)


@pytest.mark.parametrize("name", GOLDEN_V2 + GOLDEN_FULL)
def test_decode_matches_reference(name, golden_loader):
    cfg, w, z = golden_loader(name)
    dec = orc.forward(cfg, w, z["codes_ref"], "decode")
    ref = z["dec_ref"]
    rel = ((dec - ref) ** 2).sum() / (ref ** 2).sum()
    assert rel <= 1e-10
    # encode's x-hat is decode(codes) (invariant i)
    xhat = orc.decode(cfg, w, z["codes_ref"])
    assert np.abs(xhat - z["xhat_ref"]).max() <= 2e-5 * np.abs(z["xhat_ref"]).max()


def test_inference_wrapper_agreed_with_base_model(golden_loader):
    # recorded at generation time: QINCoInferenceWrapper == QINCo wherever the wrapper supports (A,B)
    for name in GOLDEN_V2 + GOLDEN_FULL:
        _, _, z = golden_loader(name)
        assert int(z["wrap_equal"]) in (1, -1)


def test_v1_codec_matches_reference(golden_loader):
    cfg, w, z = golden_loader("v1_codec")
    s = float(z["db_scale"])
    codes, mse = orc.codec_encode(cfg, w, z["x"], bs=40, db_scale=s)
    np.testing.assert_array_equal(codes, z["codes_ref"])
    assert codes.shape == (len(z["x"]), cfg["M"])
    assert abs(mse - float(z["mse_ref"])) <= 1e-4 * float(z["mse_ref"])
    dec = orc.codec_decode(cfg, w, z["codes_ref"], bs=40, db_scale=s)
    assert ((dec - z["dec_ref"]) ** 2).sum() / (z["dec_ref"] ** 2).sum() <= 1e-10


def test_distance_forms_agree():
    rng = np.random.default_rng(0)
    a = rng.standard_normal((40, 24), dtype=np.float32)
    b = rng.standard_normal((50, 24), dtype=np.float32)
    np.testing.assert_allclose(orc.approx_pairwise_distance(a, b), orc.exact_pairwise_distance(a, b), rtol=1e-4, atol=1e-4)
    bb = rng.standard_normal((40, 50, 24), dtype=np.float32)
    np.testing.assert_allclose(orc.compute_batch_distances(a[:, None], bb, approx=True),
                               orc.compute_batch_distances(a[:, None], bb, approx=False), rtol=1e-4, atol=1e-4)


def test_edge_cases():
    from qinco_b200 import synth
    cfg = synth.make_cfg(None, D=16, M=3, K=32, L=1, de=16, dh=16, A=4, B=3)
    w = synth.make_weights(cfg, seed=5, n_train=512, kmeans_iters=1)
    # empty input
    codes, xhat = orc.encode(cfg, w, np.zeros((0, 16), np.float32))
    assert codes.shape == (3, 0) and xhat.shape == (0, 16)
    assert orc.decode(cfg, w, codes).shape == (0, 16)
    # single row, and chunking invariance on a ragged batch
    x = synth.make_data(37, 16, seed=9)
    c1, h1 = orc.encode(cfg, w, x)
    c2, h2 = orc.encode(cfg, w, x, max_rows=5 * 3 * 4)
    np.testing.assert_array_equal(c1, c2)
    np.testing.assert_allclose(h1, h2, rtol=0, atol=1e-5)
    c3, _ = orc.encode(cfg, w, x[:1])
    np.testing.assert_array_equal(c3[:, 0], c1[:, 0])
    # M == 1: a plain nearest-codeword quantiser with beam 1 on exit
    cfg1 = synth.make_cfg(None, D=16, M=1, K=32, L=1, de=16, dh=16, A=0, B=4)
    w1 = synth.make_weights(cfg1, seed=6, n_train=512, kmeans_iters=1)
    c, h = orc.encode(cfg1, w1, x)
    d = orc.exact_pairwise_distance(x, w1["steps.0.codebook.weight"])
    np.testing.assert_array_equal(c[0], d.argmin(1))


@pytest.mark.parametrize("name", GOLDEN_V2 + GOLDEN_FULL + GOLDEN_IVF)
def test_torch_port_matches_reference(name, golden_loader):
    """oracle/torch_port.py (the timed CPU baseline) returns the reference's codes on every fixture."""
    from oracle.torch_port import TorchPort
    cfg, w, z = golden_loader(name)
    port = TorchPort(cfg, w)
    codes = port.forward(z["x"], "encode").numpy()
    np.testing.assert_array_equal(codes, z["codes_ref"])
    dec = port.forward(z["codes_ref"], "decode").numpy()
    assert ((dec - z["dec_ref"]) ** 2).sum() / (z["dec_ref"] ** 2).sum() <= 1e-10


@pytest.mark.parametrize("name", GOLDEN_IVF)
def test_ivf_model_matches_reference(name, golden_loader):
    """IVF-QINCo (IVFBook first step, qinco_base.py:128-196; step 1 pre-selects max(A, B) candidates, :108-112):
    the oracle returns the unmodified reference's codes (row 0 = IVF code) bit-exactly and its decode."""
    cfg, w, z = golden_loader(name)
    x = z["x"]
    xn = (x - w["data_mean"]) / np.float32(w["data_std"])
    codes, xhat = orc.encode(cfg, w, xn)
    assert codes.shape == z["codes_ref"].shape == (cfg["M"] + 1, len(x))
    np.testing.assert_array_equal(codes, z["codes_ref"])
    assert codes[0].max() < cfg["ivf_K"] and codes[1:].max() < cfg["K"]
    assert np.abs(xhat - z["xhat_ref"]).max() <= 2e-5 * np.abs(z["xhat_ref"]).max()
    dec = orc.forward(cfg, w, z["codes_ref"], "decode")
    assert ((dec - z["dec_ref"]) ** 2).sum() / (z["dec_ref"] ** 2).sum() <= 1e-10
