"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the committed reference outputs.

Tolerances (BASELINE.json north_star): integer codes round-trip bit-exact; reconstruction within 1e-4 relative MSE of
the reference PyTorch fp32 path on the same inputs.  Encode agreement is judged on MSE (fp16 tensor-core operands can
flip near-tied candidates), plus per-decision optimality for greedy (B=1) models.
"""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_FULL, GOLDEN_V2
from oracle import qinco_oracle as orc
from qinco_b200 import synth

pytestmark = pytest.mark.gpu

DEC_TOL = 1e-4        # relative MSE of decode vs reference (north_star)
STEP_TOL = 1e-5       # one MLP step, relative MSE
ENC_TOL_SMALL = 5e-3  # |MSE_ours - MSE_ref| / MSE_ref on the tiny golden samples (tens of vectors: one flipped path moves it)
ENC_TOL_LARGE = 1e-4  # ... on thousands of vectors (north_star)
# encode's returned x-hat vs decode(codes): two kernels that both evaluate u = Wx . xhat to fp32 accuracy (fp16 hi/lo
# products on the tensor core) but in a different summation order; a 1e-7 difference in e0 occasionally flips the fp16
# rounding of an activation, which a 16-block MLP amplifies.  Both stay within the north-star's 1e-4 of the reference; the
# two are held to 1e-6 of each other (the reference's own two paths agree to fp32 rounding, oracle/make_golden.py enc_dec_gap).
ENC_DEC_TOL = 1e-6


@pytest.fixture(scope="module")
def models(golden_loader):
    from qinco_b200.model import QINCo
    cache = {}

    def get(name):
        if name not in cache:
            cfg, w, z = golden_loader(name)
            cache[name] = (QINCo(cfg, w, device="cuda:0"), cfg, w, z)
        return cache[name]
    yield get
    for m, *_ in cache.values():
        m._h.close()


@pytest.mark.parametrize("name", ["s_a0_b1", "s_a16_b8", "proj_a8_b4", "l_a16_b16"])
def test_cta_pair_kernel_matches_reference(name, golden_loader):
    """The cta_group::2 variant of the MLP kernel (plan option pair = 2: one M = 256 MMA over two CTAs, half of every
    weight slab per CTA) against the same reference outputs as the default kernel."""
    from qinco_b200.model import QINCo
    cfg, w, z = golden_loader(name)
    model = QINCo(cfg, w, device="cuda:0", plan_opts={"n_tiles": 2 << 8})
    try:
        assert model._h.info(1)["pair"] == 1
        codes_ref = torch.from_numpy(z["codes_ref"]).cuda()
        dec = model(codes_ref, step="decode")
        model.synchronize()
        assert rel_mse(dec.cpu().numpy(), z["dec_ref"]) <= DEC_TOL
        x = z["x"]
        xn = (x - w["data_mean"]) / np.float32(w["data_std"])
        c = model(torch.from_numpy(x).cuda(), step="encode").cpu().numpy()
        model.synchronize()
        agree = float((c == z["codes_ref"]).all(0).mean())
        mse_ours, mse_ref = orc.mse(xn, orc.decode(cfg, w, c)), orc.mse(xn, z["xhat_ref"])
        assert agree >= 0.8 and abs(mse_ours - mse_ref) <= ENC_TOL_SMALL * mse_ref, (agree, mse_ours, mse_ref)
    finally:
        model._h.close()


def rel_mse(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(((a - b) ** 2).sum() / max((b ** 2).sum(), 1e-30))


@pytest.mark.parametrize("name", GOLDEN_V2)
def test_one_step_matches_oracle(name, models):
    """qb_debug_step == xhat + QINCoStep.forward(C_m[code], xhat)   (qinco_base.py:262-290), ragged row count."""
    model, cfg, w, z = models(name)
    rng = np.random.default_rng(7)
    n = 333
    for step in sorted({1, cfg["M"] - 1}):
        xhat = rng.standard_normal((n, cfg["D"]), dtype=np.float32)
        codes = rng.integers(0, cfg["K"], size=n).astype(np.uint8)
        ref = xhat + orc.step_mlp(cfg, w, step, w[f"steps.{step}.codebook.weight"][codes], xhat)
        xt = torch.from_numpy(xhat).cuda()
        ct = torch.from_numpy(codes).cuda()
        out = torch.empty_like(xt)
        ws = torch.empty(n * cfg["de"] * 4 + 512, dtype=torch.uint8, device="cuda")
        model._h.debug_step(step, xt.data_ptr(), ct.data_ptr(), n, out.data_ptr(), ws.data_ptr(), ws.numel(),
                            torch.cuda.current_stream().cuda_stream)
        model.synchronize()
        got = out.cpu().numpy()
        err = rel_mse(got, ref)
        if err > STEP_TOL:   # diagnostics that localise an operand-layout bug
            d = np.abs(got - ref)
            print(f"\n{name} step {step}: rel_mse={err:.3e} max|d|={d.max():.3e}")
            print("  per-32-row block max:", np.round([d[i:i + 32].max() for i in range(0, n, 32)], 4))
            print("  per-8-col block max:", np.round([d[:, j:j + 8].max() for j in range(0, cfg['D'], 8)], 4))
        assert err <= STEP_TOL, f"{name} step {step}: rel mse {err:.3e}"


@pytest.mark.parametrize("name", GOLDEN_V2)
def test_decode_matches_reference(name, models):
    model, cfg, w, z = models(name)
    codes = torch.from_numpy(z["codes_ref"]).cuda()
    dec = model(codes, step="decode")
    model.synchronize()
    assert dec.dtype == torch.float32 and tuple(dec.shape) == (codes.shape[1], cfg["D"])
    err = rel_mse(dec.cpu().numpy(), z["dec_ref"])
    assert err <= DEC_TOL, f"{name}: decode rel mse {err:.3e}"
    # normalised-space decode and int32 codes (search_tasks.py:428-445 passes int32)
    xh = model.decode(codes.int())
    assert rel_mse(xh.cpu().numpy(), z["xhat_ref"]) <= DEC_TOL


def _greedy_decisions_near_optimal(cfg, w, xn, codes_MB, tol=2e-3):
    """For B=1 models: every chosen code's fp32 distance is within tol of the step's minimum (oracle arithmetic)."""
    n = len(xn)
    xhat = np.zeros((n, cfg["D"]), np.float32)
    worst = 0.0
    for m in range(cfg["M"]):
        cb = w[f"steps.{m}.codebook.weight"]
        if m == 0:
            cand = np.broadcast_to(cb[None], (n,) + cb.shape)
        else:
            xh = xhat[:, None, :]
            cand = orc.step_mlp(cfg, w, m, np.broadcast_to(cb[None], (n,) + cb.shape), xh) + xh
        d = ((xn[:, None, :] - cand) ** 2).sum(-1)
        chosen = d[np.arange(n), codes_MB[m]]
        worst = max(worst, float(((chosen - d.min(1)) / d.min(1)).max()))
        xhat = cand[np.arange(n), codes_MB[m]].astype(np.float32)
    return worst <= tol, worst


@pytest.mark.parametrize("name", GOLDEN_V2)
def test_encode_matches_reference(name, models):
    model, cfg, w, z = models(name)
    x = z["x"]
    xn = (x - w["data_mean"]) / np.float32(w["data_std"])
    codes = model(torch.from_numpy(x).cuda(), step="encode")
    model.synchronize()
    assert codes.dtype == torch.int64 and tuple(codes.shape) == (cfg["M"], len(x))
    c = codes.cpu().numpy()
    assert c.min() >= 0 and c.max() < cfg["K"]
    agree = float((c == z["codes_ref"]).all(0).mean())
    # MSE of OUR codes decoded by the ORACLE (reference arithmetic) vs the reference's own MSE
    mse_ours = orc.mse(xn, orc.decode(cfg, w, c))
    mse_ref = orc.mse(xn, z["xhat_ref"])
    rel = abs(mse_ours - mse_ref) / mse_ref
    print(f"\n{name}: vectors with identical codes {agree:.3f}, mse ours {mse_ours:.6f} ref {mse_ref:.6f} rel {rel:.2e}")
    assert rel <= ENC_TOL_SMALL
    assert agree >= 0.8
    if cfg["A"] == 0 and cfg["B"] == 1:
        ok, worst = _greedy_decisions_near_optimal(cfg, w, xn, c)
        assert ok, f"a greedy decision is {worst:.2e} worse than the optimum"
    # encode()'s xhat is decode(codes) (SURVEY section 4 invariant i) and codes round-trip through uint8 bit-exactly
    codes2, xhat = model.encode(torch.from_numpy(xn).cuda())
    assert torch.equal(codes2, codes)
    dec = model.decode(codes2)
    model.synchronize()
    assert rel_mse(xhat.cpu().numpy(), dec.cpu().numpy()) <= ENC_DEC_TOL
    u8, _ = model.encode_u8(torch.from_numpy(xn).cuda())
    assert torch.equal(u8.t().long(), codes)


def test_encode_large_sample_mse():
    """QINCo2-S shape (BASELINE config 2), 4096 vectors: encode MSE within 1e-4 of the fp32 oracle's."""
    from qinco_b200.model import QINCo
    cfg = synth.make_cfg(None, D=128, M=8, K=256, L=2, de=128, dh=256, A=0, B=1)
    w = synth.make_weights(cfg, seed=4321, n_train=4096, kmeans_iters=2)
    x = synth.make_data(4096, 128, seed=1234)
    model = QINCo(cfg, w)
    codes = model(torch.from_numpy(x).cuda(), step="encode").cpu().numpy()
    model.synchronize()
    ref_codes, ref_xhat = orc.encode(cfg, w, x)
    mse_ours = orc.mse(x, orc.decode(cfg, w, codes))
    mse_ref = orc.mse(x, ref_xhat)
    agree = float((codes == ref_codes).all(0).mean())
    rel = abs(mse_ours - mse_ref) / mse_ref
    print(f"\nS 4096: identical-code vectors {agree:.4f}, mse ours {mse_ours:.6f} ref {mse_ref:.6f} rel {rel:.2e}")
    assert rel <= ENC_TOL_LARGE
    assert agree >= 0.97
    ok, worst = _greedy_decisions_near_optimal(cfg, w, x[:512], codes[:, :512])
    assert ok, worst
    # size-independent properties on a bigger batch: determinism, chunking invariance, decode(encode) consistency
    xb = torch.from_numpy(synth.make_data(40000, 128, seed=99)).cuda()
    c1, xh1 = model.encode_u8(xb)
    c2, _ = model.encode_u8(xb)
    c3 = torch.cat([model.encode_u8(xb[i:i + 7777])[0] for i in range(0, len(xb), 7777)])
    model.synchronize()
    assert torch.equal(c1, c2) and torch.equal(c1, c3)
    assert rel_mse(model.decode_u8(c1).cpu().numpy(), xh1.cpu().numpy()) <= ENC_DEC_TOL
    model._h.close()


def test_edge_cases(models):
    model, cfg, w, z = models("tiny_a8_b4")
    D, M = cfg["D"], cfg["M"]
    # empty input
    codes = model(torch.zeros((0, D), device="cuda"), step="encode")
    assert tuple(codes.shape) == (M, 0) and codes.dtype == torch.int64
    assert tuple(model(codes, step="decode").shape) == (0, D)
    # single row and ragged sizes agree with the batched result
    x = torch.from_numpy(z["x"]).cuda()
    full = model(x, step="encode")
    for n in (1, 2, 37):
        assert torch.equal(model(x[:n], step="encode"), full[:, :n])
    # out-of-range codes are rejected
    bad = full.clone()
    bad[0, 0] = cfg["K"]
    with pytest.raises(IndexError):      # flagged by the pack kernel, surfaced (once) at the next synchronisation point
        model(bad, step="decode")
        model.synchronize()
    model.synchronize()                  # the report cleared the word: the model stays usable
    assert torch.equal(model(x, step="encode"), full)
    neg = full.clone()
    neg[1, 3] = -1
    with pytest.raises(IndexError):
        model.decode(neg.int())
        model.synchronize()
    # strided int32 views, as the IVF search passes them (search_tasks.py:428-445: codes_int32[i:j].T)
    vm = full.t().contiguous().int()
    assert torch.equal(model.decode(vm.T), model.decode(full))
    with pytest.raises(AssertionError):
        model.decode(full[:-1])
    # workspace too small -> error status, no launch
    from qinco_b200._lib import QbError
    xs = x[:8].contiguous()
    out = torch.empty((8, M), dtype=torch.uint8, device="cuda")
    ws = torch.empty(4096, dtype=torch.uint8, device="cuda")
    with pytest.raises(QbError) as e:
        model._h.encode(xs.data_ptr(), 8, False, out.data_ptr(), None, ws.data_ptr(), ws.numel(), 0)
    assert e.value.code == -3
    model.synchronize()


def test_single_step_model():
    """M == 1: a plain nearest-codeword quantiser (qinco_base.py:218,263), beam 1 on exit."""
    from qinco_b200.model import QINCo
    cfg = synth.make_cfg(None, D=16, M=1, K=32, L=1, de=16, dh=16, A=0, B=4)
    w = synth.make_weights(cfg, seed=6, n_train=512, kmeans_iters=1, data_mean=0.3, data_std=1.7)
    x = synth.make_data(50, 16, seed=9, mean=0.3, std=1.7)
    model = QINCo(cfg, w)
    codes = model(torch.from_numpy(x).cuda(), step="encode")
    ref = orc.forward(cfg, w, x, "encode")
    np.testing.assert_array_equal(codes.cpu().numpy(), ref)
    dec = model(codes, step="decode").cpu().numpy()
    assert rel_mse(dec, orc.forward(cfg, w, ref, "decode")) <= 1e-12
    model._h.close()


def test_v1_codec_matches_reference(golden_loader):
    from qinco_b200 import codec
    cfg, w, z = golden_loader("v1_codec")
    model = codec.QINCoV1(cfg=cfg, weights=w, db_scale=float(z["db_scale"]))
    codes = codec.encode(model, z["x"], bs=40, is_float16=False)
    assert codes.dtype == np.int64 and codes.shape == z["codes_ref"].shape
    agree = float((codes == z["codes_ref"]).all(1).mean())
    assert agree >= 0.9, agree
    dec = codec.decode(model, z["codes_ref"], bs=40, is_float16=False)
    assert dec.dtype == np.float32
    assert rel_mse(dec, z["dec_ref"]) <= DEC_TOL
    # the v1 model's own tensor surface (model_qinco.py:91-118)
    xt = torch.from_numpy(z["x"] / np.float32(z["db_scale"])).cuda()
    c, xh = model.encode(xt)
    assert tuple(c.shape) == (len(xt), cfg["M"]) and c.dtype == torch.int64
    np.testing.assert_array_equal(c.cpu().numpy(), codes)
    assert rel_mse(model.decode(c).cpu().numpy(), xh.cpu().numpy()) <= ENC_DEC_TOL
    # from a v1-keyed state dict
    m2 = codec.QINCoV1(synth.to_v1_state(cfg, w), db_scale=float(z["db_scale"]))
    np.testing.assert_array_equal(codec.encode(m2, z["x"], bs=96, verbose=False), codes)


# ---------------------------------------------------------------------------------------------------------------- IVF
from conftest import GOLDEN_IVF  # noqa: E402


@pytest.mark.parametrize("name", GOLDEN_IVF)
def test_ivf_model_matches_reference(name, golden_loader):
    """IVF-QINCo: IVF codes (row 0) against the reference's arg-min, decode of the reference's codes,
    encode MSE, surface shapes ([M + 1, n] int64) and errors."""
    from qinco_b200.model import QINCo
    cfg, w, z = golden_loader(name)
    model = QINCo(cfg, w, device="cuda:0")
    try:
        x = z["x"]
        xn = (x - w["data_mean"]) / np.float32(w["data_std"])
        dec = model(torch.from_numpy(z["codes_ref"]).cuda(), step="decode")
        model.synchronize()
        assert rel_mse(dec.cpu().numpy(), z["dec_ref"]) <= DEC_TOL
        codes = model(torch.from_numpy(x).cuda(), step="encode")
        model.synchronize()
        assert codes.dtype == torch.int64 and tuple(codes.shape) == (cfg["M"] + 1, len(x))
        c = codes.cpu().numpy()
        assert c[0].max() < cfg["ivf_K"] and c[1:].max() < cfg["K"] and c.min() >= 0
        # the IVF code is an fp32 arg-min over distinct centroids: it must agree with the reference (ties aside)
        assert (c[0] == z["codes_ref"][0]).mean() >= 0.98
        agree = float((c == z["codes_ref"]).all(0).mean())
        mse_ours, mse_ref = orc.mse(xn, orc.decode(cfg, w, c)), orc.mse(xn, z["xhat_ref"])
        print(f"\n{name}: identical vectors {agree:.3f} mse ours {mse_ours:.5f} ref {mse_ref:.5f}")
        assert agree >= 0.8 and abs(mse_ours - mse_ref) <= ENC_TOL_SMALL * mse_ref
        codes2, xhat = model.encode(torch.from_numpy(xn).cuda())
        assert torch.equal(codes2, codes)
        assert rel_mse(xhat.cpu().numpy(), model.decode(codes2).cpu().numpy()) <= ENC_DEC_TOL
        with pytest.raises(IndexError):
            bad = z["codes_ref"].copy()
            bad[0, 0] = cfg["ivf_K"]
            model.decode(torch.from_numpy(bad))
            model.synchronize()
        model.synchronize()
        # the model surface the IVF search reads (search_tasks.py:449)
        cent = model.qinco_model.steps[0].ivf_centroids.weight
        assert cent.is_cuda and tuple(cent.shape) == (cfg["ivf_K"], cfg["D"])
        np.testing.assert_array_equal(cent.cpu().numpy(), w["steps.0.ivf_centroids.weight"])
        with pytest.raises(Exception):       # the plain entry points refuse IVF models
            model.encode_u8(torch.from_numpy(xn).cuda())
        assert model(torch.zeros((0, cfg["D"])).cuda(), step="encode").shape == (cfg["M"] + 1, 0)
    finally:
        model._h.close()


def test_ivf_assign_large_matches_oracle():
    """The IVF arg-min kernel alone at a ragged size: 5000 vectors x 3001 centroids, d = 96, with normalisation."""
    from qinco_b200.model import QINCo
    cfg = synth.make_cfg(None, D=96, M=2, K=32, L=1, de=96, dh=32, A=0, B=1, ivf_K=3001)
    w = synth.make_weights(cfg, seed=3, n_train=6000, kmeans_iters=1, data_mean=0.3, data_std=1.7)
    x = synth.make_data(5000, 96, seed=8, mean=0.3, std=1.7)
    xn = (x - w["data_mean"]) / np.float32(w["data_std"])
    ref = orc.approx_pairwise_distance(xn, w["steps.0.ivf_centroids.weight"]).argmin(-1)
    model = QINCo(cfg, w, device="cuda:0")
    try:
        codes = model(torch.from_numpy(x).cuda(), step="encode").cpu().numpy()
        model.synchronize()
        same = (codes[0] == ref)
        assert same.mean() >= 0.999, same.mean()
        # where they differ the two centroids are tied to fp32 rounding
        d = orc.exact_pairwise_distance(xn[~same], w["steps.0.ivf_centroids.weight"])
        rows = np.arange((~same).sum())
        assert np.all(np.abs(d[rows, codes[0][~same]] - d[rows, ref[~same]]) <= 1e-4 * d[rows, ref[~same]])
    finally:
        model._h.close()


@pytest.mark.parametrize("shape", [
    dict(D=768, M=3, K=256, L=2, de=384, dh=384, A=16, B=32, qinco1_mode=False),      # Contriever shape (BASELINE config 5)
    dict(D=96, M=6, K=256, L=2, de=384, dh=384, A=16, B=16, qinco1_mode=False),       # Deep1B shape (BASELINE config 4)
])
def test_baseline_config_shapes_vs_oracle(shape):
    """The wide-beam shapes of BASELINE configs 4 / 5 (fewer steps and blocks so the numpy oracle stays fast):
    out_proj chunks over D = 768, beam 32, D = 96 with de = 384."""
    from qinco_b200.model import QINCo
    cfg = synth.make_cfg(None, **shape)
    w = synth.make_weights(cfg, seed=17, n_train=2048, kmeans_iters=1)
    x = synth.make_data(96, cfg["D"], seed=23)
    ref_codes, ref_xhat = orc.encode(cfg, w, x)
    model = QINCo(cfg, w, device="cuda:0")
    try:
        codes = model(torch.from_numpy(x).cuda(), step="encode").cpu().numpy()
        dec = model(torch.from_numpy(ref_codes).cuda(), step="decode").cpu().numpy()
        model.synchronize()
        assert rel_mse(dec, orc.decode(cfg, w, ref_codes)) <= DEC_TOL
        agree = float((codes == ref_codes).all(0).mean())
        mse_ours, mse_ref = orc.mse(x, orc.decode(cfg, w, codes)), orc.mse(x, ref_xhat)
        print(f"\n{shape['D']}: identical vectors {agree:.3f} mse ours {mse_ours:.5f} ref {mse_ref:.5f}")
        assert agree >= 0.8 and abs(mse_ours - mse_ref) <= ENC_TOL_SMALL * mse_ref
    finally:
        model._h.close()


def test_one_step_random_shapes_vs_oracle():
    """qb_debug_step (prep + the tcgen05 MLP kernel in apply mode) on a dozen random legal shapes: odd multiples of 16,
    with / without projections and outer skip, ragged row counts."""
    from qinco_b200.model import QINCo
    rng = np.random.default_rng(77)
    tried = 0
    while tried < 12:
        D = int(rng.choice([16, 32, 48, 64, 96, 128, 192]))
        de = int(rng.choice([D, 16 * int(rng.integers(1, 25))]))
        dh = 16 * int(rng.integers(1, 25))
        L = int(rng.integers(0, 4))
        K = int(rng.choice([16, 64, 100, 256]))
        cfg = synth.make_cfg(None, D=D, M=2, K=K, L=L, de=de, dh=dh, A=0, B=1, qinco1_mode=bool(rng.integers(0, 2)))
        w = synth.make_weights(cfg, seed=int(rng.integers(1, 1000)), n_train=max(256, K), kmeans_iters=1)
        try:
            model = QINCo(cfg, w, device="cuda:0")
        except Exception as e:     # shapes the planner refuses must say why
            assert "TMEM" in str(e) or "shared memory" in str(e), (cfg, str(e))
            continue
        tried += 1
        try:
            n = int(rng.integers(1, 700))
            xhat = rng.standard_normal((n, D), dtype=np.float32)
            codes = rng.integers(0, K, size=n).astype(np.uint8)
            ref = xhat + orc.step_mlp(cfg, w, 1, w["steps.1.codebook.weight"][codes], xhat)
            xt, ct = torch.from_numpy(xhat).cuda(), torch.from_numpy(codes).cuda()
            out = torch.empty_like(xt)
            ws = torch.empty(n * de * 4 + 512, dtype=torch.uint8, device="cuda")
            model._h.debug_step(1, xt.data_ptr(), ct.data_ptr(), n, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                torch.cuda.current_stream().cuda_stream)
            model.synchronize()
            err = rel_mse(out.cpu().numpy(), ref)
            assert err <= STEP_TOL, (cfg, n, err)
        finally:
            model._h.close()


def test_encode_random_configs_vs_oracle():
    """Whole encode / decode on random small configurations (odd K, D, de, dh; A in {0, 3, 8}; beams 1..6; 2..5 steps;
    sometimes an IVF first step): codes mostly identical to the fp32 oracle, MSE within tolerance, decode within 1e-4."""
    from qinco_b200.model import QINCo
    rng = np.random.default_rng(4242)
    for it in range(10):
        D = int(rng.choice([16, 32, 48, 96]))
        de = int(rng.choice([D, 16 * int(rng.integers(1, 13))]))
        K = int(rng.choice([16, 50, 64, 256]))
        A = int(rng.choice([0, 3, 8]))
        B = int(rng.integers(1, 7))
        A = min(A, K)
        if A and A * 1 < B:            # a beam step needs at least B candidates: keep A*F_in >= B for every step
            A = B
        kw = dict(D=D, M=int(rng.integers(2, 6)), K=K, L=int(rng.integers(0, 3)), de=de, dh=16 * int(rng.integers(1, 9)),
                  A=A, B=min(B, K), qinco1_mode=bool(rng.integers(0, 2)))
        if it % 3 == 2:
            kw["ivf_K"] = int(rng.integers(5, 300))
        cfg = synth.make_cfg(None, **kw)
        w = synth.make_weights(cfg, seed=100 + it, n_train=1024, kmeans_iters=1, data_mean=0.1 * it, data_std=1.0 + 0.1 * it)
        n = int(rng.integers(1, 130))
        x = synth.make_data(n, D, seed=it, mean=0.1 * it, std=1.0 + 0.1 * it)
        xn = (x - w["data_mean"]) / np.float32(w["data_std"])
        ref_codes, ref_xhat = orc.encode(cfg, w, xn)
        model = QINCo(cfg, w, device="cuda:0")
        try:
            c = model(torch.from_numpy(x).cuda(), step="encode").cpu().numpy()
            dec = model(torch.from_numpy(ref_codes).cuda(), step="decode").cpu().numpy()
            model.synchronize()
            assert c.shape == ref_codes.shape
            assert rel_mse(dec, orc.forward(cfg, w, ref_codes, "decode")) <= DEC_TOL, cfg
            agree = float((c == ref_codes).all(0).mean())
            mse_ours, mse_ref = orc.mse(xn, orc.decode(cfg, w, c)), orc.mse(xn, ref_xhat)
            assert agree >= 0.7 and abs(mse_ours - mse_ref) <= 1e-2 * mse_ref, (cfg, n, agree, mse_ours, mse_ref)
        finally:
            model._h.close()


def test_pq_qinco_matches_oracle():
    """PQ-QINCo (qinco_v1/model_qinco.py:185-234): two sub-quantizers with their own db_scale and an OPQ rotation, through
    the v1 codec loop (codec_qinco.py:25-72 drives any object with .encode/.decode/.db_scale)."""
    from qinco_b200 import codec
    rng = np.random.default_rng(5)
    subs, cfgs, ws, scales = [], [], [], (1.0, 2.5)
    for i, d in enumerate((16, 32)):
        cfg = synth.make_cfg(None, D=d, M=3, K=64, L=1, de=d, dh=32, A=0, B=1, qinco1_mode=True)
        w = synth.make_weights(cfg, seed=40 + i, n_train=1024, kmeans_iters=1)
        cfgs.append(cfg); ws.append(w)
        subs.append(codec.QINCoV1(cfg=cfg, weights=w, db_scale=scales[i]))
    opq, _ = np.linalg.qr(rng.standard_normal((48, 48)))
    opq = opq.astype(np.float32)
    model = codec.PQQINCoV1(subs, opq_matrix=opq)
    x = rng.standard_normal((203, 48)).astype(np.float32)
    codes = codec.encode(model, x, bs=64, verbose=False)
    y = codec.decode(model, codes, bs=50, verbose=False)
    assert codes.shape == (203, 6) and y.shape == (203, 48)
    # oracle: rotate, quantise each slice in its own scale, rotate back
    xr = x @ opq.T
    ref_codes, ref = [], np.zeros_like(xr)
    d0 = 0
    for cfg, w, sc in zip(cfgs, ws, scales):
        d1 = d0 + cfg["D"]
        c, xh = orc.encode(cfg, w, xr[:, d0:d1] / np.float32(sc))
        ref_codes.append(c.T)
        ref[:, d0:d1] = xh * np.float32(sc)
        d0 = d1
    ref_codes, ref = np.concatenate(ref_codes, 1), ref @ opq
    assert (codes == ref_codes).all(1).mean() >= 0.9
    same = (codes == ref_codes).all(1)
    assert rel_mse(y[same], ref[same]) <= DEC_TOL
    for q in subs:
        q._m._h.close()


# ------------------------------------------------------------------------------------- BASELINE configs, full depth
@pytest.mark.parametrize("name", GOLDEN_FULL)
def test_full_depth_fixture_matches_reference(name, golden_loader):
    """The BASELINE configurations at their real depth (L = 16, every step; 64 rows produced by the unmodified
    reference): decode of the reference's codes within 1e-4, encode MSE / code agreement against the reference."""
    from qinco_b200.model import QINCo
    cfg, w, z = golden_loader(name)
    model = QINCo(cfg, w, device="cuda:0")
    try:
        x = z["x"]
        xn = (x - w["data_mean"]) / np.float32(w["data_std"])
        dec = model(torch.from_numpy(z["codes_ref"]).cuda(), step="decode")
        model.synchronize()
        err = rel_mse(dec.cpu().numpy(), z["dec_ref"])
        assert err <= DEC_TOL, f"{name}: decode rel mse {err:.3e}"
        codes, xhat = model.encode(torch.from_numpy(xn).cuda())
        model.synchronize()
        c = codes.cpu().numpy()
        agree = float((c == z["codes_ref"]).all(0).mean())
        from oracle.torch_port import TorchPort
        port = TorchPort(cfg, w)
        d_ours = ((xn - port.decode(c).numpy()) ** 2).sum(1)
        d_ref = ((xn - z["xhat_ref"]) ** 2).sum(1)
        rel = abs(d_ours.mean() - d_ref.mean()) / d_ref.mean()
        print(f"\n{name}: identical vectors {agree:.3f}, mse ours {d_ours.mean():.5f} ref {d_ref.mean():.5f} rel {rel:.2e}, "
              f"worst vector {np.abs(d_ours - d_ref).max() / d_ref.mean():.2e}")
        assert rel <= FULL_TOL[name][0] and agree >= FULL_TOL[name][1]
        assert rel_mse(xhat.cpu().numpy(), model.decode(codes).cpu().numpy()) <= ENC_DEC_TOL
    finally:
        model._h.close()


# (|MSE_ours - MSE_ref| / MSE_ref, fraction of vectors with identical codes) accepted on the 64-row full-depth fixtures
FULL_TOL = {"full_q1": (5e-3, 0.8), "full_l_b16": (5e-3, 0.8), "full_deep_m16": (5e-3, 0.7), "full_contr_b32": (5e-3, 0.7)}

# SURVEY section 8(d) "Parity subsets": the first 10 000 rows (S / Q1) or 1 024 rows (L) of the workload's tensor
CONTRACT = [("q1", 10000), ("c2", 10000), ("c2a16", 10000), ("c3", 1024), ("c4", 1024), ("c5", 1024)]


@pytest.mark.parametrize("workload,rows", CONTRACT)
def test_encode_mse_at_contract_sample(workload, rows):
    """SURVEY section 8(d) pass criterion on every BASELINE configuration at the contract's sample size, against the
    PyTorch-CPU restatement of the reference (oracle/torch_port.py, bit-identical to the reference on all fixtures):
        |MSE_ours - MSE_ref| / MSE_ref <= 1e-4     (x-hat of OUR codes obtained with the REFERENCE decoder)
    plus decode within 1e-4 on the reference's codes.  Where codes differ (fp16 tensor-core operands flip near-ties),
    the per-vector errors must be statistically indistinguishable: mean signed difference within 3 standard errors."""
    import bench
    from oracle.torch_port import TorchPort
    from qinco_b200.model import QINCo
    wl = bench.WORKLOADS[workload]
    cfg, w, x = bench.make_model_inputs(wl, rows, 0)
    xs = x.numpy()
    port = TorchPort(cfg, w, threads=None)
    ref_codes, ref_xhat = port.encode(xs)
    ref_codes, ref_xhat = ref_codes.numpy(), ref_xhat.numpy()
    model = QINCo(cfg, w, device="cuda:0")
    try:
        ours = model(x.cuda(), step="encode").cpu().numpy()
        dec = model(torch.from_numpy(ref_codes).cuda(), step="decode").cpu().numpy()
        model.synchronize()
    finally:
        model._h.close()
    assert rel_mse(dec, port.decode(ref_codes).numpy()) <= DEC_TOL
    d_ref = ((xs - ref_xhat) ** 2).sum(1).astype(np.float64)
    d_ours = ((xs - port.decode(ours).numpy()) ** 2).sum(1).astype(np.float64)
    rel = abs(d_ours.mean() - d_ref.mean()) / d_ref.mean()
    same = (ours == ref_codes).all(0)
    delta = (d_ours - d_ref)[~same]
    sem = delta.std(ddof=1) / np.sqrt(len(delta)) if len(delta) > 1 else 0.0
    print(f"\n{workload} n={rows}: identical vectors {same.mean():.4f}, mse ours {d_ours.mean():.6f} ref {d_ref.mean():.6f} "
          f"rel {rel:.2e}; differing vectors: mean delta {delta.mean() if len(delta) else 0:.3e} +- {sem:.1e} "
          f"(relative to the mse: {(delta.mean() if len(delta) else 0) / d_ref.mean():.2e})")
    assert rel <= ENC_TOL_LARGE, (workload, rel)
    if len(delta) > 30:
        assert abs(delta.mean()) <= 4 * sem + 1e-12, "the differing vectors are systematically better or worse than the reference's"


# ------------------------------------------------------------------------------------------- single-launch decode
@pytest.mark.parametrize("name", GOLDEN_V2 + ["ivf_a8_b4", "ivf_a0_b1"])
def test_decode_loop_kernel(name, golden_loader, monkeypatch):
    """The one-launch decode (every tile walks all steps inside qb_mlp_kernel<.., kLoop>, u = Wx . xhat on the tensor core
    as fp16 hi/lo products) against the reference AND against the per-step launch sequence it replaces: same result to
    fp32 rounding, 1 launch instead of 2 (M - 1) + 1, ragged / single-row batches."""
    from qinco_b200.model import QINCo
    cfg, w, z = golden_loader(name)
    loop = QINCo(cfg, w, device="cuda:0")
    monkeypatch.setenv("QB_NO_DECODE_LOOP", "1")
    steps = QINCo(cfg, w, device="cuda:0")
    monkeypatch.delenv("QB_NO_DECODE_LOOP")
    try:
        assert loop._h.info(1)["decode_loop"] == 1 and steps._h.info(1)["decode_loop"] == 0
        codes = torch.from_numpy(z["codes_ref"]).cuda()
        l0 = loop.launch_count
        a = loop(codes, step="decode")
        n_launch = loop.launch_count - l0
        b = steps(codes, step="decode")
        loop.synchronize(); steps.synchronize()
        assert n_launch == 2, n_launch                  # the code pack kernel + ONE decode kernel
        assert rel_mse(a.cpu().numpy(), z["dec_ref"]) <= DEC_TOL
        assert rel_mse(a.cpu().numpy(), b.cpu().numpy()) <= ENC_DEC_TOL
        # a bigger, ragged batch of random codes (several tile sets per CTA), and single rows
        rng = np.random.default_rng(5)
        S = codes.shape[0]
        big = rng.integers(0, cfg["K"], size=(S, 40000 if cfg["D"] <= 32 else 3001))
        if cfg.get("ivf_K"):
            big[0] = rng.integers(0, cfg["ivf_K"], size=big.shape[1])
        bt = torch.from_numpy(big).cuda()
        ya, yb = loop.decode(bt), steps.decode(bt)
        loop.synchronize(); steps.synchronize()
        assert rel_mse(ya.cpu().numpy(), yb.cpu().numpy()) <= ENC_DEC_TOL
        assert torch.equal(loop.decode(bt[:, :1]), ya[:1]) and torch.equal(loop.decode(bt[:, 129:130]), ya[129:130])
        assert torch.equal(loop.decode(bt), ya)          # deterministic
    finally:
        loop._h.close(); steps._h.close()


# ------------------------------------------------------------------------------------------------ fused beam selection
@pytest.mark.parametrize("name,mode", [("s_a0_b1", "lite"), ("q1_l4", "lite"), ("ivf_a0_b1", "lite"), ("s_a0_b1", "full"),
                                       ("q1_l4", "full"), ("ivf_a0_b1", "full"), ("l_a16_b16", "b"), ("proj_a8_b4", "b"),
                                       ("full_l_b16", "b"), ("full_deep_m16", "b")])
def test_fused_selection_matches_unfused(name, mode, golden_loader, monkeypatch):
    """Selection fused into the score launch (distance -> arg-min inside the kernel, no `dist` array, no select launch;
    reference qinco_base.py:343-372) against the unfused launch sequence.  The arithmetic is the same, so codes AND xhat must
    be bit-identical.  `lite` (default for beam-1 resident launches): the winner's xhat' comes from a 1/256-size update
    launch; `full` (QB_FUSE_FULL=1): the score launch writes xhat' and the history itself; `b` (default for the QINCo2-L
    family): running top-F_out per vector and the winners' xhat' inside the score launch, tiles of a vector walked by one CTA."""
    from qinco_b200.model import QINCo
    cfg, w, z = golden_loader(name)
    if mode == "full":
        monkeypatch.setenv("QB_FUSE_FULL", "1")
    fused = QINCo(cfg, w, device="cuda:0")
    monkeypatch.delenv("QB_FUSE_FULL", raising=False)
    monkeypatch.setenv("QB_NO_FUSE", "1")
    plain = QINCo(cfg, w, device="cuda:0")
    monkeypatch.delenv("QB_NO_FUSE")
    try:
        ivf = bool(cfg.get("ivf_K"))
        enc = (lambda m, x: m.encode_ivf_u8(x)) if ivf else (lambda m, x: (None,) + m.encode_u8(x))
        steps = cfg["M"] - (0 if ivf else 1)
        for n in ((len(z["x"]), 1, 129, 39999) if mode != "b" else (len(z["x"]), 1, 300, 2049)):
            x = torch.from_numpy(synth.make_data(n, cfg["D"], seed=n)).cuda() if n != len(z["x"]) else \
                torch.from_numpy((z["x"] - w["data_mean"]) / np.float32(w["data_std"])).cuda()
            l0, p0 = fused.launch_count, plain.launch_count
            iv_a, c_a, x_a = enc(fused, x)
            n_fused = fused.launch_count - l0
            iv_b, c_b, x_b = enc(plain, x)
            n_plain = plain.launch_count - p0
            fused.synchronize(); plain.synchronize()
            assert torch.equal(c_a, c_b), (name, n, float((c_a == c_b).all(1).float().mean()))
            assert torch.equal(x_a, x_b)
            if ivf:
                assert torch.equal(iv_a, iv_b)
            if n < 128:          # one chunk: step 0, then per step prep + score (+ the update launch in lite mode) vs 4 launches
                assert n_plain == 1 + 4 * steps and n_fused == 1 + (3 if mode == "lite" else 2) * steps, (n_plain, n_fused)
                if mode == "b":
                    assert fused._h.info(1)["n_tiles"] == 1
        # codes only (no xhat wanted)
        c_only = fused(torch.from_numpy(z["x"]).cuda(), step="encode")
        fused.synchronize()
        np.testing.assert_array_equal(c_only.cpu().numpy(), plain(torch.from_numpy(z["x"]).cuda(), step="encode").cpu().numpy())
    finally:
        fused._h.close(); plain._h.close()


@pytest.mark.parametrize("name", ["s_a0_b1", "s_a16_b8", "proj_a8_b4", "l_a16_b16", "tiny_a0_b4"])
def test_weight_multicast_kernel_matches_default(name, golden_loader):
    """The weight-multicast variant (plan option mcast = 2: 2-CTA clusters, each CTA streams half of every slab into both
    rings) computes exactly what the single-CTA kernel computes: identical codes, xhat and decode, bit for bit."""
    from qinco_b200.model import QINCo
    cfg, w, z = golden_loader(name)
    on = QINCo(cfg, w, device="cuda:0", plan_opts={"n_tiles": 2 << 16})
    off = QINCo(cfg, w, device="cuda:0", plan_opts={"n_tiles": 1 << 16})
    try:
        rng = np.random.default_rng(3)
        for n in (len(z["x"]), 1, 700):
            x = torch.from_numpy(rng.standard_normal((n, cfg["D"]), dtype=np.float32)).cuda()
            c_on, x_on = on.encode(x)
            c_off, x_off = off.encode(x)
            on.synchronize(); off.synchronize()
            assert torch.equal(c_on, c_off) and torch.equal(x_on, x_off), (name, n)
            assert torch.equal(on.decode(c_on), off.decode(c_on))
        dec = on(torch.from_numpy(z["codes_ref"]).cuda(), step="decode")
        on.synchronize()
        assert rel_mse(dec.cpu().numpy(), z["dec_ref"]) <= DEC_TOL
    finally:
        on._h.close(); off._h.close()


def test_ivf_tensor_core_argmin_matches_fp32(monkeypatch):
    """The tensor-core IVF arg-min (fp16 hi/lo products, qb_ivf_tc_kernel) against the fp32 CUDA-core kernel and the oracle:
    20 000 vectors x 70 001 centroids at d = 128 (ragged last part), with normalisation.  Differences must be rounding ties."""
    from qinco_b200.model import QINCo
    cfg = synth.make_cfg(None, D=128, M=2, K=32, L=1, de=128, dh=32, A=0, B=1, ivf_K=70001)
    w = synth.make_weights(cfg, seed=3, n_train=1024, kmeans_iters=1, data_mean=0.1, data_std=1.3)
    x = synth.make_data(20000, 128, seed=8, mean=0.1, std=1.3)
    tc_model = QINCo(cfg, w, device="cuda:0")
    monkeypatch.setenv("QB_IVF_CC", "1")
    cc_model = QINCo(cfg, w, device="cuda:0")
    monkeypatch.delenv("QB_IVF_CC")
    try:
        xt = torch.from_numpy(x).cuda()
        a = tc_model(xt, step="encode")[0].cpu().numpy()
        b = cc_model(xt, step="encode")[0].cpu().numpy()
        tc_model.synchronize(); cc_model.synchronize()
        same = a == b
        assert same.mean() >= 0.9995, same.mean()
        xn = (x - w["data_mean"]) / np.float32(w["data_std"])
        cent = w["steps.0.ivf_centroids.weight"]
        idx = np.nonzero(~same)[0]
        if len(idx):          # where the two disagree the two centroids are tied to fp32 rounding
            da = ((xn[idx] - cent[a[idx]]).astype(np.float64) ** 2).sum(1)
            db = ((xn[idx] - cent[b[idx]]).astype(np.float64) ** 2).sum(1)
            assert np.all(np.abs(da - db) <= 2e-5 * db), (da, db)
        # against the oracle's fp32 arg-min on a slice (the numpy matrix is 2000 x 70001)
        ref = orc.approx_pairwise_distance(xn[:2000], cent).argmin(-1)
        assert (a[:2000] == ref).mean() >= 0.999
        # the centroid is the starting beam: decode of the IVF code alone reproduces it
        assert tc_model.launch_count > 0
    finally:
        tc_model._h.close(); cc_model._h.close()
