"""Generate tests/golden/*.npz by running the UNMODIFIED reference in the dev container.

    python oracle/make_golden.py            # needs /root/reference (read-only) + torch CPU

Each fixture holds: the model hyper-parameters, the seeds that regenerate the synthetic
weights (qinco_b200/synth.py) and their sha256, the input rows, and what the reference
returned for them:
    codes_ref    [M, n] int64   QINCo.encode / QINCo.forward(step="encode")  (qinco_base.py:454-485)
    xhat_ref     [n, D] fp32    x-hat returned by encode (normalised space)
    dec_ref      [n, D] fp32    forward(codes, step="decode") (data space)       (qinco_base.py:536-537)
    wrap_equal   bool           QINCoInferenceWrapper produced the same codes (only where it supports (A,B))
and for the v1 case the outputs of qinco_v1/codec_qinco.py encode()/decode().
Small models also carry their weights so the fixture is self-contained.
"""
from __future__ import annotations

import io
import json
import os
import sys
import contextlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from qinco_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# name -> (cfg kwargs, n rows, data mean/std, weight seed)
CASES = {
    # tiny shapes: cheap oracle pins (weights stored in the fixture)
    "tiny_q1":        (dict(D=16, M=4, K=64, L=2, de=16, dh=32, A=0, B=1, qinco1_mode=True), 96, (0.0, 1.0), 11),
    "tiny_a0_b1":     (dict(D=16, M=4, K=64, L=2, de=16, dh=32, A=0, B=1, qinco1_mode=False), 96, (0.5, 2.0), 12),
    "tiny_a8_b4":     (dict(D=16, M=5, K=64, L=2, de=32, dh=48, A=8, B=4, qinco1_mode=False), 96, (0.0, 1.0), 13),
    "tiny_a0_b4":     (dict(D=16, M=4, K=64, L=1, de=16, dh=32, A=0, B=4, qinco1_mode=False), 64, (0.0, 1.0), 14),
    # shapes the CUDA path is built for (weights regenerated from the seed, digest pinned)
    "s_a0_b1":        (dict(D=128, M=8, K=256, L=2, de=128, dh=256, A=0, B=1, qinco1_mode=False), 64, (0.0, 1.0), 21),
    "s_a16_b8":       (dict(D=128, M=8, K=256, L=2, de=128, dh=256, A=16, B=8, qinco1_mode=False), 48, (0.0, 1.0), 22),
    "s_a0_b4":        (dict(D=128, M=4, K=256, L=2, de=128, dh=256, A=0, B=4, qinco1_mode=False), 24, (0.0, 1.0), 23),
    "proj_a8_b4":     (dict(D=96, M=4, K=256, L=3, de=192, dh=160, A=8, B=4, qinco1_mode=False), 40, (0.25, 1.5), 24),
    "q1_l4":          (dict(D=128, M=4, K=256, L=4, de=128, dh=256, A=0, B=1, qinco1_mode=True), 32, (0.0, 1.0), 25),
    "l_a16_b16":      (dict(D=128, M=4, K=256, L=4, de=384, dh=384, A=16, B=16, qinco1_mode=False), 16, (0.0, 1.0), 26),
    # the BASELINE configurations at FULL depth (L = 16, all M steps), 64 rows each: QINCo1 preset (config 1), QINCo2-L
    # 8x8 beam 16 (config 3), the Deep1B shape 16x8 d=96 (config 4), the Contriever shape d=768 beam 32 (config 5)
    "full_q1":        (dict(D=128, M=8, K=256, L=16, de=128, dh=256, A=0, B=1, qinco1_mode=True), 64, (0.0, 1.0), 41),
    "full_l_b16":     (dict(D=128, M=8, K=256, L=16, de=384, dh=384, A=16, B=16, qinco1_mode=False), 64, (0.0, 1.0), 42),
    "full_deep_m16":  (dict(D=96, M=16, K=256, L=16, de=384, dh=384, A=16, B=16, qinco1_mode=False), 64, (0.0, 1.0), 43),
    "full_contr_b32": (dict(D=768, M=8, K=256, L=16, de=384, dh=384, A=16, B=32, qinco1_mode=False), 64, (0.0, 1.0), 44),
    # IVF-QINCo: step 0 = arg-min over ivf_K centroids, then M implicit-codebook steps; codes [M+1, n]
    "ivf_a8_b4":      (dict(D=32, M=3, K=64, L=2, de=32, dh=48, A=8, B=4, qinco1_mode=False, ivf_K=200), 96, (0.25, 1.5), 31),
    "ivf_a4_b8":      (dict(D=32, M=3, K=64, L=1, de=48, dh=32, A=4, B=8, qinco1_mode=False, ivf_K=77), 64, (0.0, 1.0), 32),
    "ivf_a0_b1":      (dict(D=128, M=3, K=256, L=2, de=128, dh=256, A=0, B=1, qinco1_mode=False, ivf_K=1000), 48, (0.0, 1.0), 33),
}
V1_CASE = ("v1_codec", dict(D=128, M=8, K=256, L=2, de=128, dh=256, A=0, B=1, qinco1_mode=True), 96, 3.5, 31)

DATA_SEED = 1234
GAIN = 0.5


def gen_case(name, kw, n, mean_std, wseed):
    import torch
    cfg = synth.make_cfg(None, **kw)
    small = cfg["D"] <= 32
    w = synth.make_weights(cfg, seed=wseed, gain=GAIN, n_train=2048 if small else 4096, kmeans_iters=2,
                           data_mean=mean_std[0], data_std=mean_std[1])
    x = synth.make_data(n, cfg["D"], seed=DATA_SEED + wseed, mean=mean_std[0], std=mean_std[1])
    model = ref_loader.build_v2(cfg, w)
    with torch.no_grad():
        xt = torch.from_numpy(x)
        codes_fwd = model(xt, step="encode")
        codes, xhat = model.encode((xt - model.data_mean) / model.data_std)
        assert torch.equal(codes, codes_fwd)
        dec = model(codes, step="decode")
        # invariant: encode's x-hat is decode(codes) (exact for small GEMM shapes,
        # fp32 rounding apart once the batched shapes pick different BLAS kernels)
        enc_dec_gap = float((model.decode(codes) - xhat).abs().max())
        assert enc_dec_gap <= 1e-4 * float(xhat.abs().max()), enc_dec_gap
    wrap_equal = -1
    if (cfg["A"] > 0 or cfg["B"] == 1) and not cfg.get("ivf_K"):      # the wrapper's A=0 encoder assumes B=1 (qinco_inference.py:126)
        wrap = ref_loader.build_v2(cfg, w, inference=True)
        with torch.no_grad():
            wrap_equal = int(torch.equal(wrap(xt, step="encode"), codes))
    out = dict(
        cfg=json.dumps(cfg), wseed=wseed, gain=GAIN, n_train=2048 if small else 4096, kmeans_iters=2,
        data_mean=mean_std[0], data_std=mean_std[1], digest=synth.weights_digest(w),
        x=x, codes_ref=codes.numpy().astype(np.int64), xhat_ref=xhat.numpy(), dec_ref=dec.numpy(),
        wrap_equal=wrap_equal, enc_dec_gap=enc_dec_gap,
    )
    if small:
        for k, v in w.items():
            out["w:" + k] = v
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(f"{name}: n={n} wrap_equal={wrap_equal} mse={float(((xhat - (xt - model.data_mean) / model.data_std) ** 2).sum(1).mean()):.4f}")


def gen_v1():
    name, kw, n, db_scale, wseed = V1_CASE
    cfg = synth.make_cfg(None, **kw)
    w = synth.make_weights(cfg, seed=wseed, gain=GAIN, n_train=4096, kmeans_iters=2)
    x = synth.make_data(n, cfg["D"], seed=DATA_SEED + wseed, std=db_scale)
    model = ref_loader.build_v1(cfg, synth.to_v1_state(cfg, w), db_scale=db_scale)
    codec = ref_loader.v1_codec()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        codes = codec.encode(model, x, bs=40, is_float16=False)        # ragged last batch on purpose
        dec = codec.decode(model, codes, bs=40, is_float16=False)
    mse_line = [l for l in buf.getvalue().replace("\r", "\n").splitlines() if "MSE=" in l][-1]
    mse_ref = float(mse_line.split("MSE=")[1])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), cfg=json.dumps(cfg), wseed=wseed, gain=GAIN, n_train=4096,
                        kmeans_iters=2, db_scale=db_scale, digest=synth.weights_digest(w), x=x,
                        codes_ref=codes.astype(np.int64), dec_ref=dec, mse_ref=mse_ref)
    print(f"{name}: n={n} codes{codes.shape} mse_ref={mse_ref:g}")


PAIRWISE_CASE = ("pairwise_ivf", dict(D=32, M=4, K=16, Mt=8, ivf_K=40), 77)


def gen_pairwise():
    import torch
    name, kw, n = PAIRWISE_CASE
    book, comb, imap = synth.make_pairwise_tables(seed=2024, **kw)
    rng = np.random.default_rng(5)
    codes = rng.integers(0, kw["K"], (kw["M"], n)).astype(np.int64)
    ivf = rng.integers(0, kw["ivf_K"], n).astype(np.int64)
    ref = ref_loader.build_pairwise(book, comb, imap, kw["K"])
    with torch.no_grad():
        out = ref(torch.from_numpy(codes), torch.from_numpy(ivf)).numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), cfg=json.dumps(kw), seed=2024, codes=codes, ivf_codes=ivf, out_ref=out)
    print(f"{name}: n={n} out{out.shape}")


if __name__ == "__main__":
    assert ref_loader.available(), "needs /root/reference"
    os.makedirs(OUT, exist_ok=True)
    only = set(sys.argv[1:])
    for name, (kw, n, ms, ws) in CASES.items():
        if not only or name in only:
            gen_case(name, kw, n, ms, ws)
    if not only or V1_CASE[0] in only:
        gen_v1()
    if not only or PAIRWISE_CASE[0] in only:
        gen_pairwise()
