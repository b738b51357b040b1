"""CPU oracle: numpy fp32 restatement of the QINCo / QINCo2 encode-decode loop.

TEST INFRASTRUCTURE ONLY.  Nothing under qinco_b200/ imports this module; it is
used by tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline /
`--impl reference` legs, always as the checker or the timed CPU baseline and
never as the product path.

Parity status: the reference (facebookresearch/Qinco @ 5a324954) has no tests,
golden vectors or known-answer fixtures for this path, so
the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, run in the dev
container by oracle/make_golden.py and committed under tests/golden/*.npz
(tests/test_oracle_golden.py: codes identical, x-hat within fp32 rounding).

Every function cites the reference lines it restates (paths relative to
/root/reference).  State-dict keys are the reference's own (qinco_b200/synth.py).
"""
from __future__ import annotations

import numpy as np

AUTO_MAX_K_VALUE = 32  # qinco/utils.py:298


# ---------------------------------------------------------------------------
# distances (qinco/utils.py:301-388, qinco_v1/utils.py:28-52)
# ---------------------------------------------------------------------------
def approx_pairwise_distance(a, b):
    """|a|^2 + |b|^2 - 2 a.b^T   (qinco/utils.py:336-346; v1 utils.py:28-38)."""
    return (a * a).sum(-1)[:, None] + (b * b).sum(-1)[None, :] - 2.0 * (a @ b.T)


def exact_pairwise_distance(a, b):
    """sum((a-b)^2)   (qinco/utils.py:325-333)."""
    return ((a[:, None, :] - b[None, :, :]) ** 2).sum(-1)


def pairwise_distances(a, b, approx="auto"):
    """qinco/utils.py:301-322 — approx form only when both sides exceed 32 rows."""
    if approx == "auto":
        approx = len(a) > AUTO_MAX_K_VALUE and len(b) > AUTO_MAX_K_VALUE
    return approx_pairwise_distance(a, b) if approx else exact_pairwise_distance(a, b)


def compute_batch_distances(a, b, approx="auto"):
    """a:[n,1,D], b:[n,R,D] -> [n,R]   (qinco/utils.py:349-388; chunking there only bounds memory)."""
    n, one, D = a.shape
    assert one == 1
    R = b.shape[1]
    if approx == "auto":
        approx = one > AUTO_MAX_K_VALUE or R > AUTO_MAX_K_VALUE
    if not approx:
        return ((a - b) ** 2).sum(-1)
    an = (a * a).sum(-1)  # [n,1]
    bn = (b * b).sum(-1)  # [n,R]
    ab = np.einsum("nd,nrd->nr", a[:, 0, :], b, optimize=True)
    return an + bn - 2.0 * ab


def topk_smallest(d, k):
    """indices of the k smallest along the last axis, ascending (torch.topk(largest=False); ties -> lower index)."""
    if k == 1:
        return d.argmin(-1)[..., None]
    return np.argsort(d, axis=-1, kind="stable")[..., :k]


# ---------------------------------------------------------------------------
# the implicit-codebook MLP  f_m(c, xhat)
# ---------------------------------------------------------------------------
def step_mlp(cfg, w, m, c, xhat):
    """QINCoStep.forward (qinco/model/qinco_base.py:262-280) with QConcat (:60-64) and QBlockFFN (:93-97).

    c, xhat: [..., D] fp32.  Returns the conditioned codeword [..., D].
    QINCo1 mode drops the outer skip (:276-278 / qinco_inference.py:29,40); the v1 model
    (qinco_v1/model_qinco.py:39-47) is the same expression with De == D.
    """
    p = f"steps.{m}."
    D, De = cfg["D"], cfg["de"]
    e = c @ w[p + "in_proj.weight"].T if De != D else c
    cc = np.concatenate([e, np.broadcast_to(xhat, e.shape[:-1] + (D,))], axis=-1)
    e = e + (cc @ w[p + "concat.mlp.weight"].T + w[p + "concat.mlp.bias"])
    for l in range(cfg["L"]):
        h = np.maximum(e @ w[p + f"residual_blocks.{l}.up_proj.weight"].T, 0.0)
        e = e + h @ w[p + f"residual_blocks.{l}.down_proj.weight"].T
    o = e @ w[p + "out_proj.weight"].T if De != D else e
    if not cfg["qinco1_mode"]:
        o = o + c
    return o.astype(np.float32, copy=False)


# ---------------------------------------------------------------------------
# decode (qinco_base.py:447-452, 282-290; qinco_inference.py:66-75; v1 model_qinco.py:91-95)
# ---------------------------------------------------------------------------
def n_steps(cfg):
    """Number of quantisation steps = rows of the code matrix: cfg._M_ivf = M (+1 with an IVF first step)."""
    return cfg["M"] + (1 if cfg.get("ivf_K") else 0)


def decode(cfg, w, codes_MB):
    """codes [M_ivf, n] int -> xhat [n, D] fp32, normalised space."""
    codes_MB = np.asarray(codes_MB).astype(np.int64)
    S = n_steps(cfg)
    assert codes_MB.shape[0] == S
    if cfg.get("ivf_K"):                                 # IVFBook.decode (qinco_base.py:176-183): centroid lookup
        xhat = w["steps.0.ivf_centroids.weight"][codes_MB[0]].astype(np.float32)
    else:
        xhat = w["steps.0.codebook.weight"][codes_MB[0]].astype(np.float32)
    for m in range(1, S):
        c = w[f"steps.{m}.codebook.weight"][codes_MB[m]]
        xhat = xhat + step_mlp(cfg, w, m, c, xhat)
    return xhat


# ---------------------------------------------------------------------------
# encode (qinco_base.py:454-485 loop; :292-374 one beam step; :114-121 pre-selection)
# ---------------------------------------------------------------------------
def _encode_step(cfg, w, m, x, xhat_BFD, hist):
    """One beam step.  x:[n,D]; xhat_BFD:[n,F,D]; hist: list of [n,F] int64 (one per earlier step)."""
    K, D, M, A, Bw = cfg["K"], cfg["D"], n_steps(cfg), cfg["A"], cfg["B"]
    n, F_in, _ = xhat_BFD.shape
    F_out = Bw if m < M - 1 else 1                       # qinco_base.py:310 (M = cfg._M_ivf)
    if m == 0 and cfg.get("ivf_K"):                      # IVFBook.encode / quantize (qinco_base.py:148-174): F = 1, arg-min
        cent = w["steps.0.ivf_centroids.weight"]
        codes = approx_pairwise_distance(x, cent).argmin(-1)
        return cent[codes].reshape(n, 1, D).astype(np.float32), [codes.reshape(n, 1).astype(np.int64)]
    cb = w[f"steps.{m}.codebook.weight"]
    if m == 0:                                           # codebook_only (:218,:263): candidates are raw codewords
        cand = np.broadcast_to(cb[None, None], (n, F_in, K, D)) + xhat_BFD[:, :, None, :]
        idx = None
        C = K
    else:
        if A > 0:                                        # :316-324, substep :114-121
            n_codes = max(A, Bw) if (m == 1 and cfg.get("ivf_K")) else A   # qinco_base.py:108-112
            r = (x[:, None, :] - xhat_BFD).reshape(n * F_in, D)
            d_pre = pairwise_distances(r, w[f"steps.{m}.substep.codebook.weight"])
            idx = topk_smallest(d_pre, n_codes).reshape(n, F_in, n_codes)
            c = cb[idx]                                  # [n,F,A,D]
            C = n_codes
        else:                                            # :326 all K codewords
            idx = None
            c = np.broadcast_to(cb[None, None], (n, F_in, K, D))
            C = K
        xh = xhat_BFD[:, :, None, :]
        cand = step_mlp(cfg, w, m, c, xh) + xh           # :329-335
    cand = cand.reshape(n, F_in * C, D)
    dist = compute_batch_distances(x[:, None, :], cand)  # :343-345
    sel = topk_smallest(dist, F_out)                     # :346   [n,F_out] flat = f*C + a
    f_par, a_sel = sel // C, sel % C
    code = np.take_along_axis(idx.reshape(n, F_in * C), sel, axis=1) if idx is not None else a_sel  # :349-354
    new_hist = [np.take_along_axis(h, f_par, axis=1) for h in hist] + [code]                          # :357-372
    xhat_next = np.take_along_axis(cand, sel[:, :, None], axis=1)                                     # :363-369
    return xhat_next.astype(np.float32, copy=False), new_hist


def encode(cfg, w, x, max_rows=65536):
    """x [n, D] fp32 (normalised space) -> (codes [M, n] int64, xhat [n, D] fp32).

    Chunked over vectors so that at most `max_rows` candidate rows are alive, like
    enc_max_bs in the reference (qinco_base.py:456-472); chunking does not change results.
    """
    x = np.ascontiguousarray(x, dtype=np.float32)
    n = len(x)
    per_vec = cfg["B"] * (cfg["A"] or cfg["K"])
    bs = max(1, max_rows // per_vec)
    codes = np.empty((n_steps(cfg), n), np.int64)
    xhat = np.empty((n, cfg["D"]), np.float32)
    for i0 in range(0, n, bs):
        xb = x[i0:i0 + bs]
        xh = np.zeros((len(xb), 1, cfg["D"]), np.float32)  # :475
        hist = []
        for m in range(n_steps(cfg)):
            xh, hist = _encode_step(cfg, w, m, xb, xh, hist)
        assert xh.shape[1] == 1                            # :480-482
        codes[:, i0:i0 + bs] = np.stack([h[:, 0] for h in hist])
        xhat[i0:i0 + bs] = xh[:, 0]
    return codes, xhat


# ---------------------------------------------------------------------------
# the public surfaces
# ---------------------------------------------------------------------------
def forward(cfg, w, x_in, step):
    """QINCo.forward / QINCoInferenceWrapper.forward (qinco_base.py:524-539, qinco_inference.py:272-283)."""
    mean, std = w["data_mean"], np.float32(w["data_std"])
    if step == "encode":
        codes, _ = encode(cfg, w, (np.asarray(x_in, np.float32) - mean) / std)
        return codes
    if step == "decode":
        return decode(cfg, w, x_in) * std + mean
    raise ValueError(f"{step=}")


def codec_encode(cfg, w, x, bs, db_scale=1.0):
    """qinco_v1/codec_qinco.py:25-46 — numpy [N,D] -> (codes [N,M] int64, MSE in data space)."""
    out, err = [], 0.0
    for i0 in range(0, len(x), bs):
        batch = np.asarray(x[i0:i0 + bs], np.float32) / np.float32(db_scale)
        codes, recons = encode(cfg, w, batch)
        err += float(((recons - batch) ** 2).sum()) * db_scale ** 2
        out.append(codes.T)
    return np.concatenate(out), err / len(x)


def codec_decode(cfg, w, codes, bs, db_scale=1.0):
    """qinco_v1/codec_qinco.py:54-72 — numpy [N,M] -> [N,D] fp32."""
    out = []
    for i0 in range(0, len(codes), bs):
        out.append(decode(cfg, w, np.asarray(codes[i0:i0 + bs]).T) * np.float32(db_scale))
    return np.concatenate(out)


def mse(x, xhat):
    """mean over vectors of the squared error (qinco/utils.py:87-97)."""
    return float(((np.asarray(x, np.float64) - np.asarray(xhat, np.float64)) ** 2).sum(-1).mean())


def pairwise_decode(codebook_MKD, combine_mvals_m, ivf_code_map, K_base, codes_MB, ivf_codes):
    """PairwiseDecoderIVF.forward + map_codes (reference qinco/search/pairwise_decoder.py:88-93, :126-130), numpy fp32.

    codes_MB [M, n] ints, ivf_codes [n] ints -> [n, D] float32; the rows are added in table order, like the reference."""
    codes_MB = np.asarray(codes_MB, np.int64)
    ext = np.concatenate([codes_MB, np.asarray(ivf_code_map, np.int64)[np.asarray(ivf_codes, np.int64)].T])   # :128
    comb = ext[np.asarray(combine_mvals_m[0], np.int64)] * int(K_base) + ext[np.asarray(combine_mvals_m[1], np.int64)]   # :129
    book = np.asarray(codebook_MKD, np.float32)
    xhat = book[0][comb[0]].copy()                                                                               # :90
    for j in range(1, book.shape[0]):
        xhat += book[j][comb[j]]                                                                                 # :91-92
    return xhat
