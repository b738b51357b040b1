"""CPU tests of the host logic: the C-ABI library loads and exports every declared symbol, fails loudly without a GPU,
and the tcgen05 step planner + weight packer are right — checked by replaying the op list in numpy (a software model
of the kernel's dataflow: fp16 operands, fp32 accumulators in 'TMEM' columns) against the oracle's MLP.
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import qinco_oracle as orc
from qinco_b200 import _lib, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

OP_DTYPE = np.dtype([("w_off", "<u4"), ("slab_bytes", "<u4"), ("last_bytes", "<u4"), ("n", "<u2"), ("ks", "<u2"),
                     ("k_total", "<u2"), ("a_off", "<u2"), ("d_col", "<u2"), ("n_slab", "u1"), ("a_src", "u1"),
                     ("accumulate", "u1"), ("wait_a", "u1"), ("wait_d", "u1"), ("commit", "u1"), ("wait_a2_slab", "u1"),
                     ("pad", "u1", 3)])
PLAN_FIELDS = ["D", "De", "Dh", "L", "K", "has_proj", "skip", "n_tiles", "tmem_alloc_cols", "n_ops_block", "n_ops_out",
               "hc", "n_hchunk", "oc", "n_ochunk", "tmem_e_col", "tmem_h_col", "tmem_tile_cols", "smem_tres", "smem_ring",
               "slot_bytes", "n_stage", "smem_total", "block_w_bytes", "w_blob_bytes", "pair", "h_split", "n_ops_pre", "ae_chunks",
               "e_split", "mcast", "plan_view"]
A_E, A_H = 0, 1
BAR_AE_READY, BAR_AH_READY, BAR_HACC_FREE, BAR_HACC_FULL, BAR_EACC_FULL = 1, 2, 3, 4, 5


@pytest.fixture(scope="module")
def lib():
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "qinco_b200.h")).read()
    declared = set(re.findall(r"\b(qb_[a-z0-9_]+)\s*\(", header))
    declared -= {"qb_model_desc", "qb_status"}
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.qb_version() >= 100
    assert OP_DTYPE.itemsize == 32


def test_no_cpu_fallback(lib):
    """Without a CUDA device the product path must fail loudly (QB_ERR_CUDA), never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = synth.make_cfg(None, D=16, M=2, K=32, L=1, de=16, dh=16)
    w = synth.make_weights(cfg, seed=1, n_train=256, kmeans_iters=1)
    with pytest.raises(_lib.QbError) as e:
        _lib.Handle(cfg, w)
    assert e.value.code == -2
    from qinco_b200.model import QINCo
    with pytest.raises(RuntimeError):
        QINCo(cfg, w, device="cuda:0")
    with pytest.raises(RuntimeError):
        QINCo(cfg, w, device="cpu")


def test_invalid_descriptions_are_rejected(lib):
    cfg = synth.make_cfg(None, D=24, M=2, K=32, L=1, de=24, dh=16)   # D not a multiple of 16
    w = synth.make_weights(cfg, seed=1, n_train=256, kmeans_iters=1)
    with pytest.raises(_lib.QbError) as e:
        _lib.Handle(cfg, w)
    assert e.value.code == -1 and "multiple of 16" in str(e.value)


def export_plan(lib, cfg, opts=None):
    o = (C.c_int32 * 5)(*(opts or [0] * 5))
    plan = (C.c_int32 * 32)()
    ops = np.zeros(256, OP_DTYPE)
    n = lib.qb_plan_export(cfg["D"], cfg["de"], cfg["dh"], cfg["L"], cfg["K"], int(cfg["qinco1_mode"]), o, plan, 32,
                           ops.ctypes.data_as(C.c_void_p), 256)
    assert n >= 0, n
    return dict(zip(PLAN_FIELDS, list(plan))), ops[:n]


def pack(lib, cfg, w, m, plan, opts=None):
    o = (C.c_int32 * 5)(*(opts or [0] * 5))
    L = cfg["L"]
    ups = _lib._PtrArray([w[f"steps.{m}.residual_blocks.{l}.up_proj.weight"] for l in range(L)])
    downs = _lib._PtrArray([w[f"steps.{m}.residual_blocks.{l}.down_proj.weight"] for l in range(L)])
    outp = _lib._f32(w[f"steps.{m}.out_proj.weight"]) if cfg["de"] != cfg["D"] else None
    blob = np.zeros((plan["w_blob_bytes"] + 1) // 2, np.uint16)
    fp = C.POINTER(C.c_float)
    lib.qb_plan_pack.argtypes = [C.c_int] * 6 + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    rc = lib.qb_plan_pack(cfg["D"], cfg["de"], cfg["dh"], L, cfg["K"], int(cfg["qinco1_mode"]), C.cast(o, C.c_void_p),
                          C.cast(ups.arr, C.c_void_p), C.cast(downs.arr, C.c_void_p),
                          outp.ctypes.data_as(C.c_void_p) if outp is not None else None,
                          blob.ctypes.data_as(C.c_void_p), len(blob))
    assert rc == 0, rc
    return blob


def tables(lib, cfg, w, m):
    D, De, K = cfg["D"], cfg["de"], cfg["K"]
    t_blk, cb_blk, wx_t = np.zeros(De * K, np.float32), np.zeros(D * K, np.float32), np.zeros(D * De, np.float32)
    inp = _lib._f32(w[f"steps.{m}.in_proj.weight"]) if De != D else None
    vp = C.c_void_p
    lib.qb_plan_tables.argtypes = [C.c_int] * 3 + [vp] * 7
    p = lambda a: a.ctypes.data_as(vp) if a is not None else None  # noqa: E731
    lib.qb_plan_tables(D, De, K, p(_lib._f32(w[f"steps.{m}.codebook.weight"])), p(inp),
                       p(_lib._f32(w[f"steps.{m}.concat.mlp.weight"])), p(_lib._f32(w[f"steps.{m}.concat.mlp.bias"])),
                       p(t_blk), p(cb_blk), p(wx_t))
    T = t_blk.reshape(De // 4, K, 4).transpose(1, 0, 2).reshape(K, De)
    CB = cb_blk.reshape(D // 4, K, 4).transpose(1, 0, 2).reshape(K, D)
    return T, CB, wx_t.reshape(D, De)


def f16(a):
    return a.astype(np.float16).astype(np.float32)


def replay(plan, ops, blob, T, CB, WxT, codes, xhat):
    """Software model of qb_mlp_kernel for one tile of rows: returns xhat + f_m(C_m[code], xhat).

    Every op is replayed slab by slab exactly as the MMA warp walks it (K=16 MMAs over ring slots), so the slab
    offsets / sizes the producer streams are checked together with the operand offsets.
    """
    n = len(codes)
    De, D, L = plan["De"], plan["D"], plan["L"]
    hcol = plan["tmem_h_col"]
    tmem = np.zeros((n, 512), np.float32)
    e = T[codes] + xhat @ WxT                          # init epilogue
    tmem[:, :De] = e
    state = dict(AE=None, AH=None, ae_pending=f16(e), hw=0, oq=0)
    out = np.zeros((n, D), np.float32)
    bytes_view = blob.view(np.uint8)

    def run(op, base, out_phase):
        if op["wait_a"] == BAR_AE_READY:
            state["AE"] = state["ae_pending"]
        elif op["wait_a"] == BAR_AH_READY:
            state["AH"] = f16(np.maximum(tmem[:, hcol:hcol + state["hw"]], 0))   # in-place packed fp16 operand
        nn, ks, kt = int(op["n"]), int(op["ks"]), int(op["k_total"])
        pair = bool(plan["pair"])
        assert op["slab_bytes"] == nn * ks * 2 <= plan["slot_bytes"] * (2 if pair else 1)
        if pair:       # cta_group::2 shapes: N % 16 with A in shared memory, N % 32 with A in TMEM
            assert nn % (32 if op["a_src"] == A_H else 16) == 0
        assert op["n_slab"] == -(-kt // ks) and op["last_bytes"] == nn * (kt - (int(op["n_slab"]) - 1) * ks) * 2
        if op["wait_a2_slab"]:     # H chunk handed over in two K halves: the second wait sits exactly on the half boundary
            assert plan["h_split"] and op["a_src"] == A_H and op["wait_a"] == BAR_AH_READY
            assert int(op["wait_a2_slab"]) * ks == kt // 2 and kt == plan["hc"]
        c0 = int(op["d_col"])
        acc = bool(op["accumulate"])
        for s in range(int(op["n_slab"])):
            kk = min(ks, kt - s * ks)
            off = base + int(op["w_off"]) + s * int(op["slab_bytes"])
            slab = bytes_view[off: off + nn * kk * 2].view(np.float16).astype(np.float32)
            if pair:   # CTA r of the pair streams rows [r n/2, (r+1) n/2), each half laid out [k/8][n/2][8]
                W = np.concatenate([h.reshape(kk // 8, nn // 2, 8).transpose(1, 0, 2).reshape(nn // 2, kk)
                                    for h in slab.reshape(2, -1)])
            else:
                W = slab.reshape(kk // 8, nn, 8).transpose(1, 0, 2).reshape(nn, kk)
            if op["a_src"] == A_E:
                k0 = int(op["a_off"]) * 8 + s * ks
                a = state["AE"][:, k0:k0 + kk]
            else:
                k0 = (int(op["a_off"]) - hcol) * 2 + s * ks
                a = state["AH"][:, k0:k0 + kk]
            assert a.shape[1] == kk
            prod = a @ W.T
            if acc:
                tmem[:, c0:c0 + nn] += prod
            else:
                tmem[:, c0:c0 + nn] = prod
            acc = True
        if op["commit"] == BAR_HACC_FULL:
            state["hw"] = nn
            if out_phase:
                q = state["oq"]
                out[:, q * plan["oc"]: q * plan["oc"] + nn] = tmem[:, hcol:hcol + nn]
                state["oq"] += 1

    for l in range(L):
        for op in ops[:plan["n_ops_block"]]:
            run(op, l * plan["block_w_bytes"], False)
        assert ops[plan["n_ops_block"] - 1]["commit"] == BAR_EACC_FULL
        state["ae_pending"] = f16(tmem[:, :De])
    for op in ops[plan["n_ops_block"]:]:
        run(op, 0, True)
    if plan["has_proj"]:
        assert state["oq"] == plan["n_ochunk"]
        o = out
    else:
        o = tmem[:, :D].copy()
    if plan["skip"]:
        o = o + CB[codes]
    return xhat + o


SHAPES = {
    "S": dict(D=128, M=2, K=256, L=2, de=128, dh=256, A=0, B=1, qinco1_mode=False),
    "Q1": dict(D=128, M=2, K=256, L=3, de=128, dh=256, A=0, B=1, qinco1_mode=True),
    "L": dict(D=128, M=2, K=256, L=2, de=384, dh=384, A=0, B=1, qinco1_mode=False),
    "deep": dict(D=96, M=2, K=256, L=2, de=384, dh=384, A=0, B=1, qinco1_mode=False),
    "contriever": dict(D=768, M=2, K=64, L=1, de=384, dh=384, A=0, B=1, qinco1_mode=False),
    "odd": dict(D=96, M=2, K=100, L=3, de=192, dh=160, A=0, B=1, qinco1_mode=False),
    "tiny": dict(D=16, M=2, K=64, L=2, de=32, dh=48, A=0, B=1, qinco1_mode=False),
    "noblock": dict(D=32, M=2, K=16, L=0, de=32, dh=32, A=0, B=1, qinco1_mode=False),
}


@pytest.mark.parametrize("name", list(SHAPES))
@pytest.mark.parametrize("opts", [None, [64, 1, 8192, 3, 32], [0, 0, 32768, 0, 0], [0, 2 << 8, 0, 0, 0], [0, 0, 0, 2 << 8, 0]])
def test_op_list_replay_matches_oracle(lib, name, opts):
    cfg = synth.make_cfg(None, **SHAPES[name])
    w = synth.make_weights(cfg, seed=3, n_train=512, kmeans_iters=1, fp16_exact=True)
    if opts and opts[1] >> 8 == 2 and cfg["L"] > 0 and any(w % 32 for w in ([cfg["de"]] if cfg["de"] <= 256 else [cfg["de"] // 2])):
        pytest.skip("CTA-pair mode needs down-projection widths that are multiples of 32")
    plan, ops = export_plan(lib, cfg, opts)
    assert plan["pair"] == (1 if opts and opts[1] >> 8 == 2 else 0)
    # structural invariants of the plan
    assert plan["n_stage"] >= 2 and plan["n_tiles"] in (1, 2)
    assert plan["smem_total"] + 15 * 1024 <= 227 * 1024      # static shared memory of the resident variant: 14 KB
    assert plan["n_tiles"] * plan["tmem_tile_cols"] <= plan["tmem_alloc_cols"] <= 512
    assert plan["tmem_alloc_cols"] & (plan["tmem_alloc_cols"] - 1) == 0
    for op in ops:
        assert op["n"] % 16 == 0 and 16 <= op["n"] <= 256 and op["ks"] % 16 == 0 and op["w_off"] % 16 == 0
    blob = pack(lib, cfg, w, 1, plan, opts)
    T, CB, WxT = tables(lib, cfg, w, 1)
    rng = np.random.default_rng(0)
    n = 64
    codes = rng.integers(0, cfg["K"], n)
    xhat = rng.standard_normal((n, cfg["D"]), dtype=np.float32)
    got = replay(plan, ops, blob, T, CB, WxT, codes, xhat)
    ref = xhat + orc.step_mlp(cfg, w, 1, w["steps.1.codebook.weight"][codes], xhat)
    rel = float(((got - ref) ** 2).sum() / (ref ** 2).sum())
    assert rel <= 1e-6, rel   # only the fp16 rounding of activations separates them (weights are fp16-exact)


def test_planner_replay_on_random_shapes(lib):
    """Property test: for random legal shapes (multiples of 16, with and without projections / skip / pre-selection
    irrelevant here) the planner's op list + the packer's blob replay to the oracle's MLP.  Seeded, ~25 shapes."""
    rng = np.random.default_rng(2024)
    done = 0
    for _ in range(60):
        D = int(rng.choice([16, 32, 48, 64, 96, 128, 192]))
        de = int(rng.choice([D, D, 16 * int(rng.integers(1, 25))]))
        dh = 16 * int(rng.integers(1, 25))
        L = int(rng.integers(0, 4))
        K = int(rng.choice([16, 64, 100, 256]))
        q1 = bool(rng.integers(0, 2))
        opts = [int(rng.choice([0, 64, 128])), int(rng.choice([0, 1])), int(rng.choice([0, 8192, 16384, 32768])), 0,
                int(rng.choice([0, 32, 64]))]
        cfg = synth.make_cfg(None, D=D, M=2, K=K, L=L, de=de, dh=dh, A=0, B=1, qinco1_mode=q1)
        o = (C.c_int32 * 5)(*opts)
        plan_raw = (C.c_int32 * 32)()
        ops = np.zeros(256, OP_DTYPE)
        n_ops = lib.qb_plan_export(D, de, dh, L, K, int(q1), o, plan_raw, 32, ops.ctypes.data_as(C.c_void_p), 256)
        if n_ops < 0:          # shapes the planner refuses (e.g. de too wide for TMEM) must be refused consistently
            assert de + 32 > 512 or n_ops in (-1, -2)
            continue
        plan, ops = dict(zip(PLAN_FIELDS, list(plan_raw))), ops[:n_ops]
        assert plan["smem_total"] + 15 * 1024 <= 227 * 1024 and plan["n_tiles"] * plan["tmem_tile_cols"] <= 512
        w = synth.make_weights(cfg, seed=int(rng.integers(1, 1000)), n_train=max(256, K), kmeans_iters=1, fp16_exact=True)
        blob = pack(lib, cfg, w, 1, plan, opts)
        T, CB, WxT = tables(lib, cfg, w, 1)
        n = 16
        codes = rng.integers(0, K, n)
        xhat = rng.standard_normal((n, D), dtype=np.float32)
        got = replay(plan, ops, blob, T, CB, WxT, codes, xhat)
        ref = xhat + orc.step_mlp(cfg, w, 1, w["steps.1.codebook.weight"][codes], xhat)
        rel = float(((got - ref) ** 2).sum() / (ref ** 2).sum())
        assert rel <= 1e-6, (cfg, opts, rel)
        done += 1
        if done >= 25:
            break
    assert done >= 15


# ---------------------------------------------------------------------------------------------------- decode-loop plan
UOP = 1 << 10       # opts5[3] bit 10: the decode-loop plan (pre-ops computing u = Wx . xhat on the tensor core)


def pack_pre(lib, cfg, w, m, plan, opts, blob):
    o = (C.c_int32 * 5)(*opts)
    D, De = cfg["D"], cfg["de"]
    wx = np.ascontiguousarray(w[f"steps.{m}.concat.mlp.weight"][:, De:], dtype=np.float32)       # [De][D]
    lib.qb_plan_pack_pre.argtypes = [C.c_int] * 6 + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    rc = lib.qb_plan_pack_pre(D, De, cfg["dh"], cfg["L"], cfg["K"], int(cfg["qinco1_mode"]), C.cast(o, C.c_void_p),
                              wx.ctypes.data_as(C.c_void_p), blob.ctypes.data_as(C.c_void_p), len(blob))
    assert rc == 0, rc


def replay_pre(plan, ops, blob, T, codes, xhat):
    """Software model of the decode loop's step start: Eacc = T_m[code], then the pre-ops accumulate the fp16 hi/lo
    products of the operand [xhat_hi | xhat_lo] (k-chunks 0 .. 2D/8 of the A_E buffer) -> e0 = T_m[code] + Wx . xhat."""
    D, De = plan["D"], plan["De"]
    hi = f16(xhat)
    a_x = np.concatenate([hi, f16(xhat - hi)], axis=1)                 # what the previous step's final epilogue writes
    assert a_x.shape[1] // 8 <= plan["ae_chunks"]
    e = T[codes].copy()
    bytes_view = blob.view(np.uint8)
    first = plan["n_ops_block"] + plan["n_ops_out"]
    pre = ops[first:first + plan["n_ops_pre"]]
    assert len(pre) == plan["n_ops_pre"] > 0
    assert pre[0]["wait_a"] == BAR_AE_READY and pre[-1]["commit"] == BAR_EACC_FULL
    for op in pre:
        nn, ks, kt = int(op["n"]), int(op["ks"]), int(op["k_total"])
        assert op["a_src"] == A_E and op["a_off"] == 0 and op["accumulate"] == 1 and kt in (D, 2 * D)
        assert op["slab_bytes"] == nn * ks * 2 <= plan["slot_bytes"]
        c0 = int(op["d_col"])
        for s in range(int(op["n_slab"])):
            kk = min(ks, kt - s * ks)
            off = int(op["w_off"]) + s * int(op["slab_bytes"])
            W = bytes_view[off: off + nn * kk * 2].view(np.float16).astype(np.float32).reshape(kk // 8, nn, 8).transpose(1, 0, 2).reshape(nn, kk)
            e[:, c0:c0 + nn] += a_x[:, s * ks: s * ks + kk] @ W.T
    return e


@pytest.mark.parametrize("name", ["S", "Q1", "L", "deep", "odd", "tiny", "noblock"])
def test_decode_loop_plan_replay(lib, name):
    """The decode-loop plan keeps the plain plan's block / out_proj ops (one shared weight blob) and its pre-ops reproduce
    e0 = T_m[code] + Wcat[:, De:] . xhat to ~fp32 accuracy (fp16 hi/lo split on both operands, three products)."""
    cfg = synth.make_cfg(None, **SHAPES[name])
    w = synth.make_weights(cfg, seed=3, n_train=512, kmeans_iters=1)          # NOT fp16-exact: the lo parts matter
    base, base_ops = export_plan(lib, cfg, None)
    opts = [base["hc"], 1 << 8, base["slot_bytes"], UOP, 0]
    plan, ops = export_plan(lib, cfg, opts)
    assert plan["n_ops_pre"] > 0 and plan["smem_tres"] == -1
    assert plan["ae_chunks"] == max(cfg["de"], 2 * cfg["D"]) // 8
    assert plan["smem_total"] + 15 * 1024 <= 227 * 1024
    nb = len(base_ops)
    assert plan["n_ops_block"] == base["n_ops_block"] and plan["n_ops_out"] == base["n_ops_out"]
    assert ops[:nb].tobytes() == base_ops.tobytes()
    blob = pack(lib, cfg, w, 1, plan, opts)
    np.testing.assert_array_equal(blob[: len(pack(lib, cfg, w, 1, base, None))][: base["w_blob_bytes"] // 2],
                                  pack(lib, cfg, w, 1, base, None)[: base["w_blob_bytes"] // 2])
    pack_pre(lib, cfg, w, 1, plan, opts, blob)
    T, CB, WxT = tables(lib, cfg, w, 1)
    rng = np.random.default_rng(1)
    n = 48
    codes = rng.integers(0, cfg["K"], n)
    xhat = (3.0 * rng.standard_normal((n, cfg["D"]))).astype(np.float32)
    e0 = replay_pre(plan, ops, blob, T, codes, xhat)
    ref = T[codes].astype(np.float64) + xhat.astype(np.float64) @ WxT.astype(np.float64)
    u = xhat.astype(np.float64) @ WxT.astype(np.float64)
    assert np.abs(e0 - ref).max() <= 2e-6 * max(np.abs(u).max(), 1.0)      # a single fp16 product would be ~5e-4


def test_decode_loop_plan_refuses_chunked_out_proj(lib):
    cfg = synth.make_cfg(None, **SHAPES["contriever"])
    o = (C.c_int32 * 5)(0, 1 << 8, 0, UOP, 0)
    plan = (C.c_int32 * 32)()
    ops = np.zeros(256, OP_DTYPE)
    assert lib.qb_plan_export(cfg["D"], cfg["de"], cfg["dh"], cfg["L"], cfg["K"], 0, o, plan, 32, ops.ctypes.data_as(C.c_void_p), 256) < 0


def test_hot_kernels_do_not_spill():
    """ptxas -v of the last build: the tcgen05 kernels keep their spills tiny.  A second inlined copy of one epilogue helper
    once pushed every qb_mlp_kernel variant to ~900 B of spills per thread and cost 25 % on BASELINE config 2 while every
    parity test stayed green -- spills are re-read from L2 on this kernel (208 KB of shared memory leave almost no L1)."""
    log = os.path.join(ROOT, "qinco_b200", "build", "ptxas.log")
    if not os.path.exists(log):
        pytest.skip("no ptxas log (library not built on this machine)")
    text = open(log).read()
    entries = re.findall(r"Compiling entry function '(\S+)'.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", text, re.S)
    hot = [(n, int(st), int(ld)) for n, _, st, ld in entries if "qb_mlp_kernel" in n or "qb_prep_tc" in n or "qb_ivf_tc" in n]
    assert len(hot) >= 10
    worst = max(hot, key=lambda t: t[1])
    assert worst[1] <= 128 and max(t[2] for t in hot) <= 256, worst


def test_library_override_must_exist(monkeypatch):
    """QINCO_B200_LIB (A-B builds of compile-time variants) never falls back silently to the default build."""
    import importlib
    from qinco_b200 import _lib
    monkeypatch.setenv("QINCO_B200_LIB", "/nonexistent/libqinco_b200_variant.so")
    saved = _lib._lib
    _lib._lib = None
    try:
        with pytest.raises(RuntimeError, match="QINCO_B200_LIB"):
            _lib.load()
    finally:
        _lib._lib = saved


@pytest.mark.parametrize("shape,view", [
    (dict(D=128, de=128, dh=256, L=2), 1),                          # BASELINE config 2 (QINCo2-S)
    (dict(D=128, de=128, dh=256, L=16, qinco1_mode=True), 1),       # config 1 (QINCo1)
    (dict(D=128, de=384, dh=384, L=16), 2),                         # config 3 (QINCo2-L)
    (dict(D=96, de=384, dh=384, L=16), 2),                          # config 4 (Deep1B shape)
    (dict(D=768, de=384, dh=384, L=16), 2),                         # config 5 (Contriever shape)
    (dict(D=64, de=64, dh=128, L=2), 0),                            # anything else: the generic kernels
])
def test_baseline_shapes_get_their_compile_time_plan_view(lib, shape, view):
    """The fixed-shape kernels are only chosen when the planner's output matches their constants field by field; if the
    planner changes, this test (not a silent 6-9 % slowdown) says so."""
    cfg = synth.make_cfg(None, M=8, K=256, A=16, B=16, **shape)
    plan, _ = export_plan(lib, cfg)
    assert plan["plan_view"] == view, plan
    if view == 2 and shape["D"] <= 128:       # the decode-loop plan of the L family uses the same view (its pre-ops are run-time
        # fields); d = 768 has no decode-loop plan (out_proj in six chunks)
        loop, _ = export_plan(lib, cfg, [plan["hc"], (1 << 8) | (2 << 16), plan["slot_bytes"], 1 << 10, 0])
        assert loop["plan_view"] == 2 and loop["n_ops_pre"] > 0
