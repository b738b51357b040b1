"""Host-side formats: checkpoint dicts, encoded-database files, raw bit strings, vecs files."""
import os

import numpy as np
import pytest
import torch

from qinco_b200 import io, synth


def test_bitstrings_known_answer_and_round_trip():
    # faiss.pack_bitstrings semantics: value j at bits [j*nbits, (j+1)*nbits), least-significant bit first
    codes = np.array([[1, 2, 3], [7, 0, 5]])
    p = io.pack_bitstrings(codes, 3)
    assert p.shape == (2, 2) and p.dtype == np.uint8
    assert p.tolist() == [[1 | (2 << 3) | ((3 & 3) << 6), 3 >> 2], [7 | (0 << 3) | ((5 & 3) << 6), 5 >> 2]]
    rng = np.random.default_rng(0)
    for M, K in ((8, 256), (16, 256), (5, 100), (3, 2), (7, 1 << 12)):
        nbits = int(np.ceil(np.log2(K)))
        c = rng.integers(0, K, (37, M))
        q = io.pack_bitstrings(c, nbits)
        assert q.shape == (37, io.code_size_bytes(M, K))
        assert np.array_equal(io.unpack_bitstrings(q, nbits, M), c)
    with pytest.raises(AssertionError):
        io.pack_bitstrings(np.array([[8]]), 3)


def test_raw_and_encoded_db_files(tmp_path):
    rng = np.random.default_rng(1)
    codes = rng.integers(0, 256, (50, 8))
    raw = str(tmp_path / "codes.raw")
    io.write_raw_codes(raw, codes, 256)
    assert os.path.getsize(raw) == 50 * 8
    assert np.array_equal(io.read_raw_codes(raw, 8, 256), codes)
    out = str(tmp_path / "db.npz")
    io.save_encoded_db(out, [codes[:20], codes[20:]], K=256, M=8, D=128)
    back, meta = io.load_encoded_db(out)
    assert meta == dict(n_parts=2, K=256, M=8, D=128) and np.array_equal(back, codes) and back.dtype == np.int64
    z = np.load(str(tmp_path / "db.part_1.npz"))
    assert list(z.keys()) == ["codes"] and z["codes"].shape == (30, 8)


def test_vecs_readers(tmp_path):
    x = np.arange(24, dtype=np.float32).reshape(4, 6)
    f = str(tmp_path / "x.fvecs")
    np.concatenate([np.full((4, 1), 6, np.int32).view(np.float32), x], axis=1).tofile(f)
    assert np.array_equal(io.read_vectors(f), x)
    b = str(tmp_path / "x.bvecs")
    xb = (np.arange(24) % 251).astype(np.uint8).reshape(4, 6)
    np.concatenate([np.tile(np.array([6], np.int32).view(np.uint8), (4, 1)), xb], axis=1).tofile(b)
    assert np.array_equal(io.read_vectors(b), xb.astype(np.float32))
    n = str(tmp_path / "x.npy")
    np.save(n, x.astype(np.float64))
    assert io.read_vectors(n).dtype == np.float32


def test_v2_checkpoint_dict(tmp_path):
    """A file in the layout of save_model (qinco/utils.py:118-136) with legacy keys, read back like load_saved_model_data."""
    cfg = synth.make_cfg(None, D=32, M=3, K=16, L=2, de=48, dh=64, A=4, B=2)
    w = synth.make_weights(cfg, seed=1, n_train=256, kmeans_iters=1)
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in w.items()}
    sd["steps.0.substep.codebook.weight"] = torch.zeros(16, 32)                       # dropped by the loader
    sd["steps.1.residual_blocks.0.in_proj.weight"] = sd.pop("steps.1.in_proj.weight")    # legacy location
    sd["steps.1.xtarget_mean"] = torch.zeros(32)                                      # training-only buffer, ignored later
    path = str(tmp_path / "ckpt.pt")
    torch.save({"epoch": 3, "model": sd, "optimizer": None, "scheduler": None, "logger": None,
                "parameters": {"K": 16, "M": 3, "de": 48, "dh": 64, "L": 2, "A": 4, "B": 2, "ivf_in_use": False,
                               "qinco1_mode": False}, "data_dim": 32}, path)
    got_cfg, got_sd = io.load_v2_checkpoint(path, dict(B=8, A=None))
    assert got_cfg == dict(D=32, M=3, K=16, L=2, de=48, dh=64, A=4, B=8, qinco1_mode=False)
    assert "steps.0.substep.codebook.weight" not in got_sd and "steps.1.in_proj.weight" in got_sd
    assert np.array_equal(got_sd["steps.2.codebook.weight"], np.asarray(w["steps.2.codebook.weight"]))
    with pytest.raises(ValueError):      # A > 0 requested for a model trained with A = 0 (qinco/utils.py:166-169)
        io.cfg_from_v2_checkpoint({"model": sd, "parameters": {"A": 0, "M": 3, "K": 16, "L": 2}, "data_dim": 32}, dict(A=8))
    ivf_cfg = io.cfg_from_v2_checkpoint({"model": sd, "parameters": {"ivf_in_use": True, "ivf_K": 1024, "L": 2}, "data_dim": 32})
    assert ivf_cfg["ivf_K"] == 1024 and ivf_cfg["M"] == 2          # steps.0 is the IVF step: M = _M_ivf - 1
    # parameters missing: shapes are inferred from the tensors
    assert io.cfg_from_v2_checkpoint({"model": io.clean_v2_state_dict(sd)})["dh"] == 64


def test_v1_checkpoint_state_dict(tmp_path):
    cfg = synth.make_cfg(None, D=16, M=3, K=8, L=2, de=16, dh=24, A=0, B=1, qinco1_mode=True)
    w = synth.make_weights(cfg, seed=2, n_train=128, kmeans_iters=1)
    v1 = {k: torch.from_numpy(np.asarray(v)) for k, v in synth.to_v1_state(cfg, w).items()}
    path = str(tmp_path / "v1.pt")
    torch.save({"state_dict": v1, "db_scale": 2.5}, path)
    sd, scale = io.load_v1_checkpoint(path)
    assert scale == 2.5
    cfg2, w2 = synth.from_v1_state(sd)
    assert cfg2["M"] == 3 and cfg2["L"] == 2 and cfg2["dh"] == 24
    assert np.array_equal(w2["steps.1.concat.mlp.weight"], np.asarray(w["steps.1.concat.mlp.weight"]))


@pytest.mark.gpu
def test_cli_round_trip(tmp_path):
    """--encode / --decode through the CLI, v1 (npy + raw) and v2 (encoded database) formats."""
    from qinco_b200 import cli
    cfg = synth.make_cfg(None, D=32, M=4, K=64, L=1, de=32, dh=32, A=0, B=1, qinco1_mode=True)
    w = synth.make_weights(cfg, seed=5, n_train=1024, kmeans_iters=1)
    x = synth.make_data(300, 32, seed=9)
    xin = str(tmp_path / "x.npy")
    np.save(xin, x)
    m1 = str(tmp_path / "v1.pt")
    torch.save({"state_dict": {k: torch.from_numpy(np.asarray(v)) for k, v in synth.to_v1_state(cfg, w).items()}, "db_scale": 1.0}, m1)
    cli.main(["--encode", "--model", m1, "--i", xin, "--o", str(tmp_path / "c.npy")])
    cli.main(["--encode", "--model", m1, "--i", xin, "--o", str(tmp_path / "c.raw"), "--raw"])
    c = np.load(str(tmp_path / "c.npy"))
    assert c.shape == (300, 4) and np.array_equal(io.read_raw_codes(str(tmp_path / "c.raw"), 4, 64), c)
    cli.main(["--decode", "--model", m1, "--i", str(tmp_path / "c.raw"), "--o", str(tmp_path / "y.npy"), "--raw"])
    y = np.load(str(tmp_path / "y.npy"))
    assert y.shape == (300, 32) and ((y - x) ** 2).sum() < (x ** 2).sum()
    cfg2 = synth.make_cfg(None, D=32, M=4, K=64, L=1, de=32, dh=32, A=8, B=4)
    w2 = synth.make_weights(cfg2, seed=6, n_train=1024, kmeans_iters=1)
    m2 = str(tmp_path / "v2.pt")
    torch.save({"model": {k: torch.from_numpy(np.asarray(v)) for k, v in w2.items()},
                "parameters": {k: cfg2[k] for k in ("K", "M", "de", "dh", "L", "A", "B", "qinco1_mode")}, "data_dim": 32}, m2)
    cli.main(["--encode", "--v2", "--model", m2, "--i", xin, "--o", str(tmp_path / "db.npz")])
    codes, meta = io.load_encoded_db(str(tmp_path / "db.npz"))
    assert codes.shape == (300, 4) and meta["K"] == 64 and meta["D"] == 32
    cli.main(["--decode", "--v2", "--model", m2, "--i", str(tmp_path / "db.npz"), "--o", str(tmp_path / "y2.npy")])
    y2 = np.load(str(tmp_path / "y2.npy"))
    assert y2.shape == (300, 32) and ((y2 - x) ** 2).sum() < (x ** 2).sum()


@pytest.mark.gpu
def test_cli_ivf_round_trip(tmp_path):
    """IVF-QINCo through the CLI: centroids from a separate .npy (cfg.ivf_centroids), codes [n, M + 1] with the IVF code in
    column 0 (like model(batch, step="encode").T), host-buffer entry points qb_encode_ivf_host / qb_decode_ivf_host."""
    from oracle import qinco_oracle as orc
    from qinco_b200 import cli
    cfg = synth.make_cfg(None, D=32, M=3, K=64, L=1, de=32, dh=32, A=8, B=4, ivf_K=120)
    w = synth.make_weights(cfg, seed=7, n_train=1024, kmeans_iters=1)
    x = synth.make_data(257, 32, seed=4)
    xin, cent, ck = str(tmp_path / "x.npy"), str(tmp_path / "cent.npy"), str(tmp_path / "ivf.pt")
    np.save(xin, x)
    np.save(cent, w["steps.0.ivf_centroids.weight"])
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in w.items() if k != "steps.0.ivf_centroids.weight"}
    torch.save({"model": sd, "parameters": {"K": 64, "M": 3, "de": 32, "dh": 32, "L": 1, "A": 8, "B": 4, "ivf_in_use": True,
                                             "ivf_K": 120, "qinco1_mode": False}, "data_dim": 32}, ck)
    cli.main(["--encode", "--v2", "--model", ck, "--ivf_centroids", cent, "--i", xin, "--o", str(tmp_path / "db.npz")])
    codes, meta = io.load_encoded_db(str(tmp_path / "db.npz"))
    assert codes.shape == (257, 4) and codes[:, 0].max() < 120 and codes[:, 1:].max() < 64
    ref_codes, _ = orc.encode(cfg, w, x)
    assert (codes.T == ref_codes).all(0).mean() >= 0.8
    cli.main(["--decode", "--v2", "--model", ck, "--ivf_centroids", cent, "--i", str(tmp_path / "db.npz"), "--o", str(tmp_path / "y.npy")])
    y = np.load(str(tmp_path / "y.npy"))
    ref = orc.decode(cfg, w, codes.T)
    assert ((y - ref) ** 2).sum() / (ref ** 2).sum() <= 1e-4


def test_cfg_normalisation_accepts_reference_style_objects():
    """model.normalize_cfg: plain dicts and objects shaped like the reference's SharedCfgState (qinco/utils.py:16-40),
    with and without an IVF first step (cfg.ivf_in_use, cfg._M_ivf = M + 1; qinco/qinco_tasks.py:378-383)."""
    from types import SimpleNamespace
    from qinco_b200.model import normalize_cfg
    ref = SimpleNamespace(_D=96, M=16, _M_ivf=16, K=256, L=16, de=384, dh=384, A=16, B=32, qinco1_mode=False,
                          ivf_in_use=False, ivf_K=1048576, _ivf_book=None)
    assert normalize_cfg(ref) == dict(D=96, M=16, K=256, L=16, de=384, dh=384, A=16, B=32, qinco1_mode=False)
    ref.ivf_in_use, ref._M_ivf, ref._ivf_book = True, 17, object()
    got = normalize_cfg(ref)
    assert got["M"] == 16 and got["ivf_K"] == 1048576
    d = normalize_cfg(dict(D=128, M=8, K=256, L=2, de=None, dh=256, A=0, B=1, qinco1_mode=True))
    assert d["de"] == 128 and "ivf_K" not in d
    assert normalize_cfg(dict(D=128, M=8, K=256, L=2, de=128, dh=256, ivf_K=4096))["ivf_K"] == 4096
    with pytest.raises(ValueError):
        normalize_cfg(dict(M=8, K=256, L=2, dh=256))


def test_v1_pickled_module_goes_through_the_restricted_unpickler(tmp_path):
    """A pickled v1 nn.Module (what reference qinco_v1/codec_qinco.py:112 loads) is refused by default, read on request
    without the reference on sys.path, and a pickle that names anything else is rejected."""
    import pickle
    import sys
    import types
    cfg = synth.make_cfg(None, D=16, M=3, K=32, L=2, de=16, dh=32, A=0, B=1, qinco1_mode=True)
    w = synth.make_weights(cfg, seed=3, n_train=512, kmeans_iters=1)
    v1 = synth.to_v1_state(cfg, w)
    fake = types.ModuleType("model_qinco")                 # stands in for the reference's module while pickling

    class QINCoStep(torch.nn.Module):
        pass

    class QINCo(torch.nn.Module):
        pass

    for cls in (QINCoStep, QINCo):
        cls.__module__, cls.__qualname__ = "model_qinco", cls.__name__
        setattr(fake, cls.__name__, cls)
    sys.modules["model_qinco"] = fake
    try:
        model = QINCo()
        model.codebook0 = torch.nn.Embedding(32, 16)
        for m in range(1, 3):
            step = QINCoStep()
            step.codebook = torch.nn.Embedding(32, 16)
            step.MLPconcat = torch.nn.Linear(32, 16)
            for l in range(2):
                step.add_module(f"residual_block{l}", torch.nn.Sequential(torch.nn.Linear(16, 32, bias=False), torch.nn.ReLU(),
                                                                          torch.nn.Linear(32, 16, bias=False)))
            model.add_module(f"step{m}", step)
        model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in v1.items()})
        model.db_scale = 3.25
        path = str(tmp_path / "v1_module.pt")
        torch.save(model, path)
    finally:
        del sys.modules["model_qinco"]
    with pytest.raises(RuntimeError, match="allow_pickled_module"):
        io.load_v1_checkpoint(path)
    sd, scale = io.load_v1_checkpoint(path, allow_pickled_module=True)
    assert scale == 3.25 and set(sd) == set(v1)
    for k in v1:
        np.testing.assert_array_equal(sd[k], v1[k])

    class Evil:
        def __reduce__(self):
            import os
            return (os.system, ("echo pwned > /dev/null",))

    bad = str(tmp_path / "evil.pt")
    torch.save({"state_dict": Evil()}, bad)
    with pytest.raises(pickle.UnpicklingError):
        io.load_v1_checkpoint(bad, allow_pickled_module=True)
