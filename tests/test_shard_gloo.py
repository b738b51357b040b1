"""N > 1 path on CPU: row sharding + the single all-gather of uint8 codes, world size 2 and 3 over gloo.

The encoder here is the numpy oracle wrapped in the reference's call surface (the CUDA model needs a GPU); what is
under test is the host logic of qinco_b200/shard.py, the counterpart of the reference's `encode_database`
(qinco/search/search_tasks.py:85-137).
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import qinco_oracle as orc
from qinco_b200 import shard, synth


class OracleModel:
    """Reference call surface (model(x, step="encode") -> LongTensor [M, n]) on top of the CPU oracle."""

    def __init__(self, cfg, w):
        self.cfg, self.w, self.M, self.ivf_K = cfg, w, cfg["M"], cfg.get("ivf_K", 0)

    def __call__(self, x, step="encode"):
        assert step == "encode"
        return torch.from_numpy(orc.forward(self.cfg, self.w, x.numpy(), "encode"))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _cfg(ivf):
    return synth.make_cfg(None, D=16, M=3, K=32, L=1, de=16, dh=16, A=4, B=2, **({"ivf_K": 300} if ivf else {}))


def _worker(rank, world, port, n, out_dir, ivf=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = _cfg(ivf)
    w = synth.make_weights(cfg, seed=5, n_train=512, kmeans_iters=1)
    x = torch.from_numpy(synth.make_data(n, 16, seed=9))
    s, e = shard.shard_range(n, rank, world)
    codes = shard.encode_sharded(OracleModel(cfg, w), x[s:e], n, batch=7)
    if ivf:
        np.save(os.path.join(out_dir, f"ivf_{rank}.npy"), codes[0].numpy())
        codes = codes[1]
    np.save(os.path.join(out_dir, f"codes_{rank}.npy"), codes.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_covers_rows_once():
    for n in (0, 1, 7, 64, 1000):
        for world in (1, 2, 3, 8):
            rows = []
            for r in range(world):
                s, e = shard.shard_range(n, r, world)
                assert 0 <= s <= e <= n
                rows += list(range(s, e))
            assert rows == list(range(n))


@pytest.mark.parametrize("world,n", [(2, 37), (3, 10), (2, 1)])
def test_encode_sharded_matches_single_process(tmp_path, world, n):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    cfg = synth.make_cfg(None, D=16, M=3, K=32, L=1, de=16, dh=16, A=4, B=2)
    w = synth.make_weights(cfg, seed=5, n_train=512, kmeans_iters=1)
    ref = orc.forward(cfg, w, synth.make_data(n, 16, seed=9), "encode").T.astype(np.uint8)
    for r in range(world):
        got = np.load(tmp_path / f"codes_{r}.npy")
        assert got.dtype == np.uint8 and got.shape == (n, cfg["M"])
        np.testing.assert_array_equal(got, ref)      # every rank ends with the full, identical code matrix


def test_encode_sharded_ivf_model(tmp_path):
    """IVF-QINCo: the int32 IVF codes travel as four extra byte columns of the one all-gather."""
    world, n = 2, 23
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, str(tmp_path), True), nprocs=world, join=True)
    cfg = _cfg(True)
    w = synth.make_weights(cfg, seed=5, n_train=512, kmeans_iters=1)
    ref = orc.forward(cfg, w, synth.make_data(n, 16, seed=9), "encode")
    for r in range(world):
        ivf, codes = np.load(tmp_path / f"ivf_{r}.npy"), np.load(tmp_path / f"codes_{r}.npy")
        assert ivf.dtype == np.int32 and ivf.shape == (n,) and codes.shape == (n, cfg["M"])
        np.testing.assert_array_equal(ivf, ref[0])
        np.testing.assert_array_equal(codes, ref[1:].T.astype(np.uint8))
