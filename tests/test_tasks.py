"""The eval / encode drivers (qinco_b200/tasks.py): host logic on CPU with the oracle behind the
reference's call surface (gloo, world size 2), and on the GPU with the CUDA model (1 GPU; NCCL when 2 are visible)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import qinco_oracle as orc
from qinco_b200 import io, shard, synth, tasks


class OracleModel:
    """model(x, step=...) of the reference on top of the CPU oracle."""

    def __init__(self, cfg, w):
        self.cfg, self.w = cfg, w
        self.M, self.K, self.D = cfg["M"], cfg["K"], cfg["D"]
        self.ivf_K = cfg.get("ivf_K", 0)
        self.M_ivf = self.M + (1 if self.ivf_K else 0)

    def __call__(self, x, step="encode"):
        return torch.from_numpy(np.asarray(orc.forward(self.cfg, self.w, x.numpy(), step)))


def _small(ivf=False):
    cfg = synth.make_cfg(None, D=16, M=3, K=32, L=1, de=16, dh=16, A=4, B=2, **({"ivf_K": 50} if ivf else {}))
    return cfg, synth.make_weights(cfg, seed=5, n_train=512, kmeans_iters=1, data_mean=0.2, data_std=1.5)


def test_compute_mse_reports_like_the_reference():
    cfg, w = _small()
    x = synth.make_data(75, 16, seed=3, mean=0.2, std=1.5)
    lines = []
    res = tasks.compute_MSE(OracleModel(cfg, w), x, batch=32, device="cpu", timed=True, out=lines.append)
    ref = orc.forward(cfg, w, orc.forward(cfg, w, x, "encode"), "decode")
    assert res["n_vecs"] == 75 and abs(res["MSE"] - orc.mse(x, ref)) <= 1e-6 * res["MSE"]
    text = "\n".join(lines)
    assert "Test metrics: [[MSE=" in text and "test_codeword_entropy=" in text
    assert "Encoding time / vector:" in text and "Decoding time / vector:" in text and "μs" in text
    assert 0 < res["entropy"] <= np.log2(cfg["K"])


def test_reference_shard_range_matches_encode_database():
    for n in (0, 1, 7, 64, 1001):
        for world in (1, 2, 3, 8):
            rows = []
            for r in range(world):
                s, e = shard.reference_shard_range(n, r, world)
                assert (s, e) == ((n // world) * r, (n // world) * (r + 1) if r < world - 1 else n)   # search_tasks.py:103-104
                rows += list(range(s, e))
            assert rows == list(range(n))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _encode_worker(rank, world, port, n, out, ivf):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg, w = _small(ivf)
    x = synth.make_data(n, 16, seed=9, mean=0.2, std=1.5)
    tasks.encode_database(OracleModel(cfg, w), x, out, batch=5, device="cpu", out=lambda *a, **k: None)
    dist.destroy_process_group()


@pytest.mark.parametrize("ivf", [False, True])
def test_encode_database_part_files(tmp_path, ivf):
    """world size 2 over gloo: `<out>.npz` + `<out>.part_r.npz` as the reference writes them, readable back to the codes
    of a single-process run (search_tasks.py:122-131, search_utils.py:33-78)."""
    world, n = 2, 23
    out = str(tmp_path / "db.npz")
    mp.spawn(_encode_worker, args=(world, _free_port(), n, out, ivf), nprocs=world, join=True)
    cfg, w = _small(ivf)
    ref = orc.forward(cfg, w, synth.make_data(n, 16, seed=9, mean=0.2, std=1.5), "encode").T
    codes, meta = io.load_encoded_db(out)
    assert meta == dict(n_parts=2, K=32, M=3, D=16)
    assert codes.dtype == np.int64 and codes.shape == ref.shape
    np.testing.assert_array_equal(codes, ref)
    p0 = np.load(out[:-4] + ".part_0.npz")["codes"]
    assert len(p0) == n // 2                                   # floor split, the last rank takes the remainder


def test_shard_fallback_refuses_to_truncate():
    """A model with the reference's surface whose cfg says ivf_in_use (and no ivf_K attribute) is handled as IVF; codes that
    do not fit a byte raise instead of wrapping."""
    cfg, w = _small(True)

    class RefLike:                                       # like reference QINCo: M == cfg._M_ivf, IVF only visible in cfg
        def __init__(self):
            self.cfg = type("Cfg", (), {"ivf_in_use": True, "ivf_K": 50})()
            self.M = cfg["M"] + 1

        def __call__(self, x, step="encode"):
            return torch.from_numpy(orc.forward(cfg, w, x.numpy(), "encode"))

    x = torch.from_numpy(synth.make_data(9, 16, seed=1, mean=0.2, std=1.5))
    ivf, codes = shard.encode_sharded(RefLike(), x, 9, gather=False)
    ref = orc.forward(cfg, w, x.numpy(), "encode")
    np.testing.assert_array_equal(ivf.numpy(), ref[0])
    np.testing.assert_array_equal(codes.numpy(), ref[1:].T)

    class Wide:
        M = 2

        def __call__(self, x, step="encode"):
            return torch.full((2, len(x)), 300, dtype=torch.int64)

    with pytest.raises(ValueError):
        shard.encode_sharded(Wide(), x, 9, gather=False)


# ------------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_eval_task_on_gpu(capsys):
    """task=eval_time through the CUDA model on the QINCo2-S synthetic workload: same MSE as decoding our codes with the
    oracle, report lines present."""
    import bench
    from qinco_b200.model import QINCo
    cfg, w, x = bench.make_model_inputs(bench.WORKLOADS["c2"], 3000, 0)
    model = QINCo(cfg, w, device="cuda:0")
    try:
        res = tasks.compute_MSE(model, x.numpy(), batch=1024, device=model.device, timed=True)
        codes = model(x.cuda(), step="encode").cpu().numpy()
        model.synchronize()
        assert abs(res["MSE"] - orc.mse(x.numpy(), orc.decode(cfg, w, codes))) <= 1e-5 * res["MSE"]
        assert "Encoding time / vector" in capsys.readouterr().out
    finally:
        model._h.close()


@pytest.mark.gpu
def test_encode_task_single_gpu(tmp_path):
    import bench
    from qinco_b200.model import QINCo
    cfg, w, x = bench.make_model_inputs(bench.WORKLOADS["c2a16"], 2500, 0)
    model = QINCo(cfg, w, device="cuda:0")
    try:
        out = str(tmp_path / "db.npz")
        codes = tasks.encode_database(model, x.numpy(), out, batch=1024, device=model.device)
        full = model(x.cuda(), step="encode").T.cpu().numpy()
        model.synchronize()
        np.testing.assert_array_equal(codes, full)
        loaded, meta = io.load_encoded_db(out)
        np.testing.assert_array_equal(loaded, full)
        assert meta == dict(n_parts=1, K=256, M=8, D=128)
    finally:
        model._h.close()


def _nccl_worker(rank, world, port, n, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import bench
    from qinco_b200.model import QINCo
    cfg, w, x = bench.make_model_inputs(bench.WORKLOADS["c2a16"], n, 0)
    model = QINCo(cfg, w, device=dev)
    s, e = shard.shard_range(n, rank, world)
    codes = shard.encode_sharded(model, x[s:e].to(dev), n, batch=777)
    torch.cuda.synchronize(dev)
    model.synchronize()
    if rank == 0:
        full = model(x.to(dev), step="encode").T.to(torch.uint8)
        assert torch.equal(codes, full)
    np.save(os.path.join(out_dir, f"codes_{rank}.npy"), codes.cpu().numpy())
    # the current device of the process is untouched by model creation on another ordinal (DeviceGuard)
    assert torch.cuda.current_device() == rank
    dist.barrier()
    model._h.close()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_encode_sharded_two_gpus_nccl(tmp_path):
    """shard.encode_sharded with the CUDA model over NCCL (skipped when fewer than 2 GPUs are visible)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    n = 3001
    mp.spawn(_nccl_worker, args=(2, _free_port(), n, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "codes_0.npy"), np.load(tmp_path / "codes_1.npy")
    assert a.shape == (n, 8) and np.array_equal(a, b)
