#!/usr/bin/env python
"""Benchmark of the QINCo2 encode hot path on B200 (contract: see the task brief / DESIGN.md section "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c2a16|c3|c3a0|c4|c5|q1|livf|c5pw] [--n VECTORS_PER_GPU]
    python bench.py --impl reference ...      # the reference's PyTorch-CPU path (oracle/torch_port.py) on the host cores
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one encode pass over one batch of synthetic vectors (n per GPU, resident in HBM for `value`; pinned host
buffers through the C-ABI host call for `e2e`).  Rows are sharded by rank (weak scaling: n per GPU is fixed) and the
final uint8 codes are all-gathered once per step with NCCL, inside the timed region.  One JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "vectors/sec encoded (8x8 RQ, d=128)"
UNIT = "vectors/s"

# BASELINE.json configs; model hyper-parameters are the reference presets
# (reference config/model_args/qinco2-S.yaml, qinco2-L.yaml, qinco1.yaml)
WORKLOADS = {
    "c2": dict(name="QINCo2-S 8x8 K=256 d=128 A=0 beam=1", cfg=dict(D=128, M=8, K=256, L=2, de=128, dh=256, A=0, B=1, qinco1_mode=False), n=1_000_000),
    "c2a16": dict(name="QINCo2-S 8x8 K=256 d=128 A=16 beam=1", cfg=dict(D=128, M=8, K=256, L=2, de=128, dh=256, A=16, B=1, qinco1_mode=False), n=1_000_000),
    "c3": dict(name="QINCo2-L 8x8 K=256 d=128 A=16 beam=16", cfg=dict(D=128, M=8, K=256, L=16, de=384, dh=384, A=16, B=16, qinco1_mode=False), n=100_000),
    "c3a0": dict(name="QINCo2-L 8x8 K=256 d=128 A=0 beam=16", cfg=dict(D=128, M=8, K=256, L=16, de=384, dh=384, A=0, B=16, qinco1_mode=False), n=8_192),
    "c4": dict(name="QINCo2-L 16x8 K=256 d=96 (Deep1B shape) A=16 beam=16", cfg=dict(D=96, M=16, K=256, L=16, de=384, dh=384, A=16, B=16, qinco1_mode=False), n=50_000),
    "c5": dict(name="QINCo2-L 8x8 K=256 d=768 (Contriever shape) A=16 beam=32", cfg=dict(D=768, M=8, K=256, L=16, de=384, dh=384, A=16, B=32, qinco1_mode=False), n=50_000),
    "livf": dict(name="IVF-QINCo2-L 8x8 K=256 d=128 A=16 beam=16, IVF 65536 centroids", cfg=dict(D=128, M=8, K=256, L=16, de=384, dh=384, A=16, B=16, qinco1_mode=False, ivf_K=65536), n=50_000),
    "ivf1m": dict(name="IVF-QINCo2-S 8x8 K=256 d=128 A=16 beam=1, IVF 1048576 centroids (billion-scale index shape)", cfg=dict(D=128, M=8, K=256, L=2, de=128, dh=256, A=16, B=1, qinco1_mode=False, ivf_K=1 << 20), n=189_440),      # 10 waves of 148 x 128 vectors
    "q1": dict(name="QINCo1 8x8 K=256 d=128 L=16 beam=1", cfg=dict(D=128, M=8, K=256, L=16, de=128, dh=256, A=0, B=1, qinco1_mode=True), n=200_000),
}


def flops_min_per_candidate(cfg):
    """hoisted minimum MACs per candidate row x 2."""
    D, De, Dh, L = cfg["D"], cfg["de"], cfg["dh"], cfg["L"]
    p = D * De if De != D else 0
    return 2 * (2 * L * De * Dh + p + D)


def encode_flops_min_per_vector(cfg):
    D, De, M, K, A, B = cfg["D"], cfg["de"], cfg["M"], cfg["K"], cfg["A"], cfg["B"]
    C = A or K
    per_cand = flops_min_per_candidate(cfg) // 2
    ivf_K = cfg.get("ivf_K") or 0
    mac = (ivf_K or K) * D                      # step 0: plain codebook, or the IVF arg-min over ivf_K centroids
    for m in range(1, M + (1 if ivf_K else 0)):
        f_in = 1 if (ivf_K and m == 1) else B    # the step after an IVF step starts from one beam ...
        c = (max(A, B) if A else K) if (ivf_K and m == 1) else C      # ... and pre-selects max(A, B) candidates
        mac += (f_in * K * D if A > 0 else 0) + f_in * c * per_cand + f_in * D * De
    return 2 * mac


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(tflops=float(j.get("bf16_tflops_sustained") or j["bf16_tflops"]), hbm=float(j["hbm_gbs"]),
                    source="MEASURED_PEAKS.json (bf16 dense, sustained)")
    return dict(tflops=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [t.strip() for t in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                smax.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), reasons=sorted(reasons), samples=len(sm))
        return out


def make_model_inputs(wl, n, rank):
    import torch
    from qinco_b200 import synth
    cfg = synth.make_cfg(None, **wl["cfg"])
    w = synth.make_weights(cfg, seed=4321, gain=0.5, n_train=8192 if (cfg.get("ivf_K") or 0) <= 65536 else 2048,
                           kmeans_iters=1 if cfg.get("ivf_K") else 3)
    g = torch.Generator().manual_seed(1234 + rank)
    x = torch.randn(n, cfg["D"], generator=g, dtype=torch.float32)
    return cfg, w, x


def contract_sample(cfg):
    """SURVEY.md section 8(d) "Parity subsets": the first 10 000 rows for the small models (S / QINCo1), 1 024 for QINCo2-L
    (and for models with a 2^20-centroid IVF step, whose CPU arg-min alone is 268 MFLOP per vector)."""
    return 1024 if (cfg["de"] >= 384 or (cfg.get("ivf_K") or 0) > 65536) else 10000


def cpu_port_rate(cfg, w, x_np):
    """Reference PyTorch-CPU encode + decode (oracle/torch_port.py) on the contract's parity sample:
    (encode vec/s, sample size, codes, xhat, threads, port, decode vec/s)."""
    import torch
    from oracle.torch_port import TorchPort
    threads = os.cpu_count() or 1
    port = TorchPort(cfg, w, threads=threads)
    n = min(len(x_np), contract_sample(cfg))
    port.encode(x_np[:min(n, 64)])                       # thread-pool / allocator warm-up
    t0 = time.perf_counter()
    codes, xhat = port.encode(x_np[:n])
    dt = time.perf_counter() - t0
    t0 = time.perf_counter()
    port.decode(codes)
    dt_dec = time.perf_counter() - t0
    return n / dt, n, codes.numpy(), xhat.numpy(), torch.get_num_threads(), port, n / dt_dec


def parity_block(cfg, model, port, xs, ours, ref_codes, ref_xhat, dev):
    """decode MSE vs the reference on ITS codes; encode MSE of OUR codes decoded with the reference arithmetic."""
    import torch
    dec_ours = model.decode(torch.from_numpy(ref_codes).to(dev)).cpu().numpy()
    dec_ref = port.decode(ref_codes).numpy()
    d_ref = ((xs - ref_xhat) ** 2).sum(1).astype(np.float64)
    d_ours = ((xs - port.decode(ours).numpy()) ** 2).sum(1).astype(np.float64)
    same = (ours == ref_codes).all(0)
    delta = (d_ours - d_ref)[~same]
    return {"sample": int(len(xs)),
            "decode_rel_mse_vs_ref": float(((dec_ours - dec_ref) ** 2).sum() / (dec_ref ** 2).sum()),
            "encode_mse_ref": float(d_ref.mean()), "encode_mse_ours": float(d_ours.mean()),
            "encode_mse_rel_diff": float(abs(d_ours.mean() - d_ref.mean()) / d_ref.mean()),
            "vectors_with_identical_codes": float(same.mean()),
            "differing_vectors_mean_delta_over_mse": float(delta.mean() / d_ref.mean()) if len(delta) else 0.0,
            "differing_vectors_sem_over_mse": float(delta.std(ddof=1) / np.sqrt(len(delta)) / d_ref.mean()) if len(delta) > 1 else 0.0}


def decode_flops_per_vector(cfg):
    """hoisted form, per vector: (M - 1) x (the MLP of one candidate + u = Wx . xhat)"""
    D_, De_, Dh_, L_, M_ = cfg["D"], cfg["de"], cfg["dh"], cfg["L"], cfg["M"]
    return 2 * (M_ - 1 + (1 if cfg.get("ivf_K") else 0)) * (2 * L_ * De_ * Dh_ + (D_ * De_ if De_ != D_ else 0) + D_ * De_)


def timed_encode(model, x_dev, steps, warmup, local, sample_clocks=True):
    """`warmup` untimed + `steps` timed encode passes over x_dev (resident): (ms total, kinds, launches, clocks, codes)."""
    import torch
    h = model._h
    for _ in range(warmup):
        model.encode_u8(x_dev, normalize=True, want_xhat=False)
    torch.cuda.synchronize()
    h.timing_read()
    h.timing_enable(True)
    l0 = h.launch_count
    sampler = ClockSampler(local) if sample_clocks else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        codes, _ = model.encode_u8(x_dev, normalize=True, want_xhat=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    kinds = h.timing_read()
    h.timing_enable(False)
    return ms, kinds, h.launch_count - l0, clocks, codes


def score_roofline(cfg, kinds, peaks):
    ms_score, n_score, rows_score = kinds["mlp_score"]
    fl = flops_min_per_candidate(cfg)
    achieved = rows_score * fl / (ms_score * 1e-3) / 1e12 if ms_score > 0 else 0.0
    return {"bound": "tensor", "kernel": "qb_mlp_kernel (score)", "achieved": achieved, "peak": peaks["tflops"],
            "unit": "TFLOP/s", "frac": achieved / peaks["tflops"], "peak_source": peaks["source"], "launches": int(n_score),
            "avg_launch_ms": ms_score / max(n_score, 1), "flops_per_row": fl, "rows_per_launch": rows_score / max(n_score, 1),
            "kernel_share_of_step": ms_score / max(sum(v[0] for v in kinds.values()), 1e-9)}


def block_beam16(args, local, dev):
    """BASELINE config 3 (QINCo2-L 8x8, A=16, beam 16) beside the beam-1 headline: value, roofline, parity, clocks."""
    import torch
    from qinco_b200.model import QINCo
    wl = WORKLOADS["c3"]
    n = args.beam16_n
    cfg, w, x_host = make_model_inputs(wl, n, 0)
    model = QINCo(cfg, w, device=dev)
    x_dev = x_host.to(dev)
    k = 3
    ms, kinds, launches, clocks, codes = timed_encode(model, x_dev, k, 1, local)
    peaks = load_peaks()
    out = {"workload": wl["name"], "vectors_per_step": n, "steps": k, "warmup": 1, "value": n * k / (ms * 1e-3), "unit": UNIT,
           "ms_per_step": ms / k, "gpu_launches": int(launches), "flops_min_per_vector": encode_flops_min_per_vector(cfg),
           "tflops_whole_step": n * k / (ms * 1e-3) * encode_flops_min_per_vector(cfg) / 1e12,
           "roofline": score_roofline(cfg, kinds, peaks), "kernel_ms_per_step": {kk: v[0] / k for kk, v in kinds.items()},
           "clocks": clocks}
    if not args.no_cpu_baseline:
        rate, n_s, ref_codes, ref_xhat, threads, port, dec_rate = cpu_port_rate(cfg, w, x_host[:contract_sample(cfg)].numpy())
        out["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "decode_value": dec_rate,
                               "sample": f"first {n_s} vectors, oracle/torch_port.py"}
        ours = codes[:n_s].cpu().numpy().T.astype(np.int64)
        out["parity"] = parity_block(cfg, model, port, x_host[:n_s].numpy(), ours, ref_codes, ref_xhat, dev)
    model.synchronize()
    model._h.close()
    return out


def block_torch_gpu(cfg, w, x_host, dev, n_sample):
    """The reference's operator sequence run by PyTorch on THIS GPU (oracle/torch_port.py, device=cuda): fp32, and all-fp16
    like the reference's own GPU wrapper (qinco_inference.py:303-317).  Timed with CUDA events after one warm-up pass."""
    import torch
    from oracle.torch_port import TorchPort
    out = {"sample": f"first {n_sample} vectors of the workload, candidate rows processed in chunks of 2^22",
           "what": "oracle/torch_port.py on cuda:0 (torch Linear/cat/bmm/topk launches per step, no hoisting)"}
    xs = x_host[:n_sample]
    for name, dt in (("fp32", torch.float32), ("fp16", torch.float16)):
        try:
            port = TorchPort(cfg, w, device=dev, dtype=dt)
            port.encode(xs[:4096], max_rows=1 << 22)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            codes, xhat = port.encode(xs, max_rows=1 << 22)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            d0.record()
            port.decode(codes)
            d1.record()
            torch.cuda.synchronize()
            out[name] = {"value": n_sample / (ms * 1e-3), "unit": UNIT, "decode_value": n_sample / (d0.elapsed_time(d1) * 1e-3),
                         "encode_mse": float(((xs.to(dev).float() - xhat.float()) ** 2).sum(1).mean().item())}
            del port
            torch.cuda.empty_cache()
        except Exception as e:          # an out-of-memory torch path must not take the bench line down
            out[name] = {"error": f"{type(e).__name__}: {str(e)[:120]}"}
    return out


def block_c1(args, dev):
    """BASELINE config 1: QINCo1 8x8 d=128, 10 000 vectors, encode AND decode, with the reference's own protocol
    (compute_MSE, qinco_tasks.py:98-145: batches of 1024, a forced host read closing each timer; us per vector), on the host
    cores (PyTorch-CPU port of the reference) and through the B200 path (same protocol, and codec.encode/decode bs=4096)."""
    import io as _io
    import torch
    from oracle.torch_port import TorchPort
    from qinco_b200 import codec, tasks
    from qinco_b200.model import QINCo
    wl = WORKLOADS["q1"]
    n = args.c1_rows
    cfg, w, x_host = make_model_inputs(wl, n, 0)
    xs = x_host.numpy()
    threads = os.cpu_count() or 1
    port = TorchPort(cfg, w, threads=threads)
    port.forward(xs[:64], "encode")
    t_enc = t_dec = 0.0
    codes_ref, dec_ref = [], []
    for i0 in range(0, n, 1024):
        b = xs[i0:i0 + 1024]
        t0 = time.time()
        c = port.forward(b, "encode")
        _ = float(c[-1].reshape(-1)[-1])
        t1 = time.time()
        xh = port.forward(c, "decode")
        _ = float(xh.reshape(-1)[-1])
        t2 = time.time()
        t_enc += t1 - t0
        t_dec += t2 - t1
        codes_ref.append(c.numpy())
        dec_ref.append(xh.numpy())
    codes_ref, dec_ref = np.concatenate(codes_ref, 1), np.concatenate(dec_ref)
    out = {"workload": "QINCo1 8x8 K=256 d=128 L=16 beam=1, 10k synthetic fp32 vectors (BASELINE config 1)", "vectors": n,
           "batch": 1024,
           "cpu": {"kind": "port", "cores": torch.get_num_threads(), "encode_us_per_vector": t_enc / n * 1e6,
                   "decode_us_per_vector": t_dec / n * 1e6, "encode_vectors_per_s": n / t_enc, "decode_vectors_per_s": n / t_dec,
                   "mse": float(((xs - dec_ref) ** 2).sum(1).mean())}}
    model = QINCo(cfg, w, device=dev)
    res = tasks.compute_MSE(model, xs, 1024, dev, timed=False, out=lambda *a, **k: None)
    out["b200_same_protocol"] = {"encode_us_per_vector": res["t_encode"] / n * 1e6, "decode_us_per_vector": res["t_decode"] / n * 1e6,
                                 "encode_vectors_per_s": n / res["t_encode"], "decode_vectors_per_s": n / res["t_decode"],
                                 "mse": res["MSE"], "api": "model(batch, step=...) per 1024-row batch with a host read (tasks.compute_MSE)"}
    ours = model(x_host.to(dev), step="encode").cpu().numpy()
    out["parity"] = parity_block(cfg, model, port, xs, ours, codes_ref, port.decode(codes_ref).numpy(), dev)
    model.synchronize()
    model._h.close()
    v1 = codec.QINCoV1(cfg=cfg, weights=w, db_scale=1.0, device=dev)
    codec.encode(v1, xs[:4096], bs=4096, verbose=False)
    t0 = time.time()
    c1 = codec.encode(v1, xs, bs=4096, verbose=False)
    t1 = time.time()
    y1 = codec.decode(v1, c1, bs=4096, verbose=False)
    t2 = time.time()
    out["b200_codec_qinco"] = {"encode_us_per_vector": (t1 - t0) / n * 1e6, "decode_us_per_vector": (t2 - t1) / n * 1e6,
                               "mse": float(((xs - y1) ** 2).sum(1).mean()), "codes_equal_model_path": bool(np.array_equal(c1.T, ours)),
                               "api": "codec.encode / codec.decode (numpy in, numpy out, bs=4096)"}
    v1._m._h.close()
    return out


def block_small_batch(model, x_dev, dev):
    """What every caller of the reference does: cfg.batch = 1024 rows per call, then a sync (search_tasks.py:107-116)."""
    import torch
    out = {}
    for bs in (1024, 4096):
        xb = x_dev[:bs].contiguous()
        for _ in range(3):
            c = model(xb, step="encode")
            model(c, step="decode")
        torch.cuda.synchronize()
        k = 20
        t0 = time.perf_counter()
        for _ in range(k):
            c = model(xb, step="encode")
            _ = float(c[-1].reshape(-1)[-1].cpu())
        t1 = time.perf_counter()
        for _ in range(k):
            y = model(c, step="decode")
            _ = float(y.reshape(-1)[-1].cpu())
        t2 = time.perf_counter()
        out[str(bs)] = {"encode_vectors_per_s": bs * k / (t1 - t0), "encode_ms_per_call": (t1 - t0) / k * 1e3,
                        "decode_vectors_per_s": bs * k / (t2 - t1), "decode_ms_per_call": (t2 - t1) / k * 1e3}
    out["api"] = 'model(x, step="encode") / model(codes, step="decode") per call, each closed by a host read'
    return out


def run_reference(args, wl):
    """--impl reference: the reference's CPU path, rank 0 only, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle.torch_port import TorchPort
    cfg, w, x = make_model_inputs(wl, 4096, 0)
    threads = os.cpu_count() or 1
    port = TorchPort(cfg, w, threads=threads)
    xn = x.numpy()
    t0 = time.perf_counter()
    port.encode(xn[:128])
    rate0 = 128 / (time.perf_counter() - t0)
    n = int(max(64, min(4096, rate0 * 4.0)) // 64 * 64)       # ~4 s of CPU work per step
    for _ in range(args.warmup):
        port.encode(xn[:max(64, n // 8)])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        port.encode(xn[:n])
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    sample = f"{n} of the workload's vectors per step (bounded CPU sample), torch CPU fp32, {torch.get_num_threads()} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "vectors_per_step": n, "device": "cpu"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def traffic_of(workload):
    """Measured DRAM bytes per launch of the workload's dominant kernel (ncu --set full capture, profiles/traffic.json)."""
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    return json.load(open(tp)).get(workload) if os.path.exists(tp) else None


def main_pairwise(args):
    """--workload c5pw: the pairwise additive decoder at the Contriever shape of BASELINE
    config 5 (d=768, 8x8 codes -> 16 tables of 65 536 rows, 3.2 GB).  HBM-bound gather-accumulate: the roofline block
    is bytes, not flops."""
    import torch
    from qinco_b200.pairwise import PairwiseDecoderIVF
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    D, M, K, Mt, ivf_K = 768, 8, 256, 16, 4096
    n = args.n or 500_000
    g = torch.Generator().manual_seed(77)
    book = torch.randn((Mt, K * K, D), generator=g, dtype=torch.float32)
    comb = torch.stack([torch.randint(0, M + 5, (Mt,), generator=g), torch.randint(0, M + 5, (Mt,), generator=g)])
    imap = torch.randint(0, K, (ivf_K, 5), generator=g)
    dec = PairwiseDecoderIVF(dict(codebook_MKD=book, combine_mvals_m=comb, ivf_code_map=imap), K=K, M=M, device=dev)
    g2 = torch.Generator().manual_seed(1234 + rank)
    codes_h = torch.randint(0, K, (n, M), generator=g2, dtype=torch.uint8).pin_memory()
    ivf_h = torch.randint(0, ivf_K, (n,), generator=g2, dtype=torch.int32).pin_memory()
    codes, ivf = codes_h.to(dev), ivf_h.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        out = dec.decode_u8(codes, ivf)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = dec.launch_count
    e0.record()
    for _ in range(args.steps):
        out = dec.decode_u8(codes, ivf)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    clocks = sampler.stop() if sampler else None
    launches = dec.launch_count - l0
    value = world * n * args.steps / (ms * 1e-3)
    bytes_per_vec = Mt * D * 4 + D * 4 + M + 4
    # end to end: host codes in, decoded vectors back on the host (what the search's re-ranking loop does per batch)
    e2e = None
    if not args.no_e2e:
        ne = min(n, 65536)
        out_pin = torch.empty((ne, D), dtype=torch.float32).pin_memory()
        barrier()
        t0 = time.perf_counter()
        k = max(1, min(args.steps, 3))
        for _ in range(k):
            o = dec.decode_u8(codes_h[:ne].to(dev, non_blocking=True), ivf_h[:ne].to(dev, non_blocking=True))
            out_pin.copy_(o, non_blocking=True)
            torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        e2e = {"value": world * ne * k / dt, "unit": "vectors/s", "h2d_bytes_per_step": ne * (M + 4), "d2h_bytes_per_step": ne * D * 4,
               "api": "PairwiseDecoderIVF.decode_u8 on pinned host codes + copy of the decoded vectors back to the host", "steps": k}
    if rank == 0:
        peaks = load_peaks()
        achieved = value / world * bytes_per_vec / 1e9
        line = {"metric": "vectors/sec decoded (pairwise additive decoder, d=768, 16 tables of 65536 rows)", "value": value,
                "unit": "vectors/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "pairwise decoder forward, Contriever shape of BASELINE config 5", "D": D, "M": M, "K": K,
                           "tables": Mt, "ivf_K": ivf_K, "vectors_per_gpu_per_step": n, "table_bytes": dec.table_bytes,
                           "l2": "table (3.2 GB) and outputs far larger than L2"},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "kernel": "qb_pairwise_kernel", "achieved": achieved, "peak": peaks["hbm"], "unit": "GB/s",
                             "frac": achieved / peaks["hbm"], "traffic": traffic_of("c5pw") if n == 500_000 else None, "bytes_per_vector": bytes_per_vec,
                             "peak_source": "MEASURED_PEAKS.json (hbm_gbs)" if "MEASURED" in peaks["source"] else peaks["source"]},
                "clocks": clocks}
        if e2e:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline:
            from oracle import qinco_oracle as orc
            ns = 20000
            cm = codes_h[:ns].numpy().T.astype(np.int64)
            t0 = time.perf_counter()
            ref = orc.pairwise_decode(book.numpy(), comb.numpy(), imap.numpy(), K, cm, ivf_h[:ns].numpy())
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": ns / dt, "unit": "vectors/s", "cores": 1, "kind": "port",
                                    "sample": f"first {ns} vectors, oracle/qinco_oracle.py pairwise_decode (numpy gather + add)"}
            line["parity"] = {"sample": ns, "bit_identical": bool(np.array_equal(out[:ns].cpu().numpy(), ref))}
        print(json.dumps(line))
    dec.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=list(WORKLOADS) + ["c5pw"])
    ap.add_argument("--n", "--vectors", dest="n", type=int, default=0,
                    help="vectors per GPU per step (default: the workload's); use --vectors under torchrun, whose own parser trips over --n")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-decode", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the beam16 / torch_gpu_baseline / c1 / small_batch blocks of the default line")
    ap.add_argument("--beam16-n", type=int, default=100_000, help="vectors per step of the beam16 block (BASELINE config 3)")
    ap.add_argument("--c1-rows", type=int, default=10_000, help="vectors of the config-1 block (QINCo1, CPU encode + decode)")
    ap.add_argument("--plan-opts", default="", help="kernel planner overrides, e.g. slot_bytes=8192,pair=1,hc=64,max_stage=4 (pair / mcast: 0 auto, 1 off, 2 on)")
    args = ap.parse_args()
    if args.workload == "c5pw":
        args.warmup = max(args.warmup, 3)
        main_pairwise(args)
        return
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n = args.n or wl["n"]

    from qinco_b200.model import QINCo
    cfg, w, x_host = make_model_inputs(wl, n, rank)
    plan_opts = None
    if args.plan_opts:
        kv = dict(t.split("=") for t in args.plan_opts.split(","))
        plan_opts = {k: int(v) for k, v in kv.items() if k in ("hc", "slot_bytes", "max_stage", "max_slab_k", "stagger")}
        plan_opts["n_tiles"] = int(kv.get("n_tiles", 0)) | (int(kv.get("pair", 0)) << 8) | (int(kv.get("mcast", 0)) << 16)
    model = QINCo(cfg, w, device=dev, plan_opts=plan_opts)
    h = model._h
    x_pin = x_host.pin_memory()
    x_dev = x_pin.to(dev, non_blocking=True)
    from qinco_b200 import shard
    ivf = bool(cfg.get("ivf_K"))
    if ivf:      # the decode / e2e blocks below use the plain (non-IVF) entry points
        args.no_e2e = args.no_decode = True

    last_ivf = [None]

    def step():
        """One pass: this rank's rows through the kernels and (N > 1) the ONE all-gather of the uint8 codes, both inside
        qinco_b200.shard.encode_sharded -- the multi-GPU form of the reference's encode_database."""
        if world > 1:
            got = shard.encode_sharded(model, x_dev, world * n, batch=n)
            codes_all = got[1] if ivf else got
            if ivf:
                last_ivf[0] = got[0][rank * n:(rank + 1) * n]
            return codes_all[rank * n:(rank + 1) * n]
        if ivf:
            last_ivf[0], codes, _ = model.encode_ivf_u8(x_dev, normalize=True, want_xhat=False)
        else:
            codes, _ = model.encode_u8(x_dev, normalize=True, want_xhat=False)
        return codes

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    barrier()
    h.timing_read()
    h.timing_enable(True)
    launches0 = h.launch_count
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        codes = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    launches = h.launch_count - launches0
    kinds = h.timing_read()
    h.timing_enable(False)
    model.synchronize()
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * n * args.steps / (ms * 1e-3)

    # ---- decode of the codes just produced (device resident): the matching fused gather + MLP + accumulate path
    dec = None
    if not args.no_decode:
        model.decode_u8(codes, denormalize=True)
        barrier()
        # every pass timed on its own (CUDA events); the median is reported, min / max beside it
        k_dec = 5
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k_dec)]
        for a_ev, b_ev in evs:
            a_ev.record()
            xdec = model.decode_u8(codes, denormalize=True)
            b_ev.record()
        barrier()
        per_pass = sorted(a_ev.elapsed_time(b_ev) for a_ev, b_ev in evs)
        td = torch.tensor([per_pass[k_dec // 2]], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
        dec_ms = float(td.item())
        D_, M_ = cfg["D"], cfg["M"]
        dec_flops = decode_flops_per_vector(cfg)
        dec_tflops = n / (dec_ms * 1e-3) * dec_flops / 1e12
        pk = load_peaks()
        dec = {"value": world * n / (dec_ms * 1e-3), "unit": "vectors/s decoded", "ms_per_pass": dec_ms, "passes": k_dec,
               "ms_per_pass_min_max": [per_pass[0], per_pass[-1]],
               "tflops": dec_tflops, "flops_per_vector": dec_flops,
               "roofline": {"bound": "tensor", "kernel": "qb_mlp_kernel (decode)", "achieved": dec_tflops, "peak": pk["tflops"],
                            "unit": "TFLOP/s", "frac": dec_tflops / pk["tflops"], "peak_source": pk["source"],
                            "hbm_gbs_algorithmic": n / (dec_ms * 1e-3) * (4 * D_ + M_) / 1e9},
               "gpu_launches_per_pass": None,
               "mse_vs_input": float(((xdec[: min(n, 65536)] - x_dev[: min(n, 65536)]) ** 2).sum(1).mean().item())}
        l0 = h.launch_count
        model.decode_u8(codes, denormalize=True)
        dec["gpu_launches_per_pass"] = int(h.launch_count - l0)

    # ---- end to end: pinned host buffers through the C-ABI host call (H2D + encode + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        x_np = x_pin.numpy()
        h.encode_host(x_np[: min(n, 65536)], normalize=True)       # warm the staging buffers
        barrier()
        t0 = time.perf_counter()
        k_e2e = max(1, min(args.steps, 3))
        for _ in range(k_e2e):
            codes_host, _ = h.encode_host(x_np, normalize=True)
            if world > 1:
                gathered = torch.empty((world * n, cfg["M"]), dtype=torch.uint8, device=dev)
                dist.all_gather_into_tensor(gathered, torch.from_numpy(codes_host).to(dev))
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * n * k_e2e / float(tt.item()), "unit": UNIT, "h2d_bytes_per_step": n * cfg["D"] * 4,
               "d2h_bytes_per_step": n * cfg["M"], "api": "qb_encode_host (C ABI, host buffers; what codec.encode calls)",
               "steps": k_e2e}
        assert np.array_equal(codes_host, codes.cpu().numpy()), "host path and device path disagree"

    if rank == 0:
        peaks = load_peaks()
        step_ms = {k: v[0] / args.steps for k, v in kinds.items()}
        roof = score_roofline(cfg, kinds, peaks)
        roof["traffic"] = traffic_of(args.workload)
        roof["hbm_gbs_algorithmic"] = value / world * (4 * cfg["D"] + cfg["M"]) / 1e9
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 operands, f32 accumulate/residual/distances", "data": "synthetic",
            "config": {"workload": wl["name"], "vectors_per_gpu_per_step": n, "global_vectors_per_step": world * n,
                       "sharding": f"rows over {world} rank(s), one NCCL all-gather of uint8 codes per step (shard.encode_sharded)",
                       "l2": f"inputs larger than L2 ({n * cfg['D'] * 4 / 2**20:.0f} MiB of vectors per step)"},
            "gpu_launches": int(launches),
            "flops_min_per_vector": encode_flops_min_per_vector(cfg),
            "tflops_whole_step": value / world * encode_flops_min_per_vector(cfg) / 1e12,
            "roofline": roof,
            "kernel_ms_per_step": step_ms,
            "clocks": clocks,
        }
        if e2e:
            out["e2e"] = e2e
        if dec:
            out["decode"] = dec
        if ivf and kinds["ivf"][1] > 0:
            # the IVF arg-min launch on its own: useful flops = 2 D ivf_K per vector (what an fp32 evaluation costs; the kernel
            # issues three fp16 products per useful one), graded against the tf32 rate = half the measured bf16 peak
            ms_ivf, n_ivf, rows_ivf = kinds["ivf"]
            useful = rows_ivf * 2.0 * cfg["D"] * cfg["ivf_K"] / (ms_ivf * 1e-3) / 1e12
            out["ivf"] = {"kernel": "qb_ivf_tc_kernel" if cfg["D"] <= 128 else "qb_ivf_assign_kernel", "centroids": cfg["ivf_K"],
                          "ms_per_step": ms_ivf / args.steps, "vectors_per_s_ivf_only": rows_ivf / (ms_ivf * 1e-3),
                          "useful_tflops": useful, "issued_tflops_fp16": 3 * useful if cfg["D"] <= 128 else None,
                          "roofline": {"bound": "tensor", "peak": peaks["tflops"] / 2, "unit": "TFLOP/s (tf32-equivalent)",
                                       "achieved": useful, "frac": useful / (peaks["tflops"] / 2),
                                       "peak_source": "half of " + peaks["source"]}}
        if world == 1 and not args.no_cpu_baseline:
            ns = min(n, contract_sample(cfg))
            rate, n_s, ref_codes, ref_xhat, threads, port, dec_rate = cpu_port_rate(cfg, w, x_host[:ns].numpy())
            out["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                   "decode": {"value": dec_rate, "unit": "vectors/s decoded"},
                                   "sample": f"first {n_s} vectors of the workload (the contract's parity sample, SURVEY 8d), "
                                             f"oracle/torch_port.py (reference op sequence in PyTorch CPU fp32), "
                                             f"os.cpu_count()={os.cpu_count()}"}
            # parity on the same sample (decode MSE vs ref; encode MSE of our codes decoded by the reference arithmetic)
            ours = codes[:n_s].cpu().numpy().T.astype(np.int64)
            if ivf:      # the reference's code matrix has the IVF code as row 0
                ours = np.concatenate([last_ivf[0][:n_s].cpu().numpy().astype(np.int64)[None, :], ours])
            out["parity"] = parity_block(cfg, model, port, x_host[:n_s].numpy(), ours, ref_codes, ref_xhat, dev)
        if world == 1 and not args.no_extras and args.workload == "c2":
            out["small_batch"] = block_small_batch(model, x_dev, dev)
            model.synchronize()
            model._h.close()
            del x_dev
            torch.cuda.empty_cache()
            out["torch_gpu_baseline"] = block_torch_gpu(cfg, w, x_host, dev, min(n, 65536))
            out["beam16"] = block_beam16(args, local, dev)
            if not args.no_cpu_baseline:
                out["c1"] = block_c1(args, dev)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
