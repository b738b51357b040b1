"""Decode-only timing probe: python tools/decode_probe.py WORKLOAD N [passes]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from qinco_b200.model import QINCo
wl = bench.WORKLOADS[sys.argv[1]]
n = int(sys.argv[2])
k = int(sys.argv[3]) if len(sys.argv) > 3 else 5
cfg, w, x = bench.make_model_inputs(wl, 1024, 0)
m = QINCo(cfg, w, device="cuda:0")
g = torch.Generator().manual_seed(1)
codes = torch.randint(0, cfg["K"], (n, cfg["M"]), generator=g, dtype=torch.uint8).cuda()
for _ in range(2):
    m.decode_u8(codes)
torch.cuda.synchronize()
ts = []
for _ in range(k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); m.decode_u8(codes); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
m.synchronize()
print(sys.argv[1], n, "env", {k: v for k, v in os.environ.items() if k.startswith("QB_")}, "ms", [round(t, 2) for t in ts], "info", m._h.info(1))
