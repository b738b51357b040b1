#!/usr/bin/env python
"""Print a phase timeline from a QB_MLP_TRACE dump (CTA 0): usage trace_summary.py file [n_events]"""
import sys
ev = {0: [], 1: [], 2: []}
for line in open(sys.argv[1]):
    if line.startswith("#"):
        print(line.strip()); continue
    r, t, i = line.split()
    ev[int(r)].append((int(t), int(i, 16)))
t0 = min(e[0][0] for e in ev.values() if e)
names = {1: "init start", 2: "init done (AE_READY)", 3: "HACC_FULL seen", 4: "H-epi done (AH_READY)", 5: "EACC_FULL seen", 6: "E-epi done (AE_READY)", 8: "tile done"}
allev = []
names[7] = "final done"; names[9] = "pre-wait start"; names[10] = "prefetch issued"; names[11] = "cb loads issued"; names[12] = "final: ld issue"; names[13] = "final: ld landed"; names[14] = "init: stores issued"
for t, i in ev[0]: allev.append((t - t0, "EPI ", ("Y " if i & 0x80 else "X ") + names.get(i & 0x7f, hex(i))))
for t, i in ev[1]:
    kind = {0x200: "ops ready, wait slab", 0x300: "slab ready, issue", 0x400: "issued"}[i & 0xf00]
    allev.append((t - t0, "MMA ", f"{'Y' if i & 0x80 else 'X'} op {i & 0x7f:2d} {kind}"))
for t, i in ev[2]:
    kind = {0x200: "ops ready, wait slab", 0x300: "slab ready, issue", 0x400: "issued"}[i & 0xf00]
    allev.append((t - t0, "MMA ", f"Y op {i & 0x7f:2d} {kind}"))
allev.sort()
n = int(sys.argv[2]) if len(sys.argv) > 2 else 150
prev = 0
for t, who, what in allev[:n]:
    print(f"{t:9d} (+{t - prev:6d}) {who} {what}")
    prev = t
# per-tile duration
starts = [t for t, i in ev[0] if i == 1]
if len(starts) > 2:
    d = [b - a for a, b in zip(starts, starts[1:])]
    print("tile period (cycles): min", min(d), "median", sorted(d)[len(d) // 2], "max", max(d), "n", len(d))
