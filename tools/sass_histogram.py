#!/usr/bin/env python
"""SASS opcode histogram per kernel of the built library: python tools/sass_histogram.py > profiles/rNN_sass_histogram.txt"""
import collections, os, re, subprocess, sys
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "qinco_b200", "libqinco_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, hist = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1); hist[fn] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        hist[fn][m.group(2).split(".")[0]] += 1
keys = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "REDUX", "HMMA", "FFMA", "FFMA2", "LDS", "STS", "LDG", "STG", "ATOMG", "RED", "BAR", "SHFL", "LDL", "STL"]
print("# SASS opcode histogram of qinco_b200/libqinco_b200.so (cuobjdump -sass, sm_100a), one line per kernel")
print("# UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk (TMA bulk copy, incl. multicast),")
print("# SYNCS = mbarrier ops, REDUX = redux.sync, ATOMG = global atomics (error word, packed arg-min), LDL / STL = local-memory (spill) instructions")
print("# qb_mlp_kernel<kScore, kResident, kPair, kLoop, kFuse, kMcast>")
for f, c in hist.items():
    d = subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip().replace("(anonymous namespace)::", "")
    d = re.sub(r"\(.*", "", d).replace("void ", "")
    print(f"{d[:72]:72s} total={sum(c.values()):6d} " + " ".join(f"{k}={c[k]}" for k in keys if c[k]))
