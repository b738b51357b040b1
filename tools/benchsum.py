#!/usr/bin/env python
"""One-line summary of bench.py JSON lines: usage benchsum.py file..."""
import json, sys
for f in sys.argv[1:]:
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        k = j.get("kernel_ms_per_step", {})
        print(f"{f}: {j['value']:.0f} vec/s frac {j['roofline']['frac']:.3f} ms/step {j['ms_per_step']:.1f} "
              f"kernels {{{', '.join(f'{a}: {b:.1f}' for a, b in k.items())}}} clocks {j['clocks'].get('sm_mhz')} {j['clocks'].get('reasons')} "
              f"parity {j.get('parity', {}).get('vectors_with_identical_codes')}")
    except Exception as e:  # noqa: BLE001
        print(f, "unreadable:", e)
