#!/usr/bin/env python
"""One-line-per-kernel summary of bench.py JSON lines:  python scripts/bench_brief.py file.json [...]"""
import json
import sys

for path in sys.argv[1:]:
    for line in open(path):
        line = line.strip()
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        if "ms_per_step" not in d:
            print(path, d)
            continue
        e2e = d.get("e2e", {}).get("value", 0.0)
        print(f"{path}: {d['ms_per_step']:.4f} ms/step  value {d['value'] / 1e6:.2f} M/s  e2e {e2e / 1e6:.2f} M/s  n_gpus {d.get('n_gpus')}")
        for k in d.get("roofline", {}).get("kernels", []):
            print(f"   {k['ms'] * 1e3:7.1f} us  {k['frac']:.3f}  {k['kernel'][:90]}")
