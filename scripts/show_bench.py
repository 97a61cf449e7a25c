#!/usr/bin/env python
"""Pretty-print a bench.py JSON line (the kernel-class table in particular)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "e2e", "gpu_launches", "clocks", "fp32", "hook", "cpu_baseline"):
    if k in d:
        print(k, d[k])
r = d["roofline"]
print("TOP:", r["kernel"], f'{r["kernel_ms"]*1e3:.1f} us', f'frac {r["frac"]:.3f}')
for k in r.get("kernels", []):
    print(f'{k["ms"]*1e3:8.1f} us  {k["achieved"]:8.0f} GB/s  {k["frac"]:.2f}  {k.get("tflops",0):6.1f} TF  {k["kernel"][:72]}', k.get("error", ""))
print("sum of classes:", r.get("kernels_total_ms"))
for name, v in (d.get("retrieval") or {}).items():
    if isinstance(v, dict):
        print(name, {a: b for a, b in v.items() if a not in ("roofline",)})
