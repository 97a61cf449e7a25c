"""Time every top-K path with CUDA events: bf16 tensor-core at D = 96 / 256, the fp32 index with its candidate pass on the
tensor cores (3-way bf16 split), and the fp32 SIMT kernel.  usage: time_topk_modes.py [Q] [N]"""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import functional as F

Q = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
K = 100
g = torch.Generator(device="cuda").manual_seed(3)


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


for D in (96, 256):
    items = torch.randn((N, D), device="cuda", generator=g) * 0.3
    q = torch.randn((Q, D), device="cuda", generator=g) * 0.3
    ib, qb = items.bfloat16(), q.bfloat16()
    ms = timed(lambda: F.topk(qb, ib, K))
    print(json.dumps({"path": "bf16 tcgen05", "Q": Q, "N": N, "D": D, "ms": ms, "qps": Q / ms * 1e3,
                      "tflops": 2.0 * Q * N * D / ms / 1e9}), flush=True)
    del ib, qb
    split = F.split_bf16x3(items, item_layout=True)
    ms = timed(lambda: F.topk_f32_tc(q, items, split, K))
    print(json.dumps({"path": "fp32 index, tcgen05 candidates (hi|lo split, 3 products)", "Q": Q, "N": N, "D": D, "ms": ms,
                      "qps": Q / ms * 1e3, "tflops_mma": 6.0 * Q * N * D / ms / 1e9}), flush=True)
    del split
    Qs = min(Q, 4096)
    ms = timed(lambda: F.topk(q[:Qs], items, K), reps=2)
    print(json.dumps({"path": "fp32 SIMT", "Q": Qs, "N": N, "D": D, "ms": ms, "qps": Qs / ms * 1e3,
                      "tflops": 2.0 * Qs * N * D / ms / 1e9}), flush=True)
    del items, q
