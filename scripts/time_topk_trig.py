"""Sweep of the list-compaction trigger of the bf16 top-K (TTAM_TOPK_TRIG) at config-3 size; results must not change."""
import sys, json, os
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import functional as F
Q, N, K, D = 100_000, 2_000_000, 100, int(sys.argv[1]) if len(sys.argv) > 1 else 96
g = torch.Generator(device="cuda").manual_seed(3)
ib = (torch.randn((N, D), device="cuda", generator=g) * 0.3).bfloat16()
qb = (torch.randn((Q, D), device="cuda", generator=g) * 0.3).bfloat16()
ref = None
for trig in sys.argv[2:] or ("448", "320", "256", "224", "192", "160"):
    os.environ["TTAM_TOPK_TRIG"] = trig
    out = F.topk(qb, ib, K); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = F.topk(qb, ib, K); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    if ref is None:
        ref = out
    same = torch.equal(ref[0], out[0]) and torch.equal(ref[1], out[1])
    print(json.dumps({"D": D, "trig": trig, "ms": best, "tflops": 2.0 * Q * N * D / best / 1e9, "same_result": same}), flush=True)
