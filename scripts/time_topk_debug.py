"""Time decomposition of the bf16 top-K score kernel (TTAM_TOPK_DEBUG bits: 1 = no drain, 2 = no appends, 8 = MMA does not wait)."""
import sys, json, os
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import functional as F
Q, N, K, D = 100_000, 2_000_000, 100, int(sys.argv[1]) if len(sys.argv) > 1 else 96
g = torch.Generator(device="cuda").manual_seed(3)
ib = (torch.randn((N, D), device="cuda", generator=g) * 0.3).bfloat16()
qb = (torch.randn((Q, D), device="cuda", generator=g) * 0.3).bfloat16()
for dbg in ("0", "2", "1", "9"):
    os.environ["TTAM_TOPK_DEBUG"] = dbg
    F.topk(qb, ib, K); torch.cuda.synchronize()
    best = 1e9
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); F.topk(qb, ib, K); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(json.dumps({"D": D, "debug": dbg, "ms": best, "tflops": 2.0 * Q * N * D / best / 1e9}), flush=True)
