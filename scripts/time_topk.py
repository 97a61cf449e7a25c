"""Time the bf16 tensor-core top-K at BASELINE config 3 sizes (Q x 2M x 96, K=100) with CUDA events."""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import functional as F

Q = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
D, K = 96, 100
g = torch.Generator(device="cuda").manual_seed(3)
items = (torch.randn((N, D), device="cuda", generator=g) * 0.3).bfloat16()
q = (torch.randn((Q, D), device="cuda", generator=g) * 0.3).bfloat16()
F.topk(q[:1024], items, K)
torch.cuda.synchronize()
ts = []
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ids, sc = F.topk(q, items, K); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = min(ts)
fl = 2.0 * Q * N * D
print(json.dumps({"Q": Q, "N": N, "ms": ts, "qps": Q / (ms * 1e-3), "tflops": fl / (ms * 1e-3) / 1e12}))
