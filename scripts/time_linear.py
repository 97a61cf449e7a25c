"""Time the tower GEMMs at BASELINE config-2 shapes (item tower: 49152 rows): fp32 SIMT vs TF32 tcgen05, CUDA events."""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import functional as F

R, NI, Fd, H, D = 49152, 2_000_000, 605, 192, 96
only = sys.argv[1] if len(sys.argv) > 1 else "all"
g = torch.Generator(device="cuda").manual_seed(1)
X = F.round_tf32_(F.pad_cols(torch.randn((NI, Fd), device="cuda", generator=g), always_copy=True))   # engine layout
idx = torch.randint(0, NI, (R,), device="cuda", generator=g)
W1 = F.round_tf32_(F.pad_cols(torch.randn((H, Fd), device="cuda", generator=g) * 0.05, always_copy=True))
b1 = torch.zeros(H, device="cuda")
W2 = torch.randn((D, H), device="cuda", generator=g) * 0.05
h = torch.randn((R, H), device="cuda", generator=g)
dy = torch.randn((R, D), device="cuda", generator=g)
dh = torch.randn((R, H), device="cuda", generator=g)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return min(ts)

cases = {
    "fwd1": (lambda p: F.linear_fwd(X, W1, b1, gather=idx, act="relu", precision=p, x_rounded=True, w_rounded=True), R * (608 * 4 + 8) + R * H * 4, 2.0 * R * Fd * H),
    "fwd2": (lambda p: F.linear_fwd(h, W2, None, precision=p), R * H * 4 + R * D * 4, 2.0 * R * H * D),
    "dgrad2": (lambda p: F.linear_dgrad(dy, W2, aux=h, relu_mask=True, precision=p), R * D * 4 + 2 * R * H * 4, 2.0 * R * H * D),
    "wgrad1": (lambda p: F.linear_wgrad(dh, X, gather=idx, precision=p, x_rounded=True), R * (608 * 4 + 8) + R * H * 4, 2.0 * R * Fd * H),
    "wgrad2": (lambda p: F.linear_wgrad(dy, h, precision=p), R * (H + D) * 4, 2.0 * R * H * D),
}
for name, (fn, nbytes, flops) in cases.items():
    if only != "all" and only != name:
        continue
    out = {"case": name}
    for p in ("fp32", "tf32"):
        us = timeit(lambda: fn(p))
        out[p] = {"us": round(us, 1), "GBs": round(nbytes / us / 1e3, 1), "TFLOPs": round(flops / us / 1e6, 1)}
    print(json.dumps(out))
