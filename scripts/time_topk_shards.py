"""One rank's share of the item-sharded retrieval, on one GPU: local top-K of all Q queries over NI/W items (per-kernel
times from torch.profiler) and the W-way merge of a query block.  Shows what does not shrink with the shard."""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from torch.profiler import profile, ProfilerActivity
from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import functional as F

Q, NI, D, K = 100_000, 2_000_000, 96, 100
g = torch.Generator(device="cuda").manual_seed(3)
items = (torch.randn((NI, D), device="cuda", generator=g) * 0.3).bfloat16()
q = (torch.randn((Q, D), device="cuda", generator=g) * 0.3).bfloat16()
for W in (1, 2, 4, 8):
    sh = items[: NI // W].contiguous()
    F.topk(q, sh, K); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        ids, sc = F.topk(q, sh, K)
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    per = {}
    for e in ev:
        per[e.name[:48]] = per.get(e.name[:48], 0.0) + (e.time_range.end - e.time_range.start) / 1e3
    tot = sum(per.values())
    # the merge of this rank's query block: [Q/W, W, K]
    qb = Q // W
    mi = torch.randint(0, NI, (qb, W, K), device="cuda", generator=g)
    ms = torch.randn((qb, W, K), device="cuda", generator=g).sort(dim=2, descending=True).values
    F.topk_merge(mi, ms, K); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); F.topk_merge(mi, ms, K); e1.record(); torch.cuda.synchronize()
    print(json.dumps({"W": W, "items_per_rank": NI // W, "local_topk_ms": round(tot, 3), "merge_ms": round(e0.elapsed_time(e1), 3),
                      "kernels_ms": {k: round(v, 3) for k, v in sorted(per.items(), key=lambda x: -x[1])}}))
