#!/usr/bin/env python
"""Device timeline of one CUDA-graph-replayed training step on one GPU (torch.profiler): start offset, duration, kernel.
    python scripts/timeline_step.py [--small]"""
import argparse
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import two_tower_augmented_with_adaptive_mimic_mechanism_b200 as tt  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    a = ap.parse_args()
    c = dict(bench.CFG)
    if a.small:
        c.update(NU=c["NU"] // 8, NI=c["NI"] // 8)
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(1)
    ux, ix = bench.make_features(c["NI"], c["NU"], c["F"], c["n_cat"], c["n_auth"], dev, gen)
    tower = {"type": "tower", "id_embedding": {"params": {"embedding_dim": c["D"], "sparse": True}},
             "feature_encoder": {"type": "mlp", "hidden_dims": [c["H"]], "activation": "relu", "output_dim": c["D"], "dropout": 0.0},
             "fusion": "gated", "adaptive_mimic": {"hidden_dim": c["Hg"]}}
    model = tt.TwoTowerModel(tt.build_tower_encoder(tower, num_embeddings=c["NU"], feature_dim=c["F"], device=dev),
                             tt.build_tower_encoder(tower, num_embeddings=c["NI"], feature_dim=c["F"], device=dev),
                             adaptive_mimic=tt.AdaptiveMimicMechanism(num_users=c["NU"], num_items=c["NI"], embedding_dim=c["D"]).to(dev))
    eng = tt.FusedEngine(model, optimizer="adamw", lr=1e-3, weight_decay=0.01, precision="tf32",
                         loss_weights={"mimic_user": 0.15, "mimic_item": 0.15}, max_steps=256)
    users, pos, neg = bench.make_batches(24, c, dev, gen)
    for s in range(12):
        eng.train_step(users[s], pos[s], neg[s], ux, ix, graph=not a.no_graph)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for s in range(12, 18):
            eng.train_step(users[s], pos[s], neg[s], ux, ix, graph=not a.no_graph)
        torch.cuda.synchronize()
    ev = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
    starts = [e.time_range.start for e in ev if "advance_step" in e.name]
    lo, hi = starts[-2], starts[-1]
    busy = 0.0
    for e in ev:
        if lo <= e.time_range.start < hi:
            d = e.time_range.end - e.time_range.start
            busy += d
            print(f"{e.time_range.start - lo:8.1f} {d:7.1f} us  {e.name[:80]}")
    print(f"step span {hi - lo:.1f} us, summed kernel time {busy:.1f} us")


if __name__ == "__main__":
    main()
