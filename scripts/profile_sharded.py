"""torch.profiler timeline of the row-sharded training step (launch under torchrun, N >= 2): where the step's wall time
goes on rank 0 - our kernels, NCCL, torch glue ops, idle gaps."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch, torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
import bench
import two_tower_augmented_with_adaptive_mimic_mechanism_b200 as tt
from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import sharding as S

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
c = dict(bench.CFG); c.update(NU=c["NU"] // 4, NI=c["NI"] // 4)
gen = torch.Generator(device=dev).manual_seed(1 + rank)
nu_l, ni_l = S.shard_size(c["NU"], rank, world), S.shard_size(c["NI"], rank, world)
ux, ix = bench.make_features(ni_l, nu_l, c["F"], c["n_cat"], c["n_auth"], dev, gen)
tower = {"type": "tower", "id_embedding": {"params": {"embedding_dim": c["D"], "sparse": True}},
         "feature_encoder": {"type": "mlp", "hidden_dims": [c["H"]], "activation": "relu", "output_dim": c["D"], "dropout": 0.0},
         "fusion": "gated", "adaptive_mimic": {"hidden_dim": c["Hg"]}}
model = tt.TwoTowerModel(tt.build_tower_encoder(tower, num_embeddings=nu_l, feature_dim=c["F"], device=dev),
                         tt.build_tower_encoder(tower, num_embeddings=ni_l, feature_dim=c["F"], device=dev),
                         adaptive_mimic=tt.AdaptiveMimicMechanism(num_users=nu_l, num_items=ni_l, embedding_dim=c["D"]).to(dev))
eng = tt.FusedEngine(model, optimizer="adamw", lr=1e-3, weight_decay=0.01, precision="tf32",
                     loss_weights={"mimic_user": 0.15, "mimic_item": 0.15}, max_steps=64)
route = os.environ.get("TTAM_ROUTE", "peer")
sh = tt.ShardedEngine(eng, static=route != "dynamic", peer=route == "peer")
graph = route != "dynamic"
users, pos, neg = bench.make_batches(16, c, dev, gen)
for s in range(8):          # calibration steps (eager) + the step that records the graphs + one replay
    sh.train_step(users[s], pos[s], neg[s], ux, ix, graph=graph)
dist.barrier(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for s in range(8, 14):
        sh.train_step(users[s], pos[s], neg[s], ux, ix, graph=graph)
    torch.cuda.synchronize()
if rank == 0 and os.environ.get("TTAM_TIMELINE"):
    # one step's device timeline (rank 0): start offset, duration, stream, kernel
    ev = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
    starts = [e.time_range.start for e in ev if "advance_step" in e.name]
    lo, hi = starts[-2], starts[-1]
    for e in ev:
        if lo <= e.time_range.start < hi:
            print(f"{e.time_range.start - lo:8.1f} {e.time_range.end - e.time_range.start:7.1f} us  {e.name[:70]}")
if rank == 0:
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    t0, t1 = min(e.time_range.start for e in ev), max(e.time_range.end for e in ev)
    busy = sum(e.time_range.end - e.time_range.start for e in ev)
    print(f"route {route}, world {world}, slots {sh.last_exchange_rows}, fallback steps {sh.fallback_steps}")
    print(f"6 steps: span {(t1 - t0) / 6:.0f} us/step, GPU busy {busy / 6:.0f} us/step")
    agg = {}
    for e in ev:
        k = ("nccl" if "nccl" in e.name.lower() else "ttam/cub" if ("ttam" in e.name or "tcg" in e.name or "cub" in e.name.lower()) else "torch glue: " + e.name[:60])
        a = agg.setdefault(k, [0.0, 0]); a[0] += e.time_range.end - e.time_range.start; a[1] += 1
    for k, (v, n) in sorted(agg.items(), key=lambda x: -x[1][0])[:18]:
        print(f"{v / 6:8.1f} us/step {n / 6:6.1f} launches/step  {k}")
dist.barrier(); torch.cuda.synchronize(); sys.stdout.flush()
os._exit(0)
