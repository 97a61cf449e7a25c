#!/usr/bin/env python
"""bench.py — train samples/s (headline) and top-100 retrieval queries/s of the B200 two-tower hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): synthetic 1M users x 2M books, 96-dim towers, 605 -> 192 -> 96 feature MLPs,
gated fusion, adaptive mimic, batch 8192, 5 sampled negatives per positive, AdamW + SparseAdam.
One "step" = one optimisation step on one batch.  Prints ONE JSON line (see the keys below).

  value         samples/s with the batch index tensors already resident in HBM (CUDA-graph replay of the step)
  e2e           samples/s through the public API with HOST index buffers: pinned H2D copy of (users, pos, neg)
                and a D2H read of the loss inside the timed region, every step
  roofline      the dominant kernel of the step, timed alone with CUDA events, against MEASURED_PEAKS.json
  cpu_baseline  oracle/torch_port.py (the reference's CPU PyTorch path restated) on this box's host cores,
                on a bounded sample of the same workload
  retrieval     top-100 exact inner-product search, bf16/fp32 corpus of 2M x 96 (BASELINE.json configs[2])

N > 1 (torchrun, one rank per GPU): tables / optimiser state / feature matrices row-sharded, B samples per rank (weak
scaling), `--route peer` (default: fixed-capacity slots, row payloads by NVLink peer loads/stores, the step replayed as CUDA
graphs), `static` (same slots over NCCL all-to-alls) or `dynamic` (per-step split sizes, eager; diagnostic).

`--impl reference` times the CPU port alone (rank 0 only) and prints the same line shape.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

CFG = dict(NU=1_000_000, NI=2_000_000, D=96, H=192, Hg=96, F=605, B=8192, N=5, lr=1e-3, wd=0.01,
           lambdas=(0.15, 0.15), n_cat=300, n_auth=300)


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=d["hbm_gbs"], tf=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm=6650.0, tf=1590.0, source="fallback")


# ------------------------------------------------------------------------------------------------
# synthetic data
# ------------------------------------------------------------------------------------------------
def make_features(n_items, n_users, F, n_cat, n_auth, device, gen, cheap_users=False):
    item_x = torch.zeros((n_items, F), dtype=torch.float32, device=device)
    rows = torch.arange(n_items, device=device)
    for depth, w in enumerate((1.0, 0.5, 0.5)):           # 2-3 category columns, weights 1, 1/2
        cols = torch.randint(0, n_cat, (n_items,), device=device, generator=gen)
        keep = torch.ones(n_items, dtype=torch.bool, device=device) if depth < 2 else \
            (torch.rand(n_items, device=device, generator=gen) < 0.5)
        item_x[rows[keep], cols[keep]] = w
    # author one-hot with Zipf-1.0 popularity
    pa = 1.0 / torch.arange(1, n_auth + 1, device=device, dtype=torch.float64)
    auth = torch.multinomial(pa / pa.sum(), n_items, replacement=True, generator=gen)
    item_x[rows, n_cat + auth] = 1.0
    item_x[:, n_cat + n_auth:] = torch.randn((n_items, F - n_cat - n_auth), device=device, generator=gen)
    if cheap_users:                                       # CPU arm: timing does not depend on the feature values
        return item_x[torch.randint(0, n_items, (n_users,), device=device, generator=gen)], item_x
    user_x = torch.empty((n_users, F), dtype=torch.float32, device=device)
    for s in range(0, n_users, 65536):                    # user features = mean of 8 item rows (features.py:302-313)
        e = min(n_users, s + 65536)
        pick = torch.randint(0, n_items, (e - s, 8), device=device, generator=gen)
        user_x[s:e] = item_x[pick.reshape(-1)].view(e - s, 8, F).mean(1)
    return user_x, item_x


def make_batches(steps, c, device, gen):
    B, N = c["B"], c["N"]
    pop = 1.0 / torch.arange(1, c["NI"] + 1, device=device, dtype=torch.float64) ** 1.05
    pop = (pop / pop.sum()).float()
    users = torch.randint(0, c["NU"], (steps, B), device=device, generator=gen)
    pos = torch.multinomial(pop, steps * B, replacement=True, generator=gen).view(steps, B)
    neg = torch.randint(0, c["NI"], (steps, B, N), device=device, generator=gen)
    return users, pos, neg


class ClockSampler:
    """SM clock / throttle reasons during the timed region (B200_PROFILING.md recipe).  In-process NVML (nvidia_ml_py)
    every 25 ms (the timed region of the default run is ~45 ms): spawning nvidia-smi from a thread takes the driver lock often enough to slow the timed launches down
    (the first version of this bench measured `value` 10 % below `e2e` because of it); nvidia-smi is the fallback."""

    def __init__(self, index):
        self.rows, self.stop, self.index = [], threading.Event(), index
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
            else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        flag = lambda name: "Active" if (r & getattr(n, name, 0)) else "Not Active"
        return [str(sm), str(mx), flag("nvmlClocksThrottleReasonHwSlowdown"), flag("nvmlClocksThrottleReasonHwThermalSlowdown"),
                flag("nvmlClocksThrottleReasonSwThermalSlowdown"), flag("nvmlClocksThrottleReasonSwPowerCap")]

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop.is_set():
            try:
                if self.nvml is not None:
                    self.rows.append(self._sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.025 if self.nvml is not None else 0.5)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(r[2 + j].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's CPU PyTorch path, restated (oracle/torch_port.py)
# ------------------------------------------------------------------------------------------------
def cpu_arm(c, steps, warmup, seed=1234):
    from oracle import torch_port
    torch.manual_seed(seed)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    dev = torch.device("cpu")
    gen = torch.Generator().manual_seed(seed)
    t0 = time.time()
    user_x, item_x = make_features(c["NI"], c["NU"], c["F"], c["n_cat"], c["n_auth"], dev, gen, cheap_users=True)
    model = torch_port.Model(c["NU"], c["NI"], c["D"], c["F"], c["H"], c["Hg"], sparse=True, dropout=0.0)
    opts = torch_port.build_optimizers(model, lr=c["lr"], weight_decay=c["wd"])
    users, pos, neg = make_batches(steps + warmup, c, dev, gen)
    setup = time.time() - t0
    for s in range(warmup):
        torch_port.train_step(model, opts, users[s], pos[s], neg[s], user_x, item_x, c["lambdas"])
    t1 = time.time()
    for s in range(warmup, warmup + steps):
        # negatives come from the per-row Python sampler like in the reference (no positives dict: no rejection)
        n = torch_port.sample_negatives(users[s], c["NI"], None, c["N"])
        torch_port.train_step(model, opts, users[s], pos[s], n, user_x, item_x, c["lambdas"])
    dt = time.time() - t1
    return dict(value=steps * c["B"] / dt, ms_per_step=1e3 * dt / steps, cores=cores, setup_s=setup,
                sample=f"{steps} steps of B={c['B']} at full table sizes (NU={c['NU']}, NI={c['NI']}), {warmup} warm-up")


def cpu_retrieval(c, n_items, n_queries, k=100):
    from oracle import torch_port
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(5)
    items = (torch.randn((n_items, c["D"]), generator=g) * 0.3).bfloat16().float()
    q = (torch.randn((n_queries, c["D"]), generator=g) * 0.3).bfloat16().float()
    t0 = time.time()
    torch_port.flat_ip_topk(q, items, k)
    return n_queries / (time.time() - t0)


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("TTAM_PRECISION", "tf32"), choices=["tf32", "fp32"],
                    help="tower GEMMs: tf32 = tcgen05 kind::tf32 (round-to-nearest operands, fp32 accumulate in TMEM; the "
                         "product path), fp32 = SIMT FFMA (bit-faithful arithmetic)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-retrieval", action="store_true")
    ap.add_argument("--batch", type=int, default=None, help="samples per GPU per step (default 8192; sweep: 4096..65536)")
    ap.add_argument("--mode", default="hybrid", choices=["hybrid", "sparse", "dense"],
                    help="optimiser sweep (BASELINE configs[4]): hybrid = AdamW + SparseAdam (reference default), "
                         "sparse = embedding-only towers, mimic off, all SparseAdam, dense = sparse:false (AdamW semantics on every table)")
    ap.add_argument("--no-graph", action="store_true", help="N=1: launch the step eagerly instead of replaying a CUDA graph (diagnostic)")
    ap.add_argument("--route", default="peer", choices=["static", "peer", "dynamic"],
                    help="N>1: static = fixed-capacity slots, the whole sharded step replays as CUDA graphs; peer = static, and the "
                         "row payloads travel by NVLink peer loads/stores instead of NCCL all-to-alls; dynamic = per-step "
                         "split sizes, eager launches (diagnostic)")
    ap.add_argument("--small", action="store_true", help="1/16-size tables (debugging only; not a valid bench line)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    c = dict(CFG)
    if args.small:
        c.update(NU=c["NU"] // 16, NI=c["NI"] // 16)
    if args.batch:
        c["B"] = int(args.batch)
    W = max(args.warmup, 3)
    K = args.steps
    workload = (f"synthetic {c['NU']} users x {c['NI']} items, D={c['D']}, F={c['F']}->H={c['H']}->D MLPs, gated fusion, "
                f"adaptive mimic, B={c['B']}, {c['N']} sampled negatives, AdamW+SparseAdam")
    if args.mode == "sparse":
        workload = (f"synthetic {c['NU']} users x {c['NI']} items, D={c['D']}, embedding-only towers, no mimic, B={c['B']}, "
                    f"{c['N']} sampled negatives, all-sparse SparseAdam")
    elif args.mode == "dense":
        workload = workload.replace("AdamW+SparseAdam", "all-dense AdamW (sparse:false, lazy-exact rows)")

    if args.impl == "reference":
        if rank != 0:
            return 0
        ksteps = min(K, 8)
        r = cpu_arm(c, ksteps, min(W, 2))
        line = {"impl": "reference", "metric": "train samples/sec", "value": r["value"], "unit": "samples/s",
                "n_gpus": args.gpus, "steps": ksteps, "warmup": min(W, 2), "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload},
                "cpu_baseline": {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import two_tower_augmented_with_adaptive_mimic_mechanism_b200 as tt
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import functional as F
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import sharding as S
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    torch.manual_seed(1234 + rank)
    # N > 1: tables, optimiser state and feature matrices are ROW-SHARDED (row r on rank r % N), the batch is
    # data-parallel with B samples per rank (weak scaling): SURVEY 8(e).  N = 1: everything on the one GPU.
    nu_l, ni_l = S.shard_size(c["NU"], rank, world), S.shard_size(c["NI"], rank, world)
    user_x, item_x = make_features(ni_l, nu_l, c["F"], c["n_cat"], c["n_auth"], dev, gen)
    tower = {"type": "tower", "id_embedding": {"params": {"embedding_dim": c["D"], "sparse": args.mode != "dense"}},
             "feature_encoder": {"type": "mlp", "hidden_dims": [c["H"]], "activation": "relu", "output_dim": c["D"], "dropout": 0.0},
             "fusion": "gated", "adaptive_mimic": {"hidden_dim": c["Hg"]}}
    if args.mode == "sparse":
        tower = {"type": "embedding", "params": {"embedding_dim": c["D"], "sparse": True}}
    mimic = None if args.mode == "sparse" else \
        tt.AdaptiveMimicMechanism(num_users=nu_l, num_items=ni_l, embedding_dim=c["D"]).to(dev)
    model = tt.TwoTowerModel(tt.build_tower_encoder(tower, num_embeddings=nu_l, feature_dim=c["F"], device=dev),
                             tt.build_tower_encoder(tower, num_embeddings=ni_l, feature_dim=c["F"], device=dev),
                             adaptive_mimic=mimic)
    if world > 1:                      # replicas of the dense weights start identical
        for name, prm in model.named_parameters():
            if "embedding.weight" not in name and "augmented.weight" not in name:
                dist.broadcast(prm.data, src=0)
    eng = tt.FusedEngine(model, optimizer="adamw", lr=c["lr"], weight_decay=c["wd"], precision=args.precision,
                         loss_weights={} if mimic is None else {"mimic_user": c["lambdas"][0], "mimic_item": c["lambdas"][1]},
                         max_steps=4 * (K + W) + 64)
    sh = tt.ShardedEngine(eng, static=args.route != "dynamic", peer=args.route == "peer") if world > 1 else None
    users, pos, neg = make_batches(K + W, c, dev, gen)          # global row ids
    if sh is not None:
        # N > 1: a few extra untimed steps before the W warm-up steps: the 4 calibration steps of the slot capacities (eager),
        # the step that records the CUDA graphs (or, on the dynamic route, the steps that grow the engine's buffers), and the
        # opening of the NCCL / NVLink channels; none of it belongs in a timed region.
        pu, pp, pn = make_batches(6, c, dev, gen)
        for s in range(6):
            sh.train_step(pu[s], pp[s], pn[s], user_x, item_x, graph=not args.no_graph)
        del pu, pp, pn
    h_users, h_pos, h_neg = (t.cpu().pin_memory() for t in (users, pos, neg))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step(u, p, n):
        if sh is not None:
            return sh.train_step(u, p, n, user_x, item_x, graph=not args.no_graph)
        return eng.train_step(u, p, n, user_x, item_x, graph=not args.no_graph)

    # ---- value: batch index tensors resident in HBM (N = 1: CUDA-graph replay of the step)
    launches0 = F.lib().ttam_launch_count()
    for s in range(W):
        step(users[s], pos[s], neg[s])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        for s in range(W, W + K):
            step(users[s], pos[s], neg[s])
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1)
    # ---- e2e: host buffers -> H2D -> step -> D2H loss, every step
    d_u = torch.empty_like(users[0]); d_p = torch.empty_like(pos[0]); d_n = torch.empty_like(neg[0])
    loss_host = torch.empty(4, dtype=torch.float32).pin_memory()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for s in range(W, W + K):
        d_u.copy_(h_users[s], non_blocking=True); d_p.copy_(h_pos[s], non_blocking=True); d_n.copy_(h_neg[s], non_blocking=True)
        loss = step(d_u, d_p, d_n)
        loss_host.copy_(loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()          # the caller reads the loss every step (training.py:830)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    B, N, D, Fd, H = c["B"], c["N"], c["D"], c["F"], c["H"]

    # ---- roofline of the dominant kernel, timed alone (cold L2)
    pk = peaks()
    items_idx = torch.cat([pos[W], neg[W].reshape(-1)]) % ni_l
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def time_alone(fn, reps=5):
        tot = 0.0
        for _ in range(reps):
            flush.zero_()
            k0.record(); fn(); k1.record()
            torch.cuda.synchronize()
            tot += k0.elapsed_time(k1)
        return tot / reps

    R = items_idx.numel()
    if eng.item.fe_layers:
        # layer 1 of the item tower (feature-row gather fused into the GEMM loader): the largest byte stream of the step
        W1, b1 = eng.item.fe_layers[0]
        Xi = eng._x(item_x)
        tc = args.precision != "fp32"
        W1p = F.round_tf32_(F.pad_cols(W1, always_copy=True)) if tc else W1     # the layout the engine hands the kernel
        hd = torch.empty((R, H), device=dev)
        tk = time_alone(lambda: F.linear_fwd(Xi, W1p, b1, gather=items_idx, act="relu", out=hd, precision=args.precision,
                                             x_rounded=tc, w_rounded=tc))
        flops = 2.0 * R * Fd * H
        bytes_alg = R * (Fd * 4 + 8) + H * Fd * 4 + R * H * 4
        kname = "item tower layer 1: X[idx] . W1^T + b1, relu"
    else:
        # embedding-only towers: the item-table row gather (nn.Embedding.forward)
        out_rows = torch.empty((R, D), device=dev)
        tk = time_alone(lambda: F.gather_rows(eng.item.table, items_idx, out=out_rows))
        flops = 0.0
        bytes_alg = R * (2 * D * 4 + 8)
        kname = "item table row gather: E[idx]"
    roof = {"bound": "hbm", "achieved": bytes_alg / (tk * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s"}
    # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed `ncu --set full` capture
    traffic = None
    tfile = ROOT / "profiles" / "roofline_traffic.json"
    if tfile.exists():
        traffic = json.loads(tfile.read_text()).get(f"{kname}|{args.precision}|R={R}")
    roof.update(frac=roof["achieved"] / roof["peak"], traffic=traffic, kernel=kname,
                kernel_ms=tk, peak_source=pk["source"], algorithmic_bytes=bytes_alg, algorithmic_flops=flops,
                tensor_tflops=flops / (tk * 1e-3) / 1e12)

    steps_launched = 2 * K + W
    line = {"metric": "train samples/sec", "value": world * K * B / (ms * 1e-3), "unit": "samples/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"fp32": "f32", "tf32": "tf32"}[args.precision], "data": "synthetic",
            "config": {"workload": workload,
                       "parallelism": ("1 gpu, CUDA-graph replay" if not args.no_graph else "1 gpu, eager launches") if world == 1 else
                       f"{world} gpus: tables/features row-sharded, batch data-parallel ({B} samples per gpu), 3 all-to-all + 1 all-reduce per step, "
                       + (f"fixed-capacity slots ({sh.last_exchange_rows[0] // world}/{sh.last_exchange_rows[1] // world} user/item rows per "
                          f"rank pair), CUDA-graph replay incl. collectives, {sh.fallback_steps} dynamic-route fallback steps"
                          + (", row payloads by NVLink peer loads/stores fused into the un-bucket/re-bucket kernels (2 device barriers per step)"
                             if sh.peer else ", row payloads by equal-split NCCL all-to-alls")
                          if args.route != "dynamic" else "per-step split sizes, eager launches"),
                       "l2_policy": "inputs larger than L2: each step gathers from 11.8 GB of tables/features"},
            "e2e": {"value": world * K * B / (ms_e2e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": B * 8 * (2 + N),
                    "d2h_bytes_per_step": 16},
            "gpu_launches": None, "clocks": clk.summary(), "roofline": roof}
    if world == 1:
        line["gpu_launches"] = int(getattr(eng, "launches_per_step", 0)) * K
    else:
        line["gpu_launches"] = (int(sh.launches_per_step) * K if args.route != "dynamic" and not args.no_graph else
                                int((F.lib().ttam_launch_count() - launches0) * K / steps_launched))

    if not args.no_retrieval:
        r = bench_retrieval(tt, c, dev, pk, world=world, rank=rank)
        if rank == 0:
            line["retrieval"] = r
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        del eng, model, user_x, item_x
        torch.cuda.empty_cache()
        r = cpu_arm(c, 4, 1)
        line["cpu_baseline"] = {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": "port",
                                "sample": r["sample"], "ms_per_step": r["ms_per_step"]}
        if "retrieval" in line:
            line["retrieval"]["cpu_baseline_qps"] = cpu_retrieval(c, 200_000, 2048)
            line["retrieval"]["cpu_sample"] = "2048 queries x 200k items fp32 matmul + torch.topk, all host cores"
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        # leave without tearing the communicator down: destroy_process_group() behind CUDA graphs that hold NCCL kernels
        # was seen to hang; every rank has printed / synchronised by now
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)
    return 0


def bench_retrieval(tt, c, dev, pk, Q=100_000, NI=2_000_000, K=100, world=1, rank=0):
    """top-100 over the full 2M x 96 corpus (BASELINE configs[2]); Q queries per launch sequence.  N > 1: the corpus is
    item-sharded (NI/N contiguous items per rank), every rank scores all Q queries against its shard and merges the
    lists of its own query block after an all-to-all."""
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import functional as F
    import torch.distributed as dist
    g = torch.Generator(device=dev).manual_seed(7)              # same corpus / queries on every rank
    items = (torch.randn((NI, c["D"]), device=dev, generator=g) * 0.3)
    q = (torch.randn((Q - Q % world, c["D"]), device=dev, generator=g) * 0.3)
    out = {}
    if world > 1:
        per = (NI + world - 1) // world
        lo, hi = rank * per, min(NI, (rank + 1) * per)
        index = tt.ShardedFlatIPIndex(items[lo:hi].contiguous(), dtype=torch.bfloat16, contiguous_offset=lo)
        del items
        index.search(q, K)
        dist.barrier(); torch.cuda.synchronize()
        best = float("inf")
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); index.search(q, K); e1.record()
            dist.barrier(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = min(best, float(t[0]))
        flops = 2.0 * q.shape[0] * NI * c["D"]
        out["bf16"] = {"queries_per_s": q.shape[0] / (best * 1e-3), "ms": best, "queries": q.shape[0], "items": NI,
                       "items_per_gpu": hi - lo, "tflops": flops / (best * 1e-3) / 1e12,
                       "frac_of_bf16_peak": flops / (best * 1e-3) / 1e12 / (pk["tf"] * world)}
        return out
    for name, (qi, it) in {"bf16": (q.bfloat16(), items.bfloat16()), "f32": (q[:1024], items)}.items():
        try:
            F.topk(qi, it, K)                          # warm-up at the timed shape (sizes the workspace)
            torch.cuda.synchronize()
            msr = float("inf")
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                F.topk(qi, it, K)
                e1.record()
                torch.cuda.synchronize()
                msr = min(msr, e0.elapsed_time(e1))
            flops = 2.0 * qi.shape[0] * NI * c["D"]
            out[name] = {"queries_per_s": qi.shape[0] / (msr * 1e-3), "ms": msr, "queries": qi.shape[0], "items": NI,
                         "tflops": flops / (msr * 1e-3) / 1e12, "frac_of_bf16_peak": flops / (msr * 1e-3) / 1e12 / pk["tf"]}
        except Exception as e:  # noqa: BLE001
            out[name] = {"error": str(e)[:200]}
    return out


if __name__ == "__main__":
    sys.exit(main())
