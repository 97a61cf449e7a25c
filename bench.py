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


L2_POLICY = "inputs larger than L2: each step gathers its rows from ~12 GB of tables / optimiser state / features (126 MB L2)"


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
    if c["NI"] <= 1 << 24:
        pos = torch.multinomial(pop, steps * B, replacement=True, generator=gen).view(steps, B)
    else:   # torch.multinomial stops at 2^24 categories (config 4: 50M items): inverse-CDF draw from the same distribution
        cdf = torch.cumsum(pop.double(), 0)
        u = torch.rand(steps * B, device=device, dtype=torch.float64, generator=gen) * cdf[-1]
        pos = torch.searchsorted(cdf, u).clamp_(max=c["NI"] - 1).view(steps, B)
        del cdf
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
            self.stop.wait(0.002 if self.nvml is not None else 0.5)

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
# per-kernel-class timings (roofline) and the drop-in hook measurement
# ------------------------------------------------------------------------------------------------
def kernel_function(name: str, precision: str) -> str:
    """CUDA kernel function behind an entry of the kernel table (csrc/): the grouping of the ncu launch lists."""
    n = name.lower()
    tc = precision != "fp32"
    if n.startswith("gemm") and "wgrad" in n:
        return "gemm_tf32_tma_wgrad_kernel + splitk_reduce (gemm_tma.cu)" if tc else "gemm_f32_kernel wgrad + splitk_reduce (gemm_simt.cu)"
    if n.startswith("gemm layer 1"):
        return "gemm_tf32_kernel (gemm_tc.cu)" if tc else "gemm_f32_kernel (gemm_simt.cu)"
    if n.startswith("gemm"):
        return "gemm_tf32_tma_kernel (gemm_tma.cu)" if tc else "gemm_f32_kernel (gemm_simt.cu)"
    for key, fn in (("bag_fwd", "bag_fwd_kernel (bag.cu)"), ("bag_wgrad_tc", "bag_wgrad_tc_kernel + splitk_reduce (gemm_tma.cu)"),
                    ("bag_wgrad", "bag_wgrad_kernel + reduce (bag.cu)"), ("gather_rows", "gather_rows_kernel (rows.cu)"),
                    ("gate_fwd", "gate_fwd_vec_kernel (rows.cu)"), ("gate_bwd", "gate_bwd_vec_kernel (rows.cu)"), ("loss_aug", "loss_aug_vec_kernel (rows.cu)"),
                    ("loss_fwd", "loss_vec_kernel (rows.cu)"), ("inbatch", "inbatch loss (inbatch.cu + GEMMs)"), ("sort_rows", "cub::DeviceRadixSort + find_long_segments"),
                    ("sparse_adam", "sparse_adam_rows_kernel (optim.cu)"), ("lazy_catchup", "lazy_catchup_kernel (optim.cu)"), ("lazy_rows", "lazy_rows_kernel (optim.cu)")):
        if n.startswith(key):
            return fn
    return name


def time_kernel_classes(eng, F, c, dev, users, pos, neg, user_x, item_x, nu_l, ni_l, precision, pk):
    """Each kernel class of the item side of one step (49 152 rows at B = 8192: 6/7 of the step's rows), launched alone on
    the engine's own buffers.  Algorithmic bytes = operands read once + results written once."""
    B, N, D, Fd, H = c["B"], c["N"], c["D"], c["F"], c["H"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    clean = os.environ.get("TTAM_FLUSH", "clean") != "dirty"
    def flush_l2():
        # a cold L2 that holds CLEAN lines: reading 256 MB evicts what the kernel left behind without leaving 126 MB of dirty
        # lines whose write-backs the next kernel would pay for (a memset flush showed up as 90-130 MB of DRAM writes inside
        # an 18 us kernel under ncu).  TTAM_FLUSH=dirty restores the memset.
        if clean:
            flush.view(torch.int64).max()
        else:
            flush.zero_()

    def alone(fn, reps=5, prep=None):
        fn()
        tot = 0.0
        for _ in range(reps):
            if prep is not None:
                prep()
            flush_l2()
            torch.cuda._sleep(400_000)      # the host enqueues fn() while the GPU spins: k0 -> k1 is device time only
            k0.record(); fn(); k1.record()
            torch.cuda.synchronize()
            tot += k0.elapsed_time(k1)
        return tot / reps

    inbatch = getattr(eng, "loss_kind", "sampled") == "inbatch"
    idx = ((pos if inbatch else torch.cat([pos, neg.reshape(-1)])) % ni_l).contiguous()
    R = idx.numel()
    out = []

    def add(name, fn, nbytes, flops=0.0, prep=None, **extra):
        try:
            ms = alone(fn, prep=prep)
        except Exception as e:  # noqa: BLE001
            out.append({"kernel": name, "error": str(e)[:200], "ms": 0.0, "achieved": 0.0, "frac": 0.0, "bytes": int(nbytes), "flops": flops})
            return
        gbs = nbytes / (ms * 1e-3) / 1e9
        out.append({"kernel": name, "ms": ms, "bytes": int(nbytes), "flops": flops, "achieved": gbs, "frac": gbs / pk["hbm"],
                    "tflops": flops / (ms * 1e-3) / 1e12, **extra})

    plan, T = eng.item, eng.tables
    if not plan.fe_layers:
        rows = torch.empty((R, D), device=dev)
        add("gather_rows: E_item[idx]", lambda: F.gather_rows(plan.table, idx, out=rows), R * (2 * D * 4 + 8))
        return out
    (W1, b1), (W2, b2) = plan.fe_layers
    G1, c1, G2, c2 = plan.gate
    Hg = G1.shape[0]
    tc = precision != "fp32"
    f32 = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
    z, hd, a, pre2, g, t, o, q = f32(R, 2 * D), f32(R, H), f32(R, Hg), f32(R, D), f32(R, D), f32(R, D), f32(R, D), f32(R, D)
    dt = torch.randn((R, D), device=dev) * 1e-4
    dpre2, dz, dpre1, dhd = f32(R, D), f32(R, 2 * D), f32(R, Hg), f32(R, H)
    Xi = eng._x(item_x)
    bag = getattr(Xi, "_ttam_bag", None)
    if bag is not None:
        nnz = int(bag.rowptr[-1]) / bag.shape[0]
        row_bytes = 8 + 16 + nnz * 8 + bag.T * 4
        # SURVEY 8(d) "feature gather, bag form": per row nnz*(4+4) + tail*4 bytes of CSR + nnz*H*4 bytes of W1^T rows (those
        # are served from shared memory / L2, not HBM: the figure is the survey's algorithmic one) + the output row
        w_rows = (nnz + bag.T) * H * 4
        # `bytes` = what has to cross HBM (index + CSR row + dense tail in, the hidden row out); the W1^T rows a row touches
        # (nnz * H * 4 B, SURVEY 8(d)'s second term) are served from shared memory and are reported beside it
        add("bag_fwd: layer 1 of the item tower from CSR rows (b1 + sum_j x_j W1[:, j], relu)",
            lambda: F.bag_linear_fwd(bag, idx, W1, b1, act="relu", out=hd, round_tf32_out=tc),
            R * (row_bytes + H * 4), 2.0 * R * (nnz + bag.T) * H, smem_operand_bytes=int(R * w_rows))
    else:
        W1p = F.round_tf32_(F.pad_cols(W1, always_copy=True)) if tc else W1
        add("gemm layer 1 fwd: X[idx] . W1^T + b1, relu",
            lambda: F.linear_fwd(Xi, W1p, b1, gather=idx, act="relu", out=hd, precision=precision, x_rounded=tc, w_rounded=tc),
            R * (Fd * 4 + 8 + H * 4) + H * Fd * 4, 2.0 * R * Fd * H)
    F.gather_rows(plan.table, idx, out=z[:, :D])
    add("gather_rows: E_item[idx] -> z[:, :D]", lambda: F.gather_rows(plan.table, idx, out=z[:, :D]), R * (2 * D * 4 + 8))
    # the GEMMs exactly as ttam_tower_fwd / ttam_tower_bwd issue them (csrc/tower.cu): on the tensor-core path against the
    # per-step TF32-rounded (forward) and rounded + transposed (data gradient) weight copies, i.e. the TMA-fed kernel
    if tc:
        (W2r, G1r, G2r), (W2rT, G1rT, G2rT) = F.prepare_weights([W2, G1, G2])
        fw = lambda x, w, bias, out, **kw: F.linear_fwd(x, w, bias, out=out, precision=precision, w_rounded=True, **kw)
        dg = lambda dy, wT, out, **kw: F.linear_dgrad(dy, wT, out=out, precision=precision, w_transposed=True, **kw)
        W2f, G1f, G2f, W2b, G1b, G2b = W2r, G1r, G2r, W2rT, G1rT, G2rT
    else:
        fw = lambda x, w, bias, out, x_rounded=False, out_rounded=False, **kw: F.linear_fwd(x, w, bias, out=out, precision=precision, **kw)
        dg = lambda dy, w, out, x_rounded=False, out_rounded=False, **kw: F.linear_dgrad(dy, w, out=out, precision=precision, **kw)
        W2f, G1f, G2f, W2b, G1b, G2b = W2, G1, G2, W2, G1, G2
    add("gemm layer 2 fwd: f = h . W2^T + b2", lambda: fw(hd, W2f, b2, z[:, D:], x_rounded=tc and bag is not None),
        R * (H + D) * 4 + D * H * 4, 2.0 * R * H * D)
    add("gemm gate 1 fwd: a = relu([e;f] . G1^T + c1)", lambda: fw(z, G1f, c1, a, act="relu", out_rounded=tc),
        R * (2 * D + Hg) * 4 + Hg * 2 * D * 4, 2.0 * R * 2 * D * Hg)
    add("gemm gate 2 fwd: pre2 = a . G2^T + c2", lambda: fw(a, G2f, c2, pre2, x_rounded=tc),
        R * (Hg + D) * 4 + D * Hg * 4, 2.0 * R * Hg * D)
    add("gate_fwd: sigmoid, blend, + A_item[idx]", lambda: F.gate_fwd(z, pre2, aug_table=plan.aug, idx=idx, g=g, t=t, o=o, q=q),
        R * (2 * D + D + D + 4 * D) * 4 + R * 8)
    ou, tu, qu = f32(B, D).normal_(), f32(B, D).normal_(), f32(B, D).normal_()
    if inbatch:
        add("inbatch_loss_fwd_bwd: S = o_u o_p^T, row softmax, dS . o_p, dS^T . o_u, mimic",
            lambda: F.inbatch_loss_fwd_bwd(ou, o, t_u=tu, t_p=t, q_u=qu, q_p=q, lambda_u=0.15, lambda_i=0.15, precision=precision),
            4 * B * B * 4 + 12 * B * D * 4, 3 * 2.0 * B * B * D)
    elif getattr(eng, "_can_fuse_aug_loss", lambda n: False)(N) and eng.user.aug is not None:
        # what the one-GPU step launches: the augmentation add folded into the loss (o = t + A[idx] in registers)
        uidx = (users % nu_l).contiguous()
        add("loss_aug_fwd_bwd: o = t + A[idx] in registers, dots, BCE, mimic MSEs, all gradients",
            lambda: F.loss_aug_fwd_bwd(tu, t, eng.user.aug, plan.aug, uidx, idx, mimic=True, lambda_u=0.15, lambda_i=0.15),
            (R + B) * (2 * D * 4 + 8) + (R + 3 * B) * D * 4)
    else:
        add("loss_fwd_bwd: dots, BCE, mimic MSEs, all gradients", lambda: F.loss_fwd_bwd(ou, o, t_u=tu, t_p=t[:B], q_u=qu, q_p=q[:B], lambda_u=0.15, lambda_i=0.15),
            2 * (R + 4 * B) * D * 4)
    add("gate_bwd", lambda: F.gate_bwd(dt, z, g, dpre2=dpre2, dz=dz), R * (D + 2 * D + D + D + 2 * D) * 4)
    add("gemm gate 2 dgrad (relu mask)", lambda: dg(dpre2, G2b, dpre1, aux=a, relu_mask=True, out_rounded=tc),
        R * (D + 2 * Hg) * 4, 2.0 * R * Hg * D)
    add("gemm gate 1 dgrad (accumulate into dz)", lambda: dg(dpre1, G1b, dz, accumulate=True, x_rounded=tc),
        R * (Hg + 4 * D) * 4, 2.0 * R * 2 * D * Hg)
    add("gemm layer 2 dgrad (relu mask)", lambda: dg(dz[:, D:], W2b, dhd, aux=hd, relu_mask=True),
        R * (D + 2 * H) * 4, 2.0 * R * H * D)
    add("gemm gate 2 wgrad", lambda: F.linear_wgrad(dpre2, a, precision=precision, x_rounded=tc), R * (D + Hg) * 4, 2.0 * R * Hg * D)
    add("gemm gate 1 wgrad", lambda: F.linear_wgrad(dpre1, z, precision=precision), R * (Hg + 2 * D) * 4, 2.0 * R * 2 * D * Hg)
    add("gemm layer 2 wgrad", lambda: F.linear_wgrad(dz[:, D:], hd, precision=precision, x_rounded=tc and bag is not None), R * (D + H) * 4, 2.0 * R * H * D)
    if bag is not None:
        if tc:   # what ttam_tower_bwd launches on the tensor-core path: CSR rows expanded into the tcgen05 operand tile (+ split-K reduce)
            add("bag_wgrad_tc: layer 1 weight gradient, CSR rows expanded into the MMA operand tile + split-K reduce",
                lambda: F.bag_linear_wgrad(bag, idx, dhd, precision="tf32"), R * (row_bytes + H * 4), 2.0 * R * Fd * H)
        else:
            add("bag_wgrad: layer 1 weight gradient (deterministic column-owner scatter)", lambda: F.bag_linear_wgrad(bag, idx, dhd),
                R * (row_bytes + H * 4), 2.0 * R * (nnz + bag.T) * H, smem_operand_bytes=int(R * w_rows))
    else:
        add("gemm layer 1 wgrad: dh^T . X[idx]", lambda: F.linear_wgrad(dhd, Xi, gather=idx, precision=precision, x_rounded=tc),
            R * (Fd * 4 + 8 + H * 4) + H * Fd * 4, 2.0 * R * Fd * H)
    srt = eng._sort(idx, "bench", ni_l)
    add("sort_rows + find_long_segments (touched item rows)", lambda: eng._sort(idx, "bench", ni_l), R * (8 + 8 + 4) * 2)
    uniq = int(torch.unique(idx).numel())
    e_tab, a_tab = T["item_encoder.embedding.weight"], T.get("adaptive_mimic.item_augmented.weight")
    if e_tab.mode == "sparse_adam":
        add("sparse_adam_rows: segment-reduce + SparseAdam on E_item", lambda: eng._update_table(e_tab, srt, dz[:, :D]),
            R * (D * 4 + 12) + uniq * 6 * D * 4)
    if a_tab is not None:
        # a caught-up row is skipped, so every repetition starts from saved stamps.  (1) the stamps the NEXT step of this run would
        # find - the work the timed steps actually did; (2) every touched row 49 steps behind with non-zero moments - what a long
        # run settles at for this workload (an item row is touched every ~49 steps): arithmetic-bound replay, shown for the record
        stamps = a_tab.last_step.clone()
        add("lazy_catchup: zero-gradient replay of the touched A_item rows (gaps as the next step of this run finds them)",
            lambda: eng._catchup(a_tab, srt[0]), uniq * (6 * D * 4 + 8) + R * 8, prep=lambda: a_tab.last_step.copy_(stamps))
        gap = max(1, min(49, int(eng.t) - 1))
        def behind():
            a_tab.last_step.copy_(stamps)
            a_tab.last_step[idx] = max(0, int(eng.t) - 1 - gap)
        add(f"lazy_catchup, steady-state probe: every touched row {gap} steps behind", lambda: eng._catchup(a_tab, srt[0]),
            uniq * (6 * D * 4 + 8) + R * 8, prep=behind, probe=True)
        a_tab.last_step.copy_(stamps)
        eng._catchup(a_tab, srt[0])
        add("lazy_rows: segment-reduce + lazy-exact AdamW on A_item", lambda: eng._update_table(a_tab, srt, dt), R * (D * 4 + 12) + uniq * (6 * D * 4 + 8))
    return out


class _Interactions(torch.utils.data.Dataset):
    """(user_idx, item_idx) pairs held as two tensors, like the reference's InteractionDataset (datasets.py:12-45)."""

    def __init__(self, users, items):
        self._users, self._items = users.cpu(), items.cpu()

    def __len__(self):
        return self._users.shape[0]

    def __getitem__(self, i):
        return self._users[i], self._items[i]


def bench_hook(tt, eng, model, c, K, users, pos, user_x, item_x, dev, precision):
    from torch import nn
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import hooks
    hooks.OPTIONS.precision, hooks.OPTIONS.graph, hooks.OPTIONS.sampler, hooks.OPTIONS.stats = precision, True, "device", None
    object.__setattr__(model, hooks._ENGINE_ATTR, eng)
    sparse = [p for n, p in model.named_parameters() if n.endswith("encoder.embedding.weight") and eng.tables[n].mode == "sparse_adam"]
    dense = [p for p in model.parameters() if all(p is not q for q in sparse)]
    opts = [torch.optim.AdamW(dense, lr=c["lr"], weight_decay=c["wd"])] + ([torch.optim.SparseAdam(sparse, lr=c["lr"])] if sparse else [])
    loader = torch.utils.data.DataLoader(_Interactions(users, pos), batch_size=c["B"], shuffle=True)
    kw = dict(optimizers=opts, criterion=nn.BCEWithLogitsLoss(), negatives_per_positive=c["N"], num_items=item_x.shape[0],
              user_positive_items={}, user_features=user_x, item_features=item_x, device=dev, gradient_clip_norm=None,
              loss_weights={"mimic_user": c["lambdas"][0], "mimic_item": c["lambdas"][1]} if eng.mimic else {},
              item_category_tensor=None, major_category_id=None)
    hooks._train_one_epoch(model, loader, **kw)              # warm-up epoch (graph capture for this batch size)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    loss = hooks._train_one_epoch(model, loader, **kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"samples_per_s": users.numel() / dt, "steps": K, "seconds": dt, "epoch_mean_loss": loss,
            "what": "hooks._train_one_epoch over a DataLoader of K*B interactions (device sampler, device batch iterator, CUDA-graph replay; "
                    "includes the epoch-end flush of the lazily-updated tables and the host read of the losses)"}


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
    ap.add_argument("--no-fp32", action="store_true", help="skip the fp32-GEMM re-timing of the step")
    ap.add_argument("--no-hook", action="store_true", help="skip the hooks._train_one_epoch throughput measurement")
    ap.add_argument("--batch", type=int, default=None, help="samples per GPU per step (default 8192; sweep: 4096..65536)")
    ap.add_argument("--mode", default="hybrid", choices=["hybrid", "sparse", "dense"],
                    help="optimiser sweep (BASELINE configs[4]): hybrid = AdamW + SparseAdam (reference default), "
                         "sparse = embedding-only towers, mimic off, all SparseAdam, dense = sparse:false (AdamW semantics on every table)")
    ap.add_argument("--loss", default="sampled", choices=["sampled", "inbatch"],
                    help="sampled = the reference's loss (BCE over 1 positive + 5 sampled negatives, training.py:770-803; the headline); "
                         "inbatch = in-batch softmax over the batch's positives (BASELINE configs[1] wording; an extension the reference "
                         "does not have - a separately labelled line)")
    ap.add_argument("--no-graph", action="store_true", help="N=1: launch the step eagerly instead of replaying a CUDA graph (diagnostic)")
    ap.add_argument("--route", default="peer", choices=["static", "peer", "dynamic"],
                    help="N>1: static = fixed-capacity slots, the whole sharded step replays as CUDA graphs; peer = static, and the "
                         "row payloads travel by NVLink peer loads/stores instead of NCCL all-to-alls; dynamic = per-step "
                         "split sizes, eager launches (diagnostic)")
    ap.add_argument("--small", action="store_true", help="1/16-size tables (debugging only; not a valid bench line)")
    ap.add_argument("--config", type=int, default=2, choices=[2, 4],
                    help="2 = BASELINE configs[1] (1M x 2M, D=96; the headline), 4 = configs[3] (10M users x 50M items, D=256, "
                         "H=512: row-sharded over 8 GPUs, ~70 GB per GPU; a separately labelled record, run with --gpus 8)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    c = dict(CFG)
    if args.config == 4:
        c.update(NU=10_000_000, NI=50_000_000, D=256, H=512, Hg=256)
    if args.small:
        c.update(NU=c["NU"] // 16, NI=c["NI"] // 16)
    if args.batch:
        c["B"] = int(args.batch)
    W = max(args.warmup, 3)
    K = args.steps
    workload = (f"synthetic {c['NU']} users x {c['NI']} items, D={c['D']}, F={c['F']}->H={c['H']}->D MLPs, gated fusion, "
                f"adaptive mimic, dropout=0, B={c['B']}, {c['N']} sampled negatives, AdamW+SparseAdam")
    if args.mode == "sparse":
        workload = (f"synthetic {c['NU']} users x {c['NI']} items, D={c['D']}, embedding-only towers, no mimic, B={c['B']}, "
                    f"{c['N']} sampled negatives, all-sparse SparseAdam")
    elif args.mode == "dense":
        workload = workload.replace("AdamW+SparseAdam", "all-dense AdamW (sparse:false, lazy-exact rows)")
    if args.loss == "inbatch":
        workload = workload.replace(f"{c['N']} sampled negatives", "IN-BATCH SOFTMAX over the batch's positives (extension: not a reference loss)")

    if args.impl == "reference":
        if rank != 0:
            return 0
        # same K and W as the GPU arm (a CPU step of the full workload is ~0.5 s: the default 50 + 5 steps take ~30 s after ~1 min
        # of table / feature setup); only an extreme K is bounded so that the run still ends within a few minutes
        ksteps, wsteps = min(K, 200), min(W, 20)
        r = cpu_arm(c, ksteps, wsteps)
        line = {"impl": "reference", "metric": "train samples/sec", "value": r["value"], "unit": "samples/s",
                "n_gpus": args.gpus, "steps": ksteps, "warmup": wsteps, "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload, "l2_policy": L2_POLICY},
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
                         max_steps=4 * (K + W) + 64, loss=args.loss)
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
    clk = ClockSampler(local)
    clk.__enter__()                     # samples SM clocks / throttle reasons across both timed regions
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
    clk.__exit__()
    ms_e2e = f0.elapsed_time(f1)
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    B, N, D, Fd, H = c["B"], c["N"], c["D"], c["F"], c["H"]

    # ---- roofline: every kernel class of the (item-side) step timed alone with CUDA events on the launching stream, cold L2
    # (256 MB flush before each repetition); the class with the largest time is the line's `roofline`, all of them are
    # listed under roofline.kernels
    pk = peaks()
    kernels = []
    try:
        kernels = time_kernel_classes(eng, F, c, dev, users[W], pos[W], neg[W], user_x, item_x, nu_l, ni_l, args.precision, pk)
    except Exception as e:  # noqa: BLE001 - the headline numbers above do not depend on this table
        kernels = [{"kernel": "error", "error": str(e)[:300], "ms": 0.0, "achieved": 0.0, "frac": 0.0, "bytes": 0, "flops": 0.0}]
    # The dominant kernel = the CUDA kernel (function) with the largest summed time over its launches in the step - the same
    # grouping as the ncu launch list under profiles/ (share per kernel function).  achieved = sum of algorithmic bytes /
    # sum of launch durations = bytes per launch / average launch duration.  Probes are not part of the timed step.
    groups = {}
    for k in kernels:
        if k.get("probe") or not k.get("ms"):
            continue
        g = groups.setdefault(kernel_function(k["kernel"], args.precision), {"ms": 0.0, "bytes": 0, "flops": 0.0, "launches": 0, "members": []})
        g["ms"] += k["ms"]; g["bytes"] += k["bytes"]; g["flops"] += k["flops"]; g["launches"] += 1; g["members"].append(k["kernel"])
    if groups:
        fn, g = max(groups.items(), key=lambda kv: kv[1]["ms"])
        top = {"kernel": fn, "ms": g["ms"] / g["launches"], "bytes": g["bytes"] / g["launches"], "flops": g["flops"] / g["launches"],
               "achieved": g["bytes"] / (g["ms"] * 1e-3) / 1e9, "launches": g["launches"], "members": g["members"], "class_ms": g["ms"]}
        top["frac"] = top["achieved"] / pk["hbm"]
    else:
        top = dict(kernels[0], launches=1, members=[kernels[0]["kernel"]], class_ms=kernels[0]["ms"])
    traffic = None
    tfile = ROOT / "profiles" / "roofline_traffic.json"
    if tfile.exists():
        # the committed capture is of the default workload (config 2, B = 8192): other shapes have no traffic figure
        traffic = json.loads(tfile.read_text()).get(f"{top['kernel']}|{args.precision}") if (args.config == 2 and not args.batch and not args.small) else None
    roof = {"bound": "hbm", "achieved": top["achieved"], "peak": pk["hbm"], "unit": "GB/s", "frac": top["frac"], "traffic": traffic,
            "kernel": top["kernel"], "kernel_ms": top["ms"], "launches_per_step_item_side": top["launches"], "class_ms": top["class_ms"],
            "members": top["members"], "peak_source": pk["source"], "algorithmic_bytes": top["bytes"],
            "algorithmic_flops": top["flops"], "tensor_tflops": top["flops"] / (top["ms"] * 1e-3) / 1e12 if top["ms"] else 0.0,
            "kernels": kernels, "kernels_total_ms": sum(k["ms"] for k in kernels if not k.get("probe")),
            "l2_flush": "256 MB read before every repetition (cold L2, clean lines)"}

    # ---- the same step with fp32 SIMT GEMMs (the reference's arithmetic), printed next to the TF32 headline
    fp32_line = None
    if world == 1 and args.precision == "tf32" and not args.no_fp32:
        try:
            eng.precision = "fp32"
            eng._graphs.clear()
            for s in range(3):
                step(users[s], pos[s], neg[s])
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k32 = min(K, 20)
            g0.record()
            for s in range(W, W + k32):
                step(users[s], pos[s], neg[s])
            g1.record()
            barrier()
            ms32 = g0.elapsed_time(g1)
            fp32_line = {"value": k32 * B / (ms32 * 1e-3), "unit": "samples/s", "ms_per_step": ms32 / k32, "steps": k32,
                         "dtype": "f32", "note": "tower GEMMs as fp32 FFMA (gemm_f32_kernel); everything else identical"}
        finally:
            eng.precision = args.precision
            eng._graphs.clear()

    # ---- through the drop-in hook: hooks._train_one_epoch (the function scripts/train_b200.py binds over the reference's
    # training.py:700-833) fed by a DataLoader over K*B interactions, device sampler + device batch iterator + graph replay
    hook_line = None
    if world == 1 and not args.no_hook and args.loss == "sampled":
        try:
            hook_line = bench_hook(tt, eng, model, c, K, users[W:W + K].reshape(-1), pos[W:W + K].reshape(-1), user_x, item_x, dev, args.precision)
        except Exception as e:  # noqa: BLE001
            hook_line = {"error": str(e)[:300]}

    steps_launched = 2 * K + W
    line = {"metric": "train samples/sec", "value": world * K * B / (ms * 1e-3), "unit": "samples/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"fp32": "f32", "tf32": "tf32"}[args.precision], "data": "synthetic",
            "config": {"workload": workload, "l2_policy": L2_POLICY},
            "run": {"parallelism": ("1 gpu, CUDA-graph replay" if not args.no_graph else "1 gpu, eager launches") if world == 1 else
                       f"{world} gpus: tables/features row-sharded, batch data-parallel ({B} samples per gpu), 3 all-to-all + 1 all-reduce per step, "
                       + (f"fixed-capacity slots ({sh.last_exchange_rows[0] // world}/{sh.last_exchange_rows[1] // world} user/item rows per "
                          f"rank pair), CUDA-graph replay incl. collectives, {sh.fallback_steps} dynamic-route fallback steps"
                          + (", row payloads by NVLink peer loads/stores fused into the un-bucket/re-bucket kernels (2 device barriers per step)"
                             if sh.peer else ", row payloads by equal-split NCCL all-to-alls")
                          if args.route != "dynamic" else "per-step split sizes, eager launches")},
            "e2e": {"value": world * K * B / (ms_e2e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": B * 8 * (2 + N),
                    "d2h_bytes_per_step": 16},
            "gpu_launches": None, "clocks": clk.summary(), "roofline": roof}
    if fp32_line is not None:
        line["fp32"] = fp32_line
    if hook_line is not None:
        line["hook"] = hook_line
        if "samples_per_s" in hook_line:
            hook_line["fraction_of_e2e"] = hook_line["samples_per_s"] / line["e2e"]["value"]
    if world == 1:
        line["gpu_launches"] = int(getattr(eng, "launches_per_step", 0)) * K
    else:
        line["gpu_launches"] = (int(sh.launches_per_step) * K if args.route != "dynamic" and not args.no_graph else
                                int((F.lib().ttam_launch_count() - launches0) * K / steps_launched))

    if not args.no_retrieval:
        # config 4 searches its own corpus (50M x 256, item-sharded); config 2 the 2M x 96 corpus of BASELINE configs[2]
        r = bench_retrieval(tt, c, dev, pk, NI=c["NI"] if args.config == 4 else 2_000_000, world=world, rank=rank)
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
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import sharding as S
    import torch.distributed as dist
    g = torch.Generator(device=dev).manual_seed(7)              # same queries on every rank
    q = (torch.randn((Q - Q % world, c["D"]), device=dev, generator=g) * 0.3)
    out = {}
    if world > 1:
        # W = G x S grid (ShardedFlatIPIndex(query_groups=G)): rank g*S+s holds item shard s of S and scores query group g of
        # G.  G = 1 is the named config (NI/W items per GPU); larger G trades G copies of the corpus for G times fewer
        # queries per rank (the per-query cost of a stream does not shrink with the shard: DESIGN 4.3).  Every G that
        # divides W and whose shard fits is timed; the fastest is the `bf16` line, all of them are listed under `grid`.
        flops = 2.0 * q.shape[0] * NI * c["D"]
        grid = []
        for G in (1, 2, 4, 8):
            if world % G or (NI * G // world) * c["D"] * 6 > 40e9:
                continue
            _, shard, n_sh = S.retrieval_grid(rank, world, G)
            per = (NI + n_sh - 1) // n_sh
            lo, hi = shard * per, min(NI, (shard + 1) * per)
            gs = torch.Generator(device=dev).manual_seed(70 + shard)  # the ranks that hold shard s draw the same rows
            index = tt.ShardedFlatIPIndex(torch.randn((hi - lo, c["D"]), device=dev, generator=gs) * 0.3, dtype=torch.bfloat16,
                                          contiguous_offset=lo, query_groups=G)
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
            grid.append({"query_groups": G, "item_shards": n_sh, "items_per_gpu": hi - lo, "ms": best,
                         "queries_per_s": q.shape[0] / (best * 1e-3)})
            del index
            torch.cuda.empty_cache()
        top = min(grid, key=lambda r: r["ms"])
        out["bf16"] = {"queries_per_s": top["queries_per_s"], "ms": top["ms"], "queries": q.shape[0], "items": NI,
                       "items_per_gpu": top["items_per_gpu"], "query_groups": top["query_groups"],
                       "tflops": flops / (top["ms"] * 1e-3) / 1e12,
                       "frac_of_bf16_peak": flops / (top["ms"] * 1e-3) / 1e12 / (pk["tf"] * world), "grid": grid}
        return out
    items = (torch.randn((NI, c["D"]), device=dev, generator=g) * 0.3)
    # bf16: BASELINE configs[2] (bf16 corpus).  f32: the fp32 index the drop-in evaluation hooks build (FlatIPIndex default):
    # candidate pass on the tensor cores over a 3-way bf16 split (3x the MMA work), canonical fp32 re-score - bit-identical
    # to f32_simt, the SIMT kernel that deeper (k > 128) and paged searches use.
    items_split = F.split_bf16x3(items, item_layout=True)
    paths = {"bf16": (q.bfloat16(), items.bfloat16(), lambda a, b: F.topk(a, b, K), 1.0),
             "f32": (q, items, lambda a, b: F.topk_f32_tc(a, b, items_split, K), 3.0),
             "f32_simt": (q[:4096], items, lambda a, b: F.topk(a, b, K), 1.0)}
    for name, (qi, it, fn, mma_mult) in paths.items():
        try:
            fn(qi, it)                                 # warm-up at the timed shape (sizes the workspace)
            torch.cuda.synchronize()
            msr = float("inf")
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn(qi, it)
                e1.record()
                torch.cuda.synchronize()
                msr = min(msr, e0.elapsed_time(e1))
            flops = 2.0 * qi.shape[0] * NI * c["D"]
            tf = flops / (msr * 1e-3) / 1e12
            out[name] = {"queries_per_s": qi.shape[0] / (msr * 1e-3), "ms": msr, "queries": qi.shape[0], "items": NI,
                         "tflops": tf, "frac_of_bf16_peak": tf / pk["tf"],
                         "roofline": {"bound": "tensor", "achieved": tf * mma_mult, "peak": pk["tf"], "unit": "TFLOP/s",
                                      "frac": tf * mma_mult / pk["tf"],
                                      "traffic": None, "peak_source": pk["source"] + " (sustained bf16)",
                                      "algorithmic_flops": flops, "tensor_core_flops": flops * mma_mult,
                                      "kernel": "score_topk_kernel + finalize (whole ttam_topk call)"}}
            if name == "f32_simt":
                out[name]["roofline"].update(bound="fp32 SIMT (canonical mul + add per product, no FMA)",
                                             kernel="gemm_f32_kernel + select_topk_kernel per 4096-item chunk")
            # end to end through the public API: HOST queries (pinned) -> H2D -> search -> ids + scores D2H
            hq = qi.cpu().pin_memory()
            dq = torch.empty_like(qi)
            h_ids = torch.empty((qi.shape[0], K), dtype=torch.int64).pin_memory()
            h_sc = torch.empty((qi.shape[0], K), dtype=torch.float32).pin_memory()
            best = float("inf")
            for _ in range(2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                dq.copy_(hq, non_blocking=True)
                ids, sc = fn(dq, it)
                h_ids.copy_(ids, non_blocking=True); h_sc.copy_(sc, non_blocking=True)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            out[name]["e2e"] = {"value": qi.shape[0] / (best * 1e-3), "unit": "queries/s", "ms": best,
                                "h2d_bytes": hq.numel() * hq.element_size(), "d2h_bytes": h_ids.numel() * 8 + h_sc.numel() * 4}
        except Exception as e:  # noqa: BLE001
            out[name] = {"error": str(e)[:200]}
    return out


if __name__ == "__main__":
    sys.exit(main())
