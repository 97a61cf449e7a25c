#!/usr/bin/env python
"""The HBM-bound row kernels alone at BASELINE config-2 shapes (item side: 2 M x 96 fp32 tables with both moments,
49 152 touched positions per step = 8192 Zipf(1.05) positives + 40 960 uniform negatives): CUDA-event timings with a
cold L2 and algorithmic bytes / time, for `ncu --set full` captures of the same launches.

    python scripts/time_rows.py [--reps 5] [--rows 2000000] [--gap 49]
"""
import argparse
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import functional as F  # noqa: E402

PEAK = 6553.6


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--rows", type=int, default=2_000_000)
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--gap", type=int, default=49)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(1)
    NI, D, B, N = a.rows, 96, a.batch, 5
    p = torch.randn((NI, D), device=dev) * 0.02
    m = torch.randn((NI, D), device=dev) * 2e-5
    v = (m / 3).square() + 1e-18
    last = torch.zeros(NI, dtype=torch.int32, device=dev)
    w = 1.0 / torch.arange(1, NI + 1, device=dev, dtype=torch.float64) ** 1.05
    pos = torch.multinomial(w, B, replacement=True, generator=gen)
    neg = torch.randint(0, NI, (B * N,), device=dev, generator=gen)
    idx = torch.cat([pos, neg]).contiguous()
    R = idx.numel()
    uniq = int(torch.unique(idx).numel())
    grad = torch.randn((R, D), device=dev) * 1e-4
    z = torch.empty((R, 2 * D), device=dev)
    step = 200
    scal = F.adam_scalar_table(step + 2, 1e-3, (0.9, 0.999), dev)
    sidx, perm = F.sort_rows(idx, NI)
    ll = F.find_long_segments(sidx)
    print(f"R={R} unique={uniq} long segments={int(ll[0])} longest={int(torch.bincount(idx).max())}", flush=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    clean = os.environ.get("TTAM_FLUSH", "clean") != "dirty"

    def flush_l2():
        # a cold L2 of CLEAN lines (a 256 MB read); TTAM_FLUSH=dirty: the memset flush, whose write-backs land in the timed kernel
        if clean:
            flush.view(torch.int64).max()
        else:
            flush.zero_()

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def alone(fn, prep=None):
        fn()
        tot = 0.0
        for _ in range(a.reps):
            if prep is not None:
                prep()
            flush_l2()
            torch.cuda._sleep(400_000)      # the host enqueues fn() while the GPU spins: e0 -> e1 is device time only
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / a.reps * 1e3

    def report(name, us, nbytes):
        gbs = nbytes / us / 1e3
        print(f"{name:34s} {us:7.1f} us  {nbytes / 1e6:7.1f} MB  {gbs:7.0f} GB/s  {gbs / PEAK:.3f} of the measured copy peak", flush=True)

    def behind():
        last[idx] = step - 1 - a.gap

    def caught_up():
        last[idx] = step - 1

    todo = {
        "gather_rows": (lambda: F.gather_rows(p, idx, out=z[:, :D]), None, R * (2 * D * 4 + 8)),
        "sparse_adam_rows": (lambda: F.sparse_adam_rows(p, m, v, sidx, perm, grad, lr=1e-3, step=step, scalars=scal, long_list=ll), None,
                             R * (D * 4 + 12) + uniq * 6 * D * 4),
        "lazy_rows": (lambda: F.lazy_rows("adamw", p, m, v, last, sidx, perm, grad, scalars=scal, lr=1e-3, weight_decay=0.01, step=step,
                                          long_list=ll), caught_up, R * (D * 4 + 12) + uniq * (6 * D * 4 + 8)),
        f"lazy_catchup (gap {a.gap})": (lambda: F.lazy_catchup("adamw", p, m, v, last, sidx, scalars=scal, lr=1e-3, weight_decay=0.01, step=step),
                                        behind, uniq * (6 * D * 4 + 8) + R * 8),
        "lazy_catchup (gap 1)": (lambda: F.lazy_catchup("adamw", p, m, v, last, sidx, scalars=scal, lr=1e-3, weight_decay=0.01, step=step),
                                 lambda: last.__setitem__(idx, step - 2), uniq * (6 * D * 4 + 8) + R * 8),
        "sort_rows": (lambda: F.sort_rows(idx, NI, sorted_idx=sidx, perm=perm), None, R * (8 + 8 + 4) * 2),
    }
    # what the memory system gives random 384-byte rows when the launch is long enough to hide ramp-up and tail: the same
    # gather over 1 M random rows (384 MB read at random + 384 MB written in order)
    big_idx = torch.randint(0, NI, (1 << 20,), device=dev, generator=gen)
    big_out = torch.empty((1 << 20, D), device=dev)
    todo["gather_rows, 1 M random rows"] = (lambda: F.gather_rows(p, big_idx, out=big_out), None, (1 << 20) * (2 * D * 4 + 8))
    for name, (fn, prep, nbytes) in todo.items():
        if a.only and a.only not in name:
            continue
        report(name, alone(fn, prep), nbytes)


if __name__ == "__main__":
    main()
