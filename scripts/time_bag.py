#!/usr/bin/env python
"""Bag-form layer 1 alone at BASELINE config-2 shapes (item side: 49 152 rows, ~3.5 sparse + 5 dense non-zeros of 605;
user side: 8 192 rows, ~30 + 5): CUDA-event timings with a cold L2, for `ncu` captures of bag_fwd_kernel / bag_wgrad_kernel.

    python scripts/time_bag.py [--reps 5] [--side item|user|both]
"""
import argparse
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bench import make_features  # noqa: E402
from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import functional as F  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--side", default="both")
    ap.add_argument("--items", type=int, default=400_000)
    a = ap.parse_args()
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(1)
    H, Fd = 192, 605
    user_x, item_x = make_features(a.items, a.items // 2, Fd, 300, 300, dev, gen)
    W = torch.randn((H, Fd), device=dev) * 0.05
    b = torch.randn(H, device=dev) * 0.1
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    clean = os.environ.get("TTAM_FLUSH", "clean") != "dirty"

    def flush_l2():
        # a cold L2 of CLEAN lines (a 256 MB read); TTAM_FLUSH=dirty: the memset flush, whose write-backs land in the timed kernel
        if clean:
            flush.view(torch.int64).max()
        else:
            flush.zero_()

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def alone(fn):
        fn()
        tot = 0.0
        for _ in range(a.reps):
            flush_l2()
            torch.cuda._sleep(400_000)      # the host enqueues fn() while the GPU spins: e0 -> e1 is device time only
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / a.reps * 1e3

    for side, X, R in (("item", item_x, 49152), ("user", user_x, 8192)):
        if a.side not in ("both", side):
            continue
        bag = F.BagMatrix.build(X)
        idx = torch.randint(0, X.shape[0], (R,), device=dev, generator=gen)
        hd = torch.empty((R, H), device=dev)
        dh = torch.randn((R, H), device=dev) * 1e-3
        dw, db = torch.empty((H, Fd), device=dev), torch.empty(H, device=dev)
        t_f = alone(lambda: F.bag_linear_fwd(bag, idx, W, b, act="relu", out=hd, round_tf32_out=True))
        t_w = alone(lambda: F.bag_linear_wgrad(bag, idx, dh, dw=dw, db=db))
        t_tc = alone(lambda: F.bag_linear_wgrad(bag, idx, dh, dw=dw, db=db, precision="tf32"))
        print(f"{side}: R={R} mean nnz {bag.mean_nnz:.1f} (max sparse {bag.max_nnz}, tail {bag.T})  fwd {t_f:.1f} us  wgrad {t_w:.1f} us  "
              f"wgrad on the tensor cores {t_tc:.1f} us", flush=True)


if __name__ == "__main__":
    main()
