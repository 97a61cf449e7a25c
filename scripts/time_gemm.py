#!/usr/bin/env python
"""The tower-chain GEMMs at BASELINE config-2 shapes (49 152 item rows), cp.async kernel (gemm_tc.cu) against the TMA-fed
persistent kernel (gemm_tma.cu): CUDA events, cold L2.

    python scripts/time_gemm.py [--reps 5]
"""
import argparse
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import functional as F  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--rows", type=int, default=49152)
    a = ap.parse_args()
    dev = torch.device("cuda")
    R, D, H = a.rows, 96, 192
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

    for name, (M, N, K) in {"layer 2 fwd  [R,192]x[96,192]^T": (R, D, H), "gate 1 fwd   [R,192]x[96,192]^T": (R, D, 2 * D),
                            "gate 2 fwd   [R,96]x[96,96]^T": (R, D, D), "gate 2 dgrad [R,96]x[96,96]": (R, D, D),
                            "gate 1 dgrad [R,96]x[96,192]": (R, D, 2 * D), "layer 2 dgrad[R,96]x[96,192]": (R, D, H)}.items():
        dgrad = "dgrad" in name
        x = torch.randn((M, N if dgrad else K), device=dev)
        W = torch.randn((N, K), device=dev) * 0.05
        b = torch.randn(N, device=dev)
        (Wr,), (WrT,) = F.prepare_weights([W])
        if dgrad:
            out = torch.empty((M, K), device=dev)
            aux = torch.randn((M, K), device=dev)
            t0 = alone(lambda: F.linear_dgrad(x, W, out=out, aux=aux, relu_mask=True, precision="tf32"))
            t1 = alone(lambda: F.linear_dgrad(x, WrT, out=out, aux=aux, relu_mask=True, precision="tf32", w_transposed=True))
            nbytes = M * (N + 2 * K) * 4
        else:
            out = torch.empty((M, N), device=dev)
            t0 = alone(lambda: F.linear_fwd(x, W, b, act="relu", out=out, precision="tf32"))
            t1 = alone(lambda: F.linear_fwd(x, Wr, b, act="relu", out=out, precision="tf32", w_rounded=True))
            xr = F.round_tf32_(x.clone())
            t2 = alone(lambda: F.linear_fwd(xr, Wr, b, act="relu", out=out, precision="tf32", w_rounded=True, x_rounded=True))
            nbytes = M * (N + K) * 4
        extra = "" if dgrad else f"   pre-rounded A {t2:6.1f} us ({nbytes / t2 / 1e3:5.0f} GB/s)"
        print(f"{name:36s} cp.async {t0:6.1f} us ({nbytes / t0 / 1e3:5.0f} GB/s)   TMA {t1:6.1f} us ({nbytes / t1 / 1e3:5.0f} GB/s){extra}", flush=True)


if __name__ == "__main__":
    main()
