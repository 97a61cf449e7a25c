"""Shared helpers for the parity tests (CPU side; no CUDA here)."""
from __future__ import annotations

from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"

TRAIN_CASES = {
    # name: kwargs for oracle.spec_from_state + train_step
    "train_gated_mlp": dict(optimizer="adamw"),
    "train_gated_mlp_cal": dict(optimizer="adamw"),
    "train_gated_mlp_cosine": dict(optimizer="adamw"),
    "train_embedding_only": dict(optimizer="adamw"),
    "train_dense_adam": dict(optimizer="adam", sparse=False),
    "train_sgd": dict(optimizer="sgd"),
    "train_linear_sum": dict(optimizer="adamw", fusion="sum"),
}


def load_case(name):
    z = np.load(GOLDEN / f"{name}.npz")
    d = {k: z[k] for k in z.files}
    NU, NI, D, H, Hg, F, B, N, steps = (int(v) for v in d["meta_dims"])
    lr, wd, b1, b2, lu, li, lc, mom = (float(v) for v in d["hyper"])
    meta = dict(NU=NU, NI=NI, D=D, H=H, Hg=Hg, F=F, B=B, N=N, steps=steps, lr=lr, wd=wd,
                betas=(b1, b2), lambdas=(lu, li, lc), momentum=mom)
    init = {k[len("init/"):]: v.copy() for k, v in d.items() if k.startswith("init/")}
    return d, meta, init


def state_after(d, step):
    pre = f"after{step}/"
    return {k[len(pre):]: v for k, v in d.items() if k.startswith(pre)}
