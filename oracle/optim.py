"""numpy fp32 restatement of the optimisers the reference builds (training.py:1311-1346).

TEST INFRASTRUCTURE (see oracle/__init__.py).  The arithmetic lives in third-party torch
(torch>=2.0,<3.0, installed 2.11.0):
  * SparseAdam : torch/optim/_functional.py:24-84, torch/optim/sparse_adam.py:63-125
  * AdamW/Adam : torch/optim/adam.py:347-547  (_single_tensor_adam, the CPU default)
  * SGD        : torch/optim/sgd.py (_single_tensor_sgd)
Pinned by tests/golden/optim.npz (torch itself run on hand-fed gradients).
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32


class OptState:
    """Per-parameter optimiser slots: {'step', 'exp_avg', 'exp_avg_sq', 'momentum_buffer'}."""

    def __init__(self) -> None:
        self.slots: dict = {}

    def slot(self, name):
        return self.slots.setdefault(name, {"step": 0})


def coalesce(idx, vals):
    """torch sparse coalesce: unique sorted row ids; duplicates summed in stable (original) order."""
    order = np.argsort(idx, kind="stable")
    sidx = idx[order]
    rows, start = np.unique(sidx, return_index=True)
    out = np.zeros((rows.size, vals.shape[1]), dtype=F32)
    seg = np.searchsorted(rows, sidx)
    svals = vals[order]
    for j in range(sidx.size):           # sequential fp32 accumulation, like coalesce_sparse_cpu
        out[seg[j]] = (out[seg[j]] + svals[j]).astype(F32)
    return rows, out


def sparse_adam_step(p, st, idx, vals, *, lr, betas=(0.9, 0.999), eps=1e-8):
    """torch.optim.SparseAdam on one table; `p` updated in place.  Returns the touched row ids."""
    b1, b2 = betas
    if "exp_avg" not in st:
        st["exp_avg"] = np.zeros_like(p)
        st["exp_avg_sq"] = np.zeros_like(p)
    st["step"] += 1
    step = st["step"]
    rows, g = coalesce(np.asarray(idx), np.asarray(vals, dtype=F32))
    if g.size == 0:
        return rows
    m_old = st["exp_avg"][rows]
    v_old = st["exp_avg_sq"][rows]
    dm = ((g - m_old) * F32(1 - b1)).astype(F32)
    st["exp_avg"][rows] = (m_old + dm).astype(F32)
    dv = ((g * g - v_old) * F32(1 - b2)).astype(F32)
    st["exp_avg_sq"][rows] = (v_old + dv).astype(F32)
    numer = (dm + m_old).astype(F32)
    denom = (np.sqrt((dv + v_old).astype(F32)) + F32(eps)).astype(F32)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    step_size = lr * math.sqrt(bc2) / bc1
    p[rows] = (p[rows] + F32(-step_size) * (numer / denom)).astype(F32)
    return rows


def adam_zero_grad_scalars(step, lr, betas=(0.9, 0.999)):
    """(step_size, sqrt(bias_correction2)) of _single_tensor_adam for a given step (Python doubles)."""
    b1, b2 = betas
    return lr / (1 - b1 ** step), (1 - b2 ** step) ** 0.5


def dense_step(kind, p, g, st, *, lr, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-8, momentum=0.0):
    """One step of AdamW / Adam / SGD on a dense tensor `p` with dense grad `g` (in place)."""
    st["step"] += 1
    step = st["step"]
    g = np.asarray(g, dtype=F32)
    if kind in ("adamw", "adam"):
        b1, b2 = betas
        if "exp_avg" not in st:
            st["exp_avg"] = np.zeros_like(p)
            st["exp_avg_sq"] = np.zeros_like(p)
        m, v = st["exp_avg"], st["exp_avg_sq"]
        if weight_decay != 0:
            if kind == "adamw":
                p *= F32(1 - lr * weight_decay)                  # adam.py: param.mul_(1 - lr * weight_decay)
            else:
                g = (g + F32(weight_decay) * p).astype(F32)      # grad.add(param, alpha=weight_decay)
        m += ((g - m) * F32(1 - b1)).astype(F32)                 # exp_avg.lerp_(grad, 1 - beta1)
        v *= F32(b2)
        v += (F32(1 - b2) * g * g).astype(F32)                   # mul_(beta2).addcmul_(grad, grad, value=1-beta2)
        step_size, bc2s = adam_zero_grad_scalars(step, lr, betas)
        denom = (np.sqrt(v) / F32(bc2s) + F32(eps)).astype(F32)
        p += (F32(-step_size) * (m / denom)).astype(F32)         # addcdiv_(exp_avg, denom, value=-step_size)
        return
    if kind == "sgd":
        if weight_decay != 0:
            g = (g + F32(weight_decay) * p).astype(F32)
        if momentum != 0:
            if "momentum_buffer" not in st:
                st["momentum_buffer"] = g.copy()
            else:
                st["momentum_buffer"] *= F32(momentum)
                st["momentum_buffer"] += g
            g = st["momentum_buffer"]
        p += (F32(-lr) * g).astype(F32)
        return
    raise ValueError(f"Unsupported optimizer: {kind}")
