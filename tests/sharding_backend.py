"""An oracle-backed stand-in for FusedEngine's three phases, so that the N-rank exchange logic of
`ShardedEngine` / `sharding.Exchange` can be exercised on CPU with gloo (tests/test_sharding.py).

Every rank holds: its rows of the four tables (+ numpy optimiser state), a replica of the dense weights.
Arithmetic = the numpy oracle (reference-faithful: dense AdamW over the whole local augmentation shard).
TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch

import oracle
from oracle.optim import dense_step, sparse_adam_step

F32 = np.float32
TABLES = ("user_encoder.embedding.weight", "item_encoder.embedding.weight",
          "adaptive_mimic.user_augmented.weight", "adaptive_mimic.item_augmented.weight")


def shard_state(state: dict, rank: int, world: int) -> dict:
    return {k: (np.ascontiguousarray(v[rank::world]) if k in TABLES else v.copy()) for k, v in state.items()}


def unshard_state(shards: list) -> dict:
    world = len(shards)
    out = {}
    for k, v in shards[0].items():
        if k in TABLES:
            n = sum(s[k].shape[0] for s in shards)
            full = np.empty((n,) + v.shape[1:], v.dtype)
            for r, s in enumerate(shards):
                full[r::world] = s[k]
            out[k] = full
        else:
            out[k] = v
    return out


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


class OracleEngine:
    def __init__(self, state_shard, *, lr, weight_decay, betas, lambdas):
        self.state = state_shard
        self.spec = oracle.spec_from_state(state_shard)
        self.opt = oracle.OptState()
        self.lr, self.wd, self.betas = lr, weight_decay, betas
        self.lambda_u, self.lambda_i = float(lambdas[0]), float(lambdas[1])
        self.mimic = self.spec.mimic
        self.t = 0

    def begin_step(self):
        self.t += 1

    def _forward_phase(self, users, items, Xu, Xi):
        u, it = users.numpy(), items.numpy()
        cu = oracle.tower_forward(self.state, "user", self.spec.user, u, None if Xu is None else Xu.numpy()[u], train=True)
        ci = oracle.tower_forward(self.state, "item", self.spec.item, it, None if Xi is None else Xi.numpy()[it], train=True)
        out = {}
        for name, c, idx, tab in (("cu", cu, u, "adaptive_mimic.user_augmented.weight"),
                                  ("ci", ci, it, "adaptive_mimic.item_augmented.weight")):
            q = self.state[tab][idx] if self.mimic else None
            out[name] = SimpleNamespace(t=_t(c["t"]), q=None if q is None else _t(q), cache=c, idx=idx)
        return out

    def _loss_phase(self, o_u, o_i, t_u, t_p, q_u, q_p, items, B, N, batch_fraction=1.0):
        D = o_u.shape[1]
        n = lambda x: None if x is None else x.numpy()
        L = oracle.loss_forward_backward(o_u.numpy(), o_i[:B].numpy(), o_i[B:].numpy().reshape(B, N, D), t_u=n(t_u), t_p=n(t_p),
                                         q_u=n(q_u), q_p=n(q_p), lambda_u=self.lambda_u, lambda_i=self.lambda_i)
        f = F32(batch_fraction)
        loss = np.array([L["loss"], L["bce"], L["mimic_user"] or 0.0, L["mimic_item"] or 0.0], F32) * f
        do_u = L["do_u"] * f
        do_i = np.concatenate([L["do_p"], L["do_n"].reshape(B * N, D)]) * f
        dq_u = dq_p = None
        if q_u is not None:
            dq_u = do_u + L["dq_u_extra"] * f
            dq_p = do_i[:B] + L["dq_p_extra"] * f
        return _t(loss), _t(do_u), _t(do_i), None if dq_u is None else _t(dq_u), None if dq_p is None else _t(dq_p)

    def _backward_phase(self, ctx, do_u, do_i, dq_user, dq_item, dense_grad_hook=None):
        st, spec = self.state, self.spec
        grads, sparse_rows = {}, {}
        oracle.tower_backward(st, "item", spec.item, ctx["ci"].cache, do_i.numpy(), grads, sparse_rows, train=True)
        oracle.tower_backward(st, "user", spec.user, ctx["cu"].cache, do_u.numpy(), grads, sparse_rows, train=True)
        dense_names = sorted(k for k in st if k not in TABLES and k in grads)
        if dense_grad_hook is not None:
            ts = [torch.from_numpy(grads[k]) for k in dense_names]      # shares memory: reduced in place
            dense_grad_hook(ts)
        for k in dense_names:
            dense_step("adamw", st[k], grads[k], self.opt.slot(k), lr=self.lr, weight_decay=self.wd)
        if self.mimic:                      # reference: dense gradient + AdamW over EVERY row of the (local) table
            for tab, c, dq in (("adaptive_mimic.user_augmented.weight", ctx["cu"], dq_user),
                               ("adaptive_mimic.item_augmented.weight", ctx["ci"], dq_item)):
                g = np.zeros_like(st[tab])
                np.add.at(g, c.idx, dq.numpy())
                dense_step("adamw", st[tab], g, self.opt.slot(tab), lr=self.lr, weight_decay=self.wd)
        self.touched = {}
        for tab in TABLES[:2]:
            ix = np.concatenate([r[0] for r in sparse_rows[tab]])
            v = np.concatenate([r[1] for r in sparse_rows[tab]], axis=0)
            self.touched[tab] = sparse_adam_step(st[tab], self.opt.slot(tab), ix, v, lr=self.lr, betas=self.betas)
