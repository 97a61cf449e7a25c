"""N-GPU form of the training step and of retrieval (BASELINE.json configs[2..4]).

Training — row-sharded tables, data-parallel batch, owner-compute:
  every rank holds rows r with r % W == rank of the ID tables, the augmentation tables, their optimiser state and the
  feature matrices, plus a replica of the small MLP / gate weights.  Per step and per rank (B local samples):
    1. bucket the B user ids and the B(1+N) item ids by owner          all-to-all #1  (int64 ids)
    2. OWNER: sort + lazy catch-up + tower forward on the rows it owns   (FusedEngine._forward_phase)
       and return [t | q] (2 D floats per row)                          all-to-all #2
    3. REQUESTER: fused loss forward/backward on its B samples, means over the GLOBAL batch
       (batch_fraction = 1/W), send [dL/dt | dL/dq] back               all-to-all #3
    4. OWNER: tower backward + segment-reduce + row-wise SparseAdam / lazy AdamW on its shard;
       all-reduce (sum) of the dense-weight gradients, identical dense AdamW on every rank
  The loss a rank returns is its share of the global loss: the shares add up (global_loss()).

Retrieval — item-sharded corpus: each rank scores ALL queries against its shard (local top-K with global ids), an
all-to-all by query block brings the W partial lists of a query to one rank, ttam_topk_merge merges them under
(-score, +id).  The result equals the one-GPU result bit for bit (same canonical scores, same order).

The collectives are torch.distributed (NCCL over NVLink 5 on the GPU box; gloo in the CPU tests, which drive this file
with an oracle-backed engine: tests/test_sharding.py).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import sharding as S


class ShardedEngine:
    """Drives the three phases of `FusedEngine` (or any object with the same phase methods) across ranks."""

    def __init__(self, engine, group=None) -> None:
        self.eng = engine
        self.group = group
        self.world = S._world(group)
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.last_exchange_rows = (0, 0)

    def _hook(self, grads: list) -> None:
        S.all_reduce_flat(grads, self.group)

    @torch.no_grad()
    def train_step(self, users: torch.Tensor, pos: torch.Tensor, neg: torch.Tensor, user_x_shard, item_x_shard):
        """users [B], pos [B], neg [B, N]: GLOBAL row ids of this rank's samples.  user_x_shard / item_x_shard: the rows
        of the feature matrices this rank owns (sharding.shard_rows).  Returns this rank's share loss[4] of the global
        loss {total, bce, mimic_user, mimic_item}."""
        eng, W = self.eng, self.world
        B, N = neg.shape
        items = torch.cat([pos.reshape(-1), neg.reshape(-1)])
        ex_u, ex_i = S.Exchange.build_many([users, items], W, self.group)
        if ex_u.n_owned == 0 or ex_i.n_owned == 0:
            raise RuntimeError("a rank owns none of the rows requested in this step; use a larger batch")
        self.last_exchange_rows = (ex_u.n_owned, ex_i.n_owned)
        eng.begin_step()
        ctx = eng._forward_phase(ex_u.local_rows.contiguous(), ex_i.local_rows.contiguous(), user_x_shard, item_x_shard)
        cu, ci = ctx["cu"], ctx["ci"]
        mimic = bool(eng.mimic)
        D = cu.t.shape[1]
        if mimic:
            tq_u = ex_u.to_requester(torch.cat([cu.t, cu.q], dim=1))
            tq_i = ex_i.to_requester(torch.cat([ci.t, ci.q], dim=1))
            t_u, q_u = tq_u[:, :D].contiguous(), tq_u[:, D:].contiguous()
            t_i, q_i = tq_i[:, :D].contiguous(), tq_i[:, D:].contiguous()
            o_u, o_i = t_u + q_u, t_i + q_i
            loss, do_u, do_i, dq_u, dq_p = eng._loss_phase(o_u, o_i, t_u, t_i[:B].contiguous(), q_u, q_i[:B].contiguous(),
                                                           items, B, N, batch_fraction=1.0 / W)
            g_u = ex_u.to_owner(torch.cat([do_u, dq_u], dim=1))
            g_i = ex_i.to_owner(torch.cat([do_i, torch.cat([dq_p, do_i[B:]], dim=0)], dim=1))
            eng._backward_phase(ctx, g_u[:, :D].contiguous(), g_i[:, :D].contiguous(), g_u[:, D:].contiguous(),
                                g_i[:, D:].contiguous(), dense_grad_hook=self._hook if W > 1 else None)
        else:
            o_u = ex_u.to_requester(cu.t)
            o_i = ex_i.to_requester(ci.t)
            loss, do_u, do_i, _, _ = eng._loss_phase(o_u.contiguous(), o_i.contiguous(), None, None, None, None, items, B, N,
                                                     batch_fraction=1.0 / W)
            eng._backward_phase(ctx, ex_u.to_owner(do_u), ex_i.to_owner(do_i), None, None,
                                dense_grad_hook=self._hook if W > 1 else None)
        return loss

    def global_loss(self, loss: torch.Tensor) -> torch.Tensor:
        out = loss.clone()
        if self.world > 1:
            dist.all_reduce(out, group=self.group)
        return out


def build_sharded_engine(model_shard, *, group=None, **engine_kwargs) -> ShardedEngine:
    """`model_shard`: a TwoTowerModel whose tables hold this rank's rows (num_embeddings = sharding.shard_size(...))."""
    from .engine import FusedEngine
    return ShardedEngine(FusedEngine(model_shard, **engine_kwargs), group=group)


class ShardedFlatIPIndex:
    """Item-sharded exact inner-product index.  Rank r holds `item_embeddings_shard` = rows r, r+W, r+2W, ... of the
    corpus (the training layout), or a contiguous block when `contiguous_offset` is given."""

    def __init__(self, item_embeddings_shard: torch.Tensor, *, group=None, dtype=torch.bfloat16, normalize: bool = False,
                 contiguous_offset: Optional[int] = None) -> None:
        from .retrieval import FlatIPIndex
        self.group = group
        self.world = S._world(group)
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.contiguous_offset = contiguous_offset
        self.local = FlatIPIndex(item_embeddings_shard, normalize=normalize, dtype=dtype,
                                 id_offset=0 if contiguous_offset is None else int(contiguous_offset))

    def search(self, queries: torch.Tensor, k: int):
        """queries [Q, D]: the SAME Q queries on every rank (Q a multiple of W).  Returns (ids, scores) of this rank's
        query block [rank*Q/W, (rank+1)*Q/W) over the WHOLE corpus."""
        from . import functional as F
        ids, scores = self.local.search(queries, k)
        if self.contiguous_offset is None and self.world > 1:
            ids = torch.where(ids >= 0, ids * self.world + self.rank, ids)     # local row -> global id (monotone: order kept)
        return S.merge_topk_shards(ids, scores, k, F.topk_merge, self.group)
