"""Row-sharding arithmetic and the row exchange of the N-GPU step.

Tables (ID embeddings, augmentation tables, their optimiser state) and feature matrices are sharded by row:
global row `r` lives on rank `r % W` at local row `r // W`.  The batch is data-parallel.  One `Exchange` per index
set per step routes requester rows to their owners and back:

    requester: idx[R]  --bucket by owner, all-to-all-->  owner: recv_idx[R'] (global ids), local_rows = recv_idx // W
    owner:  rows[R', C] --to_requester (all-to-all back + un-bucket)--> requester: rows[R, C] in the order of idx
    requester: grads[R, C] --to_owner (bucket + all-to-all)--> owner: grads[R', C] in the order of recv_idx

Pure torch + torch.distributed: the same code runs on gloo/CPU (tests/test_sharding.py, world_size 2) and on
nccl/CUDA (NVLink 5 all-to-all).  Nothing here computes model arithmetic.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


def owner_of(idx: torch.Tensor, world: int) -> torch.Tensor:
    return idx % world


def local_row(idx: torch.Tensor, world: int) -> torch.Tensor:
    return torch.div(idx, world, rounding_mode="floor")


def global_row(local: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    return local * world + rank


def shard_size(n: int, rank: int, world: int) -> int:
    """Rows of an n-row table that rank `rank` owns."""
    return (n - rank + world - 1) // world


def shard_rows(full: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """The rows of `full` that rank `rank` owns, in local-row order."""
    return full[rank::world].contiguous()


def unshard_rows(shards: list) -> torch.Tensor:
    """Inverse of shard_rows over all ranks (tests / checkpoint export)."""
    world = len(shards)
    n = sum(s.shape[0] for s in shards)
    out = shards[0].new_empty((n,) + tuple(shards[0].shape[1:]))
    for r, s in enumerate(shards):
        out[r::world] = s
    return out


def bucket_order(owner: torch.Tensor, world: int) -> torch.Tensor:
    """Stable permutation that groups positions by owner.  CUDA: one radix pass over log2(W) bits (ttam_sort_rows);
    CPU (gloo tests): torch's stable argsort."""
    if owner.is_cuda:
        from . import functional as F
        _, perm = F.sort_rows(owner.contiguous(), world)
        return perm
    return torch.argsort(owner, stable=True)


def _world(group) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


class Exchange:
    """All-to-all route of one index set.  `idx` [R] int64 global row ids requested by this rank."""

    def __init__(self, idx: torch.Tensor, world: Optional[int] = None, group=None, _counts=None) -> None:
        self.group = group
        self.world = W = _world(group) if world is None else int(world)
        idx = idx.reshape(-1)
        self.n_req = idx.numel()
        if W == 1:
            self.order = None
            self.recv_idx = idx
            self.local_rows = idx
            self.send_splits = self.recv_splits = [self.n_req]
            return
        owner = owner_of(idx, W)
        self.order = bucket_order(owner, W)                     # bucket by owner, original order kept inside a bucket
        if _counts is None:
            send_counts = torch.bincount(owner, minlength=W)
            recv_counts = torch.empty_like(send_counts)
            dist.all_to_all_single(recv_counts, send_counts, group=group)
            # the split sizes of the payload exchanges are host integers: one small D2H per step
            both = torch.stack([send_counts, recv_counts]).cpu()
            self.send_splits, self.recv_splits = both[0].tolist(), both[1].tolist()
        else:
            self.send_splits, self.recv_splits = _counts
        send_idx = idx[self.order]
        self.recv_idx = idx.new_empty(sum(self.recv_splits))
        dist.all_to_all_single(self.recv_idx, send_idx, self.recv_splits, self.send_splits, group=group)
        self.local_rows = local_row(self.recv_idx, W)

    @staticmethod
    def build_many(index_sets: list, world: Optional[int] = None, group=None) -> list:
        """Exchanges for several index sets of one step with ONE count all-to-all and ONE host synchronisation."""
        W = _world(group) if world is None else int(world)
        if W == 1:
            return [Exchange(ix, 1, group) for ix in index_sets]
        n = len(index_sets)
        send = torch.stack([torch.bincount(owner_of(ix.reshape(-1), W), minlength=W) for ix in index_sets], dim=1)  # [W, n]
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send.contiguous(), group=group)
        both = torch.stack([send, recv]).cpu()                  # the step's only D2H of sizes
        return [Exchange(ix, W, group, _counts=(both[0][:, j].tolist(), both[1][:, j].tolist()))
                for j, ix in enumerate(index_sets)]

    @property
    def n_owned(self) -> int:
        return self.recv_idx.numel()

    def to_requester(self, rows_owner: torch.Tensor) -> torch.Tensor:
        """rows_owner [R', C] (order of recv_idx) -> [R, C] in the order of the requester's idx."""
        if self.world == 1:
            return rows_owner
        C = rows_owner.shape[1:]
        bucketed = rows_owner.new_empty((self.n_req,) + tuple(C))
        dist.all_to_all_single(bucketed, rows_owner.contiguous(), self.send_splits, self.recv_splits, group=self.group)
        out = torch.empty_like(bucketed)
        out[self.order] = bucketed
        return out

    def to_owner(self, rows_req: torch.Tensor) -> torch.Tensor:
        """rows_req [R, C] (order of idx) -> [R', C] in the order of recv_idx."""
        if self.world == 1:
            return rows_req
        C = rows_req.shape[1:]
        bucketed = rows_req[self.order].contiguous()
        out = rows_req.new_empty((self.n_owned,) + tuple(C))
        dist.all_to_all_single(out, bucketed, self.recv_splits, self.send_splits, group=self.group)
        return out


def all_reduce_flat(tensors: list, group=None) -> None:
    """Sum the tensors over the ranks in ONE collective (the dense MLP / gate gradients: ~1.3 MB at D=96, H=192).
    When the tensors are consecutive views of one flat buffer (FusedEngine lays its dense gradients out that way) the
    buffer is reduced in place, without staging copies."""
    if _world(group) == 1 or not tensors:
        return
    base = tensors[0]._base if tensors[0]._base is not None else None
    if base is not None and base.dim() == 1 and all(t._base is base for t in tensors):
        # one span of the shared buffer covers them all (alignment gaps and gradients nobody asked for ride along)
        lo = min(t.data_ptr() for t in tensors)
        hi = max(t.data_ptr() + t.numel() * t.element_size() for t in tensors)
        first = (lo - base.data_ptr()) // base.element_size()
        dist.all_reduce(base[first:first + (hi - lo) // base.element_size()], group=group)
        return
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, group=group)
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n


def merge_topk_shards(ids: torch.Tensor, scores: torch.Tensor, k: int, merge_fn, group=None):
    """Item-sharded retrieval: every rank holds the partial top-K lists of ALL queries over ITS item shard
    (ids [Q, K'] global ids, scores [Q, K']).  Query block b goes to rank b: all-to-all of [Q/W, K'] blocks, then a
    per-query W-way merge under (-score, +id) by `merge_fn(ids [q, W, K'], scores) -> (ids [q, k], scores [q, k])`.
    Returns this rank's query block (rank r owns queries [r*Q/W, (r+1)*Q/W); Q must be a multiple of W)."""
    W = _world(group)
    Q, K1 = ids.shape
    if W == 1:
        return merge_fn(ids.view(Q, 1, K1), scores.view(Q, 1, K1), k)
    if Q % W != 0:
        raise ValueError(f"number of queries ({Q}) must be a multiple of the world size ({W})")
    q = Q // W
    ids_in, sc_in = torch.empty_like(ids), torch.empty_like(scores)
    dist.all_to_all_single(ids_in, ids.contiguous(), group=group)        # [W, q, K']: part w = rank w's list for MY queries
    dist.all_to_all_single(sc_in, scores.contiguous(), group=group)
    ids_in = ids_in.view(W, q, K1).transpose(0, 1).contiguous()
    sc_in = sc_in.view(W, q, K1).transpose(0, 1).contiguous()
    return merge_fn(ids_in, sc_in, k)
