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


def gather_rows_from_shards(shard: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Inverse of `shard_rows` across the ranks of `group`: every rank passes the rows it owns (local-row order) and gets the
    full [n_total, ...] tensor back (row r = rank r % W's local row r // W).  Shards differ by at most one row: they are
    padded to the longest for the all-gather.  Used for checkpoints in the reference's layout and for the corpus export."""
    W = _world(group)
    if W == 1:
        return shard
    rank = dist.get_rank(group)
    if shard.shape[0] != shard_size(n_total, rank, W):
        raise ValueError(f"rank {rank} holds {shard.shape[0]} rows of a {n_total}-row table, expected {shard_size(n_total, rank, W)}")
    longest = shard_size(n_total, 0, W)
    padded = shard.new_zeros((longest,) + tuple(shard.shape[1:]))
    padded[: shard.shape[0]].copy_(shard)
    parts = [torch.empty_like(padded) for _ in range(W)]
    dist.all_gather(parts, padded.contiguous(), group=group)
    return unshard_rows([parts[r][: shard_size(n_total, r, W)] for r in range(W)])


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


def retrieval_grid(rank: int, world: int, query_groups: int):
    """(query group g, item shard s, number of item shards S) of `rank` in a query-group x item-shard grid, W = G*S.
    The S ranks of a query group are contiguous (g*S .. g*S+S-1): after the per-group merge rank g*S+s holds query block
    g*S+s of W, the same contract as pure item sharding (G = 1)."""
    if query_groups < 1 or world % query_groups != 0:
        raise ValueError(f"query_groups ({query_groups}) must divide the world size ({world})")
    S_ = world // query_groups
    return rank // S_, rank % S_, S_


def retrieval_subgroups(world: int, query_groups: int, group=None) -> list:
    """One process group per query group (every rank must call this, in the same order: torch.distributed.new_group is
    collective over the default group).  [None] when there is a single query group spanning the whole `group`."""
    if query_groups == 1:
        return [group]
    S_ = world // query_groups
    return [dist.new_group(list(range(g * S_, (g + 1) * S_))) for g in range(query_groups)]


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


# ------------------------------------------------------------------------------------------------------------------
# Static-shape route: the same exchange with FIXED-capacity slots, so that a whole sharded step has no data-dependent
# shape, needs no host read of split sizes, and replays as one CUDA graph (collectives included).
def _round_cap(cap: int, n_req: int) -> int:
    """Multiple of 128 rows (the GEMM tile height: W*cap rows leave no ragged tile), never more than all requests."""
    return min((int(cap) + 127) // 128 * 128, (n_req + 127) // 128 * 128)


def default_slot_capacity(n_req: int, world: int) -> int:
    """Starting slots per (requester, owner) pair: what uniformly distributed ids need - n_req/W rows per owner plus six
    standard deviations.  Skewed ids (a Zipf head puts all duplicates of a hot row on one owner) overflow this once: the
    step that does not fit runs on the dynamic route and `grown_slot_capacity` sizes the slots from what it saw."""
    per = (n_req + world - 1) // world
    return _round_cap(per + 6.0 * per ** 0.5 + 32, n_req)


def calibrated_slot_capacity(max_seen: int, n_req: int, world: int) -> int:
    """Slots after the calibration steps of a shape: the largest bucket seen so far on any rank + five standard deviations
    of a bucket count (sqrt(n_req/W) bounds it) + 32 rows.  Keeps an overflow - an eager step plus a re-capture of the
    step's graphs - out of steady state when the ids are skewed but stationary (a Zipf head adds a constant to one
    owner's buckets: at W = 4 the uniform starting capacity sits only ~3.5 sigma above that owner's mean)."""
    per = (n_req + world - 1) // world
    return _round_cap(max_seen + 5.0 * per ** 0.5 + 32, n_req)


def grown_slot_capacity(max_count: int, n_req: int) -> int:
    """Slots after an overflow: the largest bucket any rank had in that step + 6 % + 64 rows."""
    return _round_cap(max_count * 1.06 + 64, n_req)


class SlotExchange:
    """All-to-all route of one index set with `cap` slots per (requester, owner) pair.

    Every collective is an equal-split all-to-all of W*cap rows.  A requester's unused slots are PADDING: they repeat the
    real ids of their bucket (cyclically), so the owner's set of touched rows is exactly the set of requested rows (the
    SparseAdam / lazy-AdamW row sets stay bit-identical to the unpadded step), their tower outputs are dropped on the
    way back, and their gradient rows are zeros (x + 0.0 == x: segment sums and weight gradients are unchanged).

    `plan(idx)` buckets the ids and raises `flag` (device int32) when a bucket overflows `cap` or is empty (no id to
    pad with); `agree()` max-reduces the flag over the ranks.  The caller reads it once per step and falls back to the
    dynamic `Exchange` for that step when it is set - the only host read of the static step.
    Same attribute / method names as `Exchange`: the step code does not care which one routes its rows."""

    def __init__(self, n_req: int, cap: int, world: int, group=None, device="cpu") -> None:
        self.group, self.world, self.n_req, self.cap = group, int(world), int(n_req), int(cap)
        W, dev = self.world, torch.device(device)
        self.n_slots = W * self.cap
        self.send_idx = torch.zeros(self.n_slots + 1, dtype=torch.int64, device=dev)      # +1: dump slot of overflowing ids
        self.slot_of = torch.zeros(self.n_req, dtype=torch.int64, device=dev)
        self.recv_idx = torch.zeros(self.n_slots, dtype=torch.int64, device=dev)
        self.local_rows = torch.zeros(self.n_slots, dtype=torch.int64, device=dev)
        self.flag = torch.zeros(1, dtype=torch.int32, device=dev)
        self._arange_w = torch.arange(W, device=dev)
        self._arange_r = torch.arange(self.n_req, device=dev)
        self._bufs: dict = {}
        self.req_of = None                   # int32 [n_slots], CUDA route only: request held by a slot, -1 = padding
        self.peer = None                     # symmetric-memory handle when the rows travel by peer loads / stores

    # ---- NVLink peer route: no all-to-all for the row payloads ---------------------------------------------------
    def enable_peer(self, D: int) -> None:
        """Allocate this exchange's row buffers in symmetric memory (torch.distributed._symmetric_memory: every rank's
        buffer is mapped into every other rank's address space over NVLink) and exchange the mappings.  COLLECTIVE.
        own_t / own_q: what this rank computes for the slots it serves - requesters LOAD their rows from here
        (ttam_slot_unpack with one peer base per owner); recv_a / recv_b: the gradient rows of those slots -
        requesters STORE them here (ttam_slot_pack).  The caller separates producers from consumers with
        `peer_barrier()` (a device-side barrier over NVLink signal pads, ~6 us, CUDA-graph capturable)."""
        import torch.distributed._symmetric_memory as symm
        group = self.group if self.group is not None else dist.group.WORLD
        buf = symm.empty((4, self.n_slots, D), dtype=torch.float32, device=self.send_idx.device)
        buf.zero_()
        self.peer = symm.rendezvous(buf, group)
        self._peer_buf = buf
        self.own_t, self.own_q, self.recv_a, self.recv_b = buf[0], buf[1], buf[2], buf[3]
        self._peer_ptrs = [int(p) for p in self.peer.buffer_ptrs]
        self._peer_rank = dist.get_rank(group)
        self._peer_D = D
        # the bucketed ids live in symmetric memory too: owners load them, no id all-to-all (the all-reduce of the
        # overflow flag that follows every plan() orders those loads after every rank's plan)
        ids = symm.empty(self.n_slots + 1, dtype=torch.int64, device=self.send_idx.device)
        ids.zero_()
        self._peer_ids = symm.rendezvous(ids, group)
        self._peer_id_ptrs = [int(p) + self._peer_rank * self.cap * 8 for p in self._peer_ids.buffer_ptrs]
        self.send_idx = ids

    def peer_barrier(self) -> None:
        self.peer.barrier(channel=0)

    def _peer_bases(self, which: int) -> list:
        """Address, in every rank's buffer, of the rows exchanged with THIS rank: sub-buffer `which`, bucket = my rank."""
        D = self._peer_D
        off = (which * self.n_slots + self._peer_rank * self.cap) * D * 4
        return [p + off for p in self._peer_ptrs]

    def publish(self, t_owner: torch.Tensor, q_owner: Optional[torch.Tensor] = None) -> None:
        """Make the rows this rank computed loadable by its peers: nothing to do when the tower already wrote them into
        own_t / own_q (ShardedEngine seeds the engine's buffers with them), one local copy otherwise."""
        for dst, src in ((self.own_t, t_owner), (self.own_q, q_owner)):
            if src is not None and src.data_ptr() != dst.data_ptr():
                dst.copy_(src[: self.n_slots])

    @property
    def n_owned(self) -> int:
        return self.n_slots

    def _buf(self, name, cols, dtype, like):
        key = (name, cols, dtype)
        t = self._bufs.get(key)
        if t is None:
            t = torch.zeros((self.n_slots + 1, cols), dtype=dtype, device=like.device)
            self._bufs[key] = t
        return t

    def plan(self, idx: torch.Tensor) -> None:
        """Bucket `idx` [n_req] by owner into the slot layout (no communication, no host read).
        CUDA: ttam_slot_plan (4 launches); CPU (gloo tests): the same layout from torch ops."""
        W, cap, R = self.world, self.cap, self.n_req
        idx = idx.reshape(-1)
        if idx.is_cuda:
            from . import functional as F
            if self.req_of is None:
                self.req_of = torch.zeros(self.n_slots, dtype=torch.int32, device=idx.device)
            F.slot_plan(idx, W, cap, send_idx=self.send_idx, slot_of=self.slot_of, req_of=self.req_of, flag=self.flag)
            return
        owner = owner_of(idx, W)
        order = bucket_order(owner, W).long()                  # stable: original order inside a bucket
        so, sid = owner[order], idx[order]
        counts = (owner.unsqueeze(1) == self._arange_w).sum(0)
        starts = torch.cumsum(counts, 0) - counts
        j = self._arange_r - starts[so]
        slot = torch.where(j < cap, so * cap + j, torch.full_like(j, self.n_slots))
        self.slot_of.index_copy_(0, order, slot)
        # padding slots repeat the bucket's real ids cyclically (slot j >= count holds the id of slot (j - count) % count;
        # an empty bucket is flagged and its content is meaningless)
        real = torch.clamp(counts, min=1, max=cap)
        jj = torch.arange(cap, device=idx.device).unsqueeze(0).expand(W, cap)
        src_j = torch.where(jj < counts.unsqueeze(1), jj, (jj - counts.unsqueeze(1)) % real.unsqueeze(1))
        src_pos = torch.clamp(starts.unsqueeze(1) + src_j, max=R - 1)
        self.send_idx[: self.n_slots].view(W, cap).copy_(sid[src_pos])
        self.send_idx.scatter_(0, slot, sid)                    # overflowing ids land in the dump slot
        self.flag.copy_(((counts > cap) | (counts == 0)).any().to(torch.int32).reshape(1))

    def agree(self) -> None:
        if self.world > 1:
            dist.all_reduce(self.flag, op=dist.ReduceOp.MAX, group=self.group)

    def exchange_ids(self) -> None:
        if self.peer is not None:
            from . import functional as F
            F.slot_ids(self._peer_id_ptrs, self.cap, self.recv_idx, self.local_rows)
            return
        if self.world > 1:
            dist.all_to_all_single(self.recv_idx, self.send_idx[: self.n_slots], group=self.group)
        else:
            self.recv_idx.copy_(self.send_idx[: self.n_slots])
        torch.div(self.recv_idx, self.world, rounding_mode="floor", out=self.local_rows)

    def _a2a(self, out: torch.Tensor, inp: torch.Tensor) -> None:
        dist.all_to_all_single(out, inp, group=self.group)

    def pull(self, t_owner: torch.Tensor, q_owner: Optional[torch.Tensor] = None):
        """Forward exchange.  t_owner / q_owner [W*cap, D]: what this rank computed for the slots it received.
        Returns (t, q, o = t + q), each [n_req, D] in the order of the requester's ids (q None and o = t without q_owner)."""
        D, n = t_owner.shape[1], self.n_slots
        if self.peer is not None:            # published + barrier already passed: load straight from the owners' buffers
            from . import functional as F
            t = self._buf("t", D, t_owner.dtype, t_owner)[: self.n_req]
            q = self._buf("q", D, t_owner.dtype, t_owner)[: self.n_req] if q_owner is not None else None
            o = self._buf("o", D, t_owner.dtype, t_owner)[: self.n_req] if q_owner is not None else None
            F.slot_unpack(self._peer_bases(0), self._peer_bases(1) if q_owner is not None else None, D, self.cap, self.slot_of, D,
                          t_out=t, q_out=q, o_out=o)
            return t, q, (o if q is not None else t)
        srcs = []
        for name, rows in (("back_t", t_owner), ("back_q", q_owner)):
            if rows is None:
                srcs.append(None)
            elif self.world > 1:
                back = self._buf(name, D, rows.dtype, rows)
                self._a2a(back[:n], rows[:n].contiguous())
                srcs.append(back)
            else:
                srcs.append(rows)
        if t_owner.is_cuda:
            from . import functional as F
            t = self._buf("t", D, t_owner.dtype, t_owner)[: self.n_req]
            q = self._buf("q", D, t_owner.dtype, t_owner)[: self.n_req] if q_owner is not None else None
            o = self._buf("o", D, t_owner.dtype, t_owner)[: self.n_req] if q_owner is not None else None
            step = self.cap * srcs[0].stride(0) * 4
            bases = lambda x: None if x is None else [x.data_ptr() + w * step for w in range(self.world)]
            F.slot_unpack(bases(srcs[0]), bases(srcs[1]), srcs[0].stride(0), self.cap, self.slot_of, D, t_out=t, q_out=q, o_out=o)
            return t, q, (o if q is not None else t)
        pad = lambda x: x if x.shape[0] > n else torch.cat([x, x.new_zeros((1, D))])      # dump slot of the ids that did not fit
        t = pad(srcs[0]).index_select(0, self.slot_of)
        if q_owner is None:
            return t, None, t
        q = pad(srcs[1]).index_select(0, self.slot_of)
        return t, q, t + q

    def push(self, a: torch.Tensor, b0: Optional[torch.Tensor] = None, b1: Optional[torch.Tensor] = None):
        """Backward exchange.  a [n_req, D] and, optionally, b rows (b0[r] for r < len(b0), b1[r] beyond) in the order of
        the requester's ids.  Returns (ga, gb) [W*cap, D] in this rank's slot order; padding slots carry zeros."""
        D, n = a.shape[1], self.n_slots
        has_b = b0 is not None or b1 is not None
        if self.peer is not None:            # store straight into the owners' receive buffers; the caller's barrier follows
            from . import functional as F
            F.slot_pack(a, b0, b1, self.req_of, self.cap, self._peer_bases(2), self._peer_bases(3) if has_b else None, D)
            return self.recv_a, (self.recv_b if has_b else None)
        direct = self.world == 1
        send_a = self._buf("out_a" if direct else "send_a", D, a.dtype, a)
        send_b = self._buf("out_b" if direct else "send_b", D, a.dtype, a) if has_b else None
        if a.is_cuda:
            from . import functional as F
            step = self.cap * D * 4
            bases = lambda x: None if x is None else [x.data_ptr() + w * step for w in range(self.world)]
            F.slot_pack(a, b0, b1, self.req_of, self.cap, bases(send_a), bases(send_b), D)
        else:
            send_a.zero_()
            send_a.index_copy_(0, self.slot_of, a)
            if has_b:
                n0 = 0 if b0 is None else b0.shape[0]
                b = b0 if n0 >= self.n_req else (b1 if n0 == 0 else torch.cat([b0, b1[n0:]]))
                send_b.zero_()
                send_b.index_copy_(0, self.slot_of, b)
        if direct:
            return send_a[:n], (send_b[:n] if has_b else None)
        out_a = self._buf("out_a", D, a.dtype, a)
        self._a2a(out_a[:n], send_a[:n])
        out_b = None
        if has_b:
            out_b = self._buf("out_b", D, a.dtype, a)
            self._a2a(out_b[:n], send_b[:n])
        return out_a[:n], (out_b[:n] if has_b else None)
