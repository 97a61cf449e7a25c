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

import os
from typing import Optional

import torch
import torch.distributed as dist

from . import functional as F
from . import sharding as S


class _Static:
    """Buffers of the static-shape step for one (B, N): input copies, the two slot exchanges, the combined flag."""

    def __init__(self, B, N, world, group, device, cap_u, cap_i):
        self.B, self.N = B, N
        self.users = torch.zeros(B, dtype=torch.int64, device=device)
        self.items = torch.zeros(B * (1 + N), dtype=torch.int64, device=device)
        self.ex_u = S.SlotExchange(B, cap_u, world, group, device)
        self.ex_i = S.SlotExchange(B * (1 + N), cap_i, world, group, device)
        self.flag = torch.zeros(1, dtype=torch.int32, device=device)
        self.flag_host = torch.zeros(1, dtype=torch.int32)
        if torch.device(device).type == "cuda":
            self.flag_host = self.flag_host.pin_memory()
        self.graph_plan = self.graph_main = None
        self.graph_bwd = None               # speculative replay: graph_main holds the forward + loss half, this the rest
        self.flag_event = torch.cuda.Event() if torch.device(device).type == "cuda" else None
        self.loss = None
        self.x_key = None
        self.calib_left = 0                 # calibration steps still to run on this shape (see ShardedEngine._calibrate)
        self.max_seen = [0, 0]


class ShardedEngine:
    """Drives the three phases of `FusedEngine` (or any object with the same phase methods) across ranks.

    static=False: every step sizes its all-to-alls from the step's own counts (one host read per step, eager launches).
    static=True:  fixed-capacity slots per (requester, owner) pair (`sharding.SlotExchange`): no data-dependent shape, so
                  with graph=True the whole step - towers, loss, optimisers AND the collectives - is two CUDA-graph
                  replays (plan, main) around the step's single host read (the overflow flag).  A step whose ids do not
                  fit the slots runs on the dynamic route instead (same results), and the capacity is re-sized from it.
    peer=True (with static): the ids and row payloads of the exchanges do not go through NCCL at all - requesters load the
                  owners' rows and store their gradient rows over NVLink peer mappings (symmetric memory), fused into
                  the un-bucket / re-bucket kernels, between two device-side barriers per step.  Falls back to the NCCL
                  slot route (with a warning) when symmetric memory cannot be set up on every rank."""

    def __init__(self, engine, group=None, *, static: bool = False, capacity=None, peer: bool = False) -> None:
        self.eng = engine
        self.group = group
        self.world = S._world(group)
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.last_exchange_rows = (0, 0)
        self.static = bool(static)
        self.capacity = capacity              # None or (cap_users, cap_items)
        self._auto_capacity = capacity is None
        self._calibrated: set = set()
        self.peer = bool(peer) and self.static and self.world > 1
        self._static: dict = {}
        self.fallback_steps = 0
        self.launches_per_step = 0
        self.peer_error = None

    def _hook(self, grads: list) -> None:
        S.all_reduce_flat(grads, self.group)

    # ---- the step body, shared by both routes ---------------------------------------------------------------------
    def _fused_slot_loss(self, N: int, D: int) -> bool:
        eng = self.eng
        if os.environ.get("TTAM_FUSE_SLOT_LOSS", "1") == "0" or getattr(eng, "loss_kind", "sampled") != "sampled":
            return False
        if eng.lambda_c > 0 and eng.cat_tensor is not None and eng.major is not None:
            return False
        return F.loss_aug_supported(N, D)

    def _speculative(self, st: "_Static") -> bool:
        """Can the forward + loss half of the step be launched before the host has read the overflow flag?  It must not
        change anything a discarded step would have to undo: that holds for the peer route with the fused pull + loss + push
        kernel (buffers only; the lazy catch-up it runs is value-preserving; the device step counter is rewound)."""
        eng = self.eng
        return (os.environ.get("TTAM_SPECULATE", "1") != "0" and st.ex_i.peer is not None and bool(eng.mimic)
                and self._fused_slot_loss(st.N, eng.user.out_dim))

    def _body_fwd(self, ex_u, ex_i, B, N, user_x_shard, item_x_shard):
        """Owner-side forward, then pull, loss and push as ONE kernel: the pair's rows are loaded from their owners and its
        gradient rows stored into the owners' receive buffers (zeroed at the start of the step: padding slots keep zero
        rows).  Writes buffers only (see _speculative)."""
        eng = self.eng
        ctx = eng._forward_phase(ex_u.local_rows.contiguous(), ex_i.local_rows.contiguous(), user_x_shard, item_x_shard)
        cu, ci = ctx["cu"], ctx["ci"]
        ex_u.publish(cu.t, cu.q)
        ex_i.publish(ci.t, ci.q)
        ex_i.peer_barrier()               # every owner's rows are in place
        loss = eng._misc("loss", (4,), torch.float32)
        F.loss_slots_fwd_bwd(ex_u._peer_bases(0), ex_u._peer_bases(1), ex_i._peer_bases(0), ex_i._peer_bases(1),
                             ex_u._peer_bases(2), ex_u._peer_bases(3), ex_i._peer_bases(2), ex_i._peer_bases(3),
                             ex_u.cap, ex_i.cap, ex_u.slot_of, ex_i.slot_of, B, N, cu.t.shape[1], lambda_u=eng.lambda_u,
                             lambda_i=eng.lambda_i, loss=loss, batch_fraction=1.0 / self.world)
        ex_i.peer_barrier()               # every requester's gradient rows have landed
        return ctx, loss

    def _body_bwd(self, ctx, ex_u, ex_i) -> None:
        self.eng._backward_phase(ctx, ex_u.recv_a, ex_i.recv_a, ex_u.recv_b, ex_i.recv_b,
                                 dense_grad_hook=self._hook if self.world > 1 else None)

    def _body(self, ex_u, ex_i, items, B, N, user_x_shard, item_x_shard):
        eng, W = self.eng, self.world
        mimic = bool(eng.mimic)
        if (isinstance(ex_u, S.SlotExchange) and ex_i.peer is not None and mimic
                and self._fused_slot_loss(N, eng.user.out_dim)):
            ctx, loss = self._body_fwd(ex_u, ex_i, B, N, user_x_shard, item_x_shard)
            self._body_bwd(ctx, ex_u, ex_i)
            return loss
        ctx = eng._forward_phase(ex_u.local_rows.contiguous(), ex_i.local_rows.contiguous(), user_x_shard, item_x_shard)
        cu, ci = ctx["cu"], ctx["ci"]
        D = cu.t.shape[1]
        hook = self._hook if W > 1 else None
        bf = 1.0 / W
        if isinstance(ex_u, S.SlotExchange):      # static route: fused un-bucket / re-bucket kernels around the collectives
            peer = ex_i.peer is not None          # rows travel by NVLink peer loads / stores between two device barriers
            if peer:
                ex_u.publish(cu.t, cu.q if mimic else None)
                ex_i.publish(ci.t, ci.q if mimic else None)
                ex_i.peer_barrier()               # every owner's rows are in place
            if mimic:
                t_u, q_u, o_u = ex_u.pull(cu.t, cu.q)
                t_i, q_i, o_i = ex_i.pull(ci.t, ci.q)
                loss, do_u, do_i, dq_u, dq_p = eng._loss_phase(o_u, o_i, t_u, t_i[:B], q_u, q_i[:B], items, B, N, batch_fraction=bf)
                gu_a, gu_b = ex_u.push(do_u, dq_u)
                gi_a, gi_b = ex_i.push(do_i, dq_p, do_i)          # aug-table rows: dq of the positives, do of the negatives
                if peer:
                    ex_i.peer_barrier()           # every requester's gradient rows have landed
                eng._backward_phase(ctx, gu_a, gi_a, gu_b, gi_b, dense_grad_hook=hook)
            else:
                _, _, o_u = ex_u.pull(cu.t)
                _, _, o_i = ex_i.pull(ci.t)
                loss, do_u, do_i, _, _ = eng._loss_phase(o_u, o_i, None, None, None, None, items, B, N, batch_fraction=bf)
                gu_a, gi_a = ex_u.push(do_u)[0], ex_i.push(do_i)[0]
                if peer:
                    ex_i.peer_barrier()
                eng._backward_phase(ctx, gu_a, gi_a, None, None, dense_grad_hook=hook)
            return loss
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
                                g_i[:, D:].contiguous(), dense_grad_hook=hook)
        else:
            o_u = ex_u.to_requester(cu.t)
            o_i = ex_i.to_requester(ci.t)
            loss, do_u, do_i, _, _ = eng._loss_phase(o_u.contiguous(), o_i.contiguous(), None, None, None, None, items, B, N,
                                                     batch_fraction=1.0 / W)
            eng._backward_phase(ctx, ex_u.to_owner(do_u), ex_i.to_owner(do_i), None, None, dense_grad_hook=hook)
        return loss

    # ---- dynamic route ----------------------------------------------------------------------------------------------
    def _dynamic_step(self, users, pos, neg, user_x_shard, item_x_shard):
        B, N = neg.shape
        items = torch.cat([pos.reshape(-1), neg.reshape(-1)])
        ex_u, ex_i = S.Exchange.build_many([users, items], self.world, self.group)
        if ex_u.n_owned == 0 or ex_i.n_owned == 0:
            raise RuntimeError("a rank owns none of the rows requested in this step; use a larger batch")
        self.last_exchange_rows = (ex_u.n_owned, ex_i.n_owned)
        self._last_max_bucket = (max(ex_u.send_splits), max(ex_i.send_splits))
        self.eng.begin_step()
        return self._body(ex_u, ex_i, items, B, N, user_x_shard, item_x_shard)

    # ---- static route -----------------------------------------------------------------------------------------------
    def _static_state(self, B, N, device) -> _Static:
        st = self._static.get((B, N))
        if st is None:
            cap_u, cap_i = self.capacity or (S.default_slot_capacity(B, self.world),
                                             S.default_slot_capacity(B * (1 + N), self.world))
            if self.world > 1:
                # equal-split exchanges: every rank must bring the same B and N and use the same capacities (checked once
                # per shape; the dynamic route has no such requirement)
                mine = torch.tensor([B, N, cap_u, cap_i], dtype=torch.int64, device=device)
                lo, hi = mine.clone(), mine.clone()
                dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=self.group)
                dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=self.group)
                if not torch.equal(lo, hi):
                    raise ValueError(f"static route: ranks disagree on (B, N, slot capacities): min {lo.tolist()}, max {hi.tolist()}; "
                                     "use the same per-rank batch on every rank, or static=False")
            st = _Static(B, N, self.world, self.group, device, cap_u, cap_i)
            if self.peer:
                # row payloads over NVLink peer memory: the exchanges own symmetric buffers, and the towers write their
                # t / q rows straight into them (the engine looks its buffers up by name before allocating)
                eng = self.eng
                ok = 1
                try:
                    for ex, plan in ((st.ex_u, eng.user), (st.ex_i, eng.item)):
                        ex.enable_peer(plan.out_dim)
                except Exception as e:  # noqa: BLE001 - no symmetric memory on this system: every rank must learn it
                    ok, self.peer_error = 0, f"{type(e).__name__}: {e}"
                agree = torch.tensor([ok], dtype=torch.int32, device=device)
                dist.all_reduce(agree, op=dist.ReduceOp.MIN, group=self.group)
                if int(agree) == 1:
                    for ex, plan, bufs in ((st.ex_u, eng.user, eng.bufs_u), (st.ex_i, eng.item, eng.bufs_i)):
                        if plan.D == plan.out_dim:
                            bufs["t"], bufs["q"] = ex.own_t, ex.own_q
                else:
                    # same slots, equal-split NCCL all-to-alls instead of peer loads / stores (route "static")
                    import warnings
                    warnings.warn("symmetric memory is not available on every rank "
                                  f"({getattr(self, 'peer_error', 'a peer failed')}); using the NCCL slot route")
                    self.peer = False
                    st = _Static(B, N, self.world, self.group, device, cap_u, cap_i)
            if self._auto_capacity and self.world > 1 and (B, N) not in self._calibrated:
                st.calib_left = self.CALIBRATION_STEPS
            self._static[(B, N)] = st
        return st

    CALIBRATION_STEPS = 4

    def _calibrate(self, st: _Static) -> None:
        """First steps of a shape (automatic capacities only): record the largest bucket any rank fills - one bincount per
        index set, one MAX all-reduce and one host read per calibration step - and, when they are over, re-size the slots to
        `sharding.calibrated_slot_capacity` if that is more than they hold.  Calibration steps run eagerly; the (possibly
        re-sized) state records its graphs on the first step after them, i.e. still inside the caller's warm-up."""
        W, B, N = self.world, st.B, st.N
        mx = torch.stack([torch.bincount(S.owner_of(st.users, W), minlength=W).max(),
                          torch.bincount(S.owner_of(st.items, W), minlength=W).max()])
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=self.group)
        seen = mx.tolist()
        st.max_seen = [max(a, int(b)) for a, b in zip(st.max_seen, seen)]
        st.calib_left -= 1
        if st.calib_left == 0:
            self._calibrated.add((B, N))
            cap_u = max(st.ex_u.cap, S.calibrated_slot_capacity(st.max_seen[0], B, W))
            cap_i = max(st.ex_i.cap, S.calibrated_slot_capacity(st.max_seen[1], B * (1 + N), W))
            if (cap_u, cap_i) != (st.ex_u.cap, st.ex_i.cap):
                self.capacity = (cap_u, cap_i)
                del self._static[(B, N)]

    def _plan(self, st: _Static) -> None:
        st.ex_u.plan(st.users)
        st.ex_i.plan(st.items)
        torch.maximum(st.ex_u.flag, st.ex_i.flag, out=st.flag)
        if self.world > 1:                  # every rank takes the same route
            dist.all_reduce(st.flag, op=dist.ReduceOp.MAX, group=self.group)
        st.flag_host.copy_(st.flag, non_blocking=True)
        # Queued BEHIND the flag's copy, so that they run while the host reads the flag and launches the main half (the GPU
        # would idle through that round trip otherwise): the id exchange (valid whatever the flag says - the all-reduce
        # above has ordered every rank's plan before it) and the zeroing of the receive buffers of the fused loss.
        st.ex_u.exchange_ids()
        st.ex_i.exchange_ids()
        eng = self.eng
        if st.ex_i.peer is not None and eng.mimic and self._fused_slot_loss(st.N, eng.user.out_dim):
            for ex in (st.ex_u, st.ex_i):          # the fused loss writes real slots only: padding slots must read as zero rows
                ex.recv_a.zero_(); ex.recv_b.zero_()

    def _main(self, st: _Static, user_x_shard, item_x_shard):
        return self._body(st.ex_u, st.ex_i, st.items, st.B, st.N, user_x_shard, item_x_shard)

    def _overflowed(self, st: _Static) -> bool:
        if st.flag.is_cuda:
            torch.cuda.current_stream(st.flag.device).synchronize()      # the step's only host read
        return bool(int(st.flag_host[0]))

    def _grow(self, B, N) -> None:
        """After a step that did not fit (it ran on the dynamic route): size the slots from the largest bucket any rank had
        in that step (+6 % + 64 rows); graphs and buffers of this shape are rebuilt on next use."""
        st = self._static[(B, N)]
        big = torch.tensor(self._last_max_bucket, dtype=torch.int64, device=st.flag.device)
        if self.world > 1:
            dist.all_reduce(big, op=dist.ReduceOp.MAX, group=self.group)
        big_u, big_i = (int(v) for v in big.tolist())
        cap_u = st.ex_u.cap if big_u <= st.ex_u.cap else S.grown_slot_capacity(big_u, B)
        cap_i = st.ex_i.cap if big_i <= st.ex_i.cap else S.grown_slot_capacity(big_i, B * (1 + N))
        # the dynamic step may have re-allocated engine buffers the recorded graphs point into: never replay them again
        st.graph_plan = st.graph_main = st.graph_bwd = None
        if (cap_u, cap_i) != (st.ex_u.cap, st.ex_i.cap):       # (an EMPTY bucket also raises the flag: nothing to grow then)
            self.capacity = (cap_u, cap_i)
            del self._static[(B, N)]

    def _static_step(self, users, pos, neg, user_x_shard, item_x_shard, graph: bool):
        eng = self.eng
        B, N = neg.shape
        st = self._static_state(B, N, users.device)
        st.users.copy_(users)
        st.items[:B].copy_(pos.reshape(-1))
        st.items[B:].copy_(neg.reshape(-1))
        x_key = (None if user_x_shard is None else user_x_shard.data_ptr(), None if item_x_shard is None else item_x_shard.data_ptr())
        # no graph is recorded before the calibration steps of the shape are over: the slots may still be re-sized
        use_graph = graph and users.is_cuda and st.calib_left == 0
        replay = use_graph and st.graph_main is not None and st.x_key == x_key
        speculated = False
        if replay:
            st.graph_plan.replay()
            if st.graph_bwd is not None:
                # Forward + loss go out BEFORE the host looks at the flag: its copy lands while they run, so the host round
                # trip (read the flag, launch the rest) no longer leaves the GPU idle.  A step that overflowed throws the
                # speculative half away: it wrote buffers only, the device step counter is rewound below.
                st.flag_event.record()
                eng.begin_step()
                st.graph_main.replay()
                st.flag_event.synchronize()
                speculated = True
        else:
            self._plan(st)
        if (bool(int(st.flag_host[0])) if speculated else self._overflowed(st)):
            if speculated:
                eng.t -= 1
                eng.state.sub_(torch.tensor([1, 1 << 36], dtype=torch.int64, device=eng.state.device))
            self.fallback_steps += 1
            loss = self._dynamic_step(users, pos, neg, user_x_shard, item_x_shard)
            self._calibrated.add((B, N))         # the slots are about to be sized from a real overflow
            self._grow(B, N)
            return loss
        calibrating = st.calib_left > 0
        self.last_exchange_rows = (st.ex_u.n_slots, st.ex_i.n_slots)
        if speculated:
            st.graph_bwd.replay()
            return st.loss
        eng.begin_step()
        if replay:
            st.graph_main.replay()
            return st.loss
        count = getattr(eng, "launch_count", None)
        c0 = count() if count is not None else 0
        loss = self._main(st, user_x_shard, item_x_shard)
        if count is not None:
            self.launches_per_step = count() - c0            # libttam launches of one step (what a graph replay re-issues)
        if use_graph:
            # that eager step sized every buffer and opened the NCCL channels; record both halves for the next steps
            # (capture does not execute anything, so the step counters stay where they are)
            torch.cuda.synchronize()
            gp, gm = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(gp, capture_error_mode="thread_local"):
                self._plan(st)
            if self._speculative(st):
                gb = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gm, capture_error_mode="thread_local"):
                    ctx, st.loss = self._body_fwd(st.ex_u, st.ex_i, st.B, st.N, user_x_shard, item_x_shard)
                with torch.cuda.graph(gb, pool=gm.pool(), capture_error_mode="thread_local"):
                    self._body_bwd(ctx, st.ex_u, st.ex_i)
                st.graph_bwd = gb
            else:
                with torch.cuda.graph(gm, capture_error_mode="thread_local"):
                    st.loss = self._main(st, user_x_shard, item_x_shard)
            st.graph_plan, st.graph_main, st.x_key = gp, gm, x_key
        if calibrating:
            self._calibrate(st)
        return loss

    @torch.no_grad()
    def train_step(self, users: torch.Tensor, pos: torch.Tensor, neg: torch.Tensor, user_x_shard, item_x_shard, *,
                   graph: bool = False):
        """users [B], pos [B], neg [B, N]: GLOBAL row ids of this rank's samples.  user_x_shard / item_x_shard: the rows
        of the feature matrices this rank owns (sharding.shard_rows).  Returns this rank's share loss[4] of the global
        loss {total, bce, mimic_user, mimic_item} (valid until the next step).  graph=True (static route, CUDA): replay
        the step as CUDA graphs; B and N must then be the same on every rank."""
        if self.static:
            return self._static_step(users, pos, neg, user_x_shard, item_x_shard, graph)
        return self._dynamic_step(users, pos, neg, user_x_shard, item_x_shard)

    # ---- checkpoints in the reference's layout (reference training.py:141-182) ---------------------------
    ROW_SHARDED = ("encoder.embedding.weight", "_augmented.weight")

    def _is_row_sharded(self, name: str) -> bool:
        return name.endswith(self.ROW_SHARDED)

    @torch.no_grad()
    def full_state_dict(self, num_users: int, num_items: int, *, to_cpu: bool = True) -> dict:
        """`model.state_dict()` of the UNSHARDED model, with the reference's key names, assembled on every rank: lazily-updated
        rows are flushed first, row-sharded tables ({user,item}_encoder.embedding.weight, adaptive_mimic.*_augmented.weight)
        are gathered from all ranks, replicated tensors are taken from this rank.  What `_clone_state_dict` /
        `_save_checkpoint` would have stored for a one-process run.  COLLECTIVE."""
        flush = getattr(self.eng, "flush", None)
        if flush is not None:
            flush()
        out = {}
        for name, t in self.eng.model.state_dict().items():
            if self._is_row_sharded(name):
                n = num_users if ("user_encoder" in name or "user_augmented" in name) else num_items
                t = S.gather_rows_from_shards(t.detach(), n, self.group)
            out[name] = t.detach().cpu().clone() if to_cpu else t.detach().clone()
        return out

    @torch.no_grad()
    def load_full_state_dict(self, state: dict) -> None:
        """Load an unsharded (reference-layout) state_dict: this rank keeps rows r % W == rank of the row-sharded tables."""
        mine = {}
        for name, t in state.items():
            mine[name] = S.shard_rows(t, self.rank, self.world) if self._is_row_sharded(name) else t
        self.eng.model.load_state_dict(mine)

    def _reference_param_groups(self):
        """(dense names, sparse names) in the order the reference hands the parameters to its two optimisers
        (`_collect_parameter_groups`, training.py:276-309): per encoder the embedding weight first (SparseAdam when the
        embedding is sparse), then its other parameters, then whatever the model has left (the augmentation tables)."""
        model = self.eng.model
        names = {id(p): n for n, p in model.named_parameters()}
        dense, sparse, seen = [], [], set()

        def add(p, coll):
            if id(p) not in seen:
                seen.add(id(p))
                coll.append(names[id(p)])
        for enc in (model.user_encoder, model.item_encoder):
            emb = getattr(enc, "embedding", None)
            if isinstance(emb, torch.nn.Embedding):
                add(emb.weight, sparse if getattr(emb, "sparse", False) else dense)
            for n, p in enc.named_parameters():
                if n != "embedding.weight":
                    add(p, dense)
        for p in model.parameters():
            add(p, dense)
        return dense, sparse

    def _torch_optimizers(self, dense, sparse):
        """Empty stand-ins of the reference's optimisers (training.py:1315-1346): they only supply `param_groups` in the
        layout of the installed torch version."""
        eng = self.eng
        mk = lambda names: [torch.nn.Parameter(torch.empty(0)) for _ in names]
        hp = lambda k, d: getattr(eng, k, d)
        opts = []
        if dense:
            kind = hp("kind", "adamw")
            if kind == "sgd":
                opts.append(torch.optim.SGD(mk(dense), lr=hp("lr", 1e-3), weight_decay=hp("wd", 0.0), momentum=hp("momentum", 0.0)))
            else:
                cls = torch.optim.AdamW if kind == "adamw" else torch.optim.Adam
                opts.append(cls(mk(dense), lr=hp("lr", 1e-3), weight_decay=hp("wd", 0.0), betas=hp("dense_betas", (0.9, 0.999)),
                                eps=hp("eps", 1e-8)))
        if sparse:
            opts.append(torch.optim.SparseAdam(mk(sparse), lr=hp("lr", 1e-3), betas=hp("sparse_betas", (0.9, 0.999)), eps=hp("eps", 1e-8)))
        return opts

    @torch.no_grad()
    def save_checkpoint(self, path, num_users: int, num_items: int, *, epoch: int = 0, metric_name=None, metric_value=None):
        """Rank 0 writes the reference's checkpoint (training.py:150-182): `model_state_dict` un-sharded, and
        `optimizer_state_dicts` = one `optimizer.state_dict()` per reference optimiser - the dense one, then SparseAdam - in
        torch's own layout ({'state': {index: {step, exp_avg, exp_avg_sq}}, 'param_groups': [...]}, parameters in the order
        the reference passes them), so `optimizer.load_state_dict` of a reference run accepts it.  Row-sharded moments are
        gathered like the tables.  COLLECTIVE."""
        import time
        model_state = self.full_state_dict(num_users, num_items)
        dense, sparse = self._reference_param_groups()
        opts = self._torch_optimizers(dense, sparse)
        st = self.eng.optimizer_state() if hasattr(self.eng, "optimizer_state") else {}
        dicts = []
        for names, opt in zip([g for g in (dense, sparse) if g], opts):
            sd = opt.state_dict()
            is_sparse = isinstance(opt, torch.optim.SparseAdam)
            for j, name in enumerate(names):            # every rank walks the same names: the gathers are collective
                ent = st.get(name)
                if ent is None:
                    continue
                slot = {"step": int(ent["step"]) if is_sparse else torch.tensor(float(ent["step"]))}
                for k in ("exp_avg", "exp_avg_sq"):
                    v = ent.get(k)
                    if v is None:
                        continue
                    if self._is_row_sharded(name):
                        n = num_users if ("user_encoder" in name or "user_augmented" in name) else num_items
                        v = S.gather_rows_from_shards(v, n, self.group)
                    key = "momentum_buffer" if (k == "exp_avg" and isinstance(opt, torch.optim.SGD)) else k
                    slot[key] = v.detach().cpu().clone()
                if int(ent["step"]) > 0:
                    sd["state"][j] = slot
            dicts.append(sd)
        if self.rank == 0:
            torch.save({"epoch": epoch, "model_state_dict": model_state, "optimizer_state_dicts": dicts,
                        "metric_name": metric_name, "metric_value": metric_value, "timestamp": time.time()}, path)
        if self.world > 1:
            dist.barrier(group=self.group)

    @torch.no_grad()
    def load_checkpoint(self, path) -> dict:
        """Resume from a checkpoint in the reference's layout (one written by `save_checkpoint`, by the one-GPU hooks, or by
        the reference itself): every rank reads the file and keeps its rows of the tables and of their moments.  Returns
        the checkpoint's metadata (epoch, metric_name, metric_value)."""
        ck = torch.load(path, map_location="cpu", weights_only=False)
        self.load_full_state_dict(ck["model_state_dict"])
        dense, sparse = self._reference_param_groups()
        state, step = {}, 0
        for names, sd in zip([g for g in (dense, sparse) if g], ck.get("optimizer_state_dicts", [])):
            for j, slot in sd.get("state", {}).items():
                name = names[int(j)]
                step = max(step, int(float(slot.get("step", 0))))
                ent = {}
                for k_src, k_dst in (("exp_avg", "exp_avg"), ("momentum_buffer", "exp_avg"), ("exp_avg_sq", "exp_avg_sq")):
                    v = slot.get(k_src)
                    if v is not None:
                        ent[k_dst] = S.shard_rows(v, self.rank, self.world) if self._is_row_sharded(name) else v
                state[name] = ent
        if hasattr(self.eng, "load_optimizer_state"):
            self.eng.load_optimizer_state(state, step)
        self._static.clear()            # recorded graphs address the old step state
        return {k: ck.get(k) for k in ("epoch", "metric_name", "metric_value")}

    @torch.no_grad()
    def export_item_embeddings(self, path, item_x_shard, num_items: int):
        """`item_embeddings.npy` of the reference (training.py:682-697: the encoded corpus next to the FAISS index): every rank
        encodes the items it owns (eval mode, augmentation added), the shards are gathered, rank 0 writes the [NI, D] fp32
        array.  Returns the full matrix on every rank.  COLLECTIVE."""
        import numpy as np
        local = self.eng.encode_all("item", item_x_shard)
        full = S.gather_rows_from_shards(local, num_items, self.group)
        if self.rank == 0 and path is not None:
            np.save(path, full.cpu().numpy())
        if self.world > 1:
            dist.barrier(group=self.group)
        return full

    def global_loss(self, loss: torch.Tensor) -> torch.Tensor:
        out = loss.clone()
        if self.world > 1:
            dist.all_reduce(out, group=self.group)
        return out


def build_sharded_engine(model_shard, *, group=None, static: bool = False, **engine_kwargs) -> ShardedEngine:
    """`model_shard`: a TwoTowerModel whose tables hold this rank's rows (num_embeddings = sharding.shard_size(...))."""
    from .engine import FusedEngine
    return ShardedEngine(FusedEngine(model_shard, **engine_kwargs), group=group, static=static)


class ShardedFlatIPIndex:
    """Item-sharded exact inner-product index.

    query_groups = 1 (default; BASELINE configs[2]: NI/W items per GPU): rank r holds `item_embeddings_shard` = rows r, r+W,
    r+2W, ... of the corpus (the training layout), or a contiguous block when `contiguous_offset` is given; every rank
    scores ALL queries against its shard.
    query_groups = G > 1: a G x S grid (W = G*S, `sharding.retrieval_grid`): rank g*S+s holds item shard s of S (same two
    layouts, with S in the place of W) and scores only query group g (Q/G queries); the merge runs inside the group's S
    ranks.  Same FLOPs per rank, but the per-query work that does not shrink with the shard (candidate lists, finalize:
    DESIGN 4.3) is paid for Q/G queries instead of Q - at the price of G copies of the corpus across the GPUs.
    Either way rank r returns query block r of W, bit-identical to the one-GPU result."""

    def __init__(self, item_embeddings_shard: torch.Tensor, *, group=None, dtype=torch.bfloat16, normalize: bool = False,
                 contiguous_offset: Optional[int] = None, query_groups: int = 1) -> None:
        from .retrieval import FlatIPIndex
        self.group = group
        self.world = S._world(group)
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.query_groups = int(query_groups)
        self.qgroup, self.shard, self.n_shards = S.retrieval_grid(self.rank, self.world, self.query_groups)
        self.merge_group = S.retrieval_subgroups(self.world, self.query_groups, group)[self.qgroup] if self.world > 1 else group
        self.contiguous_offset = contiguous_offset
        self.local = FlatIPIndex(item_embeddings_shard, normalize=normalize, dtype=dtype,
                                 id_offset=0 if contiguous_offset is None else int(contiguous_offset))

    def search(self, queries: torch.Tensor, k: int):
        """queries [Q, D]: the SAME Q queries on every rank (Q a multiple of W).  Returns (ids, scores) of this rank's
        query block [rank*Q/W, (rank+1)*Q/W) over the WHOLE corpus."""
        from . import functional as F
        Q = queries.shape[0]
        if Q % max(self.world, 1) != 0:
            raise ValueError(f"number of queries ({Q}) must be a multiple of the world size ({self.world})")
        if self.query_groups > 1:
            per = Q // self.query_groups
            queries = queries[self.qgroup * per:(self.qgroup + 1) * per]
        ids, scores = self.local.search(queries, k)
        if self.contiguous_offset is None and self.n_shards > 1:
            ids = torch.where(ids >= 0, ids * self.n_shards + self.shard, ids)   # local row -> global id (monotone: order kept)
        return S.merge_topk_shards(ids, scores, k, F.topk_merge, self.merge_group)
