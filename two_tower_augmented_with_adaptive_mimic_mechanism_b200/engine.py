"""Fused training step, evaluation loss, corpus encoder and retrieval for a `TwoTowerModel`.

This is the B200 replacement of the loop body of `_train_one_epoch` (reference training.py:726-831):
towers -> mimic -> sampled-negative BCE -> backward -> optimiser step, as a fixed sequence of libttam launches
on pre-allocated buffers.  No autograd, no dense table gradient: the duplicate gradient rows of the ID and
augmentation tables are segment-reduced and applied row-wise (SparseAdam for `sparse=True` tables,
lazy-exact AdamW/Adam/SGD for everything the reference hands to its dense optimiser).

The engine updates the model's own parameter storage in place, so `state_dict()` / checkpoints keep working.
"""
from __future__ import annotations

import os
import weakref
from dataclasses import dataclass
from typing import Optional

import torch

from . import functional as F
from .tower_ops import TowerPlan, plan_from_module, splits_backward, tower_backward, tower_forward


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


@dataclass
class _Table:
    """Optimiser state of one row-addressed table."""
    name: str
    weight: torch.Tensor
    mode: str                      # "sparse_adam" | "lazy"
    m: Optional[torch.Tensor] = None
    v: Optional[torch.Tensor] = None
    last_step: Optional[torch.Tensor] = None   # int32 [N], lazy tables only


class _Rebuild(Exception):
    """set_hyper: the change needs a freshly built engine (optimiser kind before the first step)."""


class FusedEngine:
    def __init__(self, model, *, optimizer: str = "adamw", lr: float = 1e-3, weight_decay: float = 0.0,
                 momentum: float = 0.0, dense_betas=(0.9, 0.999), sparse_betas=(0.9, 0.999), eps: float = 1e-8,
                 loss_weights: Optional[dict] = None, precision: str = "fp32", seed: int = 1234,
                 max_steps: int = 1 << 16, item_category_tensor=None, major_category_id=None, bag="auto",
                 loss: str = "sampled") -> None:
        self.model = model
        mimic = getattr(model, "adaptive_mimic", None)
        self.user: TowerPlan = plan_from_module(model.user_encoder, mimic.user_augmented if mimic is not None else None)
        self.item: TowerPlan = plan_from_module(model.item_encoder, mimic.item_augmented if mimic is not None else None)
        self.mimic = mimic is not None
        dev = self.user.table.device
        if dev.type != "cuda":
            raise F._lib.TtamError(f"model is on {dev}: the fused engine needs CUDA (sm_100a) parameters; there is no CPU path")
        F.lib()  # fail loudly now if libttam.so is missing
        self.device = dev
        self.kind = optimizer.lower()
        if self.kind not in ("adamw", "adam", "sgd"):
            raise ValueError(f"Unsupported optimizer: {optimizer}")
        self.lr, self.wd, self.momentum, self.eps = float(lr), float(weight_decay), float(momentum), float(eps)
        self.dense_betas, self.sparse_betas = tuple(dense_betas), tuple(sparse_betas)
        w = loss_weights or {}
        self.lambda_u = float(w.get("mimic_user", 0.0))
        self.lambda_i = float(w.get("mimic_item", 0.0))
        self.lambda_c = float(w.get("category_alignment", 0.0))
        self.cat_tensor, self.major = item_category_tensor, major_category_id
        self.precision = precision
        # "sampled": the reference's loss, BCE over 1 positive + N sampled negatives per sample (training.py:770-803).
        # "inbatch": softmax over the batch's own positives (BASELINE configs[1]; an extension, see ttam_inbatch_loss_fwd_bwd)
        if loss not in ("sampled", "inbatch"):
            raise ValueError("loss must be 'sampled' or 'inbatch'")
        self.loss_kind = loss
        # layer 1 of the feature encoders from the bag (CSR + dense tail) form of the feature matrices: "auto" = whenever
        # a matrix is sparse enough for the bag kernels (<= 64 non-zeros per row outside the dense tail), False = always
        # the dense GEMM (X[idx] . W1^T), True = like "auto" but a matrix that cannot be converted raises
        self.bag = bag if os.environ.get("TTAM_BAG", "1") != "0" else False      # TTAM_BAG=0: A/B switch for measurements
        self._bags: dict = {}
        self.seed = int(seed)
        self.max_steps = int(max_steps)
        self.t = 0                         # optimiser steps taken
        self.dirty = False                 # lazy tables hold rows that are behind `t`
        # ---- dense (small) parameters: one fused launch
        seen, self.dense = set(), []
        for p in self.user.dense_params() + self.item.dense_params():
            if id(p) not in seen:
                seen.add(id(p))
                self.dense.append(p)
        need_m = self.kind != "sgd" or self.momentum != 0.0
        # dense gradients live in ONE flat buffer (views handed to the wgrad kernels): a data-parallel all-reduce needs
        # no staging copies
        pad = lambda n: (n + 63) // 64 * 64            # every view starts on a 256-byte boundary
        self.dense_grad_flat = torch.zeros(sum(pad(p.numel()) for p in self.dense), dtype=torch.float32, device=dev)
        off = 0
        self._dense_grad_views = {}
        for p in self.dense:
            self._dense_grad_views[id(p)] = self.dense_grad_flat[off:off + p.numel()]
            off += pad(p.numel())
        self.dense_m = [torch.zeros_like(p) for p in self.dense] if need_m else None
        self.dense_v = [torch.zeros_like(p) for p in self.dense] if self.kind != "sgd" else None
        # ---- tables
        self.tables: dict[str, _Table] = {}
        for side, plan in (("user", self.user), ("item", self.item)):
            self._add_table(f"{side}_encoder.embedding.weight", plan.table, "sparse_adam" if plan.sparse else "lazy")
            if plan.aug is not None:
                self._add_table(f"adaptive_mimic.{side}_augmented.weight", plan.aug, "lazy")
        # ---- step state + bias-correction tables (device resident: the step is CUDA-graph replayable)
        self.state = F.new_step_state(dev, 0, 0)
        self._build_scalar_tables()
        self.bufs_u = self._grad_bufs(self.user)
        self.bufs_i = self._grad_bufs(self.item)
        self.misc: dict = {}
        self._graphs: dict = {}
        self._xpad: dict = {}
        self._side_stream = torch.cuda.Stream(device=dev)
        self._aug_stream = torch.cuda.Stream(device=dev)
        self._comm_stream = torch.cuda.Stream(device=dev)
        # the sort + lazy catch-up branch is a chain of small dependent kernels that the forward's last launch (o = t + A[idx])
        # waits for: high priority, so that its CTAs are placed ahead of the towers' persistent GEMM CTAs whenever an SM frees up
        # (device timeline of a step: the chain's kernels started 20-30 us late each, and the tower waited for them)
        prio = -1 if os.environ.get("TTAM_SORT_PRIORITY", "1") != "0" else 0
        self._sort_streams = {"u": torch.cuda.Stream(device=dev, priority=prio), "i": torch.cuda.Stream(device=dev, priority=prio)}
        self._overlap_sort = os.environ.get("TTAM_OVERLAP_SORT", "1") != "0"
        # o = t + A[idx] formed inside the loss kernel (ttam_loss_aug_fwd_bwd) instead of by a separate augment launch per tower
        self._fuse_aug_loss = os.environ.get("TTAM_FUSE_AUG_LOSS", "1") != "0"
        self._wgrad_streams = {"u": torch.cuda.Stream(device=dev), "i": torch.cuda.Stream(device=dev)}
        self._split_wgrad = os.environ.get("TTAM_WGRAD_STREAM", "1") != "0"
        self._use_aug_stream = os.environ.get("TTAM_AUG_STREAM", "1") != "0"
        self._interleave = os.environ.get("TTAM_INTERLEAVE_WGRAD", "1") != "0"

    # --------------------------------------------------------------------------------------------
    def _build_scalar_tables(self) -> None:
        self.scal_dense = F.adam_scalar_table(self.max_steps, self.lr, self.dense_betas, self.device)
        self.scal_sparse = F.adam_scalar_table(self.max_steps, self.lr, self.sparse_betas, self.device)

    def _ensure_steps(self, t: int) -> None:
        """The bias-correction tables are indexed by the optimiser step; they grow (doubling) instead of the engine
        refusing step max_steps + 1.  Kernels - and captured graphs - address the tables by pointer, so growing them
        invalidates every captured graph (re-captured on the next graph step)."""
        if t <= self.max_steps:
            return
        self.max_steps = max(2 * self.max_steps, t + 1024)
        torch.cuda.current_stream(self.device).synchronize()     # in-flight steps still read the old tables
        self._build_scalar_tables()
        F.note_realloc()

    def set_hyper(self, *, optimizer=None, lr=None, weight_decay=None, momentum=None, dense_betas=None, sparse_betas=None) -> bool:
        """Change optimiser hyper-parameters in place.  Returns True when anything changed.  The optimiser KIND can only
        change before the first step (the moment buffers depend on it)."""
        new = dict(kind=self.kind if optimizer is None else optimizer.lower(), lr=self.lr if lr is None else float(lr),
                   wd=self.wd if weight_decay is None else float(weight_decay),
                   momentum=self.momentum if momentum is None else float(momentum),
                   dense_betas=self.dense_betas if dense_betas is None else tuple(dense_betas),
                   sparse_betas=self.sparse_betas if sparse_betas is None else tuple(sparse_betas))
        old = dict(kind=self.kind, lr=self.lr, wd=self.wd, momentum=self.momentum, dense_betas=self.dense_betas,
                   sparse_betas=self.sparse_betas)
        if new == old:
            return False
        if new["kind"] != self.kind or (self.kind == "sgd" and (new["momentum"] != 0.0) != (self.momentum != 0.0)):
            if self.t > 0:
                raise ValueError("the optimiser kind cannot change after the first step: build a new engine")
            raise _Rebuild()
        if self.dirty:
            self.flush()          # rows that are behind were to be replayed with the OLD hyper-parameters
        self.lr, self.wd, self.momentum = new["lr"], new["wd"], new["momentum"]
        self.dense_betas, self.sparse_betas = new["dense_betas"], new["sparse_betas"]
        torch.cuda.current_stream(self.device).synchronize()
        self._build_scalar_tables()
        F.note_realloc()
        return True

    def _grad_bufs(self, plan: TowerPlan) -> dict:
        """Buffer dict of one tower, pre-seeded with the weight-gradient views of the flat dense-gradient buffer
        (tower_ops looks gradients up as dw<id> [shape of W] and db<id> [n, 1])."""
        bufs = {}
        for p in plan.dense_params():
            v = self._dense_grad_views[id(p)]
            if p.dim() == 2:
                bufs[f"dw{id(p)}"] = v.view(p.shape)
            else:
                bufs[f"db{id(p)}"] = v.view(p.shape[0], 1)
        return bufs

    def _add_table(self, name, weight, mode):
        t = _Table(name=name, weight=weight, mode=mode)
        if mode == "sparse_adam" or self.kind != "sgd":
            t.m, t.v = torch.zeros_like(weight), torch.zeros_like(weight)
        elif self.momentum != 0.0:
            t.m = torch.zeros_like(weight)
        if mode == "lazy":
            t.last_step = torch.zeros(weight.shape[0], dtype=torch.int32, device=weight.device)
        self.tables[name] = t

    def _x(self, X):
        """Feature matrix as the GEMM loaders want it.  The tensor-core path keeps ONE private copy per feature matrix:
        rows zero-padded to a multiple of 4 floats (F = 605 -> ld = 608: 16-byte aligned rows for the 16-byte LDGSTS)
        and values rounded to TF32 once, here, so that the GEMMs need no rounding pass over this operand.  The copy is
        reused for as long as the caller keeps passing the same tensor."""
        if X is None:
            return X
        bag = None
        if self.bag:
            bag = self._bag_for(X)
            if bag is None and self.bag is True:
                raise ValueError("bag=True: the feature matrix has rows with more than 64 non-zeros outside its dense tail")
            if bag is not None and (bag.wgrad or self.precision == "fp32"):
                X._ttam_bag = bag          # the bag form serves both directions (or the dense wgrad reads X as it is)
                return X
        if self.precision == "fp32":
            return X
        # keyed by the tensor OBJECT (weak reference) and its version counter: a new tensor that lands on a freed
        # matrix's address, or an in-place edit of the matrix, gets a fresh copy
        key = (X.data_ptr(), tuple(X.shape))
        hit = self._xpad.get(key)
        if hit is not None and (hit[0]() is not X or hit[1] != X._version):
            hit = None
        if hit is None:
            cp = F.round_tf32_(F.pad_cols(X, always_copy=True))   # private copy: padded rows, TF32-representable values
            cp._ttam_tf32 = True
            self._xpad = {k: v for k, v in self._xpad.items() if k[0] != key[0]}
            hit = (weakref.ref(X), X._version, cp)
            self._xpad[key] = hit
            F.note_realloc()       # graphs captured on the old copy are stale
        if bag is not None:
            hit[2]._ttam_bag = bag     # forward from the bag form, weight gradient from the padded TF32 copy
        return hit[2]

    def _bag_for(self, X):
        """BagMatrix of a feature matrix, built once per tensor object / version (None: too dense for the bag kernels)."""
        key = (X.data_ptr(), tuple(X.shape))
        hit = self._bags.get(key)
        if hit is not None and (hit[0]() is not X or hit[1] != X._version):
            hit = None
        if hit is None:
            hit = (weakref.ref(X), X._version, F.BagMatrix.build(X))
            self._bags = {k: v for k, v in self._bags.items() if k[0] != key[0]}
            self._bags[key] = hit
            F.note_realloc()
        return hit[2]

    def _categories(self):
        """(category tensor on the device, number of categories) - the count costs one host read, once per tensor."""
        key = (self.cat_tensor.data_ptr(), self.cat_tensor.numel())
        hit = self.__dict__.get("_cat_cache")
        if hit is None or hit[0] != key:
            cat = self.cat_tensor.to(self.device, torch.int64).contiguous()
            hit = (key, cat, int(cat.max()) + 1 if cat.numel() else 1)
            self._cat_cache = hit
        return hit[1], hit[2]

    def _misc(self, name, shape, dtype):
        t = self.misc.get(name)
        if t is None or t.shape[0] < shape[0] or t.shape[1:] != tuple(shape[1:]) or t.dtype != dtype:
            rows = shape[0] if shape[0] < 4096 else (int(shape[0] * 1.125) + 1023) // 1024 * 1024   # headroom: see tower_ops._buf
            if name in self.misc:
                F.note_realloc()
            t = torch.empty((rows,) + tuple(shape[1:]), dtype=dtype, device=self.device)
            self.misc[name] = t
        return t[: shape[0]]

    # --------------------------------------------------------------------------------------------
    def _sort(self, idx, tag, num_rows):
        """Stable sort of the touched rows of one index set (once per step; shared by every table it addresses)."""
        R = idx.numel()
        sorted_idx = self._misc(f"sorted_{tag}", (R,), torch.int64)
        perm = self._misc(f"perm_{tag}", (R,), torch.int32)
        F.sort_rows(idx, num_rows, sorted_idx=sorted_idx, perm=perm)
        long_list = self._misc(f"long_{tag}", (F.lib().ttam_long_segments_bytes(R) // 4,), torch.int32)
        F.find_long_segments(sorted_idx, out=long_list)
        return sorted_idx, perm, long_list

    def _lazy_kw(self):
        return dict(scalars=self.scal_dense, lr=self.lr, weight_decay=self.wd, betas=self.dense_betas, eps=self.eps,
                    momentum=self.momentum, step=self.t, state=self.state)

    def _catchup(self, tab: _Table, sorted_idx):
        """Lazy tables: replay the zero-gradient steps of the rows this step reads, before the forward reads them."""
        if tab.mode == "lazy":
            F.lazy_catchup(self.kind, tab.weight, tab.m, tab.v, tab.last_step, sorted_idx, **self._lazy_kw())

    def _update_table(self, tab: _Table, sort, grad_a, grad_b=None):
        sorted_idx, perm, long_list = sort
        if tab.mode == "sparse_adam":
            F.sparse_adam_rows(tab.weight, tab.m, tab.v, sorted_idx, perm, grad_a, grad_b, lr=self.lr,
                               betas=self.sparse_betas, eps=self.eps, step=self.t, scalars=self.scal_sparse,
                               state=self.state, long_list=long_list)
        else:
            F.lazy_rows(self.kind, tab.weight, tab.m, tab.v, tab.last_step, sorted_idx, perm, grad_a, grad_b,
                        long_list=long_list, **self._lazy_kw())

    # ---- the step, in three phases (the row-sharded engine runs phases 1 and 3 on the rows a rank OWNS and phase 2
    # on the samples it was GIVEN, with an all-to-all in between; on one GPU they run back to back) -----------------
    # The user tower and the item tower are independent until the loss, and again from the loss to the dense optimiser:
    # the user-side launches (8192 rows: 64 GEMM tiles for 148 SMs) run on a side stream next to the item-side ones
    # (49 152 rows) instead of in front of them.  Under CUDA-graph capture this becomes two parallel branches.
    def _fork(self, which: str = "side"):
        side = self._side_stream if which == "side" else self._aug_stream
        side.wait_stream(torch.cuda.current_stream(self.device))
        return side

    def _join(self, side) -> None:
        torch.cuda.current_stream(self.device).wait_stream(side)

    def _tower_side(self, side: str, plan: TowerPlan, idx, X, bufs, rng_base, fused_loss: bool = False):
        """Sort + lazy catch-up + tower forward of one side, on the current stream.
        Only the augmentation add (o = t + A[idx]) reads a lazily-updated table when the ID table is a SparseAdam one, so
        the sort of the step's row ids and the catch-up of A run on their own stream NEXT TO the gather + MLP + gate of
        the tower and are joined just before the add (~50 us off the item side's critical path)."""
        T, tag = self.tables, side[0]
        e_tab = T[f"{side}_encoder.embedding.weight"]
        a_tab = T.get(f"adaptive_mimic.{side}_augmented.weight")
        kw = dict(gather=True, train=True, bufs=bufs, seed=self.seed, rng_base=rng_base, state=self.state,
                  precision=self.precision, want_q=self.mimic)
        if not (self._overlap_sort and a_tab is not None and plan.aug is not None and e_tab.mode != "lazy"):
            sort = self._sort(idx, tag, plan.table.shape[0])
            for tab in (e_tab, a_tab):
                if tab is not None:
                    self._catchup(tab, sort[0])
            return sort, tower_forward(plan, idx, X, **kw)
        cur = torch.cuda.current_stream(self.device)
        ss = self._sort_streams[tag]
        ss.wait_stream(cur)
        with torch.cuda.stream(ss), F.ws_scope(side + "_sort"):
            sort = self._sort(idx, tag, plan.table.shape[0])
            self._catchup(a_tab, sort[0])
        c = tower_forward(plan, idx, X, augment=False, **kw)
        cur.wait_stream(ss)
        if fused_loss:      # the loss kernel gathers A[idx] itself (after the catch-up joined just above)
            c.o = c.q = None
            return sort, c
        R = idx.numel()
        from .tower_ops import _buf
        o = _buf(bufs, "o", (R, plan.out_dim), self.device)
        q = _buf(bufs, "q", (R, plan.out_dim), self.device) if self.mimic else None
        F.augment_fwd(c.t, plan.aug, idx, out=o, q_out=q)
        c.o, c.q = o, q
        return sort, c

    def _forward_phase(self, users, items, Xu, Xi, fused_loss: bool = False):
        """Sort + lazy catch-up + both towers for the rows `users` / `items` (row ids of THIS engine's tables).
        fused_loss: the caller's loss kernel adds the augmentation rows itself (cu.o / ci.o stay None)."""
        Xu, Xi = self._x(Xu), self._x(Xi)
        F.advance_step(self.state, rng_stride=1 << 36)
        side = self._fork()
        with torch.cuda.stream(side), F.ws_scope("user"):
            sort_u, cu = self._tower_side("user", self.user, users, Xu, self.bufs_u, 0, fused_loss)
        with F.ws_scope("item"):
            sort_i, ci = self._tower_side("item", self.item, items, Xi, self.bufs_i, 1 << 35, fused_loss)
        self._join(side)
        return dict(sort_u=sort_u, sort_i=sort_i, cu=cu, ci=ci)

    def _loss_phase(self, o_u, o_i, t_u, t_p, q_u, q_p, items, B, N, batch_fraction=1.0):
        """Fused loss forward + backward on one (local) batch.  Returns loss[4] and the gradient row blocks."""
        D = o_u.shape[1]
        loss = self._misc("loss", (4,), torch.float32)
        do_u = self._misc("do_u", (B, D), torch.float32)
        do_i = self._misc("do_i", (B * (1 + N), D), torch.float32)
        dq_u = self._misc("dq_u", (B, D), torch.float32) if self.mimic else None
        dq_p = self._misc("dq_p", (B, D), torch.float32) if self.mimic else None
        if self.mimic:
            F.loss_fwd_bwd(o_u, o_i, t_u=t_u, t_p=t_p, q_u=q_u, q_p=q_p, lambda_u=self.lambda_u,
                           lambda_i=self.lambda_i, out=(loss, do_u, do_i, dq_u, dq_p), batch_fraction=batch_fraction)
        else:
            F.loss_fwd_bwd(o_u, o_i, out=(loss, do_u, do_i, None, None), batch_fraction=batch_fraction)
        if self.lambda_c > 0 and self.cat_tensor is not None and self.major is not None:
            if batch_fraction != 1.0:
                raise NotImplementedError("category-alignment loss is not available with a sharded batch")
            cat, n_cat = self._categories()
            F.category_alignment(items, o_i, cat, n_cat, int(self.major), lambda_c=self.lambda_c, loss_out=loss,
                                 grad_a=do_i, grad_b=dq_p if self.mimic else None, B=B)
        return loss, do_u, do_i, dq_u, dq_p

    def _can_fuse_aug_loss(self, N: int) -> bool:
        """The augmentation add can move into the loss kernel when both towers take the overlapped-sort branch of _tower_side
        (their add is a separate launch there), the loss is the sampled-negative BCE without the category term (that one reads
        o_i), and the shape is one the fused kernel covers."""
        if not (self._fuse_aug_loss and self._overlap_sort and self.loss_kind != "inbatch" and self.mimic and N >= 1):
            return False
        if self.lambda_c > 0 and self.cat_tensor is not None and self.major is not None:
            return False
        T = self.tables
        for side, plan in (("user", self.user), ("item", self.item)):
            a_tab = T.get(f"adaptive_mimic.{side}_augmented.weight")
            if a_tab is None or plan.aug is None or T[f"{side}_encoder.embedding.weight"].mode == "lazy" or plan.D != plan.out_dim:
                return False
            if not plan.aug.is_contiguous():
                return False
        return F.loss_aug_supported(N, self.user.out_dim)

    def _loss_aug_phase(self, t_u, t_i, users, items, B, N):
        """Loss forward + backward with o = t + A[idx] formed inside the kernel (no o / q buffers, no augment launches)."""
        D = t_u.shape[1]
        loss = self._misc("loss", (4,), torch.float32)
        do_u = self._misc("do_u", (B, D), torch.float32)
        do_i = self._misc("do_i", (B * (1 + N), D), torch.float32)
        dq_u = self._misc("dq_u", (B, D), torch.float32)
        dq_p = self._misc("dq_p", (B, D), torch.float32)
        F.loss_aug_fwd_bwd(t_u, t_i, self.user.aug, self.item.aug, users, items, mimic=True, lambda_u=self.lambda_u,
                           lambda_i=self.lambda_i, out=(loss, do_u, do_i, dq_u, dq_p))
        return loss, do_u, do_i, dq_u, dq_p

    def _inbatch_phase(self, o_u, o_p, t_u, t_p, q_u, q_p, B):
        """In-batch softmax loss forward + backward on the B (user, positive) pairs of the step."""
        D = o_u.shape[1]
        loss = self._misc("loss", (4,), torch.float32)
        do_u, do_i = self._misc("do_u", (B, D), torch.float32), self._misc("do_i", (B, D), torch.float32)
        dq_u = self._misc("dq_u", (B, D), torch.float32) if self.mimic else None
        dq_p = self._misc("dq_p", (B, D), torch.float32) if self.mimic else None
        kw = dict(t_u=t_u, t_p=t_p, q_u=q_u, q_p=q_p, lambda_u=self.lambda_u, lambda_i=self.lambda_i) if self.mimic else {}
        F.inbatch_loss_fwd_bwd(o_u, o_p, out=(loss, do_u, do_i, dq_u, dq_p), precision=self.precision, **kw)
        return loss, do_u, do_i, dq_u, dq_p

    def _backward_phase(self, ctx, do_u, do_i, dq_user, dq_item, dense_grad_hook=None):
        """Tower backward for the rows of `ctx`, row-wise table updates, dense optimiser.
        dq_user / dq_item: gradient rows of the augmentation tables, each a tensor or a (first rows, other rows) pair.
        dense_grad_hook(list of gradient tensors): called before the dense optimiser (data-parallel all-reduce)."""
        T = self.tables
        sort_u, sort_i, cu, ci = ctx["sort_u"], ctx["sort_i"], ctx["cu"], ctx["ci"]
        grads: dict = {}
        grads_u: dict = {}
        pair = lambda g: g if isinstance(g, tuple) else (g, None)
        shared = {id(p) for p in self.user.dense_params()} & {id(p) for p in self.item.dense_params()}
        if shared:   # towers that share dense weights accumulate into the same gradient buffers: keep them in order
            side = None
        else:
            side = self._fork()
        # ---- row-wise optimisers of the augmentation tables: their gradient rows came out of the loss kernel, nothing
        # in the tower backward feeds them -> a third branch
        aug = self._fork("aug") if (self.mimic and self._use_aug_stream) else None
        if aug is not None:
            with torch.cuda.stream(aug), F.ws_scope("aug"):
                self._update_table(T["adaptive_mimic.user_augmented.weight"], sort_u, *pair(dq_user))
                self._update_table(T["adaptive_mimic.item_augmented.weight"], sort_i, *pair(dq_item))
        # ---- tower backward + row-wise optimisers of the ID tables (no dense table gradient), user side next to item side.
        # With a dense_grad_hook (data-parallel all-reduce of the weight gradients) the hook runs on its own stream as soon
        # as BOTH towers' weight gradients exist, next to the row-wise table updates instead of after them.
        # The weight gradients of a tower feed nothing but the dense optimiser at the very end: they run on their own
        # stream (phase 2 of the composite backward) next to the row-wise update of the tower's ID table, which only
        # waits for the data-gradient chain (phase 1).
        cur = torch.cuda.current_stream(self.device)
        overlap = dense_grad_hook is not None and side is not None
        split = self._split_wgrad and side is not None and splits_backward(cu) and splits_backward(ci)
        wg_u, wg_i = self._wgrad_streams["u"], self._wgrad_streams["i"]
        with (torch.cuda.stream(side) if side is not None else _null()), F.ws_scope("user"):
            de_u = tower_backward(self.user, cu, do_u, grads_u if side is not None else grads, bufs=self.bufs_u, state=self.state,
                                  precision=self.precision, phase=1 if split else 0, wgrad_stream=wg_u if (split and self._interleave) else None)
            if split:
                wg_u.wait_stream(side)
                with torch.cuda.stream(wg_u), F.ws_scope("user_wgrad"):
                    tower_backward(self.user, cu, do_u, grads_u, bufs=self.bufs_u, state=self.state, precision=self.precision, phase=2)
                if overlap:
                    self._comm_stream.wait_stream(wg_u)
            elif overlap:
                self._comm_stream.wait_stream(side)
            self._update_table(T["user_encoder.embedding.weight"], sort_u, de_u)
            if split:
                side.wait_stream(wg_u)
        with F.ws_scope("item"):
            de_i = tower_backward(self.item, ci, do_i, grads, bufs=self.bufs_i, state=self.state, precision=self.precision,
                                  phase=1 if split else 0, wgrad_stream=wg_i if (split and self._interleave) else None)
            if split:
                wg_i.wait_stream(cur)
                with torch.cuda.stream(wg_i), F.ws_scope("item_wgrad"):
                    tower_backward(self.item, ci, do_i, grads, bufs=self.bufs_i, state=self.state, precision=self.precision, phase=2)
                if overlap:
                    self._comm_stream.wait_stream(wg_i)
            elif overlap:
                self._comm_stream.wait_stream(cur)
            if not overlap:
                self._update_table(T["item_encoder.embedding.weight"], sort_i, de_i)
        if side is not None:
            grads.update(grads_u)
        # ---- dense optimiser inputs: the MLP / gate / projection tensors that received a gradient
        ps, gs, ms, vs = [], [], [], []
        for j, p in enumerate(self.dense):
            g = grads.get(id(p))
            if g is None:
                continue
            ps.append(p.data); gs.append(g.view(p.shape))
            if self.dense_m is not None:
                ms.append(self.dense_m[j])
            if self.dense_v is not None:
                vs.append(self.dense_v[j])
        if overlap:
            if ps:
                with torch.cuda.stream(self._comm_stream):
                    dense_grad_hook(gs)
            with F.ws_scope("item"):
                self._update_table(T["item_encoder.embedding.weight"], sort_i, de_i)
            cur.wait_stream(self._comm_stream)
        if split:
            cur.wait_stream(wg_i)
        if side is not None:
            self._join(side)
        if aug is not None:
            self._join(aug)
        elif self.mimic:
            self._update_table(T["adaptive_mimic.user_augmented.weight"], sort_u, *pair(dq_user))
            self._update_table(T["adaptive_mimic.item_augmented.weight"], sort_i, *pair(dq_item))
        if ps:
            if dense_grad_hook is not None and not overlap:
                dense_grad_hook(gs)
            F.dense_step(self.kind, ps, gs, ms if self.dense_m is not None else None,
                         vs if self.dense_v is not None else None, lr=self.lr, weight_decay=self.wd,
                         betas=self.dense_betas, eps=self.eps, momentum=self.momentum, step=self.t,
                         scalars=self.scal_dense, state=self.state)

    def _step_body(self, users, items, B, N, Xu, Xi):
        launches0 = F.lib().ttam_launch_count()
        fused = self._can_fuse_aug_loss(N)
        ctx = self._forward_phase(users, items, Xu, Xi, fused_loss=fused)
        cu, ci = ctx["cu"], ctx["ci"]
        if fused and cu.o is None and ci.o is None:
            loss, do_u, do_i, dq_u, dq_p = self._loss_aug_phase(cu.t, ci.t, users, items, B, N)
            self._backward_phase(ctx, do_u, do_i, dq_u, (dq_p, do_i[B:]) if self.mimic else None)
        elif self.loss_kind == "inbatch":
            loss, do_u, do_i, dq_u, dq_p = self._inbatch_phase(cu.o, ci.o, cu.t, ci.t, cu.q, ci.q, B)
            self._backward_phase(ctx, do_u, do_i, dq_u, dq_p)
        else:
            if self.mimic:
                loss, do_u, do_i, dq_u, dq_p = self._loss_phase(cu.o, ci.o, cu.t, ci.t[:B], cu.q, ci.q[:B], items, B, N)
            else:
                loss, do_u, do_i, dq_u, dq_p = self._loss_phase(cu.o, ci.o, None, None, None, None, items, B, N)
            self._backward_phase(ctx, do_u, do_i, dq_u, (dq_p, do_i[B:]) if self.mimic else None)
        self.launches_per_step = F.lib().ttam_launch_count() - launches0
        return loss

    @staticmethod
    def launch_count() -> int:
        """libttam kernel launches issued so far by this process (host counter; graph replays are not counted)."""
        return int(F.lib().ttam_launch_count())

    def begin_step(self) -> None:
        """Advance the host-side step counter (the device-side one advances inside the forward phase)."""
        self._ensure_steps(self.t + 1)
        self.t += 1
        self.dirty = True

    @torch.no_grad()
    def train_step(self, users: torch.Tensor, pos: torch.Tensor, neg: torch.Tensor, user_x, item_x, *,
                   graph: bool = False) -> torch.Tensor:
        """One optimisation step on a batch (users [B], pos [B], neg [B,N], all int64 on the device).
        Returns a device tensor loss[4] = {total, bce | ce, mimic_user, mimic_item} (valid until the next step).
        loss="inbatch": `neg` is ignored (may be None) - the other positives of the batch are the negatives."""
        if self.loss_kind == "inbatch":
            neg = None
        elif neg is None:
            raise ValueError("the sampled-negative loss needs neg [B, N]")
        B, N = (users.shape[0], 0) if neg is None else neg.shape
        self._ensure_steps(self.t + 1)
        if graph:
            return self._graph_step(users, pos, neg, user_x, item_x)
        items = self._misc("items", (B * (1 + N),), torch.int64)
        items[:B].copy_(pos)
        if N:
            items[B:].copy_(neg.reshape(-1))
        users = users.contiguous()
        self.t += 1
        self.dirty = True
        return self._step_body(users, items, B, N, user_x, item_x)

    def _graph_step(self, users, pos, neg, user_x, item_x):
        B, N = (users.shape[0], 0) if neg is None else neg.shape
        key = (B, N, None if user_x is None else user_x.data_ptr(), None if item_x is None else item_x.data_ptr())
        entry = self._graphs.get(key)
        if entry is not None and entry[4] != F.alloc_generation():
            # a buffer / workspace / scalar table the captured launches address has been reallocated since: every graph of
            # this engine may hold dangling pointers
            self._graphs.clear()
            entry = None
        if entry is None:
            su = torch.empty(B, dtype=torch.int64, device=self.device)
            si = torch.empty(B * (1 + N), dtype=torch.int64, device=self.device)
            su.copy_(users); si[:B].copy_(pos)
            if N:
                si[B:].copy_(neg.reshape(-1))
            # warm-up outside capture sizes every buffer; run it on copies of nothing: it IS a real step
            self.t += 1
            loss = self._step_body(su, si, B, N, user_x, item_x)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            # capture advances the device step once without executing: compensate afterwards
            with torch.cuda.graph(g):
                loss = self._step_body(su, si, B, N, user_x, item_x)
            entry = (g, su, si, loss, F.alloc_generation())
            self._graphs[key] = entry
            self.dirty = True
            return loss
        g, su, si, loss, _ = entry
        su.copy_(users); si[:B].copy_(pos)
        if N:
            si[B:].copy_(neg.reshape(-1))
        self.t += 1
        self.dirty = True
        g.replay()
        return loss

    # --------------------------------------------------------------------------------------------
    @torch.no_grad()
    def flush(self) -> None:
        """Bring every lazily-updated table row up to the current step (before any full-table read)."""
        if not self.dirty or self.t == 0:
            return
        for tab in self.tables.values():
            if tab.mode == "lazy":
                F.lazy_flush(self.kind, tab.weight, tab.m, tab.v, tab.last_step, scalars=self.scal_dense, lr=self.lr,
                             weight_decay=self.wd, betas=self.dense_betas, eps=self.eps, momentum=self.momentum,
                             step=self.t)
        self.dirty = False

    @torch.no_grad()
    def encode(self, side: str, idx: torch.Tensor, X, *, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Eval-mode tower + augmentation for rows `idx` (reference training.py:613-643, 1019-1026)."""
        self.flush()
        plan = self.user if side == "user" else self.item
        bufs = self.misc.setdefault(f"enc_{side}", {})
        c = tower_forward(plan, idx.contiguous(), self._x(X), gather=True, train=False, bufs=bufs, precision=self.precision)
        if out is None:
            return c.o.clone()
        out.copy_(c.o)
        return out

    @torch.no_grad()
    def encode_all(self, side: str, X, chunk: int = 65536) -> torch.Tensor:
        plan = self.user if side == "user" else self.item
        n = plan.table.shape[0]
        out = torch.empty((n, plan.out_dim), dtype=torch.float32, device=self.device)
        for s in range(0, n, chunk):
            e = min(n, s + chunk)
            self.encode(side, torch.arange(s, e, device=self.device, dtype=torch.int64), X, out=out[s:e])
        return out

    def _param_version(self) -> tuple:
        """Changes whenever the parameters may have: our own steps (t) and torch-side writes (load_state_dict bumps the
        tensors' version counters; our kernels write through raw pointers and do not)."""
        return (self.t, tuple(p._version for p in self.model.parameters()))

    @torch.no_grad()
    def corpus(self, item_x) -> torch.Tensor:
        """Eval-mode embeddings of the whole item corpus [NI, D] (reference `_encode_item_embeddings`, training.py:613-643),
        encoded once per parameter version: the validation and test evaluations of an epoch, `_score_all_items_for_user`
        and the final index all share it (the reference re-encodes the corpus for each of them)."""
        key = (self._param_version(), None if item_x is None else (item_x.data_ptr(), item_x._version), bool(self.model.training))
        hit = self.__dict__.get("_corpus_cache")
        if hit is None or hit[0] != key:
            hit = (key, self.encode_all("item", item_x), {})
            self._corpus_cache = hit
        return hit[1]

    @torch.no_grad()
    def corpus_index(self, item_x, *, normalize: bool = False, dtype=torch.float32):
        """FlatIPIndex over `corpus(item_x)`, cached next to it."""
        from .retrieval import FlatIPIndex
        self.corpus(item_x)
        cache = self._corpus_cache[2]
        k = (bool(normalize), dtype)
        if k not in cache:
            cache[k] = FlatIPIndex(self._corpus_cache[1], normalize=normalize, dtype=dtype)
        return cache[k]

    @torch.no_grad()
    def eval_loss(self, users, pos, neg, user_x, item_x) -> torch.Tensor:
        """`_compute_loss` batch body (reference training.py:862-911): BCE on eval-mode augmented embeddings."""
        B, N = neg.shape
        items = torch.cat([pos.reshape(-1), neg.reshape(-1)])
        o_u = self.encode("user", users, user_x)
        o_i = self.encode("item", items, item_x)
        loss, *_ = F.loss_fwd_bwd(o_u, o_i, backward=False)
        return loss

    # --------------------------------------------------------------------------------------------
    @torch.no_grad()
    def load_optimizer_state(self, state: dict, step: int) -> None:
        """Inverse of `optimizer_state`: {name: {exp_avg, exp_avg_sq}} (this engine's shapes; missing entries stay zero) and
        the number of optimiser steps taken.  Every row of a lazily-updated table counts as up to date at `step` (what
        `optimizer_state` exported was flushed).  Resuming continues the bias corrections and the dropout counters from there."""
        step = int(step)
        self._ensure_steps(step + 1)
        names = {id(p): n for n, p in self.model.named_parameters()}
        slots = {name: (tab.m, tab.v) for name, tab in self.tables.items()}
        for j, p in enumerate(self.dense):
            slots[names.get(id(p), f"dense{j}")] = (None if self.dense_m is None else self.dense_m[j],
                                                    None if self.dense_v is None else self.dense_v[j])
        for name, ent in state.items():
            if name not in slots:
                raise KeyError(f"optimizer state for an unknown parameter: {name}")
            for dst, key in zip(slots[name], ("exp_avg", "exp_avg_sq")):
                src = ent.get(key)
                if dst is not None and src is not None:
                    dst.copy_(torch.as_tensor(src).to(dst.device, dst.dtype).view_as(dst))
        for tab in self.tables.values():
            if tab.last_step is not None:
                tab.last_step.fill_(step)
        self.t = step
        self.dirty = False
        self.state.copy_(F.new_step_state(self.device, step, step << 36))      # _forward_phase advances the counter by 1 << 36 a step
        self._graphs.clear()

    def optimizer_state(self) -> dict:
        """Per-parameter optimiser state in torch.optim layout ({name: {step, exp_avg, exp_avg_sq}})."""
        self.flush()
        out = {}
        for name, tab in self.tables.items():
            out[name] = {"step": self.t, "exp_avg": tab.m, "exp_avg_sq": tab.v}
        names = {id(p): n for n, p in self.model.named_parameters()}
        for j, p in enumerate(self.dense):
            out[names.get(id(p), f"dense{j}")] = {
                "step": self.t,
                "exp_avg": None if self.dense_m is None else self.dense_m[j],
                "exp_avg_sq": None if self.dense_v is None else self.dense_v[j]}
        return out
