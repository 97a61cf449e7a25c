"""Tower arithmetic on top of the C ABI: forward / backward of one tower as sequences of libttam launches.

Two users:
  * the nn.Module API of `models.py` (autograd.Function wrappers at the bottom): generic, allocates;
  * `engine.FusedEngine`: the fused training step / corpus encoder (no autograd, no dense table gradients).

Math follows SURVEY.md Appendix A (reference encoders.py:102-168,221-255; adaptive_mimic.py:59-105).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Optional

import torch
from torch import nn

from . import functional as F


# ------------------------------------------------------------------------------------------------
# plan: what a tower is, as tensors
# ------------------------------------------------------------------------------------------------
@dataclass
class TowerPlan:
    table: torch.Tensor                       # [N, D] ID embedding
    sparse: bool
    fe_kind: str = "none"                     # none | identity | linear | mlp
    fe_layers: list = field(default_factory=list)   # [(W [out,in], b [out])] in order
    activation: str = "relu"
    dropout: float = 0.0
    fusion: str = "identity"                  # identity | sum | concat | gated
    gate: Optional[tuple] = None              # (G1, c1, G2, c2)
    proj: Optional[tuple] = None              # (Wp, bp)
    aug: Optional[torch.Tensor] = None        # [N, D] augmentation table of the mimic mechanism

    @property
    def D(self) -> int:
        return self.table.shape[1]

    @property
    def out_dim(self) -> int:
        return self.proj[0].shape[0] if self.fusion == "concat" else self.D

    def dense_params(self) -> list:
        ps = [t for wb in self.fe_layers for t in wb]
        if self.fusion == "gated":
            ps += list(self.gate)
        if self.fusion == "concat":
            ps += list(self.proj)
        return ps


def plan_from_module(tower, aug_table: Optional[nn.Embedding] = None) -> TowerPlan:
    emb = tower.embedding
    if emb.padding_idx is not None or emb.max_norm is not None:
        raise ValueError("padding_idx / max_norm are not supported by the B200 tower kernels")
    plan = TowerPlan(table=emb.weight, sparse=bool(emb.sparse))
    fe = tower.feature_encoder
    plan.fusion = tower.fusion if fe is not None else "identity"
    if fe is not None and plan.fusion != "identity":
        kind = getattr(fe, "kind", None)
        if kind is None or kind == "custom":   # a reference-built wrapper: infer from the network
            net = fe.network
            kind = "identity" if isinstance(net, nn.Identity) else "linear" if isinstance(net, nn.Linear) else "mlp"
        plan.fe_kind = kind
        if kind != "identity":
            lins = [m for m in ([fe.network] if isinstance(fe.network, nn.Linear) else fe.network) if isinstance(m, nn.Linear)]
            plan.fe_layers = [(m.weight, m.bias) for m in lins]
        if kind == "mlp":
            plan.activation = _activation_name(fe.network)
            plan.dropout = next((float(m.p) for m in fe.network if isinstance(m, nn.Dropout)), 0.0)
    if plan.fusion == "gated":
        if tower.adaptive_mimic is None:
            raise ValueError("Adaptive mimic fusion requires a mimic module.")
        g = tower.adaptive_mimic.gate_network
        plan.gate = (g[0].weight, g[0].bias, g[2].weight, g[2].bias)
    if plan.fusion == "concat":
        plan.proj = (tower.projection.weight, tower.projection.bias)
    if aug_table is not None:
        plan.aug = aug_table.weight
    return plan


def _activation_name(seq) -> str:
    for m in seq:
        for name, cls in (("relu", nn.ReLU), ("gelu", nn.GELU), ("tanh", nn.Tanh), ("selu", nn.SELU)):
            if isinstance(m, cls):
                return name
    return "relu"


# ------------------------------------------------------------------------------------------------
# forward / backward on raw tensors
# ------------------------------------------------------------------------------------------------
class Cache(dict):
    __getattr__ = dict.get


def _headroom(rows: int) -> int:
    return (int(rows * 1.125) + 1023) // 1024 * 1024


def _buf(bufs, name, shape, device):
    """Pre-allocated buffer `name` (engine) or a fresh tensor (module API)."""
    if bufs is not None:
        t = bufs.get(name)
        if t is None or t.shape[0] < shape[0] or t.shape[1:] != tuple(shape[1:]):
            # grow with headroom: the row count of a row-sharded step changes a little from step to step, and every
            # reallocation is a cudaMalloc + synchronisation
            rows = shape[0] if t is None and shape[0] < 4096 else _headroom(shape[0])
            if t is not None:
                F.note_realloc()       # captured graphs that address the old buffer are stale from here on
            t = torch.empty((rows,) + tuple(shape[1:]), dtype=torch.float32, device=device)
            bufs[name] = t
        return t[: shape[0]]
    return torch.empty(shape, dtype=torch.float32, device=device)


def _w1_rounded(W: torch.Tensor, bufs) -> torch.Tensor:
    """Layer-1 weight of the composite tensor-core path: padded copy (16-byte aligned rows), rounded to TF32 here so
    that the GEMM skips its rounding pass (refreshed every call: the optimiser rewrites the weight each step)."""
    key = f"wpad{id(W)}"
    view = F.pad_cols(W, out=bufs.get(key), always_copy=True)
    bufs[key] = view._base if view._base is not None else view
    return F.round_tf32_(view)


def _w_for(W: torch.Tensor, bufs, precision: str) -> torch.Tensor:
    """Weight operand of a forward GEMM: the tensor-core path wants 16-byte aligned rows, so a weight whose row length
    is not a multiple of 4 floats (the 605-wide layer 1) is copied into a padded buffer (refreshed on every call: the
    optimiser rewrites the weight each step; 0.5 MB)."""
    if precision == "fp32" or W.shape[1] % 4 == 0:
        return W
    if bufs is None:
        return F.pad_cols(W)
    key = f"wpad{id(W)}"
    view = F.pad_cols(W, out=bufs.get(key))
    bufs[key] = view._base if view._base is not None else view
    return view


# ------------------------------------------------------------------------------------------------
# composite path: the whole gated / 1-hidden-layer-ReLU tower as ONE C call per direction (ttam_tower_fwd / _bwd)
# ------------------------------------------------------------------------------------------------
def _composite_ok(plan: TowerPlan, X, gather: bool, bufs) -> bool:
    return (bufs is not None and gather and X is not None and plan.fusion == "gated" and plan.fe_kind == "mlp"
            and len(plan.fe_layers) == 2 and plan.activation == "relu" and plan.gate is not None
            and plan.fe_layers[1][0].shape[0] == plan.D)


def _bag_of(X, H: int):
    """The bag form attached to a feature matrix by the engine (`FusedEngine._x`), when the bag kernels take this layer."""
    bag = getattr(X, "_ttam_bag", None) if X is not None else None
    if bag is not None and F.bag_supported(H, bag.shape[1], bag.T, bag.max_nnz):
        return bag
    return None


def _tower_desc(plan: TowerPlan, X, W1, p_drop, seed, rng_base, state, precision, augment=True, bag=None):
    (_, b1), (W2, b2) = plan.fe_layers
    G1, c1, G2, c2 = plan.gate
    d = F._lib.TowerDesc()
    d.table, d.table_rows, d.D = plan.table.data_ptr(), plan.table.shape[0], plan.D
    d.aug = plan.aug.data_ptr() if (plan.aug is not None and augment) else None
    d.X, d.ldx, d.F = X.data_ptr(), X.stride(0), X.shape[1]
    if bag is not None:
        d.bag_rowptr, d.bag_entries = bag.rowptr.data_ptr(), bag.entries.data_ptr()
        d.bag_tail = None if bag.tail is None else bag.tail.data_ptr()
        d.bag_T, d.bag_tail_start, d.bag_max_nnz = bag.T, bag.tail_start, bag.max_nnz
        d.bag_wgrad = 1 if bag.wgrad else 0
        ws = F.workspace(F.lib().ttam_bag_linear_workspace_bytes(0, W1.shape[0], X.shape[1]), X.device, "bag_fwd")
        d.bag_scratch, d.bag_scratch_bytes = ws.data_ptr(), ws.numel()
    d.W1, d.ldw1, d.b1, d.H = W1.data_ptr(), W1.stride(0), b1.data_ptr(), W1.shape[0]
    d.W2, d.b2 = W2.data_ptr(), b2.data_ptr()
    d.G1, d.c1, d.Hg, d.G2, d.c2 = G1.data_ptr(), c1.data_ptr(), G1.shape[0], G2.data_ptr(), c2.data_ptr()
    d.dropout_p, d.precision, d.seed, d.rng_base = float(p_drop), F.PREC[precision], int(seed), int(rng_base)
    d.state = None if state is None else state.data_ptr()
    return d


def _tower_forward_composite(plan, idx, X, *, train, bufs, seed, rng_base, state, precision, want_q, augment) -> Cache:
    R, D, dev = idx.numel(), plan.D, plan.table.device
    H, Hg = plan.fe_layers[0][0].shape[0], plan.gate[0].shape[0]
    p_drop = plan.dropout if train else 0.0
    bag = _bag_of(X, H)
    pre = precision != "fp32" and bool(getattr(X, "_ttam_tf32", False))   # the engine's private rounded copy of X (dense layer 1 / wgrad)
    W1 = _w1_rounded(plan.fe_layers[0][0], bufs) if (precision != "fp32" and bag is None) else plan.fe_layers[0][0]
    c = Cache(idx=idx, X=X, gather=True, train=train, R=R, seed=seed, rng_base=rng_base, mode="gated", composite=True)
    c.z = _buf(bufs, "z", (R, 2 * D), dev)
    c.hd, c.pre = [_buf(bufs, "hd0", (R, H), dev)], [None]
    c.a, pre2 = _buf(bufs, "a", (R, Hg), dev), _buf(bufs, "pre2", (R, D), dev)
    c.g, c.t = _buf(bufs, "g", (R, D), dev), _buf(bufs, "t", (R, D), dev)
    has_aug = plan.aug is not None and augment
    o = _buf(bufs, "o", (R, D), dev) if has_aug else None
    q = _buf(bufs, "q", (R, D), dev) if (has_aug and want_q) else None
    c.desc = _tower_desc(plan, X, W1, p_drop, seed, rng_base, state, precision, augment, bag)
    c.desc.x_rounded, c.desc.w1_rounded = int(pre), int(precision != "fp32" and bag is None)
    c.bag = bag            # keeps the CSR tensors alive as long as the cache
    if precision != "fp32" and os.environ.get("TTAM_NO_PREP", "0") == "0":
        # TF32-rounded and rounded-transposed copies of the three small weights (ttam_prepare_weights inside the forward call)
        (_, _), (W2, _) = plan.fe_layers
        G1, _, G2, _ = plan.gate
        for name, W in (("W2r", W2), ("G1r", G1), ("G2r", G2)):
            setattr(c.desc, name, _buf(bufs, f"prep_{name}", (W.shape[0], W.shape[1]), dev).data_ptr())
            setattr(c.desc, name + "T", _buf(bufs, f"prep_{name}T", (W.shape[1], W.shape[0]), dev).data_ptr())
    b = F._lib.TowerBufs()
    b.z, b.hd, b.a, b.pre2, b.g, b.t = (t.data_ptr() for t in (c.z, c.hd[0], c.a, pre2, c.g, c.t))
    b.o, b.q = (None if o is None else o.data_ptr()), (None if q is None else q.data_ptr())
    c.cbufs = b
    F.check(F.lib().ttam_tower_fwd(c.desc, idx.data_ptr(), R, b, F._stream()), "tower_fwd")
    c.o, c.q = (o if has_aug else c.t), q
    return c


def _tower_backward_composite(plan, c: Cache, dt, grads: dict, bufs, phase: int = 0, wgrad_stream=None):
    """phase 0: whole backward.  phase 1: data-gradient chain only (returns dL/de; the weight gradients are registered in
    `grads` but not yet computed).  phase 2: the weight / bias gradients of a cache whose phase 1 has run (any stream that
    is ordered after it)."""
    R, D, dev = c.R, plan.D, dt.device
    L = F.lib()
    if phase == 2:
        if c.wgrads_issued:        # phase 1 already interleaved them on the weight-gradient stream
            return None
        g = c.bwd_g
        g.phase = 2
        ws = F.workspace(L.ttam_tower_bwd_workspace_bytes(c.desc, R), dev, "wgrad")
        F.check(L.ttam_tower_bwd(c.desc, c.idx.data_ptr(), R, c.cbufs, c.bwd_dt.data_ptr(), g, ws.data_ptr(), ws.numel(), F._stream()),
                "tower_bwd")
        return None
    (W1, b1), (W2, b2) = plan.fe_layers
    G1, c1, G2, c2 = plan.gate
    g = F._lib.TowerGrads()
    dz = _buf(bufs, "dz", (R, 2 * D), dev)
    g.dz, g.dpre2 = dz.data_ptr(), _buf(bufs, "dpre2", (R, D), dev).data_ptr()
    g.dpre1, g.dhd = _buf(bufs, "dpre1", (R, G1.shape[0]), dev).data_ptr(), _buf(bufs, "dpre_h0", (R, W1.shape[0]), dev).data_ptr()
    acc = id(W1) in grads
    for W, b, fw, fb in ((W1, b1, "dW1", "db1"), (W2, b2, "dW2", "db2"), (G1, c1, "dG1", "dc1"), (G2, c2, "dG2", "dc2")):
        if not acc:
            grads[id(W)] = _buf(bufs, f"dw{id(W)}", tuple(W.shape), dev)
            grads[id(b)] = _buf(bufs, f"db{id(b)}", (b.shape[0], 1), dev).view(-1)
        setattr(g, fw, grads[id(W)].data_ptr()); setattr(g, fb, grads[id(b)].data_ptr())
    g.accumulate = 1 if acc else 0
    g.phase = phase
    dt = dt if dt.is_contiguous() else dt.contiguous()
    c.bwd_g, c.bwd_dt = g, dt
    if phase == 1 and wgrad_stream is not None:
        # Interleaved: link i of the data-gradient chain on the current stream, weight gradient i on `wgrad_stream` as soon
        # as that link has run (it needs nothing later), instead of all four weight gradients after the whole chain.
        cur = torch.cuda.current_stream(dev)
        with F.ws_scope(F._ws_scope[0] + "_wgrad"):
            ws = F.workspace(L.ttam_tower_bwd_workspace_bytes(c.desc, R), dev, "wgrad")
        for i in range(4):
            g.phase = 10 + i
            F.check(L.ttam_tower_bwd(c.desc, c.idx.data_ptr(), R, c.cbufs, dt.data_ptr(), g, ws.data_ptr(), ws.numel(), cur.cuda_stream),
                    "tower_bwd")
            wgrad_stream.wait_stream(cur)
            g.phase = 20 + i
            F.check(L.ttam_tower_bwd(c.desc, c.idx.data_ptr(), R, c.cbufs, dt.data_ptr(), g, ws.data_ptr(), ws.numel(),
                                     wgrad_stream.cuda_stream), "tower_bwd")
        g.phase = 1
        c.wgrads_issued = True
        return dz[:, :D]
    ws = F.workspace(L.ttam_tower_bwd_workspace_bytes(c.desc, R), dev, "wgrad")
    F.check(L.ttam_tower_bwd(c.desc, c.idx.data_ptr(), R, c.cbufs, dt.data_ptr(), g, ws.data_ptr(), ws.numel(), F._stream()),
            "tower_bwd")
    return dz[:, :D]


def tower_forward(plan: TowerPlan, idx: torch.Tensor, X: Optional[torch.Tensor], *, gather: bool, train: bool,
                  bufs: Optional[dict] = None, seed: int = 0, rng_base: int = 0, state=None, precision="fp32",
                  want_q: bool = False, augment: bool = True) -> Cache:
    """idx [R] int64.  X: feature matrix; if `gather` its rows are X[idx] (fused index_select, reference
    training.py:743-775), else X is already [R, F].  Returns a cache with t (base output), o (= t + aug[idx]
    when the plan has an augmentation table), q and the intermediates the backward needs."""
    if _composite_ok(plan, X, gather, bufs) and idx.numel() > 0:
        return _tower_forward_composite(plan, idx, X, train=train, bufs=bufs, seed=seed, rng_base=rng_base, state=state,
                                        precision=precision, want_q=want_q, augment=augment)
    R, D = idx.numel(), plan.D
    dev = plan.table.device
    c = Cache(idx=idx, X=X, gather=gather, train=train, R=R, seed=seed, rng_base=rng_base)
    gidx = idx if gather else None
    use_feat = plan.fusion != "identity" and plan.fe_kind != "none" and X is not None
    p_drop = plan.dropout if train else 0.0
    zcat = use_feat and plan.fusion in ("gated", "concat")
    Df = plan.fe_layers[-1][0].shape[0] if plan.fe_layers else (X.shape[1] if use_feat else 0)
    if zcat:
        z = _buf(bufs, "z", (R, D + Df), dev)
        e_view, f_view = z[:, :D], z[:, D:]
        c.z = z
    else:
        e_view = _buf(bufs, "e", (R, D), dev)
        f_view = _buf(bufs, "f", (R, Df), dev) if use_feat else None
    if not (use_feat and plan.fusion == "sum"):
        F.gather_rows(plan.table, idx, out=e_view)                               # e = E[idx]
    if not use_feat:
        c.mode = "identity"
        t = e_view
    else:
        c.mode = plan.fusion
        # ---- feature encoder
        if plan.fe_kind == "identity":
            if gather:
                F.gather_rows(X, idx, out=f_view)
            else:
                f_view.copy_(X)
        elif plan.fe_kind == "linear":
            W, b = plan.fe_layers[0]
            c.bag = _bag_of(X, W.shape[0]) if gather else None
            if c.bag is not None:
                tmp = _buf(bufs, "f_bag", (R, W.shape[0]), dev)      # 16-byte aligned rows for the bag kernel's stores
                F.bag_linear_fwd(c.bag, idx, W, b, out=tmp)
                f_view.copy_(tmp)
            else:
                F.linear_fwd(X, _w_for(W, bufs, precision), b, gather=gidx, out=f_view, precision=precision)
        else:
            h_in, g_in = X, gidx
            c.pre, c.hd = [], []
            off = rng_base
            c.bag = _bag_of(X, plan.fe_layers[0][0].shape[0]) if gather else None
            for li, (W, b) in enumerate(plan.fe_layers[:-1]):
                H = W.shape[0]
                hd = _buf(bufs, f"hd{li}", (R, H), dev)
                if li == 0 and c.bag is not None:            # bag form of layer 1 (fp32 FMA over the row's non-zeros)
                    if plan.activation == "relu":
                        F.bag_linear_fwd(c.bag, idx, W, b, act="relu", out=hd, dropout_p=p_drop, seed=seed, offset=off, state=state)
                        c.pre.append(None)
                    else:
                        pre = _buf(bufs, f"pre{li}", (R, H), dev)
                        F.bag_linear_fwd(c.bag, idx, W, b, out=pre)
                        F.act_fwd(pre, act=plan.activation, out=hd, dropout_p=p_drop, seed=seed, offset=off, state=state)
                        c.pre.append(pre)
                    c.hd.append(hd)
                    off += R * H
                    h_in, g_in = hd, None
                    continue
                W = _w_for(W, bufs, precision)
                if plan.activation == "relu":
                    F.linear_fwd(h_in, W, b, gather=g_in, act="relu", out=hd, dropout_p=p_drop, seed=seed, offset=off,
                                 state=state, precision=precision)
                    c.pre.append(None)
                else:
                    pre = _buf(bufs, f"pre{li}", (R, H), dev)
                    F.linear_fwd(h_in, W, b, gather=g_in, out=pre, precision=precision)
                    F.act_fwd(pre, act=plan.activation, out=hd, dropout_p=p_drop, seed=seed, offset=off, state=state)
                    c.pre.append(pre)
                c.hd.append(hd)
                off += R * H
                h_in, g_in = hd, None
            W, b = plan.fe_layers[-1]
            F.linear_fwd(h_in, _w_for(W, bufs, precision), b, gather=g_in, out=f_view, precision=precision)
        # ---- fusion
        if plan.fusion == "sum":
            if Df != D:
                raise ValueError("Feature encoder output dimension must match id embedding dimension for 'sum' fusion.")
            t = _buf(bufs, "t", (R, D), dev)
            F.augment_fwd(f_view, plan.table, idx, out=t)                       # t = f + E[idx]
        elif plan.fusion == "concat":
            Wp, bp = plan.proj
            t = _buf(bufs, "t", (R, Wp.shape[0]), dev)
            F.linear_fwd(z, Wp, bp, out=t, precision=precision)
        else:  # gated
            if Df != D:
                raise ValueError("Adaptive mimic requires feature encoder output to match id embedding dimension.")
            G1, c1, G2, c2 = plan.gate
            a = _buf(bufs, "a", (R, G1.shape[0]), dev)
            pre2 = _buf(bufs, "pre2", (R, D), dev)
            F.linear_fwd(z, G1, c1, act="relu", out=a, precision=precision)
            F.linear_fwd(a, G2, c2, out=pre2, precision=precision)
            g = _buf(bufs, "g", (R, D), dev)
            t = _buf(bufs, "t", (R, D), dev)
            c.a, c.g = a, g
            if plan.aug is not None and augment:
                o = _buf(bufs, "o", (R, D), dev)
                q = _buf(bufs, "q", (R, D), dev) if want_q else None
                F.gate_fwd(z, pre2, aug_table=plan.aug, idx=idx, g=g, t=t, o=o, q=q)
                c.t, c.o, c.q = t, o, q
                return c
            F.gate_fwd(z, pre2, g=g, t=t)
    c.t = t
    if plan.aug is not None and augment:
        o = _buf(bufs, "o", (R, plan.out_dim), dev)
        q = _buf(bufs, "q", (R, plan.out_dim), dev) if want_q else None
        F.augment_fwd(t, plan.aug, idx, out=o, q_out=q)
        c.o, c.q = o, q
    else:
        c.o = t
    return c


def splits_backward(c: Cache) -> bool:
    """True when tower_backward(..., phase=1) / (phase=2) is available for this forward cache."""
    return bool(c.composite)


def tower_backward(plan: TowerPlan, c: Cache, dt: torch.Tensor, grads: dict, *, bufs: Optional[dict] = None,
                   state=None, precision="fp32", phase: int = 0, wgrad_stream=None):
    """dt [R, out_dim] = dL/dt.  Dense weight gradients are written to grads[id(param)] (accumulated when the
    key exists).  Returns de [R, D] (a view; rows of dL/dE[idx], duplicates NOT yet summed).
    phase (composite towers only, see `splits_backward`): 1 = data-gradient chain now, weight gradients later by a
    phase-2 call on any stream ordered after this one."""
    if c.composite:
        return _tower_backward_composite(plan, c, dt, grads, bufs, phase, wgrad_stream)
    if phase != 0:
        raise ValueError("only composite (gated MLP, tensor-core) towers split their backward")
    R, D = c.R, plan.D
    dev = dt.device
    X, gidx = c.X, (c.idx if c.gather else None)

    def wgrad(dy, x, W, b, gather=None):
        dw = grads.get(id(W))
        acc = dw is not None
        if not acc:
            dw = _buf(bufs, f"dw{id(W)}", tuple(W.shape), dev) if bufs is not None else torch.empty_like(W)
            db = (_buf(bufs, f"db{id(b)}", (b.shape[0], 1), dev).view(-1) if bufs is not None else torch.empty_like(b))
            grads[id(W)], grads[id(b)] = dw, db
        else:
            db = grads[id(b)]
        if gather is not None and c.bag is not None and c.bag.wgrad and x is X:
            dyc = dy
            if dy.stride(0) % 4 != 0 or dy.data_ptr() % 16 != 0 or (dy.shape[1] > 1 and dy.stride(1) != 1):
                dyc = dy.contiguous()
            F.bag_linear_wgrad(c.bag, gather, dyc, dw=dw, db=db, accumulate=acc)
            return
        F.linear_wgrad(dy, x, gather=gather, dw=dw, db=db, accumulate=acc, precision=precision)

    if c.mode == "identity":
        return dt
    if c.mode == "sum":
        de = df = dt
    elif c.mode == "concat":
        Wp, bp = plan.proj
        wgrad(dt, c.z, Wp, bp)
        dz = _buf(bufs, "dz", tuple(c.z.shape), dev)
        F.linear_dgrad(dt, Wp, out=dz, precision=precision)
        de, df = dz[:, :D], dz[:, D:]
    else:  # gated
        G1, c1, G2, c2 = plan.gate
        dpre2 = _buf(bufs, "dpre2", (R, D), dev)
        dz = _buf(bufs, "dz", (R, 2 * D), dev)
        F.gate_bwd(dt, c.z, c.g, dpre2=dpre2, dz=dz)
        wgrad(dpre2, c.a, G2, c2)
        dpre1 = _buf(bufs, "dpre1", tuple(c.a.shape), dev)
        F.linear_dgrad(dpre2, G2, out=dpre1, aux=c.a, relu_mask=True, precision=precision)
        wgrad(dpre1, c.z, G1, c1)
        F.linear_dgrad(dpre1, G1, out=dz, accumulate=True, precision=precision)
        de, df = dz[:, :D], dz[:, D:]
    # ---- feature encoder (features are constants: no dx)
    if plan.fe_kind == "linear":
        W, b = plan.fe_layers[0]
        wgrad(df, X, W, b, gather=gidx)
    elif plan.fe_kind == "mlp":
        p_drop = plan.dropout if c.train else 0.0
        dy = df
        n_hidden = len(plan.fe_layers) - 1
        offs = [c.rng_base]
        for li in range(n_hidden):
            offs.append(offs[-1] + R * plan.fe_layers[li][0].shape[0])
        for li in range(n_hidden, 0, -1):
            W, b = plan.fe_layers[li]
            hd = c.hd[li - 1]
            wgrad(dy, hd, W, b)
            dpre = _buf(bufs, f"dpre_h{li - 1}", tuple(hd.shape), dev)
            if plan.activation == "relu":
                scale = 1.0 / (1.0 - p_drop) if p_drop > 0 else 1.0
                F.linear_dgrad(dy, W, out=dpre, aux=hd, relu_mask=True, scale=scale, precision=precision)
            else:
                F.linear_dgrad(dy, W, out=dpre, precision=precision)
                F.act_bwd(dpre, c.pre[li - 1], act=plan.activation, out=dpre, dropout_p=p_drop, seed=c.seed,
                          offset=offs[li - 1], state=state)
            dy = dpre
        W, b = plan.fe_layers[0]
        wgrad(dy, X, W, b, gather=gidx)
    return de


# ------------------------------------------------------------------------------------------------
# nn.Module API: autograd.Function wrappers (generic path; the fused engine does not use autograd)
# ------------------------------------------------------------------------------------------------
def _need_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise F._lib.TtamError(f"{what} is on {t.device}: this package runs on CUDA (sm_100a) only, there is no CPU path")


def _table_grad(table: torch.Tensor, idx: torch.Tensor, rows: torch.Tensor, sparse: bool):
    rows = rows.contiguous()
    if sparse:   # same layout nn.Embedding(sparse=True) produces: uncoalesced COO (reference training.py:822)
        return torch.sparse_coo_tensor(idx.view(1, -1), rows, table.shape)
    return torch.zeros_like(table).index_add_(0, idx, rows)


class _TowerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tower, idx, feats, seed, *params):
        plan = plan_from_module(tower)
        c = tower_forward(plan, idx, feats, gather=False, train=tower.training, seed=seed, augment=False)
        ctx.plan, ctx.cache, ctx.params = plan, c, params
        return c.t

    @staticmethod
    def backward(ctx, dt):
        plan, c = ctx.plan, ctx.cache
        grads: dict = {}
        de = tower_backward(plan, c, dt.contiguous(), grads)
        out = []
        for p in ctx.params:
            if p is plan.table:
                out.append(_table_grad(plan.table, c.idx, de, plan.sparse))
            else:
                out.append(grads.get(id(p)))
        return (None, None, None, None, *out)


_seed_counter = [0x5EED]


def tower_module_forward(tower, indices: torch.Tensor, features: Optional[torch.Tensor]) -> torch.Tensor:
    _need_cuda(tower.embedding.weight, "tower parameters")
    idx = indices.reshape(-1).contiguous()
    feats = None if features is None else features.reshape(idx.numel(), -1).contiguous().float()
    params = [tower.embedding.weight] + [p for n, p in tower.named_parameters() if n != "embedding.weight"]
    _seed_counter[0] += 1
    out = _TowerFn.apply(tower, idx, feats, _seed_counter[0], *params)
    return out.reshape(*indices.shape, out.shape[-1])


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, b, act):
        y = F.linear_fwd(x, W, b, act=act)
        ctx.save_for_backward(x, W, y)
        ctx.act = act
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W, y = ctx.saved_tensors
        dy = dy.contiguous()
        if ctx.act == "relu":
            dy = dy * (y > 0)
        dw, db = F.linear_wgrad(dy, x)
        dx = F.linear_dgrad(dy, W) if ctx.needs_input_grad[0] else None
        return dx, dw, db, None


def feature_encoder_forward(fe, inputs: torch.Tensor) -> torch.Tensor:
    """Stand-alone FeatureEncoderWrapper.forward (only used when a caller invokes the sub-module directly)."""
    _need_cuda(inputs, "features")
    x = inputs.reshape(-1, inputs.shape[-1]).contiguous().float()
    if fe.kind == "identity" or isinstance(fe.network, nn.Identity):
        return inputs
    lins = fe.linear_layers()
    if fe.kind == "mlp" and fe.activation != "relu":
        raise NotImplementedError("direct call of a non-ReLU MLP feature encoder: use the TowerEncoder forward")
    for li, m in enumerate(lins):
        last = li == len(lins) - 1
        x = _LinearFn.apply(x, m.weight, m.bias, "none" if last else "relu")
        if not last and fe.dropout and fe.training:
            x = torch.nn.functional.dropout(x, fe.dropout, True)
    return x.reshape(*inputs.shape[:-1], x.shape[-1])


class _GateFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, e, f, G1, c1, G2, c2):
        R, D = e.shape
        z = torch.cat([e, f], dim=1)
        a = F.linear_fwd(z, G1, c1, act="relu")
        pre2 = F.linear_fwd(a, G2, c2)
        g = torch.empty_like(e)
        t = torch.empty_like(e)
        F.gate_fwd(z, pre2, g=g, t=t)
        ctx.save_for_backward(z, a, g, G1, G2)
        return t

    @staticmethod
    def backward(ctx, dt):
        z, a, g, G1, G2 = ctx.saved_tensors
        R, D = g.shape
        dt = dt.contiguous()
        dpre2 = torch.empty_like(g)
        dz = torch.empty_like(z)
        F.gate_bwd(dt, z, g, dpre2=dpre2, dz=dz)
        dG2, dc2 = F.linear_wgrad(dpre2, a)
        dpre1 = F.linear_dgrad(dpre2, G2, aux=a, relu_mask=True)
        dG1, dc1 = F.linear_wgrad(dpre1, z)
        F.linear_dgrad(dpre1, G1, out=dz, accumulate=True)
        return dz[:, :D].contiguous(), dz[:, D:].contiguous(), dG1, dc1, dG2, dc2


def gate_forward(gate, id_repr: torch.Tensor, feature_repr: torch.Tensor) -> torch.Tensor:
    _need_cuda(id_repr, "id_repr")
    D = id_repr.shape[-1]
    n = gate.gate_network
    t = _GateFn.apply(id_repr.reshape(-1, D).contiguous(), feature_repr.reshape(-1, D).contiguous(),
                      n[0].weight, n[0].bias, n[2].weight, n[2].bias)
    return t.reshape(id_repr.shape)


class _AugmentFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, base, table, idx, sparse):
        o = torch.empty_like(base)
        q = torch.empty_like(base)
        F.augment_fwd(base, table, idx, out=o, q_out=q)
        ctx.save_for_backward(idx, table)
        ctx.sparse = sparse
        return o, q

    @staticmethod
    def backward(ctx, do, dq):
        idx, table = ctx.saved_tensors
        rows = do if dq is None else (dq if do is None else do + dq)
        return do, _table_grad(table, idx, rows, ctx.sparse), None, None


def augment(table: nn.Embedding, indices: torch.Tensor, base: torch.Tensor):
    """AdaptiveMimicMechanism._apply_aug (reference adaptive_mimic.py:88-105): returns (base + A[idx], A[idx])."""
    if indices.dtype != torch.long:
        raise ValueError("Adaptive mimic indices must be torch.long tensors.")
    _need_cuda(base, "base embedding")
    D = base.shape[-1]
    o, q = _AugmentFn.apply(base.reshape(-1, D).contiguous(), table.weight, indices.reshape(-1).contiguous(),
                            bool(table.sparse))
    return o.reshape(base.shape), q.reshape(base.shape)


def mimic_forward(mech, user_indices, item_indices, user_embedding, item_embedding):
    """AdaptiveMimicMechanism.forward (reference adaptive_mimic.py:40-68)."""
    aug_u, q_u = augment(mech.user_augmented, user_indices, user_embedding)
    aug_i, q_i = augment(mech.item_augmented, item_indices, item_embedding)
    loss_u = torch.nn.functional.mse_loss(q_u, item_embedding.detach())
    loss_i = torch.nn.functional.mse_loss(q_i, user_embedding.detach())
    return aug_u, aug_i, loss_u, loss_i
