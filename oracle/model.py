"""numpy fp32 restatement of the reference towers, mimic mechanism, loss and one training step.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function cites the reference lines it follows;
backward passes are written out by hand (the reference relies on autograd, training.py:822) so that
the derivation in SURVEY.md Appendix A is itself pinned by the golden fixtures.
State is a dict {reference state_dict key -> np.ndarray(float32)}.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field

import numpy as np

from .optim import OptState, dense_step, sparse_adam_step

F32 = np.float32


# --------------------------------------------------------------------------------------------
# model description
# --------------------------------------------------------------------------------------------
@dataclass
class TowerSpec:
    fe_type: str = "none"          # none | identity | linear | mlp        encoders.py:102-146
    fe_layers: list = field(default_factory=list)   # state-dict prefixes of the Linear layers, in order
    activation: str = "relu"       # relu | gelu | tanh | selu              encoders.py:68-78
    dropout: float = 0.0           # encoders.py:136-137
    fusion: str = "identity"       # identity | sum | concat | gated        encoders.py:203-255
    sparse: bool = True            # nn.Embedding(sparse=...)               encoders.py:54-60


@dataclass
class ModelSpec:
    user: TowerSpec
    item: TowerSpec
    mimic: bool


def _tower_spec(state, side, activation, dropout, fusion, sparse) -> TowerSpec:
    pre = f"{side}_encoder."
    spec = TowerSpec(activation=activation, dropout=dropout, sparse=sparse)
    if pre + "feature_encoder.network.weight" in state:
        spec.fe_type = "linear"
        spec.fe_layers = [pre + "feature_encoder.network"]
    else:
        idxs = sorted({int(m.group(1)) for k in state
                       for m in [re.match(re.escape(pre) + r"feature_encoder\.network\.(\d+)\.weight$", k)] if m})
        if idxs:
            spec.fe_type = "mlp"
            spec.fe_layers = [pre + f"feature_encoder.network.{i}" for i in idxs]
    if fusion is not None:
        spec.fusion = fusion
    elif pre + "adaptive_mimic.gate_network.0.weight" in state:
        spec.fusion = "gated"
    elif pre + "projection.weight" in state:
        spec.fusion = "concat"
    elif spec.fe_type != "none":
        spec.fusion = "sum"
    if spec.fe_type == "none" and fusion is None:
        spec.fusion = "identity"
    return spec


def spec_from_state(state, *, activation="relu", dropout=0.0, fusion_user=None, fusion_item=None,
                    sparse=True) -> ModelSpec:
    """Infer the module structure from reference state_dict keys (SURVEY 8(b) key list)."""
    return ModelSpec(
        user=_tower_spec(state, "user", activation, dropout, fusion_user, sparse),
        item=_tower_spec(state, "item", activation, dropout, fusion_item, sparse),
        mimic="adaptive_mimic.user_augmented.weight" in state,
    )


# --------------------------------------------------------------------------------------------
# activations (encoders.py:68-78 -> torch.nn.ReLU/GELU/Tanh/SELU)
# --------------------------------------------------------------------------------------------
_SELU_ALPHA = 1.6732632423543772848170429916717
_SELU_SCALE = 1.0507009873554804934193349852946


def _erf(x):
    try:
        from scipy.special import erf  # scipy is in the image
        return erf(x.astype(np.float64)).astype(F32)
    except Exception:  # pragma: no cover
        import math
        return np.vectorize(math.erf)(x.astype(np.float64)).astype(F32)


def _act(name, x):
    if name == "relu":
        return np.maximum(x, F32(0))
    if name == "tanh":
        return np.tanh(x)
    if name == "gelu":
        return (F32(0.5) * x * (F32(1) + _erf(x / F32(np.sqrt(2.0))))).astype(F32)
    if name == "selu":
        return (F32(_SELU_SCALE) * np.where(x > 0, x, F32(_SELU_ALPHA) * (np.exp(np.minimum(x, 0)) - F32(1)))).astype(F32)
    raise ValueError(f"Unsupported activation '{name}'")


def _dact(name, pre, out):
    """d act / d pre, given both the pre-activation and the output."""
    if name == "relu":
        return (out > 0).astype(F32)
    if name == "tanh":
        return F32(1) - out * out
    if name == "gelu":
        pdf = np.exp(F32(-0.5) * pre * pre) * F32(1.0 / np.sqrt(2.0 * np.pi))
        cdf = F32(0.5) * (F32(1) + _erf(pre / F32(np.sqrt(2.0))))
        return (cdf + pre * pdf).astype(F32)
    if name == "selu":
        return np.where(pre > 0, F32(_SELU_SCALE), F32(_SELU_SCALE * _SELU_ALPHA) * np.exp(np.minimum(pre, 0))).astype(F32)
    raise ValueError(name)


def _sigmoid(x):
    return (F32(1) / (F32(1) + np.exp(-x))).astype(F32)


def _linear(x, w, b):
    return (x @ w.T + b).astype(F32)


# --------------------------------------------------------------------------------------------
# tower forward / backward
# --------------------------------------------------------------------------------------------
def tower_forward(state, side, spec: TowerSpec, idx, x=None, *, train=False, masks=None):
    """TowerEncoder.forward (encoders.py:221-255).  Returns a cache dict; cache['t'] is the output.

    `masks`: optional list of {0,1} float arrays, one per hidden layer, standing in for nn.Dropout's
    Bernoulli draw (train mode only; scale 1/(1-p) applied here, encoders.py:136-137)."""
    pre = f"{side}_encoder."
    E = state[pre + "embedding.weight"]
    e = E[idx]                                             # encoders.py:223
    c = {"idx": idx, "e": e, "x": x, "spec": spec}
    if spec.fusion == "identity" or spec.fe_type == "none" or x is None:   # encoders.py:225-231
        c["t"] = e
        c["mode"] = "identity"
        return c
    # feature encoder (encoders.py:102-146)
    if spec.fe_type == "identity":
        f = x
    elif spec.fe_type == "linear":
        f = _linear(x, state[spec.fe_layers[0] + ".weight"], state[spec.fe_layers[0] + ".bias"])
    else:
        h = x
        c["pre_h"], c["h"], c["hd"] = [], [], []
        for li, name in enumerate(spec.fe_layers[:-1]):
            p_ = _linear(h, state[name + ".weight"], state[name + ".bias"])
            a_ = _act(spec.activation, p_)
            hd = a_
            if train and spec.dropout > 0:
                m = masks[li]
                hd = (a_ * m / F32(1.0 - spec.dropout)).astype(F32)
            c["pre_h"].append(p_)
            c["h"].append(a_)
            c["hd"].append(hd)
            h = hd
        last = spec.fe_layers[-1]
        f = _linear(h, state[last + ".weight"], state[last + ".bias"])
    c["f"] = f
    c["mode"] = spec.fusion
    if spec.fusion == "sum":                                # encoders.py:235-240
        if f.shape[-1] != e.shape[-1]:
            raise ValueError("Feature encoder output dimension must match id embedding dimension for 'sum' fusion.")
        c["t"] = (e + f).astype(F32)
    elif spec.fusion == "concat":                           # encoders.py:242-244
        z = np.concatenate([e, f], axis=-1)
        c["z"] = z
        c["t"] = _linear(z, state[pre + "projection.weight"], state[pre + "projection.bias"])
    elif spec.fusion == "gated":                            # encoders.py:149-168
        g1w, g1b = state[pre + "adaptive_mimic.gate_network.0.weight"], state[pre + "adaptive_mimic.gate_network.0.bias"]
        g2w, g2b = state[pre + "adaptive_mimic.gate_network.2.weight"], state[pre + "adaptive_mimic.gate_network.2.bias"]
        z = np.concatenate([e, f], axis=-1)
        a = np.maximum(_linear(z, g1w, g1b), F32(0))
        g = _sigmoid(_linear(a, g2w, g2b))
        c.update(z=z, a=a, g=g)
        c["t"] = (g * e + (F32(1) - g) * f).astype(F32)
    else:
        raise ValueError(f"Unsupported fusion strategy: {spec.fusion}")
    return c


def tower_backward(state, side, spec: TowerSpec, c, dt, grads, sparse_rows, masks=None, train=False):
    """Backward of tower_forward (SURVEY Appendix A 'Backward').  Accumulates dense weight grads
    into `grads[name]` and appends (idx, de) to sparse_rows[embedding key]."""
    pre = f"{side}_encoder."

    def acc(name, g):
        grads[name] = g.astype(F32) if name not in grads else (grads[name] + g).astype(F32)

    def lin_bwd(name, x_in, dy):
        acc(name + ".weight", dy.T @ x_in)
        acc(name + ".bias", dy.sum(axis=0))
        return (dy @ state[name + ".weight"]).astype(F32)

    e = c["e"]
    if c["mode"] == "identity":
        de = dt
        sparse_rows.setdefault(pre + "embedding.weight", []).append((c["idx"], de.astype(F32)))
        return
    f = c["f"]
    if c["mode"] == "sum":
        de, df = dt, dt
    elif c["mode"] == "concat":
        dz = lin_bwd(pre + "projection", c["z"], dt)
        D = e.shape[-1]
        de, df = dz[:, :D], dz[:, D:]
    else:  # gated
        g, a, z = c["g"], c["a"], c["z"]
        D = e.shape[-1]
        dg = dt * (e - f)
        dpre2 = (dg * g * (F32(1) - g)).astype(F32)
        da = lin_bwd(pre + "adaptive_mimic.gate_network.2", a, dpre2)
        dpre1 = (da * (a > 0)).astype(F32)
        dz = lin_bwd(pre + "adaptive_mimic.gate_network.0", z, dpre1)
        de = (dt * g + dz[:, :D]).astype(F32)
        df = (dt * (F32(1) - g) + dz[:, D:]).astype(F32)
    sparse_rows.setdefault(pre + "embedding.weight", []).append((c["idx"], de.astype(F32)))
    # feature encoder backward (features are constants: no dx)
    if spec.fe_type == "identity":
        return
    if spec.fe_type == "linear":
        lin_bwd(spec.fe_layers[0], c["x"], df)
        return
    dh = lin_bwd(spec.fe_layers[-1], c["hd"][-1] if c["hd"] else c["x"], df)
    for li in range(len(spec.fe_layers) - 2, -1, -1):
        if train and spec.dropout > 0:
            dh = (dh * masks[li] / F32(1.0 - spec.dropout)).astype(F32)
        dpre = (dh * _dact(spec.activation, c["pre_h"][li], c["h"][li])).astype(F32)
        x_in = c["hd"][li - 1] if li > 0 else c["x"]
        dh = lin_bwd(spec.fe_layers[li], x_in, dpre)


# --------------------------------------------------------------------------------------------
# loss (training.py:770-803, adaptive_mimic.py:59-68)
# --------------------------------------------------------------------------------------------
def _softplus(x):
    # BCEWithLogits: max(x,0) - x*y + log1p(exp(-|x|))
    return (np.maximum(x, F32(0)) + np.log1p(np.exp(-np.abs(x)))).astype(F32)


def loss_forward_backward(o_u, o_p, o_n, *, t_u=None, t_p=None, q_u=None, q_p=None,
                          lambda_u=0.0, lambda_i=0.0, need_grad=True):
    """Fused loss: sampled-negative BCE-with-logits + the two mimic MSEs.

    o_u [B,D], o_p [B,D], o_n [B,N,D] final (augmented) embeddings; t_*: base tower outputs (mimic
    targets, detached); q_*: augmentation rows of the positive pairs.
    Returns dict(loss, bce, mimic_user, mimic_item, do_u, do_p, do_n, dq_u_extra, dq_p_extra)."""
    B, N, D = o_n.shape
    M = B * (1 + N)
    s_pos = (o_u * o_p).sum(-1).astype(F32)                       # training.py:770
    s_neg = (o_u[:, None, :] * o_n).sum(-1).astype(F32)           # training.py:786-787
    bce = (_softplus(-s_pos).sum(dtype=F32) + _softplus(s_neg).sum(dtype=F32)) / F32(M)   # :789-798 mean
    out = {"bce": F32(bce), "s_pos": s_pos, "s_neg": s_neg}
    total = F32(bce)
    mu = mi = None
    if q_u is not None:
        mu = F32(((q_u - t_p) ** 2).mean(dtype=F32))              # adaptive_mimic.py:66
        mi = F32(((q_p - t_u) ** 2).mean(dtype=F32))              # adaptive_mimic.py:67
        if lambda_u > 0:
            total = F32(total + F32(lambda_u) * mu)               # training.py:800-801
        if lambda_i > 0:
            total = F32(total + F32(lambda_i) * mi)               # training.py:802-803
    out.update(loss=total, mimic_user=mu, mimic_item=mi)
    if not need_grad:
        return out
    ds_pos = ((_sigmoid(s_pos) - F32(1)) / F32(M)).astype(F32)
    ds_neg = (_sigmoid(s_neg) / F32(M)).astype(F32)
    out["do_u"] = (ds_pos[:, None] * o_p + (ds_neg[:, :, None] * o_n).sum(1)).astype(F32)
    out["do_p"] = (ds_pos[:, None] * o_u).astype(F32)
    out["do_n"] = (ds_neg[:, :, None] * o_u[:, None, :]).astype(F32)
    if q_u is not None:
        out["dq_u_extra"] = (F32(lambda_u if lambda_u > 0 else 0.0) * F32(2.0) * (q_u - t_p) / F32(B * D)).astype(F32)
        out["dq_p_extra"] = (F32(lambda_i if lambda_i > 0 else 0.0) * F32(2.0) * (q_p - t_u) / F32(B * D)).astype(F32)
    return out


def inbatch_loss_forward_backward(o_u, o_p, *, t_u=None, t_p=None, q_u=None, q_p=None, lambda_u=0.0, lambda_i=0.0,
                                  need_grad=True):
    """EXTENSION (BASELINE.json configs[1] names "in-batch negatives"; the reference has no such loss - its only training loss
    is the sampled-negative BCE of training.py:770-803 - so this definition is pinned by tests/test_oracle_inbatch.py
    against torch autograd, not against the reference: parity UNPINNED with respect to the reference).

    In-batch softmax: scores S = o_u o_p^T [B, B]; row b's positive is column b, the other B-1 items of the batch are its
    negatives; L_ce = mean_b (logsumexp_j S[b, j] - S[b, b]) (= F.cross_entropy(S, arange(B))).  The mimic terms are the
    reference's (adaptive_mimic.py:66-67) and enter the total exactly as in training.py:800-803.
    Returns dict(loss, ce, mimic_user, mimic_item, do_u, do_p, dq_u_extra, dq_p_extra)."""
    B, D = o_u.shape
    S = (o_u.astype(F32) @ o_p.astype(F32).T).astype(F32)
    m = S.max(axis=1, keepdims=True)
    e = np.exp(S - m, dtype=F32)
    z = e.sum(axis=1, keepdims=True, dtype=F32)
    lse = (np.log(z) + m)[:, 0].astype(F32)
    ce = F32((lse - np.diagonal(S)).sum(dtype=F32) / F32(B))
    out = {"ce": ce, "scores": S}
    total = ce
    mu = mi = None
    if q_u is not None:
        mu = F32(((q_u - t_p) ** 2).mean(dtype=F32))
        mi = F32(((q_p - t_u) ** 2).mean(dtype=F32))
        if lambda_u > 0:
            total = F32(total + F32(lambda_u) * mu)
        if lambda_i > 0:
            total = F32(total + F32(lambda_i) * mi)
    out.update(loss=total, mimic_user=mu, mimic_item=mi)
    if not need_grad:
        return out
    P = (e / z).astype(F32)
    P[np.arange(B), np.arange(B)] -= F32(1)
    P /= F32(B)                                                   # dL/dS
    out["do_u"] = (P @ o_p).astype(F32)
    out["do_p"] = (P.T @ o_u).astype(F32)
    if q_u is not None:
        out["dq_u_extra"] = (F32(lambda_u if lambda_u > 0 else 0.0) * F32(2.0) * (q_u - t_p) / F32(B * D)).astype(F32)
        out["dq_p_extra"] = (F32(lambda_i if lambda_i > 0 else 0.0) * F32(2.0) * (q_p - t_u) / F32(B * D)).astype(F32)
    return out


def _cov(m):
    """_compute_covariance (training.py:530-538)."""
    if m.shape[0] <= 1:
        return np.zeros((m.shape[1], m.shape[1]), dtype=F32)
    cen = m - m.mean(axis=0, keepdims=True, dtype=F32)
    return (cen.T @ cen / F32(m.shape[0] - 1)).astype(F32)


def category_alignment_loss(item_idx, emb, cat_tensor, major, need_grad=True):
    """_category_alignment_loss (training.py:541-579) + its gradient wrt `emb`."""
    zero = (F32(0), np.zeros_like(emb) if need_grad else None)
    if cat_tensor is None or major is None or item_idx.size == 0:
        return zero
    cats = cat_tensor[item_idx]
    uniq = np.unique(cats)
    if uniq.size <= 1:
        return zero
    mm = cats == major
    if mm.sum() < 2:
        return zero
    major_cov = _cov(emb[mm])
    loss = F32(0)
    diffs = []
    for cid in uniq.tolist():
        if cid == major:
            continue
        m = cats == cid
        if m.sum() < 2:
            continue
        d = (_cov(emb[m]) - major_cov).astype(F32)
        loss = F32(loss + (d * d).sum(dtype=F32))
        diffs.append((m, d))
    if not diffs:
        return zero
    n_c = len(diffs)
    loss = F32(loss / F32(n_c))
    if not need_grad:
        return loss, None
    grad = np.zeros_like(emb)
    g_major = np.zeros_like(major_cov)
    for m, d in diffs:
        G = (F32(2.0 / n_c) * d).astype(F32)           # dL/dCov_c (symmetric)
        xs = emb[m]
        cen = xs - xs.mean(axis=0, keepdims=True, dtype=F32)
        grad[m] += (F32(2.0 / (xs.shape[0] - 1)) * (cen @ G)).astype(F32)
        g_major -= G
    xs = emb[mm]
    cen = xs - xs.mean(axis=0, keepdims=True, dtype=F32)
    grad[mm] += (F32(2.0 / (xs.shape[0] - 1)) * (cen @ g_major)).astype(F32)
    return loss, grad.astype(F32)


# --------------------------------------------------------------------------------------------
# one training step (training.py:726-831)
# --------------------------------------------------------------------------------------------
def _dense_param_names(state, spec: ModelSpec):
    """_collect_parameter_groups (training.py:276-309): which tensors go to the dense optimiser."""
    sparse = set()
    if spec.user.sparse:
        sparse.add("user_encoder.embedding.weight")
    if spec.item.sparse:
        sparse.add("item_encoder.embedding.weight")
    return [k for k in state if k not in sparse], sorted(sparse)


def train_step(state, opt: OptState, spec: ModelSpec, users, pos, neg, user_x, item_x, *,
               lr=1e-3, weight_decay=0.01, betas=(0.9, 0.999), optimizer="adamw", momentum=0.0,
               lambdas=(0.0, 0.0, 0.0), cat_tensor=None, major=None, masks=None, loss="sampled"):
    """One iteration of the `_train_one_epoch` loop body.  Mutates `state` and `opt` in place;
    returns dict(loss, bce, mimic_user, mimic_item, cal, touched_user_rows, touched_item_rows).
    loss="inbatch" (EXTENSION, not in the reference): the in-batch softmax of inbatch_loss_forward_backward replaces the
    sampled-negative BCE; `neg` is ignored."""
    lam_u, lam_i, lam_c = (float(v) for v in lambdas)
    if loss == "inbatch":
        neg = np.zeros((len(users), 0), dtype=np.int64)
    B, N = neg.shape
    neg_flat = neg.reshape(-1)
    train = True
    mk = masks or {}
    cu = tower_forward(state, "user", spec.user, users, None if user_x is None else user_x[users], train=train, masks=mk.get("user"))
    cp = tower_forward(state, "item", spec.item, pos, None if item_x is None else item_x[pos], train=train, masks=mk.get("pos"))
    cn = tower_forward(state, "item", spec.item, neg_flat, None if item_x is None else item_x[neg_flat], train=train, masks=mk.get("neg"))
    t_u, t_p, t_n = cu["t"], cp["t"], cn["t"]
    D = t_u.shape[-1]
    if spec.mimic:
        Au, Ai = state["adaptive_mimic.user_augmented.weight"], state["adaptive_mimic.item_augmented.weight"]
        q_u, q_p, q_n = Au[users], Ai[pos], Ai[neg_flat]        # adaptive_mimic.py:97-105
        o_u, o_p, o_n = (t_u + q_u).astype(F32), (t_p + q_p).astype(F32), (t_n + q_n).astype(F32)
        if loss == "inbatch":
            L = inbatch_loss_forward_backward(o_u, o_p, t_u=t_u, t_p=t_p, q_u=q_u, q_p=q_p, lambda_u=lam_u, lambda_i=lam_i)
        else:
            L = loss_forward_backward(o_u, o_p, o_n.reshape(B, N, D), t_u=t_u, t_p=t_p, q_u=q_u, q_p=q_p,
                                      lambda_u=lam_u, lambda_i=lam_i)
    else:
        o_u, o_p, o_n = t_u, t_p, t_n
        L = inbatch_loss_forward_backward(o_u, o_p) if loss == "inbatch" else loss_forward_backward(o_u, o_p, o_n.reshape(B, N, D))
    if loss == "inbatch":
        L["do_n"], L["bce"] = np.zeros((0, D), dtype=F32), L["ce"]
    total = L["loss"]
    do_u, do_p, do_n = L["do_u"], L["do_p"], L["do_n"].reshape(B * N, D)
    cal = None
    if lam_c > 0:                                               # training.py:805-820
        comb_idx = np.concatenate([pos, neg_flat])
        comb_emb = np.concatenate([o_p, o_n], axis=0)
        cal, gcal = category_alignment_loss(comb_idx, comb_emb, cat_tensor, major)
        total = F32(total + F32(lam_c) * cal)
        do_p = (do_p + F32(lam_c) * gcal[:B]).astype(F32)
        do_n = (do_n + F32(lam_c) * gcal[B:]).astype(F32)

    grads: dict = {}
    sparse_rows: dict = {}
    # autograd runs the later-created graph first: negatives, then positives, then users
    tower_backward(state, "item", spec.item, cn, do_n, grads, sparse_rows, masks=mk.get("neg"), train=train)
    tower_backward(state, "item", spec.item, cp, do_p, grads, sparse_rows, masks=mk.get("pos"), train=train)
    tower_backward(state, "user", spec.user, cu, do_u, grads, sparse_rows, masks=mk.get("user"), train=train)
    if spec.mimic:
        gu = np.zeros_like(state["adaptive_mimic.user_augmented.weight"])
        gi = np.zeros_like(state["adaptive_mimic.item_augmented.weight"])
        np.add.at(gu, users, do_u + L["dq_u_extra"])            # embedding_dense_backward
        np.add.at(gi, neg_flat, do_n)
        np.add.at(gi, pos, do_p + L["dq_p_extra"])
        grads["adaptive_mimic.user_augmented.weight"] = gu
        grads["adaptive_mimic.item_augmented.weight"] = gi

    dense_names, sparse_names = _dense_param_names(state, spec)
    for name in dense_names:                                    # training.py:826 (dense optimiser first)
        if name in grads:
            g = grads[name]
        elif name in sparse_rows:                               # nn.Embedding(sparse=False): dense index_add
            g = np.zeros_like(state[name])
            for ix, v in sparse_rows[name]:
                np.add.at(g, ix, v)
        else:
            continue                                            # parameter received no gradient (p.grad is None)
        dense_step(optimizer, state[name], g, opt.slot(name), lr=lr, weight_decay=weight_decay,
                   momentum=momentum)
    touched = {}
    for name in sparse_names:                                   # training.py:827 (SparseAdam)
        if name not in sparse_rows:
            continue
        ix = np.concatenate([r[0] for r in sparse_rows[name]])
        v = np.concatenate([r[1] for r in sparse_rows[name]], axis=0)
        touched[name] = sparse_adam_step(state[name], opt.slot(name), ix, v, lr=lr, betas=betas)
    return {"loss": float(total), "bce": float(L["bce"]),
            "mimic_user": None if L["mimic_user"] is None else float(L["mimic_user"]),
            "mimic_item": None if L["mimic_item"] is None else float(L["mimic_item"]),
            "cal": None if cal is None else float(cal), "touched": touched}


# --------------------------------------------------------------------------------------------
# eval-mode helpers
# --------------------------------------------------------------------------------------------
def encode_users(state, spec: ModelSpec, idx, user_x):
    """eval-mode user tower + augment_users (training.py:1019-1026)."""
    c = tower_forward(state, "user", spec.user, idx, None if user_x is None else user_x[idx])
    t = c["t"]
    if spec.mimic:
        t = (t + state["adaptive_mimic.user_augmented.weight"][idx]).astype(F32)
    return t


def encode_items(state, spec: ModelSpec, idx, item_x):
    """_encode_item_embeddings (training.py:613-643): eval-mode item tower + augment_items."""
    c = tower_forward(state, "item", spec.item, idx, None if item_x is None else item_x[idx])
    t = c["t"]
    if spec.mimic:
        t = (t + state["adaptive_mimic.item_augmented.weight"][idx]).astype(F32)
    return t


def eval_loss(state, spec: ModelSpec, users, pos, neg, user_x, item_x):
    """_compute_loss (training.py:836-914): BCE only, on augmented eval-mode embeddings."""
    B, N = neg.shape
    o_u = encode_users(state, spec, users, user_x)
    o_p = encode_items(state, spec, pos, item_x)
    o_n = encode_items(state, spec, neg.reshape(-1), item_x).reshape(B, N, -1)
    return float(loss_forward_backward(o_u, o_p, o_n, need_grad=False)["bce"])
