"""CPU PyTorch restatement of the reference's training step and retrieval, used ONLY as the timed CPU
baseline (`bench.py` cpu_baseline / `--impl reference`) and as a second checker in tests.

TEST / BASELINE INFRASTRUCTURE (see oracle/__init__.py).  The reference is pure Python and cannot travel
to the GPU box (/root/reference does not exist there), so the "reference CPU PyTorch path" the north star
asks to be timed beside the GPU is restated here with the same stock torch building blocks the reference
uses — nn.Embedding(sparse=True), nn.Linear, autograd, torch.optim.AdamW + torch.optim.SparseAdam,
nn.BCEWithLogitsLoss, the per-row Python negative sampler and torch.topk — in the same order:
  towers            reference encoders.py:102-168,221-255
  mimic             reference adaptive_mimic.py:40-105
  step              reference training.py:726-831   (sampler: samplers.py:36-83)
  param groups      reference training.py:276-309, optimisers :1311-1346
  retrieval         reference training.py:613-643 (encode), 330-384 / IndexFlatIP (exact inner product)
Parity: pinned by tests/test_oracle_golden.py::test_torch_port_matches_reference_golden against the same
fixtures the numpy oracle is pinned to.
"""
from __future__ import annotations

import torch
from torch import nn


class Tower(nn.Module):
    def __init__(self, n, D, F, H, Hg, sparse=True, dropout=0.0):
        super().__init__()
        self.embedding = nn.Embedding(n, D, sparse=sparse)
        nn.init.normal_(self.embedding.weight, std=0.02)
        self.has_features = F > 0
        if self.has_features:
            layers = [nn.Linear(F, H), nn.ReLU()]
            if dropout:
                layers.append(nn.Dropout(dropout))
            layers.append(nn.Linear(H, D))
            self.mlp = nn.Sequential(*layers)
            self.gate = nn.Sequential(nn.Linear(2 * D, Hg), nn.ReLU(), nn.Linear(Hg, D), nn.Sigmoid())
            for m in list(self.mlp):
                if isinstance(m, nn.Linear):
                    nn.init.xavier_uniform_(m.weight)

    def forward(self, idx, x=None):
        e = self.embedding(idx)
        if not self.has_features or x is None:
            return e
        f = self.mlp(x)
        g = self.gate(torch.cat([e, f], dim=-1))
        return g * e + (1.0 - g) * f


class Model(nn.Module):
    def __init__(self, NU, NI, D, F, H, Hg, sparse=True, dropout=0.0, mimic=True):
        super().__init__()
        self.user = Tower(NU, D, F, H, Hg, sparse, dropout)
        self.item = Tower(NI, D, F, H, Hg, sparse, dropout)
        self.aug_user = nn.Embedding(NU, D) if mimic else None
        self.aug_item = nn.Embedding(NI, D) if mimic else None
        if mimic:
            nn.init.normal_(self.aug_user.weight, std=0.02)
            nn.init.normal_(self.aug_item.weight, std=0.02)

    # reference state_dict key <-> attribute map, so that golden states can be loaded
    KEYMAP = (("user_encoder.embedding.", "user.embedding."), ("item_encoder.embedding.", "item.embedding."),
              ("user_encoder.feature_encoder.network.", "user.mlp."), ("item_encoder.feature_encoder.network.", "item.mlp."),
              ("user_encoder.adaptive_mimic.gate_network.", "user.gate."), ("item_encoder.adaptive_mimic.gate_network.", "item.gate."),
              ("adaptive_mimic.user_augmented.", "aug_user."), ("adaptive_mimic.item_augmented.", "aug_item."))

    def load_reference_state(self, state):
        mapped = {}
        for k, v in state.items():
            for a, b in self.KEYMAP:
                if k.startswith(a):
                    mapped[b + k[len(a):]] = torch.as_tensor(v).clone()
        self.load_state_dict(mapped, strict=True)

    def reference_state(self):
        out = {}
        for k, v in self.state_dict().items():
            for a, b in self.KEYMAP:
                if k.startswith(b):
                    out[a + k[len(b):]] = v.detach().numpy().copy()
        return out


def build_optimizers(model: Model, lr=1e-3, weight_decay=0.01, betas=(0.9, 0.999)):
    sparse = [p for p in (model.user.embedding.weight, model.item.embedding.weight) if model.user.embedding.sparse]
    ids = {id(p) for p in sparse}
    dense = [p for p in model.parameters() if id(p) not in ids]
    opts = [torch.optim.AdamW(dense, lr=lr, weight_decay=weight_decay)]
    if sparse:
        opts.append(torch.optim.SparseAdam(sparse, lr=lr, betas=betas))
    return opts


def sample_negatives(users, num_items, positives, n):
    """Per-row loop with rejection of known positives and <= 10 resampling rounds (reference samplers.py:36-83)."""
    out = torch.empty((users.shape[0], n), dtype=torch.long)
    for r, u in enumerate(users.tolist()):
        s = torch.randint(0, num_items, (n,))
        pos = positives.get(int(u)) if positives else None
        if pos:
            pt = torch.tensor(sorted(pos), dtype=torch.long)
            bad = torch.isin(s, pt)
            tries = 0
            while bad.any():
                s[bad] = torch.randint(0, num_items, (int(bad.sum()),))
                bad = torch.isin(s, pt)
                tries += 1
                if tries > 10:
                    raise RuntimeError("Exceeded resampling attempts while drawing negatives.")
        out[r] = s
    return out


def train_step(model: Model, opts, users, pos, neg, user_x, item_x, lambdas=(0.15, 0.15)):
    """One iteration of the reference loop body; returns the loss as a Python float (the reference syncs per step)."""
    model.train()
    for o in opts:
        o.zero_grad()
    B, N = neg.shape
    t_u = model.user(users, None if user_x is None else user_x.index_select(0, users))
    t_p = model.item(pos, None if item_x is None else item_x.index_select(0, pos))
    nf = neg.reshape(-1)
    t_n = model.item(nf, None if item_x is None else item_x.index_select(0, nf))
    lu = li = None
    if model.aug_user is not None:
        q_u, q_p = model.aug_user(users), model.aug_item(pos)
        o_u, o_p = t_u + q_u, t_p + q_p
        lu = torch.nn.functional.mse_loss(q_u, t_p.detach())
        li = torch.nn.functional.mse_loss(q_p, t_u.detach())
        o_n = t_n + model.aug_item(nf)
    else:
        o_u, o_p, o_n = t_u, t_p, t_n
    o_n = o_n.view(B, N, -1)
    logits = torch.cat([(o_u * o_p).sum(-1), (o_u.unsqueeze(1) * o_n).sum(-1).reshape(-1)])
    labels = torch.cat([torch.ones(B), torch.zeros(B * N)])
    loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, labels)
    if lu is not None and lambdas[0] > 0:
        loss = loss + lambdas[0] * lu
    if li is not None and lambdas[1] > 0:
        loss = loss + lambdas[1] * li
    loss.backward()
    for o in opts:
        o.step()
    return float(loss.item())


@torch.no_grad()
def encode_items(model: Model, item_x, batch=8192):
    model.eval()
    out = []
    for s in range(0, model.item.embedding.num_embeddings, batch):
        idx = torch.arange(s, min(s + batch, model.item.embedding.num_embeddings))
        t = model.item(idx, None if item_x is None else item_x.index_select(0, idx))
        if model.aug_item is not None:
            t = t + model.aug_item(idx)
        out.append(t)
    return torch.cat(out)


@torch.no_grad()
def flat_ip_topk(queries, items, k, chunk=1024):
    """Exact inner-product search, `chunk` queries per GEMM (what IndexFlatIP.search does with BLAS + heaps)."""
    ids = []
    for s in range(0, queries.shape[0], chunk):
        sc = queries[s:s + chunk] @ items.T
        ids.append(torch.topk(sc, k=min(k, items.shape[0]), dim=1).indices)
    return torch.cat(ids)
