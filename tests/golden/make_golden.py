#!/usr/bin/env python
"""Generate golden input/output vectors by RUNNING THE UNMODIFIED REFERENCE on CPU.

Run in the dev container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference (pure Python/PyTorch) is imported read-only from /root/reference with a
two-attribute matplotlib stub (src/reporting/plots.py:8 imports matplotlib, which the image
lacks).  Nothing from the reference is copied: only *numbers it produces* are stored, as small
.npz fixtures next to this script.  The fixtures pin the oracle (oracle/*.py) and, through it,
the CUDA path.

What is exercised, unmodified:
  * src.models.build_tower_encoder / AdaptiveMimicMechanism / TwoTowerModel
  * src.pipelines.training._train_one_epoch (one call per batch so that per-step losses are
    visible), _compute_loss, _encode_item_embeddings, _score_all_items_for_user,
    _evaluate_model (FAISS branch, with a brute-force stand-in for faiss.IndexFlatIP, which is an
    exact inner-product index; the host-side filtering code that consumes it runs unmodified),
    _collect_parameter_groups, _category_alignment_loss
  * torch.optim.AdamW / Adam / SGD + torch.optim.SparseAdam (the optimisers training.py:1311-1346 builds)
  * src.evaluation.compute_ranking_metrics
The negative sampler is patched only to *record* what it drew (training.py:730), because
negatives are an input of the hot path.
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def _import_reference():
    if not REF.exists():
        raise SystemExit("reference not mounted at /root/reference; goldens can only be regenerated in the dev container")
    sys.path.insert(0, str(REF))
    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
    import src.pipelines.training as training  # noqa: E402
    import src.models as models  # noqa: E402
    import src.evaluation as evaluation  # noqa: E402
    return training, models, evaluation


class _FlatIP:
    """Exact inner-product index: what faiss.IndexFlatIP computes (training.py:672-675,955-958).

    Ties are returned in (-score, +id) order, the canonical order of this project."""

    def __init__(self, dim: int) -> None:
        self.dim = dim
        self.x = np.zeros((0, dim), dtype=np.float32)

    def add(self, x: np.ndarray) -> None:
        self.x = np.concatenate([self.x, np.asarray(x, dtype=np.float32)], axis=0)

    def search(self, q: np.ndarray, k: int):
        import torch

        scores = (torch.from_numpy(np.ascontiguousarray(q)) @ torch.from_numpy(self.x).T).numpy()
        n = scores.shape[1]
        ids = np.full((q.shape[0], k), -1, dtype=np.int64)
        dist = np.full((q.shape[0], k), -np.inf, dtype=np.float32)
        for r in range(q.shape[0]):
            order = np.lexsort((np.arange(n), -scores[r].astype(np.float64)))[: min(k, n)]
            ids[r, : order.size] = order
            dist[r, : order.size] = scores[r, order]
        return dist, ids


def _fake_faiss():
    mod = types.ModuleType("faiss")
    mod.IndexFlatIP = _FlatIP

    def normalize_L2(x):
        n = np.sqrt((x.astype(np.float32) ** 2).sum(axis=1, keepdims=True))
        n[n == 0] = 1.0
        x /= n

    mod.normalize_L2 = normalize_L2
    return mod


def _tower_cfg(D, H, Hg, fe_type="mlp", fusion="gated", sparse=True, dropout=0.0, activation="relu"):
    cfg = {
        "type": "tower",
        "id_embedding": {"params": {"embedding_dim": D, "sparse": sparse}, "init": {"type": "normal", "std": 0.02}},
        "feature_encoder": {"type": fe_type, "hidden_dims": [H] if fe_type == "mlp" else None,
                            "activation": activation, "output_dim": D, "dropout": dropout},
        "fusion": fusion,
        "output_dim": D,
    }
    if Hg is not None:
        cfg["adaptive_mimic"] = {"hidden_dim": Hg}
    return cfg


def _features(rng, n, F, n_cat, n_auth):
    """Same column layout as build_item_feature_matrix (features.py:242-247):
    category multi-hot (weights 1, 1/2) | author one-hot | z-scored numerics/text stats."""
    x = np.zeros((n, F), dtype=np.float32)
    for r in range(n):
        cats = rng.choice(n_cat, size=rng.integers(1, 4), replace=False)
        for depth, c in enumerate(cats):
            x[r, c] = 1.0 / (depth + 1) if depth < 2 else 0.5
        x[r, n_cat + rng.integers(0, n_auth)] = 1.0
    x[:, n_cat + n_auth:] = rng.standard_normal((n, F - n_cat - n_auth)).astype(np.float32)
    return x


def _state_np(module):
    return {k: v.detach().cpu().numpy().copy() for k, v in module.state_dict().items()}


def make_train_case(name, *, seed, NU, NI, D, H, Hg, n_cat, n_auth, n_dense, B, N, steps,
                    optimizer="adamw", lr=1e-3, wd=0.01, betas=(0.9, 0.999), lambdas=(0.15, 0.15, 0.0),
                    fe_type="mlp", fusion="gated", mimic=True, sparse=True, tower_type="tower",
                    momentum=0.0, with_eval=False, similarity="dot", activation="relu"):
    import torch
    from torch import nn

    training, models, evaluation = _import_reference()
    training._seed_everything(seed)
    rng = np.random.default_rng(seed)
    F = n_cat + n_auth + n_dense
    device = torch.device("cpu")

    item_x = _features(rng, NI, F, n_cat, n_auth)
    # user rows = mean of a few item rows (features.py:302-313)
    user_x = np.stack([item_x[rng.choice(NI, size=4, replace=False)].mean(axis=0) for _ in range(NU)]).astype(np.float32)
    user_feat = torch.from_numpy(user_x)
    item_feat = torch.from_numpy(item_x)

    if tower_type == "embedding":
        ucfg = {"type": "embedding", "params": {"embedding_dim": D, "sparse": sparse}, "init": {"type": "normal", "std": 0.02}}
        icfg = dict(ucfg)
        ue = models.build_tower_encoder(ucfg, num_embeddings=NU, feature_dim=F, device=device)
        ie = models.build_tower_encoder(icfg, num_embeddings=NI, feature_dim=F, device=device)
    else:
        ue = models.build_tower_encoder(_tower_cfg(D, H, Hg, fe_type, fusion, sparse, activation=activation), num_embeddings=NU, feature_dim=F, device=device)
        ie = models.build_tower_encoder(_tower_cfg(D, H, Hg, fe_type, fusion, sparse, activation=activation), num_embeddings=NI, feature_dim=F, device=device)
    mm = models.AdaptiveMimicMechanism(num_users=NU, num_items=NI, embedding_dim=D, init_std=0.02) if mimic else None
    sim = training._select_similarity(similarity)
    model = models.TwoTowerModel(ue, ie, similarity=sim, adaptive_mimic=mm)
    # make biases non-trivial so that bias handling is pinned too
    with torch.no_grad():
        for n_, p in model.named_parameters():
            if n_.endswith(".bias"):
                p.copy_(0.05 * torch.randn_like(p))

    dense_params, sparse_params = training._collect_parameter_groups(model)
    optimizers = []
    if dense_params:
        if optimizer == "adamw":
            optimizers.append(torch.optim.AdamW(dense_params, lr=lr, weight_decay=wd))
        elif optimizer == "adam":
            optimizers.append(torch.optim.Adam(dense_params, lr=lr, weight_decay=wd))
        else:
            optimizers.append(torch.optim.SGD(dense_params, lr=lr, weight_decay=wd, momentum=momentum))
    if sparse_params:
        optimizers.append(torch.optim.SparseAdam(sparse_params, lr=lr, betas=betas))
    criterion = nn.BCEWithLogitsLoss()

    # interactions: zipf-ish item popularity so duplicate rows occur inside a batch
    pop = 1.0 / np.arange(1, NI + 1) ** 1.05
    pop /= pop.sum()
    positives: dict[int, set[int]] = {}
    batches = []
    for _ in range(steps):
        u = rng.integers(0, NU, size=B)
        it = rng.choice(NI, size=B, p=pop)
        batches.append((torch.from_numpy(u.astype(np.int64)), torch.from_numpy(it.astype(np.int64))))
        for a, b in zip(u.tolist(), it.tolist()):
            positives.setdefault(int(a), set()).add(int(b))

    cat_tensor = torch.from_numpy(rng.integers(0, 4, size=NI).astype(np.int64))
    cat_tensor[: NI // 2] = 0  # category 0 is the major one
    major = 0

    recorded_negs = []
    real_sampler = training.sample_negative_items

    def recording_sampler(*a, **k):
        out = real_sampler(*a, **k)
        recorded_negs.append(out.clone())
        return out

    training.sample_negative_items = recording_sampler

    out = {"meta_dims": np.array([NU, NI, D, H if H else 0, Hg if Hg else 0, F, B, N, steps], dtype=np.int64),
           "hyper": np.array([lr, wd, betas[0], betas[1], lambdas[0], lambdas[1], lambdas[2], momentum], dtype=np.float64),
           "user_x": user_x, "item_x": item_x, "cat_tensor": cat_tensor.numpy(), "major": np.array(major)}
    for k, v in _state_np(model).items():
        out["init/" + k] = v

    losses = []
    try:
        for s, (u, it) in enumerate(batches):
            loss = training._train_one_epoch(
                model, [(u, it)], optimizers=optimizers, criterion=criterion,
                negatives_per_positive=N, num_items=NI, user_positive_items=positives,
                user_features=user_feat, item_features=item_feat, device=device,
                gradient_clip_norm=None,
                loss_weights={"mimic_user": lambdas[0], "mimic_item": lambdas[1], "category_alignment": lambdas[2]},
                item_category_tensor=cat_tensor, major_category_id=major)
            losses.append(loss)
            out[f"step{s}/users"] = u.numpy()
            out[f"step{s}/pos"] = it.numpy()
            out[f"step{s}/neg"] = recorded_negs[-1].numpy()
            for k, v in _state_np(model).items():
                if s == 0 or s == steps - 1:
                    out[f"after{s}/" + k] = v
        out["losses"] = np.array(losses, dtype=np.float64)

        # optimiser state after the last step (exp_avg / exp_avg_sq by parameter name)
        name_of = {id(p): n for n, p in model.named_parameters()}
        for opt in optimizers:
            for p, st in opt.state.items():
                for key in ("exp_avg", "exp_avg_sq", "momentum_buffer"):
                    if key in st and st[key] is not None:
                        out[f"opt/{name_of[id(p)]}/{key}"] = st[key].detach().to_dense().numpy().copy() if st[key].is_sparse else st[key].detach().numpy().copy()

        if with_eval:
            # --- _compute_loss on one held-out batch
            u = torch.from_numpy(rng.integers(0, NU, size=B).astype(np.int64))
            it = torch.from_numpy(rng.integers(0, NI, size=B).astype(np.int64))
            recorded_negs.clear()
            ev = training._compute_loss(model, [(u, it)], criterion=criterion, negatives_per_positive=N,
                                        num_items=NI, user_positive_items=positives, user_features=user_feat,
                                        item_features=item_feat, device=device)
            out["eval/users"], out["eval/pos"], out["eval/neg"] = u.numpy(), it.numpy(), recorded_negs[-1].numpy()
            out["eval/loss"] = np.array(ev)
            # --- corpus encoding
            emb = training._encode_item_embeddings(model, num_items=NI, item_features=item_feat, device=device, batch_size=17)
            out["eval/item_embeddings"] = emb.numpy()
            # --- user embeddings (eval-mode tower + augment_users, training.py:1019-1026)
            model.eval()
            with torch.no_grad():
                allu = torch.arange(NU)
                ub = model.user_encoder({"indices": allu, "features": user_feat})
                uo = mm.augment_users(allu, ub) if mm is not None else ub
            out["eval/user_embeddings"] = uo.numpy()
            # --- brute-force top-k for a few users
            tk = []
            for uidx in range(4):
                tk.append(training._score_all_items_for_user(model, user_idx=uidx, top_k=10, num_items=NI,
                                                             user_features=user_feat, item_features=item_feat,
                                                             device=device, batch_size=23))
            out["eval/score_all_topk"] = np.array(tk, dtype=np.int64)
            # --- _evaluate_model through the FAISS branch with the exact flat-IP stand-in
            import pandas as pd
            training.faiss = _fake_faiss()
            res = training._prepare_faiss_resources(model, num_items=NI, item_features=item_feat, device=device,
                                                    similarity_module=sim, batch_size=19)
            val_u = rng.integers(0, NU, size=30)
            val_i = rng.integers(0, NI, size=30)
            val = pd.DataFrame({"user_idx": val_u, "item_idx": val_i})
            k_values = [5, 10, 20]
            preds, gts = training._evaluate_model(
                model, train_positive_map=positives, val_interactions=val, item_feature_tensor=item_feat,
                user_feature_tensor=user_feat, device=device, num_items=NI, candidate_samples=50,
                k_values=k_values, rng=np.random.default_rng(0), faiss_resources=res,
                faiss_search_k=4 * max(k_values))
            training.faiss = None
            users_sorted = sorted(preds)
            out["eval/val_users"] = val_u.astype(np.int64)
            out["eval/val_items"] = val_i.astype(np.int64)
            out["eval/pred_users"] = np.array(users_sorted, dtype=np.int64)
            width = max(len(preds[u_]) for u_ in users_sorted)
            pm = np.full((len(users_sorted), width), -1, dtype=np.int64)
            for r, u_ in enumerate(users_sorted):
                pm[r, : len(preds[u_])] = preds[u_]
            out["eval/pred_items"] = pm
            # train positives as CSR so that the blocked sets can be rebuilt
            keys = sorted(positives)
            out["eval/pos_keys"] = np.array(keys, dtype=np.int64)
            out["eval/pos_ptr"] = np.cumsum([0] + [len(positives[k_]) for k_ in keys]).astype(np.int64)
            out["eval/pos_vals"] = np.array([v for k_ in keys for v in sorted(positives[k_])], dtype=np.int64)
            met = evaluation.compute_ranking_metrics(preds, gts, k_values)
            out["eval/metrics"] = np.array([[met.recall[k_], met.precision[k_], met.ndcg[k_], met.hit_rate[k_], met.map[k_]] for k_ in k_values] + [[met.mrr] * 5], dtype=np.float64)
    finally:
        training.sample_negative_items = real_sampler

    path = OUT / f"{name}.npz"
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({path.stat().st_size/1024:.1f} KiB)  losses={losses}")


def make_optimizer_case():
    """torch.optim.SparseAdam / AdamW / Adam / SGD on hand-fed gradients: pins oracle/optim.py and
    the lazy-exact AdamW replay (rows that receive zero gradient for several steps)."""
    import torch

    torch.manual_seed(7)
    rng = np.random.default_rng(7)
    N, D, steps = 37, 8, 6
    out = {}
    p0 = (0.02 * torch.randn(N, D)).numpy()
    out["p0"] = p0
    idx_steps = [rng.integers(0, N, size=rng.integers(3, 12)).astype(np.int64) for _ in range(steps)]
    val_steps = [rng.standard_normal((ix.size, D)).astype(np.float32) * 0.01 for ix in idx_steps]
    for s in range(steps):
        out[f"idx{s}"] = idx_steps[s]
        out[f"val{s}"] = val_steps[s]
    # SparseAdam
    p = torch.nn.Parameter(torch.from_numpy(p0.copy()))
    opt = torch.optim.SparseAdam([p], lr=1e-3, betas=(0.9, 0.999))
    for s in range(steps):
        p.grad = torch.sparse_coo_tensor(torch.from_numpy(idx_steps[s])[None, :], torch.from_numpy(val_steps[s]), (N, D))
        opt.step()
        out[f"sparse_adam/p{s}"] = p.detach().numpy().copy()
    out["sparse_adam/exp_avg"] = opt.state[p]["exp_avg"].numpy().copy()
    out["sparse_adam/exp_avg_sq"] = opt.state[p]["exp_avg_sq"].numpy().copy()
    # dense optimisers with index_add'ed dense grads (what nn.Embedding(sparse=False) produces)
    for name, ctor in (("adamw", lambda q: torch.optim.AdamW([q], lr=1e-3, weight_decay=0.01)),
                       ("adam", lambda q: torch.optim.Adam([q], lr=1e-3, weight_decay=0.01)),
                       ("sgd", lambda q: torch.optim.SGD([q], lr=1e-3, weight_decay=0.01, momentum=0.9))):
        p = torch.nn.Parameter(torch.from_numpy(p0.copy()))
        opt = ctor(p)
        for s in range(steps):
            g = torch.zeros(N, D)
            g.index_add_(0, torch.from_numpy(idx_steps[s]), torch.from_numpy(val_steps[s]))
            p.grad = g
            opt.step()
            out[f"{name}/p{s}"] = p.detach().numpy().copy()
        for key in ("exp_avg", "exp_avg_sq", "momentum_buffer"):
            if key in opt.state[p]:
                out[f"{name}/{key}"] = opt.state[p][key].numpy().copy()
    path = OUT / "optim.npz"
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({path.stat().st_size/1024:.1f} KiB)")


def make_metrics_case():
    """Known answers held by the reference's own tests (tests/test_metrics.py:4-19)."""
    _, _, evaluation = _import_reference()
    preds = {0: [3, 2, 1], 1: [4, 5, 6]}
    gts = {0: {1, 2}, 1: {4}}
    ks = [1, 2, 3]
    m = evaluation.compute_ranking_metrics(preds, gts, ks)
    assert m.recall[1] == 0.5 and m.precision[1] == 0.5 and m.hit_rate[1] == 0.5 and abs(m.mrr - 0.75) < 1e-6
    np.savez_compressed(OUT / "metrics.npz",
                        recall=np.array([m.recall[k] for k in ks]), precision=np.array([m.precision[k] for k in ks]),
                        ndcg=np.array([m.ndcg[k] for k in ks]), hit=np.array([m.hit_rate[k] for k in ks]),
                        map=np.array([m.map[k] for k in ks]), mrr=np.array(m.mrr))
    print("wrote metrics.npz")


def main():
    import torch

    torch.set_num_threads(1)  # float goldens depend on the thread count
    common = dict(NU=40, NI=64, n_cat=9, n_auth=7, n_dense=5, B=24, N=3, steps=3)
    make_train_case("train_gated_mlp", seed=1234, D=16, H=32, Hg=None, with_eval=True, **common)
    make_train_case("train_gated_mlp_cal", seed=4321, D=16, H=32, Hg=24, lambdas=(0.15, 0.15, 0.01), **common)
    make_train_case("train_gated_mlp_cosine", seed=99, D=16, H=32, Hg=None, with_eval=True, similarity="cosine", **common)
    make_train_case("train_embedding_only", seed=11, D=16, H=None, Hg=None, tower_type="embedding", mimic=False, **common)
    make_train_case("train_dense_adam", seed=12, D=16, H=32, Hg=None, optimizer="adam", sparse=False, **common)
    make_train_case("train_sgd", seed=13, D=16, H=32, Hg=None, optimizer="sgd", momentum=0.9, **common)
    make_train_case("train_linear_sum", seed=14, D=16, H=None, Hg=None, fe_type="linear", fusion="sum", **common)
    # concat fusion (+ projection, encoders.py:211-217,242-244) with a GELU feature MLP (encoders.py:68-78)
    make_train_case("train_concat_gelu", seed=15, D=16, H=32, Hg=None, fusion="concat", activation="gelu", **common)
    make_optimizer_case()
    make_metrics_case()


if __name__ == "__main__":
    main()
