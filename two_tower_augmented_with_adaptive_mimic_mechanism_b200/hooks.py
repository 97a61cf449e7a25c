"""Replacements for the hot functions of the reference's `src/pipelines/training.py`, with the reference's
exact signatures and return values, so that `scripts/train.py` runs unchanged:

    import src.pipelines.training as training
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import hooks
    hooks.install(training)          # then training.run_training(cfg) as usual

  _train_one_epoch        reference training.py:700-833   -> FusedEngine.train_step per batch
  _compute_loss           reference training.py:836-914   -> FusedEngine.eval_loss per batch
  _encode_item_embeddings reference training.py:613-643   -> FusedEngine.encode_all
  _evaluate_model         reference training.py:917-1043  -> batched exact top-K + the reference's host filter
  _score_all_items_for_user reference training.py:330-384 -> exact top-K over the encoded corpus

Negative sampling stays the reference's `sample_negative_items` (it is an input of the hot path, SURVEY 8(a)
row S) unless a device sampler is supplied.
"""
from __future__ import annotations

from typing import Iterable

import numpy as np
import torch
from torch import nn

from .engine import FusedEngine
from .retrieval import FlatIPIndex, evaluate_users

_ENGINE_ATTR = "_ttam_engine"


def _hyper_from_optimizers(optimizers) -> dict:
    """Recover the hyper-parameters `_run_single_experiment` chose (training.py:1311-1350) from the torch
    optimisers it built; they are only used as configuration carriers here."""
    hp = dict(optimizer="adamw", lr=1e-3, weight_decay=0.0, momentum=0.0, dense_betas=(0.9, 0.999),
              sparse_betas=(0.9, 0.999))
    for opt in optimizers:
        g = opt.param_groups[0]
        if isinstance(opt, torch.optim.SparseAdam):
            hp["sparse_betas"] = tuple(g["betas"])
            hp["lr"] = g["lr"]
        else:
            hp["lr"] = g["lr"]
            hp["weight_decay"] = g.get("weight_decay", 0.0)
            if isinstance(opt, torch.optim.AdamW):
                hp["optimizer"], hp["dense_betas"] = "adamw", tuple(g["betas"])
            elif isinstance(opt, torch.optim.Adam):
                hp["optimizer"], hp["dense_betas"] = "adam", tuple(g["betas"])
            elif isinstance(opt, torch.optim.SGD):
                hp["optimizer"], hp["momentum"] = "sgd", g.get("momentum", 0.0)
            else:
                raise ValueError(f"Unsupported optimizer: {type(opt).__name__}")
    return hp


def engine_for(model, optimizers=(), **overrides) -> FusedEngine:
    eng = getattr(model, _ENGINE_ATTR, None)
    if eng is None:
        hp = _hyper_from_optimizers(optimizers)
        hp.update(overrides)
        eng = FusedEngine(model, **hp)
        object.__setattr__(model, _ENGINE_ATTR, eng)
    return eng


def _publish_optimizer_state(eng: FusedEngine, optimizers) -> None:
    """Expose the engine's moments through the torch optimisers so that `optimizer.state_dict()` in
    `_save_checkpoint` (training.py:150-182) keeps its layout."""
    st = eng.optimizer_state()
    by_id = {id(p): n for n, p in eng.model.named_parameters()}
    for opt in optimizers:
        for group in opt.param_groups:
            for p in group["params"]:
                s = st.get(by_id.get(id(p)))
                if s is None:
                    continue
                slot = opt.state[p]
                slot["step"] = torch.tensor(float(s["step"])) if not isinstance(opt, torch.optim.SparseAdam) else s["step"]
                if s["exp_avg"] is not None:
                    key = "momentum_buffer" if isinstance(opt, torch.optim.SGD) else "exp_avg"
                    slot[key] = s["exp_avg"]
                if s["exp_avg_sq"] is not None and not isinstance(opt, torch.optim.SGD):
                    slot["exp_avg_sq"] = s["exp_avg_sq"]


def _sampler():
    try:
        from src.data.samplers import sample_negative_items  # the reference's sampler, when importable
        return sample_negative_items
    except Exception:  # pragma: no cover - standalone use
        from .sampler import sample_negative_items
        return sample_negative_items


def _train_one_epoch(model, dataloader, *, optimizers, criterion, negatives_per_positive, num_items,
                     user_positive_items, user_features, item_features, device, gradient_clip_norm=None,
                     loss_weights=None, item_category_tensor=None, major_category_id=None) -> float:
    if gradient_clip_norm is not None:
        raise NotImplementedError("gradient clipping is not supported by the fused step (reference default: null)")
    if not isinstance(criterion, nn.BCEWithLogitsLoss):
        raise ValueError("the fused step implements nn.BCEWithLogitsLoss (reference training.py:1366) only")
    model.train()
    eng = engine_for(model, optimizers)
    w = loss_weights or {}
    eng.lambda_u, eng.lambda_i = float(w.get("mimic_user", 0.0)), float(w.get("mimic_item", 0.0))
    eng.lambda_c = float(w.get("category_alignment", 0.0))
    eng.cat_tensor = None if item_category_tensor is None else item_category_tensor.to(device)
    eng.major = major_category_id
    sample = _sampler()
    losses, sizes = [], []
    for users, pos in dataloader:
        users, pos = users.to(device), pos.to(device)
        neg = sample(users, num_items=num_items, positives=user_positive_items,
                     num_negatives=negatives_per_positive, device=device)
        loss = eng.train_step(users, pos, neg, user_features, item_features)
        losses.append(loss[0:1].clone())       # no host sync inside the loop (reference syncs per step, :830)
        sizes.append(users.shape[0])
    eng.flush()
    _publish_optimizer_state(eng, optimizers)
    if not losses:
        return 0.0
    per_step = torch.cat(losses).double().cpu().numpy()
    total = int(np.sum(sizes))
    return float(np.dot(per_step, np.asarray(sizes, dtype=np.float64)) / max(total, 1))


def _compute_loss(model, dataloader, *, criterion, negatives_per_positive, num_items, user_positive_items,
                  user_features, item_features, device) -> float:
    model.eval()
    eng = engine_for(model)
    sample = _sampler()
    losses, sizes = [], []
    for users, pos in dataloader:
        users, pos = users.to(device), pos.to(device)
        neg = sample(users, num_items=num_items, positives=user_positive_items,
                     num_negatives=negatives_per_positive, device=device)
        losses.append(eng.eval_loss(users, pos, neg, user_features, item_features)[1:2].clone())
        sizes.append(users.shape[0])
    if not losses:
        return 0.0
    per_step = torch.cat(losses).double().cpu().numpy()
    return float(np.dot(per_step, np.asarray(sizes, dtype=np.float64)) / max(int(np.sum(sizes)), 1))


def _encode_item_embeddings(model, *, num_items, item_features, device, batch_size: int = 8192) -> torch.Tensor:
    if num_items == 0:
        return torch.empty((0, 0), dtype=torch.float32)
    model.eval()
    eng = engine_for(model)
    feats = item_features if (item_features is not None and item_features.numel() > 0) else None
    return eng.encode_all("item", feats, chunk=max(int(batch_size), 65536)).cpu()


def _evaluate_model(model, *, train_positive_map, val_interactions, item_feature_tensor, user_feature_tensor,
                    device, num_items, candidate_samples, k_values: Iterable[int], rng, faiss_resources=None,
                    faiss_search_k: int = 0):
    """Exact full-corpus retrieval for every validation user (the FAISS branch's semantics; the sampling
    branch is an approximation of it that exists only because FAISS is optional in the reference)."""
    if val_interactions.empty:
        return {}, {}
    model.eval()
    k_values = list(k_values)
    eng = engine_for(model)
    gts: dict[int, set[int]] = {}
    for user_idx, group in val_interactions.groupby("user_idx"):
        gt = set(map(int, group["item_idx"].tolist()))
        if gt:
            gts[int(user_idx)] = gt
    if not gts:
        return {}, {}
    users = list(gts)
    ifeat = item_feature_tensor if (item_feature_tensor is not None and item_feature_tensor.numel() > 0) else None
    ufeat = user_feature_tensor if (user_feature_tensor is not None and user_feature_tensor.numel() > 0) else None
    corpus = eng.encode_all("item", ifeat)
    cosine = isinstance(model.similarity, nn.CosineSimilarity)
    if faiss_resources is not None and "normalize" in faiss_resources:
        cosine = bool(faiss_resources["normalize"])
    index = FlatIPIndex(corpus, normalize=cosine)
    uidx = torch.tensor(users, device=device, dtype=torch.long)
    uemb = eng.encode("user", uidx, ufeat)
    preds = evaluate_users(index, uemb, users, gts, train_positive_map, k_values, search_k=int(faiss_search_k))
    return preds, gts


def _score_all_items_for_user(model, *, user_idx: int, top_k: int, num_items: int, user_features, item_features,
                              device, batch_size: int = 50000) -> list[int]:
    if num_items == 0:
        return []
    model.eval()
    eng = engine_for(model)
    ifeat = item_features if (item_features is not None and item_features.numel() > 0) else None
    ufeat = user_features if (user_features is not None and user_features.numel() > 0) else None
    corpus = eng.encode_all("item", ifeat)
    index = FlatIPIndex(corpus, normalize=isinstance(model.similarity, nn.CosineSimilarity))
    q = eng.encode("user", torch.tensor([user_idx], device=device, dtype=torch.long), ufeat)
    ids, _ = index.search(q, min(top_k, num_items))
    return ids[0].cpu().tolist()


def install(training_module) -> None:
    """Assign the fused implementations onto the reference's `src.pipelines.training` module and swap its
    model classes for the B200 ones; the reference file itself is untouched."""
    from . import models
    for name in ("_train_one_epoch", "_compute_loss", "_encode_item_embeddings", "_evaluate_model",
                 "_score_all_items_for_user"):
        setattr(training_module, name, globals()[name])
    for name in ("AdaptiveMimicMechanism", "TwoTowerModel", "build_tower_encoder"):
        if hasattr(training_module, name):
            setattr(training_module, name, getattr(models, name))
