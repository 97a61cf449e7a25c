"""numpy restatement of the reference retrieval / evaluation path.

TEST INFRASTRUCTURE (see oracle/__init__.py).
  * exact inner-product search  = faiss.IndexFlatIP (faiss-cpu>=1.7.4, pyproject.toml:24; call sites
    training.py:672-675,955-958) and the brute-force `_score_all_items_for_user` (training.py:330-384)
  * host-side filtering         = `_evaluate_model._retrieve_with_faiss` (training.py:944-972)
  * metrics                     = src/evaluation/metrics.py:49-116

Canonical result order of this project: descending score, ascending item id on ties.
Canonical score: fp32, products rounded to fp32 and accumulated sequentially over d = 0..D-1
(for bf16 inputs every product is exact in fp32, so this is also what an FMA chain gives).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def canonical_scores(q, items, chunk=1 << 16):
    """[Q,D] x [N,D] -> [Q,N] fp32, sequential accumulation order d = 0..D-1."""
    q = np.ascontiguousarray(q, dtype=F32)
    items = np.ascontiguousarray(items, dtype=F32)
    Q, D = q.shape
    N = items.shape[0]
    out = np.empty((Q, N), dtype=F32)
    for s in range(0, N, chunk):
        it = items[s:s + chunk]
        acc = np.zeros((Q, it.shape[0]), dtype=F32)
        for d in range(D):
            acc += q[:, d:d + 1] * it[None, :, d]
        out[:, s:s + chunk] = acc
    return out


def topk_canonical(scores, k, ids=None):
    """Top-k of each row under (-score, +id).  Returns (ids [Q,k] int64, scores [Q,k])."""
    Q, N = scores.shape
    k = min(k, N)
    base = np.arange(N, dtype=np.int64) if ids is None else np.asarray(ids, dtype=np.int64)
    out_i = np.empty((Q, k), dtype=np.int64)
    out_s = np.empty((Q, k), dtype=F32)
    for r in range(Q):
        order = np.lexsort((base, -scores[r].astype(np.float64)))[:k]
        out_i[r] = base[order]
        out_s[r] = scores[r, order]
    return out_i, out_s


def l2_normalize(x, eps=1e-12):
    """F.normalize(x, dim=-1) / faiss.normalize_L2."""
    n = np.sqrt((x.astype(F32) ** 2).sum(axis=1, keepdims=True, dtype=F32))
    return (x / np.maximum(n, F32(eps))).astype(F32)


def score_all_items_topk(user_emb, item_emb, k, cosine=False):
    """`_score_all_items_for_user` (training.py:330-384) result set, in canonical order."""
    u = user_emb.reshape(1, -1)
    it = item_emb
    if cosine:
        u, it = l2_normalize(u), l2_normalize(it)
    return topk_canonical(canonical_scores(u, it), k)[0][0]


def evaluate_flat_ip(user_emb_of, item_emb, val_users, val_items, train_pos, k_values, *,
                     search_k=0, cosine=False):
    """`_evaluate_model` through the FAISS branch (training.py:917-1043).

    user_emb_of: callable user_idx -> [D] embedding (eval-mode tower + augment_users)."""
    max_k = max(k_values)
    items = l2_normalize(item_emb) if cosine else item_emb
    groups: dict[int, list[int]] = {}
    for u, i in zip(np.asarray(val_users).tolist(), np.asarray(val_items).tolist()):
        groups.setdefault(int(u), []).append(int(i))
    preds, gts = {}, {}
    for u in sorted(groups):                                   # DataFrame.groupby sorts keys
        gt = set(map(int, groups[u]))
        if not gt:
            continue
        gts[u] = gt
        q = np.asarray(user_emb_of(u), dtype=F32).reshape(1, -1)
        if cosine:
            q = l2_normalize(q)
        blocked = set(train_pos.get(u, set()))
        search_limit = max(max_k + len(gt), 1)                  # training.py:956
        sk = max(search_k, search_limit + len(blocked))         # training.py:957
        cand = topk_canonical(canonical_scores(q, items), sk)[0][0].tolist()
        filtered, seen = [], set()
        for it in cand:                                         # training.py:961-968
            if it in blocked or it in seen or it < 0:
                continue
            filtered.append(int(it))
            seen.add(int(it))
            if len(filtered) >= search_limit:
                break
        for it in gt:                                           # training.py:969-971
            if it not in seen:
                filtered.append(it)
        preds[u] = filtered[:max_k]
    return preds, gts


def _dcg(rel):
    return sum(r / np.log2(i + 2) for i, r in enumerate(rel))


def ranking_metrics(preds, gts, k_values):
    """compute_ranking_metrics (src/evaluation/metrics.py:74-116) as a plain dict."""
    acc = {m: {k: [] for k in k_values} for m in ("recall", "precision", "ndcg", "hit_rate", "map")}
    mrr = []
    ks = sorted(k_values)
    max_k = max(ks)
    for u, pred in preds.items():
        gt = gts.get(u, set())
        if not gt:
            continue
        for k in ks:
            topk = pred[:k]
            hits = len(set(topk) & gt)
            acc["recall"][k].append(hits / max(len(gt), 1))
            acc["precision"][k].append(hits / max(k, 1))
            acc["hit_rate"][k].append(1.0 if hits > 0 else 0.0)
            rel = [1 if it in gt else 0 for it in pred[:k]]
            ideal = _dcg([1] * min(k, len(gt)))
            acc["ndcg"][k].append(0.0 if ideal == 0 else _dcg(rel) / ideal)
            h, sp = 0, 0.0
            for i, it in enumerate(pred[:k], start=1):
                if it in gt:
                    h += 1
                    sp += h / i
            acc["map"][k].append(sp / min(len(gt), k))
        rr = 0.0
        for i, it in enumerate(pred[:max_k], start=1):
            if it in gt:
                rr = 1.0 / i
                break
        mrr.append(rr)
    out = {m: {k: (float(np.mean(v)) if v else 0.0) for k, v in d.items()} for m, d in acc.items()}
    out["mrr"] = float(np.mean(mrr)) if mrr else 0.0
    return out
