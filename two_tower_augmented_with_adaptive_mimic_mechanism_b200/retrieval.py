"""Full-corpus retrieval on the GPU: exact inner-product top-K for many queries at once.

Replaces the reference's three per-user paths — `faiss.IndexFlatIP.search` (training.py:944-972), the
candidate-sampling fallback (:974-1009) and `_score_all_items_for_user` (:330-384) — by one batched
launch sequence.  Result order is canonical: descending score, ascending item id on ties.
"""
from __future__ import annotations

from typing import Iterable, Mapping, Sequence

import numpy as np
import torch

from . import functional as F


def l2_normalize(x: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """F.normalize(x, dim=-1) / faiss.normalize_L2 (reference training.py:364-370, 669-671)."""
    n = torch.sqrt((x * x).sum(dim=1, keepdim=True))
    return x / n.clamp_min(eps)


class FlatIPIndex:
    """Exact inner-product index over an item corpus resident in HBM (the IndexFlatIP stand-in).

    `dtype=torch.float32` keeps the reference's fp32 scores: searches with k <= 128 run their candidate pass on the
    tcgen05 tensor cores over a [hi | lo] bf16 split of the corpus (kept beside it, built on first use: 4 x D bytes per item)
    and re-score the survivors from the fp32 rows - ids and scores are bit-identical to the SIMT kernel, which deeper
    or paged searches still use (`tensor_cores=False` forces it everywhere).  `torch.bfloat16` rounds corpus and queries
    to bf16 and scores them on the tensor cores directly (BASELINE config 3)."""

    def __init__(self, item_embeddings: torch.Tensor, *, normalize: bool = False, dtype=torch.float32,
                 id_offset: int = 0, tensor_cores: bool = True) -> None:
        if not item_embeddings.is_cuda:
            raise F._lib.TtamError("FlatIPIndex needs a CUDA corpus; there is no CPU path")
        x = item_embeddings.float()
        if normalize:
            x = l2_normalize(x)
        self.normalize = normalize
        self.dtype = dtype
        self.id_offset = int(id_offset)
        self.items = x.contiguous() if dtype == torch.float32 else F.cast_bf16(x.contiguous())
        self.ntotal, self.d = self.items.shape
        self.tensor_cores = bool(tensor_cores) and dtype == torch.float32 and self.d <= F.TOPK_TC_MAX_D
        self._items_split = None

    @property
    def max_k(self) -> int:
        """Largest k one launch sequence returns (deeper rankings: `search(..., after=)` page by page, fp32 only)."""
        return F.TOPK_F32_MAX_K if self.dtype == torch.float32 else F.TOPK_BF16_MAX_K

    def _prepare(self, queries: torch.Tensor) -> torch.Tensor:
        q = queries.float()
        if self.normalize:
            q = l2_normalize(q)
        return q.contiguous() if self.dtype == torch.float32 else F.cast_bf16(q.contiguous())

    def search(self, queries: torch.Tensor, k: int, *, after=None):
        """queries [Q, D] -> (ids [Q, k] int64, scores [Q, k] fp32); missing slots (k > ntotal) hold id -1.
        after = (scores [Q], ids [Q]): only results ranked strictly after that pair (fp32 index)."""
        q = self._prepare(queries)
        k_eff = min(int(k), self.ntotal)
        if self.tensor_cores and after is None and k_eff <= F.TOPK_BF16_MAX_K:
            if self._items_split is None:
                self._items_split = F.split_bf16x3(self.items, item_layout=True)
            ids, scores = F.topk_f32_tc(q, self.items, self._items_split, k_eff, id_offset=self.id_offset)
        else:
            ids, scores = F.topk(q, self.items, k_eff, id_offset=self.id_offset, after=after)
        if k_eff < k:
            pad_i = torch.full((q.shape[0], k - k_eff), -1, dtype=torch.int64, device=q.device)
            pad_s = torch.full((q.shape[0], k - k_eff), float("-inf"), dtype=torch.float32, device=q.device)
            ids, scores = torch.cat([ids, pad_i], 1), torch.cat([scores, pad_s], 1)
        return ids, scores


    def search_deep(self, queries: torch.Tensor, k: int):
        """`search` for k beyond `max_k`: the ranking is walked down page by page (each page = one launch sequence that
        only admits items ranked after the previous page's last result).  Returns (ids, scores) as numpy [Q, k]."""
        k = min(int(k), self.ntotal)
        page = min(k, self.max_k)
        ids, scores = self.search(queries, page)
        out_i, out_s = [ids], [scores]
        got = page
        if got < k and self.dtype != torch.float32:
            raise F._lib.TtamError(f"a ranking deeper than {self.max_k} needs an fp32 index (dtype=torch.float32)")
        while got < k:
            ids, scores = self.search(queries, min(self.max_k, k - got), after=(out_s[-1][:, -1], out_i[-1][:, -1]))
            out_i.append(ids); out_s.append(scores)
            got += ids.shape[1]
        return torch.cat(out_i, 1).cpu().numpy(), torch.cat(out_s, 1).cpu().numpy()


def score_candidates(queries: torch.Tensor, corpus: torch.Tensor, cand: torch.Tensor, *, cosine: bool = False) -> torch.Tensor:
    """scores [R, C] of query r against corpus[cand[r, c]] (-inf where cand < 0): the scoring half of the reference's
    candidate-sampling evaluation (training.py:986-1005); with `cosine` both sides are L2-normalised first (:999-1003)."""
    q, it = queries.float(), corpus.float()
    if cosine:
        q, it = l2_normalize(q), l2_normalize(it)
    return F.score_pairs(q, it, cand)


def filter_candidates(candidate_ids: list[int], blocked: set[int], ground_truth: set[int], max_k: int) -> list[int]:
    """Host-side post-filter of `_retrieve_with_faiss` (reference training.py:959-972), semantics kept verbatim:
    drop blocked / duplicate / negative ids, stop at max_k + |gt|, append unseen ground truth, truncate."""
    limit = max(max_k + len(ground_truth), 1)
    kept: list[int] = []
    seen: set[int] = set()
    for it in candidate_ids:
        if it < 0 or it in blocked or it in seen:
            continue
        kept.append(int(it))
        seen.add(int(it))
        if len(kept) >= limit:
            break
    for it in ground_truth:
        if it not in seen:
            kept.append(it)
    return kept[:max_k]


def filter_block(ids: np.ndarray, need: Sequence[int], users: Sequence[int], ground_truth: Mapping[int, set],
                 train_positive_map: Mapping[int, set], max_k: int, deeper=None) -> dict:
    """`filter_candidates` for a whole block of users at once (the reference filters one user per Python
    iteration, training.py:959-972).  ids [n, K] int64: row r holds user users[r]'s candidates, best first (-1 = none);
    only its first need[r] columns count (the reference asks FAISS for search_k = need[r] results).
    Rows that keep at least max_k candidates - nearly all of them - are handled by array operations: blocked-item
    membership is one searchsorted over (row, item) keys, the survivors' first max_k columns one stable argsort.  Rows
    that keep fewer (their ground truth gets appended) or repeat an id go through `filter_candidates` itself, so the
    result is the reference's for every row.
    deeper(rows, k) -> {row: its first k candidates}: called once, for the rare rows that keep fewer than max_k of their
    K columns although the reference would have asked for more than K results (a user with more training positives than
    one launch returns)."""
    ids = np.asarray(ids, dtype=np.int64)
    n, K = ids.shape
    need = np.asarray(need, dtype=np.int64).reshape(n, 1)
    valid = (ids >= 0) & (np.arange(K, dtype=np.int64)[None, :] < need)
    rows, items = [], []
    for r, u in enumerate(users):
        b = train_positive_map.get(u)
        if b:
            rows.append(np.full(len(b), r, dtype=np.int64))
            items.append(np.fromiter(b, dtype=np.int64, count=len(b)))
    if rows and ids.size:
        rows, items = np.concatenate(rows), np.concatenate(items)
        M = int(max(ids.max(), items.max())) + 1
        bk = np.sort(rows * M + items)
        ck = np.arange(n, dtype=np.int64)[:, None] * M + np.where(valid, ids, 0)
        pos = np.minimum(np.searchsorted(bk, ck), bk.size - 1)
        valid &= bk[pos] != ck
    out = {}
    take = min(max_k, K)
    order = np.argsort(~valid, axis=1, kind="stable")[:, :take]          # columns of the survivors first, in their order
    sel = np.take_along_axis(ids, order, axis=1)
    fast = valid.sum(axis=1) >= max_k
    if take > 1:
        srt = np.sort(sel, axis=1)
        fast &= ~(srt[:, 1:] == srt[:, :-1]).any(axis=1)                 # a repeated id: the reference drops it, go slow
    # rows that need a longer candidate list than the K columns given: fetched in ONE call, deeper(rows, k) -> {row: list}
    short = [r for r in range(n) if not fast[r] and int(need[r, 0]) > K] if deeper is not None else []
    longer = deeper(short, max(int(need[r, 0]) for r in short)) if short else {}
    for r, u in enumerate(users):
        if fast[r]:
            out[u] = sel[r].tolist()
        else:
            k_r = int(need[r, 0])
            row = longer[r][:k_r] if r in longer else ids[r, :k_r].tolist()
            out[u] = filter_candidates(row, set(train_positive_map.get(u, ())), ground_truth[u], max_k)
    return out


def evaluate_users(index: FlatIPIndex, user_embeddings: torch.Tensor, user_ids: list[int],
                   ground_truth: Mapping[int, set[int]], train_positive_map: Mapping[int, set[int]],
                   k_values: Iterable[int], search_k: int = 0, query_block: int = 8192):
    """Batched `_evaluate_model` (FAISS branch): one top-K launch sequence per block of users, then the
    reference's filtering on the host, a block at a time (`filter_block`).  user_embeddings[r] belongs to user_ids[r]."""
    max_k = max(k_values)
    preds: dict[int, list[int]] = {}
    # the reference asks for search_k = max(faiss_search_k, max_k + |gt| + |blocked|) per user; a block uses its max
    for s in range(0, len(user_ids), query_block):
        blk = user_ids[s:s + query_block]
        need = [max(search_k, max(max_k + len(ground_truth[u]), 1) + len(train_positive_map.get(u, ()))) for u in blk]
        # one launch returns at most index.max_k results per query; search_k grows with a user's number of training
        # positives (training.py:956-958), so a heavy user can ask for more: such a row is only walked further down
        # (search_deep) when its first max_k columns do not already hold max_k unblocked items
        # (an fp32 index first asks for at most the 128 results its tensor-core pass returns; the rows that fall short go
        # through the SIMT kernel together)
        k_blk = min(max(need), index.ntotal, F.TOPK_BF16_MAX_K if getattr(index, "tensor_cores", False) else index.max_k)
        q_blk = user_embeddings[s:s + len(blk)]
        ids, _ = index.search(q_blk, k_blk)

        def deeper(rows, k, q_blk=q_blk):
            sel = torch.as_tensor(rows, dtype=torch.int64, device=q_blk.device)
            got = index.search_deep(q_blk.index_select(0, sel), k)[0]
            return {r: got[j].tolist() for j, r in enumerate(rows)}
        preds.update(filter_block(ids.cpu().numpy(), need, blk, ground_truth, train_positive_map, max_k, deeper=deeper))
    return preds
