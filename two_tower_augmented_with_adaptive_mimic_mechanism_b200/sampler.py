"""Device-side uniform negative sampler with rejection of known positives (reference samplers.py:11-85).

Same contract as the reference's per-row Python loop — `num_negatives` uniform draws per user from
[0, num_items), members of `positives[user]` redrawn for at most `max_rounds` rounds — but vectorised
over the batch: the positives are a sorted int64 key array (user * num_items + item) searched with
`torch.searchsorted`.  Statistical (not bit) parity: the draws come from the device generator.
"""
from __future__ import annotations

from typing import Mapping

import torch


class PositiveSet:
    """Sorted (user, item) keys resident on the device."""

    def __init__(self, positives: Mapping[int, set[int]] | None, num_items: int, device) -> None:
        self.num_items = int(num_items)
        if positives:
            keys = [int(u) * self.num_items + int(i) for u, items in positives.items() for i in items]
            self.keys = torch.tensor(sorted(keys), dtype=torch.int64, device=device)
        else:
            self.keys = torch.empty(0, dtype=torch.int64, device=device)

    def contains(self, users: torch.Tensor, items: torch.Tensor) -> torch.Tensor:
        if self.keys.numel() == 0:
            return torch.zeros_like(items, dtype=torch.bool)
        q = users * self.num_items + items
        pos = torch.searchsorted(self.keys, q).clamp_max(self.keys.numel() - 1)
        return self.keys[pos] == q


_cache: dict = {}
_draws = 0   # Philox counter base: advances with every call so that successive batches draw fresh numbers


def sample_negative_items(users: torch.Tensor, *, num_items: int, positives, num_negatives: int, device,
                          max_rounds: int = 10, generator=None) -> torch.Tensor:
    if num_negatives <= 0:
        raise ValueError("num_negatives must be greater than zero.")
    if num_items <= 1:
        raise ValueError("num_items must be greater than one.")
    pset = positives if isinstance(positives, PositiveSet) else None
    if pset is None:
        # single-entry cache keyed by the identity of the caller's dict; the dict itself is kept alive in the entry
        # so that its id() cannot be recycled by another object while the entry exists
        key = (id(positives), int(num_items), str(device))
        entry = _cache.get(key)
        if entry is None or entry[0] is not positives:
            entry = (positives, PositiveSet(positives, num_items, device))
            _cache.clear()
            _cache[key] = entry
        pset = entry[1]
    users = users.to(device).contiguous()
    B = users.shape[0]
    if users.is_cuda:
        # one launch: every (row, slot) draws and re-draws on its own (csrc/sampler.cu)
        from . import functional as F
        global _draws
        neg = torch.empty((B, num_negatives), dtype=torch.int64, device=users.device)
        fail = torch.zeros(1, dtype=torch.int32, device=users.device)
        seed = int(generator.initial_seed()) if generator is not None else int(torch.initial_seed())
        F.check(F.lib().ttam_sample_negatives(users.data_ptr(), B, num_negatives, int(num_items), F._ptr(pset.keys if pset.keys.numel() else None),
                                              pset.keys.numel(), int(max_rounds), seed & 0xFFFFFFFFFFFFFFFF, _draws, None,
                                              neg.data_ptr(), fail.data_ptr(), F._stream()), "sample_negatives")
        _draws += B * num_negatives * (max_rounds + 1)
        if pset.keys.numel() and int(fail.item()):
            raise RuntimeError("Exceeded resampling attempts while drawing negatives.")
        return neg
    neg = torch.randint(0, num_items, (B, num_negatives), device=device, generator=generator)
    if pset.keys.numel() == 0:
        return neg
    u2 = users.view(-1, 1).expand(B, num_negatives)
    bad = pset.contains(u2, neg)
    attempts = 0
    while bool(bad.any()):               # host-side index logic only (CPU tensors: unit tests of the contract)
        redraw = torch.randint(0, num_items, (B, num_negatives), device=device, generator=generator)
        neg = torch.where(bad, redraw, neg)
        bad = pset.contains(u2, neg)
        attempts += 1
        if attempts > max_rounds:
            raise RuntimeError("Exceeded resampling attempts while drawing negatives.")
    return neg
