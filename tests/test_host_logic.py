"""Host-side logic that needs no GPU: candidate filtering, sampler, hyper-parameter recovery from the reference's
optimiser objects."""
import numpy as np
import pytest
import torch

import oracle
from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import engine, hooks, retrieval, sampler


def test_filter_candidates_matches_oracle_semantics():
    blocked, gt = {3, 4}, {9, 11}
    cand = [4, 7, 7, -1, 3, 9, 1, 2, 5, 6, 8]
    got = retrieval.filter_candidates(cand, blocked, gt, max_k=5)
    # drop blocked/dup/negative, stop at max_k+|gt| = 7 kept, append unseen gt, truncate to 5
    assert got == [7, 9, 1, 2, 5]
    assert retrieval.filter_candidates([], set(), {2}, 3) == [2]


def test_sampler_contract_on_cpu():
    users = torch.tensor([0, 1, 0])
    pos = {0: {0, 1}, 1: {5}}        # P(a slot is still blocked after the 10 re-draws) = (1/3)^11: not a flaky test
    neg = sampler.sample_negative_items(users, num_items=6, positives=pos, num_negatives=2, device=torch.device("cpu"))
    assert neg.shape == (3, 2)
    for r, u in enumerate(users.tolist()):
        assert not (set(neg[r].tolist()) & pos[u])
    with pytest.raises(ValueError, match="num_negatives"):
        sampler.sample_negative_items(users, num_items=6, positives=pos, num_negatives=0, device=torch.device("cpu"))
    with pytest.raises(ValueError, match="num_items"):
        sampler.sample_negative_items(users, num_items=1, positives=pos, num_negatives=1, device=torch.device("cpu"))
    with pytest.raises(RuntimeError, match="resampling"):
        sampler.sample_negative_items(torch.tensor([0]), num_items=2, positives={0: {0, 1}}, num_negatives=1, device=torch.device("cpu"))


def test_hyperparameters_are_recovered_from_reference_optimizers():
    w = [torch.nn.Parameter(torch.zeros(2, 2))]
    e = [torch.nn.Parameter(torch.zeros(4, 2))]
    hp = hooks._hyper_from_optimizers([torch.optim.AdamW(w, lr=3e-3, weight_decay=0.02),
                                       torch.optim.SparseAdam(e, lr=3e-3, betas=(0.8, 0.99))])
    assert hp["optimizer"] == "adamw" and hp["lr"] == 3e-3 and hp["weight_decay"] == 0.02
    assert hp["sparse_betas"] == (0.8, 0.99) and hp["dense_betas"] == (0.9, 0.999)
    hp = hooks._hyper_from_optimizers([torch.optim.SGD(w, lr=0.1, momentum=0.9, weight_decay=0.0)])
    assert hp["optimizer"] == "sgd" and hp["momentum"] == 0.9
    with pytest.raises(ValueError):
        hooks._hyper_from_optimizers([torch.optim.RMSprop(w)])
