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


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours): one JSON line with the contract's keys, a
    cpu_baseline describing the run and an e2e block that repeats the value with zero copy bytes.  1/16-size tables here."""
    import json
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    out = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--small", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=str(root))
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["metric"] == "train samples/sec" and line["unit"] == "samples/s"
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_filter_block_equals_per_user_filter():
    """The block-wise candidate filter returns, for every user, exactly what the reference's per-user loop
    (`filter_candidates`, training.py:959-972) returns: blocked ids, -1 padding, per-user search_k truncation, rows that fall
    short of max_k (ground truth appended), repeated ids."""
    import numpy as np
    rng = np.random.default_rng(0)
    for trial in range(30):
        n, K, NI = int(rng.integers(1, 40)), int(rng.integers(1, 60)), int(rng.integers(5, 400))
        max_k = int(rng.integers(1, 25))
        users = rng.permutation(1000)[:n].tolist()
        ids = np.stack([rng.permutation(max(NI, K))[:K] for _ in range(n)]).astype(np.int64)
        ids[ids >= NI] = -1                                                     # padding where the corpus is smaller than K
        if trial % 5 == 0:
            ids[:, K // 2:] = ids[:, : K - K // 2]                              # repeated ids
        need = rng.integers(1, K + 1, size=n).tolist()
        blocked = {u: set(rng.integers(0, NI, size=int(rng.integers(0, NI))).tolist()) for u in users if rng.random() < 0.7}
        gt = {u: set(rng.integers(0, NI, size=int(rng.integers(1, 4))).tolist()) for u in users}
        got = retrieval.filter_block(ids, need, users, gt, blocked, max_k)
        for r, u in enumerate(users):
            want = retrieval.filter_candidates(ids[r, : need[r]].tolist(), set(blocked.get(u, ())), gt[u], max_k)
            assert got[u] == want, (trial, r)
    assert retrieval.filter_block(np.zeros((0, 4), np.int64), [], [], {}, {}, 3) == {}


def test_filter_block_fetches_longer_lists_in_one_call():
    """Rows whose K columns keep fewer than max_k candidates although the reference would have searched deeper
    (need > K) are handed to `deeper` together, once; every row ends up with the reference's list."""
    import numpy as np
    rng = np.random.default_rng(3)
    n, K, NI, max_k = 12, 16, 300, 10
    full = np.stack([rng.permutation(NI) for _ in range(n)]).astype(np.int64)
    users = list(range(100, 100 + n))
    gt = {u: set(rng.integers(0, NI, size=2).tolist()) for u in users}
    blocked, need = {}, []
    for r, u in enumerate(users):
        b = set(full[r, : (14 if r % 3 == 0 else 2)].tolist()) - gt[u]          # every third row: its top 14 are blocked
        blocked[u] = b
        need.append(max_k + len(gt[u]) + len(b))
    calls = []

    def deeper(rows, k):
        calls.append((list(rows), k))
        return {r: full[r, :k].tolist() for r in rows}
    got = retrieval.filter_block(full[:, :K], need, users, gt, blocked, max_k, deeper=deeper)
    assert len(calls) == 1 and calls[0][0] == [r for r in range(n) if r % 3 == 0]
    for r, u in enumerate(users):
        assert got[u] == retrieval.filter_candidates(full[r, : need[r]].tolist(), blocked[u], gt[u], max_k), r


def test_evaluate_users_asks_a_tensor_core_index_for_128_first():
    """evaluate_users against a stand-in index (no GPU): an fp32 index with tensor-core candidates is asked for at most 128
    results per user, the users whose blocked items leave fewer than max_k of those get longer lists through ONE
    search_deep call per block, and every prediction equals the reference's per-user filter over the full ranking."""
    import numpy as np
    import torch
    rng = np.random.default_rng(9)
    NU, NI, max_k = 30, 900, 10
    full = np.stack([rng.permutation(NI) for _ in range(NU)]).astype(np.int64)

    class Index:
        tensor_cores, ntotal, max_k = True, NI, 1024
        calls = []

        def search(self, q, k):
            self.calls.append(("search", int(q.shape[0]), int(k)))
            rows = q[:, 0].long().numpy()
            return torch.from_numpy(full[rows, :k]), None

        def search_deep(self, q, k):
            self.calls.append(("deep", int(q.shape[0]), int(k)))
            rows = q[:, 0].long().numpy()
            return full[rows, :k], None
    gt = {u: set(rng.integers(0, NI, size=2).tolist()) for u in range(NU)}
    blocked = {}
    for u in range(NU):
        n_top = 125 if u % 5 == 0 else int(rng.integers(0, 40))           # every fifth user: nearly all of the top 128 blocked
        blocked[u] = set(full[u, :n_top].tolist()) - gt[u]
    idx = Index()
    q = torch.arange(NU, dtype=torch.float32).view(-1, 1)                 # query r "is" user r
    preds = retrieval.evaluate_users(idx, q, list(range(NU)), gt, blocked, [5, max_k], query_block=16)
    assert [c for c in idx.calls if c[0] == "search"] == [("search", 16, 128), ("search", 14, 128)]
    assert len([c for c in idx.calls if c[0] == "deep"]) == 2             # one batched call per block
    for u in range(NU):
        need = max(max_k + len(gt[u]), 1) + len(blocked[u])
        assert preds[u] == retrieval.filter_candidates(full[u, :need].tolist(), blocked[u], gt[u], max_k), u


def test_bag_matrix_layout_roundtrip():
    """functional.BagMatrix (CSR over the sparse columns + dense tail) reproduces the dense matrix; a matrix with a row of more
    than 64 sparse non-zeros is refused (the engine keeps the dense GEMM path for it)."""
    import torch
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200.functional import BagMatrix
    rng = np.random.default_rng(0)
    N, F = 300, 37
    x = np.zeros((N, F), dtype=np.float32)
    for r in range(N):
        x[r, rng.choice(F - 5, size=rng.integers(0, 6), replace=False)] = rng.choice([1.0, 0.5], size=1)
    x[:, F - 5:] = rng.standard_normal((N, 5)).astype(np.float32)
    x[7, F - 2] = 0.0                                     # a zero inside the dense tail stays a stored value
    bag = BagMatrix.build(torch.from_numpy(x), chunk_rows=64)
    assert bag is not None and bag.tail_start == F - 5 and bag.T == 5
    dense = np.zeros_like(x)
    ptr, ent = bag.rowptr.numpy(), bag.entries.numpy()
    for r in range(N):
        seg = ent[ptr[r]:ptr[r + 1]]
        assert np.all(np.diff(seg[:, 0]) > 0)             # column order inside a row
        dense[r, seg[:, 0]] = seg[:, 1].view(np.float32)
    dense[:, bag.tail_start:] = bag.tail.numpy()
    assert np.array_equal(dense, x)
    assert BagMatrix.build(torch.ones((4, 100))) is None  # every column dense: 92 sparse non-zeros per row
    few = torch.zeros((5, 12)); few[:, 3] = 1.0
    b2 = BagMatrix.build(few)
    assert b2.T == 0 and b2.tail is None and b2.rowptr.tolist() == [0, 1, 2, 3, 4, 5]
