"""Retrieval / evaluation parity: corpus encoding, eval loss, exact top-K (canonical order), the batched
`_evaluate_model` replacement with the reference's host filter, and ranking metrics."""
import numpy as np
import pytest
import torch

import oracle
from helpers import TRAIN_CASES, build_model, load_case, state_after

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tt():
    import two_tower_augmented_with_adaptive_mimic_mechanism_b200 as pkg
    return pkg


@pytest.mark.parametrize("Q,N,D,K", [(1, 64, 16, 10), (300, 5000, 96, 100), (17, 4097, 128, 100), (5, 50, 8, 100),
                                     (64, 20000, 96, 128)])
def test_topk_f32_bit_exact_ids(tt, Q, N, D, K):
    rng = np.random.default_rng(N)
    q = rng.standard_normal((Q, D)).astype(np.float32)
    items = rng.standard_normal((N, D)).astype(np.float32)
    items[N // 2] = items[N // 3]                      # exact duplicates -> ties broken by id
    items[N - 1] = items[0]
    ids, scores = tt.functional.topk(torch.from_numpy(q).cuda(), torch.from_numpy(items).cuda(), K, id_offset=1000)
    ref_s = oracle.canonical_scores(q, items)
    ref_i, ref_v = oracle.topk_canonical(ref_s, K)
    assert np.array_equal(ids.cpu().numpy(), ref_i + 1000)
    assert np.array_equal(scores.cpu().numpy(), ref_v)          # canonical sequential-fp32 scores, bit for bit


def test_topk_all_ties(tt):
    q = torch.ones(3, 8, device="cuda")
    items = torch.ones(500, 8, device="cuda")
    ids, _ = tt.functional.topk(q, items, 20)
    assert ids.cpu().tolist() == [list(range(20))] * 3


def test_topk_merge_equals_unsharded(tt):
    rng = np.random.default_rng(1)
    Q, N, D, K, parts = 50, 8000, 32, 100, 8
    q = rng.standard_normal((Q, D)).astype(np.float32)
    items = np.round(rng.standard_normal((N, D)) * 4).astype(np.float32) / 4      # coarse values -> many exact ties
    dq, di = torch.from_numpy(q).cuda(), torch.from_numpy(items).cuda()
    full_i, full_s = tt.functional.topk(dq, di, K)
    shard = N // parts
    pi, ps = [], []
    for r in range(parts):
        i_, s_ = tt.functional.topk(dq, di[r * shard:(r + 1) * shard].contiguous(), K, id_offset=r * shard)
        pi.append(i_); ps.append(s_)
    mi, ms = tt.functional.topk_merge(torch.stack(pi, 1), torch.stack(ps, 1), K)
    assert torch.equal(mi, full_i) and torch.equal(ms, full_s)
    ref_i, _ = oracle.topk_canonical(oracle.canonical_scores(q, items), K)
    assert np.array_equal(mi.cpu().numpy(), ref_i)


@pytest.mark.parametrize("name,cosine", [("train_gated_mlp", False), ("train_gated_mlp_cosine", True)])
def test_eval_path_matches_reference_golden(tt, name, cosine, monkeypatch):
    monkeypatch.setattr(tt.hooks.OPTIONS, "precision", "fp32")     # the goldens are fp32; the hooks' default engine is tf32
    d, meta, _ = load_case(name)
    state = state_after(d, meta["steps"] - 1)
    model = build_model(meta, TRAIN_CASES[name], state, "cuda")
    if not cosine:
        model.similarity = torch.nn.Identity()      # anything that is not CosineSimilarity == dot product
    model.eval()
    eng = tt.FusedEngine(model, max_steps=4)
    ux, ix = torch.from_numpy(d["user_x"]).cuda(), torch.from_numpy(d["item_x"]).cuda()
    # _compute_loss
    ev = eng.eval_loss(*(torch.from_numpy(d[f"eval/{k}"]).cuda() for k in ("users", "pos", "neg")), ux, ix)
    assert float(ev[1]) == pytest.approx(float(d["eval/loss"]), rel=5e-6)
    # _encode_item_embeddings (through the hook, CPU result like the reference)
    emb = tt.hooks._encode_item_embeddings(model, num_items=meta["NI"], item_features=ix, device=torch.device("cuda"), batch_size=17)
    assert emb.device.type == "cpu"
    np.testing.assert_allclose(emb.numpy(), d["eval/item_embeddings"], rtol=2e-5, atol=2e-6)
    uemb = eng.encode("user", torch.arange(meta["NU"], device="cuda"), ux)
    np.testing.assert_allclose(uemb.cpu().numpy(), d["eval/user_embeddings"], rtol=2e-5, atol=2e-6)
    # _score_all_items_for_user
    for u in range(4):
        ids = tt.hooks._score_all_items_for_user(model, user_idx=u, top_k=10, num_items=meta["NI"], user_features=ux,
                                                 item_features=ix, device=torch.device("cuda"))
        assert ids == d["eval/score_all_topk"][u].tolist()
    # _evaluate_model + metrics
    import pandas as pd
    keys, ptr, vals = d["eval/pos_keys"], d["eval/pos_ptr"], d["eval/pos_vals"]
    train_pos = {int(k): set(vals[ptr[i]:ptr[i + 1]].tolist()) for i, k in enumerate(keys)}
    val = pd.DataFrame({"user_idx": d["eval/val_users"], "item_idx": d["eval/val_items"]})
    k_values = [5, 10, 20]
    preds, gts = tt.hooks._evaluate_model(model, train_positive_map=train_pos, val_interactions=val, item_feature_tensor=ix,
                                          user_feature_tensor=ux, device=torch.device("cuda"), num_items=meta["NI"],
                                          candidate_samples=50, k_values=k_values, rng=np.random.default_rng(0),
                                          faiss_resources={"normalize": cosine}, faiss_search_k=4 * max(k_values))
    assert sorted(preds) == d["eval/pred_users"].tolist()
    for r, u in enumerate(d["eval/pred_users"].tolist()):
        row = d["eval/pred_items"][r]
        assert preds[u] == row[row >= 0].tolist(), f"user {u}"
    met = oracle.ranking_metrics(preds, gts, k_values)
    for r, k in enumerate(k_values):
        got = [met["recall"][k], met["precision"][k], met["ndcg"][k], met["hit_rate"][k], met["map"][k]]
        np.testing.assert_allclose(got, d["eval/metrics"][r], rtol=1e-12)


def test_device_sampler_excludes_positives(tt):
    """reference tests/test_samplers.py:6-19 semantics, on the device."""
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200.sampler import sample_negative_items
    users = torch.tensor([0, 1, 0, 2], device="cuda")
    positives = {0: {0, 1, 2}, 1: {9}, 2: set()}
    neg = sample_negative_items(users, num_items=10, positives=positives, num_negatives=6, device=torch.device("cuda"))
    assert neg.shape == (4, 6) and neg.dtype == torch.long
    for r, u in enumerate(users.cpu().tolist()):
        assert not (set(neg[r].cpu().tolist()) & positives[u])
    assert int(neg.min()) >= 0 and int(neg.max()) < 10


def test_device_sampler_kernel_statistics_and_failure(tt):
    """csrc/sampler.cu: draws are uniform over the non-positive items, fresh on every call, and a user whose
    positives cover every item raises like the reference (samplers.py:77-80)."""
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200.sampler import sample_negative_items
    dev_ = torch.device("cuda")
    NI = 50
    positives = {u: set(range(0, NI, 10)) for u in range(8)}           # items 0, 10, .. 40 are positives of users 0..7
    users = torch.arange(8, device="cuda").repeat(4096)
    neg = sample_negative_items(users, num_items=NI, positives=positives, num_negatives=5, device=dev_)
    assert neg.shape == (8 * 4096, 5) and neg.dtype == torch.long
    assert int((neg % 10 == 0).sum()) == 0 and int(neg.min()) >= 1 and int(neg.max()) < NI
    allowed = torch.tensor([i for i in range(NI) if i % 10], device="cuda")
    counts = torch.bincount(neg.reshape(-1), minlength=NI)[allowed].double()
    expect = neg.numel() / allowed.numel()
    assert float(((counts - expect) ** 2 / expect).sum()) < 95.0         # chi-square, 44 dof: p ~ 1e-5 at 95
    neg2 = sample_negative_items(users, num_items=NI, positives=positives, num_negatives=5, device=dev_)
    assert not torch.equal(neg, neg2)
    with pytest.raises(RuntimeError, match="Exceeded resampling attempts"):
        sample_negative_items(torch.zeros(64, dtype=torch.long, device="cuda"), num_items=3, positives={0: {0, 1, 2}},
                              num_negatives=4, device=dev_)
    with pytest.raises(tt.TtamError, match="num_negatives must be greater than zero"):
        tt.functional.check(tt.lib().ttam_sample_negatives(users.data_ptr(), 8, 0, NI, None, 0, 10, 0, 0, None,
                                                          neg.data_ptr(), neg.data_ptr(), 0), "sample_negatives")


# ---------------------------------------------------------------------------------------------
# tcgen05 path (bf16 operands): ids AND canonical scores bit-exact against the oracle on the same bf16 values
# ---------------------------------------------------------------------------------------------
def _bf16_case(Q, N, D, seed):
    g = torch.Generator().manual_seed(seed)
    q = (torch.randn((Q, D), generator=g) * 0.3).bfloat16()
    items = (torch.randn((N, D), generator=g) * 0.3).bfloat16()
    return q, items


@pytest.mark.parametrize("Q,N,D,K", [(300, 5000, 96, 100), (1, 300, 96, 100), (257, 70001, 96, 100), (64, 4096, 64, 10),
                                     (130, 9000, 128, 128), (40, 3000, 32, 50), (33, 2000, 16, 100), (512, 200000, 96, 100)])
def test_topk_bf16_tensor_core_bit_exact(tt, Q, N, D, K):
    q, items = _bf16_case(Q, N, D, N + D)
    items[N // 2] = items[N // 3]                     # exact duplicate rows -> exact score ties, broken by id
    items[N - 1] = items[0]
    ids, scores = tt.functional.topk(q.cuda(), items.cuda(), K, id_offset=7)
    ref_s = oracle.canonical_scores(q.float().numpy(), items.float().numpy())
    ref_i, ref_v = oracle.topk_canonical(ref_s, K)
    assert np.array_equal(ids.cpu().numpy(), ref_i + 7)
    assert np.array_equal(scores.cpu().numpy(), ref_v)


def test_topk_bf16_degenerate_inputs_take_the_exact_fallback(tt):
    """All scores tie: the candidate set cannot be proven complete from tensor-core scores, so every query is
    re-done by the brute-force kernel; the answer is still canonical (ids 0..K-1)."""
    q = torch.ones(5, 96).bfloat16().cuda()
    items = torch.ones(3000, 96).bfloat16().cuda()
    ids, scores = tt.functional.topk(q, items, 20)
    assert ids.cpu().tolist() == [list(range(20))] * 5
    assert torch.all(scores == 96.0)
    # two distinct score levels, many ties inside each
    items2 = torch.ones(3000, 96)
    items2[1000:1010] = 2.0
    ids2, _ = tt.functional.topk(q, items2.bfloat16().cuda(), 20)
    assert ids2.cpu().tolist() == [list(range(1000, 1010)) + list(range(10))] * 5


def test_topk_bf16_full_corpus_agrees_with_fp32_path(tt):
    """BASELINE config 3 corpus size (2M x 96 bf16): the tensor-core result equals the fp32 SIMT kernel's
    (canonical arithmetic on the same bf16 values) for a sample of the queries."""
    g = torch.Generator(device="cuda").manual_seed(3)
    items = (torch.randn((2_000_000, 96), device="cuda", generator=g) * 0.3).bfloat16()
    q = (torch.randn((4096, 96), device="cuda", generator=g) * 0.3).bfloat16()
    ids, scores = tt.functional.topk(q, items, 100)
    sel = torch.arange(0, 4096, 32, device="cuda")
    ref_i, ref_s = tt.functional.topk(q[sel].float(), items.float(), 100)
    assert torch.equal(ids[sel], ref_i) and torch.equal(scores[sel], ref_s)
    assert bool((scores[:, :-1] >= scores[:, 1:]).all())           # sortedness over all 4096 queries
    assert int(ids.min()) >= 0 and int(ids.max()) < 2_000_000


def test_flat_ip_index_bf16(tt):
    torch.manual_seed(1234)
    emb = torch.randn(5000, 96, device="cuda")
    q = torch.randn(10, 96, device="cuda")
    idx = tt.retrieval.FlatIPIndex(emb, dtype=torch.bfloat16)
    ids, sc = idx.search(q, 10)
    ref = oracle.topk_canonical(oracle.canonical_scores(q.bfloat16().float().cpu().numpy(), emb.bfloat16().float().cpu().numpy()), 10)[0]
    assert np.array_equal(ids.cpu().numpy(), ref)


# ---------------------------------------------------------------------------------------------
# D up to 256 on the tcgen05 path (BASELINE config 4 is 256-dim): the box-ring variant of the kernel
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("Q,N,D,K", [(300, 5000, 256, 100), (130, 30001, 256, 100), (130, 9000, 192, 128), (33, 2000, 144, 50),
                                     (140, 60000, 256, 100)])
def test_topk_bf16_tensor_core_wide_rows(tt, Q, N, D, K):
    q, items = _bf16_case(Q, N, D, N + D)
    items[N // 2] = items[N // 3]
    items[N - 1] = items[0]
    ids, scores = tt.functional.topk(q.cuda(), items.cuda(), K, id_offset=7)
    ref_s = oracle.canonical_scores(q.float().numpy(), items.float().numpy())
    ref_i, ref_v = oracle.topk_canonical(ref_s, K)
    assert np.array_equal(ids.cpu().numpy(), ref_i + 7)
    assert np.array_equal(scores.cpu().numpy(), ref_v)


def test_topk_bf16_box_ring_equals_tile_ring(tt, monkeypatch):
    """The same D = 96 search through both shapes of the item ring (TTAM_TOPK_RING forces the box ring)."""
    q, items = _bf16_case(700, 150000, 96, 5)
    a_i, a_s = tt.functional.topk(q.cuda(), items.cuda(), 100)
    monkeypatch.setenv("TTAM_TOPK_RING", "1")
    b_i, b_s = tt.functional.topk(q.cuda(), items.cuda(), 100)
    assert torch.equal(a_i, b_i) and torch.equal(a_s, b_s)


# ---------------------------------------------------------------------------------------------
# fp32 index with the candidate pass on the tensor cores (3-way bf16 split): bit-exact fp32 ids and scores
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("Q,N,D,K", [(300, 5000, 96, 100), (1, 300, 96, 100), (257, 70001, 96, 100), (64, 4096, 64, 10),
                                     (130, 9000, 128, 128), (40, 3000, 40, 50), (33, 2000, 8, 100), (100, 30000, 256, 100),
                                     (260, 100000, 96, 100)])
def test_topk_f32_tensor_core_candidates_bit_exact(tt, Q, N, D, K):
    rng = np.random.default_rng(N + D)
    q = rng.standard_normal((Q, D)).astype(np.float32)
    items = rng.standard_normal((N, D)).astype(np.float32)
    items[N // 2] = items[N // 3]                      # exact duplicates -> ties broken by id
    items[N - 1] = items[0]
    dq, di = torch.from_numpy(q).cuda(), torch.from_numpy(items).cuda()
    split = tt.functional.split_bf16x3(di, item_layout=True)
    ids, scores = tt.functional.topk_f32_tc(dq, di, split, K, id_offset=1000)
    ref_i, ref_v = oracle.topk_canonical(oracle.canonical_scores(q, items), K)
    assert np.array_equal(ids.cpu().numpy(), ref_i + 1000)
    assert np.array_equal(scores.cpu().numpy(), ref_v)


def test_topk_f32_tensor_core_normalised_rows_and_near_ties(tt):
    """Cosine-style corpus (unit rows, scores packed into [-1, 1]) with clusters of near-duplicates whose scores differ
    only below bf16 resolution: the split keeps them apart or the re-score / exact fallback does."""
    rng = np.random.default_rng(11)
    Q, N, D, K = 200, 60000, 96, 100
    items = rng.standard_normal((N, D)).astype(np.float32)
    items[1000:1200] = items[999] + 1e-4 * rng.standard_normal((200, D)).astype(np.float32)
    items /= np.linalg.norm(items, axis=1, keepdims=True)
    q = rng.standard_normal((Q, D)).astype(np.float32)
    q[:20] = items[999] + 0.05 * rng.standard_normal((20, D)).astype(np.float32)      # queries that land in the cluster
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    idx = tt.retrieval.FlatIPIndex(torch.from_numpy(items).cuda())
    ids, scores = idx.search(torch.from_numpy(q).cuda(), K)
    assert idx._items_split is not None                # the tensor-core path ran
    ref_i, ref_v = oracle.topk_canonical(oracle.canonical_scores(q, items), K)
    assert np.array_equal(ids.cpu().numpy(), ref_i)
    assert np.array_equal(scores.cpu().numpy(), ref_v)
    simt = tt.retrieval.FlatIPIndex(torch.from_numpy(items).cuda(), tensor_cores=False)
    s_i, s_s = simt.search(torch.from_numpy(q).cuda(), K)
    assert torch.equal(s_i, ids) and torch.equal(s_s, scores)


def test_topk_f32_deep_and_paged(tt):
    """K = 1024 in one call (pending-list select kernel) and a 2500-deep ranking walked page by page."""
    rng = np.random.default_rng(5)
    Q, N, D = 9, 30000, 32
    q = rng.standard_normal((Q, D)).astype(np.float32)
    items = np.round(rng.standard_normal((N, D)) * 4).astype(np.float32) / 4          # coarse values: exact ties
    ref_i, ref_v = oracle.topk_canonical(oracle.canonical_scores(q, items), 2500)
    idx = tt.retrieval.FlatIPIndex(torch.from_numpy(items).cuda())
    ids, scores = idx.search(torch.from_numpy(q).cuda(), 1024)
    assert np.array_equal(ids.cpu().numpy(), ref_i[:, :1024]) and np.array_equal(scores.cpu().numpy(), ref_v[:, :1024])
    d_i, d_s = idx.search_deep(torch.from_numpy(q).cuda(), 2500)
    assert np.array_equal(d_i, ref_i) and np.array_equal(d_s, ref_v)


def test_evaluate_users_heavy_users_match_the_per_user_filter(tt):
    """Users whose blocked set is far larger than one launch returns (search_k = max_k + |gt| + |blocked|,
    training.py:956-958), next to ordinary ones: predictions equal the reference's filter applied to the full ranking."""
    rng = np.random.default_rng(21)
    NU, NI, D, max_k = 40, 6000, 32, 20
    items = rng.standard_normal((NI, D)).astype(np.float32)
    users = rng.standard_normal((NU, D)).astype(np.float32)
    full = oracle.topk_canonical(oracle.canonical_scores(users, items), NI)[0]
    gt, blocked = {}, {}
    for u in range(NU):
        gt[u] = set(rng.choice(NI, size=3, replace=False).tolist())
        if u % 7 == 0:      # heavy: the top of the ranking is blocked
            blocked[u] = set(full[u, :1500].tolist()) - gt[u]
        elif u % 7 == 1:    # top-128 nearly all blocked: falls short of max_k after the tensor-core pass
            blocked[u] = set(full[u, :120].tolist()) - gt[u]
        else:
            blocked[u] = set(rng.choice(NI, size=int(rng.integers(0, 60)), replace=False).tolist()) - gt[u]
    idx = tt.retrieval.FlatIPIndex(torch.from_numpy(items).cuda())
    preds = tt.retrieval.evaluate_users(idx, torch.from_numpy(users).cuda(), list(range(NU)), gt, blocked, [5, max_k])
    for u in range(NU):
        need = max(max_k + len(gt[u]), 1) + len(blocked[u])
        want = tt.retrieval.filter_candidates(full[u, :need].tolist(), blocked[u], gt[u], max_k)
        assert preds[u] == want, u
