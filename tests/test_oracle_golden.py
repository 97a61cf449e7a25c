"""Pin the numpy oracle to fixtures produced by running the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest

import oracle
from helpers import GOLDEN, TRAIN_CASES, load_case, state_after

# fp32; the reference golden was produced single-threaded, BLAS summation order differs from numpy's
RTOL, ATOL = 2e-5, 2e-7


@pytest.mark.parametrize("name", sorted(TRAIN_CASES))
def test_train_steps_match_reference(name):
    d, meta, state = load_case(name)
    kw = TRAIN_CASES[name]
    spec = oracle.spec_from_state(state, fusion_user=kw.get("fusion"), fusion_item=kw.get("fusion"),
                                  sparse=kw.get("sparse", True), activation=kw.get("activation", "relu"))
    opt = oracle.OptState()
    for s in range(meta["steps"]):
        out = oracle.train_step(state, opt, spec, d[f"step{s}/users"], d[f"step{s}/pos"], d[f"step{s}/neg"],
                                d["user_x"], d["item_x"], lr=meta["lr"], weight_decay=meta["wd"],
                                betas=meta["betas"], optimizer=kw["optimizer"], momentum=meta["momentum"],
                                lambdas=meta["lambdas"], cat_tensor=d["cat_tensor"], major=int(d["major"]))
        assert out["loss"] == pytest.approx(float(d["losses"][s]), rel=2e-6, abs=1e-7)
        if s in (0, meta["steps"] - 1):
            ref = state_after(d, s)
            assert set(ref) == set(state)
            for k in ref:
                # Adam normalises the gradient: elements whose gradient is ~0 may flip sign -> atol of 2*lr on a few
                np.testing.assert_allclose(state[k], ref[k], rtol=RTOL, atol=2e-6, err_msg=f"{name} step {s} {k}")
        # bit-exact: the rows SparseAdam touches
        for tname, rows in out["touched"].items():
            side = "users" if tname.startswith("user") else None
            expect = np.unique(d[f"step{s}/users"]) if side else np.unique(np.concatenate([d[f"step{s}/pos"], d[f"step{s}/neg"].reshape(-1)]))
            assert np.array_equal(rows, expect)
    for k, v in d.items():
        if k.startswith("opt/"):
            _, pname, slot = k.split("/")
            np.testing.assert_allclose(opt.slots[pname][slot], v, rtol=1e-4, atol=1e-9, err_msg=k)


def test_optimizers_match_torch():
    z = np.load(GOLDEN / "optim.npz")
    p0 = z["p0"]
    steps = 6
    p = p0.copy()
    st = {"step": 0}
    for s in range(steps):
        rows = oracle.sparse_adam_step(p, st, z[f"idx{s}"], z[f"val{s}"], lr=1e-3)
        assert np.array_equal(rows, np.unique(z[f"idx{s}"]))
        np.testing.assert_allclose(p, z[f"sparse_adam/p{s}"], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(st["exp_avg"], z["sparse_adam/exp_avg"], rtol=1e-6, atol=1e-10)
    np.testing.assert_allclose(st["exp_avg_sq"], z["sparse_adam/exp_avg_sq"], rtol=1e-6, atol=1e-12)
    for kind in ("adamw", "adam", "sgd"):
        p = p0.copy()
        st = {"step": 0}
        for s in range(steps):
            g = np.zeros_like(p)
            np.add.at(g, z[f"idx{s}"], z[f"val{s}"])
            oracle.dense_step(kind, p, g, st, lr=1e-3, weight_decay=0.01, momentum=0.9)
            np.testing.assert_allclose(p, z[f"{kind}/p{s}"], rtol=1e-6, atol=1e-8, err_msg=f"{kind} {s}")


@pytest.mark.parametrize("name,cosine", [("train_gated_mlp", False), ("train_gated_mlp_cosine", True)])
def test_eval_path_matches_reference(name, cosine):
    d, meta, _ = load_case(name)
    state = {k: v.copy() for k, v in state_after(d, meta["steps"] - 1).items()}
    spec = oracle.spec_from_state(state)
    # _compute_loss
    ev = oracle.eval_loss(state, spec, d["eval/users"], d["eval/pos"], d["eval/neg"], d["user_x"], d["item_x"])
    assert ev == pytest.approx(float(d["eval/loss"]), rel=2e-6)
    # _encode_item_embeddings
    items = oracle.encode_items(state, spec, np.arange(meta["NI"]), d["item_x"])
    np.testing.assert_allclose(items, d["eval/item_embeddings"], rtol=RTOL, atol=ATOL)
    users = oracle.encode_users(state, spec, np.arange(meta["NU"]), d["user_x"])
    np.testing.assert_allclose(users, d["eval/user_embeddings"], rtol=RTOL, atol=ATOL)
    # _score_all_items_for_user: same id SET (torch.topk tie order is arbitrary); no exact ties here, so same order
    for u in range(4):
        ids = oracle.score_all_items_topk(d["eval/user_embeddings"][u], d["eval/item_embeddings"], 10, cosine=cosine)
        assert ids.tolist() == d["eval/score_all_topk"][u].tolist()
    # _evaluate_model (FAISS branch) + compute_ranking_metrics
    keys, ptr, vals = d["eval/pos_keys"], d["eval/pos_ptr"], d["eval/pos_vals"]
    train_pos = {int(k): set(vals[ptr[i]:ptr[i + 1]].tolist()) for i, k in enumerate(keys)}
    k_values = [5, 10, 20]
    preds, gts = oracle.evaluate_flat_ip(lambda u: d["eval/user_embeddings"][u], d["eval/item_embeddings"],
                                         d["eval/val_users"], d["eval/val_items"], train_pos, k_values,
                                         search_k=4 * max(k_values), cosine=cosine)
    assert sorted(preds) == d["eval/pred_users"].tolist()
    for r, u in enumerate(d["eval/pred_users"].tolist()):
        row = d["eval/pred_items"][r]
        assert preds[u] == row[row >= 0].tolist(), f"user {u}"
    met = oracle.ranking_metrics(preds, gts, k_values)
    ref = d["eval/metrics"]
    for r, k in enumerate(k_values):
        got = [met["recall"][k], met["precision"][k], met["ndcg"][k], met["hit_rate"][k], met["map"][k]]
        np.testing.assert_allclose(got, ref[r], rtol=1e-12)
    assert met["mrr"] == pytest.approx(ref[-1][0], rel=1e-12)


def test_metric_known_answers():
    """The reference's own known-answer test (reference tests/test_metrics.py:4-19)."""
    preds = {0: [3, 2, 1], 1: [4, 5, 6]}
    gts = {0: {1, 2}, 1: {4}}
    ks = [1, 2, 3]
    m = oracle.ranking_metrics(preds, gts, ks)
    z = np.load(GOLDEN / "metrics.npz")
    for key, zk in (("recall", "recall"), ("precision", "precision"), ("ndcg", "ndcg"), ("hit_rate", "hit"), ("map", "map")):
        np.testing.assert_allclose([m[key][k] for k in ks], z[zk])
    assert m["recall"][1] == 0.5 and m["precision"][1] == 0.5 and m["hit_rate"][1] == 0.5
    assert m["recall"][3] > m["recall"][1]
    assert abs(m["mrr"] - 0.75) < 1e-6


def test_canonical_topk_ties_break_by_id():
    s = np.array([[1.0, 3.0, 3.0, 2.0, 3.0]], dtype=np.float32)
    ids, sc = oracle.topk_canonical(s, 4)
    assert ids.tolist() == [[1, 2, 4, 3]]


@pytest.mark.parametrize("name", ["train_gated_mlp", "train_gated_mlp_cosine"])
def test_torch_port_matches_reference_golden(name):
    """The torch CPU port timed by bench.py as the CPU baseline reproduces the reference's numbers."""
    import torch
    from oracle import torch_port
    torch.set_num_threads(1)
    d, meta, init = load_case(name)
    model = torch_port.Model(meta["NU"], meta["NI"], meta["D"], meta["F"], meta["H"], meta["Hg"] or meta["D"])
    model.load_reference_state(init)
    opts = torch_port.build_optimizers(model, lr=meta["lr"], weight_decay=meta["wd"], betas=meta["betas"])
    ux, ix = torch.from_numpy(d["user_x"]), torch.from_numpy(d["item_x"])
    for s in range(meta["steps"]):
        loss = torch_port.train_step(model, opts, torch.from_numpy(d[f"step{s}/users"]), torch.from_numpy(d[f"step{s}/pos"]),
                                     torch.from_numpy(d[f"step{s}/neg"]), ux, ix, lambdas=meta["lambdas"][:2])
        assert loss == pytest.approx(float(d["losses"][s]), rel=1e-6)
    got, ref = model.reference_state(), state_after(d, meta["steps"] - 1)
    assert set(got) == set(ref)
    for k in ref:
        np.testing.assert_allclose(got[k], ref[k], rtol=1e-5, atol=1e-7, err_msg=k)
    np.testing.assert_allclose(torch_port.encode_items(model, ix).numpy(), d["eval/item_embeddings"], rtol=1e-5, atol=1e-7)
