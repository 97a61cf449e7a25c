"""End-to-end parity of the fused training step against (1) the golden fixtures produced by the unmodified
reference and (2) the numpy oracle on larger seeded inputs.

Tolerances (fp32 everywhere; only the summation order differs from the reference's CPU BLAS):
  losses            rel 5e-6
  updated tensors   rtol 5e-5, atol 2e-6   (Adam normalises the gradient, so elements whose gradient is ~0
                                            can move by up to ~lr with either sign: atol = 2e-3 * lr)
  SparseAdam touched-row sets: bit-exact.
"""
import numpy as np
import pytest
import torch

import oracle
from helpers import TRAIN_CASES, build_model, load_case, model_state_np, state_after, synthetic_gated as _synthetic

pytestmark = pytest.mark.gpu
RTOL, ATOL = 5e-5, 2e-6


def _engine(model, meta, kw, **extra):
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import FusedEngine
    lu, li, lc = meta["lambdas"]
    return FusedEngine(model, optimizer=kw["optimizer"], lr=meta["lr"], weight_decay=meta["wd"], momentum=meta["momentum"],
                       sparse_betas=meta["betas"], loss_weights={"mimic_user": lu, "mimic_item": li, "category_alignment": lc},
                       max_steps=64, **extra)


@pytest.mark.parametrize("name", sorted(TRAIN_CASES))
def test_fused_step_matches_reference_golden(name):
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import functional as F
    d, meta, init = load_case(name)
    kw = TRAIN_CASES[name]
    model = build_model(meta, kw, init, "cuda")
    eng = _engine(model, meta, kw, item_category_tensor=torch.from_numpy(d["cat_tensor"]).cuda(), major_category_id=int(d["major"]))
    ux, ix = torch.from_numpy(d["user_x"]).cuda(), torch.from_numpy(d["item_x"]).cuda()
    for s in range(meta["steps"]):
        u, p, n = (torch.from_numpy(d[f"step{s}/{k}"]).cuda() for k in ("users", "pos", "neg"))
        loss = eng.train_step(u, p, n, ux, ix)
        assert float(loss[0]) == pytest.approx(float(d["losses"][s]), rel=5e-6, abs=1e-7)
        # bit-exact: the row sets the sparse optimiser touches
        su, _ = F.sort_rows(u, meta["NU"])
        si, _ = F.sort_rows(torch.cat([p, n.reshape(-1)]), meta["NI"])
        assert np.array_equal(F.unique_rows(su).cpu().numpy(), np.unique(d[f"step{s}/users"]))
        assert np.array_equal(F.unique_rows(si).cpu().numpy(), np.unique(np.concatenate([d[f"step{s}/pos"], d[f"step{s}/neg"].reshape(-1)])))
        if s in (0, meta["steps"] - 1):
            eng.flush()
            got, ref = model_state_np(model), state_after(d, s)
            assert set(got) == set(ref)
            for k in ref:
                np.testing.assert_allclose(got[k], ref[k], rtol=RTOL, atol=ATOL, err_msg=f"{name} step {s} {k}")
    st = eng.optimizer_state()
    for k, v in d.items():
        if k.startswith("opt/"):
            _, pname, slot = k.split("/")
            key = "exp_avg" if slot == "momentum_buffer" else slot
            np.testing.assert_allclose(st[pname][key].cpu().numpy(), v, rtol=2e-4, atol=1e-7, err_msg=k)


def test_train_one_epoch_hook_replays_reference_golden(monkeypatch):
    """hooks._train_one_epoch with the reference's call signature (training.py:700-716): a DataLoader of (users, pos)
    batches, the torch optimisers `_run_single_experiment` would build (used as configuration carriers), the reference's
    sampler replaced by a replay of the golden negatives.  Result: the golden state, the sample-weighted mean loss the
    reference returns (training.py:833), and the moments published into `optimizer.state` for `_save_checkpoint`."""
    import two_tower_augmented_with_adaptive_mimic_mechanism_b200 as tt
    from torch import nn
    name = "train_gated_mlp"
    d, meta, init = load_case(name)
    model = build_model(meta, TRAIN_CASES[name], init, "cuda")
    sparse = [model.user_encoder.embedding.weight, model.item_encoder.embedding.weight]
    dense = [p for p in model.parameters() if all(p is not q for q in sparse)]
    opts = [torch.optim.AdamW(dense, lr=meta["lr"], weight_decay=meta["wd"]),
            torch.optim.SparseAdam(sparse, lr=meta["lr"], betas=meta["betas"])]
    steps = meta["steps"]
    batches = [(torch.from_numpy(d[f"step{s}/users"]), torch.from_numpy(d[f"step{s}/pos"])) for s in range(steps)]
    negs = iter([torch.from_numpy(d[f"step{s}/neg"]).cuda() for s in range(steps)])
    monkeypatch.setattr(tt.hooks.OPTIONS, "sampler", "reference")       # the caller's DataLoader order + its sampler:
    monkeypatch.setattr(tt.hooks.OPTIONS, "reference_sampler", lambda users, **kw: next(negs))   # here a replay
    monkeypatch.setattr(tt.hooks.OPTIONS, "precision", "fp32")
    monkeypatch.setattr(tt.hooks.OPTIONS, "graph", False)
    lu, li, _ = meta["lambdas"]
    ux, ix = torch.from_numpy(d["user_x"]).cuda(), torch.from_numpy(d["item_x"]).cuda()
    mean_loss = tt.hooks._train_one_epoch(model, batches, optimizers=opts, criterion=nn.BCEWithLogitsLoss(),
                                          negatives_per_positive=meta["N"], num_items=meta["NI"], user_positive_items={},
                                          user_features=ux, item_features=ix, device=torch.device("cuda"),
                                          gradient_clip_norm=None, loss_weights={"mimic_user": lu, "mimic_item": li},
                                          item_category_tensor=None, major_category_id=None)
    sizes = np.array([b[0].shape[0] for b in batches], np.float64)
    assert mean_loss == pytest.approx(float(np.dot(d["losses"][:steps], sizes) / sizes.sum()), rel=5e-6)
    got, ref = model_state_np(model), state_after(d, steps - 1)
    for k in ref:
        np.testing.assert_allclose(got[k], ref[k], rtol=RTOL, atol=ATOL, err_msg=k)
    st = opts[1].state[sparse[0]]
    np.testing.assert_allclose(st["exp_avg"].cpu().numpy(), d["opt/user_encoder.embedding.weight/exp_avg"], rtol=2e-4, atol=1e-7)
    assert model.training
    with pytest.raises(NotImplementedError):
        tt.hooks._train_one_epoch(model, [], optimizers=opts, criterion=nn.BCEWithLogitsLoss(), negatives_per_positive=5,
                                  num_items=meta["NI"], user_positive_items={}, user_features=ux, item_features=ix,
                                  device=torch.device("cuda"), gradient_clip_norm=1.0)


@pytest.mark.parametrize("D,H,Hg,F,B,graph", [(96, 192, 96, 605, 512, False), (128, 256, 128, 64, 256, False),
                                              (96, 192, 96, 608, 512, True), (256, 512, 256, 605, 384, False)])
def test_fused_step_matches_oracle_at_tower_shapes(D, H, Hg, F, B, graph):
    """North-star tower shapes (96-dim, 192->96 MLP, F=605), duplicate-heavy Zipf positives, 3 steps, and the same
    steps replayed from a captured CUDA graph."""
    NU, NI, N = 3000, 5000, 5
    st, user_x, item_x, batches = _synthetic(7, NU, NI, D, H, Hg, F, B, N)
    meta = dict(NU=NU, NI=NI, D=D, H=H, Hg=Hg, F=F, lr=1e-3, wd=0.01, momentum=0.0, betas=(0.9, 0.999), lambdas=(0.15, 0.15, 0.0))
    kw = dict(optimizer="adamw")
    model = build_model(meta, kw, st, "cuda")
    eng = _engine(model, meta, kw)
    ref_state = {k: v.copy() for k, v in st.items()}
    spec = oracle.spec_from_state(ref_state)
    opt = oracle.OptState()
    ux, ix = torch.from_numpy(user_x).cuda(), torch.from_numpy(item_x).cuda()
    for u, p, n in batches:
        ref = oracle.train_step(ref_state, opt, spec, u, p, n, user_x, item_x, lr=1e-3, weight_decay=0.01, lambdas=(0.15, 0.15, 0.0))
        loss = eng.train_step(torch.from_numpy(u).cuda(), torch.from_numpy(p).cuda(), torch.from_numpy(n).cuda(), ux, ix, graph=graph)
        got = loss.cpu().numpy()
        assert got[0] == pytest.approx(ref["loss"], rel=5e-6)
        assert got[1] == pytest.approx(ref["bce"], rel=5e-6)
        assert got[2] == pytest.approx(ref["mimic_user"], rel=5e-6)
        assert got[3] == pytest.approx(ref["mimic_item"], rel=5e-6)
    eng.flush()
    got = model_state_np(model)
    for k in ref_state:
        # Adam normalises every update to ~lr: where a gradient is ~0 its sign is decided by the summation order.  At most one
        # element in 100 000 may miss the bar, and then by less than lr per step.
        bad = np.abs(got[k] - ref_state[k]) > ATOL + RTOL * np.abs(ref_state[k])
        assert bad.sum() <= max(0, int(1e-5 * bad.size)), (k, int(bad.sum()))
        assert np.abs(got[k] - ref_state[k]).max() <= 1e-3 * len(batches), k
    # rows never touched keep their bits in the sparse tables; in the aug tables they decay (AdamW on every row)
    untouched = np.setdiff1d(np.arange(NU), np.unique(np.concatenate([b[0] for b in batches])))
    assert np.array_equal(got["user_encoder.embedding.weight"][untouched], st["user_encoder.embedding.weight"][untouched])
    a0, a1 = st["adaptive_mimic.user_augmented.weight"][untouched], got["adaptive_mimic.user_augmented.weight"][untouched]
    np.testing.assert_allclose(a1, a0 * np.float32(1 - 1e-3 * 0.01) ** 3, rtol=1e-6)


@pytest.mark.parametrize("graph,D,H,Hg", [(False, 96, 192, 96), (True, 96, 192, 96), (False, 256, 512, 256), (False, 128, 256, 128)])
def test_fused_step_tf32_tensor_cores_within_tolerance(graph, D, H, Hg):
    """The same three steps with every tower GEMM on the tcgen05 TF32 path.  Stated tolerance: losses rel 2e-3; updated
    parameters: mean |diff| <= 2e-5, at most 0.1 % of the elements off by more than 3e-4 per step, none by more than
    2.5 lr per step (Adam normalises every update to ~lr = 1e-3, so where a gradient is ~0 a rounding-level
    disagreement about it moves the weight by up to ~lr: the update is ill-conditioned there, in any precision);
    touched-row index sets bit-exact."""
    NU, NI, N, F, B = 3000, 5000, 5, 605, 512
    st, user_x, item_x, batches = _synthetic(7, NU, NI, D, H, Hg, F, B, N)
    meta = dict(NU=NU, NI=NI, D=D, H=H, Hg=Hg, F=F, lr=1e-3, wd=0.01, momentum=0.0, betas=(0.9, 0.999), lambdas=(0.15, 0.15, 0.0))
    kw = dict(optimizer="adamw")
    model = build_model(meta, kw, st, "cuda")
    eng = _engine(model, meta, kw, precision="tf32")
    ref_state = {k: v.copy() for k, v in st.items()}
    spec, opt = oracle.spec_from_state(ref_state), oracle.OptState()
    ux, ix = torch.from_numpy(user_x).cuda(), torch.from_numpy(item_x).cuda()
    for u, p, n in batches:
        ref = oracle.train_step(ref_state, opt, spec, u, p, n, user_x, item_x, lr=1e-3, weight_decay=0.01, lambdas=(0.15, 0.15, 0.0))
        loss = eng.train_step(torch.from_numpy(u).cuda(), torch.from_numpy(p).cuda(), torch.from_numpy(n).cuda(), ux, ix, graph=graph)
        got = loss.cpu().numpy()
        assert got[0] == pytest.approx(ref["loss"], rel=2e-3)
        assert got[1] == pytest.approx(ref["bce"], rel=2e-3)
    eng.flush()
    got = model_state_np(model)
    for k in ref_state:
        d = np.abs(got[k] - ref_state[k])
        assert d.max() <= 2.5e-3 * len(batches), (k, d.max())
        assert (d > 3e-4 * len(batches)).sum() <= max(2, 1e-3 * d.size), (k, (d > 3e-4 * len(batches)).sum(), d.size)
        assert d.mean() <= 2e-5, (k, d.mean())
    touched = np.unique(np.concatenate([b[0] for b in batches]))
    changed = np.nonzero((got["user_encoder.embedding.weight"] != st["user_encoder.embedding.weight"]).any(1))[0]
    assert np.array_equal(changed, touched)


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_fused_step_concat_gelu_golden(precision):
    """Concat fusion + projection with a GELU feature MLP (the non-composite launch sequence) against the golden the
    unmodified reference produced; TF32 variant within the stated TF32 tolerance."""
    name = "train_concat_gelu"
    d, meta, init = load_case(name)
    kw = TRAIN_CASES[name]
    model = build_model(meta, kw, init, "cuda")
    eng = _engine(model, meta, kw, precision=precision)
    ux, ix = torch.from_numpy(d["user_x"]).cuda(), torch.from_numpy(d["item_x"]).cuda()
    for s in range(meta["steps"]):
        u, p, n = (torch.from_numpy(d[f"step{s}/{k}"]).cuda() for k in ("users", "pos", "neg"))
        loss = eng.train_step(u, p, n, ux, ix)
        assert float(loss[0]) == pytest.approx(float(d["losses"][s]), rel=5e-6 if precision == "fp32" else 2e-3)
    eng.flush()
    got, ref = model_state_np(model), state_after(d, meta["steps"] - 1)
    assert set(got) == set(ref)
    for k in ref:
        if precision == "fp32":
            np.testing.assert_allclose(got[k], ref[k], rtol=RTOL, atol=ATOL, err_msg=k)
        else:
            assert np.abs(got[k] - ref[k]).mean() <= 2e-5, k


def test_dropout_training_step_runs_and_differs_per_step():
    NU, NI, D, H, Hg, F, B, N = 500, 800, 32, 64, 32, 24, 128, 3
    st, user_x, item_x, batches = _synthetic(3, NU, NI, D, H, Hg, F, B, N)
    import two_tower_augmented_with_adaptive_mimic_mechanism_b200 as tt
    cfg = {"type": "tower", "id_embedding": {"params": {"embedding_dim": D, "sparse": True}},
           "feature_encoder": {"type": "mlp", "hidden_dims": [H], "output_dim": D, "dropout": 0.15}, "fusion": "gated",
           "adaptive_mimic": {"hidden_dim": Hg}}
    ue = tt.build_tower_encoder(cfg, num_embeddings=NU, feature_dim=F)
    ie = tt.build_tower_encoder(cfg, num_embeddings=NI, feature_dim=F)
    model = tt.TwoTowerModel(ue, ie, adaptive_mimic=tt.AdaptiveMimicMechanism(num_users=NU, num_items=NI, embedding_dim=D)).cuda()
    eng = tt.FusedEngine(model, lr=1e-3, weight_decay=0.01, loss_weights={"mimic_user": 0.15, "mimic_item": 0.15}, max_steps=16)
    ux, ix = torch.from_numpy(user_x).cuda(), torch.from_numpy(item_x).cuda()
    u, p, n = (torch.from_numpy(a).cuda() for a in batches[0])
    l1 = eng.train_step(u, p, n, ux, ix).clone()
    h1 = eng.bufs_u["hd0"][:B].clone()
    l2 = eng.train_step(u, p, n, ux, ix).clone()
    h2 = eng.bufs_u["hd0"][:B].clone()
    assert torch.isfinite(l1).all() and torch.isfinite(l2).all()
    z1, z2 = (h1 == 0), (h2 == 0)
    assert 0.1 < float(z1.float().mean()) < 0.9          # relu zeros + ~15% dropped
    assert not torch.equal(z1, z2)                       # a fresh mask every step


def test_module_api_forward_backward_matches_oracle():
    """The nn.Module path (autograd.Function wrappers) that `scripts/train.py` would use without the fused hook."""
    d, meta, init = load_case("train_gated_mlp")
    model = build_model(meta, TRAIN_CASES["train_gated_mlp"], init, "cuda")
    model.train()
    u = torch.from_numpy(d["step0/users"]).cuda()
    p = torch.from_numpy(d["step0/pos"]).cuda()
    ux, ix = torch.from_numpy(d["user_x"]).cuda(), torch.from_numpy(d["item_x"]).cuda()
    t_u = model.user_encoder({"indices": u, "features": ux.index_select(0, u)})
    t_p = model.item_encoder({"indices": p, "features": ix.index_select(0, p)})
    o_u, o_p, lu, li = model.adaptive_mimic(user_indices=u, item_indices=p, user_embedding=t_u, item_embedding=t_p)
    spec = oracle.spec_from_state(init)
    cu = oracle.tower_forward(init, "user", spec.user, d["step0/users"], d["user_x"][d["step0/users"]])
    cp = oracle.tower_forward(init, "item", spec.item, d["step0/pos"], d["item_x"][d["step0/pos"]])
    np.testing.assert_allclose(t_u.detach().cpu().numpy(), cu["t"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(t_p.detach().cpu().numpy(), cp["t"], rtol=2e-5, atol=2e-6)
    qu = init["adaptive_mimic.user_augmented.weight"][d["step0/users"]]
    assert float(lu) == pytest.approx(float(((qu - cp["t"]) ** 2).mean()), rel=1e-5)
    loss = (o_u * o_p).sum(-1).mean() + 0.15 * lu + 0.15 * li
    loss.backward()
    g = model.user_encoder.embedding.weight.grad
    assert g.is_sparse and g._nnz() == u.numel()
    assert model.adaptive_mimic.user_augmented.weight.grad.shape == (meta["NU"], meta["D"])
    w = model.user_encoder.feature_encoder.network[0].weight
    assert w.grad is not None and torch.isfinite(w.grad).all() and float(w.grad.abs().sum()) > 0
    # reference shape tests (reference tests/test_encoders.py:6-26, tests/test_adaptive_mimic.py:6-33)
    assert t_u.shape == (u.numel(), meta["D"]) and o_p.shape == t_p.shape and lu.ndim == 0 and float(li) >= 0
    n = torch.from_numpy(d["step0/neg"]).cuda()
    t_n = model.item_encoder({"indices": n.reshape(-1), "features": ix.index_select(0, n.reshape(-1))})
    assert model.adaptive_mimic.augment_items(n.reshape(-1), t_n).shape == t_n.shape


@pytest.mark.parametrize("name,graph", [("train_gated_mlp", False), ("train_gated_mlp", True), ("train_embedding_only", True)])
def test_static_slot_route_matches_reference_golden(name, graph):
    """The static-shape sharded step (sharding.SlotExchange: fixed slots, padding = repeated real ids with zero gradient
    rows) on the real kernels, world size 1: slot padding is live because the batches are not multiples of 128.  Same
    bars as the plain fused step; with graph=True steps 1.. are CUDA-graph replays (plan + main)."""
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import ShardedEngine
    d, meta, init = load_case(name)
    kw = TRAIN_CASES[name]
    model = build_model(meta, kw, init, "cuda")
    sh = ShardedEngine(_engine(model, meta, kw), static=True)
    ux, ix = torch.from_numpy(d["user_x"]).cuda(), torch.from_numpy(d["item_x"]).cuda()
    for s in range(meta["steps"]):
        u, p, n = (torch.from_numpy(d[f"step{s}/{k}"]).cuda() for k in ("users", "pos", "neg"))
        loss = sh.train_step(u, p, n, ux, ix, graph=graph)
        assert float(loss[0]) == pytest.approx(float(d["losses"][s]), rel=5e-6, abs=1e-7)
        assert sh.last_exchange_rows[0] % 128 == 0 and sh.last_exchange_rows[0] >= u.numel()
    assert sh.fallback_steps == 0
    sh.eng.flush()
    got, ref = model_state_np(model), state_after(d, meta["steps"] - 1)
    for k in ref:
        np.testing.assert_allclose(got[k], ref[k], rtol=RTOL, atol=ATOL, err_msg=f"{name} {k}")


def test_static_slot_route_overflow_falls_back_and_grows():
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import ShardedEngine
    name = "train_gated_mlp"
    d, meta, init = load_case(name)
    kw = TRAIN_CASES[name]
    model = build_model(meta, kw, init, "cuda")
    sh = ShardedEngine(_engine(model, meta, kw), static=True, capacity=(8, 8))
    ux, ix = torch.from_numpy(d["user_x"]).cuda(), torch.from_numpy(d["item_x"]).cuda()
    for s in range(meta["steps"]):
        u, p, n = (torch.from_numpy(d[f"step{s}/{k}"]).cuda() for k in ("users", "pos", "neg"))
        loss = sh.train_step(u, p, n, ux, ix, graph=True)
        assert float(loss[0]) == pytest.approx(float(d["losses"][s]), rel=5e-6, abs=1e-7)
    assert sh.fallback_steps >= 1 and sh.capacity[0] > 8
    sh.eng.flush()
    got, ref = model_state_np(model), state_after(d, meta["steps"] - 1)
    for k in ref:
        np.testing.assert_allclose(got[k], ref[k], rtol=RTOL, atol=ATOL, err_msg=f"{name} {k}")


@pytest.mark.parametrize("graph", [False, True])
def test_resume_from_optimizer_state_continues_bit_identically(graph):
    """FusedEngine.load_optimizer_state: an engine rebuilt from the parameters and moments after step 1 (what a checkpoint
    holds) takes step 2 exactly as the engine that never stopped - bias corrections, lazily-updated rows and the SparseAdam
    moments continue where they were."""
    name = "train_gated_mlp"
    d, meta, init = load_case(name)
    kw = TRAIN_CASES[name]
    ux, ix = torch.from_numpy(d["user_x"]).cuda(), torch.from_numpy(d["item_x"]).cuda()
    batches = [tuple(torch.from_numpy(d[f"step{s}/{k}"]).cuda() for k in ("users", "pos", "neg")) for s in range(meta["steps"])]
    model_a = build_model(meta, kw, init, "cuda")
    eng_a = _engine(model_a, meta, kw)
    for s in range(2):
        eng_a.train_step(*batches[s], ux, ix, graph=graph)
    eng_a.flush()                                                                          # what a checkpoint does first
    saved_params = {k: v.clone() for k, v in model_a.state_dict().items()}
    saved_opt = {k: {kk: (vv.clone() if torch.is_tensor(vv) else vv) for kk, vv in ent.items()} for k, ent in eng_a.optimizer_state().items()}
    loss_a = eng_a.train_step(*batches[2], ux, ix, graph=graph).clone()
    eng_a.flush()
    model_b = build_model(meta, kw, {k: v.cpu().numpy() for k, v in saved_params.items()}, "cuda")
    eng_b = _engine(model_b, meta, kw)
    eng_b.load_optimizer_state(saved_opt, 2)
    loss_b = eng_b.train_step(*batches[2], ux, ix, graph=graph).clone()
    eng_b.flush()
    torch.cuda.synchronize()
    assert torch.equal(loss_a, loss_b)
    sa, sb = model_a.state_dict(), model_b.state_dict()
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    oa, ob = eng_a.optimizer_state(), eng_b.optimizer_state()
    for k in oa:
        for slot in ("exp_avg", "exp_avg_sq"):
            if oa[k][slot] is not None:
                assert torch.equal(oa[k][slot], ob[k][slot]), (k, slot)
    assert float(loss_b[0]) == pytest.approx(float(d["losses"][2]), rel=5e-6, abs=1e-7)
