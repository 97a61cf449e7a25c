"""N-rank host logic on CPU: sharding arithmetic, the all-to-all row exchange and the sharded training step
(world_size 2, gloo), the item-sharded top-K merge.  The compute inside the step is the numpy oracle
(tests/sharding_backend.py); what is under test is routing, ordering, global-batch normalisation and the
dense-gradient all-reduce of `ShardedEngine` — the code the GPU path runs unchanged over NCCL."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, str(Path(__file__).resolve().parent))
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

import oracle  # noqa: E402
from helpers import load_case  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _spawn(fn, world, *args):
    port = _free_port()
    mp.spawn(_entry, args=(world, port, fn, args), nprocs=world, join=True)


def _entry(rank, world, port, fn, args):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fn(rank, world, *args)
    finally:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
def test_shard_arithmetic_roundtrip():
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import sharding as S
    for n, W in [(10, 3), (7, 8), (64, 4), (1, 2)]:
        full = torch.arange(n * 2).view(n, 2)
        shards = [S.shard_rows(full, r, W) for r in range(W)]
        assert [s.shape[0] for s in shards] == [S.shard_size(n, r, W) for r in range(W)]
        assert torch.equal(S.unshard_rows(shards), full)
        idx = torch.arange(n)
        for r in range(W):
            mine = idx[S.owner_of(idx, W) == r]
            assert torch.equal(full[mine], shards[r][S.local_row(mine, W)])
            assert torch.equal(S.global_row(S.local_row(mine, W), r, W), mine)


def test_exchange_world1_is_identity():
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import sharding as S
    idx = torch.tensor([5, 1, 5, 0])
    ex = S.Exchange(idx, 1)
    rows = torch.randn(4, 3)
    assert torch.equal(ex.local_rows, idx) and ex.to_requester(rows) is rows and ex.to_owner(rows) is rows


def _exchange_worker(rank, world):
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import sharding as S
    g = torch.Generator().manual_seed(100 + rank)
    n_rows = 37
    table = torch.arange(n_rows * 3, dtype=torch.float32).view(n_rows, 3)          # the same "global table" on every rank
    shard = S.shard_rows(table, rank, world)
    for R in (1, 8, 50 + 7 * rank):                                               # ragged: ranks request different counts
        idx = torch.randint(0, n_rows, (R,), generator=g)
        if R == 8:
            idx[:] = rank                                                         # every row to a single owner
        ex = S.Exchange(idx, world)
        assert torch.equal(S.owner_of(ex.recv_idx, world), torch.full_like(ex.recv_idx, rank))
        got = ex.to_requester(shard[ex.local_rows])                               # owner-side gather, routed back
        assert torch.equal(got, table[idx])
        # gradients: what the owner receives, accumulated by local row, equals the global scatter-add restricted to it
        grads = torch.randn(R, 3, generator=g)
        recv = ex.to_owner(grads)
        acc = torch.zeros_like(shard).index_add_(0, ex.local_rows, recv)
        all_idx = [torch.empty(0, dtype=torch.int64)] * world
        all_g = [None] * world
        dist.all_gather_object(all_idx, idx)
        dist.all_gather_object(all_g, grads)
        full = torch.zeros_like(table)
        for i, gr in zip(all_idx, all_g):
            full.index_add_(0, i, gr)
        torch.testing.assert_close(acc, S.shard_rows(full, rank, world), rtol=1e-6, atol=1e-6)
    flat = [torch.full((3,), float(rank + 1)), torch.full((2, 2), float(10 * (rank + 1)))]
    S.all_reduce_flat(flat)
    assert torch.equal(flat[0], torch.full((3,), 3.0)) and torch.equal(flat[1], torch.full((2, 2), 30.0))


def test_exchange_roundtrip_world2():
    _spawn(_exchange_worker, 2)


def _slot_exchange_worker(rank, world):
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import sharding as S
    g = torch.Generator().manual_seed(200 + rank)
    n_rows, R = 41, 64
    table = torch.arange(n_rows * 3, dtype=torch.float32).view(n_rows, 3)
    shard = S.shard_rows(table, rank, world)
    ex = S.SlotExchange(R, 48, world)                                             # 48 slots per (requester, owner) pair
    for trial in range(3):
        idx = torch.randint(0, n_rows, (R,), generator=g)
        ex.plan(idx); ex.agree()
        assert int(ex.flag) == 0
        ex.exchange_ids()
        assert ex.recv_idx.numel() == world * 48
        assert torch.equal(S.owner_of(ex.recv_idx, world), torch.full_like(ex.recv_idx, rank))   # padding included
        # padding repeats real ids: the owner's touched-row set is exactly the union of what the ranks requested from it
        all_idx = [None] * world
        dist.all_gather_object(all_idx, idx)
        want = torch.unique(torch.cat([i[S.owner_of(i, world) == rank] for i in all_idx]))
        assert torch.equal(torch.unique(ex.recv_idx), want)
        got_t, got_q, got_o = ex.pull(shard[ex.local_rows], 2 * shard[ex.local_rows])
        assert torch.equal(got_t, table[idx]) and torch.equal(got_q, 2 * table[idx]) and torch.equal(got_o, 3 * table[idx])
        only_t = ex.pull(shard[ex.local_rows])
        assert torch.equal(only_t[0], table[idx]) and only_t[1] is None and only_t[2] is only_t[0]
        grads = torch.randn(R, 3, generator=g)
        b0 = torch.randn(10, 3, generator=g)
        recv, recv_b = ex.push(grads, b0, grads)                                  # zeros in the padding slots
        want_b = torch.cat([b0, grads[10:]])
        acc_b = torch.zeros_like(shard).index_add_(0, ex.local_rows, recv_b)
        all_b = [None] * world
        dist.all_gather_object(all_b, want_b)
        full_b = torch.zeros_like(table)
        for i, gr in zip(all_idx, all_b):
            full_b.index_add_(0, i, gr)
        torch.testing.assert_close(acc_b, S.shard_rows(full_b, rank, world), rtol=1e-6, atol=1e-6)
        assert ex.push(grads)[1] is None
        acc = torch.zeros_like(shard).index_add_(0, ex.local_rows, recv)
        all_g = [None] * world
        dist.all_gather_object(all_g, grads)
        full = torch.zeros_like(table)
        for i, gr in zip(all_idx, all_g):
            full.index_add_(0, i, gr)
        torch.testing.assert_close(acc, S.shard_rows(full, rank, world), rtol=1e-6, atol=1e-6)
    # overflow on ONE rank only, and an empty bucket: every rank sees the flag after agree()
    idx = torch.randint(0, n_rows, (R,), generator=g)
    if rank == 0:
        idx = idx - idx % world + 1                                               # all 64 ids to owner 1: > 48 slots, bucket 0 empty
    ex.plan(idx)
    assert int(ex.flag) == (1 if rank == 0 else 0)
    ex.agree()
    assert int(ex.flag) == 1
    assert S.default_slot_capacity(49152, 8) % 128 == 0 and S.default_slot_capacity(49152, 8) >= 49152 // 8
    assert S.default_slot_capacity(100, 2) <= 128


def test_slot_exchange_roundtrip_world2():
    _spawn(_slot_exchange_worker, 2)


# ---------------------------------------------------------------------------------------------
def _train_worker(rank, world, case, steps, out_dir, route="dynamic"):
    from sharding_backend import OracleEngine, shard_state
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import sharding as S
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200.sharded import ShardedEngine
    d, meta, init = load_case(case)
    lu, li, _ = meta["lambdas"]
    eng = OracleEngine(shard_state(init, rank, world), lr=meta["lr"], weight_decay=meta["wd"], betas=meta["betas"],
                       lambdas=(lu, li))
    if route == "dynamic":
        sh = ShardedEngine(eng)
    else:      # "static": slots sized by default; "static_tight": 8 slots per pair -> the first step overflows, falls back, grows
        sh = ShardedEngine(eng, static=True, capacity=(8, 8) if route == "static_tight" else None)
    ux = S.shard_rows(torch.from_numpy(d["user_x"]), rank, world) if "user_x" in d and d["user_x"].size else None
    ix = S.shard_rows(torch.from_numpy(d["item_x"]), rank, world) if "item_x" in d and d["item_x"].size else None
    losses = []
    for s in range(steps):
        u, p, n = (torch.from_numpy(d[f"step{s}/{k}"]) for k in ("users", "pos", "neg"))
        B = u.shape[0] // world
        sl = slice(rank * B, (rank + 1) * B)
        share = sh.train_step(u[sl], p[sl], n[sl], ux, ix)
        losses.append(sh.global_loss(share).numpy())
    np.savez(Path(out_dir) / f"rank{rank}.npz", losses=np.stack(losses), **{"st/" + k: v for k, v in eng.state.items()},
             **{"touched/" + k: v for k, v in eng.touched.items()}, fallback=np.array(sh.fallback_steps))


@pytest.mark.parametrize("route", ["dynamic", "static", "static_tight"])
@pytest.mark.parametrize("case", ["train_gated_mlp", "train_embedding_only"])
def test_sharded_step_world2_equals_single_process(case, route, tmp_path):
    """Two ranks, each with half of the batch and half of the rows == the one-process oracle on the whole batch:
    touched-row sets bit-exact (after mapping local rows back to global ids), values within fp32 summation tolerance."""
    from sharding_backend import TABLES, unshard_state
    world, steps = 2, 2
    d, meta, init = load_case(case)
    if (d["step0/users"].shape[0] // world) * world != d["step0/users"].shape[0]:
        pytest.skip("batch not divisible")
    _spawn(_train_worker, world, case, steps, str(tmp_path), route)
    ranks = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    for z in ranks:                                            # the tight slots must have taken the dynamic route at least once
        assert (int(z["fallback"]) > 0) == (route == "static_tight")
    got = unshard_state([{k[3:]: z[k] for k in z.files if k.startswith("st/")} for z in ranks])
    ref_state = {k: v.copy() for k, v in init.items()}
    spec, opt = oracle.spec_from_state(ref_state), oracle.OptState()
    lu, li, _ = meta["lambdas"]
    has_x = "user_x" in d and d["user_x"].size
    for s in range(steps):
        ref = oracle.train_step(ref_state, opt, spec, d[f"step{s}/users"], d[f"step{s}/pos"], d[f"step{s}/neg"],
                                d["user_x"] if has_x else None, d["item_x"] if has_x else None, lr=meta["lr"],
                                weight_decay=meta["wd"], betas=meta["betas"], lambdas=(lu, li, 0.0))
        for z in ranks:                                        # every rank reports the same global loss
            assert z["losses"][s][0] == pytest.approx(ref["loss"], rel=2e-6)
            assert z["losses"][s][1] == pytest.approx(ref["bce"], rel=2e-6)
    for k, v in ref_state.items():
        np.testing.assert_allclose(got[k], v, rtol=2e-5, atol=2e-7, err_msg=k)
    for tab in TABLES[:2]:                                     # SparseAdam touched rows of the last step, as global ids
        glob = np.sort(np.concatenate([ranks[r]["touched/" + tab] * world + r for r in range(world)]))
        assert np.array_equal(glob, ref["touched"][tab])


# ---------------------------------------------------------------------------------------------
def _np_merge(ids, scores, k):
    """(-score, +id) W-way merge in numpy (stands in for ttam_topk_merge on CPU)."""
    q = ids.shape[0]
    ids2, sc2 = ids.reshape(q, -1).numpy(), scores.reshape(q, -1).numpy()
    out_i, out_s = np.empty((q, k), np.int64), np.empty((q, k), np.float32)
    for r in range(q):
        order = np.lexsort((ids2[r], -sc2[r]))[:k]
        out_i[r], out_s[r] = ids2[r][order], sc2[r][order]
    return torch.from_numpy(out_i), torch.from_numpy(out_s)


def _topk_worker(rank, world, out_dir):
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import sharding as S
    rng = np.random.default_rng(3)
    Q, NI, D, K = 8, 101, 16, 10
    q = rng.standard_normal((Q, D)).astype(np.float32)
    items = rng.standard_normal((NI, D)).astype(np.float32)
    items[50] = items[3]; items[77] = items[3]                                  # exact ties across shards
    scores = oracle.canonical_scores(q, items)
    mine = np.arange(rank, NI, world)
    li, ls = oracle.topk_canonical(scores[:, mine], K, ids=mine)                 # local top-K with global ids
    ids, sc = S.merge_topk_shards(torch.from_numpy(li), torch.from_numpy(ls), K, _np_merge)
    np.savez(Path(out_dir) / f"topk{rank}.npz", ids=ids.numpy(), sc=sc.numpy())


def test_item_sharded_topk_merge_world2(tmp_path):
    _spawn(_topk_worker, 2, str(tmp_path))
    rng = np.random.default_rng(3)
    Q, NI, D, K = 8, 101, 16, 10
    q = rng.standard_normal((Q, D)).astype(np.float32)
    items = rng.standard_normal((NI, D)).astype(np.float32)
    items[50] = items[3]; items[77] = items[3]
    ref_i, ref_s = oracle.topk_canonical(oracle.canonical_scores(q, items), K)
    got_i = np.concatenate([np.load(tmp_path / f"topk{r}.npz")["ids"] for r in range(2)])
    got_s = np.concatenate([np.load(tmp_path / f"topk{r}.npz")["sc"] for r in range(2)])
    assert np.array_equal(got_i, ref_i) and np.array_equal(got_s, ref_s)


# ---------------------------------------------------------------------------------------------
class _TinyModel(torch.nn.Module):
    """state_dict keys of the reference model (the row-sharded four + one replicated tensor)."""

    def __init__(self, nu, ni, d):
        super().__init__()
        mk = lambda n, sparse=False: torch.nn.Embedding(n, d, sparse=sparse)
        self.user_encoder = torch.nn.Module(); self.user_encoder.embedding = mk(nu, True)
        self.item_encoder = torch.nn.Module(); self.item_encoder.embedding = mk(ni, True)
        self.item_encoder.dense = torch.nn.Linear(d, d)
        self.adaptive_mimic = torch.nn.Module()
        self.adaptive_mimic.user_augmented = mk(nu); self.adaptive_mimic.item_augmented = mk(ni)


class _TinyEngine:
    def __init__(self, model):
        self.model, self.flushed = model, 0

    def flush(self):
        self.flushed += 1

    def optimizer_state(self):
        return {n: {"step": 3, "exp_avg": p.detach() * 2, "exp_avg_sq": p.detach() ** 2} for n, p in self.model.named_parameters()}

    def load_optimizer_state(self, state, step):
        self.loaded = (state, step)


def _checkpoint_worker(rank, world, out_dir):
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import sharding as S
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200.sharded import ShardedEngine
    NU, NI, D = 11, 14, 3                                                          # ragged: ranks own 6/5 and 7/7 rows
    torch.manual_seed(0)
    full = _TinyModel(NU, NI, D)                                                   # the same "unsharded model" on every rank
    ref = {k: v.clone() for k, v in full.state_dict().items()}
    sh = ShardedEngine(_TinyEngine(_TinyModel(S.shard_size(NU, rank, world), S.shard_size(NI, rank, world), D)))
    sh.load_full_state_dict(ref)                                                   # reference checkpoint -> this rank's rows
    mine = sh.eng.model.state_dict()
    assert torch.equal(mine["user_encoder.embedding.weight"], ref["user_encoder.embedding.weight"][rank::world])
    assert torch.equal(mine["item_encoder.dense.weight"], ref["item_encoder.dense.weight"])
    got = sh.full_state_dict(NU, NI)                                               # ... and back
    assert sh.eng.flushed == 1 and set(got) == set(ref)
    for k in ref:
        assert torch.equal(got[k], ref[k]), k
    path = Path(out_dir) / "ckpt.pt"
    sh.save_checkpoint(path, NU, NI, epoch=4, metric_name="recall@10", metric_value=0.5)
    ck = torch.load(path, weights_only=False)
    assert set(ck) == {"epoch", "model_state_dict", "optimizer_state_dicts", "metric_name", "metric_value", "timestamp"}
    assert ck["epoch"] == 4 and all(torch.equal(ck["model_state_dict"][k], ref[k]) for k in ref)
    # optimizer_state_dicts = [dense optimiser, SparseAdam] in torch's own layout, parameters in the reference's order
    # (training.py:276-309, 1315-1346): the reference's optimisers load them as they are
    dense_p = list(full.item_encoder.dense.parameters()) + [full.adaptive_mimic.user_augmented.weight, full.adaptive_mimic.item_augmented.weight]
    sparse_p = [full.user_encoder.embedding.weight, full.item_encoder.embedding.weight]
    adamw, sadam = torch.optim.AdamW(dense_p), torch.optim.SparseAdam(sparse_p)
    adamw.load_state_dict(ck["optimizer_state_dicts"][0])
    sadam.load_state_dict(ck["optimizer_state_dicts"][1])
    assert torch.equal(adamw.state[full.adaptive_mimic.item_augmented.weight]["exp_avg"], 2 * ref["adaptive_mimic.item_augmented.weight"])
    assert torch.equal(sadam.state[full.user_encoder.embedding.weight]["exp_avg_sq"], ref["user_encoder.embedding.weight"] ** 2)
    assert float(adamw.state[dense_p[0]]["step"]) == 3.0 and sadam.state[sparse_p[1]]["step"] == 3
    # ... and back: tables and moments return to the ranks that own their rows
    with torch.no_grad():
        for prm in sh.eng.model.parameters():
            prm.zero_()
    meta_back = sh.load_checkpoint(path)
    assert meta_back == {"epoch": 4, "metric_name": "recall@10", "metric_value": 0.5}
    assert torch.equal(sh.eng.model.state_dict()["item_encoder.embedding.weight"], ref["item_encoder.embedding.weight"][rank::world])
    state, step = sh.eng.loaded
    assert step == 3 and set(state) == set(ref)
    assert torch.equal(state["adaptive_mimic.user_augmented.weight"]["exp_avg"], 2 * ref["adaptive_mimic.user_augmented.weight"][rank::world])
    assert torch.equal(state["item_encoder.dense.weight"]["exp_avg_sq"], ref["item_encoder.dense.weight"] ** 2)
    with pytest.raises(ValueError):
        S.gather_rows_from_shards(torch.zeros(1, D), NU, None)                     # wrong shard length for this rank


def test_sharded_checkpoint_roundtrip_world2(tmp_path):
    _spawn(_checkpoint_worker, 2, str(tmp_path))


def _ragged_static_worker(rank, world):
    from sharding_backend import OracleEngine, shard_state
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200.sharded import ShardedEngine
    d, meta, init = load_case("train_embedding_only")
    eng = OracleEngine(shard_state(init, rank, world), lr=meta["lr"], weight_decay=meta["wd"], betas=meta["betas"], lambdas=(0.0, 0.0))
    sh = ShardedEngine(eng, static=True)
    u, p, n = (torch.from_numpy(d[f"step0/{k}"]) for k in ("users", "pos", "neg"))
    B = 8 + 2 * rank                                                               # ranks bring different batch sizes
    with pytest.raises(ValueError, match="ranks disagree"):
        sh.train_step(u[:B], p[:B], n[:B], None, None)
    sh2 = ShardedEngine(eng)                                                       # the dynamic route takes ragged batches
    sh2.train_step(u[:B], p[:B], n[:B], None, None)


def test_static_route_rejects_ragged_batches_world2():
    _spawn(_ragged_static_worker, 2)


def _calibration_worker(rank, world):
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import sharding as S
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200.sharded import ShardedEngine
    sh = ShardedEngine(object(), static=True)                                      # automatic capacities
    B, N = 4096, 5
    st = sh._static_state(B, N, "cpu")
    cap0 = (st.ex_u.cap, st.ex_i.cap)
    assert st.calib_left == ShardedEngine.CALIBRATION_STEPS and cap0[1] == S.default_slot_capacity(B * 6, world)
    g = torch.Generator().manual_seed(7 + rank)
    for step in range(ShardedEngine.CALIBRATION_STEPS):
        st.users.copy_(torch.randint(0, 10**6, (B,), generator=g))
        items = torch.randint(0, 10**6, (B * 6,), generator=g)
        if rank == 1:
            items[: B // 2] = 0                                                    # a hot row on owner 0, seen by ONE rank only
        st.items.copy_(items)
        assert (B, N) in sh._static
        sh._calibrate(st)
    # every rank re-sizes alike: the hot bucket (about B*6/2 + B/4 ids) does not fit the uniform starting capacity
    hot = int(torch.bincount(items % world, minlength=world).max()) if rank == 1 else 0
    assert (B, N) not in sh._static and (B, N) in sh._calibrated
    assert sh.capacity[0] == cap0[0] and sh.capacity[1] > cap0[1] and sh.capacity[1] % 128 == 0
    t = torch.tensor([hot]); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert sh.capacity[1] >= int(t) + 5 * int((B * 6 / world) ** 0.5)
    st2 = sh._static_state(B, N, "cpu")                                            # rebuilt with the new capacity, no second calibration
    assert st2.ex_i.cap == sh.capacity[1] and st2.calib_left == 0
    # explicit capacities are respected: no calibration
    sh3 = ShardedEngine(object(), static=True, capacity=(2048 + 128, 12288 + 256))
    assert sh3._static_state(B, N, "cpu").calib_left == 0


def test_slot_capacity_calibration_world2():
    _spawn(_calibration_worker, 2)


def _grid_topk_worker(rank, world, out_dir):
    """Query-group x item-shard grid (W = 4, G = 2): the host logic of ShardedFlatIPIndex(query_groups=2) with the local
    search done by the oracle (FlatIPIndex itself needs CUDA)."""
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import sharding as S
    G = 2
    g, s, n_shards = S.retrieval_grid(rank, world, G)
    assert (g, s, n_shards) == (rank // 2, rank % 2, 2)
    groups = S.retrieval_subgroups(world, G)                                      # collective: every rank builds both groups
    rng = np.random.default_rng(11)
    Q, NI, D, K = 16, 203, 12, 9
    q = rng.standard_normal((Q, D)).astype(np.float32)
    items = rng.standard_normal((NI, D)).astype(np.float32)
    items[120] = items[7]; items[121] = items[7]                                  # exact ties across the two shards
    per = Q // G
    mine_q = q[g * per:(g + 1) * per]                                             # this rank's query group
    mine_i = np.arange(s, NI, n_shards)                                           # training layout with S in the place of W
    li, ls = oracle.topk_canonical(oracle.canonical_scores(mine_q, items[mine_i]), K)          # local rows of the shard
    li = np.where(li >= 0, li * n_shards + s, li)                                 # -> global ids
    ids, sc = S.merge_topk_shards(torch.from_numpy(li), torch.from_numpy(ls), K, _np_merge, groups[g])
    assert ids.shape == (Q // world, K)
    np.savez(Path(out_dir) / f"grid{rank}.npz", ids=ids.numpy(), sc=sc.numpy())
    with pytest.raises(ValueError):
        S.retrieval_grid(rank, world, 3)


def test_query_group_item_shard_grid_world4(tmp_path):
    _spawn(_grid_topk_worker, 4, str(tmp_path))
    rng = np.random.default_rng(11)
    Q, NI, D, K = 16, 203, 12, 9
    q = rng.standard_normal((Q, D)).astype(np.float32)
    items = rng.standard_normal((NI, D)).astype(np.float32)
    items[120] = items[7]; items[121] = items[7]
    ref_i, ref_s = oracle.topk_canonical(oracle.canonical_scores(q, items), K)
    got_i = np.concatenate([np.load(tmp_path / f"grid{r}.npz")["ids"] for r in range(4)])       # rank r holds query block r
    got_s = np.concatenate([np.load(tmp_path / f"grid{r}.npz")["sc"] for r in range(4)])
    assert np.array_equal(got_i, ref_i) and np.array_equal(got_s, ref_s)


class _CpuFlatIP:
    """numpy stand-in for retrieval.FlatIPIndex (which needs CUDA): same constructor / search contract, oracle arithmetic."""

    def __init__(self, item_embeddings, *, normalize=False, dtype=None, id_offset=0):
        self.items, self.id_offset = item_embeddings.numpy(), int(id_offset)

    def search(self, queries, k):
        ids, sc = oracle.topk_canonical(oracle.canonical_scores(queries.numpy(), self.items), k)
        return torch.from_numpy(np.where(ids >= 0, ids + self.id_offset, ids)), torch.from_numpy(sc)


def _sharded_index_worker(rank, world, out_dir):
    """ShardedFlatIPIndex itself (the class bench.py drives at N > 1), both item layouts, G = 1 and G = 2, on gloo with the
    CUDA pieces (FlatIPIndex, ttam_topk_merge) replaced by their oracle equivalents."""
    import two_tower_augmented_with_adaptive_mimic_mechanism_b200.functional as F
    import two_tower_augmented_with_adaptive_mimic_mechanism_b200.retrieval as R
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200.sharded import ShardedFlatIPIndex
    R.FlatIPIndex = _CpuFlatIP
    F.topk_merge = _np_merge
    rng = np.random.default_rng(23)
    Q, NI, D, K = 16, 157, 10, 7
    q = torch.from_numpy(rng.standard_normal((Q, D)).astype(np.float32))
    items = rng.standard_normal((NI, D)).astype(np.float32)
    items[90] = items[5]                                                          # a tie across shards
    items_t = torch.from_numpy(items)
    out = {}
    for G in (1, 2):
        S_ = world // G
        s = rank % S_
        # training layout: shard s = rows s, s+S, ...
        idx = ShardedFlatIPIndex(items_t[s::S_].contiguous(), query_groups=G)
        out[f"stride{G}"] = idx.search(q, K)[0].numpy()
        # contiguous blocks
        per = (NI + S_ - 1) // S_
        lo, hi = s * per, min(NI, (s + 1) * per)
        idx = ShardedFlatIPIndex(items_t[lo:hi].contiguous(), contiguous_offset=lo, query_groups=G)
        out[f"block{G}"] = idx.search(q, K)[0].numpy()
    with pytest.raises(ValueError):
        ShardedFlatIPIndex(items_t, query_groups=3)
    np.savez(Path(out_dir) / f"index{rank}.npz", **out)


def test_sharded_flat_ip_index_world4(tmp_path):
    _spawn(_sharded_index_worker, 4, str(tmp_path))
    rng = np.random.default_rng(23)
    Q, NI, D, K = 16, 157, 10, 7
    q = rng.standard_normal((Q, D)).astype(np.float32)
    items = rng.standard_normal((NI, D)).astype(np.float32)
    items[90] = items[5]
    ref_i, _ = oracle.topk_canonical(oracle.canonical_scores(q, items), K)
    for key in ("stride1", "block1", "stride2", "block2"):
        got = np.concatenate([np.load(tmp_path / f"index{r}.npz")[key] for r in range(4)])      # rank r: query block r
        assert np.array_equal(got, ref_i), key
