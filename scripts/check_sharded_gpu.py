"""torchrun -N ranks (NCCL, one GPU each): the row-sharded step on the real kernels, every route
(dynamic / static / peer), against the golden fixtures produced by the unmodified reference
(tests/golden/train_gated_mlp.npz, train_embedding_only.npz), and at tower shapes (D = 96, F = 605, 192 -> 96 MLPs,
B = 128 samples per rank, TF32 tensor-core GEMMs as in the bench) against the numpy oracle's un-sharded step.

Each rank holds rows r % W == rank of the four tables and of the feature matrices, takes B/W samples of every golden
batch, and runs ShardedEngine.train_step; the un-sharded result must match the reference state after the last step
(rtol 5e-5 / atol 2e-6, losses rel 5e-6: the bars of tests/test_gpu_train.py).  Steps 1.. of the static and peer routes
are CUDA-graph replays.  Prints one line per (case, route); exit code 1 on any mismatch."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from helpers import TRAIN_CASES, build_model, load_case, state_after, synthetic_gated  # noqa: E402

TABLES = ("user_encoder.embedding.weight", "item_encoder.embedding.weight",
          "adaptive_mimic.user_augmented.weight", "adaptive_mimic.item_augmented.weight")


def main():
    import two_tower_augmented_with_adaptive_mimic_mechanism_b200 as tt
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import sharding as S
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    bad = 0
    for case in ("train_gated_mlp", "train_embedding_only"):
        for route in ("dynamic", "static", "peer"):
            d, meta, init = load_case(case)
            kw = TRAIN_CASES[case]
            shard = {k: (np.ascontiguousarray(v[rank::world]) if k in TABLES else v) for k, v in init.items()}
            m2 = dict(meta, NU=S.shard_size(meta["NU"], rank, world), NI=S.shard_size(meta["NI"], rank, world))
            model = build_model(m2, kw, shard, dev)
            lu, li, _ = meta["lambdas"]
            eng = tt.FusedEngine(model, optimizer=kw["optimizer"], lr=meta["lr"], weight_decay=meta["wd"], sparse_betas=meta["betas"],
                                 loss_weights={"mimic_user": lu, "mimic_item": li}, max_steps=64)
            # explicit (default-sized) capacities: no calibration steps, so steps 1.. of the 3-step goldens replay the graphs
            Bl, Nn = meta["B"] // world, meta["N"]
            caps = (S.default_slot_capacity(Bl, world), S.default_slot_capacity(Bl * (1 + Nn), world))
            sh = tt.ShardedEngine(eng, static=route != "dynamic", peer=route == "peer", capacity=caps)
            has_x = "user_x" in d and d["user_x"].size
            ux = S.shard_rows(torch.from_numpy(d["user_x"]), rank, world).to(dev) if has_x else None
            ix = S.shard_rows(torch.from_numpy(d["item_x"]), rank, world).to(dev) if has_x else None
            ok = True
            for s in range(meta["steps"]):
                u, p, n = (torch.from_numpy(d[f"step{s}/{k}"]) for k in ("users", "pos", "neg"))
                B = u.shape[0] // world
                sl = slice(rank * B, (rank + 1) * B)
                share = sh.train_step(u[sl].to(dev), p[sl].to(dev), n[sl].to(dev), ux, ix, graph=True)
                loss = float(sh.global_loss(share)[0])
                ref = float(d["losses"][s])
                if (u.shape[0] // world) * world == u.shape[0] and abs(loss - ref) > 5e-6 * abs(ref) + 1e-7:
                    ok = False
                    print(f"[rank {rank}] {case} {route} step {s}: loss {loss} != {ref}", flush=True)
            eng.flush()
            torch.cuda.synchronize()
            ref_state = state_after(d, meta["steps"] - 1)
            worst = 0.0
            for k, v in model.state_dict().items():
                mine = v.detach().cpu().numpy()
                want = ref_state[k][rank::world] if k in TABLES else ref_state[k]
                err = np.abs(mine - want) - (2e-6 + 5e-5 * np.abs(want))
                worst = max(worst, float(err.max()))
                if err.max() > 0:
                    ok = False
                    print(f"[rank {rank}] {case} {route}: {k} off by {float(np.abs(mine - want).max()):.3e}", flush=True)
            if has_x and "eval/item_embeddings" in d and route == "peer":
                # corpus export (reference _encode_item_embeddings, training.py:613-643) and the reference-layout checkpoint
                emb = sh.export_item_embeddings(None, ix, meta["NI"]).cpu().numpy()
                if not np.allclose(emb, d["eval/item_embeddings"], rtol=2e-5, atol=2e-6):
                    ok = False
                    print(f"[rank {rank}] {case}: exported item embeddings differ from the reference's", flush=True)
                full = sh.full_state_dict(meta["NU"], meta["NI"])
                for k, v in ref_state.items():
                    if not np.allclose(full[k].numpy(), v, rtol=5e-5, atol=2e-6):
                        ok = False
                        print(f"[rank {rank}] {case}: gathered state_dict[{k}] differs", flush=True)
            flag = torch.tensor([0 if ok else 1], device=dev)
            dist.all_reduce(flag)
            if rank == 0:
                print(f"{case:24s} {route:8s} world={world} fallback_steps={sh.fallback_steps} "
                      f"{'OK' if int(flag) == 0 else 'MISMATCH'} (worst margin {worst:.2e})", flush=True)
            bad += int(flag)
            del sh, eng, model
    bad += tower_shape_case(tt, S, rank, world, dev)
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(1 if bad else 0)


def tower_shape_case(tt, S, rank, world, dev) -> int:
    """North-star tower shapes on the routes the bench uses (peer when world > 1, static otherwise), precision tf32 and fp32:
    W ranks x 128 samples against the oracle's step on the whole batch."""
    import oracle
    NU, NI, D, H, Hg, F, N, steps = 3000, 5000, 96, 192, 96, 605, 5, 5
    B = 128 * world
    st, user_x, item_x, batches = synthetic_gated(7, NU, NI, D, H, Hg, F, B, N, steps=steps)
    # step 2 asks for ONE item 768 times per rank: its owner's bucket overflows the slots on a step that is replayed from the
    # recorded graphs (steps 0 / 1 record and replay them), i.e. after the speculative forward half has already been launched;
    # the step must come out of the dynamic route unharmed, the slots are re-sized, steps 3 / 4 record and replay again
    hot = batches[2]
    hot_neg = hot[2].copy()
    hot_neg[:, :3] = 17                      # (the other two columns stay random: every rank still owns some of the rows)
    batches[2] = (hot[0], np.full_like(hot[1], 17), hot_neg)
    # explicit capacities (no calibration steps: the graphs are recorded on step 0), tight enough for step 2 to overflow
    Ri = 128 * (1 + N)
    caps = (S.default_slot_capacity(128, world), Ri if world == 1 else (Ri // world + 128 + 127) // 128 * 128)
    ref_state = {k: v.copy() for k, v in st.items()}
    spec, opt = oracle.spec_from_state(ref_state), oracle.OptState()
    ref_losses = [oracle.train_step(ref_state, opt, spec, u, p, n, user_x, item_x, lr=1e-3, weight_decay=0.01,
                                    lambdas=(0.15, 0.15, 0.0))["loss"] for u, p, n in batches]
    bad = 0
    for precision, ltol, mean_tol in (("fp32", 5e-6, 2e-7), ("tf32", 2e-3, 3e-5)):
        meta = dict(NU=S.shard_size(NU, rank, world), NI=S.shard_size(NI, rank, world), D=D, H=H, Hg=Hg, F=F)
        shard = {k: (np.ascontiguousarray(v[rank::world]) if k in TABLES else v) for k, v in st.items()}
        model = build_model(meta, dict(optimizer="adamw"), shard, dev)
        eng = tt.FusedEngine(model, optimizer="adamw", lr=1e-3, weight_decay=0.01, precision=precision,
                             loss_weights={"mimic_user": 0.15, "mimic_item": 0.15}, max_steps=64)
        sh = tt.ShardedEngine(eng, static=True, peer=world > 1, capacity=caps)
        ux = S.shard_rows(torch.from_numpy(user_x), rank, world).to(dev)
        ix = S.shard_rows(torch.from_numpy(item_x), rank, world).to(dev)
        ok = True
        for s, (u, p, n) in enumerate(batches):
            sl = slice(rank * 128, (rank + 1) * 128)
            share = sh.train_step(torch.from_numpy(u[sl]).to(dev), torch.from_numpy(p[sl]).to(dev), torch.from_numpy(n[sl]).to(dev),
                                  ux, ix, graph=True)
            loss = float(sh.global_loss(share)[0])
            if abs(loss - ref_losses[s]) > ltol * abs(ref_losses[s]):
                ok = False
                print(f"[rank {rank}] tower shapes {precision} step {s}: loss {loss} != {ref_losses[s]}", flush=True)
        eng.flush()
        torch.cuda.synchronize()
        worst = 0.0
        for k, v in model.state_dict().items():
            want = ref_state[k][rank::world] if k in TABLES else ref_state[k]
            worst = max(worst, float(np.abs(v.detach().cpu().numpy() - want).mean()))
        if worst > mean_tol:
            ok = False
            print(f"[rank {rank}] tower shapes {precision}: parameters mean |diff| {worst:.3e} > {mean_tol}", flush=True)
        # bit-exact: the rows of the sparse user table this rank changed are exactly the users of the batches it owns
        touched = np.unique(np.concatenate([b[0] for b in batches]))
        mine = touched[touched % world == rank] // world
        changed = np.nonzero((model.state_dict()["user_encoder.embedding.weight"].cpu().numpy() != shard["user_encoder.embedding.weight"]).any(1))[0]
        if not np.array_equal(changed, mine):
            ok = False
            print(f"[rank {rank}] tower shapes {precision}: touched-row set differs", flush=True)
        if world > 1 and sh.fallback_steps != 1:
            ok = False
            print(f"[rank {rank}] tower shapes {precision}: expected exactly one overflow step, saw {sh.fallback_steps}", flush=True)
        flag = torch.tensor([0 if ok else 1], device=dev)
        dist.all_reduce(flag)
        if rank == 0:
            print(f"{'tower_shapes_' + precision:24s} {'peer' if sh.peer else 'static':8s} world={world} fallback_steps={sh.fallback_steps} "
                  f"{'OK' if int(flag) == 0 else 'MISMATCH'} (params mean |diff| {worst:.2e})", flush=True)
        bad += int(flag)
        del sh, eng, model
    return bad


if __name__ == "__main__":
    main()
