"""torchrun -N ranks (NCCL, one GPU each): the row-sharded step on the real kernels, every route
(dynamic / static / peer), against the golden fixtures produced by the unmodified reference
(tests/golden/train_gated_mlp.npz, train_embedding_only.npz).

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
from helpers import TRAIN_CASES, build_model, load_case, state_after  # noqa: E402

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
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(1 if bad else 0)


if __name__ == "__main__":
    main()
