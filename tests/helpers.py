"""Shared helpers for the parity tests (CPU side; no CUDA here)."""
from __future__ import annotations

from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"

TRAIN_CASES = {
    # name: kwargs for oracle.spec_from_state + train_step
    "train_gated_mlp": dict(optimizer="adamw"),
    "train_gated_mlp_cal": dict(optimizer="adamw"),
    "train_gated_mlp_cosine": dict(optimizer="adamw"),
    "train_embedding_only": dict(optimizer="adamw"),
    "train_dense_adam": dict(optimizer="adam", sparse=False),
    "train_sgd": dict(optimizer="sgd"),
    "train_linear_sum": dict(optimizer="adamw", fusion="sum"),
    "train_concat_gelu": dict(optimizer="adamw", fusion="concat", activation="gelu"),
}


def load_case(name):
    z = np.load(GOLDEN / f"{name}.npz")
    d = {k: z[k] for k in z.files}
    NU, NI, D, H, Hg, F, B, N, steps = (int(v) for v in d["meta_dims"])
    lr, wd, b1, b2, lu, li, lc, mom = (float(v) for v in d["hyper"])
    meta = dict(NU=NU, NI=NI, D=D, H=H, Hg=Hg, F=F, B=B, N=N, steps=steps, lr=lr, wd=wd,
                betas=(b1, b2), lambdas=(lu, li, lc), momentum=mom)
    init = {k[len("init/"):]: v.copy() for k, v in d.items() if k.startswith("init/")}
    return d, meta, init


def state_after(d, step):
    pre = f"after{step}/"
    return {k[len(pre):]: v for k, v in d.items() if k.startswith(pre)}


def synthetic_gated(seed, NU, NI, D, H, Hg, F, B, N, steps=3):
    """Seeded gated-MLP model state (reference key layout), feature matrices in the reference's column layout and `steps`
    batches with Zipf positives (duplicate-heavy) at caller-chosen tower shapes."""
    rng = np.random.default_rng(seed)
    st = {}
    for side, n in (("user", NU), ("item", NI)):
        pre = f"{side}_encoder."
        st[pre + "embedding.weight"] = (rng.standard_normal((n, D)) * 0.02).astype(np.float32)
        st[pre + "feature_encoder.network.0.weight"] = (rng.standard_normal((H, F)) * np.sqrt(2.0 / (H + F))).astype(np.float32)
        st[pre + "feature_encoder.network.0.bias"] = (rng.standard_normal(H) * 0.05).astype(np.float32)
        st[pre + "feature_encoder.network.2.weight"] = (rng.standard_normal((D, H)) * np.sqrt(2.0 / (H + D))).astype(np.float32)
        st[pre + "feature_encoder.network.2.bias"] = (rng.standard_normal(D) * 0.05).astype(np.float32)
        st[pre + "adaptive_mimic.gate_network.0.weight"] = (rng.standard_normal((Hg, 2 * D)) * np.sqrt(2.0 / (Hg + 2 * D))).astype(np.float32)
        st[pre + "adaptive_mimic.gate_network.0.bias"] = (rng.standard_normal(Hg) * 0.05).astype(np.float32)
        st[pre + "adaptive_mimic.gate_network.2.weight"] = (rng.standard_normal((D, Hg)) * np.sqrt(2.0 / (Hg + D))).astype(np.float32)
        st[pre + "adaptive_mimic.gate_network.2.bias"] = (rng.standard_normal(D) * 0.05).astype(np.float32)
        st[f"adaptive_mimic.{side}_augmented.weight"] = (rng.standard_normal((n, D)) * 0.02).astype(np.float32)
    item_x = np.zeros((NI, F), dtype=np.float32)
    for r in range(NI):
        item_x[r, rng.choice(F - 5, size=3, replace=False)] = (1.0, 0.5, 1.0)
    item_x[:, F - 5:] = rng.standard_normal((NI, 5)).astype(np.float32)
    user_x = np.stack([item_x[rng.integers(0, NI, size=4)].mean(0) for _ in range(NU)]).astype(np.float32)
    pop = 1.0 / np.arange(1, NI + 1) ** 1.05
    pop /= pop.sum()
    batches = [(rng.integers(0, NU, size=B).astype(np.int64), rng.choice(NI, size=B, p=pop).astype(np.int64),
                rng.integers(0, NI, size=(B, N)).astype(np.int64)) for _ in range(steps)]
    return st, user_x, item_x, batches


# ---------------------------------------------------------------------------------------------
# GPU-side helpers (import torch lazily so that the CPU-only oracle tests stay light)
# ---------------------------------------------------------------------------------------------
def tower_cfg(meta, kw, state, side):
    D, H, Hg = meta["D"], meta["H"], meta["Hg"]
    sparse = kw.get("sparse", True)
    pre = f"{side}_encoder."
    if not any(k.startswith(pre + "feature_encoder") for k in state):
        return {"type": "embedding", "params": {"embedding_dim": D, "sparse": sparse}}
    fe_type = "linear" if pre + "feature_encoder.network.weight" in state else "mlp"
    cfg = {"type": "tower",
           "id_embedding": {"params": {"embedding_dim": D, "sparse": sparse}},
           "feature_encoder": {"type": fe_type, "hidden_dims": [H] if fe_type == "mlp" else None,
                               "activation": kw.get("activation", "relu"), "output_dim": D, "dropout": 0.0},
           "fusion": kw.get("fusion", "gated"), "output_dim": D}
    if Hg:
        cfg["adaptive_mimic"] = {"hidden_dim": Hg}
    return cfg


def build_model(meta, kw, state, device, pkg=None):
    """Our TwoTowerModel with the golden case's structure and `state` loaded (state: name -> np.ndarray)."""
    import torch
    if pkg is None:
        import two_tower_augmented_with_adaptive_mimic_mechanism_b200 as pkg
    ue = pkg.build_tower_encoder(tower_cfg(meta, kw, state, "user"), num_embeddings=meta["NU"], feature_dim=meta["F"])
    ie = pkg.build_tower_encoder(tower_cfg(meta, kw, state, "item"), num_embeddings=meta["NI"], feature_dim=meta["F"])
    mm = None
    if "adaptive_mimic.user_augmented.weight" in state:
        mm = pkg.AdaptiveMimicMechanism(num_users=meta["NU"], num_items=meta["NI"], embedding_dim=meta["D"])
    model = pkg.TwoTowerModel(ue, ie, adaptive_mimic=mm)
    missing = model.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in state.items()}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return model.to(device)


def model_state_np(model):
    return {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
