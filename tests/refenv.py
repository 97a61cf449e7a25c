"""Test-side access to the UNMODIFIED reference (baseline/_ref, installed by scripts/install_reference.py; falls back to
/root/reference in the build container) and a numpy stand-in for the one third-party dependency of the path that cannot be
installed here: `faiss.IndexFlatIP` (faiss-cpu>=1.7.4, reference pyproject.toml:24; call sites training.py:672-675,696,
955-958).  IndexFlatIP is exact inner-product search; the stand-in restates that with the project's canonical order
(descending score, ascending id).  TEST INFRASTRUCTURE: nothing in the product package imports this file."""
from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]


def reference_root():
    for cand in (ROOT / "baseline" / "_ref", Path("/root/reference")):
        if (cand / "src" / "pipelines" / "training.py").exists():
            return cand
    return None


def load_training():
    """Fresh import of the reference's `src.pipelines.training` (every call: hooks.install mutates the module)."""
    ref = reference_root()
    if ref is None:
        return None
    sys.path.insert(0, str(ROOT / "scripts"))
    from train_b200 import stub_matplotlib
    stub_matplotlib()
    for name in [n for n in sys.modules if n == "src" or n.startswith("src.")]:
        del sys.modules[name]
    if str(ref) not in sys.path:
        sys.path.insert(0, str(ref))
    import src.pipelines.training as training
    return training


class _IndexFlatIP:
    def __init__(self, d: int) -> None:
        self.d, self.x = int(d), np.zeros((0, int(d)), dtype=np.float32)

    @property
    def ntotal(self) -> int:
        return self.x.shape[0]

    def add(self, x) -> None:
        self.x = np.concatenate([self.x, np.asarray(x, dtype=np.float32)], axis=0)

    def search(self, q, k: int):
        from oracle.retrieval import canonical_scores, topk_canonical
        q = np.asarray(q, dtype=np.float32)
        k_eff = min(int(k), self.ntotal)
        ids, sc = topk_canonical(canonical_scores(q, self.x), k_eff)
        if k_eff < k:
            ids = np.concatenate([ids, np.full((q.shape[0], k - k_eff), -1, dtype=np.int64)], axis=1)
            sc = np.concatenate([sc, np.full((q.shape[0], k - k_eff), -np.inf, dtype=np.float32)], axis=1)
        return sc, ids


def fake_faiss():
    m = types.ModuleType("faiss")

    def normalize_L2(x):                       # in place, like faiss.normalize_L2
        n = np.sqrt((x.astype(np.float32) ** 2).sum(axis=1, keepdims=True, dtype=np.float32))
        x /= np.maximum(n, np.float32(1e-12))      # faiss skips zero rows; none occur here

    def write_index(index, path):
        np.save(str(path) + ".npy", index.x)

    m.IndexFlatIP, m.normalize_L2, m.write_index = _IndexFlatIP, normalize_L2, write_index
    return m


def config1(data_root, out_dir, *, device="cpu", dropout=0.0, epochs=2, batch_size=512, category_alignment=0.01,
            similarity="cosine", dim=128, hidden=256, books_file="books.csv", users_file="users.csv", early=True):
    """configs/default.yaml of the reference (model / training / evaluation blocks, values as shipped) with the data root,
    artefact paths, device and - for parity - the dropout rate replaced."""
    out = Path(out_dir)

    def tower():
        return {"type": "tower",
                "id_embedding": {"params": {"embedding_dim": dim, "sparse": True}, "init": {"type": "normal", "std": 0.02}},
                "feature_encoder": {"type": "mlp", "hidden_dims": [hidden], "activation": "relu", "output_dim": dim,
                                    "dropout": dropout},
                "fusion": "gated", "output_dim": dim}
    return {
        "experiment": {"name": "baseline_two_tower", "seed": 1234, "benchmark_report": str(out / "benchmark_summary.md"),
                       "grid": {}},
        "data": {"root": str(data_root), "books_file": books_file, "users_file": users_file, "train_fraction": 0.85,
                 "test_fraction": 0.15, "books_limit": None, "interactions_limit": 2000000, "min_user_interactions": 3,
                 "min_item_interactions": 6,
                 "feature_params": {"numeric_columns": ["average_rating", "price", "rating_number"], "category_top_k": 300,
                                    "author_top_k": 300, "user_aggregation": "mean"}},
        "model": {"user_encoder": tower(), "item_encoder": tower(), "similarity": similarity, "device": device,
                  "adaptive_mimic": {"enabled": True, "init_std": 0.02}},
        "training": {"batch_size": batch_size, "num_epochs": epochs, "learning_rate": 0.001, "weight_decay": 0.01,
                     "optimizer": "adamw", "negatives_per_positive": 5, "gradient_clip_norm": None,
                     "loss_weights": {"mimic_user": 0.15, "mimic_item": 0.15, "category_alignment": category_alignment},
                     "early_stopping": {"enabled": early, "metric": "recall@10", "mode": "max", "patience": 2, "min_delta": 0.0005},
                     "checkpointing": {"enabled": True, "dir": str(out / "checkpoints"), "save_best_only": True,
                                       "keep_last": True,
                                       "filename_template": "{experiment}_{metric}_{value:.4f}_epoch{epoch}.pt"}},
        "evaluation": {"metrics_k": [5, 10, 20], "candidate_samples": 50, "holdout": "latest_per_user",
                       "faiss": {"enabled": True, "search_k_multiplier": 4, "batch_size": 8192,
                                 "index_path": str(out / "faiss" / "items.index"),
                                 "embedding_path": str(out / "faiss" / "item_embeddings.npy")}},
        "recommendations": {"sample_users": 2, "top_k": 5},
        "diagnostics": {"item_sample_size": 10, "user_sample_size": 100, "neighbor_k": 5,
                        "report_path": str(out / "reports" / "recommendation_report.md"),
                        "loss_plot_path": str(out / "reports" / "loss_curve.png"),
                        "embedding_summary_path": str(out / "reports" / "embedding_diagnostics.json"),
                        "feature_corr_top_k": 15},
        "logging": {"level": "INFO"},
    }
