"""Runs BASELINE configs[0] twice through the reference's UNMODIFIED `run_training` (training.py:1882-1897): once as the
reference runs it (CPU PyTorch; FAISS replaced by the numpy stand-in of tests/refenv.py, or absent), once with
`hooks.install` (B200 kernels), and compares what the two runs return and write.  Shared by tests/test_dropin_config1.py
and scripts/dropin_report.py.  TEST INFRASTRUCTURE."""
from __future__ import annotations

import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "tests", ROOT / "scripts"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

import refenv  # noqa: E402


def make_data(out: Path, **kw) -> dict:
    from make_config1_data import generate
    return generate(out, **kw)


def _result(res, out_dir: Path, seconds: float) -> dict:
    res = res[0] if isinstance(res, list) else res
    ck = sorted((out_dir / "checkpoints").glob("*_last.pt"))
    metrics = lambda m: None if m is None else {name: dict(getattr(m, name)) for name in ("recall", "precision", "ndcg", "hit_rate", "map")}
    return {"train_loss": list(res.history.train_loss), "val_loss": list(res.history.val_loss),
            "test_loss": list(res.history.test_loss), "val_metrics": metrics(res.val_metrics),
            "test_metrics": metrics(res.test_metrics), "best_epoch": res.best_epoch, "seconds": seconds,
            "checkpoint": torch.load(ck[0], map_location="cpu", weights_only=False) if ck else None,
            "files": sorted(str(p.relative_to(out_dir)) for p in out_dir.rglob("*") if p.is_file())}


def run_reference(data_root: Path, out_dir: Path, *, with_faiss: bool = True, record=None, **cfg_kw) -> dict:
    """The reference as it is, on the CPU.  record: optional dict that receives the per-call predictions of `_evaluate_model`."""
    tr = refenv.load_training()
    if with_faiss:
        tr.faiss = refenv.fake_faiss()
    if record is not None:
        inner = tr._evaluate_model

        def spy(*a, **k):
            out = inner(*a, **k)
            record.setdefault("predictions", []).append(out[0])
            return out
        tr._evaluate_model = spy
    cfg = refenv.config1(data_root, out_dir, device="cpu", **cfg_kw)
    t0 = time.time()
    res = tr.run_training(cfg)
    return _result(res, Path(out_dir), time.time() - t0)


def run_hooked(data_root: Path, out_dir: Path, *, precision="tf32", graph=True, sampler="device", eval_mode="exact",
               record=None, **cfg_kw) -> dict:
    """The same `run_training`, with hooks.install and model.device = cuda."""
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import hooks
    tr = refenv.load_training()
    stats: dict = {}
    hooks.install(tr, precision=precision, graph=graph, sampler=sampler, eval_mode=eval_mode, stats=stats)
    if record is not None:
        inner = tr._evaluate_model

        def spy(*a, **k):
            out = inner(*a, **k)
            record.setdefault("predictions", []).append(out[0])
            return out
        tr._evaluate_model = spy
    cfg = refenv.config1(data_root, out_dir, device="cuda", **cfg_kw)
    t0 = time.time()
    res = tr.run_training(cfg)
    out = _result(res, Path(out_dir), time.time() - t0)
    out["stats"] = stats
    return out


def compare(ref: dict, got: dict) -> dict:
    """Relative loss differences, metric differences, checkpoint layout / value differences."""
    rel = lambda a, b: float(np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.maximum(np.abs(np.asarray(a)), 1e-12))) if len(a) else 0.0
    out = {"train_loss_rel": rel(ref["train_loss"], got["train_loss"]), "val_loss_rel": rel(ref["val_loss"], got["val_loss"]),
           "test_loss_rel": rel(ref["test_loss"], got["test_loss"]), "best_epoch": (ref["best_epoch"], got["best_epoch"])}
    for split in ("val_metrics", "test_metrics"):
        a, b = ref[split], got[split]
        if a and b:
            out[split + "_max_abs"] = max(abs(a[n][k] - b[n][k]) for n in a for k in a[n])
    ca, cb = ref["checkpoint"], got["checkpoint"]
    if ca and cb:
        out["ckpt_keys_equal"] = sorted(ca) == sorted(cb)
        sa, sb = ca["model_state_dict"], cb["model_state_dict"]
        out["state_keys_equal"] = list(sa) == list(sb) and all(sa[k].shape == sb[k].shape for k in sa)
        diffs = {k: (sa[k].float() - sb[k].float().cpu()).abs() for k in sa}
        out["state_mean_abs"] = max(float(d.mean()) for d in diffs.values())
        out["state_max_abs"] = max(float(d.max()) for d in diffs.values())
        oa, ob = ca["optimizer_state_dicts"], cb["optimizer_state_dicts"]
        same = len(oa) == len(ob)
        if same:
            for x, y in zip(oa, ob):
                same &= sorted(x) == sorted(y) and sorted(x["state"]) == sorted(y["state"])
                same &= [g["params"] for g in x["param_groups"]] == [g["params"] for g in y["param_groups"]]
                for i in x["state"]:
                    same &= sorted(x["state"][i]) == sorted(y["state"].get(i, {}))
        out["optimizer_layout_equal"] = bool(same)
    out["files_missing"] = sorted(set(ref["files"]) - set(got["files"]))
    return out


def prediction_agreement(ref_preds: list, got_preds: list) -> float:
    """Fraction of (evaluation call, user) pairs whose predicted id lists are identical."""
    same = total = 0
    for a, b in zip(ref_preds, got_preds):
        for u, ids in a.items():
            total += 1
            same += int(list(ids) == list(b.get(u, ())))
    return same / max(total, 1)
