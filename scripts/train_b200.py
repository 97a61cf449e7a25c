#!/usr/bin/env python
"""`scripts/train.py` of the reference (scripts/train.py:20-35) with the B200 hot path installed.

    python scripts/train_b200.py --config configs/default.yaml [--reference /path/to/reference] \
        [--precision tf32|fp32] [--no-graph] [--sampler device|reference] [--eval-mode exact|reference] [--data-root DIR]

Nothing of the reference is modified: its package is put on sys.path (`--reference`, $TTAM_REFERENCE, or the repo's own
baseline/_ref install), `hooks.install` re-binds seven functions and three classes of `src.pipelines.training` in memory,
`model.device` defaults to "cuda", and `run_training(config)` runs as usual - data loading, splits, early stopping,
checkpoints, reports are the reference's own code.  matplotlib is not part of this image: a stub makes
`src.reporting.plots` importable (the loss-curve PNG is skipped).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def stub_matplotlib() -> None:
    try:
        import matplotlib  # noqa: F401
        return
    except ImportError:
        pass
    m, p = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
    m.use = lambda *a, **k: None

    def _unavailable(*a, **k):
        raise ValueError("matplotlib is not installed: loss curves are skipped")   # save_loss_curves' callers catch ValueError

    for name in ("subplots", "figure", "plot", "savefig", "close"):
        setattr(p, name, _unavailable)
    m.pyplot = p
    sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = m, p


def reference_root(arg: str | None) -> Path:
    for cand in (arg, os.environ.get("TTAM_REFERENCE"), ROOT / "baseline" / "_ref"):
        if cand and (Path(cand) / "src" / "pipelines" / "training.py").exists():
            return Path(cand)
    raise SystemExit("reference not found: pass --reference DIR (a checkout holding src/pipelines/training.py) or run "
                     "scripts/install_reference.py")


def load_training_module(ref: Path):
    stub_matplotlib()
    if str(ref) not in sys.path:
        sys.path.insert(0, str(ref))
    import src.pipelines.training as training
    return training


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--config", type=Path, default=None)
    ap.add_argument("--reference", default=None)
    ap.add_argument("--precision", default="tf32", choices=["tf32", "fp32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--sampler", default="device", choices=["device", "reference"])
    ap.add_argument("--eval-mode", default="exact", choices=["exact", "reference"])
    ap.add_argument("--data-root", default=None, help="override data.root (and use books.csv / users.csv found there)")
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--set", action="append", default=[], metavar="dotted.key=json", help="override a config entry")
    a = ap.parse_args(argv)
    ref = reference_root(a.reference)
    training = load_training_module(ref)
    from src.utils import load_config, set_by_dotted_path
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import hooks
    cfg = load_config(a.config or ref / "configs" / "default.yaml")
    cfg.setdefault("model", {})["device"] = a.device
    if a.data_root:
        cfg.setdefault("data", {})["root"] = a.data_root
    for item in a.set:
        key, _, val = item.partition("=")
        set_by_dotted_path(cfg, key, json.loads(val))
    stats: dict = {}
    hooks.install(training, precision=a.precision, graph=not a.no_graph, sampler=a.sampler, eval_mode=a.eval_mode, stats=stats)
    t0 = time.time()
    training.run_training(cfg)
    if stats.get("train_seconds"):
        print(json.dumps({"train_samples": stats["train_samples"], "train_seconds": stats["train_seconds"],
                          "samples_per_s_in_train_one_epoch": stats["train_samples"] / stats["train_seconds"],
                          "epoch_samples_per_s": [n / s for n, s in zip(stats["epoch_samples"], stats["epoch_seconds"])],
                          "wall_seconds": time.time() - t0}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
