#!/usr/bin/env python
"""Drop-in report for BASELINE configs[0]: the reference's CPU run next to the hooked B200 runs of the same unmodified
`run_training`, on generated CSVs.  Prints one JSON object (comparisons + timings); used to set the thresholds of
tests/test_dropin_config1.py and kept under profiles/.

    python scripts/dropin_report.py [--books 2000 --users 600 --epochs 2] [--out gpurun_out/dropin_report.json]
"""
from __future__ import annotations

import argparse
import json
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tests"))

import dropin  # noqa: E402


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--books", type=int, default=2000)
    ap.add_argument("--users", type=int, default=600)
    ap.add_argument("--per-user", type=int, default=24)
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--batch-size", type=int, default=512)
    ap.add_argument("--out", default=None)
    ap.add_argument("--skip-reference", action="store_true")
    a = ap.parse_args()
    report = {}
    with tempfile.TemporaryDirectory() as tmp:
        tmp = Path(tmp)
        report["data"] = dropin.make_data(tmp / "data", books=a.books, users=a.users, per_user=a.per_user)
        kw = dict(epochs=a.epochs, batch_size=a.batch_size)
        rec_ref, ref = {}, None
        if not a.skip_reference:
            ref = dropin.run_reference(tmp / "data", tmp / "ref", record=rec_ref, **kw)
            report["reference_cpu"] = {k: ref[k] for k in ("train_loss", "val_loss", "test_loss", "val_metrics", "seconds")}
        for name, opts in (("fp32_parity", dict(precision="fp32", graph=False, sampler="reference")),
                           ("tf32_parity", dict(precision="tf32", graph=False, sampler="reference")),
                           ("fast", dict(precision="tf32", graph=True, sampler="device"))):
            rec = {}
            got = dropin.run_hooked(tmp / "data", tmp / name, record=rec, **opts, **kw)
            entry = {k: got[k] for k in ("train_loss", "val_loss", "test_loss", "val_metrics", "seconds", "stats")}
            if ref is not None:
                entry["vs_reference"] = dropin.compare(ref, got)
                entry["prediction_agreement"] = dropin.prediction_agreement(rec_ref.get("predictions", []), rec.get("predictions", []))
            report[name] = entry
        # reference WITHOUT FAISS (its candidate-sampling evaluation) against eval_mode="reference"
        if not a.skip_reference:
            rec_a, rec_b = {}, {}
            ref2 = dropin.run_reference(tmp / "data", tmp / "ref_nofaiss", with_faiss=False, record=rec_a, **kw)
            got2 = dropin.run_hooked(tmp / "data", tmp / "sampling", precision="fp32", graph=False, sampler="reference",
                                     eval_mode="reference", record=rec_b, **kw)
            report["sampling_eval"] = {"vs_reference": dropin.compare(ref2, got2),
                                       "prediction_agreement": dropin.prediction_agreement(rec_a["predictions"], rec_b["predictions"])}
    text = json.dumps(report, indent=1, default=str)
    print(text)
    if a.out:
        Path(a.out).parent.mkdir(parents=True, exist_ok=True)
        Path(a.out).write_text(text)
    return 0


if __name__ == "__main__":
    sys.exit(main())
