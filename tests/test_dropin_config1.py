"""BASELINE configs[0] through the UNMODIFIED reference `run_training` (training.py:1882-1897) with `hooks.install`, against the
reference's own CPU run of the same config on the same generated CSVs (SURVEY 8(b) fused-step hook, 8(d) config 1), and the
reference's own test files against the drop-in `src.models`.  Needs the reference install (baseline/_ref, see
scripts/install_reference.py) and a GPU."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

import dropin
import refenv

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(refenv.reference_root() is None, reason="reference install (baseline/_ref) not present")]

KW = dict(epochs=2, batch_size=512)


@pytest.fixture(scope="module")
def runs(tmp_path_factory):
    tmp = tmp_path_factory.mktemp("cfg1")
    dropin.make_data(tmp / "data", books=2000, users=600, per_user=24)
    rec = {}
    ref = dropin.run_reference(tmp / "data", tmp / "ref", record=rec, **KW)
    return tmp, ref, rec


@pytest.mark.parametrize("precision,loss_tol,state_tol", [("fp32", 2e-5, 2e-5), ("tf32", 2e-3, 2e-4)])
def test_run_training_hooked_matches_reference_cpu(runs, precision, loss_tol, state_tol):
    """Same seeds, same batches, same negatives (sampler="reference": the global torch generator is consumed exactly as in the
    CPU run), dropout 0.  Per-epoch train / validation / test losses, ranking metrics, predictions, checkpoint layout."""
    tmp, ref, rec_ref = runs
    rec = {}
    got = dropin.run_hooked(tmp / "data", tmp / f"hook_{precision}", precision=precision, graph=False, sampler="reference",
                            record=rec, **KW)
    c = dropin.compare(ref, got)
    assert c["train_loss_rel"] <= loss_tol and c["val_loss_rel"] <= loss_tol and c["test_loss_rel"] <= loss_tol, c
    assert c["ckpt_keys_equal"] and c["state_keys_equal"] and c["optimizer_layout_equal"], c
    assert c["state_mean_abs"] <= state_tol, c
    assert not [f for f in c["files_missing"] if ".png" not in f and "items.index" not in f], c   # no matplotlib / FAISS file format here
    agree = dropin.prediction_agreement(rec_ref["predictions"], rec["predictions"])
    assert agree >= (0.98 if precision == "fp32" else 0.5), agree    # identical top-20 LISTS; tf32 swaps near-ties
    assert c["val_metrics_max_abs"] <= (0.002 if precision == "fp32" else 0.01), c


def test_run_training_fast_path(runs):
    """What install() delivers by default: TF32 tensor cores, CUDA-graph replay, device sampler + device batch iterator.
    Different random draws than the CPU run, so the comparison is statistical."""
    tmp, ref, _ = runs
    got = dropin.run_hooked(tmp / "data", tmp / "hook_fast", **KW)
    for a, b in zip(ref["train_loss"], got["train_loss"]):
        assert abs(a - b) <= 0.05 * abs(a), (ref["train_loss"], got["train_loss"])
    for a, b in zip(ref["val_loss"], got["val_loss"]):
        assert abs(a - b) <= 0.08 * abs(a), (ref["val_loss"], got["val_loss"])
    assert got["stats"]["train_steps"] > 0 and got["stats"]["train_samples"] > 0
    assert got["checkpoint"] is not None and sorted(got["checkpoint"]) == sorted(ref["checkpoint"])


def test_sampling_evaluation_branch_matches_reference(runs):
    """Without FAISS the reference evaluates on ground truth + 50 sampled candidates (training.py:974-1009); eval_mode="reference"
    reproduces that branch, rng draws included."""
    tmp, _, _ = runs
    rec_a, rec_b = {}, {}
    ref = dropin.run_reference(tmp / "data", tmp / "ref_nofaiss", with_faiss=False, record=rec_a, epochs=1, batch_size=512)
    got = dropin.run_hooked(tmp / "data", tmp / "hook_sampling", precision="fp32", graph=False, sampler="reference",
                            eval_mode="reference", record=rec_b, epochs=1, batch_size=512)
    c = dropin.compare(ref, got)
    assert c["train_loss_rel"] <= 2e-5, c
    assert dropin.prediction_agreement(rec_a["predictions"], rec_b["predictions"]) >= 0.97
    assert c["val_metrics_max_abs"] <= 0.005, c


def test_reference_test_files_pass_against_dropin():
    """tests/test_encoders.py, test_adaptive_mimic.py, test_samplers.py of the reference, unmodified, with `src.models` (and
    the sampler) re-bound to this package (tests/ref_dropin_plugin.py)."""
    ref = refenv.reference_root()
    files = [str(ref / "tests" / n) for n in ("test_encoders.py", "test_adaptive_mimic.py", "test_samplers.py")]
    if not all(Path(f).exists() for f in files):
        pytest.skip("reference tests not installed")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([str(dropin.ROOT / "tests"), str(dropin.ROOT)]))
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "ref_dropin_plugin", "-p", "no:cacheprovider", "--rootdir", str(ref / "tests"), *files],
                       capture_output=True, text=True, env=env, cwd=str(ref), timeout=600)
    assert r.returncode == 0 and "4 passed" in r.stdout, r.stdout + r.stderr
