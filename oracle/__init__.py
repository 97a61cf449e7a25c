"""CPU oracle for the two-tower train + retrieval hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain numpy (fp32) restatement of what the reference computes on the hot path
named by BASELINE.json (`src/models/*`, `src/pipelines/training.py:700-1043`, `src/data/samplers.py`,
and the third-party arithmetic it reaches: `torch.optim.SparseAdam` = torch/optim/_functional.py:24-84,
`torch.optim.AdamW/Adam` = torch/optim/adam.py:347-547 (`_single_tensor_adam`), `torch.optim.SGD`,
and `faiss.IndexFlatIP` = exact inner product, faiss-cpu>=1.7.4 per pyproject.toml:24).

Parity status: PINNED.  Every function here is checked in tests/test_oracle_golden.py against the
fixtures under tests/golden/*.npz, which were produced by *running the unmodified reference itself*
(tests/golden/make_golden.py, committed).  The reference's own tests hold no numeric vectors for this
path except the metric known-answers of tests/test_metrics.py, which are
reproduced in tests/test_oracle_golden.py as well.  The FAISS boundary has no reference test at all;
it is pinned by the exact-inner-product definition only.

Rules: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package, and only as the checker or as the CPU arm being timed.  The product package
never imports it and has no CPU fallback.
"""
from .model import (  # noqa: F401
    ModelSpec,
    spec_from_state,
    tower_forward,
    tower_backward,
    loss_forward_backward,
    inbatch_loss_forward_backward,
    category_alignment_loss,
    train_step,
    eval_loss,
    encode_items,
    encode_users,
)
from .optim import sparse_adam_step, dense_step, OptState  # noqa: F401
from .retrieval import (  # noqa: F401
    canonical_scores,
    topk_canonical,
    evaluate_flat_ip,
    ranking_metrics,
    score_all_items_topk,
)
