"""In-batch softmax loss (extension; oracle definition pinned against torch autograd in tests/test_oracle_inbatch.py, NOT
against the reference, which has no such loss): kernel vs oracle, fused step vs oracle step, size-independent properties at the
bench's batch size."""
import numpy as np
import pytest
import torch

import oracle
from helpers import build_model, model_state_np, synthetic_gated

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("B,D,mimic,precision", [(1, 8, False, "fp32"), (7, 16, True, "fp32"), (64, 96, True, "fp32"), (300, 96, True, "fp32"),
                                                 (1000, 96, True, "tf32"), (257, 128, False, "tf32"), (2048, 96, True, "tf32")])
def test_inbatch_loss_kernel_matches_oracle(B, D, mimic, precision):
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import functional as F
    rng = np.random.default_rng(B * 100 + D)
    mk = lambda s: (rng.standard_normal((B, D)) * s).astype(np.float32)
    t_u, t_p, q_u, q_p = mk(0.3), mk(0.3), mk(0.05), mk(0.05)
    o_u, o_p = t_u + q_u, t_p + q_p
    kw = dict(t_u=t_u, t_p=t_p, q_u=q_u, q_p=q_p, lambda_u=0.15, lambda_i=0.25) if mimic else {}
    ref = oracle.inbatch_loss_forward_backward(o_u, o_p, **kw)
    kwd = {k: (dev(v) if isinstance(v, np.ndarray) else v) for k, v in kw.items()}
    loss, do_u, do_p, dq_u, dq_p = F.inbatch_loss_fwd_bwd(dev(o_u), dev(o_p), precision=precision, **kwd)
    ltol, gtol = (5e-6, 2e-5) if precision == "fp32" else (2e-3, 3e-3)
    got = loss.cpu().numpy()
    assert got[0] == pytest.approx(float(ref["loss"]), rel=ltol) and got[1] == pytest.approx(float(ref["ce"]), rel=ltol)
    scale = np.abs(ref["do_u"]).max() + 1e-12
    assert np.abs(do_u.cpu().numpy() - ref["do_u"]).max() <= gtol * scale
    assert np.abs(do_p.cpu().numpy() - ref["do_p"]).max() <= gtol * scale
    if mimic:
        assert got[2] == pytest.approx(float(ref["mimic_user"]), rel=5e-6) and got[3] == pytest.approx(float(ref["mimic_item"]), rel=5e-6)
        assert np.abs(dq_u.cpu().numpy() - (ref["do_u"] + ref["dq_u_extra"])).max() <= gtol * scale
        assert np.abs(dq_p.cpu().numpy() - (ref["do_p"] + ref["dq_p_extra"])).max() <= gtol * scale
    fwd_only, *_ = F.inbatch_loss_fwd_bwd(dev(o_u), dev(o_p), backward=False, precision=precision, **kwd)
    assert torch.equal(fwd_only, loss)


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_inbatch_loss_properties_at_bench_batch(precision):
    """B = 8192: identical item rows give exactly log(B) whatever the users are, and then every gradient row of the users is
    (sum_j P_bj) c = 0; permuting the pairs leaves the loss unchanged."""
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import functional as F
    B, D = 8192, 96
    g = torch.Generator(device="cuda").manual_seed(0)
    o_u = torch.randn((B, D), device="cuda", generator=g) * 0.3
    c = torch.randn((1, D), device="cuda", generator=g) * 0.3
    loss, do_u, do_p, _, _ = F.inbatch_loss_fwd_bwd(o_u, c.expand(B, D).contiguous(), precision=precision)
    assert float(loss[0]) == pytest.approx(np.log(B), rel=2e-6)
    assert float(do_u.abs().max()) <= 1e-7                      # rounding of sum_j P_bj c; regular gradient rows are ~1e-5
    o_p = torch.randn((B, D), device="cuda", generator=g) * 0.3
    l0 = float(F.inbatch_loss_fwd_bwd(o_u, o_p, precision=precision)[0][0])
    perm = torch.randperm(B, device="cuda", generator=g)
    l1 = float(F.inbatch_loss_fwd_bwd(o_u[perm].contiguous(), o_p[perm].contiguous(), precision=precision)[0][0])
    assert l1 == pytest.approx(l0, rel=1e-5)


@pytest.mark.parametrize("precision,graph", [("fp32", False), ("tf32", True)])
def test_fused_step_inbatch_matches_oracle(precision, graph):
    """The whole step with loss="inbatch" (towers, in-batch softmax + mimic, backward, SparseAdam / lazy AdamW) against the
    oracle's step with the same loss."""
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import FusedEngine
    NU, NI, D, H, Hg, F_, B = 3000, 5000, 96, 192, 96, 605, 512
    st, user_x, item_x, batches = synthetic_gated(11, NU, NI, D, H, Hg, F_, B, 5)
    meta = dict(NU=NU, NI=NI, D=D, H=H, Hg=Hg, F=F_)
    model = build_model(meta, dict(optimizer="adamw"), st, "cuda")
    eng = FusedEngine(model, optimizer="adamw", lr=1e-3, weight_decay=0.01, loss_weights={"mimic_user": 0.15, "mimic_item": 0.15},
                      max_steps=16, precision=precision, loss="inbatch")
    ref_state = {k: v.copy() for k, v in st.items()}
    spec, opt = oracle.spec_from_state(ref_state), oracle.OptState()
    ux, ix = torch.from_numpy(user_x).cuda(), torch.from_numpy(item_x).cuda()
    for u, p, _ in batches:
        ref = oracle.train_step(ref_state, opt, spec, u, p, None, user_x, item_x, lr=1e-3, weight_decay=0.01,
                                lambdas=(0.15, 0.15, 0.0), loss="inbatch")
        loss = eng.train_step(torch.from_numpy(u).cuda(), torch.from_numpy(p).cuda(), None, ux, ix, graph=graph)
        assert float(loss[0]) == pytest.approx(ref["loss"], rel=5e-6 if precision == "fp32" else 2e-3)
    eng.flush()
    got = model_state_np(model)
    for k in ref_state:
        d = np.abs(got[k] - ref_state[k])
        if precision == "fp32":
            bad = d > 2e-6 + 5e-5 * np.abs(ref_state[k])
            assert bad.sum() <= max(1, int(1e-4 * bad.size)), (k, int(bad.sum()))     # Adam-normalised ~0 gradients (1/B-scaled here)
        else:
            assert d.mean() <= 5e-5, (k, d.mean())
    touched = np.unique(np.concatenate([b[1] for b in batches]))       # item rows: only the positives are touched
    changed = np.nonzero((got["item_encoder.embedding.weight"] != st["item_encoder.embedding.weight"]).any(1))[0]
    assert np.array_equal(changed, touched)
