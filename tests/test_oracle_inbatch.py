"""The in-batch softmax extension of the oracle (oracle/model.py::inbatch_loss_forward_backward) against torch autograd.
The reference has no in-batch loss: this pins the DEFINITION the round-2 kernel will be tested
against, not reference parity."""
import numpy as np
import pytest
import torch

import oracle


@pytest.mark.parametrize("B,D,mimic", [(1, 8, False), (7, 16, True), (64, 96, True), (33, 96, False)])
def test_inbatch_softmax_matches_torch_autograd(B, D, mimic):
    rng = np.random.default_rng(B * 100 + D)
    mk = lambda: (rng.standard_normal((B, D)) * 0.5).astype(np.float32)
    t_u, t_p, q_u, q_p = mk(), mk(), mk() * 0.1, mk() * 0.1
    lu, li = 0.15, 0.25
    kw = dict(t_u=t_u, t_p=t_p, q_u=q_u, q_p=q_p, lambda_u=lu, lambda_i=li) if mimic else {}
    got = oracle.inbatch_loss_forward_backward(t_u + q_u, t_p + q_p, **kw)
    T = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in dict(t_u=t_u, t_p=t_p, q_u=q_u, q_p=q_p).items()}
    o_u, o_p = T["t_u"] + T["q_u"], T["t_p"] + T["q_p"]
    o_u.retain_grad(); o_p.retain_grad()
    ce = torch.nn.functional.cross_entropy(o_u @ o_p.T, torch.arange(B))
    loss = ce
    if mimic:        # adaptive_mimic.py:66-67: the targets are detached
        mu = ((T["q_u"] - T["t_p"].detach()) ** 2).mean()
        mi = ((T["q_p"] - T["t_u"].detach()) ** 2).mean()
        loss = ce + lu * mu + li * mi
    loss.backward()
    assert float(got["ce"]) == pytest.approx(ce.item(), rel=2e-6, abs=1e-7)
    assert float(got["loss"]) == pytest.approx(loss.item(), rel=2e-6, abs=1e-7)
    np.testing.assert_allclose(got["do_u"], o_u.grad.numpy(), rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(got["do_p"], o_p.grad.numpy(), rtol=2e-5, atol=1e-7)
    if mimic:        # dL/dq = dL/do + the mimic term
        np.testing.assert_allclose(got["do_u"] + got["dq_u_extra"], T["q_u"].grad.numpy(), rtol=2e-5, atol=1e-7)
        np.testing.assert_allclose(got["do_p"] + got["dq_p_extra"], T["q_p"].grad.numpy(), rtol=2e-5, atol=1e-7)
        assert float(got["mimic_user"]) == pytest.approx(mu.item(), rel=2e-6)


def test_inbatch_softmax_properties():
    """Size-independent properties: B identical pairs give log(B); permuting the pairs of a batch leaves the loss unchanged
    and permutes the gradients the same way."""
    B, D = 16, 8
    o = np.ones((B, D), np.float32)
    assert float(oracle.inbatch_loss_forward_backward(o, o)["ce"]) == pytest.approx(np.log(B), rel=1e-6)
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal((B, D)).astype(np.float32), rng.standard_normal((B, D)).astype(np.float32)
    g = oracle.inbatch_loss_forward_backward(a, b)
    perm = rng.permutation(B)
    gp = oracle.inbatch_loss_forward_backward(a[perm], b[perm])
    assert float(gp["ce"]) == pytest.approx(float(g["ce"]), rel=1e-6)
    np.testing.assert_allclose(gp["do_u"], g["do_u"][perm], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(gp["do_p"], g["do_p"][perm], rtol=1e-5, atol=1e-7)
