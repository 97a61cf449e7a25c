"""Kernel-level parity through the C ABI (functional.py is a 1:1 ctypes wrapper) against the numpy oracle."""
import numpy as np
import pytest
import torch

import oracle
from oracle.optim import dense_step as o_dense_step, sparse_adam_step as o_sparse_adam
from helpers import GOLDEN

pytestmark = pytest.mark.gpu

# fp32 SIMT kernels vs numpy fp32: same arithmetic, different summation order
RTOL, ATOL = 2e-5, 2e-6


@pytest.fixture(scope="module")
def F():
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import functional
    assert functional.lib().ttam_device_ok() == 1
    return functional


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("R,ncols", [(0, 96), (1, 96), (1000, 96), (333, 21), (4097, 608), (77, 605)])
def test_gather_rows_bit_exact(F, R, ncols):
    rng = np.random.default_rng(R + ncols)
    table = rng.standard_normal((5000, ncols)).astype(np.float32)
    idx = rng.integers(0, 5000, size=R).astype(np.int64)
    out = F.gather_rows(dev(table), dev(idx))
    assert np.array_equal(out.cpu().numpy(), table[idx])


def test_gather_into_strided_view(F):
    rng = np.random.default_rng(0)
    table = rng.standard_normal((100, 16)).astype(np.float32)
    idx = rng.integers(0, 100, size=50).astype(np.int64)
    z = torch.zeros((50, 32), device="cuda")
    F.gather_rows(dev(table), dev(idx), out=z[:, :16])
    assert np.array_equal(z[:, :16].cpu().numpy(), table[idx]) and float(z[:, 16:].abs().sum()) == 0.0


def test_cast_bf16_rne(F):
    torch.manual_seed(1234)
    x = torch.randn(513, 96, device="cuda")
    x[0, 0], x[0, 1], x[0, 2] = float("inf"), -0.0, 1.0 + 2 ** -8     # tie -> even
    got = F.cast_bf16(x)
    assert torch.equal(got.view(torch.int16), x.to(torch.bfloat16).view(torch.int16))


@pytest.mark.parametrize("M,N,K,act,gather", [(257, 32, 21, "relu", True), (64, 16, 32, "none", False),
                                              (1000, 192, 605, "relu", True), (130, 96, 192, "tanh", False),
                                              (65, 24, 32, "gelu", False), (65, 24, 32, "selu", False)])
def test_linear_fwd(F, M, N, K, act, gather):
    rng = np.random.default_rng(M)
    X = rng.standard_normal((2000, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32) * 0.1
    idx = rng.integers(0, 2000, size=M).astype(np.int64)
    x = X[idx] if gather else X[:M]
    ref = x @ W.T + b
    if act != "none":
        ref = oracle.model._act(act, ref.astype(np.float32))
    got = F.linear_fwd(dev(X) if gather else dev(x), dev(W), dev(b), gather=dev(idx) if gather else None, act=act)
    np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=RTOL, atol=ATOL)


def test_linear_dgrad_and_wgrad(F):
    rng = np.random.default_rng(5)
    M, N, K = 999, 48, 37
    X = rng.standard_normal((300, K)).astype(np.float32)
    idx = rng.integers(0, 300, size=M).astype(np.int64)
    dy = rng.standard_normal((M, N)).astype(np.float32)
    W = rng.standard_normal((N, K)).astype(np.float32)
    aux = rng.standard_normal((M, K)).astype(np.float32)
    got = F.linear_dgrad(dev(dy), dev(W), aux=dev(aux), relu_mask=True, scale=1.25)
    np.testing.assert_allclose(got.cpu().numpy(), (dy @ W) * (aux > 0) * 1.25, rtol=RTOL, atol=1e-5)
    base = rng.standard_normal((M, K)).astype(np.float32)
    out = dev(base)
    F.linear_dgrad(dev(dy), dev(W), out=out, accumulate=True)
    np.testing.assert_allclose(out.cpu().numpy(), base + dy @ W, rtol=RTOL, atol=1e-5)
    dw, db = F.linear_wgrad(dev(dy), dev(X), gather=dev(idx))
    np.testing.assert_allclose(dw.cpu().numpy(), dy.T @ X[idx], rtol=1e-4, atol=2e-4)
    np.testing.assert_allclose(db.cpu().numpy(), dy.sum(0), rtol=1e-4, atol=2e-4)
    # deterministic: same launch twice gives the same bits
    dw2, db2 = F.linear_wgrad(dev(dy), dev(X), gather=dev(idx))
    assert torch.equal(dw, dw2) and torch.equal(db, db2)


# ---------------------------------------------------------------------------------------------
# TF32 tensor-core path (tcgen05 kind::tf32, operands rounded to 10-bit mantissas, fp32 accumulate).
# Tolerance: |err| <= 2 * 2^-11 * sum_k |a_k b_k| (two rounded operands); checked against an fp64 product.
# ---------------------------------------------------------------------------------------------
def _tf32_close(got, a, b, extra=0.0):
    """got ~= a @ b with the TF32 operand-rounding bound (a: [M,K], b: [K,N])."""
    ref = a.astype(np.float64) @ b.astype(np.float64)
    bound = 2.0 * 2.0 ** -11 * (np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64)) + 1e-5 + extra
    err = np.abs(got.astype(np.float64) - ref)
    assert (err <= bound).all(), f"max err {err.max():.3e}, bound at that point {bound.flat[err.argmax()]:.3e}"
    # and on average it is much better than the bound (round-to-nearest, errors cancel)
    assert err.mean() <= 0.2 * bound.mean()


@pytest.mark.parametrize("M,N,K,gather,pad", [(1000, 192, 605, True, True), (1000, 192, 605, True, False),
                                               (130, 96, 192, False, False), (257, 96, 96, False, False),
                                               (64, 16, 32, False, False), (4099, 256, 64, True, False),
                                               (333, 40, 21, True, False), (128, 300, 128, False, False)])
def test_linear_fwd_tf32(F, M, N, K, gather, pad):
    rng = np.random.default_rng(M + N)
    X = rng.standard_normal((2000, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    b = (rng.standard_normal(N) * 0.1).astype(np.float32)
    idx = rng.integers(0, 2000, size=M).astype(np.int64)
    x = X[idx] if gather else X[:M]
    xd = dev(X) if gather else dev(x)
    wd = dev(W)
    if pad:                                # 16-byte aligned rows: the vectorised loader
        xd, wd = F.pad_cols(xd), F.pad_cols(wd)
        assert xd.stride(0) % 4 == 0 and wd.stride(0) % 4 == 0
    got = F.linear_fwd(xd, wd, dev(b), gather=dev(idx) if gather else None, act="none", precision="tf32").cpu().numpy()
    _tf32_close(got - b, x, W.T)
    got_relu = F.linear_fwd(xd, wd, dev(b), gather=dev(idx) if gather else None, act="relu", precision="tf32").cpu().numpy()
    assert np.array_equal(got_relu, np.maximum(got, 0))


def test_linear_dgrad_and_wgrad_tf32(F):
    rng = np.random.default_rng(11)
    for M, N, K in [(999, 48, 37), (4096, 96, 192), (3000, 192, 605), (700, 96, 96)]:
        X = rng.standard_normal((500, K)).astype(np.float32)
        idx = rng.integers(0, 500, size=M).astype(np.int64)
        dy = rng.standard_normal((M, N)).astype(np.float32)
        W = rng.standard_normal((N, K)).astype(np.float32)
        aux = rng.standard_normal((M, K)).astype(np.float32)
        got = F.linear_dgrad(dev(dy), dev(W), precision="tf32").cpu().numpy()
        _tf32_close(got, dy, W)
        got_m = F.linear_dgrad(dev(dy), dev(W), aux=dev(aux), relu_mask=True, scale=1.25, precision="tf32").cpu().numpy()
        np.testing.assert_allclose(got_m, got * (aux > 0) * np.float32(1.25), rtol=1e-6, atol=1e-7)
        base = rng.standard_normal((M, K)).astype(np.float32)
        out = dev(base)
        F.linear_dgrad(dev(dy), dev(W), out=out, accumulate=True, precision="tf32")
        np.testing.assert_allclose(out.cpu().numpy(), base + got, rtol=1e-6, atol=1e-6)
        xd = F.pad_cols(dev(X))
        dw, db = F.linear_wgrad(dev(dy), xd, gather=dev(idx), precision="tf32")
        _tf32_close(dw.cpu().numpy(), dy.T, X[idx], extra=1e-4)
        np.testing.assert_allclose(db.cpu().numpy(), dy.sum(0), rtol=1e-4, atol=5e-4)
        dw2, _ = F.linear_wgrad(dev(dy), xd, gather=dev(idx), precision="tf32")
        assert torch.equal(dw, dw2)
        acc = dev(np.ones((N, K), np.float32))
        F.linear_wgrad(dev(dy), xd, gather=dev(idx), dw=acc, db=dev(np.zeros(N, np.float32)), accumulate=True, precision="tf32")
        np.testing.assert_allclose(acc.cpu().numpy(), dw.cpu().numpy() + 1.0, rtol=1e-6, atol=1e-5)


@pytest.mark.parametrize("seed,n,D,ncat,NI", [(0, 96, 16, 4, 500), (1, 300, 8, 7, 500), (2, 10, 4, 6, 500), (3, 6000, 96, 40, 3000),
                                              (4, 2048, 128, 300, 5000)])
def test_category_alignment_kernel_matches_oracle(F, seed, n, D, ncat, NI):
    """csrc/catalign.cu vs the oracle's restatement of training.py:530-579 (loss and gradient), a dominant major
    category (as in the real data: almost every book's primary category is the same) and many small ones."""
    rng = np.random.default_rng(seed)
    cat = rng.integers(0, ncat, size=NI).astype(np.int64)
    cat[: NI // 2] = 0
    idx = rng.integers(0, NI, size=n).astype(np.int64)
    emb = rng.standard_normal((n, D)).astype(np.float32)
    ref_l, ref_g = oracle.category_alignment_loss(idx, emb, cat, 0)
    lam, B = 0.25, n // 3
    loss = dev(np.array([1.5, 0, 0, 0], np.float32))
    ga, gb = dev(np.ones((n, D), np.float32)), dev(np.zeros((B, D), np.float32))
    cal, _ = F.category_alignment(dev(idx), dev(emb), dev(cat), ncat, 0, lambda_c=lam, loss_out=loss, grad_a=ga, grad_b=gb, B=B)
    assert float(cal[0]) == pytest.approx(float(ref_l), rel=2e-4, abs=1e-7)
    assert float(loss[0]) == pytest.approx(1.5 + lam * float(ref_l), rel=2e-4)
    scale = max(1e-6, float(np.abs(ref_g).max()))
    np.testing.assert_allclose(ga.cpu().numpy() - 1.0, lam * ref_g, rtol=2e-3, atol=2e-4 * scale)
    np.testing.assert_allclose(gb.cpu().numpy(), lam * ref_g[:B], rtol=2e-3, atol=2e-4 * scale)
    # deterministic
    ga2 = dev(np.ones((n, D), np.float32))
    F.category_alignment(dev(idx), dev(emb), dev(cat), ncat, 0, lambda_c=lam, grad_a=ga2)
    assert torch.equal(ga, ga2)


def test_category_alignment_kernel_degenerate_cases(F):
    torch.manual_seed(1234)
    emb = torch.randn(6, 4).cuda()
    idx = torch.arange(6).cuda()
    for cats in (torch.zeros(6, dtype=torch.long), torch.tensor([0, 1, 1, 1, 2, 2])):     # one category; major has < 2 rows
        g = torch.zeros_like(emb)
        cal, _ = F.category_alignment(idx, emb, cats.cuda(), 3, 0, grad_a=g)
        assert float(cal[0]) == 0.0 and not bool(g.any())
    cal, g = F.category_alignment(idx[:0], emb[:0], torch.zeros(6, dtype=torch.long).cuda(), 3, 0)   # empty batch
    assert float(cal[0]) == 0.0 and g.shape == (0, 4)
    with pytest.raises(Exception, match="not supported"):
        F.category_alignment(torch.arange(4).cuda(), torch.randn(4, 256).cuda(), torch.zeros(4, dtype=torch.long).cuda(), 1, 0)


def test_dropout_mask_is_reproducible_and_scaled(F):
    M, N, K, p = 512, 64, 16, 0.25
    x = torch.ones(M, K, device="cuda")
    W = torch.ones(N, K, device="cuda") / K
    y1 = F.linear_fwd(x, W, None, act="relu", dropout_p=p, seed=7, offset=123)
    y2 = F.linear_fwd(x, W, None, act="relu", dropout_p=p, seed=7, offset=123)
    y3 = F.linear_fwd(x, W, None, act="relu", dropout_p=p, seed=8, offset=123)
    assert torch.equal(y1, y2) and not torch.equal(y1, y3)
    vals = torch.unique(y1).cpu().tolist()
    assert len(vals) == 2 and vals[0] == 0.0 and abs(vals[1] - 1.0 / (1 - p)) < 1e-6
    keep = float((y1 > 0).float().mean())
    assert abs(keep - (1 - p)) < 0.02
    # the stand-alone activation kernel draws the same mask as the fused epilogue
    pre = F.linear_fwd(x, W, None)
    y4 = F.act_fwd(pre, act="relu", dropout_p=p, seed=7, offset=123)
    assert torch.equal(y1, y4)
    # device step state shifts the counters
    st = F.new_step_state("cuda", 0, 0)
    F.advance_step(st, rng_stride=1 << 20)
    y5 = F.linear_fwd(x, W, None, act="relu", dropout_p=p, seed=7, offset=123, state=st)
    y6 = F.linear_fwd(x, W, None, act="relu", dropout_p=p, seed=7, offset=123 + (1 << 20))
    assert torch.equal(y5, y6) and int(st[0].item()) == 1


@pytest.mark.parametrize("act", ["relu", "gelu", "tanh", "selu"])
def test_act_bwd(F, act):
    rng = np.random.default_rng(3)
    pre = rng.standard_normal((50, 24)).astype(np.float32)
    dy = rng.standard_normal((50, 24)).astype(np.float32)
    out = oracle.model._act(act, pre)
    ref = dy * oracle.model._dact(act, pre, out)
    got = F.act_bwd(dev(dy), dev(pre), act=act)
    np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=1e-5, atol=1e-6)


def test_gate_and_augment(F):
    rng = np.random.default_rng(9)
    R, D = 300, 96
    z = rng.standard_normal((R, 2 * D)).astype(np.float32)
    pre2 = rng.standard_normal((R, D)).astype(np.float32)
    A = rng.standard_normal((50, D)).astype(np.float32)
    idx = rng.integers(0, 50, size=R).astype(np.int64)
    e, f = z[:, :D], z[:, D:]
    g_ref = oracle.model._sigmoid(pre2)
    t_ref = g_ref * e + (1 - g_ref) * f
    g, t, o, q = (torch.empty(R, D, device="cuda") for _ in range(4))
    F.gate_fwd(dev(z), dev(pre2), g=g, t=t, o=o, q=q, aug_table=dev(A), idx=dev(idx))
    np.testing.assert_allclose(g.cpu().numpy(), g_ref, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(t.cpu().numpy(), t_ref, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(o.cpu().numpy(), t_ref + A[idx], rtol=1e-5, atol=1e-6)
    assert np.array_equal(q.cpu().numpy(), A[idx])
    dt = rng.standard_normal((R, D)).astype(np.float32)
    dpre2, dz = torch.empty(R, D, device="cuda"), torch.empty(R, 2 * D, device="cuda")
    F.gate_bwd(dev(dt), dev(z), g, dpre2=dpre2, dz=dz)
    np.testing.assert_allclose(dpre2.cpu().numpy(), dt * (e - f) * g_ref * (1 - g_ref), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(dz.cpu().numpy(), np.concatenate([dt * g_ref, dt * (1 - g_ref)], 1), rtol=1e-5, atol=1e-6)
    o2, q2 = torch.empty(R, D, device="cuda"), torch.empty(R, D, device="cuda")
    F.augment_fwd(dev(t_ref), dev(A), dev(idx), out=o2, q_out=q2)
    np.testing.assert_allclose(o2.cpu().numpy(), t_ref + A[idx], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("B,N,D,mimic", [(24, 3, 16, True), (1, 1, 96, False), (8192, 5, 96, True), (100, 16, 128, True)])
def test_loss_fwd_bwd(F, B, N, D, mimic):
    rng = np.random.default_rng(B)
    o_u = rng.standard_normal((B, D)).astype(np.float32) * 0.3
    o_i = rng.standard_normal((B * (1 + N), D)).astype(np.float32) * 0.3
    kw = {}
    okw = {}
    if mimic:
        t_u, t_p, q_u, q_p = (rng.standard_normal((B, D)).astype(np.float32) * 0.1 for _ in range(4))
        kw = dict(t_u=dev(t_u), t_p=dev(t_p), q_u=dev(q_u), q_p=dev(q_p), lambda_u=0.15, lambda_i=0.2)
        okw = dict(t_u=t_u, t_p=t_p, q_u=q_u, q_p=q_p, lambda_u=0.15, lambda_i=0.2)
    ref = oracle.loss_forward_backward(o_u, o_i[:B], o_i[B:].reshape(B, N, D), **okw)
    loss, do_u, do_i, dq_u, dq_p = F.loss_fwd_bwd(dev(o_u), dev(o_i), **kw)
    loss = loss.cpu().numpy()
    assert loss[0] == pytest.approx(float(ref["loss"]), rel=2e-6)
    assert loss[1] == pytest.approx(float(ref["bce"]), rel=2e-6)
    np.testing.assert_allclose(do_u.cpu().numpy(), ref["do_u"], rtol=1e-4, atol=1e-9)
    np.testing.assert_allclose(do_i[:B].cpu().numpy(), ref["do_p"], rtol=1e-4, atol=1e-9)
    np.testing.assert_allclose(do_i[B:].cpu().numpy(), ref["do_n"].reshape(B * N, D), rtol=1e-4, atol=1e-9)
    if mimic:
        assert loss[2] == pytest.approx(float(ref["mimic_user"]), rel=2e-6)
        assert loss[3] == pytest.approx(float(ref["mimic_item"]), rel=2e-6)
        np.testing.assert_allclose(dq_u.cpu().numpy(), ref["do_u"] + ref["dq_u_extra"], rtol=1e-4, atol=1e-9)
        np.testing.assert_allclose(dq_p.cpu().numpy(), ref["do_p"] + ref["dq_p_extra"], rtol=1e-4, atol=1e-9)
    # forward only (the _compute_loss path)
    l2, *_ = F.loss_fwd_bwd(dev(o_u), dev(o_i), backward=False)
    assert float(l2[1]) == pytest.approx(float(ref["bce"]), rel=2e-6)


@pytest.mark.parametrize("B,N,D", [(24, 3, 16), (8192, 5, 96), (100, 8, 128), (7, 1, 96)])
def test_loss_with_augmentation_folded_in_equals_two_calls(F, B, N, D):
    """ttam_loss_aug_fwd_bwd (o = t + A[idx] formed inside the loss kernel) against ttam_augment_fwd + ttam_loss_fwd_bwd:
    the same fp32 operations in the same order, so every output is bit-identical (duplicate ids and an id that repeats
    between positives and negatives included)."""
    rng = np.random.default_rng(B + N + D)
    NU, NI = 50, 40                                                # small tables: many duplicate ids
    t_u = dev((rng.standard_normal((B, D)) * 0.3).astype(np.float32))
    t_i = dev((rng.standard_normal((B * (1 + N), D)) * 0.3).astype(np.float32))
    A_u = dev((rng.standard_normal((NU, D)) * 0.05).astype(np.float32))
    A_i = dev((rng.standard_normal((NI, D)) * 0.05).astype(np.float32))
    users = dev(rng.integers(0, NU, size=B).astype(np.int64))
    items = dev(rng.integers(0, NI, size=B * (1 + N)).astype(np.int64))
    o_u, q_u = torch.empty_like(t_u), torch.empty_like(t_u)
    o_i, q_i = torch.empty_like(t_i), torch.empty_like(t_i)
    F.augment_fwd(t_u, A_u, users, out=o_u, q_out=q_u)
    F.augment_fwd(t_i, A_i, items, out=o_i, q_out=q_i)
    ref = F.loss_fwd_bwd(o_u, o_i, t_u=t_u, t_p=t_i[:B], q_u=q_u, q_p=q_i[:B], lambda_u=0.15, lambda_i=0.25)
    got = F.loss_aug_fwd_bwd(t_u, t_i, A_u, A_i, users, items, mimic=True, lambda_u=0.15, lambda_i=0.25)
    for a, b, name in zip(ref, got, ("loss", "do_u", "do_i", "dq_u", "dq_p")):
        assert torch.equal(a, b), name
    # forward only
    fwd = F.loss_aug_fwd_bwd(t_u, t_i, A_u, A_i, users, items, mimic=True, lambda_u=0.15, lambda_i=0.25, backward=False)
    assert torch.equal(fwd[0], ref[0]) and fwd[1] is None


def test_sort_and_unique_rows_bit_exact(F):
    rng = np.random.default_rng(2)
    idx = rng.integers(0, 2_000_000, size=57344).astype(np.int64)
    idx[:5000] = rng.integers(0, 50, size=5000)          # heavy duplicates
    s, perm = F.sort_rows(dev(idx), 2_000_000)
    order = np.argsort(idx, kind="stable")
    assert np.array_equal(s.cpu().numpy(), idx[order])
    assert np.array_equal(perm.cpu().numpy(), order.astype(np.int32))   # stable: duplicates keep original order
    assert np.array_equal(F.unique_rows(s).cpu().numpy(), np.unique(idx))
    one = dev(np.array([7], dtype=np.int64))
    s1, p1 = F.sort_rows(one, 10)
    assert s1.cpu().tolist() == [7] and p1.cpu().tolist() == [0]


def _run_sparse_adam(F, p0, idx_steps, val_steps, lr, betas, split=False, use_state=False):
    p = dev(p0.copy())
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    scal = F.adam_scalar_table(len(idx_steps) + 1, lr, betas, "cuda")
    st = F.new_step_state("cuda") if use_state else None
    touched = []
    for s, (ix, val) in enumerate(zip(idx_steps, val_steps)):
        if st is not None:
            F.advance_step(st)
        sidx, perm = F.sort_rows(dev(ix), p0.shape[0])
        touched.append(F.unique_rows(sidx).cpu().numpy())
        if split and len(ix) > 2:
            h = len(ix) // 2
            F.sparse_adam_rows(p, m, v, sidx, perm, dev(val[:h]), dev(val[h:]), lr=lr, betas=betas, step=s + 1, scalars=scal, state=st)
        else:
            F.sparse_adam_rows(p, m, v, sidx, perm, dev(val), lr=lr, betas=betas, step=s + 1, scalars=scal, state=st)
    return p.cpu().numpy(), m.cpu().numpy(), v.cpu().numpy(), touched


@pytest.mark.parametrize("split,use_state", [(False, False), (True, False), (False, True)])
def test_sparse_adam_matches_torch_golden(F, split, use_state):
    z = np.load(GOLDEN / "optim.npz")
    steps = 6
    p, m, v, touched = _run_sparse_adam(F, z["p0"], [z[f"idx{s}"] for s in range(steps)], [z[f"val{s}"] for s in range(steps)],
                                        1e-3, (0.9, 0.999), split, use_state)
    np.testing.assert_allclose(p, z[f"sparse_adam/p{steps - 1}"], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(m, z["sparse_adam/exp_avg"], rtol=1e-6, atol=1e-10)
    np.testing.assert_allclose(v, z["sparse_adam/exp_avg_sq"], rtol=1e-6, atol=1e-12)
    for s in range(steps):
        assert np.array_equal(touched[s], np.unique(z[f"idx{s}"]))


def test_sparse_adam_wide_rows_bit_exact_vs_oracle(F):
    """D=96/128/256 rows, many duplicates: the segment sum runs in original order like coalesce, so the result
    is bit-identical to the sequential-fp32 oracle."""
    for D in (96, 128, 256, 20):
        rng = np.random.default_rng(D)
        N, R = 500, 3000
        p0 = (rng.standard_normal((N, D)) * 0.02).astype(np.float32)
        idxs = [rng.integers(0, N, size=R).astype(np.int64) for _ in range(3)]
        vals = [(rng.standard_normal((R, D)) * 0.01).astype(np.float32) for _ in range(3)]
        p, m, v, _ = _run_sparse_adam(F, p0, idxs, vals, 1e-3, (0.9, 0.999))
        po, st = p0.copy(), {"step": 0}
        for ix, val in zip(idxs, vals):
            o_sparse_adam(po, st, ix, val, lr=1e-3)
        np.testing.assert_allclose(m, st["exp_avg"], rtol=0, atol=0)
        np.testing.assert_allclose(v, st["exp_avg_sq"], rtol=0, atol=0)
        np.testing.assert_allclose(p, po, rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("D", [96, 256, 20])
def test_long_segments_block_sum_matches_oracle(F, D):
    """Zipf-like batches: a few rows own hundreds of gradient rows.  With the long-segment list those are summed by a
    whole block (different, but fixed, summation order): same touched rows, values within fp32 summation tolerance,
    and bit-identical from run to run."""
    rng = np.random.default_rng(7 + D)
    N, R = 400, 6000
    hot = rng.integers(0, 5, size=R // 2)                        # 5 rows take half of the batch
    idx = np.concatenate([hot, rng.integers(0, N, size=R - R // 2)]).astype(np.int64)
    rng.shuffle(idx)
    val = (rng.standard_normal((R, D)) * 0.01).astype(np.float32)
    p0 = (rng.standard_normal((N, D)) * 0.02).astype(np.float32)
    outs = []
    for rep in range(2):
        p = dev(p0.copy()); m, v = torch.zeros_like(p), torch.zeros_like(p)
        sidx, perm = F.sort_rows(dev(idx), N)
        ll = F.find_long_segments(sidx)
        n_long = int(ll[0])
        counts = np.bincount(idx, minlength=N)
        assert n_long == int((counts > 16).sum())
        assert set(sidx[ll[1:1 + n_long].long()].cpu().tolist()) == set(np.nonzero(counts > 16)[0].tolist())
        F.sparse_adam_rows(p, m, v, sidx, perm, dev(val), lr=1e-3, step=1, long_list=ll)
        # the lazy AdamW path on a second table, same list
        p2 = dev(p0.copy()); m2, v2 = torch.zeros_like(p2), torch.zeros_like(p2)
        last = torch.zeros(N, dtype=torch.int32, device="cuda")
        scal = F.adam_scalar_table(2, 1e-3, (0.9, 0.999), "cuda")
        F.lazy_rows("adamw", p2, m2, v2, last, sidx, perm, dev(val), scalars=scal, lr=1e-3, weight_decay=0.01, step=1,
                    long_list=ll)
        outs.append((p.cpu().numpy(), m.cpu().numpy(), p2.cpu().numpy(), last.cpu().numpy()))
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a, b)
    po, st = p0.copy(), {"step": 0}
    o_sparse_adam(po, st, idx, val, lr=1e-3)
    # a different (fixed) summation order over up to ~600 rows of magnitude 1e-2: |err(sum)| <~ 600 * 1e-2 * 2^-23
    np.testing.assert_allclose(outs[0][1], st["exp_avg"], rtol=2e-5, atol=2e-7)
    # Adam divides by sqrt(v)+eps: where the summed gradient is ~0 the update is ill-conditioned (bounded by lr)
    gsum = np.zeros_like(p0); np.add.at(gsum, idx, val)
    ok = np.abs(gsum) > 1e-4
    np.testing.assert_allclose(outs[0][0][ok], po[ok], rtol=1e-5, atol=2e-6)
    assert np.abs(outs[0][0] - po).max() <= 1.1e-3
    # lazy AdamW with long list == without
    p3 = dev(p0.copy()); m3, v3 = torch.zeros_like(p3), torch.zeros_like(p3)
    last3 = torch.zeros(N, dtype=torch.int32, device="cuda")
    sidx, perm = F.sort_rows(dev(idx), N)
    F.lazy_rows("adamw", p3, m3, v3, last3, sidx, perm, dev(val), scalars=F.adam_scalar_table(2, 1e-3, (0.9, 0.999), "cuda"),
                lr=1e-3, weight_decay=0.01, step=1)
    np.testing.assert_allclose(outs[0][2][ok], p3.cpu().numpy()[ok], rtol=1e-5, atol=2e-6)
    assert np.array_equal(outs[0][3], last3.cpu().numpy())


@pytest.mark.parametrize("D,with_list", [(96, False), (96, True), (256, True), (20, False)])
def test_row_windows_segment_shapes(F, D, with_list):
    """The window kernels (one warp per 32 sorted positions, csrc/optim.cu): segments of every length 1..70 laid end to end,
    so that heads fall on every lane, segments cross window ends, run on behind the 16 positions a window carries
    (only without a long-segment list) and the list ends inside a window.  m and v are bit-identical to the
    sequential-order oracle where the window role sums (length <= 16, or no list); p within 1e-6."""
    rng = np.random.default_rng(D + int(with_list))
    lens = list(range(1, 71)) + [1] * 37 + [16, 17, 15, 33, 32, 31, 48, 49, 47, 1, 2, 3, 5]
    rows = rng.permutation(len(lens) * 3)[: len(lens)]
    idx = np.concatenate([np.full(n, r, dtype=np.int64) for r, n in zip(rows, lens)])
    rng.shuffle(idx)
    R, N = idx.size, len(lens) * 3
    assert R % 32 != 0
    val = (rng.standard_normal((R, D)) * 0.01).astype(np.float32)
    p0 = (rng.standard_normal((N, D)) * 0.02).astype(np.float32)
    p = dev(p0.copy()); m, v = torch.zeros_like(p), torch.zeros_like(p)
    sidx, perm = F.sort_rows(dev(idx), N)
    ll = F.find_long_segments(sidx) if with_list else None
    F.sparse_adam_rows(p, m, v, sidx, perm, dev(val), lr=1e-3, step=1, long_list=ll)
    po, st = p0.copy(), {"step": 0}
    o_sparse_adam(po, st, idx, val, lr=1e-3)
    counts = np.bincount(idx, minlength=N)
    exact = (counts <= 16) if with_list else np.ones(N, bool)
    assert np.array_equal(m.cpu().numpy()[exact], st["exp_avg"][exact])
    assert np.array_equal(v.cpu().numpy()[exact], st["exp_avg_sq"][exact])
    np.testing.assert_allclose(m.cpu().numpy(), st["exp_avg"], rtol=2e-5, atol=2e-7)
    untouched = counts == 0
    assert np.array_equal(p.cpu().numpy()[untouched], p0[untouched])
    gsum = np.zeros_like(p0); np.add.at(gsum, idx, val)
    ok = np.abs(gsum) > 1e-4
    np.testing.assert_allclose(p.cpu().numpy()[ok], po[ok], rtol=1e-5, atol=2e-6)
    # the lazily-updated table through the same windows: stamps, and equality with the one-segment-at-a-time result
    p2 = dev(p0.copy()); m2, v2 = torch.zeros_like(p2), torch.zeros_like(p2)
    last = torch.zeros(N, dtype=torch.int32, device="cuda")
    scal = F.adam_scalar_table(4, 1e-3, (0.9, 0.999), "cuda")
    F.lazy_catchup("adamw", p2, m2, v2, last, sidx, scalars=scal, lr=1e-3, weight_decay=0.01, step=3)
    assert np.array_equal(last.cpu().numpy(), np.where(counts > 0, 2, 0))
    np.testing.assert_allclose(p2.cpu().numpy()[~untouched], (p0 * np.float32(1 - 1e-5) * np.float32(1 - 1e-5))[~untouched], rtol=2e-7)
    F.lazy_rows("adamw", p2, m2, v2, last, sidx, perm, dev(val), scalars=scal, lr=1e-3, weight_decay=0.01, step=3, long_list=ll)
    assert np.array_equal(last.cpu().numpy(), np.where(counts > 0, 3, 0))
    assert np.array_equal(p2.cpu().numpy()[untouched], p0[untouched])
    np.testing.assert_allclose(m2.cpu().numpy(), 0.1 * gsum, rtol=3e-5, atol=3e-8)


@pytest.mark.parametrize("kind", ["adamw", "adam", "sgd"])
def test_lazy_rows_equal_dense_optimizer(F, kind):
    """Lazy-exact replay: touching a few rows per step + a final flush == the dense optimiser stepping every row
    every step (torch golden, rows with zero gradient for several steps included)."""
    z = np.load(GOLDEN / "optim.npz")
    steps, lr, wd, mom = 6, 1e-3, 0.01, 0.9
    p = dev(z["p0"].copy())
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    last = torch.zeros(p.shape[0], dtype=torch.int32, device="cuda")
    scal = F.adam_scalar_table(steps + 1, lr, (0.9, 0.999), "cuda")
    for s in range(steps):
        sidx, perm = F.sort_rows(dev(z[f"idx{s}"]), p.shape[0])
        F.lazy_rows(kind, p, m, v, last, sidx, perm, dev(z[f"val{s}"]), scalars=scal, lr=lr, weight_decay=wd,
                    momentum=mom, step=s + 1)
        if s in (2, steps - 1):      # a mid-run flush must not change the end result
            F.lazy_flush(kind, p, m, v, last, scalars=scal, lr=lr, weight_decay=wd, momentum=mom, step=s + 1)
            np.testing.assert_allclose(p.cpu().numpy(), z[f"{kind}/p{s}"], rtol=1e-6, atol=1e-8, err_msg=f"{kind} step {s}")
    assert int(last.min()) == steps
    if kind != "sgd":
        np.testing.assert_allclose(m.cpu().numpy(), z[f"{kind}/exp_avg"], rtol=1e-5, atol=1e-10)
        np.testing.assert_allclose(v.cpu().numpy(), z[f"{kind}/exp_avg_sq"], rtol=1e-5, atol=1e-12)


@pytest.mark.parametrize("gap", [1, 7, 120, 1000])
def test_lazy_replay_against_step_by_step_fp32(F, gap):
    """Zero-gradient replay of `gap` AdamW steps (csrc/optim.cu replay(): closed form of v, MUFU reciprocal) against the
    reference's own arithmetic: the same steps one by one in fp32 (numpy float32 ops = torch CPU's, torch/optim/adam.py:
    347-547 with grad = 0).  Bound on the parameters: 5e-8 absolute (a few ulp of |p| ~ 6e-2: fused vs unfused final
    add) + 3e-6 of what the update terms moved them by in total; moments: 2e-6 (+1e-7 per step) relative - the closed
    form b2^k * v against k rounded multiplications; rows whose moments are still zero only decay.
    (fp32 step by step is itself up to 6e-7 away from the float64 recurrence at gap 1000: the rounding of p * (1 - lr*wd)
    is not random from one step to the next.)"""
    rng = np.random.default_rng(gap)
    N, D, lr, wd, b1, b2, eps = 64, 96, 1e-3, 0.01, 0.9, 0.999, 1e-8
    t0 = 5                                                       # the rows were last updated at step t0
    p0 = (rng.standard_normal((N, D)) * 0.02).astype(np.float32)
    m0 = (rng.standard_normal((N, D)) * 2e-5).astype(np.float32)
    # moments as Adam leaves them for gradients of a mean over 49k samples: |m| ~ 2e-5, sqrt(v) ~ |m| / 3
    v0 = ((m0.astype(np.float64) / 3.0) ** 2 * rng.uniform(0.5, 2.0, (N, D)) + 1e-18).astype(np.float32)
    m0[:8] = 0.0; v0[:8] = 0.0                                   # rows no gradient has reached yet
    p, m, v = dev(p0.copy()), dev(m0.copy()), dev(v0.copy())
    last = torch.full((N,), t0, dtype=torch.int32, device="cuda")
    step = t0 + gap + 1                                          # catch-up brings the rows to step - 1
    scal = F.adam_scalar_table(step + 1, lr, (b1, b2), "cuda")
    sidx, _ = F.sort_rows(torch.arange(N, device="cuda"), N)
    F.lazy_catchup("adamw", p, m, v, last, sidx, scalars=scal, lr=lr, weight_decay=wd, betas=(b1, b2), eps=eps, step=step)
    f = np.float32
    P, M, V = p0.copy(), m0.copy(), v0.copy()
    decay, c1, c2, e32 = f(1.0 - lr * wd), f(1.0 - b1), f(b2), f(eps)
    for t in range(t0 + 1, step):
        P = P * decay
        M = M + (f(0.0) - M) * c1                                # lerp_(grad = 0, 1 - beta1)
        V = V * c2
        denom = np.sqrt(V) / f(np.sqrt(1.0 - b2 ** t)) + e32
        P = P + f(-(lr / (1.0 - b1 ** t))) * (M / denom)         # addcdiv_(exp_avg, denom, value=-step_size)
    assert P.dtype == np.float32 and int(last.min()) == step - 1
    moved = np.abs(P.astype(np.float64) - p0.astype(np.float64) * float(decay) ** gap)   # what the update terms contributed
    assert moved.max() < 0.05
    err = np.abs(p.cpu().numpy().astype(np.float64) - P)
    assert (err <= 5e-8 + 3e-6 * moved).all(), (err.max(), moved.max())
    np.testing.assert_allclose(m.cpu().numpy(), M, rtol=2e-6 + 1e-7 * gap, atol=1e-36)
    np.testing.assert_allclose(v.cpu().numpy(), V, rtol=4e-6 + 2e-7 * gap, atol=1e-36)
    assert np.array_equal(m.cpu().numpy()[:8], np.zeros((8, D), np.float32))


def test_dense_step_matches_oracle(F):
    rng = np.random.default_rng(4)
    shapes = [(32, 21), (32,), (16, 32), (16,), (192, 605)]
    for kind in ("adamw", "adam", "sgd"):
        ps = [rng.standard_normal(s).astype(np.float32) * 0.1 for s in shapes]
        ref = [p.copy() for p in ps]
        sts = [{"step": 0} for _ in shapes]
        dp = [dev(p) for p in ps]
        dm = [torch.zeros_like(p) for p in dp]
        dv = [torch.zeros_like(p) for p in dp]
        for step in range(1, 4):
            gs = [rng.standard_normal(s).astype(np.float32) * 0.01 for s in shapes]
            F.dense_step(kind, dp, [dev(g) for g in gs], dm, dv, lr=1e-3, weight_decay=0.01, momentum=0.9, step=step)
            for r, g, st in zip(ref, gs, sts):
                o_dense_step(kind, r, g, st, lr=1e-3, weight_decay=0.01, momentum=0.9)
        for a, r in zip(dp, ref):
            np.testing.assert_allclose(a.cpu().numpy(), r, rtol=1e-5, atol=1e-6, err_msg=kind)   # Adam divides by sqrt(v) ~ |g|: ulps of g move p by ulps of lr


@pytest.mark.parametrize("W,R,cap", [(1, 77, 128), (2, 5000, 2688), (8, 49152, 7936), (8, 3000, 300), (4, 1024, 512)])
def test_slot_plan_pack_unpack_bit_exact(W, R, cap):
    """csrc/slots.cu against the torch-op restatement of the same layout (sharding.SlotExchange CPU path, the one the
    gloo tests run): slot layout, padding ids, flag, un-bucket (t, q, o = t + q) and re-bucket with zero padding rows."""
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import functional as F
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import sharding as S
    g = torch.Generator().manual_seed(W * 1000 + R)
    idx = torch.randint(0, 100000, (R,), generator=g)
    idx[: R // 8] = idx[0]                                       # a hot row: one bucket runs ahead of the others
    ref, dev = S.SlotExchange(R, cap, W), S.SlotExchange(R, cap, W, device="cuda")
    ref.plan(idx)
    dev.plan(idx.cuda())
    n = W * cap
    assert int(dev.flag) == int(ref.flag)
    if int(ref.flag):
        counts = torch.bincount(idx % W, minlength=W)
        assert bool((counts > cap).any()) or bool((counts == 0).any())
        fits = ref.slot_of < n                                   # what fits is laid out identically; the rest is marked
        assert torch.equal(dev.slot_of.cpu()[fits], ref.slot_of[fits]) and bool((dev.slot_of.cpu()[~fits] == n).all())
        return
    assert torch.equal(dev.send_idx[:n].cpu(), ref.send_idx[:n])
    assert torch.equal(dev.slot_of.cpu(), ref.slot_of)
    req_of = dev.req_of.cpu().long()
    held = req_of >= 0
    assert int(held.sum()) == R and torch.equal(ref.slot_of[req_of[held]], torch.nonzero(held).view(-1))
    D = 96
    t_own, q_own = torch.randn(n, D, generator=g), torch.randn(n, D, generator=g)
    tc, qc = t_own.cuda(), q_own.cuda()
    t, q, o = (torch.empty(R, D, device="cuda") for _ in range(3))
    bases = lambda x: [x.data_ptr() + w * cap * D * 4 for w in range(W)]
    F.slot_unpack(bases(tc), bases(qc), D, cap, dev.slot_of, D, t_out=t, q_out=q, o_out=o)
    assert torch.equal(t.cpu(), t_own[ref.slot_of]) and torch.equal(q.cpu(), q_own[ref.slot_of])
    assert torch.equal(o.cpu(), t_own[ref.slot_of] + q_own[ref.slot_of])
    F.slot_unpack(bases(tc), None, D, cap, dev.slot_of, D, t_out=None, q_out=None, o_out=o)
    assert torch.equal(o.cpu(), t_own[ref.slot_of])
    a, b0, n0 = torch.randn(R, D, generator=g), torch.randn(R, D, generator=g), R // 6
    ga, gb = torch.full((n, D), 7.0, device="cuda"), torch.full((n, D), 7.0, device="cuda")
    ac = a.cuda()
    F.slot_pack(ac, b0[:n0].cuda(), ac, dev.req_of, cap, bases(ga), bases(gb), D)
    want_a = torch.zeros(n + 1, D).index_copy_(0, ref.slot_of, a)[:n]
    want_b = torch.zeros(n + 1, D).index_copy_(0, ref.slot_of, torch.cat([b0[:n0], a[n0:]]))[:n]
    assert torch.equal(ga.cpu(), want_a) and torch.equal(gb.cpu(), want_b)


def test_slot_ids_equals_all_to_all_of_send_idx():
    """ttam_slot_ids with W 'requesters' emulated by W local buffers: what an equal-split all-to-all would deliver to owner
    `me` (requester w's bucket `me`), and local_rows = id // W."""
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import functional as F
    W, cap, me = 4, 96, 2
    g = torch.Generator().manual_seed(5)
    send = [torch.randint(0, 1 << 40, (W * cap,), generator=g) for _ in range(W)]        # send_idx of every requester
    dev = [s.cuda() for s in send]
    recv, rows = (torch.empty(W * cap, dtype=torch.int64, device="cuda") for _ in range(2))
    F.slot_ids([d.data_ptr() + me * cap * 8 for d in dev], cap, recv, rows)
    want = torch.cat([s.view(W, cap)[me] for s in send])
    assert torch.equal(recv.cpu(), want) and torch.equal(rows.cpu(), want // W)


# ---- bag form of layer 1 (csrc/bag.cu): the "EmbeddingBag" restatement of index_select + nn.Linear ----------------
def _sparse_features(rng, n, F, nnz_lo, nnz_hi, tail):
    """Reference feature layout (features.py:242-252): multi-hot columns with weights 1 / 0.5, then `tail` dense columns."""
    x = np.zeros((n, F), dtype=np.float32)
    for r in range(n):
        k = int(rng.integers(nnz_lo, nnz_hi + 1))
        if k:
            x[r, rng.choice(F - tail, size=k, replace=False)] = rng.choice([1.0, 0.5, 0.125], size=k)
    if tail:
        x[:, F - tail:] = rng.standard_normal((n, tail)).astype(np.float32)
    return x


@pytest.mark.parametrize("R,H,Fd,nnz,tail,act", [(1000, 192, 605, (2, 5), 5, "relu"), (8192, 192, 605, (20, 45), 5, "relu"),
                                                (49152, 192, 605, (2, 5), 5, "relu"), (77, 32, 21, (0, 3), 5, "none"),
                                                (513, 256, 418, (0, 64), 0, "relu"), (300, 512, 605, (1, 9), 8, "none"),
                                                (1, 64, 40, (3, 3), 2, "relu"), (0, 64, 40, (3, 3), 2, "relu")])
def test_bag_linear_fwd_and_wgrad_match_dense_product(F, R, H, Fd, nnz, tail, act):
    """Oracle = the dense product the reference computes (encoders.py:133 on X.index_select(0, idx)), in float64.
    fp32 FMA over the row's non-zeros: |err| <= 1e-6 * sum |x_j||w_j| (+ bias); weight gradient likewise over the rows."""
    rng = np.random.default_rng(R + H)
    NI = 3000
    X = _sparse_features(rng, NI, Fd, nnz[0], nnz[1], tail)
    W = (rng.standard_normal((H, Fd)) / np.sqrt(8)).astype(np.float32)
    b = (rng.standard_normal(H) * 0.1).astype(np.float32)
    pop = 1.0 / np.arange(1, NI + 1) ** 1.05
    idx = rng.choice(NI, size=R, p=pop / pop.sum()).astype(np.int64)       # duplicate-heavy, like the positives
    bag = F.BagMatrix.build(dev(X))
    assert bag is not None and bag.T == tail and F.bag_supported(H, Fd, bag.T, bag.max_nnz)
    y = F.bag_linear_fwd(bag, dev(idx), dev(W), dev(b), act=act).cpu().numpy()
    xr = X[idx].astype(np.float64)
    ref = xr @ W.astype(np.float64).T + b
    bound = 2e-6 * (np.abs(xr) @ np.abs(W.astype(np.float64)).T + np.abs(b)) + 1e-7
    if act == "relu":
        ref = np.maximum(ref, 0)
    assert y.shape == (R, H) and np.all(np.abs(y - ref) <= bound)
    if R == 0:
        return
    dh = (rng.standard_normal((R, H)) * 1e-3).astype(np.float32)
    dw, db = F.bag_linear_wgrad(bag, dev(idx), dev(dh))
    ref_w = dh.astype(np.float64).T @ xr
    bound_w = 4e-6 * (np.abs(dh.astype(np.float64)).T @ np.abs(xr)) + 1e-9
    assert np.all(np.abs(dw.cpu().numpy() - ref_w) <= bound_w)
    np.testing.assert_allclose(db.cpu().numpy(), dh.astype(np.float64).sum(0), rtol=1e-5, atol=1e-7)
    # accumulate, and run-to-run bit-reproducibility (no atomics: fixed ownership and summation order)
    dw2, db2 = dw.clone(), db.clone()
    F.bag_linear_wgrad(bag, dev(idx), dev(dh), dw=dw2, db=db2, accumulate=True)
    np.testing.assert_allclose(dw2.cpu().numpy(), 2 * dw.cpu().numpy(), rtol=1e-6, atol=1e-12)
    dw3, db3 = F.bag_linear_wgrad(bag, dev(idx), dev(dh))
    assert torch.equal(dw3, dw) and torch.equal(db3, db)


@pytest.mark.parametrize("R,H,Fd,nnz,tail", [(1000, 192, 605, (2, 5), 5), (49152, 192, 605, (2, 5), 5), (8192, 192, 605, (20, 45), 5),
                                             (333, 96, 200, (0, 3), 0), (4097, 256, 1100, (1, 12), 8), (31, 32, 40, (1, 2), 3)])
def test_bag_linear_wgrad_tensor_cores_match_dense_product(F, R, H, Fd, nnz, tail):
    """bag_wgrad_tc_kernel (CSR rows expanded into the tcgen05 operand tile) against the fp64 product dh^T X[idx]: TF32 operand
    rounding bound per element, exact (fp32) bias gradient, deterministic from run to run, duplicate and out-of-order ids."""
    rng = np.random.default_rng(R + H)
    NI = 3000
    X = np.zeros((NI, Fd), np.float32)
    for r in range(NI):
        k = rng.integers(nnz[0], nnz[1] + 1)
        cols = rng.choice(Fd - tail, size=min(k, Fd - tail), replace=False)
        X[r, cols] = rng.choice([1.0, 0.5, 1.0 / 3.0, 0.25], size=cols.size).astype(np.float32)
        if tail:
            X[r, Fd - tail:] = rng.standard_normal(tail).astype(np.float32)
    bag = F.BagMatrix.build(dev(X))
    assert bag is not None
    idx = dev(rng.integers(0, NI, size=R).astype(np.int64))
    dh_np = (rng.standard_normal((R, H)) * 1e-2).astype(np.float32)
    dh = dev(dh_np)
    dw, db = F.bag_linear_wgrad(bag, idx, dh, precision="tf32")
    dw2, db2 = F.bag_linear_wgrad(bag, idx, dh, precision="tf32")
    assert torch.equal(dw, dw2) and torch.equal(db, db2)
    Xg = X[idx.cpu().numpy()].astype(np.float64)
    ref = dh_np.astype(np.float64).T @ Xg
    bound = 2.0 * 2.0 ** -11 * (np.abs(dh_np.astype(np.float64)).T @ np.abs(Xg)) + 1e-7
    err = np.abs(dw.cpu().numpy().astype(np.float64) - ref)
    assert (err <= bound).all(), (err.max(), bound[err > bound][:3] if (err > bound).any() else None)
    np.testing.assert_allclose(db.cpu().numpy(), dh_np.astype(np.float64).sum(0), rtol=2e-5, atol=1e-6)
    # accumulate into an existing gradient
    base = torch.ones_like(dw)
    dw3, _ = F.bag_linear_wgrad(bag, idx, dh, dw=base, db=torch.zeros_like(db), accumulate=True, precision="tf32")
    np.testing.assert_allclose(dw3.cpu().numpy(), dw.cpu().numpy() + 1.0, rtol=0, atol=2e-7)


def test_bag_linear_fwd_dropout_mask_equals_gemm_path(F):
    """Same Philox element numbering as ttam_linear_fwd: the two layer-1 paths drop the same elements for the same
    (seed, offset), so the backward mask (h > 0) of either is valid for both."""
    rng = np.random.default_rng(5)
    NI, R, H, Fd = 500, 300, 64, 40
    X = _sparse_features(rng, NI, Fd, 2, 4, 2)
    W = (np.abs(rng.standard_normal((H, Fd))) + 0.1).astype(np.float32)     # positive pre-activations: zeros are drops
    X = np.abs(X)
    b = np.ones(H, dtype=np.float32)
    idx = rng.integers(0, NI, size=R).astype(np.int64)
    bag = F.BagMatrix.build(dev(X))
    a = F.bag_linear_fwd(bag, dev(idx), dev(W), dev(b), act="relu", dropout_p=0.3, seed=99, offset=1 << 20)
    g = F.linear_fwd(dev(X), dev(W), dev(b), gather=dev(idx), act="relu", dropout_p=0.3, seed=99, offset=1 << 20)
    assert torch.equal(a == 0, g == 0) and 0.2 < float((a == 0).float().mean()) < 0.4
    np.testing.assert_allclose(a.cpu().numpy(), g.cpu().numpy(), rtol=2e-5, atol=1e-6)


def test_bag_linear_fwd_tf32_rounded_output(F):
    rng = np.random.default_rng(6)
    X = _sparse_features(rng, 200, 40, 2, 4, 2)
    W = rng.standard_normal((64, 40)).astype(np.float32)
    bag = F.BagMatrix.build(dev(X))
    idx = dev(np.arange(200, dtype=np.int64))
    y = F.bag_linear_fwd(bag, idx, dev(W), None)
    yr = F.bag_linear_fwd(bag, idx, dev(W), None, round_tf32_out=True)
    assert torch.equal(yr, F.round_tf32_(y.clone()))


# ---- TMA-fed persistent TF32 GEMM (csrc/gemm_tma.cu): forward with a prepared weight, dgrad with its transposed copy ----
@pytest.mark.parametrize("M,N,K", [(49152, 96, 192), (130, 96, 192), (257, 96, 96), (64, 16, 32), (4099, 256, 64), (128, 300, 128),
                                   (1000, 512, 256), (5, 96, 100), (20000, 192, 96)])
def test_tma_gemm_forward_matches_cp_async_kernel_bit_for_bit(F, M, N, K):
    """Same operands, same TF32 rounding, same K order of the tcgen05.mma: the TMA-fed kernel must reproduce the cp.async
    kernel's bits (and with them its error bound against the fp64 product); ReLU, bias, rounded output, pre-rounded A."""
    rng = np.random.default_rng(M + N + K)
    x = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    b = (rng.standard_normal(N) * 0.1).astype(np.float32)
    (Wr,), (WrT,) = F.prepare_weights([dev(W)])
    assert torch.equal(Wr, F.round_tf32_(dev(W))) and torch.equal(WrT, Wr.t().contiguous())
    ref = F.linear_fwd(dev(x), dev(W), dev(b), act="relu", precision="tf32")                      # cp.async kernel (weight not prepared)
    got = F.linear_fwd(dev(x), Wr, dev(b), act="relu", precision="tf32", w_rounded=True)          # TMA kernel
    assert torch.equal(got, ref)
    _tf32_close(F.linear_fwd(dev(x), Wr, dev(b), precision="tf32", w_rounded=True).cpu().numpy() - b, x, W.T)
    xr = F.round_tf32_(dev(x))
    assert torch.equal(F.linear_fwd(xr, Wr, dev(b), act="relu", precision="tf32", w_rounded=True, x_rounded=True), ref)
    got_r = F.linear_fwd(dev(x), Wr, dev(b), act="relu", precision="tf32", w_rounded=True, out_rounded=True)
    assert torch.equal(got_r, F.round_tf32_(ref.clone()))
    # strided output view (z[:, D:]) and strided input view
    z = torch.zeros((M, N + 32), device="cuda")
    F.linear_fwd(dev(x), Wr, dev(b), act="relu", out=z[:, 32:], precision="tf32", w_rounded=True)
    assert torch.equal(z[:, 32:], ref) and float(z[:, :32].abs().sum()) == 0.0


@pytest.mark.parametrize("M,N,K", [(49152, 96, 192), (4096, 96, 96), (3000, 192, 96), (700, 96, 96), (999, 48, 36), (333, 512, 256)])
def test_tma_gemm_dgrad_with_transposed_weight(F, M, N, K):
    rng = np.random.default_rng(M + 3 * N + K)
    dy = rng.standard_normal((M, N)).astype(np.float32)
    W = rng.standard_normal((N, K)).astype(np.float32)
    aux = rng.standard_normal((M, K)).astype(np.float32)
    (_,), (WrT,) = F.prepare_weights([dev(W)])
    ref = F.linear_dgrad(dev(dy), dev(W), precision="tf32")
    got = F.linear_dgrad(dev(dy), WrT, precision="tf32", w_transposed=True)
    assert torch.equal(got, ref)
    _tf32_close(got.cpu().numpy(), dy, W)
    got_m = F.linear_dgrad(dev(dy), WrT, aux=dev(aux), relu_mask=True, scale=1.25, precision="tf32", w_transposed=True)
    assert torch.equal(got_m, F.linear_dgrad(dev(dy), dev(W), aux=dev(aux), relu_mask=True, scale=1.25, precision="tf32"))
    base = rng.standard_normal((M, K)).astype(np.float32)
    out, out_ref = dev(base), dev(base)
    F.linear_dgrad(dev(dy), WrT, out=out, accumulate=True, precision="tf32", w_transposed=True)
    F.linear_dgrad(dev(dy), dev(W), out=out_ref, accumulate=True, precision="tf32")
    assert torch.equal(out, out_ref)
    got_r = F.linear_dgrad(dev(dy), WrT, precision="tf32", w_transposed=True, out_rounded=True)
    assert torch.equal(got_r, F.round_tf32_(ref.clone()))


@pytest.mark.parametrize("M,N,K", [(49152, 96, 96), (49152, 96, 192), (8192, 96, 192), (4096, 192, 96), (999, 48, 36), (3000, 192, 608),
                                   (700, 512, 256), (33, 96, 96), (20000, 256, 512)])
def test_tma_wgrad_matches_cp_async_kernel(F, M, N, K, monkeypatch):
    """dw = dy^T x and db = colsum(dy) from the TMA-fed MN-major kernel against the cp.async kernel (same TF32 rounding; the split
    of the row range differs, so the fp32 partial sums are combined in a different order) and against the fp64 product."""
    rng = np.random.default_rng(M + N + K)
    dy = rng.standard_normal((M, N)).astype(np.float32)
    x = rng.standard_normal((M, K)).astype(np.float32)
    monkeypatch.setenv("TTAM_NO_TMA_WGRAD", "1")
    dw_ref, db_ref = F.linear_wgrad(dev(dy), dev(x), precision="tf32")
    monkeypatch.delenv("TTAM_NO_TMA_WGRAD")
    dw, db = F.linear_wgrad(dev(dy), dev(x), precision="tf32")
    _tf32_close(dw.cpu().numpy(), dy.T, x, extra=1e-4)
    np.testing.assert_allclose(dw.cpu().numpy(), dw_ref.cpu().numpy(), rtol=1e-4, atol=2e-3 * np.sqrt(M / 1000))
    np.testing.assert_allclose(db.cpu().numpy(), dy.astype(np.float64).sum(0), rtol=1e-4, atol=5e-4 * np.sqrt(M / 1000))
    np.testing.assert_allclose(db.cpu().numpy(), db_ref.cpu().numpy(), rtol=1e-5, atol=1e-3)
    dw2, db2 = F.linear_wgrad(dev(dy), dev(x), precision="tf32")
    assert torch.equal(dw, dw2) and torch.equal(db, db2)                      # deterministic
    xr = F.round_tf32_(dev(x))
    dw3, _ = F.linear_wgrad(dev(dy), xr, precision="tf32", x_rounded=True)
    assert torch.equal(dw3, dw)
    acc = dev(np.ones((N, K), np.float32))
    F.linear_wgrad(dev(dy), dev(x), dw=acc, db=dev(np.zeros(N, np.float32)), accumulate=True, precision="tf32")
    np.testing.assert_allclose(acc.cpu().numpy(), dw.cpu().numpy() + 1.0, rtol=1e-6, atol=1e-5)
