"""The closed form behind the zero-gradient replay of the lazily-updated tables (csrc/optim.cu replay()), on CPU:
with grad = 0 AdamW gives v_t = b2^(t-t0) v, so sqrt(v_t)/sqrt(bc2_t) = [sqrt(v)/b2^(t0/2)] * h_t with the per-step scalar
h_t = b2^(t/2)/sqrt(1-b2^t) that `functional.adam_scalar_table` stores at [4t+3].  This restates the kernel's loop in numpy
fp32 and compares it with the reference's step-by-step arithmetic (torch/optim/adam.py:347-547 with a zero gradient);
the GPU kernel itself is held to the same bound in tests/test_gpu_kernels.py::test_lazy_replay_against_step_by_step_fp32."""
import numpy as np
import pytest

from two_tower_augmented_with_adaptive_mimic_mechanism_b200.functional import adam_scalar_table

f = np.float32


def _step_by_step(p, m, v, t0, t1, lr, wd, b1, b2, eps):
    p, m, v = p.copy(), m.copy(), v.copy()
    for t in range(t0 + 1, t1 + 1):
        p = p * f(1.0 - lr * wd)
        m = m + (f(0.0) - m) * f(1.0 - b1)
        v = v * f(b2)
        p = p + f(-(lr / (1.0 - b1 ** t))) * (m / (np.sqrt(v) / f(np.sqrt(1.0 - b2 ** t)) + f(eps)))
    return p, m, v


def _closed_form(p, m, v, t0, t1, tab, lr, wd, b1, eps):
    """The kernel's loop: scalars a_t = tab[4t], sqrt(bc2_t) = tab[4t+1], h_t = tab[4t+3]."""
    p, m = p.copy(), m.copy()
    g_from = f(1.0) if t0 == 0 else f(tab[4 * t0 + 3] * tab[4 * t0 + 1])
    ratio = f(f(tab[4 * t1 + 3] * tab[4 * t1 + 1]) / g_from)
    w0 = (np.sqrt(v) / g_from).astype(f)
    for t in range(t0 + 1, t1 + 1):
        p = p * f(1.0 - lr * wd)
        m = m + (f(0.0) - m) * f(1.0 - b1)
        p = p + f(-tab[4 * t]) * (m / (w0 * f(tab[4 * t + 3]) + f(eps)))
    return p, m, (v * ratio * ratio).astype(f)


@pytest.mark.parametrize("t0,gap", [(0, 3), (5, 1), (5, 120), (40, 1000)])
def test_closed_form_replay_equals_step_by_step(t0, gap):
    rng = np.random.default_rng(t0 + gap)
    lr, wd, b1, b2, eps = 1e-3, 0.01, 0.9, 0.999, 1e-8
    p0 = (rng.standard_normal((32, 96)) * 0.02).astype(f)
    m0 = (rng.standard_normal((32, 96)) * 2e-5).astype(f)
    v0 = ((m0.astype(np.float64) / 3.0) ** 2 * rng.uniform(0.5, 2.0, m0.shape) + 1e-18).astype(f)
    if t0 == 0:
        m0[:], v0[:] = 0.0, 0.0                                  # nothing has been applied at step 0: moments are zero
    t1 = t0 + gap
    tab = adam_scalar_table(t1 + 1, lr, (b1, b2), "cpu").numpy()
    P, M, V = _step_by_step(p0, m0, v0, t0, t1, lr, wd, b1, b2, eps)
    p, m, v = _closed_form(p0, m0, v0, t0, t1, tab, lr, wd, b1, eps)
    assert p.dtype == f and P.dtype == f
    moved = np.abs(P.astype(np.float64) - p0.astype(np.float64) * float(f(1.0 - lr * wd)) ** gap)
    assert (np.abs(p.astype(np.float64) - P) <= 5e-8 + 3e-6 * moved).all()
    np.testing.assert_array_equal(m, M)                          # the same fp32 recurrence
    np.testing.assert_allclose(v, V, rtol=4e-6 + 2e-7 * gap, atol=1e-36)


def test_scalar_table_columns():
    lr, b1, b2 = 1e-3, 0.9, 0.999
    tab = adam_scalar_table(50, lr, (b1, b2), "cpu").numpy().astype(np.float64).reshape(-1, 4)
    t = np.arange(1, 51)
    np.testing.assert_allclose(tab[1:, 0], lr / (1 - b1 ** t), rtol=1e-6)
    np.testing.assert_allclose(tab[1:, 1], np.sqrt(1 - b2 ** t), rtol=1e-6)
    np.testing.assert_allclose(tab[1:, 2], lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t), rtol=1e-6)
    b2f = float(np.float32(b2))
    np.testing.assert_allclose(tab[1:, 3], b2f ** (t / 2) / np.sqrt(1 - b2 ** t), rtol=1e-6)
    np.testing.assert_allclose(tab[1:, 3] * tab[1:, 1], b2f ** (t / 2), rtol=1e-6)      # how the kernel recovers b2^(t/2)
