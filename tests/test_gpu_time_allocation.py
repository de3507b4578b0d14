"""GPU tier: batched time-allocation search (BASELINE configs[2]) against its
numpy checker and against the properties of the method.  The reference has no such search —
parity is between the two implementations only (UNPINNED, see oracle/minsnap_oracle.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _problems(rng, B, n, K):
    wp = np.cumsum(rng.normal(0, 1.0, (B, n + 1, K)), axis=1)
    T = rng.uniform(0.6, 1.6, (B, n))
    t = np.concatenate([np.zeros((B, 1)), np.cumsum(T, axis=1)], axis=1)
    return wp, t


@pytest.mark.parametrize("n,K", [(3, 3), (6, 4), (10, 3)])
def test_matches_the_numpy_checker(n, K):
    from oracle import minsnap_oracle as mo
    import drone_path_planning_python_b200 as mst
    rng = np.random.default_rng(40 + n)
    B, iters = 6, 4
    wp, t = _problems(rng, B, n, K)
    t_new, cost = mst.optimize_time_allocation(wp, t, iters=iters)
    t_new, cost = t_new.cpu().numpy(), cost.cpu().numpy()
    for b in range(B):
        want_t, want_cost = mo.optimize_time_allocation(wp[b], t[b], iters=iters)
        np.testing.assert_allclose(cost[:, b], want_cost, rtol=1e-7)
        np.testing.assert_allclose(t_new[b], want_t, rtol=0, atol=1e-6 * t[b, -1])


def test_properties_on_a_large_batch():
    import drone_path_planning_python_b200 as mst
    rng = np.random.default_rng(7)
    B, n, K = 3000, 10, 3
    wp, t = _problems(rng, B, n, K)
    t_new, cost = mst.optimize_time_allocation(wp, t, iters=5)
    assert cost.shape == (6, B) and bool(torch.isfinite(cost).all())
    assert bool((cost[1:] <= cost[:-1]).all())                  # never accepts a worse allocation
    assert float((cost[-1] / cost[0]).median()) < 0.8           # and does find better ones
    t_new = t_new.cpu().numpy()
    assert np.array_equal(t_new[:, 0], t[:, 0]) and np.array_equal(t_new[:, -1], t[:, -1])
    T = np.diff(t_new, axis=1)
    assert (T >= 0.1 * t[:, -1:] / n * (1 - 1e-12)).all()       # the duration floor
    # the reported cost is the cost of the returned stamps
    coef, dur, info = mst.solve_batch(wp, t_new)
    assert int((info != 0).sum()) == 0
    np.testing.assert_allclose(mst.snap_cost(coef, dur).cpu().numpy(), cost[-1].cpu().numpy(), rtol=1e-9)
    # problems do not influence each other
    sub, cost_sub = mst.optimize_time_allocation(wp[5:9], t[5:9], iters=5)
    assert np.array_equal(sub.cpu().numpy(), t_new[5:9])


def test_uniform_stamps_are_a_fixed_point_for_evenly_spaced_collinear_waypoints():
    import drone_path_planning_python_b200 as mst
    n, K = 6, 3
    wp = np.linspace(0, 1, n + 1)[None, :, None] * np.array([1.0, 2.0, -1.0])[None, None, :]
    t = np.linspace(0, 6, n + 1)[None]
    skew = t.copy()
    skew[0, 1:-1] += np.array([0.3, -0.2, 0.25, -0.3, 0.1])
    t_new, cost = mst.optimize_time_allocation(np.repeat(wp, 2, 0), np.concatenate([t, skew]), iters=12)
    cost = cost.cpu().numpy()
    assert cost[-1, 0] <= cost[0, 0]
    # by symmetry the optimum is symmetric about the middle; from skewed stamps the search moves towards it
    T1 = np.diff(t_new[1].cpu().numpy())
    T0 = np.diff(skew[0])
    assert np.abs(T1 - T1[::-1]).max() < np.abs(T0 - T0[::-1]).max()
    assert cost[-1, 1] < cost[0, 1]


def test_degenerate_sizes():
    import drone_path_planning_python_b200 as mst
    wp = np.array([[[0.0, 0, 0], [1, 1, 1]]])
    t = np.array([[0.0, 2.0]])
    t_new, cost = mst.optimize_time_allocation(wp, t, iters=3)       # one piece: nothing to move
    assert np.array_equal(t_new.cpu().numpy(), t) and cost.shape == (4, 1)
    t_new, cost = mst.optimize_time_allocation(np.zeros((0, 5, 3)), np.zeros((0, 5)), iters=2)
    assert t_new.shape == (0, 5) and cost.shape == (3, 0)


def test_time_gradient_matches_the_checker_and_central_differences():
    from oracle import minsnap_oracle as mo
    import drone_path_planning_python_b200 as mst
    rng = np.random.default_rng(5)
    B, n, K = 40, 8, 4
    wp, t = _problems(rng, B, n, K)
    coef, dur, info = mst.solve_batch(wp, t)
    grad = mst.time_gradient(coef).cpu().numpy()
    want = np.stack([mo.time_gradient(c) for c in coef.cpu().numpy()])
    np.testing.assert_allclose(grad, want, rtol=1e-12, atol=1e-9 * np.abs(want).max())
    # and it is the derivative of the cost the solver + cost kernels produce
    T = np.diff(t, axis=1)
    h = 1e-6
    for i in (0, n // 2, n - 1):
        Tp, Tm = T.copy(), T.copy()
        Tp[:, i] += h
        Tm[:, i] -= h
        Jp = mst.snap_cost(*mst.solve_batch(wp, np.concatenate([t[:, :1], t[:, :1] + np.cumsum(Tp, 1)], 1))[:2])
        Jm = mst.snap_cost(*mst.solve_batch(wp, np.concatenate([t[:, :1], t[:, :1] + np.cumsum(Tm, 1)], 1))[:2])
        fd = ((Jp - Jm) / (2 * h)).cpu().numpy()
        np.testing.assert_allclose(grad[:, i], fd, rtol=1e-5, atol=1e-5 * np.abs(fd).max())
