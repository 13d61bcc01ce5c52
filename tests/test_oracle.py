"""The oracle (oracle/insider_oracle.cpp) against the committed golden vectors (produced by the independently written
NumPy/SciPy twin, tests/golden/make_golden.py), analytic optimality conditions, and the reference's edge cases."""
import os
import sys

import numpy as np
import pytest

from oracle import numpy_twin, oracle

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import make_golden  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", sorted(make_golden.CASES))
def test_oracle_matches_golden(name):
    N, P, K, tuning, alpha, lam, Q, iters = make_golden.CASES[name]
    pb, tr, te, F0, V0 = make_golden.make_inputs(name)
    g = np.load(os.path.join(GOLD, f"{name}.npz"))
    r = oracle.optimize(pb.Y, F0, V0, pb.confounder, pb.X, tr, te, 1 if Q else 0, K, lam, lam, alpha, tuning, 1e-12, 1e-5, iters, perm_mode=1, seed=77)
    assert r.iters_run == int(g["iters_run"])
    assert r.cd_sweeps == int(g["cd_sweeps"])
    np.testing.assert_allclose(r.column_factor, g["V"], rtol=0, atol=1e-11 * np.abs(g["V"]).max())
    for i, f in enumerate(r.factors):
        np.testing.assert_allclose(f, g[f"F{i}"], rtol=0, atol=1e-11 * np.abs(g[f"F{i}"]).max())
    np.testing.assert_allclose(r.loss, float(g["loss"]), rtol=1e-13)
    np.testing.assert_allclose([c["loss"] for c in r.checks], g["check_loss"], rtol=1e-13)
    if tuning == 1:
        np.testing.assert_allclose(r.test_rmse, float(g["test_rmse"]), rtol=1e-13)
    else:
        assert np.isnan(r.test_rmse)                   # reference leaves it unset (src/utils.cpp:61-63)


def _cd_problem(seed, n=60, K=7):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, K))
    beta = rng.normal(size=K) * (rng.random(K) < 0.6)
    y = X @ beta + 0.3 * rng.normal(size=n)
    return X, y, X.T @ X, X.T @ y


@pytest.mark.parametrize("alpha", [0.2, 0.5, 1.0])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_strong_cd_satisfies_kkt(alpha, mode):
    """Elastic-net KKT: X'(y - Xb) - l2 b = la*sign(b) on the support, |.| <= la off it; every permutation mode (R-stream,
    counter, identity) reaches the same minimiser."""
    lam = 3.0
    X, y, G, b = _cd_problem(11)
    beta, sweeps, rounds = oracle.strong_cd(X, y, np.zeros(7), lam, alpha, G, b, tol=1e-14, perm_mode=mode, seed=3, r_seed=5)
    la, l2 = lam * alpha, lam * (1 - alpha)
    grad = X.T @ (y - X @ beta) - l2 * beta
    on = beta != 0
    np.testing.assert_allclose(grad[on], la * np.sign(beta[on]), atol=2e-6)
    assert np.all(np.abs(grad[~on]) <= la + 2e-6)
    ref, _, _ = oracle.strong_cd(X, y, np.zeros(7), lam, alpha, G, b, tol=1e-14, perm_mode=2)
    np.testing.assert_allclose(beta, ref, atol=1e-6)


def test_strong_cd_cpp_vs_twin():
    X, y, G, b = _cd_problem(5)
    from oracle.r_rng import randperm_b
    for alpha in (0.3, 1.0):
        bo, sw, _ = oracle.strong_cd(X, y, 0.01 * np.ones(7), 2.0, alpha, G, b, tol=1e-7, perm_mode=1, seed=9, als_iter=4, gene=12)
        st, draw = {}, [0]

        def perm(inc_idx, K):
            d = draw[0]; draw[0] += 1
            return randperm_b(9, 4, 12, d, K, inc_idx)
        bt = numpy_twin.strong_coordinate_descent(X, y, 0.01 * np.ones(7), 2.0, alpha, G, b, 1e-7, perm, st)
        np.testing.assert_allclose(bo, bt, atol=1e-13)
        assert sw == st["sweeps"]


def test_screening_excludes_and_kkt_readmits():
    """coordinate_descent.cpp:74-78 zeroes screened coordinates of the warm start; :118-124 re-admits violators."""
    rng = np.random.default_rng(2)
    X = rng.normal(size=(80, 5))
    y = 0.5 * X[:, 0] + 0.4 * X[:, 1] + 0.1 * rng.normal(size=80)
    G, b = X.T @ X, X.T @ y
    lam, alpha = 30.0, 0.9
    assert np.any(np.abs(b) < alpha * (2 * lam - np.abs(b).max()))       # something is screened
    beta, _, rounds = oracle.strong_cd(X, y, np.ones(5), lam, alpha, G, b, tol=1e-12, perm_mode=2)
    la, l2 = lam * alpha, lam * (1 - alpha)
    grad = X.T @ (y - X @ beta) - l2 * beta
    assert np.all(np.abs(grad[beta == 0]) <= la + 1e-6)
    assert rounds >= 1


def test_loss_monotone_and_row_normal_equations():
    name = "masked_cd"
    N, P, K, tuning, alpha, lam, Q, iters = make_golden.CASES[name]
    pb, tr, te, F0, V0 = make_golden.make_inputs(name)
    r = oracle.optimize(pb.Y, F0, V0, pb.confounder, None, tr, te, 0, K, lam, lam, alpha, 1, 1e-12, 1e-9, 60, perm_mode=1, seed=1)
    losses = [c["loss"] for c in r.checks]
    assert all(b <= a + 1e-9 * abs(a) for a, b in zip(losses, losses[1:]))
    assert [c["iter"] for c in r.checks] == [-1, 0, 10, 20, 30, 40, 50, 60]            # src/optimize.cpp:322,381
    assert r.iters_run == 61                                                             # while (iter <= max_iter)


def test_decay_ladder_and_convergence_break():
    pb, tr, te, F0, V0 = make_golden.make_inputs("dense_cd")
    r = oracle.optimize(pb.Y, F0, V0, pb.confounder, None, tr, te, 0, 5, 2.0, 2.0, 0.4, 0, 1e-4, 1e-5, 500, perm_mode=1, seed=1)
    assert r.iters_run < 500 and r.iters_run % 10 == 0                                   # break happens on a check iteration, iter not incremented
    last = r.checks[-1]
    prev = r.checks[-2]
    assert (prev["loss"] - last["loss"]) / prev["loss"] < 1e-4
    for c in r.checks[1:]:
        d = c["delta_loss"] / 1000
        expect = next((t for t in (1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1e-1) if d <= t), 1.0)  # src/optimize.cpp:389-403
        assert c["decay"] == expect


def test_invalid_arguments_are_rejected():
    pb, tr, te, F0, V0 = make_golden.make_inputs("dense_cd")
    with pytest.raises(RuntimeError):
        oracle.optimize(pb.Y, F0, V0, pb.confounder, None, tr, te, 0, 5, 1, 1, 0.1, 2, 1e-9, 1e-5, 3)      # tuning not in {0,1}
    with pytest.raises(RuntimeError):
        oracle.optimize(pb.Y, F0, V0, pb.confounder, None, tr, te, 3, 5, 1, 1, 0.1, 1, 1e-9, 1e-5, 3)      # inc_continuous
    bad = pb.confounder.copy(); bad[:, 0] += 1                                                             # levels not 1..L
    with pytest.raises(RuntimeError):
        oracle.optimize(pb.Y, F0, V0, bad, None, tr, te, 0, 5, 1, 1, 0.1, 1, 1e-9, 1e-5, 3)
    with pytest.raises(RuntimeError):                                                                      # empty test set, tuning=1
        oracle.optimize(pb.Y, F0, V0, pb.confounder, None, tr, np.zeros_like(te), 0, 5, 1, 1, 0.1, 1, 1e-9, 1e-5, 3)


def test_fit_interaction_math():
    rng = np.random.default_rng(3)
    N, P, K, L = 30, 40, 4, 5
    V = rng.normal(size=(K, P)); R = rng.normal(size=(N, P))
    z = np.concatenate([np.arange(1, L + 1), rng.integers(1, L + 1, N - L)]).astype(np.int32)
    tr = (rng.random((N, P)) < 0.9).astype(np.int32)
    for tuning in (0, 1):
        out = oracle.fit_interaction(R, tr, L, z, V, tuning)
        for s in range(1, L + 1):
            rows = np.flatnonzero(z == s)
            XtX = np.zeros((K, K)); Xty = np.zeros(K)
            for k in rows:
                m = tr[k] != 0 if tuning == 1 else np.ones(P, bool)
                XtX += V[:, m] @ V[:, m].T; Xty += V[:, m] @ R[k, m]
            np.testing.assert_allclose(out[s - 1], np.linalg.solve(XtX, Xty), rtol=1e-10)


@pytest.mark.parametrize("case", range(12))
def test_oracle_matches_twin_on_random_small_shapes(case):
    """C++ oracle == NumPy/SciPy twin on seeded random small problems - N in [5, 64], P in [3, 50], K in [1, 8],
    1-3 confounders, masked / dense, alpha in {0, 0.4, 1}, with and without a continuous covariate - factors to 1e-11, identical
    sweep and iteration counts."""
    from insider_b200 import synth
    rng = np.random.default_rng(500 + case)
    K = int(rng.integers(1, 9))
    C = int(rng.integers(1, 4))
    levels = tuple(int(v) for v in rng.integers(1, 5, size=C))
    N = int(rng.integers(max(5, max(levels) + 1), 65))
    P = int(rng.integers(3, 51))
    Q = int(case % 4 == 3)
    tuning, alpha = case % 2, (0.0, 0.4, 1.0)[case % 3]
    lam = float(rng.choice([0.5, 2.0, 8.0]))
    pb = synth.with_continuous(N=N, P=P, K=K, levels=levels, Q=max(Q, 1), seed=case)
    X = pb.X if Q else None
    tr, te = synth.random_masks(N, P, 0.15, case + 1)
    F0, V0 = synth.init_factors(pb.levels, K, P, Q=Q, seed=case + 2)
    args = (pb.Y, F0, V0, pb.confounder, X, tr, te, Q, K, lam, lam, alpha, tuning, 1e-12, 1e-5, 10)
    r = oracle.optimize(*args, perm_mode=1, seed=31)
    t = numpy_twin.optimize(*args, perm_mode=1, seed=31)
    assert (r.iters_run, r.cd_sweeps) == (t["iters_run"], t["cd_sweeps"]), (N, P, K, levels, Q, tuning, alpha)
    np.testing.assert_allclose(r.column_factor, t["column_factor"], rtol=0, atol=1e-11 * max(1e-3, np.abs(t["column_factor"]).max()))
    for a, b in zip(r.factors, t["row_matrices"]):
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-11 * max(1e-3, np.abs(b).max()))
    np.testing.assert_allclose(r.loss, t["loss"], rtol=1e-12)
