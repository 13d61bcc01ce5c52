"""CPU model of the arithmetic the CUDA elastic-net solvers run (k_cd_dense.cu / k_cd.cu), against the residual-form oracle.

The reference (src/coordinate_descent.cpp:57-127) keeps an explicit residual and stops a round of sweeps on the difference of two
full loss evaluations. The kernels never see X or y: their state is p_k = (X'y - X'X beta)_k + beta_k XtX_kk (the reference's `upper`,
:94), an update multiplies by a tabulated 1 / (XtX_kk + lambda (1 - alpha)), and the stopping quantity is the sum of exact per-update
loss decrements. This file states that arithmetic in NumPy, operation for operation like cd_step() of k_cd_dense.cu, and checks on
seeded random problems that it takes exactly as many sweeps as the residual-form oracle and lands on the same coefficients - the
CPU-side counterpart of the sweep-count assertions of the -m gpu tests."""
import numpy as np
import pytest

from oracle import oracle
from oracle.r_rng import randperm_b


def cd_pform(G, b, w0, lam, alpha, tol, seed, als_iter):
    """One column solve in the kernels' form. Returns (beta, sweeps, KKT rounds)."""
    K = b.size
    la, l2 = lam * alpha, lam * (1.0 - alpha)
    active = ~(np.abs(b) < alpha * (2.0 * lam - np.max(np.abs(b))))            # strong-rule screen, coordinate_descent.cpp:74-78
    beta = np.where(active, w0, 0.0)
    d = np.diag(G).copy()
    inv, hden = 1.0 / (d + l2), 0.5 * (d + l2)                                 # k_cd_table: per-row constants
    p = (b - G @ beta) + beta * d                                              # the state: `upper` of every coordinate
    sweeps = rounds = 0
    full_order = {}
    while True:
        rounds += 1
        inc = np.flatnonzero(active)
        while True:
            if sweeps not in full_order:                                       # one order per (seed, ALS iteration, sweep index), shared by all genes
                full_order[sweeps] = randperm_b(seed, als_iter, 0, sweeps, K)
            dl = 0.0
            for k in full_order[sweeps]:
                if not active[k]:
                    continue
                up = p[k]
                t1 = abs(up) - la
                nb = np.copysign(t1, up) * inv[k] if t1 > 0.0 else 0.0         # :99-104 with the reciprocal multiply
                bo = beta[k]
                dlt = nb - bo
                dl += dlt * (hden[k] * (nb + bo) - up) + la * (abs(nb) - abs(bo))   # exact loss decrement of this update
                beta[k] = nb
                pk = p[k]
                p -= dlt * G[:, k]                                             # p_l -= delta XtX_lk (l != k); p_k itself does not move
                p[k] = pk
            sweeps += 1
            if not abs(dl) > tol:                                              # :114 |pre_loss - loss| > tol
                break
        viol = ~active & (np.abs(p) > la)                                      # :118-124 (beta_e = 0, so p_e = q_e = -gradient)
        if not viol.any():
            return beta, sweeps, rounds
        active |= viol


def _problem(rng, n, K, support):
    X = rng.normal(size=(n, K)) * rng.uniform(0.3, 2.0, size=K)
    X[:, 1:] += 0.4 * X[:, :1]                                                  # correlated columns: many sweeps
    beta = rng.normal(size=K) * (rng.random(K) < support)
    y = X @ beta + 0.5 * rng.normal(size=n)
    return X, y, X.T @ X, X.T @ y


@pytest.mark.parametrize("case", range(24))
def test_pform_model_takes_the_oracles_sweeps(case):
    rng = np.random.default_rng(1000 + case)
    K = int(rng.integers(2, 25))
    n = int(rng.integers(K + 5, 120))
    X, y, G, b = _problem(rng, n, K, support=rng.uniform(0.2, 0.9))
    lam = float(rng.choice([0.5, 3.0, 10.0, 40.0]))
    alpha = float(rng.choice([0.1, 0.4, 0.7, 1.0]))
    tol = float(rng.choice([1e-5, 1e-7, 1e-9]))
    w0 = 0.001 * rng.normal(size=K) if case % 3 else rng.normal(size=K)        # cold start / warm start
    seed, it = 17 + case, case % 7
    bo, sw, rd = oracle.strong_cd(X, y, w0, lam, alpha, G, b, tol=tol, perm_mode=1, seed=seed, als_iter=it)
    bm, swm, rdm = cd_pform(G, b, w0, lam, alpha, tol, seed, it)
    assert (swm, rdm) == (sw, rd), (K, n, lam, alpha, tol)
    assert ((bm == 0) == (bo == 0)).all()                                      # identical zero pattern
    np.testing.assert_allclose(bm, bo, rtol=0, atol=1e-11 * max(1.0, np.abs(bo).max()))


def test_pform_model_screen_and_readmission():
    """A problem whose strong-rule screen excludes coordinates that the KKT check has to bring back (:118-124): same number of
    rounds and sweeps in both forms."""
    rng = np.random.default_rng(7)
    found = 0
    for trial in range(8000):
        K, n = 6, 40
        mix = np.eye(K) + 0.8 * rng.normal(size=(K, K)) * (rng.random((K, K)) < 0.4)    # columns that share (also negatively) their parts
        X = rng.normal(size=(n, K)) @ mix
        y = X @ (2.0 * rng.normal(size=K)) + 0.3 * rng.normal(size=n)
        G, b = X.T @ X, X.T @ y
        lam, alpha = float(np.max(np.abs(b))) * rng.uniform(0.55, 0.95), float(rng.uniform(0.5, 1.0))
        bo, sw, rd = oracle.strong_cd(X, y, np.zeros(K), lam, alpha, G, b, tol=1e-7, perm_mode=1, seed=trial)
        if rd < 2:                                                             # nothing re-admitted (the usual case)
            continue
        found += 1
        bm, swm, rdm = cd_pform(G, b, np.zeros(K), lam, alpha, 1e-7, trial, 0)
        assert (swm, rdm) == (sw, rd)
        assert ((bm == 0) == (bo == 0)).all()
        np.testing.assert_allclose(bm, bo, rtol=0, atol=1e-11 * max(1.0, np.abs(bo).max()))
    assert found >= 5
