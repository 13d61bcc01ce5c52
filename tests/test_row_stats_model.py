"""CPU model of the row phase in SUFFICIENT-STATISTIC form (lib.cu:run_iteration) against the reference's
residual form (src/optimize.cpp:332-362 as restated by oracle/numpy_twin.py).

The reference keeps an N x P residual and adds / subtracts one block's contribution around every block solve. The CUDA path makes ONE
pass over Y per iteration for B_k = sum_j m_kj y_kj v_j, G = V V' and (masked) the complement Grams D_k = sum_{j: m_kj = 0} v_j v_j',
and then works on factor-sized data only: level s of block c solves (sum_{k in s} (G - D_k) + lambda I) a = sum_{k in s} [B_k -
(G - D_k)(u_k - a_old)], the row factor u_k is shifted after every block (Gauss-Seidel order kept), a continuous covariate runs the
reference's cyclic coordinate loop on H = sum x_k^2 (G - D_k), T = sum x_k (B_k - (G - D_k) u_k). This file states that algebra in
NumPy and checks it against the residual form on seeded problems: both tunings, several blocks, two continuous covariates."""
import numpy as np
import pytest
from scipy.linalg import solve

from oracle import numpy_twin


def row_phase_reference(Y, M, F, V, Z, X, lam, tuning):
    """The row phase of one ALS iteration exactly as numpy_twin.optimize() runs it (src/optimize.cpp:332-362)."""
    F = [f.copy() for f in F]
    C = Z.shape[1]
    onehot = [np.equal.outer(Z[:, c], np.unique(Z[:, c])).astype(float) for c in range(C)]
    U = sum(onehot[c] @ F[c] for c in range(C)) + (X @ F[C] if X is not None else 0.0)
    residual = Y - U @ V
    gram = V @ V.T
    n_blocks = C + (X is not None)
    for c in range(n_blocks):
        if c < C:
            residual = residual + onehot[c] @ F[c] @ V
            numpy_twin.optimize_row(residual, M, F[c], V, Z[:, c], gram, lam, tuning)
        else:
            for q in range(X.shape[1]):
                residual = residual + np.outer(X[:, q], F[c][q, :] @ V)
                F[c][q, :] = numpy_twin.optimize_continuous_v2(residual, M, F[c][q, :], V, X[:, q], gram, lam, tuning)
                if q != X.shape[1] - 1:
                    residual = residual - np.outer(X[:, q], F[c][q, :] @ V)
        if c != n_blocks - 1:
            residual = residual - onehot[c] @ F[c] @ V
    return F


def row_phase_statistics(Y, M, F, V, Z, X, lam, tuning):
    """The same update from the statistics of one pass over Y (what k_row_b / k_row_comp_gram / k_level_* / k_cont_* compute)."""
    F = [f.copy() for f in F]
    N, K, C = Y.shape[0], V.shape[0], Z.shape[1]
    G = V @ V.T
    if tuning == 1:
        B = (M * Y) @ V.T                                                        # k_row_b
        Gk = np.stack([G - V[:, M[k] == 0] @ V[:, M[k] == 0].T for k in range(N)])   # G - D_k, k_row_comp_gram
    else:
        B = Y @ V.T
        Gk = np.broadcast_to(G, (N, K, K))
    U = sum(F[c][Z[:, c] - 1] for c in range(C)) + (X @ F[C] if X is not None else 0.0)
    for c in range(C):                                                           # Gauss-Seidel over blocks, :335-362
        for s in np.unique(Z[:, c]):
            ids = np.flatnonzero(Z[:, c] == s)
            a_old = F[c][s - 1].copy()
            A = Gk[ids].sum(axis=0) + lam * np.eye(K)                            # k_level_gram (+ lambda on the diagonal in k_level_factor)
            rhs = sum(B[k] - Gk[k] @ (U[k] - a_old) for k in ids)                # k_row_rhs + per-level sum
            a_new = solve(A, rhs, assume_a="pos")
            F[c][s - 1] = a_new
            U[ids] += a_new - a_old                                              # k_level_update shifts the rows of U
    if X is not None:
        W = F[C]
        for q in range(X.shape[1]):
            x, w = X[:, q], W[q].copy()
            H = np.einsum("k,kij->ij", x * x, Gk)                                # k_cont_partial
            T = sum(x[k] * (B[k] - Gk[k] @ U[k]) for k in range(N)) + H @ w      # with this covariate's own contribution taken out of u_k
            if tuning == 1:
                while True:                                                      # cyclic order and stop of :102-126
                    pre = w.copy()
                    for i in range(K):
                        w[i] = (T[i] - H[i] @ w + H[i, i] * w[i]) / (H[i, i] + lam)
                    if np.sum(np.abs(pre - w)) < 1e-1:
                        break
            else:
                w = solve((x @ x) * G + lam * np.eye(K), T, assume_a="pos")      # :127-131
            U += np.outer(x, w - W[q])
            W[q] = w
    return F


def _problem(seed, N, P, K, levels, Q):
    rng = np.random.default_rng(seed)
    Z = np.column_stack([np.concatenate([np.arange(1, L + 1), rng.integers(1, L + 1, N - L)]) for L in levels])
    for c in range(Z.shape[1]):
        rng.shuffle(Z[:, c])
    X = rng.normal(size=(N, Q)) if Q else None
    F = [0.3 * rng.normal(size=(L, K)) for L in levels] + ([0.3 * rng.normal(size=(Q, K))] if Q else [])
    V = rng.normal(size=(K, P)) * (rng.random((1, P)) < 0.8)
    U = sum(F[c][Z[:, c] - 1] for c in range(len(levels))) + (X @ F[-1] if Q else 0.0)
    Y = np.maximum(0.0, 1.0 + U @ V + 0.3 * rng.normal(size=(N, P)))
    M = (rng.random((N, P)) < 0.85).astype(float)
    Fstart = [f + 0.2 * rng.normal(size=f.shape) for f in F]
    return Y, M, Fstart, V, Z, X


@pytest.mark.parametrize("tuning", [1, 0])
@pytest.mark.parametrize("shape", [(40, 60, 5, (2, 5, 3), 0), (36, 50, 4, (3, 4), 2), (50, 30, 7, (6,), 1), (30, 80, 3, (2, 3, 4, 5), 0)])
def test_statistic_form_equals_residual_form(shape, tuning):
    N, P, K, levels, Q = shape
    Y, M, F, V, Z, X = _problem(11 + N + tuning, N, P, K, levels, Q)
    lam = 2.5
    ref = row_phase_reference(Y, M if tuning else None, F, V, Z, X, lam, tuning)
    got = row_phase_statistics(Y, M, F, V, Z, X, lam, tuning)
    for a, b in zip(got, ref):
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-10 * max(1.0, np.abs(b).max()))
