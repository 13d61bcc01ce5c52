"""ORACLE — TEST INFRASTRUCTURE ONLY.

An independently written NumPy/SciPy statement of the same reference algorithm, used to cross-check the
C++ oracle (oracle/insider_oracle.cpp) on small problems. It deliberately uses the libraries the reference
uses underneath Armadillo (BLAS products through NumPy, LAPACK ``posv`` through
``scipy.linalg.solve(assume_a='pos')``) and is written array-at-a-time, so that an error in the loop-level
C++ restatement does not repeat here. Function names and argument order follow the reference sources.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import solve

from .r_rng import RRng, randperm_b


def compute_loss(residual, beta, lam, alpha):
    """src/utils.cpp:46-49"""
    return np.sum(np.square(residual)) / 2 + (1 - alpha) * lam * np.sum(np.square(beta)) / 2 + alpha * lam * np.sum(np.abs(beta))


def strong_coordinate_descent(X, y, wstart, lam, alpha, XtX, Xty, tol=1e-5, perm=None, stats=None):
    """src/coordinate_descent.cpp:57-127. ``perm(inc_idx, K)`` returns the visiting order of the active coordinates
    (indices into ``inc_idx``)."""
    beta = np.array(wstart, dtype=float)
    K = beta.size
    active = np.ones(K)
    ex_idx = np.flatnonzero(np.abs(Xty) < alpha * (2 * lam - np.max(np.abs(Xty))))      # :74
    active[ex_idx] = 0
    beta[ex_idx] = 0
    residual = y - X @ beta                                                               # :79
    iter_loss = compute_loss(residual, beta, lam, alpha)
    sweeps = 0
    while True:
        inc_idx = np.flatnonzero(active)
        ex_idx = np.flatnonzero(active == 0)
        while True:
            pre_loss = iter_loss
            order = perm(inc_idx, K) if perm is not None else np.arange(inc_idx.size)
            for i in range(inc_idx.size):
                k = inc_idx[order[i]]
                upper = residual @ X[:, k] + beta[k] * XtX[k, k]                          # :94
                if abs(upper) > lam * alpha:
                    update = np.sign(upper) * max(abs(upper) - lam * alpha, 0.0) / (XtX[k, k] + lam * (1 - alpha))
                else:
                    update = 0.0
                if update != beta[k]:
                    residual = residual - (update - beta[k]) * X[:, k]
                    beta[k] = update
            iter_loss = compute_loss(residual, beta, lam, alpha)
            sweeps += 1
            if not abs(pre_loss - iter_loss) > tol:                                        # :114
                break
        grad = XtX[np.ix_(ex_idx, inc_idx)] @ beta[inc_idx] - Xty[ex_idx]                 # :118
        viol = np.flatnonzero(np.abs(grad) > alpha * lam)
        if viol.size == 0:
            break
        active[ex_idx[viol]] = 1
    if stats is not None:
        stats["sweeps"] = stats.get("sweeps", 0) + sweeps
    return beta


def optimize_row(residual, indicator, updating_factor, c_factor, updating_confd, gram, lam, tuning):
    """src/optimize.cpp:139-198 (updates ``updating_factor`` in place)."""
    K = c_factor.shape[0]
    for s in np.unique(updating_confd):
        ids = np.flatnonzero(updating_confd == s)
        if tuning == 1:
            XtX = np.zeros((K, K))
            Xty = np.zeros(K)
            for k in ids:
                nz = np.flatnonzero(indicator[k, :])
                zero = np.flatnonzero(indicator[k, :] == 0)
                XtX += gram - c_factor[:, zero] @ c_factor[:, zero].T                     # :170
                Xty += c_factor[:, nz] @ residual[k, nz]                                  # :171
        else:
            Xtys = c_factor @ residual.T                                                  # :180
            XtX = ids.size * gram
            Xty = Xtys[:, ids].sum(axis=1)
        XtX = XtX + lam * np.eye(K)
        updating_factor[s - 1, :] = solve(XtX, Xty, assume_a="pos")


def optimize_continuous_v2(data, indicator, updating_factor, c_factor, updating_confd, gram, lam, tuning):
    """src/optimize.cpp:77-137; returns the updated 1 x K factor row."""
    w = np.array(updating_factor, dtype=float)
    x = updating_confd
    K = c_factor.shape[0]
    if tuning == 1:
        resid = data - np.outer(x, w @ c_factor)
        sq_factor = np.square(c_factor)
        sq_x = x ** 2
        norm_factor = sq_factor.sum(axis=1)
        zero = indicator == 0
        while True:
            pre = w.copy()
            for i in range(K):
                resid = resid + w[i] * np.outer(x, c_factor[i, :])
                Xty = x @ (indicator * resid) @ c_factor[i, :]                            # :111
                XtX = np.sum(sq_x * (norm_factor[i] - (zero * sq_factor[i, :]).sum(axis=1)))   # :112-115
                w[i] = Xty / (XtX + lam)
                resid = resid - w[i] * np.outer(x, c_factor[i, :])
            if np.sum(np.abs(pre - w)) < 1e-1:                                            # :122
                break
        return w
    Xty = c_factor @ data.T @ x
    XtX = (x @ x) * gram + lam * np.eye(K)
    return solve(XtX, Xty, assume_a="pos")


def optimize_col(data, indicator, row_factor, c_factor, lam, alpha, tuning, tol, perm_for_gene, stats=None):
    """src/optimize.cpp:200-253 (updates ``c_factor`` in place)."""
    K, P = c_factor.shape
    gram = row_factor.T @ row_factor
    if tuning == 1:
        for j in range(P):
            sel = np.flatnonzero(indicator[:, j])
            off = np.flatnonzero(indicator[:, j] == 0)
            feature = row_factor[sel, :]
            XtX = gram - row_factor[off, :].T @ row_factor[off, :]                        # :218-219
            outcome = data[sel, j]
            Xty = feature.T @ outcome
            if alpha == 0.0:
                c_factor[:, j] = solve(XtX + lam * np.eye(K), Xty, assume_a="pos")
            else:
                c_factor[:, j] = strong_coordinate_descent(feature, outcome, c_factor[:, j], lam, alpha, XtX, Xty, tol, perm_for_gene(j), stats)
    else:
        Xty = row_factor.T @ data
        if alpha == 0.0:
            c_factor[:, :] = solve(gram + lam * np.eye(K), Xty, assume_a="pos")
        else:
            for j in range(P):
                c_factor[:, j] = strong_coordinate_descent(row_factor, data[:, j], c_factor[:, j], lam, alpha, gram, Xty[:, j], tol, perm_for_gene(j), stats)


def evaluate(residual, train_mask, test_mask, tuning):
    """src/utils.cpp:56-77 -> (sum_residual, train_rmse, test_rmse)"""
    if tuning == 0:
        s = np.sum(np.square(residual))
        return s, np.sqrt(s / residual.size), np.nan
    s = np.sum(np.square(residual[train_mask]))
    return s, np.sqrt(s / train_mask.sum()), np.sqrt(np.mean(np.square(residual[test_mask])))


def compute_loss_global(cfd_factor, column_factor, lambda1, lambda2, alpha, sum_residual):
    """src/utils.cpp:79-102"""
    row_reg = sum(lambda1 * np.linalg.norm(f, "fro") ** 2 for f in cfd_factor)
    col_reg = lambda2 * (1 - alpha) * np.linalg.norm(column_factor, "fro") ** 2
    l1_reg = lambda2 * alpha * np.sum(np.abs(column_factor))
    return sum_residual / 2 + row_reg / 2 + col_reg / 2 + l1_reg


def optimize(data, cfd_factors, column_factor, cfd_indicators, ctns_confounder, train_indicator, test_indicator,
             inc_continuous, latent_dim, lambda1=1.0, lambda2=1.0, alpha=0.1, tuning=1, global_tol=1e-10, sub_tol=1e-5,
             max_iter=10000, perm_mode=1, seed=0, r_seed=1):
    """src/optimize.cpp:256-422. Returns a dict shaped like the reference's R list plus the check log."""
    data = np.asarray(data, dtype=float)
    N, P = data.shape
    F = [np.array(f, dtype=float) for f in cfd_factors]
    V = np.array(column_factor, dtype=float)
    Z = np.asarray(cfd_indicators, dtype=int).reshape(N, -1)
    C = Z.shape[1]
    X = None if ctns_confounder is None else np.asarray(ctns_confounder, dtype=float).reshape(N, -1)
    M = None if train_indicator is None else np.asarray(train_indicator, dtype=float)
    train_mask = None if train_indicator is None else np.asarray(train_indicator) != 0
    test_mask = None if test_indicator is None else np.asarray(test_indicator) != 0
    cfd_num = C + (1 if inc_continuous == 1 else 0)
    rstream = RRng(r_seed)
    onehot = [np.equal.outer(Z[:, c], np.unique(Z[:, c])).astype(float) for c in range(C)]   # :294-313

    def row_factor():
        U = np.zeros((N, latent_dim))
        for c in range(C):
            U += onehot[c] @ F[c]
        if inc_continuous == 1:
            U += X @ F[C]
        return U

    U = row_factor()
    residual = data - U @ V
    sum_residual, train_rmse, test_rmse = evaluate(residual, train_mask, test_mask, tuning)
    loss = compute_loss_global(F, V, lambda1, lambda2, alpha, sum_residual)
    checks = [dict(iter=-1, sum_residual=sum_residual, train_rmse=train_rmse, test_rmse=test_rmse, loss=loss, delta_loss=0.0, decay=1.0)]
    decay = 1.0
    it = 0
    stats = {}
    while it <= max_iter:
        gram = V @ V.T
        for c in range(cfd_num):
            if c < C:
                residual = residual + onehot[c] @ F[c] @ V
                optimize_row(residual, M, F[c], V, Z[:, c], gram, lambda1, tuning)
            else:
                for q in range(X.shape[1]):
                    residual = residual + np.outer(X[:, q], F[c][q, :] @ V)
                    F[c][q, :] = optimize_continuous_v2(residual, M, F[c][q, :], V, X[:, q], gram, lambda1, tuning)
                    if q != X.shape[1] - 1:
                        residual = residual - np.outer(X[:, q], F[c][q, :] @ V)
            if c != cfd_num - 1:
                residual = residual - onehot[c] @ F[c] @ V
        U = row_factor()

        def perm_for_gene(j, it=it):
            state = {"draw": 0}

            def perm(inc_idx, K):
                d = state["draw"]
                state["draw"] += 1
                if perm_mode == 0:
                    return rstream.randperm(inc_idx.size)
                if perm_mode == 1:
                    return randperm_b(seed, it, j, d, K, inc_idx)
                return np.arange(inc_idx.size)
            return perm

        optimize_col(data, M, U, V, lambda2, alpha, tuning, sub_tol * decay, perm_for_gene, stats)
        residual = data - U @ V
        if it % 10 == 0:
            pre_loss = loss
            sum_residual, train_rmse, test_rmse = evaluate(residual, train_mask, test_mask, tuning)
            loss = compute_loss_global(F, V, lambda1, lambda2, alpha, sum_residual)
            delta = pre_loss - loss
            for thr in (1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1e-1):
                if delta / 1000 <= thr:
                    decay = thr
                    break
            else:
                decay = 1.0
            checks.append(dict(iter=it, sum_residual=sum_residual, train_rmse=train_rmse, test_rmse=test_rmse, loss=loss, delta_loss=delta, decay=decay))
            if (pre_loss - loss) / pre_loss < global_tol:
                break
        it += 1
    return dict(row_matrices=F, column_factor=V, train_rmse=train_rmse, test_rmse=test_rmse, loss=loss, iters_run=it,
                checks=checks, cd_sweeps=stats.get("sweeps", 0))
