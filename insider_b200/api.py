"""Host-side mirror of the reference's R API (R is not available in this environment; see INTEGRATION.md for the
Rcpp shim that binds the same C ABI from R).

    insider()   reference R/insider.R:18-67      object constructor (train/test split, interaction column, params)
    tune()      reference R/insider.R:81-176     two-phase grid (rank sweep, then lambda x alpha at the best rank)
    fit()       reference R/insider.R:190-216    final fit, stores cfd_matrices / column_factor / test_rmse
    optimize()  reference R/RcppExports.R:20-22  the 16-argument call that the C ABI replaces
    ratio_splitter()  reference R/utils.R:78-117

Same names, argument order, defaults and error messages as the reference. All compute goes through
libinsider_b200.so on a B200; there is no CPU fallback.
"""
from __future__ import annotations

import math

import numpy as np

from . import _cabi

_default_ctx = None


def default_context() -> _cabi.Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = _cabi.Context(0)
    return _default_ctx


def set_default_context(ctx) -> None:
    global _default_ctx
    _default_ctx = ctx


class InsiderObject(dict):
    """The reference's S3 list of class "insider" (R/insider.R:24): fields are reachable as obj['data'] etc."""
    r_class = "insider"


def init_parameters(size, init_mean=0.0, init_std=0.001, rng=None):
    """reference R/utils.R:40-43 (rnorm). The Python mirror draws from a NumPy generator; the R shim uses R's rnorm."""
    rng = rng if rng is not None else np.random.default_rng()
    return rng.normal(init_mean, init_std, size)


def ratio_splitter(data, ratio=0.1, rm_na_col=True, seed=123):
    """reference R/utils.R:78-117; the masks come from insider_b200_split (R-exact Mersenne-Twister sample())."""
    data = np.array(data, dtype=np.float64, order="F", copy=True)
    tr, te, na = _cabi.split(data, ratio, seed)
    trainb, testb, nab = tr != 0, te != 0, na != 0
    data[nab] = 0.0
    testset = np.zeros_like(data)
    testset[testb] = data[testb]
    data[testb] = 0.0
    num_per_col = (data != 0).sum(axis=0)
    print(f"number of all zero columns removed: {int((num_per_col == 0).sum())} ")
    keep = num_per_col != 0 if rm_na_col else np.ones(data.shape[1], bool)
    return dict(trainset=data[:, keep], testset=testset[:, keep], train_indicator=trainb[:, keep], test_indicator=testb[:, keep],
                na_indicator=nab[:, keep])


def insider(data, confounder, ctns_confounder=None, interaction_idx=None, split_ratio=0.1, global_tol=1e-9, sub_tol=1e-5,
            tuning_iter=30, max_iter=50000):
    """reference R/insider.R:18-67"""
    data = np.asarray(data, dtype=np.float64)
    confounder = np.asarray(confounder)
    dataset = ratio_splitter(data, ratio=split_ratio)
    obj = InsiderObject()
    obj["data"] = np.array(data, order="F", copy=True)          # :25-26 (the NA zeroing there is a no-op in the reference)
    if interaction_idx is not None and np.issubdtype(np.asarray(interaction_idx).dtype, np.integer) and len(interaction_idx) > 1:
        idx = np.asarray(interaction_idx) - 1                    # R indices are 1-based
        if np.max(interaction_idx) > confounder.shape[1]:
            raise ValueError("The interaction_idx is out of the range of confounder!")
        sub = confounder[:, idx]
        _, first = np.unique(sub, axis=0, return_index=True)     # unique() keeps first-appearance order (:34)
        uniq = sub[np.sort(first)]
        ind = np.zeros(confounder.shape[0], dtype=confounder.dtype)
        for k in range(uniq.shape[0]):
            ind[np.all(sub == uniq[k], axis=1)] = k + 1          # :36-39
        obj["confounder"] = np.column_stack([confounder[:, 0], ind, confounder[:, 1:]])   # :40 interaction is column 2
    elif interaction_idx is None:
        obj["confounder"] = np.array(confounder, copy=True)
    else:
        raise ValueError("The interaction_idx should be integers and its length must be greater than or equal to 2!")
    if ctns_confounder is not None:
        obj["inc_continuous"] = 1
        obj["ctns_confounder"] = np.asarray(ctns_confounder, dtype=np.float64).reshape(confounder.shape[0], -1)
    else:
        obj["inc_continuous"] = 0
        obj["ctns_confounder"] = np.zeros((confounder.shape[0], 1))
    obj["train_indicator"] = np.asfortranarray(dataset["train_indicator"], dtype=np.int32)   # :57-59
    obj["test_indicator"] = np.asfortranarray(dataset["test_indicator"], dtype=np.int32)
    obj["na_indicator"] = np.asfortranarray(dataset["na_indicator"], dtype=np.int32)
    obj["params"] = dict(global_tol=global_tol, sub_tol=sub_tol, tuning_iter=tuning_iter, max_iter=max_iter)
    return obj


def optimize(data, cfd_factors, column_factor, cfd_indicators, ctns_confounder, train_indicator, test_indicator, inc_continuous,
             latent_dim, lambda1=1.0, lambda2=1.0, alpha=0.1, tuning=1, global_tol=1e-10, sub_tol=1e-5, max_iter=10000, *, seed=0,
             perm_mode=_cabi.PERM_COUNTER, verbose=0, ctx=None, resident=None):
    """The reference's ``optimize()`` (R/RcppExports.R:20-22 -> src/optimize.cpp:256-422), same 16 positional arguments.

    Returns the reference's list as a dict: row_matrices {factor0..}, column_factor, train_rmse, test_rmse, loss (plus
    iters_run / checks / cd_sweeps diagnostics). Like the reference, the passed factor arrays are updated in place when
    they are Fortran-ordered float64 arrays.
    """
    ctx = ctx or default_context()
    if inc_continuous not in (0, 1):
        raise ValueError("The value of prarameter inc_continuous can only be 0 or 1.")
    fac = _cabi.HostFactors(cfd_factors, column_factor, latent_dim)
    opt = _cabi.default_options()
    opt.lambda1, opt.lambda2, opt.alpha, opt.tuning = float(lambda1), float(lambda2), float(alpha), int(tuning)
    opt.global_tol, opt.sub_tol, opt.max_iter = float(global_tol), float(sub_tol), int(max_iter)
    opt.seed, opt.perm_mode, opt.verbose = int(seed), int(perm_mode), int(verbose)
    if resident is not None:
        out = resident.optimize(fac, opt)
    else:
        prob = _cabi.HostProblem(data, cfd_indicators, ctns_confounder, train_indicator if tuning == 1 or train_indicator is not None else None,
                                 test_indicator, inc_continuous)
        out = ctx.optimize(prob, fac, opt)
    for src, dst in zip(fac.factors, cfd_factors):               # in-place semantics of src/optimize.cpp:283-284
        if isinstance(dst, np.ndarray) and dst.shape == src.shape:
            dst[...] = src
    if isinstance(column_factor, np.ndarray) and column_factor.shape == fac.V.shape:
        column_factor[...] = fac.V
    out["row_matrices"] = {f"factor{i}": f for i, f in enumerate(fac.factors)}
    out["column_factor"] = fac.V
    return out


def _init_factors(obj, latent_rank, rng):
    conf = obj["confounder"]
    flist = [np.asfortranarray(init_parameters(len(np.unique(conf[:, i])) * latent_rank, rng=rng).reshape((-1, latent_rank), order="F"))
             for i in range(conf.shape[1])]
    if obj["inc_continuous"] == 1:
        q = obj["ctns_confounder"].shape[1]
        flist.append(np.asfortranarray(init_parameters(q * latent_rank, rng=rng).reshape((q, latent_rank), order="F")))
    V = np.asfortranarray(init_parameters(latent_rank * obj["data"].shape[1], rng=rng).reshape((latent_rank, -1), order="F"))
    return flist, V


def _resident_for(obj, ctx, masks=True):
    prob = _cabi.HostProblem(obj["data"], obj["confounder"], obj["ctns_confounder"], obj["train_indicator"] if masks else None,
                             obj["test_indicator"] if masks else None, obj["inc_continuous"])
    return ctx.upload(prob)


def tune(obj, latent_dimension=None, lambda_=0.1, alpha=0.0, *, seed=0, ctx=None, ctxs=None, write_csv=True):
    """reference R/insider.R:81-176. Returns dict(rank_tuning, latent_rank, reg_tuning).

    Grid points are independent fits. With ``ctxs`` (a list of contexts, e.g. one per GPU) the points of each phase are
    run as replicas by insider_b200_tune_batch, every context holding its own resident copy of the data — no
    communication. Each point draws its initial factors from a generator seeded by (seed, phase, point index), so the
    results do not depend on how points are scheduled (the reference uses whatever R's RNG stream holds at that moment).
    """
    type_msg = "TUNNING: The element of latent_dimension, lambda, and alpha should be integer, numeric, and numeric."
    if latent_dimension is None:                                  # :83 is.integer(NULL) is FALSE
        raise ValueError(type_msg)
    ld = np.atleast_1d(latent_dimension)
    try:
        lam = np.atleast_1d(np.asarray(lambda_, dtype=float))
        alp = np.atleast_1d(np.asarray(alpha, dtype=float))
    except (TypeError, ValueError):
        raise ValueError(type_msg) from None
    if not np.issubdtype(ld.dtype, np.integer):
        raise ValueError(type_msg)
    if len(ld) <= 1 and (len(lam) <= 1 and len(alp) <= 1):
        raise ValueError("TUNNING: The length of either latent_dimension or lambda and alpha should be greater than 1.")
    ctx_list = list(ctxs) if ctxs else [ctx or default_context()]
    p = obj["params"]
    residents = [_resident_for(obj, c) for c in ctx_list]       # all fits of a context share one upload

    def run_points(phase, points):
        """points: list of (K, lambda1, lambda2, alpha); returns the fitted dicts in order. One insider_b200_tune_batch call:
        the library spreads the points over the contexts (one host thread per context, shared queue)."""
        facs, opts = [], []
        for i, (K, l1, l2, a) in enumerate(points):
            flist, V = _init_factors(obj, int(K), np.random.default_rng([seed, phase, i]))
            facs.append(_cabi.HostFactors(flist, V, int(K)))
            o = _cabi.default_options()
            o.lambda1, o.lambda2, o.alpha, o.tuning = float(l1), float(l2), float(a), 1
            o.global_tol, o.sub_tol, o.max_iter, o.seed = float(p["global_tol"]), float(p["sub_tol"]), int(p["tuning_iter"]), int(seed)
            opts.append(o)
        out, _ = _cabi.tune_batch(residents, facs, opts)
        for f, o in zip(facs, out):
            o["row_matrices"] = {f"factor{i}": m for i, m in enumerate(f.factors)}
            o["column_factor"] = f.V
        return out

    rank_tuning, reg_tuning = None, None
    try:
        if len(ld) > 1:                                           # :98-132
            if len(lam) == 1 and len(alp) == 1:
                pts = [(int(k), float(lam[0]), float(lam[0]), float(alp[0])) for k in ld]
            else:
                pts = [(int(k), 0.1, 0.1, 0.0) for k in ld]       # :120-121
            for k in ld:
                print("Latent rank: ", k, "---------------------------------")
            fitted = run_points(0, pts)
            rank_tuning = np.array([[pt[0], f["train_rmse"], f["test_rmse"]] for pt, f in zip(pts, fitted)], dtype=float)
            if write_csv:
                np.savetxt("insider_rank_tuning_result.csv", rank_tuning, delimiter=",", header="rank,train_rmse,test_rmse", comments="")
        latent_rank = int(ld[np.nanargmin(rank_tuning[:, 2])]) if len(ld) > 1 else int(ld[0])   # :135-139 which.min skips NA
        if len(lam) > 1 or len(alp) > 1:                          # :142-174, expand.grid: lambda varies fastest
            pts = []
            for a0 in alp:
                for l0 in lam:
                    l, a = round(float(l0), 2), round(float(a0), 2)
                    print("parameter grid:", f"{l},{a}", "---------------------------------")
                    pts.append((latent_rank, l, l, a))
            fitted = run_points(1, pts)
            reg_tuning = np.array([[pt[1], pt[3], f["train_rmse"], f["test_rmse"]] for pt, f in zip(pts, fitted)], dtype=float)
            if write_csv:
                np.savetxt(f"insider_R{latent_rank}_reg_tuning_result.csv", reg_tuning, delimiter=",", header="lambda,alpha,train_rmse,test_rmse",
                           comments="")
    finally:
        for r in residents:
            r.release()
    return dict(rank_tuning=rank_tuning, latent_rank=latent_rank, reg_tuning=reg_tuning)


def fit(obj, latent_dimension=None, lambda_=None, alpha=None, partition=0, *, seed=0, ctx=None, verbose=0):
    """reference R/insider.R:190-216"""
    ctx = ctx or default_context()
    p = obj["params"]
    rng = np.random.default_rng(seed)
    flist, V = _init_factors(obj, int(latent_dimension), rng)
    indicator = obj["train_indicator"] + obj["test_indicator"]    # :207
    fitted = optimize(obj["data"], flist, V, obj["confounder"], obj["ctns_confounder"], indicator if partition == 1 else None,
                      obj["na_indicator"] if partition == 1 else None, obj["inc_continuous"], int(latent_dimension), lambda_, lambda_, alpha,
                      partition, p["global_tol"], p["sub_tol"], p["max_iter"], seed=seed, ctx=ctx, verbose=verbose)
    obj["cfd_matrices"] = fitted["row_matrices"]                  # :211-213
    obj["column_factor"] = fitted["column_factor"]
    obj["test_rmse"] = fitted["test_rmse"]
    obj["fit_info"] = {k: fitted[k] for k in ("train_rmse", "loss", "iters_run", "checks", "cd_sweeps", "loop_ms")}
    return obj


def strong_coordinate_descent(X, y, wstart, lambda_, alpha, XtX, Xty, tol, *, seed=0, ctx=None):
    """reference R/RcppExports.R:8-10 -> src/coordinate_descent.cpp:57-127, same 8 arguments (X and y are not needed in
    covariance form and may be None). Returns beta (K,)."""
    ctx = ctx or default_context()
    beta, _ = ctx.strong_cd(np.asarray(XtX, dtype=np.float64), np.asarray(Xty, dtype=np.float64).reshape(-1), np.asarray(wstart, dtype=np.float64).reshape(-1),
                            float(lambda_), float(alpha), tol=float(tol), seed=seed)
    return beta[:, 0]


def optimize_continuous_v2(data, indicator, updating_factor, c_factor, updating_confd, gram, lambda_, tuning, *, ctx=None):
    """reference R/RcppExports.R:16-18 -> src/optimize.cpp:77-137, same 8 arguments; `updating_factor` is updated in place
    like the reference's `rowvec&` (and returned). `gram` is accepted for signature compatibility: V V' is recomputed."""
    ctx = ctx or default_context()
    if tuning not in (0, 1):
        raise ValueError("Parameter tuning should be either 0 or 1!")
    w = ctx.optimize_continuous(data, indicator if tuning == 1 else None, updating_factor, c_factor, updating_confd, lambda_, tuning)
    if isinstance(updating_factor, np.ndarray) and updating_factor.size == w.size:
        updating_factor.reshape(-1)[...] = w
    return w


def glm_interaction(residual, train_indicator, interaction_indicator, column_factor, tol=1e-10, n_cores=10, *, ctx=None):
    """reference R/glm_interaction.R:2-30: list(coeff_matrix, pval_matrix). Like the reference, train_indicator, tol and n_cores
    are accepted and unused (the regression runs on every entry of the residual rows)."""
    ctx = ctx or default_context()
    coeff, pval = ctx.glm_interaction(residual, interaction_indicator, column_factor)
    return [coeff, pval]
