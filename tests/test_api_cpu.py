"""Host logic of insider_b200/api.py (the mirror of the reference's R API) on CPU, with the device calls replaced by a recording
fake: what is checked here is argument plumbing and grid semantics, line by line against R/insider.R and R/utils.R - not numerics
(those are the -m gpu tests' business). The split itself (insider_b200_split) is host code of the library and runs for real."""
import numpy as np
import pytest

from insider_b200 import _cabi, api


def _data(N=30, P=24, seed=3, na=0):
    rng = np.random.default_rng(seed)
    Y = np.asfortranarray(rng.gamma(2.0, 1.0, size=(N, P)))
    if na:
        Y.reshape(-1, order="F")[rng.choice(N * P, size=na, replace=False)] = np.nan
    conf = np.column_stack([rng.integers(1, 3, N), rng.integers(1, 4, N), np.arange(N) % 5 + 1]).astype(np.int64)
    conf[:2, 0] = (1, 2)
    conf[:3, 1] = (1, 2, 3)
    return Y, conf


# ---------------------------------------------------------------------------------------------------------- ratio_splitter / insider()
def test_ratio_splitter_partitions_the_entries_like_the_reference(capsys):
    """R/utils.R:78-117: NA entries are neither train nor test; floor(n_existing * ratio) test entries; the trainset is zero at
    test and NA positions; the testset holds the held-out values and zeros elsewhere."""
    Y, _ = _data(na=17)
    d = api.ratio_splitter(Y, ratio=0.1, rm_na_col=False)
    tr, te, na = d["train_indicator"], d["test_indicator"], d["na_indicator"]
    assert (na == np.isnan(Y)).all()
    assert ((tr.astype(int) + te.astype(int) + na.astype(int)) == 1).all()
    n_existing = int((~np.isnan(Y)).sum())
    assert int(te.sum()) == int(np.floor(n_existing * 0.1))                              # :89-91
    assert (d["trainset"][te] == 0).all() and (d["trainset"][na] == 0).all()
    assert (d["trainset"][tr] == Y[tr]).all()
    assert (d["testset"][te] == Y[te]).all() and (d["testset"][~te] == 0).all()
    assert "number of all zero columns removed: 0" in capsys.readouterr().out            # :103
    # same seed, same masks (set.seed(seed) inside the function, :88); another seed, another test set
    d2 = api.ratio_splitter(Y, ratio=0.1, rm_na_col=False)
    d3 = api.ratio_splitter(Y, ratio=0.1, rm_na_col=False, seed=7)
    assert (d2["test_indicator"] == te).all() and (d3["test_indicator"] != te).any()


def test_ratio_splitter_drops_all_zero_columns_from_its_outputs_only(capsys):
    """R/utils.R:102-109: with rm.na.col = TRUE the columns whose training part is all zero leave every returned matrix."""
    Y, _ = _data()
    Y[:, 5] = 0.0
    Y[:, 11] = np.nan
    d = api.ratio_splitter(Y, ratio=0.1, rm_na_col=True)
    assert "number of all zero columns removed: 2" in capsys.readouterr().out
    for k in ("trainset", "testset", "train_indicator", "test_indicator", "na_indicator"):
        assert d[k].shape == (Y.shape[0], Y.shape[1] - 2), k


def test_insider_object_has_the_reference_layout():
    """R/insider.R:18-67: field names, the integer indicators (:57-59), the N x 1 zero matrix for absent continuous covariates
    (:52-54), params (:61-64), data stored unchanged (:25-26: the NA zeroing there indexes with a NULL field and does nothing)."""
    Y, conf = _data()
    obj = api.insider(Y, conf, split_ratio=0.2, global_tol=1e-8, sub_tol=1e-4, tuning_iter=7, max_iter=123)
    assert list(obj) == ["data", "confounder", "inc_continuous", "ctns_confounder", "train_indicator", "test_indicator", "na_indicator", "params"]
    assert obj.r_class == "insider"
    assert (obj["data"] == Y).all() and obj["data"] is not Y
    assert (obj["confounder"] == conf).all()
    assert obj["inc_continuous"] == 0 and obj["ctns_confounder"].shape == (Y.shape[0], 1) and not obj["ctns_confounder"].any()
    for k in ("train_indicator", "test_indicator", "na_indicator"):
        assert obj[k].dtype == np.int32 and obj[k].shape == Y.shape and obj[k].flags.f_contiguous, k
    assert int(obj["test_indicator"].sum()) == int(np.floor(Y.size * 0.2))
    assert obj["params"] == dict(global_tol=1e-8, sub_tol=1e-4, tuning_iter=7, max_iter=123)
    X = np.random.default_rng(0).normal(size=(Y.shape[0], 2))
    objc = api.insider(Y, conf, ctns_confounder=X)
    assert objc["inc_continuous"] == 1 and (objc["ctns_confounder"] == X).all()


def test_insider_interaction_column_and_its_errors():
    """R/insider.R:28-47: levels of the interaction in order of first appearance, inserted as column 2; the two stop() messages."""
    Y, conf = _data()
    obj = api.insider(Y, conf, interaction_idx=np.array([1, 2]))
    c = obj["confounder"]
    assert c.shape == (conf.shape[0], conf.shape[1] + 1)
    assert (c[:, 0] == conf[:, 0]).all() and (c[:, 2:] == conf[:, 1:]).all()
    seen = {}
    expect = [seen.setdefault((a, b), len(seen) + 1) for a, b in zip(conf[:, 0].tolist(), conf[:, 1].tolist())]
    assert c[:, 1].tolist() == expect
    with pytest.raises(ValueError, match="out of the range of confounder"):
        api.insider(Y, conf, interaction_idx=np.array([1, 9]))
    for bad in (np.array([1.0, 2.0]), np.array([1])):
        with pytest.raises(ValueError, match="should be integers and its length must be greater than or equal to 2"):
            api.insider(Y, conf, interaction_idx=bad)


# ---------------------------------------------------------------------------------------------------------- tune()
class _FakeResident:
    def __init__(self):
        self.released = False

    def release(self):
        self.released = True


def _fake_backend(monkeypatch, rmse):
    """Replaces the upload and insider_b200_tune_batch; records every batch as a list of (K, lambda1, lambda2, alpha, options)."""
    calls, residents = [], []

    def resident_for(obj, ctx, masks=True):
        residents.append(_FakeResident())
        return residents[-1]

    def tune_batch(res, facs, opts):
        calls.append([(f.struct.K, o.lambda1, o.lambda2, o.alpha, o) for f, o in zip(facs, opts)])
        outs = []
        for f, o in zip(facs, opts):
            tr, te = rmse(f.struct.K, o.lambda1, o.alpha)
            outs.append(dict(train_rmse=tr, test_rmse=te, loss=0.0, iters_run=o.max_iter + 1, checks=[], cd_sweeps=0, loop_ms=0.0))
        return outs, [0] * len(facs)

    monkeypatch.setattr(api, "_resident_for", resident_for)
    monkeypatch.setattr(_cabi, "tune_batch", tune_batch)
    return calls, residents


def test_tune_two_phase_grid_follows_the_reference(monkeypatch, tmp_path, capsys):
    """R/insider.R:81-176: rank sweep at the fallback (0.1, 0.1, 0) when lambda / alpha are vectors (:120-121), which.min of the
    test RMSE (:136), then expand.grid(lambda, alpha) with lambda varying fastest and both rounded to 2 digits (:145-150), every
    fit with tuning = 1 and max_iter = tuning_iter (:116-121, :163-164); tables and CSV side effects (:130, :172)."""
    Y, conf = _data()
    obj = api.insider(Y, conf, tuning_iter=7, global_tol=1e-8, sub_tol=1e-4)
    monkeypatch.chdir(tmp_path)
    capsys.readouterr()

    def rmse(K, lam, alpha):                       # two ranks tie for the best test RMSE: which.min takes the first (and skips NA)
        return 1.0 / K, {4: float("nan"), 6: 0.5, 8: 0.5, 10: 0.7}[K] + 0.01 * lam + 0.1 * abs(alpha - 0.3)

    calls, residents = _fake_backend(monkeypatch, rmse)
    lam, alp = [1.004, 3.0, 5.126], [0.2, 0.3]
    out = api.tune(obj, latent_dimension=np.array([4, 6, 8, 10]), lambda_=lam, alpha=alp, seed=5, ctx=object())
    assert len(calls) == 2 and len(residents) == 1 and residents[0].released
    # phase 1
    assert [(c[0], c[1], c[2], c[3]) for c in calls[0]] == [(k, 0.1, 0.1, 0.0) for k in (4, 6, 8, 10)]
    assert out["latent_rank"] == 6
    assert out["rank_tuning"].shape == (4, 3) and out["rank_tuning"][:, 0].tolist() == [4, 6, 8, 10]
    assert np.allclose(out["rank_tuning"][:, 1], [1 / 4, 1 / 6, 1 / 8, 1 / 10])
    # phase 2
    grid = [(l, a) for a in alp for l in lam]                                             # expand.grid: first factor fastest
    assert [(c[0], c[1], c[2], c[3]) for c in calls[1]] == [(6, round(l, 2), round(l, 2), round(a, 2)) for l, a in grid]
    assert out["reg_tuning"].shape == (6, 4)
    assert out["reg_tuning"][:, 0].tolist() == [1.0, 3.0, 5.13, 1.0, 3.0, 5.13] and out["reg_tuning"][:, 1].tolist() == [0.2] * 3 + [0.3] * 3
    for batch in calls:
        for _, _, _, _, o in batch:
            assert (o.tuning, o.max_iter, o.global_tol, o.sub_tol, o.seed) == (1, 7, 1e-8, 1e-4, 5)
    assert (tmp_path / "insider_rank_tuning_result.csv").exists() and (tmp_path / "insider_R6_reg_tuning_result.csv").exists()
    printed = capsys.readouterr().out
    assert printed.count("Latent rank: ") == 4 and printed.count("parameter grid:") == 6 and "parameter grid: 5.13,0.2" in printed


def test_tune_rank_sweep_with_scalar_penalties_and_penalty_grid_at_a_given_rank(monkeypatch, tmp_path):
    """R/insider.R:116-118: scalar lambda and alpha are used for the rank sweep as they are, and nothing else runs (:142);
    :137-139: a single latent_dimension skips the sweep and is the rank of the penalty grid."""
    Y, conf = _data()
    obj = api.insider(Y, conf)
    monkeypatch.chdir(tmp_path)
    calls, _ = _fake_backend(monkeypatch, lambda K, lam, alpha: (1.0, abs(K - 7) + lam))
    out = api.tune(obj, latent_dimension=np.array([5, 7, 9]), lambda_=2.0, alpha=0.25, ctx=object())
    assert len(calls) == 1 and [(c[0], c[1], c[2], c[3]) for c in calls[0]] == [(k, 2.0, 2.0, 0.25) for k in (5, 7, 9)]
    assert out["latent_rank"] == 7 and out["reg_tuning"] is None
    calls.clear()
    out = api.tune(obj, latent_dimension=np.array([9]), lambda_=[1.0, 2.0], alpha=0.4, ctx=object(), write_csv=False)
    assert len(calls) == 1 and [(c[0], c[1], c[3]) for c in calls[0]] == [(9, 1.0, 0.4), (9, 2.0, 0.4)]
    assert out["latent_rank"] == 9 and out["rank_tuning"] is None and out["reg_tuning"].shape == (2, 4)
    assert not list(tmp_path.glob("insider_R9_*"))


def test_tune_initial_factors_do_not_depend_on_the_schedule(monkeypatch):
    """Every grid point draws its initial factors from (seed, phase, index) - the documented replacement for "whatever R's RNG
    stream holds at that moment" - with the reference's shapes (:107-114): one L_c x K matrix per confounder column, Q x K for
    the continuous block, K x P for the genes."""
    Y, conf = _data()
    X = np.random.default_rng(1).normal(size=(Y.shape[0], 2))
    obj = api.insider(Y, conf, ctns_confounder=X)
    seen = []

    def tune_batch(res, facs, opts):
        seen.append([(f.factors, f.V) for f in facs])
        return [dict(train_rmse=1.0, test_rmse=float(i), loss=0.0, iters_run=1, checks=[], cd_sweeps=0, loop_ms=0.0) for i in range(len(facs))], [0] * len(facs)

    monkeypatch.setattr(api, "_resident_for", lambda obj, ctx, masks=True: _FakeResident())
    monkeypatch.setattr(_cabi, "tune_batch", tune_batch)
    api.tune(obj, latent_dimension=np.array([3, 5]), lambda_=1.0, alpha=0.1, seed=11, ctx=object(), write_csv=False)
    api.tune(obj, latent_dimension=np.array([3, 5]), lambda_=1.0, alpha=0.1, seed=11, ctxs=[object(), object()], write_csv=False)
    (a, b), (a2, b2) = seen[0], seen[1]
    for (fa, va), (fb, vb), K in ((a, a2, 3), (b, b2, 5)):
        assert [f.shape for f in fa] == [(2, K), (3, K), (5, K), (2, K)] and va.shape == (K, Y.shape[1])
        assert all((x == y).all() for x, y in zip(fa, fb)) and (va == vb).all()
        assert abs(np.std(va)) < 0.003                                                   # init_parameters: N(0, 0.001^2), R/utils.R:40-43
    assert not (a[1][:3, :3] == b[1][:3, :3]).all()


def test_tune_argument_checks_have_the_reference_messages():
    """R/insider.R:83-89"""
    obj = {"params": {}}
    for ld in (None, np.array([4.0, 6.0])):
        with pytest.raises(ValueError, match="should be integer, numeric, and numeric"):
            api.tune(obj, latent_dimension=ld, lambda_=[1.0, 2.0], alpha=0.1)
    with pytest.raises(ValueError, match="should be integer, numeric, and numeric"):
        api.tune(obj, latent_dimension=np.array([4, 6]), lambda_=["a", "b"], alpha=0.1)
    with pytest.raises(ValueError, match="should be greater than 1"):
        api.tune(obj, latent_dimension=np.array([4]), lambda_=1.0, alpha=0.1)


# ---------------------------------------------------------------------------------------------------------- fit() / optimize()
class _RecordingCtx:
    def __init__(self):
        self.calls = []

    def optimize(self, prob, fac, opt):
        self.calls.append((prob, fac, opt))
        fac.V[...] = 2.0
        for i, f in enumerate(fac.factors):
            f[...] = 10.0 + i
        return dict(train_rmse=0.5, test_rmse=0.25, loss=3.0, iters_run=4, checks=[], cd_sweeps=9, loop_ms=1.0)


@pytest.mark.parametrize("partition", [0, 1])
def test_fit_passes_the_reference_arguments(partition):
    """R/insider.R:190-216: tuning = partition; the "train" indicator is train + test and the "test" indicator is the NA mask
    (:207-209; only read when partition = 1); max_iter / tolerances from params; cfd_matrices, column_factor, test_rmse stored."""
    Y, conf = _data(na=9)
    Yz = np.where(np.isnan(Y), 0.0, Y)
    obj = api.insider(Yz, conf, max_iter=77, global_tol=1e-7, sub_tol=1e-3)
    obj["na_indicator"] = np.asfortranarray(np.isnan(Y), dtype=np.int32)
    rc = _RecordingCtx()
    out = api.fit(obj, latent_dimension=4, lambda_=3.0, alpha=0.4, partition=partition, seed=2, ctx=rc)
    assert out is obj and len(rc.calls) == 1
    prob, fac, opt = rc.calls[0]
    assert (opt.tuning, opt.max_iter, opt.global_tol, opt.sub_tol, opt.lambda1, opt.lambda2, opt.alpha) == (partition, 77, 1e-7, 1e-3, 3.0, 3.0, 0.4)
    assert fac.struct.K == 4 and fac.struct.n_factors == conf.shape[1] and fac.V.shape == (4, Y.shape[1])
    assert prob.struct.inc_continuous == 0 and prob.struct.C == conf.shape[1] and (prob.Y == Yz).all()
    if partition == 1:
        assert prob.struct.mask_kind == _cabi.MASK_INT32
        assert (prob.train == obj["train_indicator"] + obj["test_indicator"]).all() and (prob.test == obj["na_indicator"]).all()
    else:
        assert prob.struct.mask_kind == _cabi.MASK_NONE and prob.train is None
    assert sorted(obj["cfd_matrices"]) == ["factor0", "factor1", "factor2"]
    assert (obj["cfd_matrices"]["factor1"] == 11.0).all() and (obj["column_factor"] == 2.0).all() and obj["test_rmse"] == 0.25


def test_optimize_updates_the_callers_factors_in_place_and_checks_inc_continuous():
    """src/optimize.cpp:283-284 (factors aliased, not copied) and :270-273 (the inc_continuous check, as an exception here)."""
    Y, conf = _data()
    A = [np.zeros((int(conf[:, c].max()), 3), order="F") for c in range(conf.shape[1])]
    V = np.zeros((3, Y.shape[1]), order="F")
    rc = _RecordingCtx()
    out = api.optimize(Y, A, V, conf, np.zeros((Y.shape[0], 1)), None, None, 0, 3, 2.0, 5.0, 0.3, 0, 1e-9, 1e-5, 40, ctx=rc)
    assert (V == 2.0).all() and all((a == 10.0 + i).all() for i, a in enumerate(A))
    assert sorted(out["row_matrices"]) == ["factor0", "factor1", "factor2"] and out["column_factor"].shape == V.shape
    opt = rc.calls[0][2]
    assert (opt.lambda1, opt.lambda2, opt.alpha, opt.tuning, opt.max_iter) == (2.0, 5.0, 0.3, 0, 40)
    with pytest.raises(ValueError, match="inc_continuous can only be 0 or 1"):
        api.optimize(Y, A, V, conf, None, None, None, 2, 3, ctx=rc)


def test_host_problem_rejects_the_rm_na_col_shape_mismatch():
    """R/utils.R:104-109 against R/insider.R:25: the indicators lose all-zero columns, object$data keeps them; the reference then reads
    out of bounds in C++ - here the mismatch is an error before anything is uploaded."""
    Y, conf = _data()
    Y[:, 3] = 0.0
    obj = api.insider(Y, conf)
    assert obj["train_indicator"].shape[1] == Y.shape[1] - 1 and obj["data"].shape == Y.shape
    with pytest.raises(ValueError, match="drops all-zero columns from the indicators only"):
        _cabi.HostProblem(obj["data"], obj["confounder"], None, obj["train_indicator"], obj["test_indicator"], 0)
