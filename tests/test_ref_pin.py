"""Pins the oracle against the reference ITSELF: oracle/_ref/libinsider_ref.so is the reference's own optimize.cpp,
coordinate_descent.cpp and utils.cpp, compiled where they lie under /root/reference against the small Armadillo / Rcpp API shim of
oracle/ref_shim/ (there is no R, Rcpp, Armadillo, BLAS or LAPACK in this image), serial, with arma::randperm drawn from R's
Mersenne-Twister. The oracle runs in mode A (the same R stream, one thread). What is compared is therefore every loop, update order,
index computation, stopping rule and decay ladder of the reference against their restatement; the dense linear algebra underneath
(products, Cholesky solves) is the shim's on one side and the oracle's own loops on the other.

Tolerances: factors 1e-10 (max-abs relative to the matrix's max-abs; measured 1e-13..1e-15), loss / RMSE 1e-12 relative, identical
numbers of elastic-net sweeps. Skipped when neither the built library nor /root/reference is there (the GPU box has the library: it
travels with the snapshot)."""
import numpy as np
import pytest

from insider_b200 import synth
from oracle import oracle, ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="neither oracle/_ref/libinsider_ref.so nor /root/reference is present")


@pytest.fixture(scope="module", autouse=True)
def _serial_oracle():
    oracle.set_threads(1)
    yield


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(1e-300, np.abs(b).max()))


def compare(pb, tr, te, F0, V0, K, lam, alpha, tuning, iters, X=None, gtol=1e-12, stol=1e-5, r_seed=42):
    inc = 1 if X is not None else 0
    ro = oracle.optimize(pb.Y, F0, V0, pb.confounder, X, tr, te, inc, K, lam, lam, alpha, tuning, gtol, stol, iters, perm_mode=0, r_seed=r_seed)
    Fr, Vr, trm, tem, loss = ref.optimize(pb.Y, F0, V0, pb.confounder, X, tr, te, inc, K, lam, lam, alpha, tuning, gtol, stol, iters, r_seed=r_seed)
    assert rel(Vr, ro.column_factor) <= 1e-10
    for a, b in zip(Fr, ro.factors):
        assert rel(a, b) <= 1e-10
    assert abs(loss - ro.loss) <= 1e-12 * abs(ro.loss)
    assert abs(trm - ro.train_rmse) <= 1e-12 * ro.train_rmse
    if tuning == 1:
        assert abs(tem - ro.test_rmse) <= 1e-12 * ro.test_rmse
    return ro


def test_compiled_sources_are_the_reference_files_unmodified():
    """oracle/ref_shim/SOURCES.sha256 records which files of the reference checkout oracle/_ref is compiled from (oracle/Makefile compiles them
    in place); where the checkout is present their hashes must match - nothing is patched, copied or pre-processed."""
    import hashlib
    import os
    src = os.path.join(ref.REFERENCE, "src")
    if not os.path.isdir(src):
        pytest.skip("reference checkout not present (the library was built where it is)")
    listed = [ln.split() for ln in open(os.path.join(os.path.dirname(ref.__file__), "ref_shim", "SOURCES.sha256")) if ln.strip()]
    assert {os.path.basename(n) for _, n in listed} >= {"optimize.cpp", "coordinate_descent.cpp", "utils.cpp"}
    for digest, name in listed:
        with open(os.path.normpath(os.path.join(src, name)), "rb") as f:
            assert hashlib.sha256(f.read()).hexdigest() == digest, name
    mk = open(os.path.join(os.path.dirname(ref.__file__), "Makefile")).read()
    assert "$(REF)/src/optimize.cpp" in mk and "$(REF)/src/coordinate_descent.cpp" in mk and "$(REF)/src/utils.cpp" in mk


def test_r_rng_known_answers_in_the_shim():
    # set.seed(123); runif(3) in R
    assert np.allclose(ref.r_unif(123, 3), [0.2875775201246142, 0.7883051354438066, 0.4089769218116999], rtol=0, atol=1e-15)
    assert np.array_equal(ref.r_unif(2024, 50), oracle.r_unif(2024, 50))


@pytest.mark.parametrize("tuning", [1, 0])
@pytest.mark.parametrize("alpha", [0.0, 0.4])
def test_optimize_matches_the_reference_sources(tuning, alpha):
    N, P, K = 60, 90, 6
    pb = synth.ageing_like(N=N, P=P, K=K, n_donors=11, seed=3)
    tr, te = synth.random_masks(N, P, 0.1, 6)
    F0, V0 = synth.init_factors(pb.levels, K, P, seed=7)
    ro = compare(pb, tr, te, F0, V0, K, 3.0, alpha, tuning, 12)
    assert ro.iters_run == 13


@pytest.mark.parametrize("tuning", [1, 0])
def test_optimize_at_the_ageing_row_count_and_rank(tuning):
    """377 samples, 3 confounders + pid x sid interaction, K = 23, lambda = 10, alpha = 0.4 (BASELINE.json config 1's design) on 320 genes."""
    N, P, K = 377, 320, 23
    pb = synth.ageing_like(N=N, P=P, K=K, seed=3)
    tr, te = synth.random_masks(N, P, 0.1, 6)
    F0, V0 = synth.init_factors(pb.levels, K, P, seed=7)
    ro = compare(pb, tr, te, F0, V0, K, 10.0, 0.4, tuning, 4)
    assert ro.cd_sweeps > 100000


@pytest.mark.parametrize("tuning", [1, 0])
def test_optimize_with_continuous_covariates(tuning):
    N, P, K = 50, 70, 5
    pb = synth.with_continuous(N=N, P=P, K=K, levels=(3, 4), Q=2, seed=11)
    tr, te = synth.random_masks(N, P, 0.15, 2)
    F0, V0 = synth.init_factors(pb.levels, K, P, Q=2, seed=5)
    compare(pb, tr, te, F0, V0, K, 2.0, 0.3, tuning, 11, X=pb.X)


@pytest.mark.parametrize("case", range(16))
def test_random_shapes_designs_and_penalties(case):
    """Seeded random problems: 1-3 categorical confounders with 2-9 levels, 0-2 continuous covariates, K = 1..7, alpha in {0, 0.3, 1},
    both tunings, masking ratios 5-40 %."""
    rng = np.random.default_rng(1000 + case)
    N, P, K = int(rng.integers(14, 44)), int(rng.integers(9, 40)), int(rng.integers(1, 8))
    levels = tuple(int(v) for v in rng.integers(2, 10, size=int(rng.integers(1, 4))))
    Q = int(rng.integers(0, 3))
    pb = synth.with_continuous(N=N, P=P, K=K, levels=levels, Q=Q, seed=case + 50)
    tr, te = synth.random_masks(N, P, float(rng.uniform(0.05, 0.4)), case)
    F0, V0 = synth.init_factors(pb.levels, K, P, Q=Q, seed=case)
    alpha = [0.0, 0.3, 1.0][case % 3]
    compare(pb, tr, te, F0, V0, K, float(rng.uniform(0.5, 6.0)), alpha, case % 2, 11, X=pb.X if Q else None, r_seed=case + 1)


def test_run_to_convergence_decay_ladder_and_break():
    """global_tol reached: the break iteration and the sub_tol decay ladder (src/optimize.cpp:381-407) decide the final state; equal final
    factors mean the restatement broke at the same evaluation with the same ladder."""
    N, P, K = 40, 50, 4
    pb = synth.ageing_like(N=N, P=P, K=K, n_donors=7, seed=9)
    tr, te = synth.random_masks(N, P, 0.1, 3)
    F0, V0 = synth.init_factors(pb.levels, K, P, seed=1)
    ro = compare(pb, tr, te, F0, V0, K, 5.0, 0.4, 1, 2000, gtol=1e-6)
    assert 10 <= ro.iters_run < 2000 and any(c["decay"] < 1.0 for c in ro.checks)


@pytest.mark.parametrize("alpha,lam", [(0.4, 10.0), (0.9, 4.0), (0.05, 1.0)])
def test_strong_coordinate_descent_matches_the_reference_source(alpha, lam):
    rng = np.random.default_rng(5)
    n, K = 80, 9
    X = rng.standard_normal((n, K)) @ (rng.standard_normal((K, K)) + 2.0)
    y = X @ (rng.standard_normal(K) * (rng.random(K) < 0.5)) + rng.standard_normal(n)
    XtX, Xty = X.T @ X, X.T @ y
    for w0 in (np.zeros(K), rng.standard_normal(K)):
        bo, sw_o, _ = oracle.strong_cd(X, y, w0, lam, alpha, XtX, Xty, tol=1e-9, perm_mode=0, r_seed=77)
        br, sw_r = ref.strong_cd(X, y, w0, lam, alpha, XtX, Xty, tol=1e-9, r_seed=77)
        assert sw_o == sw_r                                # one randperm per sweep in the reference
        assert np.abs(bo - br).max() <= 1e-12 * max(1.0, np.abs(br).max())
        assert np.array_equal(bo == 0.0, br == 0.0)        # the same coordinates are thresholded to exactly zero


@pytest.mark.parametrize("tuning", [1, 0])
def test_optimize_continuous_v2_matches_the_reference_source(tuning):
    rng = np.random.default_rng(8)
    N, P, K = 45, 60, 5
    V = rng.standard_normal((K, P))
    x = rng.standard_normal(N)
    w_true = rng.standard_normal(K)
    Y = np.outer(x, w_true) @ V + 0.1 * rng.standard_normal((N, P))
    ind = (rng.random((N, P)) > 0.12).astype(np.int32)
    w0 = rng.standard_normal(K) * 0.01
    G = V @ V.T
    wo = oracle.optimize_continuous_v2(Y, ind, w0, V, x, G, 2.0, tuning)
    wr = ref.optimize_continuous_v2(Y, ind, w0, V, x, G, 2.0, tuning)
    assert np.abs(wo - wr).max() <= 1e-11 * np.abs(wr).max()
