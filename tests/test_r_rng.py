"""R RNG restatement against well-known R outputs (the only externally pinned answers available offline) and the
product's split against the oracle's split (two independent implementations of R's sample())."""
import numpy as np

from oracle import oracle
from oracle.r_rng import RRng, ratio_splitter, randperm_b


def test_unif_known_answers():
    np.testing.assert_allclose(RRng(123).unif(3), [0.2875775, 0.7883051, 0.4089769], atol=5e-8)
    np.testing.assert_allclose(RRng(42).unif(2), [0.914806, 0.9370754], atol=5e-8)
    np.testing.assert_allclose(RRng(1).unif(2), [0.2655087, 0.3721239], atol=5e-8)


def test_sample_known_answers():
    assert RRng(123).sample(np.arange(1, 11)).tolist() == [3, 10, 2, 8, 6, 9, 1, 7, 5, 4]
    assert RRng(42).sample(np.arange(1, 11)).tolist() == [1, 5, 10, 8, 2, 4, 6, 9, 7, 3]
    assert RRng(1).sample(np.arange(1, 11)).tolist() == [9, 4, 7, 1, 2, 5, 3, 10, 6, 8]


def test_rnorm_known_answers():
    np.testing.assert_allclose(RRng(123).rnorm(3), [-0.56047565, -0.23017749, 1.55870831], atol=5e-9)
    np.testing.assert_allclose(RRng(1).rnorm(3), [-0.6264538, 0.1836433, -0.8356286], atol=5e-8)


def test_cpp_stream_matches_numpy_stream():
    np.testing.assert_array_equal(oracle.r_unif(123, 2000), RRng(123).unif(2000))
    # arma::randperm emulation: same stream, same ordering
    assert oracle.randperm_r(7, 23).tolist() == RRng(7).randperm(23).tolist()


def test_counter_permutation_cpp_vs_numpy():
    for args in [(0, 0, 0, 0, 23), (5, 3, 77, 2, 23), (2**63 + 11, 40, 44476, 9, 30), (1, 2, 3, 4, 1)]:
        a, b = oracle.randperm_b(*args), randperm_b(*args)
        assert a.tolist() == b.tolist()
        assert sorted(a.tolist()) == list(range(args[-1]))
    # the key does not depend on the gene: every gene at the same (ALS iteration, sweep) shares the visiting order
    assert oracle.randperm_b(5, 3, 77, 2, 23).tolist() == oracle.randperm_b(5, 3, 1234, 2, 23).tolist()
    # restricted to an active set: the active coordinates in the order of the full permutation
    for inc in ([0, 1, 2, 3], [2, 5, 11, 12, 20, 22], [7], list(range(23))):
        full = randperm_b(5, 3, 0, 2, 23)
        want = [c for c in full.tolist() if c in inc]
        a, b = oracle.randperm_b_inc(5, 3, 2, 23, inc), randperm_b(5, 3, 0, 2, 23, inc)
        assert [inc[i] for i in a.tolist()] == want == [inc[i] for i in b.tolist()]


def test_product_split_is_bit_exact_with_oracle_split():
    from insider_b200 import _cabi
    rng = np.random.default_rng(0)
    for shape, ratio in [((377, 200), 0.1), ((50, 31), 0.25), ((7, 5), 0.1)]:
        d = rng.random(shape)
        d[rng.random(shape) < 0.05] = np.nan          # NA entries are never sampled (R/utils.R:84-90)
        tr, te, na = _cabi.split(d, ratio, 123)
        s = ratio_splitter(d, ratio, rm_na_col=False, seed=123)
        assert np.array_equal(te != 0, s["test_indicator"])
        assert np.array_equal(tr != 0, s["train_indicator"])
        assert np.array_equal(na != 0, s["na_indicator"])
        assert te.sum() == int(np.floor((~np.isnan(d)).sum() * ratio))
        assert not np.any((tr != 0) & (te != 0))


def test_split_edge_cases():
    from insider_b200 import _cabi
    d = np.ones((3, 2))
    tr, te, na = _cabi.split(d, 0.0, 123)             # ratio 0: nothing held out
    assert te.sum() == 0 and tr.sum() == 6
    d[:] = np.nan                                      # all NA
    tr, te, na = _cabi.split(d, 0.5, 123)
    assert te.sum() == 0 and tr.sum() == 0 and na.sum() == 6


def test_sample2_vectorised_statement_equals_attempt_loop():
    """oracle/r_rng.py states do_sample2 twice (attempt loop / "first k distinct in-range values of the attempt stream"): same
    draws AND same RNG stream position afterwards, for one and two uniforms per attempt (bits < 16 / >= 16)."""
    for n, k, seed in [(70000, 30000, 1), (20000, 9000, 5), (1 << 16, 20000, 9), (100003, 50001, 3)]:
        a, b = RRng(seed), RRng(seed)
        assert np.array_equal(a._sample2(n, k), b._sample2_loop(n, k))
        assert a.unif_rand() == b.unif_rand()


def test_product_split_hashed_branch_is_bit_exact_with_oracle_split():
    """n > 1e7 non-NA entries and k <= n/2: sample.int takes its hashed branch (do_sample2) - the branch the 377 x 44477 mask of
    BASELINE config 2/3 takes (R/utils.R:91, 16.8 M entries). csrc/rsplit.cpp:85-98 against the independent NumPy statement.
    Both are restatements from memory of R's sources: UNVERIFIED against R itself (no R offline), see DESIGN.md section 6."""
    from insider_b200 import _cabi
    N, P = 377, 26600                                  # 10 028 200 entries
    rng = np.random.default_rng(7)
    d = rng.random((N, P))
    d[rng.random((N, P)) < 0.001] = np.nan
    n = int((~np.isnan(d)).sum())
    assert n > 1e7
    tr, te, na = _cabi.split(d, 0.1, 123)
    s = ratio_splitter(d, 0.1, rm_na_col=False, seed=123)
    assert int(te.sum()) == n // 10
    assert np.array_equal(te != 0, s["test_indicator"])
    assert np.array_equal(tr != 0, s["train_indicator"])
    assert np.array_equal(na != 0, s["na_indicator"])
    assert not np.any((tr != 0) & (te != 0))
