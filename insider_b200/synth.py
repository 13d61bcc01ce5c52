"""Synthetic data of the shapes the reference names. No network, no reference data files: the ageing
toy data (`data/ageing_example_data.RData`) is absent from the reference mount, so every config uses this generator.

Design (ageing-shaped): donor `did` (107 levels, every donor >= 1 sample), phenotype `pid` (2 levels, a function of the
donor), tissue `sid` (8 levels); interaction pid x sid = dense rank of unique pairs in first-appearance order, placed as
column 2 like reference R/insider.R:34-40. Truth follows the reference's own simulation (tests/simulation.rmd:19-74):
Gaussian factors, 30 % of gene-factor columns zeroed, Gaussian noise; expression is clipped at 0 and a fraction of the
entries is set to exact 0 to mimic log2(x+1) zeros (README.md:38-43).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

DEFAULT_SEED = 20240311


@dataclass
class SynthProblem:
    Y: np.ndarray                 # N x P, Fortran order
    confounder: np.ndarray        # N x C int32, 1-based levels (interaction column included when requested)
    X: np.ndarray | None          # N x Q continuous covariates
    levels: list                  # L_c per confounder column
    truth: dict = field(default_factory=dict)

    @property
    def shape(self):
        return self.Y.shape


def _levels_cover(rng, n, L):
    """n level ids in 1..L with every level used at least once."""
    z = np.concatenate([np.arange(1, L + 1), rng.integers(1, L + 1, size=max(0, n - L))])[:n]
    rng.shuffle(z)
    return z.astype(np.int32)


def interaction_column(a, b):
    """Dense rank of unique (a, b) pairs in first-appearance order (reference R/insider.R:34-39)."""
    seen, out = {}, np.zeros(len(a), dtype=np.int32)
    for i, key in enumerate(zip(a.tolist(), b.tolist())):
        if key not in seen:
            seen[key] = len(seen) + 1
        out[i] = seen[key]
    return out


def _expression(rng, U, V, mu, noise, zero_frac):
    N, P = U.shape[0], V.shape[1]
    Y = np.empty((N, P), order="F")
    bs = max(1, (1 << 24) // max(1, N))                           # column blocks keep the big shapes light on memory
    for j0 in range(0, P, bs):
        j1 = min(P, j0 + bs)
        blk = mu + U @ V[:, j0:j1] + rng.normal(scale=noise, size=(N, j1 - j0))
        np.maximum(blk, 0.0, out=blk)
        blk[rng.random(blk.shape) < zero_frac] = 0.0
        Y[:, j0:j1] = blk
    return Y


def _truth_v(rng, K, P):
    V = rng.normal(size=(K, P))
    V[:, rng.random(P) < 0.3] = 0.0                                # tests/simulation.rmd:25-26
    return V


def ageing_like(N=377, P=5000, K=23, n_donors=107, n_pid=2, n_sid=8, interaction=True, zero_frac=0.35, noise=0.5, mu=3.0,
                seed=DEFAULT_SEED) -> SynthProblem:
    """Configs 1-3: 377 x 5000 (toy) / 377 x 44477 (full), levels (2, <=16, 8, 107)."""
    rng = np.random.default_rng(seed)
    n_donors = min(n_donors, N)
    did = _levels_cover(rng, N, n_donors)
    donor_pid = np.concatenate([np.arange(1, n_pid + 1), rng.integers(1, n_pid + 1, size=max(0, n_donors - n_pid))])[:n_donors]
    pid = donor_pid[did - 1].astype(np.int32)
    if len(np.unique(pid)) < n_pid:                                # tiny N: force coverage
        pid[:n_pid] = np.arange(1, n_pid + 1)
    sid = _levels_cover(rng, N, min(n_sid, N))
    cols = [pid, sid, did]
    if interaction:
        cols = [pid, interaction_column(pid, sid), sid, did]      # position 2 (R/insider.R:40)
    conf = np.asfortranarray(np.stack(cols, axis=1).astype(np.int32))
    levels = [int(conf[:, c].max()) for c in range(conf.shape[1])]
    C = conf.shape[1]
    A = [rng.normal(size=(L, K)) / np.sqrt(C) for L in levels]
    V = _truth_v(rng, K, P)
    U = sum(A[c][conf[:, c] - 1] for c in range(C))
    return SynthProblem(_expression(rng, U, V, mu, noise, zero_frac), conf, None, levels, dict(A=A, V=V))


def with_continuous(N=5000, P=20000, K=20, levels=(4, 6, 10, 50), Q=2, zero_frac=0.2, noise=0.5, mu=3.0, seed=DEFAULT_SEED) -> SynthProblem:
    """Config 5: categorical + continuous covariates (optimize_continuous_v2 path). Q = 0 gives a plain multi-way design
    (config 4, GTEx-scale: levels=(54, 948), N=17382, P=56200, K=30)."""
    rng = np.random.default_rng(seed)
    conf = np.asfortranarray(np.stack([_levels_cover(rng, N, min(L, N)) for L in levels], axis=1))
    lv = [int(conf[:, c].max()) for c in range(conf.shape[1])]
    C = conf.shape[1]
    A = [rng.normal(size=(L, K)) / np.sqrt(C + (1 if Q else 0)) for L in lv]
    U = sum(A[c][conf[:, c] - 1] for c in range(C))
    X = W = None
    if Q:
        X = np.asfortranarray(rng.normal(size=(N, Q)))
        W = rng.normal(size=(Q, K)) / np.sqrt(C + 1)
        U = U + X @ W
    V = _truth_v(rng, K, P)
    return SynthProblem(_expression(rng, U, V, mu, noise, zero_frac), conf, X, lv, dict(A=A, W=W, V=V))


def gtex_like(N=17382, P=56200, K=30, n_tissue=54, n_donor=948, **kw) -> SynthProblem:
    return with_continuous(N=N, P=P, K=K, levels=(n_tissue, n_donor), Q=0, zero_frac=0.35, **kw)


def init_factors(levels, K, P, Q=0, seed=1, std=0.001):
    """N(0, 0.001^2) initial factors (reference R/utils.R:40-43), Fortran order like R matrices."""
    rng = np.random.default_rng(seed)
    F = [np.asfortranarray(rng.normal(0.0, std, size=(L, K))) for L in levels]
    if Q:
        F.append(np.asfortranarray(rng.normal(0.0, std, size=(Q, K))))
    V = np.asfortranarray(rng.normal(0.0, std, size=(K, P)))
    return F, V


def random_masks(N, P, ratio=0.1, seed=7):
    """Cheap stand-in for ratio_splitter when the R-exact split is not the thing under test."""
    rng = np.random.default_rng(seed)
    test = np.asfortranarray(rng.random((N, P)) < ratio)
    return np.asfortranarray(~test, dtype=np.int32), np.asfortranarray(test, dtype=np.int32)
