"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product path).

NumPy restatement of the parts of R's RNG that the reference's R layer relies on:

* ``set.seed`` / Mersenne-Twister / ``unif_rand``      (R src/main/RNG.c; not in the reference tree)
* ``R_unif_index`` ("Rejection" sample.kind, R >= 3.6) and ``sample.int`` without replacement
  (classic path and the hashed ``sample2`` path taken when n > 1e7 and k <= n/2)
* ``rnorm`` by inversion (``init_parameters``: reference R/utils.R:40-43)
* ``ratio_splitter``                                    (reference R/utils.R:78-117)
* ``arma::randperm`` under RcppArmadillo                (reference src/coordinate_descent.cpp:89)

Known answers this module is pinned against (tests/test_r_rng.py): ``set.seed(123); runif(3)``,
``set.seed(42); runif(2)``, ``set.seed(123); sample(1:10)``, ``set.seed(42); sample(1:10)``,
``set.seed(123); rnorm(3)``.  The hashed ``sample2`` path has no known answer available offline
(marked UNVERIFIED; only the *set* of drawn indices matters for the mask).
"""
from __future__ import annotations

import math

import numpy as np

_I2_32M1 = 2.328306437080797e-10
_SCALE = 2.3283064365386963e-10


class RRng:
    """R's default generator (Mersenne-Twister, Inversion, Rejection)."""

    def __init__(self, seed: int):
        self.set_seed(seed)

    def set_seed(self, seed: int) -> None:
        s = np.uint32(seed & 0xFFFFFFFF)
        with np.errstate(over="ignore"):
            for _ in range(50):                       # initial scrambling (RNG.c: Randomize / RNG_Init)
                s = np.uint32(np.uint32(69069) * s + np.uint32(1))
            key = np.empty(625, dtype=np.uint32)
            for j in range(625):
                s = np.uint32(np.uint32(69069) * s + np.uint32(1))
                key[j] = s
        self._bg = np.random.MT19937()
        # i_seed[0] is `mti`, forced to 624 by FixupSeeds; i_seed[1..624] is the MT state.
        self._bg.state = {"bit_generator": "MT19937", "state": {"key": key[1:].copy(), "pos": 624}}
        self._buf = np.empty(0, dtype=np.float64)
        self._pos = 0

    # -- raw stream -------------------------------------------------------------------------------
    def unif(self, n: int) -> np.ndarray:
        """n successive unif_rand() values."""
        out = np.empty(n, dtype=np.float64)
        got = 0
        while got < n:
            if self._pos >= self._buf.size:
                raw = self._bg.random_raw(max(4096, n - got)).astype(np.float64)
                u = raw * _SCALE
                u[u <= 0.0] = 0.5 * _I2_32M1
                u[(1.0 - u) <= 0.0] = 1.0 - 0.5 * _I2_32M1
                self._buf, self._pos = u, 0
            take = min(n - got, self._buf.size - self._pos)
            out[got:got + take] = self._buf[self._pos:self._pos + take]
            self._pos += take
            got += take
        return out

    def unif_rand(self) -> float:
        return float(self.unif(1)[0])

    # -- R_unif_index ------------------------------------------------------------------------------
    def _rbits(self, bits: int) -> int:
        v = 0
        n = 0
        while n <= bits:
            v1 = int(math.floor(self.unif_rand() * 65536))
            v = 65536 * v + v1
            n += 16
        if bits < 64:
            v &= (1 << bits) - 1
        return v

    def unif_index(self, dn: int) -> int:
        if dn <= 0:
            return 0
        bits = int(math.ceil(math.log2(dn)))
        while True:
            dv = self._rbits(bits)
            if dv < dn:
                return dv

    # -- sample.int(n, k), replace = FALSE ---------------------------------------------------------
    def sample_int(self, n: int, k: int) -> np.ndarray:
        """1-based indices, as ``sample.int(n, k)``."""
        if n > 1e7 and k <= n / 2:
            return self._sample2(n, k)
        x = np.arange(n, dtype=np.int64)
        y = np.empty(k, dtype=np.int64)
        nn = n
        for i in range(k):
            j = self.unif_index(nn)
            y[i] = x[j] + 1
            nn -= 1
            x[j] = x[nn]
        return y

    def _sample2(self, n: int, k: int) -> np.ndarray:
        """do_sample2 (R src/main/unique.c, the `useHash` branch of sample.int: n > 1e7 and k <= n/2):
        ``for i < k: for j < 100: y[i] = R_unif_index(n) + 1; if (!isDuplicated(y, i)) break``.
        Restated from memory of the R sources (not available offline, so there is no known answer from R itself: UNVERIFIED
        against R). Vectorised statement: every attempt of R_unif_index consumes the same number of uniforms, so the attempts
        form one fixed stream; the result is its first k distinct in-range values, in order of first appearance. (The two
        statements differ only if 100 consecutive redraws all hit duplicates: probability < (k/n)^100.)"""
        if k <= 4096:
            return self._sample2_loop(n, k)
        bits = int(math.ceil(math.log2(n)))
        per = bits // 16 + 1                                       # unif_rand() calls per rbits()
        out = np.empty(0, dtype=np.int64)
        seen_sorted = np.empty(0, dtype=np.int64)
        while out.size < k:
            m = max(1024, int((k - out.size) * (2.0 ** bits / n) * 1.2) + 1024)
            u = self.unif(m * per).reshape(m, per)
            v = np.zeros(m, dtype=np.int64)
            for c in range(per):
                v = 65536 * v + np.floor(u[:, c] * 65536).astype(np.int64)
            v &= (1 << bits) - 1
            keep = v < n
            v, pos = v[keep], np.flatnonzero(keep)
            if seen_sorted.size:
                dup = np.isin(v, seen_sorted)
                v, pos = v[~dup], pos[~dup]
            uniq, first = np.unique(v, return_index=True)          # first occurrence of every value
            order = np.argsort(first)
            vals, used = uniq[order], pos[first[order]]
            need = k - out.size
            if vals.size > need:
                # rewind the uniforms that belong to attempts after the last one consumed
                last_attempt = used[need - 1]
                self._pos -= (m - 1 - last_attempt) * per
                vals = vals[:need]
            out = np.concatenate([out, vals + 1])
            seen_sorted = np.sort(out - 1)
        return out

    def _sample2_loop(self, n: int, k: int) -> np.ndarray:
        """The same, attempt by attempt (slow; cross-checks the vectorised statement on small cases)."""
        seen = set()
        y = np.empty(k, dtype=np.int64)
        for i in range(k):
            v = 0
            for _ in range(100):
                v = self.unif_index(n) + 1
                if v not in seen:
                    break
            seen.add(v)
            y[i] = v
        return y

    def sample(self, x: np.ndarray, k: int | None = None) -> np.ndarray:
        """``sample(x, k)`` for length(x) > 1."""
        x = np.asarray(x)
        if k is None:
            k = x.size
        return x[self.sample_int(x.size, k) - 1]

    # -- rnorm (INVERSION) -------------------------------------------------------------------------
    def rnorm(self, n: int, mean: float = 0.0, sd: float = 1.0) -> np.ndarray:
        from scipy.special import ndtri  # R uses Wichura AS241; agreement to the last ulp is unverified

        u = self.unif(2 * n)
        big = 134217728.0
        uu = np.floor(big * u[0::2]) + u[1::2]
        return mean + sd * ndtri(uu / big)

    # -- arma::randperm(n) under RcppArmadillo -----------------------------------------------------
    def randperm(self, n: int) -> np.ndarray:
        vals = (self.unif(n) * 2147483647.0).astype(np.int64)   # int(Rf_runif(0, RAND_MAX))
        return np.argsort(vals, kind="stable")


def init_parameters(rng: RRng, size: int, init_mean: float = 0.0, init_std: float = 0.001) -> np.ndarray:
    """reference R/utils.R:40-43."""
    return rng.rnorm(size, init_mean, init_std)


def ratio_splitter(data: np.ndarray, ratio: float = 0.1, rm_na_col: bool = True, seed: int = 123) -> dict:
    """reference R/utils.R:78-117. ``data`` is N x P (any memory order); indices are column-major like R."""
    data = np.array(data, dtype=np.float64, order="F", copy=True)
    n, p = data.shape
    na = np.isnan(data)
    data[na] = 0.0
    train = ~na
    rng = RRng(seed)                                                   # :89 set.seed(seed)
    existing = np.flatnonzero(~na.ravel(order="F")) + 1                # :90 1-based column-major linear index
    k = int(math.floor(existing.size * ratio))
    test_idx = rng.sample(existing, k) if existing.size > 1 else existing[:k]   # :91
    test = np.zeros(n * p, dtype=bool)
    test[test_idx - 1] = True
    test = test.reshape((n, p), order="F")
    testset = np.zeros_like(data)
    testset[test] = data[test]
    data[test] = 0.0                                                   # :98
    train[test] = False                                                # :100
    nz = (data != 0).sum(axis=0)                                       # :102
    keep = nz != 0 if rm_na_col else np.ones(p, dtype=bool)
    return {
        "trainset": data[:, keep], "testset": testset[:, keep],
        "train_indicator": train[:, keep], "test_indicator": test[:, keep], "na_indicator": na[:, keep],
        "n_zero_cols": int((nz == 0).sum()),
    }


def randperm_b(seed: int, als_iter: int, gene: int, draw: int, n: int, inc=None) -> np.ndarray:
    """Counter-based permutation (mode B) — the NumPy statement of oracle/insider_oracle.cpp:randperm.

    The key (seed, als_iter, draw) selects a table permutation of all ``n`` coordinates (``gene`` is not part of the key:
    every gene at sweep ``draw`` shares the order). With ``inc`` (ascending active coordinates out of n) the result is the
    visiting order of the active set as indices into ``inc``."""
    m64 = (1 << 64) - 1

    def mix64(z):
        z &= m64
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m64
        return z ^ (z >> 31)

    pk = mix64(seed + 0x9E3779B97F4A7C15 * (1 + als_iter)) ^ mix64(draw * 0x8CB92BA72F3D8DD7 + 0x2545F4914F6CDD1D)
    t = (pk >> 20) & 4095                                     # table entry (4096 permutations per size)
    key = mix64(0x1F83D9ABFB41BD6B ^ ((n << 32) | t))
    vals = np.array([mix64(key + 0x9E3779B97F4A7C15 * (i + 1)) >> 38 for i in range(n)], dtype=np.int64)
    order = np.argsort(vals, kind="stable")
    if inc is None:
        return order
    rank = np.empty(n, dtype=np.int64)
    rank[order] = np.arange(n)
    return np.argsort(rank[np.asarray(inc)], kind="stable")
