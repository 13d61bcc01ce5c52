"""Developer tool: latency of ONE warp of the two elastic-net solvers (us per sweep of the warp's slowest gene), through
insider_b200_strong_cd. 32 genes on a shared matrix -> k_cd_dense (thread per gene); 4 genes with per-column copies of the
same matrix -> k_cd_persistent (8 lanes per gene). The tail of every early ALS iteration and every small multi-GPU shard
runs at this latency. Usage: python tools/cd_lone_warp.py [K]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from insider_b200 import _cabi  # noqa: E402


def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 23
    rng = np.random.default_rng(3)
    N = 377
    Z = rng.standard_normal((N, K))
    A = rng.standard_normal((K, K)) + 6.0 * np.ones((K, K))          # strongly correlated columns: slow coordinate descent
    U = Z @ A
    XtX = U.T @ U
    ctx = _cabi.Context(0)
    for name, n, shared in (("k_cd_dense   (32 genes, shared matrix)", 32, True), ("k_cd_persistent (4 genes, per-gene matrices)", 4, False)):
        Y = rng.standard_normal((N, n)) * 3.0
        Xty = U.T @ Y
        w0 = np.zeros((K, n))
        G = XtX if shared else np.broadcast_to(XtX, (n, K, K)).copy()
        res = []
        for tol in (1e-2, 1e-13):
            best = 1e9
            for rep in range(4):
                t0 = time.perf_counter()
                beta, sweeps = ctx.strong_cd(G, Xty, w0, 10.0, 0.4, tol=tol, seed=5, als_iter=2)
                best = min(best, time.perf_counter() - t0)
            res.append((best, int(sweeps.max()), float(sweeps.mean())))
        (t0_, s0, _), (t1_, s1, m1) = res
        print(f"{name}: {1e6 * (t1_ - t0_) / max(1, s1 - s0):.3f} us per sweep  (slowest gene {s0} -> {s1} sweeps, mean {m1:.0f}; call {1e3 * t0_:.2f} -> {1e3 * t1_:.2f} ms)")
    ctx.close()


if __name__ == "__main__":
    main()
