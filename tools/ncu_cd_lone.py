"""ncu target: the dense CD kernel with very few warps (latency-bound regime). Usage: python tools/ncu_cd_lone.py [n_genes] [K]"""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from insider_b200 import _cabi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
K = int(sys.argv[2]) if len(sys.argv) > 2 else 23
rng = np.random.default_rng(0)
N = 377
U = rng.normal(size=(N, K)) @ (np.eye(K) + 0.9 * rng.normal(size=(K, K)) / np.sqrt(K))     # correlated columns: slow CD
Y = U @ rng.normal(size=(K, n)) + rng.normal(size=(N, n))
G = U.T @ U
Xty = U.T @ Y
w0 = np.zeros((K, n))
ctx = _cabi.Context(0)
for rep in range(3):
    t0 = time.perf_counter()
    beta, sweeps = ctx.strong_cd(G, Xty, w0, 10.0, 0.4, tol=1e-9, seed=1)
    dt = time.perf_counter() - t0
    print(f"n={n} K={K}: sweeps mean {sweeps.mean():.0f} max {sweeps.max()}  wall {dt * 1e3:.3f} ms  -> {dt * 1e6 / sweeps.max():.3f} us per sweep-round (incl. launch + copies)", flush=True)
