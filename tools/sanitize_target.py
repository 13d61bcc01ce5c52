"""Small target for compute-sanitizer: dense + masked fits and the batched solver on tiny shapes (both CD kernels, every phase)."""
import sys
sys.path.insert(0, ".")
import numpy as np
from insider_b200 import _cabi, synth

ctx = _cabi.Context(0)
for tuning, alpha, Q in ((0, 0.4, 0), (1, 0.4, 0), (0, 0.0, 0), (1, 0.3, 2), (0, 0.3, 2)):
    N, P, K = 40, 96, 6
    pb = synth.with_continuous(N=N, P=P, K=K, levels=(3, 5, 4), Q=Q, seed=3) if Q else synth.ageing_like(N=N, P=P, K=K, n_donors=9, seed=3)
    tr, te = synth.random_masks(N, P, 0.1, 4)
    F0, V0 = synth.init_factors(pb.levels, K, P, Q=Q, seed=5)
    prob = _cabi.HostProblem(pb.Y, pb.confounder, pb.X, tr, te, 1 if Q else 0)
    fac = _cabi.HostFactors(F0, V0, K)
    opt = _cabi.default_options()
    opt.lambda1 = opt.lambda2 = 2.0
    opt.alpha, opt.tuning, opt.global_tol, opt.sub_tol, opt.max_iter, opt.seed = alpha, tuning, 1e-12, 1e-5, 3, 11
    out = ctx.optimize(prob, fac, opt)
    print(f"tuning={tuning} alpha={alpha} Q={Q}: loss {out['loss']:.6g} sweeps {out['cd_sweeps']}", flush=True)
rng = np.random.default_rng(1)
for K in (5, 23, 30):
    X = rng.normal(size=(50, K)); G = X.T @ X; Y = X @ rng.normal(size=(K, 70)); Xty = X.T @ Y; w0 = np.zeros((K, 70))
    b1, s1 = ctx.strong_cd(G, Xty, w0, 0.5 * np.abs(Xty).max(), 0.7, tol=1e-7, seed=2)
    b2, s2 = ctx.strong_cd(np.stack([G] * 70), Xty, w0, 0.5 * np.abs(Xty).max(), 0.7, tol=1e-7, seed=2)
    print(f"strong_cd K={K}: sweeps {s1.sum()} / {s2.sum()}, max diff {np.abs(b1 - b2).max():.2e}", flush=True)
print("SANITIZE TARGET DONE")
