"""Developer experiment: how much faster is iteration t when the slot order of the dense CD solver comes from the TRUE sweep
counts of that iteration (upper bound of what a predictor can give)? Usage: python tools/order_potential.py [tuning]"""
import sys
sys.path.insert(0, ".")
import numpy as np
from insider_b200 import _cabi, synth
N, P, K = 377, 44477, 23
tuning = int(sys.argv[1]) if len(sys.argv) > 1 else 0
pb = synth.ageing_like(N=N, P=P, K=K)
tr, te = synth.random_masks(N, P, 0.1, 7)
F0, V0 = synth.init_factors(pb.levels, K, P, seed=1)
ctx = _cabi.Context(0)
res = ctx.upload(_cabi.HostProblem(pb.Y, pb.confounder, None, tr if tuning else None, te if tuning else None, 0))
opt = _cabi.default_options(); opt.lambda1 = opt.lambda2 = 10.0
opt.alpha, opt.tuning, opt.global_tol, opt.sub_tol, opt.max_iter, opt.seed = 0.4, tuning, 1e-12, 1e-5, 10 ** 6, 1
s = res.begin(_cabi.HostFactors(F0, V0, K), opt)
_, ms0 = s.step(1); sw0 = s.sweeps(P).copy()
s.end(read_factors=False)
for name, hint in (("no hint", None), ("true counts", sw0), ("true counts + 30 % noise", (sw0 * np.random.default_rng(0).uniform(0.7, 1.3, P)).astype(np.int32)),
                   ("rank-0.8 predictor", (sw0 + np.random.default_rng(1).normal(0, 0.75 * sw0.std(), P)).clip(1).astype(np.int32))):
    s = res.begin(_cabi.HostFactors(F0, V0, K), opt)
    if hint is not None:
        s.hint_sweeps(hint)
    _, ms = s.step(1)
    assert (s.sweeps(P) == sw0).all()
    s.end(read_factors=False)
    print(f"iteration 0 with {name}: {ms:.2f} ms", flush=True)
