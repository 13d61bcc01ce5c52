"""Developer experiment: dump what the device knows before the first column update (row factor after the first row update) and
the true per-gene sweep counts of iterations 0 and 1, to study predictors offline. Usage: python tools/order_predictor_data.py"""
import sys
sys.path.insert(0, ".")
import numpy as np
from insider_b200 import _cabi, synth
N, P, K = 377, 44477, 23
pb = synth.ageing_like(N=N, P=P, K=K)
F0, V0 = synth.init_factors(pb.levels, K, P, seed=1)
ctx = _cabi.Context(0)
res = ctx.upload(_cabi.HostProblem(pb.Y, pb.confounder, None, None, None, 0))
opt = _cabi.default_options(); opt.lambda1 = opt.lambda2 = 10.0
opt.alpha, opt.tuning, opt.global_tol, opt.sub_tol, opt.max_iter, opt.seed = 0.4, 0, 1e-12, 1e-5, 10 ** 6, 1
fac = _cabi.HostFactors(F0, V0, K)
s = res.begin(fac, opt)
s.step(1); sw0 = s.sweeps(P).copy()
s.read()
A0 = [f.copy() for f in fac.factors]
s.step(1); sw1 = s.sweeps(P).copy()
s.end(read_factors=False)
np.savez_compressed("gpurun_out/order_predictor_data.npz", sw0=sw0, sw1=sw1, **{f"A{i}": a for i, a in enumerate(A0)})
print("saved", sw0.mean(), sw1.mean())
