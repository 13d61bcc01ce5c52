"""Short target for ncu: a few ALS iterations at the bench shape. Usage: python tools/ncu_target.py [tuning] [iters] [P] [N] [K]"""
import sys
sys.path.insert(0, ".")
from insider_b200 import _cabi, synth

tuning = int(sys.argv[1]) if len(sys.argv) > 1 else 0
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
P = int(sys.argv[3]) if len(sys.argv) > 3 else 44477
N = int(sys.argv[4]) if len(sys.argv) > 4 else 377
K = int(sys.argv[5]) if len(sys.argv) > 5 else 23
pb = synth.ageing_like(N=N, P=P, K=K)
tr, te = synth.random_masks(N, P, 0.1, 7)
F0, V0 = synth.init_factors(pb.levels, K, P, seed=1)
ctx = _cabi.Context(0)
res = ctx.upload(_cabi.HostProblem(pb.Y, pb.confounder, None, tr if tuning else None, te if tuning else None, 0))
opt = _cabi.default_options()
opt.lambda1 = opt.lambda2 = 10.0
opt.alpha, opt.tuning, opt.global_tol, opt.sub_tol, opt.max_iter, opt.seed = 0.4, tuning, 1e-12, 1e-5, 10 ** 6, 1
s = res.begin(_cabi.HostFactors(F0, V0, K), opt)
done, ms = s.step(iters)
out = s.end(read_factors=False)
print(f"tuning={tuning} iters={iters} ms={ms:.2f} sweeps={out['cd_sweeps']} loss={out['loss']:.8g}")
