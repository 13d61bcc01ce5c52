"""Full fit() of the ageing-shaped problem to global_tol through the one-shot C ABI (host buffers): time-to-tol.
Usage: python tools/time_to_tol.py [P] [tuning] [global_tol] [max_iter]"""
import sys
import time

sys.path.insert(0, ".")
from insider_b200 import _cabi, synth

P = int(sys.argv[1]) if len(sys.argv) > 1 else 44477
tuning = int(sys.argv[2]) if len(sys.argv) > 2 else 0
gtol = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-9
max_iter = int(sys.argv[4]) if len(sys.argv) > 4 else 50000
N, K = 377, 23
pb = synth.ageing_like(N=N, P=P, K=K)
tr, te = synth.random_masks(N, P, 0.1, 7)
F0, V0 = synth.init_factors(pb.levels, K, P, seed=1)
ctx = _cabi.Context(0)
prob = _cabi.HostProblem(pb.Y, pb.confounder, None, tr if tuning else None, te if tuning else None, 0)
opt = _cabi.default_options()
opt.lambda1 = opt.lambda2 = 10.0
opt.alpha, opt.tuning, opt.global_tol, opt.sub_tol, opt.max_iter, opt.seed = 0.4, tuning, gtol, 1e-5, max_iter, 1
for rep in range(2):
    fac = _cabi.HostFactors(F0, V0, K)
    t0 = time.perf_counter()
    out = ctx.optimize(prob, fac, opt)
    t = time.perf_counter() - t0
    print(f"rep {rep}: P={P} tuning={tuning} global_tol={gtol}: iterations {out['iters_run']}, wall {t:.3f} s (device loop {out['loop_ms'] / 1e3:.3f} s), "
          f"sweeps/gene-iter {out['cd_sweeps'] / P / max(1, out['iters_run']):.1f}, loss {out['loss']:.10g}, train_rmse {out['train_rmse']:.6g}, "
          f"decays {[c['decay'] for c in out['checks']][-4:]}, launches {out['kernel_launches']}", flush=True)
