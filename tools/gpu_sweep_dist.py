"""Developer tool: per-gene CD sweep distribution per ALS iteration on the GPU (dense path), with the lockstep
efficiency of 32-gene warps. Usage: python tools/gpu_sweep_dist.py [P] [iters] [first_iteration]"""
import sys
sys.path.insert(0, ".")
import numpy as np
from insider_b200 import _cabi, synth

P = int(sys.argv[1]) if len(sys.argv) > 1 else 44477
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
N, K = 377, 23
pb = synth.ageing_like(N=N, P=P, K=K)
F0, V0 = synth.init_factors(pb.levels, K, P, seed=1)
ctx = _cabi.Context(0)
res = ctx.upload(_cabi.HostProblem(pb.Y, pb.confounder, None, None, None, 0))
opt = _cabi.default_options()
opt.lambda1 = opt.lambda2 = 10.0
opt.alpha, opt.tuning, opt.global_tol, opt.sub_tol, opt.max_iter, opt.seed = 0.4, 0, 1e-12, 1e-5, 10 ** 6, 1
s = res.begin(_cabi.HostFactors(F0, V0, K), opt)
prev = None
if skip:
    s.step(skip)
for it in range(skip, skip + iters):
    done, ms = s.step(1)
    sw = s.sweeps(P).astype(np.int64)
    n = P // 32 * 32
    w = sw[:n].reshape(-1, 32)
    eff = w.sum() / (w.max(1).sum() * 32)
    effs = float("nan")
    if prev is not None:
        o = np.argsort(-prev, kind="stable")
        ws = sw[o][:n].reshape(-1, 32)
        effs = ws.sum() / (ws.max(1).sum() * 32)
    print(f"it {it}: {ms:.2f} ms  mean {sw.mean():.1f}  pct(50,90,99,99.9,100) {np.percentile(sw, [50, 90, 99, 99.9, 100]).astype(int)}  "
          f"warp-max mean {w.max(1).mean():.0f} max {w.max(1).max()}  lockstep eff natural {eff:.2f} sorted-by-prev {effs:.2f}", flush=True)
    prev = sw
s.end(read_factors=False)
