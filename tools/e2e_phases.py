"""Developer tool: where the end-to-end (host buffers in, factors out) time of the one-shot call goes. Usage: python tools/e2e_phases.py"""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import torch
from insider_b200 import _cabi, synth

N, P, K = 377, 44477, 23
pb = synth.ageing_like(N=N, P=P, K=K)
F0, V0 = synth.init_factors(pb.levels, K, P, seed=1)
Yp = torch.empty((P, N), dtype=torch.float64, pin_memory=True)
Yp.numpy()[...] = pb.Y.T
Yhost = Yp.numpy().T
ctx = _cabi.Context(0)
prob = _cabi.HostProblem(Yhost, pb.confounder, None, None, None, 0)
opt = _cabi.default_options()
opt.lambda1 = opt.lambda2 = 10.0
opt.alpha, opt.tuning, opt.global_tol, opt.sub_tol, opt.max_iter, opt.seed = 0.4, 0, 1e-12, 1e-5, 19, 1
for rep in range(3):
    t0 = time.perf_counter(); res = ctx.upload(prob); torch.cuda.synchronize(); t1 = time.perf_counter()
    fac = _cabi.HostFactors(F0, V0, K)
    s = res.begin(fac, opt); torch.cuda.synchronize(); t2 = time.perf_counter()
    done, ms = s.step(1000); t3 = time.perf_counter()
    out = s.end(); t4 = time.perf_counter()
    res.release(); t5 = time.perf_counter()
    print(f"rep {rep}: upload {1e3*(t1-t0):.2f} ms, begin {1e3*(t2-t1):.2f}, step {1e3*(t3-t2):.2f} (device {ms:.2f}), end {1e3*(t4-t3):.2f}, release {1e3*(t5-t4):.2f}; iters {out['iters_run']}")
    t0 = time.perf_counter(); fac = _cabi.HostFactors(F0, V0, K); o = ctx.optimize(prob, fac, opt); t1 = time.perf_counter()
    print(f"       one-shot optimize {1e3*(t1-t0):.2f} ms (loop {o['loop_ms']:.2f})")
