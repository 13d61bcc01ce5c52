"""Developer tool: one-shot optimize timing under torchrun (N ranks). Usage: torchrun ... tools/e2e_n2.py"""
import os, sys, time
sys.path.insert(0, ".")
import torch, torch.distributed as tdist
from insider_b200 import _cabi, synth, dist as ibdist
local = int(os.environ.get("LOCAL_RANK", 0)); rank = int(os.environ.get("RANK", 0))
torch.cuda.set_device(local)
tdist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = ibdist.make_context(local)
N, P, K = 377, 44477, 23
pb = synth.ageing_like(N=N, P=P, K=K)
F0, V0 = synth.init_factors(pb.levels, K, P, seed=1)
prob = _cabi.HostProblem(pb.Y, pb.confounder, None, None, None, 0)
opt = _cabi.default_options(); opt.lambda1 = opt.lambda2 = 10.0
opt.alpha, opt.tuning, opt.global_tol, opt.sub_tol, opt.max_iter, opt.seed = 0.4, 0, 1e-12, 1e-5, 19, 1
for rep in range(3):
    tdist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter(); res = ctx.upload(prob); torch.cuda.synchronize(); t1 = time.perf_counter()
    s = res.begin(_cabi.HostFactors(F0, V0, K), opt); torch.cuda.synchronize(); t2 = time.perf_counter()
    s.step(1); t3 = time.perf_counter()
    s.step(1000); t4 = time.perf_counter()
    s.end(); t5 = time.perf_counter(); res.release()
    if rank == 0:
        print(f"rep {rep}: upload {1e3*(t1-t0):.1f} begin {1e3*(t2-t1):.1f} first step {1e3*(t3-t2):.1f} rest {1e3*(t4-t3):.1f} end {1e3*(t5-t4):.1f} ms", flush=True)
ctx.close(); tdist.destroy_process_group()
