"""Multi-GPU parity check (launch with torchrun, one rank per GPU): the gene-sharded fit must reproduce the oracle and the
single-GPU fit. Usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/mgpu_check.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as tdist  # noqa: E402

from insider_b200 import _cabi, dist as ibdist, synth  # noqa: E402


def rel(a, b):
    return float(np.abs(a - b).max() / max(1e-300, np.abs(b).max()))


def main():
    rank, world, local = ibdist.env_rank()
    torch.cuda.set_device(local)
    tdist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = ibdist.make_context(local)
    ok = True
    for (N, P, K, tuning, alpha, Q) in [(120, 1003, 8, 1, 0.4, 0), (120, 1003, 8, 0, 0.4, 0), (377, 2100, 23, 1, 0.4, 0), (90, 700, 6, 1, 0.3, 2), (64, 40, 5, 0, 0.0, 0)]:
        pb = synth.with_continuous(N=N, P=P, K=K, levels=(3, 5, 4), Q=Q, seed=3) if Q else synth.ageing_like(N=N, P=P, K=K, n_donors=21, seed=3)
        tr, te = synth.random_masks(N, P, 0.1, 4)
        F0, V0 = synth.init_factors(pb.levels, K, P, Q=Q, seed=5)
        prob = _cabi.HostProblem(pb.Y, pb.confounder, pb.X, tr, te, 1 if Q else 0)
        fac = _cabi.HostFactors(F0, V0, K)
        opt = _cabi.default_options()
        opt.lambda1 = opt.lambda2 = 4.0
        opt.alpha, opt.tuning, opt.global_tol, opt.sub_tol, opt.max_iter, opt.seed = alpha, tuning, 1e-12, 1e-5, 11, 7
        out = ctx.optimize(prob, fac, opt)
        if rank == 0:
            from oracle import oracle
            ro = oracle.optimize(pb.Y, F0, V0, pb.confounder, pb.X, tr, te, 1 if Q else 0, K, 4.0, 4.0, alpha, tuning, 1e-12, 1e-5, 11, perm_mode=1, seed=7)
            dv = rel(fac.V, ro.column_factor)
            da = max(rel(a, b) for a, b in zip(fac.factors, ro.factors))
            dl = abs(out["loss"] - ro.loss) / ro.loss
            good = dv < 1e-8 and da < 1e-8 and dl < 1e-10 and out["iters_run"] == ro.iters_run and out["cd_sweeps"] * 1 >= 0
            ok &= good
            print(f"world={world} N={N} P={P} K={K} tuning={tuning} alpha={alpha} Q={Q}: iters {out['iters_run']}/{ro.iters_run} dV={dv:.2e} dA={da:.2e} "
                  f"dloss={dl:.2e} local sweeps {out['cd_sweeps']} (oracle total {ro.cd_sweeps}) {'OK' if good else 'MISMATCH'}", flush=True)
    tdist.barrier()
    ctx.close()
    tdist.destroy_process_group()
    if rank == 0:
        print("MGPU SUMMARY:", "ALL OK" if ok else "FAILED", flush=True)
        sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
