"""Developer diagnostic (not a test): runs the CUDA path against the oracle on small problems in every mode and prints
the discrepancies. Usage on a GPU box: python tools/gpu_dev_check.py [quick]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from insider_b200 import _cabi, synth  # noqa: E402
from oracle import oracle  # noqa: E402


def rel(a, b):
    return float(np.abs(a - b).max() / max(1e-300, np.abs(b).max()))


def run_case(ctx, name, N, P, K, tuning, alpha, lam=2.0, Q=0, iters=12, seed=3, gtol=1e-12):
    if Q:
        pb = synth.with_continuous(N=N, P=P, K=K, levels=(3, 5, 4), Q=Q, seed=seed)
    else:
        pb = synth.ageing_like(N=N, P=P, K=K, n_donors=min(11, N), seed=seed)
    tr, te = synth.random_masks(N, P, 0.1, seed + 1)
    F0, V0 = synth.init_factors(pb.levels, K, P, Q=Q, seed=seed + 2)
    inc = 1 if Q else 0
    t0 = time.time()
    ro = oracle.optimize(pb.Y, F0, V0, pb.confounder, pb.X, tr, te, inc, K, lam, lam, alpha, tuning, gtol, 1e-5, iters, perm_mode=1, seed=11)
    t_or = time.time() - t0
    prob = _cabi.HostProblem(pb.Y, pb.confounder, pb.X, tr, te, inc)
    fac = _cabi.HostFactors(F0, V0, K)
    opt = _cabi.default_options()
    opt.lambda1 = opt.lambda2 = lam
    opt.alpha, opt.tuning, opt.global_tol, opt.sub_tol, opt.max_iter, opt.seed = alpha, tuning, gtol, 1e-5, iters, 11
    t0 = time.time()
    try:
        rg = ctx.optimize(prob, fac, opt)
    except Exception as e:  # noqa: BLE001
        print(f"[{name}] GPU FAILED: {e}")
        return False
    t_gpu = time.time() - t0
    dv = rel(fac.V, ro.column_factor)
    da = max(rel(a, b) for a, b in zip(fac.factors, ro.factors))
    dl = abs(rg["loss"] - ro.loss) / abs(ro.loss)
    dt = abs(rg["train_rmse"] - ro.train_rmse) / ro.train_rmse
    dte = abs(rg["test_rmse"] - ro.test_rmse) / ro.test_rmse if tuning == 1 else 0.0
    ok = dv < 1e-8 and da < 1e-8 and dl < 1e-10 and rg["iters_run"] == ro.iters_run
    print(f"[{name}] N={N} P={P} K={K} tuning={tuning} alpha={alpha} Q={Q}: iters {rg['iters_run']}/{ro.iters_run} dV={dv:.2e} dA={da:.2e} "
          f"dloss={dl:.2e} dtrain={dt:.2e} dtest={dte:.2e} sweeps {rg['cd_sweeps']}/{ro.cd_sweeps} launches={rg['kernel_launches']} "
          f"t_gpu={t_gpu:.3f}s t_oracle={t_or:.3f}s {'OK' if ok else 'MISMATCH'}")
    if not ok:
        for i, (cg, co) in enumerate(zip(rg["checks"], ro.checks)):
            print(f"    check {i}: iter {cg['iter']}/{co['iter']} loss {cg['loss']:.12g}/{co['loss']:.12g} sse {cg['sum_residual']:.12g}/{co['sum_residual']:.12g} "
                  f"rowreg {cg['row_reg']:.6g}/{co['row_reg']:.6g} colreg {cg['col_reg']:.6g}/{co['col_reg']:.6g} l1 {cg['l1_reg']:.6g}/{co['l1_reg']:.6g}")
    return ok


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    ctx = _cabi.Context(0)
    results = []
    # initial evaluation only (max_iter such that a single iteration runs) then longer runs
    results.append(run_case(ctx, "dense-1it", 40, 48, 5, 0, 0.4, iters=0))
    results.append(run_case(ctx, "masked-1it", 40, 48, 5, 1, 0.4, iters=0))
    results.append(run_case(ctx, "dense-cd", 40, 48, 5, 0, 0.4))
    results.append(run_case(ctx, "masked-cd", 40, 48, 5, 1, 0.4))
    results.append(run_case(ctx, "dense-ridge", 40, 48, 5, 0, 0.0))
    results.append(run_case(ctx, "masked-ridge", 40, 48, 5, 1, 0.0))
    results.append(run_case(ctx, "masked-lasso", 40, 48, 5, 1, 1.0))
    results.append(run_case(ctx, "dense-cont", 50, 64, 6, 0, 0.4, Q=2))
    results.append(run_case(ctx, "masked-cont", 50, 64, 6, 1, 0.4, Q=2))
    if not quick:
        results.append(run_case(ctx, "K23-dense", 377, 500, 23, 0, 0.4, lam=10.0))
        results.append(run_case(ctx, "K23-masked", 377, 500, 23, 1, 0.4, lam=10.0))
        results.append(run_case(ctx, "K30-masked", 100, 333, 30, 1, 0.3, lam=3.0))
        results.append(run_case(ctx, "K9-masked", 90, 200, 9, 1, 0.3, lam=3.0))
        results.append(run_case(ctx, "bigN-dense", 1000, 160, 12, 0, 0.4, lam=4.0, iters=5))
        results.append(run_case(ctx, "bigN-masked", 1000, 160, 12, 1, 0.4, lam=4.0, iters=5))
    print("SUMMARY:", sum(results), "of", len(results), "cases OK")


if __name__ == "__main__":
    main()
