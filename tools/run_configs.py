"""Runs the non-headline BASELINE.json configs at full size on one B200 and prints timings + sanity properties.
Usage: python tools/run_configs.py [tune] [gtex] [cont]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from insider_b200 import _cabi, api, synth

which = set(sys.argv[1:]) or {"tune", "gtex", "cont"}
ctx = _cabi.Context(0)
api.set_default_context(ctx)


def opts(lam, alpha, tuning, max_iter, gtol=1e-9):
    o = _cabi.default_options()
    o.lambda1 = o.lambda2 = lam
    o.alpha, o.tuning, o.global_tol, o.sub_tol, o.max_iter, o.seed = alpha, tuning, gtol, 1e-5, max_iter, 1
    return o


if "tune" in which:
    # config 3: tune() grid on the 377 x 44477 shape: ranks 10..30 step 2, lambda 1..19 step 2, alpha {.2,.3,.4,.5}, tuning_iter 30
    t0 = time.time()
    pb = synth.ageing_like(N=377, P=44477, K=23, interaction=False)
    t_gen = time.time() - t0
    t0 = time.time()
    obj = api.insider(pb.Y, pb.confounder, None, np.array([1, 2]), split_ratio=0.1, tuning_iter=30)
    t_obj = time.time() - t0
    t0 = time.time()
    res = api.tune(obj, np.arange(10, 31, 2, dtype=np.int64), np.arange(1.0, 20.0, 2.0), np.array([0.2, 0.3, 0.4, 0.5]), seed=1, write_csv=False)
    t_tune = time.time() - t0
    print(f"[config 3] tune grid 11 ranks + 40 (lambda, alpha) points x 31 iterations on 377x44477: {t_tune:.1f} s total ({t_tune / 51:.2f} s per fit); "
          f"data gen {t_gen:.1f} s, insider() incl. R-exact split of 16.8M entries {t_obj:.1f} s; chosen rank {res['latent_rank']}; "
          f"best test rmse {res['reg_tuning'][:, 3].min():.5f}", flush=True)
    print("  rank_tuning test rmse:", np.round(res["rank_tuning"][:, 2], 5).tolist(), flush=True)

if "cont" in which:
    # config 5: 5000 x 20000, 4 categorical + 2 continuous, K = 20
    pb = synth.with_continuous(N=5000, P=20000, K=20, levels=(4, 6, 10, 50), Q=2)
    tr, te = synth.random_masks(5000, 20000, 0.1, 3)
    F0, V0 = synth.init_factors(pb.levels, 20, 20000, Q=2, seed=1)
    prob = _cabi.HostProblem(pb.Y, pb.confounder, pb.X, tr, te, 1)
    t0 = time.time(); res = ctx.upload(prob); t_up = time.time() - t0
    for tuning in (1, 0):
        fac = _cabi.HostFactors(F0, V0, 20)
        s = res.begin(fac, opts(5.0, 0.4, tuning, 10 ** 6))
        times = [s.step(1)[1] for _ in range(21)]
        out = s.end()
        losses = [c["loss"] for c in out["checks"]]
        print(f"[config 5] 5000x20000 K=20 C=4 Q=2 tuning={tuning}: upload {t_up:.2f} s; ms/iter first {times[0]:.1f}, median {np.median(times):.2f}, last {times[-1]:.2f}; "
              f"loss {losses[0]:.6g} -> {losses[-1]:.6g} monotone={all(b <= a for a, b in zip(losses, losses[1:]))}; sweeps/gene-iter {out['cd_sweeps'] / 20000 / 21:.1f}; "
              f"train_rmse {out['train_rmse']:.5f} test_rmse {out['test_rmse']:.5f}", flush=True)
    res.release()

if "gtex" in which:
    # config 4: GTEx-scale 17382 x 56200, tissue (54) x donor (948), K = 30: 10 iterations timed
    t0 = time.time()
    pb = synth.gtex_like()
    print(f"[config 4] generated 17382x56200 ({pb.Y.nbytes / 1e9:.1f} GB) in {time.time() - t0:.0f} s", flush=True)
    F0, V0 = synth.init_factors(pb.levels, 30, 56200, seed=1)
    prob = _cabi.HostProblem(pb.Y, pb.confounder, None, None, None, 0)
    t0 = time.time(); res = ctx.upload(prob); t_up = time.time() - t0
    fac = _cabi.HostFactors(F0, V0, 30)
    s = res.begin(fac, opts(10.0, 0.4, 0, 10 ** 6))
    times = [s.step(1)[1] for _ in range(11)]
    out = s.end()
    losses = [c["loss"] for c in out["checks"]]
    print(f"[config 4] 17382x56200 K=30 dense: upload {t_up:.1f} s; ms/iter {np.round(times, 1).tolist()}; loss {losses[0]:.6g} -> {losses[-1]:.6g}; "
          f"sweeps/gene-iter {out['cd_sweeps'] / 56200 / 11:.1f}; bytes/iter (SURVEY 8d) {(16 * 17382 * 56200 + 24 * 30 * 56200 + 16 * 17382 * 30) / 1e9:.2f} GB", flush=True)
    res.release()
