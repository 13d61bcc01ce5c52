"""Generates tests/golden/*.npz. There are no reference-side golden vectors (the reference has no tests and cannot be
built here: no R / Rcpp / RcppArmadillo), so these fixtures pin the ORACLE against itself over time and against the
independently written NumPy/SciPy twin: each fixture stores the inputs' seeds and the outputs of oracle/numpy_twin.py
(BLAS + LAPACK posv). Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from insider_b200 import synth  # noqa: E402
from oracle import numpy_twin  # noqa: E402

CASES = {
    # name: (N, P, K, tuning, alpha, lambda, Q, iters)
    "masked_cd": (36, 40, 5, 1, 0.4, 2.0, 0, 12),
    "dense_cd": (36, 40, 5, 0, 0.4, 2.0, 0, 12),
    "masked_ridge": (30, 33, 4, 1, 0.0, 0.1, 0, 11),
    "dense_ridge": (30, 33, 4, 0, 0.0, 0.1, 0, 11),
    "masked_lasso": (30, 33, 4, 1, 1.0, 1.5, 0, 11),
    "masked_cont": (40, 48, 6, 1, 0.3, 2.0, 2, 11),
    "dense_cont": (40, 48, 6, 0, 0.3, 2.0, 2, 11),
}


def make_inputs(name):
    N, P, K, tuning, alpha, lam, Q, iters = CASES[name]
    seed = 100 + sorted(CASES).index(name)
    if Q:
        pb = synth.with_continuous(N=N, P=P, K=K, levels=(3, 4, 5), Q=Q, seed=seed)
    else:
        pb = synth.ageing_like(N=N, P=P, K=K, n_donors=9, seed=seed)
    tr, te = synth.random_masks(N, P, 0.1, seed + 1)
    F0, V0 = synth.init_factors(pb.levels, K, P, Q=Q, seed=seed + 2)
    return pb, tr, te, F0, V0


if __name__ == "__main__":
    out = os.path.dirname(os.path.abspath(__file__))
    for name, (N, P, K, tuning, alpha, lam, Q, iters) in CASES.items():
        pb, tr, te, F0, V0 = make_inputs(name)
        r = numpy_twin.optimize(pb.Y, F0, V0, pb.confounder, pb.X, tr, te, 1 if Q else 0, K, lam, lam, alpha, tuning, 1e-12, 1e-5, iters,
                                perm_mode=1, seed=77)
        np.savez_compressed(os.path.join(out, f"{name}.npz"), V=r["column_factor"], loss=r["loss"], train_rmse=r["train_rmse"],
                            test_rmse=r["test_rmse"], iters_run=r["iters_run"], cd_sweeps=r["cd_sweeps"],
                            **{f"F{i}": f for i, f in enumerate(r["row_matrices"])},
                            check_loss=np.array([c["loss"] for c in r["checks"]]))
        print(name, r["loss"], r["iters_run"], r["cd_sweeps"])
