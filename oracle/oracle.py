"""ORACLE — TEST INFRASTRUCTURE ONLY. ctypes binding of oracle/liboracle.so (oracle/insider_oracle.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OracleCheck(C.Structure):
    _fields_ = [("iter", C.c_int32), ("pad", C.c_int32), ("sum_residual", C.c_double), ("train_rmse", C.c_double),
                ("test_rmse", C.c_double), ("row_reg", C.c_double), ("col_reg", C.c_double), ("l1_reg", C.c_double),
                ("loss", C.c_double), ("delta_loss", C.c_double), ("decay", C.c_double)]


def build(force: bool = False, native: bool = False) -> str:
    """liboracle.so (x86-64-v3: built in the CPU container, runs on any host of the pool) or, native=True,
    liboracle_native.so compiled with -march=native ON THE MACHINE THAT RUNS IT (bench.py's timed CPU legs)."""
    name = "liboracle_native.so" if native else "liboracle.so"
    so = os.path.join(_HERE, name)
    src = os.path.join(_HERE, "insider_oracle.cpp")
    if native:
        # always rebuilt where it runs: a -march=native object from another host may use instructions this one lacks
        env = dict(os.environ); env.pop("CXX", None); env.pop("CC", None)
        subprocess.run(["make", "-B", "-C", _HERE, name], check=True, capture_output=True, env=env)
    elif force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        env = dict(os.environ); env.pop("CXX", None); env.pop("CC", None)
        subprocess.run(["make", "-C", _HERE, name], check=True, capture_output=True, env=env)
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = _load(so)
    return _LIB


def _load(so):
    L = C.CDLL(so)
    L.oracle_optimize.restype = C.c_int
    L.oracle_strong_cd.restype = C.c_int
    L.oracle_fit_interaction.restype = C.c_int
    L.oracle_optimize_continuous_v2.restype = C.c_int
    L.oracle_set_threads.restype = C.c_int
    return L


def use_native() -> bool:
    """Switches this process to the -march=native build (compiled now, on this host). Returns False (and keeps the portable
    build) when the compiler is missing or fails."""
    global _LIB
    try:
        _LIB = _load(build(native=True))
        return True
    except Exception:  # noqa: BLE001
        return False


def set_threads(n: int) -> int:
    """omp_set_num_threads(n) inside the oracle; returns omp_get_max_threads() afterwards (the team size really used)."""
    return int(lib().oracle_set_threads(C.c_int(int(n))))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None


@dataclass
class OracleResult:
    factors: list
    column_factor: np.ndarray
    train_rmse: float
    test_rmse: float
    loss: float
    iters_run: int
    checks: list = field(default_factory=list)
    cd_sweeps: int = 0
    seconds_in_loop: float = 0.0


def optimize(data, cfd_factors, column_factor, cfd_indicators, ctns_confounder, train_indicator, test_indicator,
             inc_continuous, latent_dim, lambda1=1.0, lambda2=1.0, alpha=0.1, tuning=1, global_tol=1e-10,
             sub_tol=1e-5, max_iter=10000, perm_mode=1, seed=0, r_seed=1, n_cores_row=0, n_cores_col=0) -> OracleResult:
    """Mirror of the reference's ``optimize()`` (R/RcppExports.R:20-22). Inputs are copied; results returned."""
    Y = np.asfortranarray(data, dtype=np.float64)
    N, P = Y.shape
    F = [np.array(f, dtype=np.float64, order="F", copy=True) for f in cfd_factors]
    V = np.array(column_factor, dtype=np.float64, order="F", copy=True)
    Z = np.asfortranarray(np.asarray(cfd_indicators).reshape(N, -1), dtype=np.int32)
    Cn = Z.shape[1]
    X = np.asfortranarray(np.asarray(ctns_confounder, dtype=np.float64).reshape(N, -1)) if ctns_confounder is not None else np.zeros((N, 1), order="F")
    Q = X.shape[1]
    tr = np.asfortranarray(train_indicator, dtype=np.int32) if train_indicator is not None else None
    te = np.asfortranarray(test_indicator, dtype=np.int32) if test_indicator is not None else None
    nf = len(F)
    fptr = (C.POINTER(C.c_double) * nf)(*[_dp(f) for f in F])
    frows = (C.c_int * nf)(*[f.shape[0] for f in F])
    max_checks = int(max_iter) // 10 + 3
    checks = (OracleCheck * max_checks)()
    trm, tem, loss, secs = C.c_double(), C.c_double(), C.c_double(), C.c_double()
    iters, nchk, sweeps = C.c_int(), C.c_int(), C.c_longlong()
    rc = lib().oracle_optimize(
        C.c_int(N), C.c_int(P), _dp(Y), C.c_int(nf), fptr, frows, _dp(V), C.c_int(Cn), _ip(Z), C.c_int(Q), _dp(X),
        _ip(tr), _ip(te), C.c_int(int(inc_continuous)), C.c_int(int(latent_dim)), C.c_double(lambda1), C.c_double(lambda2),
        C.c_double(alpha), C.c_int(int(tuning)), C.c_double(global_tol), C.c_double(sub_tol), C.c_uint(int(max_iter)),
        C.c_int(perm_mode), C.c_uint64(seed), C.c_uint32(r_seed), C.c_int(n_cores_row), C.c_int(n_cores_col),
        C.byref(trm), C.byref(tem), C.byref(loss), C.byref(iters), checks, C.c_int(max_checks), C.byref(nchk),
        C.byref(sweeps), C.byref(secs))
    if rc != 0:
        raise RuntimeError({1: "oracle: matrix not SPD", 2: "oracle: empty test set", 3: "oracle: invalid argument"}.get(rc, f"oracle rc={rc}"))
    recs = [{k: getattr(checks[i], k) for k, _ in OracleCheck._fields_ if k != "pad"} for i in range(min(nchk.value, max_checks))]
    return OracleResult(F, V, trm.value, tem.value, loss.value, iters.value, recs, sweeps.value, secs.value)


def strong_cd(X, y, wstart, lam, alpha, XtX, Xty, tol=1e-5, perm_mode=1, seed=0, als_iter=0, gene=0, r_seed=1):
    """Mirror of ``strong_coordinate_descent`` (src/coordinate_descent.cpp:57). Returns (beta, sweeps, rounds)."""
    X = np.asfortranarray(X, dtype=np.float64)
    n, K = X.shape
    y = np.ascontiguousarray(y, dtype=np.float64)
    w = np.ascontiguousarray(wstart, dtype=np.float64)
    G = np.asfortranarray(XtX, dtype=np.float64)
    b = np.ascontiguousarray(Xty, dtype=np.float64)
    beta = np.empty(K)
    sw, rd = C.c_int(), C.c_int()
    lib().oracle_strong_cd(C.c_int(n), C.c_int(K), _dp(X), _dp(y), _dp(w), C.c_double(lam), C.c_double(alpha), _dp(G), _dp(b),
                           C.c_double(tol), C.c_int(perm_mode), C.c_uint64(seed), C.c_uint32(als_iter), C.c_uint64(gene),
                           C.c_uint32(r_seed), _dp(beta), C.byref(sw), C.byref(rd))
    return beta, sw.value, rd.value


def optimize_continuous_v2(data, indicator, updating_factor, c_factor, updating_confd, gram, lam, tuning):
    """Mirror of ``optimize_continuous_v2`` (src/optimize.cpp:77-137). Returns the updated factor (K,)."""
    Y = np.asfortranarray(data, dtype=np.float64)
    N, P = Y.shape
    V = np.asfortranarray(c_factor, dtype=np.float64)
    K = V.shape[0]
    w = np.array(updating_factor, dtype=np.float64).reshape(-1).copy()
    x = np.ascontiguousarray(updating_confd, dtype=np.float64).reshape(-1)
    ind = np.asfortranarray(indicator, dtype=np.int32) if indicator is not None else None
    g = np.asfortranarray(gram, dtype=np.float64) if gram is not None else None
    rc = lib().oracle_optimize_continuous_v2(C.c_int(N), C.c_int(P), C.c_int(K), _dp(Y), _ip(ind), _dp(w), _dp(V), _dp(x), _dp(g),
                                             C.c_double(lam), C.c_int(tuning))
    if rc:
        raise RuntimeError(f"oracle_optimize_continuous_v2 rc={rc}")
    return w


def randperm_b(seed, als_iter, gene, draw, n):
    out = np.empty(n, dtype=np.int32)
    lib().oracle_randperm_b(C.c_uint64(seed), C.c_uint32(als_iter), C.c_uint64(gene), C.c_uint32(draw), C.c_int(n), _ip(out))
    return out


def randperm_b_inc(seed, als_iter, draw, K, inc):
    inc = np.ascontiguousarray(inc, dtype=np.int32)
    out = np.empty(inc.size, dtype=np.int32)
    lib().oracle_randperm_b_inc(C.c_uint64(seed), C.c_uint32(als_iter), C.c_uint32(draw), C.c_int(K), _ip(inc), C.c_int(inc.size), _ip(out))
    return out


def randperm_r(r_seed, n):
    out = np.empty(n, dtype=np.int32)
    lib().oracle_randperm_r(C.c_uint32(r_seed), C.c_int(n), _ip(out))
    return out


def r_unif(seed, n):
    out = np.empty(n)
    lib().oracle_r_unif(C.c_uint32(seed), C.c_int(n), _dp(out))
    return out


def chol_solve(A, b):
    A = np.array(A, dtype=np.float64, order="F", copy=True)
    b = np.array(b, dtype=np.float64, order="F", copy=True)
    nrhs = 1 if b.ndim == 1 else b.shape[1]
    rc = lib().oracle_chol_solve(C.c_int(A.shape[0]), _dp(A), _dp(b), C.c_int(nrhs))
    if rc:
        raise RuntimeError("not SPD")
    return b


def fit_interaction(residual, train_indicator, n_levels, interaction_indicator, column_factor, tuning):
    """Restated math of src/fit_interaction.cpp:10-90 (dead code in the reference)."""
    R = np.asfortranarray(residual, dtype=np.float64)
    N, P = R.shape
    V = np.asfortranarray(column_factor, dtype=np.float64)
    K = V.shape[0]
    tr = np.asfortranarray(train_indicator, dtype=np.int32) if train_indicator is not None else None
    z = np.ascontiguousarray(interaction_indicator, dtype=np.int32)
    out = np.zeros((n_levels, K), order="F")
    rc = lib().oracle_fit_interaction(C.c_int(N), C.c_int(P), C.c_int(K), _dp(R), _ip(tr), _dp(out), C.c_int(n_levels), _ip(z), _dp(V), C.c_int(tuning))
    if rc:
        raise RuntimeError(f"oracle_fit_interaction rc={rc}")
    return out
