"""TEST INFRASTRUCTURE: ctypes access to oracle/_ref/libinsider_ref.so - the REFERENCE'S OWN sources (/root/reference/src/optimize.cpp,
coordinate_descent.cpp, utils.cpp) compiled where they lie against the small Armadillo / Rcpp API shim in oracle/ref_shim/ (`make -C oracle
ref`; no R, Rcpp, Armadillo, BLAS or LAPACK in this image). Used by tests/test_ref_pin.py to pin oracle/insider_oracle.cpp; never by the product.

The library is serial (no OpenMP: the reference draws its coordinate orders from one global RNG stream inside its parallel loops) and draws
`arma::randperm` from R's Mersenne-Twister the way RcppArmadillo does (restated, like the oracle's mode A): compare against
`oracle.optimize(..., perm_mode=0, r_seed=s)`."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libinsider_ref.so")
REFERENCE = os.environ.get("INSIDER_REFERENCE", "/root/reference")
_lib = None


def build(force: bool = False) -> str | None:
    """Compile the reference's sources if they are present (this container); returns the path or None."""
    if not os.path.isdir(os.path.join(REFERENCE, "src")):
        return SO if os.path.exists(SO) else None
    if force and os.path.exists(SO):
        os.remove(SO)
    env = dict(os.environ)
    env.pop("CXX", None)
    r = subprocess.run(["make", "-C", _HERE, "ref", f"REF={REFERENCE}"], env=env, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle/_ref build failed:\n" + r.stdout[-3000:] + r.stderr[-3000:])
    return SO


def available() -> bool:
    return os.path.exists(SO) or os.path.isdir(os.path.join(REFERENCE, "src"))


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            build()
        _lib = C.CDLL(SO)
        _lib.ref_randperm_calls.restype = C.c_ulonglong
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def optimize(data, cfd_factors, column_factor, cfd_indicators, ctns_confounder, train_indicator, test_indicator, inc_continuous, latent_dim,
             lambda1=1.0, lambda2=1.0, alpha=0.1, tuning=1, global_tol=1e-10, sub_tol=1e-5, max_iter=10000, r_seed=1):
    """The reference's optimize() (src/optimize.cpp:256). Returns (factors, column_factor, train_rmse, test_rmse, loss)."""
    Y = np.asfortranarray(data, dtype=np.float64)
    N, P = Y.shape
    F = [np.array(f, dtype=np.float64, order="F", copy=True) for f in cfd_factors]
    V = np.array(column_factor, dtype=np.float64, order="F", copy=True)
    K = int(latent_dim)
    Z = np.asfortranarray(np.asarray(cfd_indicators).reshape(N, -1), dtype=np.uint32)
    X = np.asfortranarray(np.asarray(ctns_confounder, dtype=np.float64).reshape(N, -1)) if ctns_confounder is not None else None
    Q = X.shape[1] if X is not None else 0
    tr = np.asfortranarray(train_indicator, dtype=np.float64) if train_indicator is not None else None       # the reference's mat indicators
    te = np.asfortranarray(test_indicator, dtype=np.float64) if test_indicator is not None else None
    nf = len(F)
    fptr = (C.POINTER(C.c_double) * nf)(*[_dp(f) for f in F])
    frows = (C.c_int * nf)(*[f.shape[0] for f in F])
    out3 = np.zeros(3)
    rc = lib().ref_optimize(C.c_int(N), C.c_int(P), C.c_int(K), _dp(Y), C.c_int(Z.shape[1]), Z.ctypes.data_as(C.POINTER(C.c_uint)), C.c_int(Q), _dp(X),
                            C.c_int(int(inc_continuous)), _dp(tr), _dp(te), C.c_int(nf), fptr, frows, _dp(V), C.c_double(lambda1), C.c_double(lambda2),
                            C.c_double(alpha), C.c_int(int(tuning)), C.c_double(global_tol), C.c_double(sub_tol), C.c_uint(int(max_iter)),
                            C.c_uint(int(r_seed)), _dp(out3))
    if rc:
        raise RuntimeError(f"ref_optimize rc={rc}")
    return F, V, float(out3[0]), float(out3[1]), float(out3[2])


def strong_cd(X, y, wstart, lam, alpha, XtX, Xty, tol=1e-5, r_seed=1):
    """The reference's strong_coordinate_descent() (src/coordinate_descent.cpp:57). Returns (beta, sweeps)."""
    X = np.asfortranarray(X, dtype=np.float64)
    n, K = X.shape
    y = np.ascontiguousarray(y, dtype=np.float64)
    w = np.ascontiguousarray(wstart, dtype=np.float64)
    G = np.asfortranarray(XtX, dtype=np.float64)
    b = np.ascontiguousarray(Xty, dtype=np.float64)
    beta = np.empty(K)
    sw = C.c_ulonglong()
    lib().ref_strong_cd(C.c_int(n), C.c_int(K), _dp(X), _dp(y), _dp(w), C.c_double(lam), C.c_double(alpha), _dp(G), _dp(b), C.c_double(tol),
                        C.c_uint(int(r_seed)), _dp(beta), C.byref(sw))
    return beta, int(sw.value)


def optimize_continuous_v2(data, indicator, updating_factor, c_factor, updating_confd, gram, lam, tuning):
    """The reference's optimize_continuous_v2() (src/optimize.cpp:77). Returns the updated factor (K,)."""
    Y = np.asfortranarray(data, dtype=np.float64)
    N, P = Y.shape
    V = np.asfortranarray(c_factor, dtype=np.float64)
    K = V.shape[0]
    w = np.array(updating_factor, dtype=np.float64).reshape(-1).copy()
    x = np.ascontiguousarray(updating_confd, dtype=np.float64).reshape(-1)
    ind = np.asfortranarray(indicator, dtype=np.float64)
    g = np.asfortranarray(gram, dtype=np.float64)
    lib().ref_optimize_continuous_v2(C.c_int(N), C.c_int(P), C.c_int(K), _dp(Y), _dp(ind), _dp(w), _dp(V), _dp(x), _dp(g), C.c_double(lam), C.c_int(int(tuning)))
    return w


def r_unif(seed, n):
    out = np.empty(n)
    lib().ref_unif(C.c_uint(int(seed)), C.c_int(n), _dp(out))
    return out
