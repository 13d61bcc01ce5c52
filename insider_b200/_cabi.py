"""ctypes binding of libinsider_b200.so (include/insider_b200.h). No CPU fallback: if the library or a B200 is
missing, calls raise."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("INSIDER_B200_LIB") or os.path.join(_HERE, "lib", "libinsider_b200.so")   # env override: A/B builds

OK, ERR_INVALID_ARG, ERR_CUDA, ERR_NCCL, ERR_NOT_SPD, ERR_DIVERGED, ERR_EMPTY_TEST_SET, ERR_NOMEM, ERR_UNSUPPORTED = range(9)
MASK_NONE, MASK_INT32, MASK_UINT8, MASK_DOUBLE = range(4)
PERM_COUNTER, PERM_IDENTITY = 1, 2

# every symbol include/insider_b200.h declares
EXPORTED = [
    "insider_b200_default_options", "insider_b200_version", "insider_b200_ctx_create", "insider_b200_nccl_unique_id",
    "insider_b200_ctx_create_dist", "insider_b200_ctx_destroy", "insider_b200_ctx_stream", "insider_b200_optimize",
    "insider_b200_upload", "insider_b200_release", "insider_b200_optimize_resident", "insider_b200_als_begin",
    "insider_b200_als_step", "insider_b200_als_read", "insider_b200_als_end", "insider_b200_als_profile", "insider_b200_als_sweeps", "insider_b200_als_hint_sweeps",
    "insider_b200_set_profile", "insider_b200_strong_cd", "insider_b200_fit_interaction", "insider_b200_split",
    "insider_b200_tune_batch", "insider_b200_optimize_continuous", "insider_b200_glm_interaction",
]


class InsiderError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"insider_b200 error {code}: {msg}")
        self.code = code
        self.msg = msg


class Problem(C.Structure):
    _fields_ = [("N", C.c_int64), ("P", C.c_int64), ("C", C.c_int32), ("Q", C.c_int32), ("inc_continuous", C.c_int32),
                ("mask_kind", C.c_int32), ("Y", C.c_void_p), ("levels", C.c_void_p), ("X", C.c_void_p), ("train", C.c_void_p),
                ("test", C.c_void_p)]


class Factors(C.Structure):
    _fields_ = [("K", C.c_int32), ("n_factors", C.c_int32), ("factors", C.POINTER(C.c_void_p)), ("factor_rows", C.POINTER(C.c_int32)),
                ("column_factor", C.c_void_p)]


class Options(C.Structure):
    _fields_ = [("lambda1", C.c_double), ("lambda2", C.c_double), ("alpha", C.c_double), ("tuning", C.c_int32), ("perm_mode", C.c_int32),
                ("global_tol", C.c_double), ("sub_tol", C.c_double), ("max_iter", C.c_uint32), ("check_every", C.c_uint32),
                ("seed", C.c_uint64), ("verbose", C.c_int32), ("use_graph", C.c_int32)]


class Check(C.Structure):
    _fields_ = [("iter", C.c_int32), ("pad", C.c_int32), ("sum_residual", C.c_double), ("train_rmse", C.c_double), ("test_rmse", C.c_double),
                ("row_reg", C.c_double), ("col_reg", C.c_double), ("l1_reg", C.c_double), ("loss", C.c_double), ("delta_loss", C.c_double),
                ("decay", C.c_double)]


class Result(C.Structure):
    _fields_ = [("train_rmse", C.c_double), ("test_rmse", C.c_double), ("loss", C.c_double), ("iters_run", C.c_uint32),
                ("n_checks", C.c_uint32), ("checks", C.POINTER(Check)), ("max_checks", C.c_uint32), ("cd_sweeps", C.c_int64),
                ("loop_ms", C.c_double), ("h2d_bytes", C.c_double), ("d2h_bytes", C.c_double), ("kernel_launches", C.c_int64), ("cd_steps", C.c_int64)]


_lib = None


def _declare(L):
    """argtypes / restype of every entry point of include/insider_b200.h (without them ctypes passes Python ints as 32-bit
    C ints, which is wrong for size_t / int64_t arguments that travel on the stack)."""
    vp, i32, i64, u32, u64, dbl, sz, cp = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_double, C.c_size_t, C.c_char_p
    P = C.POINTER
    err = (cp, sz)
    sig = {
        "insider_b200_default_options": (None, (P(Options),)),
        "insider_b200_version": (C.c_int, ()),
        "insider_b200_ctx_create": (C.c_int, (P(vp), C.c_int) + err),
        "insider_b200_nccl_unique_id": (C.c_int, (vp,) + err),
        "insider_b200_ctx_create_dist": (C.c_int, (P(vp), C.c_int, C.c_int, C.c_int, vp) + err),
        "insider_b200_ctx_destroy": (None, (vp,)),
        "insider_b200_ctx_stream": (vp, (vp,)),
        "insider_b200_optimize": (C.c_int, (vp, P(Problem), P(Factors), P(Options), P(Result)) + err),
        "insider_b200_upload": (C.c_int, (vp, P(Problem), P(vp)) + err),
        "insider_b200_release": (None, (vp,)),
        "insider_b200_optimize_resident": (C.c_int, (vp, vp, P(Factors), P(Options), P(Result)) + err),
        "insider_b200_als_begin": (C.c_int, (vp, vp, P(Factors), P(Options), P(vp)) + err),
        "insider_b200_als_step": (C.c_int, (vp, u32, P(i32), P(dbl)) + err),
        "insider_b200_als_read": (C.c_int, (vp, P(Factors)) + err),
        "insider_b200_als_end": (C.c_int, (vp, P(Factors), P(Result)) + err),
        "insider_b200_als_profile": (C.c_int, (vp, cp, sz, P(dbl), P(i64), C.c_int)),
        "insider_b200_set_profile": (None, (vp, C.c_int)),
        "insider_b200_als_sweeps": (i64, (vp, P(i32), i64)),
        "insider_b200_als_hint_sweeps": (i64, (vp, P(i32), i64)),
        "insider_b200_strong_cd": (C.c_int, (vp, i32, i64, vp, i32, vp, vp, dbl, dbl, dbl, i32, u64, u32, u64, vp, vp) + err),
        "insider_b200_fit_interaction": (C.c_int, (vp, i64, i64, i32, vp, i32, vp, vp, i32, vp, vp, i32) + err),
        "insider_b200_split": (C.c_int, (vp, i64, i64, dbl, u32, vp, vp, vp, P(i64)) + err),
        "insider_b200_tune_batch": (C.c_int, (i32, P(vp), P(vp), i32, P(Factors), P(Options), P(Result), P(i32)) + err),
        "insider_b200_optimize_continuous": (C.c_int, (vp, i64, i64, i32, vp, i32, vp, vp, vp, vp, dbl, i32) + err),
        "insider_b200_glm_interaction": (C.c_int, (vp, i64, i64, i32, vp, i32, vp, vp, vp, vp) + err),
    }
    assert sorted(sig) == sorted(EXPORTED)
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = list(args)


def lib() -> C.CDLL:
    """Loads the CUDA library; raises if it has not been built (there is no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise InsiderError(ERR_CUDA, f"{LIB_PATH} not built; run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LIB_PATH)
        for name in EXPORTED:
            getattr(L, name)
        _declare(L)
        _lib = L
    return _lib


def _err():
    return C.create_string_buffer(512)


def _chk(rc, buf):
    if rc != OK:
        raise InsiderError(rc, buf.value.decode(errors="replace"))


def default_options() -> Options:
    o = Options()
    lib().insider_b200_default_options(C.byref(o))
    return o


def mask_kind_of(a) -> int:
    if a is None:
        return MASK_NONE
    if a.dtype == np.int32:
        return MASK_INT32
    if a.dtype in (np.uint8, np.bool_):
        return MASK_UINT8
    if a.dtype == np.float64:
        return MASK_DOUBLE
    raise TypeError(f"unsupported mask dtype {a.dtype}")


class HostProblem:
    """Keeps the numpy arrays alive and builds the insider_problem struct."""

    def __init__(self, Y, levels, X=None, train=None, test=None, inc_continuous=0):
        self.Y = np.asfortranarray(Y, dtype=np.float64)
        N, P = self.Y.shape
        self.levels = np.asfortranarray(np.asarray(levels).reshape(N, -1), dtype=np.int32) if levels is not None else np.zeros((N, 0), np.int32, order="F")
        self.X = np.asfortranarray(np.asarray(X, dtype=np.float64).reshape(N, -1)) if (X is not None and inc_continuous) else None
        if train is not None:
            dt = train.dtype if train.dtype in (np.int32, np.uint8, np.float64) else (np.uint8 if train.dtype == np.bool_ else np.int32)
            self.train = np.asfortranarray(train, dtype=dt)
            self.test = np.asfortranarray(test, dtype=dt)
        else:
            self.train = self.test = None
        # shapes are the caller's contract in C (plain pointers): check them here, where they are still known
        if self.levels.shape[0] != N:
            raise ValueError(f"levels has {self.levels.shape[0]} rows, data has {N}")
        if self.X is not None and self.X.shape[0] != N:
            raise ValueError(f"ctns_confounder has {self.X.shape[0]} rows, data has {N}")
        for name, m in (("train_indicator", self.train), ("test_indicator", self.test)):
            if m is not None and m.shape != (N, P):
                raise ValueError(f"{name} is {m.shape[0]} x {m.shape[1]} but data is {N} x {P} (ratio_splitter(rm.na.col = TRUE) drops "
                                 "all-zero columns from the indicators only: drop them from data too)")
        self.N, self.P = N, P
        p = Problem()
        p.N, p.P, p.C = N, P, self.levels.shape[1]
        p.Q = self.X.shape[1] if self.X is not None else 0
        p.inc_continuous = int(inc_continuous)
        p.mask_kind = mask_kind_of(self.train)
        p.Y = self.Y.ctypes.data
        p.levels = self.levels.ctypes.data if self.levels.size else None
        p.X = self.X.ctypes.data if self.X is not None else None
        p.train = self.train.ctypes.data if self.train is not None else None
        p.test = self.test.ctypes.data if self.test is not None else None
        self.struct = p


class HostFactors:
    """In/out factor buffers (copied on construction; read `.factors` / `.V` after a call)."""

    def __init__(self, cfd_factors, column_factor, K):
        self.factors = [np.array(f, dtype=np.float64, order="F", copy=True) for f in cfd_factors]
        self.V = np.array(column_factor, dtype=np.float64, order="F", copy=True)
        n = len(self.factors)
        K = int(K)
        if self.V.ndim != 2 or self.V.shape[0] != K:
            raise ValueError(f"column_factor must be latent_dim x P = {K} x P, got {self.V.shape}")
        for i, f in enumerate(self.factors):
            if f.ndim != 2 or f.shape[1] != K:
                raise ValueError(f"cfd_factors[{i}] must have latent_dim = {K} columns, got {f.shape}")
        self._ptrs = (C.c_void_p * n)(*[f.ctypes.data for f in self.factors])
        self._rows = (C.c_int32 * n)(*[f.shape[0] for f in self.factors])
        s = Factors()
        s.K, s.n_factors = int(K), n
        s.factors = C.cast(self._ptrs, C.POINTER(C.c_void_p))
        s.factor_rows = C.cast(self._rows, C.POINTER(C.c_int32))
        s.column_factor = self.V.ctypes.data
        self.struct = s


def make_result(max_checks: int):
    buf = (Check * max(1, max_checks))()
    r = Result()
    r.checks = C.cast(buf, C.POINTER(Check))
    r.max_checks = max(1, max_checks)
    return r, buf


def result_dict(r: Result, buf) -> dict:
    checks = [{k: getattr(buf[i], k) for k, _ in Check._fields_ if k != "pad"} for i in range(r.n_checks)]
    return dict(train_rmse=r.train_rmse, test_rmse=r.test_rmse, loss=r.loss, iters_run=r.iters_run, checks=checks, cd_sweeps=r.cd_sweeps,
                loop_ms=r.loop_ms, h2d_bytes=r.h2d_bytes, d2h_bytes=r.d2h_bytes, kernel_launches=r.kernel_launches, cd_steps=r.cd_steps)


class Context:
    def __init__(self, device: int = 0, rank: int = 0, world: int = 1, nccl_id: bytes | None = None):
        self.h = C.c_void_p()
        e = _err()
        if world > 1:
            _chk(lib().insider_b200_ctx_create_dist(C.byref(self.h), device, rank, world, nccl_id, e, len(e)), e)
        else:
            _chk(lib().insider_b200_ctx_create(C.byref(self.h), device, e, len(e)), e)
        self.rank, self.world, self.device = rank, world, device

    @staticmethod
    def nccl_unique_id() -> bytes:
        b = C.create_string_buffer(128)
        e = _err()
        _chk(lib().insider_b200_nccl_unique_id(b, e, len(e)), e)
        return b.raw

    def set_profile(self, on: bool):
        lib().insider_b200_set_profile(self.h, int(on))

    def close(self):
        if self.h:
            lib().insider_b200_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- one-shot ---------------------------------------------------------------------------------
    def optimize(self, prob: HostProblem, fac: HostFactors, opt: Options) -> dict:
        r, buf = make_result(int(opt.max_iter) // max(1, int(opt.check_every) or 10) + 3)
        e = _err()
        _chk(lib().insider_b200_optimize(self.h, C.byref(prob.struct), C.byref(fac.struct), C.byref(opt), C.byref(r), e, len(e)), e)
        return result_dict(r, buf)

    def upload(self, prob: HostProblem) -> "Resident":
        return Resident(self, prob)

    def strong_cd(self, XtX, Xty, wstart, lam, alpha, tol=1e-5, perm_mode=PERM_COUNTER, seed=0, als_iter=0, gene0=0):
        Xty = np.asfortranarray(Xty, dtype=np.float64)
        if Xty.ndim == 1:
            Xty = Xty.reshape(-1, 1, order="F")
        K, n = Xty.shape
        w = np.asfortranarray(np.asarray(wstart, dtype=np.float64).reshape(K, n, order="F"))
        G = np.asarray(XtX, dtype=np.float64)
        shared = G.ndim == 2
        G = np.ascontiguousarray(G) if shared else np.ascontiguousarray(G)   # [n][K][K] symmetric: layout-agnostic
        beta = np.zeros((K, n), order="F")
        sweeps = np.zeros(n, dtype=np.int32)
        e = _err()
        _chk(lib().insider_b200_strong_cd(self.h, C.c_int32(K), C.c_int64(n), C.c_void_p(G.ctypes.data), C.c_int32(int(shared)),
                                          C.c_void_p(Xty.ctypes.data), C.c_void_p(w.ctypes.data), C.c_double(lam), C.c_double(alpha),
                                          C.c_double(tol), C.c_int32(perm_mode), C.c_uint64(seed), C.c_uint32(als_iter), C.c_uint64(gene0),
                                          C.c_void_p(beta.ctypes.data), C.c_void_p(sweeps.ctypes.data), e, len(e)), e)
        return beta, sweeps

    def fit_interaction(self, residual, train, n_levels, indicator, column_factor, tuning):
        R = np.asfortranarray(residual, dtype=np.float64)
        N, P = R.shape
        V = np.asfortranarray(column_factor, dtype=np.float64)
        K = V.shape[0]
        tr = None
        if train is not None:
            tr = np.asfortranarray(train, dtype=train.dtype if train.dtype in (np.int32, np.uint8, np.float64) else np.int32)
        z = np.ascontiguousarray(indicator, dtype=np.int32)
        out = np.zeros((n_levels, K), order="F")
        e = _err()
        _chk(lib().insider_b200_fit_interaction(self.h, C.c_int64(N), C.c_int64(P), C.c_int32(K), C.c_void_p(R.ctypes.data),
                                                C.c_int32(mask_kind_of(tr)), C.c_void_p(tr.ctypes.data if tr is not None else None),
                                                C.c_void_p(out.ctypes.data), C.c_int32(n_levels), C.c_void_p(z.ctypes.data),
                                                C.c_void_p(V.ctypes.data), C.c_int32(tuning), e, len(e)), e)
        return out


    def optimize_continuous(self, data, indicator, updating_factor, c_factor, updating_confd, lam, tuning):
        """optimize_continuous_v2 (src/optimize.cpp:77-137): returns the updated factor row (K,)."""
        Y = np.asfortranarray(data, dtype=np.float64)
        N, P = Y.shape
        V = np.asfortranarray(c_factor, dtype=np.float64)
        K = V.shape[0]
        w = np.array(updating_factor, dtype=np.float64).reshape(-1).copy()
        x = np.ascontiguousarray(updating_confd, dtype=np.float64).reshape(-1)
        if V.shape[1] != P or w.size != K or x.size != N:
            raise ValueError("optimize_continuous: shapes must be data N x P, c_factor K x P, updating_factor K, updating_confd N")
        ind = None
        if indicator is not None:
            ind = np.asfortranarray(indicator, dtype=indicator.dtype if indicator.dtype in (np.int32, np.uint8, np.float64) else np.int32)
            if ind.shape != (N, P):
                raise ValueError("optimize_continuous: indicator must be N x P")
        e = _err()
        _chk(lib().insider_b200_optimize_continuous(self.h, N, P, K, Y.ctypes.data, mask_kind_of(ind), ind.ctypes.data if ind is not None else None,
                                                    w.ctypes.data, V.ctypes.data, x.ctypes.data, float(lam), int(tuning), e, len(e)), e)
        return w

    def glm_interaction(self, residual, interaction_indicator, column_factor):
        """glm_interaction (R/glm_interaction.R:2-30): returns (coeff_matrix, pval_matrix), each n_levels x K."""
        R = np.asfortranarray(residual, dtype=np.float64)
        N, P = R.shape
        V = np.asfortranarray(column_factor, dtype=np.float64)
        K = V.shape[0]
        z = np.ascontiguousarray(interaction_indicator, dtype=np.int32).reshape(-1)
        if V.shape[1] != P or z.size != N:
            raise ValueError("glm_interaction: shapes must be residual N x P, column_factor K x P, interaction_indicator N")
        L = int(z.max())
        coeff = np.zeros((L, K), order="F")
        pval = np.zeros((L, K), order="F")
        e = _err()
        _chk(lib().insider_b200_glm_interaction(self.h, N, P, K, R.ctypes.data, L, z.ctypes.data, V.ctypes.data, coeff.ctypes.data, pval.ctypes.data,
                                                e, len(e)), e)
        return coeff, pval


def tune_batch(residents, facs, opts):
    """insider_b200_tune_batch: the grid points (facs[i], opts[i]) as replicas over the contexts of `residents` (one resident copy of
    the same problem per context). Returns (list of result dicts, context index that ran each point)."""
    n_ctx, n = len(residents), len(facs)
    ctxs = (C.c_void_p * n_ctx)(*[r.ctx.h for r in residents])
    ress = (C.c_void_p * n_ctx)(*[r.h for r in residents])
    F = (Factors * n)(*[f.struct for f in facs])
    O = (Options * n)(*opts)
    R = (Result * n)()
    bufs = []
    for i in range(n):
        mc = int(opts[i].max_iter) // max(1, int(opts[i].check_every) or 10) + 3
        buf = (Check * mc)()
        R[i].checks = C.cast(buf, C.POINTER(Check))
        R[i].max_checks = mc
        bufs.append(buf)
    who = (C.c_int32 * max(1, n))()
    e = _err()
    _chk(lib().insider_b200_tune_batch(n_ctx, ctxs, ress, n, F, O, R, who, e, len(e)), e)
    return [result_dict(R[i], bufs[i]) for i in range(n)], [int(who[i]) for i in range(n)]


class Resident:
    def __init__(self, ctx: Context, prob: HostProblem):
        self.ctx = ctx
        self.prob = prob
        self.h = C.c_void_p()
        self.owner = True
        e = _err()
        _chk(lib().insider_b200_upload(ctx.h, C.byref(prob.struct), C.byref(self.h), e, len(e)), e)

    @classmethod
    def share(cls, ctx: Context, other: "Resident") -> "Resident":
        """The same device-resident problem seen from another context ON THE SAME DEVICE (a resident problem is read-only
        after upload): several contexts = several streams whose fits overlap (tune() replicas on one GPU). Not an owner."""
        if ctx.device != other.ctx.device or ctx.world != 1:
            raise ValueError("a resident problem can only be shared between single-GPU contexts on the same device")
        r = cls.__new__(cls)
        r.ctx, r.prob, r.h, r.owner = ctx, other.prob, other.h, False
        return r

    def optimize(self, fac: HostFactors, opt: Options) -> dict:
        r, buf = make_result(int(opt.max_iter) // max(1, int(opt.check_every) or 10) + 3)
        e = _err()
        _chk(lib().insider_b200_optimize_resident(self.ctx.h, self.h, C.byref(fac.struct), C.byref(opt), C.byref(r), e, len(e)), e)
        return result_dict(r, buf)

    def begin(self, fac: HostFactors, opt: Options) -> "Session":
        return Session(self, fac, opt)

    def release(self):
        if self.h and self.owner:
            lib().insider_b200_release(self.h)
        self.h = C.c_void_p()

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class Session:
    def __init__(self, res: Resident, fac: HostFactors, opt: Options):
        self.res, self.fac, self.opt = res, fac, opt
        self.h = C.c_void_p()
        e = _err()
        _chk(lib().insider_b200_als_begin(res.ctx.h, res.h, C.byref(fac.struct), C.byref(opt), C.byref(self.h), e, len(e)), e)

    def step(self, n_iters: int = 1):
        """Runs up to n_iters iterations; returns (done, device_ms)."""
        done, ms = C.c_int32(), C.c_double()
        e = _err()
        _chk(lib().insider_b200_als_step(self.h, C.c_uint32(n_iters), C.byref(done), C.byref(ms), e, len(e)), e)
        return bool(done.value), ms.value

    def read(self):
        e = _err()
        _chk(lib().insider_b200_als_read(self.h, C.byref(self.fac.struct), e, len(e)), e)
        return [f.copy() for f in self.fac.factors], self.fac.V.copy()

    def profile(self) -> dict:
        names = C.create_string_buffer(4096)
        ms = (C.c_double * 64)()
        calls = (C.c_int64 * 64)()
        n = lib().insider_b200_als_profile(self.h, names, len(names), ms, calls, 64)
        ks = names.value.decode().split("\n") if n else []
        return {ks[i]: (ms[i], calls[i]) for i in range(n)}

    def sweeps(self, n: int) -> np.ndarray:
        """Per-gene coordinate-descent sweep counts of the last iteration (dense elastic-net path; diagnostics)."""
        out = np.zeros(n, dtype=np.int32)
        got = lib().insider_b200_als_sweeps(self.h, out.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int64(n))
        return out[:got]

    def hint_sweeps(self, hint) -> int:
        """Expected sweep counts of the next iteration (orders the dense solver's work; results do not depend on it)."""
        h = np.ascontiguousarray(hint, dtype=np.int32)
        return int(lib().insider_b200_als_hint_sweeps(self.h, h.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int64(h.size)))

    def end(self, read_factors: bool = True) -> dict:
        r, buf = make_result(int(self.opt.max_iter) // max(1, int(self.opt.check_every) or 10) + 3)
        e = _err()
        h, self.h = self.h, C.c_void_p()
        _chk(lib().insider_b200_als_end(h, C.byref(self.fac.struct) if read_factors else None, C.byref(r), e, len(e)), e)
        return result_dict(r, buf)

    def __del__(self):
        try:
            if self.h:
                self.end(read_factors=False)
        except Exception:
            pass


def split(data, ratio=0.1, seed=123):
    """insider_b200_split: R-exact ratio_splitter masks -> (train, test, na) int32 N x P (Fortran order)."""
    Y = np.asfortranarray(data, dtype=np.float64)
    N, P = Y.shape
    tr = np.zeros((N, P), dtype=np.int32, order="F")
    te = np.zeros((N, P), dtype=np.int32, order="F")
    na = np.zeros((N, P), dtype=np.int32, order="F")
    nt = C.c_int64()
    e = _err()
    _chk(lib().insider_b200_split(C.c_void_p(Y.ctypes.data), C.c_int64(N), C.c_int64(P), C.c_double(ratio), C.c_uint32(seed),
                                  C.c_void_p(tr.ctypes.data), C.c_void_p(te.ctypes.data), C.c_void_p(na.ctypes.data), C.byref(nt), e, len(e)), e)
    return tr, te, na
