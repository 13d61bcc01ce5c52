// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product; only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this library.
//
// CPU restatement (C++17 + OpenMP, no third-party dependencies) of the alternating-optimisation fit of
// kai0511/insider. Every function cites the reference file:line it follows (paths relative to the
// reference checkout). The restatement keeps the reference's *structure*: a dense N x P residual that is
// added to / subtracted from per confounder block, per-level normal equations with the complement-Gram
// trick, per-gene elastic-net coordinate descent in RESIDUAL form with a loss-difference stopping rule,
// evaluation every 10th iteration with the decay ladder.
//
// PARITY PIN: the reference ships no golden vectors / known-answer tests for this path, and as shipped it cannot be built here
// (it needs R + Rcpp + RcppArmadillo + BLAS/LAPACK, none present). Its ALGORITHM sources, however, compile where they lie against
// a small Armadillo / Rcpp API shim of ours (oracle/ref_shim/, `make -C oracle ref` -> oracle/_ref/libinsider_ref.so), and
// tests/test_ref_pin.py runs this restatement (mode A, one thread) against them: optimize() on ridge / elastic-net x masked /
// dense fits, with continuous covariates, to convergence through the decay ladder, and at 377 x 320 with K = 23;
// strong_coordinate_descent() and optimize_continuous_v2() directly - factors to 1e-13..1e-15, identical sweep counts. What that
// pin does NOT cover is the arithmetic below the reference's own code: the shim's products / Cholesky stand in for BLAS / LAPACK,
// and arma::randperm under R's RNG is restated from memory on both sides (R's runif known answers hold).
// Further pins: (1) an independently written NumPy/SciPy twin (oracle/numpy_twin.py), (2) analytic optimality checks (KKT /
// normal-equation residuals / monotone loss) in tests/, (3) R-RNG known answers.
//
// Third-party arithmetic the reference delegates to and that is restated here:
//   * arma::solve(A, b, solve_opts::likely_sympd)  -> LAPACK dposv-style Cholesky (chol_solve below). The
//     reference falls back to an approximate SVD solve (with a warning) when Cholesky fails; the oracle
//     reports an error instead (documented deviation, never hit on SPD systems with lambda > 0).
//   * arma::randperm(n) under RcppArmadillo       -> n draws int(Rf_runif(0, RAND_MAX)) sorted ascending,
//     indices returned (fn_randperm / Alt_R_RNG; version unpinned). Mode A below.
//   * R's Mersenne-Twister, set.seed scrambling, unif_rand (RNG.c)  -> RRng below.
//
// Permutation modes for the coordinate order (coordinate_descent.cpp:89):
//   mode 0 (A) R-stream-faithful: one global R RNG consumed gene after gene (single-thread semantics).
//   mode 1 (B) counter-based: key (seed, als_iter, draw) - draw = index of the sweep within the gene's solve - selects one
//              of 4096 table permutations of all K coordinates, each built like randperm (sort 26-bit random keys, index
//              tie-break); a gene visits its ACTIVE coordinates in that order (a uniformly random order of the active
//              set, like randperm(|inc|)). The key does not depend on the gene, so every gene that is at sweep d of ALS
//              iteration t uses the same order: that is what lets the GPU run one gene per thread with a warp-uniform
//              coordinate. Identical on CPU and GPU; parity at scale is defined in this mode.
//   mode 2     identity order (no shuffling) - for analytic tests.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---------------------------------------------------------------------------------------------------
// R's default RNG (Mersenne-Twister) — R src/main/RNG.c (not in the reference tree; restated).
struct RRng {
    uint32_t mt[624];
    int mti;
    void set_seed(uint32_t seed) {
        for (int j = 0; j < 50; ++j) seed = 69069u * seed + 1u;          // initial scrambling
        uint32_t dummy0 = 0;
        for (int j = 0; j < 625; ++j) {                                   // RNG_Init: n_seed = 625
            seed = 69069u * seed + 1u;
            if (j == 0) dummy0 = seed; else mt[j - 1] = seed;
        }
        (void)dummy0;
        mti = 624;                                                        // FixupSeeds: dummy[0] = 624
    }
    uint32_t genrand() {
        static const uint32_t mag01[2] = {0x0u, 0x9908b0dfu};
        uint32_t y;
        if (mti >= 624) {
            int kk;
            for (kk = 0; kk < 624 - 397; ++kk) {
                y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
                mt[kk] = mt[kk + 397] ^ (y >> 1) ^ mag01[y & 1u];
            }
            for (; kk < 623; ++kk) {
                y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
                mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ mag01[y & 1u];
            }
            y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
            mt[623] = mt[396] ^ (y >> 1) ^ mag01[y & 1u];
            mti = 0;
        }
        y = mt[mti++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    double unif_rand() {
        const double i2_32m1 = 2.328306437080797e-10;
        double x = (double)genrand() * 2.3283064365386963e-10;
        if (x <= 0.0) return 0.5 * i2_32m1;
        if ((1.0 - x) <= 0.0) return 1.0 - 0.5 * i2_32m1;
        return x;
    }
};

inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

struct PermSrc {
    int mode = 1;            // 0 = A (R stream), 1 = B (counter), 2 = identity
    uint64_t seed = 0;
    uint32_t als_iter = 0;
    uint64_t gene = 0;
    uint32_t draw = 0;       // index of the randperm call within this gene's solve
    RRng* r = nullptr;
};

// arma::randperm(n) as used at coordinate_descent.cpp:89: visiting order of the n = |inc| active coordinates, as indices
// into inc (ascending list of active coordinates out of K).
void randperm(PermSrc& ps, int n, int* ord, const int* inc = nullptr, int K = 0) {
    if (n <= 0) { ps.draw++; return; }
    std::vector<std::pair<uint32_t, int>> pk(n);
    if (ps.mode == 0) {
        for (int i = 0; i < n; ++i) {
            double u = 0.0 + (2147483647.0 - 0.0) * ps.r->unif_rand();   // Rf_runif(0, RAND_MAX)
            pk[i] = {(uint32_t)(int)u, i};
        }
    } else if (ps.mode == 1) {
        // mode B: the per-sweep key selects one of PERM_T table permutations of all K coordinates; every table entry is built
        // like randperm itself (sort K 26-bit random keys ascending, ties by index) from a fixed table key. Active coordinate
        // inc[i] is visited at its rank among the active ones.
        if (!inc || K <= 0) K = n;
        const uint64_t sel = mix64(ps.seed + 0x9E3779B97F4A7C15ull * (1ull + ps.als_iter)) ^
                            mix64((uint64_t)ps.draw * 0x8CB92BA72F3D8DD7ull + 0x2545F4914F6CDD1Dull);
        const uint64_t t = (sel >> 20) & 4095ull;
        const uint64_t key = mix64(0x1F83D9ABFB41BD6Bull ^ (((uint64_t)K << 32) | t));
        std::vector<std::pair<uint32_t, int>> full(K);
        for (int c = 0; c < K; ++c) full[c] = {(uint32_t)(mix64(key + 0x9E3779B97F4A7C15ull * (uint64_t)(c + 1)) >> 38), c};
        std::stable_sort(full.begin(), full.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
        std::vector<int> rank(K);
        for (int i = 0; i < K; ++i) rank[full[i].second] = i;
        for (int i = 0; i < n; ++i) pk[i] = {(uint32_t)rank[inc ? inc[i] : i], i};
    } else {
        for (int i = 0; i < n; ++i) pk[i] = {(uint32_t)i, i};
    }
    std::stable_sort(pk.begin(), pk.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    for (int i = 0; i < n; ++i) ord[i] = pk[i].second;
    ps.draw++;
}

// LAPACK dposv-style solve (lower Cholesky, forward + back substitution). A is K x K column-major and is
// overwritten; b is overwritten with the solution. Returns 0, or j+1 when the leading minor j is not PD.
int chol_solve(int K, double* A, double* b, int nrhs = 1, int ldb = 0) {
    if (ldb == 0) ldb = K;
    for (int j = 0; j < K; ++j) {
        double d = A[j + (size_t)j * K];
        for (int m = 0; m < j; ++m) d -= A[j + (size_t)m * K] * A[j + (size_t)m * K];
        if (!(d > 0.0)) return j + 1;
        d = std::sqrt(d);
        A[j + (size_t)j * K] = d;
        for (int i = j + 1; i < K; ++i) {
            double s = A[i + (size_t)j * K];
            for (int m = 0; m < j; ++m) s -= A[i + (size_t)m * K] * A[j + (size_t)m * K];
            A[i + (size_t)j * K] = s / d;
        }
    }
    for (int r = 0; r < nrhs; ++r) {
        double* x = b + (size_t)r * ldb;
        for (int i = 0; i < K; ++i) {
            double s = x[i];
            for (int m = 0; m < i; ++m) s -= A[i + (size_t)m * K] * x[m];
            x[i] = s / A[i + (size_t)i * K];
        }
        for (int i = K - 1; i >= 0; --i) {
            double s = x[i];
            for (int m = i + 1; m < K; ++m) s -= A[m + (size_t)i * K] * x[m];
            x[i] = s / A[i + (size_t)i * K];
        }
    }
    return 0;
}

inline double sgn(double x) { return (x > 0.0) - (x < 0.0); }

// src/utils.cpp:46-49  compute_loss(residual, beta, lambda, alpha)
double compute_loss_vec(int n, const double* r, int K, const double* beta, double lambda, double alpha) {
    double ss = 0.0, sb = 0.0, ab = 0.0;
    for (int i = 0; i < n; ++i) ss += r[i] * r[i];
    for (int k = 0; k < K; ++k) { sb += beta[k] * beta[k]; ab += std::fabs(beta[k]); }
    return ss / 2 + (1 - alpha) * lambda * sb / 2 + alpha * lambda * ab;
}

struct CdStats { int sweeps = 0; int rounds = 0; };

// optional per-gene sweep-count sink (diagnostics for tools/): g_sweep_sink[als_iter * P + gene]
int* g_sweep_sink = nullptr; long long g_sweep_sink_len = 0;

// src/coordinate_descent.cpp:57-127  strong_coordinate_descent()
// X is n x K column-major (the rows of the row factor selected for this gene), y the selected outcomes.
void strong_cd(int n, int K, const double* X, const double* y, const double* wstart, double lambda, double alpha,
               const double* XtX, const double* Xty, double tol, PermSrc& ps, double* beta, CdStats* st) {
    std::vector<double> residual(n);
    std::vector<int> active(K, 1), inc, ex, ord(K);
    for (int k = 0; k < K; ++k) beta[k] = wstart[k];                                    // :68
    double maxabs = 0.0;
    for (int k = 0; k < K; ++k) maxabs = std::max(maxabs, std::fabs(Xty[k]));
    const double thr = alpha * (2 * lambda - maxabs);                                   // :74
    for (int k = 0; k < K; ++k) if (std::fabs(Xty[k]) < thr) { active[k] = 0; beta[k] = 0.0; }   // :75-78
    for (int i = 0; i < n; ++i) residual[i] = 0.0;                                      // :79  y - X*beta
    for (int k = 0; k < K; ++k) { const double b = beta[k]; const double* xk = X + (size_t)k * n; for (int i = 0; i < n; ++i) residual[i] += xk[i] * b; }
    for (int i = 0; i < n; ++i) residual[i] = y[i] - residual[i];
    double iter_loss = compute_loss_vec(n, residual.data(), K, beta, lambda, alpha);    // :80
    double pre_loss;
    while (true) {                                                                      // :82
        inc.clear(); ex.clear();
        for (int k = 0; k < K; ++k) (active[k] ? inc : ex).push_back(k);                // :83-84
        do {
            pre_loss = iter_loss;                                                       // :87
            randperm(ps, (int)inc.size(), ord.data(), inc.data(), K);                   // :89
            for (size_t i = 0; i < inc.size(); ++i) {
                const int k = inc[ord[i]];                                              // :92
                const double* xk = X + (size_t)k * n;
                double dot = 0.0;
                for (int t = 0; t < n; ++t) dot += residual[t] * xk[t];
                const double upper = dot + beta[k] * XtX[k + (size_t)k * K];            // :94
                double update;
                if (std::fabs(upper) > lambda * alpha)                                  // :99-104
                    update = sgn(upper) * std::max(std::fabs(upper) - lambda * alpha, 0.0) / (XtX[k + (size_t)k * K] + lambda * (1 - alpha));
                else
                    update = 0.0;
                if (update != beta[k]) {                                                // :106-109
                    const double d = update - beta[k];
                    for (int t = 0; t < n; ++t) residual[t] -= d * xk[t];
                    beta[k] = update;
                }
            }
            iter_loss = compute_loss_vec(n, residual.data(), K, beta, lambda, alpha);   // :112
            if (st) st->sweeps++;
        } while (std::fabs(pre_loss - iter_loss) > tol);                                // :114
        if (st) st->rounds++;
        bool any = false;                                                               // :118-124
        for (int e : ex) {
            double g = 0.0;
            for (int k : inc) g += XtX[e + (size_t)k * K] * beta[k];
            g -= Xty[e];
            if (std::fabs(g) > alpha * lambda) { active[e] = 1; any = true; }
        }
        if (!any) break;
    }
}

struct Problem {
    int N, P, C, Q, K, inc_continuous, tuning;
    const double* data; const int32_t* levels; const double* ctns; const int32_t* train; const int32_t* test;
    std::vector<int> L;          // levels per confounder
};

// src/optimize.cpp:139-198  optimize_row()
int optimize_row(const Problem& pb, const std::vector<double>& residual, double* A /*L x K*/, int L, const double* V,
                 const int32_t* z, const std::vector<double>& gram, double lambda, int n_cores) {
    const int N = pb.N, P = pb.P, K = pb.K;
    int err = 0;
    if (pb.tuning == 1) {
#pragma omp parallel for num_threads(n_cores) schedule(dynamic, 1)
        for (int s = 1; s <= L; ++s) {                                                  // :153 (seq = 1..L)
            std::vector<double> XtX((size_t)K * K, 0.0), Xty(K, 0.0), Zc((size_t)K * K);
            int n_rows = 0;
            for (int k = 0; k < N; ++k) {
                if (z[k] != s) continue;                                                // :159
                ++n_rows;
                std::fill(Zc.begin(), Zc.end(), 0.0);
                for (int j = 0; j < P; ++j) {
                    const double* vj = V + (size_t)j * K;
                    if (pb.train[k + (size_t)j * N] != 0) {                             // :162,171
                        const double o = residual[k + (size_t)j * N];
                        for (int a = 0; a < K; ++a) Xty[a] += vj[a] * o;
                    } else {                                                            // :163,170 complement
                        for (int b = 0; b < K; ++b) { const double vb = vj[b]; for (int a = 0; a < K; ++a) Zc[a + (size_t)b * K] += vj[a] * vb; }
                    }
                }
                for (size_t t = 0; t < XtX.size(); ++t) XtX[t] += gram[t] - Zc[t];      // :170
            }
            if (n_rows == 0) continue;            // level id absent: reference's unique() would skip it
            for (int a = 0; a < K; ++a) XtX[a + (size_t)a * K] += lambda;               // :174
            if (chol_solve(K, XtX.data(), Xty.data())) {
#pragma omp atomic write
                err = 1;
                continue;
            }
            for (int a = 0; a < K; ++a) A[(s - 1) + (size_t)a * L] = Xty[a];            // :175
        }
    } else {
        // :180  Xtys = c_factor * trans(residual)  (K x N)
        std::vector<double> Xtys((size_t)K * N, 0.0);
#pragma omp parallel for num_threads(n_cores) schedule(static)
        for (int k = 0; k < N; ++k) {
            double* o = Xtys.data() + (size_t)k * K;
            for (int j = 0; j < P; ++j) {
                const double r = residual[k + (size_t)j * N];
                const double* vj = V + (size_t)j * K;
                for (int a = 0; a < K; ++a) o[a] += vj[a] * r;
            }
        }
#pragma omp parallel for num_threads(n_cores) schedule(dynamic, 1)
        for (int s = 1; s <= L; ++s) {                                                  // :183
            std::vector<double> XtX((size_t)K * K), Xty(K, 0.0);
            int n_rows = 0;
            for (int k = 0; k < N; ++k) if (z[k] == s) { ++n_rows; for (int a = 0; a < K; ++a) Xty[a] += Xtys[a + (size_t)k * K]; }   // :188
            if (n_rows == 0) continue;
            for (size_t t = 0; t < XtX.size(); ++t) XtX[t] = n_rows * gram[t];          // :186
            for (int a = 0; a < K; ++a) XtX[a + (size_t)a * K] += lambda;               // :187
            if (chol_solve(K, XtX.data(), Xty.data())) {
#pragma omp atomic write
                err = 1;
                continue;
            }
            for (int a = 0; a < K; ++a) A[(s - 1) + (size_t)a * L] = Xty[a];            // :190
        }
    }
    return err;
}

// src/optimize.cpp:77-137  optimize_continuous_v2()
// `data` is the residual with this covariate's contribution added back (optimize.cpp:344-345).
int optimize_continuous_v2(const Problem& pb, const std::vector<double>& data, double* w /*K, stride ldw*/, int ldw,
                           const double* V, const double* x, const std::vector<double>& gram, double lambda) {
    const int N = pb.N, P = pb.P, K = pb.K;
    if (pb.tuning == 1) {
        std::vector<double> resid((size_t)N * P);
        // :84  resid = data - x * w * V
        std::vector<double> wv(P);
#pragma omp parallel for schedule(static)
        for (int j = 0; j < P; ++j) {
            double s = 0.0; for (int a = 0; a < K; ++a) s += w[(size_t)a * ldw] * V[a + (size_t)j * K];
            for (int k = 0; k < N; ++k) resid[k + (size_t)j * N] = data[k + (size_t)j * N] - x[k] * s;
        }
        std::vector<double> sq_x(N), norm_factor(K, 0.0), pre(K);
        for (int k = 0; k < N; ++k) sq_x[k] = x[k] * x[k];                              // :90
        for (int j = 0; j < P; ++j) for (int a = 0; a < K; ++a) norm_factor[a] += V[a + (size_t)j * K] * V[a + (size_t)j * K];
        while (true) {                                                                  // :102
            for (int a = 0; a < K; ++a) pre[a] = w[(size_t)a * ldw];
            for (int i = 0; i < K; ++i) {                                               // :104 cyclic order
                const double wi = w[(size_t)i * ldw];
                double Xty = 0.0, XtX = 0.0;
                // :107 resid += w_i x V[i,:] ; :111 Xty = x' (M o resid) V[i,:]'
                std::vector<double> row_xty(N, 0.0), row_zero(N, 0.0);
                int nthr = 1;
#ifdef _OPENMP
                nthr = omp_get_max_threads();
#endif
                std::vector<std::vector<double>> lx(nthr, std::vector<double>(N, 0.0)), lz(nthr, std::vector<double>(N, 0.0));
#pragma omp parallel num_threads(nthr)
                {
                    int tid = 0;
#ifdef _OPENMP
                    tid = omp_get_thread_num();
#endif
                    double* mx = lx[tid].data(); double* mz = lz[tid].data();
#pragma omp for schedule(static)
                    for (int j = 0; j < P; ++j) {
                        const double vij = V[i + (size_t)j * K];
                        for (int k = 0; k < N; ++k) {
                            double& r = resid[k + (size_t)j * N];
                            r += wi * x[k] * vij;
                            if (pb.train[k + (size_t)j * N] != 0) mx[k] += r * vij; else mz[k] += vij * vij;
                        }
                    }
                }
                for (int t = 0; t < nthr; ++t) for (int k = 0; k < N; ++k) { row_xty[k] += lx[t][k]; row_zero[k] += lz[t][k]; }   // fixed order
                for (int k = 0; k < N; ++k) Xty += x[k] * row_xty[k];
                for (int k = 0; k < N; ++k) XtX += sq_x[k] * (norm_factor[i] - row_zero[k]);   // :112-115
                const double nw = Xty / (XtX + lambda);                                 // :117
                w[(size_t)i * ldw] = nw;
#pragma omp parallel for schedule(static)
                for (int j = 0; j < P; ++j) {                                           // :118
                    const double vij = V[i + (size_t)j * K];
                    for (int k = 0; k < N; ++k) resid[k + (size_t)j * N] -= nw * x[k] * vij;
                }
            }
            double diff = 0.0;
            for (int a = 0; a < K; ++a) diff += std::fabs(pre[a] - w[(size_t)a * ldw]);
            if (diff < 1e-1) break;                                                     // :122
        }
        return 0;
    }
    // :127-131  (x'x * gram + lambda I) w = V * data' * x
    std::vector<double> Xty(K, 0.0), XtX((size_t)K * K);
    double xx = 0.0; for (int k = 0; k < N; ++k) xx += x[k] * x[k];
    std::vector<double> dx(P, 0.0);
#pragma omp parallel for schedule(static)
    for (int j = 0; j < P; ++j) { double s = 0.0; for (int k = 0; k < N; ++k) s += data[k + (size_t)j * N] * x[k]; dx[j] = s; }
    for (int j = 0; j < P; ++j) for (int a = 0; a < K; ++a) Xty[a] += V[a + (size_t)j * K] * dx[j];
    for (size_t t = 0; t < XtX.size(); ++t) XtX[t] = xx * gram[t];
    for (int a = 0; a < K; ++a) XtX[a + (size_t)a * K] += lambda;
    if (chol_solve(K, XtX.data(), Xty.data())) return 1;
    for (int a = 0; a < K; ++a) w[(size_t)a * ldw] = Xty[a];
    return 0;
}

// src/optimize.cpp:200-253  optimize_col()  (called with `data`, not the residual: optimize.cpp:376)
int optimize_col(const Problem& pb, const double* U /*N x K col-major*/, double* V, double lambda, double alpha, double tol,
                 int n_cores, int perm_mode, uint64_t seed, uint32_t als_iter, RRng* rstream, long long* sweeps_total) {
    const int N = pb.N, P = pb.P, K = pb.K;
    std::vector<double> gram((size_t)K * K, 0.0);
    for (int b = 0; b < K; ++b) for (int a = 0; a < K; ++a) {                           // :205 / :234
        double s = 0.0; for (int i = 0; i < N; ++i) s += U[i + (size_t)a * N] * U[i + (size_t)b * N];
        gram[a + (size_t)b * K] = s;
    }
    int err = 0;
    long long sweeps = 0;
    const int threads = (perm_mode == 0) ? 1 : n_cores;   // mode A consumes one global stream: single-thread semantics
    if (pb.tuning == 1) {
#pragma omp parallel for num_threads(threads) schedule(dynamic, 100) reduction(+ : sweeps)
        for (int j = 0; j < P; ++j) {                                                   // :215
            std::vector<int> sel; sel.reserve(N);
            std::vector<double> XtX(gram), Xty(K, 0.0);
            for (int i = 0; i < N; ++i) {
                if (pb.train[i + (size_t)j * N] != 0) sel.push_back(i);                 // :216
                else for (int b = 0; b < K; ++b) { const double ub = U[i + (size_t)b * N]; for (int a = 0; a < K; ++a) XtX[a + (size_t)b * K] -= U[i + (size_t)a * N] * ub; }   // :218-219
            }
            const int n = (int)sel.size();
            std::vector<double> feature((size_t)n * K), outcome(n);
            for (int t = 0; t < n; ++t) outcome[t] = pb.data[sel[t] + (size_t)j * N];    // :220-221
            for (int a = 0; a < K; ++a) for (int t = 0; t < n; ++t) feature[t + (size_t)a * n] = U[sel[t] + (size_t)a * N];   // :217
            for (int a = 0; a < K; ++a) { double s = 0.0; for (int t = 0; t < n; ++t) s += feature[t + (size_t)a * n] * outcome[t]; Xty[a] = s; }   // :222
            double* vj = V + (size_t)j * K;
            if (alpha == 0.0) {                                                         // :224-226
                for (int a = 0; a < K; ++a) XtX[a + (size_t)a * K] += lambda;
                if (chol_solve(K, XtX.data(), Xty.data())) {
#pragma omp atomic write
                    err = 1;
                } else for (int a = 0; a < K; ++a) vj[a] = Xty[a];
            } else {                                                                    // :228
                PermSrc ps; ps.mode = perm_mode; ps.seed = seed; ps.als_iter = als_iter; ps.gene = (uint64_t)j; ps.r = rstream;
                std::vector<double> beta(K); CdStats st;
                strong_cd(n, K, feature.data(), outcome.data(), vj, lambda, alpha, XtX.data(), Xty.data(), tol, ps, beta.data(), &st);
                for (int a = 0; a < K; ++a) vj[a] = beta[a];
                sweeps += st.sweeps;
                if (g_sweep_sink && (long long)als_iter * P + j < g_sweep_sink_len) g_sweep_sink[(size_t)als_iter * P + j] = st.sweeps;
            }
        }
    } else {
        if (alpha == 0.0) {                                                             // :237-240
            std::vector<double> XtX(gram);
            for (int a = 0; a < K; ++a) XtX[a + (size_t)a * K] += lambda;
            std::vector<double> Xty((size_t)K * P);
#pragma omp parallel for num_threads(n_cores) schedule(static)
            for (int j = 0; j < P; ++j) for (int a = 0; a < K; ++a) { double s = 0.0; for (int i = 0; i < N; ++i) s += U[i + (size_t)a * N] * pb.data[i + (size_t)j * N]; Xty[a + (size_t)j * K] = s; }
            if (chol_solve(K, XtX.data(), Xty.data(), P, K)) err = 1;
            else std::memcpy(V, Xty.data(), sizeof(double) * (size_t)K * P);
        } else {
#pragma omp parallel for num_threads(threads) schedule(dynamic, 100) reduction(+ : sweeps)
            for (int j = 0; j < P; ++j) {                                               // :245-247
                std::vector<double> Xty(K), beta(K);
                const double* yj = pb.data + (size_t)j * N;
                for (int a = 0; a < K; ++a) { double s = 0.0; for (int i = 0; i < N; ++i) s += U[i + (size_t)a * N] * yj[i]; Xty[a] = s; }   // :235
                PermSrc ps; ps.mode = perm_mode; ps.seed = seed; ps.als_iter = als_iter; ps.gene = (uint64_t)j; ps.r = rstream;
                CdStats st;
                double* vj = V + (size_t)j * K;
                strong_cd(N, K, U, yj, vj, lambda, alpha, gram.data(), Xty.data(), tol, ps, beta.data(), &st);
                for (int a = 0; a < K; ++a) vj[a] = beta[a];
                sweeps += st.sweeps;
                if (g_sweep_sink && (long long)als_iter * P + j < g_sweep_sink_len) g_sweep_sink[(size_t)als_iter * P + j] = st.sweeps;
            }
        }
    }
    if (sweeps_total) *sweeps_total += sweeps;
    return err;
}

// residual (+/-)= (Z_c A_c) V   — src/optimize.cpp:338,354 ; rows gathered instead of a dense one-hot GEMM
void add_block(const Problem& pb, std::vector<double>& residual, const double* A, int L, const int32_t* z, const double* V, double sign) {
    const int N = pb.N, P = pb.P, K = pb.K;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < P; ++j) {
        const double* vj = V + (size_t)j * K;
        double* rj = residual.data() + (size_t)j * N;
        for (int k = 0; k < N; ++k) {
            double s = 0.0; const int r = z[k] - 1;
            for (int a = 0; a < K; ++a) s += A[r + (size_t)a * L] * vj[a];
            rj[k] += sign * s;
        }
    }
}
// residual (+/-)= x_q w_q V   — src/optimize.cpp:344,348
void add_continuous(const Problem& pb, std::vector<double>& residual, const double* w, int ldw, const double* x, const double* V, double sign) {
    const int N = pb.N, P = pb.P, K = pb.K;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < P; ++j) {
        double s = 0.0; for (int a = 0; a < K; ++a) s += w[(size_t)a * ldw] * V[a + (size_t)j * K];
        double* rj = residual.data() + (size_t)j * N;
        for (int k = 0; k < N; ++k) rj[k] += sign * x[k] * s;
    }
}

// src/optimize.cpp:286-289 and :365-373  row_factor = sum_c A_c[z_c - 1, :] + X W
void build_row_factor(const Problem& pb, double* const* F, double* U) {
    const int N = pb.N, K = pb.K;
    std::fill(U, U + (size_t)N * K, 0.0);
    for (int c = 0; c < pb.C; ++c) {
        const int L = pb.L[c]; const int32_t* z = pb.levels + (size_t)c * N;
        for (int a = 0; a < K; ++a) for (int k = 0; k < N; ++k) U[k + (size_t)a * N] += F[c][(z[k] - 1) + (size_t)a * L];
    }
    if (pb.inc_continuous == 1) {
        const double* W = F[pb.C];
        for (int a = 0; a < K; ++a) for (int k = 0; k < N; ++k) {
            double s = 0.0; for (int q = 0; q < pb.Q; ++q) s += pb.ctns[k + (size_t)q * N] * W[q + (size_t)a * pb.Q];
            U[k + (size_t)a * N] += s;
        }
    }
}

// src/utils.cpp:52-54 predict() + src/optimize.cpp:321,378 residual = data - predictions
void predict_residual(const Problem& pb, const double* U, const double* V, std::vector<double>& residual) {
    const int N = pb.N, P = pb.P, K = pb.K;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < P; ++j) {
        const double* vj = V + (size_t)j * K;
        for (int k = 0; k < N; ++k) {
            double s = 0.0; for (int a = 0; a < K; ++a) s += U[k + (size_t)a * N] * vj[a];
            residual[k + (size_t)j * N] = pb.data[k + (size_t)j * N] - s;
        }
    }
}

// src/utils.cpp:56-77 evaluate()
int evaluate(const Problem& pb, const std::vector<double>& residual, double& sum_residual, double& train_rmse, double& test_rmse,
             long long n_train, long long n_test) {
    const int N = pb.N, P = pb.P;
    if (pb.tuning == 0) {
        std::vector<double> rows(N, 0.0);                                               // :62 accu(sum(square(residual), 1))
        for (int j = 0; j < P; ++j) for (int k = 0; k < N; ++k) rows[k] += residual[k + (size_t)j * N] * residual[k + (size_t)j * N];
        double s = 0.0; for (int k = 0; k < N; ++k) s += rows[k];
        sum_residual = s;
        train_rmse = std::sqrt(sum_residual / ((double)P * (double)N));                  // :63
    } else {
        double st = 0.0, ss = 0.0;
        for (int j = 0; j < P; ++j) for (int k = 0; k < N; ++k) {
            const double r = residual[k + (size_t)j * N];
            if (pb.train[k + (size_t)j * N] != 0) st += r * r;                          // :65
            if (pb.test[k + (size_t)j * N] != 0) ss += r * r;                           // :67
        }
        sum_residual = st;
        train_rmse = std::sqrt(sum_residual / (double)n_train);                         // :66
        if (n_test == 0) return 2;                                                      // arma::mean of an empty vector throws
        test_rmse = std::sqrt(ss / (double)n_test);
    }
    return 0;
}

// src/utils.cpp:79-102 compute_loss(field, ...)
double compute_loss_global(const Problem& pb, double* const* F, const int* frows, int n_factors, const double* V, double lambda1,
                           double lambda2, double alpha, double sum_residual, double* parts) {
    double row_reg = 0.0;
    for (int c = 0; c < n_factors; ++c) {
        double s = 0.0; const size_t n = (size_t)frows[c] * pb.K;
        for (size_t t = 0; t < n; ++t) s += F[c][t] * F[c][t];
        row_reg += lambda1 * s;                                                         // :85  lambda1 * ||A_c||_F^2
    }
    double v2 = 0.0, v1 = 0.0;
    const size_t nv = (size_t)pb.K * pb.P;
    for (size_t t = 0; t < nv; ++t) { v2 += V[t] * V[t]; v1 += std::fabs(V[t]); }
    const double col_reg = lambda2 * (1 - alpha) * v2;                                  // :88
    const double l1_reg = lambda2 * alpha * v1;                                         // :91
    if (parts) { parts[0] = sum_residual / 2; parts[1] = row_reg / 2; parts[2] = col_reg / 2; parts[3] = l1_reg; }
    return sum_residual / 2 + row_reg / 2 + col_reg / 2 + l1_reg;                       // :93
}

}  // namespace

extern "C" {

// One record per evaluation (initial + every 10th iteration).
struct oracle_check {
    int32_t iter;          // -1 for the initial evaluation (optimize.cpp:320-323)
    int32_t pad;
    double sum_residual, train_rmse, test_rmse, row_reg, col_reg, l1_reg, loss, delta_loss, decay;
};

int oracle_version(void) { return 1; }
void oracle_set_sweep_sink(int* buf, long long len) { g_sweep_sink = buf; g_sweep_sink_len = len; }

// src/coordinate_descent.cpp:57-127 — single-column entry (KATs).
int oracle_strong_cd(int n, int K, const double* X, const double* y, const double* wstart, double lambda, double alpha,
                     const double* XtX, const double* Xty, double tol, int perm_mode, uint64_t seed, uint32_t als_iter,
                     uint64_t gene, uint32_t r_seed, double* beta_out, int* sweeps_out, int* rounds_out) {
    RRng r; r.set_seed(r_seed);
    PermSrc ps; ps.mode = perm_mode; ps.seed = seed; ps.als_iter = als_iter; ps.gene = gene; ps.r = &r;
    CdStats st;
    strong_cd(n, K, X, y, wstart, lambda, alpha, XtX, Xty, tol, ps, beta_out, &st);
    if (sweeps_out) *sweeps_out = st.sweeps;
    if (rounds_out) *rounds_out = st.rounds;
    return 0;
}

// The counter-based permutation of mode B, exposed so tests can pin the GPU permutation bit-for-bit.
void oracle_randperm_b(uint64_t seed, uint32_t als_iter, uint64_t gene, uint32_t draw, int n, int* ord) {
    PermSrc ps; ps.mode = 1; ps.seed = seed; ps.als_iter = als_iter; ps.gene = gene; ps.draw = draw;
    randperm(ps, n, ord);
}
// the same restricted to an active set: inc = ascending active coordinates out of K; ord = visiting order as indices into inc
void oracle_randperm_b_inc(uint64_t seed, uint32_t als_iter, uint32_t draw, int K, const int* inc, int n_inc, int* ord) {
    PermSrc ps; ps.mode = 1; ps.seed = seed; ps.als_iter = als_iter; ps.draw = draw;
    randperm(ps, n_inc, ord, inc, K);
}

// src/optimize.cpp:256-422  optimize()
//   factors: n_factors = C (+1 when inc_continuous) column-major matrices, factor c is L_c x K (continuous: Q x K);
//   they and column_factor (K x P) are updated IN PLACE like the reference (optimize.cpp:283-284).
//   levels: N x C column-major, 1-based, every column's values exactly 1..L_c.
//   train/test: N x P column-major int32 0/1 (R integer matrices), may be NULL when tuning == 0.
// Returns 0 ok, 1 not SPD, 2 empty test set (tuning=1), 3 invalid argument.
int oracle_optimize(int N, int P, const double* data, int n_factors, double* const* factors, const int* factor_rows,
                    double* column_factor, int C, const int32_t* levels, int Q, const double* ctns, const int32_t* train,
                    const int32_t* test, int inc_continuous, int latent_dim, double lambda1, double lambda2, double alpha,
                    int tuning, double global_tol, double sub_tol, unsigned max_iter, int perm_mode, uint64_t seed,
                    uint32_t r_seed, int n_cores_row, int n_cores_col, double* train_rmse_out, double* test_rmse_out,
                    double* loss_out, int* iters_run_out, oracle_check* checks, int max_checks, int* n_checks_out,
                    long long* cd_sweeps_out, double* seconds_in_loop_out) {
    if (inc_continuous != 0 && inc_continuous != 1) return 3;                           // :270-273
    if (tuning != 0 && tuning != 1) return 3;                                           // :193-195 etc.
    if (n_factors != C + (inc_continuous ? 1 : 0)) return 3;
    if (tuning == 1 && (!train || !test)) return 3;
    Problem pb; pb.N = N; pb.P = P; pb.C = C; pb.Q = Q; pb.K = latent_dim; pb.inc_continuous = inc_continuous; pb.tuning = tuning;
    pb.data = data; pb.levels = levels; pb.ctns = ctns; pb.train = train; pb.test = test;
    const int K = latent_dim;
    for (int c = 0; c < C; ++c) {
        int L = 0; for (int k = 0; k < N; ++k) { const int v = levels[k + (size_t)c * N]; if (v < 1) return 3; L = std::max(L, v); }
        if (L != factor_rows[c]) return 3;       // reference indexes row (level-1): levels must be 1..L_c
        std::vector<char> seen(L + 1, 0); for (int k = 0; k < N; ++k) seen[levels[k + (size_t)c * N]] = 1;
        for (int v = 1; v <= L; ++v) if (!seen[v]) return 3;
        pb.L.push_back(L);
    }
    if (inc_continuous && factor_rows[C] != Q) return 3;
#ifdef _OPENMP
    const int maxt = omp_get_max_threads();
#else
    const int maxt = 1;
#endif
    const int row_cores = std::max(1, std::min(n_cores_row > 0 ? n_cores_row : 10, maxt));   // optimize.cpp:140 default 10
    const int col_cores = std::max(1, std::min(n_cores_col > 0 ? n_cores_col : 30, maxt));   // optimize.cpp:376 literal 30
    RRng rstream; rstream.set_seed(r_seed);

    long long n_train = 0, n_test = 0;                                                  // :316-317
    if (tuning == 1) for (size_t t = 0; t < (size_t)N * P; ++t) { n_train += train[t] != 0; n_test += test[t] != 0; }

    std::vector<double> U((size_t)N * K), residual((size_t)N * P), gram((size_t)K * K);
    double* V = column_factor;
    build_row_factor(pb, factors, U.data());                                            // :286-289
    predict_residual(pb, U.data(), V, residual);                                        // :320-321
    double sum_residual = 0, train_rmse = std::numeric_limits<double>::quiet_NaN(), test_rmse = std::numeric_limits<double>::quiet_NaN();
    int rc = evaluate(pb, residual, sum_residual, train_rmse, test_rmse, n_train, n_test);   // :322
    if (rc) return rc;
    double parts[4];
    double loss = compute_loss_global(pb, factors, factor_rows, n_factors, V, lambda1, lambda2, alpha, sum_residual, parts);   // :323
    int n_checks = 0;
    auto record = [&](int it, double delta, double decay) {
        if (checks && n_checks < max_checks) checks[n_checks] = {it, 0, sum_residual, train_rmse, test_rmse, parts[1], parts[2], parts[3], loss, delta, decay};
        ++n_checks;
    };
    record(-1, 0.0, 1.0);
    double pre_loss, decay = 1.0;
    unsigned iter = 0;
    long long sweeps = 0;
    int err = 0;
#ifdef _OPENMP
    const double t0 = omp_get_wtime();
#endif
    while (iter <= max_iter) {                                                          // :325
        for (int b = 0; b < K; ++b) for (int a = 0; a < K; ++a) {                       // :332 gram = V V'
            double s = 0.0; for (int j = 0; j < P; ++j) s += V[a + (size_t)j * K] * V[b + (size_t)j * K];
            gram[a + (size_t)b * K] = s;
        }
        for (int c = 0; c < n_factors; ++c) {                                           // :335
            if (c < C) {
                const int32_t* z = levels + (size_t)c * N;
                add_block(pb, residual, factors[c], pb.L[c], z, V, +1.0);               // :338
                err |= optimize_row(pb, residual, factors[c], pb.L[c], V, z, gram, lambda1, row_cores);   // :339
                if (c != n_factors - 1) add_block(pb, residual, factors[c], pb.L[c], z, V, -1.0);   // :353-355
            } else {
                double* W = factors[c];
                for (int q = 0; q < Q; ++q) {                                           // :342-350
                    const double* x = ctns + (size_t)q * N;
                    add_continuous(pb, residual, W + q, Q, x, V, +1.0);                 // :344
                    err |= optimize_continuous_v2(pb, residual, W + q, Q, V, x, gram, lambda1);   // :345
                    if (q != Q - 1) add_continuous(pb, residual, W + q, Q, x, V, -1.0); // :347-349
                }
                // :353 — the continuous block is always last (c == n_factors-1): nothing subtracted.
            }
        }
        if (err) return 1;
        build_row_factor(pb, factors, U.data());                                        // :365-373
        err |= optimize_col(pb, U.data(), V, lambda2, alpha, sub_tol * decay, col_cores, perm_mode, seed, iter, &rstream, &sweeps);   // :376
        if (err) return 1;
        predict_residual(pb, U.data(), V, residual);                                    // :377-378
        if (iter % 10 == 0) {                                                           // :381
            pre_loss = loss;
            rc = evaluate(pb, residual, sum_residual, train_rmse, test_rmse, n_train, n_test);
            if (rc) return rc;
            loss = compute_loss_global(pb, factors, factor_rows, n_factors, V, lambda1, lambda2, alpha, sum_residual, parts);
            const double delta_loss = pre_loss - loss;                                  // :386
            if (delta_loss / 1000 <= 1e-6) decay = 1e-6;                                // :389-403
            else if (delta_loss / 1000 <= 1e-5) decay = 1e-5;
            else if (delta_loss / 1000 <= 1e-4) decay = 1e-4;
            else if (delta_loss / 1000 <= 1e-3) decay = 1e-3;
            else if (delta_loss / 1000 <= 1e-2) decay = 1e-2;
            else if (delta_loss / 1000 <= 1e-1) decay = 1e-1;
            else decay = 1.0;
            record((int)iter, delta_loss, decay);
            if ((pre_loss - loss) / pre_loss < global_tol) { ++iter; --iter; break; }    // :405-407 (iter not incremented on break)
        }
        iter++;                                                                         // :409
    }
#ifdef _OPENMP
    if (seconds_in_loop_out) *seconds_in_loop_out = omp_get_wtime() - t0;
#else
    if (seconds_in_loop_out) *seconds_in_loop_out = 0.0;
#endif
    if (train_rmse_out) *train_rmse_out = train_rmse;
    if (test_rmse_out) *test_rmse_out = test_rmse;     // NaN when tuning == 0 (reference returns it uninitialised: utils.cpp:61-63)
    if (loss_out) *loss_out = loss;
    if (iters_run_out) *iters_run_out = (int)iter;     // value of `iter` when the loop ended
    if (n_checks_out) *n_checks_out = n_checks;
    if (cd_sweeps_out) *cd_sweeps_out = sweeps;
    return 0;
}

// src/fit_interaction.cpp:10-90 — stand-alone, un-regularised per-level normal equations on a residual.
// (Dead code in the reference: the header/definition disagree and `.row(i) = <colvec>` would throw. The
// oracle restates the intended math: interactions[s-1,:] = solve(sum XtX, sum Xty).)
int oracle_fit_interaction(int N, int P, int K, const double* residual, const int32_t* train, double* interactions, int L,
                           const int32_t* z, const double* V, int tuning) {
    if (tuning != 0 && tuning != 1) return 3;
    std::vector<double> gram((size_t)K * K, 0.0);
    for (int b = 0; b < K; ++b) for (int a = 0; a < K; ++a) { double s = 0.0; for (int j = 0; j < P; ++j) s += V[a + (size_t)j * K] * V[b + (size_t)j * K]; gram[a + (size_t)b * K] = s; }
    int err = 0;
    for (int s = 1; s <= L; ++s) {
        std::vector<double> XtX((size_t)K * K, 0.0), Xty(K, 0.0);
        int n_rows = 0;
        for (int k = 0; k < N; ++k) {
            if (z[k] != s) continue;
            ++n_rows;
            if (tuning == 1) {                                                          // :37-52
                for (int j = 0; j < P; ++j) if (train[k + (size_t)j * N] != 0) {
                    const double* vj = V + (size_t)j * K; const double o = residual[k + (size_t)j * N];
                    for (int b = 0; b < K; ++b) { for (int a = 0; a < K; ++a) XtX[a + (size_t)b * K] += vj[a] * vj[b]; }
                    for (int a = 0; a < K; ++a) Xty[a] += vj[a] * o;
                }
            } else {                                                                    // :59-81
                for (size_t t = 0; t < XtX.size(); ++t) XtX[t] += gram[t];
                for (int j = 0; j < P; ++j) { const double* vj = V + (size_t)j * K; const double o = residual[k + (size_t)j * N]; for (int a = 0; a < K; ++a) Xty[a] += vj[a] * o; }
            }
        }
        if (n_rows == 0) continue;
        if (chol_solve(K, XtX.data(), Xty.data())) { err = 1; continue; }               // :54 / :82
        for (int a = 0; a < K; ++a) interactions[(s - 1) + (size_t)a * L] = Xty[a];
    }
    return err;
}

// src/optimize.cpp:77-137 — stand-alone entry with the reference's 8 arguments (`_insider_optimize_continuous_v2`,
// src/RcppExports.cpp:69-84). data/indicator N x P column-major, w (K) in/out, V K x P, x (N), gram K x K (tuning = 0 only).
int oracle_optimize_continuous_v2(int N, int P, int K, const double* data, const int32_t* indicator, double* w, const double* V,
                                  const double* x, const double* gram, double lambda, int tuning) {
    if (tuning != 0 && tuning != 1) return 3;
    if (tuning == 1 && !indicator) return 3;
    Problem pb; pb.N = N; pb.P = P; pb.C = 0; pb.Q = 1; pb.K = K; pb.inc_continuous = 1; pb.tuning = tuning;
    pb.data = data; pb.train = indicator; pb.test = indicator;
    std::vector<double> d(data, data + (size_t)N * P), g((size_t)K * K, 0.0);
    if (gram) g.assign(gram, gram + (size_t)K * K);
    else for (int b = 0; b < K; ++b) for (int a = 0; a < K; ++a) { double s = 0.0; for (int j = 0; j < P; ++j) s += V[a + (size_t)j * K] * V[b + (size_t)j * K]; g[a + (size_t)b * K] = s; }
    return optimize_continuous_v2(pb, d, w, 1, V, x, g, lambda);
}

// OpenMP team control for the timed CPU-baseline legs of bench.py: launchers such as torchrun export OMP_NUM_THREADS=1, and the
// team sizes of oracle_optimize are clamped to omp_get_max_threads(). Returns the team size now in effect.
int oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n; return 1;
#endif
}

// Helpers exposed for unit tests.
int oracle_chol_solve(int K, double* A, double* b, int nrhs) { return chol_solve(K, A, b, nrhs, K); }
void oracle_r_unif(uint32_t seed, int n, double* out) { RRng r; r.set_seed(seed); for (int i = 0; i < n; ++i) out[i] = r.unif_rand(); }
void oracle_randperm_r(uint32_t r_seed, int n, int* ord) { RRng r; r.set_seed(r_seed); PermSrc ps; ps.mode = 0; ps.r = &r; randperm(ps, n, ord); }

}  // extern "C"
