// Row-side (confounder factor) updates in sufficient-statistic form.
//
// The reference adds one confounder's contribution back to a dense N x P residual, solves per level, and subtracts
// it again (src/optimize.cpp:335-362), i.e. 2C-1 read-modify-write passes over N x P per iteration. With
//     B_k  = sum_j m_kj y_kj v_j          (k_row_b)
//     Gk_k = sum_j m_kj v_j v_j^T = G - D_k   (k_gram_v, k_row_comp_gram)
// the normal equations of level s of confounder c (src/optimize.cpp:150-176 / :178-191) are
//     XtX_s = sum_{k in s} Gk_k + lambda I ,   Xty_s = sum_{k in s} [ B_k - Gk_k (u_k - a_{c,s}) ]
// where u_k is the current row factor (Gauss-Seidel: blocks updated earlier in the same iteration are already in u_k).
// In the dense path (tuning = 0) Gk_k = G for every row.
#include "common.cuh"
#include "kernels.cuh"

namespace ib {

namespace {

// shared with k_cd.cu in spirit: warp Cholesky on an odd-pitch shared matrix
__device__ bool chol_factor(double* S, int ld, int K, int lane) {
    bool ok = true;
    for (int j = 0; j < K; ++j) {
        const double ajj = S[j * ld + j];
        if (!(ajj > 0.0)) ok = false;
        const double dj = sqrt(ajj);
        __syncwarp();
        double lij = 0.0;
        if (lane == j) S[j * ld + j] = dj;
        if (lane > j && lane < K) { lij = S[lane * ld + j] / dj; S[lane * ld + j] = lij; }
        __syncwarp();
        for (int m = j + 1; m < K; ++m)
            if (lane >= m && lane < K) S[lane * ld + m] = fma(-lij, S[m * ld + j], S[lane * ld + m]);
        __syncwarp();
    }
    return ok;
}
__device__ double chol_subst(const double* S, int ld, int K, int lane, double b) {
    for (int i = 0; i < K; ++i) {
        double xi = b / S[i * ld + i];
        xi = __shfl_sync(FULL, xi, i);
        if (lane == i) b = xi;
        else if (lane > i && lane < K) b = fma(-S[lane * ld + i], xi, b);
    }
    for (int i = K - 1; i >= 0; --i) {
        double xi = b / S[i * ld + i];
        xi = __shfl_sync(FULL, xi, i);
        if (lane == i) b = xi;
        else if (lane < i) b = fma(-S[i * ld + lane], xi, b);
    }
    return b;
}

// GL[s][e] = sum_{k in s} (G[e] - D[k][e])
__global__ void __launch_bounds__(256) k_level_gram(const int* __restrict__ rows_sorted, const int* __restrict__ level_start,
                                                    const double* __restrict__ G, const double* __restrict__ D, double* __restrict__ GL, int KK) {
    const int s = blockIdx.x;
    const int b = level_start[s], e = level_start[s + 1];
    for (int x = threadIdx.x; x < KK; x += 256) {
        const double gx = G[x];
        double acc = 0.0;
        for (int r = b; r < e; ++r) acc += gx - D[(size_t)rows_sorted[r] * KK + x];
        GL[(size_t)s * KK + x] = acc;
    }
}

// warp per row: T_k = B_k - M_k (u_k - a), M_k = G - D_k (masked) or G. `sub_own` = 0 drops the "- a" term (continuous block).
__global__ void __launch_bounds__(256) k_row_rhs(int N, int KP, const int* __restrict__ level_of_row, const double* __restrict__ A,
                                                 const double* __restrict__ B, const double* __restrict__ G, const double* __restrict__ D,
                                                 const double* __restrict__ U, double* __restrict__ T) {
    const int lane = threadIdx.x & 31;
    const int k = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (k >= N) return;
    double w = 0.0;
    if (lane < KP) {
        w = U[(size_t)k * KP + lane];
        if (A) w -= A[(size_t)level_of_row[k] * KP + lane];
    }
    double acc = 0.0;
    const double* Dk = D ? D + (size_t)k * KP * KP : nullptr;
    for (int m = 0; m < KP; ++m) {
        const double wm = __shfl_sync(FULL, w, m);
        if (lane < KP) {
            double mv = G[m * KP + lane];                 // symmetric: column m read as row m (coalesced)
            if (Dk) mv -= Dk[m * KP + lane];
            acc = fma(mv, wm, acc);
        }
    }
    if (lane < KP) T[(size_t)k * KP + lane] = B[(size_t)k * KP + lane] - acc;
}

// one warp per level: assemble, solve, write A, shift the rows of U
__global__ void __launch_bounds__(32) k_level_solve(int K, int KP, int masked, const int* __restrict__ rows_sorted,
                                                    const int* __restrict__ level_start, double* __restrict__ A, const double* __restrict__ G,
                                                    const double* __restrict__ GL, const double* __restrict__ T, double lambda,
                                                    double* __restrict__ U, int* err_flag) {
    extern __shared__ double S[];                      // [KP][KP+1]
    const int ld = KP + 1;
    const int s = blockIdx.x, lane = threadIdx.x;
    const int b = level_start[s], e = level_start[s + 1];
    if (e == b) return;
    const double ns = (double)(e - b);
    for (int x = lane; x < KP * KP; x += 32) {
        const int r = x / KP, c = x % KP;
        double v = masked ? GL[(size_t)s * KP * KP + x] : ns * G[x];   // optimize.cpp:170 / :186
        if (r == c) v += lambda;                                       // :174 / :187
        S[r * ld + c] = v;
    }
    double rhs = 0.0;
    if (lane < KP)
        for (int r = b; r < e; ++r) rhs += T[(size_t)rows_sorted[r] * KP + lane];
    __syncwarp();
    const bool ok = chol_factor(S, ld, K, lane);
    const double x = chol_subst(S, ld, K, lane, rhs);                  // :175 / :190
    if (!ok && lane == 0) atomicExch(err_flag, 1);
    if (lane < K) {
        const double old = A[(size_t)s * KP + lane];
        const double dlt = x - old;
        A[(size_t)s * KP + lane] = x;
        for (int r = b; r < e; ++r) U[(size_t)rows_sorted[r] * KP + lane] += dlt;
    }
}

// continuous covariate, stage 1: per chunk of 64 rows  H_c = sum x_k^2 M_k ,  T_c = sum x_k (B_k - M_k u_k)
__global__ void __launch_bounds__(256) k_cont_partial(int N, int KP, const double* __restrict__ x, const double* __restrict__ B,
                                                      const double* __restrict__ G, const double* __restrict__ D, const double* __restrict__ U,
                                                      double* __restrict__ scratch) {
    __shared__ double tw[8][32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r0 = blockIdx.x * 64, r1 = min(N, r0 + 64);
    const int KK = KP * KP;
    double* out = scratch + (size_t)blockIdx.x * (KK + KP);
    // H: element-parallel, rows in order
    for (int e = tid; e < KK; e += 256) {
        const double gx = G[e];
        double acc = 0.0;
        for (int k = r0; k < r1; ++k) { const double xk = x[k]; acc = fma(xk * xk, D ? gx - D[(size_t)k * KK + e] : gx, acc); }
        out[e] = acc;
    }
    // T: warp w takes rows r0+w, r0+w+8, ... ; then the 8 warps are combined in order
    double tacc = 0.0;
    for (int k = r0 + warp; k < r1; k += 8) {
        const double w = (lane < KP) ? U[(size_t)k * KP + lane] : 0.0;
        const double* Dk = D ? D + (size_t)k * KK : nullptr;
        double acc = 0.0;
        for (int m = 0; m < KP; ++m) {
            const double wm = __shfl_sync(FULL, w, m);
            if (lane < KP) { double mv = G[m * KP + lane]; if (Dk) mv -= Dk[m * KP + lane]; acc = fma(mv, wm, acc); }
        }
        if (lane < KP) tacc = fma(x[k], B[(size_t)k * KP + lane] - acc, tacc);
    }
    tw[warp][lane] = tacc;
    __syncthreads();
    if (tid < KP) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += tw[w][tid];
        out[KK + tid] = s;
    }
}

// continuous covariate, stage 2 (one block): reduce chunks, update w (src/optimize.cpp:102-126 or :127-131), shift U
__global__ void __launch_bounds__(256) k_cont_final(int N, int K, int KP, int masked, int n_chunks, const double* __restrict__ x,
                                                    const double* __restrict__ scratch, double lambda, double* __restrict__ w,
                                                    double* __restrict__ U, int* err_flag) {
    extern __shared__ double sm[];                     // H [KP][KP+1], Tq [KP], dw [KP]
    const int ld = KP + 1, KK = KP * KP;
    double* H = sm; double* Tq = H + KP * ld; double* dw = Tq + KP;
    const int tid = threadIdx.x, lane = tid & 31;
    for (int e = tid; e < KK + KP; e += 256) {
        double s = 0.0;
        for (int c = 0; c < n_chunks; ++c) s += scratch[(size_t)c * (KK + KP) + e];
        if (e < KK) H[(e / KP) * ld + (e % KP)] = s; else Tq[e - KK] = s;
    }
    __syncthreads();
    if (tid < 32) {
        double wl = (lane < K) ? w[lane] : 0.0;
        const double w_old = wl;
        if (masked) {
            double tl = (lane < K) ? Tq[lane] : 0.0;
            for (int it = 0; it < 100000; ++it) {                                // while(1)  :102
                const double pre = wl;
                for (int i = 0; i < K; ++i) {                                    // cyclic order  :104
                    const double hii = H[i * ld + i];
                    const double wi = __shfl_sync(FULL, wl, i), ti = __shfl_sync(FULL, tl, i);
                    const double xty = fma(wi, hii, ti);                         // :111
                    const double nw = xty / (hii + lambda);                      // :117
                    const double dlt = nw - wi;
                    if (lane == i) wl = nw;
                    if (lane < K) tl = fma(-dlt, H[i * ld + lane], tl);          // :118 in statistic form
                }
                const double diff = warp_sum((lane < K) ? fabs(pre - wl) : 0.0);
                if (diff < 1e-1) break;                                          // :122
            }
        } else {
            // (x'x G + lambda I) w = V data' x = Tq + H w_old     :127-131
            double rhs = (lane < K) ? Tq[lane] : 0.0;
            for (int m = 0; m < K; ++m) { const double wm = __shfl_sync(FULL, w_old, m); if (lane < K) rhs = fma(H[m * ld + lane], wm, rhs); }
            if (lane < K) H[lane * ld + lane] += lambda;
            __syncwarp();
            const bool ok = chol_factor(H, ld, K, lane);
            wl = chol_subst(H, ld, K, lane, rhs);
            if (!ok && lane == 0) atomicExch(err_flag, 1);
        }
        if (lane < KP) { dw[lane] = (lane < K) ? wl - w_old : 0.0; if (lane < K) w[lane] = wl; }
    }
    __syncthreads();
    for (int64_t e = tid; e < (int64_t)N * KP; e += 256) {
        const int k = (int)(e / KP), l = (int)(e % KP);
        U[e] = fma(x[k], dw[l], U[e]);
    }
}

__global__ void __launch_bounds__(256) k_build_u(int N, int KP, int ldT, int C, const RowDesign* __restrict__ designs, int Q,
                                                 const double* __restrict__ X, const double* __restrict__ W, double* __restrict__ U,
                                                 double* __restrict__ Ut) {
    const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (e >= (int64_t)N * KP) return;
    const int k = (int)(e / KP), l = (int)(e % KP);
    double s = 0.0;
    for (int c = 0; c < C; ++c) s += designs[c].A[(size_t)designs[c].level_of_row[k] * KP + l];      // optimize.cpp:366-369
    for (int q = 0; q < Q; ++q) s = fma(X[(size_t)q * N + k], W[(size_t)q * KP + l], s);             // :371-373
    U[e] = s;
    Ut[(size_t)l * ldT + k] = s;
}

}  // namespace

void launch_level_gram(const Geom& g, const RowDesign& d, const double* G, const double* D, double* GL, cudaStream_t st) {
    k_level_gram<<<d.L, 256, 0, st>>>(d.rows_sorted, d.level_start, G, D, GL, g.KP * g.KP);
}

void launch_row_rhs(const Geom& g, bool masked, const RowDesign& d, const double* B, const double* G, const double* D, const double* U,
                    double* T, cudaStream_t st) {
    k_row_rhs<<<(g.N + 7) / 8, 256, 0, st>>>(g.N, g.KP, d.level_of_row, d.A, B, G, masked ? D : nullptr, U, T);
}

void launch_level_solve(const Geom& g, bool masked, const RowDesign& d, const double* G, const double* GL, const double* T, double lambda,
                        double* U, int* err_flag, cudaStream_t st) {
    const size_t smem = (size_t)g.KP * (g.KP + 1) * 8;
    k_level_solve<<<d.L, 32, smem, st>>>(g.K, g.KP, masked ? 1 : 0, d.rows_sorted, d.level_start, d.A, G, GL, T, lambda, U, err_flag);
}

size_t continuous_scratch_elems(const Geom& g) { return (size_t)((g.N + 63) / 64) * (g.KP * g.KP + g.KP); }

void launch_continuous(const Geom& g, bool masked, const double* x, double* w, const double* B, const double* G, const double* D,
                       double lambda, double* U, double* scratch, int* err_flag, cudaStream_t st) {
    const int n_chunks = (g.N + 63) / 64;
    k_cont_partial<<<n_chunks, 256, 0, st>>>(g.N, g.KP, x, B, G, masked ? D : nullptr, U, scratch);
    const size_t smem = ((size_t)g.KP * (g.KP + 1) + 2 * g.KP) * 8;
    k_cont_final<<<1, 256, smem, st>>>(g.N, g.K, g.KP, masked ? 1 : 0, n_chunks, x, scratch, lambda, w, U, err_flag);
}

void launch_build_u(const Geom& g, int C, const RowDesign* designs_dev, int Q, const double* X, const double* W, double* U, double* Ut,
                    cudaStream_t st) {
    const int64_t n = (int64_t)g.N * g.KP;
    k_build_u<<<(int)((n + 255) / 256), 256, 0, st>>>(g.N, g.KP, g.ldT, C, designs_dev, Q, X, W, U, Ut);
}

}  // namespace ib
