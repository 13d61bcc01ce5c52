// Row-side (confounder factor) updates in sufficient-statistic form.
//
// The reference adds one confounder's contribution back to a dense N x P residual, solves per level, and subtracts
// it again (src/optimize.cpp:335-362), i.e. 2C-1 read-modify-write passes over N x P per iteration. With
//     B_k  = sum_j m_kj y_kj v_j          (k_row_b)
//     Gk_k = sum_j m_kj v_j v_j^T = G - D_k   (accumulated in k_row_b; k_row_comp_gram)
// the normal equations of level s of confounder c (src/optimize.cpp:150-176 / :178-191) are
//     XtX_s = sum_{k in s} Gk_k + lambda I ,   Xty_s = sum_{k in s} [ B_k - Gk_k (u_k - a_{c,s}) ]
// where u_k is the current row factor (Gauss-Seidel: blocks updated earlier in the same iteration are already in u_k).
// In the dense path (tuning = 0) Gk_k = G for every row, so Xty_s = sum B_k - G (sum u_k - n_s a_s).
//
// XtX_s does not depend on the factors, so all levels of all confounders are assembled and Cholesky-factorised up front
// by one launch (k_level_factor: one warp per level, batched K x K factorisations in shared memory); the per-confounder
// launches then only form right-hand sides, substitute, and shift the rows of U.
#include <cooperative_groups.h>

#include "common.cuh"
#include "kernels.cuh"

namespace cg = cooperative_groups;

namespace ib {

namespace {

constexpr int LG_ROWS = 32;      // rows per k_level_gram chunk

// warp Cholesky on an odd-pitch shared matrix (lane = row). Each lane only ever writes its own row, the pivot column
// travels by shuffle, so the trailing update needs no barrier inside the m loop.
__device__ bool chol_factor(double* S, int ld, int K, int lane) {
    bool ok = true;
    for (int j = 0; j < K; ++j) {
        __syncwarp();
        const double ajj = S[j * ld + j];
        if (!(ajj > 0.0)) ok = false;
        const double dj = sqrt(ajj);
        double lij = 0.0;
        if (lane == j) S[j * ld + j] = dj;
        if (lane > j && lane < K) { lij = S[lane * ld + j] / dj; S[lane * ld + j] = lij; }
        for (int m = j + 1; m < K; ++m) {
            const double lmj = __shfl_sync(FULL, lij, m);
            if (lane >= m && lane < K) S[lane * ld + m] = fma(-lij, lmj, S[lane * ld + m]);
        }
    }
    __syncwarp();
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
// substitution with a stored factor (staged in shared memory): Lf[KP*KP] row-major lower triangle, Lf[KP*KP + i] = 1 / L_ii
__device__ double chol_subst_stored(const double* Lf, int KP, int K, int lane, double b) {
    const double* inv = Lf + KP * KP;
    for (int i = 0; i < K; ++i) {
        double xi = b * inv[i];
        xi = __shfl_sync(FULL, xi, i);
        if (lane == i) b = xi;
        else if (lane > i && lane < K) b = fma(-Lf[lane * KP + i], xi, b);
    }
    for (int i = K - 1; i >= 0; --i) {
        double xi = b * inv[i];
        xi = __shfl_sync(FULL, xi, i);
        if (lane == i) b = xi;
        else if (lane < i) b = fma(-Lf[i * KP + lane], xi, b);
    }
    return b;
}

// The same substitution with the factor in registers (lane = row: its row for the forward pass, its column of L for the
// backward pass): per step one multiply, one shuffle, one FMA instead of two dependent shared-memory loads around them.
// Operation for operation identical to chol_subst_stored.
template <int KP>
__device__ __forceinline__ double chol_subst_regs(const double* __restrict__ Lf, int K, int lane, double b) {
    double Lrow[KP], Lcol[KP];
    const int lr = (lane < KP) ? lane : 0;
#pragma unroll
    for (int i = 0; i < KP; i += 2) {
        const double2 v = *reinterpret_cast<const double2*>(Lf + lr * KP + i);
        Lrow[i] = v.x; Lrow[i + 1] = v.y;
    }
#pragma unroll
    for (int i = 0; i < KP; ++i) Lcol[i] = Lf[i * KP + lr];
    const double inv = Lf[KP * KP + lr];
    // no `i < K` guards: for the padding rows K..KP-1 the stored inverse diagonal is 0 and the right-hand side is 0, so their
    // steps are exact no-ops - and the whole substitution is one basic block
#pragma unroll
    for (int i = 0; i < KP; ++i) {
        const double xi = __shfl_sync(FULL, b * inv, i);
        const double upd = fma(-Lrow[i], xi, b);
        b = (lane == i) ? xi : ((lane > i && lane < K) ? upd : b);
    }
#pragma unroll
    for (int i = KP - 1; i >= 0; --i) {
        const double xi = __shfl_sync(FULL, b * inv, i);
        const double upd = fma(-Lcol[i], xi, b);
        b = (lane == i) ? xi : ((lane < i) ? upd : b);
    }
    return b;
}
// asynchronous global -> shared copy of n doubles (n even, 16-byte aligned both sides) by one warp; gs_async_wait() before use
__device__ __forceinline__ void gs_async_copy(double* dst, const double* src, int n, int lane) {
    for (int x = 2 * lane; x < n; x += 64)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(dst + x)), "l"(src + x) : "memory");
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}
__device__ __forceinline__ void gs_async_wait() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// masked path: GLp[level][chunk][e] = sum over the chunk's rows of (G[e] - D[row][e]); grid (total levels, max chunks)
__global__ void __launch_bounds__(256) k_level_gram(const LevelTable* __restrict__ tab, const double* __restrict__ G, const double* __restrict__ D,
                                                    double* __restrict__ GLp, int KK, int max_chunks) {
    __shared__ int rows[LG_ROWS];
    const LevelTable t = tab[blockIdx.x];
    const int b = t.row_begin + blockIdx.y * LG_ROWS, e = min(t.row_end, b + LG_ROWS);
    if (b >= e) return;
    if (threadIdx.x < e - b) rows[threadIdx.x] = t.rows_sorted[b + threadIdx.x];
    __syncthreads();
    const int n = e - b;
    double* out = GLp + ((size_t)blockIdx.x * max_chunks + blockIdx.y) * KK;
    for (int x = threadIdx.x; x < KK; x += 256) {
        const double gx = G[x];
        double acc = 0.0;
#pragma unroll 4
        for (int r = 0; r < n; ++r) acc += gx - D[(size_t)rows[r] * KK + x];
        out[x] = acc;
    }
}

// one warp per level (all confounders): assemble XtX_s + lambda I, factorise, store L and 1/diag
__global__ void __launch_bounds__(32) k_level_factor(const LevelTable* __restrict__ tab, int K, int KP, int masked, const double* __restrict__ G,
                                                     const double* __restrict__ GLp, int max_chunks, double lambda, double* __restrict__ Lfac,
                                                     int* err_flag) {
    extern __shared__ double S[];                      // [KP][KP+1]
    const int ld = KP + 1, KK = KP * KP, lane = threadIdx.x;
    const LevelTable t = tab[blockIdx.x];
    const int n_rows = t.row_end - t.row_begin;
    const int n_chunks = (n_rows + LG_ROWS - 1) / LG_ROWS;
    for (int x = lane; x < KK; x += 32) {
        const int r = x / KP, c = x % KP;
        double v;
        if (masked) {                                                          // optimize.cpp:170
            v = 0.0;
            for (int y = 0; y < n_chunks; ++y) v += GLp[((size_t)blockIdx.x * max_chunks + y) * KK + x];
        } else {
            v = (double)n_rows * G[x];                                         // :186
        }
        if (r == c) v += lambda;                                               // :174 / :187
        S[r * ld + c] = v;
    }
    __syncwarp();
    const bool ok = chol_factor(S, ld, K, lane);
    if (!ok && lane == 0) atomicExch(err_flag, 1);
    double* out = Lfac + (size_t)blockIdx.x * (KK + KP);
    for (int x = lane; x < KK; x += 32) out[x] = S[(x / KP) * ld + (x % KP)];
    if (lane < KP) out[KK + lane] = (lane < K) ? 1.0 / S[lane * ld + lane] : 0.0;
}

// masked path: warp per row, T_k = B_k - (G - D_k)(u_k - a_{c,z(k)})
__global__ void __launch_bounds__(256) k_row_rhs(int N, int KP, const int* __restrict__ level_of_row, const double* __restrict__ A,
                                                 const double* __restrict__ B, const double* __restrict__ G, const double* __restrict__ D,
                                                 const double* __restrict__ U, double* __restrict__ T) {
    const int lane = threadIdx.x & 31;
    const int k = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (k >= N) return;
    double w = 0.0;
    if (lane < KP) w = U[(size_t)k * KP + lane] - A[(size_t)level_of_row[k] * KP + lane];
    double acc = 0.0;
    const double* Dk = D + (size_t)k * KP * KP;
    for (int m = 0; m < KP; ++m) {
        const double wm = __shfl_sync(FULL, w, m);
        if (lane < KP) acc = fma(G[m * KP + lane] - Dk[m * KP + lane], wm, acc);   // symmetric: column m read as row m (coalesced)
    }
    if (lane < KP) T[(size_t)k * KP + lane] = B[(size_t)k * KP + lane] - acc;
}

// per level of one confounder (block of 4 warps): right-hand side, substitution with the stored factor, A and U update
//   masked: rhs = sum_{k in s} T_k                         (T from k_row_rhs)
//   dense : rhs = sum_{k in s} B_k - G (sum_{k in s} u_k - n_s a_s)
__global__ void __launch_bounds__(128) k_level_update(int K, int KP, int masked, const int* __restrict__ rows_sorted,
                                                      const int* __restrict__ level_start, double* __restrict__ A, const double* __restrict__ G,
                                                      const double* __restrict__ B, const double* __restrict__ T, const double* __restrict__ Lfac,
                                                      int lfac_base, double* __restrict__ U) {
    constexpr int RCH = 256;                           // rows staged per pass
    __shared__ int rows[RCH];
    __shared__ double part[4][2][32];
    __shared__ double delta_s[32];
    __shared__ double Lf_s[32 * 32 + 32];              // this level's Cholesky factor + inverse diagonal
    const int s = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = level_start[s], e = level_start[s + 1];
    if (e == b) return;
    {
        const double* src = Lfac + (size_t)(lfac_base + s) * (KP * KP + KP);
        copy_batched(Lf_s, src, KP * KP + KP, threadIdx.x, 128);
    }
    double p0 = 0.0, p1 = 0.0;     // masked: p0 = sum T ; dense: p0 = sum B, p1 = sum u
    const double* src0 = masked ? T : B;
    for (int c0 = b; c0 < e; c0 += RCH) {
        const int n = min(RCH, e - c0);
        __syncthreads();
        for (int x = threadIdx.x; x < n; x += 128) rows[x] = rows_sorted[c0 + x];
        __syncthreads();
        if (lane < KP) {
            int r = warp;
            for (; r + 12 < n; r += 16) {              // four rows in flight per warp
                const int k0 = rows[r], k1 = rows[r + 4], k2 = rows[r + 8], k3 = rows[r + 12];
                const double a0 = src0[(size_t)k0 * KP + lane], a1 = src0[(size_t)k1 * KP + lane], a2 = src0[(size_t)k2 * KP + lane], a3 = src0[(size_t)k3 * KP + lane];
                p0 += a0; p0 += a1; p0 += a2; p0 += a3;
                if (!masked) {
                    const double u0 = U[(size_t)k0 * KP + lane], u1 = U[(size_t)k1 * KP + lane], u2 = U[(size_t)k2 * KP + lane], u3 = U[(size_t)k3 * KP + lane];
                    p1 += u0; p1 += u1; p1 += u2; p1 += u3;
                }
            }
            for (; r < n; r += 4) {
                const int k = rows[r];
                p0 += src0[(size_t)k * KP + lane];
                if (!masked) p1 += U[(size_t)k * KP + lane];
            }
        }
    }
    part[warp][0][lane] = p0; part[warp][1][lane] = p1;
    __syncthreads();
    if (warp == 0) {
        double rhs = (part[0][0][lane] + part[1][0][lane]) + (part[2][0][lane] + part[3][0][lane]);
        const double a_old = (lane < KP) ? A[(size_t)s * KP + lane] : 0.0;
        if (!masked) {
            const double su = (part[0][1][lane] + part[1][1][lane]) + (part[2][1][lane] + part[3][1][lane]);
            const double w = su - (double)(e - b) * a_old;
            double acc = 0.0;
            for (int m = 0; m < KP; ++m) {
                const double wm = __shfl_sync(FULL, w, m);
                if (lane < KP) acc = fma(G[m * KP + lane], wm, acc);
            }
            rhs -= acc;
        }
        const double x = chol_subst_stored(Lf_s, KP, K, lane, rhs);   // :175 / :190
        double dlt = 0.0;
        if (lane < K) { dlt = x - a_old; A[(size_t)s * KP + lane] = x; }
        delta_s[lane] = dlt;
    }
    __syncthreads();
    const double dlt = delta_s[lane];
    for (int c0 = b; c0 < e; c0 += RCH) {
        const int n = min(RCH, e - c0);
        if (e - b > RCH) {                             // indices of this pass (single pass: still staged from above)
            __syncthreads();
            for (int x = threadIdx.x; x < n; x += 128) rows[x] = rows_sorted[c0 + x];
            __syncthreads();
        }
        if (lane < KP) {
            int r = warp;
            for (; r + 12 < n; r += 16) {
                double* q0 = U + (size_t)rows[r] * KP + lane; double* q1 = U + (size_t)rows[r + 4] * KP + lane;
                double* q2 = U + (size_t)rows[r + 8] * KP + lane; double* q3 = U + (size_t)rows[r + 12] * KP + lane;
                const double u0 = *q0, u1 = *q1, u2 = *q2, u3 = *q3;
                *q0 = u0 + dlt; *q1 = u1 + dlt; *q2 = u2 + dlt; *q3 = u3 + dlt;
            }
            for (; r < n; r += 4) U[(size_t)rows[r] * KP + lane] += dlt;
        }
    }
}

// ---- dense path (tuning = 0): the whole Gauss-Seidel sweep over confounders without touching N-length data ----------
// sum_{k in s} u_k = sum_{c' != c} sum_{s'} n(c,s;c',s') a_{c',s'} + n_s a_{c,s} + (sum_{k in s} x_k) W, where n(.) are the
// co-occurrence counts of the design (fixed). The right-hand side of level s of confounder c is therefore
//     SB_{c,s} - G ( sum_{c' != c} sum_{s'} n a_{c',s'} + Sx_{c,s} W ),        SB_{c,s} = sum_{k in s} B_k
// k_level_sumB forms SB for every level of every confounder in one launch; k_rows_dense_gs then runs the C block updates
// back to back in ONE block (one warp per level, __syncthreads between confounders), reading only factor-sized data.
__global__ void __launch_bounds__(128) k_level_sumB(const LevelTable* __restrict__ tab, int KP, const double* __restrict__ B, double* __restrict__ SB) {
    constexpr int RCH = 256;
    __shared__ int rows[RCH];
    __shared__ double part[4][32];
    const LevelTable t = tab[blockIdx.x];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double p0 = 0.0;
    for (int c0 = t.row_begin; c0 < t.row_end; c0 += RCH) {
        const int n = min(RCH, t.row_end - c0);
        __syncthreads();
        for (int x = threadIdx.x; x < n; x += 128) rows[x] = t.rows_sorted[c0 + x];
        __syncthreads();
        if (lane < KP) {
            int r = warp;
            for (; r + 12 < n; r += 16) {
                const double a0 = B[(size_t)rows[r] * KP + lane], a1 = B[(size_t)rows[r + 4] * KP + lane];
                const double a2 = B[(size_t)rows[r + 8] * KP + lane], a3 = B[(size_t)rows[r + 12] * KP + lane];
                p0 += a0; p0 += a1; p0 += a2; p0 += a3;
            }
            for (; r < n; r += 4) p0 += B[(size_t)rows[r] * KP + lane];
        }
    }
    part[warp][lane] = p0;
    __syncthreads();
    if (warp == 0 && lane < KP) SB[(size_t)blockIdx.x * KP + lane] = (part[0][lane] + part[1][lane]) + (part[2][lane] + part[3][lane]);
}

struct DenseGsArgs {
    int C, K, KP, Q;
    const int* lvl_first;        // [C+1] first global level index of each confounder
    const int* co_ptr;           // [total_levels+1] CSR over levels: co-occurring (other-confounder) levels
    const int* co_row;           // [nnz] row of A_all (global level index) of the co-occurring level
    const double* co_cnt;        // [nnz] number of samples in both levels
    const double* Sx;            // [total_levels][Q] sum of each continuous covariate over the level's samples (or null)
    const double* W;             // [Q][KP] continuous factor (or null)
    double* A_all;               // [total_levels][KP]
    const double* SB;            // [total_levels][KP]
    const double* G;             // [KP*KP]
    const double* Lfac;          // [total_levels][KP*KP + KP]
    int a_in_smem;               // 1: the categorical factors fit in shared memory next to the per-warp factor buffers
    int n_lbuf;                  // factor buffers per warp (2: the next confounder's factor is staged ahead)
    int csr_in_smem;             // 1: co_row / co_cnt staged in shared memory
    // fused row-factor rebuild (fast path, no continuous covariates): U = sum_c A_c[z_c], Ut, UtU (src/optimize.cpp:365-369)
    int fuse_u, N, ldT;
    const int* row_lv;           // [N][C] global level index (row of A_all) of every (sample, confounder)
    double* U; double* Ut; double* UtU;
};
constexpr int GS_CLUSTER = 8;       // CTAs per cluster (portable maximum)
constexpr int GS_WARPS = 16;        // warps per CTA -> 128 levels in flight

// one level: w = sum over co-occurring levels of count * a_{c',s'} (+ Sx W), rhs = SB - G w, substitution with the staged factor
template <int KPT>
__device__ __forceinline__ double gs_level_solve(const DenseGsArgs& a, int lv, const double* As, bool a_in_smem, const double* Gs, const double* Lw, int lane,
                                                 const int* co_row, const double* co_cnt, int e0, int e1, double sb) {
    const int KP = a.KP, K = a.K;
    double w = 0.0;
    for (int eb = e0; eb < e1; eb += 32) {                                     // 32 CSR entries per batch, broadcast by shuffle
        const int n = min(32, e1 - eb);
        const int myrow = (lane < n) ? co_row[eb + lane] : 0;
        const double mycnt = (lane < n) ? co_cnt[eb + lane] : 0.0;
        int j = 0;
        for (; j + 3 < n; j += 4) {
            const int r0 = __shfl_sync(FULL, myrow, j), r1 = __shfl_sync(FULL, myrow, j + 1), r2 = __shfl_sync(FULL, myrow, j + 2), r3 = __shfl_sync(FULL, myrow, j + 3);
            const double c0 = __shfl_sync(FULL, mycnt, j), c1 = __shfl_sync(FULL, mycnt, j + 1), c2 = __shfl_sync(FULL, mycnt, j + 2), c3 = __shfl_sync(FULL, mycnt, j + 3);
            double v0 = 0, v1 = 0, v2 = 0, v3 = 0;
            if (lane < KP) {
                if (a_in_smem) { v0 = As[(size_t)r0 * KP + lane]; v1 = As[(size_t)r1 * KP + lane]; v2 = As[(size_t)r2 * KP + lane]; v3 = As[(size_t)r3 * KP + lane]; }
                else { v0 = __ldcg(As + (size_t)r0 * KP + lane); v1 = __ldcg(As + (size_t)r1 * KP + lane); v2 = __ldcg(As + (size_t)r2 * KP + lane); v3 = __ldcg(As + (size_t)r3 * KP + lane); }
            }
            w = fma(c0, v0, w); w = fma(c1, v1, w); w = fma(c2, v2, w); w = fma(c3, v3, w);
        }
        for (; j < n; ++j) {
            const int r0 = __shfl_sync(FULL, myrow, j);
            const double c0 = __shfl_sync(FULL, mycnt, j);
            if (lane < KP) w = fma(c0, a_in_smem ? As[(size_t)r0 * KP + lane] : __ldcg(As + (size_t)r0 * KP + lane), w);
        }
    }
    if (lane < KP) for (int q = 0; q < a.Q; ++q) w = fma(a.Sx[(size_t)lv * a.Q + q], a.W[(size_t)q * KP + lane], w);
    double acc = 0.0;
    const int lg = (lane < KPT) ? lane : 0;
#pragma unroll
    for (int m = 0; m < KPT; ++m) {                                            // G w, same order of accumulation as before
        const double wm = __shfl_sync(FULL, w, m);
        acc = fma(Gs[m * KPT + lg], wm, acc);
    }
    const double rhs = (lane < KP) ? sb - acc : 0.0;
    __syncwarp();
    return chol_subst_regs<KPT>(Lw, K, lane, rhs);                             // :190
}

template <int KPT>
__global__ void __cluster_dims__(GS_CLUSTER, 1, 1) __launch_bounds__(GS_WARPS * 32, 1) k_rows_dense_gs(DenseGsArgs a) {
    // shared memory: G [KP*KP] | A copy [total_levels*KP] (if a.a_in_smem) | per-warp factor buffers [a.n_lbuf][KP*KP + KP]
    extern __shared__ double gs_smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int KP = a.KP, K = a.K, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gw = (int)cluster.block_rank() * GS_WARPS + warp, n_gw = GS_CLUSTER * GS_WARPS;
    const int FK = KP * KP + KP;
    const int n_lv = a.lvl_first[a.C];
    double* Gs = gs_smem;
    double* As = gs_smem + KP * KP;                                            // this CTA's copy of all categorical factors
    double* Lw0 = As + (a.a_in_smem ? (size_t)n_lv * KP : 0) + (size_t)warp * a.n_lbuf * FK;
    for (int x = threadIdx.x; x < KP * KP; x += blockDim.x) Gs[x] = a.G[x];
    // the design's co-occurrence lists (constant), staged once when they fit: a level of a small confounder co-occurs with every
    // level of the others (131 entries for the 2-level one), and every 32-entry batch read from global memory is an L2 round trip
    const int nnz = a.co_ptr[n_lv];
    const double* co_cnt = a.co_cnt; const int* co_row = a.co_row;
    if (a.csr_in_smem) {
        double* cnt_s = Lw0 - (size_t)warp * a.n_lbuf * FK + (size_t)GS_WARPS * a.n_lbuf * FK;      // after all factor buffers
        int* row_s = reinterpret_cast<int*>(cnt_s + nnz);
        for (int x = threadIdx.x; x < nnz; x += blockDim.x) { cnt_s[x] = a.co_cnt[x]; row_s[x] = a.co_row[x]; }
        co_cnt = cnt_s; co_row = row_s;
    }
    // Fast path (a.a_in_smem && a.n_lbuf == 2): every CTA keeps all factors in shared memory; a solved level is written into
    // the copies of all 8 CTAs through distributed shared memory, so after the cluster barrier nothing has to be re-read from
    // global memory, and the Cholesky factor of a warp's level of the NEXT confounder is staged while this one is solved.
    const bool fast = a.a_in_smem && a.n_lbuf == 2;
    if (a.a_in_smem) {
        int x = threadIdx.x; const int nthr = blockDim.x, n = n_lv * KP;
        for (; x + 3 * nthr < n; x += 4 * nthr) {
            const double v0 = __ldcg(a.A_all + x), v1 = __ldcg(a.A_all + x + nthr), v2 = __ldcg(a.A_all + x + 2 * nthr), v3 = __ldcg(a.A_all + x + 3 * nthr);
            As[x] = v0; As[x + nthr] = v1; As[x + 2 * nthr] = v2; As[x + 3 * nthr] = v3;
        }
        for (; x < n; x += nthr) As[x] = __ldcg(a.A_all + x);
    }
    // scalars of this warp's first level of the next confounder, fetched one confounder ahead (global loads off the chain)
    int pe0 = 0, pe1 = 0; double psb = 0.0;
    auto prefetch_level = [&](int c) {
        const int lvn = a.lvl_first[c] + gw;
        if (lvn < a.lvl_first[c + 1]) {
            pe0 = a.co_ptr[lvn]; pe1 = a.co_ptr[lvn + 1];
            psb = (lane < KP) ? a.SB[(size_t)lvn * KP + lane] : 0.0;
            if (fast) gs_async_copy(Lw0 + (size_t)(c & 1) * FK, a.Lfac + (size_t)lvn * FK, FK, lane);
        }
    };
    if (a.C > 0) prefetch_level(0);
    cluster.sync();                                                            // every CTA's copy is complete before remote writes start
    for (int c = 0; c < a.C; ++c) {                                            // src/optimize.cpp:335 (fixed block order)
        if (!fast && a.a_in_smem && c > 0) {
            // slow path: re-read the factors (blocks updated earlier in this sweep were written by other CTAs of the cluster)
            int x = threadIdx.x; const int nthr = blockDim.x, n = n_lv * KP;
            for (; x < n; x += nthr) As[x] = __ldcg(a.A_all + x);
            __syncthreads();
        }
        const double* Asrc = a.a_in_smem ? As : a.A_all;
        int it = 0;
        for (int lv = a.lvl_first[c] + gw; lv < a.lvl_first[c + 1]; lv += n_gw, ++it) {
            double* Lw = Lw0 + (size_t)((fast && it == 0) ? (c & 1) : 0) * FK;
            int e0, e1; double sb;
            if (it == 0) { e0 = pe0; e1 = pe1; sb = psb; }
            else { e0 = a.co_ptr[lv]; e1 = a.co_ptr[lv + 1]; sb = (lane < KP) ? a.SB[(size_t)lv * KP + lane] : 0.0; }
            if (fast && it == 0) { gs_async_wait(); }                          // staged ahead (asynchronous copy)
            else { __syncwarp(); copy_batched(Lw, a.Lfac + (size_t)lv * FK, FK, lane, 32); }
            __syncwarp();
            if (it == 0 && c + 1 < a.C) prefetch_level(c + 1);                 // next confounder's level: factor + scalars
            const double x = gs_level_solve<KPT>(a, lv, Asrc, a.a_in_smem != 0, Gs, Lw, lane, co_row, co_cnt, e0, e1, sb);
            if (lane < K) {
                a.A_all[(size_t)lv * KP + lane] = x;
                if (fast) {
#pragma unroll
                    for (int r = 0; r < GS_CLUSTER; ++r) cluster.map_shared_rank(As, r)[(size_t)lv * KP + lane] = x;
                }
            }
            __syncwarp();
        }
        if (it == 0 && c + 1 < a.C) prefetch_level(c + 1);                     // no level in this confounder: still look ahead
        __threadfence();
        cluster.sync();                                                        // block c is complete and visible cluster-wide
    }
    if (a.fuse_u) {
        // Every CTA's shared copy now holds all updated factors: rebuild the row factor right here (k_build_u + k_gram_u_final
        // were two more dependent launches). CTA r takes a contiguous chunk of rows; the chunks' partial U'U are combined by
        // CTA 0 through distributed shared memory in rank order (fixed order: bitwise reproducible).
        const int rank = (int)cluster.block_rank();
        const int rows_per = (a.N + GS_CLUSTER - 1) / GS_CLUSTER;
        const int rb = rank * rows_per, re = min(a.N, rb + rows_per);
        const int ldu = KP + 1;
        double* us = As + (size_t)n_lv * KP;                                   // [rows_per][KP + 1]   (the factor buffers are idle now)
        double* part = us + (size_t)rows_per * ldu;                            // [KP * KP]
        for (int x = threadIdx.x; x < (re - rb) * KP; x += blockDim.x) {
            const int kr = x / KP, l = x % KP, k = rb + kr;
            double s = 0.0;
            for (int c = 0; c < a.C; ++c) s += As[(size_t)__ldg(a.row_lv + (size_t)k * a.C + c) * KP + l];   // optimize.cpp:366-369
            a.U[(size_t)k * KP + l] = s;
            a.Ut[(size_t)l * a.ldT + k] = s;
            us[kr * ldu + l] = s;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < KP * KP; e += blockDim.x) {
            const int ca = e / KP, cb = e % KP;
            double s = 0.0;
            for (int i = 0; i < re - rb; ++i) s = fma(us[i * ldu + ca], us[i * ldu + cb], s);
            part[e] = s;
        }
        cluster.sync();
        if (rank == 0) {
            for (int e = threadIdx.x; e < KP * KP; e += blockDim.x) {
                double s = 0.0;
#pragma unroll
                for (int r = 0; r < GS_CLUSTER; ++r) s += cluster.map_shared_rank(part, r)[e];
                a.UtU[e] = s;
            }
        }
        cluster.sync();                                                        // the partials stay alive until CTA 0 has read them
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
    for (int e = tid; e < KK; e += 256) {
        const double gx = G[e];
        double acc = 0.0;
        for (int k = r0; k < r1; ++k) { const double xk = x[k]; acc = fma(xk * xk, D ? gx - D[(size_t)k * KK + e] : gx, acc); }
        out[e] = acc;
    }
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

// U = sum_c A_c[z_c] + X W for a chunk of 32 rows; also Ut and the chunk's partial U^T U (fixed-order reduce follows)
__global__ void __launch_bounds__(256) k_build_u(int N, int KP, int ldT, int C, const RowDesign* __restrict__ designs, int Q,
                                                 const double* __restrict__ X, const double* __restrict__ W, double* __restrict__ U,
                                                 double* __restrict__ Ut, double* __restrict__ UtU_parts) {
    __shared__ double us[32][33];
    const int r0 = blockIdx.x * 32;
    for (int x = threadIdx.x; x < 32 * KP; x += 256) {
        const int kr = x / KP, l = x % KP, k = r0 + kr;
        double s = 0.0;
        if (k < N) {
            for (int c = 0; c < C; ++c) s += designs[c].A[(size_t)designs[c].level_of_row[k] * KP + l];   // optimize.cpp:366-369
            for (int q = 0; q < Q; ++q) s = fma(X[(size_t)q * N + k], W[(size_t)q * KP + l], s);          // :371-373
            U[(size_t)k * KP + l] = s;
            Ut[(size_t)l * ldT + k] = s;
        }
        us[kr][l] = s;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < KP * KP; e += 256) {
        const int a = e / KP, b = e % KP;
        double s = 0.0;
#pragma unroll 8
        for (int i = 0; i < 32; ++i) s = fma(us[i][a], us[i][b], s);
        UtU_parts[(size_t)blockIdx.x * KP * KP + e] = s;
    }
}
__global__ void __launch_bounds__(256) k_gram_u_final(const double* __restrict__ parts, double* __restrict__ UtU, int n_parts, int KK) {
    for (int e = blockIdx.x * 256 + threadIdx.x; e < KK; e += gridDim.x * 256) {
        double s = 0.0;
        for (int p = 0; p < n_parts; ++p) s += parts[(size_t)p * KK + e];
        UtU[e] = s;
    }
}

}  // namespace

void launch_level_gram(const Geom& g, const LevelTable* tab_dev, int total_levels, int max_chunks, const double* G, const double* D, double* GLp,
                       cudaStream_t st) {
    dim3 grid(total_levels, max_chunks);
    k_level_gram<<<grid, 256, 0, st>>>(tab_dev, G, D, GLp, g.KP * g.KP, max_chunks);
}

void launch_level_factor(const Geom& g, bool masked, const LevelTable* tab_dev, int total_levels, int max_chunks, const double* G,
                         const double* GLp, double lambda, double* Lfac, int* err_flag, cudaStream_t st) {
    // (a register-resident variant measured 37 us against 19 us for this one: with one warp per level the
    // 276 shuffle + select steps of the unrolled trailing update are slower than the shared-memory read-modify-writes)
    const size_t smem = (size_t)g.KP * (g.KP + 1) * 8;
    k_level_factor<<<total_levels, 32, smem, st>>>(tab_dev, g.K, g.KP, masked ? 1 : 0, G, GLp, max_chunks, lambda, Lfac, err_flag);
}

void launch_row_rhs(const Geom& g, const RowDesign& d, const double* B, const double* G, const double* D, const double* U, double* T,
                    cudaStream_t st) {
    k_row_rhs<<<(g.N + 7) / 8, 256, 0, st>>>(g.N, g.KP, d.level_of_row, d.A, B, G, D, U, T);
}

void launch_level_update(const Geom& g, bool masked, const RowDesign& d, int lfac_base, const double* G, const double* B, const double* T,
                         const double* Lfac, double* U, cudaStream_t st) {
    k_level_update<<<d.L, 128, 0, st>>>(g.K, g.KP, masked ? 1 : 0, d.rows_sorted, d.level_start, d.A, G, B, T, Lfac, lfac_base, U);
}

void launch_level_sumB(const Geom& g, const LevelTable* tab_dev, int total_levels, const double* B, double* SB, cudaStream_t st) {
    if (total_levels) k_level_sumB<<<total_levels, 128, 0, st>>>(tab_dev, g.KP, B, SB);
}

bool rows_dense_gs_can_fuse_u(const Geom& g, int Q, int total_levels, int max_levels, int nnz) {
    // the fast path of k_rows_dense_gs (all factors + two factor buffers per warp in shared memory), no continuous covariates,
    // and a row chunk + K x K partial that fit into the idle factor buffers
    const size_t FK = (size_t)g.KP * g.KP + g.KP, limit = 227 * 1024;
    size_t smem = ((size_t)g.KP * g.KP + GS_WARPS * FK) * 8 + (size_t)total_levels * g.KP * 8;
    if (smem > limit) return false;
    if (smem + (size_t)nnz * 12 + 8 <= limit) smem += (size_t)nnz * 12 + 8;
    if (!(max_levels <= GS_CLUSTER * GS_WARPS && smem + GS_WARPS * FK * 8 <= limit)) return false;
    const size_t rows_per = (size_t)(g.N + GS_CLUSTER - 1) / GS_CLUSTER;
    return Q == 0 && rows_per * (g.KP + 1) + (size_t)g.KP * g.KP <= 2 * GS_WARPS * FK;
}

void launch_rows_dense_gs(const Geom& g, const DenseGs& d, int C, int Q, int total_levels, int max_levels, int nnz, double* A_all, const double* W,
                          const double* SB, const double* G, const double* Lfac, cudaStream_t st) {
    launch_rows_dense_gs_ex(g, d, C, Q, total_levels, max_levels, nnz, A_all, W, SB, G, Lfac, nullptr, nullptr, nullptr, nullptr, st);
}

void launch_rows_dense_gs_ex(const Geom& g, const DenseGs& d, int C, int Q, int total_levels, int max_levels, int nnz, double* A_all, const double* W,
                             const double* SB, const double* G, const double* Lfac, const int* row_lv, double* U, double* Ut, double* UtU,
                             cudaStream_t st) {
    DenseGsArgs a{C, g.K, g.KP, Q, d.lvl_first, d.co_ptr, d.co_row, d.co_cnt, d.Sx, W, A_all, SB, G, Lfac, 0, 1, 0, 0, g.N, g.ldT, row_lv, U, Ut, UtU};
    a.fuse_u = (U != nullptr && rows_dense_gs_can_fuse_u(g, Q, total_levels, max_levels, nnz)) ? 1 : 0;
    const size_t FK = (size_t)g.KP * g.KP + g.KP, limit = 227 * 1024;
    size_t smem = ((size_t)g.KP * g.KP + GS_WARPS * FK) * 8;
    if (smem + (size_t)total_levels * g.KP * 8 <= limit) { a.a_in_smem = 1; smem += (size_t)total_levels * g.KP * 8; }
    if (smem + (size_t)nnz * 12 + 8 <= limit) { a.csr_in_smem = 1; smem += (size_t)nnz * 12 + 8; }
    // second factor buffer (staging ahead): only when every warp has at most one level per confounder
    if (a.a_in_smem && max_levels <= GS_CLUSTER * GS_WARPS && smem + GS_WARPS * FK * 8 <= limit) { a.n_lbuf = 2; smem += GS_WARPS * FK * 8; }
#define GS_LAUNCH(KPv) { cudaFuncSetAttribute(k_rows_dense_gs<KPv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit); \
                         k_rows_dense_gs<KPv><<<GS_CLUSTER, GS_WARPS * 32, smem, st>>>(a); }
    switch (g.NT) { case 1: GS_LAUNCH(8) break; case 2: GS_LAUNCH(16) break; case 3: GS_LAUNCH(24) break; default: GS_LAUNCH(32) break; }
#undef GS_LAUNCH
}

size_t continuous_scratch_elems(const Geom& g) { return (size_t)((g.N + 63) / 64) * (g.KP * g.KP + g.KP); }

void launch_continuous(const Geom& g, bool masked, const double* x, double* w, const double* B, const double* G, const double* D,
                       double lambda, double* U, double* scratch, int* err_flag, cudaStream_t st) {
    const int n_chunks = (g.N + 63) / 64;
    k_cont_partial<<<n_chunks, 256, 0, st>>>(g.N, g.KP, x, B, G, masked ? D : nullptr, U, scratch);
    const size_t smem = ((size_t)g.KP * (g.KP + 1) + 2 * g.KP) * 8;
    k_cont_final<<<1, 256, smem, st>>>(g.N, g.K, g.KP, masked ? 1 : 0, n_chunks, x, scratch, lambda, w, U, err_flag);
}

int build_u_parts(const Geom& g) { return (g.N + 31) / 32; }

void launch_build_u(const Geom& g, int C, const RowDesign* designs_dev, int Q, const double* X, const double* W, double* U, double* Ut,
                    double* UtU, cudaStream_t st) {
    // UtU buffer: [KP*KP] result followed by build_u_parts(g) partial blocks
    const int n_parts = build_u_parts(g);
    double* parts = UtU + (size_t)g.KP * g.KP;
    k_build_u<<<n_parts, 256, 0, st>>>(g.N, g.KP, g.ldT, C, designs_dev, Q, X, W, U, Ut, parts);
    k_gram_u_final<<<(g.KP * g.KP + 255) / 256, 256, 0, st>>>(parts, UtU, n_parts, g.KP * g.KP);
}

}  // namespace ib
