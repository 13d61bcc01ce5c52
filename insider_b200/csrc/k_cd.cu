// Per-gene column update: masked Gram build + elastic-net coordinate descent / ridge solve.
//
//   replaces  optimize_col()               src/optimize.cpp:200-253
//             strong_coordinate_descent()  src/coordinate_descent.cpp:57-127
//             compute_loss(vec, vec, ..)   src/utils.cpp:46-49   (only loss *differences* are needed, see below)
//
// Arithmetic form. The reference keeps an explicit residual r = y - X beta (n ~ 0.9 N entries) and evaluates
// r.x_k per coordinate. Here the same iteration runs in covariance form on q = X'y - X'X beta (q_k == r.x_k):
//     upper_k = q_k + beta_k XtX_kk ; soft-threshold ; q -= (new - old) XtX[:,k]
// which needs K instead of n flops per coordinate and no access to X or y. The stopping rule
// |pre_loss - loss| <= tol (coordinate_descent.cpp:114) is evaluated from the exact per-coordinate loss decrement
//     dL_k = delta (delta XtX_kk / 2 - q_k) + lambda(1-alpha)(new^2 - old^2)/2 + lambda alpha (|new| - |old|)
// summed over the sweep, i.e. the same quantity without the cancellation of subtracting two O(|r|^2) numbers.
// The KKT re-admission test (coordinate_descent.cpp:118-119) is |q_e| > alpha lambda for excluded e because
// beta_e = 0 there. Visit order: counter-based permutation identical to the oracle's mode B.
//
// Mapping: 8 lanes per gene (4 genes per warp), lane li owns coordinates c = s*8 + li, s < SL = KP/8.
#include "common.cuh"
#include "kernels.cuh"

namespace ib {

namespace {

constexpr int LPG = 8;            // lanes per gene
constexpr int GPW = 4;            // genes per warp
constexpr int CD_WARPS = 4;       // warps per block
constexpr int MAX_SWEEPS = 200000;

__device__ __forceinline__ double grp_sum(double v) {
    v += __shfl_xor_sync(FULL, v, 4); v += __shfl_xor_sync(FULL, v, 2); v += __shfl_xor_sync(FULL, v, 1);
    return v;
}
__device__ __forceinline__ double grp_max(double v) {
    v = fmax(v, __shfl_xor_sync(FULL, v, 4)); v = fmax(v, __shfl_xor_sync(FULL, v, 2)); v = fmax(v, __shfl_xor_sync(FULL, v, 1));
    return v;
}
template <int SL> __device__ __forceinline__ double sel(const double (&a)[SL], int s) {
    double v = a[0];
#pragma unroll
    for (int i = 1; i < SL; ++i) v = (s == i) ? a[i] : v;
    return v;
}

// ---------------------------------------------------------------------------------------------------------------
// warp-level Cholesky (lane = row) on a K x K matrix in shared memory with leading dimension ld (odd).
__device__ bool warp_chol_factor(double* S, int ld, int K, int lane) {
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
// solves L L' x = b ; lane l holds b_l on entry and x_l on return
__device__ double warp_chol_subst(const double* S, int ld, int K, int lane, double b) {
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

// ---------------------------------------------------------------------------------------------------------------
// elastic-net coordinate descent for the 4 genes of a warp (one 8-lane group each).
//   Xs : this group's symmetric KP x KP matrix in shared memory (leading dimension ld, zero padded)
//   sh : per-group scratch, 5*KP doubles  [beta | q | d | den | 1/den] indexed by coordinate
//   ord_s : 32 bytes per group
// Data lives in two layouts. "Coordinate layout": lane li, slot s holds coordinate c = s*8 + li. "Position layout":
// lane li, slot s holds the coordinate visited at step i = s*8 + li of the current sweep (inactive ones behind the
// n_inc active ones). Each sweep re-labels registers into position layout through `sh`, so the unrolled step loop has a
// compile-time owner lane and slot: no dynamic register indexing anywhere in the sweep.
template <int SL>
__device__ void group_cd(const double* Xs, int ld, int K, int li, bool gvalid, const double (&xty)[SL], double (&beta)[SL], double lambda,
                         double alpha, double tol, int perm_mode, uint64_t seed, uint32_t als_iter, uint64_t gene, double* sh,
                         unsigned char* ord_s, int& sweeps_out) {
    constexpr int KP = SL * LPG;
    const int lane = threadIdx.x & 31;
    const int grp_shift = (lane >> 3) << 3;
    const double la = lambda * alpha, l2 = lambda * (1.0 - alpha);
    double* Bc = sh; double* Qc = sh + KP; double* Dc = sh + 2 * KP; double* DENc = sh + 3 * KP; double* RDc = sh + 4 * KP;
    uint32_t inc = 0;                                                       // active coordinates (group-uniform)
    {
        double mx = 0.0;
#pragma unroll
        for (int s = 0; s < SL; ++s) { const int c = s * LPG + li; mx = fmax(mx, (c < K) ? fabs(xty[s]) : 0.0); }
        mx = grp_max(mx);
        const double thr = alpha * (2.0 * lambda - mx);                    // coordinate_descent.cpp:74
        double q[SL];
#pragma unroll
        for (int s = 0; s < SL; ++s) {
            const int c = s * LPG + li;
            const bool a = (c < K) && !(fabs(xty[s]) < thr);
            if (!a) beta[s] = 0.0;                                         // :75-78
            const uint32_t b = __ballot_sync(FULL, a);
            inc |= ((b >> grp_shift) & 0xffu) << (LPG * s);
            q[s] = (c < K) ? xty[s] : 0.0;
        }
        // q = X'y - X'X beta
#pragma unroll
        for (int ms = 0; ms < SL; ++ms)
            for (int ml = 0; ml < LPG; ++ml) {
                const int m = ms * LPG + ml;
                const double bm = __shfl_sync(FULL, beta[ms], ml, LPG);
                if (m < K && bm != 0.0) {
#pragma unroll
                    for (int s = 0; s < SL; ++s) q[s] = fma(-Xs[m * ld + s * LPG + li], bm, q[s]);
                }
            }
#pragma unroll
        for (int s = 0; s < SL; ++s) {
            const int c = s * LPG + li;
            const double d = Xs[c * ld + c], den = d + l2;
            Bc[c] = beta[s]; Qc[c] = q[s]; Dc[c] = d; DENc[c] = den; RDc[c] = 1.0 / den;
        }
    }
    const uint32_t valid_mask = (K >= 32) ? 0xffffffffu : ((1u << K) - 1u);
    int n_inc = __popc(inc);
    bool done = !gvalid;
    uint32_t draw = 0;
    int sweeps = 0;
    while (true) {
        // ---- visiting order (coordinate_descent.cpp:89): rank of every coordinate, in coordinate layout
        int pos[SL];
        {
            uint32_t key[SL];
            const uint64_t pk = perm_key(seed, als_iter, gene, draw);
#pragma unroll
            for (int s = 0; s < SL; ++s) {
                const int c = s * LPG + li;
                const uint32_t below = (1u << c) - 1u;
                const int p_act = __popc(inc & below);
                key[s] = (perm_mode == 1) ? ((perm_value(pk, p_act) << 5) | (uint32_t)c) : (uint32_t)c;
                // inactive coordinates keep ascending order behind the active ones; padding (c >= K) stays last
                pos[s] = (c < K) ? n_inc + __popc(~inc & valid_mask & below) : c;
            }
            int rank[SL];
#pragma unroll
            for (int s = 0; s < SL; ++s) rank[s] = 0;
#pragma unroll
            for (int ms = 0; ms < SL; ++ms)
#pragma unroll
                for (int ml = 0; ml < LPG; ++ml) {
                    const int m = ms * LPG + ml;
                    const uint32_t km = __shfl_sync(FULL, key[ms], ml, LPG);
                    const uint32_t am = (inc >> m) & 1u;
#pragma unroll
                    for (int s = 0; s < SL; ++s) rank[s] += (int)(am & (uint32_t)(km < key[s]));
                }
#pragma unroll
            for (int s = 0; s < SL; ++s) if ((inc >> (s * LPG + li)) & 1u) pos[s] = rank[s];
        }
        ++draw;
        __syncwarp();
#pragma unroll
        for (int s = 0; s < SL; ++s) ord_s[pos[s]] = (unsigned char)(s * LPG + li);
        __syncwarp();
        // ---- gather into position layout
        int cd[SL];
        double b[SL], q[SL], d[SL], den[SL], rd[SL];
#pragma unroll
        for (int s = 0; s < SL; ++s) {
            const int c = ord_s[s * LPG + li];
            cd[s] = c; b[s] = Bc[c]; q[s] = Qc[c]; d[s] = Dc[c]; den[s] = DENc[c]; rd[s] = RDc[c];
        }
        int nmax = done ? 0 : n_inc;
        nmax = max(nmax, __shfl_xor_sync(FULL, nmax, 8));
        nmax = max(nmax, __shfl_xor_sync(FULL, nmax, 16));
        const int n_on = done ? 0 : n_inc;
        double dl = 0.0;
        // ---- one sweep; step i is owned by lane (i & 7), slot (i >> 3)
#pragma unroll
        for (int i = 0; i < KP; ++i) {
            if (i >= nmax) break;
            constexpr int dummy = 0; (void)dummy;
            const int si = i >> 3, ow = i & 7;
            const double bo = b[si], qk = q[si], dk = d[si];
            const double up = fma(bo, dk, qk);                             // :94
            const double t1 = fabs(up) - la;
            double nb = 0.0;
            if (t1 > 0.0) {                                                // :99-104
                const double num = copysign(t1, up);
                nb = num * rd[si];                                         // correctly rounded num / den (Markstein)
                nb = fma(fma(-den[si], nb, num), rd[si], nb);
            }
            double dlt = nb - bo;
            if (i >= n_on) dlt = 0.0;
            const double dkk = __shfl_sync(FULL, dlt, ow, LPG);
            const int kk = __shfl_sync(FULL, cd[si], ow, LPG);
            if (dkk != 0.0) {                                              // :106-109
                if (li == ow) {
                    const double s3 = fma(0.5 * l2, nb + bo, fma(0.5 * dk, dlt, -qk));
                    dl = fma(dlt, s3, dl);
                    dl = fma(la, fabs(nb) - fabs(bo), dl);
                    b[si] = nb;
                }
                const double* xr = Xs + kk * ld;
#pragma unroll
                for (int s = 0; s < SL; ++s) q[s] = fma(-dkk, xr[cd[s]], q[s]);
            }
        }
        // ---- scatter back to coordinate layout
#pragma unroll
        for (int s = 0; s < SL; ++s) { Bc[cd[s]] = b[s]; Qc[cd[s]] = q[s]; }
        const double delta = grp_sum(dl);
        // inner do-while ends (:114) -> KKT check on the excluded set (:118-124)
        const bool inner_end = !done && (!(fabs(delta) > tol) || sweeps + 1 >= MAX_SWEEPS);
        uint32_t vmask = 0;
        if (inner_end) {
#pragma unroll
            for (int s = 0; s < SL; ++s) {
                const int i = s * LPG + li;
                if (i >= n_inc && cd[s] < K && fabs(q[s]) > la) vmask |= 1u << cd[s];
            }
        }
        vmask |= __shfl_xor_sync(FULL, vmask, 4); vmask |= __shfl_xor_sync(FULL, vmask, 2); vmask |= __shfl_xor_sync(FULL, vmask, 1);
        if (!done) {
            ++sweeps;
            if (inner_end) {
                if (vmask == 0u || sweeps >= MAX_SWEEPS) done = true;
                else { inc |= vmask; n_inc = __popc(inc); }
            }
        }
        if (__all_sync(FULL, done)) break;
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < SL; ++s) beta[s] = Bc[s * LPG + li];
    sweeps_out = gvalid ? sweeps : 0;
}

struct SolveArgs {
    const uint32_t* trC; const double* U; const double* UtU; const double* Xty; double* V;
    int N, K, KP, ldV, Wp; int64_t P; int64_t gene0;
    double lambda, alpha; const double* tol; const uint32_t* als_iter; uint64_t seed; int perm_mode;
    unsigned long long* sweeps; int* err_flag;
};

// builds XtX_j = UtU - sum_{i: m_ij = 0} u_i u_i^T for one gene with the whole warp (DMMA rank-4 updates)
template <int SL>
__device__ void warp_masked_gram(const SolveArgs& a, int64_t gene, double* Xs, int ld) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    double acc[SL][SL][2];
#pragma unroll
    for (int i = 0; i < SL; ++i)
#pragma unroll
        for (int j = 0; j < SL; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const int nW = (a.N + 31) >> 5;
    int rows[4]; int cnt = 0;
    auto flush = [&]() {
        const int myrow = (t == 0) ? rows[0] : (t == 1) ? rows[1] : (t == 2) ? rows[2] : rows[3];
        double f[SL];
#pragma unroll
        for (int n = 0; n < SL; ++n) f[n] = (t < cnt) ? __ldg(a.U + (size_t)myrow * a.KP + 8 * n + g) : 0.0;
#pragma unroll
        for (int n1 = 0; n1 < SL; ++n1)
#pragma unroll
            for (int n2 = n1; n2 < SL; ++n2) dmma(acc[n1][n2][0], acc[n1][n2][1], f[n1], f[n2]);
        cnt = 0;
    };
    rows[0] = rows[1] = rows[2] = rows[3] = 0;
    for (int w0 = 0; w0 < nW; w0 += 32) {
        uint32_t z = 0;
        const int wi = w0 + lane;
        if (wi < nW) {
            z = ~__ldg(a.trC + gene * a.Wp + wi);
            const int lim = a.N - 32 * wi;
            if (lim < 32) z &= (1u << lim) - 1u;
        }
        const int wn = min(32, nW - w0);
        for (int w = 0; w < wn; ++w) {
            uint32_t zw = __shfl_sync(FULL, z, w);
            while (zw) {
                const int b = __ffs(zw) - 1;
                zw &= zw - 1;
                rows[cnt++] = 32 * (w0 + w) + b;
                if (cnt == 4) flush();
            }
        }
    }
    if (cnt > 0) flush();
#pragma unroll
    for (int n1 = 0; n1 < SL; ++n1)
#pragma unroll
        for (int n2 = n1; n2 < SL; ++n2)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int ra = 8 * n1 + g, cb = 8 * n2 + 2 * t + e;
                const double v = a.UtU[ra * a.KP + cb] - acc[n1][n2][e];
                Xs[ra * ld + cb] = v;
                Xs[cb * ld + ra] = v;
            }
    __syncwarp();
}

template <int SL, bool MASKED>
__global__ void __launch_bounds__(CD_WARPS * 32) k_col_solve(SolveArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int KP = SL * 8;
    constexpr int XLD = KP + 1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, grp = lane >> 3, li = lane & 7;
    double* Xall = reinterpret_cast<double*>(smem_raw);
    // MASKED: one matrix per gene [CD_WARPS][GPW][KP*XLD]; dense: one shared matrix (+ one factor for ridge)
    const int n_mats = MASKED ? CD_WARPS * GPW : 1;
    double* sh_all = Xall + (size_t)n_mats * KP * XLD;                    // [CD_WARPS*GPW][5*KP]
    unsigned char* ord_all = reinterpret_cast<unsigned char*>(sh_all + (size_t)CD_WARPS * GPW * 5 * KP);
    const int64_t j0 = ((int64_t)blockIdx.x * CD_WARPS + warp) * GPW;
    const double tol = *a.tol;
    const uint32_t als_iter = *a.als_iter;
    bool spd_ok = true;

    if (!MASKED) {
        for (int x = tid; x < KP * KP; x += blockDim.x) {
            const int r = x / KP, c = x % KP;
            double v = a.UtU[x];
            if (a.alpha == 0.0 && r == c) v += a.lambda;                    // optimize.cpp:238
            Xall[r * XLD + c] = v;
        }
        __syncthreads();
        if (a.alpha == 0.0) {
            if (warp == 0) spd_ok = warp_chol_factor(Xall, XLD, a.K, lane);
            __syncthreads();
        }
    }
    double* Xw = MASKED ? Xall + (size_t)warp * GPW * KP * XLD : Xall;
    if (MASKED) {
        for (int gi = 0; gi < GPW; ++gi)
            if (j0 + gi < a.P) warp_masked_gram<SL>(a, j0 + gi, Xw + (size_t)gi * KP * XLD, XLD);
    }
    if (a.alpha == 0.0) {
        // ridge: (XtX + lambda I) v = Xty  — optimize.cpp:224-226 (masked) / :237-240 (dense)
        for (int gi = 0; gi < GPW; ++gi) {
            const int64_t j = j0 + gi;
            if (j >= a.P) break;
            double b = (lane < a.K) ? a.Xty[j * a.ldV + lane] : 0.0;
            if (MASKED) {
                double* Xs = Xw + (size_t)gi * KP * XLD;
                if (lane < a.K) Xs[lane * XLD + lane] += a.lambda;
                __syncwarp();
                spd_ok &= warp_chol_factor(Xs, XLD, a.K, lane);
                b = warp_chol_subst(Xs, XLD, a.K, lane, b);
            } else {
                b = warp_chol_subst(Xall, XLD, a.K, lane, b);
            }
            if (lane < a.K) a.V[j * a.ldV + lane] = b;
        }
        if (!spd_ok && lane == 0) atomicExch(a.err_flag, 1);
        return;
    }
    const int64_t j = j0 + grp;
    const bool gvalid = j < a.P;
    double xty[SL], beta[SL];
#pragma unroll
    for (int s = 0; s < SL; ++s) {
        const int c = s * LPG + li;
        xty[s] = (gvalid && c < a.K) ? a.Xty[j * a.ldV + c] : 0.0;
        beta[s] = (gvalid && c < a.K) ? a.V[j * a.ldV + c] : 0.0;
    }
    const double* Xs = MASKED ? Xw + (size_t)grp * KP * XLD : Xall;
    int sweeps = 0;
    group_cd<SL>(Xs, XLD, a.K, li, gvalid, xty, beta, a.lambda, a.alpha, tol, a.perm_mode, a.seed, als_iter, (uint64_t)(a.gene0 + j),
                 sh_all + (size_t)(warp * GPW + grp) * 5 * KP, ord_all + (warp * GPW + grp) * 32, sweeps);
    if (gvalid) {
#pragma unroll
        for (int s = 0; s < SL; ++s) { const int c = s * LPG + li; if (c < a.K) a.V[j * a.ldV + c] = beta[s]; }
    }
    // sweep statistics: one atomic per warp
    int sw = (li == 0) ? sweeps : 0;
    sw += __shfl_xor_sync(FULL, sw, 8); sw += __shfl_xor_sync(FULL, sw, 16);
    if (lane == 0 && sw > 0 && a.sweeps) atomicAdd(a.sweeps, (unsigned long long)sw);
}

// stand-alone batched solver
struct BatchArgs {
    const double* XtX; int shared; const double* Xty; const double* w0; double* beta; int* sweeps;
    int K, KP; int64_t n; double lambda, alpha, tol; int perm_mode; uint64_t seed; uint32_t als_iter; uint64_t gene0;
};
template <int SL>
__global__ void __launch_bounds__(CD_WARPS * 32) k_cd_batch(BatchArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int KP = SL * 8;
    constexpr int XLD = KP + 1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, grp = lane >> 3, li = lane & 7;
    double* Xall = reinterpret_cast<double*>(smem_raw);
    double* sh_all = Xall + (size_t)CD_WARPS * GPW * KP * XLD;
    unsigned char* ord_all = reinterpret_cast<unsigned char*>(sh_all + (size_t)CD_WARPS * GPW * 5 * KP);
    const int64_t j0 = ((int64_t)blockIdx.x * CD_WARPS + warp) * GPW;
    double* Xw = Xall + (size_t)warp * GPW * KP * XLD;
    for (int gi = 0; gi < GPW; ++gi) {
        const int64_t j = j0 + gi;
        if (j >= a.n) break;
        const double* src = a.XtX + (a.shared ? 0 : (size_t)j * a.K * a.K);
        for (int x = lane; x < KP * KP; x += 32) {
            const int r = x / KP, c = x % KP;
            Xw[(size_t)gi * KP * XLD + r * XLD + c] = (r < a.K && c < a.K) ? src[r + (size_t)c * a.K] : 0.0;
        }
    }
    __syncwarp();
    const int64_t j = j0 + grp;
    const bool gvalid = j < a.n;
    double xty[SL], beta[SL];
#pragma unroll
    for (int s = 0; s < SL; ++s) {
        const int c = s * LPG + li;
        xty[s] = (gvalid && c < a.K) ? a.Xty[j * a.K + c] : 0.0;
        beta[s] = (gvalid && c < a.K) ? a.w0[j * a.K + c] : 0.0;
    }
    int sweeps = 0;
    group_cd<SL>(Xw + (size_t)grp * KP * XLD, XLD, a.K, li, gvalid, xty, beta, a.lambda, a.alpha, a.tol, a.perm_mode, a.seed, a.als_iter,
                 a.gene0 + (uint64_t)j, sh_all + (size_t)(warp * GPW + grp) * 5 * KP, ord_all + (warp * GPW + grp) * 32, sweeps);
    if (gvalid) {
#pragma unroll
        for (int s = 0; s < SL; ++s) { const int c = s * LPG + li; if (c < a.K) a.beta[j * a.K + c] = beta[s]; }
        if (li == 0 && a.sweeps) a.sweeps[j] = sweeps;
    }
}

template <typename KernelT>
void opt_in_smem(KernelT k, size_t bytes) {
    if (bytes > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

}  // namespace

void launch_col_solve(const Geom& g, bool masked, const uint32_t* trC, const double* U, const double* UtU, const double* Xty, double* V,
                      const CdParams& p, unsigned long long* sweeps, int* err_flag, cudaStream_t st) {
    SolveArgs a{};
    a.trC = trC; a.U = U; a.UtU = UtU; a.Xty = Xty; a.V = V;
    a.N = g.N; a.K = g.K; a.KP = g.KP; a.ldV = g.ldV; a.Wp = g.Wp; a.P = g.P; a.gene0 = g.gene0;
    a.lambda = p.lambda; a.alpha = p.alpha; a.tol = p.tol; a.als_iter = p.als_iter; a.seed = p.seed; a.perm_mode = p.perm_mode;
    a.sweeps = sweeps; a.err_flag = err_flag;
    const int genes_per_block = CD_WARPS * GPW;
    const int blocks = (int)((g.P + genes_per_block - 1) / genes_per_block);
    const int XLD = g.KP + 1;
    const size_t mats = masked ? (size_t)CD_WARPS * GPW : 1;
    const size_t smem = mats * g.KP * XLD * 8 + (size_t)CD_WARPS * GPW * 5 * g.KP * 8 + CD_WARPS * GPW * 32;
#define LAUNCH_CS(SLv)                                                                                                 \
    if (masked) { opt_in_smem(k_col_solve<SLv, true>, smem); k_col_solve<SLv, true><<<blocks, CD_WARPS * 32, smem, st>>>(a); } \
    else { opt_in_smem(k_col_solve<SLv, false>, smem); k_col_solve<SLv, false><<<blocks, CD_WARPS * 32, smem, st>>>(a); }
    switch (g.NT) { case 1: LAUNCH_CS(1) break; case 2: LAUNCH_CS(2) break; case 3: LAUNCH_CS(3) break; default: LAUNCH_CS(4) break; }
#undef LAUNCH_CS
}

void launch_cd_batch(int K, int64_t n, const double* XtX, bool shared, const double* Xty, const double* w0, double lambda, double alpha,
                     double tol, int perm_mode, uint64_t seed, uint32_t als_iter, uint64_t gene0, double* beta, int* sweeps,
                     cudaStream_t st) {
    BatchArgs a{};
    a.XtX = XtX; a.shared = shared ? 1 : 0; a.Xty = Xty; a.w0 = w0; a.beta = beta; a.sweeps = sweeps;
    a.K = K; a.KP = round_up(K, 8); a.n = n; a.lambda = lambda; a.alpha = alpha; a.tol = tol; a.perm_mode = perm_mode; a.seed = seed;
    a.als_iter = als_iter; a.gene0 = gene0;
    const int genes_per_block = CD_WARPS * GPW;
    const int blocks = (int)((n + genes_per_block - 1) / genes_per_block);
    const size_t smem = (size_t)CD_WARPS * GPW * a.KP * (a.KP + 1) * 8 + (size_t)CD_WARPS * GPW * 5 * a.KP * 8 + CD_WARPS * GPW * 32;
#define LAUNCH_CB(SLv) { opt_in_smem(k_cd_batch<SLv>, smem); k_cd_batch<SLv><<<blocks, CD_WARPS * 32, smem, st>>>(a); }
    switch (a.KP / 8) { case 1: LAUNCH_CB(1) break; case 2: LAUNCH_CB(2) break; case 3: LAUNCH_CB(3) break; default: LAUNCH_CB(4) break; }
#undef LAUNCH_CB
}

}  // namespace ib
