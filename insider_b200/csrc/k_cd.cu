// Per-gene column update: masked Gram build (DMMA) + elastic-net coordinate descent / ridge solve.
//
//   replaces  optimize_col()               src/optimize.cpp:200-253
//             strong_coordinate_descent()  src/coordinate_descent.cpp:57-127
//             compute_loss(vec, vec, ..)   src/utils.cpp:46-49   (only loss *differences* are needed, see below)
//
// Arithmetic form. The reference keeps an explicit residual r = y - X beta (n ~ 0.9 N entries) and evaluates r.x_k per
// coordinate. Here the same iteration runs in covariance form on q = X'y - X'X beta (q_k == r.x_k):
//     upper_k = q_k + beta_k XtX_kk ; soft-threshold ; q -= (new - old) XtX[:,k]
// which needs K instead of n flops per coordinate and no access to X or y. The stopping rule |pre_loss - loss| <= tol
// (coordinate_descent.cpp:114) is evaluated from the exact per-coordinate loss decrement
//     dL_k = (new - old) ((XtX_kk + l2)(new + old)/2 - upper_k) + lambda alpha (|new| - |old|)
// summed over the sweep: the same quantity without the cancellation of subtracting two O(|r|^2) numbers. The KKT
// re-admission test (coordinate_descent.cpp:118-119) is |q_e| > alpha lambda for excluded e because beta_e = 0 there.
// State form: p_k = q_k + beta_k XtX_kk (the "upper" of coordinate_descent.cpp:94) is what the lanes hold - see k_cd_dense.cu.
// Visit order: counter-based permutation identical to the oracle's mode B.
//
// Mapping (k_cd_persistent): LPG lanes per gene (8, or 4 in the long solves of the first ALS iterations), 32 / LPG genes per warp;
// lane li of a group holds p of coordinates li, li + LPG, .. in registers for the whole solve (coordinate layout; no per-sweep
// relayout). Groups pull genes from an atomic queue (sweep counts vary 4x between genes) and run independently: each has its own
// sweep index and therefore its own visiting order. A step for coordinate k: every lane computes the same scalar update from the
// group's shared beta / diagonal arrays and the upper of k - which arrived one step EARLIER (look-ahead, see the kernel) - then
// updates its own p with its elements of row k of the gene's Gram matrix (LPG consecutive doubles per group, the groups of a warp
// interleaved per 128-byte line: conflict-free). History of the step: profiles/r02_masked_solver_versions.txt.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "kernels.cuh"

namespace ib {

namespace {

constexpr int CD_WARPS = 2;       // warps per block (8 genes, 43 KB of per-gene tables at K = 23: 5 blocks = 40 genes per SM)
constexpr int MAX_SWEEPS = 200000;

__device__ __forceinline__ double grp_sum(double v) {
    v += __shfl_xor_sync(FULL, v, 4); v += __shfl_xor_sync(FULL, v, 2); v += __shfl_xor_sync(FULL, v, 1);
    return v;
}
// c ? a : b as one selp (kept opaque: nvcc otherwise turns a select chain over q[] into a dynamically indexed local array)
__device__ __forceinline__ double selp64(double a, double b, int c) {
    double r;
    asm("{\n .reg .pred p;\n setp.ne.s32 p, %3, 0;\n selp.f64 %0, %1, %2, p;\n}\n" : "=d"(r) : "d"(a), "d"(b), "r"(c));
    return r;
}
__device__ __forceinline__ double lds64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(addr));
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
// k_col_gram: XtX_j = UtU - sum_{i: m_ij = 0} u_i u_i^T  for every gene (src/optimize.cpp:216-219), one warp per gene,
// DMMA rank-4 updates over gathered rows of U. Output [P][KP*KP] row-major (both triangles).
template <int SL>
__global__ void __launch_bounds__(256) k_col_gram(const uint32_t* __restrict__ trC, const double* __restrict__ U, const double* __restrict__ UtU,
                                                  double* __restrict__ out, int N, int KP, int Wp, int64_t P) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int64_t gene = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (gene >= P) return;
    double acc[SL][SL][2];
#pragma unroll
    for (int i = 0; i < SL; ++i)
#pragma unroll
        for (int j = 0; j < SL; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const int nW = (N + 31) >> 5;
    int rows[4] = {0, 0, 0, 0};
    int cnt = 0;
    auto flush = [&]() {
        const int myrow = (t == 0) ? rows[0] : (t == 1) ? rows[1] : (t == 2) ? rows[2] : rows[3];
        double f[SL];
#pragma unroll
        for (int n = 0; n < SL; ++n) f[n] = (t < cnt) ? __ldg(U + (size_t)myrow * KP + 8 * n + g) : 0.0;
#pragma unroll
        for (int n1 = 0; n1 < SL; ++n1)
#pragma unroll
            for (int n2 = n1; n2 < SL; ++n2) dmma(acc[n1][n2][0], acc[n1][n2][1], f[n1], f[n2]);
        cnt = 0;
    };
    for (int w0 = 0; w0 < nW; w0 += 32) {
        uint32_t z = 0;
        const int wi = w0 + lane;
        if (wi < nW) {
            z = ~__ldg(trC + gene * Wp + wi);
            const int lim = N - 32 * wi;
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
    double* o = out + (size_t)gene * KP * KP;
#pragma unroll
    for (int n1 = 0; n1 < SL; ++n1)
#pragma unroll
        for (int n2 = n1; n2 < SL; ++n2)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int ra = 8 * n1 + g, cb = 8 * n2 + 2 * t + e;
                const double v = UtU[ra * KP + cb] - acc[n1][n2][e];
                o[ra * KP + cb] = v;
                o[cb * KP + ra] = v;
            }
}

// ---------------------------------------------------------------------------------------------------------------
struct CdArgs {
    const double* Xsh;            // shared K x K matrix (dense path), element (r, c) at r*xs_r + c*xs_c
    const double* Xall;           // per-gene matrices (masked path / batch), gene j at j*x_stride, same strides
    int64_t x_stride; int xs_r, xs_c;
    const double* Xty; const double* W0; double* Vout;     // per gene, stride ldv (Vout may alias W0)
    int64_t ldv;
    int K; int64_t P; int64_t gene0;
    double lambda, alpha;
    const double* tol_dev; double tol_host;
    const uint32_t* als_iter_dev; uint32_t als_iter_host;
    uint64_t seed; int perm_mode;
    unsigned long long* sweeps_total; unsigned long long* steps_total; int* sweeps_per_gene;
    unsigned int* queue;          // atomic gene counter (zeroed before launch)
    const int* order;             // optional [P]: the queue hands out gene order[i] (longest expected solves first)
    const unsigned char* perm_table;   // rank + order tables (common.cuh)
};

// persistent elastic-net solver: every LPG-lane group repeatedly claims a gene, solves it, writes it back.
//
// Step (round 2, "look-ahead"): the upper of the NEXT coordinate is fetched from its owner lane before this step's update is
// known (the shuffle overlaps the soft-threshold chain) and brought up to date by every lane with one FMA on X[k][k_next] -
// bitwise the FMA its owner applies. The chain between consecutive steps is |p| - la -> copysign -> * 1/den -> two selects ->
// new - old -> FMA: no shuffle, no select, no shared-memory load on it (the operands of step i + 1 are loaded during step i;
// the order row sits in registers and the sweep is fully unrolled, so operand addresses never wait for a load either).
// Measured (profiles/r02_masked_solver_versions.txt): first 13 masked iterations 237 -> 202 ms, lone-warp sweep 2.45 -> 1.59 us.
//
// Matrix layout in shared memory (round 2): row r is SL lines of 16 doubles; a line holds LPG consecutive columns of 16 / LPG
// genes (neighbouring groups of a warp) side by side: element (r, c) of gene slot h at (r*SL + c/LPG)*16 + h*LPG + c%LPG. The
// groups of a warp read rows of different matrices at every step; the groups served in one half-warp phase of an LDS.64 sit in
// disjoint bank ranges whatever their rows are (rows of pitch KP + 1 started at arbitrary banks: 30 % of the wavefronts were
// conflicts, profiles/r02_ncu_k_cd_persistent_la_masked.txt).
template <int LPGv> struct GroupGeom {
    static constexpr int LPG = LPGv, GPW = 32 / LPGv, LOG = (LPGv == 8) ? 3 : 2, GPL = 16 / LPGv;
    static constexpr int WARPS = (LPGv == 8) ? CD_WARPS : 1;      // 8 genes (41.7 KB of tables at K = 23) per block either way
    static constexpr int OW = 8 / LPGv;                           // 32-bit words of the 32-byte order row per lane
};

template <int KP, int LPGv, bool PERGENE>
__global__ void __launch_bounds__(GroupGeom<LPGv>::WARPS * 32) k_cd_persistent(CdArgs a) {
    using GG = GroupGeom<LPGv>;
    constexpr int LPG = GG::LPG, GPW = GG::GPW, LOG = GG::LOG, GPL = GG::GPL, NWARP = GG::WARPS, OW = GG::OW;
    constexpr int SL = KP / LPG;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // (Measured and dropped, profiles/r02_masked_solver_versions.txt v14: keeping only the lower triangle of every matrix - 2.4 instead of
    // 4.6 KB per gene, 9 instead of 5 one-warp blocks per SM, both access directions bank-conflict free with element t = r(r+1)/2 + c of
    // gene slot h at 4t + h - costs a compare, two selects and an add per load: 92 instead of 62 instructions per step, iteration 0
    // 101 instead of 83 ms.)
    constexpr int MAT = KP * SL * 16;                                          // doubles per set of GPL interleaved matrices
    auto XI = [](int r, int c) -> int { return (r * SL + (c >> LOG)) * 16 + (c & (LPG - 1)); };
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, grp = lane >> LOG, li = lane & (LPG - 1);
    const int grp_shift = grp << LOG;
    constexpr uint32_t lmask = (1u << LPG) - 1u;
    const uint32_t gmask = lmask << grp_shift;
    double* Xall_s = reinterpret_cast<double*>(smem_raw);
    constexpr int n_mats = PERGENE ? NWARP * GPW : 1;
    double* sh_all = Xall_s + (size_t)((n_mats + GPL - 1) / GPL) * MAT;        // [NWARP*GPW][3*KP]
    unsigned char* ord_all = reinterpret_cast<unsigned char*>(sh_all + (size_t)NWARP * GPW * 3 * KP);
    const int gslot = warp * GPW + grp;
    double* Xs = PERGENE ? Xall_s + (size_t)(gslot / GPL) * MAT + (gslot % GPL) * LPG : Xall_s;
    // row k, this lane's slot s: one shared-memory byte address = lane constant + k * row pitch
    const uint32_t xs_base = smem_u32(Xs);
    uint32_t cA[SL];
#pragma unroll
    for (int s = 0; s < SL; ++s) cA[s] = xs_base + (uint32_t)(s * 16 + li) * 8u;
    // (measured, masked iteration 0: the explicit-address form wins with 8 lanes per gene, 91.8 -> 87.3 ms, and loses with 4, 83.1 -> 91.9 ms)
    auto XR = [&](int k, int s) -> double {
        if (LPG == 8) return lds64(cA[s] + (uint32_t)k * (uint32_t)(SL * 16 * 8));
        return Xs[(k * SL + s) * 16 + li];
    };
    double* sh = sh_all + (size_t)gslot * 3 * KP;
    unsigned char* ord_s = ord_all + gslot * 32;
    double* Bc = sh; double* DRc = sh + KP;            // beta [KP] | ((XtX_kk + l2)/2, 1/(XtX_kk + l2)) pairs [2*KP]   (16-byte aligned: KP % 8 == 0)
    const int K = a.K;
    const double tol = a.tol_dev ? *a.tol_dev : a.tol_host;
    const uint32_t als_iter = a.als_iter_dev ? *a.als_iter_dev : a.als_iter_host;
    const double la = a.lambda * a.alpha, l2 = a.lambda * (1.0 - a.alpha);
    const uint64_t key_iter = mix64(a.seed + 0x9E3779B97F4A7C15ull * (1ull + als_iter));   // perm_key(): first factor

    if (!PERGENE) {
        for (int x = tid; x < KP * KP; x += blockDim.x) {
            const int r = x / KP, c = x % KP;
            Xall_s[XI(r, c)] = (r < K && c < K && r != c) ? a.Xsh[(size_t)r * a.xs_r + (size_t)c * a.xs_c] : 0.0;   // zero diagonal (p form)
        }
        __syncthreads();
    }

    // per-group solver state
    int64_t gene = 0;
    bool active = false, retired = false;
    uint32_t inc = 0, draw = 0;
    int n_inc = 0, sweeps = 0;
    uint32_t row_w[OW], row_draw = 0xffffffffu;                // prefetched order-table words (see below)
#pragma unroll
    for (int w = 0; w < OW; ++w) row_w[w] = 0;
    unsigned long long sweeps_acc = 0, steps_acc = 0;
    double q[SL];                                              // p = X'y - (X'X - diag) beta of coordinates s*LPG + li
#pragma unroll
    for (int s = 0; s < SL; ++s) q[s] = 0.0;

    while (true) {
        // ---- claim and set up a new gene (divergent per group; group-local masks only)
        if (!active && !retired) {
            unsigned int jn = 0;
            if (li == 0) jn = atomicAdd(a.queue, 1u);
            jn = __shfl_sync(gmask, jn, 0, LPG);
            if ((int64_t)jn >= a.P) {
                retired = true;
            } else {
                gene = a.order ? (int64_t)a.order[jn] : (int64_t)jn;
                if (PERGENE) {
                    // per-gene matrix global -> shared, eight independent loads in flight per lane (a plain strided loop
                    // pays one L2 round trip per element)
                    const double* src = a.Xall + (size_t)gene * a.x_stride;
                    for (int x0 = li; x0 < KP * KP; x0 += 8 * LPG) {
                        double v[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int x = x0 + u * LPG, r = x / KP, c = x % KP;
                            v[u] = (x < KP * KP && r < K && c < K) ? src[(size_t)r * a.xs_r + (size_t)c * a.xs_c] : 0.0;
                        }
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int x = x0 + u * LPG;
                            if (x < KP * KP) Xs[XI(x / KP, x % KP)] = v[u];
                        }
                    }
                    __syncwarp(gmask);
                }
                double xty[SL], beta[SL];
                double mx = 0.0;
#pragma unroll
                for (int s = 0; s < SL; ++s) {
                    const int c = s * LPG + li;
                    xty[s] = (c < K) ? a.Xty[gene * a.ldv + c] : 0.0;
                    beta[s] = (c < K) ? a.W0[gene * a.ldv + c] : 0.0;
                    mx = fmax(mx, fabs(xty[s]));
                }
#pragma unroll
                for (int o = LPG / 2; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(gmask, mx, o));
                const double thr = a.alpha * (2.0 * a.lambda - mx);                // coordinate_descent.cpp:74
                inc = 0;
#pragma unroll
                for (int s = 0; s < SL; ++s) {
                    const int c = s * LPG + li;
                    const bool on = (c < K) && !(fabs(xty[s]) < thr);
                    if (!on) beta[s] = 0.0;                                        // :75-78
                    const uint32_t bal = __ballot_sync(gmask, on);
                    inc |= ((bal >> grp_shift) & lmask) << (LPG * s);
                    q[s] = xty[s];
                }
                // group-shared scalars of every coordinate, then the diagonal is zeroed in the shared copy: the solver keeps
                // p = q + beta * diag (k_cd_dense.cu, "state form"), whose own entry does not move when its coordinate is updated
#pragma unroll
                for (int s = 0; s < SL; ++s) {
                    const int c = s * LPG + li;
                    const double d = PERGENE ? Xs[XI(c, c)] : ((c < K) ? a.Xsh[(size_t)c * a.xs_r + (size_t)c * a.xs_c] : 0.0);
                    // an EXCLUDED coordinate gets 1/den = 0: its step computes new = +-0 = old, delta = +-0, and every FMA of the step is
                    // an exact no-op - no test of the active set inside the step (5 instructions of ~60)
                    Bc[c] = beta[s]; DRc[2 * c] = 0.5 * (d + l2); DRc[2 * c + 1] = ((inc >> c) & 1u) ? 1.0 / (d + l2) : 0.0;
                }
                __syncwarp(gmask);
#pragma unroll
                for (int s = 0; s < SL; ++s) if (PERGENE) Xs[XI(s * LPG + li, s * LPG + li)] = 0.0;
                __syncwarp(gmask);
                // p = X'y - (X'X - diag) beta   (:79, in covariance form)
#pragma unroll
                for (int ms = 0; ms < SL; ++ms)
                    for (int ml = 0; ml < LPG; ++ml) {
                        const int m = ms * LPG + ml;
                        const double bm = __shfl_sync(gmask, beta[ms], ml, LPG);
                        if (m < K && bm != 0.0) {
#pragma unroll
                            for (int s = 0; s < SL; ++s) q[s] = fma(-XR(m, s), bm, q[s]);
                        }
                    }
                __syncwarp(gmask);
                n_inc = __popc(inc); draw = 0; sweeps = 0; active = true; row_draw = 0xffffffffu;
            }
        }
        if (__all_sync(FULL, retired)) break;

        // ---- visiting order of this group's sweep (coordinate_descent.cpp:89): the per-sweep key selects a table permutation of
        //      all K coordinates (shared by every gene at this sweep index); active coordinates are visited in that order. Lane
        //      li holds the 4-byte words li, li + LPG, .. of the 32-byte order row; the row of the NEXT sweep is prefetched one
        //      sweep ahead.
        {
            auto row_word = [&](uint32_t dr, int w) -> uint32_t {
                const uint32_t wi = (uint32_t)(li + w * LPG);
                if (a.perm_mode != 1) { const uint32_t c0 = 4u * wi; return c0 | ((c0 + 1) << 8) | ((c0 + 2) << 16) | ((c0 + 3) << 24); }
                const uint64_t pk = key_iter ^ mix64((uint64_t)dr * 0x8CB92BA72F3D8DD7ull + 0x2545F4914F6CDD1Dull);
                return __ldg(reinterpret_cast<const uint32_t*>(a.perm_table + PERM_TABLE_HALF + ((size_t)(K - 1) * PERM_T + perm_select(pk)) * 32) + wi);
            };
#pragma unroll
            for (int w = 0; w < OW; ++w) {
                if (row_draw != draw) row_w[w] = row_word(draw, w);                       // new gene
                reinterpret_cast<uint32_t*>(ord_s)[li + w * LPG] = row_w[w];
                row_w[w] = row_word(draw + 1, w);                                         // consumed by the next sweep
            }
            row_draw = draw + 1;
        }
        ++draw;
        __syncwarp();
        // ---- one sweep: step i visits coordinate k = ord[i] of every group (groups are at different sweeps: k differs per group)
        // (a retired group keeps stepping on its last gene's state: a converged solve, never written back)
        double dl = 0.0;
        {
            uint32_t ow[KP / 4];
#pragma unroll
            for (int w = 0; w < KP / 4; ++w) ow[w] = reinterpret_cast<const uint32_t*>(ord_s)[w];
            auto ob = [&](int i) -> int { return (int)((ow[i >> 2] >> (8 * (i & 3))) & 0xffu); };
            auto sel_slot = [&](int kk) -> double {                                       // q of coordinate kk in its owner lane
                double v = q[0];
#pragma unroll
                for (int s = 1; s < SL; ++s) v = selp64(q[s], v, (kk >> LOG) == s);
                return v;
            };
            int k = ob(0), kn = (1 < K) ? ob(1) : ob(0);                                      // (K == KP - 7 == 1 is possible)
            double xr[SL];
#pragma unroll
            for (int s = 0; s < SL; ++s) xr[s] = XR(k, s);
            double2 dr = *reinterpret_cast<const double2*>(DRc + 2 * k);                  // (XtX_kk + l2) / 2, 1 / (XtX_kk + l2)
            double bo = Bc[k], xkn = Xs[XI(k, kn)];
            double up = __shfl_sync(FULL, sel_slot(k), k & (LPG - 1), LPG);               // :94 - the state is the upper itself (p form)
#pragma unroll
            for (int i = 0; i < KP; ++i) {
                if (i >= KP - 7 && i >= K) break;                                             // KP - 7 <= K <= KP: the first KP - 7 steps always run
                const int knn = (i + 2 < KP && (i + 2 < KP - 7 || i + 2 < K)) ? ob((i + 2 < KP) ? i + 2 : KP - 1) : kn;
                // operands of step i + 1 (addresses depend on the order only) and its upper as of BEFORE this step's update
                double xrn[SL];
#pragma unroll
                for (int s = 0; s < SL; ++s) xrn[s] = XR(kn, s);
                const double2 drn = *reinterpret_cast<const double2*>(DRc + 2 * kn);
                const double xknn = Xs[XI(kn, knn)];
                const double shn = __shfl_sync(FULL, sel_slot(kn), kn & (LPG - 1), LPG);
                const double bon = Bc[kn];
                // the step itself (coordinate_descent.cpp:94-109)
                const double t1 = fabs(up) - la;
                double nb = copysign(t1, up) * dr.y;                                          // :99-104 (excluded coordinate: dr.y = 0)
                nb = (__double2hiint(t1) >= 0) ? nb : 0.0;
                const double dlt = nb - bo;
                const double nd = -dlt;
                const double upn = fma(nd, xkn, shn);                                         // what the owner of k_next computes below
#pragma unroll
                for (int s = 0; s < SL; ++s) q[s] = fma(nd, xr[s], q[s]);
                if (li == (k & (LPG - 1))) Bc[k] = nb;                                        // :106-109
                // exact loss decrement: dlt ((XtX_kk + l2)(new + old)/2 - upper) + lambda alpha (|new| - |old|)
                dl = fma(dlt, fma(dr.x, nb + bo, -up), dl);
                dl = fma(la, fabs(nb) - fabs(bo), dl);
                k = kn; kn = knn; dr = drn; bo = bon; xkn = xknn; up = upn;
#pragma unroll
                for (int s = 0; s < SL; ++s) xr[s] = xrn[s];
            }
        }
        __syncwarp();
        // inner do-while ends (:114) -> KKT check on the excluded set (:118-124); every lane of the group holds the same dl
        const bool inner_end = active && (!(fabs(dl) > tol) || sweeps + 1 >= MAX_SWEEPS);
        uint32_t vmask = 0;
        if (inner_end) {
#pragma unroll
            for (int s = 0; s < SL; ++s) {
                const int c = s * LPG + li;
                if (c < K && !((inc >> c) & 1u) && fabs(q[s]) > la) vmask |= 1u << c;     // |XtX[e,inc] beta - Xty_e| = |q_e| (beta_e = 0)
            }
        }
        if (__any_sync(FULL, inner_end)) {                                                // (most sweeps end nowhere: one vote instead of the shuffles)
#pragma unroll
            for (int o = LPG / 2; o > 0; o >>= 1) vmask |= __shfl_xor_sync(FULL, vmask, o);
        }
        if (active) {
            ++sweeps;
            steps_acc += (unsigned long long)n_inc;     // coordinate updates attempted (statistics only)
            if (inner_end) {
                if (vmask == 0u || sweeps >= MAX_SWEEPS) {
                    // finished: write the gene back
#pragma unroll
                    for (int s = 0; s < SL; ++s) { const int c = s * LPG + li; if (c < K) a.Vout[gene * a.ldv + c] = Bc[c] + 0.0; }   // (+ 0.0: an excluded coordinate may hold -0)
                    if (li == 0) { sweeps_acc += (unsigned long long)sweeps; if (a.sweeps_per_gene) a.sweeps_per_gene[gene] = sweeps; }
                    active = false;
                } else {
                    // re-admitted coordinates get their 1/den back: 1 / (2 * ((XtX_kk + l2) / 2)), bitwise the set-up's quotient
#pragma unroll
                    for (int s = 0; s < SL; ++s) { const int c = s * LPG + li; if ((vmask >> c) & 1u) DRc[2 * c + 1] = 1.0 / (2.0 * DRc[2 * c]); }
                    inc |= vmask; n_inc = __popc(inc);
                }
            }
        }
    }
    // one atomic per warp for the sweep statistics
    if (li != 0) steps_acc = 0;
#pragma unroll
    for (int o = LPG; o < 32; o <<= 1) { sweeps_acc += __shfl_xor_sync(FULL, sweeps_acc, o); steps_acc += __shfl_xor_sync(FULL, steps_acc, o); }
    if (lane == 0 && sweeps_acc && a.sweeps_total) atomicAdd(a.sweeps_total, sweeps_acc);
    if (lane == 0 && steps_acc && a.steps_total) atomicAdd(a.steps_total, steps_acc);
}

// ---------------------------------------------------------------------------------------------------------------
// ridge (alpha == 0): (XtX + lambda I) v = Xty   — optimize.cpp:224-226 (masked) / :237-240 (dense); one warp per gene
struct RidgeArgs {
    const double* UtU; const double* Xall; const double* Xty; double* V;
    int K, KP, ldV; int64_t P; double lambda; int* err_flag;
};
template <bool MASKED>
__global__ void __launch_bounds__(128) k_col_ridge(RidgeArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int KP = a.KP, XLD = KP + 1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double* S0 = reinterpret_cast<double*>(smem_raw);
    bool ok = true;
    if (!MASKED) {
        for (int x = tid; x < KP * KP; x += blockDim.x) {
            const int r = x / KP, c = x % KP;
            S0[r * XLD + c] = a.UtU[x] + ((r == c) ? a.lambda : 0.0);               // optimize.cpp:238
        }
        __syncthreads();
        if (warp == 0) ok = warp_chol_factor(S0, XLD, a.K, lane);
        __syncthreads();
    }
    const int64_t j = (int64_t)blockIdx.x * 4 + warp;
    if (j < a.P) {
        double b = (lane < a.K) ? a.Xty[j * a.ldV + lane] : 0.0;
        if (MASKED) {
            double* S = S0 + (size_t)warp * KP * XLD;
            const double* src = a.Xall + (size_t)j * KP * KP;
            for (int x0 = lane; x0 < KP * KP; x0 += 8 * 32) {
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) { const int x = x0 + u * 32; v[u] = (x < KP * KP) ? src[x] : 0.0; }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int x = x0 + u * 32, r = x / KP, c = x % KP;
                    if (x < KP * KP) S[r * XLD + c] = v[u] + ((r == c) ? a.lambda : 0.0);   // :225
                }
            }
            __syncwarp();
            ok = warp_chol_factor(S, XLD, a.K, lane);
            b = warp_chol_subst(S, XLD, a.K, lane, b);
        } else {
            b = warp_chol_subst(S0, XLD, a.K, lane, b);
        }
        if (lane < a.K) a.V[j * a.ldV + lane] = b;
    }
    if (!ok && lane == 0) atomicExch(a.err_flag, 1);
}

template <typename KernelT>
void opt_in_smem(KernelT k, size_t bytes) {
    if (bytes > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

template <int LPGv>
size_t cd_smem_bytes(int KP, bool pergene) {
    using GG = GroupGeom<LPGv>;
    const size_t genes = (size_t)GG::WARPS * GG::GPW;
    const size_t sets = pergene ? genes / GG::GPL : 1;                         // 16 / LPG matrices interleaved per 128-byte line
    return sets * KP * (KP / LPGv) * 16 * 8 + genes * 3 * KP * 8 + genes * 32;
}

template <int KP, int LPGv, bool PG>
void launch_cd_kp(const CdArgs& a, int sm_count, cudaStream_t st) {
    using GG = GroupGeom<LPGv>;
    const size_t smem = cd_smem_bytes<LPGv>(KP, PG);
    const int64_t genes_per_block = GG::WARPS * GG::GPW;
    int64_t blocks = (a.P + genes_per_block - 1) / genes_per_block;
    int per_sm = (int)std::min<size_t>(16, (227 * 1024) / (smem + 1024));      // persistent: enough blocks to fill every SM
    if (per_sm < 1) per_sm = 1;
    blocks = std::min<int64_t>(blocks, (int64_t)sm_count * per_sm);
    opt_in_smem(k_cd_persistent<KP, LPGv, PG>, smem);
    k_cd_persistent<KP, LPGv, PG><<<(int)blocks, GG::WARPS * 32, smem, st>>>(a);
}

// long_solves: hundreds of sweeps per gene expected (first ALS iterations). 4 lanes per gene issue 7.8 instead of 13.3 warp
// instructions per gene-step (8 genes per warp share the scalar chain) and win there (iteration 0: 92 -> 83 ms); 8 lanes per gene
// set a gene up twice as fast and have the shorter lone-warp sweep (1.6 against 1.95 us): better from ~100 sweeps per gene down.
void launch_cd(const CdArgs& a, int KP, bool pergene, int sm_count, bool long_solves, cudaStream_t st) {
    cudaMemsetAsync(a.queue, 0, sizeof(unsigned int), st);
    static const int lpg_env = [] { const char* e = getenv("INSIDER_B200_CD_LPG"); return e ? atoi(e) : 0; }();   // A/B switch (profiles/r02_masked_solver_versions.txt)
    const int lpg = (lpg_env == 4 || lpg_env == 8) ? lpg_env : (long_solves ? 4 : 8);
#define LAUNCH_CD(KPv)                                                                                                    \
    { if (lpg == 8) { if (pergene) launch_cd_kp<KPv, 8, true>(a, sm_count, st); else launch_cd_kp<KPv, 8, false>(a, sm_count, st); }   \
      else { if (pergene) launch_cd_kp<KPv, 4, true>(a, sm_count, st); else launch_cd_kp<KPv, 4, false>(a, sm_count, st); } }
    switch (KP / 8) { case 1: LAUNCH_CD(8) break; case 2: LAUNCH_CD(16) break; case 3: LAUNCH_CD(24) break; default: LAUNCH_CD(32) break; }
#undef LAUNCH_CD
}

}  // namespace

void launch_col_gram(const Geom& g, const uint32_t* trC, const double* U, const double* UtU, double* XtXall, cudaStream_t st) {
    const int blocks = (int)((g.P + 7) / 8);
    if (blocks == 0) return;
    switch (g.NT) {
        case 1: k_col_gram<1><<<blocks, 256, 0, st>>>(trC, U, UtU, XtXall, g.N, g.KP, g.Wp, g.P); break;
        case 2: k_col_gram<2><<<blocks, 256, 0, st>>>(trC, U, UtU, XtXall, g.N, g.KP, g.Wp, g.P); break;
        case 3: k_col_gram<3><<<blocks, 256, 0, st>>>(trC, U, UtU, XtXall, g.N, g.KP, g.Wp, g.P); break;
        default: k_col_gram<4><<<blocks, 256, 0, st>>>(trC, U, UtU, XtXall, g.N, g.KP, g.Wp, g.P); break;
    }
}

void launch_col_solve(const Geom& g, bool masked, const double* UtU, const double* XtXall, const double* Xty, double* V, const CdParams& p,
                      unsigned long long* sweeps, unsigned long long* steps, unsigned int* queue, const unsigned char* perm_table, int* err_flag,
                      int sm_count, int* sweeps_per_gene, const int* order, bool long_solves, cudaStream_t st) {
    if (g.P == 0) return;
    if (p.alpha == 0.0) {
        RidgeArgs r{UtU, XtXall, Xty, V, g.K, g.KP, g.ldV, g.P, p.lambda, err_flag};
        const int blocks = (int)((g.P + 3) / 4);
        const size_t smem = (size_t)(masked ? 4 : 1) * g.KP * (g.KP + 1) * 8;
        if (masked) k_col_ridge<true><<<blocks, 128, smem, st>>>(r);
        else k_col_ridge<false><<<blocks, 128, smem, st>>>(r);
        return;
    }
    CdArgs a{};
    a.Xsh = UtU; a.Xall = XtXall; a.x_stride = (int64_t)g.KP * g.KP; a.xs_r = g.KP; a.xs_c = 1;
    a.Xty = Xty; a.W0 = V; a.Vout = V; a.ldv = g.ldV; a.K = g.K; a.P = g.P; a.gene0 = g.gene0;
    a.lambda = p.lambda; a.alpha = p.alpha; a.tol_dev = p.tol; a.als_iter_dev = p.als_iter; a.seed = p.seed; a.perm_mode = p.perm_mode;
    a.sweeps_total = sweeps; a.steps_total = steps; a.sweeps_per_gene = sweeps_per_gene; a.order = order; a.queue = queue; a.perm_table = perm_table;
    launch_cd(a, g.KP, masked, sm_count, long_solves, st);
}

void launch_cd_batch(int K, int64_t n, const double* XtX, bool shared, const double* Xty, const double* w0, double lambda, double alpha,
                     double tol, int perm_mode, uint64_t seed, uint32_t als_iter, uint64_t gene0, double* beta, int* sweeps,
                     unsigned int* queue, const unsigned char* perm_table, int sm_count, cudaStream_t st) {
    CdArgs a{};
    a.perm_table = perm_table;
    a.Xsh = XtX; a.Xall = XtX; a.x_stride = (int64_t)K * K; a.xs_r = 1; a.xs_c = K;      // caller's column-major K x K
    a.Xty = Xty; a.W0 = w0; a.Vout = beta; a.ldv = K; a.K = K; a.P = n; a.gene0 = (int64_t)gene0;
    a.lambda = lambda; a.alpha = alpha; a.tol_dev = nullptr; a.tol_host = tol; a.als_iter_dev = nullptr; a.als_iter_host = als_iter;
    a.seed = seed; a.perm_mode = perm_mode; a.sweeps_total = nullptr; a.steps_total = nullptr; a.sweeps_per_gene = sweeps; a.queue = queue;
    launch_cd(a, round_up(K, 8), !shared, sm_count, false, st);
}

}  // namespace ib
