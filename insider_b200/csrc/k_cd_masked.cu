// Masked (tuning = 1) column update with elastic net: per-gene Gram matrices that never exist as K x K arrays in HBM.
//
//   replaces  optimize_col() tuning = 1 branch   src/optimize.cpp:203-230   (XtX_j = U'U - sum_{i: m_ij = 0} u_i u_i')
//             strong_coordinate_descent()        src/coordinate_descent.cpp:57-127
//
// Round 1 wrote every gene's full K x K matrix to HBM (k_col_gram: 205 MB per iteration at 377 x 44 477, K = 23) and an
// 8-lanes-per-gene solver fetched it back with 72 scattered loads per lane, kept it in 4.8 KB of shared memory per gene (15 %
// occupancy, 29 % of its shared-memory wavefronts bank conflicts) and replicated the scalar part of every coordinate update in
// all 8 lanes (profiles/r01_ncu_k_cd_persistent_v6_masked_A.txt: 221 MB read for 18 MB of algorithmic input).
//
// Here:
//   k_col_gram_tiles  builds the matrices of 32 genes (one SLOT tile: genes in the order the solver will take them), one warp
//                     per gene at a time: the masked-out rows are listed first, then gathered 16 at a time (12 independent loads
//                     in flight per lane instead of one L2 round trip per 4 rows) into DMMA rank-4 updates. Only the lower
//                     triangle is kept (zero on the diagonal, the diagonal and 1 / (XtX_kk + l2) separately), staged gene-major
//                     and written in the solver's own shared-memory layout - element e of slot s at [e][s] - so that a tile is
//                     ONE contiguous block (89 KB at K = 23):  [ KT (KT+1)/2 triangle | KT diagonal | KT reciprocals ] x 32
//   k_cd_masked       one warp = one tile, ONE GENE PER THREAD like the dense solver (k_cd_dense.cu): the tile arrives by one
//                     TMA bulk copy; p = q + beta * diag (KT doubles) lives in registers in COORDINATE order, beta in a
//                     thread-private shared-memory column. A step on the warp-uniform coordinate k (the visiting order depends
//                     only on the sweep index, common.cuh) reads row k of its own gene's matrix - element (k, l) sits at
//                     tri(max, min): row part at a run-time base + compile-time offset, column part at compile-time base +
//                     run-time offset - so the sweep is straight-line code (a first version dispatched every step through
//                     `switch (k)` to 24 specialised 1 KB bodies: instruction-cache misses made it slower than round 1's kernel,
//                     profiles/r02_masked_solver_versions.txt). Only `up = p[k]` needs a run-time register index: a 24-way
//                     switch of single moves. Every lane reads its own gene's elements: [e][lane] is conflict-free.
// Arithmetic per coordinate: the p form of k_cd_dense.cu, operation for operation.
#include <cstdio>
#include <cstdlib>
#include <utility>

#include "common.cuh"
#include "kernels.cuh"

namespace ib {

namespace {

constexpr int MAX_SWEEPS_M = 200000;

__host__ __device__ constexpr int tri(int r, int c) { return r >= c ? r * (r + 1) / 2 + c : c * (c + 1) / 2 + r; }
__host__ __device__ constexpr int tile_elems(int KT) { return KT * (KT + 1) / 2 + 2 * KT; }   // triangle | diagonal | reciprocals

// ---------------------------------------------------------------------------------------------------------------
// k_col_gram_tiles: block = GT_WARPS warps = one tile of 32 slots, warp w builds slots w, w + GT_WARPS, ...
constexpr int GT_WARPS = 16;
constexpr int GT_GROUP = 1024;           // rows per mask-scan group (32 words): bounds the per-warp row list

template <int SL>
__global__ void __launch_bounds__(GT_WARPS * 32, 2) k_col_gram_tiles(const uint32_t* __restrict__ trC, const double* __restrict__ U, const double* __restrict__ UtU,
                                                                  const int* __restrict__ order, double* __restrict__ tiles, int N, int K, int KP, int KT,
                                                                  int Wp, int64_t P, double l2, int list_len) {
    extern __shared__ double st[];                                            // [32][E1] staging | per-warp row lists
    const int E = tile_elems(KT), E1 = E | 1, NTRI = KT * (KT + 1) / 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    unsigned short* list = reinterpret_cast<unsigned short*>(st + (size_t)32 * E1) + (size_t)warp * list_len;   // rows relative to the group
    const int nW = (N + 31) >> 5;
    for (int s = warp; s < 32; s += GT_WARPS) {
        const int64_t slot = (int64_t)blockIdx.x * 32 + s;
        double* out = st + (size_t)s * E1;
        if (slot >= P) {                                                      // padding slot: zero matrix, zero reciprocals
            for (int e = lane; e < E; e += 32) out[e] = 0.0;
            continue;
        }
        const int64_t gene = order ? (int64_t)order[slot] : slot;
        double acc[SL][SL][2];
#pragma unroll
        for (int i = 0; i < SL; ++i)
#pragma unroll
            for (int j = 0; j < SL; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        for (int w0 = 0; w0 < nW; w0 += 32) {
            // rows of this group whose train bit is 0, listed in ascending order (lane = mask word; exclusive prefix of popcounts)
            uint32_t z = 0;
            const int wi = w0 + lane;
            if (wi < nW) {
                z = ~__ldg(trC + gene * Wp + wi);
                const int lim = N - 32 * wi;
                if (lim < 32) z &= (1u << lim) - 1u;
            }
            int incl = __popc(z);
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
            const int total = __shfl_sync(FULL, incl, 31);
            int pos = incl - __popc(z);
            __syncwarp();
            while (z) { const int b = __ffs(z) - 1; z &= z - 1; list[pos++] = (unsigned short)(32 * lane + b); }
            __syncwarp();
            // 16 rows (4 rank-4 updates) per batch: all gathers of a batch are independent
            for (int b0 = 0; b0 < total; b0 += 16) {
                double f[4][SL];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int idx = b0 + 4 * q + t;
                    const int row = (idx < total) ? 32 * w0 + (int)list[idx] : -1;
#pragma unroll
                    for (int n = 0; n < SL; ++n) f[q][n] = (row >= 0) ? __ldg(U + (size_t)row * KP + 8 * n + g) : 0.0;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int n1 = 0; n1 < SL; ++n1)
#pragma unroll
                        for (int n2 = n1; n2 < SL; ++n2) dmma(acc[n1][n2][0], acc[n1][n2][1], f[q][n1], f[q][n2]);
            }
        }
#pragma unroll
        for (int n1 = 0; n1 < SL; ++n1)
#pragma unroll
            for (int n2 = n1; n2 < SL; ++n2)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int ra = 8 * n1 + g, cb = 8 * n2 + 2 * t + e;          // element (ra, cb) of the symmetric matrix
                    if (cb >= ra && cb < KT) {
                        const double v = (cb < K) ? UtU[ra * KP + cb] - acc[n1][n2][e] : 0.0;   // src/optimize.cpp:218
                        if (cb == ra) {
                            out[cb * (cb + 1) / 2 + ra] = 0.0;                // p form: a step never moves its own p_k
                            out[NTRI + ra] = v;
                            out[NTRI + KT + ra] = (ra < K) ? 1.0 / (v + l2) : 0.0;
                        } else out[cb * (cb + 1) / 2 + ra] = v;
                    }
                }
    }
    __syncthreads();
    double* dst = tiles + (size_t)blockIdx.x * E * 32;
    for (int x = threadIdx.x; x < E * 32; x += GT_WARPS * 32) dst[x] = st[(size_t)(x & 31) * E1 + (x >> 5)];
}

// ---------------------------------------------------------------------------------------------------------------
struct CdMaskedArgs {
    const double* tiles;             // [n_tiles][E][32]
    const double* Xty; double* V;    // per gene, stride ldv
    int64_t ldv;
    int K; int64_t P;
    double la, l2, alpha, lambda;
    const double* tol_dev; const uint32_t* als_iter_dev;
    uint64_t seed; int perm_mode;
    unsigned long long* sweeps_total; unsigned long long* steps_total;
    int* sweeps_per_gene;
    const int* order;
    const unsigned char* perm_table;
};

// p[k] for a run-time, warp-uniform k: 24-way switch of single moves
template <int KT>
__device__ __forceinline__ double select_reg(const double (&p)[KT], int k) {
    double v = p[0];
#define CDM_SEL(C) case C: if constexpr (C < KT) v = p[C < KT ? C : 0]; break;
    switch (k) {
        CDM_SEL(1) CDM_SEL(2) CDM_SEL(3) CDM_SEL(4) CDM_SEL(5) CDM_SEL(6) CDM_SEL(7)
        CDM_SEL(8) CDM_SEL(9) CDM_SEL(10) CDM_SEL(11) CDM_SEL(12) CDM_SEL(13) CDM_SEL(14) CDM_SEL(15)
        CDM_SEL(16) CDM_SEL(17) CDM_SEL(18) CDM_SEL(19) CDM_SEL(20) CDM_SEL(21) CDM_SEL(22) CDM_SEL(23)
        CDM_SEL(24) CDM_SEL(25) CDM_SEL(26) CDM_SEL(27) CDM_SEL(28) CDM_SEL(29) CDM_SEL(30) CDM_SEL(31)
        default: break;
    }
#undef CDM_SEL
    return v;
}

// row k (run time) of this thread's gene: x[l] = XtX[k][l] for l != k, 0 for l == k. Element (k, l) sits at tri(max, min) of the
// lower triangle; the byte offsets of a whole row come from a small shared table (off[k][l], built once per block) as 128-bit
// loads, so a row costs one integer add per element (selecting between "row base + l" and "column base_l + k" per element took
// 6 integer instructions each and made the step issue-bound: 230 instructions per step against 125 now).
template <int KT>
__device__ __forceinline__ void load_row(double (&x)[KT], const unsigned char* __restrict__ Tb, const uint32_t* __restrict__ off_s, int k) {
    const uint4* orow = reinterpret_cast<const uint4*>(off_s + k * KT);
#pragma unroll
    for (int q = 0; q < KT / 4; ++q) {
        const uint4 o = orow[q];
        x[4 * q + 0] = *reinterpret_cast<const double*>(Tb + o.x);
        x[4 * q + 1] = *reinterpret_cast<const double*>(Tb + o.y);
        x[4 * q + 2] = *reinterpret_cast<const double*>(Tb + o.z);
        x[4 * q + 3] = *reinterpret_cast<const double*>(Tb + o.w);
    }
}

template <int KT>
__global__ void __launch_bounds__(32, 2) k_cd_masked(CdMaskedArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int E = tile_elems(KT), NTRI = KT * (KT + 1) / 2;
    double* tile = reinterpret_cast<double*>(smem_raw);                       // [E][32]
    double* beta_s = tile + (size_t)E * 32;                                   // [KT][32] thread-private columns
    uint32_t* off_s = reinterpret_cast<uint32_t*>(beta_s + (size_t)KT * 32);   // [KT][KT] byte offset of element (k, l) in a thread's column
    unsigned char* ord_s = reinterpret_cast<unsigned char*>(off_s + KT * KT);  // [2][32] visiting order of this / the next sweep
    uint64_t* bar = reinterpret_cast<uint64_t*>(ord_s + 64);
    const int lane = threadIdx.x, K = a.K;
    const int64_t slot = (int64_t)blockIdx.x * 32 + lane;
    bool active = slot < a.P;
    const int64_t gene = active ? (a.order ? (int64_t)a.order[slot] : slot) : 0;
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        mbar_expect_tx(bar, (uint32_t)(E * 32 * 8));
        tma_load_1d(tile, a.tiles + (size_t)blockIdx.x * E * 32, (uint32_t)(E * 32 * 8), bar);
    }
    for (int x = lane; x < KT * KT; x += 32) off_s[x] = (uint32_t)tri(x / KT, x % KT) * 256u;
    const double la = a.la, hl2 = 0.5 * a.l2;
    const double tol = *a.tol_dev;
    const uint32_t als_iter = *a.als_iter_dev;
    const uint64_t key_iter = mix64(a.seed + 0x9E3779B97F4A7C15ull * (1ull + als_iter));   // perm_key(): first factor
    // order row of sweep dr: word (lane & 7) of the 32-byte row (identity when perm_mode != 1)
    auto row_word = [&](uint32_t dr) -> uint32_t {
        const uint32_t c0 = 4u * (lane & 7);
        if (a.perm_mode != 1) return c0 | ((c0 + 1) << 8) | ((c0 + 2) << 16) | ((c0 + 3) << 24);
        const uint64_t pk = key_iter ^ mix64((uint64_t)dr * 0x8CB92BA72F3D8DD7ull + 0x2545F4914F6CDD1Dull);
        return __ldg(reinterpret_cast<const uint32_t*>(a.perm_table + PERM_TABLE_HALF + ((size_t)(K - 1) * PERM_T + perm_select(pk)) * 32) + (lane & 7));
    };
    uint32_t row_w = row_word(0);
    // ---- this thread's gene (coordinate order throughout)
    double p[KT];
    uint32_t inc = 0;
    double* bs = beta_s + lane;
    {
        const double* xp = a.Xty + gene * a.ldv;
        const double* wp = a.V + gene * a.ldv;
        double mx = 0.0;
#pragma unroll
        for (int c = 0; c < KT; ++c) {
            p[c] = (active && c < K) ? xp[c] : 0.0;
            mx = fmax(mx, fabs(p[c]));
        }
        const double thr = a.alpha * (2.0 * a.lambda - mx);                   // coordinate_descent.cpp:74
#pragma unroll
        for (int c = 0; c < KT; ++c) {
            const bool on = active && (c < K) && !(fabs(p[c]) < thr);
            bs[c * 32] = on ? wp[c] : 0.0;                                    // :75-78
            inc |= (on ? 1u : 0u) << c;
        }
    }
    __syncwarp();
    mbar_wait(bar, 0);
    const double* T = tile + lane;
    const unsigned char* Tb = reinterpret_cast<const unsigned char*>(T);
    // p = X'y - (X'X - diag) beta for the screened warm start (:79 in covariance form), coordinates in ascending order
    for (int m = 0; m < K; ++m) {
        double x[KT];
        load_row<KT>(x, Tb, off_s, m);
        const double bm = bs[m * 32];
#pragma unroll
        for (int l = 0; l < KT; ++l) p[l] = fma(-x[l], bm, p[l]);
    }
    int sweeps = 0, cur = 0;
    unsigned long long steps_acc = 0;
    uint32_t draw = 0;
    while (true) {
        if (lane < 8) reinterpret_cast<uint32_t*>(ord_s + 32 * cur)[lane] = row_w;
        row_w = row_word(draw + 1);                                           // consumed by the next sweep
        __syncwarp();
        const unsigned char* oc = ord_s + 32 * cur;
        double dl = 0.0;
        const uint32_t incs = active ? inc : 0u;                              // finished genes: every step is a no-op
        // (fetching the next step's row one step early into a second register set - 208 registers - was measured slower: 485
        //  against 445 us per steady-state launch, profiles/r02_masked_solver_versions.txt)
        int k = oc[0];
        for (int i = 0; i < K; ++i) {
            const int kn = oc[(i + 1 < K) ? i + 1 : i];
            double x[KT];
            load_row<KT>(x, Tb, off_s, k);                                    // depends on k only: in flight during the scalar chain
            const double d = T[(size_t)(NTRI + k) * 32];
            const double rinv = T[(size_t)(NTRI + KT + k) * 32];
            const double bo = bs[k * 32];
            const bool on = (incs >> k) & 1u;
            const double up = select_reg<KT>(p, k);                           // coordinate_descent.cpp:94 (p form)
            const double t1 = fabs(up) - la;
            double nb = copysign(t1, up) * rinv;                              // :99-104
            nb = (__double2hiint(t1) >= 0) ? nb : 0.0;
            nb = on ? nb : bo;
            const double dlt = nb - bo;
            const double hden = fma(d, 0.5, hl2);                             // (XtX_kk + l2) / 2, exactly
            dl = fma(dlt, fma(hden, nb + bo, -up), dl);
            dl = fma(la, fabs(nb) - fabs(bo), dl);
            bs[k * 32] = nb;                                                  // :106-109
            const double nd = -dlt;
#pragma unroll
            for (int l = 0; l < KT; ++l) p[l] = fma(nd, x[l], p[l]);          // x[k] == 0: p_k does not move
            k = kn;
        }
        // ---- end of the sweep for this gene: inner do-while test (:114), KKT re-admission (:118-124)
        if (active) {
            ++sweeps;
            steps_acc += (unsigned long long)__popc(inc);
            if (!(fabs(dl) > tol) || sweeps >= MAX_SWEEPS_M) {
                uint32_t vmask = 0;
#pragma unroll
                for (int c = 0; c < KT; ++c)
                    if (c < K && !((inc >> c) & 1u) && fabs(p[c]) > la) vmask |= 1u << c;     // |q_e| = |p_e| (beta_e = 0)
                if (vmask == 0u || sweeps >= MAX_SWEEPS_M) {
                    double* vp = a.V + gene * a.ldv;
#pragma unroll
                    for (int c = 0; c < KT; ++c) if (c < K) vp[c] = bs[c * 32];
                    if (a.sweeps_per_gene) a.sweeps_per_gene[gene] = sweeps;
                    active = false;
                } else inc |= vmask;
            }
        }
        if (!__any_sync(FULL, active)) break;
        ++draw; cur ^= 1;
    }
    unsigned long long sw = (slot < a.P) ? (unsigned long long)sweeps : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sw += __shfl_xor_sync(FULL, sw, o); steps_acc += __shfl_xor_sync(FULL, steps_acc, o); }
    if (lane == 0) {
        if (a.sweeps_total && sw) atomicAdd(a.sweeps_total, sw);
        if (a.steps_total && steps_acc) atomicAdd(a.steps_total, steps_acc);
    }
}

template <int KT>
void launch_masked_kt(const CdMaskedArgs& a, cudaStream_t st) {
    const size_t smem = (size_t)(tile_elems(KT) + KT) * 32 * 8 + (size_t)KT * KT * 4 + 64 + 16;
    const int blocks = (int)((a.P + 31) / 32);
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_cd_masked<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    // two one-warp blocks per SM need the full shared-memory carve-out (the driver's default heuristic may settle for less)
    cudaFuncSetAttribute(k_cd_masked<KT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (getenv("INSIDER_B200_TRACE")) {
        static bool once = false;
        if (!once) {
            once = true; int nb = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_cd_masked<KT>, 32, smem);
            fprintf(stderr, "[insider_b200] k_cd_masked<%d>: %zu B shared memory per block, %d resident blocks per SM\n", KT, smem, nb);
        }
    }
    k_cd_masked<KT><<<blocks, 32, smem, st>>>(a);
}

}  // namespace

size_t cd_masked_tile_doubles(int K, int64_t P) {
    const int KT = (K + 3) / 4 * 4;
    return (size_t)((P + 31) / 32) * tile_elems(KT) * 32;
}

void launch_col_gram_tiles(const Geom& g, const uint32_t* trC, const double* U, const double* UtU, const int* order, double lambda, double alpha,
                           double* tiles, cudaStream_t st) {
    if (g.P <= 0) return;
    const int KT = (g.K + 3) / 4 * 4;
    const int blocks = (int)((g.P + 31) / 32);
    const int list_len = std::min(GT_GROUP, round_up(g.N, 32));
    const size_t smem = (size_t)32 * (tile_elems(KT) | 1) * 8 + (size_t)GT_WARPS * list_len * 2;
    const double l2 = lambda * (1.0 - alpha);
#define LAUNCH_GT(SLv)                                                                                                               \
    { if (smem > 48 * 1024) cudaFuncSetAttribute(k_col_gram_tiles<SLv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
      cudaFuncSetAttribute(k_col_gram_tiles<SLv>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);   \
      k_col_gram_tiles<SLv><<<blocks, GT_WARPS * 32, smem, st>>>(trC, U, UtU, order, tiles, g.N, g.K, g.KP, KT, g.Wp, g.P, l2, list_len); }
    switch (g.NT) { case 1: LAUNCH_GT(1) break; case 2: LAUNCH_GT(2) break; case 3: LAUNCH_GT(3) break; default: LAUNCH_GT(4) break; }
#undef LAUNCH_GT
}

void launch_cd_masked(const Geom& g, const double* tiles, const double* Xty, double* V, const CdParams& p, unsigned long long* sweeps,
                      unsigned long long* steps, int* sweeps_per_gene, const int* order, const unsigned char* perm_table, cudaStream_t st) {
    if (g.P <= 0) return;
    CdMaskedArgs a{};
    a.tiles = tiles; a.Xty = Xty; a.V = V; a.ldv = g.ldV; a.K = g.K; a.P = g.P;
    a.lambda = p.lambda; a.alpha = p.alpha; a.la = p.lambda * p.alpha; a.l2 = p.lambda * (1.0 - p.alpha);
    a.tol_dev = p.tol; a.als_iter_dev = p.als_iter; a.seed = p.seed; a.perm_mode = p.perm_mode;
    a.sweeps_total = sweeps; a.steps_total = steps; a.sweeps_per_gene = sweeps_per_gene; a.order = order; a.perm_table = perm_table;
    const int KT = (g.K + 3) / 4 * 4;
    switch (KT / 4) {
        case 1: launch_masked_kt<4>(a, st); break;
        case 2: launch_masked_kt<8>(a, st); break;
        case 3: launch_masked_kt<12>(a, st); break;
        case 4: launch_masked_kt<16>(a, st); break;
        case 5: launch_masked_kt<20>(a, st); break;
        case 6: launch_masked_kt<24>(a, st); break;
        case 7: launch_masked_kt<28>(a, st); break;
        default: launch_masked_kt<32>(a, st); break;
    }
}

}  // namespace ib
