// Streaming passes over the N x P expression matrix (HBM-bound by design; FP64 DMMA for the dense contractions).
//
//   k_row_b    B_partial = (M o Y) V^T          replaces the per-row gathers + dgemv of src/optimize.cpp:161-171 (tuning=1)
//                                               and Xtys = V R^T of src/optimize.cpp:180 (tuning=0), in sufficient-statistic form
//   k_col_xty  Xty = U^T (M o Y)                replaces U[sel,:]^T y[sel] of src/optimize.cpp:216-222 and U^T Y of :235
//   k_sse      sum m (y - u.v)^2, |V|^2, |V|_1  replaces predict + residual + evaluate + the V terms of compute_loss
//                                               (src/utils.cpp:52-102, src/optimize.cpp:377-384) as one fused pass
//
// All three stream tiles of TG = 16 genes x R rows through shared memory with cp.async.bulk (TMA 1-D bulk copies)
// completing on mbarriers, NSTAGE deep, and contract them with mma.sync m8n8k4 f64 (DMMA). Pitches are chosen
// (pitch % 8 == 4) so every fragment load is bank-conflict free. Accumulation order is fixed => bitwise reproducible.
#include <algorithm>
#include <cstdlib>
#include <stdexcept>

#include <cuda.h>

#include "common.cuh"
#include "kernels.cuh"

namespace ib {

namespace {

constexpr int THREADS = 256;
constexpr int NWARPS = 8;
static_assert(NWARPS == 2 * (TG / 4), "k_row_b: two warps per k-step share the V V^T tiles");
constexpr int MT_PER_WARP = 6;          // 8 warps x 6 m-tiles x 8 rows = 384 rows per slab
constexpr int SLAB_ROWS_BIG = 384;      // k_row_b multi-slab
constexpr int SLAB_ROWS_SMALL = 128;    // k_col_xty / k_sse multi-slab
constexpr int SINGLE_SLAB_MAX_N = 384;     // 48 m-tiles of 8 rows

// one TMA tensor copy: the box at (c0 = element along the contiguous dimension, c1 = row of the outer dimension) of a 2-D tensor map
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(smem_u32(dst_smem)),
                 "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

struct StreamArgs {
    const double* Y; const uint32_t* trC; const uint32_t* teC; const double* V; const double* Ut;
    double* out;                 // Bp / Xty / partial
    double* out2;                // k_row_b: Gp[(split*4 + ks)][KP*KP] partial V V^T of k-step ks (slab 0 blocks; warps 2ks, 2ks+1)
    int N, K, KP, ldY, ldV, ldT, Wp;
    int n_tiles;                 // gene tiles in total
    int R, n_slabs, pitchS, pitchU;
    int n_stages;
    int n_splits;                // k_row_b: gene splits (grid.x); others: number of blocks
    int scratch_sep;             // k_col_xty: 1 = dedicated cross-warp scratch (stage buffer too small to alias)
    const int* lv_ptr;           // k_row_b, dense single-slab fast path: sum the block's B rows per confounder level (CSR over all
    const int* lv_rows;          // levels of all confounders: rows lv_rows[lv_ptr[l] .. lv_ptr[l+1]) ascending) and store
    int n_levels, n_lv_rows;     // out[split][n_levels][KP] instead of out[split][N][KP] (k_level_sumB folded into the epilogue)
};

// balanced partition of n items over parts
__device__ __forceinline__ void split_range(int n, int parts, int idx, int& b, int& e) {
    int q = n / parts, r = n % parts;
    b = idx * q + min(idx, r);
    e = b + q + (idx < r ? 1 : 0);
}

// zero the masked-out entries of a landed Y piece: thread (c = tid/16, w = tid%16) owns word w of gene c
__device__ __forceinline__ void premask(double* Ys, int pitchS, uint32_t word, int w, int c, int rows_here) {
    int base = 32 * w;
    if (base >= rows_here) return;
    uint32_t z = ~word;
    int lim = rows_here - base;
    if (lim < 32) z &= (1u << lim) - 1u;
    while (z) {
        int b = __ffs(z) - 1;
        z &= z - 1;
        Ys[c * pitchS + base + b] = 0.0;
    }
}

// upper-triangle tile pairs (n1 <= n2) of an NT x NT tile grid in row-major order: idx -> n1, n2 (compile-time recursion so
// that register arrays are only ever indexed by constants)
template <int NT, int IDX, int N1 = 0, bool IN_ROW = (IDX < NT - N1)>
struct GramPair { static constexpr int n1 = GramPair<NT, IDX - (NT - N1), N1 + 1>::n1, n2 = GramPair<NT, IDX - (NT - N1), N1 + 1>::n2; };
template <int NT, int IDX, int N1>
struct GramPair<NT, IDX, N1, true> { static constexpr int n1 = N1, n2 = N1 + IDX; };

// V V^T tiles of one k-step: tiles j (first half, warp 2 ks) or GH + j (second half, warp 2 ks + 1), selected in registers
template <int NT, int J>
__device__ __forceinline__ void gram_steps(double (&gacc)[(NT * (NT + 1) / 2 + 1) / 2][2], const double (&bg)[NT], int ghalf) {
    constexpr int NP = NT * (NT + 1) / 2, GH = (NP + 1) / 2;
    if constexpr (J < GH) {
        constexpr int J1 = (GH + J < NP) ? GH + J : NP - 1;                 // clamped: computed, not stored
        const double ga = ghalf ? bg[GramPair<NT, J1>::n1] : bg[GramPair<NT, J>::n1];
        const double gb = ghalf ? bg[GramPair<NT, J1>::n2] : bg[GramPair<NT, J>::n2];
        dmma(gacc[J][0], gacc[J][1], ga, gb);
        gram_steps<NT, J + 1>(gacc, bg, ghalf);
    }
}
template <int NT, int J>
__device__ __forceinline__ void gram_store(const double (&gacc)[(NT * (NT + 1) / 2 + 1) / 2][2], double* Gp, int KP, int ghalf, int g, int t) {
    constexpr int NP = NT * (NT + 1) / 2, GH = (NP + 1) / 2;
    if constexpr (J < GH) {
        constexpr int J1 = (GH + J < NP) ? GH + J : NP - 1;
        if (!ghalf || GH + J < NP) {
            const int n1 = ghalf ? GramPair<NT, J1>::n1 : GramPair<NT, J>::n1;
            const int n2 = ghalf ? GramPair<NT, J1>::n2 : GramPair<NT, J>::n2;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int ra = 8 * n1 + g, cb = 8 * n2 + 2 * t + e;
                Gp[ra * KP + cb] = gacc[J][e];
                Gp[cb * KP + ra] = gacc[J][e];
            }
        }
        gram_store<NT, J + 1>(gacc, Gp, KP, ghalf, g, t);
    }
}

// ------------------------------------------------------------------------------------------------------------
// k_row_b: grid (n_splits, n_slabs). Each block accumulates B[rows of its slab][K] over its gene tiles.
template <int NT, bool MASKED>
__global__ void __launch_bounds__(THREADS, 1) k_row_b(StreamArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int slab = blockIdx.y, r0 = slab * a.R;
    const int rows_here = min(a.R, a.ldY - r0);
    const int mt_here = (min(a.N - r0, a.R) + 7) / 8;
    int t0, t1;
    split_range(a.n_tiles, a.n_splits, blockIdx.x, t0, t1);
    const int n_items = t1 - t0;

    const int ysz = TG * a.pitchS + 8;       // doubles per Y piece (+slack for the 8-row tile overhang)
    const int vsz = TG * a.ldV;
    const int stage_doubles = ysz + vsz;
    double* stage0 = reinterpret_cast<double*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage0 + (size_t)a.n_stages * stage_doubles);
    const int S = a.n_stages;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    // slack doubles are never written by TMA: zero them once so overhang rows read finite values
    for (int s = tid; s < S * 8; s += THREADS) stage0[(size_t)(s / 8) * stage_doubles + TG * a.pitchS + (s % 8)] = 0.0;
    __syncthreads();

    auto issue = [&](int item) {
        const int s = item % S;
        double* Ys = stage0 + (size_t)s * stage_doubles;
        double* Vs = Ys + ysz;
        const int64_t gene = (int64_t)(t0 + item) * TG;
        const uint32_t ybytes = (uint32_t)rows_here * 8u;
        mbar_expect_tx(&bars[s], ybytes * TG + (uint32_t)vsz * 8u);
        if (a.pitchS == a.ldY && rows_here == a.ldY) {
            tma_load_1d(Ys, a.Y + gene * a.ldY, ybytes * TG, &bars[s]);
        } else {
            for (int c = 0; c < TG; ++c) tma_load_1d(Ys + c * a.pitchS, a.Y + (gene + c) * a.ldY + r0, ybytes, &bars[s]);
        }
        tma_load_1d(Vs, a.V + gene * a.ldV, (uint32_t)vsz * 8u, &bars[s]);
    };
    if (tid == 0) for (int i = 0; i < S - 1 && i < n_items; ++i) issue(i);

    double acc[MT_PER_WARP][NT][2];
#pragma unroll
    for (int i = 0; i < MT_PER_WARP; ++i)
#pragma unroll
        for (int n = 0; n < NT; ++n) acc[i][n][0] = acc[i][n][1] = 0.0;
    // gram = V V^T (src/optimize.cpp:332) rides along (slab 0 only): k-step ks of a tile contributes 4 genes to each of the
    // NP upper-triangle tiles of sum_j v_j v_j^T; warps 2 ks and 2 ks + 1 take half of those tiles each. The DMMAs are issued
    // UNCONDITIONALLY with register-selected operands: a predicated-off DMMA still occupies the pipe (measured: with the gram
    // tiles under `if (ks == warp)` every warp paid for all 24 of them, +33 % DMMA time).
    constexpr int NP = NT * (NT + 1) / 2, GH = (NP + 1) / 2;
    const bool do_gram = (slab == 0) && (a.out2 != nullptr);
    const int gks = warp >> 1, ghalf = warp & 1;
    double gacc[GH][2];
#pragma unroll
    for (int j = 0; j < GH; ++j) gacc[j][0] = gacc[j][1] = 0.0;

    for (int item = 0; item < n_items; ++item) {
        const int s = item % S;
        double* Ys = stage0 + (size_t)s * stage_doubles;
        const double* Vs = Ys + ysz;
        if (tid == 0 && item + S - 1 < n_items) { fence_proxy_async(); issue(item + S - 1); }
        uint32_t word = 0;
        if (MASKED) {
            const int c = tid >> 4, w = tid & 15;
            if (32 * w < rows_here) word = __ldg(a.trC + ((int64_t)(t0 + item) * TG + c) * a.Wp + (r0 >> 5) + w);
        }
        mbar_wait(&bars[s], (uint32_t)((item / S) & 1));
        if (MASKED) {
            premask(Ys, a.pitchS, word, tid & 15, tid >> 4, rows_here);
            __syncthreads();
        }
        if (do_gram) {
            double bg[NT];
#pragma unroll
            for (int n = 0; n < NT; ++n) bg[n] = Vs[(4 * gks + t) * a.ldV + 8 * n + g];
            gram_steps<NT, 0>(gacc, bg, ghalf);
        }
#pragma unroll
        for (int ks = 0; ks < TG / 4; ++ks) {
            double b[NT];
#pragma unroll
            for (int n = 0; n < NT; ++n) b[n] = Vs[(4 * ks + t) * a.ldV + 8 * n + g];
            const double* yrow = Ys + (4 * ks + t) * a.pitchS + g;
#pragma unroll
            for (int i = 0; i < MT_PER_WARP; ++i) {
                const int mt = warp + NWARPS * i;
                if (mt < mt_here) {
                    const double av = yrow[8 * mt];
#pragma unroll
                    for (int n = 0; n < NT; ++n) dmma(acc[i][n][0], acc[i][n][1], av, b[n]);
                }
            }
        }
        __syncthreads();   // stage s may be refilled
    }
    if (do_gram) {
        double* Gp = a.out2 + ((size_t)blockIdx.x * (TG / 4) + gks) * a.KP * a.KP;
        gram_store<NT, 0>(gacc, Gp, a.KP, ghalf, g, t);
    }
    if (a.lv_ptr != nullptr) {
        // Dense fast path (single slab): the row update only needs the sums of B over the rows of every confounder level
        // (k_rows.cu: rhs = SB - G w). Stage this block's B in shared memory (the stage ring is idle: every bulk copy has landed
        // and been consumed) together with the level lists, and sum it per level in the fixed order of the level's row list:
        // 133 x KP instead of 377 x KP doubles per partial, and no separate k_level_sumB launch.
        double* Bs = stage0;                                                  // [N][KP]
        int* ptr_s = reinterpret_cast<int*>(Bs + (size_t)a.N * a.KP);         // [n_levels + 1]
        int* rows_s = ptr_s + a.n_levels + 1;                                 // [n_lv_rows]
        __syncthreads();
        for (int x = tid; x <= a.n_levels; x += THREADS) ptr_s[x] = __ldg(a.lv_ptr + x);
        for (int x = tid; x < a.n_lv_rows; x += THREADS) rows_s[x] = __ldg(a.lv_rows + x);
#pragma unroll
        for (int i = 0; i < MT_PER_WARP; ++i) {
            const int mt = warp + NWARPS * i;
            const int row = 8 * mt + g;
            if (mt < mt_here && row < a.N) {
#pragma unroll
                for (int n = 0; n < NT; ++n) {
                    const int k = 8 * n + 2 * t;
                    Bs[row * a.KP + k] = acc[i][n][0];
                    Bs[row * a.KP + k + 1] = acc[i][n][1];
                }
            }
        }
        __syncthreads();
        double* SBp = a.out + (size_t)blockIdx.x * a.n_levels * a.KP;
        // a warp per level, lanes = columns (KP <= 32), 4 independent partial sums combined in a fixed order
        for (int lv = warp; lv < a.n_levels; lv += NWARPS) {
            if (lane < a.KP) {
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                int r = ptr_s[lv];
                const int re = ptr_s[lv + 1];
                for (; r + 4 <= re; r += 4) {
                    s0 += Bs[rows_s[r] * a.KP + lane]; s1 += Bs[rows_s[r + 1] * a.KP + lane];
                    s2 += Bs[rows_s[r + 2] * a.KP + lane]; s3 += Bs[rows_s[r + 3] * a.KP + lane];
                }
                for (; r < re; ++r) s0 += Bs[rows_s[r] * a.KP + lane];
                SBp[(size_t)lv * a.KP + lane] = (s0 + s1) + (s2 + s3);
            }
        }
        return;
    }
    // store the block's partial: Bp[split][N][KP]
    double* Bp = a.out + (size_t)blockIdx.x * a.N * a.KP;
#pragma unroll
    for (int i = 0; i < MT_PER_WARP; ++i) {
        const int mt = warp + NWARPS * i;
        const int row = r0 + 8 * mt + g;
        if (mt < mt_here && row < a.N) {
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                const int k = 8 * n + 2 * t;
                if (k < a.KP) Bp[(size_t)row * a.KP + k] = acc[i][n][0];
                if (k + 1 < a.KP) Bp[(size_t)row * a.KP + k + 1] = acc[i][n][1];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// k_col_xty: grid (n_blocks), single-slab geometry (N <= 384 rows; launch_col_xty sends the rest to k_col_xty_slabs). Items =
// the block's gene tiles, streamed through one ring of stages. The 8 warps form TWO GROUPS of 4 (one warp of each group per SM sub-partition) that take alternate tiles: inside a group
// the reduction over rows is split over the 4 warps and combined through shared memory in a fixed order; the groups only
// meet in the stage ring. While one group combines and stores its tile (no DMMA work, 3 group barriers) the other is in
// its contraction and keeps the DMMA pipe busy - one warp with 6 independent accumulators saturates a sub-partition's
// pipe. (With all 8 warps on the same tile the pipe idled 26 % of the kernel in that epilogue: pc sampling in
// profiles/r01_ncu_streaming_kernels_dense_v2.txt.)
// Stage hand-over: item i + S is issued into item i's stage by the group that consumed item i. An mbarrier parity wait can
// only tell the current phase from the one before it, and the two groups are not ordered against each other, so a group
// first waits (issued[]) until ITS item has been issued into the stage and only then for the stage's phase.
constexpr int XG_WARPS = NWARPS / 2;           // warps per group
constexpr int XG_THREADS = XG_WARPS * 32;

__device__ __forceinline__ void group_bar(int grp) { asm volatile("bar.sync %0, %1;\n" ::"r"(1 + grp), "r"(XG_THREADS) : "memory"); }

template <int NT, bool MASKED, bool RESIDENT_U>
__global__ void __launch_bounds__(THREADS, 1) k_col_xty(StreamArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int grp = warp / XG_WARPS, gw = warp % XG_WARPS, gtid = tid % XG_THREADS;
    int t0, t1;
    split_range(a.n_tiles, a.n_splits, blockIdx.x, t0, t1);
    const int n_tiles_blk = t1 - t0;
    const int n_items = n_tiles_blk * a.n_slabs;
    const int S = a.n_stages;

    const int ysz = TG * a.pitchS;
    const int usz = a.KP * a.pitchU;
    const int stage_doubles = ysz + (RESIDENT_U ? 0 : usz);
    double* Ures = reinterpret_cast<double*>(smem_raw);                       // resident Ut (if any)
    double* stage0 = Ures + (RESIDENT_U ? usz : 0);
    double* scratch_own = stage0 + (size_t)S * stage_doubles;                // [2][XG_WARPS][NT*2*64] when scratch_sep
    uint64_t* bars = reinterpret_cast<uint64_t*>(scratch_own + (a.scratch_sep ? NWARPS * NT * 2 * 64 : 0));   // S stage barriers + 1 for Ut
    volatile int* issued = reinterpret_cast<volatile int*>(bars + 5);         // [4] last item issued into each stage (S <= 4; 8 words reserved)

    if (tid == 0) {
        for (int s = 0; s <= S; ++s) mbar_init(&bars[s], 1);
        for (int s = 0; s < 4; ++s) issued[s] = -1;
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](int item) {
        const int s = item % S;
        double* Ys = stage0 + (size_t)s * stage_doubles;
        const int tile = t1 - 1 - item / a.n_slabs, slab = item % a.n_slabs;      // (backwards: see below)
        const int r0 = slab * a.R;
        const int rows_here = min(a.R, a.ldY - r0);
        const int64_t gene = (int64_t)tile * TG;
        const uint32_t ybytes = (uint32_t)rows_here * 8u;
        uint32_t total = ybytes * TG;
        if (!RESIDENT_U) total += (uint32_t)a.KP * ybytes;
        mbar_expect_tx(&bars[s], total);
        if (a.pitchS == a.ldY && rows_here == a.ldY) {
            tma_load_1d(Ys, a.Y + gene * a.ldY, ybytes * TG, &bars[s]);
        } else {
            for (int c = 0; c < TG; ++c) tma_load_1d(Ys + c * a.pitchS, a.Y + (gene + c) * a.ldY + r0, ybytes, &bars[s]);
        }
        if (!RESIDENT_U) {
            double* Us = Ys + ysz;
            for (int k = 0; k < a.KP; ++k) tma_load_1d(Us + k * a.pitchU, a.Ut + (size_t)k * a.ldT + r0, ybytes, &bars[s]);
        }
        issued[s] = item;
    };
    if (tid == 0) {
        if (RESIDENT_U) {
            mbar_expect_tx(&bars[S], (uint32_t)usz * 8u);
            tma_load_1d(Ures, a.Ut, (uint32_t)usz * 8u, &bars[S]);
        }
        for (int i = 0; i < S && i < n_items; ++i) issue(i);                  // item i + S is issued by the group that consumed item i
    }
    __syncthreads();                                                          // issued[] of the first S items is visible to both groups
    if (RESIDENT_U) mbar_wait(&bars[S], 0);

    double acc[NT][2][2];
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int m = 0; m < 2; ++m) acc[n][m][0] = acc[n][m][1] = 0.0;

    for (int tl = grp; tl < n_tiles_blk; tl += 2) {
        // The block walks its tile range BACKWARDS: k_row_b, the other pass over Y of an iteration, walks the same range forwards, so each
        // pass starts on the tiles the previous one touched last - the part of the 134 MB matrix that is still in the 126 MB L2
        // (measured: k_row_b 45.7 -> 44.0 us, k_col_xty 40.4 -> 39.5 us; two blocks per SM for k_row_b: 46.1 us, dropped).
        const int tile = t1 - 1 - tl;
        for (int slab = 0; slab < a.n_slabs; ++slab) {
            const int item = tl * a.n_slabs + slab;
            const int s = item % S;
            double* Ys = stage0 + (size_t)s * stage_doubles;
            const double* Us = RESIDENT_U ? Ures : (Ys + ysz);
            const int r0 = slab * a.R;
            const int rows_here = min(a.R, a.ldY - r0);
            uint32_t word[2] = {0, 0};
            if (MASKED) {
                const int w = gtid & 15;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int c = (gtid >> 4) + 8 * h;
                    if (32 * w < rows_here) word[h] = __ldg(a.trC + ((int64_t)tile * TG + c) * a.Wp + (r0 >> 5) + w);
                }
            }
            while (issued[s] < item) { }                                       // see "stage hand-over" above
            mbar_wait(&bars[s], (uint32_t)((item / S) & 1));
            if (MASKED) {
#pragma unroll
                for (int h = 0; h < 2; ++h) premask(Ys, a.pitchS, word[h], gtid & 15, (gtid >> 4) + 8 * h, rows_here);
                group_bar(grp);
            }
            const int KS = rows_here >> 2;
            int k0, k1;
            split_range(KS, XG_WARPS, gw, k0, k1);
            for (int ks = k0; ks < k1; ++ks) {
                double av[NT], bv[2];
#pragma unroll
                for (int n = 0; n < NT; ++n) av[n] = Us[(8 * n + g) * a.pitchU + 4 * ks + t];
#pragma unroll
                for (int m = 0; m < 2; ++m) bv[m] = Ys[(8 * m + g) * a.pitchS + 4 * ks + t];
#pragma unroll
                for (int n = 0; n < NT; ++n)
#pragma unroll
                    for (int m = 0; m < 2; ++m) dmma(acc[n][m][0], acc[n][m][1], av[n], bv[m]);
            }
            if (slab == a.n_slabs - 1) {
                // cross-warp reduction of the group in a fixed order, then store the K x 16 tile of Xty. The scratch
                // [XG_WARPS][NT*2*64] aliases this item's (fully consumed) stage buffer when that is large enough.
                double* scratch = a.scratch_sep ? scratch_own + (size_t)grp * XG_WARPS * (NT * 2 * 64) : Ys;
                group_bar(grp);
                double* sc = scratch + gw * (NT * 2 * 64);
#pragma unroll
                for (int n = 0; n < NT; ++n)
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        sc[(n * 2 + m) * 64 + lane * 2 + 0] = acc[n][m][0];
                        sc[(n * 2 + m) * 64 + lane * 2 + 1] = acc[n][m][1];
                        acc[n][m][0] = acc[n][m][1] = 0.0;
                    }
                group_bar(grp);
                for (int x = gtid; x < NT * 2 * 64; x += XG_THREADS) {
                    double sum = 0.0;
#pragma unroll
                    for (int w = 0; w < XG_WARPS; ++w) sum += scratch[w * (NT * 2 * 64) + x];
                    const int tl2 = x >> 6, ln = (x & 63) >> 1, e = x & 1;
                    const int n = tl2 >> 1, m = tl2 & 1;
                    const int k = 8 * n + (ln >> 2), gene = 8 * m + 2 * (ln & 3) + e;
                    if (k < a.ldV) a.out[((int64_t)tile * TG + gene) * a.ldV + k] = sum;
                }
            }
            group_bar(grp);                                                    // the group is done with stage s: refill it
            if (gtid == 0 && item + S < n_items) { fence_proxy_async(); issue(item + S); }
        }
    }
}

// k_col_xty_slabs: the multi-slab form (N > 384 rows: U^T does not fit in shared memory and a tile's accumulators live across its
// row slabs). Items = (group of XW_GT gene tiles, slab) of the block's tile range, all 8 warps on the same item; the reduction over
// the slab's rows is split over warps. The slab of U^T travels with every item, so the group is XW_GT = 4 tiles wide: 64 KB of Y
// per 32 KB of U^T, and every A fragment of U^T is reused for 4 x 2 B fragments. An item arrives by TWO tensor-map TMA copies
// (boxes of pitchS rows x 64 genes of Y and pitchS rows x KP rows of U^T; the box is 4 rows wider than the slab so that the
// dense box IS the bank-conflict-free pitch): measured on the GTEx-scale shard, a bulk copy costs the SM ~100 cycles whatever
// its size, and round 1's 48 copies of 1 KB per 16-gene item ran at 0.15 of the HBM roofline (profiles/r02_gtex_shard_streaming.txt).
constexpr int XW_GT = 4;

template <int NT, bool MASKED>
__global__ void __launch_bounds__(THREADS, 1) k_col_xty_slabs(StreamArgs a, const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmU) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    int t0, t1;
    split_range(a.n_tiles, a.n_splits, blockIdx.x, t0, t1);
    const int n_groups = (t1 - t0 + XW_GT - 1) / XW_GT;
    const int n_items = n_groups * a.n_slabs;
    const int S = a.n_stages;

    const int ysz = XW_GT * TG * a.pitchS;
    const int usz = a.KP * a.pitchU;
    const int stage_doubles = ysz + usz;
    double* stage0 = reinterpret_cast<double*>(smem_raw);
    double* scratch_own = stage0 + (size_t)S * stage_doubles;                // [NWARPS][NT*2*64] when scratch_sep
    uint64_t* bars = reinterpret_cast<uint64_t*>(scratch_own + (a.scratch_sep ? NWARPS * NT * 2 * 64 : 0));   // S stage barriers

    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // (a partial last group still loads the full box: genes of the next block's range or zero fill beyond P_pad, never used)
    auto issue = [&](int item) {
        const int s = item % S;
        double* Ys = stage0 + (size_t)s * stage_doubles;
        const int grp = item / a.n_slabs, slab = item % a.n_slabs;
        mbar_expect_tx(&bars[s], (uint32_t)stage_doubles * 8u);
        tma_load_2d(Ys, &tmY, slab * a.R, (t0 + grp * XW_GT) * TG, &bars[s]);
        tma_load_2d(Ys + ysz, &tmU, slab * a.R, 0, &bars[s]);
    };
    if (tid == 0)
        for (int i = 0; i < S - 1 && i < n_items; ++i) issue(i);

    double acc[XW_GT][NT][2][2];
#pragma unroll
    for (int j = 0; j < XW_GT; ++j)
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int m = 0; m < 2; ++m) acc[j][n][m][0] = acc[j][n][m][1] = 0.0;

    for (int item = 0; item < n_items; ++item) {
        const int s = item % S;
        double* Ys = stage0 + (size_t)s * stage_doubles;
        const double* Us = Ys + ysz;
        const int grp = item / a.n_slabs, slab = item % a.n_slabs;
        const int tile0 = t0 + grp * XW_GT;
        const int nt = min(XW_GT, t1 - tile0);
        const int r0 = slab * a.R;
        const int rows_here = min(a.R, a.ldY - r0);
        if (tid == 0 && item + S - 1 < n_items) { fence_proxy_async(); issue(item + S - 1); }
        uint32_t word[XW_GT];
        if (MASKED) {
            const int c = tid >> 4, w = tid & 15;
#pragma unroll
            for (int j = 0; j < XW_GT; ++j) {
                word[j] = 0;
                if (j < nt && 32 * w < rows_here) word[j] = __ldg(a.trC + ((int64_t)(tile0 + j) * TG + c) * a.Wp + (r0 >> 5) + w);
            }
        }
        mbar_wait(&bars[s], (uint32_t)((item / S) & 1));
        if (MASKED) {
#pragma unroll
            for (int j = 0; j < XW_GT; ++j)
                if (j < nt) premask(Ys + j * TG * a.pitchS, a.pitchS, word[j], tid & 15, tid >> 4, rows_here);
            __syncthreads();
        }
        const int KS = rows_here >> 2;
        int k0, k1;
        split_range(KS, NWARPS, warp, k0, k1);
        for (int ks = k0; ks < k1; ++ks) {
            double av[NT];
#pragma unroll
            for (int n = 0; n < NT; ++n) av[n] = Us[(8 * n + g) * a.pitchU + 4 * ks + t];
#pragma unroll
            for (int j = 0; j < XW_GT; ++j) {
                if (j < nt) {
                    double bv[2];
#pragma unroll
                    for (int m = 0; m < 2; ++m) bv[m] = Ys[(j * TG + 8 * m + g) * a.pitchS + 4 * ks + t];
#pragma unroll
                    for (int n = 0; n < NT; ++n)
#pragma unroll
                        for (int m = 0; m < 2; ++m) dmma(acc[j][n][m][0], acc[j][n][m][1], av[n], bv[m]);
                }
            }
        }
        if (slab == a.n_slabs - 1) {
            // cross-warp reduction in a fixed order, tile by tile, then store the K x 16 tile of Xty. The scratch
            // [NWARPS][NT*2*64] aliases this item's (fully consumed) stage buffer when that is large enough.
            double* scratch = a.scratch_sep ? scratch_own : Ys;
#pragma unroll
            for (int j = 0; j < XW_GT; ++j) {
                if (j < nt) {
                    __syncthreads();
                    double* sc = scratch + warp * (NT * 2 * 64);
#pragma unroll
                    for (int n = 0; n < NT; ++n)
#pragma unroll
                        for (int m = 0; m < 2; ++m) {
                            sc[(n * 2 + m) * 64 + lane * 2 + 0] = acc[j][n][m][0];
                            sc[(n * 2 + m) * 64 + lane * 2 + 1] = acc[j][n][m][1];
                            acc[j][n][m][0] = acc[j][n][m][1] = 0.0;
                        }
                    __syncthreads();
                    for (int x = tid; x < NT * 2 * 64; x += THREADS) {
                        double sum = 0.0;
#pragma unroll
                        for (int w = 0; w < NWARPS; ++w) sum += scratch[w * (NT * 2 * 64) + x];
                        const int tl = x >> 6, ln = (x & 63) >> 1, e = x & 1;
                        const int n = tl >> 1, m = tl & 1;
                        const int k = 8 * n + (ln >> 2), gene = 8 * m + 2 * (ln & 3) + e;
                        if (k < a.ldV) a.out[((int64_t)(tile0 + j) * TG + gene) * a.ldV + k] = sum;
                    }
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------------------
// k_sse: grid (n_blocks). pred = U V per (tile, slab), residual against the Y piece, masked sums of squares.
// Stage layout: resident U^T: [Y piece + 8 | V tile]; row slabs (N > 384): [Y box | U^T box | V tile], the two boxes by one
// tensor-map TMA copy each (see k_col_xty_slabs: 3 copies per item instead of 49).
template <int NT, bool MASKED, bool RESIDENT_U>
__global__ void __launch_bounds__(THREADS, 1) k_sse(StreamArgs a, const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmU) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    int t0, t1;
    split_range(a.n_tiles, a.n_splits, blockIdx.x, t0, t1);
    const int n_items = (t1 - t0) * a.n_slabs;
    const int S = a.n_stages;

    const int ysz = TG * a.pitchS + (RESIDENT_U ? 8 : 0);
    const int vsz = TG * a.ldV;
    const int usz = a.KP * a.pitchU;
    const int stage_doubles = RESIDENT_U ? ysz + vsz : (ysz + usz + vsz + 15) / 16 * 16;      // (boxes land 128-byte aligned)
    double* Ures = reinterpret_cast<double*>(smem_raw);
    double* stage0 = Ures + (RESIDENT_U ? usz : 0);
    double* red = stage0 + (size_t)S * stage_doubles;                          // [NWARPS][4]
    uint32_t* mk = reinterpret_cast<uint32_t*>(red + NWARPS * 4);              // [2][TG][16] mask words (train, test)
    uint64_t* bars = reinterpret_cast<uint64_t*>(mk + 2 * TG * 16);

    if (tid == 0) {
        for (int s = 0; s <= S; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](int item) {
        const int s = item % S;
        double* Ys = stage0 + (size_t)s * stage_doubles;
        const int tile = t0 + item / a.n_slabs, slab = item % a.n_slabs;
        const int r0 = slab * a.R;
        const int64_t gene = (int64_t)tile * TG;
        if (!RESIDENT_U) {
            mbar_expect_tx(&bars[s], (uint32_t)(ysz + usz + vsz) * 8u);
            tma_load_2d(Ys, &tmY, r0, (int)gene, &bars[s]);
            tma_load_2d(Ys + ysz, &tmU, r0, 0, &bars[s]);
            tma_load_1d(Ys + ysz + usz, a.V + gene * a.ldV, (uint32_t)vsz * 8u, &bars[s]);
            return;
        }
        double* Vs = Ys + ysz;
        const int rows_here = min(a.R, a.ldY - r0);
        const uint32_t ybytes = (uint32_t)rows_here * 8u;
        mbar_expect_tx(&bars[s], ybytes * TG + (uint32_t)vsz * 8u);
        if (a.pitchS == a.ldY && rows_here == a.ldY) {
            tma_load_1d(Ys, a.Y + gene * a.ldY, ybytes * TG, &bars[s]);
        } else {
            for (int c = 0; c < TG; ++c) tma_load_1d(Ys + c * a.pitchS, a.Y + (gene + c) * a.ldY + r0, ybytes, &bars[s]);
        }
        tma_load_1d(Vs, a.V + gene * a.ldV, (uint32_t)vsz * 8u, &bars[s]);
    };
    if (tid == 0) {
        if (RESIDENT_U) {
            mbar_expect_tx(&bars[S], (uint32_t)usz * 8u);
            tma_load_1d(Ures, a.Ut, (uint32_t)usz * 8u, &bars[S]);
        }
        for (int i = 0; i < S - 1 && i < n_items; ++i) issue(i);
    }
    if (RESIDENT_U) mbar_wait(&bars[S], 0);

    double sse_tr = 0.0, sse_te = 0.0, v2 = 0.0, v1 = 0.0;
    for (int item = 0; item < n_items; ++item) {
        const int s = item % S;
        const double* Ys = stage0 + (size_t)s * stage_doubles;
        const double* Vs = RESIDENT_U ? Ys + ysz : Ys + ysz + usz;
        const double* Us = RESIDENT_U ? Ures : (Ys + ysz);
        const int tile = t0 + item / a.n_slabs, slab = item % a.n_slabs;
        const int r0 = slab * a.R;
        const int rows_here = min(a.R, a.ldY - r0);
        const int mt_here = (min(a.N - r0, a.R) + 7) / 8;
        if (tid == 0 && item + S - 1 < n_items) { fence_proxy_async(); issue(item + S - 1); }
        if (MASKED) {
            const int c = tid >> 4, w = tid & 15;
            uint32_t wtr = 0, wte = 0;
            if (32 * w < rows_here) {
                const int64_t idx = ((int64_t)tile * TG + c) * a.Wp + (r0 >> 5) + w;
                wtr = __ldg(a.trC + idx);
                wte = __ldg(a.teC + idx);
            }
            mk[tid] = wtr;
            mk[TG * 16 + tid] = wte;
        }
        mbar_wait(&bars[s], (uint32_t)((item / S) & 1));
        __syncthreads();
        if (slab == 0) {
            for (int x = tid; x < vsz; x += THREADS) { const double v = Vs[x]; v2 = fma(v, v, v2); v1 += fabs(v); }
        }
        double bv[2][NT * 2];
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int ks = 0; ks < NT * 2; ++ks) bv[m][ks] = Vs[(8 * m + g) * a.ldV + 4 * ks + t];
        for (int mt = warp; mt < mt_here; mt += NWARPS) {
            double p[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
            for (int ks = 0; ks < NT * 2; ++ks) {
                const double av = Us[(4 * ks + t) * a.pitchU + 8 * mt + g];
#pragma unroll
                for (int m = 0; m < 2; ++m) dmma(p[m][0], p[m][1], av, bv[m][ks]);
            }
            const int lr = 8 * mt + g;              // row within the slab
            const bool row_ok = (r0 + lr) < a.N;
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int gene = 8 * m + 2 * t + e;
                    const double r = Ys[gene * a.pitchS + lr] - p[m][e];
                    const double r2 = r * r;
                    if (MASKED) {
                        const uint32_t wtr = mk[gene * 16 + (lr >> 5)], wte = mk[TG * 16 + gene * 16 + (lr >> 5)];
                        if (row_ok && ((wtr >> (lr & 31)) & 1u)) sse_tr += r2;
                        if (row_ok && ((wte >> (lr & 31)) & 1u)) sse_te += r2;
                    } else {
                        if (row_ok) sse_tr += r2;
                    }
                }
        }
        __syncthreads();
    }
    sse_tr = warp_sum(sse_tr); sse_te = warp_sum(sse_te); v2 = warp_sum(v2); v1 = warp_sum(v1);
    if (lane == 0) { red[warp * 4 + 0] = sse_tr; red[warp * 4 + 1] = sse_te; red[warp * 4 + 2] = v2; red[warp * 4 + 3] = v1; }
    __syncthreads();
    if (tid < 4) {
        double s = 0.0;
        for (int w = 0; w < NWARPS; ++w) s += red[w * 4 + tid];
        a.out[(size_t)blockIdx.x * 4 + tid] = s;
    }
}

// ------------------------------------------------------------------------------------------------------------
constexpr size_t SMEM_LIMIT = 227 * 1024;

// The opt-in is a per-device function attribute: no caching here (a context on another device must opt in again and
// the call is cheap next to a kernel launch).
template <typename KernelT>
void set_smem(KernelT k, size_t bytes) {
    if (bytes > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

StreamArgs base_args(const Geom& g) {
    StreamArgs a{};
    a.N = g.N; a.K = g.K; a.KP = g.KP; a.ldY = g.ldY; a.ldV = g.ldV; a.ldT = g.ldT; a.Wp = g.Wp; a.n_tiles = g.n_tiles;
    return a;
}

}  // namespace

int row_b_default_splits(const Geom& g, int sm_count) {
    const int n_slabs = (g.N <= SINGLE_SLAB_MAX_N) ? 1 : (g.ldY + SLAB_ROWS_BIG - 1) / SLAB_ROWS_BIG;
    int splits = sm_count / n_slabs;
    if (splits < 1) splits = 1;
    if (splits > g.n_tiles) splits = g.n_tiles;
    return splits;
}
size_t row_b_partial_elems(const Geom& g, int n_splits) { return (size_t)n_splits * g.N * g.KP; }
int stream_default_blocks(const Geom& g, int sm_count) { return sm_count < g.n_tiles ? sm_count : g.n_tiles; }

bool row_b_levels_supported(const Geom& g, int n_levels, int n_lv_rows) {
    // single slab, and the staged N x KP block + the level lists fit in two stages of the ring (a stage holds 16 x ldY doubles)
    return g.N <= SINGLE_SLAB_MAX_N && (size_t)g.N * g.KP + ((size_t)n_levels + 1 + n_lv_rows + 1) / 2 <= 2 * ((size_t)TG * g.ldY + 8 + (size_t)TG * g.ldV);
}

void launch_row_b(const Geom& g, bool masked, const double* Y, const uint32_t* trC, const double* V, double* Bp, double* Gp, int n_splits,
                  cudaStream_t st) { launch_row_b_ex(g, masked, Y, trC, V, Bp, Gp, n_splits, nullptr, nullptr, 0, 0, st); }

void launch_row_b_ex(const Geom& g, bool masked, const double* Y, const uint32_t* trC, const double* V, double* Bp, double* Gp, int n_splits,
                     const int* lv_ptr, const int* lv_rows, int n_levels, int n_lv_rows, cudaStream_t st) {
    StreamArgs a = base_args(g);
    a.Y = Y; a.trC = trC; a.V = V; a.out = Bp; a.out2 = Gp; a.n_splits = n_splits;
    a.lv_ptr = lv_ptr; a.lv_rows = lv_rows; a.n_levels = n_levels; a.n_lv_rows = n_lv_rows;
    if (g.N <= SINGLE_SLAB_MAX_N) { a.R = g.ldY; a.n_slabs = 1; a.pitchS = g.ldY; }
    else { a.R = SLAB_ROWS_BIG; a.n_slabs = (g.ldY + a.R - 1) / a.R; a.pitchS = pitch4(a.R); }
    const size_t stage = ((size_t)TG * a.pitchS + 8 + (size_t)TG * g.ldV) * 8;
    int S = (int)((SMEM_LIMIT - 64) / stage);
    if (S > 4) S = 4;
    if (S < 2) S = 2;
    a.n_stages = S;
    const size_t smem = S * stage + 8 * (S + 1);
    dim3 grid(n_splits, a.n_slabs);
#define LAUNCH_RB(NTv)                                                                                         \
    if (masked) { set_smem(k_row_b<NTv, true>, smem); k_row_b<NTv, true><<<grid, THREADS, smem, st>>>(a); }    \
    else { set_smem(k_row_b<NTv, false>, smem); k_row_b<NTv, false><<<grid, THREADS, smem, st>>>(a); }
    switch (g.NT) { case 1: LAUNCH_RB(1) break; case 2: LAUNCH_RB(2) break; case 3: LAUNCH_RB(3) break; default: LAUNCH_RB(4) break; }
#undef LAUNCH_RB
}

namespace {
// 2-D tensor map over a pitched FP64 matrix [outer][pitch] (box: box_inner contiguous elements x box_outer rows), through the
// driver entry point (the library links only the runtime)
CUtensorMap tensor_map_2d(const double* base, uint64_t inner, uint64_t outer, uint64_t pitch_elems, uint32_t box_inner, uint32_t box_outer) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) fn = nullptr;
        return reinterpret_cast<EncodeFn>(fn);
    }();
    alignas(64) CUtensorMap m{};
    const cuuint64_t dims[2] = {inner, outer};
    const cuuint64_t strides[1] = {pitch_elems * 8};
    const cuuint32_t box[2] = {box_inner, box_outer};
    const cuuint32_t estr[2] = {1, 1};
    if (!encode || encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        throw std::runtime_error("cuTensorMapEncodeTiled failed");
    return m;
}

// geometry shared by k_col_xty and k_sse
bool resident_geometry(const Geom& g, StreamArgs& a, size_t extra_stage_doubles, size_t fixed_bytes) {
    const size_t ures = (size_t)g.KP * g.ldT * 8;
    const size_t stage = ((size_t)TG * g.ldY + extra_stage_doubles) * 8;
    if (g.N <= SINGLE_SLAB_MAX_N && ures + 2 * stage + fixed_bytes <= SMEM_LIMIT) {
        a.R = g.ldY; a.n_slabs = 1; a.pitchS = g.ldY; a.pitchU = g.ldT;
        int S = (int)((SMEM_LIMIT - ures - fixed_bytes) / stage);
        a.n_stages = S > 4 ? 4 : S;
        return true;
    }
    a.R = SLAB_ROWS_SMALL; a.n_slabs = (g.ldY + a.R - 1) / a.R; a.pitchS = pitch4(a.R); a.pitchU = a.pitchS;
    const size_t stage2 = ((size_t)TG * a.pitchS + (size_t)g.KP * a.pitchU + extra_stage_doubles) * 8;
    int S = (int)((SMEM_LIMIT - fixed_bytes) / stage2);
    a.n_stages = S > 4 ? 4 : (S < 2 ? 2 : S);
    return false;
}
}  // namespace

void launch_col_xty(const Geom& g, bool masked, const double* Y, const uint32_t* trC, const double* Ut, double* Xty, int n_blocks,
                    cudaStream_t st) {
    StreamArgs a = base_args(g);
    a.Y = Y; a.trC = trC; a.Ut = Ut; a.out = Xty; a.n_splits = n_blocks;
    const size_t need = (size_t)NWARPS * g.NT * 2 * 64;                       // cross-warp scratch, doubles
    const bool small = (size_t)TG * g.ldY < need;                             // resident path's stage cannot hold it
    const size_t fixed = 8 * 8 + (small ? need * 8 : 0);
    const bool res = resident_geometry(g, a, 0, fixed);
    size_t stage, smem;
    alignas(64) CUtensorMap tmY{}, tmU{};
    if (res) {
        stage = (size_t)TG * a.pitchS * 8;
        a.scratch_sep = (stage / 8 < need) ? 1 : 0;
        smem = (size_t)g.KP * a.pitchU * 8 + a.n_stages * stage + 8 * 8 + (a.scratch_sep ? need * 8 : 0);
    } else {
        // row slabs (N > 384): XW_GT tiles + the slab of U^T per stage, double buffered
        stage = ((size_t)XW_GT * TG * a.pitchS + (size_t)g.KP * a.pitchU) * 8;
        a.n_stages = (int)std::min<size_t>(4, std::max<size_t>(2, (SMEM_LIMIT - 8 * 8) / stage));
        a.scratch_sep = ((size_t)XW_GT * TG * a.pitchS < need) ? 1 : 0;
        smem = a.n_stages * stage + 8 * 8 + (a.scratch_sep ? need * 8 : 0);
        tmY = tensor_map_2d(Y, (uint64_t)g.ldY, (uint64_t)g.P_pad, (uint64_t)g.ldY, (uint32_t)a.pitchS, (uint32_t)(XW_GT * TG));
        tmU = tensor_map_2d(Ut, (uint64_t)g.ldT, (uint64_t)g.KP, (uint64_t)g.ldT, (uint32_t)a.pitchU, (uint32_t)g.KP);
    }
    // single slab with U^T resident: two-group kernel; row slabs: all warps on one item, accumulators across slabs
#define LAUNCH_CX(NTv, M) { set_smem(k_col_xty<NTv, M, true>, smem); k_col_xty<NTv, M, true><<<n_blocks, THREADS, smem, st>>>(a); }
#define LAUNCH_CS(NTv, M) { set_smem(k_col_xty_slabs<NTv, M>, smem); k_col_xty_slabs<NTv, M><<<n_blocks, THREADS, smem, st>>>(a, tmY, tmU); }
#define LAUNCH_CX2(NTv)                                                       \
    if (masked) { if (res) LAUNCH_CX(NTv, true) else LAUNCH_CS(NTv, true) } \
    else { if (res) LAUNCH_CX(NTv, false) else LAUNCH_CS(NTv, false) }
    switch (g.NT) { case 1: LAUNCH_CX2(1) break; case 2: LAUNCH_CX2(2) break; case 3: LAUNCH_CX2(3) break; default: LAUNCH_CX2(4) break; }
#undef LAUNCH_CX2
#undef LAUNCH_CX
#undef LAUNCH_CS
}

void launch_sse(const Geom& g, bool masked, const double* Y, const uint32_t* trC, const uint32_t* teC, const double* Ut, const double* V,
                double* partial, int n_blocks, cudaStream_t st) {
    StreamArgs a = base_args(g);
    a.Y = Y; a.trC = trC; a.teC = teC; a.Ut = Ut; a.V = V; a.out = partial; a.n_splits = n_blocks;
    const size_t fixed = (size_t)NWARPS * 4 * 8 + 2 * TG * 16 * 4 + 8 * 8;
    const size_t extra = 8 + (size_t)TG * g.ldV;
    const bool res = resident_geometry(g, a, extra, fixed);
    size_t stage = ((size_t)TG * a.pitchS + extra) * 8;
    alignas(64) CUtensorMap tmY{}, tmU{};
    if (!res) {
        stage = ((size_t)TG * a.pitchS + (size_t)g.KP * a.pitchU + (size_t)TG * g.ldV + 15) / 16 * 16 * 8;
        a.n_stages = (int)std::min<size_t>(4, std::max<size_t>(2, (SMEM_LIMIT - fixed) / stage));
        tmY = tensor_map_2d(Y, (uint64_t)g.ldY, (uint64_t)g.P_pad, (uint64_t)g.ldY, (uint32_t)a.pitchS, (uint32_t)TG);
        tmU = tensor_map_2d(Ut, (uint64_t)g.ldT, (uint64_t)g.KP, (uint64_t)g.ldT, (uint32_t)a.pitchU, (uint32_t)g.KP);
    }
    const size_t smem = (res ? (size_t)g.KP * a.pitchU * 8 : 0) + a.n_stages * stage + fixed;
#define LAUNCH_SS(NTv, M, RS) { set_smem(k_sse<NTv, M, RS>, smem); k_sse<NTv, M, RS><<<n_blocks, THREADS, smem, st>>>(a, tmY, tmU); }
#define LAUNCH_SS2(NTv)                                                       \
    if (masked) { if (res) LAUNCH_SS(NTv, true, true) else LAUNCH_SS(NTv, true, false) } \
    else { if (res) LAUNCH_SS(NTv, false, true) else LAUNCH_SS(NTv, false, false) }
    switch (g.NT) { case 1: LAUNCH_SS2(1) break; case 2: LAUNCH_SS2(2) break; case 3: LAUNCH_SS2(3) break; default: LAUNCH_SS2(4) break; }
#undef LAUNCH_SS2
#undef LAUNCH_SS
}

}  // namespace ib
