// Elastic-net coordinate descent for the dense (tuning = 0) column update: every gene shares XtX = U'U.
//
//   replaces  optimize_col() tuning = 0 branch   src/optimize.cpp:232-248
//             strong_coordinate_descent()        src/coordinate_descent.cpp:57-127
//
// Mapping: ONE GENE PER THREAD, the whole solver state (q = X'y - X'X beta and beta, 2 x KT doubles) in registers. The
// visiting order of a sweep depends only on (seed, ALS iteration, sweep index) - common.cuh:perm_key - so all 32 genes of a
// warp start at sweep 0 together and visit the SAME coordinate at every step: the coordinate is warp-uniform. Per sweep
// the warp
//   * relabels q and beta into visiting order (thread-private transposition through shared memory), so that step i of the
//     fully unrolled sweep touches registers q[i], b[i]: no dynamic register index, no dispatch, one basic block per sweep;
//   * reads row i of a copy of XtX permuted into the same order, which arrives for the NEXT sweep while this one runs:
//     fetched with one TMA bulk copy from the pre-permuted set of all 4096 orders (iterations with many sweeps), or built
//     in place from the coordinate-order table (steady state: a few sweeps per gene do not pay for 22 MB of tables).
// What bounds it (ncu, profiles/r01_ncu_k_cd_dense_*.txt): every step needs the 24 doubles of the table row in every
// thread, and a broadcast LDS costs one shared-memory wavefront per double: the shared-memory data pipe runs at 92-93 % of
// its peak with the FP64 pipe at ~40 %. Measured alternatives: XtX through the constant bank (ptxas emits LDCU.128 into two
// uniform-register quads and serialises on them: 1.6x slower); two genes per thread sharing every row load (half the
// wavefronts per gene, but half the warps at the same cycles-per-instruction: 1.15x slower); `switch (k)` on coordinate-
// order registers instead of relabelling (same wavefronts, plus dispatch: 1.05x slower).
// Also measured and dropped: a gene split over 4 lanes (lone-warp sweep 4.1 us against 2.45 us: the shuffle sits on the
// dependency chain), static per-SM chunks of the ordered gene list instead of the hardware block scheduler (much slower:
// the long genes end up on few SMs), the 8-lanes-per-gene kernel of k_cd.cu on this path (2x slower at every shard size,
// profiles/r01_cd_dense_vs_group_shard_sizes.txt).
// A warp runs until its slowest gene has converged (finished genes idle); genes are handed out in the order of their
// previous sweep counts (`order`, k_cd_order) so that a warp's genes finish together. The first ALS iterations, where a few
// genes need 100x the median number of sweeps, run in PHASES (draw0/cap/state below): a launch stops every gene at sweep
// `cap`, parks the unconverged ones (q, beta, active set, last |loss decrement|), k_cd_order compacts and re-sorts them
// and the next launch resumes them in full warps - bitwise the same iterates, only regrouped. Arithmetic per coordinate (covariance form,
// exact loss decrements, correctly rounded division) is identical to k_cd.cu - see the header comment there.
#include <algorithm>
#include <utility>

#include "common.cuh"
#include "kernels.cuh"

// State form (round 2). The solver keeps p_k = q_k + beta_k XtX_kk ("upper" of coordinate_descent.cpp:94) instead of
// q_k = X'y_k - (X'X beta)_k: an update of coordinate k leaves p_k unchanged (q_k -= delta XtX_kk, beta_k += delta) and moves
// p_l by -delta XtX_lk for l != k, so a step reads its upper straight from a register and the chain between two consecutive
// steps is  |p| - la -> copysign -> * 1/(XtX_kk + l2) -> two selects -> new - old -> FMA into the next p : 4 FP64 operations
// instead of 7 (the multiplication by the tabulated reciprocal replaces the correctly rounded division: <= 1 ulp in beta, far
// inside the 1e-8 parity tolerance; sweep counts stay identical to the residual-form oracle on every fixture). 31 instead of
// 39 FP64 instructions per coordinate update. For excluded coordinates beta = 0, so p_e = q_e and the KKT test is unchanged.
// -DCD_QFORM=1 builds the round-1 arithmetic (q form, Markstein-corrected division) for A/B measurements.
#ifndef CD_QFORM
#define CD_QFORM 0
#endif

namespace ib {

namespace {

constexpr int DW = 1;                // warps per block: one-warp blocks spread evenly over the SMs
constexpr int MAX_SWEEPS_D = 200000;

struct CdDenseArgs {
    const double* XtX;               // K x K, element (r, c) at r*xs_r + c*xs_c
    int xs_r, xs_c;
    const double* table;             // [KT][KT + 4] prepared by k_cd_table (row r: XtX[r][:], XtX_rr, 1/(XtX_rr + l2), 0, 0)
    const double* tables_all;        // optional [PERM_T][KT][KT + 4]: the table permuted into every visiting order (k_cd_tables_all)
    const double* Xty; const double* W0; double* Vout;     // per gene, stride ldv (Vout may alias W0)
    int64_t ldv;
    int K; int64_t P;
    double lambda, alpha;
    double la, l2;                   // lambda*alpha, lambda*(1-alpha) (set by launch())
    const double* tol_dev; double tol_host;
    const uint32_t* als_iter_dev; uint32_t als_iter_host;
    uint64_t seed; int perm_mode;
    unsigned long long* sweeps_total; unsigned long long* steps_total;
    int* sweeps_per_gene;            // optional [P]
    const int* order;                // optional [P]: slot i solves gene order[i]
    const unsigned char* perm_table;
    // phased execution (optional): a launch runs sweeps [draw0, cap) of every gene it is given and parks the unconverged ones
    const int* n_slots_dev;          // number of slots of this launch (device; nullptr: P)
    uint32_t draw0, cap;             // first sweep index of this launch (> 0: resume from `state`), sweep index to stop at
    double* state;                   // [P][2*KT] q | beta in coordinate order of the parked genes
    uint32_t* state_inc;             // [P] active set (coordinate space) of the parked genes
    float* state_dl;                 // [P] |loss decrement| of a parked gene's last sweep (orders the next phase)
    int* alive;                      // [P] 1: parked (unconverged at `cap`), 0: finished
};

// byte j of a packed byte array held in 32-bit words (compile-time j)
template <int NW>
__device__ __forceinline__ uint32_t byte_of(const uint32_t (&w)[NW], int j) { return (w[j >> 2] >> (8 * (j & 3))) & 0xffu; }

// One coordinate update at POSITION I of the sweep (compile-time) for this thread's gene. q and b are held in visiting
// order (position layout), Xp is the warp's XtX table permuted into the same order:
//   row[0..KT)  XtX[k_I][k_l]   row[KT] XtX_kk   row[KT+1] 1/(XtX_kk + l2)   (row pitch KT + 4 keeps 16-byte alignment)
template <int KT, int I>
__device__ __forceinline__ void cd_step(double (&q)[KT], double (&b)[KT], const double* __restrict__ Xp, uint32_t incp, double la, double l2, double& dl) {
    const double* row = Xp + I * (KT + 4);                                   // compile-time shared-memory offset, warp-uniform
    const bool on = (incp >> I) & 1u;
    const double bo = b[I];
#if CD_QFORM
    // one broadcast load for the row constants (a broadcast LDS costs a wavefront per double and the kernel is bound by them):
    // d and 1/(d + l2) come from the table, d + l2 and (d + l2)/2 are recomputed (same roundings as the table's)
    const double2 dr = *reinterpret_cast<const double2*>(row + KT);          // d, 1/den
    const double den = dr.x + l2, hden = 0.5 * den;
    const double up = fma(bo, dr.x, q[I]);                                   // coordinate_descent.cpp:94
    const double t1 = fabs(up) - la;
    const double num = copysign(t1, up);
    double nb = num * dr.y;                                                  // :99-104, correctly rounded num / den
    nb = fma(fma(-den, nb, num), dr.y, nb);
#else
    const double2 dr = *reinterpret_cast<const double2*>(row + KT);          // (XtX_kk + l2) / 2, 1 / (XtX_kk + l2)
    const double hden = dr.x;
    const double up = q[I];                                                  // p form: the state IS the upper of :94
    const double t1 = fabs(up) - la;
    double nb = copysign(t1, up) * dr.y;                                     // :99-104
    (void)l2;
#endif
    nb = (__double2hiint(t1) >= 0) ? nb : 0.0;                               // t1 > 0 (t1 == +0 gives nb == 0 either way); integer test: one FP64-pipe op less
    nb = on ? nb : bo;                                                       // excluded coordinate / finished gene: no-op
    const double dlt = nb - bo;
    // exact loss decrement of this update: dlt ((XtX_kk + l2)(new + old)/2 - upper) + lambda alpha (|new| - |old|)
    dl = fma(dlt, fma(hden, nb + bo, -up), dl);
    dl = fma(la, fabs(nb) - fabs(bo), dl);
    b[I] = nb;                                                               // :106-109
    const double nd = -dlt;
#pragma unroll
    for (int l = 0; l < KT; l += 2) {
        const double2 x = *reinterpret_cast<const double2*>(row + l);
#if CD_QFORM
        q[l] = fma(nd, x.x, q[l]);
        q[l + 1] = fma(nd, x.y, q[l + 1]);
#else
        if (l != I) q[l] = fma(nd, x.x, q[l]);                               // p_k itself does not move
        if (l + 1 != I) q[l + 1] = fma(nd, x.y, q[l + 1]);
#endif
    }
}

// row I of a table in visiting order `ow` (order bytes): Xn[I][l] = XtX[k_I][k_l], per-row constants copied from row k_I.
// Positions >= K are the identity and hit zero rows of Xs. Lane l writes column l (and l + 32 when the row is longer).
template <int KT, int I>
__device__ __forceinline__ void build_row(double* __restrict__ Xn, const double* __restrict__ Xs, const uint32_t (&ow)[KT / 4], int srcA, int lane) {
    constexpr int XLD = KT + 4;
    const double* srow = Xs + byte_of(ow, I) * XLD;
    if (lane < XLD) Xn[I * XLD + lane] = srow[srcA];
    if (XLD > 32 && lane + 32 < XLD) Xn[I * XLD + lane + 32] = srow[lane + 32];
}
template <int KT, int... Is>
__device__ __forceinline__ void build_table(std::integer_sequence<int, Is...>, double* __restrict__ Xn, const double* __restrict__ Xs,
                                            const uint32_t (&ow)[KT / 4], int srcA, int lane) {
    (build_row<KT, Is>(Xn, Xs, ow, srcA, lane), ...);
}
// One sweep on table Xp, straight-line over all KT positions (the KT - K <= 3 padding positions are never active and their
// table rows are zero: exact no-ops, and no branch keeps the sweep one basic block). The rows of the NEXT sweep's table are
// built in between, off the critical path.
template <int KT, int... Is>
__device__ __forceinline__ void cd_sweep(std::integer_sequence<int, Is...>, double (&q)[KT], double (&b)[KT], const double* __restrict__ Xp,
                                         uint32_t incp, double la, double l2, double& dl, double* __restrict__ Xn, const double* __restrict__ Xs,
                                         const uint32_t (&ow)[KT / 4], int srcA, int lane) {
    ((cd_step<KT, Is>(q, b, Xp, incp, la, l2, dl), build_row<KT, Is>(Xn, Xs, ow, srcA, lane)), ...);
}
// the same without the table build: the next table arrives by a TMA bulk copy from the pre-permuted set (no LSU wavefronts)
template <int KT, int... Is>
__device__ __forceinline__ void cd_sweep_nobuild(std::integer_sequence<int, Is...>, double (&q)[KT], double (&b)[KT], const double* __restrict__ Xp,
                                                 uint32_t incp, double la, double l2, double& dl) {
    (cd_step<KT, Is>(q, b, Xp, incp, la, l2, dl), ...);
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// tables_all[t] = the prepared table permuted into visiting order t of the K coordinates (order table of common.cuh):
// out[t][i][l] = table[k_i][k_l] for l < KT (positions >= K: identity), out[t][i][KT..KT+3] = table[k_i][KT..KT+3]
__global__ void __launch_bounds__(128) k_cd_tables_all(const double* __restrict__ table, const unsigned char* __restrict__ perm_table, int K, int KT,
                                                       double* __restrict__ out) {
    __shared__ unsigned char ord[32];
    const int XLD = KT + 4, t = blockIdx.x;
    if (threadIdx.x < 32) ord[threadIdx.x] = perm_table[PERM_TABLE_HALF + ((size_t)(K - 1) * PERM_T + t) * 32 + threadIdx.x];
    __syncthreads();
    double* o = out + (size_t)t * KT * XLD;
    for (int x = threadIdx.x; x < KT * XLD; x += blockDim.x) {
        const int i = x / XLD, l = x % XLD;
        const int ki = (i < K) ? ord[i] : i;
        const int kl = (l < KT) ? ((l < K) ? ord[l] : l) : l;
        o[x] = table[ki * XLD + kl];
    }
}


// The XtX table in coordinate order, ready to be copied into shared memory by every block: row r = XtX[r][0..KT) (zero padded),
// XtX_rr, 1/(XtX_rr + l2), 0, 0. One small launch per column update instead of 21 dependent global loads per thread in each of
// the 1390 one-warp blocks.
__global__ void __launch_bounds__(256) k_cd_table(const double* __restrict__ XtX, int xs_r, int xs_c, int K, int KT, double l2, double* __restrict__ out) {
    const int XLD = KT + 4;
    for (int x = threadIdx.x; x < KT * XLD; x += blockDim.x) {
        const int r = x / XLD, c = x % XLD;
        double v = 0.0;
        if (r < K) {
            if (c < KT) v = (c < K) ? XtX[(size_t)r * xs_r + (size_t)c * xs_c] : 0.0;
#if CD_QFORM
            else if (c == KT) v = XtX[(size_t)r * xs_r + (size_t)r * xs_c];
#else
            else if (c == KT) v = 0.5 * (XtX[(size_t)r * xs_r + (size_t)r * xs_c] + l2);
#endif
            else if (c == KT + 1) v = 1.0 / (XtX[(size_t)r * xs_r + (size_t)r * xs_c] + l2);
        }
        out[x] = v;
    }
}

// MINB one-warp blocks per SM. MINB = 8 (255 registers): the lone-warp sweep takes 2.3 us against 2.8 us at 200 and 3.6 us at
// 168 registers (ptxas keeps more table rows in flight) - the regime of the long early iterations, whose launches end with
// the latency of their longest genes, and of the small shards of a multi-GPU run. MINB = 10 (200 registers): all 1390 blocks
// of the 44 477-gene problem are resident at once; at steady state (4 sweeps per gene) a second wave of blocks costs more
// than the slower sweep. lib.cu switches after the first iterations.
template <int KT, bool TMA>
__device__ __forceinline__ void cd_dense_body(const CdDenseArgs& a) {
    constexpr int XLD = KT + 4;
    constexpr int NW = KT / 4;                         // 32-bit words holding KT position bytes
    constexpr int WBUF = (KT * 32 > KT * XLD) ? KT * 32 : KT * XLD;          // one warp buffer: permuted table, then relabel scratch
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double* Xs = reinterpret_cast<double*>(smem_raw);                        // block: XtX table in coordinate order
    double* Xw = Xs + KT * XLD + warp * 2 * WBUF;                            // warp: two buffers (current / next sweep)
    unsigned char* bytes = reinterpret_cast<unsigned char*>(Xs + KT * XLD + DW * 2 * WBUF) + warp * 144;
    unsigned char* ord_s = bytes;                                            // [2][32] visiting order of the current / next sweep
    unsigned char* rank_s = bytes + 64;                                      // [32] rank of every coordinate in the next order
    unsigned char* np_s = bytes + 96;                                        // [32] next position of the value at current position j
    uint64_t* bars = reinterpret_cast<uint64_t*>(bytes + 128);               // [2] arrival of the two table buffers (TMA)
    const int K = a.K;
    const double la = a.la, l2 = a.l2;
    const double tol = a.tol_dev ? *a.tol_dev : a.tol_host;
    const uint32_t als_iter = a.als_iter_dev ? *a.als_iter_dev : a.als_iter_host;
    const uint64_t key_iter = mix64(a.seed + 0x9E3779B97F4A7C15ull * (1ull + als_iter));   // perm_key(): first factor
    constexpr bool use_tma = TMA;                                             // (launch() picks TMA only with tables_all and perm_mode 1)

    const int64_t n_slots = a.n_slots_dev ? (int64_t)*a.n_slots_dev : a.P;
    if ((int64_t)blockIdx.x * (DW * 32) >= n_slots) return;                  // a later phase usually has far fewer genes than blocks
    {   // prepared table -> shared memory, 16 bytes per load, all loads of a thread independent
        const double2* src = reinterpret_cast<const double2*>(a.table);
        double2* dst = reinterpret_cast<double2*>(Xs);
        constexpr int N2 = KT * XLD / 2;
#pragma unroll
        for (int x0 = 0; x0 < N2; x0 += DW * 32 * 4) {
            double2 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int x = x0 + u * DW * 32 + tid; if (x < N2) v[u] = __ldg(src + x); }
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int x = x0 + u * DW * 32 + tid; if (x < N2) dst[x] = v[u]; }
        }
    }
    ord_s[lane] = (unsigned char)lane;                                       // order before the first sweep: identity (positions = coordinates)
    if (use_tma && lane == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
    __syncthreads();

    // ---- this thread's gene; positions = coordinates until the first sweep relabels them
    const int64_t slot = (int64_t)blockIdx.x * (DW * 32) + tid;
    bool active = slot < n_slots;
    const int64_t gene = active ? (a.order ? (int64_t)a.order[slot] : slot) : 0;
    double q[KT], b[KT];
    uint32_t incp = 0;                                                       // active set, indexed by POSITION
    if (a.draw0 == 0u) {
        const double* xp = a.Xty + gene * a.ldv;
        const double* wp = a.W0 + gene * a.ldv;
        double mx = 0.0;
#pragma unroll
        for (int c = 0; c < KT; ++c) {
            q[c] = (active && c < K) ? xp[c] : 0.0;
            b[c] = (active && c < K) ? wp[c] : 0.0;
            mx = fmax(mx, fabs(q[c]));
        }
        const double thr = a.alpha * (2.0 * a.lambda - mx);                  // coordinate_descent.cpp:74
#pragma unroll
        for (int c = 0; c < KT; ++c) {
            const bool on = active && (c < K) && !(fabs(q[c]) < thr);
            if (!on) b[c] = 0.0;                                             // :75-78
            incp |= (on ? 1u : 0u) << c;
        }
        // q = X'y - X'X beta   (:79, in covariance form); coordinates in ascending order like k_cd.cu
#pragma unroll
        for (int m = 0; m < KT; ++m) {
            const double bm = b[m];
#pragma unroll
            for (int l = 0; l < KT; l += 2) {
                const double2 x = *reinterpret_cast<const double2*>(&Xs[m * XLD + l]);
#if CD_QFORM
                q[l] = fma(-x.x, bm, q[l]);
                q[l + 1] = fma(-x.y, bm, q[l + 1]);
#else
                if (l != m) q[l] = fma(-x.x, bm, q[l]);                      // p = X'y - (X'X - diag) beta
                if (l + 1 != m) q[l + 1] = fma(-x.y, bm, q[l + 1]);
#endif
            }
        }
    } else {                                                                 // resume a gene parked by the previous phase
        const double* sp = a.state + (size_t)gene * (2 * KT);
#pragma unroll
        for (int c = 0; c < KT; ++c) { q[c] = active ? sp[c] : 0.0; b[c] = active ? sp[KT + c] : 0.0; }
        incp = active ? a.state_inc[gene] : 0u;
    }
    const uint32_t full = (K >= 32) ? 0xffffffffu : ((1u << K) - 1u);
    int sweeps = (int)a.draw0;                                               // every gene of a launch is at the same sweep index
    unsigned long long steps_acc = 0;

    // table rows of sweep `dr`: lanes 0-7 fetch the 8 words of the order row (coordinate at every position), lanes 8-15 the
    // rank row (position of every coordinate); identity when perm_mode != 1
    auto row_word = [&](uint32_t dr) -> uint32_t {
        const uint32_t c0 = 4u * (lane & 7);
        if (a.perm_mode != 1) return c0 | ((c0 + 1) << 8) | ((c0 + 2) << 16) | ((c0 + 3) << 24);
        const uint64_t pk = key_iter ^ mix64((uint64_t)dr * 0x8CB92BA72F3D8DD7ull + 0x2545F4914F6CDD1Dull);
        const unsigned char* rowp = a.perm_table + ((lane & 8) ? 0 : PERM_TABLE_HALF) + ((size_t)(K - 1) * PERM_T + perm_select(pk)) * 32;
        return __ldg(reinterpret_cast<const uint32_t*>(rowp) + (lane & 7));
    };
    // relabels this gene's state into the next visiting order (np_s: next position of the value at current position j) through
    // the thread-private column of `scratch` (a table buffer that is dead at that point)
    auto relabel = [&](double* scratch) {
        uint32_t npw[NW];
#pragma unroll
        for (int w = 0; w < NW; ++w) npw[w] = reinterpret_cast<const uint32_t*>(np_s)[w];
        double* col = scratch + lane;
#pragma unroll
        for (int j = 0; j < KT; ++j) col[byte_of(npw, j) * 32] = q[j];
#pragma unroll
        for (int i = 0; i < KT; ++i) q[i] = col[i * 32];
#pragma unroll
        for (int j = 0; j < KT; ++j) col[byte_of(npw, j) * 32] = b[j];
#pragma unroll
        for (int i = 0; i < KT; ++i) b[i] = col[i * 32];
        if (__any_sync(FULL, incp != full && incp != 0u)) {                  // screened coordinates: rare
            uint32_t n = 0;
#pragma unroll
            for (int j = 0; j < KT; ++j) n |= ((incp >> j) & 1u) << byte_of(npw, j);
            incp = n;
        }
    };
    // publishes the fetched table rows as the order `into` (0/1) and np_s = positions in it of the values now ordered by `from`
    auto publish_next = [&](uint32_t row_w, int into, int from) {
        if (lane < 8) reinterpret_cast<uint32_t*>(ord_s + 32 * into)[lane] = row_w;
        else if (lane < 16) reinterpret_cast<uint32_t*>(rank_s)[lane - 8] = row_w;
        __syncwarp();
        np_s[lane] = (lane < K) ? rank_s[ord_s[32 * from + lane]] : (unsigned char)lane;
        __syncwarp();
    };
    // ---- prologue: order of sweep 0, state relabelled into it, its table built; order words of sweep 1 in flight
    uint32_t draw = a.draw0;
    publish_next(row_word(draw), 1, 0);
    uint32_t row_w = row_word(draw + 1);
    relabel(Xw);
    __syncwarp();
    int cur = 1;                                                             // ord_s[cur]: order of the sweep about to run; table in Xw + cur*WBUF
    {
        uint32_t ow[NW];
#pragma unroll
        for (int w = 0; w < NW; ++w) ow[w] = reinterpret_cast<const uint32_t*>(ord_s + 32 * cur)[w];
        const int srcA = (lane < K) ? (int)ord_s[32 * cur + lane] : lane;
        build_table<KT>(std::make_integer_sequence<int, KT>{}, Xw + cur * WBUF, Xs, ow, srcA, lane);
    }
    uint32_t tma_count[2] = {0u, 0u};                                        // fetches issued into each buffer (mbarrier phase parity)
    while (true) {
        // order of the next sweep (prefetched words) and the relabelling into it; then its table is built during this sweep
        publish_next(row_w, cur ^ 1, cur);                                   // (its __syncwarp also orders the table build / relabel of the last sweep)
        row_w = row_word(draw + 2);                                          // consumed by the sweep after the next one
        const double* Xp = Xw + cur * WBUF;
        double* Xn = Xw + (cur ^ 1) * WBUF;
        double dl = 0.0;
        if constexpr (use_tma) {
            // the next sweep's table comes from the pre-permuted set: one bulk copy, issued now, awaited after this sweep. The
            // buffer was the relabel scratch of the last sweep (generic-proxy accesses, ordered by the __syncwarp above).
            if (lane == 0) {
                const uint64_t pk = key_iter ^ mix64((uint64_t)(draw + 1) * 0x8CB92BA72F3D8DD7ull + 0x2545F4914F6CDD1Dull);
                fence_proxy_async_smem();
                mbar_expect_tx(&bars[cur ^ 1], (uint32_t)(KT * XLD * 8));
                tma_load_1d(Xn, a.tables_all + (size_t)perm_select(pk) * (KT * XLD), (uint32_t)(KT * XLD * 8), &bars[cur ^ 1]);
            }
            ++tma_count[cur ^ 1];
            cd_sweep_nobuild<KT>(std::make_integer_sequence<int, KT>{}, q, b, Xp, incp, la, l2, dl);
        } else {
            uint32_t ow[NW];
#pragma unroll
            for (int w = 0; w < NW; ++w) ow[w] = reinterpret_cast<const uint32_t*>(ord_s + 32 * (cur ^ 1))[w];
            const int srcA = (lane < K) ? (int)ord_s[32 * (cur ^ 1) + lane] : lane;
            cd_sweep<KT>(std::make_integer_sequence<int, KT>{}, q, b, Xp, incp, la, l2, dl, Xn, Xs, ow, srcA, lane);
        }
        // ---- end of the sweep for this gene: inner do-while test (:114), KKT re-admission (:118-124)
        if (active) {
            ++sweeps;
            steps_acc += (unsigned long long)__popc(incp);
            if (!(fabs(dl) > tol) || sweeps >= MAX_SWEEPS_D) {
                uint32_t vmask = 0;
#pragma unroll
                for (int i = 0; i < KT; ++i)
                    if (i < K && !((incp >> i) & 1u) && fabs(q[i]) > la) vmask |= 1u << i;    // |XtX[e,inc] beta - Xty_e| = |q_e| (beta_e = 0)
                if (vmask == 0u || sweeps >= MAX_SWEEPS_D) {
                    double* vp = a.Vout + gene * a.ldv;
                    const unsigned char* oc = ord_s + 32 * cur;
#pragma unroll
                    for (int i = 0; i < KT; ++i) if (i < K) vp[oc[i]] = b[i];
                    if (a.sweeps_per_gene) a.sweeps_per_gene[gene] = sweeps;
                    if (a.alive) a.alive[gene] = 0;
                    active = false; incp = 0;
                } else incp |= vmask;
            }
            if (active && (uint32_t)sweeps >= a.cap) {
                // park: state back in coordinate order (position i holds coordinate oc[i]); the next phase resumes at sweep `cap`
                double* sp = a.state + (size_t)gene * (2 * KT);
                const unsigned char* oc = ord_s + 32 * cur;
                uint32_t inc = 0;
#pragma unroll
                for (int i = 0; i < KT; ++i) {
                    const int c = (i < K) ? (int)oc[i] : i;
                    sp[c] = q[i]; sp[KT + c] = b[i];
                    inc |= ((incp >> i) & 1u) << c;
                }
                a.state_inc[gene] = inc;
                a.state_dl[gene] = (float)fabs(dl);
                a.alive[gene] = 1;
                active = false; incp = 0;
            }
        }
        if (!__any_sync(FULL, active)) break;
        __syncwarp();                                                        // every lane is done reading this sweep's table
        relabel(Xw + cur * WBUF);                                            // ... which now serves as the relabel scratch
        ++draw; cur ^= 1;
        if (use_tma) mbar_wait(&bars[cur], (tma_count[cur] - 1u) & 1u);     // the table of the sweep about to run has landed
    }
    // (a fetch may still be in flight when the warp leaves the loop: wait for it before the block's shared memory is released)
    if (use_tma && tma_count[cur ^ 1] > 0u) mbar_wait(&bars[cur ^ 1], (tma_count[cur ^ 1] - 1u) & 1u);
    // one atomic per warp for the statistics
    unsigned long long sw = (slot < n_slots) ? (unsigned long long)(sweeps - (int)a.draw0) : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sw += __shfl_xor_sync(FULL, sw, o); steps_acc += __shfl_xor_sync(FULL, steps_acc, o); }
    if (lane == 0) {
        if (a.sweeps_total && sw) atomicAdd(a.sweeps_total, sw);
        if (a.steps_total && steps_acc) atomicAdd(a.steps_total, steps_acc);
    }
}

template <int KT, bool TMA> __global__ void __launch_bounds__(DW * 32, 8) k_cd_dense(CdDenseArgs a) { cd_dense_body<KT, TMA>(a); }
// (__launch_bounds__(32, 10) makes ptxas stop at 168 registers; __maxnreg__ lets it use the 200 that 10 blocks per SM allow)
template <int KT> __global__ void __maxnreg__(200) k_cd_dense_r200(CdDenseArgs a) { cd_dense_body<KT, false>(a); }

// Slot order for the next launch: genes sorted by descending sweep count of the previous iteration (bucketed to ~3 %:
// exponent + 5 mantissa bits), so that the 32 genes of a warp finish together and the longest warps start first.
// Consecutive iterations correlate at 0.95+ (tools/gpu_sweep_dist.py): lockstep efficiency 0.58 -> 0.85. The order within a
// bucket depends on atomics; results do not depend on the order at all (every gene is solved independently).
constexpr int ORDER_BUCKETS = 32 * 19;
__device__ __forceinline__ int order_key(int s) {
    const uint32_t v = (uint32_t)max(s, 0) + 1u;
    const int e = 31 - __clz(v);
    const uint32_t m = ((v << (31 - e)) >> 26) & 31u;
    return ORDER_BUCKETS - 1 - min(ORDER_BUCKETS - 1, e * 32 + (int)m);     // descending
}
constexpr int ORDER_BLOCKS = 32, ORDER_THREADS = 256;
#ifndef ORDER_EVERY_MASK
#define ORDER_EVERY_MASK 7        // re-sort every 8th iteration after the first 8 (-DORDER_EVERY_MASK=0 = every iteration: measured 3 % slower at steady state)
#endif
// adds 1 to counter[key] for every lane of the warp, one atomic per distinct key (sweep counts cluster on a few buckets);
// returns the value this lane's increment would have received. key < 0: lane does not take part.
__device__ __forceinline__ int warp_agg_inc(int* counter, int key) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned m = __match_any_sync(FULL, key);
    const int leader = __ffs(m) - 1;
    int base = 0;
    if ((int)lane == leader && key >= 0) base = atomicAdd(&counter[key], __popc(m));
    base = __shfl_sync(FULL, base, leader);
    return base + __popc(m & ((1u << lane) - 1u));
}
// work: [ORDER_BUCKETS] bucket totals | arrive counter | done counter  (all zero between launches)
// Two uses: (1) before a column update: all P genes, key = sweep count of the previous iteration; (2) between the phases of a
// column update (alive != nullptr): only the parked genes, key = |loss decrement| of their last sweep relative to tol (the
// larger, the more sweeps remain: rank correlation 0.8-0.9), and their number goes to *n_slots.
__device__ __forceinline__ int order_key_dl(float dl, float tol) {
    const int b = (int)(__log2f(fmaxf(dl, tol) / tol) * 8.0f);               // 8 buckets per octave above tol
    return ORDER_BUCKETS - 1 - min(ORDER_BUCKETS - 1, max(0, b));            // descending
}
__global__ void __launch_bounds__(ORDER_THREADS) k_cd_order(const int* __restrict__ sweeps, int P, int* __restrict__ order, int* __restrict__ work,
                                                            const uint32_t* __restrict__ als_iter, const int* __restrict__ alive,
                                                            const float* __restrict__ state_dl, const double* __restrict__ tol_dev, int* __restrict__ n_slots) {
    // sweep counts of consecutive iterations correlate at 0.95+: after the first iterations a new order every 8th is enough
    if (als_iter) { const uint32_t it = *als_iter; if (it >= 8u && (it & (uint32_t)ORDER_EVERY_MASK) != 0u) return; }
    const float tolf = (alive && tol_dev) ? (float)*tol_dev : 1e-5f;
    auto key_of = [&](int j) -> int { return alive ? (alive[j] ? order_key_dl(state_dl[j], tolf) : -1) : order_key(sweeps[j]); };
    __shared__ int hist[ORDER_BUCKETS];      // this block's count per bucket, later its scatter cursor
    __shared__ int boff[ORDER_BUCKETS];      // start of this block's genes inside the bucket, later + start of the bucket
    __shared__ int last;
    int* gtot = work; int* arrive = work + ORDER_BUCKETS; int* done = arrive + 1;
    const int tid = threadIdx.x, gtid = blockIdx.x * ORDER_THREADS + tid, gsz = gridDim.x * ORDER_THREADS;
    const int Pw = (P + 31) / 32 * 32;
    for (int x = tid; x < ORDER_BUCKETS; x += ORDER_THREADS) hist[x] = 0;
    __syncthreads();
    for (int j = gtid; j < Pw; j += gsz) warp_agg_inc(hist, j < P ? key_of(j) : -1);
    __syncthreads();
    for (int x = tid; x < ORDER_BUCKETS; x += ORDER_THREADS) { boff[x] = hist[x] ? atomicAdd(&gtot[x], hist[x]) : 0; hist[x] = 0; }
    // grid barrier (all ORDER_BLOCKS blocks are co-resident: far fewer than SMs)
    __threadfence();
    __syncthreads();
    if (tid == 0) { atomicAdd(arrive, 1); while (atomicAdd(arrive, 0) < (int)gridDim.x) { } }
    __syncthreads();
    if (tid < 32) {                                                          // exclusive scan of the totals, one warp, 19 buckets per lane
        constexpr int PER = ORDER_BUCKETS / 32;
        int loc[PER], sum = 0;
#pragma unroll
        for (int u = 0; u < PER; ++u) { loc[u] = sum; sum += __ldcg(&gtot[tid * PER + u]); }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, incl, o); if (tid >= o) incl += v; }
        const int base = incl - sum;
#pragma unroll
        for (int u = 0; u < PER; ++u) boff[tid * PER + u] += base + loc[u];
        if (n_slots && blockIdx.x == 0 && tid == 31) *n_slots = incl;        // genes that take part
    }
    __syncthreads();
    for (int j = gtid; j < Pw; j += gsz) {
        const int k = j < P ? key_of(j) : -1;
        const int r = warp_agg_inc(hist, k);
        if (k >= 0) order[boff[k] + r] = j;
    }
    // the last block to finish zeroes the work area for the next launch
    __threadfence();
    __syncthreads();
    if (tid == 0) last = (atomicAdd(done, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (last) {
        for (int x = tid; x < ORDER_BUCKETS + 2; x += ORDER_THREADS) work[x] = 0;
    }
}

template <int KT, int MINB>
void launch_kt(const CdDenseArgs& a, cudaStream_t st) {
    constexpr int XLD = KT + 4;
    constexpr int WBUF = (KT * 32 > KT * XLD) ? KT * 32 : KT * XLD;
    const int blocks = (int)((a.P + DW * 32 - 1) / (DW * 32));
    const size_t smem = (size_t)(KT * XLD + DW * 2 * WBUF) * 8 + DW * 144;
    const bool tma = a.tables_all != nullptr && a.perm_mode == 1;
    auto kern = (MINB == 8) ? (tma ? k_cd_dense<KT, true> : k_cd_dense<KT, false>) : k_cd_dense_r200<KT>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<blocks, DW * 32, smem, st>>>(a);
}

// `table`: prepared by launch_cd_dense_table(); resident_all: prefer the variant whose blocks all fit at once
void launch(CdDenseArgs a, const double* table, bool resident_all, cudaStream_t st) {
    a.la = a.lambda * a.alpha; a.l2 = a.lambda * (1.0 - a.alpha);
    if (a.P <= 0) return;
    const int KT = (a.K + 3) / 4 * 4;
    a.table = table;
#define LAUNCH_KT(KTv) if (resident_all) launch_kt<KTv, 10>(a, st); else launch_kt<KTv, 8>(a, st); break;
    switch (KT / 4) {
        case 1: LAUNCH_KT(4)
        case 2: LAUNCH_KT(8)
        case 3: LAUNCH_KT(12)
        case 4: LAUNCH_KT(16)
        case 5: LAUNCH_KT(20)
        case 6: LAUNCH_KT(24)
        case 7: LAUNCH_KT(28)
        default: LAUNCH_KT(32)
    }
#undef LAUNCH_KT
}

}  // namespace

size_t cd_dense_table_elems() { return (size_t)32 * 36; }
void launch_cd_dense_table(int K, const double* XtX, int xs_r, int xs_c, double lambda, double alpha, double* table, cudaStream_t st) {
    k_cd_table<<<1, 256, 0, st>>>(XtX, xs_r, xs_c, K, (K + 3) / 4 * 4, lambda * (1.0 - alpha), table);
}
size_t cd_dense_tables_all_elems(int K) { const int KT = (K + 3) / 4 * 4; return (size_t)PERM_T * KT * (KT + 4); }
void launch_cd_dense_tables_all(int K, const double* table, const unsigned char* perm_table, double* tables_all, cudaStream_t st) {
    k_cd_tables_all<<<PERM_T, 128, 0, st>>>(table, perm_table, K, (K + 3) / 4 * 4, tables_all);
}

void launch_cd_dense(const Geom& g, const double* UtU, const double* Xty, double* V, const CdParams& p, unsigned long long* sweeps,
                     unsigned long long* steps, int* sweeps_per_gene, const int* order, const unsigned char* perm_table, const double* table,
                     bool resident_all, const CdPhaseState* ps, uint32_t draw0, uint32_t cap, const double* tables_all, cudaStream_t st) {
    CdDenseArgs a{};
    a.XtX = UtU; a.xs_r = g.KP; a.xs_c = 1;
    a.Xty = Xty; a.W0 = V; a.Vout = V; a.ldv = g.ldV; a.K = g.K; a.P = g.P;
    a.lambda = p.lambda; a.alpha = p.alpha; a.tol_dev = p.tol; a.als_iter_dev = p.als_iter; a.seed = p.seed; a.perm_mode = p.perm_mode;
    a.sweeps_total = sweeps; a.steps_total = steps; a.sweeps_per_gene = sweeps_per_gene; a.order = order; a.perm_table = perm_table;
    a.draw0 = draw0; a.cap = cap; a.tables_all = tables_all;
    if (ps) { a.state = ps->state; a.state_inc = ps->inc; a.state_dl = ps->dl; a.alive = ps->alive; a.n_slots_dev = draw0 ? ps->n_slots : nullptr; }
    launch(a, table, resident_all, st);
}

size_t cd_order_work_ints() { return ORDER_BUCKETS + 2; }
void launch_cd_order(const int* sweeps_per_gene, int64_t P, int* order, int* work, const uint32_t* als_iter, cudaStream_t st) {
    if (P > 0) k_cd_order<<<ORDER_BLOCKS, ORDER_THREADS, 0, st>>>(sweeps_per_gene, (int)P, order, work, als_iter, nullptr, nullptr, nullptr, nullptr);
}
void launch_cd_order_parked(const CdPhaseState& ps, int64_t P, int* order, int* work, const double* tol_dev, cudaStream_t st) {
    if (P > 0) k_cd_order<<<ORDER_BLOCKS, ORDER_THREADS, 0, st>>>(nullptr, (int)P, order, work, nullptr, ps.alive, ps.dl, tol_dev, ps.n_slots);
}

void launch_cd_dense_batch(int K, int64_t n, const double* XtX, const double* Xty, const double* w0, double lambda, double alpha, double tol,
                           int perm_mode, uint64_t seed, uint32_t als_iter, double* beta, int* sweeps, const unsigned char* perm_table,
                           double* table, cudaStream_t st) {
    CdDenseArgs a{};
    a.XtX = XtX; a.xs_r = 1; a.xs_c = K;                                     // caller's column-major K x K
    a.Xty = Xty; a.W0 = w0; a.Vout = beta; a.ldv = K; a.K = K; a.P = n;
    a.lambda = lambda; a.alpha = alpha; a.tol_dev = nullptr; a.tol_host = tol; a.als_iter_dev = nullptr; a.als_iter_host = als_iter;
    a.seed = seed; a.perm_mode = perm_mode; a.sweeps_per_gene = sweeps; a.perm_table = perm_table;
    a.draw0 = 0u; a.cap = 0xffffffffu;
    launch_cd_dense_table(K, XtX, 1, K, lambda, alpha, table, st);
    launch(a, table, false, st);
}

}  // namespace ib
