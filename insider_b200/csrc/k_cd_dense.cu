// Elastic-net coordinate descent for the dense (tuning = 0) column update: every gene shares XtX = U'U.
//
//   replaces  optimize_col() tuning = 0 branch   src/optimize.cpp:232-248
//             strong_coordinate_descent()        src/coordinate_descent.cpp:57-127
//
// Mapping: ONE GENE PER THREAD, the whole solver state (q = X'y - X'X beta and beta, 2 x KT doubles) in registers.
// The visiting order of a sweep depends only on (seed, ALS iteration, sweep index) - common.cuh:perm_key - so all 32
// genes of a warp start at sweep 0 together and visit the SAME coordinate k at every step: k is warp-uniform. That makes
//   * the XtX row of the step a shared-memory broadcast (one LDS.128 wavefront serves 32 genes x 2 columns; the 8-lanes-
//     per-gene kernel in k_cd.cu spends ~3 wavefronts per gene-coordinate and is bound by the shared-memory pipe), and
//   * `switch (k)` a uniform branch into a block whose register indices are compile-time constants,
// so a coordinate update is K DFMAs + a 6-deep FP64 chain per thread and the kernel runs on the FP64 pipe.
// A warp runs until its slowest gene has converged (finished lanes idle); genes can be handed out in the order of their
// previous sweep counts (`order`) so that a warp's genes finish together. Arithmetic per coordinate (covariance form,
// exact loss decrements, correctly rounded division) is identical to k_cd.cu - see the header comment there.
#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"

namespace ib {

namespace {

constexpr int DW = 4;                // warps per block
constexpr int MAX_SWEEPS_D = 200000;

struct CdDenseArgs {
    const double* XtX;               // K x K, element (r, c) at r*xs_r + c*xs_c
    int xs_r, xs_c;
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
    const int* order;                // optional [P]: thread i solves gene order[i]
    const unsigned char* perm_table;
};

// volatile: the table is loop-invariant, and without it the compiler hoists all KT x (KT+4) loads out of the sweep loop
__device__ __forceinline__ double2 lds128(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

// one coordinate update of coordinate C (compile-time) for this thread's gene; row = shared-memory row C of the table
//   row[0..KT)  XtX[C][:]     row[KT] XtX_CC     row[KT+1] XtX_CC + l2     row[KT+2] 1/(XtX_CC + l2)   row[KT+3] (XtX_CC + l2)/2
template <int KT, int C>
__device__ __forceinline__ void cd_step(double (&q)[KT], double (&b)[KT], uint32_t xbase, bool on, double la, double& dl) {
    constexpr uint32_t row = (uint32_t)C * (KT + 4) * 8u;                     // compile-time shared-memory offset
    const double2 dd = lds128(xbase + row + KT * 8u);                        // d, den
    const double2 rr = lds128(xbase + row + (KT + 2) * 8u);                  // 1/den, den/2
    const double bo = b[C];
    const double up = fma(bo, dd.x, q[C]);                                   // coordinate_descent.cpp:94
    const double t1 = fabs(up) - la;
    const double num = copysign(t1, up);
    double nb = num * rr.x;                                                  // :99-104, correctly rounded num / den
    nb = fma(fma(-dd.y, nb, num), rr.x, nb);
    nb = (t1 > 0.0) ? nb : 0.0;
    nb = on ? nb : bo;                                                       // excluded coordinate / finished gene: no-op
    const double dlt = nb - bo;
    // exact loss decrement of this update: dlt ((XtX_CC + l2)(new + old)/2 - upper) + lambda alpha (|new| - |old|)
    dl = fma(dlt, fma(rr.y, nb + bo, -up), dl);
    dl = fma(la, fabs(nb) - fabs(bo), dl);
    b[C] = nb;                                                               // :106-109
    const double nd = -dlt;
#pragma unroll
    for (int l = 0; l < KT; l += 2) {
        const double2 x = lds128(xbase + row + l * 8u);
        q[l] = fma(nd, x.x, q[l]);
        q[l + 1] = fma(nd, x.y, q[l + 1]);
    }
}

template <int KT>
__global__ void __launch_bounds__(DW * 32) k_cd_dense(CdDenseArgs a) {
    constexpr int XLD = KT + 4;
    __shared__ __align__(16) double Xs[KT * XLD];
    __shared__ __align__(16) unsigned char ord_s[DW][2][32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = a.K;
    const double la = a.la, l2 = a.l2;
    const double tol = a.tol_dev ? *a.tol_dev : a.tol_host;
    const uint32_t als_iter = a.als_iter_dev ? *a.als_iter_dev : a.als_iter_host;
    const uint64_t key_iter = mix64(a.seed + 0x9E3779B97F4A7C15ull * (1ull + als_iter));   // perm_key(): first factor

    for (int x = tid; x < KT * XLD; x += DW * 32) {
        const int r = x / XLD, c = x % XLD;
        double v = 0.0;
        if (r < K) {
            if (c < KT) v = (c < K) ? a.XtX[(size_t)r * a.xs_r + (size_t)c * a.xs_c] : 0.0;
            else {
                const double d = a.XtX[(size_t)r * a.xs_r + (size_t)r * a.xs_c], den = d + l2;
                v = (c == KT) ? d : (c == KT + 1) ? den : (c == KT + 2) ? 1.0 / den : 0.5 * den;
            }
        } else if (c == KT + 2) v = 1.0;
        Xs[x] = v;
    }
    __syncthreads();

    // ---- this thread's gene
    const int64_t slot = (int64_t)blockIdx.x * (DW * 32) + tid;
    bool active = slot < a.P;
    const int64_t gene = active ? (a.order ? (int64_t)a.order[slot] : slot) : 0;
    double q[KT], b[KT];
    uint32_t inc = 0;
    {
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
            inc |= (on ? 1u : 0u) << c;
        }
        // q = X'y - X'X beta   (:79, in covariance form); coordinates in ascending order like k_cd.cu
#pragma unroll
        for (int m = 0; m < KT; ++m) {
            const double bm = b[m];
#pragma unroll
            for (int l = 0; l < KT; l += 2) {
                const double2 x = *reinterpret_cast<const double2*>(&Xs[m * XLD + l]);
                q[l] = fma(-x.x, bm, q[l]);
                q[l + 1] = fma(-x.y, bm, q[l + 1]);
            }
        }
    }
    int sweeps = 0;
    unsigned long long steps_acc = 0;
    const uint32_t xbase = smem_u32(Xs);

    // visiting order of sweep `draw`: order-table row of K coordinates (identity when perm_mode != 1)
    auto row_word = [&](uint32_t dr) -> uint32_t {
        if (a.perm_mode != 1) { const uint32_t c0 = 4u * (lane & 7); return c0 | ((c0 + 1) << 8) | ((c0 + 2) << 16) | ((c0 + 3) << 24); }
        const uint64_t pk = key_iter ^ mix64((uint64_t)dr * 0x8CB92BA72F3D8DD7ull + 0x2545F4914F6CDD1Dull);
        return __ldg(reinterpret_cast<const uint32_t*>(a.perm_table + PERM_TABLE_HALF + ((size_t)(K - 1) * PERM_T + perm_select(pk)) * 32) + (lane & 7));
    };
    uint32_t draw = 0;
    uint32_t row_w = row_word(0);
    int cur = 0;
    while (true) {
        if (lane < 8) reinterpret_cast<uint32_t*>(ord_s[warp][cur])[lane] = row_w;
        __syncwarp();
        row_w = row_word(draw + 1);                                          // prefetched, consumed by the next sweep
        const unsigned char* ord = ord_s[warp][cur];
        double dl = 0.0;
        int k = ord[0];
        for (int i = 0; i < K; ++i) {
            const int kn = ord[(i + 1 < K) ? i + 1 : i];
            const bool on = (inc >> k) & 1u;
            switch (k) {
#define CD_CASE(Cv) case Cv: if constexpr (Cv < KT) cd_step<KT, (Cv < KT ? Cv : 0)>(q, b, xbase, on, la, dl); break;
                CD_CASE(0) CD_CASE(1) CD_CASE(2) CD_CASE(3) CD_CASE(4) CD_CASE(5) CD_CASE(6) CD_CASE(7)
                CD_CASE(8) CD_CASE(9) CD_CASE(10) CD_CASE(11) CD_CASE(12) CD_CASE(13) CD_CASE(14) CD_CASE(15)
                CD_CASE(16) CD_CASE(17) CD_CASE(18) CD_CASE(19) CD_CASE(20) CD_CASE(21) CD_CASE(22) CD_CASE(23)
                CD_CASE(24) CD_CASE(25) CD_CASE(26) CD_CASE(27) CD_CASE(28) CD_CASE(29) CD_CASE(30) CD_CASE(31)
#undef CD_CASE
                default: break;
            }
            k = kn;
        }
        // ---- end of the sweep for this gene: inner do-while test (:114), KKT re-admission (:118-124)
        if (active) {
            ++sweeps;
            steps_acc += (unsigned long long)__popc(inc);
            if (!(fabs(dl) > tol) || sweeps >= MAX_SWEEPS_D) {
                uint32_t vmask = 0;
#pragma unroll
                for (int c = 0; c < KT; ++c)
                    if (c < K && !((inc >> c) & 1u) && fabs(q[c]) > la) vmask |= 1u << c;     // |XtX[e,inc] beta - Xty_e| = |q_e| (beta_e = 0)
                if (vmask == 0u || sweeps >= MAX_SWEEPS_D) {
                    double* vp = a.Vout + gene * a.ldv;
#pragma unroll
                    for (int c = 0; c < KT; ++c) if (c < K) vp[c] = b[c];
                    if (a.sweeps_per_gene) a.sweeps_per_gene[gene] = sweeps;
                    active = false; inc = 0;
                } else inc |= vmask;
            }
        }
        if (!__any_sync(FULL, active)) break;
        ++draw; cur ^= 1;
    }
    // one atomic per warp for the statistics
    unsigned long long sw = (unsigned long long)sweeps;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sw += __shfl_xor_sync(FULL, sw, o); steps_acc += __shfl_xor_sync(FULL, steps_acc, o); }
    if (lane == 0) {
        if (a.sweeps_total && sw) atomicAdd(a.sweeps_total, sw);
        if (a.steps_total && steps_acc) atomicAdd(a.steps_total, steps_acc);
    }
}

void launch(CdDenseArgs a, cudaStream_t st) {
    a.la = a.lambda * a.alpha; a.l2 = a.lambda * (1.0 - a.alpha);
    const int blocks = (int)((a.P + DW * 32 - 1) / (DW * 32));
    if (blocks == 0) return;
    switch ((a.K + 3) / 4) {
        case 1: k_cd_dense<4><<<blocks, DW * 32, 0, st>>>(a); break;
        case 2: k_cd_dense<8><<<blocks, DW * 32, 0, st>>>(a); break;
        case 3: k_cd_dense<12><<<blocks, DW * 32, 0, st>>>(a); break;
        case 4: k_cd_dense<16><<<blocks, DW * 32, 0, st>>>(a); break;
        case 5: k_cd_dense<20><<<blocks, DW * 32, 0, st>>>(a); break;
        case 6: k_cd_dense<24><<<blocks, DW * 32, 0, st>>>(a); break;
        case 7: k_cd_dense<28><<<blocks, DW * 32, 0, st>>>(a); break;
        default: k_cd_dense<32><<<blocks, DW * 32, 0, st>>>(a); break;
    }
}

}  // namespace

void launch_cd_dense(const Geom& g, const double* UtU, const double* Xty, double* V, const CdParams& p, unsigned long long* sweeps,
                     unsigned long long* steps, int* sweeps_per_gene, const int* order, const unsigned char* perm_table, cudaStream_t st) {
    CdDenseArgs a{};
    a.XtX = UtU; a.xs_r = g.KP; a.xs_c = 1;
    a.Xty = Xty; a.W0 = V; a.Vout = V; a.ldv = g.ldV; a.K = g.K; a.P = g.P;
    a.lambda = p.lambda; a.alpha = p.alpha; a.tol_dev = p.tol; a.als_iter_dev = p.als_iter; a.seed = p.seed; a.perm_mode = p.perm_mode;
    a.sweeps_total = sweeps; a.steps_total = steps; a.sweeps_per_gene = sweeps_per_gene; a.order = order; a.perm_table = perm_table;
    launch(a, st);
}

void launch_cd_dense_batch(int K, int64_t n, const double* XtX, const double* Xty, const double* w0, double lambda, double alpha, double tol,
                           int perm_mode, uint64_t seed, uint32_t als_iter, double* beta, int* sweeps, const unsigned char* perm_table,
                           cudaStream_t st) {
    CdDenseArgs a{};
    a.XtX = XtX; a.xs_r = 1; a.xs_c = K;                                     // caller's column-major K x K
    a.Xty = Xty; a.W0 = w0; a.Vout = beta; a.ldv = K; a.K = K; a.P = n;
    a.lambda = lambda; a.alpha = alpha; a.tol_dev = nullptr; a.tol_host = tol; a.als_iter_dev = nullptr; a.als_iter_host = als_iter;
    a.seed = seed; a.perm_mode = perm_mode; a.sweeps_per_gene = sweeps; a.perm_table = perm_table;
    launch(a, st);
}

}  // namespace ib
