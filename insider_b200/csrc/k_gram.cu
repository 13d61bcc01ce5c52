// Gram matrices and fixed-order reductions.
//
//   (gram = V V^T, src/optimize.cpp:332, is accumulated inside k_row_b: k_stream.cu)
//   k_row_comp_gram  per-row complement  sum_{j: m_ij=0} v_j v_j^T  src/optimize.cpp:163,170 (c_factor.cols(zero_idx) * trans(..))
//   k_reduce         deterministic sum of per-block partial buffers
#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"

namespace ib {

namespace {

// warp per (row, split): DMMA rank-4 updates over the genes whose train bit is 0
template <int NT>
__global__ void __launch_bounds__(256) k_row_comp_gram(const uint32_t* __restrict__ trR, const double* __restrict__ V, double* __restrict__ Dp,
                                                       int N, int KP, int ldV, int WPr, int64_t P, int n_splits) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wid >= (int64_t)N * n_splits) return;
    const int row = (int)(wid / n_splits), split = (int)(wid % n_splits);
    const int wq = WPr / n_splits, wr = WPr % n_splits;
    const int w_begin = split * wq + min(split, wr), w_end = w_begin + wq + (split < wr ? 1 : 0);
    double acc[NT][NT][2];
#pragma unroll
    for (int i = 0; i < NT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    int64_t genes[4] = {0, 0, 0, 0};
    int cnt = 0;
    auto flush = [&]() {
        const int64_t mine = (t == 0) ? genes[0] : (t == 1) ? genes[1] : (t == 2) ? genes[2] : genes[3];
        double f[NT];
#pragma unroll
        for (int n = 0; n < NT; ++n) f[n] = (t < cnt) ? __ldg(V + mine * ldV + 8 * n + g) : 0.0;
#pragma unroll
        for (int n1 = 0; n1 < NT; ++n1)
#pragma unroll
            for (int n2 = n1; n2 < NT; ++n2) dmma(acc[n1][n2][0], acc[n1][n2][1], f[n1], f[n2]);
        cnt = 0;
    };
    for (int w0 = w_begin; w0 < w_end; w0 += 32) {
        uint32_t z = 0;
        const int wi = w0 + lane;
        if (wi < w_end) {
            z = ~__ldg(trR + (size_t)row * WPr + wi);
            const int64_t lim = P - 32 * (int64_t)wi;       // genes >= P are padding, never counted
            if (lim <= 0) z = 0; else if (lim < 32) z &= (1u << lim) - 1u;
        }
        const int wn = min(32, w_end - w0);
        for (int w = 0; w < wn; ++w) {
            uint32_t zw = __shfl_sync(FULL, z, w);
            while (zw) {
                const int b = __ffs(zw) - 1;
                zw &= zw - 1;
                genes[cnt++] = 32 * (int64_t)(w0 + w) + b;
                if (cnt == 4) flush();
            }
        }
    }
    if (cnt > 0) flush();
    double* out = Dp + ((size_t)split * N + row) * KP * KP;
#pragma unroll
    for (int n1 = 0; n1 < NT; ++n1)
#pragma unroll
        for (int n2 = n1; n2 < NT; ++n2)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int ra = 8 * n1 + g, cb = 8 * n2 + 2 * t + e;
                out[ra * KP + cb] = acc[n1][n2][e];
                out[cb * KP + ra] = acc[n1][n2][e];
            }
}

// Fixed-order reduction of per-block partial buffers, up to three jobs per launch (B, G, D). A block owns 32 consecutive
// elements; its 8 warps each sum every 8th partial (the loads of one warp are coalesced and independent), then the 8
// warp sums are combined in ascending order: deterministic, and short dependency chains even for hundreds of partials.
struct ReduceJobs { double* out[3]; const double* parts[3]; long long n[3]; int n_parts[3]; int flat[3]; int first_block[4]; };
__global__ void __launch_bounds__(256) k_reduce(ReduceJobs jobs) {
    __shared__ double sm[8][33];
    int j = 0;
    while (j < 2 && (int)blockIdx.x >= jobs.first_block[j + 1]) ++j;
    const long long n = jobs.n[j];
    const int np = jobs.n_parts[j];
    if (jobs.flat[j]) {
        // few partials: one thread per element, 256 elements per block
        const long long i = ((long long)blockIdx.x - jobs.first_block[j]) * 256 + threadIdx.x;
        if (i >= n) return;
        const double* p = jobs.parts[j] + i;
        double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        int q = 0;
        for (; q + 4 <= np; q += 4) {
            const double v0 = p[(size_t)q * n], v1 = p[(size_t)(q + 1) * n], v2 = p[(size_t)(q + 2) * n], v3 = p[(size_t)(q + 3) * n];
            s0 += v0; s1 += v1; s2 += v2; s3 += v3;
        }
        for (; q < np; ++q) s0 += p[(size_t)q * n];
        jobs.out[j][i] = (s0 + s1) + (s2 + s3);
        return;
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long i = ((long long)blockIdx.x - jobs.first_block[j]) * 32 + lane;
    double s = 0.0;
    if (i < n) {
        const double* p = jobs.parts[j] + i;
        // eight independent loads in flight per thread (the V V^T job has ~600 partials on few blocks: its chain of L2 round
        // trips is the length of this kernel)
        double sa[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        int q = w;
        for (; q + 56 < np; q += 64) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = p[(size_t)(q + 8 * u) * n];
#pragma unroll
            for (int u = 0; u < 8; ++u) sa[u] += v[u];
        }
        for (; q < np; q += 8) sa[0] += p[(size_t)q * n];
        s = ((sa[0] + sa[1]) + (sa[2] + sa[3])) + ((sa[4] + sa[5]) + (sa[6] + sa[7]));
    }
    sm[w][lane] = s;
    __syncthreads();
    if (w == 0 && i < n) {
        double t = sm[0][lane];
#pragma unroll
        for (int k = 1; k < 8; ++k) t += sm[k][lane];
        jobs.out[j][i] = t;
    }
}

// G = V V^T (src/optimize.cpp:332) as its own kernel: GV_BLOCKS blocks of 8 warps, a warp takes every (8 GV_BLOCKS)-th group of
// 4 genes as one DMMA rank-4 update of the NT (NT + 1) / 2 upper tiles (operands straight from global memory); the 8 warps of a
// block are combined through shared memory, the GV_BLOCKS block partials by the last block to finish - every sum in a fixed
// order. Few blocks on purpose: the kernel is a latency chain (loads -> block sum -> cross-block sum) and the cross-block sum
// costs one L2 round trip per 8 partials (a first version with 148 partials summed serially took 84 us, this one ~8).
constexpr int GV_BLOCKS = 64, GV_WARPS = 4;
template <int NT>
__global__ void __launch_bounds__(GV_WARPS * 32) k_gram_v(const double* __restrict__ V, int ldV, int KP, int64_t P, double* __restrict__ parts, double* __restrict__ G,
                                                unsigned int* counter, CheckState* bump) {
    __shared__ double sm[GV_WARPS][NT * 8 * NT * 8];
    __shared__ int last;
    const int KK = KP * KP, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    double acc[NT][NT][2];
#pragma unroll
    for (int i = 0; i < NT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const int64_t n_groups = (P + 3) / 4;
    for (int64_t q = (int64_t)blockIdx.x * GV_WARPS + warp; q < n_groups; q += (int64_t)gridDim.x * GV_WARPS) {
        const int64_t j = 4 * q + t;
        double f[NT];
#pragma unroll
        for (int n = 0; n < NT; ++n) f[n] = (j < P) ? V[j * ldV + 8 * n + g] : 0.0;
#pragma unroll
        for (int n1 = 0; n1 < NT; ++n1)
#pragma unroll
            for (int n2 = n1; n2 < NT; ++n2) dmma(acc[n1][n2][0], acc[n1][n2][1], f[n1], f[n2]);
    }
#pragma unroll
    for (int n1 = 0; n1 < NT; ++n1)
#pragma unroll
        for (int n2 = n1; n2 < NT; ++n2)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int ra = 8 * n1 + g, cb = 8 * n2 + 2 * t + e;
                sm[warp][ra * KP + cb] = acc[n1][n2][e];
                sm[warp][cb * KP + ra] = acc[n1][n2][e];
            }
    __syncthreads();
    for (int e = tid; e < KK; e += GV_WARPS * 32) {
        double s = sm[0][e];
#pragma unroll
        for (int w = 1; w < GV_WARPS; ++w) s += sm[w][e];
        parts[(size_t)blockIdx.x * KK + e] = s;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) last = (atomicAdd(counter, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!last) return;
    __threadfence();
    const int nb = (int)gridDim.x;
    for (int e = tid; e < KK; e += GV_WARPS * 32) {
        double s = 0.0;
        int b = 0;
        for (; b + 8 <= nb; b += 8) {                                         // 8 independent loads in flight, summed in block order
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldcg(parts + (size_t)(b + u) * KK + e);
#pragma unroll
            for (int u = 0; u < 8; ++u) s += v[u];
        }
        for (; b < nb; ++b) s += __ldcg(parts + (size_t)b * KK + e);
        G[e] = s;
    }
    if (tid == 0) { *counter = 0u; if (bump) bump->als_iter += 1; }
}

}  // namespace

int gram_v_parts(int64_t P) { return (int)std::max<int64_t>(1, std::min<int64_t>(GV_BLOCKS, (P + 31) / 32)); }
void launch_gram_v(const Geom& g, const double* V, double* parts, double* G, unsigned int* counter, CheckState* bump, cudaStream_t st) {
    const int nb = gram_v_parts(g.P);
    switch (g.NT) {
        case 1: k_gram_v<1><<<nb, GV_WARPS * 32, 0, st>>>(V, g.ldV, g.KP, g.P, parts, G, counter, bump); break;
        case 2: k_gram_v<2><<<nb, GV_WARPS * 32, 0, st>>>(V, g.ldV, g.KP, g.P, parts, G, counter, bump); break;
        case 3: k_gram_v<3><<<nb, GV_WARPS * 32, 0, st>>>(V, g.ldV, g.KP, g.P, parts, G, counter, bump); break;
        default: k_gram_v<4><<<nb, GV_WARPS * 32, 0, st>>>(V, g.ldV, g.KP, g.P, parts, G, counter, bump); break;
    }
}

void launch_row_comp_gram(const Geom& g, const uint32_t* trR, const double* V, double* Dp, int n_splits, cudaStream_t st) {
    const int64_t warps = (int64_t)g.N * n_splits;
    const int blocks = (int)((warps + 7) / 8);
    switch (g.NT) {
        case 1: k_row_comp_gram<1><<<blocks, 256, 0, st>>>(trR, V, Dp, g.N, g.KP, g.ldV, g.WPr, g.P, n_splits); break;
        case 2: k_row_comp_gram<2><<<blocks, 256, 0, st>>>(trR, V, Dp, g.N, g.KP, g.ldV, g.WPr, g.P, n_splits); break;
        case 3: k_row_comp_gram<3><<<blocks, 256, 0, st>>>(trR, V, Dp, g.N, g.KP, g.ldV, g.WPr, g.P, n_splits); break;
        default: k_row_comp_gram<4><<<blocks, 256, 0, st>>>(trR, V, Dp, g.N, g.KP, g.ldV, g.WPr, g.P, n_splits); break;
    }
}

void launch_reduce_jobs(int n_jobs, double* const* out, const double* const* parts, const int64_t* n_elems, const int* n_parts, cudaStream_t st) {
    ReduceJobs jobs{};
    int blocks = 0;
    for (int j = 0; j < 3; ++j) {
        jobs.first_block[j] = blocks;
        if (j < n_jobs) {
            jobs.out[j] = out[j]; jobs.parts[j] = parts[j]; jobs.n[j] = n_elems[j]; jobs.n_parts[j] = n_parts[j];
            jobs.flat[j] = n_parts[j] < 32 ? 1 : 0;
            blocks += jobs.flat[j] ? (int)((n_elems[j] + 255) / 256) : (int)((n_elems[j] + 31) / 32);
        } else { jobs.out[j] = nullptr; jobs.parts[j] = nullptr; jobs.n[j] = 0; jobs.n_parts[j] = 0; jobs.flat[j] = 1; }
    }
    jobs.first_block[3] = blocks;
    for (int j = n_jobs; j < 3; ++j) jobs.first_block[j] = blocks;   // empty jobs own no blocks
    if (blocks) k_reduce<<<blocks, 256, 0, st>>>(jobs);
}

void launch_reduce_partials(double* out, const double* partials, int64_t n_elems, int n_parts, cudaStream_t st) {
    double* o[1] = {out}; const double* p[1] = {partials}; int64_t n[1] = {n_elems}; int np[1] = {n_parts};
    launch_reduce_jobs(1, o, p, n, np, st);
}

}  // namespace ib
