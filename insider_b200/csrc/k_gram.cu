// Gram matrices and fixed-order reductions.
//
//   k_gram_v         gram = V V^T                                   src/optimize.cpp:332
//   k_row_comp_gram  per-row complement  sum_{j: m_ij=0} v_j v_j^T  src/optimize.cpp:163,170 (c_factor.cols(zero_idx) * trans(..))
//   k_reduce         deterministic sum of per-block partial buffers
#include "common.cuh"
#include "kernels.cuh"

namespace ib {

namespace {

__global__ void __launch_bounds__(256) k_gram_v(const double* __restrict__ V, double* __restrict__ Gp, int KP, int ldV, int64_t P_pad,
                                                int n_blocks) {
    extern __shared__ double vs[];                 // [GT][ldV]
    constexpr int GT = 64;
    const int tid = threadIdx.x;
    const int64_t per = (P_pad + n_blocks - 1) / n_blocks;
    const int64_t j0 = (int64_t)blockIdx.x * per, j1 = min(P_pad, j0 + per);
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int64_t jb = j0; jb < j1; jb += GT) {
        const int n = (int)min((int64_t)GT, j1 - jb);
        __syncthreads();
        for (int x = tid; x < n * ldV; x += 256) vs[x] = V[jb * ldV + x];
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int e = tid + 256 * r;
            if (e < KP * KP) {
                const int a = e / KP, b = e % KP;
                double s = acc[r];
                for (int j = 0; j < n; ++j) s = fma(vs[j * ldV + a], vs[j * ldV + b], s);
                acc[r] = s;
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int e = tid + 256 * r;
        if (e < KP * KP) Gp[(size_t)blockIdx.x * KP * KP + e] = acc[r];
    }
}

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

__global__ void __launch_bounds__(256) k_reduce(double* __restrict__ out, const double* __restrict__ parts, int64_t n, int n_parts) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int p = 0; p < n_parts; ++p) s += parts[(size_t)p * n + i];
    out[i] = s;
}

}  // namespace

void launch_gram_v(const Geom& g, const double* V, double* Gp, int n_blocks, cudaStream_t st) {
    const size_t smem = (size_t)64 * g.ldV * 8;
    k_gram_v<<<n_blocks, 256, smem, st>>>(V, Gp, g.KP, g.ldV, g.P_pad, n_blocks);
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

void launch_reduce_partials(double* out, const double* partials, int64_t n_elems, int n_parts, cudaStream_t st) {
    const int blocks = (int)((n_elems + 255) / 256);
    k_reduce<<<blocks, 256, 0, st>>>(out, partials, n_elems, n_parts);
}

}  // namespace ib
