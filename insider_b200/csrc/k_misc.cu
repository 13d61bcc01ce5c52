// Mask packing, check/decay ladder and small utilities.
#include <cmath>
#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"
#include "../../include/insider_b200.h"

namespace ib {

namespace {

// warp per (gene, word): lane b tests element (32 w + b) of the gene's column
template <typename T>
__global__ void __launch_bounds__(256) k_pack_mask(const T* __restrict__ src, int64_t N, int64_t n_genes, int Wp, uint32_t* __restrict__ dst) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (wid >= n_genes * Wp) return;
    const int64_t j = wid / Wp; const int w = (int)(wid % Wp);
    const int64_t i = 32 * (int64_t)w + lane;
    bool on = false;
    if (i < N) on = src[j * N + i] != (T)0;
    const uint32_t word = __ballot_sync(FULL, on);
    if (lane == 0) dst[j * Wp + w] = word;
}

__global__ void __launch_bounds__(256) k_transpose_mask(const uint32_t* __restrict__ trC, int64_t N, int64_t P_pad, int Wp, int WPr,
                                                        uint32_t* __restrict__ trR) {
    const int64_t x = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (x >= N * WPr) return;
    const int64_t i = x / WPr; const int wj = (int)(x % WPr);
    uint32_t word = 0;
    for (int b = 0; b < 32; ++b) {
        const int64_t j = 32 * (int64_t)wj + b;
        if (j < P_pad) word |= ((trC[j * Wp + (i >> 5)] >> (i & 31)) & 1u) << b;
    }
    trR[x] = word;
}

__global__ void __launch_bounds__(256) k_count_bits(const uint32_t* __restrict__ m, int64_t n, unsigned long long* out) {
    unsigned long long c = 0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) c += __popc(m[i]);
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// dst[r*dst_ld + c] = src[r*src_ld + c] for c < width: re-pitches a matrix on the device. Host <-> device transfers of the
// pitched layouts go through one contiguous copy + this kernel: cudaMemcpy2D with 44477 rows of 3 KB (Y) or 184 B (V) runs
// at 19 GB/s and 1.3 GB/s, a contiguous copy from pinned memory at ~55 GB/s.
__global__ void __launch_bounds__(256) k_repitch(double* __restrict__ dst, int64_t dst_ld, const double* __restrict__ src, int64_t src_ld, int width,
                                                 int64_t rows) {
    const int64_t n = rows * width;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const int64_t r = i / width; const int c = (int)(i - r * width);
        dst[r * dst_ld + c] = src[r * src_ld + c];
    }
}

__global__ void k_sse_reduce(const double* __restrict__ partial, int n_blocks, CheckState* st) {
    // 4 quantities x 32 lanes: lane l sums blocks l, l+32, ... then a fixed xor tree over lanes
    __shared__ double red[4][32];
    const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double s = 0.0;
    for (int b = lane; b < n_blocks; b += 32) s += partial[(size_t)b * 4 + q];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    if (lane == 0) (&st->sse_train)[q] = s;
    (void)red;
}

// src/utils.cpp:56-102 (rmse, loss terms) and src/optimize.cpp:381-408 (delta, decay ladder, convergence)
__global__ void __launch_bounds__(256) k_check(CheckState* st, const double* __restrict__ A_all, int64_t n_A, int initial, int iter, insider_check* rec) {
    __shared__ double red[256];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n_A; i += 256) s = fma(A_all[i], A_all[i], s);
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x != 0) return;
    const double row_reg = st->lambda1 * red[0];                                   // utils.cpp:85
    const double col_reg = st->lambda2 * (1.0 - st->alpha) * st->v2;               // :88
    const double l1_reg = st->lambda2 * st->alpha * st->v1;                        // :91
    const double sse = st->sse_train;
    double train_rmse, test_rmse = nan("");
    if (st->tuning == 0) train_rmse = sqrt(sse / st->np_total);                    // :63
    else { train_rmse = sqrt(sse / st->n_train); test_rmse = sqrt(st->sse_test / st->n_test); }   // :66-67
    const double loss = sse / 2 + row_reg / 2 + col_reg / 2 + l1_reg;              // :93
    double delta = 0.0, decay = 1.0;
    if (initial) {
        st->loss = loss; st->pre_loss = loss; st->decay = 1.0; st->tol = st->sub_tol; st->converged = 0;
    } else {
        const double pre = st->loss;                                               // optimize.cpp:382
        delta = pre - loss;                                                        // :386
        if (delta / 1000 <= 1e-6) decay = 1e-6;                                    // :389-403
        else if (delta / 1000 <= 1e-5) decay = 1e-5;
        else if (delta / 1000 <= 1e-4) decay = 1e-4;
        else if (delta / 1000 <= 1e-3) decay = 1e-3;
        else if (delta / 1000 <= 1e-2) decay = 1e-2;
        else if (delta / 1000 <= 1e-1) decay = 1e-1;
        else decay = 1.0;
        st->pre_loss = pre; st->loss = loss; st->decay = decay; st->tol = st->sub_tol * decay;   // :376 sub_tol * decay
        st->converged = ((pre - loss) / pre < st->global_tol) ? 1 : 0;             // :405
    }
    st->diverged = isfinite(loss) ? 0 : 1;
    st->row_reg = row_reg; st->train_rmse = train_rmse; st->test_rmse = test_rmse; st->delta_loss = delta;
    if (rec) {
        rec->iter = initial ? -1 : iter; rec->pad = 0;
        rec->sum_residual = sse; rec->train_rmse = train_rmse; rec->test_rmse = test_rmse;
        rec->row_reg = row_reg / 2; rec->col_reg = col_reg / 2; rec->l1_reg = l1_reg; rec->loss = loss; rec->delta_loss = delta; rec->decay = decay;
    }
}

__global__ void k_bump_iter(CheckState* st) { st->als_iter += 1; }

// part[block][i] = sum over the block's genes of Y[j][i]^2 (threads along rows: coalesced); a fixed-order reduce follows
__global__ void __launch_bounds__(256) k_row_sumsq(const double* __restrict__ Y, int N, int ldY, int64_t P, double* __restrict__ part) {
    const int64_t per = (P + gridDim.x - 1) / gridDim.x;
    const int64_t j0 = (int64_t)blockIdx.x * per, j1 = (j0 + per < P) ? j0 + per : P;
    for (int i = threadIdx.x; i < N; i += 256) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int64_t j = j0;
        for (; j + 4 <= j1; j += 4) {
            const double a = Y[j * ldY + i], b = Y[(j + 1) * ldY + i], c = Y[(j + 2) * ldY + i], d = Y[(j + 3) * ldY + i];
            s0 = fma(a, a, s0); s1 = fma(b, b, s1); s2 = fma(c, c, s2); s3 = fma(d, d, s3);
        }
        for (; j < j1; ++j) { const double a = Y[j * ldY + i]; s0 = fma(a, a, s0); }
        part[(size_t)blockIdx.x * N + i] = (s0 + s1) + (s2 + s3);
    }
}

}  // namespace

void launch_pack_mask(const void* src, int kind, int64_t N, int64_t n_genes, int Wp, uint32_t* dstC, cudaStream_t st) {
    const int64_t warps = n_genes * Wp;
    const int blocks = (int)((warps + 7) / 8);
    if (blocks == 0) return;
    if (kind == INSIDER_MASK_INT32) k_pack_mask<int32_t><<<blocks, 256, 0, st>>>((const int32_t*)src, N, n_genes, Wp, dstC);
    else if (kind == INSIDER_MASK_UINT8) k_pack_mask<uint8_t><<<blocks, 256, 0, st>>>((const uint8_t*)src, N, n_genes, Wp, dstC);
    else k_pack_mask<double><<<blocks, 256, 0, st>>>((const double*)src, N, n_genes, Wp, dstC);
}

void launch_transpose_mask(const uint32_t* trC, int64_t N, int64_t P_pad, int Wp, int WPr, uint32_t* trR, cudaStream_t st) {
    const int64_t n = N * WPr;
    k_transpose_mask<<<(int)((n + 255) / 256), 256, 0, st>>>(trC, N, P_pad, Wp, WPr, trR);
}

void launch_repitch(double* dst, int64_t dst_ld, const double* src, int64_t src_ld, int width, int64_t rows, cudaStream_t st) {
    if (rows <= 0 || width <= 0) return;
    const int64_t n = rows * width;
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
    k_repitch<<<blocks, 256, 0, st>>>(dst, dst_ld, src, src_ld, width, rows);
}
void launch_count_bits(const uint32_t* m, int64_t n_words, unsigned long long* out, cudaStream_t st) {
    if (n_words == 0) return;
    int blocks = (int)((n_words + 255) / 256);
    if (blocks > 1184) blocks = 1184;
    k_count_bits<<<blocks, 256, 0, st>>>(m, n_words, out);
}

void launch_sse_reduce(const double* partial, int n_blocks, CheckState* state, cudaStream_t st) { k_sse_reduce<<<1, 128, 0, st>>>(partial, n_blocks, state); }

void launch_check(CheckState* state, const double* A_all, int64_t n_A, int initial, int iter, void* record_out, cudaStream_t st) {
    k_check<<<1, 256, 0, st>>>(state, A_all, n_A, initial, iter, (insider_check*)record_out);
}

void launch_bump_iter(CheckState* state, cudaStream_t st) { k_bump_iter<<<1, 1, 0, st>>>(state); }

void launch_row_sumsq(const Geom& g, const double* Y, double* part, int n_blocks, cudaStream_t st) {
    if (g.P > 0) k_row_sumsq<<<n_blocks, 256, 0, st>>>(Y, g.N, g.ldY, g.P, part);
}

}  // namespace ib
