// Microbenchmark (B200): where should the shared XtX row of the dense elastic-net step come from?
//   smem : 12 broadcast LDS.128 per step (what k_cd_dense does), row index dynamic
//   const: `switch (k)` to 24 step bodies whose 24 FMA operands are compile-time addresses in a __constant__ table
//          (no LDS at all; the question is whether the constant cache sustains a 4.6 KB table visited in random order)
//   dmma : the rank-1 update q -= delta x row on the FP64 tensor pipe WITHOUT leaving the thread-per-gene layout: tile j of 12 holds
//          coordinates 2j, 2j+1 of all 32 genes as an m8n8k4 accumulator (thread (r, c) = gene 4r + c), A = the thread's own delta,
//          B = block-diagonal [c'' == c'] x row[2j + e] (one non-broadcast LDS.64 per tile and lane): 12 DMMAs per step, each doing
//          64 useful of its 256 FMAs
// One warp per block like k_cd_dense; reports cycles per warp-step for a lone warp and for 8 blocks per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mb_const_table tools/mb_const_table.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int KT = 24;
__constant__ double ctab[KT * KT];
__constant__ unsigned char cord[4096];          // visiting sequence (random coordinates)

template <int R>
__device__ __forceinline__ void step_const(double (&q)[KT], double la) {
    const double up = q[R];
    const double t1 = fabs(up) - la;
    double nb = copysign(t1, up) * ctab[R * KT + R];
    nb = (__double2hiint(t1) >= 0) ? nb : 0.0;
    const double nd = -nb * 1e-3;
#pragma unroll
    for (int l = 0; l < KT; ++l) if (l != R) q[l] = fma(nd, ctab[R * KT + l], q[l]);
}

__global__ void __launch_bounds__(32, 8) k_const(double* out, int steps, double la) {
    double q[KT];
#pragma unroll
    for (int l = 0; l < KT; ++l) q[l] = 1.0 + 0.01 * l + 1e-3 * threadIdx.x;
    for (int i = 0; i < steps; ++i) {
        const int k = cord[(i + blockIdx.x * 7) & 4095];
        switch (k) {
#define C(r) case r: step_const<r>(q, la); break;
            C(0) C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12) C(13) C(14) C(15) C(16) C(17) C(18) C(19) C(20) C(21) C(22) C(23)
#undef C
        }
    }
    double s = 0;
#pragma unroll
    for (int l = 0; l < KT; ++l) s += q[l];
    out[blockIdx.x * 32 + threadIdx.x] = s;
}

// position order like k_cd_dense: the step reads q[I] at compile-time I, the ROW is dynamic (shared memory, broadcast)
template <int I>
__device__ __forceinline__ void step_smem(double (&q)[KT], const double* row, double la) {
    const double up = q[I];
    const double t1 = fabs(up) - la;
    double nb = copysign(t1, up) * row[I];
    nb = (__double2hiint(t1) >= 0) ? nb : 0.0;
    const double nd = -nb * 1e-3;
#pragma unroll
    for (int l = 0; l < KT; l += 2) {
        const double2 x = *reinterpret_cast<const double2*>(row + l);
        if (l != I) q[l] = fma(nd, x.x, q[l]);
        if (l + 1 != I) q[l + 1] = fma(nd, x.y, q[l + 1]);
    }
}
template <int... Is>
__device__ __forceinline__ void sweep_smem(double (&q)[KT], const double* tab, const unsigned char* ord, double la) {
    ((step_smem<Is>(q, tab + ord[Is] * KT, la)), ...);
}
__global__ void __launch_bounds__(32, 8) k_smem(const double* tab_g, double* out, int steps, double la) {
    __shared__ __align__(16) double tab[KT * KT];
    __shared__ unsigned char ord[4096];
    for (int x = threadIdx.x; x < KT * KT; x += 32) tab[x] = tab_g[x];
    for (int x = threadIdx.x; x < 4096; x += 32) ord[x] = cord[x];
    __syncwarp();
    double q[KT];
#pragma unroll
    for (int l = 0; l < KT; ++l) q[l] = 1.0 + 0.01 * l + 1e-3 * threadIdx.x;
    for (int i = 0; i < steps; i += KT) {
        const unsigned char* o = ord + ((i + blockIdx.x * 7) & 4095 & ~31);
        sweep_smem<0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23>(q, tab, o, la);
    }
    double s = 0;
#pragma unroll
    for (int l = 0; l < KT; ++l) s += q[l];
    out[blockIdx.x * 32 + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(32, 8) k_dmma(const double* tab_g, double* out, int steps, double la) {
    __shared__ __align__(16) double tab[KT * KT];
    __shared__ unsigned char ord[4096];
    for (int x = threadIdx.x; x < KT * KT; x += 32) tab[x] = tab_g[x];
    for (int x = threadIdx.x; x < 4096; x += 32) ord[x] = cord[x];
    __syncwarp();
    const int lane = threadIdx.x, kk = lane & 3, nn = lane >> 2;       // B fragment element (k = lane % 4, n = lane / 4)
    const bool nz = kk == (nn >> 1);                                   // block diagonal: column n belongs to c' = n / 2
    const int e = nn & 1;
    double q[KT];                                                      // q[2j], q[2j+1]: accumulator pair of tile j
#pragma unroll
    for (int l = 0; l < KT; ++l) q[l] = 1.0 + 0.01 * l + 1e-3 * threadIdx.x;
    for (int i = 0; i < steps; ++i) {
        const int k = ord[(i + blockIdx.x * 7) & 4095];
        const double* row = tab + k * KT;
        // scalar chain on this thread's gene (coordinate k is dynamic here: the microbenchmark reads a fixed register, the real kernel
        // would relabel as k_cd_dense does)
        const double up = q[0];
        const double t1 = fabs(up) - la;
        double nb = copysign(t1, up) * row[k];
        nb = (__double2hiint(t1) >= 0) ? nb : 0.0;
        const double nd = -nb * 1e-3;
#pragma unroll
        for (int j = 0; j < KT / 2; ++j) {
            const double b = nz ? row[2 * j + e] : 0.0;
            dmma884(q[2 * j], q[2 * j + 1], nd, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int l = 0; l < KT; ++l) s += q[l];
    out[blockIdx.x * 32 + threadIdx.x] = s;
}

int main() {
    double h[KT * KT];
    srand(1);
    for (int i = 0; i < KT * KT; ++i) h[i] = 1e-3 * (rand() % 1000) / 1000.0;
    unsigned char ho[4096];
    for (int i = 0; i < 4096; ++i) ho[i] = (unsigned char)(rand() % KT);
    CK(cudaMemcpyToSymbol(ctab, h, sizeof(h)));
    CK(cudaMemcpyToSymbol(cord, ho, sizeof(ho)));
    double *tab_g, *out;
    CK(cudaMalloc(&tab_g, sizeof(h))); CK(cudaMemcpy(tab_g, h, sizeof(h), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&out, 148 * 16 * 32 * 8));
    int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int steps = 24 * 20000;
    for (int blocks : {1, 148, 148 * 4, 148 * 8}) {
        for (int v = 0; v < 3; ++v) {
            float best = 1e9;
            for (int rep = 0; rep < 3; ++rep) {
                CK(cudaEventRecord(e0));
                if (v == 0) k_smem<<<blocks, 32>>>(tab_g, out, steps, 0.5); else if (v == 1) k_const<<<blocks, 32>>>(out, steps, 0.5);
                else k_dmma<<<blocks, 32>>>(tab_g, out, steps, 0.5);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
            }
            CK(cudaGetLastError());
            const double clk = best * 1e-3 * khz * 1e3;
            const double per_sm = (blocks + 147) / 148;
            printf("%s blocks %5d: %.3f ms, %.1f clk per warp-step (one warp), %.1f clk per warp-step per SM\n", v == 0 ? "smem " : v == 1 ? "const" : "dmma ", blocks, best,
                   clk / steps, clk / steps / per_sm);
        }
    }
    return 0;
}
