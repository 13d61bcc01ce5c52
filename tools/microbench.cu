// Microbenchmarks that fix the design constants for the FP64 hot path on B200 (sm_100a):
//   DFMA issue peak, DMMA (mma.sync f64) peak for the shapes ptxas accepts, streaming HBM read bandwidth,
//   and shared-memory broadcast/unique load rate. Output is one line per probe on stdout.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// m8n8k4: A 1 reg, B 1 reg, C/D 2 regs per lane.
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__global__ void dmma884_kernel(double* out, int iters, double a, double b) {
    double c[8][2];
    for (int t = 0; t < 8; ++t) { c[t][0] = threadIdx.x + t; c[t][1] = t; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int t = 0; t < 8; ++t) dmma884(c[t][0], c[t][1], a, b);
    }
    double s = 0; for (int t = 0; t < 8; ++t) s += c[t][0] + c[t][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// dependent chain (latency)
__global__ void dmma884_lat_kernel(double* out, int iters, double a, double b, long long* cyc) {
    double c0 = threadIdx.x, c1 = 1;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) dmma884(c0, c1, a, b);
    long long t1 = clock64();
    out[threadIdx.x] = c0 + c1;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void dfma_lat_kernel(double* out, int iters, double a, double b, long long* cyc) {
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) x = fma(x, a, b);
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void ddiv_lat_kernel(double* out, int iters, double a, long long* cyc) {
    double x = threadIdx.x + 3.0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) x = a / x + 1.5;
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void shfl_lat_kernel(double* out, int iters, long long* cyc) {
    double x = threadIdx.x + 3.0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31) + 1.0;
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

#if defined(HAVE_M16)
__device__ __forceinline__ void dmma1684(double (&d)[4], const double (&a)[2], double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__global__ void dmma1684_kernel(double* out, int iters, double av, double bv) {
    double c[4][4]; double a[2] = {av, av + 1};
    for (int t = 0; t < 4; ++t) for (int q = 0; q < 4; ++q) c[t][q] = threadIdx.x + t + q;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int t = 0; t < 4; ++t) dmma1684(c[t], a, bv);
    }
    double s = 0; for (int t = 0; t < 4; ++t) for (int q = 0; q < 4; ++q) s += c[t][q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dmma1688_kernel(double* out, int iters, double av, double bv) {
    double c[4][4]; double a[4] = {av, av + 1, av + 2, av + 3}; double b[2] = {bv, bv + 1};
    for (int t = 0; t < 4; ++t) for (int q = 0; q < 4; ++q) c[t][q] = threadIdx.x + t + q;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int t = 0; t < 4; ++t) dmma1688(c[t], a, b);
    }
    double s = 0; for (int t = 0; t < 4; ++t) for (int q = 0; q < 4; ++q) s += c[t][q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dmma16816_kernel(double* out, int iters, double av, double bv) {
    double c[4][4]; double a[8], b[4];
    for (int q = 0; q < 8; ++q) a[q] = av + q;
    for (int q = 0; q < 4; ++q) b[q] = bv + q;
    for (int t = 0; t < 4; ++t) for (int q = 0; q < 4; ++q) c[t][q] = threadIdx.x + t + q;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int t = 0; t < 4; ++t) dmma16816(c[t], a, b);
    }
    double s = 0; for (int t = 0; t < 4; ++t) for (int q = 0; q < 4; ++q) s += c[t][q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
#endif

// Verifies the internal accumulation order of m8n8k4: compares D against sequential-fma and other orders.
__global__ void dmma_order_kernel(const double* A, const double* B, double* D) {
    // A is 8x4 row-major, B is 4x8 (col-major fragment: B[k][n]), one warp.
    int lane = threadIdx.x;
    double a = A[(lane >> 2) * 4 + (lane & 3)];
    double b = B[(lane & 3) * 8 + (lane >> 2)];
    double d0 = 0.0, d1 = 0.0;
    dmma884(d0, d1, a, b);
    D[(lane >> 2) * 8 + (lane & 3) * 2 + 0] = d0;
    D[(lane >> 2) * 8 + (lane & 3) * 2 + 1] = d1;
}

__global__ void hbm_read_kernel(const double2* __restrict__ in, size_t n2, double* out) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    double s = 0;
    for (; i + 3 * stride < n2; i += 4 * stride) {
        double2 a = in[i], b = in[i + stride], c = in[i + 2 * stride], d = in[i + 3 * stride];
        s += a.x + a.y + b.x + b.y + c.x + c.y + d.x + d.y;
    }
    for (; i < n2; i += stride) { double2 a = in[i]; s += a.x + a.y; }
    if (s == 123.456) out[0] = s;
}

// smem: broadcast LDS.128 vs unique LDS.64 issue rates
__global__ void lds_kernel(double* out, int iters, int mode) {
    __shared__ double sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
    __syncthreads();
    double s = 0;
    int lane = threadIdx.x & 31;
    if (mode == 0) {  // broadcast 128-bit
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 8; ++u) { double2 v = *reinterpret_cast<const double2*>(&sm[((i + u) * 2) & 4094]); s += v.x + v.y; }
        }
    } else {          // unique 64-bit per lane, conflict-free
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 8; ++u) { s += sm[(((i + u) * 32) & 4064) + lane]; }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> static float time_ms(F f, int reps = 5) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s sms %d clock_khz %d smem_optin %zu l2 %d\n", p.name, p.multiProcessorCount, p.clockRate, p.sharedMemPerBlockOptin, p.l2CacheSize);
    int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 16 * 1024));
    long long* cyc; CK(cudaMallocManaged(&cyc, 8));
    const int iters = 20000;
    for (int bps : {1, 2, 4}) for (int thr : {128, 256, 512}) {
        float ms = time_ms([&] { dfma_kernel<<<sms * bps, thr>>>(out, iters, 1.0000001, 1e-9); });
        double fl = 2.0 * 8 * iters * (double)sms * bps * thr;
        printf("dfma blocks/sm %d threads %d : %.2f TFLOP/s (%.3f ms)\n", bps, thr, fl / ms / 1e9, ms);
    }
    for (int bps : {1, 2, 4}) for (int thr : {128, 256, 512}) {
        float ms = time_ms([&] { dmma884_kernel<<<sms * bps, thr>>>(out, iters / 4, 1.0000001, 1e-9); });
        double fl = 2.0 * 256 * 8 * (iters / 4) * (double)sms * bps * (thr / 32);
        printf("dmma m8n8k4 blocks/sm %d threads %d : %.2f TFLOP/s (%.3f ms)\n", bps, thr, fl / ms / 1e9, ms);
    }
#if defined(HAVE_M16)
    for (int thr : {128, 256, 512}) {
        float ms = time_ms([&] { dmma1684_kernel<<<sms * 2, thr>>>(out, iters / 4, 1.0000001, 1e-9); });
        printf("dmma m16n8k4 threads %d : %.2f TFLOP/s\n", thr, 2.0 * 512 * 4 * (iters / 4) * (double)sms * 2 * (thr / 32) / ms / 1e9);
        ms = time_ms([&] { dmma1688_kernel<<<sms * 2, thr>>>(out, iters / 4, 1.0000001, 1e-9); });
        printf("dmma m16n8k8 threads %d : %.2f TFLOP/s\n", thr, 2.0 * 1024 * 4 * (iters / 4) * (double)sms * 2 * (thr / 32) / ms / 1e9);
        ms = time_ms([&] { dmma16816_kernel<<<sms * 2, thr>>>(out, iters / 4, 1.0000001, 1e-9); });
        printf("dmma m16n8k16 threads %d : %.2f TFLOP/s\n", thr, 2.0 * 2048 * 4 * (iters / 4) * (double)sms * 2 * (thr / 32) / ms / 1e9);
    }
#endif
    dmma884_lat_kernel<<<1, 32>>>(out, 4096, 1.0000001, 1e-9, cyc); CK(cudaDeviceSynchronize());
    printf("dmma m8n8k4 dependent latency: %.1f cycles\n", (double)*cyc / 4096);
    dfma_lat_kernel<<<1, 32>>>(out, 4096, 1.0000001, 1e-9, cyc); CK(cudaDeviceSynchronize());
    printf("dfma dependent latency: %.1f cycles\n", (double)*cyc / 4096);
    ddiv_lat_kernel<<<1, 32>>>(out, 4096, 1.7, cyc); CK(cudaDeviceSynchronize());
    printf("ddiv+dadd dependent latency: %.1f cycles\n", (double)*cyc / 4096);
    shfl_lat_kernel<<<1, 32>>>(out, 4096, cyc); CK(cudaDeviceSynchronize());
    printf("shfl64+dadd dependent latency: %.1f cycles\n", (double)*cyc / 4096);

    {   // accumulation order of the DMMA
        double hA[32], hB[32], hD[64], *dA, *dB, *dD;
        // choose values so different summation orders round differently
        unsigned long long s = 88172645463325252ull;
        auto rnd = [&] { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) / 9007199254740992.0 - 0.5; };
        int same_seq = 0, same_rev = 0, same_pair = 0, same_nofma = 0, total = 0;
        CK(cudaMalloc(&dA, 256)); CK(cudaMalloc(&dB, 256)); CK(cudaMalloc(&dD, 512));
        for (int trial = 0; trial < 200; ++trial) {
            for (int i = 0; i < 32; ++i) { hA[i] = rnd() * (1 + 1e3 * (i % 3)); hB[i] = rnd(); }
            CK(cudaMemcpy(dA, hA, 256, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB, 256, cudaMemcpyHostToDevice));
            dmma_order_kernel<<<1, 32>>>(dA, dB, dD); CK(cudaMemcpy(hD, dD, 512, cudaMemcpyDeviceToHost));
            for (int m = 0; m < 8; ++m) for (int n = 0; n < 8; ++n) {
                double a0 = hA[m * 4], a1 = hA[m * 4 + 1], a2 = hA[m * 4 + 2], a3 = hA[m * 4 + 3];
                double b0 = hB[n], b1 = hB[8 + n], b2 = hB[16 + n], b3 = hB[24 + n];
                double seq = __builtin_fma(a3, b3, __builtin_fma(a2, b2, __builtin_fma(a1, b1, __builtin_fma(a0, b0, 0.0))));
                double rev = __builtin_fma(a0, b0, __builtin_fma(a1, b1, __builtin_fma(a2, b2, __builtin_fma(a3, b3, 0.0))));
                double pr = __builtin_fma(a1, b1, a0 * b0) + __builtin_fma(a3, b3, a2 * b2);
                volatile double p0 = a0 * b0, p1 = a1 * b1, p2 = a2 * b2, p3 = a3 * b3;
                double nof = ((p0 + p1) + p2) + p3;
                double d = hD[m * 8 + n];
                same_seq += d == seq; same_rev += d == rev; same_pair += d == pr; same_nofma += d == nof; ++total;
            }
        }
        printf("dmma order: of %d outputs match seq-fma %d, rev-fma %d, pairwise %d, no-fma-seq %d\n", total, same_seq, same_rev, same_pair, same_nofma);
    }
    {
        size_t bytes = (size_t)2 << 30; double2* buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 0, bytes));
        for (int bps : {2, 4, 8, 16}) {
            float ms = time_ms([&] { hbm_read_kernel<<<sms * bps, 256>>>(buf, bytes / 16, out); });
            printf("hbm read blocks/sm %d : %.1f GB/s\n", bps, bytes / ms / 1e6);
        }
        CK(cudaFree(buf));
    }
    for (int mode : {0, 1}) {
        float ms = time_ms([&] { lds_kernel<<<sms * 2, 512>>>(out, 20000, mode); });
        double insts = 8.0 * 20000 * sms * 2 * 16;
        printf("lds mode %d (%s): %.2f warp-LDS/clk/SM at %d kHz nominal\n", mode, mode ? "unique64" : "bcast128", insts / (ms * 1e-3) / sms / (p.clockRate * 1e3), p.clockRate);
    }
    return 0;
}
