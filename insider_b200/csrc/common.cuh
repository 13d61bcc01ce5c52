// Shared device/host helpers for libinsider_b200 (sm_100a only).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

namespace ib {

constexpr int WARP = 32;
constexpr unsigned FULL = 0xffffffffu;
constexpr int KMAX = 32;           // latent_dim limit: one lane per coordinate in the per-gene solver
constexpr int TG = 16;             // genes per streamed tile (two 8-wide DMMA n-tiles)

// Pitches: a leading dimension p with p % 8 == 4 makes every DMMA fragment load (address = t*p + g or g*p + t,
// g in 0..7, t in 0..3) bank-conflict-free for 64-bit shared-memory accesses.
__host__ __device__ inline int pitch4(int n) { int p = (n + 7) / 8 * 8 + 4; return (p - 8 >= n) ? p - 8 : p; }
__host__ __device__ inline int round_up(int n, int m) { return (n + m - 1) / m * m; }

// ---------------------------------------------------------------------------------------------------------------
// counter-based permutation source (mode B) — must match oracle/insider_oracle.cpp:randperm bit-for-bit
__host__ __device__ inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// Key of sweep `draw` of a gene's solve in ALS iteration `als_iter`. The gene is NOT part of the key: every gene at the same
// sweep index shares the visiting order (each gene still sees a fresh uniformly random order per sweep), which makes the
// coordinate warp-uniform in the thread-per-gene solver (k_cd_dense.cu).
__host__ __device__ inline uint64_t perm_key(uint64_t seed, uint32_t als_iter, uint32_t draw) {
    return mix64(seed + 0x9E3779B97F4A7C15ull * (1ull + als_iter)) ^ mix64((uint64_t)draw * 0x8CB92BA72F3D8DD7ull + 0x2545F4914F6CDD1Dull);
}
__host__ __device__ inline uint32_t perm_value(uint64_t key, int pos) {
    return (uint32_t)(mix64(key + 0x9E3779B97F4A7C15ull * (uint64_t)(pos + 1)) >> 38);   // 26-bit sort key
}
// Permutation tables: for every size n = 1..32, PERM_T permutations, each built like arma::randperm (sort n random
// keys ascending, ties by index). A sweep selects t = perm_select(perm_key(seed, als_iter, draw)) with n = K (all
// coordinates); a gene visits its active coordinates in that order. Identical in oracle/insider_oracle.cpp.
//   rank table : table[((n-1)*PERM_T + t)*32 + p]                = position of coordinate p in the visiting order
//   order table: table[PERM_TABLE_HALF + ((n-1)*PERM_T + t)*32 + i] = coordinate visited at position i
constexpr int PERM_T = 4096;
constexpr int PERM_NMAX = 32;
constexpr size_t PERM_TABLE_HALF = (size_t)PERM_NMAX * PERM_T * 32;
constexpr size_t PERM_TABLE_BYTES = 2 * PERM_TABLE_HALF;
__host__ __device__ inline uint64_t perm_table_key(int n, int t) { return mix64(0x1F83D9ABFB41BD6Bull ^ (((uint64_t)n << 32) | (uint64_t)t)); }
__host__ __device__ inline uint32_t perm_select(uint64_t pk) { return (uint32_t)(pk >> 20) & (uint32_t)(PERM_T - 1); }
inline void build_perm_table(unsigned char* table /* [PERM_TABLE_BYTES] */) {
    for (int n = 1; n <= PERM_NMAX; ++n)
        for (int t = 0; t < PERM_T; ++t) {
            const uint64_t key = perm_table_key(n, t);
            uint32_t v[PERM_NMAX];
            for (int p = 0; p < n; ++p) v[p] = perm_value(key, p);
            unsigned char* row = table + ((size_t)(n - 1) * PERM_T + t) * 32;
            unsigned char* ord = row + PERM_TABLE_HALF;
            for (int p = 0; p < 32; ++p) {
                int r = 0;
                if (p < n) { for (int m = 0; m < n; ++m) r += (v[m] < v[p]) || (v[m] == v[p] && m < p); } else r = p;
                row[p] = (unsigned char)r;
                ord[r] = (unsigned char)p;
            }
        }
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------------------------
// warp helpers
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

// Copies n doubles global -> shared with a warp/group of `nthr` threads (thread index `t`), eight independent loads in
// flight per thread: a plain `for (x = t; x < n; x += nthr) dst[x] = src[x]` loop issues its loads one latency at a time.
__device__ __forceinline__ void copy_batched(double* dst, const double* __restrict__ src, int n, int t, int nthr) {
    int x = t;
    for (; x + 7 * nthr < n; x += 8 * nthr) {
        const double v0 = src[x], v1 = src[x + nthr], v2 = src[x + 2 * nthr], v3 = src[x + 3 * nthr];
        const double v4 = src[x + 4 * nthr], v5 = src[x + 5 * nthr], v6 = src[x + 6 * nthr], v7 = src[x + 7 * nthr];
        dst[x] = v0; dst[x + nthr] = v1; dst[x + 2 * nthr] = v2; dst[x + 3 * nthr] = v3;
        dst[x + 4 * nthr] = v4; dst[x + 5 * nthr] = v5; dst[x + 6 * nthr] = v6; dst[x + 7 * nthr] = v7;
    }
    for (; x < n; x += nthr) dst[x] = src[x];
}

// FP64 tensor-core MMA: D(8x8) += A(8x4, row) * B(4x8, col). Lane l: g = l>>2, t = l&3.
//   a = A[g][t], b = B[t][g], d0 = D[g][2t], d1 = D[g][2t+1].
// Measured on B200 (profiles/r01_microbench_fp64_hbm.txt): result == fma(a3,b3,fma(a2,b2,fma(a1,b1,fma(a0,b0,c)))).
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------------------------------------------
// mbarrier + TMA bulk copy (cp.async.bulk, 1-D): global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// bytes must be a multiple of 16, src and dst 16-byte aligned
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
#endif  // __CUDACC__

}  // namespace ib
