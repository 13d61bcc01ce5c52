// Developer probe: which SM does block b of a 1390 x 32-thread grid land on, and when does it start?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(32, 8) probe(int* smid, unsigned long long* t0, int spin_base) {
    extern __shared__ double sm[];
    unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
    unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (threadIdx.x == 0) { smid[blockIdx.x] = (int)s; t0[blockIdx.x] = t; }
    // spin: earlier blocks longer (like slots sorted by descending work)
    const long long spin = (long long)spin_base * (2000 - blockIdx.x);
    const long long c0 = clock64();
    double acc = sm[threadIdx.x];
    while (clock64() - c0 < spin) acc = acc * 1.0000001 + 1e-9;
    if (acc == 12345.678) sm[0] = acc;
}
int main() {
    const int nb = 1390;
    int* d_s; unsigned long long* d_t;
    cudaMalloc(&d_s, nb * 4); cudaMalloc(&d_t, nb * 8);
    probe<<<nb, 32, 17792>>>(d_s, d_t, 100);
    cudaDeviceSynchronize();
    int hs[nb]; unsigned long long ht[nb];
    cudaMemcpy(hs, d_s, nb * 4, cudaMemcpyDeviceToHost); cudaMemcpy(ht, d_t, nb * 8, cudaMemcpyDeviceToHost);
    unsigned long long tmin = ht[0]; for (int i = 0; i < nb; ++i) if (ht[i] < tmin) tmin = ht[i];
    printf("first 48 blocks: smid (start us)\n");
    for (int i = 0; i < 48; ++i) printf("%d:%d(%.1f) ", i, hs[i], (ht[i] - tmin) / 1e3);
    printf("\nblocks 1180..1200:\n");
    for (int i = 1180; i < 1200; ++i) printf("%d:%d(%.1f) ", i, hs[i], (ht[i] - tmin) / 1e3);
    int cnt[256] = {0}; int first8[256] = {0};
    for (int i = 0; i < nb; ++i) { cnt[hs[i]]++; if (i < 148) first8[hs[i]]++; }
    int mx = 0, mn = 1 << 30, nsm = 0, mx148 = 0;
    for (int s = 0; s < 256; ++s) if (cnt[s]) { ++nsm; if (cnt[s] > mx) mx = cnt[s]; if (cnt[s] < mn) mn = cnt[s]; if (first8[s] > mx148) mx148 = first8[s]; }
    printf("\n%d SMs used, blocks per SM min %d max %d; among the first 148 blocks the busiest SM got %d\n", nsm, mn, mx, mx148);
    return 0;
}
