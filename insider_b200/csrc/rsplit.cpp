// insider_b200_split — the train/test split of ratio_splitter() (reference R/utils.R:78-117), bit-exact with R:
// set.seed(seed) (Mersenne-Twister, R's seed scrambling), sample(existing_idx, floor(n * ratio)) without replacement
// using R >= 3.6 "Rejection" sampling (R_unif_index), including the hashed sample2 path R takes when n > 1e7.
// The R sources are not part of the reference tree; this is a restatement pinned by the
// known answers in tests/test_r_rng.py. Host code: R's generator is one sequential stream.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/insider_b200.h"

namespace {

struct RMersenne {
    uint32_t mt[624];
    int mti = 624;
    explicit RMersenne(uint32_t seed) {
        for (int j = 0; j < 50; ++j) seed = 69069u * seed + 1u;
        for (int j = 0; j < 625; ++j) {
            seed = 69069u * seed + 1u;
            if (j > 0) mt[j - 1] = seed;      // i_seed[0] holds mti, forced to 624 by FixupSeeds
        }
    }
    uint32_t next() {
        static const uint32_t mag[2] = {0u, 0x9908b0dfu};
        if (mti >= 624) {
            int k = 0;
            for (; k < 227; ++k) { uint32_t y = (mt[k] & 0x80000000u) | (mt[k + 1] & 0x7fffffffu); mt[k] = mt[k + 397] ^ (y >> 1) ^ mag[y & 1u]; }
            for (; k < 623; ++k) { uint32_t y = (mt[k] & 0x80000000u) | (mt[k + 1] & 0x7fffffffu); mt[k] = mt[k - 227] ^ (y >> 1) ^ mag[y & 1u]; }
            uint32_t y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
            mt[623] = mt[396] ^ (y >> 1) ^ mag[y & 1u];
            mti = 0;
        }
        uint32_t y = mt[mti++];
        y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
        return y;
    }
    double unif() {
        const double lo = 2.328306437080797e-10;
        double x = (double)next() * 2.3283064365386963e-10;
        if (x <= 0.0) return 0.5 * lo;
        if (1.0 - x <= 0.0) return 1.0 - 0.5 * lo;
        return x;
    }
    double rbits(int bits) {
        int64_t v = 0;
        for (int n = 0; n <= bits; n += 16) v = 65536 * v + (int)std::floor(unif() * 65536);
        if (bits < 64) v &= ((int64_t)1 << bits) - 1;
        return (double)v;
    }
    double unif_index(double dn) {
        if (dn <= 0) return 0.0;
        const int bits = (int)std::ceil(std::log2(dn));
        double dv;
        do { dv = rbits(bits); } while (dn <= dv);
        return dv;
    }
};

}  // namespace

extern "C" int insider_b200_split(const double* data, int64_t N, int64_t P, double ratio, uint32_t seed, int32_t* train, int32_t* test,
                                  int32_t* na, int64_t* n_test, char* errbuf, size_t errlen) {
    if (!data || !train || !test || N <= 0 || P <= 0 || !(ratio >= 0.0 && ratio <= 1.0)) {
        if (errbuf && errlen) snprintf(errbuf, errlen, "insider_b200_split: bad argument");
        return INSIDER_ERR_INVALID_ARG;
    }
    const int64_t total = N * P;
    std::vector<int64_t> existing;                       // 0-based column-major linear indices of non-NA entries (R/utils.R:90)
    existing.reserve((size_t)total);
    for (int64_t i = 0; i < total; ++i) {
        const bool isna = std::isnan(data[i]);
        if (na) na[i] = isna ? 1 : 0;
        train[i] = isna ? 0 : 1;
        test[i] = 0;
        if (!isna) existing.push_back(i);
    }
    const int64_t n = (int64_t)existing.size();
    const int64_t k = (int64_t)std::floor((double)n * ratio);   // :91
    RMersenne rng(seed);                                  // :89 set.seed(seed)
    int64_t drawn = 0;
    if (n > 1 && k > 0) {
        if ((double)n > 1e7 && k <= n / 2) {
            // sample.int useHash branch -> do_sample2: redraw (at most 100 times) while the value was already taken
            std::vector<uint64_t> seen((size_t)(n + 63) / 64, 0);
            for (int64_t i = 0; i < k; ++i) {
                int64_t v = 0;
                for (int j = 0; j < 100; ++j) {
                    v = (int64_t)rng.unif_index((double)n);
                    if (!((seen[(size_t)v >> 6] >> (v & 63)) & 1ull)) break;
                }
                seen[(size_t)v >> 6] |= 1ull << (v & 63);
                const int64_t idx = existing[(size_t)v];
                if (!test[idx]) { test[idx] = 1; train[idx] = 0; ++drawn; }
            }
        } else {
            // classic partial Fisher-Yates of do_sample: x[j] = x[--n]
            std::vector<int64_t> x(existing);
            int64_t nn = n;
            for (int64_t i = 0; i < k; ++i) {
                const int64_t j = (int64_t)rng.unif_index((double)nn);
                const int64_t idx = x[(size_t)j];
                x[(size_t)j] = x[(size_t)--nn];
                test[idx] = 1; train[idx] = 0; ++drawn;
            }
        }
    }
    if (n_test) *n_test = drawn;
    return INSIDER_OK;
}
