// Host driver of libinsider_b200: context, resident problems, ALS sessions, the C ABI of include/insider_b200.h.
//
// The outer loop follows src/optimize.cpp:256-422 step for step (initial evaluation, gram, Gauss-Seidel over
// confounder blocks, row-factor rebuild, column update, every-10th-iteration evaluation with the decay ladder and
// the relative-decrease stopping rule) but keeps everything device-resident: the host only launches kernels and,
// on check iterations, reads back one small record to decide whether to stop.
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/insider_b200.h"
#include "common.cuh"
#include "kernels.cuh"

using namespace ib;

namespace {

// ---- minimal NCCL binding (dlopen: the single-GPU path has no NCCL dependency) --------------------------------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
struct NcclApi {
    void* h = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool load(std::string& err) {
        if (h) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
        if (!h) { err = std::string("cannot dlopen libnccl.so.2: ") + dlerror(); return false; }
#define SYM(f, n) *(void**)(&f) = dlsym(h, n); if (!f) { err = std::string("missing NCCL symbol ") + n; return false; }
        SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
        SYM(AllReduce, "ncclAllReduce") SYM(Broadcast, "ncclBroadcast") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
        return true;
    }
};
NcclApi g_nccl;
constexpr int NCCL_FLOAT64 = 8, NCCL_SUM = 0;

void set_err(char* buf, size_t len, const char* fmt, ...) {
    if (!buf || len == 0) return;
    va_list ap; va_start(ap, fmt); vsnprintf(buf, len, fmt, ap); va_end(ap);
}

struct Err { int code; std::string msg; };
#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) throw Err{INSIDER_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)}; } while (0)
#define REQUIRE(c, msg) do { if (!(c)) throw Err{INSIDER_ERR_INVALID_ARG, msg}; } while (0)

// Stream-ordered allocations from the device's default memory pool (release threshold raised at context creation), so
// that the 51 back-to-back fits of tune() reuse their buffers instead of paying cudaMalloc/cudaFree every time.
struct DevPool {                       // frees everything it handed out
    std::vector<void*> ptrs;
    cudaStream_t stream = nullptr;
    template <typename T> T* get(size_t n, bool zero = true, cudaStream_t st = 0) {
        if (n == 0) n = 1;
        if (st) stream = st;
        T* p = nullptr;
        cudaError_t e = cudaMallocAsync((void**)&p, n * sizeof(T), stream);
        if (e != cudaSuccess) { cudaGetLastError(); throw Err{INSIDER_ERR_NOMEM, std::string("cudaMallocAsync failed: ") + cudaGetErrorString(e)}; }
        ptrs.push_back(p);
        if (zero) cudaMemsetAsync(p, 0, n * sizeof(T), stream);
        return p;
    }
    ~DevPool() { for (void* p : ptrs) cudaFreeAsync(p, stream); }
};

}  // namespace

struct insider_ctx {
    int device = 0, sm_count = 148, rank = 0, world = 1;
    cudaStream_t stream = nullptr;
    cudaStream_t side = nullptr;            // second stream: small independent kernels of an iteration run beside the main chain
    bool no_side = false;                   // INSIDER_B200_NO_SIDE_STREAM=1: everything on the main stream (debugging)
    ncclComm_t comm = nullptr;
    bool profile = false;
    unsigned char* perm_table = nullptr;    // rank tables of the counter-based permutation source (common.cuh)
    // pinned bounce buffers for uploads from PAGEABLE host memory (what R passes): see h2d_staged()
    void* pin[2] = {nullptr, nullptr};
    cudaEvent_t pin_ev[2] = {nullptr, nullptr};
    bool pin_busy[2] = {false, false};
};

struct insider_resident {
    insider_ctx* ctx = nullptr;
    int64_t N = 0, P = 0, j0 = 0, Pl = 0, Pl_pad = 0;     // global genes P; this rank owns [j0, j0 + Pl)
    int C = 0, Q = 0, inc_continuous = 0, has_masks = 0;
    int ldY = 0, ldT = 0, Wp = 0, WPr = 0;
    double *Y = nullptr, *X = nullptr;
    uint32_t *trC = nullptr, *teC = nullptr, *trR = nullptr;
    std::vector<int> L;
    std::vector<int*> level_of_row, rows_sorted, level_start;
    std::vector<std::vector<int>> level_start_host, lor_host;
    // dense-path Gauss-Seidel tables (design only): co-occurrence CSR over all levels, per-level sums of X
    int total_levels = 0;
    int *gs_lvl_first = nullptr, *gs_co_ptr = nullptr, *gs_co_row = nullptr;
    int *gs_row_lv = nullptr;               // [N][C] global level index of every (sample, confounder)
    int *lv_ptr = nullptr, *lv_rows = nullptr;   // CSR over all levels: the rows of each level, ascending ([total_levels + 1], [C * N])
    int gs_nnz = 0;
    double *gs_co_cnt = nullptr, *gs_Sx = nullptr;
    double n_train = 0, n_test = 0;
    double h2d_bytes = 0;
    DevPool pool;
};

struct ProfEntry { std::string name; cudaEvent_t e0, e1; };

struct insider_session {
    insider_ctx* ctx = nullptr;
    insider_resident* r = nullptr;
    insider_options opt{};
    Geom g{};
    bool masked = false;
    int n_factors = 0;
    std::vector<int> frows;
    std::vector<size_t> a_off;          // offset (doubles) of factor c inside A_all
    size_t n_A = 0;
    double *V = nullptr, *Xty = nullptr, *A_all = nullptr, *U = nullptr, *Ut = nullptr, *UtU = nullptr;
    double *stats = nullptr, *B = nullptr, *G = nullptr, *D = nullptr;   // stats = [B | G | D] (one all-reduce)
    size_t stats_elems = 0;
    double *Bp = nullptr, *Gp = nullptr, *Dp = nullptr, *GLp = nullptr, *T = nullptr, *cont_scratch = nullptr, *sse_part = nullptr;
    double* XtXall = nullptr;           // masked ridge path (alpha = 0): per-gene Gram matrices [P_l][KP*KP]
    double* gram_tiles = nullptr;       // masked elastic net: per-gene lower triangles in 32-gene slot tiles (k_cd_masked.cu)
    bool masked_v6 = true;              // masked elastic net: k_col_gram + k_cd_persistent (8 lanes per gene, dynamic gene queue). The
                                        // thread-per-gene tile solver of k_cd_masked.cu (INSIDER_B200_MASKED_TILES=1) is faster at steady
                                        // state (0.94 against 1.14 ms per iteration) but slower in the first, unordered iterations that
                                        // dominate a 31-iteration tune() fit (profiles/r02_masked_solver_versions.txt)
    unsigned int* queue = nullptr;      // gene queue of the persistent CD kernel
    double* Vfull = nullptr;            // world > 1: gathered V for the final download
    double* Vpack = nullptr;            // world > 1: the same without pitch (K x P contiguous)
    std::vector<int> lfac_base;         // first level-table index of each confounder
    int total_levels = 0, max_chunks = 1;
    LevelTable* tab_dev = nullptr;
    double* Lfac = nullptr;             // [total_levels][KP*KP + KP] Cholesky factors + inverse diagonals
    double* SB = nullptr;               // dense path: [total_levels][KP] per-level sums of B
    // dense fast chain (single slab, no continuous covariates, factors fit the cluster kernel's shared memory): k_row_b stores
    // per-level sums, V V' is its own kernel right after the column update (so the level factorisations run beside the pass
    // over Y), the row-factor rebuild is fused into the Gauss-Seidel cluster kernel
    bool dense_fast = false;
    double* SBp = nullptr;              // [rb_splits][total_levels][KP]
    double* gv_parts = nullptr;         // k_gram_v block partials
    unsigned int* gv_counter = nullptr;
    int rb_splits = 1, d_splits = 1, stream_blocks = 1;
    RowDesign* designs_dev = nullptr;
    std::vector<RowDesign> designs;
    CheckState* state = nullptr;
    insider_check* records_dev = nullptr;
    unsigned long long* sweeps_dev = nullptr;
    int* sweeps_gene = nullptr;          // dense CD: sweeps of every local gene in the last iteration
    int* cd_order = nullptr;             // dense CD: slot -> gene, sorted by the last iteration's sweep counts
    int* cd_order_work = nullptr;
    double* cd_table = nullptr;          // dense CD: XtX table prepared once per column update
    double* cd_tables_all = nullptr;     // dense CD: the table in every visiting order (iterations with many sweeps per gene)
    CdPhaseState cd_phase{};             // dense CD: parked genes between the phases of a column update
    int* cd_order_parked = nullptr;      // dense CD: slot order of a later phase
    uint32_t cd_phase0 = 512;            // sweeps of the first phase (doubling afterwards)
    int* err_dev = nullptr;
    uint32_t max_records = 0, n_records = 0;
    uint32_t iter = 0;
    bool done = false;
    insider_check last{};
    std::vector<insider_check> records;
    double loop_ms = 0, h2d = 0, d2h = 0;
    int64_t launches = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_fork[2] = {nullptr, nullptr}, ev_join[2] = {nullptr, nullptr};   // fork/join of the side stream
    // one ALS iteration (run_iteration + k_bump_iter), replayed. Three variants of the dense elastic-net solve: [0] the first
    // iterations (thousands of CD sweeps per gene, no useful prediction of the counts: phases with re-grouping), [1] the next
    // ones (hundreds of sweeps, counts known from the previous iteration: one launch, fastest sweep), [2] afterwards (a few
    // sweeps per gene: all blocks resident at once)
    cudaGraphExec_t iter_graphs[3] = {nullptr, nullptr, nullptr};
    int64_t launches_per_iter_v[3] = {0, 0, 0};
    int graph_variant = 0;
    bool graph_failed = false;
    std::vector<ProfEntry> prof;
    std::map<std::string, std::pair<double, int64_t>> prof_acc;
    DevPool pool;
};

namespace {

struct Launch {                         // counts launches and optionally brackets them with events
    insider_session* s; const char* name; cudaEvent_t e0 = nullptr, e1 = nullptr;
    Launch(insider_session* s_, const char* n, int count = 1) : s(s_), name(n) {
        s->launches += count;
        if (s->ctx->profile) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, s->ctx->stream); }
    }
    ~Launch() { if (e0) { cudaEventRecord(e1, s->ctx->stream); s->prof.push_back({name, e0, e1}); } }
};

void drain_profile(insider_session* s) {
    for (auto& p : s->prof) {
        cudaEventSynchronize(p.e1);
        float ms = 0; cudaEventElapsedTime(&ms, p.e0, p.e1);
        auto& acc = s->prof_acc[p.name]; acc.first += ms; acc.second += 1;
        cudaEventDestroy(p.e0); cudaEventDestroy(p.e1);
    }
    s->prof.clear();
}

void nccl_check(int rc, const char* what) {
    if (rc != 0) throw Err{INSIDER_ERR_NCCL, std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "nccl error")};
}

void split_genes(int64_t P, int world, int rank, int64_t& j0, int64_t& n) {
    // contiguous blocks, multiples of 32 genes so the row-major mask words never straddle ranks
    const int64_t words = (P + 31) / 32;
    const int64_t q = words / world, r = words % world;
    const int64_t w0 = rank * q + std::min<int64_t>(rank, r), w1 = w0 + q + (rank < r ? 1 : 0);
    j0 = std::min(P, w0 * 32);
    n = std::min(P, w1 * 32) - j0;
}

// ---- upload -------------------------------------------------------------------------------------------------
// Host-to-device copy that does not depend on where the caller's buffer lives. From pinned memory: one cudaMemcpyAsync. From
// pageable memory (R's matrices) the driver stages every copy through its own small pinned buffer at 9-13 GB/s; here the source is
// copied by 4 host threads into two 16 MB pinned bounce buffers whose DMA transfers overlap the next chunk's memcpy
// (142 MB of Y: 15.7 -> ~7 ms, bench.py `e2e_pageable`).
constexpr size_t PIN_BYTES = (size_t)16 << 20;
void h2d_staged(insider_ctx* ctx, void* dst, const void* src, size_t bytes, cudaStream_t st) {
    cudaPointerAttributes at{};
    const bool pageable = bytes >= ((size_t)4 << 20) && cudaPointerGetAttributes(&at, src) == cudaSuccess && at.type == cudaMemoryTypeUnregistered;
    cudaGetLastError();
    if (!pageable) { CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st)); return; }
    for (int b = 0; b < 2; ++b) {
        if (!ctx->pin[b]) { CUDA_TRY(cudaHostAlloc(&ctx->pin[b], PIN_BYTES, cudaHostAllocDefault)); CUDA_TRY(cudaEventCreateWithFlags(&ctx->pin_ev[b], cudaEventDisableTiming)); }
    }
    int b = 0;
    for (size_t off = 0; off < bytes; off += PIN_BYTES, b ^= 1) {
        const size_t n = std::min(PIN_BYTES, bytes - off);
        if (ctx->pin_busy[b]) { CUDA_TRY(cudaEventSynchronize(ctx->pin_ev[b])); ctx->pin_busy[b] = false; }   // its last DMA has read the buffer
        constexpr int T = 4;
        std::thread th[T - 1];
        const size_t part = (n / T + 63) & ~(size_t)63;
        auto piece = [&](int t) {
            const size_t o = std::min(n, (size_t)t * part), e = std::min(n, o + part);
            if (e > o) memcpy((char*)ctx->pin[b] + o, (const char*)src + off + o, e - o);
        };
        bool spawned[T - 1] = {false, false, false};
        for (int t = 1; t < T; ++t) {
            try { th[t - 1] = std::thread(piece, t); spawned[t - 1] = true; }
            catch (...) { piece(t); }                                          // no thread to be had: copy this piece here
        }
        piece(0);
        for (int t = 1; t < T; ++t) if (spawned[t - 1]) th[t - 1].join();
        CUDA_TRY(cudaMemcpyAsync((char*)dst + off, ctx->pin[b], n, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaEventRecord(ctx->pin_ev[b], st));
        ctx->pin_busy[b] = true;
    }
}

insider_resident* do_upload(insider_ctx* ctx, const insider_problem* pb) {
    REQUIRE(pb && pb->Y && pb->N > 0 && pb->P > 0, "problem: Y, N, P required");
    REQUIRE(pb->C >= 0 && (pb->C == 0 || pb->levels), "problem: levels required when C > 0");
    REQUIRE(pb->inc_continuous == 0 || pb->inc_continuous == 1, "The value of parameter inc_continuous can only be 0 or 1.");
    REQUIRE(pb->inc_continuous == 0 || (pb->X && pb->Q > 0), "inc_continuous = 1 needs X and Q > 0");
    REQUIRE(pb->C + pb->inc_continuous > 0, "at least one confounder block is required");
    REQUIRE(pb->mask_kind >= INSIDER_MASK_NONE && pb->mask_kind <= INSIDER_MASK_DOUBLE, "bad mask_kind");
    REQUIRE(pb->mask_kind == INSIDER_MASK_NONE || (pb->train && pb->test), "train and test masks required unless mask_kind = NONE");
    REQUIRE(pb->N < (1 << 30), "N too large");
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    auto* r = new insider_resident();
    try {
        r->pool.stream = st;
        r->ctx = ctx; r->N = pb->N; r->P = pb->P; r->C = pb->C; r->Q = pb->inc_continuous ? pb->Q : 0; r->inc_continuous = pb->inc_continuous;
        r->has_masks = pb->mask_kind != INSIDER_MASK_NONE;
        split_genes(r->P, ctx->world, ctx->rank, r->j0, r->Pl);
        r->Pl_pad = std::max<int64_t>(TG, (r->Pl + TG - 1) / TG * TG);
        const int N = (int)r->N;
        r->ldY = pitch4(N);
        r->ldT = pitch4(round_up(N, 8));
        r->Wp = round_up((r->ldT + 31) / 32, 4);
        r->WPr = (int)((r->Pl_pad + 31) / 32);
        // Y: column block [j0, j0+Pl) into pitch ldY, zero padded (+ slack for tile overhang reads)
        const size_t y_elems = (size_t)r->Pl_pad * r->ldY + 64;
        r->Y = r->pool.get<double>(y_elems, true, st);
        if (r->Pl > 0) {
            // contiguous H2D into a staging buffer in column chunks, re-pitched on the device
            // (32 MB staging: a large one-off allocation would cost more in memory-pool growth than the copy itself)
            const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(r->Pl, (int64_t)((size_t)32 << 20) / ((size_t)N * 8)));
            double* stage = nullptr;
            CUDA_TRY(cudaMallocAsync((void**)&stage, (size_t)chunk * N * 8, st));
            for (int64_t c0 = 0; c0 < r->Pl; c0 += chunk) {
                const int64_t n = std::min(chunk, r->Pl - c0);
                try { h2d_staged(ctx, stage, pb->Y + (size_t)(r->j0 + c0) * N, (size_t)n * N * 8, st); }
                catch (...) { cudaFreeAsync(stage, st); throw; }
                launch_repitch(r->Y + (size_t)c0 * r->ldY, r->ldY, stage, N, N, n, st);
            }
            cudaFreeAsync(stage, st);
        }
        r->h2d_bytes += (double)r->Pl * N * 8;
        // design
        for (int c = 0; c < r->C; ++c) {
            const int32_t* z = pb->levels + (size_t)c * N;
            int L = 0;
            for (int k = 0; k < N; ++k) { REQUIRE(z[k] >= 1, "levels must be 1-based positive integers"); L = std::max(L, (int)z[k]); }
            std::vector<int> cnt(L + 1, 0), lor(N), sorted(N), start(L + 1, 0);
            for (int k = 0; k < N; ++k) { lor[k] = z[k] - 1; cnt[lor[k]]++; }
            for (int l = 0; l < L; ++l) { REQUIRE(cnt[l] > 0, "each confounder column must use every level 1..L_c (reference indexes row level-1)"); start[l + 1] = start[l] + cnt[l]; }
            std::vector<int> fill(start.begin(), start.end() - 1);
            for (int k = 0; k < N; ++k) sorted[fill[lor[k]]++] = k;
            int* d_lor = r->pool.get<int>(N, false); int* d_sorted = r->pool.get<int>(N, false); int* d_start = r->pool.get<int>(L + 1, false);
            CUDA_TRY(cudaMemcpyAsync(d_lor, lor.data(), N * sizeof(int), cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(d_sorted, sorted.data(), N * sizeof(int), cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(d_start, start.data(), (L + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaStreamSynchronize(st));   // host vectors go out of scope
            r->level_start_host.push_back(start); r->lor_host.push_back(lor);
            r->L.push_back(L); r->level_of_row.push_back(d_lor); r->rows_sorted.push_back(d_sorted); r->level_start.push_back(d_start);
            r->h2d_bytes += (2.0 * N + L + 1) * 4;
        }
        if (r->inc_continuous) {
            r->X = r->pool.get<double>((size_t)N * r->Q, false);
            CUDA_TRY(cudaMemcpyAsync(r->X, pb->X, (size_t)N * r->Q * 8, cudaMemcpyHostToDevice, st));
            r->h2d_bytes += (double)N * r->Q * 8;
        }
        // dense-path tables: for every level (c, s) the co-occurring levels (c' != c, s') with their sample counts
        {
            std::vector<int> first(r->C + 1, 0);
            for (int c = 0; c < r->C; ++c) first[c + 1] = first[c] + r->L[c];
            r->total_levels = first[r->C];
            std::vector<int> ptr(r->total_levels + 1, 0), rows;
            std::vector<double> cnts, Sx((size_t)std::max(1, r->total_levels) * std::max(1, r->Q), 0.0);
            for (int c = 0; c < r->C; ++c)
                for (int l = 0; l < r->L[c]; ++l) {
                    std::map<int, int> co;
                    const int b = r->level_start_host[c][l], e = r->level_start_host[c][l + 1];
                    std::vector<int> members;
                    for (int k = 0; k < N; ++k) if (r->lor_host[c][k] == l) members.push_back(k);
                    (void)b; (void)e;
                    for (int k : members) {
                        for (int c2 = 0; c2 < r->C; ++c2) if (c2 != c) co[first[c2] + r->lor_host[c2][k]] += 1;
                        for (int q = 0; q < r->Q; ++q) Sx[(size_t)(first[c] + l) * r->Q + q] += pb->X[(size_t)q * N + k];
                    }
                    for (auto& kv : co) { rows.push_back(kv.first); cnts.push_back((double)kv.second); }
                    ptr[first[c] + l + 1] = (int)rows.size();
                }
            {
                std::vector<int> row_lv((size_t)std::max(1, N * r->C)), lptr(r->total_levels + 1, 0), lrows((size_t)std::max(1, N * r->C));
                for (int c = 0; c < r->C; ++c) {
                    for (int k = 0; k < N; ++k) row_lv[(size_t)k * r->C + c] = first[c] + r->lor_host[c][k];
                    std::vector<int> fill(r->L[c]);
                    for (int l = 0; l < r->L[c]; ++l) { lptr[first[c] + l] = c * N + r->level_start_host[c][l]; fill[l] = lptr[first[c] + l]; }
                    for (int k = 0; k < N; ++k) lrows[fill[r->lor_host[c][k]]++] = k;
                }
                lptr[r->total_levels] = N * r->C;
                r->gs_row_lv = r->pool.get<int>(row_lv.size(), false);
                r->lv_ptr = r->pool.get<int>(lptr.size(), false);
                r->lv_rows = r->pool.get<int>(lrows.size(), false);
                CUDA_TRY(cudaMemcpyAsync(r->gs_row_lv, row_lv.data(), row_lv.size() * 4, cudaMemcpyHostToDevice, st));
                CUDA_TRY(cudaMemcpyAsync(r->lv_ptr, lptr.data(), lptr.size() * 4, cudaMemcpyHostToDevice, st));
                CUDA_TRY(cudaMemcpyAsync(r->lv_rows, lrows.data(), lrows.size() * 4, cudaMemcpyHostToDevice, st));
                CUDA_TRY(cudaStreamSynchronize(st));
            }
            r->gs_nnz = (int)rows.size();
            r->gs_lvl_first = r->pool.get<int>(first.size(), false);
            r->gs_co_ptr = r->pool.get<int>(ptr.size(), false);
            r->gs_co_row = r->pool.get<int>(std::max<size_t>(1, rows.size()), false);
            r->gs_co_cnt = r->pool.get<double>(std::max<size_t>(1, cnts.size()), false);
            r->gs_Sx = r->pool.get<double>(Sx.size(), false);
            CUDA_TRY(cudaMemcpyAsync(r->gs_lvl_first, first.data(), first.size() * 4, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(r->gs_co_ptr, ptr.data(), ptr.size() * 4, cudaMemcpyHostToDevice, st));
            if (!rows.empty()) {
                CUDA_TRY(cudaMemcpyAsync(r->gs_co_row, rows.data(), rows.size() * 4, cudaMemcpyHostToDevice, st));
                CUDA_TRY(cudaMemcpyAsync(r->gs_co_cnt, cnts.data(), cnts.size() * 8, cudaMemcpyHostToDevice, st));
            }
            CUDA_TRY(cudaMemcpyAsync(r->gs_Sx, Sx.data(), Sx.size() * 8, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaStreamSynchronize(st));
        }
        // masks -> bit planes
        if (r->has_masks) {
            const size_t words = (size_t)r->Pl_pad * r->Wp;
            r->trC = r->pool.get<uint32_t>(words, true, st);
            r->teC = r->pool.get<uint32_t>(words, true, st);
            r->trR = r->pool.get<uint32_t>((size_t)N * r->WPr, true, st);
            const size_t esz = pb->mask_kind == INSIDER_MASK_INT32 ? 4 : pb->mask_kind == INSIDER_MASK_UINT8 ? 1 : 8;
            const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(r->Pl, (int64_t)((size_t)512 << 20) / (esz * N)));
            void* tmp = nullptr;
            CUDA_TRY(cudaMallocAsync(&tmp, (size_t)chunk * N * esz, st));
            try {
                for (int pass = 0; pass < 2; ++pass) {
                    const char* src = (const char*)(pass == 0 ? pb->train : pb->test);
                    uint32_t* dst = pass == 0 ? r->trC : r->teC;
                    for (int64_t c0 = 0; c0 < r->Pl; c0 += chunk) {
                        const int64_t n = std::min(chunk, r->Pl - c0);
                        h2d_staged(ctx, tmp, src + (size_t)(r->j0 + c0) * N * esz, (size_t)n * N * esz, st);
                        launch_pack_mask(tmp, pb->mask_kind, N, n, r->Wp, dst + (size_t)c0 * r->Wp, st);
                        r->h2d_bytes += (double)n * N * esz;
                    }
                }
                launch_transpose_mask(r->trC, N, r->Pl_pad, r->Wp, r->WPr, r->trR, st);
                unsigned long long* cnt = r->pool.get<unsigned long long>(2, true, st);
                launch_count_bits(r->trC, (int64_t)words, cnt, st);
                launch_count_bits(r->teC, (int64_t)words, cnt + 1, st);
                unsigned long long h[2];
                CUDA_TRY(cudaMemcpyAsync(h, cnt, 16, cudaMemcpyDeviceToHost, st));
                CUDA_TRY(cudaStreamSynchronize(st));
                r->n_train = (double)h[0]; r->n_test = (double)h[1];
            } catch (...) { cudaFreeAsync(tmp, st); throw; }
            cudaFreeAsync(tmp, st);
            if (ctx->world > 1) {
                double* d2 = r->pool.get<double>(2, false);
                double hv[2] = {r->n_train, r->n_test};
                CUDA_TRY(cudaMemcpyAsync(d2, hv, 16, cudaMemcpyHostToDevice, st));
                nccl_check(g_nccl.AllReduce(d2, d2, 2, NCCL_FLOAT64, NCCL_SUM, ctx->comm, st), "ncclAllReduce(counts)");
                CUDA_TRY(cudaMemcpyAsync(hv, d2, 16, cudaMemcpyDeviceToHost, st));
                CUDA_TRY(cudaStreamSynchronize(st));
                r->n_train = hv[0]; r->n_test = hv[1];
            }
        }
        CUDA_TRY(cudaStreamSynchronize(st));
        CUDA_TRY(cudaGetLastError());
    } catch (...) { delete r; throw; }
    return r;
}

// ---- factor marshalling (host column-major L x K  <->  device row-major L x KP) -----------------------------------
void upload_factors(insider_session* s, const insider_factors* f) {
    const int K = s->g.K, KP = s->g.KP;
    std::vector<double> tmp(s->n_A, 0.0);
    for (int c = 0; c < s->n_factors; ++c) {
        const int L = s->frows[c];
        const double* src = f->factors[c];
        for (int l = 0; l < L; ++l) for (int k = 0; k < K; ++k) tmp[s->a_off[c] + (size_t)l * KP + k] = src[l + (size_t)k * L];
    }
    CUDA_TRY(cudaMemcpyAsync(s->A_all, tmp.data(), s->n_A * 8, cudaMemcpyHostToDevice, s->ctx->stream));
    const insider_resident* r = s->r;
    if (r->Pl > 0) {                       // contiguous copy into the (not yet used) Xty buffer, re-pitched on the device
        CUDA_TRY(cudaMemcpyAsync(s->Xty, f->column_factor + (size_t)r->j0 * K, (size_t)r->Pl * K * 8, cudaMemcpyHostToDevice, s->ctx->stream));
        launch_repitch(s->V, s->g.ldV, s->Xty, K, K, r->Pl, s->ctx->stream);
    }
    CUDA_TRY(cudaStreamSynchronize(s->ctx->stream));
    s->h2d += (double)s->n_A * 8 + (double)r->Pl * K * 8;
}

void download_factors(insider_session* s, const insider_factors* f) {
    const int K = s->g.K, KP = s->g.KP;
    cudaStream_t st = s->ctx->stream;
    std::vector<double> tmp(s->n_A);
    CUDA_TRY(cudaMemcpyAsync(tmp.data(), s->A_all, s->n_A * 8, cudaMemcpyDeviceToHost, st));
    const insider_resident* r = s->r;
    if (s->ctx->world == 1) {
        // pack V (pitch ldV) into the Xty buffer (free between iterations: k_col_xty rewrites it) and download it contiguously
        launch_repitch(s->Xty, K, s->V, s->g.ldV, K, r->Pl, st);
        CUDA_TRY(cudaMemcpyAsync(f->column_factor, s->Xty, (size_t)r->Pl * K * 8, cudaMemcpyDeviceToHost, st));
    } else {
        // gather every rank's gene block with one broadcast per rank, then one download of the full K x P matrix
        for (int rk = 0; rk < s->ctx->world; ++rk) {
            int64_t j0, n; split_genes(r->P, s->ctx->world, rk, j0, n);
            if (n == 0) continue;
            nccl_check(g_nccl.Broadcast(s->V, s->Vfull + (size_t)j0 * s->g.ldV, (size_t)n * s->g.ldV, NCCL_FLOAT64, rk, s->ctx->comm, st), "ncclBroadcast(V)");
        }
        launch_repitch(s->Vpack, K, s->Vfull, s->g.ldV, K, r->P, st);          // contiguous K x P, one download
        CUDA_TRY(cudaMemcpyAsync(f->column_factor, s->Vpack, (size_t)r->P * K * 8, cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    for (int c = 0; c < s->n_factors; ++c) {
        const int L = s->frows[c];
        double* dst = f->factors[c];
        for (int l = 0; l < L; ++l) for (int k = 0; k < K; ++k) dst[l + (size_t)k * L] = tmp[s->a_off[c] + (size_t)l * KP + k];
    }
    s->d2h += (double)s->n_A * 8 + (double)(s->ctx->world == 1 ? r->Pl : r->P) * K * 8;
}

// ---- evaluation (src/optimize.cpp:320-323 and :381-408) -----------------------------------------------------------
void evaluate(insider_session* s, bool initial) {
    cudaStream_t st = s->ctx->stream;
    insider_resident* r = s->r;
    { Launch l(s, "k_sse"); launch_sse(s->g, s->masked, r->Y, r->trC, r->teC, s->Ut, s->V, s->sse_part, s->stream_blocks, st); }
    { Launch l(s, "k_sse_reduce"); launch_sse_reduce(s->sse_part, s->stream_blocks, s->state, st); }
    if (s->ctx->world > 1) nccl_check(g_nccl.AllReduce(&s->state->sse_train, &s->state->sse_train, 4, NCCL_FLOAT64, NCCL_SUM, s->ctx->comm, st), "ncclAllReduce(sse)");
    insider_check* rec = s->records_dev + std::min(s->n_records, s->max_records - 1);
    { Launch l(s, "k_check"); launch_check(s->state, s->A_all, (int64_t)s->n_A, initial ? 1 : 0, (int)s->iter, rec, st); }
    struct { insider_check rec; } hostrec;
    CheckState hs;
    CUDA_TRY(cudaMemcpyAsync(&hostrec.rec, rec, sizeof(insider_check), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(&hs, s->state, sizeof(CheckState), cudaMemcpyDeviceToHost, st));
    int err = 0;
    CUDA_TRY(cudaMemcpyAsync(&err, s->err_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    s->d2h += sizeof(insider_check) + sizeof(CheckState) + 4;
    s->last = hostrec.rec;
    s->records.push_back(hostrec.rec);
    s->n_records++;
    if (s->opt.verbose && s->ctx->rank == 0) {
        // the reference's console lines (src/optimize.cpp:327-329,387; src/utils.cpp:70-76,95-100)
        const int it = initial ? 0 : hostrec.rec.iter;
        printf("insider iter %d: train rmse = %.12g\n", it, hostrec.rec.train_rmse);
        if (s->opt.tuning == 1) printf("insider iter %d: test rmse = %.12g\n", it, hostrec.rec.test_rmse);
        printf("total_residual\t%.12g;\nrow_reg_loss:\t%.12g;\ncol_reg_loss:\t%.12g;\nl1_reg_loss:\t%.12g.\n", hostrec.rec.sum_residual / 2,
               hostrec.rec.row_reg, hostrec.rec.col_reg, hostrec.rec.l1_reg);
        if (!initial) printf("Delta loss for iter %d:%.12g\n", it, hostrec.rec.delta_loss);
        fflush(stdout);
    }
    if (err) throw Err{INSIDER_ERR_NOT_SPD, "a normal-equation matrix was not positive definite"};
    if (hs.diverged) throw Err{INSIDER_ERR_DIVERGED, "loss is NaN or Inf"};
    if (!initial && hs.converged) s->done = true;
}

// ---- one ALS iteration (src/optimize.cpp:331-378) -------------------------------------------------------------------
// Side-stream sections: small kernels that do not depend on each other run beside the main chain (also inside the captured
// graph, where they become parallel branches). With per-kernel profiling on everything stays on the main stream.
void run_column_update(insider_session* s);

struct SideSection {
    insider_session* s; int idx; cudaStream_t side;
    SideSection(insider_session* s_, int i) : s(s_), idx(i), side((s_->ctx->profile || s_->ctx->no_side) ? s_->ctx->stream : s_->ctx->side) {
        if (side != s->ctx->stream) { cudaEventRecord(s->ev_fork[idx], s->ctx->stream); cudaStreamWaitEvent(side, s->ev_fork[idx], 0); }
    }
    void join() { if (side != s->ctx->stream) { cudaEventRecord(s->ev_join[idx], side); cudaStreamWaitEvent(s->ctx->stream, s->ev_join[idx], 0); } }
};

// ALS iterations whose masked elastic-net solves still take hundreds of sweeps per gene (k_cd.cu: 4 lanes per gene until then)
constexpr uint32_t CD_LONG_SOLVE_ITERS = 6;

bool run_iteration(insider_session* s) {
    cudaStream_t st = s->ctx->stream;
    insider_resident* r = s->r;
    const Geom& g = s->g;
    const int KK = g.KP * g.KP;
    if (s->dense_fast) {
        // Two independent chains meet at the Gauss-Seidel kernel: [G = V V' (:332) -> factorisation of every level's normal
        // equations] needs V only, [pass over Y -> per-level sums of B] is the long one: they run side by side (one GPU). With
        // several GPUs SB and G are adjacent and travel in ONE all-reduce, after which the factorisations follow on the chain.
        SideSection sec0(s, 0);
        // the device-side iteration counter (permutation keys of the column update) is bumped HERE, beside the pass over Y, for the
        // iteration that just ended: nothing before the join below reads it (6 us off the dependent chain of every iteration)
        if (s->iter > 0) { Launch l(s, "k_bump_iter"); launch_bump_iter(s->state, sec0.side); }
        { Launch l(s, "k_gram_v"); launch_gram_v(g, s->V, s->gv_parts, s->G, s->gv_counter, nullptr, sec0.side); }
        if (s->ctx->world == 1) { Launch l(s, "k_level_factor"); launch_level_factor(g, false, s->tab_dev, s->total_levels, s->max_chunks, s->G, s->GLp, s->opt.lambda1, s->Lfac, s->err_dev, sec0.side); }
        { Launch l(s, "k_row_b"); launch_row_b_ex(g, false, r->Y, nullptr, s->V, s->SBp, nullptr, s->rb_splits, r->lv_ptr, r->lv_rows, s->total_levels, g.N * r->C, st); }
        { Launch l(s, "k_reduce"); launch_reduce_partials(s->SB, s->SBp, (int64_t)s->total_levels * g.KP, s->rb_splits, st); }
        sec0.join();
        if (s->ctx->world > 1) {
            nccl_check(g_nccl.AllReduce(s->SB, s->SB, (size_t)s->total_levels * g.KP + KK, NCCL_FLOAT64, NCCL_SUM, s->ctx->comm, st), "ncclAllReduce(SB|G)");
            Launch l(s, "k_level_factor"); launch_level_factor(g, false, s->tab_dev, s->total_levels, s->max_chunks, s->G, s->GLp, s->opt.lambda1, s->Lfac, s->err_dev, st);
        }
        DenseGs dg{r->gs_lvl_first, r->gs_co_ptr, r->gs_co_row, r->gs_co_cnt, r->gs_Sx};
        { Launch l(s, "k_rows_dense_gs"); launch_rows_dense_gs_ex(g, dg, r->C, r->Q, s->total_levels, *std::max_element(r->L.begin(), r->L.end()), r->gs_nnz, s->A_all, nullptr, s->SB, s->G, s->Lfac, r->gs_row_lv, s->U, s->Ut, s->UtU, st); }
        run_column_update(s);
        return true;
    }
    // sufficient statistics of the row update: G = V V' (:332), B = (M o Y) V', D_k = complement Grams
    { Launch l(s, "k_row_b"); launch_row_b(g, s->masked, r->Y, r->trC, s->V, s->Bp, s->Gp, s->rb_splits, st); }
    if (s->masked) { Launch l(s, "k_row_comp_gram"); launch_row_comp_gram(g, r->trR, s->V, s->Dp, s->d_splits, st); }
    {
        double* outs[3] = {s->B, s->G, s->D};
        const double* parts[3] = {s->Bp, s->Gp, s->Dp};
        const int64_t ns[3] = {(int64_t)g.N * g.KP, KK, (int64_t)g.N * KK};
        const int nps[3] = {s->rb_splits, s->rb_splits * ROW_B_GRAM_PARTS, s->d_splits};
        Launch l(s, "k_reduce");
        launch_reduce_jobs(s->masked ? 3 : 2, outs, parts, ns, nps, st);
    }
    if (s->ctx->world > 1) nccl_check(g_nccl.AllReduce(s->stats, s->stats, s->stats_elems, NCCL_FLOAT64, NCCL_SUM, s->ctx->comm, st), "ncclAllReduce(stats)");
    // normal-equation matrices of every level of every confounder: assembled and factorised once (they do not depend on A)
    SideSection sec0(s, 0);                // dense path: the per-level sums of B run beside the factorisations
    if (r->C > 0) {
        if (s->masked) { Launch l(s, "k_level_gram"); launch_level_gram(g, s->tab_dev, s->total_levels, s->max_chunks, s->G, s->D, s->GLp, st); }
        if (!s->masked) { Launch l(s, "k_level_sumB"); launch_level_sumB(g, s->tab_dev, s->total_levels, s->B, s->SB, sec0.side); }
        { Launch l(s, "k_level_factor"); launch_level_factor(g, s->masked, s->tab_dev, s->total_levels, s->max_chunks, s->G, s->GLp, s->opt.lambda1, s->Lfac, s->err_dev, st); }
    }
    sec0.join();
    // Gauss-Seidel over confounder blocks (:335-362)
    if (s->masked) {
        for (int c = 0; c < r->C; ++c) {
            const RowDesign& d = s->designs[c];
            { Launch l(s, "k_row_rhs"); launch_row_rhs(g, d, s->B, s->G, s->D, s->U, s->T, st); }
            { Launch l(s, "k_level_update"); launch_level_update(g, true, d, s->lfac_base[c], s->G, s->B, s->T, s->Lfac, s->U, st); }
        }
    } else if (r->C > 0) {
        // dense path: per-level sums of B once, then all C block updates in one single-block launch on factor-sized data
        DenseGs dg{r->gs_lvl_first, r->gs_co_ptr, r->gs_co_row, r->gs_co_cnt, r->gs_Sx};
        { Launch l(s, "k_rows_dense_gs"); launch_rows_dense_gs(g, dg, r->C, r->Q, s->total_levels, r->L.empty() ? 0 : *std::max_element(r->L.begin(), r->L.end()), r->gs_nnz, s->A_all, r->inc_continuous ? s->A_all + s->a_off[r->C] : nullptr, s->SB, s->G, s->Lfac, st); }
        if (r->inc_continuous) {   // the continuous block below works on the row factor: rebuild it with the new A_c
            Launch l(s, "k_build_u", 2);
            launch_build_u(g, r->C, s->designs_dev, r->Q, r->X, s->A_all + s->a_off[r->C], s->U, s->Ut, s->UtU, st);
        }
    }
    if (r->inc_continuous) {
        double* W = s->A_all + s->a_off[r->C];
        for (int q = 0; q < r->Q; ++q) {
            Launch l(s, "k_continuous", 2);
            launch_continuous(g, s->masked, r->X + (size_t)q * g.N, W + (size_t)q * g.KP, s->B, s->G, s->D, s->opt.lambda1, s->U, s->cont_scratch, s->err_dev, st);
        }
    }
    // row factor rebuild (:365-373) and column update (:376)
    { Launch l(s, "k_build_u", 2); launch_build_u(g, r->C, s->designs_dev, r->Q, r->X, r->inc_continuous ? s->A_all + s->a_off[r->C] : nullptr, s->U, s->Ut, s->UtU, st); }
    run_column_update(s);
    return false;
}

// column update (:376): Xty = U'(M o Y), then ridge / elastic net per gene
void run_column_update(insider_session* s) {
    cudaStream_t st = s->ctx->stream;
    insider_resident* r = s->r;
    const Geom& g = s->g;
    static const bool dense_group = getenv("INSIDER_B200_DENSE_GROUP") != nullptr;   // A/B: 8-lanes-per-gene solver on the dense path
    const bool dense_cd = !s->masked && s->opt.alpha != 0.0 && !dense_group;
    SideSection sec1(s, 1);                // dense elastic net: slot order and XtX table (neither needs Xty) beside the pass over Y
    if (dense_cd) {
        { Launch l(s, "k_cd_order"); launch_cd_order(s->sweeps_gene, g.P, s->cd_order, s->cd_order_work, &s->state->als_iter, sec1.side); }
        { Launch l(s, "k_cd_table"); launch_cd_dense_table(g.K, s->UtU, g.KP, 1, s->opt.lambda2, s->opt.alpha, s->cd_table, sec1.side); }
        if (s->cd_tables_all && s->graph_variant < 2) {   // thousands to tens of sweeps per gene: 22 MB of tables once, no per-sweep build
            Launch l(s, "k_cd_tables_all"); launch_cd_dense_tables_all(g.K, s->cd_table, s->ctx->perm_table, s->cd_tables_all, sec1.side);
        }
    }
    const bool masked_cd = s->masked && s->opt.alpha != 0.0;
    const bool masked_tiles = masked_cd && !s->masked_v6;
    if (masked_tiles) { Launch l(s, "k_cd_order"); launch_cd_order(s->sweeps_gene, g.P, s->cd_order, s->cd_order_work, &s->state->als_iter, sec1.side); }
    { Launch l(s, "k_col_xty"); launch_col_xty(g, s->masked, r->Y, r->trC, s->Ut, s->Xty, s->stream_blocks, st); }
    sec1.join();
    CdParams p{s->opt.lambda2, s->opt.alpha, &s->state->tol, &s->state->als_iter, s->opt.seed, s->opt.perm_mode};
    if (masked_tiles) {
        // per-gene matrices straight into the solver's tile layout (slot order), then thread-per-gene coordinate descent
        { Launch l(s, "k_col_gram_tiles"); launch_col_gram_tiles(g, r->trC, s->U, s->UtU, s->cd_order, s->opt.lambda2, s->opt.alpha, s->gram_tiles, st); }
        { Launch l(s, "k_cd_masked"); launch_cd_masked(g, s->gram_tiles, s->Xty, s->V, p, s->sweeps_dev, s->sweeps_dev + 1, s->sweeps_gene, s->cd_order, s->ctx->perm_table, st); }
        return;
    }
    if (s->masked) { Launch l(s, "k_col_gram"); launch_col_gram(g, r->trC, s->U, s->UtU, s->XtXall, st); }
    if (dense_cd && s->graph_variant == 0) {
        // first iterations (hundreds to thousands of sweeps per gene, counts spread 10x): phases of doubling length. A warp runs
        // until its slowest gene is done, so after every phase the unconverged genes are re-grouped by how far they still are
        // from the stopping rule: lane efficiency 0.47 -> 0.8 in iteration 0 (profiles/r01_cd_phases.txt). Launches whose
        // phase has no gene left exit at once.
        constexpr int N_PHASES = 9;
        uint32_t draw0 = 0, cap = s->cd_phase0;
        for (int ph = 0; ph < N_PHASES; ++ph) {
            const bool last = ph == N_PHASES - 1;
            if (ph > 0) { Launch l(s, "k_cd_order"); launch_cd_order_parked(s->cd_phase, g.P, s->cd_order_parked, s->cd_order_work, &s->state->tol, st); }
            Launch l(s, "k_cd_dense");
            launch_cd_dense(g, s->UtU, s->Xty, s->V, p, s->sweeps_dev, s->sweeps_dev + 1, s->sweeps_gene, ph ? s->cd_order_parked : s->cd_order,
                            s->ctx->perm_table, s->cd_table, false, &s->cd_phase, draw0, last ? 0xffffffffu : cap, s->cd_tables_all, st);
            draw0 = cap; cap += std::max<uint32_t>(1u, cap / 2);                       // 512, 768, 1152, ... x1.5
        }
    } else if (dense_cd) {
        Launch l(s, "k_cd_dense");
        launch_cd_dense(g, s->UtU, s->Xty, s->V, p, s->sweeps_dev, s->sweeps_dev + 1, s->sweeps_gene, s->cd_order, s->ctx->perm_table, s->cd_table,
                        s->graph_variant == 2, nullptr, 0u, 0xffffffffu, s->graph_variant < 2 ? s->cd_tables_all : nullptr, st);
    } else {
        const bool group_cd = s->opt.alpha != 0.0;
        if (group_cd) { Launch l(s, "k_cd_order"); launch_cd_order(s->sweeps_gene, g.P, s->cd_order, s->cd_order_work, &s->state->als_iter, st); }
        Launch l(s, "k_col_solve");
        launch_col_solve(g, s->masked, s->UtU, s->XtXall, s->Xty, s->V, p, s->sweeps_dev, s->sweeps_dev + 1, s->queue, s->ctx->perm_table, s->err_dev, s->ctx->sm_count,
                         group_cd ? s->sweeps_gene : nullptr, group_cd ? s->cd_order : nullptr, s->iter < CD_LONG_SOLVE_ITERS, st);
    }
}

insider_session* do_begin(insider_ctx* ctx, insider_resident* r, const insider_factors* f, const insider_options* o) {
    REQUIRE(r && f && o, "null argument");
    // a resident problem is read-only after upload: contexts on the same device may share it (tune() replicas on one GPU)
    REQUIRE(r->ctx == ctx || (r->ctx->device == ctx->device && r->ctx->world == 1 && ctx->world == 1), "resident problem belongs to a context on another device");
    REQUIRE(o->tuning == 0 || o->tuning == 1, "Parameter tuning should be either 0 or 1!");
    if (f->K < 1 || f->K > KMAX) throw Err{INSIDER_ERR_UNSUPPORTED, "latent_dim = " + std::to_string(f->K) + " is not supported: libinsider_b200 handles latent_dim 1..32"};
    REQUIRE(f->n_factors == r->C + r->inc_continuous, "n_factors must equal C + inc_continuous");
    REQUIRE(f->factors && f->factor_rows && f->column_factor, "factor buffers required");
    REQUIRE(o->tuning == 0 || r->has_masks, "tuning = 1 needs train/test masks");
    for (int c = 0; c < r->C; ++c) REQUIRE(f->factor_rows[c] == r->L[c], "factor rows must equal the number of levels of the confounder");
    if (r->inc_continuous) REQUIRE(f->factor_rows[r->C] == r->Q, "continuous factor must be Q x K");
    if (o->tuning == 1 && r->n_test == 0) throw Err{INSIDER_ERR_EMPTY_TEST_SET, "tuning = 1 but the test indicator is empty"};
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    auto* s = new insider_session();
    try {
        s->pool.stream = st;
        s->ctx = ctx; s->r = r; s->opt = *o;
        if (s->opt.check_every == 0) s->opt.check_every = 10;
        if (s->opt.perm_mode == 0) s->opt.perm_mode = INSIDER_PERM_COUNTER;
        REQUIRE(s->opt.perm_mode == INSIDER_PERM_COUNTER || s->opt.perm_mode == INSIDER_PERM_IDENTITY, "bad perm_mode");
        s->masked = o->tuning == 1;
        Geom& g = s->g;
        g.N = (int)r->N; g.K = f->K; g.KP = round_up(f->K, 8); g.NT = g.KP / 8;
        g.ldY = r->ldY; g.ldV = pitch4(g.KP); g.ldT = r->ldT; g.Wp = r->Wp; g.WPr = r->WPr;
        g.P = r->Pl; g.P_pad = r->Pl_pad; g.n_tiles = (int)(r->Pl_pad / TG); g.gene0 = r->j0;
        s->n_factors = f->n_factors;
        size_t off = 0;
        for (int c = 0; c < f->n_factors; ++c) { s->frows.push_back(f->factor_rows[c]); s->a_off.push_back(off); off += (size_t)f->factor_rows[c] * g.KP; }
        s->n_A = off;
        const int KK = g.KP * g.KP;
        s->V = s->pool.get<double>((size_t)g.P_pad * g.ldV + 64, true, st);
        s->Xty = s->pool.get<double>((size_t)g.P_pad * g.ldV + 64, true, st);
        s->A_all = s->pool.get<double>(s->n_A, true, st);
        s->U = s->pool.get<double>((size_t)g.N * g.KP, true, st);
        s->Ut = s->pool.get<double>((size_t)g.KP * g.ldT + 64, true, st);
        s->UtU = s->pool.get<double>((size_t)KK * (1 + build_u_parts(g)), true, st);
        s->stats_elems = (size_t)g.N * g.KP + KK + (s->masked ? (size_t)g.N * KK : 0);
        s->stats = s->pool.get<double>(s->stats_elems, true, st);
        s->B = s->stats; s->G = s->B + (size_t)g.N * g.KP; s->D = s->masked ? s->G + KK : nullptr;
        s->rb_splits = row_b_default_splits(g, ctx->sm_count);
        s->Bp = s->pool.get<double>(row_b_partial_elems(g, s->rb_splits), true, st);
        s->Gp = s->pool.get<double>((size_t)s->rb_splits * ROW_B_GRAM_PARTS * KK, true, st);
        s->stream_blocks = stream_default_blocks(g, ctx->sm_count);
        s->sse_part = s->pool.get<double>((size_t)s->stream_blocks * 4, true, st);
        s->T = s->pool.get<double>((size_t)g.N * g.KP, true, st);
        // level table over all confounders
        {
            std::vector<LevelTable> tab;
            for (int c = 0; c < r->C; ++c) {
                s->lfac_base.push_back((int)tab.size());
                for (int l = 0; l < r->L[c]; ++l) {
                    const int b = r->level_start_host[c][l], e = r->level_start_host[c][l + 1];
                    tab.push_back(LevelTable{r->rows_sorted[c], b, e});
                    s->max_chunks = std::max(s->max_chunks, (e - b + 31) / 32);
                }
            }
            s->total_levels = (int)tab.size();
            s->tab_dev = s->pool.get<LevelTable>(std::max<size_t>(1, tab.size()), true, st);
            if (!tab.empty()) CUDA_TRY(cudaMemcpyAsync(s->tab_dev, tab.data(), tab.size() * sizeof(LevelTable), cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            s->Lfac = s->pool.get<double>((size_t)std::max(1, s->total_levels) * (KK + g.KP), true, st);
            s->SB = s->pool.get<double>((size_t)std::max(1, s->total_levels) * g.KP, true, st);
        }
        if (s->masked) {
            s->d_splits = std::max(1, std::min(16, (ctx->sm_count * 16 + g.N - 1) / g.N));
            s->d_splits = std::min(s->d_splits, std::max(1, g.WPr));
            s->Dp = s->pool.get<double>((size_t)s->d_splits * g.N * KK, true, st);
            s->GLp = s->pool.get<double>((size_t)std::max(1, s->total_levels) * s->max_chunks * KK, true, st);
            { const char* e = getenv("INSIDER_B200_MASKED_TILES"); s->masked_v6 = !(e && e[0] == '1'); }
            if (o->alpha == 0.0 || s->masked_v6) s->XtXall = s->pool.get<double>((size_t)std::max<int64_t>(1, g.P) * KK, true, st);
            else s->gram_tiles = s->pool.get<double>(cd_masked_tile_doubles(g.K, std::max<int64_t>(1, g.P)), false, st);
        } else {
            s->GLp = s->pool.get<double>(1, true, st);
        }
        if (r->inc_continuous) s->cont_scratch = s->pool.get<double>(continuous_scratch_elems(g), true, st);
        s->dense_fast = !s->masked && r->C > 0 && !r->inc_continuous && row_b_levels_supported(g, s->total_levels, g.N * r->C) && !getenv("INSIDER_B200_NO_FAST_CHAIN") &&
                        rows_dense_gs_can_fuse_u(g, 0, s->total_levels, *std::max_element(r->L.begin(), r->L.end()), r->gs_nnz);
        if (s->dense_fast) {
            s->SB = s->pool.get<double>((size_t)s->total_levels * g.KP + KK, true, st);     // [SB | G]: one all-reduce
            s->G = s->SB + (size_t)s->total_levels * g.KP;
            s->SBp = s->pool.get<double>((size_t)s->rb_splits * s->total_levels * g.KP, true, st);
            s->gv_parts = s->pool.get<double>((size_t)gram_v_parts(std::max<int64_t>(1, g.P)) * KK, true, st);
            s->gv_counter = s->pool.get<unsigned int>(1, true, st);
        }
        if (ctx->world > 1) { s->Vfull = s->pool.get<double>((size_t)r->P * g.ldV, true, st); s->Vpack = s->pool.get<double>((size_t)r->P * f->K, false, st); }
        for (int c = 0; c < r->C; ++c) s->designs.push_back(RowDesign{r->L[c], r->level_of_row[c], r->rows_sorted[c], r->level_start[c], s->A_all + s->a_off[c]});
        s->designs_dev = s->pool.get<RowDesign>(std::max(1, r->C), true, st);
        if (r->C) CUDA_TRY(cudaMemcpyAsync(s->designs_dev, s->designs.data(), r->C * sizeof(RowDesign), cudaMemcpyHostToDevice, st));
        s->state = s->pool.get<CheckState>(1, true, st);
        CheckState hs{};
        hs.sub_tol = o->sub_tol; hs.global_tol = o->global_tol; hs.tol = o->sub_tol; hs.decay = 1.0;
        hs.n_train = r->n_train; hs.n_test = r->n_test; hs.np_total = (double)r->N * (double)r->P;
        hs.lambda1 = o->lambda1; hs.lambda2 = o->lambda2; hs.alpha = o->alpha; hs.tuning = o->tuning; hs.als_iter = 0;
        CUDA_TRY(cudaMemcpyAsync(s->state, &hs, sizeof(hs), cudaMemcpyHostToDevice, st));
        s->max_records = (uint32_t)std::min<uint64_t>((uint64_t)o->max_iter / s->opt.check_every + 3, 1u << 20);
        s->records_dev = s->pool.get<insider_check>(s->max_records, true, st);
        s->sweeps_dev = s->pool.get<unsigned long long>(2, true, st);   // [sweeps, coordinate steps]
        s->queue = s->pool.get<unsigned int>(1, true, st);
        s->sweeps_gene = s->pool.get<int>((size_t)std::max<int64_t>(1, g.P), true, st);
        s->cd_order = s->pool.get<int>((size_t)std::max<int64_t>(1, g.P), true, st);
        s->cd_order_work = s->pool.get<int>(cd_order_work_ints(), true, st);
        s->cd_table = s->pool.get<double>(cd_dense_table_elems(), true, st);
        if (!s->masked && s->opt.alpha != 0.0 && s->opt.perm_mode == INSIDER_PERM_COUNTER && !getenv("INSIDER_B200_CD_NO_TMA"))
            s->cd_tables_all = s->pool.get<double>(cd_dense_tables_all_elems(g.K), false, st);
        {
            const size_t Pn = (size_t)std::max<int64_t>(1, g.P);
            s->cd_phase.state = s->pool.get<double>(Pn * 2 * (size_t)round_up(g.K, 4), false, st);
            s->cd_phase.inc = s->pool.get<uint32_t>(Pn, true, st);
            s->cd_phase.dl = s->pool.get<float>(Pn, true, st);
            s->cd_phase.alive = s->pool.get<int>(Pn, true, st);
            s->cd_phase.n_slots = s->pool.get<int>(1, true, st);
            s->cd_order_parked = s->pool.get<int>(Pn, true, st);
            if (const char* e = getenv("INSIDER_B200_CD_PHASE0")) { const long v = atol(e); if (v >= 1) s->cd_phase0 = (uint32_t)v; }
        }
        s->err_dev = s->pool.get<int>(1, true, st);
        CUDA_TRY(cudaEventCreate(&s->ev0)); CUDA_TRY(cudaEventCreate(&s->ev1));
        for (int i = 0; i < 2; ++i) {
            CUDA_TRY(cudaEventCreateWithFlags(&s->ev_fork[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&s->ev_join[i], cudaEventDisableTiming));
        }
        upload_factors(s, f);
        // initial row factor and evaluation (:286-289, :320-323)
        { Launch l(s, "k_build_u", 2); launch_build_u(g, r->C, s->designs_dev, r->Q, r->X, r->inc_continuous ? s->A_all + s->a_off[r->C] : nullptr, s->U, s->Ut, s->UtU, st); }
        evaluate(s, true);
    } catch (...) { if (s->ev0) cudaEventDestroy(s->ev0); if (s->ev1) cudaEventDestroy(s->ev1); delete s; throw; }
    return s;
}

// ALS iteration from which the dense solver runs with all its blocks resident (see k_cd_dense.cu): by then a gene needs tens of
// sweeps, not thousands
constexpr uint32_t CD_VARIANT_SWITCH_ITER = 24;
// ALS iterations that run the dense solver in phases (measured on the ageing-shaped fit: iteration 0 69 -> 44 ms, 1 and 2
// -10 %, from iteration 3 on the previous iteration's counts order the work better than a mid-solve re-grouping)
constexpr uint32_t CD_PHASED_ITERS = 3;

// one iteration + iteration-counter bump, replayed as a CUDA graph unless per-kernel profiling is on
void launch_iteration(insider_session* s) {
    cudaStream_t st = s->ctx->stream;
    const int v = (s->iter >= CD_VARIANT_SWITCH_ITER) ? 2 : (s->iter >= CD_PHASED_ITERS ? 1 : 0);
    s->graph_variant = v;
    // Graph replay pays off at steady state (a dozen short dependent kernels per iteration). The first iterations are a few
    // multi-millisecond solver launches: captured, their instantiation sat inside iteration 0 (measured 48.4 ms against 39.9 ms
    // with plain launches) - they are launched directly, and the steady-state graph is built while the GPU still works on them.
    const bool want_graph = s->opt.use_graph >= 0 && !s->ctx->profile && !s->graph_failed && (v == 2 || s->opt.use_graph > 0);
    if (!want_graph) {
        if (!run_iteration(s)) { Launch l(s, "k_bump_iter"); launch_bump_iter(s->state, st); }
        return;
    }
    if (!s->iter_graphs[v]) {
        const int64_t before = s->launches;
        cudaGraph_t graph = nullptr;
        CUDA_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
        try {
            if (!run_iteration(s)) { launch_bump_iter(s->state, st); s->launches += 1; }
        } catch (...) { cudaStreamEndCapture(st, &graph); if (graph) cudaGraphDestroy(graph); throw; }
        cudaError_t e = cudaStreamEndCapture(st, &graph);
        s->launches_per_iter_v[v] = s->launches - before;
        s->launches = before;
        if (e == cudaSuccess) e = cudaGraphInstantiate(&s->iter_graphs[v], graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (e != cudaSuccess) {                       // fall back to plain launches (still the CUDA path)
            cudaGetLastError(); s->iter_graphs[v] = nullptr; s->graph_failed = true;
            if (!run_iteration(s)) { Launch l(s, "k_bump_iter"); launch_bump_iter(s->state, st); }
            return;
        }
    }
    CUDA_TRY(cudaGraphLaunch(s->iter_graphs[v], st));
    s->launches += s->launches_per_iter_v[v];
}

void do_step(insider_session* s, uint32_t n_iters, int32_t* done, double* ms) {
    CUDA_TRY(cudaSetDevice(s->ctx->device));
    cudaStream_t st = s->ctx->stream;
    CUDA_TRY(cudaEventRecord(s->ev0, st));
    uint32_t ran = 0;
    while (!s->done && ran < n_iters) {
        if (s->iter > s->opt.max_iter) { s->done = true; break; }                  // while(iter <= max_iter)  :325
        // the device-side iteration counter (permutation keys) is bumped at the end of the iteration; the evaluation of a
        // check iteration therefore receives the iteration index from the host
        launch_iteration(s);
        if (s->iter % s->opt.check_every == 0) evaluate(s, false);                 // :381
        if (s->done) break;                                                        // :405-407 (iter not incremented on break)
        s->iter++; ran++;                                                          // :409
        if (s->iter > s->opt.max_iter) s->done = true;
    }
    CUDA_TRY(cudaEventRecord(s->ev1, st));
    CUDA_TRY(cudaEventSynchronize(s->ev1));
    CUDA_TRY(cudaGetLastError());
    float t = 0; CUDA_TRY(cudaEventElapsedTime(&t, s->ev0, s->ev1));
    s->loop_ms += t;
    if (ms) *ms = t;
    if (done) *done = s->done ? 1 : 0;
    if (s->ctx->profile) drain_profile(s);
}

void fill_result(insider_session* s, insider_result* res) {
    if (!res) return;
    unsigned long long sw2[2] = {0, 0}; int err = 0;
    CUDA_TRY(cudaMemcpy(sw2, s->sweeps_dev, 16, cudaMemcpyDeviceToHost));
    const unsigned long long sw = sw2[0];
    CUDA_TRY(cudaMemcpy(&err, s->err_dev, 4, cudaMemcpyDeviceToHost));
    res->train_rmse = s->last.train_rmse; res->test_rmse = s->last.test_rmse; res->loss = s->last.loss;
    res->iters_run = s->iter;
    res->n_checks = (uint32_t)std::min<size_t>(s->records.size(), res->checks ? res->max_checks : s->records.size());
    if (res->checks) for (uint32_t i = 0; i < res->n_checks; ++i) res->checks[i] = s->records[i];
    res->cd_sweeps = (int64_t)sw; res->cd_steps = (int64_t)sw2[1]; res->loop_ms = s->loop_ms; res->h2d_bytes = s->h2d; res->d2h_bytes = s->d2h; res->kernel_launches = s->launches;
    if (err) throw Err{INSIDER_ERR_NOT_SPD, "a normal-equation matrix was not positive definite"};
}

void destroy_session(insider_session* s) {
    if (!s) return;
    for (auto& p : s->prof) { cudaEventDestroy(p.e0); cudaEventDestroy(p.e1); }
    for (auto& ge : s->iter_graphs) if (ge) cudaGraphExecDestroy(ge);
    if (s->ev0) cudaEventDestroy(s->ev0);
    for (int i = 0; i < 2; ++i) { if (s->ev_fork[i]) cudaEventDestroy(s->ev_fork[i]); if (s->ev_join[i]) cudaEventDestroy(s->ev_join[i]); }
    if (s->ev1) cudaEventDestroy(s->ev1);
    delete s;
}

template <typename F> int guarded(char* errbuf, size_t errlen, F&& fn) {
    try { fn(); return INSIDER_OK; }
    catch (const Err& e) { set_err(errbuf, errlen, "%s", e.msg.c_str()); cudaGetLastError(); return e.code; }
    catch (const std::bad_alloc&) { set_err(errbuf, errlen, "host out of memory"); return INSIDER_ERR_NOMEM; }
    catch (const std::exception& e) { set_err(errbuf, errlen, "%s", e.what()); return INSIDER_ERR_INVALID_ARG; }
    catch (...) { set_err(errbuf, errlen, "unknown error"); return INSIDER_ERR_INVALID_ARG; }
}

int create_ctx(insider_ctx** out, int device, int rank, int world, const void* id, char* errbuf, size_t errlen) {
    return guarded(errbuf, errlen, [&] {
        REQUIRE(out, "null out pointer");
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0) throw Err{INSIDER_ERR_CUDA, "no CUDA device available (libinsider_b200 has no CPU fallback)"};
        REQUIRE(device >= 0 && device < n, "bad device index");
        cudaDeviceProp prop; CUDA_TRY(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10) throw Err{INSIDER_ERR_CUDA, "libinsider_b200 is built for sm_100a (B200) only"};
        CUDA_TRY(cudaSetDevice(device));
        auto* c = new insider_ctx();
        c->device = device; c->sm_count = prop.multiProcessorCount; c->rank = rank; c->world = world;
        try {
            CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
            CUDA_TRY(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
            { const char* e = getenv("INSIDER_B200_NO_SIDE_STREAM"); c->no_side = e && e[0] == '1'; }
            {
                std::vector<unsigned char> tab(PERM_TABLE_BYTES);
                build_perm_table(tab.data());
                CUDA_TRY(cudaMalloc(&c->perm_table, tab.size()));
                CUDA_TRY(cudaMemcpy(c->perm_table, tab.data(), tab.size(), cudaMemcpyHostToDevice));
            }
            {
                cudaMemPool_t mp = nullptr;
                if (cudaDeviceGetDefaultMemPool(&mp, device) == cudaSuccess) { uint64_t thr = UINT64_MAX; cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &thr); }
            }
            if (world > 1) {
                std::string err;
                if (!g_nccl.load(err)) throw Err{INSIDER_ERR_NCCL, err};
                ncclUniqueId uid; memcpy(&uid, id, sizeof(uid));
                nccl_check(g_nccl.CommInitRank(&c->comm, world, uid, rank), "ncclCommInitRank");
            }
        } catch (...) { if (c->perm_table) cudaFree(c->perm_table); if (c->side) cudaStreamDestroy(c->side); if (c->stream) cudaStreamDestroy(c->stream); delete c; throw; }
        *out = c;
    });
}

}  // namespace

// ================================================================================================================
extern "C" {

int insider_b200_version(void) { return INSIDER_B200_VERSION; }

void insider_b200_default_options(insider_options* o) {
    if (!o) return;
    memset(o, 0, sizeof(*o));
    o->lambda1 = 1.0; o->lambda2 = 1.0; o->alpha = 0.1; o->tuning = 1;             // src/optimize.cpp:257
    o->global_tol = 1e-10; o->sub_tol = 1e-5; o->max_iter = 10000; o->check_every = 10;
    o->perm_mode = INSIDER_PERM_COUNTER; o->seed = 0; o->verbose = 0; o->use_graph = 0;
}

int insider_b200_ctx_create(insider_ctx** out, int device, char* errbuf, size_t errlen) { return create_ctx(out, device, 0, 1, nullptr, errbuf, errlen); }

int insider_b200_nccl_unique_id(void* out128, char* errbuf, size_t errlen) {
    return guarded(errbuf, errlen, [&] {
        REQUIRE(out128, "null buffer");
        std::string err;
        if (!g_nccl.load(err)) throw Err{INSIDER_ERR_NCCL, err};
        ncclUniqueId uid; nccl_check(g_nccl.GetUniqueId(&uid), "ncclGetUniqueId");
        memcpy(out128, &uid, sizeof(uid));
    });
}

int insider_b200_ctx_create_dist(insider_ctx** out, int device, int rank, int world, const void* nccl_id128, char* errbuf, size_t errlen) {
    if (world < 1 || rank < 0 || rank >= world || (world > 1 && !nccl_id128)) { set_err(errbuf, errlen, "bad rank/world/nccl id"); return INSIDER_ERR_INVALID_ARG; }
    return create_ctx(out, device, rank, world, nccl_id128, errbuf, errlen);
}

void insider_b200_ctx_destroy(insider_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    if (c->perm_table) cudaFree(c->perm_table);
    for (int b = 0; b < 2; ++b) { if (c->pin_ev[b]) cudaEventDestroy(c->pin_ev[b]); if (c->pin[b]) cudaFreeHost(c->pin[b]); }
    if (c->side) cudaStreamDestroy(c->side);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

void* insider_b200_ctx_stream(insider_ctx* c) { return c ? (void*)c->stream : nullptr; }
void insider_b200_set_profile(insider_ctx* c, int on) { if (c) c->profile = on != 0; }

int insider_b200_upload(insider_ctx* ctx, const insider_problem* prob, insider_resident** out, char* errbuf, size_t errlen) {
    return guarded(errbuf, errlen, [&] { REQUIRE(ctx && out, "null argument"); *out = do_upload(ctx, prob); });
}
void insider_b200_release(insider_resident* r) { if (r) { cudaSetDevice(r->ctx->device); delete r; } }

int insider_b200_als_begin(insider_ctx* ctx, insider_resident* r, const insider_factors* fac, const insider_options* opt, insider_session** out,
                           char* errbuf, size_t errlen) {
    return guarded(errbuf, errlen, [&] { REQUIRE(ctx && out, "null argument"); *out = do_begin(ctx, r, fac, opt); });
}
int insider_b200_als_step(insider_session* s, uint32_t n_iters, int32_t* done, double* ms, char* errbuf, size_t errlen) {
    return guarded(errbuf, errlen, [&] { REQUIRE(s, "null session"); do_step(s, n_iters, done, ms); });
}
int insider_b200_als_read(insider_session* s, const insider_factors* fac, char* errbuf, size_t errlen) {
    return guarded(errbuf, errlen, [&] { REQUIRE(s && fac, "null argument"); download_factors(s, fac); });
}
int insider_b200_als_end(insider_session* s, const insider_factors* fac, insider_result* res, char* errbuf, size_t errlen) {
    int rc = guarded(errbuf, errlen, [&] {
        REQUIRE(s, "null session");
        if (fac) download_factors(s, fac);
        fill_result(s, res);
    });
    destroy_session(s);
    return rc;
}
int insider_b200_als_profile(insider_session* s, char* names, size_t names_len, double* ms, int64_t* calls, int max_entries) {
    if (!s) return 0;
    std::string joined; int n = 0;
    for (auto& kv : s->prof_acc) {
        if (n >= max_entries) break;
        if (n) joined += "\n";
        joined += kv.first;
        if (ms) ms[n] = kv.second.first;
        if (calls) calls[n] = kv.second.second;
        ++n;
    }
    if (names && names_len) { strncpy(names, joined.c_str(), names_len - 1); names[names_len - 1] = 0; }
    return n;
}

int64_t insider_b200_als_sweeps(insider_session* s, int32_t* out, int64_t n) {
    if (!s || !out || n <= 0 || s->opt.alpha == 0.0) return 0;
    n = std::min<int64_t>(n, s->g.P);
    if (cudaSetDevice(s->ctx->device) != cudaSuccess) return 0;
    if (cudaMemcpyAsync(out, s->sweeps_gene, (size_t)n * 4, cudaMemcpyDeviceToHost, s->ctx->stream) != cudaSuccess) return 0;
    if (cudaStreamSynchronize(s->ctx->stream) != cudaSuccess) return 0;
    return n;
}

int64_t insider_b200_als_hint_sweeps(insider_session* s, const int32_t* hint, int64_t n) {
    if (!s || !hint || n <= 0 || s->opt.alpha == 0.0) return 0;
    n = std::min<int64_t>(n, s->g.P);
    if (cudaSetDevice(s->ctx->device) != cudaSuccess) return 0;
    if (cudaMemcpyAsync(s->sweeps_gene, hint, (size_t)n * 4, cudaMemcpyHostToDevice, s->ctx->stream) != cudaSuccess) return 0;
    if (cudaStreamSynchronize(s->ctx->stream) != cudaSuccess) return 0;
    return n;
}

int insider_b200_optimize_resident(insider_ctx* ctx, insider_resident* r, const insider_factors* fac, const insider_options* opt, insider_result* res,
                                   char* errbuf, size_t errlen) {
    insider_session* s = nullptr;
    int rc = insider_b200_als_begin(ctx, r, fac, opt, &s, errbuf, errlen);
    if (rc) return rc;
    int32_t done = 0;
    while (!done) {
        rc = insider_b200_als_step(s, 1000, &done, nullptr, errbuf, errlen);
        if (rc) { destroy_session(s); return rc; }
    }
    return insider_b200_als_end(s, fac, res, errbuf, errlen);
}

int insider_b200_optimize(insider_ctx* ctx, const insider_problem* prob, const insider_factors* fac, const insider_options* opt, insider_result* res,
                          char* errbuf, size_t errlen) {
    insider_resident* r = nullptr;
    if (fac && (fac->K < 1 || fac->K > KMAX)) {                         // before anything is uploaded
        set_err(errbuf, errlen, "latent_dim = %d is not supported: libinsider_b200 handles latent_dim 1..%d", fac->K, KMAX);
        return INSIDER_ERR_UNSUPPORTED;
    }
    const bool trace = getenv("INSIDER_B200_TRACE") != nullptr;         // phase wall times of the one-shot call on stderr
    const auto t0 = std::chrono::steady_clock::now();
    int rc = insider_b200_upload(ctx, prob, &r, errbuf, errlen);
    if (rc) return rc;
    const auto t1 = std::chrono::steady_clock::now();
    const double up = r->h2d_bytes;
    rc = insider_b200_optimize_resident(ctx, r, fac, opt, res, errbuf, errlen);
    if (rc == INSIDER_OK && res) res->h2d_bytes += up;
    const auto t2 = std::chrono::steady_clock::now();
    insider_b200_release(r);
    if (trace) {
        const auto t3 = std::chrono::steady_clock::now();
        auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        fprintf(stderr, "[insider_b200 rank %d] optimize: upload %.2f ms, fit %.2f ms (device loop %.2f ms), release %.2f ms\n", ctx->rank, ms(t0, t1),
                ms(t1, t2), res ? res->loop_ms : 0.0, ms(t2, t3));
    }
    return rc;
}

int insider_b200_tune_batch(int32_t n_ctx, insider_ctx* const* ctxs, insider_resident* const* residents, int32_t n_points, const insider_factors* fac,
                            const insider_options* opt, insider_result* res, int32_t* point_ctx, char* errbuf, size_t errlen) {
    if (n_ctx < 1 || !ctxs || !residents || !fac || !opt || !res || n_points < 0) { set_err(errbuf, errlen, "null argument"); return INSIDER_ERR_INVALID_ARG; }
    for (int c = 0; c < n_ctx; ++c)
        if (!ctxs[c] || !residents[c] || ctxs[c]->world != 1) { set_err(errbuf, errlen, "tune_batch needs plain single-GPU contexts, one resident problem each"); return INSIDER_ERR_INVALID_ARG; }
    // replicas: one host thread per context pulls grid points from a shared counter (no communication between fits)
    std::atomic<int> next{0};
    std::atomic<int> first_rc{INSIDER_OK};
    std::mutex err_mu;
    std::string err_msg;
    auto worker = [&](int c) {
        char local[512];
        while (first_rc.load() == INSIDER_OK) {
            const int i = next.fetch_add(1);
            if (i >= n_points) break;
            local[0] = 0;
            const int rc = insider_b200_optimize_resident(ctxs[c], residents[c], &fac[i], &opt[i], &res[i], local, sizeof local);
            if (point_ctx) point_ctx[i] = c;
            if (rc != INSIDER_OK) {
                int expected = INSIDER_OK;
                if (first_rc.compare_exchange_strong(expected, rc)) { std::lock_guard<std::mutex> lk(err_mu); err_msg = "grid point " + std::to_string(i) + ": " + local; }
            }
        }
    };
    if (n_ctx == 1) worker(0);
    else {
        std::vector<std::thread> th;
        for (int c = 0; c < n_ctx; ++c) th.emplace_back(worker, c);
        for (auto& t : th) t.join();
    }
    if (first_rc.load() != INSIDER_OK) set_err(errbuf, errlen, "%s", err_msg.c_str());
    return first_rc.load();
}

int insider_b200_optimize_continuous(insider_ctx* ctx, int64_t N, int64_t P, int32_t K, const double* data, int32_t mask_kind, const void* indicator,
                                     double* updating_factor, const double* c_factor, const double* updating_confd, double lambda, int32_t tuning,
                                     char* errbuf, size_t errlen) {
    // The problem "one continuous covariate, nothing else": row factor u_k = x_k w. Statistics B, G (, D) of `data`, then the
    // K x K update of k_cont_partial / k_cont_final - the same launches the continuous block of a full fit runs (run_iteration).
    return guarded(errbuf, errlen, [&] {
        REQUIRE(ctx && data && updating_factor && c_factor && updating_confd, "null argument");
        REQUIRE(tuning == 0 || tuning == 1, "Parameter tuning should be either 0 or 1!");            // src/optimize.cpp:133-135
        REQUIRE(tuning == 0 || (indicator && mask_kind != INSIDER_MASK_NONE), "tuning = 1 needs the indicator matrix");
        REQUIRE(ctx->world == 1, "optimize_continuous runs on a single-GPU context");
        if (K < 1 || K > KMAX) throw Err{INSIDER_ERR_UNSUPPORTED, "latent_dim must be in 1..32"};
        insider_problem pb{};
        pb.N = N; pb.P = P; pb.C = 0; pb.Q = 1; pb.inc_continuous = 1; pb.Y = data; pb.X = updating_confd;
        pb.mask_kind = tuning == 1 ? mask_kind : INSIDER_MASK_NONE; pb.train = indicator; pb.test = indicator;
        std::unique_ptr<insider_resident, void (*)(insider_resident*)> r(do_upload(ctx, &pb), insider_b200_release);
        std::vector<double> V((size_t)K * P);
        memcpy(V.data(), c_factor, V.size() * 8);
        double* fp[1] = {updating_factor}; int32_t rows[1] = {1};
        insider_factors f{K, 1, fp, rows, V.data()};
        insider_options o; insider_b200_default_options(&o);
        o.tuning = tuning; o.lambda1 = lambda; o.max_iter = 0;
        insider_session* s = do_begin(ctx, r.get(), &f, &o);
        try {
            cudaStream_t st = ctx->stream;
            const Geom& g = s->g; const int KK = g.KP * g.KP;
            launch_row_b(g, s->masked, r->Y, r->trC, s->V, s->Bp, s->Gp, s->rb_splits, st);
            if (s->masked) launch_row_comp_gram(g, r->trR, s->V, s->Dp, s->d_splits, st);
            double* outs[3] = {s->B, s->G, s->D};
            const double* parts[3] = {s->Bp, s->Gp, s->Dp};
            const int64_t ns[3] = {(int64_t)g.N * g.KP, KK, (int64_t)g.N * KK};
            const int nps[3] = {s->rb_splits, s->rb_splits * ROW_B_GRAM_PARTS, s->d_splits};
            launch_reduce_jobs(s->masked ? 3 : 2, outs, parts, ns, nps, st);
            launch_continuous(g, s->masked, r->X, s->A_all, s->B, s->G, s->D, lambda, s->U, s->cont_scratch, s->err_dev, st);
            CUDA_TRY(cudaStreamSynchronize(st));
            download_factors(s, &f);
            int err = 0; CUDA_TRY(cudaMemcpy(&err, s->err_dev, 4, cudaMemcpyDeviceToHost));
            if (err) throw Err{INSIDER_ERR_NOT_SPD, "continuous-covariate normal equations not positive definite"};
        } catch (...) { destroy_session(s); throw; }
        destroy_session(s);
    });
}

int insider_b200_strong_cd(insider_ctx* ctx, int32_t K, int64_t n_cols, const double* XtX, int32_t shared_gram, const double* Xty, const double* wstart,
                           double lambda, double alpha, double tol, int32_t perm_mode, uint64_t seed, uint32_t als_iter, uint64_t gene0, double* beta,
                           int32_t* sweeps, char* errbuf, size_t errlen) {
    return guarded(errbuf, errlen, [&] {
        REQUIRE(ctx && XtX && Xty && wstart && beta && n_cols >= 0, "null argument");
        if (K < 1 || K > KMAX) throw Err{INSIDER_ERR_UNSUPPORTED, "K must be in 1..32"};
        if (perm_mode == 0) perm_mode = INSIDER_PERM_COUNTER;
        if (n_cols == 0) return;
        CUDA_TRY(cudaSetDevice(ctx->device));
        cudaStream_t st = ctx->stream;
        DevPool pool; pool.stream = st;
        const size_t gsz = (size_t)K * K * (shared_gram ? 1 : n_cols);
        double* dG = pool.get<double>(gsz, false); double* dx = pool.get<double>((size_t)K * n_cols, false);
        double* dw = pool.get<double>((size_t)K * n_cols, false); double* db = pool.get<double>((size_t)K * n_cols, true, st);
        int* dsw = pool.get<int>(n_cols, true, st);
        unsigned int* dq = pool.get<unsigned int>(1, true, st);
        CUDA_TRY(cudaMemcpyAsync(dG, XtX, gsz * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(dx, Xty, (size_t)K * n_cols * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(dw, wstart, (size_t)K * n_cols * 8, cudaMemcpyHostToDevice, st));
        if (shared_gram) {
            double* dtab = pool.get<double>(cd_dense_table_elems(), false);
            launch_cd_dense_batch(K, n_cols, dG, dx, dw, lambda, alpha, tol, perm_mode, seed, als_iter, db, dsw, ctx->perm_table, dtab, st);
        }
        else launch_cd_batch(K, n_cols, dG, false, dx, dw, lambda, alpha, tol, perm_mode, seed, als_iter, gene0, db, dsw, dq, ctx->perm_table, ctx->sm_count, st);
        CUDA_TRY(cudaMemcpyAsync(beta, db, (size_t)K * n_cols * 8, cudaMemcpyDeviceToHost, st));
        if (sweeps) CUDA_TRY(cudaMemcpyAsync(sweeps, dsw, (size_t)n_cols * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        CUDA_TRY(cudaGetLastError());
    });
}

int insider_b200_fit_interaction(insider_ctx* ctx, int64_t N, int64_t P, int32_t K, const double* residual, int32_t mask_kind, const void* train,
                                 double* interactions, int32_t n_levels, const int32_t* indicator, const double* column_factor, int32_t tuning,
                                 char* errbuf, size_t errlen) {
    // per-level normal equations without ridge on a caller-supplied residual: the row update of one confounder with
    // lambda = 0 and a zero row factor (then T_k = B_k and XtX_s = sum Gk_k)  — src/fit_interaction.cpp:37-54 / :59-82
    return guarded(errbuf, errlen, [&] {
        REQUIRE(ctx && residual && interactions && indicator && column_factor, "null argument");
        REQUIRE(tuning == 0 || tuning == 1, "Parameter tuning should be either 0 or 1!");
        REQUIRE(ctx->world == 1, "fit_interaction runs on a single-GPU context");
        insider_problem pb{};
        pb.N = N; pb.P = P; pb.C = 1; pb.Q = 0; pb.inc_continuous = 0; pb.Y = residual; pb.levels = indicator;
        pb.mask_kind = tuning == 1 ? mask_kind : INSIDER_MASK_NONE; pb.train = train; pb.test = train;
        std::unique_ptr<insider_resident, void (*)(insider_resident*)> r(do_upload(ctx, &pb), insider_b200_release);
        REQUIRE(r->L[0] == n_levels, "n_levels must equal the number of interaction levels");
        std::vector<double> A0((size_t)n_levels * K, 0.0), V((size_t)K * P);
        memcpy(V.data(), column_factor, V.size() * 8);
        double* fp[1] = {A0.data()}; int32_t rows[1] = {n_levels};
        insider_factors f{K, 1, fp, rows, V.data()};
        insider_options o; insider_b200_default_options(&o);
        o.tuning = tuning; o.lambda1 = 0.0; o.max_iter = 0;
        insider_session* s = do_begin(ctx, r.get(), &f, &o);
        try {
            cudaStream_t st = ctx->stream;
            const Geom& g = s->g; const int KK = g.KP * g.KP;
            launch_row_b(g, s->masked, r->Y, r->trC, s->V, s->Bp, s->Gp, s->rb_splits, st);
            launch_reduce_partials(s->G, s->Gp, KK, s->rb_splits * ROW_B_GRAM_PARTS, st);
            launch_reduce_partials(s->B, s->Bp, (int64_t)g.N * g.KP, s->rb_splits, st);
            if (s->masked) {
                launch_row_comp_gram(g, r->trR, s->V, s->Dp, s->d_splits, st);
                launch_reduce_partials(s->D, s->Dp, (int64_t)g.N * KK, s->d_splits, st);
                launch_level_gram(g, s->tab_dev, s->total_levels, s->max_chunks, s->G, s->D, s->GLp, st);
                launch_row_rhs(g, s->designs[0], s->B, s->G, s->D, s->U, s->T, st);
            }
            launch_level_factor(g, s->masked, s->tab_dev, s->total_levels, s->max_chunks, s->G, s->GLp, 0.0, s->Lfac, s->err_dev, st);
            launch_level_update(g, s->masked, s->designs[0], 0, s->G, s->B, s->T, s->Lfac, s->U, st);
            CUDA_TRY(cudaStreamSynchronize(st));
            download_factors(s, &f);
            int err = 0; CUDA_TRY(cudaMemcpy(&err, s->err_dev, 4, cudaMemcpyDeviceToHost));
            if (err) throw Err{INSIDER_ERR_NOT_SPD, "interaction normal equations not positive definite"};
        } catch (...) { destroy_session(s); throw; }
        destroy_session(s);
        memcpy(interactions, A0.data(), A0.size() * 8);
    });
}

namespace {
// regularised incomplete beta I_x(a, b) by the continued fraction (modified Lentz); lx = log(x), lxc = log(1 - x) are passed
// in because x = df / (df + t^2) is within rounding of 1 for the large degrees of freedom this is used with
double betacf(double a, double b, double x) {
    const double tiny = 1e-300, eps = 1e-16;
    const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0, d = 1.0 - qab * x / qap;
    if (fabs(d) < tiny) d = tiny;
    d = 1.0 / d;
    double h = d;
    for (int m = 1; m <= 2000000; ++m) {
        const double m2 = 2.0 * m;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d; if (fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c; if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d; h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d; if (fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c; if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < eps) break;
    }
    return h;
}
// two-sided p-value of Student's t with df degrees of freedom: I_{df/(df+t^2)}(df/2, 1/2)   (R: 2 * pt(-|t|, df))
double student_two_sided(double t, double df) {
    if (!(df > 0.0) || std::isnan(t)) return std::nan("");
    const double t2 = t * t;
    if (t2 == 0.0) return 1.0;
    if (std::isinf(t2)) return 0.0;
    const double a = 0.5 * df, b = 0.5;
    const double x = df / (df + t2), xc = t2 / (df + t2);
    const double lx = -log1p(t2 / df), lxc = log(xc);
    const double bt = exp(lgamma(a + b) - lgamma(a) - lgamma(b) + a * lx + b * lxc);
    if (x < (a + 1.0) / (a + b + 2.0)) return bt * betacf(a, b, x) / a;
    return 1.0 - bt * betacf(b, a, xc) / b;
}
}  // namespace

int insider_b200_glm_interaction(insider_ctx* ctx, int64_t N, int64_t P, int32_t K, const double* residual, int32_t n_levels,
                                 const int32_t* interaction_indicator, const double* column_factor, double* coeff, double* pval, char* errbuf,
                                 size_t errlen) {
    // Per level i the stacked regression of R/glm_interaction.R:15-24 has X'X = n_i V V' and X'y = V sum_{k in i} r_k: the dense
    // row update with lambda = 0 (the same launches as insider_b200_fit_interaction, tuning = 0). The coefficient table needs, per
    // level, RSS = sum_k |r_k|^2 - 2 a' X'y + n_i a' G a (k_row_sumsq gives the first term) and diag((n_i G)^-1).
    return guarded(errbuf, errlen, [&] {
        REQUIRE(ctx && residual && interaction_indicator && column_factor && coeff && pval, "null argument");
        REQUIRE(ctx->world == 1, "glm_interaction runs on a single-GPU context");
        if (K < 1 || K > KMAX) throw Err{INSIDER_ERR_UNSUPPORTED, "latent_dim must be in 1..32"};
        insider_problem pb{};
        pb.N = N; pb.P = P; pb.C = 1; pb.Q = 0; pb.inc_continuous = 0; pb.Y = residual; pb.levels = interaction_indicator; pb.mask_kind = INSIDER_MASK_NONE;
        std::unique_ptr<insider_resident, void (*)(insider_resident*)> r(do_upload(ctx, &pb), insider_b200_release);
        REQUIRE(r->L[0] == n_levels, "n_levels must equal the number of interaction levels");
        std::vector<double> A0((size_t)n_levels * K, 0.0), V((size_t)K * P);
        memcpy(V.data(), column_factor, V.size() * 8);
        double* fp[1] = {A0.data()}; int32_t rows[1] = {n_levels};
        insider_factors f{K, 1, fp, rows, V.data()};
        insider_options o; insider_b200_default_options(&o);
        o.tuning = 0; o.lambda1 = 0.0; o.max_iter = 0;
        insider_session* s = do_begin(ctx, r.get(), &f, &o);
        try {
            cudaStream_t st = ctx->stream;
            const Geom& g = s->g; const int KP = g.KP, KK = KP * KP, Nn = g.N;
            const int nb = std::max(1, std::min<int>(ctx->sm_count, (int)std::min<int64_t>(g.P, 1 << 20)));
            double* ss_part = s->pool.get<double>((size_t)nb * Nn, true, st);
            double* ss = s->pool.get<double>((size_t)Nn, true, st);
            launch_row_b(g, false, r->Y, nullptr, s->V, s->Bp, s->Gp, s->rb_splits, st);
            launch_reduce_partials(s->G, s->Gp, KK, s->rb_splits * ROW_B_GRAM_PARTS, st);
            launch_reduce_partials(s->B, s->Bp, (int64_t)Nn * KP, s->rb_splits, st);
            launch_row_sumsq(g, r->Y, ss_part, nb, st);
            launch_reduce_partials(ss, ss_part, Nn, nb, st);
            launch_level_factor(g, false, s->tab_dev, s->total_levels, s->max_chunks, s->G, s->GLp, 0.0, s->Lfac, s->err_dev, st);
            launch_level_update(g, false, s->designs[0], 0, s->G, s->B, s->T, s->Lfac, s->U, st);
            std::vector<double> Bh((size_t)Nn * KP), Gh(KK), ssh(Nn);
            CUDA_TRY(cudaMemcpyAsync(Bh.data(), s->B, Bh.size() * 8, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(Gh.data(), s->G, Gh.size() * 8, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(ssh.data(), ss, ssh.size() * 8, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            download_factors(s, &f);
            int err = 0; CUDA_TRY(cudaMemcpy(&err, s->err_dev, 4, cudaMemcpyDeviceToHost));
            if (err) throw Err{INSIDER_ERR_NOT_SPD, "V V' is not positive definite: the regression of glm_interaction is rank deficient"};
            // K x K host algebra of summary.glm: diag(G^-1) by Cholesky
            std::vector<double> Lc((size_t)K * K, 0.0), ginv_diag(K, 0.0);
            for (int j = 0; j < K; ++j) {
                double d = Gh[(size_t)j * KP + j];
                for (int m = 0; m < j; ++m) d -= Lc[(size_t)j * K + m] * Lc[(size_t)j * K + m];
                if (!(d > 0.0)) throw Err{INSIDER_ERR_NOT_SPD, "V V' is not positive definite"};
                Lc[(size_t)j * K + j] = sqrt(d);
                for (int i = j + 1; i < K; ++i) {
                    double v = Gh[(size_t)i * KP + j];
                    for (int m = 0; m < j; ++m) v -= Lc[(size_t)i * K + m] * Lc[(size_t)j * K + m];
                    Lc[(size_t)i * K + j] = v / Lc[(size_t)j * K + j];
                }
            }
            for (int c = 0; c < K; ++c) {                       // column c of L^-1, then (G^-1)_cc = sum_i (L^-1)_ic^2 ... via solve
                std::vector<double> y(K, 0.0);
                for (int i = c; i < K; ++i) {
                    double v = (i == c) ? 1.0 : 0.0;
                    for (int m = c; m < i; ++m) v -= Lc[(size_t)i * K + m] * y[m];
                    y[i] = v / Lc[(size_t)i * K + i];
                }
                double acc = 0.0;
                for (int i = c; i < K; ++i) acc += y[i] * y[i];
                ginv_diag[c] = acc;
            }
            const std::vector<int>& lor = r->lor_host[0];
            std::vector<double> SB((size_t)n_levels * K, 0.0), SS(n_levels, 0.0), cnt(n_levels, 0.0);
            for (int k = 0; k < Nn; ++k) {
                const int l = lor[k];
                for (int a = 0; a < K; ++a) SB[(size_t)l * K + a] += Bh[(size_t)k * KP + a];
                SS[l] += ssh[k]; cnt[l] += 1.0;
            }
            for (int l = 0; l < n_levels; ++l) {
                double aSB = 0.0, aGa = 0.0;
                for (int a = 0; a < K; ++a) {
                    const double ca = A0[l + (size_t)a * n_levels];
                    aSB += ca * SB[(size_t)l * K + a];
                    double ga = 0.0;
                    for (int b = 0; b < K; ++b) ga += Gh[(size_t)a * KP + b] * A0[l + (size_t)b * n_levels];
                    aGa += ca * ga;
                }
                const double rss = SS[l] - 2.0 * aSB + cnt[l] * aGa;
                const double df = cnt[l] * (double)P - (double)K;
                const double sigma2 = rss / df;                                    // dispersion of the gaussian family (summary.glm)
                for (int a = 0; a < K; ++a) {
                    const double ca = A0[l + (size_t)a * n_levels];
                    const double se = sqrt(sigma2 * ginv_diag[a] / cnt[l]);
                    coeff[l + (size_t)a * n_levels] = ca;
                    pval[l + (size_t)a * n_levels] = student_two_sided(ca / se, df);
                }
            }
        } catch (...) { destroy_session(s); throw; }
        destroy_session(s);
    });
}

}  // extern "C"
