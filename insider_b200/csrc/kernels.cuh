// Kernel launch interfaces of libinsider_b200 (definitions in k_*.cu). All pointers are device pointers.
//
// Device data layout (see DESIGN.md "Data layout in HBM"):
//   Y    [P_pad][ldY]   column j = gene j, ldY = pitch4(N) rows, zero padded (rows >= N, genes >= P)
//   V    [P_pad][ldV]   gene j's K-vector contiguous, ldV = pitch4(KP), KP = round_up(K, 8), zero padded
//   Xty  [P_pad][ldV]   same layout as V
//   trC/teC [P_pad][Wp] train/test mask bits, bit (i & 31) of word i>>5 of gene j; Wp = round_up(ceil(N/32), 4)
//   trR  [N][WPr]       train mask bits transposed: bit (j & 31) of word j>>5 of row i; WPr = ceil(P_pad/32)
//   U    [N][KP]        row factor, row-major, zero padded columns
//   Ut   [KP][ldT]      its transpose, ldT = pitch4(round_up(N, 8)), zero padded
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ib {

struct Geom {
    int N, K, KP, NT;        // NT = KP / 8
    int ldY, ldV, ldT, Wp, WPr;
    int64_t P, P_pad;        // local genes
    int n_tiles;             // P_pad / TG
    int64_t gene0;           // global index of local gene 0 (permutation keys use global gene ids)
};

struct LevelTable;
// ---- streaming passes over Y (k_stream.cu) ------------------------------------------------------------------
// Bp[split][N][KP] = sum over the split's genes of (M o Y) V^T ; Gp[split*4 + w][KP*KP] = partial V V^T (w = 0..3), fused
void launch_row_b(const Geom& g, bool masked, const double* Y, const uint32_t* trC, const double* V, double* Bp, double* Gp, int n_splits,
                  cudaStream_t st);
constexpr int ROW_B_GRAM_PARTS = 4;
// the same with optional outputs: Gp == nullptr skips the V V^T tiles; lv_ptr != nullptr (dense, single slab only:
// row_b_levels_supported) stores per-level sums of the block's B instead of B itself: Bp[split][n_levels][KP]. lv_ptr / lv_rows:
// CSR over all levels of all confounders (rows of a level ascending), n_lv_rows = C * N entries
void launch_row_b_ex(const Geom& g, bool masked, const double* Y, const uint32_t* trC, const double* V, double* Bp, double* Gp, int n_splits,
                     const int* lv_ptr, const int* lv_rows, int n_levels, int n_lv_rows, cudaStream_t st);
bool row_b_levels_supported(const Geom& g, int n_levels, int n_lv_rows);
size_t row_b_partial_elems(const Geom& g, int n_splits);
int row_b_default_splits(const Geom& g, int sm_count);
// Xty[j][k] = sum_i U[i][k] m_ij y_ij
void launch_col_xty(const Geom& g, bool masked, const double* Y, const uint32_t* trC, const double* Ut, double* Xty, int n_blocks,
                    cudaStream_t st);
// partial[block][4] = { sse_train, sse_test, sum v^2, sum |v| } over the block's genes; r = y - u.v
void launch_sse(const Geom& g, bool masked, const double* Y, const uint32_t* trC, const uint32_t* teC, const double* Ut, const double* V,
                double* partial, int n_blocks, cudaStream_t st);
int stream_default_blocks(const Geom& g, int sm_count);

// ---- Gram matrices (k_gram.cu) ------------------------------------------------------------------------------
// Dp[split][N][KP*KP]: per-row complement Gram sum_{j: m_ij = 0} v_j v_j^T over the split's genes
void launch_row_comp_gram(const Geom& g, const uint32_t* trR, const double* V, double* Dp, int n_splits, cudaStream_t st);
// fixed-order reductions of the partial buffers:
//   B[N][K] = sum_s Bp[s] ; G[KP*KP] = sum_b Gp[b] ; D[N][KP*KP] = sum_s Dp[s]
void launch_reduce_partials(double* out, const double* partials, int64_t n_elems, int n_parts, cudaStream_t st);
// the same for up to three buffers in one launch
void launch_reduce_jobs(int n_jobs, double* const* out, const double* const* parts, const int64_t* n_elems, const int* n_parts, cudaStream_t st);

// G[KP*KP] = V V^T over the local genes (src/optimize.cpp:332) as its own small kernel: block partials in `parts`
// (gram_v_parts(P) x KP*KP doubles) combined in block order by the last block to finish (`counter`: one zero-initialised
// unsigned int, left at zero). bump != nullptr: that block also advances the device-side ALS iteration counter.
int gram_v_parts(int64_t P);
struct CheckState;
void launch_gram_v(const Geom& g, const double* V, double* parts, double* G, unsigned int* counter, CheckState* bump, cudaStream_t st);

// ---- row side (k_rows.cu) -----------------------------------------------------------------------------------
struct RowDesign {            // one categorical confounder
    int L;                    // levels
    const int* level_of_row;  // [N] 0-based
    const int* rows_sorted;   // [N] rows grouped by level
    const int* level_start;   // [L+1]
    double* A;                // [L][KP] row-major factor
};
struct LevelTable {           // one level of one confounder (all confounders concatenated)
    const int* rows_sorted;   // the confounder's rows grouped by level
    int row_begin, row_end;   // this level's slice of rows_sorted
};
// masked path: GLp[level][chunk][KP*KP] = sum over 32-row chunks of (G - D_k)   (all levels of all confounders, one launch)
void launch_level_gram(const Geom& g, const LevelTable* tab_dev, int total_levels, int max_chunks, const double* G, const double* D, double* GLp,
                       cudaStream_t st);
// batched K x K Cholesky, one warp per level: XtX_s = sum_chunks GLp (masked) or n_s G (dense), + lambda I -> Lfac[level][KP*KP + KP]
void launch_level_factor(const Geom& g, bool masked, const LevelTable* tab_dev, int total_levels, int max_chunks, const double* G,
                         const double* GLp, double lambda, double* Lfac, int* err_flag, cudaStream_t st);
// masked path: T[k][:] = B_k - (G - D_k)(u_k - a_{c,z(k)})
void launch_row_rhs(const Geom& g, const RowDesign& d, const double* B, const double* G, const double* D, const double* U, double* T,
                    cudaStream_t st);
// per level of confounder d: right-hand side, substitution with the stored factor, update A and the rows of U
void launch_level_update(const Geom& g, bool masked, const RowDesign& d, int lfac_base, const double* G, const double* B, const double* T,
                         const double* Lfac, double* U, cudaStream_t st);
// dense path: SB[level][KP] = sum of B over the level's rows (all confounders, one launch), then the Gauss-Seidel sweep over
// all confounder blocks in one single-block launch using the design's co-occurrence counts (no N-length loops)
struct DenseGs { const int* lvl_first; const int* co_ptr; const int* co_row; const double* co_cnt; const double* Sx; };
void launch_level_sumB(const Geom& g, const LevelTable* tab_dev, int total_levels, const double* B, double* SB, cudaStream_t st);
void launch_rows_dense_gs(const Geom& g, const DenseGs& d, int C, int Q, int total_levels, int max_levels, int nnz, double* A_all, const double* W, const double* SB,
                          const double* G, const double* Lfac, cudaStream_t st);
// the same with the row-factor rebuild (U, Ut, UtU: k_build_u + k_gram_u_final) fused into the tail of the cluster kernel when
// rows_dense_gs_can_fuse_u(); otherwise the caller launches launch_build_u afterwards
void launch_rows_dense_gs_ex(const Geom& g, const DenseGs& d, int C, int Q, int total_levels, int max_levels, int nnz, double* A_all, const double* W,
                             const double* SB, const double* G, const double* Lfac, const int* row_lv /*[N][C] global level of (row, confounder)*/,
                             double* U, double* Ut, double* UtU, cudaStream_t st);
bool rows_dense_gs_can_fuse_u(const Geom& g, int Q, int total_levels, int max_levels, int nnz);
// continuous covariate q: H = sum_k x_k^2 Gk_k, Tq = sum_k x_k (B_k - Gk_k u_k); cyclic coordinate update / solve; updates w and U
void launch_continuous(const Geom& g, bool masked, const double* x, double* w /*[KP]*/, const double* B, const double* G, const double* D,
                       double lambda, double* U, double* scratch, int* err_flag, cudaStream_t st);
size_t continuous_scratch_elems(const Geom& g);
// U = sum_c A_c[z_c] + X W ; also writes Ut and UtU = U^T U (UtU buffer: KP*KP result + build_u_parts(g) partial blocks)
void launch_build_u(const Geom& g, int C, const RowDesign* designs_dev, int Q, const double* X, const double* W, double* U, double* Ut,
                    double* UtU, cudaStream_t st);
int build_u_parts(const Geom& g);

// ---- column side (k_cd.cu) ----------------------------------------------------------------------------------
struct CdParams {
    double lambda, alpha;
    const double* tol;        // device scalar: sub_tol * decay
    const uint32_t* als_iter; // device scalar
    uint64_t seed;
    int perm_mode;
};
// masked path: XtXall[j][KP*KP] = UtU - sum_{i: m_ij=0} u_i u_i^T for every local gene (DMMA gathers, one warp per gene)
void launch_col_gram(const Geom& g, const uint32_t* trC, const double* U, const double* UtU, double* XtXall, cudaStream_t st);
// per gene: alpha == 0 -> ridge solve, else elastic-net CD (persistent groups pulling genes from `queue`). Updates V in place.
// sweeps_per_gene / order (optional): record every gene's sweep count; hand the genes out in the given order (launch_cd_order
// on the previous iteration's counts: the longest solves start first, which shortens the tail of the launch).
void launch_col_solve(const Geom& g, bool masked, const double* UtU, const double* XtXall, const double* Xty, double* V, const CdParams& p,
                      unsigned long long* sweeps, unsigned long long* steps, unsigned int* queue, const unsigned char* perm_table, int* err_flag,
                      int sm_count, int* sweeps_per_gene, const int* order, bool long_solves, cudaStream_t st);
// dense path, alpha != 0: thread-per-gene elastic-net CD with the shared Gram UtU (k_cd_dense.cu). `order` (optional) maps
// thread slots to genes; `sweeps_per_gene` (optional) receives every gene's sweep count.
// `table`: cd_dense_table_elems() doubles filled by launch_cd_dense_table() from UtU (independent of Xty: may run beside
// k_col_xty); resident_all: the 200-register variant (10 one-warp blocks per SM) instead of the 255-register one (8 per SM,
// faster sweeps).
// Phased execution: a launch runs sweeps [draw0, cap) of the genes it is given; genes that have not converged at sweep `cap`
// are parked in `ps` (state in coordinate order) and taken up by the next launch (draw0 = that cap) in a new slot order
// (launch_cd_order_parked). ps == nullptr, draw0 = 0, cap = 0xffffffff: one launch runs every gene to convergence.
struct CdPhaseState {
    double* state;      // [P][2 * round_up(K, 4)]
    uint32_t* inc;      // [P]
    float* dl;          // [P]
    int* alive;         // [P]
    int* n_slots;       // device scalar: parked genes = slots of the next phase
};
size_t cd_dense_table_elems();
void launch_cd_dense_table(int K, const double* XtX, int xs_r, int xs_c, double lambda, double alpha, double* table, cudaStream_t st);
void launch_cd_dense(const Geom& g, const double* UtU, const double* Xty, double* V, const CdParams& p, unsigned long long* sweeps,
                     unsigned long long* steps, int* sweeps_per_gene, const int* order, const unsigned char* perm_table, const double* table,
                     bool resident_all, const CdPhaseState* ps, uint32_t draw0, uint32_t cap, const double* tables_all, cudaStream_t st);
// tables_all (optional): the prepared table permuted into each of the PERM_T visiting orders (cd_dense_tables_all_elems(K)
// doubles, 22 MB at K = 23); with it the solver fetches every sweep's table by one TMA bulk copy instead of building it
size_t cd_dense_tables_all_elems(int K);
void launch_cd_dense_tables_all(int K, const double* table, const unsigned char* perm_table, double* tables_all, cudaStream_t st);
// order[] for the next launch_cd_dense from the sweep counts of the last one (descending, bucketed)
// (`work`: cd_order_work_ints() ints, zero-initialised once; the kernel leaves it zeroed)
size_t cd_order_work_ints();
void launch_cd_order(const int* sweeps_per_gene, int64_t P, int* order, int* work, const uint32_t* als_iter, cudaStream_t st);
// order[] of the parked genes for the next phase (by the loss decrement of their last sweep) and their count in ps.n_slots
void launch_cd_order_parked(const CdPhaseState& ps, int64_t P, int* order, int* work, const double* tol_dev, cudaStream_t st);
void launch_cd_dense_batch(int K, int64_t n, const double* XtX, const double* Xty, const double* w0, double lambda, double alpha, double tol,
                           int perm_mode, uint64_t seed, uint32_t als_iter, double* beta, int* sweeps, const unsigned char* perm_table,
                           double* table, cudaStream_t st);
// stand-alone batched solver (insider_b200_strong_cd): XtX column-major K x K, either shared or per column [n][K*K]
void launch_cd_batch(int K, int64_t n, const double* XtX, bool shared, const double* Xty, const double* w0, double lambda, double alpha,
                     double tol, int perm_mode, uint64_t seed, uint32_t als_iter, uint64_t gene0, double* beta, int* sweeps,
                     unsigned int* queue, const unsigned char* perm_table, int sm_count, cudaStream_t st);

// ---- masked column update, round 2 (k_cd_masked.cu) ----------------------------------------------------------
// tiles[n_tiles][E][32]: lower triangles of the per-gene matrices XtX_j = UtU - sum_{i: m_ij=0} u_i u_i^T plus the
// reciprocals 1/(XtX_kk + l2), 32 genes per tile in SLOT order (slot i = gene order[i]) in the solver's shared-memory layout
size_t cd_masked_tile_doubles(int K, int64_t P);
void launch_col_gram_tiles(const Geom& g, const uint32_t* trC, const double* U, const double* UtU, const int* order, double lambda, double alpha,
                           double* tiles, cudaStream_t st);
// thread-per-gene elastic-net CD on those tiles (one warp per tile, tile fetched by one TMA bulk copy). Updates V in place.
void launch_cd_masked(const Geom& g, const double* tiles, const double* Xty, double* V, const CdParams& p, unsigned long long* sweeps,
                      unsigned long long* steps, int* sweeps_per_gene, const int* order, const unsigned char* perm_table, cudaStream_t st);

// ---- misc (k_misc.cu) ---------------------------------------------------------------------------------------
// src: n_genes columns of N mask elements (INSIDER_MASK_* kind) -> dstC[n_genes][Wp] bit-packed
void launch_pack_mask(const void* src, int kind, int64_t N, int64_t n_genes, int Wp, uint32_t* dstC, cudaStream_t st);
void launch_transpose_mask(const uint32_t* trC, int64_t N, int64_t P_pad, int Wp, int WPr, uint32_t* trR, cudaStream_t st);
void launch_count_bits(const uint32_t* m, int64_t n_words, unsigned long long* out, cudaStream_t st);
// dst[r*dst_ld + c] = src[r*src_ld + c], c < width (device re-pitch behind the contiguous host <-> device copies)
void launch_repitch(double* dst, int64_t dst_ld, const double* src, int64_t src_ld, int width, int64_t rows, cudaStream_t st);
struct CheckState {           // device-resident loop state
    double loss, pre_loss, decay, tol, sub_tol, global_tol;
    double sse_train, sse_test, v2, v1, row_reg;
    double train_rmse, test_rmse, delta_loss;
    double n_train, n_test, np_total;
    double lambda1, lambda2, alpha;
    uint32_t als_iter;
    int converged, diverged, tuning;
};
// reduce the SSE partials (fixed order) into state->{sse_train, sse_test, v2, v1}
void launch_sse_reduce(const double* partial, int n_blocks, CheckState* state, cudaStream_t st);
// row_reg = lambda1 * sum ||A_c||^2 (+ W); then loss, delta, decay ladder, convergence (src/optimize.cpp:381-408)
void launch_check(CheckState* state, const double* A_all, int64_t n_A, int initial, int iter, void* record_out, cudaStream_t st);
void launch_bump_iter(CheckState* state, cudaStream_t st);
// part[n_blocks][N]: per-row sums of squares of Y over each block's genes (glm_interaction's residual sums of squares)
void launch_row_sumsq(const Geom& g, const double* Y, double* part, int n_blocks, cudaStream_t st);

}  // namespace ib
