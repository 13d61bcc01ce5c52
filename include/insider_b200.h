/*
 * insider_b200.h — C ABI of libinsider_b200.so: the B200-native (sm_100a) replacement for the
 * alternating-optimisation fit of kai0511/insider.
 *
 * What each entry point replaces in the reference (paths relative to the reference checkout):
 *
 *   insider_b200_optimize            <-  .Call(`_insider_optimize`, 16 SEXPs)   R/RcppExports.R:20-22,
 *                                        src/RcppExports.cpp:87-110  ->  optimize()  src/optimize.cpp:256-422
 *   insider_b200_upload / _optimize_resident / _release
 *                                    <-  the same call, split so that tune()'s 51 fits on one data set
 *                                        (R/insider.R:100-174) upload the N x P matrix and masks once
 *   insider_b200_als_begin/_step/_read/_end
 *                                    <-  the `while(iter <= max_iter)` loop body, src/optimize.cpp:325-410,
 *                                        exposed iteration-by-iteration for benchmarks and parity tests
 *   insider_b200_strong_cd           <-  .Call(`_insider_strong_coordinate_descent`, 8 SEXPs)
 *                                        src/RcppExports.cpp:34-50 -> src/coordinate_descent.cpp:57-127
 *   insider_b200_optimize_continuous <-  .Call(`_insider_optimize_continuous_v2`, 8 SEXPs)
 *                                        src/RcppExports.cpp:69-84 -> optimize_continuous_v2()  src/optimize.cpp:77-137
 *   insider_b200_fit_interaction     <-  fit_interaction()  src/fit_interaction.cpp:10-90 (not exported by the reference)
 *   insider_b200_glm_interaction     <-  glm_interaction()  R/glm_interaction.R:2-30 (per-level OLS + Gaussian-GLM p-values)
 *   insider_b200_split               <-  ratio_splitter()  R/utils.R:78-117 (bit-exact train/test masks)
 *   insider_b200_tune_batch          <-  the two grid loops of tune()  R/insider.R:100-132, 145-174
 *
 * Conventions
 *   - All matrices are COLUMN-MAJOR double precision exactly as R / Armadillo hold them.
 *   - `levels` (the reference's cfd_indicators, src/optimize.cpp:256) is N x C, 1-based, and every column's
 *     values must be exactly 1..L_c (the reference indexes row `level-1`, src/optimize.cpp:175,190).
 *   - Confounder factors and column_factor are IN/OUT host buffers, updated in place like the reference
 *     (src/optimize.cpp:283-284 aliases R memory). The library never keeps host pointers after a call returns.
 *   - Functions return INSIDER_OK or an error code and write a message to errbuf; nothing exits or throws
 *     across the ABI (the reference calls exit(1) on bad flags: src/optimize.cpp:270-273).
 *   - There is no CPU fallback: every entry that computes fails with INSIDER_ERR_CUDA when no sm_100 device
 *     is usable.
 */
#ifndef INSIDER_B200_H
#define INSIDER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define INSIDER_B200_VERSION 200

enum {
    INSIDER_OK = 0,
    INSIDER_ERR_INVALID_ARG = 1,
    INSIDER_ERR_CUDA = 2,
    INSIDER_ERR_NCCL = 3,
    INSIDER_ERR_NOT_SPD = 4,        /* a normal-equation matrix was not positive definite (reference: SVD fallback + warning) */
    INSIDER_ERR_DIVERGED = 5,       /* NaN/Inf loss */
    INSIDER_ERR_EMPTY_TEST_SET = 6, /* tuning=1 with no test entries (reference: arma::mean of empty throws, src/utils.cpp:67) */
    INSIDER_ERR_NOMEM = 7,
    INSIDER_ERR_UNSUPPORTED = 8     /* latent_dim > 32 (checked before anything is uploaded) */
};

/* mask element types accepted for train/test indicators */
enum {
    INSIDER_MASK_NONE = 0,   /* no masks (only valid with tuning = 0) */
    INSIDER_MASK_INT32 = 1,  /* R integer matrix (R/insider.R:57-58) */
    INSIDER_MASK_UINT8 = 2,
    INSIDER_MASK_DOUBLE = 3  /* what the reference's C++ sees after Rcpp converts (src/optimize.cpp:256) */
};

/* coordinate-visit order of the elastic-net solver (src/coordinate_descent.cpp:89 draws it from R's global RNG,
 * which cannot be reproduced by a parallel solver; see DESIGN.md "permutations") */
enum {
    INSIDER_PERM_COUNTER = 1,  /* counter-based: key(seed, als_iter, sweep index of the solve) selects one of 4096 precomputed
                                * randperm-style permutations of the K coordinates; a gene visits its ACTIVE coordinates in that
                                * order. The gene is deliberately NOT part of the key: every gene at the same sweep index of the
                                * same ALS iteration shares the order (fresh and uniformly random per sweep, like
                                * randperm(|inc|), but not independent across genes) - DESIGN.md section 2 "Permutations" */
    INSIDER_PERM_IDENTITY = 2  /* ascending coordinate order */
};

typedef struct insider_ctx insider_ctx;           /* one per process per GPU */
typedef struct insider_resident insider_resident; /* a problem (data, masks, design) resident in HBM */
typedef struct insider_session insider_session;   /* an ALS run in progress on a resident problem */

typedef struct {
    int64_t N, P;            /* data is N x P (samples x genes) */
    int32_t C;               /* categorical confounders (columns of `levels`) */
    int32_t Q;               /* continuous covariates (columns of X); ignored unless inc_continuous == 1 */
    int32_t inc_continuous;  /* 0 or 1 (src/optimize.cpp:270-278) */
    int32_t mask_kind;       /* INSIDER_MASK_* */
    const double* Y;         /* N x P */
    const int32_t* levels;   /* N x C, 1-based level ids */
    const double* X;         /* N x Q or NULL */
    const void* train;       /* N x P of mask_kind, nonzero = member; NULL with INSIDER_MASK_NONE */
    const void* test;        /* N x P */
} insider_problem;

typedef struct {
    int32_t K;                  /* latent_dim, 1..32 */
    int32_t n_factors;          /* C + inc_continuous */
    double* const* factors;     /* factor c: L_c x K (continuous block: Q x K), column-major, in/out */
    const int32_t* factor_rows; /* rows of each factor matrix */
    double* column_factor;      /* K x P, in/out */
} insider_factors;

typedef struct {
    double lambda1, lambda2, alpha;  /* src/optimize.cpp:257 */
    int32_t tuning;                  /* 1 = masked train/test path, 0 = dense path */
    int32_t perm_mode;               /* INSIDER_PERM_* ; 0 -> INSIDER_PERM_COUNTER */
    double global_tol, sub_tol;
    uint32_t max_iter;               /* loop runs while iter <= max_iter (src/optimize.cpp:325) */
    uint32_t check_every;            /* 0 -> 10 (src/optimize.cpp:381) */
    uint64_t seed;                   /* permutation seed (the Rcpp shim draws it from R's RNG inside RNGScope) */
    int32_t verbose;                 /* 1: print the reference's per-check lines to stdout */
    int32_t use_graph;               /* 0 (default): steady-state iterations replay a CUDA graph, the first (solver-dominated) ones are
                                      * launched directly; 1: every iteration is a graph replay; -1: plain kernel launches only */
} insider_options;

/* one record per evaluation: the initial one (iter = -1, src/optimize.cpp:320-323) and every check_every-th iteration */
typedef struct {
    int32_t iter;
    int32_t pad;
    double sum_residual, train_rmse, test_rmse, row_reg, col_reg, l1_reg, loss, delta_loss, decay;
} insider_check;

typedef struct {
    double train_rmse, test_rmse, loss; /* of the last evaluated iteration (src/optimize.cpp:413-421); test_rmse is NaN when tuning = 0 */
    uint32_t iters_run;                 /* value of `iter` when the loop ended */
    uint32_t n_checks;                  /* records written (<= max_checks) */
    insider_check* checks;              /* optional caller buffer */
    uint32_t max_checks;
    int64_t cd_sweeps;                  /* total coordinate-descent sweeps over all genes and iterations */
    double loop_ms;                     /* device time of the ALS loop (CUDA events on the library's stream) */
    double h2d_bytes, d2h_bytes;        /* bytes copied by this call */
    int64_t kernel_launches;            /* kernels launched by this call */
    int64_t cd_steps;                   /* total coordinate updates attempted by the elastic-net solver (statistics) */
} insider_result;

void insider_b200_default_options(insider_options* opt);
int insider_b200_version(void);

/* context */
int insider_b200_ctx_create(insider_ctx** out, int device, char* errbuf, size_t errlen);
/* gene-sharded multi-GPU context: rank r owns a contiguous block of genes; nccl_id is the 128-byte ncclUniqueId
 * produced by insider_b200_nccl_unique_id on rank 0 and broadcast by the caller (torch.distributed / MPI / R). */
int insider_b200_nccl_unique_id(void* out128, char* errbuf, size_t errlen);
int insider_b200_ctx_create_dist(insider_ctx** out, int device, int rank, int world, const void* nccl_id128, char* errbuf, size_t errlen);
void insider_b200_ctx_destroy(insider_ctx* ctx);
/* the CUDA stream all work of this context is issued on (a cudaStream_t) */
void* insider_b200_ctx_stream(insider_ctx* ctx);

/* one-shot drop-in for `_insider_optimize` */
int insider_b200_optimize(insider_ctx* ctx, const insider_problem* prob, const insider_factors* fac, const insider_options* opt,
                          insider_result* res, char* errbuf, size_t errlen);

/* resident problems */
int insider_b200_upload(insider_ctx* ctx, const insider_problem* prob, insider_resident** out, char* errbuf, size_t errlen);
void insider_b200_release(insider_resident* r);
int insider_b200_optimize_resident(insider_ctx* ctx, insider_resident* r, const insider_factors* fac, const insider_options* opt,
                                   insider_result* res, char* errbuf, size_t errlen);

/* stepping interface (the loop body of src/optimize.cpp:325-410) */
int insider_b200_als_begin(insider_ctx* ctx, insider_resident* r, const insider_factors* fac, const insider_options* opt,
                           insider_session** out, char* errbuf, size_t errlen);
/* run up to n_iters more iterations; *done is set when the loop has ended (convergence at a check, or iter > max_iter);
 * *ms (optional) receives the device time of these iterations */
int insider_b200_als_step(insider_session* s, uint32_t n_iters, int32_t* done, double* ms, char* errbuf, size_t errlen);
/* copy the current factors to the host buffers of `fac` */
int insider_b200_als_read(insider_session* s, const insider_factors* fac, char* errbuf, size_t errlen);
/* finish: copies factors back (if fac != NULL), fills res, frees the session */
int insider_b200_als_end(insider_session* s, const insider_factors* fac, insider_result* res, char* errbuf, size_t errlen);
/* per-kernel device times of the session so far: names is a '\n'-joined list, ms[i] the accumulated time, calls[i] the
 * launch count; returns the number of distinct kernels (timing is only collected when profile != 0 at als_begin time) */
int insider_b200_als_profile(insider_session* s, char* names, size_t names_len, double* ms, int64_t* calls, int max_entries);
void insider_b200_set_profile(insider_ctx* ctx, int on);
/* diagnostics: coordinate-descent sweeps every local gene needed in the last iteration (the count strong_coordinate_descent's
 * do-while ran, src/coordinate_descent.cpp:86-114); n = number of entries of `out`, at most the context's local gene count.
 * Returns the number of entries written (0 when the column update is a ridge solve, alpha = 0). */
int64_t insider_b200_als_sweeps(insider_session* s, int32_t* out, int64_t n);
/* optional hint for the elastic-net solvers: expected sweep count of every local gene in the NEXT iteration (e.g. the
 * counts of a previous fit of the same data). It only orders the work (genes with similar counts share a warp); results do
 * not depend on it. Returns the number of entries taken. */
int64_t insider_b200_als_hint_sweeps(insider_session* s, const int32_t* hint, int64_t n);

/* batched single-column elastic-net solves (src/coordinate_descent.cpp:57-127). Problem b uses XtX[b] (K x K), Xty[b] (K),
 * wstart[b] (K); X and y of the reference signature are not needed in covariance form and are accepted as NULL.
 * shared_gram != 0: one K x K matrix for all columns. beta: K x n_cols out; sweeps: n_cols out (optional).
 * gene0 is accepted for source compatibility and IGNORED: the visiting order does not depend on the gene (see
 * INSIDER_PERM_COUNTER), so results do not depend on which columns of a larger problem a batch holds. */
int insider_b200_strong_cd(insider_ctx* ctx, int32_t K, int64_t n_cols, const double* XtX, int32_t shared_gram, const double* Xty,
                           const double* wstart, double lambda, double alpha, double tol, int32_t perm_mode, uint64_t seed,
                           uint32_t als_iter, uint64_t gene0, double* beta, int32_t* sweeps, char* errbuf, size_t errlen);

/* per-level un-regularised normal equations on a residual (src/fit_interaction.cpp:10-90):
 * interactions is n_levels x K column-major (out), indicator N (1-based), residual N x P, train as in insider_problem */
int insider_b200_fit_interaction(insider_ctx* ctx, int64_t N, int64_t P, int32_t K, const double* residual, int32_t mask_kind,
                                 const void* train, double* interactions, int32_t n_levels, const int32_t* indicator,
                                 const double* column_factor, int32_t tuning, char* errbuf, size_t errlen);

/* ratio_splitter (R/utils.R:78-117): R-exact Mersenne-Twister `sample()`; writes 0/1 int32 N x P masks.
 * Host-side (the R RNG stream is sequential); NaN entries of data are NA. Returns the number of test entries in *n_test. */
int insider_b200_split(const double* data, int64_t N, int64_t P, double ratio, uint32_t seed, int32_t* train, int32_t* test,
                       int32_t* na, int64_t* n_test, char* errbuf, size_t errlen);

/* one continuous covariate's factor (optimize_continuous_v2, src/optimize.cpp:77-137; same 8 arguments as the registered
 * `_insider_optimize_continuous_v2`, src/RcppExports.cpp:69-84, with `gram` = V V^T recomputed on the device):
 *   data N x P            the residual WITH this covariate's contribution added back (src/optimize.cpp:344)
 *   indicator N x P       train mask (mask_kind; ignored when tuning = 0)
 *   updating_factor K     in/out: the covariate's factor w (one row of cfd_matrices[[C+1]])
 *   c_factor K x P        column_factor V
 *   updating_confd N      the covariate x
 * tuning = 1: cyclic coordinate updates until sum|w_prev - w| < 0.1 (:102-126); tuning = 0: the closed form (:127-131). */
int insider_b200_optimize_continuous(insider_ctx* ctx, int64_t N, int64_t P, int32_t K, const double* data, int32_t mask_kind,
                                     const void* indicator, double* updating_factor, const double* c_factor,
                                     const double* updating_confd, double lambda, int32_t tuning, char* errbuf, size_t errlen);

/* glm_interaction (R/glm_interaction.R:2-30): for every interaction level i, ordinary least squares of the level's residual
 * rows (all P genes, no mask, no ridge) on t(column_factor) - i.e. glm(response ~ . - 1, gaussian) on the stacked data - and
 * its coefficient table. coeff, pval: n_levels x K column-major (row i-1 = level i); pval = two-sided t-test with
 * n_i P - K degrees of freedom and the residual-deviance dispersion, like coef(summary(fit))[,4]. */
int insider_b200_glm_interaction(insider_ctx* ctx, int64_t N, int64_t P, int32_t K, const double* residual, int32_t n_levels,
                                 const int32_t* interaction_indicator, const double* column_factor, double* coeff, double* pval,
                                 char* errbuf, size_t errlen);

/* grid of fits (tune(), R/insider.R:100-174): point g uses fac[g] (K, caller-initialised factors, in/out) and opt[g]
 * (lambda1, lambda2, alpha, tuning_iter, ...); results in res[g]. The points are independent fits and run as REPLICAS with no
 * communication: n_ctx contexts (e.g. one per GPU, or several on one GPU: their kernels overlap), context c holding its own
 * resident copy residents[c] of the same problem. One host thread per context pulls points from a shared queue (a K = 30 fit
 * costs several times a K = 10 fit: static partitions would idle). point_ctx (optional, n_points) receives the index of the
 * context that ran each point. Results do not depend on the schedule. Every context must be a plain single-GPU context. */
int insider_b200_tune_batch(int32_t n_ctx, insider_ctx* const* ctxs, insider_resident* const* residents, int32_t n_points,
                            const insider_factors* fac, const insider_options* opt, insider_result* res, int32_t* point_ctx,
                            char* errbuf, size_t errlen);

#ifdef __cplusplus
}
#endif
#endif /* INSIDER_B200_H */
