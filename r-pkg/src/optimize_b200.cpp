// Rcpp shim: replaces optimize() of the reference (src/optimize.cpp:256-422) by a call into libinsider_b200.
// Same exported name, same 16 arguments, same returned list (row_matrices{factor0..}, column_factor, train_rmse,
// test_rmse, loss), same in-place update of the factor matrices (src/optimize.cpp:283-284).
// NOT compiled in the insider_b200 repository (no R toolchain there); see r-pkg/README.md.
// [[Rcpp::plugins("cpp17")]]
#include <Rcpp.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>

#include "insider_b200.h"

namespace {

insider_ctx* ctx_singleton() {
    static insider_ctx* ctx = nullptr;
    if (!ctx) {
        char err[512] = "";
        const char* dev = std::getenv("INSIDER_B200_DEVICE");
        if (insider_b200_ctx_create(&ctx, dev ? std::atoi(dev) : 0, err, sizeof err) != INSIDER_OK) Rcpp::stop(err);
    }
    return ctx;
}

struct Marshalled {
    insider_problem pb{};
    insider_factors fac{};
    insider_options opt{};
    std::vector<int32_t> levels;
    std::vector<double*> fptr;
    std::vector<int32_t> frows;
};

void marshal(Marshalled& m, Rcpp::NumericMatrix data, Rcpp::List cfd_factors, Rcpp::NumericMatrix column_factor,
             Rcpp::NumericMatrix cfd_indicators, Rcpp::NumericMatrix ctns_confounder, Rcpp::IntegerMatrix train_indicator,
             Rcpp::IntegerMatrix test_indicator, int inc_continuous, int latent_dim, double lambda1, double lambda2, double alpha,
             int tuning, double global_tol, double sub_tol, unsigned int max_iter) {
    const int N = data.nrow(), P = data.ncol(), C = cfd_indicators.ncol();
    if (latent_dim < 1 || latent_dim > 32)                       // before anything is uploaded (INSIDER_ERR_UNSUPPORTED otherwise)
        Rcpp::stop("insider (B200 back end): latent_dim = %d is not supported; libinsider_b200 handles latent_dim 1..32", latent_dim);
    m.levels.resize((size_t)N * C);                              // cbind() made the indicators double (R/insider.R:40,43)
    for (size_t i = 0; i < m.levels.size(); ++i) m.levels[i] = (int32_t)cfd_indicators[i];
    m.pb.N = N; m.pb.P = P; m.pb.C = C; m.pb.Q = ctns_confounder.ncol(); m.pb.inc_continuous = inc_continuous;
    m.pb.mask_kind = INSIDER_MASK_INT32;                         // R integer matrices (R/insider.R:57-58)
    m.pb.Y = data.begin(); m.pb.levels = m.levels.data(); m.pb.X = ctns_confounder.begin();
    m.pb.train = train_indicator.begin(); m.pb.test = test_indicator.begin();
    const int nf = cfd_factors.size();
    m.fptr.resize(nf); m.frows.resize(nf);
    for (int c = 0; c < nf; ++c) { Rcpp::NumericMatrix f = cfd_factors[c]; m.fptr[c] = f.begin(); m.frows[c] = f.nrow(); }
    m.fac.K = latent_dim; m.fac.n_factors = nf; m.fac.factors = m.fptr.data(); m.fac.factor_rows = m.frows.data();
    m.fac.column_factor = column_factor.begin();
    insider_b200_default_options(&m.opt);
    m.opt.lambda1 = lambda1; m.opt.lambda2 = lambda2; m.opt.alpha = alpha; m.opt.tuning = tuning;
    m.opt.global_tol = global_tol; m.opt.sub_tol = sub_tol; m.opt.max_iter = max_iter; m.opt.verbose = 1;
    {   // set.seed() keeps controlling the fit: the permutation seed is drawn from R's RNG (RcppExports.cpp:90 RNGScope)
        Rcpp::RNGScope scope;
        m.opt.seed = (uint64_t)(R::unif_rand() * 9007199254740992.0);
    }
}

Rcpp::List as_r_list(Rcpp::List cfd_factors, Rcpp::NumericMatrix column_factor, const insider_result& res) {
    Rcpp::List row_matrices;
    for (int c = 0; c < cfd_factors.size(); ++c) row_matrices["factor" + std::to_string(c)] = cfd_factors[c];   // optimize.cpp:413-415
    return Rcpp::List::create(Rcpp::Named("row_matrices") = row_matrices, Rcpp::Named("column_factor") = column_factor,
                              Rcpp::Named("train_rmse") = res.train_rmse, Rcpp::Named("test_rmse") = res.test_rmse,
                              Rcpp::Named("loss") = res.loss);                                                  // optimize.cpp:417-421
}

}  // namespace

// [[Rcpp::export]]
Rcpp::List optimize(Rcpp::NumericMatrix data, Rcpp::List cfd_factors, Rcpp::NumericMatrix column_factor,
                    Rcpp::NumericMatrix cfd_indicators, Rcpp::NumericMatrix ctns_confounder,
                    Rcpp::IntegerMatrix train_indicator, Rcpp::IntegerMatrix test_indicator, int inc_continuous,
                    int latent_dim, double lambda1 = 1.0, double lambda2 = 1.0, double alpha = 0.1, int tuning = 1,
                    double global_tol = 1e-10, double sub_tol = 1e-5, unsigned int max_iter = 10000) {
    Marshalled m;
    marshal(m, data, cfd_factors, column_factor, cfd_indicators, ctns_confounder, train_indicator, test_indicator, inc_continuous,
            latent_dim, lambda1, lambda2, alpha, tuning, global_tol, sub_tol, max_iter);
    insider_result res{};
    char err[512] = "";
    if (insider_b200_optimize(ctx_singleton(), &m.pb, &m.fac, &m.opt, &res, err, sizeof err) != INSIDER_OK) Rcpp::stop(err);
    return as_r_list(cfd_factors, column_factor, res);
}

// Resident variant for tune(): upload once, fit many (R/insider.R:100-174 runs 51 fits on the same data).
// [[Rcpp::export]]
SEXP b200_upload(Rcpp::NumericMatrix data, Rcpp::NumericMatrix cfd_indicators, Rcpp::NumericMatrix ctns_confounder,
                 Rcpp::IntegerMatrix train_indicator, Rcpp::IntegerMatrix test_indicator, int inc_continuous) {
    Marshalled m;
    Rcpp::List none; Rcpp::NumericMatrix dummy(1, 1);
    marshal(m, data, none, dummy, cfd_indicators, ctns_confounder, train_indicator, test_indicator, inc_continuous, 1, 1, 1, 0.1, 1, 1e-10, 1e-5, 0);
    insider_resident* r = nullptr;
    char err[512] = "";
    if (insider_b200_upload(ctx_singleton(), &m.pb, &r, err, sizeof err) != INSIDER_OK) Rcpp::stop(err);
    return Rcpp::XPtr<insider_resident>(r, false);
}

// [[Rcpp::export]]
void b200_release(SEXP handle) { Rcpp::XPtr<insider_resident> p(handle); insider_b200_release(p.get()); }

// [[Rcpp::export]]
Rcpp::List b200_optimize_resident(SEXP handle, Rcpp::List cfd_factors, Rcpp::NumericMatrix column_factor, int latent_dim, double lambda1,
                                  double lambda2, double alpha, int tuning, double global_tol, double sub_tol, unsigned int max_iter) {
    Rcpp::XPtr<insider_resident> p(handle);
    const int nf = cfd_factors.size();
    std::vector<double*> fptr(nf); std::vector<int32_t> frows(nf);
    for (int c = 0; c < nf; ++c) { Rcpp::NumericMatrix f = cfd_factors[c]; fptr[c] = f.begin(); frows[c] = f.nrow(); }
    insider_factors fac{latent_dim, nf, fptr.data(), frows.data(), column_factor.begin()};
    insider_options opt; insider_b200_default_options(&opt);
    opt.lambda1 = lambda1; opt.lambda2 = lambda2; opt.alpha = alpha; opt.tuning = tuning; opt.global_tol = global_tol; opt.sub_tol = sub_tol;
    opt.max_iter = max_iter; opt.verbose = 1;
    { Rcpp::RNGScope scope; opt.seed = (uint64_t)(R::unif_rand() * 9007199254740992.0); }
    insider_result res{}; char err[512] = "";
    if (insider_b200_optimize_resident(ctx_singleton(), p.get(), &fac, &opt, &res, err, sizeof err) != INSIDER_OK) Rcpp::stop(err);
    return as_r_list(cfd_factors, column_factor, res);
}

// ---- the other registered entry points of the reference (src/RcppExports.cpp:112-116) ------------------------------------

// strong_coordinate_descent (src/coordinate_descent.cpp:57-127; R/RcppExports.R:8-10). X and y are accepted and unused: the
// solver works in covariance form on XtX, Xty.
// [[Rcpp::export]]
Rcpp::NumericVector strong_coordinate_descent(Rcpp::NumericMatrix X, Rcpp::NumericVector y, Rcpp::NumericVector wstart, double lambda, double alpha,
                                              Rcpp::NumericMatrix XtX, Rcpp::NumericVector Xty, double tol = 1e-5) {
    const int K = XtX.nrow();
    if (K < 1 || K > 32) Rcpp::stop("insider (B200 back end): %d coordinates are not supported (1..32)", K);
    if (XtX.ncol() != K || Xty.size() != K || wstart.size() != K) Rcpp::stop("strong_coordinate_descent: XtX must be K x K, Xty and wstart of length K");
    Rcpp::NumericVector beta(K);
    uint64_t seed;
    { Rcpp::RNGScope scope; seed = (uint64_t)(R::unif_rand() * 9007199254740992.0); }     // randperm drew from R's RNG (RcppExports.cpp:37)
    char err[512] = "";
    if (insider_b200_strong_cd(ctx_singleton(), K, 1, XtX.begin(), 1, Xty.begin(), wstart.begin(), lambda, alpha, tol, INSIDER_PERM_COUNTER, seed, 0, 0,
                               beta.begin(), nullptr, err, sizeof err) != INSIDER_OK) Rcpp::stop(err);
    return beta;
}

// coordinate_descent (src/coordinate_descent.cpp:10-55): the reference's variant without the strong-rule screen / KKT loop. The
// elastic net is strictly convex for alpha < 1, so both variants stop at the same minimiser to within tol; the symbol is kept
// for useDynLib(.registration = TRUE) and bound to the same solver (nothing in R/insider.R calls it).
// [[Rcpp::export]]
Rcpp::NumericVector coordinate_descent(Rcpp::NumericMatrix X, Rcpp::NumericVector y, Rcpp::NumericVector wstart, double lambda, double alpha,
                                       Rcpp::NumericMatrix XtX, Rcpp::NumericVector Xty, double tol = 1e-5) {
    return strong_coordinate_descent(X, y, wstart, lambda, alpha, XtX, Xty, tol);
}

// optimize_continuous_v2 (src/optimize.cpp:77-137; R/RcppExports.R:16-18): updating_factor is updated in place (rowvec&).
// [[Rcpp::export]]
void optimize_continuous_v2(Rcpp::NumericMatrix data, Rcpp::NumericMatrix indicator, Rcpp::NumericVector updating_factor, Rcpp::NumericMatrix c_factor,
                            Rcpp::NumericVector updating_confd, Rcpp::NumericMatrix gram, double lambda, int tuning) {
    const int N = data.nrow(), P = data.ncol(), K = c_factor.nrow();
    if (K < 1 || K > 32) Rcpp::stop("insider (B200 back end): latent_dim = %d is not supported (1..32)", K);
    if (tuning != 0 && tuning != 1) Rcpp::stop("Parameter tuning should be either 0 or 1!");       // reference: exit(1), src/optimize.cpp:133-135
    if (c_factor.ncol() != P || updating_factor.size() != K || updating_confd.size() != N || indicator.nrow() != N || indicator.ncol() != P)
        Rcpp::stop("optimize_continuous_v2: dimension mismatch");
    (void)gram;                                                  // V V' is recomputed on the device
    char err[512] = "";
    if (insider_b200_optimize_continuous(ctx_singleton(), N, P, K, data.begin(), INSIDER_MASK_DOUBLE, indicator.begin(), updating_factor.begin(),
                                         c_factor.begin(), updating_confd.begin(), lambda, tuning, err, sizeof err) != INSIDER_OK) Rcpp::stop(err);
}

// optimize_continuous (src/optimize.cpp:14-74): the reference's first version of the same update (stops on a loss decrement
// < 1e-3 instead of sum|dw| < 0.1; superseded by _v2 at src/optimize.cpp:345). Bound to the v2 entry; kept for the registration table.
// [[Rcpp::export]]
void optimize_continuous(Rcpp::NumericMatrix data, Rcpp::NumericMatrix indicator, Rcpp::NumericVector updating_factor, Rcpp::NumericMatrix c_factor,
                         Rcpp::NumericVector updating_confd, Rcpp::NumericMatrix gram, double lambda, int tuning) {
    optimize_continuous_v2(data, indicator, updating_factor, c_factor, updating_confd, gram, lambda, tuning);
}

// glm_interaction back end (R/glm_interaction.R:2-30): list(coeff_matrix, pval_matrix)
// [[Rcpp::export]]
Rcpp::List b200_glm_interaction(Rcpp::NumericMatrix residual, Rcpp::IntegerVector interaction_indicator, Rcpp::NumericMatrix column_factor) {
    const int N = residual.nrow(), P = residual.ncol(), K = column_factor.nrow();
    if (column_factor.ncol() != P || interaction_indicator.size() != N) Rcpp::stop("glm_interaction: dimension mismatch");
    int L = 0; for (int i = 0; i < N; ++i) L = std::max(L, interaction_indicator[i]);
    Rcpp::NumericMatrix coeff(L, K), pval(L, K);
    char err[512] = "";
    if (insider_b200_glm_interaction(ctx_singleton(), N, P, K, residual.begin(), L, interaction_indicator.begin(), column_factor.begin(), coeff.begin(),
                                     pval.begin(), err, sizeof err) != INSIDER_OK) Rcpp::stop(err);
    return Rcpp::List::create(coeff, pval);
}

// number of usable CUDA devices as seen by the library (0 when none: every fit then fails with INSIDER_ERR_CUDA - there is no CPU fallback)
// [[Rcpp::export]]
int b200_devices() {
    int n = 0; char err[64] = "";
    for (int d = 0; d < 16; ++d) { insider_ctx* c = nullptr; if (insider_b200_ctx_create(&c, d, err, sizeof err) != INSIDER_OK) break; insider_b200_ctx_destroy(c); ++n; }
    return n;
}
