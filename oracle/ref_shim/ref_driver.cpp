// TEST INFRASTRUCTURE: C entry points around the reference's OWN functions (compiled from /root/reference/src by oracle/Makefile
// against oracle/ref_shim/RcppArmadillo.h) so that tests can run them from Python and pin oracle/insider_oracle.cpp against them.
//   ref_optimize                 -> optimize()                    /root/reference/src/optimize.cpp:256-422
//   ref_strong_cd                -> strong_coordinate_descent()   /root/reference/src/coordinate_descent.cpp:57-127
//   ref_optimize_continuous_v2   -> optimize_continuous_v2()      /root/reference/src/optimize.cpp:77-137
// Compiled WITHOUT OpenMP: the reference draws arma::randperm from one global RNG inside its parallel loops, so only the serial
// execution is a function of the seed.
#include <RcppArmadillo.h>

#include <sstream>

using namespace arma;

// the reference's definitions (no headers declare optimize / optimize_continuous_v2: RcppExports.cpp forward-declares them too)
Rcpp::List optimize(const mat& data, Rcpp::List cfd_factors, mat& column_factor, const umat& cfd_indicators, const mat& ctns_confounder,
                    const mat& train_indicator, const mat& test_indicator, const int& inc_continuous, const int latent_dim, const double lambda1,
                    const double lambda2, const double alpha, const int tuning, const double global_tol, const double sub_tol, const unsigned int max_iter);
void optimize_continuous_v2(const mat& data, const mat& indicator, rowvec& updating_factor, const mat& c_factor, const vec& updating_confd,
                            const mat& gram, const double lambda, const int tuning);
vec strong_coordinate_descent(const mat& X, const vec& y, const vec& wstart, const double& lambda, const double& alpha, const mat& XtX,
                              const vec& Xty, const double& tol);

namespace {
struct Quiet {                                   // the reference prints its progress on std::cout
    std::ostringstream sink; std::streambuf* old;
    Quiet() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~Quiet() { std::cout.rdbuf(old); }
};
}  // namespace

extern "C" {

int ref_version(void) { return 1; }
void ref_set_seed(unsigned int seed) { arma::r_stream().set_seed(seed); }
unsigned long long ref_randperm_calls(void) { return arma::randperm_calls(); }
void ref_unif(unsigned int seed, int n, double* out) { arma::RStream r; r.set_seed(seed); for (int i = 0; i < n; ++i) out[i] = r.unif_rand(); }

// All matrices column-major like R. factors: n_factors pointers, factor c is factor_rows[c] x K, updated in place.
// levels: N x C, 1-based (the reference's cfd_indicators, an unsigned matrix). Returns the number of evaluations is not
// available from the reference: out3 = {train_rmse, test_rmse, loss} of its last evaluation.
int ref_optimize(int N, int P, int K, const double* Y, int C, const unsigned int* levels, int Q, const double* X, int inc_continuous,
                 const double* train, const double* test, int n_factors, double** factors, const int* factor_rows, double* V,
                 double lambda1, double lambda2, double alpha, int tuning, double global_tol, double sub_tol, unsigned int max_iter,
                 unsigned int r_seed, double* out3) {
    Quiet q;
    arma::r_stream().set_seed(r_seed);
    mat data(Y, (uword)N, (uword)P);
    umat ind(levels, (uword)N, (uword)C);
    mat ctns = (Q > 0 && X) ? mat(X, (uword)N, (uword)Q) : mat((uword)N, 0u);
    mat tr = train ? mat(train, (uword)N, (uword)P) : mat((uword)N, (uword)P).ones();
    mat te = test ? mat(test, (uword)N, (uword)P) : mat((uword)N, (uword)P);
    mat col(V, (uword)K, (uword)P);
    Rcpp::List fl;
    for (int c = 0; c < n_factors; ++c) fl.items.push_back(Rcpp::NumericMatrix(factors[c], factor_rows[c], K));
    Rcpp::List res = optimize(data, fl, col, ind, ctns, tr, te, inc_continuous, K, lambda1, lambda2, alpha, tuning, global_tol, sub_tol, max_iter);
    Rcpp::List& rows = *res.named["row_matrices"].l;
    for (int c = 0; c < n_factors; ++c) {
        const mat& m = rows.named["factor" + std::to_string(c)].m;
        if ((int)m.n_rows != factor_rows[c] || (int)m.n_cols != K) return 2;
        std::memcpy(factors[c], m.memptr(), sizeof(double) * m.n_elem);
    }
    const mat& vc = res.named["column_factor"].m;
    std::memcpy(V, vc.memptr(), sizeof(double) * vc.n_elem);
    out3[0] = res.named["train_rmse"].d; out3[1] = res.named["test_rmse"].d; out3[2] = res.named["loss"].d;
    return 0;
}

// X: n x K, y: n, XtX: K x K, Xty: K, wstart / beta: K
int ref_strong_cd(int n, int K, const double* X, const double* y, const double* wstart, double lambda, double alpha, const double* XtX,
                  const double* Xty, double tol, unsigned int r_seed, double* beta, unsigned long long* n_sweeps) {
    Quiet q;
    arma::r_stream().set_seed(r_seed);
    const unsigned long long c0 = arma::randperm_calls();
    mat Xm(X, (uword)n, (uword)K), G(XtX, (uword)K, (uword)K);
    vec yv(y, (uword)n, 1u), w(wstart, (uword)K, 1u), xty(Xty, (uword)K, 1u);
    vec b = strong_coordinate_descent(Xm, yv, w, lambda, alpha, G, xty, tol);
    std::memcpy(beta, b.memptr(), sizeof(double) * (size_t)K);
    if (n_sweeps) *n_sweeps = arma::randperm_calls() - c0;          // one randperm per sweep (coordinate_descent.cpp:89)
    return 0;
}

// data: N x P, indicator: N x P (doubles), updating_factor: K (in/out), c_factor: K x P, updating_confd: N, gram: K x K
int ref_optimize_continuous_v2(int N, int P, int K, const double* data, const double* indicator, double* updating_factor, const double* c_factor,
                               const double* updating_confd, const double* gram, double lambda, int tuning) {
    Quiet q;
    mat d(data, (uword)N, (uword)P), ind(indicator, (uword)N, (uword)P), cf(c_factor, (uword)K, (uword)P), g(gram, (uword)K, (uword)K);
    vec x(updating_confd, (uword)N, 1u);
    rowvec uf(updating_factor, 1u, (uword)K);
    optimize_continuous_v2(d, ind, uf, cf, x, g, lambda, tuning);
    std::memcpy(updating_factor, uf.memptr(), sizeof(double) * (size_t)K);
    return 0;
}

}  // extern "C"
