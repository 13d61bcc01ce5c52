// TEST INFRASTRUCTURE. A small, eager re-implementation of the part of the Armadillo / Rcpp API that the reference's own sources
// (src/optimize.cpp, src/coordinate_descent.cpp, src/utils.cpp of kai0511/insider) use, so that those files can be compiled
// WHERE THEY LIE under /root/reference, unmodified, without R, Rcpp, RcppArmadillo, BLAS or LAPACK (none of which exist in this
// image), into oracle/_ref/libinsider_ref.so (oracle/Makefile). The algorithm that runs - every loop, update order, stopping
// rule, index computation - is then the reference's own code; what this header supplies is the arithmetic underneath it:
//   * dense column-major matrices with value semantics, element-wise operators, products (plain triple loops), sums, find / unique
//     / elem / rows / cols / diag views, cube and field containers;
//   * solve(A, B, likely_sympd): Cholesky (falls back to LU with partial pivoting), where Armadillo would call LAPACK;
//   * randperm(n): n draws int(unif_rand() * RAND_MAX) from R's Mersenne-Twister, sorted ascending, indices returned - how
//     RcppArmadillo's alternative RNG drives arma::randperm (restated from memory, like oracle/insider_oracle.cpp mode A);
//   * Rcpp::List / NumericMatrix / Named: just enough for optimize()'s signature and return value.
// Nothing here is product code and nothing here is copied from Armadillo: only its public names and documented semantics.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace arma {

using std::cout;
using std::cerr;
using std::endl;
using std::size_t;

typedef unsigned int uword;     // RcppArmadillo's default (the reference's utils.h spells subview_row<unsigned int>)
typedef int sword;

[[noreturn]] inline void shim_fail(const char* what) { std::fprintf(stderr, "ref_shim: %s\n", what); std::abort(); }

struct SizeMat { uword n_rows, n_cols; };

template <typename T> class Mat;
template <typename T> class Cube;
typedef Mat<double> mat;
typedef Mat<double> vec;
typedef Mat<double> colvec;
typedef Mat<double> rowvec;
typedef Mat<uword> umat;
typedef Mat<uword> uvec;
typedef Cube<double> cube;
template <typename T> using subview_row = Mat<T>;
template <typename T> using subview_col = Mat<T>;

namespace solve_opts { struct opts { int flags; }; static const opts none = {0}; static const opts likely_sympd = {1}; }

// ---- views that write through to their parent (they ARE matrices for every read use) -------------------------------------------
template <typename T> struct ElemView;
template <typename T> struct BlockView;
template <typename T> struct DiagView;

template <typename T>
class Mat {
public:
    uword n_rows = 0, n_cols = 0, n_elem = 0;
    std::vector<T> mem;

    Mat() {}
    explicit Mat(uword n) : n_rows(n), n_cols(1), n_elem(n), mem(n, T(0)) {}                    // column vector, like arma::vec(n)
    Mat(uword r, uword c) : n_rows(r), n_cols(c), n_elem(r * c), mem((size_t)r * c, T(0)) {}
    Mat(const SizeMat& s) : Mat(s.n_rows, s.n_cols) {}
    Mat(T* aux, uword r, uword c, bool /*copy_aux_mem*/ = true, bool /*strict*/ = false) : n_rows(r), n_cols(c), n_elem(r * c), mem(aux, aux + (size_t)r * c) {}
    Mat(const T* aux, uword r, uword c) : n_rows(r), n_cols(c), n_elem(r * c), mem(aux, aux + (size_t)r * c) {}
    Mat(const Mat&) = default;
    Mat(Mat&&) = default;
    // assignment copies the VALUE only (a view assigned to a matrix must not drag its parent pointer along)
    Mat& operator=(const Mat& o) { n_rows = o.n_rows; n_cols = o.n_cols; n_elem = o.n_elem; mem = o.mem; return *this; }
    Mat& operator=(Mat&& o) { n_rows = o.n_rows; n_cols = o.n_cols; n_elem = o.n_elem; mem = std::move(o.mem); return *this; }
    virtual ~Mat() {}

    void set_size(uword r, uword c) { n_rows = r; n_cols = c; n_elem = r * c; mem.assign((size_t)r * c, T(0)); }
    bool is_vec() const { return n_rows == 1 || n_cols == 1; }
    bool is_empty() const { return n_elem == 0; }
    uword size() const { return n_elem; }
    T* memptr() { return mem.data(); }
    const T* memptr() const { return mem.data(); }
    T* begin() { return mem.data(); }
    const T* begin() const { return mem.data(); }

    T& operator()(uword i) { if (i >= n_elem) shim_fail("index out of range"); return mem[i]; }
    const T& operator()(uword i) const { if (i >= n_elem) shim_fail("index out of range"); return mem[i]; }
    T& operator[](uword i) { return mem[i]; }
    const T& operator[](uword i) const { return mem[i]; }
    T& operator()(uword r, uword c) { if (r >= n_rows || c >= n_cols) shim_fail("index out of range"); return mem[(size_t)c * n_rows + r]; }
    const T& operator()(uword r, uword c) const { if (r >= n_rows || c >= n_cols) shim_fail("index out of range"); return mem[(size_t)c * n_rows + r]; }
    T& at(uword r, uword c) { return mem[(size_t)c * n_rows + r]; }
    const T& at(uword r, uword c) const { return mem[(size_t)c * n_rows + r]; }

    Mat& zeros() { std::fill(mem.begin(), mem.end(), T(0)); return *this; }
    Mat& ones() { std::fill(mem.begin(), mem.end(), T(1)); return *this; }
    Mat& zeros(uword r, uword c) { set_size(r, c); return *this; }
    Mat& fill(T v) { std::fill(mem.begin(), mem.end(), v); return *this; }

    Mat t() const { Mat o(n_cols, n_rows); for (uword c = 0; c < n_cols; ++c) for (uword r = 0; r < n_rows; ++r) o.at(c, r) = at(r, c); return o; }

    // ---- reads that return copies
    Mat row(uword r) const { if (r >= n_rows) shim_fail("row out of range"); Mat o(1, n_cols); for (uword c = 0; c < n_cols; ++c) o.mem[c] = at(r, c); return o; }
    Mat col(uword c) const { if (c >= n_cols) shim_fail("col out of range"); Mat o(n_rows, 1); for (uword r = 0; r < n_rows; ++r) o.mem[r] = at(r, c); return o; }
    Mat rows(const Mat<uword>& idx) const {
        Mat o(idx.n_elem, n_cols);
        for (uword i = 0; i < idx.n_elem; ++i) { if (idx.mem[i] >= n_rows) shim_fail("rows(): index out of range"); for (uword c = 0; c < n_cols; ++c) o.at(i, c) = at(idx.mem[i], c); }
        return o;
    }
    Mat cols(const Mat<uword>& idx) const {
        Mat o(n_rows, idx.n_elem);
        for (uword j = 0; j < idx.n_elem; ++j) { if (idx.mem[j] >= n_cols) shim_fail("cols(): index out of range"); for (uword r = 0; r < n_rows; ++r) o.at(r, j) = at(r, idx.mem[j]); }
        return o;
    }
    Mat elem(const Mat<uword>& idx) const { Mat o(idx.n_elem, 1); for (uword i = 0; i < idx.n_elem; ++i) { if (idx.mem[i] >= n_elem) shim_fail("elem(): index out of range"); o.mem[i] = mem[idx.mem[i]]; } return o; }
    Mat operator()(const Mat<uword>& idx) const { return elem(idx); }
    Mat operator()(const Mat<uword>& ri, const Mat<uword>& ci) const {
        Mat o(ri.n_elem, ci.n_elem);
        for (uword j = 0; j < ci.n_elem; ++j) for (uword i = 0; i < ri.n_elem; ++i) o.at(i, j) = (*this)(ri.mem[i], ci.mem[j]);
        return o;
    }
    // ---- writable views (non-const objects)
    BlockView<T> row(uword r);
    BlockView<T> col(uword c);
    ElemView<T> elem(const Mat<uword>& idx);
    ElemView<T> operator()(const Mat<uword>& idx);
    DiagView<T> diag();
    Mat diag() const { const uword n = std::min(n_rows, n_cols); Mat o(n, 1); for (uword i = 0; i < n; ++i) o.mem[i] = at(i, i); return o; }

    // ---- compound assignment
    Mat& operator+=(const Mat& o) { same(o); for (size_t i = 0; i < mem.size(); ++i) mem[i] += o.mem[i]; return *this; }
    Mat& operator-=(const Mat& o) { same(o); for (size_t i = 0; i < mem.size(); ++i) mem[i] -= o.mem[i]; return *this; }
    Mat& operator%=(const Mat& o) { same(o); for (size_t i = 0; i < mem.size(); ++i) mem[i] *= o.mem[i]; return *this; }
    Mat& operator+=(T s) { for (auto& v : mem) v += s; return *this; }
    Mat& operator-=(T s) { for (auto& v : mem) v -= s; return *this; }
    Mat& operator*=(T s) { for (auto& v : mem) v *= s; return *this; }
    Mat& operator/=(T s) { for (auto& v : mem) v /= s; return *this; }
    void same(const Mat& o) const { if (n_rows != o.n_rows || n_cols != o.n_cols) shim_fail("element-wise operation on matrices of different size"); }
};

template <typename T>
struct BlockView : Mat<T> {                     // one row or one column of `parent`
    Mat<T>* parent; uword idx; bool is_row;
    BlockView(Mat<T>* p, uword i, bool r) : Mat<T>(r ? static_cast<const Mat<T>*>(p)->row(i) : static_cast<const Mat<T>*>(p)->col(i)), parent(p), idx(i), is_row(r) {}
    void push() {
        if (is_row) { if (this->n_elem != parent->n_cols) shim_fail("row assignment of wrong length"); for (uword c = 0; c < parent->n_cols; ++c) parent->at(idx, c) = this->mem[c]; }
        else { if (this->n_elem != parent->n_rows) shim_fail("column assignment of wrong length"); for (uword r = 0; r < parent->n_rows; ++r) parent->at(r, idx) = this->mem[r]; }
    }
    BlockView& operator=(const Mat<T>& v) {
        if (is_row ? !(v.n_rows == 1 && v.n_cols == parent->n_cols) : !(v.n_cols == 1 && v.n_rows == parent->n_rows)) shim_fail("subview assignment: incompatible dimensions");
        Mat<T>::operator=(v); push(); return *this;
    }
    BlockView& operator=(const BlockView& v) { return operator=(static_cast<const Mat<T>&>(v)); }
    BlockView& operator+=(const Mat<T>& v) { Mat<T>::operator+=(v); push(); return *this; }
    BlockView& operator-=(const Mat<T>& v) { Mat<T>::operator-=(v); push(); return *this; }
    BlockView& zeros() { Mat<T>::zeros(); push(); return *this; }
    BlockView& ones() { Mat<T>::ones(); push(); return *this; }
};
template <typename T>
struct ElemView : Mat<T> {
    Mat<T>* parent; Mat<uword> idx;
    ElemView(Mat<T>* p, const Mat<uword>& i) : Mat<T>(static_cast<const Mat<T>*>(p)->elem(i)), parent(p), idx(i) {}
    void push() { for (uword i = 0; i < idx.n_elem; ++i) parent->mem[idx.mem[i]] = this->mem[i]; }
    ElemView& operator=(const Mat<T>& v) { if (v.n_elem != idx.n_elem) shim_fail("elem assignment of wrong length"); for (uword i = 0; i < idx.n_elem; ++i) this->mem[i] = v.mem[i]; push(); return *this; }
    ElemView& zeros() { Mat<T>::zeros(); push(); return *this; }
    ElemView& ones() { Mat<T>::ones(); push(); return *this; }
    ElemView& fill(T v) { Mat<T>::fill(v); push(); return *this; }
};
template <typename T>
struct DiagView : Mat<T> {
    Mat<T>* parent;
    explicit DiagView(Mat<T>* p) : Mat<T>(static_cast<const Mat<T>*>(p)->diag()), parent(p) {}
    void push() { for (uword i = 0; i < this->n_elem; ++i) parent->at(i, i) = this->mem[i]; }
    DiagView& operator+=(T s) { Mat<T>::operator+=(s); push(); return *this; }
    DiagView& operator-=(T s) { Mat<T>::operator-=(s); push(); return *this; }
    DiagView& operator+=(const Mat<T>& v) { Mat<T>::operator+=(v); push(); return *this; }
    DiagView& zeros() { Mat<T>::zeros(); push(); return *this; }
    DiagView& ones() { Mat<T>::ones(); push(); return *this; }
};
template <typename T> BlockView<T> Mat<T>::row(uword r) { return BlockView<T>(this, r, true); }
template <typename T> BlockView<T> Mat<T>::col(uword c) { return BlockView<T>(this, c, false); }
template <typename T> ElemView<T> Mat<T>::elem(const Mat<uword>& idx) { return ElemView<T>(this, idx); }
template <typename T> ElemView<T> Mat<T>::operator()(const Mat<uword>& idx) { return ElemView<T>(this, idx); }
template <typename T> DiagView<T> Mat<T>::diag() { return DiagView<T>(this); }

// ---- cube, field ------------------------------------------------------------------------------------------------------------------
template <typename T>
class Cube {
public:
    uword n_rows = 0, n_cols = 0, n_slices = 0;
    std::vector<Mat<T>> s;
    Cube() {}
    Cube(uword r, uword c, uword n) : n_rows(r), n_cols(c), n_slices(n), s(n, Mat<T>(r, c)) {}
    Mat<T>& slice(uword i) { if (i >= n_slices) shim_fail("slice out of range"); return s[i]; }
    const Mat<T>& slice(uword i) const { if (i >= n_slices) shim_fail("slice out of range"); return s[i]; }
    Cube slices(const Mat<uword>& idx) const { Cube o(n_rows, n_cols, idx.n_elem); for (uword i = 0; i < idx.n_elem; ++i) o.s[i] = slice(idx.mem[i]); return o; }
};
template <typename OT>
class field {
public:
    uword n_elem = 0;
    std::vector<OT> v;
    field() {}
    explicit field(uword n) : n_elem(n), v(n) {}
    OT& operator()(uword i) { if (i >= n_elem) shim_fail("field index out of range"); return v[i]; }
    const OT& operator()(uword i) const { if (i >= n_elem) shim_fail("field index out of range"); return v[i]; }
};

// ---- generators --------------------------------------------------------------------------------------------------------------------
inline mat zeros(uword r, uword c) { return mat(r, c); }
inline mat zeros(const SizeMat& s) { return mat(s.n_rows, s.n_cols); }
inline vec zeros(uword n) { return vec(n); }
template <typename MT> inline MT zeros(uword n) { return MT(n); }
template <typename MT> inline MT zeros(uword r, uword c) { return MT(r, c); }
inline vec ones(uword n) { vec o(n); o.ones(); return o; }
inline mat ones(uword r, uword c) { mat o(r, c); o.ones(); return o; }
template <typename MT> inline MT ones(uword n) { MT o(n); o.ones(); return o; }
template <typename T> inline SizeMat size(const Mat<T>& X) { return SizeMat{X.n_rows, X.n_cols}; }
inline std::ostream& operator<<(std::ostream& os, const SizeMat& s) { return os << s.n_rows << 'x' << s.n_cols; }

// ---- element-wise arithmetic -----------------------------------------------------------------------------------------------------------
#define SHIM_BINOP(OP)                                                                                                                \
    template <typename T> inline Mat<T> operator OP(const Mat<T>& a, const Mat<T>& b) { a.same(b); Mat<T> o(a.n_rows, a.n_cols);      \
        for (size_t i = 0; i < o.mem.size(); ++i) o.mem[i] = a.mem[i] OP b.mem[i]; return o; }
SHIM_BINOP(+)
SHIM_BINOP(-)
#undef SHIM_BINOP
template <typename T> inline Mat<T> operator%(const Mat<T>& a, const Mat<T>& b) { a.same(b); Mat<T> o(a.n_rows, a.n_cols); for (size_t i = 0; i < o.mem.size(); ++i) o.mem[i] = a.mem[i] * b.mem[i]; return o; }
template <typename T> inline Mat<T> operator-(const Mat<T>& a) { Mat<T> o(a.n_rows, a.n_cols); for (size_t i = 0; i < o.mem.size(); ++i) o.mem[i] = -a.mem[i]; return o; }
template <typename T, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
inline Mat<T> operator*(const Mat<T>& a, S s) { Mat<T> o(a.n_rows, a.n_cols); for (size_t i = 0; i < o.mem.size(); ++i) o.mem[i] = a.mem[i] * (T)s; return o; }
template <typename T, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
inline Mat<T> operator*(S s, const Mat<T>& a) { return a * s; }
template <typename T, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
inline Mat<T> operator/(const Mat<T>& a, S s) { Mat<T> o(a.n_rows, a.n_cols); for (size_t i = 0; i < o.mem.size(); ++i) o.mem[i] = a.mem[i] / (T)s; return o; }
template <typename T, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
inline Mat<T> operator+(const Mat<T>& a, S s) { Mat<T> o(a.n_rows, a.n_cols); for (size_t i = 0; i < o.mem.size(); ++i) o.mem[i] = a.mem[i] + (T)s; return o; }
template <typename T, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
inline Mat<T> operator-(const Mat<T>& a, S s) { Mat<T> o(a.n_rows, a.n_cols); for (size_t i = 0; i < o.mem.size(); ++i) o.mem[i] = a.mem[i] - (T)s; return o; }
// matrix product (a 1 x 1 operand acts as a scalar, as in Armadillo's as_scalar-style chains)
template <typename T> inline Mat<T> operator*(const Mat<T>& a, const Mat<T>& b) {
    if (a.n_cols != b.n_rows) shim_fail("matrix product: inner dimensions differ");
    Mat<T> o(a.n_rows, b.n_cols);
    for (uword j = 0; j < b.n_cols; ++j)
        for (uword k = 0; k < a.n_cols; ++k) {
            const T bkj = b.at(k, j);
            if (bkj == T(0)) continue;
            const T* ac = a.mem.data() + (size_t)k * a.n_rows;
            T* oc = o.mem.data() + (size_t)j * a.n_rows;
            for (uword i = 0; i < a.n_rows; ++i) oc[i] += ac[i] * bkj;
        }
    return o;
}
// relational operators produce 0/1 masks
#define SHIM_REL(OP)                                                                                                                   \
    template <typename T, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>                          \
    inline Mat<uword> operator OP(const Mat<T>& a, S s) { Mat<uword> o(a.n_rows, a.n_cols);                                            \
        for (size_t i = 0; i < o.mem.size(); ++i) o.mem[i] = (a.mem[i] OP (T)s) ? 1u : 0u; return o; }
SHIM_REL(==)
SHIM_REL(!=)
SHIM_REL(<)
SHIM_REL(>)
SHIM_REL(<=)
SHIM_REL(>=)
#undef SHIM_REL

// ---- functions ------------------------------------------------------------------------------------------------------------------------
template <typename T> inline Mat<T> trans(const Mat<T>& a) { return a.t(); }
template <typename T> inline Mat<T> square(const Mat<T>& a) { Mat<T> o(a.n_rows, a.n_cols); for (size_t i = 0; i < o.mem.size(); ++i) o.mem[i] = a.mem[i] * a.mem[i]; return o; }
template <typename T, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
inline Mat<T> pow(const Mat<T>& a, S p) { Mat<T> o(a.n_rows, a.n_cols); for (size_t i = 0; i < o.mem.size(); ++i) o.mem[i] = (p == S(2)) ? a.mem[i] * a.mem[i] : (T)std::pow((double)a.mem[i], (double)p); return o; }
template <typename T> inline Mat<T> abs(const Mat<T>& a) { Mat<T> o(a.n_rows, a.n_cols); for (size_t i = 0; i < o.mem.size(); ++i) o.mem[i] = a.mem[i] < T(0) ? -a.mem[i] : a.mem[i]; return o; }
inline double sign(double x) { return (x > 0.0) ? 1.0 : ((x < 0.0) ? -1.0 : 0.0); }
template <typename T> inline T accu(const Mat<T>& a) { T s = T(0); for (const auto& v : a.mem) s += v; return s; }
template <typename T> inline T dot(const Mat<T>& a, const Mat<T>& b) { if (a.n_elem != b.n_elem) shim_fail("dot: different lengths"); T s = T(0); for (size_t i = 0; i < a.mem.size(); ++i) s += a.mem[i] * b.mem[i]; return s; }
template <typename T> inline T as_scalar(const Mat<T>& a) { if (a.n_elem != 1) shim_fail("as_scalar: not 1 x 1"); return a.mem[0]; }
// one-argument sum / mean / max: every use in the reference is on a vector, where Armadillo returns a scalar
template <typename T> inline T sum(const Mat<T>& a) { if (!a.is_vec() && a.n_elem) shim_fail("sum(X) of a matrix: only vectors are supported by the shim"); return accu(a); }
template <typename T> inline Mat<T> sum(const Mat<T>& a, int dim) {
    if (dim == 0) { Mat<T> o(1, a.n_cols); for (uword c = 0; c < a.n_cols; ++c) for (uword r = 0; r < a.n_rows; ++r) o.mem[c] += a.at(r, c); return o; }
    Mat<T> o(a.n_rows, 1); for (uword c = 0; c < a.n_cols; ++c) for (uword r = 0; r < a.n_rows; ++r) o.mem[r] += a.at(r, c); return o;
}
template <typename T> inline Mat<T> sum(const Cube<T>& q, int dim) {
    if (dim != 2) shim_fail("sum(cube, dim): only dim = 2 is supported by the shim");
    Mat<T> o(q.n_rows, q.n_cols); for (uword i = 0; i < q.n_slices; ++i) o += q.s[i]; return o;
}
template <typename T> inline double mean(const Mat<T>& a) { if (!a.is_vec() && a.n_elem) shim_fail("mean(X) of a matrix"); return a.n_elem ? (double)accu(a) / a.n_elem : 0.0; }
template <typename T> inline T max(const Mat<T>& a) { if (!a.n_elem) shim_fail("max of an empty object"); T m = a.mem[0]; for (const auto& v : a.mem) if (v > m) m = v; return m; }
template <typename T> inline T min(const Mat<T>& a) { if (!a.n_elem) shim_fail("min of an empty object"); T m = a.mem[0]; for (const auto& v : a.mem) if (v < m) m = v; return m; }
template <typename T> inline double norm(const Mat<T>& a, const char* method) {
    if (std::strcmp(method, "F") != 0 && std::strcmp(method, "fro") != 0) shim_fail("norm: only \"F\" is supported by the shim");
    double s = 0.0; for (const auto& v : a.mem) s += (double)v * (double)v; return std::sqrt(s);
}
template <typename T> inline Mat<uword> find(const Mat<T>& a) { std::vector<uword> ix; for (uword i = 0; i < a.n_elem; ++i) if (a.mem[i] != T(0)) ix.push_back(i); Mat<uword> o((uword)ix.size()); o.mem = ix; return o; }
template <typename T> inline Mat<T> unique(const Mat<T>& a) { std::vector<T> v = a.mem; std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); Mat<T> o((uword)v.size()); o.mem = v; return o; }
template <typename T> inline bool approx_equal(const Mat<T>& a, const Mat<T>& b, const char* method, double tol) {
    if (std::strcmp(method, "absdiff") != 0) shim_fail("approx_equal: only \"absdiff\" is supported by the shim");
    if (a.n_rows != b.n_rows || a.n_cols != b.n_cols) return false;
    for (size_t i = 0; i < a.mem.size(); ++i) { const double d = (double)a.mem[i] - (double)b.mem[i]; if ((d < 0 ? -d : d) > tol) return false; }
    return true;
}

// solve(A, B, likely_sympd): Cholesky A = L L' with forward / back substitution; LU with partial pivoting if A is not positive definite
inline mat solve(const mat& A, const mat& B, const solve_opts::opts& = solve_opts::none) {
    if (A.n_rows != A.n_cols || A.n_rows != B.n_rows) shim_fail("solve: incompatible dimensions");
    const uword n = A.n_rows, m = B.n_cols;
    mat L(n, n);
    bool spd = true;
    for (uword j = 0; j < n && spd; ++j) {
        double d = A.at(j, j);
        for (uword k = 0; k < j; ++k) d -= L.at(j, k) * L.at(j, k);
        if (!(d > 0.0)) { spd = false; break; }
        const double ljj = std::sqrt(d);
        L.at(j, j) = ljj;
        for (uword i = j + 1; i < n; ++i) {
            double s = A.at(i, j);
            for (uword k = 0; k < j; ++k) s -= L.at(i, k) * L.at(j, k);
            L.at(i, j) = s / ljj;
        }
    }
    mat X = B;
    if (spd) {
        for (uword c = 0; c < m; ++c) {
            for (uword i = 0; i < n; ++i) { double s = X.at(i, c); for (uword k = 0; k < i; ++k) s -= L.at(i, k) * X.at(k, c); X.at(i, c) = s / L.at(i, i); }
            for (uword ii = n; ii-- > 0;) { double s = X.at(ii, c); for (uword k = ii + 1; k < n; ++k) s -= L.at(k, ii) * X.at(k, c); X.at(ii, c) = s / L.at(ii, ii); }
        }
        return X;
    }
    mat M = A;
    for (uword j = 0; j < n; ++j) {
        uword p = j; double best = std::fabs(M.at(j, j));
        for (uword i = j + 1; i < n; ++i) if (std::fabs(M.at(i, j)) > best) { best = std::fabs(M.at(i, j)); p = i; }
        if (best == 0.0) shim_fail("solve: singular matrix");
        if (p != j) { for (uword c = 0; c < n; ++c) std::swap(M.at(j, c), M.at(p, c)); for (uword c = 0; c < m; ++c) std::swap(X.at(j, c), X.at(p, c)); }
        for (uword i = j + 1; i < n; ++i) {
            const double f = M.at(i, j) / M.at(j, j);
            if (f == 0.0) continue;
            for (uword c = j; c < n; ++c) M.at(i, c) -= f * M.at(j, c);
            for (uword c = 0; c < m; ++c) X.at(i, c) -= f * X.at(j, c);
        }
    }
    for (uword c = 0; c < m; ++c)
        for (uword ii = n; ii-- > 0;) { double s = X.at(ii, c); for (uword k = ii + 1; k < n; ++k) s -= M.at(ii, k) * X.at(k, c); X.at(ii, c) = s / M.at(ii, ii); }
    return X;
}

// ---- R's default RNG (Mersenne-Twister, set.seed() scrambling, unif_rand: R's src/main/RNG.c restated) and arma::randperm on top of it ----
struct RStream {
    uint32_t mt[624]; int mti = 625;
    void set_seed(uint32_t seed) {
        for (int j = 0; j < 50; ++j) seed = 69069u * seed + 1u;
        for (int j = 0; j < 625; ++j) { seed = 69069u * seed + 1u; if (j > 0) mt[j - 1] = seed; }
        mti = 624;
    }
    uint32_t genrand() {
        static const uint32_t mag01[2] = {0x0u, 0x9908b0dfu};
        uint32_t y;
        if (mti >= 624) {
            if (mti == 625) set_seed(4357u);
            int kk;
            for (kk = 0; kk < 624 - 397; ++kk) { y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu); mt[kk] = mt[kk + 397] ^ (y >> 1) ^ mag01[y & 1u]; }
            for (; kk < 623; ++kk) { y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu); mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ mag01[y & 1u]; }
            y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
            mt[623] = mt[396] ^ (y >> 1) ^ mag01[y & 1u];
            mti = 0;
        }
        y = mt[mti++];
        y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
        return y;
    }
    double unif_rand() {
        const double i2_32m1 = 2.328306437080797e-10;
        const double x = (double)genrand() * 2.3283064365386963e-10;
        if (x <= 0.0) return 0.5 * i2_32m1;
        if ((1.0 - x) <= 0.0) return 1.0 - 0.5 * i2_32m1;
        return x;
    }
};
inline RStream& r_stream() { static RStream s; return s; }
inline unsigned long long& randperm_calls() { static unsigned long long n = 0; return n; }
inline uvec randperm(uword n) {
    ++randperm_calls();
    std::vector<std::pair<uint32_t, uword>> key(n);
    for (uword i = 0; i < n; ++i) key[i] = {(uint32_t)(int)(2147483647.0 * r_stream().unif_rand()), i};
    std::stable_sort(key.begin(), key.end(), [](const std::pair<uint32_t, uword>& a, const std::pair<uint32_t, uword>& b) { return a.first < b.first; });
    uvec o(n);
    for (uword i = 0; i < n; ++i) o.mem[i] = key[i].second;
    return o;
}

}  // namespace arma

// ---- Rcpp: what optimize()'s signature and return statement need ---------------------------------------------------------------------------------
namespace Rcpp {

struct NumericMatrix {
    std::shared_ptr<std::vector<double>> data; int r = 0, c = 0;
    NumericMatrix() {}
    NumericMatrix(const double* p, int rows, int cols) : data(std::make_shared<std::vector<double>>(p, p + (size_t)rows * cols)), r(rows), c(cols) {}
    double* begin() { return data->data(); }
    int nrow() const { return r; }
    int ncol() const { return c; }
};

class List;
struct Value {
    int kind = 0;                         // 1 matrix, 2 scalar, 3 list
    arma::mat m; double d = 0.0; std::shared_ptr<List> l;
    Value() {}
    Value(const arma::mat& x) : kind(1), m(x) {}
    Value(double x) : kind(2), d(x) {}
    Value(const List& x);
};
struct NamedValue { std::string name; Value v; };
struct Named {
    std::string name;
    explicit Named(const std::string& n) : name(n) {}
    template <typename X> NamedValue operator=(const X& x) const { return NamedValue{name, Value(x)}; }
};
class List {
public:
    std::vector<NumericMatrix> items;                     // positional (the caller's list of factor matrices)
    std::map<std::string, Value> named;
    NumericMatrix operator[](int i) const { if (i < 0 || (size_t)i >= items.size()) arma::shim_fail("List index out of range"); return items[(size_t)i]; }
    NumericMatrix operator[](unsigned int i) const { return (*this)[(int)i]; }
    Value& operator[](const std::string& k) { return named[k]; }
    int size() const { return (int)items.size(); }
    template <typename... A> static List create(const A&... a) { List l; const NamedValue nv[] = {a...}; for (const auto& x : nv) l.named[x.name] = x.v; return l; }
};
inline Value::Value(const List& x) : kind(3), l(std::make_shared<List>(x)) {}

}  // namespace Rcpp
