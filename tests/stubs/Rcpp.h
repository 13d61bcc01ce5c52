// Minimal stand-in for <Rcpp.h>: just enough surface for `g++ -fsyntax-only` to type-check r-pkg/src/*.cpp against
// include/insider_b200.h in an environment without R (tests/test_cabi_cpu.py). It implements nothing.
#pragma once
#include <cstddef>
#include <string>
#include <vector>
typedef struct SEXPREC* SEXP;
namespace R { inline double unif_rand() { return 0.5; } }
namespace Rcpp {
template <typename T> struct Vec {
    std::vector<T> v;
    Vec() {}
    explicit Vec(int n) : v(n) {}
    T* begin() { return v.data(); }
    int size() const { return (int)v.size(); }
    T& operator[](size_t i) { return v[i]; }
};
template <typename T> struct Mat : Vec<T> {
    int r = 0, c = 0;
    Mat() {}
    Mat(int r_, int c_) : Vec<T>(r_ * c_), r(r_), c(c_) {}
    int nrow() const { return r; }
    int ncol() const { return c; }
};
typedef Vec<double> NumericVector;
typedef Vec<int> IntegerVector;
typedef Mat<double> NumericMatrix;
typedef Mat<int> IntegerMatrix;
struct Named { std::string n; explicit Named(const char* s) : n(s) {} template <typename T> Named& operator=(const T&) { return *this; } };
struct List {
    struct Proxy {
        template <typename T> Proxy& operator=(const T&) { return *this; }
        operator NumericMatrix() const { return NumericMatrix(); }
    };
    int size() const { return 0; }
    Proxy operator[](int) { return Proxy(); }
    Proxy operator[](const std::string&) { return Proxy(); }
    template <typename... A> static List create(const A&...) { return List(); }
};
template <typename T> struct XPtr {
    T* p;
    XPtr(T* q, bool) : p(q) {}
    XPtr(SEXP) : p(nullptr) {}
    T* get() { return p; }
    operator SEXP() const { return nullptr; }
};
struct RNGScope {};
template <typename... A> [[noreturn]] inline void stop(const char*, A...) { throw 1; }
}  // namespace Rcpp
