// ORACLE -- TEST INFRASTRUCTURE ONLY (see taylor_oracle.hpp header).
// Flat C API over the templated restatement so that tests/ (ctypes) and bench.py's
// cpu_baseline leg can drive it.  Two instantiations: `orc_f64_*` (T = double) and
// `orc_iv_*` (T = Interval<f64>, data passed as [lo, hi] pairs).
#include "taylor_oracle.hpp"

#include <chrono>
#include <cstdio>

using namespace orc;

static thread_local std::string g_err;
extern "C" const char* orc_last_error() { return g_err.c_str(); }

template <class F> static auto guard(F&& f) -> decltype(f()) {
  try {
    return f();
  } catch (const std::exception& e) {
    g_err = e.what();
    return decltype(f())();
  }
}

template <class T> struct Scalar;  // marshalling of T <-> double[]
template <> struct Scalar<double> {
  static constexpr int W = 1;
  static double load(const double* p) { return p[0]; }
  static void store(double* p, double x) { p[0] = x; }
};
template <> struct Scalar<Interval> {
  static constexpr int W = 2;
  static Interval load(const double* p) { return {p[0], p[1]}; }
  static void store(double* p, const Interval& x) { p[0] = x.lo; p[1] = x.hi; }
};

template <class T> static std::vector<usize> to_vec(const uint64_t* p, int n) {
  std::vector<usize> v(n);
  for (int i = 0; i < n; i++) v[i] = (p[i] == UINT64_MAX) ? UMAX : (usize)p[i];
  return v;
}

template <class T> struct Api {
  using TP = TaylorPoly<T>;
  using TE = TaylorExpansion<T>;
  static TP* make(int ndim, const uint64_t* shape, const uint64_t* degrees, const double* data) {
    return guard([&]() -> TP* {
      Arr<T> a(to_vec<T>(shape, ndim), Num<T>::zero());
      for (usize i = 0; i < a.data.size(); i++) a.data[i] = Scalar<T>::load(data + i * Scalar<T>::W);
      return new TP(std::move(a), to_vec<T>(degrees, ndim));
    });
  }
  static void shape(const TP* t, uint64_t* out) {
    for (usize i = 0; i < t->coeffs.shape.size(); i++) out[i] = t->coeffs.shape[i];
  }
  static void degrees(const TP* t, uint64_t* out) {
    for (usize i = 0; i < t->degrees_p1.size(); i++) out[i] = t->degrees_p1[i] == UMAX ? UINT64_MAX : t->degrees_p1[i];
  }
  static void data(const TP* t, double* out) {
    for (usize i = 0; i < t->coeffs.data.size(); i++) Scalar<T>::store(out + i * Scalar<T>::W, t->coeffs.data[i]);
  }
};

#define ORC_API(P, T)                                                                                         \
  extern "C" {                                                                                                \
  void* P##new(int ndim, const uint64_t* shape, const uint64_t* degrees, const double* data) {                \
    return Api<T>::make(ndim, shape, degrees, data);                                                          \
  }                                                                                                           \
  void P##free(void* t) { delete (TaylorPoly<T>*)t; }                                                         \
  int P##ndim(void* t) { return (int)((TaylorPoly<T>*)t)->coeffs.shape.size(); }                              \
  uint64_t P##len(void* t) { return ((TaylorPoly<T>*)t)->coeffs.data.size(); }                                \
  void P##shape(void* t, uint64_t* o) { Api<T>::shape((TaylorPoly<T>*)t, o); }                                \
  void P##degrees(void* t, uint64_t* o) { Api<T>::degrees((TaylorPoly<T>*)t, o); }                            \
  void P##data(void* t, double* o) { Api<T>::data((TaylorPoly<T>*)t, o); }                                    \
  int P##eq(void* a, void* b) { return *(TaylorPoly<T>*)a == *(TaylorPoly<T>*)b; }                            \
  void* P##add(void* a, void* b) {                                                                            \
    return guard([&] { return new TaylorPoly<T>(tp_add(*(TaylorPoly<T>*)a, *(TaylorPoly<T>*)b)); });          \
  }                                                                                                           \
  void* P##sub(void* a, void* b) {                                                                            \
    return guard([&] { return new TaylorPoly<T>(tp_sub(*(TaylorPoly<T>*)a, *(TaylorPoly<T>*)b)); });          \
  }                                                                                                           \
  void* P##mul(void* a, void* b) {                                                                            \
    return guard([&] { return new TaylorPoly<T>(tp_mul(*(TaylorPoly<T>*)a, *(TaylorPoly<T>*)b)); });          \
  }                                                                                                           \
  void* P##div(void* a, void* b) {                                                                            \
    return guard([&] { return new TaylorPoly<T>(tp_div(*(TaylorPoly<T>*)a, *(TaylorPoly<T>*)b)); });          \
  }                                                                                                           \
  void* P##neg(void* a) { return guard([&] { return new TaylorPoly<T>(tp_neg(*(TaylorPoly<T>*)a)); }); }      \
  void* P##exp(void* a) { return guard([&] { return new TaylorPoly<T>(((TaylorPoly<T>*)a)->exp()); }); }      \
  void* P##log(void* a) { return guard([&] { return new TaylorPoly<T>(((TaylorPoly<T>*)a)->log()); }); }      \
  void* P##pow(void* a, uint32_t e) {                                                                         \
    return guard([&] { return new TaylorPoly<T>(((TaylorPoly<T>*)a)->pow(e)); });                             \
  }                                                                                                           \
  void* P##derivative(void* a, uint64_t v, uint64_t n) {                                                      \
    return guard([&] { return new TaylorPoly<T>(((TaylorPoly<T>*)a)->derivative(v, n)); });                   \
  }                                                                                                           \
  void* P##taylor_expansion_of_coeff(void* a, uint64_t v, uint64_t n) {                                       \
    return guard([&] { return new TaylorPoly<T>(((TaylorPoly<T>*)a)->taylor_expansion_of_coeff(v, n)); });    \
  }                                                                                                           \
  void* P##shift_down(void* a, uint64_t v, uint64_t n) {                                                      \
    return guard([&] { return new TaylorPoly<T>(((TaylorPoly<T>*)a)->shift_down(v, n)); });                   \
  }                                                                                                           \
  void* P##subst_var(void* a, uint64_t v, void* s) {                                                          \
    return guard([&] { return new TaylorPoly<T>(((TaylorPoly<T>*)a)->subst_var(v, *(TaylorPoly<T>*)s)); });   \
  }                                                                                                           \
  void* P##coefficients_of_term(void* a, uint64_t v, uint64_t o) {                                            \
    return guard([&] { return new TaylorPoly<T>(((TaylorPoly<T>*)a)->coefficients_of_term(v, o)); });         \
  }                                                                                                           \
  void* P##taylor_polynomial(void* a, uint64_t v, uint64_t o) {                                               \
    return guard([&] { return new TaylorPoly<T>(((TaylorPoly<T>*)a)->taylor_polynomial(v, o)); });            \
  }                                                                                                           \
  void* P##taylor_polynomial_terms(void* a, uint64_t v, const uint64_t* orders, int n) {                      \
    return guard([&] {                                                                                        \
      return new TaylorPoly<T>(((TaylorPoly<T>*)a)->taylor_polynomial_terms(v, to_vec<T>(orders, n)));        \
    });                                                                                                       \
  }                                                                                                           \
  void* P##truncate_to_degree_p1(void* a, uint64_t d) {                                                       \
    return guard([&] {                                                                                        \
      return new TaylorPoly<T>(((TaylorPoly<T>*)a)->truncate_to_degree_p1(d == UINT64_MAX ? UMAX : d));       \
    });                                                                                                       \
  }                                                                                                           \
  void* P##remove_last_variable(void* a) {                                                                    \
    return guard([&] { return new TaylorPoly<T>(((TaylorPoly<T>*)a)->remove_last_variable()); });             \
  }                                                                                                           \
  void* P##extend_to_dim(void* a, uint64_t ndim, uint64_t d) {                                                \
    return guard([&] {                                                                                        \
      return new TaylorPoly<T>(((TaylorPoly<T>*)a)->extend_to_dim(ndim, d == UINT64_MAX ? UMAX : d));         \
    });                                                                                                       \
  }                                                                                                           \
  void* P##extend(void* a, const uint64_t* ns, int n) {                                                       \
    return guard([&] { return new TaylorPoly<T>(((TaylorPoly<T>*)a)->extend(to_vec<T>(ns, n))); });           \
  }                                                                                                           \
  void* P##zero_with(const uint64_t* d, int n) {                                                              \
    return guard([&] { return new TaylorPoly<T>(TaylorPoly<T>::zero_with(to_vec<T>(d, n))); });               \
  }                                                                                                           \
  void* P##var(uint64_t v, const double* x, uint64_t len) {                                                   \
    return guard([&] { return new TaylorPoly<T>(TaylorPoly<T>::var(v, Scalar<T>::load(x), len)); });          \
  }                                                                                                           \
  void* P##var_at_zero(uint64_t v, uint64_t len) {                                                            \
    return guard([&] { return new TaylorPoly<T>(TaylorPoly<T>::var_at_zero(v, len)); });                      \
  }                                                                                                           \
  void* P##var_with_degrees_p1(uint64_t v, const double* x, const uint64_t* d, int n) {                       \
    return guard([&] {                                                                                        \
      return new TaylorPoly<T>(TaylorPoly<T>::var_with_degrees_p1(v, Scalar<T>::load(x), to_vec<T>(d, n)));   \
    });                                                                                                       \
  }                                                                                                           \
  int P##coefficient(void* a, const uint64_t* idx, int n, double* out) {                                      \
    try {                                                                                                     \
      Scalar<T>::store(out, ((TaylorPoly<T>*)a)->coefficient(to_vec<T>(idx, n)));                             \
      return 0;                                                                                               \
    } catch (const std::exception& e) {                                                                       \
      g_err = e.what();                                                                                       \
      return 1;                                                                                               \
    }                                                                                                         \
  }                                                                                                           \
  void P##constant_term(void* a, double* out) { Scalar<T>::store(out, ((TaylorPoly<T>*)a)->constant_term()); } \
  void P##evaluate_all_one(void* a, double* out) {                                                            \
    Scalar<T>::store(out, ((TaylorPoly<T>*)a)->evaluate_all_one());                                           \
  }                                                                                                           \
  int P##is_zero(void* a) { return ((TaylorPoly<T>*)a)->is_zero(); }                                          \
  int P##is_one(void* a) { return ((TaylorPoly<T>*)a)->is_one(); }                                            \
  /* returns 1 and fills (c, m, v) if linear */                                                               \
  int P##extract_linear(void* a, double* c, double* m, uint64_t* v) {                                         \
    auto l = ((TaylorPoly<T>*)a)->extract_linear();                                                           \
    if (!l) return 0;                                                                                         \
    Scalar<T>::store(c, l->c);                                                                                \
    Scalar<T>::store(m, l->m);                                                                                \
    *v = l->v;                                                                                                \
    return 1;                                                                                                 \
  }                                                                                                           \
  /* ---- univariate TaylorExpansion ---- */                                                                  \
  void* P##te_const(const double* x) { return new TaylorExpansion<T>(TaylorExpansion<T>::constant(Scalar<T>::load(x))); } \
  void* P##te_poly(const double* xs, uint64_t n) {                                                            \
    std::vector<T> v(n, Num<T>::zero());                                                                      \
    for (uint64_t i = 0; i < n; i++) v[i] = Scalar<T>::load(xs + i * Scalar<T>::W);                           \
    return new TaylorExpansion<T>(TaylorExpansion<T>::polynomial(std::move(v)));                              \
  }                                                                                                           \
  void* P##te_var(const double* x, uint64_t order) {                                                          \
    return new TaylorExpansion<T>(TaylorExpansion<T>::var(Scalar<T>::load(x), order));                        \
  }                                                                                                           \
  void P##te_free(void* t) { delete (TaylorExpansion<T>*)t; }                                                 \
  int P##te_is_const(void* t) { return ((TaylorExpansion<T>*)t)->is_const; }                                  \
  uint64_t P##te_len(void* t) {                                                                               \
    auto* e = (TaylorExpansion<T>*)t;                                                                         \
    return e->is_const ? 1 : e->coeffs.size();                                                                \
  }                                                                                                           \
  void P##te_data(void* t, double* out) {                                                                     \
    auto* e = (TaylorExpansion<T>*)t;                                                                         \
    if (e->is_const) { Scalar<T>::store(out, e->c); return; }                                                 \
    for (usize i = 0; i < e->coeffs.size(); i++) Scalar<T>::store(out + i * Scalar<T>::W, e->coeffs[i]);      \
  }                                                                                                           \
  int P##te_eq(void* a, void* b) { return *(TaylorExpansion<T>*)a == *(TaylorExpansion<T>*)b; }               \
  void* P##te_add(void* a, void* b) { return guard([&] { return new TaylorExpansion<T>(te_add(*(TaylorExpansion<T>*)a, *(TaylorExpansion<T>*)b)); }); } \
  void* P##te_sub(void* a, void* b) { return guard([&] { return new TaylorExpansion<T>(te_sub(*(TaylorExpansion<T>*)a, *(TaylorExpansion<T>*)b)); }); } \
  void* P##te_mul(void* a, void* b) { return guard([&] { return new TaylorExpansion<T>(te_mul(*(TaylorExpansion<T>*)a, *(TaylorExpansion<T>*)b)); }); } \
  void* P##te_div(void* a, void* b) { return guard([&] { return new TaylorExpansion<T>(te_div(*(TaylorExpansion<T>*)a, *(TaylorExpansion<T>*)b)); }); } \
  void* P##te_neg(void* a) { return guard([&] { return new TaylorExpansion<T>(te_neg(*(TaylorExpansion<T>*)a)); }); } \
  void* P##te_exp(void* a) { return guard([&] { return new TaylorExpansion<T>(((TaylorExpansion<T>*)a)->exp()); }); } \
  void* P##te_log(void* a) { return guard([&] { return new TaylorExpansion<T>(((TaylorExpansion<T>*)a)->log()); }); } \
  void* P##te_pow(void* a, uint32_t e) { return guard([&] { return new TaylorExpansion<T>(te_pow(*(TaylorExpansion<T>*)a, e)); }); } \
  void* P##te_subst(void* a, void* s) { return guard([&] { return new TaylorExpansion<T>(te_subst(*(TaylorExpansion<T>*)a, *(TaylorExpansion<T>*)s)); }); } \
  void* P##te_taylor_expansion_of_coeff(void* a, uint64_t n) { return guard([&] { return new TaylorExpansion<T>(((TaylorExpansion<T>*)a)->taylor_expansion_of_coeff(n)); }); } \
  int P##te_coeff(void* a, uint64_t order, double* out) {                                                     \
    try { Scalar<T>::store(out, ((TaylorExpansion<T>*)a)->coeff(order)); return 0; }                          \
    catch (const std::exception& e) { g_err = e.what(); return 1; }                                           \
  }                                                                                                           \
  int P##te_derivative(void* a, uint64_t order, double* out) {                                                \
    try { Scalar<T>::store(out, ((TaylorExpansion<T>*)a)->derivative(order)); return 0; }                     \
    catch (const std::exception& e) { g_err = e.what(); return 1; }                                           \
  }                                                                                                           \
  }

ORC_API(orc_f64_, double)
ORC_API(orc_iv_, Interval)

extern "C" {

// Raw general product on contiguous f64 buffers (multivariate_taylor.rs:984-1012), `res` must be
// zero-initialised by the caller.  Used by the parity tests at mid sizes.
void orc_mul_raw(int ndim, const uint64_t* xs, const double* x, const uint64_t* ys, const double* y,
                 const uint64_t* rs, double* r) {
  std::vector<usize> xsv = to_vec<double>(xs, ndim), ysv = to_vec<double>(ys, ndim), rsv = to_vec<double>(rs, ndim);
  View<double> xv{x, xsv.data(), (usize)ndim}, yv{y, ysv.data(), (usize)ndim};
  ViewMut<double> rv{r, rsv.data(), (usize)ndim};
  mul(xv, yv, rv);
}

// Bounded sample of the same product: only the leading-axis output rows listed in `rows` are
// computed (exactly the j-loop of :1001-1010 for those k).  r holds the FULL result buffer; rows
// not listed are left untouched.  Returns the number of MACs executed (trip counts :975-977,
// :1002-1004) so that bench.py can turn the timing into GFLOP/s.
double orc_mul_rows(int ndim, const uint64_t* xs, const double* x, const uint64_t* ys, const double* y,
                    const uint64_t* rs, double* r, const uint64_t* rows, int nrows) {
  std::vector<usize> xsv = to_vec<double>(xs, ndim), ysv = to_vec<double>(ys, ndim), rsv = to_vec<double>(rs, ndim);
  View<double> xv{x, xsv.data(), (usize)ndim}, yv{y, ysv.data(), (usize)ndim};
  ViewMut<double> rv{r, rsv.data(), (usize)ndim};
  double macs = 0;
  std::vector<usize> xt(xsv.begin() + 1, xsv.end()), yt(ysv.begin() + 1, ysv.end()), rt(rsv.begin() + 1, rsv.end());
  double inner = ndim > 1 ? mul_macs(xt, yt, rt) : 1.0;
  for (int i = 0; i < nrows; i++) {
    usize k = rows[i];
    ViewMut<double> z = rv.index0(k);
    usize lo = sat_sub(k + 1, ysv[0]), hi = std::min(k + 1, xsv[0]);
    for (usize j = lo; j < hi; j++) {
      mul(xv.index0(j), yv.index0(k - j), z);
      macs += inner;
    }
  }
  return macs;
}

double orc_mul_macs(int ndim, const uint64_t* xs, const uint64_t* ys, const uint64_t* rs) {
  return mul_macs(to_vec<double>(xs, ndim), to_vec<double>(ys, ndim), to_vec<double>(rs, ndim));
}

double orc_next_up(double x) { return next_up(x); }
double orc_next_down(double x) { return next_down(x); }
}
