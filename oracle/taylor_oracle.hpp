// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement (C++17 templates, single thread, separate multiply/add, no FMA:
// build with -ffp-contract=off) of the reference's dense truncated Taylor arithmetic:
//   * TaylorPoly<T>      -- /root/reference/src/multivariate_taylor.rs
//   * TaylorExpansion<T> -- /root/reference/src/univariate_taylor.rs
//   * Interval<f64>      -- /root/reference/src/interval.rs (+ next_up/next_down of
//                           src/number/f64.rs:127-171) for the --bounds enclosure check
// Each function cites the reference file:line whose semantics it follows.  The reference
// is Rust and cannot be compiled in this image (no cargo/rustc), so this restatement is
// pinned against the literal vectors of the reference's in-file unit tests
// (tests/test_oracle_golden.py).  Storage is a plain row-major std::vector, not ndarray;
// truncations materialise contiguous copies (the reference keeps strided views -- same
// logical content).
//
// Parity pinning status:
//   pinned   : +,-,*,/ ,exp,log,derivative,taylor_expansion_of_coeff,subst_var,mul_linear,
//              shape algebra (reference unit tests, bit-exact f64)
//   unpinned : f64 rounding order of shift_down (third-party ndarray 0.15.6 sum_axis, not
//              vendored; its published algorithm is restated in sum_axis() below) and the
//              Interval<f64> instantiation (no reference fixture uses --bounds).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
// may use anything in this directory.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

namespace orc {

using usize = std::size_t;
constexpr usize UMAX = std::numeric_limits<usize>::max();  // Rust usize::MAX

struct OracleError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
#define ORC_ASSERT(cond, msg)              \
  do {                                     \
    if (!(cond)) throw OracleError(msg);   \
  } while (0)

static inline usize sat_sub(usize a, usize b) { return a > b ? a - b : 0; }

// ---------------------------------------------------------------------------------------
// Scalar layer.  F64: src/number/f64.rs (every op is one IEEE double op, :202-264;
// exp/ln go to libm :53-61; is_zero is `== 0.0` :181-183).
// ---------------------------------------------------------------------------------------
inline double next_up(double x) {  // f64.rs:127-148
  uint64_t bits;
  std::memcpy(&bits, &x, 8);
  const uint64_t inf_bits = 0x7ff0000000000000ULL;
  if (std::isnan(x) || bits == inf_bits) return x;
  uint64_t abs = bits & 0x7fffffffffffffffULL;
  uint64_t next = (abs == 0) ? 0x1ULL : (bits == abs ? bits + 1 : bits - 1);
  double r;
  std::memcpy(&r, &next, 8);
  return r;
}
inline double next_down(double x) {  // f64.rs:150-171
  uint64_t bits;
  std::memcpy(&bits, &x, 8);
  const uint64_t ninf_bits = 0xfff0000000000000ULL;
  if (std::isnan(x) || bits == ninf_bits) return x;
  uint64_t abs = bits & 0x7fffffffffffffffULL;
  uint64_t next = (abs == 0) ? 0x8000000000000001ULL : (bits == abs ? bits - 1 : bits + 1);
  double r;
  std::memcpy(&r, &next, 8);
  return r;
}
// F64::min / F64::max (f64.rs:69-85): `if self < other {self} else {other}` -- NaN in
// `self` makes the comparison false and returns `other`.
inline double fmin_ref(double a, double b) { return a < b ? a : b; }
inline double fmax_ref(double a, double b) { return a > b ? a : b; }

struct Interval {  // interval.rs:11-15
  double lo, hi;
  static Interval exact(double lo, double hi) { return {lo, hi}; }
  static Interval widen(double lo, double hi) { return {next_down(lo), next_up(hi)}; }  // :29-31
  bool contains(double x) const { return lo <= x && x <= hi; }                          // :34-36
  bool is_zero() const { return lo == 0.0 && hi == 0.0; }                                // :100-102
  bool is_one() const { return lo == 1.0 && hi == 1.0; }                                 // :112-114
  bool is_finite() const { return std::isfinite(lo) && std::isfinite(hi); }              // :317-319
  bool is_nan() const { return std::isnan(lo) || std::isnan(hi); }                       // :321-323
  bool operator==(const Interval& o) const { return lo == o.lo && hi == o.hi; }
};
inline Interval operator-(const Interval& a) { return {-a.hi, -a.lo}; }  // :117-124
inline Interval operator+(const Interval& a, const Interval& b) {         // :126-139
  if (a.is_zero()) return b;
  if (b.is_zero()) return a;
  return Interval::widen(a.lo + b.lo, a.hi + b.hi);
}
inline Interval operator-(const Interval& a, const Interval& b) { return a + (-b); }  // :148-155
inline Interval operator*(const Interval& a, const Interval& b) {                      // :164-190
  if ((a.is_zero() && b.is_finite()) || (a.is_finite() && b.is_zero())) return {0.0, 0.0};
  if (a.is_one()) return b;
  if (b.is_one()) return a;
  if ((-a).is_one()) return -b;
  if ((-b).is_one()) return -a;
  double p = a.lo * b.lo, q = a.lo * b.hi, r = a.hi * b.lo, s = a.hi * b.hi;
  return Interval::widen(fmin_ref(fmin_ref(fmin_ref(p, q), r), s),
                         fmax_ref(fmax_ref(fmax_ref(p, q), r), s));
}
inline Interval operator/(const Interval& a, const Interval& b) {  // :199-234
  const double inf = std::numeric_limits<double>::infinity();
  if (a.is_nan() || b.is_nan()) return {std::nan(""), std::nan("")};
  if (a.is_zero() && !b.is_zero()) return a;
  if (b.is_one()) return a;
  double lo = inf, hi = -inf;
  if (b.contains(0.0)) {
    if (0.0 <= a.lo) hi = inf; else lo = -inf;
    if (a.hi <= 0.0) lo = -inf; else hi = inf;
  }
  double p = a.lo / b.lo, q = a.lo / b.hi, r = a.hi / b.lo, s = a.hi / b.hi;
  lo = fmin_ref(fmin_ref(fmin_ref(fmin_ref(lo, p), q), r), s);
  hi = fmax_ref(fmax_ref(fmax_ref(fmax_ref(hi, p), q), r), s);
  return Interval::widen(lo, hi);
}

template <class T> struct Num;
template <> struct Num<double> {
  static double zero() { return 0.0; }
  static double one() { return 1.0; }
  static double from_u32(uint32_t u) { return (double)u; }
  static bool is_zero(double x) { return x == 0.0; }
  static bool is_one(double x) { return x == 1.0; }
  static double exp(double x) { return std::exp(x); }
  static double log(double x) { return std::log(x); }
  static bool eq(double a, double b) { return a == b; }
};
template <> struct Num<Interval> {
  static Interval zero() { return {0.0, 0.0}; }
  static Interval one() { return {1.0, 1.0}; }
  static Interval from_u32(uint32_t u) { return {(double)u, (double)u}; }  // interval.rs:79-84
  static bool is_zero(const Interval& x) { return x.is_zero(); }
  static bool is_one(const Interval& x) { return x.is_one(); }
  static Interval exp(const Interval& x) {  // interval.rs:264-269
    if (x.is_zero()) return one();
    return Interval::widen(std::exp(x.lo), std::exp(x.hi));
  }
  static Interval log(const Interval& x) {  // interval.rs:271-276
    if (x.is_one()) return zero();
    return Interval::widen(std::log(x.lo), std::log(x.hi));
  }
  static bool eq(const Interval& a, const Interval& b) { return a == b; }
};

// ---------------------------------------------------------------------------------------
// Dense row-major array + read-only / mutable views (stand-ins for ndarray ArrayD/ArrayViewD)
// ---------------------------------------------------------------------------------------
inline usize prod(const std::vector<usize>& s) {
  usize p = 1;
  for (usize x : s) p *= x;
  return p;
}

template <class T> struct View {  // contiguous row-major view of `ndim` axes
  const T* p;
  const usize* shape;
  usize ndim;
  usize len() const { usize n = 1; for (usize i = 0; i < ndim; i++) n *= shape[i]; return n; }
  usize stride0() const { usize n = 1; for (usize i = 1; i < ndim; i++) n *= shape[i]; return n; }
  View index0(usize j) const { return {p + j * stride0(), shape + 1, ndim - 1}; }
};
template <class T> struct ViewMut {
  T* p;
  const usize* shape;
  usize ndim;
  usize len() const { usize n = 1; for (usize i = 0; i < ndim; i++) n *= shape[i]; return n; }
  usize stride0() const { usize n = 1; for (usize i = 1; i < ndim; i++) n *= shape[i]; return n; }
  ViewMut index0(usize j) const { return {p + j * stride0(), shape + 1, ndim - 1}; }
  View<T> ro() const { return {p, shape, ndim}; }
};

template <class T> struct Arr {
  std::vector<usize> shape;
  std::vector<T> data;
  Arr() = default;
  Arr(std::vector<usize> s, T fill) : shape(std::move(s)), data(prod(shape), fill) {}
  usize ndim() const { return shape.size(); }
  usize len() const { return data.size(); }
  View<T> view() const { return {data.data(), shape.data(), shape.size()}; }
  ViewMut<T> view_mut() { return {data.data(), shape.data(), shape.size()}; }
  bool operator==(const Arr& o) const {
    if (shape != o.shape) return false;
    for (usize i = 0; i < data.size(); i++)
      if (!Num<T>::eq(data[i], o.data[i])) return false;
    return true;
  }
};

// Iterate all multi-indices of `shape` in row-major (logical) order.
template <class F> void for_each_index(const std::vector<usize>& shape, F&& f) {
  usize n = prod(shape);
  if (n == 0) return;
  std::vector<usize> idx(shape.size(), 0);
  for (usize c = 0; c < n; c++) {
    f(idx);
    for (usize a = shape.size(); a-- > 0;) {
      if (++idx[a] < shape[a]) break;
      idx[a] = 0;
    }
  }
}
inline usize offset_of(const std::vector<usize>& shape, const std::vector<usize>& idx) {
  usize off = 0;
  for (usize a = 0; a < shape.size(); a++) off = off * shape[a] + idx[a];
  return off;
}

// Copy of the sub-block lo[a] <= i_a < lo[a]+ext[a]  (ndarray slice_each_axis(..).to_owned()).
template <class T>
Arr<T> sub_block(const Arr<T>& a, const std::vector<usize>& lo, const std::vector<usize>& ext) {
  Arr<T> r(ext, Num<T>::zero());
  std::vector<usize> src(ext.size());
  usize c = 0;
  for_each_index(ext, [&](const std::vector<usize>& idx) {
    for (usize i = 0; i < idx.size(); i++) src[i] = idx[i] + lo[i];
    r.data[c++] = a.data[offset_of(a.shape, src)];
  });
  return r;
}
template <class T> Arr<T> slice_axis(const Arr<T>& a, usize axis, usize from, usize to) {
  std::vector<usize> lo(a.ndim(), 0), ext = a.shape;
  lo[axis] = from;
  ext[axis] = to - from;
  return sub_block(a, lo, ext);
}
// dst[leading block of src.shape] (op)= src    (slice_each_axis_mut(0..len).add_assign etc.)
template <class T, class F> void zip_leading(ViewMut<T> dst, View<T> src, F&& f) {
  std::vector<usize> dshape(dst.shape, dst.shape + dst.ndim), sshape(src.shape, src.shape + src.ndim);
  usize c = 0;
  for_each_index(sshape, [&](const std::vector<usize>& idx) {
    T& d = dst.p[offset_of(dshape, idx)];
    d = f(d, src.p[c++]);
  });
}

// multivariate_taylor.rs:958-969
inline std::optional<usize> extract_1d_len(const usize* shape, usize ndim) {
  std::optional<usize> res;
  for (usize i = 0; i < ndim; i++) {
    if (shape[i] != 1) {
      if (res.has_value()) return std::nullopt;
      res = shape[i];
    }
  }
  return res;
}

// multivariate_taylor.rs:972-982 -- each zs[k] is formed from zero by sequential `+= x*y`.
template <class T>
std::vector<T> mul_1d(const T* xs, usize xlen, const T* ys, usize ylen, usize n) {
  std::vector<T> zs(n, Num<T>::zero());
  for (usize k = 0; k < n; k++) {
    usize lo = sat_sub(k + 1, ylen);
    usize hi = std::min(k + 1, xlen);
    for (usize j = lo; j < hi; j++) zs[k] = zs[k] + xs[j] * ys[k - j];
  }
  return zs;
}

// multivariate_taylor.rs:984-1012 -- accumulating truncated N-D Cauchy product.
template <class T> void mul(View<T> xs, View<T> ys, ViewMut<T> res) {
  if (res.len() == 0) return;
  if (res.ndim == 0) {
    res.p[0] = res.p[0] + xs.p[0] * ys.p[0];
    return;
  }
  if (auto n = extract_1d_len(res.shape, res.ndim)) {
    std::vector<T> out = mul_1d(xs.p, xs.len(), ys.p, ys.len(), *n);
    for (usize i = 0; i < *n; i++) res.p[i] = res.p[i] + out[i];
    return;
  }
  usize xl = xs.shape[0], yl = ys.shape[0];
  for (usize k = 0; k < res.shape[0]; k++) {
    ViewMut<T> z = res.index0(k);
    usize lo = sat_sub(k + 1, yl);
    usize hi = std::min(k + 1, xl);
    for (usize j = lo; j < hi; j++) mul(xs.index0(j), ys.index0(k - j), z);
  }
}

// multivariate_taylor.rs:1162-1192 -- power-series quotient, recursion over all axes.
template <class T> void div(View<T> xs, View<T> ys, ViewMut<T> res) {
  if (xs.len() == 0) return;
  if (res.ndim == 0) {
    res.p[0] = xs.p[0] / ys.p[0];
    return;
  }
  std::vector<usize> cur_shape(res.shape + 1, res.shape + res.ndim);
  usize cur_len = res.stride0();
  for (usize k = 0; k < res.shape[0]; k++) {
    ViewMut<T> current = res.index0(k);
    usize lo = sat_sub(k + 1, ys.shape[0]);
    for (usize j = lo; j < k; j++) mul(res.ro().index0(j), ys.index0(k - j), current);
    for (usize i = 0; i < cur_len; i++) current.p[i] = -current.p[i];
    if (k < xs.shape[0]) {
      zip_leading(current, xs.index0(k), [](const T& d, const T& s) { return d + s; });
    }
    Arr<T> copy(cur_shape, Num<T>::zero());
    std::copy(current.p, current.p + cur_len, copy.data.begin());
    for (usize i = 0; i < cur_len; i++) current.p[i] = Num<T>::zero();
    div(copy.view(), ys.index0(0), current);
  }
}

// multivariate_taylor.rs:1271-1283
template <class T> std::vector<T> exp_1d(const T* xs, usize xlen, usize n) {
  std::vector<T> res(n, Num<T>::zero());
  res[0] = Num<T>::exp(xs[0]);
  for (usize k = 1; k < n; k++) {
    T sum = Num<T>::zero();
    usize hi = std::min(xlen, k + 1);
    for (usize j = 1; j < hi; j++) sum = sum + xs[j] * Num<T>::from_u32((uint32_t)j) * res[k - j];
    res[k] = sum / Num<T>::from_u32((uint32_t)k);
  }
  return res;
}

// multivariate_taylor.rs:1285-1317
template <class T> void exp(View<T> xs, ViewMut<T> res) {
  if (xs.len() == 0) return;
  if (res.ndim == 0) {
    res.p[0] = Num<T>::exp(xs.p[0]);
    return;
  }
  if (auto n = extract_1d_len(res.shape, res.ndim)) {
    std::vector<T> out = exp_1d(xs.p, xs.len(), *n);
    for (usize i = 0; i < *n; i++) res.p[i] = out[i];
    return;
  }
  exp(xs.index0(0), res.index0(0));
  usize cur_len = res.stride0();
  std::vector<usize> xs_sub_shape(xs.shape + 1, xs.shape + xs.ndim);
  for (usize k = 1; k < res.shape[0]; k++) {
    ViewMut<T> current = res.index0(k);
    usize hi = std::min(xs.shape[0], k + 1);
    for (usize j = 1; j < hi; j++) {
      View<T> xj = xs.index0(j);
      Arr<T> scaled(xs_sub_shape, Num<T>::zero());
      for (usize i = 0; i < scaled.data.size(); i++) scaled.data[i] = xj.p[i] * Num<T>::from_u32((uint32_t)j);
      mul(scaled.view(), res.ro().index0(k - j), current);
    }
    for (usize i = 0; i < cur_len; i++) current.p[i] = current.p[i] / Num<T>::from_u32((uint32_t)k);
  }
}

// multivariate_taylor.rs:1319-1333
template <class T> std::vector<T> log_1d(const T* xs, usize xlen, usize n) {
  std::vector<T> res(n, Num<T>::zero());
  res[0] = Num<T>::log(xs[0]);
  for (usize k = 1; k < n; k++) {
    T sum = Num<T>::zero();
    usize lo = std::max<usize>(sat_sub(k + 1, xlen), 1);
    for (usize j = lo; j < k; j++) sum = sum + xs[k - j] * res[j] * Num<T>::from_u32((uint32_t)j);
    T xk = k < xlen ? xs[k] : Num<T>::zero();
    res[k] = (xk * Num<T>::from_u32((uint32_t)k) - sum) / xs[0] / Num<T>::from_u32((uint32_t)k);
  }
  return res;
}

template <class T> struct TaylorPoly;
template <class T> TaylorPoly<T> tp_div(TaylorPoly<T> a, TaylorPoly<T> b);

// multivariate_taylor.rs:1335-1386
template <class T> void log(View<T> xs, ViewMut<T> res);

// ---------------------------------------------------------------------------------------
// TaylorPoly<T>  (multivariate_taylor.rs:13-19).  `coeffs.shape` may be smaller than
// `degrees_p1`; missing entries are known zeros.
// ---------------------------------------------------------------------------------------
template <class T> struct TaylorPoly {
  Arr<T> coeffs;
  std::vector<usize> degrees_p1;

  TaylorPoly() = default;
  TaylorPoly(Arr<T> c, std::vector<usize> d) : coeffs(std::move(c)), degrees_p1(std::move(d)) {  // :33-41
    ORC_ASSERT(coeffs.ndim() == degrees_p1.size(), "TaylorPoly::new: ndim != degrees_p1.len()");
    for (usize i = 0; i < degrees_p1.size(); i++)
      ORC_ASSERT(0 < coeffs.shape[i] && coeffs.shape[i] <= degrees_p1[i], "TaylorPoly::new: shape/degree invariant");
  }
  static TaylorPoly from_coeffs(Arr<T> c) {  // :43-46
    std::vector<usize> s = c.shape;
    return TaylorPoly(std::move(c), s);
  }
  static TaylorPoly from_scalar(T x) { return from_coeffs(Arr<T>({}, x)); }  // :626-630
  static TaylorPoly zero() { return from_scalar(Num<T>::zero()); }           // :639-641
  static TaylorPoly one() { return from_scalar(Num<T>::one()); }             // :649-651
  bool operator==(const TaylorPoly& o) const { return coeffs == o.coeffs && degrees_p1 == o.degrees_p1; }  // derive(PartialEq) :10

  usize num_vars() const { return degrees_p1.size(); }                    // :48-51
  bool is_constant() const { return coeffs.len() == 1; }                  // :68-70
  usize len_of(usize v) const { return v < degrees_p1.size() ? degrees_p1[v] : UMAX; }  // :72-79
  bool is_zero() const { return coeffs.len() == 1 && Num<T>::is_zero(coeffs.data[0]); }  // :643-645
  bool is_one() const { return coeffs.len() == 1 && Num<T>::is_one(coeffs.data[0]); }    // :653-655

  TaylorPoly extend_to_dim(usize ndim, usize degree_p1) const {  // :81-89
    ORC_ASSERT(coeffs.ndim() <= ndim, "extend_to_dim: ndim shrinks");
    TaylorPoly r = *this;
    r.coeffs.shape.resize(ndim, 1);
    r.degrees_p1.resize(ndim, degree_p1);
    return r;
  }
  // test helper `extend` :91-112 -- zero-extends the stored array to `new_size`
  TaylorPoly extend(const std::vector<usize>& new_size) const {
    Arr<T> src = coeffs;
    src.shape.resize(new_size.size(), 1);
    Arr<T> out(new_size, Num<T>::zero());
    zip_leading(out.view_mut(), src.view(), [](const T&, const T& s) { return s; });
    return TaylorPoly(std::move(out), new_size);
  }

  std::vector<usize> min_degrees_p1(const TaylorPoly& o) const {  // :114-127
    std::vector<usize> d(std::max(degrees_p1.size(), o.degrees_p1.size()), UMAX);
    for (usize v = 0; v < d.size(); v++) {
      if (v < degrees_p1.size()) d[v] = std::min(d[v], degrees_p1[v]);
      if (v < o.degrees_p1.size()) d[v] = std::min(d[v], o.degrees_p1[v]);
    }
    return d;
  }
  std::vector<usize> max_shape(const TaylorPoly& o) const {  // :129-148
    std::vector<usize> s(std::max(coeffs.ndim(), o.coeffs.ndim()), 1);
    for (usize v = 0; v < s.size(); v++) {
      if (v < coeffs.ndim()) s[v] = std::max(s[v], coeffs.shape[v]);
      if (v < o.coeffs.ndim()) s[v] = std::max(s[v], o.coeffs.shape[v]);
      if (v < degrees_p1.size()) s[v] = std::min(s[v], degrees_p1[v]);
      if (v < o.degrees_p1.size()) s[v] = std::min(s[v], o.degrees_p1[v]);
    }
    return s;
  }
  std::vector<usize> sum_shape(const TaylorPoly& o) const {  // :150-170
    std::vector<usize> s(std::max(coeffs.ndim(), o.coeffs.ndim()), 0);
    for (usize v = 0; v < s.size(); v++) {
      if (v < coeffs.ndim()) s[v] += coeffs.shape[v] - 1;
      if (v < o.coeffs.ndim()) s[v] += o.coeffs.shape[v] - 1;
      s[v] += 1;
      if (v < degrees_p1.size()) s[v] = std::min(s[v], degrees_p1[v]);
      if (v < o.degrees_p1.size()) s[v] = std::min(s[v], o.degrees_p1[v]);
    }
    return s;
  }

  TaylorPoly remove_last_variable() const {  // :172-181
    usize v = num_vars() - 1;
    Arr<T> c = coeffs;
    if (v < c.ndim()) {
      c = slice_axis(c, v, 0, 1);
      c.shape.erase(c.shape.begin() + v);
    }
    std::vector<usize> d = degrees_p1;
    d.pop_back();
    return TaylorPoly(std::move(c), std::move(d));
  }
  TaylorPoly truncate_to_degree_p1(usize degree_p1) const {  // :183-193
    std::vector<usize> d(num_vars(), degree_p1);
    TaylorPoly r = *this;
    r.truncate_degrees_p1(d);
    return r;
  }
  void truncate_degrees_p1(const std::vector<usize>& d) {  // :195-204
    for (usize v = 0; v < num_vars(); v++) {
      degrees_p1[v] = std::min(degrees_p1[v], d[v]);
      if (v < coeffs.ndim() && coeffs.shape[v] > d[v]) coeffs = slice_axis(coeffs, v, 0, d[v]);
    }
  }

  static TaylorPoly zero_with(std::vector<usize> d) {  // :208-216
    Arr<T> c(std::vector<usize>(d.size(), 1), Num<T>::zero());
    return TaylorPoly(std::move(c), std::move(d));
  }
  static TaylorPoly from_u32_with(uint32_t c, std::vector<usize> d) {  // :219-225
    return TaylorPoly(Arr<T>({}, Num<T>::from_u32(c)), std::move(d));
  }
  static TaylorPoly var_at_zero(usize v, usize len) {  // :228-237
    std::vector<usize> shape(v + 1, 1);
    shape[v] = 2;
    Arr<T> c(shape, Num<T>::zero());
    if (len > 1) c.data[1] = Num<T>::one();
    return TaylorPoly(std::move(c), std::vector<usize>(v + 1, len));
  }
  static TaylorPoly var(usize v, T x, usize len) {  // :239-248
    std::vector<usize> shape(v + 1, 1);
    shape[v] = std::min<usize>(len, 2);
    Arr<T> c(shape, Num<T>::zero());
    c.data[0] = x;
    if (len > 1) c.data[1] = Num<T>::one();
    return TaylorPoly(std::move(c), std::vector<usize>(v + 1, len));
  }
  static TaylorPoly var_with_degrees_p1(usize v, T x, std::vector<usize> d) {  // :250-259
    std::vector<usize> shape(d.size(), 1);
    shape[v] = 2;
    Arr<T> c(shape, Num<T>::zero());
    c.data[0] = x;
    // element [.., 1 (axis v), ..] of a shape that is 1 everywhere except axis v: flat index 1
    if (d[v] > 1) c.data[1] = Num<T>::one();
    return TaylorPoly(std::move(c), std::move(d));
  }

  std::optional<T> extract_constant() const {  // :262-269
    if (coeffs.len() == 1) return coeffs.data[0];
    return std::nullopt;
  }
  struct Linear { T c, m; usize v; };
  std::optional<Linear> extract_linear() const {  // :275-294
    for (usize v = 0; v < coeffs.ndim(); v++) {
      if (coeffs.shape[v] < 2) continue;
      // every entry other than [v=0,rest=0] and [v=1,rest=0] must be zero
      bool ok = true;
      usize c = 0;
      for_each_index(coeffs.shape, [&](const std::vector<usize>& idx) {
        const T& x = coeffs.data[c++];
        if (!ok) return;
        bool rest_zero = true;
        for (usize a = 0; a < idx.size(); a++)
          if (a != v && idx[a] != 0) rest_zero = false;
        bool exempt = idx[v] <= 1 && rest_zero;
        if (!exempt && !Num<T>::is_zero(x)) ok = false;
      });
      if (ok) {
        std::vector<usize> i0(coeffs.ndim(), 0), i1(coeffs.ndim(), 0);
        i1[v] = 1;
        return Linear{coeffs.data[offset_of(coeffs.shape, i0)], coeffs.data[offset_of(coeffs.shape, i1)], v};
      }
    }
    return std::nullopt;
  }
  T constant_term() const { return coeffs.data[0]; }  // :296-299

  T coefficient(const std::vector<usize>& index) const {  // :314-339
    std::vector<usize> full(coeffs.ndim(), 0);
    for (usize v = 0; v < index.size(); v++) {
      ORC_ASSERT(index[v] < len_of(v), "index out of bounds");
      if (v >= coeffs.ndim()) {
        if (index[v] != 0) return Num<T>::zero();
      } else if (index[v] >= coeffs.shape[v]) {
        return Num<T>::zero();
      } else {
        full[v] = index[v];
      }
    }
    ORC_ASSERT(index.size() >= coeffs.ndim(), "index is too short");
    return coeffs.data[offset_of(coeffs.shape, full)];
  }

  TaylorPoly coefficients_of_term(usize v, usize order) const {  // :341-358
    if (v >= coeffs.ndim()) {
      if (order == 0) return *this;
      return zero_with(degrees_p1);
    }
    if (order >= coeffs.shape[v]) return zero_with(degrees_p1);
    return TaylorPoly(slice_axis(coeffs, v, order, order + 1), degrees_p1);
  }
  TaylorPoly taylor_polynomial(usize v, usize order) const {  // :360-378
    ORC_ASSERT(v < num_vars() && order < len_of(v), "taylor_polynomial: bad v/order");
    if (v >= coeffs.ndim()) {
      if (order == 0) return *this;
      return zero_with(degrees_p1);
    }
    if (order >= coeffs.shape[v]) return *this;
    usize upper = std::min(coeffs.shape[v], order + 1);
    return TaylorPoly(slice_axis(coeffs, v, 0, upper), degrees_p1);
  }
  TaylorPoly taylor_polynomial_terms(usize v, const std::vector<usize>& orders) const {  // :380-404
    usize max_order_p1 = 1;
    if (!orders.empty()) max_order_p1 = *std::max_element(orders.begin(), orders.end()) + 1;
    if (v >= coeffs.ndim()) {
      if (std::find(orders.begin(), orders.end(), (usize)0) != orders.end()) return *this;
      return zero_with(degrees_p1);
    }
    usize upper = std::min(coeffs.shape[v], max_order_p1);
    Arr<T> result = slice_axis(coeffs, v, 0, upper);
    std::vector<char> keep(max_order_p1, 0);
    for (usize o : orders) keep[o] = 1;
    usize c = 0;
    for_each_index(result.shape, [&](const std::vector<usize>& idx) {
      if (!keep[idx[v]]) result.data[c] = Num<T>::zero();
      c++;
    });
    return TaylorPoly(std::move(result), degrees_p1);
  }

  TaylorPoly exp() const {  // :406-417
    std::vector<usize> rs = degrees_p1;
    for (usize i = 0; i < rs.size(); i++)
      if (coeffs.shape[i] == 1) rs[i] = 1;
    Arr<T> result(rs, Num<T>::zero());
    orc::exp(coeffs.view(), result.view_mut());
    return TaylorPoly(std::move(result), degrees_p1);
  }
  TaylorPoly log() const {  // :419-430
    std::vector<usize> rs = degrees_p1;
    for (usize i = 0; i < rs.size(); i++)
      if (coeffs.shape[i] == 1) rs[i] = 1;
    Arr<T> result(rs, Num<T>::zero());
    orc::log(coeffs.view(), result.view_mut());
    return TaylorPoly(std::move(result), degrees_p1);
  }
  TaylorPoly pow(uint32_t e) const;  // :433-451

  TaylorPoly derivative(usize v, usize n) const {  // :457-481
    ORC_ASSERT(v < num_vars() && n < len_of(v), "derivative: bad v/n");
    if (v >= coeffs.ndim()) {
      if (n == 0) return *this;
      return zero_with(degrees_p1);
    }
    std::vector<usize> d = degrees_p1;
    d[v] = sat_sub(d[v], n);
    if (n >= coeffs.shape[v]) return zero_with(d);
    Arr<T> result = slice_axis(coeffs, v, n, coeffs.shape[v]);
    T ff = Num<T>::one();
    for (usize i = 1; i <= n; i++) ff = ff * Num<T>::from_u32((uint32_t)i);
    std::vector<T> factor(result.shape[v], Num<T>::zero());
    for (usize k = 0; k < result.shape[v]; k++) {
      factor[k] = ff;
      ff = ff * (Num<T>::from_u32((uint32_t)(n + k + 1)) / Num<T>::from_u32((uint32_t)(k + 1)));
    }
    usize c = 0;
    for_each_index(result.shape, [&](const std::vector<usize>& idx) {
      result.data[c] = result.data[c] * factor[idx[v]];
      c++;
    });
    return TaylorPoly(std::move(result), std::move(d));
  }
  TaylorPoly taylor_expansion_of_coeff(usize v, usize n) const {  // :484-509
    ORC_ASSERT(v < num_vars() && n < len_of(v), "taylor_expansion_of_coeff: bad v/n");
    if (v >= coeffs.ndim()) {
      if (n == 0) return *this;
      return zero_with(degrees_p1);
    }
    std::vector<usize> d = degrees_p1;
    d[v] = sat_sub(d[v], n);
    if (n >= coeffs.shape[v]) return zero_with(d);
    Arr<T> result = slice_axis(coeffs, v, n, coeffs.shape[v]);
    std::vector<T> factor(result.shape[v], Num<T>::one());
    T f = Num<T>::one();
    for (usize k = 1; k < result.shape[v]; k++) {
      f = f * (Num<T>::from_u32((uint32_t)(n + k)) / Num<T>::from_u32((uint32_t)k));
      factor[k] = f;
    }
    usize c = 0;
    for_each_index(result.shape, [&](const std::vector<usize>& idx) {
      if (idx[v] >= 1) result.data[c] = result.data[c] * factor[idx[v]];  // slice 0 untouched (.skip(1))
      c++;
    });
    return TaylorPoly(std::move(result), std::move(d));
  }
  TaylorPoly shift_down(usize v, usize n) const;  // :514-536
  TaylorPoly subst_var(usize v, const TaylorPoly& subst) const;  // :540-580

  T evaluate_all_one() const {  // :583-586
    T acc = Num<T>::zero();
    for (const T& x : coeffs.data) acc = acc + x;
    return acc;
  }
  TaylorPoly mul_var(T m, usize v, std::vector<usize> shape, std::vector<usize> d) const {  // :589-608
    usize upper = std::min(shape[v] - 1, coeffs.shape[v]);
    Arr<T> self = slice_axis(coeffs, v, 0, upper);
    for (T& x : self.data) x = x * m;
    Arr<T> result(shape, Num<T>::zero());
    // result[.., 1..=upper (axis v), ..] = self[each axis clipped to shape]
    std::vector<usize> ext(self.ndim());
    for (usize a = 0; a < ext.size(); a++) ext[a] = std::min(self.shape[a], shape[a]);
    std::vector<usize> dst(ext.size());
    for_each_index(ext, [&](const std::vector<usize>& idx) {
      dst = idx;
      dst[v] += 1;
      result.data[offset_of(result.shape, dst)] = self.data[offset_of(self.shape, idx)];
    });
    return TaylorPoly(std::move(result), std::move(d));
  }
  TaylorPoly mul_linear(T c, T m, usize v, std::vector<usize> shape, std::vector<usize> d) const;  // :611-623
};

// ndarray 0.15.6 numeric_util::unrolled_fold (third-party, restated from its published source):
// 8 partial accumulators over chunks of 8, combined (p0+p4),(p1+p5),(p2+p6),(p3+p7) into acc in
// that order, then the <8 leftover elements sequentially.
template <class T> T unrolled_sum(const T* xs, usize n) {
  T acc = Num<T>::zero();
  T p[8];
  for (auto& q : p) q = Num<T>::zero();
  while (n >= 8) {
    for (int i = 0; i < 8; i++) p[i] = p[i] + xs[i];
    xs += 8;
    n -= 8;
  }
  acc = acc + (p[0] + p[4]);
  acc = acc + (p[1] + p[5]);
  acc = acc + (p[2] + p[6]);
  acc = acc + (p[3] + p[7]);
  for (usize i = 0; i < n; i++) acc = acc + xs[i];
  return acc;
}
// ndarray 0.15.6 ArrayBase::sum_axis: if `axis` is the minimum-stride axis, each lane is summed
// with unrolled_fold; otherwise `res = zeros; for subview in axis_iter(axis): res = res + subview`.
// Dimension::min_stride_axis scans the axes in REVERSE order and keeps the first minimum of
// |stride| (unit axes are not skipped), so for a standard row-major array it is always the LAST
// axis (stride 1).
template <class T> Arr<T> sum_axis(const Arr<T>& a, usize axis) {
  std::vector<usize> rshape = a.shape;
  rshape.erase(rshape.begin() + axis);
  Arr<T> res(rshape, Num<T>::zero());
  usize outer = 1, inner = 1, len = a.shape[axis];
  for (usize i = 0; i < axis; i++) outer *= a.shape[i];
  for (usize i = axis + 1; i < a.ndim(); i++) inner *= a.shape[i];
  if (axis == a.ndim() - 1) {
    for (usize o = 0; o < outer; o++) res.data[o] = unrolled_sum(a.data.data() + o * len, len);
  } else {
    for (usize l = 0; l < len; l++)
      for (usize o = 0; o < outer; o++)
        for (usize in = 0; in < inner; in++)
          res.data[o * inner + in] = res.data[o * inner + in] + a.data[(o * len + l) * inner + in];
  }
  return res;
}

template <class T> TaylorPoly<T> TaylorPoly<T>::shift_down(usize v, usize n) const {  // :514-536
  ORC_ASSERT(v < num_vars() && n < len_of(v), "shift_down: bad v/n");
  if (v >= coeffs.ndim()) return *this;
  std::vector<usize> d = degrees_p1;
  d[v] = sat_sub(d[v], n);
  Arr<T> result;
  if (coeffs.shape[v] <= n + 1) {
    result = sum_axis(coeffs, v);
    result.shape.insert(result.shape.begin() + v, 1);
  } else {
    result = slice_axis(coeffs, v, n, coeffs.shape[v]);
    Arr<T> head = sum_axis(slice_axis(coeffs, v, 0, n), v);
    // result.index_axis_mut(v, 0) += head
    std::vector<usize> hshape = head.shape;
    usize c = 0;
    std::vector<usize> dst(result.ndim());
    for_each_index(hshape, [&](const std::vector<usize>& idx) {
      usize b = 0;
      for (usize a = 0; a < result.ndim(); a++) dst[a] = (a == v) ? 0 : idx[b++];
      T& r = result.data[offset_of(result.shape, dst)];
      r = r + head.data[c++];
    });
  }
  return TaylorPoly(std::move(result), std::move(d));
}

// multivariate_taylor.rs:832-852
template <class T> void broadcast(TaylorPoly<T>& xs, TaylorPoly<T>& ys) {
  if (xs.degrees_p1.size() < ys.degrees_p1.size())
    xs.degrees_p1.insert(xs.degrees_p1.end(), ys.degrees_p1.begin() + xs.degrees_p1.size(), ys.degrees_p1.end());
  else if (ys.degrees_p1.size() < xs.degrees_p1.size())
    ys.degrees_p1.insert(ys.degrees_p1.end(), xs.degrees_p1.begin() + ys.degrees_p1.size(), xs.degrees_p1.end());
  if (xs.coeffs.ndim() < ys.coeffs.ndim()) xs.coeffs.shape.resize(ys.coeffs.ndim(), 1);
  if (ys.coeffs.ndim() < xs.coeffs.ndim()) ys.coeffs.shape.resize(xs.coeffs.ndim(), 1);
}

template <class T> TaylorPoly<T> tp_neg(TaylorPoly<T> a) {  // :902-909
  for (T& x : a.coeffs.data) x = -x;
  return a;
}

template <class T> TaylorPoly<T> tp_add(TaylorPoly<T> self, TaylorPoly<T> other) {  // :854-882
  std::vector<usize> rd = self.min_degrees_p1(other);
  broadcast(self, other);
  self.truncate_degrees_p1(rd);
  other.truncate_degrees_p1(rd);
  if (other.coeffs.len() == 1) {
    self.coeffs.data[0] = self.coeffs.data[0] + other.coeffs.data[0];
    return TaylorPoly<T>(std::move(self.coeffs), rd);
  }
  if (self.coeffs.len() == 1) {
    other.coeffs.data[0] = other.coeffs.data[0] + self.coeffs.data[0];
    return TaylorPoly<T>(std::move(other.coeffs), rd);
  }
  std::vector<usize> shape = self.max_shape(other);
  self.truncate_degrees_p1(shape);
  other.truncate_degrees_p1(shape);
  Arr<T> result(shape, Num<T>::zero());
  zip_leading(result.view_mut(), self.coeffs.view(), [](const T& d, const T& s) { return d + s; });
  zip_leading(result.view_mut(), other.coeffs.view(), [](const T& d, const T& s) { return d + s; });
  return TaylorPoly<T>(std::move(result), rd);
}

template <class T> TaylorPoly<T> tp_sub(TaylorPoly<T> self, TaylorPoly<T> other) {  // :911-937
  std::vector<usize> rd = self.min_degrees_p1(other);
  broadcast(self, other);
  self.truncate_degrees_p1(rd);
  other.truncate_degrees_p1(rd);
  if (other.coeffs.len() == 1) {
    self.coeffs.data[0] = self.coeffs.data[0] - other.coeffs.data[0];
    return TaylorPoly<T>(std::move(self.coeffs), rd);
  }
  if (self.coeffs.len() == 1) {
    other.coeffs.data[0] = other.coeffs.data[0] - self.coeffs.data[0];
    for (T& x : other.coeffs.data) x = -x;
    return TaylorPoly<T>(std::move(other.coeffs), rd);
  }
  std::vector<usize> shape = self.max_shape(other);
  Arr<T> result(shape, Num<T>::zero());
  zip_leading(result.view_mut(), self.coeffs.view(), [](const T& d, const T& s) { return d + s; });
  zip_leading(result.view_mut(), other.coeffs.view(), [](const T& d, const T& s) { return d - s; });
  return TaylorPoly<T>(std::move(result), rd);
}

template <class T> TaylorPoly<T> tp_mul(TaylorPoly<T> self, TaylorPoly<T> other) {  // :1014-1072
  std::vector<usize> d = self.min_degrees_p1(other);
  if (self.is_zero() || other.is_zero()) return TaylorPoly<T>::zero_with(d);
  broadcast(self, other);
  std::vector<usize> shape = self.sum_shape(other);
  self.truncate_degrees_p1(d);
  other.truncate_degrees_p1(d);
  if (self.is_one()) return other;
  if (other.is_one()) return self;
  if (auto c = self.extract_constant()) {
    for (T& x : other.coeffs.data) x = *c * x;
    return other;
  }
  if (auto c = other.extract_constant()) {
    for (T& x : self.coeffs.data) x = *c * x;
    return self;
  }
  if (auto lin = self.extract_linear()) {
    std::vector<usize> s = other.coeffs.shape;
    s[lin->v] = std::min(d[lin->v], s[lin->v] + 1);
    return other.mul_linear(lin->c, lin->m, lin->v, s, d);
  }
  if (auto lin = other.extract_linear()) {
    std::vector<usize> s = self.coeffs.shape;
    s[lin->v] = std::min(d[lin->v], s[lin->v] + 1);
    return self.mul_linear(lin->c, lin->m, lin->v, s, d);
  }
  Arr<T> result(shape, Num<T>::zero());
  mul(self.coeffs.view(), other.coeffs.view(), result.view_mut());
  return TaylorPoly<T>(std::move(result), d);
}

template <class T>
TaylorPoly<T> TaylorPoly<T>::mul_linear(T c, T m, usize v, std::vector<usize> shape, std::vector<usize> d) const {  // :611-623
  if (Num<T>::is_zero(c)) return mul_var(m, v, shape, d);
  return tp_add(mul_var(m, v, shape, d), tp_mul(*this, TaylorPoly<T>::from_scalar(c)));
}

template <class T> TaylorPoly<T> tp_div(TaylorPoly<T> self, TaylorPoly<T> other) {  // :1194-1231
  broadcast(self, other);
  std::vector<usize> d = self.min_degrees_p1(other);
  self.truncate_degrees_p1(d);
  other.truncate_degrees_p1(d);
  if (other.is_one()) return self;
  if (auto c = other.extract_constant()) {
    for (T& x : self.coeffs.data) x = x / *c;
    return self;
  }
  std::vector<usize> rs = d;
  for (usize i = 0; i < rs.size(); i++)
    if (other.coeffs.shape[i] == 1) rs[i] = self.coeffs.shape[i];
  Arr<T> result(rs, Num<T>::zero());
  div(self.coeffs.view(), other.coeffs.view(), result.view_mut());
  return TaylorPoly<T>(std::move(result), d);
}

template <class T> void log(View<T> xs, ViewMut<T> res) {  // :1335-1386
  if (xs.len() == 0) return;
  if (res.ndim == 0) {
    res.p[0] = Num<T>::log(xs.p[0]);
    return;
  }
  if (extract_1d_len(xs.shape, xs.ndim).has_value()) {
    auto n = extract_1d_len(res.shape, res.ndim);
    ORC_ASSERT(n.has_value(), "log: unwrap on None (result not 1-d)");
    std::vector<T> out = log_1d(xs.p, xs.len(), *n);
    for (usize i = 0; i < *n; i++) res.p[i] = out[i];
    return;
  }
  log(xs.index0(0), res.index0(0));
  usize cur_len = res.stride0();
  std::vector<usize> cur_shape(res.shape + 1, res.shape + res.ndim);
  std::vector<usize> xs_sub_shape(xs.shape + 1, xs.shape + xs.ndim);
  for (usize k = 1; k < res.shape[0]; k++) {
    ViewMut<T> current = res.index0(k);
    usize lo = std::max<usize>(sat_sub(k + 1, xs.shape[0]), 1);
    for (usize j = lo; j < k; j++) {
      View<T> rj = res.ro().index0(j);
      Arr<T> scaled(cur_shape, Num<T>::zero());
      for (usize i = 0; i < cur_len; i++) scaled.data[i] = rj.p[i] * Num<T>::from_u32((uint32_t)j);
      mul(xs.index0(k - j), scaled.view(), current);
    }
    for (usize i = 0; i < cur_len; i++) current.p[i] = -current.p[i];
    if (k < xs.shape[0]) {
      View<T> xk = xs.index0(k);
      Arr<T> scaled(xs_sub_shape, Num<T>::zero());
      for (usize i = 0; i < scaled.data.size(); i++) scaled.data[i] = Num<T>::from_u32((uint32_t)k) * xk.p[i];
      zip_leading(current, scaled.view(), [](const T& d, const T& s) { return d + s; });
    }
    Arr<T> num(cur_shape, Num<T>::zero());
    std::copy(current.p, current.p + cur_len, num.data.begin());
    Arr<T> den(xs_sub_shape, Num<T>::zero());
    View<T> x0 = xs.index0(0);
    std::copy(x0.p, x0.p + den.data.size(), den.data.begin());
    TaylorPoly<T> q = tp_div(TaylorPoly<T>(std::move(num), cur_shape), TaylorPoly<T>(std::move(den), cur_shape));
    ORC_ASSERT(q.coeffs.shape == cur_shape, "log: assign shape mismatch");
    for (usize i = 0; i < cur_len; i++) current.p[i] = q.coeffs.data[i] / Num<T>::from_u32((uint32_t)k);
  }
}

template <class T> TaylorPoly<T> TaylorPoly<T>::pow(uint32_t e) const {  // :433-451
  if (e == 0) return one();
  if (e == 1) return *this;
  TaylorPoly res = one();
  TaylorPoly base = *this;
  while (e > 0) {
    if (e & 1) res = tp_mul(res, base);
    base = tp_mul(base, base);  // note: also after the last bit (:447)
    e >>= 1;
  }
  return res;
}

template <class T> TaylorPoly<T> TaylorPoly<T>::subst_var(usize v, const TaylorPoly& subst) const {  // :540-580
  if (v >= coeffs.ndim()) return *this;
  std::vector<usize> d = min_degrees_p1(subst);
  if (subst.is_zero()) return TaylorPoly(slice_axis(coeffs, v, 0, 1), d);
  if (auto lin = subst.extract_linear()) {
    if (v == lin->v && Num<T>::is_zero(lin->c)) {
      std::vector<usize> lo(coeffs.ndim(), 0), ext(coeffs.ndim());
      for (usize a = 0; a < ext.size(); a++) ext[a] = std::min(coeffs.shape[a], d[a]);
      Arr<T> result = sub_block(coeffs, lo, ext);
      std::vector<T> factor(result.shape[v], Num<T>::one());
      T f = Num<T>::one();
      for (usize i = 0; i < factor.size(); i++) {
        factor[i] = f;
        f = f * lin->m;
      }
      usize c = 0;
      for_each_index(result.shape, [&](const std::vector<usize>& idx) {
        result.data[c] = result.data[c] * factor[idx[v]];
        c++;
      });
      return TaylorPoly(std::move(result), d);
    }
  }
  TaylorPoly res = zero_with(d);
  Arr<T> cs = coeffs;
  cs.shape.resize(std::max(cs.ndim(), d.size()), 1);
  for (usize i = cs.shape[v]; i-- > 0;) {
    std::vector<usize> lo(cs.ndim(), 0), ext(cs.ndim());
    for (usize a = 0; a < ext.size(); a++) ext[a] = std::min(cs.shape[a], d[a]);
    lo[v] = i;
    ext[v] = std::min<usize>(1, d[v]);
    ORC_ASSERT(ext[v] == 1, "subst_var: zero degree along substituted axis");
    res = tp_add(tp_mul(res, subst), TaylorPoly(sub_block(cs, lo, ext), d));
  }
  return res;
}

// ---------------------------------------------------------------------------------------
// TaylorExpansion<T>  (univariate_taylor.rs:9-13)
// ---------------------------------------------------------------------------------------
template <class T> struct TaylorExpansion {
  bool is_const;          // Constant(T) vs Polynomial{coeffs}
  T c;                    // valid when is_const
  std::vector<T> coeffs;  // valid when !is_const

  static TaylorExpansion constant(T x) { return {true, x, {}}; }
  static TaylorExpansion polynomial(std::vector<T> v) { return {false, Num<T>::zero(), std::move(v)}; }
  static TaylorExpansion zero() { return constant(Num<T>::zero()); }  // :238-241
  static TaylorExpansion one() { return constant(Num<T>::one()); }    // :250-253
  bool operator==(const TaylorExpansion& o) const {
    if (is_const != o.is_const) return false;
    if (is_const) return Num<T>::eq(c, o.c);
    if (coeffs.size() != o.coeffs.size()) return false;
    for (usize i = 0; i < coeffs.size(); i++)
      if (!Num<T>::eq(coeffs[i], o.coeffs[i])) return false;
    return true;
  }
  static TaylorExpansion var(T x, usize order) {  // :16-23
    std::vector<T> v(order + 1, Num<T>::zero());
    if (1 < v.size()) v[1] = Num<T>::one();
    v[0] = x;
    return polynomial(std::move(v));
  }
  T coeff(usize order) const {  // :25-36
    if (!is_const) {
      ORC_ASSERT(order < coeffs.size(), "coeff: index out of bounds");
      return coeffs[order];
    }
    return order == 0 ? c : Num<T>::zero();
  }
  usize order() const { return is_const ? UMAX : coeffs.size(); }  // :38-43
  T derivative(usize order) const {                                // :45-60
    if (!is_const) {
      ORC_ASSERT(order < coeffs.size(), "derivative: index out of bounds");
      T f = Num<T>::one();
      for (usize i = 1; i <= order; i++) f = f * Num<T>::from_u32((uint32_t)i);
      return f * coeffs[order];
    }
    return order == 0 ? c : Num<T>::zero();
  }
  TaylorExpansion taylor_expansion_of_coeff(usize n) const {  // :69-89
    if (is_const) {
      if (n == 0) return constant(Num<T>::exp(c));  // sic: the reference applies exp here (:73)
      return zero();
    }
    ORC_ASSERT(n <= coeffs.size(), "taylor_expansion_of_coeff: slice start out of range");
    std::vector<T> res(coeffs.begin() + n, coeffs.end());
    T f = Num<T>::one();
    for (usize k = 1; k < res.size(); k++) {
      f = f * (Num<T>::from_u32((uint32_t)(n + k)) / Num<T>::from_u32((uint32_t)k));
      res[k] = res[k] * f;
    }
    return polynomial(std::move(res));
  }
  TaylorExpansion exp() const {  // :151-168  (note the association res*coeff*j, unlike exp_1d)
    if (is_const) return constant(Num<T>::exp(c));
    usize n = coeffs.size();
    std::vector<T> res(n, Num<T>::zero());
    res[0] = Num<T>::exp(coeffs[0]);
    for (usize k = 1; k < n; k++) {
      T sum = Num<T>::zero();
      for (usize j = 1; j <= k; j++) sum = sum + res[k - j] * coeffs[j] * Num<T>::from_u32((uint32_t)j);
      res[k] = sum / Num<T>::from_u32((uint32_t)k);
    }
    return polynomial(std::move(res));
  }
  TaylorExpansion log() const {  // :170-189
    if (is_const) return constant(Num<T>::log(c));
    usize n = coeffs.size();
    std::vector<T> res(n, Num<T>::zero());
    res[0] = Num<T>::log(coeffs[0]);
    for (usize k = 1; k < n; k++) {
      T sum = Num<T>::zero();
      for (usize j = 1; j < k; j++) sum = sum + coeffs[k - j] * res[j] * Num<T>::from_u32((uint32_t)j);
      res[k] = (coeffs[k] * Num<T>::from_u32((uint32_t)k) - sum) / coeffs[0] / Num<T>::from_u32((uint32_t)k);
    }
    return polynomial(std::move(res));
  }
};

template <class T> TaylorExpansion<T> te_add(TaylorExpansion<T> self, TaylorExpansion<T> rhs) {  // :277-306
  if (rhs.is_const) {
    if (self.is_const) self.c = self.c + rhs.c;
    else self.coeffs[0] = self.coeffs[0] + rhs.c;
    return self;
  }
  std::vector<T>& ws = rhs.coeffs;
  if (self.is_const) {
    ws[0] = ws[0] + self.c;
    return TaylorExpansion<T>::polynomial(std::move(ws));
  }
  usize order = std::min(self.coeffs.size(), ws.size());
  for (usize i = 0; i < order; i++) ws[i] = ws[i] + self.coeffs[i];
  ws.resize(order, Num<T>::zero());
  return TaylorExpansion<T>::polynomial(std::move(ws));
}
template <class T> TaylorExpansion<T> te_neg(TaylorExpansion<T> a) {  // :308-319
  if (a.is_const) a.c = -a.c;
  else for (T& x : a.coeffs) x = -x;
  return a;
}
template <class T> TaylorExpansion<T> te_sub(TaylorExpansion<T> self, TaylorExpansion<T> rhs) {  // :330-362
  if (rhs.is_const) {
    if (self.is_const) self.c = self.c - rhs.c;
    else self.coeffs[0] = self.coeffs[0] - rhs.c;
    return self;
  }
  std::vector<T>& ws = rhs.coeffs;
  if (self.is_const) {
    for (T& w : ws) w = -w;
    ws[0] = ws[0] + self.c;
    return TaylorExpansion<T>::polynomial(std::move(ws));
  }
  usize order = std::min(self.coeffs.size(), ws.size());
  for (usize i = 0; i < order; i++) ws[i] = self.coeffs[i] - ws[i];
  ws.resize(order, Num<T>::zero());
  return TaylorExpansion<T>::polynomial(std::move(ws));
}
template <class T> TaylorExpansion<T> te_mul(TaylorExpansion<T> a, TaylorExpansion<T> b) {  // :364-389
  if (a.is_const && b.is_const) return TaylorExpansion<T>::constant(a.c * b.c);
  if (a.is_const || b.is_const) {
    T c = a.is_const ? a.c : b.c;
    std::vector<T> cs = a.is_const ? std::move(b.coeffs) : std::move(a.coeffs);
    for (T& x : cs) x = x * c;
    return TaylorExpansion<T>::polynomial(std::move(cs));
  }
  usize order = std::min(a.coeffs.size(), b.coeffs.size());
  std::vector<T> r(order, Num<T>::zero());
  for (usize k = 0; k < order; k++) {
    T sum = Num<T>::zero();
    for (usize j = 0; j <= k; j++) sum = sum + a.coeffs[j] * b.coeffs[k - j];
    r[k] = sum;
  }
  return TaylorExpansion<T>::polynomial(std::move(r));
}
template <class T> TaylorExpansion<T> te_div(TaylorExpansion<T> a, TaylorExpansion<T> b) {  // :397-439
  if (a.is_const && b.is_const) return TaylorExpansion<T>::constant(a.c / b.c);
  if (!a.is_const && b.is_const) {
    for (T& x : a.coeffs) x = x / b.c;
    return a;
  }
  const std::vector<T>& ws = b.coeffs;
  if (a.is_const) {
    usize order = ws.size();
    std::vector<T> r(order, Num<T>::zero());
    T scale = Num<T>::one() / ws[0];
    r[0] = a.c * scale;
    for (usize k = 1; k < order; k++) {
      T sum = Num<T>::zero();
      for (usize i = 0; i < k; i++) sum = sum - r[i] * ws[k - i];
      r[k] = scale * sum;
    }
    return TaylorExpansion<T>::polynomial(std::move(r));
  }
  usize order = std::min(a.coeffs.size(), ws.size());
  std::vector<T> r(order, Num<T>::zero());
  T scale = Num<T>::one() / ws[0];
  r[0] = scale * a.coeffs[0];
  for (usize k = 1; k < order; k++) {
    T sum = a.coeffs[k];
    for (usize i = 0; i < k; i++) sum = sum - r[i] * ws[k - i];
    r[k] = scale * sum;
  }
  return TaylorExpansion<T>::polynomial(std::move(r));
}
template <class T> TaylorExpansion<T> te_pow(const TaylorExpansion<T>& x, uint32_t e) {  // :192-203
  TaylorExpansion<T> res = TaylorExpansion<T>::one();
  TaylorExpansion<T> base = x;
  while (e > 0) {
    if (e & 1) res = te_mul(res, base);
    base = te_mul(base, base);
    e >>= 1;
  }
  return res;
}
template <class T> TaylorExpansion<T> te_subst(const TaylorExpansion<T>& self, const TaylorExpansion<T>& subst) {  // :93-115
  if (self.is_const) return self;
  if (!subst.is_const) ORC_ASSERT(subst.coeffs.size() == self.coeffs.size(), "Substitution must have the same order");
  TaylorExpansion<T> res = TaylorExpansion<T>::zero();
  for (usize i = self.coeffs.size(); i-- > 0;)
    res = te_add(te_mul(res, subst), TaylorExpansion<T>::constant(self.coeffs[i]));
  return res;
}

// Algorithmic MAC count of the general product (trip counts of :975-977 / :1002-1004).
inline double mul_macs(const std::vector<usize>& xs, const std::vector<usize>& ys, const std::vector<usize>& rs) {
  double total = 1.0;
  for (usize a = 0; a < rs.size(); a++) {
    double s = 0;
    for (usize k = 0; k < rs[a]; k++) {
      usize lo = sat_sub(k + 1, ys[a]), hi = std::min(k + 1, xs[a]);
      if (hi > lo) s += (double)(hi - lo);
    }
    total *= s;
  }
  return total;
}

}  // namespace orc
