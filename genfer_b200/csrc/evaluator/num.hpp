// Host evaluator -- scalars of the generating function.
// F64 helper semantics (src/number/f64.rs: powi :64-66, min / max :68-84, next_up / next_down :127-171), Interval<F64>
// (src/interval.rs: every operation widens by one ulp each side, :29-31), and `Num`: a constant of the GenFun as BOTH instantiations
// of the reference hold it -- T = F64 (`v`) and T = Interval<F64> (`iv`, the --bounds mode, whose constants come from
// Number::from_ratio's default implementation, number/number.rs:26-33, not from one IEEE division).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace gfe {

// ---- F64 scalar helpers ---------------------------------------------------------------------------
inline double next_up(double x) {
  uint64_t bits;
  std::memcpy(&bits, &x, 8);
  if (std::isnan(x) || bits == 0x7ff0000000000000ULL) return x;
  uint64_t abs = bits & 0x7fffffffffffffffULL, next;
  if (abs == 0) next = 1;
  else if (bits == abs) next = bits + 1;
  else next = bits - 1;
  double r;
  std::memcpy(&r, &next, 8);
  return r;
}
inline double next_down(double x) {
  uint64_t bits;
  std::memcpy(&bits, &x, 8);
  if (std::isnan(x) || bits == 0xfff0000000000000ULL) return x;
  uint64_t abs = bits & 0x7fffffffffffffffULL, next;
  if (abs == 0) next = 0x8000000000000001ULL;
  else if (bits == abs) next = bits - 1;
  else next = bits + 1;
  double r;
  std::memcpy(&r, &next, 8);
  return r;
}
inline double powi(double a, uint32_t b) {  // f64::powi = compiler-rt __powidf2 (binary exponentiation)
  double r = 1.0;
  while (true) {
    if (b & 1) r *= a;
    b /= 2;
    if (b == 0) break;
    a *= a;
  }
  return r;
}
inline double f64_min(double a, double b) { return a < b ? a : b; }   // number/f64.rs:69-75
inline double f64_max(double a, double b) { return a > b ? a : b; }   // :77-84

// ---- Interval<F64> (src/interval.rs) -------------------------------------------------------------------
struct Iv {
  double lo = 0, hi = 0;
  static Iv exact(double l, double h) { return {l, h}; }
  static Iv precisely(double x) { return {x, x}; }
  static Iv widen(double l, double h) { return {next_down(l), next_up(h)}; }
  static Iv zero() { return {0.0, 0.0}; }
  static Iv one() { return {1.0, 1.0}; }
  bool is_zero() const { return lo == 0.0 && hi == 0.0; }
  bool is_one() const { return lo == 1.0 && hi == 1.0; }
  bool is_finite() const { return std::isfinite(lo) && std::isfinite(hi); }
  bool is_nan() const { return std::isnan(lo) || std::isnan(hi); }
  bool contains(double x) const { return lo <= x && x <= hi; }
  Iv unite(double x) const { return {f64_min(lo, x), f64_max(hi, x)}; }
  bool is_point() const { return lo == hi; }
  double center() const { return (lo + hi) / 2.0; }
  Iv ensure_lower_bound(double nl) const { return lo < nl ? Iv{nl, hi} : *this; }
  Iv ensure_upper_bound(double nh) const { return hi > nh ? Iv{lo, nh} : *this; }
  Iv neg() const { return {-hi, -lo}; }
  Iv add(const Iv& r) const {
    if (is_zero()) return r;
    if (r.is_zero()) return *this;
    return widen(lo + r.lo, hi + r.hi);
  }
  Iv sub(const Iv& r) const { return add(r.neg()); }
  Iv mul(const Iv& r) const {
    if ((is_zero() && r.is_finite()) || (is_finite() && r.is_zero())) return zero();
    if (is_one()) return r;
    if (r.is_one()) return *this;
    if (neg().is_one()) return r.neg();
    if (r.neg().is_one()) return neg();
    double a = lo * r.lo, b = lo * r.hi, c = hi * r.lo, d = hi * r.hi;
    return widen(f64_min(f64_min(f64_min(a, b), c), d), f64_max(f64_max(f64_max(a, b), c), d));
  }
  Iv div(const Iv& r) const {
    if (is_nan() || r.is_nan()) return {NAN, NAN};
    if (is_zero() && !r.is_zero()) return *this;
    if (r.is_one()) return *this;
    double l = INFINITY, h = -INFINITY;
    if (r.contains(0.0)) {
      if (0.0 <= lo) h = INFINITY; else l = -INFINITY;
      if (hi <= 0.0) l = -INFINITY; else h = INFINITY;
    }
    double a = lo / r.lo, b = lo / r.hi, c = hi / r.lo, d = hi / r.hi;
    l = f64_min(f64_min(f64_min(f64_min(l, a), b), c), d);
    h = f64_max(f64_max(f64_max(f64_max(h, a), b), c), d);
    return widen(l, h);
  }
  Iv pow(uint32_t e) const {
    Iv r = widen(powi(lo, e), powi(hi, e));
    return contains(0.0) ? r.unite(0.0) : r;
  }
  Iv exp() const {   // :264-269
    if (is_zero()) return one();
    return widen(std::exp(lo), std::exp(hi));
  }
  Iv sqrt() const {
    double l = lo < 0.0 ? 0.0 : std::sqrt(lo);
    return widen(l, std::sqrt(hi));
  }
  // PartialOrd (:236-248)
  bool lt(const Iv& o) const { return !(lo == o.lo && hi == o.hi) && hi <= o.lo; }
  bool gt(const Iv& o) const { return !(lo == o.lo && hi == o.hi) && !(hi <= o.lo) && lo >= o.hi; }
};

// A constant of the generating function under T = F64 (v) and under T = Interval<F64> (iv); arithmetic acts on both.
struct Num {
  double v = 0.0;
  Iv iv;
  Num() = default;
  Num(double x) : v(x), iv{x, x} {}   // T::from(u32), T::zero(), T::one(): exact in both
  Num(double x, Iv i) : v(x), iv(i) {}
  // F64: one IEEE division (number/f64.rs:49-51).  Interval<F64>: the trait's default (number/number.rs:26-33):
  // (lo32 + hi32 * 2^32) / (lo32 + hi32 * 2^32) in interval arithmetic, i.e. widened even when the quotient is exact.
  static Num from_ratio(uint64_t numerator, uint64_t denominator) {
    const Iv two_to_32 = Iv::precisely(4294967295.0).add(Iv::one());
    const Iv n = Iv::precisely((double)(uint32_t)numerator).add(Iv::precisely((double)(uint32_t)(numerator >> 32)).mul(two_to_32));
    const Iv d = Iv::precisely((double)(uint32_t)denominator).add(Iv::precisely((double)(uint32_t)(denominator >> 32)).mul(two_to_32));
    return Num((double)numerator / (double)denominator, n.div(d));
  }
};
inline Num operator+(const Num& a, const Num& b) { return Num(a.v + b.v, a.iv.add(b.iv)); }
inline Num operator-(const Num& a, const Num& b) { return Num(a.v - b.v, a.iv.sub(b.iv)); }
inline Num operator*(const Num& a, const Num& b) { return Num(a.v * b.v, a.iv.mul(b.iv)); }
inline Num operator/(const Num& a, const Num& b) { return Num(a.v / b.v, a.iv.div(b.iv)); }
inline Num operator-(const Num& a) { return Num(-a.v, a.iv.neg()); }
inline Num exp(const Num& a) { return Num(std::exp(a.v), a.iv.exp()); }

}  // namespace gfe
