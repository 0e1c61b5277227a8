// Interval<F64> scalar of the reference's --bounds mode (src/interval.rs, src/number/f64.rs:127-171), usable on the host
// and on the device: every arithmetic operation computes with IEEE doubles and then widens the result by one ulp on each side
// through integer bit operations (next_down / next_up), so the interval always contains the exact result.
// SURVEY 8 f3 ("Interval<F64> TaylorPoly on device").
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define GTI_HD __host__ __device__ __forceinline__
#else
#define GTI_HD inline
#endif

namespace gti {

GTI_HD uint64_t f64_bits(double x) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t b;
  memcpy(&b, &x, 8);
  return b;
#endif
}
GTI_HD double bits_f64(uint64_t b) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)b);
#else
  double x;
  memcpy(&x, &b, 8);
  return x;
#endif
}
// f64.rs:127-147 / :149-171 (the standard library's next_up / next_down, strictly integer arithmetic)
GTI_HD double next_up(double x) {
  const uint64_t bits = f64_bits(x);
  if (x != x || bits == 0x7ff0000000000000ull) return x;
  const uint64_t abs = bits & 0x7fffffffffffffffull;
  const uint64_t next = abs == 0 ? 0x1ull : (bits == abs ? bits + 1 : bits - 1);
  return bits_f64(next);
}
GTI_HD double next_down(double x) {
  const uint64_t bits = f64_bits(x);
  if (x != x || bits == 0xfff0000000000000ull) return x;
  const uint64_t abs = bits & 0x7fffffffffffffffull;
  const uint64_t next = abs == 0 ? 0x8000000000000001ull : (bits == abs ? bits - 1 : bits + 1);
  return bits_f64(next);
}
GTI_HD double fmin_ref(double a, double b) { return a < b ? a : b; }   // F64::min (number/f64.rs:68-75): a if a < b else b
GTI_HD double fmax_ref(double a, double b) { return a > b ? a : b; }
GTI_HD bool finite(double x) { return x - x == 0.0; }

struct Iv {   // interval.rs:11-15
  double lo, hi;
};
GTI_HD Iv iv(double lo, double hi) { Iv r; r.lo = lo; r.hi = hi; return r; }
GTI_HD Iv iv_point(double x) { return iv(x, x); }                                   // Interval::precisely :24-26
GTI_HD Iv iv_widen(double lo, double hi) { return iv(next_down(lo), next_up(hi)); }   // :29-31
GTI_HD bool iv_is_zero(const Iv& a) { return a.lo == 0.0 && a.hi == 0.0; }            // :100-102
GTI_HD bool iv_is_one(const Iv& a) { return a.lo == 1.0 && a.hi == 1.0; }             // :112-114
GTI_HD bool iv_is_finite(const Iv& a) { return finite(a.lo) && finite(a.hi); }        // :317-319
GTI_HD bool iv_is_nan(const Iv& a) { return a.lo != a.lo || a.hi != a.hi; }
GTI_HD bool iv_contains(const Iv& a, double x) { return a.lo <= x && x <= a.hi; }
GTI_HD Iv iv_neg(const Iv& a) { return iv(-a.hi, -a.lo); }                            // :117-124
GTI_HD Iv iv_add(const Iv& a, const Iv& b) {                                          // :126-139
  if (iv_is_zero(a)) return b;
  if (iv_is_zero(b)) return a;
  return iv_widen(a.lo + b.lo, a.hi + b.hi);
}
GTI_HD Iv iv_sub(const Iv& a, const Iv& b) { return iv_add(a, iv_neg(b)); }            // :148-155
GTI_HD Iv iv_mul(const Iv& a, const Iv& b) {                                          // :164-190
  if ((iv_is_zero(a) && iv_is_finite(b)) || (iv_is_finite(a) && iv_is_zero(b))) return iv(0.0, 0.0);
  if (iv_is_one(a)) return b;
  if (iv_is_one(b)) return a;
  if (iv_is_one(iv_neg(a))) return iv_neg(b);
  if (iv_is_one(iv_neg(b))) return iv_neg(a);
  const double p = a.lo * b.lo, q = a.lo * b.hi, r = a.hi * b.lo, s = a.hi * b.hi;
  return iv_widen(fmin_ref(fmin_ref(fmin_ref(p, q), r), s), fmax_ref(fmax_ref(fmax_ref(p, q), r), s));
}
GTI_HD Iv iv_div(const Iv& a, const Iv& b) {                                          // :199-234
  if (iv_is_nan(a) || iv_is_nan(b)) return iv(NAN, NAN);
  if (iv_is_zero(a) && !iv_is_zero(b)) return a;
  if (iv_is_one(b)) return a;
  double lo = INFINITY, hi = -INFINITY;
  if (iv_contains(b, 0.0)) {
    if (0.0 <= a.lo) hi = INFINITY; else lo = -INFINITY;
    if (a.hi <= 0.0) lo = -INFINITY; else hi = INFINITY;
  }
  const double p = a.lo / b.lo, q = a.lo / b.hi, r = a.hi / b.lo, s = a.hi / b.hi;
  lo = fmin_ref(fmin_ref(fmin_ref(fmin_ref(lo, p), q), r), s);
  hi = fmax_ref(fmax_ref(fmax_ref(fmax_ref(hi, p), q), r), s);
  return iv_widen(lo, hi);
}
// exp / log (:264-276): the reference widens libm's result by one ulp.  The host uses libm too; the device's exp / log are
// accurate to 1 ulp (CUDA math API), so the device widens by two ulps to keep the enclosure property.
GTI_HD Iv iv_exp(const Iv& a) {
  if (iv_is_zero(a)) return iv(1.0, 1.0);
#if defined(__CUDA_ARCH__)
  return iv(next_down(next_down(exp(a.lo))), next_up(next_up(exp(a.hi))));
#else
  return iv_widen(std::exp(a.lo), std::exp(a.hi));
#endif
}
GTI_HD Iv iv_log(const Iv& a) {
  if (iv_is_one(a)) return iv(0.0, 0.0);
#if defined(__CUDA_ARCH__)
  return iv(next_down(next_down(log(a.lo))), next_up(next_up(log(a.hi))));
#else
  return iv_widen(std::log(a.lo), std::log(a.hi));
#endif
}
GTI_HD Iv iv_from_u32(uint32_t u) { return iv((double)u, (double)u); }                 // :79-84

}  // namespace gti
