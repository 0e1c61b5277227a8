// Host evaluator -- symbolic mode (`-s`): the generating function is turned into ONE univariate computation DAG
// (symbolic differentiation / Taylor coefficients with symbolic entries) and that DAG is evaluated over TaylorExpansion<F64>.
// Restates the reference's src/symbolic.rs (term kinds :300-311, simplifying constructors :397-600, substitute :603-704,
// derive :706-786, taylor_coeffs_with :788-844, evaluate :324-373, probs_symbolic / moments_symbolic :238-299), the generic
// TaylorExpansion<T> of src/univariate_taylor.rs instantiated at T = SymGenFun (its coefficients are terms, its arithmetic
// builds terms), and GenFun::to_computation (src/generating_function.rs:767-831, fold_coeffs :916-934).
//
// Everything here is pointer-DAG construction on the host plus constant folding of literals; the numeric work -- evaluating the
// final term over univariate Taylor expansions -- goes through the backend's univariate operators (gtu_* on the device for the
// product, the oracle's TaylorExpansion<f64> for the checker).
#pragma once
#include <functional>
#include <unordered_map>

#include "eval.hpp"

namespace gfe {

struct SymNode;
using Sym = std::shared_ptr<const SymNode>;
struct SymNode {
  enum Kind { Variable, Lit, Add, Mul, Div, Exp, Log, Pow, Max } kind;
  Var var = 0;
  double lit = 0.0;
  uint32_t n = 0;
  Sym a, b;
};

namespace sym {
inline Sym mk(SymNode::Kind k, Sym a = nullptr, Sym b = nullptr) { SymNode n; n.kind = k; n.a = std::move(a); n.b = std::move(b); return std::make_shared<const SymNode>(std::move(n)); }
inline Sym lit(double x) { SymNode n; n.kind = SymNode::Lit; n.lit = x; return std::make_shared<const SymNode>(std::move(n)); }
inline Sym var(Var v) { SymNode n; n.kind = SymNode::Variable; n.var = v; return std::make_shared<const SymNode>(std::move(n)); }
inline Sym zero() { return lit(0.0); }
inline Sym one() { return lit(1.0); }
inline bool is_lit(const Sym& s) { return s->kind == SymNode::Lit; }
inline bool lit_zero(const Sym& s) { return is_lit(s) && s->lit == 0.0; }
inline bool lit_one(const Sym& s) { return is_lit(s) && s->lit == 1.0; }
inline bool same(const Sym& x, const Sym& y) { return x.get() == y.get(); }
Sym mul(Sym lhs, Sym rhs);
Sym exp(Sym arg);
Sym pow(Sym base, uint32_t e);

inline Sym add(Sym lhs, Sym rhs) {   // :397-430
  if (lit_zero(lhs)) return rhs;
  if (lit_zero(rhs)) return lhs;
  if (is_lit(lhs) && is_lit(rhs)) return lit(lhs->lit + rhs->lit);
  if (is_lit(lhs) && rhs->kind == SymNode::Add) {
    if (is_lit(rhs->b)) return add(rhs->a, lit(lhs->lit + rhs->b->lit));
    return mk(SymNode::Add, rhs, lhs);
  }
  if (lhs->kind == SymNode::Add && is_lit(rhs)) {
    if (is_lit(lhs->a)) return add(lhs->b, lit(rhs->lit + lhs->a->lit));
    return mk(SymNode::Add, lhs, rhs);
  }
  if (lhs->kind == SymNode::Add && rhs->kind == SymNode::Add) {
    const Sym &a = lhs->a, &b = lhs->b, &c = rhs->a, &d = rhs->b;
    if (is_lit(b) && is_lit(d)) return add(add(a, c), lit(b->lit + d->lit));
    if (is_lit(b)) return add(add(a, rhs), b);
    if (is_lit(d)) return add(add(lhs, c), d);
  }
  return mk(SymNode::Add, lhs, rhs);
}

inline Sym mul(Sym lhs, Sym rhs) {   // :432-548
  // literal simplifications
  if (lit_zero(lhs) || lit_zero(rhs)) return zero();
  if (lit_one(lhs)) return rhs;
  if (lit_one(rhs)) return lhs;
  if (lhs->kind == SymNode::Exp && rhs->kind == SymNode::Exp) return exp(add(lhs->a, rhs->a));
  if (is_lit(lhs) && is_lit(rhs)) return lit(lhs->lit * rhs->lit);
  if (is_lit(lhs) && rhs->kind == SymNode::Mul) {
    if (is_lit(rhs->a)) return mk(SymNode::Mul, lit(lhs->lit * rhs->a->lit), rhs->b);
  } else if (lhs->kind == SymNode::Mul && is_lit(rhs)) {
    if (is_lit(lhs->a)) return mk(SymNode::Mul, lit(rhs->lit * lhs->a->lit), lhs->b);
  }
  // exp simplifications
  {
    const Sym* m = nullptr;
    const Sym* e = nullptr;
    if (lhs->kind == SymNode::Mul && rhs->kind == SymNode::Exp) { m = &lhs; e = &rhs; }
    else if (lhs->kind == SymNode::Exp && rhs->kind == SymNode::Mul) { m = &rhs; e = &lhs; }
    if (m) {
      const Sym &a1 = (*m)->a, &a2 = (*m)->b, &bb = (*e)->a;
      if (a2->kind == SymNode::Exp) return mul(a1, exp(add(a2->a, bb)));
      if (a1->kind == SymNode::Exp) return mul(a2, exp(add(a1->a, bb)));
    } else if (lhs->kind == SymNode::Mul && rhs->kind == SymNode::Mul) {
      const Sym &a1 = lhs->a, &a2 = lhs->b, &b1 = rhs->a, &b2 = rhs->b;
      const bool ea1 = a1->kind == SymNode::Exp, ea2 = a2->kind == SymNode::Exp, eb1 = b1->kind == SymNode::Exp, eb2 = b2->kind == SymNode::Exp;
      if (ea1 && eb1) return mul(mul(a2, b2), exp(add(a1->a, b1->a)));
      if (ea1 && eb2) return mul(mul(a2, b1), exp(add(a1->a, b2->a)));
      if (ea2 && eb1) return mul(mul(a1, b2), exp(add(a2->a, b1->a)));
      if (ea2 && eb2) return mul(mul(a1, b1), exp(add(a2->a, b2->a)));
    }
  }
  // moving literals left
  if (lhs->kind == SymNode::Mul && rhs->kind == SymNode::Mul) {
    if (is_lit(lhs->a) && is_lit(rhs->a)) return mk(SymNode::Mul, lit(lhs->a->lit * rhs->a->lit), mul(lhs->b, rhs->b));
  } else if (lhs->kind == SymNode::Mul) {
    if (is_lit(lhs->a)) return mk(SymNode::Mul, lhs->a, mul(lhs->b, rhs));
  } else if (rhs->kind == SymNode::Mul) {
    if (is_lit(rhs->a)) return mk(SymNode::Mul, rhs->a, mul(rhs->b, lhs));
  }
  // pow simplifications
  if (lhs->kind == SymNode::Mul) {
    const Sym &a1 = lhs->a, &a2 = lhs->b;
    if (same(a2, rhs)) return mul(a1, pow(a2, 2));
    if (rhs->kind == SymNode::Pow && same(a2, rhs->a)) return mul(a1, pow(a2, rhs->n + 1));
    if (rhs->kind == SymNode::Pow && a2->kind == SymNode::Pow && same(a2->a, rhs->a)) return mul(a1, pow(a2->a, a2->n + rhs->n));
  }
  if (is_lit(rhs)) return mk(SymNode::Mul, rhs, lhs);
  return mk(SymNode::Mul, lhs, rhs);
}

inline Sym div(Sym lhs, Sym rhs) {   // :550-560
  if (lit_zero(lhs)) return zero();
  if (lit_one(rhs)) return lhs;
  return mk(SymNode::Div, lhs, rhs);
}
inline Sym neg(Sym arg) { return mul(lit(-1.0), arg); }   // :562-567
inline Sym exp(Sym arg) {   // :569-584
  if (lit_zero(arg)) return one();
  if (is_lit(arg)) return lit(std::exp(arg->lit));
  if (arg->kind == SymNode::Add && is_lit(arg->b)) return mul(lit(std::exp(arg->b->lit)), exp(arg->a));
  return mk(SymNode::Exp, arg);
}
inline Sym log(Sym arg) {   // :586-601
  if (lit_one(arg)) return zero();
  if (is_lit(arg)) return lit(std::log(arg->lit));
  if (arg->kind == SymNode::Mul && is_lit(arg->a)) return add(log(arg->b), lit(std::log(arg->a->lit)));
  return mk(SymNode::Log, arg);
}
inline Sym pow(Sym base, uint32_t e) {   // :603-618
  if (e == 0) return one();
  if (e == 1) return base;
  if (lit_zero(base)) return zero();
  if (lit_one(base)) return one();
  SymNode n;
  n.kind = SymNode::Pow;
  n.a = std::move(base);
  n.n = e;
  return std::make_shared<const SymNode>(std::move(n));
}
inline Sym max(Sym a, Sym b) { return mk(SymNode::Max, a, b); }
}  // namespace sym

// SymGenFun<T> as a Number (:16-236): the coefficient type of the symbolic Taylor expansions below
struct SymVal {
  Sym root;
  SymVal() : root(sym::zero()) {}
  explicit SymVal(Sym r) : root(std::move(r)) {}
  static SymVal zero() { return SymVal(sym::zero()); }
  static SymVal one() { return SymVal(sym::one()); }
  static SymVal from_u32(uint32_t k) { return SymVal(sym::lit((double)k)); }
  bool is_zero() const { return sym::lit_zero(root); }
  bool is_one() const { return sym::lit_one(root); }
  SymVal exp() const { return SymVal(sym::exp(root)); }
  SymVal log() const { return SymVal(sym::log(root)); }
  SymVal pow(uint32_t e) const { return SymVal(sym::pow(root, e)); }
  SymVal max(const SymVal& o) const { return SymVal(sym::max(root, o.root)); }
  friend SymVal operator+(const SymVal& x, const SymVal& y) { return SymVal(sym::add(x.root, y.root)); }
  friend SymVal operator-(const SymVal& x) { return SymVal(sym::neg(x.root)); }
  friend SymVal operator-(const SymVal& x, const SymVal& y) { return x + (-y); }   // :196-201
  friend SymVal operator*(const SymVal& x, const SymVal& y) { return SymVal(sym::mul(x.root, y.root)); }
  friend SymVal operator/(const SymVal& x, const SymVal& y) { return SymVal(sym::div(x.root, y.root)); }
};

// TaylorExpansion<T> (univariate_taylor.rs) for a coefficient type with SymVal's interface
template <class T>
struct UniSeries {
  bool constant = true;
  T c;                     // Constant
  std::vector<T> coeffs;   // Polynomial
  static UniSeries cst(T x) { UniSeries r; r.constant = true; r.c = std::move(x); return r; }
  static UniSeries poly(std::vector<T> v) { UniSeries r; r.constant = false; r.coeffs = std::move(v); return r; }
  static UniSeries zero() { return cst(T::zero()); }
  static UniSeries one() { return cst(T::one()); }
  static UniSeries var(T x, size_t order) {   // :16-23
    std::vector<T> v(order + 1, T::zero());
    if (1 < v.size()) v[1] = T::one();
    v[0] = std::move(x);
    return poly(std::move(v));
  }
  T coeff(size_t order) const {   // :25-37
    if (!constant) { GFE_ASSERT(order < coeffs.size(), "coeff: index out of bounds"); return coeffs[order]; }
    return order == 0 ? c : T::zero();
  }
  UniSeries exp() const {   // :163-181
    if (constant) return cst(c.exp());
    const size_t order = coeffs.size();
    std::vector<T> res(order, T::zero());
    res[0] = coeffs[0].exp();
    for (size_t k = 1; k < order; k++) {
      T sum = T::zero();
      for (size_t j = 1; j <= k; j++) sum = sum + res[k - j] * coeffs[j] * T::from_u32((uint32_t)j);
      res[k] = sum / T::from_u32((uint32_t)k);
    }
    return poly(std::move(res));
  }
  UniSeries log() const {   // :183-203
    if (constant) return cst(c.log());
    const size_t order = coeffs.size();
    std::vector<T> res(order, T::zero());
    res[0] = coeffs[0].log();
    for (size_t k = 1; k < order; k++) {
      T sum = T::zero();
      for (size_t j = 1; j < k; j++) sum = sum + coeffs[k - j] * res[j] * T::from_u32((uint32_t)j);
      res[k] = (coeffs[k] * T::from_u32((uint32_t)k) - sum) / coeffs[0] / T::from_u32((uint32_t)k);
    }
    return poly(std::move(res));
  }
  UniSeries pow(uint32_t e) const {   // :205-217 (binary exponentiation, including the last wasted squaring)
    UniSeries res = one(), base = *this;
    while (e > 0) {
      if (e & 1) res = res * base;
      base = base * base;
      e >>= 1;
    }
    return res;
  }
  friend UniSeries operator+(const UniSeries& lhs, const UniSeries& rhs) {   // AddAssign :286-315
    if (rhs.constant) {
      if (lhs.constant) return cst(lhs.c + rhs.c);
      UniSeries r = lhs;
      r.coeffs[0] = r.coeffs[0] + rhs.c;
      return r;
    }
    std::vector<T> ws = rhs.coeffs;
    if (lhs.constant) {
      ws[0] = ws[0] + lhs.c;
      return poly(std::move(ws));
    }
    const size_t order = std::min(lhs.coeffs.size(), ws.size());
    for (size_t i = 0; i < order; i++) ws[i] = ws[i] + lhs.coeffs[i];
    ws.resize(order, T::zero());
    return poly(std::move(ws));
  }
  friend UniSeries operator-(const UniSeries& x) {   // :318-329
    if (x.constant) return cst(-x.c);
    std::vector<T> v;
    for (const T& t : x.coeffs) v.push_back(-t);
    return poly(std::move(v));
  }
  friend UniSeries operator-(const UniSeries& lhs, const UniSeries& rhs) {   // SubAssign :340-370
    if (rhs.constant) {
      if (lhs.constant) return cst(lhs.c - rhs.c);
      UniSeries r = lhs;
      r.coeffs[0] = r.coeffs[0] - rhs.c;
      return r;
    }
    std::vector<T> ws = rhs.coeffs;
    if (lhs.constant) {
      for (T& w : ws) w = -w;
      ws[0] = ws[0] + lhs.c;
      return poly(std::move(ws));
    }
    const size_t order = std::min(lhs.coeffs.size(), ws.size());
    for (size_t i = 0; i < order; i++) ws[i] = lhs.coeffs[i] - ws[i];
    ws.resize(order, T::zero());
    return poly(std::move(ws));
  }
  friend UniSeries operator*(const UniSeries& lhs, const UniSeries& rhs) {   // :372-397
    if (lhs.constant && rhs.constant) return cst(lhs.c * rhs.c);
    if (lhs.constant || rhs.constant) {
      const T& k = lhs.constant ? lhs.c : rhs.c;
      std::vector<T> v = lhs.constant ? rhs.coeffs : lhs.coeffs;
      for (T& t : v) t = t * k;
      return poly(std::move(v));
    }
    const size_t order = std::min(lhs.coeffs.size(), rhs.coeffs.size());
    std::vector<T> res(order, T::zero());
    for (size_t k = 0; k < order; k++) {
      T sum = T::zero();
      for (size_t j = 0; j <= k; j++) sum = sum + lhs.coeffs[j] * rhs.coeffs[k - j];
      res[k] = sum;
    }
    return poly(std::move(res));
  }
  friend UniSeries operator/(const UniSeries& lhs, const UniSeries& rhs) {   // :405-446
    if (lhs.constant && rhs.constant) return cst(lhs.c / rhs.c);
    if (rhs.constant) {
      std::vector<T> v = lhs.coeffs;
      for (T& t : v) t = t / rhs.c;
      return poly(std::move(v));
    }
    const std::vector<T>& ws = rhs.coeffs;
    const T scale = T::one() / ws[0];
    if (lhs.constant) {
      const size_t order = ws.size();
      std::vector<T> res(order, T::zero());
      res[0] = lhs.c * scale;
      for (size_t k = 1; k < order; k++) {
        T sum = T::zero();
        for (size_t i = 0; i < k; i++) sum = sum - res[i] * ws[k - i];
        res[k] = scale * sum;
      }
      return poly(std::move(res));
    }
    const size_t order = std::min(lhs.coeffs.size(), ws.size());
    std::vector<T> res(order, T::zero());
    res[0] = scale * lhs.coeffs[0];
    for (size_t k = 1; k < order; k++) {
      T sum = lhs.coeffs[k];
      for (size_t i = 0; i < k; i++) sum = sum - res[i] * ws[k - i];
      res[k] = scale * sum;
    }
    return poly(std::move(res));
  }
};

namespace sym {
using SymSeries = UniSeries<SymVal>;

// substitute (:603-704): variables for which `map` yields a term are replaced; untouched sub-terms keep their identity
inline Sym substitute_with(const Sym& t, const std::function<Sym(Var)>& map, std::unordered_map<const SymNode*, Sym>& cache) {
  auto it = cache.find(t.get());
  if (it != cache.end()) return it->second;
  Sym r;
  switch (t->kind) {
    case SymNode::Variable: { Sym v = map(t->var); r = v ? v : t; break; }
    case SymNode::Lit: r = t; break;
    case SymNode::Add: case SymNode::Mul: case SymNode::Div: case SymNode::Max: {
      Sym a2 = substitute_with(t->a, map, cache), b2 = substitute_with(t->b, map, cache);
      if (same(t->a, a2) && same(t->b, b2)) r = t;
      else if (t->kind == SymNode::Add) r = add(a2, b2);
      else if (t->kind == SymNode::Mul) r = mul(a2, b2);
      else if (t->kind == SymNode::Div) r = div(a2, b2);
      else r = max(a2, b2);
      break;
    }
    case SymNode::Exp: case SymNode::Log: case SymNode::Pow: {
      Sym a2 = substitute_with(t->a, map, cache);
      if (same(t->a, a2)) r = t;
      else if (t->kind == SymNode::Exp) r = exp(a2);
      else if (t->kind == SymNode::Log) r = log(a2);
      else r = pow(a2, t->n);
      break;
    }
  }
  cache[t.get()] = r;   // the reference caches shared nodes only; an unshared node is visited once either way
  return r;
}
inline Sym substitute_var(const Sym& t, Var v, const Sym& val) {
  std::unordered_map<const SymNode*, Sym> cache;
  return substitute_with(t, [&](Var w) { return w == v ? val : Sym(); }, cache);
}

// derive (:706-786)
inline Sym derive_with(const Sym& t, Var v, std::unordered_map<const SymNode*, Sym>& cache) {
  auto it = cache.find(t.get());
  if (it != cache.end()) return it->second;
  Sym r;
  switch (t->kind) {
    case SymNode::Variable: r = t->var == v ? one() : zero(); break;
    case SymNode::Lit: r = zero(); break;
    case SymNode::Add: { Sym da = derive_with(t->a, v, cache), db = derive_with(t->b, v, cache); r = add(da, db); break; }
    case SymNode::Mul: {
      Sym da = derive_with(t->a, v, cache), db = derive_with(t->b, v, cache);
      Sym x = mul(t->a, db), y = mul(t->b, da);
      r = add(x, y);
      break;
    }
    case SymNode::Div: {
      Sym da = derive_with(t->a, v, cache), db = derive_with(t->b, v, cache);
      Sym x = mul(t->a, db), y = mul(t->b, da);
      Sym b2 = pow(t->b, 2);
      r = div(add(x, neg(y)), b2);
      break;
    }
    case SymNode::Exp: { Sym da = derive_with(t->a, v, cache); r = mul(da, t); break; }
    case SymNode::Log: { Sym da = derive_with(t->a, v, cache); r = div(da, t->a); break; }
    case SymNode::Pow: {
      GFE_ASSERT(t->n != 0, "unexpected 0 exponent, should have been simplified away before");
      Sym da = derive_with(t->a, v, cache);
      Sym am1 = pow(t->a, t->n - 1);
      r = mul(mul(lit((double)t->n), da), am1);
      break;
    }
    case SymNode::Max: throw EvalError("Maximum shouldn't be differentiated.");
  }
  cache[t.get()] = r;
  return r;
}
inline Sym derive(const Sym& t, Var v) {
  std::unordered_map<const SymNode*, Sym> cache;
  return derive_with(t, v, cache);
}

// taylor_coeffs_with (:788-844): Taylor expansion in `v` (around the literal x, or around the variable itself) with terms as
// coefficients
inline SymSeries taylor_coeffs_with(const Sym& t, Var v, const double* x, size_t order, std::unordered_map<const SymNode*, SymSeries>& cache) {
  auto it = cache.find(t.get());
  if (it != cache.end()) return it->second;
  SymSeries r;
  switch (t->kind) {
    case SymNode::Variable:
      if (t->var == v) r = SymSeries::var(SymVal(x ? lit(*x) : var(v)), order);
      else r = SymSeries::cst(SymVal(t));
      break;
    case SymNode::Lit: r = SymSeries::cst(SymVal(t)); break;
    case SymNode::Add: r = taylor_coeffs_with(t->a, v, x, order, cache) + taylor_coeffs_with(t->b, v, x, order, cache); break;
    case SymNode::Mul: r = taylor_coeffs_with(t->a, v, x, order, cache) * taylor_coeffs_with(t->b, v, x, order, cache); break;
    case SymNode::Div: r = taylor_coeffs_with(t->a, v, x, order, cache) / taylor_coeffs_with(t->b, v, x, order, cache); break;
    case SymNode::Exp: r = taylor_coeffs_with(t->a, v, x, order, cache).exp(); break;
    case SymNode::Log: r = taylor_coeffs_with(t->a, v, x, order, cache).log(); break;
    case SymNode::Pow: r = taylor_coeffs_with(t->a, v, x, order, cache).pow(t->n); break;
    case SymNode::Max: throw EvalError("Maximum shouldn't be differentiated.");
  }
  cache[t.get()] = r;
  return r;
}
inline SymSeries taylor_coeffs(const Sym& t, Var v, const double* x, size_t order) {
  std::unordered_map<const SymNode*, SymSeries> cache;
  return taylor_coeffs_with(t, v, x, order, cache);
}

// fold_coeffs (generating_function.rs:916-934) of a Polynomial node
inline Sym fold_coeffs(const HostPoly& hp, size_t ndim, size_t offset, size_t stride_elems) {
  if (ndim == 0) return lit(hp.data[offset]);
  const size_t v = ndim - 1;
  // axis v is the last axis of the current view: its stride is 1 at the top level and grows as trailing axes are fixed
  size_t stride = stride_elems;
  const size_t len = (size_t)hp.shape[v];
  Sym result = zero();
  for (size_t i = len; i-- > 0;) {
    result = mul(result, var(v));
    Sym c = fold_coeffs(hp, ndim - 1, offset + i * stride, stride * len);
    result = add(result, c);
  }
  return result;
}

// GenFun::to_computation (generating_function.rs:767-831)
inline Sym to_computation(const GenFun& g) {
  const GfNode& n = *g;
  switch (n.kind) {
    case GfNode::Var: return var(n.var);
    case GfNode::Const: return lit(n.value);
    case GfNode::Add: { Sym a = to_computation(n.a); Sym b = to_computation(n.b); return add(a, b); }
    case GfNode::Neg: return neg(to_computation(n.a));
    case GfNode::Mul: { Sym a = to_computation(n.a); Sym b = to_computation(n.b); return mul(a, b); }
    case GfNode::Div: { Sym a = to_computation(n.a); Sym b = to_computation(n.b); return div(a, b); }
    case GfNode::Polynomial: return fold_coeffs(*n.poly, n.poly->shape.size(), 0, 1);
    case GfNode::Exp: return exp(to_computation(n.a));
    case GfNode::Log: return log(to_computation(n.a));
    case GfNode::Pow: return pow(to_computation(n.a), n.n);
    case GfNode::Max: { Sym a = to_computation(n.a); Sym b = to_computation(n.b); return max(a, b); }
    case GfNode::UniformMgf: { Sym gc = to_computation(n.a); return div(add(exp(gc), neg(one())), gc); }   // (e^g - 1) / g
    case GfNode::Subst: { Sym s = to_computation(n.b); return substitute_var(to_computation(n.a), n.var, s); }
    case GfNode::Derivative: {
      Sym d = to_computation(n.a);
      for (size_t i = 0; i < n.order; i++) d = derive(d, n.var);
      return d;
    }
    case GfNode::TaylorPolynomial: {
      size_t max_order = 0;
      for (size_t o : n.orders) max_order = std::max(max_order, o);
      SymSeries taylor = taylor_coeffs(to_computation(n.a), n.var, nullptr, max_order);
      std::vector<bool> keep(max_order + 1, false);
      for (size_t o : n.orders) keep[o] = true;
      Sym acc = lit(0.0);
      for (size_t i = max_order + 1; i-- > 0;) {
        if (keep[i]) acc = add(mul(acc, var(n.var)), taylor.coeff(i).root);
        else acc = mul(acc, var(n.var));
      }
      return acc;
    }
    case GfNode::TaylorCoeffAtZero: {
      const double z = 0.0;
      return taylor_coeffs(to_computation(n.a), n.var, &z, n.order).coeff(n.order).root;
    }
    case GfNode::TaylorCoeff: return taylor_coeffs(to_computation(n.a), n.var, nullptr, n.order).coeff(n.order).root;
    case GfNode::ShiftTaylorAtZero: throw EvalError("not yet implemented");   // todo!() in the reference
  }
  throw EvalError("unreachable");
}
}  // namespace sym

// Evaluation of a term over the backend's univariate Taylor expansions (SymGenFunKind::evaluate_with :337-373): memoised by
// node, every operator is one backend call (gtu_* on the device).
template <class B>
class SymEvaluator {
 public:
  using U = typename B::Uni;
  explicit SymEvaluator(B& b) : b_(b) {}
  U evaluate(const Sym& t, const std::function<U(Var)>& var_map) {
    cache_.clear();
    return eval(t, var_map);
  }
  // probs_symbolic (:238-259)
  std::vector<double> probs(const Sym& pgf, Var v, const VarSupport& vi, size_t n) {
    U var = b_.uni_var(0.0, n);
    U t = evaluate(pgf, [&](Var w) { return w == v ? var : (vi[w].is_discrete() ? b_.uni_constant(1.0) : b_.uni_constant(0.0)); });
    std::vector<double> out;
    for (size_t i = 0; i < n; i++) out.push_back(b_.uni_coeff(t, i));
    return out;
  }
  // moments_symbolic (:261-299)
  std::pair<double, std::vector<double>> moments(const Sym& pgf, Var v, const VarSupport& vi, size_t limit) {
    U var = b_.uni_var(vi[v].is_discrete() ? 1.0 : 0.0, limit);
    U t = evaluate(pgf, [&](Var w) { return w == v ? var : (vi[w].is_discrete() ? b_.uni_constant(1.0) : b_.uni_constant(0.0)); });
    std::vector<double> result;
    double factor = 1.0;
    for (size_t i = 0; i < limit; i++) {
      result.push_back(b_.uni_coeff(t, i) * factor);
      factor *= (double)(uint32_t)(i + 1);
    }
    if (vi[v].is_discrete()) return Evaluator<B>::factorial_moments_to_moments(result);
    double total = result[0];
    std::vector<double> moments;
    for (size_t i = 1; i < result.size(); i++) moments.push_back(result[i] / total);
    return {total, moments};
  }
  // evaluate_closed (:92-99)
  double closed(const Sym& t) {
    U r = evaluate(t, [&](Var) -> U { throw EvalError("term should be closed"); });
    return b_.uni_coeff(r, 0);
  }
  size_t nodes_evaluated = 0;

 private:
  U eval(const Sym& t, const std::function<U(Var)>& var_map) {
    auto it = cache_.find(t.get());
    if (it != cache_.end()) return it->second;
    U r;
    switch (t->kind) {
      case SymNode::Variable: r = var_map(t->var); break;
      case SymNode::Lit: r = b_.uni_constant(t->lit); break;
      case SymNode::Add: { U a = eval(t->a, var_map); U b = eval(t->b, var_map); r = b_.uni_add(a, b); break; }
      case SymNode::Mul: { U a = eval(t->a, var_map); U b = eval(t->b, var_map); r = b_.uni_mul(a, b); break; }
      case SymNode::Div: { U a = eval(t->a, var_map); U b = eval(t->b, var_map); r = b_.uni_div(a, b); break; }
      case SymNode::Exp: r = b_.uni_exp(eval(t->a, var_map)); break;
      case SymNode::Log: r = b_.uni_log(eval(t->a, var_map)); break;
      case SymNode::Pow: r = b_.uni_pow(eval(t->a, var_map), t->n); break;
      case SymNode::Max: { U a = eval(t->a, var_map); U b = eval(t->b, var_map); r = b_.uni_max(a, b); break; }
    }
    nodes_evaluated++;
    cache_[t.get()] = r;
    return r;
  }
  B& b_;
  std::unordered_map<const SymNode*, U> cache_;
};

}  // namespace gfe
