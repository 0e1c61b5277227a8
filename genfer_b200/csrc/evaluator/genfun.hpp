// Host evaluator -- the generating-function IR.
// Restates the reference's GenFun<T> DAG for T = F64 (src/generating_function.rs: node kinds :302-323,
// constructors :50-148, operator impls :231-293) and the recognisers of compound observations (:840-914).
// Nodes are shared (the reference uses Rc); structural equality is derive(PartialEq).
#pragma once
#include <memory>
#include <unordered_map>
#include <vector>

#include "ast.hpp"

namespace gfe {

struct GfNode;
using GenFun = std::shared_ptr<const GfNode>;

struct HostPoly {  // coefficients of a `Polynomial` node: dense row-major array over `shape`
  std::vector<uint64_t> shape;
  std::vector<double> data;
};

struct GfNode {
  enum Kind {
    Var, Const, Add, Neg, Mul, Div, Polynomial, Exp, Log, Pow, UniformMgf, Subst, Derivative, TaylorPolynomial,
    TaylorCoeffAtZero, TaylorCoeff, ShiftTaylorAtZero, Max
  } kind;
  gfe::Var var = 0;           // Var / Subst / Derivative / Taylor* / Shift
  double value = 0.0;         // Const under T = F64
  Iv bounds;                  // Const under T = Interval<F64> (--bounds): encloses the exact constant, not just `value`
  uint32_t n = 0;             // Pow exponent
  size_t order = 0;           // Derivative / TaylorCoeff(AtZero) / Shift
  std::vector<size_t> orders; // TaylorPolynomial
  GenFun a, b;                // operands (Subst: a = body, b = replacement)
  std::shared_ptr<const HostPoly> poly;
  // used_vars of this (immutable) node, filled in on first use: the reference memoises per call (:424-449), which makes
  // the 12 k `observe` statements of switchpoint re-walk the whole DAG each time (44 % of the host evaluator's profile)
  mutable size_t used_vars_memo = (size_t)-1;
};

namespace gf {
inline GenFun make(GfNode n) { return std::make_shared<const GfNode>(std::move(n)); }
inline GenFun var(Var v) { GfNode n; n.kind = GfNode::Var; n.var = v; return make(n); }
inline GenFun constant(const Num& x) { GfNode n; n.kind = GfNode::Const; n.value = x.v; n.bounds = x.iv; return make(n); }
inline GenFun zero() { return constant(0.0); }
inline GenFun one() { return constant(1.0); }
inline GenFun from_u32(uint32_t k) { return constant((double)k); }
inline GenFun from_ratio(PosRatio r) { return constant(r.to_num()); }
inline GenFun bin(GfNode::Kind k, GenFun a, GenFun b) { GfNode n; n.kind = k; n.a = std::move(a); n.b = std::move(b); return make(n); }
inline GenFun un(GfNode::Kind k, GenFun a) { GfNode n; n.kind = k; n.a = std::move(a); return make(n); }
inline GenFun add(GenFun a, GenFun b) { return bin(GfNode::Add, a, b); }
inline GenFun neg(GenFun a) { return un(GfNode::Neg, a); }
inline GenFun sub(GenFun a, GenFun b) { return add(a, neg(b)); }   // :263-268: self + (-rhs)
inline GenFun mul(GenFun a, GenFun b) { return bin(GfNode::Mul, a, b); }
inline GenFun div(GenFun a, GenFun b) { return bin(GfNode::Div, a, b); }
inline GenFun exp(GenFun a) { return un(GfNode::Exp, a); }
inline GenFun log(GenFun a) { return un(GfNode::Log, a); }
inline GenFun max(GenFun a, GenFun b) { return bin(GfNode::Max, a, b); }
inline GenFun uniform_mgf(GenFun a) { return un(GfNode::UniformMgf, a); }
inline GenFun pow(GenFun a, uint32_t e) { GfNode n; n.kind = GfNode::Pow; n.a = std::move(a); n.n = e; return make(n); }
inline GenFun polynomial(std::shared_ptr<const HostPoly> p) { GfNode n; n.kind = GfNode::Polynomial; n.poly = std::move(p); return make(n); }
inline GenFun with_var_order(GfNode::Kind k, GenFun a, Var v, size_t order) {
  GfNode n; n.kind = k; n.a = std::move(a); n.var = v; n.order = order; return make(n);
}
inline GenFun derive(GenFun a, Var v, size_t order) { return with_var_order(GfNode::Derivative, a, v, order); }
inline GenFun taylor_coeff_at_zero(GenFun a, Var v, size_t order) { return with_var_order(GfNode::TaylorCoeffAtZero, a, v, order); }
inline GenFun taylor_coeff(GenFun a, Var v, size_t order) { return with_var_order(GfNode::TaylorCoeff, a, v, order); }
inline GenFun shift_down_taylor_at_zero(GenFun a, Var v, size_t order) { return with_var_order(GfNode::ShiftTaylorAtZero, a, v, order); }
inline GenFun taylor_polynomial_at_zero(GenFun a, Var v, std::vector<size_t> orders) {
  GfNode n; n.kind = GfNode::TaylorPolynomial; n.a = std::move(a); n.var = v; n.orders = std::move(orders); return make(n);
}
inline GenFun substitute_var(GenFun a, Var v, GenFun val) {
  GfNode n; n.kind = GfNode::Subst; n.a = std::move(a); n.var = v; n.b = std::move(val); return make(n);
}

// derive(PartialEq) on Rc<GeneratingFunctionKind<T>>: structural, by value
inline bool equal(const GenFun& x, const GenFun& y) {
  if (x.get() == y.get()) return true;
  if (!x || !y || x->kind != y->kind) return false;
  switch (x->kind) {
    case GfNode::Var: return x->var == y->var;
    case GfNode::Const: return x->value == y->value;
    case GfNode::Polynomial: return x->poly->shape == y->poly->shape && x->poly->data == y->poly->data;
    case GfNode::Pow: return x->n == y->n && equal(x->a, y->a);
    case GfNode::Subst: return x->var == y->var && equal(x->a, y->a) && equal(x->b, y->b);
    case GfNode::TaylorPolynomial: return x->var == y->var && x->orders == y->orders && equal(x->a, y->a);
    case GfNode::Derivative: case GfNode::TaylorCoeffAtZero: case GfNode::TaylorCoeff: case GfNode::ShiftTaylorAtZero:
      return x->var == y->var && x->order == y->order && equal(x->a, y->a);
    case GfNode::Add: case GfNode::Mul: case GfNode::Div: case GfNode::Max: return equal(x->a, y->a) && equal(x->b, y->b);
    default: return equal(x->a, y->a);
  }
}

// used_vars (:28-47, :424-449): VarRange = max var id + 1; memoised per node like the reference's cache (the DAG is
// heavily shared: a plain recursion is exponential in the number of if-statements)
inline size_t used_vars_with(const GenFun& g, std::unordered_map<const GfNode*, size_t>& cache) {
  if (g->used_vars_memo != (size_t)-1) return g->used_vars_memo;
  size_t r;
  switch (g->kind) {
    case GfNode::Var: r = g->var + 1; break;
    case GfNode::Const: r = 0; break;
    case GfNode::Polynomial: r = g->poly->shape.size(); break;
    case GfNode::Add: case GfNode::Mul: case GfNode::Div: case GfNode::Max:
      r = std::max(used_vars_with(g->a, cache), used_vars_with(g->b, cache));
      break;
    case GfNode::Subst: {
      size_t u = used_vars_with(g->a, cache);
      if (g->var + 1 == u) u = g->var;   // VarRange::remove (ppl.rs:146-153)
      r = std::max(u, used_vars_with(g->b, cache));
      break;
    }
    case GfNode::TaylorCoeffAtZero: {
      size_t u = used_vars_with(g->a, cache);
      r = g->var + 1 == u ? g->var : u;
      break;
    }
    default: r = used_vars_with(g->a, cache); break;
  }
  g->used_vars_memo = r;
  (void)cache;
  return r;
}
inline size_t used_vars(const GenFun& g) {
  std::unordered_map<const GfNode*, size_t> cache;
  return used_vars_with(g, cache);
}

struct Recognised { Var param_var; Num scalar; GenFun inner; };
// y * exp(lambda * (x - 1)) substituted for y: observation from Poisson(lambda * Y), Y discrete (:840-866)
inline bool recognize_discrete_poisson_observation(const GenFun& g, Var aux, Recognised* out) {
  if (g->kind != GfNode::Subst) return false;
  const GenFun& r = g->b;
  if (r->kind != GfNode::Mul || !equal(r->a, var(g->var))) return false;
  if (r->b->kind != GfNode::Exp) return false;
  const GenFun& e = r->b->a;
  if (e->kind != GfNode::Mul || e->a->kind != GfNode::Const) return false;
  if (!equal(e->b, sub(var(aux), one()))) return false;
  *out = {g->var, Num(e->a->value, e->a->bounds), g->a};
  return true;
}
// y + lambda * (x - 1): observation from Poisson(lambda * Y), Y continuous (:868-890)
inline bool recognize_continuous_poisson_observation(const GenFun& g, Var aux, Recognised* out) {
  if (g->kind != GfNode::Subst) return false;
  const GenFun& r = g->b;
  if (r->kind != GfNode::Add || !equal(r->a, var(g->var))) return false;
  const GenFun& m = r->b;
  if (m->kind != GfNode::Mul || m->a->kind != GfNode::Const) return false;
  if (!equal(m->b, sub(var(aux), one()))) return false;
  *out = {g->var, Num(m->a->value, m->a->bounds), g->a};
  return true;
}
// y * (p / (1 - (1-p) x)): observation from NegBinomial(Y, p) (:892-914)
inline bool recognize_negative_binomial_observation(const GenFun& g, Var aux, Recognised* out) {
  if (g->kind != GfNode::Subst) return false;
  const GenFun& r = g->b;
  if (r->kind != GfNode::Mul || !equal(r->a, var(g->var))) return false;
  const GenFun& d = r->b;
  if (d->kind != GfNode::Div || d->a->kind != GfNode::Const) return false;
  const Num p(d->a->value, d->a->bounds);
  GenFun expected = sub(one(), mul(constant(Num(1.0) - p), var(aux)));   // compared by its f64 value
  if (!equal(d->b, expected)) return false;
  *out = {g->var, p, g->a};
  return true;
}
}  // namespace gf

}  // namespace gfe
