// Host evaluator -- GenFun::simplify / eval / probs_taylor / moments_taylor over a TaylorPoly backend.
// Restates the host recursion of the reference's src/generating_function.rs: simplify :474-545 (exact polynomial
// algebra with degrees_p1 = usize::MAX), eval_with (pointer-keyed memo, :186-222), eval :548-668, the observation
// fast paths of eval_taylor_coeff_at_zero :670-765, probs_taylor :937-967, moments_taylor :970-1005 and
// factorial_moments_to_moments :1008-1033.
//
// `B` is the TaylorPoly backend.  In the product it is the C ABI of libgenfer_taylor (GpuBackend in
// eval_api.cpp): every arithmetic operation below is then a call into the CUDA library.  The oracle instantiates
// the same host logic over its CPU restatement (oracle/oracle_eval.cpp) -- that is test infrastructure.
#pragma once
#include <cmath>
#include <map>
#include <unordered_map>

#include "genfun.hpp"
#include "support.hpp"

namespace gfe {

constexpr uint64_t UNBOUNDED = UINT64_MAX;

// A GenFun constant as the backend's scalar: its f64 value for T = F64, its enclosure for an interval scalar type (which
// provides `from_bounds(lo, hi)`, `lower()`, `upper()`).
template <class S>
inline S scalar_of(const Num& x) {
  if constexpr (std::is_same<S, double>::value) return x.v;
  else return S::from_bounds(x.iv.lo, x.iv.hi);
}
template <class S>
inline Iv bounds_of(const S& x) {
  if constexpr (std::is_same<S, double>::value) return Iv::precisely(x);   // print_moments_and_probs (main.rs:256-289)
  else return Iv::exact(x.lower(), x.upper());
}

template <class B>
class Evaluator {
 public:
  using Poly = typename B::Poly;
  // The scalar type T of the reference's TaylorPoly<T>: double in the product (GpuBackend) and in the f64 oracle; the
  // --bounds instantiations (oracle and gti_* on the device) use Interval<F64>; a GenFun constant then enters as the enclosure
  // Number::from_ratio builds for it (num.hpp), not as the rounded f64 value.
  using S = typename B::Scalar;
  explicit Evaluator(B& backend) : b_(backend) {}

  // ---- simplify (:152-177, :474-545) ---------------------------------------------------------------
  GenFun simplify(const GenFun& g) {
    simp_cache_.clear();
    std::optional<Poly> p = simplify_with(g);
    if (!p) return g;
    auto hp = std::make_shared<HostPoly>();
    hp->shape = b_.array_shape(*p);
    hp->data = b_.to_host(*p);
    return gf::polynomial(hp);
  }

  // ---- eval (:179-222) -----------------------------------------------------------------------------------
  Poly eval(const GenFun& g, const std::vector<S>& inputs, size_t degree_p1) {
    eval_cache_.clear();
    return eval_with(g, inputs, degree_p1);
  }

  // probs_taylor (:937-967): p(0..max_n) of variable v
  std::vector<S> probs_taylor(const GenFun& pgf, Var v, const VarSupport& vi, size_t max_n) {
    GFE_ASSERT(vi[v].is_discrete(), "Can only compute probabilities for discrete variables");
    std::vector<S> substs(vi.num_vars(), S(0.0));
    for (size_t i = 0; i < substs.size(); i++) substs[i] = vi[i].is_discrete() ? S(1.0) : S(0.0);
    substs[v] = S(0.0);
    Poly expansion = eval(pgf, substs, max_n + 1);
    return b_.gather_axis(expansion, v, max_n);   // `max_n` coefficient() reads as ONE gather
  }

  // moments_taylor (:970-1005): (total, raw moments of order 1..limit-1)
  std::pair<S, std::vector<S>> moments_taylor(const GenFun& pgf, Var v, const VarSupport& vi, size_t limit) {
    std::vector<S> substs(vi.num_vars(), S(0.0));
    for (size_t i = 0; i < substs.size(); i++) substs[i] = vi[i].is_discrete() ? S(1.0) : S(0.0);
    Poly expansion = eval(pgf, substs, limit);
    std::vector<S> coeffs = b_.gather_axis(expansion, v, limit);
    std::vector<S> result;
    S factor = S(1.0);
    for (size_t i = 0; i < limit; i++) {
      result.push_back(coeffs[i] * factor);
      factor = factor * S((double)(uint32_t)(i + 1));
    }
    if (vi[v].is_discrete()) return factorial_moments_to_moments(result);
    S total = result[0];
    std::vector<S> moments;
    for (size_t i = 1; i < result.size(); i++) moments.push_back(result[i] / total);
    return {total, moments};
  }

  static std::pair<S, std::vector<S>> factorial_moments_to_moments(const std::vector<S>& fm) {  // :1008-1033
    const size_t len = fm.size();
    std::vector<std::vector<S>> st(len, std::vector<S>(len, S(0.0)));
    for (size_t n = 0; n < len; n++) {
      st[n][0] = S(0.0);
      st[n][n] = S(1.0);
      for (size_t k = 1; k < n; k++) st[n][k] = st[n - 1][k - 1] + S((double)(uint32_t)k) * st[n - 1][k];
    }
    S total = fm[0];
    std::vector<S> moments(len - 1, S(0.0));
    for (size_t n = 1; n < len; n++)
      for (size_t k = 0; k <= n; k++) moments[n - 1] = moments[n - 1] + st[n][k] * fm[k];
    for (S& m : moments) m = m / total;
    return {total, moments};
  }

  size_t cache_hits = 0, nodes_evaluated = 0;

 private:
  B& b_;
  std::unordered_map<const GfNode*, std::pair<GenFun, std::optional<Poly>>> simp_cache_;
  // the entry keeps its node alive (like EvalResult.gf in the reference): temporaries built during evaluation are
  // freed again, and a later node allocated at the same address must not hit a stale entry
  struct EvalEntry { GenFun node; std::vector<S> inputs; size_t degree_p1; Poly output; };
  std::unordered_map<const GfNode*, EvalEntry> eval_cache_;

  static bool same_inputs(const std::vector<S>& a, const std::vector<S>& b) {
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); i++)
      if (!(a[i] == b[i])) return false;   // PartialEq on f64
    return true;
  }

  std::optional<Poly> simplify_with(const GenFun& g) {
    const bool shared = g.use_count() > 1;
    if (shared) {
      auto it = simp_cache_.find(g.get());
      if (it != simp_cache_.end()) return it->second.second;
    }
    std::optional<Poly> r = simplify_node(*g);
    if (shared) simp_cache_[g.get()] = std::make_pair(g, r);
    return r;
  }

  std::optional<Poly> simplify_node(const GfNode& n) {
    switch (n.kind) {
      case GfNode::Var: return b_.var_with_degrees(n.var, S(0.0), std::vector<uint64_t>(n.var + 1, UNBOUNDED));
      case GfNode::Const: return b_.from_scalar(scalar_of<S>(Num(n.value, n.bounds)));
      case GfNode::Add: case GfNode::Mul: case GfNode::Div: {
        auto p1 = simplify_with(n.a);
        auto p2 = simplify_with(n.b);
        if (!p1 || !p2) return std::nullopt;
        if (n.kind == GfNode::Add) return b_.add(*p1, *p2);
        if (n.kind == GfNode::Mul) return b_.mul(*p1, *p2);
        if (b_.extract_constant(*p2)) return b_.div(*p1, *p2);   // only division by constants stays polynomial
        return std::nullopt;
      }
      case GfNode::Neg: { auto p = simplify_with(n.a); if (!p) return std::nullopt; return b_.neg(*p); }
      case GfNode::Polynomial: case GfNode::Exp: case GfNode::Log: case GfNode::Max: case GfNode::UniformMgf:
        return std::nullopt;
      case GfNode::Pow: { auto p = simplify_with(n.a); if (!p) return std::nullopt; return b_.pow(*p, n.n); }
      case GfNode::Subst: {
        auto p = simplify_with(n.a);
        auto q = simplify_with(n.b);
        if (!p || !q) return std::nullopt;
        return b_.subst_var(*p, n.var, *q);
      }
      case GfNode::Derivative: { auto p = simplify_with(n.a); if (!p) return std::nullopt; return b_.derivative(*p, n.var, n.order); }
      case GfNode::TaylorPolynomial: {
        auto p = simplify_with(n.a);
        if (!p) return std::nullopt;
        return b_.taylor_polynomial_terms(*p, n.var, n.orders);
      }
      case GfNode::TaylorCoeffAtZero: {
        auto p = simplify_with(n.a);
        if (!p) return std::nullopt;
        Poly res = b_.coefficients_of_term(*p, n.var, n.order);
        if (n.var + 1 == b_.num_vars(res)) res = b_.remove_last_variable(res);
        return res;
      }
      case GfNode::TaylorCoeff: { auto p = simplify_with(n.a); if (!p) return std::nullopt; return b_.taylor_expansion_of_coeff(*p, n.var, n.order); }
      case GfNode::ShiftTaylorAtZero: { auto p = simplify_with(n.a); if (!p) return std::nullopt; return b_.shift_down(*p, n.var, n.order); }
    }
    return std::nullopt;
  }

  Poly eval_with(const GenFun& g, const std::vector<S>& inputs, size_t degree_p1) {
    const bool shared = g.use_count() > 1;
    if (shared) {
      auto it = eval_cache_.find(g.get());
      if (it != eval_cache_.end() && it->second.degree_p1 == degree_p1 && same_inputs(it->second.inputs, inputs)) {
        cache_hits++;
        return it->second.output;
      }
    }
    nodes_evaluated++;
    Poly r = eval_node(g, inputs, degree_p1);
    if (shared) eval_cache_[g.get()] = EvalEntry{g, inputs, degree_p1, r};
    return r;
  }

  Poly eval_node(const GenFun& g, const std::vector<S>& inputs, size_t degree_p1) {
    const GfNode& n = *g;
    switch (n.kind) {
      case GfNode::Var: return b_.var(n.var, inputs.at(n.var), degree_p1);
      case GfNode::Const: return b_.from_scalar(scalar_of<S>(Num(n.value, n.bounds)));
      case GfNode::Add: { Poly x = eval_with(n.a, inputs, degree_p1); Poly y = eval_with(n.b, inputs, degree_p1); return b_.add(x, y); }
      case GfNode::Neg: return b_.neg(eval_with(n.a, inputs, degree_p1));
      case GfNode::Mul: { Poly x = eval_with(n.a, inputs, degree_p1); Poly y = eval_with(n.b, inputs, degree_p1); return b_.mul(x, y); }
      case GfNode::Div: { Poly x = eval_with(n.a, inputs, degree_p1); Poly y = eval_with(n.b, inputs, degree_p1); return b_.div(x, y); }
      case GfNode::Polynomial: {  // :567-584
        const HostPoly& hp = *n.poly;
        Poly t = b_.new_poly(hp.shape, std::vector<uint64_t>(hp.shape.size(), UNBOUNDED), hp.data.data());
        for (size_t v = 0; v < inputs.size(); v++) t = b_.subst_var(t, v, b_.var(v, inputs[v], degree_p1));
        size_t ndim = b_.num_vars(t);
        if (ndim > inputs.size()) {
          GFE_ASSERT(ndim == inputs.size() + 1, "polynomial with too many variables");
          t = b_.remove_last_variable(t);   // auxiliary variable of an `x ~ D` event
        }
        return b_.truncate_to_degree_p1(b_.extend_to_dim(t, inputs.size(), degree_p1), degree_p1);
      }
      case GfNode::Exp: return b_.exp(eval_with(n.a, inputs, degree_p1));
      case GfNode::Log: return b_.log(eval_with(n.a, inputs, degree_p1));
      case GfNode::Max: {
        Poly s = eval_with(n.a, inputs, degree_p1);
        Poly t = eval_with(n.b, inputs, degree_p1);
        S x = b_.constant_term(s), y = b_.constant_term(t);
        return b_.from_scalar(B::scalar_max(x, y));   // F64::max (number/f64.rs:77-84)
      }
      case GfNode::Pow: return b_.pow(eval_with(n.a, inputs, degree_p1), n.n);
      case GfNode::UniformMgf: {  // (e^x - 1) / x, :597-610
        Poly x = eval_with(n.a, inputs, degree_p1);
        if (b_.constant_term(x) == S(0.0)) {
          Poly y = b_.var_at_zero(0, degree_p1 + 1);
          Poly numerator = b_.sub(b_.exp(y), b_.from_scalar(S(1.0)));
          std::vector<S> arr = b_.to_host(numerator);          // 1-D, length degree_p1 + 1
          std::vector<uint64_t> shape = b_.array_shape(numerator);
          GFE_ASSERT(shape.size() == 1 && shape[0] >= 1, "unexpected shape in UniformMgf");
          std::vector<uint64_t> nshape{shape[0] - 1};
          GFE_ASSERT(nshape[0] >= 1, "UniformMgf needs degree_p1 >= 1");
          Poly fraction = b_.new_poly(nshape, std::vector<uint64_t>{(uint64_t)degree_p1}, arr.data() + 1);   // divide by y
          return b_.subst_var(fraction, 0, x);
        }
        Poly numerator = b_.sub(b_.exp(x), b_.from_scalar(S(1.0)));
        return b_.truncate_to_degree_p1(b_.div(numerator, x), degree_p1);
      }
      case GfNode::Subst: {  // :611-629
        std::vector<S> new_inputs = inputs;
        Poly subst = eval_with(n.b, inputs, degree_p1);
        S c = b_.constant_term(subst);
        subst = b_.sub(subst, b_.from_scalar(c));
        if (n.var < inputs.size()) new_inputs[n.var] = c;
        else {
          GFE_ASSERT(n.var == inputs.size(), "substituted variable out of range");
          new_inputs.push_back(c);
        }
        Poly taylor = eval_with(n.a, new_inputs, degree_p1);
        Poly result = b_.subst_var(taylor, n.var, subst);
        if (b_.num_vars(taylor) > inputs.size()) {
          GFE_ASSERT(b_.num_vars(taylor) == inputs.size() + 1, "unexpected number of variables");
          result = b_.remove_last_variable(result);
        }
        return result;
      }
      case GfNode::Derivative: {
        Poly t = eval_with(n.a, inputs, degree_p1 + n.order);
        return b_.truncate_to_degree_p1(b_.derivative(t, n.var, n.order), degree_p1);
      }
      case GfNode::TaylorPolynomial: {
        std::vector<S> new_inputs = inputs;
        new_inputs.at(n.var) = S(0.0);
        size_t max_order = 0;
        for (size_t o : n.orders) max_order = std::max(max_order, o);
        Poly t = eval_with(n.a, new_inputs, degree_p1 + max_order);
        Poly r = b_.taylor_polynomial_terms(t, n.var, n.orders);
        r = b_.subst_var(r, n.var, b_.var(n.var, inputs[n.var], degree_p1));
        return b_.truncate_to_degree_p1(r, degree_p1);
      }
      case GfNode::TaylorCoeffAtZero: return eval_taylor_coeff_at_zero(n.a, n.var, n.order, inputs, degree_p1);
      case GfNode::TaylorCoeff: {
        Poly t = eval_with(n.a, inputs, degree_p1 + n.order);
        return b_.truncate_to_degree_p1(b_.taylor_expansion_of_coeff(t, n.var, n.order), degree_p1);
      }
      case GfNode::ShiftTaylorAtZero: {
        if (inputs.at(n.var) == S(0.0)) {
          Poly t = eval_with(n.a, inputs, degree_p1 + n.order);
          return b_.truncate_to_degree_p1(b_.shift_down(t, n.var, n.order), degree_p1);
        }
        std::vector<size_t> orders;
        for (size_t i = 0; i < n.order; i++) orders.push_back(i);
        GenFun first_terms = gf::taylor_polynomial_at_zero(n.a, n.var, orders);
        GenFun mass_on_zero = gf::substitute_var(first_terms, n.var, gf::one());
        GenFun h = gf::add(gf::div(gf::sub(n.a, first_terms), gf::pow(gf::var(n.var), (uint32_t)n.order)), mass_on_zero);
        return eval_with(h, inputs, degree_p1);
      }
    }
    throw EvalError("unreachable");
  }

  Poly eval_taylor_coeff_at_zero(const GenFun& g, Var v, size_t order, const std::vector<S>& inputs, size_t degree_p1) {  // :670-765
    gf::Recognised rec;
    if (gf::recognize_discrete_poisson_observation(g, v, &rec)) {
      // D^n(G) with D(G)(y) := lambda y G'(y), evaluated at y = e^(-lambda) y; the 1/n! is folded into the loop
      GenFun f = rec.inner;
      for (size_t k = 1; k <= order; k++)
        f = gf::mul(gf::mul(gf::derive(f, rec.param_var, 1), gf::var(rec.param_var)), gf::constant(rec.scalar / Num((double)(uint32_t)k)));
      GenFun repl = gf::mul(gf::constant(exp(-rec.scalar)), gf::var(rec.param_var));
      f = gf::substitute_var(f, rec.param_var, repl);
      return b_.truncate_to_degree_p1(eval_with(f, inputs, degree_p1), degree_p1);
    }
    if (gf::recognize_continuous_poisson_observation(g, v, &rec)) {
      GenFun f = rec.inner;
      for (size_t k = 1; k <= order; k++)
        f = gf::mul(gf::derive(f, rec.param_var, 1), gf::constant(rec.scalar / Num((double)(uint32_t)k)));
      GenFun repl = gf::sub(gf::var(rec.param_var), gf::constant(rec.scalar));
      f = gf::substitute_var(f, rec.param_var, repl);
      return b_.truncate_to_degree_p1(eval_with(f, inputs, degree_p1), degree_p1);
    }
    if (gf::recognize_negative_binomial_observation(g, v, &rec)) {
      const Num p = rec.scalar;
      std::vector<Num> lahs{Num(1.0)};   // row d of the Lah numbers times (1-p)^d / d!
      const Num one_mp = Num(1.0) - p;
      for (size_t d = 1; d <= order; d++) {
        std::vector<Num> next;
        for (size_t i = 0; i <= d; i++) {
          Num l_dm1_i = i < lahs.size() ? lahs[i] : Num(0.0);
          Num l_dm1_im1 = (1 <= i && i <= lahs.size()) ? lahs[i - 1] : Num(0.0);
          next.push_back(one_mp / Num((double)(uint32_t)d) * (l_dm1_i * Num((double)(uint32_t)(d + i - 1)) + l_dm1_im1));
        }
        lahs = next;
      }
      Poly sum = b_.zero_with(std::vector<uint64_t>(inputs.size(), degree_p1));
      std::vector<S> new_inputs = inputs;
      new_inputs.at(rec.param_var) = scalar_of<S>(p) * inputs[rec.param_var];
      Poly inner_result = eval_with(rec.inner, new_inputs, degree_p1 + order);
      Poly power = b_.from_scalar(S(1.0));
      Poly param_tp = b_.var(rec.param_var, inputs[rec.param_var], degree_p1);
      Poly p_param = b_.mul(b_.from_scalar(scalar_of<S>(p)), param_tp);
      for (const Num& lah : lahs) {
        Poly subst = b_.mul(b_.from_scalar(scalar_of<S>(p)), b_.var_at_zero(rec.param_var, degree_p1));
        Poly term = b_.mul(b_.mul(b_.subst_var(inner_result, rec.param_var, subst), power), b_.from_scalar(scalar_of<S>(lah)));
        sum = b_.add(sum, term);
        power = b_.mul(power, p_param);
        inner_result = b_.derivative(inner_result, rec.param_var, 1);
      }
      return b_.truncate_to_degree_p1(sum, degree_p1);
    }
    std::vector<S> in = inputs;
    Poly result;
    if (v == in.size()) {
      in.push_back(S(0.0));
      Poly t = eval_with(g, in, degree_p1 + order);
      result = b_.remove_last_variable(b_.coefficients_of_term(t, v, order));
    } else {
      in.at(v) = S(0.0);
      Poly t = eval_with(g, in, degree_p1 + order);
      result = b_.coefficients_of_term(t, v, order);
    }
    return b_.truncate_to_degree_p1(result, degree_p1);
  }
};

}  // namespace gfe
