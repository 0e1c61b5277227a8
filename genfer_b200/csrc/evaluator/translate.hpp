// Host evaluator -- program -> generating function.
// Restates the reference's GfTransformer (src/semantics/gf.rs): GfTranslation :11-63, events :100-205, statements
// :207-349, compound distributions :358-386, distributions :389-536, observations from distributions :538-586,
// normalize :588-634, marginalisation :637-657.  Compiler pass: builds the GenFun DAG, no arithmetic.
#pragma once
#include "genfun.hpp"
#include "support.hpp"

namespace gfe {

struct GfTranslation {
  VarSupport var_info;
  GenFun gf;
  GenFun rest;          // remaining probability mass not captured in gf
  VarSupport rest_info;

  static GfTranslation zero(size_t n) { return {VarSupport::empty(n), gf::zero(), gf::zero(), VarSupport::empty(n)}; }
  GfTranslation join(const GfTranslation& o) const {  // :36-43: branches of an if -- max of the rests
    return {var_info.join(o.var_info), gf::add(gf, o.gf), gf::max(rest, o.rest), rest_info.join(o.rest_info)};
  }
  GfTranslation plus(const GfTranslation& o) const {  // :46-57
    return {var_info.join(o.var_info), gf::add(gf, o.gf), gf::add(rest, o.rest), rest_info.join(o.rest_info)};
  }
  void scale(const Num& c) {  // :59-64
    gf = gf::mul(gf, gf::constant(c));
    rest = gf::mul(rest, gf::constant(c));
  }
};

inline GenFun marginalize_out(Var v, const GenFun& g, const VarSupport& vi) {  // :637-649
  if (v >= vi.num_vars()) {
    GFE_ASSERT(v == vi.num_vars(), "temporary variable index");
    return gf::substitute_var(g, v, gf::one());
  }
  return gf::substitute_var(g, v, vi[v].is_discrete() ? gf::one() : gf::zero());
}
inline GenFun marginalize_all(GenFun g, const VarSupport& vi) {  // :651-657
  for (size_t v = 0; v < vi.num_vars(); v++) g = marginalize_out(v, g, vi);
  return g;
}

// Event::recognize_const_prob (ppl.rs:339-365)
inline std::optional<Num> recognize_const_prob(const Event& e) {
  switch (e.kind) {
    case Event::InSet: case Event::VarComparison: return std::nullopt;
    case Event::DataFromDist:
      if (e.dist.kind == DistKind::Bernoulli) {
        if (e.data == 0) return e.dist.p.complement().to_num();
        if (e.data == 1) return e.dist.p.to_num();
        return Num(0.0);
      }
      return std::nullopt;
    case Event::Complement: {
      auto p = recognize_const_prob(*e.children[0]);
      if (!p) return std::nullopt;
      return Num(1.0) - *p;
    }
    case Event::Intersection: {
      Num r(1.0);
      for (auto& c : e.children) {
        auto p = recognize_const_prob(*c);
        if (!p) return std::nullopt;
        r = r * *p;
      }
      return r;
    }
  }
  return std::nullopt;
}

class GfTransformer {
 public:
  explicit GfTransformer(size_t unroll = 0) : unroll_(unroll) { support_.unroll = unroll; }

  std::string warnings;   // what the reference println!s while translating (stdout)

  GfTranslation semantics(const Program& p) {
    VarSupport vi = support_.init(p);
    GfTranslation t{vi, gf::one(), gf::zero(), VarSupport::empty(vi.num_vars())};
    return transform_statements(p.stmts, t);
  }

 private:
  size_t unroll_;
  SupportTransformer support_;

  static GenFun gf_in_set(Var v, const std::vector<Natural>& set, const GenFun& g) {  // :105-112
    if (set.size() == 1) return gf::mul(gf::taylor_coeff_at_zero(g, v, set[0]), gf::pow(gf::var(v), set[0]));
    std::vector<size_t> orders(set.begin(), set.end());
    return gf::taylor_polynomial_at_zero(g, v, std::move(orders));
  }
  static std::vector<Natural> upto(Natural n, bool inclusive) {
    std::vector<Natural> v;
    for (Natural i = 0; inclusive ? i <= n : i < n; i++) v.push_back(i);
    return v;
  }

  std::pair<GfTranslation, GfTranslation> transform_event(const Event& e, const GfTranslation& init) {  // :100-205
    const VarSupport& vi = init.var_info;
    GenFun g = init.gf;
    switch (e.kind) {
      case Event::InSet: g = gf_in_set(e.v1, e.set, g); break;
      case Event::VarComparison: {
        auto r1 = vi[e.v1].finite_nonempty_range(), r2 = vi[e.v2].finite_nonempty_range();
        GFE_ASSERT(r1 || r2, "Cannot compare two variables with infinite support.");
        Var scrutinee, other;
        bool reversed;
        std::pair<uint32_t, uint32_t> range;
        if (!r1) { scrutinee = e.v2; other = e.v1; reversed = false; range = *r2; }
        else if (!r2) { scrutinee = e.v1; other = e.v2; reversed = true; range = *r1; }
        else if (r1->second - r1->first <= r2->second - r2->first) { scrutinee = e.v1; other = e.v2; reversed = true; range = *r1; }
        else { scrutinee = e.v2; other = e.v1; reversed = false; range = *r2; }
        GenFun result = gf::zero();
        for (uint64_t i = range.first; i <= range.second; i++) {
          GenFun eq_i = gf_in_set(scrutinee, {(Natural)i}, g);
          GenFun summand;
          if (e.comp == Comparison::Eq) summand = gf_in_set(other, {(Natural)i}, eq_i);
          else if (e.comp == Comparison::Lt && !reversed) summand = gf_in_set(other, upto((Natural)i, false), eq_i);
          else if (e.comp == Comparison::Lt && reversed) summand = gf::sub(eq_i, gf_in_set(other, upto((Natural)i, true), eq_i));
          else if (e.comp == Comparison::Le && !reversed) summand = gf_in_set(other, upto((Natural)i, true), eq_i);
          else summand = gf::sub(eq_i, gf_in_set(other, upto((Natural)i, false), eq_i));
          result = gf::add(result, summand);
        }
        g = result;
        break;
      }
      case Event::DataFromDist:
        if (auto f = recognize_const_prob(e)) g = gf::mul(gf::constant(*f), g);
        else g = transform_data_from_dist(e.data, e.dist, vi, g);
        break;
      case Event::Complement: g = transform_event(*e.children[0], init).second.gf; break;
      case Event::Intersection: {
        GfTranslation then_r = init;
        for (auto& c : e.children) then_r = transform_event(*c, then_r).first;
        g = then_r.gf;
        break;
      }
    }
    auto info = support_.transform_event(e, init.var_info);
    auto rinfo = support_.transform_event(e, init.rest_info);
    return {GfTranslation{info.first, g, init.rest, rinfo.first},
            GfTranslation{info.second, gf::sub(init.gf, g), init.rest, rinfo.second}};
  }

  GfTranslation transform_statements(const Block& b, GfTranslation cur) {
    for (auto& s : b) cur = transform_statement(s, cur);
    return cur;
  }

  GfTranslation transform_statement(const Statement& st, const GfTranslation& init) {  // :207-349
    switch (st.kind) {
      case Statement::Sample: return transform_distribution(st.dist, st.var, init, st.add_previous_value);
      case Statement::Assign: {
        const Var v = st.var;
        GenFun g = init.gf;
        const VarSupport& vi = init.var_info;
        GenFun var = gf::var(v);
        uint32_t v_exp = st.add_previous_value ? 1 : 0;
        bool has_w = false;
        Var w = 0;
        GenFun w_subst;
        if (st.has_addend) {
          if (v == st.addend_var) {
            v_exp += st.addend_factor;
          } else if (vi[st.addend_var].is_discrete()) {
            has_w = true; w = st.addend_var;
            w_subst = gf::mul(gf::var(w), gf::pow(var, st.addend_factor));
          } else {
            GFE_ASSERT(!vi[v].is_discrete() || !st.add_previous_value, "cannot add a continuous to a discrete variable");
            has_w = true; w = st.addend_var;
            w_subst = gf::add(gf::var(w), gf::mul(var, gf::from_u32(st.addend_factor)));
          }
        }
        if (vi[v].is_discrete()) g = gf::substitute_var(g, v, gf::pow(var, v_exp));
        else g = gf::substitute_var(g, v, gf::mul(var, gf::from_u32(v_exp)));
        if (has_w) g = gf::substitute_var(g, w, w_subst);
        VarSupport nvi = support_.transform_statement(st, init.var_info);
        VarSupport nri = support_.transform_statement(st, init.rest_info);
        if (nvi[v].is_discrete()) g = gf::mul(g, gf::pow(var, st.offset));
        else g = gf::mul(g, gf::exp(gf::mul(var, gf::from_u32(st.offset))));
        return {nvi, g, init.rest, nri};
      }
      case Statement::Decrement: {
        GFE_ASSERT(init.var_info[st.var].is_discrete(), "cannot decrement continuous variables");
        VarSupport nvi = support_.transform_statement(st, init.var_info);
        VarSupport nri = support_.transform_statement(st, init.rest_info);
        return {nvi, gf::shift_down_taylor_at_zero(init.gf, st.var, st.offset), init.rest, nri};
      }
      case Statement::IfThenElse: {
        if (auto f = recognize_const_prob(*st.cond)) {  // avoids path explosion: scale AFTER both branches
          GfTranslation t = transform_statements(st.then_, init);
          GfTranslation e = transform_statements(st.else_, init);
          t.scale(*f);
          e.scale(Num(1.0) - *f);
          return t.plus(e);
        }
        auto br = transform_event(*st.cond, init);
        GfTranslation t = transform_statements(st.then_, br.first);
        GfTranslation e = transform_statements(st.else_, br.second);
        return t.join(e);
      }
      case Statement::While: {  // experimental in the reference too (:304-321)
        warnings += "WARNING: results are APPROXIMATE due to presence of loops: exact inference is only possible for loop-free programs\n";
        GfTranslation result = GfTranslation::zero(init.var_info.num_vars());
        GfTranslation rest = init;
        size_t iters = st.unroll.value_or(unroll_);
        for (size_t i = 0; i < iters; i++) {
          auto br = transform_event(*st.cond, rest);
          result = result.join(br.second);
          rest = transform_statements(st.then_, br.first);
        }
        result.rest = gf::add(result.rest, marginalize_all(rest.gf, rest.var_info));
        VarSupport inv = support_.find_while_invariant(*st.cond, st.then_, rest.var_info);
        auto ex = support_.transform_event(*st.cond, inv);
        result.rest_info = result.rest_info.join(ex.second);
        result.var_info = result.var_info.join(result.rest_info);
        return result;
      }
      case Statement::Fail: return GfTranslation::zero(init.var_info.num_vars());
      case Statement::Normalize: return transform_normalize(st.given_vars, 0, st.then_, init);
    }
    throw EvalError("unreachable");
  }

  static GenFun compound_dist(const GenFun& g, const GenFun& base, Var sampled, Var param, bool add_previous,
                              bool param_discrete, const GenFun& subst) {  // :358-386
    auto combined = [&]() { return param_discrete ? gf::mul(gf::var(param), subst) : gf::add(gf::var(param), subst); };
    if (sampled == param) {
      if (add_previous) return gf::substitute_var(g, param, combined());
      return gf::substitute_var(g, param, subst);
    }
    return gf::substitute_var(base, param, combined());
  }

  static GenFun geometric_gf(const PosRatio& p, Var v) {
    return gf::div(gf::from_ratio(p), gf::sub(gf::one(), gf::mul(gf::from_ratio(p.complement()), gf::var(v))));
  }

  GfTranslation transform_distribution(const Distribution& d, Var v, const GfTranslation& tr, bool add_previous) {  // :389-536
    GenFun base = add_previous ? tr.gf : marginalize_out(v, tr.gf, tr.var_info);
    VarSupport nvi = SupportTransformer::transform_distribution(d, v, tr.var_info, add_previous);
    VarSupport nri = SupportTransformer::transform_distribution(d, v, tr.rest_info, add_previous);
    const GenFun& g = tr.gf;
    GenFun out;
    switch (d.kind) {
      case DistKind::Dirac: {
        GenFun dirac;
        if (auto a = d.p.as_integer()) dirac = gf::pow(gf::var(v), *a);
        else dirac = gf::exp(gf::mul(gf::var(v), gf::from_ratio(d.p)));
        out = gf::mul(dirac, base);
        break;
      }
      case DistKind::Bernoulli:
        out = gf::mul(gf::add(gf::mul(gf::from_ratio(d.p), gf::var(v)), gf::from_ratio(d.p.complement())), base);
        break;
      case DistKind::BernoulliVarProb: {
        Var w = d.var;
        GenFun prob_times_gf = tr.var_info[w].is_discrete() ? gf::mul(gf::derive(g, w, 1), gf::var(w)) : gf::derive(g, w, 1);
        GenFun prob_times_base = add_previous ? prob_times_gf : marginalize_out(v, prob_times_gf, tr.var_info);
        GenFun v_term = nvi[v].is_discrete() ? gf::var(v) : gf::exp(gf::var(v));
        out = gf::add(base, gf::mul(gf::sub(v_term, gf::one()), prob_times_base));
        break;
      }
      case DistKind::BinomialVarTrials: {
        GenFun subst = gf::add(gf::mul(gf::from_ratio(d.p), gf::var(v)), gf::from_ratio(d.p.complement()));
        out = compound_dist(g, base, v, d.var, add_previous, true, subst);
        break;
      }
      case DistKind::Binomial:
        out = gf::mul(gf::pow(gf::add(gf::mul(gf::from_ratio(d.p), gf::var(v)), gf::from_ratio(d.p.complement())), d.n), base);
        break;
      case DistKind::Categorical: {
        GenFun cat = gf::zero();
        for (auto it = d.rs.rbegin(); it != d.rs.rend(); ++it) {
          cat = gf::mul(cat, gf::var(v));
          cat = gf::add(cat, gf::from_ratio(*it));
        }
        out = gf::mul(cat, base);
        break;
      }
      case DistKind::NegBinomialVarSuccesses:
        out = compound_dist(g, base, v, d.var, add_previous, true, geometric_gf(d.p, v));
        break;
      case DistKind::NegBinomial: out = gf::mul(gf::pow(geometric_gf(d.p, v), d.n), base); break;
      case DistKind::Geometric: out = gf::mul(geometric_gf(d.p, v), base); break;
      case DistKind::Poisson:
        out = gf::mul(gf::exp(gf::mul(gf::from_ratio(d.p), gf::sub(gf::var(v), gf::one()))), base);
        break;
      case DistKind::PoissonVarRate: {
        bool wd = tr.var_info[d.var].is_discrete();
        GenFun inner = gf::mul(gf::from_ratio(d.p), gf::sub(gf::var(v), gf::one()));
        GenFun subst = wd ? gf::exp(inner) : inner;
        out = compound_dist(g, base, v, d.var, add_previous, wd, subst);
        break;
      }
      case DistKind::Uniform: {
        GFE_ASSERT(d.m > d.n, "Uniform distribution cannot have length 0");
        uint32_t len = d.m - d.n;
        GenFun weight = gf::from_ratio(PosRatio(1, len));
        GenFun uni = gf::zero();
        for (uint32_t i = 0; i < len; i++) uni = gf::add(weight, gf::mul(gf::var(v), uni));
        uni = gf::mul(uni, gf::pow(gf::var(v), d.n));
        out = gf::mul(uni, base);
        break;
      }
      case DistKind::Exponential: {
        GenFun beta = gf::from_ratio(d.p);
        out = gf::mul(gf::div(beta, gf::sub(beta, gf::var(v))), base);
        break;
      }
      case DistKind::Gamma: {
        GenFun beta = gf::from_ratio(d.q);
        GenFun gamma;
        if (auto shape = d.p.as_integer()) gamma = gf::pow(gf::div(beta, gf::sub(beta, gf::var(v))), *shape);
        else gamma = gf::exp(gf::mul(gf::from_ratio(d.p), gf::sub(gf::log(beta), gf::log(gf::sub(beta, gf::var(v))))));
        out = gf::mul(gamma, base);
        break;
      }
      case DistKind::UniformCont: {
        const Num width = d.q.to_num() - d.p.to_num();
        GenFun x = gf::mul(gf::constant(width), gf::var(v));
        GenFun uni = gf::mul(gf::uniform_mgf(x), gf::exp(gf::mul(gf::from_ratio(d.p), gf::var(v))));
        out = gf::mul(uni, base);
        break;
      }
    }
    return {nvi, out, tr.rest, nri};
  }

  GenFun transform_data_from_dist(Natural data, const Distribution& d, const VarSupport& vi, const GenFun& g) {  // :538-586
    if (d.kind == DistKind::BernoulliVarProb) {
      GenFun ptg = vi[d.var].is_discrete() ? gf::mul(gf::derive(g, d.var, 1), gf::var(d.var)) : gf::derive(g, d.var, 1);
      if (data == 0) return gf::sub(g, ptg);
      if (data == 1) return ptg;
      return gf::zero();
    }
    if (d.kind == DistKind::BinomialVarTrials) {
      GenFun repl = gf::mul(gf::from_ratio(d.p.complement()), gf::var(d.var));
      return gf::mul(gf::substitute_var(gf::taylor_coeff(g, d.var, data), d.var, repl),
                     gf::pow(gf::mul(gf::from_ratio(d.p), gf::var(d.var)), data));
    }
    // general case: sample a temporary variable X_n ~ D(...), take its data-th coefficient, marginalise it out
    Var new_var = gf::used_vars(g);
    Statement sample;
    sample.kind = Statement::Sample;
    sample.var = new_var;
    sample.dist = d;
    sample.add_previous_value = false;
    GfTranslation tr{vi, g, gf::zero(), VarSupport::empty(vi.num_vars())};
    GfTranslation nt = transform_statement(sample, tr);
    GenFun c = gf::taylor_coeff_at_zero(nt.gf, new_var, data);
    return marginalize_out(new_var, c, nt.var_info);
  }

  GfTranslation transform_normalize(const std::vector<Var>& given, size_t idx, const Block& block, const GfTranslation& tr) {  // :588-634
    if (idx == given.size()) {
      GenFun total_before = marginalize_all(tr.gf, tr.var_info);
      GenFun rest_before = tr.rest;
      GfTranslation t = transform_statements(block, tr);
      GenFun total_after = marginalize_all(t.gf, t.var_info);
      GenFun min_factor = gf::div(total_before, gf::add(total_after, t.rest));
      GenFun max_factor = gf::div(gf::add(total_before, rest_before), total_after);
      return {t.var_info, gf::mul(min_factor, t.gf), gf::mul(max_factor, t.rest), t.rest_info};
    }
    Var v = given[idx];
    auto range = tr.var_info[v].finite_nonempty_range();
    GFE_ASSERT(range.has_value(), "Cannot normalize with respect to a variable whose value could not be proven to be bounded.");
    GfTranslation joined = GfTranslation::zero(tr.var_info.num_vars());
    for (uint64_t i = range->first; i <= range->second; i++) {
      GenFun summand = gf::mul(gf::taylor_coeff_at_zero(tr.gf, v, i), gf::pow(gf::var(v), (uint32_t)i));
      VarSupport vi = tr.var_info, ri = tr.rest_info;
      vi.set(v, SupportSet::point((uint32_t)i));
      ri.set(v, SupportSet::point((uint32_t)i));
      joined = joined.join(transform_normalize(given, idx + 1, block, GfTranslation{vi, summand, tr.rest, ri}));
    }
    return joined;
  }
};

}  // namespace gfe
