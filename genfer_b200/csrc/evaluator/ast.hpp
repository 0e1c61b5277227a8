// Host evaluator -- program representation.
// Restates the data types of the reference's src/ppl.rs (Natural :11-12, PosRatio :33-72, Var :96-104,
// Distribution :177-206, Comparison :300-305, Event :317-324, Statement :446-478, Program :666-670) in C++.
// This is host control logic: no Taylor arithmetic happens here.
#pragma once
#include <cstdint>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "num.hpp"

namespace gfe {

struct EvalError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
#define GFE_ASSERT(cond, msg)                 \
  do {                                        \
    if (!(cond)) throw ::gfe::EvalError(msg); \
  } while (0)

using Var = size_t;
using Natural = uint32_t;

struct PosRatio {  // ppl.rs:33-72
  uint64_t numer = 0, denom = 1;
  PosRatio() = default;
  PosRatio(uint64_t n, uint64_t d) : numer(n), denom(d) {}
  PosRatio complement() const {
    GFE_ASSERT(numer <= denom, "probability > 1");
    return {denom - numer, denom};
  }
  std::optional<uint32_t> as_integer() const {
    if (denom != 0 && numer % denom == 0 && numer / denom <= UINT32_MAX) return (uint32_t)(numer / denom);
    return std::nullopt;
  }
  // Number::from_ratio for F64 (number/f64.rs:49-51): one IEEE division
  double to_f64() const { return (double)numer / (double)denom; }
  // T::from_ratio under T = F64 and T = Interval<F64> at once (num.hpp)
  Num to_num() const { return Num::from_ratio(numer, denom); }
};

enum class DistKind {
  Dirac, Bernoulli, BernoulliVarProb, BinomialVarTrials, Binomial, Categorical, NegBinomialVarSuccesses, NegBinomial,
  Geometric, Poisson, PoissonVarRate, Uniform, Exponential, Gamma, UniformCont
};

struct Distribution {  // ppl.rs:177-206
  DistKind kind;
  PosRatio p;              // Dirac value / probability / rate / lambda / shape(Gamma) / start(UniformCont)
  PosRatio q;              // rate(Gamma) / end(UniformCont)
  Var var = 0;             // parameter variable of the *Var* kinds
  Natural n = 0, m = 0;    // Binomial / NegBinomial count; Uniform start (n) and end (m)
  std::vector<PosRatio> rs;  // Categorical
  bool uses_var() const {
    return kind == DistKind::BernoulliVarProb || kind == DistKind::BinomialVarTrials ||
           kind == DistKind::NegBinomialVarSuccesses || kind == DistKind::PoissonVarRate;
  }
};

enum class Comparison { Eq, Lt, Le };

struct Event;
using EventP = std::shared_ptr<const Event>;
struct Event {  // ppl.rs:317-324
  enum Kind { InSet, VarComparison, DataFromDist, Complement, Intersection } kind;
  Var v1 = 0, v2 = 0;
  std::vector<Natural> set;
  Comparison comp = Comparison::Eq;
  Natural data = 0;
  Distribution dist{};
  std::vector<EventP> children;  // Complement: 1 child; Intersection: n children

  static EventP in_set(Var v, std::vector<Natural> s) {
    auto e = std::make_shared<Event>();
    e->kind = InSet; e->v1 = v; e->set = std::move(s);
    return e;
  }
  static EventP var_comparison(Var a, Comparison c, Var b) {
    auto e = std::make_shared<Event>();
    e->kind = VarComparison; e->v1 = a; e->v2 = b; e->comp = c;
    return e;
  }
  static EventP data_from_dist(Natural d, Distribution dist) {
    auto e = std::make_shared<Event>();
    e->kind = DataFromDist; e->data = d; e->dist = std::move(dist);
    return e;
  }
  static EventP complement(EventP inner) {  // :372-378: double complement cancels
    if (inner->kind == Complement) return inner->children[0];
    auto e = std::make_shared<Event>();
    e->kind = Complement; e->children = {std::move(inner)};
    return e;
  }
  static EventP intersection(std::vector<EventP> es) {  // :399-413: flatten; a single conjunct is itself
    std::vector<EventP> conj;
    for (auto& e : es) {
      if (e->kind == Intersection) conj.insert(conj.end(), e->children.begin(), e->children.end());
      else conj.push_back(e);
    }
    if (conj.size() == 1) return conj[0];
    auto e = std::make_shared<Event>();
    e->kind = Intersection; e->children = std::move(conj);
    return e;
  }
  static EventP disjunction(std::vector<EventP> es) {  // :415-421: De Morgan
    if (es.size() == 1) return es[0];
    std::vector<EventP> neg;
    for (auto& e : es) neg.push_back(complement(e));
    return complement(intersection(std::move(neg)));
  }
  static EventP always() { return intersection({}); }
  static EventP never() { return complement(always()); }

  size_t used_vars() const;  // VarRange upper bound (max var id + 1), :327-337
};

struct Statement;
using Block = std::vector<Statement>;
struct Statement {  // ppl.rs:446-478
  enum Kind { Sample, Assign, Decrement, IfThenElse, While, Fail, Normalize } kind = Fail;
  Var var = 0;
  Distribution dist{};
  bool add_previous_value = false;
  bool has_addend = false;
  Natural addend_factor = 1;
  Var addend_var = 0;
  Natural offset = 0;
  EventP cond;
  Block then_, else_;         // IfThenElse; While/Normalize body in then_
  std::optional<size_t> unroll;
  std::vector<Var> given_vars;

  bool uses_observe() const;  // :613-624
  size_t used_vars() const;   // :626-661
};

struct Program {
  Block stmts;
  Var result = 0;
  std::vector<std::string> var_names;
  bool uses_observe() const {
    for (auto& s : stmts)
      if (s.uses_observe()) return true;
    return false;
  }
  size_t used_vars() const {
    size_t n = 0;
    for (auto& s : stmts) n = std::max(n, s.used_vars());
    return n;
  }
};

inline size_t dist_used_vars(const Distribution& d) { return d.uses_var() ? d.var + 1 : 0; }

inline size_t Event::used_vars() const {
  switch (kind) {
    case InSet: return v1 + 1;
    case VarComparison: return std::max(v1, v2) + 1;
    case DataFromDist: return dist_used_vars(dist);
    default: {
      size_t n = 0;
      for (auto& c : children) n = std::max(n, c->used_vars());
      return n;
    }
  }
}
inline bool Statement::uses_observe() const {
  switch (kind) {
    case Sample: case Assign: case Decrement: return false;
    case Fail: return true;
    default:
      for (auto& s : then_) if (s.uses_observe()) return true;
      for (auto& s : else_) if (s.uses_observe()) return true;
      return false;
  }
}
inline size_t Statement::used_vars() const {
  size_t n = 0;
  switch (kind) {
    case Sample: return std::max(dist_used_vars(dist), var + 1);
    case Assign: return std::max(var + 1, has_addend ? addend_var + 1 : (size_t)0);
    case Decrement: return var + 1;
    case Fail: return 0;
    case IfThenElse: case While: n = cond->used_vars(); break;
    case Normalize: break;
  }
  for (auto& s : then_) n = std::max(n, s.used_vars());
  for (auto& s : else_) n = std::max(n, s.used_vars());
  return n;
}

}  // namespace gfe
