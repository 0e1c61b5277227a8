// Host evaluator -- support (range) analysis.  All of this is INTEGER / exact-rational work and
// must be bit-exact: it decides discrete-vs-continuous (expansion point 1 vs 0), the
// ranges of variable comparisons / normalize, and the automatic probability limit.
// Restates the reference's src/support.rs (SupportSet lattice :13-17, join :61-118, saturating_sub :120-134,
// retain_only / remove_all :196-263, Add :378-441, Mul :451-519) and src/semantics/support.rs (VarSupport :9-143,
// SupportTransformer :159-386 incl. loop widening :332-350).
#pragma once
#include <algorithm>
#include <numeric>

#include "ast.hpp"

namespace gfe {

// Exact non-negative rational with +infinity (den == 0); enough for interval supports of the distributions.
struct Rat {
  __int128 num = 0, den = 1;
  Rat() = default;
  Rat(__int128 n, __int128 d) : num(n), den(d) { normalize(); }
  static Rat from_u32(uint32_t x) { return Rat(x, 1); }
  static Rat infinity() { Rat r; r.num = 1; r.den = 0; return r; }
  bool is_inf() const { return den == 0; }
  void normalize() {
    if (den == 0) { num = 1; return; }
    __int128 a = num < 0 ? -num : num, b = den;
    while (b) { __int128 t = a % b; a = b; b = t; }
    if (a > 1) { num /= a; den /= a; }
  }
  int cmp(const Rat& o) const {
    if (is_inf() || o.is_inf()) return is_inf() == o.is_inf() ? 0 : (is_inf() ? 1 : -1);
    __int128 l = num * o.den, r = o.num * den;
    return l < r ? -1 : (l > r ? 1 : 0);
  }
  bool operator<(const Rat& o) const { return cmp(o) < 0; }
  bool operator<=(const Rat& o) const { return cmp(o) <= 0; }
  bool operator>(const Rat& o) const { return cmp(o) > 0; }
  bool operator>=(const Rat& o) const { return cmp(o) >= 0; }
  bool operator==(const Rat& o) const { return cmp(o) == 0; }
  Rat min(const Rat& o) const { return *this <= o ? *this : o; }
  Rat max(const Rat& o) const { return *this >= o ? *this : o; }
  Rat operator+(const Rat& o) const {
    if (is_inf() || o.is_inf()) return infinity();
    return Rat(num * o.den + o.num * den, den * o.den);
  }
  Rat operator-(const Rat& o) const {
    if (is_inf()) return infinity();
    return Rat(num * o.den - o.num * den, den * o.den);
  }
  Rat operator*(const Rat& o) const {
    if (is_inf() || o.is_inf()) return infinity();   // (0 * inf never arises: zero supports are handled first)
    return Rat(num * o.num, den * o.den);
  }
  std::string str() const {
    if (is_inf()) return "\xe2\x88\x9e";
    auto i128 = [](__int128 v) {
      if (v == 0) return std::string("0");
      bool neg = v < 0;
      if (neg) v = -v;
      std::string s;
      while (v) { s.insert(s.begin(), char('0' + (int)(v % 10))); v /= 10; }
      return neg ? "-" + s : s;
    };
    return den == 1 ? i128(num) : i128(num) + "/" + i128(den);
  }
  double to_double() const { return is_inf() ? INFINITY : (double)num / (double)den; }
};

struct SupportSet {  // support.rs:13-17
  enum Kind { Empty, Range, Interval } kind = Empty;
  uint32_t start = 0;
  std::optional<uint32_t> end;   // Range: None = unbounded
  Rat lo, hi;                    // Interval

  static SupportSet empty() { return {}; }
  static SupportSet range(uint32_t s, std::optional<uint32_t> e) {
    SupportSet r; r.kind = Range; r.start = s; r.end = e; return r;
  }
  static SupportSet zero() { return range(0, 0); }
  static SupportSet point(uint32_t x) { return range(x, x); }
  static SupportSet naturals() { return range(0, std::nullopt); }
  static SupportSet from_range_excl(uint32_t s, uint32_t e) { return e <= s ? empty() : range(s, e - 1); }  // :322-333
  static SupportSet from_range_incl(uint32_t s, uint32_t e) { return s > e ? empty() : range(s, e); }       // :348-359
  static SupportSet interval(Rat s, Rat e) {  // :154-159
    if (s > e) return empty();
    SupportSet r; r.kind = Interval; r.lo = s; r.hi = e; return r;
  }
  static SupportSet nonneg_reals() { return interval(Rat(), Rat::infinity()); }

  bool is_empty() const { return kind == Empty; }
  bool is_zero() const { return kind == Range && start == 0 && end && *end == 0; }
  bool is_discrete() const { return kind != Interval; }   // :146-151
  bool operator==(const SupportSet& o) const {
    if (kind != o.kind) return false;
    if (kind == Range) return start == o.start && end == o.end;
    if (kind == Interval) return lo == o.lo && hi == o.hi;
    return true;
  }
  std::optional<std::pair<uint32_t, uint32_t>> finite_nonempty_range() const {  // :136-141
    if (kind == Range && end) return std::make_pair(start, *end);
    return std::nullopt;
  }
  SupportSet join(const SupportSet& o) const {  // :61-118
    if (is_empty()) return o;
    if (o.is_empty()) return *this;
    if (kind == Range && o.kind == Range)
      return range(std::min(start, o.start), (end && o.end) ? std::optional<uint32_t>(std::max(*end, *o.end)) : std::nullopt);
    auto as_iv = [](const SupportSet& s) {
      if (s.kind == Interval) return std::make_pair(s.lo, s.hi);
      return std::make_pair(Rat::from_u32(s.start), s.end ? Rat::from_u32(*s.end) : Rat::infinity());
    };
    auto a = as_iv(*this), b = as_iv(o);
    SupportSet r; r.kind = Interval; r.lo = a.first.min(b.first); r.hi = a.second.max(b.second);
    return r;
  }
  SupportSet saturating_sub(uint32_t k) const {  // :120-134
    if (kind == Range) return range(start > k ? start - k : 0, end ? std::optional<uint32_t>(*end > k ? *end - k : 0) : std::nullopt);
    if (kind == Interval) {
      SupportSet r = *this;
      r.lo = (lo - Rat::from_u32(k)).max(Rat());
      r.hi = (hi - Rat::from_u32(k)).max(Rat());
      return r;
    }
    return *this;
  }
  bool is_subset_of(const SupportSet& o) const {  // :165-194
    if (is_empty()) return true;
    if (o.is_empty()) return false;
    if (kind == Interval && o.kind == Range) return false;
    if (kind == Range && o.kind == Range) return start >= o.start && (!o.end || (end && *end <= *o.end));
    if (kind == Interval && o.kind == Interval) return lo >= o.lo && hi <= o.hi;
    return Rat::from_u32(start) >= o.lo && end && Rat::from_u32(*end) <= o.hi;
  }
  void retain_only(std::vector<uint32_t> set) {  // :196-227
    std::sort(set.begin(), set.end());
    if (kind != Range) return;
    std::optional<uint32_t> ns, ne;
    for (uint32_t v : set)
      if (start <= v && v <= end.value_or(UINT32_MAX)) {
        if (!ns) ns = v;
        ne = v;
      }
    if (ns) *this = range(*ns, ne);
    else *this = empty();
  }
  void remove_all(std::vector<uint32_t> set) {  // :229-263
    std::sort(set.begin(), set.end());
    if (kind != Range || set.empty()) return;
    for (uint32_t v : set)
      if (v == start) start = v + 1;
    if (end) {
      for (auto it = set.rbegin(); it != set.rend(); ++it)
        if (*it == *end) {
          if (*it == 0) { end = 0; start = 1; }
          else end = *it - 1;
        }
    }
    if (start > end.value_or(UINT32_MAX)) *this = empty();
  }
  bool contains(uint32_t i) const {  // :291-300
    if (kind == Range) return i >= start && (!end || i <= *end);
    if (kind == Interval) return Rat::from_u32(i) >= lo && Rat::from_u32(i) <= hi;
    return false;
  }
  SupportSet add(const SupportSet& o) const {  // :378-441
    if (is_empty()) return o;
    if (o.is_empty()) return *this;
    if (kind == Range && o.kind == Range) {
      uint64_t s = (uint64_t)start + o.start;
      std::optional<uint32_t> e;
      if (end && o.end) {
        uint64_t t = (uint64_t)*end + *o.end;
        if (t <= UINT32_MAX) e = (uint32_t)t;   // checked_add
      }
      return range(s > UINT32_MAX ? UINT32_MAX : (uint32_t)s, e);   // saturating_add
    }
    auto as_iv = [](const SupportSet& s) {
      if (s.kind == Interval) return std::make_pair(s.lo, s.hi);
      return std::make_pair(Rat::from_u32(s.start), s.end ? Rat::from_u32(*s.end) : Rat::infinity());
    };
    auto a = as_iv(*this), b = as_iv(o);
    SupportSet r; r.kind = Interval; r.lo = a.first + b.first; r.hi = a.second + b.second;
    return r;
  }
  SupportSet mul_u32(uint32_t k) const {  // :451-466
    if (kind == Range) return range(start * k, end ? std::optional<uint32_t>(*end * k) : std::nullopt);
    if (kind == Interval) { SupportSet r = *this; r.lo = lo * Rat::from_u32(k); r.hi = hi * Rat::from_u32(k); return r; }
    return *this;
  }
  std::string str() const {  // :346-368
    if (kind == Empty) return "\xe2\x88\x85";
    if (kind == Range) {
      if (end) {
        if (start == *end) return "{" + std::to_string(start) + "}";
        return "{" + std::to_string(start) + ", ..., " + std::to_string(*end) + "}";
      }
      return "{" + std::to_string(start) + ", ...}";
    }
    if (hi.is_inf()) return "[" + lo.str() + ", \xe2\x88\x9e)";
    return "[" + lo.str() + ", " + hi.str() + "]";
  }
};

inline SupportSet dist_support(const Distribution& d) {  // ppl.rs:208-238
  switch (d.kind) {
    case DistKind::Dirac:
      if (auto a = d.p.as_integer()) return SupportSet::point(*a);
      return SupportSet::interval(Rat(d.p.numer, d.p.denom), Rat(d.p.numer, d.p.denom));
    case DistKind::Bernoulli: case DistKind::BernoulliVarProb: return SupportSet::from_range_incl(0, 1);
    case DistKind::Binomial: return SupportSet::from_range_incl(0, d.n);
    case DistKind::Categorical: return SupportSet::from_range_excl(0, (uint32_t)d.rs.size());
    case DistKind::Uniform: return SupportSet::from_range_excl(d.n, d.m);
    case DistKind::Exponential: case DistKind::Gamma: return SupportSet::nonneg_reals();
    case DistKind::UniformCont: return SupportSet::interval(Rat(d.p.numer, d.p.denom), Rat(d.q.numer, d.q.denom));
    default: return SupportSet::naturals();
  }
}

struct VarSupport {  // semantics/support.rs:9-143
  bool is_empty = false;
  size_t n_empty = 0;
  std::vector<SupportSet> sets;

  static VarSupport empty(size_t n) { VarSupport v; v.is_empty = true; v.n_empty = n; return v; }
  static VarSupport zero(size_t n) { VarSupport v; v.sets.assign(n, SupportSet::zero()); return v; }
  size_t num_vars() const { return is_empty ? n_empty : sets.size(); }
  const SupportSet& operator[](Var v) const {
    static const SupportSet kEmpty = SupportSet::empty();
    return is_empty ? kEmpty : sets.at(v);
  }
  bool operator==(const VarSupport& o) const {
    if (is_empty != o.is_empty) return false;
    return is_empty ? n_empty == o.n_empty : sets == o.sets;
  }
  void push(const SupportSet& s) {
    if (is_empty) n_empty++;
    else sets.push_back(s);
  }
  void normalize() {
    if (is_empty) return;
    for (auto& s : sets)
      if (s.is_empty()) { *this = empty(sets.size()); return; }
  }
  bool is_subset_of(const VarSupport& o) const {
    if (is_empty) return true;
    if (o.is_empty) return false;
    for (size_t i = 0; i < sets.size(); i++)
      if (!sets[i].is_subset_of(o.sets[i])) return false;
    return true;
  }
  VarSupport join(const VarSupport& o) const {
    if (is_empty) return o;
    if (o.is_empty) return *this;
    VarSupport r;
    for (size_t i = 0; i < sets.size(); i++) r.sets.push_back(sets[i].join(o.sets[i]));
    r.normalize();
    return r;
  }
  template <class F> void update(Var v, F&& f) {
    if (!is_empty) f(sets.at(v));
    normalize();
  }
  void set(Var v, const SupportSet& s) { update(v, [&](SupportSet& x) { x = s; }); }
};

class SupportTransformer {  // semantics/support.rs:145-386
 public:
  size_t unroll = 0;

  VarSupport init(const Program& p) { return VarSupport::zero(p.used_vars()); }

  std::pair<VarSupport, VarSupport> transform_event(const Event& e, VarSupport init) {
    switch (e.kind) {
      case Event::InSet: {
        VarSupport then_s = init, else_s = init;
        std::vector<uint32_t> set(e.set.begin(), e.set.end());
        then_s.update(e.v1, [&](SupportSet& s) { s.retain_only(set); });
        else_s.update(e.v1, [&](SupportSet& s) { s.remove_all(set); });
        return {then_s, else_s};
      }
      case Event::DataFromDist: case Event::VarComparison: return {init, init};
      case Event::Complement: {
        auto r = transform_event(*e.children[0], init);
        return {r.second, r.first};
      }
      case Event::Intersection: {
        VarSupport else_s = VarSupport::empty(init.num_vars()), then_s = init;
        for (auto& c : e.children) {
          auto r = transform_event(*c, then_s);
          then_s = r.first;
          else_s = else_s.join(r.second);
        }
        return {then_s, else_s};
      }
    }
    throw EvalError("unreachable");
  }

  static VarSupport transform_distribution(const Distribution& d, Var v, VarSupport init, bool add_previous) {  // :250-266
    VarSupport r = init;
    if (v == r.num_vars()) r.push(SupportSet::zero());
    GFE_ASSERT(v < r.num_vars(), "variable index out of range");
    if (!add_previous) r.set(v, SupportSet::zero());
    SupportSet ds = dist_support(d);
    r.update(v, [&](SupportSet& s) { s = s.add(ds); });
    return r;
  }

  VarSupport transform_statements(const Block& b, VarSupport cur) {
    for (auto& s : b) cur = transform_statement(s, cur);
    return cur;
  }

  VarSupport transform_statement(const Statement& st, VarSupport init) {
    switch (st.kind) {
      case Statement::Sample: return transform_distribution(st.dist, st.var, init, st.add_previous_value);
      case Statement::Assign: {
        SupportSet ns = init[st.var];
        if (!st.add_previous_value) ns = SupportSet::zero();
        if (st.has_addend) ns = ns.add(init[st.addend_var].mul_u32(st.addend_factor));
        ns = ns.add(SupportSet::point(st.offset));
        init.set(st.var, ns);
        return init;
      }
      case Statement::Decrement:
        init.update(st.var, [&](SupportSet& s) { s = s.saturating_sub(st.offset); });
        return init;
      case Statement::IfThenElse: {
        auto r = transform_event(*st.cond, init);
        return transform_statements(st.then_, r.first).join(transform_statements(st.else_, r.second));
      }
      case Statement::While: {
        size_t count = st.unroll.value_or(unroll);
        if (auto fix = find_unroll_fixpoint(*st.cond, st.then_, init)) count = std::max(count, *fix);
        VarSupport pre = init, rest = VarSupport::empty(init.num_vars());
        for (size_t i = 0; i < count; i++) {
          auto it = one_iteration(pre, st.then_, *st.cond);
          rest = rest.join(it.second);
          pre = it.first;
        }
        VarSupport inv = find_while_invariant(*st.cond, st.then_, pre);
        auto ex = transform_event(*st.cond, inv);
        return rest.join(ex.second);
      }
      case Statement::Fail: return VarSupport::empty(init.num_vars());
      case Statement::Normalize: return transform_normalize(st.given_vars, 0, st.then_, init);
    }
    throw EvalError("unreachable");
  }

  std::optional<size_t> find_unroll_fixpoint(const Event& cond, const Block& body, VarSupport init) {  // :268-285
    VarSupport pre = init;
    for (size_t i = 0; i < 100; i++) {
      auto it = one_iteration(pre, body, cond);
      if (pre == it.first) return i;
      pre = it.first;
    }
    return std::nullopt;
  }

  VarSupport find_while_invariant(const Event& cond, const Block& body, VarSupport init) {  // :287-330
    VarSupport pre = init;
    for (int i = 0; i < 100; i++) {
      auto it = one_iteration(pre, body, cond);
      if (it.first.is_subset_of(pre)) return pre;
      pre = pre.join(it.first);
    }
    for (size_t i = 0; i <= 2 * pre.num_vars(); i++) {
      auto it = one_iteration(pre, body, cond);
      if (it.first.is_subset_of(pre)) return pre;
      for (size_t v = 0; v < pre.num_vars(); v++) pre.set(v, widen(pre[v], it.first[v]));
    }
    auto it = one_iteration(pre, body, cond);
    GFE_ASSERT(it.first.is_subset_of(pre), "Widening failed.");
    return pre;
  }

 private:
  static SupportSet widen(const SupportSet& cur, const SupportSet& nw) {  // :332-350
    GFE_ASSERT(cur.kind == SupportSet::Range && nw.kind == SupportSet::Range, "Cannot widen non-range supports");
    uint32_t s = cur.start <= nw.start ? cur.start : 0;
    std::optional<uint32_t> e;
    if (cur.end && nw.end && *nw.end <= *cur.end) e = cur.end;
    return SupportSet::range(s, e);
  }
  std::pair<VarSupport, VarSupport> one_iteration(VarSupport init, const Block& body, const Event& cond) {
    auto r = transform_event(cond, init);
    return {transform_statements(body, r.first), r.second};
  }
  VarSupport transform_normalize(const std::vector<Var>& given, size_t idx, const Block& block, VarSupport vi) {  // :363-385
    if (idx == given.size()) return transform_statements(block, vi);
    Var v = given[idx];
    auto range = vi[v].finite_nonempty_range();
    GFE_ASSERT(range.has_value(), "Cannot normalize with respect to a variable whose value could not be proven to be bounded.");
    VarSupport joined = VarSupport::empty(vi.num_vars());
    for (uint32_t i = range->first; i <= range->second; i++) {
      VarSupport nv = vi;
      nv.set(v, SupportSet::point(i));
      joined = joined.join(transform_normalize(given, idx + 1, block, nv));
      if (i == UINT32_MAX) break;
    }
    return joined;
  }
};

}  // namespace gfe
