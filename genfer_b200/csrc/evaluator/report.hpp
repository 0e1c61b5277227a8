// Host evaluator -- driver and report: what the reference's `genfer file.sgcl` prints (--no-timing).
// Restates src/main.rs: run_program :187-227, translate_program_to_gf :229-254, the interval post-processing and
// clamping of print_moments_and_probs_interval :301-389 (Z clamped to [0,1], rest mass, moments structure), the
// automatic limit by Markov's inequality and the probability table of print_probs :391-473, print_moments
// :541-573; moment conversions src/generating_function.rs:1035-1086; Interval<F64> src/interval.rs (every
// operation widens by one ulp each side, :29-31); F64 scalar semantics src/number/f64.rs (powi :64-66, ryu
// shortest round-trip Display :41-45, next_up / next_down :127-171).
#pragma once
#include <charconv>
#include <cstring>

#include "eval.hpp"
#include "parser.hpp"
#include "symbolic.hpp"
#include "translate.hpp"

namespace gfe {

// ryu::Buffer::format (pretty::format64 layout rules on the shortest round-trip digits)
inline std::string fmt_f64(double x) {
  if (std::isnan(x)) return "NaN";
  if (std::isinf(x)) return x > 0 ? "inf" : "-inf";
  std::string out;
  if (std::signbit(x)) { out += '-'; x = -x; }
  if (x == 0.0) return out + "0.0";
  char buf[64];
  auto res = std::to_chars(buf, buf + sizeof(buf), x, std::chars_format::scientific);
  std::string sci(buf, res.ptr);                  // d[.ddd]e[+-]XX
  size_t epos = sci.find('e');
  std::string digits;
  for (size_t i = 0; i < epos; i++)
    if (sci[i] != '.') digits += sci[i];
  int exp10 = std::stoi(sci.substr(epos + 1));    // value = d.ddd * 10^exp10
  const int length = (int)digits.size();
  const int kk = exp10 + 1;                       // 10^(kk-1) <= v < 10^kk
  const int k = kk - length;                      // v = digits * 10^k
  if (0 <= k && kk <= 16) {                       // 1234e7 -> 12340000000.0
    out += digits;
    out.append((size_t)(kk - length), '0');
    out += ".0";
  } else if (0 < kk && kk <= 16) {                // 1234e-2 -> 12.34
    out += digits.substr(0, (size_t)kk) + "." + digits.substr((size_t)kk);
  } else if (-5 < kk && kk <= 0) {                // 1234e-6 -> 0.001234
    out += "0.";
    out.append((size_t)(-kk), '0');
    out += digits;
  } else if (length == 1) {                       // 1e30
    out += digits + "e" + std::to_string(kk - 1);
  } else {                                        // 1234e30 -> 1.234e33
    out += digits.substr(0, 1) + "." + digits.substr(1) + "e" + std::to_string(kk - 1);
  }
  return out;
}

inline std::string in_interval(const Iv& iv, bool print_intervals) {  // main.rs:291-299
  if (iv.is_point()) return "= " + fmt_f64(iv.lo);
  if (!print_intervals) return "= " + fmt_f64(iv.center());
  return "\xe2\x88\x88 [" + fmt_f64(iv.lo) + ", " + fmt_f64(iv.hi) + "]";
}

// moments_to_central_moments over Interval<F64> (generating_function.rs:1035-1060)
inline std::pair<Iv, std::vector<Iv>> moments_to_central_moments(const std::vector<Iv>& moments) {
  const size_t len = moments.size() + 1;
  Iv mean = moments[0];
  std::vector<std::vector<Iv>> bc(len, std::vector<Iv>(len, Iv::zero()));
  for (size_t n = 0; n < len; n++) {
    bc[n][0] = Iv::one();
    bc[n][n] = Iv::one();
    for (size_t k = 1; k < n; k++) bc[n][k] = bc[n - 1][k - 1].add(bc[n - 1][k]);
  }
  Iv neg_mean = mean.neg();
  std::vector<Iv> central(len - 2, Iv::zero());
  for (size_t n = 2; n < len; n++) {
    for (size_t k = 1; k <= n; k++)
      central[n - 2] = central[n - 2].add(bc[n][k].mul(neg_mean.pow((uint32_t)(n - k))).mul(moments[k - 1]));
    central[n - 2] = central[n - 2].add(neg_mean.pow((uint32_t)n));
  }
  return {mean, central};
}
// central_to_standardized_moments (:1062-1086)
inline std::pair<Iv, std::vector<Iv>> central_to_standardized_moments(const std::vector<Iv>& central) {
  Iv variance = central[0];
  Iv sigma = variance.sqrt();
  std::vector<Iv> result;
  for (size_t i = 0; i + 1 < central.size(); i++) {
    const Iv& x = central[i + 1];
    if (x.is_zero() && !variance.is_nan() && !variance.is_zero()) {
      result.push_back(x);
    } else {
      Iv sp = (i % 2 == 0) ? sigma.pow((uint32_t)(i + 3)) : variance.pow((uint32_t)((i + 3) / 2));
      result.push_back(x.div(sp));
    }
  }
  return {variance, result};
}

// Minimal string builder (no iostreams: the library is loaded into processes that may carry a second C++ runtime,
// and locale-dependent stream formatting is not needed for a byte-exact report).
struct Out {
  std::string s;
  Out& operator<<(const char* t) { s += t; return *this; }
  Out& operator<<(const std::string& t) { s += t; return *this; }
  Out& operator<<(size_t v) { s += std::to_string(v); return *this; }
  const std::string& str() const { return s; }
};

struct RunOptions {
  std::optional<size_t> limit;   // --limit
  bool no_probs = false;          // --no-probs
  bool no_simplify_gf = false;    // --no-simplify-gf
  size_t unroll = 8;              // --unroll (default 8, main.rs:66-67)
  bool symbolic = false;          // -s: one symbolic computation DAG evaluated over univariate Taylor expansions (symbolic.hpp)
  bool bounds = false;            // --bounds: non-point intervals are printed as such (with an interval backend they are the result)
};

struct RunResult {
  std::string report;             // stdout of the reference with --no-timing
  std::string support;
  double total = 0, mean = 0, raw2 = 0, raw3 = 0, raw4 = 0, stddev = 0, variance = 0, central3 = 0, central4 = 0, skewness = 0, kurtosis = 0;
  size_t limit = 0;
  std::vector<double> probs, normalized_probs;
  double tail_unnorm = 0, tail_norm = 0;
  bool is_normalized = true;
  size_t nodes_evaluated = 0, cache_hits = 0;
  // the intervals behind the values above (points for an f64 run without rest mass): Z, E, raw 2..4, sigma, V, central 3, 4, S, K
  std::vector<Iv> moment_bounds, prob_bounds, normalized_prob_bounds;
};

constexpr size_t MAX_PROB_LIMIT = 1000;   // main.rs:30

template <class B>
RunResult run_program(B& backend, const std::string& source, const RunOptions& opt) {
  RunResult out;
  Out os;
  Program program = parse_program(source);
  const bool uses_observe = program.uses_observe();
  GfTransformer transformer(opt.unroll);
  GfTranslation tr = transformer.semantics(program);
  os << transformer.warnings;
  Evaluator<B> ev(backend);
  using S = typename B::Scalar;
  // The interval instantiations evaluate the unsimplified DAG: GenFun::simplify stores its polynomial forms with f64 coefficients
  // here.  (The reference simplifies in --bounds mode too; either DAG gives a valid enclosure of the same exact value.)
  if constexpr (std::is_same<S, double>::value) {
    if (!opt.no_simplify_gf) {
      tr.gf = ev.simplify(tr.gf);
      tr.rest = ev.simplify(tr.rest);
    }
  }
  const SupportSet var_info = tr.var_info[program.result];
  const SupportSet rest_info = tr.rest_info[program.result];
  out.support = var_info.str();
  os << "Support is a subset of: " << out.support << "\n\nComputing moments...\n";

  // ---- print_moments_and_probs_interval (:301-389) ----
  S rest_val;
  std::pair<S, std::vector<S>> mom;
  Sym sym_gf;
  if constexpr (std::is_same<S, double>::value) {
    if (opt.symbolic) {   // run_program (:196-209): gf.to_computation(), rest.to_computation().evaluate_closed()
      sym_gf = sym::to_computation(tr.gf);
      Sym sym_rest = sym::to_computation(tr.rest);
      SymEvaluator<B> sev(backend);
      rest_val = sev.closed(sym_rest);
      mom = sev.moments(sym_gf, program.result, tr.var_info, 5);
      out.nodes_evaluated += sev.nodes_evaluated;
    }
  } else {
    GFE_ASSERT(!opt.symbolic, "symbolic mode runs over F64 only");
  }
  if (!opt.symbolic) {
    rest_val = backend.constant_term(ev.eval(tr.rest, std::vector<S>(tr.var_info.num_vars(), S(0.0)), 1));
    mom = ev.moments_taylor(tr.gf, program.result, tr.var_info, 5);
  }
  Iv rest = bounds_of<S>(rest_val).ensure_lower_bound(0.0).ensure_upper_bound(1.0).unite(0.0);
  Iv total = bounds_of<S>(mom.first).ensure_lower_bound(0.0).ensure_upper_bound(1.0);
  const Iv total_without_rest = total;
  Iv max_rest = Iv::one().sub(total_without_rest);
  rest = rest.ensure_upper_bound(max_rest.hi);
  total = total.add(rest).ensure_upper_bound(1.0);
  std::vector<Iv> moments;
  for (const S& m : mom.second) moments.push_back(bounds_of<S>(m).ensure_lower_bound(0.0));
  {  // rest_info.to_interval() (support.rs:265-289)
    std::optional<Iv> range;
    if (rest_info.kind == SupportSet::Range) range = Iv::exact((double)rest_info.start, rest_info.end ? (double)*rest_info.end : INFINITY);
    else if (rest_info.kind == SupportSet::Interval) range = Iv::exact(rest_info.lo.to_double(), rest_info.hi.to_double());
    if (range)
      for (size_t i = 0; i < moments.size(); i++) {
        double added = rest.hi * powi(range->hi, (uint32_t)i + 1);
        moments[i] = moments[i].add(Iv::exact(0.0, added));
      }
  }
  // moments_to_moments_struct (:515-539)
  Iv raw2 = moments[1], raw3 = moments[2], raw4 = moments[3];
  auto cm = moments_to_central_moments(moments);
  Iv mean = cm.first, central3 = cm.second[1], central4 = cm.second[2];
  auto sm = central_to_standardized_moments(cm.second);
  Iv variance = sm.first, skewness = sm.second[0], kurtosis = sm.second[1];
  Iv stddev = variance.sqrt();
  for (const Iv& m : moments) GFE_ASSERT(!(m.lt(Iv::zero())), "moments must be non-negative for distributions supported on the natural numbers");
  GFE_ASSERT(!variance.lt(Iv::zero()), "variance must be non-negative");
  GFE_ASSERT(!kurtosis.lt(Iv::zero()), "kurtosis must be non-negative");
  variance = variance.ensure_lower_bound(0.0);
  stddev = stddev.ensure_lower_bound(0.0);
  kurtosis = kurtosis.ensure_lower_bound(0.0);
  const bool pi = opt.bounds || !rest.is_zero();
  os << "Total measure:             Z " << in_interval(total, pi) << "\n";
  os << "Expected value:            E " << in_interval(mean, pi) << "\n";
  os << "2nd raw moment:         \xce\xbc'_2 " << in_interval(raw2, pi) << "\n";
  os << "3rd raw moment:         \xce\xbc'_3 " << in_interval(raw3, pi) << "\n";
  os << "4th raw moment:         \xce\xbc'_4 " << in_interval(raw4, pi) << "\n";
  os << "Standard deviation:        \xcf\x83 " << in_interval(stddev, pi) << "\n";
  os << "Variance (2nd central):    V " << in_interval(variance, pi) << "\n";
  os << "3rd central moment:      \xce\xbc_3 " << in_interval(central3, pi) << "\n";
  os << "4th central moment:      \xce\xbc_4 " << in_interval(central4, pi) << "\n";
  os << "Skewness (3rd std moment): S " << in_interval(skewness, pi) << "\n";
  os << "Kurtosis (4th std moment): K " << in_interval(kurtosis, pi) << "\n";
  auto val = [](const Iv& iv) { return iv.is_point() ? iv.lo : iv.center(); };
  out.total = val(total); out.mean = val(mean); out.raw2 = val(raw2); out.raw3 = val(raw3); out.raw4 = val(raw4);
  out.stddev = val(stddev); out.variance = val(variance); out.central3 = val(central3); out.central4 = val(central4);
  out.skewness = val(skewness); out.kurtosis = val(kurtosis);
  out.moment_bounds = {total, mean, raw2, raw3, raw4, stddev, variance, central3, central4, skewness, kurtosis};

  if (!(opt.no_probs || !var_info.is_discrete() || total.is_zero())) {
    // ---- print_probs (:391-473) ----
    os << "\n";
    Iv tot = total_without_rest.add(rest).ensure_upper_bound(1.0);
    size_t limit;
    if (opt.limit) limit = *opt.limit;
    else if (tot.is_zero()) limit = 1;
    else if (auto r = var_info.finite_nonempty_range()) limit = (size_t)r->second + 1;
    else {  // Markov's inequality: P(X >= limit) <= 1/256
      auto cm2 = moments_to_central_moments(moments);
      double c4root = std::sqrt(std::sqrt(cm2.second[2].hi));
      double lim = std::ceil(cm2.first.hi + 4.0 * c4root);
      if (std::isfinite(lim)) limit = std::min((size_t)lim + 1, MAX_PROB_LIMIT);
      else {
        os << "Failed to find a limit automatically due to non-finite moments.\n";
        os << "Please specify a limit manually with `--limit`.\nUsing a limit of 2 for now.\n";
        limit = 2;
      }
    }
    os << "Computing probabilities up to " << limit << "...\n";
    const bool is_normalized = !uses_observe || tot.is_one();
    Iv mass_missing = total_without_rest;
    std::vector<S> raw;
    if constexpr (std::is_same<S, double>::value) {
      if (opt.symbolic) {
        SymEvaluator<B> sev(backend);
        raw = sev.probs(sym_gf, program.result, tr.var_info, limit);
        out.nodes_evaluated += sev.nodes_evaluated;
      }
    }
    if (!opt.symbolic) raw = ev.probs_taylor(tr.gf, program.result, tr.var_info, limit);
    for (size_t i = 0; i < limit; i++) {
      Iv p = bounds_of<S>(raw[i]);
      mass_missing = mass_missing.sub(p);
      if (rest_info.contains((uint32_t)i)) p = p.add(rest);
      GFE_ASSERT(!(p.lt(Iv::zero()) || p.gt(Iv::one())), "p(" + std::to_string(i) + ") is not a probability");
      p = p.ensure_lower_bound(0.0).ensure_upper_bound(1.0);
      out.probs.push_back(val(p));
      out.prob_bounds.push_back(p);
      if (is_normalized) {
        os << "p(" << i << ") " << in_interval(p, pi) << "\n";
      } else {
        Iv np = p.div(tot).ensure_lower_bound(0.0).ensure_upper_bound(1.0);
        os << "Unnormalized: p(" << i << ")     " << in_interval(p, pi) << "\n";
        os << "Normalized:   p(" << i << ") / Z " << in_interval(np, pi) << "\n";
        out.normalized_probs.push_back(val(np));
        out.normalized_prob_bounds.push_back(np);
      }
    }
    SupportSet up_to = SupportSet::range(0, (uint32_t)limit - 1);
    if (!rest_info.is_subset_of(up_to)) mass_missing = mass_missing.add(rest);
    if (var_info.is_subset_of(up_to)) mass_missing = Iv::zero();
    double mm_unnorm = f64_min(f64_max(mass_missing.hi, 0.0), 1.0);
    double mm_norm = f64_min(f64_max(mass_missing.div(tot).hi, 0.0), 1.0);
    if (is_normalized) {
      os << "p(n) <= " << fmt_f64(mm_unnorm) << " for all n >= " << limit << "\n";
    } else {
      os << "Unnormalized: p(n)     <= " << fmt_f64(mm_unnorm) << " for all n >= " << limit << "\n";
      os << "Normalized:   p(n) / Z <= " << fmt_f64(mm_norm) << " for all n >= " << limit << "\n";
    }
    out.limit = limit;
    out.tail_unnorm = mm_unnorm;
    out.tail_norm = mm_norm;
    out.is_normalized = is_normalized;
  }
  out.report = os.str();
  out.nodes_evaluated += ev.nodes_evaluated;
  out.cache_hits = ev.cache_hits;
  return out;
}

}  // namespace gfe
