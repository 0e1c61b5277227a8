// ORACLE -- TEST INFRASTRUCTURE ONLY.
// The product's host evaluator (genfer_b200/csrc/evaluator/*.hpp: parser, GF translation, GenFun::eval, report --
// all host control logic) instantiated over the CPU restatement of TaylorPoly<f64> (taylor_oracle.hpp) instead of
// the CUDA library.  With the reference's operation order and separate multiply/add this reproduces the
// reference's stdout byte for byte on its .expect fixtures (tests/test_oracle_sgcl.py), which pins both the
// oracle's arithmetic and the shared host logic.  Never linked into or called from the product path.
// A second instantiation over TaylorPoly<Interval<f64>> (orc_run_sgcl_bounds) gives the --bounds-style enclosure used by
// the enclosure tests; PARITY UNPINNED for that one: no reference fixture runs with --bounds.
#include <cstring>
#include <memory>

#include "../genfer_b200/csrc/evaluator/report.hpp"
#include "taylor_oracle.hpp"

namespace {

// Scalar adaptors: how the evaluator's scalar type S maps onto the element type T of the oracle's TaylorPoly<T>.
struct F64Scalar {
  using T = double;
  using S = double;
  static T raw(S x) { return x; }
  static S wrap(T x) { return x; }
  static S max(S x, S y) { return x > y ? x : y; }   // F64::max (number/f64.rs:77-84)
};
// Interval<F64> (src/interval.rs) as the evaluator sees it: constructible from an f64 constant (a point interval)
struct IvS {
  orc::Interval v{0.0, 0.0};
  IvS() = default;
  IvS(double x) : v{x, x} {}
  explicit IvS(orc::Interval i) : v(i) {}
  static IvS from_bounds(double lo, double hi) { return IvS(orc::Interval{lo, hi}); }
  double lower() const { return v.lo; }
  double upper() const { return v.hi; }
  friend IvS operator+(const IvS& a, const IvS& b) { return IvS(a.v + b.v); }
  friend IvS operator-(const IvS& a, const IvS& b) { return IvS(a.v - b.v); }
  friend IvS operator*(const IvS& a, const IvS& b) { return IvS(a.v * b.v); }
  friend IvS operator/(const IvS& a, const IvS& b) { return IvS(a.v / b.v); }
  friend bool operator==(const IvS& a, const IvS& b) { return a.v == b.v; }
};
struct IvScalar {
  using T = orc::Interval;
  using S = IvS;
  static T raw(const S& x) { return x.v; }
  static S wrap(const T& x) { return IvS(x); }
  static S max(const S& x, const S& y) {   // Interval::max (interval.rs): componentwise
    return IvS(orc::Interval{x.v.lo > y.v.lo ? x.v.lo : y.v.lo, x.v.hi > y.v.hi ? x.v.hi : y.v.hi});
  }
};

template <class A>
struct OracleBackendT {
  using T = typename A::T;
  using Scalar = typename A::S;
  using TP = orc::TaylorPoly<T>;
  using Poly = std::shared_ptr<const TP>;
  static Scalar scalar_max(const Scalar& x, const Scalar& y) { return A::max(x, y); }
  static Poly mk(TP t) { return std::make_shared<const TP>(std::move(t)); }
  static std::vector<orc::usize> us(const std::vector<uint64_t>& v) {
    std::vector<orc::usize> r;
    for (uint64_t x : v) r.push_back(x == UINT64_MAX ? orc::UMAX : (orc::usize)x);
    return r;
  }
  Poly from_scalar(const Scalar& x) { return mk(TP::from_scalar(A::raw(x))); }
  Poly var(size_t v, const Scalar& x, size_t len) { return mk(TP::var(v, A::raw(x), len)); }
  Poly var_at_zero(size_t v, size_t len) { return mk(TP::var_at_zero(v, len)); }
  Poly var_with_degrees(size_t v, const Scalar& x, const std::vector<uint64_t>& d) { return mk(TP::var_with_degrees_p1(v, A::raw(x), us(d))); }
  Poly zero_with(const std::vector<uint64_t>& d) { return mk(TP::zero_with(us(d))); }
  Poly new_poly(const std::vector<uint64_t>& shape, const std::vector<uint64_t>& degrees, const double* data) {   // f64 GenFun constants
    orc::Arr<T> a(us(shape), A::raw(Scalar(0.0)));
    for (size_t i = 0; i < a.data.size(); i++) a.data[i] = A::raw(Scalar(data[i]));
    return mk(TP(std::move(a), us(degrees)));
  }
  template <class Q = Scalar, class = std::enable_if_t<!std::is_same<Q, double>::value>>
  Poly new_poly(const std::vector<uint64_t>& shape, const std::vector<uint64_t>& degrees, const Scalar* data) {
    orc::Arr<T> a(us(shape), A::raw(Scalar(0.0)));
    for (size_t i = 0; i < a.data.size(); i++) a.data[i] = A::raw(data[i]);
    return mk(TP(std::move(a), us(degrees)));
  }
  Poly add(const Poly& a, const Poly& b) { return mk(orc::tp_add(*a, *b)); }
  Poly sub(const Poly& a, const Poly& b) { return mk(orc::tp_sub(*a, *b)); }
  Poly mul(const Poly& a, const Poly& b) { return mk(orc::tp_mul(*a, *b)); }
  Poly div(const Poly& a, const Poly& b) { return mk(orc::tp_div(*a, *b)); }
  Poly neg(const Poly& a) { return mk(orc::tp_neg(*a)); }
  Poly exp(const Poly& a) { return mk(a->exp()); }
  Poly log(const Poly& a) { return mk(a->log()); }
  Poly pow(const Poly& a, uint32_t e) { return mk(a->pow(e)); }
  Poly derivative(const Poly& a, size_t v, size_t n) { return mk(a->derivative(v, n)); }
  Poly taylor_expansion_of_coeff(const Poly& a, size_t v, size_t n) { return mk(a->taylor_expansion_of_coeff(v, n)); }
  Poly shift_down(const Poly& a, size_t v, size_t n) { return mk(a->shift_down(v, n)); }
  Poly coefficients_of_term(const Poly& a, size_t v, size_t n) { return mk(a->coefficients_of_term(v, n)); }
  Poly taylor_polynomial_terms(const Poly& a, size_t v, const std::vector<size_t>& orders) {
    return mk(a->taylor_polynomial_terms(v, std::vector<orc::usize>(orders.begin(), orders.end())));
  }
  Poly subst_var(const Poly& a, size_t v, const Poly& s) { return mk(a->subst_var(v, *s)); }
  Poly truncate_to_degree_p1(const Poly& a, size_t d) { return mk(a->truncate_to_degree_p1(d)); }
  Poly remove_last_variable(const Poly& a) { return mk(a->remove_last_variable()); }
  Poly extend_to_dim(const Poly& a, size_t ndim, size_t d) { return mk(a->extend_to_dim(ndim, d)); }
  Scalar constant_term(const Poly& a) { return A::wrap(a->constant_term()); }
  std::vector<Scalar> gather_axis(const Poly& a, size_t v, size_t count) {   // `count` coefficient() calls (:959-965)
    std::vector<Scalar> out;
    std::vector<orc::usize> idx(a->num_vars(), 0);
    for (size_t i = 0; i < count; i++) {
      idx.at(v) = i;
      out.push_back(A::wrap(a->coefficient(idx)));
    }
    return out;
  }
  size_t num_vars(const Poly& a) { return a->num_vars(); }
  std::vector<uint64_t> array_shape(const Poly& a) { return std::vector<uint64_t>(a->coeffs.shape.begin(), a->coeffs.shape.end()); }
  std::optional<Scalar> extract_constant(const Poly& a) {
    auto c = a->extract_constant();
    if (!c) return std::nullopt;
    return A::wrap(*c);
  }
  std::vector<Scalar> to_host(const Poly& a) {
    std::vector<Scalar> out;
    for (const T& x : a->coeffs.data) out.push_back(A::wrap(x));
    return out;
  }
  // ---- TaylorExpansion<T> (univariate_taylor.rs) for the symbolic mode (only instantiated for T = f64) ----
  using TE = orc::TaylorExpansion<T>;
  using Uni = std::shared_ptr<const TE>;
  static Uni umk(TE t) { return std::make_shared<const TE>(std::move(t)); }
  Uni uni_constant(double x) { return umk(TE::constant(A::raw(Scalar(x)))); }
  Uni uni_var(double x, size_t order) { return umk(TE::var(A::raw(Scalar(x)), order)); }
  Uni uni_add(const Uni& a, const Uni& b) { return umk(orc::te_add(*a, *b)); }
  Uni uni_mul(const Uni& a, const Uni& b) { return umk(orc::te_mul(*a, *b)); }
  Uni uni_div(const Uni& a, const Uni& b) { return umk(orc::te_div(*a, *b)); }
  Uni uni_exp(const Uni& a) { return umk(a->exp()); }
  Uni uni_log(const Uni& a) { return umk(a->log()); }
  Uni uni_pow(const Uni& a, uint32_t e) { return umk(orc::te_pow(*a, e)); }
  Uni uni_max(const Uni& a, const Uni& b) {
    if (!a->is_const || !b->is_const) throw gfe::EvalError("Maximum can only be applied to constant Taylor expansions.");
    return umk(TE::constant(A::raw(A::max(A::wrap(a->c), A::wrap(b->c)))));
  }
  Scalar uni_coeff(const Uni& a, size_t order) { return A::wrap(a->coeff(order)); }
};
using OracleBackend = OracleBackendT<F64Scalar>;
using IntervalBackend = OracleBackendT<IvScalar>;

}  // namespace

struct orc_sgcl_result {
  gfe::RunResult r;
};

extern "C" {
int orc_run_sgcl(const char* source, int64_t limit, int flags, uint64_t unroll, orc_sgcl_result** out, char* err, size_t err_cap) {
  try {
    gfe::RunOptions opt;
    if (limit >= 0) opt.limit = (size_t)limit;
    opt.no_probs = (flags & 1) != 0;
    opt.no_simplify_gf = (flags & 2) != 0;
    opt.bounds = (flags & 4) != 0;
    opt.symbolic = (flags & 8) != 0;
    opt.unroll = (size_t)unroll;
    auto res = std::make_unique<orc_sgcl_result>();
    if (opt.bounds) {   // run_program_intervals::<F64> (main.rs:145-185)
      IntervalBackend backend;
      res->r = gfe::run_program(backend, source, opt);
    } else {
      OracleBackend backend;
      res->r = gfe::run_program(backend, source, opt);
    }
    *out = res.release();
    return 0;
  } catch (const std::exception& e) {
    if (err && err_cap) {
      std::strncpy(err, e.what(), err_cap - 1);
      err[err_cap - 1] = '\0';
    }
    return 1;
  }
}
// --bounds-style enclosure of the evaluator's direct outputs: the same host logic over TaylorPoly<Interval<F64>>
// (no simplification pass: its polynomial form stores f64 coefficients).  GenFun constants are the f64 values the f64
// path uses together with the enclosure Number::from_ratio builds for them (evaluator/num.hpp), so this encloses both the exact
// posterior and the f64 / GPU evaluation of the DAG.
// out: [rest lo, hi, total lo, hi, raw moment 1..4 lo, hi ...] (12 doubles); probs: `limit` pairs [lo, hi].
int orc_run_sgcl_bounds(const char* source, int64_t limit, uint64_t unroll, double* out12, double* probs_lohi, char* err, size_t err_cap) {
  try {
    IntervalBackend backend;
    gfe::Program program = gfe::parse_program(source);
    gfe::GfTransformer transformer((size_t)unroll);
    gfe::GfTranslation tr = transformer.semantics(program);
    gfe::Evaluator<IntervalBackend> ev(backend);
    IvS rest = backend.constant_term(ev.eval(tr.rest, std::vector<IvS>(tr.var_info.num_vars(), IvS(0.0)), 1));
    auto mom = ev.moments_taylor(tr.gf, program.result, tr.var_info, 5);
    out12[0] = rest.v.lo; out12[1] = rest.v.hi;
    out12[2] = mom.first.v.lo; out12[3] = mom.first.v.hi;
    for (size_t i = 0; i < 4; i++) { out12[4 + 2 * i] = mom.second.at(i).v.lo; out12[5 + 2 * i] = mom.second.at(i).v.hi; }
    if (limit > 0 && probs_lohi) {
      std::vector<IvS> p = ev.probs_taylor(tr.gf, program.result, tr.var_info, (size_t)limit);
      for (size_t i = 0; i < (size_t)limit; i++) { probs_lohi[2 * i] = p.at(i).v.lo; probs_lohi[2 * i + 1] = p.at(i).v.hi; }
    }
    return 0;
  } catch (const std::exception& e) {
    if (err && err_cap) {
      std::strncpy(err, e.what(), err_cap - 1);
      err[err_cap - 1] = '\0';
    }
    return 1;
  }
}
void orc_sgcl_free(orc_sgcl_result* r) { delete r; }
const char* orc_sgcl_report(const orc_sgcl_result* r) { return r->r.report.c_str(); }
void orc_sgcl_moments(const orc_sgcl_result* r, double* out11) {
  const gfe::RunResult& x = r->r;
  const double v[11] = {x.total, x.mean, x.raw2, x.raw3, x.raw4, x.stddev, x.variance, x.central3, x.central4, x.skewness, x.kurtosis};
  std::memcpy(out11, v, sizeof(v));
}
void orc_sgcl_stats(const orc_sgcl_result* r, uint64_t* nodes, uint64_t* hits) { *nodes = r->r.nodes_evaluated; *hits = r->r.cache_hits; }
uint64_t orc_sgcl_limit(const orc_sgcl_result* r) { return r->r.probs.size(); }
void orc_sgcl_moment_bounds(const orc_sgcl_result* r, double* out22) {
  for (size_t i = 0; i < r->r.moment_bounds.size() && i < 11; i++) {
    out22[2 * i] = r->r.moment_bounds[i].lo;
    out22[2 * i + 1] = r->r.moment_bounds[i].hi;
  }
}
void orc_sgcl_prob_bounds(const orc_sgcl_result* r, double* unnormalized_pairs) {
  for (size_t i = 0; i < r->r.prob_bounds.size(); i++) {
    unnormalized_pairs[2 * i] = r->r.prob_bounds[i].lo;
    unnormalized_pairs[2 * i + 1] = r->r.prob_bounds[i].hi;
  }
}
void orc_sgcl_probs(const orc_sgcl_result* r, double* unnormalized, double* normalized) {
  const gfe::RunResult& x = r->r;
  for (size_t i = 0; i < x.probs.size(); i++) {
    if (unnormalized) unnormalized[i] = x.probs[i];
    if (normalized) normalized[i] = x.is_normalized ? x.probs[i] : x.normalized_probs[i];
  }
}
}
