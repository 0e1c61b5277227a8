// Host evaluator -- SGCL parser.
// Recursive-descent restatement of the grammar of the reference's nom parser (src/parser.rs): comments `#`, `#= =#`
// (:564-580), naturals / ratios "a/b" and decimals (:18-69), events (:137-325), distributions (:362-497),
// sample / assign / decrement (:499-528), if / loop (unrolled at parse time :540-551) / while / normalize / fail /
// observe (= if cond {} else {fail}, :306-325), `return v` (:622-635).
#pragma once
#include <cctype>
#include <cstring>

#include "ast.hpp"

namespace gfe {

class Parser {
 public:
  explicit Parser(const std::string& text) : s_(text) {}

  Program parse_program() {
    Program prog;
    vars_ = &prog.var_names;
    while (true) {
      ws();
      size_t save = pos_;
      if (keyword("return")) { pos_ = save; break; }
      if (eof()) fail("missing return statement");
      statement(prog.stmts);
    }
    ws();
    GFE_ASSERT(keyword("return"), "expected `return`");
    prog.result = expect_var(identifier());
    if (peek() == ';') pos_++;
    ws();
    if (!eof()) fail("trailing input after return");
    return prog;
  }

 private:
  const std::string& s_;
  size_t pos_ = 0;
  std::vector<std::string>* vars_ = nullptr;

  [[noreturn]] void fail(const std::string& msg) const {
    size_t line = 1;
    for (size_t i = 0; i < pos_ && i < s_.size(); i++) line += s_[i] == '\n';
    throw EvalError("Parse error (line " + std::to_string(line) + "): " + msg);
  }
  bool eof() const { return pos_ >= s_.size(); }
  char peek(size_t o = 0) const { return pos_ + o < s_.size() ? s_[pos_ + o] : '\0'; }
  bool starts_with(const char* t) const { return s_.compare(pos_, strlen(t), t) == 0; }
  bool eat(const char* t) {
    if (starts_with(t)) { pos_ += strlen(t); return true; }
    return false;
  }
  static bool ident_rest(char c) { return std::isalnum((unsigned char)c) || c == '_'; }

  void ws() {  // :564-580
    while (true) {
      while (!eof() && std::isspace((unsigned char)s_[pos_])) pos_++;
      if (starts_with("#=")) {
        size_t e = s_.find("=#", pos_);
        if (e == std::string::npos) fail("Unterminated comment: found opening `#=` but no closing `=#`");
        pos_ = e + 2;
      } else if (peek() == '#') {
        while (!eof() && s_[pos_] != '\n' && s_[pos_] != '\r') pos_++;
      } else {
        return;
      }
    }
  }
  bool keyword(const char* k) {  // tag + not(identifier_rest) (:81-83)
    size_t n = strlen(k);
    if (s_.compare(pos_, n, k) != 0) return false;
    if (pos_ + n < s_.size() && ident_rest(s_[pos_ + n])) return false;
    pos_ += n;
    return true;
  }
  bool peek_keyword(const char* k) {
    size_t save = pos_;
    bool ok = keyword(k);
    pos_ = save;
    return ok;
  }
  std::optional<std::string> try_digits() {
    size_t b = pos_;
    while (!eof() && std::isdigit((unsigned char)s_[pos_])) pos_++;
    if (pos_ == b) return std::nullopt;
    return s_.substr(b, pos_ - b);
  }
  std::optional<uint64_t> try_u64() {  // ws digit1 ws
    size_t save = pos_;
    ws();
    auto d = try_digits();
    if (!d) { pos_ = save; return std::nullopt; }
    ws();
    return std::stoull(*d);
  }
  std::optional<Natural> try_natural() {
    auto v = try_u64();
    if (!v) return std::nullopt;
    GFE_ASSERT(*v <= UINT32_MAX, "natural number too large");
    return (Natural)*v;
  }
  Natural natural() {
    auto v = try_natural();
    if (!v) fail("expected natural number");
    return *v;
  }
  std::optional<PosRatio> try_pos_ratio() {  // :41-69
    size_t save = pos_;
    ws();
    size_t inner = pos_;
    if (auto n = try_u64()) {  // a / b
      if (peek() == '/') {
        pos_++;
        auto d = try_u64();
        if (!d) fail("expected denominator");
        ws();
        return PosRatio(*n, *d);
      }
    }
    pos_ = inner;
    auto integer = try_digits();
    if (!integer) { pos_ = save; return std::nullopt; }
    PosRatio r;
    if (peek() == '.') {
      pos_++;
      auto frac = try_digits();
      if (!frac) fail("expected digits after decimal point");
      uint64_t den = 1;
      for (size_t i = 0; i < frac->size(); i++) {
        GFE_ASSERT(den <= UINT64_MAX / 10, "too many decimal digits");
        den *= 10;
      }
      r = PosRatio(std::stoull(*integer + *frac), den);
    } else {
      r = PosRatio(std::stoull(*integer), 1);
    }
    ws();
    return r;
  }
  PosRatio pos_ratio() {
    auto r = try_pos_ratio();
    if (!r) fail("expected real number");
    return *r;
  }
  std::optional<std::string> try_identifier() {  // ws ident ws (:85-95)
    size_t save = pos_;
    ws();
    if (eof() || !(std::isalpha((unsigned char)peek()) || peek() == '_')) { pos_ = save; return std::nullopt; }
    size_t b = pos_;
    while (!eof() && ident_rest(s_[pos_])) pos_++;
    std::string id = s_.substr(b, pos_ - b);
    ws();
    return id;
  }
  std::string identifier() {
    auto id = try_identifier();
    if (!id) fail("expected identifier");
    return *id;
  }
  std::optional<Var> find_var(const std::string& id) const {
    for (size_t i = 0; i < vars_->size(); i++)
      if ((*vars_)[i] == id) return i;
    return std::nullopt;
  }
  Var find_or_create_var(const std::string& id) {
    if (auto v = find_var(id)) return *v;
    vars_->push_back(id);
    return vars_->size() - 1;
  }
  Var expect_var(const std::string& id) const {
    auto v = find_var(id);
    if (!v) throw EvalError("Unknown variable " + id);
    return *v;
  }
  void expect_char(char c) {
    if (peek() != c) fail(std::string("expected `") + c + "`");
    pos_++;
  }
  void semicolon() {
    ws();
    expect_char(';');
  }
  std::vector<Natural> natural_list() {  // :30-39
    ws();
    expect_char('[');
    std::vector<Natural> out;
    if (auto n = try_natural()) {
      out.push_back(*n);
      while (peek() == ',') {
        pos_++;
        out.push_back(natural());
      }
    }
    expect_char(']');
    ws();
    return out;
  }

  // ---- events ----------------------------------------------------------------------------------
  struct Operand { bool is_var; Var var; Natural nat; };
  std::optional<Operand> try_operand() {  // :137-146
    if (auto n = try_natural()) return Operand{false, 0, *n};
    if (auto id = try_identifier()) return Operand{true, expect_var(*id), 0};
    return std::nullopt;
  }
  Operand operand() {
    auto o = try_operand();
    if (!o) fail("expected comparee");
    return *o;
  }
  static std::vector<Natural> range_to(Natural n, bool inclusive) {
    std::vector<Natural> v;
    for (Natural i = 0; inclusive ? i <= n : i < n; i++) v.push_back(i);
    return v;
  }
  static EventP event_eq(const Operand& l, const Operand& r) {  // :148-158
    if (l.is_var && r.is_var) return Event::var_comparison(l.var, Comparison::Eq, r.var);
    if (l.is_var) return Event::in_set(l.var, {r.nat});
    if (r.is_var) return Event::in_set(r.var, {l.nat});
    return l.nat == r.nat ? Event::always() : Event::never();
  }
  static EventP event_lt(const Operand& l, const Operand& r) {  // :160-171
    if (l.is_var && r.is_var) return Event::var_comparison(l.var, Comparison::Lt, r.var);
    if (l.is_var) return Event::in_set(l.var, range_to(r.nat, false));
    if (r.is_var) return Event::complement(Event::in_set(r.var, range_to(l.nat, true)));
    return l.nat < r.nat ? Event::always() : Event::never();
  }
  static EventP event_le(const Operand& l, const Operand& r) {  // :173-186
    if (l.is_var && r.is_var) return Event::var_comparison(l.var, Comparison::Le, r.var);
    if (l.is_var) return Event::in_set(l.var, range_to(r.nat, true));
    if (r.is_var) return Event::complement(Event::in_set(r.var, range_to(l.nat, false)));
    return l.nat <= r.nat ? Event::always() : Event::never();
  }
  static EventP event_in(const Operand& l, std::vector<Natural> ns) {  // :188-194
    if (l.is_var) return Event::in_set(l.var, std::move(ns));
    for (Natural n : ns)
      if (n == l.nat) return Event::always();
    return Event::never();
  }
  EventP try_comparison() {  // :196-246
    size_t save = pos_;
    auto lhs = try_operand();
    if (!lhs) { pos_ = save; return nullptr; }
    if (eat("<=") || eat("\xe2\x89\xa4")) return event_le(*lhs, operand());
    if (eat("!=") || eat("\xe2\x89\xa0")) return Event::complement(event_eq(*lhs, operand()));
    if (eat(">=") || eat("\xe2\x89\xa5")) { Operand r = operand(); return event_le(r, *lhs); }
    if (peek() == '=') { pos_++; return event_eq(*lhs, operand()); }
    if (peek() == '<') { pos_++; return event_lt(*lhs, operand()); }
    if (peek() == '>') { pos_++; Operand r = operand(); return event_lt(r, *lhs); }
    if (keyword("not in") || eat("\xe2\x88\x89")) return Event::complement(event_in(*lhs, natural_list()));
    if (keyword("in") || eat("\xe2\x88\x88")) return event_in(*lhs, natural_list());
    pos_ = save;
    return nullptr;
  }
  EventP try_data_from_dist() {  // :248-253
    size_t save = pos_;
    auto data = try_natural();
    if (!data || peek() != '~') { pos_ = save; return nullptr; }
    pos_++;
    return Event::data_from_dist(*data, distribution());
  }
  EventP atomic_event() {  // :255-283
    size_t save = pos_;
    ws();
    if (peek() == '!' ) { pos_++; return Event::complement(atomic_event()); }
    if (peek_keyword("not") && !peek_keyword("not in")) { keyword("not"); return Event::complement(atomic_event()); }
    if (peek() == '(') {
      pos_++;
      EventP e = event();
      ws();
      expect_char(')');
      return e;
    }
    pos_ = save;
    if (EventP e = try_comparison()) return e;
    if (EventP e = try_data_from_dist()) return e;
    fail("expected event");
  }
  bool eat_connective(const char* word, const char* sym) {
    size_t save = pos_;
    ws();
    if (keyword(word) || eat(sym)) return true;
    pos_ = save;
    return false;
  }
  EventP event() {  // :285-316
    EventP e = atomic_event();
    std::vector<EventP> es{e};
    if (eat_connective("and", "&&")) {
      do es.push_back(event()); while (eat_connective("and", "&&"));
      return Event::intersection(std::move(es));
    }
    if (eat_connective("or", "||")) {
      do es.push_back(event()); while (eat_connective("or", "||"));
      return Event::disjunction(std::move(es));
    }
    return e;
  }

  // ---- distributions (:362-497) -------------------------------------------------------------------
  Distribution distribution() {
    std::string name = identifier();
    Distribution d;
    expect_char('(');
    if (name == "Dirac") {
      d.kind = DistKind::Dirac; d.p = pos_ratio();
    } else if (name == "Bernoulli") {
      if (auto p = try_pos_ratio()) { d.kind = DistKind::Bernoulli; d.p = *p; }
      else { d.kind = DistKind::BernoulliVarProb; d.var = expect_var(identifier()); }
    } else if (name == "Binomial" || name == "NegBinomial") {
      const bool neg = name == "NegBinomial";
      if (auto n = try_natural()) {
        d.kind = neg ? DistKind::NegBinomial : DistKind::Binomial; d.n = *n;
      } else {
        d.kind = neg ? DistKind::NegBinomialVarSuccesses : DistKind::BinomialVarTrials;
        d.var = expect_var(identifier());
      }
      expect_char(',');
      d.p = pos_ratio();
    } else if (name == "Categorical") {
      d.kind = DistKind::Categorical;
      d.rs.push_back(pos_ratio());
      while (peek() == ',') { pos_++; d.rs.push_back(pos_ratio()); }
    } else if (name == "Geometric") {
      d.kind = DistKind::Geometric; d.p = pos_ratio();
    } else if (name == "Poisson") {
      if (auto lam = try_pos_ratio()) {
        d.p = *lam;
        if (peek() == '*') { pos_++; d.kind = DistKind::PoissonVarRate; d.var = expect_var(identifier()); }
        else d.kind = DistKind::Poisson;
      } else {
        d.kind = DistKind::PoissonVarRate; d.p = PosRatio(1, 1); d.var = expect_var(identifier());
      }
    } else if (name == "UniformDisc") {
      d.kind = DistKind::Uniform; d.n = natural(); expect_char(','); d.m = natural();
    } else if (name == "Exponential") {
      d.kind = DistKind::Exponential; d.p = pos_ratio();
    } else if (name == "Gamma") {
      d.kind = DistKind::Gamma; d.p = pos_ratio(); expect_char(','); d.q = pos_ratio();
    } else if (name == "UniformCont") {
      d.kind = DistKind::UniformCont; d.p = pos_ratio(); expect_char(','); d.q = pos_ratio();
    } else {
      throw EvalError("Unknown distribution " + name);
    }
    expect_char(')');
    return d;
  }

  // ---- statements -----------------------------------------------------------------------------------
  Block block() {  // :582-594
    ws();
    expect_char('{');
    Block out;
    while (true) {
      ws();
      if (peek() == '}') { pos_++; return out; }
      if (eof()) fail("unterminated block");
      statement(out);
    }
  }
  Statement if_event() {  // :536-558 (the `if` keyword has been consumed)
    Statement st;
    st.kind = Statement::IfThenElse;
    st.cond = event();
    st.then_ = block();
    size_t save = pos_;
    ws();
    if (keyword("else")) {
      ws();
      if (keyword("if")) st.else_.push_back(if_event());
      else st.else_ = block();
    } else {
      pos_ = save;
    }
    return st;
  }
  void assign(Block& out) {  // :499-528
    std::string lhs = identifier();
    Statement st;
    if (peek() == '~' || starts_with("+~")) {
      st.kind = Statement::Sample;
      st.add_previous_value = peek() == '+';
      pos_ += st.add_previous_value ? 2 : 1;
      st.var = find_or_create_var(lhs);
      st.dist = distribution();
    } else if (eat("-=")) {
      st.kind = Statement::Decrement;
      st.offset = natural();
      st.var = find_or_create_var(lhs);
    } else {
      st.kind = Statement::Assign;
      if (eat(":=")) st.add_previous_value = false;
      else if (eat("+=")) st.add_previous_value = true;
      else fail("expected `~`, `+~`, `:=`, `+=` or `-=`");
      // [natural '*'] identifier ['+' natural]   |   natural      (:330-351)
      size_t save = pos_;
      bool parsed = false;
      {
        Natural factor = 1;
        size_t s2 = pos_;
        auto f = try_natural();
        if (f && peek() == '*') { pos_++; factor = *f; } else pos_ = s2;
        if (auto id = try_identifier()) {
          st.has_addend = true;
          st.addend_factor = factor;
          st.addend_var = expect_var(*id);
          if (peek() == '+') { pos_++; st.offset = natural(); }
          parsed = true;
        }
      }
      if (!parsed) {
        pos_ = save;
        st.offset = natural();
      }
      st.var = find_or_create_var(lhs);
    }
    semicolon();
    out.push_back(std::move(st));
  }
  void statement(Block& out) {  // :596-620
    ws();
    if (keyword("normalize")) {
      Statement st;
      st.kind = Statement::Normalize;
      while (true) {
        size_t save = pos_;
        ws();
        if (peek() == '{') { pos_ = save; break; }
        st.given_vars.push_back(expect_var(identifier()));
      }
      st.then_ = block();
      out.push_back(std::move(st));
    } else if (keyword("if")) {
      out.push_back(if_event());
    } else if (keyword("observe")) {
      Statement st;
      st.kind = Statement::IfThenElse;
      st.cond = event();
      semicolon();
      Statement f;
      f.kind = Statement::Fail;
      st.else_.push_back(f);
      out.push_back(std::move(st));
    } else if (keyword("loop")) {
      Natural count = natural();
      Block body = block();
      for (Natural i = 0; i < count; i++) out.insert(out.end(), body.begin(), body.end());
    } else if (keyword("while")) {
      Statement st;
      st.kind = Statement::While;
      st.cond = event();
      size_t save = pos_;
      ws();
      if (keyword("unroll")) st.unroll = natural();
      else pos_ = save;
      st.then_ = block();
      out.push_back(std::move(st));
    } else if (keyword("fail")) {
      semicolon();
      Statement st;
      st.kind = Statement::Fail;
      out.push_back(st);
    } else {
      assign(out);
    }
    ws();
  }
};

inline Program parse_program(const std::string& text) { return Parser(text).parse_program(); }

}  // namespace gfe
