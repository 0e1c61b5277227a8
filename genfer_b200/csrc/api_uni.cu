// C ABI for the univariate TaylorExpansion<F64> (src/univariate_taylor.rs): enum dispatch on the
// host (Constant vs Polynomial), arithmetic on the device.
#include "kernels.cuh"

using namespace gtp;
using SerP = std::unique_ptr<gtu_series>;

namespace {

SerP new_series(Ctx& c, bool is_const, u64 n) {
  SerP s(new gtu_series());
  s->is_const = is_const;
  s->n = is_const ? 1 : n;
  s->buf = c.alloc(std::max<u64>(s->n, 1));
  return s;
}
SerP share(const gtu_series& a) { return SerP(new gtu_series(a)); }

template <class F> int wrap(gtp_ctx* ctx, F&& f) {
  try {
    if (ctx) GTP_CUDA(cudaSetDevice(ctx->device));
    f();
    return GTP_OK;
  } catch (const gtp::Error& e) {
    if (ctx) ctx->err = e.what();
    return e.code;
  } catch (const std::exception& e) {
    if (ctx) ctx->err = e.what();
    return GTP_ERR_ARG;
  }
}

// AddAssign :277-306 / SubAssign :330-362
SerP uni_addsub(Ctx& c, const gtu_series& self, const gtu_series& rhs, bool sub) {
  if (rhs.is_const) {
    // Constant+Constant, or coeffs[0] (+|-)= rhs
    SerP r = new_series(c, self.is_const, self.n);
    uni_ew(c, sub ? 6 : 5, self.ptr(), false, rhs.ptr(), true, r->buf->d, r->n);
    return r;
  }
  if (self.is_const) {
    // ws[0] += c          |   ws = -ws; ws[0] += c
    SerP r = new_series(c, false, rhs.n);
    uni_ew(c, sub ? 7 : 5, rhs.ptr(), false, self.ptr(), true, r->buf->d, r->n);
    return r;
  }
  u64 order = std::min(self.n, rhs.n);  // truncated to the shorter operand (:291, :347)
  SerP r = new_series(c, false, order);
  uni_ew(c, sub ? 1 : 0, self.ptr(), false, rhs.ptr(), false, r->buf->d, order);
  return r;
}

SerP uni_mul_op(Ctx& c, const gtu_series& a, const gtu_series& b) {  // :364-389
  if (a.is_const && b.is_const) {
    SerP r = new_series(c, true, 1);
    uni_ew(c, 2, a.ptr(), false, b.ptr(), false, r->buf->d, 1);
    return r;
  }
  if (a.is_const || b.is_const) {  // coeff *= c
    const gtu_series& p = a.is_const ? b : a;
    const gtu_series& k = a.is_const ? a : b;
    SerP r = new_series(c, false, p.n);
    uni_ew(c, 2, p.ptr(), false, k.ptr(), true, r->buf->d, p.n);
    return r;
  }
  u64 order = std::min(a.n, b.n);
  SerP r = new_series(c, false, order);
  uni_mul(c, a.ptr(), b.ptr(), r->buf->d, order);
  return r;
}

SerP uni_div_op(Ctx& c, const gtu_series& a, const gtu_series& b) {  // :397-439
  if (a.is_const && b.is_const) {
    SerP r = new_series(c, true, 1);
    uni_ew(c, 3, a.ptr(), false, b.ptr(), false, r->buf->d, 1);
    return r;
  }
  if (!a.is_const && b.is_const) {  // coeff /= c
    SerP r = new_series(c, false, a.n);
    uni_ew(c, 3, a.ptr(), false, b.ptr(), true, r->buf->d, a.n);
    return r;
  }
  u64 order = a.is_const ? b.n : std::min(a.n, b.n);
  SerP r = new_series(c, false, order);
  uni_div(c, a.ptr(), a.is_const, b.ptr(), r->buf->d, order);
  return r;
}

SerP uni_const(Ctx& c, double x) {
  SerP r = new_series(c, true, 1);
  launch_fill(c, r->buf->d, 1, x);
  return r;
}

}  // namespace

extern "C" {

int gtu_constant(gtp_ctx* c, double x, gtu_series** out) { return wrap(c, [&] { *out = uni_const(*c, x).release(); }); }
int gtu_from_coefficients(gtp_ctx* c, const double* xs, uint64_t n, gtu_series** out) {
  return wrap(c, [&] {
    GTP_CHECK(xs || n == 0, GTP_ERR_ARG, "null coefficients");
    SerP r = new_series(*c, false, n);
    if (n) GTP_CUDA(cudaMemcpyAsync(r->buf->d, xs, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    *out = r.release();
  });
}
int gtu_var(gtp_ctx* c, double x, uint64_t order, gtu_series** out) {  // :16-23
  return wrap(c, [&] {
    std::vector<double> v(order + 1, 0.0);
    if (v.size() > 1) v[1] = 1.0;
    v[0] = x;
    SerP r = new_series(*c, false, order + 1);
    GTP_CUDA(cudaMemcpyAsync(r->buf->d, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    c->sync();  // `v` is pageable and dies at scope exit
    *out = r.release();
  });
}
void gtu_free(gtp_ctx* c, gtu_series* s) {
  if (c) cudaSetDevice(c->device);
  delete s;
}
int gtu_is_constant(const gtu_series* s) { return s->is_const ? 1 : 0; }
uint64_t gtu_order(const gtu_series* s) { return s->is_const ? GTP_UNBOUNDED : s->n; }
int gtu_to_host(gtp_ctx* c, const gtu_series* s, double* out) {
  return wrap(c, [&] {
    if (s->n) GTP_CUDA(cudaMemcpyAsync(out, s->ptr(), s->n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    c->sync();
  });
}
int gtu_coeff(gtp_ctx* c, const gtu_series* s, uint64_t order, double* out) {  // :25-36
  return wrap(c, [&] {
    if (s->is_const && order != 0) { *out = 0.0; return; }
    GTP_CHECK(s->is_const || order < s->n, GTP_ERR_INDEX, "coeff: index out of bounds");
    GTP_CUDA(cudaMemcpyAsync(&c->rb_host->vals[0], s->ptr() + (s->is_const ? 0 : order), sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    c->sync();
    *out = c->rb_host->vals[0];
  });
}
int gtu_derivative(gtp_ctx* c, const gtu_series* s, uint64_t order, double* out) {  // :45-60
  return wrap(c, [&] {
    if (s->is_const) {
      if (order != 0) { *out = 0.0; return; }
      GTP_CUDA(cudaMemcpyAsync(&c->rb_host->vals[0], s->ptr(), sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    } else {
      GTP_CHECK(order < s->n, GTP_ERR_INDEX, "derivative: index out of bounds");
      uni_factorial_times(*c, s->ptr(), order, &c->rb_dev->vals[0]);
      GTP_CUDA(cudaMemcpyAsync(&c->rb_host->vals[0], &c->rb_dev->vals[0], sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    }
    c->sync();
    *out = c->rb_host->vals[0];
  });
}
int gtu_add(gtp_ctx* c, const gtu_series* a, const gtu_series* b, gtu_series** out) { return wrap(c, [&] { *out = uni_addsub(*c, *a, *b, false).release(); }); }
int gtu_sub(gtp_ctx* c, const gtu_series* a, const gtu_series* b, gtu_series** out) { return wrap(c, [&] { *out = uni_addsub(*c, *a, *b, true).release(); }); }
int gtu_mul(gtp_ctx* c, const gtu_series* a, const gtu_series* b, gtu_series** out) { return wrap(c, [&] { *out = uni_mul_op(*c, *a, *b).release(); }); }
int gtu_div(gtp_ctx* c, const gtu_series* a, const gtu_series* b, gtu_series** out) { return wrap(c, [&] { *out = uni_div_op(*c, *a, *b).release(); }); }
int gtu_neg(gtp_ctx* c, const gtu_series* a, gtu_series** out) {  // :308-319
  return wrap(c, [&] {
    SerP r = new_series(*c, a->is_const, a->n);
    uni_ew(*c, 4, a->ptr(), false, nullptr, false, r->buf->d, r->n);
    *out = r.release();
  });
}
int gtu_exp(gtp_ctx* c, const gtu_series* a, gtu_series** out) {  // :151-168
  return wrap(c, [&] {
    SerP r = new_series(*c, a->is_const, a->n);
    if (a->is_const) launch_scalar_fn(*c, 0, a->ptr(), r->buf->d);
    else uni_exp(*c, a->ptr(), r->buf->d, a->n);
    *out = r.release();
  });
}
int gtu_log(gtp_ctx* c, const gtu_series* a, gtu_series** out) {  // :170-189
  return wrap(c, [&] {
    SerP r = new_series(*c, a->is_const, a->n);
    if (a->is_const) launch_scalar_fn(*c, 1, a->ptr(), r->buf->d);
    else uni_log(*c, a->ptr(), r->buf->d, a->n);
    *out = r.release();
  });
}
int gtu_pow(gtp_ctx* c, const gtu_series* a, uint32_t e, gtu_series** out) {  // :192-203
  return wrap(c, [&] {
    SerP res = uni_const(*c, 1.0);
    SerP base = share(*a);
    while (e > 0) {
      if (e & 1) res = uni_mul_op(*c, *res, *base);
      base = uni_mul_op(*c, *base, *base);
      e >>= 1;
    }
    *out = res.release();
  });
}
int gtu_subst(gtp_ctx* c, const gtu_series* a, const gtu_series* s, gtu_series** out) {  // :93-115
  return wrap(c, [&] {
    if (a->is_const) { *out = share(*a).release(); return; }
    if (!s->is_const) GTP_CHECK(s->n == a->n, GTP_ERR_INDEX, "Substitution must have the same order");
    SerP res = uni_const(*c, 0.0);
    for (u64 i = a->n; i-- > 0;) {
      SerP p = uni_mul_op(*c, *res, *s);
      gtu_series ci;  // Constant(c_i) viewed in place
      ci.is_const = true;
      ci.n = 1;
      auto b = std::make_shared<Buf>();
      b->d = const_cast<double*>(a->ptr() + i);
      b->n = 1;
      b->owned = false;
      ci.buf = b;
      res = uni_addsub(*c, *p, ci, false);
    }
    *out = res.release();
  });
}
int gtu_taylor_expansion_of_coeff(gtp_ctx* c, const gtu_series* a, uint64_t n, gtu_series** out) {  // :69-89
  return wrap(c, [&] {
    if (a->is_const) {
      if (n == 0) {  // sic: the reference returns Constant(c.exp()) here (:73)
        SerP r = new_series(*c, true, 1);
        launch_scalar_fn(*c, 0, a->ptr(), r->buf->d);
        *out = r.release();
      } else {
        *out = uni_const(*c, 0.0).release();
      }
      return;
    }
    GTP_CHECK(n <= a->n, GTP_ERR_INDEX, "taylor_expansion_of_coeff: slice start out of range");
    SerP r = new_series(*c, false, a->n - n);
    uni_teoc(*c, a->ptr(), r->buf->d, n, a->n - n);
    *out = r.release();
  });
}
int gtu_eq(gtp_ctx* c, const gtu_series* a, const gtu_series* b, int* out) {
  return wrap(c, [&] {
    *out = 0;
    if (a->is_const != b->is_const || a->n != b->n) return;
    if (a->n == 0) { *out = 1; return; }
    launch_eq(*c, a->ptr(), b->ptr(), a->n, c->rb_dev);
    GTP_CUDA(cudaMemcpyAsync(&c->rb_host->flag, &c->rb_dev->flag, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
    c->sync();
    *out = c->rb_host->flag ? 1 : 0;
  });
}

}  // extern "C"
