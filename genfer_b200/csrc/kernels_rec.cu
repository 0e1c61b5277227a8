// Degree-recursive kernels: power-series division, exp_1d / log_1d, and small helpers used by the
// panel recurrences of exp / log (multivariate_taylor.rs:1162-1386).  "One axis sequential, all
// orthogonal axes data-parallel."
#include "kernels.cuh"

namespace gtp {

// ------------------------------------------------------------------------------------------
// Division by a divisor that varies along exactly one axis (the common p/(1-q*v) case,
// semantics/gf.rs:465-519).  Views: x (outer, x_len, inner), r (outer, r_len, inner), y (y_len).
// Per lane the reference computes (div :1170-1191 with the inner `mul` collapsing to x*y):
//     acc = 0;  for j = max(0,k+1-y_len) .. k-1:  acc = acc + r[j]*y[k-j]
//     r[k] = ((-acc) + x[k]) / y[0]          (x[k] absent for k >= x_len)
// with separately rounded multiply and add -- reproduced here bit for bit.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_div_axis_lanes(const double* __restrict__ x, const double* __restrict__ y,
                                                       double* r, u64 outer, u64 inner, u64 x_len, u64 y_len, u64 r_len) {
  const u64 total = outer * inner;
  const u64 gstride = (u64)gridDim.x * blockDim.x;
  const double y0 = y[0];
  for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gstride) {
    u64 o = t / inner, i = t - o * inner;
    const double* xp = x + o * x_len * inner + i;
    double* rp = r + o * r_len * inner + i;
    for (u64 k = 0; k < r_len; k++) {
      double acc = 0.0;
      u64 lo = (k + 1 > y_len) ? k + 1 - y_len : 0;
      for (u64 j = lo; j < k; j++) acc = __dadd_rn(acc, __dmul_rn(rp[j * inner], y[k - j]));
      double cur = -acc;
      if (k < x_len) cur = __dadd_rn(cur, xp[k * inner]);
      rp[k * inner] = __ddiv_rn(cur, y0);
    }
  }
}

// Few lanes, long axis (e.g. the 1-D population model, D ~ 516): one block per lane, column
// oriented -- once r[j] is final every thread adds r[j]*y[k-j] to its own pending acc[k].  The
// adds reach each acc[k] in ascending j, i.e. in the reference's order, so this is bit-exact too.
constexpr int DIV_COL_THREADS = 256;
constexpr int DIV_COL_PER_THREAD = 8;
__global__ void __launch_bounds__(DIV_COL_THREADS) k_div_axis_column(const double* __restrict__ x, const double* __restrict__ y,
                                                                    double* r, u64 inner, u64 x_len, u64 y_len, u64 r_len) {
  __shared__ double s_rj;
  const u64 lane = blockIdx.x;
  const u64 o = lane / inner, i = lane - o * inner;
  const double* xp = x + o * x_len * inner + i;
  double* rp = r + o * r_len * inner + i;
  const double y0 = y[0];
  double acc[DIV_COL_PER_THREAD];
#pragma unroll
  for (int q = 0; q < DIV_COL_PER_THREAD; q++) acc[q] = 0.0;
  for (u64 j = 0; j < r_len; j++) {
    // owner of coefficient j finalises it
    if ((j % DIV_COL_THREADS) == threadIdx.x) {
      double a = 0.0;
#pragma unroll
      for (int q = 0; q < DIV_COL_PER_THREAD; q++)
        if ((u64)q == j / DIV_COL_THREADS) a = acc[q];
      double cur = -a;
      if (j < x_len) cur = __dadd_rn(cur, xp[j * inner]);
      double rj = __ddiv_rn(cur, y0);
      rp[j * inner] = rj;
      s_rj = rj;
    }
    __syncthreads();
    const double rj = s_rj;
#pragma unroll
    for (int q = 0; q < DIV_COL_PER_THREAD; q++) {
      u64 k = (u64)q * DIV_COL_THREADS + threadIdx.x;
      if (k > j && k < r_len && k - j < y_len) acc[q] = __dadd_rn(acc[q], __dmul_rn(rj, y[k - j]));
    }
    __syncthreads();
  }
}

void launch_div_axis(Ctx& ctx, const double* x, const double* y, double* r, u64 outer, u64 inner, u64 x_len,
                     u64 y_len, u64 r_len, const Shape&, const Shape&, int) {
  u64 lanes = outer * inner;
  if (lanes == 0 || r_len == 0) return;
  if (lanes < (u64)ctx.sm_count * 8 && r_len >= 64 && r_len <= (u64)DIV_COL_THREADS * DIV_COL_PER_THREAD) {
    GTP_LAUNCH(ctx, k_div_axis_column, (unsigned)lanes, DIV_COL_THREADS, 0, x, y, r, inner, x_len, y_len, r_len);
  } else {
    int grid = (int)std::max<u64>(1, std::min<u64>((lanes + 127) / 128, (u64)ctx.sm_count * 32));
    GTP_LAUNCH(ctx, k_div_axis_lanes, grid, 128, 0, x, y, r, outer, inner, x_len, y_len, r_len);
  }
}

// ------------------------------------------------------------------------------------------
// General N-D division (divisor varies along >= 2 axes; rare).  r[k] depends on r[j], j <= k, j != k,
// so all k of one total degree |k| = t are independent: one launch per wavefront.
//     r[k] = (x[k] - sum_{m != 0, m <= k, m < ys} r[k-m]*y[m]) / y[0]
// Same value as the reference's nested recursion in exact arithmetic; the summation order differs
// (tolerance-checked, not bit-exact).
// ------------------------------------------------------------------------------------------
struct DivGP {
  int ndim;
  unsigned xs[MAXD], ys[MAXD], rs[MAXD];
  long long xstr[MAXD], ystr[MAXD], rstr[MAXD];
  u64 outer_total;  // prod(rs[0..ndim-2])
  const double* x;
  const double* y;
  double* r;
};
__global__ void __launch_bounds__(128) k_div_wavefront(const DivGP p, unsigned t) {
  const u64 gstride = (u64)gridDim.x * blockDim.x;
  const int nd = p.ndim;
  for (u64 lin = (u64)blockIdx.x * blockDim.x + threadIdx.x; lin < p.outer_total; lin += gstride) {
    unsigned k[MAXD], m[MAXD], mhi[MAXD];
    u64 rem = lin;
    unsigned sum = 0;
    for (int d = nd - 2; d >= 0; --d) {
      k[d] = (unsigned)(rem % p.rs[d]);
      rem /= p.rs[d];
      sum += k[d];
    }
    if (sum > t) continue;
    k[nd - 1] = t - sum;
    if (k[nd - 1] >= p.rs[nd - 1]) continue;
    long long ro = 0, xo = 0;
    bool in_x = true;
    for (int d = 0; d < nd; d++) {
      ro += (long long)k[d] * p.rstr[d];
      xo += (long long)k[d] * p.xstr[d];
      in_x &= k[d] < p.xs[d];
      mhi[d] = min(k[d], p.ys[d] - 1);  // inclusive upper bound of m[d]
      m[d] = 0;
    }
    double s = 0.0;
    // odometer over m in [0, mhi], skipping m = 0
    long long mo_r = 0, mo_y = 0;  // offsets of r[k-m] (relative to ro) and y[m]
    while (true) {
      int d = nd - 1;
      for (; d >= 0; --d) {
        if (m[d] < mhi[d]) {
          m[d]++;
          mo_r -= p.rstr[d];
          mo_y += p.ystr[d];
          break;
        }
        mo_r += (long long)m[d] * p.rstr[d];
        mo_y -= (long long)m[d] * p.ystr[d];
        m[d] = 0;
      }
      if (d < 0) break;
      s = fma(p.r[ro + mo_r], p.y[mo_y], s);
    }
    double num = in_x ? p.x[xo] - s : -s;
    p.r[ro] = num / p.y[0];
  }
}
void launch_div_general(Ctx& ctx, const double* x, const Shape& xs, const double* y, const Shape& ys, double* r,
                        const Shape& rs) {
  const int nd = (int)rs.size();
  GTP_CHECK(nd >= 1 && nd <= MAXD, GTP_ERR_ARG, "div_general: bad ndim");
  DivGP p;
  memset(&p, 0, sizeof(p));
  p.ndim = nd;
  Shape xst(nd, 1), yst(nd, 1), rst(nd, 1);
  for (int i = nd - 2; i >= 0; --i) {
    xst[i] = xst[i + 1] * xs[i + 1];
    yst[i] = yst[i + 1] * ys[i + 1];
    rst[i] = rst[i + 1] * rs[i + 1];
  }
  u64 tmax = 0;
  p.outer_total = 1;
  for (int d = 0; d < nd; d++) {
    p.xs[d] = (unsigned)xs[d];
    p.ys[d] = (unsigned)ys[d];
    p.rs[d] = (unsigned)rs[d];
    p.xstr[d] = (long long)xst[d];
    p.ystr[d] = (long long)yst[d];
    p.rstr[d] = (long long)rst[d];
    tmax += rs[d] - 1;
    if (d < nd - 1) p.outer_total *= rs[d];
  }
  p.x = x;
  p.y = y;
  p.r = r;
  int grid = (int)std::max<u64>(1, std::min<u64>((p.outer_total + 127) / 128, (u64)ctx.sm_count * 16));
  for (u64 t = 0; t <= tmax; t++) GTP_LAUNCH(ctx, k_div_wavefront, grid, 128, 0, p, (unsigned)t);
}

// ------------------------------------------------------------------------------------------
// exp_1d (:1271-1283) and log_1d (:1319-1333): sequential in k, the inner sum spread over one
// block (tree reduction => tolerance-level, not bit-level, agreement with the reference order).
// ------------------------------------------------------------------------------------------
constexpr int REC_THREADS = 256;
__device__ __forceinline__ double block_sum(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < 32) {
    t = (threadIdx.x < REC_THREADS / 32) ? sh[threadIdx.x] : 0.0;
    for (int o = 4; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
  }
  return t;  // valid in thread 0
}
// `seed`: exp / log of the constant term evaluated by the host's libm when the host knows the constant term (bit-identical
// to the reference's f64::exp / ln, number/f64.rs:53-62); otherwise CUDA's exp / log (<= 1 ulp) on the device.
__global__ void __launch_bounds__(REC_THREADS) k_exp_1d(const double* __restrict__ x, u64 xlen, double* r, u64 n,
                                                       double* scratch, int has_seed, double seed) {
  __shared__ double sh[REC_THREADS / 32];
  // scratch[j] = x[j] * j   (the reference's `xs[j] * T::from(j)` factor)
  for (u64 j = threadIdx.x; j < xlen; j += REC_THREADS) scratch[j] = __dmul_rn(x[j], (double)(unsigned)j);
  if (threadIdx.x == 0) r[0] = has_seed ? seed : exp(x[0]);
  __syncthreads();
  for (u64 k = 1; k < n; k++) {
    u64 hi = xlen < k + 1 ? xlen : k + 1;
    double part = 0.0;
    if (hi <= 32) {  // short sums: thread 0 in reference order (bit-exact for low-degree arguments)
      if (threadIdx.x == 0) {
        double sum = 0.0;
        for (u64 j = 1; j < hi; j++) sum = __dadd_rn(sum, __dmul_rn(scratch[j], r[k - j]));
        r[k] = __ddiv_rn(sum, (double)(unsigned)k);
      }
      __syncthreads();
      continue;
    }
    for (u64 j = 1 + threadIdx.x; j < hi; j += REC_THREADS) part = fma(scratch[j], r[k - j], part);
    double sum = block_sum(part, sh);
    if (threadIdx.x == 0) r[k] = __ddiv_rn(sum, (double)(unsigned)k);
    __syncthreads();
  }
}
__global__ void __launch_bounds__(REC_THREADS) k_log_1d(const double* __restrict__ x, u64 xlen, double* r, u64 n,
                                                       double* scratch, int has_seed, double seed) {
  __shared__ double sh[REC_THREADS / 32];
  // scratch[j] = r[j] * j, filled as coefficients become final
  if (threadIdx.x == 0) {
    r[0] = has_seed ? seed : log(x[0]);
    scratch[0] = 0.0;
  }
  __syncthreads();
  const double x0 = x[0];
  for (u64 k = 1; k < n; k++) {
    u64 lo = (k + 1 > xlen) ? k + 1 - xlen : 0;
    if (lo < 1) lo = 1;
    double sum;
    if (k - lo <= 32) {
      sum = 0.0;
      if (threadIdx.x == 0)
        for (u64 j = lo; j < k; j++) sum = __dadd_rn(sum, __dmul_rn(__dmul_rn(x[k - j], r[j]), (double)(unsigned)j));
    } else {
      double part = 0.0;
      for (u64 j = lo + threadIdx.x; j < k; j += REC_THREADS) part = fma(x[k - j], scratch[j], part);
      sum = block_sum(part, sh);
    }
    if (threadIdx.x == 0) {
      double xk = k < xlen ? x[k] : 0.0;
      double kk = (double)(unsigned)k;
      double v = __ddiv_rn(__ddiv_rn(__dsub_rn(__dmul_rn(xk, kk), sum), x0), kk);
      r[k] = v;
      scratch[k] = __dmul_rn(v, kk);
    }
    __syncthreads();
  }
}
void launch_exp_1d(Ctx& ctx, const double* x, u64 xlen, double* r, u64 n, const double* seed) {
  if (n == 0) return;
  BufP scratch = ctx.alloc(std::max<u64>(xlen, 1));
  GTP_LAUNCH(ctx, k_exp_1d, 1, REC_THREADS, 0, x, xlen, r, n, scratch->d, seed ? 1 : 0, seed ? *seed : 0.0);
}
void launch_log_1d(Ctx& ctx, const double* x, u64 xlen, double* r, u64 n, const double* seed) {
  if (n == 0) return;
  BufP scratch = ctx.alloc(std::max<u64>(n, 1));
  GTP_LAUNCH(ctx, k_log_1d, 1, REC_THREADS, 0, x, xlen, r, n, scratch->d, seed ? 1 : 0, seed ? *seed : 0.0);
}

__global__ void k_scalar_fn(int fn, const double* x, double* r, int has_seed, double seed) {
  if (threadIdx.x == 0 && blockIdx.x == 0) r[0] = has_seed ? seed : ((fn == 0) ? exp(x[0]) : log(x[0]));
}
void launch_scalar_fn(Ctx& ctx, int fn, const double* x, double* r, const double* seed) {
  GTP_LAUNCH(ctx, k_scalar_fn, 1, 32, 0, fn, x, r, seed ? 1 : 0, seed ? *seed : 0.0);
}

__global__ void __launch_bounds__(256) k_scale_const(const double* __restrict__ in, double* __restrict__ out, u64 n, double k, int divide) {
  u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = divide ? __ddiv_rn(in[i], k) : __dmul_rn(in[i], k);
}
void launch_scale_const(Ctx& ctx, const double* in, double* out, u64 n, double k, bool divide) {
  if (n == 0) return;
  int grid = (int)std::max<u64>(1, std::min<u64>((n + 255) / 256, (u64)ctx.sm_count * 16));
  GTP_LAUNCH(ctx, k_scale_const, grid, 256, 0, in, out, n, k, divide ? 1 : 0);
}

__global__ void __launch_bounds__(256) k_scale_rows(const double* __restrict__ in, double* __restrict__ out, u64 rows, u64 inner, u64 j0) {
  u64 total = rows * inner;
  u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    u64 j = i / inner;
    out[i] = __dmul_rn(in[i], (double)(unsigned)(j + j0));
  }
}
void launch_scale_rows(Ctx& ctx, const double* in, double* out, u64 rows, u64 inner, u64 j0, bool) {
  if (rows * inner == 0) return;
  int grid = (int)std::max<u64>(1, std::min<u64>((rows * inner + 255) / 256, (u64)ctx.sm_count * 16));
  GTP_LAUNCH(ctx, k_scale_rows, grid, 256, 0, in, out, rows, inner, j0);
}

}  // namespace gtp
