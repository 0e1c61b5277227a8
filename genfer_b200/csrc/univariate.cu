// Device kernels for the univariate TaylorExpansion<F64> (src/univariate_taylor.rs).  Series are
// short (order <= limit <= 1000, src/main.rs:30), so every op is one small launch.  Products,
// quotients and the element-wise family keep the reference's operation order with separately
// rounded multiply/add and are bit-identical to it; exp/log use a block reduction for long sums.
#include "kernels.cuh"

namespace gtp {

// Mul :376-386 -- result[k] = fold_{j=0..k}(sum + us[j]*ws[k-j]); one thread per k, same order.
__global__ void __launch_bounds__(128) k_uni_mul(const double* __restrict__ u, const double* __restrict__ w, double* r, u64 order) {
  u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= order) return;
  double sum = 0.0;
  for (u64 j = 0; j <= k; j++) sum = __dadd_rn(sum, __dmul_rn(u[j], w[k - j]));
  r[k] = sum;
}
void uni_mul(Ctx& ctx, const double* u, const double* w, double* r, u64 order) {
  if (order == 0) return;
  GTP_LAUNCH(ctx, k_uni_mul, (unsigned)((order + 127) / 128), 128, 0, u, w, r, order);
}

// Div :409-436 -- scale = 1/ws[0]; result[k] = scale * fold_{i<k}(init_k, sum - result[i]*ws[k-i]).
// Column oriented: when result[i] is final every pending k subtracts result[i]*ws[k-i]; each
// pending sum sees i ascending, i.e. the reference's fold order.
constexpr int UNI_T = 256;
constexpr int UNI_PER = 8;  // orders up to 2048
__global__ void __launch_bounds__(UNI_T) k_uni_div(const double* __restrict__ u, int u_const, const double* __restrict__ w,
                                                  double* r, u64 order) {
  __shared__ double s_ri;
  const double scale = __ddiv_rn(1.0, w[0]);
  double acc[UNI_PER];
#pragma unroll
  for (int q = 0; q < UNI_PER; q++) {
    u64 k = (u64)q * UNI_T + threadIdx.x;
    acc[q] = (k < order && !u_const) ? u[k] : 0.0;
  }
  for (u64 i = 0; i < order; i++) {
    if ((i % UNI_T) == threadIdx.x) {
      double a = 0.0;
#pragma unroll
      for (int q = 0; q < UNI_PER; q++)
        if ((u64)q == i / UNI_T) a = acc[q];
      double ri = (i == 0) ? __dmul_rn(u_const ? u[0] : a, scale) : __dmul_rn(scale, a);
      r[i] = ri;
      s_ri = ri;
    }
    __syncthreads();
    const double ri = s_ri;
#pragma unroll
    for (int q = 0; q < UNI_PER; q++) {
      u64 k = (u64)q * UNI_T + threadIdx.x;
      if (k > i && k < order) acc[q] = __dsub_rn(acc[q], __dmul_rn(ri, w[k - i]));
    }
    __syncthreads();
  }
}
void uni_div(Ctx& ctx, const double* u, bool u_const, const double* w, double* r, u64 order) {
  GTP_CHECK(order <= (u64)UNI_T * UNI_PER, GTP_ERR_ARG, "univariate order > 2048 not supported");
  if (order == 0) return;
  GTP_LAUNCH(ctx, k_uni_div, 1, UNI_T, 0, u, u_const ? 1 : 0, w, r, order);
}

__device__ __forceinline__ double uni_block_sum(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < 32) {
    t = (threadIdx.x < UNI_T / 32) ? sh[threadIdx.x] : 0.0;
    for (int o = 4; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
  }
  return t;
}
// exp :153-164 -- res[k] = (sum_{j=1..k} res[k-j]*coeffs[j]*j) / k
__global__ void __launch_bounds__(UNI_T) k_uni_exp(const double* __restrict__ c, double* r, u64 order) {
  __shared__ double sh[UNI_T / 32];
  if (threadIdx.x == 0) r[0] = exp(c[0]);
  __syncthreads();
  for (u64 k = 1; k < order; k++) {
    if (k <= 32) {
      if (threadIdx.x == 0) {
        double sum = 0.0;
        for (u64 j = 1; j <= k; j++) sum = __dadd_rn(sum, __dmul_rn(__dmul_rn(r[k - j], c[j]), (double)(unsigned)j));
        r[k] = __ddiv_rn(sum, (double)(unsigned)k);
      }
      __syncthreads();
      continue;
    }
    double part = 0.0;
    for (u64 j = 1 + threadIdx.x; j <= k; j += UNI_T) part = fma(__dmul_rn(r[k - j], c[j]), (double)(unsigned)j, part);
    double sum = uni_block_sum(part, sh);
    if (threadIdx.x == 0) r[k] = __ddiv_rn(sum, (double)(unsigned)k);
    __syncthreads();
  }
}
// log :172-185 -- res[k] = (coeffs[k]*k - sum_{j=1..k-1} coeffs[k-j]*res[j]*j) / coeffs[0] / k
__global__ void __launch_bounds__(UNI_T) k_uni_log(const double* __restrict__ c, double* r, u64 order) {
  __shared__ double sh[UNI_T / 32];
  if (threadIdx.x == 0) r[0] = log(c[0]);
  __syncthreads();
  const double c0 = c[0];
  for (u64 k = 1; k < order; k++) {
    double sum = 0.0;
    if (k <= 33) {
      if (threadIdx.x == 0)
        for (u64 j = 1; j < k; j++) sum = __dadd_rn(sum, __dmul_rn(__dmul_rn(c[k - j], r[j]), (double)(unsigned)j));
    } else {
      double part = 0.0;
      for (u64 j = 1 + threadIdx.x; j < k; j += UNI_T) part = fma(__dmul_rn(c[k - j], r[j]), (double)(unsigned)j, part);
      sum = uni_block_sum(part, sh);
    }
    if (threadIdx.x == 0) {
      double kk = (double)(unsigned)k;
      r[k] = __ddiv_rn(__ddiv_rn(__dsub_rn(__dmul_rn(c[k], kk), sum), c0), kk);
    }
    __syncthreads();
  }
}
void uni_exp(Ctx& ctx, const double* c, double* r, u64 order) {
  if (order) GTP_LAUNCH(ctx, k_uni_exp, 1, UNI_T, 0, c, r, order);
}
void uni_log(Ctx& ctx, const double* c, double* r, u64 order) {
  if (order) GTP_LAUNCH(ctx, k_uni_log, 1, UNI_T, 0, c, r, order);
}

// Element-wise family covering the Constant/Polynomial combinations of AddAssign :277-306,
// SubAssign :330-362, Mul-by-constant :370-375, Div-by-constant :403-408, Neg :308-319.
__global__ void __launch_bounds__(256) k_uni_ew(int op, const double* __restrict__ a, int a_b, const double* __restrict__ b, int b_b,
                                               double* r, u64 n) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x = a[a_b ? 0 : i];
  double y = b ? b[b_b ? 0 : i] : 0.0;
  double v;
  switch (op) {
    case 0: v = __dadd_rn(x, y); break;
    case 1: v = __dsub_rn(x, y); break;
    case 2: v = __dmul_rn(x, y); break;
    case 3: v = __ddiv_rn(x, y); break;
    case 4: v = -x; break;
    case 5: v = (i == 0) ? __dadd_rn(x, y) : x; break;          // coeffs[0] += rhs
    case 6: v = (i == 0) ? __dsub_rn(x, y) : x; break;          // coeffs[0] -= rhs
    default: v = (i == 0) ? __dadd_rn(-x, y) : -x; break;       // ws = -ws; ws[0] += c   (:340-344)
  }
  r[i] = v;
}
void uni_ew(Ctx& ctx, int op, const double* a, bool a_bcast, const double* b, bool b_bcast, double* r, u64 n) {
  if (n) GTP_LAUNCH(ctx, k_uni_ew, (unsigned)((n + 255) / 256), 256, 0, op, a, a_bcast ? 1 : 0, b, b_bcast ? 1 : 0, r, n);
}

// taylor_expansion_of_coeff :78-87
__global__ void k_uni_teoc(const double* __restrict__ in, double* out, u64 n, u64 len) {
  if (threadIdx.x || blockIdx.x) return;
  double factor = 1.0;
  if (len) out[0] = in[n];
  for (u64 k = 1; k < len; k++) {
    factor = __dmul_rn(factor, __ddiv_rn((double)(unsigned)(n + k), (double)(unsigned)k));
    out[k] = __dmul_rn(in[n + k], factor);
  }
}
void uni_teoc(Ctx& ctx, const double* in, double* out, u64 n, u64 len) { GTP_LAUNCH(ctx, k_uni_teoc, 1, 32, 0, in, out, n, len); }

// derivative :47-51 -- factorial(order) * coeffs[order]
__global__ void k_uni_fact(const double* __restrict__ in, u64 order, double* out) {
  if (threadIdx.x || blockIdx.x) return;
  double f = 1.0;
  for (u64 i = 1; i <= order; i++) f = __dmul_rn(f, (double)(unsigned)i);
  out[0] = __dmul_rn(f, in[order]);
}
void uni_factorial_times(Ctx& ctx, const double* in, u64 order, double* out) { GTP_LAUNCH(ctx, k_uni_fact, 1, 32, 0, in, order, out); }

}  // namespace gtp
