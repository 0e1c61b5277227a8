// Stand-alone FP64 register-pattern probes (developer tool; not part of libgenfer_taylor).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes tools/probes.cu && tools/probes
// Each kernel is a pure-register DFMA pattern, run on 148*CTAS CTAs of 128 threads, reporting TFLOP/s.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

// a_i = fma(a_i, m, c)
__global__ void __launch_bounds__(128) p_chain8(int iters, const double* src, double* sink) {
  double a[8]; double m = src[64], c = src[65];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = src[(threadIdx.x + i) & 63];
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 16; r++)
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = fma(a[i], m, c);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += a[i];
  if (s == 12345.678) sink[0] = s;
}
// a_i = fma(b_i, m, a_i): 2 fresh operands
template <int N> __global__ void __launch_bounds__(128) p_two(int iters, const double* src, double* sink) {
  double a[N], b[N]; double m = src[64];
#pragma unroll
  for (int i = 0; i < N; i++) { a[i] = src[(threadIdx.x + i) & 63]; b[i] = src[(threadIdx.x * 3 + i) & 63]; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int i = 0; i < N; i++) a[i] = fma(b[i], m, a[i]);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < N; i++) s += a[i];
  if (s == 12345.678) sink[0] = s;
}
// a_i = fma(b_i, c_i, a_i): 3 fresh operands
template <int N> __global__ void __launch_bounds__(128) p_three(int iters, const double* src, double* sink) {
  double a[N], b[N], c[N];
#pragma unroll
  for (int i = 0; i < N; i++) { a[i] = src[(threadIdx.x + i) & 63]; b[i] = src[(threadIdx.x * 3 + i) & 63]; c[i] = src[(threadIdx.x * 5 + i) & 63]; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int i = 0; i < N; i++) a[i] = fma(b[i], c[i], a[i]);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < N; i++) s += a[i];
  if (s == 12345.678) sink[0] = s;
}
// outer product MxN: z[i][j] += x[i]*y[j]
template <int M, int N> __global__ void __launch_bounds__(128) p_outer(int iters, const double* src, double* sink) {
  double x[M], y[N], z[M][N];
#pragma unroll
  for (int i = 0; i < M; i++) x[i] = src[(threadIdx.x + 7 * i) & 63];
#pragma unroll
  for (int j = 0; j < N; j++) y[j] = src[(threadIdx.x * 3 + 5 * j + 1) & 63];
#pragma unroll
  for (int i = 0; i < M; i++)
#pragma unroll
    for (int j = 0; j < N; j++) z[i][j] = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int i = 0; i < M; i++)
#pragma unroll
        for (int j = 0; j < N; j++) z[i][j] = fma(x[i], y[j], z[i][j]);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < M; i++)
#pragma unroll
    for (int j = 0; j < N; j++) s += z[i][j];
  if (s == 12345.678) sink[0] = s;
}
// truncated row convolution, L=16: z[k] += x[j]*y[k-j]; ORDER 0: j outer, 1: k outer
template <int L, int ORDER> __global__ void __launch_bounds__(128) p_rowconv(int iters, const double* src, double* sink) {
  double x[L], y[L], z[L];
#pragma unroll
  for (int i = 0; i < L; i++) { x[i] = src[(threadIdx.x + 7 * i) & 63]; y[i] = src[(threadIdx.x * 3 + 5 * i + 1) & 63]; z[i] = 0; }
  for (int it = 0; it < iters; it++) {
    if (ORDER == 0) {
#pragma unroll
      for (int j = 0; j < L; j++)
#pragma unroll
        for (int k = j; k < L; k++) z[k] = fma(x[j], y[k - j], z[k]);
    } else {
#pragma unroll
      for (int k = 0; k < L; k++)
#pragma unroll
        for (int j = 0; j <= k; j++) z[k] = fma(x[j], y[k - j], z[k]);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < L; i++) s += z[i];
  if (s == 12345.678) sink[0] = s;
}
// full (untruncated) LxL convolution into 2L-1 outputs
template <int L> __global__ void __launch_bounds__(128) p_fullconv(int iters, const double* src, double* sink) {
  double x[L], y[L], z[2 * L];
#pragma unroll
  for (int i = 0; i < L; i++) { x[i] = src[(threadIdx.x + 7 * i) & 63]; y[i] = src[(threadIdx.x * 3 + 5 * i + 1) & 63]; }
#pragma unroll
  for (int i = 0; i < 2 * L; i++) z[i] = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < L; j++)
#pragma unroll
      for (int k = 0; k < L; k++) z[j + k] = fma(x[j], y[k], z[j + k]);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 2 * L; i++) s += z[i];
  if (s == 12345.678) sink[0] = s;
}
// 2x2 block of row convolutions into 3 accumulator rows (the k_mul_tiled22 inner pattern)
__global__ void __launch_bounds__(128) p_block22(int iters, const double* src, double* sink) {
  double xa[16], xb[16], y0[16], y1[16], z0[16], z1[16], z2[16];
#pragma unroll
  for (int i = 0; i < 16; i++) {
    xa[i] = src[(threadIdx.x + 7 * i) & 63]; xb[i] = src[(threadIdx.x + 3 * i + 2) & 63];
    y0[i] = src[(threadIdx.x * 3 + 2 * i + 1) & 63]; y1[i] = src[(threadIdx.x * 7 + 5 * i + 3) & 63];
    z0[i] = z1[i] = z2[i] = 0.0;
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < 16; j++)
#pragma unroll
      for (int k = j; k < 16; k++) {
        z0[k] = fma(xa[j], y0[k - j], z0[k]);
        z1[k] = fma(xa[j], y1[k - j], z1[k]);
        z1[k] = fma(xb[j], y0[k - j], z1[k]);
        z2[k] = fma(xb[j], y1[k - j], z2[k]);
      }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += z0[i] + z1[i] + z2[i];
  if (s == 12345.678) sink[0] = s;
}
// one x row against NB y rows into NB accumulator rows, L = 16 (x reused across rows: candidate 1xNB blocking)
template <int NB> __global__ void __launch_bounds__(128) p_block1n(int iters, const double* src, double* sink) {
  double x[16], y[NB][16], z[NB][16];
#pragma unroll
  for (int i = 0; i < 16; i++) {
    x[i] = src[(threadIdx.x + 7 * i) & 63];
#pragma unroll
    for (int b = 0; b < NB; b++) { y[b][i] = src[(threadIdx.x * (3 + 2 * b) + 5 * i + b) & 63]; z[b][i] = 0; }
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < 16; j++)
#pragma unroll
      for (int k = j; k < 16; k++)
#pragma unroll
        for (int b = 0; b < NB; b++) z[b][k] = fma(x[j], y[b][k - j], z[b][k]);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; i++)
#pragma unroll
    for (int b = 0; b < NB; b++) s += z[b][i];
  if (s == 12345.678) sink[0] = s;
}

template <class K> int run(const char* name, K kern, double dfma_per_iter, int iters, int ctas_per_sm, const double* src, double* sink) {
  int dev_sms = 148;
  cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
  int grid = dev_sms * ctas_per_sm;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaEventRecord(e0));
    kern<<<grid, 128>>>(iters, src, sink);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
  double fl = (double)grid * 128 * iters * dfma_per_iter * 2.0;
  printf("%-28s ctas/sm %d regs %3d  %7.2f TF/s  %8.3f ms\n", name, ctas_per_sm, fa.numRegs, fl / (best * 1e-3) / 1e12, best);
  return 0;
}

int main() {
  double h[128];
  for (int i = 0; i < 128; i++) h[i] = 1.0 + 1e-3 * i;
  h[64] = 1.0000001; h[65] = 1e-9;
  double *src, *sink;
  CK(cudaMalloc(&src, sizeof(h))); CK(cudaMalloc(&sink, 64 * 8));
  CK(cudaMemcpy(src, h, sizeof(h), cudaMemcpyHostToDevice));
  for (int c : {2, 4}) {
    run("chain8 (a=fma(a,m,c))", p_chain8, 128, 8192, c, src, sink);
    run("two<16> (a=fma(b,m,a))", p_two<16>, 128, 8192, c, src, sink);
    run("two<32>", p_two<32>, 256, 4096, c, src, sink);
    run("three<16> (a=fma(b,c,a))", p_three<16>, 128, 8192, c, src, sink);
    run("three<24>", p_three<24>, 192, 4096, c, src, sink);
    run("outer<4,4>", p_outer<4, 4>, 64, 16384, c, src, sink);
    run("outer<8,8>", p_outer<8, 8>, 256, 4096, c, src, sink);
    run("outer<4,16>", p_outer<4, 16>, 256, 4096, c, src, sink);
    run("outer<2,16>", p_outer<2, 16>, 128, 8192, c, src, sink);
    run("rowconv<16> j-outer", p_rowconv<16, 0>, 136, 8192, c, src, sink);
    run("rowconv<16> k-outer", p_rowconv<16, 1>, 136, 8192, c, src, sink);
    run("rowconv<8> j-outer", p_rowconv<8, 0>, 36, 16384, c, src, sink);
    run("rowconv<32> j-outer", p_rowconv<32, 0>, 528, 2048, c, src, sink);
    run("fullconv<8>", p_fullconv<8>, 64, 16384, c, src, sink);
    run("fullconv<16>", p_fullconv<16>, 256, 4096, c, src, sink);
    run("block22", p_block22, 544, 2048, c, src, sink);
    run("block1n<2>", p_block1n<2>, 272, 4096, c, src, sink);
    run("block1n<3>", p_block1n<3>, 408, 2048, c, src, sink);
    run("block1n<4>", p_block1n<4>, 544, 2048, c, src, sink);
  }
  return 0;
}
