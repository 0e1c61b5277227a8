// Developer probe: how does the rate of tiny DEPENDENT kernel launches scale with host threads x streams on this box?
// Each thread owns a stream and launches N kernels that each touch a 512-element vector produced by the previous one.
// build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tools/mt_launch_probe tools/mt_launch_probe.cu -lpthread
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
__global__ void k_step(const double* in, double* out, int n, double s) {
  int i = threadIdx.x + blockIdx.x * blockDim.x;
  if (i < n) out[i] = in[i] * s + 1.0;
}
int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 20000;
  cudaFree(0);
  for (int T : {1, 2, 4, 8, 12, 16}) {
    std::vector<std::thread> th;
    auto t0 = std::chrono::steady_clock::now();
    for (int t = 0; t < T; t++)
      th.emplace_back([=] {
        cudaStream_t s;
        cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
        double *a, *b;
        cudaMalloc(&a, 4096); cudaMalloc(&b, 4096);
        cudaMemsetAsync(a, 0, 4096, s);
        for (int i = 0; i < N; i++) {
          k_step<<<1, 512, 0, s>>>(a, b, 512, 1.0000001);
          double* tmp = a; a = b; b = tmp;
        }
        cudaStreamSynchronize(s);
        cudaFree(a); cudaFree(b); cudaStreamDestroy(s);
      });
    for (auto& x : th) x.join();
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("threads %2d: %d launches each, %.3f s total, %.2f us per launch per thread, %.2f M launches/s aggregate\n", T, N, dt, dt / N * 1e6, T * (double)N / dt / 1e6);
  }
  return 0;
}
