// Grid-wide barrier for persistent kernels launched with cudaLaunchCooperativeKernel (all CTAs co-resident), and the
// L2 load used for data other CTAs wrote earlier in the same launch.
#pragma once
#include <cuda_runtime.h>

namespace gtp {

__device__ __forceinline__ double ldcg(const double* p) { return __ldcg(p); }

// `bar` is a zero-initialised counter that only grows; `phase` counts this CTA's passages.  Thread 0 releases the CTA's
// writes (bar.sync + fence are cumulative) and acquires everyone else's before the CTA goes on.
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned& phase) {
  __syncthreads();
  if (threadIdx.x == 0) {
    phase++;
    __threadfence();
    atomicAdd(bar, 1u);
    const unsigned target = phase * gridDim.x;
    while (*(volatile unsigned*)bar < target) { }
    __threadfence();
  }
  __syncthreads();
}

}  // namespace gtp
