// Internal types of libgenfer_taylor: context, immutable ref-counted device buffers, handles.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/genfer_taylor.h"

namespace gtp {

using u64 = uint64_t;
constexpr u64 UNB = GTP_UNBOUNDED;
using Shape = std::vector<u64>;

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
#define GTP_CHECK(cond, code, msg)                    \
  do {                                                \
    if (!(cond)) throw ::gtp::Error((code), (msg));   \
  } while (0)
#define GTP_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      throw ::gtp::Error(_e == cudaErrorMemoryAllocation ? GTP_ERR_OOM : GTP_ERR_CUDA,              \
                         std::string(#expr) + ": " + cudaGetErrorString(_e));                       \
  } while (0)

inline u64 prod(const Shape& s) {
  u64 p = 1;
  for (u64 x : s) p *= x;
  return p;
}
inline u64 sat_sub(u64 a, u64 b) { return a > b ? a - b : 0; }

struct Ctx;

// The stream outlives the context for as long as any buffer allocated on it is alive.
struct StreamCore {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own = false;
  ~StreamCore() {
    if (own && stream) cudaStreamDestroy(stream);
  }
};

// Immutable device buffer.  Freed stream-ordered (cudaFreeAsync) when the last handle drops it.
struct Buf {
  double* d = nullptr;
  u64 n = 0;       // doubles
  bool owned = true;
  std::shared_ptr<StreamCore> core;
  ~Buf() {
    if (owned && d && core) cudaFreeAsync(d, core->stream);
  }
};
using BufP = std::shared_ptr<Buf>;

// Pinned read-back page layout (host-visible, written by kernels / D2H copies)
struct Readback {
  unsigned int viol_mask;  // classify: bit v set <=> axis v is NOT a linear axis
  unsigned int flag;       // generic boolean result (eq / any)
  double vals[64];         // c, m, sums, gathered scalars ...
  volatile unsigned long long seq;  // written LAST by the single-CTA classify kernel (zero-copy path): host spins on it
};

struct Ctx {
  int device = 0;
  std::shared_ptr<StreamCore> core;
  cudaStream_t stream = nullptr;  // == core->stream
  int sm_count = 148;
  std::string err;
  Readback* rb_host = nullptr;  // pinned
  Readback* rb_dev = nullptr;   // device scratch mirrored into rb_host by cudaMemcpyAsync
  unsigned int* counter_dev = nullptr;  // "last block" tickets etc.
  double* gather_host = nullptr;        // pinned, for gtp_gather_axis / to_host staging
  u64 gather_cap = 0;
  u64 launches = 0;
  u64 rb_seq = 0;               // sequence number of the last zero-copy read-back request
  Readback* rb_host_dev = nullptr;  // device alias of rb_host (mapped pinned memory)
  int fast_mul = 1;  // 0: reference-order kernel only, 1: auto, 2: force the blocked kernel even on tiny products (tests)
  std::shared_ptr<void> blk_plans;   // per-context cache of product plans (kernels_mul_blk.cu)
  std::shared_ptr<void> slide_plans; // same for kernels_mul_slide.cu
  bool use_slide = true;             // sliding 1x2 kernel for dense cube slabs (false: 2x2-blocked kernel everywhere)
  int slide_tile = 0;                // 0: auto; 4 / 8: force the plane-tiled sliding plan with that many planes per slab
  bool fuse_mul_linear = true;       // one-pass mul_linear kernel (false: the reference's composition, for A/B tests)
  bool blk_octet = false;            // experimental octet tables for single-plane slabs (8 staged pairs = 8 lanes)
  bool blk_fold_tables = true;       // structured (folded) item tables for dense cube slabs; false: evenly dealt

  BufP alloc(u64 n_doubles);
  void sync() { GTP_CUDA(cudaStreamSynchronize(stream)); }
};

// Cached data-dependent classification of a handle (multivariate_taylor.rs:262-294, :643-655)
struct ClassInfo {
  bool known = false;
  bool linear = false;  // extract_linear() is Some
  double c = 0, m = 0;  // if linear
  u64 v = 0;            // if linear
  double first = 0;     // coeffs.first()
};

}  // namespace gtp

// The opaque handle types of the C ABI
struct gtp_poly {
  gtp::BufP buf;
  gtp::u64 off = 0;  // element offset of this polynomial inside buf (prefix views)
  gtp::Shape shape;    // coeffs.shape()
  gtp::Shape degrees;  // degrees_p1
  std::shared_ptr<gtp::ClassInfo> cls = std::make_shared<gtp::ClassInfo>();  // shared by O(1) clones of the same data
  const double* ptr() const { return buf->d + off; }
  gtp::u64 len() const { return gtp::prod(shape); }
  int ndim() const { return (int)shape.size(); }
};

struct gtu_series {
  bool is_const = false;
  gtp::BufP buf;  // 1 element for Constant, `n` for Polynomial
  gtp::u64 n = 1;
  const double* ptr() const { return buf->d; }
};

struct gtp_ctx : gtp::Ctx {};
