// Fused Horner loop of subst_var (multivariate_taylor.rs:569-579) for a substitution of at most 32 coefficients:
//
//     for i in (0..len_v).rev():   res = res * subst + self.slice(v, i)
//
// The reference spends three whole-tensor operations per step (the slice copy, a general Mul, an Add); the per-operator
// device path mirrored that: 3 launches and 7 tensor passes per step, plus a classification of every intermediate.  Here
// the WHOLE remaining loop is ONE cooperative kernel: a step reads the previous result once (ld.global.cg: other CTAs
// wrote it), reads the slice straight out of `self` (no copy) and writes the next result -- 2 passes -- with a grid
// barrier between steps; results ping-pong between two buffers that stay L2-resident for small and mid-size tensors.
//
// Arithmetic per output coefficient is exactly the per-operator path's, i.e. the reference's:
//   * the product in the reference's nesting order (`mul` :984-1012 with Y = subst): outer axes ascending in the X index,
//     the innermost non-unit axis summed from zero and then added, multiply and add separate (== k_mul_stencil /
//     k_mul_ordered);
//   * Add's general path `(0 + a?) + b?` on zero-extended operands (:873-880), or its scalar path (`first += s`, :862-865)
//     when the slice is a single coefficient.
// The host enters the fused loop only once `res` has more than one coefficient and every axis that will ever be non-unit is
// non-unit in `res` (after at most two per-operator steps): from then on Mul's data-dependent fast paths (zero, one,
// constant) cannot fire, the set of non-unit axes -- which fixes the reference's summation nesting -- no longer changes, and
// the mul_linear fast path gives the same bits as the general product (two-term sums).
#include "device_sync.cuh"
#include "kernels.cuh"

namespace gtp {

constexpr int HN_MAXE = 8;
constexpr int HN_MAXT = 32;
constexpr int HN_T = 256;

struct HornerP {
  int ne, nt;
  unsigned d[HN_MAXE];          // degrees_p1 clip (UINT_MAX: unbounded)
  unsigned sshape[HN_MAXE];     // subst extents
  unsigned slice[HN_MAXE];      // slice extents (1 along v)
  long long selfstr[HN_MAXE];   // strides of `self` along the effective axes
  long long self_vstr;          // stride of axis v in `self`
  unsigned shape0[HN_MAXE];     // shape of the incoming res
  unsigned char m[HN_MAXT][HN_MAXE];
  unsigned char group_start[HN_MAXT];
  unsigned short sidx[HN_MAXT];
  const double* self;
  const double* subst;
  const double* res0;
  double* buf[2];
  unsigned i_top, nsteps;
  int slice_scalar;
  unsigned* bar;
};

__global__ void __launch_bounds__(HN_T) k_horner(const __grid_constant__ HornerP p) {
  __shared__ double s_sv[HN_MAXT];
  __shared__ long long s_delta[HN_MAXT];   // offset of res[k - m_t] relative to res[k], for this step's strides
  if (threadIdx.x < (unsigned)p.nt) s_sv[threadIdx.x] = p.subst[p.sidx[threadIdx.x]];
  __syncthreads();
  const int ne = p.ne;
  unsigned cur[HN_MAXE], nxt[HN_MAXE], prodsh[HN_MAXE];
#pragma unroll
  for (int a = 0; a < HN_MAXE; a++) cur[a] = a < ne ? p.shape0[a] : 1u;
  unsigned phase = 0;
  const double* src = p.res0;
  for (unsigned step = 0; step < p.nsteps; step++) {
    double* dst = p.buf[step & 1u];
    const unsigned i = p.i_top - step;
    // shapes of this step: product (sum_shape :150-170), result of the Add (max_shape :129-148)
    unsigned total = 1;
    long long cstr[HN_MAXE];
#pragma unroll
    for (int a = 0; a < HN_MAXE; a++) {
      if (a < ne) {
        const unsigned long long s = (unsigned long long)cur[a] + p.sshape[a] - 1ull;
        prodsh[a] = (unsigned)min(s, (unsigned long long)p.d[a]);
        nxt[a] = max(prodsh[a], p.slice[a]);
        total *= nxt[a];
      } else {
        prodsh[a] = nxt[a] = 1u;
      }
    }
    {
      long long st = 1;
#pragma unroll
      for (int a = HN_MAXE - 1; a >= 0; --a) {
        if (a < ne) { cstr[a] = st; st *= (long long)cur[a]; } else cstr[a] = 0;
      }
    }
    __syncthreads();   // every thread is done with the previous step's deltas
    if (threadIdx.x < (unsigned)p.nt) {
      long long dl = 0;
#pragma unroll
      for (int a = 0; a < HN_MAXE; a++)
        if (a < ne) dl += (long long)p.m[threadIdx.x][a] * cstr[a];
      s_delta[threadIdx.x] = dl;
    }
    __syncthreads();
    const double* slice_base = p.self + (long long)i * p.self_vstr;
    for (unsigned lin = blockIdx.x * blockDim.x + threadIdx.x; lin < total; lin += gridDim.x * blockDim.x) {
      unsigned k[HN_MAXE];
      unsigned rem = lin;
      long long base = 0, so = 0;
      bool a_ok = true, b_ok = true;
#pragma unroll
      for (int a = HN_MAXE - 1; a >= 0; --a) {
        if (a < ne) {
          const unsigned q = rem / nxt[a];
          k[a] = rem - q * nxt[a];
          rem = q;
          base += (long long)k[a] * cstr[a];
          so += (long long)k[a] * p.selfstr[a];
          a_ok = a_ok && k[a] < prodsh[a];
          b_ok = b_ok && k[a] < p.slice[a];
        } else {
          k[a] = 0;
        }
      }
      double r = 0.0;
      double prod = 0.0;
      if (a_ok) {
        double total_v = 0.0, inner = 0.0;
        bool open = false;
        for (int t = 0; t < p.nt; t++) {
          if (p.group_start[t]) {
            if (open) total_v = __dadd_rn(total_v, inner);
            inner = 0.0;
            open = true;
#pragma unroll
            for (int a = 0; a < HN_MAXE - 1; a++)
              if (a < ne - 1) open = open && (k[a] - (unsigned)p.m[t][a]) < cur[a];   // unsigned: k >= m too
          }
          if (open) {
            const unsigned ml = p.m[t][ne - 1];
            if ((k[ne - 1] - ml) < cur[ne - 1]) inner = __dadd_rn(inner, __dmul_rn(ldcg(src + (base - s_delta[t])), s_sv[t]));
          }
        }
        if (open) total_v = __dadd_rn(total_v, inner);
        prod = total_v;
      }
      if (p.slice_scalar) {
        // Add's scalar path (:862-865): the product with `first += slice[0]`
        r = (lin == 0) ? __dadd_rn(prod, slice_base[0]) : prod;
      } else {
        if (a_ok) r = __dadd_rn(r, prod);
        if (b_ok) r = __dadd_rn(r, slice_base[so]);
      }
      dst[lin] = r;
    }
#pragma unroll
    for (int a = 0; a < HN_MAXE; a++) cur[a] = nxt[a];
    src = dst;
    if (step + 1 < p.nsteps) grid_barrier(p.bar, phase);
  }
}

// Host side.  `self`: the polynomial whose axis v is substituted (shape self_shape, already zero-extended to the common
// ndim), `d`: result degrees, `subst`: shape sshape (<= 32 coefficients, <= d), `res`: the running Horner value (shape
// rshape) BEFORE the step that adds slice i_top.  Runs the steps i_top, i_top-1, .., 0 and returns the final buffer and
// shape; false when the case is outside the kernel's domain (the caller continues with per-operator steps).
bool launch_horner(Ctx& ctx, const double* self, const Shape& self_shape, u64 v, const Shape& d, const double* subst,
                   const Shape& sshape, const double* res, const Shape& rshape, u64 i_top, BufP* out_buf, Shape* out_shape) {
  const int nd = (int)d.size();
  if ((int)self_shape.size() != nd || (int)sshape.size() != nd || (int)rshape.size() != nd) return false;
  const u64 ns = prod(sshape);
  if (ns < 2 || ns > (u64)HN_MAXT || prod(rshape) < 2) return false;
  Shape slice(nd);
  for (int a = 0; a < nd; a++) slice[a] = std::min(self_shape[a], d[a]);
  slice[v] = 1;
  // effective axes: the non-unit axes of res; every axis that can become non-unit must already be one
  std::vector<int> eff;
  for (int a = 0; a < nd; a++) {
    if (rshape[a] > 1) eff.push_back(a);
    else if (sshape[a] > 1 || slice[a] > 1) return false;
    if (sshape[a] > d[a] || sshape[a] > 255) return false;
  }
  const int ne = (int)eff.size();
  if (ne < 1 || ne > HN_MAXE) return false;
  // final shape (the same recurrence the kernel runs) and overflow checks
  Shape cur = rshape;
  const u64 nsteps = i_top + 1;
  for (u64 s = 0; s < nsteps; s++) {
    bool changed = false;
    for (int a = 0; a < nd; a++) {
      u64 pr = std::min<u64>(d[a], cur[a] + sshape[a] - 1);
      u64 nx = std::max<u64>(pr, slice[a]);
      changed |= nx != cur[a];
      cur[a] = nx;
    }
    if (!changed) break;
  }
  const u64 final_total = prod(cur);
  if (final_total >= (1ull << 31) || prod(self_shape) >= (1ull << 40)) return false;
  HornerP p;
  memset(&p, 0, sizeof(p));
  p.ne = ne;
  Shape selfst(nd, 1), sst(nd, 1);
  for (int a = nd - 2; a >= 0; --a) {
    selfst[a] = selfst[a + 1] * self_shape[a + 1];
    sst[a] = sst[a + 1] * sshape[a + 1];
  }
  for (int e = 0; e < ne; e++) {
    const int a = eff[e];
    p.d[e] = d[a] >= 0xffffffffull ? 0xffffffffu : (unsigned)d[a];
    p.sshape[e] = (unsigned)sshape[a];
    p.slice[e] = (unsigned)slice[a];
    p.selfstr[e] = (long long)selfst[a];
    p.shape0[e] = (unsigned)rshape[a];
  }
  p.self_vstr = (long long)selfst[v];
  // terms of subst in the reference's visiting order: X = res index ascending <=> m descending (lexicographic)
  struct Term { std::vector<unsigned> m; u64 idx; };
  std::vector<Term> terms;
  for (u64 lin = 0; lin < ns; lin++) {
    u64 rem = lin;
    std::vector<unsigned> full(nd);
    for (int a = nd - 1; a >= 0; --a) { full[a] = (unsigned)(rem % sshape[a]); rem /= sshape[a]; }
    Term t;
    t.idx = lin;
    for (int e = 0; e < ne; e++) t.m.push_back(full[eff[e]]);
    terms.push_back(t);
  }
  std::sort(terms.begin(), terms.end(), [](const Term& x, const Term& y) { return x.m > y.m; });
  p.nt = (int)terms.size();
  for (int t = 0; t < p.nt; t++) {
    for (int e = 0; e < ne; e++) p.m[t][e] = (unsigned char)terms[t].m[e];
    p.sidx[t] = (unsigned short)terms[t].idx;
    bool new_group = t == 0;
    for (int e = 0; e < ne - 1 && !new_group; e++) new_group = terms[t].m[e] != terms[t - 1].m[e];
    p.group_start[t] = new_group ? 1 : 0;
  }
  p.self = self;
  p.subst = subst;
  p.res0 = res;
  p.i_top = (unsigned)i_top;
  p.nsteps = (unsigned)nsteps;
  p.slice_scalar = prod(slice) == 1 ? 1 : 0;
  BufP b0 = ctx.alloc(final_total), b1 = nsteps > 1 ? ctx.alloc(final_total) : BufP();
  BufP bar = ctx.alloc(2);
  GTP_CUDA(cudaMemsetAsync(bar->d, 0, 16, ctx.stream));
  p.buf[0] = b0->d;
  p.buf[1] = b1 ? b1->d : b0->d;
  p.bar = reinterpret_cast<unsigned*>(bar->d);
  static int coop[64] = {};
  int& c = coop[ctx.device & 63];
  if (c == 0) {
    int vv = 0;
    cudaDeviceGetAttribute(&vv, cudaDevAttrCooperativeLaunch, ctx.device);
    c = vv ? 1 : -1;
  }
  if (c < 0) return false;
  static int per_sm_cached[64] = {};
  int& per_sm = per_sm_cached[ctx.device & 63];
  if (per_sm == 0) {
    GTP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_horner, HN_T, 0));
    if (per_sm < 1) per_sm = -1;
  }
  if (per_sm < 0) return false;
  const u64 want = (final_total + HN_T - 1) / HN_T;
  // small and mid-size tensors are barrier-bound (a step is ~1 us of L2 traffic): one CTA per SM keeps the barrier short;
  // HBM-sized tensors want every resident CTA for loads in flight
  const int ctas_per_sm = final_total >= (1u << 21) ? std::min(per_sm, 4) : (final_total >= (1u << 18) ? std::min(per_sm, 2) : 1);
  const unsigned grid = (unsigned)std::max<u64>(1, std::min<u64>(want, (u64)ctas_per_sm * ctx.sm_count));
  void* args[] = {(void*)&p};
  const double t0 = ctx.hist ? Ctx::now() : 0.0;
  GTP_CUDA(cudaLaunchCooperativeKernel((const void*)k_horner, dim3(grid), dim3(HN_T), args, 0, ctx.stream));
  ctx.launches++;
  if (ctx.hist) {
    (*ctx.hist)["k_horner"]++;
    ctx.t_launch += Ctx::now() - t0;
  }
  *out_buf = ((nsteps - 1) & 1) ? b1 : b0;
  *out_shape = cur;
  return true;
}

}  // namespace gtp
