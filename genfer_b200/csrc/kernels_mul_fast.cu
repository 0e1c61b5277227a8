// Register-tiled DFMA kernel for the truncated N-D product (multivariate_taylor.rs:984-1012) on dense
// operands whose last axis has the same length LT in X, Y and Z.
//
// Decomposition (DESIGN.md "product kernel"):
//   axes = [ A-axes ... | b1 | b2 | L ]
//   * L (last axis, length LT <= 16): one truncated 1-D convolution  z[0..LT) += x[0..LT) (*) y[0..LT)
//     is done entirely in REGISTERS by one thread: LT(LT+1)/2 DFMAs for 2*LT operand loads, fully
//     unrolled, no wasted (padded) multiply-adds.
//   * b1, b2 ("slab" axes): one X slab X[jA,:,:,:] and one Y slab Y[kA-jA,:,:,:] are staged in shared
//     memory (cp.async, rows padded so per-lane 16-byte row reads are bank-conflict free).  The slab-pair
//     product consists of  sum_{k1,k2} cnt(k1)*cnt(k2)  row convolutions; a host-built STEP TABLE assigns
//     them to the 64 threads of the CTA so that every lane has the same number of steps (for the triangular
//     cube case this is the k <-> D-1-k folding) and each lane's steps for one output row are contiguous.
//   * A-axes: a work unit = (output slab kA, chunk [q0,q1) of the jA iteration box).  Units are sorted by
//     size (longest first) so the hardware block scheduler performs LPT list scheduling; heavy slabs are
//     split into several units (split-K) whose partial rows meet in HBM through red.global.add.f64.
//   Output rows live in registers while a lane works on them and are flushed with RED when the lane moves
//   to another row (table order is reversed on every other slab pair so the row at the seam stays put).
#include <map>
#include <mutex>

#include "kernels.cuh"

namespace gtp {

constexpr int FT = 128;      // threads per CTA: 2 teams of 64 lanes that interleave the steps of one slab pair
constexpr int F_MAXA = 6;    // A-axes handled
constexpr unsigned E_VALID = 1u << 30;
constexpr unsigned E_NONE = 0xffffffffu;

struct FastP {
  int na;                                   // number of A axes
  unsigned xa[F_MAXA], ya[F_MAXA], ra[F_MAXA];   // lengths along the A axes
  long long xastr[F_MAXA], yastr[F_MAXA];   // element strides of the A axes in X / Y
  unsigned rows_a0;                         // number of computed rows along A axis 0 (row subset)
  unsigned x_rows, y_rows, z_rows;          // rows (of LT doubles) per X / Y / Z slab
  int nsteps;
  const unsigned* table;                    // [nsteps][FT]
  const uint4* units;                       // {packed kA index, q0, q1, k along A axis 0}
  const double* x;
  const double* y;
  double* out;
};

template <int LT> struct RowStride { static constexpr int value = ((LT / 2) % 2 == 1) ? LT : LT + 2; };

template <int LT>
__device__ __forceinline__ void rowconv(double (&z)[LT], const double* __restrict__ xs, const double* __restrict__ ys) {
  double x[LT], y[LT];
#pragma unroll
  for (int i = 0; i < LT; i += 2) {
    double2 a = *reinterpret_cast<const double2*>(xs + i);
    double2 b = *reinterpret_cast<const double2*>(ys + i);
    x[i] = a.x; x[i + 1] = a.y;
    y[i] = b.x; y[i + 1] = b.y;
  }
#pragma unroll
  for (int j = 0; j < LT; j++) {
#pragma unroll
    for (int k = j; k < LT; k++) z[k] = fma(x[j], y[k - j], z[k]);
  }
}

template <int LT>
__device__ __forceinline__ void flush_row(double (&z)[LT], double* __restrict__ dst) {
#pragma unroll
  for (int i = 0; i < LT; i++) {
    atomicAdd(dst + i, z[i]);  // result unused -> RED.E.ADD.F64
    z[i] = 0.0;
  }
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}

template <int LT>
__global__ void __launch_bounds__(FT, 3) k_mul_tiled(const FastP p) {
  constexpr int LTP = RowStride<LT>::value;
  constexpr int V2 = LT / 2;  // double2 per row
  extern __shared__ __align__(16) double smem[];
  double* Xs = smem;
  double* Ys = smem + (size_t)p.x_rows * LTP;
  const int tid = threadIdx.x;

  const uint4 unit = p.units[blockIdx.x];
  // ---- decode the output slab kA (packed index over the computed rows) ----
  unsigned k[F_MAXA], lo[F_MAXA], ext[F_MAXA];
  {
    unsigned rem = unit.x;
#pragma unroll
    for (int a = F_MAXA - 1; a >= 0; --a) {
      if (a < p.na) {
        unsigned len = (a == 0) ? p.rows_a0 : p.ra[a];
        unsigned idx = rem % len;
        rem /= len;
        k[a] = (a == 0) ? unit.w : idx;
        unsigned l = (k[a] + 1 > p.ya[a]) ? k[a] + 1 - p.ya[a] : 0;
        unsigned h = (k[a] + 1 < p.xa[a]) ? k[a] + 1 : p.xa[a];
        lo[a] = l;
        ext[a] = h > l ? h - l : 0;
      } else {
        k[a] = lo[a] = 0;
        ext[a] = 1;
      }
    }
  }
  double* out_slab = p.out + (size_t)unit.x * p.z_rows * LT;

  double z[LT];
#pragma unroll
  for (int i = 0; i < LT; i++) z[i] = 0.0;
  unsigned cur = E_NONE;

  for (unsigned q = unit.y; q < unit.z; ++q) {
    // ---- jA from the linear index q over the box prod(ext) (last A axis fastest) ----
    long long xo = 0, yo = 0;
    {
      unsigned rem = q;
#pragma unroll
      for (int a = F_MAXA - 1; a >= 0; --a) {
        if (a < p.na) {
          unsigned j = lo[a] + rem % ext[a];
          rem /= ext[a];
          xo += (long long)j * p.xastr[a];
          yo += (long long)(k[a] - j) * p.yastr[a];
        }
      }
    }
    __syncthreads();  // everyone is done with the previous slab pair
    {
      const double2* gx = reinterpret_cast<const double2*>(p.x + xo);
      const double2* gy = reinterpret_cast<const double2*>(p.y + yo);
      const int nx = (int)p.x_rows * V2, ny = (int)p.y_rows * V2;
      for (int i = tid; i < nx; i += FT) {
        int r = i / V2, c = i - r * V2;
        cp_async16(Xs + r * LTP + 2 * c, gx + i);
      }
      for (int i = tid; i < ny; i += FT) {
        int r = i / V2, c = i - r * V2;
        cp_async16(Ys + r * LTP + 2 * c, gy + i);
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
    }
    __syncthreads();

    // ---- the slab-pair product, table driven; direction alternates so the seam row stays in registers ----
    const bool fwd = ((q - unit.y) & 1u) == 0;
    int step = fwd ? 0 : p.nsteps - 1;
    const int dstep = fwd ? 1 : -1;
    unsigned e = p.table[step * FT + tid];
    for (int s = 0; s < p.nsteps; ++s) {
      step += dstep;
      unsigned en = (s + 1 < p.nsteps) ? p.table[step * FT + tid] : 0u;
      if (e & E_VALID) {
        unsigned zr = e & 1023u;
        if (zr != cur) {
          if (cur != E_NONE) flush_row<LT>(z, out_slab + (size_t)cur * LT);
          cur = zr;
        }
        rowconv<LT>(z, Xs + ((e >> 10) & 1023u) * LTP, Ys + ((e >> 20) & 1023u) * LTP);
      }
      e = en;
    }
  }
  if (cur != E_NONE) flush_row<LT>(z, out_slab + (size_t)cur * LT);
}

// ------------------------------------------------------------------------------------------
// host side: applicability, step table, work units (cached per shape signature)
// ------------------------------------------------------------------------------------------
struct FastPlan {
  BufP table, units;
  int nsteps = 0;
  unsigned n_units = 0;
  FastP p;
  size_t smem = 0;
  int lt = 0;
  bool blocked22 = false;
};

static bool lt_supported(u64 lt) { return lt == 8 || lt == 10 || lt == 12 || lt == 16; }

struct FastGeom {
  int nd, na;
  u64 lt;
  u64 xb1, xb2, yb1, yb2, rb1, rb2;
};
static bool fast_geom(const MulArgs& a, FastGeom* g) {
  int nd = a.ndim;
  if (nd < 4) return false;                       // needs at least one A axis besides b1, b2, L
  u64 lt = a.rs[nd - 1];
  if (a.xs[nd - 1] != lt || a.ys[nd - 1] != lt || !lt_supported(lt)) return false;
  if (nd - 3 > F_MAXA) return false;
  g->nd = nd;
  g->na = nd - 3;
  g->lt = lt;
  g->xb1 = a.xs[nd - 3]; g->xb2 = a.xs[nd - 2];
  g->yb1 = a.ys[nd - 3]; g->yb2 = a.ys[nd - 2];
  g->rb1 = a.rs[nd - 3]; g->rb2 = a.rs[nd - 2];
  if (g->xb1 * g->xb2 > 1023 || g->yb1 * g->yb2 > 1023 || g->rb1 * g->rb2 > 1023) return false;
  u64 ltp = ((lt / 2) % 2 == 1) ? lt : lt + 2;
  u64 smem = (g->xb1 * g->xb2 + g->yb1 * g->yb2) * ltp * 8;
  if (smem > 74 * 1024) return false;             // 3 CTAs per SM
  for (int d = 0; d < nd; d++)
    if (a.xs[d] == 0 || a.ys[d] == 0 || a.rs[d] == 0) return false;
  return true;
}

bool fast_mul_applicable(const Ctx&, const MulArgs& a) {
  FastGeom g;
  if (a.accumulate) return false;
  if (!fast_geom(a, &g)) return false;
  // enough parallel slabs to be worth it
  u64 slabs = a.row_count;
  for (int d = 1; d < g.na; d++) slabs *= a.rs[d];
  return slabs >= 64;
}

bool fast_mul_cube16(const Ctx& ctx, const MulArgs& a) {
  FastGeom g;
  if (!fast_mul_applicable(ctx, a) || !fast_geom(a, &g)) return false;
  return g.lt == 16 && g.xb1 == 16 && g.yb1 == 16 && g.rb1 == 16 && g.xb2 == 16 && g.yb2 == 16 && g.rb2 == 16;
}

// Balanced step table for one slab-pair product.  Work item = (z row, x row, y row).  Rows are dealt
// to the 64 lanes longest-first onto the least loaded lane (LPT); a lane's items for one row stay contiguous.
// Triangular 16x16 slab (the dense-cube case): the k <-> 15-k folding.  Lane (c1,c2), c in 0..7, owns
// the four rows {c1,15-c1} x {c2,15-c2} and walks the step grid (s1,s2) in 0..16 x 0..16:
//     s <= c : row c,    j = s            s > c : row 15-c, j = s-c-1
// so every lane has exactly 17*17 steps.  tid = c1*8 + c2 puts the eight c2 of one c1 in one quarter-warp:
// at any step they share j1 and their j2 (and k2-j2) are either equal or consecutive, i.e. distinct modulo 8,
// which with the 144-byte row stride makes every LDS.128 of a quarter-warp bank-conflict free.  The s2
// direction alternates with s1 so a lane switches output row only ~18 times per slab pair.
static bool build_fold_table(const FastGeom& g, std::vector<unsigned>* table, int* nsteps) {
  const u64 D = 16;
  if (g.xb1 != D || g.yb1 != D || g.rb1 != D || g.xb2 != D || g.yb2 != D || g.rb2 != D) return false;
  const int S = (int)D + 1;
  const int TEAMS = FT / 64;
  *nsteps = (S * S + TEAMS - 1) / TEAMS;
  table->assign((size_t)(*nsteps) * FT, 0u);
  for (int c1 = 0; c1 < 8; c1++)
    for (int c2 = 0; c2 < 8; c2++) {
      int step = 0;
      for (int s1 = 0; s1 < S; s1++) {
        int r1 = s1 <= c1 ? c1 : 15 - c1, j1 = s1 <= c1 ? s1 : s1 - c1 - 1;
        for (int t = 0; t < S; t++) {
          int s2 = (s1 & 1) ? S - 1 - t : t;
          int r2 = s2 <= c2 ? c2 : 15 - c2, j2 = s2 <= c2 ? s2 : s2 - c2 - 1;
          unsigned zr = (unsigned)(r1 * 16 + r2), xr = (unsigned)(j1 * 16 + j2);
          unsigned yr = (unsigned)((r1 - j1) * 16 + (r2 - j2));
          // consecutive steps of the walk go to alternating teams (both teams flush partial rows with RED)
          int team = step % TEAMS, slot = step / TEAMS;
          int lane = team * 64 + c1 * 8 + c2;
          (*table)[(size_t)slot * FT + lane] = E_VALID | zr | (xr << 10) | (yr << 20);
          step++;
        }
      }
    }
  return true;
}

static void build_table(const FastGeom& g, std::vector<unsigned>* table, int* nsteps) {
  if (build_fold_table(g, table, nsteps)) return;
  struct Row { unsigned zr; u64 work; };
  std::vector<Row> rows;
  auto cnt = [](u64 k, u64 xl, u64 yl) -> u64 {
    u64 lo = sat_sub(k + 1, yl), hi = std::min(k + 1, xl);
    return hi > lo ? hi - lo : 0;
  };
  for (u64 k1 = 0; k1 < g.rb1; k1++)
    for (u64 k2 = 0; k2 < g.rb2; k2++) {
      u64 w = cnt(k1, g.xb1, g.yb1) * cnt(k2, g.xb2, g.yb2);
      if (w) rows.push_back({(unsigned)(k1 * g.rb2 + k2), w});
    }
  std::stable_sort(rows.begin(), rows.end(), [](const Row& a, const Row& b) { return a.work > b.work; });
  std::vector<std::vector<unsigned>> lane_items(FT);
  std::vector<u64> load(FT, 0);
  for (const Row& r : rows) {
    int best = 0;
    for (int l = 1; l < FT; l++)
      if (load[l] < load[best]) best = l;
    u64 k1 = r.zr / g.rb2, k2 = r.zr % g.rb2;
    u64 lo1 = sat_sub(k1 + 1, g.yb1), hi1 = std::min(k1 + 1, g.xb1);
    u64 lo2 = sat_sub(k2 + 1, g.yb2), hi2 = std::min(k2 + 1, g.xb2);
    for (u64 j1 = lo1; j1 < hi1; j1++)
      for (u64 j2 = lo2; j2 < hi2; j2++) {
        unsigned xr = (unsigned)(j1 * g.xb2 + j2);
        unsigned yr = (unsigned)((k1 - j1) * g.yb2 + (k2 - j2));
        lane_items[best].push_back(E_VALID | r.zr | (xr << 10) | (yr << 20));
      }
    load[best] += r.work;
  }
  u64 mx = 0;
  for (int l = 0; l < FT; l++) mx = std::max<u64>(mx, lane_items[l].size());
  *nsteps = (int)mx;
  table->assign((size_t)mx * FT, 0u);
  for (int l = 0; l < FT; l++)
    for (size_t s = 0; s < lane_items[l].size(); s++) (*table)[s * FT + l] = lane_items[l][s];
}


// ------------------------------------------------------------------------------------------
// 2x2-blocked variant for the dense 16-cube slab (b1 = b2 = L = 16 in X, Y and Z).
//
// The 1x1 kernel above needs 2*16 operand loads per 136 DFMA: at full DFMA rate that is ~120 B/clk of
// the SM's 128 B/clk shared-memory bandwidth, so the LSU and the FP64 pipe saturate together (measured:
// both ~46 %).  Here one step multiplies TWO x rows (j1, a), (j1, a+1) with TWO y rows (m1, b), (m1, b+1):
// four row convolutions (544 DFMA) for 64 operand loads, accumulated into the THREE output rows
// (r1, s), (r1, s+1), (r1, s+2), s = a+b, that stay in registers while the lane walks down the
// anti-diagonal a+b = s.  Shared-memory traffic per DFMA halves.
//   lane  = (c1 in 0..7 : k <-> 15-k folding of axis b1) x (d in 0..3 : anti-diagonal pairs s/2 = d, 7-d)
//           -> 17*(d+1) + 17*(8-d) = 153 block steps for every lane, 32 lanes per team, 4 teams per CTA
//           taking every 4th step.  The block on the last anti-diagonal (s = 14) computes one row
//           convolution that falls outside the truncation (row 16) and is discarded: 18496 useful of
//           19584 executed row convolutions (94.4 %).
//   smem  : row (p, q) at p*PLANE + q*18 doubles, PLANE = 16*18 + 2: the eight lanes of a quarter-warp
//           (same d, c1 = 0..7) touch rows that differ only in p by consecutive values (or coincide), and
//           the plane skew maps those to distinct 16-byte bank groups: conflict-free LDS.128.
// ------------------------------------------------------------------------------------------
constexpr int B22_LT = 16;
constexpr int B22_ROW = 18;                   // doubles per padded row
constexpr int B22_PLANE = 16 * B22_ROW + 2;   // doubles per padded plane
constexpr int B22_SLAB = 16 * B22_PLANE;      // doubles per slab in smem

__device__ __forceinline__ int b22_off(unsigned row) { return (int)(row >> 4) * B22_PLANE + (int)(row & 15u) * B22_ROW; }

__device__ __forceinline__ void b22_flush(double (&z)[16], double* __restrict__ dst) {
#pragma unroll
  for (int i = 0; i < 16; i++) {
    atomicAdd(dst + i, z[i]);
    z[i] = 0.0;
  }
}

__global__ void __launch_bounds__(FT, 2) k_mul_tiled22(const FastP p) {
  extern __shared__ __align__(16) double smem[];
  double* Xs = smem;
  double* Ys = smem + B22_SLAB;
  const int tid = threadIdx.x;
  const uint4 unit = p.units[blockIdx.x];
  unsigned k[F_MAXA], lo[F_MAXA], ext[F_MAXA];
  {
    unsigned rem = unit.x;
#pragma unroll
    for (int a = F_MAXA - 1; a >= 0; --a) {
      if (a < p.na) {
        unsigned len = (a == 0) ? p.rows_a0 : p.ra[a];
        unsigned idx = rem % len;
        rem /= len;
        k[a] = (a == 0) ? unit.w : idx;
        unsigned l = (k[a] + 1 > p.ya[a]) ? k[a] + 1 - p.ya[a] : 0;
        unsigned h = (k[a] + 1 < p.xa[a]) ? k[a] + 1 : p.xa[a];
        lo[a] = l;
        ext[a] = h > l ? h - l : 0;
      } else {
        k[a] = lo[a] = 0;
        ext[a] = 1;
      }
    }
  }
  double* out_slab = p.out + (size_t)unit.x * 256 * 16;

  double z0[16], z1[16], z2[16];
#pragma unroll
  for (int i = 0; i < 16; i++) z0[i] = z1[i] = z2[i] = 0.0;
  unsigned cur = E_NONE;  // z row held in z0 (z1, z2 are the next two rows of the same plane)

  for (unsigned q = unit.y; q < unit.z; ++q) {
    long long xo = 0, yo = 0;
    {
      unsigned rem = q;
#pragma unroll
      for (int a = F_MAXA - 1; a >= 0; --a) {
        if (a < p.na) {
          unsigned j = lo[a] + rem % ext[a];
          rem /= ext[a];
          xo += (long long)j * p.xastr[a];
          yo += (long long)(k[a] - j) * p.yastr[a];
        }
      }
    }
    __syncthreads();
    {
      const double2* gx = reinterpret_cast<const double2*>(p.x + xo);
      const double2* gy = reinterpret_cast<const double2*>(p.y + yo);
      for (int i = tid; i < 256 * 8; i += FT) {  // 256 rows x 8 double2
        int r = i >> 3, c = i & 7;
        int off = b22_off((unsigned)r) + 2 * c;
        cp_async16(Xs + off, gx + i);
        cp_async16(Ys + off, gy + i);
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
    }
    __syncthreads();

    const bool fwd = ((q - unit.y) & 1u) == 0;
    int step = fwd ? 0 : p.nsteps - 1;
    const int dstep = fwd ? 1 : -1;
    unsigned e = p.table[step * FT + tid];
    for (int s = 0; s < p.nsteps; ++s) {
      step += dstep;
      unsigned en = (s + 1 < p.nsteps) ? p.table[step * FT + tid] : 0u;
      if (e & E_VALID) {
        const unsigned xr = e & 255u, yr = (e >> 8) & 255u, zr = (e >> 16) & 255u;
        if (zr != cur) {
          if (cur != E_NONE) {
            double* dst = out_slab + (size_t)cur * 16;
            b22_flush(z0, dst);
            b22_flush(z1, dst + 16);
            if ((cur & 15u) + 2u < 16u) b22_flush(z2, dst + 32);
            else {
#pragma unroll
              for (int i = 0; i < 16; i++) z2[i] = 0.0;
            }
          }
          cur = zr;
        }
        const double* xs = Xs + b22_off(xr);
        const double* ys = Ys + b22_off(yr);
        double y0[16], y1[16];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          double2 u = *reinterpret_cast<const double2*>(ys + i);
          double2 v = *reinterpret_cast<const double2*>(ys + B22_ROW + i);
          y0[i] = u.x; y0[i + 1] = u.y;
          y1[i] = v.x; y1[i + 1] = v.y;
        }
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          double2 xa = *reinterpret_cast<const double2*>(xs + j);            // x row a
          double2 xb = *reinterpret_cast<const double2*>(xs + B22_ROW + j);  // x row a+1
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const double xaj = h ? xa.y : xa.x, xbj = h ? xb.y : xb.x;
            const int jj = j + h;
#pragma unroll
            for (int kk = jj; kk < 16; kk++) {
              z0[kk] = fma(xaj, y0[kk - jj], z0[kk]);
              z1[kk] = fma(xaj, y1[kk - jj], z1[kk]);
              z1[kk] = fma(xbj, y0[kk - jj], z1[kk]);
              z2[kk] = fma(xbj, y1[kk - jj], z2[kk]);
            }
          }
        }
      }
      e = en;
    }
  }
  if (cur != E_NONE) {
    double* dst = out_slab + (size_t)cur * 16;
    b22_flush(z0, dst);
    b22_flush(z1, dst + 16);
    if ((cur & 15u) + 2u < 16u) b22_flush(z2, dst + 32);
  }
}

// step table of the 2x2-blocked kernel: entry = valid | xrow0 | yrow0 << 8 | zrow0 << 16
static void build_table22(std::vector<unsigned>* table, int* nsteps) {
  const int TEAMS = FT / 32;
  const int SEQ = 153;
  *nsteps = (SEQ + TEAMS - 1) / TEAMS;
  table->assign((size_t)(*nsteps) * FT, 0u);
  for (int d = 0; d < 4; d++)
    for (int c1 = 0; c1 < 8; c1++) {
      int i = 0;
      for (int phase = 0; phase < 2; phase++) {
        int half = phase == 0 ? d : 7 - d, s = 2 * half;
        for (int t = 0; t < 17; t++) {
          int s1 = phase == 0 ? t : 16 - t;  // keeps r1 fixed across the seam between the two diagonals
          int r1 = s1 <= c1 ? c1 : 15 - c1, j1 = s1 <= c1 ? s1 : s1 - c1 - 1;
          for (int ab = 0; ab <= half; ab++) {
            int a = 2 * ab, b = s - a;
            unsigned xr = (unsigned)(j1 * 16 + a), yr = (unsigned)((r1 - j1) * 16 + b), zr = (unsigned)(r1 * 16 + s);
            int team = i % TEAMS, slot = i / TEAMS;
            // warp = one d (all four teams of it): the lanes of a warp then switch output rows at (almost) the
            // same steps, so the divergent flush code runs ~3x less often; quarter-warp = one team, c1 = 0..7
            int lane = d * 32 + team * 8 + c1;
            (*table)[(size_t)slot * FT + lane] = E_VALID | xr | (yr << 8) | (zr << 16);
            i++;
          }
        }
      }
    }
}

static bool geom_is_cube16(const FastGeom& g) {
  return g.lt == 16 && g.xb1 == 16 && g.yb1 == 16 && g.rb1 == 16 && g.xb2 == 16 && g.yb2 == 16 && g.rb2 == 16;
}

struct PlanKey {
  std::vector<u64> v;
  bool operator<(const PlanKey& o) const { return v < o.v; }
};
using PlanCache = std::map<PlanKey, std::shared_ptr<FastPlan>>;
static PlanCache& plan_cache(Ctx& ctx) {
  if (!ctx.fast_plans) ctx.fast_plans = std::make_shared<PlanCache>();
  return *std::static_pointer_cast<PlanCache>(ctx.fast_plans);
}

template <int LT> static void launch_tiled(Ctx& ctx, const FastPlan& pl, const FastP& p) {
  static bool configured[64] = {};
  if (!configured[ctx.device & 63]) {
    GTP_CUDA(cudaFuncSetAttribute(k_mul_tiled<LT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 74 * 1024));
    configured[ctx.device & 63] = true;
  }
  GTP_LAUNCH(ctx, k_mul_tiled<LT>, pl.n_units, FT, pl.smem, p);
}

void launch_mul_fast(Ctx& ctx, const MulArgs& a) {
  FastGeom g;
  GTP_CHECK(fast_geom(a, &g), GTP_ERR_ARG, "fast product kernel not applicable");
  PlanKey key;
  key.v.insert(key.v.end(), a.xs.begin(), a.xs.end());
  key.v.insert(key.v.end(), a.ys.begin(), a.ys.end());
  key.v.insert(key.v.end(), a.rs.begin(), a.rs.end());
  key.v.push_back(a.row_begin);
  key.v.push_back(a.row_step);
  key.v.push_back(a.row_count);
  key.v.insert(key.v.end(), a.rows.begin(), a.rows.end());
  auto& cache = plan_cache(ctx);
  std::shared_ptr<FastPlan> pl;
  auto it = cache.find(key);
  if (it != cache.end()) {
    pl = it->second;
  } else {
    pl = std::make_shared<FastPlan>();
    FastP& p = pl->p;
    memset(&p, 0, sizeof(p));
    const int nd = g.nd, na = g.na;
    Shape xst(nd, 1), yst(nd, 1);
    for (int i = nd - 2; i >= 0; --i) {
      xst[i] = xst[i + 1] * a.xs[i + 1];
      yst[i] = yst[i + 1] * a.ys[i + 1];
    }
    p.na = na;
    for (int d = 0; d < na; d++) {
      p.xa[d] = (unsigned)a.xs[d];
      p.ya[d] = (unsigned)a.ys[d];
      p.ra[d] = (unsigned)a.rs[d];
      p.xastr[d] = (long long)xst[d];
      p.yastr[d] = (long long)yst[d];
    }
    p.rows_a0 = (unsigned)a.row_count;
    p.x_rows = (unsigned)(g.xb1 * g.xb2);
    p.y_rows = (unsigned)(g.yb1 * g.yb2);
    p.z_rows = (unsigned)(g.rb1 * g.rb2);
    std::vector<unsigned> table;
    pl->blocked22 = geom_is_cube16(g);
    if (pl->blocked22) build_table22(&table, &pl->nsteps);
    else build_table(g, &table, &pl->nsteps);
    p.nsteps = pl->nsteps;
    // ---- work units: (packed kA, q0, q1), split-K chunks, longest first ----
    u64 n_slabs = a.row_count;
    for (int d = 1; d < na; d++) n_slabs *= a.rs[d];
    struct U { unsigned ka, q0, q1, k0; };
    auto row_of = [&](u64 idx) -> u64 { return a.rows.empty() ? a.row_begin + idx * a.row_step : a.rows[idx]; };
    std::vector<U> units;
    std::vector<u64> boxes(n_slabs);
    std::vector<unsigned> k0s(n_slabs, 0);
    u64 total_pairs = 0;
    for (u64 s = 0; s < n_slabs; s++) {
      u64 rem = s, box = 1;
      for (int d = na - 1; d >= 0; --d) {
        u64 len = (d == 0) ? a.row_count : a.rs[d];
        u64 idx = rem % len;
        rem /= len;
        u64 k = (d == 0) ? row_of(idx) : idx;
        if (d == 0) k0s[s] = (unsigned)k;
        u64 lo = sat_sub(k + 1, a.ys[d]), hi = std::min(k + 1, a.xs[d]);
        box *= hi > lo ? hi - lo : 0;
      }
      boxes[s] = box;
      total_pairs += box;
    }
    // chunk size: aim for ~16 units per CTA slot so list scheduling balances to a few percent
    u64 slots = (u64)ctx.sm_count * (pl->blocked22 ? 2 : 3);
    u64 chunk = std::max<u64>(8, total_pairs / (slots * 16) + 1);
    for (u64 s = 0; s < n_slabs; s++) {
      u64 box = boxes[s];
      if (!box) continue;
      u64 parts = (box + chunk - 1) / chunk;
      u64 per = (box + parts - 1) / parts;
      for (u64 q0 = 0; q0 < box; q0 += per) units.push_back({(unsigned)s, (unsigned)q0, (unsigned)std::min(box, q0 + per), k0s[s]});
    }
    std::stable_sort(units.begin(), units.end(), [](const U& x, const U& y) { return (x.q1 - x.q0) > (y.q1 - y.q0); });
    pl->n_units = (unsigned)units.size();
    std::vector<uint4> hu(units.size());
    for (size_t i = 0; i < units.size(); i++) hu[i] = make_uint4(units[i].ka, units[i].q0, units[i].q1, units[i].k0);
    pl->table = ctx.alloc((table.size() * sizeof(unsigned) + 7) / 8 + 1);
    pl->units = ctx.alloc((hu.size() * sizeof(uint4) + 7) / 8 + 1);
    GTP_CUDA(cudaMemcpyAsync(pl->table->d, table.data(), table.size() * sizeof(unsigned), cudaMemcpyHostToDevice, ctx.stream));
    if (!hu.empty())
      GTP_CUDA(cudaMemcpyAsync(pl->units->d, hu.data(), hu.size() * sizeof(uint4), cudaMemcpyHostToDevice, ctx.stream));
    ctx.sync();  // the staging vectors are pageable and die here
    p.table = reinterpret_cast<const unsigned*>(pl->table->d);
    p.units = reinterpret_cast<const uint4*>(pl->units->d);
    u64 ltp = ((g.lt / 2) % 2 == 1) ? g.lt : g.lt + 2;
    pl->smem = pl->blocked22 ? (size_t)(2 * B22_SLAB * 8) : (size_t)((p.x_rows + p.y_rows) * ltp * 8);
    pl->lt = (int)g.lt;
    if (cache.size() > 64) cache.clear();
    cache[key] = pl;
  }
  FastP p = pl->p;
  p.x = a.x;
  p.y = a.y;
  p.out = a.out;
  u64 row_elems = 1;
  for (int d = 1; d < a.ndim; d++) row_elems *= a.rs[d];
  GTP_CUDA(cudaMemsetAsync(a.out, 0, a.row_count * row_elems * sizeof(double), ctx.stream));
  if (pl->n_units == 0) return;
  if (pl->blocked22) {
    static bool configured22[64] = {};
    if (!configured22[ctx.device & 63]) {
      GTP_CUDA(cudaFuncSetAttribute(k_mul_tiled22, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * B22_SLAB * 8));
      configured22[ctx.device & 63] = true;
    }
    GTP_LAUNCH(ctx, k_mul_tiled22, pl->n_units, FT, pl->smem, p);
    return;
  }
  switch (pl->lt) {
    case 8: launch_tiled<8>(ctx, *pl, p); break;
    case 10: launch_tiled<10>(ctx, *pl, p); break;
    case 12: launch_tiled<12>(ctx, *pl, p); break;
    case 16: launch_tiled<16>(ctx, *pl, p); break;
    default: throw Error(GTP_ERR_ARG, "unsupported LT");
  }
}

}  // namespace gtp
