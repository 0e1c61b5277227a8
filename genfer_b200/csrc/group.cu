// Multi-GPU group contexts (north_star "large products are partitioned across the 8 GPUs of one box by
// sharding the output along the leading variable axis, the smaller operand replicated via NCCL all-gather over NVLink").
//
// One process per GPU, SPMD: every rank makes the same sequence of C-ABI calls on replicated handles.  A group context
// owns one NCCL communicator on its stream.  gtp_mul partitions a general product whose result has at least `threshold`
// coefficients: rank r computes the folded-cyclic leading-axis rows k0 mod 2W in {r, 2W-1-r} (row k0 costs k0 + 1
// sub-products, multivariate_taylor.rs:1001-1010, so the fold gives every rank the same MAC count) and the result handle
// stays ROW-SHARDED.  Whoever needs the whole tensor next (the next product of a Horner chain, gtp_to_host, any gather)
// replicates it once -- one grouped NCCL broadcast per row, straight into place -- and the replica is cached in the
// handle.  Operands can be uploaded in BLOCK shards (gtp_from_host_block: 1/W of the H2D traffic per rank) and are
// replicated by one ncclAllGather over NVLink on first use.
//
// NCCL is dlopen'ed here (libnccl.so.2: torch's bundled copy when the process has already loaded it, the system's
// otherwise); libgenfer_taylor.so has no link-time dependency on it.
#include <dlfcn.h>
#include <nccl.h>

#include "kernels.cuh"

namespace gtp {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static std::shared_ptr<NcclApi> load_nccl() {
  static std::shared_ptr<NcclApi> cached;
  if (cached) return cached;
  auto api = std::make_shared<NcclApi>();
  // GTP_NCCL_LIB names the library to use.  A process that also hosts PyTorch must load the SAME libnccl.so.2 torch
  // links against (its bundled copy): the loader keys shared objects by soname, so whichever copy comes first serves both
  // (genfer_b200/_lib.py points GTP_NCCL_LIB at torch's copy).  Other hosts get the system's NCCL.
  const char* env = getenv("GTP_NCCL_LIB");
  for (const char* name : {env ? env : "libnccl.so.2", "libnccl.so.2", "libnccl.so"}) {
    api->lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
    if (api->lib) break;
  }
  GTP_CHECK(api->lib, GTP_ERR_CUDA, std::string("cannot load NCCL (libnccl.so.2): ") + (dlerror() ? dlerror() : "not found"));
  auto sym = [&](const char* n) {
    void* p = dlsym(api->lib, n);
    GTP_CHECK(p, GTP_ERR_CUDA, std::string("NCCL symbol missing: ") + n);
    return p;
  };
  api->GetUniqueId = (decltype(api->GetUniqueId))sym("ncclGetUniqueId");
  api->CommInitRank = (decltype(api->CommInitRank))sym("ncclCommInitRank");
  api->CommDestroy = (decltype(api->CommDestroy))sym("ncclCommDestroy");
  api->AllGather = (decltype(api->AllGather))sym("ncclAllGather");
  api->Broadcast = (decltype(api->Broadcast))sym("ncclBroadcast");
  api->GroupStart = (decltype(api->GroupStart))sym("ncclGroupStart");
  api->GroupEnd = (decltype(api->GroupEnd))sym("ncclGroupEnd");
  api->GetErrorString = (decltype(api->GetErrorString))sym("ncclGetErrorString");
  cached = api;
  return api;
}

#define GTP_NCCL(api, expr)                                                                                  \
  do {                                                                                                       \
    ncclResult_t _r = (expr);                                                                                \
    if (_r != ncclSuccess) throw ::gtp::Error(GTP_ERR_CUDA, std::string(#expr) + ": " + (api)->GetErrorString(_r)); \
  } while (0)

Group::~Group() {
  if (comm && api && api->CommDestroy) api->CommDestroy((ncclComm_t)comm);
}

// folded-cyclic row map (the same as genfer_b200/partition.py::rows_for_rank)
void partition_rows(u64 n_rows, int world, int rank, std::vector<u64>* out) {
  out->clear();
  const u64 period = 2 * (u64)world;
  for (u64 k = 0; k < n_rows; k++) {
    const u64 m = k % period;
    if (m == (u64)rank || m == period - 1 - (u64)rank) out->push_back(k);
  }
}
static int row_owner(u64 k, int world) {
  const u64 period = 2 * (u64)world, m = k % period;
  return (int)(m < (u64)world ? m : period - 1 - m);
}

const double* replicate(const ShardState& cs) {
  ShardState& s = const_cast<ShardState&>(cs);   // the cache of an immutable value
  if (s.full) return s.full->d;
  Ctx& c = *s.ctx;
  Group& g = *s.group;
  NcclApi* api = g.api.get();
  ncclComm_t comm = (ncclComm_t)g.comm;
  if (s.kind == ShardState::BLOCK) {
    // every rank holds `block` slices (zero padded): one all-gather; the tensor is the prefix of the padded result
    BufP full = c.alloc((u64)g.world * s.block * s.row_elems);
    const double* src = s.local ? s.local->d : s.local_ptr;
    GTP_NCCL(api, api->AllGather(src, full->d, s.block * s.row_elems, ncclFloat64, comm, c.stream));
    s.full = full;
  } else {
    BufP full = c.alloc(std::max<u64>(s.n_rows * s.row_elems, 1));
    // one broadcast per leading-axis row, root = its owner, straight into place (no permutation pass)
    std::vector<u64> next(g.world, 0);   // next local slot of every rank, in ascending row order
    GTP_NCCL(api, api->GroupStart());
    for (u64 k = 0; k < s.n_rows; k++) {
      const int root = row_owner(k, g.world);
      const u64 slot = next[root]++;
      const double* send = root == g.rank ? s.local->d + slot * s.row_elems : nullptr;
      double* recv = full->d + k * s.row_elems;
      GTP_NCCL(api, api->Broadcast(send ? (const void*)send : (const void*)recv, recv, s.row_elems, ncclFloat64, root, comm, c.stream));
    }
    GTP_NCCL(api, api->GroupEnd());
    s.full = full;
  }
  g.gathers++;
  return s.full->d;
}

}  // namespace gtp

using namespace gtp;

namespace {
template <class F> int gwrap(gtp_ctx* ctx, F&& f) {
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
Shape to_shape(const uint64_t* p, int n) { return p ? Shape(p, p + n) : Shape(); }
}  // namespace

extern "C" {

int gtp_nccl_unique_id(void* out128) {
  if (!out128) return GTP_ERR_ARG;
  try {
    auto api = load_nccl();
    ncclUniqueId id;
    if (api->GetUniqueId(&id) != ncclSuccess) return GTP_ERR_CUDA;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(out128, &id, 128);
    return GTP_OK;
  } catch (const std::exception& e) {
    fprintf(stderr, "gtp_nccl_unique_id: %s\n", e.what());
    return GTP_ERR_CUDA;
  }
}

int gtp_ctx_create_group(int device, void* cuda_stream, int rank, int world, const void* id128, gtp_ctx** out) {
  if (!out || !id128 || world < 1 || rank < 0 || rank >= world) return GTP_ERR_ARG;
  int rc = gtp_ctx_create(device, cuda_stream, out);
  if (rc != GTP_OK) return rc;
  gtp_ctx* c = *out;
  rc = gwrap(c, [&] {
    auto g = std::make_shared<Group>();
    g->rank = rank;
    g->world = world;
    g->api = load_nccl();
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t comm;
    GTP_NCCL(g->api, g->api->CommInitRank(&comm, world, id, rank));
    g->comm = comm;
    if (const char* t = getenv("GTP_PARTITION_THRESHOLD")) g->threshold = strtoull(t, nullptr, 10);
    c->group = g;
  });
  if (rc != GTP_OK) {
    fprintf(stderr, "gtp_ctx_create_group: %s\n", c->err.c_str());
    gtp_ctx_destroy(c);
    *out = nullptr;
  }
  return rc;
}

int gtp_ctx_group_info(gtp_ctx* c, int* rank, int* world, uint64_t* partitioned_products, uint64_t* gathers) {
  if (!c) return GTP_ERR_ARG;
  if (rank) *rank = c->group ? c->group->rank : 0;
  if (world) *world = c->group ? c->group->world : 1;
  if (partitioned_products) *partitioned_products = c->group ? c->group->partitioned_products : 0;
  if (gathers) *gathers = c->group ? c->group->gathers : 0;
  return GTP_OK;
}

int gtp_ctx_set_partition_threshold(gtp_ctx* c, uint64_t coefficients) {
  if (!c || !c->group) return GTP_ERR_ARG;
  c->group->threshold = coefficients;
  return GTP_OK;
}

// integer work, no device needed: the leading-axis rows of `rank` (ascending); returns their number
uint64_t gtp_partition_rows(uint64_t n_rows, int world, int rank, uint64_t* rows_out) {
  if (world < 1 || rank < 0 || rank >= world) return 0;
  std::vector<u64> rows;
  partition_rows(n_rows, world, rank, &rows);
  if (rows_out) std::copy(rows.begin(), rows.end(), rows_out);
  return rows.size();
}
// the contiguous block [lo, hi) of leading-axis slices rank holds of a block-sharded operand, and the padded block length
void gtp_partition_block(uint64_t n_slices, int world, int rank, uint64_t* lo, uint64_t* hi, uint64_t* block) {
  const u64 b = (n_slices + (u64)world - 1) / (u64)world;
  const u64 l = std::min<u64>(n_slices, (u64)rank * b), h = std::min<u64>(n_slices, l + b);
  if (lo) *lo = l;
  if (hi) *hi = h;
  if (block) *block = b;
}

static gtp_poly* make_block_poly(gtp_ctx* c, int ndim, const uint64_t* shape, const uint64_t* degrees, BufP local, const double* local_ptr) {
  Shape sh = to_shape(shape, ndim), dg = to_shape(degrees, ndim);
  GTP_CHECK(sh.size() == dg.size() && ndim >= 1 && ndim <= GTP_MAX_NDIM, GTP_ERR_SHAPE, "bad shape");
  for (int i = 0; i < ndim; i++) GTP_CHECK(0 < sh[i] && sh[i] <= dg[i], GTP_ERR_SHAPE, "need 0 < shape[i] <= degrees_p1[i]");
  auto p = std::make_unique<gtp_poly>();
  p->shape = sh;
  p->degrees = dg;
  auto s = std::make_shared<ShardState>();
  s->kind = ShardState::BLOCK;
  s->ctx = c;
  s->group = c->group;
  s->local = local;
  s->local_ptr = local_ptr;
  s->n_rows = sh[0];
  s->row_elems = 1;
  for (int i = 1; i < ndim; i++) s->row_elems *= sh[i];
  s->block = (sh[0] + c->group->world - 1) / c->group->world;
  p->shard = s;
  return p.release();
}

int gtp_from_host_block(gtp_ctx* c, int ndim, const uint64_t* shape, const uint64_t* degrees, const double* block_data, gtp_poly** out) {
  return gwrap(c, [&] {
    GTP_CHECK(out && shape && degrees && c->group, GTP_ERR_ARG, "gtp_from_host_block needs a group context");
    u64 lo, hi, block;
    gtp_partition_block(shape[0], c->group->world, c->group->rank, &lo, &hi, &block);
    u64 row_elems = 1;
    for (int i = 1; i < ndim; i++) row_elems *= shape[i];
    BufP local = c->alloc(std::max<u64>(block * row_elems, 1));
    if (hi - lo < block) GTP_CUDA(cudaMemsetAsync(local->d + (hi - lo) * row_elems, 0, (block - (hi - lo)) * row_elems * sizeof(double), c->stream));
    if (hi > lo) {
      GTP_CHECK(block_data, GTP_ERR_ARG, "null block");
      GTP_CUDA(cudaMemcpyAsync(local->d, block_data, (hi - lo) * row_elems * sizeof(double), cudaMemcpyHostToDevice, c->stream));
      cudaPointerAttributes attr;
      if (cudaPointerGetAttributes(&attr, block_data) != cudaSuccess) { cudaGetLastError(); c->sync(); }
      else if (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged) c->sync();   // see gtp_from_host
    }
    *out = make_block_poly(c, ndim, shape, degrees, local, nullptr);
  });
}

int gtp_from_device_block(gtp_ctx* c, int ndim, const uint64_t* shape, const uint64_t* degrees, const double* device_block, gtp_poly** out) {
  return gwrap(c, [&] {
    GTP_CHECK(out && shape && degrees && device_block && c->group, GTP_ERR_ARG, "gtp_from_device_block needs a group context");
    *out = make_block_poly(c, ndim, shape, degrees, nullptr, device_block);
  });
}

int gtp_is_distributed(const gtp_poly* p) { return p && p->shard && !p->shard->full ? 1 : 0; }

int gtp_replicate(gtp_ctx* c, const gtp_poly* p) {
  return gwrap(c, [&] {
    GTP_CHECK(p, GTP_ERR_ARG, "null argument");
    if (p->shard) replicate(*p->shard);
  });
}

// the rows this rank holds of a row-sharded product result (all rows of a replicated polynomial)
uint64_t gtp_local_rows(const gtp_poly* p, uint64_t* rows_out) {
  if (!p) return 0;
  if (p->shard && p->shard->kind == ShardState::ROWS) {
    if (rows_out) std::copy(p->shard->rows.begin(), p->shard->rows.end(), rows_out);
    return p->shard->rows.size();
  }
  const u64 n = p->shape.empty() ? 1 : p->shape[0];
  if (rows_out) for (u64 i = 0; i < n; i++) rows_out[i] = i;
  return n;
}

// D2H of this rank's rows only (row gtp_local_rows()[i] at out + i * prod(shape[1:])); synchronises.  No collective.
int gtp_to_host_local(gtp_ctx* c, const gtp_poly* p, double* out) {
  return gwrap(c, [&] {
    GTP_CHECK(p && out, GTP_ERR_ARG, "null argument");
    if (p->shard && p->shard->kind == ShardState::ROWS) {
      const ShardState& s = *p->shard;
      GTP_CUDA(cudaMemcpyAsync(out, s.local->d, s.rows.size() * s.row_elems * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    } else {
      GTP_CUDA(cudaMemcpyAsync(out, p->ptr(), p->len() * sizeof(double), cudaMemcpyDefault, c->stream));
    }
    c->sync();
  });
}

}  // extern "C"
