// Multi-GPU exchange steps of the ALS fit loop: one process per GPU, NCCL over NVLink 5 / NVSwitch.
// The reference has no multi-GPU path at all (`// TODO: multi-gpu support`, implicit/gpu/als.cu:169);
// the sharding is described in DESIGN.md: every rank holds full replicas of both factor matrices and
// a contiguous, nnz-balanced row shard of Cui and of Ciu; after each half the freshly solved rows are
// all-gathered into every replica.
//
// NCCL is bound lazily with dlopen so that the single-GPU product has no link-time dependency on it
// (and picks up whichever libnccl.so.2 the process already has mapped).
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include "common.h"

namespace als {
namespace {

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.handle) return ALS_OK;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  void *h = nullptr;
  for (const char *n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    set_error("NCCL is not loadable: %s", dlerror());
    return ALS_E_NCCL;
  }
#define ALS_SYM(field, name)                                           \
  *(void **)(&g_nccl.field) = dlsym(h, name);                          \
  if (!g_nccl.field) {                                                 \
    set_error("NCCL symbol %s not found", name);                       \
    return ALS_E_NCCL;                                                 \
  }
  ALS_SYM(GetUniqueId, "ncclGetUniqueId")
  ALS_SYM(CommInitRank, "ncclCommInitRank")
  ALS_SYM(CommDestroy, "ncclCommDestroy")
  ALS_SYM(AllReduce, "ncclAllReduce")
  ALS_SYM(Broadcast, "ncclBroadcast")
  ALS_SYM(AllGather, "ncclAllGather")
  ALS_SYM(GroupStart, "ncclGroupStart")
  ALS_SYM(GroupEnd, "ncclGroupEnd")
  ALS_SYM(GetErrorString, "ncclGetErrorString")
#undef ALS_SYM
  g_nccl.handle = h;
  return ALS_OK;
}

int nccl_fail(ncclResult_t r, const char *what) {
  set_error("NCCL error %d (%s) in %s", (int)r, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?", what);
  return ALS_E_NCCL;
}

#define ALS_NCCL(expr)                                   \
  do {                                                   \
    ncclResult_t r__ = (expr);                           \
    if (r__ != ncclSuccess) return nccl_fail(r__, #expr); \
  } while (0)

}  // namespace

int comm_allreduce_gramian(als_ctx *ctx, int n_floats) {
  if (ctx->world == 1 || !ctx->comm) return ALS_OK;
  // + 1: the failure flag written by gramian_reduce_kernel (see als_solver_status)
  ALS_NCCL(g_nccl.AllReduce(ctx->G, ctx->G, (size_t)n_floats + 1, ncclFloat, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
  return ALS_OK;
}

}  // namespace als

using namespace als;

static_assert(sizeof(ncclUniqueId) == ALS_COMM_ID_BYTES, "ncclUniqueId size");

ALS_API int als_comm_unique_id(void *id) {
  ALS_REQUIRE(id, "als_comm_unique_id: NULL");
  int rc = load_nccl();
  if (rc != ALS_OK) return rc;
  ncclUniqueId uid;
  ALS_NCCL(g_nccl.GetUniqueId(&uid));
  memcpy(id, &uid, sizeof(uid));
  return ALS_OK;
}

ALS_API int als_comm_init(als_ctx *ctx, int rank, int world, const void *id) {
  ALS_REQUIRE(ctx && id, "als_comm_init: NULL argument");
  ALS_REQUIRE(world >= 1 && rank >= 0 && rank < world, "als_comm_init: bad rank %d / world %d", rank, world);
  int rc = load_nccl();
  if (rc != ALS_OK) return rc;
  ALS_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  ncclComm_t comm;
  ALS_NCCL(g_nccl.CommInitRank(&comm, world, uid, rank));
  ctx->comm = comm;
  ctx->rank = rank;
  ctx->world = world;
  return ALS_OK;
}

ALS_API int als_comm_destroy(als_ctx *ctx) {
  if (!ctx || !ctx->comm) return ALS_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  g_nccl.CommDestroy((ncclComm_t)ctx->comm);
  ctx->comm = nullptr;
  ctx->rank = 0;
  ctx->world = 1;
  return ALS_OK;
}

ALS_API int als_comm_allgather_rows(als_ctx *ctx, als_factors *f, const int64_t *row_splits) {
  ALS_REQUIRE(ctx && f && row_splits, "als_comm_allgather_rows: NULL argument");
  if (ctx->world == 1) return ALS_OK;
  ALS_REQUIRE(ctx->comm, "als_comm_allgather_rows: communicator not initialised");
  ALS_REQUIRE(row_splits[0] == 0 && row_splits[ctx->world] == f->rows,
              "als_comm_allgather_rows: row_splits must run from 0 to %lld", (long long)f->rows);
  ALS_CUDA(cudaSetDevice(ctx->device));
  // variable shard sizes: one in-place broadcast per owner, fused into a single NCCL group
  ALS_NCCL(g_nccl.GroupStart());
  for (int r = 0; r < ctx->world; ++r) {
    const int64_t n = row_splits[r + 1] - row_splits[r];
    if (n <= 0) continue;
    float *p = f->d + row_splits[r] * (int64_t)f->ld;
    ALS_NCCL(g_nccl.Broadcast(p, p, (size_t)(n * f->ld), ncclFloat, r, (ncclComm_t)ctx->comm, ctx->stream));
  }
  ALS_NCCL(g_nccl.GroupEnd());
  return ALS_OK;
}

ALS_API int als_comm_allreduce_f64(als_ctx *ctx, double *values, int n, int op_max) {
  ALS_REQUIRE(ctx && values && n >= 0 && n <= 8, "als_comm_allreduce_f64: bad argument (n must be <= 8)");
  if (ctx->world == 1 || n == 0) return ALS_OK;
  ALS_REQUIRE(ctx->comm, "als_comm_allreduce_f64: communicator not initialised");
  ALS_CUDA(cudaSetDevice(ctx->device));
  ALS_CUDA(cudaMemcpyAsync(ctx->dscalars, values, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  ALS_NCCL(g_nccl.AllReduce(ctx->dscalars, ctx->dscalars, n, ncclDouble, op_max ? ncclMax : ncclSum,
                            (ncclComm_t)ctx->comm, ctx->stream));
  ALS_CUDA(cudaMemcpyAsync(values, ctx->dscalars, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  ALS_CUDA(cudaStreamSynchronize(ctx->stream));
  return ALS_OK;
}

ALS_API int als_comm_allgather_bytes(als_ctx *ctx, const void *send, void *recv, int nbytes) {
  ALS_REQUIRE(ctx && send && recv && nbytes > 0 && nbytes <= 256, "als_comm_allgather_bytes: bad argument");
  if (ctx->world == 1) {
    memcpy(recv, send, nbytes);
    return ALS_OK;
  }
  ALS_REQUIRE(ctx->comm, "als_comm_allgather_bytes: communicator not initialised");
  ALS_CUDA(cudaSetDevice(ctx->device));
  int rc = ensure_scratch(ctx, (int64_t)(ctx->world + 1) * 256);
  if (rc != ALS_OK) return rc;
  char *base = (char *)ctx->scratch;
  ALS_CUDA(cudaMemcpyAsync(base, send, nbytes, cudaMemcpyHostToDevice, ctx->stream));
  ALS_NCCL(g_nccl.AllGather(base, base + 256, nbytes, ncclChar, (ncclComm_t)ctx->comm, ctx->stream));
  ALS_CUDA(cudaMemcpyAsync(recv, base + 256, (size_t)nbytes * ctx->world, cudaMemcpyDeviceToHost, ctx->stream));
  ALS_CUDA(cudaStreamSynchronize(ctx->stream));
  return ALS_OK;
}

ALS_API int als_comm_barrier(als_ctx *ctx) {
  double one = 1.0;
  return als_comm_allreduce_f64(ctx, &one, 1, 0);
}
