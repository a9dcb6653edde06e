// Data-parallel gradient exchange: one NCCL communicator over the GPUs of the box, sum all-reduce
// of flat fp32 gradient buckets on a caller-chosen (side) stream so it overlaps backward.
// Replaces tf.contrib.distribute MirroredStrategy's cross_device_ops (NCCL all_sum on 2 packs)
//   <- /root/reference/utils/distribution_utils.py:85-98, /root/reference/core/estimator.py:570-613.
// NCCL is bound with dlopen so the library has no link-time dependency on it: single-GPU users
// and the CPU-side ABI test never need libnccl.
#include <dlfcn.h>
#include <cstdlib>
#include <cstring>
#include "internal.h"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat32 = 7, ncclSum = 0 };

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  const char* (*GetErrorString)(ncclResult_t);
};

NcclApi g_api;

int load_nccl(bsl_ctx* ctx) {
  if (ctx->nccl_lib) return BSL_OK;
  const char* names[] = {getenv("BSL_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* nm : names) {
    if (!nm || !*nm) continue;
    h = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
    if (h) break;
  }
  if (!h) return bsl_fail(ctx, BSL_ENCCL, "cannot dlopen libnccl.so.2 (%s); set BSL_NCCL_LIB", dlerror());
#define BIND(field, sym)                                                              \
  *reinterpret_cast<void**>(&g_api.field) = dlsym(h, sym);                            \
  if (!g_api.field) return bsl_fail(ctx, BSL_ENCCL, "libnccl is missing symbol %s", sym)
  BIND(GetUniqueId, "ncclGetUniqueId");
  BIND(CommInitRank, "ncclCommInitRank");
  BIND(AllReduce, "ncclAllReduce");
  BIND(CommDestroy, "ncclCommDestroy");
  BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
  ctx->nccl_lib = h;
  return BSL_OK;
}

int nccl_fail(bsl_ctx* ctx, ncclResult_t r, const char* what) {
  return bsl_fail(ctx, BSL_ENCCL, "NCCL error %d (%s) at %s", r, g_api.GetErrorString ? g_api.GetErrorString(r) : "?",
                  what);
}

}  // namespace

extern "C" {

int bsl_comm_unique_id(bsl_ctx* ctx, void* id128) {
  if (!ctx || !id128) return BSL_EINVAL;
  int rc = load_nccl(ctx);
  if (rc) return rc;
  ncclUniqueId id;
  ncclResult_t r = g_api.GetUniqueId(&id);
  if (r) return nccl_fail(ctx, r, "ncclGetUniqueId");
  memcpy(id128, &id, sizeof(id));
  return BSL_OK;
}

int bsl_comm_init(bsl_ctx* ctx, const void* id128, int rank, int world) {
  if (!ctx || !id128 || rank < 0 || world < 1 || rank >= world) return BSL_EINVAL;
  int rc = load_nccl(ctx);
  if (rc) return rc;
  BSL_CUDA(ctx, cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm;
  ncclResult_t r = g_api.CommInitRank(&comm, world, id, rank);
  if (r) return nccl_fail(ctx, r, "ncclCommInitRank");
  ctx->nccl_comm = comm;
  ctx->rank = rank;
  ctx->world = world;
  return BSL_OK;
}

int bsl_allreduce_sum_f32(bsl_ctx* ctx, float* buf, size_t n, void* stream) {
  if (!ctx || !buf) return BSL_EINVAL;
  if (ctx->world == 1) return BSL_OK;  // a 1-rank sum is the identity; nothing to enqueue
  if (!ctx->nccl_comm) return bsl_fail(ctx, BSL_ENCCL, "allreduce: communicator not initialised");
  ncclResult_t r = g_api.AllReduce(buf, buf, n, ncclFloat32, ncclSum, reinterpret_cast<ncclComm_t>(ctx->nccl_comm),
                                   as_stream(stream));
  if (r) return nccl_fail(ctx, r, "ncclAllReduce");
  return BSL_OK;
}

int bsl_comm_destroy(bsl_ctx* ctx) {
  if (!ctx) return BSL_EINVAL;
  if (ctx->nccl_comm) {
    g_api.CommDestroy(reinterpret_cast<ncclComm_t>(ctx->nccl_comm));
    ctx->nccl_comm = nullptr;
  }
  ctx->world = 1;
  ctx->rank = 0;
  return BSL_OK;
}

}  // extern "C"
