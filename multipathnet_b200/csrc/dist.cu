// dist.cu — the path's ONE collective: an all-gather of the fixed-size per-image detection records
// every rank packed on the device (post.cu), issued by the library itself on the ctx stream.
// The reference runs one replica per GPU inside one process (test_runner.lua:55-66) and joins the per-image
// results in the main thread (test_runner.lua:96-103,121-122); with one PROCESS per GPU that join is an NCCL
// all-gather over NVLink/NVSwitch. NCCL is bound at run time (dlopen of libnccl.so.2: the copy the host process
// already holds — e.g. the one PyTorch ships — else the system one), so libmpn_b200.so has no link-time NCCL
// version pin and the other entry points work on a box without NCCL.
#include "common.cuh"
#include <dlfcn.h>
#include <mutex>

namespace {

// the slice of nccl.h this file needs (NCCL 2.x ABI: ncclUniqueId = 128 opaque bytes, ncclFloat32 = 7)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
constexpr int kNcclFloat32 = 7;
static_assert(sizeof(ncclUniqueId) == MPN_DIST_ID_BYTES, "MPN_DIST_ID_BYTES must equal sizeof(ncclUniqueId)");

struct NcclApi {
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int *) = nullptr;
  std::string err;
};
NcclApi g_nccl;
std::once_flag g_nccl_once;

void nccl_load() {
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  void *h = nullptr;
  for (const char *n : names) { h = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL); if (h) break; }   // already in the process
  if (!h) for (const char *n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
  if (!h) { g_nccl.err = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "?"); return; }
  g_nccl.h = h;
  *(void **)&g_nccl.GetUniqueId = dlsym(h, "ncclGetUniqueId");
  *(void **)&g_nccl.CommInitRank = dlsym(h, "ncclCommInitRank");
  *(void **)&g_nccl.AllGather = dlsym(h, "ncclAllGather");
  *(void **)&g_nccl.CommDestroy = dlsym(h, "ncclCommDestroy");
  *(void **)&g_nccl.GetErrorString = dlsym(h, "ncclGetErrorString");
  *(void **)&g_nccl.GetVersion = dlsym(h, "ncclGetVersion");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.CommDestroy) {
    g_nccl.err = "libnccl.so.2 lacks ncclGetUniqueId / ncclCommInitRank / ncclAllGather / ncclCommDestroy"; g_nccl.h = nullptr;
  }
}
int nccl_ready(mpn_ctx *ctx) {
  std::call_once(g_nccl_once, nccl_load);
  if (!g_nccl.h) return mpn_fail(ctx, MPN_ERR_STATE, "NCCL unavailable: " + g_nccl.err);
  return MPN_OK;
}
#define MPN_NCCL(ctx, expr)                                                                        \
  do {                                                                                             \
    ncclResult_t r__ = (expr);                                                                     \
    if (r__ != 0) return mpn_fail((ctx), MPN_ERR_CUDA, std::string("NCCL error: ") +               \
                                  (g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "?"));     \
  } while (0)

}  // namespace

extern "C" {

int mpn_dist_unique_id(mpn_ctx *ctx, uint8_t *id) {
  if (!id) return MPN_ERR_ARG;
  MPN_TRY(nccl_ready(ctx));
  ncclUniqueId u;
  MPN_NCCL(ctx, g_nccl.GetUniqueId(&u));
  memcpy(id, &u, sizeof u);
  return MPN_OK;
}

int mpn_dist_init(mpn_ctx *ctx, const uint8_t *id, int32_t rank, int32_t world) {
  if (!ctx || !id) return MPN_ERR_ARG;
  MPN_CHECK_ARG(ctx, world >= 1 && rank >= 0 && rank < world, "mpn_dist_init: 0 <= rank < world");
  MPN_CHECK_ARG(ctx, !ctx->dist_comm, "mpn_dist_init: this ctx already has a communicator");
  MPN_TRY(nccl_ready(ctx));
  MPN_CUDA(ctx, cudaSetDevice(ctx->device));
  ncclUniqueId u; memcpy(&u, id, sizeof u);
  ncclComm_t comm = nullptr;
  MPN_NCCL(ctx, g_nccl.CommInitRank(&comm, world, u, rank));
  ctx->dist_comm = comm; ctx->dist_rank = rank; ctx->dist_world = world;
  return MPN_OK;
}

int mpn_dist_world(const mpn_ctx *ctx, int32_t *rank, int32_t *world) {
  if (!ctx) return MPN_ERR_ARG;
  if (rank) *rank = ctx->dist_comm ? ctx->dist_rank : 0;
  if (world) *world = ctx->dist_comm ? ctx->dist_world : 1;
  return MPN_OK;
}

int mpn_dist_all_gather_dev(mpn_ctx *ctx, const float *send_dev, int64_t n_floats, float *recv_dev) {
  if (!ctx) return MPN_ERR_ARG;
  MPN_CHECK_ARG(ctx, send_dev && recv_dev && n_floats > 0, "mpn_dist_all_gather_dev: buffers missing");
  MPN_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!ctx->dist_comm) {       // world of one: the gather is a copy (same stream ordering as the collective)
    if (recv_dev != send_dev)
      MPN_CUDA(ctx, cudaMemcpyAsync(recv_dev, send_dev, sizeof(float) * (size_t)n_floats, cudaMemcpyDeviceToDevice, ctx->stream));
    return MPN_OK;
  }
  MPN_NCCL(ctx, g_nccl.AllGather(send_dev, recv_dev, (size_t)n_floats, kNcclFloat32, (ncclComm_t)ctx->dist_comm, ctx->stream));
  ctx->collectives++;
  return MPN_OK;
}

int mpn_dist_all_gather(mpn_ctx *ctx, const float *send_dev, int64_t n_floats, float *recv_host) {
  if (!ctx) return MPN_ERR_ARG;
  MPN_CHECK_ARG(ctx, send_dev && recv_host && n_floats > 0, "mpn_dist_all_gather: buffers missing");
  const int world = ctx->dist_comm ? ctx->dist_world : 1;
  void *tmp = nullptr;
  MPN_TRY(mpn_scratch(ctx, sizeof(float) * (size_t)n_floats * world, &tmp));
  MPN_TRY(mpn_dist_all_gather_dev(ctx, send_dev, n_floats, (float *)tmp));
  MPN_CUDA(ctx, cudaMemcpyAsync(recv_host, tmp, sizeof(float) * (size_t)n_floats * world, cudaMemcpyDeviceToHost, ctx->stream));
  MPN_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MPN_OK;
}

int mpn_dist_destroy(mpn_ctx *ctx) {
  if (!ctx) return MPN_ERR_ARG;
  if (!ctx->dist_comm) return MPN_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (g_nccl.h) g_nccl.CommDestroy((ncclComm_t)ctx->dist_comm);
  ctx->dist_comm = nullptr; ctx->dist_world = 1; ctx->dist_rank = 0;
  return MPN_OK;
}

int mpn_dist_nccl_version(mpn_ctx *ctx, int32_t *version) {
  if (!version) return MPN_ERR_ARG;
  MPN_TRY(nccl_ready(ctx));
  int v = 0;
  if (g_nccl.GetVersion) g_nccl.GetVersion(&v);
  *version = v;
  return MPN_OK;
}

}  // extern "C"
