// comm.cu -- the one exchange step of the multi-GPU paths: NCCL over NVLink 5 / NVSwitch.
//
// The reference has no distributed code at all; what is sharded here are its independent units (edge checks, radius queries,
// worlds of the QMDP SSSP).  One process per GPU, one porrt_ctx per process.  A collective is issued only where a later
// device-resident stage needs every rank's slice (PRM: the valid (neighbour, new node) pairs before the CSR is assembled on
// every rank; QMDP: the dist rows of the other ranks' worlds).  NCCL is bound at run time with dlopen("libnccl.so.2") -- the
// library has no link-time NCCL dependency, and in a process that already holds an NCCL (torch) the same copy is reused.
// The 128-byte ncclUniqueId is created by porrt_comm_unique_id on rank 0 and carried to the other ranks by the host
// (torch.distributed / MPI / a file: any 128-byte broadcast).
#include <dlfcn.h>

#include <mutex>

#include "common.cuh"

namespace {
typedef struct { char internal[128]; } NcclId;          // ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES 128)
typedef void* NcclComm;
enum { kNcclInt8 = 0 };                                   // ncclDataType_t: ncclInt8 / ncclChar
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
  std::string why;
};

void nccl_load(NcclApi* out);

NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;   // two ranks' host threads of one process may get here together
  std::call_once(once, [] { nccl_load(&api); });
  return &api;
}

void nccl_load(NcclApi* out) {
  NcclApi& api = *out;
  const char* names[] = {getenv("PORRT_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    if (!nm || !*nm) continue;
    api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
    api.why = dlerror();
  }
  if (!api.handle) return;
  bool ok = true;
  auto sym = [&](const char* s) { void* p = dlsym(api.handle, s); if (!p) { ok = false; api.why = std::string("missing symbol ") + s; } return p; };
  api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
  api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
  api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
  api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
  api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
  api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
  if (!ok) { dlclose(api.handle); api.handle = nullptr; }
}

int32_t nccl_fail(porrt_ctx* ctx, const char* what, int rc) {
  NcclApi* a = nccl_api();
  return porrt_fail(ctx, PORRT_ERR_COMM, std::string(what) + ": " + (a->GetErrorString ? a->GetErrorString(rc) : "nccl error"));
}
#define NCCL_TRY(ctx, expr) do { int _r = (expr); if (_r != 0) return nccl_fail(ctx, #expr, _r); } while (0)
}  // namespace

// contiguous shard [lo, hi) of n units for `rank` of `world`: sizes differ by at most one, lower ranks take the extra unit
// (same rule as po_rrt_b200/shard.py:shard_range)
void comm_shard_range(int64_t n, int rank, int world, int64_t* lo, int64_t* hi) {
  const int64_t base = n / world, extra = n % world;
  *lo = rank * base + std::min<int64_t>(rank, extra);
  *hi = *lo + base + (rank < extra ? 1 : 0);
}

// every rank contributes `offsets[rank+1]-offsets[rank]` bytes (already in place at recv + offsets[rank] when send == nullptr)
int32_t comm_all_gatherv_dev(porrt_ctx* ctx, const void* send_dev, void* recv_dev, const int64_t* offsets /* host [world+1] */,
                             cudaStream_t st) {
  if (ctx->comm_world <= 1) {
    if (send_dev && send_dev != recv_dev && offsets[1] > offsets[0])
      CUDA_TRY(ctx, cudaMemcpyAsync((char*)recv_dev + offsets[0], send_dev, (size_t)(offsets[1] - offsets[0]), cudaMemcpyDeviceToDevice, st));
    return PORRT_OK;
  }
  NcclApi* a = nccl_api();
  if (!a->handle || !ctx->comm) return porrt_fail(ctx, PORRT_ERR_COMM, "communicator not initialised");
  const int W = ctx->comm_world, me = ctx->comm_rank;
  bool uniform = true;
  const int64_t sz0 = offsets[1] - offsets[0];
  for (int r = 0; r < W; ++r) uniform = uniform && (offsets[r + 1] - offsets[r] == sz0);
  const char* mine = send_dev ? (const char*)send_dev : (const char*)recv_dev + offsets[me];
  if (uniform) {
    if (sz0 > 0) NCCL_TRY(ctx, a->AllGather(mine, (char*)recv_dev + offsets[0], (size_t)sz0, kNcclInt8, (NcclComm)ctx->comm, st));
    return PORRT_OK;
  }
  // ragged shards: every rank's slice is padded to the longest one, ONE in-place ncclAllGather over a staging buffer, then the slices
  // are copied to their places (two extra passes over local HBM).  The grouped ncclBroadcast per root this replaces needed 2.8 ms for
  // the 105 MB of neighbour lists of a 1e6-node roadmap at 8 ranks -- an all-gather of that size is a few hundred microseconds.
  int64_t maxsz = 0;
  for (int r = 0; r < W; ++r) maxsz = std::max<int64_t>(maxsz, offsets[r + 1] - offsets[r]);
  if (maxsz <= 0) return PORRT_OK;
  maxsz = (maxsz + 15) & ~(int64_t)15;
  CUDA_TRY(ctx, ctx->comm_tmp.ensure((size_t)W * (size_t)maxsz));
  char* tmp = ctx->comm_tmp.as<char>();
  const int64_t sz_me = offsets[me + 1] - offsets[me];
  if (sz_me > 0) CUDA_TRY(ctx, cudaMemcpyAsync(tmp + (size_t)me * maxsz, mine, (size_t)sz_me, cudaMemcpyDeviceToDevice, st));
  NCCL_TRY(ctx, a->AllGather(tmp + (size_t)me * maxsz, tmp, (size_t)maxsz, kNcclInt8, (NcclComm)ctx->comm, st));
  for (int r = 0; r < W; ++r) {
    const int64_t sz = offsets[r + 1] - offsets[r];
    if (sz <= 0 || (r == me && !send_dev)) continue;   // (in place: this rank's slice is where it belongs already)
    CUDA_TRY(ctx, cudaMemcpyAsync((char*)recv_dev + offsets[r], tmp + (size_t)r * maxsz, (size_t)sz, cudaMemcpyDeviceToDevice, st));
  }
  return PORRT_OK;
}

PORRT_API int32_t porrt_comm_unique_id(uint8_t* out_id128) {
  if (!out_id128) return PORRT_ERR_INVALID_ARG;
  NcclApi* a = nccl_api();
  if (!a->handle) return PORRT_ERR_COMM;
  NcclId id;
  if (a->GetUniqueId(&id) != 0) return PORRT_ERR_COMM;
  memcpy(out_id128, id.internal, 128);
  return PORRT_OK;
}

PORRT_API int32_t porrt_comm_init(porrt_ctx* ctx, const uint8_t* id128, int32_t rank, int32_t world) {
  CTX_CHECK(ctx);
  if (world < 1 || rank < 0 || rank >= world) return porrt_fail(ctx, PORRT_ERR_INVALID_ARG, "comm_init: bad rank / world");
  if (ctx->comm) return porrt_fail(ctx, PORRT_ERR_INVALID_ARG, "comm_init: communicator already initialised");
  if (world == 1) { ctx->comm_rank = 0; ctx->comm_world = 1; return PORRT_OK; }
  if (!id128) return porrt_fail(ctx, PORRT_ERR_INVALID_ARG, "comm_init: null unique id");
  NcclApi* a = nccl_api();
  if (!a->handle) return porrt_fail(ctx, PORRT_ERR_COMM, "libnccl.so.2 not loadable: " + a->why);
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  NcclId id;
  memcpy(id.internal, id128, 128);
  NcclComm c = nullptr;
  NCCL_TRY(ctx, a->CommInitRank(&c, world, id, rank));
  ctx->comm = c; ctx->comm_rank = rank; ctx->comm_world = world;
  return PORRT_OK;
}

PORRT_API int32_t porrt_comm_destroy(porrt_ctx* ctx) {
  CTX_CHECK(ctx);
  if (ctx->comm) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    nccl_api()->CommDestroy((NcclComm)ctx->comm);
    ctx->comm = nullptr;
  }
  ctx->comm_rank = 0; ctx->comm_world = 1;
  return PORRT_OK;
}

PORRT_API int32_t porrt_comm_info(porrt_ctx* ctx, int32_t* out_rank, int32_t* out_world, int32_t* out_nccl_version) {
  CTX_CHECK(ctx);
  if (out_rank) *out_rank = ctx->comm_rank;
  if (out_world) *out_world = ctx->comm_world;
  if (out_nccl_version) {
    *out_nccl_version = 0;
    NcclApi* a = nccl_api();
    if (a->handle && a->GetVersion) { int v = 0; if (a->GetVersion(&v) == 0) *out_nccl_version = v; }
  }
  return PORRT_OK;
}

PORRT_API int32_t porrt_shard_range(int64_t n, int32_t rank, int32_t world, int64_t* out_lo, int64_t* out_hi) {
  if (n < 0 || world < 1 || rank < 0 || rank >= world || !out_lo || !out_hi) return PORRT_ERR_INVALID_ARG;
  comm_shard_range(n, rank, world, out_lo, out_hi);
  return PORRT_OK;
}

// All-gather of per-rank result slices that live on the device (validity ids, world masks, neighbour lists ...):
// rank r owns rows shard_range(n_total, r, world) of `bytes_per_unit` bytes each; recv_dev receives all n_total rows in rank order.
// send_dev may be null when the rank's slice already sits at its place inside recv_dev (in-place gather).
PORRT_API int32_t porrt_comm_all_gather_dev(porrt_ctx* ctx, const void* send_dev, void* recv_dev, int64_t n_total, int64_t bytes_per_unit) {
  CTX_CHECK(ctx);
  if (!recv_dev || n_total < 0 || bytes_per_unit <= 0) return porrt_fail(ctx, PORRT_ERR_INVALID_ARG, "comm_all_gather: bad arguments");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  std::vector<int64_t> off(ctx->comm_world + 1, 0);
  for (int r = 0; r < ctx->comm_world; ++r) {
    int64_t lo, hi;
    comm_shard_range(n_total, r, ctx->comm_world, &lo, &hi);
    off[r] = lo * bytes_per_unit; off[r + 1] = hi * bytes_per_unit;
  }
  return comm_all_gatherv_dev(ctx, send_dev, recv_dev, off.data(), ctx->stream);
}

// Ragged variant: rank r contributes counts[r] bytes (host array, identical on every rank).
PORRT_API int32_t porrt_comm_all_gatherv_dev(porrt_ctx* ctx, const void* send_dev, void* recv_dev, const int64_t* byte_counts) {
  CTX_CHECK(ctx);
  if (!recv_dev || !byte_counts) return porrt_fail(ctx, PORRT_ERR_INVALID_ARG, "comm_all_gatherv: bad arguments");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  std::vector<int64_t> off(ctx->comm_world + 1, 0);
  for (int r = 0; r < ctx->comm_world; ++r) {
    if (byte_counts[r] < 0) return porrt_fail(ctx, PORRT_ERR_INVALID_ARG, "comm_all_gatherv: negative count");
    off[r + 1] = off[r] + byte_counts[r];
  }
  return comm_all_gatherv_dev(ctx, send_dev, recv_dev, off.data(), ctx->stream);
}
