// Multi-GPU exchange of the C ABI: ONE NCCL all-gather of the per-rank global descriptors.
//
// Replaces: the single-device loop of eval/evaluate.py:454-466 when the clouds of a batch are sharded over the GPUs of a
// node (egonn_b200/parallel.py): every rank extracts its own clouds, then all ranks exchange the (clouds, 256) fp32 global
// descriptors; local descriptors / keypoints stay rank-local.  There is no collective inside the network.
//
// NCCL is resolved at RUN TIME (dlopen of the libnccl.so.2 the process already carries - PyTorch's bundled copy - or the
// path in EGN_NCCL_LIB), so libegonn_b200.so has no link-time dependency on it and single-GPU users never touch it.
// Only the five entry points below are used; their prototypes are restated here from the public nccl.h (NCCL 2.x ABI:
// ncclUniqueId is 128 opaque bytes passed by value, ncclFloat32 = 7).
#include <dlfcn.h>

#include <mutex>

#include "ctx.cuh"

namespace {

struct NcclUniqueId { char internal[EGN_COMM_ID_BYTES]; };
typedef void *NcclComm;
typedef int (*fn_get_unique_id)(NcclUniqueId *);
typedef int (*fn_comm_init_rank)(NcclComm *, int, NcclUniqueId, int);
typedef int (*fn_comm_destroy)(NcclComm);
typedef int (*fn_all_gather)(const void *, void *, size_t, int, NcclComm, cudaStream_t);
typedef const char *(*fn_error_string)(int);
constexpr int kNcclFloat32 = 7;

struct NcclApi {
  void *handle = nullptr;
  fn_get_unique_id get_unique_id = nullptr;
  fn_comm_init_rank comm_init_rank = nullptr;
  fn_comm_destroy comm_destroy = nullptr;
  fn_all_gather all_gather = nullptr;
  fn_error_string error_string = nullptr;
  char why[256] = {0};
};

NcclApi *nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char *env = getenv("EGN_NCCL_LIB");
    const char *names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      if (!n || !n[0]) continue;
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
      snprintf(api.why, sizeof(api.why), "%s", dlerror());
    }
    if (!api.handle) return;
    api.get_unique_id = (fn_get_unique_id)dlsym(api.handle, "ncclGetUniqueId");
    api.comm_init_rank = (fn_comm_init_rank)dlsym(api.handle, "ncclCommInitRank");
    api.comm_destroy = (fn_comm_destroy)dlsym(api.handle, "ncclCommDestroy");
    api.all_gather = (fn_all_gather)dlsym(api.handle, "ncclAllGather");
    api.error_string = (fn_error_string)dlsym(api.handle, "ncclGetErrorString");
    if (!(api.get_unique_id && api.comm_init_rank && api.comm_destroy && api.all_gather && api.error_string)) {
      snprintf(api.why, sizeof(api.why), "libnccl is loaded but lacks one of the five entry points used");
      api.handle = nullptr;
    }
  });
  return &api;
}

}  // namespace

struct egn_comm {
  NcclComm comm = nullptr;
  int device = 0, rank = 0, world = 1;
};

#define EGN_NCCL(api, expr)                                                                              \
  do {                                                                                                   \
    const int egn_r_ = (expr);                                                                           \
    EGN_CHECK(egn_r_ == 0, EGN_ERR_CUDA, "NCCL: %s (%s)", (api)->error_string(egn_r_), #expr);           \
  } while (0)

extern "C" {

int egn_comm_unique_id(void *id_out) {
  EGN_CHECK(id_out != nullptr, EGN_ERR_INVALID, "comm_unique_id: null buffer");
  NcclApi *api = nccl_api();
  EGN_CHECK(api->handle != nullptr, EGN_ERR_STATE, "NCCL is not available (set EGN_NCCL_LIB to libnccl.so.2): %s", api->why);
  NcclUniqueId id;
  EGN_NCCL(api, api->get_unique_id(&id));
  memcpy(id_out, &id, sizeof(id));
  return EGN_OK;
}

int egn_comm_create(egn_comm **out, int device, int rank, int world, const void *id) {
  EGN_CHECK(out && id, EGN_ERR_INVALID, "comm_create: null argument");
  EGN_CHECK(world >= 1 && rank >= 0 && rank < world, EGN_ERR_INVALID, "comm_create: rank %d of %d", rank, world);
  NcclApi *api = nccl_api();
  EGN_CHECK(api->handle != nullptr, EGN_ERR_STATE, "NCCL is not available (set EGN_NCCL_LIB to libnccl.so.2): %s", api->why);
  egn::DeviceGuard g(device);
  NcclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  egn_comm *c = new egn_comm();
  c->device = device; c->rank = rank; c->world = world;
  const int r = api->comm_init_rank(&c->comm, world, uid, rank);
  if (r != 0) {
    delete c;
    EGN_CHECK(false, EGN_ERR_CUDA, "NCCL: %s (ncclCommInitRank rank %d of %d)", api->error_string(r), rank, world);
  }
  *out = c;
  return EGN_OK;
}

int egn_comm_destroy(egn_comm *comm) {
  if (!comm) return EGN_OK;
  NcclApi *api = nccl_api();
  if (api->handle && comm->comm) {
    egn::DeviceGuard g(comm->device);
    api->comm_destroy(comm->comm);
  }
  delete comm;
  return EGN_OK;
}

int egn_allgather_global(egn_comm *comm, const float *send, float *recv, int64_t floats_per_rank, egn_stream_t stream) {
  EGN_CHECK(comm && send && recv && floats_per_rank > 0, EGN_ERR_INVALID, "allgather_global: bad argument");
  NcclApi *api = nccl_api();
  EGN_CHECK(api->handle != nullptr, EGN_ERR_STATE, "NCCL is not available: %s", api->why);
  egn::DeviceGuard g(comm->device);
  EGN_NCCL(api, api->all_gather(send, recv, (size_t)floats_per_rank, kNcclFloat32, comm->comm, (cudaStream_t)stream));
  return EGN_OK;
}

}  // extern "C"
