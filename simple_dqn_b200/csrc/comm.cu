// comm.cu — data-parallel learners: gradient all-reduce over NVLink 5 / NVSwitch through NCCL.
// libnccl.so.2 is dlopen()ed (the process normally already holds torch's bundled copy), so the
// library has no link-time NCCL dependency and single-GPU use never touches it.
#include <dlfcn.h>

#include "net.cuh"
#include "net_umma.cuh"

namespace b200 {

// Minimal NCCL ABI (nccl.h 2.x): opaque comm, 128-byte unique id, enums by value.
struct NcclUniqueId { char internal[128]; };
typedef void* ncclComm_t;
typedef int (*fn_GetUniqueId)(NcclUniqueId*);
typedef int (*fn_CommInitRank)(ncclComm_t*, int, NcclUniqueId, int);
typedef int (*fn_AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
typedef int (*fn_CommDestroy)(ncclComm_t);
typedef const char* (*fn_GetErrorString)(int);
constexpr int kNcclFloat32 = 7;  // ncclFloat32
constexpr int kNcclSum = 0;      // ncclSum

static struct {
  void* handle = nullptr;
  fn_GetUniqueId GetUniqueId = nullptr;
  fn_CommInitRank CommInitRank = nullptr;
  fn_AllReduce AllReduce = nullptr;
  fn_CommDestroy CommDestroy = nullptr;
  fn_GetErrorString GetErrorString = nullptr;
} g_nccl;

static int nccl_load() {
  if (g_nccl.handle) return B200DQN_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  B2_REQUIRE(h, B200DQN_ENCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
  g_nccl.GetUniqueId = (fn_GetUniqueId)dlsym(h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (fn_CommInitRank)dlsym(h, "ncclCommInitRank");
  g_nccl.AllReduce = (fn_AllReduce)dlsym(h, "ncclAllReduce");
  g_nccl.CommDestroy = (fn_CommDestroy)dlsym(h, "ncclCommDestroy");
  g_nccl.GetErrorString = (fn_GetErrorString)dlsym(h, "ncclGetErrorString");
  B2_REQUIRE(g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.AllReduce && g_nccl.CommDestroy, B200DQN_ENCCL,
             "libnccl.so.2 lacks an expected symbol");
  g_nccl.handle = h;
  return B200DQN_OK;
}

#define B2_CHECK_NCCL(expr)                                                                      \
  do {                                                                                           \
    int r__ = (expr);                                                                            \
    if (r__ != 0) {                                                                              \
      set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr,                                     \
                g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "nccl error");              \
      return B200DQN_ENCCL;                                                                      \
    }                                                                                            \
  } while (0)

int comm_allreduce_grads(b200dqn_net* n, cudaStream_t st) {
  B2_REQUIRE(n->nccl_comm, B200DQN_ESTATE, "communicator not initialised");
  B2_CHECK_NCCL(g_nccl.AllReduce(n->d_g, n->d_g, size_t(n->n_params), kNcclFloat32, kNcclSum,
                                 (ncclComm_t)n->nccl_comm, st));
  return B200DQN_OK;
}

// all-reduce the summed gradients of layers [l0, l1] (contiguous in d_g) on stream st
int comm_allreduce_range(b200dqn_net* n, int l0, int l1, cudaStream_t st) {
  B2_REQUIRE(n->nccl_comm, B200DQN_ESTATE, "communicator not initialised");
  float* p = n->d_g + n->lt.off[l0];
  const size_t cnt = size_t(n->lt.off[l1 + 1] - n->lt.off[l0]);
  B2_CHECK_NCCL(g_nccl.AllReduce(p, p, cnt, kNcclFloat32, kNcclSum, (ncclComm_t)n->nccl_comm, st));
  return B200DQN_OK;
}

void comm_destroy(b200dqn_net* n) {
  if (n->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)n->nccl_comm);
  n->nccl_comm = nullptr;
  n->world = 1;
  n->rank = 0;
}

}  // namespace b200

using namespace b200;

extern "C" int b200dqn_comm_unique_id(void* out_id128) {
  B2_REQUIRE(out_id128, B200DQN_EINVAL, "null id buffer");
  int rc = nccl_load();
  if (rc) return rc;
  NcclUniqueId id;
  B2_CHECK_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(out_id128, &id, sizeof(id));
  return B200DQN_OK;
}

extern "C" int b200dqn_net_comm_init(b200dqn_net* n, const void* id128, int rank, int world_size) {
  B2_REQUIRE(n && id128 && world_size >= 1 && rank >= 0 && rank < world_size, B200DQN_EINVAL,
             "net_comm_init: bad argument");
  B2_REQUIRE(!n->nccl_comm, B200DQN_ESTATE, "net_comm_init: communicator already initialised");
  int rc = nccl_load();
  if (rc) return rc;
  DeviceGuard g(n->device);
  NcclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm = nullptr;
  B2_CHECK_NCCL(g_nccl.CommInitRank(&comm, world_size, id, rank));
  n->nccl_comm = comm;
  n->rank = rank;
  n->world = world_size;
  // Warm-up collective outside any capture: NCCL sets up its channels/proxies lazily on the first call,
  // which must not happen inside the CUDA-graph capture of the train step.
  cudaStream_t ws;
  B2_CHECK_CUDA(cudaStreamCreateWithFlags(&ws, cudaStreamNonBlocking));
  B2_CHECK_CUDA(cudaMemsetAsync(n->d_g, 0, n->n_params * sizeof(float), ws));
  B2_CHECK_NCCL(g_nccl.AllReduce(n->d_g, n->d_g, size_t(n->n_params), kNcclFloat32, kNcclSum, comm, ws));
  B2_CHECK_CUDA(cudaStreamSynchronize(ws));
  B2_CHECK_CUDA(cudaStreamDestroy(ws));
  return B200DQN_OK;
}

extern "C" int b200dqn_net_comm_destroy(b200dqn_net* n) {
  B2_REQUIRE(n, B200DQN_EINVAL, "null net");
  DeviceGuard g(n->device);
  cudaDeviceSynchronize();
  // captured steps hold nodes of this communicator
  if (n->graph_exec) { cudaGraphExecDestroy(n->graph_exec); n->graph_exec = nullptr; }
  if (n->graph_train_exec) { cudaGraphExecDestroy(n->graph_train_exec); n->graph_train_exec = nullptr; }
  comm_destroy(n);
  return B200DQN_OK;
}
