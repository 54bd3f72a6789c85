// Data-parallel plumbing of the training step (train.py:464-506 under a tf.distribute strategy):
// the model owns an NCCL communicator so that a host in any language can run the gradient all-reduce through the C ABI
// (uu_comm_unique_id -> exchange 128 bytes -> uu_comm_init on every rank).  NCCL is resolved at run time with dlopen:
// inside a PyTorch process this picks up the libnccl.so.2 torch already loaded (one NCCL per process), elsewhere the
// system library.  The all-reduce runs on a side stream in buckets that follow the order in which the backward pass
// finishes regions of the flat gradient buffer, so the exchange of the strided / temporal blocks overlaps the rest of
// the backward pass (uu_train_step in uu_train.cu).
#include <dlfcn.h>
#include <cstring>

#include "model.cuh"

namespace uu {

namespace {
struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
typedef int (*PFN_GetUniqueId)(NcclUniqueId*);
typedef int (*PFN_CommInitRank)(NcclComm*, int, NcclUniqueId, int);
typedef int (*PFN_AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*PFN_CommDestroy)(NcclComm);
typedef const char* (*PFN_GetErrorString)(int);
typedef int (*PFN_GetVersion)(int*);
constexpr int NCCL_FLOAT32 = 7, NCCL_SUM = 0;

struct NcclApi {
  void* handle = nullptr;
  PFN_GetUniqueId get_unique_id = nullptr;
  PFN_CommInitRank comm_init_rank = nullptr;
  PFN_AllReduce all_reduce = nullptr;
  PFN_CommDestroy comm_destroy = nullptr;
  PFN_GetErrorString error_string = nullptr;
  PFN_GetVersion get_version = nullptr;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.handle ? &api : nullptr;
  tried = true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);       // the copy this process already uses (PyTorch's)
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return nullptr;
  api.get_unique_id = (PFN_GetUniqueId)dlsym(h, "ncclGetUniqueId");
  api.comm_init_rank = (PFN_CommInitRank)dlsym(h, "ncclCommInitRank");
  api.all_reduce = (PFN_AllReduce)dlsym(h, "ncclAllReduce");
  api.comm_destroy = (PFN_CommDestroy)dlsym(h, "ncclCommDestroy");
  api.error_string = (PFN_GetErrorString)dlsym(h, "ncclGetErrorString");
  api.get_version = (PFN_GetVersion)dlsym(h, "ncclGetVersion");
  if (!api.get_unique_id || !api.comm_init_rank || !api.all_reduce || !api.comm_destroy) return nullptr;
  api.handle = h;
  return &api;
}

#define UU_NCCL(api, expr)                                                                        \
  do {                                                                                            \
    const int _r = (expr);                                                                        \
    if (_r != 0) {                                                                                \
      set_error(std::string(#expr) + " failed: " + ((api)->error_string ? (api)->error_string(_r) : "NCCL error") + \
                " (code " + std::to_string(_r) + ")");                                            \
      return 1;                                                                                   \
    }                                                                                             \
  } while (0)
}  // namespace

int comm_allreduce_range(uu_model* m, size_t lo, size_t hi, int bucket, cudaStream_t main) {
  if (!m->nccl_comm || m->comm_world <= 1 || hi <= lo) return 0;
  NcclApi* api = nccl_api();
  UU_CHECK(api && m->grads && bucket >= 0 && bucket < 4, "communicator not usable");
  UU_CUDA(cudaEventRecord(m->comm_ev_ready[bucket], main));
  UU_CUDA(cudaStreamWaitEvent(m->comm_stream, m->comm_ev_ready[bucket], 0));
  UU_NCCL(api, api->all_reduce(m->grads + lo, m->grads + lo, hi - lo, NCCL_FLOAT32, NCCL_SUM, m->nccl_comm, m->comm_stream));
  return 0;
}

int comm_allreduce_scalar(uu_model* m, float* dev_value, cudaStream_t main) {
  if (!m->nccl_comm || m->comm_world <= 1) return 0;
  NcclApi* api = nccl_api();
  UU_CHECK(api, "communicator not usable");
  UU_CUDA(cudaEventRecord(m->comm_ev_ready[3], main));
  UU_CUDA(cudaStreamWaitEvent(m->comm_stream, m->comm_ev_ready[3], 0));
  UU_NCCL(api, api->all_reduce(dev_value, dev_value, 1, NCCL_FLOAT32, NCCL_SUM, m->nccl_comm, m->comm_stream));
  return 0;
}

int comm_join(uu_model* m, cudaStream_t main) {
  if (!m->nccl_comm || m->comm_world <= 1) return 0;
  UU_CUDA(cudaEventRecord(m->comm_ev_done, m->comm_stream));
  UU_CUDA(cudaStreamWaitEvent(main, m->comm_ev_done, 0));
  return 0;
}

void comm_destroy(uu_model* m) {
  if (m->nccl_comm) {
    NcclApi* api = nccl_api();
    if (api) api->comm_destroy(m->nccl_comm);
    m->nccl_comm = nullptr;
  }
  for (auto& e : m->comm_ev_ready)
    if (e) { cudaEventDestroy(e); e = nullptr; }
  if (m->comm_ev_done) { cudaEventDestroy(m->comm_ev_done); m->comm_ev_done = nullptr; }
  if (m->comm_stream) { cudaStreamDestroy(m->comm_stream); m->comm_stream = nullptr; }
  m->comm_world = 1; m->comm_rank = 0;
}

}  // namespace uu

using namespace uu;

extern "C" {

int uu_comm_unique_id(void* id_out, int capacity) {
  UU_CHECK(id_out && capacity >= 128, "id buffer must hold 128 bytes");
  NcclApi* api = nccl_api();
  UU_CHECK(api, "NCCL (libnccl.so.2) could not be loaded");
  NcclUniqueId id;
  UU_NCCL(api, api->get_unique_id(&id));
  memcpy(id_out, id.internal, 128);
  return 0;
}

int uu_comm_init(uu_model* m, const void* unique_id, int rank, int world) {
  UU_CHECK(m && unique_id && world >= 1 && rank >= 0 && rank < world, "bad argument");
  UU_CUDA(cudaSetDevice(m->device));
  comm_destroy(m);
  if (world == 1) return 0;
  NcclApi* api = nccl_api();
  UU_CHECK(api, "NCCL (libnccl.so.2) could not be loaded");
  NcclUniqueId id;
  memcpy(id.internal, unique_id, 128);
  NcclComm comm = nullptr;
  UU_NCCL(api, api->comm_init_rank(&comm, world, id, rank));
  m->nccl_comm = comm; m->comm_rank = rank; m->comm_world = world;
  UU_CUDA(cudaStreamCreateWithFlags(&m->comm_stream, cudaStreamNonBlocking));
  for (auto& e : m->comm_ev_ready) UU_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  UU_CUDA(cudaEventCreateWithFlags(&m->comm_ev_done, cudaEventDisableTiming));
  return 0;
}

int uu_comm_destroy(uu_model* m) {
  UU_CHECK(m, "null model");
  UU_CUDA(cudaSetDevice(m->device));
  comm_destroy(m);
  return 0;
}

int uu_comm_world_size(const uu_model* m) { return m ? m->comm_world : 0; }

int uu_allreduce_gradients(uu_model* m, void* stream) {
  UU_CHECK(m && m->grads, "no gradients: run uu_train_forward_backward first");
  UU_CUDA(cudaSetDevice(m->device));
  if (comm_allreduce_range(m, 0, m->n_alloc, 0, (cudaStream_t)stream)) return 1;
  return comm_join(m, (cudaStream_t)stream);
}

}  // extern "C"
