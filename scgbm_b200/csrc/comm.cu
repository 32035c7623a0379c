// comm.cu -- execution context + multi-GPU plumbing.  One rank per GPU; cells
// (columns of Y) are sharded over ranks.  NCCL is loaded lazily with
// dlopen so the single-GPU path has no NCCL dependency at all.
#include <dlfcn.h>
#include <mutex>
#include <nccl.h>   // types only; symbols are resolved with dlsym
#include "common.cuh"

namespace scgbm {

static thread_local std::string g_last_error;
void set_last_error(const std::string& m) { g_last_error = m; }
const char* get_last_error() { return g_last_error.c_str(); }

struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.h) break;
    }
    if (!api.h) return;
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.h, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.h, "ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.h, "ncclCommDestroy");
    api.AllReduce = (decltype(api.AllReduce))dlsym(api.h, "ncclAllReduce");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.h, "ncclGetErrorString");
  });
  if (!api.h || !api.GetUniqueId || !api.CommInitRank || !api.AllReduce)
    throw Error(SCGBM_ERR_CUDA, "NCCL (libnccl.so.2) could not be loaded");
  return api;
}

#define SCGBM_NCCL(expr)                                                               \
  do {                                                                                 \
    ncclResult_t _r = (expr);                                                          \
    if (_r != ncclSuccess) {                                                           \
      char _b[512];                                                                    \
      snprintf(_b, sizeof(_b), "NCCL error %d at %s:%d: %s", (int)_r, __FILE__, __LINE__, \
               nccl().GetErrorString ? nccl().GetErrorString(_r) : "?");               \
      throw ::scgbm::Error(SCGBM_ERR_CUDA, _b);                                        \
    }                                                                                  \
  } while (0)

struct Comm {
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0, device = 0;
  // Several ranks inside ONE process (multi.cu: one host thread per GPU) share the CUDA driver's process-wide locks:
  // a cudaFree on this GPU (implicit device synchronisation, lock held) issued while this GPU's NCCL kernel is
  // still waiting for the peer's kernel deadlocks as soon as the peer's launch needs the same lock.  Such
  // communicators therefore complete every collective before the caller goes on (one stream sync per all-reduce,
  // a few tens of microseconds); ranks in separate processes keep the fully asynchronous path.
  bool in_process = false;
};

Ctx::Ctx(int device_, Comm* comm_) : comm(comm_) {
  int cur = 0;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    throw Error(SCGBM_ERR_CUDA, "no CUDA device available: libscgbm has no CPU fallback");
  }
  if (device_ < 0) { SCGBM_CUDA(cudaGetDevice(&cur)); device_ = cur; }
  if (comm_) device_ = comm_->device;
  device = device_;
  SCGBM_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  SCGBM_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    throw Error(SCGBM_ERR_CUDA, "libscgbm is built for sm_100a (Blackwell) only");
  num_sms = prop.multiProcessorCount;
  SCGBM_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  prof.stream = stream;
  if (comm) { rank = comm->rank; nranks = comm->nranks; }
}

Ctx::~Ctx() {
  if (stream) { cudaStreamSynchronize(stream); cudaStreamDestroy(stream); }
}

void Ctx::allreduce_sum(double* dev, int64_t n) {
  if (!comm || nranks == 1 || n <= 0) return;
  ProfScope ps(prof, SCGBM_K_COMM, 1);
  SCGBM_NCCL(nccl().AllReduce(dev, dev, (size_t)n, ncclFloat64, ncclSum, comm->comm, stream));
  if (comm->in_process) SCGBM_CUDA(cudaStreamSynchronize(stream));
}
void Ctx::allreduce_max(double* dev, int64_t n) {
  if (!comm || nranks == 1 || n <= 0) return;
  ProfScope ps(prof, SCGBM_K_COMM, 1);
  SCGBM_NCCL(nccl().AllReduce(dev, dev, (size_t)n, ncclFloat64, ncclMax, comm->comm, stream));
  if (comm->in_process) SCGBM_CUDA(cudaStreamSynchronize(stream));
}

Comm* comm_create(const uint8_t* id, int nranks, int rank, int device) {
  SCGBM_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank/nranks");
  SCGBM_CUDA(cudaSetDevice(device));
  Comm* c = new Comm();
  c->nranks = nranks; c->rank = rank; c->device = device;
  ncclUniqueId uid;
  static_assert(sizeof(uid) == SCGBM_UNIQUE_ID_BYTES, "unique id size");
  memcpy(&uid, id, sizeof(uid));
  try {
    SCGBM_NCCL(nccl().CommInitRank(&c->comm, nranks, uid, rank));
  } catch (...) { delete c; throw; }
  return c;
}
void comm_unique_id(uint8_t* id) {
  ncclUniqueId uid;
  SCGBM_NCCL(nccl().GetUniqueId(&uid));
  memcpy(id, &uid, sizeof(uid));
}
void comm_destroy(Comm* c) {
  if (!c) return;
  if (c->comm && nccl().CommDestroy) nccl().CommDestroy(c->comm);
  delete c;
}
void comm_set_in_process(Comm* c) { if (c) c->in_process = true; }
int comm_rank(Comm* c) { return c ? c->rank : 0; }
int comm_nranks(Comm* c) { return c ? c->nranks : 1; }
int comm_device(Comm* c) { return c ? c->device : -1; }

}  // namespace scgbm
