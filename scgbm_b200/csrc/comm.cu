// comm.cu -- execution context + multi-GPU plumbing.  One rank per GPU; cells
// (columns of Y) are sharded over ranks.  NCCL is loaded lazily with
// dlopen so the single-GPU path has no NCCL dependency at all.
#include <dlfcn.h>
#include <condition_variable>
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

// Several ranks inside ONE process (multi.cu: one host thread per GPU).  NCCL enables peer access between the GPUs,
// and from then on a cudaFree on one GPU also waits for the kernels of its peers; a rank that frees a buffer while a
// peer's all-reduce kernel is already spinning for this rank's (not yet launched) kernel dead-locks the process --
// observed as a hang of gbm_sc(ngpu = 2) in round 2.  In-process communicators therefore (1) rendezvous on a host
// barrier before every collective, so that no rank is inside a synchronising CUDA call while another has a
// collective in flight, and (2) complete the collective before going on.  Ranks in separate processes keep the
// fully asynchronous path.  A rank that fails marks the group aborted, which releases the others with an error.
struct InprocGroup {
  std::mutex m;
  std::condition_variable cv;
  int n = 0, count = 0;
  uint64_t gen = 0;
  bool aborted = false;
  int finished = 0;                 // ranks that have left their body (successfully or not)
};

struct Comm {
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0, device = 0;
  InprocGroup* group = nullptr;     // non-null: ranks are threads of this process
};

static void inproc_arrive(Comm* c) {
  InprocGroup* g = c->group;
  std::unique_lock<std::mutex> lk(g->m);
  if (g->aborted) throw Error(SCGBM_ERR_CUDA, "another rank of this call failed; no rank proceeds");
  const uint64_t gen0 = g->gen;
  if (++g->count == g->n) {
    g->count = 0; ++g->gen;
    g->cv.notify_all();
    return;
  }
  g->cv.wait(lk, [&] { return g->gen != gen0 || g->aborted; });
  if (g->gen == gen0) throw Error(SCGBM_ERR_CUDA, "another rank of this call failed; no rank proceeds");
}

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

// One NCCL all-reduce in place.  Ranks that are threads of one process rendezvous on the host first and complete the
// collective before going on (see InprocGroup above).
static void nccl_allreduce(Ctx& ctx, double* dev, int64_t n, ncclRedOp_t op) {
  Comm* comm = ctx.comm;
  if (comm->group) inproc_arrive(comm);
  SCGBM_NCCL(nccl().AllReduce(dev, dev, (size_t)n, ncclFloat64, op, comm->comm, ctx.stream));
  if (comm->group) SCGBM_CUDA(cudaStreamSynchronize(ctx.stream));
}

// Sum over ranks.  A sum all-reduce promises every rank A correctly rounded sum of the contributions, not the SAME
// one: from four ranks on the association order -- hence the last bit -- may differ from rank to rank (two ranks
// cannot differ: a + b = b + a).  Nothing downstream may therefore assume bit-identical copies:
//   * host decisions are taken on max-reduced scalars (fit.cu, tsvd.cu) -- max IS exactly associative;
//   * the replicated state of the SVD, where a last-bit difference is carried forward and amplified, is imposed
//     from rank 0 (bcast0; svd_bcast in tsvd.cu has the measurement).
// SCGBM_SUM_MAX=1 (experiments) additionally follows every sum by a max all-reduce of the result, which makes the
// copies bit-identical at the price of a second latency-bound collective per site; measured on 4 ranks, that alone did
// not stop the drift (the eigensolver's output differed between GPUs even for bit-identical input).
void Ctx::allreduce_sum(double* dev, int64_t n) {
  if (!comm || nranks == 1 || n <= 0) return;
  static const bool sum_max = getenv("SCGBM_SUM_MAX") != nullptr;
  ProfScope ps(prof, SCGBM_K_COMM, (sum_max && nranks > 2) ? 2 : 1);
  nccl_allreduce(*this, dev, n, ncclSum);
  if (sum_max && nranks > 2) nccl_allreduce(*this, dev, n, ncclMax);
}
void Ctx::allreduce_max(double* dev, int64_t n) {
  if (!comm || nranks == 1 || n <= 0) return;
  ProfScope ps(prof, SCGBM_K_COMM, 1);
  nccl_allreduce(*this, dev, n, ncclMax);
}

void Ctx::bcast0(double* dev, int64_t n) {
  if (!comm || nranks == 1 || n <= 0) return;
  ProfScope ps(prof, SCGBM_K_COMM, 1);
  if (rank != 0) SCGBM_CUDA(cudaMemsetAsync(dev, 0, (size_t)n * sizeof(double), stream));
  nccl_allreduce(*this, dev, n, ncclSum);        // x + 0 + ... + 0 is exact in any order: rank 0's bits everywhere
}

__global__ void __launch_bounds__(1024) checksum_kernel(int64_t n, const double* __restrict__ x, double* __restrict__ out) {
  __shared__ double red[1024];
  double s = 0.0, a = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) { s += x[i] * (double)(1 + (i % 7)); a += fabs(x[i]); }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int st = 512; st > 0; st >>= 1) { if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st]; __syncthreads(); }
  const double cs = red[0];
  __syncthreads();
  red[threadIdx.x] = a;
  __syncthreads();
  for (int st = 512; st > 0; st >>= 1) { if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st]; __syncthreads(); }
  if (threadIdx.x == 0) { out[0] = cs; out[1] = -cs; out[2] = red[0]; }
}

void Ctx::rank_spread(const char* what, int call, const double* dev, int64_t n) {
  static const bool on = getenv("SCGBM_RANKCHECK") != nullptr;
  if (!on || !comm || nranks == 1 || n <= 0) return;
  double* d = nullptr;
  SCGBM_CUDA(cudaMalloc((void**)&d, 4 * sizeof(double)));
  checksum_kernel<<<1, 1024, 0, stream>>>(n, dev, d);
  allreduce_max(d, 3);
  double h[3];
  SCGBM_CUDA(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, stream));
  sync();
  cudaFree(d);
  const double spread = h[0] + h[1];
  if (rank == 0 && spread != 0.0)
    fprintf(stderr, "[scgbm-rankcheck] call %d %s: checksum spread %.3e (sum |x| %.3e, n %lld)\n", call, what, spread, h[2], (long long)n);
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
InprocGroup* inproc_group_create(int n) { InprocGroup* g = new InprocGroup(); g->n = n; return g; }
void inproc_group_destroy(InprocGroup* g) { delete g; }
void inproc_group_abort(InprocGroup* g) {
  std::lock_guard<std::mutex> lk(g->m);
  g->aborted = true;
  g->cv.notify_all();
}
// every rank calls this once when it is done (or has failed): returns when all ranks are, i.e. when no collective
// can be in flight any more and the communicators may be torn down
void inproc_group_finish(InprocGroup* g) {
  std::unique_lock<std::mutex> lk(g->m);
  ++g->finished;
  g->cv.notify_all();
  g->cv.wait(lk, [&] { return g->finished >= g->n; });
}
void comm_set_in_process(Comm* c, InprocGroup* g) { if (c) c->group = g; }
int comm_rank(Comm* c) { return c ? c->rank : 0; }
int comm_nranks(Comm* c) { return c ? c->nranks : 1; }
int comm_device(Comm* c) { return c ? c->device : -1; }

}  // namespace scgbm
