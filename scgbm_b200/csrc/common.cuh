// common.cuh -- shared host/device infrastructure for libscgbm (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <stdexcept>
#include "../../include/scgbm.h"

namespace scgbm {

// ---------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------
struct Error : public std::runtime_error {
  int status;
  Error(int st, const std::string& m) : std::runtime_error(m), status(st) {}
};

void set_last_error(const std::string& m);

#define SCGBM_CUDA(expr)                                                               \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      char _b[512];                                                                    \
      snprintf(_b, sizeof(_b), "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e),     \
               __FILE__, __LINE__, cudaGetErrorString(_e));                            \
      throw ::scgbm::Error(SCGBM_ERR_CUDA, _b);                                        \
    }                                                                                  \
  } while (0)

#define SCGBM_REQUIRE(cond, msg)                                                       \
  do {                                                                                 \
    if (!(cond)) throw ::scgbm::Error(SCGBM_ERR_INVALID, std::string(msg));            \
  } while (0)

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
static inline int64_t ceil_div(int64_t x, int64_t m) { return (x + m - 1) / m; }

constexpr int kRowPad = 64;  // leading dimensions are multiples of 64 elements

// ---------------------------------------------------------------------------------
// device buffers
// ---------------------------------------------------------------------------------
template <typename T>
struct DevBuf {
  T* p = nullptr;
  int64_t n = 0;
  DevBuf() {}
  explicit DevBuf(int64_t n_) { alloc(n_); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void alloc(int64_t n_) {
    release();
    n = n_;
    if (n > 0) {
      cudaError_t e = cudaMalloc((void**)&p, (size_t)n * sizeof(T));
      if (e != cudaSuccess) {
        p = nullptr;
        char b[256];
        snprintf(b, sizeof(b), "cudaMalloc of %.3f GB failed: %s", (double)n * sizeof(T) / 1e9,
                 cudaGetErrorString(e));
        cudaGetLastError();
        throw Error(SCGBM_ERR_NOMEM, b);
      }
    }
  }
  void zero(cudaStream_t s) { if (n > 0) SCGBM_CUDA(cudaMemsetAsync(p, 0, (size_t)n * sizeof(T), s)); }
  void release() { if (p) { cudaFree(p); p = nullptr; } n = 0; }
  T* get() const { return p; }
};

template <typename T>
struct PinnedBuf {
  T* p = nullptr;
  int64_t n = 0;
  PinnedBuf() {}
  explicit PinnedBuf(int64_t n_) { alloc(n_); }
  PinnedBuf(const PinnedBuf&) = delete;
  PinnedBuf& operator=(const PinnedBuf&) = delete;
  ~PinnedBuf() { if (p) cudaFreeHost(p); }
  void alloc(int64_t n_) {
    if (p) { cudaFreeHost(p); p = nullptr; }
    n = n_;
    if (n > 0) SCGBM_CUDA(cudaMallocHost((void**)&p, (size_t)n * sizeof(T)));
  }
};

// ---------------------------------------------------------------------------------
// CUDA-event profiler: per kernel class summed time + launch counts
// ---------------------------------------------------------------------------------
struct Profiler {
  bool timing = false;
  bool fused_only = false;   // profile level 2: CUDA events around the K4 launches only (the roofline kernel)
  cudaStream_t stream = nullptr;
  int64_t launches[SCGBM_K_NCLASS] = {0};
  double ms[SCGBM_K_NCLASS] = {0};
  struct Rec { int cls; cudaEvent_t a, b; };
  std::vector<Rec> pending;
  std::vector<cudaEvent_t> pool;
  double h2d_bytes = 0, d2h_bytes = 0;

  cudaEvent_t get_event() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e;
    SCGBM_CUDA(cudaEventCreate(&e));
    return e;
  }
  // begin/end bracket `n` kernel launches of class `cls`
  cudaEvent_t begin(int cls, int n = 1) {
    launches[cls] += n;
    if (!timing || (fused_only && cls != SCGBM_K_FUSED)) return nullptr;
    cudaEvent_t a = get_event();
    SCGBM_CUDA(cudaEventRecord(a, stream));
    return a;
  }
  void end(int cls, cudaEvent_t a) {
    if (!timing || a == nullptr) return;
    cudaEvent_t b = get_event();
    SCGBM_CUDA(cudaEventRecord(b, stream));
    pending.push_back({cls, a, b});
    if (pending.size() > 4096) resolve();
  }
  void resolve() {
    if (pending.empty()) return;
    SCGBM_CUDA(cudaStreamSynchronize(stream));
    for (auto& r : pending) {
      float t = 0;
      cudaEventElapsedTime(&t, r.a, r.b);
      ms[r.cls] += t;
      pool.push_back(r.a);
      pool.push_back(r.b);
    }
    pending.clear();
  }
  void fill(scgbm_profile* p) {
    resolve();
    for (int k = 0; k < SCGBM_K_NCLASS; ++k) { p->ms[k] = ms[k]; p->launches[k] = launches[k]; }
    p->h2d_bytes = h2d_bytes;
    p->d2h_bytes = d2h_bytes;
  }
  ~Profiler() {
    for (auto& r : pending) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : pool) cudaEventDestroy(e);
  }
};

struct ProfScope {
  Profiler& p; int cls; cudaEvent_t a;
  ProfScope(Profiler& p_, int cls_, int n = 1) : p(p_), cls(cls_) { a = p.begin(cls, n); }
  ~ProfScope() {
    try { p.end(cls, a); } catch (...) {}
  }
};

struct Comm;  // comm.cu

// per-call execution context
struct Ctx {
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  Profiler prof;
  Comm* comm = nullptr;  // may be null
  int rank = 0, nranks = 1;
  explicit Ctx(int device_, Comm* comm_ = nullptr);
  ~Ctx();
  void sync() { SCGBM_CUDA(cudaStreamSynchronize(stream)); }
  // collectives on device buffers (no-ops when single rank)
  void allreduce_sum(double* dev, int64_t n);
  void allreduce_max(double* dev, int64_t n);
  // every rank ends up with rank 0's copy of a replicated quantity, bit for bit (no-op on a single rank)
  void bcast0(double* dev, int64_t n);
  // diagnostics (SCGBM_RANKCHECK): how far the ranks' copies of a replicated quantity are apart
  // (weighted checksum, max - min over ranks; 0 when they are identical); prints on rank 0
  void rank_spread(const char* what, int call, const double* dev, int64_t n);
};

#define SCGBM_LAUNCH_CHECK() SCGBM_CUDA(cudaGetLastError())

// ---------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// 2^x on the SFU (MUFU.EX2), ~2 ulp
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// exp(x) = 2^(x*log2e) with log2e split in two so the constant carries no
// systematic relative error
__device__ __forceinline__ float exp_fast_accurate(float x) {
  const float L_hi = 1.4426950216293335f;     // fl32(log2 e)
  const float L_lo = 1.9259629911e-08f;       // log2 e - L_hi
  float t = x * L_hi;
  t = fmaf(x, L_lo, t);
  return ex2_approx(t);
}
#endif

}  // namespace scgbm
