// api.cu -- the extern "C" surface declared in include/scgbm.h.
#include "common.cuh"
#include "counts.cuh"
#include "dense_small.cuh"
#include "eta_pass.cuh"
#include "tsvd.cuh"

namespace scgbm {
const char* get_last_error();
Comm* comm_create(const uint8_t* id, int nranks, int rank, int device);
void comm_unique_id(uint8_t* id);
void comm_destroy(Comm* c);
void fit_impl(const scgbm_fit_args* a, scgbm_fit_out* out);
void fit_multi_impl(const scgbm_fit_args* a, scgbm_fit_out* out);
void get_se_multi_impl(const scgbm_se_args* a, scgbm_se_out* out);
struct FitRun;
FitRun* fit_begin_impl(const scgbm_fit_args* a);
int fit_step_impl(FitRun* r, int n_steps);
void fit_profile_impl(FitRun* r, scgbm_profile* p);
void fit_set_profile_impl(FitRun* r, int level);
void fit_end_impl(FitRun* r, scgbm_fit_out* out);
void get_se_impl(const scgbm_se_args* a, scgbm_se_out* out);
void project_impl(const scgbm_proj_args* a, scgbm_proj_out* out);
void sim_counts(const scgbm_sim_args* a, void* Y_dev, int64_t* n_saturated);
void sim_data_impl(const scgbm_simdata_args* a, scgbm_simdata_out* out);
void cci_perturb_impl(const scgbm_cci_args* a, double* out);
void dbg_tsvd_impl(const scgbm_tsvd_args* a, scgbm_tsvd_out* out);
void dbg_eval_impl(const scgbm_eval_args* a, scgbm_eval_out* out);
void dbg_prod_impl(const scgbm_prod_args* a, scgbm_prod_out* out);
}  // namespace scgbm

using namespace scgbm;

template <typename F>
static int guarded(F&& f) {
  try {
    f();
    return SCGBM_OK;
  } catch (const Error& e) {
    set_last_error(e.what());
    cudaGetLastError();
    return e.status;
  } catch (const std::bad_alloc&) {
    set_last_error("host allocation failed");
    return SCGBM_ERR_NOMEM;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return SCGBM_ERR_CUDA;
  }
}

static void use_device(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    throw Error(SCGBM_ERR_CUDA, "no CUDA device available: libscgbm has no CPU fallback");
  }
  if (device >= 0) SCGBM_CUDA(cudaSetDevice(device));
}

// each thread: NR 16-byte streaming loads and NW 16-byte streaming stores per step, 4 steps in flight
template <int NR, int NW>
__global__ void __launch_bounds__(256) membench_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n_steps) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t s0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s0 < n_steps; s0 += 4 * stride) {
    uint4 v[4][NR];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const size_t s = s0 + (size_t)u * stride;
        v[u][r] = s < n_steps ? __ldcs(in + s + (size_t)r * n_steps) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const size_t s = s0 + (size_t)u * stride;
      if (s >= n_steps) continue;
      uint4 a = v[u][0];
#pragma unroll
      for (int r = 1; r < NR; ++r) { a.x ^= v[u][r].x; a.y += v[u][r].y; }
#pragma unroll
      for (int w = 0; w < NW; ++w) __stcs(out + s + (size_t)w * n_steps, make_uint4(a.x + w, a.y, a.z, a.w));
    }
  }
}

extern "C" {

int scgbm_fit(const scgbm_fit_args* args, scgbm_fit_out* out) {
  return guarded([&] { if (args && args->ngpu > 1) fit_multi_impl(args, out); else fit_impl(args, out); });
}
int scgbm_fit_begin(const scgbm_fit_args* args, scgbm_fit_handle** h) {
  return guarded([&] { SCGBM_REQUIRE(h, "NULL handle pointer"); *h = reinterpret_cast<scgbm_fit_handle*>(fit_begin_impl(args)); });
}
int scgbm_fit_step(scgbm_fit_handle* h, int32_t n_steps, int32_t* done) {
  return guarded([&] { const int d = fit_step_impl(reinterpret_cast<FitRun*>(h), n_steps); if (done) *done = d; });
}
int scgbm_fit_profile(scgbm_fit_handle* h, scgbm_profile* prof) {
  return guarded([&] { fit_profile_impl(reinterpret_cast<FitRun*>(h), prof); });
}
int scgbm_fit_set_profile(scgbm_fit_handle* h, int32_t level) {
  return guarded([&] { fit_set_profile_impl(reinterpret_cast<FitRun*>(h), level); });
}
int scgbm_fit_end(scgbm_fit_handle* h, scgbm_fit_out* out) {
  return guarded([&] { fit_end_impl(reinterpret_cast<FitRun*>(h), out); });
}
int scgbm_get_se(const scgbm_se_args* args, scgbm_se_out* out) {
  return guarded([&] { if (args && args->ngpu > 1) get_se_multi_impl(args, out); else get_se_impl(args, out); });
}
int scgbm_project(const scgbm_proj_args* args, scgbm_proj_out* out) { return guarded([&] { project_impl(args, out); }); }
int scgbm_sim_counts(const scgbm_sim_args* args, void* Y_dev, int64_t* n_saturated) {
  return guarded([&] { SCGBM_REQUIRE(args && Y_dev, "NULL args"); sim_counts(args, Y_dev, n_saturated); });
}
int scgbm_sim_data(const scgbm_simdata_args* args, scgbm_simdata_out* out) { return guarded([&] { sim_data_impl(args, out); }); }
int scgbm_cci_perturb(const scgbm_cci_args* args, double* out) { return guarded([&] { cci_perturb_impl(args, out); }); }

int scgbm_dev_alloc(void** p, int64_t bytes, int device) {
  return guarded([&] {
    SCGBM_REQUIRE(p && bytes >= 0, "bad args");
    use_device(device);
    cudaError_t e = cudaMalloc(p, (size_t)bytes);
    if (e != cudaSuccess) { cudaGetLastError(); throw Error(SCGBM_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e)); }
  });
}
int scgbm_dev_free(void* p, int device) {
  return guarded([&] { use_device(device); SCGBM_CUDA(cudaFree(p)); });
}
int scgbm_memcpy_d2h(void* dst_host, const void* src_dev, int64_t bytes, int device) {
  return guarded([&] { use_device(device); SCGBM_CUDA(cudaMemcpy(dst_host, src_dev, (size_t)bytes, cudaMemcpyDeviceToHost)); });
}
int scgbm_memcpy_h2d(void* dst_dev, const void* src_host, int64_t bytes, int device) {
  return guarded([&] { use_device(device); SCGBM_CUDA(cudaMemcpy(dst_dev, src_host, (size_t)bytes, cudaMemcpyHostToDevice)); });
}
int scgbm_host_alloc_pinned(void** p, int64_t bytes) {
  return guarded([&] { use_device(-1); SCGBM_CUDA(cudaMallocHost(p, (size_t)bytes)); });
}
int scgbm_host_free_pinned(void* p) {
  return guarded([&] { SCGBM_CUDA(cudaFreeHost(p)); });
}

__global__ void flush_fill_kernel(int64_t n, float* p, float v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}
int scgbm_flush_l2(int device) {
  return guarded([&] {
    use_device(device);
    static thread_local void* buf = nullptr;
    static thread_local int buf_dev = -1;
    int dev = 0;
    SCGBM_CUDA(cudaGetDevice(&dev));
    const int64_t n = (int64_t)64 << 20;  // 256 MB of floats > 126 MB L2
    if (!buf || buf_dev != dev) { SCGBM_CUDA(cudaMalloc(&buf, n * sizeof(float))); buf_dev = dev; }
    flush_fill_kernel<<<1184, 256>>>(n, (float*)buf, 1.0f);
    SCGBM_CUDA(cudaGetLastError());
    SCGBM_CUDA(cudaDeviceSynchronize());
  });
}

int scgbm_comm_unique_id(uint8_t id[SCGBM_UNIQUE_ID_BYTES]) { return guarded([&] { comm_unique_id(id); }); }
int scgbm_comm_init_rank(scgbm_comm** comm, const uint8_t id[SCGBM_UNIQUE_ID_BYTES], int nranks, int rank, int device) {
  return guarded([&] {
    SCGBM_REQUIRE(comm, "NULL comm");
    use_device(device);
    *comm = reinterpret_cast<scgbm_comm*>(comm_create(id, nranks, rank, device));
  });
}
int scgbm_comm_destroy(scgbm_comm* comm) {
  return guarded([&] { comm_destroy(reinterpret_cast<Comm*>(comm)); });
}
int scgbm_comm_allreduce_f64(scgbm_comm* comm, double* host_buf, int64_t n, int op_max) {
  return guarded([&] {
    SCGBM_REQUIRE(comm && host_buf && n > 0, "bad args");
    Ctx ctx(-1, reinterpret_cast<Comm*>(comm));
    DevBuf<double> d(n);
    SCGBM_CUDA(cudaMemcpyAsync(d.get(), host_buf, n * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
    if (op_max) ctx.allreduce_max(d.get(), n); else ctx.allreduce_sum(d.get(), n);
    SCGBM_CUDA(cudaMemcpyAsync(host_buf, d.get(), n * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
    ctx.sync();
  });
}

const char* scgbm_last_error(void) { return get_last_error(); }
int scgbm_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
int scgbm_version(void) { return SCGBM_VERSION; }

int scgbm_dbg_membench(int32_t device, double gbs[2]) {
  return guarded([&] {
    SCGBM_REQUIRE(gbs, "NULL out");
    Ctx ctx(device, nullptr);
    const size_t n_steps = (size_t)1 << 26;      // 1 GiB per stream
    DevBuf<uint4> in((int64_t)n_steps), out((int64_t)n_steps * 2);
    in.zero(ctx.stream);
    cudaEvent_t a, b;
    SCGBM_CUDA(cudaEventCreate(&a)); SCGBM_CUDA(cudaEventCreate(&b));
    const int grid = ctx.num_sms * 8;
    for (int which = 0; which < 2; ++which) {
      float best = 1e30f;
      for (int rep = 0; rep < 4; ++rep) {
        SCGBM_CUDA(cudaEventRecord(a, ctx.stream));
        if (which == 0) membench_kernel<1, 1><<<grid, 256, 0, ctx.stream>>>(in.get(), out.get(), n_steps);
        else membench_kernel<1, 2><<<grid, 256, 0, ctx.stream>>>(in.get(), out.get(), n_steps);
        SCGBM_CUDA(cudaEventRecord(b, ctx.stream));
        SCGBM_CUDA(cudaEventSynchronize(b));
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        if (rep > 0) best = std::min(best, ms);
      }
      gbs[which] = (double)n_steps * 16.0 * (which == 0 ? 2.0 : 3.0) / (best * 1e-3) / 1e9;
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
  });
}

int scgbm_dbg_tsvd(const scgbm_tsvd_args* args, scgbm_tsvd_out* out) { return guarded([&] { dbg_tsvd_impl(args, out); }); }
int scgbm_dbg_prod(const scgbm_prod_args* args, scgbm_prod_out* out) { return guarded([&] { dbg_prod_impl(args, out); }); }
int scgbm_dbg_eval(const scgbm_eval_args* args, scgbm_eval_out* out) { return guarded([&] { dbg_eval_impl(args, out); }); }
int scgbm_dbg_eigh(double* A, double* w, int32_t n, int32_t device) {
  return guarded([&] {
    SCGBM_REQUIRE(A && w && n >= 1, "bad args");
    Ctx ctx(device, nullptr);
    SmallWs ws;
    DevBuf<double> dA((int64_t)n * n), dw(n), dV((int64_t)n * n);
    SCGBM_CUDA(cudaMemcpyAsync(dA.get(), A, (size_t)n * n * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
    eigh_desc(ctx, ws, n, dA.get(), dw.get(), dV.get());
    SCGBM_CUDA(cudaMemcpyAsync(A, dV.get(), (size_t)n * n * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
    SCGBM_CUDA(cudaMemcpyAsync(w, dw.get(), (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
    ctx.sync();
  });
}

}  // extern "C"
