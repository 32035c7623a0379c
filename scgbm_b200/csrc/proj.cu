// proj.cu -- K9: the projection step of gbm.sc(subset = ...) (R/fit.R:229-270):
// one Poisson GLM per cell, design [1 | U] (Ik x (M+1)), offset alpha, solved by the
// IRLS iteration fastglm(method = 3) runs [ext]: start mu = y + 0.1, WLS step by
// Cholesky, stop at |dev - devold| / (|dev| + 0.1) < tol, SE from the last solve's
// X^T W X.  Eight cells share a CTA so each x_a x_b product feeds eight weights.
#include "wgram.cuh"
#include "counts.cuh"
#include "proj_tc.cuh"

namespace scgbm {

struct ProjParams {
  int64_t Ik, J, ldY;
  const void* Y;          // packed counts, ldY x J
  const double* X;        // Ik x p design (col-major, ld Ik), column 0 = 1
  const double* off;      // Ik offsets
  int p;
  double tol; int maxit;
  double* coef;           // J x p   (col-major, ld J)
  double* se;             // J x p
  int8_t* fail;           // J
  int32_t* iters;         // J
  int64_t ncell;          // cells this launch fits: J, or the length of `list`
  const int64_t* list;    // null: cells 0..J-1; else the cells to fit (the tensor-core path's stragglers, proj_tc.cu)
};

template <typename YT> __device__ __forceinline__ double y_at(const void* Y, int64_t idx) {
  return (double)reinterpret_cast<const YT*>(Y)[idx];
}

template <int PT, typename YT>
__global__ void __launch_bounds__(WG_NT)
proj_irls_kernel(const ProjParams q) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int p = q.p, T = tri_count(p), ldf = p + 1;
  double* F_s = reinterpret_cast<double*>(smraw);                 // [WG_TR][p+1]
  double* w_s = F_s + WG_TR * ldf;                                // [WG_TR][CC]
  double* wz_s = w_s + WG_TR * WG_CC;                             // [WG_TR][CC]
  double* b_s = wz_s + WG_TR * WG_CC;                             // [CC][p]   current coefficients
  double* A_s = b_s + WG_CC * p;                                  // [CC][p*p] normal matrices
  double* rhs_s = A_s + (size_t)WG_CC * p * p;                    // [CC][p]
  double* se_s = rhs_s + WG_CC * p;                               // [CC][p]
  double* dev_s = se_s + WG_CC * p;                               // [CC] current deviance
  double* devold_s = dev_s + WG_CC;                               // [CC]
  int* state_s = reinterpret_cast<int*>(devold_s + WG_CC);        // [CC] 0 running, 1 converged, 2 failed
  int* iters_s = state_s + WG_CC;
  unsigned short* tab_a = reinterpret_cast<unsigned short*>(iters_s + WG_CC);
  unsigned short* tab_b = tab_a + T;
  fill_pair_table(p, tab_a, tab_b);
  const int64_t c0 = (int64_t)blockIdx.x * WG_CC;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;   // warp = cell slot in the eta phase
  if (threadIdx.x < WG_CC) {
    state_s[threadIdx.x] = (c0 + threadIdx.x < q.ncell) ? 0 : 1;
    iters_s[threadIdx.x] = 0;
    dev_s[threadIdx.x] = 0.0; devold_s[threadIdx.x] = 0.0;
  }
  for (int idx = threadIdx.x; idx < WG_CC * p; idx += WG_NT) { b_s[idx] = 0.0; se_s[idx] = 0.0; }
  __syncthreads();

  for (int it = 0; it <= q.maxit; ++it) {
    double acc[PT][WG_CC];
#pragma unroll
    for (int pi = 0; pi < PT; ++pi)
#pragma unroll
      for (int c = 0; c < WG_CC; ++c) acc[pi][c] = 0.0;
    double racc[WG_CC];        // rhs: threads a < p own coefficient a
#pragma unroll
    for (int c = 0; c < WG_CC; ++c) racc[c] = 0.0;
    double dev_part = 0.0;     // this thread's (gene lane, cell = warp) share of the deviance
    const bool cell_live = state_s[warp] == 0;
    const int64_t cell = (c0 + warp < q.ncell) ? (q.list ? q.list[c0 + warp] : c0 + warp) : 0;

    for (int64_t r0 = 0; r0 < q.Ik; r0 += WG_TR) {
      __syncthreads();
      for (int idx = threadIdx.x; idx < WG_TR * p; idx += WG_NT) {
        const int r = idx % WG_TR, a = idx / WG_TR;
        F_s[r * ldf + a] = (r0 + r < q.Ik) ? q.X[(int64_t)a * q.Ik + r0 + r] : 0.0;
      }
      __syncthreads();
      // eta / mu / weights for (gene = lane, cell = warp)
      {
        const int64_t g = r0 + lane;
        double w = 0.0, wz = 0.0;
        if (g < q.Ik && cell_live) {
          const double y = y_at<YT>(q.Y, cell * q.ldY + g);
          const double o = q.off[g];
          double eta, mu;
          if (it == 0) { mu = y + 0.1; eta = log(mu); }           // poisson()$initialize
          else {
            eta = o;
            for (int a = 0; a < p; ++a) eta = fma(F_s[lane * ldf + a], b_s[warp * p + a], eta);
            mu = exp(eta);
          }
          dev_part += 2.0 * ((y > 0.0 ? y * log(y / mu) : 0.0) - (y - mu));
          w = mu;
          wz = mu * (eta - o) + (y - mu);
        }
        w_s[lane * WG_CC + warp] = w;
        wz_s[lane * WG_CC + warp] = wz;
      }
      __syncthreads();
      wgram_tile<PT>(F_s, ldf, w_s, tab_a, tab_b, T, WG_TR, acc);
      if (threadIdx.x < p) {
        for (int r = 0; r < WG_TR; ++r) {
          const double x = F_s[r * ldf + threadIdx.x];
#pragma unroll
          for (int c = 0; c < WG_CC; ++c) racc[c] = fma(x, wz_s[r * WG_CC + c], racc[c]);
        }
      }
    }
    // deviance of the coefficients that produced this pass's mu
    dev_part = warp_sum(dev_part);
    __syncthreads();
    if (lane == 0 && cell_live) { devold_s[warp] = dev_s[warp]; dev_s[warp] = dev_part; }
    // scatter the normal equations
#pragma unroll
    for (int pi = 0; pi < PT; ++pi) {
      const int idx = threadIdx.x + pi * WG_NT;
      if (idx >= T) continue;
      const int a = tab_a[idx], b = tab_b[idx];
#pragma unroll
      for (int c = 0; c < WG_CC; ++c) {
        A_s[(size_t)c * p * p + b * p + a] = acc[pi][c];
        A_s[(size_t)c * p * p + a * p + b] = acc[pi][c];
      }
    }
    if (threadIdx.x < p) {
#pragma unroll
      for (int c = 0; c < WG_CC; ++c) rhs_s[c * p + threadIdx.x] = racc[c];
    }
    __syncthreads();
    // per-cell control + solve: warp c owns cell c
    if (cell_live) {
      bool done = false;
      if (it > 0) {
        const double dev = dev_s[warp], devold = devold_s[warp];
        if (!isfinite(dev)) { if (lane == 0) state_s[warp] = 2; done = true; }
        else if (fabs(dev - devold) / (fabs(dev) + 0.1) < q.tol) { if (lane == 0) state_s[warp] = 1; done = true; }
      }
      if (!done && it == q.maxit) { if (lane == 0) state_s[warp] = 1; done = true; }   // maxit reached: keep last
      if (!done) {
        double* A = A_s + (size_t)warp * p * p;
        const bool ok = warp_cholesky(A, p, lane);
        if (!ok) { if (lane == 0) state_s[warp] = 2; }
        else {
          warp_chol_solve(A, rhs_s + warp * p, p, lane);
          warp_inv_diag(A, p, lane, se_s + warp * p);
          for (int a = lane; a < p; a += 32) b_s[warp * p + a] = rhs_s[warp * p + a];
          if (lane == 0) iters_s[warp] = it + 1;
        }
      }
    }
    __syncthreads();
    bool any = false;
    for (int c = 0; c < WG_CC; ++c) any |= (state_s[c] == 0);
    if (!any) break;
  }
  // outputs
  for (int idx = threadIdx.x; idx < WG_CC * p; idx += WG_NT) {
    const int c = idx / p, a = idx % p;
    if (c0 + c >= q.ncell) continue;
    const int64_t cellj = q.list ? q.list[c0 + c] : c0 + c;
    const bool bad = state_s[c] == 2;
    q.coef[(int64_t)a * q.J + cellj] = bad ? 0.0 : b_s[c * p + a];
    q.se[(int64_t)a * q.J + cellj] = bad ? 0.0 : sqrt(se_s[c * p + a]);
  }
  if (threadIdx.x < WG_CC && c0 + threadIdx.x < q.ncell) {
    const int64_t cellj = q.list ? q.list[c0 + threadIdx.x] : c0 + threadIdx.x;
    q.fail[cellj] = state_s[threadIdx.x] == 2 ? 1 : 0;
    if (q.iters) q.iters[cellj] = iters_s[threadIdx.x];
  }
}

__global__ void build_design_kernel(int64_t Ik, int M, const double* __restrict__ U, double* __restrict__ X) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Ik * (M + 1)) return;
  const int64_t r = idx % Ik;
  const int a = (int)(idx / Ik);
  X[idx] = a == 0 ? 1.0 : U[(int64_t)(a - 1) * Ik + r];       // cbind(1, U)  R/fit.R:231
}

template <int PT>
static void launch_proj(Ctx& ctx, const ProjParams& q, int storage, unsigned grid, size_t smem) {
#define PROJ_LAUNCH(YT)                                                                                 \
  do {                                                                                                  \
    if (smem > 48 * 1024)                                                                               \
      SCGBM_CUDA(cudaFuncSetAttribute(proj_irls_kernel<PT, YT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    proj_irls_kernel<PT, YT><<<grid, WG_NT, smem, ctx.stream>>>(q);                                     \
  } while (0)
  switch (storage) {
    case SCGBM_STORE_U8: PROJ_LAUNCH(uint8_t); break;
    case SCGBM_STORE_U16: PROJ_LAUNCH(uint16_t); break;
    default: PROJ_LAUNCH(float); break;
  }
#undef PROJ_LAUNCH
  SCGBM_LAUNCH_CHECK();
}

void project_impl(const scgbm_proj_args* a, scgbm_proj_out* out) {
  SCGBM_REQUIRE(a && out && (a->Y || (a->csc_i && a->csc_p && a->csc_x)) && a->ixs && a->U && a->alpha, "NULL args");
  SCGBM_REQUIRE(out->beta && out->scores && out->se_scores && out->fail, "output buffers missing");
  SCGBM_REQUIRE(a->Ik >= 1 && a->J >= 1 && a->M >= 1, "bad shape");
  Ctx ctx(a->device, nullptr);
  ctx.prof.timing = a->profile != 0;
  cudaEvent_t ev0, ev1;
  SCGBM_CUDA(cudaEventCreate(&ev0)); SCGBM_CUDA(cudaEventCreate(&ev1));
  SCGBM_CUDA(cudaEventRecord(ev0, ctx.stream));
  const int64_t Ik = a->Ik, J = a->J;
  const int M = a->M, p = M + 1, T = tri_count(p);
  const int PT = (int)ceil_div(T, WG_NT);
  SCGBM_REQUIRE(PT <= 9 && p <= WG_PMAX, "projection: M too large (M <= 63)");
  Counts Y;
  DevBuf<double> rs_full;
  if (out->rowsum_full) rs_full.alloc(a->I_full);
  if (a->Y) {
    load_counts_rows(ctx, a->Y, a->y_dtype, a->y_on_device, a->I_full, J, a->ldY, a->ixs, Ik, Y, rs_full.get());   // Y[ixs,]  R/fit.R:237
  } else {
    // sparse Y: densify all genes on the device first (scatter), then gather the kept rows from there
    Counts full;
    load_counts_csc(ctx, a->csc_i, a->csc_p, a->csc_x, a->csc_x_dtype, a->I_full, J, nullptr, 1, SCGBM_STORE_AUTO, full);
    const int dt = full.storage == SCGBM_STORE_U8 ? SCGBM_U8 : (full.storage == SCGBM_STORE_U16 ? SCGBM_U16 : SCGBM_F32);
    load_counts_rows(ctx, full.data, dt, 1, a->I_full, J, full.ld, a->ixs, Ik, Y, rs_full.get());
    ctx.sync();
  }
  DevBuf<double> U(Ik * M), X(Ik * p), off(Ik), coef(J * p), se(J * p);
  DevBuf<int8_t> fail(J);
  DevBuf<int32_t> iters(J);
  SCGBM_CUDA(cudaMemcpyAsync(U.get(), a->U, Ik * M * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  SCGBM_CUDA(cudaMemcpyAsync(off.get(), a->alpha, Ik * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  build_design_kernel<<<(unsigned)ceil_div(Ik * p, 256), 256, 0, ctx.stream>>>(Ik, M, U.get(), X.get());
  SCGBM_LAUNCH_CHECK();
  ProjParams q;
  q.Ik = Ik; q.J = J; q.ldY = Y.ld; q.Y = Y.data; q.X = X.get(); q.off = off.get(); q.p = p;
  q.tol = a->glm_tol > 0 ? a->glm_tol : 1e-8;
  q.maxit = a->glm_maxit > 0 ? a->glm_maxit : 100;
  q.coef = coef.get(); q.se = se.get(); q.fail = fail.get(); q.iters = iters.get();
  q.ncell = J; q.list = nullptr;
  const size_t smem = (size_t)(WG_TR * (p + 1) + 2 * WG_TR * WG_CC + 3 * WG_CC * p + (size_t)WG_CC * p * p + 2 * WG_CC) * sizeof(double) +
                      2 * WG_CC * sizeof(int) + 2 * (size_t)T * sizeof(unsigned short) + 32;
  SCGBM_REQUIRE(smem <= 220 * 1024, "projection: M too large for shared memory");
  // Method (scgbm_proj_args.glm_method; SCGBM_PROJ_METHOD overrides for experiments): the tensor-core path runs all
  // cells in lock step through the fused pass's product rider (proj_tc.cu) and leaves only its stragglers to the
  // exact kernel; AUTO takes it for integer counts once there are enough cells to fill the machine.
  int method = a->glm_method;
  if (const char* e = getenv("SCGBM_PROJ_METHOD")) method = atoi(e);
  SCGBM_REQUIRE(method >= SCGBM_GLM_AUTO && method <= SCGBM_GLM_TCGEN05, "glm_method must be 0 (auto), 1 (exact) or 2 (tcgen05)");
  const bool tc_ok = proj_tc_supported(ctx, Y, M);
  SCGBM_REQUIRE(method != SCGBM_GLM_TCGEN05 || tc_ok, "glm_method = tcgen05 needs integer counts (u8 / u16 storage) and M <= 64");
  const bool use_tc = method == SCGBM_GLM_TCGEN05 || (method == SCGBM_GLM_AUTO && tc_ok && J >= 4096);
  DevBuf<int64_t> list;
  ProjTcStats tcs;
  if (use_tc) {
    q.ncell = project_tc(ctx, Y, Ik, M, U.get(), off.get(), q.tol, 30, coef.get(), se.get(), fail.get(), iters.get(), list, &tcs);
    q.list = list.get();
  }
  if (q.ncell > 0) {
    ProfScope ps(ctx.prof, SCGBM_K_PROJ, 1);
    const unsigned grid = (unsigned)ceil_div(q.ncell, WG_CC);
    switch (PT) {
      case 1: launch_proj<1>(ctx, q, Y.storage, grid, smem); break;
      case 2: launch_proj<2>(ctx, q, Y.storage, grid, smem); break;
      case 3: case 4: launch_proj<4>(ctx, q, Y.storage, grid, smem); break;
      case 5: case 6: launch_proj<6>(ctx, q, Y.storage, grid, smem); break;
      default: launch_proj<9>(ctx, q, Y.storage, grid, smem); break;
    }
  }
  out->method_used = use_tc ? SCGBM_GLM_TCGEN05 : SCGBM_GLM_EXACT;
  out->tc_sweeps = tcs.sweeps;
  out->tc_stragglers = use_tc ? q.ncell : 0;
  // V[, 1] = beta, V[, 2:(M+1)] = scores, se[-1] = se_scores   (R/fit.R:258-270)
  SCGBM_CUDA(cudaMemcpyAsync(out->beta, coef.get(), J * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  SCGBM_CUDA(cudaMemcpyAsync(out->scores, coef.get() + J, J * M * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  SCGBM_CUDA(cudaMemcpyAsync(out->se_scores, se.get() + J, J * M * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  SCGBM_CUDA(cudaMemcpyAsync(out->fail, fail.get(), J * sizeof(int8_t), cudaMemcpyDeviceToHost, ctx.stream));
  if (out->iters) SCGBM_CUDA(cudaMemcpyAsync(out->iters, iters.get(), J * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx.stream));
  if (out->rowsum_full) SCGBM_CUDA(cudaMemcpyAsync(out->rowsum_full, rs_full.get(), a->I_full * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  SCGBM_CUDA(cudaEventRecord(ev1, ctx.stream));
  ctx.sync();
  ctx.prof.d2h_bytes += (double)J * (2 * M + 1) * sizeof(double);
  float tot = 0;
  cudaEventElapsedTime(&tot, ev0, ev1);
  cudaEventDestroy(ev0); cudaEventDestroy(ev1);
  memset(&out->prof, 0, sizeof(out->prof));
  ctx.prof.fill(&out->prof);
  out->prof.total_ms = tot;
}

}  // namespace scgbm
