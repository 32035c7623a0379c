// fit.cu -- K7: host control of gbm.sc's full-data fit (R/fit.R:76-211) over device
// state.  X is never an I x J matrix: it is one of three factor slots; the only I x J
// device objects are Y (compact counts) and the residual R (two fp16 planes of the scaled residual, 4 B/entry).
#include <chrono>
#include <map>
#include "common.cuh"
#include "counts.cuh"
#include "eta_pass.cuh"
#include "tsvd.cuh"

namespace scgbm {

enum { XK_ZERO = 0, XK_LOWRANK = 1, XK_X0 = 2 };

// Ritz residual estimate at which the block-Krylov SVD stops.  irlba's own default is tol = 1e-5 (R/fit.R:115,323
// call it with defaults [ext]); the looser 1e-4 of round 1 left the spans 1.4e-4 rad from an exact (LAPACK) SVD at
// 20 000 genes x 2 000 cells (tests/golden/tall_*), outside the 1e-4 tolerance.
constexpr double kDefaultSvdTol = 1e-5;

struct Factor {
  int kind = XK_ZERO;
  int M = 0;
  DevBuf<double> U, d, V;      // ldI x M, M, ldJ x M   (orthonormal factors from the SVD)
  DevBuf<float> A32, B32;      // fp32 tiles for the eta passes: A32 = U d (* row scale), B32 = V (* col scale)
  Factor16 a16, b16;           // the same, as fp16 tensor-core operands (single batch only)
  bool has16 = false;
  void alloc(int64_t ldI, int64_t ldJ, int M_) {
    M = M_;
    U.alloc(ldI * M); d.alloc(M); V.alloc(ldJ * M); A32.alloc(ldI * M); B32.alloc(ldJ * M);
  }
};

// ---- small elementwise kernels ----------------------------------------------------
__global__ void log_kernel(int64_t n, const double* __restrict__ x, double* __restrict__ y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = log(x[i]);
}

// single CTA: alpha = logrs - log(S) on the I real rows of each batch column, then
// recentre:  shift = mean(alpha); alpha -= shift   (R/fit.R:127-132 and :106-110)
__global__ void __launch_bounds__(1024)
alpha_update_kernel(int64_t I, int64_t ldI, int nbatch, const double* __restrict__ logrs,
                    const double* __restrict__ S, const double* __restrict__ logtot, double* __restrict__ alpha,
                    double* __restrict__ shift) {
  __shared__ double red[1024];
  double s = 0.0;
  for (int64_t idx = threadIdx.x; idx < I * nbatch; idx += 1024) {
    const int64_t i = idx % I;
    const int k = (int)(idx / I);
    const double a = logrs[k * ldI + i] - (S ? log(S[k * ldI + i]) : logtot[k]);
    alpha[k * ldI + i] = a;
    s += a;
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int st = 512; st > 0; st >>= 1) {
    if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
    __syncthreads();
  }
  const double mean = red[0] / (double)(I * nbatch);
  for (int64_t idx = threadIdx.x; idx < I * nbatch; idx += 1024) {
    const int64_t i = idx % I;
    const int k = (int)(idx / I);
    alpha[k * ldI + i] -= mean;
  }
  if (threadIdx.x == 0) shift[0] = mean;
}

__global__ void beta_shift_kernel(int64_t J, double* __restrict__ beta, const double* __restrict__ shift) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < J) beta[j] += shift[0];
}

// beta_j = logcs_j - log(C_j)    (R/fit.R:134)
__global__ void beta_infer_kernel(int64_t J, const double* __restrict__ logcs, const double* __restrict__ C,
                                  double* __restrict__ beta) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < J) beta[j] = logcs[j] - log(C[j]);
}

// fp32 staging: e[i] = exp(scale * x[i]) and l[i] = x[i]; rows >= n_real get 0
__global__ void exp_stage_kernel(int64_t n_real, int64_t ld, int ncol, const double* __restrict__ x, double scale,
                                 float* __restrict__ e, float* __restrict__ l) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ld * ncol) return;
  const int64_t r = idx % ld;
  const bool ok = r < n_real;
  const double v = ok ? x[idx] : 0.0;
  if (e) e[idx] = ok ? (float)exp(scale * v) : 0.0f;
  if (l) l[idx] = (float)v;
}

// A32[i,m] = U[i,m] * d[m] * (rs ? rs[i] : 1);  rows >= n_real are zero
__global__ void stage_factor_kernel(int64_t n_real, int64_t ld, int M, const double* __restrict__ U,
                                    const double* __restrict__ d, const float* __restrict__ rs,
                                    float* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ld * M) return;
  const int64_t r = idx % ld;
  const int m = (int)(idx / ld);
  double v = 0.0;
  if (r < n_real) v = U[idx] * (d ? d[m] : 1.0) * (rs ? (double)rs[r] : 1.0);
  out[idx] = (float)v;
}

// Z = (Y - W0)/sqrt(W0) with W0 = a_i b_j:  z = y * s_i t_j - 1/(s_i t_j)    (R/fit.R:111-114)
// MAXONLY: only max |z| (as float bits, non-negative floats order like ints) -> zmax;
// otherwise the fp16 planes of scale * z.
template <typename YT, bool MAXONLY>
__global__ void __launch_bounds__(256)
init_z_kernel(int64_t I, int64_t ldI, int64_t J, const YT* __restrict__ Y,
              const float* __restrict__ s_row /* ldI x nbatch */, const float* __restrict__ t_col,
              const int32_t* __restrict__ batch, float scale, uint16_t* __restrict__ Zhi,
              uint16_t* __restrict__ Zlo, int* __restrict__ zmax) {
  // one column (cell) per block trip, 8 consecutive genes per thread trip (ldI is a multiple of 64)
  float m = 0.0f;
  for (int64_t j = blockIdx.x; j < J; j += gridDim.x) {
    const int kb = batch ? batch[j] : 0;
    const float tj = t_col[j];
    const YT* yc = Y + j * ldI;
    const float* sr = s_row + (int64_t)kb * ldI;
    for (int64_t i0 = (int64_t)threadIdx.x * 8; i0 < ldI; i0 += (int64_t)blockDim.x * 8) {
      float y[8];
      if (sizeof(YT) == 1) {
        const uint2 v = *reinterpret_cast<const uint2*>(yc + i0);
        const uint32_t w[2] = {v.x, v.y};
#pragma unroll
        for (int k = 0; k < 8; ++k) y[k] = (float)((w[k >> 2] >> (8 * (k & 3))) & 0xffu);
      } else if (sizeof(YT) == 2) {
        const uint4 v = *reinterpret_cast<const uint4*>(yc + i0);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) y[k] = (float)((w[k >> 1] >> (16 * (k & 1))) & 0xffffu);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) y[k] = (float)yc[i0 + k];
      }
      const float4 s0 = *reinterpret_cast<const float4*>(sr + i0), s1 = *reinterpret_cast<const float4*>(sr + i0 + 4);
      const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
      uint32_t hp[4], lp[4];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float z = 0.0f;
        if (i0 + k < I) {
          const float st = sv[k] * tj;
          z = y[k] * st - 1.0f / st;
        }
        if (MAXONLY) {
          m = fmaxf(m, fabsf(z));
        } else {
          uint16_t h, l;
          split_f16x2(z * scale, h, l);
          if (k & 1) { hp[k >> 1] |= (uint32_t)h << 16; lp[k >> 1] |= (uint32_t)l << 16; }
          else { hp[k >> 1] = h; lp[k >> 1] = l; }
        }
      }
      if (!MAXONLY) {
        *reinterpret_cast<uint4*>(Zhi + j * ldI + i0) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
        *reinterpret_cast<uint4*>(Zlo + j * ldI + i0) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
      }
    }
  }
  if (MAXONLY) {   // one atomic per block (a per-warp atomic on a single address serialised the whole pass)
    __shared__ float red[8];
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
      if (m > 0.0f) atomicMax(zmax, __float_as_int(m));
    }
  }
}

// W[i, j] = exp(alpha_i,b(j) + beta_j + X_ij) in fp64 (R/fit.R:189-191), column chunk
__global__ void write_w_kernel(int64_t I, int64_t ldI, int64_t j0, int64_t jc, int64_t ldJ, int M, int kind,
                               const double* __restrict__ U, const double* __restrict__ d,
                               const double* __restrict__ V, const double* __restrict__ alpha,
                               const double* __restrict__ beta, const int32_t* __restrict__ batch,
                               const double* __restrict__ alpha0, const double* __restrict__ beta0,
                               double* __restrict__ W) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= I * jc) return;
  const int64_t i = idx % I, jl = idx / I, j = j0 + jl;
  const int kb = batch ? batch[j] : 0;
  double x = 0.0;
  if (kind != XK_ZERO) {
    for (int m = 0; m < M; ++m) x += U[(int64_t)m * ldI + i] * d[m] * V[(int64_t)m * ldJ + j];
    if (kind == XK_X0) {
      x *= exp(-0.5 * (alpha0[(int64_t)kb * ldI + i] + beta0[j]));
      x = x > 8.0 ? 8.0 : x;
      x = x < -8.0 ? -8.0 : x;
    }
  }
  W[jl * I + i] = exp(alpha[(int64_t)kb * ldI + i] + x) * exp(beta[j]);
}

__global__ void planes_to_f64_kernel(int64_t n, const uint16_t* __restrict__ hi, const uint16_t* __restrict__ lo,
                                     double inv_scale, double* __restrict__ b) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) b[i] = (double)join_f16x2(hi[i], lo[i]) * inv_scale;
}

__global__ void colsum_batch_total_kernel(int64_t I, int64_t ldI, int nbatch, const double* __restrict__ rowsum,
                                          double* __restrict__ logtot) {
  // logtot[k] = log(sum_i rowsum[i,k])  == log(sum(exp(betas[batch==k])))  (R/fit.R:107)
  __shared__ double red[1024];
  for (int k = 0; k < nbatch; ++k) {
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < I; i += 1024) s += rowsum[k * ldI + i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int st = 512; st > 0; st >>= 1) {
      if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
      __syncthreads();
    }
    if (threadIdx.x == 0) logtot[k] = log(red[0]);
    __syncthreads();
  }
}

__global__ void min_kernel(int64_t n, const double* __restrict__ x, int64_t n_real, int64_t ld, double* __restrict__ out) {
  // out[0] = min over real entries (single CTA)
  __shared__ double red[1024];
  double m = 1e300;
  for (int64_t idx = threadIdx.x; idx < n; idx += 1024) {
    if (idx % ld < n_real) m = fmin(m, x[idx]);
  }
  red[threadIdx.x] = m;
  __syncthreads();
  for (int st = 512; st > 0; st >>= 1) {
    if (threadIdx.x < st) red[threadIdx.x] = fmin(red[threadIdx.x], red[threadIdx.x + st]);
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = red[0];
}

#define GRID1(n) (unsigned)ceil_div((n), 256), 256, 0, ctx.stream

// ---------------------------------------------------------------------------------
struct FitState {
  Ctx& ctx;
  const scgbm_fit_args& a;
  Counts Y;
  int64_t I, J, ldI, ldJ, J_total;
  int nbatch, M;
  DevBuf<uint16_t> Rhi, Rlo;          // residual planes (products.cuh)
  float rscale = 1.0f;                // power-of-two scale of what the planes currently hold
  ResidBuf resid() const { ResidBuf r; r.hi = Rhi.get(); r.lo = Rlo.get(); r.scale = rscale; return r; }
  DevBuf<double> alpha, beta, alpha0, beta0, logrs, logcs, S, Ccol, logtot, shift, scal, rs_local;
  DevBuf<float> f_row, la_row, g_col, lb_col, s_row, t_col;
  Factor slot[3];
  EtaWs eta;
  TsvdWs svd;
  ProdWs fuse_ws;           // fp16 planes of the SVD's start block for the fused pass's product rider
  DevBuf<double> RtQ0;      // ldJ x b: R^T Q0 delivered by the fused pass
  double maxY = 0;
  PassGeom geom;

  FitState(Ctx& c, const scgbm_fit_args& args) : ctx(c), a(args) {}

  XDesc desc(const Factor& f) const {
    XDesc x;
    if (f.kind == XK_ZERO) { x.M = 0; return x; }
    x.A = f.A32.get(); x.B = f.B32.get(); x.M = f.M;
    if (f.has16) { x.A16 = &f.a16; x.B16 = &f.b16; }
    if (f.kind == XK_LOWRANK) { x.U64 = f.U.get(); x.d64 = f.d.get(); x.V64 = f.V.get(); }
    if (f.kind == XK_X0) {
      x.clamp = 1;
      x.rowscale = nbatch > 1 ? s_row.get() : nullptr;
    }
    return x;
  }

  void stage_cols() {   // g_col = exp(beta), lb_col = beta
    ProfScope ps(ctx.prof, SCGBM_K_MISC, 1);
    exp_stage_kernel<<<GRID1(ldJ)>>>(J, ldJ, 1, beta.get(), 1.0, g_col.get(), lb_col.get());
    SCGBM_LAUNCH_CHECK();
  }
  void stage_rows() {   // f_row = exp(alpha), la_row = alpha
    ProfScope ps(ctx.prof, SCGBM_K_MISC, 1);
    exp_stage_kernel<<<GRID1(ldI * nbatch)>>>(I, ldI, nbatch, alpha.get(), 1.0, f_row.get(), la_row.get());
    SCGBM_LAUNCH_CHECK();
  }
  void stage_factor(Factor& f) {
    ProfScope ps(ctx.prof, SCGBM_K_MISC, 2);
    const float* rs = (f.kind == XK_X0 && nbatch == 1) ? s_row.get() : nullptr;
    const float* cs = (f.kind == XK_X0) ? t_col.get() : nullptr;
    stage_factor_kernel<<<GRID1(ldI * f.M)>>>(I, ldI, f.M, f.U.get(), f.d.get(), rs, f.A32.get());
    stage_factor_kernel<<<GRID1(ldJ * f.M)>>>(J, ldJ, f.M, f.V.get(), nullptr, cs, f.B32.get());
    SCGBM_LAUNCH_CHECK();
    f.has16 = false;
    if (f.M <= 64) {
      stage_factor16(ctx, f.a16, I, ldI, f.M, 0, f.U.get(), f.d.get(), rs);
      stage_factor16(ctx, f.b16, J, ldJ, f.M, 1, f.V.get(), nullptr, cs);
      f.has16 = true;
    }
  }

  void setup() {
    I = a.I; J = a.J; M = a.M; nbatch = a.batch ? std::max(1, a.nbatch) : 1;
    // Ingest.  With several ranks every rank must leave this block the same way: a rank that rejected its shard
    // (NA, negative, bad batch code, out of memory) while the others went on into the first collective would
    // leave them blocked in ncclAllReduce for ever, and a rank whose shard alone forced the u16 -> f32 storage
    // fall-back would run different kernels from its peers.  So the status and the storage format are agreed
    // (max over ranks) before anything else happens.
    auto ingest = [&](int storage_req) {
      if (!a.Y && a.csc_p) load_counts_csc(ctx, a.csc_i, a.csc_p, a.csc_x, a.csc_x_dtype, I, J, a.batch, nbatch, storage_req, Y);
      else load_counts(ctx, a.Y, a.y_dtype, a.y_on_device, I, J, a.ldY, a.batch, nbatch, storage_req, Y);
    };
    auto agreed_ingest = [&](int storage_req) -> int {      // returns the largest storage code any rank ended with
      int err = 0;
      std::string msg;
      try { ingest(storage_req); } catch (const Error& e) { err = e.status; msg = e.what(); }
      if (ctx.nranks == 1) {
        if (err) throw Error(err, msg);
        return Y.storage;
      }
      DevBuf<double> ag(2);
      double hv[2] = {(double)err, err ? 0.0 : (double)Y.storage};
      SCGBM_CUDA(cudaMemcpyAsync(ag.get(), hv, sizeof(hv), cudaMemcpyHostToDevice, ctx.stream));
      ctx.allreduce_max(ag.get(), 2);
      SCGBM_CUDA(cudaMemcpyAsync(hv, ag.get(), sizeof(hv), cudaMemcpyDeviceToHost, ctx.stream));
      ctx.sync();
      if (err) throw Error(err, msg);
      if (hv[0] != 0.0)
        throw Error((int)hv[0], "the count matrix was rejected on another rank (see that rank's message); "
                                "no rank proceeds");
      return (int)hv[1];
    };
    {
      const int agreed = agreed_ingest(a.y_storage);
      if (agreed != Y.storage) agreed_ingest(agreed);       // a peer needed the wider format: repack to match it
    }
    ldI = Y.ld; ldJ = round_up(J, kRowPad);
    geom = PassGeom{I, J, ldI, ldJ, nbatch, Y.batch.get()};
    // global stats
    scal.alloc(8);
    double h[2] = {(double)J, Y.maxY};
    rs_local.alloc(ldI * nbatch);     // this rank's row sums (the objective's constant term), before the all-reduce
    SCGBM_CUDA(cudaMemcpyAsync(rs_local.get(), Y.rowsum.get(), ldI * nbatch * sizeof(double), cudaMemcpyDeviceToDevice, ctx.stream));
    if (ctx.nranks > 1) {
      SCGBM_CUDA(cudaMemcpyAsync(scal.get(), h, sizeof(h), cudaMemcpyHostToDevice, ctx.stream));
      ctx.allreduce_sum(scal.get(), 1);
      ctx.allreduce_max(scal.get() + 1, 1);
      ctx.allreduce_sum(Y.rowsum.get(), ldI * nbatch);
      SCGBM_CUDA(cudaMemcpyAsync(h, scal.get(), sizeof(h), cudaMemcpyDeviceToHost, ctx.stream));
      ctx.sync();
    }
    J_total = (int64_t)llround(h[0]);
    maxY = h[1];
    // AUTO picks the narrowest format the counts fit: u16 first (the maximum is only known after the pack), then
    // u8 when the maximum over ALL ranks is <= 255 -- one byte per count less for the fused pass to read
    // (SCGBM_AUTO_U8=0 keeps u16: A/B measurements)
    static const bool auto_u8 = !(getenv("SCGBM_AUTO_U8") && getenv("SCGBM_AUTO_U8")[0] == '0');
    if (auto_u8 && a.y_storage == SCGBM_STORE_AUTO && Y.storage == SCGBM_STORE_U16 && maxY <= 255.0) {
      Y.maxY = std::min(Y.maxY, 255.0);
      narrow_counts_to_u8(ctx, Y);
    }
    SCGBM_REQUIRE(M >= 1, "M must be >= 1");
    SCGBM_REQUIRE(M <= std::min(I, J_total), "M must not exceed min(nrow(Y), ncol(Y))");
    SCGBM_REQUIRE(maxY > 0.0, "Count matrix Y is all zero");

    Rhi.alloc(ldI * J); Rlo.alloc(ldI * J);
    alpha.alloc(ldI * nbatch); alpha0.alloc(ldI * nbatch); logrs.alloc(ldI * nbatch); S.alloc(ldI * nbatch);
    beta.alloc(ldJ); beta0.alloc(ldJ); logcs.alloc(ldJ); Ccol.alloc(ldJ);
    logtot.alloc(nbatch); shift.alloc(1);
    f_row.alloc(ldI * nbatch + 128); la_row.alloc(ldI * nbatch); s_row.alloc(ldI * nbatch);   // f_row: zero tail (tiles)
    f_row.zero(ctx.stream);
    g_col.alloc(ldJ + 128); lb_col.alloc(ldJ); t_col.alloc(ldJ);   // g_col: zero tail for whole 128-cell tiles
    g_col.zero(ctx.stream);
    alpha.zero(ctx.stream); beta.zero(ctx.stream); logcs.zero(ctx.stream);
    for (auto& f : slot) f.alloc(ldI, ldJ, M);

    // Q11 precondition: every gene (per batch) and every cell must have a count
    {
      DevBuf<double> mn(2);
      min_kernel<<<1, 1024, 0, ctx.stream>>>(ldI * nbatch, Y.rowsum.get(), I, ldI, mn.get());
      min_kernel<<<1, 1024, 0, ctx.stream>>>(J, Y.colsum.get(), J, J, mn.get() + 1);
      SCGBM_LAUNCH_CHECK();
      double hm[2];
      SCGBM_CUDA(cudaMemcpyAsync(hm, mn.get(), sizeof(hm), cudaMemcpyDeviceToHost, ctx.stream));
      ctx.sync();
      double bad = (!(hm[0] > 0.0) || !(hm[1] > 0.0)) ? 1.0 : 0.0;
      if (ctx.nranks > 1) {   // every rank must take the same branch
        SCGBM_CUDA(cudaMemcpyAsync(mn.get(), &bad, sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
        ctx.allreduce_max(mn.get(), 1);
        SCGBM_CUDA(cudaMemcpyAsync(&bad, mn.get(), sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
        ctx.sync();
      }
      if (bad != 0.0)
        throw Error(SCGBM_ERR_INVALID,
                    "every gene (within each batch) and every cell needs at least one count: the reference "
                    "takes log(rowSums)/log(colSums) (R/fit.R:96,104) and turns NaN otherwise");
    }
    ProfScope ps(ctx.prof, SCGBM_K_MISC, 6);
    // betas = log(colSums(Y))  (R/fit.R:96,99); log.rsy (R/fit.R:102-105)
    log_kernel<<<GRID1(J)>>>(J, Y.colsum.get(), logcs.get());
    SCGBM_CUDA(cudaMemcpyAsync(beta.get(), logcs.get(), J * sizeof(double), cudaMemcpyDeviceToDevice, ctx.stream));
    log_kernel<<<GRID1(ldI * nbatch)>>>(ldI * nbatch, Y.rowsum.get(), logrs.get());
    // alphas = log.rsy - log(sum(exp(betas[batch==k])))  (R/fit.R:106-108); recentre (:109-110)
    colsum_batch_total_kernel<<<1, 1024, 0, ctx.stream>>>(I, ldI, nbatch, Y.rowsum.get(), logtot.get());
    alpha_update_kernel<<<1, 1024, 0, ctx.stream>>>(I, ldI, nbatch, logrs.get(), nullptr, logtot.get(), alpha.get(), shift.get());
    beta_shift_kernel<<<GRID1(J)>>>(J, beta.get(), shift.get());
    SCGBM_LAUNCH_CHECK();
    SCGBM_CUDA(cudaMemcpyAsync(alpha0.get(), alpha.get(), ldI * nbatch * sizeof(double), cudaMemcpyDeviceToDevice, ctx.stream));
    SCGBM_CUDA(cudaMemcpyAsync(beta0.get(), beta.get(), ldJ * sizeof(double), cudaMemcpyDeviceToDevice, ctx.stream));
    // s = exp(-alpha0/2), t = exp(-beta0/2): W0^(-1/2) = s_i t_j   (R/fit.R:111,114,117)
    exp_stage_kernel<<<GRID1(ldI * nbatch)>>>(I, ldI, nbatch, alpha.get(), -0.5, s_row.get(), nullptr);
    exp_stage_kernel<<<GRID1(ldJ)>>>(J, ldJ, 1, beta.get(), -0.5, t_col.get(), nullptr);
    SCGBM_LAUNCH_CHECK();
  }

  template <bool MAXONLY>
  void launch_init_z(float scale, int* zmax) {
    const unsigned grid = (unsigned)std::min<int64_t>(J, (int64_t)ctx.num_sms * 16);
    const int32_t* bt = Y.batch.get();
    switch (Y.storage) {
      case SCGBM_STORE_U8: init_z_kernel<uint8_t, MAXONLY><<<grid, 256, 0, ctx.stream>>>(I, ldI, J, (const uint8_t*)Y.data, s_row.get(), t_col.get(), bt, scale, Rhi.get(), Rlo.get(), zmax); break;
      case SCGBM_STORE_U16: init_z_kernel<uint16_t, MAXONLY><<<grid, 256, 0, ctx.stream>>>(I, ldI, J, (const uint16_t*)Y.data, s_row.get(), t_col.get(), bt, scale, Rhi.get(), Rlo.get(), zmax); break;
      default: init_z_kernel<float, MAXONLY><<<grid, 256, 0, ctx.stream>>>(I, ldI, J, (const float*)Y.data, s_row.get(), t_col.get(), bt, scale, Rhi.get(), Rlo.get(), zmax); break;
    }
    SCGBM_LAUNCH_CHECK();
  }

  void init_svd() {
    {
      ProfScope ps(ctx.prof, SCGBM_K_MISC, 2);
      // the planes are fp16: find max |Z| first so that the stored values stay in range
      DevBuf<int> zmax(1);
      zmax.zero(ctx.stream);
      launch_init_z<true>(1.0f, zmax.get());
      int bits = 0;
      SCGBM_CUDA(cudaMemcpyAsync(&bits, zmax.get(), sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
      ctx.sync();
      float zm;
      memcpy(&zm, &bits, sizeof(float));
      double h = zm;
      if (ctx.nranks > 1) {          // every rank must use the same scale (the products are summed)
        SCGBM_CUDA(cudaMemcpyAsync(scal.get() + 4, &h, sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
        ctx.allreduce_max(scal.get() + 4, 1);
        SCGBM_CUDA(cudaMemcpyAsync(&h, scal.get() + 4, sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
        ctx.sync();
      }
      rscale = resid_scale_for(h);
      launch_init_z<false>(rscale, nullptr);
    }
    GOperator op;
    op.I = I; op.ldI = ldI; op.J = J; op.ldJ = ldJ; op.J_total = J_total;
    op.R = resid(); op.c = 1.0; op.nterms = 0;
    TsvdOpts o = opts();
    Factor& f = slot[0];
    tsvd(ctx, svd, op, o, f.U.get(), f.d.get(), f.V.get(), nullptr, nullptr);     // R/fit.R:115
    f.kind = XK_X0;
    stage_factor(f);                                                                // R/fit.R:116-120
    // the Pearson-residual right basis stays as the warm start of the first PGD projection
  }

  TsvdOpts opts() const {
    TsvdOpts o;
    o.M = M;
    o.extra = a.svd_extra > 0 ? a.svd_extra : 12;
    o.depth = a.svd_depth > 0 ? a.svd_depth : 3;
    static const double env_tol = getenv("SCGBM_SVD_TOL") ? atof(getenv("SCGBM_SVD_TOL")) : 0.0;   // experiments
    o.tol = a.svd_tol > 0 ? a.svd_tol : (env_tol > 0 ? env_tol : kDefaultSvdTol);
    // The warm calls of the loop stop at the same estimate as the initial SVD (irlba's own 1e-5).  A decade lower
    // (SCGBM_SVD_TOL_WARM=1e-6) was measured: steady-state calls sit at 1e-8 .. 3e-7 after their two mandatory
    // products and do not notice, but the late iterations of a converging fit do -- the complete 100-iteration fit at
    // 20k x 1M took 11.9 s instead of 11.1 s (3e-6: 11.9 s as well), for final objectives 3e-10 apart.  What it would
    // buy: the objective of a REJECTED overshoot step right after a revert, a value the reference never returns,
    // moved from 6e-6 to within 1e-6 of the exact-SVD float64 reference fit on 4 and 8 ranks (tests/test_gpu_multi.py holds
    // rejected iterations to 1e-5 and accepted ones to 1e-6).
    static const double env_tolw = getenv("SCGBM_SVD_TOL_WARM") ? atof(getenv("SCGBM_SVD_TOL_WARM")) : 0.0;   // experiments
    o.tol_warm = env_tolw > 0 ? env_tolw : o.tol;
    o.seed = a.seed;
    return o;
  }
};

static double now_sec() {
  using namespace std::chrono;
  return duration<double>(steady_clock::now().time_since_epoch()).count();
}

// One fit in flight: scgbm_fit_begin / _step / _end drive it; scgbm_fit = all three.
struct FitRun {
  scgbm_fit_args a;
  Ctx ctx;
  FitState st;
  std::vector<double> LL, obj, tvec;
  std::vector<int8_t> acc;
  double lr = 1.0;                                       // R/fit.R:84
  int cur = 0, prev = -1, last = -1;                     // factor slots; prev == -1: Xt = 0 (R/fit.R:123)
  int i = 0;                                             // loop index of the last trip
  int n_iter = 0, hit_max = 0;
  bool finished = false;
  double t_prev = 0, t_cum = 0, fused_bytes = 0;
  int n_underflow_reruns = 0;
  std::map<int, double> dbg_ll;                          // SCGBM_DBG_LL="7:-inf,9:nan": objective overrides (tests only)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;

  explicit FitRun(const scgbm_fit_args& args)
      : a(args), ctx(args.device, reinterpret_cast<Comm*>(args.comm)), st(ctx, a) {
    ctx.prof.timing = a.profile != 0;
    ctx.prof.fused_only = a.profile == 2;
    if (const char* e = getenv("SCGBM_DBG_LL")) {
      std::string sv(e);
      size_t pos = 0;
      while (pos < sv.size()) {
        size_t end = sv.find(',', pos);
        if (end == std::string::npos) end = sv.size();
        const std::string tok = sv.substr(pos, end - pos);
        const size_t c = tok.find(':');
        if (c != std::string::npos) {
          const std::string v = tok.substr(c + 1);
          dbg_ll[atoi(tok.substr(0, c).c_str())] = v == "nan" ? NAN : (v == "inf" ? INFINITY : (v == "-inf" ? -INFINITY : atof(v.c_str())));
        }
        pos = end + 1;
      }
    }
    SCGBM_CUDA(cudaEventCreate(&ev0));
    SCGBM_CUDA(cudaEventCreate(&ev1));
    SCGBM_CUDA(cudaEventRecord(ev0, ctx.stream));
    st.setup();
    st.init_svd();
    LL.assign(a.max_iter, 0.0);
    acc.assign(a.max_iter, 0);
    tvec.assign(a.max_iter, 0.0);
    t_prev = now_sec();
    ctx.sync();
  }
  ~FitRun() {
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
  }

  void revert() {                                        // X <- Xt
    if (prev >= 0) { cur = prev; return; }
    int z = 0;                                           // Xt is still the zero matrix (R/fit.R:123)
    while (z == cur) ++z;
    st.slot[z].kind = XK_ZERO; st.slot[z].M = st.M;
    cur = z;
  }

  // one trip through R/fit.R:125-186; returns false when the loop has ended
  bool step() {
    if (finished) return false;
    if (i >= a.max_iter) { finished = true; return false; }
    ++i;
    n_iter = i;
    const int64_t I = st.I, J = st.J, ldI = st.ldI, ldJ = st.ldJ;
    const int M = st.M, nbatch = st.nbatch;
    const Factor& fc = st.slot[cur];
    const XDesc xc = st.desc(fc);
    static const bool trace = getenv("SCGBM_TRACE") != nullptr;
    const double tr0 = trace ? now_sec() : 0.0;
    // (a) alphas <- log.rsy - log(rowSums(exp(X + betas)))      R/fit.R:127-130
    st.stage_cols();
    const bool tc_eta = eta_tc_supported(st.geom, xc, st.Y.storage);
    if (tc_eta) pass_rowsum_tc(ctx, st.eta, st.geom, xc, st.g_col.get(), st.S.get(), st.beta.get());
    else pass_rowsum(ctx, st.eta, st.geom, xc, st.g_col.get(), st.S.get());
    ctx.allreduce_sum(st.S.get(), ldI * nbatch);
    {
      ProfScope ps(ctx.prof, SCGBM_K_MISC, 2);
      alpha_update_kernel<<<1, 1024, 0, ctx.stream>>>(I, ldI, nbatch, st.logrs.get(), st.S.get(), nullptr,
                                                      st.alpha.get(), st.shift.get());
      beta_shift_kernel<<<GRID1(J)>>>(J, st.beta.get(), st.shift.get());          // R/fit.R:131-132
      SCGBM_LAUNCH_CHECK();
    }
    ctx.rank_spread("alpha", i, st.alpha.get(), ldI * nbatch);
    st.stage_rows();
    if (a.infer_beta) {                                  // R/fit.R:133-135
      if (tc_eta && nbatch == 1) pass_colsum_tc(ctx, st.eta, st.geom, xc, st.f_row.get(), st.Ccol.get());
      else pass_colsum(ctx, st.eta, st.geom, xc, st.f_row.get(), st.Ccol.get());
      ProfScope ps(ctx.prof, SCGBM_K_MISC, 1);
      beta_infer_kernel<<<GRID1(J)>>>(J, st.logcs.get(), st.Ccol.get(), st.beta.get());
      SCGBM_LAUNCH_CHECK();
    }
    st.stage_cols();
    // (d)-(g) W, clip, objective, max W; residual Y - W           R/fit.R:136-151,321
    {
      // |Y - W| <= max.Y (W is clipped); in iterations 1-2 the clamped X0 (|x| <= 8) is folded
      // into the planes with weight <= w.max / lr afterwards, so leave room for it
      const bool x0 = fc.kind == XK_X0 || (prev >= 0 && prev != cur && st.slot[prev].kind == XK_X0);
      st.rscale = resid_scale_for(st.maxY * (x0 ? 1.0 + 10.0 / std::min(lr, 1.0) : 1.0));
    }
    // When this iteration's SVD will start from the previous left Ritz block Q0, the fused pass also
    // delivers R^T Q0 -- the SVD's first pass over the residual -- from the tiles it has in shared memory.
    bool rider = false;
    if (tc_eta && fc.kind == XK_LOWRANK && !(prev >= 0 && prev != cur && st.slot[prev].kind == XK_X0)) {
      int b = 0;
      const double* Q0 = tsvd_left_start(st.svd, I, st.J_total, st.opts(), &b);
      if (Q0 && b <= 32 && prod_uses_tc(b) && st.opts().variant == PROD_AUTO && fused_prod_supported(ctx, st.geom, xc)) {
        split_operand(ctx, st.fuse_ws, ldI, b, 32, ldI, Q0, ldI);
        if (st.RtQ0.n < ldJ * b) { st.RtQ0.alloc(ldJ * b); st.RtQ0.zero(ctx.stream); }
        FusedProd fp;
        fp.op16 = st.fuse_ws.op16.get(); fp.ldk = ldI; fp.op_inv = st.fuse_ws.opmeta.get() + 1; fp.b = b;
        fp.out = st.RtQ0.get(); fp.ldo = ldJ;
        pass_fused_tc(ctx, st.eta, st.geom, xc, st.Y, st.f_row.get(), st.g_col.get(), st.la_row.get(), st.lb_col.get(),
                      st.alpha.get(), st.beta.get(), st.rs_local.get(), (float)st.maxY, st.resid(), st.scal.get(), &fp);
        rider = true;
      }
    }
    if (rider) {
    } else if (tc_eta)
      pass_fused_tc(ctx, st.eta, st.geom, xc, st.Y, st.f_row.get(), st.g_col.get(), st.la_row.get(), st.lb_col.get(),
                    st.alpha.get(), st.beta.get(), st.rs_local.get(), (float)st.maxY, st.resid(), st.scal.get());
    else
      pass_fused(ctx, st.eta, st.geom, xc, st.Y, st.f_row.get(), st.g_col.get(), st.la_row.get(), st.lb_col.get(),
                 (float)st.maxY, st.resid(), st.scal.get());
    if (fused_bytes == 0.0) {
      XDesc xm = xc; xm.M = M;
      fused_bytes = fused_pass_bytes(st.geom, xm, st.Y.elem_bytes());
    }
    ctx.allreduce_sum(st.scal.get(), 2);
    // max W and the underflow flag of the tensor-core pass -- and the two sums once more: every accept / revert /
    // stop decision below is taken by each rank on its own copy of these scalars, so the copies must be IDENTICAL.
    // A sum all-reduce need not return the same last bit on every rank (the association order may differ between
    // ranks from 4 ranks on; with 2 it cannot, a + b = b + a); a max all-reduce does, whatever the order.
    ctx.allreduce_max(st.scal.get(), 4);
    double h[4];
    SCGBM_CUDA(cudaMemcpyAsync(h, st.scal.get(), sizeof(h), cudaMemcpyDeviceToHost, ctx.stream));
    ctx.sync();
    if (tc_eta && h[3] != 0.0) {
      // Some eta may lie below -745: the reference's exp() underflows to 0 there and log(0) = -Inf sends it
      // into the NaN/Inf branch (R/fit.R:152-157).  The tensor-core pass cannot see that (it sums Y * eta),
      // so this iteration is evaluated again by the CUDA-core pass, which restates the underflow per entry.
      pass_fused(ctx, st.eta, st.geom, xc, st.Y, st.f_row.get(), st.g_col.get(), st.la_row.get(), st.lb_col.get(),
                 (float)st.maxY, st.resid(), st.scal.get());
      rider = false;
      ++n_underflow_reruns;
      ctx.allreduce_sum(st.scal.get(), 2);
      ctx.allreduce_max(st.scal.get() + 2, 1);
      SCGBM_CUDA(cudaMemcpyAsync(h, st.scal.get(), 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
      ctx.sync();
    }
    const double tr1 = trace ? now_sec() : 0.0;
    {                                                    // R/fit.R:139-142
      const double t = now_sec();
      t_cum += t - t_prev; t_prev = t;
      tvec[i - 1] = t_cum;
    }
    double ll = h[1] - h[0];                             // R/fit.R:151
    if (!dbg_ll.empty()) {                               // test hook: pins the control flow of R/fit.R:152-157 (Q1, Q2)
      auto it = dbg_ll.find(i);
      if (it != dbg_ll.end()) ll = it->second;
    }
    const double w_max = h[2];
    LL[i - 1] = ll;
    acc[i - 1] = 0;
    if (a.verbose) fprintf(stderr, "[scgbm] iter %d LL=%.10g lr=%.4g wmax=%.6g\n", i, ll, lr, w_max);

    if (std::isnan(ll) || std::isinf(ll)) {              // R/fit.R:152-157
      revert();
      lr = lr / 2;
      if (i >= a.max_iter) finished = true;
      return !finished;
    }
    if (i >= 3) {                                        // R/fit.R:158
      const double tau = std::fabs((ll - LL[i - 3]) / ll);
      if (tau < a.tol && lr <= 1.06 && i >= a.min_iter) { finished = true; return false; }   // R/fit.R:160-162
      const double llp = LL[i - 2];
      if (std::isnan(llp))                               // quirk Q1: if (NA) errors in R
        throw Error(SCGBM_ERR_NUMERIC, "missing value where TRUE/FALSE needed (objective was NaN at the previous iteration)");
      if (ll < llp) {                                    // R/fit.R:164-167
        lr = std::max(lr / 2, 1.0);
        revert();
        if (i >= a.max_iter) finished = true;
        return !finished;
      } else {
        lr = lr * 1.05;                                  // R/fit.R:169
      }
    }
    obj.push_back(ll);                                   // R/fit.R:174
    acc[i - 1] = 1;
    if (a.progress) a.progress(i, ll, a.progress_user);  // R/fit.R:175

    // (k) pgd_irlba                                               R/fit.R:315-328
    const double m = (double)(i - 1) / (double)(i + 2);  // R/fit.R:319
    const double c = lr / w_max;                         // R/fit.R:321,323
    GOperator op;
    op.I = I; op.ldI = ldI; op.J = J; op.ldJ = ldJ; op.J_total = st.J_total;
    op.R = st.resid(); op.c = c; op.nterms = 0;
    op.RtQ0 = rider ? st.RtQ0.get() : nullptr;
    auto add_term = [&](int s, double coef) {
      if (s < 0 || coef == 0.0) return;
      Factor& f = st.slot[s];
      if (f.kind == XK_ZERO) return;
      if (f.kind == XK_X0) {                             // clamped X0 is not low-rank: fold it into R
        pass_addx(ctx, st.geom, st.desc(f), (float)(coef / c), st.resid());
        return;
      }
      LowRankTerm& t = op.t[op.nterms++];
      t.A = f.U.get(); t.lda = ldI; t.d = f.d.get(); t.B = f.V.get(); t.ldb = ldJ; t.M = f.M; t.coef = coef;
    };
    if (prev == cur || m == 0.0) add_term(cur, 1.0);     // X == Xt right after a revert (Q4)
    else { add_term(cur, 1.0 + m); add_term(prev, -m); }
    int next = 0;
    while (next == cur || next == prev) ++next;
    Factor& fn = st.slot[next];
    fn.kind = XK_LOWRANK; fn.M = M;
    TsvdOpts o = st.opts();
    int svd_passes = 0;
    double svd_est = 0.0;
    tsvd(ctx, st.svd, op, o, fn.U.get(), fn.d.get(), fn.V.get(), &svd_passes, &svd_est);   // R/fit.R:323
    st.stage_factor(fn);                                 // X <- u d v^T (never materialised; R/fit.R:324)
    if (trace) {
      ctx.sync();
      const double tr2 = now_sec();
      unsigned long long hc = 0;
      if (st.eta.exact_rows.n >= 1) SCGBM_CUDA(cudaMemcpy(&hc, st.eta.exact_rows.get(), sizeof(hc), cudaMemcpyDeviceToHost));
      fprintf(stderr, "[scgbm-trace] iter %d: eta passes %.2f ms, svd %.2f ms (%d passes, est %.2e), exact rows so far %llu\n", i,
              (tr1 - tr0) * 1e3, (tr2 - tr1) * 1e3, svd_passes, svd_est, hc);
    }
    prev = cur; cur = next; last = next;                 // R/fit.R:179-180,320
    if (i == a.max_iter) { hit_max = 1; finished = true; return false; }   // R/fit.R:182-185
    return true;
  }

  void fill_profile(scgbm_profile* p) {
    SCGBM_CUDA(cudaEventRecord(ev1, ctx.stream));
    ctx.sync();
    float tot = 0;
    cudaEventElapsedTime(&tot, ev0, ev1);
    memset(p, 0, sizeof(*p));
    ctx.prof.fill(p);
    p->total_ms = tot;
    p->fused_bytes = fused_bytes;
    p->svd_passes = st.svd.passes_total;
    p->svd_calls = st.svd.calls;
  }

  // ---- assemble (R/fit.R:188-204, 301-313) -------------------------------------
  void finish(scgbm_fit_out* out) {
    SCGBM_REQUIRE(out && out->U && out->D && out->V && out->alpha && out->beta && out->obj, "output buffers missing");
    if (last < 0)
      throw Error(SCGBM_ERR_NUMERIC, "no projected-gradient step was taken (object 'pgd' not found in the reference)");
    const int64_t I = st.I, J = st.J, ldI = st.ldI, ldJ = st.ldJ;
    const int M = st.M, nbatch = st.nbatch;
    out->n_obj = (int)obj.size(); out->n_iter = n_iter; out->hit_max_iter = hit_max;
    out->y_storage = st.Y.storage; out->max_Y = st.maxY;
    out->svd_unconverged = st.svd.unconverged; out->svd_worst_est = st.svd.worst_est;
    out->underflow_reruns = n_underflow_reruns;
    out->exact_rows = 0;
    if (st.eta.exact_rows.n >= 1) {
      unsigned long long hc = 0;
      SCGBM_CUDA(cudaMemcpyAsync(&hc, st.eta.exact_rows.get(), sizeof(hc), cudaMemcpyDeviceToHost, ctx.stream));
      ctx.sync();
      out->exact_rows = (int64_t)hc;
    }
    for (size_t k = 0; k < obj.size(); ++k) out->obj[k] = obj[k];
    for (int k = 0; k < n_iter; ++k) {
      if (out->LL) out->LL[k] = LL[k];
      if (out->accepted) out->accepted[k] = acc[k];
      if (out->time && a.time_by_iter) out->time[k] = tvec[k];
    }
    const Factor& fl = st.slot[last];
    std::vector<double> hU((size_t)ldI * M), hV((size_t)ldJ * M), hD(M), hA((size_t)ldI * nbatch), hB(ldJ);
    SCGBM_CUDA(cudaMemcpyAsync(hU.data(), fl.U.get(), hU.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
    SCGBM_CUDA(cudaMemcpyAsync(hV.data(), fl.V.get(), hV.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
    SCGBM_CUDA(cudaMemcpyAsync(hD.data(), fl.d.get(), M * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
    SCGBM_CUDA(cudaMemcpyAsync(hA.data(), st.alpha.get(), hA.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
    SCGBM_CUDA(cudaMemcpyAsync(hB.data(), st.beta.get(), hB.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
    ctx.sync();
    ctx.prof.d2h_bytes += (double)(hU.size() + hV.size() + hA.size() + hB.size()) * sizeof(double);
    // sign convention: gene 1 decides (U is replicated, so every rank flips alike)
    for (int m2 = 0; m2 < M; ++m2) {
      const bool flip = hU[(size_t)m2 * ldI + 0] < 0;                       // R/fit.R:305
      for (int64_t i2 = 0; i2 < I; ++i2) {
        const double u = hU[(size_t)m2 * ldI + i2];
        if (out->loadings) out->loadings[(size_t)m2 * I + i2] = u;          // pre-flip (Q7, R/fit.R:196)
        out->U[(size_t)m2 * I + i2] = flip ? -u : u;
      }
      for (int64_t j2 = 0; j2 < J; ++j2) {
        const double v = hV[(size_t)m2 * ldJ + j2];
        const double vf = flip ? -v : v;
        out->V[(size_t)m2 * J + j2] = vf;
        if (out->scores) out->scores[(size_t)m2 * J + j2] = vf * hD[m2];    // R/fit.R:310
      }
      out->D[m2] = hD[m2];
    }
    for (int k = 0; k < nbatch; ++k)
      for (int64_t i2 = 0; i2 < I; ++i2) out->alpha[(size_t)k * I + i2] = hA[(size_t)k * ldI + i2];
    for (int64_t j2 = 0; j2 < J; ++j2) out->beta[j2] = hB[j2];

    out->x_rank = -1;
    {
      const Factor& fx = st.slot[cur];
      if (fx.kind == XK_ZERO) out->x_rank = 0;
      else if (fx.kind == XK_LOWRANK) out->x_rank = fx.M;
      if (out->Xu && out->Xv && out->x_rank >= 0) {
        std::vector<double> xU((size_t)ldI * M), xV((size_t)ldJ * M), xD(M);
        if (fx.kind == XK_LOWRANK) {
          SCGBM_CUDA(cudaMemcpyAsync(xU.data(), fx.U.get(), xU.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
          SCGBM_CUDA(cudaMemcpyAsync(xV.data(), fx.V.get(), xV.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
          SCGBM_CUDA(cudaMemcpyAsync(xD.data(), fx.d.get(), M * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
          ctx.sync();
          ctx.prof.d2h_bytes += (double)(xU.size() + xV.size()) * sizeof(double);
        }
        for (int m2 = 0; m2 < M; ++m2) {
          for (int64_t i2 = 0; i2 < I; ++i2)
            out->Xu[(size_t)m2 * I + i2] = fx.kind == XK_LOWRANK ? xU[(size_t)m2 * ldI + i2] * xD[m2] : 0.0;
          for (int64_t j2 = 0; j2 < J; ++j2)
            out->Xv[(size_t)m2 * J + j2] = fx.kind == XK_LOWRANK ? xV[(size_t)m2 * ldJ + j2] : 0.0;
        }
      }
    }
    if (a.return_W && out->W) {                          // R/fit.R:189-191 (unclipped, current X)
      const Factor& fc = st.slot[cur];
      int64_t jc_max = std::max<int64_t>(1, (int64_t)(64ll << 20) / I);
      jc_max = std::min(jc_max, J);
      DevBuf<double> Wd(I * jc_max);
      for (int64_t j0 = 0; j0 < J; j0 += jc_max) {
        const int64_t jc = std::min(jc_max, J - j0);
        {
          ProfScope ps(ctx.prof, SCGBM_K_MISC, 1);
          write_w_kernel<<<GRID1(I * jc)>>>(I, ldI, j0, jc, ldJ, fc.M, fc.kind, fc.U.get(), fc.d.get(), fc.V.get(),
                                            st.alpha.get(), st.beta.get(), st.Y.batch.get(), st.alpha0.get(),
                                            st.beta0.get(), Wd.get());
          SCGBM_LAUNCH_CHECK();
        }
        SCGBM_CUDA(cudaMemcpyAsync(out->W + j0 * I, Wd.get(), I * jc * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
        ctx.sync();
        ctx.prof.d2h_bytes += (double)I * jc * sizeof(double);
      }
    }
    fill_profile(&out->prof);
  }
};

static void check_fit_args(const scgbm_fit_args* a) {
  SCGBM_REQUIRE(a, "NULL args");
  SCGBM_REQUIRE(a->I >= 1 && a->J >= 1, "Y must be a non-empty matrix");
  SCGBM_REQUIRE(a->Y || (a->csc_p && a->csc_i && a->csc_x), "Y is NULL (pass a dense matrix or the csc_* arrays)");
  SCGBM_REQUIRE(a->M >= 1, "M must be >= 1");
  SCGBM_REQUIRE(a->max_iter >= 1, "max.iter must be >= 1");
  SCGBM_REQUIRE(a->ngpu <= 1 || a->comm == nullptr, "ngpu > 1 and comm are mutually exclusive");
}

FitRun* fit_begin_impl(const scgbm_fit_args* a) {
  check_fit_args(a);
  SCGBM_REQUIRE(a->ngpu <= 1, "the stepwise handle drives one GPU (or one rank with comm); use scgbm_fit for ngpu > 1");
  return new FitRun(*a);
}
int fit_step_impl(FitRun* r, int n_steps) {
  SCGBM_REQUIRE(r, "NULL handle");
  SCGBM_CUDA(cudaSetDevice(r->ctx.device));
  for (int k = 0; k < n_steps; ++k)
    if (!r->step()) return 1;
  return r->finished ? 1 : 0;
}
void fit_profile_impl(FitRun* r, scgbm_profile* p) {
  SCGBM_REQUIRE(r && p, "NULL handle");
  SCGBM_CUDA(cudaSetDevice(r->ctx.device));
  r->fill_profile(p);
}
void fit_set_profile_impl(FitRun* r, int level) {
  SCGBM_REQUIRE(r, "NULL handle");
  SCGBM_CUDA(cudaSetDevice(r->ctx.device));
  r->ctx.prof.resolve();
  r->ctx.prof.timing = level != 0;
  r->ctx.prof.fused_only = level == 2;
}
void fit_end_impl(FitRun* r, scgbm_fit_out* out) {
  SCGBM_REQUIRE(r, "NULL handle");
  SCGBM_CUDA(cudaSetDevice(r->ctx.device));
  try {
    if (out) r->finish(out);
  } catch (...) { delete r; throw; }
  delete r;
}
void fit_impl(const scgbm_fit_args* a, scgbm_fit_out* out) {
  check_fit_args(a);
  SCGBM_REQUIRE(out && out->U && out->D && out->V && out->alpha && out->beta && out->obj, "output buffers missing");
  FitRun* r = new FitRun(*a);
  try {
    while (r->step()) {}
    r->finish(out);
  } catch (...) { delete r; throw; }
  delete r;
}

}  // namespace scgbm

// ---------------------------------------------------------------------------------
// diagnostic entry points (kernel-level parity tests)
// ---------------------------------------------------------------------------------
namespace scgbm {

static double host_absmax(const float* x, int64_t n) {
  double m = 0.0;
  for (int64_t i = 0; i < n; ++i) m = std::max(m, (double)std::fabs(x[i]));
  return m;
}

static void upload_padded(Ctx& ctx, const double* host, int64_t rows, int cols, int64_t ld, double* dev) {
  SCGBM_CUDA(cudaMemsetAsync(dev, 0, (size_t)ld * cols * sizeof(double), ctx.stream));
  SCGBM_CUDA(cudaMemcpy2DAsync(dev, ld * sizeof(double), host, rows * sizeof(double), rows * sizeof(double), cols,
                               cudaMemcpyHostToDevice, ctx.stream));
}

void dbg_eval_impl(const scgbm_eval_args* a, scgbm_eval_out* out) {
  SCGBM_REQUIRE(a && out && a->Y && a->beta_in && out->alpha && out->beta, "NULL args");
  Ctx ctx(a->device, nullptr);
  Counts Y;
  load_counts(ctx, a->Y, a->y_dtype, 0, a->I, a->J, a->ldY, nullptr, 1, a->y_storage, Y);
  const int64_t I = a->I, J = a->J, ldI = Y.ld, ldJ = round_up(J, kRowPad);
  const int Mx = a->Mx;
  PassGeom geom{I, J, ldI, ldJ, 1, nullptr};
  DevBuf<double> Xu(std::max<int64_t>(ldI * Mx, 1)), Xv(std::max<int64_t>(ldJ * Mx, 1)), alpha(ldI), beta(ldJ),
      logrs(ldI), logcs(ldJ), S(ldI), Cc(ldJ), shift(1), scal(8), tmp(std::max(ldI, ldJ));
  DevBuf<float> A32(std::max<int64_t>(ldI * Mx, 1)), B32(std::max<int64_t>(ldJ * Mx, 1)), f_row(ldI + 128), la_row(ldI),
      g_col(ldJ + 128), lb_col(ldJ), rs32(ldI), cs32(ldJ);
  g_col.zero(ctx.stream); f_row.zero(ctx.stream);
  DevBuf<uint16_t> Rhi(ldI * J), Rlo(ldI * J);
  ResidBuf R; R.hi = Rhi.get(); R.lo = Rlo.get(); R.scale = resid_scale_for(Y.maxY);
  alpha.zero(ctx.stream); beta.zero(ctx.stream);
  SCGBM_CUDA(cudaMemcpyAsync(beta.get(), a->beta_in, J * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  XDesc x;
  Factor16 a16, b16;
  if (Mx > 0) {
    upload_padded(ctx, a->Xu, I, Mx, ldI, Xu.get());
    upload_padded(ctx, a->Xv, J, Mx, ldJ, Xv.get());
    const float* rs = nullptr; const float* cs = nullptr;
    if (a->rowscale) {
      upload_padded(ctx, a->rowscale, I, 1, ldI, tmp.get());
      exp_stage_kernel<<<GRID1(ldI)>>>(I, ldI, 1, tmp.get(), 1.0, nullptr, rs32.get());
      rs = rs32.get();
    }
    stage_factor_kernel<<<GRID1(ldI * Mx)>>>(I, ldI, Mx, Xu.get(), nullptr, rs, A32.get());
    if (a->colscale) {
      upload_padded(ctx, a->colscale, J, 1, ldJ, tmp.get());
      exp_stage_kernel<<<GRID1(ldJ)>>>(J, ldJ, 1, tmp.get(), 1.0, nullptr, cs32.get());
      cs = cs32.get();
    }
    stage_factor_kernel<<<GRID1(ldJ * Mx)>>>(J, ldJ, Mx, Xv.get(), nullptr, cs, B32.get());
    SCGBM_LAUNCH_CHECK();
    x.A = A32.get(); x.B = B32.get(); x.M = Mx; x.clamp = a->clamp8;
    if (Mx <= 64) {
      stage_factor16(ctx, a16, I, ldI, Mx, 0, Xu.get(), nullptr, rs);
      stage_factor16(ctx, b16, J, ldJ, Mx, 1, Xv.get(), nullptr, cs);
      x.A16 = &a16; x.B16 = &b16;
    }
  }
  const bool tc_eta = eta_tc_supported(geom, x, Y.storage);
  EtaWs eta;
  log_kernel<<<GRID1(J)>>>(J, Y.colsum.get(), logcs.get());
  log_kernel<<<GRID1(ldI)>>>(ldI, Y.rowsum.get(), logrs.get());
  exp_stage_kernel<<<GRID1(ldJ)>>>(J, ldJ, 1, beta.get(), 1.0, g_col.get(), lb_col.get());
  if (tc_eta) pass_rowsum_tc(ctx, eta, geom, x, g_col.get(), S.get());
  else pass_rowsum(ctx, eta, geom, x, g_col.get(), S.get());
  alpha_update_kernel<<<1, 1024, 0, ctx.stream>>>(I, ldI, 1, logrs.get(), S.get(), nullptr, alpha.get(), shift.get());
  beta_shift_kernel<<<GRID1(J)>>>(J, beta.get(), shift.get());
  exp_stage_kernel<<<GRID1(ldI)>>>(I, ldI, 1, alpha.get(), 1.0, f_row.get(), la_row.get());
  if (a->infer_beta) {
    if (tc_eta) pass_colsum_tc(ctx, eta, geom, x, f_row.get(), Cc.get());
    else pass_colsum(ctx, eta, geom, x, f_row.get(), Cc.get());
    beta_infer_kernel<<<GRID1(J)>>>(J, logcs.get(), Cc.get(), beta.get());
  }
  exp_stage_kernel<<<GRID1(ldJ)>>>(J, ldJ, 1, beta.get(), 1.0, g_col.get(), lb_col.get());
  if (tc_eta)
    pass_fused_tc(ctx, eta, geom, x, Y, f_row.get(), g_col.get(), la_row.get(), lb_col.get(), alpha.get(), beta.get(),
                  Y.rowsum.get(), (float)Y.maxY, R, scal.get());
  else
    pass_fused(ctx, eta, geom, x, Y, f_row.get(), g_col.get(), la_row.get(), lb_col.get(), (float)Y.maxY, R, scal.get());
  double h[4];
  SCGBM_CUDA(cudaMemcpyAsync(h, scal.get(), sizeof(h), cudaMemcpyDeviceToHost, ctx.stream));
  ctx.sync();
  if (tc_eta && h[3] != 0.0) {   // possible exp() underflow (R/fit.R:152-157): the CUDA-core pass restates it, as in the fit
    pass_fused(ctx, eta, geom, x, Y, f_row.get(), g_col.get(), la_row.get(), lb_col.get(), (float)Y.maxY, R, scal.get());
    SCGBM_CUDA(cudaMemcpyAsync(h, scal.get(), 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  }
  SCGBM_CUDA(cudaMemcpyAsync(out->alpha, alpha.get(), I * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  SCGBM_CUDA(cudaMemcpyAsync(out->beta, beta.get(), J * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  ctx.sync();
  out->sum_w = h[0]; out->sum_ylogw = h[1]; out->max_w = h[2]; out->max_Y = Y.maxY; out->LL = h[1] - h[0];
  if (out->resid) {
    DevBuf<double> Rd(ldI * J);
    planes_to_f64_kernel<<<GRID1(ldI * J)>>>(ldI * J, R.hi, R.lo, 1.0 / R.scale, Rd.get());
    SCGBM_LAUNCH_CHECK();
    SCGBM_CUDA(cudaMemcpy2DAsync(out->resid, I * sizeof(double), Rd.get(), ldI * sizeof(double), I * sizeof(double), J,
                                 cudaMemcpyDeviceToHost, ctx.stream));
    ctx.sync();
  }
}

void dbg_tsvd_impl(const scgbm_tsvd_args* a, scgbm_tsvd_out* out) {
  SCGBM_REQUIRE(a && out && out->U && out->D && out->V, "NULL args");
  SCGBM_REQUIRE(a->nterms >= 0 && a->nterms <= 2, "nterms must be 0..2");
  Ctx ctx(a->device, nullptr);
  const int64_t I = a->I, J = a->J, ldI = round_up(I, kRowPad), ldJ = round_up(J, kRowPad);
  const int M = a->M;
  GOperator op;
  op.I = I; op.ldI = ldI; op.J = J; op.ldJ = ldJ; op.J_total = J;
  DevBuf<float> R;
  DevBuf<uint16_t> Rhi, Rlo;
  if (a->R) {
    R.alloc(ldI * J); Rhi.alloc(ldI * J); Rlo.alloc(ldI * J);
    R.zero(ctx.stream);
    SCGBM_CUDA(cudaMemcpy2DAsync(R.get(), ldI * sizeof(float), a->R, I * sizeof(float), I * sizeof(float), J,
                                 cudaMemcpyHostToDevice, ctx.stream));
    op.R.hi = Rhi.get(); op.R.lo = Rlo.get(); op.c = a->c;
    op.R.scale = resid_scale_for(host_absmax(a->R, I * J));
    split_planes(ctx, R.get(), ldI * J, op.R);
  }
  DevBuf<double> tA[2], td[2], tB[2];
  for (int t = 0; t < a->nterms; ++t) {
    const int Mt = a->Mt[t];
    tA[t].alloc(ldI * Mt); td[t].alloc(Mt); tB[t].alloc(ldJ * Mt);
    upload_padded(ctx, a->A[t], I, Mt, ldI, tA[t].get());
    upload_padded(ctx, a->B[t], J, Mt, ldJ, tB[t].get());
    SCGBM_CUDA(cudaMemcpyAsync(td[t].get(), a->d[t], Mt * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
    LowRankTerm& lt = op.t[op.nterms++];
    lt.A = tA[t].get(); lt.lda = ldI; lt.d = td[t].get(); lt.B = tB[t].get(); lt.ldb = ldJ; lt.M = Mt; lt.coef = a->coef[t];
  }
  TsvdOpts o;
  o.M = M;
  o.extra = a->svd_extra > 0 ? a->svd_extra : 12;
  o.depth = a->svd_depth > 0 ? a->svd_depth : 3;
  o.tol = a->svd_tol > 0 ? a->svd_tol : 1e-4;
  o.seed = a->seed;
  o.variant = a->kernel_variant;
  TsvdWs ws;
  if (a->V0) tsvd_set_warm_start(ctx, ws, op, o, a->V0);
  DevBuf<double> U(ldI * M), D(M), V(ldJ * M);
  int passes = 0;
  double est = 0;
  tsvd(ctx, ws, op, o, U.get(), D.get(), V.get(), &passes, &est);
  SCGBM_CUDA(cudaMemcpy2DAsync(out->U, I * sizeof(double), U.get(), ldI * sizeof(double), I * sizeof(double), M,
                               cudaMemcpyDeviceToHost, ctx.stream));
  SCGBM_CUDA(cudaMemcpy2DAsync(out->V, J * sizeof(double), V.get(), ldJ * sizeof(double), J * sizeof(double), M,
                               cudaMemcpyDeviceToHost, ctx.stream));
  SCGBM_CUDA(cudaMemcpyAsync(out->D, D.get(), M * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  ctx.sync();
  out->passes = passes;
  out->est = est;
}

void dbg_prod_impl(const scgbm_prod_args* a, scgbm_prod_out* out) {
  SCGBM_REQUIRE(a && out && a->R && a->P && a->Q && out->outN && out->outT, "NULL args");
  Ctx ctx(a->device, nullptr);
  const int64_t I = a->I, J = a->J, ldI = round_up(I, kRowPad), ldJ = round_up(J, kRowPad);
  const int b = a->b;
  DevBuf<float> R(ldI * J);
  DevBuf<uint16_t> Rhi(ldI * J), Rlo(ldI * J);
  R.zero(ctx.stream);
  SCGBM_CUDA(cudaMemcpy2DAsync(R.get(), ldI * sizeof(float), a->R, I * sizeof(float), I * sizeof(float), J,
                               cudaMemcpyHostToDevice, ctx.stream));
  ResidBuf rb; rb.hi = Rhi.get(); rb.lo = Rlo.get();
  rb.scale = resid_scale_for(host_absmax(a->R, I * J));
  split_planes(ctx, R.get(), ldI * J, rb);
  DevBuf<double> P(ldJ * b), Q(ldI * b), oN(ldI * b), oT(ldJ * b);
  upload_padded(ctx, a->P, J, b, ldJ, P.get());
  upload_padded(ctx, a->Q, I, b, ldI, Q.get());
  ProdWs ws;
  const int reps = std::max(1, a->reps);
  cudaEvent_t e0, e1;
  SCGBM_CUDA(cudaEventCreate(&e0)); SCGBM_CUDA(cudaEventCreate(&e1));
  float ms = 0;
  for (int pass = 0; pass < 2; ++pass) {          // pass 0 warms up, pass 1 is timed
    SCGBM_CUDA(cudaEventRecord(e0, ctx.stream));
    for (int r = 0; r < (pass ? reps : 1); ++r) {
      oN.zero(ctx.stream);
      prod_n(ctx, ws, rb, I, ldI, J, P.get(), ldJ, b, a->c, oN.get(), ldI, a->variant);
    }
    SCGBM_CUDA(cudaEventRecord(e1, ctx.stream));
    ctx.sync();
    cudaEventElapsedTime(&ms, e0, e1);
  }
  out->ms_n = ms / reps;
  for (int pass = 0; pass < 2; ++pass) {
    SCGBM_CUDA(cudaEventRecord(e0, ctx.stream));
    for (int r = 0; r < (pass ? reps : 1); ++r) {
      oT.zero(ctx.stream);
      prod_t(ctx, ws, rb, I, ldI, J, Q.get(), ldI, b, a->c, oT.get(), ldJ, a->variant);
    }
    SCGBM_CUDA(cudaEventRecord(e1, ctx.stream));
    ctx.sync();
    cudaEventElapsedTime(&ms, e0, e1);
  }
  out->ms_t = ms / reps;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  SCGBM_CUDA(cudaMemcpy2DAsync(out->outN, I * sizeof(double), oN.get(), ldI * sizeof(double), I * sizeof(double), b,
                               cudaMemcpyDeviceToHost, ctx.stream));
  SCGBM_CUDA(cudaMemcpy2DAsync(out->outT, J * sizeof(double), oT.get(), ldJ * sizeof(double), J * sizeof(double), b,
                               cudaMemcpyDeviceToHost, ctx.stream));
  ctx.sync();
  if (out->Rq) {
    DevBuf<double> Rd(ldI * J);
    planes_to_f64_kernel<<<GRID1(ldI * J)>>>(ldI * J, rb.hi, rb.lo, 1.0 / rb.scale, Rd.get());
    SCGBM_LAUNCH_CHECK();
    SCGBM_CUDA(cudaMemcpy2DAsync(out->Rq, I * sizeof(double), Rd.get(), ldI * sizeof(double), I * sizeof(double), J,
                                 cudaMemcpyDeviceToHost, ctx.stream));
    ctx.sync();
  }
}

}  // namespace scgbm
