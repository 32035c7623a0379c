// eta_pass.cu -- K3/K4 (CUDA-core version): tiles of eta are rebuilt from factor
// tiles staged in shared memory; Y is read once, coalesced and vectorised along
// genes; the residual is written once; reductions go warp-shuffle -> CTA -> fixed
// order second stage so that every run is bit-reproducible.
//
//   rowsum : S_ik = sum_{j in k} exp(beta_j + X_ij)                     R/fit.R:127-130
//   colsum : C_j  = sum_i exp(alpha_i,b(j) + X_ij)                      R/fit.R:133-135
//   fused  : W = min(exp(alpha+beta+X), max.Y); sum W; sum_nz Y log W;
//            max W; resid = Y - W                                       R/fit.R:136-151,321,323
#include "eta_pass.cuh"
#include "products.cuh"

namespace scgbm {

constexpr int ET_TI = 128;   // genes per tile  (32 lanes x 4 consecutive genes)
constexpr int ET_TJ = 64;    // cells per tile  (8 warps x 8 consecutive cells)
constexpr int ET_NT = 256;

enum { MODE_ROWSUM = 0, MODE_FUSED = 1, MODE_ADDX = 2 };

struct EtaParams {
  int64_t I, J, ldI, ldJ;
  int nbatch;
  const int32_t* batch;
  const float* A; const float* B; int M; int clamp; const float* rowscale;
  const float* f_row; const float* g_col; const float* la_row; const float* lb_col;
  const void* Y; uint16_t* Rhi; uint16_t* Rlo;
  float maxY, log_maxY, coef, rscale;
  int tiles_per_chunk; int nJt;
  double* row_part; double* scal_part;
};

template <typename YT> __device__ __forceinline__ void load_y4(const YT* p, float (&y)[4]);
template <> __device__ __forceinline__ void load_y4<uint8_t>(const uint8_t* p, float (&y)[4]) {
  const uchar4 v = __ldg(reinterpret_cast<const uchar4*>(p));
  y[0] = v.x; y[1] = v.y; y[2] = v.z; y[3] = v.w;
}
template <> __device__ __forceinline__ void load_y4<uint16_t>(const uint16_t* p, float (&y)[4]) {
  const ushort4 v = __ldg(reinterpret_cast<const ushort4*>(p));
  y[0] = v.x; y[1] = v.y; y[2] = v.z; y[3] = v.w;
}
template <> __device__ __forceinline__ void load_y4<float>(const float* p, float (&y)[4]) {
  const float4 v = __ldg(reinterpret_cast<const float4*>(p));
  y[0] = v.x; y[1] = v.y; y[2] = v.z; y[3] = v.w;
}

__device__ __forceinline__ float clamp8(float x) {
  x = x > 8.0f ? 8.0f : x;     // NaN stays NaN, as in R's X[X > 8] <- 8
  x = x < -8.0f ? -8.0f : x;
  return x;
}

// acc[a][b] = sum_m U_s[m][4tx+a] * V_s[m][8ty+b]
__device__ __forceinline__ void tile_dot(const float* __restrict__ U_s, const float* __restrict__ V_s,
                                         int M, int tx, int ty, float (&acc)[4][8]) {
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.0f;
#pragma unroll 2
  for (int m = 0; m < M; ++m) {
    const float4 u = *reinterpret_cast<const float4*>(U_s + m * ET_TI + 4 * tx);
    const float4 v0 = *reinterpret_cast<const float4*>(V_s + m * ET_TJ + 8 * ty);
    const float4 v1 = *reinterpret_cast<const float4*>(V_s + m * ET_TJ + 8 * ty + 4);
    const float uu[4] = {u.x, u.y, u.z, u.w};
    const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(uu[a], vv[b], acc[a][b]);
  }
}

template <int MODE, typename YT, bool MB>
__global__ void __launch_bounds__(ET_NT, 2)
eta_rows_kernel(const EtaParams p) {
  extern __shared__ __align__(16) float sm[];
  const int M = p.M;
  float* U_s = sm;
  float* V_s = U_s + (size_t)M * ET_TI;
  float* gj_s = V_s + (size_t)M * ET_TJ;
  float* lb_s = gj_s + ET_TJ;
  int* bt_s = reinterpret_cast<int*>(lb_s + ET_TJ);
  double* red = reinterpret_cast<double*>(bt_s + ET_TJ);   // ROWSUM: [8][128] or MB: [nbatch][128]

  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t i0 = (int64_t)blockIdx.x * ET_TI;
  const int chunk = blockIdx.y;
  const int64_t ig = i0 + 4 * tx;            // first of this thread's 4 genes
  const bool row_ok = ig < p.ldI;

  for (int idx = threadIdx.x; idx < M * ET_TI; idx += ET_NT) {
    const int m = idx / ET_TI, ii = idx % ET_TI;
    U_s[idx] = (i0 + ii < p.ldI) ? p.A[(int64_t)m * p.ldI + i0 + ii] : 0.0f;
  }
  if (MODE == MODE_ROWSUM && MB) {
    for (int idx = threadIdx.x; idx < p.nbatch * ET_TI; idx += ET_NT) red[idx] = 0.0;
  }
  float f[4] = {0, 0, 0, 0}, la[4] = {0, 0, 0, 0};
  if (MODE == MODE_FUSED && !MB && row_ok) {
#pragma unroll
    for (int a = 0; a < 4; ++a) { f[a] = p.f_row[ig + a]; la[a] = p.la_row[ig + a]; }
  }
  double rs[4] = {0, 0, 0, 0};
  double sw = 0.0, syl = 0.0;
  float mw = 0.0f;

  const int jt_begin = chunk * p.tiles_per_chunk;
  const int jt_end = min(p.nJt, jt_begin + p.tiles_per_chunk);
  for (int jt = jt_begin; jt < jt_end; ++jt) {
    const int64_t j0 = (int64_t)jt * ET_TJ;
    __syncthreads();
    for (int idx = threadIdx.x; idx < M * ET_TJ; idx += ET_NT) {
      const int m = idx / ET_TJ, jj = idx % ET_TJ;
      V_s[idx] = (j0 + jj < p.J) ? p.B[(int64_t)m * p.ldJ + j0 + jj] : 0.0f;
    }
    if (threadIdx.x < ET_TJ) {
      const int64_t j = j0 + threadIdx.x;
      const bool ok = j < p.J;
      if (MODE != MODE_ADDX) gj_s[threadIdx.x] = ok ? p.g_col[j] : 0.0f;
      if (MODE == MODE_FUSED) lb_s[threadIdx.x] = ok ? p.lb_col[j] : 0.0f;
      bt_s[threadIdx.x] = (ok && p.batch) ? p.batch[j] : 0;
    }
    __syncthreads();

    float acc[4][8];
    tile_dot(U_s, V_s, M, tx, ty, acc);

    if (MODE == MODE_ROWSUM) {
      float trow[4] = {0, 0, 0, 0};
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const int jj = 8 * ty + b;
        const float gj = gj_s[jj];
        const int kb = MB ? bt_s[jj] : 0;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          float x = acc[a][b];
          if (MB && p.rowscale && row_ok) x *= p.rowscale[(int64_t)kb * p.ldI + ig + a];
          if (p.clamp) x = clamp8(x);
          const float e = exp_fast_accurate(x) * gj;
          if (MB) { if (gj != 0.0f) atomicAdd(&red[kb * ET_TI + 4 * tx + a], (double)e); }
          else trow[a] += e;
        }
      }
      if (!MB) {
#pragma unroll
        for (int a = 0; a < 4; ++a) rs[a] += (double)trow[a];
      }
    } else {
      float tsw = 0.0f, tsyl = 0.0f;
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const int jj = 8 * ty + b;
        const int64_t j = j0 + jj;
        if (j >= p.J || !row_ok) continue;
        const int kb = MB ? bt_s[jj] : 0;
        float xv[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          float x = acc[a][b];
          if (MB && p.rowscale) x *= p.rowscale[(int64_t)kb * p.ldI + ig + a];
          if (p.clamp) x = clamp8(x);
          xv[a] = x;
        }
        ushort4* rph = reinterpret_cast<ushort4*>(p.Rhi + j * p.ldI + ig);
        ushort4* rpl = reinterpret_cast<ushort4*>(p.Rlo + j * p.ldI + ig);
        if (MODE == MODE_ADDX) {
          ushort4 h = *rph, l = *rpl;
          const float r0 = fmaf(p.coef, xv[0], join_f16x2(h.x, l.x)), r1 = fmaf(p.coef, xv[1], join_f16x2(h.y, l.y));
          const float r2 = fmaf(p.coef, xv[2], join_f16x2(h.z, l.z)), r3 = fmaf(p.coef, xv[3], join_f16x2(h.w, l.w));
          split_f16x2(r0, h.x, l.x); split_f16x2(r1, h.y, l.y); split_f16x2(r2, h.z, l.z); split_f16x2(r3, h.w, l.w);
          *rph = h; *rpl = l;
        } else {
          float y[4];
          load_y4<YT>(reinterpret_cast<const YT*>(p.Y) + j * p.ldI + ig, y);
          const float gj = gj_s[jj], lbj = lb_s[jj];
          float rr[4];
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            float fa = f[a], laa = la[a];
            if (MB) { fa = p.f_row[(int64_t)kb * p.ldI + ig + a]; laa = p.la_row[(int64_t)kb * p.ldI + ig + a]; }
            float w = exp_fast_accurate(xv[a]) * (fa * gj);
            float lw = xv[a] + (laa + lbj);
            if (w > p.maxY) { w = p.maxY; lw = p.log_maxY; }     // R/fit.R:145 (NaN is kept)
            tsw += w;
            mw = (w > mw || w != w) ? w : mw;
            if (y[a] != 0.0f) {                                   // R/fit.R:151, nz only
              if (lw < -745.13f) lw = -INFINITY;                  // exp() underflows to 0 in R's doubles
              tsyl = fmaf(y[a], lw, tsyl);
            }
            rr[a] = (y[a] - w) * p.rscale;
          }
          ushort4 h, l;
          split_f16x2(rr[0], h.x, l.x); split_f16x2(rr[1], h.y, l.y); split_f16x2(rr[2], h.z, l.z); split_f16x2(rr[3], h.w, l.w);
          *rph = h; *rpl = l;
        }
      }
      if (MODE == MODE_FUSED) { sw += (double)tsw; syl += (double)tsyl; }
    }
  }

  if (MODE == MODE_ROWSUM) {
    __syncthreads();
    if (!MB) {
#pragma unroll
      for (int a = 0; a < 4; ++a) red[ty * ET_TI + 4 * tx + a] = rs[a];
      __syncthreads();
      if (threadIdx.x < ET_TI && i0 + threadIdx.x < p.ldI) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w * ET_TI + threadIdx.x];
        p.row_part[(int64_t)chunk * p.ldI + i0 + threadIdx.x] = s;
      }
    } else {
      for (int idx = threadIdx.x; idx < p.nbatch * ET_TI; idx += ET_NT) {
        const int k = idx / ET_TI, ii = idx % ET_TI;
        if (i0 + ii < p.ldI) p.row_part[((int64_t)chunk * p.nbatch + k) * p.ldI + i0 + ii] = red[idx];
      }
    }
  }
  if (MODE == MODE_FUSED) {
    __syncthreads();
    sw = warp_sum(sw); syl = warp_sum(syl);
    // NaN-propagating max
    float mwr = mw;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float other = __shfl_xor_sync(0xffffffffu, mwr, o);
      mwr = (other > mwr || other != other) ? other : mwr;
    }
    if (tx == 0) { red[ty] = sw; red[8 + ty] = syl; red[16 + ty] = (double)mwr; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0.0, b = 0.0, c = 0.0;
      for (int w = 0; w < 8; ++w) {
        a += red[w]; b += red[8 + w];
        const double o = red[16 + w];
        c = (o > c || o != o) ? o : c;
      }
      const int64_t cta = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
      p.scal_part[cta * 4 + 0] = a; p.scal_part[cta * 4 + 1] = b; p.scal_part[cta * 4 + 2] = c;
    }
  }
}

// second stage: fixed-order sums ------------------------------------------------
__global__ void rowpart_reduce_kernel(int nchunk, int64_t n, const double* __restrict__ part, double* __restrict__ S) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  double s = 0.0;
  for (int c = 0; c < nchunk; ++c) s += part[(int64_t)c * n + idx];
  S[idx] = s;
}

__global__ void __launch_bounds__(1024)
scal_reduce_kernel(int64_t ncta, const double* __restrict__ part, double* __restrict__ out) {
  __shared__ double s0[1024], s1[1024], s2[1024];
  double a = 0.0, b = 0.0, c = 0.0;
  for (int64_t k = threadIdx.x; k < ncta; k += 1024) {
    a += part[k * 4 + 0]; b += part[k * 4 + 1];
    const double o = part[k * 4 + 2];
    c = (o > c || o != o) ? o : c;
  }
  s0[threadIdx.x] = a; s1[threadIdx.x] = b; s2[threadIdx.x] = c;
  __syncthreads();
  for (int st = 512; st > 0; st >>= 1) {
    if (threadIdx.x < st) {
      s0[threadIdx.x] += s0[threadIdx.x + st];
      s1[threadIdx.x] += s1[threadIdx.x + st];
      const double o = s2[threadIdx.x + st], m = s2[threadIdx.x];
      s2[threadIdx.x] = (o > m || o != o) ? o : m;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[0] = s0[0]; out[1] = s1[0]; out[2] = s2[0]; out[3] = 0.0; }
}

void scal_reduce_launch(Ctx& ctx, int64_t ncta, const double* part, double* out3) {
  scal_reduce_kernel<<<1, 1024, 0, ctx.stream>>>(ncta, part, out3);
  SCGBM_LAUNCH_CHECK();
}
void rowpart_reduce_launch(Ctx& ctx, int nchunk, int64_t n, const double* part, double* S) {
  ProfScope ps(ctx.prof, SCGBM_K_MISC, 1);
  rowpart_reduce_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, ctx.stream>>>(nchunk, n, part, S);
  SCGBM_LAUNCH_CHECK();
}

// column sums (infer.beta): one CTA per cell tile, loops over all gene tiles -----
template <bool MB>
__global__ void __launch_bounds__(ET_NT, 2)
eta_colsum_kernel(const EtaParams p, double* __restrict__ C) {
  extern __shared__ __align__(16) float sm[];
  const int M = p.M;
  float* U_s = sm;
  float* V_s = U_s + (size_t)M * ET_TI;
  float* fi_s = V_s + (size_t)M * ET_TJ;            // [128] f_row for single batch
  int* bt_s = reinterpret_cast<int*>(fi_s + ET_TI);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t j0 = (int64_t)blockIdx.x * ET_TJ;
  for (int idx = threadIdx.x; idx < M * ET_TJ; idx += ET_NT) {
    const int m = idx / ET_TJ, jj = idx % ET_TJ;
    V_s[idx] = (j0 + jj < p.J) ? p.B[(int64_t)m * p.ldJ + j0 + jj] : 0.0f;
  }
  if (threadIdx.x < ET_TJ) {
    const int64_t j = j0 + threadIdx.x;
    bt_s[threadIdx.x] = (j < p.J && p.batch) ? p.batch[j] : 0;
  }
  double cs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const int nIt = (int)((p.ldI + ET_TI - 1) / ET_TI);
  for (int it = 0; it < nIt; ++it) {
    const int64_t i0 = (int64_t)it * ET_TI;
    __syncthreads();
    for (int idx = threadIdx.x; idx < M * ET_TI; idx += ET_NT) {
      const int m = idx / ET_TI, ii = idx % ET_TI;
      U_s[idx] = (i0 + ii < p.ldI) ? p.A[(int64_t)m * p.ldI + i0 + ii] : 0.0f;
    }
    if (threadIdx.x < ET_TI) fi_s[threadIdx.x] = (i0 + threadIdx.x < p.ldI) ? p.f_row[i0 + threadIdx.x] : 0.0f;
    __syncthreads();
    float acc[4][8];
    tile_dot(U_s, V_s, M, tx, ty, acc);
    const int64_t ig = i0 + 4 * tx;
    const bool row_ok = ig < p.ldI;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const int kb = MB ? bt_s[8 * ty + b] : 0;
      float t = 0.0f;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        float x = acc[a][b];
        float fa = fi_s[4 * tx + a];
        if (MB && row_ok) {
          fa = p.f_row[(int64_t)kb * p.ldI + ig + a];
          if (p.rowscale) x *= p.rowscale[(int64_t)kb * p.ldI + ig + a];
        }
        if (p.clamp) x = clamp8(x);
        t += (fa != 0.0f) ? exp_fast_accurate(x) * fa : 0.0f;
      }
      cs[b] += (double)t;
    }
  }
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const double s = warp_sum(cs[b]);
    const int64_t j = j0 + 8 * ty + b;
    if (tx == 0 && j < p.J) C[j] = s;
  }
}

// ---------------------------------------------------------------------------------
static size_t eta_smem_bytes(int M, int nbatch, int mode, bool mb) {
  size_t s = (size_t)M * (ET_TI + ET_TJ) * sizeof(float) + 3 * ET_TJ * sizeof(float);
  s = (s + 15) / 16 * 16;
  size_t red = 8 * ET_TI * sizeof(double);
  if (mode == MODE_ROWSUM && mb) red = std::max(red, (size_t)nbatch * ET_TI * sizeof(double));
  return s + red + 64;
}

static void fill_params(EtaParams& p, const PassGeom& g, const XDesc& x) {
  memset(&p, 0, sizeof(p));
  p.I = g.I; p.J = g.J; p.ldI = g.ldI; p.ldJ = g.ldJ; p.nbatch = g.nbatch; p.batch = g.batch;
  p.A = x.A; p.B = x.B; p.M = x.M; p.clamp = x.clamp; p.rowscale = x.rowscale;
  p.nJt = (int)ceil_div(g.J, ET_TJ);
}

static void choose_chunks(Ctx& ctx, const PassGeom& g, int& nIt, int& nchunk, int& tiles_per_chunk) {
  nIt = (int)ceil_div(g.ldI, ET_TI);
  const int nJt = (int)ceil_div(g.J, ET_TJ);
  int want = (int)ceil_div((int64_t)ctx.num_sms * 8, nIt);
  nchunk = std::max(1, std::min(want, nJt));
  tiles_per_chunk = (int)ceil_div(nJt, nchunk);
  nchunk = (int)ceil_div(nJt, tiles_per_chunk);
}

template <typename K>
static void set_smem(K kernel, size_t bytes) {
  SCGBM_REQUIRE(bytes <= 220 * 1024, "M too large for the shared-memory factor tiles");
  if (bytes > 48 * 1024)
    SCGBM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

void pass_rowsum(Ctx& ctx, EtaWs& ws, const PassGeom& g, const XDesc& x, const float* g_col, double* S) {
  EtaParams p;
  fill_params(p, g, x);
  p.g_col = g_col;
  int nIt, nchunk, tpc;
  choose_chunks(ctx, g, nIt, nchunk, tpc);
  p.tiles_per_chunk = tpc;
  const int64_t n = g.ldI * g.nbatch;
  if (ws.row_part.n < (int64_t)nchunk * n) ws.row_part.alloc((int64_t)nchunk * n);
  p.row_part = ws.row_part.get();
  const bool mb = g.nbatch > 1;
  const size_t smem = eta_smem_bytes(x.M, g.nbatch, MODE_ROWSUM, mb);
  dim3 grid(nIt, nchunk);
  {
    ProfScope ps(ctx.prof, SCGBM_K_ROWSUM, 1);
    if (mb) {
      set_smem(eta_rows_kernel<MODE_ROWSUM, float, true>, smem);
      eta_rows_kernel<MODE_ROWSUM, float, true><<<grid, ET_NT, smem, ctx.stream>>>(p);
    } else {
      set_smem(eta_rows_kernel<MODE_ROWSUM, float, false>, smem);
      eta_rows_kernel<MODE_ROWSUM, float, false><<<grid, ET_NT, smem, ctx.stream>>>(p);
    }
    SCGBM_LAUNCH_CHECK();
  }
  {
    ProfScope ps(ctx.prof, SCGBM_K_MISC, 1);
    rowpart_reduce_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, ctx.stream>>>(nchunk, n, ws.row_part.get(), S);
    SCGBM_LAUNCH_CHECK();
  }
}

void pass_colsum(Ctx& ctx, EtaWs& ws, const PassGeom& g, const XDesc& x, const float* f_row, double* C) {
  EtaParams p;
  fill_params(p, g, x);
  p.f_row = f_row;
  const bool mb = g.nbatch > 1;
  const size_t smem = (size_t)x.M * (ET_TI + ET_TJ) * sizeof(float) + ET_TI * sizeof(float) + ET_TJ * sizeof(int) + 64;
  ProfScope ps(ctx.prof, SCGBM_K_COLSUM, 1);
  dim3 grid((unsigned)ceil_div(g.J, ET_TJ));
  if (mb) {
    set_smem(eta_colsum_kernel<true>, smem);
    eta_colsum_kernel<true><<<grid, ET_NT, smem, ctx.stream>>>(p, C);
  } else {
    set_smem(eta_colsum_kernel<false>, smem);
    eta_colsum_kernel<false><<<grid, ET_NT, smem, ctx.stream>>>(p, C);
  }
  SCGBM_LAUNCH_CHECK();
}

template <typename YT>
static void launch_fused(Ctx& ctx, const EtaParams& p, dim3 grid, size_t smem, bool mb) {
  if (mb) {
    set_smem(eta_rows_kernel<MODE_FUSED, YT, true>, smem);
    eta_rows_kernel<MODE_FUSED, YT, true><<<grid, ET_NT, smem, ctx.stream>>>(p);
  } else {
    set_smem(eta_rows_kernel<MODE_FUSED, YT, false>, smem);
    eta_rows_kernel<MODE_FUSED, YT, false><<<grid, ET_NT, smem, ctx.stream>>>(p);
  }
  SCGBM_LAUNCH_CHECK();
}

void pass_fused(Ctx& ctx, EtaWs& ws, const PassGeom& g, const XDesc& x, const Counts& y,
                const float* f_row, const float* g_col, const float* la_row, const float* lb_col,
                float maxY, ResidBuf resid, double* out3) {
  EtaParams p;
  fill_params(p, g, x);
  p.f_row = f_row; p.g_col = g_col; p.la_row = la_row; p.lb_col = lb_col;
  p.Y = y.data; p.Rhi = resid.hi; p.Rlo = resid.lo; p.rscale = resid.scale; p.maxY = maxY; p.log_maxY = logf(maxY);
  int nIt, nchunk, tpc;
  choose_chunks(ctx, g, nIt, nchunk, tpc);
  p.tiles_per_chunk = tpc;
  const int64_t ncta = (int64_t)nIt * nchunk;
  if (ws.scal_part.n < ncta * 4) ws.scal_part.alloc(ncta * 4);
  p.scal_part = ws.scal_part.get();
  const bool mb = g.nbatch > 1;
  const size_t smem = eta_smem_bytes(x.M, g.nbatch, MODE_FUSED, mb);
  dim3 grid(nIt, nchunk);
  {
    ProfScope ps(ctx.prof, SCGBM_K_FUSED, 1);
    switch (y.storage) {
      case SCGBM_STORE_U8: launch_fused<uint8_t>(ctx, p, grid, smem, mb); break;
      case SCGBM_STORE_U16: launch_fused<uint16_t>(ctx, p, grid, smem, mb); break;
      case SCGBM_STORE_F32: launch_fused<float>(ctx, p, grid, smem, mb); break;
      default: throw Error(SCGBM_ERR_INVALID, "bad storage");
    }
  }
  {
    ProfScope ps(ctx.prof, SCGBM_K_MISC, 1);
    scal_reduce_kernel<<<1, 1024, 0, ctx.stream>>>(ncta, ws.scal_part.get(), out3);
    SCGBM_LAUNCH_CHECK();
  }
}

void pass_addx(Ctx& ctx, const PassGeom& g, const XDesc& x, float coef, ResidBuf resid) {
  if (x.M == 0 || coef == 0.0f) return;
  EtaParams p;
  fill_params(p, g, x);
  p.Rhi = resid.hi; p.Rlo = resid.lo; p.coef = coef * resid.scale;
  int nIt, nchunk, tpc;
  choose_chunks(ctx, g, nIt, nchunk, tpc);
  p.tiles_per_chunk = tpc;
  const bool mb = g.nbatch > 1 && x.rowscale;
  const size_t smem = eta_smem_bytes(x.M, g.nbatch, MODE_ADDX, mb);
  dim3 grid(nIt, nchunk);
  ProfScope ps(ctx.prof, SCGBM_K_MISC, 1);
  if (mb) {
    set_smem(eta_rows_kernel<MODE_ADDX, float, true>, smem);
    eta_rows_kernel<MODE_ADDX, float, true><<<grid, ET_NT, smem, ctx.stream>>>(p);
  } else {
    set_smem(eta_rows_kernel<MODE_ADDX, float, false>, smem);
    eta_rows_kernel<MODE_ADDX, float, false><<<grid, ET_NT, smem, ctx.stream>>>(p);
  }
  SCGBM_LAUNCH_CHECK();
}

double fused_pass_bytes(const PassGeom& g, const XDesc& x, int y_elem_bytes) {
  // N*s_Y (Y read) + 4N (residual write: two fp16 planes) + 4M(I+J) (factor tiles) + 8(I*nbatch + J) (alpha, beta)
  const double N = (double)g.I * (double)g.J;
  return N * y_elem_bytes + 4.0 * N + 4.0 * x.M * ((double)g.I + (double)g.J) +
         8.0 * ((double)g.I * g.nbatch + (double)g.J);
}

}  // namespace scgbm
