// products.cu -- K5 dispatch + the CUDA-core (FFMA) version of the tall-skinny residual
// products (validation path, and blocks wider than the tcgen05 kernel supports).
// fp32 FMAs inside runs of <= 128 reduction steps, fp64 across runs, fixed-order
// second stage for the split-K partials => bit-reproducible.
#include "products.cuh"

namespace scgbm {

constexpr int PR_NT = 256;

// four consecutive (scaled) residual entries from the two fp16 planes
__device__ __forceinline__ float4 load_resid4(const uint16_t* __restrict__ hi, const uint16_t* __restrict__ lo, int64_t off) {
  const ushort4 h = __ldg(reinterpret_cast<const ushort4*>(hi + off));
  const ushort4 l = __ldg(reinterpret_cast<const ushort4*>(lo + off));
  return make_float4(join_f16x2(h.x, l.x), join_f16x2(h.y, l.y), join_f16x2(h.z, l.z), join_f16x2(h.w, l.w));
}

// fp64 column-major (rows x b, ld) -> fp32 row-major (rows x bp), zero padded
__global__ void to_f32_rowmajor(int64_t rows, int b, int bp, const double* __restrict__ src, int64_t ld,
                                float* __restrict__ dst) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * bp) return;
  const int64_t r = idx / bp;
  const int k = (int)(idx % bp);
  dst[idx] = (k < b) ? (float)src[(int64_t)k * ld + r] : 0.0f;
}

static void convert_operand(Ctx& ctx, ProdWs& ws, int64_t rows, int b, int bp, const double* src, int64_t ld) {
  if (ws.op32.n < rows * bp) ws.op32.alloc(rows * bp);
  ProfScope ps(ctx.prof, SCGBM_K_MISC, 1);
  to_f32_rowmajor<<<(unsigned)ceil_div(rows * bp, 256), 256, 0, ctx.stream>>>(rows, b, bp, src, ld, ws.op32.get());
  SCGBM_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------
// prod_n: part[chunk][k][i] = sum_{j in chunk} R[i,j] P[j,k]
// CTA tile: TI genes x BK columns; thread: 4 consecutive genes x 8 columns.
// ---------------------------------------------------------------------------------
template <int BK>
__global__ void __launch_bounds__(PR_NT, 2)
prod_n_kernel(const uint16_t* __restrict__ Rhi, const uint16_t* __restrict__ Rlo, int64_t ldI, int64_t J,
              const float* __restrict__ P, int bp,
              int k0, int64_t cells_per_chunk, double* __restrict__ part) {
  constexpr int KT = BK / 8;            // column threads
  constexpr int GT = PR_NT / KT;        // gene threads
  constexpr int TI = GT * 4;
  constexpr int KC = 16;                // cells per smem step
  __shared__ __align__(16) float As[KC][TI];
  __shared__ __align__(16) float Ps[KC][BK];
  const int gt = threadIdx.x % GT, kt = threadIdx.x / GT;
  const int64_t i0 = (int64_t)blockIdx.x * TI;
  const int chunk = blockIdx.y;
  const int64_t j_begin = (int64_t)chunk * cells_per_chunk;
  const int64_t j_end = min(J, j_begin + cells_per_chunk);
  float acc[4][8];
  double acc64[4][8];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) { acc[a][b] = 0.0f; acc64[a][b] = 0.0; }
  int since_flush = 0;
  for (int64_t j0 = j_begin; j0 < j_end; j0 += KC) {
    __syncthreads();
    // R tile: KC columns x TI genes, float4 along genes
    for (int idx = threadIdx.x; idx < KC * (TI / 4); idx += PR_NT) {
      const int jj = idx / (TI / 4), q = idx % (TI / 4);
      const int64_t j = j0 + jj, i = i0 + 4 * q;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j < j_end && i < ldI) v = load_resid4(Rhi, Rlo, j * ldI + i);
      *reinterpret_cast<float4*>(&As[jj][4 * q]) = v;
    }
    for (int idx = threadIdx.x; idx < KC * (BK / 4); idx += PR_NT) {
      const int jj = idx / (BK / 4), q = idx % (BK / 4);
      const int64_t j = j0 + jj;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j < j_end) v = __ldg(reinterpret_cast<const float4*>(P + j * bp + k0 + 4 * q));
      *reinterpret_cast<float4*>(&Ps[jj][4 * q]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int jj = 0; jj < KC; ++jj) {
      const float4 u = *reinterpret_cast<const float4*>(&As[jj][4 * gt]);
      const float4 p0 = *reinterpret_cast<const float4*>(&Ps[jj][8 * kt]);
      const float4 p1 = *reinterpret_cast<const float4*>(&Ps[jj][8 * kt + 4]);
      const float uu[4] = {u.x, u.y, u.z, u.w};
      const float pp[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(uu[a], pp[b], acc[a][b]);
    }
    if (++since_flush == 8) {
      since_flush = 0;
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) { acc64[a][b] += (double)acc[a][b]; acc[a][b] = 0.0f; }
    }
  }
  const int64_t i = i0 + 4 * gt;
  if (i < ldI) {
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      double* dst = part + ((int64_t)chunk * bp + k0 + 8 * kt + b) * ldI + i;
#pragma unroll
      for (int a = 0; a < 4; ++a) dst[a] = acc64[a][b] + (double)acc[a][b];
    }
  }
}

__global__ void prod_n_reduce(int nchunk, int64_t ldI, int b, int bp, const double* __restrict__ part,
                              double c, const float* __restrict__ c_dev, double* __restrict__ out, int64_t ldo) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ldI * b) return;
  const int64_t i = idx % ldI;
  const int k = (int)(idx / ldI);
  double s = 0.0;
  for (int ch = 0; ch < nchunk; ++ch) s += part[((int64_t)ch * bp + k) * ldI + i];
  out[(int64_t)k * ldo + i] += (c_dev ? c * (double)__ldg(c_dev) : c) * s;
}

void prod_n_reduce_launch(Ctx& ctx, int nchunk, int64_t ldI, int b, int bp, const double* part, double c,
                          const float* c_dev, double* out, int64_t ldo) {
  ProfScope ps(ctx.prof, SCGBM_K_MISC, 1);
  prod_n_reduce<<<(unsigned)ceil_div(ldI * b, 256), 256, 0, ctx.stream>>>(nchunk, ldI, b, bp, part, c, c_dev, out, ldo);
  SCGBM_LAUNCH_CHECK();
}

static bool use_tc(int variant, int b) {
  if (variant == PROD_FFMA) return false;
  if (variant == PROD_TCGEN05) { SCGBM_REQUIRE(prod_tc_supported(b), "tcgen05 products need a block of at most 64 columns"); return true; }
  static const bool force_ffma = getenv("SCGBM_PROD_FFMA") != nullptr;
  return !force_ffma && prod_tc_supported(b);
}

bool prod_uses_tc(int b) { return use_tc(PROD_AUTO, b); }

void prod_n(Ctx& ctx, ProdWs& ws, ResidBuf R, int64_t I, int64_t ldI, int64_t J, const double* P,
            int64_t ldP, int b, double c, double* out, int64_t ldo, int variant, bool use_lo) {
  (void)I;
  c /= (double)R.scale;   // the planes hold scale * R
  if (use_tc(variant, b)) { prod_n_tc(ctx, ws, R, ldI, J, P, ldP, b, c, out, ldo, use_lo); return; }
  const int BK = b <= 32 ? 32 : 64;
  const int bp = (int)round_up(b, BK);
  convert_operand(ctx, ws, J, b, bp, P, ldP);
  const int TI = (PR_NT / (BK / 8)) * 4;
  const int nIt = (int)ceil_div(ldI, TI);
  int nchunk = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div((int64_t)ctx.num_sms * 4, nIt), ceil_div(J, 256)));
  int64_t cpc = round_up(ceil_div(J, nchunk), 16);
  nchunk = (int)ceil_div(J, cpc);
  if (ws.part.n < (int64_t)nchunk * bp * ldI) ws.part.alloc((int64_t)nchunk * bp * ldI);
  {
    ProfScope ps(ctx.prof, SCGBM_K_PROD_N, bp / BK);
    dim3 grid(nIt, nchunk);
    for (int k0 = 0; k0 < bp; k0 += BK) {
      if (BK == 32) prod_n_kernel<32><<<grid, PR_NT, 0, ctx.stream>>>(R.hi, R.lo, ldI, J, ws.op32.get(), bp, k0, cpc, ws.part.get());
      else prod_n_kernel<64><<<grid, PR_NT, 0, ctx.stream>>>(R.hi, R.lo, ldI, J, ws.op32.get(), bp, k0, cpc, ws.part.get());
      SCGBM_LAUNCH_CHECK();
    }
  }
  prod_n_reduce_launch(ctx, nchunk, ldI, b, bp, ws.part.get(), c, nullptr, out, ldo);
}

// ---------------------------------------------------------------------------------
// prod_t: out[j,k] += c * sum_i R[i,j] Q[i,k]
// CTA tile: TJ cells x BK columns, loops over all genes; thread: 4 cells x 8 columns.
// ---------------------------------------------------------------------------------
template <int BK>
__global__ void __launch_bounds__(PR_NT, 2)
prod_t_kernel(const uint16_t* __restrict__ Rhi, const uint16_t* __restrict__ Rlo, int64_t ldI, int64_t J,
              const float* __restrict__ Q, int bp,
              int k0, int b, double c, double* __restrict__ out, int64_t ldo) {
  constexpr int KT = BK / 8;
  constexpr int CT = PR_NT / KT;        // cell threads
  constexpr int TJ = CT * 4;
  constexpr int KC = 32;                // genes per smem step
  __shared__ float As[TJ][KC + 1];
  __shared__ __align__(16) float Qs[KC][BK];
  const int ct = threadIdx.x % CT, kt = threadIdx.x / CT;
  const int64_t j0 = (int64_t)blockIdx.x * TJ;
  float acc[4][8];
  double acc64[4][8];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int bb = 0; bb < 8; ++bb) { acc[a][bb] = 0.0f; acc64[a][bb] = 0.0; }
  int since_flush = 0;
  for (int64_t i0 = 0; i0 < ldI; i0 += KC) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < TJ * (KC / 4); idx += PR_NT) {
      const int cell = idx / (KC / 4), q = idx % (KC / 4);
      const int64_t j = j0 + cell;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j < J) v = load_resid4(Rhi, Rlo, j * ldI + i0 + 4 * q);
      As[cell][4 * q + 0] = v.x; As[cell][4 * q + 1] = v.y;
      As[cell][4 * q + 2] = v.z; As[cell][4 * q + 3] = v.w;
    }
    for (int idx = threadIdx.x; idx < KC * (BK / 4); idx += PR_NT) {
      const int g = idx / (BK / 4), q = idx % (BK / 4);
      *reinterpret_cast<float4*>(&Qs[g][4 * q]) =
          __ldg(reinterpret_cast<const float4*>(Q + (i0 + g) * bp + k0 + 4 * q));
    }
    __syncthreads();
#pragma unroll 8
    for (int g = 0; g < KC; ++g) {
      float rr[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) rr[a] = As[ct + CT * a][g];
      const float4 q0 = *reinterpret_cast<const float4*>(&Qs[g][8 * kt]);
      const float4 q1 = *reinterpret_cast<const float4*>(&Qs[g][8 * kt + 4]);
      const float qq[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int bb = 0; bb < 8; ++bb) acc[a][bb] = fmaf(rr[a], qq[bb], acc[a][bb]);
    }
    if (++since_flush == 4) {
      since_flush = 0;
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int bb = 0; bb < 8; ++bb) { acc64[a][bb] += (double)acc[a][bb]; acc[a][bb] = 0.0f; }
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int64_t j = j0 + ct + CT * a;
    if (j >= J) continue;
#pragma unroll
    for (int bb = 0; bb < 8; ++bb) {
      const int k = k0 + 8 * kt + bb;
      if (k < b) out[(int64_t)k * ldo + j] += c * (acc64[a][bb] + (double)acc[a][bb]);
    }
  }
}

void prod_t(Ctx& ctx, ProdWs& ws, ResidBuf R, int64_t I, int64_t ldI, int64_t J, const double* Q,
            int64_t ldQ, int b, double c, double* out, int64_t ldo, int variant, bool use_lo) {
  (void)I;
  c /= (double)R.scale;
  if (use_tc(variant, b)) { prod_t_tc(ctx, ws, R, ldI, J, Q, ldQ, b, c, out, ldo, use_lo); return; }
  const int BK = b <= 32 ? 32 : 64;
  const int bp = (int)round_up(b, BK);
  convert_operand(ctx, ws, ldI, b, bp, Q, ldQ);   // Q has ldI rows (pad rows are zero)
  const int TJ = (PR_NT / (BK / 8)) * 4;
  ProfScope ps(ctx.prof, SCGBM_K_PROD_T, bp / BK);
  dim3 grid((unsigned)ceil_div(J, TJ));
  for (int k0 = 0; k0 < bp; k0 += BK) {
    if (BK == 32) prod_t_kernel<32><<<grid, PR_NT, 0, ctx.stream>>>(R.hi, R.lo, ldI, J, ws.op32.get(), bp, k0, b, c, out, ldo);
    else prod_t_kernel<64><<<grid, PR_NT, 0, ctx.stream>>>(R.hi, R.lo, ldI, J, ws.op32.get(), bp, k0, b, c, out, ldo);
    SCGBM_LAUNCH_CHECK();
  }
}

}  // namespace scgbm
