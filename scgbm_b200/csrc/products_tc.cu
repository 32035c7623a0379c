// products_tc.cu -- K5 on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
//   prod_n : part[chunk][k][i] = sum_{j in chunk} R[i,j] P[j,k]     (A = R tile, MN-major: genes contiguous)
//   prod_t : out[j,k]        += c * sum_i R[i,j] Q[i,k]             (A = R^T tile, K-major: genes contiguous)
//
// One persistent warp-specialised kernel serves both: warp 0 = TMA producer, warp 1 = MMA
// issuer (one elected thread) and TMEM owner, warps 2-5 = epilogue (one TMEM lane quarter
// each).  Per 64-step k-block the producer lands the two fp16 planes of R (2 x 16 KB) and the
// two fp16 planes of the scaled thin operand (2*BP rows x 128 B) in 128B-swizzled shared
// memory; the issuer runs 4 x (hi: N = 2*BP, lo: N = BP) UMMAs of K = 16 into an fp32 TMEM
// accumulator (kind::f16 requires A and B in the same 16-bit format -- fp16 x bf16 raises an
// illegal-instruction fault -- hence fp16 everywhere, 11-bit pieces whose products are exact
// in fp32); every TC_DRAIN k-blocks the accumulator is handed to the epilogue warps, which
// fold the two column groups and add them into fp64 registers while the issuer continues in
// the second TMEM buffer.  HBM-bound by design: 4 bytes of R per entry against 6*BP flop.
#include "products.cuh"
#include "tc_common.cuh"

namespace scgbm {

constexpr int TC_NT = 192;
constexpr int TC_KB = 64;                        // reduction steps per k-block (128 B of bf16)
constexpr int TC_DRAIN = 2;                      // k-blocks per TMEM accumulation run (128 steps: keeps the
                                                 // tensor core's truncating fp32 accumulation below the fp32 FFMA error)
constexpr int TC_PLANE_BYTES = 128 * TC_KB * 2;  // one bf16 plane of a 128 x 64 tile

template <int BP> struct TcCfg {
  static constexpr int B_BYTES = 2 * BP * 128;
  static constexpr int STAGE_BYTES = 2 * TC_PLANE_BYTES + B_BYTES;
  static constexpr int STAGES = (BP <= 32) ? 5 : 4;
  static constexpr int TMEM_STRIDE = (BP <= 32) ? 64 : 128;    // columns per accumulator buffer
  static constexpr int TMEM_COLS = 2 * TMEM_STRIDE;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

struct TcParams {
  int64_t ldI, J;
  int n_tiles, n_gene_tiles;
  int64_t cells_per_chunk;   // N mode
  int kb_t;                  // T mode: k-blocks per tile = ldI / 64
  double c;                  // T mode
  const float* op_inv_scale; // device: 1 / (power-of-two scale of the thin operand planes)
  double* out;               // N: part[nchunk][BP][ldI];  T: out[J x b] (ldo)
  int64_t ldo;
  int b;
  uint32_t a_lbo, a_sbo;     // N-mode A descriptor strides (bytes)
  int use_lo;                // 0: hi plane only (basis-building passes; 11-bit residual)
};

template <int BP, bool TRANS>
__global__ void __launch_bounds__(TC_NT, 1)
prod_tc_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
               const __grid_constant__ CUtensorMap map_b, const TcParams p) {
  using Cfg = TcCfg<BP>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;                       // [STAGES]  TMA -> MMA
  uint64_t* empty = bars + Cfg::STAGES;        // [STAGES]  MMA -> TMA
  uint64_t* tfull = bars + 2 * Cfg::STAGES;    // [2]       MMA -> epilogue
  uint64_t* tempty = tfull + 2;                // [2]       epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&map_hi); tc::tma_prefetch_desc(&map_lo); tc::tma_prefetch_desc(&map_b);
    for (int s = 0; s < Cfg::STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { tc::mbar_init(&tfull[s], 1); tc::mbar_init(&tempty[s], 128); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  auto tile_kb = [&](int tile, int64_t& c0) -> int {
    if (TRANS) { c0 = (int64_t)tile * 128; return p.kb_t; }
    const int ci = tile / p.n_gene_tiles;
    c0 = (int64_t)ci * p.cells_per_chunk;
    const int64_t c1 = min(p.J, c0 + p.cells_per_chunk);
    return (int)((c1 - c0 + TC_KB - 1) / TC_KB);
  };

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        int64_t c0;
        const int nkb = tile_kb(tile, c0);
        const int gene0 = TRANS ? 0 : (tile % p.n_gene_tiles) * 128;
        for (int kb = 0; kb < nkb; ++kb) {
          tc::mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* st = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* a_hi = st; uint8_t* a_lo = st + TC_PLANE_BYTES; uint8_t* bt = st + 2 * TC_PLANE_BYTES;
          tc::mbar_arrive_expect_tx(&full[stage], Cfg::STAGE_BYTES - (p.use_lo ? 0 : TC_PLANE_BYTES));
          if (!TRANS) {
            const int cell = (int)(c0 + (int64_t)kb * TC_KB);
            tc::tma_load_2d(&map_hi, &full[stage], a_hi, gene0, cell);
            tc::tma_load_2d(&map_hi, &full[stage], a_hi + TC_PLANE_BYTES / 2, gene0 + 64, cell);
            if (p.use_lo) {
              tc::tma_load_2d(&map_lo, &full[stage], a_lo, gene0, cell);
              tc::tma_load_2d(&map_lo, &full[stage], a_lo + TC_PLANE_BYTES / 2, gene0 + 64, cell);
            }
            tc::tma_load_2d(&map_b, &full[stage], bt, cell, 0);
          } else {
            const int gene = kb * TC_KB;
            const int cell0 = (int)c0;
            tc::tma_load_2d(&map_hi, &full[stage], a_hi, gene, cell0);
            tc::tma_load_2d(&map_hi, &full[stage], a_hi + TC_PLANE_BYTES / 2, gene, cell0 + 64);
            if (p.use_lo) {
              tc::tma_load_2d(&map_lo, &full[stage], a_lo, gene, cell0);
              tc::tma_load_2d(&map_lo, &full[stage], a_lo + TC_PLANE_BYTES / 2, gene, cell0 + 64);
            }
            tc::tma_load_2d(&map_b, &full[stage], bt, gene, 0);
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      // A = fp16 residual planes, B = fp16 operand planes, D = fp32
      constexpr uint32_t idesc_hi = tc::instr_desc(0, 0, TRANS ? 0 : 1, 0, 128, 2 * BP);
      constexpr uint32_t idesc_lo = tc::instr_desc(0, 0, TRANS ? 0 : 1, 0, 128, BP);
      uint32_t stage = 0, phase = 0, buf = 0;
      uint32_t tphase[2] = {0, 0};
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        int64_t c0;
        const int nkb = tile_kb(tile, c0);
        for (int kb = 0; kb < nkb; ++kb) {
          const int kd = kb % TC_DRAIN;
          if (kd == 0) { tc::mbar_wait(&tempty[buf], tphase[buf] ^ 1); tc::fence_after_sync(); }
          tc::mbar_wait(&full[stage], phase);
          tc::fence_after_sync();
          const uint32_t st = tc::smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t a_hi = st, a_lo = st + TC_PLANE_BYTES, bt = st + 2 * TC_PLANE_BYTES;
          const uint32_t d = tmem_base + buf * Cfg::TMEM_STRIDE;
#pragma unroll
          for (int ks = 0; ks < TC_KB / 16; ++ks) {
            uint64_t da_hi, da_lo;
            if (TRANS) {   // K-major: 32 bytes per K = 16 inside the 128-byte swizzle row
              da_hi = tc::smem_desc_sw128(a_hi + ks * 32, 16, 1024);
              da_lo = tc::smem_desc_sw128(a_lo + ks * 32, 16, 1024);
            } else {       // MN-major: K = 16 is two 8-row groups of 1024 bytes
              da_hi = tc::smem_desc_sw128(a_hi + ks * 2048, p.a_lbo, p.a_sbo);
              da_lo = tc::smem_desc_sw128(a_lo + ks * 2048, p.a_lbo, p.a_sbo);
            }
            const uint64_t db = tc::smem_desc_sw128(bt + ks * 32, 16, 1024);
            tc::mma_f16(d, da_hi, db, idesc_hi, (kd > 0 || ks > 0) ? 1u : 0u);
            if (p.use_lo) tc::mma_f16(d, da_lo, db, idesc_lo, 1u);
          }
          tc::mma_commit(&empty[stage]);
          if (kd == TC_DRAIN - 1 || kb == nkb - 1) {
            tc::mma_commit(&tfull[buf]);
            tphase[buf] ^= 1;
            buf ^= 1;
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===== epilogue: TMEM -> fp64 registers -> global =====
    const int q = warp & 3;                       // TMEM lane quarter this warp may touch
    const int row = q * 32 + lane;                // D row = tile-local gene (N mode) or cell (T mode)
    double acc[BP];
#pragma unroll
    for (int k = 0; k < BP; ++k) acc[k] = 0.0;
    uint32_t buf = 0;
    uint32_t fphase[2] = {0, 0};
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      int64_t c0;
      const int nkb = tile_kb(tile, c0);
      const int ndrain = (nkb + TC_DRAIN - 1) / TC_DRAIN;
      for (int dr = 0; dr < ndrain; ++dr) {
        tc::mbar_wait(&tfull[buf], fphase[buf]);
        fphase[buf] ^= 1;
        tc::fence_after_sync();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * Cfg::TMEM_STRIDE;
#pragma unroll
        for (int g = 0; g < BP / 16; ++g) {
          uint32_t v0[16], v1[16];
          tc::tmem_ld16(taddr + g * 16, v0);
          tc::tmem_ld16(taddr + BP + g * 16, v1);
          tc::tmem_wait_ld();
#pragma unroll
          for (int k = 0; k < 16; ++k)
            acc[g * 16 + k] += (double)__uint_as_float(v0[k]) + (double)__uint_as_float(v1[k]);
        }
        tc::fence_before_sync();
        tc::mbar_arrive(&tempty[buf]);
        buf ^= 1;
      }
      if (!TRANS) {
        const int gi = tile % p.n_gene_tiles, ci = tile / p.n_gene_tiles;
        const int64_t gene = (int64_t)gi * 128 + row;
        if (gene < p.ldI) {
          double* dst = p.out + ((int64_t)ci * BP) * p.ldI + gene;
#pragma unroll
          for (int k = 0; k < BP; ++k)
            if (k < p.b) dst[(int64_t)k * p.ldI] = acc[k];
        }
      } else {
        const int64_t cell = c0 + row;
        if (cell < p.J) {
          const double cs = p.c * (double)__ldg(p.op_inv_scale);
#pragma unroll
          for (int k = 0; k < BP; ++k)
            if (k < p.b) p.out[(int64_t)k * p.ldo + cell] += cs * acc[k];
        }
      }
#pragma unroll
      for (int k = 0; k < BP; ++k) acc[k] = 0.0;
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) { tc::fence_after_sync(); tc::tmem_dealloc<Cfg::TMEM_COLS>(tmem_base); }
}

// ---------------------------------------------------------------------------------
// thin operand: fp64 column-major (rows x b, ld) -> two fp16 planes [2*bp][ldk] of
// s * x (reduction index contiguous, zero padded for k >= b, r >= rows), with s the power
// of two that brings max |x| to [8192, 16384): fp16 keeps 11 bits per plane down to 6e-5.
// ---------------------------------------------------------------------------------
__global__ void absmax_kernel(int64_t rows, int b, const double* __restrict__ src, int64_t ld, int* __restrict__ out) {
  float m = 0.0f;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < rows * b; idx += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx / rows);
    const int64_t r = idx % rows;
    m = fmaxf(m, fabsf((float)src[(int64_t)k * ld + r]));
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(out, __float_as_int(m));   // non-negative floats order like ints
}

__device__ __forceinline__ float operand_scale(const int* absmax_bits) {
  const float m = __int_as_float(*absmax_bits);
  if (!(m > 0.0f) || !isfinite(m)) return 1.0f;
  return exp2f(floorf(log2f(16384.0f / m)));
}

__global__ void split2_kernel(int64_t rows, int b, int bp, int64_t ldk, const double* __restrict__ src, int64_t ld,
                              const int* __restrict__ absmax_bits, uint16_t* __restrict__ dst,
                              float* __restrict__ inv_scale_out) {
  const float s = operand_scale(absmax_bits);
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx == 0) *inv_scale_out = 1.0f / s;
  if (idx >= (int64_t)bp * ldk) return;
  const int k = (int)(idx / ldk);
  const int64_t r = idx % ldk;
  float x = 0.0f;
  if (k < b && r < rows) x = (float)(src[(int64_t)k * ld + r] * (double)s);
  uint16_t p1, p2;
  split_f16x2(x, p1, p2);
  dst[((int64_t)k) * ldk + r] = p1;
  dst[((int64_t)(bp + k)) * ldk + r] = p2;
}

__global__ void split_planes_kernel(int64_t n, const float* __restrict__ src, float scale, uint16_t* __restrict__ hi,
                                    uint16_t* __restrict__ lo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint16_t h, l;
  split_f16x2(src[i] * scale, h, l);
  hi[i] = h; lo[i] = l;
}

void split_planes(Ctx& ctx, const float* src, int64_t n, ResidBuf dst) {
  if (n <= 0) return;
  ProfScope ps(ctx.prof, SCGBM_K_MISC, 1);
  split_planes_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, ctx.stream>>>(n, src, dst.scale, dst.hi, dst.lo);
  SCGBM_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

CUtensorMap make_tmap_2d(const void* base, int elem_bytes, bool is_float, uint64_t inner, uint64_t outer,
                         uint64_t outer_stride_bytes, uint32_t box_inner, uint32_t box_outer, int swizzle_bytes) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    SCGBM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    if (!p || qres != cudaDriverEntryPointSuccess) throw Error(SCGBM_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  CUtensorMapDataType dt;
  if (elem_bytes == 2) dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  else if (elem_bytes == 4) dt = is_float ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT32;
  else if (elem_bytes == 1) dt = CU_TENSOR_MAP_DATA_TYPE_UINT8;
  else throw Error(SCGBM_ERR_INVALID, "make_tmap_2d: unsupported element size");
  CUtensorMap m;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {outer_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(&m, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char b[256];
    snprintf(b, sizeof(b), "cuTensorMapEncodeTiled failed (%d): dims %llu x %llu, stride %llu, box %u x %u", (int)r,
             (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)outer_stride_bytes, box_inner, box_outer);
    throw Error(SCGBM_ERR_CUDA, b);
  }
  return m;
}

bool prod_tc_supported(int b) { return b >= 1 && b <= 64; }

static int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return s ? atoi(s) : dflt;
}

template <int BP, bool TRANS>
static void launch_tc(Ctx& ctx, const CUtensorMap& mh, const CUtensorMap& ml, const CUtensorMap& mb, const TcParams& p) {
  using Cfg = TcCfg<BP>;
  auto kern = prod_tc_kernel<BP, TRANS>;
  // a function attribute belongs to the DEVICE the call is made on: no process-wide "already set" flag (several
  // GPUs may be driven from this one process, multi.cu)
  SCGBM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  const int grid = std::min(p.n_tiles, ctx.num_sms);
  kern<<<grid, TC_NT, Cfg::SMEM_BYTES, ctx.stream>>>(mh, ml, mb, p);
  SCGBM_LAUNCH_CHECK();
}

void split_operand(Ctx& ctx, ProdWs& ws, int64_t rows, int b, int bp, int64_t ldk, const double* src, int64_t ld) {
  if (ws.op16.n < 2 * (int64_t)bp * ldk) ws.op16.alloc(2 * (int64_t)bp * ldk);
  if (ws.opmeta.n < 2) ws.opmeta.alloc(2);
  int* absmax = reinterpret_cast<int*>(ws.opmeta.get());
  float* inv_scale = ws.opmeta.get() + 1;
  ProfScope ps(ctx.prof, SCGBM_K_MISC, 2);
  SCGBM_CUDA(cudaMemsetAsync(absmax, 0, sizeof(int), ctx.stream));
  const unsigned g1 = (unsigned)std::min<int64_t>(ceil_div(rows * b, 256), 4096);
  absmax_kernel<<<g1, 256, 0, ctx.stream>>>(rows, b, src, ld, absmax);
  split2_kernel<<<(unsigned)ceil_div((int64_t)bp * ldk, 256), 256, 0, ctx.stream>>>(rows, b, bp, ldk, src, ld, absmax,
                                                                              ws.op16.get(), inv_scale);
  SCGBM_LAUNCH_CHECK();
}

void prod_n_tc(Ctx& ctx, ProdWs& ws, ResidBuf R, int64_t ldI, int64_t J, const double* P, int64_t ldP,
               int b, double c, double* out, int64_t ldo, bool use_lo) {
  SCGBM_REQUIRE(prod_tc_supported(b), "prod_n_tc: block too wide");
  const int bp = b <= 32 ? 32 : 64;
  const int64_t ldk = round_up(J, 64);
  split_operand(ctx, ws, J, b, bp, ldk, P, ldP);
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.ldI = ldI; p.J = J; p.b = b;
  p.n_gene_tiles = (int)ceil_div(ldI, 128);
  const int64_t run = (int64_t)TC_KB * TC_DRAIN;
  const int64_t want = std::max<int64_t>(1, std::min<int64_t>(ceil_div((int64_t)ctx.num_sms * 8, p.n_gene_tiles), ceil_div(J, run)));
  p.cells_per_chunk = round_up(ceil_div(J, want), run);
  const int nchunk = (int)ceil_div(J, p.cells_per_chunk);
  p.n_tiles = p.n_gene_tiles * nchunk;
  if (ws.part.n < (int64_t)nchunk * bp * ldI) ws.part.alloc((int64_t)nchunk * bp * ldI);
  p.out = ws.part.get(); p.ldo = ldI;
  p.use_lo = use_lo ? 1 : 0;
  p.a_lbo = (uint32_t)env_int("SCGBM_TC_LBO", 8192);
  p.a_sbo = (uint32_t)env_int("SCGBM_TC_SBO", 1024);
  const CUtensorMap mh = make_tmap_2d(R.hi, 2, false, (uint64_t)ldI, (uint64_t)J, (uint64_t)ldI * 2, 64, 64);
  const CUtensorMap ml = make_tmap_2d(R.lo, 2, false, (uint64_t)ldI, (uint64_t)J, (uint64_t)ldI * 2, 64, 64);
  const CUtensorMap mb = make_tmap_2d(ws.op16.get(), 2, false, (uint64_t)ldk, (uint64_t)(2 * bp), (uint64_t)ldk * 2, 64, 2 * bp);
  {
    ProfScope ps(ctx.prof, SCGBM_K_PROD_N, 1);
    if (bp == 32) launch_tc<32, false>(ctx, mh, ml, mb, p);
    else launch_tc<64, false>(ctx, mh, ml, mb, p);
  }
  prod_n_reduce_launch(ctx, nchunk, ldI, b, bp, ws.part.get(), c, ws.opmeta.get() + 1, out, ldo);
}

void prod_t_tc(Ctx& ctx, ProdWs& ws, ResidBuf R, int64_t ldI, int64_t J, const double* Q, int64_t ldQ,
               int b, double c, double* out, int64_t ldo, bool use_lo) {
  SCGBM_REQUIRE(prod_tc_supported(b), "prod_t_tc: block too wide");
  const int bp = b <= 32 ? 32 : 64;
  split_operand(ctx, ws, ldI, b, bp, ldI, Q, ldQ);
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.ldI = ldI; p.J = J; p.b = b;
  p.n_gene_tiles = 1;
  p.n_tiles = (int)ceil_div(J, 128);
  p.kb_t = (int)(ldI / TC_KB);
  p.c = c; p.out = out; p.ldo = ldo; p.op_inv_scale = ws.opmeta.get() + 1; p.use_lo = use_lo ? 1 : 0;
  const CUtensorMap mh = make_tmap_2d(R.hi, 2, false, (uint64_t)ldI, (uint64_t)J, (uint64_t)ldI * 2, 64, 64);
  const CUtensorMap ml = make_tmap_2d(R.lo, 2, false, (uint64_t)ldI, (uint64_t)J, (uint64_t)ldI * 2, 64, 64);
  const CUtensorMap mb = make_tmap_2d(ws.op16.get(), 2, false, (uint64_t)ldI, (uint64_t)(2 * bp), (uint64_t)ldI * 2, 64, 2 * bp);
  ProfScope ps(ctx.prof, SCGBM_K_PROD_T, 1);
  if (bp == 32) launch_tc<32, true>(ctx, mh, ml, mb, p);
  else launch_tc<64, true>(ctx, mh, ml, mb, p);
}

}  // namespace scgbm
