// eta_tc.cu -- K3/K4 on the 5th-generation tensor cores (u8/u16 counts, M <= 64, any number of batches).
//
// The eta tile X = A B^T (A: genes x M with d folded in, B: cells x M) comes out of tcgen05
// UMMAs instead of 2M FFMAs per entry.  Each factor is scaled by a power of two and split in
// two fp16 pieces (22 bits); the K dimension of the MMA is the concatenation of three
// sections of Mp = round_up(M, 16) columns
//         gene rows:  [ a_hi | a_hi | a_lo ]        cell rows:  [ b_hi | b_lo | b_hi ]
// so that one accumulation gives a_hi b_hi + a_hi b_lo + a_lo b_hi = a.b to 2^-22 (every
// fp16 x fp16 product is exact in fp32; alpha_i and beta_j stay OUT of the MMA so that the
// truncating fp32 accumulator only ever sees |x|-sized partial sums).
//
//   fused_tc_kernel  (K4, R/fit.R:136-151,321,323): D[128 cells x 64 genes] in TMEM; epilogue
//     thread = one cell (TMEM lane), 16 consecutive genes per tcgen05.ld chunk; Y tile arrives
//     by TMA (128B swizzle: a thread's 32-byte run per chunk is conflict-free), the two fp16
//     planes of scale*(Y - W) leave by TMA store from a double-buffered smem tile.  Twelve warps:
//     factor-operand TMA, UMMA issue, retire (residual TMA store), count-tile TMA, 8 x epilogue;
//     every hand-off is an mbarrier.  The PROD instantiation (product rider, see the kernel) also
//     multiplies each outgoing residual tile with the SVD's start block and adds four drain warps.
//     sum W, sum_nz Y*x, max W are reduced fp32 per chunk -> fp64 per thread -> CTA partial;
//     sum_nz Y log W = sum Y*x + sum_i alpha_i rs_i + sum_j beta_j cs_j (+ clip correction).
//   rowsum_tc_kernel (K3, R/fit.R:127-130): D[128 genes x 128 cells]; epilogue thread = one
//     gene, accumulates g_j exp(x_ij) over the cells; touches no I x J memory (SFU-bound).
//     With the factors swapped the same kernel gives the column sums of R/fit.R:133-135.
#include "eta_pass.cuh"
#include "tc_common.cuh"

namespace scgbm {

// ---------------------------------------------------------------------------------
// factor staging: fp64 (ld x M) [* d] [* rowscale] -> fp16 rows [ld_pad][Kp], three sections
// ---------------------------------------------------------------------------------
// out[0] = max |v| (float bits), out[2] = max over rows of the row's 2-norm (float bits): |x_ij| <= |a_i| |b_j|
// bounds eta from below for the exp-underflow guard of the fused pass (ll_const_kernel)
__global__ void factor_absmax_kernel(int64_t n_real, int64_t ld, int M, const double* __restrict__ U,
                                     const double* __restrict__ d, const float* __restrict__ rs, int* __restrict__ out) {
  float m = 0.0f, nm = 0.0f;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_real; r += (int64_t)gridDim.x * blockDim.x) {
    const double rsv = rs ? (double)rs[r] : 1.0;
    double n2 = 0.0;
    for (int k = 0; k < M; ++k) {
      const double v = U[(int64_t)k * ld + r] * (d ? d[k] : 1.0) * rsv;
      m = fmaxf(m, fabsf((float)v));
      n2 += v * v;
    }
    nm = fmaxf(nm, (float)sqrt(n2) * 1.0000002f);     // rounded up: this is a bound
  }
  m = warp_max(m); nm = warp_max(nm);
  if ((threadIdx.x & 31) == 0) {
    if (m > 0.0f) atomicMax(out, __float_as_int(m));
    if (nm > 0.0f) atomicMax(out + 2, __float_as_int(nm));
  }
}

// side 0 (genes): [hi | hi | lo];  side 1 (cells): [hi | lo | hi];  rows >= n_real and columns >= M are zero
__global__ void factor_stage16_kernel(int64_t n_real, int64_t rows_pad, int64_t ld, int M, int Mp, int Kp, int side,
                                      const double* __restrict__ U, const double* __restrict__ d,
                                      const float* __restrict__ rs, const int* __restrict__ absmax_bits,
                                      uint16_t* __restrict__ out, float* __restrict__ inv_scale_out) {
  const float mx = __int_as_float(*absmax_bits);
  const float s = (mx > 0.0f && isfinite(mx)) ? exp2f(floorf(log2f(16384.0f / mx))) : 1.0f;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx == 0) *inv_scale_out = 1.0f / s;
  if (idx >= rows_pad * Mp) return;
  const int64_t r = idx / Mp;
  const int k = (int)(idx % Mp);
  uint16_t hi = 0, lo = 0;
  if (r < n_real && k < M) {
    const double v = U[(int64_t)k * ld + r] * (d ? d[k] : 1.0) * (rs ? (double)rs[r] : 1.0);
    split_f16x2((float)(v * (double)s), hi, lo);
  }
  uint16_t* row = out + r * Kp;
  row[k] = hi;
  row[Mp + k] = side == 0 ? hi : lo;
  row[2 * Mp + k] = side == 0 ? lo : hi;
  if (k == 0)
    for (int z = 3 * Mp; z < Kp; ++z) row[z] = 0;
}

void stage_factor16(Ctx& ctx, Factor16& f, int64_t n_real, int64_t ld, int M, int side, const double* U,
                    const double* d, const float* rs) {
  f.M = M;
  f.Mp = (int)round_up(M, 16);
  f.Kp = (int)round_up(3 * f.Mp, 64);
  f.rows_pad = round_up(ld, 128);
  if (f.data.n < f.rows_pad * f.Kp) f.data.alloc(f.rows_pad * f.Kp);
  if (f.meta.n < 4) f.meta.alloc(4);      // [0] max |v| (bits), [1] 1 / scale, [2] max row 2-norm (bits)
  int* absmax = reinterpret_cast<int*>(f.meta.get());
  ProfScope ps(ctx.prof, SCGBM_K_MISC, 2);
  SCGBM_CUDA(cudaMemsetAsync(absmax, 0, 4 * sizeof(int), ctx.stream));
  factor_absmax_kernel<<<(unsigned)std::min<int64_t>(ceil_div(n_real, 256), 2048), 256, 0, ctx.stream>>>(n_real, ld, M, U, d, rs, absmax);
  factor_stage16_kernel<<<(unsigned)ceil_div(f.rows_pad * f.Mp, 256), 256, 0, ctx.stream>>>(
      n_real, f.rows_pad, ld, M, f.Mp, f.Kp, side, U, d, rs, absmax, f.data.get(), f.meta.get() + 1);
  SCGBM_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------
constexpr int FT_NT = 384;            // warp 0 TMA factor loads, 1 eta MMA, 2 retire (TMA stores [+ fused product]), 3 TMA Y loads, 4-11 epilogue
constexpr int FT_NT_PROD = 512;       // + warps 12-15: fp64 drain of the product accumulator
constexpr int FT_W_RETIRE = 2, FT_W_Y = 3, FT_W_EPI0 = 4, FT_W_DRAIN0 = 12;
constexpr int FT_OP_BYTES = 64 * 128; // fused product: one operand tile, [hi 32 | lo 32 block columns][64 genes] fp16
constexpr int FT_NB = 3;              // gene-operand ring depth
constexpr int FT_NB_PROD = 2;         // ... with the product rider: one stage goes to the operand ring
constexpr int FT_NOP = 4;             // product operand ring depth (tiles are requested 3 ahead)
constexpr int FT_EPI = 256;
constexpr int FT_CELLS = 128;         // MMA M  (TMEM lanes = cells)
constexpr int FT_GENES = 64;          // MMA N  (TMEM columns = genes)
constexpr int FT_SLAB_A = FT_CELLS * 128;   // bytes of one 64-half K slab of the cell operand
constexpr int FT_SLAB_B = FT_GENES * 128;
constexpr int FT_R_BYTES = 2 * FT_CELLS * 128;   // hi + lo planes of one tile (each 128 rows x 128 B)

struct FusedTcParams {
  int64_t I, J, ldI;
  int nslab, kk;                      // K slabs of 64 halves; UMMA K steps (3*Mp/16)
  int n_cell_tiles, n_gene_tiles, gene_tiles_per_unit, n_units;
  int y_stages, y_stage_bytes;
  const float* f_row; const float* g_col;        // exp(alpha_i), exp(beta_j)   (padded, fp32)
  const float* la_row; const float* lb_col;      // alpha_i, beta_j (fp32; clip correction only)
  const float* inv_a; const float* inv_b;        // device: 1 / factor scales
  float maxY, log_maxY, rscale;
  const int32_t* batch;               // J zero-based batch codes (MB only)
  const float* rowscale;              // ldI x nbatch X0 row scaling (MB + CLAMP only) or null
  double* scal_part;                  // [cta][4]
  // fused first Krylov pass (PROD): prod_out[j, k] = sum_i R[i, j] Q0[i, k] for this pass's residual
  int prod_b;                         // block columns (<= 32)
  double* prod_out; int64_t prod_ld;  // J x prod_b, column-major
  const float* op_inv;                // device: 1 / operand scale
  uint32_t ns_prod, ns_mma, ns_drain; // back-off of the polling single-lane warps / UMMA issuer (0: spin) / drain warps
  int dbg_no_store;                   // experiment (SCGBM_DBG_NO_STORE): keep everything but the residual-plane TMA stores --
                                      // with the rider this IS an "implicit-W" product R^T Q (eta tile rebuilt, exp, fp16
                                      // split in shared memory, second UMMA) that reads only Y; results are then wrong
};

// PROD: the tile that the epilogue just wrote to shared memory for the TMA store is, as it lies there
// (128 cells x 64 genes, K-major, 128B swizzle), the A operand of prod_t's UMMA.  The retire warp
// multiplies it with the matching 64-gene tile of the fp16-split start block Q0 (the previous SVD's left
// Ritz block, known before this pass) into a second TMEM accumulator; every two tiles four drain warps
// (12-15, one per TMEM lane quarter) fold that accumulator into fp64 registers (same 128-step drain as prod_tc_kernel).  The warm SVD's
// first pass G^T Q0 then needs no pass over the planes at all.  Requires one unit per cell tile.
template <typename YT, bool CLAMP, bool MB, bool PROD, bool PK>
__global__ void __launch_bounds__(PROD ? FT_NT_PROD : FT_NT, 1)
fused_tc_kernel(const __grid_constant__ CUtensorMap map_cells, const __grid_constant__ CUtensorMap map_genes,
                const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_hi,
                const __grid_constant__ CUtensorMap map_lo, const __grid_constant__ CUtensorMap map_op,
                const FusedTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_buf = smem;                                          // nslab * 16 KB (cells)
  uint8_t* b_buf = a_buf + p.nslab * FT_SLAB_A;                   // FT_NB * nslab * 8 KB (genes)
  constexpr int NB = PROD ? FT_NB_PROD : FT_NB;
  uint8_t* y_buf = b_buf + NB * p.nslab * FT_SLAB_B;              // y_stages * y_stage_bytes
  uint8_t* r_buf = y_buf + p.y_stages * p.y_stage_bytes;          // 2 * 32 KB
  uint8_t* op_buf = r_buf + 2 * FT_R_BYTES;                       // PROD: FT_NOP * 8 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(op_buf + (PROD ? FT_NOP * FT_OP_BYTES : 0));
  uint64_t* a_full = bars;        uint64_t* a_empty = bars + 1;
  uint64_t* b_full = bars + 2;    uint64_t* b_empty = bars + 5;   // [FT_NB]
  uint64_t* d_full = bars + 8;    uint64_t* d_empty = bars + 10;  // [2]
  uint64_t* y_full = bars + 12;   uint64_t* y_empty = bars + 16;  // [<=4]
  uint64_t* r_full = bars + 20;   uint64_t* r_free = bars + 22;   // [2] residual tile buffers: epilogue <-> store warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  uint64_t* op_full = bars + 26;  uint64_t* op_empty = bars + 30; // [FT_NOP] PROD: operand tiles, retire warp <-> tensor core
  uint64_t* p_full = bars + 34;   uint64_t* p_empty = bars + 36;  // [2] PROD: product accumulators, tensor core <-> drain warps
  double* red = reinterpret_cast<double*>(bars + 38);             // [8][4]
  float* f_ring = reinterpret_cast<float*>(bars + 70);            // [<=4][64] exp(alpha_i) of the gene tile, with each Y stage
  constexpr int TMEM_COLS = PROD ? 256 : 128;                     // eta x2 at 0/64; PROD: products x2 at 128/192

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&map_cells); tc::tma_prefetch_desc(&map_genes); tc::tma_prefetch_desc(&map_y);
    tc::tma_prefetch_desc(&map_hi); tc::tma_prefetch_desc(&map_lo);
    if (PROD) tc::tma_prefetch_desc(&map_op);
    tc::mbar_init(a_full, 1); tc::mbar_init(a_empty, 1);
    for (int s = 0; s < NB; ++s) { tc::mbar_init(&b_full[s], 1); tc::mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&d_full[s], 1); tc::mbar_init(&d_empty[s], FT_EPI);
      tc::mbar_init(&r_full[s], 8); tc::mbar_init(&r_free[s], PROD ? 2 : 1);   // PROD: store drained + product UMMAs done
      tc::mbar_init(&p_full[s], 1); tc::mbar_init(&p_empty[s], 128);
    }
    for (int s = 0; s < FT_NOP; ++s) { tc::mbar_init(&op_full[s], 1); tc::mbar_init(&op_empty[s], 1); }
    for (int s = 0; s < 4; ++s) { tc::mbar_init(&y_full[s], 1); tc::mbar_init(&y_empty[s], FT_EPI); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc<TMEM_COLS>(tmem_slot);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // PROD: 512 threads start with 128 registers each; the epilogue needs ~165, so the single-lane warps
  // (warpgroup 0) and the drain warps (warpgroup 3) hand theirs over.  Each role branch re-budgets on entry
  // so that ptxas allocates per branch.   4 x 128 x (56 + 176 + 176 + 104) = 65536
  // (ptxas still aims near the launch value and rematerialises in the epilogue, ~134 registers used.  The
  // alternative that keeps 384 threads -- the four role warps folding the accumulator while their roles
  // wait -- was measured slower: 4.3 vs 3.1 ms per pass at 100k cells, the drains delay the roles.)
#define FT_REGS_PRODUCER() do { if (PROD) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;\n"); } while (0)
#define FT_REGS_EPILOGUE() do { if (PROD) asm volatile("setmaxnreg.inc.sync.aligned.u32 176;\n"); } while (0)
#define FT_REGS_DRAIN() do { if (PROD) asm volatile("setmaxnreg.dec.sync.aligned.u32 104;\n"); } while (0)

  // unit -> (cell tile, first gene tile, gene tile count)
  auto unit_info = [&](int unit, int& ct, int& gt0, int& ngt) {
    const int chunks = (p.n_gene_tiles + p.gene_tiles_per_unit - 1) / p.gene_tiles_per_unit;
    ct = unit / chunks;
    gt0 = (unit % chunks) * p.gene_tiles_per_unit;
    ngt = min(p.gene_tiles_per_unit, p.n_gene_tiles - gt0);
  };

  if (warp < 4) {
    // ===== warpgroup 0: four roles, one elected lane each =====
    FT_REGS_PRODUCER();
    const bool lead = lane == 0;
    const bool act = lead;
    // role waits: ns == 0 spins (UMMA issuer), otherwise backs off so that the poll does not eat issue slots
    auto pwait = [&](uint64_t* bar, uint32_t parity, uint32_t ns, bool) {
      if (ns) tc::mbar_wait_relaxed(bar, parity, ns); else tc::mbar_wait(bar, parity);
    };

    if (warp == 0 && act) {
      // ===== TMA producer of the factor operands (L2-resident); g = -1 is the unit's cell operand =====
      uint32_t bs = 0, bph = 0, aph = 0;
      for (int unit = blockIdx.x; unit < p.n_units; unit += gridDim.x) {
        int ct, gt0, ngt;
        unit_info(unit, ct, gt0, ngt);
        for (int g = -1; g < ngt; ++g) {
          pwait(g < 0 ? a_empty : &b_empty[bs], (g < 0 ? aph : bph) ^ 1, p.ns_prod, true);
          if (g < 0) {
            if (lead) {
              tc::mbar_arrive_expect_tx(a_full, p.nslab * FT_SLAB_A);
              for (int s = 0; s < p.nslab; ++s) tc::tma_load_2d(&map_cells, a_full, a_buf + s * FT_SLAB_A, s * 64, ct * FT_CELLS);
            }
            aph ^= 1;
          } else {
            if (lead) {
              tc::mbar_arrive_expect_tx(&b_full[bs], p.nslab * FT_SLAB_B);
              for (int s = 0; s < p.nslab; ++s)
                tc::tma_load_2d(&map_genes, &b_full[bs], b_buf + (bs * p.nslab + s) * FT_SLAB_B, s * 64, (gt0 + g) * FT_GENES);
            }
            if (++bs == NB) { bs = 0; bph ^= 1; }
          }
        }
      }
    } else if (warp == FT_W_Y && act) {
      // ===== TMA producer of the count tiles (HBM): runs y_stages tiles ahead of the epilogue,
      // independent of the MMA pipeline (ncu: the epilogue spent 10 % of its samples waiting for Y
      // while these loads were queued behind the factor ring) =====
      uint32_t ys = 0, yph = 0;
      for (int unit = blockIdx.x; unit < p.n_units; unit += gridDim.x) {
        int ct, gt0, ngt;
        unit_info(unit, ct, gt0, ngt);
        for (int g = 0; g < ngt; ++g) {
          pwait(&y_empty[ys], yph ^ 1, p.ns_prod, true);
          if (lead) {
            // f_i rides along with the count tile (single batch): the 80 KB vector does not survive in the
            // ~14 KB of L1 left beside 214 KB of shared memory, and an L2 round trip per chunk stalls the epilogue
            tc::mbar_arrive_expect_tx(&y_full[ys], p.y_stage_bytes + (MB ? 0 : FT_GENES * 4));
            tc::tma_load_2d(&map_y, &y_full[ys], y_buf + ys * p.y_stage_bytes, (gt0 + g) * FT_GENES, ct * FT_CELLS);
            if (!MB) tc::bulk_load_1d(f_ring + ys * FT_GENES, p.f_row + (int64_t)(gt0 + g) * FT_GENES, FT_GENES * 4, &y_full[ys]);
          }
          if (++ys == (uint32_t)p.y_stages) { ys = 0; yph ^= 1; }
        }
      }
    } else if (warp == 1 && act) {
      // ===== eta UMMA issuer =====
      constexpr uint32_t idesc = tc::instr_desc(0, 0, 0, 0, FT_CELLS, FT_GENES);
      uint32_t bs = 0, bph = 0, db = 0, aph = 0;
      uint32_t dph = 0;
      for (int unit = blockIdx.x; unit < p.n_units; unit += gridDim.x) {
        int ct, gt0, ngt;
        unit_info(unit, ct, gt0, ngt);
        pwait(a_full, aph, p.ns_mma, false);
        aph ^= 1;
        for (int g = 0; g < ngt; ++g) {
          pwait(&d_empty[db], ((dph >> db) & 1u) ^ 1u, p.ns_mma, true);
          pwait(&b_full[bs], bph, p.ns_mma, false);
          if (lead) {
            tc::fence_after_sync();
            const uint32_t a0 = tc::smem_u32(a_buf), b0 = tc::smem_u32(b_buf + bs * p.nslab * FT_SLAB_B);
            const uint32_t d = tmem_base + db * FT_GENES;
            for (int k = 0; k < p.kk; ++k) {
              const uint64_t da = tc::smem_desc_sw128(a0 + (k >> 2) * FT_SLAB_A + (k & 3) * 32, 16, 1024);
              const uint64_t dbd = tc::smem_desc_sw128(b0 + (k >> 2) * FT_SLAB_B + (k & 3) * 32, 16, 1024);
              tc::mma_f16(d, da, dbd, idesc, k > 0 ? 1u : 0u);
            }
            tc::mma_commit(&b_empty[bs]);
            tc::mma_commit(&d_full[db]);
            if (g == ngt - 1) tc::mma_commit(a_empty);
          }
          dph ^= 1u << db; db ^= 1;
          if (++bs == NB) { bs = 0; bph ^= 1; }
        }
      }
    } else if (warp == FT_W_RETIRE && act) {
      // ===== retire warp: ships each finished residual tile (TMA store; no CTA-wide barrier on this path)
      // and, PROD, feeds the same shared-memory tile to the product UMMAs =====
      constexpr uint32_t idesc_hi = tc::instr_desc(0, 0, 0, 0, FT_CELLS, 64);   // R_hi x [Q0_hi | Q0_lo]
      constexpr uint32_t idesc_lo = tc::instr_desc(0, 0, 0, 0, FT_CELLS, 32);   // R_lo x Q0_hi
      uint32_t rb = 0, ntile = 0;
      uint32_t fph = 0;
      uint32_t pb = 0;
      uint32_t peph = 0;
      for (int unit = blockIdx.x; unit < p.n_units; unit += gridDim.x) {
        int ct, gt0, ngt;
        unit_info(unit, ct, gt0, ngt);
        auto load_op = [&](uint32_t n, int gtile) {   // operand tile of global tile count n -> stage n % FT_NOP
          const uint32_t st = n % FT_NOP;
          pwait(&op_empty[st], ((n / FT_NOP) & 1u) ^ 1u, p.ns_prod, false);   // released by this warp's own commit
          if (lead) {
            tc::mbar_arrive_expect_tx(&op_full[st], FT_OP_BYTES);
            tc::tma_load_2d(&map_op, &op_full[st], op_buf + st * FT_OP_BYTES, gtile * FT_GENES, 0);
          }
        };
        if (PROD)
          for (int a = 0; a < FT_NOP - 1 && a < ngt; ++a) load_op(ntile + a, gt0 + a);
        for (int g = 0; g < ngt; ++g, ++ntile) {
          const int gene0 = (gt0 + g) * FT_GENES;
          // FT_NOP - 1 tiles ahead: that stage's last reader was issued one tile ago
          if (PROD && g + FT_NOP - 1 < ngt) load_op(ntile + FT_NOP - 1, gt0 + g + FT_NOP - 1);
          pwait(&r_full[rb], (fph >> rb) & 1u, p.ns_prod, true);     // all 8 epilogue warps wrote (and proxy-fenced) their part
          fph ^= 1u << rb;
          uint8_t* rt = r_buf + rb * FT_R_BYTES;
          auto ship = [&]() {
            if (!p.dbg_no_store) {
              tc::tma_store_2d(&map_hi, rt, gene0, ct * FT_CELLS);
              tc::tma_store_2d(&map_lo, rt + FT_CELLS * 128, gene0, ct * FT_CELLS);
            }
            tc::tma_store_commit();
            tc::tma_store_wait_read<1>();                    // the previous tile's store has drained the other buffer
            if (ntile >= 1) tc::mbar_arrive(&r_free[rb ^ 1]);
          };
          // store first, product UMMAs second: issuing the UMMAs ahead of the store was measured slower
          // (26.3 vs 24.9 ms per pass at 1M cells) -- the store queue is the critical resource of this kernel
          if (lead) ship();
          if (PROD) {
            const uint32_t st = ntile % FT_NOP;
            if ((g & 1) == 0) { pwait(&p_empty[pb], ((peph >> pb) & 1u) ^ 1u, 0, true); peph ^= 1u << pb; }
            pwait(&op_full[st], (ntile / FT_NOP) & 1u, 0, false);
            if (lead) {
              tc::fence_after_sync();
              const uint32_t a_hi = tc::smem_u32(rt), a_lo = a_hi + FT_CELLS * 128, bt = tc::smem_u32(op_buf + st * FT_OP_BYTES);
              const uint32_t d = tmem_base + 128 + pb * 64;
#pragma unroll
              for (int ks = 0; ks < FT_GENES / 16; ++ks) {
                const uint64_t da_hi = tc::smem_desc_sw128(a_hi + ks * 32, 16, 1024);
                const uint64_t da_lo = tc::smem_desc_sw128(a_lo + ks * 32, 16, 1024);
                const uint64_t dbd = tc::smem_desc_sw128(bt + ks * 32, 16, 1024);
                tc::mma_f16(d, da_hi, dbd, idesc_hi, ((g & 1) || ks > 0) ? 1u : 0u);
                tc::mma_f16(d, da_lo, dbd, idesc_lo, 1u);
              }
              tc::mma_commit(&op_empty[st]);
              tc::mma_commit(&r_free[rb]);
              if ((g & 1) || g == ngt - 1) tc::mma_commit(&p_full[pb]);
            }
            if ((g & 1) || g == ngt - 1) pb ^= 1;
          }
          rb ^= 1;
        }
      }
      if (lead) tc::tma_store_wait<0>();
    }
  } else if (PROD && warp >= FT_W_DRAIN0) {
    // ===== PROD drain warps: thread = one cell; every pair of gene tiles the fp32 TMEM accumulator
    // ([R_hi Q_hi + R_lo Q_hi | R_hi Q_lo], 2 x 32 columns) is folded into 32 fp64 registers =====
    FT_REGS_DRAIN();
    const int q = warp & 3;
    const int row = q * 32 + lane;
    double acc[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) acc[k] = 0.0;
    uint32_t pbuf = 0;
    uint32_t pfph = 0;
    const double sc = (double)__ldg(p.op_inv) / (double)p.rscale;
    for (int unit = blockIdx.x; unit < p.n_units; unit += gridDim.x) {
      int ct, gt0, ngt;
      unit_info(unit, ct, gt0, ngt);
      const int npair = (ngt + 1) >> 1;
      for (int pr = 0; pr < npair; ++pr) {
        tc::mbar_wait_relaxed(&p_full[pbuf], (pfph >> pbuf) & 1u, p.ns_drain);
        pfph ^= 1u << pbuf;
        tc::fence_after_sync();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + 128 + pbuf * 64;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v0[16], v1[16];
          tc::tmem_ld16(taddr + h * 16, v0);
          tc::tmem_ld16(taddr + 32 + h * 16, v1);
          tc::tmem_wait_ld();
#pragma unroll
          for (int k = 0; k < 16; ++k) acc[h * 16 + k] += (double)__uint_as_float(v0[k]) + (double)__uint_as_float(v1[k]);
        }
        tc::fence_before_sync();
        tc::mbar_arrive(&p_empty[pbuf]);
        pbuf ^= 1;
      }
      const int64_t cell = (int64_t)ct * FT_CELLS + row;
      if (cell < p.J) {
        double* dst = p.prod_out + cell;
#pragma unroll
        for (int k = 0; k < 32; ++k)
          if (k < p.prod_b) dst[(int64_t)k * p.prod_ld] = acc[k] * sc;
      }
#pragma unroll
      for (int k = 0; k < 32; ++k) acc[k] = 0.0;
    }
  } else {
    // ===== epilogue: one cell per thread, 2 chunks of 16 genes per tile =====
    FT_REGS_EPILOGUE();
    const int e = warp - FT_W_EPI0;         // 0..7
    const int q = warp & 3;                 // TMEM lane quarter of this warp
    const int ch = e >> 2;                  // column half: genes ch*32 .. ch*32+31 of the tile
    const int row = q * 32 + lane;          // tile-local cell
    const int et = threadIdx.x - FT_W_EPI0 * 32;   // 0..255
    const float inv = __ldg(p.inv_a) * __ldg(p.inv_b);
    const float L2E = 1.4426950408889634f * inv;   // log2(e) / (sA sB): one rounding of the product, <= 2^-24 |t|
    const float clampv = 8.0f / inv;
    // everything W-sized is carried pre-multiplied by the residual scale 2^k (exact): the count
    // comes out of the magic-number conversion already scaled, g_j and max.Y are scaled once
    const int kexp = (int)((__float_as_uint(p.rscale) >> 23) & 0xffu) - 127;
    const uint32_t magic_bits = (uint32_t)(150 + kexp) << 23;       // 2^(23+k)
    const float magic = __uint_as_float(magic_bits);
    const float maxYs = p.maxY * p.rscale;
    const float inv_rs = 1.0f / p.rscale;
    double sw64 = 0.0, syl64 = 0.0;
    float mw = 0.0f;
    uint32_t ys = 0, yph = 0, db = 0, rb = 0;
    uint32_t dph = 0, rfph = 0;   // phase bits per buffer (bit masks: indexed arrays would live in local memory)
    const uint32_t y_s = tc::smem_u32(y_buf), r_s = tc::smem_u32(r_buf), f_s = tc::smem_u32(f_ring);
    for (int unit = blockIdx.x; unit < p.n_units; unit += gridDim.x) {
      int ct, gt0, ngt;
      unit_info(unit, ct, gt0, ngt);
      const int64_t cell = (int64_t)ct * FT_CELLS + row;
      const bool cell_ok = cell < p.J;
      const float gj = cell_ok ? __ldg(p.g_col + cell) * p.rscale : 0.0f;
      const float lbj = cell_ok ? __ldg(p.lb_col + cell) : 0.0f;
      // multi-batch: alpha (hence f_i, la_i and the X0 row scaling) depends on the batch of this thread's cell
      const int64_t boff = (MB && cell_ok) ? (int64_t)__ldg(p.batch + cell) * p.ldI : 0;
      for (int g = 0; g < ngt; ++g) {
        const int gene0 = (gt0 + g) * FT_GENES;
        tc::mbar_wait(&d_full[db], (dph >> db) & 1u);
        dph ^= 1u << db;
        tc::mbar_wait(&y_full[ys], yph);
        tc::fence_after_sync();
        // 32-bit shared addresses: LDS / STS instead of generic loads and stores (tc_common.cuh)
        const uint32_t yt = y_s + ys * (uint32_t)p.y_stage_bytes + (uint32_t)row * (sizeof(YT) == 2 ? 128u : 64u);
        const uint32_t rt = r_s + rb * (uint32_t)FT_R_BYTES + (uint32_t)row * 128u;
        const uint32_t ft = f_s + ys * (uint32_t)(FT_GENES * 4);
        // Without the rider: issue every load of the tile's two chunks first (TMEM, Y from smem, f_i broadcast),
        // wait once, hand the TMEM buffer back to the MMA warp, then do the arithmetic (168 registers).
        // With it the kernel launches 512 threads at 128 registers and ptxas keeps the epilogue near that
        // figure whatever setmaxnreg grants, rematerialising addresses all over the front-loaded form (+30 %
        // instructions); loading one chunk at a time needs ~40 registers less and measures 2.77 vs 3.14 ms
        // per pass at 114k cells (without the rider the front-loaded form stays ahead, 2.62 vs 2.72 ms).
        uint32_t xr2[2][16];
        uint4 yv[2][2];
        float4 fv[2][4];
        auto issue_loads = [&](int c) {
          const int col = ch * 32 + c * 16;            // first gene of the chunk within the tile
          tc::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + db * FT_GENES + col, xr2[c]);
          if (sizeof(YT) == 2) {
            const int q0 = col >> 3;                   // 16-byte chunk index (8 genes each)
            yv[c][0] = tc::lds128(yt + (uint32_t)(((q0) ^ (row & 7)) << 4));
            yv[c][1] = tc::lds128(yt + (uint32_t)(((q0 + 1) ^ (row & 7)) << 4));
          } else {
            yv[c][0] = tc::lds128(yt + (uint32_t)col);
            yv[c][1] = make_uint4(0u, 0u, 0u, 0u);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k)
            fv[c][k] = MB ? __ldg(reinterpret_cast<const float4*>(p.f_row + boff + gene0 + col) + k)
                          : tc::lds128f(ft + (uint32_t)((col + 4 * k) * 4));
        };
        constexpr bool SEQ = PROD;
        if constexpr (!SEQ) {
          issue_loads(0); issue_loads(1);
          tc::tmem_wait_ld();
          tc::fence_before_sync();
          tc::mbar_arrive(&d_empty[db]);               // accumulator is in registers: the next MMA may overwrite it
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int col = ch * 32 + c * 16;
          if constexpr (SEQ) {
            issue_loads(c);
            tc::tmem_wait_ld();
            if (c == 1) { tc::fence_before_sync(); tc::mbar_arrive(&d_empty[db]); }
          }
          uint32_t (&xr)[16] = xr2[c];
          uint32_t hp[8], lp[8];
          if constexpr (PK) {
            // ---- packed form: two entries per FADD2 / FMUL2 / FFMA2, ~9 issued instructions per entry ----
            // count -> float by a byte permute into the mantissa of 2^(23+k) (PRMT) and one packed subtraction;
            // w = ex2(x L2E) (f_i g_j); clipping (R/fit.R:145) is deferred to the rare chunk that needs it.
            const tc::f32x2 L2 = tc::pk2(L2E, L2E), G2 = tc::pk2(gj, gj), NM2 = tc::pk2(-magic, -magic);
            tc::f32x2 SW2 = tc::pk2(0.0f, 0.0f), SY2 = tc::pk2(0.0f, 0.0f);
            float cm = 0.0f;
            const uint32_t yw[8] = {yv[c][0].x, yv[c][0].y, yv[c][0].z, yv[c][0].w, yv[c][1].x, yv[c][1].y, yv[c][1].z, yv[c][1].w};
            const float fs[16] = {fv[c][0].x, fv[c][0].y, fv[c][0].z, fv[c][0].w, fv[c][1].x, fv[c][1].y, fv[c][1].z, fv[c][1].w,
                                  fv[c][2].x, fv[c][2].y, fv[c][2].z, fv[c][2].w, fv[c][3].x, fv[c][3].y, fv[c][3].z, fv[c][3].w};
            auto ybits = [&](int k) -> uint32_t {     // float bits of 2^(23+k) + count k of the chunk
              if (sizeof(YT) == 2) return __byte_perm(yw[k >> 1], magic_bits, (k & 1) ? 0x7632u : 0x7610u);
              return __byte_perm(yw[k >> 2], magic_bits, 0x7640u | (uint32_t)(k & 3));
            };
#pragma unroll
            for (int pp = 0; pp < 8; ++pp) {
              float x0 = __uint_as_float(xr[2 * pp]), x1 = __uint_as_float(xr[2 * pp + 1]);
              if (MB && CLAMP && p.rowscale) {                          // X0 with batches: per (gene, batch) scaling, R/fit.R:117
                x0 *= __ldg(p.rowscale + boff + gene0 + col + 2 * pp);
                x1 *= __ldg(p.rowscale + boff + gene0 + col + 2 * pp + 1);
              }
              if (CLAMP) {                                              // R/fit.R:119-120
                x0 = fminf(fmaxf(x0, -clampv), clampv); x1 = fminf(fmaxf(x1, -clampv), clampv);
              }
              if (CLAMP || MB) { xr[2 * pp] = __float_as_uint(x0); xr[2 * pp + 1] = __float_as_uint(x1); }
              const tc::f32x2 X2 = tc::pk2(x0, x1);
              const tc::f32x2 Y2 = tc::add2(tc::pk2(__uint_as_float(ybits(2 * pp)), __uint_as_float(ybits(2 * pp + 1))), NM2);
              float t0, t1;
              tc::upk2(tc::mul2(X2, L2), t0, t1);
              const tc::f32x2 W2 = tc::mul2(tc::pk2(ex2_approx(t0), ex2_approx(t1)), tc::mul2(tc::pk2(fs[2 * pp], fs[2 * pp + 1]), G2));
              float w0, w1;
              tc::upk2(W2, w0, w1);
              cm = tc::max3(cm, w0, w1);
              SW2 = tc::add2(SW2, W2);
              SY2 = tc::fma2(Y2, X2, SY2);
              float r0, r1;
              tc::upk2(tc::sub2(Y2, W2), r0, r1);                       // 2^k (Y - W)
              tc::split_f16x2_pair(r0, r1, hp[pp], lp[pp]);
            }
            float s0, s1, u0, u1;
            tc::upk2(SW2, s0, s1); tc::upk2(SY2, u0, u1);
            float tsw = s0 + s1;
            const float tsyl = u0 + u1;
            double corr = 0.0;
            if (cm > maxYs) {   // rare: some W of this chunk exceeds max.Y -- redo it entry by entry with the clip;
                                // clipped entries take log(max.Y) instead of eta in the objective (R/fit.R:145,151)
              tsw = 0.0f;
              float rr[16];
#pragma unroll
              for (int k = 0; k < 16; ++k) {
                const float xs = __uint_as_float(xr[k]);
                const float yk = __uint_as_float(ybits(k)) - magic;
                const float w = ex2_approx(xs * L2E) * (fs[k] * gj);    // same value as above
                const float wc = fminf(w, maxYs);
                tsw += wc;
                rr[k] = yk - wc;
                if (w > maxYs && yk != 0.0f) {
                  const float lw = xs * inv + (__ldg(p.la_row + boff + gene0 + col + k) + lbj);
                  corr += (double)(yk * inv_rs) * (double)(p.log_maxY - lw);
                }
              }
#pragma unroll
              for (int pp = 0; pp < 8; ++pp) tc::split_f16x2_pair(rr[2 * pp], rr[2 * pp + 1], hp[pp], lp[pp]);
            }
            sw64 += (double)tsw;
            syl64 += (double)tsyl * (double)(inv * inv_rs) + corr;
            mw = fmaxf(mw, fminf(cm, maxYs));
          } else {
          float y[16];
          if (sizeof(YT) == 2) {
            const uint32_t w[8] = {yv[c][0].x, yv[c][0].y, yv[c][0].z, yv[c][0].w, yv[c][1].x, yv[c][1].y, yv[c][1].z, yv[c][1].w};
#pragma unroll
            for (int k = 0; k < 8; ++k) {   // exact small-int -> float on the ALU/FMA pipes (the XU pipe is busy with ex2)
              y[2 * k] = __uint_as_float(magic_bits | (w[k] & 0xffffu)) - magic;        // = 2^k * count, exact
              y[2 * k + 1] = __uint_as_float(magic_bits | (w[k] >> 16)) - magic;
            }
          } else {
            const uint32_t w[4] = {yv[c][0].x, yv[c][0].y, yv[c][0].z, yv[c][0].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              y[4 * k] = __uint_as_float(magic_bits | (w[k] & 0xffu)) - magic;
              y[4 * k + 1] = __uint_as_float(magic_bits | ((w[k] >> 8) & 0xffu)) - magic;
              y[4 * k + 2] = __uint_as_float(magic_bits | ((w[k] >> 16) & 0xffu)) - magic;
              y[4 * k + 3] = __uint_as_float(magic_bits | (w[k] >> 24)) - magic;
            }
          }
          float f[16];
#pragma unroll
          for (int k = 0; k < 4; ++k) { f[4 * k] = fv[c][k].x; f[4 * k + 1] = fv[c][k].y; f[4 * k + 2] = fv[c][k].z; f[4 * k + 3] = fv[c][k].w; }
          float tsw = 0.0f, tsyl = 0.0f, cm = 0.0f;
          float r[16];
          if (MB && CLAMP && p.rowscale) {                           // X0 with batches: per (gene, batch) scaling, R/fit.R:117
#pragma unroll
            for (int k = 0; k < 16; ++k)
              xr[k] = __float_as_uint(__uint_as_float(xr[k]) * __ldg(p.rowscale + boff + gene0 + col + k));
          }
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            float xs = __uint_as_float(xr[k]);                        // sA sB x
            if (CLAMP) xs = fminf(fmaxf(xs, -clampv), clampv);        // R/fit.R:119-120
            xr[k] = __float_as_uint(xs);
            const float w = ex2_approx(xs * L2E) * (f[k] * gj);       // 2^k W
            cm = fmaxf(cm, w);
            // The clip costs one FMNMX per entry.  With the rider it is deferred to the rare chunk that needs
            // it (measured -1 %); without, the very same change measures +3 % (2.69 vs 2.62 ms at 114k cells:
            // ptxas schedules the 168-register form worse), so that instantiation keeps the inline clip.
            if constexpr (PROD) {
              tsw += w;                                               // clipping (R/fit.R:145) is handled below
              tsyl = fmaf(y[k], xs, tsyl);                            // sum Y x (scaled)
              r[k] = y[k] - w;                                        // 2^k (Y - W)
            } else {
              const float wc = fminf(w, maxYs);                       // R/fit.R:145
              tsw += wc;
              tsyl = fmaf(y[k], xs, tsyl);
              r[k] = y[k] - wc;
            }
          }
          double corr = 0.0;
          if (cm > maxYs) {   // rare: some W of this chunk exceeds max.Y; clipped entries take log(max.Y) instead
                              // of eta in the objective.  PROD: the fast path above did not clip -- redo the chunk
            if constexpr (PROD) tsw = 0.0f;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              const float xs = __uint_as_float(xr[k]);
              const float w = ex2_approx(xs * L2E) * (f[k] * gj);     // same value as above
              if constexpr (PROD) {
                const float wc = fminf(w, maxYs);
                tsw += wc;
                r[k] = y[k] - wc;
              }
              if (w > maxYs && y[k] != 0.0f) {
                const float lw = xs * inv + (__ldg(p.la_row + boff + gene0 + col + k) + lbj);
                corr += (double)(y[k] * inv_rs) * (double)(p.log_maxY - lw);
              }
            }
          }
          sw64 += (double)tsw;
          syl64 += (double)tsyl * (double)(inv * inv_rs) + corr;
          mw = fmaxf(mw, fminf(cm, maxYs));
          // fp16 planes, 16 genes = 32 bytes per plane: two 16-byte swizzled chunks
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const __half2 h2 = __floats2half2_rn(r[2 * k], r[2 * k + 1]);
            const float2 hf = __half22float2(h2);
            const __half2 l2 = __floats2half2_rn(r[2 * k] - hf.x, r[2 * k + 1] - hf.y);
            hp[k] = *reinterpret_cast<const uint32_t*>(&h2);
            lp[k] = *reinterpret_cast<const uint32_t*>(&l2);
          }
          }   // !PK
          if (c == 0) {   // the store that read r_buf[rb] two tiles ago must have drained it
            tc::mbar_wait(&r_free[rb], ((rfph >> rb) & 1u) ^ 1u);
            rfph ^= 1u << rb;
          }
          const int q0 = col >> 3;
          const uint32_t o0 = (uint32_t)(((q0) ^ (row & 7)) << 4), o1 = (uint32_t)(((q0 + 1) ^ (row & 7)) << 4);
          tc::sts128(rt + o0, hp[0], hp[1], hp[2], hp[3]);
          tc::sts128(rt + o1, hp[4], hp[5], hp[6], hp[7]);
          tc::sts128(rt + FT_CELLS * 128 + o0, lp[0], lp[1], lp[2], lp[3]);
          tc::sts128(rt + FT_CELLS * 128 + o1, lp[4], lp[5], lp[6], lp[7]);
        }
        // the Y stage is free again
        tc::mbar_arrive(&y_empty[ys]);
        // hand this warp's part of the tile to the store warp (generic-proxy writes -> async proxy)
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&r_full[rb]);
        db ^= 1; rb ^= 1;
        if (++ys == (uint32_t)p.y_stages) { ys = 0; yph ^= 1; }
      }
    }
    // CTA partial of the three scalars (fixed order)
    sw64 = warp_sum(sw64) * (double)inv_rs; syl64 = warp_sum(syl64); mw = warp_max(mw) * inv_rs;
    if (lane == 0) { red[e * 4 + 0] = sw64; red[e * 4 + 1] = syl64; red[e * 4 + 2] = (double)mw; }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (et == 0) {
      double a = 0.0, b = 0.0, c = 0.0;
      for (int w = 0; w < 8; ++w) { a += red[w * 4]; b += red[w * 4 + 1]; c = fmax(c, red[w * 4 + 2]); }
      p.scal_part[(int64_t)blockIdx.x * 4 + 0] = a;
      p.scal_part[(int64_t)blockIdx.x * 4 + 1] = b;
      p.scal_part[(int64_t)blockIdx.x * 4 + 2] = c;
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) { tc::fence_after_sync(); tc::tmem_dealloc<TMEM_COLS>(tmem_base); }
#undef FT_REGS_PRODUCER
#undef FT_REGS_EPILOGUE
#undef FT_REGS_DRAIN
}

// out[1] += sum_i alpha_i rs_i + sum_j beta_j cs_j   (single CTA, fp64, fixed order)
// out[3]  = 1 when exp(eta) could underflow to 0 in the reference's doubles for some entry: R then has
//           log(0) = -Inf for a non-zero count and takes the NaN/Inf branch (R/fit.R:152-157).  The tensor-core
//           pass sums Y * eta directly and would never see that, so the caller re-runs such an iteration on the
//           CUDA-core pass, which restates the underflow entry by entry (eta_pass.cu).  The test is a bound:
//           eta_ij >= min(alpha) + min(beta) - xb with |x_ij| <= xb = max_i |a_i|_2 * max_j |b_j|_2 (Cauchy-Schwarz
//           on the staged factor rows; 8 for the clamped X0).
__global__ void __launch_bounds__(1024)
ll_const_kernel(int64_t I, const double* __restrict__ alpha, const double* __restrict__ rs, int64_t J,
                const double* __restrict__ beta, const double* __restrict__ cs, double* __restrict__ out,
                int xM, int xclamp, const int* __restrict__ meta_a, const int* __restrict__ meta_b) {
  __shared__ double red[1024];
  __shared__ double rmin[1024];
  double s = 0.0, mn_a = 0.0, mn_b = 1e300;
  for (int64_t i = threadIdx.x; i < I; i += 1024) { const double a = alpha[i]; s += a * rs[i]; mn_a = fmin(mn_a, a); }
  for (int64_t j = threadIdx.x; j < J; j += 1024) { const double b = beta[j]; s += b * cs[j]; mn_b = fmin(mn_b, b); }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int st = 512; st > 0; st >>= 1) {
    if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[1] += red[0];
  // the two minima (fmin drops NaN: a NaN alpha/beta already makes the objective NaN through the sums)
  red[threadIdx.x] = mn_a; rmin[threadIdx.x] = mn_b;
  __syncthreads();
  for (int st = 512; st > 0; st >>= 1) {
    if (threadIdx.x < st) {
      red[threadIdx.x] = fmin(red[threadIdx.x], red[threadIdx.x + st]);
      rmin[threadIdx.x] = fmin(rmin[threadIdx.x], rmin[threadIdx.x + st]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double xb = 0.0;
    if (xM > 0) {
      xb = (double)__int_as_float(meta_a[2]) * (double)__int_as_float(meta_b[2]);
      if (xclamp) xb = fmin(xb, 8.0);
    }
    const double lo = red[0] + rmin[0] - xb;
    out[3] = (lo < -700.0 || lo != lo) ? 1.0 : 0.0;
  }
}

// ---------------------------------------------------------------------------------
// K3 row sums:  S_i = sum_j g_j exp(x_ij);   D[128 genes x 128 cells], thread = gene
// ---------------------------------------------------------------------------------
constexpr int RT_NT = 576;            // warp 0 TMA, warp 1 MMA, warps 2-17 epilogue (SFU-bound: keep the XU pipe fed)
constexpr int RT_EPI = 512;
constexpr int RT_TILE = 128;
constexpr int RT_SLAB = RT_TILE * 128;

struct RowsumTcParams {
  int64_t J, ldI;
  int nslab, kk;
  int n_gene_tiles, n_chunks, cell_tiles_per_chunk, n_cell_tiles;
  const float* g_col;   // exp(beta_j), zero padded to a multiple of 128 cells
  const float* lane_scale;  // per-lane multiplier of x before the clamp (X0 with batches) or null
  const float* inv_a; const float* inv_b;
  double* row_part;     // [chunk][4][ldI]
  int* row_flag;        // [ldI] or null: set when some |x| of the lane's row reaches flag_x (rows re-summed in fp64)
  float flag_x;
};

template <bool CLAMP>
__global__ void __launch_bounds__(RT_NT, 1)
rowsum_tc_kernel(const __grid_constant__ CUtensorMap map_genes, const __grid_constant__ CUtensorMap map_cells,
                 const RowsumTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_buf = smem;                                   // genes: nslab * 16 KB
  uint8_t* b_buf = a_buf + p.nslab * RT_SLAB;              // cells: 2 * nslab * 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_buf + 2 * p.nslab * RT_SLAB);
  uint64_t* a_full = bars;        uint64_t* a_empty = bars + 1;
  uint64_t* b_full = bars + 2;    uint64_t* b_empty = bars + 4;
  uint64_t* d_full = bars + 6;    uint64_t* d_empty = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&map_genes); tc::tma_prefetch_desc(&map_cells);
    tc::mbar_init(a_full, 1); tc::mbar_init(a_empty, 1);
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&b_full[s], 1); tc::mbar_init(&b_empty[s], 1);
      tc::mbar_init(&d_full[s], 1); tc::mbar_init(&d_empty[s], RT_EPI);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc<256>(tmem_slot);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  auto unit_info = [&](int unit, int& gt, int& ct0, int& nct) {
    gt = unit % p.n_gene_tiles;
    const int chunk = unit / p.n_gene_tiles;
    ct0 = chunk * p.cell_tiles_per_chunk;
    nct = min(p.cell_tiles_per_chunk, p.n_cell_tiles - ct0);
  };
  const int n_units = p.n_gene_tiles * p.n_chunks;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t bs = 0, bph = 0, aph = 0;
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        int gt, ct0, nct;
        unit_info(unit, gt, ct0, nct);
        tc::mbar_wait_relaxed(a_empty, aph ^ 1);
        tc::mbar_arrive_expect_tx(a_full, p.nslab * RT_SLAB);
        for (int s = 0; s < p.nslab; ++s) tc::tma_load_2d(&map_genes, a_full, a_buf + s * RT_SLAB, s * 64, gt * RT_TILE);
        aph ^= 1;
        for (int c = 0; c < nct; ++c) {
          tc::mbar_wait_relaxed(&b_empty[bs], bph ^ 1);
          tc::mbar_arrive_expect_tx(&b_full[bs], p.nslab * RT_SLAB);
          for (int s = 0; s < p.nslab; ++s)
            tc::tma_load_2d(&map_cells, &b_full[bs], b_buf + (bs * p.nslab + s) * RT_SLAB, s * 64, (ct0 + c) * RT_TILE);
          if (++bs == 2) { bs = 0; bph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = tc::instr_desc(0, 0, 0, 0, RT_TILE, RT_TILE);
      uint32_t bs = 0, bph = 0, db = 0, aph = 0;
      uint32_t dph = 0;
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        int gt, ct0, nct;
        unit_info(unit, gt, ct0, nct);
        tc::mbar_wait(a_full, aph);
        aph ^= 1;
        for (int c = 0; c < nct; ++c) {
          tc::mbar_wait(&d_empty[db], ((dph >> db) & 1u) ^ 1u);
          tc::mbar_wait(&b_full[bs], bph);
          tc::fence_after_sync();
          const uint32_t a0 = tc::smem_u32(a_buf), b0 = tc::smem_u32(b_buf + bs * p.nslab * RT_SLAB);
          const uint32_t d = tmem_base + db * RT_TILE;
          for (int k = 0; k < p.kk; ++k) {
            const uint64_t da = tc::smem_desc_sw128(a0 + (k >> 2) * RT_SLAB + (k & 3) * 32, 16, 1024);
            const uint64_t dbd = tc::smem_desc_sw128(b0 + (k >> 2) * RT_SLAB + (k & 3) * 32, 16, 1024);
            tc::mma_f16(d, da, dbd, idesc, k > 0 ? 1u : 0u);
          }
          tc::mma_commit(&b_empty[bs]);
          tc::mma_commit(&d_full[db]);
          if (c == nct - 1) tc::mma_commit(a_empty);
          dph ^= 1u << db; db ^= 1;
          if (++bs == 2) { bs = 0; bph ^= 1; }
        }
      }
    }
  } else {
    const int e = warp - 2, q = warp & 3, ch = e >> 2;   // ch: cells ch*32 .. +31 of the tile
    const int row = q * 32 + lane;                       // tile-local gene
    const float inv = __ldg(p.inv_a) * __ldg(p.inv_b);
    const float L_hi = 1.4426950216293335f * inv, L_lo = 1.9259629911e-08f * inv;
    const float clampv = 8.0f / inv;
    const float flagv = p.flag_x / inv;                  // in the scaled units of the accumulator
    uint32_t db = 0;
    uint32_t dph = 0;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
      int gt, ct0, nct;
      unit_info(unit, gt, ct0, nct);
      double s64 = 0.0;
      float xm = 0.0f;                                   // max |x| of this lane's row over the unit (ALU pipe; the kernel is MUFU-bound)
      float lsc = 1.0f;
      if (CLAMP && p.lane_scale) {
        const int64_t lr = (int64_t)gt * RT_TILE + row;
        lsc = lr < p.ldI ? __ldg(p.lane_scale + lr) : 0.0f;
      }
      for (int c = 0; c < nct; ++c) {
        const float* gp = p.g_col + (int64_t)(ct0 + c) * RT_TILE + ch * 32;
        tc::mbar_wait(&d_full[db], (dph >> db) & 1u);
        dph ^= 1u << db;
        tc::fence_after_sync();
        float ts = 0.0f;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int col = ch * 32 + cc * 16;
          uint32_t xr[16];
          tc::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + db * RT_TILE + col, xr);
          float gj[16];
#pragma unroll
          for (int k = 0; k < 4; ++k) {     // same address in every lane: one broadcast transaction
            const float4 g4 = __ldg(reinterpret_cast<const float4*>(gp + cc * 16) + k);
            gj[4 * k] = g4.x; gj[4 * k + 1] = g4.y; gj[4 * k + 2] = g4.z; gj[4 * k + 3] = g4.w;
          }
          tc::tmem_wait_ld();
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            float xs = __uint_as_float(xr[k]);
            if (CLAMP) xs = fminf(fmaxf(xs * lsc, -clampv), clampv);
            else xm = fmaxf(xm, fabsf(xs));
            ts = fmaf(ex2_approx(fmaf(xs, L_lo, xs * L_hi)), gj[k], ts);
          }
        }
        s64 += (double)ts;
        tc::fence_before_sync();
        tc::mbar_arrive(&d_empty[db]);
        db ^= 1;
      }
      // the four column quarters of a gene meet in the fixed-order second stage: [chunk][quarter][ldI]
      const int chunk = unit / p.n_gene_tiles;
      const int64_t gene = (int64_t)gt * RT_TILE + row;
      if (gene < p.ldI) {
        p.row_part[((int64_t)chunk * 4 + ch) * p.ldI + gene] = s64;
        if (!CLAMP && p.row_flag && xm >= flagv) p.row_flag[gene] = 1;    // idempotent: any unit may raise it
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) { tc::fence_after_sync(); tc::tmem_dealloc<256>(tmem_base); }
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
bool eta_tc_supported(const PassGeom& g, const XDesc& x, int y_storage) {
  static const bool off = getenv("SCGBM_ETA_FFMA") != nullptr;
  if (off || x.M < 1 || x.M > 64 || !x.A16 || !x.B16) return false;
  if (y_storage >= 0 && y_storage != SCGBM_STORE_U8 && y_storage != SCGBM_STORE_U16) return false;
  return true;
}

// the fused first Krylov pass needs: the unclamped low-rank X, enough cell tiles to be worth it, and room
// for the operand ring beside at least two count stages
bool fused_prod_supported(Ctx& ctx, const PassGeom& g, const XDesc& x, bool any_size) {
  const char* knob = getenv("SCGBM_FUSED_PROD");   // "0": off; "force": also below the size threshold (tests)
  if ((knob && knob[0] == '0' && !any_size) || x.clamp || !x.A16) return false;
  // one unit per cell tile (see pass_fused_tc): below ~4 tiles per SM the unfilled last wave eats the gain
  const int64_t n_cell_tiles = ceil_div(g.J, FT_CELLS);
  if (n_cell_tiles < (int64_t)ctx.num_sms * 4 && !(knob && knob[0] == 'f') && !any_size) return false;
  const int nslab = x.A16->Kp / 64;
  const int fixed = nslab * FT_SLAB_A + FT_NB_PROD * nslab * FT_SLAB_B + 2 * FT_R_BYTES + 1024 + 640 + 4 * FT_GENES * 4 + FT_NOP * FT_OP_BYTES;
  return (232448 - fixed) / (FT_CELLS * FT_GENES * 2) >= 2;
}

void pass_fused_tc(Ctx& ctx, EtaWs& ws, const PassGeom& g, const XDesc& x, const Counts& y, const float* f_row,
                   const float* g_col, const float* la_row, const float* lb_col, const double* alpha,
                   const double* beta, const double* rowsum_local, float maxY, ResidBuf resid, double* out3,
                   const FusedProd* prod) {
  const Factor16& fa = *x.A16;
  const Factor16& fb = *x.B16;
  FusedTcParams p;
  memset(&p, 0, sizeof(p));
  p.I = g.I; p.J = g.J; p.ldI = g.ldI;
  p.nslab = fa.Kp / 64; p.kk = 3 * fa.Mp / 16;
  p.n_cell_tiles = (int)ceil_div(g.J, FT_CELLS);
  p.n_gene_tiles = (int)ceil_div(g.ldI, FT_GENES);
  // units = cell tiles x gene chunks: ~16+ per SM, at least 8 gene tiles each; among the admissible
  // chunk counts take the one whose last wave is fullest (the units are equal-sized, so the tail of a
  // static round-robin is the only imbalance: ncu showed 7 % of the samples idling at the kernel's end)
  int chunks = 1;
  {
    const int cmin = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div((int64_t)ctx.num_sms * 16, p.n_cell_tiles), p.n_gene_tiles / 8));
    const int cmax = (int)std::max<int64_t>(cmin, std::min<int64_t>(p.n_gene_tiles / 8, 3 * cmin + 4));
    double best = -1.0;
    for (int c = cmin; c <= cmax; ++c) {
      const int per = (int)ceil_div(p.n_gene_tiles, c);
      const int64_t units = (int64_t)p.n_cell_tiles * ceil_div(p.n_gene_tiles, per);
      const double eff = (double)units / ((double)ctx.num_sms * (double)ceil_div(units, ctx.num_sms));
      if (eff > best + 1e-9) { best = eff; chunks = c; }
    }
  }
  if (getenv("SCGBM_FUSED_CHUNKS")) chunks = std::max(1, atoi(getenv("SCGBM_FUSED_CHUNKS")));   // experiments
  const bool do_prod = prod != nullptr;
  if (do_prod) {
    SCGBM_REQUIRE(fused_prod_supported(ctx, g, x, prod->product_only) && prod->b >= 1 && prod->b <= 32,
                  "fused_tc: product fusion not available for this shape");
    // One unit owns a cell tile's whole product row: fixed summation order, the very drain cadence of prod_t
    // (bit-identical results), no partial buffers.  Measured alternative -- gene chunks with per-chunk partials
    // summed afterwards -- fills the last wave better but pays the operand/accumulator ramp per unit:
    // 4.0 vs ~3.5 ms at 125k cells.
    chunks = 1;
  }
  p.gene_tiles_per_unit = (int)ceil_div(p.n_gene_tiles, chunks);
  chunks = (int)ceil_div(p.n_gene_tiles, p.gene_tiles_per_unit);
  p.n_units = p.n_cell_tiles * chunks;
  const int ybytes = y.elem_bytes();
  p.y_stage_bytes = FT_CELLS * FT_GENES * ybytes;
  const int fixed = p.nslab * FT_SLAB_A + (do_prod ? FT_NB_PROD : FT_NB) * p.nslab * FT_SLAB_B + 2 * FT_R_BYTES + 1024 + 640 +
                    4 * FT_GENES * 4 + (do_prod ? FT_NOP * FT_OP_BYTES : 0);   // 640: barriers + CTA reduction scratch
  p.y_stages = std::min(4, (int)(((do_prod ? 232448 : 225 * 1024) - fixed) / p.y_stage_bytes));
  SCGBM_REQUIRE(p.y_stages >= 2, "fused_tc: not enough shared memory for this M");
  const int smem_bytes = fixed + p.y_stages * p.y_stage_bytes;
  p.f_row = f_row; p.g_col = g_col; p.la_row = la_row; p.lb_col = lb_col;
  p.inv_a = fa.meta.get() + 1; p.inv_b = fb.meta.get() + 1;
  p.maxY = maxY; p.log_maxY = logf(maxY); p.rscale = resid.scale;
  p.batch = g.batch; p.rowscale = x.rowscale;
  // polling back-off (ns): measured flat between 64 and 400 for the single-lane warps, with or without a
  // back-off in the UMMA issuer; the drain warps poll once per ~3 us pair of tiles
  p.ns_prod = 64; p.ns_mma = 0; p.ns_drain = 400;
  static const bool no_store = getenv("SCGBM_DBG_NO_STORE") != nullptr;
  p.dbg_no_store = (no_store || (prod && prod->product_only)) ? 1 : 0;
  const int grid = std::min(p.n_units, ctx.num_sms);
  if (ws.scal_part.n < (int64_t)grid * 4) ws.scal_part.alloc((int64_t)grid * 4);
  p.scal_part = ws.scal_part.get();
  const CUtensorMap mc = make_tmap_2d(fb.data.get(), 2, false, (uint64_t)fb.Kp, (uint64_t)fb.rows_pad, (uint64_t)fb.Kp * 2, 64, FT_CELLS, 128);
  const CUtensorMap mg = make_tmap_2d(fa.data.get(), 2, false, (uint64_t)fa.Kp, (uint64_t)fa.rows_pad, (uint64_t)fa.Kp * 2, 64, FT_GENES, 128);
  const CUtensorMap my = make_tmap_2d(y.data, ybytes, false, (uint64_t)g.ldI, (uint64_t)g.J, (uint64_t)g.ldI * ybytes, FT_GENES, FT_CELLS,
                                      ybytes == 2 ? 128 : 0);
  const CUtensorMap mh = make_tmap_2d(resid.hi, 2, false, (uint64_t)g.ldI, (uint64_t)g.J, (uint64_t)g.ldI * 2, FT_GENES, FT_CELLS, 128);
  const CUtensorMap ml = make_tmap_2d(resid.lo, 2, false, (uint64_t)g.ldI, (uint64_t)g.J, (uint64_t)g.ldI * 2, FT_GENES, FT_CELLS, 128);
  CUtensorMap mo = ml;   // unused without PROD
  if (do_prod) {
    p.prod_b = prod->b; p.prod_out = prod->out; p.prod_ld = prod->ldo; p.op_inv = prod->op_inv;

    mo = make_tmap_2d(prod->op16, 2, false, (uint64_t)prod->ldk, 64, (uint64_t)prod->ldk * 2, 64, 64, 128);
  }
  {
    ProfScope ps(ctx.prof, SCGBM_K_FUSED, 1);
#define LAUNCH_FUSED_PK(YT, CL, MBF, PR, PKF)                                                                                       \
  do {                                                                                                                         \
    SCGBM_CUDA(cudaFuncSetAttribute(fused_tc_kernel<YT, CL, MBF, PR, PKF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes)); \
    fused_tc_kernel<YT, CL, MBF, PR, PKF><<<grid, PR ? FT_NT_PROD : FT_NT, smem_bytes, ctx.stream>>>(mc, mg, my, mh, ml, mo, p);   \
  } while (0)
#define LAUNCH_FUSED(YT, CL, MBF, PR)                                                                       \
  do { if (packed) LAUNCH_FUSED_PK(YT, CL, MBF, PR, true); else LAUNCH_FUSED_PK(YT, CL, MBF, PR, false); } while (0)
#define LAUNCH_FUSED_Y(YT)                                                                                   \
  do {                                                                                                       \
    if (do_prod) { if (mb) LAUNCH_FUSED(YT, false, true, true); else LAUNCH_FUSED(YT, false, false, true); } \
    else if (mb) { if (x.clamp) LAUNCH_FUSED(YT, true, true, false); else LAUNCH_FUSED(YT, false, true, false); } \
    else { if (x.clamp) LAUNCH_FUSED(YT, true, false, false); else LAUNCH_FUSED(YT, false, false, false); }  \
  } while (0)
    const bool mb = g.nbatch > 1;
    // packed-fp32 epilogue (FADD2/FMUL2/FFMA2, PRMT count unpack, FHADD split: 9 instead of ~20 issued instructions
    // per entry).  Measured at 20k x 1M: 23.5 vs 23.9 ms without the rider, 25.9-26.7 vs 25.3-26.4 ms with it (the
    // rider's instantiation is register-capped at 128 and the 64-bit pairs cost it moves) -- inside the box-to-box
    // noise either way: the pass is bound by the HBM read/write mix, not by issue slots.  The two forms round the
    // residual differently in the last bit (objective 2e-9 relative), so ONE form serves every instantiation:
    // a fit must not change in its low bits because a shard was large enough for the rider.  Default: the plain
    // form (the rider, i.e. the hot path at scale, is not slower with it); SCGBM_FUSED_PACKED=1 selects the packed
    // one for A/B measurements (read per launch, so a test can switch it).
    const char* pk_env = getenv("SCGBM_FUSED_PACKED");
    const bool packed = pk_env && pk_env[0] != '0';
    if (ybytes == 2) LAUNCH_FUSED_Y(uint16_t); else LAUNCH_FUSED_Y(uint8_t);
#undef LAUNCH_FUSED_Y
#undef LAUNCH_FUSED
#undef LAUNCH_FUSED_PK
    SCGBM_LAUNCH_CHECK();
  }
  {
    ProfScope ps(ctx.prof, SCGBM_K_MISC, 2);
    scal_reduce_launch(ctx, grid, ws.scal_part.get(), out3);
    // alpha and the local row sums are ldI x nbatch with zero pad rows: one flat dot product
    ll_const_kernel<<<1, 1024, 0, ctx.stream>>>(g.nbatch > 1 ? g.ldI * g.nbatch : g.I, alpha, rowsum_local, g.J, beta,
                                                y.colsum.get(), out3, x.M, x.clamp,
                                                reinterpret_cast<const int*>(fa.meta.get()),
                                                reinterpret_cast<const int*>(fb.meta.get()));
    SCGBM_LAUNCH_CHECK();
  }
}

// out[r] = sum_c colfac[c] * exp(x_rc) for the rows of `lanes` (TMEM lanes) against all rows of
// `streamed` (TMEM columns).  rowsum: lanes = genes, streamed = cells, colfac = exp(beta);
// colsum: lanes = cells, streamed = genes, colfac = exp(alpha).  colfac is zero padded to whole tiles.
static void lane_sums_tc(Ctx& ctx, EtaWs& ws, const Factor16& lanes, const Factor16& streamed, int64_t ld_lanes,
                         int64_t n_streamed, const float* colfac, bool clamp, double* out, int prof_class,
                         const float* lane_scale = nullptr, int* row_flag = nullptr) {
  RowsumTcParams p;
  memset(&p, 0, sizeof(p));
  p.J = n_streamed; p.ldI = ld_lanes;
  p.nslab = lanes.Kp / 64; p.kk = 3 * lanes.Mp / 16;
  p.n_gene_tiles = (int)ceil_div(ld_lanes, RT_TILE);
  p.n_cell_tiles = (int)ceil_div(n_streamed, RT_TILE);
  int chunks = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div((int64_t)ctx.num_sms * 8, p.n_gene_tiles), ceil_div(p.n_cell_tiles, 4)));
  p.cell_tiles_per_chunk = (int)ceil_div(p.n_cell_tiles, chunks);
  p.n_chunks = (int)ceil_div(p.n_cell_tiles, p.cell_tiles_per_chunk);
  p.g_col = colfac; p.lane_scale = lane_scale;
  p.row_flag = row_flag; p.flag_x = kExactRowX;
  p.inv_a = lanes.meta.get() + 1; p.inv_b = streamed.meta.get() + 1;
  const int64_t n = ld_lanes;
  if (ws.row_part.n < (int64_t)p.n_chunks * 4 * n) ws.row_part.alloc((int64_t)p.n_chunks * 4 * n);
  p.row_part = ws.row_part.get();
  const int smem_bytes = 3 * p.nslab * RT_SLAB + 1024 + 128;
  const int n_units = p.n_gene_tiles * p.n_chunks;
  const int grid = std::min(n_units, ctx.num_sms);
  const CUtensorMap ml = make_tmap_2d(lanes.data.get(), 2, false, (uint64_t)lanes.Kp, (uint64_t)lanes.rows_pad, (uint64_t)lanes.Kp * 2, 64, RT_TILE, 128);
  const CUtensorMap ms = make_tmap_2d(streamed.data.get(), 2, false, (uint64_t)streamed.Kp, (uint64_t)streamed.rows_pad, (uint64_t)streamed.Kp * 2, 64, RT_TILE, 128);
  {
    ProfScope ps(ctx.prof, prof_class, 1);
    if (clamp) {
      SCGBM_CUDA(cudaFuncSetAttribute(rowsum_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
      rowsum_tc_kernel<true><<<grid, RT_NT, smem_bytes, ctx.stream>>>(ml, ms, p);
    } else {
      SCGBM_CUDA(cudaFuncSetAttribute(rowsum_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
      rowsum_tc_kernel<false><<<grid, RT_NT, smem_bytes, ctx.stream>>>(ml, ms, p);
    }
    SCGBM_LAUNCH_CHECK();
  }
  rowpart_reduce_launch(ctx, p.n_chunks * 4, n, ws.row_part.get(), out);
}

__global__ void mask_batch_kernel(int64_t J, int64_t n_pad, const float* __restrict__ g, const int32_t* __restrict__ batch,
                                  int k, float* __restrict__ out) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n_pad) out[j] = (j < J && batch[j] == k) ? g[j] : 0.0f;
}

// ---- exact re-summation of the flagged genes (eta_pass.cuh kExactRowX) -------------------------------------
// S[i, k] = sum_{j in batch k} exp(beta_j + x_ij) in fp64, x_ij = sum_m U_im d_m V_jm from the fp64 factors.
// (1) one CTA compacts the flags into a gene list (ascending, so the work layout is deterministic);
// (2) CTAs (cell chunk, slot lane) accumulate fixed-order partial sums part[k][slot][chunk];
// (3) one thread per (slot, batch) adds the chunks in order and overwrites S.  No host round trip.
constexpr int XR_CAP = 2048;      // genes per call; further flagged genes keep their tensor-core sum
constexpr int XR_CHUNK = 4096;    // cells per partial sum
constexpr int XR_LANES = 32;      // gridDim.y: slots are strided over these

__global__ void __launch_bounds__(1024)
exact_rows_compact_kernel(int64_t I, const int* __restrict__ row_flag, int* __restrict__ list /* [0] = count, then genes */,
                          unsigned long long* __restrict__ total) {
  __shared__ int wsum[32];
  __shared__ int base;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (int64_t i0 = 0; i0 < I; i0 += 1024) {
    const int64_t i = i0 + threadIdx.x;
    const int f = (i < I && row_flag[i]) ? 1 : 0;
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) wsum[w] = __popc(bal);
    __syncthreads();
    int off = base;
    for (int q = 0; q < w; ++q) off += wsum[q];
    off += __popc(bal & ((1u << lane) - 1u));
    if (f && off < XR_CAP) list[1 + off] = (int)i;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int q = 0; q < 32; ++q) t += wsum[q]; base += t; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int n = base < XR_CAP ? base : XR_CAP;
    list[0] = n;
    atomicAdd(total, (unsigned long long)n);
  }
}

__global__ void __launch_bounds__(256)
exact_rows_partial_kernel(int64_t ldI, int64_t J, int64_t ldJ, int M, const double* __restrict__ U,
                          const double* __restrict__ d, const double* __restrict__ V, const double* __restrict__ beta,
                          const int32_t* __restrict__ batch, int nbatch, const int* __restrict__ list, int nchunks,
                          double* __restrict__ part /* [nbatch][XR_CAP][nchunks] */) {
  const int count = list[0];
  __shared__ double a_s[64];
  __shared__ double red[8];
  const int64_t j0 = (int64_t)blockIdx.x * XR_CHUNK;
  for (int slot = blockIdx.y; slot < count; slot += gridDim.y) {
    const int64_t i = list[1 + slot];
    __syncthreads();
    if (threadIdx.x < M) a_s[threadIdx.x] = U[(int64_t)threadIdx.x * ldI + i] * d[threadIdx.x];
    __syncthreads();
    for (int k = 0; k < nbatch; ++k) {
      double s = 0.0;
      for (int t = threadIdx.x; t < XR_CHUNK; t += 256) {      // fixed cell -> thread assignment
        const int64_t j = j0 + t;
        if (j >= J || (batch && batch[j] != k)) continue;
        double x = 0.0;
        for (int m = 0; m < M; ++m) x = fma(a_s[m], V[(int64_t)m * ldJ + j], x);
        s += exp(beta[j] + x);
      }
      s = warp_sum(s);
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
      __syncthreads();
      if (threadIdx.x == 0) {
        double tsum = 0.0;
        for (int w = 0; w < 8; ++w) tsum += red[w];
        part[((int64_t)k * XR_CAP + slot) * nchunks + blockIdx.x] = tsum;
      }
      __syncthreads();
    }
  }
}

__global__ void exact_rows_finish_kernel(int64_t ldI, int nbatch, const int* __restrict__ list, int nchunks,
                                         const double* __restrict__ part, double* __restrict__ S) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int count = list[0];
  if (idx >= count * nbatch) return;
  const int slot = idx % count, k = idx / count;
  const double* p = part + ((int64_t)k * XR_CAP + slot) * nchunks;
  double s = 0.0;
  for (int c = 0; c < nchunks; ++c) s += p[c];
  S[(int64_t)k * ldI + list[1 + slot]] = s;
}

void pass_rowsum_tc(Ctx& ctx, EtaWs& ws, const PassGeom& g, const XDesc& x, const float* g_col, double* S,
                    const double* beta64) {
  static const bool no_exact = getenv("SCGBM_NO_EXACT_ROWS") != nullptr;      // experiments
  const bool exact = !no_exact && beta64 && x.U64 && x.d64 && x.V64 && !x.clamp && x.M >= 1 && x.M <= 64;
  int* flag = nullptr;
  if (exact) {
    if (ws.row_flag.n < g.ldI) ws.row_flag.alloc(g.ldI);
    if (ws.exact_rows.n < 1) { ws.exact_rows.alloc(1); ws.exact_rows.zero(ctx.stream); }
    ws.row_flag.zero(ctx.stream);
    flag = ws.row_flag.get();
  }
  if (g.nbatch == 1) {
    lane_sums_tc(ctx, ws, *x.A16, *x.B16, g.ldI, g.J, g_col, x.clamp != 0, S, SCGBM_K_ROWSUM, nullptr, flag);
  } else {
    // S_ik = sum over the cells of batch k: one pass per batch with exp(beta) masked to that batch
    // (and, for the X0 form, the batch's row scaling as a per-lane factor)
    const int64_t n_pad = g.ldJ + 128;
    if (ws.gmask.n < n_pad) ws.gmask.alloc(n_pad);
    for (int k = 0; k < g.nbatch; ++k) {
      {
        ProfScope ps(ctx.prof, SCGBM_K_MISC, 1);
        mask_batch_kernel<<<(unsigned)ceil_div(n_pad, 256), 256, 0, ctx.stream>>>(g.J, n_pad, g_col, g.batch, k, ws.gmask.get());
        SCGBM_LAUNCH_CHECK();
      }
      lane_sums_tc(ctx, ws, *x.A16, *x.B16, g.ldI, g.J, ws.gmask.get(), x.clamp != 0, S + (int64_t)k * g.ldI, SCGBM_K_ROWSUM,
                   x.rowscale ? x.rowscale + (int64_t)k * g.ldI : nullptr, flag);
    }
  }
  if (exact) {
    const int nchunks = (int)ceil_div(g.J, XR_CHUNK);
    if (ws.exact_list.n < 1 + XR_CAP) ws.exact_list.alloc(1 + XR_CAP);
    const int64_t need = (int64_t)g.nbatch * XR_CAP * nchunks;
    if (ws.exact_part.n < need) ws.exact_part.alloc(need);
    ProfScope ps(ctx.prof, SCGBM_K_ROWSUM, 3);
    exact_rows_compact_kernel<<<1, 1024, 0, ctx.stream>>>(g.I, flag, ws.exact_list.get(), ws.exact_rows.get());
    exact_rows_partial_kernel<<<dim3((unsigned)nchunks, XR_LANES), 256, 0, ctx.stream>>>(
        g.ldI, g.J, g.ldJ, x.M, x.U64, x.d64, x.V64, beta64, g.batch, g.nbatch, ws.exact_list.get(), nchunks, ws.exact_part.get());
    exact_rows_finish_kernel<<<(unsigned)ceil_div((int64_t)XR_CAP * g.nbatch, 256), 256, 0, ctx.stream>>>(
        g.ldI, g.nbatch, ws.exact_list.get(), nchunks, ws.exact_part.get(), S);
    SCGBM_LAUNCH_CHECK();
  }
}

// C_j = sum_i f_i exp(x_ij)   (R/fit.R:134): the same kernel with the roles of the factors swapped
void pass_colsum_tc(Ctx& ctx, EtaWs& ws, const PassGeom& g, const XDesc& x, const float* f_row, double* C) {
  lane_sums_tc(ctx, ws, *x.B16, *x.A16, g.ldJ, g.ldI, f_row, x.clamp != 0, C, SCGBM_K_COLSUM);
}

}  // namespace scgbm
