// counts.cu -- K1: pack the count matrix into its device format and compute the exact
// integer statistics the fit needs (max(Y), colSums, per-batch rowSums: R/fit.R:94-96,
// 102-105) while validating the input (R/fit.R:287-299).  Also the on-device
// simData-style generator (R/simulate.R:24-48,60-65) used for benchmark inputs.
#include "counts.cuh"

namespace scgbm {

struct PackStats {
  unsigned long long max_bits;  // bit pattern of the max as a non-negative double
  int flags;                    // 1: negative, 2: NaN, 4: non-integral, 8: does not fit Tout
};

template <typename T> struct OutLimit;
template <> struct OutLimit<uint8_t> { static constexpr double max = 255.0; static constexpr bool integral = true; };
template <> struct OutLimit<uint16_t> { static constexpr double max = 65535.0; static constexpr bool integral = true; };
template <> struct OutLimit<float> { static constexpr double max = 3.0e38; static constexpr bool integral = false; };

constexpr int PK_ROWS = 256, PK_COLS = 16;

// in: rows x jc (ld_in); out: ld_out x jc, rows >= n_rows are written as zero.
template <typename Tin, typename Tout>
__global__ void __launch_bounds__(PK_ROWS)
pack_stats_kernel(const Tin* __restrict__ in, int64_t ld_in, const int64_t* __restrict__ rowmap,
                  int64_t n_rows, int64_t ld_out, int64_t jc, Tout* __restrict__ out,
                  const int32_t* __restrict__ batch, double* __restrict__ colsum,
                  double* __restrict__ rowsum, PackStats* __restrict__ st, int write_out) {
  __shared__ double s_col[PK_COLS][PK_ROWS / 32];
  const int64_t i = (int64_t)blockIdx.x * PK_ROWS + threadIdx.x;
  const int64_t j0 = (int64_t)blockIdx.y * PK_COLS;
  const bool active = i < n_rows;
  const int64_t src_row = active ? (rowmap ? rowmap[i] : i) : 0;
  double racc = 0.0;
  int rb = -1;
  double vmax = 0.0;
  int flags = 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int jj = 0; jj < PK_COLS; ++jj) {
    const int64_t j = j0 + jj;
    double v = 0.0;
    if (j < jc && active) {
      v = (double)in[j * ld_in + src_row];
      if (v != v) { flags |= 2; v = 0.0; }
      if (v < 0.0) flags |= 1;
      if (OutLimit<Tout>::integral && v != floor(v)) flags |= 4;
      if (v > OutLimit<Tout>::max) flags |= 8;
      vmax = fmax(vmax, v);
      const int b = batch ? batch[j] : 0;
      if (b != rb) {
        if (rb >= 0 && racc != 0.0) atomicAdd(&rowsum[(int64_t)rb * ld_out + i], racc);
        rb = b; racc = 0.0;
      }
      racc += v;
    }
    if (j < jc && write_out && i < ld_out) out[j * ld_out + i] = (Tout)v;
    const double cs = warp_sum(v);
    if (lane == 0) s_col[jj][warp] = cs;
  }
  if (rb >= 0 && racc != 0.0) atomicAdd(&rowsum[(int64_t)rb * ld_out + i], racc);
  __syncthreads();
  if (threadIdx.x < PK_COLS) {
    const int64_t j = j0 + threadIdx.x;
    if (j < jc) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < PK_ROWS / 32; ++w) s += s_col[threadIdx.x][w];
      if (s != 0.0) atomicAdd(&colsum[j], s);   // integer-valued sums: order-independent
    }
  }
  vmax = warp_max(vmax);
  flags |= __shfl_xor_sync(0xffffffffu, flags, 16);
  flags |= __shfl_xor_sync(0xffffffffu, flags, 8);
  flags |= __shfl_xor_sync(0xffffffffu, flags, 4);
  flags |= __shfl_xor_sync(0xffffffffu, flags, 2);
  flags |= __shfl_xor_sync(0xffffffffu, flags, 1);
  if (lane == 0) {
    if (vmax > 0.0) atomicMax(&st->max_bits, (unsigned long long)__double_as_longlong(vmax));
    if (flags) atomicOr(&st->flags, flags);
  }
}

template <typename Tin, typename Tout>
static void launch_pack(Ctx& ctx, cudaStream_t s, const void* in, int64_t ld_in, const int64_t* rowmap,
                        int64_t n_rows, int64_t ld_out, int64_t jc, void* out, const int32_t* batch,
                        double* colsum, double* rowsum, PackStats* st, int write_out) {
  dim3 grid((unsigned)ceil_div(ld_out, PK_ROWS), (unsigned)ceil_div(jc, PK_COLS));
  pack_stats_kernel<Tin, Tout><<<grid, PK_ROWS, 0, s>>>((const Tin*)in, ld_in, rowmap, n_rows, ld_out, jc,
                                                        (Tout*)out, batch, colsum, rowsum, st, write_out);
  SCGBM_LAUNCH_CHECK();
}

template <typename Tout>
static void launch_pack_in(Ctx& ctx, cudaStream_t s, int y_dtype, const void* in, int64_t ld_in,
                           const int64_t* rowmap, int64_t n_rows, int64_t ld_out, int64_t jc, void* out,
                           const int32_t* batch, double* colsum, double* rowsum, PackStats* st,
                           int write_out) {
  switch (y_dtype) {
    case SCGBM_I32: launch_pack<int32_t, Tout>(ctx, s, in, ld_in, rowmap, n_rows, ld_out, jc, out, batch, colsum, rowsum, st, write_out); break;
    case SCGBM_F64: launch_pack<double, Tout>(ctx, s, in, ld_in, rowmap, n_rows, ld_out, jc, out, batch, colsum, rowsum, st, write_out); break;
    case SCGBM_U8: launch_pack<uint8_t, Tout>(ctx, s, in, ld_in, rowmap, n_rows, ld_out, jc, out, batch, colsum, rowsum, st, write_out); break;
    case SCGBM_U16: launch_pack<uint16_t, Tout>(ctx, s, in, ld_in, rowmap, n_rows, ld_out, jc, out, batch, colsum, rowsum, st, write_out); break;
    case SCGBM_F32: launch_pack<float, Tout>(ctx, s, in, ld_in, rowmap, n_rows, ld_out, jc, out, batch, colsum, rowsum, st, write_out); break;
    default: throw Error(SCGBM_ERR_INVALID, "unknown y_dtype");
  }
}

static int dtype_bytes(int y_dtype) {
  switch (y_dtype) {
    case SCGBM_I32: return 4; case SCGBM_F64: return 8; case SCGBM_U8: return 1;
    case SCGBM_U16: return 2; case SCGBM_F32: return 4;
  }
  throw Error(SCGBM_ERR_INVALID, "unknown y_dtype");
}

static void pack_dispatch(Ctx& ctx, cudaStream_t s, int storage, int y_dtype, const void* in, int64_t ld_in,
                          const int64_t* rowmap, int64_t n_rows, int64_t ld_out, int64_t jc, void* out,
                          const int32_t* batch, double* colsum, double* rowsum, PackStats* st, int write_out) {
  switch (storage) {
    case SCGBM_STORE_U8: launch_pack_in<uint8_t>(ctx, s, y_dtype, in, ld_in, rowmap, n_rows, ld_out, jc, out, batch, colsum, rowsum, st, write_out); break;
    case SCGBM_STORE_U16: launch_pack_in<uint16_t>(ctx, s, y_dtype, in, ld_in, rowmap, n_rows, ld_out, jc, out, batch, colsum, rowsum, st, write_out); break;
    case SCGBM_STORE_F32: launch_pack_in<float>(ctx, s, y_dtype, in, ld_in, rowmap, n_rows, ld_out, jc, out, batch, colsum, rowsum, st, write_out); break;
    default: throw Error(SCGBM_ERR_INVALID, "unknown storage");
  }
}

// row sums over ALL source rows of a raw (caller-typed) column block: out[i] += sum_j in[j * ld + i]
// (gbm.proj.parallel needs rowSums(Y) of every gene, R/fit.R:217, also of the genes its fit filters out)
template <typename Tin>
__global__ void __launch_bounds__(256)
raw_rowsum_kernel(const Tin* __restrict__ in, int64_t ld_in, int64_t I_src, int64_t jc, int64_t cols_per_block,
                  double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= I_src) return;
  const int64_t ja = (int64_t)blockIdx.y * cols_per_block, jb = min(jc, ja + cols_per_block);
  double s = 0.0;
  for (int64_t j = ja; j < jb; ++j) {
    const double v = (double)in[j * ld_in + i];
    s += (v == v) ? v : 0.0;
  }
  if (s != 0.0) atomicAdd(&out[i], s);        // integer-valued sums: order-independent
}

static void raw_rowsum(cudaStream_t s, int y_dtype, const void* in, int64_t ld_in, int64_t I_src, int64_t jc, double* out) {
  const int64_t cpb = 512;
  dim3 grid((unsigned)ceil_div(I_src, 256), (unsigned)ceil_div(jc, cpb));
  switch (y_dtype) {
    case SCGBM_I32: raw_rowsum_kernel<int32_t><<<grid, 256, 0, s>>>((const int32_t*)in, ld_in, I_src, jc, cpb, out); break;
    case SCGBM_F64: raw_rowsum_kernel<double><<<grid, 256, 0, s>>>((const double*)in, ld_in, I_src, jc, cpb, out); break;
    case SCGBM_U8: raw_rowsum_kernel<uint8_t><<<grid, 256, 0, s>>>((const uint8_t*)in, ld_in, I_src, jc, cpb, out); break;
    case SCGBM_U16: raw_rowsum_kernel<uint16_t><<<grid, 256, 0, s>>>((const uint16_t*)in, ld_in, I_src, jc, cpb, out); break;
    case SCGBM_F32: raw_rowsum_kernel<float><<<grid, 256, 0, s>>>((const float*)in, ld_in, I_src, jc, cpb, out); break;
    default: throw Error(SCGBM_ERR_INVALID, "unknown y_dtype");
  }
  SCGBM_LAUNCH_CHECK();
}

static void load_impl(Ctx& ctx, const void* Y, int y_dtype, int y_on_device, int64_t I_src, int64_t J,
                      int64_t ldY, const int64_t* rowmap_host, int64_t n_rows, const int32_t* batch_host,
                      int nbatch, int storage_req, Counts& c, double* full_rowsum = nullptr) {
  SCGBM_REQUIRE(Y != nullptr, "Y is NULL");
  SCGBM_REQUIRE(n_rows >= 1 && J >= 1, "Y must have at least one row and one column");
  SCGBM_REQUIRE(ldY >= I_src, "ldY < I");
  SCGBM_REQUIRE(nbatch >= 1, "nbatch < 1");
  c.I = n_rows; c.J = J; c.ld = round_up(n_rows, kRowPad); c.nbatch = nbatch;
  c.colsum.alloc(J);
  c.rowsum.alloc(c.ld * nbatch);
  DevBuf<int64_t> rowmap;
  if (rowmap_host) {
    rowmap.alloc(n_rows);
    SCGBM_CUDA(cudaMemcpyAsync(rowmap.get(), rowmap_host, n_rows * sizeof(int64_t), cudaMemcpyHostToDevice, ctx.stream));
  }
  if (batch_host && nbatch > 1) {
    std::vector<int32_t> b0(J);
    for (int64_t j = 0; j < J; ++j) {
      SCGBM_REQUIRE(batch_host[j] >= 1 && batch_host[j] <= nbatch, "batch codes must be in 1..nbatch");
      b0[j] = batch_host[j] - 1;
    }
    c.batch.alloc(J);
    SCGBM_CUDA(cudaMemcpyAsync(c.batch.get(), b0.data(), J * sizeof(int32_t), cudaMemcpyHostToDevice, ctx.stream));
    SCGBM_CUDA(cudaStreamSynchronize(ctx.stream));
  }
  DevBuf<PackStats> st(1);
  const int eb = dtype_bytes(y_dtype);

  int storage = storage_req == SCGBM_STORE_AUTO ? SCGBM_STORE_U16 : storage_req;
  for (int attempt = 0; attempt < 2; ++attempt) {
    c.storage = storage;
    // alias a resident device matrix that already has the native layout
    const bool native = y_on_device && !rowmap_host && ldY == c.ld &&
                        ((storage == SCGBM_STORE_U8 && y_dtype == SCGBM_U8) ||
                         (storage == SCGBM_STORE_U16 && y_dtype == SCGBM_U16) ||
                         (storage == SCGBM_STORE_F32 && y_dtype == SCGBM_F32));
    if (native) {
      c.owned.release();
      c.data = const_cast<void*>(Y);
    } else {
      c.owned.alloc(c.ld * J * c.elem_bytes());
      c.data = c.owned.get();
    }
    c.colsum.zero(ctx.stream);
    c.rowsum.zero(ctx.stream);
    st.zero(ctx.stream);
    if (full_rowsum) SCGBM_CUDA(cudaMemsetAsync(full_rowsum, 0, (size_t)I_src * sizeof(double), ctx.stream));
    if (y_on_device) {
      ProfScope ps(ctx.prof, SCGBM_K_PACK, 1);
      if (full_rowsum) raw_rowsum(ctx.stream, y_dtype, Y, ldY, I_src, J, full_rowsum);
      pack_dispatch(ctx, ctx.stream, storage, y_dtype, Y, ldY, rowmap.get(), n_rows, c.ld, J, c.data,
                    c.batch.get(), c.colsum.get(), c.rowsum.get(), st.get(), native ? 0 : 1);
    } else {
      // chunked H2D (copy stream) overlapped with packing (compute stream)
      int64_t cols_per_chunk = std::max<int64_t>(1, (int64_t)(256ll << 20) / (ldY * eb));
      cols_per_chunk = std::min(cols_per_chunk, J);
      DevBuf<uint8_t> raw[2];
      raw[0].alloc(cols_per_chunk * ldY * eb);
      if (cols_per_chunk < J) raw[1].alloc(cols_per_chunk * ldY * eb);
      cudaStream_t cs;
      SCGBM_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
      cudaEvent_t copied[2], packed[2];
      for (int k = 0; k < 2; ++k) {
        SCGBM_CUDA(cudaEventCreateWithFlags(&copied[k], cudaEventDisableTiming));
        SCGBM_CUDA(cudaEventCreateWithFlags(&packed[k], cudaEventDisableTiming));
      }
      try {
        SCGBM_CUDA(cudaStreamSynchronize(ctx.stream));
        int k = 0;
        int64_t nchunk = 0;
        for (int64_t j0 = 0; j0 < J; j0 += cols_per_chunk, k ^= 1, ++nchunk) {
          const int64_t jc = std::min(cols_per_chunk, J - j0);
          if (nchunk >= 2) SCGBM_CUDA(cudaStreamWaitEvent(cs, packed[k], 0));
          const uint8_t* src = (const uint8_t*)Y + (size_t)j0 * ldY * eb;
          // last column may be shorter than ldY in the caller's allocation
          const size_t bytes = ((size_t)(jc - 1) * ldY + I_src) * eb;
          SCGBM_CUDA(cudaMemcpyAsync(raw[k].get(), src, bytes, cudaMemcpyHostToDevice, cs));
          ctx.prof.h2d_bytes += (double)bytes;
          SCGBM_CUDA(cudaEventRecord(copied[k], cs));
          SCGBM_CUDA(cudaStreamWaitEvent(ctx.stream, copied[k], 0));
          {
            ProfScope ps(ctx.prof, SCGBM_K_PACK, 1);
            if (full_rowsum) raw_rowsum(ctx.stream, y_dtype, raw[k].get(), ldY, I_src, jc, full_rowsum);
            pack_dispatch(ctx, ctx.stream, storage, y_dtype, raw[k].get(), ldY, rowmap.get(), n_rows, c.ld, jc,
                          (uint8_t*)c.data + (size_t)j0 * c.ld * c.elem_bytes(),
                          c.batch.get() ? c.batch.get() + j0 : nullptr, c.colsum.get() + j0,
                          c.rowsum.get(), st.get(), 1);
          }
          SCGBM_CUDA(cudaEventRecord(packed[k], ctx.stream));
        }
        SCGBM_CUDA(cudaStreamSynchronize(ctx.stream));
      } catch (...) {
        cudaStreamSynchronize(cs); cudaStreamDestroy(cs);
        for (int k = 0; k < 2; ++k) { cudaEventDestroy(copied[k]); cudaEventDestroy(packed[k]); }
        throw;
      }
      cudaStreamSynchronize(cs); cudaStreamDestroy(cs);
      for (int k = 0; k < 2; ++k) { cudaEventDestroy(copied[k]); cudaEventDestroy(packed[k]); }
    }
    PackStats h;
    SCGBM_CUDA(cudaMemcpyAsync(&h, st.get(), sizeof(h), cudaMemcpyDeviceToHost, ctx.stream));
    SCGBM_CUDA(cudaStreamSynchronize(ctx.stream));
    if (h.flags & 2) throw Error(SCGBM_ERR_INVALID, "Count matrix Y cannot contain NA values.");
    if (h.flags & 1) throw Error(SCGBM_ERR_INVALID, "Count matrix Y cannot contain negative values.");
    if (h.flags & (4 | 8)) {
      if (storage_req == SCGBM_STORE_AUTO && storage != SCGBM_STORE_F32) { storage = SCGBM_STORE_F32; continue; }
      throw Error(SCGBM_ERR_INVALID, "Y does not fit the requested device storage format");
    }
    long long bits = (long long)h.max_bits;
    double mx;
    memcpy(&mx, &bits, sizeof(mx));
    c.maxY = mx;
    return;
  }
}

// ---------------------------------------------------------------------------------
// sparse ingest: one warp per column scatters its non-zeros into the (zeroed) dense device matrix
// ---------------------------------------------------------------------------------
template <typename Tx, typename Tout>
__global__ void csc_scatter_kernel(int64_t I, int64_t ld, int64_t jc, const int64_t* __restrict__ p /* jc + 1, chunk-relative base p[0] */,
                                   const int32_t* __restrict__ ri, const Tx* __restrict__ x, Tout* __restrict__ out,
                                   PackStats* __restrict__ st) {
  const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (j >= jc) return;
  const int64_t base = p[0], a = p[j] - base, b = p[j + 1] - base;
  int flags = 0;
  for (int64_t k = a + lane; k < b; k += 32) {
    const int32_t i = ri[k];
    double v = (double)x[k];
    if (i < 0 || i >= I) { flags |= 16; continue; }
    if (v != v) { flags |= 2; v = 0.0; }
    if (v < 0.0) flags |= 1;
    if (OutLimit<Tout>::integral && v != floor(v)) flags |= 4;
    if (v > OutLimit<Tout>::max) flags |= 8;
    out[j * ld + i] = (Tout)v;
  }
  if (flags) atomicOr(&st->flags, flags);
}

template <typename Tout>
static void launch_scatter(cudaStream_t s, int x_dtype, int64_t I, int64_t ld, int64_t jc, const int64_t* p,
                           const int32_t* ri, const void* x, void* out, PackStats* st) {
  const unsigned grid = (unsigned)ceil_div(jc * 32, 256);
  switch (x_dtype) {
    case SCGBM_F64: csc_scatter_kernel<double, Tout><<<grid, 256, 0, s>>>(I, ld, jc, p, ri, (const double*)x, (Tout*)out, st); break;
    case SCGBM_F32: csc_scatter_kernel<float, Tout><<<grid, 256, 0, s>>>(I, ld, jc, p, ri, (const float*)x, (Tout*)out, st); break;
    case SCGBM_I32: csc_scatter_kernel<int32_t, Tout><<<grid, 256, 0, s>>>(I, ld, jc, p, ri, (const int32_t*)x, (Tout*)out, st); break;
    case SCGBM_U16: csc_scatter_kernel<uint16_t, Tout><<<grid, 256, 0, s>>>(I, ld, jc, p, ri, (const uint16_t*)x, (Tout*)out, st); break;
    case SCGBM_U8: csc_scatter_kernel<uint8_t, Tout><<<grid, 256, 0, s>>>(I, ld, jc, p, ri, (const uint8_t*)x, (Tout*)out, st); break;
    default: throw Error(SCGBM_ERR_INVALID, "unknown csc_x_dtype");
  }
  SCGBM_LAUNCH_CHECK();
}

void load_counts_csc(Ctx& ctx, const int32_t* csc_i, const int64_t* csc_p, const void* csc_x, int x_dtype,
                     int64_t I, int64_t J, const int32_t* batch_host, int nbatch, int storage_req, Counts& c) {
  SCGBM_REQUIRE(csc_i && csc_p && csc_x, "sparse input needs csc_i, csc_p and csc_x");
  SCGBM_REQUIRE(I >= 1 && J >= 1 && nbatch >= 1, "Y must have at least one row and one column");
  SCGBM_REQUIRE(csc_p[0] >= 0, "csc_p must be non-decreasing");
  for (int64_t j = 0; j < J; ++j) SCGBM_REQUIRE(csc_p[j + 1] >= csc_p[j], "csc_p must be non-decreasing");
  c.I = I; c.J = J; c.ld = round_up(I, kRowPad); c.nbatch = nbatch;
  c.colsum.alloc(J);
  c.rowsum.alloc(c.ld * nbatch);
  if (batch_host && nbatch > 1) {
    std::vector<int32_t> b0(J);
    for (int64_t j = 0; j < J; ++j) {
      SCGBM_REQUIRE(batch_host[j] >= 1 && batch_host[j] <= nbatch, "batch codes must be in 1..nbatch");
      b0[j] = batch_host[j] - 1;
    }
    c.batch.alloc(J);
    SCGBM_CUDA(cudaMemcpyAsync(c.batch.get(), b0.data(), J * sizeof(int32_t), cudaMemcpyHostToDevice, ctx.stream));
    SCGBM_CUDA(cudaStreamSynchronize(ctx.stream));
  }
  DevBuf<PackStats> st(1);
  const int xb = dtype_bytes(x_dtype);
  // column chunks of at most ~32M non-zeros (and at least one column)
  const int64_t max_nnz = 32ll << 20;
  int storage = storage_req == SCGBM_STORE_AUTO ? SCGBM_STORE_U16 : storage_req;
  for (int attempt = 0; attempt < 2; ++attempt) {
    c.storage = storage;
    c.owned.alloc(c.ld * J * c.elem_bytes());
    c.data = c.owned.get();
    SCGBM_CUDA(cudaMemsetAsync(c.data, 0, (size_t)c.ld * J * c.elem_bytes(), ctx.stream));
    c.colsum.zero(ctx.stream); c.rowsum.zero(ctx.stream); st.zero(ctx.stream);
    DevBuf<int64_t> dp;
    DevBuf<int32_t> di;
    DevBuf<uint8_t> dx;
    for (int64_t j0 = 0; j0 < J;) {
      int64_t j1 = j0 + 1;
      while (j1 < J && csc_p[j1 + 1] - csc_p[j0] <= max_nnz && j1 - j0 < (1 << 22)) ++j1;
      const int64_t jc = j1 - j0, n0 = csc_p[j0], nn = csc_p[j1] - n0;
      if (dp.n < jc + 1) dp.alloc(jc + 1);
      if (di.n < nn) { di.alloc(std::max<int64_t>(nn, 1)); dx.alloc(std::max<int64_t>(nn, 1) * xb); }
      SCGBM_CUDA(cudaMemcpyAsync(dp.get(), csc_p + j0, (jc + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx.stream));
      if (nn > 0) {
        SCGBM_CUDA(cudaMemcpyAsync(di.get(), csc_i + n0, nn * sizeof(int32_t), cudaMemcpyHostToDevice, ctx.stream));
        SCGBM_CUDA(cudaMemcpyAsync(dx.get(), (const uint8_t*)csc_x + (size_t)n0 * xb, (size_t)nn * xb, cudaMemcpyHostToDevice, ctx.stream));
        ctx.prof.h2d_bytes += (double)nn * (4 + xb) + (double)(jc + 1) * 8;
        ProfScope ps(ctx.prof, SCGBM_K_PACK, 1);
        void* out = (uint8_t*)c.data + (size_t)j0 * c.ld * c.elem_bytes();
        switch (storage) {
          case SCGBM_STORE_U8: launch_scatter<uint8_t>(ctx.stream, x_dtype, I, c.ld, jc, dp.get(), di.get(), dx.get(), out, st.get()); break;
          case SCGBM_STORE_U16: launch_scatter<uint16_t>(ctx.stream, x_dtype, I, c.ld, jc, dp.get(), di.get(), dx.get(), out, st.get()); break;
          default: launch_scatter<float>(ctx.stream, x_dtype, I, c.ld, jc, dp.get(), di.get(), dx.get(), out, st.get()); break;
        }
      }
      SCGBM_CUDA(cudaStreamSynchronize(ctx.stream));   // the staging buffers are reused by the next chunk
      j0 = j1;
    }
    {   // exact integer statistics from the dense device matrix (same kernel as the dense path, stats only)
      ProfScope ps(ctx.prof, SCGBM_K_PACK, 1);
      const int native_dtype = storage == SCGBM_STORE_U8 ? SCGBM_U8 : (storage == SCGBM_STORE_U16 ? SCGBM_U16 : SCGBM_F32);
      pack_dispatch(ctx, ctx.stream, storage, native_dtype, c.data, c.ld, nullptr, I, c.ld, J, c.data, c.batch.get(),
                    c.colsum.get(), c.rowsum.get(), st.get(), 0);
    }
    PackStats h;
    SCGBM_CUDA(cudaMemcpyAsync(&h, st.get(), sizeof(h), cudaMemcpyDeviceToHost, ctx.stream));
    SCGBM_CUDA(cudaStreamSynchronize(ctx.stream));
    if (h.flags & 16) throw Error(SCGBM_ERR_INVALID, "sparse Y: row index out of range");
    if (h.flags & 2) throw Error(SCGBM_ERR_INVALID, "Count matrix Y cannot contain NA values.");
    if (h.flags & 1) throw Error(SCGBM_ERR_INVALID, "Count matrix Y cannot contain negative values.");
    if (h.flags & (4 | 8)) {
      if (storage_req == SCGBM_STORE_AUTO && storage != SCGBM_STORE_F32) { storage = SCGBM_STORE_F32; continue; }
      throw Error(SCGBM_ERR_INVALID, "Y does not fit the requested device storage format");
    }
    long long bits = (long long)h.max_bits;
    double mx;
    memcpy(&mx, &bits, sizeof(mx));
    c.maxY = mx;
    return;
  }
}

void load_counts(Ctx& ctx, const void* Y, int y_dtype, int y_on_device, int64_t I, int64_t J, int64_t ldY,
                 const int32_t* batch_host, int nbatch, int storage_req, Counts& c) {
  load_impl(ctx, Y, y_dtype, y_on_device, I, J, ldY, nullptr, I, batch_host, nbatch, storage_req, c);
}

// u16 -> u8 in a fresh buffer (16 counts per thread); the caller has checked max(Y) <= 255
__global__ void narrow_u16_u8_kernel(int64_t n16, const uint4* __restrict__ in, uint4* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n16) return;
  const uint4 a = in[2 * i], b = in[2 * i + 1];
  auto pk = [](uint32_t lo, uint32_t hi) {     // four u16 (two words) -> four u8
    return (lo & 0xffu) | ((lo >> 8) & 0xff00u) | ((hi & 0xffu) << 16) | ((hi >> 16) << 24);
  };
  out[i] = make_uint4(pk(a.x, a.y), pk(a.z, a.w), pk(b.x, b.y), pk(b.z, b.w));
}

void narrow_counts_to_u8(Ctx& ctx, Counts& c) {
  SCGBM_REQUIRE(c.storage == SCGBM_STORE_U16 && c.maxY <= 255.0, "narrow_counts_to_u8: not applicable");
  const int64_t n = c.ld * c.J;                 // ld is a multiple of 64: n is a multiple of 16
  DevBuf<uint8_t> dst(n);
  ProfScope ps(ctx.prof, SCGBM_K_PACK, 1);
  narrow_u16_u8_kernel<<<(unsigned)ceil_div(n / 16, 256), 256, 0, ctx.stream>>>(
      n / 16, reinterpret_cast<const uint4*>(c.data), reinterpret_cast<uint4*>(dst.get()));
  SCGBM_LAUNCH_CHECK();
  SCGBM_CUDA(cudaStreamSynchronize(ctx.stream));
  c.owned = std::move(dst);                     // releases the u16 copy if the library owned it (an aliased caller
  c.data = c.owned.get();                       // buffer is left alone)
  c.storage = SCGBM_STORE_U8;
}

void load_counts_rows(Ctx& ctx, const void* Y, int y_dtype, int y_on_device, int64_t I_full, int64_t J,
                      int64_t ldY, const int64_t* ixs_host, int64_t Ik, Counts& c, double* full_rowsum_dev) {
  for (int64_t r = 0; r < Ik; ++r)
    SCGBM_REQUIRE(ixs_host[r] >= 0 && ixs_host[r] < I_full, "ixs out of range");
  load_impl(ctx, Y, y_dtype, y_on_device, I_full, J, ldY, ixs_host, Ik, nullptr, 1, SCGBM_STORE_AUTO, c, full_rowsum_dev);
}

}  // namespace scgbm
