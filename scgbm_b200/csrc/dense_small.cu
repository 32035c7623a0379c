// dense_small.cu -- K6: the fp64 "small side" of the block-Krylov truncated SVD that
// replaces irlba::irlba (R/fit.R:115,323): tall-skinny Gram / multiply kernels, a
// single-CTA parallel Jacobi symmetric eigensolver, and the orthonormalisation built
// from them.  Everything stays on the device; the host only sequences launches.
#include "dense_small.cuh"

namespace scgbm {

// ---------------------------------------------------------------------------------
// ts_gram:  C[p x q] = alpha * A^T B,   A: n x p (lda), B: n x q (ldb), n large.
// Stage 1: grid (chunks, p/64, q/64); each CTA reduces its row chunk into a 64x64
// block of partial[chunk] (fp64).  Stage 2 sums the chunks in a fixed order.
// ---------------------------------------------------------------------------------
constexpr int GR_TB = 64;     // output block edge
constexpr int GR_KS = 32;     // rows per smem slab
constexpr int GR_NT = 256;    // threads: 16 x 16, each 4 x 4 outputs

__global__ void __launch_bounds__(GR_NT)
ts_gram_partial(int64_t n, int p, int q, const double* __restrict__ A, int64_t lda,
                const double* __restrict__ B, int64_t ldb, double* __restrict__ part,
                int64_t rows_per_chunk) {
  __shared__ double As[GR_KS][GR_TB + 1];
  __shared__ double Bs[GR_KS][GR_TB + 1];
  const int chunk = blockIdx.x;
  const int p0 = blockIdx.y * GR_TB, q0 = blockIdx.z * GR_TB;
  const int tp = threadIdx.x % 16, tq = threadIdx.x / 16;
  const int64_t r_begin = (int64_t)chunk * rows_per_chunk;
  const int64_t r_end = min(n, r_begin + rows_per_chunk);
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;

  for (int64_t r0 = r_begin; r0 < r_end; r0 += GR_KS) {
    // load slabs: GR_KS rows x 64 columns each; rows contiguous in global (col-major)
    for (int idx = threadIdx.x; idx < GR_KS * GR_TB; idx += GR_NT) {
      const int rr = idx % GR_KS, cc = idx / GR_KS;
      const int64_t r = r0 + rr;
      double va = 0.0, vb = 0.0;
      if (r < r_end) {
        if (p0 + cc < p) va = A[(int64_t)(p0 + cc) * lda + r];
        if (q0 + cc < q) vb = B[(int64_t)(q0 + cc) * ldb + r];
      }
      As[rr][cc] = va;
      Bs[rr][cc] = vb;
    }
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < GR_KS; ++k) {
      double av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) av[a] = As[k][tp + 16 * a];
#pragma unroll
      for (int b = 0; b < 4; ++b) bv[b] = Bs[k][tq + 16 * b];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
    }
    __syncthreads();
  }
  double* out = part + (int64_t)chunk * p * q;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int pi = p0 + tp + 16 * a, qi = q0 + tq + 16 * b;
      if (pi < p && qi < q) out[(int64_t)qi * p + pi] = acc[a][b];
    }
}

__global__ void ts_gram_reduce(int nchunk, int p, int q, const double* __restrict__ part,
                               double* __restrict__ C, int ldc, double alpha, double beta) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p * q) return;
  double s = 0.0;
  for (int c = 0; c < nchunk; ++c) s += part[(int64_t)c * p * q + idx];
  const int pi = idx % p, qi = idx / p;
  double* dst = &C[(int64_t)qi * ldc + pi];
  *dst = (beta == 0.0 ? 0.0 : beta * (*dst)) + alpha * s;
}

void ts_gram(Ctx& ctx, SmallWs& ws, int64_t n, int p, int q, const double* A, int64_t lda,
             const double* B, int64_t ldb, double* C, int ldc, double alpha, double beta) {
  if (p <= 0 || q <= 0) return;
  int64_t rows_per_chunk = 2048;
  int nchunk = (int)ceil_div(n, rows_per_chunk);
  const int max_chunks = ctx.num_sms * 4;
  if (nchunk > max_chunks) {
    rows_per_chunk = round_up(ceil_div(n, max_chunks), GR_KS);
    nchunk = (int)ceil_div(n, rows_per_chunk);
  }
  if (nchunk < 1) nchunk = 1;
  ws.need_partial((int64_t)nchunk * p * q);
  ProfScope ps(ctx.prof, SCGBM_K_SMALL, 2);
  dim3 grid(nchunk, (unsigned)ceil_div(p, GR_TB), (unsigned)ceil_div(q, GR_TB));
  ts_gram_partial<<<grid, GR_NT, 0, ctx.stream>>>(n, p, q, A, lda, B, ldb, ws.partial.get(),
                                                  rows_per_chunk);
  SCGBM_LAUNCH_CHECK();
  ts_gram_reduce<<<(unsigned)ceil_div(p * q, 256), 256, 0, ctx.stream>>>(nchunk, p, q, ws.partial.get(),
                                                                        C, ldc, alpha, beta);
  SCGBM_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------
// ts_mul:  C[n x q] = beta * C + alpha * A[n x p] * S[p x q]      (S small, col-major)
// CTA: 128 rows x 64 cols; thread: 4 rows (stride 32) x 8 cols.
// ---------------------------------------------------------------------------------
constexpr int TM_ROWS = 128, TM_COLS = 64, TM_KS = 16, TM_NT = 256;

__global__ void __launch_bounds__(TM_NT)
ts_mul_kernel(int64_t n, int p, int q, const double* __restrict__ A, int64_t lda,
              const double* __restrict__ S, int lds, double* __restrict__ C, int64_t ldc,
              double alpha, double beta) {
  __shared__ double As[TM_KS][TM_ROWS];
  __shared__ double Ss[TM_KS][TM_COLS];
  const int64_t r0 = (int64_t)blockIdx.x * TM_ROWS;
  const int c0 = blockIdx.y * TM_COLS;
  const int lane = threadIdx.x % 32, cg = threadIdx.x / 32;
  double acc[4][8];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.0;
  for (int k0 = 0; k0 < p; k0 += TM_KS) {
    for (int idx = threadIdx.x; idx < TM_KS * TM_ROWS; idx += TM_NT) {
      const int rr = idx % TM_ROWS, kk = idx / TM_ROWS;
      const int64_t r = r0 + rr;
      As[kk][rr] = (r < n && k0 + kk < p) ? A[(int64_t)(k0 + kk) * lda + r] : 0.0;
    }
    for (int idx = threadIdx.x; idx < TM_KS * TM_COLS; idx += TM_NT) {
      const int kk = idx % TM_KS, cc = idx / TM_KS;
      Ss[kk][cc] = (k0 + kk < p && c0 + cc < q) ? S[(int64_t)(c0 + cc) * lds + k0 + kk] : 0.0;
    }
    __syncthreads();
#pragma unroll 4
    for (int kk = 0; kk < TM_KS; ++kk) {
      double av[4], sv[8];
#pragma unroll
      for (int a = 0; a < 4; ++a) av[a] = As[kk][lane + 32 * a];
#pragma unroll
      for (int b = 0; b < 8; ++b) sv[b] = Ss[kk][cg * 8 + b];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = fma(av[a], sv[b], acc[a][b]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int64_t r = r0 + lane + 32 * a;
    if (r >= n) continue;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const int c = c0 + cg * 8 + b;
      if (c < q) {
        double* dst = &C[(int64_t)c * ldc + r];
        const double old = (beta == 0.0) ? 0.0 : beta * (*dst);
        *dst = old + alpha * acc[a][b];
      }
    }
  }
}

void ts_mul(Ctx& ctx, int64_t n, int p, int q, const double* A, int64_t lda, const double* S,
            int lds, double* C, int64_t ldc, double alpha, double beta) {
  if (n <= 0 || q <= 0) return;
  ProfScope ps(ctx.prof, SCGBM_K_SMALL, 1);
  dim3 grid((unsigned)ceil_div(n, TM_ROWS), (unsigned)ceil_div(q, TM_COLS));
  ts_mul_kernel<<<grid, TM_NT, 0, ctx.stream>>>(n, p, q, A, lda, S, lds, C, ldc, alpha, beta);
  SCGBM_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------
// Symmetric eigensolver: parallel cyclic two-sided Jacobi in ONE CTA.
// A (n x n, col-major, lda = n) is overwritten; on exit `w` holds eigenvalues in
// DESCENDING order and `Vout` the matching eigenvectors (columns).
// Round-robin ("circle") ordering gives n/2 disjoint rotations per round.
// ---------------------------------------------------------------------------------
constexpr int JE_NT = 1024;

__device__ __forceinline__ void jacobi_pair(int r, int k, int m, int& p, int& q) {
  // circle method on m (even) players; player m-1 is fixed
  int a, b;
  if (k == 0) { a = m - 1; b = r % (m - 1); }
  else {
    a = (r + k) % (m - 1);
    b = (r - k + (m - 1)) % (m - 1);
  }
  p = min(a, b); q = max(a, b);
}

__global__ void __launch_bounds__(JE_NT, 1)
jacobi_eigh_kernel(int n, double* __restrict__ Ag, double* __restrict__ Vg, double* __restrict__ w,
                   double* __restrict__ Vout, int use_smem, int max_sweeps, int* __restrict__ info) {
  extern __shared__ double sm[];
  const int m = (n % 2 == 0) ? n : n + 1;   // virtual size (dummy index n when odd)
  const int half = m / 2;
  double* cs = sm;                 // [half] cos
  double* sn = cs + half;          // [half] sin
  double* red = sn + half;         // [32] reduction scratch
  double* A = Ag;
  double* V = Vg;
  if (use_smem) {
    A = red + 32;
    V = A + (size_t)n * n;
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) A[idx] = Ag[idx];
  }
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) V[idx] = ((idx % n) == (idx / n)) ? 1.0 : 0.0;
  __syncthreads();

  // Frobenius norm^2 for the stopping rule
  double loc = 0.0;
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) loc += A[idx] * A[idx];
  loc = warp_sum(loc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = loc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
    v = warp_sum(v);
    if (threadIdx.x == 0) red[0] = v;
  }
  __syncthreads();
  const double fro2 = red[0];
  __syncthreads();

  int sweeps_done = 0;
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    // off-diagonal norm^2
    double off = 0.0;
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
      const int i = idx % n, j = idx / n;
      if (i != j) off += A[idx] * A[idx];
    }
    off = warp_sum(off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = off;
    __syncthreads();
    if (threadIdx.x < 32) {
      double v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
      v = warp_sum(v);
      if (threadIdx.x == 0) red[0] = v;
    }
    __syncthreads();
    const double off2 = red[0];
    __syncthreads();
    if (off2 <= 1e-30 * fro2 || fro2 == 0.0) break;
    sweeps_done = sweep + 1;

    for (int r = 0; r < m - 1; ++r) {
      // 1. rotation angles
      for (int k = threadIdx.x; k < half; k += blockDim.x) {
        int p, q;
        jacobi_pair(r, k, m, p, q);
        double c = 1.0, s = 0.0;
        if (q < n) {
          const double apq = A[(size_t)q * n + p];
          const double app = A[(size_t)p * n + p], aqq = A[(size_t)q * n + q];
          if (fabs(apq) > 1e-300 && fabs(apq) > 1e-18 * sqrt(fabs(app * aqq)) ) {
            const double tau = (aqq - app) / (2.0 * apq);
            const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            c = 1.0 / sqrt(1.0 + t * t);
            s = t * c;
          }
        }
        cs[k] = c; sn[k] = s;
      }
      __syncthreads();
      // 2. columns: A <- A J,  V <- V J
      for (int idx = threadIdx.x; idx < half * n; idx += blockDim.x) {
        const int k = idx / n, i = idx % n;
        int p, q;
        jacobi_pair(r, k, m, p, q);
        if (q >= n) continue;
        const double c = cs[k], s = sn[k];
        if (s == 0.0) continue;
        const double ap = A[(size_t)p * n + i], aq = A[(size_t)q * n + i];
        A[(size_t)p * n + i] = c * ap - s * aq;
        A[(size_t)q * n + i] = s * ap + c * aq;
        const double vp = V[(size_t)p * n + i], vq = V[(size_t)q * n + i];
        V[(size_t)p * n + i] = c * vp - s * vq;
        V[(size_t)q * n + i] = s * vp + c * vq;
      }
      __syncthreads();
      // 3. rows: A <- J^T A
      for (int idx = threadIdx.x; idx < half * n; idx += blockDim.x) {
        const int k = idx % half, j = idx / half;
        int p, q;
        jacobi_pair(r, k, m, p, q);
        if (q >= n) continue;
        const double c = cs[k], s = sn[k];
        if (s == 0.0) continue;
        const double ap = A[(size_t)j * n + p], aq = A[(size_t)j * n + q];
        A[(size_t)j * n + p] = c * ap - s * aq;
        A[(size_t)j * n + q] = s * ap + c * aq;
      }
      __syncthreads();
    }
  }
  // sort eigenvalues descending (rank by counting), write out
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double di = A[(size_t)i * n + i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const double dj = A[(size_t)j * n + j];
      if (dj > di || (dj == di && j < i)) ++rank;
    }
    w[rank] = di;
    // stash the destination column in cs/sn space is too small for n; use red-free approach below
    for (int k = 0; k < n; ++k) Vout[(size_t)rank * n + k] = V[(size_t)i * n + k];
  }
  if (threadIdx.x == 0 && info) info[0] = sweeps_done;
}

void eigh_desc(Ctx& ctx, SmallWs& ws, int n, double* A, double* w, double* Vout) {
  SCGBM_REQUIRE(n >= 1 && n <= 1024, "eigh_desc: n out of range");
  const int m = (n % 2 == 0) ? n : n + 1;
  size_t base = (size_t)(m + 32) * sizeof(double);
  size_t full = base + (size_t)2 * n * n * sizeof(double);
  int use_smem = full <= 200 * 1024;
  size_t smem = use_smem ? full : base;
  ws.need_eigV((int64_t)n * n);
  static bool attr_set = false;
  if (!attr_set) {
    SCGBM_CUDA(cudaFuncSetAttribute(jacobi_eigh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 1024));
    attr_set = true;
  }
  ProfScope ps(ctx.prof, SCGBM_K_SMALL, 1);
  int threads = n <= 32 ? 256 : (n <= 64 ? 512 : JE_NT);
  jacobi_eigh_kernel<<<1, threads, smem, ctx.stream>>>(n, A, ws.eigV.get(), w, Vout, use_smem, 40, nullptr);
  SCGBM_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------
// E[:,k] *= (lam[k] > eps*lam[0]) ? 1/sqrt(lam[k]) : 0        (lam descending)
__global__ void scale_cols_invsqrt(int n, double* __restrict__ E, const double* __restrict__ lam, double eps) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * n) return;
  const int k = idx / n;
  const double l0 = lam[0], lk = lam[k];
  const double f = (lk > eps * l0 && lk > 0.0) ? rsqrt(lk) : 0.0;
  E[idx] *= f;
}

__global__ void copy_strided_kernel(int64_t n, int q, const double* __restrict__ src, int64_t lds,
                                    double* __restrict__ dst, int64_t ldd) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * q) return;
  const int64_t r = idx % n;
  const int c = (int)(idx / n);
  dst[(int64_t)c * ldd + r] = src[(int64_t)c * lds + r];
}

void copy_mat(Ctx& ctx, int64_t n, int q, const double* src, int64_t lds, double* dst, int64_t ldd) {
  if (n * q <= 0) return;
  ProfScope ps(ctx.prof, SCGBM_K_SMALL, 1);
  copy_strided_kernel<<<(unsigned)ceil_div(n * q, 256), 256, 0, ctx.stream>>>(n, q, src, lds, dst, ldd);
  SCGBM_LAUNCH_CHECK();
}

// K (n x b, ldk) <- orthonormal basis of span(K) orthogonal to Qb (n x pb); columns
// whose direction is numerically absent become zero.  Two rounds (CholeskyQR2-like,
// but through the eigendecomposition of the Gram so rank deficiency is harmless).
void sym_orth(Ctx& ctx, SmallWs& ws, int64_t n, int b, double* K, int64_t ldk, const double* Qb,
              int64_t ldq, int pb, double* tmp_nb /* n x b scratch, ld = ldk */) {
  ws.need_small((int64_t)std::max(pb, b) * b + 3 * (int64_t)b * b + b);
  double* C = ws.small.get();                       // pb x b
  double* S = C + (int64_t)std::max(pb, b) * b;     // b x b
  double* E = S + (int64_t)b * b;                   // b x b
  double* lam = E + (int64_t)b * b;                 // b
  for (int round = 0; round < 2; ++round) {
    if (pb > 0) {
      ts_gram(ctx, ws, n, pb, b, Qb, ldq, K, ldk, C, pb, 1.0, 0.0);
      ts_mul(ctx, n, pb, b, Qb, ldq, C, pb, K, ldk, -1.0, 1.0);
    }
    ts_gram(ctx, ws, n, b, b, K, ldk, K, ldk, S, b, 1.0, 0.0);
    eigh_desc(ctx, ws, b, S, lam, E);
    {
      ProfScope ps(ctx.prof, SCGBM_K_SMALL, 1);
      scale_cols_invsqrt<<<(unsigned)ceil_div(b * b, 256), 256, 0, ctx.stream>>>(b, E, lam, 1e-12);
      SCGBM_LAUNCH_CHECK();
    }
    ts_mul(ctx, n, b, b, K, ldk, E, b, tmp_nb, ldk, 1.0, 0.0);
    copy_mat(ctx, n, b, tmp_nb, ldk, K, ldk);
  }
}

}  // namespace scgbm
