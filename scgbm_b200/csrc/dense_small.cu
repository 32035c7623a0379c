// dense_small.cu -- K6: the fp64 "small side" of the block-Krylov truncated SVD that
// replaces irlba::irlba (R/fit.R:115,323): tall-skinny Gram / multiply kernels, a
// single-CTA parallel Jacobi symmetric eigensolver, and the orthonormalisation built
// from them.  Everything stays on the device; the host only sequences launches.
#include <cooperative_groups.h>
#include "dense_small.cuh"

namespace scgbm {

// ---------------------------------------------------------------------------------
// ts_gram:  C[p x q] = alpha * A^T B,   A: n x p (lda), B: n x q (ldb), n large.
// Stage 1: grid (chunks, p/PT, q/QT); each CTA reduces its row chunk into a PT x QT
// block of partial[chunk] (fp64).  Stage 2 sums the chunks in a fixed order.
// ---------------------------------------------------------------------------------
constexpr int GR_TB = 64;     // output block edge
constexpr int GR_KS = 32;     // rows per smem slab
constexpr int GR_NT = 256;    // threads: 16 x 16, each 4 x 4 outputs

// PT x QT output block (64 or 32 each): the Krylov blocks are 32 wide at M = 20, so a fixed 64 x 64 block
// would spend half (q = 32) to 85 % (p = 20, q = 32) of its fp64 FMAs on zero padding.
template <int PT, int QT>
__global__ void __launch_bounds__(GR_NT)
ts_gram_partial(int64_t n, int p, int q, const double* __restrict__ A, int64_t lda,
                const double* __restrict__ B, int64_t ldb, double* __restrict__ part,
                int64_t rows_per_chunk) {
  constexpr int NA = PT / 16, NB = QT / 16;
  __shared__ double As[GR_KS][PT + 1];
  __shared__ double Bs[GR_KS][QT + 1];
  const int chunk = blockIdx.x;
  const int p0 = blockIdx.y * PT, q0 = blockIdx.z * QT;
  const int tp = threadIdx.x % 16, tq = threadIdx.x / 16;
  const int64_t r_begin = (int64_t)chunk * rows_per_chunk;
  const int64_t r_end = min(n, r_begin + rows_per_chunk);
  double acc[NA][NB];
#pragma unroll
  for (int a = 0; a < NA; ++a)
#pragma unroll
    for (int b = 0; b < NB; ++b) acc[a][b] = 0.0;

  for (int64_t r0 = r_begin; r0 < r_end; r0 += GR_KS) {
    // load slabs: GR_KS rows x PT (QT) columns; rows contiguous in global (col-major)
    for (int idx = threadIdx.x; idx < GR_KS * PT; idx += GR_NT) {
      const int rr = idx % GR_KS, cc = idx / GR_KS;
      const int64_t r = r0 + rr;
      As[rr][cc] = (r < r_end && p0 + cc < p) ? A[(int64_t)(p0 + cc) * lda + r] : 0.0;
    }
    for (int idx = threadIdx.x; idx < GR_KS * QT; idx += GR_NT) {
      const int rr = idx % GR_KS, cc = idx / GR_KS;
      const int64_t r = r0 + rr;
      Bs[rr][cc] = (r < r_end && q0 + cc < q) ? B[(int64_t)(q0 + cc) * ldb + r] : 0.0;
    }
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < GR_KS; ++k) {
      double av[NA], bv[NB];
#pragma unroll
      for (int a = 0; a < NA; ++a) av[a] = As[k][tp + 16 * a];
#pragma unroll
      for (int b = 0; b < NB; ++b) bv[b] = Bs[k][tq + 16 * b];
#pragma unroll
      for (int a = 0; a < NA; ++a)
#pragma unroll
        for (int b = 0; b < NB; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
    }
    __syncthreads();
  }
  double* out = part + (int64_t)chunk * p * q;
#pragma unroll
  for (int a = 0; a < NA; ++a)
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const int pi = p0 + tp + 16 * a, qi = q0 + tq + 16 * b;
      if (pi < p && qi < q) out[(int64_t)qi * p + pi] = acc[a][b];
    }
}

// one warp per output entry: lanes stride over the chunks, fixed-order shuffle tree
__global__ void ts_gram_reduce(int nchunk, int p, int q, const double* __restrict__ part,
                               double* __restrict__ C, int ldc, double alpha, double beta) {
  const int idx = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (idx >= p * q) return;
  double s = 0.0;
  for (int c = lane; c < nchunk; c += 32) s += part[(int64_t)c * p * q + idx];
  s = warp_sum(s);
  if (lane == 0) {
    const int pi = idx % p, qi = idx / p;
    double* dst = &C[(int64_t)qi * ldc + pi];
    *dst = (beta == 0.0 ? 0.0 : beta * (*dst)) + alpha * s;
  }
}

void ts_gram(Ctx& ctx, SmallWs& ws, int64_t n, int p, int q, const double* A, int64_t lda,
             const double* B, int64_t ldb, double* C, int ldc, double alpha, double beta) {
  if (p <= 0 || q <= 0) return;
  // enough chunks to fill the machine twice over, at least two smem slabs each
  const int PT = p <= 32 ? 32 : GR_TB, QT = q <= 32 ? 32 : GR_TB;
  const int64_t blocks_pq = ceil_div(p, PT) * ceil_div(q, QT);
  const int64_t want_chunks = std::max<int64_t>(1, ceil_div((int64_t)ctx.num_sms * 2, blocks_pq));
  int64_t rows_per_chunk = std::max<int64_t>(2 * GR_KS, round_up(ceil_div(n, want_chunks), GR_KS));
  int nchunk = (int)std::max<int64_t>(1, ceil_div(n, rows_per_chunk));
  ws.need_partial((int64_t)nchunk * p * q);
  ProfScope ps(ctx.prof, SCGBM_K_SMALL, 2);
  dim3 grid(nchunk, (unsigned)ceil_div(p, PT), (unsigned)ceil_div(q, QT));
  if (PT == 32 && QT == 32) ts_gram_partial<32, 32><<<grid, GR_NT, 0, ctx.stream>>>(n, p, q, A, lda, B, ldb, ws.partial.get(), rows_per_chunk);
  else if (PT == 64 && QT == 32) ts_gram_partial<64, 32><<<grid, GR_NT, 0, ctx.stream>>>(n, p, q, A, lda, B, ldb, ws.partial.get(), rows_per_chunk);
  else if (PT == 32 && QT == 64) ts_gram_partial<32, 64><<<grid, GR_NT, 0, ctx.stream>>>(n, p, q, A, lda, B, ldb, ws.partial.get(), rows_per_chunk);
  else ts_gram_partial<64, 64><<<grid, GR_NT, 0, ctx.stream>>>(n, p, q, A, lda, B, ldb, ws.partial.get(), rows_per_chunk);
  SCGBM_LAUNCH_CHECK();
  ts_gram_reduce<<<(unsigned)ceil_div((int64_t)p * q * 32, 256), 256, 0, ctx.stream>>>(nchunk, p, q, ws.partial.get(),
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

// Tall-skinny variant (q <= 64, p <= 128, large n): one thread per row keeps its q outputs in
// registers, S sits in shared memory (broadcast reads), every global access is a coalesced column
// segment and all of a thread's loads are independent -- HBM-bound instead of latency-bound.
template <int QMAX>
__global__ void __launch_bounds__(128)
ts_mul_rows(int64_t n, int p, int q, const double* __restrict__ A, int64_t lda, const double* __restrict__ S, int lds,
            double* __restrict__ C, int64_t ldc, double alpha, double beta) {
  extern __shared__ double tm_sm[];         // [p][QMAX]
  for (int idx = threadIdx.x; idx < p * QMAX; idx += blockDim.x) {
    const int k = idx / QMAX, c = idx % QMAX;
    tm_sm[idx] = c < q ? S[(int64_t)c * lds + k] : 0.0;
  }
  __syncthreads();
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  double acc[QMAX];
#pragma unroll
  for (int c = 0; c < QMAX; ++c) acc[c] = 0.0;
#pragma unroll 4
  for (int k = 0; k < p; ++k) {
    const double a = A[(int64_t)k * lda + r];
    const double* srow = tm_sm + k * QMAX;
#pragma unroll
    for (int c = 0; c < QMAX; ++c) acc[c] = fma(a, srow[c], acc[c]);
  }
#pragma unroll
  for (int c = 0; c < QMAX; ++c) {
    if (c < q) {
      double* dst = &C[(int64_t)c * ldc + r];
      const double old = (beta == 0.0) ? 0.0 : beta * (*dst);
      *dst = old + alpha * acc[c];
    }
  }
}

void ts_mul(Ctx& ctx, int64_t n, int p, int q, const double* A, int64_t lda, const double* S,
            int lds, double* C, int64_t ldc, double alpha, double beta) {
  if (n <= 0 || q <= 0) return;
  static const bool no_rows = getenv("SCGBM_MUL_OLD") != nullptr;
  if (!no_rows && q <= 64 && p <= 160 && n >= 65536) {
    ProfScope ps(ctx.prof, SCGBM_K_SMALL, 1);
    const unsigned grid = (unsigned)ceil_div(n, 128);
    if (q <= 32) {
      const size_t smem = (size_t)p * 32 * sizeof(double);
      ts_mul_rows<32><<<grid, 128, smem, ctx.stream>>>(n, p, q, A, lda, S, lds, C, ldc, alpha, beta);
    } else {
      const size_t smem = (size_t)p * 64 * sizeof(double);
      if (smem > 48 * 1024)
        SCGBM_CUDA(cudaFuncSetAttribute(ts_mul_rows<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      ts_mul_rows<64><<<grid, 128, smem, ctx.stream>>>(n, p, q, A, lda, S, lds, C, ldc, alpha, beta);
    }
    SCGBM_LAUNCH_CHECK();
    return;
  }
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

__device__ __forceinline__ double block_sum_1024(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
    t = warp_sum(t);
    if (threadIdx.x == 0) red[0] = t;
  }
  __syncthreads();
  const double out = red[0];
  __syncthreads();
  return out;
}

// A and V live in shared memory (leading dimension lds, odd so that row sweeps are
// bank-conflict free) when they fit, else in global memory (L2-resident).
__global__ void __launch_bounds__(JE_NT, 1)
jacobi_eigh_kernel(int n, double* __restrict__ Ag, double* __restrict__ Vg, double* __restrict__ w,
                   double* __restrict__ Vout, int a_smem, int v_smem, int max_sweeps) {
  extern __shared__ double sm[];
  const int m = (n % 2 == 0) ? n : n + 1;   // virtual size (dummy index n when odd)
  const int half = m / 2;
  const int lds = n | 1;
  double* cs = sm;                 // [half] cos
  double* sn = cs + half;          // [half] sin
  double* red = sn + half;         // [32]
  int* pq = reinterpret_cast<int*>(red + 32);   // [half] p | q << 16
  double* nxt = reinterpret_cast<double*>(pq + ((half + 1) & ~1));
  double* A = Ag; int lda = n;
  double* V = Vg; int ldv = n;
  if (a_smem) { A = nxt; lda = lds; nxt += (size_t)n * lds; }
  if (v_smem) { V = nxt; ldv = lds; }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
    const int i = idx % n, j = idx / n;
    if (a_smem) A[(size_t)j * lda + i] = Ag[idx];
    V[(size_t)j * ldv + i] = (i == j) ? 1.0 : 0.0;
  }
  __syncthreads();
  double loc = 0.0;
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
    const double a = A[(size_t)(idx / n) * lda + (idx % n)];
    loc += a * a;
  }
  const double fro2 = block_sum_1024(loc, red);

  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    double off = 0.0;
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
      const int i = idx % n, j = idx / n;
      if (i != j) { const double a = A[(size_t)j * lda + i]; off += a * a; }
    }
    const double off2 = block_sum_1024(off, red);
    if (off2 <= 1e-28 * fro2 || fro2 == 0.0) break;

    for (int r = 0; r < m - 1; ++r) {
      // 1. rotation angles for the half disjoint pairs of this round
      for (int k = threadIdx.x; k < half; k += blockDim.x) {
        int p, q;
        jacobi_pair(r, k, m, p, q);
        double c = 1.0, s = 0.0;
        if (q < n) {
          const double apq = A[(size_t)q * lda + p];
          const double app = A[(size_t)p * lda + p], aqq = A[(size_t)q * lda + q];
          if (fabs(apq) > 1e-300 && fabs(apq) > 1e-17 * sqrt(fabs(app * aqq))) {
            const double tau = (aqq - app) / (2.0 * apq);
            const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            c = rsqrt(1.0 + t * t);
            s = t * c;
          }
        } else {
          q = -1;
        }
        cs[k] = c; sn[k] = s; pq[k] = (q < 0 || s == 0.0) ? -1 : (p | (q << 16));
      }
      __syncthreads();
      // 2. columns: A <- A J, V <- V J   (one warp per pair, lanes over rows)
      for (int k = warp; k < half; k += nwarps) {
        const int code = pq[k];
        if (code < 0) continue;
        const int p = code & 0xffff, q = code >> 16;
        const double c = cs[k], s = sn[k];
        double* ap = A + (size_t)p * lda; double* aq = A + (size_t)q * lda;
        double* vp = V + (size_t)p * ldv; double* vq = V + (size_t)q * ldv;
        for (int i = lane; i < n; i += 32) {
          const double x = ap[i], y = aq[i];
          ap[i] = c * x - s * y; aq[i] = s * x + c * y;
          const double u = vp[i], v = vq[i];
          vp[i] = c * u - s * v; vq[i] = s * u + c * v;
        }
      }
      __syncthreads();
      // 3. rows: A <- J^T A   (one warp per pair, lanes over columns)
      for (int k = warp; k < half; k += nwarps) {
        const int code = pq[k];
        if (code < 0) continue;
        const int p = code & 0xffff, q = code >> 16;
        const double c = cs[k], s = sn[k];
        for (int j = lane; j < n; j += 32) {
          double* col = A + (size_t)j * lda;
          const double x = col[p], y = col[q];
          col[p] = c * x - s * y; col[q] = s * x + c * y;
        }
      }
      __syncthreads();
    }
  }
  // sort eigenvalues descending (rank by counting), write out
  for (int i = warp; i < n; i += nwarps) {
    const double di = A[(size_t)i * lda + i];
    int rank = 0;
    for (int j = lane; j < n; j += 32) {
      const double dj = A[(size_t)j * lda + j];
      if (dj > di || (dj == di && j < i)) ++rank;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
    if (lane == 0) w[rank] = di;
    for (int k = lane; k < n; k += 32) Vout[(size_t)rank * n + k] = V[(size_t)i * ldv + k];
  }
}

// ---------------------------------------------------------------------------------
// The same cyclic Jacobi for matrices too large for one CTA's shared memory (n > 128: the
// cold-start Ritz problems, p = 5b up to 320 at M = 50): a cooperative grid, one grid-wide
// barrier per rotation round.  A and V ping-pong between two global (L2-resident) buffers:
// every CTA derives the round's n/2 rotations redundantly from the OLD buffer, then the grid
// rewrites each 2 x 2 block  B' = J_r^T B J_c  (and each column pair of V) into the NEW one,
// so no CTA ever reads what another is writing.  Odd n is padded with a zero row/column whose
// rotations are identities.
// ---------------------------------------------------------------------------------
namespace cg = cooperative_groups;

__global__ void jacobi_coop_init_kernel(int n, int m, const double* __restrict__ A, double* __restrict__ A0,
                                        double* __restrict__ V0) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= m * m) return;
  const int i = idx % m, j = idx / m;
  A0[idx] = (i < n && j < n) ? A[(size_t)j * n + i] : 0.0;
  V0[idx] = (i == j) ? 1.0 : 0.0;
}

__global__ void __launch_bounds__(256)
jacobi_coop_kernel(int n, int m, double* __restrict__ Abuf0, double* __restrict__ Abuf1, double* __restrict__ Vbuf0,
                   double* __restrict__ Vbuf1, double* __restrict__ w, double* __restrict__ Vout, int max_sweeps) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ double smj[];
  const int half = m / 2;
  double* cs = smj;                 // [half]
  double* sn = cs + half;           // [half]
  double* red = sn + half;          // [32]
  int* pp = reinterpret_cast<int*>(red + 32);   // [half] p
  int* qq = pp + half;                          // [half] q
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t gsize = (int64_t)gridDim.x * blockDim.x;
  double* Ab[2] = {Abuf0, Abuf1};
  double* Vb[2] = {Vbuf0, Vbuf1};
  int cur = 0;
  double fro2 = 0.0;
  for (int sweep = 0; sweep <= max_sweeps; ++sweep) {
    // every CTA evaluates the same convergence test on the same data
    const double* A = Ab[cur];
    double off = 0.0, all = 0.0;
    for (int idx = threadIdx.x; idx < m * m; idx += blockDim.x) {
      const double a = A[idx];
      all += a * a;
      if (idx % m != idx / m) off += a * a;
    }
    const double off2 = block_sum_1024(off, red);
    if (sweep == 0) fro2 = block_sum_1024(all, red);
    if (off2 <= 1e-28 * fro2 || fro2 == 0.0 || sweep == max_sweeps) break;
    for (int r = 0; r < m - 1; ++r) {
      const double* Ao = Ab[cur];
      const double* Vo = Vb[cur];
      double* An = Ab[cur ^ 1];
      double* Vn = Vb[cur ^ 1];
      for (int k = threadIdx.x; k < half; k += blockDim.x) {
        int p, q;
        jacobi_pair(r, k, m, p, q);
        double c = 1.0, s = 0.0;
        const double apq = Ao[(size_t)q * m + p];
        const double app = Ao[(size_t)p * m + p], aqq = Ao[(size_t)q * m + q];
        if (fabs(apq) > 1e-300 && fabs(apq) > 1e-17 * sqrt(fabs(app * aqq))) {
          const double tau = (aqq - app) / (2.0 * apq);
          const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
          c = rsqrt(1.0 + t * t);
          s = t * c;
        }
        cs[k] = c; sn[k] = s; pp[k] = p; qq[k] = q;
      }
      __syncthreads();
      const int64_t nblk = (int64_t)half * half;
      for (int64_t item = gtid; item < nblk + (int64_t)half * m; item += gsize) {
        if (item < nblk) {        // A: rows (pr, qr) x columns (pc, qc)
          const int kr = (int)(item % half), kc = (int)(item / half);
          const int pr = pp[kr], qr = qq[kr], pc = pp[kc], qc = qq[kc];
          const double cr = cs[kr], sr = sn[kr], cc = cs[kc], sc = sn[kc];
          const double b00 = Ao[(size_t)pc * m + pr], b10 = Ao[(size_t)pc * m + qr];
          const double b01 = Ao[(size_t)qc * m + pr], b11 = Ao[(size_t)qc * m + qr];
          const double t00 = cc * b00 - sc * b01, t01 = sc * b00 + cc * b01;
          const double t10 = cc * b10 - sc * b11, t11 = sc * b10 + cc * b11;
          An[(size_t)pc * m + pr] = cr * t00 - sr * t10;
          An[(size_t)pc * m + qr] = sr * t00 + cr * t10;
          An[(size_t)qc * m + pr] = cr * t01 - sr * t11;
          An[(size_t)qc * m + qr] = sr * t01 + cr * t11;
        } else {                  // V: column pair kc, one row
          const int64_t it2 = item - nblk;
          const int i = (int)(it2 % m), kc = (int)(it2 / m);
          const int pc = pp[kc], qc = qq[kc];
          const double c = cs[kc], s2 = sn[kc];
          const double x = Vo[(size_t)pc * m + i], y = Vo[(size_t)qc * m + i];
          Vn[(size_t)pc * m + i] = c * x - s2 * y;
          Vn[(size_t)qc * m + i] = s2 * x + c * y;
        }
      }
      grid.sync();
      cur ^= 1;
    }
  }
  // sort eigenvalues descending (rank by counting) and write out; the pad index n never mixes
  const double* A = Ab[cur];
  const double* V = Vb[cur];
  const int gwarp = (int)(gtid >> 5), nwarps = (int)(gsize >> 5), lane = threadIdx.x & 31;
  for (int i = gwarp; i < n; i += nwarps) {
    const double di = A[(size_t)i * m + i];
    int rank = 0;
    for (int j = lane; j < n; j += 32) {
      const double dj = A[(size_t)j * m + j];
      if (dj > di || (dj == di && j < i)) ++rank;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
    if (lane == 0) w[rank] = di;
    for (int k = lane; k < n; k += 32) Vout[(size_t)rank * n + k] = V[(size_t)i * m + k];
  }
}

static void eigh_desc_coop(Ctx& ctx, SmallWs& ws, int n, double* A, double* w, double* Vout) {
  int m = (n % 2 == 0) ? n : n + 1;
  const int half = m / 2;
  ws.need_eigV(4 * (int64_t)m * m);
  double* A0 = ws.eigV.get();
  double* A1 = A0 + (size_t)m * m;
  double* V0 = A1 + (size_t)m * m;
  double* V1 = V0 + (size_t)m * m;
  ProfScope ps(ctx.prof, SCGBM_K_SMALL, 2);
  jacobi_coop_init_kernel<<<(unsigned)ceil_div((int64_t)m * m, 256), 256, 0, ctx.stream>>>(n, m, A, A0, V0);
  SCGBM_LAUNCH_CHECK();
  const size_t smem = (size_t)(2 * half + 32) * sizeof(double) + (size_t)2 * half * sizeof(int) + 16;
  int per_sm = 0;
  SCGBM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jacobi_coop_kernel, 256, smem));
  SCGBM_REQUIRE(per_sm >= 1, "eigh_desc: cooperative Jacobi does not fit");
  const int64_t items = (int64_t)half * half + (int64_t)half * m;
  int grid = (int)std::min<int64_t>((int64_t)ctx.num_sms * std::min(per_sm, 2), ceil_div(items, 256));
  grid = std::max(grid, 1);
  int max_sweeps = 40;
  void* args[] = {&n, &m, &A0, &A1, &V0, &V1, &w, &Vout, &max_sweeps};
  SCGBM_CUDA(cudaLaunchCooperativeKernel((void*)jacobi_coop_kernel, dim3(grid), dim3(256), args, smem, ctx.stream));
}

void eigh_desc(Ctx& ctx, SmallWs& ws, int n, double* A, double* w, double* Vout) {
  SCGBM_REQUIRE(n >= 1 && n <= 1024, "eigh_desc: n out of range");
  static const int coop_from = getenv("SCGBM_JACOBI_COOP_FROM") ? atoi(getenv("SCGBM_JACOBI_COOP_FROM")) : 129;
  if (n >= coop_from) { eigh_desc_coop(ctx, ws, n, A, w, Vout); return; }
  const int m = (n % 2 == 0) ? n : n + 1;
  const int lds = n | 1;
  const size_t base = (size_t)(m + 32) * sizeof(double) + (size_t)((m / 2 + 1) & ~1) * sizeof(int) + 16;
  const size_t mat = (size_t)n * lds * sizeof(double);
  const size_t cap = 220 * 1024;
  const int a_smem = base + mat <= cap;
  const int v_smem = a_smem && base + 2 * mat <= cap;
  const size_t smem = base + (a_smem ? mat : 0) + (v_smem ? mat : 0);
  ws.need_eigV((int64_t)n * n);
  SCGBM_CUDA(cudaFuncSetAttribute(jacobi_eigh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap));
  ProfScope ps(ctx.prof, SCGBM_K_SMALL, 1);
  int threads = n <= 16 ? 256 : (n <= 32 ? 512 : JE_NT);
  jacobi_eigh_kernel<<<1, threads, smem, ctx.stream>>>(n, A, ws.eigV.get(), w, Vout, a_smem, v_smem, 40);
  SCGBM_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------
// chol_inv: S (b x b SPD Gram, col-major) -> E = L^{-T} so that K E is orthonormal when
// S = K^T K (CholeskyQR).  Sets *flag when a pivot is too small (rank-deficient block).
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1)
chol_inv_kernel(int b, const double* __restrict__ S, double* __restrict__ E, int* __restrict__ flag) {
  extern __shared__ double smc[];
  const int ld = b | 1;
  double* Lm = smc;                    // [b][ld] row-major lower triangle
  double* X = Lm + (size_t)b * ld;     // [b][ld] inverse of L (lower)
  __shared__ int bad;
  __shared__ double dmax;
  if (threadIdx.x == 0) { bad = 0; dmax = 0.0; }
  for (int idx = threadIdx.x; idx < b * b; idx += blockDim.x) {
    const int i = idx % b, j = idx / b;
    Lm[(size_t)i * ld + j] = S[idx];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double d = 0.0;
    for (int i = 0; i < b; ++i) d = fmax(d, Lm[(size_t)i * ld + i]);
    dmax = d;
  }
  __syncthreads();
  for (int k = 0; k < b; ++k) {
    const double akk = Lm[(size_t)k * ld + k];
    if (!(akk > 1e-10 * dmax) || !(akk > 0.0)) { if (threadIdx.x == 0) bad = 1; __syncthreads(); break; }
    const double lkk = sqrt(akk);
    __syncthreads();
    for (int i = k + threadIdx.x; i < b; i += blockDim.x)
      Lm[(size_t)i * ld + k] = (i == k) ? lkk : Lm[(size_t)i * ld + k] / lkk;
    __syncthreads();
    // trailing update, lower triangle only
    const int rem = b - k - 1;
    for (int idx = threadIdx.x; idx < rem * rem; idx += blockDim.x) {
      const int i = k + 1 + idx / rem, j = k + 1 + idx % rem;
      if (j <= i) Lm[(size_t)i * ld + j] -= Lm[(size_t)i * ld + k] * Lm[(size_t)j * ld + k];
    }
    __syncthreads();
  }
  __syncthreads();
  if (bad) {
    if (threadIdx.x == 0) atomicExch(flag, 1);
    for (int idx = threadIdx.x; idx < b * b; idx += blockDim.x) E[idx] = (idx % b == idx / b) ? 1.0 : 0.0;
    return;
  }
  // X = L^{-1}: thread c solves column c by forward substitution
  for (int c = threadIdx.x; c < b; c += blockDim.x) {
    for (int i = 0; i < b; ++i) {
      if (i < c) { X[(size_t)i * ld + c] = 0.0; continue; }
      double s = (i == c) ? 1.0 : 0.0;
      for (int k = c; k < i; ++k) s -= Lm[(size_t)i * ld + k] * X[(size_t)k * ld + c];
      X[(size_t)i * ld + c] = s / Lm[(size_t)i * ld + i];
    }
  }
  __syncthreads();
  // E = X^T (upper triangular), column-major: E[i + j b] = X[j][i]
  for (int idx = threadIdx.x; idx < b * b; idx += blockDim.x) {
    const int i = idx % b, j = idx / b;
    E[idx] = X[(size_t)j * ld + i];
  }
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

// K (n x b, ldk) <- orthonormal basis of span(K) orthogonal to Qb (n x pb).  Fast path:
// CholeskyQR2 (Gram -> Cholesky -> K L^{-T}); a too-small pivot raises *chol_flag and
// the caller redoes the whole SVD call with robust = true, where the Gram is
// eigendecomposed instead and numerically absent directions become zero columns.
void sym_orth(Ctx& ctx, SmallWs& ws, int64_t n, int b, double* K, int64_t ldk, const double* Qb,
              int64_t ldq, int pb, double* tmp_nb /* n x b scratch, ld = ldk */, bool robust, int* chol_flag) {
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
    if (robust) {
      eigh_desc(ctx, ws, b, S, lam, E);
      ProfScope ps(ctx.prof, SCGBM_K_SMALL, 1);
      scale_cols_invsqrt<<<(unsigned)ceil_div(b * b, 256), 256, 0, ctx.stream>>>(b, E, lam, 1e-12);
      SCGBM_LAUNCH_CHECK();
    } else {
      const size_t smem = 2 * (size_t)b * (b | 1) * sizeof(double);
      SCGBM_REQUIRE(smem <= 200 * 1024, "sym_orth: block too large for the shared-memory Cholesky");
      if (smem > 48 * 1024)
        SCGBM_CUDA(cudaFuncSetAttribute(chol_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      ProfScope ps(ctx.prof, SCGBM_K_SMALL, 1);
      chol_inv_kernel<<<1, 256, smem, ctx.stream>>>(b, S, E, chol_flag);
      SCGBM_LAUNCH_CHECK();
    }
    ts_mul(ctx, n, b, b, K, ldk, E, b, tmp_nb, ldk, 1.0, 0.0);
    copy_mat(ctx, n, b, tmp_nb, ldk, K, ldk);
  }
}

}  // namespace scgbm
