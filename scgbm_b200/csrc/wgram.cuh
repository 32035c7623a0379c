// wgram.cuh -- shared device routines for K8 (get.se) and K9 (projection IRLS):
//   * weighted Gram accumulation  G_c[a,b] += sum_r w[r,c] F[r,a] F[r,b]  with the
//     (a,b) pairs spread over the threads of a CTA and CC targets per CTA, so each
//     product F[r,a]F[r,b] is formed once and reused for CC weights;
//   * warp-level Cholesky of an M x M SPD matrix in shared memory, solve, and the
//     diagonal of the inverse (sqrt(diag(solve(Mat))) of R/uncertainty.R:24,31).
#pragma once
#include "common.cuh"

namespace scgbm {

constexpr int WG_NT = 256;   // threads per CTA
constexpr int WG_CC = 8;     // targets (cells or genes) per CTA
constexpr int WG_TR = 32;    // reduction rows per tile

__host__ __device__ inline int tri_count(int p) { return p * (p + 1) / 2; }

// pair index -> (a <= b), packed by columns: idx = b(b+1)/2 + a
__device__ __forceinline__ void fill_pair_table(int p, unsigned short* tab_a, unsigned short* tab_b) {
  const int T = tri_count(p);
  for (int idx = threadIdx.x; idx < T; idx += blockDim.x) {
    int b = (int)((sqrtf(8.0f * idx + 1.0f) - 1.0f) * 0.5f);
    while ((b + 1) * (b + 2) / 2 <= idx) ++b;
    while (b * (b + 1) / 2 > idx) --b;
    tab_a[idx] = (unsigned short)(idx - b * (b + 1) / 2);
    tab_b[idx] = (unsigned short)b;
  }
}

// acc[pi][c] += F_s[r][a] F_s[r][b] w_s[r][c]   over the WG_TR rows of a tile
template <int PT>
__device__ __forceinline__ void wgram_tile(const double* __restrict__ F_s, int ldf, const double* __restrict__ w_s,
                                           const unsigned short* __restrict__ tab_a,
                                           const unsigned short* __restrict__ tab_b, int T, int rows,
                                           double (&acc)[PT][WG_CC]) {
  int pa[PT], pb[PT];
#pragma unroll
  for (int pi = 0; pi < PT; ++pi) {
    const int idx = threadIdx.x + pi * WG_NT;
    pa[pi] = idx < T ? tab_a[idx] : -1;
    pb[pi] = idx < T ? tab_b[idx] : 0;
  }
  for (int r = 0; r < rows; ++r) {
    double w[WG_CC];
#pragma unroll
    for (int c = 0; c < WG_CC; ++c) w[c] = w_s[r * WG_CC + c];
#pragma unroll
    for (int pi = 0; pi < PT; ++pi) {
      if (pa[pi] < 0) continue;
      const double prod = F_s[r * ldf + pa[pi]] * F_s[r * ldf + pb[pi]];
#pragma unroll
      for (int c = 0; c < WG_CC; ++c) acc[pi][c] = fma(prod, w[c], acc[pi][c]);
    }
  }
}

// One warp: in-place Cholesky A = L L^T of the p x p SPD matrix `A` (row-major, ld =
// p, full storage in shared memory).  Returns false when a pivot is not positive.
__device__ __forceinline__ bool warp_cholesky(double* A, int p, int lane) {
  bool ok = true;
  for (int k = 0; k < p; ++k) {
    __syncwarp();
    const double akk = A[k * p + k];
    if (!(akk > 0.0)) { ok = false; break; }
    const double lkk = sqrt(akk);
    __syncwarp();
    if (lane == 0) A[k * p + k] = lkk;
    for (int i = k + 1 + lane; i < p; i += 32) A[i * p + k] /= lkk;
    __syncwarp();
    // trailing update of the lower triangle: A[i][j] -= L[i][k] L[j][k], j in (k, i]
    for (int i = k + 1 + lane; i < p; i += 32) {
      const double lik = A[i * p + k];
      for (int j = k + 1; j <= i; ++j) A[i * p + j] -= lik * A[j * p + k];
    }
  }
  __syncwarp();
  return ok;
}

// One warp: diag(A^{-1})_c = sum_{k >= c} (L^{-1})[k][c]^2 ; lane c builds column c of
// L^{-1} by forward substitution in a private array (p <= WG_PMAX).
constexpr int WG_PMAX = 64;
__device__ __forceinline__ void warp_inv_diag(const double* L, int p, int lane, double* diag_out) {
  for (int c = lane; c < p; c += 32) {
    double xcol[WG_PMAX];
    double s2 = 0.0;
    for (int i = c; i < p; ++i) {
      double s = (i == c) ? 1.0 : 0.0;
      for (int k = c; k < i; ++k) s -= L[i * p + k] * xcol[k];
      const double x = s / L[i * p + i];
      xcol[i] = x;
      s2 += x * x;
    }
    diag_out[c] = s2;
  }
  __syncwarp();
}

// One warp: solve L L^T x = rhs (x overwrites rhs, length p, shared memory).
__device__ __forceinline__ void warp_chol_solve(const double* L, double* x, int p, int lane) {
  // forward: L y = rhs
  for (int i = 0; i < p; ++i) {
    double part = 0.0;
    for (int k = lane; k < i; k += 32) part += L[i * p + k] * x[k];
    part = warp_sum(part);
    __syncwarp();
    if (lane == 0) x[i] = (x[i] - part) / L[i * p + i];
    __syncwarp();
  }
  // backward: L^T x = y
  for (int i = p - 1; i >= 0; --i) {
    double part = 0.0;
    for (int k = i + 1 + lane; k < p; k += 32) part += L[k * p + i] * x[k];
    part = warp_sum(part);
    __syncwarp();
    if (lane == 0) x[i] = (x[i] - part) / L[i * p + i];
    __syncwarp();
  }
}

}  // namespace scgbm
