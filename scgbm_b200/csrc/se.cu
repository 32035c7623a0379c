// se.cu -- K8: get.se (R/uncertainty.R:14-35).
//   se_scores[j,]   = sqrt(diag(solve(U^T diag(W[,j]) U + EPS I)))     :22-26
//   se_loadings[i,] = sqrt(diag(solve(V^T diag(W[i,]) V + EPS I)))     :29-33   (V = scores)
// The J + I information matrices are formed as one Khatri-Rao Gram sweep over W
// (each U_ia U_ib product is shared by 8 cells), then inverted by a batched
// warp-per-matrix Cholesky.  All fp64.
#include "wgram.cuh"

namespace scgbm {

struct SeGramParams {
  int64_t nT;         // targets
  int64_t nR;         // reduction length
  const double* W;    // W[r * sr + c * sc]
  int64_t sr, sc;
  const double* F;    // nR x p (col-major, ld ldf)
  int64_t ldf;
  int p;
  double* G;          // nT x T packed grams, G[c * T + pair]
  int accumulate;
};

template <int PT>
__global__ void __launch_bounds__(WG_NT)
se_gram_kernel(const SeGramParams q) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int p = q.p, T = tri_count(p), ldf = p + 1;
  double* F_s = reinterpret_cast<double*>(smraw);            // [WG_TR][p+1]
  double* w_s = F_s + WG_TR * ldf;                           // [WG_TR][WG_CC]
  unsigned short* tab_a = reinterpret_cast<unsigned short*>(w_s + WG_TR * WG_CC);
  unsigned short* tab_b = tab_a + T;
  fill_pair_table(p, tab_a, tab_b);
  const int64_t c0 = (int64_t)blockIdx.x * WG_CC;
  double acc[PT][WG_CC];
#pragma unroll
  for (int pi = 0; pi < PT; ++pi)
#pragma unroll
    for (int c = 0; c < WG_CC; ++c) acc[pi][c] = 0.0;
  for (int64_t r0 = 0; r0 < q.nR; r0 += WG_TR) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < WG_TR * p; idx += WG_NT) {
      const int r = idx % WG_TR, a = idx / WG_TR;
      F_s[r * ldf + a] = (r0 + r < q.nR) ? q.F[(int64_t)a * q.ldf + r0 + r] : 0.0;
    }
    {
      int r, c;
      if (q.sr == 1) { r = threadIdx.x % WG_TR; c = threadIdx.x / WG_TR; }
      else { c = threadIdx.x % WG_CC; r = threadIdx.x / WG_CC; }
      double v = 0.0;
      if (r0 + r < q.nR && c0 + c < q.nT) v = q.W[(r0 + r) * q.sr + (c0 + c) * q.sc];
      w_s[r * WG_CC + c] = v;
    }
    __syncthreads();
    wgram_tile<PT>(F_s, ldf, w_s, tab_a, tab_b, T, WG_TR, acc);
  }
#pragma unroll
  for (int pi = 0; pi < PT; ++pi) {
    const int idx = threadIdx.x + pi * WG_NT;
    if (idx >= T) continue;
#pragma unroll
    for (int c = 0; c < WG_CC; ++c) {
      if (c0 + c >= q.nT) continue;
      double* dst = q.G + (c0 + c) * T + idx;
      *dst = (q.accumulate ? *dst : 0.0) + acc[pi][c];
    }
  }
}

// one warp per matrix: se[c, m] = sqrt(diag((G_c + EPS I)^{-1}))_m ; out is nT x M, ld ldo
__global__ void __launch_bounds__(256)
se_invert_kernel(int64_t nT, int M, const double* __restrict__ G, double EPS, double* __restrict__ out,
                 int64_t ldo, unsigned long long* __restrict__ n_singular, int warps_per_cta) {
  extern __shared__ __align__(16) double smd[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= warps_per_cta) return;
  const int T = tri_count(M);
  double* A = smd + (size_t)warp * (M * M + M);
  double* dg = A + M * M;
  const int64_t c = (int64_t)blockIdx.x * warps_per_cta + warp;
  if (c >= nT) return;
  const double* g = G + c * T;
  for (int idx = lane; idx < T; idx += 32) {
    int b = (int)((sqrtf(8.0f * idx + 1.0f) - 1.0f) * 0.5f);
    while ((b + 1) * (b + 2) / 2 <= idx) ++b;
    while (b * (b + 1) / 2 > idx) --b;
    const int a = idx - b * (b + 1) / 2;
    const double v = g[idx] + (a == b ? EPS : 0.0);
    A[b * M + a] = v;      // lower triangle is what the Cholesky reads (row b >= col a)
    A[a * M + b] = v;
  }
  __syncwarp();
  const bool ok = warp_cholesky(A, M, lane);
  if (ok) warp_inv_diag(A, M, lane, dg);
  __syncwarp();
  for (int m = lane; m < M; m += 32) out[(int64_t)m * ldo + c] = ok ? sqrt(dg[m]) : nan("");
  if (!ok && lane == 0) atomicAdd(n_singular, 1ull);
}

// W chunk from factors (extension, W == NULL): W[i, jl] = exp(alpha_i,b(j) + beta_j + (Xu Xv^T)_ij)
__global__ void se_eval_w_kernel(int64_t I, int64_t j0, int64_t jc, int Mx, const double* __restrict__ Xu,
                                 const double* __restrict__ Xv, int64_t J, const double* __restrict__ alpha,
                                 const double* __restrict__ beta, const int32_t* __restrict__ batch,
                                 double* __restrict__ W) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= I * jc) return;
  const int64_t i = idx % I, jl = idx / I, j = j0 + jl;
  const int kb = batch ? batch[j] - 1 : 0;
  double x = 0.0;
  for (int m = 0; m < Mx; ++m) x += Xu[(int64_t)m * I + i] * Xv[(int64_t)m * J + j];
  W[idx] = exp(alpha[(int64_t)kb * I + i] + x) * exp(beta[j]);
}

template <typename K>
static void launch_pt(int PT, K&& f) {
  switch (PT) {
    case 1: f(std::integral_constant<int, 1>()); break;
    case 2: f(std::integral_constant<int, 2>()); break;
    case 3: case 4: f(std::integral_constant<int, 4>()); break;
    case 5: case 6: f(std::integral_constant<int, 6>()); break;
    default: f(std::integral_constant<int, 9>()); break;
  }
}

static void se_gram(Ctx& ctx, const SeGramParams& q) {
  const int T = tri_count(q.p);
  const int PT = (int)ceil_div(T, WG_NT);
  SCGBM_REQUIRE(PT <= 9, "get.se: M too large (M <= 64)");
  const size_t smem = (size_t)(WG_TR * (q.p + 1) + WG_TR * WG_CC) * sizeof(double) + 2 * (size_t)T * sizeof(unsigned short) + 16;
  ProfScope ps(ctx.prof, SCGBM_K_SE, 1);
  const unsigned grid = (unsigned)ceil_div(q.nT, WG_CC);
  launch_pt(PT, [&](auto pt) {
    constexpr int P = decltype(pt)::value;
    if (smem > 48 * 1024)
      SCGBM_CUDA(cudaFuncSetAttribute(se_gram_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    se_gram_kernel<P><<<grid, WG_NT, smem, ctx.stream>>>(q);
  });
  SCGBM_LAUNCH_CHECK();
}

void se_invert(Ctx& ctx, int64_t nT, int M, const double* G, double EPS, double* out, int64_t ldo,
               unsigned long long* n_singular) {
  SCGBM_REQUIRE(M <= WG_PMAX, "get.se: M must be <= 64");
  const size_t per_warp = (size_t)(M * M + M) * sizeof(double);
  int wpc = (int)std::min<size_t>(8, (200 * 1024) / per_warp);
  SCGBM_REQUIRE(wpc >= 1, "get.se: M too large for the shared-memory Cholesky");
  const size_t smem = per_warp * wpc;
  if (smem > 48 * 1024)
    SCGBM_CUDA(cudaFuncSetAttribute(se_invert_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope ps(ctx.prof, SCGBM_K_SE, 1);
  se_invert_kernel<<<(unsigned)ceil_div(nT, wpc), 256, smem, ctx.stream>>>(nT, M, G, EPS, out, ldo, n_singular, wpc);
  SCGBM_LAUNCH_CHECK();
}

void get_se_impl(const scgbm_se_args* a, scgbm_se_out* out) {
  SCGBM_REQUIRE(a && out && a->U && a->scores && out->se_scores && out->se_loadings, "NULL args");
  SCGBM_REQUIRE(a->I >= 1 && a->J >= 1 && a->M >= 1, "bad shape");
  SCGBM_REQUIRE(a->W || (a->alpha && a->beta && (a->Mx == 0 || (a->Xu && a->Xv))), "get.se needs W or (alpha, beta, X factors)");
  Ctx ctx(a->device, reinterpret_cast<Comm*>(a->comm));
  ctx.prof.timing = a->profile != 0;
  cudaEvent_t ev0, ev1;
  SCGBM_CUDA(cudaEventCreate(&ev0)); SCGBM_CUDA(cudaEventCreate(&ev1));
  SCGBM_CUDA(cudaEventRecord(ev0, ctx.stream));
  const int64_t I = a->I, J = a->J;
  const int M = a->M, T = tri_count(M);
  DevBuf<double> U(I * M), S(J * M), GI(I * T), seI(I * M);
  DevBuf<unsigned long long> nsing(1);
  nsing.zero(ctx.stream);
  GI.zero(ctx.stream);
  SCGBM_CUDA(cudaMemcpyAsync(U.get(), a->U, I * M * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  SCGBM_CUDA(cudaMemcpyAsync(S.get(), a->scores, J * M * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  ctx.prof.h2d_bytes += (double)(I + J) * M * sizeof(double);
  DevBuf<double> Xu, Xv, al, be;
  DevBuf<int32_t> bt;
  const int nbatch = std::max(1, a->nbatch);
  if (!a->W) {
    if (a->Mx > 0) {
      Xu.alloc(I * a->Mx); Xv.alloc(J * a->Mx);
      SCGBM_CUDA(cudaMemcpyAsync(Xu.get(), a->Xu, I * a->Mx * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
      SCGBM_CUDA(cudaMemcpyAsync(Xv.get(), a->Xv, J * a->Mx * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
    }
    al.alloc(I * nbatch); be.alloc(J);
    SCGBM_CUDA(cudaMemcpyAsync(al.get(), a->alpha, I * nbatch * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
    SCGBM_CUDA(cudaMemcpyAsync(be.get(), a->beta, J * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
    if (a->batch && nbatch > 1) {
      bt.alloc(J);
      SCGBM_CUDA(cudaMemcpyAsync(bt.get(), a->batch, J * sizeof(int32_t), cudaMemcpyHostToDevice, ctx.stream));
    }
  }
  int64_t jc_max = std::max<int64_t>(WG_CC, (int64_t)(256ll << 20) / (I * (int64_t)sizeof(double)));
  jc_max = std::min(jc_max, J);
  DevBuf<double> Wd(I * jc_max), GJ(jc_max * T), seJ(jc_max * M);
  for (int64_t j0 = 0; j0 < J; j0 += jc_max) {
    const int64_t jc = std::min(jc_max, J - j0);
    if (a->W) {
      SCGBM_CUDA(cudaMemcpyAsync(Wd.get(), a->W + j0 * I, I * jc * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
      ctx.prof.h2d_bytes += (double)I * jc * sizeof(double);
    } else {
      ProfScope ps(ctx.prof, SCGBM_K_SE, 1);
      se_eval_w_kernel<<<(unsigned)ceil_div(I * jc, 256), 256, 0, ctx.stream>>>(
          I, j0, jc, a->Mx, Xu.get(), Xv.get(), J, al.get(), be.get(), bt.get(), Wd.get());
      SCGBM_LAUNCH_CHECK();
    }
    // cells of this chunk: reduce over genes (contiguous in W)
    SeGramParams qj{jc, I, Wd.get(), 1, I, U.get(), I, M, GJ.get(), 0};
    se_gram(ctx, qj);
    se_invert(ctx, jc, M, GJ.get(), a->EPS, seJ.get(), jc, nsing.get());
    SCGBM_CUDA(cudaMemcpy2DAsync(out->se_scores + j0, J * sizeof(double), seJ.get(), jc * sizeof(double),
                                 jc * sizeof(double), M, cudaMemcpyDeviceToHost, ctx.stream));
    // genes: accumulate over this chunk's cells (stride I in W)
    SeGramParams qi{I, jc, Wd.get(), I, 1, S.get() + j0, J, M, GI.get(), 1};
    se_gram(ctx, qi);
    ctx.sync();   // Wd is reused by the next chunk's H2D
  }
  ctx.allreduce_sum(GI.get(), I * T);
  se_invert(ctx, I, M, GI.get(), a->EPS, seI.get(), I, nsing.get());
  SCGBM_CUDA(cudaMemcpyAsync(out->se_loadings, seI.get(), I * M * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  unsigned long long hs = 0;
  SCGBM_CUDA(cudaMemcpyAsync(&hs, nsing.get(), sizeof(hs), cudaMemcpyDeviceToHost, ctx.stream));
  SCGBM_CUDA(cudaEventRecord(ev1, ctx.stream));
  ctx.sync();
  ctx.prof.d2h_bytes += (double)(I + J) * M * sizeof(double);
  float tot = 0;
  cudaEventElapsedTime(&tot, ev0, ev1);
  cudaEventDestroy(ev0); cudaEventDestroy(ev1);
  memset(&out->prof, 0, sizeof(out->prof));
  ctx.prof.fill(&out->prof);
  out->prof.total_ms = tot;
  out->n_singular = (int64_t)hs;
  if (hs > 0) {
    char b[200];
    snprintf(b, sizeof(b), "get.se: %lld information matrices are not positive definite (R's solve() would stop; try EPS > 0)", (long long)hs);
    throw Error(SCGBM_ERR_NUMERIC, b);
  }
}

}  // namespace scgbm
