// proj_tc.cu -- K9 on the tensor cores: the per-cell Poisson GLMs of gbm.sc(subset = ...) (R/fit.R:248-264) in
// lock step over ALL cells, every I x J operation on tcgen05.
//
// fastglm(x = [1 | U], y = Y[, j], offset = alpha, family = poisson()) is Newton's method on a concave likelihood
// (IRLS with the canonical link IS Newton): with X = [1 | U] (Ik x p, p = M + 1), eta_j = alpha + X theta_j,
//     gradient  g_j = X^T (y_j - mu_j)             Hessian  H_j = X^T diag(mu_j) X           theta_j += H_j^-1 g_j.
// Over all cells at once both are tall-skinny products with the residual R = Y - W of the CURRENT coefficients
// against the Khatri-Rao expansion K of the design, K[i, (a,b)] = X[i,a] X[i,b] (a <= b; T = p (p + 1) / 2 columns,
// the pairs (0, b) being X itself because X[, 0] = 1):
//     R^T K [j, (0,b)] = g_j[b]                    Y^T K - R^T K = W^T K = H_j            (one row per cell)
// and W = exp(alpha_i + theta_j0 + U theta_j) is exactly what the fit's fused likelihood pass evaluates (f_i = e^alpha,
// g_j = e^theta_j0, low-rank term U x scores).  So one IRLS sweep = ceil((T + 1) / 32) launches of fused_tc_kernel's
// product rider in `product_only` form (eta_tc.cu: eta tile by UMMA, exp, residual split into two fp16 planes in
// shared memory, second UMMA with a 32-column block of K, fp64 drain every 128 genes; nothing I x J is written),
// followed by one warp-per-cell kernel that assembles the p x p system, solves it by Cholesky and updates theta.
// Y^T K, needed once, costs no pass of its own: at the start theta = (log(cs_j / sum_i e^alpha_i), 0, ..) makes W
// rank one, so Y^T K = R_0^T K + g_j (f^T K) with the second term summed in fp64 from the very floats the kernel used.
//
// Convergence mirrors fastglm's |dev - devold| / (|dev| + 0.1) < tol with the PREDICTED decrease of the step, the
// Newton decrement g^T H^-1 g (they agree to second order; the decrement is computed in fp64 from the products and is
// free of the ~1e-7 relative noise a recomputed deviance would carry at 22-bit eta): the step is taken and, when its
// decrement is below the tolerance, the cell is done -- coefficients after the step, standard errors from that last
// solve's X^T W X, which is what fastglm returns.  A deviance that rises (Newton overshoot far from the optimum)
// halves the step.  Cells with no counts, a failed factorisation, a non-finite deviance, and the last percent of
// slow cells go to the exact fp64 kernel (proj.cu) as a list, which also owns fastglm's failure semantics.
// Accuracy against the exact kernel (measured, 20k x 60k, M = 20): coefficients 4e-7 relative (22-bit eta,
// ex2.approx); standard errors 4e-4 relative -- "the last solve's X^T W X" is evaluated one step before the returned
// coefficients on either path, and the two paths' last steps differ.  Newton solves per cell: 3.3 on average
// against fastglm's 6.0 IRLS iterations (the null model is a better start than mu = y + 0.1 for sparse counts).
#include "proj_tc.cuh"
#include "eta_pass.cuh"
#include "wgram.cuh"

namespace scgbm {

static constexpr int kBlk = 32;        // rider block width (FusedProd::b)
static constexpr double kStragglerShare = 0.02;   // below this share of running cells the exact kernel takes the rest

__device__ __forceinline__ void pair_of(int idx, int& a, int& b) {   // idx = b (b + 1) / 2 + a, a <= b
  int bb = (int)((sqrtf(8.0f * (float)idx + 1.0f) - 1.0f) * 0.5f);
  while ((bb + 1) * (bb + 2) / 2 <= idx) ++bb;
  while (bb * (bb + 1) / 2 > idx) --bb;
  b = bb; a = idx - bb * (bb + 1) / 2;
}

// K (ldI x Tp, column-major): columns < T the pair products of X = [1 | U], column T the offset, the rest zero
__global__ void kr_build_kernel(int64_t Ik, int64_t ldI, int M, int T, int Tp, const double* __restrict__ U,
                                const double* __restrict__ off, double* __restrict__ K) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ldI * Tp) return;
  const int64_t i = idx % ldI;
  const int t = (int)(idx / ldI);
  double v = 0.0;
  if (i < Ik) {
    if (t < T) {
      int a, b;
      pair_of(t, a, b);
      const double xa = a == 0 ? 1.0 : U[(int64_t)(a - 1) * ldI + i];
      const double xb = b == 0 ? 1.0 : U[(int64_t)(b - 1) * ldI + i];
      v = xa * xb;
    } else if (t == T) {
      v = off[i];
    }
  }
  K[idx] = v;
}

// f_row = float(exp(off)) (n entries, zero beyond Ik), la_row = float(off) (n_la entries)
__global__ void stage_rows_kernel(int64_t Ik, int64_t n, int64_t n_la, const double* __restrict__ off, float* __restrict__ f,
                                  float* __restrict__ la) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool ok = i < Ik;
  f[i] = ok ? (float)exp(off[i]) : 0.0f;
  if (i < n_la) la[i] = ok ? (float)off[i] : 0.0f;
}

// fK[t] = sum_i (double) f_row[i] * K[i, t]     (one CTA per column, fixed order)
__global__ void __launch_bounds__(256) fk_kernel(int64_t Ik, int64_t ldI, const float* __restrict__ f, const double* __restrict__ K,
                                                 double* __restrict__ fK) {
  __shared__ double red[256];
  const int t = blockIdx.x;
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < Ik; i += 256) s += (double)f[i] * K[(int64_t)t * ldI + i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) fK[t] = red[0];
}

// per cell: C_j = sum_{y > 0} y log y (warp per cell)
template <typename YT>
__global__ void ylogy_kernel(int64_t Ik, int64_t J, int64_t ldY, const YT* __restrict__ Y, double* __restrict__ Cj) {
  const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (j >= J) return;
  const YT* col = Y + j * ldY;
  double s = 0.0;
  for (int64_t i = lane; i < Ik; i += 32) {
    const double y = (double)col[i];
    if (y > 0.0) s += y * log(y);
  }
  s = warp_sum(s);
  if (lane == 0) Cj[j] = s;
}

// theta_0 = (log(cs_j / sum_i e^off_i), 0, ...): the null model's MLE; cells without counts go to the exact kernel
__global__ void irls_init_kernel(int64_t J, int64_t ldJ, int p, const double* __restrict__ cs, const float* __restrict__ f_row,
                                 int64_t Ik, double* __restrict__ theta, double* __restrict__ theta_prev, int* __restrict__ state,
                                 int* __restrict__ iters, int* __restrict__ halv, double* __restrict__ dev_prev) {
  __shared__ double red[256];
  double s = 0.0;                                  // every CTA re-sums the 20k offsets (L2-resident): no extra launch
  for (int64_t i = threadIdx.x; i < Ik; i += 256) s += (double)f_row[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
    __syncthreads();
  }
  const double tot = red[0];
  const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (j >= ldJ) return;
  const bool real = j < J;
  const double c = real ? cs[j] : 0.0;
  for (int a = 0; a < p; ++a) { theta[(int64_t)a * ldJ + j] = 0.0; theta_prev[(int64_t)a * ldJ + j] = 0.0; }
  if (real && c > 0.0) theta[j] = log(c / tot);
  state[j] = !real ? 1 : (c > 0.0 ? 0 : 3);
  iters[j] = 0; halv[j] = 0; dev_prev[j] = 0.0;
}

// g_col = float(exp(theta_0)) (zero beyond J, n = ldJ + 128), lb_col = float(theta_0)
__global__ void stage_cols_kernel(int64_t J, int64_t n, int64_t ldJ, const double* __restrict__ theta0, float* __restrict__ g,
                                  float* __restrict__ lb) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const bool ok = j < J;
  g[j] = ok ? (float)exp(theta0[j]) : 0.0f;
  if (j < ldJ) lb[j] = ok ? (float)theta0[j] : 0.0f;
}

// Y^T K = R_0^T K + g_j (f^T K): W_0 = f_i g_j is what the kernel multiplied (ex2(0) = 1 exactly)
__global__ void ytk_build_kernel(int64_t J, int64_t ldJ, int Tp, const double* __restrict__ RtK, const float* __restrict__ g_col,
                                 const double* __restrict__ fK, double* __restrict__ YtK) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ldJ * Tp) return;
  const int64_t j = idx % ldJ;
  const int t = (int)(idx / ldJ);
  YtK[idx] = j < J ? RtK[idx] + (double)g_col[j] * fK[t] : 0.0;
}

struct IrlsTcParams {
  int64_t J, ldJ;
  int p, T, it;
  double tol;
  const double* YtK; const double* RtK; const double* Cj;
  double* theta; double* theta_prev; double* dev_prev; double* se;
  int* state;        // 0 running, 1 converged, 3 hand over to the exact kernel
  int* iters; int* halv;
  int* n_active;
};

constexpr int IRLS_WPB = 4;
__global__ void __launch_bounds__(IRLS_WPB * 32) irls_update_kernel(const IrlsTcParams q) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int p = q.p, wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t j = (int64_t)blockIdx.x * IRLS_WPB + wib;
  if (j >= q.J || q.state[j] != 0) return;               // warp-uniform; only warp-level synchronisation below
  double* A = reinterpret_cast<double*>(smraw) + (size_t)wib * (p * p + 3 * p);
  double* g = A + p * p;
  double* dl = g + p;
  double* dg = dl + p;
  const int64_t ldJ = q.ldJ;
  for (int idx = lane; idx < q.T; idx += 32) {
    int a, b;
    pair_of(idx, a, b);
    const double r = q.RtK[(int64_t)idx * ldJ + j];
    const double h = q.YtK[(int64_t)idx * ldJ + j] - r;   // (W^T K)[j, (a,b)]
    A[b * p + a] = h; A[a * p + b] = h;
    if (a == 0) g[b] = r;                                 // X^T (y - mu)
  }
  __syncwarp();
  // deviance of the current coefficients: 2 [sum y log y - sum y eta - sum (y - mu)]
  double s = 0.0;
  for (int b = lane; b < p; b += 32) s += q.theta[(int64_t)b * ldJ + j] * q.YtK[(int64_t)(b * (b + 1) / 2) * ldJ + j];
  s = warp_sum(s);
  const double dev = 2.0 * (q.Cj[j] - (s + q.YtK[(int64_t)q.T * ldJ + j]) - g[0]);
  const double devp = q.dev_prev[j];
  auto still_running = [&]() { if (lane == 0) atomicAdd(q.n_active, 1); };
  if (!isfinite(dev)) { if (lane == 0) q.state[j] = 3; return; }
  if (q.iters[j] > 0 && dev > devp + 1e-6 * (fabs(devp) + 0.1)) {
    // the last step overshot: go back half way (theta_prev / dev_prev keep describing the point before the step)
    if (q.halv[j] >= 12) { if (lane == 0) q.state[j] = 3; return; }
    for (int a = lane; a < p; a += 32)
      q.theta[(int64_t)a * ldJ + j] = 0.5 * (q.theta[(int64_t)a * ldJ + j] + q.theta_prev[(int64_t)a * ldJ + j]);
    if (lane == 0) q.halv[j] += 1;
    still_running();
    return;
  }
  for (int a = lane; a < p; a += 32) dl[a] = g[a];
  __syncwarp();
  if (!warp_cholesky(A, p, lane)) { if (lane == 0) q.state[j] = 3; return; }
  warp_chol_solve(A, dl, p, lane);
  warp_inv_diag(A, p, lane, dg);
  double dec = 0.0;                                       // Newton decrement g^T H^-1 g = predicted drop of the deviance
  bool finite = true;
  for (int a = lane; a < p; a += 32) { dec += g[a] * dl[a]; finite &= isfinite(dl[a]) && isfinite(dg[a]); }
  dec = warp_sum(dec);
  finite = __all_sync(0xffffffffu, finite);
  if (!finite || !(dec >= 0.0)) { if (lane == 0) q.state[j] = 3; return; }
  for (int a = lane; a < p; a += 32) {
    const double th = q.theta[(int64_t)a * ldJ + j];
    q.theta_prev[(int64_t)a * ldJ + j] = th;
    q.theta[(int64_t)a * ldJ + j] = th + dl[a];
    q.se[(int64_t)a * ldJ + j] = sqrt(dg[a]);             // SE of the last solve's X^T W X (fastglm's save_se)
  }
  if (lane == 0) {
    q.dev_prev[j] = dev;
    q.iters[j] += 1;
    q.halv[j] = 0;
  }
  if (dec / (fabs(dev) + 0.1) < q.tol) { if (lane == 0) q.state[j] = 1; }
  else still_running();
}

// results of the converged cells in the exact kernel's layout (J x p, ld J); stragglers are listed for it
__global__ void irls_finish_kernel(int64_t J, int64_t ldJ, int p, const double* __restrict__ theta, const double* __restrict__ se_in,
                                   const int* __restrict__ state, const int* __restrict__ iters_in, double* __restrict__ coef,
                                   double* __restrict__ se, int8_t* __restrict__ fail, int32_t* __restrict__ iters,
                                   int64_t* __restrict__ list, unsigned long long* __restrict__ nlist) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= J) return;
  if (state[j] == 1) {
    for (int a = 0; a < p; ++a) {
      coef[(int64_t)a * J + j] = theta[(int64_t)a * ldJ + j];
      se[(int64_t)a * J + j] = se_in[(int64_t)a * ldJ + j];
    }
    fail[j] = 0;
    if (iters) iters[j] = iters_in[j];
  } else {
    list[atomicAdd(nlist, 1ull)] = j;                    // order does not matter: every cell is fitted on its own
  }
}

bool proj_tc_supported(Ctx& ctx, const Counts& Y, int M) {
  if (!(Y.storage == SCGBM_STORE_U8 || Y.storage == SCGBM_STORE_U16) || M < 1 || M > 64) return false;
  // the rider's shared-memory budget at this M (one K slab per 64 halves of the three-section operand)
  const int nslab = (int)round_up(3 * round_up(M, 16), 64) / 64;
  const int fixed = nslab * 128 * 128 + 2 * nslab * 64 * 128 + 2 * (2 * 128 * 128) + 1024 + 640 + 4 * 64 * 4 + 4 * (64 * 128);
  (void)ctx;
  return (232448 - fixed) / (128 * 64 * 2) >= 2;
}

int64_t project_tc(Ctx& ctx, const Counts& Y, int64_t Ik, int M, const double* U_dev, const double* off_dev, double tol,
                   int max_sweeps, double* coef, double* se, int8_t* fail, int32_t* iters, DevBuf<int64_t>& list,
                   ProjTcStats* stats) {
  const int64_t J = Y.J, ldI = Y.ld, ldJ = round_up(J, kRowPad);
  const int p = M + 1, T = tri_count(p), Tp = (int)round_up(T + 1, kBlk), nblk = Tp / kBlk;
  SCGBM_REQUIRE(Y.I == Ik && proj_tc_supported(ctx, Y, M), "project_tc: unsupported input");
  cudaStream_t st = ctx.stream;
#define G1(n) (unsigned)ceil_div((n), 256), 256, 0, st

  // design, its Khatri-Rao expansion (+ the offset column), the fp16 operand planes of every 32-column block
  DevBuf<double> Upad(ldI * M), offpad(ldI), K(ldI * Tp), fK(Tp);
  Upad.zero(st); offpad.zero(st);
  SCGBM_CUDA(cudaMemcpy2DAsync(Upad.get(), ldI * sizeof(double), U_dev, Ik * sizeof(double), Ik * sizeof(double), M,
                               cudaMemcpyDeviceToDevice, st));
  SCGBM_CUDA(cudaMemcpyAsync(offpad.get(), off_dev, Ik * sizeof(double), cudaMemcpyDeviceToDevice, st));
  DevBuf<float> f_row(ldI + 128), la_row(ldI), g_col(ldJ + 128), lb_col(ldJ);
  {
    ProfScope ps(ctx.prof, SCGBM_K_MISC, 2);
    kr_build_kernel<<<G1(ldI * Tp)>>>(Ik, ldI, M, T, Tp, Upad.get(), offpad.get(), K.get());
    stage_rows_kernel<<<G1(ldI + 128)>>>(Ik, ldI + 128, ldI, offpad.get(), f_row.get(), la_row.get());
    SCGBM_LAUNCH_CHECK();
  }
  {
    ProfScope ps(ctx.prof, SCGBM_K_MISC, 1);
    fk_kernel<<<Tp, 256, 0, st>>>(Ik, ldI, f_row.get(), K.get(), fK.get());
    SCGBM_LAUNCH_CHECK();
  }
  std::vector<ProdWs> blk(nblk);
  for (int c = 0; c < nblk; ++c) split_operand(ctx, blk[c], ldI, kBlk, kBlk, ldI, K.get() + (int64_t)c * kBlk * ldI, ldI);
  Factor16 fa, fb;
  stage_factor16(ctx, fa, Ik, ldI, M, 0, Upad.get(), nullptr, nullptr);

  // per-cell state
  DevBuf<double> theta(ldJ * p), theta_prev(ldJ * p), se_w(ldJ * p), dev_prev(ldJ), Cj(ldJ), YtK(ldJ * Tp), RtK(ldJ * Tp), scal(8);
  DevBuf<int> state(ldJ), it_w(ldJ), halv(ldJ), n_active(1);
  se_w.zero(st); Cj.zero(st); RtK.zero(st); scal.zero(st);
  {
    ProfScope ps(ctx.prof, SCGBM_K_PROJ, 2);
    if (Y.storage == SCGBM_STORE_U8)
      ylogy_kernel<uint8_t><<<G1(J * 32)>>>(Ik, J, Y.ld, reinterpret_cast<const uint8_t*>(Y.data), Cj.get());
    else
      ylogy_kernel<uint16_t><<<G1(J * 32)>>>(Ik, J, Y.ld, reinterpret_cast<const uint16_t*>(Y.data), Cj.get());
    irls_init_kernel<<<(unsigned)ceil_div(ldJ, 256), 256, 0, st>>>(J, ldJ, p, Y.colsum.get(), f_row.get(), Ik, theta.get(),
                                                                  theta_prev.get(), state.get(), it_w.get(), halv.get(), dev_prev.get());
    SCGBM_LAUNCH_CHECK();
  }

  // W is not clipped in a GLM; the planes' scale only has to keep |Y - W| inside fp16 for any W the iteration can
  // reasonably visit (a W beyond the cap is clipped there, the deviance rises, the step is halved)
  const float cap = (float)std::max(4096.0, 64.0 * (Y.maxY + 1.0));
  ResidBuf resid;
  resid.hi = reinterpret_cast<uint16_t*>(Y.data);      // product_only: never written; the maps want a valid address
  resid.lo = reinterpret_cast<uint16_t*>(Y.data);
  resid.scale = resid_scale_for((double)cap);
  PassGeom geom;
  geom.I = Ik; geom.J = J; geom.ldI = ldI; geom.ldJ = ldJ; geom.nbatch = 1; geom.batch = nullptr;
  EtaWs ews;
  XDesc x;
  x.M = M; x.clamp = 0; x.A16 = &fa; x.B16 = &fb;

  const size_t smem_upd = (size_t)IRLS_WPB * (p * p + 3 * p) * sizeof(double);
  if (smem_upd > 48 * 1024)
    SCGBM_CUDA(cudaFuncSetAttribute(irls_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_upd));
  int sweeps = 0;
  int64_t launches = 0;
  for (int it = 0; it < max_sweeps; ++it) {
    {
      ProfScope ps(ctx.prof, SCGBM_K_MISC, 1);
      stage_cols_kernel<<<G1(ldJ + 128)>>>(J, ldJ + 128, ldJ, theta.get(), g_col.get(), lb_col.get());
      SCGBM_LAUNCH_CHECK();
    }
    stage_factor16(ctx, fb, J, ldJ, M, 1, theta.get() + ldJ, nullptr, nullptr);
    for (int c = 0; c < nblk; ++c) {
      FusedProd fp;
      fp.op16 = blk[c].op16.get(); fp.ldk = ldI; fp.op_inv = blk[c].opmeta.get() + 1; fp.b = kBlk;
      fp.out = RtK.get() + (int64_t)c * kBlk * ldJ; fp.ldo = ldJ; fp.product_only = true;
      pass_fused_tc(ctx, ews, geom, x, Y, f_row.get(), g_col.get(), la_row.get(), lb_col.get(), offpad.get(), theta.get(),
                    Y.rowsum.get(), cap, resid, scal.get(), &fp);
      ++launches;
    }
    ProfScope ps(ctx.prof, SCGBM_K_PROJ, 2);
    if (it == 0) {
      ytk_build_kernel<<<G1(ldJ * Tp)>>>(J, ldJ, Tp, RtK.get(), g_col.get(), fK.get(), YtK.get());
      SCGBM_LAUNCH_CHECK();
    }
    n_active.zero(st);
    IrlsTcParams q;
    q.J = J; q.ldJ = ldJ; q.p = p; q.T = T; q.it = it; q.tol = tol;
    q.YtK = YtK.get(); q.RtK = RtK.get(); q.Cj = Cj.get();
    q.theta = theta.get(); q.theta_prev = theta_prev.get(); q.dev_prev = dev_prev.get(); q.se = se_w.get();
    q.state = state.get(); q.iters = it_w.get(); q.halv = halv.get(); q.n_active = n_active.get();
    irls_update_kernel<<<(unsigned)ceil_div(J, IRLS_WPB), IRLS_WPB * 32, smem_upd, st>>>(q);
    SCGBM_LAUNCH_CHECK();
    int h = 0;
    SCGBM_CUDA(cudaMemcpyAsync(&h, n_active.get(), sizeof(int), cudaMemcpyDeviceToHost, st));
    ctx.sync();
    ++sweeps;
    static const bool trace = getenv("SCGBM_TRACE") != nullptr;
    if (trace) fprintf(stderr, "[scgbm-trace] projection sweep %d: %d of %lld cells still running\n", it, h, (long long)J);
    if (h == 0) break;
    // Newton from the null model converges quadratically: after 3-4 sweeps a percent or so of the cells is left,
    // and a further lock-step sweep over ALL cells (~0.25 us per cell) costs more than the exact kernel on the
    // remainder (~11 us per straggler).  Measured at 20k x 60k: 23 % / 1.6 % / 0.2 % running after sweeps 2 / 3 / 4.
    if (it >= 2 && (double)h < kStragglerShare * (double)J) break;
  }
  // hand the results over; whoever did not converge is listed for the exact kernel
  if (list.n < J) list.alloc(J);
  DevBuf<unsigned long long> nlist(1);
  nlist.zero(st);
  {
    ProfScope ps(ctx.prof, SCGBM_K_PROJ, 1);
    irls_finish_kernel<<<G1(J)>>>(J, ldJ, p, theta.get(), se_w.get(), state.get(), it_w.get(), coef, se, fail, iters,
                                  list.get(), nlist.get());
    SCGBM_LAUNCH_CHECK();
  }
  unsigned long long n = 0;
  SCGBM_CUDA(cudaMemcpyAsync(&n, nlist.get(), sizeof(n), cudaMemcpyDeviceToHost, st));
  ctx.sync();
  if (stats) { stats->sweeps = sweeps; stats->rider_launches = launches; stats->stragglers = (int64_t)n; stats->blocks = nblk; }
#undef G1
  return (int64_t)n;
}

}  // namespace scgbm
