// simulate.cu -- the reference's generators on the device:
//   scgbm_sim_counts   Y ~ Poisson / NB (exp(alpha_i + beta_j + (U D V^T)_ij)) for GIVEN factors (R/simulate.R:37-38)
//   scgbm_sim_data     simData / simNull / simNullNB end to end: Haar U, V by Cholesky-QR of Gaussian blocks,
//                      D, alpha, beta and the counts, all drawn on the GPU        (R/simulate.R:24-48, 51-57, 60-65)
//   scgbm_cci_perturb  V + N(0, se_V^2), `reps` replicates in one launch           (R/cci.R:69-71, :57)
// One counter-based stream (philox.cuh): element e of draw family f uses the Philox counter e under the key
// seed + f * GOLDEN, so results do not depend on sharding or launch geometry.
#include "dense_small.cuh"
#include "philox.cuh"

namespace scgbm {

// draw families (keys derived from the user's seed); the counts keep the bare seed (scgbm_sim_counts' stream)
constexpr uint64_t kKeyStep = 0x9E3779B97F4A7C15ull;
enum { FAM_U = 1, FAM_V = 2, FAM_ALPHA = 3, FAM_BETA = 4, FAM_MU = 5, FAM_CCI = 6 };
__host__ __device__ inline uint64_t fam_key(uint64_t seed, int fam) { return seed + kKeyStep * (uint64_t)fam; }

template <typename Tout>
__global__ void __launch_bounds__(256)
sim_counts_kernel(int64_t I, int64_t ldv, int64_t ld, int64_t j_offset, int d, const float* __restrict__ Ud,
                  const float* __restrict__ V, const float* __restrict__ alpha, const float* __restrict__ beta,
                  double nb_theta, uint64_t seed, Tout* __restrict__ Y, unsigned long long* __restrict__ n_sat) {
  extern __shared__ float s_v[];   // d floats: V row of this column
  const int64_t j = blockIdx.y;
  for (int m = threadIdx.x; m < d; m += blockDim.x) s_v[m] = V[(int64_t)m * ldv + j];
  __syncthreads();
  const float bj = beta[j];
  unsigned long long sat = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ld; i += (int64_t)gridDim.x * blockDim.x) {
    Tout outv = (Tout)0;
    if (i < I) {
      float x = alpha[i] + bj;
      for (int m = 0; m < d; ++m) x = fmaf(Ud[(int64_t)m * I + i], s_v[m], x);
      double lam = exp((double)x);
      Philox rng(seed, (uint64_t)(j_offset + j) * (uint64_t)I + (uint64_t)i);
      if (nb_theta > 0.0) lam = rng.gamma(nb_theta) * (lam / nb_theta);
      double y = rng.poisson(lam);
      if (y > SimLimit<Tout>::max) { y = SimLimit<Tout>::max; ++sat; }
      outv = (Tout)y;
    }
    Y[j * ld + i] = outv;
  }
  if (sat) atomicAdd(n_sat, sat);
}

__global__ void f64_to_f32_scaled(int64_t n, int q, const double* __restrict__ src, int64_t lds,
                                  const double* __restrict__ colscale, float* __restrict__ dst, int64_t ldd) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * q) return;
  const int64_t r = idx % n;
  const int c = (int)(idx / n);
  dst[(int64_t)c * ldd + r] = (float)(src[(int64_t)c * lds + r] * (colscale ? colscale[c] : 1.0));
}

// counts for device-resident fp32 factors: Ud = U diag(D) (I x d, ld I), Vf (rows = this call's cells, ld ldv)
static int64_t sim_counts_dev(Ctx& ctx, int64_t I, int64_t J, int64_t ld, int64_t j_offset, int d, const float* Ud,
                              const float* Vf, int64_t ldv, const float* al, const float* be, double nb_theta,
                              uint64_t seed, int out_dtype, void* Y_dev) {
  DevBuf<unsigned long long> nsat(1);
  nsat.zero(ctx.stream);
  const unsigned gx = (unsigned)std::min<int64_t>(ceil_div(ld, 256), 64);
  const int64_t slab = 65535;              // blockIdx.y limit: loop over column slabs
  const int es = out_dtype == SCGBM_U8 ? 1 : (out_dtype == SCGBM_U16 ? 2 : 4);
  for (int64_t j0 = 0; j0 < J; j0 += slab) {
    const int64_t jc = std::min(slab, J - j0);
    dim3 g(gx, (unsigned)jc);
    const size_t sm = std::max(d, 1) * sizeof(float);
    void* yp = (uint8_t*)Y_dev + (size_t)j0 * ld * es;
#define SIM_LAUNCH(T)                                                                                         \
    sim_counts_kernel<T><<<g, 256, sm, ctx.stream>>>(I, ldv, ld, j_offset + j0, d, Ud, Vf + j0, al, be + j0,   \
                                                     nb_theta, seed, (T*)yp, nsat.get())
    switch (out_dtype) {
      case SCGBM_U8: SIM_LAUNCH(uint8_t); break;
      case SCGBM_U16: SIM_LAUNCH(uint16_t); break;
      case SCGBM_F32: SIM_LAUNCH(float); break;
      case SCGBM_I32: SIM_LAUNCH(int32_t); break;
      default: throw Error(SCGBM_ERR_INVALID, "sim: unsupported out_dtype");
    }
#undef SIM_LAUNCH
    SCGBM_LAUNCH_CHECK();
  }
  unsigned long long h = 0;
  SCGBM_CUDA(cudaMemcpyAsync(&h, nsat.get(), sizeof(h), cudaMemcpyDeviceToHost, ctx.stream));
  ctx.sync();
  return (int64_t)h;
}

void sim_counts(const scgbm_sim_args* a, void* Y_dev, int64_t* n_saturated) {
  SCGBM_REQUIRE(a->I >= 1 && a->J >= 1 && a->ld >= a->I, "sim: bad shape");
  SCGBM_REQUIRE(a->d >= 0 && a->d <= 1024, "sim: bad d");
  SCGBM_REQUIRE(a->alpha && a->beta, "sim: alpha/beta required");
  Ctx ctx(a->device, nullptr);
  const int64_t I = a->I, J = a->J;
  const int d = a->d;
  DevBuf<double> tmp(std::max<int64_t>(std::max(I, J) * std::max(d, 1), 1)), Dd(std::max(d, 1));
  DevBuf<float> Ud(std::max<int64_t>(I * d, 1)), Vf(std::max<int64_t>(J * d, 1)), al(I), be(J);
  auto conv = [&](const double* host, int64_t n, int q, const double* colscale_dev, float* dst) {
    SCGBM_CUDA(cudaMemcpyAsync(tmp.get(), host, n * q * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
    f64_to_f32_scaled<<<(unsigned)ceil_div(n * q, 256), 256, 0, ctx.stream>>>(n, q, tmp.get(), n, colscale_dev, dst, n);
    SCGBM_LAUNCH_CHECK();
  };
  if (d > 0) {
    SCGBM_REQUIRE(a->U && a->V && a->D, "sim: U, V, D required when d > 0");
    SCGBM_CUDA(cudaMemcpyAsync(Dd.get(), a->D, d * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
    conv(a->U, I, d, Dd.get(), Ud.get());
    conv(a->V, J, d, nullptr, Vf.get());
  }
  conv(a->alpha, I, 1, nullptr, al.get());
  conv(a->beta, J, 1, nullptr, be.get());
  const int64_t ns = sim_counts_dev(ctx, I, J, a->ld, a->j_offset, d, Ud.get(), Vf.get(), J, al.get(), be.get(),
                                    a->nb_theta, a->seed, a->out_dtype, Y_dev);
  if (n_saturated) *n_saturated = ns;
}

// ---------------------------------------------------------------------------------
// simData / simNull / simNullNB
// ---------------------------------------------------------------------------------
// out[c * ld + r] = mean + sd * z(key; counter (row_offset + r) * q + c),  r < n, c < q
__global__ void normal_fill_kernel(int64_t n, int q, int64_t ld, uint64_t key, int64_t row_offset, double mean,
                                   double sd, double* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * q) return;
  const int64_t r = idx % n;
  const int c = (int)(idx / n);
  Philox rng(key, (uint64_t)(row_offset + r) * (uint64_t)q + (uint64_t)c);
  out[(int64_t)c * ld + r] = mean + sd * rng.normal();
}

// out[r] = log(lo + (hi - lo) * u(key; counter r))     (simNull: mus <- runif(I, 1, 10); alpha = log mu)
__global__ void log_uniform_fill_kernel(int64_t n, uint64_t key, double lo, double hi, double* __restrict__ out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  Philox rng(key, (uint64_t)r);
  out[r] = log(lo + (hi - lo) * rng.uniform());
}

// U[1, m] < 0 -> flip column m of U and of V     (R/simulate.R:28-33)
__global__ void sign_fix_kernel(int64_t I, int64_t J, int d, double* __restrict__ U, double* __restrict__ V) {
  const int m = blockIdx.y;
  if (m >= d) return;
  const bool flip = U[(int64_t)m * I] < 0.0;
  if (!flip) return;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // V first (reads U[0, m] above before any thread of another block may have flipped it: row 0 goes last)
  if (idx < J) V[(int64_t)m * J + idx] = -V[(int64_t)m * J + idx];
  if (idx >= 1 && idx < I) U[(int64_t)m * I + idx] = -U[(int64_t)m * I + idx];
}
__global__ void sign_fix_row0_kernel(int64_t I, int d, double* __restrict__ U) {
  const int m = threadIdx.x;
  if (m < d && U[(int64_t)m * I] < 0.0) U[(int64_t)m * I] = -U[(int64_t)m * I];
}

void sim_data_impl(const scgbm_simdata_args* a, scgbm_simdata_out* out) {
  SCGBM_REQUIRE(a && out && out->Y, "NULL args");
  SCGBM_REQUIRE(a->model >= SCGBM_SIM_DATA && a->model <= SCGBM_SIM_NULL_NB, "unknown model");
  const int64_t I = a->I, J = a->J, Jt = a->J_total > 0 ? a->J_total : a->J, j0 = a->j_offset;
  SCGBM_REQUIRE(I >= 1 && J >= 1 && j0 >= 0 && j0 + J <= Jt, "sim: bad shape");
  SCGBM_REQUIRE(a->ldY >= I, "sim: ldY < I");
  const bool bilinear = a->model == SCGBM_SIM_DATA;
  const int d = bilinear ? a->d : 0;
  if (bilinear) SCGBM_REQUIRE(d >= 1 && d <= 96 && d <= std::min(I, Jt), "sim: d must be in 1..min(I, J, 96)");
  const double K = a->K > 0.0 ? a->K : 2.0;
  double theta = a->theta;
  if (a->model == SCGBM_SIM_NULL) theta = 0.0;
  if (a->model == SCGBM_SIM_NULL_NB && !(theta > 0.0)) theta = 1.0;       // R default theta = 1
  Ctx ctx(a->device, nullptr);
  SmallWs ws;
  DevBuf<double> U(std::max<int64_t>(I * d, 1)), V(std::max<int64_t>(Jt * d, 1)), tmp(std::max<int64_t>(std::max(I, Jt) * d, 1));
  DevBuf<double> Dd(std::max(d, 1)), al(I), be(Jt);
  DevBuf<int> flag(1);
  std::vector<double> Dh(std::max(d, 1), 0.0);
  auto grid = [](int64_t n) { return (unsigned)ceil_div(n, 256); };
  if (bilinear) {
    // U <- rustiefel(I, d), V <- rustiefel(J, d): Q factor of a Gaussian block (R/simulate.R:25-26)
    normal_fill_kernel<<<grid(I * d), 256, 0, ctx.stream>>>(I, d, I, fam_key(a->seed, FAM_U), 0, 0.0, 1.0, U.get());
    normal_fill_kernel<<<grid(Jt * d), 256, 0, ctx.stream>>>(Jt, d, Jt, fam_key(a->seed, FAM_V), 0, 0.0, 1.0, V.get());
    SCGBM_LAUNCH_CHECK();
    flag.zero(ctx.stream);
    sym_orth(ctx, ws, I, d, U.get(), I, nullptr, I, 0, tmp.get(), false, flag.get());
    sym_orth(ctx, ws, Jt, d, V.get(), Jt, nullptr, Jt, 0, tmp.get(), false, flag.get());
    int hf = 0;
    SCGBM_CUDA(cudaMemcpyAsync(&hf, flag.get(), sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
    ctx.sync();
    if (hf) throw Error(SCGBM_ERR_NUMERIC, "sim: Gaussian block numerically rank deficient");
    dim3 g(grid(std::max(I, Jt)), (unsigned)d);
    sign_fix_kernel<<<g, 256, 0, ctx.stream>>>(I, Jt, d, U.get(), V.get());
    sign_fix_row0_kernel<<<1, 256, 0, ctx.stream>>>(I, d, U.get());
    SCGBM_LAUNCH_CHECK();
    const double s = std::sqrt((double)I) + std::sqrt((double)Jt);
    for (int m = 0; m < d; ++m) Dh[m] = d == 1 ? K * s : K * s + (s - K * s) * (double)m / (double)(d - 1);   // seq(K s, s, length.out = d)
    SCGBM_CUDA(cudaMemcpyAsync(Dd.get(), Dh.data(), d * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
    normal_fill_kernel<<<grid(I), 256, 0, ctx.stream>>>(I, 1, I, fam_key(a->seed, FAM_ALPHA), 0, a->alpha_mean, 1.0, al.get());
    normal_fill_kernel<<<grid(Jt), 256, 0, ctx.stream>>>(Jt, 1, Jt, fam_key(a->seed, FAM_BETA), 0, a->beta_mean, 1.0, be.get());
    SCGBM_LAUNCH_CHECK();
  } else {
    log_uniform_fill_kernel<<<grid(I), 256, 0, ctx.stream>>>(I, fam_key(a->seed, FAM_MU), 1.0, 10.0, al.get());
    SCGBM_LAUNCH_CHECK();
    be.zero(ctx.stream);
  }
  // fp32 staging for the count kernel: Ud = U diag(D), this call's rows of V
  DevBuf<float> Ud(std::max<int64_t>(I * d, 1)), Vf(std::max<int64_t>(J * d, 1)), alf(I), bef(J);
  if (d > 0) {
    f64_to_f32_scaled<<<grid(I * d), 256, 0, ctx.stream>>>(I, d, U.get(), I, Dd.get(), Ud.get(), I);
    f64_to_f32_scaled<<<grid(J * d), 256, 0, ctx.stream>>>(J, d, V.get() + j0, Jt, nullptr, Vf.get(), J);
  }
  f64_to_f32_scaled<<<grid(I), 256, 0, ctx.stream>>>(I, 1, al.get(), I, nullptr, alf.get(), I);
  f64_to_f32_scaled<<<grid(J), 256, 0, ctx.stream>>>(J, 1, be.get() + j0, J, nullptr, bef.get(), J);
  SCGBM_LAUNCH_CHECK();
  const int es = a->out_dtype == SCGBM_U8 ? 1 : (a->out_dtype == SCGBM_U16 ? 2 : 4);
  DevBuf<uint8_t> Ytmp;
  void* Yd = out->Y;
  if (!a->y_on_device) { Ytmp.alloc(a->ldY * J * es); Yd = Ytmp.get(); }
  out->n_saturated = sim_counts_dev(ctx, I, J, a->ldY, j0, d, Ud.get(), Vf.get(), J, alf.get(), bef.get(), theta,
                                    a->seed, a->out_dtype, Yd);
  if (!a->y_on_device)
    SCGBM_CUDA(cudaMemcpyAsync(out->Y, Yd, (size_t)a->ldY * J * es, cudaMemcpyDeviceToHost, ctx.stream));
  if (out->U && d > 0) SCGBM_CUDA(cudaMemcpyAsync(out->U, U.get(), I * d * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  if (out->D && d > 0) memcpy(out->D, Dh.data(), d * sizeof(double));
  if (out->alpha) SCGBM_CUDA(cudaMemcpyAsync(out->alpha, al.get(), I * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  if (out->beta) SCGBM_CUDA(cudaMemcpyAsync(out->beta, be.get() + j0, J * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  std::vector<double> Vh;
  if (out->V && d > 0) {
    Vh.resize((size_t)J * d);
    SCGBM_CUDA(cudaMemcpy2DAsync(Vh.data(), J * sizeof(double), V.get() + j0, Jt * sizeof(double), J * sizeof(double), d,
                                 cudaMemcpyDeviceToHost, ctx.stream));
  }
  ctx.sync();
  if (out->V && d > 0)                                       // val$V <- t(diag(D) %*% t(V))   (R/simulate.R:42)
    for (int m = 0; m < d; ++m)
      for (int64_t j = 0; j < J; ++j) out->V[(size_t)m * J + j] = Vh[(size_t)m * J + j] * Dh[m];
}

// ---------------------------------------------------------------------------------
// CCI perturbation sampling   (R/cci.R:69-71)
// ---------------------------------------------------------------------------------
__global__ void cci_perturb_kernel(int64_t n /* J*M */, int reps, int rep_offset, uint64_t key,
                                   const double* __restrict__ scores, const double* __restrict__ se,
                                   double* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * reps) return;
  const int64_t e = idx % n;
  const int r = (int)(idx / n);
  Philox rng(key, (uint64_t)(rep_offset + r) * (uint64_t)n + (uint64_t)e);
  out[idx] = (scores ? scores[e] : 0.0) + se[e] * rng.normal();    // V + rnorm(J*M, sd = se_V)
}

void cci_perturb_impl(const scgbm_cci_args* a, double* out) {
  SCGBM_REQUIRE(a && out && a->se_scores, "NULL args");
  SCGBM_REQUIRE(a->J >= 1 && a->M >= 1 && a->reps >= 1 && a->rep_offset >= 0, "cci: bad shape");
  Ctx ctx(a->device, nullptr);
  const int64_t n = a->J * a->M;
  DevBuf<double> sc, se(n), tmp;
  SCGBM_CUDA(cudaMemcpyAsync(se.get(), a->se_scores, n * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  if (a->scores) {
    sc.alloc(n);
    SCGBM_CUDA(cudaMemcpyAsync(sc.get(), a->scores, n * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  }
  double* od = out;
  if (!a->out_on_device) { tmp.alloc(n * a->reps); od = tmp.get(); }
  cci_perturb_kernel<<<(unsigned)ceil_div(n * a->reps, 256), 256, 0, ctx.stream>>>(
      n, a->reps, a->rep_offset, fam_key(a->seed, FAM_CCI), sc.get(), se.get(), od);
  SCGBM_LAUNCH_CHECK();
  if (!a->out_on_device)
    SCGBM_CUDA(cudaMemcpyAsync(out, od, (size_t)n * a->reps * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  ctx.sync();
}

}  // namespace scgbm
