// tsvd.cuh -- warm-started, restarted block-Krylov truncated SVD of the operator
//     G = sum_t coef_t * A_t diag(d_t) B_t^T  +  c * R
// (low-rank terms exact in fp64 through their factors, R the materialised residual
// in two bf16 planes).  Replaces irlba::irlba(V + (lr/w.max)(Y - W), nv = M) at R/fit.R:323 and
// irlba::irlba(Z, nv = M) at R/fit.R:115.
#pragma once
#include "common.cuh"
#include "dense_small.cuh"
#include "products.cuh"

namespace scgbm {

struct LowRankTerm {
  const double* A = nullptr; int64_t lda = 0;   // ldI x M
  const double* d = nullptr;                    // M (device)
  const double* B = nullptr; int64_t ldb = 0;   // ldJ x M
  int M = 0;
  double coef = 0.0;
};

struct GOperator {
  int64_t I = 0, ldI = 0, J = 0, ldJ = 0;
  int64_t J_total = 0;     // over all ranks
  ResidBuf R;              // bf16 hi/lo planes of the residual (R.hi == nullptr: no dense term)
  double c = 0.0;
  int nterms = 0;
  LowRankTerm t[2];
  // R^T Q0 (ldJ x b) for Q0 = the start block tsvd_left_start() announced, computed by the fused
  // likelihood pass that produced R (eta_pass.cuh: FusedProd); nullptr: tsvd reads the planes itself
  const double* RtQ0 = nullptr;
};

struct TsvdOpts {
  int M = 0;
  int extra = 12;       // block = M + extra
  int depth = 3;        // Krylov extensions per cycle (warm start)
  int cold_depth = 7;   // ... on a cold start (fewer, deeper cycles; the p = 8b Ritz problem runs on the cooperative Jacobi)
  double tol = 1e-4;    // Ritz residual estimate (cold start)
  double tol_warm = 0;  // ... of the warm calls inside the loop (0: = tol)
  int max_cycles = 40;  // restarts (cold start on gap-less spectra)
  int max_warm_cycles = 8;  // ... of a warm call that is still above the tolerance after its first cycle
  uint64_t seed = 0;
  int variant = PROD_AUTO;
  int fast = 0;         // allow 2-pass calls (no Krylov extension) while the previous estimate is < tol^2
};

struct TsvdWs {
  SmallWs small;
  ProdWs prod;
  DevBuf<double> Qb, Zb, tmpI, T, Tw, S, theta, C, est, Ub, Vb;
  DevBuf<int> flag;       // set by the Cholesky orthonormalisation when a pivot vanishes
  bool robust = false;    // eigen-based orthonormalisation (rank-deficient operators)
  int pmax = 0;
  DevBuf<double> Vwarm;   // ldJ x b warm start for the next call
  bool warm = false;
  int b = 0;
  int64_t passes_total = 0, calls = 0;
  double last_est = 1e300;  // residual estimate of the last call that ran a Krylov extension
  int fast_streak = 0;      // consecutive calls that skipped the extension
  int warm_calls = 0;       // consecutive warm calls so far (Ub then belongs to the same operator family)
  bool lowprec = false;     // cold start, early cycles: products read only the residual's hi plane
  bool start_hi = false;    // warm call: the start block Q0 = orth(G V0) is built from the hi plane only
  int unconverged = 0;      // calls that returned with a Ritz residual estimate >= tol (cycle budget exhausted)
  double worst_est = 0.0;   // the largest such estimate
};

// outU: ldI x M, outD: M, outV: ldJ x M (device, column-major)
// upload a J x b start block (host, column-major) as the warm start of the next call
void tsvd_set_warm_start(Ctx& ctx, TsvdWs& ws, const GOperator& op, const TsvdOpts& o, const double* V0_host);

// The next warm call starts its Krylov space from a left block that is already known (the previous
// call's left Ritz block): returns it (ldI x *b) or nullptr when the next call starts differently.
const double* tsvd_left_start(const TsvdWs& ws, int64_t I, int64_t J_total, const TsvdOpts& o, int* b);

void tsvd(Ctx& ctx, TsvdWs& ws, const GOperator& op, const TsvdOpts& o, double* outU, double* outD,
          double* outV, int* passes_out, double* est_out);

}  // namespace scgbm
