// proj_tc.cuh -- K9 on the tensor cores (proj_tc.cu): lock-step Newton / IRLS over all cells with the fused pass's
// product rider delivering gradient and Hessian (R/fit.R:248-264)
#pragma once
#include "common.cuh"
#include "counts.cuh"

namespace scgbm {

struct ProjTcStats {
  int sweeps = 0, blocks = 0;
  int64_t rider_launches = 0, stragglers = 0;
};

bool proj_tc_supported(Ctx& ctx, const Counts& Y, int M);

// Y: the kept genes' counts (Ik x J, u8 / u16); U_dev Ik x M (ld Ik), off_dev Ik -- device, fp64.
// coef / se: J x (M + 1) device (ld J), fail / iters: J.  Cells that converged are written; the others (no counts,
// failed solve, sweep budget) are returned as `list` (device, first `return value` entries) for the exact kernel.
int64_t project_tc(Ctx& ctx, const Counts& Y, int64_t Ik, int M, const double* U_dev, const double* off_dev, double tol,
                   int max_sweeps, double* coef, double* se, int8_t* fail, int32_t* iters, DevBuf<int64_t>& list,
                   ProjTcStats* stats);

}  // namespace scgbm
