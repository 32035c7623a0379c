// products.cuh -- K5: tall-skinny products with the materialised fp32 residual R
// (ldI x J, column-major) inside the truncated SVD that replaces irlba (R/fit.R:323):
//     prod_n : out[ldI x b] += c * R   P      (reduction over cells; partial per rank)
//     prod_t : out[J   x b] += c * R^T Q      (reduction over genes; local)
// Operands P/Q arrive as fp64 column-major blocks and are converted to the layout
// the kernel wants; accumulation is fp32 inside short runs, fp64 across them.
#pragma once
#include "common.cuh"

namespace scgbm {

struct ProdWs {
  DevBuf<float> op32;      // converted operand: rows x bp, row-major
  DevBuf<double> part;     // prod_n split-K partials
};

enum { PROD_AUTO = 0, PROD_FFMA = 1, PROD_TCGEN05 = 2 };

void prod_n(Ctx& ctx, ProdWs& ws, const float* R, int64_t I, int64_t ldI, int64_t J, const double* P,
            int64_t ldP, int b, double c, double* out, int64_t ldo, int variant);

void prod_t(Ctx& ctx, ProdWs& ws, const float* R, int64_t I, int64_t ldI, int64_t J, const double* Q,
            int64_t ldQ, int b, double c, double* out, int64_t ldo, int variant);

}  // namespace scgbm
