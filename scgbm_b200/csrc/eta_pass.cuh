// eta_pass.cuh -- K3/K4: the dense "evaluate eta" passes of the loop body
// (R/fit.R:127-151, :321).  X is never materialised: it is rebuilt per tile from
// shared-memory-staged factor tiles.
#pragma once
#include "common.cuh"
#include "counts.cuh"
#include "products.cuh"

namespace scgbm {

// tensor-core form of one side of X (eta_tc.cu): fp16 rows [rows_pad][Kp] holding the three
// K sections of the scaled, two-piece factor; meta[1] = 1 / scale (device)
struct Factor16 {
  DevBuf<uint16_t> data;
  DevBuf<float> meta;
  int M = 0, Mp = 0, Kp = 0;
  int64_t rows_pad = 0;
};

// X_ij = clamp?( rowscale?[i, b(j)] * sum_m A[i,m] B[j,m] )
struct XDesc {
  const float* A = nullptr;   // ldI x M (col-major), d (and single-batch row scale) folded in
  const float* B = nullptr;   // ldJ x M (col-major), column scale folded in
  int M = 0;                  // 0 -> X == 0
  int clamp = 0;              // clamp to [-8, 8]   (R/fit.R:119-120)
  const float* rowscale = nullptr;  // ldI x nbatch or null (multi-batch X0 only)
  const Factor16* A16 = nullptr;    // genes (side 0) and cells (side 1) for the tcgen05 passes, or null
  const Factor16* B16 = nullptr;
  // fp64 factors of an unclamped low-rank X = U diag(d) V^T (null otherwise): rows whose |x| grows beyond what the
  // 22-bit tensor-core eta resolves to 1e-5 are re-summed exactly from these (pass_rowsum_tc)
  const double* U64 = nullptr; const double* d64 = nullptr; const double* V64 = nullptr;
};

struct PassGeom {
  int64_t I, J, ldI, ldJ;
  int nbatch;
  const int32_t* batch;  // J zero-based or null
};

struct EtaWs {
  DevBuf<double> row_part;   // [chunks][nbatch][ldI]
  DevBuf<double> scal_part;  // [ctas][4]
  DevBuf<double> scal;       // [4]: sum_w, sum_ylogw, max_w, spare
  DevBuf<float> gmask;       // exp(beta) masked to one batch (multi-batch row sums on tcgen05)
  DevBuf<int> row_flag;      // ldI: genes whose row sum is redone in fp64 (max |x| over the row >= kExactRowX)
  DevBuf<unsigned long long> exact_rows;   // [1] running count of such rows (diagnostic)
  DevBuf<int> exact_list;    // [1 + cap]: count, then the flagged genes in ascending order
  DevBuf<double> exact_part; // [nbatch][cap][cell chunks] fixed-order partial sums
};

// |x| beyond which a gene's row sum S_i = sum_j exp(beta_j + x_ij) is recomputed in fp64: the tensor-core eta carries
// 22 bits, i.e. an absolute error of |x| 2^-22 (x some cancellation), and alpha_i = log rs_i - log S_i inherits it
// undiluted when a few large entries dominate the row (sparse genes the fit drives to |x| ~ 50): measured at
// 20 000 genes, alpha error 4.6e-6 for rows with max |x| < 16, up to 3.4e-5 beyond (tolerance 1e-5).
constexpr float kExactRowX = 12.0f;

// S[i,k] = sum_{j in batch k} g_j * exp(X_ij)      (R/fit.R:129)   -> S (ldI x nbatch, fp64)
void pass_rowsum(Ctx& ctx, EtaWs& ws, const PassGeom& g, const XDesc& x, const float* g_col, double* S);

// C[j] = sum_i f_{i,b(j)} * exp(X_ij)               (R/fit.R:134)   -> C (J, fp64)
void pass_colsum(Ctx& ctx, EtaWs& ws, const PassGeom& g, const XDesc& x, const float* f_row, double* C);

// W = min(f_i g_j exp(X_ij), max.Y); resid = Y - W (fp16 hi/lo planes of resid.scale * (Y - W), ldI x J); scalars:
// out[0] = sum W, out[1] = sum_{Y != 0} Y log W, out[2] = max W    (R/fit.R:136-151,321)
void pass_fused(Ctx& ctx, EtaWs& ws, const PassGeom& g, const XDesc& x, const Counts& y,
                const float* f_row, const float* g_col, const float* la_row, const float* lb_col,
                float maxY, ResidBuf resid, double* out3);

// resid += coef * X_ij   (iterations 1-2: the clamped, non-low-rank X0 terms of G)
void pass_addx(Ctx& ctx, const PassGeom& g, const XDesc& x, float coef, ResidBuf resid);

// ---- tcgen05 versions (eta_tc.cu): u8/u16 counts, M <= 64 (column sums: single batch) ------
bool eta_tc_supported(const PassGeom& g, const XDesc& x, int y_storage /* -1: pass without Y */);
// side 0 = genes [hi|hi|lo], side 1 = cells [hi|lo|hi]; value = U[r,k] * d[k] * rs[r]
void stage_factor16(Ctx& ctx, Factor16& f, int64_t n_real, int64_t ld, int M, int side, const double* U,
                    const double* d, const float* rs);
// beta64 (J, fp64) together with x.U64/d64/V64 enables the exact re-summation of extreme rows (kExactRowX)
void pass_rowsum_tc(Ctx& ctx, EtaWs& ws, const PassGeom& g, const XDesc& x, const float* g_col, double* S,
                    const double* beta64 = nullptr);
// f_row / g_col must be zero padded by 128 entries beyond ldI / ldJ (whole 128-wide tiles are read)
void pass_colsum_tc(Ctx& ctx, EtaWs& ws, const PassGeom& g, const XDesc& x, const float* f_row, double* C);
// alpha/beta (fp64) and this rank's row sums of Y close the objective:
//   sum_nz Y log W = sum Y x + sum_i alpha_i rs_i + sum_j beta_j cs_j   (+ clip correction)
// out3[3] = 1 flags an iteration in which exp(eta) may underflow in the reference's doubles (log(0) = -Inf for a
// non-zero count, R/fit.R:152-157): the caller must then re-run it with pass_fused, which restates that.
// Optional rider of the fused pass: out (J x b, column-major) = R^T Q0 for the residual this pass produces,
// Q0 given as the two scaled fp16 planes prod_t would build ([hi 32 | lo 32][ldk]; products.cuh).
struct FusedProd {
  const uint16_t* op16; int64_t ldk;
  const float* op_inv;        // device: 1 / operand scale
  int b;                      // block columns, <= 32
  double* out; int64_t ldo;
  // product only (proj_tc.cu): no residual planes are written -- the launch reads Y, rebuilds W and delivers
  // (Y - W)^T Q0; `resid` then only carries the scale.  Implies any shard size (no "worth it" threshold).
  bool product_only = false;
};
bool fused_prod_supported(Ctx& ctx, const PassGeom& g, const XDesc& x, bool any_size = false);
void pass_fused_tc(Ctx& ctx, EtaWs& ws, const PassGeom& g, const XDesc& x, const Counts& y, const float* f_row,
                   const float* g_col, const float* la_row, const float* lb_col, const double* alpha,
                   const double* beta, const double* rowsum_local, float maxY, ResidBuf resid, double* out3,
                   const FusedProd* prod = nullptr);
// fixed-order second stages shared by both implementations
void scal_reduce_launch(Ctx& ctx, int64_t ncta, const double* part, double* out3);
void rowpart_reduce_launch(Ctx& ctx, int nchunk, int64_t n, const double* part, double* S);

// algorithmic bytes of one fused-pass launch (SURVEY 8d / BASELINE.md section 4)
double fused_pass_bytes(const PassGeom& g, const XDesc& x, int y_elem_bytes);

}  // namespace scgbm
