// products.cuh -- K5: tall-skinny products with the materialised residual R = Y - W
// (ldI x J, column-major) inside the truncated SVD that replaces irlba (R/fit.R:323):
//     prod_n : out[ldI x b] += c * R   P      (reduction over cells; partial per rank)
//     prod_t : out[J   x b] += c * R^T Q      (reduction over genes; local)
//
// R is stored as TWO fp16 planes of the scaled residual, scale * R = hi + lo with
// hi = fp16(scale R), lo = fp16(scale R - hi): the same 4 bytes per entry as fp32, 22
// significant bits (two bf16 planes would keep 16 and cost a factor 10 in parity with the
// reference), round-to-nearest, and each plane is directly a tcgen05 kind::f16 operand, so
// TMA feeds the tensor cores without any in-kernel conversion pass.  `scale` is a power
// of two chosen by the writer so that |scale R| <= 32768 (fp16 overflows at 65504); the
// products fold 1/scale into their constant.  The thin operand is scaled by its own power
// of two and split into two fp16 planes (22 bits; fp16 x fp16 products are exact in fp32),
// concatenated along N:
//     D[:, 0:2bp]  = hi . [P1|P2]        D[:, 0:bp] += lo . P1
//     result[:, k] = D[:, k] + D[:, bp+k]
// (the dropped lo.P2 term is <= 2^-22 relative, ~2^-24 rms).  Accumulators live in TMEM
// (fp32) for 128 reduction steps, then drain into fp64 registers of the epilogue warps, so
// long reductions (20k genes / 1M cells) keep fp64 accuracy.
#pragma once
#include <cuda_fp16.h>
#include <cmath>
#include "common.cuh"

namespace scgbm {

struct ResidBuf {
  uint16_t* hi = nullptr;   // fp16 bit patterns, ldI x J
  uint16_t* lo = nullptr;
  float scale = 1.0f;       // stored value = scale * R (power of two)
};

// largest power of two s with s * bound <= 32768 (clamped to 2^-60 .. 2^14)
inline float resid_scale_for(double bound) {
  if (!(bound > 0.0) || !std::isfinite(bound)) return 1.0f;
  int k = (int)std::floor(std::log2(32768.0 / bound));
  k = std::max(-60, std::min(14, k));
  return (float)std::ldexp(1.0, k);
}

struct ProdWs {
  DevBuf<float> op32;       // FFMA path: converted operand, rows x bp row-major
  DevBuf<uint16_t> op16;    // tcgen05 path: 2 fp16 planes of the scaled operand, [2*bp][ldk]
  DevBuf<float> opmeta;     // [0] max |operand| (float bits), [1] 1 / operand scale
  DevBuf<double> part;      // prod_n split-K partials  [nchunk][bp][ldI]
};

// tcgen05 path: ws.op16 <- the two scaled fp16 planes [2*bp][ldk] of the fp64 operand src (rows x b, ld),
// ws.opmeta[1] <- 1 / scale  (device side; no host round trip)
void split_operand(Ctx& ctx, ProdWs& ws, int64_t rows, int b, int bp, int64_t ldk, const double* src, int64_t ld);
bool prod_uses_tc(int b);     // the dispatch prod_t / prod_n take for PROD_AUTO

enum { PROD_AUTO = 0, PROD_FFMA = 1, PROD_TCGEN05 = 2 };

// use_lo = false: the tcgen05 kernel reads only the hi plane (11-bit residual, half the bytes) --
// good enough for passes that merely build a Krylov basis; every pass whose output enters the
// Rayleigh-Ritz projection (prod_t) keeps both planes.
void prod_n(Ctx& ctx, ProdWs& ws, ResidBuf R, int64_t I, int64_t ldI, int64_t J, const double* P,
            int64_t ldP, int b, double c, double* out, int64_t ldo, int variant, bool use_lo = true);

void prod_t(Ctx& ctx, ProdWs& ws, ResidBuf R, int64_t I, int64_t ldI, int64_t J, const double* Q,
            int64_t ldQ, int b, double c, double* out, int64_t ldo, int variant, bool use_lo = true);

// tcgen05 implementations (products_tc.cu); b <= 64
bool prod_tc_supported(int b);
void prod_n_tc(Ctx& ctx, ProdWs& ws, ResidBuf R, int64_t ldI, int64_t J, const double* P, int64_t ldP,
               int b, double c, double* out, int64_t ldo, bool use_lo);
void prod_t_tc(Ctx& ctx, ProdWs& ws, ResidBuf R, int64_t ldI, int64_t J, const double* Q, int64_t ldQ,
               int b, double c, double* out, int64_t ldo, bool use_lo);
void prod_n_reduce_launch(Ctx& ctx, int nchunk, int64_t ldI, int b, int bp, const double* part, double c,
                          const float* c_dev /* optional extra device factor */, double* out, int64_t ldo);

// fp32 (ld x ncol, column-major) -> fp16 hi/lo planes of dst.scale * src (diagnostics, tests)
void split_planes(Ctx& ctx, const float* src, int64_t n, ResidBuf dst);

#ifdef __CUDACC__
__device__ __forceinline__ float bf16_bits_to_float(uint16_t b) { return __uint_as_float(((uint32_t)b) << 16); }
// round-to-nearest-even bf16 of a float (NaN stays NaN, Inf stays Inf)
__device__ __forceinline__ uint16_t float_to_bf16_bits(float f) {
  uint32_t u = __float_as_uint(f);
  if ((u & 0x7f800000u) == 0x7f800000u) return (uint16_t)((u >> 16) | ((u & 0xffffu) ? 0x40u : 0u));
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
// residual planes: r (already scaled) = hi + lo, both fp16
__device__ __forceinline__ void split_f16x2(float r, uint16_t& hi, uint16_t& lo) {
  const __half h = __float2half_rn(r);
  hi = __half_as_ushort(h);
  lo = __half_as_ushort(__float2half_rn(r - __half2float(h)));
}
__device__ __forceinline__ float join_f16x2(uint16_t hi, uint16_t lo) {
  return __half2float(__ushort_as_half(hi)) + __half2float(__ushort_as_half(lo));
}
#endif

}  // namespace scgbm
