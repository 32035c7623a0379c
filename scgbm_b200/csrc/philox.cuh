// philox.cuh -- counter-based RNG for the on-device generators (simData recipe, CCI perturbations).
#pragma once
#include "common.cuh"

namespace scgbm {

// ---------------------------------------------------------------------------------
// on-device generators: Y ~ Poisson / NB(exp(alpha_i + beta_j + (U D V^T)_ij)), Gaussian / uniform draws
// counter-based Philox4x32-10 keyed by (seed, element index) so shards are
// reproducible independently of the launch geometry.
// ---------------------------------------------------------------------------------
struct Philox {
  uint32_t c[4], k[2], out[4];
  int have;
  __device__ Philox(uint64_t seed, uint64_t idx) {
    c[0] = (uint32_t)idx; c[1] = (uint32_t)(idx >> 32); c[2] = 0; c[3] = 0x5C6B0000u;
    k[0] = (uint32_t)seed; k[1] = (uint32_t)(seed >> 32);
    have = 0;
  }
  __device__ void round_fn(uint32_t* ctr, const uint32_t* key) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr[0]), lo0 = 0xD2511F53u * ctr[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr[2]), lo1 = 0xCD9E8D57u * ctr[2];
    const uint32_t n0 = hi1 ^ ctr[1] ^ key[0], n1 = lo1, n2 = hi0 ^ ctr[3] ^ key[1], n3 = lo0;
    ctr[0] = n0; ctr[1] = n1; ctr[2] = n2; ctr[3] = n3;
  }
  __device__ void refill() {
    uint32_t ctr[4] = {c[0], c[1], c[2], c[3]};
    uint32_t key[2] = {k[0], k[1]};
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      round_fn(ctr, key);
      key[0] += 0x9E3779B9u; key[1] += 0xBB67AE85u;
    }
    out[0] = ctr[0]; out[1] = ctr[1]; out[2] = ctr[2]; out[3] = ctr[3];
    c[2] += 1;
    have = 4;
  }
  __device__ uint32_t next() {
    if (have == 0) refill();
    return out[--have];
  }
  __device__ double uniform() {  // (0,1), 53-bit
    const uint64_t a = next() >> 5, b = next() >> 6;
    return ((double)a * 67108864.0 + (double)b + 0.5) * (1.0 / 9007199254740992.0);
  }
  __device__ double normal() {
    const double u1 = uniform(), u2 = uniform();
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
  }
  __device__ double gamma(double shape) {  // Marsaglia-Tsang
    double boost = 1.0;
    if (shape < 1.0) { boost = pow(uniform(), 1.0 / shape); shape += 1.0; }
    const double d = shape - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * d);
    for (int it = 0; it < 1000; ++it) {
      double x = normal(), v = 1.0 + cc * x;
      if (v <= 0.0) continue;
      v = v * v * v;
      const double u = uniform();
      if (u < 1.0 - 0.0331 * x * x * x * x) return boost * d * v;
      if (log(u) < 0.5 * x * x + d * (1.0 - v + log(v))) return boost * d * v;
    }
    return boost * d;
  }
  __device__ double poisson(double lam) {
    if (!(lam > 0.0)) return 0.0;
    if (lam < 12.0) {   // inversion by sequential search
      const double u = uniform();
      double p = exp(-lam), q = p;
      int k = 0;
      while (u > q && k < 200) { ++k; p *= lam / k; q += p; }
      return (double)k;
    }
    // PTRS (Hoermann 1993)
    const double slam = sqrt(lam), loglam = log(lam);
    const double b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
    const double invalpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2.0);
    for (int it = 0; it < 1000; ++it) {
      const double U = uniform() - 0.5, V = uniform();
      const double us = 0.5 - fabs(U);
      const double k = floor((2.0 * a / us + b) * U + lam + 0.43);
      if (us >= 0.07 && V <= vr) return k;
      if (k < 0.0 || (us < 0.013 && V > us)) continue;
      if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -lam + k * loglam - lgamma(k + 1.0)) return k;
    }
    return floor(lam);
  }
};

template <typename Tout> struct SimLimit;
template <> struct SimLimit<uint8_t> { static constexpr double max = 255.0; };
template <> struct SimLimit<uint16_t> { static constexpr double max = 65535.0; };
template <> struct SimLimit<float> { static constexpr double max = 16777216.0; };
template <> struct SimLimit<int32_t> { static constexpr double max = 2147483647.0; };


}  // namespace scgbm
