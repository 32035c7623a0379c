// HBM rate for a given read:write mix -- what can the fused pass's 1:2 read:write traffic reach at all?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/micro/membench scripts/micro/membench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); exit(1); } } while (0)

// each thread: NR 16-byte loads, NW 16-byte stores per step
template <int NR, int NW>
__global__ void __launch_bounds__(256) mix(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n_steps) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_steps; s += stride) {
    uint4 v[NR > 0 ? NR : 1];
#pragma unroll
    for (int r = 0; r < NR; ++r) v[r] = __ldcs(in + s + (size_t)r * n_steps);
#pragma unroll
    for (int r = 0; r < NR; ++r) { acc.x ^= v[r].x; acc.y += v[r].y; acc.z ^= v[r].z; acc.w += v[r].w; }
#pragma unroll
    for (int w = 0; w < NW; ++w) __stcs(out + s + (size_t)w * n_steps, make_uint4(acc.x + w, acc.y, acc.z, acc.w));
  }
  if (NW == 0 && acc.x == 0x12345678u) out[0] = acc;
}

template <int NR, int NW>
void run(const char* name, uint4* in, uint4* out, size_t n_steps) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  const int grid = 148 * 8;
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaEventRecord(a));
    for (int k = 0; k < 3; ++k) mix<NR, NW><<<grid, 256>>>(in, out, n_steps);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
  }
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  const double bytes = 3.0 * (double)n_steps * 16.0 * (NR + NW);
  printf("%-22s read:write %d:%d  %.1f GB in %.2f ms -> %.0f GB/s\n", name, NR, NW, bytes / 1e9, ms, bytes / ms / 1e6);
}

int main() {
  const size_t n_steps = (size_t)1 << 28;     // 4 GiB per stream
  uint4 *in, *out;
  CK(cudaMalloc(&in, n_steps * 16 * 2));
  CK(cudaMalloc(&out, n_steps * 16 * 4));
  CK(cudaMemset(in, 1, n_steps * 16 * 2));
  run<1, 0>("pure read", in, out, n_steps);
  run<0, 1>("pure write", in, out, n_steps);
  run<1, 1>("copy", in, out, n_steps);
  run<1, 2>("fused-pass mix", in, out, n_steps);
  run<2, 1>("2 reads 1 write", in, out, n_steps);
  run<1, 4>("1 read 4 writes", in, out, n_steps);
  return 0;
}
