// Random 4-byte loads from an L2-resident table: loads per clock per SM against the number
// of independent loads each lane keeps in flight (K) and the table size.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_gather l2_gather.cu && ./l2_gather
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int K>
__global__ void __launch_bounds__(1024, 1) k(const uint32_t *tab, uint32_t mask, uint32_t *out, int iters, long long *cyc) {
  uint32_t h = (blockIdx.x * 1024 + threadIdx.x) * 2654435761u + 12345u, acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    uint32_t v[K];
#pragma unroll
    for (int i = 0; i < K; i++) {
      h = h * 1664525u + 1013904223u;
      v[i] = __ldg(tab + ((h >> 7) & mask));
    }
#pragma unroll
    for (int i = 0; i < K; i++) acc += v[i];
  }
  long long t1 = clock64();
  out[blockIdx.x * 1024 + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int K> void run(const uint32_t *tab, uint32_t words, uint32_t *out, long long *cyc) {
  const int iters = 2048 / K;
  k<K><<<148, 1024>>>(tab, words - 1, out, iters, cyc); cudaDeviceSynchronize();
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a); k<K><<<148, 1024>>>(tab, words - 1, out, iters, cyc); cudaEventRecord(b); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, a, b);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double loads = 1024.0 * iters * K;  // per SM
  printf("table %6.1f MB  K=%2d  %.3f loads/clk/SM  %.1f G loads/s chip  (%.3f ms)\n", words * 4 / 1048576.0, K, loads / h,
         loads * 148 / ms / 1e6, ms);
}
int main() {
  uint32_t *out; long long *cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  for (uint32_t mb : {1u, 16u, 64u, 256u}) {
    const uint32_t words = mb * 262144u;
    uint32_t *tab; cudaMalloc(&tab, (size_t)words * 4); cudaMemset(tab, 1, (size_t)words * 4);
    run<1>(tab, words, out, cyc); run<4>(tab, words, out, cyc); run<16>(tab, words, out, cyc);
    cudaFree(tab);
  }
  return 0;
}
