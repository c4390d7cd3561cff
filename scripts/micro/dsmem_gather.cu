// Microbenchmark: random 4-byte gathers from distributed shared memory (cluster) vs local smem.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_gather dsmem_gather.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

constexpr int WORDS = 32768;  // 128 KB per CTA

template <int CS>
__global__ void __launch_bounds__(1024, 1) k_gather(uint32_t *out, int iters, int remote_mode) {
  extern __shared__ uint32_t tab[];
  cg::cluster_group cl = cg::this_cluster();
  for (int i = threadIdx.x; i < WORDS; i += blockDim.x) tab[i] = i * 2654435761u;
  cl.sync();
  const uint32_t rank = cl.block_rank();
  uint32_t base[CS];
#pragma unroll
  for (int r = 0; r < CS; r++) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(tab), m;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(m) : "r"(a), "r"(r));
    base[r] = m;
  }
  uint32_t x = threadIdx.x * 747796405u + blockIdx.x * 2891336453u + 1;
  uint32_t acc = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      x = x * 1664525u + 1013904223u;
      const uint32_t idx = __umulhi(x, (uint32_t)WORDS);
      uint32_t r;
      if (remote_mode == 0) r = rank;                      // all local
      else if (remote_mode == 1) r = (x >> 3) % CS;        // uniform over the cluster
      else r = (rank + 1 + ((x >> 3) % (CS - 1 ? CS - 1 : 1))) % CS;  // always remote
      uint32_t b = base[0];
#pragma unroll
      for (int q = 1; q < CS; q++) if (r == q) b = base[q];
      uint32_t v;
      asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(b + idx * 4));
      acc += v;
    }
  }
  cl.sync();
  if (acc == 0x12345678) out[0] = acc;
}

template <int CS>
void run(int mode, int iters) {
  uint32_t *out;
  cudaMalloc(&out, 4);
  cudaFuncSetAttribute(k_gather<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, WORDS * 4);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148 / CS * CS);
  cfg.blockDim = dim3(1024);
  cfg.dynamicSmemBytes = WORDS * 4;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; rep++) {
    cudaEventRecord(e0);
    cudaError_t err = cudaLaunchKernelEx(&cfg, k_gather<CS>, out, iters, mode);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (err != cudaSuccess || cudaGetLastError() != cudaSuccess) { printf("CS=%d mode=%d launch error %s\n", CS, mode, cudaGetErrorString(err)); return; }
    if (rep == 1) {
      double lookups = (double)cfg.gridDim.x * 1024 * iters * 16;
      printf("cluster=%d mode=%d (%s): %.3f ms, %.1f G lookups/s, %.2f lookups/clk/SM @1.9GHz\n", CS, mode,
             mode == 0 ? "local" : mode == 1 ? "uniform" : "remote", ms, lookups / ms / 1e6,
             lookups / (ms * 1e-3) / cfg.gridDim.x / 1.9e9);
    }
  }
  cudaFree(out);
}

int main() {
  const int iters = 2000;
  run<1>(0, iters);
  run<2>(0, iters); run<2>(1, iters); run<2>(2, iters);
  run<4>(0, iters); run<4>(1, iters); run<4>(2, iters);
  run<8>(0, iters); run<8>(1, iters); run<8>(2, iters);
  return 0;
}
