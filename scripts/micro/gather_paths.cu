// Random 4/16-byte gathers from an L2-resident table through the three paths an SM has:
//   ldg   ld.global.nc per lane (L1tex: one wavefront per distinct 128-byte line)
//   tex   tex1Dfetch per lane (texture path of the same L1tex unit)
//   tma   cp.async.bulk of 16 bytes per lane into shared memory, mbarrier completion
//   mix   every warp issues both kinds: does TMA add to the LDG rate or share its limit?
// Prints gathers per clock per SM.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_paths gather_paths.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t rnd(uint32_t &h) { h = h * 1664525u + 1013904223u; return h >> 7; }

__global__ void __launch_bounds__(1024, 1) k_ldg(const uint32_t *tab, uint32_t mask, uint32_t *out, int iters, long long *cyc) {
  uint32_t h = (blockIdx.x * 1024 + threadIdx.x) * 2654435761u + 12345u, acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    uint32_t v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __ldg(tab + (rnd(h) & mask));
#pragma unroll
    for (int i = 0; i < 8; i++) acc += v[i];
  }
  long long t1 = clock64();
  out[blockIdx.x * 1024 + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__global__ void __launch_bounds__(1024, 1) k_tex(cudaTextureObject_t tex, uint32_t mask, uint32_t *out, int iters, long long *cyc) {
  uint32_t h = (blockIdx.x * 1024 + threadIdx.x) * 2654435761u + 12345u, acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    uint32_t v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = tex1Dfetch<uint32_t>(tex, (int)(rnd(h) & mask));
#pragma unroll
    for (int i = 0; i < 8; i++) acc += v[i];
  }
  long long t1 = clock64();
  out[blockIdx.x * 1024 + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// every lane issues NB bulk copies of 16 bytes per iteration into its own shared-memory slots;
// MIXL extra LDG gathers per iteration alongside
template <int NB, int MIXL>
__global__ void __launch_bounds__(1024, 1) k_tma(const uint32_t *tab, uint32_t mask, uint32_t *out, int iters, long long *cyc) {
  extern __shared__ __align__(16) uint8_t sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t *slots = sm + (size_t)warp * 32 * NB * 16;                 // per warp: 32 lanes x NB x 16 B
  uint64_t *bar = reinterpret_cast<uint64_t *>(sm + 32 * 32 * NB * 16) + warp;
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(bar);
  const uint32_t slot_a = (uint32_t)__cvta_generic_to_shared(slots) + lane * NB * 16;
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_a), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t h = (blockIdx.x * 1024 + threadIdx.x) * 2654435761u + 12345u, acc = 0, phase = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(32 * NB * 16) : "memory");
    __syncwarp();
#pragma unroll
    for (int i = 0; i < NB; i++) {
      const uint32_t *src = tab + ((rnd(h) & mask) & ~3u);  // 16-byte aligned
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];"
                   ::"r"(slot_a + i * 16), "l"(src), "r"(bar_a) : "memory");
    }
    uint32_t v[MIXL > 0 ? MIXL : 1];
#pragma unroll
    for (int i = 0; i < MIXL; i++) v[i] = __ldg(tab + (rnd(h) & mask));
    asm volatile(
        "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}" ::"r"(bar_a), "r"(phase) : "memory");
    phase ^= 1;
#pragma unroll
    for (int i = 0; i < NB; i++) acc += *reinterpret_cast<volatile uint32_t *>(slots + lane * NB * 16 + i * 16);
#pragma unroll
    for (int i = 0; i < MIXL; i++) acc += v[i];
    __syncwarp();
  }
  long long t1 = clock64();
  out[blockIdx.x * 1024 + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  uint32_t *out; long long *cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const uint32_t words = 16u * 262144u;  // 16 MB
  uint32_t *tab; cudaMalloc(&tab, (size_t)words * 4); cudaMemset(tab, 1, (size_t)words * 4);
  cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = tab;
  rd.res.linear.desc = cudaCreateChannelDesc<uint32_t>(); rd.res.linear.sizeInBytes = (size_t)words * 4;
  cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
  cudaTextureObject_t tex; cudaCreateTextureObject(&tex, &rd, &td, nullptr);
  auto report = [&](const char *name, double per_iter, int iters) {
    cudaDeviceSynchronize();
    cudaError_t e = cudaGetLastError();
    long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s %s  %.3f gathers/clk/SM\n", name, e == cudaSuccess ? "ok" : cudaGetErrorString(e), 1024.0 * iters * per_iter / h);
  };
  const int iters = 256;
  for (int rep = 0; rep < 2; rep++) {
    k_ldg<<<148, 1024>>>(tab, words - 1, out, iters, cyc); report("ldg x8", 8, iters);
    k_tex<<<148, 1024>>>(tex, words - 1, out, iters, cyc); report("tex x8", 8, iters);
    {
      auto f = k_tma<4, 0>; size_t smem = 32 * 32 * 4 * 16 + 32 * 8;
      cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      f<<<148, 1024, smem>>>(tab, words - 1, out, iters, cyc); report("tma 16B x4", 4, iters);
    }
    {
      auto f = k_tma<1, 0>; size_t smem = 32 * 32 * 1 * 16 + 32 * 8;
      f<<<148, 1024, smem>>>(tab, words - 1, out, iters, cyc); report("tma 16B x1", 1, iters);
    }
    {
      auto f = k_tma<2, 6>; size_t smem = 32 * 32 * 2 * 16 + 32 * 8;
      cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      f<<<148, 1024, smem>>>(tab, words - 1, out, iters, cyc); report("mix tma x2 + ldg x6", 8, iters);
    }
    {
      auto f = k_tma<4, 8>; size_t smem = 32 * 32 * 4 * 16 + 32 * 8;
      cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      f<<<148, 1024, smem>>>(tab, words - 1, out, iters, cyc); report("mix tma x4 + ldg x8", 12, iters);
    }
  }
  return 0;
}
