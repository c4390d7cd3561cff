// Integer-pipe throughput on sm_100a: IMAD, IMAD.HI, SHF, LOP3 and mixes, per SM sub-partition.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_pipes int_pipes.cu && ./int_pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 512
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(uint32_t *out, uint32_t m, uint32_t s, long long *cyc) {
  uint32_t a[8];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 2654435761u + i;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0 || MODE == 4 || MODE == 6) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(s));
      if (MODE == 1 || MODE == 5) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(s));
      if (MODE == 2 || MODE == 4 || MODE == 5) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(s));
      if (MODE == 3 || MODE == 6) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(m), "r"(s));
      if (MODE == 8) { unsigned long long w; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(a[i]), "r"(m)); a[i] = (uint32_t)w ^ (uint32_t)(w >> 32); }
      if (MODE == 9) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(s));
      if (MODE == 10) asm volatile("{.reg .pred p; setp.ne.u32 p, %2, 0; selp.b32 %0, %0, %1, p;}" : "+r"(a[i]) : "r"(m), "r"(s));
      if (MODE == 11) asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.b32 %0, %0, %2, p;}" : "+r"(a[i]) : "r"(m), "r"(s));
      if (MODE == 12) asm volatile("bfind.u32 %0, %0;" : "+r"(a[i]));
      if (MODE == 13) asm volatile("popc.b32 %0, %0;" : "+r"(a[i]));
      if (MODE == 14) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1);
      if (MODE == 15) a[i] = __ballot_sync(0xffffffffu, a[i] & 1);
      if (MODE == 16) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(s));
      if (MODE == 17) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(s));
                        asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(s));
                        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(m), "r"(s)); }
      if (MODE == 18) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "n"(51712), "r"(s));
      if (MODE == 7) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(s));
                       asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(s));
                       asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(s));
                       asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(s));
                       asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(m), "r"(s)); }
    }
  }
  long long t1 = clock64();
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) r ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE> void run(const char *name, int per_iter) {
  uint32_t *out; long long *cyc, h;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  k<MODE><<<148, 1024>>>(out, 0x9E3779B1u, 7, cyc); cudaDeviceSynchronize();
  k<MODE><<<148, 1024>>>(out, 0x9E3779B1u, 7, cyc); cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  // warp instructions per SMSP = 8 warps * ITER * 8 * per_iter
  double inst = 8.0 * ITER * 8 * per_iter;
  printf("%-28s %8lld cycles  %.3f warp-inst/clk/SMSP  (%.2f clk per inst)\n", name, h, inst / h, h / inst);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("IMAD", 1); run<1>("IMAD.HI", 1); run<2>("SHF", 1); run<3>("LOP3", 1);
  run<4>("IMAD+SHF", 2); run<5>("IMAD.HI+SHF", 2); run<6>("IMAD+LOP3", 2);
  run<7>("IMAD,IMAD.HI,SHF,SHF,LOP3", 5);
  run<8>("IMAD.WIDE (+xor)", 2); run<9>("PRMT", 1); run<10>("SEL", 1); run<11>("ISETP+SEL", 2);
  run<12>("FLO", 1); run<13>("POPC", 1); run<14>("SHFL", 1); run<15>("VOTE(+and)", 2); run<16>("IADD", 1);
  run<17>("IMAD,SHF,LOP3", 3); run<18>("IMAD imm", 1);
  return 0;
}
