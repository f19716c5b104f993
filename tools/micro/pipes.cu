// Microbenchmark (sm_100a): issue rate of the instructions the encoder epilogues are made of, per SM sub-partition.
// One block per launch; nw warps per sub-partition (threads = 128 * nw); every warp runs `iters` x 64 independent
// instructions of one kind (8 accumulator chains).  Prints clocks per warp-instruction per sub-partition.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
template <int KIND>
__global__ void k(float* out, u64* clk, int iters) {
  u64 a[8]; float f[8]; uint32_t h[8];
  for (int i = 0; i < 8; ++i) { f[i] = threadIdx.x * 1e-3f + i; a[i] = ((u64)__float_as_uint(f[i]) << 32) | __float_as_uint(f[i] * 0.5f); h[i] = i; }
  const u64 c1 = 0x3f8000013f800001ull;
  __syncthreads();
  const u64 t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (KIND == 0) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(c1));
        if (KIND == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(a[i]) : "l"(c1));
        if (KIND == 2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
        if (KIND == 3) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(1.0001f));
        if (KIND == 4) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(f[(i + 1) & 7]), "f"(f[(i + 2) & 7]));
        if (KIND == 5) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(f[i]), "f"(f[(i + 1) & 7]));
        if (KIND == 6) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(f[i]));
        if (KIND == 7) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(c1)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i])); }
        if (KIND == 8) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(c1)); asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(f[(i + 1) & 7]), "f"(f[(i + 2) & 7])); }
      }
  }
  const u64 t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += f[i] + __uint_as_float((uint32_t)a[i]) + h[i];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
template <int KIND> void run(const char* name, int per) {
  float* out; u64* clk; cudaMalloc(&out, 4096); cudaMalloc(&clk, 8);
  for (int nw : {1, 2, 4}) {
    const int iters = 200;
    k<KIND><<<1, 128 * nw>>>(out, clk, iters); k<KIND><<<1, 128 * nw>>>(out, clk, iters);
    u64 c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
    printf("%-28s warps/SMSP %d : %.2f clk per warp-instr per SMSP\n", name, nw, (double)c / (iters * 64.0 * per * nw));
  }
}
int main() {
  run<0>("add.f32x2 (FADD2)", 1); run<1>("fma.f32x2 (FFMA2)", 1); run<3>("fma.f32 (FFMA)", 1); run<4>("max3.f32 (FMNMX3)", 1);
  run<5>("cvt.bf16x2 (F2FP)", 1); run<2>("ex2 (MUFU.EX2)", 1); run<6>("tanh (MUFU.TANH)", 1); run<7>("FADD2 + EX2 pair", 2); run<8>("FADD2 + FMNMX3 pair", 2);
  return 0;
}
