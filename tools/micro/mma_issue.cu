// Microbenchmark (sm_100a): is the ~60 clk per small tcgen05.mma a tensor-pipe limit or an issue-side limit?
// nwarps issuing threads (one per warp) each issue `cnt` MMAs (M=128, N=32, K=16) into their own accumulator, with the
// descriptors precomputed (unrolled by 4).  Prints clk per MMA per issuing thread and aggregate.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../adafortitran_b200/csrc/tc_ptx.cuh"
using namespace aft::ptx;
typedef unsigned long long u64;
constexpr uint32_t kHi128 = (uint32_t)(desc_k_sw128_const() >> 32);
__device__ __forceinline__ uint64_t d128(uint32_t saddr, int ks) {
  return ((uint64_t)kHi128 << 32) | (((uint32_t)desc_k_sw128_const() | ((saddr >> 4) & 0x3FFF)) + (uint32_t)ks * 2);
}
__global__ void __launch_bounds__(128, 1) k(u64* out, int cnt, int nissue, int N) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0;
  const uint32_t bar = sb + 40960, tptr = bar + 64;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(bar + 8 * i, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(tptr, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tptr));
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && w < nissue) {
    const uint32_t idesc = make_idesc_bf16(128, N, false, false);
    const uint64_t a0 = d128(sb, 0), a1 = d128(sb, 1), a2 = d128(sb, 2), a3 = d128(sb, 3);
    const uint64_t b0 = d128(sb + 16384, 0), b1 = d128(sb + 16384, 1), b2 = d128(sb + 16384, 2), b3 = d128(sb + 16384, 3);
    const uint32_t d = tmem + 128 * w;
    for (int rep = 0; rep < 2; ++rep) {
      const u64 t0 = clock64();
      for (int i = 0; i < cnt; i += 4) { mma_ss(d, a0, b0, idesc, true); mma_ss(d, a1, b1, idesc, true); mma_ss(d, a2, b2, idesc, true); mma_ss(d, a3, b3, idesc, true); }
      mma_commit(bar + 8 * w);
      mbar_wait(bar + 8 * w, rep);
      out[w] = clock64() - t0;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}
int main() {
  u64* out; cudaMalloc(&out, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 50 * 1024);
  const int cnt = 256;
  for (int N : {32, 128})
    for (int ni = 1; ni <= 4; ++ni) {
      cudaMemset(out, 0, 64);
      k<<<1, 128, 50 * 1024>>>(out, cnt, ni, N);
      u64 h[4]; if (cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("error\n"); return 1; }
      u64 mx = 0; for (int i = 0; i < ni; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("N=%3d issuing threads %d: %.1f clk per MMA per thread, %.1f clk per MMA aggregate\n", N, ni, (double)mx / cnt, (double)mx / (cnt * ni));
    }
  return 0;
}
