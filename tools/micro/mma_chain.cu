// Microbenchmark (sm_100a): issue-to-completion behaviour of tcgen05.mma chains.  One CTA, one issuing thread.
// For N in {32, 64, 96, 128} (M = 128, K = 16, bf16, SS mode, SWIZZLE_128B K-major operands in zeroed shared memory):
//   dep   : `cnt` MMAs accumulating into the SAME accumulator columns
//   indep : `cnt` MMAs alternating between FOUR disjoint accumulators
//   ts    : dependent chain with the A operand in tensor memory (the P.V form), N = 32
// prints clocks per MMA (clock64 around issue ... commit -> mbarrier wait).
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
__global__ void __launch_bounds__(128, 1) k(u64* out, int cnt) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0;
  const uint32_t bar = sb + 40960, tptr = bar + 16;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(tptr, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tptr));
  if (threadIdx.x == 0) {
    uint32_t par = 0;
    int slot = 0;
    const int Ns[4] = {32, 64, 96, 128};
    for (int ni = 0; ni < 4; ++ni) {
      const int N = Ns[ni];
      const uint32_t idesc = make_idesc_bf16(128, N, false, false);
      for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
          const u64 t0 = clock64();
          for (int i = 0; i < cnt; ++i) mma_ss(tmem + (mode ? (i & 3) * 128 : 0), d128(sb, i & 3), d128(sb + 16384, i & 3), idesc, true);
          mma_commit(bar);
          mbar_wait(bar, par); par ^= 1;
          const u64 t1 = clock64();
          if (rep) out[slot++] = t1 - t0;
        }
      }
    }
    const uint32_t idpv = make_idesc_bf16(128, 32, false, false);
    for (int rep = 0; rep < 2; ++rep) {
      const u64 t0 = clock64();
      for (int i = 0; i < cnt; ++i) mma_ts(tmem, tmem + 256 + (i & 3) * 8, d128(sb + 16384, i & 3), idpv, true);
      mma_commit(bar);
      mbar_wait(bar, par); par ^= 1;
      const u64 t1 = clock64();
      if (rep) out[slot++] = t1 - t0;
    }
    // M = 64 (N = 32, 96), and N = 192 / 256 at M = 128
    {
      const uint32_t ids[4] = {make_idesc_bf16(64, 32, false, false), make_idesc_bf16(64, 96, false, false), make_idesc_bf16(128, 192, false, false),
                               make_idesc_bf16(128, 256, false, false)};
      for (int v = 0; v < 4; ++v)
        for (int rep = 0; rep < 2; ++rep) {
          const u64 t0 = clock64();
          for (int i = 0; i < cnt; ++i) mma_ss(tmem, d128(sb, i & 3), d128(sb + 8192, i & 3), ids[v], true);
          mma_commit(bar);
          mbar_wait(bar, par); par ^= 1;
          const u64 t1 = clock64();
          if (rep) out[slot++] = t1 - t0;
        }
    }
    // latency of a single small MMA group: 2 MMAs + commit + wait (the S tile form), 1 MMA
    for (int g = 1; g <= 8; g *= 2) {
      u64 acc = 0;
      for (int rep = 0; rep < 9; ++rep) {
        const u64 t0 = clock64();
        for (int i = 0; i < g; ++i) mma_ss(tmem, d128(sb, i & 3), d128(sb + 16384, i & 3), make_idesc_bf16(128, 32, false, false), i > 0);
        mma_commit(bar);
        mbar_wait(bar, par); par ^= 1;
        if (rep) acc += clock64() - t0;
      }
      out[slot++] = acc / 8;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}
int main() {
  u64* out; cudaMalloc(&out, 64 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 50 * 1024);
  const int cnt = 256;
  k<<<1, 128, 50 * 1024>>>(out, cnt);
  u64 h[64]; cudaError_t e = cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  const int Ns[4] = {32, 64, 96, 128};
  int s = 0;
  for (int ni = 0; ni < 4; ++ni) {
    printf("N=%3d  dependent chain: %.1f clk/MMA   independent (4 accumulators): %.1f clk/MMA\n", Ns[ni], (double)h[s] / cnt, (double)h[s + 1] / cnt);
    s += 2;
  }
  printf("N= 32  TS (A in TMEM) dependent chain: %.1f clk/MMA\n", (double)h[s++] / cnt);
  const char* nm[4] = {"M=64 N=32", "M=64 N=96", "M=128 N=192", "M=128 N=256"};
  for (int v = 0; v < 4; ++v) printf("%s dependent chain: %.1f clk/MMA\n", nm[v], (double)h[s++] / cnt);
  for (int g = 1; g <= 8; g *= 2) printf("group of %d MMAs (N=32) issue -> commit -> mbarrier wait returns: %llu clk\n", g, h[s++]);
  return 0;
}
