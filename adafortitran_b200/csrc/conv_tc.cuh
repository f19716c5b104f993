// ConvEnhancer (reference src/models/blocks/enhancers.py:12-20) for the AFT_BF16 path: the two wide layers
// (8 -> 32 and 32 -> 8 channels, 92 % of the stack's FLOPs) run on the 5th-gen tensor cores as an implicit GEMM;
// the 1 -> 8 and 8 -> 1 layers stay on CUDA cores.
//
// Implicit GEMM by *shifted descriptors*: an activation map is stored in shared memory as K-major, NON-swizzled
// UMMA operand planes
//        plane[g][p][8 channels] (bf16, 16 bytes per position),   g = channel group of 8,
// over the zero-padded 122 x 16 position grid p = (row+1)*16 + (col+1).  Eight consecutive positions are one UMMA
// core matrix (8 rows x 16 bytes, contiguous), so the A operand "128 output positions starting at p0, as seen through
// filter tap (dy,dx)" is just the same plane addressed at p0 + (dy-1)*16 + (dx-1): one tcgen05.mma per (tap, 16 input
// channels) with a shifted start address -- no im2col copy.  M tiles run over padded positions (16 tiles of 128);
// results at border positions are discarded (written back as the zeros the next layer's padding needs).
//
//   conv2:  A = a1 planes (8 channels + 8 zero channels = K 16), B = [32 cout][16] per tap    ->  9 MMAs / tile, N = 32
//   conv3:  A = mid planes (32 channels = 2 K-steps),            B = [8+8 zero cout][16]       -> 18 MMAs / tile, N = 16
// Accumulators: conv2 fills all 512 TMEM columns (16 tiles x 32), conv3 reuses the first 256.
#pragma once

#include "aft_internal.cuh"
#include "tc_layout.cuh"
#include "tc_ptx.cuh"

namespace aft {
namespace convtc {

using namespace ptx;

constexpr int kThreads = 512;                 // 16 warps; thread 0 also issues the MMAs
constexpr int kPosGuard = 32;                 // positions of slack before p = 0 (taps reach back 17 positions)
constexpr int kPosAlloc = 2112;               // 32 guard + 2048 (16 M-tiles) + 32 guard
constexpr int kPlaneBytes = kPosAlloc * 16;   // one 8-channel group: 33,792 bytes
constexpr int kTiles = 16;

// packed parameters of one conv stack (global memory, built by conv_tc_pack): byte offsets
constexpr int kPkW2 = 0;                      // 9 taps x [2 halves][32 cout][8] bf16 = 9 x 1024
constexpr int kPkW3 = 9216;                   // 18 (tap, kstep) x [2 halves][16 cout][8] bf16 = 18 x 512
constexpr int kPkF32 = 18432;                 // fp32: w0[9][8] | b0[8] | b1[32] | b2[8] | w3[9][8] | b3[1] (+pad) = 200 floats
constexpr int kPkBytes = 18432 + 800;         // 19,232
constexpr int kF_w0 = 0, kF_b0 = 72, kF_b1 = 80, kF_b2 = 112, kF_w3 = 120, kF_b3 = 192;

// shared-memory map of one conv-stack workspace (bytes, relative to a 1024-aligned base)
constexpr int OFF_A1 = 0;                               // 2 groups (second one zero) -- conv3's output a3 reuses group 0
constexpr int OFF_MID = 2 * kPlaneBytes;                // 4 groups
constexpr int OFF_PK = 6 * kPlaneBytes;                 // packed weights (kPkBytes)
constexpr int OFF_IN = OFF_PK + 19456;                  // fp32 padded input plane [122][16] (7,808 B); reused for the fp32 output
constexpr int OFF_BAR = OFF_IN + 7808;                  // 2 mbarriers + TMEM pointer + 24 input floats
constexpr int kStackSmemBytes = OFF_BAR + 256;          // 230,272
// After a stack has run, the mid planes are dead until the next conv2 epilogue rewrites positions [0, 2048) of every
// group; the fp32 result and the callers' scratch live there (behind the 512-byte front guard, which must stay zero).
constexpr int OFF_OUT = OFF_MID + kPosGuard * 16;                   // fp32 result, unpadded [1680]
constexpr int OFF_SCRATCH = OFF_MID + kPlaneBytes + kPosGuard * 16; // 32,768 bytes of caller scratch (group 1)

__device__ __forceinline__ uint64_t desc_k_none(uint32_t saddr, uint32_t lbo_bytes) {
  return make_smem_desc(saddr, lbo_bytes, 128, kSwizzleNone);   // SBO = 128: next 8 rows (positions / couts)
}

constexpr uint32_t kIdescConv2 = make_idesc_bf16(128, 32, false, false);
constexpr uint32_t kIdescConv3 = make_idesc_bf16(128, 16, false, false);

__device__ __forceinline__ bool interior(int p) {   // padded position -> is it a real pixel?
  const int r = p >> 4, c = p & 15;
  return r >= 1 && r <= kGridH && c >= 1 && c <= kGridW;
}

// zero the activation planes once per CTA (guards and borders must stay zero; interiors are rewritten every image)
__device__ __forceinline__ void stack_init(uint8_t* smem, const uint8_t* __restrict__ pack) {
  uint4* z = reinterpret_cast<uint4*>(smem + OFF_A1);
  for (int i = threadIdx.x; i < 6 * kPlaneBytes / 16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
  const uint4* src = reinterpret_cast<const uint4*>(pack);
  uint4* dst = reinterpret_cast<uint4*>(smem + OFF_PK);
  for (int i = threadIdx.x; i < kPkBytes / 16; i += blockDim.x) dst[i] = src[i];
  float* in = reinterpret_cast<float*>(smem + OFF_IN);
  for (int i = threadIdx.x; i < kPlane; i += blockDim.x) in[i] = 0.f;
}

// Runs the stack on the fp32 padded plane at OFF_IN (interior filled by the caller, border zero) and leaves the fp32
// result, unpadded [1680], at OFF_OUT.  All kThreads threads call it.  `bar` = shared address of two mbarriers
// (count 1 each), `phase` = number of stacks this CTA has run before (parity of both barriers).
__device__ __forceinline__ void stack_run(uint8_t* smem, uint32_t sb, uint32_t tmem, uint32_t bar, uint32_t phase) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* fw = reinterpret_cast<const float*>(smem + OFF_PK + kPkF32);
  const float* in = reinterpret_cast<const float*>(smem + OFF_IN);

  // ---- conv1 (1 -> 8, ReLU) on CUDA cores: one interior position per thread iteration -> a1 group 0 (bf16)
  for (int px = tid; px < kPix; px += kThreads) {
    const int r = px / kGridW, c = px - r * kGridW;
    float acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = fw[kF_b0 + o];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float a = in[(r + t / 3) * kPW + c + t % 3];
#pragma unroll
      for (int o = 0; o < 8; ++o) acc[o] = fmaf(a, fw[kF_w0 + t * 8 + o], acc[o]);
    }
    const int p = (r + 1) * kPW + c + 1;
    *reinterpret_cast<uint4*>(smem + OFF_A1 + (kPosGuard + p) * 16) =
        make_uint4(pack_bf16x2(fmaxf(acc[0], 0.f), fmaxf(acc[1], 0.f)), pack_bf16x2(fmaxf(acc[2], 0.f), fmaxf(acc[3], 0.f)),
                   pack_bf16x2(fmaxf(acc[4], 0.f), fmaxf(acc[5], 0.f)), pack_bf16x2(fmaxf(acc[6], 0.f), fmaxf(acc[7], 0.f)));
  }
  fence_proxy_async_smem();
  __syncthreads();

  // ---- conv2 (8 -> 32) on the tensor core: 16 tiles x 9 taps, K = 16 (8 channels + 8 zero channels)
  if (warp == 0) {   // converged warp, elected lane issues (keeps the descriptor math in uniform registers)
    const bool el = elect_one();
    tc_fence_after_sync();
    const uint32_t a1 = sb + OFF_A1 + kPosGuard * 16;
#pragma unroll 1
    for (int i = 0; i < kTiles; ++i)
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int shift = (t / 3 - 1) * kPW + (t % 3 - 1);
        mma_ss(tmem + i * 32, desc_k_none(a1 + (i * 128 + shift) * 16, kPlaneBytes), desc_k_none(sb + OFF_PK + kPkW2 + t * 1024, 512),
               kIdescConv2, t > 0, el);
      }
    mma_commit(bar, el);
  }
  mbar_wait(bar, phase & 1);
  tc_fence_after_sync();

  // ---- conv2 epilogue: + bias, ReLU, zero the border positions -> mid planes (4 groups of 8 channels, bf16)
  {
    const int q = warp & 3, part = warp >> 2;
#pragma unroll 1
    for (int i = part; i < kTiles; i += 4) {
      const int p = i * 128 + q * 32 + lane;
      uint32_t acc[32];
      tmem_ld16p(tmem + ((uint32_t)(q * 32) << 16) + i * 32, acc);
      tmem_ld16p(tmem + ((uint32_t)(q * 32) << 16) + i * 32 + 16, acc + 16);
      tmem_wait_ld();
      const bool in_img = interior(p);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float v0 = fmaxf(__uint_as_float(acc[g * 8 + 2 * j]) + fw[kF_b1 + g * 8 + 2 * j], 0.f);
          const float v1 = fmaxf(__uint_as_float(acc[g * 8 + 2 * j + 1]) + fw[kF_b1 + g * 8 + 2 * j + 1], 0.f);
          pk[j] = in_img ? pack_bf16x2(v0, v1) : 0u;
        }
        *reinterpret_cast<uint4*>(smem + OFF_MID + g * kPlaneBytes + (kPosGuard + p) * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
  }
  tc_fence_before_sync();
  fence_proxy_async_smem();
  __syncthreads();

  // ---- conv3 (32 -> 8) on the tensor core: 16 tiles x 9 taps x 2 K-steps, N = 16 (8 real + 8 zero output channels)
  if (warp == 0) {
    const bool el = elect_one();
    tc_fence_after_sync();
    const uint32_t mid = sb + OFF_MID + kPosGuard * 16;
#pragma unroll 1
    for (int i = 0; i < kTiles; ++i)
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int shift = (t / 3 - 1) * kPW + (t % 3 - 1);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
          mma_ss(tmem + i * 16, desc_k_none(mid + 2 * ks * kPlaneBytes + (i * 128 + shift) * 16, kPlaneBytes),
                 desc_k_none(sb + OFF_PK + kPkW3 + (t * 2 + ks) * 512, 256), kIdescConv3, (t | ks) != 0, el);
      }
    mma_commit(bar + 8, el);
  }
  mbar_wait(bar + 8, phase & 1);
  tc_fence_after_sync();

  // ---- conv3 epilogue: + bias, ReLU, zero borders -> a3 (bf16, reuses the a1 group-0 plane: conv2 has consumed it)
  {
    const int q = warp & 3, part = warp >> 2;
#pragma unroll 1
    for (int i = part; i < kTiles; i += 4) {
      const int p = i * 128 + q * 32 + lane;
      uint32_t acc[8];
      tmem_ld8p(tmem + ((uint32_t)(q * 32) << 16) + i * 16, acc);
      tmem_wait_ld();
      const bool in_img = interior(p);
      uint32_t pk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float v0 = fmaxf(__uint_as_float(acc[2 * j]) + fw[kF_b2 + 2 * j], 0.f);
        const float v1 = fmaxf(__uint_as_float(acc[2 * j + 1]) + fw[kF_b2 + 2 * j + 1], 0.f);
        pk[j] = in_img ? pack_bf16x2(v0, v1) : 0u;
      }
      *reinterpret_cast<uint4*>(smem + OFF_A1 + (kPosGuard + p) * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();

  // ---- conv4 (8 -> 1, no activation) on CUDA cores -> fp32 result, unpadded, at OFF_IN
  for (int px = tid; px < kPix; px += kThreads) {
    const int r = px / kGridW, c = px - r * kGridW;
    float acc = fw[kF_b3];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int p = (r + t / 3) * kPW + c + t % 3;
      const uint4 a = *reinterpret_cast<const uint4*>(smem + OFF_A1 + (kPosGuard + p) * 16);
      const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc = fmaf(__uint_as_float(w[j] << 16), fw[kF_w3 + t * 8 + 2 * j], acc);
        acc = fmaf(__uint_as_float(w[j] & 0xFFFF0000u), fw[kF_w3 + t * 8 + 2 * j + 1], acc);
      }
    }
    reinterpret_cast<float*>(smem + OFF_OUT)[px] = acc;
  }
  __syncthreads();
}

}  // namespace convtc
}  // namespace aft
